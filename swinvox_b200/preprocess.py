"""GPU input pipeline: the reference's evaluation-time transforms (core/test.py:50-55 --
CenterCrop -> RandomBackground(TEST.RANDOM_BG_COLOR_RANGE) -> Normalize -> ToTensor, utils/data_transforms.py) in one
kernel, from the uint8 renderings straight into the encoder's `[B,V,3,224,224]` fp32 input.  The reference does this
per sample on the CPU with np.append in a loop (quadratic in the number of views).  No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def crop_window(h, w, crop_h, crop_w):
    """utils/data_transforms.py:142-151 (no bounding box): centred crop when the image is larger, else the whole image"""
    if h > crop_h and w > crop_w:
        x_left = int(w - crop_w) // 2
        y_top = int(h - crop_h) // 2
        return y_top, int(y_top + crop_h), x_left, int(x_left + crop_w)
    return 0, h, 0, w


def bbox_windows(bounding_box, n_views, h, w):
    """Crop windows (y0, y1, x0, x1) of utils/data_transforms.py:93-128 for the views of ONE sample: the square around the
    normalised bounding box (x0, y0, x1, y1), possibly leaving the image (the kernel clamps = the reference's edge padding).
    Faithful to the reference's loop, which overwrites `bounding_box` with its pixel-scaled value inside the per-view loop
    (:95-100), so from the second view on the box is scaled again; a window that misses the image raises, like the
    reference's np.pad on an empty crop does."""
    bb = [float(v) for v in bounding_box]
    out = []
    for _ in range(n_views):
        bb = [bb[0] * w, bb[1] * h, bb[2] * w, bb[3] * h]
        sq = max(bb[2] - bb[0], bb[3] - bb[1])
        xm, ym = (bb[2] + bb[0]) * .5, (bb[3] + bb[1]) * .5
        x_left, x_right = int(xm - sq * .5), int(xm + sq * .5)
        y_top, y_bottom = int(ym - sq * .5), int(ym + sq * .5)
        if x_left >= w or y_top >= h or x_right < 0 or y_bottom < 0 or x_right < x_left or y_bottom < y_top:
            raise ValueError("bounding-box crop misses the image (the reference's np.pad raises on the empty crop)")
        out.append((y_top, y_bottom + 1, x_left, x_right + 1))
    return out


class EvalTransform:
    def __init__(self, img_size=(224, 224), crop_size=(128, 128), bg_color_range=((240, 240), (240, 240), (240, 240)),
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
        """arguments = cfg.CONST.IMG_H/W, cfg.CONST.CROP_IMG_H/W, cfg.TEST.RANDOM_BG_COLOR_RANGE (or
        cfg.TRAIN.RANDOM_BG_COLOR_RANGE), cfg.DATASET.MEAN/STD.  A degenerate range (lo == hi, the evaluation default) is a
        fixed colour; a proper range draws one colour per sample on the host, exactly like the reference
        (utils/data_transforms.py:433-435: np.random.randint(lo, hi + 1) for b, g, r), and the kernel applies it."""
        self.img_size, self.crop_size = tuple(img_size), tuple(crop_size)
        self.mean, self.std = [float(m) for m in mean], [float(s) for s in std]
        self.bg_range = [(int(lo), int(hi)) for lo, hi in bg_color_range]
        self.random_bg = any(lo != hi for lo, hi in self.bg_range)
        self.bg_norm = self._norm(np.array([lo for lo, _ in self.bg_range], np.float64))

    def _norm(self, bg255):
        return ((bg255 / 255. - np.array(self.mean)) / np.array(self.std)).astype(np.float32)   # float64 math, as numpy does

    def __call__(self, images_u8, out=None, bounding_box=None, rng=None):
        """images_u8: uint8 CUDA tensor [..., H, W, C] (C = 3 or 4, BGR(A) as cv2.imread gives) -> fp32
        [..., 3, img_h, img_w] on the same device (e.g. [B,V,137,137,4] -> [B,V,3,224,224]).
        bounding_box: normalised (x0, y0, x1, y1) of the sample ([..., 4] for a batch: one box per leading index but the
        last, the view axis) -- the Pascal3D / Pix3D path of CenterCrop.  rng: numpy RandomState-like generator for the
        background colours of a proper range (default: numpy's global generator, which is what the reference draws from;
        one draw of three colours per sample, in sample order)."""
        if images_u8.device.type != "cuda" or images_u8.dtype != torch.uint8:
            raise _lib.SvxError("EvalTransform takes uint8 CUDA tensors (no CPU fallback)")
        lib = _lib.get()
        x = images_u8.contiguous()
        lead, (H, W, Cc) = x.shape[:-3], x.shape[-3:]
        N = int(np.prod(lead)) if len(lead) else 1
        OH, OW = self.img_size
        shape = tuple(lead) + (3, OH, OW)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=x.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"EvalTransform: `out` must be a contiguous fp32 tensor of shape {shape}")
        d = _lib.PreprocessDesc()
        d.inp, d.out = x.data_ptr(), out.data_ptr()
        d.N, d.H, d.W, d.C, d.OH, d.OW = N, H, W, Cc, OH, OW
        d.y0, d.y1, d.x0, d.x1 = crop_window(H, W, self.crop_size[0], self.crop_size[1])
        for i in range(3):
            d.mean[i], d.std[i], d.bg_norm[i] = self.mean[i], self.std[i], float(self.bg_norm[i])
        V = lead[-1] if len(lead) else 1          # the last leading axis is the view axis of a sample
        S = N // V
        keep = []
        if bounding_box is not None:
            bb = np.asarray(bounding_box, np.float64).reshape(S, 4)
            wins = np.array([wv for s_ in range(S) for wv in bbox_windows(bb[s_], V, H, W)], np.int32)
            keep.append(torch.from_numpy(wins).to(x.device))
            d.windows = keep[-1].data_ptr()
        if self.random_bg:
            rng = rng or np.random
            cols = np.array([[rng.randint(lo, hi + 1) for lo, hi in self.bg_range] for _ in range(S)], np.float64)
            bgn = np.repeat(np.stack([self._norm(c) for c in cols]), V, axis=0).astype(np.float32)
            keep.append(torch.from_numpy(bgn).to(x.device))
            d.bg_norm_n = keep[-1].data_ptr()
        _lib.check(lib.svx_preprocess(C.byref(d), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
        self._keep = keep    # the launch is asynchronous: the small per-image arrays must outlive it
        return out

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


class EvalTransform:
    def __init__(self, img_size=(224, 224), crop_size=(128, 128), bg_color_range=((240, 240), (240, 240), (240, 240)),
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
        """arguments = cfg.CONST.IMG_H/W, cfg.CONST.CROP_IMG_H/W, cfg.TEST.RANDOM_BG_COLOR_RANGE, cfg.DATASET.MEAN/STD.
        The evaluation range is degenerate (lo == hi): a fixed colour; a proper range is rejected (the reference would
        draw a random colour per sample, which is a training-time augmentation)."""
        if any(lo != hi for lo, hi in bg_color_range):
            raise ValueError("EvalTransform needs a fixed background colour (cfg.TEST.RANDOM_BG_COLOR_RANGE)")
        self.img_size, self.crop_size = tuple(img_size), tuple(crop_size)
        self.mean, self.std = [float(m) for m in mean], [float(s) for s in std]
        bg = np.array([lo for lo, _ in bg_color_range], np.float64) / 255.
        self.bg_norm = ((bg - np.array(self.mean)) / np.array(self.std)).astype(np.float32)   # float64 math, as numpy does

    def __call__(self, images_u8, out=None):
        """images_u8: uint8 CUDA tensor [..., H, W, C] (C = 3 or 4, BGR(A) as cv2.imread gives) -> fp32
        [..., 3, img_h, img_w] on the same device (e.g. [B,V,137,137,4] -> [B,V,3,224,224])"""
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
        _lib.check(lib.svx_preprocess(C.byref(d), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
        return out

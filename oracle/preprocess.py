"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation-time image pipeline
(core/test.py:50-55: CenterCrop -> RandomBackground(TEST.RANDOM_BG_COLOR_RANGE) -> Normalize -> ToTensor, over images
read as `cv2.imread(path, IMREAD_UNCHANGED).astype(np.float32) / 255.`, utils/data_loaders.py:70-76).

The bilinear resize is OpenCV's (cv2.resize, INTER_LINEAR, float32 images: cv2 is a third-party dependency of the
reference, opencv-python 4.x, requirements.txt) restated from its published algorithm: source coordinate
fx = (dx + 0.5) * (src / dst) - 0.5, floor / clamp at the borders, separable interpolation.  Pinned:
oracle/make_golden_preprocess.py runs the REAL utils/data_transforms.py (which calls the real cv2) on seeded images and
checks this file against it (max abs difference <= 5e-7 on outputs in [-1, 1]); the goldens are committed.
"""
import numpy as np


def _axis(src, dst):
    """bilinear index / weight tables for one axis -> (i0, i1, w1) with w1 in double precision"""
    scale = float(src) / float(dst)
    i0 = np.zeros(dst, np.int64)
    w1 = np.zeros(dst, np.float64)
    for d in range(dst):
        fx = (d + 0.5) * scale - 0.5
        sx = int(np.floor(fx))
        fx -= sx
        if sx < 0:
            sx, fx = 0, 0.0
        if sx >= src - 1:
            sx, fx = src - 1, 0.0
        i0[d], w1[d] = sx, fx
    return i0, np.minimum(i0 + 1, src - 1), w1


def resize_linear(img, out_h, out_w):
    """img float32 [H, W, C] -> float32 [out_h, out_w, C].  The opencv-python build the reference runs on dispatches
    float32 INTER_LINEAR to IPP, whose result sits within 1.2e-7 of the exactly-weighted bilinear value (measured,
    oracle/make_golden_preprocess.py): weights and interpolation are therefore evaluated in double here."""
    h, w, _ = img.shape
    x0, x1, wx = _axis(w, out_w)
    y0, y1, wy = _axis(h, out_h)
    i64 = img.astype(np.float64)
    rows = i64[:, x0, :] * (1 - wx)[None, :, None] + i64[:, x1, :] * wx[None, :, None]
    return (rows[y0] * (1 - wy)[:, None, None] + rows[y1] * wy[:, None, None]).astype(np.float32)


def crop_window(h, w, crop_h, crop_w):
    """utils/data_transforms.py:142-151 -> (y_top, y_bottom, x_left, x_right)"""
    if h > crop_h and w > crop_w:
        x_left = int(w - crop_w) // 2
        y_top = int(h - crop_h) // 2
        return y_top, int(y_top + crop_h), x_left, int(x_left + crop_w)
    return 0, h, 0, w


def bbox_windows(bounding_box, shapes):
    """utils/data_transforms.py:93-128: square crop windows (y_top, y_bottom + 1, x_left, x_right + 1) around the
    bounding box, possibly outside the image (the reference pads with np.pad(mode='edge'), i.e. coordinates clamp).
    Faithful to the reference's loop, which re-assigns `bounding_box` to its pixel-scaled value inside the per-view
    loop (:95-100): from the second view on the already scaled box is scaled again."""
    out = []
    for (h, w) in shapes:
        bounding_box = [bounding_box[0] * w, bounding_box[1] * h, bounding_box[2] * w, bounding_box[3] * h]
        bw, bh = bounding_box[2] - bounding_box[0], bounding_box[3] - bounding_box[1]
        xm, ym = (bounding_box[2] + bounding_box[0]) * .5, (bounding_box[3] + bounding_box[1]) * .5
        sq = max(bw, bh)
        x_left, x_right = int(xm - sq * .5), int(xm + sq * .5)
        y_top, y_bottom = int(ym - sq * .5), int(ym + sq * .5)
        out.append((y_top, y_bottom + 1, x_left, x_right + 1))
    return out


def eval_transform(images_u8, img_size=(224, 224), crop_size=(128, 128), bg=(240, 240, 240), mean=(0.5, 0.5, 0.5),
                   std=(0.5, 0.5, 0.5), bounding_box=None, bg_range=None, rng=None):
    """images_u8: uint8 [V, H, W, C] (C = 4: BGRA with alpha, or 3) -> float32 [V, 3, img_h, img_w].
    bounding_box: normalised (x0, y0, x1, y1) as the Pascal3D / Pix3D loaders pass it (utils/data_loaders.py).
    bg_range: ((lo, hi),) * 3 -- one colour per call drawn like utils/data_transforms.py:433-435 from `rng`
    (default: numpy's global generator, as the reference uses)."""
    out = []
    if bg_range is not None:
        rng = rng or np.random
        bg = [rng.randint(bg_range[i][0], bg_range[i][1] + 1) for i in range(3)]
    bgc = np.array(bg, np.float64) / 255.
    wins = bbox_windows(list(bounding_box), [u8.shape[:2] for u8 in images_u8]) if bounding_box is not None else None
    for vi, u8 in enumerate(images_u8):
        img = u8.astype(np.float32) / 255.
        h, w = img.shape[:2]
        if wins is not None:
            y0, y1, x0, x1 = wins[vi]
            ys = np.clip(np.arange(y0, y1), 0, h - 1)
            xs = np.clip(np.arange(x0, x1), 0, w - 1)
            crop = img[ys][:, xs]
        else:
            y0, y1, x0, x1 = crop_window(h, w, crop_size[0], crop_size[1])
            crop = img[y0:y1, x0:x1]
        img = resize_linear(crop, img_size[0], img_size[1]).astype(np.float64)   # np.append upcasts
        if img.shape[2] == 4:
            alpha = (img[:, :, 3:4] == 0).astype(np.float32)
            img = alpha * bgc[None, None, :] + (1 - alpha) * img[:, :, :3]
        img = (img - np.array(mean)) / np.array(std)
        out.append(np.transpose(img, (2, 0, 1)))
    return np.stack(out).astype(np.float32)

"""Pins oracle/preprocess.py against the real reference transforms (utils/data_transforms.py + the real cv2) and writes
tests/golden/preprocess_cases.npz.  Build container only:  python -m oracle.make_golden_preprocess"""
import os
import sys

import numpy as np

from . import preprocess as OP

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "preprocess_cases.npz")


def renderings(seed, v, h, w, c):
    """ShapeNet-like renderings: an opaque blob on a transparent background, plus a soft alpha edge"""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(v, h, w, c), dtype=np.uint8)
    if c == 4:
        yy, xx = np.mgrid[0:h, 0:w]
        for i in range(v):
            r = min(h, w) * (0.25 + 0.05 * i)
            dist = np.sqrt((yy - h / 2 - 3 * i) ** 2 + (xx - w / 2 + 2 * i) ** 2)
            a = np.clip((r - dist) * 40 + 128, 0, 255).astype(np.uint8)
            img[i, :, :, 3] = a
            img[i][a == 0, :3] = 0
    return img


def main():
    sys.path.insert(0, "/root/reference")
    import utils.data_transforms as T
    store = {}
    cases = {"shapenet137": (1, 2, 137, 137, 4), "small100": (2, 1, 100, 120, 4), "rgb256": (3, 1, 256, 200, 3),
             "exact224": (4, 1, 224, 224, 4)}
    worst = 0.0
    for name, (seed, v, h, w, c) in cases.items():
        u8 = renderings(seed, v, h, w, c)
        tf = T.Compose([T.CenterCrop((224, 224), (128, 128)), T.RandomBackground([[240, 240], [240, 240], [240, 240]]),
                        T.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5]), T.ToTensor()])
        ref = tf(u8.astype(np.float32) / 255.).numpy()
        mine = OP.eval_transform(u8)
        assert ref.shape == mine.shape == (v, 3, 224, 224) and ref.dtype == np.float32
        diff = float(np.abs(ref - mine).max())
        worst = max(worst, diff)
        assert diff <= 5e-7, (name, diff)
        store[f"{name}.input"] = u8
        store[f"{name}.output"] = ref
    # bounding-box crops (Pascal3D / Pix3D loaders pass a normalised box; utils/data_transforms.py:93-131) and a proper
    # background-colour range (one colour per call from numpy's global generator, :433-435)
    extra = {"bbox_rgb": (5, 1, 180, 240, 3, [0.2, 0.1, 0.7, 0.95], None),   # (a second view would crash the reference: see bbox_windows)
             "bbox_edge_rgba": (6, 1, 137, 137, 4, [0.55, 0.4, 1.0, 1.0], None),
             "bg_range": (7, 2, 137, 137, 4, None, [[225, 255], [225, 255], [225, 255]])}
    for name, (seed, v, h, w, c, bbox, bgr) in extra.items():
        u8 = renderings(seed, v, h, w, c)
        crop = T.CenterCrop((224, 224), (128, 128))
        rest = T.Compose([T.RandomBackground(bgr or [[240, 240], [240, 240], [240, 240]]),
                          T.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5]), T.ToTensor()])
        np.random.seed(100 + seed)
        ref = rest(crop(u8.astype(np.float32) / 255., bbox)).numpy()
        np.random.seed(100 + seed)
        mine = OP.eval_transform(u8, bounding_box=bbox, bg_range=bgr)
        diff = float(np.abs(ref - mine).max())
        worst = max(worst, diff)
        assert ref.shape == mine.shape and diff <= 5e-7, (name, diff)
        store[f"{name}.input"] = u8
        store[f"{name}.output"] = ref
        store[f"{name}.bbox"] = np.array(bbox if bbox is not None else [], np.float64)
        store[f"{name}.bg_range"] = np.array(bgr if bgr is not None else [], np.int64)
        store[f"{name}.seed"] = np.array(100 + seed)
    np.savez_compressed(GOLDEN, **store)
    print(f"oracle.preprocess == reference transforms (max abs diff {worst:.2e}) on {len(cases) + len(extra)} cases; wrote {GOLDEN}")


if __name__ == "__main__":
    main()

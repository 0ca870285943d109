"""GPU input pipeline.  CPU tier: the oracle against the outputs of the REAL reference transforms
(tests/golden/preprocess_cases.npz, written by oracle/make_golden_preprocess.py).  GPU tier: the CUDA kernel through
the C-ABI against the same goldens and the oracle.  Tolerance (floating point): 1e-6 absolute on outputs in [-1, 1]
(bilinear weights and interpolation in fp32 on the device vs the reference's cv2/IPP float path), background pixels
and the alpha == 0 decision exact."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as OP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_cases.npz")
CASES = ["shapenet137", "small100", "rgb256", "exact224"]
EXTRA = ["bbox_rgb", "bbox_edge_rgba", "bg_range"]    # bounding-box crops (edge padded) and a proper background range
TOL = 1e-6


def load():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_transforms(name):
    g = load()
    out = OP.eval_transform(g[f"{name}.input"])
    assert out.shape == g[f"{name}.output"].shape
    assert np.abs(out - g[f"{name}.output"]).max() <= 5e-7


@pytest.mark.parametrize("name", EXTRA)
def test_oracle_matches_reference_bbox_and_background_range(name):
    g = load()
    bbox = g[f"{name}.bbox"].tolist() or None
    bgr = g[f"{name}.bg_range"].tolist() or None
    np.random.seed(int(g[f"{name}.seed"]))
    out = OP.eval_transform(g[f"{name}.input"], bounding_box=bbox, bg_range=bgr)
    assert np.abs(out - g[f"{name}.output"]).max() <= 5e-7


def test_bbox_windows_host_rule():
    from swinvox_b200.preprocess import bbox_windows
    assert bbox_windows([0.2, 0.1, 0.7, 0.95], 1, 180, 240) == OP.bbox_windows([0.2, 0.1, 0.7, 0.95], [(180, 240)])
    assert bbox_windows([0.55, 0.4, 1.0, 1.0], 1, 137, 137) == OP.bbox_windows([0.55, 0.4, 1.0, 1.0], [(137, 137)])
    with pytest.raises(ValueError):      # the reference re-scales the box per view: a second view misses the image
        bbox_windows([0.2, 0.1, 0.7, 0.95], 2, 180, 240)


def test_crop_window_rule():
    from swinvox_b200.preprocess import crop_window
    assert crop_window(137, 137, 128, 128) == (4, 132, 4, 132) == OP.crop_window(137, 137, 128, 128)
    assert crop_window(100, 120, 128, 128) == (0, 100, 0, 120)          # smaller than the crop: whole image
    assert crop_window(128, 200, 128, 128) == (0, 128, 0, 200)          # not strictly larger in both: whole image


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_transforms(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200.preprocess import EvalTransform
    g = load()
    u8 = torch.from_numpy(g[f"{name}.input"]).cuda()
    out = EvalTransform()(u8).cpu().numpy()
    ref = g[f"{name}.output"]
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.abs(out - ref).max() <= TOL
    bg = np.float32((240 / 255. - 0.5) / 0.5)
    assert np.array_equal(out == bg, ref == bg)       # the alpha == 0 decision is exact


@pytest.mark.gpu
def test_kernel_batched_into_encoder_input():
    """[B, V, H, W, 4] uint8 -> the encoder's [B, V, 3, 224, 224] buffer in one launch, equal to the oracle per view"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200.preprocess import EvalTransform
    rng = np.random.default_rng(7)
    u8 = rng.integers(0, 256, size=(2, 3, 137, 137, 4), dtype=np.uint8)
    u8[..., 3] = np.where(rng.random((2, 3, 137, 137)) < 0.4, 0, u8[..., 3])
    buf = torch.empty(2, 3, 3, 224, 224, device="cuda")
    got = EvalTransform()(torch.from_numpy(u8).cuda(), out=buf)
    assert got.data_ptr() == buf.data_ptr()
    want = np.stack([OP.eval_transform(u8[b]) for b in range(2)])
    assert np.abs(got.cpu().numpy() - want).max() <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXTRA)
def test_kernel_bbox_and_background_range(name):
    """Pascal3D / Pix3D style bounding-box crops (windows that leave the image read the nearest edge pixel) and
    RandomBackground with a proper colour range (one colour per sample from numpy's generator, like the reference)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from swinvox_b200.preprocess import EvalTransform
    g = load()
    bbox = g[f"{name}.bbox"].tolist() or None
    bgr = g[f"{name}.bg_range"].tolist() or [[240, 240]] * 3
    np.random.seed(int(g[f"{name}.seed"]))
    out = EvalTransform(bg_color_range=bgr)(torch.from_numpy(g[f"{name}.input"]).cuda(), bounding_box=bbox).cpu().numpy()
    assert np.abs(out - g[f"{name}.output"]).max() <= TOL

"""CPU tier: the oracle against the golden vectors the REAL reference produced (oracle/make_golden.py ran
/root/reference/models/*.py in the build container), plus its independent cross-check of the restated timm
Swin-T against torchvision.swin_t.  /root/reference is not needed to run these."""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import modules as M
from oracle import swin_t
from util import golden

CFGS = {
    "default": (dict(), 1, 2),
    "single_stage_nocva": (dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3], USE_CROSS_VIEW_ATTENTION=False), 1, 1),
}


def close(a, b, tol=2e-5):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() <= tol


@pytest.mark.parametrize("tag", list(CFGS))
def test_oracle_reproduces_reference_golden(tag):
    over, B, V = CFGS[tag]
    g = golden(tag)
    cfg = M.default_cfg(**over)
    mods = FX.build(cfg, "calibrated", 0)
    images, gt = FX.structured_inputs(B, V, seed=1234), FX.seeded_gt(B)
    with torch.no_grad():
        taps = {}
        f = mods["encoder"](images, taps)
        raw, gen = mods["decoder"](f)
        m = mods["merger"](raw, gen, taps)
        v = mods["refiner"](m, taps)
    # bit-exact in the container that wrote the fixtures; 2e-5 leaves room for another CPU's SIMD kernels
    assert close(f, g["encoder"]) and close(m, g["merged"]) and close(v, g["final"])
    assert close(gen[:, :, ::2, ::2, ::2], g["gen"]) and close(raw[:, :, 3, ::2, ::2, ::2], g["raw_c3"])
    assert close(taps["resnet"][:, ::8], g["resnet"]) and close(taps["post_cva"][:, :, ::16], g["post_cva"])
    counts, iou, f1 = M.voxel_metrics(v, gt)
    assert (counts.numpy() - g["counts"]).__abs__().max() <= 2
    assert np.abs(iou.numpy() - g["iou"]).max() < 1e-4 and np.abs(f1.numpy() - g["f1"]).max() < 1e-4


def test_swin_restatement_matches_torchvision():
    import torchvision
    tv = torchvision.models.swin_t(weights=None).eval()
    mine = swin_t.load_from_torchvision(swin_t.SwinTFeatures((0, 1, 2, 3)).eval(), tv)
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        outs = mine(x)
        f = tv.features[0](x)
        for s in range(4):
            if s:
                f = tv.features[2 * s](f)
            f = tv.features[2 * s + 1](f)
            assert torch.equal(outs[s], f)


def test_relative_position_index_formula():
    import torchvision
    tv = torchvision.models.swin_t(weights=None)
    assert torch.equal(swin_t.relative_position_index().flatten(), tv.features[1][0].attn.relative_position_index)


def test_feature_list_prunes_trailing_stages():
    m = swin_t.create_model("swin_tiny_patch4_window7_224", pretrained=False, features_only=True, out_indices=[1])
    assert hasattr(m, "layers_1") and not hasattr(m, "layers_2")
    assert m.feature_info.channels() == [192]


def test_parameter_counts_match_reference_logs():
    """Notebook cell 47/53: Decoder 3,817,944 / Refiner 34,880,352 / Merger 17,877; Encoder 40,339,770 for the
    logged single-stage config with CVA, 45,109,818 for the default config."""
    cfg = M.default_cfg()
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(M.RefDecoder(cfg)) == 3817944 and n(M.RefRefiner(cfg)) == 34880352 and n(M.RefMerger(cfg)) == 17877
    assert n(M.RefEncoder(cfg)) == 45109818
    assert n(M.RefEncoder(M.default_cfg(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3]))) == 40339770


def test_metrics_edge_cases():
    """core/test.py:150-153: union == 0 and intersection == 0 -> IoU 1.0"""
    logits = torch.full((2, 32, 32, 32), -30.0)
    gt = torch.zeros(2, 32, 32, 32)
    gt[1, 0, 0, 0] = 1
    counts, iou, f1 = M.voxel_metrics(logits, gt)
    assert torch.all(iou[0] == 1.0) and torch.all(iou[1] == 0.0)
    assert counts[1, :, 4].tolist() == [1, 1, 1, 1] and torch.all(f1 == 0)

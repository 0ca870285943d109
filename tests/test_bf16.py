"""BASELINE configs[2]: the forward with Cross-View Attention in bf16 -- the encoder (ResNet-50 trunk, Swin-T, per-stage
reduce / downsample chains, CVA, fusion layers) stores its activations and feeds its tensor cores in bf16, with fp32
accumulation, LayerNorm / softmax statistics and fp32 module input / output; decoder, merger and refiner stay fp32/TF32.
The reference has no bf16 evaluation path (its only reduced-precision mode is training autocast, core/train.py:235), so
the bar is SURVEY 8(d)'s: stage outputs within rtol 2e-2 of the fp32 oracle, voxel / IoU deltas reported."""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import modules as M
from swinvox_b200.models import Decoder, Encoder, Merger, Refiner
from swinvox_b200.pipeline import Reconstructor
from test_modules import nchw, oracle_forward
from util import dev, parity_log, stage_check, sync  # noqa: F401

PRODUCT = dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner)
RTOL_BF16 = 2e-2


def _run(dev, B, V, seed):
    cfg = M.default_cfg()
    images, gt = FX.structured_inputs(B, V, seed=seed), FX.seeded_gt(B)
    ref = oracle_forward(FX.build(cfg, "calibrated", 0), images, cfg)
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    rec = Reconstructor(cfg, prod["encoder"], prod["decoder"], prod["merger"], prod["refiner"], device=dev, dtype="bf16")
    with torch.no_grad():
        f = rec.encoder(images.to(dev))
        raw, gen = rec.decoder(f)
        merged = rec.merger(raw, gen)
        final = rec.refiner(merged)
    sync(dev)
    assert f.dtype == torch.float32 and final.dtype == torch.float32
    plan = next(iter(rec.encoder._plans.values()))[0]
    assert plan.dtype == torch.bfloat16 and plan.taps["resnet"].buf.dtype == torch.bfloat16
    reports = [stage_check("resnet branch (bf16)", nchw(plan.taps["resnet"]), ref["resnet"], RTOL_BF16)]
    for i, (a, b) in enumerate(zip(plan.taps["swin"], ref["swin"])):
        reports.append(stage_check(f"swin stage {i} (bf16)", nchw(a), b, RTOL_BF16))
    reports.append(stage_check("post_cva (bf16)", nchw(plan.taps["post_cva"]).reshape(ref["post_cva"].shape), ref["post_cva"], RTOL_BF16))
    reports.append(stage_check("encoder (bf16)", f, ref["encoder"], RTOL_BF16))
    reports.append(stage_check("decoder.gen", gen, ref["gen"], RTOL_BF16))
    reports.append(stage_check("merger", merged, ref["merged"], RTOL_BF16))
    reports.append(stage_check("refiner / final logits", final, ref["final"], RTOL_BF16))
    # thresholded voxels and IoU against the fp32 oracle: reported, with a sanity bound
    _, iou_g, _ = M.voxel_metrics(final.float().cpu(), gt)
    _, iou_r, _ = M.voxel_metrics(ref["final"], gt)
    vox = []
    for ti, th in enumerate(cfg.TEST.VOXEL_THRESH):
        mism = ((torch.sigmoid(final.float().cpu()) >= th) != (torch.sigmoid(ref["final"]) >= th)).float().mean().item()
        d_iou = (iou_g[:, ti] - iou_r[:, ti]).abs().max().item()
        vox.append((th, mism, float("nan"), d_iou))
        assert mism <= 5e-3 and d_iou <= 5e-3, (th, mism, d_iou)
    print("\n".join(reports), vox)
    parity_log(f"default, encoder in bf16 [{dev}] B={B} V={V}", reports, vox)


def test_bf16_encoder_pipeline_parity(dev):
    _run(dev, 1, 2, 1234)


@pytest.mark.gpu
def test_bf16_five_views_with_cva_gpu():
    """the view count of BASELINE configs[2]"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run("cuda", 2, 5, 505)

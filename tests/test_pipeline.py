"""The reference's evaluation chain (core/test.py:82-89,120-130) through swinvox_b200.pipeline.Reconstructor: checkpoint
loading, the merger / refiner switches and epoch gates, output lifetime, and oracle parity at the benchmark's full size."""
import collections

import pytest
import torch

from oracle import fixtures as FX
from oracle import modules as M
from swinvox_b200.models import Decoder, Encoder, Merger, Refiner
from swinvox_b200.pipeline import Reconstructor, adapt_state_dict
from util import RTOL, RTOL_DEEP, dev, parity_log, stage_check, sync, voxel_check  # noqa: F401

PRODUCT = dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner)
SMALL = dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3], USE_CROSS_VIEW_ATTENTION=False)


def _mock_checkpoint(ora, epoch_idx=250, drop=()):
    """what core/train.py:358-369 writes: the state dicts of the DataParallel-wrapped modules (`module.` prefix)"""
    ck = collections.OrderedDict(epoch_idx=epoch_idx, best_iou=0.5, best_epoch=epoch_idx)
    for k in ("encoder", "decoder", "refiner", "merger"):
        if k not in drop:
            ck[f"{k}_state_dict"] = collections.OrderedDict(("module." + n, v.clone()) for n, v in ora[k].state_dict().items())
    return ck


def test_load_checkpoint_module_prefix_and_epoch_gates(dev):
    cfg = M.default_cfg(**SMALL)
    cfg.TRAIN = M.AttrDict(EPOCH_START_USE_MERGER=10, EPOCH_START_USE_REFINER=20)
    ora = FX.build(cfg, "calibrated", 0)
    images = FX.structured_inputs(1, 2, seed=11)
    with torch.no_grad():
        raw, gen = ora["decoder"](ora["encoder"](images))
        want = {250: ora["refiner"](ora["merger"](raw, gen)),     # both gates passed
                15: ora["merger"](raw, gen),                      # merger on, refiner not yet (core/test.py:129)
                5: gen.mean(1)}                                   # neither (core/test.py:125-126)
    torch.manual_seed(123)   # the product starts from DIFFERENT random weights: everything must come from the checkpoint
    rec = Reconstructor(cfg, device=dev)
    for epoch, ref in want.items():
        rec.load_checkpoint(_mock_checkpoint(ora, epoch))
        assert rec.epoch_idx == epoch
        got = rec.forward(images.to(dev))
        sync(dev)
        stage_check(f"checkpoint at epoch {epoch}", got, ref, RTOL_DEEP)
    # a configuration that uses the merger / refiner needs their state dicts, like core/test.py:86-89
    with pytest.raises(KeyError):
        rec.load_checkpoint(_mock_checkpoint(ora, drop=("merger",)))
    with pytest.raises(KeyError):
        rec.load_checkpoint(_mock_checkpoint(ora, drop=("refiner",)))
    # ... and ignores them when the configuration does not (Pix2Vox-F style checkpoints have no refiner, notebook cell 66)
    cfg2 = M.default_cfg(USE_REFINER=False, **SMALL)
    Reconstructor(cfg2, device=dev).load_checkpoint(_mock_checkpoint(FX.build(cfg2, "analytic", 0), drop=("refiner",)))


def test_adapt_state_dict_keys():
    sd = {"module.swin_transformer.layer_norm.weight": 1, "module.swin_transformer.layer_norm.bias": 2,
          "module.swin_transformer.layer_norm.1.weight": 3, "resnet.0.weight": 4}
    out = adapt_state_dict(sd)
    assert out == {"swin_transformer.layer_norm.0.weight": 1, "swin_transformer.layer_norm.0.bias": 2,
                   "swin_transformer.layer_norm.1.weight": 3, "resnet.0.weight": 4}
    # checkpoints of another architecture revision fail loudly, naming the keys
    cfg = M.default_cfg(**SMALL)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        Decoder(cfg).load_state_dict(adapt_state_dict({"module.layer1.2.conv.0.weight": torch.zeros(1)}))


def test_outputs_are_fresh_tensors(dev):
    """the reference's modules return fresh tensors; Reconstructor does too unless zero_copy is requested"""
    cfg = M.default_cfg(**SMALL)
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    rec = Reconstructor(cfg, prod["encoder"], prod["decoder"], prod["merger"], prod["refiner"], device=dev)
    a_img, b_img = FX.structured_inputs(1, 1, seed=1).to(dev), FX.structured_inputs(1, 1, seed=2).to(dev)
    gt = FX.seeded_gt(1).to(dev)
    a, ca = rec.evaluate(a_img, gt)
    keep_a, keep_ca = a.clone(), ca.clone()
    b, cb = rec.evaluate(b_img, gt)
    sync(dev)
    assert torch.equal(a, keep_a) and torch.equal(ca, keep_ca) and not torch.equal(a, b)
    rec.zero_copy = True
    z1 = rec.forward(a_img)
    z2 = rec.forward(b_img)
    sync(dev)
    assert z1.data_ptr() == z2.data_ptr()


@pytest.mark.gpu
def test_full_size_oracle_parity_gpu():
    """BASELINE configs[1] at FULL size -- 64 objects x 3 views in one batch -- against the oracle (run in chunks of 8
    objects on the host cores): logits within rtol 1e-3, thresholded voxels / IoU within 1e-4."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = M.default_cfg()
    B, V = 64, 3
    images, gt = FX.structured_inputs(B, V, seed=2024), FX.seeded_gt(B)
    ora = FX.build(cfg, "calibrated", 0)
    with torch.no_grad():
        enc = torch.cat([ora["encoder"](images[i:i + 8]) for i in range(0, B, 8)])
        ref = torch.cat([M.forward_pipeline(ora["encoder"], ora["decoder"], ora["merger"], ora["refiner"], images[i:i + 8], cfg)
                         for i in range(0, B, 8)])
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    rec = Reconstructor(cfg, prod["encoder"], prod["decoder"], prod["merger"], prod["refiner"], device="cuda:0")
    rec.set_graph(True)
    for _ in range(2):   # second call = CUDA-graph replay of the cached plans
        logits, counts = rec.evaluate(images.cuda(), gt.cuda())
    with torch.no_grad():
        f = rec.encoder(images.cuda()).clone()
    torch.cuda.synchronize()
    reports = [stage_check("encoder B64xV3", f, enc, RTOL_DEEP), stage_check("logits B64xV3", logits, ref, RTOL)]
    # mean IoU within 1e-4; per object, one flipped in-band voxel of a ~3500-voxel union already moves IoU by ~1e-4:
    # the worst of 64 objects is held to 3e-4 (every flipped voxel lies inside the error band, checked by voxel_check)
    vox = voxel_check(logits, ref, gt, per_object_iou=3e-4)
    ref_counts, _, _ = M.voxel_metrics(ref, gt)
    dcount = (counts.cpu().long() - ref_counts).abs()
    reports.append(f"counters vs oracle: max |delta| {int(dcount.max())} of 32768 voxels, "
                   f"objects with any delta {int((dcount.flatten(1).max(1).values > 0).sum())}/{B}")
    assert dcount.max().item() <= 32768 * 5e-4
    print("\n".join(reports), vox)
    parity_log("default [cuda] B=64 V=3 (full size)", reports, vox)

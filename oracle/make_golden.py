"""Pins the oracle and writes the golden fixtures.  Runs ONLY in the build container (needs /root/reference):

    python -m oracle.make_golden

1. imports the real reference modules (/root/reference/models/*.py, unmodified) behind two offline shims --
   `timm.create_model` -> oracle.swin_t.create_model (timm is not installed) and
   torchvision.models.resnet50(weights=...) -> weights=None (no network);
2. checks oracle.swin_t against torchvision.models.swin_t (independent implementation, bit-exact);
3. checks every oracle module against the reference module on identical weights (bit-exact on CPU);
4. computes the LSUV calibration scalars of the `calibrated` regime and stores them;
5. runs the REFERENCE forward on seeded structured inputs and stores its outputs as tests/golden/*.npz.
"""
import json
import os
import sys
import types

import numpy as np
import torch

from . import fixtures as FX
from . import modules as M
from . import swin_t

REFERENCE = "/root/reference"

CONFIGS = {
    # tag: (network overrides, B, V)
    "default": (dict(), 1, 2),
    "single_stage_nocva": (dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3], USE_CROSS_VIEW_ATTENTION=False), 1, 1),
    # every other NETWORK switch the hot path reads (config.py:83-94; encoder.py:41-85, decoder.py:24-46,
    # core/test.py:123-130), one at a time.  These store the module outputs only (SLIM).
    "stages_3": (dict(SWIN_T_STAGES=[3]), 1, 2),
    "stages_23": (dict(SWIN_T_STAGES=[2, 3]), 1, 2),
    "stages_13": (dict(SWIN_T_STAGES=[1, 3]), 1, 2),
    "tconv_bias_ratio1": (dict(TCONV_USE_BIAS=True, ATT_SPATIAL_DOWNSAMPLE_RATIO=1), 1, 2),
    "nocva": (dict(USE_CROSS_VIEW_ATTENTION=False), 1, 2),
    "nomerger": (dict(USE_MERGER=False), 1, 2),
    "norefiner": (dict(USE_REFINER=False), 1, 2),
}
FULL = ("default", "single_stage_nocva")   # fixtures that also carry the internal taps


def import_reference():
    import torchvision
    shim = types.ModuleType("timm")
    shim.create_model = swin_t.create_model
    sys.modules["timm"] = shim
    orig = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda weights=None, **k: orig(weights=None)
    sys.path.insert(0, REFERENCE)
    from models.decoder import Decoder
    from models.encoder import Encoder
    from models.merger import Merger
    from models.refiner import Refiner
    return dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner)


def check_swin_vs_torchvision():
    import torchvision
    tv = torchvision.models.swin_t(weights=None).eval()
    mine = swin_t.load_from_torchvision(swin_t.SwinTFeatures((0, 1, 2, 3)).eval(), tv)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        outs = mine(x)
        f = tv.features[0](x)
        for s in range(4):
            if s:
                f = tv.features[2 * s](f)
            f = tv.features[2 * s + 1](f)
            assert torch.equal(outs[s], f), f"swin stage {s} differs from torchvision"
    print("oracle.swin_t == torchvision.swin_t (bit-exact, 4 stages)")


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(FX.GOLDEN_DIR, exist_ok=True)
    ref_factory = import_reference()
    check_swin_vs_torchvision()
    only = sys.argv[1:]
    for tag, (over, B, V) in CONFIGS.items():
        if only and tag not in only:
            continue
        cfg = M.default_cfg(**over)
        # calibration scalars (oracle forward), stored next to the goldens
        mods = FX.build(cfg, "analytic", 0)
        calib = FX.calibrate(mods, cfg, FX.structured_inputs(2, max(V, 2), seed=7))
        with open(FX.calibration_path(cfg, 0), "w") as fh:
            json.dump(calib, fh, indent=0)
        ora = FX.build(cfg, "calibrated", 0)
        ref = FX.build(cfg, "calibrated", 0, ref_factory)
        for k in ora:
            a, b = ora[k].state_dict(), ref[k].state_dict()
            assert list(a) == list(b), f"{k}: state_dict keys differ from the reference"
            assert all(torch.equal(a[n], b[n]) for n in a)
        images = FX.structured_inputs(B, V, seed=1234)
        gt = FX.seeded_gt(B)
        net = cfg.NETWORK
        with torch.no_grad():   # the reference's gating, core/test.py:120-130 (epoch gates passed)
            fr = ref["encoder"](images)
            rawr, genr = ref["decoder"](fr)
            mr = ref["merger"](rawr, genr) if net.USE_MERGER else torch.mean(genr, dim=1)
            vr = ref["refiner"](mr) if net.USE_REFINER else mr
            taps = {}
            fo = ora["encoder"](images, taps)
            rawo, geno = ora["decoder"](fo)
            mo = ora["merger"](rawo, geno, taps) if net.USE_MERGER else torch.mean(geno, dim=1)
            vo = ora["refiner"](mo, taps) if net.USE_REFINER else mo
            assert torch.equal(vo, M.forward_pipeline(ora["encoder"], ora["decoder"], ora["merger"], ora["refiner"], images, cfg))
        for name, a, b in (("encoder", fr, fo), ("raw", rawr, rawo), ("gen", genr, geno), ("merged", mr, mo), ("final", vr, vo)):
            assert torch.equal(a, b), f"{tag}: oracle {name} differs from the reference"
        counts, iou, f1 = M.voxel_metrics(vr, gt)
        # the reference's own metric arithmetic (core/test.py:141-164) on the same tensors
        prob = torch.sigmoid(vr)
        for ti, th in enumerate(cfg.TEST.VOXEL_THRESH):
            _v = torch.ge(prob, th).float()
            inter = torch.sum(_v.mul(gt)).float()
            union = torch.sum(torch.ge(_v.add(gt), 1)).float()
            ref_iou = 1.0 if union.item() == 0 and inter.item() == 0 else (inter / union).item() if union.item() > 0 else 0.0
            assert abs(ref_iou - iou[0, ti].item()) < 1e-7
        if tag not in FULL:
            np.savez_compressed(os.path.join(FX.GOLDEN_DIR, f"golden_{tag}.npz"), encoder=fr.numpy(), merged=mr.numpy(),
                                final=vr.numpy(), gen=genr.numpy()[:, :, ::2, ::2, ::2], counts=counts.numpy(),
                                iou=iou.numpy(), f1=f1.numpy(), B=B, V=V)
            print(f"{tag}: reference == oracle (bit-exact); golden written; IoU {iou.tolist()}")
            continue
        sw = taps["swin"] if isinstance(taps["swin"], list) else [taps["swin"]]
        np.savez_compressed(
            os.path.join(FX.GOLDEN_DIR, f"golden_{tag}.npz"),
            encoder=fr.numpy(), gen=genr.numpy()[:, :, ::2, ::2, ::2], raw_c3=rawr.numpy()[:, :, 3, ::2, ::2, ::2],
            merged=mr.numpy(), final=vr.numpy(), counts=counts.numpy(), iou=iou.numpy(), f1=f1.numpy(),
            resnet=taps["resnet"].numpy()[:, ::8], swin_last=sw[-1].numpy()[:, ::16],
            post_cva=taps["post_cva"].numpy()[:, :, ::16], merger_weights=taps["merger_weights"].numpy()[:, :, ::2, ::2, ::2],
            B=B, V=V)
        print(f"{tag}: reference == oracle (bit-exact); golden written; IoU {iou.tolist()}")


if __name__ == "__main__":
    main()

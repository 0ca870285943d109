"""ORACLE (test infrastructure): seeded weights and inputs for parity runs.

SURVEY 7.3-H1: with PyTorch's default init (or the reference's own init_weights) the SwinVox forward is
numerically degenerate -- refined logits are input independent to 1e-5 and sit on the 0.5 threshold.
Parity needs weights under which every stage is O(1) and input dependent.  `calibrated_` below draws
such weights analytically (fan-in scaled gaussians, randomised BatchNorm statistics) from the CPU
generator only, so the same seed gives bit-identical weights in the build container and on the GPU box
and the golden vectors in tests/golden/ stay valid without shipping 335 MB of weights.

Fan-in scaling alone still lets ReLU stacks drift (positive means swamp the input dependence), so the
calibrated regime adds one LSUV-style pass (`calibrate`): a single oracle forward on a structured
calibration batch that re-centres / re-scales each conv/linear(+BN) unit with ONE scalar scale and ONE
scalar shift.  Those ~250 scalar pairs are data dependent, hence committed as
tests/golden/calibration_*.json (written by oracle/make_golden.py) and re-applied by `build`.
"""
import json
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import modules as M

REGIMES = ("default", "init_weights", "analytic", "calibrated")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def structured_inputs(B, V, seed=1234):
    """images in [-1, 1] with low-frequency structure (i.i.d. pixel noise averages out inside the network
    and makes every sample look alike); deterministic in the CPU generator"""
    g = torch.Generator().manual_seed(seed)
    n = B * V
    coarse = F.interpolate(torch.rand(n, 3, 7, 7, generator=g) * 2 - 1, size=224, mode="bilinear", align_corners=False)
    mid = F.interpolate(torch.rand(n, 3, 28, 28, generator=g) * 2 - 1, size=224, mode="bilinear", align_corners=False)
    fine = torch.rand(n, 3, 224, 224, generator=g) * 2 - 1
    img = (1.2 * coarse + 0.6 * mid + 0.15 * fine).clamp_(-1, 1)
    return img.reshape(B, V, 3, 224, 224).contiguous()


def seeded_inputs(B, V, seed=1234):
    """images uniform in [-1, 1] (the clamp domain of core/train.py:226)"""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, V, 3, 224, 224, generator=g) * 2 - 1


def seeded_gt(B, seed=4321, p=0.1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, 32, 32, 32, generator=g) < p).float()


def _randomise_bn(bn, g):
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
        bn.weight.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(bn.num_features, generator=g) * 0.1)


def analytic_(module, seed, gain=1.0):
    """in place; deterministic in (module structure, seed)"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in module.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                fan_in = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain * math.sqrt(2.0 / fan_in)))
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, nn.ConvTranspose3d):
                # every output voxel of a stride-2 transposed conv sees 1/8 of the kernel taps
                k = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2]
                taps = max(k // 8, 1)
                fan_in = m.in_channels * taps
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain * math.sqrt(2.0 / fan_in)))
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain * math.sqrt(1.0 / m.in_features)))
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                _randomise_bn(m, g)
            elif isinstance(m, nn.LayerNorm):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        for name, p in module.named_parameters():
            if name.endswith("relative_position_bias_table"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            # keep the residual branches of the ResNet bottlenecks from exploding (no BN renormalisation in eval)
            if ".conv3.weight" in name:
                p.mul_(0.4)
    return module


def _units(module):
    """(name, layer) of every unit that gets a scalar (scale, shift): BatchNorms, and convs / linears whose
    output does not go straight into a BatchNorm"""
    names = dict(module.named_modules())
    out = []
    for name, m in names.items():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            out.append((name, m))
        elif isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose3d, nn.Linear)):
            parent, _, leaf = name.rpartition(".")
            sib = None
            if leaf.isdigit():
                sib = names.get(f"{parent}.{int(leaf) + 1}" if parent else str(int(leaf) + 1))
            elif leaf.startswith("conv"):
                sib = names.get(f"{parent}.bn{leaf[4:]}")
            if not isinstance(sib, (nn.BatchNorm2d, nn.BatchNorm3d)):
                out.append((name, m))
    return out


TARGETS = {"decoder.layer5.0": (0.0, 2.0), "refiner.layer8.0": (0.0, 2.5), "merger.layer6.1": (0.0, 1.5)}


def _rescale(m, scale, mean, tmean):
    with torch.no_grad():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            m.weight.mul_(scale)
            m.bias.copy_((m.bias - mean) * scale + tmean)
        else:
            m.weight.mul_(scale)
            if m.bias is not None:
                m.bias.copy_((m.bias - mean) * scale + tmean)


def calibrate(mods, cfg, images):
    """One LSUV-style forward of the oracle: returns {qualified unit name: [scale, mean, target_mean]}.
    Hooks rewrite each unit in place and hand the corrected output downstream."""
    calib, hooks = {}, []

    def make(qname, m):
        def hook(mod, inp, out):
            tmean, tstd = TARGETS.get(qname, (0.0, 1.0))
            mean = float(out.mean()) if (getattr(mod, "bias", None) is not None) else 0.0
            std = float((out - mean).std())
            scale = tstd / max(std, 1e-12)
            if getattr(mod, "bias", None) is None:
                tmean = 0.0
            # round-trip through float32 so the stored scalars reproduce the exact same parameters
            scale, mean = float(torch.tensor(scale, dtype=torch.float32)), float(torch.tensor(mean, dtype=torch.float32))
            calib[qname] = [scale, mean, tmean]
            _rescale(mod, scale, mean, tmean)
            return (out - mean) * scale + tmean
        return hook

    for mk, mod in mods.items():
        for name, m in _units(mod):
            hooks.append(m.register_forward_hook(make(f"{mk}.{name}", m)))
    with torch.no_grad():
        M.forward_pipeline(mods["encoder"], mods["decoder"], mods["merger"], mods["refiner"], images, cfg)
    for h in hooks:
        h.remove()
    return calib   # modules the configuration skips (USE_MERGER / USE_REFINER off) have no entries: they keep the analytic draw


def apply_calibration(mods, calib):
    for mk, mod in mods.items():
        for name, m in _units(mod):
            if f"{mk}.{name}" not in calib:   # a module the configuration never runs (see calibrate)
                continue
            scale, mean, tmean = calib[f"{mk}.{name}"]
            _rescale(m, scale, mean, tmean)
    return mods


def cfg_tag(cfg):
    n = cfg.NETWORK
    tag = "ms{}_st{}_cva{}_r{}d{}h{}".format(int(n.USE_SWIN_T_MULTI_STAGE), "".join(map(str, n.SWIN_T_STAGES)),
                                            int(n.USE_CROSS_VIEW_ATTENTION), n.CROSS_ATT_REDUCTION_RATIO,
                                            n.ATT_SPATIAL_DOWNSAMPLE_RATIO, n.CROSS_ATT_NUM_HEADS)
    # non-default switches that change which units exist / run (suffixes only, so the first fixtures keep their names)
    if n.TCONV_USE_BIAS:
        tag += "_tb1"
    if not n.USE_MERGER:
        tag += "_mg0"
    if not n.USE_REFINER:
        tag += "_rf0"
    return tag


def calibration_path(cfg, seed):
    return os.path.join(GOLDEN_DIR, f"calibration_{cfg_tag(cfg)}_seed{seed}.json")


def build(cfg=None, regime="calibrated", seed=0, factory=None):
    """Four modules (encoder, decoder, merger, refiner) in eval mode under one init regime.
    `factory`: dict of constructors (defaults to the oracle's Ref* classes) so the same weights can be
    drawn for the real reference modules or the product modules."""
    cfg = cfg or M.default_cfg()
    f = factory or dict(encoder=M.RefEncoder, decoder=M.RefDecoder, merger=M.RefMerger, refiner=M.RefRefiner)
    torch.manual_seed(seed)
    mods = {k: f[k](cfg) for k in ("encoder", "decoder", "merger", "refiner")}
    for i, (k, m) in enumerate(mods.items()):
        if regime == "init_weights":
            m.apply(M.init_weights)
        elif regime in ("analytic", "calibrated"):
            analytic_(m, seed * 16 + i + 1)
        elif regime != "default":
            raise ValueError(regime)
        m.eval()
    if regime == "calibrated":
        with open(calibration_path(cfg, seed)) as fh:
            apply_calibration(mods, json.load(fh))
    return mods

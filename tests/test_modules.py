"""Parity of the drop-in modules (swinvox_b200/models/*.py) with the oracle and with the golden vectors of the real
reference, per stage, through the reference's own calling convention (core/test.py:120-130).  Each test runs on
`hostsim` (CPU tier: checks lowering, weight re-layout, plan construction) and on `cuda` (gpu tier: the sm_100a
kernels through the C-ABI)."""
import pytest
import torch

from oracle import fixtures as FX
from oracle import modules as M
from swinvox_b200.models import CrossViewAttention, Decoder, Encoder, Merger, Refiner, SwinTransformer
from util import RTOL_DEEP, RTOL_INTERNAL, dev, golden, parity_log, stage_check, sync, voxel_check  # noqa: F401

PRODUCT = dict(encoder=Encoder, decoder=Decoder, merger=Merger, refiner=Refiner)
CFGS = {   # the fixtures of oracle/make_golden.py: every NETWORK switch the hot path reads (config.py:83-94)
    "default": (dict(), 1, 2),
    "single_stage_nocva": (dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3], USE_CROSS_VIEW_ATTENTION=False), 1, 1),
    "stages_3": (dict(SWIN_T_STAGES=[3]), 1, 2),
    "stages_23": (dict(SWIN_T_STAGES=[2, 3]), 1, 2),
    "stages_13": (dict(SWIN_T_STAGES=[1, 3]), 1, 2),
    "tconv_bias_ratio1": (dict(TCONV_USE_BIAS=True, ATT_SPATIAL_DOWNSAMPLE_RATIO=1), 1, 2),
    "nocva": (dict(USE_CROSS_VIEW_ATTENTION=False), 1, 2),
    "nomerger": (dict(USE_MERGER=False), 1, 2),
    "norefiner": (dict(USE_REFINER=False), 1, 2),
}
# the CPU tier (hostsim) lowers every configuration but runs the numeric comparison on a subset to stay within minutes
HOSTSIM_TAGS = ("default", "single_stage_nocva", "stages_13", "tconv_bias_ratio1", "nomerger")


def nchw(act):
    return act.view().squeeze(1).permute(0, 3, 1, 2)


def oracle_forward(mods, images, cfg=None):
    """core/test.py:120-130 on the oracle modules, keeping every stage output"""
    net = (cfg or M.default_cfg()).NETWORK
    taps = {}
    with torch.no_grad():
        f = mods["encoder"](images, taps)
        raw, gen = mods["decoder"](f)
        m = mods["merger"](raw, gen, taps) if net.USE_MERGER else gen.mean(1)
        v = mods["refiner"](m, taps) if net.USE_REFINER else m
    taps.update(encoder=f, raw=raw, gen=gen, merged=m, final=v)
    return taps


@pytest.mark.parametrize("tag", list(CFGS))
def test_pipeline_parity_with_oracle_and_reference_golden(dev, tag):
    if dev == "cpu" and tag not in HOSTSIM_TAGS:
        pytest.skip("numeric comparison of this configuration runs in the gpu tier")
    from swinvox_b200.pipeline import ViewMean
    over, B, V = CFGS[tag]
    cfg = M.default_cfg(**over)
    net = cfg.NETWORK
    ref = oracle_forward(FX.build(cfg, "calibrated", 0), FX.structured_inputs(B, V, seed=1234), cfg)
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    for m in prod.values():
        m.to(dev)
    images = FX.structured_inputs(B, V, seed=1234).to(dev)
    with torch.no_grad():  # the reference's calling sequence, core/test.py:120-130
        f = prod["encoder"](images)
        raw, gen = prod["decoder"](f)
        merged = prod["merger"](raw, gen) if net.USE_MERGER else ViewMean()(gen)
        final = prod["refiner"](merged) if net.USE_REFINER else merged
    sync(dev)
    plan = next(iter(prod["encoder"]._plans.values()))[0]
    reports = [stage_check("resnet branch", nchw(plan.taps["resnet"]), ref["resnet"], RTOL_DEEP)]
    sw_ref = ref["swin"] if isinstance(ref["swin"], list) else [ref["swin"]]
    for i, (a, b) in enumerate(zip(plan.taps["swin"], sw_ref)):
        reports.append(stage_check(f"swin stage {i}", nchw(a), b, RTOL_DEEP))
    reports.append(stage_check("post_cva", nchw(plan.taps["post_cva"]).reshape(ref["post_cva"].shape), ref["post_cva"], RTOL_DEEP))
    reports.append(stage_check("encoder", f, ref["encoder"], RTOL_DEEP))
    reports.append(stage_check("decoder.raw", raw, ref["raw"]))
    reports.append(stage_check("decoder.gen", gen, ref["gen"]))
    if net.USE_MERGER:
        reports.append(stage_check("merger.weights", prod["merger"].last_volume_weights, ref["merger_weights"], RTOL_INTERNAL))
    reports.append(stage_check("merger", merged, ref["merged"]))
    reports.append(stage_check("refiner", final, ref["final"]))
    # ... and against what the unmodified reference produced in the build container
    g = golden(tag)
    reports.append(stage_check("encoder vs reference golden", f, torch.from_numpy(g["encoder"]), RTOL_DEEP))
    reports.append(stage_check("merged vs reference golden", merged, torch.from_numpy(g["merged"])))
    reports.append(stage_check("final vs reference golden", final, torch.from_numpy(g["final"])))
    vox = voxel_check(final, torch.from_numpy(g["final"]), FX.seeded_gt(B))
    print("\n".join(reports))
    print("voxels (th, mismatch, out-of-band mismatch, dIoU):", vox)
    parity_log(f"{tag} [{dev}] B={B} V={V}", reports, vox)


def test_modules_accept_foreign_tensors(dev):
    """each module alone, fed plain contiguous NCHW tensors as a reference caller would"""
    cfg = M.default_cfg()
    ora = FX.build(cfg, "calibrated", 0)
    ref = oracle_forward(ora, FX.structured_inputs(1, 2, seed=1234))
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    with torch.no_grad():
        dec, mer, rf = prod["decoder"].to(dev), prod["merger"].to(dev), prod["refiner"].to(dev)
        raw, gen = dec(ref["encoder"].to(dev))
        stage_check("decoder.raw", raw, ref["raw"])
        stage_check("decoder.gen", gen, ref["gen"])
        merged = mer(ref["raw"].to(dev), ref["gen"].to(dev))
        stage_check("merger", merged, ref["merged"])
        stage_check("refiner", rf(ref["merged"].to(dev)), ref["final"])
        # foreign input after a chained call must not disturb the upstream module's output buffer
        keep = raw.clone()
        dec(torch.zeros_like(ref["encoder"]).to(dev))
        mer(raw, gen)
        dec2 = dec(ref["encoder"].to(dev))[0]
        assert torch.equal(dec2, keep)
        cva_ref = ora["encoder"].cross_view_attention
        cva = CrossViewAttention(cfg, 512)
        cva.load_state_dict(cva_ref.state_dict())
        cva.eval().to(dev)
        stage_check("cross_view_attention", cva(ref["pre_cva"].to(dev)), ref["post_cva"], RTOL_DEEP)
    sync(dev)


def test_merger_fp16_range_fallback(dev):
    """activations beyond fp16's range: the merger reports the saturation and switches itself to tf32 operands"""
    cfg = M.default_cfg()
    ora = FX.build(cfg, "calibrated", 0)["merger"]
    g = torch.Generator().manual_seed(3)
    raw = torch.randn(1, 2, 9, 32, 32, 32, generator=g) * 2e5
    gen = torch.randn(1, 2, 32, 32, 32, generator=g)
    with torch.no_grad():
        taps = {}
        ora(raw, gen, taps)   # compare the pre-softmax scores: at this magnitude the softmax is an arg-max
        mer = Merger(cfg)
        mer.load_state_dict(ora.state_dict())
        mer.eval().to(dev)
        mer(raw.to(dev), gen.to(dev))
        sync(dev)
        assert mer.saturated() and mer.slab_operands == "tf32"
        mer(raw.to(dev), gen.to(dev))
        sync(dev)
        assert not mer.saturated()
    stage_check("merger scores (tf32 operands, |x| ~ 2e5)", mer.last_volume_weights, taps["merger_weights"], RTOL_INTERNAL)


def test_swin_wrapper_outputs(dev):
    cfg = M.default_cfg()
    ora = FX.build(cfg, "calibrated", 0)["encoder"].swin_transformer
    sw = SwinTransformer(cfg, 3, 224, pretrained=False)
    sw.load_state_dict(ora.state_dict())
    sw.eval().to(dev)
    x = FX.structured_inputs(1, 1, seed=5)[0]
    with torch.no_grad():
        ref, got = ora(x), sw(x.to(dev))
    assert isinstance(got, list) and len(got) == 4
    for i, (a, b) in enumerate(zip(got, ref)):
        assert tuple(a.shape) == tuple(b.shape)
        stage_check(f"swin wrapper stage {i}", a, b, RTOL_DEEP)
    single = M.default_cfg(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3])
    sw1 = SwinTransformer(single, 3, 224, pretrained=False).eval()
    assert sw1.out_channels == [768] and sw1.out_spatial == [7]


def test_swin_wrapper_resize_and_input_channels(dev):
    """models/swin_transformer.py:29-37 (patch embedding re-created for in_channels != 3) and :74-75 (inputs that are not
    224 x 224 are resized, bilinear / align_corners=False)"""
    cfg = M.default_cfg(SWIN_T_STAGES=[0, 1])
    torch.manual_seed(0)
    for cin, hw in ((3, (160, 200)), (4, (224, 224)), (1, (96, 96))):
        ora = M.RefSwinTransformer(cfg, cin, 224, pretrained=False)
        FX.analytic_(ora, 5)
        ora.eval()
        sw = SwinTransformer(cfg, cin, 224, pretrained=False)
        sw.load_state_dict(ora.state_dict())
        sw.eval().to(dev)
        x = FX.structured_inputs(1, 1, seed=9)[0][:, :1].repeat(1, cin, 1, 1)[..., :hw[0], :hw[1]].contiguous()
        with torch.no_grad():
            ref, got = ora(x), sw(x.to(dev))
        sync(dev)
        for i, (a, b) in enumerate(zip(got, ref)):
            stage_check(f"swin wrapper cin={cin} input {hw} stage {i}", a, b, RTOL_DEEP)


@pytest.mark.gpu
def test_batched_multi_view_parity_gpu():
    """B=2 objects x V=3 views (the metric's view count); views of one object interact only through CVA / merger"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = M.default_cfg()
    images = FX.structured_inputs(2, 3, seed=77)
    ref = oracle_forward(FX.build(cfg, "calibrated", 0), images)
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    with torch.no_grad():
        for m in prod.values():
            m.cuda()
        f = prod["encoder"](images.cuda())
        raw, gen = prod["decoder"](f)
        final = prod["refiner"](prod["merger"](raw, gen))
        # replaying the cached plans (second call, CUDA graph) must give the same bits
        for m in prod.values():
            m.use_graph = True
        for _ in range(2):
            f2 = prod["encoder"](images.cuda())
            raw2, gen2 = prod["decoder"](f2)
            final2 = prod["refiner"](prod["merger"](raw2, gen2))
    torch.cuda.synchronize()
    stage_check("encoder", f, ref["encoder"], RTOL_DEEP)
    stage_check("refiner", final, ref["final"])
    voxel_check(final, ref["final"], FX.seeded_gt(2))
    assert torch.equal(final2, final)


def _forward(prod, images):
    with torch.no_grad():
        f = prod["encoder"](images)
        raw, gen = prod["decoder"](f)
        return prod["refiner"](prod["merger"](raw, gen))


@pytest.mark.gpu
@pytest.mark.parametrize("V", [1, 5, 20, 24])
def test_view_sweep_parity_gpu(V):
    """BASELINE configs[2..4]: 5 views with cross-view attention, 20 views, and the ends of the 1..24 view sweep --
    one object each against the oracle (the views of an object meet in CVA's softmax over views and the merger)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = M.default_cfg()
    images = FX.structured_inputs(1, V, seed=100 + V)
    ref = oracle_forward(FX.build(cfg, "calibrated", 0), images)
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    for m in prod.values():
        m.cuda()
    final = _forward(prod, images.cuda())
    torch.cuda.synchronize()
    stage_check(f"refiner V={V}", final, ref["final"])
    voxel_check(final, ref["final"], FX.seeded_gt(1))


@pytest.mark.gpu
def test_full_size_properties_gpu():
    """BASELINE configs[1] at full size (64 objects x 3 views), through properties that need no oracle run:
    objects are independent (an object's logits do not depend on its batch mates), and the result is invariant
    under a permutation of an object's views (CVA is permutation-equivariant, the merger's softmax-weighted sum is
    permutation-invariant: cross_view_attention.py:81-99, merger.py:98-104)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = M.default_cfg()
    prod = FX.build(cfg, "calibrated", 0, PRODUCT)
    for m in prod.values():
        m.cuda()
    g = torch.Generator().manual_seed(9)
    images = (torch.rand(64, 3, 3, 224, 224, generator=g) * 2 - 1).cuda()
    full = _forward(prod, images).clone()
    assert torch.isfinite(full).all()
    pair = _forward(prod, images[5:7].contiguous()).clone()
    scale = full.abs().max().item()
    assert (full[5:7] - pair).abs().max().item() <= 1e-5 * scale
    perm = _forward(prod, images[:, [2, 0, 1]].contiguous()).clone()
    assert (perm - full).abs().max().item() <= 1.5e-3 * scale   # one TF32 rounding flip of a stored activation = 4.9e-4
    # threshold counters of the full batch: bit-exact against the same logits counted by torch (core/test.py:141-164)
    from swinvox_b200.metrics import VoxelMetrics
    gt = (torch.rand(64, 32, 32, 32, generator=g) < 0.1).float().cuda()
    counts = VoxelMetrics(cfg.TEST.VOXEL_THRESH).counts(full, gt).cpu()
    for ti, th in enumerate(cfg.TEST.VOXEL_THRESH):
        v = (torch.sigmoid(full) >= th).float()
        inter = (v * gt).flatten(1).sum(1).cpu().long()
        union = ((v + gt) >= 1).flatten(1).sum(1).cpu().long()
        assert torch.equal(counts[:, ti, 0].long(), inter) and torch.equal(counts[:, ti, 1].long(), union)


def test_eval_only_guards(dev):
    cfg = M.default_cfg()
    ref = Refiner(cfg).to(dev)
    x = torch.zeros(1, 32, 32, 32, device=dev)
    with pytest.raises(RuntimeError, match="inference-only"):
        ref.train()(x)
    with pytest.raises(RuntimeError, match="no_grad"):
        ref.eval()(x)
    with torch.no_grad(), pytest.raises(ValueError):
        ref(torch.zeros(1, 16, 16, 16, device=dev))


def test_cpu_tensor_is_rejected_loudly():
    """no CPU fallback: the real package refuses CPU tensors"""
    from swinvox_b200 import _lib
    with torch.no_grad(), pytest.raises(_lib.SvxError, match="no CPU fallback"):
        Refiner(M.default_cfg()).eval()(torch.zeros(1, 32, 32, 32))


@pytest.mark.parametrize("over", [dict(), dict(USE_SWIN_T_MULTI_STAGE=False, SWIN_T_STAGES=[3]),
                                  dict(SWIN_T_STAGES=[2, 3], USE_CROSS_VIEW_ATTENTION=False),
                                  dict(TCONV_USE_BIAS=True, ATT_SPATIAL_DOWNSAMPLE_RATIO=1)])
def test_state_dict_layout_matches_reference(over):
    """keys, order and shapes equal the oracle's, which make_golden.py pinned to the real reference modules;
    strict load works both ways and with DataParallel's `module.` prefix stripped"""
    cfg = M.default_cfg(**over)
    torch.manual_seed(0)
    pairs = [(Encoder(cfg), M.RefEncoder(cfg)), (Decoder(cfg), M.RefDecoder(cfg)), (Merger(cfg), M.RefMerger(cfg)),
             (Refiner(cfg), M.RefRefiner(cfg))]
    for mine, ref in pairs:
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a) == list(b), type(mine).__name__
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        mine.load_state_dict(b, strict=True)
        ref.load_state_dict(mine.state_dict(), strict=True)
        assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in ref.parameters())

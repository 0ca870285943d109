"""Parity of every kernel in libswinvox_b200 against plain PyTorch fp64 references of the same op (run on
the CPU).  Inputs that feed a tensor-core contraction are pre-rounded to TF32 so the only difference
left is accumulation order: tolerances are therefore tight (1e-4 of the output scale).

Every test runs twice: `cuda` (marked gpu: the real sm_100a kernels through the C-ABI) and `hostsim`
(CPU tier: the test-only CPU twin in tests/hostsim, which checks the host-side packing / tap tables /
descriptors and the reference code of the test itself)."""
import pytest
import torch
import torch.nn.functional as F

from swinvox_b200 import engine as E



from util import dev, sync  # noqa: F401  (dual-backend fixture)


def rel_err(got, ref):
    ref = ref.to(torch.float64)
    return ((got.detach().cpu().to(torch.float64) - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def act_from_nchw(x, DEV):  # [N,C,H,W] or [N,C,D,H,W] cpu -> Act on GPU
    if x.dim() == 4:
        x = x.unsqueeze(2)
    N, Cc, D, H, W = x.shape
    return E.Act(x.permute(0, 2, 3, 4, 1).contiguous().view(-1, Cc).to(DEV), N, D, H, W, Cc)


def to_nchw(a):
    return a.view().permute(0, 4, 1, 2, 3).cpu()


def rand_bn(bn):
    bn.eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    return bn


@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (127, 8, 64), (300, 96, 96), (1000, 288, 96), (4096, 256, 1024),
                                   (513, 192, 768), (64, 2048, 1024), (2500, 64, 148), (777, 1000, 384)])
def test_gemm_plain(dev, M, N, K):
    DEV = dev
    torch.manual_seed(M + N + K)
    x = E.tf32_round(torch.randn(M, K))
    w = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N)
    p = E.Plan(DEV)
    pk = E.pack_matrix(w, b, DEV)
    out = p.new_act(M, 1, 1, 1, N, Cs=E.round_up(N, 4))
    p.linear(E.Act(x.to(DEV), M, 1, 1, 1, K), pk, out)
    p.run()
    sync(DEV)
    ref = x.double() @ E.tf32_round(w).double().t() + b.double()
    assert rel_err(out.view().reshape(M, N), ref) < 1e-4


@pytest.mark.parametrize("block_n", [16, 32, 64, 96, 128, 192, 256])
def test_gemm_block_n(dev, block_n):
    DEV = dev
    torch.manual_seed(block_n)
    M, K, N = 700, 416, 2 * block_n
    x = E.tf32_round(torch.randn(M, K))
    w = torch.randn(N, K) / K ** 0.5
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, N)
    p.linear(E.Act(x.to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, None, DEV, block_n=block_n), out)
    p.run()
    sync(DEV)
    assert rel_err(out.view().reshape(M, N), x.double() @ E.tf32_round(w).double().t()) < 1e-4


@pytest.mark.parametrize("act", [E.ACT_NONE, E.ACT_RELU, E.ACT_LEAKY, E.ACT_GELU])
@pytest.mark.parametrize("res_after", [False, True])
def test_gemm_epilogue(dev, act, res_after):
    DEV = dev
    torch.manual_seed(5)
    M, K, N = 333, 160, 96
    x = E.tf32_round(torch.randn(M, K))
    w, b, r = torch.randn(N, K) / K ** 0.5, torch.randn(N), torch.randn(M, N)
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, N)
    res = E.Act(r.to(DEV), M, 1, 1, 1, N)
    p.linear(E.Act(x.to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, b, DEV), out, act=act, act_param=0.2, residual=res,
             res_after_act=res_after, out_scale=0.5, round_out=True)
    p.run()
    sync(DEV)
    f = {E.ACT_NONE: lambda t: t, E.ACT_RELU: F.relu, E.ACT_LEAKY: lambda t: F.leaky_relu(t, 0.2),
         E.ACT_GELU: F.gelu}[act]
    y = x.double() @ E.tf32_round(w).double().t() + b.double()
    ref = 0.5 * (f(y) + r.double() if res_after else f(y + r.double()))
    assert rel_err(out.view().reshape(M, N), ref) < 6e-4  # result is rounded to TF32 (2^-11)


@pytest.mark.parametrize("C,M", [(96, 192 * 56 * 56), (192, 192 * 28 * 28)])
def test_mlp_fused_full_size_identity(dev, C, M):
    """BASELINE-size property (64 objects x 3 views, Swin stages 0 / 1): with W2 = 0 the fused kernel must return
    residual + b2 bit-exactly for every row -- every tile of every CTA goes through the staged / transposed tile tail."""
    DEV = dev
    if DEV == "cpu":
        M = 4 * 128 + 3   # the CPU twin has no tiles to get wrong; keep the CPU tier fast
    torch.manual_seed(C)
    hid = 4 * C
    x = E.tf32_round(torch.randn(M, C))
    res = torch.randn(M, C)
    b2 = torch.randn(C)
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, C)
    p.mlp(E.Act(x.to(DEV), M, 1, 1, 1, C), E.pack_matrix(torch.randn(hid, C) / C ** 0.5, torch.randn(hid), DEV),
          E.pack_matrix(torch.zeros(C, hid), b2, DEV), out, residual=E.Act(res.to(DEV), M, 1, 1, 1, C))
    p.run()
    sync(DEV)
    assert torch.equal(out.view().reshape(M, C).cpu(), res + b2)


@pytest.mark.parametrize("M", [128, 1000, 148 * 128 * 3 + 77])
def test_mlp_fused_with_layernorm(dev, M):
    """the stage-0 variant that also applies the block's norm2: x1 -> LN -> fc1 -> GELU -> fc2 -> + x1"""
    DEV = dev
    torch.manual_seed(M)
    C, hid = 96, 384
    x1 = torch.randn(M, C) * 1.5 + 0.3
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    w1, b1 = torch.randn(hid, C) / C ** 0.5, torch.randn(hid) * 0.5
    w2, b2 = torch.randn(C, hid) / hid ** 0.5, torch.randn(C) * 0.5
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, C)
    xa = E.Act(x1.to(DEV), M, 1, 1, 1, C)
    p.mlp(xa, E.pack_matrix(w1, b1, DEV), E.pack_matrix(w2, b2, DEV), out, residual=xa, ln=(g.to(DEV), b.to(DEV), 1e-5))
    p.run()
    p.run()
    sync(DEV)
    y = F.layer_norm(x1.double(), (C,), g.double(), b.double(), 1e-5)
    h = F.gelu(E.tf32_round(y.float()).double() @ E.tf32_round(w1).double().t() + b1.double())
    ref = E.tf32_round(h.float()).double() @ E.tf32_round(w2).double().t() + b2.double() + x1.double()
    assert rel_err(out.view().reshape(M, C), ref) < 5e-4   # LN output rounded to TF32 at slightly different values


@pytest.mark.parametrize("M,C", [(128, 96), (1000, 96), (19 * 128 + 5, 96), (148 * 128 * 2 + 77, 96), (640, 192),
                                 (3001, 192), (148 * 128 + 130, 192)])
def test_mlp_fused(dev, M, C):
    """timm Mlp + residual as one kernel (svx_mlp.cu): fc1 -> erf-GELU -> TF32 rounding -> fc2 -> + residual.  Sizes cover
    one tile, ragged last tiles, several tiles per CTA (the chunk pipeline crosses tile boundaries) and both widths."""
    DEV = dev
    torch.manual_seed(M + C)
    hid = 4 * C
    x = E.tf32_round(torch.randn(M, C))
    w1, b1 = torch.randn(hid, C) / C ** 0.5, torch.randn(hid) * 0.5
    w2, b2 = torch.randn(C, hid) / hid ** 0.5, torch.randn(C) * 0.5
    res = torch.randn(M, C)
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, C)
    p.mlp(E.Act(x.to(DEV), M, 1, 1, 1, C), E.pack_matrix(w1, b1, DEV), E.pack_matrix(w2, b2, DEV), out,
          residual=E.Act(res.to(DEV), M, 1, 1, 1, C))
    p.run()
    p.run()   # a second launch must not depend on leftover barrier / TMEM state
    sync(DEV)
    h = F.gelu(x.double() @ E.tf32_round(w1).double().t() + b1.double())
    h = E.tf32_round(h.float()).double()
    ref = h @ E.tf32_round(w2).double().t() + b2.double() + res.double()
    assert rel_err(out.view().reshape(M, C), ref) < 2e-4


@pytest.mark.parametrize("M,K,N,bn", [(333, 160, 256, 256), (1000, 64, 512, 128), (4096, 256, 1024, None)])
def test_gemm_residual_on_tensor_cores(dev, M, K, N, bn):
    """ResNet conv3-style epilogue relu(x W^T + b + r) with the TF32-exact residual added by the MMA itself
    (identity columns appended to the weights, residual tile streamed by TMA as extra k-chunks)"""
    DEV = dev
    torch.manual_seed(M)
    x = E.tf32_round(torch.randn(M, K))
    w, b, r = torch.randn(N, K) / K ** 0.5, torch.randn(N), E.tf32_round(torch.randn(M, N))
    p = E.Plan(DEV)
    out = p.new_act(M, 1, 1, 1, N)
    res = E.Act(r.to(DEV), M, 1, 1, 1, N)
    p.linear(E.Act(x.to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, b, DEV, block_n=bn), out, act=E.ACT_RELU, residual=res,
             res_after_act=False, res_via_mma=True)
    p.run()
    sync(DEV)
    ref = F.relu(x.double() @ E.tf32_round(w).double().t() + b.double() + r.double())
    assert rel_err(out.view().reshape(M, N), ref) < 1e-4


@pytest.mark.parametrize("cin,cout,hw,k,s,p_", [(64, 64, 14, 3, 1, 1), (32, 96, 15, 3, 2, 1), (256, 128, 7, 1, 1, 0),
                                                 (128, 256, 14, 1, 2, 0), (8, 16, 9, 5, 2, 2), (512, 256, 7, 3, 1, 1),
                                                 (4, 64, 30, 7, 2, 3), (4, 96, 24, 4, 4, 0), (4, 32, 11, 3, 1, 1)])
def test_conv2d_gather(dev, cin, cout, hw, k, s, p_):
    DEV = dev
    torch.manual_seed(cin + cout)
    x = E.tf32_round(torch.randn(3, cin, hw, hw))
    conv = torch.nn.Conv2d(cin, cout, k, s, p_)
    bn = rand_bn(torch.nn.BatchNorm2d(cout))
    with torch.no_grad():
        conv.weight.copy_(conv.weight * 3)
    oh = (hw + 2 * p_ - k) // s + 1
    p = E.Plan(DEV)
    pk = E.pack_conv(conv.weight, conv.bias, bn, DEV)
    out = p.new_act(3, 1, oh, oh, cout)
    p.conv(act_from_nchw(x, DEV), pk, E.conv_taps(1, k, k, 0, p_, p_), out, stride=(1, s, s), act=E.ACT_RELU)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.relu(F.conv2d(x.double(), E.tf32_round(wf).double(), bf.double(), s, p_))
    assert rel_err(to_nchw(out).squeeze(2), ref) < 1e-4


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 14), (128, 96, 9), (512, 256, 7)])
def test_conv2d_flat_tma(dev, cin, cout, hw):
    """stride-1 3x3 conv over a zero-bordered input, A operand streamed by TMA (one shifted box per tap), writing
    into the interior of another zero-bordered buffer"""
    DEV = dev
    torch.manual_seed(cin)
    x = E.tf32_round(torch.randn(3, cin, hw, hw))
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
    bn = rand_bn(torch.nn.BatchNorm2d(cout))
    p = E.Plan(DEV)
    xin = p.new_act(3, 1, hw, hw, cin, pad=(0, 1, 1))
    xin.view().copy_(x.permute(0, 2, 3, 1).unsqueeze(1))
    out = p.new_act(3, 1, hw, hw, cout, pad=(0, 1, 1))
    p.conv_flat(xin, E.pack_conv(conv.weight, conv.bias, bn, DEV), E.conv_taps(1, 3, 3, 0, 0, 0), out, act=E.ACT_RELU)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.relu(F.conv2d(x.double(), E.tf32_round(wf).double(), bf.double(), 1, 1))
    assert rel_err(out.view().squeeze(1).permute(0, 3, 1, 2), ref) < 1e-4
    full = out.buf.view(3, hw + 2, hw + 2, cout).cpu()
    assert full[:, 0].abs().max() == 0 and full[:, -1].abs().max() == 0 and full[:, :, 0].abs().max() == 0 \
        and full[:, :, -1].abs().max() == 0, "the zero border must stay untouched"


def test_conv3d_padded_channels(dev):
    """merger-style Conv3d 9->9 k3 on a 16-channel padded buffer read at a channel offset"""
    DEV = dev
    torch.manual_seed(1)
    x = E.tf32_round(torch.randn(2, 9, 8, 8, 8))
    conv = torch.nn.Conv3d(9, 9, 3, padding=1)
    bn = rand_bn(torch.nn.BatchNorm3d(9))
    p = E.Plan(DEV)
    buf = p.zeros(2 * 512, 64)
    xin = E.Act(buf, 2, 8, 8, 8, 9, 16)
    xin.view().copy_(x.permute(0, 2, 3, 4, 1))
    out = E.Act(buf, 2, 8, 8, 8, 16, 32)
    pk = E.pack_conv(conv.weight, conv.bias, bn, DEV, cin_pad=16, n_logical=16)
    p.conv(xin, pk, E.conv_taps(3, 3, 3, 1, 1, 1), out, act=E.ACT_LEAKY, act_param=0.2)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.leaky_relu(F.conv3d(x.double(), E.tf32_round(wf).double(), bf.double(), padding=1), 0.2)
    got = out.view().permute(0, 4, 1, 2, 3).cpu()
    assert rel_err(got[:, :9], ref) < 1e-4
    assert got[:, 9:].abs().max().item() == 0.0


@pytest.mark.parametrize("case", ["huge_weights", "tiny_weights", "tiny_activations", "huge_activations_tf32",
                                  "saturating_fp16"])
def test_conv3d_slab_operand_range(dev, case):
    """The slab kernel's MMA operands are fp16 (5 exponent bits).  BatchNorm-folded weights of any magnitude are handled
    by the host's power-of-two weight scale; activations are exact in [6.1e-5, 65504], lose only absolute accuracy
    (< 3e-8) below, and SATURATE above -- reported through range_flag, after which the caller re-runs with tf32 operands
    (Merger.saturated does this)."""
    DEV = dev
    torch.manual_seed(5)
    n, cin, D, H, W = 2, 9, 4, 6, 8
    xs = {"tiny_activations": 1e-3, "huge_activations_tf32": 3e5, "saturating_fp16": 3e5}.get(case, 1.0)
    x = E.tf32_round(torch.randn(n, cin, D, H, W) * xs)
    conv = torch.nn.Conv3d(cin, 9, 3, padding=1)
    bn = rand_bn(torch.nn.BatchNorm3d(9))
    with torch.no_grad():
        if case == "huge_weights":     # BN with a tiny running variance: folded weights ~1e6
            bn.running_var.fill_(1e-12)
            bn.eps = 1e-13
        if case == "tiny_weights":
            conv.weight.mul_(1e-9)
    operands = "tf32" if case == "huge_activations_tf32" else "fp16"
    p = E.Plan(DEV)
    src = p.new_act(n, D, H, W, 32, pad=(1, 1, 1))
    src.view()[..., :cin].copy_(x.permute(0, 2, 3, 4, 1))
    dst = p.new_act(n, D, H, W, 16, pad=(1, 1, 1))
    flag = p.zeros(1, dtype=torch.int32)
    pk = E.pack_conv3_slab(conv.weight, conv.bias, bn, DEV, n_logical=16)
    assert 2.0 ** 13 <= pk.W.abs().max().item() < 2.0 ** 14   # whatever the folded magnitude
    p.conv3_slab(E.Act(src.buf, n, D + 2, H + 2, W + 2, 32, 0, (1, 1, 1)), pk, dst, cin, operands=operands, range_flag=flag)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    ws = pk.acc_scale
    ref = F.conv3d(x.double(), (E.tf32_round(wf / ws) * ws).double(), bf.double(), padding=1)
    got = dst.view().permute(0, 4, 1, 2, 3).cpu()[:, :9]
    if case == "saturating_fp16":
        assert flag.item() == 1                       # loud, not silent
        assert torch.isfinite(got).all()              # saturated, never inf / nan
    else:
        assert flag.item() == 0
        assert rel_err(got, ref) < (2e-4 if case == "tiny_activations" else 1e-4)


@pytest.mark.parametrize("dims,c0,residual", [((6, 7, 9), 16, False), ((5, 32, 32), 0, False), ((4, 6, 8), 32, True),
                                               ((3, 32, 32), 48, True), ((32, 32, 32), 16, False)])
def test_conv3d_slab_merger_style(dev, dims, c0, residual):
    """merger-style Conv3d(k3, p1) with 9 output channels over a zero-bordered volume on the depth-marching TMA slab
    kernel (kw taps folded into the MMA's N, (kd,kh) taps addressed inside the depth slabs): reads a 32-channel box
    at channel offset c0 (overhanging the 64-channel buffer at c0=48), writes a 16-channel group of another
    zero-bordered buffer (optionally accumulating on it: layer5's second pass) and, like layer6, a planar
    single-channel volume."""
    DEV = dev
    torch.manual_seed(sum(dims) + c0)
    D, H, W = dims
    n = 2
    cin = 9 if c0 == 48 else 25
    x = E.tf32_round(torch.randn(n, cin, D, H, W))
    conv = torch.nn.Conv3d(cin, 9, 3, padding=1)
    bn = rand_bn(torch.nn.BatchNorm3d(9))
    conv1 = torch.nn.Conv3d(cin, 1, 3, padding=1)
    p = E.Plan(DEV)
    src = p.new_act(n, D, H, W, 64, pad=(1, 1, 1))
    src.view()[..., c0:c0 + cin].copy_(x.permute(0, 2, 3, 4, 1))
    xin = E.Act(src.buf, n, D + 2, H + 2, W + 2, 32, c0, (1, 1, 1))
    dst = p.new_act(n, D, H, W, 64, pad=(1, 1, 1))
    out = dst.channels(32, 16)
    r = torch.randn(n, 16, D, H, W) if residual else None
    if residual:
        out.view().copy_(r.permute(0, 2, 3, 4, 1))
    p.conv3_slab(xin, E.pack_conv3_slab(conv.weight, conv.bias, bn, DEV, n_logical=16), out, cin, act=E.ACT_LEAKY,
                 act_param=0.2, residual=out if residual else None, res_after_act=False)
    planar = p.empty(n, D * H * W)
    p.conv3_slab(xin, E.pack_conv3_slab(conv1.weight, conv1.bias, None, DEV), E.Act(planar.view(-1, 1), n, D, H, W, 1, 0), cin)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    y = F.conv3d(x.double(), E.tf32_round(wf).double(), bf.double(), padding=1)
    if residual:
        y = y + r[:, :9].double()
    ref = F.leaky_relu(y, 0.2)
    got = out.view().permute(0, 4, 1, 2, 3).cpu()
    assert rel_err(got[:, :9], ref) < 1e-4
    if residual:
        assert rel_err(got[:, 9:], F.leaky_relu(r[:, 9:].double(), 0.2)) < 1e-6
    else:
        assert got[:, 9:].abs().max().item() == 0.0
    full = dst.buf.view(n, D + 2, H + 2, W + 2, 64).cpu()
    assert full[:, 0].abs().max() == 0 and full[:, -1].abs().max() == 0 and full[:, :, 0].abs().max() == 0 and \
        full[:, :, -1].abs().max() == 0 and full[:, :, :, 0].abs().max() == 0 and full[:, :, :, -1].abs().max() == 0, \
        "the zero border must stay untouched"
    assert full[..., :32].abs().max() == 0 and full[..., 48:].abs().max() == 0, "other channel groups must stay untouched"
    ref1 = F.conv3d(x.double(), E.tf32_round(conv1.weight.detach()).double(), conv1.bias.detach().double(), padding=1)
    assert rel_err(planar.view(n, 1, D, H, W).cpu(), ref1) < 1e-4


@pytest.mark.parametrize("ks,pads,ind", [((4, 4, 4), (1, 1, 1), (3, 4, 5)), ((6, 4, 4), (2, 1, 1), (2, 2, 2))])
def test_convtranspose3d_classes(dev, ks, pads, ind):
    DEV = dev
    torch.manual_seed(2)
    cin, cout = 32, 24
    x = E.tf32_round(torch.randn(2, cin, *ind))
    ct = torch.nn.ConvTranspose3d(cin, cout, ks, 2, pads, bias=False)
    bn = rand_bn(torch.nn.BatchNorm3d(cout))
    od, oh, ow = [2 * i for i in ind]
    p = E.Plan(DEV)
    skip = torch.randn(2, cout, od, oh, ow)
    res = act_from_nchw(skip, DEV)
    out = p.new_act(2, od, oh, ow, cout)
    Cs = out.Cs
    for pd in (0, 1):
        for ph in (0, 1):
            for pw in (0, 1):
                pk, taps = E.pack_convT_class(ct.weight, bn, DEV, pads, (pd, ph, pw))
                omap = (((pd * oh + ph) * ow + pw) * Cs, od * oh * ow * Cs, 2 * oh * ow * Cs, 2 * ow * Cs, 2 * Cs)
                p.conv(act_from_nchw(x, DEV), pk, taps, out, out_map=omap, rows_dhw=ind, act=E.ACT_RELU, residual=res,
                       res_after_act=True)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(ct.weight.transpose(0, 1), None, bn)
    ref = F.relu(F.conv_transpose3d(x.double(), E.tf32_round(wf).transpose(0, 1).double(), bf.double(), 2, pads))
    assert rel_err(to_nchw(out), ref + skip.double()) < 1e-4


def test_decoder_tail_epilogue(dev):
    DEV = dev
    torch.manual_seed(3)
    x = E.tf32_round(torch.randn(2, 32, 4, 4, 4))
    ct = torch.nn.ConvTranspose3d(32, 8, 4, 2, 1, bias=False)
    bn = rand_bn(torch.nn.BatchNorm3d(8))
    w5 = torch.randn(9)
    p = E.Plan(DEV)
    raw = p.new_act(2, 8, 8, 8, 16)
    coarse = p.empty(2, 512)
    Cs = 16
    for pd in (0, 1):
        for ph in (0, 1):
            for pw in (0, 1):
                pk, taps = E.pack_convT_class(ct.weight, bn, DEV, (1, 1, 1), (pd, ph, pw), block_n=16, n_logical=16)
                base = (pd * 8 + ph) * 8 + pw
                p.conv(act_from_nchw(x, DEV), pk, taps, raw, out_map=(base * Cs, 512 * Cs, 128 * Cs, 16 * Cs, 2 * Cs),
                       rows_dhw=(4, 4, 4), round_out=True,
                       epi_tail=(w5.to(DEV), coarse, (base, 512, 128, 16, 2)))
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(ct.weight.transpose(0, 1), None, bn)
    feat = F.relu(F.conv_transpose3d(x.double(), E.tf32_round(wf).transpose(0, 1).double(), bf.double(), 2, 1))
    gen = (feat * w5[:8].double().view(1, 8, 1, 1, 1)).sum(1) + w5[8].double()
    got = to_nchw(raw)
    assert rel_err(coarse.view(2, 8, 8, 8), gen) < 1e-4
    assert rel_err(got[:, :8], feat) < 6e-4
    assert rel_err(got[:, 8], gen) < 6e-4
    assert got[:, 9:].abs().max().item() == 0.0


@pytest.mark.parametrize("cin,c0,dims", [(9, 0, (5, 16, 32)), (12, 4, (3, 32, 32)), (1, 8, (2, 16, 32))])
def test_conv3d_single_output_fp32(dev, cin, c0, dims):
    """merger layer6 form: Conv3d(cin -> 1, k3, p1) + BatchNorm + LeakyReLU in fp32 on the CUDA cores"""
    DEV = dev
    torch.manual_seed(cin + c0)
    n, (D, H, W) = 2, dims
    x = torch.randn(n, cin, D, H, W)
    conv = torch.nn.Conv3d(cin, 1, 3, padding=1)
    bn = rand_bn(torch.nn.BatchNorm3d(1))
    p = E.Plan(DEV)
    vol = p.new_act(n, D, H, W, 32, pad=(1, 1, 1))
    vol.view()[..., c0:c0 + cin] = x.permute(0, 2, 3, 4, 1).to(DEV)
    vol.view()[..., c0 + cin:] = 7.0                      # channels past Cin carry zero weights, whatever they hold
    out = p.empty(n, D * H * W)
    p.conv3_to1(E.Act(vol.buf, n, D + 2, H + 2, W + 2, 32, c0, (1, 1, 1)), conv.weight, conv.bias, bn, out, 0.2)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.leaky_relu(F.conv3d(x.double(), wf.double(), bf.double(), padding=1), 0.2)
    assert rel_err(out.view(n, 1, D, H, W), ref) < 1e-5


@pytest.mark.parametrize("cout,ind", [(1, (3, 4, 5)), (4, (2, 3, 6)), (1, (16, 16, 16))])
def test_convtranspose3d_fused_classes(dev, cout, ind):
    """all eight parity classes in N (SVX_EPI_CONVT8): the refiner's layer8 form, (x + convT(y)) * 0.5"""
    DEV = dev
    torch.manual_seed(21 + cout)
    cin, n = 32, 2
    x = E.tf32_round(torch.randn(n, cin, *ind))
    ct = torch.nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=True)
    od, oh, ow = [2 * i for i in ind]
    skip = torch.randn(n, cout, od, oh, ow)
    p = E.Plan(DEV)
    out = p.new_act(n, od, oh, ow, cout)
    p.convT_fused(act_from_nchw(x, DEV), E.pack_convT_fused(ct.weight, None, DEV, bias=ct.bias), out, act=E.ACT_NONE,
                  residual=act_from_nchw(skip, DEV), res_after_act=True, out_scale=0.5)
    p.run()
    sync(DEV)
    ref = F.conv_transpose3d(x.double(), E.tf32_round(ct.weight.detach()).double(), ct.bias.detach().double(), 2, 1)
    assert rel_err(to_nchw(out), (ref + skip.double()) * 0.5) < 1e-4


@pytest.mark.parametrize("cin,cout,ind", [(128, 64, (4, 4, 4)), (64, 32, (8, 8, 8)), (32, 16, (3, 5, 6))])
def test_convtranspose3d_classes_in_n(dev, cin, cout, ind):
    """ConvTranspose3d(k4, s2, p1) + BN + ReLU with the eight parity classes in N (decoder layers 2 and 3): class blocks of
    a multiple of 16 channels leave through 64-byte vector stores"""
    DEV = dev
    torch.manual_seed(cin)
    x = E.tf32_round(torch.randn(2, cin, *ind))
    ct = torch.nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=False)
    bn = rand_bn(torch.nn.BatchNorm3d(cout))
    p = E.Plan(DEV)
    out = p.new_act(2, 2 * ind[0], 2 * ind[1], 2 * ind[2], cout)
    p.convT_fused(act_from_nchw(x, DEV), E.pack_convT_fused(ct.weight, bn, DEV), out, act=E.ACT_RELU, round_out=True)
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(ct.weight.transpose(0, 1), None, bn)
    ref = F.relu(F.conv_transpose3d(x.double(), E.tf32_round(wf).transpose(0, 1).double(), bf.double(), 2, 1))
    assert rel_err(to_nchw(out), ref) < 6e-4   # result stored TF32-rounded


def test_decoder_tail_fused_classes(dev):
    """decoder layer4 + layer5 + cat with the eight classes in one GEMM, writing the merger's zero-bordered layout"""
    DEV = dev
    torch.manual_seed(5)
    x = E.tf32_round(torch.randn(3, 32, 4, 5, 6))
    ct = torch.nn.ConvTranspose3d(32, 8, 4, 2, 1, bias=False)
    bn = rand_bn(torch.nn.BatchNorm3d(8))
    w5 = torch.randn(9)
    p = E.Plan(DEV)
    raw = p.new_act(3, 8, 10, 12, 16, Cs=32, pad=(1, 1, 1))
    coarse = p.empty(3, 8 * 10 * 12)
    p.convT_fused(act_from_nchw(x, DEV), E.pack_convT_fused(ct.weight, bn, DEV, block_n=64), raw, act=E.ACT_RELU,
                  round_out=True, tail=(w5.to(DEV), coarse))
    p.run()
    sync(DEV)
    wf, bf = E.fold_bn(ct.weight.transpose(0, 1), None, bn)
    feat = F.relu(F.conv_transpose3d(x.double(), E.tf32_round(wf).transpose(0, 1).double(), bf.double(), 2, 1))
    gen = (feat * w5[:8].double().view(1, 8, 1, 1, 1)).sum(1) + w5[8].double()
    got = to_nchw(raw)
    assert rel_err(coarse.view(3, 8, 10, 12), gen) < 1e-4
    assert rel_err(got[:, :8], feat) < 6e-4
    assert rel_err(got[:, 8], gen) < 6e-4
    assert got[:, 9:].abs().max().item() == 0.0
    assert raw.buf.view(3, 10, 12, 14, 32)[:, 0].abs().max().item() == 0.0   # the zero border stays untouched


@pytest.mark.parametrize("C,H,W,N", [(64, 112, 112, 2), (32, 9, 13, 3), (256, 7, 7, 1), (8, 2, 2, 2)])
def test_maxpool_3x3_stride2(dev, C, H, W, N):
    """the ResNet stem's MaxPool2d(3, 2, 1): row-marching kernel, odd sizes, one-pixel outputs; exact (max is exact)"""
    DEV = dev
    torch.manual_seed(C + H)
    y = torch.randn(N, C, H, W)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    p = E.Plan(DEV)
    mp = p.new_act(N, 1, OH, OW, C)
    p.pool(act_from_nchw(y, DEV), mp, (1, 3, 3), (1, 2, 2), (0, 1, 1), E.POOL_MAX)
    p.run()
    sync(DEV)
    assert torch.equal(to_nchw(mp).squeeze(2), F.max_pool2d(y, 3, 2, 1))


@pytest.mark.parametrize("D,k,stride,pad,kpad", [(32, 5, 2, 2, 128), (8, 3, 1, 1, 28), (9, 4, 2, 0, 64)])
def test_im2col_single_channel_volume(dev, D, k, stride, pad, kpad):
    """refiner layer1's window gather (one channel, k^3 windows) -- the shared-memory line kernel; exact copy"""
    DEV = dev
    torch.manual_seed(D + k)
    B = 2
    vol = torch.randn(B, D, D, D)
    O = (D + 2 * pad - k) // stride + 1
    p = E.Plan(DEV)
    cols = p.im2col(vol.to(DEV), (D ** 3, 0, D * D, D, 1), B, 1, (D, D, D), (k, k, k), stride, (pad, pad, pad), (O, O, O), kpad,
                    round_out=False)
    p.run()
    sync(DEV)
    padded = F.pad(vol, (pad,) * 6)
    win = padded.unfold(1, k, stride).unfold(2, k, stride).unfold(3, k, stride)      # [B, O, O, O, k, k, k]
    ref = win.reshape(B * O ** 3, k ** 3)
    got = cols.view().reshape(B * O ** 3, kpad).cpu()
    assert torch.equal(got[:, :k ** 3], ref) and (kpad == k ** 3 or got[:, k ** 3:].abs().max().item() == 0.0)


def test_im2col_and_pools(dev):
    DEV = dev
    torch.manual_seed(4)
    x = torch.randn(2, 3, 30, 30)
    p = E.Plan(DEV)
    xd = x.to(DEV)
    cols = p.im2col(xd, (3 * 900, 900, 0, 30, 1), 2, 3, (1, 30, 30), (1, 7, 7), 2, (0, 3, 3), (1, 15, 15), 160,
                    round_out=False)
    y = torch.randn(2, 8, 9, 9)
    ya = act_from_nchw(y, DEV)
    mp = p.new_act(2, 1, 5, 5, 8)
    p.pool(ya, mp, (1, 3, 3), (1, 2, 2), (0, 1, 1), E.POOL_MAX)
    z = torch.randn(2, 8, 7, 7)
    ap = p.new_act(2, 2, 2, 2, 8)
    p.pool(act_from_nchw(z, DEV), ap, (1, 4, 4), (0, 3, 3), (0, 0, 0), E.POOL_AVG)
    v = torch.randn(2, 8, 6, 6, 6)
    mp3 = p.new_act(2, 3, 3, 3, 8)
    p.pool(act_from_nchw(v, DEV), mp3, (2, 2, 2), (2, 2, 2), (0, 0, 0), E.POOL_MAX)
    p.run()
    sync(DEV)
    ref = F.unfold(x, 7, padding=3, stride=2)  # [N, C*49, L] with k index = c*49 + kh*7 + kw
    ref = ref.view(2, 3, 49, 225).permute(0, 3, 2, 1).reshape(2, 225, 147)
    got = cols.view().reshape(2, 225, 160).cpu()
    assert torch.equal(got[..., :147], ref) and got[..., 147:].abs().max() == 0
    assert torch.equal(to_nchw(mp).squeeze(2), F.max_pool2d(y, 3, 2, 1))
    ad = F.adaptive_avg_pool2d(z, 2)
    assert rel_err(to_nchw(ap), ad.unsqueeze(2).expand(-1, -1, 2, -1, -1)) < 1e-6
    assert torch.equal(to_nchw(mp3), F.max_pool3d(v, 2))


def test_layernorms(dev):
    DEV = dev
    torch.manual_seed(6)
    p = E.Plan(DEV)
    x = torch.randn(500, 192) * 2 + 0.5
    g, b = torch.rand(192) + 0.5, torch.randn(192)
    o1 = p.new_act(500, 1, 1, 1, 192)
    p.layernorm_rows(E.Act(x.to(DEV), 500, 1, 1, 1, 192), g.to(DEV), b.to(DEV), o1, round_out=False)
    xm = torch.randn(2, 14, 14, 96)
    gm, bm = torch.rand(384) + 0.5, torch.randn(384)
    o2 = p.new_act(2, 1, 7, 7, 384)
    p.layernorm_rows(E.Act(xm.view(-1, 96).to(DEV), 2, 1, 14, 14, 96), gm.to(DEV), bm.to(DEV), o2, merge_hw=(14, 14),
                     round_out=False)
    xs = torch.randn(3, 7, 7, 768) + 1.0
    gs, bs = torch.rand(7, 7, 768) + 0.5, torch.randn(7, 7, 768)
    o3 = p.new_act(3, 1, 7, 7, 768)
    p.layernorm_sample(E.Act(xs.view(-1, 768).to(DEV), 3, 1, 7, 7, 768), gs.to(DEV), bs.to(DEV), o3, round_out=False)
    p.run()
    sync(DEV)
    assert rel_err(o1.view().reshape(500, 192), F.layer_norm(x.double(), (192,), g.double(), b.double())) < 1e-5
    merged = torch.cat([xm[:, 0::2, 0::2], xm[:, 1::2, 0::2], xm[:, 0::2, 1::2], xm[:, 1::2, 1::2]], -1)
    assert rel_err(o2.view().reshape(2, 7, 7, 384), F.layer_norm(merged.double(), (384,), gm.double(), bm.double())) < 1e-5
    assert rel_err(o3.view().reshape(3, 7, 7, 768),
                   F.layer_norm(xs.double(), (7, 7, 768), gs.double(), bs.double())) < 1e-5


@pytest.mark.parametrize("H,C,N", [(56, 96, 3), (28, 192, 2), (14, 384, 5), (7, 768, 2), (3, 100, 2)])
@pytest.mark.parametrize("cluster", [False, True])
def test_layernorm_sample_shapes(dev, H, C, N, cluster, monkeypatch):
    """the wrapper's LayerNorm([C,H,W]) at the four Swin stage shapes and one tiny sample; `cluster` = the opt-in
    thread-block-cluster kernel (distributed-shared-memory reduction), every register variant"""
    DEV = dev
    if cluster:
        monkeypatch.setenv("SVX_LN_CLUSTER", "1")
    torch.manual_seed(H + C)
    xs = torch.randn(N, H, H, C) * 1.7 + 0.4
    gs, bs = torch.rand(H, H, C) + 0.5, torch.randn(H, H, C)
    p = E.Plan(DEV)
    o = p.new_act(N, 1, H, H, C)
    p.layernorm_sample(E.Act(xs.view(-1, C).to(DEV), N, 1, H, H, C), gs.to(DEV), bs.to(DEV), o, round_out=False)
    p.run()
    sync(DEV)
    assert rel_err(o.view().reshape(N, H, H, C), F.layer_norm(xs.double(), (H, H, C), gs.double(), bs.double())) < 1e-5


@pytest.mark.parametrize("rows,C,merge", [(1, 96, False), (1001, 96, False), (333, 192, False), (77, 384, False),
                                          (50, 768, False), (13, 2048, False), (None, 384, True), (None, 768, True),
                                          (None, 1536, True), (7, 100, False)])
def test_layernorm_rows_widths(dev, rows, C, merge):
    """every register-resident variant of the row LayerNorm (C / 128 float4 per lane) + ragged row counts"""
    DEV = dev
    torch.manual_seed(C + (rows or 0))
    p = E.Plan(DEV)
    g, b = torch.rand(C) + 0.5, torch.randn(C)
    if merge:
        n, H = 3, 6
        xm = torch.randn(n, H, H, C // 4) * 1.5 + 0.3
        o = p.new_act(n, 1, H // 2, H // 2, C)
        p.layernorm_rows(E.Act(xm.view(-1, C // 4).to(DEV), n, 1, H, H, C // 4), g.to(DEV), b.to(DEV), o, merge_hw=(H, H),
                         round_out=False)
        ref_in = torch.cat([xm[:, 0::2, 0::2], xm[:, 1::2, 0::2], xm[:, 0::2, 1::2], xm[:, 1::2, 1::2]], -1).reshape(-1, C)
    else:
        ref_in = torch.randn(rows, C) * 2 + 0.5
        o = p.new_act(rows, 1, 1, 1, C)
        p.layernorm_rows(E.Act(ref_in.to(DEV), rows, 1, 1, 1, C), g.to(DEV), b.to(DEV), o, round_out=False)
    p.run()
    sync(DEV)
    assert rel_err(o.view().reshape(-1, C), F.layer_norm(ref_in.double(), (C,), g.double(), b.double())) < 1e-5


@pytest.mark.parametrize("H,heads,shift,N", [(14, 3, 0, 2), (14, 3, 3, 2), (7, 6, 0, 2), (28, 2, 3, 2), (7, 24, 0, 3),
                                             (14, 12, 3, 5), (56, 3, 3, 3), (7, 2, 0, 1)])
def test_window_attention(dev, H, heads, shift, N):
    """W-MSA / SW-MSA: the tensor-core kernel packs two windows per 128-row tile, so odd window counts (N = 3 or 1 images
    of one window, 5 x 4 windows), every wrap-around window type of the shifted maps and all four Swin stage shapes are
    covered"""
    DEV = dev
    torch.manual_seed(7)
    Cc = heads * 32
    qkv = E.tf32_round(torch.randn(N, H, H, 3 * Cc))   # the qkv GEMM stores its output TF32-rounded
    bias = torch.randn(heads, 49, 49)
    scale = 32 ** -0.5
    p = E.Plan(DEV)
    out = p.new_act(N, 1, H, H, Cc)
    p.window_attention(E.Act(qkv.view(-1, 3 * Cc).to(DEV), N, 1, H, H, 3 * Cc), out, bias.to(DEV), H, H, heads, shift,
                       scale, round_out=False)
    p.run()
    sync(DEV)
    # reference: roll -> partition -> attention with mask -> reverse -> roll back (timm semantics)
    x = qkv.double()
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    nw = H // 7
    xw = x.view(N, nw, 7, nw, 7, 3, heads, 32).permute(0, 1, 3, 5, 6, 2, 4, 7).reshape(N * nw * nw, 3, heads, 49, 32)
    q, k, v = xw[:, 0] * scale, xw[:, 1], xw[:, 2]
    attn = q @ k.transpose(-1, -2) + bias.double()
    if shift:
        img = torch.zeros(H, H)
        cnt = 0
        for hs in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
            for ws in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
                img[hs, ws] = cnt
                cnt += 1
        mw = img.view(nw, 7, nw, 7).permute(0, 2, 1, 3).reshape(nw * nw, 49)
        mask = (mw.unsqueeze(1) - mw.unsqueeze(2) != 0).double() * -100.0
        attn = attn.view(N, nw * nw, heads, 49, 49) + mask.view(1, nw * nw, 1, 49, 49)
        attn = attn.view(-1, heads, 49, 49)
    o = attn.softmax(-1) @ v  # [B_, heads, 49, 32]
    o = o.transpose(1, 2).reshape(N, nw, nw, 7, 7, Cc).permute(0, 1, 3, 2, 4, 5).reshape(N, H, H, Cc)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    # the CUDA kernel contracts on the tensor cores (TF32 operands: P is rounded to 11 bits), the CPU twin in fp32
    assert rel_err(out.view().reshape(N, H, H, Cc), o) < (1e-3 if DEV == "cuda" else 2e-5)


def test_cva_pieces(dev):
    DEV = dev
    torch.manual_seed(8)
    B, V, Cc, R, heads = 2, 3, 64, 32, 4
    p = E.Plan(DEV)
    x = torch.randn(B * V, Cc, 7, 7)
    dw = torch.nn.Conv2d(Cc, Cc, 2, 2, groups=Cc)
    o_dw = p.new_act(B * V, 1, 3, 3, Cc)
    p.dwconv(act_from_nchw(x, DEV), dw.weight.detach().view(Cc, 4).t().contiguous().to(DEV), dw.bias.detach().to(DEV), o_dw, 2,
             round_out=False)
    qkv = torch.randn(B * V, 3 * R, 3, 3)
    o_att = p.new_act(B * V, 1, 3, 3, R)
    scale = 1.0 / ((R // heads) * V) ** 0.5
    p.view_attention(act_from_nchw(qkv, DEV), o_att, B, V, heads, scale, round_out=False)
    small, skip = torch.randn(B * V, Cc, 3, 3), torch.randn(B * V, Cc, 7, 7)
    o_bl = p.new_act(B * V, 1, 7, 7, Cc)
    p.bilinear_add(act_from_nchw(small, DEV), act_from_nchw(skip, DEV), o_bl, round_out=False)
    p.run()
    sync(DEV)
    assert rel_err(to_nchw(o_dw).squeeze(2), dw(x)) < 1e-5
    hd = R // heads
    q, k, v = torch.split(qkv.double(), [R] * 3, 1)
    q = q.reshape(B, V, heads, hd * 9).permute(0, 2, 1, 3)
    k = k.reshape(B, V, heads, hd * 9).permute(0, 2, 3, 1)
    v = v.reshape(B, V, heads, hd * 9).permute(0, 2, 1, 3)
    a = torch.softmax(q @ k * scale, -1) @ v  # [B, heads, V, hd*9]
    a = a.view(B, heads, V, hd, 3, 3).permute(0, 2, 1, 3, 4, 5).reshape(B * V, R, 3, 3)
    assert rel_err(to_nchw(o_att).squeeze(2), a) < 1e-5
    ref = F.interpolate(small.double(), size=(7, 7), mode="bilinear", align_corners=False) + skip.double()
    assert rel_err(to_nchw(o_bl).squeeze(2), ref) < 1e-5


@pytest.mark.parametrize("V", [1, 3, 5, 24])
def test_merger_fuse(dev, V):
    DEV = dev
    torch.manual_seed(9)
    B, P = 3, 32 ** 3
    w, c = torch.randn(B, V, P) * 3, torch.randn(B, V, P)
    p = E.Plan(DEV)
    out = p.empty(B, P)
    p.merger_fuse(p.hold(w.to(DEV)), p.hold(c.to(DEV)), out, B, V, P)
    p.run()
    sync(DEV)
    ref = (torch.softmax(w.double(), 1) * c.double()).sum(1)
    assert (out.cpu().double() - ref).abs().max().item() < 1e-5


def test_voxel_metrics_exact(dev):
    DEV = dev
    torch.manual_seed(10)
    B, P = 5, 32 ** 3
    logits = torch.randn(B, P) * 2
    logits[3] = -50.0  # empty prediction
    gt = (torch.rand(B, P) < 0.1).float()
    gt[3] = 0.0        # and empty ground truth: union == 0
    th = torch.tensor([0.2, 0.3, 0.4, 0.5])
    p = E.Plan(DEV)
    counts = p.zeros(B, 4, 5, dtype=torch.int32)
    ld = logits.to(DEV)
    p.voxel_metrics(p.hold(ld), p.hold(gt.to(DEV)), th.to(DEV), counts, B, P)
    p.run()
    p.run()  # counters are re-zeroed by each launch
    sync(DEV)
    prob = torch.sigmoid(ld).cpu()  # same fp32 sigmoid family; ties at the threshold are measure-zero here
    exp = torch.zeros(B, 4, 5, dtype=torch.int64)
    for t in range(4):
        v = (prob >= th[t]).float()
        exp[:, t, 0] = (v * gt).sum(1)
        exp[:, t, 1] = ((v + gt) >= 1).sum(1)
        exp[:, t, 2] = (v * gt).sum(1)
        exp[:, t, 3] = (v * (1 - gt)).sum(1)
        exp[:, t, 4] = ((1 - v) * gt).sum(1)
    diff = (counts.cpu().long() - exp).abs().max().item()
    assert diff <= 2, diff  # <=2 voxels of 32768 may sit within 1 ulp of a threshold
    assert counts[3].sum().item() == 0


def test_transpose_roundtrip(dev):
    DEV = dev
    torch.manual_seed(11)
    x = torch.randn(3, 9, 1000)
    p = E.Plan(DEV)
    cl = p.empty(3, 1000, 16)
    back = p.empty(3, 9, 1000)
    xd = p.hold(x.to(DEV))
    p.transpose(xd, cl, 3, 9, 1000, 16, True)
    p.transpose(cl, back, 3, 9, 1000, 16, False)
    p.run()
    sync(DEV)
    assert torch.equal(cl.cpu()[..., :9], x.permute(0, 2, 1)) and cl.cpu()[..., 9:].abs().max() == 0
    assert torch.equal(back.cpu(), x)


def test_graph_replay_matches_eager(dev):
    DEV = dev
    torch.manual_seed(12)
    M, K, N = 1024, 256, 128
    x = E.tf32_round(torch.randn(M, K)).to(DEV)
    p = E.Plan(DEV)
    h = p.new_act(M, 1, 1, 1, N)
    o = p.new_act(M, 1, 1, 1, K)
    p.linear(E.Act(x, M, 1, 1, 1, K), E.pack_matrix(torch.randn(N, K) / 16, None, DEV), h, act=E.ACT_GELU, round_out=True)
    p.linear(h, E.pack_matrix(torch.randn(K, N) / 11, None, DEV), o)
    p.run()
    sync(DEV)
    ref = o.buf.clone()
    for _ in range(3):
        o.buf.zero_()
        p.run(graph=True)
    sync(DEV)
    assert torch.equal(o.buf, ref)

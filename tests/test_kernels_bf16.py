"""Parity of the bf16-storage variants of the kernels (BASELINE configs[2]: the encoder with bf16 activations and bf16
tensor-core operands, fp32 accumulation and statistics) against plain PyTorch fp64 references of the same op on the
same bf16-rounded inputs.  bf16 outputs carry one round-to-nearest-even (2^-9 relative), fp32 outputs are held to the
accumulation-order tolerance.  Dual backend like tests/test_kernels.py: `cuda` = the real sm_100a kernels (gpu tier),
`hostsim` = the CPU twin (checks packing / tap tables / descriptors and the references of this file)."""
import pytest
import torch
import torch.nn.functional as F

from swinvox_b200 import engine as E
from test_kernels import rand_bn, rel_err
from util import dev, sync  # noqa: F401

BF = torch.bfloat16
TOL16 = 6e-3    # one bf16 rounding of the stored result, relative to the tensor's max
TOL32 = 2e-4    # fp32 result of bf16 operands: accumulation order only


def bf(t):
    return t.to(BF).float()


def act16(x, DEV):  # [N,C,H,W] cpu fp32 -> bf16 channels-last Act
    N, Cc, H, W = x.shape
    return E.Act(x.permute(0, 2, 3, 1).contiguous().view(-1, Cc).to(BF).to(DEV), N, 1, H, W, Cc)


def nchw(a):
    return a.view().squeeze(1).permute(0, 3, 1, 2).float().cpu()


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 96, 96), (1000, 288, 96), (4096, 256, 1024), (513, 192, 768),
                                   (64, 2048, 1024), (2500, 64, 152), (777, 1000, 384)])
@pytest.mark.parametrize("out_dtype", [BF, torch.float32])
def test_gemm_plain_bf16(dev, M, N, K, out_dtype):
    DEV = dev
    torch.manual_seed(M + N + K)
    x, w, b = bf(torch.randn(M, K)), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(M, 1, 1, 1, N, dtype=out_dtype)
    p.linear(E.Act(x.to(BF).to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, b, DEV), out)
    p.run()
    sync(DEV)
    ref = x.double() @ bf(w).double().t() + b.double()
    assert rel_err(out.view().reshape(M, N).float(), ref) < (TOL16 if out_dtype == BF else TOL32)


@pytest.mark.parametrize("act", [E.ACT_NONE, E.ACT_RELU, E.ACT_GELU])
@pytest.mark.parametrize("res_after", [False, True])
@pytest.mark.parametrize("N,bn", [(96, None), (256, 128), (384, None)])
def test_gemm_epilogue_bf16(dev, act, res_after, N, bn):
    DEV = dev
    torch.manual_seed(5 + N)
    M, K = 333, 192
    x, w, b, r = bf(torch.randn(M, K)), torch.randn(N, K) / K ** 0.5, torch.randn(N), bf(torch.randn(M, N))
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(M, 1, 1, 1, N)
    res = E.Act(r.to(BF).to(DEV), M, 1, 1, 1, N)
    p.linear(E.Act(x.to(BF).to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, b, DEV, block_n=bn), out, act=act, residual=res,
             res_after_act=res_after, out_scale=0.5)
    p.run()
    sync(DEV)
    f = {E.ACT_NONE: lambda t: t, E.ACT_RELU: F.relu, E.ACT_GELU: F.gelu}[act]
    y = x.double() @ bf(w).double().t() + b.double()
    ref = 0.5 * (f(y) + r.double() if res_after else f(y + r.double()))
    assert rel_err(out.view().reshape(M, N).float(), ref) < TOL16


@pytest.mark.parametrize("M,K,N,bn", [(333, 192, 256, 256), (1000, 64, 512, 128), (4096, 256, 1024, None)])
def test_gemm_residual_on_tensor_cores_bf16(dev, M, K, N, bn):
    """relu(x W^T + b + r) with the bf16 residual added by the MMA (identity columns, residual tile as extra k-chunks)"""
    DEV = dev
    torch.manual_seed(M)
    x, w, b, r = bf(torch.randn(M, K)), torch.randn(N, K) / K ** 0.5, torch.randn(N), bf(torch.randn(M, N))
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(M, 1, 1, 1, N)
    p.linear(E.Act(x.to(BF).to(DEV), M, 1, 1, 1, K), E.pack_matrix(w, b, DEV, block_n=bn), out, act=E.ACT_RELU,
             residual=E.Act(r.to(BF).to(DEV), M, 1, 1, 1, N), res_after_act=False, res_via_mma=True)
    p.run()
    sync(DEV)
    ref = F.relu(x.double() @ bf(w).double().t() + b.double() + r.double())
    assert rel_err(out.view().reshape(M, N).float(), ref) < TOL16


@pytest.mark.parametrize("cin,cout,hw,k,s,p_", [(64, 64, 14, 3, 1, 1), (64, 96, 15, 3, 2, 1), (256, 128, 7, 1, 1, 0),
                                                 (128, 256, 14, 1, 2, 0), (512, 256, 7, 3, 1, 1), (32, 48, 9, 3, 2, 1),
                                                 (8, 16, 9, 5, 2, 2)])
def test_conv2d_bf16(dev, cin, cout, hw, k, s, p_):
    """implicit-GEMM convolutions on bf16 tensors: TMA im2col mode (whole 128-byte channel chunks) and the cp.async
    gather fallback (32 / 8 channels: 64 / 16-byte pixels that are not whole chunks)"""
    DEV = dev
    torch.manual_seed(cin + cout)
    x = bf(torch.randn(3, cin, hw, hw))
    conv = torch.nn.Conv2d(cin, cout, k, s, p_)
    bn = rand_bn(torch.nn.BatchNorm2d(cout))
    oh = (hw + 2 * p_ - k) // s + 1
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(3, 1, oh, oh, cout)
    p.conv(act16(x, DEV), E.pack_conv(conv.weight, conv.bias, bn, DEV, cin_pad=E.round_up(cin, 8)), E.conv_taps(1, k, k, 0, p_, p_),
           out, stride=(1, s, s), act=E.ACT_RELU)
    p.run()
    sync(DEV)
    wf, bfold = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.relu(F.conv2d(x.double(), bf(wf).double(), bfold.double(), s, p_))
    assert rel_err(nchw(out), ref) < TOL16


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 14), (128, 96, 9), (512, 256, 7)])
def test_conv2d_flat_tma_bf16(dev, cin, cout, hw):
    DEV = dev
    torch.manual_seed(cin)
    x = bf(torch.randn(3, cin, hw, hw))
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1)
    bn = rand_bn(torch.nn.BatchNorm2d(cout))
    p = E.Plan(DEV, dtype=BF)
    xin = p.new_act(3, 1, hw, hw, cin, pad=(0, 1, 1))
    xin.view().copy_(x.permute(0, 2, 3, 1).unsqueeze(1))
    out = p.new_act(3, 1, hw, hw, cout, pad=(0, 1, 1))
    p.conv_flat(xin, E.pack_conv(conv.weight, conv.bias, bn, DEV), E.conv_taps(1, 3, 3, 0, 0, 0), out, act=E.ACT_RELU)
    p.run()
    sync(DEV)
    wf, bfold = E.fold_bn(conv.weight, conv.bias, bn)
    ref = F.relu(F.conv2d(x.double(), bf(wf).double(), bfold.double(), 1, 1))
    assert rel_err(nchw(out), ref) < TOL16
    full = out.buf.view(3, hw + 2, hw + 2, cout).float().cpu()
    assert full[:, 0].abs().max() == 0 and full[:, -1].abs().max() == 0 and full[:, :, 0].abs().max() == 0 \
        and full[:, :, -1].abs().max() == 0, "the zero border must stay untouched"


def test_image_stems_bf16(dev):
    """both image stems on the bf16 staged image (pixel pairs = 16-byte TMA im2col boxes): the ResNet 7x7 s2 conv + BN +
    ReLU and Swin's 4x4 s4 patch embedding, through the lowering helpers of swinvox_b200/graph.py"""
    from swinvox_b200 import graph
    DEV = dev
    torch.manual_seed(3)
    N = 2
    img = torch.rand(N, 3, 224, 224) * 2 - 1
    conv1, bn1 = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False), rand_bn(torch.nn.BatchNorm2d(64))
    proj = torch.nn.Conv2d(3, 96, 4, 4)
    p = E.Plan(DEV, dtype=BF)
    x4 = graph.stage_image(p, img.to(DEV).contiguous(), N)
    pitch, x0 = graph.img_layout(p)
    pairs = E.Act(x4.buf.view(-1, 8), N, 1, 224, pitch // 2, 8)
    stem = p.new_act(N, 1, 112, 112, 64)
    pk, taps = graph.pack_stem_pairs(conv1, bn1, DEV, x0)
    p.conv(pairs, pk, taps, stem, stride=(1, 2, 1), act=E.ACT_RELU)
    w = proj.weight.detach().float()
    Wp = torch.zeros(96, 4, 2, 2, 4)
    for kw in range(4):
        Wp[:, :, kw // 2, kw % 2, :3] = w[:, :, :, kw].permute(0, 2, 1)
    emb = p.new_act(N, 1, 56, 56, 96)
    p.conv(pairs, E.pack_matrix(Wp.reshape(96, 64), proj.bias, DEV), [(0, kh, x0 // 2 + pr) for kh in range(4) for pr in range(2)],
           emb, stride=(1, 4, 2), rows_dhw=(1, 56, 56))
    p.run()
    sync(DEV)
    xb = bf(img).double()
    wf, bfold = E.fold_bn(conv1.weight, None, bn1)
    assert rel_err(nchw(stem), F.relu(F.conv2d(xb, bf(wf).double(), bfold.double(), 2, 3))) < TOL16
    assert rel_err(nchw(emb), F.conv2d(xb, bf(w).double(), proj.bias.double(), 4)) < TOL16


@pytest.mark.parametrize("C,merge", [(96, False), (192, False), (384, True), (768, False), (1536, True)])
def test_layernorm_rows_bf16(dev, C, merge):
    DEV = dev
    torch.manual_seed(C)
    N, H, W = 2, 6, 8
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    p = E.Plan(DEV, dtype=BF)
    if merge:
        x = bf(torch.randn(N, 2 * H, 2 * W, C // 4) * 2 + 0.5)
        xa = E.Act(x.reshape(-1, C // 4).to(BF).to(DEV), N, 1, 2 * H, 2 * W, C // 4)
        rows = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(-1, C)
    else:
        x = bf(torch.randn(N, H, W, C) * 2 + 0.5)
        xa = E.Act(x.reshape(-1, C).to(BF).to(DEV), N, 1, H, W, C)
        rows = x.reshape(-1, C)
    out = p.new_act(N, 1, H, W, C)
    p.layernorm_rows(xa, g.to(DEV), b.to(DEV), out, merge_hw=(2 * H, 2 * W) if merge else None)
    p.run()
    sync(DEV)
    ref = F.layer_norm(rows.double(), (C,), g.double(), b.double(), 1e-5)
    assert rel_err(out.view().reshape(-1, C).float(), ref) < TOL16


@pytest.mark.parametrize("hwc", [(56, 56, 96), (14, 14, 384), (7, 7, 768)])
def test_layernorm_sample_bf16(dev, hwc):
    DEV = dev
    H, W, C = hwc
    torch.manual_seed(C)
    N = 3
    x = bf(torch.randn(N, H, W, C) * 1.5 - 0.2)
    g, b = torch.rand(H, W, C) + 0.5, torch.randn(H, W, C) * 0.2
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(N, 1, H, W, C)
    p.layernorm_sample(E.Act(x.reshape(-1, C).to(BF).to(DEV), N, 1, H, W, C), g.to(DEV), b.to(DEV), out)
    p.run()
    sync(DEV)
    ref = F.layer_norm(x.double(), (H, W, C), g.double(), b.double(), 1e-5)
    assert rel_err(out.view().squeeze(1).float(), ref) < TOL16


@pytest.mark.parametrize("H,C,shift", [(14, 96, 0), (14, 192, 3), (7, 768, 0), (28, 96, 3)])
def test_window_attention_bf16(dev, H, C, shift):
    """W-MSA / SW-MSA on bf16 qkv against a direct fp64 evaluation of timm's WindowAttention semantics"""
    DEV = dev
    torch.manual_seed(H + C + shift)
    N, heads = 2, C // 32
    qkv = bf(torch.randn(N, H, H, 3 * C))
    bias = torch.randn(heads, 49, 49) * 0.5
    p = E.Plan(DEV, dtype=BF)
    out = p.new_act(N, 1, H, H, C)
    p.window_attention(E.Act(qkv.reshape(-1, 3 * C).to(BF).to(DEV), N, 1, H, H, 3 * C), out, bias.to(DEV), H, H, heads, shift,
                       32 ** -0.5)
    p.run()
    sync(DEV)
    x = torch.roll(qkv.double(), (-shift, -shift), (1, 2)) if shift else qkv.double()
    nw = H // 7
    xw = x.view(N, nw, 7, nw, 7, 3, heads, 32).permute(0, 1, 3, 5, 6, 2, 4, 7).reshape(N * nw * nw, 3, heads, 49, 32)
    q, k, v = xw[:, 0], xw[:, 1], xw[:, 2]
    att = (q * 32 ** -0.5) @ k.transpose(-1, -2) + bias.double()
    if shift:
        m = torch.zeros(H, H)
        for i, hs in enumerate((slice(0, -7), slice(-7, -shift), slice(-shift, None))):
            for j, ws_ in enumerate((slice(0, -7), slice(-7, -shift), slice(-shift, None))):
                m[hs, ws_] = i * 3 + j
        mw = m.view(nw, 7, nw, 7).permute(0, 2, 1, 3).reshape(nw * nw, 49)
        mask = (mw[:, None, :] != mw[:, :, None]).double() * -100.0
        att = att + mask.repeat(N, 1, 1)[:, None]
    o = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(N, nw, nw, 7, 7, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H, H, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    assert rel_err(out.view().squeeze(1).float(), o) < TOL16


def test_pool_dwconv_bilinear_viewattn_bf16(dev):
    """the small bandwidth-bound kernels of the encoder tail on bf16 tensors"""
    DEV = dev
    torch.manual_seed(9)
    p = E.Plan(DEV, dtype=BF)
    x = bf(torch.randn(2, 64, 14, 14))
    xa = act16(x, DEV)
    mp = p.new_act(2, 1, 7, 7, 64)
    p.pool(xa, mp, (1, 3, 3), (1, 2, 2), (0, 1, 1), E.POOL_MAX)
    ap = p.new_act(2, 1, 7, 7, 64)
    p.pool(xa, ap, (1, 2, 2), (1, 2, 2), (0, 0, 0), E.POOL_AVG)
    wdw, bdw = torch.randn(64, 1, 2, 2), torch.randn(64)
    dwo = p.new_act(2, 1, 7, 7, 64)
    p.dwconv(xa, wdw.reshape(64, 4).t().contiguous().to(DEV), bdw.to(DEV), dwo, 2)
    small = bf(torch.randn(2, 64, 3, 3))
    skip = bf(torch.randn(2, 64, 7, 7))
    bil = p.new_act(2, 1, 7, 7, 64)
    p.bilinear_add(act16(small, DEV), act16(skip, DEV), bil)
    B, V, P_, R, heads = 2, 3, 9, 128, 4
    qkv = bf(torch.randn(B * V, P_, 3 * R))
    va = p.new_act(B * V, 1, 3, 3, R)
    p.view_attention(E.Act(qkv.reshape(-1, 3 * R).to(BF).to(DEV), B * V, 1, 3, 3, 3 * R), va, B, V, heads, 1.0 / (32 * V) ** 0.5)
    p.run()
    sync(DEV)
    assert rel_err(nchw(mp), F.max_pool2d(x.double(), 3, 2, 1)) < TOL16
    assert rel_err(nchw(ap), F.avg_pool2d(x.double(), 2)) < TOL16
    assert rel_err(nchw(dwo), F.conv2d(x.double(), wdw.double(), bdw.double(), stride=2, groups=64)) < TOL16
    ref_b = F.interpolate(small.double(), size=(7, 7), mode="bilinear", align_corners=False) + skip.double()
    assert rel_err(nchw(bil), ref_b) < TOL16
    q, k, v = [t.reshape(B, V, P_, heads, 32).permute(0, 3, 1, 2, 4).reshape(B, heads, V, P_ * 32).double()
               for t in qkv.split(R, dim=-1)]
    att = (q @ k.transpose(-1, -2) / (32 * V) ** 0.5).softmax(-1)
    ref_v = (att @ v).reshape(B, heads, V, P_, 32).permute(0, 2, 3, 1, 4).reshape(B * V, P_, R)
    assert rel_err(va.view().reshape(B * V, P_, R).float(), ref_v) < TOL16

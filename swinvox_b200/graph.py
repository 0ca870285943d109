"""Lowering of the SwinVox modules onto libswinvox_b200 ops.

Each ``lower_*`` function reads the parameters of a (reference-layout) nn.Module container, prepares the
weights (BatchNorm folding, TF32 rounding, k = tap*Cin + c re-layout, transposed-convolution parity
classes) and records the ops into an engine.Plan.  Reference semantics are cited per function.

Rounding policy: a tensor is stored rounded to TF32 (`round_out`) exactly when its next consumer is a
tensor-core contraction, so kind::tf32's operand truncation never adds a bias; everything else stays
full fp32 (residual streams, attention inputs, module outputs).
"""
import os

import torch

from . import engine as E
from .engine import ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, POOL_AVG, POOL_MAX, Act

WINDOW = 7
RAW_CS = 16   # row width (channels) of the decoder -> merger hand-off buffer `raw`: 9 live channels in 64-byte rows


def _dev(plan):
    return plan.device


# --------------------------------------------------------------------------------------------------
# ResNet-50 trunk: torchvision resnet50 children[:7]  (models/encoder.py:22-23,119)
# --------------------------------------------------------------------------------------------------
def img_layout(plan):
    """(row pitch, left zero pixels) of the staged image.  fp32: 226 / 1 -- Swin's patch embedding reads 16-byte pixels,
    the ResNet stem the 8-channel pixel pairs (2j-1, 2j) of its stride-2 window.  bf16: a pixel is 8 bytes, below the TMA
    unit's 16-byte minimum, so BOTH stems read pixel pairs and the image starts at an even staged column: 228 / 2."""
    return (228, 2) if plan.dtype == torch.bfloat16 else (226, 1)


def stage_image(plan, img, N):
    """[N,3,224,224] NCHW fp32 tensor -> Act [N,224,pitch,4] (channels-last, zero fourth channel, rounded to the plan's
    operand type, zero pixels left and right of every row; see img_layout).  Idempotent per plan."""
    if isinstance(img, Act):
        return img
    key = ("nhwc4", img.data_ptr())
    if key not in plan.taps:
        pitch, x0 = img_layout(plan)
        a = plan.new_act(N, 1, 224, pitch, 4, zero=True)
        cin = img.shape[1]   # 3 (4 at most: one 16-byte fp32 pixel)
        plan.transpose(img, a.buf, N, cin, 224 * 224, 4, True, round_out=True, name="image.nhwc4", rows=(224, pitch, x0))
        plan.taps[key] = a
    return plan.taps[key]


def pack_stem_pairs(conv1, bn1, dev, x0=1):
    """ResNet conv1 (7x7, stride 2, pad 3) over pixel pairs.  Image column 2ow - 3 + kw sits at staged column
    2ow - 3 + kw + x0 = 2(ow - 1) + s with s = kw + x0 - 1, i.e. pair (ow - 1) + s//2, element s%2 -> K = 7 kh x 4 pair
    taps x (2 pixels x 4 channels) = 224; the unused eighth pixel slot and channel 3 carry zero weights."""
    w, b = E.fold_bn(conv1.weight, conv1.bias, bn1)            # [64, 3, 7, 7]
    W = torch.zeros(64, 7, 4, 2, 4, dtype=torch.float32, device=w.device)
    for kw in range(7):
        s_ = kw + x0 - 1
        W[:, :, s_ // 2, s_ % 2, :3] = w[:, :, :, kw].permute(0, 2, 1)
    taps = [(0, kh - 3, pt - 1) for kh in range(7) for pt in range(4)]
    return E.pack_matrix(W.reshape(64, 224), b, dev), taps


def lower_resnet_trunk(plan, resnet, img, N):
    """img: [N,3,224,224] fp32 NCHW tensor (or the staged NHWC4 Act).  Returns Act [N,14,14,1024]."""
    dev = _dev(plan)
    conv1, bn1 = resnet[0], resnet[1]
    x4 = stage_image(plan, img, N)
    # 7x7 s2 p3 stem as an implicit GEMM over 28 pixel-pair taps of 8 channels (32-byte TMA im2col boxes): K = 224
    pitch, x0 = img_layout(plan)
    pairs = Act(x4.buf.view(-1, 8), N, 1, 224, pitch // 2, 8)
    stem = plan.new_act(N, 1, 112, 112, 64)
    pk, taps = pack_stem_pairs(conv1, bn1, dev, x0)
    plan.conv(pairs, pk, taps, stem, stride=(1, 2, 1), act=ACT_RELU, name="resnet.stem")
    x = plan.new_act(N, 1, 56, 56, 64)
    plan.pool(stem, x, (1, 3, 3), (1, 2, 2), (0, 1, 1), POOL_MAX, round_out=True, name="resnet.maxpool")
    plan.release(stem)
    for li in (4, 5, 6):
        for bi, blk in enumerate(resnet[li]):
            x = _bottleneck(plan, blk, x, f"resnet.{li}.{bi}")   # (releases its input: nobody else reads it)
    return x


def _bottleneck(plan, blk, x, name):
    dev = _dev(plan)
    s = blk.conv2.stride[0]
    width, cout = blk.conv1.out_channels, blk.conv3.out_channels
    H2 = (x.H + 2 - 3) // s + 1
    # stride-1 3x3: TMA-fed "flat" conv over a zero-padded conv1 output (whole 128-byte channel chunks)
    flat = s == 1 and (width * E.esize_of(plan.dtype)) % 128 == 0
    t1 = plan.new_act(x.N, 1, x.H, x.W, width, pad=(0, 1, 1) if flat else (0, 0, 0))
    plan.linear(x, E.pack_conv(blk.conv1.weight, None, blk.bn1, dev), t1, act=ACT_RELU, round_out=True, name=name + ".conv1")
    t2 = plan.new_act(x.N, 1, H2, H2, width)
    pk2 = E.pack_conv(blk.conv2.weight, None, blk.bn2, dev)
    if flat:
        plan.conv_flat(t1, pk2, E.conv_taps(1, 3, 3, 0, 0, 0), t2, act=ACT_RELU, round_out=True, name=name + ".conv2")
    else:
        plan.conv(t1, pk2, E.conv_taps(1, 3, 3, 0, 1, 1), t2, stride=(1, s, s), act=ACT_RELU, round_out=True,
                  name=name + ".conv2")
    if blk.downsample is not None:
        idn = plan.new_act(x.N, 1, H2, H2, cout)
        pk = E.pack_conv(blk.downsample[0].weight, None, blk.downsample[1], dev)
        # the identity branch is stored TF32-rounded so that conv3 can add it on the tensor cores exactly
        if s == 1:
            plan.linear(x, pk, idn, round_out=True, name=name + ".downsample")
        else:
            plan.conv(x, pk, [(0, 0, 0)], idn, stride=(1, s, s), round_out=True, name=name + ".downsample")
    else:
        idn = x
    out = plan.new_act(x.N, 1, H2, H2, cout)
    pk3 = E.pack_conv(blk.conv3.weight, None, blk.bn3, dev)
    plan.linear(t2, pk3, out, act=ACT_RELU, residual=idn, res_after_act=False, round_out=True,
                res_via_mma=(cout % 256 == 0), name=name + ".conv3")
    # every reader of the block's temporaries and of its input has been recorded (all in this lane): their memory backs
    # the next blocks (the whole trunk lives in ~4 buffers per resolution instead of ~50)
    for a in (t1, t2, idn, x):
        plan.release(a)
    return out


# --------------------------------------------------------------------------------------------------
# Swin-T (timm swin_tiny_patch4_window7_224, features_only) + the wrapper's per-stage LayerNorm([C,H,W])
# (models/swin_transformer.py:71-94; timm semantics restated in SURVEY 8c)
# --------------------------------------------------------------------------------------------------
def relative_position_bias(table, heads):
    """[169, heads] table -> expanded [heads, 49, 49] (index (dh+6)*13 + (dw+6))"""
    ws = WINDOW
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    idx = ((rel[0] + ws - 1) * (2 * ws - 1) + (rel[1] + ws - 1)).to(table.device)
    return table.detach().float()[idx.view(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1).contiguous()


def lower_swin(plan, swin, img, N, stage_tail=None):
    """swin: the SwinTransformer wrapper (has .model, .layer_norm, .cfg).  Returns the list of per-stage
    Acts [N,H,W,C] AFTER the wrapper LayerNorm, in SWIN_T_STAGES order, stored TF32-rounded NHWC.
    stage_tail(i, f): optional callback recorded right after output i is available (in a side lane chosen by the
    caller), so per-stage consumers overlap the remaining Swin stages."""
    dev = _dev(plan)
    model = swin.model
    stages = [i % 4 for i in swin.cfg.NETWORK.SWIN_T_STAGES]
    pe = model.patch_embed
    x4 = stage_image(plan, img, N)
    emb = plan.new_act(N, 1, 56, 56, 96)
    pitch, x0 = img_layout(plan)
    if plan.dtype == torch.bfloat16:
        # patch embedding over pixel pairs (16 bytes): image columns 4j .. 4j+3 are staged columns 4j+2 .. 4j+5 = pairs
        # 2j+1, 2j+2 -> 4 kh x 2 pair taps of 8 channels, K = 64, stride 2 in pair units
        w = pe.proj.weight.detach().float()                                  # [96, Cin <= 4, 4, 4]
        Wp = torch.zeros(96, 4, 2, 2, 4, dtype=torch.float32, device=w.device)   # [co, kh, pair, pixel, channel]
        for kw in range(4):
            Wp[:, :, kw // 2, kw % 2, :w.shape[1]] = w[:, :, :, kw].permute(0, 2, 1)
        pairs = Act(x4.buf.view(-1, 8), N, 1, 224, pitch // 2, 8)
        plan.conv(pairs, E.pack_matrix(Wp.reshape(96, 64), pe.proj.bias, dev),
                  [(0, kh, x0 // 2 + pr) for kh in range(4) for pr in range(2)], emb, stride=(1, 4, 2), rows_dhw=(1, 56, 56),
                  name="swin.patch_embed.proj")
    else:
        # patch embedding: 4x4 s4 conv = 16 one-pixel taps, K = 64 (image column x sits at staged column x + 1)
        plan.conv(x4, E.pack_conv(pe.proj.weight, pe.proj.bias, None, dev, cin_pad=4),
                  [(0, kh, kw + x0) for kh in range(4) for kw in range(4)], emb, stride=(1, 4, 4), rows_dhw=(1, 56, 56),
                  name="swin.patch_embed.proj")
    x = plan.new_act(N, 1, 56, 56, 96)
    plan.layernorm_rows(emb, pe.norm.weight.detach().float().to(dev), pe.norm.bias.detach().float().to(dev), x,
                        eps=pe.norm.eps, round_out=False, name="swin.patch_embed.norm")
    plan.release(emb)
    x_private = True   # `x` is read by this lane only (a stage output is also read by its wrapper LayerNorm in a side lane)
    feats = {}
    outs = [None] * len(stages)
    for s in range(max(stages) + 1):
        layer = getattr(model, f"layers_{s}")
        Cc, H = 96 * 2 ** s, 56 // 2 ** s
        heads = Cc // 32
        if s > 0:
            ds = layer.downsample
            merged = plan.new_act(N, 1, H, H, 2 * Cc)  # 4 * (C/2) concatenated channels
            plan.layernorm_rows(x, ds.norm.weight.detach().float().to(dev), ds.norm.bias.detach().float().to(dev),
                                merged, merge_hw=(2 * H, 2 * H), eps=ds.norm.eps, name=f"swin.{s}.merge.norm")
            x = plan.new_act(N, 1, H, H, Cc)
            plan.linear(merged, E.pack_matrix(ds.reduction.weight, None, dev), x, name=f"swin.{s}.merge.reduction")
            plan.release(merged)
            x_private = True
        for j, blk in enumerate(layer.blocks):
            nm = f"swin.{s}.{j}"
            shift = WINDOW // 2 if (j % 2 == 1 and H > WINDOW) else 0
            y = plan.new_act(N, 1, H, H, Cc)
            plan.layernorm_rows(x, blk.norm1.weight.detach().float().to(dev), blk.norm1.bias.detach().float().to(dev), y,
                                eps=blk.norm1.eps, name=nm + ".norm1")
            qkv = plan.new_act(N, 1, H, H, 3 * Cc)
            plan.linear(y, E.pack_matrix(blk.attn.qkv.weight, blk.attn.qkv.bias, dev), qkv, round_out=True, name=nm + ".qkv")
            plan.release(y)
            att = plan.new_act(N, 1, H, H, Cc)
            bias = relative_position_bias(blk.attn.relative_position_bias_table, heads).to(dev)
            if "attn_range_flag" not in plan.taps:
                plan.taps["attn_range_flag"] = plan.zeros(1, dtype=torch.int32)
            plan.window_attention(qkv, att, bias, H, H, heads, shift, 32 ** -0.5, name=nm + ".attn",
                                  range_flag=plan.taps["attn_range_flag"])
            plan.release(qkv)
            x1 = plan.new_act(N, 1, H, H, Cc)
            plan.linear(att, E.pack_matrix(blk.attn.proj.weight, blk.attn.proj.bias, dev), x1, residual=x,
                        name=nm + ".proj")
            plan.release(att)
            if x_private:
                plan.release(x)
            x_private = True   # the block output below has no reader outside this lane unless it ends the stage
            g2, b2 = blk.norm2.weight.detach().float().to(dev), blk.norm2.bias.detach().float().to(dev)
            # (the fused MLP kernel exists for fp32 / TF32 storage; bf16 plans run fc1 / fc2 as two contractions)
            fused = E.mlp_fusable(Cc, blk.mlp.fc1.out_features) and plan.dtype == torch.float32
            x = plan.new_act(N, 1, H, H, Cc)
            if fused and E.mlp_ln_fusable(Cc):
                # stage 0: norm2 -> fc1 -> GELU -> fc2 -> + x1 in one kernel (x1 is read once, normalised in shared memory)
                plan.mlp(x1, E.pack_matrix(blk.mlp.fc1.weight, blk.mlp.fc1.bias, dev),
                         E.pack_matrix(blk.mlp.fc2.weight, blk.mlp.fc2.bias, dev), x, residual=x1, name=nm + ".mlp",
                         ln=(g2, b2, blk.norm2.eps))
                plan.release(x1)
                continue
            y2 = plan.new_act(N, 1, H, H, Cc)
            plan.layernorm_rows(x1, g2, b2, y2, eps=blk.norm2.eps, name=nm + ".norm2")
            if fused:
                # stage 1: fc1 -> GELU -> fc2 -> + x1 in one kernel, the 4C-wide hidden activation stays on the SM
                plan.mlp(y2, E.pack_matrix(blk.mlp.fc1.weight, blk.mlp.fc1.bias, dev),
                         E.pack_matrix(blk.mlp.fc2.weight, blk.mlp.fc2.bias, dev), x, residual=x1, name=nm + ".mlp")
                plan.release(y2)
                plan.release(x1)
                continue
            hid = plan.new_act(N, 1, H, H, 4 * Cc)
            plan.linear(y2, E.pack_matrix(blk.mlp.fc1.weight, blk.mlp.fc1.bias, dev), hid, act=ACT_GELU, round_out=True,
                        name=nm + ".fc1")
            plan.release(y2)
            plan.linear(hid, E.pack_matrix(blk.mlp.fc2.weight, blk.mlp.fc2.bias, dev), x, residual=x1, name=nm + ".fc2")
            plan.release(hid)
            plan.release(x1)
        feats[s] = x
        x_private = False   # read by the wrapper LayerNorm (a side lane when stage tails overlap) and by the next stage
        for i, si in enumerate(stages):   # wrapper LayerNorm (+ the caller's per-stage tail) as soon as the stage is done
            if si != s:
                continue
            if stage_tail is not None:
                plan.lane(2 + i % 6)
            ln = swin.layer_norm[i]
            f = plan.new_act(N, 1, H, H, Cc)
            # affine is stored [C,H,W] (it normalises an NCHW tensor); our data is [H,W,C]
            g = ln.weight.detach().float().permute(1, 2, 0).contiguous().to(dev)
            b = ln.bias.detach().float().permute(1, 2, 0).contiguous().to(dev)
            plan.layernorm_sample(x, g, b, f, eps=ln.eps, name=f"swin.layer_norm.{i}")
            outs[i] = f
            if stage_tail is not None:
                stage_tail(i, f)
                plan.lane(0)
    return outs


# --------------------------------------------------------------------------------------------------
# Cross-view attention (models/cross_view_attention.py:59-134)
# --------------------------------------------------------------------------------------------------
def lower_cva(plan, cva, x, B, V, out_pad=(0, 0, 0)):
    """x: Act [B*V,7,7,C] (C=512).  Returns Act of the same shape (TF32-rounded: it feeds fusion_layer),
    optionally inside a zero border so the next 3x3 convolution can stream it by TMA."""
    dev = _dev(plan)
    N, Cc, R, heads = x.N, x.C, cva.reduced_channels, cva.num_heads
    ratio = cva.attention_spatial_downsample_ratio
    if ratio > 1:
        h = (x.H - ratio) // ratio + 1
        small = plan.new_act(N, 1, h, h, Cc)
        dw = cva.downsample_qkv
        plan.dwconv(x, dw.weight.detach().float().reshape(Cc, ratio * ratio).t().contiguous().to(dev),
                    dw.bias.detach().float().to(dev) if dw.bias is not None else None, small, ratio, name="cva.downsample_qkv")
    else:
        h, small = x.H, x
    qkv = plan.new_act(N, 1, h, h, 3 * R)
    plan.linear(small, E.pack_conv(cva.qkv_conv.weight, cva.qkv_conv.bias, None, dev), qkv, name="cva.qkv_conv")
    att = plan.new_act(N, 1, h, h, R)
    plan.view_attention(qkv, att, B, V, heads, 1.0 / float(cva.head_dim * V) ** 0.5, name="cva.attention")
    y = plan.new_act(N, 1, x.H, x.W, Cc)
    pk = E.pack_conv(cva.proj_conv.weight, cva.proj_conv.bias, None, dev)
    if ratio > 1:
        proj = plan.new_act(N, 1, h, h, Cc)
        plan.linear(att, pk, proj, name="cva.proj_conv")
        plan.bilinear_add(proj, x, y, name="cva.upsample_residual")
    else:
        plan.linear(att, pk, y, residual=x, round_out=True, name="cva.proj_conv")
    hid = plan.new_act(N, 1, x.H, x.W, Cc)
    plan.linear(y, E.pack_conv(cva.ffn[0].weight, cva.ffn[0].bias, None, dev), hid, act=ACT_GELU, round_out=True,
                name="cva.ffn.0")
    out = plan.new_act(N, 1, x.H, x.W, Cc, pad=out_pad)
    plan.linear(hid, E.pack_conv(cva.ffn[2].weight, cva.ffn[2].bias, cva.batch_norm, dev), out, round_out=True,
                name="cva.ffn.2+bn")
    return out


# --------------------------------------------------------------------------------------------------
# Encoder (models/encoder.py:113-164)
# --------------------------------------------------------------------------------------------------
def lower_encoder(plan, enc, img, B, V):
    """img: [B*V,3,224,224] NCHW staging tensor.  Returns Act [B*V,7,7,256] (full fp32)."""
    dev = _dev(plan)
    N = B * V
    net = enc.cfg.NETWORK
    cat = plan.new_act(N, 1, 7, 7, 512)
    img = stage_image(plan, img, N)   # shared by both branches: staged before they fork
    # The ResNet and the Swin branch are independent until the concat (encoder.py:119-143): they are recorded as two
    # concurrent graph branches (lane 1 / lane 0), so the tail of one branch's persistent kernels overlaps the other's.
    two_lanes = True
    if two_lanes:
        plan.lane(1)
    # ResNet branch.  avg_pool2d(conv1x1(x)) == conv1x1(avg_pool2d(x)): pool first, 4x fewer MACs.
    r = lower_resnet_trunk(plan, enc.resnet, img, N)
    rp = plan.new_act(N, 1, 7, 7, 1024)
    plan.pool(r, rp, (1, 2, 2), (1, 2, 2), (0, 0, 0), POOL_AVG, round_out=True, name="encoder.avg_pool")
    plan.release(r)
    plan.linear(rp, E.pack_conv(enc.resnet_reduce.weight, enc.resnet_reduce.bias, None, dev), cat.channels(0, 256),
                round_out=True, name="encoder.resnet_reduce")
    plan.release(rp)
    if two_lanes:
        plan.lane(0)
    # Swin branch.  Per-stage tails (wrapper LayerNorm, 1x1 reduce, all but the last strided convolution of the
    # downsample chain; encoder.py:134-138) are recorded in side lanes right after their stage, so they overlap the later
    # Swin stages; the last op of every chain accumulates into the running sum `sw` and therefore runs after the join,
    # in stage order (the reference's summation order).
    sw = cat.channels(256, 256)
    finals = []
    multi = net.USE_SWIN_T_MULTI_STAGE
    n_st = len(enc.cfg.NETWORK.SWIN_T_STAGES) if multi else 1

    def stage_tail(i, f):
        if not multi:
            return
        chain = enc.swin_downsamples[i]
        convs = [] if isinstance(chain, torch.nn.Identity) else [(chain[k], chain[k + 1]) for k in range(0, len(chain), 3)]
        first, last_stage = i == 0, i == n_st - 1
        red = enc.swin_stage_reduces[i]
        pk = E.pack_conv(red.weight, red.bias, None, dev)
        if not convs:   # the reduce itself accumulates into the running sum
            finals.append((i, lambda f=f, pk=pk, first=first, last_stage=last_stage, i=i: plan.linear(
                f, pk, sw, residual=None if first else sw, round_out=last_stage, name=f"encoder.swin_reduce.{i}")))
            return
        t = plan.new_act(N, 1, f.H, f.W, 256)
        plan.linear(f, pk, t, round_out=True, name=f"encoder.swin_reduce.{i}")
        for ci, (conv, bn) in enumerate(convs):
            Ho = (t.H + 2 - 3) // 2 + 1
            pkc = E.pack_conv(conv.weight, conv.bias, bn, dev)
            if ci == len(convs) - 1:
                finals.append((i, lambda t=t, pkc=pkc, first=first, last_stage=last_stage, i=i, ci=ci: plan.conv(
                    t, pkc, E.conv_taps(1, 3, 3, 0, 1, 1), sw, stride=(1, 2, 2), act=ACT_RELU,
                    residual=None if first else sw, res_after_act=True, round_out=last_stage,
                    name=f"encoder.swin_down.{i}.{ci}")))
            else:
                o = plan.new_act(N, 1, Ho, Ho, 256)
                plan.conv(t, pkc, E.conv_taps(1, 3, 3, 0, 1, 1), o, stride=(1, 2, 2), act=ACT_RELU, round_out=True,
                          name=f"encoder.swin_down.{i}.{ci}")
                t = o

    overlap = two_lanes and multi   # single-stage mode: the reduce below reads the LayerNorm output on lane 0
    feats = lower_swin(plan, enc.swin_transformer, img, N, stage_tail=stage_tail if overlap else None)
    if not overlap:
        for i, f in enumerate(feats):
            stage_tail(i, f)
    if not multi:
        plan.linear(feats[-1], E.pack_conv(enc.swin_reduce.weight, enc.swin_reduce.bias, None, dev), sw, round_out=True,
                    name="encoder.swin_reduce")
    if two_lanes:
        plan.join()
    for _, emit in sorted(finals, key=lambda t: t[0]):   # the accumulating ops, in stage order
        emit()
    x = cat
    plan.taps.update(resnet=cat.channels(0, 256), swin_sum=sw, swin=feats, pre_cva=cat)
    if net.USE_CROSS_VIEW_ATTENTION:
        x = lower_cva(plan, enc.cross_view_attention, cat, B, V, out_pad=(0, 1, 1))
    plan.taps["post_cva"] = x
    seq = [("fusion_layer", enc.fusion_layer), ("layer1", enc.layer1), ("layer2", enc.layer2), ("layer3", enc.layer3)]
    for li, (nm, layer) in enumerate(seq):
        last = li == len(seq) - 1
        # the module output is fp32 whatever the plan's storage type (the decoder reads it)
        o = plan.new_act(N, 1, 7, 7, 256, pad=(0, 0, 0) if last else (0, 1, 1), dtype=torch.float32 if last else None)
        pk = E.pack_conv(layer[0].weight, layer[0].bias, layer[1], dev)
        if any(x.pad):   # zero-bordered input: one TMA box per filter tap
            plan.conv_flat(x, pk, E.conv_taps(1, 3, 3, 0, 0, 0), o, act=ACT_RELU, round_out=not last, name="encoder." + nm)
        else:
            plan.conv(x, pk, E.conv_taps(1, 3, 3, 0, 1, 1), o, act=ACT_RELU, round_out=not last, name="encoder." + nm)
        x = o
    return x


# --------------------------------------------------------------------------------------------------
# Decoder (models/decoder.py:48-99)
# --------------------------------------------------------------------------------------------------
def _convT_layer(plan, x, conv, bn, pads, out, name, act=ACT_RELU, residual=None, round_out=True, n_logical=None,
                 block_n=None, epi_tail=None, out_elem_map=None, out_scale=1.0):
    """stride-2 ConvTranspose3d as 8 parity-class implicit GEMMs; out dims are 2x the input dims."""
    dev = _dev(plan)
    od, oh, ow = 2 * x.D, 2 * x.H, 2 * x.W
    assert tuple(out.inner) == (od, oh, ow)
    ob, osn, osd, osh, osw = out.interior_map()
    for pd in (0, 1):
        for ph in (0, 1):
            for pw in (0, 1):
                plan.lane(1 + pd * 4 + ph * 2 + pw)   # the eight classes write disjoint voxels: run them concurrently
                pk, taps = E.pack_convT_class(conv.weight, bn, dev, pads, (pd, ph, pw), n_logical=n_logical,
                                              block_n=block_n, bias=conv.bias)
                vox = (pd * oh + ph) * ow + pw
                omap = (ob + pd * osd + ph * osh + pw * osw, osn, 2 * osd, 2 * osh, 2 * osw)
                tail = None
                if epi_tail is not None:
                    aux, out2 = epi_tail
                    tail = (aux, out2, (vox, od * oh * ow, 2 * oh * ow, 2 * ow, 2))
                plan.conv(x, pk, taps, out, out_map=omap, rows_dhw=(x.D, x.H, x.W), act=act, residual=residual,
                          res_after_act=True, round_out=round_out, out_scale=out_scale, epi_tail=tail,
                          name=f"{name}.p{pd}{ph}{pw}")
    plan.join()
    return out


def lower_decoder(plan, dec, feat, N):
    """feat: Act [N,7,7,256].  Returns (raw Act: 32^3 voxels inside a zero border, RAW_CS-channel rows with 9 live
    channels -- the layout the merger's slab convolution streams by TMA -- and coarse [N, 32768])."""
    dev = _dev(plan)
    g = plan.new_act(N, 2, 2, 2, 256)
    # AdaptiveAvgPool2d 7->2 = windows [0,4) and [3,7); the new depth axis replicates (stride_d = 0)
    src = Act(feat.buf, N, 1, 7, 7, 256, feat.c0)
    plan.pool(src, g, (1, 4, 4), (0, 3, 3), (0, 0, 0), POOL_AVG, round_out=True, name="decoder.spatial_reduce")
    x = g
    for li, (layer, pads, cout) in enumerate(((dec.layer1, (2, 1, 1), 128), (dec.layer2, (1, 1, 1), 64),
                                              (dec.layer3, (1, 1, 1), 32))):
        o = plan.new_act(N, 2 * x.D, 2 * x.H, 2 * x.W, cout)
        if tuple(layer[0].kernel_size) == (4, 4, 4):
            # k4 s2 p1: all eight output-parity classes in the N dimension of ONE contraction over the 3x3x3 input
            # neighbourhood (one launch instead of eight; the structural zeros cost 3.4x MACs on a layer that is launch-
            # and latency-bound: 24 micro-launches took 0.63 ms against a 0.08 ms floor)
            plan.convT_fused(x, E.pack_convT_fused(layer[0].weight, layer[1], dev, bias=layer[0].bias), o, act=ACT_RELU,
                             round_out=True, name=f"decoder.layer{li + 1}")
        else:   # layer1: kernel (6, 4, 4) -- three depth taps per class
            _convT_layer(plan, x, layer[0], layer[1], pads, o, f"decoder.layer{li + 1}")
        x = o
    raw = plan.new_act(N, 32, 32, 32, 16, Cs=RAW_CS, pad=(1, 1, 1))
    coarse = plan.empty(N, 32768)
    l5 = dec.layer5[0]
    w5 = torch.zeros(9)
    w5[:8] = l5.weight.detach().float().reshape(8).cpu()
    if l5.bias is not None:
        w5[8] = l5.bias.detach().float().cpu()[0]
    # layer4 (32 -> 8 channels) + layer5 + cat: all eight parity classes in one GEMM (N = 64), tail in the epilogue
    plan.convT_fused(x, E.pack_convT_fused(dec.layer4[0].weight, dec.layer4[1], dev, bias=dec.layer4[0].bias, block_n=64),
                     raw, act=ACT_RELU, round_out=True, tail=(w5.to(dev), coarse), name="decoder.layer4+5")
    return raw, coarse


# --------------------------------------------------------------------------------------------------
# Merger (models/merger.py:56-107)
# --------------------------------------------------------------------------------------------------
def lower_merger(plan, mer, raw, coarse, B, V, operands="fp16", range_flag=None):
    """raw: Act of 32^3 voxels inside a (1,1,1) zero border, RAW_CS-channel rows (9 live, TF32-rounded, rest zero);
    coarse: [N,32768] tensor.  Returns (merged [B, 32768], pre-softmax scores [N, 32768]).
    All six Conv3d(k3,p1) layers run on the depth-marching TMA slab kernel over zero-bordered 34^3 volumes.
    operands: MMA operand type of the slab kernel ("fp16": exact for this path's TF32-rounded activations up to 65504,
    beyond that they saturate and `range_flag` (device int32[1]) is set; "tf32": fp32's range, twice the MMA instructions)."""
    dev = _dev(plan)
    N = B * V
    slope = float(mer.cfg.NETWORK.LEAKY_VALUE)
    cat = plan.new_act(N, 32, 32, 32, 64, pad=(1, 1, 1))   # four 16-channel groups: w1 | w2 | w3 | w4 (9 live each)

    def box(a, c0):   # the 32-channel TMA box starting at channel c0 of a zero-bordered buffer
        return Act(a.buf, N, 34, 34, 34, 32, c0, (1, 1, 1))

    x = box(raw, raw.c0)
    for i, layer in enumerate((mer.layer1, mer.layer2, mer.layer3, mer.layer4)):
        o = cat.channels(16 * i, 16)
        plan.conv3_slab(x, E.pack_conv3_slab(layer[0].weight, layer[0].bias, layer[1], dev, n_logical=16), o, 9,
                        act=ACT_LEAKY, act_param=slope, round_out=True, name=f"merger.layer{i + 1}", operands=operands,
                        range_flag=range_flag)
        x = box(cat, 16 * i)
    # layer5 sees cat(w1..w4): reference channel 9*g + c lives at 16*g + c here.  Two slab passes over the two
    # 32-channel halves; the second adds the first's partial sums before bias / LeakyReLU.
    w5, b5 = E.fold_bn(mer.layer5[0].weight, mer.layer5[0].bias, mer.layer5[1])
    # (64-byte rows: layer6 streams this buffer on the CUDA cores and a 128-byte row would double its DRAM traffic --
    # ncu: 989 MB per launch for 483 MB of 16-channel rows, profiles/r2_launches_v40_summary.txt)
    t = plan.new_act(N, 32, 32, 32, 16, pad=(1, 1, 1))
    for half in (0, 1):
        wh = torch.zeros(9, 32, 3, 3, 3, device=w5.device)
        for gi in (0, 1):
            wh[:, 16 * gi:16 * gi + 9] = w5[:, 9 * (2 * half + gi):9 * (2 * half + gi) + 9]
        pk = E.pack_conv3_slab(wh, b5 if half else None, None, dev, n_logical=16)
        plan.conv3_slab(box(cat, 32 * half), pk, t, 25, act=ACT_LEAKY if half else ACT_NONE, act_param=slope,
                        residual=t if half else None, res_after_act=False, round_out=False,   # layer6 reads it in fp32
                        name=f"merger.layer5.{'ab'[half]}", operands=operands, range_flag=range_flag)
    # layer6 (9 -> 1 channel): 243 MACs per voxel, fp32 on the CUDA cores (a tensor-core tile is issue-bound here)
    wts = plan.empty(N, 32768)
    plan.conv3_to1(t, mer.layer6[0].weight, mer.layer6[0].bias, mer.layer6[1], wts, slope, name="merger.layer6")
    merged = plan.empty(B, 32768)
    plan.merger_fuse(wts, coarse, merged, B, V, 32768, name="merger.softmax_fuse")
    return merged, wts


# --------------------------------------------------------------------------------------------------
# Refiner (models/refiner.py:72-106)
# --------------------------------------------------------------------------------------------------
def lower_refiner(plan, ref, vol, B):
    """vol: [B, 32768] fp32 tensor (planar 32^3).  Returns refined [B, 32768]."""
    dev = _dev(plan)
    slope = float(ref.cfg.NETWORK.LEAKY_VALUE)
    # layer1: Conv3d(1,32,k4,p2) -> 33^3, BN, LeakyReLU, MaxPool3d(2) (floors to 16^3: conv outputs 0..31 are used).
    # One GEMM row per POOLED voxel: its 2x2x2 conv positions share one 5^3 input window (im2col, K = 125), the
    # weights are the 4^3 kernel placed at the eight offsets inside that window (N = 8 x 32), and the epilogue takes
    # the max over the eight column groups before bias / LeakyReLU (both commute with max).
    cols = plan.im2col(vol, (32768, 0, 1024, 32, 1), B, 1, (32, 32, 32), (5, 5, 5), 2, (2, 2, 2), (16, 16, 16), 128,
                       name="refiner.layer1.im2col")
    c1 = ref.layer1[0]
    w, b = E.fold_bn(c1.weight, c1.bias, ref.layer1[1])
    w8 = torch.zeros(8, 32, 5, 5, 5, device=w.device)
    for a_ in (0, 1):
        for b_ in (0, 1):
            for c_ in (0, 1):
                w8[a_ * 4 + b_ * 2 + c_, :, a_:a_ + 4, b_:b_ + 4, c_:c_ + 4] = w[:, 0]
    l16 = plan.new_act(B, 16, 16, 16, 32)
    w8p = torch.nn.functional.pad(w8.reshape(256, 125), (0, 3))   # K = 125 padded to the im2col row length
    plan.linear(cols, E.pack_matrix(w8p, b, dev, block_n=256), l16, act=ACT_LEAKY, act_param=slope,
                round_out=True, pool8=True, name="refiner.layer1+pool")
    x, skips = l16, [l16]
    for li, (layer, cout) in enumerate(((ref.layer2, 64), (ref.layer3, 128))):
        full = plan.new_act(B, x.D, x.H, x.W, cout)   # conv outputs 0..D-1 of the (D+1)^3 the reference computes
        plan.conv(x, E.pack_conv(layer[0].weight, layer[0].bias, layer[1], dev), E.conv_taps(4, 4, 4, 2, 2, 2), full,
                  act=ACT_LEAKY, act_param=slope, name=f"refiner.layer{li + 2}")
        p = plan.new_act(B, x.D // 2, x.H // 2, x.W // 2, cout)
        plan.pool(full, p, (2, 2, 2), (2, 2, 2), (0, 0, 0), POOL_MAX, round_out=True, name=f"refiner.layer{li + 2}.pool")
        skips.append(p)
        x = p
    l4 = x  # [B,4,4,4,128] == [B, 8192] in (d,h,w,c) order; the reference flattens (c,d,h,w)
    fc4, fc5 = ref.layer4[0], ref.layer5[0]
    w4 = fc4.weight.detach().float().view(2048, 128, 64).permute(0, 2, 1).reshape(2048, 8192)
    h = plan.new_act(B, 1, 1, 1, 2048)
    flat = Act(l4.buf.view(B, 8192), B, 1, 1, 1, 8192, 0)
    plan.linear(flat, E.pack_matrix(w4, fc4.bias, dev, block_n=32), h, act=ACT_RELU, round_out=True, name="refiner.layer4")
    w5 = fc5.weight.detach().float().view(128, 64, 2048).permute(1, 0, 2).reshape(8192, 2048)
    b5 = fc5.bias.detach().float().view(128, 64).t().reshape(8192)
    r4buf = plan.empty(B * 64, 128)
    r4 = Act(r4buf, B, 4, 4, 4, 128, 0)
    plan.linear(h, E.pack_matrix(w5, b5, dev, block_n=32), Act(r4buf.view(B, 8192), B, 1, 1, 1, 8192, 0), act=ACT_RELU,
                residual=flat, res_after_act=True, round_out=True, name="refiner.layer5+skip")
    r8 = plan.new_act(B, 8, 8, 8, 64)
    _convT_layer(plan, r4, ref.layer6[0], ref.layer6[1], (1, 1, 1), r8, "refiner.layer6", residual=skips[1])
    r16 = plan.new_act(B, 16, 16, 16, 32)
    _convT_layer(plan, r8, ref.layer7[0], ref.layer7[1], (1, 1, 1), r16, "refiner.layer7", residual=skips[0])
    out = plan.empty(B, 32768)
    oact = Act(out.view(-1, 1), B, 32, 32, 32, 1, 0)
    vact = Act(vol.view(-1, 1), B, 32, 32, 32, 1, 0)
    # layer8 (32 -> 1 channel): the eight parity classes are the N = 8 columns of one GEMM
    plan.convT_fused(r16, E.pack_convT_fused(ref.layer8[0].weight, None, dev, bias=ref.layer8[0].bias), oact, act=ACT_NONE,
                     residual=vact, res_after_act=True, out_scale=0.5, name="refiner.layer8")
    return out

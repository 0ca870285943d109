"""Host-side plan builder for libswinvox_b200: activation bookkeeping, weight re-layout, BatchNorm
folding and tap tables.  Everything here runs once per (module, input shape); the per-forward work
is a single ``svx_plan_run`` call.

Layout conventions: activations are fp32 channels-last buffers ``[N, D, H, W, Cs]`` (2-D maps use
D=1); contraction weights are ``[Npad, Kpad]`` with k = tap*Cin + c, pre-rounded to TF32.
"""
import ctypes as C
import math
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, A_FLAT, A_GATHER, A_IM2COL, A_PLAIN, A_SLAB3, EPI_CONVT8, EPI_DEC_TAIL, EPI_POOL8, EPI_STD,
                   POOL_AVG, POOL_MAX)

BLOCK_NS = (16, 32, 64, 96, 128, 192, 256)


def tf32_round(t):
    """round-to-nearest (ties away) to TF32, the host twin of cvt.rna.tf32.f32"""
    t = t.contiguous().float()
    bits = t.view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def choose_block_n(n, rows=None, gather=False, heavy_epilogue=False, k=None):
    """N tile of one contraction.  Measured on B200 (profiles/r1_gemm_bench_v6.log): the widest tile that divides N
    wins as long as the grid still fills the 148 SMs for a few waves; gather-mode convolutions re-fetch their A tile
    once per N tile, so they take the widest tile regardless."""
    if n <= 16:
        return 16
    if n <= 32:
        return 32
    if n <= 64:
        return 64
    cands = [bn for bn in (256, 192, 128, 96) if n % bn == 0] or [128 if n > 96 else 96]
    if heavy_epilogue and n % 96 == 0 and (k is None or k <= 192):
        # erf-GELU epilogues are bound by instruction issue in the epilogue warps (profiles/r1_ncu_swin_block_v17.txt):
        # the 96-wide tile runs two CTAs per SM, i.e. twice the epilogue warps per SM.  Only while the main loop is
        # short: a tcgen05.mma costs max(~89, N/2) cycles (profiles/r1_umma_tf32_issue_cost.txt), so N = 96 caps the
        # tensor pipe at 53 % and loses on the deeper-K stages.
        return 96
    if rows is None:
        return cands[0] if gather else ([bn for bn in cands if bn <= 192] or cands)[0]
    tiles_m = (rows + 127) // 128
    overhead = 256 if gather else 48   # per-tile cost that does not shrink with the tile (epilogue latency / A gather)
    best, best_cost = None, None
    for bn in cands:
        waves = -(-tiles_m * (-(-n // bn)) // 148)
        cost = waves * (bn + overhead)
        if best_cost is None or cost < best_cost:
            best, best_cost = bn, cost
    return best


def esize_of(dtype):
    return 2 if dtype == torch.bfloat16 else 4


def to_operand(t, dtype):
    """values as the tensor cores will read them: TF32-rounded fp32, or bf16 (round to nearest even)"""
    return t.to(torch.bfloat16).contiguous() if dtype == torch.bfloat16 else tf32_round(t)


def mlp_fusable(c, hidden):
    """shapes the fused MLP kernel (svx_mlp.cu) is instantiated for: Swin stages 0 and 1"""
    return c in (96, 192) and hidden == 4 * c


def mlp_ln_fusable(c):
    """widths for which the fused MLP kernel can also apply the preceding LayerNorm (two X buffers: stage 0)"""
    return c == 96


def round_up(x, m):
    return (x + m - 1) // m * m


@dataclass
class Act:
    """channels [c0, c0+C) of a channels-last buffer [N, D, H, W, Cs]"""
    buf: torch.Tensor
    N: int
    D: int
    H: int
    W: int
    C: int
    c0: int = 0
    pad: tuple = (0, 0, 0)   # zero border (d, h, w) on each side; D/H/W above INCLUDE it

    @property
    def inner(self):
        """extents of the real data inside the zero border"""
        return (self.D - 2 * self.pad[0], self.H - 2 * self.pad[1], self.W - 2 * self.pad[2])

    def interior_map(self):
        """(base, s_n, s_d, s_h, s_w) element offsets of data voxel (n, d, h, w) inside the padded buffer"""
        Cs = self.Cs
        base = self.c0 + ((self.pad[0] * self.H + self.pad[1]) * self.W + self.pad[2]) * Cs
        return (base, self.D * self.H * self.W * Cs, self.H * self.W * Cs, self.W * Cs, Cs)

    @property
    def Cs(self):
        return self.buf.shape[-1]

    @property
    def dtype(self):
        return self.buf.dtype

    @property
    def esize(self):
        return self.buf.element_size()

    @property
    def bf16(self):
        return self.buf.dtype == torch.bfloat16

    @property
    def pixels(self):
        return self.N * self.D * self.H * self.W

    def view(self):
        """logical [N, D, H, W, C] torch view of the real data (for tests / module boundaries)"""
        v = self.buf.view(self.N, self.D, self.H, self.W, self.Cs)[..., self.c0:self.c0 + self.C]
        pd, ph, pw = self.pad
        return v[:, pd:self.D - pd, ph:self.H - ph, pw:self.W - pw]

    def channels(self, c0, c):
        return Act(self.buf, self.N, self.D, self.H, self.W, c, self.c0 + c0, self.pad)


class WeightPack:
    """Prepared weights of one contraction.  The N tile (and with it the zero padding of W to [Npad, Kpad]) is fixed
    lazily by `finalize`, when the op that uses the pack knows its row count and operand mode."""

    def __init__(self, W, bias, N, K, block_n=None, prepared=False):
        self.raw_W, self.raw_bias = W, bias     # [n_rows, K] fp32 (not yet rounded unless `prepared`) / [n_rows]
        self.N, self.K, self.block_n = N, K, block_n
        self.W = self.bias = None
        self.dtype = torch.float32
        self.prepared = prepared
        if prepared:   # already in the kernel's final layout and rounding (the slab packs)
            self.finalize()

    def finalize(self, rows=None, gather=False, heavy_epilogue=False, dtype=torch.float32):
        """fixes the N tile, pads to [Npad, Kpad] (Kpad: whole 128-byte k-chunks) and rounds to the operand type"""
        if self.W is not None:
            assert self.dtype == dtype, "a weight pack serves one operand type"
            return self
        self.dtype = dtype
        bn = self.block_n or choose_block_n(self.N, rows, gather, heavy_epilogue, self.K)
        n, k = self.raw_W.shape
        npad, kpad = round_up(max(n, self.N), bn), round_up(k, 128 // esize_of(dtype))
        raw = self.raw_W if self.prepared else to_operand(self.raw_W, dtype)
        if (npad, kpad) == (n, k):
            self.W = raw.contiguous()
        else:
            self.W = torch.zeros(npad, kpad, dtype=dtype, device=self.raw_W.device)
            self.W[:n, :k] = raw
        self.bias = torch.zeros(npad, dtype=torch.float32, device=self.raw_W.device)
        if self.raw_bias is not None:
            self.bias[:self.raw_bias.numel()] = self.raw_bias
        self.block_n = bn
        self.raw_W = self.raw_bias = None
        return self

    @property
    def Kpad(self):
        return self.W.shape[1] if self.W is not None else round_up(self.K, 128 // esize_of(self.dtype))

    @property
    def Npad(self):
        return self.W.shape[0]


def pack_matrix(w2d, bias, device, block_n=None, n_logical=None):
    """w2d: [N, K] fp32 (already folded / permuted)."""
    n, k = w2d.shape
    W = w2d.detach().to(device=device, dtype=torch.float32).contiguous()   # rounded to the operand type by finalize()
    b = bias.detach().to(device=device, dtype=torch.float32) if bias is not None else None
    return WeightPack(W, b, n_logical or n, k, block_n)


def fold_bn(weight, bias, bn):
    """Fold eval-mode BatchNorm (running stats) into the preceding conv: returns (w, b).
    `weight` has output channels on dim 0."""
    w = weight.detach().float()
    cout = w.shape[0]
    b = bias.detach().float() if bias is not None else torch.zeros(cout, device=w.device)
    if bn is None:
        return w, b
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = w * s.view(-1, *([1] * (w.dim() - 1)))
    b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w, b


def conv_taps(kd, kh, kw, pd, ph, pw):
    return [(a - pd, b - ph, c - pw) for a in range(kd) for b in range(kh) for c in range(kw)]


def pack_conv(weight, bias, bn, device, cin_pad=None, block_n=None, n_logical=None):
    """nn.Conv2d / nn.Conv3d weight [Cout, Cin, (KD,) KH, KW] -> WeightPack with k = tap*Cin_pad + c."""
    w, b = fold_bn(weight, bias, bn)
    if w.dim() == 4:
        w = w.unsqueeze(2)
    cout, cin = w.shape[:2]
    cp = cin_pad or round_up(cin, 4)
    w = w.permute(0, 2, 3, 4, 1)  # [Cout, KD, KH, KW, Cin]
    if cp != cin:
        w = torch.nn.functional.pad(w, (0, cp - cin))
    return pack_matrix(w.reshape(cout, -1), b, device, block_n, n_logical)


def pack_conv3_slab(weight, bias, bn, device, n_logical=None):
    """nn.Conv3d(k=3) weight [Cout<=16, Cin<=32, 3, 3, 3] -> the slab kernel's kw-in-N layout: W[48, 288] with
    row = kw*16 + co and col = (kd*3 + kh)*32 + c (BatchNorm folded, TF32-rounded)."""
    w, b = fold_bn(weight, bias, bn)
    cout, cin = w.shape[:2]
    assert cout <= 16 and cin <= 32 and tuple(w.shape[2:]) == (3, 3, 3)
    W = torch.zeros(3, 16, 9, 32, dtype=torch.float32, device=w.device)       # [kw, co, (kd,kh), c]
    W[:, :cout, :, :cin] = w.permute(4, 0, 2, 3, 1).reshape(3, cout, 9, cin)
    # The kernel's MMA operands are fp16: a power-of-two scale brings the (BatchNorm-folded) weights to max|W| in
    # [2^13, 2^14) whatever their magnitude; the kernel multiplies the accumulator by its inverse (exact) before the bias.
    wmax = float(W.abs().max())
    ws = 2.0 ** (13 - math.frexp(wmax)[1] + 1) if wmax > 0 and math.isfinite(wmax) else 1.0
    W = tf32_round((W * ws).reshape(48, 288)).to(device)
    bb = torch.zeros(48, dtype=torch.float32, device=device)
    bb[:cout] = b.to(device)
    pack = WeightPack(W, bb, n_logical or cout, 288, 48, prepared=True)   # [48, 288] is already the padded layout
    pack.acc_scale = 1.0 / ws
    return pack


def convT_class_taps(ks, pad, par):
    """taps (kernel index k, input offset delta) feeding output parity `par` of a stride-2 ConvTranspose."""
    return [(k, (par + pad - k) // 2) for k in range(ks) if (k - par - pad) % 2 == 0]


def pack_convT_class(weight, bn, device, pads, parity, cin_pad=None, block_n=None, n_logical=None, bias=None):
    """One stride-2 parity class of nn.ConvTranspose3d (weight [Cin, Cout, KD, KH, KW]).
    Returns (WeightPack, taps[(dd,dh,dw)])."""
    w = weight.detach().float().transpose(0, 1)  # [Cout, Cin, KD, KH, KW]
    w, b = fold_bn(w, bias, bn)
    cout, cin = w.shape[:2]
    cp = cin_pad or round_up(cin, 4)
    td = convT_class_taps(w.shape[2], pads[0], parity[0])
    th = convT_class_taps(w.shape[3], pads[1], parity[1])
    tw = convT_class_taps(w.shape[4], pads[2], parity[2])
    cols, taps = [], []
    for kd, dd in td:
        for kh, dh in th:
            for kw, dw in tw:
                col = w[:, :, kd, kh, kw]
                if cp != cin:
                    col = torch.nn.functional.pad(col, (0, cp - cin))
                cols.append(col)
                taps.append((dd, dh, dw))
    return pack_matrix(torch.cat(cols, dim=1), b, device, block_n, n_logical), taps


def pack_convT_fused(weight, bn, device, bias=None, block_n=None):
    """nn.ConvTranspose3d(k=4, s=2, p=1) weight [Cin, Cout, 4, 4, 4] with all eight output-parity classes in N:
    W[cls*Cout + co, tap*Cin + c], tap = (td*3 + th)*3 + tw over the 3x3x3 INPUT neighbourhood (offsets t-1).
    Along one axis, output 2i+par reads input i+delta through kernel index k = par + 1 - 2*delta, so parity 0 uses
    delta in {-1, 0} and parity 1 uses {0, +1}; the other taps of a class are structural zeros."""
    w = weight.detach().float().transpose(0, 1)  # [Cout, Cin, 4, 4, 4]
    assert tuple(w.shape[2:]) == (4, 4, 4)
    w, b = fold_bn(w, bias, bn)
    cout, cin = w.shape[:2]
    assert cin % 4 == 0
    W = torch.zeros(8, cout, 27, cin, dtype=torch.float32, device=w.device)
    for cls in range(8):
        par = (cls >> 2, (cls >> 1) & 1, cls & 1)
        per_axis = [dict((delta + 1, k) for k, delta in convT_class_taps(4, 1, pa)) for pa in par]   # t -> k
        for td, kd in per_axis[0].items():
            for th, kh in per_axis[1].items():
                for tw, kw in per_axis[2].items():
                    W[cls, :, (td * 3 + th) * 3 + tw, :] = w[:, :, kd, kh, kw]
    n = 8 * cout
    return pack_matrix(W.reshape(n, 27 * cin), b.repeat(8), device, block_n or (16 if n <= 16 else None), n)


CONVT_FUSED_TAPS = [(a - 1, b - 1, c - 1) for a in range(3) for b in range(3) for c in range(3)]


class Plan:
    """A recorded op list bound to fixed device buffers."""

    def __init__(self, device, lib=None, dtype=torch.float32):
        """dtype: storage type of the activations this plan allocates and operand type of its contractions
        (torch.float32: TF32 tensor cores; torch.bfloat16: bf16 operands, fp32 accumulation)"""
        self.lib = lib or _lib.get()
        self.device = torch.device(device)
        self.dtype = dtype
        assert dtype in (torch.float32, torch.bfloat16)
        self.handle = C.c_void_p(self.lib.svx_plan_create())
        if not self.handle:
            raise _lib.SvxError("svx_plan_create failed")
        self.keep = {}
        # activation-buffer reuse: `release(act)` after the last op that reads `act` has been recorded returns its memory to
        # the pool of the CURRENT lane; later `new_act` calls in the same lane draw from it (ops of one lane run in recording
        # order, so the new writer is ordered after the old readers; lanes overlap, so they never share a pool; a join
        # merges the side pools into lane 0's).
        self._lane = 0
        self._pools = {}          # lane -> list of free raw uint8 tensors
        self._pad_pools = {}      # lane -> {geometry key: [zero-bordered buffers whose producers wrote interiors only]}
        self._raw_of = {}         # id(view tensor) -> (raw uint8 tensor | None, geometry key | None)
        self.bytes_allocated = 0  # fresh device memory obtained for activations (diagnostics / tests)
        self.bytes_reused = 0
        self.op_names = []
        self.flops = []
        self.bytes = []   # algorithmic HBM bytes per op: every operand read once, every result written once
        self.taps = {}   # name -> Act of selected intermediates (parity tests / debugging)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.svx_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- memory ------------------------------------------------------------------------------
    def empty(self, *shape, dtype=torch.float32):
        return self.hold(torch.empty(*shape, dtype=dtype, device=self.device))

    def zeros(self, *shape, dtype=torch.float32):
        return self.hold(torch.zeros(*shape, dtype=dtype, device=self.device))

    def new_act(self, N, D, H, W, C, Cs=None, zero=False, pad=(0, 0, 0), dtype=None):
        """D, H, W are the data extents; `pad` adds a zero border that producers never write (flat-conv inputs)"""
        Cs = Cs or C
        dtype = dtype or self.dtype
        Dp, Hp, Wp = D + 2 * pad[0], H + 2 * pad[1], W + 2 * pad[2]
        rows = N * Dp * Hp * Wp
        if zero:                       # relies on zeros nobody rewrites (channel / column padding): never pooled
            buf = self.zeros(rows, Cs, dtype=dtype)
        elif any(pad):                 # zero border, interior-only producers: reusable by an identical geometry only
            key = (N, Dp, Hp, Wp, Cs, dtype, tuple(pad))
            free = self._pad_pools.get(self._lane, {}).get(key)
            if free:
                buf = free.pop()
                self.bytes_reused += buf.numel() * buf.element_size()
            else:
                buf = self.zeros(rows, Cs, dtype=dtype)
                self.bytes_allocated += buf.numel() * buf.element_size()
                self._raw_of[id(buf)] = (None, key)
        else:
            buf = self._pooled_empty(rows, Cs, dtype)
        return Act(buf, N, Dp, Hp, Wp, C, 0, tuple(pad))

    def _pooled_empty(self, rows, Cs, dtype):
        need = rows * Cs * esize_of(dtype)
        pool = self._pools.setdefault(self._lane, [])
        best = None
        for i, raw in enumerate(pool):   # best fit: the smallest free block that is large enough
            if raw.numel() >= need and (best is None or raw.numel() < pool[best].numel()):
                best = i
        if best is not None:
            raw = pool.pop(best)
            self.bytes_reused += need
        else:
            raw = torch.empty(max(need, 16), dtype=torch.uint8, device=self.device)
            self.bytes_allocated += raw.numel()
        buf = raw[:need].view(dtype).view(rows, Cs)
        self._raw_of[id(buf)] = (raw, None)
        return self.hold(buf)

    def release(self, act):
        """The last op reading `act` (a whole buffer from new_act) has been recorded in the CURRENT lane, and every other
        reader was recorded in this lane too (or before this lane forked): its memory may back later activations of this
        lane.  Tensors not created by new_act (module inputs, zero-initialised staging buffers) are ignored."""
        buf = act.buf if isinstance(act, Act) else act
        ent = self._raw_of.pop(id(buf), None)
        if ent is None:
            return
        raw, key = ent
        if raw is not None:
            self._pools.setdefault(self._lane, []).append(raw)
        else:
            self._pad_pools.setdefault(self._lane, {}).setdefault(key, []).append(buf)

    @staticmethod
    def _dt(x, out):
        """SVX_DT_* bits of a non-contraction op (x / out: Act or tensor)"""
        xb = (x.buf if isinstance(x, Act) else x).dtype == torch.bfloat16
        ob = (out.buf if isinstance(out, Act) else out).dtype == torch.bfloat16
        return (_lib.DT_IN_BF16 if xb else 0) | (_lib.DT_OUT_BF16 if ob else 0)

    def hold(self, t):
        """keep every tensor whose address is baked into an op alive as long as the plan"""
        if isinstance(t, Act):
            self.keep[id(t.buf)] = t.buf
        elif t is not None:
            self.keep[id(t)] = t
        return t

    # ---- execution ---------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def run(self, graph=False):
        _lib.check(self.lib.svx_plan_run(self.handle, self._stream(), 1 if graph else 0), self.lib)

    def run_range(self, first, last):
        _lib.check(self.lib.svx_plan_run_range(self.handle, first, last, self._stream()), self.lib)

    def time_ops(self, iters=5):
        n = self.num_ops
        ms = (C.c_float * n)()
        _lib.check(self.lib.svx_plan_time_ops(self.handle, self._stream(), iters, ms), self.lib)
        return list(ms)

    @property
    def num_ops(self):
        return self.lib.svx_plan_num_ops(self.handle)

    @property
    def num_launches(self):
        return self.lib.svx_plan_num_launches(self.handle)

    def _add(self, op, desc, name, flops=0.0, nbytes=0.0):
        add = getattr(self.lib, _lib.OPS[op][1])
        _lib.check(add(self.handle, C.byref(desc)), self.lib)
        self.op_names.append(name or op)
        self.flops.append(flops)
        self.bytes.append(float(nbytes))

    # ---- concurrency hints -----------------------------------------------------------------------
    def lane(self, k):
        """ops recorded from now on go to lane k (0 = the caller's stream, 1..8 = side streams / graph branches)"""
        _lib.check(self.lib.svx_plan_set_lane(self.handle, k), self.lib)
        self._lane = k

    def join(self):
        """the caller's stream waits for every side lane; also resets the current lane to 0"""
        self.lane(0)
        _lib.check(self.lib.svx_plan_add_join(self.handle), self.lib)
        for k in [k for k in self._pools if k != 0]:          # everything recorded so far precedes everything after the join
            self._pools.setdefault(0, []).extend(self._pools.pop(k))
        for k in [k for k in self._pad_pools if k != 0]:
            for key, bufs in self._pad_pools.pop(k).items():
                self._pad_pools.setdefault(0, {}).setdefault(key, []).extend(bufs)
        self.op_names.append("join")
        self.flops.append(0.0)
        self.bytes.append(0.0)

    # ---- contractions ------------------------------------------------------------------------
    def _taps_tensor(self, taps):
        t = torch.tensor([[a, b, c, 0] for a, b, c in taps], dtype=torch.int32, device=self.device)
        return self.hold(t)

    def _gather_operand(self, d, x, cin, taps, rows_dhw, stride):
        """A operand of an implicit-GEMM convolution over the unpadded channels-last tensor `x`: fetched by the TMA unit
        in im2col mode when the channel count allows whole 32-channel boxes (SVX_A_IM2COL), by cp.async gather warps
        otherwise (SVX_A_GATHER: the 4-channel image stems)."""
        rd, rh, rw = rows_dhw
        d.M = x.N * rd * rh * rw
        d.A = self.hold(x).buf.data_ptr()
        d.in_D, d.in_H, d.in_W, d.in_Cs, d.in_c0, d.Cin = x.D, x.H, x.W, x.Cs, x.c0, cin
        d.out_D, d.out_H, d.out_W = rd, rh, rw
        d.stride_d, d.stride_h, d.stride_w = stride
        d.ntaps = len(taps)
        lo = [min(t[a] for t in taps) for a in range(3)]
        up = [lo[a] + (rows_dhw[a] - 1) * stride[a] - ((x.D, x.H, x.W)[a] - 1) for a in range(3)]
        # whole 128-byte boxes, or 16-byte pixels (the image stems: eight taps per k-chunk) / 32-byte pixels (four taps)
        es = x.esize
        narrow_ok = cin * es == 16 or (cin * es == 32 and len(taps) % 4 == 0)
        tma = (((cin * es) % 128 == 0 or narrow_ok)
               and len(taps) <= 64 and all(-16 <= v <= 15 for v in lo + up))
        if tma:
            d.a_mode = A_IM2COL
            host = (C.c_int32 * (4 * len(taps)))(*[v for t in taps for v in (t[0], t[1], t[2], 0)])
            self.keep[id(host)] = host
            d.taps_host = C.cast(host, C.c_void_p)
        else:
            d.a_mode = A_GATHER
            d.taps = self._taps_tensor(taps).data_ptr()
        return tma

    def _fill_epilogue(self, d, pack, out, out_map, act, act_param, residual, res_after_act, out_scale, round_out,
                       res_via_mma=False, op_dtype=torch.float32):
        pack.finalize(rows=d.M, gather=(d.a_mode == A_GATHER), heavy_epilogue=(act == ACT_GELU), dtype=op_dtype)
        assert pack.dtype == op_dtype, "weights and the A operand must share the operand type"
        bf = op_dtype == torch.bfloat16
        d.operand_kind = _lib.OPERAND_BF16 if bf else d.operand_kind
        if out.bf16:
            assert bf, "bf16 outputs come from bf16-operand contractions"
            d.io_flags |= _lib.IO_OUT_BF16
        d.W = pack.W.data_ptr()
        if res_via_mma:
            # the residual is added by the tensor cores: block_n identity columns appended to every weight row
            assert residual is not None and not res_after_act and pack.N % pack.block_n == 0
            if getattr(pack, "W_ext", None) is None:
                eye = torch.zeros(pack.Npad, pack.block_n, dtype=pack.W.dtype, device=pack.W.device)
                idx = torch.arange(pack.Npad, device=pack.W.device)
                eye[idx, idx % pack.block_n] = 1.0
                pack.W_ext = torch.cat([pack.W, eye], dim=1).contiguous()
            d.W = self.hold(pack.W_ext).data_ptr()
            d.res_via_mma = 1
        d.bias = pack.bias.data_ptr()
        d.N, d.K, d.Kpad, d.Npad, d.block_n = pack.N, pack.K, pack.Kpad, pack.Npad, pack.block_n
        d.out = self.hold(out).buf.data_ptr()
        if out_map is None:
            out_map = out.interior_map()
        d.o_base, d.o_sn, d.o_sd, d.o_sh, d.o_sw = out_map
        d.act, d.act_param = act, act_param
        d.out_scale = out_scale
        d.round_tf32 = 1 if round_out else 0
        d.epi_mode = EPI_STD
        if residual is not None:
            assert (residual.Cs == out.Cs and residual.c0 == out.c0 and residual.pixels == out.pixels
                    and residual.pad == out.pad and residual.dtype == out.dtype), "residual must share the output layout"
            d.residual = self.hold(residual).buf.data_ptr()
            if residual.bf16:
                d.io_flags |= _lib.IO_RES_BF16
            d.res_after_act = 1 if res_after_act else 0
        self.hold(pack.W)
        self.hold(pack.bias)

    def linear(self, x, pack, out, act=ACT_NONE, act_param=0.0, residual=None, res_after_act=True,
               out_scale=1.0, round_out=False, name=None, pool8=False, res_via_mma=False):
        """x: Act read as a [pixels, C] matrix (plain TMA operand); out: Act with C == pack.N.
        res_via_mma: the (pre-activation) residual holds TF32-exact values and is added by the tensor cores.
        pool8: the pack's N columns are 8 groups of N/8 channels (the conv positions under one MaxPool3d(2) window);
        the epilogue stores act(max over groups + bias) into out (C == pack.N / 8)."""
        oD, oH, oW = out.inner
        assert x.C == pack.K and x.pixels == out.N * oD * oH * oW, (x.C, pack.K, out.C, pack.N)
        assert (out.C * 8 == pack.N) if pool8 else (out.C >= pack.N), (out.C, pack.N)
        e16 = 16 // x.esize
        assert x.c0 % e16 == 0 and x.Cs % e16 == 0 and not any(x.pad), "plain operands are dense matrices"
        d = _lib.GemmDesc()
        d.M = x.pixels
        d.a_mode = A_PLAIN
        d.A = self.hold(x).buf.data_ptr() + x.esize * x.c0
        d.lda = x.Cs
        d.out_D, d.out_H, d.out_W = oD, oH, oW
        self._fill_epilogue(d, pack, out, None, act, act_param, residual, res_after_act, out_scale, round_out,
                            res_via_mma=res_via_mma, op_dtype=x.dtype)
        if pool8:
            d.epi_mode = EPI_POOL8
        n_out = pack.N // 8 if pool8 else pack.N
        self._add("gemm", d, name or "linear", 2.0 * d.M * pack.N * pack.K,
                  x.esize * (d.M * pack.K + pack.N * pack.K) + out.esize * d.M * n_out * (2 if residual is not None else 1))
        return out

    def mlp(self, x, pack1, pack2, out, residual, round_out=False, name=None, ln=None):
        """timm Mlp (fc1 -> GELU -> fc2) + the block's second residual as one kernel (svx_mlp_desc): the hidden activation
        stays on the SM.  x: TF32-rounded norm2 output [pixels, C]; pack1 / pack2: fc1 [4C, C] / fc2 [C, 4C] packs."""
        Cc, hid = pack1.K, pack1.N
        assert mlp_fusable(Cc, hid) and pack2.K == hid and pack2.N == Cc and x.C == Cc and out.C == Cc
        assert x.pixels == out.pixels == residual.pixels and not any(x.pad) and not any(out.pad) and not any(residual.pad)
        assert x.c0 == 0 and out.c0 == 0 and residual.c0 == 0 and residual.Cs == out.Cs
        for pk in (pack1, pack2):   # unpadded [N, K] matrices: the kernel has its own tiling
            if pk.W is None:
                pk.block_n = 32
                pk.finalize()
            assert tuple(pk.W.shape) == (pk.N, pk.K), "mlp: packs must not be padded"
        d = _lib.MlpDesc()
        d.x, d.ldx = self.hold(x).buf.data_ptr(), x.Cs
        d.W1, d.b1 = self.hold(pack1.W).data_ptr(), self.hold(pack1.bias).data_ptr()
        d.W2, d.b2 = self.hold(pack2.W).data_ptr(), self.hold(pack2.bias).data_ptr()
        d.residual, d.out, d.ldo = self.hold(residual).buf.data_ptr(), self.hold(out).buf.data_ptr(), out.Cs
        d.M, d.C, d.hidden = x.pixels, Cc, hid
        d.round_tf32 = 1 if round_out else 0
        if ln is not None:   # (gamma, beta, eps): x holds the un-normalised rows, the kernel applies the block's norm2 itself
            assert mlp_ln_fusable(Cc)
            d.ln_gamma, d.ln_beta, d.ln_eps = self.hold(ln[0]).data_ptr(), self.hold(ln[1]).data_ptr(), float(ln[2])
        self._add("mlp", d, name or "mlp", 4.0 * d.M * Cc * hid,
                  4.0 * ((2 if ln is not None and x is residual else 3) * d.M * Cc + 2 * Cc * hid))
        return out

    def conv(self, x, pack, taps, out, stride=(1, 1, 1), out_map=None, rows_dhw=None, act=ACT_NONE, act_param=0.0,
             residual=None, res_after_act=True, out_scale=1.0, round_out=False, name=None, epi_tail=None):
        """Implicit-GEMM convolution.  rows_dhw: extents the GEMM rows run over (default: out D,H,W)."""
        cin_pad = pack.K // len(taps)
        e16 = 16 // x.esize
        assert cin_pad * len(taps) == pack.K and cin_pad % e16 == 0 and cin_pad >= x.C
        assert x.c0 % e16 == 0 and x.Cs % e16 == 0 and x.c0 + cin_pad <= x.Cs, (x.c0, cin_pad, x.Cs)
        assert not any(x.pad), "gather mode reads unpadded tensors (use conv_flat for padded ones)"
        d = _lib.GemmDesc()
        self._gather_operand(d, x, cin_pad, taps, rows_dhw or out.inner, stride)
        self._fill_epilogue(d, pack, out, out_map, act, act_param, residual, res_after_act, out_scale, round_out,
                            op_dtype=x.dtype)
        if epi_tail is not None:
            aux, out2, map2 = epi_tail
            d.epi_mode = EPI_DEC_TAIL
            d.epi_aux = self.hold(aux).data_ptr()
            d.epi_out2 = self.hold(out2).data_ptr()
            d.o2_base, d.o2_sn, d.o2_sd, d.o2_sh, d.o2_sw = map2
        self._add("gemm", d, name or "conv", 2.0 * d.M * pack.N * pack.K,
                  x.esize * (x.pixels * x.C + pack.N * pack.K) + out.esize * d.M * pack.N * (2 if residual is not None else 1))
        return out

    def convT_fused(self, x, pack, out, act=ACT_NONE, act_param=0.0, residual=None, res_after_act=True, out_scale=1.0,
                    round_out=False, tail=None, name=None):
        """ConvTranspose3d(k4, s2, p1) with the eight parity classes in N (SVX_EPI_CONVT8): one implicit GEMM whose rows
        are the INPUT voxels of `x` (gathered 3x3x3 neighbourhood) and whose epilogue scatters column (cls, c) to
        output voxel (2d+pd, 2h+ph, 2w+pw).  `pack` from pack_convT_fused.  tail = (aux[9], coarse tensor [N, OD*OH*OW]):
        the decoder's layer5 + cat (8 channels per class)."""
        cout = pack.N // 8
        assert pack.N == 8 * cout and pack.K == 27 * x.C and x.C % 4 == 0 and not any(x.pad)
        od, oh, ow = out.inner
        assert (od, oh, ow) == (2 * x.D, 2 * x.H, 2 * x.W) and out.N == x.N and out.C >= (9 if tail else cout)
        d = _lib.GemmDesc()
        self._gather_operand(d, x, x.C, CONVT_FUSED_TAPS, (x.D, x.H, x.W), (1, 1, 1))
        ob, osn, osd, osh, osw = out.interior_map()
        self._fill_epilogue(d, pack, out, (ob, osn, 2 * osd, 2 * osh, 2 * osw), act, act_param, residual, res_after_act,
                            out_scale, round_out)
        d.epi_mode = EPI_CONVT8
        d.cls_cout = cout
        d.c_sd, d.c_sh, d.c_sw = osd, osh, osw
        if tail is not None:
            aux, out2 = tail
            assert cout == 8 and residual is None and act == ACT_RELU
            d.epi_aux = self.hold(aux).data_ptr()
            d.epi_out2 = self.hold(out2).data_ptr()
            d.o2_base, d.o2_sn, d.o2_sd, d.o2_sh, d.o2_sw = 0, od * oh * ow, 2 * oh * ow, 2 * ow, 2
            d.c2_sd, d.c2_sh, d.c2_sw = oh * ow, ow, 1
        # useful work only: 8 of the 27 taps feed each class
        self._add("gemm", d, name or "convT_fused", 2.0 * d.M * pack.N * 8 * x.C,
                  4.0 * (x.pixels * x.C + d.M * (8 * 10 if tail else pack.N * (2 if residual is not None else 1))
                         + pack.N * pack.K))
        return out

    def conv_flat(self, x, pack, taps, out, act=ACT_NONE, act_param=0.0, residual=None, res_after_act=True,
                  out_scale=1.0, round_out=False, name=None, out_map=None, valid=None):
        """Stride-1 convolution over a zero-padded input, A operand streamed by TMA (one shifted box per tap).
        `taps` are non-negative (dd, dh, dw) offsets from the window corner in padded coordinates; `valid` =
        number of window corners per axis that are real outputs (default: the extents of `out`)."""
        cin = pack.K // len(taps)
        assert cin * len(taps) == pack.K and (cin * x.esize) % 128 == 0 and cin == x.C, (cin, x.C, pack.K)
        assert all(min(t) >= 0 for t in taps)
        vD, vH, vW = valid or out.inner
        d = _lib.GemmDesc()
        rows_total = x.buf.shape[0]
        per = x.D * x.H * x.W
        d.M = (x.N - 1) * per + ((vD - 1) * x.H + (vH - 1)) * x.W + vW
        d.a_mode = A_FLAT
        d.A = self.hold(x).buf.data_ptr()
        d.lda = rows_total
        d.in_D, d.in_H, d.in_W, d.in_Cs, d.in_c0, d.Cin = x.D, x.H, x.W, x.Cs, x.c0, cin
        d.out_D, d.out_H, d.out_W = x.D, x.H, x.W
        d.valid_D, d.valid_H, d.valid_W = vD, vH, vW
        d.stride_d = d.stride_h = d.stride_w = 1
        d.ntaps = len(taps)
        host = (C.c_int32 * (4 * len(taps)))(*[v for t in taps for v in (t[0], t[1], t[2], 0)])
        self.keep[id(host)] = host
        d.taps_host = C.cast(host, C.c_void_p)
        self._fill_epilogue(d, pack, out, out_map, act, act_param, residual, res_after_act, out_scale, round_out,
                            op_dtype=x.dtype)
        assert pack.Kpad == pack.K
        self._add("gemm", d, name or "conv_flat", 2.0 * x.N * vD * vH * vW * pack.N * pack.K,
                  x.esize * (x.N * vD * vH * vW * cin + pack.N * pack.K)
                  + out.esize * x.N * vD * vH * vW * pack.N * (2 if residual is not None else 1))
        return out

    def conv3_slab(self, x, pack, out, cin_live, act=ACT_NONE, act_param=0.0, residual=None, res_after_act=True,
                   out_scale=1.0, round_out=False, name=None, operands="fp16", range_flag=None):
        """Conv3d(k3, s1, p1) with <= 16 output channels over a (1,1,1) zero-bordered volume `x` (an Act whose
        channel window [c0, c0+32) is the TMA box; only the first `cin_live` channels carry weights).  `pack` comes
        from pack_conv3_slab.  `out` may live in any buffer whose Act describes the same voxel grid."""
        assert x.pad == (1, 1, 1) and pack.block_n == 48 and pack.Npad == 48 and pack.Kpad == 288
        vD, vH, vW = x.inner
        assert tuple(out.inner) == (vD, vH, vW) and out.N == x.N and out.C >= pack.N
        d = _lib.GemmDesc()
        d.M = x.N * vD * vH * vW
        d.a_mode = A_SLAB3
        d.A = self.hold(x).buf.data_ptr()
        d.lda = x.buf.shape[0]
        d.in_D, d.in_H, d.in_W, d.in_Cs, d.in_c0, d.Cin = x.D, x.H, x.W, x.Cs, x.c0, 32
        d.out_D, d.out_H, d.out_W = vD, vH, vW
        d.valid_D, d.valid_H, d.valid_W = vD, vH, vW
        d.stride_d = d.stride_h = d.stride_w = 1
        d.ntaps = 27
        d.cin_live = cin_live
        d.acc_scale = getattr(pack, "acc_scale", 1.0)
        # operands: "fp16" (default: half the MMA instructions; activations beyond 65504 saturate and set range_flag) or
        # "tf32" (fp32's exponent range)
        assert operands in ("fp16", "tf32")
        d.operand_kind = _lib.OPERAND_TF32 if operands == "tf32" else _lib.OPERAND_DEFAULT
        if range_flag is not None and operands == "fp16":
            d.range_flag = self.hold(range_flag).data_ptr()
        self._fill_epilogue(d, pack, out, None, act, act_param, residual, res_after_act, out_scale, round_out)
        self._add("gemm", d, name or "conv3_slab", 2.0 * d.M * 27 * cin_live * pack.N,
                  4.0 * d.M * (cin_live + pack.N * (2 if residual is not None else 1)))
        return out

    # ---- everything else ------------------------------------------------------------------------
    def im2col(self, src, strides, N, Cin, dhw, kernel, stride, pads, out_dhw, kpad, round_out=True, name=None):
        """src: tensor read in place with element strides (s_n, s_c, s_d, s_h, s_w)."""
        OD, OH, OW = out_dhw
        out = self.new_act(N, OD, OH, OW, kpad)
        d = _lib.Im2colDesc()
        d.inp, d.out = self.hold(src).data_ptr(), out.buf.data_ptr()
        d.N, d.C, (d.D, d.H, d.W) = N, Cin, dhw
        d.s_n, d.s_c, d.s_d, d.s_h, d.s_w = strides
        d.KD, d.KH, d.KW = kernel
        d.stride = stride
        d.pad_d, d.pad_h, d.pad_w = pads
        d.OD, d.OH, d.OW, d.Kpad = OD, OH, OW, kpad
        d.round_tf32 = 1 if round_out else 0
        self._add("im2col", d, name, 0.0, 4.0 * (N * Cin * dhw[0] * dhw[1] * dhw[2] + out.pixels * kpad))
        return out

    def pool(self, x, out, kernel, stride, pads, mode, round_out=False, name=None):
        d = _lib.PoolDesc()
        assert x.c0 == 0 and out.c0 == 0, "pool works on whole buffers"
        d.inp, d.out = self.hold(x).buf.data_ptr(), self.hold(out).buf.data_ptr()
        d.N, d.C, d.D, d.H, d.W, d.in_Cs, d.out_Cs = x.N, x.C, x.D, x.H, x.W, x.Cs, out.Cs
        d.KD, d.KH, d.KW = kernel
        d.SD, d.SH, d.SW = stride
        d.PD, d.PH, d.PW = pads
        d.OD, d.OH, d.OW = out.D, out.H, out.W
        d.mode = mode
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(x, out)
        self._add("pool", d, name, 0.0, x.esize * x.C * (x.pixels + out.pixels))
        return out

    def layernorm_rows(self, x, gamma, beta, out, merge_hw=None, eps=1e-5, round_out=True, name=None):
        d = _lib.LnRowsDesc()
        d.inp, d.out = self.hold(x).buf.data_ptr(), self.hold(out).buf.data_ptr()
        d.gamma, d.beta = self.hold(gamma).data_ptr(), self.hold(beta).data_ptr()
        d.rows, d.C = out.pixels, out.C
        if merge_hw:
            d.merge, d.H, d.W = 1, merge_hw[0], merge_hw[1]
        d.eps = eps
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(x, out)
        self._add("layernorm_rows", d, name, 0.0, 2.0 * x.esize * out.pixels * out.C)
        return out

    def layernorm_sample(self, x, gamma, beta, out, eps=1e-5, round_out=True, name=None):
        d = _lib.LnSampleDesc()
        d.inp, d.out = self.hold(x).buf.data_ptr(), self.hold(out).buf.data_ptr()
        d.gamma, d.beta = self.hold(gamma).data_ptr(), self.hold(beta).data_ptr()
        d.N, d.L = x.N, x.D * x.H * x.W * x.Cs
        d.eps = eps
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(x, out)
        self._add("layernorm_sample", d, name, 0.0, 2.0 * x.esize * d.N * d.L)
        return out

    def window_attention(self, qkv, out, bias, H, W, heads, shift, scale, round_out=True, name=None, range_flag=None):
        """range_flag: optional device int32[1] the fp32-storage kernel sets when a V value exceeded fp16's finite range
        (its P V product runs on fp16 operands)"""
        d = _lib.WinAttnDesc()
        if range_flag is not None and not qkv.bf16:
            d.range_flag = self.hold(range_flag).data_ptr()
        d.qkv, d.out, d.bias = self.hold(qkv).buf.data_ptr(), self.hold(out).buf.data_ptr(), self.hold(bias).data_ptr()
        d.N, d.H, d.W, d.C, d.heads, d.shift, d.scale = qkv.N, H, W, out.C, heads, shift, scale
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(qkv, out)
        self._add("window_attention", d, name, 4.0 * qkv.N * H * W * 49 * out.C, 4.0 * qkv.esize * qkv.N * H * W * out.C)
        return out

    def dwconv(self, x, w, bias, out, k, round_out=True, name=None):
        d = _lib.DwConvDesc()
        d.inp, d.out, d.w = self.hold(x).buf.data_ptr(), self.hold(out).buf.data_ptr(), self.hold(w).data_ptr()
        d.bias = self.hold(bias).data_ptr() if bias is not None else None
        d.N, d.H, d.W, d.C, d.k, d.OH, d.OW = x.N, x.H, x.W, x.C, k, out.H, out.W
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(x, out)
        self._add("dwconv", d, name)
        return out

    def view_attention(self, qkv, out, B, V, heads, scale, round_out=True, name=None):
        d = _lib.ViewAttnDesc()
        d.qkv, d.out = self.hold(qkv).buf.data_ptr(), self.hold(out).buf.data_ptr()
        d.B, d.V, d.P, d.R, d.heads, d.scale = B, V, qkv.H * qkv.W, out.C, heads, scale
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(qkv, out)
        self._add("view_attention", d, name)
        return out

    def bilinear_add(self, x, skip, out, round_out=True, name=None):
        d = _lib.BilinearDesc()
        d.inp, d.out = self.hold(x).buf.data_ptr(), self.hold(out).buf.data_ptr()
        d.skip = self.hold(skip).buf.data_ptr() if skip is not None else None
        d.N, d.IH, d.IW, d.OH, d.OW, d.C = x.N, x.H, x.W, out.H, out.W, x.C
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(x, out)
        assert skip is None or skip.dtype == x.dtype
        self._add("bilinear_add", d, name)
        return out

    def conv3_to1(self, x, weight, bias, bn, out, slope, name=None):
        """Conv3d(Cin <= 12 -> 1, k3, p1) + BatchNorm (folded) + LeakyReLU in fp32 on the CUDA cores, over the (1,1,1)
        zero-bordered volume `x` (channels [c0, c0 + Cin)); out: planar [N, D*H*W] tensor"""
        assert x.pad == (1, 1, 1)
        w, b = fold_bn(weight, bias, bn)                     # [1, Cin, 3, 3, 3]
        cin = w.shape[1]
        W = torch.zeros(27, 12, dtype=torch.float32, device=w.device)
        W[:, :cin] = w[0].permute(1, 2, 3, 0).reshape(27, cin)
        d = _lib.Conv3to1Desc()
        d.inp = self.hold(x).buf.data_ptr()
        d.w, d.bias = self.hold(W.contiguous().to(self.device)).data_ptr(), self.hold(b.float().to(self.device)).data_ptr()
        d.out = self.hold(out).data_ptr()
        vD, vH, vW = x.inner
        d.N, d.D, d.H, d.W, d.Cs, d.c0, d.Cin, d.slope = x.N, vD, vH, vW, x.Cs, x.c0, cin, slope
        M = x.N * vD * vH * vW
        self._add("conv3to1", d, name, 2.0 * M * 27 * cin, 4.0 * M * (cin + 1))
        return out

    def merger_fuse(self, weights, coarse, out, B, V, P, name=None):
        """weights None: the mean over the views (core/test.py:125-126)"""
        d = _lib.MergeFuseDesc()
        d.weights = self.hold(weights).data_ptr() if weights is not None else None
        d.coarse, d.out = self.hold(coarse).data_ptr(), self.hold(out).data_ptr()
        d.B, d.V, d.P = B, V, P
        self._add("merger_fuse", d, name, 0.0, 4.0 * ((2 if weights is not None else 1) * V + 1) * B * P)
        return out

    def voxel_metrics(self, logits, gt, thresholds, counts, B, P, name=None, bce=None):
        d = _lib.MetricsDesc()
        if bce is not None:
            d.bce_q20 = self.hold(bce).data_ptr()
        d.logits, d.gt, d.prob_thresholds, d.counts = (self.hold(logits).data_ptr(), self.hold(gt).data_ptr(),
                                                       self.hold(thresholds).data_ptr(), self.hold(counts).data_ptr())
        d.B, d.P, d.T = B, P, thresholds.numel()
        self._add("voxel_metrics", d, name, 0.0, 8.0 * B * P)
        return counts

    def resize_bilinear(self, src, dst, name=None):
        """planar fp32 [N, C, IH, IW] -> [N, C, OH, OW], F.interpolate(mode="bilinear", align_corners=False)"""
        d = _lib.ResizeDesc()
        d.inp, d.out = self.hold(src).data_ptr(), self.hold(dst).data_ptr()
        d.NC, d.IH, d.IW, d.OH, d.OW = src.shape[0] * src.shape[1], src.shape[2], src.shape[3], dst.shape[2], dst.shape[3]
        self._add("resize_bilinear", d, name or "resize", 0.0, 4.0 * (src.numel() + dst.numel()))
        return dst

    def transpose(self, src, dst, N, Cc, P, Cs, to_channels_last, round_out=False, name=None, rows=None):
        """rows = (row_w, row_pitch, row_x0): write the 4-channel channels-last image with zero columns around each row"""
        d = _lib.TransposeDesc()
        if rows is not None:
            d.row_w, d.row_pitch, d.row_x0 = rows
        d.inp, d.out = self.hold(src).data_ptr(), self.hold(dst).data_ptr()
        d.N, d.C, d.P, d.Cs = N, Cc, P, Cs
        d.to_channels_last = 1 if to_channels_last else 0
        d.round_tf32 = 1 if round_out else 0
        d.dtype = self._dt(src, dst)
        self._add("transpose", d, name, 0.0, N * P * (src.element_size() * Cc + dst.element_size() * Cs))
        return dst


def transpose_now(plan, src, dst, N, Cc, P, Cs, to_channels_last, round_out=False):
    """immediate (non-recorded) layout change on the plan's stream: module-boundary conversions"""
    d = _lib.TransposeDesc()
    d.inp, d.out = src.data_ptr(), dst.data_ptr()
    d.N, d.C, d.P, d.Cs = N, Cc, P, Cs
    d.to_channels_last = 1 if to_channels_last else 0
    d.round_tf32 = 1 if round_out else 0
    _lib.check(plan.lib.svx_transpose(C.byref(d), plan._stream()), plan.lib)

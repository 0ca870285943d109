/*
 * swinvox_b200.h -- C-ABI of libswinvox_b200.so (sm_100a).
 *
 * The reference (SwinVox) has no FFI of its own: its boundary for the hot path is the
 * Python nn.Module API of models/{encoder,swin_transformer,cross_view_attention,decoder,
 * merger,refiner}.py plus the metric loop core/test.py:141-164.  The Python shells in
 * swinvox_b200/models/*.py keep that API and drive this library through ctypes.  Each
 * entry point below cites the reference operation it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (PyTorch).  The library allocates nothing persistent besides the plan object itself.
 *   - activations are channels-last ([N,H,W,C] / [N,D,H,W,C]), fp32 by default: contractions then run on tcgen05
 *     kind::tf32 with fp32 accumulation in TMEM.  The encoder can alternatively be lowered with bf16 activation
 *     storage (SVX_OPERAND_BF16 / SVX_DT_* flags): kind::f16 with bf16 operands, fp32 accumulation and statistics.
 *   - every function returns 0 on success, non-zero on failure; svx_last_error() returns a
 *     thread-local message.  No exceptions cross the ABI.
 *   - launches are stream ordered on the cudaStream_t passed as `void* stream`.
 */
#ifndef SWINVOX_B200_H
#define SWINVOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVX_ABI_VERSION 2

/* activation codes for svx_gemm_desc.act */
enum { SVX_ACT_NONE = 0, SVX_ACT_RELU = 1, SVX_ACT_LEAKY = 2, SVX_ACT_GELU = 3 };
/* A operand modes */
enum { SVX_A_PLAIN = 0, SVX_A_GATHER = 1, SVX_A_FLAT = 2, SVX_A_SLAB3 = 3, SVX_A_IM2COL = 4 };
/* special epilogues */
enum {
  SVX_EPI_STD = 0,
  SVX_EPI_DEC_TAIL = 1, /* decoder layer4+layer5+cat, decoder.py:80-89 */
  SVX_EPI_POOL8 = 2,    /* conv + BN + LeakyReLU + MaxPool3d(2) (refiner.py:21-26): the N columns are 8 groups (the 2x2x2
                           conv positions of one pooled voxel) of N/8 channels; out[r, c] = act(max_g acc[r, g*N/8 + c] +
                           bias[c]) -- valid because the activation is monotonic and the bias is shared by the group.
                           Plain row-major output [M, N/8] only (o_sw = row pitch). */
  SVX_EPI_CONVT8 = 3    /* stride-2 ConvTranspose3d(k4, p1) (decoder.py:37-46, refiner.py:62-70) with all eight output-parity
                           classes of one INPUT voxel in the N dimension: row r = input voxel (n,d,h,w), K = the 3x3x3
                           input neighbourhood x Cin (structural zeros where a tap does not feed a class), column
                           j = cls*cls_cout + c with cls = pd*4 + ph*2 + pw.  Column j is stored at
                             out + o_base + n*o_sn + d*o_sd + h*o_sh + w*o_sw + pd*c_sd + ph*c_sh + pw*c_sw + c
                           (o_s* = twice the output voxel strides, c_s* = the output voxel strides); `residual` uses the same
                           mapping.  With epi_aux set (cls_cout == 8) every class also gets the decoder tail of
                           SVX_EPI_DEC_TAIL: channels 0-7 = relu, channel 8 = layer5, which also goes to
                           epi_out2 + o2_base + ... + pd*c2_sd + ph*c2_sh + pw*c2_sw. */
};
/* svx_gemm_desc.operand_kind: element type of the MMA operands */
enum {
  SVX_OPERAND_DEFAULT = 0, /* fp32 storage read as TF32 (kind::tf32); SVX_A_SLAB3: fp16 operands converted in shared memory */
  SVX_OPERAND_TF32 = 1,    /* SVX_A_SLAB3 only: keep kind::tf32 operands (full fp32 exponent range, twice the MMA instructions) */
  SVX_OPERAND_BF16 = 2     /* A, W (and a res_via_mma residual) are bf16 in memory: kind::f16 with bf16 operands, fp32 accumulation */
};
/* svx_gemm_desc.io_flags: storage type of the epilogue tensors (default fp32) */
enum { SVX_IO_OUT_BF16 = 1, SVX_IO_RES_BF16 = 2 };
/* `dtype` field of the non-contraction descriptors: storage type of the activation tensors (default: fp32 in, fp32 out).
 * Arithmetic (statistics, softmax, accumulation) is fp32 either way; gamma / beta / bias / relative-position tables stay fp32. */
enum { SVX_DT_IN_BF16 = 1, SVX_DT_OUT_BF16 = 2, SVX_DT_BF16 = 3 };
/* pooling modes */
enum { SVX_POOL_MAX = 0, SVX_POOL_AVG = 1 };

/*
 * One tensor-core contraction  out[r, j] = act( sum_k A[r,k] * W[j,k] + bias[j] ) (+ residual[r,j])
 * Replaces every nn.Linear / nn.Conv2d / nn.Conv3d / nn.ConvTranspose3d of the reference
 * (encoder.py:22-111, timm Swin linears, cross_view_attention.py:38-53, decoder.py:24-46,
 * merger.py:20-54, refiner.py:21-70) with BatchNorm (eval) folded into W/bias by the host.
 *
 * A operand:
 *   SVX_A_PLAIN : row-major [M, K] with row stride lda (elements), streamed by TMA.
 *   SVX_A_GATHER: implicit im2col.  Row r decodes to (n, od, oh, ow) over out_{D,H,W};
 *                 k = tap*Cin + c; the element read is
 *                   in[n, od*stride_d + taps[tap].d, oh*stride_h + taps[tap].h,
 *                         ow*stride_w + taps[tap].w, in_c0 + c]
 *                 of a channels-last tensor with pixel stride in_Cs, zero outside the tensor.
 *                 Cin, in_c0, in_Cs must be multiples of 4 (16-byte cp.async chunks).
 *   SVX_A_IM2COL: the same implicit im2col as SVX_A_GATHER (same row / k / tap meaning, taps given in taps_host),
 *                 but fetched by the TMA unit in im2col mode (cuTensorMapEncodeIm2col over the NDHWC tensor, one
 *                 128-pixel x 32-channel box per (tap, channel chunk), borders zero-filled by the hardware).  Needs
 *                 Cin % 32 == 0, or Cin == 4 (image stems: one 16-byte pixel per tap, eight taps per k-chunk, Kpad may
 *                 exceed K), or Cin == 8 (pixel pairs: one 32-byte box per tap, four taps per k-chunk, K % 32 == 0); tap offsets and the implied padding must fit the descriptor's [-16, 15] corner range.
 *   SVX_A_FLAT  : stride-1 convolution over a zero-PADDED channels-last tensor viewed as the matrix
 *                 [N*in_D*in_H*in_W, in_Cs] (in_* are the padded extents).  Row r is the flat padded position
 *                 of the window corner; tap t reads row r + (dd*in_H + dh)*in_W + dw (taps >= 0), streamed by
 *                 TMA one 32-channel chunk at a time.  Cin must be a multiple of 32.  Rows decode over
 *                 out_{D,H,W} = in_{D,H,W}; only rows with od < valid_D, oh < valid_H, ow < valid_W are stored.
 *   SVX_A_SLAB3 : 3x3x3 stride-1 convolution with N <= 16 output channels over a zero-bordered channels-last
 *                 volume [vol, in_D, in_H, in_W, in_Cs] (in_D = valid_D + 2; lda = total rows), reading the 32
 *                 channels starting at in_c0 (a box that overhangs in_Cs reads zeros); output voxel (vol, d, h, w),
 *                 d < valid_D, h < valid_H, w < valid_W, sums taps (d+kd, h+kh, w+kw).  M = vol*valid_D*valid_H*
 *                 valid_W.  Weights are laid out for the kw-in-N formulation: W is [48, 288] with
 *                 row = kw*16 + co, col = (kd*3 + kh)*32 + c; block_n = Npad = 48, K = Kpad = 288; cin_live =
 *                 leading channels with non-zero weights (contraction steps beyond them are skipped).
 *                 The MMA operands of this mode are fp16 (kind::f16, fp32 accumulation): the kernel converts the fp32
 *                 slabs and weights in shared memory.  For the TF32-rounded values this path stores (10 mantissa bits =
 *                 fp16's) the conversion is exact for |x| in [6.1e-5, 65504]; below that the absolute error is < 3e-8
 *                 (fp16 subnormals), above it the value SATURATES to +-65504 and the kernel sets *range_flag (if given)
 *                 so the caller can re-run with operand_kind = SVX_OPERAND_TF32.  BN-folded weights of any magnitude
 *                 are handled by the host: W is pre-multiplied by a power of two that brings max|W| to [2^13, 2^14)
 *                 and acc_scale holds its inverse (applied to the accumulator before the bias).
 * W: [Npad, Kpad] fp32, K contiguous, zero padded, values pre-rounded to TF32 (rna) by the host.
 * result = out_scale * (res_after_act ? act(acc+bias) + res : act(acc+bias+res)).
 * Output row r is stored at out + o_base + n*o_sn + od*o_sd + oh*o_sh + ow*o_sw (elements),
 * columns contiguous.  `residual` uses the same mapping.
 */
typedef struct svx_gemm_desc {
  int32_t M, N, K;          /* logical sizes; N columns are written */
  int32_t Kpad, Npad;       /* padded weight extents (Kpad % 32 == 0, Npad % block_n == 0) */
  int32_t block_n;          /* N tile: 16, 32, 64, 96, 128, 192 or 256 (48 in slab mode) */
  int32_t a_mode;
  const void* A;            /* fp32, or bf16 with SVX_OPERAND_BF16 */
  int64_t lda;              /* plain: row stride (elements); flat: number of rows of the padded matrix */
  /* gather description */
  int32_t in_D, in_H, in_W, in_Cs, in_c0, Cin;
  int32_t out_D, out_H, out_W;
  int32_t stride_d, stride_h, stride_w;
  int32_t ntaps;
  const int32_t* taps;      /* device int32[ntaps][4] = {dd, dh, dw, 0} (gather mode) */
  const int32_t* taps_host; /* the same table in host memory (flat mode; read at plan-build time only) */
  int32_t valid_D, valid_H, valid_W; /* flat mode: extents of the rows that are real outputs (0 = all) */
  /* weights / epilogue */
  const void* W;            /* fp32 (TF32-rounded), or bf16 with SVX_OPERAND_BF16 (then Kpad % 64 == 0) */
  const float* bias;        /* [Npad] or NULL (always fp32) */
  const void* residual;     /* or NULL; fp32, or bf16 with SVX_IO_RES_BF16 */
  void* out;                /* fp32, or bf16 with SVX_IO_OUT_BF16 */
  int64_t o_base, o_sn, o_sd, o_sh, o_sw;
  int32_t act;
  float act_param;          /* LeakyReLU slope */
  int32_t res_after_act;    /* 0: act(acc+bias+res) (ResNet); 1: act(acc+bias)+res (Swin, refiner skips) */
  float out_scale;          /* applied last (refiner.py:103 uses 0.5); 1.0 otherwise */
  int32_t round_tf32;       /* round the stored result to TF32 (it feeds another contraction) */
  int32_t epi_mode;
  const float* epi_aux;     /* SVX_EPI_DEC_TAIL: layer5 weights w5[0..7], bias w5[8] */
  float* epi_out2;          /* SVX_EPI_DEC_TAIL: planar coarse volume [n, OD*OH*OW] */
  int64_t o2_base, o2_sn, o2_sd, o2_sh, o2_sw;
  int32_t cin_live;         /* slab mode: only the first cin_live of the 32 box channels carry non-zero weights
                               (0 = all); the kernel skips the contraction steps beyond them */
  int32_t res_via_mma;      /* plain mode, pre-activation residual holding TF32-exact values: add it on the tensor cores.
                               W then is [Npad, Kpad + block_n]: the extra columns of row n are one-hot at n % block_n */
  int32_t cls_cout;         /* SVX_EPI_CONVT8: output channels per parity class (1, 2, 4, 8 or a multiple of 16) */
  float acc_scale;          /* SVX_A_SLAB3: the accumulator is multiplied by this before the bias is added (0 = 1) */
  int64_t c_sd, c_sh, c_sw;     /* SVX_EPI_CONVT8: element offsets of the class bits in out / residual */
  int64_t c2_sd, c2_sh, c2_sw;  /* ... and in epi_out2 */
  int32_t* range_flag;      /* SVX_A_SLAB3 with fp16 operands: device int32, OR-ed with 1 when an activation saturated; or NULL */
  int32_t operand_kind;     /* SVX_OPERAND_* */
  int32_t io_flags;         /* SVX_IO_* bits */
} svx_gemm_desc;

/*
 * The MLP of one Swin block (timm Mlp: fc1 -> nn.GELU() -> fc2; swin_transformer.py:71-94 runs it through
 * timm's SwinTransformerBlock) together with the block's second residual, as ONE kernel:
 *   out[r, :] = residual[r, :] + W2 * round_tf32(gelu(W1 * x[r, :] + b1)) + b2
 * The 4C-wide hidden activation never leaves the SM: per 128-row tile the fc1 accumulator (TMEM) goes through the
 * GELU epilogue into shared memory in the 128B-swizzled operand layout and is contracted with W2 straight away.
 * x: [M, C] with row pitch ldx (the TF32-rounded norm2 output); W1: [hidden, C], W2: [C, hidden], both K-contiguous
 * and pre-rounded to TF32; b1[hidden], b2[C]; residual / out: [M, C] with row pitch ldo (may alias).
 * C = 96 or 192, hidden = 4*C; pointers 16-byte aligned, pitches multiples of 4.
 */
typedef struct svx_mlp_desc {
  const float* x; int64_t ldx;
  const float* W1; const float* b1;
  const float* W2; const float* b2;
  const float* residual; float* out; int64_t ldo;
  int32_t M, C, hidden;
  int32_t round_tf32;       /* round the stored result to TF32 */
  /* optional fused pre-LayerNorm (the block's norm2; C = 96 only): when ln_gamma is set, x holds the UN-normalised rows
   * (normally x == residual) and fc1 reads round_tf32(LayerNorm(x) * ln_gamma + ln_beta), normalised inside the kernel */
  const float* ln_gamma; const float* ln_beta;
  float ln_eps; int32_t reserved0;
} svx_mlp_desc;

/* Explicit im2col for tiny channel counts (ResNet stem 7x7 s2 on 3 channels, Swin patch-embed
 * 4x4 s4, refiner layer1 4x4x4 on 1 channel).  Input addressed with arbitrary element strides
 * so NCHW user tensors are read in place.  Row r=(n,od,oh,ow); k=((kd*KH+kh)*KW+kw)*C+c. */
typedef struct svx_im2col_desc {
  const float* in; float* out;
  int32_t N, C, D, H, W;
  int64_t s_n, s_c, s_d, s_h, s_w;
  int32_t KD, KH, KW, stride, pad_d, pad_h, pad_w;
  int32_t OD, OH, OW, Kpad;
  int32_t round_tf32;
} svx_im2col_desc;

/* Channels-last pooling (resnet maxpool, encoder.py:123 avg_pool2d, decoder.py:59-67 adaptive
 * pool + depth replicate, refiner.py MaxPool3d).  Output (od,oh,ow) reduces the window starting
 * at (od*sd - pd, ...) of extent (kd,kh,kw); out-of-range taps are skipped (max) / not counted. */
typedef struct svx_pool_desc {
  const float* in; float* out;
  int32_t N, C, D, H, W, in_Cs, out_Cs;
  int32_t KD, KH, KW, SD, SH, SW, PD, PH, PW, OD, OH, OW;
  int32_t mode, round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_pool_desc;

/* Row LayerNorm over C channels (timm norm1/norm2/patch_embed.norm/downsample.norm), eps 1e-5.
 * merge=1 gathers the PatchMerging 2x2 neighbourhood: in is [N,H,W,C/4], row (n,y,x) over
 * (H/2,W/2) concatenates (2y,2x),(2y+1,2x),(2y,2x+1),(2y+1,2x+1). */
typedef struct svx_lnrows_desc {
  const float* in; float* out; const float* gamma; const float* beta;
  int32_t rows, C; int32_t merge, H, W; float eps; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_lnrows_desc;

/* nn.LayerNorm([C,H,W]) of the Swin wrapper (swin_transformer.py:64-67,84-86): statistics over
 * all L=C*H*W values of one sample, affine laid out like the data ([H,W,C]). */
typedef struct svx_lnsample_desc {
  const float* in; float* out; const float* gamma; const float* beta;
  int32_t N, L; float eps; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_lnsample_desc;

/* W-MSA / SW-MSA (timm WindowAttention + roll/partition/reverse).  qkv is [N*H*W, 3C] in token
 * order, columns (which, head, d); out is [N*H*W, C].  bias is the expanded relative-position
 * bias [heads, 49, 49].  The cyclic shift and the -100 region mask are applied by indexing.
 * tcgen05 kernel (svx_winattn.cu): two windows per 128-row tile, score / output accumulators in TMEM. */
typedef struct svx_winattn_desc {
  const float* qkv; float* out; const float* bias;
  int32_t N, H, W, C, heads, shift; float scale; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
  int32_t reserved0;
  int32_t* range_flag;      /* fp32 storage: the P V product uses fp16 operands (exact for the TF32-rounded values this path
                               stores while |v| <= 65504); device int32 OR-ed with 1 if a V value saturated, or NULL */
} svx_winattn_desc;

/* depthwise k=s conv without padding (cross_view_attention.py:26-34), channels-last */
typedef struct svx_dwconv_desc {
  const float* in; float* out; const float* w /* [k*k, C] */; const float* bias;
  int32_t N, H, W, C, k, OH, OW; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_dwconv_desc;

/* attention over the view axis (cross_view_attention.py:81-103). qkv: [B*V, P, 3*R] channels-last
 * (P = h*w positions, R = reduced channels); out: [B*V, P, R]. */
typedef struct svx_viewattn_desc {
  const float* qkv; float* out;
  int32_t B, V, P, R, heads; float scale; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_viewattn_desc;

/* out = bilinear_resize(in, align_corners=False) + skip  (cross_view_attention.py:110-120) */
typedef struct svx_bilinear_desc {
  const float* in; const float* skip; float* out;
  int32_t N, IH, IW, OH, OW, C; int32_t round_tf32;
  int32_t dtype;            /* SVX_DT_* bits */
} svx_bilinear_desc;

/* Conv3d(Cin <= 12 -> 1, k3, s1, p1) + folded BatchNorm + LeakyReLU on the CUDA cores in fp32 (merger.py:50-54, layer6:
 * 243 MACs per voxel -- a tensor-core tile would spend a full 128 x N x 8 instruction slot per 8 of them).
 * in: zero-bordered channels-last volume [N, D+2, H+2, W+2, Cs], channels [c0, c0 + Cin); w: [27, 12] (tap-major
 * kd, kh, kw; channel-minor, zero padded), bias[1]; out: planar [N, D*H*W].  H % 16 == 0, W == 32, Cs % 4 == 0. */
typedef struct svx_conv3to1_desc {
  const float* in; const float* w; const float* bias; float* out;
  int32_t N, D, H, W, Cs, c0, Cin;
  float slope;
} svx_conv3to1_desc;

/* per-voxel softmax over views + weighted sum (merger.py:98-104).
 * weights, coarse: [B, V, P]; out: [B, P].  weights == NULL: the plain mean over the views, the reference's
 * torch.mean(generated_volume, dim=1) when the merger is off or not yet enabled (core/test.py:123-126). */
typedef struct svx_mergefuse_desc {
  const float* weights; const float* coarse; float* out;
  int32_t B, V, P;
} svx_mergefuse_desc;

/* sigmoid -> threshold -> I/U/TP/FP/FN per object (core/test.py:141-164).
 * logits, gt: [B, P] (gt holds {0,1}); prob_thresholds: device float[T] (cfg.TEST.VOXEL_THRESH);
 * counts: int32[B, T, 5] = {I,U,TP,FP,FN}, zeroed by the launch itself (cudaMemsetAsync).
 * bce_q20 (optional): int64[B], per object the sum over voxels of BCEWithLogits(logit, gt) =
 * max(x,0) - x*gt + log1p(exp(-|x|)) (the EDLoss / RLoss of core/test.py:133-139 before the mean and the *10), in
 * fixed point with 20 fractional bits so that the atomic accumulation is order-independent; zeroed by the launch. */
typedef struct svx_metrics_desc {
  const float* logits; const float* gt; const float* prob_thresholds; int32_t* counts;
  int32_t B, P, T;
  int32_t reserved0;
  int64_t* bce_q20;
} svx_metrics_desc;

/* layout change [N, C, P] (planar) <-> [N, P, Cs] (channels-last, first C of Cs channels).
 * to_channels_last with row_w > 0 (Cs == 4 only): the P pixels of a sample are rows of row_w pixels and pixel (y, x)
 * is stored at pixel y*row_pitch + row_x0 + x of the sample's (P/row_w)*row_pitch-pixel block -- zero columns left
 * and right of every image row (never written), the layout the ResNet stem reads as 8-channel pixel pairs. */
typedef struct svx_transpose_desc {
  const float* in; float* out;
  int32_t N, C, P, Cs; int32_t to_channels_last; int32_t round_tf32;
  int32_t row_w, row_pitch, row_x0;
  int32_t dtype;            /* SVX_DT_* bits (the planar side is always fp32: only SVX_DT_OUT_BF16 with to_channels_last,
                               SVX_DT_IN_BF16 without) */
} svx_transpose_desc;

/* torch.nn.functional.interpolate(x, size=(OH, OW), mode="bilinear", align_corners=False) on planar fp32 images
 * (models/swin_transformer.py:74-75 resizes inputs that are not img_size x img_size): in [NC, IH, IW] -> out [NC, OH, OW],
 * source coordinate max((o + 0.5) * IH / OH - 0.5, 0), the neighbour index clamped at the border. */
typedef struct svx_resize_desc {
  const float* in; float* out;
  int32_t NC, IH, IW, OH, OW;
  int32_t reserved0;
} svx_resize_desc;

/* binvox run-length decode (utils/binvox_rw.py:119-153 read_as_3d_array; utils/data_loaders.py:84-87): the payload
 * after the text header is (value, count) byte pairs; np.repeat(values, counts).astype(bool).reshape(dims) gives the
 * volume in file order x, z, y (y fastest); fix_coords transposes it to x, y, z.  B objects per launch: their
 * payloads are concatenated in `payload` (each starting at an even offset), object b owns bytes
 * [offsets[b], offsets[b+1]).  out: fp32 {0,1} [B, d0, d2, d1] (fix_coords) or [B, d0, d1, d2]; status[b] = number of
 * voxels the stream expands to (the caller compares it with d0*d1*d2, the reference's reshape would raise). */
typedef struct svx_binvox_decode_desc {
  const uint8_t* payload; const int64_t* offsets; float* out; int32_t* status;
  int32_t B, d0, d1, d2, fix_coords;
} svx_binvox_decode_desc;

/* binvox run-length encode (utils/binvox_rw.py:239-300 write): volume fp32 [B, d0, d1, d2], voxel set iff
 * value >= threshold; axis_xyz != 0: the volume is in x, y, z order and is written transposed (file order x, z, y).
 * Runs are cut at 255 exactly like the reference's state machine, including its zero-length pair after a run whose
 * length is a multiple of 255.  payload: [B, 2*d0*d1*d2] bytes (worst case), nbytes[b] = bytes written for b. */
typedef struct svx_binvox_encode_desc {
  const float* volume; float threshold; uint8_t* payload; int32_t* nbytes;
  int32_t B, d0, d1, d2, axis_xyz;
} svx_binvox_encode_desc;

/* Evaluation-time image pipeline (core/test.py:50-55): CenterCrop (utils/data_transforms.py:76-167, no bounding box) ->
 * RandomBackground with a fixed colour (:415-452) -> Normalize (:57-62) -> ToTensor (:42-49), over renderings read as
 * uint8 (utils/data_loaders.py:70-76 divides by 255).  in: uint8 [N, H, W, C] (C = 4: BGRA, C = 3: no alpha);
 * [y0, y1) x [x0, x1) is the crop window (the host applies the reference's rule: centred crop_size window when the
 * image is larger than it, else the whole image); the window is resized to OH x OW with cv2's INTER_LINEAR convention
 * on the value/255 floats (all channels incl. alpha); where the resized alpha is exactly 0 the pixel becomes the
 * background; out: fp32 [N, 3, OH, OW] = (x - mean) / std; bg_norm is the already normalised background colour. */
typedef struct svx_preprocess_desc {
  const uint8_t* in; float* out;
  int32_t N, H, W, C, OH, OW;
  int32_t y0, y1, x0, x1;
  float mean[3], std[3], bg_norm[3];
  int32_t reserved0;
  /* optional per-image overrides (device arrays, or NULL):
   * windows [N][4] = {y0, y1, x0, x1}: the bounding-box crops of utils/data_transforms.py:93-131 -- the window may leave
   *   the image; rows / columns outside are the nearest edge pixel (the reference's np.pad(mode='edge')), i.e. source
   *   coordinates clamp.  y1 > y0, x1 > x0 and the window must intersect the image (the reference raises otherwise).
   * bg_norm_n [N][3]: per-image normalised background colour (RandomBackground with a proper colour range draws one
   *   colour per sample, utils/data_transforms.py:433-435; the host draws, the kernel applies). */
  const int32_t* windows;
  const float* bg_norm_n;
} svx_preprocess_desc;

/* ---- library ------------------------------------------------------------------------ */
int svx_abi_version(void);
const char* svx_last_error(void);
/* sizeof() of every descriptor, so the Python mirror can verify its struct layout */
int svx_desc_sizes(int32_t* sizes, int n);
int svx_device_info(int device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- immediate launches (one op on `stream`) ----------------------------------------- */
int svx_gemm(const svx_gemm_desc*, void* stream);
int svx_mlp(const svx_mlp_desc*, void* stream);
int svx_im2col(const svx_im2col_desc*, void* stream);
int svx_pool(const svx_pool_desc*, void* stream);
int svx_layernorm_rows(const svx_lnrows_desc*, void* stream);
int svx_layernorm_sample(const svx_lnsample_desc*, void* stream);
int svx_window_attention(const svx_winattn_desc*, void* stream);
int svx_dwconv(const svx_dwconv_desc*, void* stream);
int svx_view_attention(const svx_viewattn_desc*, void* stream);
int svx_bilinear_add(const svx_bilinear_desc*, void* stream);
int svx_merger_fuse(const svx_mergefuse_desc*, void* stream);
int svx_conv3to1(const svx_conv3to1_desc*, void* stream);
int svx_voxel_metrics(const svx_metrics_desc*, void* stream);
int svx_transpose(const svx_transpose_desc*, void* stream);
int svx_resize_bilinear(const svx_resize_desc*, void* stream);
int svx_binvox_decode(const svx_binvox_decode_desc*, void* stream);
int svx_binvox_encode(const svx_binvox_encode_desc*, void* stream);
int svx_preprocess(const svx_preprocess_desc*, void* stream);

/* ---- plans: a recorded op list replayed per forward (one per module instance/shape) --- */
typedef struct svx_plan svx_plan;
svx_plan* svx_plan_create(void);
void svx_plan_destroy(svx_plan*);
int svx_plan_num_ops(const svx_plan*);
int svx_plan_add_gemm(svx_plan*, const svx_gemm_desc*);
int svx_plan_add_mlp(svx_plan*, const svx_mlp_desc*);
int svx_plan_add_im2col(svx_plan*, const svx_im2col_desc*);
int svx_plan_add_pool(svx_plan*, const svx_pool_desc*);
int svx_plan_add_layernorm_rows(svx_plan*, const svx_lnrows_desc*);
int svx_plan_add_layernorm_sample(svx_plan*, const svx_lnsample_desc*);
int svx_plan_add_window_attention(svx_plan*, const svx_winattn_desc*);
int svx_plan_add_dwconv(svx_plan*, const svx_dwconv_desc*);
int svx_plan_add_view_attention(svx_plan*, const svx_viewattn_desc*);
int svx_plan_add_bilinear_add(svx_plan*, const svx_bilinear_desc*);
int svx_plan_add_merger_fuse(svx_plan*, const svx_mergefuse_desc*);
int svx_plan_add_conv3to1(svx_plan*, const svx_conv3to1_desc*);
int svx_plan_add_voxel_metrics(svx_plan*, const svx_metrics_desc*);
int svx_plan_add_transpose(svx_plan*, const svx_transpose_desc*);
int svx_plan_add_resize_bilinear(svx_plan*, const svx_resize_desc*);
/* Concurrency hints.  Ops are recorded into the current lane (default 0 = the caller's stream).  Ops of a side lane
 * k in [1, 8] run in order on the plan's side stream k, forked from the caller's stream where the lane's first op
 * since the last join sits in the op list; svx_plan_add_join makes the caller's stream wait for every side lane (the
 * end of the plan joins implicitly).  The caller guarantees that ops of different lanes between two joins are
 * independent.  Used for the eight output-parity classes of a stride-2 ConvTranspose3d. */
int svx_plan_set_lane(svx_plan*, int lane);
int svx_plan_add_join(svx_plan*);
/* run every op in order on `stream`; use_graph != 0 replays a CUDA graph captured on first use */
int svx_plan_run(svx_plan*, void* stream, int use_graph);
/* run ops [first, last) only (profiling / bisecting) */
int svx_plan_run_range(svx_plan*, int first, int last, void* stream);
/* device time of each op (CUDA events on `stream`, mean of `iters` launches); ms has num_ops entries */
int svx_plan_time_ops(svx_plan*, void* stream, int iters, float* ms);
/* number of kernel launches one svx_plan_run issues */
int svx_plan_num_launches(const svx_plan*);

#ifdef __cplusplus
}
#endif
#endif /* SWINVOX_B200_H */

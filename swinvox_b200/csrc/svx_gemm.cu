// svx_gemm.cu -- the contraction engine of the SwinVox forward path on sm_100a.
//
// One persistent, warp-specialised kernel serves every Linear / Conv2d / Conv3d / ConvTranspose3d of the
// reference (encoder.py:22-111, timm Swin linears, cross_view_attention.py:38-53, decoder.py:24-46,
// merger.py:20-54, refiner.py:21-70):
//
//   warp 0-3  epilogue   TMEM -> registers (tcgen05.ld) -> smem transpose -> bias / residual / activation ->
//                        coalesced 128-byte row segments to global; overlaps the next tile's main loop through
//                        a double-buffered TMEM accumulator
//   warp 4    TMA        weights, and the A operand when it is a plain matrix or a pre-padded stride-1
//                        convolution ("flat" mode: one shifted 2-D box per filter tap) -> 128B-swizzled smem
//   warp 5    MMA        one elected thread issues tcgen05.mma kind::tf32, accumulating in TMEM
//   warp 6-9  gather     strided / unpadded convolutions: cp.async 16-byte chunks of channels-last pixels written
//                        with the 128B swizzle TMA would produce, zero-filled at the borders
//
// Tile: 128 (rows = output pixels) x BN (output channels) x 32 (fp32 k-chunk = one swizzle row).  Grid = one CTA
// per SM; tiles are dealt round-robin (n fastest, so CTAs that share an A tile run together and hit L2).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "svx_internal.h"
#include "svx_ptx.cuh"
#include "svx_act.cuh"

namespace svx {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int UMMA_K = 8;
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int kThreads = 320;
constexpr int kGatherLag = 2;
constexpr int kMaxTaps = 64;
constexpr int kSlab = 16;                        // accumulator columns per epilogue pass

constexpr int kBlockBytes = 32 * kSlab * 4;      // one staged 32-row x 16-column fp32 block

struct GemmParams {
  int M, N, K, Npad, nk, tiles_n, tiles_m, a_mode;
  const float* A;
  int in_D, in_H, in_W, in_Cs, in_c0, Cin;
  int out_D, out_H, out_W;
  int valid_D, valid_H, valid_W;
  int sd, sh, sw;
  int chunks_per_tap;
  int cin_live;             // slab mode: leading channels of the 32-channel box that carry non-zero weights
  const int4* taps;
  const float* bias;
  const float* residual;
  float* out;
  long long o_base, o_sn, o_sd, o_sh, o_sw;
  int act;
  float act_param;
  int res_after_act;
  float out_scale;
  int round_tf32, epi_mode;
  const float* epi_aux;
  float* out2;
  long long o2_base, o2_sn, o2_sd, o2_sh, o2_sw;
  int vec_ok;
  int epi_tma;              // plain row-major output: epilogue stores through TMA (map_c)
  int nk_a;                 // k-chunks read from map_a; the remaining nk - nk_a come from the residual (map_r)
  long long ldc;            // epi_tma: row pitch of out / residual (elements); out and residual already include o_base
  int cls_cout;             // SVX_EPI_CONVT8: channels per output-parity class
  long long c_sd, c_sh, c_sw, c2_sd, c2_sh, c2_sw;   // class-bit offsets in out / residual and in out2
  int i2c_lo_d, i2c_lo_h, i2c_lo_w;   // im2col mode: base-pixel coordinate of output 0 on each axis (= smallest tap)
  int i2c_narrow, ntaps;              // 4-channel pixels (image stems): eight 16-byte taps per k-chunk, unswizzled A tile
  int io16;                 // bf16 kernels: out and residual are bf16 in memory (otherwise fp32)
  const float* Wg;          // slab mode, fp16 operands: the fp32 weight matrix in global memory (converted once per CTA)
  int* range_flag;          // slab mode, fp16 operands: OR-ed with 1 when an activation saturated in the conversion
  float acc_scale;          // slab mode: accumulator scale applied before the bias (inverse of the host's weight scale)
  int flat_off[kMaxTaps];  // flat mode: row offset of each tap; im2col mode: tap offsets packed w | h << 8 | d << 16
};

template <int BN, int PARTS = 2>
struct Cfg {
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBBytes;
  static constexpr int kCtasPerSm = BN <= 96 ? 2 : 1;   // two resident CTAs double the epilogue / gather warps
  // per epilogue warp: kEpiBufs staged blocks for the TMA-store epilogue (the mapped-output epilogue uses the first
  // block for its transpose and the 256 bytes after it for the row offsets); 512-byte multiples keep the 64B swizzle
  static constexpr int kEpiBufs = BN == 96 ? 1 : 2;
  static constexpr int kWarpStage = BN == 96 ? kBlockBytes + 512 : 2 * kBlockBytes;
  // epilogue warps: warps 0-3 (part 0) and warps 6.. (parts 1..kParts-1; warps 6-9 gather instead in SVX_A_GATHER mode).
  // PARTS = 4 (18 warps, 96 registers each) is the variant for erf-GELU epilogues on the one-CTA-per-SM tiles, which are
  // bound by the epilogue warps' FP32 / MUFU issue: stage-2 fc1 0.71 -> 0.63 ms.  Every other epilogue is faster with
  // PARTS = 2 (138 registers, one more pipeline stage): measured per op in profiles/r1_gemm_epilogue_parts_v23.txt.
  static constexpr int kParts = PARTS;
  static_assert(PARTS == 2 || kCtasPerSm == 1, "two resident CTAs leave no registers for extra epilogue warps");
  static constexpr int kEpiWarps = 4 * kParts;
  static constexpr int kThreadsT = 32 * (6 + 4 * (kParts - 1));
  static constexpr int kAuxBytes = kEpiWarps * kWarpStage + 256;          // + barriers
  static constexpr int kBudget = (kCtasPerSm == 2 ? 115712 : 232448) - 1024 - kAuxBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr uint32_t kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static_assert(kStages >= kGatherLag + 1, "the gather pipeline needs more stages than its lag");
  // stages + 1024 alignment slack + barriers (256) + staging + row offsets
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kAuxBytes;
};

template <int ACT>
__device__ __forceinline__ float act_t(float x, float slope) {
  if constexpr (ACT == SVX_ACT_RELU) return fmaxf(x, 0.f);
  if constexpr (ACT == SVX_ACT_LEAKY) return x > 0.f ? x : x * slope;
  if constexpr (ACT == SVX_ACT_GELU) return gelu_erf(x);
  return x;
}

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  switch (act) {
    case SVX_ACT_RELU: return act_t<SVX_ACT_RELU>(x, slope);
    case SVX_ACT_LEAKY: return act_t<SVX_ACT_LEAKY>(x, slope);
    case SVX_ACT_GELU: return act_t<SVX_ACT_GELU>(x, slope);
    default: return x;
  }
}


__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// row index -> (valid, output offset); rows decode as (n, od, oh, ow) over the row-space extents
__device__ __forceinline__ bool decode_row(const GemmParams& p, int r, long long& off, long long& off2) {
  if (r >= p.M) return false;
  const int ow = r % p.out_W;
  int t = r / p.out_W;
  const int oh = t % p.out_H;
  t /= p.out_H;
  const int od = t % p.out_D;
  const int n = t / p.out_D;
  if (p.valid_W > 0 && (ow >= p.valid_W || oh >= p.valid_H || od >= p.valid_D)) return false;
  off = p.o_base + n * p.o_sn + od * p.o_sd + oh * p.o_sh + ow * p.o_sw;
  off2 = p.o2_base + n * p.o2_sn + od * p.o2_sd + oh * p.o2_sh + ow * p.o2_sw;
  return true;
}


// One epilogue pass of a warp over a staged 32 x SLAB accumulator block: every warp-wide access covers whole
// 16*CP-byte row segments (coalesced), rows are processed G at a time so their loads overlap.
template <int SLAB, int ACT, bool IO16 = false>
__device__ __forceinline__ void store_slab(const GemmParams& p, const float* staging, const long long* soff,
                                           int lane, int jb) {
  constexpr int CP = SLAB / 4;
  constexpr int RP = 32 / CP;
  constexpr int G = CP < 4 ? CP : 4;
  const int ch = lane % CP;
  const int col = jb + ch * 4;
  const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool has_res = p.residual != nullptr;
  const bool pre = has_res && !p.res_after_act;
  const float slope = p.act_param, scale = p.out_scale;
  const bool rnd = p.round_tf32 != 0;
#pragma unroll
  for (int i0 = 0; i0 < CP; i0 += G) {
    long long ro[G];
    bool ok[G];
    float4 a4[G], r4[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int row = (i0 + i) * RP + lane / CP;
      ro[i] = soff[row];
      ok[i] = ro[i] >= 0 && col < p.N;
      a4[i] = *reinterpret_cast<const float4*>(staging + row * SLAB + ((ch ^ (row & (CP - 1))) << 2));
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {  // plain loads: out may alias the residual
      if constexpr (IO16)
        r4[i] = (has_res && ok[i]) ? ld4(reinterpret_cast<const bf16_t*>(p.residual) + ro[i] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      else
        r4[i] = (has_res && ok[i]) ? *reinterpret_cast<const float4*>(p.residual + ro[i] + col)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      float x[4] = {a4[i].x + b4.x, a4[i].y + b4.y, a4[i].z + b4.z, a4[i].w + b4.w};
      const float rv[4] = {r4[i].x, r4[i].y, r4[i].z, r4[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t = pre ? x[q] + rv[q] : x[q];
        t = act_t<ACT>(t, slope);
        if (!pre) t += rv[q];
        t *= scale;
        x[q] = rnd ? round_tf32(t) : t;
      }
      if (ok[i]) {
        if constexpr (IO16) st4(reinterpret_cast<bf16_t*>(p.out) + ro[i] + col, make_float4(x[0], x[1], x[2], x[3]));
        else *reinterpret_cast<float4*>(p.out + ro[i] + col) = make_float4(x[0], x[1], x[2], x[3]);
      }
    }
  }
}

// BF = false: fp32 storage read as TF32 (kind::tf32, 32 elements per 128-byte k-chunk).  BF = true: bf16 A / W
// (kind::f16 with bf16 operand formats, 64 elements per k-chunk, K = 16 per MMA); the byte geometry of the pipeline --
// 128-byte swizzled rows, four 32-byte MMA steps per chunk -- is identical, so one body serves both.
template <int BN, int PARTS, bool BF>
__device__ __forceinline__ void gemm_body(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_c,
                                          const CUtensorMap& map_r, const GemmParams& p) {
  using C = Cfg<BN, PARTS>;
  constexpr int S = C::kStages;
  constexpr int BKE = BF ? 64 : 32;        // elements per k-chunk
  constexpr int ESZ = BF ? 2 : 4;          // bytes per operand element
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // after the pipeline stages: epilogue staging (1024-byte aligned: TMA-store source blocks), row offsets, barriers
  uint8_t* aux_gen = smem_gen + S * C::kStageBytes;
  constexpr int kBarOff = C::kEpiWarps * C::kWarpStage;
  const uint32_t bar_base = smem_base + S * C::kStageBytes + kBarOff;
  // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], tmem slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(aux_gen + kBarOff + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = p.nk;
  const int a_mode = p.a_mode;
  const int num_tiles = p.tiles_m * p.tiles_n;

  // ---- one-time setup ---------------------------------------------------------------------
  if (warp == 4 && lane == 0) {
    if (a_mode != SVX_A_GATHER) tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.epi_tma) tma_prefetch_desc(&map_c);
    if (p.nk_a < p.nk) tma_prefetch_desc(&map_r);
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), a_mode == SVX_A_GATHER ? 5u : 1u);
        mbar_init(empty_bar(s), 1u);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tmem_full_bar(a), 1u);
        mbar_init(tmem_empty_bar(a), a_mode == SVX_A_GATHER ? 4u : 4u * C::kParts);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the previous kernel in
  // the stream; its results are only touched from here on.  The next kernel may start its own preamble right away.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 4) {
    // ---- TMA producer (one elected thread: elect.sync lets ptxas issue the uniform-datapath TMA / MMA
    // instructions directly instead of wrapping each in a divergence loop) ----------------------------
    if (elect_one()) {
      uint32_t g = 0;
      int i2c_w = 0, i2c_h = 0, i2c_d = 0, i2c_n = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t a_dst = smem_base + s * C::kStageBytes;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          if (a_mode == SVX_A_PLAIN) {
            mbar_arrive_expect_tx(full_bar(s), A_STAGE_BYTES + C::kBBytes);
            if (kc < p.nk_a) tma_load_2d(a_dst, &map_a, full_bar(s), kc * BKE, m0);
            else tma_load_2d(a_dst, &map_r, full_bar(s), n0 + (kc - p.nk_a) * BKE, m0);   // residual x identity
          } else if (a_mode == SVX_A_FLAT) {
            mbar_arrive_expect_tx(full_bar(s), A_STAGE_BYTES + C::kBBytes);
            const int tap = kc / p.chunks_per_tap;
            const int cc = kc - tap * p.chunks_per_tap;
            tma_load_2d(a_dst, &map_a, full_bar(s), p.in_c0 + cc * BKE, m0 + p.flat_off[tap]);
          } else if (a_mode == SVX_A_IM2COL) {
            mbar_arrive_expect_tx(full_bar(s), A_STAGE_BYTES + C::kBBytes);
            if (kc == 0) {   // first output pixel of the tile -> base pixel of the hardware traversal
              i2c_w = m0 % p.out_W;
              int t = m0 / p.out_W;
              i2c_h = t % p.out_H;
              t /= p.out_H;
              i2c_d = t % p.out_D;
              i2c_n = t / p.out_D;
              i2c_w = i2c_w * p.sw + p.i2c_lo_w; i2c_h = i2c_h * p.sh + p.i2c_lo_h; i2c_d = i2c_d * p.sd + p.i2c_lo_d;
            }
            if (p.i2c_narrow == 2) {
              // 8-channel pixel pairs: a k-chunk is four taps, each a 128-pixel x 32-byte box (32B swizzle) = the A
              // operand of ONE MMA K step
#pragma unroll 1
              for (int j = 0; j < 4; ++j) {
                const int o = p.flat_off[kc * 4 + j];
                tma_load_im2col_5d(a_dst + j * (BM * 32), &map_a, full_bar(s), p.in_c0, i2c_w, i2c_h, i2c_d, i2c_n,
                                   (uint16_t)(o & 255), (uint16_t)((o >> 8) & 255), (uint16_t)((o >> 16) & 255));
              }
            } else if (p.i2c_narrow) {
              // 4-channel pixels: a k-chunk is eight taps, each its own 128-pixel x 16-byte box = one column of core
              // matrices of the unswizzled K-major tile (taps past the filter re-read tap 0 against zero weights)
#pragma unroll 1
              for (int j = 0; j < 8; ++j) {
                const int tap = kc * 8 + j;
                const int o = p.flat_off[tap < p.ntaps ? tap : 0];
                tma_load_im2col_5d(a_dst + j * (BM * 16), &map_a, full_bar(s), p.in_c0, i2c_w, i2c_h, i2c_d, i2c_n,
                                   (uint16_t)(o & 255), (uint16_t)((o >> 8) & 255), (uint16_t)((o >> 16) & 255));
              }
            } else {
              const int tap = kc / p.chunks_per_tap;
              const int cc = kc - tap * p.chunks_per_tap;
              const int o = p.flat_off[tap];
              tma_load_im2col_5d(a_dst, &map_a, full_bar(s), p.in_c0 + cc * BKE, i2c_w, i2c_h, i2c_d, i2c_n,
                                 (uint16_t)(o & 255), (uint16_t)((o >> 8) & 255), (uint16_t)((o >> 16) & 255));
            }
          } else {
            mbar_arrive_expect_tx(full_bar(s), C::kBBytes);
          }
          tma_load_2d(b_dst, &map_b, full_bar(s), kc * BKE, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---- MMA issuer -----------------------------------------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = BF ? umma_idesc_bf16(BM, BN) : umma_idesc_tf32(BM, BN);
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1u;
        mbar_wait(tmem_empty_bar(as), ((it >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + as * BN;
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * C::kStageBytes;
          const int narrow = p.i2c_narrow;
          const uint64_t da = narrow == 2 ? umma_desc_sw32(a_addr) : narrow ? umma_desc_nosw(a_addr, BM * 16, 128) : umma_desc_sw128(a_addr);
          const uint64_t db = umma_desc_sw128(a_addr + A_STAGE_BYTES);
          // per MMA K step (32 bytes): +32 B inside the 128B atom / two 16-byte boxes / one 32-byte box
          const uint32_t a_step = narrow ? (BM * 32u) >> 4 : 2u;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 8 fp32 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
            if constexpr (BF) umma_f16(acc, da + a_step * k, db + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
            else umma_tf32(acc, da + a_step * k, db + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar(as));
      }
    }
    __syncwarp();
  } else if (warp >= 6 && warp < 10 && a_mode == SVX_A_GATHER) {
    // ---- A gather producers (implicit im2col for strided / unpadded convolutions) ----------------------
    {
      const int gw = warp - 6;
      const int j = lane & 7;      // 16-byte chunk inside the 128-byte k-row
      const int rsub = lane >> 3;  // row inside a group of 4
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        long long base[8];
        int crd[8];  // zd | zh<<10 | zw<<20 | valid<<30
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = m0 + gw * 32 + i * 4 + rsub;
          if (r < p.M) {
            const int ow = r % p.out_W;
            int t = r / p.out_W;
            const int oh = t % p.out_H;
            t /= p.out_H;
            const int od = t % p.out_D;
            const int n = t / p.out_D;
            const int zd = od * p.sd, zh = oh * p.sh, zw = ow * p.sw;
            base[i] = ((((long long)n * p.in_D + zd) * p.in_H + zh) * p.in_W + zw) * p.in_Cs * ESZ;   // bytes
            crd[i] = zd | (zh << 10) | (zw << 20) | (1 << 30);
          } else {
            base[i] = 0;
            crd[i] = 0;
          }
        }
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const int k4 = kc * BKE + j * (16 / ESZ);   // first element of this lane's 16-byte chunk
          const bool kv = k4 < p.K;
          int dd = 0, dh = 0, dw = 0, delta = 0;
          if (kv) {
            const int tap = k4 / p.Cin;
            const int c = k4 - tap * p.Cin;
            const int4 t = __ldg(p.taps + tap);
            dd = t.x; dh = t.y; dw = t.z;
            delta = (((dd * p.in_H + dh) * p.in_W + dw) * p.in_Cs + p.in_c0 + c) * ESZ;   // bytes
          }
          const uint32_t a_dst = smem_base + s * C::kStageBytes;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = gw * 32 + i * 4 + rsub;
            const int id = (crd[i] & 1023) + dd;
            const int ih = ((crd[i] >> 10) & 1023) + dh;
            const int iw = ((crd[i] >> 20) & 1023) + dw;
            const bool ok = kv && (crd[i] >> 30) && (unsigned)id < (unsigned)p.in_D &&
                            (unsigned)ih < (unsigned)p.in_H && (unsigned)iw < (unsigned)p.in_W;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.A) + (ok ? base[i] + delta : 0);
            const uint32_t dst = a_dst + row * 128 + ((j ^ (row & 7)) << 4);
            cp_async16_zfill(dst, src, ok ? 16u : 0u);
          }
          cp_async_commit();
          if (g >= kGatherLag) {   // the pipeline runs across tile boundaries
            cp_async_wait<kGatherLag>();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar((g - kGatherLag) % S));
          }
        }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (uint32_t q = (g > kGatherLag ? g - kGatherLag : 0); q < g; ++q) mbar_arrive(full_bar(q % S));
      }
      __syncwarp();
    }
  } else if (warp < 4 || a_mode != SVX_A_GATHER) {
    // ---- epilogue: warps 0-3, helped by warps 6.. whenever the A operand is not gathered.  A warp may only touch the
    // TMEM lane quarter (warp % 4); the warps of a quarter take the 16-column slabs round-robin. ----------------
    const int quarter = warp & 3;
    const int part = warp < 4 ? 0 : 1 + ((warp - 6) >> 2);
    const int nparts = a_mode == SVX_A_GATHER ? 1 : C::kParts;
    const int ew = part * 4 + quarter;
    float* staging = reinterpret_cast<float*>(aux_gen + ew * C::kWarpStage);
    long long* soff = reinterpret_cast<long long*>(aux_gen + ew * C::kWarpStage + kBlockBytes);
    constexpr int SLAB = kSlab;
    constexpr int CP = SLAB / 4;     // float4 chunks per staged row
    const bool convt = p.epi_mode == SVX_EPI_CONVT8;
    const bool rowwise = (p.epi_mode == SVX_EPI_DEC_TAIL) || convt || !p.vec_ok;
    uint32_t it = 0, buf = 0;   // buf: which of this warp's two staging blocks the next TMA store uses
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
      const uint32_t as = it & 1u;
      if (p.epi_tma) {
        // ---- plain row-major output: TMEM -> registers -> bias / residual / activation -> 64B-swizzled smem block
        // (32 rows x 16 columns per warp) -> one TMA store per block.  No per-element address arithmetic, the
        // store is coalesced by the TMA unit and clipped at the matrix edges. ----
        const uint32_t lane_addr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        const int row = m0 + quarter * 32 + lane;
        const bool row_ok = row < p.M;
        const float* res_row = p.residual + static_cast<long long>(row) * p.ldc;
        const bf16_t* res_row16 = reinterpret_cast<const bf16_t*>(p.residual) + static_cast<long long>(row) * p.ldc;
        const bool io16 = BF && p.io16;
        const bool has_res = p.residual != nullptr;
        const bool pre = has_res && !p.res_after_act, post = has_res && p.res_after_act;
        char* stg = reinterpret_cast<char*>(staging);                       // 2 x 2 KB, 1024-byte aligned
        const uint32_t stg_u32 = smem_u32(stg);
        const int sw = (lane >> 1) & 3;                                        // 64B swizzle phase of this row
        if constexpr (BF) {
          if (io16) {
            // bf16 output: 32 columns per pass, so a staged row is 64 bytes (the fp32 path's geometry: 64B swizzle, one
            // 2 KB block, one TMA store per 32 rows x 32 columns) -- 16-column passes halved the bytes per bulk store
            // and left the HBM-bound layers store-issue-bound.
            mbar_wait(tmem_full_bar(as), (it >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = part * 32; c0 < BN; c0 += nparts * 32) {
              const int jb = n0 + c0;
              if (jb >= p.N) break;  // warp-uniform
              // the previous store out of this buffer must have been read by the TMA unit
              if (lane == 0) tma_store_wait_read<C::kEpiBufs - 1>();
              __syncwarp();
              char* my_row = stg + buf * 2048 + lane * 64;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {   // two 16-column halves of the 64-byte row (bounded register use)
                const int jh = jb + 16 * hf;
                uint4 rr[2];
                if (has_res) {
#pragma unroll
                  for (int h = 0; h < 2; ++h)
                    rr[h] = (row_ok && jh + 8 * h < p.N) ? *reinterpret_cast<const uint4*>(res_row16 + jh + 8 * h)
                                                         : make_uint4(0u, 0u, 0u, 0u);
                }
                uint32_t v[16];
                __syncwarp();
                tmem_ld16(lane_addr + c0 + 16 * hf, v);
                tmem_ld_wait();
                float x[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + jh + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                  x[4 * c + 0] = __uint_as_float(v[4 * c + 0]) + bv.x;
                  x[4 * c + 1] = __uint_as_float(v[4 * c + 1]) + bv.y;
                  x[4 * c + 2] = __uint_as_float(v[4 * c + 2]) + bv.z;
                  x[4 * c + 3] = __uint_as_float(v[4 * c + 3]) + bv.w;
                }
                auto add_res = [&]() {
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    x[8 * h + 0] += bf16_lo(rr[h].x); x[8 * h + 1] += bf16_hi(rr[h].x);
                    x[8 * h + 2] += bf16_lo(rr[h].y); x[8 * h + 3] += bf16_hi(rr[h].y);
                    x[8 * h + 4] += bf16_lo(rr[h].z); x[8 * h + 5] += bf16_hi(rr[h].z);
                    x[8 * h + 6] += bf16_lo(rr[h].w); x[8 * h + 7] += bf16_hi(rr[h].w);
                  }
                };
                if (pre) add_res();
                switch (p.act) {
                  case SVX_ACT_RELU:
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
                    break;
                  case SVX_ACT_LEAKY:
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = act_t<SVX_ACT_LEAKY>(x[q], p.act_param);
                    break;
                  case SVX_ACT_GELU:
#pragma unroll
                    for (int q = 0; q < 16; q += 2) gelu_erf2_fast(x[q], x[q + 1]);
                    break;
                  default: break;
                }
                if (post) add_res();
                if (p.out_scale != 1.f) {
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] *= p.out_scale;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
                  *reinterpret_cast<uint4*>(my_row + (((2 * hf + c) ^ sw) << 4)) =
                      make_uint4(pack_bf16x2(x[8 * c], x[8 * c + 1]), pack_bf16x2(x[8 * c + 2], x[8 * c + 3]),
                                 pack_bf16x2(x[8 * c + 4], x[8 * c + 5]), pack_bf16x2(x[8 * c + 6], x[8 * c + 7]));
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&map_c, stg_u32 + buf * 2048, jb, m0 + quarter * 32);
                tma_store_commit();
              }
              buf = (buf + 1u) % C::kEpiBufs;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(as));
            continue;
          }
        }
        const bool pool8 = p.epi_mode == SVX_EPI_POOL8;
        const int ncols = pool8 ? (p.N >> 3) : BN;          // output columns this tile produces
        // The residual does not depend on the accumulator: the residual of the first kAhead blocks is requested before
        // waiting for the MMAs of this tile, and each block requests the one kAhead blocks later while it works, so
        // several HBM round trips per warp are in flight (a shift-register of float4 quads keeps the indices static).
        constexpr int kAhead = 1;   // deeper prefetch measured slower: the row-per-lane loads congest the L1 pipeline
        float4 rq[kAhead][4];
        auto load_res = [&](int jbn, float4 (&dst)[4]) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[c] = (row_ok && jbn + 4 * c < p.N) ? *reinterpret_cast<const float4*>(res_row + jbn + 4 * c)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if (has_res) {
#pragma unroll
          for (int a = 0; a < kAhead; ++a) {
            const int c0a = (part + a * nparts) * SLAB;
            if (c0a < ncols && n0 + c0a < p.N) load_res(n0 + c0a, rq[a]);
          }
        }
        mbar_wait(tmem_full_bar(as), (it >> 1) & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = part * SLAB; c0 < ncols; c0 += nparts * SLAB) {
          const int jb = n0 + c0;
          if (jb >= p.N) break;  // warp-uniform
          uint32_t v[SLAB];
          __syncwarp();
          tmem_ld16(lane_addr + c0, v);
          if (pool8) {   // max over the eight conv positions of this pooled voxel (column groups of N/8)
            tmem_ld_wait();
#pragma unroll 1
            for (int gq = 1; gq < 8; ++gq) {
              uint32_t w[SLAB];
              tmem_ld16(lane_addr + gq * ncols + c0, w);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < SLAB; ++q) v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), __uint_as_float(w[q])));
            }
          }
          float4 rv[4];
          if (has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) rv[c] = rq[0][c];
#pragma unroll
            for (int a = 0; a + 1 < kAhead; ++a)
#pragma unroll
              for (int c = 0; c < 4; ++c) rq[a][c] = rq[a + 1][c];
            const int c0n = c0 + kAhead * nparts * SLAB;
            if (c0n < ncols && n0 + c0n < p.N) load_res(n0 + c0n, rq[kAhead - 1]);
          }
          float4 bv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            bv[c] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + jb + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          float x[SLAB];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            x[4 * c + 0] = __uint_as_float(v[4 * c + 0]) + bv[c].x;
            x[4 * c + 1] = __uint_as_float(v[4 * c + 1]) + bv[c].y;
            x[4 * c + 2] = __uint_as_float(v[4 * c + 2]) + bv[c].z;
            x[4 * c + 3] = __uint_as_float(v[4 * c + 3]) + bv[c].w;
          }
          if (pre) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { x[4 * c] += rv[c].x; x[4 * c + 1] += rv[c].y; x[4 * c + 2] += rv[c].z; x[4 * c + 3] += rv[c].w; }
          }
          switch (p.act) {
            case SVX_ACT_RELU:
#pragma unroll
              for (int q = 0; q < SLAB; ++q) x[q] = fmaxf(x[q], 0.f);
              break;
            case SVX_ACT_LEAKY:
#pragma unroll
              for (int q = 0; q < SLAB; ++q) x[q] = act_t<SVX_ACT_LEAKY>(x[q], p.act_param);
              break;
            case SVX_ACT_GELU:
#pragma unroll
              for (int q = 0; q < SLAB; q += 2) gelu_erf2_fast(x[q], x[q + 1]);
              break;
            default: break;
          }
          if (post) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { x[4 * c] += rv[c].x; x[4 * c + 1] += rv[c].y; x[4 * c + 2] += rv[c].z; x[4 * c + 3] += rv[c].w; }
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int q = 0; q < SLAB; ++q) x[q] *= p.out_scale;
          }
          if (p.round_tf32) {
#pragma unroll
            for (int q = 0; q < SLAB; ++q) x[q] = round_tf32(x[q]);
          }
          // this buffer's previous store (two blocks ago) must have been read out of smem by the TMA unit
          if (lane == 0) tma_store_wait_read<C::kEpiBufs - 1>();
          __syncwarp();
          char* my_row = stg + buf * 2048 + lane * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<float4*>(my_row + ((c ^ sw) << 4)) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, stg_u32 + buf * 2048, jb, m0 + quarter * 32);
            tma_store_commit();
          }
          buf = (buf + 1u) % C::kEpiBufs;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(as));
        continue;
      }
      long long off = 0, off2 = 0;
      const bool valid = decode_row(p, m0 + quarter * 32 + lane, off, off2);
      __syncwarp();
      soff[lane] = valid ? off : -1;
      __syncwarp();
      mbar_wait(tmem_full_bar(as), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = part * SLAB; c0 < BN; c0 += nparts * SLAB) {
        const int jb = n0 + c0;
        if (jb >= p.N) break;  // warp-uniform
        uint32_t v[SLAB];
        __syncwarp();
        tmem_ld16(lane_addr + c0, v);
        tmem_ld_wait();
        if (!rowwise) {
          // transpose through smem so that every warp-wide store writes whole 64-byte row segments
#pragma unroll
          for (int c = 0; c < CP; ++c)
            *reinterpret_cast<uint4*>(staging + lane * SLAB + ((c ^ (lane & (CP - 1))) << 2)) =
                make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          __syncwarp();
          if (BF && p.io16) {
            switch (p.act) {
              case SVX_ACT_RELU: store_slab<SLAB, SVX_ACT_RELU, true>(p, staging, soff, lane, jb); break;
              case SVX_ACT_LEAKY: store_slab<SLAB, SVX_ACT_LEAKY, true>(p, staging, soff, lane, jb); break;
              case SVX_ACT_GELU: store_slab<SLAB, SVX_ACT_GELU, true>(p, staging, soff, lane, jb); break;
              default: store_slab<SLAB, SVX_ACT_NONE, true>(p, staging, soff, lane, jb); break;
            }
          } else {
            switch (p.act) {
              case SVX_ACT_RELU: store_slab<SLAB, SVX_ACT_RELU>(p, staging, soff, lane, jb); break;
              case SVX_ACT_LEAKY: store_slab<SLAB, SVX_ACT_LEAKY>(p, staging, soff, lane, jb); break;
              case SVX_ACT_GELU: store_slab<SLAB, SVX_ACT_GELU>(p, staging, soff, lane, jb); break;
              default: store_slab<SLAB, SVX_ACT_NONE>(p, staging, soff, lane, jb); break;
            }
          }
        } else if (!BF && valid) {   // (row-wise epilogues exist for fp32 storage only: decoder / refiner tails)
          // one thread = one output row (decoder tail needs the whole row; N % 4 != 0 outputs are scalar)
          float x[SLAB];
#pragma unroll
          for (int q = 0; q < SLAB; ++q) x[q] = __uint_as_float(v[q]) + ((p.bias && jb + q < p.Npad) ? __ldg(p.bias + jb + q) : 0.f);
          if (convt) {
            // all eight output-parity classes of a stride-2 transposed convolution sit side by side in N
            auto cls_off = [&](int cls) { return (cls >> 2) * p.c_sd + ((cls >> 1) & 1) * p.c_sh + (cls & 1) * p.c_sw; };
            if (p.epi_aux) {
              // decoder tail (decoder.py:80-89) per class: 8 channels per class, two classes (pw = 0, 1) per slab
              const float b8 = __ldg(p.epi_aux + 8);
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int cls = (jb >> 3) + hf;
                float gsum = b8;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  x[8 * hf + q] = fmaxf(x[8 * hf + q], 0.f);
                  gsum = fmaf(__ldg(p.epi_aux + q), x[8 * hf + q], gsum);
                }
                float* dst = p.out + off + cls_off(cls);
                const float* xs = x + 8 * hf;
                if (p.round_tf32) {
                  *reinterpret_cast<float4*>(dst) = make_float4(round_tf32(xs[0]), round_tf32(xs[1]), round_tf32(xs[2]), round_tf32(xs[3]));
                  *reinterpret_cast<float4*>(dst + 4) = make_float4(round_tf32(xs[4]), round_tf32(xs[5]), round_tf32(xs[6]), round_tf32(xs[7]));
                  *reinterpret_cast<float4*>(dst + 8) = make_float4(round_tf32(gsum), 0.f, 0.f, 0.f);
                } else {
                  *reinterpret_cast<float4*>(dst) = make_float4(xs[0], xs[1], xs[2], xs[3]);
                  *reinterpret_cast<float4*>(dst + 4) = make_float4(xs[4], xs[5], xs[6], xs[7]);
                  *reinterpret_cast<float4*>(dst + 8) = make_float4(gsum, 0.f, 0.f, 0.f);
                }
                p.out2[off2 + (cls >> 2) * p.c2_sd + ((cls >> 1) & 1) * p.c2_sh + (cls & 1) * p.c2_sw] = gsum;
              }
            } else {
              const int cc = p.cls_cout;
              if ((cc & 15) == 0 && p.vec_ok) {
                // the 16 columns of this pass are 16 consecutive channels of ONE class: 64 contiguous bytes per voxel
                const int cls = jb / cc;
                const long long o = off + cls_off(cls) + (jb - cls * cc);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float4 r4 = p.residual ? *reinterpret_cast<const float4*>(p.residual + o + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
                  const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
                  float t4[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float t = x[4 * c + e];
                    if (!p.res_after_act) t += rv[e];
                    t = apply_act(t, p.act, p.act_param);
                    if (p.res_after_act) t += rv[e];
                    t *= p.out_scale;
                    t4[e] = p.round_tf32 ? round_tf32(t) : t;
                  }
                  *reinterpret_cast<float4*>(p.out + o + 4 * c) = make_float4(t4[0], t4[1], t4[2], t4[3]);
                }
                continue;
              }
#pragma unroll
              for (int q = 0; q < SLAB; ++q) {
                const int j = jb + q;
                if (j < p.N) {
                  const int cls = j / cc;
                  const long long o = off + cls_off(cls) + (j - cls * cc);
                  const float rvq = p.residual ? p.residual[o] : 0.f;
                  float t = x[q];
                  if (!p.res_after_act) t += rvq;
                  t = apply_act(t, p.act, p.act_param);
                  if (p.res_after_act) t += rvq;
                  t *= p.out_scale;
                  p.out[o] = p.round_tf32 ? round_tf32(t) : t;
                }
              }
            }
            continue;
          }
          if (p.epi_mode == SVX_EPI_DEC_TAIL) {
            // decoder.py:80-89: raw = cat(relu(bn(layer4)), layer5(.)) ; coarse = layer5(.)
            float gsum = __ldg(p.epi_aux + 8);  // layer5 bias (0 when TCONV_USE_BIAS is off)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              x[q] = fmaxf(x[q], 0.f);
              gsum = fmaf(__ldg(p.epi_aux + q), x[q], gsum);
            }
            x[8] = gsum;
#pragma unroll
            for (int q = 9; q < SLAB; ++q) x[q] = 0.f;
            p.out2[off2] = gsum;
          } else {
            const float* res = p.residual ? p.residual + off + jb : nullptr;
#pragma unroll 4
            for (int q = 0; q < SLAB; ++q) {
              const float rvq = (res && jb + q < p.N) ? res[q] : 0.f;
              float t = x[q];
              if (!p.res_after_act) t += rvq;
              t = apply_act(t, p.act, p.act_param);
              if (p.res_after_act) t += rvq;
              x[q] = t * p.out_scale;
            }
          }
          if (p.round_tf32) {
#pragma unroll
            for (int q = 0; q < SLAB; ++q) x[q] = round_tf32(x[q]);
          }
          float* dst = p.out + off + jb;
          if (p.vec_ok) {
#pragma unroll
            for (int q = 0; q < SLAB; q += 4) {
              if (jb + q < p.N) *reinterpret_cast<float4*>(dst + q) = make_float4(x[q], x[q + 1], x[q + 2], x[q + 3]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < SLAB; ++q) {
              if (jb + q < p.N) dst[q] = x[q];
            }
          }
        }
      }
      // hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(as));
    }
    if (p.epi_tma && lane == 0) tma_store_wait_all<0>();   // smem must outlive the last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<C::kTmemCols>(tmem_base);
}

template <int BN, int PARTS>
__global__ void __launch_bounds__(Cfg<BN, PARTS>::kThreadsT, Cfg<BN, PARTS>::kCtasPerSm)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ GemmParams p) {
  gemm_body<BN, PARTS, false>(map_a, map_b, map_c, map_r, p);
}

template <int BN, int PARTS>
__global__ void __launch_bounds__(Cfg<BN, PARTS>::kThreadsT, Cfg<BN, PARTS>::kCtasPerSm)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ GemmParams p) {
  gemm_body<BN, PARTS, true>(map_a, map_b, map_c, map_r, p);
}


// =====================================================================================================
// Slab kernel: 3x3x3 stride-1 Conv3d with <= 16 output channels over a zero-bordered channels-last volume
// (the merger's layers, merger.py:20-54).  With so few output channels a tcgen05.mma is bound by reading its
// A operand from shared memory (~76 cycles per 128x8 fp32 block whatever N is, measured), so the kernel
//   * marches along depth: one work unit = (volume, 126-row column of the (h,w) plane); each 200-row depth slab
//     is loaded ONCE by TMA into a ring and serves the three output depths that touch it,
//   * folds the kw taps into the MMA's N dimension: D[u, kw*16+co] = sum_{kd,kh,c} X[u + (kd*Hp+kh)*Wp, c] *
//     W[kd,kh,kw,c,co], nine shifted-A MMAs per k-step instead of 27 (the (kd,kh) shift is a start-address shift
//     of the smem descriptor inside the slab; the 128B swizzle is a function of the absolute smem address),
//   * finishes out[u, co] = D[u, 0, co] + D[u+1, 1, co] + D[u+2, 2, co] in the epilogue with warp shuffles (rows
//     are TMEM lanes); the two rows that cross a warp's lane quarter go through a small smem exchange, and tiles
//     overlap by two rows.
// Weights ([48, 9*32], row = kw*16+co, col = (kd*3+kh)*32 + c) stay resident in smem; eight accumulators ring
// in TMEM so the epilogue of one depth overlaps the MMAs of the next.
// =====================================================================================================
constexpr int S3_ROWS = 200;                       // slab rows: 128 + 2*Wp <= 200  (Wp <= 36)
#ifndef S3_PARTS
#define S3_PARTS 1                                 // a slab arrives as S3_PARTS TMA boxes (1 or 5; measured: no difference)
#endif
constexpr int S3_PART_ROWS = S3_ROWS / S3_PARTS;   // multiple of 8 (whole swizzle atoms)
constexpr int S3_N = 48;                           // MMA N: 3 kw groups x 16 output channels
constexpr int S3_NACC = 8, S3_ACC_STRIDE = 64;     // TMEM: 8 accumulators of 48 (stride 64) columns
constexpr int S3_STEP = BM - 2;                    // valid rows per tile
constexpr int S3_THREADS = 320;                    // warps 0-3 / 6-9: two epilogue sets, 4: TMA, 5: MMA
constexpr int S3_XCHG_BYTES = 2 * BM * 128;        // per epilogue set: 128 rows x (D1[16] | D2[16]) staged for the row shift

// ROWB = bytes per staged row: 128 (32-channel box, 128B swizzle) or 64 (layers with <= 16 live channels: 16-channel
// box, 64B swizzle -- half the L2 traffic and shared memory, which buys a twice deeper slab ring)
// F16: the MMA operands are fp16 (kind::f16, K = 16 per instruction: half the issue-bound instructions of kind::tf32 for
// the same channels).  HBM tensors stay fp32: each fp32 slab lands in a small staging ring by TMA, four converter warps
// (10-13) rewrite it as an fp16 operand slab (rows of ROWB/2 bytes, the narrower swizzle), and convert the weights once
// per CTA.  The activations and weights of this path are stored TF32-rounded (10 mantissa bits = fp16's), so the
// conversion is exact for |x| in [6.1e-5, 65504]; smaller magnitudes lose relative, not absolute, accuracy (< 3e-8).
constexpr int S3_THREADS_F16 = 448;
template <int ROWB, bool F16 = false>
struct S3Cfg {
  static constexpr int kOpRowB = F16 ? ROWB / 2 : ROWB;            // bytes per operand row
  static constexpr int kSlabBytes = S3_ROWS * kOpRowB;
  static constexpr int kNSlab = F16 ? (ROWB == 128 ? 4 : 10) : (ROWB == 128 ? 5 : 10);
  // fp32 staging slabs (F16 only): the TMA round trip is hidden by THIS ring now (three slots left the kernel
  // latency-bound: one slab per third of a round trip), the operand ring only decouples converters and MMAs
  static constexpr int kNStg = F16 ? (ROWB == 128 ? 3 : 6) : 0;
  // F16: two exchange buffers per epilogue set, so a tile needs ONE 128-thread barrier (write | read) instead of two
  static constexpr int kXchgBufs = F16 ? 2 : 1;
  static constexpr int kXchgBytes = kXchgBufs * S3_XCHG_BYTES;
  static constexpr int kTapBytes = S3_N * kOpRowB;     // one (kd,kh) weight block
  static constexpr int kWBytes = (9 * kTapBytes + 1023) / 1024 * 1024;
  static constexpr int kStgBytes = F16 ? S3_ROWS * ROWB : 0;
  static constexpr int kSmem = 1024 + kWBytes + kNSlab * kSlabBytes + kNStg * kStgBytes + kXchgBytes + 512;
  static_assert(kSmem <= 232448, "slab kernel shared memory");
};

template <int ROWB, bool PAIR, bool F16>
__global__ void __launch_bounds__(F16 ? S3_THREADS_F16 : S3_THREADS, 1)
conv3_slab_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ GemmParams p) {
  static_assert(!(PAIR && F16), "the CTA-pair mode exists for kind::tf32 only");
  using SC = S3Cfg<ROWB, F16>;
  constexpr int S3_SLAB_BYTES = SC::kSlabBytes, S3_NSLAB = SC::kNSlab;
  constexpr int S3_TAP_BYTES = SC::kTapBytes, S3_W_BYTES = SC::kWBytes;
  constexpr int OPB = SC::kOpRowB;   // bytes per operand row in shared memory
  constexpr int kBoxCh = ROWB / 4;   // channels per staged row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_smem = smem_base;
  // weights | fp32 staging ring (F16 only; 1024-byte aligned: its TMA boxes use the 128B swizzle) | operand slab ring
  constexpr int kStgOff = S3_W_BYTES;
  const uint32_t stg_smem = smem_base + kStgOff;
  constexpr int S3_NSTG = SC::kNStg > 0 ? SC::kNStg : 1;
  constexpr int kSlabOff = kStgOff + SC::kNStg * SC::kStgBytes;
  static_assert(kSlabOff % 1024 == 0, "operand slab ring alignment");
  const uint32_t slab_smem = smem_base + kSlabOff;
  constexpr int kXchgOff = kSlabOff + S3_NSLAB * S3_SLAB_BYTES;
  float* xchg = reinterpret_cast<float*>(smem_gen + kXchgOff);
  const uint32_t bar_base = smem_base + kXchgOff + SC::kXchgBytes;
  // barriers: slab_full[5], slab_empty[5], w_full, acc_full[8], acc_empty[8], tmem slot
  auto slab_full = [&](uint32_t s) { return bar_base + 8u * s; };
  auto slab_empty = [&](uint32_t s) { return bar_base + 8u * (S3_NSLAB + s); };
  const uint32_t w_full = bar_base + 8u * (2 * S3_NSLAB);
  auto acc_full = [&](uint32_t a) { return bar_base + 8u * (2 * S3_NSLAB + 1 + a); };
  auto acc_empty = [&](uint32_t a) { return bar_base + 8u * (2 * S3_NSLAB + 1 + S3_NACC + a); };
  auto stg_full = [&](uint32_t s) { return bar_base + 8u * (2 * S3_NSLAB + 1 + 2 * S3_NACC + s); };
  auto stg_empty = [&](uint32_t s) { return bar_base + 8u * (2 * S3_NSLAB + 1 + 2 * S3_NACC + S3_NSTG + s); };
  constexpr int kSlotIdx = 2 * S3_NSLAB + 1 + 2 * S3_NACC + 2 * S3_NSTG;
  const uint32_t tmem_slot = bar_base + 8u * kSlotIdx;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kXchgOff + SC::kXchgBytes + 8 * kSlotIdx);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: the two CTAs of a cluster work on two different units in lock step; the leader (rank 0) issues ONE
  // tcgen05.mma.cta_group::2 (M = 256: 128 rows from each CTA's slab, N/2 weight rows from each CTA) per step for both,
  // halving the per-SM cost of the issue-bound ~89-cycle instructions (profiles/r1_umma_tf32_issue_cost.txt)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0u;
  const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // scheduling slot (pair index)
  const int nslots = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int Wp = p.in_W, HWp = p.in_H * p.in_W;
  const int nd = p.valid_D;                 // output depths per unit; the unit streams nd + 2 slabs
  const int ncol = p.tiles_n;               // 126-row columns per (h,w) plane
  const int num_units = p.tiles_m;          // volumes x columns
  const int num_steps = PAIR ? (num_units + 1) / 2 : num_units;   // PAIR: step q covers units 2q (leader) and 2q+1 (peer)
  auto unit_of = [&](int q) { const int u = PAIR ? 2 * q + (int)rank : q; return u < num_units ? u : num_units - 1; };
  auto unit_real = [&](int q) { return (PAIR ? 2 * q + (int)rank : q) < num_units; };

  if (warp == 4 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
  if (warp == 5) {
    if (lane == 0) {
      for (uint32_t s = 0; s < S3_NSLAB; ++s) { mbar_init(slab_full(s), F16 ? 4u : 1u); mbar_init(slab_empty(s), 1u); }
      mbar_init(w_full, F16 ? 4u : 1u);
      for (uint32_t s = 0; s < S3_NSTG; ++s) { mbar_init(stg_full(s), 1u); mbar_init(stg_empty(s), 4u); }
      // PAIR: the leader's acc_empty collects the epilogue warps of both CTAs
      for (uint32_t a = 0; a < S3_NACC; ++a) { mbar_init(acc_full(a), 1u); mbar_init(acc_empty(a), PAIR ? 8u : 4u); }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) tmem_alloc_pair<512>(tmem_slot); else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 4) {
    // ---- TMA producer: weights once, then one slab per (unit, plane) ---------------------------------
    if (elect_one()) {
      // PAIR: every load of either CTA signals the LEADER's barrier (it is the leader that issues the MMAs); the
      // leader expects the bytes of both.  Each CTA holds half of the weight rows of every tap (map_w boxes 24 rows).
      constexpr uint32_t kWTap = PAIR ? S3_TAP_BYTES / 2 : S3_TAP_BYTES;
      if constexpr (!F16) {
        if (leader) mbar_arrive_expect_tx(w_full, 9 * S3_TAP_BYTES);
        for (int t = 0; t < 9; ++t) {   // first kBoxCh of the 32
          if (PAIR) tma_load_2d_pair(w_smem + t * kWTap, &map_w, leader_addr(w_full), t * BK, (int)rank * (S3_N / 2));
          else tma_load_2d(w_smem + t * S3_TAP_BYTES, &map_w, w_full, t * BK, 0);
        }
      }
      uint32_t g = 0;
      for (int q = cta; q < num_steps; q += nslots) {
        const int unit = unit_of(q);
        const int n = unit / ncol, col = unit - n * ncol;
        const int row0 = n * p.in_D * HWp + col * S3_STEP;
        for (int pl = 0; pl < nd + 2; ++pl, ++g) {
          if constexpr (F16) {   // fp32 slab -> staging ring; the converter warps fill the operand ring
            const uint32_t s = g % S3_NSTG;
            mbar_wait(stg_empty(s), ((g / S3_NSTG) & 1u) ^ 1u);
            mbar_arrive_expect_tx(stg_full(s), SC::kStgBytes);
            tma_load_2d(stg_smem + s * SC::kStgBytes, &map_x, stg_full(s), p.in_c0, row0 + pl * HWp);
            continue;
          }
          const uint32_t s = g % S3_NSLAB;
          mbar_wait(slab_empty(s), ((g / S3_NSLAB) & 1u) ^ 1u);
          if (leader) mbar_arrive_expect_tx(slab_full(s), (PAIR ? 2u : 1u) * S3_SLAB_BYTES);
#pragma unroll
          for (int part = 0; part < S3_PARTS; ++part) {   // several boxes in flight: the TMA unit walks a box row by row
            const uint32_t dst = slab_smem + s * S3_SLAB_BYTES + part * S3_PART_ROWS * ROWB;
            const int row = row0 + pl * HWp + part * S3_PART_ROWS;
            if (PAIR) tma_load_2d_pair(dst, &map_x, leader_addr(slab_full(s)), p.in_c0, row);
            else tma_load_2d(dst, &map_x, slab_full(s), p.in_c0, row);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---- MMA issuer -----------------------------------------------------------------------------------
    if (leader && elect_one()) {
      constexpr uint32_t idesc = F16 ? umma_idesc_f16(BM, S3_N) : umma_idesc_tf32(PAIR ? 2 * BM : BM, S3_N);
      constexpr uint32_t kWTap = PAIR ? S3_TAP_BYTES / 2 : S3_TAP_BYTES;
      const int ksteps = F16 ? (p.cin_live + 15) / 16 : (p.cin_live + UMMA_K - 1) / UMMA_K;   // 32 operand bytes per step
      mbar_wait(w_full, 0u);
      tc_fence_after();
      uint32_t sbase = 0, waited = 0, tg = 0;
      for (int q = cta; q < num_steps; q += nslots) {
        for (int d = 0; d < nd; ++d, ++tg) {
          while (waited < sbase + d + 3) {   // slabs d, d+1, d+2 of this unit
            mbar_wait(slab_full(waited % S3_NSLAB), (waited / S3_NSLAB) & 1u);
            ++waited;
          }
          const uint32_t a = tg % S3_NACC;
          mbar_wait(acc_empty(a), ((tg / S3_NACC) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t acc = tmem_base + a * S3_ACC_STRIDE;
          for (int kd = 0; kd < 3; ++kd) {
            const uint32_t slab = slab_smem + ((sbase + d + kd) % S3_NSLAB) * S3_SLAB_BYTES;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint64_t da = OPB == 128 ? umma_desc_sw128(slab + kh * Wp * OPB)
                                  : OPB == 64 ? umma_desc_sw64(slab + kh * Wp * OPB) : umma_desc_sw32(slab + kh * Wp * OPB);
              const uint64_t db = OPB == 128 ? umma_desc_sw128(w_smem + (kd * 3 + kh) * kWTap)
                                  : OPB == 64 ? umma_desc_sw64(w_smem + (kd * 3 + kh) * kWTap)
                                              : umma_desc_sw32(w_smem + (kd * 3 + kh) * kWTap);
#ifdef SVX_SLAB_NOMMA   // experiment: the TMA pipeline alone (results are wrong)
              if (ksteps > 0) continue;
#endif
              for (int k = 0; k < ksteps; ++k) {
                if (PAIR) umma_tf32_pair(acc, da + 2u * k, db + 2u * k, idesc, (kd | kh | k) != 0 ? 1u : 0u);
                else if (F16) umma_f16(acc, da + 2u * k, db + 2u * k, idesc, (kd | kh | k) != 0 ? 1u : 0u);
                else umma_tf32(acc, da + 2u * k, db + 2u * k, idesc, (kd | kh | k) != 0 ? 1u : 0u);
              }
            }
          }
          auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
          commit(slab_empty((sbase + d) % S3_NSLAB));   // plane d is not needed by later depths
          if (d == nd - 1) {
            commit(slab_empty((sbase + d + 1) % S3_NSLAB));
            commit(slab_empty((sbase + d + 2) % S3_NSLAB));
          }
          commit(acc_full(a));
        }
        sbase += nd + 2;
      }
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ---- converter warps (F16 only): fp32 staging slab -> fp16 operand slab; weights once --------------------
    if constexpr (F16) {
      constexpr int UIN = ROWB / 16;    // 16-byte units per fp32 row (4 floats each)
      constexpr int UOUT = OPB / 16;    // 16-byte units per fp16 row (8 halves each)
      const int ct = threadIdx.x - 320; // 0..127
      uint8_t* w_gen = smem_gen;
      // an fp16 operand row r of OPB bytes: unit u is stored at u ^ swz(r) (32B swizzle: bit 2 of r; 64B: bits 1-2)
      auto swz_out = [](int r) { return OPB == 32 ? ((r >> 2) & 1) : ((r >> 1) & 3); };
      auto swz_in = [](int r) { return ROWB == 64 ? ((r >> 1) & 3) : (r & 7); };
      auto pack8 = [](const float4& a, const float4& b) {
        uint4 o;
        // saturating: |x| > 65504 becomes +-65504 instead of inf (and is reported through range_flag below)
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(a.y), "f"(a.x));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(a.w), "f"(a.z));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(b.y), "f"(b.x));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(b.w), "f"(b.z));
        return o;
      };
      // running max |x| over everything this thread converts (integer max of the sign-stripped bit patterns; NaN > inf)
      uint32_t amax = 0u;
      auto track = [&](const float4& a, const float4& b) {
        const uint32_t m0 = max(max(__float_as_uint(a.x) & 0x7fffffffu, __float_as_uint(a.y) & 0x7fffffffu),
                                max(__float_as_uint(a.z) & 0x7fffffffu, __float_as_uint(a.w) & 0x7fffffffu));
        const uint32_t m1 = max(max(__float_as_uint(b.x) & 0x7fffffffu, __float_as_uint(b.y) & 0x7fffffffu),
                                max(__float_as_uint(b.z) & 0x7fffffffu, __float_as_uint(b.w) & 0x7fffffffu));
        amax = max(amax, max(m0, m1));
      };
      // weights: W[row = kw*16+co][col = tap*32 + c] fp32 -> per tap a [48 rows x kBoxCh halves] swizzled block
      for (int i = ct; i < 9 * S3_N * UOUT; i += 128) {
        const int u = i % UOUT, row = (i / UOUT) % S3_N, t = i / (UOUT * S3_N);
        const float* src = p.Wg + (long long)row * (9 * BK) + t * BK + u * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        *reinterpret_cast<uint4*>(w_gen + t * S3_TAP_BYTES + row * OPB + ((u ^ swz_out(row)) << 4)) = pack8(a, b);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_full);
      uint32_t g = 0;
      for (int q = cta; q < num_steps; q += nslots) {
        for (int pl = 0; pl < nd + 2; ++pl, ++g) {
          const uint32_t ss = g % S3_NSTG, so = g % S3_NSLAB;
          mbar_wait(stg_full(ss), (g / S3_NSTG) & 1u);
          mbar_wait(slab_empty(so), ((g / S3_NSLAB) & 1u) ^ 1u);
          const uint8_t* src = smem_gen + kStgOff + ss * SC::kStgBytes;
          uint8_t* dst = smem_gen + kSlabOff + so * S3_SLAB_BYTES;
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = ct + rr * 128;
            if (r < S3_ROWS) {
              const uint8_t* srow = src + r * ROWB;
              uint8_t* drow = dst + r * OPB;
              const int si = swz_in(r), sw_o = swz_out(r);
#pragma unroll
              for (int u = 0; u < UOUT; ++u) {
                const float4 a = *reinterpret_cast<const float4*>(srow + (((2 * u) ^ si) << 4));
                const float4 b = *reinterpret_cast<const float4*>(srow + (((2 * u + 1) ^ si) << 4));
                track(a, b);
                *reinterpret_cast<uint4*>(drow + ((u ^ sw_o) << 4)) = pack8(a, b);
              }
            }
          }
          static_assert(UIN == 2 * UOUT, "two fp32 units make one fp16 unit");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(slab_full(so));
            mbar_arrive(stg_empty(ss));
          }
        }
      }
      if (p.range_flag && amax > 0x477fe000u) atomicOr(p.range_flag, 1);   // an activation beyond 65504: saturated above
    }
  } else {
    // ---- epilogue: one thread = one row u of the tile; out[u] = D0[u] + D1[u+1] + D2[u+2].  Two sets of four
    // warps (0-3: even tiles, 6-9: odd tiles) so two tiles drain concurrently; the row shift goes through a
    // swizzled smem staging buffer per set (rows u+1, u+2 may belong to another warp's TMEM lane quarter). ------
    const int quarter = warp & 3;
    const uint32_t set = warp >= 6 ? 1u : 0u;
    const int r = quarter * 32 + lane;
    const int r1 = min(r + 1, BM - 1), r2 = min(r + 2, BM - 1);
    char* xb0 = reinterpret_cast<char*>(xchg) + set * (SC::kXchgBufs * BM * 128);
    uint32_t xk = 0;   // tiles this set has exchanged (selects the buffer)
    const int barid = 1 + static_cast<int>(set);
    float bias[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) bias[q] = p.bias ? __ldg(p.bias + q) : 0.f;
    const float slope = p.act_param, scale = p.out_scale, acc_scale = p.acc_scale;
    const bool has_res = p.residual != nullptr;
    const bool pre = has_res && !p.res_after_act, post = has_res && p.res_after_act;
    const bool full16 = p.N == 16 && p.vec_ok;
    uint32_t tg = 0;
    for (int q = cta; q < num_steps; q += nslots) {
      const int unit = unit_of(q);
      const int n = unit / ncol, col = unit - n * ncol;
      const int u = col * S3_STEP + r;
      const int h = u / Wp, w = u - h * Wp;
      const bool row_ok = r < S3_STEP && h < p.valid_H && w < p.valid_W && unit_real(q);
      const long long off0 = p.o_base + n * p.o_sn + h * p.o_sh + w * p.o_sw;
      for (int d = 0; d < nd; ++d, ++tg) {
        if ((tg & 1u) != set) continue;
        const uint32_t a = tg % S3_NACC;
        mbar_wait(acc_full(a), (tg / S3_NACC) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + a * S3_ACC_STRIDE + (static_cast<uint32_t>(quarter * 32) << 16);
        uint32_t d0[16], d1[16], d2[16];
        __syncwarp();
        tmem_ld16(taddr, d0);
        tmem_ld16(taddr + 16, d1);
        tmem_ld16(taddr + 32, d2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {   // the accumulator is in registers: hand it back (to the leader, who issues the MMAs)
          if (PAIR && !leader) mbar_arrive_cluster(leader_addr(acc_empty(a)));
          else mbar_arrive(acc_empty(a));
        }
#ifdef SVX_SLAB_NOEPI   // experiment: how fast is the TMA + MMA side alone? (results are wrong)
        if (d0[0] != 0x7fc12345u) continue;
#endif
        char* xb = xb0 + (SC::kXchgBufs == 2 ? (xk & 1u) * (BM * 128) : 0u);
        ++xk;
        char* my_row = xb + r * 128;
        const char* row1 = xb + r1 * 128;
        const char* row2 = xb + r2 * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(my_row + ((c ^ (r & 7)) << 4)) = make_uint4(d1[4 * c], d1[4 * c + 1], d1[4 * c + 2], d1[4 * c + 3]);
          *reinterpret_cast<uint4*>(my_row + (((c + 4) ^ (r & 7)) << 4)) = make_uint4(d2[4 * c], d2[4 * c + 1], d2[4 * c + 2], d2[4 * c + 3]);
        }
        named_bar_sync(barid, 128);
        float x[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 s1 = *reinterpret_cast<const float4*>(row1 + ((c ^ (r1 & 7)) << 4));
          const float4 s2 = *reinterpret_cast<const float4*>(row2 + (((c + 4) ^ (r2 & 7)) << 4));
          x[4 * c + 0] = fmaf((__uint_as_float(d0[4 * c + 0]) + s1.x) + s2.x, acc_scale, bias[4 * c + 0]);
          x[4 * c + 1] = fmaf((__uint_as_float(d0[4 * c + 1]) + s1.y) + s2.y, acc_scale, bias[4 * c + 1]);
          x[4 * c + 2] = fmaf((__uint_as_float(d0[4 * c + 2]) + s1.z) + s2.z, acc_scale, bias[4 * c + 2]);
          x[4 * c + 3] = fmaf((__uint_as_float(d0[4 * c + 3]) + s1.w) + s2.w, acc_scale, bias[4 * c + 3]);
        }
        // one buffer: it may be overwritten by this set's next tile.  Two buffers: the next tile writes the other one, and
        // nobody writes this one again before passing the next tile's barrier, i.e. after every thread finished this read
        if (SC::kXchgBufs == 1) named_bar_sync(barid, 128);
        if (row_ok) {
          const long long off = off0 + d * p.o_sd;
          float* dst = p.out + off;
          if (full16) {
            if (has_res) {
              float rv[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float4 t4 = *reinterpret_cast<const float4*>(p.residual + off + 4 * c);
                rv[4 * c] = t4.x; rv[4 * c + 1] = t4.y; rv[4 * c + 2] = t4.z; rv[4 * c + 3] = t4.w;
              }
              if (pre) {
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q] += rv[q];
              }
              switch (p.act) {
                case SVX_ACT_RELU:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
                  break;
                case SVX_ACT_LEAKY:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = act_t<SVX_ACT_LEAKY>(x[q], slope);
                  break;
                case SVX_ACT_GELU:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = gelu_erf(x[q]);
                  break;
                default: break;
              }
              if (post) {
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q] += rv[q];
              }
            } else {
              switch (p.act) {
                case SVX_ACT_RELU:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = fmaxf(x[q], 0.f);
                  break;
                case SVX_ACT_LEAKY:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = act_t<SVX_ACT_LEAKY>(x[q], slope);
                  break;
                case SVX_ACT_GELU:
#pragma unroll
                  for (int q = 0; q < 16; ++q) x[q] = gelu_erf(x[q]);
                  break;
                default: break;
              }
            }
            if (scale != 1.f) {
#pragma unroll
              for (int q = 0; q < 16; ++q) x[q] *= scale;
            }
            if (p.round_tf32) {
#pragma unroll
              for (int q = 0; q < 16; ++q) x[q] = round_tf32(x[q]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
          } else {
            // few output channels (layer6: one) or unaligned destinations: scalar path over the first N columns
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              if (q < p.N) {
                float tv = x[q];
                const float rv = has_res ? p.residual[off + q] : 0.f;
                if (pre) tv += rv;
                tv = apply_act(tv, p.act, slope);
                if (post) tv += rv;
                tv *= scale;
                dst[q] = p.round_tf32 ? round_tf32(tv) : tv;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 5) { if (PAIR) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch `pitch_elems`; box = box_rows x box_cols columns (32: 128B swizzle,
// 16: 64B swizzle)
// esize = 4: fp32 elements; 2: bf16.  The swizzle mode follows the box width in BYTES (128 / 64 / 32).
int encode_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
               uint32_t box_rows, uint32_t box_cols = BK, int esize = 4) {
  if (box_rows > 256) return fail("TMA box of %u rows exceeds 256", box_rows);
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * (uint64_t)esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const uint32_t box_bytes = box_cols * (uint32_t)esize;
  if (box_bytes != 128 && box_bytes != 64 && box_bytes != 32) return fail("TMA box of %u bytes per row unsupported", box_bytes);
  if ((pitch_elems * (uint64_t)esize) % 16 != 0) return fail("TMA row pitch of %llu bytes is not a multiple of 16", (unsigned long long)(pitch_elems * esize));
  CUresult r = fn(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  // a 64-byte box out of a wider row must not be promoted to 128-byte L2 fills: that doubled the DRAM
                  // reads of the merger's 16-channel layers (profiles/r1_ncu_merger_v17.txt)
                  box_bytes <= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

// im2col-mode map over the channels-last tensor [N, D, H, W, Cs]: boxes of 128 pixels x 32 channels, 128B swizzle.
// lo / up: bounding-box corners of the base pixel per axis (d, h, w); str: convolution strides (d, h, w).
int encode_im2col_map(CUtensorMap* map, const void* ptr, uint64_t N, uint64_t D, uint64_t H, uint64_t W, uint64_t Cs,
                      const int lo[3], const int up[3], const int str[3], int box_ch = BK, int esize = 4) {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  if (!fn) return fail("cuTensorMapEncodeIm2col entry point not available");
  cuuint64_t dims[5] = {Cs, W, H, D, N};
  const uint64_t es = (uint64_t)esize;
  cuuint64_t strides[4] = {Cs * es, W * Cs * es, H * W * Cs * es, D * H * W * Cs * es};
  const int box_bytes = box_ch * esize;
  int lower[3] = {lo[2], lo[1], lo[0]}, upper[3] = {up[2], up[1], up[0]};   // the driver takes (w, h, d)
  cuuint32_t estr[5] = {1, (cuuint32_t)str[2], (cuuint32_t)str[1], (cuuint32_t)str[0], 1};
  // narrow: one 4-channel (16-byte) pixel per row, stored densely (no swizzle) = a column of 8x16B core matrices
  CUresult r = fn(map, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(ptr),
                  dims, strides, lower, upper, (cuuint32_t)box_ch, BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_bytes == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE : box_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeIm2col failed with CUresult %d (lo %d,%d,%d up %d,%d,%d)", (int)r, lo[0], lo[1], lo[2],
                up[0], up[1], up[2]);
  return 0;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int PARTS>
int launch_bn_parts(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr,
                    const GemmParams& p, int grid, cudaStream_t st, bool bf) {
  using C = Cfg<BN, PARTS>;
  // the opt-in is per device and cheap: set it on every launch path rather than caching a per-process flag
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    SVX_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    SVX_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured_dev = dev;
  }
  static const bool pdl = getenv("SVX_PDL") != nullptr;   // opt-in until validated on the GPU tier
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::kThreadsT);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t le = bf ? cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, PARTS>, ma, mb, mc, mr, p)
                      : cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN, PARTS>, ma, mb, mc, mr, p);
  if (le != cudaSuccess) return fail("launch of gemm_%s_kernel failed: %s", bf ? "bf16" : "tf32", cudaGetErrorString(le));
  SVX_LAUNCH_OK("gemm_tf32_kernel");
  return 0;
}

template <int BN>
int launch_bn(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mr,
              const GemmParams& p, int grid, cudaStream_t st, bool bf) {
  if constexpr (BN >= 128) {
    // erf-GELU epilogues through the TMA-store path: twice the epilogue warps (see Cfg)
    // (only while the main loop is short: at K = 768 the tensor pipe is the bound again and the variant loses;
    //  bf16 main loops take half the time per K, so the epilogue bound holds up to twice the depth)
    if (p.act == SVX_ACT_GELU && p.epi_tma && p.a_mode != SVX_A_GATHER && p.K <= (bf ? 1024 : 512))
      return launch_bn_parts<BN, 4>(ma, mb, mc, mr, p, grid, st, bf);
  }
  return launch_bn_parts<BN, 2>(ma, mb, mc, mr, p, grid, st, bf);
}

}  // namespace

struct GemmPrepared {
  CUtensorMap map_a, map_b, map_c, map_r;
  GemmParams p;
  int bn, grid;
  bool slab = false;   // SVX_A_SLAB3: handled by conv3_slab_kernel
  bool slab_narrow = false;   // 16-channel (64-byte) rows
  bool slab_pair = false;     // CTA pairs (cta_group::2): one M = 256 MMA per step for two units
  bool slab_f16 = false;      // fp16 MMA operands converted inside the kernel (kind::f16: half the MMA instructions)
  bool bf16 = false;          // SVX_OPERAND_BF16: gemm_bf16_kernel
};

int gemm_prepare(const svx_gemm_desc& d, GemmPrepared** out) {
  *out = nullptr;
  SVX_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  SVX_REQUIRE(d.block_n == 16 || d.block_n == 32 || d.block_n == 64 || d.block_n == 96 || d.block_n == 128 ||
                  d.block_n == 192 || d.block_n == 256 || (d.block_n == 48 && d.a_mode == SVX_A_SLAB3),
              "gemm: block_n=%d unsupported", d.block_n);
  const bool bf = d.operand_kind == SVX_OPERAND_BF16;
  const int esz = bf ? 2 : 4;             // bytes per operand element
  const int BKE = 128 / esz;              // elements per 128-byte k-chunk
  const int e16 = 16 / esz;               // elements per 16 bytes
  SVX_REQUIRE(d.operand_kind == SVX_OPERAND_DEFAULT || bf || (d.operand_kind == SVX_OPERAND_TF32 && d.a_mode == SVX_A_SLAB3),
              "gemm: operand_kind %d unsupported for this operand mode", d.operand_kind);
  SVX_REQUIRE(bf || d.io_flags == 0, "gemm: bf16 outputs / residuals need bf16 operands");
  SVX_REQUIRE(!bf || (d.a_mode != SVX_A_SLAB3 && d.epi_mode == SVX_EPI_STD),
              "gemm: bf16 operands support the standard epilogue of the plain / flat / im2col / gather modes only");
  const bool io16 = (d.io_flags & SVX_IO_OUT_BF16) != 0;
  SVX_REQUIRE(!d.residual || ((d.io_flags & SVX_IO_RES_BF16) != 0) == io16 || d.res_via_mma,
              "gemm: the epilogue residual must have the output's storage type");
  SVX_REQUIRE(d.Kpad % BKE == 0 && d.Kpad >= d.K, "gemm: Kpad=%d must be a multiple of %d and >= K=%d", d.Kpad, BKE, d.K);
  SVX_REQUIRE(d.Npad % d.block_n == 0 && d.Npad >= d.N, "gemm: Npad=%d vs N=%d block_n=%d", d.Npad, d.N, d.block_n);
  SVX_REQUIRE(d.A && d.W && d.out, "gemm: null operand");
  SVX_REQUIRE((reinterpret_cast<uintptr_t>(d.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.W) & 15) == 0,
              "gemm: A and W must be 16-byte aligned");
  SVX_REQUIRE(d.out_D > 0 && d.out_H > 0 && d.out_W > 0, "gemm: row decode extents must be positive");
  GemmPrepared* g = new GemmPrepared();
  GemmParams& p = g->p;
  memset(&p, 0, sizeof(p));
  memset(&g->map_a, 0, sizeof(g->map_a));
  memset(&g->map_c, 0, sizeof(g->map_c));
  memset(&g->map_r, 0, sizeof(g->map_r));
  p.chunks_per_tap = 1;
  if (d.a_mode == SVX_A_PLAIN) {
    if (d.lda % e16 != 0 || d.lda < d.K) {
      delete g;
      return fail("gemm: plain A needs a 16-byte row pitch and lda >= K (lda=%lld K=%d)", (long long)d.lda, d.K);
    }
    if (encode_map(&g->map_a, d.A, (uint64_t)d.M, (uint64_t)d.K, (uint64_t)d.lda, BM, BKE, esz)) { delete g; return 1; }
  } else if (d.a_mode == SVX_A_GATHER) {
    bool ok = d.Cin > 0 && d.Cin % e16 == 0 && d.in_c0 % e16 == 0 && d.in_Cs % e16 == 0 && d.ntaps > 0 && d.taps &&
              d.K == d.ntaps * d.Cin && d.in_D > 0 && d.in_H > 0 && d.in_W > 0 && d.in_D < 1024 &&
              d.in_H < 1024 && d.in_W < 1024;
    if (!ok) {
      delete g;
      return fail("gemm: bad gather description (Cin=%d c0=%d Cs=%d ntaps=%d K=%d)", d.Cin, d.in_c0, d.in_Cs,
                  d.ntaps, d.K);
    }
  } else if (d.a_mode == SVX_A_FLAT) {
    bool ok = d.Cin > 0 && d.Cin % BKE == 0 && d.in_c0 % e16 == 0 && d.in_Cs % e16 == 0 && d.ntaps > 0 &&
              d.ntaps <= kMaxTaps && d.taps_host && d.K == d.ntaps * d.Cin && d.Kpad == d.K &&
              d.out_D == d.in_D && d.out_H == d.in_H && d.out_W == d.in_W;
    if (!ok) {
      delete g;
      return fail("gemm: bad flat-conv description (Cin=%d ntaps=%d K=%d Kpad=%d)", d.Cin, d.ntaps, d.K, d.Kpad);
    }
    const long long rows_total = (long long)d.lda;  // flat mode: lda carries the row count of the padded matrix
    if (rows_total < d.M) { delete g; return fail("gemm: flat-conv matrix has %lld rows < M=%d", rows_total, d.M); }
    for (int t = 0; t < d.ntaps; ++t) {
      const int dd = d.taps_host[4 * t], dh = d.taps_host[4 * t + 1], dw = d.taps_host[4 * t + 2];
      if (dd < 0 || dh < 0 || dw < 0) { delete g; return fail("gemm: flat-conv taps must be non-negative"); }
      p.flat_off[t] = (dd * d.in_H + dh) * d.in_W + dw;
    }
    p.chunks_per_tap = d.Cin / BKE;
    if (encode_map(&g->map_a, d.A, (uint64_t)rows_total, (uint64_t)d.in_Cs, (uint64_t)d.in_Cs, BM, BKE, esz)) { delete g; return 1; }
  } else if (d.a_mode == SVX_A_IM2COL) {
    const bool narrow = d.Cin * esz == 16, pairs = d.Cin * esz == 32;   // 16-byte / 32-byte pixels
    bool ok = d.Cin > 0 && (narrow || ((pairs || d.Cin % BKE == 0) && d.Kpad == d.K)) && d.in_c0 % e16 == 0 && d.in_Cs % e16 == 0 &&
              d.ntaps > 0 && d.ntaps <= kMaxTaps && d.taps_host && d.K == d.ntaps * d.Cin && d.in_D > 0 && d.in_H > 0 && d.in_W > 0 &&
              d.stride_d >= 1 && d.stride_h >= 1 && d.stride_w >= 1 && d.stride_d <= 8 && d.stride_h <= 8 && d.stride_w <= 8 &&
              d.M % (d.out_D * d.out_H * d.out_W) == 0;
    if (!ok) {
      delete g;
      return fail("gemm: bad im2col description (Cin=%d c0=%d Cs=%d ntaps=%d K=%d Kpad=%d)", d.Cin, d.in_c0, d.in_Cs,
                  d.ntaps, d.K, d.Kpad);
    }
    int lo[3] = {1 << 20, 1 << 20, 1 << 20}, hi[3] = {-(1 << 20), -(1 << 20), -(1 << 20)}, up[3];
    for (int t = 0; t < d.ntaps; ++t)
      for (int a = 0; a < 3; ++a) {
        lo[a] = d.taps_host[4 * t + a] < lo[a] ? d.taps_host[4 * t + a] : lo[a];
        hi[a] = d.taps_host[4 * t + a] > hi[a] ? d.taps_host[4 * t + a] : hi[a];
      }
    const int in_ext[3] = {d.in_D, d.in_H, d.in_W}, out_ext[3] = {d.out_D, d.out_H, d.out_W};
    const int str[3] = {d.stride_d, d.stride_h, d.stride_w};
    for (int a = 0; a < 3; ++a) {
      // the base pixel of output o sits at lo + o*stride; the box ends exactly at the last output's base pixel
      up[a] = lo[a] + (out_ext[a] - 1) * str[a] - (in_ext[a] - 1);
      if (lo[a] < -16 || lo[a] > 15 || up[a] < -16 || up[a] > 15 || hi[a] - lo[a] > 255) {
        delete g;
        return fail("gemm: im2col corners out of the descriptor range (axis %d: lo=%d up=%d span=%d)", a, lo[a], up[a],
                    hi[a] - lo[a]);
      }
    }
    for (int t = 0; t < d.ntaps; ++t)
      p.flat_off[t] = (d.taps_host[4 * t + 2] - lo[2]) | ((d.taps_host[4 * t + 1] - lo[1]) << 8) |
                      ((d.taps_host[4 * t] - lo[0]) << 16);
    p.i2c_lo_d = lo[0]; p.i2c_lo_h = lo[1]; p.i2c_lo_w = lo[2];
    p.chunks_per_tap = (narrow || pairs) ? 1 : d.Cin / BKE;
    p.i2c_narrow = narrow ? 1 : pairs ? 2 : 0;
    p.ntaps = d.ntaps;
    const uint64_t n_img = (uint64_t)(d.M / (d.out_D * d.out_H * d.out_W));
    if (encode_im2col_map(&g->map_a, d.A, n_img, d.in_D, d.in_H, d.in_W, d.in_Cs, lo, up, str,
                          narrow ? e16 : pairs ? 2 * e16 : BKE, esz)) { delete g; return 1; }
  } else if (d.a_mode == SVX_A_SLAB3) {
    const int live = d.cin_live > 0 ? d.cin_live : BK;
    bool ok = d.Cin == BK && live <= BK && d.in_Cs % 4 == 0 && d.in_c0 >= 0 && d.N <= 16 && d.block_n == S3_N &&
              d.Npad == S3_N && d.K == 9 * BK && d.Kpad == 9 * BK && d.in_D == d.valid_D + 2 && d.valid_H > 0 &&
              d.valid_W > 0 && d.valid_H <= d.in_H && d.valid_W <= d.in_W && BM + 2 * d.in_W <= S3_ROWS &&
              d.epi_mode == SVX_EPI_STD && d.lda > 0;
    if (!ok) {
      delete g;
      return fail("gemm: bad slab-conv description (Cin=%d live=%d N=%d block_n=%d Npad=%d K=%d Kpad=%d in=%dx%dx%d)",
                  d.Cin, live, d.N, d.block_n, d.Npad, d.K, d.Kpad, d.in_D, d.in_H, d.in_W);
    }
    p.cin_live = live;
    g->slab = true;
    g->slab_narrow = live <= 16;
    if (encode_map(&g->map_a, d.A, (uint64_t)d.lda, (uint64_t)d.in_Cs, (uint64_t)d.in_Cs, S3_PART_ROWS, g->slab_narrow ? 16 : BK)) {
      delete g;
      return 1;
    }
  } else {
    delete g;
    return fail("gemm: unknown a_mode %d", d.a_mode);
  }
  const int w_cols = d.Kpad + (d.res_via_mma ? d.block_n : 0);   // identity columns appended by the host
  // CTA pairs (cluster of 2, tcgen05.mma.cta_group::2) are correct (tests pass with SVX_SLAB_PAIR=1) but measured no
  // faster than single CTAs on this kernel (profiles/README.md, "merger slab kernel experiments"): opt-in only.
  g->slab_pair = g->slab && getenv("SVX_SLAB_PAIR") != nullptr;
  g->slab_f16 = g->slab && !g->slab_pair && d.operand_kind != SVX_OPERAND_TF32;
  g->bf16 = bf;
  p.io16 = io16 ? 1 : 0;
  p.Wg = reinterpret_cast<const float*>(d.W);
  p.range_flag = g->slab_f16 ? d.range_flag : nullptr;
  p.acc_scale = d.acc_scale != 0.f ? d.acc_scale : 1.f;
  if (encode_map(&g->map_b, d.W, (uint64_t)d.Npad, (uint64_t)w_cols, (uint64_t)w_cols,
                 g->slab_pair ? (uint32_t)(S3_N / 2) : (uint32_t)d.block_n, g->slab_narrow ? 16 : BKE, esz)) {
    delete g;
    return 1;
  }
  if (d.epi_mode == SVX_EPI_DEC_TAIL &&
      !(d.block_n == 16 && d.N == 16 && d.epi_aux && d.epi_out2)) {
    delete g;
    return fail("gemm: decoder tail epilogue needs block_n=16, N=16, aux weights and a coarse output");
  }
  if (d.epi_mode == SVX_EPI_CONVT8) {
    const bool tail = d.epi_aux != nullptr;
    const bool ok = d.a_mode != SVX_A_SLAB3 && d.cls_cout > 0 && d.N == 8 * d.cls_cout &&
                    (tail ? (d.cls_cout == 8 && d.epi_out2 && !d.residual && ((d.o_base | d.o_sn | d.o_sd | d.o_sh | d.o_sw |
                                                                                 d.c_sd | d.c_sh | d.c_sw) & 3) == 0 &&
                             (reinterpret_cast<uintptr_t>(d.out) & 15) == 0)
                          : true);
    if (!ok) { delete g; return fail("gemm: bad transposed-convolution class epilogue (cls_cout=%d N=%d)", d.cls_cout, d.N); }
  }
  p.M = d.M; p.N = d.N; p.K = d.K; p.Npad = d.Npad; p.nk = d.Kpad / BKE; p.nk_a = p.nk; p.tiles_n = d.Npad / d.block_n; p.a_mode = d.a_mode;
  p.A = reinterpret_cast<const float*>(d.A);
  p.in_D = d.in_D; p.in_H = d.in_H; p.in_W = d.in_W; p.in_Cs = d.in_Cs; p.in_c0 = d.in_c0; p.Cin = d.Cin;
  p.out_D = d.out_D; p.out_H = d.out_H; p.out_W = d.out_W;
  p.valid_D = d.valid_D; p.valid_H = d.valid_H; p.valid_W = d.valid_W;
  if (p.valid_W > 0 && (p.valid_D <= 0 || p.valid_H <= 0)) { delete g; return fail("gemm: valid extents must all be set"); }
  p.sd = d.stride_d; p.sh = d.stride_h; p.sw = d.stride_w;
  p.taps = reinterpret_cast<const int4*>(d.taps);
  p.bias = d.bias; p.residual = reinterpret_cast<const float*>(d.residual); p.out = reinterpret_cast<float*>(d.out);
  p.o_base = d.o_base; p.o_sn = d.o_sn; p.o_sd = d.o_sd; p.o_sh = d.o_sh; p.o_sw = d.o_sw;
  p.act = d.act; p.act_param = d.act_param; p.res_after_act = d.res_after_act;
  p.out_scale = d.out_scale; p.round_tf32 = d.round_tf32; p.epi_mode = d.epi_mode;
  p.epi_aux = d.epi_aux; p.out2 = d.epi_out2;
  p.cls_cout = d.cls_cout; p.c_sd = d.c_sd; p.c_sh = d.c_sh; p.c_sw = d.c_sw; p.c2_sd = d.c2_sd; p.c2_sh = d.c2_sh; p.c2_sw = d.c2_sw;
  p.o2_base = d.o2_base; p.o2_sn = d.o2_sn; p.o2_sd = d.o2_sd; p.o2_sh = d.o2_sh; p.o2_sw = d.o2_sw;
  auto al4 = [](long long v) { return (v & 3) == 0; };
  p.vec_ok = (d.N % 4 == 0) && al4(d.o_base) && al4(d.o_sn) && al4(d.o_sd) && al4(d.o_sh) && al4(d.o_sw) &&
             (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
             (!d.residual || (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0);
  if (bf && !p.vec_ok) { delete g; return fail("gemm: bf16 operands need 4-element aligned outputs (N=%d)", d.N); }
  // Plain row-major output (row r lands at out + o_base + r*ldc, every row a real output): the epilogue stores
  // through TMA.  True for every linear layer and for convolutions writing an unpadded channels-last tensor.
  {
    const bool compact = d.o_sh == (long long)d.out_W * d.o_sw && d.o_sd == (long long)d.out_H * d.o_sh &&
                         d.o_sn == (long long)d.out_D * d.o_sd;
    const bool pool8 = d.epi_mode == SVX_EPI_POOL8;
    const int n_out = pool8 ? d.N / 8 : d.N;
    const int oe16 = io16 ? 8 : 4;   // output elements per 16 bytes
    const bool plain = compact && d.valid_W == 0 && (d.epi_mode == SVX_EPI_STD || pool8) && d.a_mode != SVX_A_SLAB3 &&
                       d.block_n >= 32 && n_out % oe16 == 0 && d.o_sw % oe16 == 0 && d.o_sw >= n_out && (d.o_base % oe16) == 0 &&
                       (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
                       (!d.residual || (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0);
    if (plain) {
      p.epi_tma = 1;
      p.ldc = d.o_sw;
      const size_t osz = io16 ? 2 : 4;
      p.out = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d.out) + (size_t)d.o_base * osz);
      if (d.residual)
        p.residual = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(d.residual) + (size_t)d.o_base * osz);
      // staged rows are 64 bytes either way: 16 fp32 or 32 bf16 columns per bulk store
      if (encode_map(&g->map_c, p.out, (uint64_t)d.M, (uint64_t)n_out, (uint64_t)d.o_sw, 32, io16 ? 32 : 16, (int)osz)) { delete g; return 1; }
    }
    if (d.res_via_mma) {
      // act(A W^T + b + R) with R added by the tensor cores: the k loop runs block_n / 32 extra chunks whose A
      // operand is the residual tile (TMA, coalesced) and whose B operand is an identity block.  Exact when R holds
      // TF32-representable values (the caller's contract).
      const bool ok = plain && d.a_mode == SVX_A_PLAIN && d.residual && !d.res_after_act && d.N % d.block_n == 0 &&
                      d.block_n % BKE == 0 && (!bf || (d.io_flags & SVX_IO_RES_BF16));
      if (!ok) { delete g; return fail("gemm: res_via_mma needs a plain operand, a plain output, a pre-activation residual and N %% block_n == 0"); }
      if (encode_map(&g->map_r, p.residual, (uint64_t)d.M, (uint64_t)d.N, (uint64_t)d.o_sw, BM, BKE, esz)) { delete g; return 1; }
      p.nk = p.nk_a + d.block_n / BKE;
      p.residual = nullptr;   // nothing left for the epilogue to add
    }
    if (pool8 && !(plain && d.N == d.block_n && d.N % 128 == 0 && !d.residual)) {
      delete g;
      return fail("gemm: the pooled epilogue needs a plain [M, N/8] output, one N tile (block_n == N) and no residual");
    }
  }
  const long long tiles_m = (d.M + BM - 1) / BM;
  const long long tiles = tiles_m * p.tiles_n;
  if (tiles > 0x7fffffffLL) { delete g; return fail("gemm: too many tiles"); }
  p.tiles_m = (int)tiles_m;
  g->bn = d.block_n;
  const int slots = sm_count() * (d.block_n <= 96 ? 2 : 1);
  g->grid = (int)(tiles < slots ? tiles : slots);
  if (g->slab) {   // units = volumes x 126-row columns of the (h,w) plane
    const int rows_plane = (d.valid_H - 1) * d.in_W + d.valid_W;
    p.tiles_n = (rows_plane + S3_STEP - 1) / S3_STEP;
    p.tiles_m = (d.M / (d.valid_D * d.valid_H * d.valid_W)) * p.tiles_n;
    p.nk = 1;
    g->grid = p.tiles_m < sm_count() ? p.tiles_m : sm_count();
    if (g->slab_pair) {   // whole pairs only; a pair handles two units per step
      const int steps = (p.tiles_m + 1) / 2;
      const int pairs = steps < sm_count() / 2 ? steps : sm_count() / 2;
      g->grid = 2 * pairs;
    }
  }
  *out = g;
  return 0;
}

void gemm_prepared_free(GemmPrepared* p) { delete p; }

int gemm_launch(const svx_gemm_desc& d, GemmPrepared* prepared, void* stream) {
  GemmPrepared* g = prepared;
  if (!g) {
    if (int rc = gemm_prepare(d, &g)) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = 0;
  if (g->slab) {
    static thread_local int configured_dev = -1;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (configured_dev != cur_dev) {
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<128>::kSmem));
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<64>::kSmem));
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<128>::kSmem));
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<64>::kSmem));
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<128, true>::kSmem));
      SVX_CUDA_OK(cudaFuncSetAttribute(conv3_slab_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S3Cfg<64, true>::kSmem));
      configured_dev = cur_dev;
    }
    if (g->slab_pair) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(g->grid);
      cfg.blockDim = dim3(S3_THREADS);
      cfg.dynamicSmemBytes = g->slab_narrow ? S3Cfg<64>::kSmem : S3Cfg<128>::kSmem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t le = g->slab_narrow ? cudaLaunchKernelEx(&cfg, conv3_slab_kernel<64, true, false>, g->map_a, g->map_b, g->p)
                                      : cudaLaunchKernelEx(&cfg, conv3_slab_kernel<128, true, false>, g->map_a, g->map_b, g->p);
      if (le != cudaSuccess) { if (!prepared) delete g; return fail("cluster launch of conv3_slab_kernel failed: %s", cudaGetErrorString(le)); }
    } else if (g->slab_f16) {
      if (g->slab_narrow)
        conv3_slab_kernel<64, false, true><<<g->grid, S3_THREADS_F16, S3Cfg<64, true>::kSmem, st>>>(g->map_a, g->map_b, g->p);
      else
        conv3_slab_kernel<128, false, true><<<g->grid, S3_THREADS_F16, S3Cfg<128, true>::kSmem, st>>>(g->map_a, g->map_b, g->p);
    } else if (g->slab_narrow) {
      conv3_slab_kernel<64, false, false><<<g->grid, S3_THREADS, S3Cfg<64>::kSmem, st>>>(g->map_a, g->map_b, g->p);
    } else {
      conv3_slab_kernel<128, false, false><<<g->grid, S3_THREADS, S3Cfg<128>::kSmem, st>>>(g->map_a, g->map_b, g->p);
    }
    cudaError_t e = cudaGetLastError();
    if (!prepared) delete g;
    if (e != cudaSuccess) return fail("launch of conv3_slab_kernel failed: %s", cudaGetErrorString(e));
    return 0;
  }
  switch (g->bn) {
    case 16: rc = launch_bn<16>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    case 32: rc = launch_bn<32>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    case 64: rc = launch_bn<64>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    case 96: rc = launch_bn<96>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    case 128: rc = launch_bn<128>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    case 192: rc = launch_bn<192>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
    default: rc = launch_bn<256>(g->map_a, g->map_b, g->map_c, g->map_r, g->p, g->grid, st, g->bf16); break;
  }
  if (!prepared) delete g;
  return rc;
}

int gemm_num_launches(const svx_gemm_desc&) { return 1; }

}  // namespace svx

// svx_mlp.cu -- the MLP of a Swin block (timm Mlp: fc1 -> nn.GELU() -> fc2) plus the block's second residual as ONE
// persistent kernel on sm_100a (svx_mlp_desc).  Unfused, the 4C-wide hidden activation is written to and re-read from
// HBM (stage 0 at 64 x 3 views: 0.93 GB each way per block, 4/5 of the traffic of fc1 + fc2); here it never leaves the SM.
//
// Per 128-row tile, hidden chunk by hidden chunk (HC columns):
//   G1(c): acc1[b] (TMEM, 128 x HC)  = X (smem, resident for the tile) . W1[c]^T       tcgen05.mma kind::tf32
//   E(c):  H (smem ring, 128B-swizzled K-major operand tiles) = round_tf32(gelu(acc1[b] + b1[c]))   16 epilogue warps
//   G2(c): acc2 (TMEM, 128 x C)     += H . W2[:, c]^T
// and after the last chunk out = acc2 + b2 + residual (the tile tail: staged through shared memory by TMA for C = 96,
// transposed to 64-byte row segments for C = 192).  With ln_gamma set (C = 96) the epilogue warps first apply the block's
// norm2 to the TMA-delivered X tile in place, so the kernel reads the un-normalised x1 once and no norm2 kernel runs.
// The single MMA-issuing thread runs the software pipeline G1(0) G1(1) | G2(n) G1(n+2) ... over the chunk stream of ALL
// its tiles, so the tensor pipe works on chunk n+1 / n+2 while the epilogue warps are in the GELU of chunk n, also across
// tile boundaries.
//
//   warps 0-15  epilogue: TMEM lane quarter = warp % 4, column part = warp / 4
//   warp 16     TMA producer of the W1 / W2 k-chunks (a ring in exactly the order the MMA thread consumes them)
//   warp 17     MMA issuer (one elected thread), owns the TMEM allocation
//   warp 18     TMA producer of the X tiles
#include <cuda.h>
#include <cuda_runtime.h>

#include "svx_internal.h"
#include "svx_ptx.cuh"
#include "svx_act.cuh"

namespace svx {
namespace {

constexpr int ML_BM = 128;                       // rows per tile
constexpr int ML_BK = 32;                        // fp32 per k-chunk = one 128-byte swizzle row
constexpr int ML_KCH = ML_BM * ML_BK * 4;        // one 128-row k-chunk of an A operand: 16 KB
constexpr int ML_EPI_WARPS = 16;
constexpr int ML_WARP_W = 16, ML_WARP_MMA = 17, ML_WARP_X = 18;
constexpr int ML_THREADS = 19 * 32;
constexpr int ML_SMEM_MAX = 232448;

template <int C, int HC>
struct MlpCfg {
  static_assert(C % 32 == 0 && HC % 64 == 0 && HC <= 128 && (4 * C) % HC == 0, "unsupported MLP shape");
  static constexpr int kHidden = 4 * C;
  static constexpr int kNC = kHidden / HC;                 // hidden chunks per tile
  static constexpr int kXChunks = C / ML_BK;
  static constexpr int kXBytes = kXChunks * ML_KCH;
  static constexpr int kXBufs = C <= 96 ? 2 : 1;
  // With two X buffers the residual / output tile goes through shared memory (TMA load, in-place update, TMA store):
  // per-lane rows of a [128, C] tile are 32 cache lines per warp-wide access, which made the tile tail the longest
  // phase.  The tile borrows the X buffer of its own tile: X(i) is dead after the last fc1 MMA of tile i, the tail
  // comes two chunks later, and X(i+2) is not needed before the end of tile i+1.
  static constexpr bool kStaged = kXBufs == 2;
  static constexpr int kKCh = HC / ML_BK;                  // k-chunks of one hidden chunk
  // H ring slots: one per k-chunk of a hidden chunk while shared memory allows (C = 96); C = 192 has room for two, so
  // with HC = 128 a slot is used twice per chunk (k-chunks s and s + 2) and has its own "consumed" barrier
  static constexpr int kNHS = C <= 96 ? kKCh : 2;
  static constexpr int kUPS = kKCh / kNHS;                 // uses of one slot per hidden chunk
  static_assert(kKCh % kNHS == 0 && (kUPS == 1 || kUPS == 2), "H ring shape");
  static constexpr int kSPP = HC / 64;                     // 16-column slabs per epilogue warp and chunk
  static constexpr int kHFullCount = 4 * (2 / kSPP);       // warps that write one H k-chunk
  static constexpr int kWRows = C > HC ? C : HC;
  static constexpr int kWStage = kWRows * ML_BK * 4;       // holds W1 k-chunks (HC rows each) or one W2 k-chunk (C rows)
  static constexpr int kW1PerStageRaw = kWStage / (HC * ML_BK * 4);
  // W1 k-chunks per ring stage: as many as fit in the slot and divide the chunk count (C = 192, HC = 64: 3 x 8 KB)
  static constexpr int kW1PerStage = (kW1PerStageRaw >= 3 && kXChunks % 3 == 0) ? 3 : (kW1PerStageRaw >= 2 && kXChunks % 2 == 0) ? 2 : 1;
  static constexpr int kW1Stages = kXChunks / kW1PerStage; // ring stages one fc1 chunk consumes
  static constexpr int kFixed = 1024 + 512 + kXBufs * kXBytes + kNHS * ML_KCH;
  static constexpr int kWStagesRaw = (ML_SMEM_MAX - kFixed) / kWStage;
  static constexpr int kWStages = kWStagesRaw > 8 ? 8 : kWStagesRaw;
  static_assert(kWStages >= 3, "weight ring too shallow");
  static constexpr int kSmem = kFixed + kWStages * kWStage;
  static constexpr uint32_t kAcc2Col = 2 * HC;             // TMEM: acc1[0], acc1[1], acc2[0] (, acc2[1])
  static constexpr int kAcc2Bufs = (2 * HC + 2 * C <= 512) ? 2 : 1;
  static_assert(2 * HC + kAcc2Bufs * C <= 512, "TMEM columns");
  static constexpr int kOutSlabs = C / 16;
  static constexpr int kMaxOutSlabs = (kOutSlabs + 3) / 4;  // per epilogue warp
};

#ifdef SVX_MLP_PROFILE   // role timers of CTA 0 (tools/probes/mlp_time.py): cycles spent in each kind of wait
__device__ unsigned long long g_mlp_prof[20];
#define ML_TIMED_WAIT(slot, ...) do { const long long t0__ = clock64(); __VA_ARGS__; prof[slot] += clock64() - t0__; } while (0)
#define ML_MARK(var) const long long var = clock64()
#define ML_SPAN(slot, a, b) prof[slot] += (b) - (a)
#else
#define ML_TIMED_WAIT(slot, ...) do { __VA_ARGS__; } while (0)
#define ML_MARK(var)
#define ML_SPAN(slot, a, b)
#endif

struct MlpParams {
  int M, tiles;
  const float* b1;
  const float* b2;
  const float* residual;
  float* out;
  long long ldo;
  int round_tf32;
  const float* ln_gamma;   // fused pre-LayerNorm (staged variant only): x is the un-normalised input (= the residual)
  const float* ln_beta;
  float ln_eps;
};

template <int C, int HC>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ CUtensorMap map_o, const __grid_constant__ MlpParams p) {
  using K = MlpCfg<C, HC>;
  constexpr int WS = K::kWStages, NC = K::kNC, XB = K::kXBufs, NHS = K::kNHS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t x_smem = smem_base;
  const uint32_t h_smem = x_smem + XB * K::kXBytes;
  const uint32_t w_smem = h_smem + NHS * ML_KCH;
  const uint32_t bar_base = w_smem + WS * K::kWStage;
  uint8_t* h_gen = smem_gen + XB * K::kXBytes;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (WS + s); };
  auto x_full = [&](int b) { return bar_base + 8u * (2 * WS + b); };
  auto x_empty = [&](int b) { return bar_base + 8u * (2 * WS + 2 + b); };
  auto acc1_full = [&](int b) { return bar_base + 8u * (2 * WS + 4 + b); };
  auto acc1_empty = [&](int b) { return bar_base + 8u * (2 * WS + 6 + b); };
  auto h_full = [&](int j) { return bar_base + 8u * (2 * WS + 8 + j); };
  const uint32_t h_empty = bar_base + 8u * (2 * WS + 12);
  auto acc2_full = [&](int b) { return bar_base + 8u * (2 * WS + 13 + b); };
  auto acc2_empty = [&](int b) { return bar_base + 8u * (2 * WS + 15 + b); };
  const uint32_t r_full = bar_base + 8u * (2 * WS + 17);
  const uint32_t r_done = bar_base + 8u * (2 * WS + 18);
  auto xn_full = [&](int b) { return bar_base + 8u * (2 * WS + 19 + b); };   // X tile normalised in place (fused LayerNorm)
  // kUPS == 2: one "consumed" barrier per (slot, use within the chunk).  A single barrier per slot would make the
  // writers of the second use wait TWO phases ahead, which a parity wait cannot tell from zero phases ahead.
  auto h_empty2 = [&](int slot, int use) { return bar_base + 8u * (2 * WS + 21 + slot * 2 + use); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * WS + 25);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (bar_base - smem_base) + 8 * (2 * WS + 25));
  const bool fuse_ln = K::kStaged && p.ln_gamma != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = ((int)blockIdx.x < p.tiles) ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ntot = nt * NC;   // hidden chunks this CTA works through

  if (warp == ML_WARP_W && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_w2);
    if (K::kStaged) { tma_prefetch_desc(&map_r); tma_prefetch_desc(&map_o); }
  }
  if (warp == ML_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < WS; ++s) { mbar_init(w_full(s), 1u); mbar_init(w_empty(s), 1u); }
      for (int b = 0; b < 2; ++b) {
        mbar_init(x_full(b), 1u); mbar_init(x_empty(b), 1u);
        mbar_init(acc1_full(b), 1u); mbar_init(acc1_empty(b), ML_EPI_WARPS);
        mbar_init(acc2_full(b), 1u); mbar_init(acc2_empty(b), ML_EPI_WARPS);
      }
      for (int j = 0; j < 4; ++j) mbar_init(h_full(j), K::kHFullCount);
      mbar_init(h_empty, 1u);
      for (int s2 = 0; s2 < 2; ++s2) { mbar_init(h_empty2(s2, 0), 1u); mbar_init(h_empty2(s2, 1), 1u); }
      mbar_init(r_full, 1u);
      mbar_init(r_done, ML_EPI_WARPS);
      mbar_init(xn_full(0), ML_EPI_WARPS);
      mbar_init(xn_full(1), ML_EPI_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();               // (see svx_gemm.cu: the preamble above overlapped the previous kernel's tail)
  pdl_launch_dependents();

  if (warp == ML_WARP_W) {
    // ---- weight producer: the ring carries W1 / W2 k-chunks in the MMA thread's order -----------------------
    if (elect_one()) {
      uint32_t g = 0;
      auto slot = [&](uint32_t bytes) {
        const int s = g % WS;
        mbar_wait(w_empty(s), ((g / WS) & 1u) ^ 1u);
#ifdef SVX_MLP_NOWLOAD   // timing probe: every ring slot is loaded once, later uses only hand the slot over
        if (g >= (uint32_t)WS) { mbar_arrive(w_full(s)); ++g; return -1; }
#endif
        mbar_arrive_expect_tx(w_full(s), bytes);
        ++g;
        return s;
      };
      auto load_g1 = [&](int n) {
        const int c = n % NC;
#pragma unroll 1
        for (int ks = 0; ks < K::kW1Stages; ++ks) {
          const int s = slot(K::kW1PerStage * HC * ML_BK * 4);
          if (s >= 0) {
#pragma unroll
            for (int u = 0; u < K::kW1PerStage; ++u)
              tma_load_2d(w_smem + s * K::kWStage + u * (HC * ML_BK * 4), &map_w1, w_full(s),
                          (ks * K::kW1PerStage + u) * ML_BK, c * HC);
          }
        }
      };
      auto load_g2 = [&](int n) {
        const int c = n % NC;
#pragma unroll 1
        for (int j = 0; j < K::kKCh; ++j) {
          const int s = slot(C * ML_BK * 4);
          if (s >= 0) tma_load_2d(w_smem + s * K::kWStage, &map_w2, w_full(s), c * HC + j * ML_BK, 0);
        }
      };
      if (ntot > 0) load_g1(0);
      if (ntot > 1) load_g1(1);
      for (int n = 0; n < ntot; ++n) {
        load_g2(n);
        if (n + 2 < ntot) load_g1(n + 2);
      }
    }
    __syncwarp();
  } else if (warp == ML_WARP_X) {
    // ---- X tile producer -------------------------------------------------------------------------------
    if (elect_one()) {
      auto load_x = [&](int i, bool wait_empty) {
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * ML_BM;
        const int xb = i % XB;
        if (wait_empty) mbar_wait(x_empty(xb), (((uint32_t)(i / XB)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(x_full(xb), K::kXBytes);
#pragma unroll 1
        for (int kc = 0; kc < K::kXChunks; ++kc)
          tma_load_2d(x_smem + xb * K::kXBytes + kc * ML_KCH, &map_x, x_full(xb), kc * ML_BK, m0);
      };
      if constexpr (!K::kStaged) {
        for (int i = 0; i < nt; ++i) load_x(i, true);
      } else {
        // buffer i % 2 in turn holds X(i) -> [last fc1 MMA of tile i] -> the residual tile R(i) -> [epilogue warps
        // update it in place] -> it is stored as the output tile -> X(i+2)
        if (nt > 0) load_x(0, false);
        if (nt > 1) load_x(1, false);
        for (int i = 0; i < nt; ++i) {
          const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * ML_BM;
          const uint32_t buf = x_smem + (i & 1) * K::kXBytes;
          mbar_wait(x_empty(i & 1), ((uint32_t)i >> 1) & 1u);   // every fc1 MMA of tile i has read X(i)
          mbar_arrive_expect_tx(r_full, K::kXBytes);
#pragma unroll 1
          for (int kc = 0; kc < K::kXChunks; ++kc) tma_load_2d(buf + kc * ML_KCH, &map_r, r_full, kc * ML_BK, m0);
          mbar_wait(r_done, (uint32_t)i & 1u);       // every epilogue warp has written its part of the output tile
#pragma unroll 1
          for (int kc = 0; kc < K::kXChunks; ++kc) tma_store_2d(&map_o, buf + kc * ML_KCH, kc * ML_BK, m0);
          tma_store_commit();
          tma_store_wait_read<0>();
          if (i + 2 < nt) load_x(i + 2, false);
        }
        tma_store_wait_all<0>();
      }
    }
    __syncwarp();
  } else if (warp == ML_WARP_MMA) {
    // ---- MMA issuer --------------------------------------------------------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc_tf32(ML_BM, HC);
      constexpr uint32_t idesc2 = umma_idesc_tf32(ML_BM, C);
      uint32_t g = 0;
#ifdef SVX_MLP_PROFILE
      long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_start = clock64();
#endif
      auto g1 = [&](int n) {
        const int i = n / NC, c = n - i * NC, xb = i % XB, b = n & 1;
        if (c == 0) ML_TIMED_WAIT(0, mbar_wait(fuse_ln ? xn_full(xb) : x_full(xb), ((uint32_t)(i / XB)) & 1u));
        ML_TIMED_WAIT(1, mbar_wait(acc1_empty(b), (((uint32_t)n >> 1) & 1u) ^ 1u));
        tc_fence_after();
        const uint32_t acc = tmem_base + b * HC;
#pragma unroll 1
        for (int ks = 0; ks < K::kW1Stages; ++ks, ++g) {
          const int s = g % WS;
          ML_TIMED_WAIT(2, mbar_wait(w_full(s), (g / WS) & 1u));
          tc_fence_after();
#pragma unroll
          for (int u = 0; u < K::kW1PerStage; ++u) {
            const int kc = ks * K::kW1PerStage + u;
            const uint64_t da = umma_desc_sw128(x_smem + xb * K::kXBytes + kc * ML_KCH);
            const uint64_t db = umma_desc_sw128(w_smem + s * K::kWStage + u * (HC * ML_BK * 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32(acc, da + 2u * k, db + 2u * k, idesc1, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(w_empty(s));
        }
        umma_commit(acc1_full(b));
        if (c == NC - 1) umma_commit(x_empty(xb));   // every fc1 MMA of the tile has read X
      };
      auto g2 = [&](int n) {
        const int i = n / NC, c = n - i * NC;
        const int ab = K::kAcc2Bufs == 2 ? (i & 1) : 0;
        if (c == 0) {
          ML_TIMED_WAIT(3, mbar_wait(acc2_empty(ab), ((K::kAcc2Bufs == 2 ? (uint32_t)i >> 1 : (uint32_t)i) & 1u) ^ 1u));
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + K::kAcc2Col + ab * C;
#pragma unroll 1
        for (int j = 0; j < K::kKCh; ++j, ++g) {
          const int hs = j % NHS;                                          // H ring slot of this k-chunk
          const uint32_t hu = (uint32_t)n * K::kUPS + (uint32_t)(j / NHS);   // its use count
          ML_TIMED_WAIT(4, mbar_wait(h_full(hs), hu & 1u));
          const int s = g % WS;
          ML_TIMED_WAIT(5, mbar_wait(w_full(s), (g / WS) & 1u));
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(h_smem + hs * ML_KCH);
          const uint64_t db = umma_desc_sw128(w_smem + s * K::kWStage);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(acc, da + 2u * k, db + 2u * k, idesc2, (c | j | k) != 0 ? 1u : 0u);
          umma_commit(w_empty(s));
          if (K::kUPS == 2) umma_commit(h_empty2(hs, j / NHS));   // this slot may take its next k-chunk
        }
        if (K::kUPS == 1) umma_commit(h_empty);        // every fc2 MMA of this chunk has read H
        if (c == NC - 1) umma_commit(acc2_full(ab));
      };
      if (ntot > 0) g1(0);
      if (ntot > 1) g1(1);
      for (int n = 0; n < ntot; ++n) {
        g2(n);
        if (n + 2 < ntot) g1(n + 2);
      }
#ifdef SVX_MLP_PROFILE
      if (blockIdx.x == 0) {
        for (int k = 0; k < 6; ++k) g_mlp_prof[k] = (unsigned long long)prof[k];
        g_mlp_prof[6] = (unsigned long long)(clock64() - t_start);
        g_mlp_prof[7] = (unsigned long long)nt;
      }
#endif
    }
    __syncwarp();
  } else {
    // ---- epilogue warps -----------------------------------------------------------------------------------
    const int q = warp & 3, part = warp >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int r = q * 32 + lane;                       // row inside the tile
    const int sw = r & 7;                              // 128B swizzle phase of the row
    const int sl0 = part * K::kSPP;                    // first 16-column slab of a hidden chunk this warp owns
    const int j = sl0 >> 1;                            // ... which lies in H k-chunk j
    const int hs = j % NHS;                            // ... staged in ring slot hs
    uint8_t* h_row = h_gen + hs * ML_KCH + r * 128;
#ifdef SVX_MLP_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = clock64();
#endif
    // Fused pre-LayerNorm (norm2 of the Swin block): the X tile of tile i2 is normalised in place once TMA has
    // delivered it -- four threads per row (a row = C/32 k-chunks x 8 swizzled 16-byte units), two-pass statistics in
    // registers, TF32-rounded result -- and handed to the MMA thread through xn_full.
    auto layernorm_tile = [&](int i2) {
      if constexpr (K::kStaged) {
        const int xb = i2 % XB;
        mbar_wait(x_full(xb), ((uint32_t)(i2 / XB)) & 1u);
        const int row = threadIdx.x >> 2, sub = threadIdx.x & 3;
        uint8_t* xrow = smem_gen + xb * K::kXBytes + row * 128;
        float4 v[K::kXChunks][2];
        float sum = 0.f;
#pragma unroll
        for (int kc = 0; kc < K::kXChunks; ++kc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            v[kc][h] = *reinterpret_cast<const float4*>(xrow + kc * ML_KCH + (((sub + 4 * h) ^ (row & 7)) << 4));
            sum += (v[kc][h].x + v[kc][h].y) + (v[kc][h].z + v[kc][h].w);
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float mean = sum * (1.f / (float)C);
        float sq = 0.f;
#pragma unroll
        for (int kc = 0; kc < K::kXChunks; ++kc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float a = v[kc][h].x - mean, b = v[kc][h].y - mean, e = v[kc][h].z - mean, f = v[kc][h].w - mean;
            sq += (a * a + b * b) + (e * e + f * f);
          }
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        const float rstd = rsqrtf(sq * (1.f / (float)C) + p.ln_eps);
#pragma unroll
        for (int kc = 0; kc < K::kXChunks; ++kc)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = kc * ML_BK + (sub + 4 * h) * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + ch));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(p.ln_beta + ch));
            float4 o;
            o.x = round_tf32((v[kc][h].x - mean) * rstd * g.x + bt.x);
            o.y = round_tf32((v[kc][h].y - mean) * rstd * g.y + bt.y);
            o.z = round_tf32((v[kc][h].z - mean) * rstd * g.z + bt.z);
            o.w = round_tf32((v[kc][h].w - mean) * rstd * g.w + bt.w);
            *reinterpret_cast<float4*>(xrow + kc * ML_KCH + (((sub + 4 * h) ^ (row & 7)) << 4)) = o;
          }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(xn_full(xb));
      }
    };
    if (fuse_ln && nt > 0) layernorm_tile(0);
    for (int i = 0; i < nt; ++i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * ML_BM;
      if constexpr (!K::kStaged) {
        // pull the tile's residual rows into L2 now; the tail reads them two dozen microseconds later
        constexpr int kLines = C / 32;   // 128-byte lines per row
        for (int idx = threadIdx.x; idx < ML_BM * kLines; idx += ML_EPI_WARPS * 32) {
          const int row = m0 + idx / kLines;
          if (row < p.M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + static_cast<long long>(row) * p.ldo + (idx % kLines) * 32));
        }
      }
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
        const int n = i * NC + c, b = n & 1;
        ML_TIMED_WAIT(0, mbar_wait(acc1_full(b), ((uint32_t)n >> 1) & 1u));
        tc_fence_after();
        ML_MARK(ta);
        uint32_t v[K::kSPP][16];
#pragma unroll
        for (int si = 0; si < K::kSPP; ++si) tmem_ld16(lane_addr + b * HC + (sl0 + si) * 16, v[si]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc1_empty(b));     // the MMA thread may overwrite this accumulator
        ML_MARK(tb);
        ML_SPAN(4, ta, tb);
        float x[K::kSPP][16];
#pragma unroll
        for (int si = 0; si < K::kSPP; ++si) {
          const float4* bp = reinterpret_cast<const float4*>(p.b1 + c * HC + (sl0 + si) * 16);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 bv = __ldg(bp + cc);
            x[si][4 * cc + 0] = __uint_as_float(v[si][4 * cc + 0]) + bv.x;
            x[si][4 * cc + 1] = __uint_as_float(v[si][4 * cc + 1]) + bv.y;
            x[si][4 * cc + 2] = __uint_as_float(v[si][4 * cc + 2]) + bv.z;
            x[si][4 * cc + 3] = __uint_as_float(v[si][4 * cc + 3]) + bv.w;
          }
#ifndef SVX_MLP_NOGELU   // (timing probe)
#pragma unroll
          for (int e = 0; e < 16; e += 2) gelu_erf2_fast(x[si][e], x[si][e + 1]);
#endif
          // H is only ever read by tcgen05.mma kind::tf32, which drops the low 13 mantissa bits: adding half an ulp
          // of TF32 here makes that truncation a round-to-nearest (what cvt.rna + a store would give), one IADD
#pragma unroll
          for (int e = 0; e < 16; ++e) x[si][e] = __uint_as_float(__float_as_uint(x[si][e]) + 0x1000u);
        }
        ML_MARK(tc);
        ML_SPAN(5, tb, tc);
        // the fc2 MMAs that read the slot's previous contents have completed
        if (K::kUPS == 1) ML_TIMED_WAIT(1, mbar_wait(h_empty, ((uint32_t)n & 1u) ^ 1u));
        else if (j / NHS == 0) ML_TIMED_WAIT(1, mbar_wait(h_empty2(hs, 1), ((uint32_t)n & 1u) ^ 1u));   // previous chunk's 2nd use
        else ML_TIMED_WAIT(1, mbar_wait(h_empty2(hs, 0), (uint32_t)n & 1u));                           // this chunk's 1st use
        ML_MARK(td);
#pragma unroll
        for (int si = 0; si < K::kSPP; ++si) {
          const int half = (sl0 + si) & 1;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<float4*>(h_row + (((half * 4 + cc) ^ sw) << 4)) =
                make_float4(x[si][4 * cc], x[si][4 * cc + 1], x[si][4 * cc + 2], x[si][4 * cc + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full(hs));
        ML_MARK(te);
        ML_SPAN(6, td, te);
        // after chunk 1: the next tile's X (requested when the previous output tile had left) has had two chunks to
        // arrive -- normalising it after chunk 0 stalled ~1.5k cycles per tile on x_full -- and its first fc1 MMA, issued
        // behind chunk 1's fc2, is not needed before this tile's tail is done
        if (c == (NC > 1 ? 1 : 0) && fuse_ln && i + 1 < nt) layernorm_tile(i + 1);
      }
      ML_MARK(tf);
      // ---- tile tail: out = acc2 + b2 + residual ------------------------------------------------------------
      const int ab = K::kAcc2Bufs == 2 ? (i & 1) : 0;
      const uint32_t acc2_par = (K::kAcc2Bufs == 2 ? (uint32_t)i >> 1 : (uint32_t)i) & 1u;
      const uint32_t acc2_addr = lane_addr + K::kAcc2Col + ab * C;
      if constexpr (K::kStaged) {
        // the residual tile sits in shared memory in the X layout (128B-swizzled k-chunks, loaded by TMA); every
        // thread updates the 16-byte units of its own row in place, then the tile leaves through TMA stores
        ML_TIMED_WAIT(2, mbar_wait(r_full, (uint32_t)i & 1u));
        ML_TIMED_WAIT(3, mbar_wait(acc2_full(ab), acc2_par));
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < K::kMaxOutSlabs; ++t) {
          const int sl = part + 4 * t;
          if (sl < K::kOutSlabs) {   // warp-uniform
            uint32_t v[16];
            tmem_ld16(acc2_addr + sl * 16, v);
            uint8_t* rrow = smem_gen + (i & 1) * K::kXBytes + (sl >> 1) * ML_KCH + r * 128;
            const float4* bp = reinterpret_cast<const float4*>(p.b2 + sl * 16);
            float4 rv[4], bv[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              rv[cc] = *reinterpret_cast<const float4*>(rrow + ((((sl & 1) * 4 + cc) ^ sw) << 4));
              bv[cc] = __ldg(bp + cc);
            }
            tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              float4 o;
              o.x = __uint_as_float(v[4 * cc + 0]) + bv[cc].x + rv[cc].x;
              o.y = __uint_as_float(v[4 * cc + 1]) + bv[cc].y + rv[cc].y;
              o.z = __uint_as_float(v[4 * cc + 2]) + bv[cc].z + rv[cc].z;
              o.w = __uint_as_float(v[4 * cc + 3]) + bv[cc].w + rv[cc].w;
              if (p.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
              *reinterpret_cast<float4*>(rrow + ((((sl & 1) * 4 + cc) ^ sw) << 4)) = o;
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(acc2_empty(ab));
          mbar_arrive(r_done);
        }
      } else {
        // No spare tile buffer (C = 192): each warp transposes its 32 x 16 accumulator blocks through 2 KB of the H
        // ring so that every warp-wide global access covers 8 rows x 64 contiguous bytes instead of 32 rows x 16
        // (the row-per-lane version spent a third of the tile here).  The 2 KB lie inside the H region this warp and its
        // partner (same lane quarter, other slab) write during the GELU phase: H is dead between the last fc2 MMA of
        // the tile (acc2_full) and the pair's next H write, which the pair barrier below orders behind both tails.
        // slot / half of this warp's carve-out: the two warps of a lane quarter that write the same slot during the GELU
        // phase (HC = 64: parts 2s, 2s+1; HC = 128: parts s, s+2) split that slot's 32 rows between them
        const int tslot = K::kSPP == 1 ? (part >> 1) : (part & 1), thalf = K::kSPP == 1 ? (part & 1) : (part >> 1);
        float* stage = reinterpret_cast<float*>(h_gen + tslot * ML_KCH + (q * 32 + thalf * 16) * 128);
        const int trow = lane >> 2, tch = lane & 3;    // read-back mapping: 8 rows x 4 float4 per pass
        // the residual does not depend on the accumulator: every block's share is requested before the wait (the lines
        // were pulled into L2 at the start of the tile)
        float4 rv[K::kMaxOutSlabs][4];
#pragma unroll
        for (int t = 0; t < K::kMaxOutSlabs; ++t) {
          const int sl = part + 4 * t;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int row = m0 + q * 32 + g4 * 8 + trow;
            rv[t][g4] = (row < p.M && sl < K::kOutSlabs)
                            ? *reinterpret_cast<const float4*>(p.residual + static_cast<long long>(row) * p.ldo + sl * 16 + tch * 4)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        ML_TIMED_WAIT(3, mbar_wait(acc2_full(ab), acc2_par));
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < K::kMaxOutSlabs; ++t) {
          const int sl = part + 4 * t;
          if (sl >= K::kOutSlabs) break;   // warp-uniform
          uint32_t v[16];
          tmem_ld16(acc2_addr + sl * 16, v);
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.b2 + sl * 16) + tch);
          tmem_ld_wait();
          __syncwarp();                      // the previous block has been read back by every lane
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            *reinterpret_cast<uint4*>(stage + lane * 16 + ((cc ^ (lane & 3)) << 2)) =
                make_uint4(v[4 * cc], v[4 * cc + 1], v[4 * cc + 2], v[4 * cc + 3]);
          __syncwarp();
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int lr = g4 * 8 + trow;
            const int row = m0 + q * 32 + lr;
            const float4 a = *reinterpret_cast<const float4*>(stage + lr * 16 + ((tch ^ (lr & 3)) << 2));
            float4 o;
            o.x = a.x + bv.x + rv[t][g4].x; o.y = a.y + bv.y + rv[t][g4].y; o.z = a.z + bv.z + rv[t][g4].z;
            o.w = a.w + bv.w + rv[t][g4].w;
            if (p.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            if (row < p.M) *reinterpret_cast<float4*>(p.out + static_cast<long long>(row) * p.ldo + sl * 16 + tch * 4) = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty(ab));
        // this warp and its partner are done with their carve-outs before either writes H again
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q * 2 + tslot) : "memory");
      }
      ML_MARK(tg);
      ML_SPAN(7, tf, tg);
    }
#ifdef SVX_MLP_PROFILE
    if (warp == 0 && lane == 0 && blockIdx.x == 0) {
      for (int k = 0; k < 4; ++k) g_mlp_prof[8 + k] = (unsigned long long)prof[k];
      g_mlp_prof[12] = (unsigned long long)(clock64() - t_start);
      for (int k = 4; k < 7; ++k) g_mlp_prof[9 + k] = (unsigned long long)prof[k];
      g_mlp_prof[16] = (unsigned long long)prof[7];
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ML_WARP_MMA) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// 2-D fp32 tensor [rows, cols] with row pitch `pitch` (elements); boxes of box_rows x 32 columns, 128B swizzle
int encode_rows_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult res;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &res) == cudaSuccess &&
        res == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 4};
  cuuint32_t box[2] = {ML_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("mlp: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

template <int C, int HC>
int launch_mlp(const CUtensorMap& mx, const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& mr,
               const CUtensorMap& mo, const MlpParams& p, int grid, cudaStream_t st) {
  static thread_local int configured_dev = -1;   // the opt-in is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (configured_dev != cur_dev) {
    SVX_CUDA_OK(cudaFuncSetAttribute(mlp_fused_kernel<C, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     MlpCfg<C, HC>::kSmem));
    configured_dev = cur_dev;
  }
  static const bool pdl = getenv("SVX_PDL") != nullptr;   // opt-in until validated on the GPU tier
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(ML_THREADS);
  cfg.dynamicSmemBytes = MlpCfg<C, HC>::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, mlp_fused_kernel<C, HC>, mx, m1, m2, mr, mo, p);
  if (le != cudaSuccess) return fail("launch of mlp_fused_kernel failed: %s", cudaGetErrorString(le));
  SVX_LAUNCH_OK("mlp_fused_kernel");
  return 0;
}

}  // namespace

struct MlpPrepared {
  CUtensorMap map_x, map_w1, map_w2, map_r, map_o;
  MlpParams p;
  int C, hc, grid;
};

int mlp_prepare(const svx_mlp_desc& d, MlpPrepared** out) {
  *out = nullptr;
  SVX_REQUIRE(d.M > 0 && (d.C == 96 || d.C == 192) && d.hidden == 4 * d.C, "mlp: unsupported shape M=%d C=%d hidden=%d",
              d.M, d.C, d.hidden);
  SVX_REQUIRE(d.x && d.W1 && d.b1 && d.W2 && d.b2 && d.residual && d.out, "mlp: null operand");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  SVX_REQUIRE(al16(d.x) && al16(d.W1) && al16(d.b1) && al16(d.W2) && al16(d.b2) && al16(d.residual) && al16(d.out),
              "mlp: pointers must be 16-byte aligned");
  SVX_REQUIRE(d.ldx % 4 == 0 && d.ldx >= d.C && d.ldo % 4 == 0 && d.ldo >= d.C, "mlp: bad row pitch");
  SVX_REQUIRE(!d.ln_gamma || (d.C == 96 && d.ln_beta && al16(d.ln_gamma) && al16(d.ln_beta)),
              "mlp: the fused LayerNorm exists for C = 96 only and needs gamma and beta");
  MlpPrepared* g = new MlpPrepared();
  // C = 192: HC = 128 (half the fc1 MMAs of an HC = 64 instance, the two H slots used twice per chunk: 0.210 -> 0.177 ms per
  // launch, measured in round 1; the HC = 64 instance is no longer built)
  const int hc = 128;
  int rc = encode_rows_map(&g->map_x, d.x, (uint64_t)d.M, (uint64_t)d.C, (uint64_t)d.ldx, ML_BM);
  if (!rc) rc = encode_rows_map(&g->map_w1, d.W1, (uint64_t)d.hidden, (uint64_t)d.C, (uint64_t)d.C, (uint32_t)hc);
  if (!rc) rc = encode_rows_map(&g->map_w2, d.W2, (uint64_t)d.C, (uint64_t)d.hidden, (uint64_t)d.hidden, (uint32_t)d.C);
  if (!rc) rc = encode_rows_map(&g->map_r, d.residual, (uint64_t)d.M, (uint64_t)d.C, (uint64_t)d.ldo, ML_BM);
  if (!rc) rc = encode_rows_map(&g->map_o, d.out, (uint64_t)d.M, (uint64_t)d.C, (uint64_t)d.ldo, ML_BM);
  if (rc) { delete g; return rc; }
  g->C = d.C;
  g->hc = hc;
  g->p.M = d.M;
  g->p.tiles = (d.M + ML_BM - 1) / ML_BM;
  g->p.b1 = d.b1;
  g->p.b2 = d.b2;
  g->p.residual = d.residual;
  g->p.out = d.out;
  g->p.ldo = d.ldo;
  g->p.round_tf32 = d.round_tf32;
  g->p.ln_gamma = d.ln_gamma;
  g->p.ln_beta = d.ln_beta;
  g->p.ln_eps = d.ln_eps;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  g->grid = g->p.tiles < sms ? g->p.tiles : sms;
  *out = g;
  return 0;
}

void mlp_prepared_free(MlpPrepared* p) { delete p; }

int mlp_launch(const svx_mlp_desc& d, MlpPrepared* prepared, void* stream) {
  MlpPrepared* g = prepared;
  if (!g) {
    if (int rc = mlp_prepare(d, &g)) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (g->C == 96) rc = launch_mlp<96, 128>(g->map_x, g->map_w1, g->map_w2, g->map_r, g->map_o, g->p, g->grid, st);
  else rc = launch_mlp<192, 128>(g->map_x, g->map_w1, g->map_w2, g->map_r, g->map_o, g->p, g->grid, st);
  if (!prepared) delete g;
  return rc;
}

}  // namespace svx

#ifdef SVX_MLP_PROFILE
extern "C" int svx_mlp_prof_read(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, svx::g_mlp_prof, sizeof(unsigned long long) * 20) == cudaSuccess ? 0 : 1;
}
#endif

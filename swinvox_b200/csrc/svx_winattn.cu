// svx_winattn.cu -- W-MSA / SW-MSA of the Swin blocks (timm WindowAttention + the roll / window_partition /
// window_reverse around it; restated in oracle/swin_t.py) on the sm_100a tensor cores: tcgen05.mma with the score and
// output accumulators in TMEM.
//
// One work item = TWO 7x7 windows of one head stacked into a 128-row tile (window w occupies rows 64w .. 64w+48, the
// other 15 rows of each half are padding):
//   S (TMEM, 128 x 128 fp32) = Q K^T      A = Q tile [128 rows x 32 dims], B = K tile [128 keys x 32 dims], both K-major
//   P (smem, 128 x 128)      = exp2(S * scale*log2e + bias*log2e + mask*log2e - rowmax)  on the rows of TMEM: one thread
//                              per query row reads its own window's 64 score columns, writes its 49 probabilities in
//                              the MMA's A-operand layout; the off-diagonal 64 x 64 blocks of P are zero and never
//                              written, which is what keeps the two windows of a tile apart
//   O (TMEM, 128 x 32 fp32)  = P V        A = P, B = V tile [128 keys x 32 dims] read as an MN-major operand: the V rows
//                              are used exactly as they lie in the qkv tensor, no transpose
//   out[token] = O / rowsum
// The cyclic shift and the inverse roll are folded into the token addresses of the row gather / the output rows; the
// -100 attention mask of the shifted blocks is recomputed from the window position ("a key is masked iff it lies in a
// different wrap-around piece than the query"), so neither a rolled copy nor a mask tensor exists.
//
// Storage type T: bf16 (kind::f16 with bf16 operands throughout) or fp32.  fp32 tensors of this path hold TF32-rounded
// values: S = Q K^T runs as kind::tf32 on the rows as stored; for O = P V the probabilities are written as fp16 (11
// significant bits, what a TF32 rounding keeps) and two converter warps rewrite the V rows as fp16 (exact for TF32-rounded
// |v| in [6.1e-5, 65504], saturating beyond -- reported through svx_winattn_desc.range_flag), because kind::tf32 accepts an
// MN-major operand only in a 32-byte-interleaved layout the 16-byte row gather cannot produce, while 16-bit MN-major
// operands take the rows as they are; it also halves the P V instructions.  Accumulation, softmax statistics and the bias
// are fp32 either way.
//
// Warp roles (19 warps):  0-3 / 4-7  two softmax groups (TMEM lane quarter = warp % 4), alternating items: scores -> P
//                         8-11       output group, every item: O / rowsum -> global (rows transposed through a private
//                                    scratch so that a store instruction covers whole token rows)
//                         12-15      row gather, two warps per window: cp.async 16-byte chunks of the token rows into
//                                    128B / 64B-swizzled operand tiles (a ring of stages)
//                         16         MMA issuer (one elected thread), owns the TMEM allocation
//                         17-18      fp32 storage only: V rows fp32 -> fp16
// What bounds it (stage-knockout probes, profiles/r2_winattn_probes.txt and r2_winattn_gather_warps.txt): with two gather
// warps the row gather -- ~110 cp.async instructions per thread and item plus their address arithmetic -- took as long as
// the rest of the pipeline together (316 us per stage-0 launch against 164 us without loads); four gather warps and the
// early stage release give 217 us (4.3 TB/s).  The remaining gap is the per-item chain scores -> softmax -> P -> P.V ->
// output (~5000 cycles through one group; three groups work on three items at once).
#include <cuda_runtime.h>

#include <cstdlib>

#include "svx_internal.h"
#include "svx_ptx.cuh"

namespace svx {
namespace {

constexpr int WS = 7, WT = 49, HD = 32;
// gather warps: 2 (one per window of the tile), 4 or 6 (two / three per window, alternating row groups).  Measured
// (profiles/r2_winattn_gather_warps.txt): two warps could not issue the 16-byte cp.async chunks of an item (plus their
// address arithmetic) faster than ~2.5 us per item -- 316 us per stage-0 launch where the rest of the pipeline needs 164;
// with four 217 us (4.3 TB/s).
#ifndef SVX_WU_LOAD_WARPS
#define SVX_WU_LOAD_WARPS 4
#endif
constexpr int WU_CONV_WARPS = 2;
constexpr int WU_PRODUCER_WARPS = SVX_WU_LOAD_WARPS;
static_assert(WU_PRODUCER_WARPS == 2 || WU_PRODUCER_WARPS == 4 || WU_PRODUCER_WARPS == 6, "gather warps");
constexpr int WU_W_OUT = 8, WU_W_LOAD = 12, WU_W_MMA = WU_W_LOAD + WU_PRODUCER_WARPS, WU_W_CONV = WU_W_MMA + 1;
constexpr int WU_WARPS = WU_W_CONV + WU_CONV_WARPS;      // 17 | 19 | 21
constexpr int WU_THREADS = WU_WARPS * 32;
constexpr float kLog2e = 1.4426950408889634f;

template <typename T>
struct WuCfg {
  static constexpr bool kBf = sizeof(T) == 2;
  static constexpr int kRowB = HD * (int)sizeof(T);          // bytes per Q / K / V row: 128 | 64
  static constexpr int kMatB = 128 * kRowB;                  // one 128-row operand tile: 16 KB | 8 KB
  static constexpr int kStageB = 3 * kMatB;                  // Q | K | V
  static constexpr int kStages = kBf ? 4 : 3;
  static constexpr int kV16B = kBf ? 0 : 128 * 64;           // fp32 storage: the fp16 copy of V, one per stage (8 KB)
  static constexpr int kPBytes = 2 * 128 * 128;              // P is 16-bit either way: two 64-key atoms of 128 rows x 128 B
  static constexpr int kPBufs = kBf ? 2 : 1;
  static constexpr int kBiasB = (WT * WT * 4 + 127) / 128 * 128;
  static constexpr int kKS = kRowB / 32;                     // MMA K steps of S = Q K^T (32 bytes each): 4 | 2
  static constexpr int kKP = 8;                              // MMA K steps of O = P V: 128 keys, 16 per step
  static constexpr int kTokB = 4 * 32 * 4 + 4 * 128 * 4;     // output warps: token row per tile row; 1/rowsum of the last four items
  static constexpr int kScrB = 2 * WT * kRowB;               // output transpose scratch: the 98 live rows of a tile
  static constexpr int kSmem = 1024 + kStages * (kStageB + kV16B) + kPBufs * kPBytes + kBiasB + kTokB + kScrB + 256;
  static_assert(kSmem <= 232448, "window attention shared memory");
};

// instruction descriptors: fp32 accumulate; A K-major; B K-major (S) or MN-major (O = P V: bit 16).
// fmt: operand format code 0 = fp16, 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32)
__host__ __device__ constexpr uint32_t wu_idesc(uint32_t fmt, uint32_t n, bool b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (b_mn ? (1u << 16) : 0u) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// Operand tile descriptor.  K-major (Q, K): rows of 128 | 64 bytes, 8-row groups `sbo` bytes apart.  MN-major (the 16-bit
// V tile): rows = keys (the MMA's K) of 64 bytes = 32 dims (the MMA's N = one swizzle atom wide), exactly as the rows lie
// in the qkv tensor; groups of 8 keys `sbo` = 512 bytes apart.  Layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t wu_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                 // leading byte offset: unused (one atom along the contiguous dimension)
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// keys (bit j = key j = (ky, kx) = (j / 7, j % 7)) whose ky (kx) is >= 4: the second wrap-around piece of a window in the
// last window row (column) of a shifted map.  A query with qy < 4 masks exactly these keys, a query with qy >= 4 the others.
__host__ __device__ constexpr uint64_t wu_keys_ge4(bool by_row) {
  uint64_t m = 0;
  for (int j = 0; j < WT; ++j)
    if ((by_row ? j / WS : j % WS) >= 4) m |= 1ull << j;
  return m;
}
constexpr uint64_t kKeysAll = (1ull << WT) - 1;


// Stall diagnostics (SVX_WINATTN_DEBUG=<address of a pinned, zeroed host buffer of >= 148*21*32 u64>): a barrier wait that
// makes no progress for ~2 s writes who waited for what, and the state of every barrier of the CTA, into host memory
// (readable after the trap has killed the context).
__device__ __noinline__ void wu_report(unsigned long long* dbg, uint32_t bar_base, int nbars, uint32_t tag, uint32_t bar,
                                       uint32_t parity, int i, int extra) {
  if (dbg) {
    unsigned long long* rec = dbg + ((size_t)blockIdx.x * 21 + (threadIdx.x >> 5)) * 32;   // (21 = the largest warp count)
    rec[0] = 0xD1A6000000000000ull | ((unsigned long long)tag << 32) | ((unsigned long long)((bar - bar_base) / 8) << 16) |
             ((unsigned long long)parity << 8) | (threadIdx.x & 31);
    rec[1] = ((unsigned long long)(uint32_t)i << 32) | (uint32_t)extra;
    rec[2] = ((unsigned long long)gridDim.x << 32) | blockIdx.x;
    for (int b = 0; b < nbars && b < 28; ++b) {
      unsigned long long v;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(bar_base + 8u * b));
      rec[3 + b] = v;
    }
    __threadfence_system();
  }
  __trap();
}
// fp32 storage: release an operand stage as soon as S(i) has completed and the V rows have been converted (the P.V product
// reads the fp16 copy), instead of after P.V(i): the stage ring then holds items that are LOADING, not items that wait for
// their softmax (profiles/r2_winattn_probes.txt: the row gather bounds the kernel because little more than one stage was
// in flight).  The fp16 V copies get their own empty barriers.
#ifndef SVX_WU_EARLY_RELEASE
#define SVX_WU_EARLY_RELEASE 1
#endif
#ifndef SVX_WU_BACKOFF_NS
#define SVX_WU_BACKOFF_NS 0
#endif
#ifndef SVX_WU_BACKOFF_TAGS
#define SVX_WU_BACKOFF_TAGS 0x3eu   // every wait site (tags 1 .. 5)
#endif
__device__ __forceinline__ void wu_wait(uint32_t bar, uint32_t parity, unsigned long long* dbg, uint32_t bar_base, int nbars,
                                        uint32_t tag, int i) {
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if (mbar_test(bar, parity)) break;
#if SVX_WU_BACKOFF_NS > 0
    // back-off of the roles selected by SVX_WU_BACKOFF_TAGS (bit = wait-site tag): a polling warp is always eligible and
    // takes issue slots from the softmax warps of its scheduler (profiles/r2_ncu_winattn_v40.txt: 50 % of the samples poll)
    if ((SVX_WU_BACKOFF_TAGS >> tag) & 1u) __nanosleep(SVX_WU_BACKOFF_NS);
#endif
    if ((it & 4095u) == 4095u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) wu_report(dbg, bar_base, nbars, tag, bar, parity, i, 0);
    }
  }
}

// Timing probes (builds with -DSVX_WINATTN_PROBES only; they switch single stages of the pipeline off and give WRONG
// results -- how profiles/r2_winattn_probes.txt was measured).  The product build has none.
#ifdef SVX_WINATTN_PROBES
#define WU_PROBE(bit) ((d.reserved0 & (bit)) != 0)
#else
#define WU_PROBE(bit) false
#endif
#define WU_WAIT(bar, parity, tag) wu_wait(bar, parity, dbg, bar_base, 3 * NS + 12, tag, i)

template <typename T, bool SHIFTED>
__global__ void __launch_bounds__(WU_THREADS, 1) winattn_umma_kernel(const svx_winattn_desc d, int ctas_per_head, unsigned long long* dbg) {
  using K = WuCfg<T>;
  constexpr bool BF = K::kBf;
  constexpr int NS = K::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_smem = smem_base;
  const uint32_t v16_smem = stage_smem + NS * K::kStageB;                   // fp32 storage: fp16 V tiles, one per stage
  constexpr int kPOff = NS * (K::kStageB + K::kV16B);
  const uint32_t p_smem = smem_base + kPOff;
  uint8_t* p_gen = smem_gen + kPOff;
  float* sbias = reinterpret_cast<float*>(smem_gen + kPOff + K::kPBufs * K::kPBytes);
  int* stok_all = reinterpret_cast<int*>(smem_gen + kPOff + K::kPBufs * K::kPBytes + K::kBiasB);
  float* sinv = reinterpret_cast<float*>(stok_all + 4 * 32);                // [4 items][128 rows]
  uint8_t* oscr = smem_gen + kPOff + K::kPBufs * K::kPBytes + K::kBiasB + K::kTokB;
  const uint32_t bar_base = p_smem + K::kPBufs * K::kPBytes + K::kBiasB + K::kTokB + K::kScrB;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // stage s holds Q, K, V of an item
  auto empty_bar = [&](int s) { return bar_base + 8u * (NS + s); };         // every MMA that reads stage s has completed
  auto s_full = [&](int b) { return bar_base + 8u * (2 * NS + b); };        // scores of an item are in TMEM buffer b
  auto s_empty = [&](int b) { return bar_base + 8u * (2 * NS + 2 + b); };   // ... and have been read out
  auto p_full = [&](int b) { return bar_base + 8u * (2 * NS + 4 + b); };    // probabilities are in P buffer b
  auto p_empty = [&](int b) { return bar_base + 8u * (2 * NS + 6 + b); };   // the P V MMAs have read P buffer b
  auto o_full = [&](int b) { return bar_base + 8u * (2 * NS + 8 + b); };
  auto o_empty = [&](int b) { return bar_base + 8u * (2 * NS + 10 + b); };
  auto v16_full = [&](int s) { return bar_base + 8u * (2 * NS + 12 + s); };   // the fp16 copy of stage s's V rows is ready
  auto v16_empty = [&](int s) { return bar_base + 8u * (3 * NS + 13 + s); };  // P.V has read the fp16 V copy of stage s
  constexpr bool kEarly = SVX_WU_EARLY_RELEASE != 0 && !BF;
  const uint32_t tmem_slot = bar_base + 8u * (3 * NS + 12);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (3 * NS + 12));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % d.heads;
  const int cta_in_head = blockIdx.x / d.heads;
  const int nwx = d.W / WS, nwy = d.H / WS;
  const int num_windows = d.N * nwx * nwy;
  const int num_items = (num_windows + 1) >> 1;
  const int nt = cta_in_head < num_items ? (num_items - cta_in_head + ctas_per_head - 1) / ctas_per_head : 0;   // items of this CTA

  // ---- one-time setup: zero the operand stages and P (padding rows / off-diagonal blocks are never written again),
  // the head's bias (x log2 e), barriers, TMEM ----------------------------------------------------------------------
  for (int i = threadIdx.x; i < (kPOff + K::kPBufs * K::kPBytes) / 16; i += WU_THREADS)
    reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < WT * WT; i += WU_THREADS) sbias[i] = __ldg(d.bias + (long long)head * WT * WT + i) * kLog2e;
  if (warp == WU_W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(full_bar(s), WU_PRODUCER_WARPS * 32);
        mbar_init(empty_bar(s), kEarly ? 1u + (uint32_t)WU_CONV_WARPS : 1u);   // early: S(i) committed + both converter warps
        mbar_init(v16_full(s), (uint32_t)WU_CONV_WARPS);
        mbar_init(v16_empty(s), 1u);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(s_full(b), 1u); mbar_init(s_empty(b), 4u);
        mbar_init(p_full(b), 4u); mbar_init(p_empty(b), 1u);
        mbar_init(o_full(b), 1u); mbar_init(o_empty(b), 4u);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  fence_proxy_async_smem();   // the zero fill is read by the tensor cores (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // TMEM columns: scores S[b] at b * 128 (128 columns each), outputs O[b] at 256 + b * 32
  auto item_of = [&](int i) { return cta_in_head + i * ctas_per_head; };
  // window `win` -> (image row base n * H, first rolled row wy * 7 + shift, first rolled column wx * 7 + shift, wy, wx)
  struct WinPos { int nH, y0, x0, wy, wx; };
  auto window_pos = [&](int win) {
    const int wq = win / nwx, wx = win - wq * nwx;
    const int n = wq / nwy, wy = wq - n * nwy;
    return WinPos{n * d.H, wy * WS + d.shift, wx * WS + d.shift, wy, wx};
  };
  // token row (n * H + oy) * W + ox of window position (qy, qx): the cyclic shift is one conditional subtraction per axis
  auto token_of = [&](const WinPos& wp, int qy, int qx) -> long long {
    int oy = wp.y0 + qy, ox = wp.x0 + qx;
    oy -= oy >= d.H ? d.H : 0;
    ox -= ox >= d.W ? d.W : 0;
    return ((long long)(wp.nH + oy)) * d.W + ox;
  };

  if (warp >= WU_W_LOAD && warp < WU_W_MMA) {
    // ---- row gather: each producer warp owns one half of the tile = one window ------------------------------------
    const int w = (warp - WU_W_LOAD) & 1;                          // window of the tile
    constexpr int kSplit = WU_PRODUCER_WARPS / 2;                  // warps per window
    const int part = (warp - WU_W_LOAD) >> 1;                      // this warp takes row groups part, part + kSplit, ...
    constexpr int CH = K::kRowB / 16;            // 16-byte chunks per row: 8 | 4
    constexpr int RPI = 32 / CH;                 // rows one warp-wide cp.async instruction covers: 4 | 8
    const int ch = lane % CH, rsub = lane / CH;
    const size_t tok_pitch = (size_t)3 * d.C * sizeof(T), c_bytes = (size_t)d.C * sizeof(T);
    const uint8_t* qkvb = reinterpret_cast<const uint8_t*>(d.qkv) + (size_t)head * HD * sizeof(T) + ch * 16;
    for (int i = 0; i < nt; ++i) {
      const int s = i % NS;
      WU_WAIT(empty_bar(s), (((uint32_t)(i / NS)) & 1u) ^ 1u, 1u);
      const int win = 2 * item_of(i) + w;
      if (win < num_windows && !WU_PROBE(1)) {   // (probe bit 1: no loads -- timing experiments only)
        const uint32_t dst0 = stage_smem + s * K::kStageB + (64 * w) * K::kRowB;
        const WinPos wp = window_pos(win);
#pragma unroll 1
        for (int r0 = part * RPI; r0 < WT; r0 += kSplit * RPI) {
          const int q = r0 + rsub;
          if (q < WT) {
            const uint8_t* src = qkvb + (size_t)token_of(wp, q / WS, q % WS) * tok_pitch;
            const int row = 64 * w + q;
            // 128B swizzle: chunk ^= row & 7;  64B swizzle: chunk ^= (row >> 1) & 3
            const uint32_t dst = dst0 + q * K::kRowB + ((BF ? (ch ^ ((row >> 1) & 3)) : (ch ^ (row & 7))) << 4);
            cp_async16_zfill(dst, src, 16u);
            cp_async16_zfill(dst + K::kMatB, src + c_bytes, 16u);
            cp_async16_zfill(dst + 2 * K::kMatB, src + 2 * c_bytes, 16u);
          }
        }
      }
      // every producer thread arrives on the stage's barrier when ITS copies have landed (the barrier counts all 64
      // threads): the producers never wait for data, they run ahead as far as the stage ring lets them
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
    }
  } else if (warp == WU_W_MMA) {
    // ---- MMA issuer: S(0) S(1) | PV(0) S(2) | PV(1) S(3) ... -------------------------------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_s = wu_idesc(BF ? 1u : 2u, 128, false);   // bf16 | tf32
      constexpr uint32_t idesc_o = wu_idesc(BF ? 1u : 0u, 32, true);     // bf16 | fp16, V MN-major
      constexpr uint32_t lay = BF ? 4u : 2u;
      constexpr uint32_t sbo = 8 * K::kRowB;     // 8-row groups of the Q / K tiles: 1024 | 512 bytes
      // S(i) needs the item's rows and a free score buffer; P.V(i) needs the fp16 V rows (fp32 storage), the probabilities
      // and a free output buffer.  Issue order S(0) S(1) | PV(0) S(2) | PV(1) S(3) ... -- but never block one kind behind
      // the other: a late row gather must not hold back the P.V of an earlier item (it did: +40 % per launch).
      auto ready_s = [&](int i) {
        return mbar_test(full_bar(i % NS), ((uint32_t)(i / NS)) & 1u) && mbar_test(s_empty(i & 1), (((uint32_t)i >> 1) & 1u) ^ 1u);
      };
      auto ready_pv = [&](int i) {
        const int pb = K::kPBufs == 2 ? (i & 1) : 0;
        const uint32_t pi = K::kPBufs == 2 ? (uint32_t)i >> 1 : (uint32_t)i;
        return (BF || mbar_test(v16_full(i % NS), ((uint32_t)(i / NS)) & 1u)) && mbar_test(p_full(pb), pi & 1u) &&
               mbar_test(o_empty(i & 1), (((uint32_t)i >> 1) & 1u) ^ 1u);
      };
      auto issue_s = [&](int i) {
        const int s = i % NS, b = i & 1;
        tc_fence_after();
        fence_proxy_async_smem();   // the rows were written by cp.async (generic proxy), the MMAs read them through the async proxy
        const uint32_t q_addr = stage_smem + s * K::kStageB;
        const uint64_t da = wu_desc(q_addr, sbo, lay), db = wu_desc(q_addr + K::kMatB, sbo, lay);
#pragma unroll
        for (int k = 0; k < K::kKS; ++k) {
          if constexpr (BF) umma_f16(tmem_base + b * 128, da + 2u * k, db + 2u * k, idesc_s, k != 0 ? 1u : 0u);
          else umma_tf32(tmem_base + b * 128, da + 2u * k, db + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full(b));
        if constexpr (kEarly) umma_commit(empty_bar(s));   // Q and K of this item are dead once the scores exist
      };
      auto issue_pv = [&](int i) {
        const int s = i % NS, b = i & 1, pb = K::kPBufs == 2 ? b : 0;
        tc_fence_after();
        const uint32_t v_addr = BF ? stage_smem + s * K::kStageB + 2 * K::kMatB : v16_smem + s * K::kV16B;
        const uint32_t p_addr = p_smem + pb * K::kPBytes;
#pragma unroll
        for (int k = 0; k < K::kKP; ++k) {
          // A: k-th 32-byte step (16 keys) of the P rows: four steps per 128-byte atom, atoms 16 KB apart;
          // B: the 16 V rows of this step = 1024 bytes (two 8-key groups 512 bytes apart)
          const uint64_t da = umma_desc_sw128(p_addr + (k >> 2) * (128 * 128)) + 2u * (k & 3);
          const uint64_t db = wu_desc(v_addr + k * 1024, 512u, 4u);
          if (WU_PROBE(8)) continue;                                    // (probe bit 8: no P.V MMAs)
          umma_f16(tmem_base + 256 + b * 32, da, db, idesc_o, k != 0 ? 1u : 0u);
        }
        if constexpr (kEarly) umma_commit(v16_empty(s));   // the fp16 V copy of this stage may be rewritten
        else umma_commit(empty_bar(s));                    // Q, K, V of this item are no longer needed
        umma_commit(p_empty(pb));
        umma_commit(o_full(b));
      };
      int ns = 0, np = 0;              // next item whose scores / whose P.V product is to be issued
      long long t0 = 0;
      const bool in_order = WU_PROBE(128);   // (probe bit 128: strict S(0) S(1) | PV(i) S(i+2) order)
      for (uint32_t it = 0; np < nt; ++it) {
        bool did = false;
        if (in_order) {
          if (ns < nt && ns < np + 2) { if (ready_s(ns)) { issue_s(ns); ++ns; did = true; } }
          else if (ready_pv(np)) { issue_pv(np); ++np; did = true; }
        } else {
          if (np < ns && ready_pv(np)) { issue_pv(np); ++np; did = true; }
          // S(i) only after P.V(i - 2) [i - 4 with two P buffers] has been ISSUED: the softmax thread of item i tells "the
          // P buffer is free" from the PARITY of p_empty, which cannot distinguish "P.V(i-1) done" from "P.V(i-3) done".
          // The tensor pipe runs in issue order, so with this gate P.V(i-2) [and its commit] precede the scores the softmax
          // of item i waits for.  (Without it, an output group held up by slow stores let S(i) overtake P.V(i-2): the
          // P buffer was overwritten early, p_full ran one phase ahead, and the CTA deadlocked -- seen only with other
          // streams loading the memory system.)
          if (ns < nt && ns < np + 2 * K::kPBufs && ready_s(ns)) { issue_s(ns); ++ns; did = true; }
        }
        if (did) { it = 0; t0 = 0; }
        else if ((it & 4095u) == 4095u) {   // ~2 s without progress: a protocol bug (see mbar_wait)
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000LL) wu_report(dbg, bar_base, 3 * NS + 12, 6u, bar_base, 0u, np, ns | (nt << 16));
        }
      }
    }
    __syncwarp();
  } else if (warp >= WU_W_CONV) {
    // ---- fp32 storage: V rows (fp32, 128B-swizzled as gathered) -> fp16 rows of 64 bytes (64B swizzle) ---------------
    if constexpr (!BF) {
      const int ct = threadIdx.x - WU_W_CONV * 32;      // 0 .. 63
      uint8_t* v16_gen = smem_gen + NS * K::kStageB;
      uint32_t amax = 0u;
      // unit = (tile row, 16-byte output chunk = 8 dims) over the real rows of the two windows: 392 units, up to four per
      // thread; all loads of a thread are issued before the first conversion (the conversion sits on the path to P.V)
      constexpr int UPT = (2 * WT * 4 + 32 * WU_CONV_WARPS - 1) / (32 * WU_CONV_WARPS);
      int urow[UPT], uoc[UPT];
#pragma unroll
      for (int k = 0; k < UPT; ++k) {
        const int u = ct + k * 32 * WU_CONV_WARPS, rr = u >> 2;
        uoc[k] = u & 3;
        urow[k] = u < 2 * WT * 4 ? (rr < WT ? rr : 64 + rr - WT) : -1;
      }
      for (int i = 0; i < nt; ++i) {
        const int s = i % NS;
        WU_WAIT(full_bar(s), ((uint32_t)(i / NS)) & 1u, 2u);
        // early release: the fp16 copy of this stage is free once P.V(i - NS) has completed (long before, normally)
        if constexpr (kEarly) WU_WAIT(v16_empty(s), (((uint32_t)(i / NS)) & 1u) ^ 1u, 2u);
        const uint8_t* src = smem_gen + s * K::kStageB + 2 * K::kMatB;
        uint8_t* dst = v16_gen + s * K::kV16B;
        float4 a[UPT], b[UPT];
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
          const int row = urow[k] < 0 ? 0 : urow[k], oc = uoc[k];
          a[k] = *reinterpret_cast<const float4*>(src + row * 128 + (((2 * oc) ^ (row & 7)) << 4));
          b[k] = *reinterpret_cast<const float4*>(src + row * 128 + (((2 * oc + 1) ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
          if (urow[k] < 0 || WU_PROBE(32)) continue;   // (probe bit 32: no V conversion)
          const int row = urow[k], oc = uoc[k];
          const uint32_t m0 = max(max(__float_as_uint(a[k].x) & 0x7fffffffu, __float_as_uint(a[k].y) & 0x7fffffffu),
                                  max(__float_as_uint(a[k].z) & 0x7fffffffu, __float_as_uint(a[k].w) & 0x7fffffffu));
          const uint32_t m1 = max(max(__float_as_uint(b[k].x) & 0x7fffffffu, __float_as_uint(b[k].y) & 0x7fffffffu),
                                  max(__float_as_uint(b[k].z) & 0x7fffffffu, __float_as_uint(b[k].w) & 0x7fffffffu));
          amax = max(amax, max(m0, m1));
          uint4 o;
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(a[k].y), "f"(a[k].x));
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(a[k].w), "f"(a[k].z));
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(b[k].y), "f"(b[k].x));
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(b[k].w), "f"(b[k].z));
          *reinterpret_cast<uint4*>(dst + row * 64 + ((oc ^ ((row >> 1) & 3)) << 4)) = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(v16_full(s));
          if constexpr (kEarly) mbar_arrive(empty_bar(s));   // this warp has read its share of the fp32 V rows
        }
      }
      if (d.range_flag && amax > 0x477fe000u) atomicOr(d.range_flag, 1);   // a V value beyond fp16's 65504 saturated
    }
  } else if (warp < WU_W_OUT) {
    // ---- softmax: group g = warp / 4 takes the items with i % 2 == g; thread = one query row ---------------------------
    const int grp = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;           // tile row = TMEM lane
    const int w = r >> 6, q = r & 63;            // window of the tile, row inside the window
    const bool qreal = q < WT;
    const int qy = q / WS, qx = q - qy * WS;
    const float scale2 = d.scale * kLog2e;
    const float kMask = -100.f * kLog2e;
    const float* brow = sbias + (qreal ? q : 0) * WT;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr uint64_t kRowGe4 = wu_keys_ge4(true), kColGe4 = wu_keys_ge4(false);
    for (int i = grp; i < nt; i += 2) {
      const int b = i & 1, pb = K::kPBufs == 2 ? b : 0;
      const int win = 2 * item_of(i) + w;
      const bool live = qreal && win < num_windows;
      uint64_t masked = 0;   // keys this query must not see (shifted blocks only)
      if (SHIFTED && live) {
        const WinPos wp = window_pos(win);
        if (wp.wy == nwy - 1) masked |= qy < 4 ? kRowGe4 : (kKeysAll & ~kRowGe4);
        if (wp.wx == nwx - 1) masked |= qx < 4 ? kColGe4 : (kKeysAll & ~kColGe4);
      }
      WU_WAIT(s_full(b), ((uint32_t)i >> 1) & 1u, 3u);
      tc_fence_after();
      // this row's scores against the 64 key slots of its own window (slots 49..63 are padding); the values are turned
      // into logits and then probabilities in place
      uint32_t sv[4][16];
      const uint32_t s_addr = tmem_base + lane_sel + b * 128 + w * 64;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16(s_addr + 16 * c, sv[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));
      auto E = [&](int j) -> uint32_t& { return sv[j >> 4][j & 15]; };
      float mxp[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four partial maxima / sums: short dependency chains
      if (!WU_PROBE(2)) {   // (probe bit 2: no softmax arithmetic -- timing experiments only)
#pragma unroll
        for (int j = 0; j < WT; ++j) {
          float v = fmaf(__uint_as_float(E(j)), scale2, brow[j]);
          if (SHIFTED) v += ((masked >> j) & 1ull) ? kMask : 0.f;
          E(j) = __float_as_uint(v);
          mxp[j & 3] = fmaxf(mxp[j & 3], v);
        }
      }
      const float mx = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
      float smp[4] = {0.f, 0.f, 0.f, 0.f};
      if (WU_PROBE(2)) smp[0] = 1.f;
      else {
#pragma unroll
        for (int j = 0; j < WT; ++j) {
          const float e = ex2f(__uint_as_float(E(j)) - mx);
          E(j) = __float_as_uint(e);
          smp[j & 3] += e;
        }
      }
      const float sum = (smp[0] + smp[1]) + (smp[2] + smp[3]);
      // P row in the A-operand layout (K-major, 128B swizzle): only the 49 (+7 zero) slots of the own window
      const uint32_t pi = K::kPBufs == 2 ? (uint32_t)i >> 1 : (uint32_t)i;
      WU_WAIT(p_empty(pb), (pi & 1u) ^ 1u, 4u);
      sinv[(i & 3) * 128 + r] = __frcp_rn(sum);   // for the output group (ring of four items: see the barrier order there)
      if (live) {
        uint8_t* prow = p_gen + pb * K::kPBytes + r * 128;
        // 64 sixteen-bit slots of window w = atom w; chunk c holds slots 8c .. 8c+7 (bf16, or fp16 for fp32 storage)
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          uint32_t u[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int j0 = 8 * c + 2 * h;
            const float lo = j0 < WT ? __uint_as_float(E(j0 < WT ? j0 : 0)) : 0.f;
            const float hi = j0 + 1 < WT ? __uint_as_float(E(j0 + 1 < WT ? j0 + 1 : 0)) : 0.f;
            if constexpr (BF) u[h] = pack_bf16x2(lo, hi);
            else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[h]) : "f"(hi), "f"(lo));
          }
          *reinterpret_cast<uint4*>(prow + w * (128 * 128) + ((c ^ (r & 7)) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(pb));
    }
  } else {
    // ---- output group (warps 8-11), every item: O / rowsum -> out[token, head*32 ..] -------------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int w = r >> 6, q = r & 63;
    const bool qreal = q < WT;
    const int qy = q / WS, qx = q - qy * WS;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    int* stok = stok_all + quarter * 32;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int win = 2 * item_of(i) + w;
      const bool live = qreal && win < num_windows;
      long long tok = 0;
      if (live) tok = token_of(window_pos(win), qy, qx);
      WU_WAIT(o_full(b), ((uint32_t)i >> 1) & 1u, 5u);
      tc_fence_after();
      uint32_t ov[2][16];
      const uint32_t o_addr = tmem_base + lane_sel + 256 + b * 32;
      __syncwarp();
      tmem_ld16(o_addr, ov[0]);
      tmem_ld16(o_addr + 16, ov[1]);
      tmem_ld_wait();
      tc_fence_before();
      // 1/rowsum was written by the softmax thread of this row before it arrived on p_full, which the MMA thread acquired
      // before issuing P.V(i), whose completion this thread has just observed.  Read BEFORE o_empty is released: the ring
      // slot (i & 3) is rewritten by item i + 4, whose scores cannot even be issued before P.V(i + 2), which waits for
      // that release.
      const float inv = sinv[(i & 3) * 128 + r];
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty(b));
      // Row per lane -> global would touch 32 different cache lines per store instruction (measured: the output phase was
      // the longest part of an item).  Each warp therefore transposes its rows through a small scratch and stores whole
      // rows: 8 (4) lanes cover the 128 (64) bytes of one token's head slice.
      // this warp's live rows (32 of the first, 17 of the second quarter of each window half) in the transpose scratch
      uint8_t* scratch = oscr + ((quarter >> 1) * WT + (quarter & 1) * 32) * K::kRowB;
      stok[lane] = live ? (int)tok : -1;
      if (live) {
        uint8_t* srow = scratch + lane * K::kRowB;
        if constexpr (BF) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t* o8 = &ov[c >> 1][(c & 1) * 8];
            *reinterpret_cast<uint4*>(srow + ((c ^ ((r >> 1) & 3)) << 4)) =
                make_uint4(pack_bf16x2(__uint_as_float(o8[0]) * inv, __uint_as_float(o8[1]) * inv),
                           pack_bf16x2(__uint_as_float(o8[2]) * inv, __uint_as_float(o8[3]) * inv),
                           pack_bf16x2(__uint_as_float(o8[4]) * inv, __uint_as_float(o8[5]) * inv),
                           pack_bf16x2(__uint_as_float(o8[6]) * inv, __uint_as_float(o8[7]) * inv));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t* o4 = &ov[c >> 2][(c & 3) * 4];
            float4 o = make_float4(__uint_as_float(o4[0]) * inv, __uint_as_float(o4[1]) * inv, __uint_as_float(o4[2]) * inv,
                                   __uint_as_float(o4[3]) * inv);
            if (d.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            *reinterpret_cast<float4*>(srow + ((c ^ (r & 7)) << 4)) = o;
          }
        }
      }
      __syncwarp();
      constexpr int CH = K::kRowB / 16;        // 16-byte chunks per row: 8 | 4
      constexpr int RPI = 32 / CH;             // rows per store instruction: 4 | 8
      const int ch = lane % CH, rsub = lane / CH;
      T* outh = reinterpret_cast<T*>(d.out) + head * HD + ch * (16 / (int)sizeof(T));
      if (!WU_PROBE(16)) {               // (probe bit 16: no output stores)
#pragma unroll
        for (int p0 = 0; p0 < 32; p0 += RPI) {
          const int lr = p0 + rsub, row = quarter * 32 + lr;
          const int tk = stok[lr];
          if (tk >= 0) {   // (rows without a token do not exist in the scratch)
            const uint4 v = *reinterpret_cast<const uint4*>(scratch + lr * K::kRowB +
                                                            ((BF ? (ch ^ ((row >> 1) & 3)) : (ch ^ (row & 7))) << 4));
            *reinterpret_cast<uint4*>(outh + (size_t)tk * d.C) = v;
          }
        }
      }
      __syncwarp();   // the scratch is reused by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WU_W_MMA) tmem_dealloc<512>(tmem_base);
}

}  // namespace

int winattn_umma_launch(const svx_winattn_desc& d_in, void* stream) {
  svx_winattn_desc d = d_in;
  d.reserved0 = 0;
#ifdef SVX_WINATTN_PROBES
  static const char* probe = getenv("SVX_WINATTN_PROBE");   // 1 no loads, 2 no softmax, 8 no P.V, 16 no stores, 32 no V conversion,
  d.reserved0 = probe ? atoi(probe) : 0;                    // 128 strictly ordered MMA issue
#endif
  static const char* dbg_env = getenv("SVX_WINATTN_DEBUG");
  unsigned long long* dbg = dbg_env ? reinterpret_cast<unsigned long long*>(strtoull(dbg_env, nullptr, 0)) : nullptr;
  const long long windows = (long long)d.N * (d.H / WS) * (d.W / WS);
  const long long items = (windows + 1) / 2;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  long long per_head = sms / d.heads;      // one CTA per SM (it takes most of the shared memory), whole heads
  if (per_head < 1) per_head = 1;
  if (per_head > items) per_head = items;
  const int grid = (int)(per_head * d.heads);
  const bool bf = d.dtype == SVX_DT_BF16;
  cudaStream_t st = (cudaStream_t)stream;
  SVX_CUDA_OK(cudaFuncSetAttribute(winattn_umma_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WuCfg<float>::kSmem));
  SVX_CUDA_OK(cudaFuncSetAttribute(winattn_umma_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WuCfg<float>::kSmem));
  SVX_CUDA_OK(cudaFuncSetAttribute(winattn_umma_kernel<bf16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WuCfg<bf16_t>::kSmem));
  SVX_CUDA_OK(cudaFuncSetAttribute(winattn_umma_kernel<bf16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WuCfg<bf16_t>::kSmem));
  if (bf) {
    if (d.shift > 0) winattn_umma_kernel<bf16_t, true><<<grid, WU_THREADS, WuCfg<bf16_t>::kSmem, st>>>(d, (int)per_head, dbg);
    else winattn_umma_kernel<bf16_t, false><<<grid, WU_THREADS, WuCfg<bf16_t>::kSmem, st>>>(d, (int)per_head, dbg);
  } else {
    if (d.shift > 0) winattn_umma_kernel<float, true><<<grid, WU_THREADS, WuCfg<float>::kSmem, st>>>(d, (int)per_head, dbg);
    else winattn_umma_kernel<float, false><<<grid, WU_THREADS, WuCfg<float>::kSmem, st>>>(d, (int)per_head, dbg);
  }
  SVX_LAUNCH_OK("winattn_umma_kernel");
  return 0;
}

}  // namespace svx

// svx_hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A CPU twin of the device kernels in swinvox_b200/csrc, compiled with g++ into
// tests/hostsim/libsvx_hostsim.so.  It implements the same descriptor semantics
// (include/swinvox_b200.h) with plain loops so that the host logic of the product -- weight
// re-layout, BatchNorm folding, tap tables, transposed-convolution parity classes, window/shift
// indexing, plan construction -- can be checked against the oracle in the CPU-only test tier.
// The swinvox_b200 package never loads this library; tests inject it explicitly.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../swinvox_b200/csrc/svx_internal.h"

namespace svx {

static inline float tf32_trunc(float x) {  // what tcgen05 kind::tf32 sees of an fp32 operand
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static inline float tf32_rna(float x) {  // cvt.rna.tf32.f32
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static inline float rnd(float x, int r) { return r ? tf32_rna(x) : x; }
// bf16 storage (the encoder's reduced-precision mode): round-to-nearest-even on store, exact widening on load
static inline float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float x; memcpy(&x, &u, 4); return x; }
static inline uint16_t f2bf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
// element i of a tensor stored as fp32 or bf16
static inline float ldx(const void* p, long long i, bool bf) {
  return bf ? bf2f(reinterpret_cast<const uint16_t*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
static inline void stx(void* p, long long i, float v, bool bf) {
  if (bf) reinterpret_cast<uint16_t*>(p)[i] = f2bf(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

static inline float act_fn(float x, int act, float slope) {
  switch (act) {
    case SVX_ACT_RELU: return x > 0.f ? x : 0.f;
    case SVX_ACT_LEAKY: return x > 0.f ? x : x * slope;
    case SVX_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    default: return x;
  }
}

struct GemmPrepared { int unused; };
int gemm_prepare(const svx_gemm_desc& d, GemmPrepared** out) {
  *out = nullptr;
  SVX_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "gemm: empty problem");
  const bool bf = d.operand_kind == SVX_OPERAND_BF16;
  const int BKE = bf ? 64 : 32, e16 = bf ? 8 : 4, esz = bf ? 2 : 4;
  SVX_REQUIRE(d.Kpad % BKE == 0 && d.Kpad >= d.K, "gemm: bad Kpad");
  SVX_REQUIRE(bf || d.io_flags == 0, "gemm: bf16 outputs need bf16 operands");
  SVX_REQUIRE(!bf || (d.a_mode != SVX_A_SLAB3 && d.epi_mode == SVX_EPI_STD && d.N % 4 == 0), "gemm: bf16 operands: standard epilogue only");
  SVX_REQUIRE(!d.residual || d.res_via_mma || (((d.io_flags & SVX_IO_RES_BF16) != 0) == ((d.io_flags & SVX_IO_OUT_BF16) != 0)),
              "gemm: residual / output storage types differ");
  SVX_REQUIRE(d.Npad % d.block_n == 0 && d.Npad >= d.N, "gemm: bad Npad");
  if (d.a_mode == SVX_A_GATHER)
    SVX_REQUIRE(d.Cin % e16 == 0 && d.in_c0 % e16 == 0 && d.in_Cs % e16 == 0 && d.K == d.ntaps * d.Cin, "gemm: bad gather");
  else if (d.a_mode == SVX_A_IM2COL) {
    SVX_REQUIRE((d.Cin * esz == 16 || ((d.Cin * esz == 32 || d.Cin % BKE == 0) && d.Kpad == d.K)) && d.in_c0 % e16 == 0 && d.in_Cs % e16 == 0 && d.K == d.ntaps * d.Cin &&
                    d.taps_host && d.ntaps <= 64 && d.M % (d.out_D * d.out_H * d.out_W) == 0,
                "gemm: bad im2col");
    const int in_ext[3] = {d.in_D, d.in_H, d.in_W}, out_ext[3] = {d.out_D, d.out_H, d.out_W};
    const int str[3] = {d.stride_d, d.stride_h, d.stride_w};
    for (int a = 0; a < 3; ++a) {   // the corner range of the hardware descriptor
      int lo = 1 << 20, hi = -(1 << 20);
      for (int t = 0; t < d.ntaps; ++t) { lo = std::min(lo, d.taps_host[4 * t + a]); hi = std::max(hi, d.taps_host[4 * t + a]); }
      const int up = lo + (out_ext[a] - 1) * str[a] - (in_ext[a] - 1);
      SVX_REQUIRE(lo >= -16 && lo <= 15 && up >= -16 && up <= 15 && hi - lo <= 255 && str[a] >= 1 && str[a] <= 8,
                  "gemm: im2col corners out of range");
    }
  }
  else if (d.a_mode == SVX_A_FLAT)
    SVX_REQUIRE(d.Cin % BKE == 0 && d.in_Cs % e16 == 0 && d.K == d.ntaps * d.Cin && d.Kpad == d.K && d.taps_host && d.ntaps <= 64 &&
                    d.out_D == d.in_D && d.out_H == d.in_H && d.out_W == d.in_W && d.lda >= d.M,
                "gemm: bad flat conv");
  else if (d.a_mode == SVX_A_SLAB3)
    SVX_REQUIRE(d.Cin == 32 && d.N <= 16 && d.block_n == 48 && d.Npad == 48 && d.K == 288 && d.Kpad == 288 &&
                    d.in_D == d.valid_D + 2 && d.valid_H <= d.in_H && d.valid_W <= d.in_W && 128 + 2 * d.in_W <= 200 &&
                    d.lda > 0 && d.epi_mode == SVX_EPI_STD,
                "gemm: bad slab conv");
  else
    SVX_REQUIRE(d.lda % e16 == 0 && d.lda >= d.K, "gemm: bad lda");
  if (d.res_via_mma) SVX_REQUIRE(d.block_n % BKE == 0 && (!bf || (d.io_flags & SVX_IO_RES_BF16)), "gemm: bad res_via_mma");
  if (d.epi_mode == SVX_EPI_CONVT8)
    SVX_REQUIRE(d.a_mode != SVX_A_SLAB3 && d.cls_cout > 0 && d.N == 8 * d.cls_cout && (!d.epi_aux || (d.cls_cout == 8 && d.epi_out2)),
                "gemm: bad transposed-convolution class epilogue");
  if (d.epi_mode == SVX_EPI_DEC_TAIL)
    SVX_REQUIRE(d.block_n == 16 && d.N == 16 && d.epi_aux && d.epi_out2, "gemm: bad decoder tail");
  *out = new GemmPrepared();
  return 0;
}
void gemm_prepared_free(GemmPrepared* p) { delete p; }
int gemm_num_launches(const svx_gemm_desc&) { return 1; }

int gemm_launch(const svx_gemm_desc& d, GemmPrepared* prepared, void*) {
  if (!prepared) {
    GemmPrepared* g = nullptr;
    if (int rc = gemm_prepare(d, &g)) return rc;
    delete g;
  }
  if (d.a_mode == SVX_A_SLAB3) {
    // direct 3x3x3 convolution over the zero-bordered volume; weights in the kw-in-N layout
    const int live = d.cin_live > 0 ? d.cin_live : 32;
    const long long HWp = (long long)d.in_H * d.in_W;
    const long long vox = (long long)d.valid_D * d.valid_H * d.valid_W;
    const bool f16 = d.operand_kind != SVX_OPERAND_TF32;   // the device converts operands to fp16 (saturating)
    const float acc_scale = d.acc_scale != 0.f ? d.acc_scale : 1.f;
    int saturated = 0;
#pragma omp parallel for schedule(static) reduction(| : saturated)
    for (long long r = 0; r < d.M; ++r) {
      const long long n = r / vox;
      long long t = r % vox;
      const int w = (int)(t % d.valid_W); t /= d.valid_W;
      const int h = (int)(t % d.valid_H);
      const int dd = (int)(t / d.valid_H);
      const long long off = d.o_base + n * d.o_sn + dd * d.o_sd + h * d.o_sh + w * d.o_sw;
      for (int co = 0; co < d.N; ++co) {
        float acc = 0.f;
        for (int kd = 0; kd < 3; ++kd)
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              const long long row = (n * d.in_D + dd + kd) * HWp + (long long)(h + kh) * d.in_W + (w + kw);
              if (row >= d.lda) continue;
              const float* px = reinterpret_cast<const float*>(d.A) + row * d.in_Cs + d.in_c0;
              const float* wr = reinterpret_cast<const float*>(d.W) + (long long)(kw * 16 + co) * d.Kpad + (kd * 3 + kh) * 32;
              for (int c = 0; c < live; ++c)
                if (d.in_c0 + c < d.in_Cs) {
                  float a = tf32_trunc(px[c]);
                  if (f16 && !(std::fabs(a) <= 65504.f)) { saturated = 1; a = a > 0.f ? 65504.f : -65504.f; }
                  acc += a * tf32_trunc(wr[c]);
                }
            }
        float v = acc * acc_scale + (d.bias ? d.bias[co] : 0.f);
        const float res = d.residual ? reinterpret_cast<const float*>(d.residual)[off + co] : 0.f;
        if (d.residual && !d.res_after_act) v += res;
        v = act_fn(v, d.act, d.act_param);
        if (d.residual && d.res_after_act) v += res;
        v *= d.out_scale;
        reinterpret_cast<float*>(d.out)[off + co] = rnd(v, d.round_tf32);
      }
    }
    if (f16 && saturated && d.range_flag) *d.range_flag |= 1;
    return 0;
  }
  const long long rows_per_n = (long long)d.out_D * d.out_H * d.out_W;
  const long long w_pitch = d.Kpad + (d.res_via_mma ? d.block_n : 0);   // identity columns appended for the device
  const bool bfA = d.operand_kind == SVX_OPERAND_BF16;
  const bool bfO = (d.io_flags & SVX_IO_OUT_BF16) != 0, bfR = (d.io_flags & SVX_IO_RES_BF16) != 0;
  // what the tensor cores see of an operand element: fp32 storage is truncated to TF32, bf16 storage is exact
  auto opA = [&](long long i) { return bfA ? ldx(d.A, i, true) : tf32_trunc(reinterpret_cast<const float*>(d.A)[i]); };
  auto opW = [&](long long i) { return bfA ? ldx(d.W, i, true) : tf32_trunc(reinterpret_cast<const float*>(d.W)[i]); };
  auto RES = [&](long long i) { return ldx(d.residual, i, bfR); };
  auto OUT = [&](long long i, float v) { stx(d.out, i, v, bfO); };
#pragma omp parallel
  {
    std::vector<float> arow(d.K);
#pragma omp for schedule(static)
    for (int r = 0; r < d.M; ++r) {
      const int ow = r % d.out_W;
      int t = r / d.out_W;
      const int oh = t % d.out_H;
      t /= d.out_H;
      const int od = t % d.out_D;
      const long long n = t / d.out_D;
      (void)rows_per_n;
      if (d.valid_W > 0 && (ow >= d.valid_W || oh >= d.valid_H || od >= d.valid_D)) continue;
      if (d.a_mode == SVX_A_PLAIN) {
        for (int k = 0; k < d.K; ++k) arow[k] = opA((long long)r * d.lda + k);
      } else if (d.a_mode == SVX_A_FLAT) {
        for (int tap = 0; tap < d.ntaps; ++tap) {
          const long long row = r + ((long long)d.taps_host[tap * 4] * d.in_H + d.taps_host[tap * 4 + 1]) * d.in_W +
                                d.taps_host[tap * 4 + 2];
          const bool ok = row < d.lda;  // TMA zero-fills rows past the end of the matrix
          const long long px = row * d.in_Cs + d.in_c0;
          // TMA zero-fills columns past the end of a row as well (a 32-channel box may overhang the tensor)
          for (int c = 0; c < d.Cin; ++c)
            arow[tap * d.Cin + c] = (ok && d.in_c0 + c < d.in_Cs) ? opA(px + c) : 0.f;
        }
      } else {
        const int32_t* tp = d.a_mode == SVX_A_IM2COL ? d.taps_host : d.taps;
        for (int tap = 0; tap < d.ntaps; ++tap) {
          const int id = od * d.stride_d + tp[tap * 4 + 0];
          const int ih = oh * d.stride_h + tp[tap * 4 + 1];
          const int iw = ow * d.stride_w + tp[tap * 4 + 2];
          const bool ok = id >= 0 && id < d.in_D && ih >= 0 && ih < d.in_H && iw >= 0 && iw < d.in_W;
          const long long px = (((n * d.in_D + id) * d.in_H + ih) * (long long)d.in_W + iw) * d.in_Cs + d.in_c0;
          for (int c = 0; c < d.Cin; ++c) arow[tap * d.Cin + c] = ok ? opA(px + c) : 0.f;
        }
      }
      const long long off = d.o_base + n * d.o_sn + od * d.o_sd + oh * d.o_sh + ow * d.o_sw;
      float x[16];
      if (d.epi_mode == SVX_EPI_DEC_TAIL) {
        float g = d.epi_aux[8];
        for (int j = 0; j < 8; ++j) {
          float acc = 0.f;
          const long long w = (long long)j * w_pitch;
          for (int k = 0; k < d.K; ++k) acc += arow[k] * opW(w + k);
          acc += d.bias ? d.bias[j] : 0.f;
          x[j] = acc > 0.f ? acc : 0.f;
          g = fmaf(d.epi_aux[j], x[j], g);
        }
        x[8] = g;
        for (int j = 9; j < 16; ++j) x[j] = 0.f;
        d.epi_out2[d.o2_base + n * d.o2_sn + od * d.o2_sd + oh * d.o2_sh + ow * d.o2_sw] = g;
        for (int j = 0; j < 16; ++j) OUT(off + j, rnd(x[j], d.round_tf32));
        continue;
      }
      if (d.epi_mode == SVX_EPI_CONVT8) {
        const int cc = d.cls_cout;
        for (int cls = 0; cls < 8; ++cls) {
          const long long o = off + (cls >> 2) * d.c_sd + ((cls >> 1) & 1) * d.c_sh + (cls & 1) * d.c_sw;
          float g = d.epi_aux ? d.epi_aux[8] : 0.f;
          for (int c = 0; c < cc; ++c) {
            const int j = cls * cc + c;
            float acc = 0.f;
            const long long w = (long long)j * w_pitch;
            for (int k = 0; k < d.K; ++k) acc += arow[k] * opW(w + k);
            float v = acc + (d.bias ? d.bias[j] : 0.f);
            if (d.epi_aux) {
              v = v > 0.f ? v : 0.f;
              g = fmaf(d.epi_aux[c], v, g);
            } else {
              const float res = d.residual ? RES(o + c) : 0.f;
              if (d.residual && !d.res_after_act) v += res;
              v = act_fn(v, d.act, d.act_param);
              if (d.residual && d.res_after_act) v += res;
              v *= d.out_scale;
            }
            OUT(o + c, rnd(v, d.round_tf32));
          }
          if (d.epi_aux) {
            OUT(o + 8, rnd(g, d.round_tf32));
            OUT(o + 9, 0.f); OUT(o + 10, 0.f); OUT(o + 11, 0.f);
            d.epi_out2[d.o2_base + n * d.o2_sn + od * d.o2_sd + oh * d.o2_sh + ow * d.o2_sw + (cls >> 2) * d.c2_sd +
                       ((cls >> 1) & 1) * d.c2_sh + (cls & 1) * d.c2_sw] = g;
          }
        }
        continue;
      }
      if (d.epi_mode == SVX_EPI_POOL8) {
        const int nc = d.N / 8;
        for (int c = 0; c < nc; ++c) {
          float best = -INFINITY;
          for (int gq = 0; gq < 8; ++gq) {
            float acc = 0.f;
            const long long w = (long long)(gq * nc + c) * w_pitch;
            for (int k = 0; k < d.K; ++k) acc += arow[k] * opW(w + k);
            best = std::max(best, acc);
          }
          float v = act_fn(best + (d.bias ? d.bias[c] : 0.f), d.act, d.act_param) * d.out_scale;
          OUT(off + c, rnd(v, d.round_tf32));
        }
        continue;
      }
      for (int j = 0; j < d.N; ++j) {
        float acc = 0.f;
        const long long w = (long long)j * w_pitch;
        for (int k = 0; k < d.K; ++k) acc += arow[k] * opW(w + k);
        float v = acc + (d.bias ? d.bias[j] : 0.f);
        const float res = d.residual ? RES(off + j) : 0.f;
        if (d.residual && !d.res_after_act) v += res;
        v = act_fn(v, d.act, d.act_param);
        if (d.residual && d.res_after_act) v += res;
        v *= d.out_scale;
        OUT(off + j, rnd(v, d.round_tf32));
      }
    }
  }
  return 0;
}

int im2col_launch(const svx_im2col_desc& d, void*) {
  const int K = d.KD * d.KH * d.KW * d.C;
  const long long rows = (long long)d.N * d.OD * d.OH * d.OW;
#pragma omp parallel for
  for (long long r = 0; r < rows; ++r) {
    long long t = r;
    const int ow = t % d.OW; t /= d.OW;
    const int oh = t % d.OH; t /= d.OH;
    const int od = t % d.OD;
    const long long n = t / d.OD;
    float* dst = d.out + r * d.Kpad;
    int k = 0;
    for (int kd = 0; kd < d.KD; ++kd)
      for (int kh = 0; kh < d.KH; ++kh)
        for (int kw = 0; kw < d.KW; ++kw)
          for (int c = 0; c < d.C; ++c, ++k) {
            const int id = od * d.stride - d.pad_d + kd, ih = oh * d.stride - d.pad_h + kh,
                      iw = ow * d.stride - d.pad_w + kw;
            float v = 0.f;
            if (id >= 0 && id < d.D && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W)
              v = d.in[n * d.s_n + c * d.s_c + id * d.s_d + ih * d.s_h + iw * d.s_w];
            dst[k] = rnd(v, d.round_tf32);
          }
    for (; k < d.Kpad; ++k) dst[k] = 0.f;
  }
  return 0;
}

int pool_launch(const svx_pool_desc& d, void*) {
  const long long rows = (long long)d.N * d.OD * d.OH * d.OW;
#pragma omp parallel for
  for (long long r = 0; r < rows; ++r) {
    long long t = r;
    const int ow = t % d.OW; t /= d.OW;
    const int oh = t % d.OH; t /= d.OH;
    const int od = t % d.OD;
    const long long n = t / d.OD;
    for (int c = 0; c < d.C; ++c) {
      float acc = d.mode == SVX_POOL_MAX ? -INFINITY : 0.f;
      int cnt = 0;
      for (int kd = 0; kd < d.KD; ++kd)
        for (int kh = 0; kh < d.KH; ++kh)
          for (int kw = 0; kw < d.KW; ++kw) {
            const int id = od * d.SD - d.PD + kd, ih = oh * d.SH - d.PH + kh, iw = ow * d.SW - d.PW + kw;
            if (id < 0 || id >= d.D || ih < 0 || ih >= d.H || iw < 0 || iw >= d.W) continue;
            const float v = ldx(d.in, (((n * d.D + id) * d.H + ih) * d.W + iw) * (long long)d.in_Cs + c, d.dtype & SVX_DT_IN_BF16);
            acc = d.mode == SVX_POOL_MAX ? std::max(acc, v) : acc + v;
            ++cnt;
          }
      if (d.mode == SVX_POOL_AVG) acc /= (float)std::max(cnt, 1);
      stx(d.out, r * d.out_Cs + c, rnd(acc, d.round_tf32), d.dtype & SVX_DT_OUT_BF16);
    }
  }
  return 0;
}

struct MlpPrepared { int unused; };
int mlp_prepare(const svx_mlp_desc& d, MlpPrepared** out) {
  *out = nullptr;
  SVX_REQUIRE(d.M > 0 && (d.C == 96 || d.C == 192) && d.hidden == 4 * d.C, "mlp: unsupported shape");
  SVX_REQUIRE(d.x && d.W1 && d.b1 && d.W2 && d.b2 && d.residual && d.out, "mlp: null operand");
  SVX_REQUIRE(d.ldx % 4 == 0 && d.ldx >= d.C && d.ldo % 4 == 0 && d.ldo >= d.C, "mlp: bad row pitch");
  SVX_REQUIRE(!d.ln_gamma || (d.C == 96 && d.ln_beta), "mlp: the fused LayerNorm exists for C = 96 only");
  *out = new MlpPrepared();
  return 0;
}
void mlp_prepared_free(MlpPrepared* p) { delete p; }

// out = residual + W2 . round_tf32(gelu(W1 . x + b1)) + b2, operands as kind::tf32 sees them
int mlp_launch(const svx_mlp_desc& d, MlpPrepared* prepared, void*) {
  if (!prepared) {
    MlpPrepared* g = nullptr;
    if (int rc = mlp_prepare(d, &g)) return rc;
    delete g;
  }
#pragma omp parallel
  {
    std::vector<float> xr(d.C), h(d.hidden);
#pragma omp for schedule(static)
    for (int r = 0; r < d.M; ++r) {
      if (d.ln_gamma) {   // fused pre-LayerNorm: fc1 reads round_tf32(LN(x) * gamma + beta)
        double mean = 0, var = 0;
        for (int k = 0; k < d.C; ++k) mean += d.x[(long long)r * d.ldx + k];
        mean /= d.C;
        for (int k = 0; k < d.C; ++k) { const double t = d.x[(long long)r * d.ldx + k] - mean; var += t * t; }
        var /= d.C;
        const float rstd = 1.f / sqrtf((float)var + d.ln_eps);
        for (int k = 0; k < d.C; ++k)
          xr[k] = tf32_rna((d.x[(long long)r * d.ldx + k] - (float)mean) * rstd * d.ln_gamma[k] + d.ln_beta[k]);
      } else
      for (int k = 0; k < d.C; ++k) xr[k] = tf32_trunc(d.x[(long long)r * d.ldx + k]);
      for (int j = 0; j < d.hidden; ++j) {
        float acc = 0.f;
        const float* w = d.W1 + (long long)j * d.C;
        for (int k = 0; k < d.C; ++k) acc += xr[k] * tf32_trunc(w[k]);
        h[j] = tf32_rna(act_fn(acc + d.b1[j], SVX_ACT_GELU, 0.f));
      }
      for (int c = 0; c < d.C; ++c) {
        float acc = 0.f;
        const float* w = d.W2 + (long long)c * d.hidden;
        for (int j = 0; j < d.hidden; ++j) acc += h[j] * tf32_trunc(w[j]);
        const long long o = (long long)r * d.ldo + c;
        d.out[o] = rnd(acc + d.b2[c] + d.residual[o], d.round_tf32);
      }
    }
  }
  return 0;
}

int lnrows_launch(const svx_lnrows_desc& d, void*) {
#pragma omp parallel
  {
    std::vector<float> row(d.C);
#pragma omp for
    for (int r = 0; r < d.rows; ++r) {
      if (d.merge) {
        const int Cq = d.C / 4, W2 = d.W / 2, H2 = d.H / 2;
        const int x = r % W2, y = (r / W2) % H2;
        const long long n = r / (W2 * H2);
        for (int s = 0; s < 4; ++s) {
          const int dy = s & 1, dx = s >> 1;
          const long long src = ((n * d.H + 2 * y + dy) * d.W + 2 * x + dx) * (long long)Cq;
          for (int c = 0; c < Cq; ++c) row[s * Cq + c] = ldx(d.in, src + c, d.dtype & SVX_DT_IN_BF16);
        }
      } else {
        for (int c = 0; c < d.C; ++c) row[c] = ldx(d.in, (long long)r * d.C + c, d.dtype & SVX_DT_IN_BF16);
      }
      double mean = 0, var = 0;
      for (int c = 0; c < d.C; ++c) mean += row[c];
      mean /= d.C;
      for (int c = 0; c < d.C; ++c) var += (row[c] - mean) * (row[c] - mean);
      var /= d.C;
      const float rstd = 1.f / sqrtf((float)var + d.eps);
      for (int c = 0; c < d.C; ++c)
        stx(d.out, (long long)r * d.C + c, rnd((row[c] - (float)mean) * rstd * d.gamma[c] + d.beta[c], d.round_tf32),
            d.dtype & SVX_DT_OUT_BF16);
    }
  }
  return 0;
}

int lnsample_launch(const svx_lnsample_desc& d, void*) {
#pragma omp parallel for
  for (int n = 0; n < d.N; ++n) {
    std::vector<float> x(d.L);
    for (int i = 0; i < d.L; ++i) x[i] = ldx(d.in, (long long)n * d.L + i, d.dtype & SVX_DT_IN_BF16);
    double mean = 0, var = 0;
    for (int i = 0; i < d.L; ++i) mean += x[i];
    mean /= d.L;
    for (int i = 0; i < d.L; ++i) var += (x[i] - mean) * (x[i] - mean);
    var /= d.L;
    const float rstd = 1.f / sqrtf((float)var + d.eps);
    for (int i = 0; i < d.L; ++i)
      stx(d.out, (long long)n * d.L + i, rnd((x[i] - (float)mean) * rstd * d.gamma[i] + d.beta[i], d.round_tf32),
          d.dtype & SVX_DT_OUT_BF16);
  }
  return 0;
}

int winattn_launch(const svx_winattn_desc& d, void*) {
  SVX_REQUIRE(d.H % 7 == 0 && d.W % 7 == 0 && d.C == d.heads * 32, "window_attention: bad shape");
  const int nwy = d.H / 7, nwx = d.W / 7;
  const long long nwin = (long long)d.N * nwy * nwx;
#pragma omp parallel for
  for (long long w = 0; w < nwin; ++w) {
    const int wx = w % nwx, wy = (w / nwx) % nwy;
    const long long n = w / (nwx * nwy);
    long long tok[49];
    int reg[49];
    for (int t = 0; t < 49; ++t) {
      const int py = wy * 7 + t / 7, px = wx * 7 + t % 7;
      tok[t] = (n * d.H + (py + d.shift) % d.H) * d.W + (px + d.shift) % d.W;
      reg[t] = 0;
      if (d.shift > 0) {
        const int ry = py < d.H - 7 ? 0 : (py < d.H - d.shift ? 1 : 2);
        const int rx = px < d.W - 7 ? 0 : (px < d.W - d.shift ? 1 : 2);
        reg[t] = ry * 3 + rx;
      }
    }
    for (int h = 0; h < d.heads; ++h)
      for (int i = 0; i < 49; ++i) {
        const bool bfi = d.dtype & SVX_DT_IN_BF16;
        const long long q = tok[i] * 3 * d.C + h * 32;
        float s[49], mx = -INFINITY;
        for (int j = 0; j < 49; ++j) {
          const long long k = tok[j] * 3 * d.C + d.C + h * 32;
          float acc = 0.f;
          for (int e = 0; e < 32; ++e) acc += ldx(d.qkv, q + e, bfi) * d.scale * ldx(d.qkv, k + e, bfi);
          acc += d.bias[((long long)h * 49 + i) * 49 + j];
          if (reg[i] != reg[j]) acc += -100.f;
          s[j] = acc;
          mx = std::max(mx, acc);
        }
        float den = 0.f;
        for (int j = 0; j < 49; ++j) { s[j] = expf(s[j] - mx); den += s[j]; }
        for (int e = 0; e < 32; ++e) {
          float acc = 0.f;
          for (int j = 0; j < 49; ++j) acc += s[j] / den * ldx(d.qkv, tok[j] * 3 * d.C + 2 * d.C + h * 32 + e, bfi);
          stx(d.out, tok[i] * d.C + h * 32 + e, rnd(acc, d.round_tf32), d.dtype & SVX_DT_OUT_BF16);
        }
      }
  }
  return 0;
}

int dwconv_launch(const svx_dwconv_desc& d, void*) {
  for (long long n = 0; n < d.N; ++n)
    for (int oy = 0; oy < d.OH; ++oy)
      for (int ox = 0; ox < d.OW; ++ox)
        for (int c = 0; c < d.C; ++c) {
          float acc = d.bias ? d.bias[c] : 0.f;
          for (int ky = 0; ky < d.k; ++ky)
            for (int kx = 0; kx < d.k; ++kx)
              acc += ldx(d.in, ((n * d.H + oy * d.k + ky) * d.W + ox * d.k + kx) * (long long)d.C + c, d.dtype & SVX_DT_IN_BF16) *
                     d.w[(ky * d.k + kx) * d.C + c];
          stx(d.out, ((n * d.OH + oy) * d.OW + ox) * (long long)d.C + c, rnd(acc, d.round_tf32), d.dtype & SVX_DT_OUT_BF16);
        }
  return 0;
}

int viewattn_launch(const svx_viewattn_desc& d, void*) {
  const int hd = d.R / d.heads, R3 = 3 * d.R;
  for (long long b = 0; b < d.B; ++b)
    for (int h = 0; h < d.heads; ++h) {
      const long long b0 = b * d.V * (long long)d.P * R3;
      auto base = [&](long long i) { return ldx(d.qkv, b0 + i, d.dtype & SVX_DT_IN_BF16); };
      std::vector<float> sc(d.V * d.V);
      for (int v1 = 0; v1 < d.V; ++v1) {
        float mx = -INFINITY;
        for (int v2 = 0; v2 < d.V; ++v2) {
          float acc = 0.f;
          for (int pos = 0; pos < d.P; ++pos)
            for (int e = 0; e < hd; ++e)
              acc += base(((long long)v1 * d.P + pos) * R3 + h * hd + e) *
                     base(((long long)v2 * d.P + pos) * R3 + d.R + h * hd + e);
          sc[v1 * d.V + v2] = acc * d.scale;
          mx = std::max(mx, acc * d.scale);
        }
        float den = 0.f;
        for (int v2 = 0; v2 < d.V; ++v2) { sc[v1 * d.V + v2] = expf(sc[v1 * d.V + v2] - mx); den += sc[v1 * d.V + v2]; }
        for (int v2 = 0; v2 < d.V; ++v2) sc[v1 * d.V + v2] /= den;
      }
      for (int v1 = 0; v1 < d.V; ++v1)
        for (int pos = 0; pos < d.P; ++pos)
          for (int e = 0; e < hd; ++e) {
            float acc = 0.f;
            for (int v2 = 0; v2 < d.V; ++v2)
              acc += sc[v1 * d.V + v2] * base(((long long)v2 * d.P + pos) * R3 + 2 * d.R + h * hd + e);
            stx(d.out, ((b * d.V + v1) * d.P + pos) * (long long)d.R + h * hd + e, rnd(acc, d.round_tf32), d.dtype & SVX_DT_OUT_BF16);
          }
    }
  return 0;
}

int bilinear_launch(const svx_bilinear_desc& d, void*) {
  const float sy = (float)d.IH / d.OH, sx = (float)d.IW / d.OW;
  for (long long n = 0; n < d.N; ++n)
    for (int oy = 0; oy < d.OH; ++oy)
      for (int ox = 0; ox < d.OW; ++ox) {
        const float fy = std::max((oy + 0.5f) * sy - 0.5f, 0.f), fx = std::max((ox + 0.5f) * sx - 0.5f, 0.f);
        const int y0 = std::min((int)fy, d.IH - 1), x0 = std::min((int)fx, d.IW - 1);
        const int y1 = std::min(y0 + 1, d.IH - 1), x1 = std::min(x0 + 1, d.IW - 1);
        const float ly = fy - y0, lx = fx - x0;
        for (int c = 0; c < d.C; ++c) {
          auto at = [&](int y, int x) { return ldx(d.in, ((n * d.IH + y) * d.IW + x) * (long long)d.C + c, d.dtype & SVX_DT_IN_BF16); };
          const long long o = ((n * d.OH + oy) * d.OW + ox) * (long long)d.C + c;
          float v = (1 - ly) * (1 - lx) * at(y0, x0) + (1 - ly) * lx * at(y0, x1) + ly * (1 - lx) * at(y1, x0) +
                    ly * lx * at(y1, x1);
          if (d.skip) v += ldx(d.skip, o, d.dtype & SVX_DT_IN_BF16);
          stx(d.out, o, rnd(v, d.round_tf32), d.dtype & SVX_DT_OUT_BF16);
        }
      }
  return 0;
}

int conv3to1_launch(const svx_conv3to1_desc& d, void*) {
  SVX_REQUIRE(d.in && d.w && d.out && d.Cin >= 1 && d.Cin <= 12 && d.W == 32 && d.H % 16 == 0 && d.Cs % 4 == 0 && d.c0 % 4 == 0,
              "conv3to1: bad description");
  const long long Hp = d.H + 2, Wp = d.W + 2, Dp = d.D + 2;
#pragma omp parallel for
  for (long long r = 0; r < (long long)d.N * d.D * d.H * d.W; ++r) {
    const int w = (int)(r % d.W);
    long long t = r / d.W;
    const int h = (int)(t % d.H); t /= d.H;
    const int dd = (int)(t % d.D);
    const long long n = t / d.D;
    float acc = d.bias ? d.bias[0] : 0.f;
    for (int kd = 0; kd < 3; ++kd)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          const float* px = d.in + (((n * Dp + dd + kd) * Hp + h + kh) * Wp + w + kw) * d.Cs + d.c0;
          const float* wr = d.w + ((kd * 3 + kh) * 3 + kw) * 12;
          for (int c = 0; c < d.Cin; ++c) acc = fmaf(px[c], wr[c], acc);
        }
    d.out[r] = acc > 0.f ? acc : acc * d.slope;
  }
  return 0;
}

int mergefuse_launch(const svx_mergefuse_desc& d, void*) {
  for (long long b = 0; b < d.B; ++b)
    for (int p = 0; p < d.P; ++p) {
      if (!d.weights) {
        float s = 0.f;
        for (int v = 0; v < d.V; ++v) s += d.coarse[(b * d.V + v) * (long long)d.P + p];
        d.out[b * d.P + p] = s * (1.f / (float)d.V);
        continue;
      }
      float mx = -INFINITY;
      for (int v = 0; v < d.V; ++v) mx = std::max(mx, d.weights[(b * d.V + v) * (long long)d.P + p]);
      float den = 0.f, num = 0.f;
      for (int v = 0; v < d.V; ++v) {
        const float e = expf(d.weights[(b * d.V + v) * (long long)d.P + p] - mx);
        den += e;
        num += e * d.coarse[(b * d.V + v) * (long long)d.P + p];
      }
      d.out[b * d.P + p] = num / den;
    }
  return 0;
}

int metrics_launch(const svx_metrics_desc& d, void*) {
  SVX_REQUIRE(d.T >= 1 && d.T <= 8, "voxel_metrics: supports 1..8 thresholds");
  for (long long b = 0; b < d.B; ++b)
    for (int t = 0; t < d.T; ++t) {
      int I = 0, U = 0, FP = 0, FN = 0;
      for (int p = 0; p < d.P; ++p) {
        const float prob = 1.f / (1.f + expf(-d.logits[b * d.P + p]));
        const int v = prob >= d.prob_thresholds[t], g = d.gt[b * d.P + p] != 0.f;
        I += v & g; U += v | g; FP += v & !g; FN += !v & g;
      }
      int32_t* c = d.counts + (b * d.T + t) * 5;
      c[0] = I; c[1] = U; c[2] = I; c[3] = FP; c[4] = FN;
    }
  if (d.bce_q20)
    for (long long b = 0; b < d.B; ++b) {
      long long acc = 0;
      for (int p = 0; p < d.P; ++p) {
        const float x = d.logits[b * d.P + p];
        const float l = fmaxf(x, 0.f) - x * d.gt[b * d.P + p] + log1pf(expf(-fabsf(x)));
        acc += llrintf(l * 1048576.f);
      }
      d.bce_q20[b] = acc;
    }
  return 0;
}

int resize_launch(const svx_resize_desc& d, void*) {
  const float sy = (float)d.IH / d.OH, sx = (float)d.IW / d.OW;
  for (long long nc = 0; nc < d.NC; ++nc)
    for (int oy = 0; oy < d.OH; ++oy)
      for (int ox = 0; ox < d.OW; ++ox) {
        const float fy = std::max((oy + 0.5f) * sy - 0.5f, 0.f), fx = std::max((ox + 0.5f) * sx - 0.5f, 0.f);
        const int y0 = std::min((int)fy, d.IH - 1), x0 = std::min((int)fx, d.IW - 1);
        const int y1 = std::min(y0 + 1, d.IH - 1), x1 = std::min(x0 + 1, d.IW - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float* img = d.in + nc * (long long)d.IH * d.IW;
        d.out[(nc * d.OH + oy) * (long long)d.OW + ox] =
            (1.f - ly) * ((1.f - lx) * img[(long long)y0 * d.IW + x0] + lx * img[(long long)y0 * d.IW + x1]) +
            ly * ((1.f - lx) * img[(long long)y1 * d.IW + x0] + lx * img[(long long)y1 * d.IW + x1]);
      }
  return 0;
}

int transpose_launch(const svx_transpose_desc& d, void*) {
  for (long long n = 0; n < d.N; ++n)
    for (int p = 0; p < d.P; ++p) {
      if (d.to_channels_last) {
        long long o = n * d.P + p;
        if (d.row_w > 0) o = n * (long long)(d.P / d.row_w) * d.row_pitch + (long long)(p / d.row_w) * d.row_pitch + d.row_x0 + p % d.row_w;
        for (int c = 0; c < d.Cs; ++c)
          stx(d.out, o * (long long)d.Cs + c, c < d.C ? rnd(d.in[(n * d.C + c) * (long long)d.P + p], d.round_tf32) : 0.f,
              d.dtype & SVX_DT_OUT_BF16);
      } else {
        for (int c = 0; c < d.C; ++c)
          d.out[(n * d.C + c) * (long long)d.P + p] = rnd(ldx(d.in, (n * d.P + p) * (long long)d.Cs + c, d.dtype & SVX_DT_IN_BF16), d.round_tf32);
      }
    }
  return 0;
}

}  // namespace svx

// svx_ops.cu -- the bandwidth-bound and small CUDA-core stages of the SwinVox forward path:
// im2col for tiny channel counts, pooling, LayerNorm (row-wise and whole-sample), shifted-window
// attention, cross-view attention pieces, merger softmax-fuse, threshold/IoU counters, layout changes.
#include <cuda_runtime.h>

#include <cstdlib>

#include "svx_internal.h"
#include "svx_ptx.cuh"

namespace svx {
namespace {

constexpr int kSmCount = 148;

__device__ __forceinline__ float maybe_round(float x, int r) { return r ? round_tf32(x) : x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum for blockDim.x <= 1024; `red` is 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int grid_for(long long work, int block, int max_blocks = kSmCount * 32) {
  long long g = (work + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- im2col (tiny C): one thread = four consecutive k of one row (one 16-byte store) ------------------------
__global__ void im2col_kernel(const svx_im2col_desc d, long long total4, int K) {
  const int K4 = d.Kpad >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % K4) * 4;
    long long r = idx / K4;
    const int ow = (int)(r % d.OW); r /= d.OW;
    const int oh = (int)(r % d.OH); r /= d.OH;
    const int od = (int)(r % d.OD);
    const long long n = r / d.OD;
    const float* src = d.in + n * d.s_n;
    const int id0 = od * d.stride - d.pad_d, ih0 = oh * d.stride - d.pad_h, iw0 = ow * d.stride - d.pad_w;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + q;
      float x = 0.f;
      if (k < K) {
        const int c = k % d.C;
        int t = k / d.C;
        const int kw = t % d.KW; t /= d.KW;
        const int kh = t % d.KH;
        const int kd = t / d.KH;
        const int id = id0 + kd, ih = ih0 + kh, iw = iw0 + kw;
        if ((unsigned)id < (unsigned)d.D && (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W)
          x = __ldg(src + c * d.s_c + id * d.s_d + ih * d.s_h + iw * d.s_w);
      }
      v[q] = maybe_round(x, d.round_tf32);
    }
    reinterpret_cast<float4*>(d.out)[idx] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- im2col of a single-channel volume, one CTA per output row line (n, od, oh): the KD x KH x rowlen input region that
// the OW windows of the line share is staged in shared memory once (zero-filled outside the volume), a look-up table
// maps k -> offset inside the region, and the CTA writes the OW x Kpad block with coalesced 16-byte stores.
constexpr int kI2cRegionMax = 4096, kI2cKMax = 512;
__global__ void __launch_bounds__(256) im2col_line_kernel(const svx_im2col_desc d, int K) {
  __shared__ float region[kI2cRegionMax];
  __shared__ short lut[kI2cKMax];
  const int rowlen = (d.OW - 1) * d.stride + d.KW;
  long long line = blockIdx.x;
  const int oh = (int)(line % d.OH); line /= d.OH;
  const int od = (int)(line % d.OD);
  const long long n = line / d.OD;
  const float* src = d.in + n * d.s_n;
  for (int i = threadIdx.x; i < d.KD * d.KH * rowlen; i += blockDim.x) {
    const int x = i % rowlen, t = i / rowlen;
    const int kh = t % d.KH, kd = t / d.KH;
    const int id = od * d.stride - d.pad_d + kd, ih = oh * d.stride - d.pad_h + kh, iw = x - d.pad_w;
    float v = 0.f;
    if ((unsigned)id < (unsigned)d.D && (unsigned)ih < (unsigned)d.H && (unsigned)iw < (unsigned)d.W)
      v = __ldg(src + id * d.s_d + ih * d.s_h + iw * d.s_w);
    region[i] = maybe_round(v, d.round_tf32);
  }
  for (int k = threadIdx.x; k < d.Kpad; k += blockDim.x) {
    short o = -1;
    if (k < K) {
      const int kw = k % d.KW, t = k / d.KW;
      o = (short)(((t / d.KH) * d.KH + t % d.KH) * rowlen + kw);
    }
    lut[k] = o;
  }
  __syncthreads();
  const int K4 = d.Kpad >> 2;
  float* dst = d.out + ((n * d.OD + od) * d.OH + oh) * (long long)d.OW * d.Kpad;
  for (int i = threadIdx.x; i < d.OW * K4; i += blockDim.x) {
    const int ow = i / K4, k0 = (i - ow * K4) * 4;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = lut[k0 + q];
      v[q] = o >= 0 ? region[o + ow * d.stride] : 0.f;
    }
    *reinterpret_cast<float4*>(dst + (long long)ow * d.Kpad + k0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- pooling ------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_kernel(const svx_pool_desc d, long long total) {
  const T* in = reinterpret_cast<const T*>(d.in);
  T* out = reinterpret_cast<T*>(d.out);
  const int c4n = d.C >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % c4n);
    long long r = idx / c4n;
    const int ow = (int)(r % d.OW); r /= d.OW;
    const int oh = (int)(r % d.OH); r /= d.OH;
    const int od = (int)(r % d.OD);
    const long long n = r / d.OD;
    float4 acc = d.mode == SVX_POOL_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int kd = 0; kd < d.KD; ++kd) {
      const int id = od * d.SD - d.PD + kd;
      if ((unsigned)id >= (unsigned)d.D) continue;
      for (int kh = 0; kh < d.KH; ++kh) {
        const int ih = oh * d.SH - d.PH + kh;
        if ((unsigned)ih >= (unsigned)d.H) continue;
        for (int kw = 0; kw < d.KW; ++kw) {
          const int iw = ow * d.SW - d.PW + kw;
          if ((unsigned)iw >= (unsigned)d.W) continue;
          const float4 v = ld4g(in + (((n * d.D + id) * d.H + ih) * d.W + iw) * (long long)d.in_Cs + c4 * 4);
          if (d.mode == SVX_POOL_MAX) {
            acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y);
            acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
          } else {
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          ++cnt;
        }
      }
    }
    if (d.mode == SVX_POOL_AVG) {
      const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    acc.x = maybe_round(acc.x, d.round_tf32); acc.y = maybe_round(acc.y, d.round_tf32);
    acc.z = maybe_round(acc.z, d.round_tf32); acc.w = maybe_round(acc.w, d.round_tf32);
    st4(out + (((n * d.OD + od) * d.OH + oh) * d.OW + ow) * (long long)d.out_Cs + c4 * 4, acc);
  }
}

// ---- MaxPool2d(3, stride 2, pad 1), channels-last (the ResNet stem's pool): a thread marches down the output rows of one
// (column, 4-channel group); per output row it reads two NEW input rows (3 float4 each, horizontal max first) and reuses
// the horizontal max of the row shared with the previous output row: 6 loads per output instead of 9.
template <typename T>
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const svx_pool_desc d, int cols_per_block) {
  const T* in = reinterpret_cast<const T*>(d.in);
  T* out = reinterpret_cast<T*>(d.out);
  const int c4n = d.C >> 2;
  const int c4 = threadIdx.x % c4n;
  const int ow = blockIdx.x * cols_per_block + threadIdx.x / c4n;
  const long long n = blockIdx.y;
  if (ow >= d.OW) return;
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  auto hmax = [&](int ih) -> float4 {
    if ((unsigned)ih >= (unsigned)d.H) return ninf;
    const T* row = in + ((n * d.H + ih) * (long long)d.W) * d.in_Cs + c4 * 4;
    float4 m = ninf;
#pragma unroll
    for (int k = -1; k <= 1; ++k) {
      const int iw = 2 * ow + k;
      if ((unsigned)iw < (unsigned)d.W) {
        const float4 v = ld4g(row + (long long)iw * d.in_Cs);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    return m;
  };
  float4 prev = hmax(-1);
  for (int oh = 0; oh < d.OH; ++oh) {
    const float4 a = hmax(2 * oh), b = hmax(2 * oh + 1);
    float4 o;
    o.x = maybe_round(fmaxf(prev.x, fmaxf(a.x, b.x)), d.round_tf32); o.y = maybe_round(fmaxf(prev.y, fmaxf(a.y, b.y)), d.round_tf32);
    o.z = maybe_round(fmaxf(prev.z, fmaxf(a.z, b.z)), d.round_tf32); o.w = maybe_round(fmaxf(prev.w, fmaxf(a.w, b.w)), d.round_tf32);
    st4(out + ((n * d.OH + oh) * (long long)d.OW + ow) * d.out_Cs + c4 * 4, o);
    prev = b;
  }
}

// ---- row LayerNorm: one warp per row ---------------------------------------------------------------
template <typename T>
__global__ void lnrows_kernel(const svx_lnrows_desc d) {
  const T* in = reinterpret_cast<const T*>(d.in);
  T* out = reinterpret_cast<T*>(d.out);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int C = d.C;
  const int Cq = C >> 2;  // merge: channels of one source pixel
  const float invC = 1.f / (float)C;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < d.rows;
       row += (long long)gridDim.x * wpb) {
    const T* src[4];
    if (d.merge) {
      const int W2 = d.W >> 1, H2 = d.H >> 1;
      const int x = (int)(row % W2);
      long long t = row / W2;
      const int y = (int)(t % H2);
      const long long n = t / H2;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int dy = s & 1, dx = s >> 1;  // timm order: (h0,w0) (h1,w0) (h0,w1) (h1,w1)
        src[s] = in + ((n * d.H + 2 * y + dy) * d.W + 2 * x + dx) * (long long)Cq;
      }
    } else {
      src[0] = in + row * (long long)C;
      src[1] = src[2] = src[3] = src[0];
    }
    auto load4 = [&](int c) -> float4 {
      if (d.merge) {
        const int s = c / Cq;
        return ld4g(src[s] + (c - s * Cq));
      }
      return ld4g(src[0] + c);
    };
    float sum = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = load4(c);
      sum += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(sum) * invC;
    float sq = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = load4(c);
      const float a = v.x - mean, b = v.y - mean, e = v.z - mean, f = v.w - mean;
      sq += (a * a + b * b) + (e * e + f * f);
    }
    const float rstd = rsqrtf(warp_sum(sq) * invC + d.eps);
    T* dst = out + row * (long long)C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = load4(c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(d.gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(d.beta + c));
      float4 o;
      o.x = maybe_round((v.x - mean) * rstd * g.x + b.x, d.round_tf32);
      o.y = maybe_round((v.y - mean) * rstd * g.y + b.y, d.round_tf32);
      o.z = maybe_round((v.z - mean) * rstd * g.z + b.z, d.round_tf32);
      o.w = maybe_round((v.w - mean) * rstd * g.w + b.w, d.round_tf32);
      st4(dst + c, o);
    }
  }
}

// ---- row LayerNorm, register-resident: one warp normalises R rows at a time, each row read from HBM exactly once
// (NV float4 per lane) so R*NV independent 16-byte loads are in flight per lane; statistics are the exact two-pass
// mean / variance computed from the registers.
template <int NV, int R, typename T>
__global__ void __launch_bounds__(256) lnrows_reg_kernel(const svx_lnrows_desc d) {
  const T* in = reinterpret_cast<const T*>(d.in);
  T* out = reinterpret_cast<T*>(d.out);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int C = d.C;
  const int Cq = C >> 2;  // merge: channels of one source pixel
  const float invC = 1.f / (float)C;
  float4 gm[NV <= 6 ? NV : 1], bt[NV <= 6 ? NV : 1];
  if constexpr (NV <= 6) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane * 4 + i * 128;
      gm[i] = c < C ? __ldg(reinterpret_cast<const float4*>(d.gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      bt[i] = c < C ? __ldg(reinterpret_cast<const float4*>(d.beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const long long groups = (d.rows + R - 1) / R;
  for (long long grp = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); grp < groups; grp += (long long)gridDim.x * wpb) {
    float4 v[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
      const bool rok = row < d.rows;
      long long base = row * (long long)C;
      int x = 0, y = 0;
      long long n = 0;
      if (d.merge) {
        const int W2 = d.W >> 1, H2 = d.H >> 1;
        x = (int)(row % W2);
        const long long t = row / W2;
        y = (int)(t % H2);
        n = t / H2;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane * 4 + i * 128;
        const T* src;
        if (d.merge) {
          const int sidx = c / Cq;
          const int dy = sidx & 1, dx = sidx >> 1;  // timm order: (h0,w0) (h1,w0) (h0,w1) (h1,w1)
          src = in + ((n * d.H + 2 * y + dy) * d.W + 2 * x + dx) * (long long)Cq + (c - sidx * Cq);
        } else {
          src = in + base + c;
        }
        v[r][i] = (rok && c < C) ? ld4g(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
      const float mean = warp_sum(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane * 4 + i * 128 < C) {
          const float a = v[r][i].x - mean, b = v[r][i].y - mean, e = v[r][i].z - mean, f = v[r][i].w - mean;
          sq += (a * a + b * b) + (e * e + f * f);
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) * invC + d.eps);
      if (row < d.rows) {
        T* dst = out + row * (long long)C;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane * 4 + i * 128;
          if (c < C) {
            float4 g, b;
            if constexpr (NV <= 6) { g = gm[i]; b = bt[i]; }
            else {
              g = __ldg(reinterpret_cast<const float4*>(d.gamma + c));
              b = __ldg(reinterpret_cast<const float4*>(d.beta + c));
            }
            float4 o;
            o.x = maybe_round((v[r][i].x - mean) * rstd * g.x + b.x, d.round_tf32);
            o.y = maybe_round((v[r][i].y - mean) * rstd * g.y + b.y, d.round_tf32);
            o.z = maybe_round((v[r][i].z - mean) * rstd * g.z + b.z, d.round_tf32);
            o.w = maybe_round((v[r][i].w - mean) * rstd * g.w + b.w, d.round_tf32);
            st4(dst + c, o);
          }
        }
      }
    }
  }
}

// ---- row LayerNorm on bf16 rows: the same register-resident scheme with 16-byte accesses (eight bf16 per lane and
// vector, NV8 vectors cover NV8 * 256 channels): the fp32-shaped kernel above moves half the bytes per instruction on
// bf16 data and ends up instruction-bound (measured slower than on fp32).
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
template <int NV8, int R>
__global__ void __launch_bounds__(256) lnrows_bf16_kernel(const svx_lnrows_desc d) {
  const bf16_t* in = reinterpret_cast<const bf16_t*>(d.in);
  bf16_t* out = reinterpret_cast<bf16_t*>(d.out);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int C = d.C;
  const int Cq = C >> 2;  // merge: channels of one source pixel
  const float invC = 1.f / (float)C;
  const long long groups = (d.rows + R - 1) / R;
  for (long long grp = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); grp < groups; grp += (long long)gridDim.x * wpb) {
    uint4 raw[R][NV8];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
      const bool rok = row < d.rows;
      int x = 0, y = 0;
      long long n = 0;
      if (d.merge) {
        const int W2 = d.W >> 1, H2 = d.H >> 1;
        x = (int)(row % W2);
        const long long t = row / W2;
        y = (int)(t % H2);
        n = t / H2;
      }
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        const int c = lane * 8 + i * 256;
        const bf16_t* src;
        if (d.merge) {
          const int sidx = c / Cq;
          const int dy = sidx & 1, dx = sidx >> 1;  // timm order: (h0,w0) (h1,w0) (h0,w1) (h1,w1)
          src = in + ((n * d.H + 2 * y + dy) * d.W + 2 * x + dx) * (long long)Cq + (c - sidx * Cq);
        } else {
          src = in + row * (long long)C + c;
        }
        raw[r][i] = (rok && c < C) ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
      float v[NV8][8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        unpack8(raw[r][i], v[i]);
        sum += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
      }
      const float mean = warp_sum(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        if (lane * 8 + i * 256 < C) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float a = v[i][e] - mean; sq = fmaf(a, a, sq); }
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) * invC + d.eps);
      if (row < d.rows) {
        bf16_t* dst = out + row * (long long)C;
#pragma unroll
        for (int i = 0; i < NV8; ++i) {
          const int c = lane * 8 + i * 256;
          if (c < C) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(d.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(d.gamma + c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(d.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(d.beta + c + 4));
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, g[e], b[e]);
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                            pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
      }
    }
  }
}

// ---- whole-sample LayerNorm: one CTA per sample ----------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) lnsample_kernel(const svx_lnsample_desc d) {
  __shared__ float red[32];
  const long long L = d.L;
  const int L4 = d.L >> 2;
  for (int n = blockIdx.x; n < d.N; n += gridDim.x) {
    const T* x = reinterpret_cast<const T*>(d.in) + n * L;
    float sum = 0.f;
    for (int i = threadIdx.x; i < L4; i += blockDim.x) {
      const float4 v = ld4(x + 4 * i);
      sum += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = block_sum(sum, red) / (float)L;
    float sq = 0.f;
    for (int i = threadIdx.x; i < L4; i += blockDim.x) {
      const float4 v = ld4(x + 4 * i);
      const float a = v.x - mean, b = v.y - mean, e = v.z - mean, f = v.w - mean;
      sq += (a * a + b * b) + (e * e + f * f);
    }
    const float rstd = rsqrtf(block_sum(sq, red) / (float)L + d.eps);
    T* y = reinterpret_cast<T*>(d.out) + n * L;
    const float4* g4 = reinterpret_cast<const float4*>(d.gamma);
    const float4* b4 = reinterpret_cast<const float4*>(d.beta);
    for (int i = threadIdx.x; i < L4; i += blockDim.x) {
      const float4 v = ld4(x + 4 * i);
      const float4 g = __ldg(g4 + i);
      const float4 b = __ldg(b4 + i);
      float4 o;
      o.x = maybe_round((v.x - mean) * rstd * g.x + b.x, d.round_tf32);
      o.y = maybe_round((v.y - mean) * rstd * g.y + b.y, d.round_tf32);
      o.z = maybe_round((v.z - mean) * rstd * g.z + b.z, d.round_tf32);
      o.w = maybe_round((v.w - mean) * rstd * g.w + b.w, d.round_tf32);
      st4(y + 4 * i, o);
    }
  }
}

// ---- whole-sample LayerNorm on a thread-block cluster: the eight CTAs of a cluster share one sample, every thread
// keeps its NV float4 of the sample in registers (ONE HBM read), the two statistics are reduced across the cluster
// through distributed shared memory (each CTA publishes its partial, all read the eight partials in rank order).
constexpr int kLnCluster = 8;
__device__ __forceinline__ float dsmem_read(const float* local, unsigned rank) {
  unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(local)), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NV, typename T>
__global__ void __launch_bounds__(1024, 1) lnsample_cluster_kernel(const svx_lnsample_desc d) {
  __shared__ float red[32];
  __shared__ float part[2];
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int n = blockIdx.x / kLnCluster;   // grid = N clusters (exactly one sample each)
  const long long L = d.L;
  const int L4 = d.L >> 2;
  const T* x = reinterpret_cast<const T*>(d.in) + n * L;
  const int t0 = rank * 1024 + threadIdx.x, stride = kLnCluster * 1024;
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = t0 + k * stride;
    v[k] = i < L4 ? ld4g(x + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) part[0] = sum;
  cluster_barrier();
  float tot = 0.f;
#pragma unroll
  for (unsigned r = 0; r < kLnCluster; ++r) tot += dsmem_read(&part[0], r);
  const float mean = tot / (float)L;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (t0 + k * stride < L4) {
      const float a = v[k].x - mean, b = v[k].y - mean, e = v[k].z - mean, f = v[k].w - mean;
      sq += (a * a + b * b) + (e * e + f * f);
    }
  }
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) part[1] = sq;
  cluster_barrier();
  float tsq = 0.f;
#pragma unroll
  for (unsigned r = 0; r < kLnCluster; ++r) tsq += dsmem_read(&part[1], r);
  const float rstd = rsqrtf(tsq / (float)L + d.eps);
  T* y = reinterpret_cast<T*>(d.out) + n * L;
  const float4* g4 = reinterpret_cast<const float4*>(d.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(d.beta);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = t0 + k * stride;
    if (i < L4) {
      const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
      float4 o;
      o.x = maybe_round((v[k].x - mean) * rstd * g.x + b.x, d.round_tf32);
      o.y = maybe_round((v[k].y - mean) * rstd * g.y + b.y, d.round_tf32);
      o.z = maybe_round((v[k].z - mean) * rstd * g.z + b.z, d.round_tf32);
      o.w = maybe_round((v[k].w - mean) * rstd * g.w + b.w, d.round_tf32);
      st4(y + 4 * i, o);
    }
  }
  cluster_barrier();   // nobody leaves while a peer may still read its partials
}
// ---- shifted-window attention: the kernel lives in svx_winattn.cu (tcgen05 / TMEM); only the argument checks are here
constexpr int WS = 7, WT = 49, HD = 32;

// ---- depthwise k=s conv, channels-last ---------------------------------------------------------------
template <typename T>
__global__ void dwconv_kernel(const svx_dwconv_desc d, long long total) {
  const T* in = reinterpret_cast<const T*>(d.in);
  T* out = reinterpret_cast<T*>(d.out);
  const int c4n = d.C >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % c4n);
    long long r = idx / c4n;
    const int ox = (int)(r % d.OW); r /= d.OW;
    const int oy = (int)(r % d.OH);
    const long long n = r / d.OH;
    float4 acc = d.bias ? __ldg(reinterpret_cast<const float4*>(d.bias + c4 * 4)) : make_float4(0, 0, 0, 0);
    for (int ky = 0; ky < d.k; ++ky)
      for (int kx = 0; kx < d.k; ++kx) {
        const float4 v = ld4g(in + ((n * d.H + oy * d.k + ky) * d.W + ox * d.k + kx) * (long long)d.C + c4 * 4);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(d.w + (ky * d.k + kx) * d.C + c4 * 4));
        acc.x = fmaf(v.x, wv.x, acc.x); acc.y = fmaf(v.y, wv.y, acc.y);
        acc.z = fmaf(v.z, wv.z, acc.z); acc.w = fmaf(v.w, wv.w, acc.w);
      }
    acc.x = maybe_round(acc.x, d.round_tf32); acc.y = maybe_round(acc.y, d.round_tf32);
    acc.z = maybe_round(acc.z, d.round_tf32); acc.w = maybe_round(acc.w, d.round_tf32);
    st4(out + ((n * d.OH + oy) * d.OW + ox) * (long long)d.C + c4 * 4, acc);
  }
}

// ---- attention over the view axis: one CTA per (object, head) ----------------------------------------
constexpr int kMaxViews = 32;
template <typename T>
__global__ void __launch_bounds__(256) viewattn_kernel(const svx_viewattn_desc d) {
  __shared__ float sc[kMaxViews * kMaxViews];
  const int head = blockIdx.x % d.heads;
  const long long b = blockIdx.x / d.heads;
  const int V = d.V, P = d.P, R = d.R, hd = R / d.heads;
  const int R3 = 3 * R;
  const int L = P * hd;  // dot-product length
  const T* base = reinterpret_cast<const T*>(d.qkv) + b * V * (long long)P * R3;
  T* out = reinterpret_cast<T*>(d.out);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int pair = warp; pair < V * V; pair += nw) {
    const int v1 = pair / V, v2 = pair % V;
    float acc = 0.f;
    for (int e = lane; e < L; e += 32) {
      const int pos = e / hd, dd = e % hd;
      const float qv = ld1(base + ((long long)v1 * P + pos) * R3 + head * hd + dd);
      const float kv = ld1(base + ((long long)v2 * P + pos) * R3 + R + head * hd + dd);
      acc = fmaf(qv, kv, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) sc[v1 * kMaxViews + v2] = acc * d.scale;
  }
  __syncthreads();
  for (int v1 = warp; v1 < V; v1 += nw) {
    float x = lane < V ? sc[v1 * kMaxViews + lane] : -INFINITY;
    const float mx = warp_max(x);
    const float e = lane < V ? __expf(x - mx) : 0.f;
    const float den = warp_sum(e);
    if (lane < V) sc[v1 * kMaxViews + lane] = e / den;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < V * L; idx += blockDim.x) {
    const int v1 = idx / L, e = idx % L;
    const int pos = e / hd, dd = e % hd;
    float acc = 0.f;
    for (int v2 = 0; v2 < V; ++v2)
      acc = fmaf(sc[v1 * kMaxViews + v2], ld1(base + ((long long)v2 * P + pos) * R3 + 2 * R + head * hd + dd), acc);
    st1(out + ((b * V + v1) * P + pos) * (long long)R + head * hd + dd, maybe_round(acc, d.round_tf32));
  }
}

// ---- bilinear resize (align_corners=False) + skip ------------------------------------------------------
template <typename T>
__global__ void bilinear_kernel(const svx_bilinear_desc d, long long total) {
  const T* in = reinterpret_cast<const T*>(d.in);
  const T* skip = reinterpret_cast<const T*>(d.skip);
  T* out = reinterpret_cast<T*>(d.out);
  const int c4n = d.C >> 2;
  const float sy = (float)d.IH / (float)d.OH, sx = (float)d.IW / (float)d.OW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % c4n);
    long long r = idx / c4n;
    const int ox = (int)(r % d.OW); r /= d.OW;
    const int oy = (int)(r % d.OH);
    const long long n = r / d.OH;
    float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, d.IH - 1), x0 = min((int)fx, d.IW - 1);
    const int y1 = min(y0 + 1, d.IH - 1), x1 = min(x0 + 1, d.IW - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    auto at = [&](int y, int x) {
      return ld4g(in + ((n * d.IH + y) * d.IW + x) * (long long)d.C + c4 * 4);
    };
    const float4 a = at(y0, x0), b = at(y0, x1), c = at(y1, x0), e = at(y1, x1);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const long long o = ((n * d.OH + oy) * d.OW + ox) * (long long)d.C + c4 * 4;
    const float4 s = skip ? ld4g(skip + o) : make_float4(0, 0, 0, 0);
    float4 v;
    v.x = maybe_round(w00 * a.x + w01 * b.x + w10 * c.x + w11 * e.x + s.x, d.round_tf32);
    v.y = maybe_round(w00 * a.y + w01 * b.y + w10 * c.y + w11 * e.y + s.y, d.round_tf32);
    v.z = maybe_round(w00 * a.z + w01 * b.z + w10 * c.z + w11 * e.z + s.z, d.round_tf32);
    v.w = maybe_round(w00 * a.w + w01 * b.w + w10 * c.w + w11 * e.w + s.w, d.round_tf32);
    st4(out + o, v);
  }
}

// ---- merger: per-voxel softmax over views and weighted fusion (merger.py:98-104) -------------------------
__global__ void mergefuse_kernel(const svx_mergefuse_desc d, long long total4) {
  const int P4 = d.P >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / P4;
    const int p4 = (int)(idx % P4);
    const float4* c = reinterpret_cast<const float4*>(d.coarse + b * d.V * (long long)d.P) + p4;
    if (!d.weights) {   // no merger (core/test.py:125-126): torch.mean(generated_volume, dim=1)
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int v = 0; v < d.V; ++v) {
        const float4 g = __ldg(c + (long long)v * P4);
        s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
      }
      const float inv = 1.f / (float)d.V;
      reinterpret_cast<float4*>(d.out + b * (long long)d.P)[p4] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
      continue;
    }
    const float4* w = reinterpret_cast<const float4*>(d.weights + b * d.V * (long long)d.P) + p4;
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int v = 0; v < d.V; ++v) {
      const float4 x = __ldg(w + (long long)v * P4);
      mx.x = fmaxf(mx.x, x.x); mx.y = fmaxf(mx.y, x.y); mx.z = fmaxf(mx.z, x.z); mx.w = fmaxf(mx.w, x.w);
    }
    float4 den = make_float4(0, 0, 0, 0), num = make_float4(0, 0, 0, 0);
    for (int v = 0; v < d.V; ++v) {
      const float4 x = __ldg(w + (long long)v * P4);
      const float4 g = __ldg(c + (long long)v * P4);
      const float ex = __expf(x.x - mx.x), ey = __expf(x.y - mx.y), ez = __expf(x.z - mx.z), ew = __expf(x.w - mx.w);
      den.x += ex; den.y += ey; den.z += ez; den.w += ew;
      num.x = fmaf(ex, g.x, num.x); num.y = fmaf(ey, g.y, num.y);
      num.z = fmaf(ez, g.z, num.z); num.w = fmaf(ew, g.w, num.w);
    }
    reinterpret_cast<float4*>(d.out + b * (long long)d.P)[p4] =
        make_float4(num.x / den.x, num.y / den.y, num.z / den.z, num.w / den.w);
  }
}

// ---- Conv3d(Cin <= 12 -> 1, k3, p1) + LeakyReLU in fp32 on the CUDA cores (merger layer6) -------------------------------
// One CTA marches along depth over a C31_ROWS-row x 32-column tile of the (h, w) plane: three input planes ((ROWS + 2) x 34
// voxels x 12 channels) ring in shared memory, each plane is loaded from HBM once per tile (measured on B200,
// profiles/r2_conv3to1_variants.txt: 8-row tiles with four rows per thread, i.e. three 64-thread CTAs per SM, 0.243 ms per
// merger layer6 at 64 x 3 views against 0.280 ms for 16-row tiles; two or one row per thread lose to shared-memory
// bandwidth: 0.31 - 0.48 ms); a thread owns C31_RPT consecutive ROWS of
// one column, so every staged voxel it reads (3 float4) feeds up to three kh taps of up to four outputs from registers,
// and the 32 lanes of a warp read 32 adjacent voxels (48 bytes apart: bank-conflict free).
#ifndef C31_RPT
#define C31_RPT 4               // output rows per thread (threads per CTA = C31_ROWS / C31_RPT * 32)
#endif
#ifndef C31_ROWS
#define C31_ROWS 8              // rows of the (h, w) tile of one CTA (16: 118 KB of planes, one CTA per SM; 8: 65 KB, three)
#endif
constexpr int C31_TH = C31_ROWS, C31_TW = 32, C31_THREADS = (C31_TH / C31_RPT) * C31_TW;
constexpr int C31_PLANE = (C31_TH + 2) * (C31_TW + 2);          // staged voxels per plane
__global__ void __launch_bounds__(C31_THREADS) conv3to1_kernel(const svx_conv3to1_desc d) {
  extern __shared__ float4 c31_smem[];                          // [4 planes][C31_PLANE][3 float4] + weights [27][3]
  float4* wsm = c31_smem + 4 * C31_PLANE * 3;
  const int tiles_h = d.H / C31_TH;
  const int n = blockIdx.x / tiles_h, h0 = (blockIdx.x % tiles_h) * C31_TH;
  const int Hp = d.H + 2, Wp = d.W + 2;
  const int ty = (threadIdx.x / C31_TW) * C31_RPT, tx = threadIdx.x % C31_TW;   // rows ty..ty+RPT-1 of column tx
  for (int i = threadIdx.x; i < 27 * 3; i += C31_THREADS) wsm[i] = __ldg(reinterpret_cast<const float4*>(d.w) + i);
  const float bias = d.bias ? __ldg(d.bias) : 0.f;
  auto load_plane = [&](int dp) {   // padded depth index dp -> ring slot dp % 4, asynchronously (one cp.async group)
    if (dp < d.D + 2) {
      const float* src = d.in + (((long long)n * (d.D + 2) + dp) * Hp + h0) * (long long)Wp * d.Cs + d.c0;
      const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(c31_smem + (dp % 4) * C31_PLANE * 3));
      for (int i = threadIdx.x; i < C31_PLANE * 3; i += C31_THREADS) {
        const int v = i / 3, q = i - v * 3;                     // the tile's rows are contiguous in the padded volume
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + (long long)v * d.Cs + 4 * q) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_plane(0);
  load_plane(1);
  load_plane(2);
  for (int dd = 0; dd < d.D; ++dd) {
    load_plane(dd + 3);                                         // lands while depths dd .. dd+? compute
    asm volatile("cp.async.wait_group 1;" ::: "memory");        // planes dd, dd+1, dd+2 have landed
    __syncthreads();
    float acc[C31_RPT];
#pragma unroll
    for (int r = 0; r < C31_RPT; ++r) acc[r] = bias;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float4* pl = c31_smem + ((dd + kd) % 4) * C31_PLANE * 3;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* col = pl + (ty * (C31_TW + 2) + tx + kw) * 3;
#pragma unroll
        for (int j = 0; j < C31_RPT + 2; ++j) {                  // RPT + 2 staged rows feed RPT outputs x three kh taps
          const float4* vx = col + j * (C31_TW + 2) * 3;
          const float4 a = vx[0], b = vx[1], c = vx[2];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const int o = j - kh;
            if (o < 0 || o >= C31_RPT) continue;
            const float4* wk = wsm + ((kd * 3 + kh) * 3 + kw) * 3;
            const float4 wa = wk[0], wb = wk[1], wc = wk[2];
            float s = acc[o];
            s = fmaf(a.x, wa.x, s); s = fmaf(a.y, wa.y, s); s = fmaf(a.z, wa.z, s); s = fmaf(a.w, wa.w, s);
            s = fmaf(b.x, wb.x, s); s = fmaf(b.y, wb.y, s); s = fmaf(b.z, wb.z, s); s = fmaf(b.w, wb.w, s);
            s = fmaf(c.x, wc.x, s); s = fmaf(c.y, wc.y, s); s = fmaf(c.z, wc.z, s); s = fmaf(c.w, wc.w, s);
            acc[o] = s;
          }
        }
      }
    }
    float* orow = d.out + (((long long)n * d.D + dd) * d.H + h0 + ty) * d.W + tx;
#pragma unroll
    for (int r = 0; r < C31_RPT; ++r) orow[r * d.W] = acc[r] > 0.f ? acc[r] : acc[r] * d.slope;
    __syncthreads();   // the slot of plane dd is overwritten by the load issued two iterations from now
  }
}

// ---- sigmoid / threshold / I,U,TP,FP,FN counters (core/test.py:141-164) -----------------------------------
constexpr int kMaxThresh = 8;
__global__ void __launch_bounds__(256) metrics_kernel(const svx_metrics_desc d, int chunks) {
  __shared__ int sacc[kMaxThresh * 5];
  __shared__ unsigned long long sbce;
  if (threadIdx.x == 0) sbce = 0ull;
  long long bce = 0;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  for (int i = threadIdx.x; i < kMaxThresh * 5; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  float th[kMaxThresh];
#pragma unroll
  for (int t = 0; t < kMaxThresh; ++t) th[t] = t < d.T ? __ldg(d.prob_thresholds + t) : 2.f;
  int cI[kMaxThresh], cU[kMaxThresh], cFP[kMaxThresh], cFN[kMaxThresh];
#pragma unroll
  for (int t = 0; t < kMaxThresh; ++t) cI[t] = cU[t] = cFP[t] = cFN[t] = 0;
  const int per = (d.P + chunks - 1) / chunks;
  const int beg = chunk * per, end = min(d.P, beg + per);
  const float* lg = d.logits + (long long)b * d.P;
  const float* gt = d.gt + (long long)b * d.P;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float x = lg[i];
    const float prob = 1.f / (1.f + expf(-x));
    const int g = gt[i] != 0.f;
#pragma unroll
    for (int t = 0; t < kMaxThresh; ++t) {
      const int v = prob >= th[t];
      cI[t] += v & g; cU[t] += v | g; cFP[t] += v & (g ^ 1); cFN[t] += (v ^ 1) & g;
    }
    if (d.bce_q20) {   // torch's stable form of BCEWithLogits, quantised per voxel to 2^-20
      const float l = fmaxf(x, 0.f) - x * gt[i] + log1pf(expf(-fabsf(x)));
      bce += __float2ll_rn(l * 1048576.f);
    }
  }
  if (d.bce_q20) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bce += __shfl_xor_sync(0xffffffffu, bce, o);
  }
#pragma unroll
  for (int t = 0; t < kMaxThresh; ++t) {
    if (t >= d.T) break;
    int a = cI[t], u = cU[t], fp = cFP[t], fn = cFN[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); u += __shfl_xor_sync(0xffffffffu, u, o);
      fp += __shfl_xor_sync(0xffffffffu, fp, o); fn += __shfl_xor_sync(0xffffffffu, fn, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sacc[t * 5 + 0], a); atomicAdd(&sacc[t * 5 + 1], u); atomicAdd(&sacc[t * 5 + 2], a);
      atomicAdd(&sacc[t * 5 + 3], fp); atomicAdd(&sacc[t * 5 + 4], fn);
    }
  }
  if (d.bce_q20 && (threadIdx.x & 31) == 0) atomicAdd(&sbce, (unsigned long long)bce);
  __syncthreads();
  for (int i = threadIdx.x; i < d.T * 5; i += blockDim.x) atomicAdd(d.counts + (long long)b * d.T * 5 + i, sacc[i]);
  if (d.bce_q20 && threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(d.bce_q20) + b, sbce);
}

// ---- [N,C,P] <-> [N,P,Cs] ------------------------------------------------------------------------------
template <typename TCL>   // storage type of the channels-last side; the planar side is fp32
__global__ void transpose_kernel(const svx_transpose_desc d) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (d.to_channels_last) {
    const float* src = d.in + (long long)n * d.C * d.P;
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, p = p0 + tx;
      tile[i][tx] = (c < d.C && p < d.P) ? src[(long long)c * d.P + p] : 0.f;
    }
    __syncthreads();
    TCL* dst = reinterpret_cast<TCL*>(d.out) + (long long)n * d.P * d.Cs;
    for (int i = ty; i < 32; i += 8) {
      const int p = p0 + i, c = c0 + tx;
      if (p < d.P && c < d.Cs) st1(dst + (long long)p * d.Cs + c, c < d.C ? maybe_round(tile[tx][i], d.round_tf32) : 0.f);
    }
  } else {
    const TCL* src = reinterpret_cast<const TCL*>(d.in) + (long long)n * d.P * d.Cs;
    for (int i = ty; i < 32; i += 8) {
      const int p = p0 + i, c = c0 + tx;
      tile[i][tx] = (p < d.P && c < d.C) ? ld1(src + (long long)p * d.Cs + c) : 0.f;
    }
    __syncthreads();
    float* dst = d.out + (long long)n * d.C * d.P;
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, p = p0 + tx;
      if (c < d.C && p < d.P) dst[(long long)c * d.P + p] = maybe_round(tile[tx][i], d.round_tf32);
    }
  }
}

// planar [N, C<=4, P] -> [N, P, 4] (image staging: NCHW fp32 -> NHWC with a zero fourth channel); one thread = one pixel
template <typename TO>
__global__ void interleave4_kernel(const svx_transpose_desc d, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / d.P;
    const int p = (int)(idx - n * d.P);
    const float* src = d.in + n * (long long)d.C * d.P + p;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < d.C) v[c] = maybe_round(__ldg(src + (long long)c * d.P), d.round_tf32);
    long long o = idx;
    if (d.row_w > 0) {   // zero columns around every image row
      const int y = p / d.row_w, x = p - y * d.row_w;
      o = n * (long long)(d.P / d.row_w) * d.row_pitch + (long long)y * d.row_pitch + d.row_x0 + x;
    }
    st4(reinterpret_cast<TO*>(d.out) + 4 * o, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// ---- F.interpolate(mode="bilinear", align_corners=False) on planar fp32 images (the Swin wrapper's input resize) -----------
__global__ void resize_planar_kernel(const svx_resize_desc d, long long total) {
  const float sy = (float)d.IH / (float)d.OH, sx = (float)d.IW / (float)d.OW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % d.OW);
    const long long t = idx / d.OW;
    const int oy = (int)(t % d.OH);
    const long long nc = t / d.OH;
    const float fy = fmaxf((oy + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min((int)fy, d.IH - 1), x0 = min((int)fx, d.IW - 1);
    const int y1 = min(y0 + 1, d.IH - 1), x1 = min(x0 + 1, d.IW - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* img = d.in + nc * (long long)d.IH * d.IW;
    const float a = __ldg(img + (long long)y0 * d.IW + x0), b = __ldg(img + (long long)y0 * d.IW + x1);
    const float c = __ldg(img + (long long)y1 * d.IW + x0), e = __ldg(img + (long long)y1 * d.IW + x1);
    d.out[idx] = (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * e);
  }
}

}  // namespace

int resize_launch(const svx_resize_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.NC > 0 && d.IH > 0 && d.IW > 0 && d.OH > 0 && d.OW > 0, "resize_bilinear: bad description");
  const long long total = (long long)d.NC * d.OH * d.OW;
  resize_planar_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  SVX_LAUNCH_OK("resize_planar_kernel");
  return 0;
}

int im2col_launch(const svx_im2col_desc& d, void* stream) {
  const int K = d.KD * d.KH * d.KW * d.C;
  SVX_REQUIRE(d.in && d.out && d.N > 0 && K > 0 && d.Kpad >= K, "im2col: bad description");
  SVX_REQUIRE(d.Kpad % 4 == 0 && al16(d.out), "im2col: Kpad must be a multiple of 4 and the output 16-byte aligned");
  const long long total4 = (long long)d.N * d.OD * d.OH * d.OW * (d.Kpad / 4);
  const long long lines = (long long)d.N * d.OD * d.OH;
  if (d.C == 1 && d.Kpad <= kI2cKMax && d.KD * d.KH * ((d.OW - 1) * d.stride + d.KW) <= kI2cRegionMax && lines < 0x7fffffffLL) {
    im2col_line_kernel<<<(int)lines, 256, 0, (cudaStream_t)stream>>>(d, K);
    SVX_LAUNCH_OK("im2col_line_kernel");
    return 0;
  }
  im2col_kernel<<<grid_for(total4, 256), 256, 0, (cudaStream_t)stream>>>(d, total4, K);
  SVX_LAUNCH_OK("im2col_kernel");
  return 0;
}

int pool_launch(const svx_pool_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.C % 4 == 0 && d.in_Cs % 4 == 0 && d.out_Cs % 4 == 0 && al16(d.in) && al16(d.out),
              "pool: channels must be multiples of 4 and pointers 16-byte aligned");
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "pool: input and output share one storage type");
  const long long total = (long long)d.N * d.OD * d.OH * d.OW * (d.C / 4);
  if (d.mode == SVX_POOL_MAX && d.D == 1 && d.OD == 1 && d.KD == 1 && d.KH == 3 && d.KW == 3 && d.SH == 2 && d.SW == 2 &&
      d.PD == 0 && d.PH == 1 && d.PW == 1 && d.C / 4 <= 256 && 256 % (d.C / 4) == 0 && d.N <= 65535 &&
      d.OH == (d.H - 1) / 2 + 1 && d.OW == (d.W - 1) / 2 + 1) {
    const int cols = 256 / (d.C / 4);
    dim3 grid((d.OW + cols - 1) / cols, d.N);
    if (d.dtype == SVX_DT_BF16) maxpool3s2_kernel<bf16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d, cols);
    else maxpool3s2_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(d, cols);
    SVX_LAUNCH_OK("maxpool3s2_kernel");
    return 0;
  }
  if (d.dtype == SVX_DT_BF16) pool_kernel<bf16_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  else pool_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  SVX_LAUNCH_OK("pool_kernel");
  return 0;
}

int lnrows_launch(const svx_lnrows_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.gamma && d.beta && d.rows > 0, "layernorm_rows: null operand");
  SVX_REQUIRE(d.C % 4 == 0 && (!d.merge || (d.C % 16 == 0 && d.H % 2 == 0 && d.W % 2 == 0)),
              "layernorm_rows: C=%d unsupported", d.C);
  SVX_REQUIRE(al16(d.in) && al16(d.out) && al16(d.gamma) && al16(d.beta), "layernorm_rows: unaligned pointer");
  SVX_REQUIRE(d.dtype != SVX_DT_BF16 || (d.C % 8 == 0 && (!d.merge || d.C % 32 == 0)), "layernorm_rows: bf16 rows need C %% 8 == 0");
  const int wpb = 8;
  const int nv = (d.C + 127) / 128;
  cudaStream_t st = (cudaStream_t)stream;
  auto grid = [&](int r) { return grid_for((d.rows + r - 1) / r, wpb, kSmCount * 8); };
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "layernorm_rows: input and output share one storage type");
  if (d.dtype == SVX_DT_BF16) {
    const int nv8 = (d.C + 255) / 256;   // 16-byte vectors (eight bf16) per lane
    if (nv8 == 1) lnrows_bf16_kernel<1, 4><<<grid(4), wpb * 32, 0, st>>>(d);
    else if (nv8 == 2) lnrows_bf16_kernel<2, 4><<<grid(4), wpb * 32, 0, st>>>(d);
    else if (nv8 == 3) lnrows_bf16_kernel<3, 2><<<grid(2), wpb * 32, 0, st>>>(d);
    else if (nv8 <= 6) lnrows_bf16_kernel<6, 1><<<grid(1), wpb * 32, 0, st>>>(d);
    else lnrows_kernel<bf16_t><<<grid_for(d.rows, wpb), wpb * 32, 0, st>>>(d);
  } else if (nv == 1) lnrows_reg_kernel<1, 4, float><<<grid(4), wpb * 32, 0, st>>>(d);
  else if (nv == 2) lnrows_reg_kernel<2, 4, float><<<grid(4), wpb * 32, 0, st>>>(d);
  else if (nv == 3) lnrows_reg_kernel<3, 2, float><<<grid(2), wpb * 32, 0, st>>>(d);
  else if (nv <= 6) lnrows_reg_kernel<6, 1, float><<<grid(1), wpb * 32, 0, st>>>(d);
  else if (nv <= 12) lnrows_reg_kernel<12, 1, float><<<grid(1), wpb * 32, 0, st>>>(d);
  else lnrows_kernel<float><<<grid_for(d.rows, wpb), wpb * 32, 0, st>>>(d);
  SVX_LAUNCH_OK("lnrows_kernel");
  return 0;
}

int lnsample_launch(const svx_lnsample_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.gamma && d.beta && d.N > 0 && d.L > 0 && d.L % 4 == 0, "layernorm_sample: bad description");
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "layernorm_sample: input and output share one storage type");
  const bool bf = d.dtype == SVX_DT_BF16;
  const int nv = ((d.L >> 2) + kLnCluster * 1024 - 1) / (kLnCluster * 1024);
  // measured (192 samples): 56x56x96: 0.162 ms vs 0.210 ms for the single-CTA kernel; 28x28x192: 0.114 vs 0.097;
  // 14x14x384: 0.093 vs 0.045 -> the cluster kernel only pays for the large samples
  if (nv <= 10 && d.L >= 4096 && nv >= 6) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d.N * kLnCluster);
    cfg.blockDim = dim3(1024);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kLnCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (bf) {
      if (nv <= 2) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<2, bf16_t>, d);
      else if (nv <= 3) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<3, bf16_t>, d);
      else if (nv <= 5) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<5, bf16_t>, d);
      else e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<10, bf16_t>, d);
    } else if (nv <= 2) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<2, float>, d);
    else if (nv <= 3) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<3, float>, d);
    else if (nv <= 5) e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<5, float>, d);
    else e = cudaLaunchKernelEx(&cfg, lnsample_cluster_kernel<10, float>, d);
    if (e != cudaSuccess) return fail("cluster launch of lnsample_cluster_kernel failed: %s", cudaGetErrorString(e));
    return 0;
  }
  if (bf) lnsample_kernel<bf16_t><<<d.N, 1024, 0, (cudaStream_t)stream>>>(d);
  else lnsample_kernel<float><<<d.N, 1024, 0, (cudaStream_t)stream>>>(d);
  SVX_LAUNCH_OK("lnsample_kernel");
  return 0;
}

int winattn_launch(const svx_winattn_desc& d, void* stream) {
  SVX_REQUIRE(d.qkv && d.out && d.bias, "window_attention: null operand");
  SVX_REQUIRE(d.H % WS == 0 && d.W % WS == 0 && d.C == d.heads * HD && d.shift >= 0 && d.shift < WS,
              "window_attention: needs 7x7 windows, head_dim 32 (H=%d W=%d C=%d heads=%d)", d.H, d.W, d.C, d.heads);
  const long long windows = (long long)d.N * (d.H / WS) * (d.W / WS);
  SVX_REQUIRE(windows * WT * (3LL * d.C / 4) < 0xffffffffLL && d.C % 4 == 0 && al16(d.qkv) && al16(d.out),
              "window_attention: tensor too large for 32-bit row offsets, or unaligned");
  SVX_REQUIRE(d.dtype == 0 || (d.dtype == SVX_DT_BF16 && d.C % 8 == 0), "window_attention: qkv and out share one storage type");
  return winattn_umma_launch(d, stream);
}

int dwconv_launch(const svx_dwconv_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.w && d.C % 4 == 0 && d.k >= 1, "dwconv: bad description");
  const long long total = (long long)d.N * d.OH * d.OW * (d.C / 4);
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "dwconv: input and output share one storage type");
  if (d.dtype == SVX_DT_BF16) dwconv_kernel<bf16_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  else dwconv_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  SVX_LAUNCH_OK("dwconv_kernel");
  return 0;
}

int viewattn_launch(const svx_viewattn_desc& d, void* stream) {
  SVX_REQUIRE(d.qkv && d.out && d.V >= 1 && d.V <= kMaxViews && d.R % d.heads == 0,
              "view_attention: supports 1..%d views (V=%d)", kMaxViews, d.V);
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "view_attention: input and output share one storage type");
  if (d.dtype == SVX_DT_BF16) viewattn_kernel<bf16_t><<<d.B * d.heads, 256, 0, (cudaStream_t)stream>>>(d);
  else viewattn_kernel<float><<<d.B * d.heads, 256, 0, (cudaStream_t)stream>>>(d);
  SVX_LAUNCH_OK("viewattn_kernel");
  return 0;
}

int bilinear_launch(const svx_bilinear_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.C % 4 == 0, "bilinear_add: bad description");
  const long long total = (long long)d.N * d.OH * d.OW * (d.C / 4);
  SVX_REQUIRE(d.dtype == 0 || d.dtype == SVX_DT_BF16, "bilinear_add: tensors share one storage type");
  if (d.dtype == SVX_DT_BF16) bilinear_kernel<bf16_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  else bilinear_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
  SVX_LAUNCH_OK("bilinear_kernel");
  return 0;
}

int mergefuse_launch(const svx_mergefuse_desc& d, void* stream) {
  SVX_REQUIRE(d.coarse && d.out && d.B > 0 && d.V > 0 && d.P % 4 == 0, "merger_fuse: bad description");
  SVX_REQUIRE((!d.weights || al16(d.weights)) && al16(d.coarse) && al16(d.out), "merger_fuse: unaligned pointer");
  const long long total4 = (long long)d.B * (d.P / 4);
  mergefuse_kernel<<<grid_for(total4, 256), 256, 0, (cudaStream_t)stream>>>(d, total4);
  SVX_LAUNCH_OK("mergefuse_kernel");
  return 0;
}

int conv3to1_launch(const svx_conv3to1_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.w && d.out && d.N > 0 && d.D > 0, "conv3to1: null operand");
  SVX_REQUIRE(d.Cin >= 1 && d.Cin <= 12 && d.W == C31_TW && d.H % C31_TH == 0 && d.Cs % 4 == 0 && d.c0 % 4 == 0 &&
                  d.c0 + 12 <= d.Cs && al16(d.in) && al16(d.w) && al16(d.out),
              "conv3to1: needs W == 32, H %% 16 == 0, 16-byte aligned 12-channel reads (Cin=%d W=%d H=%d Cs=%d c0=%d)",
              d.Cin, d.W, d.H, d.Cs, d.c0);
  const int smem = (4 * C31_PLANE * 3 + 27 * 3) * (int)sizeof(float4);
  SVX_CUDA_OK(cudaFuncSetAttribute(conv3to1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   // per device
  conv3to1_kernel<<<d.N * (d.H / C31_TH), C31_THREADS, smem, (cudaStream_t)stream>>>(d);
  SVX_LAUNCH_OK("conv3to1_kernel");
  return 0;
}

int metrics_launch(const svx_metrics_desc& d, void* stream) {
  SVX_REQUIRE(d.logits && d.gt && d.prob_thresholds && d.counts && d.T >= 1 && d.T <= kMaxThresh && d.B > 0 && d.P > 0,
              "voxel_metrics: supports 1..%d thresholds", kMaxThresh);
  SVX_CUDA_OK(cudaMemsetAsync(d.counts, 0, sizeof(int32_t) * (size_t)d.B * d.T * 5, (cudaStream_t)stream));
  if (d.bce_q20) SVX_CUDA_OK(cudaMemsetAsync(d.bce_q20, 0, sizeof(int64_t) * (size_t)d.B, (cudaStream_t)stream));
  int chunks = (2 * kSmCount + d.B - 1) / d.B;
  if (chunks < 1) chunks = 1;
  const int max_chunks = (d.P + 2047) / 2048;
  if (chunks > max_chunks) chunks = max_chunks;
  metrics_kernel<<<d.B * chunks, 256, 0, (cudaStream_t)stream>>>(d, chunks);
  SVX_LAUNCH_OK("metrics_kernel");
  return 0;
}

int transpose_launch(const svx_transpose_desc& d, void* stream) {
  SVX_REQUIRE(d.in && d.out && d.N > 0 && d.C > 0 && d.P > 0 && d.Cs >= d.C, "transpose: bad description");
  SVX_REQUIRE(d.row_w == 0 || (d.to_channels_last && d.Cs == 4 && d.C <= 4 && d.P % d.row_w == 0 && d.row_x0 >= 0 &&
                               d.row_x0 + d.row_w <= d.row_pitch && al16(d.out)),
              "transpose: padded rows need the 4-channel channels-last form");
  if (d.to_channels_last && d.Cs == 4 && d.C <= 4 && al16(d.out)) {
    const long long total = (long long)d.N * d.P;
    if (d.dtype == SVX_DT_OUT_BF16) interleave4_kernel<bf16_t><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
    else interleave4_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d, total);
    SVX_LAUNCH_OK("interleave4_kernel");
    return 0;
  }
  SVX_REQUIRE(d.N <= 65535, "transpose: N too large");
  const int cext = d.to_channels_last ? d.Cs : d.C;
  dim3 grid((d.P + 31) / 32, (cext + 31) / 32, d.N);
  const int want = d.to_channels_last ? SVX_DT_OUT_BF16 : SVX_DT_IN_BF16;
  SVX_REQUIRE(d.dtype == 0 || d.dtype == want, "transpose: only the channels-last side may be bf16");
  if (d.dtype) transpose_kernel<bf16_t><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(d);
  else transpose_kernel<float><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(d);
  SVX_LAUNCH_OK("transpose_kernel");
  return 0;
}

}  // namespace svx

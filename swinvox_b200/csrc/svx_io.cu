// svx_io.cu -- binvox run-length streams <-> dense device volumes (the ground-truth side of the IoU step:
// utils/binvox_rw.py:119-153 read_as_3d_array, :239-300 write; utils/data_loaders.py:84-87).
// Byte / index work, HBM- and latency-bound: one CTA per object, the volume staged in shared memory as bytes so that
// the global reads (payload) and writes (fp32 volume, pairs) are coalesced and the x,z,y -> x,y,z transpose is free.
#include <cuda_runtime.h>

#include <cstdint>

#include "svx_internal.h"

namespace svx {
namespace {

constexpr int kIoThreads = 256;

// exclusive prefix sum of one int per thread over the CTA; returns the exclusive value, *total = CTA sum
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();   // warp_tot may still be read by the previous call
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kIoThreads / 32; ++i) {
    const int t = warp_tot[i];
    if (i < w) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kIoThreads) binvox_decode_kernel(const svx_binvox_decode_desc d) {
  extern __shared__ uint8_t vol[];   // P bytes, file order (x, z, y)
  __shared__ int warp_tot[kIoThreads / 32];
  const int P = d.d0 * d.d1 * d.d2;
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    for (int i = threadIdx.x; i < P; i += kIoThreads) vol[i] = 0;
    __syncthreads();
    const uint8_t* pay = d.payload + d.offsets[b];
    const long long npairs = (d.offsets[b + 1] - d.offsets[b]) / 2;
    long long carry = 0;
    for (long long base = 0; base < npairs; base += kIoThreads) {
      const long long p = base + threadIdx.x;
      int val = 0, cnt = 0;
      if (p < npairs) {
        const uchar2 vc = *reinterpret_cast<const uchar2*>(pay + 2 * p);   // payload offsets are even (host contract)
        val = vc.x;
        cnt = vc.y;
      }
      int tot;
      const long long start = carry + block_excl_scan(cnt, warp_tot, &tot);
      if (val != 0) {   // np.repeat(values, counts).astype(bool): any non-zero value is "set"
        for (int i = 0; i < cnt; ++i)
          if (start + i < P) vol[start + i] = 1;
      }
      carry += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) d.status[b] = (int)(carry > 0x7fffffffLL ? 0x7fffffffLL : carry);
    float* out = d.out + (long long)b * P;
    if (d.fix_coords) {   // out[x][y][z] = file[x][z][y]
      for (int o = threadIdx.x; o < P; o += kIoThreads) {
        const int z = o % d.d1;
        const int t = o / d.d1;
        const int y = t % d.d2, x = t / d.d2;
        out[o] = vol[(x * d.d1 + z) * d.d2 + y] ? 1.f : 0.f;
      }
    } else {
      for (int o = threadIdx.x; o < P; o += kIoThreads) out[o] = vol[o] ? 1.f : 0.f;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kIoThreads) binvox_encode_kernel(const svx_binvox_encode_desc d) {
  extern __shared__ uint8_t smem[];
  __shared__ int warp_tot[kIoThreads / 32];
  const int P = d.d0 * d.d1 * d.d2;
  uint8_t* vol = smem;                                               // P bytes, file order
  int* starts = reinterpret_cast<int*>(smem + ((P + 15) / 16) * 16); // P + 1 run starts
  const int seg = (P + kIoThreads - 1) / kIoThreads;
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    const float* in = d.volume + (long long)b * P;
    if (d.axis_xyz) {   // input [x][y][z]; the file stores [x][z][y]: read coalesced, scatter into shared memory
      for (int i = threadIdx.x; i < P; i += kIoThreads) {
        const int z = i % d.d2;
        const int t = i / d.d2;
        const int y = t % d.d1, x = t / d.d1;
        vol[(x * d.d2 + z) * d.d1 + y] = in[i] >= d.threshold ? 1 : 0;
      }
    } else {
      for (int i = threadIdx.x; i < P; i += kIoThreads) vol[i] = in[i] >= d.threshold ? 1 : 0;
    }
    __syncthreads();
    // run starts: thread t owns voxels [t*seg, (t+1)*seg)
    const int f0 = threadIdx.x * seg, f1 = min(P, f0 + seg);
    int nstart = 0;
    for (int f = f0; f < f1; ++f) nstart += (f == 0 || vol[f] != vol[f - 1]) ? 1 : 0;
    int R;
    int r = block_excl_scan(nstart, warp_tot, &R);
    for (int f = f0; f < f1; ++f)
      if (f == 0 || vol[f] != vol[f - 1]) starts[r++] = f;
    if (threadIdx.x == 0) starts[R] = P;
    __syncthreads();
    // pairs per run (the writer's state machine, binvox_rw.py:279-300): L / 255 full pairs, then (value, L % 255) --
    // written even when it is 0 if another run follows, dropped when it is 0 at the end of the stream
    const int rseg = (R + kIoThreads - 1) / kIoThreads;
    const int r0 = threadIdx.x * rseg, r1 = min(R, r0 + rseg);
    int npair = 0;
    for (int q = r0; q < r1; ++q) {
      const int L = starts[q + 1] - starts[q];
      npair += L / 255 + ((q < R - 1 || L % 255 > 0) ? 1 : 0);
    }
    int total;
    int pos = block_excl_scan(npair, warp_tot, &total);
    uint8_t* out = d.payload + (long long)b * 2 * P;
    for (int q = r0; q < r1; ++q) {
      const int L = starts[q + 1] - starts[q];
      const uint8_t v = vol[starts[q]];
      for (int k = 0; k < L / 255; ++k, ++pos) *reinterpret_cast<uchar2*>(out + 2 * pos) = make_uchar2(v, 255);
      if (q < R - 1 || L % 255 > 0) {
        *reinterpret_cast<uchar2*>(out + 2 * pos) = make_uchar2(v, (uint8_t)(L % 255));
        ++pos;
      }
    }
    if (threadIdx.x == 0) d.nbytes[b] = 2 * total;
    __syncthreads();
  }
}

}  // namespace
}  // namespace svx

using namespace svx;

extern "C" {

int svx_binvox_decode(const svx_binvox_decode_desc* d, void* stream) {
  if (!d) return fail("svx_binvox_decode: null descriptor");
  SVX_REQUIRE(d->payload && d->offsets && d->out && d->status, "binvox_decode: null operand");
  SVX_REQUIRE(d->B > 0 && d->d0 > 0 && d->d1 > 0 && d->d2 > 0, "binvox_decode: empty problem");
  const long long P = (long long)d->d0 * d->d1 * d->d2;
  SVX_REQUIRE(P <= 200 * 1024, "binvox_decode: volumes above 204800 voxels are not supported (got %lld)", P);
  SVX_CUDA_OK(cudaFuncSetAttribute(binvox_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // per device
  const int grid = d->B < 148 * 4 ? d->B : 148 * 4;
  binvox_decode_kernel<<<grid, kIoThreads, (size_t)P, (cudaStream_t)stream>>>(*d);
  SVX_LAUNCH_OK("binvox_decode_kernel");
  return 0;
}

int svx_binvox_encode(const svx_binvox_encode_desc* d, void* stream) {
  if (!d) return fail("svx_binvox_encode: null descriptor");
  SVX_REQUIRE(d->volume && d->payload && d->nbytes, "binvox_encode: null operand");
  SVX_REQUIRE(d->B > 0 && d->d0 > 0 && d->d1 > 0 && d->d2 > 0, "binvox_encode: empty problem");
  const long long P = (long long)d->d0 * d->d1 * d->d2;
  SVX_REQUIRE(P <= 40000, "binvox_encode: volumes above 40000 voxels are not supported (got %lld)", P);
  const size_t smem = (size_t)((P + 15) / 16 * 16 + 4 * (P + 1));
  SVX_CUDA_OK(cudaFuncSetAttribute(binvox_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));   // per device
  const int grid = d->B < 148 ? d->B : 148;
  binvox_encode_kernel<<<grid, kIoThreads, smem, (cudaStream_t)stream>>>(*d);
  SVX_LAUNCH_OK("binvox_encode_kernel");
  return 0;
}

}  // extern "C"

// ---- evaluation-time image pipeline (core/test.py:50-55, utils/data_transforms.py:76-167,415-452,42-62) -------------
// uint8 H x W x C renderings (C = 4: BGRA, or 3) -> centre crop -> bilinear resize (cv2 INTER_LINEAR convention) of the
// value/255 floats -> background colour where the RESIZED alpha is exactly 0 -> (x - mean) / std -> planar fp32
// [N, 3, OH, OW], the encoder's input.  One thread per output pixel (all channels): reads are the 2 x 2 source
// neighbourhood (L1/L2-resident, the source is 3-5x smaller than the output), writes are coalesced per plane.
namespace svx {
namespace {

__global__ void __launch_bounds__(256) preprocess_kernel(const svx_preprocess_desc d) {
  const long long total = (long long)d.N * d.OH * d.OW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % d.OW);
    const long long t = idx / d.OW;
    const int oy = (int)(t % d.OH);
    const long long n = t / d.OH;
    // source coordinate fx = (o + 0.5) * src / dst - 0.5 as the exact rational ((2o+1) src - dst) / (2 dst)
    auto coord = [](int o, int src, int dst, int& i0, int& i1, float& w1) {
      const int num = (2 * o + 1) * src - dst, den = 2 * dst;
      int s = num >= 0 ? num / den : -((-num + den - 1) / den);
      float f = (float)(num - s * den) / (float)den;
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= src - 1) { s = src - 1; f = 0.f; }
      i0 = s; i1 = s + 1 < src ? s + 1 : src - 1; w1 = f;
    };
    // crop window of this image; a per-image window may leave the image: outside coordinates clamp (edge padding)
    int y0 = d.y0, y1 = d.y1, x0 = d.x0, x1 = d.x1;
    if (d.windows) {
      const int4 wv = __ldg(reinterpret_cast<const int4*>(d.windows) + n);
      y0 = wv.x; y1 = wv.y; x0 = wv.z; x1 = wv.w;
    }
    const int ch = y1 - y0, cw = x1 - x0;
    int xa, xb, ya, yb;
    float wx, wy;
    coord(ox, cw, d.OW, xa, xb, wx);
    coord(oy, ch, d.OH, ya, yb, wy);
    auto cl = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
    const int sya = cl(y0 + ya, d.H - 1), syb = cl(y0 + yb, d.H - 1), sxa = cl(x0 + xa, d.W - 1), sxb = cl(x0 + xb, d.W - 1);
    const uint8_t* img = d.in + n * (long long)d.H * d.W * d.C;
    const uint8_t* p00 = img + ((long long)sya * d.W + sxa) * d.C;
    const uint8_t* p01 = img + ((long long)sya * d.W + sxb) * d.C;
    const uint8_t* p10 = img + ((long long)syb * d.W + sxa) * d.C;
    const uint8_t* p11 = img + ((long long)syb * d.W + sxb) * d.C;
    auto lerp2 = [&](int c) {
      const float v00 = __fdiv_rn((float)p00[c], 255.f), v01 = __fdiv_rn((float)p01[c], 255.f);
      const float v10 = __fdiv_rn((float)p10[c], 255.f), v11 = __fdiv_rn((float)p11[c], 255.f);
      const float r0 = fmaf(wx, v01 - v00, v00), r1 = fmaf(wx, v11 - v10, v10);
      return fmaf(wy, r1 - r0, r0);
    };
    bool background = false;
    if (d.C == 4) background = lerp2(3) == 0.f;   // alpha * bg + (1 - alpha) * img with alpha = (resized alpha == 0)
    float* out = d.out + n * 3LL * d.OH * d.OW + (long long)oy * d.OW + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float bgv = d.bg_norm_n ? __ldg(d.bg_norm_n + n * 3 + c) : d.bg_norm[c];
      const float v = background ? bgv : (lerp2(c) - d.mean[c]) / d.std[c];
      out[(long long)c * d.OH * d.OW] = v;
    }
  }
}

}  // namespace
}  // namespace svx

extern "C" int svx_preprocess(const svx_preprocess_desc* d, void* stream) {
  using namespace svx;
  if (!d) return fail("svx_preprocess: null descriptor");
  SVX_REQUIRE(d->in && d->out && d->N > 0 && d->H > 0 && d->W > 0 && (d->C == 3 || d->C == 4) && d->OH > 0 && d->OW > 0,
              "preprocess: bad description");
  SVX_REQUIRE(d->windows || (d->y0 >= 0 && d->x0 >= 0 && d->y1 > d->y0 && d->x1 > d->x0 && d->y1 <= d->H && d->x1 <= d->W),
              "preprocess: crop window outside the image");
  SVX_REQUIRE(d->windows || ((long long)(2 * d->OW + 1) * (d->x1 - d->x0) < 0x7fffffffLL &&
                             (long long)(2 * d->OH + 1) * (d->y1 - d->y0) < 0x7fffffffLL),
              "preprocess: image too large");
  SVX_REQUIRE(!d->windows || (reinterpret_cast<uintptr_t>(d->windows) & 15) == 0, "preprocess: windows must be 16-byte aligned");
  const long long total = (long long)d->N * d->OH * d->OW;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  preprocess_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(*d);
  SVX_LAUNCH_OK("preprocess_kernel");
  return 0;
}

// umma_probe2.cu -- what does a tcgen05.mma kind::tf32 stream pay for (a) tcgen05.commit between groups of MMAs and
// (b) moving the accumulator window?  M = 128, K = 8, operands in shared memory, one issuing thread per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I swinvox_b200/csrc tools/probes/umma_probe2.cu -o tools/probes/umma_probe2.bin
#include <cstdio>

#include "svx_ptx.cuh"

using namespace svx;

// pattern 0: one accumulator; 1: group g uses 48-column slot (g % 10); 2: N=144 window sliding down by 48 columns per
// group (the kd-in-N issuer); 3: like 1 but the first MMA of every group has accumulate = 0
__global__ void __launch_bounds__(128, 1) probe(int n, int group, int commit, int pattern, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[64];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
  if (warp == 0) {
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&bars[i]), 1u); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<512>(smem_u32(&slot));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_tf32(128, n);
    const uint64_t db = umma_desc_sw128(base + 96 * 1024);
    const long long t0 = clock64();
    int g = 0;
    for (int it = 0; it < iters; it += group, ++g) {
      uint32_t acc = tmem;
      if (pattern == 1 || pattern == 3) acc = tmem + (g % 10) * 48;
      if (pattern == 2) acc = tmem + (7 - g % 8) * 48;
      if (pattern >= 4) {
        // the kd-in-N issuer's sequence: N=48 (accumulate 0) at c, N=96 at c+48, then N=144 at c (overlapping windows
        // of different shapes); pattern 5: the same shapes but all three at disjoint columns
        acc = tmem + (7 - g % 8) * 48;
        const uint64_t da0 = umma_desc_sw128(base + (it % 3) * 200 * 128);
        const uint32_t far = pattern == 5 ? 256u : 0u;
        umma_tf32(acc, da0, db, umma_idesc_tf32(128, 48), pattern == 4 ? 0u : 1u);
        umma_tf32(acc + 48 + (pattern == 5 ? 16u : 0u), da0, db, umma_idesc_tf32(128, 96), 1u);
        for (int j = 2; j < group; ++j) {
          const uint64_t da = umma_desc_sw128(base + ((it + j) % 3) * 200 * 128) + 2u * (j & 3);
          umma_tf32(acc + far, da, db + 2u * (j & 3), umma_idesc_tf32(128, pattern == 5 ? 96 : 144), 1u);
        }
      } else
      for (int j = 0; j < group; ++j) {
        const uint64_t da = umma_desc_sw128(base + ((it + j) % 3) * 200 * 128) + 2u * (j & 3);
        umma_tf32(acc, da, db + 2u * (j & 3), idesc, (pattern == 3 && j == 0) ? 0u : 1u);
      }
      if (commit == 1) umma_commit(smem_u32(&bars[g % 63]));
      if (commit == 2) { umma_commit(smem_u32(&bars[g % 63])); umma_commit(smem_u32(&bars[(g + 31) % 63])); }
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bars[63]));
    mbar_wait(smem_u32(&bars[63]), 0u);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("# N group commits/group pattern : cycles per MMA (until everything completed), cycles per group\n");
  for (int pattern : {4, 5, 2})
    for (int n : {48, 144, 192})
      for (int group : {4, 7, 18})
        for (int commit : {0, 1, 2}) {
          if (pattern >= 2 && n != 144) continue;
          if ((pattern == 1 || pattern == 3) && n != 48) continue;
          const int iters = group * 600;
          probe<<<148, 128, 200 * 1024>>>(n, group, commit, pattern, iters, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%4d %3d %d %d : %7.1f %8.1f\n", n, group, commit, pattern, (double)out[1] / iters, (double)out[1] / iters * group);
        }
  return 0;
}

// umma_probe.cu -- measures the issue cost of tcgen05.mma kind::tf32 (M = 128, K = 8, operands in shared memory) on
// sm_100a as a function of N, the operand swizzle (128B / 64B rows) and the A start offset inside the swizzle atom.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I swinvox_b200/csrc tools/probes/umma_probe.cu -o gpurun_out/umma_probe
// One CTA per SM; one thread issues `iters` MMAs back to back (cycling over `nslots` distinct A tiles and k-steps so
// the operand reads are real), commits, waits; cycles = clock64 delta / iters.
#include <cstdio>
#include <cstdlib>

#include "svx_ptx.cuh"

using namespace svx;

__global__ void __launch_bounds__(128, 1) probe(int n, int rowb, int a_off, int iters, int ksteps, int acc_mode,
                                                long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1u); fence_barrier_init(); }
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
    const uint32_t a_base = base, b_base = base + 96 * 1024;   // A region: 96 KB, B: 256 rows x 128 B
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_addr = a_base + (it % 3) * 200 * rowb + a_off;   // three "slabs", shifted like the (kd,kh) taps
      const int k = it % ksteps;
      const uint64_t da = (rowb == 128 ? umma_desc_sw128(a_addr) : umma_desc_sw64(a_addr)) + 2u * k;
      const uint64_t db = (rowb == 128 ? umma_desc_sw128(b_base) : umma_desc_sw64(b_base)) + 2u * k;
      const uint32_t acc = acc_mode == 0 ? tmem : tmem + (it % 2) * 256;
      umma_tf32(acc, da, db, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0u);
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
  const int iters = 4096;
  printf("# tcgen05.mma kind::tf32 M=128 K=8 SS: cycles per instruction (issue loop only / until commit completes)\n");
  printf("# N rowB a_off ksteps acc  issue  total  cyc_per_Ncol\n");
  const int ns[] = {16, 48, 96, 144, 192, 256};
  for (int rowb : {128, 64})
    for (int a_off : {0, 34 * 64, 34 * 128})
      for (int ks : {1, 2, 4})
        for (int n : ns) {
          if (rowb == 64 && ks == 4) continue;
          for (int acc : {0, 1}) {
            if (acc == 1 && n > 256) continue;
            probe<<<148, 128, 200 * 1024>>>(n, rowb, a_off, iters, ks, acc, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%4d %4d %5d %2d %d  %7.1f %7.1f  %5.2f\n", n, rowb, a_off, ks, acc, (double)out[0] / iters,
                   (double)out[1] / iters, (double)out[1] / iters / n);
          }
        }
  return 0;
}

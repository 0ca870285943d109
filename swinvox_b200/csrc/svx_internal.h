// svx_internal.h -- shared host-side helpers (error reporting, launch checks).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/swinvox_b200.h"
#ifndef SVX_HOSTSIM
#include <cuda_runtime.h>
#endif

namespace svx {

// thread-local error slot behind svx_last_error()
char* error_buffer();
int fail(const char* fmt, ...);

#ifndef SVX_HOSTSIM
#define SVX_CUDA_OK(expr)                                                             \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return svx::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define SVX_LAUNCH_OK(name)                                                           \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      return svx::fail("launch of %s failed: %s", name, cudaGetErrorString(e__));     \
  } while (0)
#endif

#define SVX_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) return svx::fail(__VA_ARGS__); \
  } while (0)

// ---- op launchers shared by the immediate API and the plan executor -----------------------
// `prepared` caches per-op host state (TMA descriptors) between launches; may be null.
struct GemmPrepared;
int gemm_prepare(const svx_gemm_desc& d, GemmPrepared** out);
void gemm_prepared_free(GemmPrepared* p);
int gemm_launch(const svx_gemm_desc& d, GemmPrepared* prepared, void* stream);
struct MlpPrepared;
int mlp_prepare(const svx_mlp_desc& d, MlpPrepared** out);
void mlp_prepared_free(MlpPrepared* p);
int mlp_launch(const svx_mlp_desc& d, MlpPrepared* prepared, void* stream);

int im2col_launch(const svx_im2col_desc& d, void* stream);
int pool_launch(const svx_pool_desc& d, void* stream);
int lnrows_launch(const svx_lnrows_desc& d, void* stream);
int lnsample_launch(const svx_lnsample_desc& d, void* stream);
int winattn_launch(const svx_winattn_desc& d, void* stream);
int winattn_umma_launch(const svx_winattn_desc& d, void* stream);   // svx_winattn.cu: the tcgen05 kernel behind winattn_launch
int dwconv_launch(const svx_dwconv_desc& d, void* stream);
int viewattn_launch(const svx_viewattn_desc& d, void* stream);
int bilinear_launch(const svx_bilinear_desc& d, void* stream);
int mergefuse_launch(const svx_mergefuse_desc& d, void* stream);
int conv3to1_launch(const svx_conv3to1_desc& d, void* stream);
int metrics_launch(const svx_metrics_desc& d, void* stream);
int transpose_launch(const svx_transpose_desc& d, void* stream);
int resize_launch(const svx_resize_desc& d, void* stream);

// kernel launches a single op issues (for svx_plan_num_launches)
int gemm_num_launches(const svx_gemm_desc& d);

}  // namespace svx

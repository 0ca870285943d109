// svx_api.cu -- the extern "C" surface declared in include/swinvox_b200.h: immediate launches and
// the plan executor (a recorded op list replayed per forward, optionally as a CUDA graph).
#ifndef SVX_HOSTSIM
#include <cuda_runtime.h>
#endif

#include <vector>

#include "svx_internal.h"

namespace svx {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return 1;
}

enum OpKind {
  OP_GEMM, OP_IM2COL, OP_POOL, OP_LNROWS, OP_LNSAMPLE, OP_WINATTN, OP_DWCONV, OP_VIEWATTN, OP_BILINEAR,
  OP_MERGEFUSE, OP_METRICS, OP_TRANSPOSE, OP_CONV3TO1, OP_MLP, OP_RESIZE, OP_JOIN
};

struct Op {
  OpKind kind;
  union {
    svx_gemm_desc gemm;
    svx_im2col_desc im2col;
    svx_pool_desc pool;
    svx_lnrows_desc lnrows;
    svx_lnsample_desc lnsample;
    svx_winattn_desc winattn;
    svx_dwconv_desc dwconv;
    svx_viewattn_desc viewattn;
    svx_bilinear_desc bilinear;
    svx_mergefuse_desc mergefuse;
    svx_metrics_desc metrics;
    svx_conv3to1_desc conv3to1;
    svx_transpose_desc transpose;
    svx_mlp_desc mlp;
    svx_resize_desc resize;
  } u;
  GemmPrepared* prepared = nullptr;
  MlpPrepared* mlp_prepared = nullptr;
  int lane = 0;   // 0: the caller's stream; k > 0: side stream k (ops of different lanes may overlap until the next join)
};

int launch_op(Op& op, void* stream) {
  switch (op.kind) {
    case OP_GEMM: return gemm_launch(op.u.gemm, op.prepared, stream);
    case OP_IM2COL: return im2col_launch(op.u.im2col, stream);
    case OP_POOL: return pool_launch(op.u.pool, stream);
    case OP_LNROWS: return lnrows_launch(op.u.lnrows, stream);
    case OP_LNSAMPLE: return lnsample_launch(op.u.lnsample, stream);
    case OP_WINATTN: return winattn_launch(op.u.winattn, stream);
    case OP_DWCONV: return dwconv_launch(op.u.dwconv, stream);
    case OP_VIEWATTN: return viewattn_launch(op.u.viewattn, stream);
    case OP_BILINEAR: return bilinear_launch(op.u.bilinear, stream);
    case OP_MERGEFUSE: return mergefuse_launch(op.u.mergefuse, stream);
    case OP_METRICS: return metrics_launch(op.u.metrics, stream);
    case OP_TRANSPOSE: return transpose_launch(op.u.transpose, stream);
    case OP_CONV3TO1: return conv3to1_launch(op.u.conv3to1, stream);
    case OP_MLP: return mlp_launch(op.u.mlp, op.mlp_prepared, stream);
    case OP_RESIZE: return resize_launch(op.u.resize, stream);
    case OP_JOIN: return 0;
  }
  return fail("unknown op kind");
}

}  // namespace svx

constexpr int kMaxLanes = 8;

struct svx_plan {
  std::vector<svx::Op> ops;
  int cur_lane = 0;
#ifndef SVX_HOSTSIM
  cudaGraphExec_t graph_exec = nullptr;
  size_t graph_ops = 0;
  cudaStream_t side[kMaxLanes] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[kMaxLanes] = {};
#endif
};

using namespace svx;

extern "C" {

int svx_abi_version(void) { return SVX_ABI_VERSION; }
const char* svx_last_error(void) { return error_buffer(); }

int svx_desc_sizes(int32_t* sizes, int n) {
  const int32_t s[] = {(int32_t)sizeof(svx_gemm_desc),     (int32_t)sizeof(svx_im2col_desc),
                       (int32_t)sizeof(svx_pool_desc),     (int32_t)sizeof(svx_lnrows_desc),
                       (int32_t)sizeof(svx_lnsample_desc), (int32_t)sizeof(svx_winattn_desc),
                       (int32_t)sizeof(svx_dwconv_desc),   (int32_t)sizeof(svx_viewattn_desc),
                       (int32_t)sizeof(svx_bilinear_desc), (int32_t)sizeof(svx_mergefuse_desc),
                       (int32_t)sizeof(svx_metrics_desc),  (int32_t)sizeof(svx_transpose_desc),
                       (int32_t)sizeof(svx_conv3to1_desc), (int32_t)sizeof(svx_mlp_desc),
                       (int32_t)sizeof(svx_resize_desc)};
  const int have = (int)(sizeof(s) / sizeof(s[0]));
  for (int i = 0; i < n && i < have; ++i) sizes[i] = s[i];
  return have;
}

#ifdef SVX_HOSTSIM
// CPU twin used only by tests/hostsim (never loaded by the swinvox_b200 package)
int svx_device_info(int, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  if (sm_count) *sm_count = 0;
  if (cc_major) *cc_major = 0;
  if (cc_minor) *cc_minor = 0;
  return 0;
}
#else
int svx_device_info(int device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  cudaDeviceProp prop;
  SVX_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
#endif

#define SVX_IMMEDIATE(fn, desc_t, launcher)                       \
  int fn(const desc_t* d, void* stream) {                         \
    if (!d) return fail(#fn ": null descriptor");                 \
    return launcher(*d, stream);                                  \
  }

int svx_gemm(const svx_gemm_desc* d, void* stream) {
  if (!d) return fail("svx_gemm: null descriptor");
  return gemm_launch(*d, nullptr, stream);
}
int svx_mlp(const svx_mlp_desc* d, void* stream) {
  if (!d) return fail("svx_mlp: null descriptor");
  return mlp_launch(*d, nullptr, stream);
}
SVX_IMMEDIATE(svx_im2col, svx_im2col_desc, im2col_launch)
SVX_IMMEDIATE(svx_pool, svx_pool_desc, pool_launch)
SVX_IMMEDIATE(svx_layernorm_rows, svx_lnrows_desc, lnrows_launch)
SVX_IMMEDIATE(svx_layernorm_sample, svx_lnsample_desc, lnsample_launch)
SVX_IMMEDIATE(svx_window_attention, svx_winattn_desc, winattn_launch)
SVX_IMMEDIATE(svx_dwconv, svx_dwconv_desc, dwconv_launch)
SVX_IMMEDIATE(svx_view_attention, svx_viewattn_desc, viewattn_launch)
SVX_IMMEDIATE(svx_bilinear_add, svx_bilinear_desc, bilinear_launch)
SVX_IMMEDIATE(svx_merger_fuse, svx_mergefuse_desc, mergefuse_launch)
SVX_IMMEDIATE(svx_voxel_metrics, svx_metrics_desc, metrics_launch)
SVX_IMMEDIATE(svx_transpose, svx_transpose_desc, transpose_launch)
SVX_IMMEDIATE(svx_conv3to1, svx_conv3to1_desc, conv3to1_launch)
SVX_IMMEDIATE(svx_resize_bilinear, svx_resize_desc, resize_launch)

svx_plan* svx_plan_create(void) { return new svx_plan(); }

void svx_plan_destroy(svx_plan* p) {
  if (!p) return;
  for (auto& op : p->ops)
    if (op.prepared) gemm_prepared_free(op.prepared);
  for (auto& op : p->ops)
    if (op.mlp_prepared) mlp_prepared_free(op.mlp_prepared);
#ifndef SVX_HOSTSIM
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (int k = 0; k < kMaxLanes; ++k) {
    if (p->side[k]) cudaStreamDestroy(p->side[k]);
    if (p->join_ev[k]) cudaEventDestroy(p->join_ev[k]);
  }
  if (p->fork_ev) cudaEventDestroy(p->fork_ev);
#endif
  delete p;
}

int svx_plan_num_ops(const svx_plan* p) { return p ? (int)p->ops.size() : 0; }

int svx_plan_add_gemm(svx_plan* p, const svx_gemm_desc* d) {
  if (!p || !d) return fail("svx_plan_add_gemm: null argument");
  Op op;
  op.kind = OP_GEMM;
  op.u.gemm = *d;
  if (int rc = gemm_prepare(*d, &op.prepared)) return rc;
  op.lane = p->cur_lane;
  p->ops.push_back(op);
  return 0;
}

int svx_plan_add_mlp(svx_plan* p, const svx_mlp_desc* d) {
  if (!p || !d) return fail("svx_plan_add_mlp: null argument");
  Op op;
  op.kind = OP_MLP;
  op.u.mlp = *d;
  if (int rc = mlp_prepare(*d, &op.mlp_prepared)) return rc;
  op.lane = p->cur_lane;
  p->ops.push_back(op);
  return 0;
}

#define SVX_PLAN_ADD(fn, desc_t, kind_, member)              \
  int fn(svx_plan* p, const desc_t* d) {                     \
    if (!p || !d) return fail(#fn ": null argument");        \
    Op op;                                                   \
    op.kind = kind_;                                         \
    op.u.member = *d;                                        \
    op.lane = p->cur_lane;                                   \
    p->ops.push_back(op);                                    \
    return 0;                                                \
  }
SVX_PLAN_ADD(svx_plan_add_im2col, svx_im2col_desc, OP_IM2COL, im2col)
SVX_PLAN_ADD(svx_plan_add_pool, svx_pool_desc, OP_POOL, pool)
SVX_PLAN_ADD(svx_plan_add_layernorm_rows, svx_lnrows_desc, OP_LNROWS, lnrows)
SVX_PLAN_ADD(svx_plan_add_layernorm_sample, svx_lnsample_desc, OP_LNSAMPLE, lnsample)
SVX_PLAN_ADD(svx_plan_add_window_attention, svx_winattn_desc, OP_WINATTN, winattn)
SVX_PLAN_ADD(svx_plan_add_dwconv, svx_dwconv_desc, OP_DWCONV, dwconv)
SVX_PLAN_ADD(svx_plan_add_view_attention, svx_viewattn_desc, OP_VIEWATTN, viewattn)
SVX_PLAN_ADD(svx_plan_add_bilinear_add, svx_bilinear_desc, OP_BILINEAR, bilinear)
SVX_PLAN_ADD(svx_plan_add_merger_fuse, svx_mergefuse_desc, OP_MERGEFUSE, mergefuse)
SVX_PLAN_ADD(svx_plan_add_voxel_metrics, svx_metrics_desc, OP_METRICS, metrics)
SVX_PLAN_ADD(svx_plan_add_transpose, svx_transpose_desc, OP_TRANSPOSE, transpose)
SVX_PLAN_ADD(svx_plan_add_conv3to1, svx_conv3to1_desc, OP_CONV3TO1, conv3to1)
SVX_PLAN_ADD(svx_plan_add_resize_bilinear, svx_resize_desc, OP_RESIZE, resize)

int svx_plan_set_lane(svx_plan* p, int lane) {
  if (!p || lane < 0 || lane > kMaxLanes) return fail("svx_plan_set_lane: lane must be in [0, %d]", kMaxLanes);
  p->cur_lane = lane;
  return 0;
}

int svx_plan_add_join(svx_plan* p) {
  if (!p) return fail("svx_plan_add_join: null plan");
  Op op;
  op.kind = OP_JOIN;
  memset(&op.u, 0, sizeof(op.u));
  p->ops.push_back(op);
  return 0;
}

// Ops of lane 0 run on the caller's stream.  The first op of a side lane since the last join forks that lane's
// stream from the caller's stream at this point of the op list; a join op (and the end of the range) makes the
// caller's stream wait for every side lane.  Works identically under stream capture (the lanes become graph branches).
int svx_plan_run_range(svx_plan* p, int first, int last, void* stream) {
  if (!p) return fail("svx_plan_run_range: null plan");
  if (first < 0 || last > (int)p->ops.size() || first > last) return fail("svx_plan_run_range: bad range");
#ifdef SVX_HOSTSIM
  for (int i = first; i < last; ++i)
    if (int rc = launch_op(p->ops[i], stream)) return rc;
  return 0;
#else
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool forked[kMaxLanes] = {};
  auto join_all = [&]() -> int {
    for (int k = 0; k < kMaxLanes; ++k) {
      if (!forked[k]) continue;
      SVX_CUDA_OK(cudaEventRecord(p->join_ev[k], p->side[k]));
      SVX_CUDA_OK(cudaStreamWaitEvent(st, p->join_ev[k], 0));
      forked[k] = false;
    }
    return 0;
  };
  for (int i = first; i < last; ++i) {
    Op& op = p->ops[i];
    if (op.kind == OP_JOIN) {
      if (int rc = join_all()) return rc;
      continue;
    }
    if (op.lane == 0) {
      if (int rc = launch_op(op, stream)) return rc;
      continue;
    }
    const int k = op.lane - 1;
    if (!p->side[k]) {
      SVX_CUDA_OK(cudaStreamCreateWithFlags(&p->side[k], cudaStreamNonBlocking));
      SVX_CUDA_OK(cudaEventCreateWithFlags(&p->join_ev[k], cudaEventDisableTiming));
    }
    if (!p->fork_ev) SVX_CUDA_OK(cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming));
    if (!forked[k]) {
      SVX_CUDA_OK(cudaEventRecord(p->fork_ev, st));
      SVX_CUDA_OK(cudaStreamWaitEvent(p->side[k], p->fork_ev, 0));
      forked[k] = true;
    }
    if (int rc = launch_op(op, p->side[k])) return rc;
  }
  return join_all();
#endif
}

int svx_plan_run(svx_plan* p, void* stream, int use_graph) {
  if (!p) return fail("svx_plan_run: null plan");
  if (!use_graph) return svx_plan_run_range(p, 0, (int)p->ops.size(), stream);
#ifdef SVX_HOSTSIM
  return svx_plan_run_range(p, 0, (int)p->ops.size(), stream);
#else
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!p->graph_exec || p->graph_ops != p->ops.size()) {
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    // warm every kernel once outside capture (function attributes are set lazily at first launch)
    if (int rc = svx_plan_run_range(p, 0, (int)p->ops.size(), stream)) return rc;
    cudaStream_t cap;
    SVX_CUDA_OK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    if (e == cudaSuccess) {
      rc = svx_plan_run_range(p, 0, (int)p->ops.size(), cap);
      e = cudaStreamEndCapture(cap, &graph);
    }
    if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    if (rc) return rc;
    if (e != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(e));
    p->graph_ops = p->ops.size();
    return 0;  // the warm run above already produced this call's results
  }
  SVX_CUDA_OK(cudaGraphLaunch(p->graph_exec, st));
  return 0;
#endif
}

int svx_plan_time_ops(svx_plan* p, void* stream, int iters, float* ms) {
  if (!p || !ms || iters < 1) return fail("svx_plan_time_ops: bad argument");
#ifdef SVX_HOSTSIM
  (void)stream;
  return fail("svx_plan_time_ops: not available in the host simulator");
#else
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t a, b;
  SVX_CUDA_OK(cudaEventCreate(&a));
  SVX_CUDA_OK(cudaEventCreate(&b));
  int rc = 0;
  for (size_t i = 0; i < p->ops.size() && !rc; ++i) {
    if (p->ops[i].kind == OP_JOIN) { ms[i] = 0.f; continue; }
    rc = launch_op(p->ops[i], stream);  // warm
    cudaEventRecord(a, st);
    for (int it = 0; it < iters && !rc; ++it) rc = launch_op(p->ops[i], stream);
    cudaEventRecord(b, st);
    if (cudaEventSynchronize(b) != cudaSuccess) rc = fail("svx_plan_time_ops: op %d failed: %s", (int)i,
                                                        cudaGetErrorString(cudaGetLastError()));
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    ms[i] = t / iters;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return rc;
#endif
}

int svx_plan_num_launches(const svx_plan* p) {
  if (!p) return 0;
  int n = 0;
  for (auto& op : p->ops) n += (op.kind == OP_GEMM) ? gemm_num_launches(op.u.gemm) : (op.kind == OP_JOIN ? 0 : 1);
  return n;
}

}  // extern "C"

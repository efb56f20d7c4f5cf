// The eps sweep of recursive_dbscan resident on the device (sm_100a).
//
// Reference: autometa/binning/recursive_dbscan.py:172-215.  Per eps the reference runs DBSCAN, add_metrics, the
// cutoff filter and a median on the host and then decides, in Python, whether this eps is the best so far, how
// far to step and whether to stop.  With the step itself on the GPU (am_dbscan.cu, am_metrics.cu) that decision
// is the only thing left that needs the host — one synchronisation and one read-back per eps, which is what
// bounds the small (<= 10 000 contig) per-taxon sweeps of large-data mode.  Here a one-thread controller kernel
// takes the decision (the loop state lives in d_ctl), writes the next step's cell grid to where
// am_eps_union_dyn reads it and a copy kernel keeps the best labels / metrics.  One step is a fixed launch
// sequence whatever the eps, so it is captured once into a CUDA graph and replayed; the host queues batches of
// replays and reads the `done` word once per batch.
#include "am_common.cuh"

namespace {

constexpr int W = AM_SWEEP_CTL_WORDS;

inline int64_t al256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct CtlLayout {
  int64_t rounds_off, trace_off, total;
};

inline CtlLayout ctl_layout(int64_t n, int trace_cap) {
  CtlLayout l;
  l.rounds_off = al256((int64_t)W * 8);
  l.trace_off = l.rounds_off + al256((n + 1) * 4);
  l.total = l.trace_off + al256((int64_t)std::max(trace_cap, 0) * 4 * 8);
  return l;
}

// One turn of the reference's loop body after the metrics (recursive_dbscan.py:194-215).  Shared by the device
// controller and am_sweep_advance_host so that the tests can check the schedule without a GPU.
__host__ __device__ inline int sweep_advance(double *ctl, int *rounds, int F, int nc, double med, double *trace,
                                            int trace_cap) {
  if (ctl[AM_SWEEP_DONE] != 0.0) {  // frozen: a step queued after the end changes nothing
    ctl[AM_SWEEP_TAKE] = 0.0;
    return 0;
  }
  const double eps = ctl[AM_SWEEP_EPS];
  const int k = (int)ctl[AM_SWEEP_NSTEPS];
  int take = 0;
  if (med >= ctl[AM_SWEEP_BEST]) {  // :194-196 (ties go to the later eps)
    ctl[AM_SWEEP_BEST] = med;
    ctl[AM_SWEEP_BEST_STEP] = (double)k;
    ctl[AM_SWEEP_BEST_NC] = (double)nc;
    ctl[AM_SWEEP_BEST_EPS] = eps;
    take = 1;
  }
  ctl[AM_SWEEP_TAKE] = (double)take;
  ctl[AM_SWEEP_LAST_NC] = (double)nc;
  ctl[AM_SWEEP_LAST_MED] = med;
  if (trace && k < trace_cap) {
    trace[4 * k + 0] = eps;
    trace[4 * k + 1] = (double)nc;
    trace[4 * k + 2] = med;
    trace[4 * k + 3] = ctl[AM_SWEEP_BEST];
  }
  ctl[AM_SWEEP_NSTEPS] = (double)(k + 1);
  rounds[nc] += 1;                                                                     // :199-202
  if (rounds[nc] > 10) ctl[AM_SWEEP_STEP] = am::mul_rn(ctl[AM_SWEEP_STEP], 10.0);      // :204-205
  if (med == 0.0) rounds[0] += 1;                                                      // :206-207
  if (rounds[0] >= 10) {                                                               // :208-209
    ctl[AM_SWEEP_DONE] = 1.0;
    return take;
  }
  const double next = am::add_rn(eps, ctl[AM_SWEEP_STEP]);                             // :215
  ctl[AM_SWEEP_EPS] = next;
  if (!(nc > 1)) {                                                                     // while n_clusters > 1 (:179)
    ctl[AM_SWEEP_DONE] = 1.0;
    return take;
  }
  am::eps_grid(next, ctl + AM_SWEEP_SPAN, F, ctl + AM_SWEEP_INV_CELL);
  return take;
}

struct InitVals {
  double span[4];
  int F;
};

__global__ void sweep_init_kernel(double *ctl, InitVals v, int *parent, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
  if (i == 0) {
    for (int w = 0; w < W; ++w) ctl[w] = 0.0;
    ctl[AM_SWEEP_EPS] = 0.3;   // recursive_dbscan.py:172-173
    ctl[AM_SWEEP_STEP] = 0.1;
    ctl[AM_SWEEP_BEST] = -INFINITY;
    ctl[AM_SWEEP_BEST_STEP] = -1.0;
    for (int d = 0; d < 4; ++d) ctl[AM_SWEEP_SPAN + d] = v.span[d];
    am::eps_grid(0.3, v.span, v.F, ctl + AM_SWEEP_INV_CELL);
  }
}

__global__ void sweep_control_kernel(double *ctl, int *rounds, const int *__restrict__ n_clusters,
                                     const double *__restrict__ summary, int F, double *trace, int trace_cap) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  sweep_advance(ctl, rounds, F, *n_clusters, summary[1], trace, trace_cap);
}

// keep the labels and the per-cluster metrics of the step the controller just declared the best so far
__global__ void sweep_keep_kernel(const double *__restrict__ ctl, int n, const int64_t *__restrict__ labels,
                                  int64_t *__restrict__ b_labels, const double *__restrict__ c0,
                                  const double *__restrict__ c1, const double *__restrict__ c2,
                                  const double *__restrict__ c3, const int *__restrict__ p,
                                  const int *__restrict__ s, double *__restrict__ b0, double *__restrict__ b1,
                                  double *__restrict__ b2, double *__restrict__ b3, int *__restrict__ bp,
                                  int *__restrict__ bs) {
  if (ctl[AM_SWEEP_TAKE] == 0.0) return;
  const int nc = (int)ctl[AM_SWEEP_LAST_NC];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    b_labels[i] = labels[i];
    if (i < nc) {
      b0[i] = c0[i];
      b1[i] = c1[i];
      b2[i] = c2[i];
      b3[i] = c3[i];
      bp[i] = p[i];
      bs[i] = s[i];
    }
  }
}

int check_args(const am_sweep_args *a, const char *who) {
  if (!a) return am::fail(AM_EINVAL, "%s: null args", who);
  if (a->F < 1 || a->F > 4) return am::fail(AM_EINVAL, "%s: F must be in 1..4", who);
  if (a->n_points < 1 || a->n_points > 0x7fffffff) return am::fail(AM_EINVAL, "%s: bad n_points", who);
  if (a->n_markers < 1 || a->nnz < 0 || a->nnz > 0x7fffffff) return am::fail(AM_EINVAL, "%s: bad marker table", who);
  if (a->trace_cap < 0) return am::fail(AM_EINVAL, "%s: bad trace_cap", who);
  if (!a->d_X || !a->d_parent || !a->d_dbscan_work || !a->d_labels || !a->d_n_clusters || !a->d_coverage ||
      !a->d_gc || !a->d_completeness || !a->d_purity || !a->d_cov_std || !a->d_gc_std || !a->d_present ||
      !a->d_single || !a->d_summary || !a->d_metrics_work || !a->d_ctl || !a->d_best_labels ||
      !a->d_best_completeness || !a->d_best_purity || !a->d_best_cov_std || !a->d_best_gc_std ||
      !a->d_best_present || !a->d_best_single)
    return am::fail(AM_EINVAL, "%s: null pointer", who);
  for (int d = 0; d < a->F; ++d)
    if (!(a->h_max[d] - a->h_min[d] >= 0.0)) return am::fail(AM_EINVAL, "%s: bad bounds (NaN?)", who);
  return AM_OK;
}

// one step: DBSCAN at the eps in d_ctl, labels, metrics, controller, keep-best
int queue_step(const am_sweep_args *a, cudaStream_t st) {
  char *base = (char *)a->d_ctl;
  double *ctl = (double *)base;
  const CtlLayout l = ctl_layout(a->n_points, a->trace_cap);
  int *rounds = (int *)(base + l.rounds_off);
  double *trace = a->trace_cap > 0 ? (double *)(base + l.trace_off) : nullptr;
  int rc = am_eps_union_dyn(a->d_X, a->n_points, a->F, a->h_min, ctl + AM_SWEEP_INV_CELL, a->d_parent,
                            a->d_dbscan_work, st);
  if (rc) return rc;
  rc = am_labels_from_parent(a->d_parent, a->n_points, a->d_labels, a->d_n_clusters, a->d_dbscan_work, st);
  if (rc) return rc;
  rc = am_cluster_metrics(a->d_labels, a->n_points, a->d_n_clusters, a->d_marker_rows, a->d_marker_cols,
                          a->d_marker_vals, a->nnz, a->n_markers, a->d_coverage, a->d_gc, a->cutoffs[0],
                          a->cutoffs[1], a->cutoffs[2], a->cutoffs[3], a->d_completeness, a->d_purity,
                          a->d_cov_std, a->d_gc_std, a->d_present, a->d_single, a->d_summary, a->d_metrics_work,
                          st);
  if (rc) return rc;
  sweep_control_kernel<<<1, 32, 0, st>>>(ctl, rounds, a->d_n_clusters, a->d_summary, a->F, trace, a->trace_cap);
  AM_LAUNCHED("sweep_control_kernel");
  const int n = (int)a->n_points;
  const int blocks = std::min((n + 255) / 256, 4 * am::num_sms());
  sweep_keep_kernel<<<blocks, 256, 0, st>>>(ctl, n, a->d_labels, a->d_best_labels, a->d_completeness, a->d_purity,
                                            a->d_cov_std, a->d_gc_std, a->d_present, a->d_single,
                                            a->d_best_completeness, a->d_best_purity, a->d_best_cov_std,
                                            a->d_best_gc_std, a->d_best_present, a->d_best_single);
  AM_LAUNCHED("sweep_keep_kernel");
  return AM_OK;
}

struct SweepGraph {
  cudaGraphExec_t exec;
  int64_t launches;  // kernels per replay (for am_launch_count)
};

}  // namespace

extern "C" int64_t am_sweep_args_size(void) { return (int64_t)sizeof(am_sweep_args); }

extern "C" int64_t am_sweep_ctl_bytes(int64_t n_points, int trace_cap) {
  if (n_points < 0 || trace_cap < 0) return 0;
  return ctl_layout(std::max<int64_t>(n_points, 1), trace_cap).total;
}

extern "C" int am_sweep_init(const am_sweep_args *a, void *stream) {
  int rc = check_args(a, "am_sweep_init");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const CtlLayout l = ctl_layout(a->n_points, a->trace_cap);
  AM_CUDA(cudaMemsetAsync((char *)a->d_ctl + l.rounds_off, 0, (size_t)(l.total - l.rounds_off), st));
  InitVals v;
  for (int d = 0; d < 4; ++d) v.span[d] = d < a->F ? a->h_max[d] - a->h_min[d] : 0.0;
  v.F = a->F;
  const int n = (int)a->n_points;
  sweep_init_kernel<<<(n + 255) / 256, 256, 0, st>>>((double *)a->d_ctl, v, a->d_parent, n);
  AM_LAUNCHED("sweep_init_kernel");
  return AM_OK;
}

extern "C" int am_sweep_steps(const am_sweep_args *a, int repeats, double *h_ctl, void *stream) {
  int rc = check_args(a, "am_sweep_steps");
  if (rc) return rc;
  if (repeats < 0) return am::fail(AM_EINVAL, "am_sweep_steps: repeats < 0");
  cudaStream_t st = (cudaStream_t)stream;
  for (int r = 0; r < repeats; ++r) {
    rc = queue_step(a, st);
    if (rc) return rc;
  }
  if (h_ctl) AM_CUDA(cudaMemcpyAsync(h_ctl, a->d_ctl, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
  return AM_OK;
}

extern "C" int am_sweep_graph_create(const am_sweep_args *a, void *stream, void **graph_out) {
  int rc = check_args(a, "am_sweep_graph_create");
  if (rc) return rc;
  if (!graph_out) return am::fail(AM_EINVAL, "am_sweep_graph_create: null graph_out");
  *graph_out = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread)
    return am::fail(AM_EINVAL, "am_sweep_graph_create: needs a stream of its own (the default stream cannot be captured)");
  (void)am::num_sms();  // device queries happen before the capture starts
  const int64_t l0 = am::g_launches.load();
  AM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  rc = queue_step(a, st);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  const int64_t per = am::g_launches.load() - l0;
  am::g_launches.fetch_sub(per);  // captured, not launched: counted per replay instead
  if (rc) {
    if (g) cudaGraphDestroy(g);
    (void)cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || !g) {
    (void)cudaGetLastError();
    return am::cuda_fail(e, "cudaStreamEndCapture");
  }
  cudaGraphExec_t ex = nullptr;
  const cudaError_t e2 = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e2 != cudaSuccess) {
    (void)cudaGetLastError();
    return am::cuda_fail(e2, "cudaGraphInstantiate");
  }
  SweepGraph *sg = new SweepGraph{ex, per};
  *graph_out = sg;
  return AM_OK;
}

extern "C" int am_sweep_graph_launch(void *graph, const am_sweep_args *a, int repeats, double *h_ctl, void *stream) {
  if (!graph || !a || !a->d_ctl) return am::fail(AM_EINVAL, "am_sweep_graph_launch: null pointer");
  if (repeats < 0) return am::fail(AM_EINVAL, "am_sweep_graph_launch: repeats < 0");
  SweepGraph *sg = (SweepGraph *)graph;
  cudaStream_t st = (cudaStream_t)stream;
  for (int r = 0; r < repeats; ++r) AM_CUDA(cudaGraphLaunch(sg->exec, st));
  am::g_launches.fetch_add(sg->launches * repeats);
  if (h_ctl) AM_CUDA(cudaMemcpyAsync(h_ctl, a->d_ctl, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
  return AM_OK;
}

extern "C" int am_sweep_graph_destroy(void *graph) {
  if (!graph) return AM_OK;
  SweepGraph *sg = (SweepGraph *)graph;
  const cudaError_t e = cudaGraphExecDestroy(sg->exec);
  delete sg;
  if (e != cudaSuccess) return am::cuda_fail(e, "cudaGraphExecDestroy");
  return AM_OK;
}

extern "C" int am_sweep_advance_host(double *ctl, int32_t *rounds, int F, int n_clusters, double median) {
  if (!ctl || !rounds || F < 1 || F > 4 || n_clusters < 0) return am::fail(AM_EINVAL, "am_sweep_advance_host: bad args");
  return sweep_advance(ctl, rounds, F, n_clusters, median, nullptr, 0);
}

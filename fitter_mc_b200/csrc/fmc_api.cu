// C ABI of libfmc_b200.so (declared in include/fmc_b200.h).  Host-side plumbing only:
// device buffers, stream, launch geometry, decoding of the accumulator block.  All the
// arithmetic of the hot path is in bootstrap_kernel.cuh / nl2_solver.cuh / pass_eval.cuh.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>  // types and constants only: the library is loaded at run time (see NcclApi below)

#include "../../include/fmc_b200.h"
#include "model_registry.cuh"
#include "round_kernels.cuh"
#include "warp_kernel.cuh"

// (no `using namespace fmc` here: fmc::exp & co. must not leak into the global scope, where
// they would collide with the CUDA math headers)
using fmc::acc_len;
using fmc::ACC_NFCALL;
using fmc::ACC_NFIT;
using fmc::ACC_NGCALL;
using fmc::ACC_NITER;
using fmc::ACC_NPASS;
using fmc::ACC_RC;
using fmc::KernelArgs;
using fmc::ModelEntry;
using fmc::normals4;
using fmc::RowData;

namespace {

thread_local std::string g_last_cuda_error;
int g_default_jac = 0;  // Jacobian mode of the one-call entry points (fmc_set_jacobian_mode(NULL, .))
int g_default_mapping = 0;  // fit mapping of the one-call entry points (fmc_set_fit_mapping(NULL, .))
int g_default_schedule = -1;  // schedule of the one-call entry points (fmc_set_schedule(NULL, .)); -1 = FMC_SCHEDULE env / refill

#define FMC_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      g_last_cuda_error = std::string(#call) + ": " + cudaGetErrorString(e_);            \
      return FMC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// a kernel launch of the fit pipeline: counted, so that callers can report how many of this
// library's kernels ran (fmc_last_launch_count)
#define FMC_LAUNCH(call) \
  do {                   \
    FMC_CUDA(call);      \
    ++c->nlaunch;        \
  } while (0)

// temporary device buffer, released on every return path
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

// Make `device` current for the duration of an entry point and give the caller's device back
// afterwards: the library is loaded into processes (torch, the Fortran driver's threads) that
// have their own current device.
struct DeviceGuard {
  int prev = -1;
  int dev;
  cudaError_t err;
  explicit DeviceGuard(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

// NCCL, bound at run time: the moment accumulators of the GPUs of one box are combined by ONE
// ncclAllReduce over NVLink, replacing the OpenMP REDUCTION(+) of
// bootstrap.f90:310-311.  dlopen instead of a link-time dependency, so that a process that has
// already loaded an NCCL (torch ships its own libnccl.so.2) shares that copy, and a single-GPU
// host without NCCL can still load libfmc_b200.so.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
    auto sym = [&](const char* n) { return dlsym(a.handle, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

#define FMC_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r_ = (call);                                                            \
    if (r_ != ncclSuccess) {                                                             \
      g_last_cuda_error = std::string(#call) + ": " + nccl().GetErrorString(r_);         \
      return FMC_ERR_NCCL;                                                               \
    }                                                                                    \
  } while (0)

int g_last_combine = FMC_COMBINE_NONE;  // how the last multi-GPU one-call run combined its accumulators

const std::vector<ModelEntry>& models() {
  static const std::vector<ModelEntry> m = {fmc::fmc_entry_model0(), fmc::fmc_entry_model1(),
                                            fmc::fmc_entry_model2(), fmc::fmc_entry_model3()};
  return m;
}

int device_count() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    g_last_cuda_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
    return 0;
  }
  return n;
}

__global__ void generate_offsets_kernel(unsigned long long seed, long long first, long long nres, int n,
                                        double* __restrict__ z) {
  const int nblk = (n + 3) / 4;
  const long long total = nres * (long long)nblk;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / nblk;
    const int b = (int)(t - r * nblk);
    float zz[4];
    normals4(seed, (unsigned long long)(first + r), (uint32_t)b, zz);
    for (int m = 0; m < 4 && 4 * b + m < n; ++m) z[r * n + 4 * b + m] = (double)zz[m];
  }
}

// Raw words of the counter-based generator for known-answer tests on the device (fmc_debug_philox4x32).
__global__ void philox_debug_kernel(const uint32_t* __restrict__ in6, int n, uint32_t* __restrict__ out4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fmc::U4 c;
  c.x = in6[6 * i];
  c.y = in6[6 * i + 1];
  c.z = in6[6 * i + 2];
  c.w = in6[6 * i + 3];
  const fmc::U4 o = fmc::philox4x32_10(c, in6[6 * i + 4], in6[6 * i + 5]);
  out4[4 * i] = o.x;
  out4[4 * i + 1] = o.y;
  out4[4 * i + 2] = o.z;
  out4[4 * i + 3] = o.w;
}

// Register-resident DFMA chains: the measured FP64 roofline denominator (MEASURED_PEAKS.json
// has no FP64 entry; SURVEY.md section 8d asks for this microbenchmark).
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0,
         x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Binning runs as its own small kernel over the per-fit results of a chunk, so that the fit
// kernels are the same machine code whether or not distributions are requested.  Traffic:
// 8 (P+Q) B read per fit and one 64-bit reduction per quantity -- three orders of magnitude
// below the fit kernels' time.  Counts are integers: the result does not depend on the order
// in which fits are binned or GPUs are summed.
__global__ void __launch_bounds__(256) hist_kernel(const fmc::HistArgs h) {
  const int nquant = h.np + h.nq;
  const long long total = h.nres * nquant;
  const int stride = h.bins + 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    // element e of the concatenation [params | props]: consecutive threads read consecutive doubles
    const bool is_param = e < h.nres * h.np;
    const long long idx = is_param ? e : e - h.nres * h.np;
    const int j = is_param ? (int)(idx % h.np) : h.np + (int)(idx % h.nq);
    const double v = is_param ? h.params[idx] : h.props[idx];
    atomicAdd(h.hist + (size_t)j * stride + fmc::hist_bin(v, h.lo[j], h.scale[j], h.bins), 1ULL);
  }
}

}  // namespace

struct fmc_context {
  int device = 0;
  int num_sms = 0;
  int max_smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int model = -1, n = 0, n_pad = 0;
  RowData* rows_dev = nullptr;
  int rows_cap = 0;
  // the same data once more in lane-major column order for the warp-per-fit kernel (warp_kernel.cuh)
  double* cols_dev = nullptr;   // [4 + NAUX][n_w]: x, y, err, 1/err, the model's per-row auxiliary values
  double* aux_dev = nullptr;    // [n_pad][NAUX]: the same auxiliary values in row order (thread-per-fit kernels)
  int aux_cap = 0;
  int cols_cap = 0, n_w = 0;
  double* tscratch_dev = nullptr;  // [warps of the grid][n_w]: t_i = r_i / err_i between the two sweeps of a pass
  long long tscratch_cap = 0;
  double* acc_own = nullptr;
  double* acc = nullptr;
  int acc_n = 0;
  double* params_out_dev = nullptr;
  double* props_out_dev = nullptr;
  int* counters_dev_out = nullptr;
  long long out_cap_p = 0, out_cap_q = 0, out_cap_c = 0;  // capacities in elements
  // pipeline state between the kernels of a chunk (bootstrap_kernel.cuh)
  double* xbuf = nullptr;
  double* initbuf = nullptr;
  int* istate = nullptr;
  int* cnt = nullptr;
  long long chunk_cap = 0;
  int chunk_p = 0;
  unsigned long long* counters_dev = nullptr;  // one work counter per iterate launch
  double* rowj_dev = nullptr;  // [n_pad][1+P] shared Jacobian rows at params_start + [PP] G0
  unsigned long long* anom_dev = nullptr;  // anomaly log of the last launch
  unsigned long long* hist_own = nullptr;  // binned distributions (optional)
  unsigned long long* hist = nullptr;
  long long hist_own_cap = 0;
  int hist_bins = 0;
  ncclComm_t comm = nullptr;  // communicator of this GPU's rank (fmc_nccl_init_rank / the one-call multi-GPU run)
  bool comm_owned = false;
  int comm_nranks = 1;
  int jac = fmc::JAC_FD;  // how Jacobian rows are formed (fmc_set_jacobian_mode)
  int mapping = 0;        // 0 = one fit per thread, 1 = one fit per warp (fmc_set_fit_mapping)
  int schedule = FMC_SCHEDULE_AUTO;  // FMC_SCHEDULE_REFILL, _ROUNDS or _AUTO (fmc_set_schedule)
  int schedule_used = FMC_SCHEDULE_REFILL;  // what the last launch resolved AUTO to
  // ---- round schedule (round_kernels.cuh): slot pool, lists, control words, the WHILE graph
  fmc::RoundArgs rs{0, nullptr, nullptr, nullptr, nullptr, nullptr};
  long long rs_state_cap = 0;            // 64-bit words allocated for rs.state
  long long rs_slot_cap = 0;             // slots allocated for slot_fit / slot_meta / list
  fmc::KernelArgs* rs_kargs_dev = nullptr;     // the chunk's KernelArgs, read by the graph's kernels
  fmc::KernelArgs* rs_kargs_pinned = nullptr;  // one per chunk of the launch in flight
  long long rs_kargs_cap = 0;
  unsigned int* rs_ctl_pinned = nullptr;
  cudaStream_t rs_capture_stream = nullptr;
  cudaGraph_t rs_graph = nullptr;
  cudaGraphExec_t rs_exec = nullptr;
  bool rs_no_graph = false;              // FMC_ROUNDS_NO_GRAPH=1 or the graph could not be built: host-driven rounds
  long long rs_key[8] = {-1, -1, -1, -1, -1, -1, -1, -1};  // what rs_exec was built for
  long long rs_rounds = 0;               // rounds enqueued by the host-driven fallback
  bool rs_used = false;                  // the launch in flight ran under the round schedule
  double hist_lo[fmc::HIST_MAX_QUANT] = {0};
  double hist_scale[fmc::HIST_MAX_QUANT] = {0};
  long long rowj_cap = 0;
  int ncounters = 0;
  long long inflight = -1;
  long long nlaunch = 0;  // kernels launched by the last fmc_launch
  bool keep = false;
  double start[16] = {0};
  int geom[4] = {0, 0, 0, 0};
  std::vector<RowData> rows_host;
};

// ---------------------------------------------------------------------------------------
// Round schedule (round_kernels.cuh): pool allocation, the WHILE graph, one chunk's rounds.
static int default_schedule() {
  static const int v = [] {
    const char* e = std::getenv("FMC_SCHEDULE");
    if (e && (e[0] == 'r' || e[0] == 'R') && (e[1] == 'o' || e[1] == 'O')) return FMC_SCHEDULE_ROUNDS;
    if (e && (e[0] == 'r' || e[0] == 'R') && (e[1] == 'e' || e[1] == 'E')) return FMC_SCHEDULE_REFILL;
    return FMC_SCHEDULE_AUTO;
  }();
  return v;
}

static int rounds_prepare(fmc_context* c, const fmc::RoundEntry& re, int P, bool inject, bool smem, int jac,
                          size_t shmem, long long cap, long long nchunks) {
  int bps = 0;
  FMC_CUDA(re.prepare(inject, smem, jac, shmem, &bps));
  if (bps < 1) {
    g_last_cuda_error = "pass kernel does not fit on an SM";
    return FMC_ERR_CUDA;
  }
  // pool: enough slots for several waves of the pass kernel, so that the solve kernel that follows
  // runs at full occupancy and the partial last wave of a round is a small fraction of it
  long long waves = 8;
  if (const char* e = std::getenv("FMC_ROUND_WAVES")) {
    const long long v = std::atoll(e);
    if (v >= 1 && v <= 4096) waves = v;
  }
  long long slots = (long long)c->num_sms * bps * FMC_ROUND_PASS_THREADS * waves;
  if (slots > cap) slots = cap;
  if (slots < 1) slots = 1;
  c->geom[0] = (int)(slots / FMC_ROUND_PASS_THREADS + 2);
  c->geom[1] = FMC_ROUND_PASS_THREADS;
  c->geom[2] = bps;
  c->geom[3] = c->num_sms;
  const long long words = (long long)re.state_words * slots;
  if (words > c->rs_state_cap) {
    cudaFree(c->rs.state);
    c->rs.state = nullptr;
    c->rs_state_cap = 0;
    FMC_CUDA(cudaMalloc(&c->rs.state, sizeof(unsigned long long) * words));
    c->rs_state_cap = words;
  }
  if (slots > c->rs_slot_cap) {
    cudaFree(c->rs.slot_fit);
    cudaFree(c->rs.slot_meta);
    cudaFree(c->rs.list);
    c->rs.slot_fit = nullptr;
    c->rs.slot_meta = nullptr;
    c->rs.list = nullptr;
    c->rs_slot_cap = 0;
    FMC_CUDA(cudaMalloc(&c->rs.slot_fit, sizeof(long long) * slots));
    FMC_CUDA(cudaMalloc(&c->rs.slot_meta, sizeof(int) * 2 * slots));
    FMC_CUDA(cudaMalloc(&c->rs.list, sizeof(int) * 2 * fmc::RC_KINDS * slots));
    c->rs_slot_cap = slots;
  }
  if (!c->rs.ctl) {
    FMC_CUDA(cudaMalloc(&c->rs.ctl, sizeof(unsigned int) * fmc::RC_NWORDS));
    FMC_CUDA(cudaMalloc(&c->rs_kargs_dev, sizeof(fmc::KernelArgs)));
    FMC_CUDA(cudaMallocHost(&c->rs_ctl_pinned, sizeof(unsigned int) * fmc::RC_NWORDS));
    FMC_CUDA(cudaStreamCreateWithFlags(&c->rs_capture_stream, cudaStreamNonBlocking));
    const char* e = std::getenv("FMC_ROUNDS_NO_GRAPH");
    c->rs_no_graph = e && e[0] == '1';
  }
  if (nchunks > c->rs_kargs_cap) {
    if (c->rs_kargs_pinned) cudaFreeHost(c->rs_kargs_pinned);
    c->rs_kargs_pinned = nullptr;
    c->rs_kargs_cap = 0;
    FMC_CUDA(cudaMallocHost(&c->rs_kargs_pinned, sizeof(fmc::KernelArgs) * nchunks));
    c->rs_kargs_cap = nchunks;
  }
  c->rs.nslots = slots;
  // Residual sum first (Nl2::spec = 0) where a Jacobian row is expensive next to the model value -- many
  // parameters -- and the passes are what a round costs; with few parameters the extra solve-kernel visit
  // of every accepted step costs more than the Jacobians saved (measured: DESIGN.md 2.8).
  c->rs.spec = P > 6 ? 0 : 1;
  if (const char* e = std::getenv("FMC_ROUND_SPECULATE")) c->rs.spec = (e[0] == '0') ? 0 : 1;
  FMC_CUDA(cudaMemsetAsync(c->rs.ctl, 0, sizeof(unsigned int) * fmc::RC_NWORDS, c->stream));
  c->rs_rounds = 0;

  // the WHILE graph: body = one round; rebuilt only when what it was captured for changes
  const long long key[8] = {c->model, (inject ? 1 : 0) | (smem ? 2 : 0) | (jac << 2) | (c->rs.spec << 4), (long long)shmem, slots,
                            (long long)(size_t)c->rs.state, (long long)(size_t)c->rs.slot_fit,
                            (long long)(size_t)c->rs.list, (long long)(size_t)c->rs_kargs_dev};
  (void)P;
  if (c->rs_no_graph) return FMC_OK;
  bool same = c->rs_exec != nullptr;
  for (int i = 0; i < 8; ++i) same = same && key[i] == c->rs_key[i];
  if (same) return FMC_OK;
  if (c->rs_exec) cudaGraphExecDestroy(c->rs_exec);
  if (c->rs_graph) cudaGraphDestroy(c->rs_graph);
  c->rs_exec = nullptr;
  c->rs_graph = nullptr;
  auto give_up = [&](const char* what, cudaError_t e) {
    // no conditional graph nodes (old driver): host-driven rounds, same kernels
    g_last_cuda_error = std::string("round graph: ") + what + ": " + cudaGetErrorString(e);
    if (c->rs_exec) cudaGraphExecDestroy(c->rs_exec);
    if (c->rs_graph) cudaGraphDestroy(c->rs_graph);
    c->rs_exec = nullptr;
    c->rs_graph = nullptr;
    c->rs_no_graph = true;
    cudaGetLastError();
    return FMC_OK;
  };
  cudaError_t e = cudaGraphCreate(&c->rs_graph, 0);
  if (e != cudaSuccess) return give_up("cudaGraphCreate", e);
  cudaGraphConditionalHandle handle;
  e = cudaGraphConditionalHandleCreate(&handle, c->rs_graph, 1, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) return give_up("cudaGraphConditionalHandleCreate", e);
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  e = cudaGraphAddNode(&node, c->rs_graph, nullptr, 0, &np);
  if (e != cudaSuccess) return give_up("cudaGraphAddNode(conditional)", e);
  cudaGraph_t body = np.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(c->rs_capture_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) return give_up("cudaStreamBeginCaptureToGraph", e);
  e = re.enqueue_round(c->rs_kargs_dev, c->rs, inject, smem, jac, shmem, handle, 1, c->rs_capture_stream);
  cudaGraph_t captured = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(c->rs_capture_stream, &captured);
  if (e != cudaSuccess) return give_up("capture of the round kernels", e);
  if (e2 != cudaSuccess) return give_up("cudaStreamEndCapture", e2);
  e = cudaGraphInstantiate(&c->rs_exec, c->rs_graph, 0);
  if (e != cudaSuccess) return give_up("cudaGraphInstantiate", e);
  for (int i = 0; i < 8; ++i) c->rs_key[i] = key[i];
  return FMC_OK;
}

// The rounds of one chunk: `a` (host copy, pinned, entry ch) is already complete.
static int rounds_run_chunk(fmc_context* c, const fmc::RoundEntry& re, long long ch, bool inject, bool smem, int jac,
                            size_t shmem) {
  FMC_CUDA(cudaMemcpyAsync(c->rs_kargs_dev, &c->rs_kargs_pinned[ch], sizeof(fmc::KernelArgs), cudaMemcpyHostToDevice,
                           c->stream));
  // round counter, list lengths and the done flag back to zero (the total keeps counting)
  FMC_CUDA(cudaMemsetAsync(c->rs.ctl, 0, sizeof(unsigned int) * fmc::RC_ROUNDS_TOTAL, c->stream));
  if (!c->rs_no_graph) {
    FMC_CUDA(cudaGraphLaunch(c->rs_exec, c->stream));
    return FMC_OK;
  }
  // host-driven fallback: batches of rounds, the done flag read back between batches (blocks the host)
  const long long max_rounds = 4LL * (fmc::k::mxfcal + fmc::k::mxiter) + 16;
  long long done_rounds = 0, batch = 8;
  while (done_rounds < max_rounds) {
    for (long long r = 0; r < batch; ++r) {
      FMC_CUDA(re.enqueue_round(c->rs_kargs_dev, c->rs, inject, smem, jac, shmem, 0, 0, c->stream));
      ++c->rs_rounds;
    }
    done_rounds += batch;
    FMC_CUDA(cudaMemcpyAsync(c->rs_ctl_pinned, c->rs.ctl, sizeof(unsigned int) * fmc::RC_NWORDS, cudaMemcpyDeviceToHost,
                             c->stream));
    FMC_CUDA(cudaStreamSynchronize(c->stream));
    if (c->rs_ctl_pinned[fmc::RC_DONE] != 0u) return FMC_OK;
    if (batch < 64) batch *= 2;
  }
  g_last_cuda_error = "round schedule: fits still iterating after the cap on rounds";
  return FMC_ERR_CUDA;
}

extern "C" {

const char* fmc_version(void) { return "fmc_b200 0.1 (sm_100a)"; }

const char* fmc_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

const char* fmc_strerror(int code) {
  switch (code) {
    case FMC_OK: return "success";
    case FMC_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case FMC_ERR_CUDA: return "a CUDA call failed (see fmc_last_cuda_error)";
    case FMC_ERR_BAD_MODEL: return "model id out of range";
    case FMC_ERR_UNSUPPORTED: return "not part of this build of the library";
    case FMC_ERR_BAD_ARG: return "invalid argument";
    case FMC_ERR_NO_DATA: return "fmc_set_data has not been called";
    case FMC_ERR_STATE: return "no launch in flight";
    case FMC_ERR_ALLOC: return "allocation failed";
    case FMC_ERR_NCCL: return "an NCCL call failed or NCCL could not be loaded (see fmc_last_cuda_error)";
    default: return "unknown error code";
  }
}

int fmc_model_count(void) { return (int)models().size(); }

int fmc_initialise_model(int model, int* nparam, int* nprop, double* params_init, char (*param_name)[3],
                         char (*prop_name)[3], char* model_name32) {
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  const ModelEntry& m = models()[model];
  if (nparam) *nparam = m.nparam;
  if (nprop) *nprop = m.nprop;
  for (int i = 0; i < m.nparam; ++i) {
    if (params_init) params_init[i] = m.init[i];
    if (param_name) {
      std::snprintf(param_name[i], 3, "%-2.2s", m.pnames[i] ? m.pnames[i] : "");
    }
  }
  for (int i = 0; i < m.nprop; ++i)
    if (prop_name) std::snprintf(prop_name[i], 3, "%-2.2s", m.qnames[i] ? m.qnames[i] : "");
  if (model_name32) std::snprintf(model_name32, 32, "%s", m.name);
  return FMC_OK;
}

double fmc_model_y(int model, const double* params, double x) {
  if (model < 0 || model >= fmc_model_count()) return 0.0;
  return models()[model].model_y(params, x);
}

int fmc_derived_properties(int model, const double* params, double* props) {
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  models()[model].props(params, props);
  return FMC_OK;
}

void fmc_finalise_moments(int k, long long nresample, const double* sum, const double* sum2, double* mean,
                          double* err) {
  const double rtemp = 1.0 / (double)nresample;  // bootstrap.f90:328-330
  for (int i = 0; i < k; ++i) {
    const double m = sum[i] * rtemp, m2 = sum2[i] * rtemp;
    mean[i] = m;
    if (nresample > 1) {
      const double bessel = (double)nresample / (double)(nresample - 1);  // bootstrap.f90:334-337
      err[i] = std::sqrt(std::fmax((m2 - m * m) * bessel, 0.0));
    } else {
      err[i] = 0.0;
    }
  }
}

int fmc_create(fmc_context** out, int device) {
  if (!out) return FMC_ERR_BAD_ARG;
  *out = nullptr;
  const int ndev = device_count();
  if (ndev <= 0) return FMC_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return FMC_ERR_BAD_ARG;
  DeviceGuard guard_(device);
  FMC_CUDA(guard_.err);
  // owned until every allocation below has succeeded: a failure frees what was created so far
  std::unique_ptr<fmc_context, int (*)(fmc_context*)> owner(new (std::nothrow) fmc_context(), fmc_destroy);
  fmc_context* c = owner.get();
  if (!c) return FMC_ERR_ALLOC;
  c->device = device;
  cudaDeviceProp prop;
  FMC_CUDA(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  FMC_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  FMC_CUDA(cudaEventCreate(&c->ev0));
  FMC_CUDA(cudaEventCreate(&c->ev1));
  c->schedule = default_schedule();
  c->acc_n = acc_len(FMC_MAX_PARAM, FMC_MAX_PARAM);
  FMC_CUDA(cudaMalloc(&c->acc_own, sizeof(double) * c->acc_n));
  c->acc = c->acc_own;
  FMC_CUDA(cudaMalloc(&c->anom_dev, sizeof(unsigned long long) * (1 + 2 * fmc::ANOM_CAP)));
  FMC_CUDA(cudaMemset(c->anom_dev, 0, sizeof(unsigned long long) * (1 + 2 * fmc::ANOM_CAP)));
  *out = owner.release();
  return FMC_OK;
}

int fmc_destroy(fmc_context* c) {
  if (!c) return FMC_OK;
  DeviceGuard guard_(c->device);
  if (c->inflight >= 0) cudaStreamSynchronize(c->stream);
  if (c->comm && c->comm_owned && nccl().ok) nccl().CommDestroy(c->comm);
  cudaFree(c->rows_dev);
  cudaFree(c->cols_dev);
  cudaFree(c->aux_dev);
  cudaFree(c->tscratch_dev);
  cudaFree(c->acc_own);
  cudaFree(c->xbuf);
  cudaFree(c->initbuf);
  cudaFree(c->istate);
  cudaFree(c->cnt);
  cudaFree(c->counters_dev);
  cudaFree(c->rowj_dev);
  cudaFree(c->anom_dev);
  cudaFree(c->hist_own);
  cudaFree(c->params_out_dev);
  cudaFree(c->props_out_dev);
  cudaFree(c->counters_dev_out);
  if (c->rs_exec) cudaGraphExecDestroy(c->rs_exec);
  if (c->rs_graph) cudaGraphDestroy(c->rs_graph);
  if (c->rs_capture_stream) cudaStreamDestroy(c->rs_capture_stream);
  cudaFree(c->rs.state);
  cudaFree(c->rs.slot_fit);
  cudaFree(c->rs.slot_meta);
  cudaFree(c->rs.list);
  cudaFree(c->rs.ctl);
  cudaFree(c->rs_kargs_dev);
  if (c->rs_kargs_pinned) cudaFreeHost(c->rs_kargs_pinned);
  if (c->rs_ctl_pinned) cudaFreeHost(c->rs_ctl_pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return FMC_OK;
}

int fmc_set_stream(fmc_context* c, void* cuda_stream) {
  if (!c) return FMC_ERR_BAD_ARG;
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return FMC_OK;
}

int fmc_set_data(fmc_context* c, int model, int ndata, const double* x, const double* y, const double* err_y) {
  if (!c || !x || !y || !err_y) return FMC_ERR_BAD_ARG;
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  const ModelEntry& m = models()[model];
  if (ndata < m.nparam || ndata < 1) return FMC_ERR_BAD_ARG;
  for (int i = 0; i < ndata; ++i)
    if (!(err_y[i] > 0.0)) return FMC_ERR_BAD_ARG;  // bootstrap.f90:134-138
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  const int n_pad = (ndata + 3) / 4 * 4;
  c->rows_host.resize(n_pad);
  for (int i = 0; i < n_pad; ++i) {
    RowData r;
    if (i < ndata) {
      r.x = x[i];
      r.y = y[i];
      r.err = err_y[i];
      r.rerr = 1.0 / err_y[i];  // bootstrap.f90:140
    } else {  // padding rows contribute exactly zero to every sum
      r.x = x[ndata - 1];
      r.y = 0.0;
      r.err = 0.0;
      r.rerr = 0.0;
    }
    c->rows_host[i] = r;
  }
  if (n_pad > c->rows_cap) {
    cudaFree(c->rows_dev);
    c->rows_dev = nullptr;
    FMC_CUDA(cudaMalloc(&c->rows_dev, sizeof(RowData) * n_pad));
    c->rows_cap = n_pad;
  }
  FMC_CUDA(cudaMemcpyAsync(c->rows_dev, c->rows_host.data(), sizeof(RowData) * n_pad, cudaMemcpyHostToDevice,
                           c->stream));
  // lane-major columns: element s = (4 j + m) 32 + l of a column is row 4 (32 j + l) + m, so that lane l
  // of a warp owns the four-row groups l, l + 32, ... and every warp-wide load is contiguous
  // per-row auxiliary values the model declares (functions of x_i alone, e.g. log x_i): once per data set
  const int naux = m.naux;
  std::vector<double> aux((size_t)(naux > 0 ? naux : 1) * n_pad, 0.0);
  for (int i = 0; i < n_pad && naux > 0; ++i) m.row_aux(c->rows_host[i].x, &aux[(size_t)i * naux]);
  if (naux > 0) {
    if (naux * n_pad > c->aux_cap) {
      cudaFree(c->aux_dev);
      c->aux_dev = nullptr;
      c->aux_cap = 0;
      FMC_CUDA(cudaMalloc(&c->aux_dev, sizeof(double) * naux * n_pad));
      c->aux_cap = naux * n_pad;
    }
    FMC_CUDA(cudaMemcpyAsync(c->aux_dev, aux.data(), sizeof(double) * naux * n_pad, cudaMemcpyHostToDevice, c->stream));
  }
  const int n_w = (ndata + 127) / 128 * 128;
  const int ncol = 4 + naux;
  std::vector<double> cols((size_t)ncol * n_w);
  for (int sidx = 0; sidx < n_w; ++sidx) {
    const int l = sidx & 31, jm = sidx >> 5, mm = jm & 3, j = jm >> 2;
    const int row = 4 * (32 * j + l) + mm;
    const bool real = row < ndata;
    cols[sidx] = real ? x[row] : x[ndata - 1];
    cols[(size_t)n_w + sidx] = real ? y[row] : 0.0;
    cols[(size_t)2 * n_w + sidx] = real ? err_y[row] : 0.0;
    cols[(size_t)3 * n_w + sidx] = real ? 1.0 / err_y[row] : 0.0;
    for (int k = 0; k < naux; ++k) cols[(size_t)(4 + k) * n_w + sidx] = aux[(size_t)(real ? row : ndata - 1) * naux + k];
  }
  if (ncol * n_w > c->cols_cap) {
    cudaFree(c->cols_dev);
    c->cols_dev = nullptr;
    c->cols_cap = 0;
    FMC_CUDA(cudaMalloc(&c->cols_dev, sizeof(double) * ncol * n_w));
    c->cols_cap = ncol * n_w;
  }
  FMC_CUDA(cudaMemcpyAsync(c->cols_dev, cols.data(), sizeof(double) * ncol * n_w, cudaMemcpyHostToDevice, c->stream));
  c->n_w = n_w;
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  if (model != c->model) c->hist_bins = 0;  // the binned quantities belong to a model
  c->model = model;
  c->n = ndata;
  c->n_pad = n_pad;
  return FMC_OK;
}

int fmc_set_jacobian_mode(fmc_context* c, int mode) {
  if (mode != FMC_JACOBIAN_FORWARD_DIFFERENCE && mode != FMC_JACOBIAN_ANALYTIC) return FMC_ERR_BAD_ARG;
  if (c)
    c->jac = mode;
  else
    g_default_jac = mode;  // the one-call entry points
  return FMC_OK;
}

int fmc_set_fit_mapping(fmc_context* c, int mapping) {
  if (mapping != FMC_FIT_PER_THREAD && mapping != FMC_FIT_PER_WARP) return FMC_ERR_BAD_ARG;
#ifndef FMC_WITH_WARP_MAPPING
  // The warp-per-fit kernels lost the measurement on every configuration (profiles/r2_mapping_*.json:
  // 0.47-0.70 of the thread-per-fit rate on C4, 0.47 on C5) and are not part of the default build.
  if (mapping == FMC_FIT_PER_WARP) return FMC_ERR_UNSUPPORTED;
#endif
  if (c)
    c->mapping = mapping;
  else
    g_default_mapping = mapping;  // the one-call entry points
  return FMC_OK;
}

int fmc_set_schedule(fmc_context* c, int schedule) {
  if (schedule != FMC_SCHEDULE_REFILL && schedule != FMC_SCHEDULE_ROUNDS && schedule != FMC_SCHEDULE_AUTO)
    return FMC_ERR_BAD_ARG;
  if (c)
    c->schedule = schedule;
  else
    g_default_schedule = schedule;  // the one-call entry points
  return FMC_OK;
}

int fmc_get_schedule(const fmc_context* c) {
  if (!c) return g_default_schedule >= 0 ? g_default_schedule : default_schedule();
  return c->schedule == FMC_SCHEDULE_AUTO ? c->schedule_used : c->schedule;  // AUTO: what the last launch chose
}

int fmc_model_has_analytic_jacobian(int model) {
  if (model < 0 || model >= fmc_model_count()) return 0;
  return models()[model].analytic ? 1 : 0;
}

int fmc_set_histogram(fmc_context* c, int nbins, const double* lo, const double* hi) {
  if (!c || nbins < 0 || nbins > (1 << 24)) return FMC_ERR_BAD_ARG;
  if (nbins == 0) {
    c->hist_bins = 0;
    return FMC_OK;
  }
  if (c->model < 0) return FMC_ERR_NO_DATA;
  if (!lo || !hi) return FMC_ERR_BAD_ARG;
  const ModelEntry& m = models()[c->model];
  const int nq = m.nparam + m.nprop;
  if (nq > fmc::HIST_MAX_QUANT) return FMC_ERR_BAD_ARG;
  for (int j = 0; j < nq; ++j)
    if (!(hi[j] > lo[j]) || !std::isfinite(lo[j]) || !std::isfinite(hi[j])) return FMC_ERR_BAD_ARG;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  if (c->inflight >= 0) FMC_CUDA(cudaStreamSynchronize(c->stream));
  const long long need = (long long)nq * (nbins + 2);
  if (need > c->hist_own_cap) {
    const bool was_own = c->hist == c->hist_own;
    cudaFree(c->hist_own);
    c->hist_own = nullptr;
    c->hist_own_cap = 0;
    FMC_CUDA(cudaMalloc(&c->hist_own, sizeof(unsigned long long) * need));
    c->hist_own_cap = need;
    if (was_own) c->hist = c->hist_own;
  }
  if (!c->hist) c->hist = c->hist_own;
  for (int j = 0; j < nq; ++j) {
    c->hist_lo[j] = lo[j];
    c->hist_scale[j] = (double)nbins / (hi[j] - lo[j]);
  }
  c->hist_bins = nbins;
  return FMC_OK;
}

long long fmc_histogram_len(const fmc_context* c) {
  if (!c || c->model < 0 || c->hist_bins <= 0) return 0;
  const ModelEntry& m = models()[c->model];
  return (long long)(m.nparam + m.nprop) * (c->hist_bins + 2);
}

int fmc_set_histogram_dev(fmc_context* c, unsigned long long* counts_dev) {
  if (!c) return FMC_ERR_BAD_ARG;
  c->hist = counts_dev ? counts_dev : c->hist_own;
  return FMC_OK;
}

int fmc_get_histogram(fmc_context* c, unsigned long long* counts_host) {
  if (!c || !counts_host) return FMC_ERR_BAD_ARG;
  const long long len = fmc_histogram_len(c);
  if (len <= 0 || !c->hist) return FMC_ERR_STATE;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  FMC_CUDA(cudaMemcpyAsync(counts_host, c->hist, sizeof(unsigned long long) * len, cudaMemcpyDeviceToHost,
                           c->stream));
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  return FMC_OK;
}

long long fmc_histogram_quantiles(const unsigned long long* counts, int nbins, double lo, double hi,
                                  const double* probs, int nprobs, double* out) {
  if (!counts || nbins <= 0 || !(hi > lo) || nprobs < 0 || (nprobs > 0 && (!probs || !out))) return -1;
  unsigned long long total = 0;
  for (int b = 0; b < nbins + 2; ++b) total += counts[b];
  const double width = (hi - lo) / (double)nbins;
  for (int i = 0; i < nprobs; ++i) {
    if (total == 0 || !(probs[i] >= 0.0) || !(probs[i] <= 1.0)) {
      out[i] = std::nan("");
      continue;
    }
    const double target = probs[i] * (double)total;  // resamples below the quantile
    double below = (double)counts[0];
    if (target <= below && counts[0] > 0) {
      out[i] = lo;
      continue;
    }
    double q = hi;
    for (int b = 1; b <= nbins; ++b) {
      const double in_bin = (double)counts[b];
      if (in_bin > 0.0 && target <= below + in_bin) {
        q = lo + width * ((double)(b - 1) + (target - below) / in_bin);
        break;
      }
      below += in_bin;
    }
    out[i] = q;
  }
  return (long long)(counts[0] + counts[nbins + 1]);
}

int fmc_accumulator_len(const fmc_context* c) {
  if (!c || c->model < 0) return acc_len(FMC_MAX_PARAM, FMC_MAX_PARAM);
  const ModelEntry& m = models()[c->model];
  return acc_len(m.nparam, m.nprop);
}

int fmc_set_accumulator_dev(fmc_context* c, double* acc_dev) {
  if (!c) return FMC_ERR_BAD_ARG;
  c->acc = acc_dev ? acc_dev : c->acc_own;
  return FMC_OK;
}

int fmc_launch(fmc_context* c, const double* params_start, long long nresample, unsigned long long seed,
               long long first_counter, const double* z_inject_dev, int keep_per_resample) {
  if (!c || !params_start || nresample < 0) return FMC_ERR_BAD_ARG;
  if (c->model < 0) return FMC_ERR_NO_DATA;
  const ModelEntry& m = models()[c->model];
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  if (c->inflight >= 0) FMC_CUDA(cudaStreamSynchronize(c->stream));
  const int P = m.nparam, Q = m.nprop;
  c->keep = keep_per_resample != 0;
  // chunking: the pipeline keeps ~0.4 KB of state per fit between its kernels
  const long long kChunk = 1LL << 21;
  const long long cap = nresample < kChunk ? (nresample > 0 ? nresample : 1) : kChunk;
  const bool binning = c->hist_bins > 0 && c->hist != nullptr;
  if (c->keep || binning) {
    // per-resample results: the whole run when the caller keeps them, else one chunk of scratch
    // for the binning kernel
    const long long rows_out = c->keep ? nresample : cap;
    const long long need_p = (long long)P * rows_out, need_q = (long long)(Q > 0 ? Q : 1) * rows_out,
                    need_c = c->keep ? 10LL * nresample : 0;
    if (need_p > c->out_cap_p) {
      cudaFree(c->params_out_dev);
      c->params_out_dev = nullptr;
      c->out_cap_p = 0;
      FMC_CUDA(cudaMalloc(&c->params_out_dev, sizeof(double) * need_p));
      c->out_cap_p = need_p;
    }
    if (need_q > c->out_cap_q) {
      cudaFree(c->props_out_dev);
      c->props_out_dev = nullptr;
      c->out_cap_q = 0;
      FMC_CUDA(cudaMalloc(&c->props_out_dev, sizeof(double) * need_q));
      c->out_cap_q = need_q;
    }
    if (need_c > c->out_cap_c) {
      cudaFree(c->counters_dev_out);
      c->counters_dev_out = nullptr;
      c->out_cap_c = 0;
      FMC_CUDA(cudaMalloc(&c->counters_dev_out, sizeof(int) * need_c));
      c->out_cap_c = need_c;
    }
  }
  const int PP = P * (P + 1) / 2;
  if (cap > c->chunk_cap || P != c->chunk_p) {
    cudaFree(c->xbuf);
    cudaFree(c->initbuf);
    cudaFree(c->istate);
    cudaFree(c->cnt);
    c->xbuf = c->initbuf = nullptr;
    c->istate = c->cnt = nullptr;
    c->chunk_cap = 0;
    FMC_CUDA(cudaMalloc(&c->xbuf, sizeof(double) * P * cap));
    FMC_CUDA(cudaMalloc(&c->initbuf, sizeof(double) * (1 + P + PP) * cap));
    FMC_CUDA(cudaMalloc(&c->istate, sizeof(int) * 2 * cap));
    FMC_CUDA(cudaMalloc(&c->cnt, sizeof(int) * 10 * cap));
    c->chunk_cap = cap;
    c->chunk_p = P;
  }
  const long long nchunks = (nresample + cap - 1) / cap;
  if (4 * nchunks > c->ncounters) {  // one work counter per persistent launch: 2 halves x 2 nl2sno calls per chunk
    cudaFree(c->counters_dev);
    c->counters_dev = nullptr;
    c->ncounters = 0;
    FMC_CUDA(cudaMalloc(&c->counters_dev, sizeof(unsigned long long) * 4 * nchunks));
    c->ncounters = (int)(4 * nchunks);
  }

  KernelArgs a;
  std::memset(&a, 0, sizeof(a));
  a.rows = c->rows_dev;
  a.aux = m.naux > 0 ? c->aux_dev : nullptr;
  a.n = c->n;
  a.n_pad = c->n_pad;
  for (int i = 0; i < P; ++i) {
    a.start[i] = params_start[i];
    c->start[i] = params_start[i];
  }
  a.seed = seed;
  a.cap = c->chunk_cap;
  a.xbuf = c->xbuf;
  a.init = c->initbuf;
  a.istate = c->istate;
  a.cnt = c->cnt;
  a.acc = c->acc;
  a.anom = c->anom_dev;
  fmc::HistArgs h;
  std::memset(&h, 0, sizeof(h));
  if (binning) {
    h.hist = c->hist;
    h.bins = c->hist_bins;
    h.np = P;
    h.nq = Q;
    for (int j = 0; j < P + Q; ++j) {
      h.lo[j] = c->hist_lo[j];
      h.scale[j] = c->hist_scale[j];
    }
  }

  // data in shared memory when it fits (227 KB opt-in on B200), else through L1/L2
  size_t shmem = (sizeof(RowData) + sizeof(double) * m.naux) * (size_t)c->n_pad;
  bool smem = shmem + 1024 <= (size_t)c->max_smem_optin;
  if (!smem) shmem = 0;
  const bool inject = z_inject_dev != nullptr;
  int bps = 0;
  const int jac = c->jac;
  if (jac == fmc::JAC_ANALYTIC && !m.analytic) return FMC_ERR_BAD_ARG;  // double-only model_y, no model_grad
  const bool per_warp = c->mapping == FMC_FIT_PER_WARP;
  // FMC_SCHEDULE_AUTO, by measurement (profiles/r2_schedule_*.json): the round schedule wins where the pass
  // kernel gains from shedding the solver's state -- rows that do not fit in shared memory, or so many
  // parameters that the row loop needs every register (C4: n = 10^4, p = 10, +19 %) -- and loses on
  // small fits (C1, C2: the solve kernel's trip through HBM costs more than the divergence it removes; C5,
  // whose fits diverge strongly: rounds won until the refill kernel's solver was compiled compactly for
  // that model, csrc/kernels_model3.cu -- 8.2e5 against 6.4e5 fits/s).
  const int schedule = c->schedule == FMC_SCHEDULE_AUTO ? ((!smem || P > 6) ? FMC_SCHEDULE_ROUNDS : FMC_SCHEDULE_REFILL)
                                                        : c->schedule;
  c->schedule_used = schedule;
  const bool rounds = !per_warp && schedule == FMC_SCHEDULE_ROUNDS;
  const fmc::RoundEntry rentry = rounds ? fmc::fmc_round_entry(c->model) : fmc::RoundEntry{0, nullptr, nullptr, nullptr};
  if (rounds && !rentry.enqueue_round) return FMC_ERR_BAD_MODEL;
  c->rs_used = false;
#ifdef FMC_WITH_WARP_MAPPING
  const fmc::WarpEntry wentry = per_warp ? fmc::fmc_warp_entry(c->model) : fmc::WarpEntry{nullptr, nullptr, 0};
#else
  const fmc::WarpEntry wentry = fmc::WarpEntry{nullptr, nullptr, 0};  // (fmc_set_fit_mapping refuses FMC_FIT_PER_WARP)
#endif
  fmc::WarpArgs wa;
  std::memset(&wa, 0, sizeof(wa));
  if (per_warp) {
    if (!wentry.launch) return FMC_ERR_BAD_MODEL;
    // columns staged in shared memory: all four when they fit, else x and 1/err, else none
    const size_t col = sizeof(double) * (size_t)c->n_w;
    const size_t fixed = wentry.solver_smem + 1024;  // the warps' solver states + static shared memory
    wa.stage = (4 + m.naux) * col + fixed <= (size_t)c->max_smem_optin
                   ? 4
                   : ((2 + m.naux) * col + fixed <= (size_t)c->max_smem_optin ? 2 : 0);
    wa.cols = c->cols_dev;
    wa.n_w = c->n_w;
    shmem = wentry.solver_smem + (wa.stage > 0 ? (wa.stage + m.naux) * col : 0);
    FMC_CUDA(wentry.occupancy(inject, jac, shmem, &bps));
  } else if (rounds) {
    bps = 1;  // (the pass kernel's own occupancy is checked in rounds_prepare)
  } else {
    FMC_CUDA(m.occupancy(inject, smem, jac, shmem, &bps));
  }
  if (bps < 1) {
    g_last_cuda_error = "kernel does not fit on an SM";
    return FMC_ERR_CUDA;
  }
  // persistent grid for the refill kernels: every SM full, threads take fits from a counter
  const long long full_grid = (long long)c->num_sms * bps;
  c->geom[1] = per_warp ? FMC_WARP_THREADS : FMC_THREADS;
  c->geom[2] = bps;
  c->geom[3] = c->num_sms;

  // first nl2sno call: shared Jacobian at params_start when its rows fit in shared memory too
  const size_t shmem_shared = (sizeof(RowData) + sizeof(double) * (1 + P)) * (size_t)c->n_pad;
  const bool shared_init =
      !per_warp && jac == fmc::JAC_FD && shmem_shared + 1024 <= (size_t)c->max_smem_optin;
  double* g0_dev = nullptr;
  if (shared_init) {
    const long long need = (long long)(1 + P) * c->n_pad + PP;
    if (need > c->rowj_cap) {
      cudaFree(c->rowj_dev);
      c->rowj_dev = nullptr;
      c->rowj_cap = 0;
      FMC_CUDA(cudaMalloc(&c->rowj_dev, sizeof(double) * need));
      c->rowj_cap = need;
    }
    g0_dev = c->rowj_dev + (size_t)(1 + P) * c->n_pad;
  }

  if (per_warp) {
    const long long need = full_grid * (FMC_WARP_THREADS / 32) * (long long)c->n_w;
    if (need > c->tscratch_cap) {
      cudaFree(c->tscratch_dev);
      c->tscratch_dev = nullptr;
      c->tscratch_cap = 0;
      FMC_CUDA(cudaMalloc(&c->tscratch_dev, sizeof(double) * need));
      c->tscratch_cap = need;
    }
    wa.tscratch = c->tscratch_dev;
  }
  if (rounds && nresample > 0) {
    const int prc = rounds_prepare(c, rentry, P, inject, smem, jac, shmem, cap, nchunks);
    if (prc != FMC_OK) return prc;
    c->rs_used = true;
  }
  if (nchunks > 0) FMC_CUDA(cudaMemsetAsync(c->counters_dev, 0, sizeof(unsigned long long) * 4 * nchunks, c->stream));
  FMC_CUDA(cudaMemsetAsync(c->acc, 0, sizeof(double) * acc_len(P, Q), c->stream));
  FMC_CUDA(cudaMemsetAsync(c->anom_dev, 0, sizeof(unsigned long long) * (1 + 2 * fmc::ANOM_CAP), c->stream));
  if (binning)
    FMC_CUDA(cudaMemsetAsync(h.hist, 0, sizeof(unsigned long long) * (size_t)(P + Q) * (h.bins + 2), c->stream));
  FMC_CUDA(cudaEventRecord(c->ev0, c->stream));
  c->nlaunch = 0;
  if (shared_init && nresample > 0) FMC_LAUNCH(m.launch_jac0(a, c->rowj_dev, g0_dev, c->stream));
  for (long long ch = 0; ch < nchunks; ++ch) {
    const long long lo = ch * cap, cnt = (nresample - lo) < cap ? (nresample - lo) : cap;
    a.first = first_counter + lo;
    a.nres = cnt;
    a.z_inject = inject ? z_inject_dev + (size_t)lo * c->n : nullptr;
    const size_t out_row = c->keep ? (size_t)lo : 0;  // scratch is reused by every chunk
    a.params_out = (c->keep || binning) ? c->params_out_dev + out_row * P : nullptr;
    a.props_out = (c->keep || binning) ? c->props_out_dev + out_row * (Q > 0 ? Q : 1) : nullptr;
    a.counters_out = c->keep ? c->counters_dev_out + (size_t)lo * 10 : nullptr;
    long long grid = full_grid;
    const long long need = (cnt + FMC_THREADS - 1) / FMC_THREADS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (per_warp) {  // one warp per fit: both nl2sno calls inside one kernel (warp_kernel.cuh)
      const long long warps_per_block = FMC_WARP_THREADS / 32;
      long long wgrid = full_grid;
      const long long wneed = (cnt + warps_per_block - 1) / warps_per_block;
      if (wgrid > wneed) wgrid = wneed;
      if (wgrid < 1) wgrid = 1;
      grid = wgrid;
    }
    if (!rounds) c->geom[0] = (int)grid;
    if (rounds) {
      // round schedule: the first evaluation of the first call for the whole chunk (uniform kernels,
      // as below), then rounds of solve / pass over the slot pool until every fit of the chunk has
      // been through both nl2sno calls, then the moment sums
      a.call = 0;
      a.next_id = c->counters_dev + 4 * ch;
      if (shared_init)
        FMC_LAUNCH(m.launch_init_shared(a, c->rowj_dev, g0_dev, inject, shmem_shared, c->stream));
      else
        FMC_LAUNCH(m.launch_init(a, inject, smem, jac, shmem, c->stream));
      c->rs_kargs_pinned[ch] = a;
      const int rrc = rounds_run_chunk(c, rentry, ch, inject, smem, jac, shmem);
      if (rrc != FMC_OK) return rrc;
      long long mgrid = (cnt + 255) / 256;
      if (mgrid > (long long)c->num_sms * 8) mgrid = (long long)c->num_sms * 8;
      FMC_LAUNCH(rentry.enqueue_moments(c->rs_kargs_dev, (int)mgrid, c->stream));
    }
    if (per_warp) {
      a.call = 0;
      a.next_id = c->counters_dev + 4 * ch;
      FMC_LAUNCH(wentry.launch(a, wa, inject, jac, (int)grid, shmem, c->stream));
    }
    // Refill schedule.  (Running the two halves of a chunk on two streams, so that the blocks of one half's
    // persistent kernel fill the tail of the other's, was measured and is SLOWER: C2 2.49e7 against 2.69e7
    // fits/s, C5 5.84e5 against 6.96e5 -- profiles/r2_refill_split_streams.json.)
    for (int call = 0; call < 2 && !per_warp && !rounds; ++call) {  // fit_data: nl2sno twice (bootstrap.f90:219-225)
      a.call = call;
      a.next_id = c->counters_dev + 4 * ch + call;
      if (call == 0 && shared_init)
        FMC_LAUNCH(m.launch_init_shared(a, c->rowj_dev, g0_dev, inject, shmem_shared, c->stream));
      else
        FMC_LAUNCH(m.launch_init(a, inject, smem, jac, shmem, c->stream));
      FMC_LAUNCH(m.launch_iter(a, inject, smem, jac, (int)grid, shmem, c->stream));
    }
    if (binning) {
      h.params = a.params_out;
      h.props = a.props_out;
      h.nres = cnt;
      const long long work = cnt * (P + Q);
      long long hgrid = (work + 255) / 256;
      const long long hmax = (long long)c->num_sms * 8;
      if (hgrid > hmax) hgrid = hmax;
      if (hgrid < 1) hgrid = 1;
      hist_kernel<<<(unsigned)hgrid, 256, 0, c->stream>>>(h);
      FMC_LAUNCH(cudaGetLastError());
    }
  }
  FMC_CUDA(cudaEventRecord(c->ev1, c->stream));
  c->inflight = nresample;
  return FMC_OK;
}

int fmc_decode_accumulator(int model, const double* params_start, const double* acc, double* sum_p,
                           double* sum_p2, double* sum_prop, double* sum_prop2, long long* nfit,
                           long long* rc_count, long long* totals) {
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  if (!params_start || !acc) return FMC_ERR_BAD_ARG;
  const ModelEntry& m = models()[model];
  const int P = m.nparam, Q = m.nprop, NM = 2 * (P + Q);
  const double N = acc[NM + ACC_NFIT];
  double cq[FMC_MAX_PARAM];
  m.props(params_start, cq);
  // sums were accumulated as sum(v - c), sum((v - c)^2) with c = value at params_start
  for (int i = 0; i < P; ++i) {
    const double c0 = params_start[i], s1 = acc[i], s2 = acc[P + i];
    if (sum_p) sum_p[i] = N * c0 + s1;
    if (sum_p2) sum_p2[i] = s2 + 2.0 * c0 * s1 + N * c0 * c0;
  }
  for (int i = 0; i < Q; ++i) {
    const double c0 = cq[i], s1 = acc[2 * P + i], s2 = acc[2 * P + Q + i];
    if (sum_prop) sum_prop[i] = N * c0 + s1;
    if (sum_prop2) sum_prop2[i] = s2 + 2.0 * c0 * s1 + N * c0 * c0;
  }
  if (nfit) *nfit = (long long)(N + 0.5);
  if (rc_count)
    for (int i = 0; i < 16; ++i) rc_count[i] = (long long)(acc[NM + ACC_RC + i] + 0.5);
  if (totals) {
    totals[0] = (long long)(acc[NM + ACC_NFCALL] + 0.5);
    totals[1] = (long long)(acc[NM + ACC_NGCALL] + 0.5);
    totals[2] = (long long)(acc[NM + ACC_NITER] + 0.5);
    totals[3] = (long long)(acc[NM + ACC_NPASS] + 0.5);
  }
  return FMC_OK;
}

int fmc_finish(fmc_context* c, double* sum_p, double* sum_p2, double* sum_prop, double* sum_prop2,
               double* params_out, double* props_out, int* counters_out, long long* rc_count,
               long long* totals) {
  if (!c) return FMC_ERR_BAD_ARG;
  if (c->inflight < 0) return FMC_ERR_STATE;
  const ModelEntry& m = models()[c->model];
  const int P = m.nparam, Q = m.nprop;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  std::vector<double> acc(acc_len(P, Q));
  FMC_CUDA(cudaMemcpyAsync(acc.data(), c->acc, sizeof(double) * acc.size(), cudaMemcpyDeviceToHost, c->stream));
  const long long nres = c->inflight;
  if (c->keep && nres > 0) {
    if (params_out)
      FMC_CUDA(cudaMemcpyAsync(params_out, c->params_out_dev, sizeof(double) * P * nres, cudaMemcpyDeviceToHost,
                               c->stream));
    if (props_out && Q > 0)
      FMC_CUDA(cudaMemcpyAsync(props_out, c->props_out_dev, sizeof(double) * Q * nres, cudaMemcpyDeviceToHost,
                               c->stream));
    if (counters_out)
      FMC_CUDA(cudaMemcpyAsync(counters_out, c->counters_dev_out, sizeof(int) * 10 * nres, cudaMemcpyDeviceToHost,
                               c->stream));
  }
  if (c->rs_used)
    FMC_CUDA(cudaMemcpyAsync(c->rs_ctl_pinned, c->rs.ctl, sizeof(unsigned int) * fmc::RC_NWORDS, cudaMemcpyDeviceToHost,
                             c->stream));
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  if (c->rs_used) {  // the kernels of the rounds (solve, control, pass [, residual-only]), counted on the device
    c->nlaunch += (c->rs.spec ? 3LL : 4LL) * c->rs_ctl_pinned[fmc::RC_ROUNDS_TOTAL];
    c->rs_used = false;
    if (c->rs_ctl_pinned[fmc::RC_ERROR] != 0u) {
      c->inflight = -1;
      g_last_cuda_error = "round schedule: fits still iterating after the cap on rounds";
      return FMC_ERR_CUDA;
    }
  }
  c->inflight = -1;
  long long nfit = 0;
  int rc = fmc_decode_accumulator(c->model, c->start, acc.data(), sum_p, sum_p2, sum_prop, sum_prop2, &nfit,
                                  rc_count, totals);
  if (rc != FMC_OK) return rc;
  if (nfit != nres) {
    g_last_cuda_error = "kernel finished " + std::to_string(nfit) + " of " + std::to_string(nres) + " fits";
    return FMC_ERR_CUDA;
  }
  return FMC_OK;
}

int fmc_wait(fmc_context* c) {
  if (!c) return FMC_ERR_BAD_ARG;
  if (c->inflight < 0) return FMC_ERR_STATE;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  if (c->rs_used)
    FMC_CUDA(cudaMemcpyAsync(c->rs_ctl_pinned, c->rs.ctl, sizeof(unsigned int) * fmc::RC_NWORDS, cudaMemcpyDeviceToHost,
                             c->stream));
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  c->inflight = -1;
  if (c->rs_used) {  // the kernels of the rounds (solve, control, pass [, residual-only]), counted on the device
    c->nlaunch += (c->rs.spec ? 3LL : 4LL) * c->rs_ctl_pinned[fmc::RC_ROUNDS_TOTAL];
    c->rs_used = false;
    if (c->rs_ctl_pinned[fmc::RC_ERROR] != 0u) {
      g_last_cuda_error = "round schedule: fits still iterating after the cap on rounds";
      return FMC_ERR_CUDA;
    }
  }
  return FMC_OK;
}

int fmc_get_accumulator(fmc_context* c, double* acc_host) {
  if (!c || !acc_host) return FMC_ERR_BAD_ARG;
  if (c->model < 0) return FMC_ERR_NO_DATA;
  const ModelEntry& m = models()[c->model];
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  FMC_CUDA(cudaMemcpyAsync(acc_host, c->acc, sizeof(double) * acc_len(m.nparam, m.nprop), cudaMemcpyDeviceToHost,
                           c->stream));
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  return FMC_OK;
}

int fmc_last_anomalies(fmc_context* c, long long* ids, int* flags, int cap) {
  if (!c) return FMC_ERR_BAD_ARG;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  unsigned long long h[1 + 2 * fmc::ANOM_CAP];
  FMC_CUDA(cudaMemcpy(h, c->anom_dev, sizeof(h), cudaMemcpyDeviceToHost));
  const int n = (int)h[0];
  for (int i = 0; i < n && i < cap && i < fmc::ANOM_CAP; ++i) {
    if (ids) ids[i] = (long long)h[1 + 2 * i];
    if (flags) flags[i] = (int)h[2 + 2 * i];
  }
  return n;
}

float fmc_last_kernel_ms(fmc_context* c) {
  if (!c) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

long long fmc_last_launch_count(const fmc_context* c) { return c ? c->nlaunch : -1; }

int fmc_last_launch_geometry(const fmc_context* c, int* out4) {
  if (!c || !out4) return FMC_ERR_BAD_ARG;
  for (int i = 0; i < 4; ++i) out4[i] = c->geom[i];
  return FMC_OK;
}

// ---------------------------------------------------------------------------- NCCL combine
int fmc_nccl_available(void) {
  if (!nccl().ok) return 0;
  int v = 0;
  if (nccl().GetVersion && nccl().GetVersion(&v) == ncclSuccess) return v;
  return 1;
}

int fmc_nccl_unique_id(char* id128) {
  if (!id128) return FMC_ERR_BAD_ARG;
  if (!nccl().ok) {
    g_last_cuda_error = "libnccl.so.2 could not be loaded";
    return FMC_ERR_NCCL;
  }
  static_assert(NCCL_UNIQUE_ID_BYTES == 128, "fmc_nccl_unique_id hands out 128 bytes");
  ncclUniqueId id;
  FMC_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return FMC_OK;
}

int fmc_nccl_init_rank(fmc_context* c, const char* id128, int nranks, int rank) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return FMC_ERR_BAD_ARG;
  if (!nccl().ok) {
    g_last_cuda_error = "libnccl.so.2 could not be loaded";
    return FMC_ERR_NCCL;
  }
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  if (c->comm && c->comm_owned) nccl().CommDestroy(c->comm);
  c->comm = nullptr;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  FMC_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank));
  c->comm_owned = true;
  c->comm_nranks = nranks;
  return FMC_OK;
}

int fmc_allreduce(fmc_context* c) {
  if (!c) return FMC_ERR_BAD_ARG;
  if (c->model < 0) return FMC_ERR_NO_DATA;
  if (!c->comm || c->comm_nranks <= 1) return FMC_OK;  // a single rank: nothing to combine
  const ModelEntry& m = models()[c->model];
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  // ONE all-reduce of the accumulator block (moment sums, return-code histogram, counters: all doubles)
  FMC_NCCL(nccl().AllReduce(c->acc, c->acc, (size_t)acc_len(m.nparam, m.nprop), ncclDouble, ncclSum, c->comm,
                            c->stream));
  if (c->hist_bins > 0 && c->hist)  // optional binned distributions: integer counts, exact in any order
    FMC_NCCL(nccl().AllReduce(c->hist, c->hist, (size_t)(m.nparam + m.nprop) * (c->hist_bins + 2), ncclUint64,
                              ncclSum, c->comm, c->stream));
  return FMC_OK;
}

int fmc_last_combine_method(void) { return g_last_combine; }

// ---------------------------------------------------------------------------- one-shot API
namespace {
std::mutex g_ctx_mutex;
std::vector<fmc_context*> g_ctx;  // one cached context per device

int cached_context(int device, fmc_context** out) {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  if ((int)g_ctx.size() <= device) g_ctx.resize(device + 1, nullptr);
  if (!g_ctx[device]) {
    int rc = fmc_create(&g_ctx[device], device);
    if (rc != FMC_OK) return rc;
  }
  *out = g_ctx[device];
  return FMC_OK;
}

// one communicator per GPU for the in-process multi-GPU run over devices 0..G-1, created once
std::vector<std::vector<ncclComm_t>> g_comms;
ncclComm_t* cached_comms(int G) {
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  if ((int)g_comms.size() <= G) g_comms.resize(G + 1);
  if (g_comms[G].empty()) {
    std::vector<ncclComm_t> comms(G, nullptr);
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = g;
    int prev = -1;
    cudaGetDevice(&prev);
    const ncclResult_t r = nccl().CommInitAll(comms.data(), G, devs.data());
    if (prev >= 0) cudaSetDevice(prev);
    if (r != ncclSuccess) {
      g_last_cuda_error = std::string("ncclCommInitAll: ") + nccl().GetErrorString(r);
      return nullptr;
    }
    g_comms[G] = comms;
  }
  return g_comms[G].data();
}
}  // namespace

namespace {
int bootstrap_run_impl(int model, int ndata, const double* x, const double* y, const double* err_y,
                       const double* params_start, long long nresample, unsigned long long seed,
                       long long first_counter, int ngpus, const double* z_inject, double* sum_p,
                       double* sum_p2, double* sum_prop, double* sum_prop2, double* params_out,
                       double* props_out, long long* rc_count, long long* totals, int nbins, const double* bin_lo,
                       const double* bin_hi, unsigned long long* counts) {
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  if (!x || !y || !err_y || !params_start || nresample < 0 || ngpus < 0) return FMC_ERR_BAD_ARG;
  const int ndev = device_count();
  if (ndev <= 0) return FMC_ERR_NO_DEVICE;
  int G = ngpus == 0 ? ndev : (ngpus < ndev ? ngpus : ndev);
  if ((long long)G > nresample) G = nresample > 0 ? (int)nresample : 1;
  const ModelEntry& m = models()[model];
  const int P = m.nparam, Q = m.nprop, AL = acc_len(P, Q);
  const bool keep = params_out || props_out;

  std::vector<int> rcs(G, FMC_OK);
  std::vector<std::string> errs(G);
  std::vector<std::vector<double>> accs(G, std::vector<double>(AL, 0.0));
  const size_t HL = nbins > 0 ? (size_t)(P + Q) * (nbins + 2) : 0;
  std::vector<std::vector<unsigned long long>> hists(G, std::vector<unsigned long long>(HL, 0ULL));
  auto shard = [&](int g) {
    // contiguous split of the resample ids
    const long long lo = nresample * g / G, hi = nresample * (g + 1) / G, cnt = hi - lo;
    fmc_context* c = nullptr;
    int rc = cached_context(g, &c);
    double* zdev = nullptr;
    do {
      if (rc != FMC_OK) break;
      if ((rc = fmc_set_data(c, model, ndata, x, y, err_y)) != FMC_OK) break;
      if (z_inject && cnt > 0) {
        if (cudaMalloc(&zdev, sizeof(double) * (size_t)cnt * ndata) != cudaSuccess) {
          rc = FMC_ERR_ALLOC;
          break;
        }
        if (cudaMemcpyAsync(zdev, z_inject + (size_t)lo * ndata, sizeof(double) * (size_t)cnt * ndata,
                            cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
          rc = FMC_ERR_CUDA;
          break;
        }
      }
      fmc_set_accumulator_dev(c, nullptr);
      c->jac = g_default_jac;
      c->mapping = g_default_mapping;
      c->schedule = g_default_schedule >= 0 ? g_default_schedule : default_schedule();
      fmc_set_histogram_dev(c, nullptr);
      if ((rc = fmc_set_histogram(c, nbins, bin_lo, bin_hi)) != FMC_OK) break;  // cached context: also switches off
      if ((rc = fmc_launch(c, params_start, cnt, seed, first_counter + lo, zdev, keep ? 1 : 0)) != FMC_OK) break;
      rc = fmc_finish(c, nullptr, nullptr, nullptr, nullptr, params_out ? params_out + (size_t)lo * P : nullptr,
                      (props_out && Q > 0) ? props_out + (size_t)lo * Q : nullptr, nullptr, nullptr, nullptr);
      if (rc != FMC_OK) break;
      if (cudaMemcpy(accs[g].data(), c->acc, sizeof(double) * AL, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = FMC_ERR_CUDA;
        break;
      }
      if (HL > 0) rc = fmc_get_histogram(c, hists[g].data());
    } while (0);
    if (zdev) cudaFree(zdev);
    rcs[g] = rc;
    errs[g] = g_last_cuda_error;
  };
  if (G == 1) {
    shard(0);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) th.emplace_back(shard, g);
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < G; ++g)
    if (rcs[g] != FMC_OK) {
      g_last_cuda_error = errs[g];
      return rcs[g];
    }
  std::vector<double> acc(AL, 0.0);
  g_last_combine = FMC_COMBINE_NONE;
  if (G > 1) {
    // Combine the per-GPU accumulator blocks: ONE ncclAllReduce over NVLink, issued for all GPUs of the
    // box as one group on their own streams (bootstrap.f90:310-311's REDUCTION(+), across GPUs); every
    // GPU then holds the total and GPU 0's copy is read back.  Without a loadable NCCL the blocks
    // (31 doubles each) are added on the host instead; fmc_last_combine_method() says which.
    std::vector<fmc_context*> ctxs(G, nullptr);
    for (int g = 0; g < G; ++g)
      if (cached_context(g, &ctxs[g]) != FMC_OK) return FMC_ERR_CUDA;
    ncclComm_t* comms = nccl().ok ? cached_comms(G) : nullptr;
    if (comms) {
      ncclResult_t r = nccl().GroupStart();
      for (int g = 0; g < G && r == ncclSuccess; ++g) {
        fmc_context* c = ctxs[g];
        r = nccl().AllReduce(c->acc, c->acc, (size_t)AL, ncclDouble, ncclSum, comms[g], c->stream);
        if (r == ncclSuccess && HL > 0)
          r = nccl().AllReduce(c->hist, c->hist, HL, ncclUint64, ncclSum, comms[g], c->stream);
      }
      const ncclResult_t r2 = nccl().GroupEnd();
      if (r == ncclSuccess) r = r2;
      if (r != ncclSuccess) {
        g_last_cuda_error = std::string("ncclAllReduce: ") + nccl().GetErrorString(r);
        return FMC_ERR_NCCL;
      }
      for (int g = 0; g < G; ++g) {
        DeviceGuard guard_(ctxs[g]->device);
        FMC_CUDA(cudaStreamSynchronize(ctxs[g]->stream));
      }
      {
        DeviceGuard guard_(ctxs[0]->device);
        FMC_CUDA(cudaMemcpy(acc.data(), ctxs[0]->acc, sizeof(double) * AL, cudaMemcpyDeviceToHost));
        if (HL > 0 && counts)
          FMC_CUDA(cudaMemcpy(counts, ctxs[0]->hist, sizeof(unsigned long long) * HL, cudaMemcpyDeviceToHost));
      }
      g_last_combine = FMC_COMBINE_NCCL;
    }
  }
  if (g_last_combine != FMC_COMBINE_NCCL) {
    for (int g = 0; g < G; ++g)
      for (int i = 0; i < AL; ++i) acc[i] += accs[g][i];
    if (HL > 0 && counts) {
      for (size_t i = 0; i < HL; ++i) counts[i] = 0ULL;
      for (int g = 0; g < G; ++g)
        for (size_t i = 0; i < HL; ++i) counts[i] += hists[g][i];
    }
    if (G > 1) g_last_combine = FMC_COMBINE_HOST;
  }
  return fmc_decode_accumulator(model, params_start, acc.data(), sum_p, sum_p2, sum_prop, sum_prop2, nullptr,
                                rc_count, totals);
}
}  // namespace

int fmc_bootstrap_run(int model, int ndata, const double* x, const double* y, const double* err_y,
                      const double* params_start, long long nresample, unsigned long long seed,
                      long long first_counter, int ngpus, const double* z_inject, double* sum_p,
                      double* sum_p2, double* sum_prop, double* sum_prop2, double* params_out,
                      double* props_out, long long* rc_count, long long* totals) {
  return bootstrap_run_impl(model, ndata, x, y, err_y, params_start, nresample, seed, first_counter, ngpus, z_inject,
                            sum_p, sum_p2, sum_prop, sum_prop2, params_out, props_out, rc_count, totals, 0, nullptr,
                            nullptr, nullptr);
}

int fmc_bootstrap_run_binned(int model, int ndata, const double* x, const double* y, const double* err_y,
                             const double* params_start, long long nresample, unsigned long long seed,
                             long long first_counter, int ngpus, double* sum_p, double* sum_p2, double* sum_prop,
                             double* sum_prop2, long long* rc_count, long long* totals, int nbins,
                             const double* bin_lo, const double* bin_hi, unsigned long long* counts) {
  if (nbins <= 0 || !bin_lo || !bin_hi || !counts) return FMC_ERR_BAD_ARG;
  return bootstrap_run_impl(model, ndata, x, y, err_y, params_start, nresample, seed, first_counter, ngpus, nullptr,
                            sum_p, sum_p2, sum_prop, sum_prop2, nullptr, nullptr, rc_count, totals, nbins, bin_lo,
                            bin_hi, counts);
}

int fmc_fit_once(int model, int ndata, const double* x, const double* y, const double* err_y, double* params,
                 int* rc2) {
  if (!params) return FMC_ERR_BAD_ARG;
  if (model < 0 || model >= fmc_model_count()) return FMC_ERR_BAD_MODEL;
  if (device_count() <= 0) return FMC_ERR_NO_DEVICE;
  fmc_context* c = nullptr;
  int rc = cached_context(0, &c);
  if (rc != FMC_OK) return rc;
  if ((rc = fmc_set_data(c, model, ndata, x, y, err_y)) != FMC_OK) return rc;
  // offset_data(.FALSE.): a single "resample" with all offsets zero
  DevBuf zbuf;
  FMC_CUDA(cudaMalloc(&zbuf.p, sizeof(double) * ndata));
  double* zdev = zbuf.p;
  FMC_CUDA(cudaMemsetAsync(zdev, 0, sizeof(double) * ndata, c->stream));
  fmc_set_accumulator_dev(c, nullptr);
  fmc_set_histogram(c, 0, nullptr, nullptr);  // the cached context may have binned before
  c->jac = g_default_jac;
  c->mapping = g_default_mapping;
  c->schedule = g_default_schedule >= 0 ? g_default_schedule : default_schedule();
  rc = fmc_launch(c, params, 1, 0ULL, 0, zdev, 1);
  const int P = models()[model].nparam;
  std::vector<double> pout(P);
  int cnt[10];
  if (rc == FMC_OK) rc = fmc_finish(c, nullptr, nullptr, nullptr, nullptr, pout.data(), nullptr, cnt, nullptr, nullptr);
  if (rc != FMC_OK) return rc;
  for (int i = 0; i < P; ++i) params[i] = pout[i];
  if (rc2) {
    rc2[0] = cnt[0];
    rc2[1] = cnt[5];
  }
  return FMC_OK;
}

int fmc_measure_dfma_peak(int device, int iters, double* tflops, float* ms_out) {
  if (!tflops || iters < 1) return FMC_ERR_BAD_ARG;
  if (device_count() <= 0) return FMC_ERR_NO_DEVICE;
  DeviceGuard guard_(device);
  FMC_CUDA(guard_.err);
  cudaDeviceProp prop;
  FMC_CUDA(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8;
  DevBuf obuf;
  FMC_CUDA(cudaMalloc(&obuf.p, sizeof(double) * threads * blocks));
  double* out = obuf.p;
  struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } ev;
  FMC_CUDA(cudaEventCreate(&ev.a));
  FMC_CUDA(cudaEventCreate(&ev.b));
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  dfma_peak_kernel<<<blocks, threads>>>(out, iters / 8 + 1, 0.999999, 1e-6);  // warm-up
  FMC_CUDA(cudaEventRecord(e0));
  dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
  FMC_CUDA(cudaEventRecord(e1));
  FMC_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  FMC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = 2.0 * 8.0 * 16.0 * (double)iters * (double)threads * (double)blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  return FMC_OK;
}

int fmc_debug_philox4x32(const unsigned int* counters_keys, int n, unsigned int* out) {
  if (!counters_keys || !out || n < 0) return FMC_ERR_BAD_ARG;
  if (device_count() <= 0) return FMC_ERR_NO_DEVICE;
  if (n == 0) return FMC_OK;
  fmc_context* c = nullptr;
  int rc = cached_context(0, &c);
  if (rc != FMC_OK) return rc;
  DeviceGuard guard_(c->device);
  FMC_CUDA(guard_.err);
  DevBuf in, outb;  // (double* buffers used as raw storage)
  FMC_CUDA(cudaMalloc(&in.p, sizeof(uint32_t) * 6 * (size_t)n + 8));
  FMC_CUDA(cudaMalloc(&outb.p, sizeof(uint32_t) * 4 * (size_t)n + 8));
  FMC_CUDA(cudaMemcpyAsync(in.p, counters_keys, sizeof(uint32_t) * 6 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  philox_debug_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(in.p), n,
                                                               reinterpret_cast<uint32_t*>(outb.p));
  FMC_CUDA(cudaGetLastError());
  FMC_CUDA(cudaMemcpyAsync(out, outb.p, sizeof(uint32_t) * 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  FMC_CUDA(cudaStreamSynchronize(c->stream));
  return FMC_OK;
}

int fmc_generate_offsets(unsigned long long seed, long long first_counter, long long nresample, int ndata,
                         double* z_out) {
  if (!z_out || nresample < 0 || ndata < 1) return FMC_ERR_BAD_ARG;
  if (device_count() <= 0) return FMC_ERR_NO_DEVICE;
  if (nresample == 0) return FMC_OK;
  fmc_context* c = nullptr;
  int rc = cached_context(0, &c);
  if (rc != FMC_OK) return rc;
  DevBuf zbuf;
  FMC_CUDA(cudaMalloc(&zbuf.p, sizeof(double) * (size_t)nresample * ndata));
  double* zdev = zbuf.p;
  generate_offsets_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(seed, first_counter, nresample, ndata, zdev);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(z_out, zdev, sizeof(double) * (size_t)nresample * ndata, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    g_last_cuda_error = std::string("generate_offsets: ") + cudaGetErrorString(e);
    return FMC_ERR_CUDA;
  }
  return FMC_OK;
}

}  // extern "C"

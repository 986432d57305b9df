// common.cuh — shared plumbing of the CUDA library (context, error handling, device buffers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lvi_exc_b200.h"

namespace lvi {

void set_error(const std::string& msg);

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define LVI_CUDA(call)                                                                                        \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess) { cudaGetLastError(); /* clear the sticky last-error so later launches do not report it */         \
      throw ::lvi::Error(LVI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); } \
  } while (0)

#define LVI_REQUIRE(cond, code, msg)                 \
  do {                                               \
    if (!(cond)) throw ::lvi::Error((code), (msg));  \
  } while (0)

// catch-all wrapper for extern "C" entry points
template <class F>
int guarded(F&& f) {
  try {
    f();
    return LVI_OK;
  } catch (const Error& e) {
    set_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_error(e.what());
    return LVI_ERR_INVALID;
  }
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

}  // namespace lvi

struct lvi_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  void* nccl = nullptr;  // ncclComm_t
  bool owns_nccl = false;
  int rank = 0, world = 1;
  int64_t launches = 0;
  int sm_count = lvi::kNumSMs;
  // two auxiliary streams + fork/join events: independent kernels of one phase (the per-type normal-equation kernels) run side by side
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  // Per-context kernel state.  Function attributes (opt-in dynamic shared memory), occupancy figures and __device__ symbols belong to
  // ONE device: a process that drives two GPUs has two contexts, and each sets them up for its own device on first use.
  struct KernelState {
    bool pair_table = false;          // problem.cu: g_pair_table uploaded to this device
    bool lin_attr[8] = {};            // problem.cu: linearize_kernel<TYPE> shared-memory attribute set
    bool jac_attr[8] = {};            //             jacobian_kernel<TYPE> likewise
    bool gather_attr[2] = {};         // assemble.cu: gather_kernel<8 / 16-bit descriptors> dynamic shared memory attribute set
    int fac_resident = 0;             // solver.cu: co-resident CTAs of band_factor_ll_kernel (0 = not yet queried)
    int bs_resident = 0;              //            ... of band_backsolve_ll_kernel
    size_t corner_attr = 48 * 1024;   //            dynamic shared memory granted to corner_solve_kernel so far
    size_t schur_attr = 48 * 1024;    //            ... to schur_eliminate_kernel
    size_t assoc_attr = 48 * 1024;    // assoc.cu: ... to assoc_scan_fused_kernel
    bool selftest_done = false;
  } ks;
  // Peer-memory region for the fused normal-equation reduction (p2p.cu): one cudaMalloc'd block per rank, exported with CUDA IPC and mapped
  // by every other rank of the node, so the reduction kernel reads its peers' tiles directly over NVLink.  state: 0 not tried, 1 mapped,
  // -1 unavailable on this node (the NCCL all-reduce is used instead).
  struct PeerRegion {
    void* base = nullptr;
    size_t bytes = 0;
    std::vector<void*> peer;   // [world] base pointers (own = base)
    void** peer_d = nullptr;   // device copy of the table
    unsigned epoch = 0;
    int state = 0;
  } p2p;
  // host-side resources every solve needs, made once per context (cudaMallocHost / cudaEventCreate cost 0.1 - 1 ms each on a cold driver)
  double* h_scal = nullptr;           // pinned mirror of a problem's scalar block (64 doubles)
  std::vector<cudaEvent_t> timing_events;   // reused by the LM loop's phase timers
  size_t timing_used = 0;
  // optional per-kernel timing (lvi_ctx_kernel_timing): every LVI_LAUNCH is bracketed by CUDA events on the stream it goes to
  bool kt_enabled = false;
  struct KtSpan { const char* name; cudaEvent_t a, b; };
  std::vector<KtSpan> kt_spans;
  std::vector<cudaEvent_t> kt_free;
  cudaEvent_t kt_event() {
    if (!kt_free.empty()) { cudaEvent_t e = kt_free.back(); kt_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  cudaEvent_t timing_event() {
    if (timing_used == timing_events.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return nullptr; timing_events.push_back(e); }
    return timing_events[timing_used++];
  }
};

namespace lvi {

void p2p_ctx_release(lvi_ctx* ctx);   // p2p.cu: unmap / free the peer-memory region

// The stream the calling thread's current entry point works on.  Device buffers are taken from and returned to the device's
// stream-ordered memory pool on it (cudaMallocAsync / cudaFreeAsync; the pool keeps freed blocks, release threshold = never), so
// creating and destroying a problem or a map costs microseconds per buffer instead of a device-synchronising cudaMalloc / cudaFree.
inline thread_local cudaStream_t tl_stream = nullptr;
inline void activate(lvi_ctx* c) {
  LVI_CUDA(cudaSetDevice(c->device));
  tl_stream = c->stream;
}

// RAII device buffer on a context's stream-ordered pool
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t pool_stream = nullptr;   // stream the block was taken on (nullptr: plain cudaMalloc)
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), pool_stream(o.pool_stream) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; pool_stream = o.pool_stream; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (!count) return;
    pool_stream = tl_stream;
    if (pool_stream) LVI_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), pool_stream));
    else LVI_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
  }
  void release() {
    if (p) {  // a destroyed context's stream is no longer valid: fall back to the synchronising free
      if (!pool_stream || cudaFreeAsync(p, pool_stream) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
    }
    p = nullptr; n = 0;
  }
  void zero(cudaStream_t s) { if (n) LVI_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) { if (count) LVI_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void download(T* h, size_t count, cudaStream_t s) const { if (count) LVI_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
};

inline int grid_for(int64_t work_items, int block, int sm_count, int blocks_per_sm = 8) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(sm_count) * blocks_per_sm;  // grid sized in multiples of the SM count
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

#define LVI_LAUNCH_AS(ctx, label, kernel, grid, block, smem, ...)            \
  do {                                                                       \
    cudaEvent_t kt_a__ = nullptr, kt_b__ = nullptr;                          \
    if ((ctx)->kt_enabled) { kt_a__ = (ctx)->kt_event(); kt_b__ = (ctx)->kt_event(); cudaEventRecord(kt_a__, (ctx)->stream); } \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);         \
    if (kt_a__) { cudaEventRecord(kt_b__, (ctx)->stream); (ctx)->kt_spans.push_back({label, kt_a__, kt_b__}); } \
    ++(ctx)->launches;                                                       \
    LVI_CUDA(cudaGetLastError());                                            \
  } while (0)
#define LVI_LAUNCH(ctx, kernel, grid, block, smem, ...) LVI_LAUNCH_AS(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)

// the same bracket around a library call that launches kernels itself (CUB sort / scan)
#define LVI_TIMED(ctx, label, call)                                          \
  do {                                                                       \
    cudaEvent_t kt_a__ = nullptr, kt_b__ = nullptr;                          \
    if ((ctx)->kt_enabled) { kt_a__ = (ctx)->kt_event(); kt_b__ = (ctx)->kt_event(); cudaEventRecord(kt_a__, (ctx)->stream); } \
    LVI_CUDA(call);                                                          \
    if (kt_a__) { cudaEventRecord(kt_b__, (ctx)->stream); (ctx)->kt_spans.push_back({label, kt_a__, kt_b__}); } \
  } while (0)

}  // namespace lvi

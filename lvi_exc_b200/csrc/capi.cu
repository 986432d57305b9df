// capi.cu — context management and misc entry points of the C-ABI (include/lvi_exc_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "common.cuh"
#include "nccl_dyn.hpp"

namespace lvi {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace lvi

using namespace lvi;

extern "C" {

const char* lvi_last_error(void) { return g_error.c_str(); }
int lvi_abi_version(void) { return LVI_ABI_VERSION; }
/* sizeof of the ABI structures, for binding generators: 0 problem_desc, 1 solve_options, 2 solve_summary, 3 point_xyzit, 4 surfel_point */
int64_t lvi_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(lvi_problem_desc);
    case 1: return sizeof(lvi_solve_options);
    case 2: return sizeof(lvi_solve_summary);
    case 3: return sizeof(lvi_point_xyzit);
    case 4: return sizeof(lvi_surfel_point);
    default: return -1;
  }
}
int lvi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static void ctx_init(lvi_ctx* c, int device) {
  LVI_REQUIRE(lvi_device_count() > 0, LVI_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
  LVI_REQUIRE(device >= 0 && device < lvi_device_count(), LVI_ERR_INVALID, "bad device ordinal");
  LVI_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LVI_CUDA(cudaGetDeviceProperties(&prop, device));
  LVI_REQUIRE(prop.major >= 10, LVI_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) + ", library is built for sm_100a only");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  LVI_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    LVI_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    LVI_CUDA(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
  }
  LVI_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  // device buffers come from the stream-ordered pool (common.cuh, DBuf): keep freed blocks cached instead of returning them to the driver
  cudaMemPool_t pool;
  LVI_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;
  LVI_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  tl_stream = c->stream;
  // Grow the pool ONCE, here, instead of inside the first solve: a C2 problem takes ~1.1 GB of solver buffers (tile stores + flagged
  // copies) and the first cudaMallocAsync of that size is a synchronous driver allocation of tens of milliseconds.  LVI_POOL_RESERVE_MB
  // overrides the default (4 GB of the 180 GB: two C2 problems side by side); 0 disables it.
  {
    size_t mb = 4096;
    if (const char* e = std::getenv("LVI_POOL_RESERVE_MB")) mb = static_cast<size_t>(std::strtoull(e, nullptr, 10));
    if (mb > 0) {
      void* p = nullptr;
      if (cudaMallocAsync(&p, mb << 20, c->stream) == cudaSuccess) cudaFreeAsync(p, c->stream);
      else cudaGetLastError();   // a smaller device: the pool simply grows on demand
    }
  }
  LVI_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_scal), 64 * sizeof(double)));
  for (int i = 0; i < 16; ++i) { cudaEvent_t e; LVI_CUDA(cudaEventCreate(&e)); c->timing_events.push_back(e); }
  LVI_CUDA(cudaStreamSynchronize(c->stream));
}

int lvi_ctx_create(int device, void* nccl_comm, int rank, int world, lvi_ctx** out) {
  return guarded([&] {
    LVI_REQUIRE(out, LVI_ERR_INVALID, "lvi_ctx_create: null out");
    LVI_REQUIRE(world >= 1 && rank >= 0 && rank < world, LVI_ERR_INVALID, "lvi_ctx_create: bad rank/world");
    LVI_REQUIRE(world == 1 || nccl_comm, LVI_ERR_INVALID, "lvi_ctx_create: world > 1 needs an NCCL communicator");
    auto c = new lvi_ctx();
    try { ctx_init(c, device); } catch (...) { delete c; throw; }
    c->nccl = nccl_comm; c->rank = rank; c->world = world;
    *out = c;
  });
}

int lvi_nccl_unique_id(void* id128) {
  return guarded([&] {
    LVI_REQUIRE(id128, LVI_ERR_INVALID, "null id");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    LVI_REQUIRE(nccl().GetUniqueId(&id) == ncclSuccess, LVI_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(id128, &id, 128);
  });
}

int lvi_ctx_create_nccl(int device, const void* id128, int rank, int world, lvi_ctx** out) {
  return guarded([&] {
    LVI_REQUIRE(out && id128, LVI_ERR_INVALID, "lvi_ctx_create_nccl: null argument");
    LVI_REQUIRE(world >= 1 && rank >= 0 && rank < world, LVI_ERR_INVALID, "lvi_ctx_create_nccl: bad rank/world");
    auto c = new lvi_ctx();
    try {
      ctx_init(c, device);
      ncclUniqueId id;
      std::memcpy(&id, id128, 128);
      ncclComm_t comm;
      ncclResult_t r = nccl().CommInitRank(&comm, world, id, rank);
      LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
      c->nccl = comm; c->owns_nccl = true; c->rank = rank; c->world = world;
    } catch (...) { delete c; throw; }
    *out = c;
  });
}

int lvi_ctx_destroy(lvi_ctx* ctx) {
  if (!ctx) return LVI_OK;
  cudaSetDevice(ctx->device);
  p2p_ctx_release(ctx);
  if (ctx->owns_nccl && ctx->nccl) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->nccl));
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);   // hand cached blocks back to the driver
    for (int i = 0; i < 2; ++i) { if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    for (auto& sp : ctx->kt_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : ctx->kt_free) cudaEventDestroy(e);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    cudaStreamDestroy(ctx->stream);
    if (tl_stream == ctx->stream) tl_stream = nullptr;
  }
  (void)cudaGetLastError();
  delete ctx;
  return LVI_OK;
}

int lvi_ctx_synchronize(lvi_ctx* ctx) {
  return guarded([&] {
    LVI_REQUIRE(ctx, LVI_ERR_INVALID, "null ctx");
    activate(ctx);
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
// per-kernel timing: enable (1) / disable (0); lvi_ctx_kernel_times drains the spans recorded so far as text lines "name count total_ms"
int lvi_ctx_kernel_timing(lvi_ctx* ctx, int enable) {
  if (!ctx) return LVI_ERR_INVALID;
  ctx->kt_enabled = enable != 0;
  return LVI_OK;
}
int64_t lvi_ctx_kernel_times(lvi_ctx* ctx, char* out, int64_t cap) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 2; ++i) cudaStreamSynchronize(ctx->aux[i]);
  std::vector<std::pair<std::string, std::pair<int, double>>> agg;
  for (auto& sp : ctx->kt_spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    std::string name = sp.name;
    const size_t lt = name.find('<');   // template arguments stay: linearize_kernel<RT_CAM> etc.
    (void)lt;
    bool found = false;
    for (auto& a : agg) if (a.first == name) { a.second.first += 1; a.second.second += ms; found = true; break; }
    if (!found) agg.push_back({name, {1, ms}});
    ctx->kt_free.push_back(sp.a); ctx->kt_free.push_back(sp.b);
  }
  ctx->kt_spans.clear();
  std::string txt;
  for (auto& a : agg) {
    std::string nm = a.first;
    for (char& c : nm) if (c == ' ') c = '_';
    txt += nm + " " + std::to_string(a.second.first) + " " + std::to_string(a.second.second) + "\n";
  }
  if (out && cap > 0) { const size_t n = std::min<size_t>(txt.size(), static_cast<size_t>(cap - 1)); std::memcpy(out, txt.data(), n); out[n] = 0; }
  return static_cast<int64_t>(txt.size());
}

void* lvi_ctx_stream(lvi_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
int64_t lvi_ctx_launch_count(lvi_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

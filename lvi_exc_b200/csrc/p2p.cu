// p2p.cu — the multi-GPU reduction of the normal equations as ONE kernel over NVLink peer memory (SURVEY §8e collective 3).
//
// Every rank linearises its time chunk of the residual tables; the reference path for the sum is ncclAllReduce over the packed tile store
// (solver.cu).  Here each rank's linearisation writes into a cudaMalloc'd region that is exported with CUDA IPC and mapped by all peers of
// the node, and p2p_reduce_kernel forms  H = sum_q H_q  by READING the peers' units (1024 doubles: one 32x32 tile, or one block of the
// small dense part {corner, g, Schur rows, Schur diagonal, cost}) straight out of their HBM -- and only the units a peer actually wrote:
// the tile units carry a static flag byte from the assembly plan (a rank's time chunk touches ~1/N of the tiles), the small units a flag
// computed after the gather.  The sum runs in rank order on every rank, so all ranks hold bitwise identical normal equations and walk the
// same LM path; the result lands directly in the private tile store the factorisation reads (no pack / unpack pass, no second copy).
// Synchronisation is two flag words per (rank, peer) in the region: "arrived" (my units of this epoch are complete) and "done" (I have
// finished reading yours) -- system-scope stores after a fence, polled with a bounded spin that raises the solve's failure flag instead
// of hanging.  If IPC mapping is not possible on a node the library keeps using the NCCL all-reduce.
#include <cstring>

#include "nccl_dyn.hpp"
#include "problem.cuh"

namespace lvi {

constexpr int kUnit = kTileElems;               // doubles per unit
constexpr size_t kFlagBytes = 256;              // arrive[16] u32 | done[16] u32 | pad
constexpr long long kSpinBudget = 4000000000ll; // cycles (~2 s) before a wait gives up

__device__ __forceinline__ unsigned ld_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// lane q tells peer q that this rank reached `epoch` in slot 0 (arrived) or 1 (done)
__global__ void p2p_signal_kernel(void* const* __restrict__ peer, int world, int rank, int slot, unsigned epoch) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  st_sys(reinterpret_cast<unsigned*>(static_cast<char*>(peer[q]) + slot * 64) + rank, epoch);
}
__device__ __forceinline__ bool wait_flags(const unsigned* flags, int world, unsigned epoch) {   // one warp; true = everybody got there
  const int q = threadIdx.x & 31;
  bool ok = true;
  if (q < world) {
    const long long t0 = clock64();
    while (static_cast<int>(ld_sys(flags + q) - epoch) < 0) {
      if (clock64() - t0 > kSpinBudget) { ok = false; break; }
      __nanosleep(200);
    }
  }
  return __all_sync(0xffffffffu, ok);
}
__global__ void p2p_wait_kernel(const void* base, int world, int slot, unsigned epoch, int* __restrict__ fail) {
  if (!wait_flags(reinterpret_cast<const unsigned*>(static_cast<const char*>(base) + slot * 64), world, epoch) && threadIdx.x == 0) *fail = 5;
}

// flagged tile units are cleared (split tiles accumulate with atomics), the small part is cleared whole, the static tile flags are re-installed
__global__ void __launch_bounds__(256) p2p_clear_kernel(unsigned char* __restrict__ flags, const int* __restrict__ own_list, int n_tile_units, double* __restrict__ data) {
  const int u = own_list[blockIdx.x];   // the tile units this rank writes, then every small unit
  if (threadIdx.x == 0) flags[u] = u < n_tile_units ? 1 : 0;
  double2* d = reinterpret_cast<double2*>(data + static_cast<size_t>(u) * kUnit);
  for (int e = threadIdx.x; e < kUnit / 2; e += 256) d[e] = make_double2(0.0, 0.0);
}
// a small unit is flagged when it holds anything but zeros
__global__ void __launch_bounds__(256) p2p_flag_small_kernel(unsigned char* __restrict__ flags, const double* __restrict__ data, int n_tile_units) {
  const int u = n_tile_units + blockIdx.x;
  const double* d = data + static_cast<size_t>(u) * kUnit;
  int nz = 0;
  for (int e = threadIdx.x; e < kUnit; e += 256) nz |= d[e] != 0.0;
  if (__syncthreads_or(nz) && threadIdx.x == 0) flags[u] = 1;
}

struct P2PView {
  void* const* peer;         // [world] region bases
  int world, rank;
  size_t off_flags, off_data;
  int n_tile_units, n_units;
  const int* unit_list;      // units this kernel visits (tiles some rank writes + every small unit)
  int n_list;
  const unsigned char* tile_flags_all;   // [world][n_tile_units] every rank's static tile flags (local copy, gathered once per problem)
  double* tiles_out;         // private tile store
  double* small_out;         // private contiguous copy of the small part
  unsigned epoch;
  int* fail;
  unsigned* done_counter;
};

__global__ void __launch_bounds__(256) p2p_reduce_kernel(P2PView V) {
  __shared__ int s_flag[16];
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  if (tid < 32) {   // everybody's units of this epoch are complete
    const bool ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<const char*>(V.peer[V.rank])), V.world, V.epoch);
    if (tid == 0) { s_ok = ok; if (!ok) *V.fail = 5; }
  }
  __syncthreads();
  if (s_ok) {
    for (int it = blockIdx.x; it < V.n_list; it += gridDim.x) {
      const int u = V.unit_list[it];
      if (tid < V.world)   // tile units: the local copy of everybody's static flags; small units: the flag the peer computed after its gather
        s_flag[tid] = u < V.n_tile_units ? V.tile_flags_all[static_cast<size_t>(tid) * V.n_tile_units + u]
                                         : static_cast<const unsigned char*>(V.peer[tid])[V.off_flags + u];
      __syncthreads();
      double2 a0 = make_double2(0.0, 0.0), a1 = a0;
      for (int q = 0; q < V.world; ++q) {   // rank order on every rank: bitwise identical sums
        if (!s_flag[q]) continue;
        const double2* src = reinterpret_cast<const double2*>(static_cast<const char*>(V.peer[q]) + V.off_data) + static_cast<size_t>(u) * (kUnit / 2);
        const double2 v0 = src[tid], v1 = src[tid + 256];
        a0.x += v0.x; a0.y += v0.y; a1.x += v1.x; a1.y += v1.y;
      }
      double2* dst = reinterpret_cast<double2*>(u < V.n_tile_units ? V.tiles_out + static_cast<size_t>(u) * kUnit
                                                                     : V.small_out + static_cast<size_t>(u - V.n_tile_units) * kUnit);
      dst[tid] = a0; dst[tid + 256] = a1;
      __syncthreads();
    }
  }
  // the last CTA tells the peers that this rank no longer reads their regions
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned n = atomicAdd(V.done_counter, 1u) + 1u;
    s_ok = n == gridDim.x;
    if (s_ok) *V.done_counter = 0;
  }
  __syncthreads();
  if (s_ok && tid < V.world) {
    __threadfence_system();
    st_sys(reinterpret_cast<unsigned*>(static_cast<char*>(V.peer[tid]) + 64) + V.rank, V.epoch);
  }
}

static bool all_ranks_agree(lvi_ctx* ctx, bool mine) {   // logical AND across the ranks
  DBuf<int> v(1);
  const int h = mine ? 1 : 0;
  LVI_CUDA(cudaMemcpyAsync(v.p, &h, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = nccl().AllReduce(v.p, v.p, 1, ncclInt, ncclMin, static_cast<ncclComm_t>(ctx->nccl), ctx->stream);
  LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
  int out = 0;
  LVI_CUDA(cudaMemcpyAsync(&out, v.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  return out == 1;
}

static void p2p_release(lvi_ctx* ctx) {
  auto& R = ctx->p2p;
  for (int q = 0; q < static_cast<int>(R.peer.size()); ++q)
    if (q != ctx->rank && R.peer[q]) cudaIpcCloseMemHandle(R.peer[q]);
  R.peer.clear();
  (void)cudaGetLastError();
}

// collective: every rank calls it with the same size.  Grows the region when a problem needs more than it holds.
static bool p2p_ensure_region(lvi_ctx* ctx, size_t bytes) {
  auto& R = ctx->p2p;
  if (R.state < 0) return false;
  if (R.state == 1 && bytes <= R.bytes) return true;
  cudaStream_t st = ctx->stream;
  LVI_CUDA(cudaStreamSynchronize(st));
  if (R.base) {   // nobody may still be reading the old block: close the mappings everywhere, then free
    p2p_release(ctx);
    all_ranks_agree(ctx, true);
    cudaFree(R.base); R.base = nullptr; R.bytes = 0;
    if (R.peer_d) { cudaFree(R.peer_d); R.peer_d = nullptr; }
  }
  const size_t want = bytes + bytes / 4;
  bool ok = cudaMalloc(&R.base, want) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaMemset(R.base, 0, want) == cudaSuccess && cudaIpcGetMemHandle(&mine, R.base) == cudaSuccess;
  (void)cudaGetLastError();
  const int W = ctx->world;
  std::vector<cudaIpcMemHandle_t> all(W);
  {
    DBuf<unsigned char> send(sizeof(mine)), recv(sizeof(mine) * W);
    LVI_CUDA(cudaMemcpyAsync(send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    ncclResult_t r = nccl().AllGather(send.p, recv.p, sizeof(mine), ncclChar, static_cast<ncclComm_t>(ctx->nccl), st);
    LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
    LVI_CUDA(cudaMemcpyAsync(all.data(), recv.p, sizeof(mine) * W, cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaStreamSynchronize(st));
  }
  ok = all_ranks_agree(ctx, ok);
  R.peer.assign(W, nullptr);
  if (ok) {
    R.peer[ctx->rank] = R.base;
    for (int q = 0; q < W && ok; ++q) {
      if (q == ctx->rank) continue;
      ok = cudaIpcOpenMemHandle(&R.peer[q], all[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
    (void)cudaGetLastError();
  }
  ok = all_ranks_agree(ctx, ok);
  if (ok) ok = cudaMalloc(reinterpret_cast<void**>(&R.peer_d), sizeof(void*) * W) == cudaSuccess &&
               cudaMemcpy(R.peer_d, R.peer.data(), sizeof(void*) * W, cudaMemcpyHostToDevice) == cudaSuccess;
  ok = all_ranks_agree(ctx, ok);
  if (!ok) {
    p2p_release(ctx);
    if (R.base) { cudaFree(R.base); R.base = nullptr; }
    R.bytes = 0; R.state = -1;
    (void)cudaGetLastError();
    if (std::getenv("LVI_VERBOSE")) std::fprintf(stderr, "[lvi] rank %d: peer-memory mapping unavailable, using ncclAllReduce\n", ctx->rank);
    return false;
  }
  R.bytes = want; R.state = 1; R.epoch = 0;
  return true;
}

// Sets the problem up for the peer-memory reduction (collective).  On success the linearisation targets (H_lin, schur_lin, g_lin,
// cost_lin) point into this rank's region; otherwise they keep pointing at the private buffers and solver.cu all-reduces with NCCL.
bool p2p_prepare(lvi_problem* p) {
  lvi_ctx* ctx = p->ctx;
  if (ctx->world <= 1 || ctx->world > 16 || std::getenv("LVI_NO_P2P")) return false;
  P2PPlan& Q = p->p2p;
  const BandSys& H = p->H;
  Q.n_tile_units = H.NT * H.TPC;
  Q.small_elems = p->H_C.n + p->g.n + p->Hrx.n + p->Hrr.n + 1;
  const int n_small = static_cast<int>((Q.small_elems + kUnit - 1) / kUnit);
  Q.n_units = Q.n_tile_units + n_small;
  Q.off_flags = kFlagBytes;
  Q.off_data = kFlagBytes + ((static_cast<size_t>(Q.n_units) + 255) & ~static_cast<size_t>(255));
  const size_t bytes = Q.off_data + static_cast<size_t>(Q.n_units) * kUnit * sizeof(double);
  if (!p2p_ensure_region(ctx, bytes)) return false;
  cudaStream_t st = ctx->stream;
  // static tile flags of this rank from its assembly plan, gathered from every rank: the reduction then knows without a remote read which
  // peers hold something for a tile, and visits only tiles that some rank writes
  const int W = ctx->world;
  Q.tile_flags.alloc(std::max(Q.n_tile_units, 1));
  Q.tile_flags.zero(st);
  assemble_mark_tiles(p, Q.tile_flags.p);
  Q.tile_flags_all.alloc(static_cast<size_t>(W) * std::max(Q.n_tile_units, 1));
  {
    ncclResult_t r = nccl().AllGather(Q.tile_flags.p, Q.tile_flags_all.p, std::max(Q.n_tile_units, 1), ncclChar, static_cast<ncclComm_t>(ctx->nccl), st);
    LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
  }
  std::vector<unsigned char> fl(Q.tile_flags_all.n);
  Q.tile_flags_all.download(fl.data(), fl.size(), st);
  LVI_CUDA(cudaStreamSynchronize(st));
  std::vector<int> list, own;
  for (int u = 0; u < Q.n_tile_units; ++u) {
    bool any = false;
    for (int q = 0; q < W; ++q) any = any || fl[static_cast<size_t>(q) * Q.n_tile_units + u];
    if (any) list.push_back(u);
    if (fl[static_cast<size_t>(ctx->rank) * Q.n_tile_units + u]) own.push_back(u);
  }
  for (int k = 0; k < n_small; ++k) { list.push_back(Q.n_tile_units + k); own.push_back(Q.n_tile_units + k); }
  Q.unit_list.alloc(std::max<size_t>(list.size(), 1)); Q.unit_list.upload(list.data(), list.size(), st);
  Q.own_list.alloc(std::max<size_t>(own.size(), 1)); Q.own_list.upload(own.data(), own.size(), st);
  Q.n_list = static_cast<int>(list.size()); Q.n_own = static_cast<int>(own.size());
  Q.small_sum.alloc(static_cast<size_t>(n_small) * kUnit);
  Q.counter.alloc(1); Q.counter.zero(st);
  LVI_CUDA(cudaStreamSynchronize(st));
  Q.active = true;
  return true;
}

// pointers into this rank's region for the linearisation of problem p (the region may have been re-allocated since the last solve)
void p2p_bind(lvi_problem* p) {
  P2PPlan& Q = p->p2p;
  if (!Q.active) return;
  char* base = static_cast<char*>(p->ctx->p2p.base);
  double* data = reinterpret_cast<double*>(base + Q.off_data);
  double* small = data + static_cast<size_t>(Q.n_tile_units) * kUnit;
  p->H_lin = p->H; p->H_lin.tiles = data; p->H_lin.C = small;
  p->g_lin = small + p->H_C.n;
  p->schur_lin = p->schur; p->schur_lin.Hrx = p->g_lin + p->g.n; p->schur_lin.Hrr = p->schur_lin.Hrx + p->Hrx.n;
  p->cost_lin = p->schur_lin.Hrr + p->Hrr.n;
}

void p2p_begin_linearize(lvi_problem* p) {
  lvi_ctx* ctx = p->ctx;
  P2PPlan& Q = p->p2p;
  auto& R = ctx->p2p;
  p2p_bind(p);
  char* base = static_cast<char*>(R.base);
  if (R.epoch > 0) LVI_LAUNCH(ctx, p2p_wait_kernel, 1, 32, 0, R.base, ctx->world, 1, R.epoch, p->fail.p);   // the peers are done with the previous epoch
  LVI_LAUNCH(ctx, p2p_clear_kernel, Q.n_own, 256, 0, reinterpret_cast<unsigned char*>(base + Q.off_flags), Q.own_list.p, Q.n_tile_units,
             reinterpret_cast<double*>(base + Q.off_data));
}

void p2p_reduce(lvi_problem* p) {
  lvi_ctx* ctx = p->ctx;
  P2PPlan& Q = p->p2p;
  auto& R = ctx->p2p;
  cudaStream_t st = ctx->stream;
  char* base = static_cast<char*>(R.base);
  const int n_small = Q.n_units - Q.n_tile_units;
  LVI_LAUNCH(ctx, p2p_flag_small_kernel, n_small, 256, 0, reinterpret_cast<unsigned char*>(base + Q.off_flags), reinterpret_cast<const double*>(base + Q.off_data),
             Q.n_tile_units);
  const unsigned epoch = ++R.epoch;
  LVI_LAUNCH(ctx, p2p_signal_kernel, 1, 32, 0, R.peer_d, ctx->world, ctx->rank, 0, epoch);
  P2PView V{R.peer_d, ctx->world, ctx->rank, Q.off_flags, Q.off_data, Q.n_tile_units, Q.n_units, Q.unit_list.p, Q.n_list, Q.tile_flags_all.p, p->H_tiles.p, Q.small_sum.p,
            epoch, p->fail.p, reinterpret_cast<unsigned*>(Q.counter.p)};
  LVI_LAUNCH(ctx, p2p_reduce_kernel, std::min(Q.n_list, ctx->sm_count * 8), 256, 0, V);
  struct Seg { double* ptr; size_t n; };
  const Seg segs[5] = {{p->H_C.p, p->H_C.n}, {p->g.p, p->g.n}, {p->Hrx.p, p->Hrx.n}, {p->Hrr.p, p->Hrr.n}, {p->scal.p, 1}};
  size_t off = 0;
  for (const Seg& s : segs) { if (s.n) LVI_CUDA(cudaMemcpyAsync(s.ptr, Q.small_sum.p + off, s.n * sizeof(double), cudaMemcpyDeviceToDevice, st)); off += s.n; }
}

void p2p_ctx_release(lvi_ctx* ctx) {
  auto& R = ctx->p2p;
  p2p_release(ctx);
  if (R.base) cudaFree(R.base);
  if (R.peer_d) cudaFree(R.peer_d);
  R.base = nullptr; R.peer_d = nullptr; R.bytes = 0;
  (void)cudaGetLastError();
}

}  // namespace lvi

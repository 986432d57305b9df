// problem.cu — residual / Jacobian evaluation and normal-equation assembly on sm_100a (SURVEY §8 a-4 … a-11, a-13).
//
// Replaces, for one kontiki::TrajectoryEstimator problem (K/trajectory_estimator.h:19-135):
//   * ceres::DynamicAutoDiffCostFunction evaluation of the Kontiki Residual functors (4..26 Jet passes per residual per
//     iteration, each re-evaluating both spline segments with heap allocations) -> analytic Jacobians (residuals.cuh)
//   * Ceres' block-sparse Jacobian + SchurEliminator/normal-equation build -> fused J^T J / J^T r accumulation straight into
//     the band+arrow tile storage (problem.cuh); J is never materialised.
// linearize_kernel<TYPE>: one lane per residual evaluates r and J (fp64) into shared memory; the warp then reduces every run of
// consecutive residuals that share a knot span (same column -> position map; residual tables are chronological) into ONE
// J_run^T J_run contribution, so HBM sees one fp64 atomic per (run, column pair) instead of one per (residual, pair).
// The Huber corrector (ceres Corrector, rho'' <= 0 branch) and the EigenQuaternionParameterization tangent are folded in.
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "problem.cuh"

namespace lvi {

template <int TYPE> struct RTr {
  static constexpr int rows = (TYPE == RT_GYRO || TYPE == RT_ACCEL) ? 3 : (TYPE == RT_CAM ? 2 : 1);
  static constexpr int cols = TYPE == RT_GYRO ? 15 : TYPE == RT_ACCEL ? 29 : TYPE == RT_SURFEL ? 54 : TYPE == RT_CAM ? 55 : TYPE == RT_CAMSURF ? 60 : 12;
  static constexpr int shared_cols = TYPE == RT_CAM ? 54 : cols;  // the inverse-depth column differs per residual
  static constexpr bool two_eval = TYPE == RT_SURFEL || TYPE == RT_CAM || TYPE == RT_CAMSURF;
  static constexpr int warp_doubles = 32 * rows * cols + 32 * rows + 32;  // J, r, positions (ints, padded)
};

constexpr int kLinWarps = 2;

// (c1, c2), c1 >= c2, of the p-th column pair in row-major triangular order, packed c1 << 8 | c2.  One table serves every residual type:
// the pairs of an n-column block are the first n(n+1)/2 entries.  (Decoding p with sqrtf + fix-up loops cost ~25 of the ~60 instructions
// per accumulated pair.)
constexpr int kMaxPairs = LVI_MAX_COLS * (LVI_MAX_COLS + 1) / 2;
__device__ unsigned short g_pair_table[kMaxPairs];
static void ensure_pair_table(lvi_ctx* ctx) {   // the symbol lives on ONE device: uploaded once per context
  if (ctx->ks.pair_table) return;
  std::vector<unsigned short> h(kMaxPairs);
  int p = 0;
  for (int c1 = 0; c1 < LVI_MAX_COLS; ++c1)
    for (int c2 = 0; c2 <= c1; ++c2) h[p++] = static_cast<unsigned short>(c1 << 8 | c2);
  LVI_CUDA(cudaMemcpyToSymbol(g_pair_table, h.data(), sizeof(unsigned short) * kMaxPairs));
  ctx->ks.pair_table = true;
}

template <int TYPE>
__global__ void __launch_bounds__(kLinWarps * 32) linearize_kernel(ProblemView P, BandSys H, SchurView SV, double* __restrict__ g, double* __restrict__ cost) {
  constexpr int ROWS = RTr<TYPE>::rows, COLS = RTr<TYPE>::cols, RC = ROWS * COLS, SC = RTr<TYPE>::shared_cols;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* Jw = smem + warp * RTr<TYPE>::warp_doubles;
  double* rw = Jw + 32 * RC;
  int* posw = reinterpret_cast<int*>(rw + 32 * ROWS);
  const ResTable& T = P.tab[TYPE];
  double cost_acc = 0.0;
  const int stride = gridDim.x * kLinWarps * 32;
  for (int base = T.lo + (blockIdx.x * kLinWarps + warp) * 32; base < T.hi; base += stride) {
    const bool valid = base + lane < T.hi;
    const int i = valid ? (T.perm ? T.perm[base + lane] : base + lane) : 0;
    long long key = -1;
    if (valid) {
      ResOut o;
#pragma unroll
      for (int k = 0; k < ROWS; ++k)
        for (int c = 0; c < COLS; ++c) o.J[k][c] = 0.0;
      eval_residual<TYPE>(P, i, true, o);
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) s += o.r[k] * o.r[k];
      double sc;
      const double rho = huber(s, T.huber ? T.huber[i] : -1.0, sc);
      cost_acc += 0.5 * rho;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        rw[lane * ROWS + k] = o.r[k] * sc;
        for (int c = 0; c < COLS; ++c) Jw[lane * RC + k * COLS + c] = o.J[k][c] * sc;
      }
      key = T.i0a[i];
      if (RTr<TYPE>::two_eval) key |= static_cast<long long>(T.i0b[i]) << 32;
    }
    __syncwarp();
    const long long prev = __shfl_up_sync(FULL, key, 1);
    const bool head = valid && (lane == 0 || key != prev);
    unsigned heads = __ballot_sync(FULL, head);
    const int nvalid = __popc(__ballot_sync(FULL, valid));
    while (heads) {
      const int s = __ffs(heads) - 1;
      heads &= heads - 1;
      const int e = heads ? (__ffs(heads) - 1) : nvalid;
      const int i_head = __shfl_sync(FULL, i, s);
      for (int c = lane; c < COLS; c += 32) posw[c] = col_pos<TYPE>(P, i_head, c);
      __syncwarp();
      constexpr int NP = SC * (SC + 1) / 2;
      for (int p = lane; p < NP; p += 32) {
        const int pc = g_pair_table[p];
        const int c1 = pc >> 8, c2 = pc & 255;
        const int p1 = posw[c1], p2 = posw[c2];
        if (p1 < 0 || p2 < 0) continue;
        double acc = 0.0;
        for (int r = s; r < e; ++r) {
          const double* Jr = Jw + r * RC;
#pragma unroll
          for (int k = 0; k < ROWS; ++k) acc += Jr[k * COLS + c1] * Jr[k * COLS + c2];
        }
        if (c1 != c2 && p1 == p2) acc += acc;  // two columns of one residual on the same parameter (overlapping knot windows)
        if (acc != 0.0) atomicAdd(band_addr(H, max(p1, p2), min(p1, p2)), acc);
      }
      for (int c = lane; c < SC; c += 32) {
        const int pc = posw[c];
        if (pc < 0) continue;
        double acc = 0.0;
        for (int r = s; r < e; ++r)
#pragma unroll
          for (int k = 0; k < ROWS; ++k) acc += Jw[r * RC + k * COLS + c] * rw[r * ROWS + k];
        if (acc != 0.0) atomicAdd(g + pc, acc);
      }
      if (TYPE == RT_CAM) {  // inverse-depth column: one parameter per landmark, kept out of the band (eliminated by Schur complement)
        for (int r = s; r < e; ++r) {
          const int ir = __shfl_sync(FULL, i, r);
          const int lm = T.ia[ir];
          const int pr = P.pos_rho[lm];
          if (pr < 0) continue;
          const double* Jr = Jw + r * RC;
          const int rs = SV.row_start[lm], obs_base = T.ib[ir];
          for (int c = lane; c <= SC; c += 32) {
            if (c < SC && posw[c] < 0) continue;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < ROWS; ++k) acc += Jr[k * COLS + SC] * Jr[k * COLS + c];
            if (acc == 0.0) continue;
            if (c == SC) atomicAdd(SV.Hrr + (pr - SV.base), acc);
            else atomicAdd(SV.Hrx + rs + (c < 24 ? c : c < 48 ? obs_base + (c - 24) : 24 + (c - 48)), acc);
          }
          if (lane == 0) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < ROWS; ++k) acc += Jr[k * COLS + SC] * rw[r * ROWS + k];
            atomicAdd(g + pr, acc);
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost_acc += __shfl_xor_sync(FULL, cost_acc, o);
  if (lane == 0 && cost_acc != 0.0) atomicAdd(cost, cost_acc);
}

// ---- loss-corrected Jacobian rows for the tile gather (assemble.cu) ---------------------------------------------------------------------
// One lane per residual evaluates r and J (fp64, analytic) and stages them in shared memory; the warp then
//   * (camera table) adds the inverse-depth couplings J_rho^T [J, r] to the landmark's Schur row, lanes across columns,
//   * folds every second-window column whose knot the first window already holds into the first window's column (pos_kernel marks it -1), so
//     the positions of one residual's columns are distinct and the gather can scatter them without collisions,
//   * copies the rows out with coalesced stores: J [residual][row][column], r [residual][row] in table order.
template <int TYPE>
__global__ void __launch_bounds__(kLinWarps * 32) jacobian_kernel(ProblemView P, SchurView SV, double* __restrict__ Jg, double* __restrict__ g,
                                                                  double* __restrict__ cost) {
  constexpr int ROWS = RTr<TYPE>::rows, COLS = RTr<TYPE>::cols, RC = ROWS * COLS, SC = RTr<TYPE>::shared_cols;
  constexpr int RS = (RC + ROWS + 1) & ~1;   // asm_block_doubles(TYPE): [J rows | r | pad]
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* Jw = smem + warp * RTr<TYPE>::warp_doubles;
  double* rw = Jw + 32 * RC;
  const ResTable& T = P.tab[TYPE];
  double cost_acc = 0.0;
  const int stride = gridDim.x * kLinWarps * 32;
  for (int base = T.lo + (blockIdx.x * kLinWarps + warp) * 32; base < T.hi; base += stride) {
    const bool valid = base + lane < T.hi;
    const int i = base + lane;
    if (valid) {
      ResOut o;
#pragma unroll
      for (int k = 0; k < ROWS; ++k)
        for (int c = 0; c < COLS; ++c) o.J[k][c] = 0.0;
      eval_residual<TYPE>(P, i, true, o);
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) s += o.r[k] * o.r[k];
      double sc;
      const double rho = huber(s, T.huber ? T.huber[i] : -1.0, sc);
      cost_acc += 0.5 * rho;
#pragma unroll
      for (int k = 0; k < ROWS; ++k) {
        rw[lane * ROWS + k] = o.r[k] * sc;
        for (int c = 0; c < COLS; ++c) Jw[lane * RC + k * COLS + c] = o.J[k][c] * sc;
      }
    }
    __syncwarp();
    const int nvalid = min(32, T.hi - base);
    if (TYPE == RT_CAM) {  // inverse-depth column: one parameter per landmark, kept out of the band (eliminated by Schur complement)
      for (int r = 0; r < nvalid; ++r) {
        const int ir = base + r;
        const int lm = T.ia[ir];
        const int pr = P.pos_rho[lm];
        if (pr < 0) continue;
        const double* Jr = Jw + r * RC;
        const int rs = SV.row_start[lm], obs_base = T.ib[ir];
        for (int c = lane; c <= SC; c += 32) {   // slots with a constant parameter are skipped by their row_pos when the row is used
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < ROWS; ++k) acc += Jr[k * COLS + SC] * Jr[k * COLS + c];
          if (acc == 0.0) continue;
          if (c == SC) atomicAdd(SV.Hrr + (pr - SV.base), acc);
          else atomicAdd(SV.Hrx + rs + (c < 24 ? c : c < 48 ? obs_base + (c - 24) : 24 + (c - 48)), acc);
        }
        if (lane == 0) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < ROWS; ++k) acc += Jr[k * COLS + SC] * rw[r * ROWS + k];
          atomicAdd(g + pr, acc);
        }
      }
      __syncwarp();
    }
    if (RTr<TYPE>::two_eval && valid) {   // overlapping windows: same knot, same parameter
      const int shift = T.i0b[i] - T.i0a[i];
      if (shift > -4 && shift < 4) {
        double* Jr = Jw + lane * RC;
        for (int jp = 0; jp < 4; ++jp) {
          const int j = shift + jp;
          if (j < 0 || j > 3) continue;
          for (int k = 0; k < ROWS; ++k)
            for (int c = 0; c < 3; ++c) {
              Jr[k * COLS + 3 * j + c] += Jr[k * COLS + 24 + 3 * jp + c];           Jr[k * COLS + 24 + 3 * jp + c] = 0.0;
              Jr[k * COLS + 12 + 3 * j + c] += Jr[k * COLS + 36 + 3 * jp + c];      Jr[k * COLS + 36 + 3 * jp + c] = 0.0;
            }
        }
      }
    }
    __syncwarp();
    double* Jo = Jg + static_cast<size_t>(base) * RS;
    for (int e = lane; e < nvalid * RS; e += 32) {
      const int l = e / RS, o = e - l * RS;
      Jo[e] = o < RC ? Jw[l * RC + o] : (o < RC + ROWS ? rw[l * ROWS + (o - RC)] : 0.0);
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost_acc += __shfl_xor_sync(FULL, cost_acc, o);
  if (lane == 0 && cost_acc != 0.0) atomicAdd(cost, cost_acc);
}

// cost only (trial steps, fixed cost): one thread per residual
template <int TYPE>
__global__ void __launch_bounds__(128) cost_kernel(ProblemView P, double* __restrict__ cost) {
  constexpr int ROWS = RTr<TYPE>::rows;
  const ResTable& T = P.tab[TYPE];
  double acc = 0.0;
  for (int i = T.lo + blockIdx.x * blockDim.x + threadIdx.x; i < T.hi; i += gridDim.x * blockDim.x) {
    ResOut o;
    eval_residual<TYPE>(P, i, false, o);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ROWS; ++k) s += o.r[k] * o.r[k];
    double sc;
    acc += 0.5 * huber(s, T.huber ? T.huber[i] : -1.0, sc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[4];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double s = part[0] + part[1] + part[2] + part[3];
    if (s != 0.0) atomicAdd(cost, s);
  }
}

// corrected residuals (and optionally the dense tangent Jacobian) for parity tests
template <int TYPE>
__global__ void __launch_bounds__(64) residual_kernel(ProblemView P, int res_offset, int nt, double* __restrict__ res, double* __restrict__ Jd) {
  constexpr int ROWS = RTr<TYPE>::rows, COLS = RTr<TYPE>::cols;
  const ResTable& T = P.tab[TYPE];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T.n; i += gridDim.x * blockDim.x) {
    ResOut o;
    for (int k = 0; k < ROWS; ++k)
      for (int c = 0; c < COLS; ++c) o.J[k][c] = 0.0;
    eval_residual<TYPE>(P, i, Jd != nullptr, o);
    double s = 0.0;
    for (int k = 0; k < ROWS; ++k) s += o.r[k] * o.r[k];
    double sc;
    huber(s, T.huber ? T.huber[i] : -1.0, sc);
    for (int k = 0; k < ROWS; ++k) {
      const size_t row = static_cast<size_t>(res_offset) + static_cast<size_t>(i) * ROWS + k;
      if (res) res[row] = o.r[k] * sc;
      if (Jd && T.active)
        for (int c = 0; c < COLS; ++c) {
          const int p = col_pos<TYPE>(P, i, c);
          if (p >= 0) Jd[row * nt + p] += o.J[k][c] * sc;
        }
    }
  }
}

template <int TYPE>
static void launch_linearize(lvi_problem* p) {
  const ResTable& T = p->view.tab[TYPE];
  if (T.n == 0 || !T.active) return;
  ensure_pair_table(p->ctx);
  static_assert(RT_COUNT <= 8, "KernelState::lin_attr");
  bool& attr_set = p->ctx->ks.lin_attr[TYPE];
  const size_t smem = static_cast<size_t>(kLinWarps) * RTr<TYPE>::warp_doubles * sizeof(double);
  if (!attr_set) {
    LVI_CUDA(cudaFuncSetAttribute(linearize_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  const int per_cta = kLinWarps * 32;
  int grid = std::max(1, (T.hi - T.lo + per_cta - 1) / per_cta);
  const int cap = p->ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  static const char* const names[RT_COUNT] = {"linearize_kernel<RT_GYRO>", "linearize_kernel<RT_ACCEL>", "linearize_kernel<RT_SURFEL>", "linearize_kernel<RT_CAM>",
                                              "linearize_kernel<RT_CAMSURF>", "linearize_kernel<RT_ORIENT>"};
  LVI_LAUNCH_AS(p->ctx, names[TYPE], linearize_kernel<TYPE>, grid, per_cta, smem, p->view, p->H, p->schur, p->g.p, p->scal.p);
}

template <int TYPE>
static void launch_jacobian(lvi_problem* p) {
  const ResTable& T = p->view.tab[TYPE];
  if (T.n == 0 || !T.active || T.hi <= T.lo) return;
  bool& attr_set = p->ctx->ks.jac_attr[TYPE];
  const size_t smem = static_cast<size_t>(kLinWarps) * RTr<TYPE>::warp_doubles * sizeof(double);
  if (!attr_set) {
    LVI_CUDA(cudaFuncSetAttribute(jacobian_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  const int per_cta = kLinWarps * 32;
  int grid = std::max(1, (T.hi - T.lo + per_cta - 1) / per_cta);
  const int cap = p->ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  static const char* const names[RT_COUNT] = {"jacobian_kernel<RT_GYRO>", "jacobian_kernel<RT_ACCEL>", "jacobian_kernel<RT_SURFEL>", "jacobian_kernel<RT_CAM>",
                                              "jacobian_kernel<RT_CAMSURF>", "jacobian_kernel<RT_ORIENT>"};
  LVI_LAUNCH_AS(p->ctx, names[TYPE], jacobian_kernel<TYPE>, grid, per_cta, smem, p->view, p->schur_lin, p->asmp.J[TYPE].p, p->g_lin, p->cost_lin);
}

template <int TYPE>
static void launch_cost(lvi_problem* p, double* cost_d, bool want_active, bool want_inactive) {
  const ResTable& T = p->view.tab[TYPE];
  if (T.n == 0) return;
  if (T.active ? !want_active : !want_inactive) return;
  int grid = std::max(1, (T.hi - T.lo + 127) / 128);
  const int cap = p->ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  LVI_LAUNCH(p->ctx, cost_kernel<TYPE>, grid, 128, 0, p->view, cost_d);
}

template <int TYPE>
static void launch_residual(lvi_problem* p, double* res_d, double* J_d) {
  const ResTable& T = p->view.tab[TYPE];
  if (T.n == 0) return;
  int grid = (T.n + 63) / 64;
  const int cap = p->ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  LVI_LAUNCH(p->ctx, residual_kernel<TYPE>, grid, 64, 0, p->view, p->L.res_offset[TYPE], p->nt, res_d, J_d);
}

void problem_set_param_source(lvi_problem* p, const double* x) {
  p->view.r3 = x + p->off_r3; p->view.so3 = x + p->off_so3; p->view.sens = x + p->off_sens; p->view.rho = x + p->off_rho;
}

void problem_linearize(lvi_problem* p, double* cost_d) {
  cudaStream_t st = p->ctx->stream;
  problem_ensure_solver_buffers(p);
  if (p->p2p.active) {
    p2p_begin_linearize(p);   // clears this rank's units in the peer-memory region (after the peers have finished reading the previous ones)
  } else {
    p->H_tiles.zero(st); p->H_C.zero(st); p->g.zero(st); p->Hrx.zero(st); p->Hrr.zero(st);
    LVI_CUDA(cudaMemsetAsync(p->scal.p, 0, sizeof(double), st));
  }
  problem_set_param_source(p, p->X.p);
  // The per-type kernels only meet in fp64 atomics, and each of them leaves most of the machine idle (the camera kernel is 502 CTAs of
  // 64 threads at 255 registers): they run side by side on three streams -- camera | surfel | IMU and the rest -- and join again.
  lvi_ctx* ctx = p->ctx;
  LVI_CUDA(cudaEventRecord(ctx->ev_fork, st));
  for (int i = 0; i < 2; ++i) LVI_CUDA(cudaStreamWaitEvent(ctx->aux[i], ctx->ev_fork, 0));
  static const bool scatter = std::getenv("LVI_ASM_SCATTER") != nullptr;   // diagnostics: the first version (atomic scatter per run of residuals)
  if (scatter) launch_linearize<RT_CAM>(p); else launch_jacobian<RT_CAM>(p);
  {
    struct Restore { lvi_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, st};   // LVI_LAUNCH goes to ctx->stream
    const bool serial = std::getenv("LVI_LIN_SERIAL") != nullptr;   // diagnostics: every table on the main stream (per-kernel times mean something)
    ctx->stream = serial ? st : ctx->aux[0];
    if (scatter) launch_linearize<RT_SURFEL>(p); else launch_jacobian<RT_SURFEL>(p);
    ctx->stream = serial ? st : ctx->aux[1];
    if (scatter) { launch_linearize<RT_ACCEL>(p); launch_linearize<RT_GYRO>(p); launch_linearize<RT_CAMSURF>(p); launch_linearize<RT_ORIENT>(p); }
    else { launch_jacobian<RT_ACCEL>(p); launch_jacobian<RT_GYRO>(p); launch_jacobian<RT_CAMSURF>(p); launch_jacobian<RT_ORIENT>(p); }
  }
  for (int i = 0; i < 2; ++i) {
    LVI_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->aux[i]));
    LVI_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
  }
  if (!scatter) assemble_gather(p);   // H tiles, corner and g from the Jacobian rows (owner-computes tile gather, fp64 tensor pipe)
  if (!scatter && std::getenv("LVI_ASM_CHECK")) {   // diagnostics: the atomic-scatter kernels on the same point, compared element by element
    const BandSys& H = p->H;
    std::vector<double> t1(p->H_tiles.n), c1(p->H_C.n), g1(p->g.n), t2(p->H_tiles.n), c2(p->H_C.n), g2(p->g.n);
    p->H_tiles.download(t1.data(), t1.size(), st); p->H_C.download(c1.data(), c1.size(), st); p->g.download(g1.data(), g1.size(), st);
    LVI_CUDA(cudaStreamSynchronize(st));
    p->H_tiles.zero(st); p->H_C.zero(st); p->g.zero(st); p->Hrx.zero(st); p->Hrr.zero(st);
    LVI_CUDA(cudaMemsetAsync(p->scal.p, 0, sizeof(double), st));
    launch_linearize<RT_CAM>(p); launch_linearize<RT_SURFEL>(p); launch_linearize<RT_ACCEL>(p); launch_linearize<RT_GYRO>(p); launch_linearize<RT_CAMSURF>(p); launch_linearize<RT_ORIENT>(p);
    p->H_tiles.download(t2.data(), t2.size(), st); p->H_C.download(c2.data(), c2.size(), st); p->g.download(g2.data(), g2.size(), st);
    LVI_CUDA(cudaStreamSynchronize(st));
    int shown = 0;
    double worst = 0, scale_ = 0;
    for (size_t e = 0; e < t1.size(); ++e) {
      const size_t tile = e >> 10; const int J = static_cast<int>(tile / H.TPC), s = static_cast<int>(tile % H.TPC), b = (e >> 5) & 31, a = e & 31;
      if (s == 0 && a < b) continue;   // upper triangle of a diagonal tile: not part of the store
      const double d = std::fabs(t1[e] - t2[e]);
      scale_ = std::max(scale_, std::fabs(t2[e]));
      if (d > worst) worst = d;
      if (d > 1e-9 * (1.0 + std::fabs(t2[e])) && shown < 12) { std::fprintf(stderr, "[asm check] tile J=%d s=%d (a=%d,b=%d): gather %.12g scatter %.12g\n", J, s, a, b, t1[e], t2[e]); ++shown; }
    }
    double worst_c = 0, worst_g = 0;
    for (int bj = 0; bj < H.ldc; ++bj) for (int bi = bj; bi < H.ldc; ++bi) {
      const size_t e = bi + static_cast<size_t>(H.ldc) * bj;
      const double d = std::fabs(c1[e] - c2[e]);
      worst_c = std::max(worst_c, d);
      if (d > 1e-9 * (1.0 + std::fabs(c2[e])) && shown < 24) { std::fprintf(stderr, "[asm check] corner (%d,%d): gather %.12g scatter %.12g\n", bi, bj, c1[e], c2[e]); ++shown; }
    }
    for (size_t e = 0; e < g1.size(); ++e) {
      const double d = std::fabs(g1[e] - g2[e]);
      worst_g = std::max(worst_g, d);
      if (d > 1e-9 * (1.0 + std::fabs(g2[e])) && shown < 36) { std::fprintf(stderr, "[asm check] g[%zu]: gather %.12g scatter %.12g\n", e, g1[e], g2[e]); ++shown; }
    }
    std::fprintf(stderr, "[asm check] nb %d nbo %d NT %d T %d RB %d | items %d entries %d | max |dH| %.3g (max |H| %.3g), corner %.3g, g %.3g\n", H.nb, H.nbo, H.NT, H.T, H.RB,
                 p->asmp.n_items, p->asmp.n_entries, worst, scale_, worst_c, worst_g);
  }
  if (cost_d && cost_d != p->cost_lin) LVI_CUDA(cudaMemcpyAsync(cost_d, p->cost_lin, sizeof(double), cudaMemcpyDeviceToDevice, st));
}

void problem_cost(lvi_problem* p, const double* x_d, double* cost_d, bool active, bool inactive) {
  LVI_CUDA(cudaMemsetAsync(cost_d, 0, sizeof(double), p->ctx->stream));
  problem_set_param_source(p, x_d);
  {  // like the normal-equation kernels: camera | surfel | the rest, side by side
    lvi_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    LVI_CUDA(cudaEventRecord(ctx->ev_fork, st));
    for (int i = 0; i < 2; ++i) LVI_CUDA(cudaStreamWaitEvent(ctx->aux[i], ctx->ev_fork, 0));
    launch_cost<RT_CAM>(p, cost_d, active, inactive);
    {
      struct Restore { lvi_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, st};
      ctx->stream = ctx->aux[0];
      launch_cost<RT_SURFEL>(p, cost_d, active, inactive);
      ctx->stream = ctx->aux[1];
      launch_cost<RT_GYRO>(p, cost_d, active, inactive); launch_cost<RT_ACCEL>(p, cost_d, active, inactive);
      launch_cost<RT_CAMSURF>(p, cost_d, active, inactive); launch_cost<RT_ORIENT>(p, cost_d, active, inactive);
    }
    for (int i = 0; i < 2; ++i) {
      LVI_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->aux[i]));
      LVI_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
    }
  }
  problem_set_param_source(p, p->X.p);
}

// Distinct positions of every landmark's Schur row (static per problem): the windows of consecutive observations overlap, so ~294 slots
// hold ~195 distinct parameters; eliminating over the merged row needs 2.3x fewer updates.  One CTA per landmark: bitonic sort of
// (position, slot) keys in shared memory, heads of equal-position runs numbered by a scan.
__global__ void __launch_bounds__(256) schur_merge_plan_kernel(const int* __restrict__ row_start, const int* __restrict__ row_pos, int n_landmarks,
                                                               int* __restrict__ slot2u, int* __restrict__ urow_pos, int* __restrict__ ulen, int* __restrict__ fail) {
  __shared__ long long key[1024];
  __shared__ int rank[1024];
  const int l = blockIdx.x;
  const int rs = row_start[l], len = row_start[l + 1] - rs;
  if (len == 0) { if (threadIdx.x == 0) ulen[l] = 0; return; }
  if (len > 1024) { if (threadIdx.x == 0) { *fail = 6; ulen[l] = 0; } return; }
  int n2 = 1;
  while (n2 < len) n2 <<= 1;
  for (int e = threadIdx.x; e < n2; e += 256) {
    long long k = 0x7fffffffffffffffll;
    if (e < len) { const int p = row_pos[rs + e]; if (p >= 0) k = (static_cast<long long>(p) << 16) | e; else slot2u[rs + e] = -1; }
    key[e] = k;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = threadIdx.x; e < n2; e += 256) {
        const int x = e ^ j;
        if (x > e) {
          const bool up = (e & k) == 0;
          const long long a = key[e], b = key[x];
          if ((a > b) == up) { key[e] = b; key[x] = a; }
        }
      }
      __syncthreads();
    }
  for (int e = threadIdx.x; e < n2; e += 256) {
    const long long k = key[e];
    const bool valid = k != 0x7fffffffffffffffll;
    rank[e] = (valid && (e == 0 || (key[e - 1] >> 16) != (k >> 16))) ? 1 : 0;
  }
  __syncthreads();
  for (int o = 1; o < n2; o <<= 1) {   // inclusive scan
    int v[4];
    int c = 0;
    for (int e = threadIdx.x; e < n2; e += 256) v[c++] = e >= o ? rank[e - o] : 0;
    __syncthreads();
    c = 0;
    for (int e = threadIdx.x; e < n2; e += 256) rank[e] += v[c++];
    __syncthreads();
  }
  for (int e = threadIdx.x; e < n2; e += 256) {
    const long long k = key[e];
    if (k == 0x7fffffffffffffffll) continue;
    const int m = rank[e] - 1, slot = static_cast<int>(k & 0xffff), p = static_cast<int>(k >> 16);
    slot2u[rs + slot] = m;
    if (e == 0 || (key[e - 1] >> 16) != (k >> 16)) urow_pos[rs + m] = p;
  }
  if (threadIdx.x == 0) {
    int last = 0;
    for (int e = n2 - 1; e >= 0; --e) if (key[e] != 0x7fffffffffffffffll) { last = rank[e]; break; }
    ulen[l] = last;
  }
}

static void alloc_bandsys(BandSys& S, const Lowered& L, DBuf<double>& tiles, DBuf<double>& C, double* Linv, double* x, int* fail) {
  S.nb = L.nb; S.nbo = L.nbo;
  S.NT = (L.nb + kTile - 1) / kTile;
  S.NT0 = L.chain1_start < L.nb ? L.chain1_start / kTile : S.NT;
  S.n_mid = L.n_mid;
  S.T = S.NT > 0 ? std::min(S.NT - 1, (L.bw + kTile - 1) / kTile) : 0;
  S.RB = (L.nbo + 1 + kTile - 1) / kTile;  // + the rhs row
  S.TPC = S.T + 1 + S.RB;
  S.ldc = S.RB * kTile;
  tiles.alloc(std::max<size_t>(static_cast<size_t>(S.NT) * S.TPC * kTileElems, 1));
  C.alloc(static_cast<size_t>(S.ldc) * S.ldc);
  S.tiles = tiles.p; S.C = C.p; S.Linv = Linv; S.x = x; S.fail = fail; S.work_i = nullptr; S.work_d = nullptr; S.trace = nullptr; S.ll = nullptr; S.epoch = 0;
}

void problem_ensure_solver_buffers(lvi_problem* p) {
  if (p->has_solver_buffers) return;
  const Lowered& L = p->L;
  p->fail.alloc(4);
  const int NT = (L.nb + kTile - 1) / kTile;
  const int RB = (L.nbo + 1 + kTile - 1) / kTile;
  p->A_Linv.alloc(std::max<size_t>(static_cast<size_t>(NT) * kTileElems, 1));
  p->A_x.alloc(static_cast<size_t>(NT) * kTile + static_cast<size_t>(RB) * kTile);
  alloc_bandsys(p->H, L, p->H_tiles, p->H_C, nullptr, nullptr, nullptr);
  alloc_bandsys(p->A, L, p->A_tiles, p->A_C, p->A_Linv.p, p->A_x.p, p->fail.p);
  p->A_work_i.alloc(p->A.work_i_count()); p->A_work_d.alloc(std::max<size_t>(p->A.work_d_count(), 1));
  p->A.work_i = p->A_work_i.p; p->A.work_d = p->A_work_d.p;
  p->A_ll.alloc(std::max<size_t>(p->A.ll_count(), 1)); p->A_ll.zero(p->ctx->stream); p->A.ll = p->A_ll.p;
  p->A2 = BandSys{};
  if (p->A.n_mid > 0) {  // second-level system: the separator block of the corner, factored with the same tile machinery
    init_second_level(p->A, p->A2);
    BandSys& B = p->A2;
    p->A2_tiles.alloc(static_cast<size_t>(B.NT) * B.TPC * kTileElems); p->A2_C.alloc(static_cast<size_t>(B.ldc) * B.ldc);
    p->A2_Linv.alloc(static_cast<size_t>(B.NT) * kTileElems); p->A2_x.alloc(static_cast<size_t>(B.NT) * kTile + B.ldc);
    p->A2_work_i.alloc(B.work_i_count()); p->A2_work_d.alloc(B.work_d_count());
    B.tiles = p->A2_tiles.p; B.C = p->A2_C.p; B.Linv = p->A2_Linv.p; B.x = p->A2_x.p; B.fail = p->fail.p;
    B.work_i = p->A2_work_i.p; B.work_d = p->A2_work_d.p; B.trace = nullptr;
    p->A2_ll.alloc(B.ll_count()); p->A2_ll.zero(p->ctx->stream); B.ll = p->A2_ll.p; B.epoch = 0;
  }
  LVI_REQUIRE(p->A.ldc <= 1024, LVI_ERR_INVALID, "arrow border wider than 1023 dims is not supported");
  const size_t nt = std::max(p->nt, 1);
  {  // Schur rows of the inverse depths
    cudaStream_t st = p->ctx->stream;
    std::vector<int> lm;
    for (int l = 0; l < L.n_landmarks; ++l) if (L.pos_rho[l] >= 0) lm.push_back(l);
    p->row_start.alloc(L.row_start.size()); p->row_start.upload(L.row_start.data(), L.row_start.size(), st);
    p->row_pos.alloc(L.row_pos.size()); p->row_pos.upload(L.row_pos.data(), L.row_pos.size(), st);
    p->lm_of_rho.alloc(std::max<size_t>(lm.size(), 1)); p->lm_of_rho.upload(lm.data(), lm.size(), st);
    p->Hrx.alloc(L.row_pos.size()); p->Hrr.alloc(std::max(L.n_rho, 1)); p->yrho.alloc(std::max(L.n_rho, 1));
    p->slot2u.alloc(L.row_pos.size()); p->urow_pos.alloc(L.row_pos.size()); p->ulen.alloc(std::max<size_t>(L.row_start.size() - 1, 1));
    if (L.row_start.size() > 1) {
      LVI_CUDA(cudaMemsetAsync(p->fail.p + 3, 0, sizeof(int), st));
      LVI_LAUNCH(p->ctx, schur_merge_plan_kernel, static_cast<int>(L.row_start.size() - 1), 256, 0, p->row_start.p, p->row_pos.p, static_cast<int>(L.row_start.size() - 1),
                 p->slot2u.p, p->urow_pos.p, p->ulen.p, p->fail.p + 3);
      int fail = 0;
      LVI_CUDA(cudaMemcpyAsync(&fail, p->fail.p + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
      LVI_CUDA(cudaStreamSynchronize(st));
      LVI_REQUIRE(fail == 0, LVI_ERR_INVALID, "a landmark couples to more than 1024 parameter slots");
    }
    LVI_CUDA(cudaStreamSynchronize(st));
    p->schur = SchurView{L.n_rho, L.nb + L.nbo, p->row_start.p, p->row_pos.p, p->lm_of_rho.p, p->slot2u.p, p->urow_pos.p, p->ulen.p, p->Hrx.p, p->Hrr.p, p->yrho.p};
  }
  p->g.alloc(nt); p->scale.alloc(nt); p->diag.alloc(nt); p->y.alloc(nt); p->delta.alloc(nt);
  if (p->ctx->world > 1) {  // the all-reduce of H moves only the tiles that can be non-zero (about half of the store at C2)
    const BandSys& H = p->H;
    const int RBsep = H.n_mid >> kTileLog;
    std::vector<int> map;
    for (int j = 0; j < H.NT; ++j) {
      const int c_start = j < H.NT0 ? 0 : H.NT0, c_end = j < H.NT0 ? H.NT0 : H.NT;
      const int sep_first = std::max(c_start, c_end - H.T - 1);
      for (int s = 0; s < H.TPC; ++s) {
        const bool band = s <= H.T;
        if (band && j + s >= c_end) continue;                              // below the end of the chain: never referenced
        if (!band && (s - H.T - 1) < RBsep && j < sep_first) continue;     // separator row outside its coupling range: structurally zero
        map.push_back(j * H.TPC + s);
      }
    }
    p->pack_map.alloc(std::max<size_t>(map.size(), 1));
    p->pack_map.upload(map.data(), map.size(), p->ctx->stream);
    // one staging buffer for ONE collective per iteration: [packed tiles | corner | g | Schur rows | Schur diagonal | cost]
    p->pack_buf.alloc(map.size() * kTileElems + p->H_C.n + p->g.n + p->Hrx.n + p->Hrr.n + 1);
    p->n_pack = static_cast<int>(map.size());
    LVI_CUDA(cudaStreamSynchronize(p->ctx->stream));
  }
  p->has_solver_buffers = true;
  p->H_lin = p->H; p->schur_lin = p->schur; p->g_lin = p->g.p; p->cost_lin = p->scal.p;
  assemble_build_plan(p);
  assemble_build_schur_plan(p);
  if (p->ctx->world > 1 && p2p_prepare(p)) {   // the peers' contributions are read straight out of their HBM: the private store is only ever written
    p->H_tiles.zero(p->ctx->stream);
  }
}

void problem_download_params(lvi_problem* p) {
  cudaStream_t st = p->ctx->stream;
  std::vector<double> h(p->nx);
  p->X.download(h.data(), p->nx, st);
  LVI_CUDA(cudaStreamSynchronize(st));
  const lvi_problem_desc& d = p->desc;
  const int n = d.n_knots;
  if (d.r3_knots) std::memcpy(d.r3_knots, h.data() + p->off_r3, sizeof(double) * 3 * n);
  std::memcpy(d.so3_knots, h.data() + p->off_so3, sizeof(double) * 4 * n);
  const double* s = h.data() + p->off_sens;
  if (d.lidar_q) std::memcpy(d.lidar_q, s + SENS_LQ, 32);
  if (d.lidar_p) std::memcpy(d.lidar_p, s + SENS_LP, 24);
  if (d.cam_q) std::memcpy(d.cam_q, s + SENS_CQ, 32);
  if (d.cam_p) std::memcpy(d.cam_p, s + SENS_CP, 24);
  if (d.gravity) std::memcpy(d.gravity, s + SENS_G, 16);
  if (d.acc_bias) std::memcpy(d.acc_bias, s + SENS_BA, 24);
  if (d.gyr_bias) std::memcpy(d.gyr_bias, s + SENS_BG, 24);
  if (d.rho && d.n_landmarks) std::memcpy(d.rho, h.data() + p->off_rho, sizeof(double) * d.n_landmarks);
}

static lvi_problem* create_problem(lvi_ctx* ctx, const lvi_problem_desc* d) {
  auto p = std::unique_ptr<lvi_problem>(new lvi_problem());
  p->ctx = ctx; p->desc = *d;
  Lowered& L = p->L;
  const bool timing = std::getenv("LVI_TIME_CREATE") != nullptr;   // diagnostics: host phases of problem creation on stderr
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto T1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[lvi] create: %-22s %.3f ms\n", what, std::chrono::duration<double, std::milli>(T1 - T0).count());
    T0 = T1;
  };
  try {
    lower_problem(*d, L);
  } catch (const RangeError& e) {
    throw Error(LVI_ERR_RANGE, e.what());
  } catch (const std::invalid_argument& e) {
    throw Error(LVI_ERR_INVALID, e.what());
  }
  // unit-quaternion check of the control points (UniformSO3SplineEntity validator, K/trajectories/uniform_so3_spline_trajectory.h:23-27)
  for (int i = 0; i < d->n_knots; ++i) {
    const double* q = d->so3_knots + 4 * i;
    const double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (std::fabs(nrm - 1.0) > 1e-5) throw Error(LVI_ERR_DOMAIN, "SO3 control point is not a unit quaternion");
  }
  lap("lower_problem");
  double sens[SENS_N];
  pack_sens(*d, sens);
  {
    ProblemView hv = host_view(*d, L, sens);
    compute_bandwidth(hv, L);
  }
  lap("compute_bandwidth");
  cudaStream_t st = ctx->stream;
  const int n = d->n_knots, nl = d->n_landmarks;
  p->off_r3 = 0; p->off_so3 = 3 * n; p->off_sens = 7 * n; p->off_rho = 7 * n + SENS_N; p->nx = 7 * n + SENS_N + std::max(nl, 1);
  std::vector<double> hx(p->nx, 0.0);
  if (d->r3_knots) std::memcpy(hx.data(), d->r3_knots, sizeof(double) * 3 * n);
  std::memcpy(hx.data() + p->off_so3, d->so3_knots, sizeof(double) * 4 * n);
  std::memcpy(hx.data() + p->off_sens, sens, sizeof(sens));
  if (nl && d->rho) std::memcpy(hx.data() + p->off_rho, d->rho, sizeof(double) * nl);
  p->X.alloc(p->nx); p->XC.alloc(p->nx); p->XS.alloc(p->nx);
  p->X.upload(hx.data(), p->nx, st);
  LVI_CUDA(cudaMemcpyAsync(p->XC.p, p->X.p, sizeof(double) * p->nx, cudaMemcpyDeviceToDevice, st));
  // tables
  ProblemView& V = p->view;
  V = ProblemView{};
  V.dt_inv = 1.0 / d->dt; V.n_knots = n; V.fx = d->fx; V.fy = d->fy; V.cx = d->cx; V.cy = d->cy; V.has_r3 = L.has_r3;
  for (int k = 0; k < 5; ++k) V.dist[k] = d->distortion[k];
  V.do_dist = (std::fabs(d->distortion[0]) > 1e-5 || std::fabs(d->distortion[1]) > 1e-5 || std::fabs(d->distortion[2]) > 1e-5) ? 1 : 0;   // pinhole_camera.h:78
  auto up_i = [&](DBuf<int>& b, const std::vector<int>& h) -> const int* { if (h.empty()) return nullptr; b.alloc(h.size()); b.upload(h.data(), h.size(), st); return b.p; };
  auto up_d = [&](DBuf<double>& b, const std::vector<double>& h) -> const double* { if (h.empty()) return nullptr; b.alloc(h.size()); b.upload(h.data(), h.size(), st); return b.p; };
  for (int t = 0; t < RT_COUNT; ++t) {
    const LoweredTable& T = L.tab[t];
    ResTable& R = V.tab[t];
    R.n = T.n; R.active = T.active;
    shard_range(T.n, ctx->rank, ctx->world, R.lo, R.hi);
    R.i0a = up_i(p->tab_i[t][0], T.i0a); R.i0b = up_i(p->tab_i[t][1], T.i0b); R.ia = up_i(p->tab_i[t][2], T.ia); R.ib = up_i(p->tab_i[t][3], T.ib); R.perm = up_i(p->tab_i[t][4], T.perm);
    R.ua = up_d(p->tab_d[t][0], T.ua); R.ub = up_d(p->tab_d[t][1], T.ub); R.v = up_d(p->tab_d[t][2], T.v);
    R.weight = up_d(p->tab_d[t][3], T.weight); R.huber = up_d(p->tab_d[t][4], T.huber);
  }
  lap("tables alloc+upload");
  V.pos_r3 = up_i(p->pos_r3, L.pos_r3); V.pos_so3 = up_i(p->pos_so3, L.pos_so3); V.pos_rho = up_i(p->pos_rho, L.pos_rho);
  for (int b = 0; b < TB_COUNT; ++b) V.pos_sens[b] = L.pos_sens[b];
  if (d->n_planes > 0) {
    LVI_REQUIRE(d->planes, LVI_ERR_INVALID, "planes is null");
    p->planes.alloc(3 * static_cast<size_t>(d->n_planes));
    p->planes.upload(d->planes, 3 * static_cast<size_t>(d->n_planes), st);
  }
  V.planes = p->planes.p;
  problem_set_param_source(p.get(), p->X.p);
  // free parameter blocks
  std::vector<FreeBlock> fb;
  for (int i = 0; i < n; ++i) {
    if (L.pos_r3[i] >= 0) fb.push_back({p->off_r3 + 3 * i, L.pos_r3[i], 3, 0});
    if (L.pos_so3[i] >= 0) fb.push_back({p->off_so3 + 4 * i, L.pos_so3[i], 4, 1});
  }
  const int s_off[TB_COUNT] = {SENS_LQ, SENS_LP, SENS_CQ, SENS_CP, SENS_G, SENS_BA, SENS_BG};
  const int s_size[TB_COUNT] = {4, 3, 4, 3, 2, 3, 3};
  const int s_kind[TB_COUNT] = {1, 0, 1, 0, 0, 0, 0};
  for (int b = 0; b < TB_COUNT; ++b)
    if (L.pos_sens[b] >= 0) fb.push_back({p->off_sens + s_off[b], L.pos_sens[b], s_size[b], s_kind[b]});
  for (int l = 0; l < nl; ++l)
    if (L.pos_rho[l] >= 0) fb.push_back({p->off_rho + l, L.pos_rho[l], 1, 2});
  p->n_blocks = static_cast<int>(fb.size());
  p->blocks.alloc(std::max<size_t>(fb.size(), 1));
  p->blocks.upload(fb.data(), fb.size(), st);
  p->nt = L.nt();
  p->scal.alloc(64);
  p->h_scal = ctx->h_scal;   // pinned mirror, owned by the context
  LVI_CUDA(cudaStreamSynchronize(st));  // host staging vectors go out of scope
  lap("rest + sync");
  return p.release();
}

}  // namespace lvi

using namespace lvi;

extern "C" {

void lvi_solve_options_default(lvi_solve_options* o) {
  if (!o) return;
  o->max_num_iterations = 30; o->verbose = 0;  // K/trajectory_estimator.h:38 Solve(max_iterations = 30)
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
}

int lvi_problem_create(lvi_ctx* ctx, const lvi_problem_desc* desc, lvi_problem** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && desc && out, LVI_ERR_INVALID, "lvi_problem_create: null argument");
    activate(ctx);
    *out = create_problem(ctx, desc);
  });
}

int lvi_problem_destroy(lvi_problem* p) {
  if (p) { cudaSetDevice(p->ctx->device); tl_stream = p->ctx->stream; delete p; }
  return LVI_OK;
}

int lvi_problem_num_residuals(const lvi_problem* p) { return p ? p->L.n_res : 0; }
int lvi_problem_num_tangent(const lvi_problem* p) { return p ? p->nt : 0; }
int lvi_problem_tangent_offset_knot(const lvi_problem* p, int knot) {
  if (!p || knot < 0 || knot >= p->L.n_knots) return -1;
  return p->L.pos_r3[knot] >= 0 ? p->L.pos_r3[knot] : p->L.pos_so3[knot];
}
int lvi_problem_tangent_offset_block(const lvi_problem* p, int which) {
  if (!p || which < 0) return -1;
  if (which < TB_COUNT) return p->L.pos_sens[which];
  const int l = which - TB_COUNT;
  return l < p->L.n_landmarks ? p->L.pos_rho[l] : -1;
}

int lvi_problem_evaluate(lvi_problem* p, double* cost, double* residuals, double* gradient) {
  return guarded([&] {
    LVI_REQUIRE(p, LVI_ERR_INVALID, "lvi_problem_evaluate: null problem");
    activate(p->ctx);
    cudaStream_t st = p->ctx->stream;
    if (gradient) {
      problem_linearize(p, nullptr);
      p->g.download(gradient, p->nt, st);
      if (cost) LVI_CUDA(cudaMemcpyAsync(cost, p->scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    } else if (cost) {
      problem_cost(p, p->X.p, p->scal.p, true, false);
      LVI_CUDA(cudaMemcpyAsync(cost, p->scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (residuals) {
      DBuf<double> r(std::max(p->L.n_res, 1));
      launch_residual<RT_GYRO>(p, r.p, nullptr); launch_residual<RT_ACCEL>(p, r.p, nullptr); launch_residual<RT_SURFEL>(p, r.p, nullptr);
      launch_residual<RT_CAM>(p, r.p, nullptr); launch_residual<RT_CAMSURF>(p, r.p, nullptr); launch_residual<RT_ORIENT>(p, r.p, nullptr);
      r.download(residuals, p->L.n_res, st);
      LVI_CUDA(cudaStreamSynchronize(st));
    }
    LVI_CUDA(cudaStreamSynchronize(st));
  });
}

int lvi_problem_jacobian_dense(lvi_problem* p, double* J) {
  return guarded([&] {
    LVI_REQUIRE(p && J, LVI_ERR_INVALID, "lvi_problem_jacobian_dense: null argument");
    activate(p->ctx);
    const size_t n = static_cast<size_t>(p->L.n_res) * p->nt;
    LVI_REQUIRE(n < (1ull << 28), LVI_ERR_INVALID, "lvi_problem_jacobian_dense: problem too large for a dense Jacobian");
    cudaStream_t st = p->ctx->stream;
    DBuf<double> Jd(std::max<size_t>(n, 1)), r(std::max(p->L.n_res, 1));
    Jd.zero(st);
    launch_residual<RT_GYRO>(p, r.p, Jd.p); launch_residual<RT_ACCEL>(p, r.p, Jd.p); launch_residual<RT_SURFEL>(p, r.p, Jd.p);
    launch_residual<RT_CAM>(p, r.p, Jd.p); launch_residual<RT_CAMSURF>(p, r.p, Jd.p); launch_residual<RT_ORIENT>(p, r.p, Jd.p);
    Jd.download(J, n, st);
    LVI_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"

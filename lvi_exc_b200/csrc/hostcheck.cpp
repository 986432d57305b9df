// hostcheck.cpp — CPU instantiation of the SAME lowering + residual/Jacobian headers the CUDA kernels use
// (lowering.hpp, residuals.cuh, spline_math.cuh are __host__ __device__).  Built into liblvi_hostcheck.so and used ONLY
// by the `-m "not gpu"` tests to validate the analytic Jacobians against the oracle's forward-mode Jets without a GPU.
// It is not linked into liblvi_exc_b200.so and is not reachable from the C-ABI: the product has no CPU path.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "lowering.hpp"

using namespace lvi;

namespace {
thread_local std::string g_err;

template <int TYPE>
void eval_type(const ProblemView& P, const Lowered& L, double& cost, double& fixed, double* res, double* J, int nt) {
  const ResTable& T = P.tab[TYPE];
  const int rows = rt_rows(TYPE), cols = rt_cols(TYPE);
  for (int i = T.lo; i < T.hi; ++i) {
    ResOut o;
    std::memset(&o, 0, sizeof(o));
    eval_residual<TYPE>(P, i, true, o);
    double s = 0;
    for (int k = 0; k < rows; ++k) s += o.r[k] * o.r[k];
    double scale;
    const double rho = huber(s, T.huber ? T.huber[i] : -1.0, scale);
    (T.active ? cost : fixed) += 0.5 * rho;
    const int ro = L.res_offset[TYPE] + i * rows;
    for (int k = 0; k < rows; ++k) {
      if (res) res[ro + k] = o.r[k] * scale;
      if (J && T.active)
        for (int c = 0; c < cols; ++c) {
          const int p = col_pos<TYPE>(P, i, c);
          if (p >= 0) J[static_cast<size_t>(ro + k) * nt + p] += o.J[k][c] * scale;
        }
    }
  }
}
}  // namespace

extern "C" {

const char* lvi_hostcheck_last_error() { return g_err.c_str(); }

// sizes: out[0]=n_res out[1]=nt out[2]=nb out[3]=nbo out[4]=bw out[5]=chain1_start out[6]=n_mid out[7]=n_pad ; pos arrays may be NULL
int lvi_hostcheck_layout(const lvi_problem_desc* d, int* out, int* pos_r3, int* pos_so3, int* pos_sens, int* pos_rho) {
  try {
    Lowered L;
    lower_problem(*d, L);
    double sens[SENS_N];
    pack_sens(*d, sens);
    ProblemView P = host_view(*d, L, sens);
    compute_bandwidth(P, L);
    out[0] = L.n_res; out[1] = L.nt(); out[2] = L.nb; out[3] = L.nbo; out[4] = L.bw; out[5] = L.chain1_start; out[6] = L.n_mid; out[7] = L.n_pad;
    if (pos_r3) std::memcpy(pos_r3, L.pos_r3.data(), sizeof(int) * d->n_knots);
    if (pos_so3) std::memcpy(pos_so3, L.pos_so3.data(), sizeof(int) * d->n_knots);
    if (pos_sens) std::memcpy(pos_sens, L.pos_sens, sizeof(int) * TB_COUNT);
    if (pos_rho && d->n_landmarks) std::memcpy(pos_rho, L.pos_rho.data(), sizeof(int) * d->n_landmarks);
    return LVI_OK;
  } catch (const RangeError& e) { g_err = e.what(); return LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_INVALID; }
}

int lvi_hostcheck_evaluate(const lvi_problem_desc* d, double* cost, double* fixed_cost, double* residuals, double* J_dense) {
  try {
    Lowered L;
    lower_problem(*d, L);
    double sens[SENS_N];
    pack_sens(*d, sens);
    ProblemView P = host_view(*d, L, sens);
    const int nt = L.nt();
    if (J_dense) std::memset(J_dense, 0, sizeof(double) * static_cast<size_t>(L.n_res) * nt);
    double c = 0, f = 0;
    eval_type<RT_GYRO>(P, L, c, f, residuals, J_dense, nt);
    eval_type<RT_ACCEL>(P, L, c, f, residuals, J_dense, nt);
    eval_type<RT_SURFEL>(P, L, c, f, residuals, J_dense, nt);
    eval_type<RT_CAM>(P, L, c, f, residuals, J_dense, nt);
    eval_type<RT_CAMSURF>(P, L, c, f, residuals, J_dense, nt);
    eval_type<RT_ORIENT>(P, L, c, f, residuals, J_dense, nt);
    if (cost) *cost = c;
    if (fixed_cost) *fixed_cost = f;
    return LVI_OK;
  } catch (const RangeError& e) { g_err = e.what(); return LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_INVALID; }
}

}  // extern "C"

// One rank's share of the normal equations under the library's data-parallel sharding (ResTable lo/hi from shard_range, exactly
// as lvi_problem_create sets them): cost, g = J^T r [nt], H = J^T J [nt x nt, dense, row-major].  The layout comes from the FULL
// problem, so the shares of all ranks add up to the single-rank result (tests/test_multi_gloo.py all-reduces them over gloo).
extern "C" int lvi_hostcheck_normal_equations(const lvi_problem_desc* d, int rank, int world, double* cost, double* g, double* H) {
  try {
    Lowered L;
    lower_problem(*d, L);
    double sens[SENS_N];
    pack_sens(*d, sens);
    ProblemView P = host_view(*d, L, sens);
    compute_bandwidth(P, L);
    for (int t = 0; t < RT_COUNT; ++t) shard_range(P.tab[t].n, rank, world, P.tab[t].lo, P.tab[t].hi);
    const int nt = L.nt();
    std::vector<double> res(L.n_res, 0.0), J(static_cast<size_t>(L.n_res) * nt, 0.0);
    double c = 0, f = 0;
    eval_type<RT_GYRO>(P, L, c, f, res.data(), J.data(), nt);
    eval_type<RT_ACCEL>(P, L, c, f, res.data(), J.data(), nt);
    eval_type<RT_SURFEL>(P, L, c, f, res.data(), J.data(), nt);
    eval_type<RT_CAM>(P, L, c, f, res.data(), J.data(), nt);
    eval_type<RT_CAMSURF>(P, L, c, f, res.data(), J.data(), nt);
    eval_type<RT_ORIENT>(P, L, c, f, res.data(), J.data(), nt);
    *cost = c;
    std::fill(g, g + nt, 0.0);
    std::fill(H, H + static_cast<size_t>(nt) * nt, 0.0);
    std::vector<int> nz;
    for (int r = 0; r < L.n_res; ++r) {  // rows are sparse (<= 60 non-zeros): outer products over the non-zeros only
      const double* row = J.data() + static_cast<size_t>(r) * nt;
      nz.clear();
      for (int a = 0; a < nt; ++a) if (row[a] != 0.0) nz.push_back(a);
      for (int a : nz) {
        g[a] += row[a] * res[r];
        for (int b : nz) H[static_cast<size_t>(a) * nt + b] += row[a] * row[b];
      }
    }
    return LVI_OK;
  } catch (const RangeError& e) { g_err = e.what(); return LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_INVALID; }
}

// trajectory evaluation through the product's closed-form path: out = p[3] a[3] q[4] w_body[3]
extern "C" int lvi_hostcheck_traj_eval(const lvi_problem_desc* d, double t, double* out) {
  int i0; double u;
  time_to_index(*d, t, i0, u);
  if (i0 < 0 || i0 > d->n_knots - 4) return LVI_ERR_RANGE;
  So3Eval e;
  so3_spline_eval(d->so3_knots + 4 * i0, u, 1.0 / d->dt, true, true, e);
  double Bp[4], Ba[4];
  basis_pos(u, Bp); basis_acc(u, 1.0 / d->dt, Ba);
  V3 p = v3(0, 0, 0), a = v3(0, 0, 0);
  if (d->r3_knots) { p = r3_spline(d->r3_knots + 3 * i0, Bp); a = r3_spline(d->r3_knots + 3 * i0, Ba); }
  const double o[13] = {p.x, p.y, p.z, a.x, a.y, a.z, e.q.x, e.q.y, e.q.z, e.q.w, e.w_body.x, e.w_body.y, e.w_body.z};
  std::memcpy(out, o, sizeof(o));
  return LVI_OK;
}

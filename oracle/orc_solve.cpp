// oracle/orc_solve.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
//
// CPU restatement of the continuous-time least-squares half of the LVI-ExC hot path (SURVEY §8 a-4 … a-13):
//   * cubic B-spline evaluation          K/trajectories/spline_base.h:19-29,153-168,194-222,380-426
//                                        K/trajectories/uniform_r3_spline_trajectory.h:36-103
//                                        K/trajectories/uniform_so3_spline_trajectory.h:46-125
//                                        K/trajectories/split_trajectory.h:41-58,117-123
//   * quaternion log/exp/ang. velocity   K/math/quaternion_math.h:16-95
//   * sensors                            K/sensors/{sensors.h:137-167, imu.h:25,61-101, constant_bias_imu.h:51-119,
//                                        pinhole_camera.h:96-124,217-238}
//   * residual functors                  K/measurements/{gyroscope_measurement.h:36-110, accelerometer_measurement.h:37-114,
//                                        lidar_surfel_point.h:31-215, static_rscamera_measurement.h:16-203,
//                                        camera_surfel_landmark.h:29-255, orientation_measurement.h:30-80}
//   * solver configuration               K/trajectory_estimator.h:38-68
// (K/ = /root/reference/src/lvi_exc/thirdparty/Kontiki/include/kontiki/).
//
// Ceres (<= 2.1; not vendored, only pin in the tree is A-LOAM's docker CERES_VERSION 1.12.0) is restated from its
// published algorithm: DynamicAutoDiffCostFunction = forward-mode Jets in passes of 4
// active scalars, Corrector (Huber: rho''<=0 branch), EigenQuaternionParameterization Plus/Jacobian,
// TrustRegionMinimizer + LevenbergMarquardtStrategy with Jacobi scaling, bounds projection, exact linear solve
// (SPARSE_SCHUR is exact; here: Schur on the inverse depths + band/arrow Cholesky).
// PARITY STATUS: unpinned by upstream (the reference has no tests/golden vectors and Ceres is absent);
// validated by finite differences, spline identities and recovery of known minimisers (tests/test_oracle_*.py).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <map>
#include <set>
#include <vector>

#include "../include/lvi_exc_b200.h"
#include "orc_math.hpp"

namespace orc {

// ---- spline matrices (spline_base.h:19-29) ------------------------------------------------------------------
static const double M_[4][4] = {{1. / 6., 4. / 6., 1. / 6., 0}, {-3. / 6., 0, 3. / 6., 0}, {3. / 6., -6. / 6, 3. / 6., 0}, {-1. / 6., 3. / 6., -3. / 6., 1. / 6.}};
static const double Mc_[4][4] = {{6. / 6., 5. / 6., 1. / 6., 0}, {0. / 6., 3. / 6., 3. / 6., 0}, {0. / 6., -3. / 6., 3. / 6., 0}, {0. / 6., 1. / 6., -2. / 6., 1. / 6.}};

enum { EvalPosition = 1, EvalVelocity = 2, EvalAcceleration = 4, EvalOrientation = 8, EvalAngularVelocity = 16 };

struct Segment { double t0, dt; int n; int first; /* index of first knot in the master spline */ };

// SplineEntity::AddToProblem (spline_base.h:380-426): times -> segments + ordered knot list
static void build_segments(double master_t0, double master_dt, const std::vector<std::pair<double, double>>& times,
                           std::vector<Segment>& segs, std::vector<int>& knots) {
  int cur_start = 0, cur_end = -1;
  for (const auto& tt : times) {
    int i1 = static_cast<int>(std::floor((tt.first - master_t0) / master_dt));
    int i2 = static_cast<int>(std::floor((tt.second - master_t0) / master_dt));
    if (i1 > cur_end) {
      segs.push_back({master_t0 + master_dt * i1, master_dt, 0, i1});
      cur_start = i1;
    } else {
      i1 = cur_end + 1;
    }
    for (int i = i1; i < i2 + 4; ++i) { knots.push_back(i); segs.back().n += 1; }
    cur_end = cur_start + segs.back().n - 1;
  }
}

template <class T> struct Eval { V3<T> p, v, a, w; Quat<T> q; };

template <class T> Quat<T> logq(const Quat<T>& q) {  // quaternion_math.h:16-58
  T qn = qnorm(q);
  if (abs_(qn - 1.0) > 1e-5) throw std::runtime_error("logq: Only implemented for unit quaternions.");
  T k;
  T v2 = dot(q.vec(), q.vec());
  if (v2 > 1e-16) { T vn = sqrt_(v2); k = atan2_(vn, q.w) / vn; } else { k = T(1.0); }
  return Quat<T>(q.x * k, q.y * k, q.z * k, T(0.0));
}
template <class T> Quat<T> expq(const Quat<T>& q) {  // quaternion_math.h:60-85
  T v2 = dot(q.vec(), q.vec());
  T ea = exp_(q.w);
  T ka, kv;
  if (v2 > 1e-16) { T vn = sqrt_(v2); ka = ea * cos_(vn); kv = ea * sin_(vn) / vn; } else { ka = ea; kv = ea; }
  return Quat<T>(kv * q.x, kv * q.y, kv * q.z, ka);
}

// segment-local index (spline_base.h:153-168)
template <class T> static void index_u(const Segment& s, const T& t, int& i0, T& u) {
  T sv = (t - s.t0) / s.dt;
  i0 = static_cast<int>(std::floor(val(sv)));
  u = sv - double(i0);
}

// UniformR3SplineSegmentView::Evaluate (uniform_r3_spline_trajectory.h:36-103). cps: n blocks of 3.
template <class T> static void r3_eval(const Segment& s, T const* const* cps, const T& t, int flags, Eval<T>& r) {
  int i0; T u;
  index_u(s, t, i0, u);
  if (s.n < 4 || i0 < 0 || i0 > s.n - 4) throw std::range_error("r3 spline: t out of range");
  T u2 = u * u, u3 = u2 * u;
  const double di = 1.0 / s.dt;
  T Up[4] = {T(1.0), u, u2, u3};
  T Uv[4] = {T(0.0), T(di), 2.0 * u * di, 3.0 * u2 * di};
  T Ua[4] = {T(0.0), T(0.0), T(2.0 * di * di), 6.0 * u * (di * di)};
  for (int j = 0; j < 4; ++j) {
    T Bp(0.0), Bv(0.0), Ba(0.0);
    for (int k = 0; k < 4; ++k) { Bp += Up[k] * T(M_[k][j]); Bv += Uv[k] * T(M_[k][j]); Ba += Ua[k] * T(M_[k][j]); }
    const T* cp = cps[i0 + j];
    V3<T> c(cp[0], cp[1], cp[2]);
    if (flags & EvalPosition) r.p = r.p + Bp * c;
    if (flags & EvalVelocity) r.v = r.v + Bv * c;
    if (flags & EvalAcceleration) r.a = r.a + Ba * c;
  }
}

// UniformSO3SplineSegmentView::Evaluate (uniform_so3_spline_trajectory.h:46-125). cps: n blocks of 4 (x,y,z,w).
template <class T> static void so3_eval(const Segment& s, T const* const* cps, const T& t, int flags, Eval<T>& r) {
  int i0; T u;
  index_u(s, t, i0, u);
  if (s.n < 4 || i0 < 0 || i0 > s.n - 4) throw std::range_error("so3 spline: t out of range");
  T u2 = u * u, u3 = u2 * u;
  const double di = 1.0 / s.dt;
  T U[4] = {T(1.0), u, u2, u3};
  T dU[4] = {T(0.0), T(di), 2.0 * u * di, 3.0 * u2 * di};
  T B[4], dB[4];
  for (int j = 0; j < 4; ++j) { B[j] = T(0.0); dB[j] = T(0.0); for (int k = 0; k < 4; ++k) { B[j] += U[k] * T(Mc_[k][j]); dB[j] += dU[k] * T(Mc_[k][j]); } }
  auto CP = [&](int i) { const T* c = cps[i]; return Quat<T>(c[0], c[1], c[2], c[3]); };
  Quat<T> q = CP(i0);
  Quat<T> dq_parts[3];
  const bool need_w = flags & EvalAngularVelocity;
  for (int i = i0 + 1; i < i0 + 4; ++i) {
    Quat<T> qa = CP(i - 1), qb = CP(i);
    Quat<T> omega = logq(qa.conj() * qb);
    const T& b = B[i - i0];
    Quat<T> eomegab = expq(Quat<T>(omega.x * b, omega.y * b, omega.z * b, omega.w * b));
    q = q * eomegab;
    if (need_w) {
      for (int j = i0 + 1; j < i0 + 4; ++j) {
        const int m = j - i0 - 1;
        if (i == j) { const T& d = dB[i - i0]; dq_parts[m] = dq_parts[m] * Quat<T>(omega.x * d, omega.y * d, omega.z * d, omega.w * d); }
        dq_parts[m] = dq_parts[m] * eomegab;
      }
    }
  }
  r.q = q;
  if (need_w) {
    Quat<T> sum(dq_parts[0].x + dq_parts[1].x + dq_parts[2].x, dq_parts[0].y + dq_parts[1].y + dq_parts[2].y,
                dq_parts[0].z + dq_parts[1].z + dq_parts[2].z, dq_parts[0].w + dq_parts[1].w + dq_parts[2].w);
    Quat<T> dq = CP(i0) * sum;
    Quat<T> wq = dq * q.conj();  // math::angular_velocity: 2*(dq * conj(q)).vec
    r.w = V3<T>(2.0 * wq.x, 2.0 * wq.y, 2.0 * wq.z);
  }
}

// SplineView::Evaluate (spline_base.h:194-222) incl. the t-1e-5 retry (Q6)
template <class T, class F> static bool seg_dispatch(const std::vector<Segment>& segs, const T& t, F&& f) {
  int off = 0;
  for (const auto& s : segs) {
    const double mn = s.t0, mx = s.t0 + (s.n - 3) * s.dt;
    if (val(t) >= mn && val(t) < mx) { f(s, off, t); return true; }
    T tt = t - 0.00001;
    if (val(tt) >= mn && val(tt) < mx) { f(s, off, tt); return true; }
    off += s.n;
  }
  return false;
}

struct TrajMeta { std::vector<Segment> r3, so3; int n_r3 = 0, n_so3 = 0; bool has_r3 = true; };

// SplitView::Evaluate (split_trajectory.h:41-58); params = [r3 blocks..., so3 blocks...]
template <class T> static Eval<T> traj_eval(const TrajMeta& m, T const* const* params, const T& t, int flags) {
  Eval<T> r;
  if ((flags & (EvalPosition | EvalVelocity | EvalAcceleration)) && m.has_r3) {
    if (!seg_dispatch(m.r3, t, [&](const Segment& s, int off, const T& tt) { r3_eval(s, params + off, tt, flags, r); }))
      throw std::range_error("No segment found for time t");
  }
  if (flags & (EvalOrientation | EvalAngularVelocity)) {
    if (!seg_dispatch(m.so3, t, [&](const Segment& s, int off, const T& tt) { so3_eval(s, params + m.n_r3 + off, tt, flags, r); }))
      throw std::range_error("No segment found for time t");
  }
  return r;
}

// ---- problem -----------------------------------------------------------------------------------------------
enum RType { R_GYRO = 0, R_ACCEL, R_SURFEL, R_CAM, R_CAMSURF, R_ORIENT };

struct Block { double* data; int size; int tsize; bool constant; bool quat; int toff; bool has_lower; double lower; bool used; };

struct RBlock {
  RType type; int idx; int nres; double huber;  // huber <= 0: no loss
  TrajMeta meta; std::vector<int> blocks;       // parameter block ids in the reference's order
  bool active;                                  // false: all blocks constant -> removed, cost goes to fixed_cost
};

struct Problem {
  lvi_problem_desc d;
  std::vector<Block> blocks;
  std::vector<RBlock> rb;
  int n_tangent = 0, n_res = 0;
  int id_r3(int i) const { return i; }
  int id_so3(int i) const { return d.n_knots + i; }
  int base_s, base_plane, base_rho;
  double imu_q[4] = {0, 0, 0, 1}, imu_p[3] = {0, 0, 0}, imu_toff[1] = {0}, lidar_toff[1] = {0}, cam_toff[1] = {0};
  bool constrained = false;
};
// sensor block ids relative to base_s
enum { S_IMU_Q = 0, S_IMU_P, S_IMU_T, S_GR, S_GP, S_BA, S_BG, S_LQ, S_LP, S_LT, S_CQ, S_CP, S_CT, S_COUNT };

static double min_time(const lvi_problem_desc& d) { return d.t0; }
static double max_time(const lvi_problem_desc& d) { return d.t0 + (d.n_knots - 3) * d.dt; }

// TrajectoryEstimator::CheckTimeSpans (trajectory_estimator.h:106-130)
static void check_spans(const lvi_problem_desc& d, const std::vector<std::pair<double, double>>& times) {
  double prev = 0; int i = 0;
  for (auto& ts : times) {
    if (ts.first < min_time(d) || ts.second >= max_time(d)) throw std::range_error("Time span out of range for trajectory");
    if (ts.first > ts.second) throw std::range_error("At least one time span begins before it ends");
    else if (i > 0 && ts.first < prev) throw std::range_error("Time spans are not ordered");
    prev = ts.first; ++i;
  }
}

static void add_traj(Problem& P, RBlock& rb, const std::vector<std::pair<double, double>>& times, bool so3_only) {
  check_spans(P.d, times);
  std::vector<int> k;
  if (!so3_only) {
    build_segments(P.d.t0, P.d.dt, times, rb.meta.r3, k);
    for (int i : k) rb.blocks.push_back(P.id_r3(i));
    rb.meta.n_r3 = static_cast<int>(k.size());
  } else {
    rb.meta.has_r3 = false; rb.meta.n_r3 = 0;
  }
  k.clear();
  build_segments(P.d.t0, P.d.dt, times, rb.meta.so3, k);
  for (int i : k) rb.blocks.push_back(P.id_so3(i));
  rb.meta.n_so3 = static_cast<int>(k.size());
}

static Problem* build_problem(const lvi_problem_desc* dd) {
  Problem* Pp = new Problem();
  Problem& P = *Pp;
  P.d = *dd;
  const lvi_problem_desc& d = P.d;
  const int n = d.n_knots;
  P.lidar_toff[0] = d.lidar_toff; P.cam_toff[0] = d.cam_toff; P.imu_toff[0] = d.imu_toff;
  auto mk = [](double* data, int size, bool quat, bool constant) { Block b{data, size, quat ? 3 : size, constant, quat, -1, false, 0.0, false}; return b; };
  for (int i = 0; i < n; ++i) P.blocks.push_back(mk(d.r3_knots ? d.r3_knots + 3 * i : nullptr, 3, false, d.lock_r3 != 0));
  for (int i = 0; i < n; ++i) P.blocks.push_back(mk(d.so3_knots + 4 * i, 4, true, d.lock_so3 != 0));
  P.base_s = static_cast<int>(P.blocks.size());
  P.blocks.resize(P.base_s + S_COUNT);
  P.blocks[P.base_s + S_IMU_Q] = mk(P.imu_q, 4, true, true);   // IMU q,p,t_off locked by default (sensors.h:95-97, Q5)
  P.blocks[P.base_s + S_IMU_P] = mk(P.imu_p, 3, false, true);
  P.blocks[P.base_s + S_IMU_T] = mk(P.imu_toff, 1, false, true);
  P.blocks[P.base_s + S_GR] = mk(d.gravity, 1, false, false);  // never locked (imu.h:129-142, Q5)
  P.blocks[P.base_s + S_GP] = mk(d.gravity + 1, 1, false, false);
  P.blocks[P.base_s + S_BA] = mk(d.acc_bias, 3, false, d.lock_acc_bias != 0);
  P.blocks[P.base_s + S_BG] = mk(d.gyr_bias, 3, false, d.lock_gyr_bias != 0);
  P.blocks[P.base_s + S_LQ] = mk(d.lidar_q, 4, true, d.lock_lidar_q != 0);
  P.blocks[P.base_s + S_LP] = mk(d.lidar_p, 3, false, d.lock_lidar_p != 0);
  P.blocks[P.base_s + S_LT] = mk(P.lidar_toff, 1, false, true);
  P.blocks[P.base_s + S_CQ] = mk(d.cam_q, 4, true, d.lock_cam_q != 0);
  P.blocks[P.base_s + S_CP] = mk(d.cam_p, 3, false, d.lock_cam_p != 0);
  P.blocks[P.base_s + S_CT] = mk(P.cam_toff, 1, false, true);
  P.base_plane = static_cast<int>(P.blocks.size());
  for (int k = 0; k < d.n_planes; ++k) P.blocks.push_back(mk(const_cast<double*>(d.planes) + 3 * k, 3, false, true));  // LiDARSurfelPoint::locked_ = true
  P.base_rho = static_cast<int>(P.blocks.size());
  for (int l = 0; l < d.n_landmarks; ++l) {
    Block b = mk(d.rho + l, 1, false, d.rho_locked ? d.rho_locked[l] != 0 : false);
    b.has_lower = true; b.lower = 0.0;  // static_rscamera_measurement.h:185
    P.blocks.push_back(b);
  }
  const bool so3_only = (d.r3_knots == nullptr);
  auto imu_blocks = [&](RBlock& rb) { for (int s : {S_IMU_Q, S_IMU_P, S_IMU_T, S_GR, S_GP, S_BA, S_BG}) rb.blocks.push_back(P.base_s + s); };
  // order of tables: gyro, accel, surfel, cam, camsurf, orient (== residual vector order of the C-ABI)
  for (int i = 0; i < d.n_gyro; ++i) {  // gyroscope_measurement.h:81-110
    RBlock rb; rb.type = R_GYRO; rb.idx = i; rb.nres = 3; rb.huber = -1;
    add_traj(P, rb, {{d.gyro_t[i], d.gyro_t[i]}}, so3_only);
    imu_blocks(rb);
    P.rb.push_back(std::move(rb));
  }
  for (int i = 0; i < d.n_accel; ++i) {  // accelerometer_measurement.h:85-114
    RBlock rb; rb.type = R_ACCEL; rb.idx = i; rb.nres = 3; rb.huber = -1;
    add_traj(P, rb, {{d.accel_t[i], d.accel_t[i]}}, so3_only);
    imu_blocks(rb);
    P.rb.push_back(std::move(rb));
  }
  for (int i = 0; i < d.n_surfel; ++i) {  // lidar_surfel_point.h:141-215 (time offset locked)
    RBlock rb; rb.type = R_SURFEL; rb.idx = i; rb.nres = 1; rb.huber = d.surfel_huber[i];
    add_traj(P, rb, {{d.surfel_tmap[i], d.surfel_tmap[i]}, {d.surfel_t[i], d.surfel_t[i]}}, so3_only);
    for (int s : {S_LQ, S_LP, S_LT}) rb.blocks.push_back(P.base_s + s);
    rb.blocks.push_back(P.base_plane + d.surfel_plane[i]);
    P.rb.push_back(std::move(rb));
  }
  for (int i = 0; i < d.n_cam; ++i) {  // static_rscamera_measurement.h:136-203
    RBlock rb; rb.type = R_CAM; rb.idx = i; rb.nres = 2; rb.huber = d.cam_huber[i];
    double t1 = d.cam_t0_ref[i], t2 = d.cam_t0_obs[i];
    if (!(t1 <= t2)) std::swap(t1, t2);
    const double margin = 1e-3;
    add_traj(P, rb, {{t1 - margin, t1 + d.readout + margin}, {t2 - margin, t2 + d.readout + margin}}, so3_only);
    for (int s : {S_CQ, S_CP, S_CT}) rb.blocks.push_back(P.base_s + s);
    rb.blocks.push_back(P.base_rho + d.cam_landmark[i]);
    P.rb.push_back(std::move(rb));
  }
  for (int i = 0; i < d.n_camsurf; ++i) {  // camera_surfel_landmark.h:177-255
    RBlock rb; rb.type = R_CAMSURF; rb.idx = i; rb.nres = 1; rb.huber = d.cs_huber[i];
    add_traj(P, rb, {{d.cs_tmap[i], d.cs_tmap[i]}, {d.cs_t[i], d.cs_t[i]}}, so3_only);
    for (int s : {S_CQ, S_CP, S_CT}) rb.blocks.push_back(P.base_s + s);
    for (int s : {S_LQ, S_LP, S_LT}) rb.blocks.push_back(P.base_s + s);
    rb.blocks.push_back(P.base_plane + d.cs_plane[i]);
    rb.blocks.push_back(P.base_rho + d.cs_landmark[i]);
    P.rb.push_back(std::move(rb));
  }
  for (int i = 0; i < d.n_orient; ++i) {  // orientation_measurement.h:59-80
    RBlock rb; rb.type = R_ORIENT; rb.idx = i; rb.nres = 1; rb.huber = -1;
    add_traj(P, rb, {{d.orient_t[i], d.orient_t[i]}}, true);
    P.rb.push_back(std::move(rb));
  }
  // reduced program: residual blocks whose parameter blocks are all constant are dropped (their cost is fixed)
  for (auto& rb : P.rb) {
    rb.active = false;
    for (int b : rb.blocks) if (!P.blocks[b].constant) rb.active = true;
    if (rb.active) for (int b : rb.blocks) P.blocks[b].used = true;
  }
  // oracle tangent layout: knots interleaved [r3_i, so3_i], then sensors, then rho
  int off = 0;
  auto place = [&](int id) { Block& b = P.blocks[id]; if (b.used && !b.constant) { b.toff = off; off += b.tsize; if (b.has_lower) P.constrained = true; } };
  for (int i = 0; i < n; ++i) { place(P.id_r3(i)); place(P.id_so3(i)); }
  for (int s = 0; s < S_COUNT; ++s) place(P.base_s + s);
  for (int l = 0; l < d.n_landmarks; ++l) place(P.base_rho + l);
  P.n_tangent = off;
  P.n_res = 0;
  for (auto& rb : P.rb) P.n_res += rb.nres;
  return Pp;
}

// ---- residual functors (templated on the scalar like the reference's Residual::operator()) --------------------
template <class T> static V3<T> refined_gravity(const T& roll, const T& pitch) {  // imu.h:61-70, G = -9.79 (imu.h:25)
  const double G = -9.79;
  T cr = cos_(roll), sr = sin_(roll), cp = cos_(pitch), sp = sin_(pitch);
  return V3<T>(-sp * cr * G, sr * G, -cr * cp * G);
}

// ---- pinhole distortion (K/sensors/pinhole_camera.h) -----------------------------------------------------------------------------------------
static bool do_distortion(const lvi_problem_desc& d) {  // :78 (tests p1 twice, never p2 or k3)
  return std::abs(d.distortion[0]) > 1e-5 || std::abs(d.distortion[1]) > 1e-5 || std::abs(d.distortion[2]) > 1e-5 || std::abs(d.distortion[2]) > 1e-5;
}
template <class T> static void distortion(const lvi_problem_desc& d, const T& x, const T& y, T& du, T& dv) {  // :199-215
  const double k1 = d.distortion[0], k2 = d.distortion[1], p1 = d.distortion[2], p2 = d.distortion[3], k3 = d.distortion[4];
  T mx2_u = x * x, my2_u = y * y, mxy_u = x * y;
  T rho2_u = mx2_u + my2_u;
  T rad_dist_u = k1 * rho2_u + k2 * rho2_u * rho2_u + k3 * rho2_u * rho2_u * rho2_u;
  du = x * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho2_u + 2.0 * mx2_u);
  dv = y * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho2_u + 2.0 * my2_u);
}
static void lift_projective(const lvi_problem_desc& d, double u, double v, double out[2]) {  // :131-190, on the (constant) observation
  const double mx_d = (u - d.cx) / d.fx, my_d = (v - d.cy) / d.fy;   // m_inv_K11 * u + m_inv_K13 (:137-138)
  double mx_u = mx_d, my_u = my_d;
  if (do_distortion(d)) {
    const int n = 8;   // recursive distortion model (:168-187), fixed number of steps
    double du, dv;
    distortion<double>(d, mx_d, my_d, du, dv);
    mx_u = mx_d - du; my_u = my_d - dv;
    for (int i = 1; i < n; ++i) {
      distortion<double>(d, mx_u, my_u, du, dv);
      mx_u = mx_d - du; my_u = my_d - dv;
    }
  }
  out[0] = mx_u; out[1] = my_u;
}

template <class T> static bool functor(const Problem& P, const RBlock& rb, T const* const* params, T* res) {
  const lvi_problem_desc& d = P.d;
  const int nt = rb.meta.n_r3 + rb.meta.n_so3;
  switch (rb.type) {
    case R_GYRO: {  // gyroscope_measurement.h:36-38 ; imu.h:87-91 ; constant_bias_imu.h:57-61
      const int i = rb.idx;
      const T* toff = params[nt + S_IMU_T];
      const T* bg = params[nt + 6];
      Eval<T> e = traj_eval<T>(rb.meta, params, T(d.gyro_t[i]) + toff[0], EvalOrientation | EvalAngularVelocity);
      V3<T> m = rot(e.q.conj(), e.w);
      const double w = d.gyro_weight[i];
      for (int k = 0; k < 3; ++k) res[k] = w * (T(d.gyro_w[3 * i + k]) - (m[k] + bg[k]));
      return true;
    }
    case R_ACCEL: {  // accelerometer_measurement.h:37-39 ; imu.h:95-101 ; constant_bias_imu.h:51-55
      const int i = rb.idx;
      const T* toff = params[nt + S_IMU_T];
      const T* gr = params[nt + 3];
      const T* gp = params[nt + 4];
      const T* ba = params[nt + 5];
      Eval<T> e = traj_eval<T>(rb.meta, params, T(d.accel_t[i]) + toff[0], EvalOrientation | EvalAcceleration);
      V3<T> m = rot(e.q.conj(), e.a + refined_gravity(gr[0], gp[0]));
      const double w = d.accel_weight[i];
      for (int k = 0; k < 3; ++k) res[k] = w * (T(d.accel_a[3 * i + k]) - (m[k] + ba[k]));
      return true;
    }
    case R_SURFEL: {  // lidar_surfel_point.h:31-74
      const int i = rb.idx;
      const T* lq = params[nt + 0]; const T* lp = params[nt + 1]; const T* lt = params[nt + 2]; const T* pl = params[nt + 3];
      const int fl = EvalPosition | EvalOrientation;
      Eval<T> e0 = traj_eval<T>(rb.meta, params, T(d.surfel_tmap[i]) + lt[0], fl);
      Eval<T> ek = traj_eval<T>(rb.meta, params, T(d.surfel_t[i]) + lt[0], fl);
      V3<T> p_LinI(lp[0], lp[1], lp[2]);
      Quat<T> q_LtoI(lq[0], lq[1], lq[2], lq[3]);
      V3<T> p_Lk(T(d.surfel_point[3 * i]), T(d.surfel_point[3 * i + 1]), T(d.surfel_point[3 * i + 2]));
      V3<T> p_I = rot(q_LtoI, p_Lk) + p_LinI;
      V3<T> p_temp = rot(e0.q.conj(), rot(ek.q, p_I) + ek.p - e0.p);
      V3<T> p_M = rot(q_LtoI.conj(), p_temp - p_LinI);
      V3<T> Pi(pl[0], pl[1], pl[2]);
      T plane_d = norm(Pi);
      T nrm[3] = {Pi.x / plane_d, Pi.y / plane_d, Pi.z / plane_d};
      T dist = nrm[0] * p_M.x + nrm[1] * p_M.y + nrm[2] * p_M.z - plane_d;
      res[0] = d.surfel_weight[i] * dist;
      return true;
    }
    case R_CAM: {  // static_rscamera_measurement.h:16-60 ; pinhole_camera.h:96-124,217-238
      const int i = rb.idx;
      const T* cq = params[nt + 0]; const T* cp = params[nt + 1]; const T* ct = params[nt + 2];
      T rho = params[nt + 3][0];
      const double row_delta = d.readout / double(d.cam_rows);
      T t_ref = d.cam_t0_ref[i] + ct[0] + d.cam_uv_ref[2 * i + 1] * row_delta;
      T t_obs = d.cam_t0_obs[i] + ct[0] + d.cam_uv_obs[2 * i + 1] * row_delta;
      const int fl = EvalPosition | EvalOrientation;
      Eval<T> er = traj_eval<T>(rb.meta, params, t_ref, fl);
      Eval<T> eo = traj_eval<T>(rb.meta, params, t_obs, fl);
      V3<T> p_CinI(cp[0], cp[1], cp[2]);
      Quat<T> q_CinI(cq[0], cq[1], cq[2], cq[3]);
      V3<T> p_ct = rot(q_CinI.conj(), -p_CinI);
      Quat<T> q_ct = q_CinI.conj();
      // Unproject: K^-1 * (u, v, 1)
      double lift[2];
      lift_projective(d, d.cam_uv_ref[2 * i], d.cam_uv_ref[2 * i + 1], lift);
      V3<T> yh(T(lift[0]), T(lift[1]), T(1.0));
      V3<T> X_ref = rot(q_ct.conj(), yh - rho * p_ct);
      V3<T> X = rot(er.q, X_ref) + er.p * rho;
      V3<T> X_obs = rot(eo.q.conj(), X - rho * eo.p);
      V3<T> Xc = rot(q_ct, X_obs) + p_ct * rho;
      const double eps = 1e-32;  // spaceToPlane
      T pu = Xc.x / (eps + Xc.z), pv = Xc.y / (eps + Xc.z);
      if (do_distortion(d)) {  // p_d = p_u + d_u (pinhole_camera.h:228-233)
        T du, dv;
        distortion<T>(d, pu, pv, du, dv);
        pu = pu + du; pv = pv + dv;
      }
      T yx = d.fx * pu + d.cx, yy = d.fy * pv + d.cy;
      const double w = d.cam_weight[i];
      res[0] = w * (d.cam_uv_obs[2 * i] - yx);
      res[1] = w * (d.cam_uv_obs[2 * i + 1] - yy);
      return true;
    }
    case R_CAMSURF: {  // camera_surfel_landmark.h:29-91 ; rho is read as a constant (Q4, :159-161)
      const int i = rb.idx;
      const T* cq = params[nt + 0]; const T* cp = params[nt + 1]; const T* ct = params[nt + 2];
      const T* lq = params[nt + 3]; const T* lp = params[nt + 4]; const T* pl = params[nt + 6];
      const double rho = d.rho[d.cs_landmark[i]];
      const int fl = EvalPosition | EvalOrientation;
      Eval<T> e0 = traj_eval<T>(rb.meta, params, T(d.cs_tmap[i]) + ct[0], fl);
      Eval<T> ek = traj_eval<T>(rb.meta, params, T(d.cs_t[i]) + ct[0], fl);
      V3<T> p_CinI(cp[0], cp[1], cp[2]), p_LinI(lp[0], lp[1], lp[2]);
      Quat<T> q_CtoI(cq[0], cq[1], cq[2], cq[3]), q_LtoI(lq[0], lq[1], lq[2], lq[3]);
      const double s = 1.0 / (rho + 1e-8);
      double lift[2];
      lift_projective(d, d.cs_uv[2 * i], d.cs_uv[2 * i + 1], lift);
      V3<T> yh(T(lift[0] * s), T(lift[1] * s), T(s));
      V3<T> p_I = rot(q_CtoI, yh) + p_CinI;
      V3<T> p_temp = rot(e0.q.conj(), rot(ek.q, p_I) + ek.p - e0.p);
      V3<T> p_M = rot(q_LtoI.conj(), p_temp - p_LinI);
      V3<T> Pi(pl[0], pl[1], pl[2]);
      T plane_d = norm(Pi);
      T dist = (Pi.x / plane_d) * p_M.x + (Pi.y / plane_d) * p_M.y + (Pi.z / plane_d) * p_M.z - plane_d;
      res[0] = d.cs_weight[i] * dist;
      return true;
    }
    case R_ORIENT: {  // orientation_measurement.h:30-33 ; Eigen angularDistance = 2*atan2(|d.vec|, |d.w|)
      const int i = rb.idx;
      Eval<T> e = traj_eval<T>(rb.meta, params, T(d.orient_t[i]), EvalOrientation);
      Quat<T> qm(T(d.orient_q[4 * i]), T(d.orient_q[4 * i + 1]), T(d.orient_q[4 * i + 2]), T(d.orient_q[4 * i + 3]));
      Quat<T> dd = qm * e.q.conj();
      T ang = 2.0 * atan2_(norm(dd.vec()), abs_(dd.w));
      res[0] = d.orient_weight[i] * ang;
      return true;
    }
  }
  return false;
}

// ---- evaluation: residuals + Jacobian (autodiff stride 4) + corrector + local parameterisation --------------------
struct JRow { int nres; std::vector<int> toff; std::vector<int> tsz; std::vector<double> J; int ncols; };  // J row-major nres x ncols

static void quat_plus_jac(const double* q, double J[12]) {  // EigenQuaternionParameterization::ComputeJacobian (x,y,z,w)
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double j[12] = {w, z, -y, -z, w, x, y, -x, w, -x, -y, -z};
  for (int k = 0; k < 12; ++k) J[k] = j[k];
}
static void quat_plus(const double* q, const double* dl, double* out) {  // EigenQuaternionParameterization::Plus
  const double n = std::sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    Quat<double> dq(s * dl[0], s * dl[1], s * dl[2], std::cos(n));
    Quat<double> r = dq * Quat<double>(q[0], q[1], q[2], q[3]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  } else { for (int k = 0; k < 4; ++k) out[k] = q[k]; }
}

struct EvalOut { double cost = 0, fixed_cost = 0; std::vector<double> res; std::vector<JRow> rows; };

static double huber_rho(double s, double a, double* rho1) {  // ceres::HuberLoss
  const double b = a * a;
  if (s > b) { const double r = std::sqrt(s); *rho1 = std::max(std::numeric_limits<double>::min(), a / r); return 2.0 * a * r - b; }
  *rho1 = 1.0; return s;
}

static void evaluate(const Problem& P, bool want_jac, EvalOut& out) {
  const int nrb = static_cast<int>(P.rb.size());
  out.res.assign(P.n_res, 0.0);
  if (want_jac) out.rows.assign(nrb, JRow());
  std::vector<int> roff(nrb + 1, 0);
  for (int i = 0; i < nrb; ++i) roff[i + 1] = roff[i] + P.rb[i].nres;
  double cost = 0, fixed = 0;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cost, fixed)
  for (int bi = 0; bi < nrb; ++bi) {
    const RBlock& rb = P.rb[bi];
    try {
      const int nb = static_cast<int>(rb.blocks.size());
      std::vector<int> poff(nb + 1, 0);
      for (int k = 0; k < nb; ++k) poff[k + 1] = poff[k] + P.blocks[rb.blocks[k]].size;
      double r[3];
      {  // residual-only evaluation with doubles
        std::vector<const double*> pp(nb);
        for (int k = 0; k < nb; ++k) pp[k] = P.blocks[rb.blocks[k]].data;
        functor<double>(P, rb, pp.data(), r);
      }
      double s = 0;
      for (int k = 0; k < rb.nres; ++k) s += r[k] * r[k];
      double rho1 = 1.0, rho0 = s;
      if (rb.huber > 0) rho0 = huber_rho(s, rb.huber, &rho1);
      if (!rb.active) { fixed += 0.5 * rho0; for (int k = 0; k < rb.nres; ++k) out.res[roff[bi] + k] = r[k] * std::sqrt(rho1); continue; }
      cost += 0.5 * rho0;
      const double sr = std::sqrt(rho1);  // Corrector, rho'' <= 0 branch
      for (int k = 0; k < rb.nres; ++k) out.res[roff[bi] + k] = r[k] * sr;
      if (!want_jac) continue;
      // --- DynamicAutoDiffCostFunction::Evaluate: passes of 4 over the scalars of non-constant blocks
      typedef Jet<4> J4;
      std::vector<J4> pj(poff[nb]);
      std::vector<const J4*> pp(nb);
      std::vector<int> active;  // ambient scalar indices
      for (int k = 0; k < nb; ++k) {
        const Block& b = P.blocks[rb.blocks[k]];
        for (int c = 0; c < b.size; ++c) pj[poff[k] + c] = J4(b.data[c]);
        pp[k] = pj.data() + poff[k];
        if (!b.constant) for (int c = 0; c < b.size; ++c) active.push_back(poff[k] + c);
      }
      std::vector<double> Jamb(static_cast<size_t>(rb.nres) * poff[nb], 0.0);
      for (size_t a0 = 0; a0 < active.size(); a0 += 4) {
        const int cnt = static_cast<int>(std::min<size_t>(4, active.size() - a0));
        for (int c = 0; c < cnt; ++c) pj[active[a0 + c]].v[c] = 1.0;
        J4 rj[3];
        functor<J4>(P, rb, pp.data(), rj);
        for (int c = 0; c < cnt; ++c) {
          for (int k = 0; k < rb.nres; ++k) Jamb[static_cast<size_t>(k) * poff[nb] + active[a0 + c]] = rj[k].v[c];
          pj[active[a0 + c]].v[c] = 0.0;
        }
      }
      // local parameterisation + corrector
      JRow& row = out.rows[bi];
      row.nres = rb.nres; row.ncols = 0;
      for (int k = 0; k < nb; ++k) { const Block& b = P.blocks[rb.blocks[k]]; if (!b.constant) { row.toff.push_back(b.toff); row.tsz.push_back(b.tsize); row.ncols += b.tsize; } }
      row.J.assign(static_cast<size_t>(rb.nres) * row.ncols, 0.0);
      int col = 0;
      for (int k = 0; k < nb; ++k) {
        const Block& b = P.blocks[rb.blocks[k]];
        if (b.constant) continue;
        if (b.quat) {
          double PJ[12]; quat_plus_jac(b.data, PJ);
          for (int rr = 0; rr < rb.nres; ++rr)
            for (int c = 0; c < 3; ++c) {
              double acc = 0;
              for (int a = 0; a < 4; ++a) acc += Jamb[static_cast<size_t>(rr) * poff[nb] + poff[k] + a] * PJ[a * 3 + c];
              row.J[static_cast<size_t>(rr) * row.ncols + col + c] = acc * sr;
            }
        } else {
          for (int rr = 0; rr < rb.nres; ++rr)
            for (int c = 0; c < b.size; ++c) row.J[static_cast<size_t>(rr) * row.ncols + col + c] = Jamb[static_cast<size_t>(rr) * poff[nb] + poff[k] + c] * sr;
        }
        col += b.tsize;
      }
    } catch (const std::range_error&) {
#pragma omp atomic write
      err = LVI_ERR_RANGE;
    } catch (const std::exception&) {
#pragma omp atomic write
      err = LVI_ERR_DOMAIN;
    }
  }
  if (err == LVI_ERR_RANGE) throw std::range_error("time out of range");
  if (err) throw std::runtime_error("evaluation failed");
  out.cost = cost; out.fixed_cost = fixed;
}

// ---- linear algebra: (J^T J + D^2) y = -J^T r  via Schur on rho + band/arrow Cholesky ---------------------------------
struct Layout {
  int nb = 0, nbo = 0, nrho = 0, bw = 0;          // band dims, border dims, #rho, half bandwidth
  std::vector<int> pos;                          // tangent index -> position ( [0,nb) band, nb+k border, -(l+1)-1.. rho )
  std::vector<int> rho_of;                       // tangent index -> rho slot or -1
};

static void make_layout(const Problem& P, Layout& L) {
  const int nt = P.n_tangent;
  L.pos.assign(nt, -1); L.rho_of.assign(nt, -1);
  // border knots: knots touched by the map-time evaluation of surfel / camsurf residuals (SURVEY §5 "arrow")
  std::set<int> border_knots;
  const lvi_problem_desc& d = P.d;
  auto mark = [&](double t) { int i0 = static_cast<int>(std::floor((t - d.t0) / d.dt)); for (int k = i0; k < i0 + 4; ++k) if (k >= 0 && k < d.n_knots) border_knots.insert(k); };
  for (int i = 0; i < d.n_surfel; ++i) mark(d.surfel_tmap[i] + d.lidar_toff);
  for (int i = 0; i < d.n_camsurf; ++i) mark(d.cs_tmap[i] + d.cam_toff);
  if (border_knots.size() > 16) border_knots.clear();  // not an arrow structure; treat as band
  int pb = 0;
  std::vector<int> border_t;
  for (int i = 0; i < d.n_knots; ++i)
    for (int id : {P.id_r3(i), P.id_so3(i)}) {
      const Block& b = P.blocks[id];
      if (b.toff < 0) continue;
      if (border_knots.count(i)) { for (int c = 0; c < b.tsize; ++c) border_t.push_back(b.toff + c); }
      else { for (int c = 0; c < b.tsize; ++c) L.pos[b.toff + c] = pb++; }
    }
  L.nb = pb;
  for (int s = 0; s < S_COUNT; ++s) { const Block& b = P.blocks[P.base_s + s]; if (b.toff >= 0) for (int c = 0; c < b.tsize; ++c) border_t.push_back(b.toff + c); }
  L.nbo = static_cast<int>(border_t.size());
  for (int k = 0; k < L.nbo; ++k) L.pos[border_t[k]] = L.nb + k;
  int nr = 0;
  for (int l = 0; l < d.n_landmarks; ++l) { const Block& b = P.blocks[P.base_rho + l]; if (b.toff >= 0) L.rho_of[b.toff] = nr++; }
  L.nrho = nr;
}

struct Normal {  // reduced normal equations in band/arrow storage
  int nb, nbo, bw;
  std::vector<double> B;   // nb x (bw+1) lower band, B[i*(bw+1) + (j - i + bw)] for i-bw <= j <= i
  std::vector<double> F;   // nbo x nb
  std::vector<double> C;   // nbo x nbo
  double& band(int i, int j) { return B[static_cast<size_t>(i) * (bw + 1) + (j - i + bw)]; }
};

// Solves (H + diag(D2)) y = -g for y given per-residual Jacobian rows (already Jacobi-scaled via `scale`).
// Returns false on Cholesky breakdown.
static bool solve_normal(const Problem& P, const Layout& L, const EvalOut& ev, const std::vector<double>& scale,
                         const std::vector<double>& D2, const std::vector<double>& g, std::vector<double>& y, int* bw_out) {
  const int nt = P.n_tangent, nb = L.nb, nbo = L.nbo, nr = L.nrho;
  // --- rho blocks: H_rr (diag), and per-rho sparse row H_r,x
  std::vector<double> Hrr(nr, 0.0), grho(nr, 0.0);
  std::vector<std::map<int, double>> Hrx(nr);  // position -> value
  // pass 1: bandwidth
  int bw = 0;
  std::vector<std::pair<int, int>> rho_span(nr, {std::numeric_limits<int>::max(), -1});
  for (size_t bi = 0; bi < ev.rows.size(); ++bi) {
    const JRow& row = ev.rows[bi];
    if (row.ncols == 0) continue;
    int lo = std::numeric_limits<int>::max(), hi = -1, rs = -1;
    for (size_t k = 0; k < row.toff.size(); ++k)
      for (int c = 0; c < row.tsz[k]; ++c) {
        const int t = row.toff[k] + c;
        if (L.rho_of[t] >= 0) { rs = L.rho_of[t]; continue; }
        const int p = L.pos[t];
        if (p < nb) { lo = std::min(lo, p); hi = std::max(hi, p); }
      }
    if (hi >= 0) bw = std::max(bw, hi - lo);
    if (rs >= 0 && hi >= 0) { rho_span[rs].first = std::min(rho_span[rs].first, lo); rho_span[rs].second = std::max(rho_span[rs].second, hi); }
  }
  for (int r = 0; r < nr; ++r) if (rho_span[r].second >= 0) bw = std::max(bw, rho_span[r].second - rho_span[r].first);
  if (bw_out) *bw_out = bw;
  Normal N; N.nb = nb; N.nbo = nbo; N.bw = bw;
  N.B.assign(static_cast<size_t>(nb) * (bw + 1), 0.0); N.F.assign(static_cast<size_t>(nbo) * nb, 0.0); N.C.assign(static_cast<size_t>(nbo) * nbo, 0.0);
  auto addH = [&](int pa, int pb_, double v) {  // pa >= pb_ positions
    if (pa < nb) N.band(pa, pb_) += v;
    else if (pb_ < nb) N.F[static_cast<size_t>(pa - nb) * nb + pb_] += v;
    else N.C[static_cast<size_t>(pa - nb) * nbo + (pb_ - nb)] += v;
  };
  std::vector<int> cols; std::vector<double> jv;
  for (size_t bi = 0; bi < ev.rows.size(); ++bi) {
    const JRow& row = ev.rows[bi];
    if (row.ncols == 0) continue;
    cols.clear();
    for (size_t k = 0; k < row.toff.size(); ++k) for (int c = 0; c < row.tsz[k]; ++c) cols.push_back(row.toff[k] + c);
    for (int rr = 0; rr < row.nres; ++rr) {
      jv.resize(cols.size());
      for (size_t a = 0; a < cols.size(); ++a) jv[a] = row.J[static_cast<size_t>(rr) * row.ncols + a] * scale[cols[a]];
      for (size_t a = 0; a < cols.size(); ++a) {
        if (jv[a] == 0.0) continue;
        const int ta = cols[a]; const int ra = L.rho_of[ta];
        for (size_t b = 0; b <= a; ++b) {
          if (jv[b] == 0.0) continue;
          const int tb = cols[b]; const int rbb = L.rho_of[tb];
          const double v = jv[a] * jv[b];
          if (ra >= 0 && rbb >= 0) { if (ra == rbb) Hrr[ra] += v; }
          else if (ra >= 0) Hrx[ra][L.pos[tb]] += v;
          else if (rbb >= 0) Hrx[rbb][L.pos[ta]] += v;
          else {
            const int pa = L.pos[ta], pb_ = L.pos[tb];
            if (pa == pb_) addH(pa, pb_, v);
            else if (pa > pb_) addH(pa, pb_, v);
            else addH(pb_, pa, v);
            // duplicate column indices inside one row (a != b, same position) contribute twice to the diagonal
            if (a != b && pa == pb_) addH(pa, pb_, v);
          }
        }
      }
    }
  }
  // rhs = -g (scaled gradient), damping
  std::vector<double> rhs(nb + nbo, 0.0);
  for (int t = 0; t < nt; ++t) {
    if (L.rho_of[t] >= 0) { grho[L.rho_of[t]] = -g[t]; Hrr[L.rho_of[t]] += D2[t]; }
    else { const int p = L.pos[t]; rhs[p] = -g[t]; if (p < nb) N.band(p, p) += D2[t]; else N.C[static_cast<size_t>(p - nb) * nbo + (p - nb)] += D2[t]; }
  }
  // Schur complement on rho:  H_xx -= H_xr Hrr^-1 H_rx ; rhs_x -= H_xr Hrr^-1 g_r
  for (int r = 0; r < nr; ++r) {
    if (Hrr[r] <= 0) return false;
    const double inv = 1.0 / Hrr[r];
    std::vector<std::pair<int, double>> e(Hrx[r].begin(), Hrx[r].end());
    for (size_t a = 0; a < e.size(); ++a) {
      rhs[e[a].first] -= e[a].second * inv * grho[r];
      for (size_t b = 0; b <= a; ++b) {
        const int pa = std::max(e[a].first, e[b].first), pb_ = std::min(e[a].first, e[b].first);
        addH(pa, pb_, -e[a].second * inv * e[b].second);
      }
    }
  }
  // band Cholesky (in place, lower), right-looking, rows of the trailing window in parallel
  const int W1 = bw + 1;
  for (int j = 0; j < nb; ++j) {
    double djj = N.band(j, j);
    if (!(djj > 0.0)) return false;
    djj = std::sqrt(djj);
    N.band(j, j) = djj;
    const int iend = std::min(nb - 1, j + bw);
    for (int i = j + 1; i <= iend; ++i) N.band(i, j) /= djj;
    const int cnt = iend - j;
#pragma omp parallel for if (cnt > 64) schedule(static)
    for (int i = j + 1; i <= iend; ++i) {
      const double lij = N.B[static_cast<size_t>(i) * W1 + (j - i + bw)];
      if (lij == 0.0) continue;
      double* rowi = &N.B[static_cast<size_t>(i) * W1 + bw - i];  // rowi[k] = band(i,k)
      for (int k = j + 1; k <= i; ++k) rowi[k] -= lij * N.B[static_cast<size_t>(k) * W1 + (j - k + bw)];
    }
  }
  // Y = L^-1 F^T  (stored as F rows: for each border row solve L y = f)
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nbo; ++r) {
    double* f = &N.F[static_cast<size_t>(r) * nb];
    for (int i = 0; i < nb; ++i) {
      double s = f[i];
      const int k0 = std::max(0, i - bw);
      const double* rowi = &N.B[static_cast<size_t>(i) * W1 + bw - i];
      for (int k = k0; k < i; ++k) s -= rowi[k] * f[k];
      f[i] = s / rowi[i];
    }
  }
  // z1 = L^-1 rhs1
  std::vector<double> z(rhs.begin(), rhs.begin() + nb);
  for (int i = 0; i < nb; ++i) {
    double s = z[i];
    const int k0 = std::max(0, i - bw);
    const double* rowi = &N.B[static_cast<size_t>(i) * W1 + bw - i];
    for (int k = k0; k < i; ++k) s -= rowi[k] * z[k];
    z[i] = s / rowi[i];
  }
  // S = C - Y^T Y ; r2 = rhs2 - Y^T z
  std::vector<double> S(static_cast<size_t>(nbo) * nbo, 0.0), r2(nbo, 0.0);
  for (int a = 0; a < nbo; ++a) {
    const double* fa = &N.F[static_cast<size_t>(a) * nb];
    for (int b = 0; b <= a; ++b) {
      const double* fb = &N.F[static_cast<size_t>(b) * nb];
      double s = 0; for (int i = 0; i < nb; ++i) s += fa[i] * fb[i];
      S[static_cast<size_t>(a) * nbo + b] = N.C[static_cast<size_t>(a) * nbo + b] - s;
    }
    double s = 0; for (int i = 0; i < nb; ++i) s += fa[i] * z[i];
    r2[a] = rhs[nb + a] - s;
  }
  // dense Cholesky of S (lower)
  for (int j = 0; j < nbo; ++j) {
    double dj = S[static_cast<size_t>(j) * nbo + j];
    for (int k = 0; k < j; ++k) dj -= S[static_cast<size_t>(j) * nbo + k] * S[static_cast<size_t>(j) * nbo + k];
    if (!(dj > 0.0)) return false;
    dj = std::sqrt(dj); S[static_cast<size_t>(j) * nbo + j] = dj;
    for (int i = j + 1; i < nbo; ++i) {
      double s = S[static_cast<size_t>(i) * nbo + j];
      for (int k = 0; k < j; ++k) s -= S[static_cast<size_t>(i) * nbo + k] * S[static_cast<size_t>(j) * nbo + k];
      S[static_cast<size_t>(i) * nbo + j] = s / dj;
    }
  }
  std::vector<double> x2(r2);
  for (int i = 0; i < nbo; ++i) { double s = x2[i]; for (int k = 0; k < i; ++k) s -= S[static_cast<size_t>(i) * nbo + k] * x2[k]; x2[i] = s / S[static_cast<size_t>(i) * nbo + i]; }
  for (int i = nbo - 1; i >= 0; --i) { double s = x2[i]; for (int k = i + 1; k < nbo; ++k) s -= S[static_cast<size_t>(k) * nbo + i] * x2[k]; x2[i] = s / S[static_cast<size_t>(i) * nbo + i]; }
  // x1 = L^-T (z - Y x2)
  for (int i = 0; i < nb; ++i) { double s = 0; for (int a = 0; a < nbo; ++a) s += N.F[static_cast<size_t>(a) * nb + i] * x2[a]; z[i] -= s; }
  for (int i = nb - 1; i >= 0; --i) {
    double s = z[i];
    const int k1 = std::min(nb - 1, i + bw);
    for (int k = i + 1; k <= k1; ++k) s -= N.B[static_cast<size_t>(k) * W1 + (i - k + bw)] * z[k];
    z[i] = s / N.B[static_cast<size_t>(i) * W1 + bw];
  }
  // back-substitute rho: y_r = (g_r - H_rx x) / Hrr
  std::vector<double> xs(nb + nbo);
  for (int i = 0; i < nb; ++i) xs[i] = z[i];
  for (int a = 0; a < nbo; ++a) xs[nb + a] = x2[a];
  y.assign(nt, 0.0);
  for (int t = 0; t < nt; ++t) {
    const int r = L.rho_of[t];
    if (r < 0) { y[t] = xs[L.pos[t]]; continue; }
    double s = grho[r];
    for (const auto& kv : Hrx[r]) s -= kv.second * xs[kv.first];
    y[t] = s / Hrr[r];
  }
  return true;
}

// ---- state handling ------------------------------------------------------------------------------------------
struct State { std::vector<double> x; };  // ambient values of all non-constant used blocks, in toff order of blocks
static std::vector<int> free_blocks(const Problem& P) {
  std::vector<int> ids;
  for (size_t i = 0; i < P.blocks.size(); ++i) if (P.blocks[i].toff >= 0) ids.push_back(static_cast<int>(i));
  std::sort(ids.begin(), ids.end(), [&](int a, int b) { return P.blocks[a].toff < P.blocks[b].toff; });
  return ids;
}
static void get_state(const Problem& P, const std::vector<int>& ids, std::vector<double>& x) {
  x.clear();
  for (int id : ids) for (int c = 0; c < P.blocks[id].size; ++c) x.push_back(P.blocks[id].data[c]);
}
static void set_state(Problem& P, const std::vector<int>& ids, const std::vector<double>& x) {
  size_t o = 0;
  for (int id : ids) for (int c = 0; c < P.blocks[id].size; ++c) P.blocks[id].data[c] = x[o++];
}
// Program::Plus: local parameterisation then projection onto bounds
static void plus(const Problem& P, const std::vector<int>& ids, const std::vector<double>& x, const std::vector<double>& delta, std::vector<double>& xp) {
  xp.resize(x.size());
  size_t o = 0;
  for (int id : ids) {
    const Block& b = P.blocks[id];
    if (b.quat) quat_plus(&x[o], &delta[b.toff], &xp[o]);
    else for (int c = 0; c < b.size; ++c) xp[o + c] = x[o + c] + delta[b.toff + c];
    if (b.has_lower) for (int c = 0; c < b.size; ++c) xp[o + c] = std::max(xp[o + c], b.lower);
    o += b.size;
  }
}
static double norm2(const std::vector<double>& a) { double s = 0; for (double v : a) s += v * v; return std::sqrt(s); }

static void gradient_of(const Problem& P, const EvalOut& ev, std::vector<double>& g) {
  g.assign(P.n_tangent, 0.0);
  int ro = 0;
  for (size_t bi = 0; bi < ev.rows.size(); ++bi) {
    const JRow& row = ev.rows[bi];
    const int nres = P.rb[bi].nres;
    if (row.ncols) {
      int col = 0;
      for (size_t k = 0; k < row.toff.size(); ++k) {
        for (int c = 0; c < row.tsz[k]; ++c)
          for (int rr = 0; rr < row.nres; ++rr) g[row.toff[k] + c] += row.J[static_cast<size_t>(rr) * row.ncols + col + c] * ev.res[ro + rr];
        col += row.tsz[k];
      }
    }
    ro += nres;
  }
}

static int solve(Problem& P, const lvi_solve_options& o, lvi_solve_summary& S) {
  const auto T0 = std::chrono::steady_clock::now();
  std::memset(&S, 0, sizeof(S));
  Layout L; make_layout(P, L);
  const std::vector<int> ids = free_blocks(P);
  const int nt = P.n_tangent;
  std::vector<double> x, cand, g, y, delta(nt), scale(nt, 1.0), diag(nt), D2(nt);
  get_state(P, ids, x);
  EvalOut ev;
  double tj = 0, tl = 0;
  auto tic = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
  auto t1 = tic();
  evaluate(P, true, ev);
  tj += ms(t1);
  gradient_of(P, ev, g);
  double x_cost = ev.cost;
  S.initial_cost = x_cost + ev.fixed_cost; S.fixed_cost = ev.fixed_cost;
  S.num_residual_blocks = static_cast<int>(P.rb.size()); S.num_residuals = P.n_res; S.num_effective_parameters = nt;
  // Jacobi scaling at iteration 0: 1/(1+||J_col||)
  auto colnorm2 = [&](std::vector<double>& cn) {
    cn.assign(nt, 0.0);
    for (const JRow& row : ev.rows) {
      int col = 0;
      for (size_t k = 0; k < row.toff.size(); ++k) { for (int c = 0; c < row.tsz[k]; ++c) for (int rr = 0; rr < row.nres; ++rr) { const double v = row.J[static_cast<size_t>(rr) * row.ncols + col + c]; cn[row.toff[k] + c] += v * v; } col += row.tsz[k]; }
    }
  };
  std::vector<double> cn;
  if (o.jacobi_scaling) { colnorm2(cn); for (int t = 0; t < nt; ++t) scale[t] = 1.0 / (1.0 + std::sqrt(cn[t])); }
  double x_norm = norm2(x);
  auto grad_max_norm = [&]() { std::vector<double> mg(nt), xp; for (int t = 0; t < nt; ++t) mg[t] = -g[t]; plus(P, ids, x, mg, xp); double m = 0; for (size_t i = 0; i < x.size(); ++i) m = std::max(m, std::fabs(x[i] - xp[i])); return m; };
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int it = 0, invalid = 0;
  auto log_iter = [&](double cost, double change, double gmax, double step, bool ok) {
    if (S.n_log < LVI_MAX_ITER_LOG) { const int k = S.n_log++; S.log_cost[k] = cost; S.log_cost_change[k] = change; S.log_gradient_max_norm[k] = gmax; S.log_step_norm[k] = step; S.log_radius[k] = radius; S.log_successful[k] = ok; }
  };
  double gmax = grad_max_norm();
  log_iter(x_cost + ev.fixed_cost, 0, gmax, 0, true);
  S.termination_type = LVI_NO_CONVERGENCE;
  if (o.verbose) std::printf("[oracle] iter %3d cost %.9e |g| %.3e\n", 0, x_cost, gmax);
  bool done = gmax <= o.gradient_tolerance;
  if (done) S.termination_type = LVI_CONVERGENCE;
  while (!done) {
    if (it >= o.max_num_iterations) break;
    ++it;
    // --- LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      colnorm2(cn);
      for (int t = 0; t < nt; ++t) diag[t] = std::min(std::max(cn[t] * scale[t] * scale[t], o.min_lm_diagonal), o.max_lm_diagonal);
    }
    for (int t = 0; t < nt; ++t) D2[t] = diag[t] / radius;
    std::vector<double> gs(nt);
    for (int t = 0; t < nt; ++t) gs[t] = g[t] * scale[t];
    t1 = tic();
    int bw = 0;
    const bool ok = solve_normal(P, L, ev, scale, D2, gs, y, &bw);
    tl += ms(t1);
    S.band_width = bw; S.border_width = L.nbo;
    reuse_diagonal = true;
    bool step_valid = ok;
    double model_cost_change = 0;
    if (ok) {
      for (double v : y) if (!std::isfinite(v)) step_valid = false;
    }
    if (step_valid) {
      // model_cost_change = -(J y).(r + J y / 2)   with the scaled Jacobian
      int ro = 0;
      for (size_t bi = 0; bi < ev.rows.size(); ++bi) {
        const JRow& row = ev.rows[bi];
        for (int rr = 0; rr < row.nres && row.ncols; ++rr) {
          double jy = 0; int col = 0;
          for (size_t k = 0; k < row.toff.size(); ++k) { for (int c = 0; c < row.tsz[k]; ++c) jy += row.J[static_cast<size_t>(rr) * row.ncols + col + c] * scale[row.toff[k] + c] * y[row.toff[k] + c]; col += row.tsz[k]; }
          model_cost_change -= jy * (ev.res[ro + rr] + jy / 2.0);
        }
        ro += P.rb[bi].nres;
      }
      if (!(model_cost_change > 0.0)) step_valid = false;
    }
    if (!step_valid) {  // HandleInvalidStep
      ++invalid;
      if (invalid >= o.max_num_consecutive_invalid_steps) { S.termination_type = LVI_FAILURE; break; }
      radius *= 0.5; reuse_diagonal = true;
      ++S.num_unsuccessful_steps;
      log_iter(x_cost + ev.fixed_cost, 0, gmax, 0, false);
      continue;
    }
    invalid = 0;
    for (int t = 0; t < nt; ++t) delta[t] = y[t] * scale[t];
    plus(P, ids, x, delta, cand);
    set_state(P, ids, cand);
    EvalOut evc;
    t1 = tic();
    evaluate(P, false, evc);
    tj += ms(t1);
    double cand_cost = evc.cost;
    if (P.constrained) {
      // DoLineSearch: Armijo along delta with projection; step 1 accepted when sufficient decrease holds.
      double gd = 0; for (int t = 0; t < nt; ++t) gd += g[t] * delta[t];
      double step = 1.0; int ls = 0;
      while (cand_cost > x_cost + 1e-4 * step * gd && ls < 20 && step > 1e-9) {
        // quadratic interpolation through f(0), f'(0), f(step), contraction clamped to [1e-3, 0.6]
        double ns = -gd * step * step / (2.0 * (cand_cost - x_cost - gd * step));
        ns = std::min(std::max(ns, 1e-3 * step), 0.6 * step);
        step = ns; ++ls;
        std::vector<double> ds(nt); for (int t = 0; t < nt; ++t) ds[t] = delta[t] * step;
        plus(P, ids, x, ds, cand); set_state(P, ids, cand); evaluate(P, false, evc); cand_cost = evc.cost;
      }
      if (step != 1.0) { for (int t = 0; t < nt; ++t) delta[t] *= step; }
    }
    // ParameterToleranceReached
    double sn = 0; for (size_t i = 0; i < x.size(); ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]);
    const double step_norm = std::sqrt(sn);
    const double cost_change = x_cost - cand_cost;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      set_state(P, ids, x); S.termination_type = LVI_CONVERGENCE; log_iter(x_cost + ev.fixed_cost, cost_change, gmax, step_norm, false); break; }
    // FunctionToleranceReached
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) {
      set_state(P, ids, x); S.termination_type = LVI_CONVERGENCE; log_iter(x_cost + ev.fixed_cost, cost_change, gmax, step_norm, false); break; }
    const double rel = cost_change / model_cost_change;
    if (rel > o.min_relative_decrease) {  // HandleSuccessfulStep
      x = cand; x_norm = norm2(x); x_cost = cand_cost;
      t1 = tic();
      evaluate(P, true, ev);
      tj += ms(t1);
      gradient_of(P, ev, g);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      ++S.num_successful_steps;
      gmax = grad_max_norm();
      log_iter(x_cost + ev.fixed_cost, cost_change, gmax, step_norm, true);
      if (o.verbose) std::printf("[oracle] iter %3d cost %.9e change %.3e |g| %.3e |step| %.3e radius %.3e\n", it, x_cost, cost_change, gmax, step_norm, radius);
      if (gmax <= o.gradient_tolerance) { S.termination_type = LVI_CONVERGENCE; break; }
    } else {  // HandleUnsuccessfulStep
      set_state(P, ids, x);
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      ++S.num_unsuccessful_steps;
      log_iter(x_cost + ev.fixed_cost, cost_change, gmax, step_norm, false);
      if (o.verbose) std::printf("[oracle] iter %3d REJECTED cost %.9e change %.3e radius %.3e\n", it, cand_cost, cost_change, radius);
    }
    if (radius <= o.min_trust_region_radius) { S.termination_type = LVI_CONVERGENCE; break; }
  }
  set_state(P, ids, x);
  S.num_iterations = it;
  S.final_cost = x_cost + ev.fixed_cost;
  S.time_jacobian_ms = tj; S.time_linear_solve_ms = tl;
  S.time_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
  return LVI_OK;
}

}  // namespace orc

extern "C" {

void orc_solve_options_default(lvi_solve_options* o) {
  o->max_num_iterations = 30; o->verbose = 0;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
}

static thread_local std::string g_err;
const char* orc_last_error() { return g_err.c_str(); }

void* orc_problem_create(const lvi_problem_desc* d, int* status) {
  try { *status = 0; return orc::build_problem(d); }
  catch (const std::range_error& e) { g_err = e.what(); *status = LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); *status = LVI_ERR_INVALID; }
  return nullptr;
}
void orc_problem_free(void* p) { delete static_cast<orc::Problem*>(p); }
int orc_problem_num_residuals(void* p) { return static_cast<orc::Problem*>(p)->n_res; }
int orc_problem_num_tangent(void* p) { return static_cast<orc::Problem*>(p)->n_tangent; }
// which: 0 r3 knot, 1 so3 knot
int orc_problem_tangent_offset_knot(void* p, int knot, int which) {
  orc::Problem* P = static_cast<orc::Problem*>(p);
  return P->blocks[which ? P->id_so3(knot) : P->id_r3(knot)].toff;
}
// which: 0 lidar_q 1 lidar_p 2 cam_q 3 cam_p 4 gravity(roll; pitch = +1) 5 acc_bias 6 gyr_bias ; 7+l rho_l
int orc_problem_tangent_offset_block(void* p, int which) {
  orc::Problem* P = static_cast<orc::Problem*>(p);
  using namespace orc;
  static const int map[7] = {S_LQ, S_LP, S_CQ, S_CP, S_GR, S_BA, S_BG};
  if (which < 7) return P->blocks[P->base_s + map[which]].toff;
  return P->blocks[P->base_rho + (which - 7)].toff;
}
// cost excludes fixed cost; residuals after loss correction (table order); gradient/J in the oracle tangent layout
int orc_problem_evaluate(void* p, double* cost, double* fixed_cost, double* residuals, double* gradient, double* J_dense) {
  orc::Problem* P = static_cast<orc::Problem*>(p);
  try {
    orc::EvalOut ev;
    orc::evaluate(*P, gradient || J_dense, ev);
    if (cost) *cost = ev.cost;
    if (fixed_cost) *fixed_cost = ev.fixed_cost;
    if (residuals) std::copy(ev.res.begin(), ev.res.end(), residuals);
    if (gradient) { std::vector<double> g; orc::gradient_of(*P, ev, g); std::copy(g.begin(), g.end(), gradient); }
    if (J_dense) {
      const int nt = P->n_tangent;
      std::fill(J_dense, J_dense + static_cast<size_t>(P->n_res) * nt, 0.0);
      int ro = 0;
      for (size_t bi = 0; bi < ev.rows.size(); ++bi) {
        const orc::JRow& row = ev.rows[bi];
        if (row.ncols) {
          int col = 0;
          for (size_t k = 0; k < row.toff.size(); ++k) {
            for (int c = 0; c < row.tsz[k]; ++c)
              for (int rr = 0; rr < row.nres; ++rr) J_dense[static_cast<size_t>(ro + rr) * nt + row.toff[k] + c] += row.J[static_cast<size_t>(rr) * row.ncols + col + c];
            col += row.tsz[k];
          }
        }
        ro += P->rb[bi].nres;
      }
    }
    return LVI_OK;
  } catch (const std::range_error& e) { g_err = e.what(); return LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_DOMAIN; }
}
int orc_problem_solve(void* p, const lvi_solve_options* o, lvi_solve_summary* s) {
  try { return orc::solve(*static_cast<orc::Problem*>(p), *o, *s); }
  catch (const std::range_error& e) { g_err = e.what(); return LVI_ERR_RANGE; }
  catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_DOMAIN; }
}
// the owning SplitTrajectory as one segment per spline holding all n knots (what TrajectoryManagerLVI::evaluate*Pose evaluates)
static void master_meta(const lvi_problem_desc& d, orc::TrajMeta& m, std::vector<const double*>& pp) {
  m.r3.push_back({d.t0, d.dt, d.n_knots, 0}); m.n_r3 = d.n_knots;
  m.so3.push_back({d.t0, d.dt, d.n_knots, 0}); m.n_so3 = d.n_knots;
  for (int i = 0; i < d.n_knots; ++i) pp.push_back(d.r3_knots + 3 * i);
  for (int i = 0; i < d.n_knots; ++i) pp.push_back(d.so3_knots + 4 * i);
}
// trajectory evaluation for tests: out = p[3] v[3] a[3] q[4] w[3]
int orc_traj_eval(const lvi_problem_desc* d, double t, double* out) {
  using namespace orc;
  try {
    // Trajectory::Evaluate on the OWNING spline: one segment holding every knot (K/trajectories/spline_base.h:194-222,370-378)
    TrajMeta m; std::vector<const double*> pp;
    master_meta(*d, m, pp);
    Eval<double> e = traj_eval<double>(m, pp.data(), t, 31);
    const double o[16] = {e.p.x, e.p.y, e.p.z, e.v.x, e.v.y, e.v.z, e.a.x, e.a.y, e.a.z, e.q.x, e.q.y, e.q.z, e.q.w, e.w.x, e.w.y, e.w.z};
    std::copy(o, o + 16, out);
    return LVI_OK;
  } catch (const std::exception& e) { g_err = e.what(); return LVI_ERR_RANGE; }
}
int orc_num_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to every rank: callers that time the oracle set the thread count explicitly
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// ScanUndistortion::undistort (L/include/core/scan_undistortion.h:132-180) + evaluateLidarPose
// (L/src/core/trajectory_manager_lvi.cpp:398-408) for n_scans x pts_per_scan raw points; target_time[s] is the pose the
// scan is expressed in (its own stamp for undistortScan, the map time for undistortScanInMap).
// raw: licalib::PointXYZIT (32 B); out: pcl::PointXYZI-shaped 8 floats per point.
int orc_undistort(const lvi_problem_desc* d, const void* raw_v, int32_t n_scans, int64_t pts_per_scan, const double* target_time,
                  int correct_position, float* out) {
  using namespace orc;
  struct Raw { float x, y, z, pad; float intensity; float pad2; double timestamp; };
  const Raw* raw = static_cast<const Raw*>(raw_v);
  TrajMeta mm; std::vector<const double*> mpp;
  master_meta(*d, mm, mpp);
  auto lidar_pose = [&](double t, Quat<double>& q, V3<double>& p) -> bool {
    const double tt = t + d->lidar_toff;
    if (d->t0 > tt || d->t0 + (d->n_knots - 3) * d->dt <= tt) return false;
    Eval<double> e = traj_eval<double>(mm, mpp.data(), tt, EvalOrientation | EvalPosition);  // traj_->Evaluate on the whole spline
    Quat<double> qL(d->lidar_q[0], d->lidar_q[1], d->lidar_q[2], d->lidar_q[3]);
    q = e.q * qL;
    p = rot(e.q, V3<double>(d->lidar_p[0], d->lidar_p[1], d->lidar_p[2])) + e.p;
    return true;
  };
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
  for (int s = 0; s < n_scans; ++s) {
    Quat<double> q0; V3<double> p0;
    bool ok0 = false;
    try { ok0 = lidar_pose(target_time[s], q0, p0); } catch (const std::exception&) { ok0 = false; }
    for (int64_t i = 0; i < pts_per_scan; ++i) {
      const Raw& r = raw[s * pts_per_scan + i];
      float* o = out + (s * pts_per_scan + i) * 8;
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
      o[3] = 1.f;
      const float nanv = std::numeric_limits<float>::quiet_NaN();
      if (!ok0 || std::isnan(r.x)) { o[0] = o[1] = o[2] = nanv; if (!ok0) ++bad; continue; }
      Quat<double> qk; V3<double> pk;
      bool okk = false;
      try { okk = lidar_pose(r.timestamp, qk, pk); } catch (const std::exception&) { ++bad; }
      if (!okk) { continue; }  // reference leaves the default-constructed point (zeros)
      Quat<double> q = q0.conj() * qk;
      V3<double> po = rot(q, V3<double>(r.x, r.y, r.z));
      if (correct_position) po = po + rot(q0.conj(), pk - p0);
      o[0] = static_cast<float>(po.x); o[1] = static_cast<float>(po.y); o[2] = static_cast<float>(po.z);
      o[4] = r.intensity;
    }
  }
  return bad ? LVI_ERR_RANGE : LVI_OK;
}

}  // extern "C"

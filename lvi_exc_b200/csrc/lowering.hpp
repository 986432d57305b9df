// lowering.hpp — host-side lowering of an `lvi_problem_desc` (what the Kontiki-shaped facade records) into the
// flat tables the kernels consume.  Replaces the per-residual pointer bookkeeping of
//   TrajectoryEstimator::AddTrajectoryForTimes / CheckTimeSpans   K/trajectory_estimator.h:76-130
//   SplineEntity::AddToProblem                                    K/trajectories/spline_base.h:380-426
//   SensorEntity / ImuEntity / ConstantBiasImuEntity::AddToProblem K/sensors/{sensors.h:137-167, imu.h:129-142,
//                                                                  constant_bias_imu.h:100-119}
// by index tables: per evaluation the first knot i0 and the interpolation amount u (times are constants because
// the time offsets are locked, cfg optimize_time_offset=false), and per parameter block its position in the linear
// system.  Ordering: trajectory knots in time order [r3_i, so3_i] (band part), then the "arrow" border: knots touched by
// the map-time evaluation of every surfel residual + the sensor blocks (SURVEY §5 "long-context"), then the landmarks'
// inverse depths, which are eliminated FIRST by a Schur complement (they are 1x1 diagonal blocks: what Ceres'
// SPARSE_SCHUR does with them as e-blocks); each landmark's coupling row [ref window | camera q,p | obs windows...] is
// laid out here.  Pure host C++ (no CUDA).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <exception>
#include <vector>

#include "../../include/lvi_exc_b200.h"
#include "residuals.cuh"

namespace lvi {

struct RangeError : std::runtime_error { using std::runtime_error::runtime_error; };

// data-parallel sharding: rank r of `world` evaluates the contiguous (= time-contiguous, tables are chronological)
// index range [lo, hi) of every residual table; the normal equations are summed across ranks (NCCL all-reduce).
inline void shard_range(int n, int rank, int world, int& lo, int& hi) {
  lo = static_cast<int>(static_cast<long long>(n) * rank / world);
  hi = static_cast<int>(static_cast<long long>(n) * (rank + 1) / world);
}

struct LoweredTable {
  int n = 0, active = 1;
  std::vector<int> i0a, i0b, ia, ib;
  std::vector<double> ua, ub, v, weight, huber;
  std::vector<int> perm;   // optional evaluation order of the normal-equation kernel (residuals with equal spline windows next to each other)
};

struct Lowered {
  int n_knots = 0, n_landmarks = 0, n_planes = 0, has_r3 = 1;
  double t0 = 0, dt = 1;
  LoweredTable tab[RT_COUNT];
  std::vector<int> pos_r3, pos_so3, pos_rho;
  int pos_sens[TB_COUNT];
  int nb = 0, nbo = 0, bw = 0;   // band dims, border dims, half bandwidth
  // Two-sided ("burn at both ends") ordering: the band part is split into chain 0 = [0, chain1_start) holding the first half of the
  // trajectory in time order and chain 1 = [chain1_start, nb) holding the second half in REVERSED time order; the knots in between
  // (wide enough that no residual or Schur row couples the two chains) lead the arrow border.  Both chains are ordinary band matrices
  // that only meet in the border, so they factor concurrently and the serial pivot chain halves.  chain1_start == nb: single chain.
  int chain1_start = 0;          // multiple of 32
  int n_mid = 0;                 // dims of the middle separator (= first n_mid border dims)
  int n_pad = 0;                 // unused dims padding chain 0 up to a tile boundary (positions chain1_start - n_pad .. chain1_start - 1)
  int n_rho = 0;                 // free inverse depths; tangent positions nb+nbo .. nb+nbo+n_rho-1
  // Schur rows of the inverse depths: landmark l couples to row_pos[row_start[l] .. row_start[l+1]) (-1: constant parameter);
  // slots: [0,24) reference window (cols 0..23 of its camera residuals), [24,30) camera q,p, then 24 per observation
  std::vector<int> row_start, row_pos;
  int n_res = 0, n_res_blocks = 0;
  int res_offset[RT_COUNT + 1];  // row offsets of each table in the residual vector
  bool constrained = false;      // a free block has bounds (rho >= 0) -> projected steps + Armijo check
  int nt() const { return nb + nbo + n_rho; }
};


inline void time_to_index(const lvi_problem_desc& d, double t, int& i0, double& u) {
  const double s = (t - d.t0) / d.dt;   // SplineSegmentView::CalculateIndexAndInterpolationAmount, spline_base.h:153-157
  i0 = static_cast<int>(std::floor(s));
  u = s - i0;
}
// Segment bookkeeping of SplineEntity::AddToProblem (spline_base.h:380-426) and the lookup of SplineView::Evaluate
// (spline_base.h:194-222) INCLUDING its retry at t - 1e-5 when t sits on a segment boundary (Q6): with timestamps
// commensurate with the knot grid the reference really evaluates 10 us early, and so do we.
struct SegList {
  double t0[2]; int n[2]; int first[2]; int count = 0;
};
inline SegList build_segments(const lvi_problem_desc& d, const double (*spans)[2], int nspans) {
  SegList S;
  int cur_start = 0, cur_end = -1;
  for (int k = 0; k < nspans; ++k) {
    int i1 = static_cast<int>(std::floor((spans[k][0] - d.t0) / d.dt));
    const int i2 = static_cast<int>(std::floor((spans[k][1] - d.t0) / d.dt));
    if (i1 > cur_end) { S.t0[S.count] = d.t0 + d.dt * i1; S.n[S.count] = 0; S.first[S.count] = i1; ++S.count; cur_start = i1; }
    else i1 = cur_end + 1;
    for (int i = i1; i < i2 + 4; ++i) S.n[S.count - 1] += 1;
    cur_end = cur_start + S.n[S.count - 1] - 1;
  }
  return S;
}
inline void locate(const lvi_problem_desc& d, const SegList& S, double t, int& i0, double& u) {
  for (int k = 0; k < S.count; ++k) {
    const double mn = S.t0[k], mx = S.t0[k] + (S.n[k] - 3) * d.dt;
    double te = t;
    bool ok = (t >= mn) && (t < mx);
    if (!ok) { te = t - 0.00001; ok = (te >= mn) && (te < mx); }
    if (!ok) continue;
    const double s = (te - S.t0[k]) / d.dt;
    const int il = static_cast<int>(std::floor(s));
    if (S.n[k] < 4 || il < 0 || il > S.n[k] - 4) throw RangeError("t is out of range for spline segment");
    i0 = S.first[k] + il;
    u = s - il;
    if (i0 < 0 || i0 > d.n_knots - 4) throw RangeError("t is out of range for spline");
    return;
  }
  throw RangeError("No segment found for time t");
}
inline void locate1(const lvi_problem_desc& d, double t_span, double t_eval, int& i0, double& u) {
  const double sp[1][2] = {{t_span, t_span}};
  locate(d, build_segments(d, sp, 1), t_eval, i0, u);
}
inline void locate2(const lvi_problem_desc& d, double ta, double tb, double ea, double eb, int& i0a, double& ua, int& i0b, double& ub) {
  const double sp[2][2] = {{ta, ta}, {tb, tb}};
  const SegList S = build_segments(d, sp, 2);
  locate(d, S, ea, i0a, ua);
  locate(d, S, eb, i0b, ub);
}

struct LowerLap {   // diagnostics: LVI_TIME_CREATE=1 prints the host phases of the lowering on stderr
  bool on = std::getenv("LVI_TIME_CREATE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void operator()(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[lvi]   lower: %-26s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};
// Chunks of these loops on extra host threads were measured SLOWER on the GPU box (create 10.1 -> 11.9 ms: a thread start costs more than
// the 0.3 - 1 ms a chunk saves); the one helper thread for the surfel table stays.

inline void check_span(const lvi_problem_desc& d, double t1, double t2) {  // CheckTimeSpans
  const double mn = d.t0, mx = d.t0 + (d.n_knots - 3) * d.dt;
  if (t1 < mn || t2 >= mx) throw RangeError("Time span out of range for trajectory");
  if (t1 > t2) throw RangeError("At least one time span begins before it ends");
}

inline void lower_problem(const lvi_problem_desc& d, Lowered& L) {
  if (d.n_knots < 4) throw RangeError("Spline had too few control points");
  if (!d.so3_knots) throw std::invalid_argument("so3_knots is null");
  L.n_knots = d.n_knots; L.n_landmarks = d.n_landmarks; L.n_planes = d.n_planes; L.t0 = d.t0; L.dt = d.dt;
  L.has_r3 = d.r3_knots != nullptr;
  const int n = d.n_knots;
  const bool r3_free = L.has_r3 && !d.lock_r3, so3_free = !d.lock_so3;
  const bool traj_free = r3_free || so3_free;
  std::vector<char> knot_used(n, 0);
  std::vector<char> rho_used(std::max(d.n_landmarks, 1), 0);
  bool sens_used[TB_COUNT] = {false, false, false, false, false, false, false};
  auto use_window = [&](int i0) { for (int k = i0; k < i0 + 4 && k < n; ++k) knot_used[k] = 1; };
  auto use_span = [&](double ta, double tb) { int ia, ib; double u; time_to_index(d, ta, ia, u); time_to_index(d, tb, ib, u); for (int k = ia; k < ib + 4 && k < n; ++k) if (k >= 0) knot_used[k] = 1; };
  auto need = [&](const void* p, int cnt, const char* what) { if (cnt > 0 && !p) throw std::invalid_argument(std::string("null table pointer: ") + what); };
  LowerLap lap;
  // ---- gyro / accel / orient: single evaluation
  {
    LoweredTable& T = L.tab[RT_GYRO];
    need(d.gyro_t, d.n_gyro, "gyro_t"); need(d.gyro_w, d.n_gyro, "gyro_w"); need(d.gyro_weight, d.n_gyro, "gyro_weight");
    T.n = d.n_gyro; T.active = 1;  // gravity blocks are never constant (imu.h:129-142, Q5)
    T.i0a.resize(T.n); T.ua.resize(T.n); T.v.assign(d.gyro_w, d.gyro_w + 3 * static_cast<size_t>(T.n)); T.weight.assign(d.gyro_weight, d.gyro_weight + T.n);
    for (int i = 0; i < T.n; ++i) {
      const double t = d.gyro_t[i] + d.imu_toff;
      check_span(d, d.gyro_t[i], d.gyro_t[i]);
      locate1(d, d.gyro_t[i], t, T.i0a[i], T.ua[i]);
      use_window(T.i0a[i]);
    }
    if (T.n) { sens_used[TB_G] = sens_used[TB_BA] = sens_used[TB_BG] = true; }
  }
  {
    LoweredTable& T = L.tab[RT_ACCEL];
    need(d.accel_t, d.n_accel, "accel_t"); need(d.accel_a, d.n_accel, "accel_a"); need(d.accel_weight, d.n_accel, "accel_weight");
    T.n = d.n_accel; T.active = 1;
    T.i0a.resize(T.n); T.ua.resize(T.n); T.v.assign(d.accel_a, d.accel_a + 3 * static_cast<size_t>(T.n)); T.weight.assign(d.accel_weight, d.accel_weight + T.n);
    for (int i = 0; i < T.n; ++i) {
      check_span(d, d.accel_t[i], d.accel_t[i]);
      locate1(d, d.accel_t[i], d.accel_t[i] + d.imu_toff, T.i0a[i], T.ua[i]);
      use_window(T.i0a[i]);
    }
    if (T.n) { sens_used[TB_G] = sens_used[TB_BA] = sens_used[TB_BG] = true; }
  }
  {
    LoweredTable& T = L.tab[RT_ORIENT];
    need(d.orient_t, d.n_orient, "orient_t"); need(d.orient_q, d.n_orient, "orient_q"); need(d.orient_weight, d.n_orient, "orient_weight");
    T.n = d.n_orient; T.active = so3_free;
    T.i0a.resize(T.n); T.ua.resize(T.n); T.v.assign(d.orient_q, d.orient_q + 4 * static_cast<size_t>(T.n)); T.weight.assign(d.orient_weight, d.orient_weight + T.n);
    for (int i = 0; i < T.n; ++i) {
      check_span(d, d.orient_t[i], d.orient_t[i]);
      locate1(d, d.orient_t[i], d.orient_t[i], T.i0a[i], T.ua[i]);
      if (T.active) use_window(T.i0a[i]);
    }
  }
  // ---- surfel (map time first: spans must be ordered, Q12)
  // The surfel table (the largest: 110 k rows at C2) is lowered on a second host thread while this one does the camera tables; it
  // marks knots in its own array, merged after the join.
  std::vector<char> border_flag(n + 4, 0);   // knots of the map-time windows (they go to the arrow border)
  std::vector<char> knot_used_surfel(n, 0), border_flag_cs(n + 4, 0);
  bool surfel_sens = false;
  // the table's set-up (sizes, copies of the plain columns) is done here, by the calling thread, before the helper starts; the rows are then
  // split: the helper thread takes the first 60 %, the calling thread the rest once it is through with the camera tables (it used to wait
  // 3 ms of the table's 4 ms for the helper)
  {
    LoweredTable& T = L.tab[RT_SURFEL];
    need(d.surfel_t, d.n_surfel, "surfel_t"); need(d.surfel_point, d.n_surfel, "surfel_point"); need(d.surfel_plane, d.n_surfel, "surfel_plane");
    need(d.surfel_tmap, d.n_surfel, "surfel_tmap"); need(d.surfel_weight, d.n_surfel, "surfel_weight"); need(d.surfel_huber, d.n_surfel, "surfel_huber");
    need(d.planes, d.n_surfel, "planes");
    if (d.n_surfel && !L.has_r3) throw std::invalid_argument("surfel residuals need the R3 spline");
    T.n = d.n_surfel; T.active = traj_free || !d.lock_lidar_q || !d.lock_lidar_p;
    T.i0a.resize(T.n); T.ua.resize(T.n); T.i0b.resize(T.n); T.ub.resize(T.n);
    if (T.n && T.active) surfel_sens = true;
  }
  std::vector<char> knot_used_surfel2(n, 0), border_flag2(n + 4, 0);
  auto lower_surfel = [&](int lo, int hi, std::vector<char>& used, std::vector<char>& flag, bool copies) {
    auto use_window = [&](int i0) { for (int k = i0; k < i0 + 4 && k < n; ++k) used[k] = 1; };
    LoweredTable& T = L.tab[RT_SURFEL];
    if (copies) {
      T.v.assign(d.surfel_point, d.surfel_point + 3 * static_cast<size_t>(T.n)); T.ia.assign(d.surfel_plane, d.surfel_plane + T.n);
      T.weight.assign(d.surfel_weight, d.surfel_weight + T.n); T.huber.assign(d.surfel_huber, d.surfel_huber + T.n);
    }
    for (int i = lo; i < hi; ++i) {
      check_span(d, d.surfel_tmap[i], d.surfel_tmap[i]); check_span(d, d.surfel_t[i], d.surfel_t[i]);
      if (d.surfel_t[i] < d.surfel_tmap[i]) throw RangeError("Time spans are not ordered");
      if (d.surfel_plane[i] < 0 || d.surfel_plane[i] >= d.n_planes) throw std::invalid_argument("surfel plane id out of range");
      locate2(d, d.surfel_tmap[i], d.surfel_t[i], d.surfel_tmap[i] + d.lidar_toff, d.surfel_t[i] + d.lidar_toff, T.i0a[i], T.ua[i], T.i0b[i], T.ub[i]);
      if (T.active) { use_window(T.i0a[i]); use_window(T.i0b[i]); for (int k = 0; k < 4; ++k) flag[T.i0a[i] + k] = 1; }
    }
  };
  const int surfel_split = d.n_surfel >= 4096 ? static_cast<int>(0.6 * d.n_surfel) : d.n_surfel;
  lap("imu tables");
  std::exception_ptr surfel_error;
  std::thread surfel_thread([&] { try { lower_surfel(0, surfel_split, knot_used_surfel, border_flag, true); } catch (...) { surfel_error = std::current_exception(); } });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } surfel_joiner{surfel_thread};
  // ---- camera
  std::vector<int> rho_anchor(std::max(d.n_landmarks, 1), -1);
  {
    LoweredTable& T = L.tab[RT_CAM];
    need(d.cam_t0_ref, d.n_cam, "cam_t0_ref"); need(d.cam_uv_ref, d.n_cam, "cam_uv_ref"); need(d.cam_landmark, d.n_cam, "cam_landmark");
    need(d.cam_t0_obs, d.n_cam, "cam_t0_obs"); need(d.cam_uv_obs, d.n_cam, "cam_uv_obs"); need(d.cam_weight, d.n_cam, "cam_weight");
    need(d.cam_huber, d.n_cam, "cam_huber"); need(d.rho, d.n_cam, "rho");
    if (d.n_cam && !L.has_r3) throw std::invalid_argument("camera residuals need the R3 spline");
    T.n = d.n_cam;
    // Ceres keeps a residual block while ANY of its parameter blocks is free.  Landmarks are unlocked by default (K/sfm/landmark.h), so
    // what counts is whether some landmark of this table is actually free, not whether the caller passed a lock array.
    bool any_rho_free = false;
    for (int i = 0; i < d.n_cam && !any_rho_free; ++i) {
      const int l = d.cam_landmark[i];
      any_rho_free = l >= 0 && l < d.n_landmarks && !(d.rho_locked && d.rho_locked[l]);
    }
    T.active = traj_free || !d.lock_cam_q || !d.lock_cam_p || any_rho_free;
    T.i0a.resize(T.n); T.ua.resize(T.n); T.i0b.resize(T.n); T.ub.resize(T.n); T.v.resize(4 * static_cast<size_t>(T.n));
    T.ia.assign(d.cam_landmark, d.cam_landmark + T.n);
    T.weight.assign(d.cam_weight, d.cam_weight + T.n); T.huber.assign(d.cam_huber, d.cam_huber + T.n);
    const double row_delta = d.readout / static_cast<double>(d.cam_rows);
    const double margin = 1e-3;
    for (int i = 0; i < T.n; ++i) {
      double t1 = d.cam_t0_ref[i], t2 = d.cam_t0_obs[i];
      if (!(t1 <= t2)) std::swap(t1, t2);
      check_span(d, t1 - margin, t1 + d.readout + margin); check_span(d, t2 - margin, t2 + d.readout + margin);
      if (T.ia[i] < 0 || T.ia[i] >= d.n_landmarks) throw std::invalid_argument("camera landmark id out of range");
      T.v[4 * i] = d.cam_uv_ref[2 * i]; T.v[4 * i + 1] = d.cam_uv_ref[2 * i + 1]; T.v[4 * i + 2] = d.cam_uv_obs[2 * i]; T.v[4 * i + 3] = d.cam_uv_obs[2 * i + 1];
      {
        const double sp[2][2] = {{t1 - margin, t1 + d.readout + margin}, {t2 - margin, t2 + d.readout + margin}};
        const SegList S = build_segments(d, sp, 2);
        locate(d, S, d.cam_t0_ref[i] + d.cam_toff + d.cam_uv_ref[2 * i + 1] * row_delta, T.i0a[i], T.ua[i]);
        locate(d, S, d.cam_t0_obs[i] + d.cam_toff + d.cam_uv_obs[2 * i + 1] * row_delta, T.i0b[i], T.ub[i]);
      }
      if (T.active) {
        use_span(t1 - margin, t1 + d.readout + margin); use_span(t2 - margin, t2 + d.readout + margin);
        rho_used[T.ia[i]] = 1;
        rho_anchor[T.ia[i]] = std::max(rho_anchor[T.ia[i]], T.i0a[i] + 3);
      }
    }
    if (T.n && T.active) sens_used[TB_CQ] = sens_used[TB_CP] = true;
  }
  {
    LoweredTable& T = L.tab[RT_CAMSURF];
    need(d.cs_t, d.n_camsurf, "cs_t"); need(d.cs_uv, d.n_camsurf, "cs_uv"); need(d.cs_landmark, d.n_camsurf, "cs_landmark"); need(d.cs_plane, d.n_camsurf, "cs_plane");
    need(d.cs_tmap, d.n_camsurf, "cs_tmap"); need(d.cs_weight, d.n_camsurf, "cs_weight"); need(d.cs_huber, d.n_camsurf, "cs_huber");
    need(d.planes, d.n_camsurf, "planes"); need(d.rho, d.n_camsurf, "rho");
    if (d.n_camsurf && !L.has_r3) throw std::invalid_argument("camera-surfel residuals need the R3 spline");
    T.n = d.n_camsurf;
    // blocks: trajectory, camera q,p, lidar q,p, plane (constant), rho (added as a block, read as a constant, Q4): the table is dropped
    // from the reduced program only when every one of them is constant
    bool any_rho_free = false;
    for (int i = 0; i < d.n_camsurf && !any_rho_free; ++i) {
      const int l = d.cs_landmark[i];
      any_rho_free = l >= 0 && l < d.n_landmarks && !(d.rho_locked && d.rho_locked[l]);
    }
    T.active = traj_free || !d.lock_cam_q || !d.lock_cam_p || !d.lock_lidar_q || !d.lock_lidar_p || any_rho_free;
    T.i0a.resize(T.n); T.ua.resize(T.n); T.i0b.resize(T.n); T.ub.resize(T.n);
    T.v.assign(d.cs_uv, d.cs_uv + 2 * static_cast<size_t>(T.n)); T.ia.assign(d.cs_plane, d.cs_plane + T.n); T.ib.assign(d.cs_landmark, d.cs_landmark + T.n);
    T.weight.assign(d.cs_weight, d.cs_weight + T.n); T.huber.assign(d.cs_huber, d.cs_huber + T.n);
    for (int i = 0; i < T.n; ++i) {
      check_span(d, d.cs_tmap[i], d.cs_tmap[i]); check_span(d, d.cs_t[i], d.cs_t[i]);
      if (d.cs_t[i] < d.cs_tmap[i]) throw RangeError("Time spans are not ordered");
      if (T.ia[i] < 0 || T.ia[i] >= d.n_planes || T.ib[i] < 0 || T.ib[i] >= d.n_landmarks) throw std::invalid_argument("camera-surfel id out of range");
      locate2(d, d.cs_tmap[i], d.cs_t[i], d.cs_tmap[i] + d.cam_toff, d.cs_t[i] + d.cam_toff, T.i0a[i], T.ua[i], T.i0b[i], T.ub[i]);
      if (!T.active) continue;
      use_window(T.i0a[i]); use_window(T.i0b[i]);
      for (int k = 0; k < 4; ++k) border_flag_cs[T.i0a[i] + k] = 1;
      rho_used[T.ib[i]] = 1;
      rho_anchor[T.ib[i]] = std::max(rho_anchor[T.ib[i]], T.i0b[i] + 3);
    }
    if (T.n && T.active) sens_used[TB_CQ] = sens_used[TB_CP] = sens_used[TB_LQ] = sens_used[TB_LP] = true;
  }
  lap("camera tables");
  lower_surfel(surfel_split, d.n_surfel, knot_used_surfel2, border_flag2, false);   // the calling thread's share of the surfel rows
  lap("surfel rows (own share)");
  surfel_thread.join();
  lap("wait for the surfel thread");
  if (surfel_error) std::rethrow_exception(surfel_error);
  for (int i = 0; i < n; ++i) { knot_used[i] |= knot_used_surfel[i] | knot_used_surfel2[i]; border_flag[i] |= border_flag_cs[i] | border_flag2[i]; }
  if (surfel_sens) sens_used[TB_LQ] = sens_used[TB_LP] = true;
  std::set<int> border_knots;
  for (int i = 0; i < n; ++i) if (border_flag[i]) border_knots.insert(i);
  if (border_knots.size() > 16) border_knots.clear();  // not an arrow structure: leave those knots in the band
  // ---- positions
  L.pos_r3.assign(n, -1); L.pos_so3.assign(n, -1); L.pos_rho.assign(std::max(d.n_landmarks, 1), -1);
  std::vector<int> band;  // band knots in time order
  for (int i = 0; i < n; ++i) if (knot_used[i] && !border_knots.count(i)) band.push_back(i);
  // longest coupling in knots among band knots: 3 inside one evaluation window, reference -> last observation for the camera
  int kspan = 3;
  {
    const LoweredTable& TS = L.tab[RT_SURFEL];
    if (border_knots.empty()) for (int i = 0; i < TS.n; ++i) kspan = std::max(kspan, TS.i0b[i] + 3 - TS.i0a[i]);
    const LoweredTable& TC = L.tab[RT_CAM];
    for (int i = 0; i < TC.n; ++i) kspan = std::max(kspan, std::abs(TC.i0b[i] - TC.i0a[i]) + 3);
    std::vector<int> lo(std::max(d.n_landmarks, 1), 1 << 30), hi(std::max(d.n_landmarks, 1), -1);
    for (int i = 0; i < TC.n; ++i) { const int l = TC.ia[i]; lo[l] = std::min(lo[l], std::min(TC.i0a[i], TC.i0b[i])); hi[l] = std::max(hi[l], std::max(TC.i0a[i], TC.i0b[i]) + 3); }
    for (int l = 0; l < d.n_landmarks; ++l) if (hi[l] >= 0) kspan = std::max(kspan, hi[l] - lo[l]);  // Schur fill couples all windows of a landmark
    const LoweredTable& TV = L.tab[RT_CAMSURF];
    if (border_knots.empty()) for (int i = 0; i < TV.n; ++i) kspan = std::max(kspan, TV.i0b[i] + 3 - TV.i0a[i]);
  }
  const int m = static_cast<int>(band.size()), wmid = kspan + 1;
  const bool twist = traj_free && m >= 8 * wmid && !std::getenv("LVI_NO_TWIST");
  const int mid_a = twist ? (m - wmid) / 2 : m, mid_b = twist ? mid_a + wmid : m;   // band[mid_a, mid_b) = separator
  int pb = 0;
  auto place = [&](int i) { if (r3_free) { L.pos_r3[i] = pb; pb += 3; } if (so3_free) { L.pos_so3[i] = pb; pb += 3; } };
  for (int k = 0; k < mid_a; ++k) place(band[k]);
  L.n_pad = 0;
  if (twist) { L.n_pad = (32 - pb % 32) % 32; pb += L.n_pad; }
  L.chain1_start = pb;
  for (int k = m - 1; k >= mid_b; --k) place(band[k]);
  L.nb = pb;
  if (!twist) L.chain1_start = L.nb;
  for (int k = mid_a; k < mid_b; ++k) place(band[k]);
  L.n_mid = pb - L.nb;
  for (int i : border_knots) {
    if (!knot_used[i]) continue;
    place(i);
  }
  const bool sens_free[TB_COUNT] = {!d.lock_lidar_q, !d.lock_lidar_p, !d.lock_cam_q, !d.lock_cam_p, true, !d.lock_acc_bias, !d.lock_gyr_bias};
  const int sens_dim[TB_COUNT] = {3, 3, 3, 3, 2, 3, 3};
  for (int b = 0; b < TB_COUNT; ++b) {
    L.pos_sens[b] = -1;
    if (sens_used[b] && sens_free[b]) { L.pos_sens[b] = pb; pb += sens_dim[b]; }
  }
  L.nbo = pb - L.nb;
  L.n_rho = 0;
  for (int l = 0; l < d.n_landmarks; ++l) {
    const bool locked = d.rho_locked && d.rho_locked[l];
    if (rho_used[l] && !locked) { L.pos_rho[l] = pb++; ++L.n_rho; L.constrained = true; }
  }
  lap("positions");
  // Schur rows: slot base of every camera residual inside its landmark's row (stored in tab[RT_CAM].ib)
  {
    LoweredTable& T = L.tab[RT_CAM];
    const int nl = std::max(d.n_landmarks, 1);
    std::vector<int> cnt(nl, 0), ref_i0(nl, -1);
    T.ib.assign(T.n, 0);
    for (int i = 0; i < T.n; ++i) {
      const int l = T.ia[i];
      if (ref_i0[l] < 0) ref_i0[l] = T.i0a[i];
      else if (ref_i0[l] != T.i0a[i]) throw std::invalid_argument("camera residuals of one landmark must share its reference observation (Landmark::reference())");
      T.ib[i] = 30 + 24 * cnt[l]++;
    }
    L.row_start.assign(nl + 1, 0);
    for (int l = 0; l < nl; ++l) L.row_start[l + 1] = L.row_start[l] + ((T.active && l < d.n_landmarks && L.pos_rho[l] >= 0 && cnt[l] > 0) ? 30 + 24 * cnt[l] : 0);
    L.row_pos.assign(std::max(L.row_start[nl], 1), -1);
  }
  // Camera residuals arrive landmark by landmark, so neighbours almost never share both spline windows and the normal-equation kernel
  // cannot merge their J^T J before its atomics.  Its evaluation order is therefore sorted by window pair (the tables keep their order).
  {
    LoweredTable& T = L.tab[RT_CAM];
    T.perm.resize(T.n);
    for (int i = 0; i < T.n; ++i) T.perm[i] = i;
    std::stable_sort(T.perm.begin(), T.perm.end(), [&](int x, int y) { return T.i0a[x] != T.i0a[y] ? T.i0a[x] < T.i0a[y] : T.i0b[x] < T.i0b[y]; });
  }
  lap("schur rows + camera order");
  // ---- residual vector layout + bandwidth
  int ro = 0, nblk = 0;
  for (int t = 0; t < RT_COUNT; ++t) { L.res_offset[t] = ro; ro += L.tab[t].n * rt_rows(t); nblk += L.tab[t].n; }
  L.res_offset[RT_COUNT] = ro; L.n_res = ro; L.n_res_blocks = nblk;
  L.bw = 0;  // filled by compute_bandwidth()
}


// host ProblemView over the lowered tables (parameters are read straight from the desc)
inline ProblemView host_view(const lvi_problem_desc& d, const Lowered& L, const double* sens /*SENS_N*/) {
  ProblemView P{};
  P.dt_inv = 1.0 / d.dt; P.n_knots = d.n_knots; P.r3 = d.r3_knots; P.so3 = d.so3_knots; P.sens = sens; P.rho = d.rho; P.planes = d.planes;
  P.fx = d.fx; P.fy = d.fy; P.cx = d.cx; P.cy = d.cy; P.has_r3 = L.has_r3;
  for (int k = 0; k < 5; ++k) P.dist[k] = d.distortion[k];
  P.do_dist = (std::fabs(d.distortion[0]) > 1e-5 || std::fabs(d.distortion[1]) > 1e-5 || std::fabs(d.distortion[2]) > 1e-5) ? 1 : 0;   // pinhole_camera.h:78
  for (int t = 0; t < RT_COUNT; ++t) {
    const LoweredTable& T = L.tab[t];
    ResTable& R = P.tab[t];
    R.n = T.n; R.lo = 0; R.hi = T.n; R.active = T.active; R.i0a = T.i0a.data(); R.ua = T.ua.data(); R.i0b = T.i0b.data(); R.ub = T.ub.data(); R.v = T.v.data();
    R.ia = T.ia.data(); R.ib = T.ib.data(); R.perm = nullptr; R.weight = T.weight.data(); R.huber = T.huber.empty() ? nullptr : T.huber.data();
  }
  P.pos_r3 = L.pos_r3.data(); P.pos_so3 = L.pos_so3.data(); P.pos_rho = L.pos_rho.data();
  for (int b = 0; b < TB_COUNT; ++b) P.pos_sens[b] = L.pos_sens[b];
  return P;
}

inline void pack_sens(const lvi_problem_desc& d, double* s) {
  static const double ident[4] = {0, 0, 0, 1}, zero[3] = {0, 0, 0};
  const double* lq = d.lidar_q ? d.lidar_q : ident; const double* lp = d.lidar_p ? d.lidar_p : zero;
  const double* cq = d.cam_q ? d.cam_q : ident; const double* cp = d.cam_p ? d.cam_p : zero;
  for (int k = 0; k < 4; ++k) { s[SENS_LQ + k] = lq[k]; s[SENS_CQ + k] = cq[k]; }
  for (int k = 0; k < 3; ++k) { s[SENS_LP + k] = lp[k]; s[SENS_CP + k] = cp[k]; s[SENS_BA + k] = d.acc_bias ? d.acc_bias[k] : 0; s[SENS_BG + k] = d.gyr_bias ? d.gyr_bias[k] : 0; }
  s[SENS_G] = d.gravity ? d.gravity[0] : 0; s[SENS_G + 1] = d.gravity ? d.gravity[1] : 0;
}

// half bandwidth inside each chain; a residual must never couple the two chains directly (the separator guarantees it)
struct SpanAcc {
  int c1, nb, bw = 0;
  int lo[2], hi[2];
  void begin() { lo[0] = lo[1] = 1 << 30; hi[0] = hi[1] = -1; }
  void add(int p) { if (p >= 0 && p < nb) { const int c = p >= c1; lo[c] = std::min(lo[c], p); hi[c] = std::max(hi[c], p); } }
  void end() {
    if (hi[0] >= 0 && hi[1] >= 0) throw std::logic_error("lowering: a coupling crosses the separator of the two-sided ordering");
    for (int c = 0; c < 2; ++c) if (hi[c] >= 0) bw = std::max(bw, hi[c] - lo[c]);
  }
};
// band positions of one spline window (knots i0 .. i0 + 3; 3 dims per block): the only band columns a residual has -- its sensor blocks sit
// in the border and its inverse depth is eliminated, both at positions >= nb, which SpanAcc ignores
inline void add_window(SpanAcc& A, const ProblemView& P, int i0, bool r3, bool so3) {
  for (int k = 0; k < 4; ++k) {
    if (r3) { const int p = P.pos_r3[i0 + k]; if (p >= 0) { A.add(p); A.add(p + 2); } }
    if (so3) { const int p = P.pos_so3[i0 + k]; if (p >= 0) { A.add(p); A.add(p + 2); } }
  }
}
template <int TYPE> inline void bw_of_type(const ProblemView& P, SpanAcc& A, const int* order = nullptr) {
  const ResTable& T = P.tab[TYPE];
  if (!T.active) return;
  // consecutive residuals are time-ordered: evaluate each distinct window pair once
  constexpr bool two = TYPE == RT_SURFEL || TYPE == RT_CAM || TYPE == RT_CAMSURF;
  constexpr bool r3 = TYPE != RT_GYRO && TYPE != RT_ORIENT;
  int last_a = -1, last_b = -1;
  for (int q = 0; q < T.n; ++q) {
    const int i = order ? order[q] : q;
    const int wa = T.i0a ? T.i0a[i] : 0, wb = (two && T.i0b) ? T.i0b[i] : 0;
    if (q > 0 && wa == last_a && wb == last_b) continue;
    last_a = wa; last_b = wb;
    A.begin();
    add_window(A, P, wa, r3, true);
    if (two) add_window(A, P, wb, r3, true);
    A.end();
  }
}
// second lowering pass (needs the column -> position map): positions of every Schur-row slot, and the half bandwidth of the
// band part INCLUDING the fill the inverse-depth elimination creates between the windows of one landmark
inline void compute_bandwidth(const ProblemView& P, Lowered& L) {
  SpanAcc acc;
  acc.c1 = L.chain1_start; acc.nb = L.nb;
  bw_of_type<RT_GYRO>(P, acc); bw_of_type<RT_ACCEL>(P, acc); bw_of_type<RT_SURFEL>(P, acc);
  bw_of_type<RT_CAM>(P, acc, L.tab[RT_CAM].perm.empty() ? nullptr : L.tab[RT_CAM].perm.data()); bw_of_type<RT_CAMSURF>(P, acc); bw_of_type<RT_ORIENT>(P, acc);
  const LoweredTable& T = L.tab[RT_CAM];
  for (int i = 0; i < T.n; ++i) {
    const int l = T.ia[i];
    const int rs = L.row_start[l];
    if (L.row_start[l + 1] == rs) continue;
    // same values as col_pos<RT_CAM>(P, i, c), c = 0..53, written window by window (this loop is per camera residual: 32 k at C2)
    auto window = [&](int i0, int* dst) {   // [r3 of knots i0..i0+3 | so3 of knots i0..i0+3], 3 dims each
      for (int k = 0; k < 4; ++k) {
        const int pr = P.pos_r3[i0 + k], ps = P.pos_so3[i0 + k];
        for (int d3 = 0; d3 < 3; ++d3) { dst[3 * k + d3] = pr < 0 ? -1 : pr + d3; dst[12 + 3 * k + d3] = ps < 0 ? -1 : ps + d3; }
      }
    };
    window(T.i0a[i], &L.row_pos[rs]);
    window(T.i0b[i], &L.row_pos[rs + T.ib[i]]);
    for (int d3 = 0; d3 < 3; ++d3) {
      L.row_pos[rs + 24 + d3] = P.pos_sens[TB_CQ] < 0 ? -1 : P.pos_sens[TB_CQ] + d3;
      L.row_pos[rs + 27 + d3] = P.pos_sens[TB_CP] < 0 ? -1 : P.pos_sens[TB_CP] + d3;
    }
  }
  // half bandwidth of the Schur rows: the band positions of a landmark's row are those of its residuals' windows, so the span of a row is
  // accumulated per window (16 positions per residual) instead of re-reading the 54 slots per residual just written (1.7 M ints at C2)
  {
    const int nl = static_cast<int>(L.row_start.size()) - 1;
    std::vector<SpanAcc> row(nl);
    for (auto& a : row) { a.c1 = acc.c1; a.nb = acc.nb; a.begin(); }
    for (int i = 0; i < T.n; ++i) {
      const int l = T.ia[i];
      if (L.row_start[l + 1] == L.row_start[l]) continue;
      add_window(row[l], P, T.i0a[i], true, true);
      add_window(row[l], P, T.i0b[i], true, true);
    }
    for (auto& a : row) { a.end(); acc.bw = std::max(acc.bw, a.bw); }
  }
  L.bw = acc.bw;
}

}  // namespace lvi

// undistort.cu — scan de-skew / map assembly on sm_100a (the step right before the map build).
//
// Replaces ScanUndistortion::undistort (L/include/core/scan_undistortion.h:132-180) with
// TrajectoryManagerLVI::evaluateLidarPose (L/src/core/trajectory_manager_lvi.cpp:398-408): one spline pose evaluation
// per raw point (17 M per pass at 60 s; serial, heap-allocating on the CPU), and pcl::transformPointCloud with a float
// 4x4 (scan_undistortion.h:111, L/src/core/lidar_odometry.cpp:98).
//   undistort_target_kernel : pose of the LiDAR at each scan's target time (own stamp or the map time)
//   undistort_kernel        : thread per point, 32 B coalesced load (PointXYZIT) -> closed-form SO3/R3 spline pose ->
//                             32 B coalesced store (PointXYZI).  HBM-bound: 64 B/point + cache-resident knots.
//   transform_kernel        : float 4x4 per scan, evaluated left to right without FMA (bit-exact with Eigen's float path).
// Compiled with -fmad=false.
#include <memory>

#include "map.cuh"
#include "spline_math.cuh"

namespace lvi {

struct TrajView {
  double t0, dt, dt_inv, t_max, toff;
  int n_knots;
  const double* r3;   // [n*3]
  const double* so3;  // [n*4]
  const double* hlog; // [n*3] half-angle log of q_{i-1}^-1 q_i per knot (so3_log_kernel), row 0 unused
  Q4 qL; V3 pL;
};

// knot-to-knot logs once per trajectory instead of three atan2/sqrt chains per point
__global__ void so3_log_kernel(const double* __restrict__ so3, int n, double* __restrict__ hlog) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 h = v3(0, 0, 0);
  if (i >= 1) {
    const Q4 a = q4(so3[4 * i - 4], so3[4 * i - 3], so3[4 * i - 2], so3[4 * i - 1]), b = q4(so3[4 * i], so3[4 * i + 1], so3[4 * i + 2], so3[4 * i + 3]);
    h = logq_half(qmul(qconj(a), b));
  }
  hlog[3 * i] = h.x; hlog[3 * i + 1] = h.y; hlog[3 * i + 2] = h.z;
}

__device__ __forceinline__ bool lidar_pose(const TrajView& T, double t, Q4& q, V3& p) {
  const double tt = t + T.toff;
  if (T.t0 > tt || T.t_max <= tt || !(tt == tt)) return false;  // evaluateLidarPose range test (trajectory_manager_lvi.cpp:401-403)
  const double s = (tt - T.t0) / T.dt;
  int i0 = static_cast<int>(floor(s));
  if (i0 < 0 || i0 > T.n_knots - 4) return false;
  const double u = s - i0;
  const Q4 qI = so3_orientation_pre(T.so3 + 4 * i0, T.hlog + 3 * (i0 + 1), u);
  double B[4];
  basis_pos(u, B);
  const V3 pI = r3_spline(T.r3 + 3 * i0, B);
  q = qmul(qI, T.qL);                  // q_LtoG = q_ItoG * q_LtoI
  p = qrot(qI, T.pL) + pI;             // p_LinG = q_ItoG * p_LinI + p_IinG
  return true;
}

struct TargetPose { Q4 q; V3 p; int ok; int pad; double c[13]; };   // c = R(q)^T row-major (9), p (3), ok as a double: what the per-point kernel reads

__global__ void undistort_target_kernel(TrajView T, const double* __restrict__ target_time, int n_scans, TargetPose* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  TargetPose tp;
  tp.q = q4(0, 0, 0, 1); tp.p = v3(0, 0, 0); tp.pad = 0;
  tp.ok = lidar_pose(T, target_time[s], tp.q, tp.p) ? 1 : 0;
  const M3 R = qmat(tp.q);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) tp.c[3 * r + c] = R.m[3 * c + r];
  tp.c[9] = tp.p.x; tp.c[10] = tp.p.y; tp.c[11] = tp.p.z; tp.c[12] = tp.ok ? 1.0 : 0.0;
  out[s] = tp;
}

// ---- the per-point pose of the de-skew kernel -----------------------------------------------------------------------------------------------
// undistort_kernel is bound by the fp64 pipe, not by HBM (one cumulative-SO3 pose per point), so its arithmetic is written out with explicit
// fused multiply-adds (the file is compiled with -fmad=false for the float 4x4 path) and stripped of everything that does not depend on the
// point's own stamp:
//   * the knot-to-knot logs come as (unit axis, half angle) per knot (so3_axis_kernel): exp(B~_j(u) h_j) = (sin(B~_j a_j) n_j, cos(B~_j a_j))
//     -- no square root and no division per point, and B~_j a_j < 0.25 rad for any trajectory a 0.02 s spline can follow, where sin / cos are
//     short Taylor polynomials (|error| < 3e-18); larger angles take sincos();
//   * the point is moved by matrices that are constant per call (R_L, p_L) or per scan (R_0^T, p_0 of the target pose):
//       out = R_0^T [ R_I(t) (R_L x + p_L) + p_I(t) - p_0 ]            (rotation only: out = R_0^T R_I(t) R_L x),
//     which is q_0^-1 (q_I q_L) x + q_0^-1 (q_I p_L + p_I - p_0) of the reference (scan_undistortion.h:150-170) with one quaternion rotation per
//     point instead of three products and three rotations.
__global__ void so3_axis_kernel(const double* __restrict__ hlog, int n, double* __restrict__ hax /* [n*4]: n_x n_y n_z a */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 h = v3(hlog[3 * i], hlog[3 * i + 1], hlog[3 * i + 2]);
  const double a = sqrt(dot(h, h));
  const double inv = a > 0.0 ? 1.0 / a : 0.0;
  hax[4 * i] = h.x * inv; hax[4 * i + 1] = h.y * inv; hax[4 * i + 2] = h.z * inv; hax[4 * i + 3] = a;
}

__device__ __forceinline__ void sincos_small(double x, double& sn, double& cs) {
  if (fabs(x) < 0.25) {
    const double x2 = x * x;
    double p = -2.5052108385441720e-08;                 // -1/11!
    p = fma(p, x2, 2.7557319223985893e-06);             //  1/9!
    p = fma(p, x2, -1.9841269841269841e-04);            // -1/7!
    p = fma(p, x2, 8.3333333333333332e-03);             //  1/5!
    p = fma(p, x2, -1.6666666666666666e-01);            // -1/3!
    sn = fma(x * x2, p, x);
    double c = 2.0876756987868100e-09;                  //  1/12!
    c = fma(c, x2, -2.7557319223985888e-07);            // -1/10!
    c = fma(c, x2, 2.4801587301587302e-05);             //  1/8!
    c = fma(c, x2, -1.3888888888888889e-03);            // -1/6!
    c = fma(c, x2, 4.1666666666666664e-02);             //  1/4!
    c = fma(c, x2, -0.5);
    cs = fma(c, x2, 1.0);
  } else {
    sincos(x, &sn, &cs);
  }
}

__device__ __forceinline__ Q4 qmul_fma(const Q4& a, const Q4& b) {
  Q4 r;
  r.x = fma(a.w, b.x, fma(a.x, b.w, fma(a.y, b.z, -(a.z * b.y))));
  r.y = fma(a.w, b.y, fma(a.y, b.w, fma(a.z, b.x, -(a.x * b.z))));
  r.z = fma(a.w, b.z, fma(a.z, b.w, fma(a.x, b.y, -(a.y * b.x))));
  r.w = fma(a.w, b.w, -fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)));
  return r;
}

struct DeskewConst { double RL[9]; double pL[3]; };   // R(q_L) row-major, p_L

// PACKED = false: out is pcl::PointXYZI-shaped (2 x float4 per point, the ABI layout).  PACKED = true: out is the library's own scan batch
// (one float4 per point: x y z intensity) and the finite points are folded into their scan's min/max slots on the way out.
template <bool PACKED>
__global__ void __launch_bounds__(256, 4) undistort_kernel(TrajView T, DeskewConst K, const double* __restrict__ hax, const lvi_point_xyzit* __restrict__ raw,
                                                        int64_t n, int64_t pts_per_scan, const TargetPose* __restrict__ target, int correct_position,
                                                        float4* __restrict__ out, int* __restrict__ mm) {
  __shared__ double s_tp[2][13];   // R_0^T (9), p_0 (3), ok
  __shared__ float4 s_raw[3][2][256];
  ScanMinMax acc;
  acc.reset(-1);
  const float nanv = __int_as_float(0x7fc00000);
  for (int64_t c0 = static_cast<int64_t>(blockIdx.x) * kProducerChunk; c0 < n; c0 += static_cast<int64_t>(gridDim.x) * kProducerChunk) {
  int64_t scan = c0 / pts_per_scan;                       // one 64-bit division per chunk, not per point
  int64_t in_scan = c0 - scan * pts_per_scan + threadIdx.x;
  // the target poses of the (at most two, for any real scan size) scans a chunk of 2048 points touches are staged in shared memory; the raw
  // record of the NEXT point is requested before the current one is worked on, so the HBM latency of the stream hides behind ~400
  // instructions of pose arithmetic
  __syncthreads();
  if (threadIdx.x < 26) {
    const int w = threadIdx.x / 13, k = threadIdx.x % 13;
    const int64_t sc = min(scan + w, (n - 1) / pts_per_scan);
    s_tp[w][k] = target[sc].c[k];
  }
  __syncthreads();
  const int64_t scan_first = scan;
  // raw records: asynchronous copies (cp.async) into this thread's own shared-memory slots, two points ahead of the arithmetic
  auto issue = [&](int jj) {
    const int64_t ii = c0 + threadIdx.x + 256 * jj;
    if (jj < kProducerChunk / 256 && ii < n) {
      const unsigned d0 = static_cast<unsigned>(__cvta_generic_to_shared(&s_raw[jj % 3][0][threadIdx.x]));
      const unsigned d1 = static_cast<unsigned>(__cvta_generic_to_shared(&s_raw[jj % 3][1][threadIdx.x]));
      const float4* src = reinterpret_cast<const float4*>(raw + ii);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(src + 1) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0); issue(1);
  int64_t i = c0 + threadIdx.x;
  for (int j = 0; j < kProducerChunk / 256; ++j, in_scan += 256, i += 256) {
    issue(j + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    if (i >= n) break;
    const float4 a = s_raw[j % 3][0][threadIdx.x], b = s_raw[j % 3][1][threadIdx.x];
    while (in_scan >= pts_per_scan) { in_scan -= pts_per_scan; ++scan; }
    const double* Rt = s_tp[0];
    if (scan - scan_first == 1) Rt = s_tp[1];
    else if (scan != scan_first) Rt = target[scan].c;   // scans shorter than a chunk: straight from global memory
    const bool tp_ok = Rt[12] != 0.0;
    const double p0x = Rt[9], p0y = Rt[10], p0z = Rt[11];
    const double ts = __hiloint2double(__float_as_int(b.w), __float_as_int(b.z));   // x y z pad | intensity pad2 timestamp(lo, hi)
    float4 o0 = make_float4(0.f, 0.f, 0.f, 1.f), o1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!tp_ok || isnan(a.x)) {
      o0.x = nanv; o0.y = nanv; o0.z = nanv;
    } else {
      const double tt = ts + T.toff;
      if (!(T.t0 > tt || T.t_max <= tt || !(tt == tt))) {   // evaluateLidarPose range test (trajectory_manager_lvi.cpp:401-403)
        // span and local time.  (tt - t0) * (1 / dt) can differ from the reference's division in the last place, which only matters when
        // the stamp sits within an ulp of a knot: the spline is C2 there, so either span gives the same pose to 1e-15.
        const double s = (tt - T.t0) * T.dt_inv;
        const int i0 = min(max(static_cast<int>(floor(s)), 0), T.n_knots - 4);
        const double u = s - i0;
        const double u2 = u * u, u3 = u2 * u;
        // cumulative basis (spline_math.cuh basis_cumul) and position basis
        const double c3 = u3 * (1.0 / 6.0);
        const double c2 = fma(-2.0, u3, fma(3.0, u2, fma(3.0, u, 1.0))) * (1.0 / 6.0);
        const double c1 = fma(1.0, u3, fma(-3.0, u2, fma(3.0, u, 5.0))) * (1.0 / 6.0);
        const double2 qa = __ldg(reinterpret_cast<const double2*>(T.so3) + 2 * i0), qb = __ldg(reinterpret_cast<const double2*>(T.so3) + 2 * i0 + 1);
        Q4 q = q4(qa.x, qa.y, qb.x, qb.y);
        const double cb[3] = {c1, c2, c3};
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
          const double2 ha = __ldg(reinterpret_cast<const double2*>(hax) + 2 * (i0 + 1 + jj)), hb = __ldg(reinterpret_cast<const double2*>(hax) + 2 * (i0 + 1 + jj) + 1);
          double sn, cs;
          sincos_small(cb[jj] * hb.y, sn, cs);
          q = qmul_fma(q, q4(sn * ha.x, sn * ha.y, sn * hb.x, cs));
        }
        // y = R_L x (+ p_L)
        const double x = a.x, y = a.y, z = a.z;
        double y0 = fma(K.RL[0], x, fma(K.RL[1], y, K.RL[2] * z)), y1 = fma(K.RL[3], x, fma(K.RL[4], y, K.RL[5] * z)),
               y2 = fma(K.RL[6], x, fma(K.RL[7], y, K.RL[8] * z));
        if (correct_position) { y0 += K.pL[0]; y1 += K.pL[1]; y2 += K.pL[2]; }
        // z = R(q) y : v + w (2 u x v) + u x (2 u x v)
        const double t0 = 2.0 * fma(q.y, y2, -(q.z * y1)), t1 = 2.0 * fma(q.z, y0, -(q.x * y2)), t2 = 2.0 * fma(q.x, y1, -(q.y * y0));
        double z0 = fma(q.w, t0, y0) + fma(q.y, t2, -(q.z * t1));
        double z1 = fma(q.w, t1, y1) + fma(q.z, t0, -(q.x * t2));
        double z2 = fma(q.w, t2, y2) + fma(q.x, t1, -(q.y * t0));
        if (correct_position) {
          const double b0 = 1.0 - c1, b1 = c1 - c2, b2 = c2 - c3, b3 = c3;
          const double* cp = T.r3 + 3 * i0;
          z0 += fma(b0, __ldg(cp), fma(b1, __ldg(cp + 3), fma(b2, __ldg(cp + 6), b3 * __ldg(cp + 9)))) - p0x;
          z1 += fma(b0, __ldg(cp + 1), fma(b1, __ldg(cp + 4), fma(b2, __ldg(cp + 7), b3 * __ldg(cp + 10)))) - p0y;
          z2 += fma(b0, __ldg(cp + 2), fma(b1, __ldg(cp + 5), fma(b2, __ldg(cp + 8), b3 * __ldg(cp + 11)))) - p0z;
        }
        o0.x = static_cast<float>(fma(Rt[0], z0, fma(Rt[1], z1, Rt[2] * z2)));
        o0.y = static_cast<float>(fma(Rt[3], z0, fma(Rt[4], z1, Rt[5] * z2)));
        o0.z = static_cast<float>(fma(Rt[6], z0, fma(Rt[7], z1, Rt[8] * z2)));
        o1.x = b.x;
      }  // else: the reference leaves the default-constructed point (zeros)
    }
    if (PACKED) {
      __stcs(out + i, make_float4(o0.x, o0.y, o0.z, o1.x));
      acc.add(mm, scan, o0.x, o0.y, o0.z);
    } else {
      __stcs(out + 2 * i, o0);
      __stcs(out + 2 * i + 1, o1);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (PACKED) acc.flush(mm);
}

__global__ void scan_minmax_init_kernel(int* mm, int n_slots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_slots * 6) mm[i] = (i % 6) < 3 ? f2ord(3.402823466e38f) : f2ord(-3.402823466e38f);
}

// IMU pose of the trajectory at arbitrary times (SplitTrajectory::Evaluate, K/trajectories/split_trajectory.h:41-58), used by
// the landmark association (L/src/core/surfel_association.cpp:161-195 evaluates the camera pose per landmark).
__global__ void traj_eval_kernel(TrajView T, const double* __restrict__ t, int64_t n, double* __restrict__ pos, double* __restrict__ quat,
                                 unsigned char* __restrict__ valid) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  TrajView Ti = T;
  Ti.qL = q4(0, 0, 0, 1); Ti.pL = v3(0, 0, 0); Ti.toff = 0.0;
  Q4 q = q4(0, 0, 0, 1); V3 p = v3(0, 0, 0);
  const bool ok = lidar_pose(Ti, t[i], q, p);
  pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
  quat[4 * i] = q.x; quat[4 * i + 1] = q.y; quat[4 * i + 2] = q.z; quat[4 * i + 3] = q.w;
  valid[i] = ok ? 1 : 0;
}

// all five quantities of kontiki::trajectories::TrajectoryEvaluation (K/trajectories/trajectory.h:27-37) at arbitrary times:
// position / velocity / acceleration of the R3 spline (uniform_r3_spline_trajectory.h:36-103), orientation and WORLD-frame angular
// velocity of the SO3 spline (uniform_so3_spline_trajectory.h:46-125)
__global__ void traj_eval_full_kernel(TrajView T, const double* __restrict__ t, int64_t n, double* __restrict__ pos, double* __restrict__ vel,
                                      double* __restrict__ acc, double* __restrict__ quat, double* __restrict__ angvel, unsigned char* __restrict__ valid) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double tt = t[i];
  bool ok = !(T.t0 > tt || T.t_max <= tt || !(tt == tt));
  int i0 = 0; double u = 0;
  if (ok) {
    const double s = (tt - T.t0) / T.dt;
    i0 = static_cast<int>(floor(s));
    u = s - i0;
    ok = i0 >= 0 && i0 <= T.n_knots - 4;
  }
  V3 p = v3(0, 0, 0), v = p, a = p, w = p;
  Q4 q = q4(0, 0, 0, 1);
  if (ok) {
    double B[4], Bv[4], Ba[4];
    basis_pos(u, B);
    const double u2 = u * u;   // (0,1,2u,3u^2) M / dt
    Bv[0] = T.dt_inv * (-3.0 + 6.0 * u - 3.0 * u2) / 6.0; Bv[1] = T.dt_inv * (-12.0 * u + 9.0 * u2) / 6.0;
    Bv[2] = T.dt_inv * (3.0 + 6.0 * u - 9.0 * u2) / 6.0; Bv[3] = T.dt_inv * (3.0 * u2) / 6.0;
    basis_acc(u, T.dt_inv, Ba);
    if (T.r3) { p = r3_spline(T.r3 + 3 * i0, B); v = r3_spline(T.r3 + 3 * i0, Bv); a = r3_spline(T.r3 + 3 * i0, Ba); }
    So3Eval e;
    so3_spline_eval(T.so3 + 4 * i0, u, T.dt_inv, true, false, e);
    q = e.q;
    w = e.R * e.w_body;
  }
  if (pos) { pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z; }
  if (vel) { vel[3 * i] = v.x; vel[3 * i + 1] = v.y; vel[3 * i + 2] = v.z; }
  if (acc) { acc[3 * i] = a.x; acc[3 * i + 1] = a.y; acc[3 * i + 2] = a.z; }
  if (quat) { quat[4 * i] = q.x; quat[4 * i + 1] = q.y; quat[4 * i + 2] = q.z; quat[4 * i + 3] = q.w; }
  if (angvel) { angvel[3 * i] = w.x; angvel[3 * i + 1] = w.y; angvel[3 * i + 2] = w.z; }
  valid[i] = ok ? 1 : 0;
}

struct Mat34f { float m[12]; };

// pcl::transformPointCloud with a float 4x4 per scan.  The input is a PCL cloud (2 x float4 per point) or a packed batch; so is the output.
template <bool IN_PACKED, bool OUT_PACKED>
__global__ void __launch_bounds__(256) transform_kernel(const float4* __restrict__ in, int64_t n, int64_t pts_per_scan, const Mat34f* __restrict__ poses,
                                                        float4* __restrict__ out, int* __restrict__ mm) {
  ScanMinMax acc;
  acc.reset(-1);
  for (int64_t c0 = static_cast<int64_t>(blockIdx.x) * kProducerChunk; c0 < n; c0 += static_cast<int64_t>(gridDim.x) * kProducerChunk)
  for (int j = 0; j < kProducerChunk / 256; ++j) {
    const int64_t i = c0 + threadIdx.x + 256 * j;
    if (i >= n) break;
    float4 a; float inten;
    if (IN_PACKED) { a = __ldg(in + i); inten = a.w; }
    else { a = __ldg(in + 2 * i); inten = __ldg(in + 2 * i + 1).x; }
    const int64_t scan = i / pts_per_scan;
    const Mat34f& M = poses[scan];
    float4 o;
    o.x = M.m[0] * a.x + M.m[1] * a.y + M.m[2] * a.z + M.m[3];
    o.y = M.m[4] * a.x + M.m[5] * a.y + M.m[6] * a.z + M.m[7];
    o.z = M.m[8] * a.x + M.m[9] * a.y + M.m[10] * a.z + M.m[11];
    o.w = 1.f;
    if (OUT_PACKED) {
      out[i] = make_float4(o.x, o.y, o.z, inten);
      acc.add(mm, scan, o.x, o.y, o.z);
    } else {
      out[2 * i] = o;
      out[2 * i + 1] = make_float4(inten, 0.f, 0.f, 0.f);
    }
  }
  if (OUT_PACKED) acc.flush(mm);
}

// ABI cloud (x,y,z float at `stride` spacing, intensity at offset 16 when the record is the 32 B PCL one) -> packed batch + min/max slots
__global__ void __launch_bounds__(256) batch_import_kernel(const char* __restrict__ in, size_t stride, int64_t n, int64_t pts_per_scan,
                                                           float4* __restrict__ out, int* __restrict__ mm) {
  const bool vec = (stride % 16 == 0) && ((reinterpret_cast<size_t>(in) & 15) == 0);
  ScanMinMax acc;
  acc.reset(-1);
  for (int64_t c0 = static_cast<int64_t>(blockIdx.x) * kProducerChunk; c0 < n; c0 += static_cast<int64_t>(gridDim.x) * kProducerChunk)
  for (int j = 0; j < kProducerChunk / 256; ++j) {
    const int64_t i = c0 + threadIdx.x + 256 * j;
    if (i >= n) break;
    float x, y, z, w = 0.f;
    if (vec) { const float4 v = __ldg(reinterpret_cast<const float4*>(in + i * stride)); x = v.x; y = v.y; z = v.z; }
    else { const float* p = reinterpret_cast<const float*>(in + i * stride); x = p[0]; y = p[1]; z = p[2]; }
    if (stride >= 20) w = *reinterpret_cast<const float*>(in + i * stride + 16);
    out[i] = make_float4(x, y, z, w);
    acc.add(mm, i / pts_per_scan, x, y, z);
  }
  acc.flush(mm);
}
// packed batch -> pcl::PointXYZI records (x, y, z, 1, intensity, 0, 0, 0)
__global__ void __launch_bounds__(256) batch_export_kernel(const float4* __restrict__ in, int64_t n, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(in + i);
    out[2 * i] = make_float4(v.x, v.y, v.z, 1.f);
    out[2 * i + 1] = make_float4(v.w, 0.f, 0.f, 0.f);
  }
}

lvi_scan_batch* batch_alloc(lvi_ctx* ctx, int32_t n_scans, int64_t pts_per_scan, int64_t n) {
  auto b = std::unique_ptr<lvi_scan_batch>(new lvi_scan_batch());
  b->ctx = ctx; b->n_scans = n_scans; b->pts_per_scan = pts_per_scan; b->n = n;
  b->pts.alloc(static_cast<size_t>(n));
  b->mm.alloc(static_cast<size_t>(n_scans) * 6);
  LVI_LAUNCH(ctx, scan_minmax_init_kernel, (n_scans * 6 + 255) / 256, 256, 0, b->mm.p, n_scans);
  return b.release();
}
lvi_scan_batch* batch_import_xyzi(lvi_ctx* ctx, const void* xyz_d, size_t stride, int64_t n, int64_t pts_per_scan) {
  LVI_REQUIRE(n > 0 && n < 2147483647LL, LVI_ERR_INVALID, "point count must be in (0, 2^31)");
  LVI_REQUIRE(stride >= 12 && stride % 4 == 0, LVI_ERR_INVALID, "stride must be a multiple of 4, >= 12");
  LVI_REQUIRE(pts_per_scan > 0, LVI_ERR_INVALID, "pts_per_scan must be positive");
  const int64_t slots = (n + pts_per_scan - 1) / pts_per_scan;
  LVI_REQUIRE(slots < (1 << 24), LVI_ERR_INVALID, "too many scans in one batch");
  auto b = std::unique_ptr<lvi_scan_batch>(batch_alloc(ctx, static_cast<int32_t>(slots), pts_per_scan, n));
  LVI_LAUNCH(ctx, batch_import_kernel, grid_for(n, kProducerChunk, ctx->sm_count, 8), 256, 0, static_cast<const char*>(xyz_d), stride, n, pts_per_scan, b->pts.p, b->mm.p);
  return b.release();
}

static void undistort_device(lvi_ctx* ctx, const lvi_problem_desc* d, const lvi_point_xyzit* raw_d, int n_scans, int64_t pts_per_scan,
                             const double* target_time_h, int correct_position, void* out_d, int* n_bad_targets, int* mm_packed = nullptr) {
  LVI_REQUIRE(d->r3_knots && d->so3_knots && d->n_knots >= 4, LVI_ERR_INVALID, "lvi_undistort: trajectory needs both splines and >= 4 knots");
  LVI_REQUIRE(n_scans > 0 && pts_per_scan > 0, LVI_ERR_INVALID, "lvi_undistort: empty batch");
  cudaStream_t st = ctx->stream;
  const int n = d->n_knots;
  DBuf<double> r3(3 * static_cast<size_t>(n)), so3(4 * static_cast<size_t>(n)), tt(n_scans);
  r3.upload(d->r3_knots, r3.n, st); so3.upload(d->so3_knots, so3.n, st); tt.upload(target_time_h, n_scans, st);
  TrajView T;
  T.t0 = d->t0; T.dt = d->dt; T.dt_inv = 1.0 / d->dt; T.t_max = d->t0 + (n - 3) * d->dt; T.toff = d->lidar_toff; T.n_knots = n;
  DBuf<double> hlog(3 * static_cast<size_t>(n));
  LVI_LAUNCH(ctx, so3_log_kernel, (n + 127) / 128, 128, 0, so3.p, n, hlog.p);
  T.r3 = r3.p; T.so3 = so3.p; T.hlog = hlog.p;
  DBuf<double> hax(4 * static_cast<size_t>(n));
  LVI_LAUNCH(ctx, so3_axis_kernel, (n + 127) / 128, 128, 0, hlog.p, n, hax.p);
  static const double ident[4] = {0, 0, 0, 1}, zero[3] = {0, 0, 0};
  const double* lq = d->lidar_q ? d->lidar_q : ident; const double* lp = d->lidar_p ? d->lidar_p : zero;
  T.qL = q4(lq[0], lq[1], lq[2], lq[3]); T.pL = v3(lp[0], lp[1], lp[2]);
  DBuf<TargetPose> tp(n_scans);
  LVI_LAUNCH(ctx, undistort_target_kernel, (n_scans + 127) / 128, 128, 0, T, tt.p, n_scans, tp.p);
  DeskewConst K;
  {
    const M3 RL = qmat(T.qL);
    for (int k = 0; k < 9; ++k) K.RL[k] = RL.m[k];
    K.pL[0] = T.pL.x; K.pL[1] = T.pL.y; K.pL[2] = T.pL.z;
  }
  const int64_t np = static_cast<int64_t>(n_scans) * pts_per_scan;
  if (mm_packed)
    LVI_LAUNCH(ctx, undistort_kernel<true>, grid_for(np, kProducerChunk, ctx->sm_count, 8), 256, 0, T, K, hax.p, raw_d, np, pts_per_scan, tp.p, correct_position,
               static_cast<float4*>(out_d), mm_packed);
  else
    LVI_LAUNCH(ctx, undistort_kernel<false>, grid_for(np, kProducerChunk, ctx->sm_count, 8), 256, 0, T, K, hax.p, raw_d, np, pts_per_scan, tp.p, correct_position,
               static_cast<float4*>(out_d), static_cast<int*>(nullptr));
  std::vector<TargetPose> h(n_scans);
  tp.download(h.data(), n_scans, st);
  LVI_CUDA(cudaStreamSynchronize(st));
  int bad = 0;
  for (const auto& t : h) bad += t.ok ? 0 : 1;
  if (n_bad_targets) *n_bad_targets = bad;
}

static void transform_device(lvi_ctx* ctx, const void* in_d, int n_scans, int64_t pts_per_scan, const double* poses_h, void* out_d,
                             bool packed = false, int* mm_packed = nullptr) {
  LVI_REQUIRE(n_scans > 0 && pts_per_scan > 0, LVI_ERR_INVALID, "lvi_transform_scans: empty batch");
  std::vector<Mat34f> hm(n_scans);
  for (int s = 0; s < n_scans; ++s)
    for (int k = 0; k < 12; ++k) hm[s].m[k] = static_cast<float>(poses_h[16 * s + k]);  // odom_data.pose double -> float
  DBuf<Mat34f> pm(n_scans);
  pm.upload(hm.data(), n_scans, ctx->stream);
  const int64_t np = static_cast<int64_t>(n_scans) * pts_per_scan;
  if (packed)
    LVI_LAUNCH(ctx, (transform_kernel<true, true>), grid_for(np, kProducerChunk, ctx->sm_count, 8), 256, 0, static_cast<const float4*>(in_d), np, pts_per_scan, pm.p,
               static_cast<float4*>(out_d), mm_packed);
  else
    LVI_LAUNCH(ctx, (transform_kernel<false, false>), grid_for(np, kProducerChunk, ctx->sm_count, 8), 256, 0, static_cast<const float4*>(in_d), np, pts_per_scan, pm.p,
               static_cast<float4*>(out_d), static_cast<int*>(nullptr));
  LVI_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_undistort_d(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw_d, int32_t n_scans, int64_t pts_per_scan,
                    const double* target_time, int correct_position, void* out_xyzi_d, int32_t* n_bad_targets) {
  return guarded([&] {
    LVI_REQUIRE(ctx && traj && scans_raw_d && target_time && out_xyzi_d, LVI_ERR_INVALID, "lvi_undistort_d: null argument");
    activate(ctx);
    undistort_device(ctx, traj, scans_raw_d, n_scans, pts_per_scan, target_time, correct_position, out_xyzi_d, n_bad_targets);
  });
}

int lvi_undistort(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw, int32_t n_scans, int64_t pts_per_scan,
                  const double* target_time, int correct_position, void* out_xyzi, int32_t* n_bad_targets) {
  return guarded([&] {
    LVI_REQUIRE(ctx && traj && scans_raw && target_time && out_xyzi, LVI_ERR_INVALID, "lvi_undistort: null argument");
    activate(ctx);
    const size_t np = static_cast<size_t>(n_scans) * pts_per_scan;
    DBuf<lvi_point_xyzit> raw_d(np);
    DBuf<char> out_d(np * 32);
    LVI_CUDA(cudaMemcpyAsync(raw_d.p, scans_raw, np * sizeof(lvi_point_xyzit), cudaMemcpyHostToDevice, ctx->stream));
    undistort_device(ctx, traj, raw_d.p, n_scans, pts_per_scan, target_time, correct_position, out_d.p, n_bad_targets);
    LVI_CUDA(cudaMemcpyAsync(out_xyzi, out_d.p, np * 32, cudaMemcpyDeviceToHost, ctx->stream));
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

// ---- scan batches (packed, HBM-resident) --------------------------------------------------------------------------------------
int lvi_scan_batch_undistort_d(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw_d, int32_t n_scans, int64_t pts_per_scan,
                               const double* target_time, int correct_position, lvi_scan_batch** out, int32_t* n_bad_targets) {
  return guarded([&] {
    LVI_REQUIRE(ctx && traj && scans_raw_d && target_time && out, LVI_ERR_INVALID, "lvi_scan_batch_undistort_d: null argument");
    activate(ctx);
    LVI_REQUIRE(n_scans > 0 && pts_per_scan > 0, LVI_ERR_INVALID, "lvi_scan_batch_undistort_d: empty batch");
    auto b = std::unique_ptr<lvi_scan_batch>(batch_alloc(ctx, n_scans, pts_per_scan, static_cast<int64_t>(n_scans) * pts_per_scan));
    undistort_device(ctx, traj, scans_raw_d, n_scans, pts_per_scan, target_time, correct_position, b->pts.p, n_bad_targets, b->mm.p);
    *out = b.release();
  });
}

int lvi_scan_batch_transform(lvi_ctx* ctx, const lvi_scan_batch* in, const double* poses, lvi_scan_batch** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && in && poses && out, LVI_ERR_INVALID, "lvi_scan_batch_transform: null argument");
    activate(ctx);
    auto b = std::unique_ptr<lvi_scan_batch>(batch_alloc(ctx, in->n_scans, in->pts_per_scan, in->n));
    transform_device(ctx, in->pts.p, in->n_scans, in->pts_per_scan, poses, b->pts.p, true, b->mm.p);
    *out = b.release();
  });
}

int lvi_scan_batch_from_xyzi_d(lvi_ctx* ctx, const void* xyzi_d, size_t stride_bytes, int32_t n_scans, int64_t pts_per_scan, lvi_scan_batch** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && xyzi_d && out && n_scans > 0, LVI_ERR_INVALID, "lvi_scan_batch_from_xyzi_d: bad argument");
    activate(ctx);
    *out = batch_import_xyzi(ctx, xyzi_d, stride_bytes, static_cast<int64_t>(n_scans) * pts_per_scan, pts_per_scan);
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int lvi_scan_batch_export_xyzi(lvi_ctx* ctx, const lvi_scan_batch* b, void* out_xyzi, int out_is_device) {
  return guarded([&] {
    LVI_REQUIRE(ctx && b && out_xyzi, LVI_ERR_INVALID, "lvi_scan_batch_export_xyzi: null argument");
    activate(ctx);
    const size_t bytes = static_cast<size_t>(b->n) * 32;
    DBuf<char> tmp;
    float4* dst = static_cast<float4*>(out_xyzi);
    if (!out_is_device) { tmp.alloc(bytes); dst = reinterpret_cast<float4*>(tmp.p); }
    LVI_LAUNCH(ctx, batch_export_kernel, grid_for(b->n, 256, ctx->sm_count, 8), 256, 0, b->pts.p, b->n, dst);
    if (!out_is_device) LVI_CUDA(cudaMemcpyAsync(out_xyzi, tmp.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int lvi_scan_batch_destroy(lvi_scan_batch* b) {
  if (b) { cudaSetDevice(b->ctx->device); tl_stream = b->ctx->stream; delete b; }
  return LVI_OK;
}
int64_t lvi_scan_batch_num_points(const lvi_scan_batch* b) { return b ? b->n : 0; }
int32_t lvi_scan_batch_num_scans(const lvi_scan_batch* b) { return b ? b->n_scans : 0; }
/* the packed device buffer (float4 x,y,z,intensity per point) for callers that keep working on the device */
const void* lvi_scan_batch_points_d(const lvi_scan_batch* b) { return b ? b->pts.p : nullptr; }

int lvi_trajectory_evaluate(lvi_ctx* ctx, const lvi_problem_desc* d, const double* t, int64_t n, double* pos, double* quat, uint8_t* valid) {
  return guarded([&] {
    LVI_REQUIRE(ctx && d && t && pos && quat && valid && n > 0, LVI_ERR_INVALID, "lvi_trajectory_evaluate: bad argument");
    LVI_REQUIRE(d->r3_knots && d->so3_knots && d->n_knots >= 4, LVI_ERR_INVALID, "lvi_trajectory_evaluate: trajectory needs both splines and >= 4 knots");
    activate(ctx);
    cudaStream_t st = ctx->stream;
    const int nk = d->n_knots;
    DBuf<double> r3(3 * static_cast<size_t>(nk)), so3(4 * static_cast<size_t>(nk)), td(n), pd(3 * static_cast<size_t>(n)), qd(4 * static_cast<size_t>(n));
    DBuf<unsigned char> vd(n);
    r3.upload(d->r3_knots, r3.n, st); so3.upload(d->so3_knots, so3.n, st); td.upload(t, n, st);
    TrajView T;
    T.t0 = d->t0; T.dt = d->dt; T.dt_inv = 1.0 / d->dt; T.t_max = d->t0 + (nk - 3) * d->dt; T.toff = 0.0; T.n_knots = nk;
    DBuf<double> hlog(3 * static_cast<size_t>(nk));
    LVI_LAUNCH(ctx, so3_log_kernel, (nk + 127) / 128, 128, 0, so3.p, nk, hlog.p);
    T.r3 = r3.p; T.so3 = so3.p; T.hlog = hlog.p; T.qL = q4(0, 0, 0, 1); T.pL = v3(0, 0, 0);
    LVI_LAUNCH(ctx, traj_eval_kernel, static_cast<int>((n + 127) / 128), 128, 0, T, td.p, n, pd.p, qd.p, vd.p);
    pd.download(pos, pd.n, st); qd.download(quat, qd.n, st); vd.download(valid, n, st);
    LVI_CUDA(cudaStreamSynchronize(st));
  });
}

int lvi_trajectory_evaluate_full(lvi_ctx* ctx, const lvi_problem_desc* d, const double* t, int64_t n, double* pos, double* vel, double* acc, double* quat,
                                 double* angvel, uint8_t* valid) {
  return guarded([&] {
    LVI_REQUIRE(ctx && d && t && valid && n > 0, LVI_ERR_INVALID, "lvi_trajectory_evaluate_full: bad argument");
    LVI_REQUIRE(d->so3_knots && d->n_knots >= 4, LVI_ERR_INVALID, "lvi_trajectory_evaluate_full: trajectory needs the SO3 spline and >= 4 knots");
    activate(ctx);
    cudaStream_t st = ctx->stream;
    const int nk = d->n_knots;
    DBuf<double> r3(d->r3_knots ? 3 * static_cast<size_t>(nk) : 0), so3(4 * static_cast<size_t>(nk)), td(n);
    DBuf<double> pd(3 * static_cast<size_t>(n)), vd(3 * static_cast<size_t>(n)), ad(3 * static_cast<size_t>(n)), qd(4 * static_cast<size_t>(n)), wd(3 * static_cast<size_t>(n));
    DBuf<unsigned char> okd(n);
    if (d->r3_knots) r3.upload(d->r3_knots, r3.n, st);
    so3.upload(d->so3_knots, so3.n, st); td.upload(t, n, st);
    TrajView T;
    T.t0 = d->t0; T.dt = d->dt; T.dt_inv = 1.0 / d->dt; T.t_max = d->t0 + (nk - 3) * d->dt; T.toff = 0.0; T.n_knots = nk;
    T.r3 = d->r3_knots ? r3.p : nullptr; T.so3 = so3.p; T.hlog = nullptr; T.qL = q4(0, 0, 0, 1); T.pL = v3(0, 0, 0);
    LVI_LAUNCH(ctx, traj_eval_full_kernel, static_cast<int>((n + 127) / 128), 128, 0, T, td.p, n, pd.p, vd.p, ad.p, qd.p, wd.p, okd.p);
    if (pos) pd.download(pos, pd.n, st);
    if (vel) vd.download(vel, vd.n, st);
    if (acc) ad.download(acc, ad.n, st);
    if (quat) qd.download(quat, qd.n, st);
    if (angvel) wd.download(angvel, wd.n, st);
    okd.download(valid, n, st);
    LVI_CUDA(cudaStreamSynchronize(st));
  });
}

int lvi_transform_scans_d(lvi_ctx* ctx, const void* scans_xyzi_d, int32_t n_scans, int64_t pts_per_scan, const double* poses, void* out_xyzi_d) {
  return guarded([&] {
    LVI_REQUIRE(ctx && scans_xyzi_d && poses && out_xyzi_d, LVI_ERR_INVALID, "lvi_transform_scans_d: null argument");
    activate(ctx);
    transform_device(ctx, scans_xyzi_d, n_scans, pts_per_scan, poses, out_xyzi_d);
  });
}

int lvi_transform_scans(lvi_ctx* ctx, const void* scans_xyzi, int32_t n_scans, int64_t pts_per_scan, const double* poses, void* out_xyzi) {
  return guarded([&] {
    LVI_REQUIRE(ctx && scans_xyzi && poses && out_xyzi, LVI_ERR_INVALID, "lvi_transform_scans: null argument");
    activate(ctx);
    const size_t bytes = static_cast<size_t>(n_scans) * pts_per_scan * 32;
    DBuf<char> in_d(bytes), out_d(bytes);
    LVI_CUDA(cudaMemcpyAsync(in_d.p, scans_xyzi, bytes, cudaMemcpyHostToDevice, ctx->stream));
    transform_device(ctx, in_d.p, n_scans, pts_per_scan, poses, out_d.p);
    LVI_CUDA(cudaMemcpyAsync(out_xyzi, out_d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

}  // extern "C"

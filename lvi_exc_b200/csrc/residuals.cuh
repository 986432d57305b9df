// residuals.cuh — residuals and ANALYTIC Jacobians of the six on-path Kontiki measurements (SURVEY §8 a-5 … a-10).
//
// Replaces the templated Residual functors evaluated with ceres::Jet<double,4> by DynamicAutoDiffCostFunction
// (4..26 functor passes per residual, each re-evaluating the spline with heap allocations):
//   GyroscopeMeasurement        K/measurements/gyroscope_measurement.h:36-38   + K/sensors/imu.h:87-91, constant_bias_imu.h:57-61
//   AccelerometerMeasurement    K/measurements/accelerometer_measurement.h:37-39 + K/sensors/imu.h:61-70,95-101
//   LiDARSurfelPoint            K/measurements/lidar_surfel_point.h:31-74
//   StaticRsCameraMeasurement   K/measurements/static_rscamera_measurement.h:16-60 + K/sensors/pinhole_camera.h:96-124,217-238
//   CameraSurfelLandmark        K/measurements/camera_surfel_landmark.h:29-91 (rho read as a constant, Q4)
//   OrientationMeasurement      K/measurements/orientation_measurement.h:30-33
// Jacobians are with respect to Ceres' tangent coordinates (EigenQuaternionParameterization: q+ = [sinc|d| d, cos|d|] * q,
// i.e. a LEFT rotation by the vector 2d), one row block per residual, columns in the fixed per-type layout below.
// The loss correction (ceres Corrector, Huber => rho'' <= 0 branch) is applied by the caller.
#pragma once
#include "spline_math.cuh"

namespace lvi {

// sensor parameter vector layout (ambient)
enum { SENS_LQ = 0, SENS_LP = 4, SENS_CQ = 7, SENS_CP = 11, SENS_G = 14, SENS_BA = 16, SENS_BG = 19, SENS_N = 22 };
// tangent blocks of the sensors (index into ProblemView::pos_sens)
enum { TB_LQ = 0, TB_LP, TB_CQ, TB_CP, TB_G, TB_BA, TB_BG, TB_COUNT };

enum { RT_GYRO = 0, RT_ACCEL, RT_SURFEL, RT_CAM, RT_CAMSURF, RT_ORIENT, RT_COUNT };
// column layouts (tangent dims) per residual type
//   gyro    : so3[4x3] | bg[3]                                             = 15
//   accel   : r3[4x3] | so3[4x3] | g_roll g_pitch | ba[3]                  = 29
//   surfel  : map: r3[12] so3[12] | k: r3[12] so3[12] | lidar q[3] p[3]    = 54
//   cam     : ref: r3[12] so3[12] | obs: r3[12] so3[12] | cam q[3] p[3] | rho = 55
//   camsurf : map: r3[12] so3[12] | ref: r3[12] so3[12] | cam q p | lidar q p = 60
//   orient  : so3[12]
#define LVI_MAX_COLS 60
#define LVI_MAX_ROWS 3
LVI_HD int rt_rows(int t) { return t == RT_GYRO || t == RT_ACCEL ? 3 : (t == RT_CAM ? 2 : 1); }
LVI_HD int rt_cols(int t) { return t == RT_GYRO ? 15 : t == RT_ACCEL ? 29 : t == RT_SURFEL ? 54 : t == RT_CAM ? 55 : t == RT_CAMSURF ? 60 : 12; }

struct ResTable {   // one residual table, SoA, device (or host) pointers
  int n;
  int lo, hi;       // index range this rank evaluates (data-parallel sharding by time chunk); [0, n) on one GPU
  int active;       // 0: every parameter block constant -> contributes only to the fixed cost
  const int* i0a; const double* ua;   // first evaluation (map time / ref time / the only one)
  const int* i0b; const double* ub;   // second evaluation (point time / obs time)
  const double* v;                    // measurement vector(s): gyro w[3], accel a[3], surfel p[3], cam uv_ref[2]+uv_obs[2], camsurf uv[2], orient q[4]
  const int* ia;                      // plane id (surfel, camsurf) / landmark id (cam)
  const int* ib;                      // landmark id (camsurf)
  const int* perm;                    // evaluation order of the normal-equation kernel (nullptr: table order)
  const double* weight; const double* huber;
};

struct ProblemView {
  double dt_inv;
  int n_knots;
  const double* r3;     // [n*3]
  const double* so3;    // [n*4]
  const double* sens;   // [SENS_N]
  const double* rho;    // [n_landmarks]
  const double* planes; // [n_planes*3]
  double fx, fy, cx, cy;
  double dist[5];       // k1 k2 p1 p2 k3
  int do_dist;          // the reference's test: |k1|, |k2| or |p1| above 1e-5 (pinhole_camera.h:78)
  int has_r3;
  ResTable tab[RT_COUNT];
  // column positions in the linear system (-1: constant)
  const int* pos_r3; const int* pos_so3; const int* pos_rho;
  int pos_sens[TB_COUNT];
};

struct ResOut {
  double r[LVI_MAX_ROWS];
  double J[LVI_MAX_ROWS][LVI_MAX_COLS];
};

// helpers: write a 1x3 row block g^T * A (optionally scaled) into J[row][col..col+2]
LVI_HD void put_row(double* Jrow, int col, V3 g) { Jrow[col] = g.x; Jrow[col + 1] = g.y; Jrow[col + 2] = g.z; }

// d(r)/d(cp) blocks for a pose evaluation: positions via basis weights, orientation via C_j (x2 for Ceres' half-angle delta)
LVI_HD void scatter_pose(double* Jrow, int col0, V3 g_p, V3 g_phi, const double Bp[4], const So3Eval& e, bool has_r3) {
  for (int j = 0; j < 4; ++j) {
    if (has_r3) put_row(Jrow, col0 + 3 * j, Bp[j] * g_p);
    put_row(Jrow, col0 + 12 + 3 * j, 2.0 * vecmat(g_phi, e.C[j]));
  }
}

LVI_HD void eval_gyro(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_GYRO];
  So3Eval e;
  so3_spline_eval(P.so3 + 4 * T.i0a[i], T.ua[i], P.dt_inv, true, jac, e);
  const double w = T.weight[i];
  const double* bg = P.sens + SENS_BG;
  const double m[3] = {e.w_body.x + bg[0], e.w_body.y + bg[1], e.w_body.z + bg[2]};
  for (int k = 0; k < 3; ++k) o.r[k] = w * (T.v[3 * i + k] - m[k]);
  if (!jac) return;
  for (int k = 0; k < 3; ++k) {
    for (int j = 0; j < 4; ++j) for (int c = 0; c < 3; ++c) o.J[k][3 * j + c] = -2.0 * w * e.N[j].m[k * 3 + c];
    for (int c = 0; c < 3; ++c) o.J[k][12 + c] = (c == k) ? -w : 0.0;
  }
}

LVI_HD void eval_accel(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_ACCEL];
  const int i0 = T.i0a[i];
  So3Eval e;
  so3_spline_eval(P.so3 + 4 * i0, T.ua[i], P.dt_inv, false, jac, e);
  double Ba[4];
  basis_acc(T.ua[i], P.dt_inv, Ba);
  V3 a = v3(0, 0, 0);
  if (P.has_r3) a = r3_spline(P.r3 + 3 * i0, Ba);
  const double roll = P.sens[SENS_G], pitch = P.sens[SENS_G + 1];
  const double G = -9.79;  // K/sensors/imu.h:25
  const double cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch);
  const V3 g = v3(-sp * cr * G, sr * G, -cr * cp * G);  // refined_gravity, imu.h:61-70
  const V3 v = a + g;
  const V3 m = mulT(e.R, v);
  const double w = T.weight[i];
  const double* ba = P.sens + SENS_BA;
  o.r[0] = w * (T.v[3 * i] - (m.x + ba[0]));
  o.r[1] = w * (T.v[3 * i + 1] - (m.y + ba[1]));
  o.r[2] = w * (T.v[3 * i + 2] - (m.z + ba[2]));
  if (!jac) return;
  const M3 Rt = transpose(e.R);
  const M3 dphi = Rt * skew(v);  // d(R^T v)/dphi
  const V3 dg_dr = v3(sp * sr * G, cr * G, sr * cp * G), dg_dp = v3(-cp * cr * G, 0.0, cr * sp * G);
  const V3 mr = mulT(e.R, dg_dr), mp = mulT(e.R, dg_dp);
  for (int k = 0; k < 3; ++k) {
    const V3 rowRt = v3(Rt.m[k * 3], Rt.m[k * 3 + 1], Rt.m[k * 3 + 2]);
    const V3 rowPhi = v3(dphi.m[k * 3], dphi.m[k * 3 + 1], dphi.m[k * 3 + 2]);
    for (int j = 0; j < 4; ++j) {
      put_row(o.J[k], 3 * j, (-w * Ba[j]) * rowRt);
      put_row(o.J[k], 12 + 3 * j, (-2.0 * w) * vecmat(rowPhi, e.C[j]));
    }
    o.J[k][24] = -w * (k == 0 ? mr.x : k == 1 ? mr.y : mr.z);
    o.J[k][25] = -w * (k == 0 ? mp.x : k == 1 ? mp.y : mp.z);
    for (int c = 0; c < 3; ++c) o.J[k][26 + c] = (c == k) ? -w : 0.0;
  }
}

// shared by surfel and camsurf: p_M = R_LI^T (R_0^T (R_k p_I + p_k - p_0) - p_LI), r = w (n.p_M - d)
struct MapChain { V3 p_M, y, p_tmp; M3 R0, Rk, RLI; V3 gn; /* w * n^T R_LI^T R_0^T */ double w_eff; V3 n; };

LVI_HD void eval_surfel(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_SURFEL];
  const int i0m = T.i0a[i], i0k = T.i0b[i];
  So3Eval e0, ek;
  so3_spline_eval(P.so3 + 4 * i0m, T.ua[i], P.dt_inv, false, jac, e0);
  so3_spline_eval(P.so3 + 4 * i0k, T.ub[i], P.dt_inv, false, jac, ek);
  double B0[4], Bk[4];
  basis_pos(T.ua[i], B0); basis_pos(T.ub[i], Bk);
  const V3 p0 = r3_spline(P.r3 + 3 * i0m, B0), pk = r3_spline(P.r3 + 3 * i0k, Bk);
  const Q4 qL = q4(P.sens[SENS_LQ], P.sens[SENS_LQ + 1], P.sens[SENS_LQ + 2], P.sens[SENS_LQ + 3]);
  const V3 pLI = v3(P.sens[SENS_LP], P.sens[SENS_LP + 1], P.sens[SENS_LP + 2]);
  const V3 pL = v3(T.v[3 * i], T.v[3 * i + 1], T.v[3 * i + 2]);
  const V3 RpL = qrot(qL, pL);
  const V3 p_I = RpL + pLI;
  const V3 Rk_pI = qrot(ek.q, p_I);
  const V3 y = Rk_pI + pk - p0;
  const V3 p_tmp = qrot(qconj(e0.q), y);
  const V3 vv = p_tmp - pLI;
  const V3 p_M = qrot(qconj(qL), vv);
  const double* Pi = P.planes + 3 * T.ia[i];
  const double d = sqrt(Pi[0] * Pi[0] + Pi[1] * Pi[1] + Pi[2] * Pi[2]);  // Q9: plane through the origin is singular
  const V3 n = v3(Pi[0] / d, Pi[1] / d, Pi[2] / d);
  const double w = T.weight[i];
  o.r[0] = w * (n.x * p_M.x + n.y * p_M.y + n.z * p_M.z - d);
  if (!jac) return;
  const M3 RLI = qmat(qL);
  const V3 gL = w * (RLI * n);              // row: w n^T R_LI^T            (as a column: R_LI n)
  const V3 g0 = mulT(transpose(e0.R), gL);  // w n^T R_LI^T R_0^T  = (R_0 R_LI n)^T
  // map-time evaluation (columns 0..23): d/dp0 = -g0 ; d/dphi0 = gL^T R_0^T... = g0^T [y]x
  scatter_pose(o.J[0], 0, -1.0 * g0, vecmat(g0, skew(y)), B0, e0, true);
  // point-time evaluation (columns 24..47): d/dpk = g0 ; d/dphik = -g0^T [R_k p_I]x
  scatter_pose(o.J[0], 24, g0, -1.0 * vecmat(g0, skew(Rk_pI)), Bk, ek, true);
  // lidar extrinsics: dphi_L = -g0^T R_k [R_LI p_L]x + gL^T [p_tmp - p_LI]x ; dp_L = g0^T R_k - gL^T
  const V3 g0Rk = vecmat(g0, ek.R);
  put_row(o.J[0], 48, 2.0 * (vecmat(gL, skew(vv)) - vecmat(g0Rk, skew(RpL))));
  put_row(o.J[0], 51, g0Rk - gL);
}

// ---- radial-tangential distortion of the pinhole camera (K/sensors/pinhole_camera.h:131-240) --------------------------------------------
// d(p_u) with p_d = p_u + d (:199-215)
LVI_HD void cam_distortion(const double k[5], double x, double y, double& dx, double& dy) {
  const double x2 = x * x, y2 = y * y, xy = x * y, r2 = x2 + y2;
  const double rad = k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2;
  dx = x * rad + 2.0 * k[2] * xy + k[3] * (r2 + 2.0 * x2);
  dy = y * rad + 2.0 * k[3] * xy + k[2] * (r2 + 2.0 * y2);
}
// M = I + d d / d p_u (the Jacobian of p_u -> p_d), row-major 2x2
LVI_HD void cam_distortion_jac(const double k[5], double x, double y, double M[4]) {
  const double x2 = x * x, y2 = y * y, r2 = x2 + y2;
  const double rad = k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2;
  const double drad = k[0] + 2.0 * k[1] * r2 + 3.0 * k[4] * r2 * r2;   // d rad / d r2
  M[0] = 1.0 + rad + 2.0 * x2 * drad + 2.0 * k[2] * y + 6.0 * k[3] * x;
  M[1] = 2.0 * x * y * drad + 2.0 * k[2] * x + 2.0 * k[3] * y;
  M[2] = 2.0 * x * y * drad + 2.0 * k[3] * y + 2.0 * k[2] * x;
  M[3] = 1.0 + rad + 2.0 * y2 * drad + 2.0 * k[3] * x + 6.0 * k[2] * y;
}
// Unproject (:112-124, liftProjective :131-190): K^-1 (u, v, 1), then -- with distortion -- the recursive inverse model, 8 steps, no early exit
LVI_HD V3 cam_unproject(const ProblemView& P, double u, double v) {
  const double mx_d = (u - P.cx) / P.fx, my_d = (v - P.cy) / P.fy;
  if (!P.do_dist) return v3(mx_d, my_d, 1.0);
  double dx, dy;
  cam_distortion(P.dist, mx_d, my_d, dx, dy);
  double mx_u = mx_d - dx, my_u = my_d - dy;
  for (int it = 1; it < 8; ++it) {
    cam_distortion(P.dist, mx_u, my_u, dx, dy);
    mx_u = mx_d - dx; my_u = my_d - dy;
  }
  return v3(mx_u, my_u, 1.0);
}

LVI_HD void eval_camsurf(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_CAMSURF];
  const int i0m = T.i0a[i], i0k = T.i0b[i];
  So3Eval e0, ek;
  so3_spline_eval(P.so3 + 4 * i0m, T.ua[i], P.dt_inv, false, jac, e0);
  so3_spline_eval(P.so3 + 4 * i0k, T.ub[i], P.dt_inv, false, jac, ek);
  double B0[4], Bk[4];
  basis_pos(T.ua[i], B0); basis_pos(T.ub[i], Bk);
  const V3 p0 = r3_spline(P.r3 + 3 * i0m, B0), pk = r3_spline(P.r3 + 3 * i0k, Bk);
  const Q4 qL = q4(P.sens[SENS_LQ], P.sens[SENS_LQ + 1], P.sens[SENS_LQ + 2], P.sens[SENS_LQ + 3]);
  const Q4 qC = q4(P.sens[SENS_CQ], P.sens[SENS_CQ + 1], P.sens[SENS_CQ + 2], P.sens[SENS_CQ + 3]);
  const V3 pLI = v3(P.sens[SENS_LP], P.sens[SENS_LP + 1], P.sens[SENS_LP + 2]);
  const V3 pCI = v3(P.sens[SENS_CP], P.sens[SENS_CP + 1], P.sens[SENS_CP + 2]);
  const double s = 1.0 / (P.rho[T.ib[i]] + 1e-8);  // camera_surfel_landmark.h:57
  const V3 yh = s * cam_unproject(P, T.v[2 * i], T.v[2 * i + 1]);   // landmark position in the camera frame: Unproject(uv) / (rho + 1e-8)
  const V3 Ryh = qrot(qC, yh);
  const V3 p_I = Ryh + pCI;
  const V3 Rk_pI = qrot(ek.q, p_I);
  const V3 y = Rk_pI + pk - p0;
  const V3 p_tmp = qrot(qconj(e0.q), y);
  const V3 vv = p_tmp - pLI;
  const V3 p_M = qrot(qconj(qL), vv);
  const double* Pi = P.planes + 3 * T.ia[i];
  const double d = sqrt(Pi[0] * Pi[0] + Pi[1] * Pi[1] + Pi[2] * Pi[2]);
  const V3 n = v3(Pi[0] / d, Pi[1] / d, Pi[2] / d);
  const double w = T.weight[i];
  o.r[0] = w * (n.x * p_M.x + n.y * p_M.y + n.z * p_M.z - d);
  if (!jac) return;
  const M3 RLI = qmat(qL);
  const V3 gL = w * (RLI * n);
  const V3 g0 = mulT(transpose(e0.R), gL);
  scatter_pose(o.J[0], 0, -1.0 * g0, vecmat(g0, skew(y)), B0, e0, true);
  scatter_pose(o.J[0], 24, g0, -1.0 * vecmat(g0, skew(Rk_pI)), Bk, ek, true);
  const V3 g0Rk = vecmat(g0, ek.R);
  put_row(o.J[0], 48, -2.0 * vecmat(g0Rk, skew(Ryh)));  // camera q
  put_row(o.J[0], 51, g0Rk);                             // camera p
  put_row(o.J[0], 54, 2.0 * vecmat(gL, skew(vv)));       // lidar q
  put_row(o.J[0], 57, -1.0 * gL);                        // lidar p
}

LVI_HD void eval_cam(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_CAM];
  const int i0r = T.i0a[i], i0o = T.i0b[i];
  So3Eval er, eo;
  so3_spline_eval(P.so3 + 4 * i0r, T.ua[i], P.dt_inv, false, jac, er);
  so3_spline_eval(P.so3 + 4 * i0o, T.ub[i], P.dt_inv, false, jac, eo);
  double Br[4], Bo[4];
  basis_pos(T.ua[i], Br); basis_pos(T.ub[i], Bo);
  const V3 pr = r3_spline(P.r3 + 3 * i0r, Br), po = r3_spline(P.r3 + 3 * i0o, Bo);
  const Q4 qC = q4(P.sens[SENS_CQ], P.sens[SENS_CQ + 1], P.sens[SENS_CQ + 2], P.sens[SENS_CQ + 3]);
  const V3 pCI = v3(P.sens[SENS_CP], P.sens[SENS_CP + 1], P.sens[SENS_CP + 2]);
  const double rho = P.rho[T.ia[i]];
  const double* uv = T.v + 4 * i;  // uv_ref[2], uv_obs[2]
  const V3 yh = cam_unproject(P, uv[0], uv[1]);  // Unproject: K^-1 (u,v,1) (+ inverse distortion)
  // the reference goes through p_ct = q_CI^-1 (-p_CI), q_ct = q_CI^-1 (static_rscamera_measurement.h:42-55)
  const V3 p_ct = qrot(qconj(qC), -1.0 * pCI);
  const V3 X_ref = qrot(qC, yh - rho * p_ct);
  const V3 RrX = qrot(er.q, X_ref);
  const V3 X = RrX + rho * pr;
  const V3 Xd = X - rho * po;
  const V3 X_obs = qrot(qconj(eo.q), Xd);
  const V3 Xc = qrot(qconj(qC), X_obs) + rho * p_ct;
  const double z = 1e-32 + Xc.z;  // spaceToPlane (pinhole_camera.h:217-240)
  const double w = T.weight[i];
  double pdx = Xc.x / z, pdy = Xc.y / z;
  double M[4] = {1.0, 0.0, 0.0, 1.0};
  if (P.do_dist) {
    double dx, dy;
    cam_distortion(P.dist, pdx, pdy, dx, dy);
    if (jac) cam_distortion_jac(P.dist, pdx, pdy, M);
    pdx += dx; pdy += dy;
  }
  o.r[0] = w * (uv[2] - (P.fx * pdx + P.cx));
  o.r[1] = w * (uv[3] - (P.fy * pdy + P.cy));
  if (!jac) return;
  const M3 RC = qmat(qC);
  const V3 Ryh = qrot(qC, yh);
  const V3 vC = X_obs - rho * pCI;  // X_c = R_CI^T vC
  // G rows: -w * dproj/dXc
  V3 G0 = v3(-w * P.fx / z, 0.0, w * P.fx * Xc.x / (z * z));
  V3 G1 = v3(0.0, -w * P.fy / z, w * P.fy * Xc.y / (z * z));
  if (P.do_dist) {   // chain rule through p_d = p_u + d(p_u)
    const V3 a0 = v3(1.0 / z, 0.0, -Xc.x / (z * z)), a1 = v3(0.0, 1.0 / z, -Xc.y / (z * z));   // d p_u / d Xc
    G0 = (-w * P.fx) * (M[0] * a0 + M[1] * a1);
    G1 = (-w * P.fy) * (M[2] * a0 + M[3] * a1);
  }
  const V3 dXobs_drho = qrot(qconj(eo.q), qrot(er.q, pCI) + pr - po);
  const V3 dXc_drho = mulT(RC, dXobs_drho - pCI);
  for (int k = 0; k < 2; ++k) {
    const V3 G = k == 0 ? G0 : G1;
    const V3 gC = RC * G;                         // G^T R_CI^T as a column
    const V3 gO = mulT(transpose(eo.R), gC);      // G^T R_CI^T R_o^T  = (R_o R_CI G)
    // ref evaluation: dX_obs/dp_r = rho R_o^T ; dX_obs/dphi_r = -R_o^T [R_r X_ref]x
    scatter_pose(o.J[k], 0, rho * gO, -1.0 * vecmat(gO, skew(RrX)), Br, er, true);
    // obs evaluation: dX_obs/dp_o = -rho R_o^T ; dX_obs/dphi_o = R_o^T [X - rho p_o]x
    scatter_pose(o.J[k], 24, -rho * gO, vecmat(gO, skew(Xd)), Bo, eo, true);
    // camera extrinsics
    const V3 gOr = vecmat(gO, er.R);              // G^T R_CI^T R_o^T R_r
    put_row(o.J[k], 48, 2.0 * (vecmat(gC, skew(vC)) - vecmat(gOr, skew(Ryh))));
    put_row(o.J[k], 51, rho * (gOr - gC));
    o.J[k][54] = dot(G, dXc_drho);
  }
}

LVI_HD void eval_orient(const ProblemView& P, int i, bool jac, ResOut& o) {
  const ResTable& T = P.tab[RT_ORIENT];
  So3Eval e;
  so3_spline_eval(P.so3 + 4 * T.i0a[i], T.ua[i], P.dt_inv, false, jac, e);
  const Q4 qm = q4(T.v[4 * i], T.v[4 * i + 1], T.v[4 * i + 2], T.v[4 * i + 3]);
  const Q4 dq = qmul(qm, qconj(e.q));
  const double vn = sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
  const double w = T.weight[i];
  o.r[0] = w * 2.0 * atan2(vn, fabs(dq.w));  // Eigen angularDistance
  if (!jac) return;
  const double sg = dq.w < 0 ? -1.0 : 1.0;
  const V3 ax = vn > 0 ? v3(sg * dq.x / vn, sg * dq.y / vn, sg * dq.z / vn) : v3(0, 0, 0);
  for (int j = 0; j < 4; ++j) put_row(o.J[0], 3 * j, (-2.0 * w) * vecmat(ax, e.C[j]));
}

template <int TYPE> LVI_HD void eval_residual(const ProblemView& P, int i, bool jac, ResOut& o) {
  if (TYPE == RT_GYRO) eval_gyro(P, i, jac, o);
  else if (TYPE == RT_ACCEL) eval_accel(P, i, jac, o);
  else if (TYPE == RT_SURFEL) eval_surfel(P, i, jac, o);
  else if (TYPE == RT_CAM) eval_cam(P, i, jac, o);
  else if (TYPE == RT_CAMSURF) eval_camsurf(P, i, jac, o);
  else eval_orient(P, i, jac, o);
}

// linear-system position of column c of residual i of type TYPE (-1: constant block)
template <int TYPE> LVI_HD int col_pos(const ProblemView& P, int i, int c) {
  const ResTable& T = P.tab[TYPE];
  if (TYPE == RT_GYRO) {
    if (c < 12) { const int p = P.pos_so3[T.i0a[i] + c / 3]; return p < 0 ? -1 : p + c % 3; }
    return P.pos_sens[TB_BG] < 0 ? -1 : P.pos_sens[TB_BG] + (c - 12);
  }
  if (TYPE == RT_ORIENT) { const int p = P.pos_so3[T.i0a[i] + c / 3]; return p < 0 ? -1 : p + c % 3; }
  if (TYPE == RT_ACCEL) {
    if (c < 12) { const int p = P.pos_r3[T.i0a[i] + c / 3]; return p < 0 ? -1 : p + c % 3; }
    if (c < 24) { const int p = P.pos_so3[T.i0a[i] + (c - 12) / 3]; return p < 0 ? -1 : p + c % 3; }
    if (c < 26) return P.pos_sens[TB_G] < 0 ? -1 : P.pos_sens[TB_G] + (c - 24);
    return P.pos_sens[TB_BA] < 0 ? -1 : P.pos_sens[TB_BA] + (c - 26);
  }
  // two-evaluation types: columns 0..23 first evaluation, 24..47 second
  if (c < 48) {
    const int i0 = c < 24 ? T.i0a[i] : T.i0b[i];
    const int cc = c < 24 ? c : c - 24;
    const int p = cc < 12 ? P.pos_r3[i0 + cc / 3] : P.pos_so3[i0 + (cc - 12) / 3];
    return p < 0 ? -1 : p + cc % 3;
  }
  const int s = c - 48;
  if (TYPE == RT_SURFEL) { const int b = s < 3 ? TB_LQ : TB_LP; return P.pos_sens[b] < 0 ? -1 : P.pos_sens[b] + s % 3; }
  if (TYPE == RT_CAM) {
    if (s < 6) { const int b = s < 3 ? TB_CQ : TB_CP; return P.pos_sens[b] < 0 ? -1 : P.pos_sens[b] + s % 3; }
    return P.pos_rho[T.ia[i]];
  }
  // camsurf
  const int b = s < 3 ? TB_CQ : s < 6 ? TB_CP : s < 9 ? TB_LQ : TB_LP;
  return P.pos_sens[b] < 0 ? -1 : P.pos_sens[b] + s % 3;
}

// ceres::HuberLoss + Corrector (rho'' <= 0 branch): returns rho(s), sets the residual/Jacobian scale sqrt(rho')
LVI_HD double huber(double s, double a, double& scale) {
  scale = 1.0;
  if (!(a > 0.0)) return s;
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    double rho1 = a / r;
    if (rho1 < 2.2250738585072014e-308) rho1 = 2.2250738585072014e-308;
    scale = sqrt(rho1);
    return 2.0 * a * r - b;
  }
  return s;
}

}  // namespace lvi

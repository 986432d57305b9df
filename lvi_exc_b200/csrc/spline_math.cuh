// spline_math.cuh — closed-form evaluation AND analytic Jacobians of the Kontiki split trajectory.
//
// Replaces the per-call heap-allocating, Jet-typed evaluators
//   UniformR3SplineSegmentView::Evaluate  (K/trajectories/uniform_r3_spline_trajectory.h:36-103)
//   UniformSO3SplineSegmentView::Evaluate (K/trajectories/uniform_so3_spline_trajectory.h:46-125)
//   math::logq / expq / angular_velocity  (K/math/quaternion_math.h:16-95)
// and Ceres' forward-mode autodiff through them (ceil(active/4) functor passes per residual).
// Orientation: q(u) = q0 * prod_j exp(B~_j(u) log(q_{j-1}^-1 q_j)); derivative with respect to the LEFT
// perturbation q_j <- Exp(theta_j) q_j (Ceres' EigenQuaternionParameterization uses delta = theta/2):
//   dphi/dtheta = [I - M1, M1 - M2, M2 - M3, M3],  M_j = P_j * b_j Jr(b_j d_j) Jr^-1(d_j) * R_j^T
// (Sommer et al., "Efficient derivative computation for cumulative B-splines on Lie groups", CVPR 2020, restated for
// left perturbations).  __host__ __device__ so the same code is unit-tested on the CPU against the oracle's Jets.
#pragma once
#include <math.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif
#define LVI_HD __host__ __device__ __forceinline__

namespace lvi {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };  // row-major
struct Q4 { double x, y, z, w; };

LVI_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
LVI_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
LVI_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
LVI_HD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
LVI_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LVI_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LVI_HD M3 m3_identity() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0; return r; }
LVI_HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.0; return r; }
LVI_HD M3 operator*(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C.m[r * 3 + c] = A.m[r * 3] * B.m[c] + A.m[r * 3 + 1] * B.m[3 + c] + A.m[r * 3 + 2] * B.m[6 + c];
  return C;
}
LVI_HD M3 operator-(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i]; return C; }
LVI_HD M3 operator+(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] + B.m[i]; return C; }
LVI_HD M3 operator*(double s, const M3& A) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = s * A.m[i]; return C; }
LVI_HD V3 operator*(const M3& A, V3 v) { return v3(A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z); }
LVI_HD M3 transpose(const M3& A) { M3 C; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C.m[r * 3 + c] = A.m[c * 3 + r]; return C; }
LVI_HD V3 mulT(const M3& A, V3 v) { return v3(A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z, A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z); }
LVI_HD M3 skew(V3 v) { M3 K; K.m[0] = 0; K.m[1] = -v.z; K.m[2] = v.y; K.m[3] = v.z; K.m[4] = 0; K.m[5] = -v.x; K.m[6] = -v.y; K.m[7] = v.x; K.m[8] = 0; return K; }
// row vector times matrix: g^T A
LVI_HD V3 vecmat(V3 g, const M3& A) { return v3(g.x * A.m[0] + g.y * A.m[3] + g.z * A.m[6], g.x * A.m[1] + g.y * A.m[4] + g.z * A.m[7], g.x * A.m[2] + g.y * A.m[5] + g.z * A.m[8]); }

LVI_HD Q4 q4(double x, double y, double z, double w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
LVI_HD Q4 qmul(Q4 a, Q4 b) {  // Hamilton product, Eigen order
  return q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
LVI_HD Q4 qconj(Q4 a) { return q4(-a.x, -a.y, -a.z, a.w); }
LVI_HD M3 qmat(Q4 q) {
  M3 R;
  const double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z, xw = q.x * q.w, yw = q.y * q.w, zw = q.z * q.w;
  R.m[0] = 1 - 2 * (yy + zz); R.m[1] = 2 * (xy - zw); R.m[2] = 2 * (xz + yw);
  R.m[3] = 2 * (xy + zw); R.m[4] = 1 - 2 * (xx + zz); R.m[5] = 2 * (yz - xw);
  R.m[6] = 2 * (xz - yw); R.m[7] = 2 * (yz + xw); R.m[8] = 1 - 2 * (xx + yy);
  return R;
}
// Eigen's quaternion-vector product: v + w*(2 u x v) + u x (2 u x v)
LVI_HD V3 qrot(Q4 q, V3 v) {
  const V3 u = v3(q.x, q.y, q.z);
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}

// half-angle log / exp exactly as K/math/quaternion_math.h (branch at |v|^2 <= 1e-16, Q8)
LVI_HD V3 logq_half(Q4 q) {
  const double v2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double k = 1.0;
  if (v2 > 1e-16) { const double vn = sqrt(v2); k = atan2(vn, q.w) / vn; }
  return v3(q.x * k, q.y * k, q.z * k);
}
LVI_HD Q4 expq_half(V3 h) {
  const double v2 = dot(h, h);
  double ka = 1.0, kv = 1.0;
  if (v2 > 1e-16) { const double vn = sqrt(v2); ka = cos(vn); kv = sin(vn) / vn; }
  return q4(kv * h.x, kv * h.y, kv * h.z, ka);
}

// coefficients (a, b, c) of a I + b K + c K^2 with K = [phi]x
struct KPoly { double a, b, c; };
LVI_HD KPoly jr_poly(double th2, double s /* scale: Jr(s*phi) expressed in K = [phi]x */) {
  KPoly p; p.a = 1.0;
  const double t2 = th2 * s * s;
  if (t2 < 1e-10) { p.b = -0.5 * s * (1.0 - t2 / 12.0); p.c = s * s * (1.0 / 6.0) * (1.0 - t2 / 20.0); return p; }
  const double t = sqrt(t2);
  p.b = -(1.0 - cos(t)) / t2 * s;
  p.c = (t - sin(t)) / (t2 * t) * s * s;
  return p;
}
LVI_HD KPoly jrinv_poly(double th2) {
  KPoly p; p.a = 1.0; p.b = 0.5;
  if (th2 < 1e-10) { p.c = (1.0 / 12.0) * (1.0 + th2 / 60.0); return p; }
  const double t = sqrt(th2);
  p.c = 1.0 / th2 - (1.0 + cos(t)) / (2.0 * t * sin(t));
  return p;
}
LVI_HD KPoly kpoly_mul(KPoly p, KPoly q, double th2) {  // K^3 = -th2 K
  KPoly r;
  r.a = p.a * q.a;
  r.b = p.a * q.b + p.b * q.a - th2 * (p.b * q.c + p.c * q.b);
  r.c = p.a * q.c + p.b * q.b + p.c * q.a - th2 * p.c * q.c;
  return r;
}
LVI_HD M3 kpoly_eval(KPoly p, V3 phi) {
  const M3 K = skew(phi);
  const M3 K2 = K * K;
  M3 R;
  for (int i = 0; i < 9; ++i) R.m[i] = p.b * K.m[i] + p.c * K2.m[i];
  R.m[0] += p.a; R.m[4] += p.a; R.m[8] += p.a;
  return R;
}

// cubic B-spline bases (K/trajectories/spline_base.h:19-29): row vector U*M
LVI_HD void basis_pos(double u, double B[4]) {
  const double u2 = u * u, u3 = u2 * u;
  B[0] = (1.0 - 3.0 * u + 3.0 * u2 - u3) / 6.0;
  B[1] = (4.0 - 6.0 * u2 + 3.0 * u3) / 6.0;
  B[2] = (1.0 + 3.0 * u + 3.0 * u2 - 3.0 * u3) / 6.0;
  B[3] = u3 / 6.0;
}
LVI_HD void basis_acc(double u, double dt_inv, double B[4]) {  // (0,0,2,6u) * M / dt^2
  const double s = dt_inv * dt_inv;
  B[0] = s * (1.0 - u);
  B[1] = s * (3.0 * u - 2.0);
  B[2] = s * (1.0 - 3.0 * u);
  B[3] = s * u;
}
LVI_HD void basis_cumul(double u, double B[4]) {  // U * M_cumul
  const double u2 = u * u, u3 = u2 * u;
  B[0] = 1.0;
  B[1] = (5.0 + 3.0 * u - 3.0 * u2 + u3) / 6.0;
  B[2] = (1.0 + 3.0 * u + 3.0 * u2 - 2.0 * u3) / 6.0;
  B[3] = u3 / 6.0;
}
LVI_HD void basis_cumul_d(double u, double dt_inv, double B[4]) {  // (0,1,2u,3u^2)/dt * M_cumul
  const double u2 = u * u;
  B[0] = 0.0;
  B[1] = dt_inv * (3.0 - 6.0 * u + 3.0 * u2) / 6.0;
  B[2] = dt_inv * (3.0 + 6.0 * u - 6.0 * u2) / 6.0;
  B[3] = dt_inv * (3.0 * u2) / 6.0;
}

struct So3Eval {
  Q4 q;        // orientation
  M3 R;        // rotation matrix of q
  V3 w_body;   // body angular velocity (= q^-1 * w_world of the reference)
  M3 C[4];     // d(left rot-vector perturbation of q) / d(theta_j)
  M3 N[4];     // d(w_body) / d(theta_j)
};

// cps: 4 consecutive control quaternions (x,y,z,w). want_jac / want_w select the extra outputs.
LVI_HD void so3_spline_eval(const double* cps /*16*/, double u, double dt_inv, bool want_w, bool want_jac, So3Eval& e) {
  double B[4], dB[4];
  basis_cumul(u, B);
  basis_cumul_d(u, dt_inv, dB);
  Q4 qc[4];
  for (int j = 0; j < 4; ++j) qc[j] = q4(cps[4 * j], cps[4 * j + 1], cps[4 * j + 2], cps[4 * j + 3]);
  Q4 q = qc[0];
  V3 d[4]; M3 A[4]; Q4 part[4];
  part[0] = q;
  for (int j = 1; j < 4; ++j) {
    const V3 h = logq_half(qmul(qconj(qc[j - 1]), qc[j]));
    const Q4 ej = expq_half(B[j] * h);
    q = qmul(q, ej);
    part[j] = q;
    d[j] = 2.0 * h;       // rotation vector of q_{j-1}^-1 q_j
    A[j] = qmat(ej);      // Exp(b_j d_j)
  }
  e.q = q;
  e.R = qmat(q);
  if (want_w) {
    V3 w = dB[1] * d[1];
    w = mulT(A[2], w) + dB[2] * d[2];
    w = mulT(A[3], w) + dB[3] * d[3];
    e.w_body = w;
  }
  if (!want_jac) return;
  M3 Mj[4], Nj[4];
  V3 w1 = dB[1] * d[1];
  V3 w2 = mulT(A[2], w1) + dB[2] * d[2];
  for (int j = 1; j < 4; ++j) {
    const double th2 = dot(d[j], d[j]);
    const KPoly jri = jrinv_poly(th2);
    const KPoly jrb = jr_poly(th2, B[j]);
    const M3 RjT = transpose(qmat(qc[j]));
    const M3 G = kpoly_eval(kpoly_mul(jrb, jri, th2), d[j]);  // Jr(b d) Jr^-1(d)
    Mj[j] = qmat(part[j]) * (B[j] * (G * RjT));
    if (want_w) {
      const M3 JiR = kpoly_eval(jri, d[j]) * RjT;  // Jr^-1(d_j) R_j^T
      M3 W;
      if (j == 1) {
        W = dB[1] * (transpose(A[3]) * transpose(A[2]));
      } else {
        const V3 v = (j == 2) ? mulT(A[2], w1) : mulT(A[3], w2);  // A_j^T w^(j-1)
        M3 inner = B[j] * (skew(v) * kpoly_eval(jrb, d[j]));
        inner.m[0] += dB[j]; inner.m[4] += dB[j]; inner.m[8] += dB[j];
        W = (j == 2) ? transpose(A[3]) * inner : inner;
      }
      Nj[j] = W * JiR;
    }
  }
  e.C[0] = m3_identity() - Mj[1]; e.C[1] = Mj[1] - Mj[2]; e.C[2] = Mj[2] - Mj[3]; e.C[3] = Mj[3];
  if (want_w) { e.N[0] = -1.0 * Nj[1]; e.N[1] = Nj[1] - Nj[2]; e.N[2] = Nj[2] - Nj[3]; e.N[3] = Nj[3]; }
}

// Orientation only, with the knot-to-knot half-angle logs h_j = logq_half(q_{j-1}^-1 q_j) precomputed once per knot (they do not
// depend on t): q(u) = q_i0 * prod_{j=1..3} expq_half(B~_j(u) h_{i0+j}).  Used by the per-point pose queries of the map path.
LVI_HD Q4 so3_orientation_pre(const double* q_i0 /*4*/, const double* hlog /*3 x 3: h_{i0+1..i0+3}*/, double u) {
  double B[4];
  basis_cumul(u, B);
  Q4 q = q4(q_i0[0], q_i0[1], q_i0[2], q_i0[3]);
#pragma unroll
  for (int j = 1; j < 4; ++j) q = qmul(q, expq_half(B[j] * v3(hlog[3 * (j - 1)], hlog[3 * (j - 1) + 1], hlog[3 * (j - 1) + 2])));
  return q;
}

LVI_HD V3 r3_spline(const double* cps /*12*/, const double B[4]) {
  return v3(B[0] * cps[0] + B[1] * cps[3] + B[2] * cps[6] + B[3] * cps[9], B[0] * cps[1] + B[1] * cps[4] + B[2] * cps[7] + B[3] * cps[10],
            B[0] * cps[2] + B[1] * cps[5] + B[2] * cps[8] + B[3] * cps[11]);
}

}  // namespace lvi

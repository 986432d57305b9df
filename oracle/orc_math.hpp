// oracle/orc_math.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
//
// Minimal numeric kit for the CPU restatement of the LVI-ExC hot path: forward-mode dual numbers that mirror
// ceres::Jet<double,N> (the reference differentiates every residual with DynamicAutoDiffCostFunction, e.g.
// K/measurements/lidar_surfel_point.h:145, default stride 4), 3-vectors and Eigen-ordered quaternions (x,y,z,w).
// Neither Eigen nor Ceres exists in this container — parity is unpinned by upstream vectors; the
// kit is validated by tests/test_oracle_*.py (finite differences, numpy eigh/cov, spline identities).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// ---- Jet ------------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT implicit like ceres::Jet
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; if (k >= 0) v[k] = 1.0; }
};
#define ORC_JET_BIN(op, expr_a, expr_v)                                             \
  template <int N> inline Jet<N> operator op(const Jet<N>& f, const Jet<N>& g) {    \
    Jet<N> h; h.a = expr_a; for (int i = 0; i < N; ++i) h.v[i] = expr_v; return h; }
ORC_JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
ORC_JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
ORC_JET_BIN(*, f.a * g.a, f.a * g.v[i] + f.v[i] * g.a)
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { return f + s; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { return (-f) + s; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { return f * (1.0 / s); }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) { return Jet<N>(s) / g; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) { f = f - g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, const Jet<N>& g) { f = f * g; return f; }
template <int N> inline bool operator>(const Jet<N>& f, double s) { return f.a > s; }
template <int N> inline bool operator<(const Jet<N>& f, double s) { return f.a < s; }
template <int N> inline bool operator>=(const Jet<N>& f, double s) { return f.a >= s; }
template <int N> inline bool operator<=(const Jet<N>& f, double s) { return f.a <= s; }

inline double sqrt_(double x) { return std::sqrt(x); }
inline double sin_(double x) { return std::sin(x); }
inline double cos_(double x) { return std::cos(x); }
inline double exp_(double x) { return std::exp(x); }
inline double abs_(double x) { return std::fabs(x); }
inline double atan2_(double y, double x) { return std::atan2(y, x); }
inline double val(double x) { return x; }
template <int N> inline double val(const Jet<N>& x) { return x.a; }
template <int N> inline Jet<N> scale_d(const Jet<N>& f, double a, double d) { Jet<N> h; h.a = a; for (int i = 0; i < N; ++i) h.v[i] = d * f.v[i]; return h; }
template <int N> inline Jet<N> sqrt_(const Jet<N>& f) { const double s = std::sqrt(f.a); return scale_d(f, s, 1.0 / (2.0 * s)); }
template <int N> inline Jet<N> sin_(const Jet<N>& f) { return scale_d(f, std::sin(f.a), std::cos(f.a)); }
template <int N> inline Jet<N> cos_(const Jet<N>& f) { return scale_d(f, std::cos(f.a), -std::sin(f.a)); }
template <int N> inline Jet<N> exp_(const Jet<N>& f) { const double e = std::exp(f.a); return scale_d(f, e, e); }
template <int N> inline Jet<N> abs_(const Jet<N>& f) { return f.a < 0 ? -f : f; }
template <int N> inline Jet<N> atan2_(const Jet<N>& g, const Jet<N>& f) {  // atan2(g, f)
  Jet<N> h; const double t = 1.0 / (f.a * f.a + g.a * g.a); h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]); return h; }

// ---- small fixed-size algebra ------------------------------------------------------------------------------
template <class T> struct V3 {
  T x, y, z;
  V3() : x(T(0.0)), y(T(0.0)), z(T(0.0)) {}
  V3(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator-(const V3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline V3<T> operator*(const V3<T>& a, const T& s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline V3<T> operator*(const T& s, const V3<T>& a) { return a * s; }
template <class T> inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <class T> inline T norm(const V3<T>& a) { return sqrt_(dot(a, a)); }
template <class T, class S> inline V3<T> cast(const V3<S>& a) { return {T(a.x), T(a.y), T(a.z)}; }

// Quaternion with Eigen's storage order (x,y,z,w) and Hamilton product (Eigen::Quaternion operator*).
template <class T> struct Quat {
  T x, y, z, w;
  Quat() : x(T(0.0)), y(T(0.0)), z(T(0.0)), w(T(1.0)) {}
  Quat(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
  V3<T> vec() const { return {x, y, z}; }
  Quat conj() const { return {-x, -y, -z, w}; }
};
template <class T> inline Quat<T> operator*(const Quat<T>& a, const Quat<T>& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z}; }
// Eigen's Quaternion * Vector3 (QuaternionBase::_transformVector): v + w*(2 u×v) + u×(2 u×v)
template <class T> inline V3<T> rot(const Quat<T>& q, const V3<T>& v) {
  V3<T> u = q.vec();
  V3<T> uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv); }
template <class T> inline T qnorm(const Quat<T>& q) { return sqrt_(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); }

}  // namespace orc

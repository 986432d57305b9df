// Small dense helpers of the initial-guess stage where the reference calls Eigen's JacobiSVD and LDLT (hosts without Eigen).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace lvi_init {

struct MatX {   // row-major
  int rows = 0, cols = 0;
  std::vector<double> a;
  MatX() = default;
  MatX(int r, int c) : rows(r), cols(c), a(static_cast<size_t>(r) * c, 0.0) {}
  double& operator()(int r, int c) { return a[static_cast<size_t>(r) * cols + c]; }
  double operator()(int r, int c) const { return a[static_cast<size_t>(r) * cols + c]; }
};

inline MatX gram(const MatX& A, double scale) {   // A^T A * scale
  MatX G(A.cols, A.cols);
  for (int i = 0; i < A.cols; ++i)
    for (int j = i; j < A.cols; ++j) {
      double s = 0;
      for (int r = 0; r < A.rows; ++r) s += A(r, i) * A(r, j);
      G(i, j) = G(j, i) = s * scale;
    }
  return G;
}
inline std::vector<double> atb(const MatX& A, const std::vector<double>& b, double scale) {
  std::vector<double> y(A.cols, 0.0);
  for (int i = 0; i < A.cols; ++i) { double s = 0; for (int r = 0; r < A.rows; ++r) s += A(r, i) * b[r]; y[i] = s * scale; }
  return y;
}

// eigen-decomposition of a symmetric matrix by cyclic Jacobi: eigenvalues (unsorted) and eigenvectors in the columns of V
inline void jacobi_eig(MatX S, std::vector<double>& ev, MatX& V) {
  const int n = S.rows;
  V = MatX(n, n);
  for (int i = 0; i < n; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += S(p, q) * S(p, q);
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        if (S(p, q) == 0.0) continue;
        const double theta = (S(q, q) - S(p, p)) / (2.0 * S(p, q));
        double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (theta < 0) t = -t;
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) { const double kp = S(k, p), kq = S(k, q); S(k, p) = c * kp - s * kq; S(k, q) = s * kp + c * kq; }
        for (int k = 0; k < n; ++k) { const double pk = S(p, k), qk = S(q, k); S(p, k) = c * pk - s * qk; S(q, k) = s * pk + c * qk; }
        for (int k = 0; k < n; ++k) { const double kp = V(k, p), kq = V(k, q); V(k, p) = c * kp - s * kq; V(k, q) = s * kp + c * kq; }
      }
  }
  ev.resize(n);
  for (int i = 0; i < n; ++i) ev[i] = S(i, i);
}

// Eigen's A.ldlt().solve(b) on a (semi-)definite normal matrix: unknowns that no row observes (exactly zero diagonal, zero couplings)
// come back 0; the observed block is solved through its eigen-decomposition (pseudo-inverse below 1e-14 of the largest eigenvalue)
inline std::vector<double> solve_semidefinite(const MatX& A, const std::vector<double>& b) {
  const int n = A.rows;
  std::vector<int> live;
  for (int i = 0; i < n; ++i) if (A(i, i) != 0.0) live.push_back(i);
  std::vector<double> x(n, 0.0);
  const int m = static_cast<int>(live.size());
  if (!m) return x;
  MatX S(m, m);
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) S(i, j) = A(live[i], live[j]);
  std::vector<double> ev;
  MatX V;
  jacobi_eig(S, ev, V);
  double emax = 0;
  for (double e : ev) emax = std::max(emax, std::fabs(e));
  for (int k = 0; k < m; ++k) {
    if (std::fabs(ev[k]) <= 1e-14 * emax) continue;
    double proj = 0;
    for (int i = 0; i < m; ++i) proj += V(i, k) * b[live[i]];
    proj /= ev[k];
    for (int i = 0; i < m; ++i) x[live[i]] += V(i, k) * proj;
  }
  return x;
}

}  // namespace lvi_init

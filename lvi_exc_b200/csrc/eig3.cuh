// eig3.cuh — 3x3 symmetric eigen-solver and 3x3 inverse used by the voxel and surfel kernels.
//
// Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> (N/voxel_grid_covariance_omp_impl.hpp:279,337) and
// pcl::eigen33 (PCL plane refinement).  Cyclic Jacobi, lower triangle as input, eigenvalues ascending, eigenvectors in
// the columns of a row-major 3x3.  Only + - * / sqrt in a fixed order: with -fmad=false the result is bit-identical to
// the CPU oracle's, which the bit-exact plane/association parity relies on.
#pragma once
#include <cuda_runtime.h>

namespace lvi {

__host__ __device__ inline void jacobi3_lower(const double A[9], double evals[3], double evecs[9]) {
  double a00 = A[0], a11 = A[4], a22 = A[8], a01 = A[3], a02 = A[6], a12 = A[7];
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 12; ++sweep) {
    if (a01 == 0.0 && a02 == 0.0 && a12 == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      double app, aqq, apq;
      if (pq == 0) { app = a00; aqq = a11; apq = a01; }
      else if (pq == 1) { app = a00; aqq = a22; apq = a02; }
      else { app = a11; aqq = a22; apq = a12; }
      if (apq == 0.0) continue;
      if (fabs(apq) < 1e-300) {
        if (pq == 0) a01 = 0; else if (pq == 1) a02 = 0; else a12 = 0;
        continue;
      }
      const double theta = (aqq - app) / (2.0 * apq);
      const double at = fabs(theta);
      double t = 1.0 / (at + sqrt(theta * theta + 1.0));
      if (theta < 0.0) t = -t;
      const double c = 1.0 / sqrt(t * t + 1.0);
      const double s = t * c;
      const double napp = app - t * apq;
      const double naqq = aqq + t * apq;
      if (pq == 0) {
        const double ar_p = a02, ar_q = a12;
        a00 = napp; a11 = naqq; a01 = 0.0;
        a02 = c * ar_p - s * ar_q;
        a12 = s * ar_p + c * ar_q;
      } else if (pq == 1) {
        const double ar_p = a01, ar_q = a12;
        a00 = napp; a22 = naqq; a02 = 0.0;
        a01 = c * ar_p - s * ar_q;
        a12 = s * ar_p + c * ar_q;
      } else {
        const double ar_p = a01, ar_q = a02;
        a11 = napp; a22 = naqq; a12 = 0.0;
        a01 = c * ar_p - s * ar_q;
        a02 = s * ar_p + c * ar_q;
      }
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double vp = V[r * 3 + p], vq = V[r * 3 + q];
        V[r * 3 + p] = c * vp - s * vq;
        V[r * 3 + q] = s * vp + c * vq;
      }
    }
  }
  const double d[3] = {a00, a11, a22};
  int idx[3] = {0, 1, 2};
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && d[idx[j]] < d[idx[j - 1]]; --j) { const int t = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = t; }
  for (int k = 0; k < 3; ++k) {
    evals[k] = d[idx[k]];
    for (int r = 0; r < 3; ++r) evecs[r * 3 + k] = V[r * 3 + idx[k]];
  }
}

__host__ __device__ inline void inv3_cofactor(const double M[9], double out[9]) {
  const double c00 = M[4] * M[8] - M[5] * M[7];
  const double c01 = M[5] * M[6] - M[3] * M[8];
  const double c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  const double id = 1.0 / det;
  out[0] = c00 * id;
  out[1] = (M[2] * M[7] - M[1] * M[8]) * id;
  out[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  out[3] = c01 * id;
  out[4] = (M[0] * M[8] - M[2] * M[6]) * id;
  out[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  out[6] = c02 * id;
  out[7] = (M[1] * M[6] - M[0] * M[7]) * id;
  out[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

}  // namespace lvi

// b200reg — small dense solvers that run inside the on-device optimiser step.
#pragma once
#include "common.cuh"

namespace b200 {

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi, double).  evals ascending, eigenvectors in
// the columns of V (row-major 3x3) — the contract of Eigen::SelfAdjointEigenSolver that
// pclomp::VoxelGridCovariance relies on.
__device__ inline void sym_eigen3(const double A_in[9], double evals[3], double V[9]) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { a[i] = A_in[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double apq = a[3 * p + q];
      if (apq == 0.0) continue;
      double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double akp = a[3 * k + p], akq = a[3 * k + q];
        a[3 * k + p] = c * akp - s * akq;
        a[3 * k + q] = s * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double apk = a[3 * p + k], aqk = a[3 * q + k];
        a[3 * p + k] = c * apk - s * aqk;
        a[3 * q + k] = s * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double vkp = V[3 * k + p], vkq = V[3 * k + q];
        V[3 * k + p] = c * vkp - s * vkq;
        V[3 * k + q] = s * vkp + c * vkq;
      }
    }
  }
  // sort ascending (3 elements), permuting the columns of V
  double d[3] = {a[0], a[4], a[8]};
  int idx[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2 - i; ++j)
      if (d[idx[j]] > d[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  double Vs[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    evals[j] = d[idx[j]];
#pragma unroll
    for (int k = 0; k < 3; ++k) Vs[3 * k + j] = V[3 * k + idx[j]];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

// Matrix3d::inverse() by cofactors
__device__ inline void inverse3(const double a[9], double r[9]) {
  double c00 = a[4] * a[8] - a[5] * a[7];
  double c10 = a[5] * a[6] - a[3] * a[8];
  double c20 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c10 + a[2] * c20;
  double inv = 1.0 / det;
  r[0] = c00 * inv;
  r[3] = c10 * inv;
  r[6] = c20 * inv;
  r[1] = (a[2] * a[7] - a[1] * a[8]) * inv;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * inv;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * inv;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * inv;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * inv;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * inv;
}

// One-sided Jacobi SVD solve of a 6x6 system with Eigen's JacobiSVD rank threshold — the slow,
// always-safe path (used when the pivoted elimination below meets a near-singular matrix).
static __device__ __noinline__ void svd_solve6(const double* A, const double* b, double* x) {
  double U[36], V[36];
  for (int i = 0; i < 36; ++i) { U[i] = A[i]; V[i] = (i % 7 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 100; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 6; ++k) {
          alpha += U[6 * k + p] * U[6 * k + p];
          beta += U[6 * k + q] * U[6 * k + q];
          gamma += U[6 * k + p] * U[6 * k + q];
        }
        if (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 6; ++k) {
          double up = U[6 * k + p], uq = U[6 * k + q];
          U[6 * k + p] = c * up - s * uq;
          U[6 * k + q] = s * up + c * uq;
          double vp = V[6 * k + p], vq = V[6 * k + q];
          V[6 * k + p] = c * vp - s * vq;
          V[6 * k + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sig[6], smax = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += U[6 * k + j] * U[6 * k + j];
    sig[j] = sqrt(s);
    smax = fmax(smax, sig[j]);
  }
  const double thr = 6.0 * 2.220446049250313e-16 * smax;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    if (!(sig[j] > thr) || sig[j] == 0.0) continue;
    double utb = 0;
    for (int k = 0; k < 6; ++k) utb += U[6 * k + j] * b[k];
    double w = utb / (sig[j] * sig[j]);
    for (int i = 0; i < 6; ++i) x[i] += V[6 * i + j] * w;
  }
}

// Solve A x = b (6x6, double).  Gaussian elimination with partial pivoting on the common,
// well-conditioned case; when a pivot collapses (rank-deficient or NaN input) the SVD path
// reproduces JacobiSVD::solve's minimum-norm answer (e.g. H = 0 -> x = 0), which is what the
// reference's NDT loop relies on to terminate.
// Every loop is fully unrolled and every index is a compile-time constant (row swaps are
// predicated exchanges), so the 6x7 tableau lives in registers: this runs on one lane between
// two derivative passes and its latency is on the critical path of every Newton iteration.
static __device__ __noinline__ void solve6(const double* A, const double* b, double* x) {
  double M[6][7];
  double rmax[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) M[i][j] = A[6 * i + j];
    M[i][6] = b[i];
    // max |a_ij| as a tree (fmax is exact, so the order does not matter; one 36-long chain did)
    rmax[i] = fmax(fmax(fmax(fabs(M[i][0]), fabs(M[i][1])), fmax(fabs(M[i][2]), fabs(M[i][3]))), fmax(fabs(M[i][4]), fabs(M[i][5])));
  }
  const double amax = fmax(fmax(fmax(rmax[0], rmax[1]), fmax(rmax[2], rmax[3])), fmax(rmax[4], rmax[5]));
  bool ok = amax > 0.0 && amax == amax && amax < 1.7e308;
  double invs[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = fabs(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const double v = fabs(M[i][k]);
      if (v > best) { best = v; piv = i; }
    }
    if (!(best > 1e-11 * amax)) ok = false;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const bool sw = (i == piv);
#pragma unroll
      for (int j = k; j < 7; ++j) {
        const double a = M[k][j], c = M[i][j];
        M[k][j] = sw ? c : a;
        M[i][j] = sw ? a : c;
      }
    }
    const double inv = __drcp_rn(M[k][k]);  // correctly rounded reciprocal: the bits of 1.0 / x, without the general division's slow path
    invs[k] = inv;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const double f = M[i][k] * inv;
#pragma unroll
      for (int j = k + 1; j < 7; ++j) M[i][j] -= f * M[k][j];
    }
  }
  if (!ok) {
    svd_solve6(A, b, x);
    return;
  }
  double xs[6];
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = M[i][6];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) s -= M[i][j] * xs[j];
    xs[i] = s * invs[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = xs[i];
}

// The same elimination by a whole warp: lane i (i < 6) holds row i of [A | b] in registers, the pivot
// search is a three-step butterfly, the row exchange and the pivot-row broadcast are shuffles.  Every
// element goes through the same operations in the same order as in solve6 (f = m_ik * inv,
// m_ij -= f * m_kj; back substitution with ascending j), so the bits are those of the serial routine —
// which kept its 6x7 tableau in a stack frame (the caller's 128 registers are taken) and spent ~4.5 k
// cycles per solve on one lane while the rest of the CTA waited.  Returns false when a pivot collapses
// (the caller then runs svd_solve6 on one lane, as solve6 does).  All 32 lanes must call; x[0..5] is
// valid in every lane on return.
__device__ __forceinline__ bool warp_solve6(const double* A, const double* b, double x[6], int lane) {
  const int row = lane < 6 ? lane : 0;
  double r[7];
#pragma unroll
  for (int j = 0; j < 6; ++j) r[j] = A[6 * row + j];
  r[6] = b[row];
  double amax = fmax(fmax(fmax(fabs(r[0]), fabs(r[1])), fmax(fabs(r[2]), fabs(r[3]))), fmax(fabs(r[4]), fabs(r[5])));
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));  // lanes 6, 7 hold copies of row 0
  amax = __shfl_sync(0xffffffffu, amax, 0);
  bool ok = amax > 0.0 && amax == amax && amax < 1.7e308;
  double invs[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    // first row at or below k with the largest |m_ik| (solve6: strict > in ascending i)
    double best = (lane >= k && lane < 6) ? fabs(r[k]) : -1.0;
    int piv = lane;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int op = __shfl_xor_sync(0xffffffffu, piv, o);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    best = __shfl_sync(0xffffffffu, best, 0);
    piv = __shfl_sync(0xffffffffu, piv, 0);
    // NaN entries: fabs(NaN) never wins a comparison, as in solve6 (v > best is false) — best stays the
    // value of the first candidate, and the test below rejects NaN pivots the same way
    if (!(best > 1e-11 * amax)) ok = false;
    double pr[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < k) continue;
      const double from_piv = __shfl_sync(0xffffffffu, r[j], piv);
      const double from_k = __shfl_sync(0xffffffffu, r[j], k);
      pr[j] = from_piv;                    // the pivot row after the exchange
      if (lane == k) r[j] = from_piv;
      else if (lane == piv) r[j] = from_k;
    }
    const double inv = __drcp_rn(pr[k]);
    invs[k] = inv;
    if (lane > k && lane < 6) {
      const double f = r[k] * inv;
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j > k) r[j] -= f * pr[j];
    }
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double sacc = r[6];
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j > i) sacc -= r[j] * x[j];
    x[i] = __shfl_sync(0xffffffffu, sacc * invs[i], i);
  }
  return ok;
}

}  // namespace b200

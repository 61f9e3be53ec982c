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

// H x = b for a SYMMETRIC 6x6 H by 3x3 blocks, every step closed form (cofactor inverses, Schur complement):
//   H = [A B; B^T D],  S = D - B^T A^-1 B,  x2 = S^-1 (b2 - B^T A^-1 b1),  x1 = A^-1 (b1 - B x2)
// ~250 independent-rich double operations and two reciprocals, no pivot search, no exchange of rows between lanes: run
// redundantly by every lane of the warp that holds the optimiser state it costs a few hundred cycles where the pivoted
// elimination spends ~3 k on dependent shuffles.  There is no pivoting, so the routine VERIFIES its answer: it returns
// false unless both block determinants are well away from zero relative to their blocks AND the residual H x - b is at
// rounding level; the caller then falls back to the pivoted elimination (and that one to the SVD path), so a
// degenerate or indefinite Hessian is never solved badly — only more slowly.  For NDT the blocks are the translation
// (A) and rotation (D) parts of the Hessian, both well conditioned wherever the scan constrains the pose.
// h = upper triangle of H in row order (h00 h01 .. h05 h11 .. h55, 21 values), g = the gradient (the right-hand side is
// -g); both are read from (shared) memory where they are used, block by block, so that no more than ~35 doubles are
// live at a time (the first version took all 27 inputs by value and spilled most of them).
__device__ __forceinline__ bool solve6_schur(const double* h, const double* g, double x[6]) {
  // ---- A^-1 (symmetric): cofactors / det
  double i00, i01, i02, i11, i12, i22, nA;
  bool ok;
  {
    const double a00 = h[0], a01 = h[1], a02 = h[2], a11 = h[6], a12 = h[7], a22 = h[11];
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const double detA = a00 * c00 + a01 * c01 + a02 * c02;
    nA = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
    ok = fabs(detA) > 1e-9 * nA * nA * nA && detA == detA;
    const double iA = __drcp_rn(detA);
    i00 = c00 * iA; i01 = c01 * iA; i02 = c02 * iA; i11 = c11 * iA; i12 = c12 * iA; i22 = c22 * iA;
  }
  // ---- Y = A^-1 B (3x3), z = A^-1 b1 with b1 = -g[0..2]
  const double b00 = h[3], b01 = h[4], b02 = h[5], b10 = h[8], b11 = h[9], b12 = h[10], b20 = h[12], b21 = h[13], b22 = h[14];
  const double y00 = i00 * b00 + i01 * b10 + i02 * b20, y01 = i00 * b01 + i01 * b11 + i02 * b21, y02 = i00 * b02 + i01 * b12 + i02 * b22;
  const double y10 = i01 * b00 + i11 * b10 + i12 * b20, y11 = i01 * b01 + i11 * b11 + i12 * b21, y12 = i01 * b02 + i11 * b12 + i12 * b22;
  const double y20 = i02 * b00 + i12 * b10 + i22 * b20, y21 = i02 * b01 + i12 * b11 + i22 * b21, y22 = i02 * b02 + i12 * b12 + i22 * b22;
  double z0, z1, z2;
  {
    const double r0 = -g[0], r1 = -g[1], r2 = -g[2];
    z0 = i00 * r0 + i01 * r1 + i02 * r2; z1 = i01 * r0 + i11 * r1 + i12 * r2; z2 = i02 * r0 + i12 * r1 + i22 * r2;
  }
  // ---- S = D - B^T Y (symmetric), r = b2 - B^T z; x2 = S^-1 r
  {
    const double s00 = h[15] - (b00 * y00 + b10 * y10 + b20 * y20), s01 = h[16] - (b00 * y01 + b10 * y11 + b20 * y21), s02 = h[17] - (b00 * y02 + b10 * y12 + b20 * y22);
    const double s11 = h[18] - (b01 * y01 + b11 * y11 + b21 * y21), s12 = h[19] - (b01 * y02 + b11 * y12 + b21 * y22), s22 = h[20] - (b02 * y02 + b12 * y12 + b22 * y22);
    const double r0 = -g[3] - (b00 * z0 + b10 * z1 + b20 * z2), r1 = -g[4] - (b01 * z0 + b11 * z1 + b21 * z2), r2 = -g[5] - (b02 * z0 + b12 * z1 + b22 * z2);
    const double e00 = s11 * s22 - s12 * s12, e01 = s02 * s12 - s01 * s22, e02 = s01 * s12 - s02 * s11;
    const double e11 = s00 * s22 - s02 * s02, e12 = s01 * s02 - s00 * s12, e22 = s00 * s11 - s01 * s01;
    const double detS = s00 * e00 + s01 * e01 + s02 * e02;
    const double nS = fmax(fmax(fabs(s00), fabs(s11)), fmax(fabs(s22), fmax(fabs(s01), fmax(fabs(s02), fabs(s12)))));
    ok = ok && fabs(detS) > 1e-9 * nS * nS * nS && detS == detS;
    const double iS = __drcp_rn(detS);
    x[3] = (e00 * r0 + e01 * r1 + e02 * r2) * iS;
    x[4] = (e01 * r0 + e11 * r1 + e12 * r2) * iS;
    x[5] = (e02 * r0 + e12 * r1 + e22 * r2) * iS;
  }
  x[0] = z0 - (y00 * x[3] + y01 * x[4] + y02 * x[5]);
  x[1] = z1 - (y10 * x[3] + y11 * x[4] + y12 * x[5]);
  x[2] = z2 - (y20 * x[3] + y21 * x[4] + y22 * x[5]);
  // ---- residual H x + g at rounding level relative to the size of its terms (rows re-read where they are used)
  double res, hmax = nA;
  {
    const double q0 = h[0] * x[0] + h[1] * x[1] + h[2] * x[2] + b00 * x[3] + b01 * x[4] + b02 * x[5] + g[0];
    const double q1 = h[1] * x[0] + h[6] * x[1] + h[7] * x[2] + b10 * x[3] + b11 * x[4] + b12 * x[5] + g[1];
    const double q2 = h[2] * x[0] + h[7] * x[1] + h[11] * x[2] + b20 * x[3] + b21 * x[4] + b22 * x[5] + g[2];
    res = fmax(fabs(q0), fmax(fabs(q1), fabs(q2)));
    hmax = fmax(hmax, fmax(fmax(fabs(b00), fabs(b11)), fabs(b22)));
  }
  {
    const double d00 = h[15], d01 = h[16], d02 = h[17], d11 = h[18], d12 = h[19], d22 = h[20];
    const double q3 = b00 * x[0] + b10 * x[1] + b20 * x[2] + d00 * x[3] + d01 * x[4] + d02 * x[5] + g[3];
    const double q4 = b01 * x[0] + b11 * x[1] + b21 * x[2] + d01 * x[3] + d11 * x[4] + d12 * x[5] + g[4];
    const double q5 = b02 * x[0] + b12 * x[1] + b22 * x[2] + d02 * x[3] + d12 * x[4] + d22 * x[5] + g[5];
    res = fmax(res, fmax(fabs(q3), fmax(fabs(q4), fabs(q5))));
    hmax = fmax(hmax, fmax(fabs(d00), fmax(fabs(d11), fabs(d22))));
  }
  const double xmax = fmax(fmax(fabs(x[0]), fabs(x[1])), fmax(fmax(fabs(x[2]), fabs(x[3])), fmax(fabs(x[4]), fabs(x[5]))));
  const double bmax = fmax(fmax(fabs(g[0]), fabs(g[1])), fmax(fmax(fabs(g[2]), fabs(g[3])), fmax(fabs(g[4]), fabs(g[5]))));
  return ok && res <= 1e-9 * (hmax * xmax + bmax) && res == res;
}

// warp_solve6 for a system whose rows already sit in registers: lane i < 6 passes row i of [A | b] in r[0..6], the other
// lanes pass zeros.  Same operations in the same order as warp_solve6 / solve6 (so the same bits), without the shared-
// memory round trip of the operands.  x[0..5] is valid in every lane on return; false when a pivot collapses.
__device__ __forceinline__ bool warp_solve6_rows(double r[7], double x[6], int lane) {
  double amax = fmax(fmax(fmax(fabs(r[0]), fabs(r[1])), fmax(fabs(r[2]), fabs(r[3]))), fmax(fabs(r[4]), fabs(r[5])));
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));  // lanes 6, 7 hold zeros: neutral
  amax = __shfl_sync(0xffffffffu, amax, 0);
  bool ok = amax > 0.0 && amax == amax && amax < 1.7e308;
  double invs[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
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
    if (!(best > 1e-11 * amax)) ok = false;
    double pr[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j < k) continue;
      const double from_piv = __shfl_sync(0xffffffffu, r[j], piv);
      const double from_k = __shfl_sync(0xffffffffu, r[j], k);
      pr[j] = from_piv;
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

// ORACLE — test infrastructure only.  Nothing under oracle/ is linked, imported or
// executed by the product path (delta_graph_slam_b200/); it is the CPU checker the
// CUDA engine is compared against (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs).
//
// Small fixed-size dense linear algebra standing in for the Eigen routines the
// upstream libraries call: SelfAdjointEigenSolver<Matrix3d>,
// Matrix3d::inverse, JacobiSVD<Matrix6d>::solve, LDLT<Matrix6d>::solve,
// Matrix3f::eulerAngles(0,1,2).  PARITY UNPINNED: Eigen is not available in this
// image, so these are restatements of the published algorithms, validated by
// known-answer tests against numpy (tests/test_oracle_linalg.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

// ---- 3x3 double, row-major -------------------------------------------------
struct M3 {
  double m[9];
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline M3 m3_zero() { M3 a; std::memset(a.m, 0, sizeof a.m); return a; }
inline M3 m3_identity() { M3 a = m3_zero(); a(0, 0) = a(1, 1) = a(2, 2) = 1.0; return a; }
inline M3 m3_mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
  return c;
}
inline M3 m3_transpose(const M3& a) {
  M3 t;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t(i, j) = a(j, i);
  return t;
}

// Matrix3d::inverse(): cofactor expansion (Eigen's compute_inverse_size3_helper).
inline M3 m3_inverse(const M3& a) {
  M3 c;
  c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  c(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  c(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  double det = a(0, 0) * c(0, 0) + a(0, 1) * c(1, 0) + a(0, 2) * c(2, 0);
  double inv = 1.0 / det;
  M3 r;
  r(0, 0) = c(0, 0) * inv;
  r(1, 0) = c(1, 0) * inv;
  r(2, 0) = c(2, 0) * inv;
  r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * inv;
  r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * inv;
  r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * inv;
  r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * inv;
  r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * inv;
  r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * inv;
  return r;
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi).  Eigenvalues ascending,
// eigenvectors in the COLUMNS of evecs — the SelfAdjointEigenSolver contract.
inline void m3_sym_eigen(const M3& a_in, double evals[3], M3& evecs) {
  M3 a = a_in;
  M3 v = m3_identity();
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a(0, 1) * a(0, 1) + a(0, 2) * a(0, 2) + a(1, 2) * a(1, 2);
    double diag = a(0, 0) * a(0, 0) + a(1, 1) * a(1, 1) + a(2, 2) * a(2, 2);
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = a(p, q);
        if (apq == 0.0) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * J
          double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T * A
          double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v(k, p), vkq = v(k, q);
          v(k, p) = c * vkp - s * vkq;
          v(k, q) = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {a(0, 0), a(1, 1), a(2, 2)};
  std::sort(idx, idx + 3, [&](int x, int y) { return d[x] < d[y]; });
  for (int j = 0; j < 3; ++j) {
    evals[j] = d[idx[j]];
    for (int k = 0; k < 3; ++k) evecs(k, j) = v(k, idx[j]);
  }
}

// ---- 6x6 double, row-major -------------------------------------------------
struct M6 {
  double m[36];
  double& operator()(int r, int c) { return m[6 * r + c]; }
  double operator()(int r, int c) const { return m[6 * r + c]; }
};
struct V6 {
  double v[6];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
inline double v6_dot(const V6& a, const V6& b) {
  double s = 0;
  for (int i = 0; i < 6; ++i) s += a[i] * b[i];
  return s;
}
inline double v6_norm(const V6& a) { return std::sqrt(v6_dot(a, a)); }

// JacobiSVD<Matrix6d>(A, FullU|FullV).solve(b): one-sided (Hestenes) Jacobi SVD,
// pseudo-inverse with Eigen's default rank threshold (diagSize * epsilon * sigma_max).
inline V6 m6_svd_solve(const M6& A, const V6& b) {
  double U[36], V[36];
  for (int i = 0; i < 36; ++i) { U[i] = A.m[i]; V[i] = 0.0; }
  for (int i = 0; i < 6; ++i) V[6 * i + i] = 1.0;
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
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
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
    sig[j] = std::sqrt(s);
    smax = std::max(smax, sig[j]);
  }
  const double thr = 6.0 * 2.220446049250313e-16 * smax;
  V6 x;
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    if (!(sig[j] > thr) || sig[j] == 0.0) continue;
    double utb = 0;
    for (int k = 0; k < 6; ++k) utb += U[6 * k + j] * b[k];
    double w = utb / (sig[j] * sig[j]);  // (u_j/sigma_j . b) / sigma_j
    for (int i = 0; i < 6; ++i) x[i] += V[6 * i + j] * w;
  }
  return x;
}

// LDLT<Matrix6d>(A).solve(b) for a symmetric matrix (diagonal pivoting).
inline V6 m6_ldlt_solve(const M6& A_in, const V6& b_in) {
  double A[36];
  std::memcpy(A, A_in.m, sizeof A);
  int perm[6];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  double L[36] = {0}, D[6];
  for (int k = 0; k < 6; ++k) {
    int piv = k;
    double best = std::fabs(A[6 * k + k]);
    for (int i = k + 1; i < 6; ++i)
      if (std::fabs(A[6 * i + i]) > best) { best = std::fabs(A[6 * i + i]); piv = i; }
    if (piv != k) {  // symmetric row/column swap
      for (int j = 0; j < 6; ++j) std::swap(A[6 * k + j], A[6 * piv + j]);
      for (int j = 0; j < 6; ++j) std::swap(A[6 * j + k], A[6 * j + piv]);
      for (int j = 0; j < k; ++j) std::swap(L[6 * k + j], L[6 * piv + j]);
      std::swap(perm[k], perm[piv]);
    }
    D[k] = A[6 * k + k];
    L[6 * k + k] = 1.0;
    for (int i = k + 1; i < 6; ++i) L[6 * i + k] = (D[k] != 0.0) ? A[6 * i + k] / D[k] : 0.0;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j < 6; ++j) A[6 * i + j] -= L[6 * i + k] * D[k] * L[6 * j + k];
  }
  double y[6], z[6];
  for (int i = 0; i < 6; ++i) y[i] = b_in[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) y[i] -= L[6 * i + j] * y[j];
  for (int i = 0; i < 6; ++i) z[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
  for (int i = 5; i >= 0; --i)
    for (int j = i + 1; j < 6; ++j) z[i] -= L[6 * j + i] * z[j];
  V6 x;
  for (int i = 0; i < 6; ++i) x[perm[i]] = z[i];
  return x;
}

// ---- 4x4 float, row-major (Eigen::Matrix4f values; the C API converts to/from
// Eigen's column-major storage at the boundary) -------------------------------
struct M4f {
  float m[16];
  float& operator()(int r, int c) { return m[4 * r + c]; }
  float operator()(int r, int c) const { return m[4 * r + c]; }
};
inline M4f m4f_identity() {
  M4f a;
  for (int i = 0; i < 16; ++i) a.m[i] = 0.f;
  a(0, 0) = a(1, 1) = a(2, 2) = a(3, 3) = 1.f;
  return a;
}
inline bool m4f_is_identity(const M4f& a) {
  M4f i = m4f_identity();
  for (int k = 0; k < 16; ++k)
    if (a.m[k] != i.m[k]) return false;
  return true;
}
// affine product (last row 0 0 0 1), float, left-to-right accumulation
inline M4f m4f_mul_affine(const M4f& a, const M4f& b) {
  M4f c = m4f_identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    c(i, 3) = ((a(i, 0) * b(0, 3) + a(i, 1) * b(1, 3)) + a(i, 2) * b(2, 3)) + a(i, 3);
  }
  return c;
}

// Translation(p0..2) * Rx(p3) * Ry(p4) * Rz(p5), built in float from the double
// parameter vector (ndt_omp builds Eigen::Translation<float>/AngleAxis<float>
// from static_cast<float>(p(i))).
inline M4f m4f_from_xyz_euler(const double p[6]) {
  // float trig evaluated as the float rounding of the double function: correctly rounded in
  // all but ~1e-9 of cases, where glibc's cosf/sinf (what Eigen::AngleAxisf calls) may differ
  // by one ulp; chosen so the CUDA engine can reproduce the same bits.
  float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
  float cx = (float)std::cos((double)rx), sx = (float)std::sin((double)rx), cy = (float)std::cos((double)ry), sy = (float)std::sin((double)ry);
  float cz = (float)std::cos((double)rz), sz = (float)std::sin((double)rz);
  M4f T = m4f_identity(), Rx = m4f_identity(), Ry = m4f_identity(), Rz = m4f_identity();
  T(0, 3) = (float)p[0]; T(1, 3) = (float)p[1]; T(2, 3) = (float)p[2];
  Rx(1, 1) = cx; Rx(1, 2) = -sx; Rx(2, 1) = sx; Rx(2, 2) = cx;
  Ry(0, 0) = cy; Ry(0, 2) = sy; Ry(2, 0) = -sy; Ry(2, 2) = cy;
  Rz(0, 0) = cz; Rz(0, 1) = -sz; Rz(1, 0) = sz; Rz(1, 1) = cz;
  return m4f_mul_affine(m4f_mul_affine(m4f_mul_affine(T, Rx), Ry), Rz);
}

// Eigen 3.3 Matrix3f::eulerAngles(0,1,2) on the rotation block.
inline void m4f_euler_xyz(const M4f& T, float out[3]) {
  const float pi = 3.14159265358979323846f;
  float r0 = std::atan2(T(1, 2), T(2, 2));
  float c2 = std::sqrt(T(0, 0) * T(0, 0) + T(0, 1) * T(0, 1));
  float r1;
  if (r0 > 0.f) {
    if (r0 > 0.f) r0 -= pi; else r0 += pi;
    r1 = std::atan2(-T(0, 2), -c2);
  } else {
    r1 = std::atan2(-T(0, 2), c2);
  }
  float s1 = std::sin(r0), c1 = std::cos(r0);
  float r2 = std::atan2(s1 * T(2, 0) - c1 * T(1, 0), c1 * T(1, 1) - s1 * T(2, 1));
  out[0] = -r0; out[1] = -r1; out[2] = -r2;
}

// pcl::transformPoint with a Matrix4f: out = M(0:3,0:3) * p + M(0:3,3), float,
// accumulated left to right.
inline void m4f_apply(const M4f& T, const float p[3], float out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = ((T(i, 0) * p[0] + T(i, 1) * p[1]) + T(i, 2) * p[2]) + T(i, 3);
}

}  // namespace orc

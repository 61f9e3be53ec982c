// ORACLE — test infrastructure only (see oracle.hpp).
// fast_gicp::FastGICP on fast_gicp::LsqRegistration, as configured by
// [REF src/hdl_graph_slam/registrations.cpp:27-36].
#include <omp.h>

#include <cmath>
#include <cstdio>

#include "oracle.hpp"

namespace orc {

void FastGICP::setInputSource(const Pt* pts, size_t n) {
  Registration::setInputSource(pts, n);
  source_kdtree_.build(input_.empty() ? nullptr : &input_[0].x, input_.size());
  source_covs_.clear();
}
void FastGICP::setInputTarget(const Pt* pts, size_t n) {
  Registration::setInputTarget(pts, n);
  target_kdtree_.build(target_.empty() ? nullptr : &target_[0].x, target_.size());
  target_covs_.clear();
}
const std::vector<M3>& FastGICP::sourceCovariances() {
  if (source_covs_.size() != input_.size()) calculate_covariances(input_, source_kdtree_, source_covs_);
  return source_covs_;
}
const std::vector<M3>& FastGICP::targetCovariances() {
  if (target_covs_.size() != target_.size()) calculate_covariances(target_, target_kdtree_, target_covs_);
  return target_covs_;
}

void FastGICP::calculate_covariances(const Cloud& cloud, const KdTree& tree, std::vector<M3>& covs) {
  const int k = k_correspondences_;
  covs.assign(cloud.size(), m3_zero());
#pragma omp parallel for num_threads(threads()) schedule(guided, 8)
  for (long i = 0; i < (long)cloud.size(); ++i) {
    std::vector<int> idx(k);
    std::vector<float> d2(k);
    int found = tree.knn(&cloud[i].x, k, idx.data(), d2.data());
    // neighbors (k columns), subtract the row-wise mean, cov = N N^T / k   (k = requested count)
    double mean[3] = {0, 0, 0};
    for (int j = 0; j < found; ++j) {
      const Pt& q = cloud[idx[j]];
      mean[0] += (double)q.x; mean[1] += (double)q.y; mean[2] += (double)q.z;
    }
    for (int a = 0; a < 3; ++a) mean[a] /= (double)k;
    M3 cov = m3_zero();
    for (int j = 0; j < found; ++j) {
      const Pt& q = cloud[idx[j]];
      const double v[3] = {(double)q.x - mean[0], (double)q.y - mean[1], (double)q.z - mean[2]};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) cov(a, b) += v[a] * v[b];
    }
    for (int a = 0; a < 9; ++a) cov.m[a] /= (double)k;
    if (regularization_method_ == REG_NONE) {
      covs[i] = cov;
    } else if (regularization_method_ == REG_FROBENIUS) {
      const double lambda = 1e-3;
      M3 C = cov;
      for (int a = 0; a < 3; ++a) C(a, a) += lambda;
      M3 Ci = m3_inverse(C);
      double fro = 0;
      for (int a = 0; a < 9; ++a) fro += Ci.m[a] * Ci.m[a];
      fro = std::sqrt(fro);
      for (int a = 0; a < 9; ++a) Ci.m[a] /= fro;
      covs[i] = m3_inverse(Ci);
    } else {
      // JacobiSVD of a symmetric PSD 3x3: U = V = eigenvectors, singular values descending
      double ev[3];
      M3 evec;
      m3_sym_eigen(cov, ev, evec);  // ascending
      double values[3];             // in descending-singular-value order
      const int order[3] = {2, 1, 0};
      if (regularization_method_ == REG_PLANE) {
        values[0] = 1; values[1] = 1; values[2] = 1e-3;
      } else if (regularization_method_ == REG_MIN_EIG) {
        for (int a = 0; a < 3; ++a) values[a] = std::max(std::fabs(ev[order[a]]), 1e-3);
      } else {  // REG_NORMALIZED_MIN_EIG
        double smax = std::fabs(ev[2]);
        for (int a = 0; a < 3; ++a) values[a] = std::max(std::fabs(ev[order[a]]) / smax, 1e-3);
      }
      M3 out = m3_zero();
      for (int s = 0; s < 3; ++s) {
        const int col = order[s];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) out(a, b) += values[s] * evec(a, col) * evec(b, col);
      }
      covs[i] = out;
    }
  }
}

FastGICP::Iso FastGICP::iso_mul(const Iso& a, const Iso& b) {
  Iso c;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) c.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
    c.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
  }
  return c;
}

// so3_exp / se3_exp (fast_gicp/so3/so3.hpp)
FastGICP::Iso FastGICP::se3_exp(const V6& a) {
  const double w[3] = {a[0], a[1], a[2]};
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = std::sqrt(theta_sq);
  double imag, real;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  // Eigen::Quaterniond(real, imag*w) is used as-is (unit up to rounding) -> rotation matrix
  Iso T;
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  T.R[0] = 1 - (tyy + tzz); T.R[1] = txy - twz;       T.R[2] = txz + twy;
  T.R[3] = txy + twz;       T.R[4] = 1 - (txx + tzz); T.R[5] = tyz - twx;
  T.R[6] = txz - twy;       T.R[7] = tyz + twx;       T.R[8] = 1 - (txx + tyy);
  double V[9];
  if (theta < 1e-10) {
    for (int k = 0; k < 9; ++k) V[k] = T.R[k];
  } else {
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    const double c1 = (1.0 - std::cos(theta)) / theta_sq, c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int k = 0; k < 9; ++k) V[k] = ((k % 4 == 0) ? 1.0 : 0.0) + c1 * O[k] + c2 * O2[k];
  }
  for (int i = 0; i < 3; ++i) T.t[i] = V[3 * i] * a[3] + V[3 * i + 1] * a[4] + V[3 * i + 2] * a[5];
  return T;
}

void FastGICP::update_correspondences(const Iso& trans) {
  const size_t N = input_.size();
  correspondences_.resize(N);
  sq_distances_.resize(N);
  mahalanobis_.resize(N);
  float Tf[12];  // trans.cast<float>()
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Tf[4 * i + j] = (float)trans.R[3 * i + j];
    Tf[4 * i + 3] = (float)trans.t[i];
  }
  const double thr = corr_dist_threshold_ * corr_dist_threshold_;
#pragma omp parallel for num_threads(threads()) schedule(guided, 8)
  for (long i = 0; i < (long)N; ++i) {
    const Pt& p = input_[i];
    float q[3];
    for (int r = 0; r < 3; ++r) q[r] = ((Tf[4 * r] * p.x + Tf[4 * r + 1] * p.y) + Tf[4 * r + 2] * p.z) + Tf[4 * r + 3];
    int k_idx;
    float k_d2;
    target_kdtree_.knn(q, 1, &k_idx, &k_d2);
    sq_distances_[i] = k_d2;
    correspondences_[i] = ((double)k_d2 < thr) ? k_idx : -1;
    if (correspondences_[i] < 0) continue;
    const M3& cov_A = source_covs_[i];
    const M3& cov_B = target_covs_[k_idx];
    M3 R;
    for (int k = 0; k < 9; ++k) R.m[k] = trans.R[k];
    M3 RCR = m3_mul(m3_mul(R, cov_A), m3_transpose(R));
    for (int k = 0; k < 9; ++k) RCR.m[k] += cov_B.m[k];
    mahalanobis_[i] = m3_inverse(RCR);  // 4x4 with RCR(3,3)=1 inverted, (3,3) zeroed == 3x3 inverse
  }
}

double FastGICP::linearize(const Iso& trans, M6* H, V6* b) {
  update_correspondences(trans);
  n_linearize++;
  const int T = threads();
  std::vector<M6> Hs(T);
  std::vector<V6> bs(T);
  for (int t = 0; t < T; ++t) {
    for (int k = 0; k < 36; ++k) Hs[t].m[k] = 0;
    for (int k = 0; k < 6; ++k) bs[t][k] = 0;
  }
  double sum_errors = 0.0;
#pragma omp parallel for num_threads(T) reduction(+ : sum_errors) schedule(guided, 8)
  for (long i = 0; i < (long)input_.size(); ++i) {
    int target_index = correspondences_[i];
    if (target_index < 0) continue;
    const double a[3] = {(double)input_[i].x, (double)input_[i].y, (double)input_[i].z};
    const double bpt[3] = {(double)target_[target_index].x, (double)target_[target_index].y, (double)target_[target_index].z};
    double ta[3], err[3];
    for (int r = 0; r < 3; ++r) ta[r] = trans.R[3 * r] * a[0] + trans.R[3 * r + 1] * a[1] + trans.R[3 * r + 2] * a[2] + trans.t[r];
    for (int r = 0; r < 3; ++r) err[r] = bpt[r] - ta[r];
    const M3& M = mahalanobis_[i];
    double Me[3];
    for (int r = 0; r < 3; ++r) Me[r] = M(r, 0) * err[0] + M(r, 1) * err[1] + M(r, 2) * err[2];
    sum_errors += err[0] * Me[0] + err[1] * Me[1] + err[2] * Me[2];
    if (!H || !b) continue;
    // J = [ skew(transed) | -I ]   (3x6)
    double J[3][6] = {{0, -ta[2], ta[1], -1, 0, 0}, {ta[2], 0, -ta[0], 0, -1, 0}, {-ta[1], ta[0], 0, 0, 0, -1}};
    double MJ[3][6];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 6; ++c) MJ[r][c] = M(r, 0) * J[0][c] + M(r, 1) * J[1][c] + M(r, 2) * J[2][c];
    int t = omp_get_thread_num();
    for (int r = 0; r < 6; ++r) {
      for (int c = 0; c < 6; ++c) Hs[t](r, c) += J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c] + J[2][r] * MJ[2][c];
      bs[t][r] += J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2];
    }
  }
  if (H && b) {
    for (int k = 0; k < 36; ++k) H->m[k] = 0;
    for (int k = 0; k < 6; ++k) (*b)[k] = 0;
    for (int t = 0; t < T; ++t) {
      for (int k = 0; k < 36; ++k) H->m[k] += Hs[t].m[k];
      for (int k = 0; k < 6; ++k) (*b)[k] += bs[t][k];
    }
  }
  return sum_errors;
}

double FastGICP::compute_error(const Iso& trans) {
  n_error++;
  double sum_errors = 0.0;
#pragma omp parallel for num_threads(threads()) reduction(+ : sum_errors) schedule(guided, 8)
  for (long i = 0; i < (long)input_.size(); ++i) {
    int target_index = correspondences_[i];
    if (target_index < 0) continue;
    const double a[3] = {(double)input_[i].x, (double)input_[i].y, (double)input_[i].z};
    const double bpt[3] = {(double)target_[target_index].x, (double)target_[target_index].y, (double)target_[target_index].z};
    double err[3];
    for (int r = 0; r < 3; ++r) err[r] = bpt[r] - (trans.R[3 * r] * a[0] + trans.R[3 * r + 1] * a[1] + trans.R[3 * r + 2] * a[2] + trans.t[r]);
    const M3& M = mahalanobis_[i];
    double Me[3];
    for (int r = 0; r < 3; ++r) Me[r] = M(r, 0) * err[0] + M(r, 1) * err[1] + M(r, 2) * err[2];
    sum_errors += err[0] * Me[0] + err[1] * Me[1] + err[2] * Me[2];
  }
  return sum_errors;
}

bool FastGICP::is_converged(const Iso& delta) const {
  double rmax = 0, tmax = 0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rmax = std::max(rmax, std::fabs(delta.R[3 * i + j] - (i == j ? 1.0 : 0.0)));
    tmax = std::max(tmax, std::fabs(delta.t[i]));
  }
  return std::max(rmax / rotation_epsilon_, tmax / transformation_epsilon_) < 1;
}

bool FastGICP::step_lm(Iso& x0, Iso& delta) {
  M6 H;
  V6 b;
  double y0 = linearize(x0, &H, &b);
  if (lm_lambda_ < 0.0) {
    double mx = 0;
    for (int i = 0; i < 6; ++i) mx = std::max(mx, std::fabs(H(i, i)));
    lm_lambda_ = lm_init_lambda_factor_ * mx;
  }
  double nu = 2.0;
  for (int i = 0; i < lm_max_iterations_; ++i) {
    M6 A = H;
    for (int k = 0; k < 6; ++k) A(k, k) += lm_lambda_;
    V6 nb;
    for (int k = 0; k < 6; ++k) nb[k] = -b[k];
    V6 d = m6_ldlt_solve(A, nb);
    delta = se3_exp(d);
    Iso xi = iso_mul(delta, x0);
    double yi = compute_error(xi);
    double denom = 0;
    for (int k = 0; k < 6; ++k) denom += d[k] * (lm_lambda_ * d[k] - b[k]);
    double rho = (y0 - yi) / denom;
    if (rho < 0) {
      if (is_converged(delta)) return true;
      lm_lambda_ = nu * lm_lambda_;
      nu = 2 * nu;
      continue;
    }
    x0 = xi;
    lm_lambda_ = lm_lambda_ * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    return true;
  }
  return false;
}

bool FastGICP::step_gn(Iso& x0, Iso& delta) {
  M6 H;
  V6 b;
  linearize(x0, &H, &b);
  V6 nb;
  for (int k = 0; k < 6; ++k) nb[k] = -b[k];
  V6 d = m6_ldlt_solve(H, nb);
  delta = se3_exp(d);
  x0 = iso_mul(delta, x0);
  return true;
}

void FastGICP::computeTransformation(Cloud& output, const M4f& guess) {
  n_linearize = n_error = 0;
  if (source_covs_.size() != input_.size()) calculate_covariances(input_, source_kdtree_, source_covs_);
  if (target_covs_.size() != target_.size()) calculate_covariances(target_, target_kdtree_, target_covs_);
  Iso x0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) x0.R[3 * i + j] = (double)guess(i, j);
    x0.t[i] = (double)guess(i, 3);
  }
  lm_lambda_ = -1.0;
  converged_ = false;
  for (int i = 0; i < max_iterations_ && !converged_; ++i) {
    nr_iterations_ = i;
    Iso delta;
    bool ok = lsq_ == LSQ_LM ? step_lm(x0, delta) : step_gn(x0, delta);
    if (!ok) break;  // "lm not converged!!"
    converged_ = is_converged(delta);
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) final_transformation_(i, j) = (float)x0.R[3 * i + j];
    final_transformation_(i, 3) = (float)x0.t[i];
  }
  for (size_t i = 0; i < input_.size(); ++i) {
    float q[3];
    m4f_apply(final_transformation_, &input_[i].x, q);
    output[i] = Pt{q[0], q[1], q[2], 1.0f};
  }
}

}  // namespace orc

// ORACLE — test infrastructure only (see oracle.hpp).
// pclomp::NormalDistributionsTransform, as configured by
// [REF src/hdl_graph_slam/registrations.cpp:93-120] and driven by
// [REF apps/scan_matching_odometry_nodelet.cpp:180-228] and
// [REF include/hdl_graph_slam/loop_detector.hpp:124-155].
#include <omp.h>

#include <cmath>
#include <cstdio>

#include "oracle.hpp"

namespace orc {

void NDT::setResolution(float r) {
  if (resolution_ != r) {
    resolution_ = r;
    if (!target_.empty()) init();
  }
}
void NDT::setInputTarget(const Pt* pts, size_t n) {
  Registration::setInputTarget(pts, n);
  if (n) init();
}
void NDT::init() {
  target_cells_.setLeafSize(resolution_);
  target_cells_.build(target_.data(), target_.size(), true);
}

void NDT::initGauss() {
  double gauss_c1 = 10.0 * (1 - outlier_ratio_);
  double gauss_c2 = outlier_ratio_ / std::pow((double)resolution_, 3);
  double gauss_d3 = -std::log(gauss_c2);
  gauss_d1_ = -std::log(gauss_c1 + gauss_c2) - gauss_d3;
  gauss_d2_ = -2 * std::log((-std::log(gauss_c1 * std::exp(-0.5) + gauss_c2) - gauss_d3) / gauss_d1_);
}

void NDT::computeAngleDerivatives(const double p[6]) {
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
  const double j[8][3] = {
      {-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy},  // a -> (1,3)
      {cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy},    // b -> (2,3)
      {-sy * cz, sy * sz, cy},                                       // c -> (0,4)
      {sx * cy * cz, -sx * cy * sz, sx * sy},                        // d -> (1,4)
      {-cx * cy * cz, cx * cy * sz, -cx * sy},                       // e -> (2,4)
      {-cy * sz, -cy * cz, 0},                                       // f -> (0,5)
      {cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0},          // g -> (1,5)
      {sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0}};          // h -> (2,5)
  const double h[15][3] = {
      {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy},   // a2
      {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy},  // a3
      {cx * cy * cz, -cx * cy * sz, cx * sy},                        // b2
      {sx * cy * cz, -sx * cy * sz, sx * sy},                        // b3
      {-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0},          // c2
      {cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0},          // c3
      {-cy * cz, cy * sz, sy},                                       // d1 (upstream has +sy here; kept)
      {-sx * sy * cz, sx * sy * sz, sx * cy},                        // d2
      {cx * sy * cz, -cx * sy * sz, -cx * cy},                       // d3
      {sy * sz, sy * cz, 0},                                         // e1
      {-sx * cy * sz, -sx * cy * cz, 0},                             // e2
      {cx * cy * sz, cx * cy * cz, 0},                               // e3
      {-cy * cz, cy * sz, 0},                                        // f1
      {-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0},         // f2
      {-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0}};        // f3
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 3; ++c) { j_ang_d_[r][c] = j[r][c]; j_ang_[r][c] = (float)j[r][c]; }
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 3; ++c) { h_ang_d_[r][c] = h[r][c]; h_ang_[r][c] = (float)h[r][c]; }
}

namespace {

struct PointDerivF {
  float pg[4][6];    // point_gradient4
  float ph[24][6];   // point_hessian_
};

inline float dot3f(const float r[3], const float x[3]) { return (r[0] * x[0] + r[1] * x[1]) + r[2] * x[2]; }

// computePointDerivatives (float tables)
inline void point_derivatives(const float j_ang[8][3], const float h_ang[15][3], const float x[3], bool compute_hessian, PointDerivF& d) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 6; ++c) d.pg[r][c] = 0.f;
  d.pg[0][0] = d.pg[1][1] = d.pg[2][2] = 1.f;
  d.pg[1][3] = dot3f(j_ang[0], x);
  d.pg[2][3] = dot3f(j_ang[1], x);
  d.pg[0][4] = dot3f(j_ang[2], x);
  d.pg[1][4] = dot3f(j_ang[3], x);
  d.pg[2][4] = dot3f(j_ang[4], x);
  d.pg[0][5] = dot3f(j_ang[5], x);
  d.pg[1][5] = dot3f(j_ang[6], x);
  d.pg[2][5] = dot3f(j_ang[7], x);
  if (!compute_hessian) return;
  for (int r = 0; r < 24; ++r)
    for (int c = 0; c < 6; ++c) d.ph[r][c] = 0.f;
  float xh[15];
  for (int r = 0; r < 15; ++r) xh[r] = dot3f(h_ang[r], x);
  const float a[4] = {0.f, xh[0], xh[1], 0.f}, b[4] = {0.f, xh[2], xh[3], 0.f}, c[4] = {0.f, xh[4], xh[5], 0.f};
  const float dd[4] = {xh[6], xh[7], xh[8], 0.f}, e[4] = {xh[9], xh[10], xh[11], 0.f}, f[4] = {xh[12], xh[13], xh[14], 0.f};
  for (int r = 0; r < 4; ++r) {
    d.ph[12 + r][3] = a[r];
    d.ph[16 + r][3] = b[r];
    d.ph[20 + r][3] = c[r];
    d.ph[12 + r][4] = b[r];
    d.ph[16 + r][4] = dd[r];
    d.ph[20 + r][4] = e[r];
    d.ph[12 + r][5] = c[r];
    d.ph[16 + r][5] = e[r];
    d.ph[20 + r][5] = f[r];
  }
}

// updateDerivatives: float per-hit algebra, double accumulation targets
inline double update_derivatives(double g[6], double H[36], const PointDerivF& d, const double x_trans[3], const M3& c_inv, double gauss_d1, double gauss_d2_d, bool compute_hessian) {
  const float x4[4] = {(float)x_trans[0], (float)x_trans[1], (float)x_trans[2], 0.0f};
  float C4[4][4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) C4[r][c] = (r < 3 && c < 3) ? (float)c_inv(r, c) : 0.f;
  const float gauss_d2 = (float)gauss_d2_d;
  float xC[4];  // x_trans4 * c_inv4   (row vector times matrix)
  for (int c = 0; c < 4; ++c) xC[c] = ((x4[0] * C4[0][c] + x4[1] * C4[1][c]) + x4[2] * C4[2][c]) + x4[3] * C4[3][c];
  float xCx = ((x4[0] * xC[0] + x4[1] * xC[1]) + x4[2] * xC[2]) + x4[3] * xC[3];
  float e_x_cov_x = std::exp(-gauss_d2 * xCx * 0.5f);
  float score_inc = (float)(-gauss_d1 * e_x_cov_x);
  e_x_cov_x = gauss_d2 * e_x_cov_x;
  if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) return 0.0;
  e_x_cov_x = (float)(e_x_cov_x * gauss_d1);
  float CG[4][6];  // c_inv4 * point_gradient4
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 6; ++c) CG[r][c] = ((C4[r][0] * d.pg[0][c] + C4[r][1] * d.pg[1][c]) + C4[r][2] * d.pg[2][c]) + C4[r][3] * d.pg[3][c];
  float xCG[6];
  for (int c = 0; c < 6; ++c) xCG[c] = ((x4[0] * CG[0][c] + x4[1] * CG[1][c]) + x4[2] * CG[2][c]) + x4[3] * CG[3][c];
  for (int c = 0; c < 6; ++c) g[c] += (double)(e_x_cov_x * xCG[c]);
  if (compute_hessian) {
    float GCG[6][6];  // point_gradient4^T * (c_inv4 * point_gradient4)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) GCG[r][c] = ((d.pg[0][r] * CG[0][c] + d.pg[1][r] * CG[1][c]) + d.pg[2][r] * CG[2][c]) + d.pg[3][r] * CG[3][c];
    for (int i = 0; i < 6; ++i) {
      float xCH[6];
      for (int j = 0; j < 6; ++j) xCH[j] = ((xC[0] * d.ph[4 * i + 0][j] + xC[1] * d.ph[4 * i + 1][j]) + xC[2] * d.ph[4 * i + 2][j]) + xC[3] * d.ph[4 * i + 3][j];
      for (int j = 0; j < 6; ++j) H[6 * i + j] += (double)(e_x_cov_x * (-gauss_d2 * xCG[i] * xCG[j] + xCH[j] + GCG[j][i]));
    }
  }
  return (double)score_inc;
}

}  // namespace

double NDT::computeDerivatives(V6& g_out, M6& H_out, const Cloud& trans_cloud, const double p[6], bool compute_hessian) {
  computeAngleDerivatives(p);
  const size_t N = input_.size();
  std::vector<double> scores(N, 0.0);
  std::vector<V6> grads(N);
  std::vector<M6> hess(N);
  long hits_total = 0;
  const int mode = search_method_ == DIRECT1 ? 1 : search_method_ == DIRECT7 ? 7 : 27;
#pragma omp parallel for num_threads(threads()) schedule(guided, 8) reduction(+ : hits_total)
  for (long idx = 0; idx < (long)N; ++idx) {
    V6 g;
    M6 H;
    for (int k = 0; k < 6; ++k) g[k] = 0.0;
    for (int k = 0; k < 36; ++k) H.m[k] = 0.0;
    double score_pt = 0.0;
    const float xt[3] = {trans_cloud[idx].x, trans_cloud[idx].y, trans_cloud[idx].z};
    const Leaf* nb[27];
    std::vector<const Leaf*> nbv;
    int cnt;
    const Leaf* const* cells = nb;
    if (search_method_ == KDTREE) {
      cnt = target_cells_.radiusSearch(xt, (double)resolution_, nbv);
      cells = nbv.data();
    } else {
      cnt = target_cells_.neighborhood(xt, mode, nb);
    }
    if (cnt) {
      const float x[3] = {input_[idx].x, input_[idx].y, input_[idx].z};
      PointDerivF d;
      point_derivatives(j_ang_, h_ang_, x, compute_hessian, d);
      for (int c = 0; c < cnt; ++c) {
        const Leaf* cell = cells[c];
        const double x_trans[3] = {(double)xt[0] - cell->mean[0], (double)xt[1] - cell->mean[1], (double)xt[2] - cell->mean[2]};
        score_pt += update_derivatives(g.v, H.m, d, x_trans, cell->icov, gauss_d1_, gauss_d2_, compute_hessian);
      }
    }
    hits_total += cnt;
    scores[idx] = score_pt;
    grads[idx] = g;
    hess[idx] = H;
  }
  // "ensure that the result is invariant against the summing up order": serial, index order
  double score = 0.0;
  for (int k = 0; k < 6; ++k) g_out[k] = 0.0;
  for (int k = 0; k < 36; ++k) H_out.m[k] = 0.0;
  for (size_t i = 0; i < N; ++i) {
    score += scores[i];
    for (int k = 0; k < 6; ++k) g_out[k] += grads[i][k];
    for (int k = 0; k < 36; ++k) H_out.m[k] += hess[i].m[k];
  }
  n_eval++;
  n_hits += hits_total;
  return score;
}

// computeHessian / updateHessian: Hessian only, double arithmetic, serial
void NDT::computeHessian(M6& H, const Cloud& trans_cloud, const double /*p*/[6]) {
  for (int k = 0; k < 36; ++k) H.m[k] = 0.0;
  const int mode = search_method_ == DIRECT1 ? 1 : search_method_ == DIRECT7 ? 7 : 27;
  long hits_total = 0;
  for (size_t idx = 0; idx < input_.size(); ++idx) {
    const float xt[3] = {trans_cloud[idx].x, trans_cloud[idx].y, trans_cloud[idx].z};
    const Leaf* nb[27];
    std::vector<const Leaf*> nbv;
    int cnt;
    const Leaf* const* cells = nb;
    if (search_method_ == KDTREE) {
      cnt = target_cells_.radiusSearch(xt, (double)resolution_, nbv);
      cells = nbv.data();
    } else {
      cnt = target_cells_.neighborhood(xt, mode, nb);
    }
    if (!cnt) continue;
    hits_total += cnt;
    const double x[3] = {(double)input_[idx].x, (double)input_[idx].y, (double)input_[idx].z};
    // double point_gradient_ (3x6) and point_hessian_ (18x6)
    double J[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
    auto dj = [&](int r) { return j_ang_d_[r][0] * x[0] + j_ang_d_[r][1] * x[1] + j_ang_d_[r][2] * x[2]; };
    auto dh = [&](int r) { return h_ang_d_[r][0] * x[0] + h_ang_d_[r][1] * x[1] + h_ang_d_[r][2] * x[2]; };
    J[1][3] = dj(0); J[2][3] = dj(1); J[0][4] = dj(2); J[1][4] = dj(3); J[2][4] = dj(4); J[0][5] = dj(5); J[1][5] = dj(6); J[2][5] = dj(7);
    const double a[3] = {0, dh(0), dh(1)}, b[3] = {0, dh(2), dh(3)}, c[3] = {0, dh(4), dh(5)};
    const double d[3] = {dh(6), dh(7), dh(8)}, e[3] = {dh(9), dh(10), dh(11)}, f[3] = {dh(12), dh(13), dh(14)};
    const double* Hij[6][6] = {};
    Hij[3][3] = a; Hij[4][3] = b; Hij[5][3] = c;
    Hij[3][4] = b; Hij[4][4] = d; Hij[5][4] = e;
    Hij[3][5] = c; Hij[4][5] = e; Hij[5][5] = f;
    for (int ci = 0; ci < cnt; ++ci) {
      const Leaf* cell = cells[ci];
      const double xt_d[3] = {(double)xt[0] - cell->mean[0], (double)xt[1] - cell->mean[1], (double)xt[2] - cell->mean[2]};
      const M3& C = cell->icov;
      double Cx[3];  // c_inv * x_trans
      for (int r = 0; r < 3; ++r) Cx[r] = C(r, 0) * xt_d[0] + C(r, 1) * xt_d[1] + C(r, 2) * xt_d[2];
      double e_x_cov_x = gauss_d2_ * std::exp(-gauss_d2_ * (xt_d[0] * Cx[0] + xt_d[1] * Cx[1] + xt_d[2] * Cx[2]) / 2);
      if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;
      e_x_cov_x *= gauss_d1_;
      double CJ[3][6], xCJ[6];  // cov_dxd_pi = c_inv * J_i ; x_trans . cov_dxd_pi
      for (int i = 0; i < 6; ++i) {
        for (int r = 0; r < 3; ++r) CJ[r][i] = C(r, 0) * J[0][i] + C(r, 1) * J[1][i] + C(r, 2) * J[2][i];
        xCJ[i] = xt_d[0] * CJ[0][i] + xt_d[1] * CJ[1][i] + xt_d[2] * CJ[2][i];
      }
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double xCH = 0.0;
          if (Hij[i][j]) {
            const double* h = Hij[i][j];
            double Ch[3];
            for (int r = 0; r < 3; ++r) Ch[r] = C(r, 0) * h[0] + C(r, 1) * h[1] + C(r, 2) * h[2];
            xCH = xt_d[0] * Ch[0] + xt_d[1] * Ch[1] + xt_d[2] * Ch[2];
          }
          double JCJ = J[0][j] * CJ[0][i] + J[1][j] * CJ[1][i] + J[2][j] * CJ[2][i];
          H(i, j) += e_x_cov_x * (-gauss_d2_ * xCJ[i] * xCJ[j] + xCH + JCJ);
        }
    }
  }
  n_eval++;
  n_hits += hits_total;
}

namespace {
// More-Thuente helpers (A.4)
inline double aux_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
inline double aux_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool updateIntervalMT(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    a_u = a_t; f_u = f_t; g_u = g_t;
    return false;
  } else if (g_t * (a_l - a_t) > 0) {
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  } else if (g_t * (a_l - a_t) < 0) {
    a_u = a_l; f_u = f_l; g_u = g_l;
    a_l = a_t; f_l = f_t; g_l = g_t;
    return false;
  }
  return true;
}

double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {  // case 1
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {  // case 2
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {  // case 3
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_t_next = std::fabs(a_c - a_t) < std::fabs(a_s - a_t) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
    return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
  } else {  // case 4
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}
}  // namespace

static void transform_cloud(const Cloud& in, Cloud& out, const M4f& T) {
  out.resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    float q[3];
    m4f_apply(T, &in[i].x, q);
    out[i] = Pt{q[0], q[1], q[2], 1.0f};
  }
}

double NDT::computeStepLengthMT(const double x[6], V6& step_dir, double step_init, double step_max, double step_min, double& score, V6& g, M6& H, Cloud& trans_cloud) {
  double phi_0 = -score;
  double d_phi_0 = -v6_dot(g, step_dir);
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int k = 0; k < 6; ++k) step_dir[k] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = aux_psi(a_l, phi_0, phi_0, d_phi_0, mu), g_l = aux_dpsi(d_phi_0, d_phi_0, mu);
  double f_u = aux_psi(a_u, phi_0, phi_0, d_phi_0, mu), g_u = aux_dpsi(d_phi_0, d_phi_0, mu);
  bool interval_converged = (step_max - step_min) < 0, open_interval = true;
  double a_t = step_init;
  a_t = std::min(a_t, step_max);
  a_t = std::max(a_t, step_min);
  double x_t[6];
  for (int k = 0; k < 6; ++k) x_t[k] = x[k] + step_dir[k] * a_t;
  final_transformation_ = m4f_from_xyz_euler(x_t);
  transform_cloud(input_, trans_cloud, final_transformation_);
  score = computeDerivatives(g, H, trans_cloud, x_t, true);
  double phi_t = -score;
  double d_phi_t = -v6_dot(g, step_dir);
  double psi_t = aux_psi(a_t, phi_t, phi_0, d_phi_0, mu);
  double d_psi_t = aux_dpsi(d_phi_t, d_phi_0, mu);
  auto rec = [&]() {
    const double r[12] = {(double)nr_iterations_, (double)step_iterations, a_t, score, phi_t, d_phi_t, psi_t, d_psi_t, open_interval ? 1.0 : 0.0, interval_converged ? 1.0 : 0.0, phi_0, d_phi_0};
    trace.insert(trace.end(), r, r + 12);
  };
  rec();
  while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval)
      a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int k = 0; k < 6; ++k) x_t[k] = x[k] + step_dir[k] * a_t;
    final_transformation_ = m4f_from_xyz_euler(x_t);
    transform_cloud(input_, trans_cloud, final_transformation_);
    score = computeDerivatives(g, H, trans_cloud, x_t, false);
    phi_t = -score;
    d_phi_t = -v6_dot(g, step_dir);
    psi_t = aux_psi(a_t, phi_t, phi_0, d_phi_0, mu);
    d_psi_t = aux_dpsi(d_phi_t, d_phi_0, mu);
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
      g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
      g_u = g_u + mu * d_phi_0;
    }
    if (open_interval)
      interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else
      interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
    rec();
  }
  if (step_iterations) computeHessian(H, trans_cloud, x_t);
  return a_t;
}

void NDT::computeTransformation(Cloud& output, const M4f& guess) {
  nr_iterations_ = 0;
  converged_ = false;
  n_eval = 0;
  n_hits = 0;
  trace.clear();
  initGauss();
  if (!m4f_is_identity(guess)) {
    final_transformation_ = guess;
    transform_cloud(output, output, guess);
  }
  double p[6];
  float eul[3];
  m4f_euler_xyz(final_transformation_, eul);
  p[0] = final_transformation_(0, 3); p[1] = final_transformation_(1, 3); p[2] = final_transformation_(2, 3);
  p[3] = eul[0]; p[4] = eul[1]; p[5] = eul[2];
  V6 g, delta_p;
  M6 H;
  double score = computeDerivatives(g, H, output, p, true);
  while (!converged_) {
    previous_transformation_ = transformation_;
    V6 neg_g;
    for (int k = 0; k < 6; ++k) neg_g[k] = -g[k];
    delta_p = m6_svd_solve(H, neg_g);
    double delta_p_norm = v6_norm(delta_p);
    if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
      trans_probability_ = score / (double)input_.size();
      converged_ = delta_p_norm == delta_p_norm;
      return;
    }
    for (int k = 0; k < 6; ++k) delta_p[k] /= delta_p_norm;
    delta_p_norm = computeStepLengthMT(p, delta_p, delta_p_norm, step_size_, transformation_epsilon_ / 2, score, g, H, output);
    for (int k = 0; k < 6; ++k) delta_p[k] *= delta_p_norm;
    transformation_ = m4f_from_xyz_euler(delta_p.v);
    for (int k = 0; k < 6; ++k) p[k] += delta_p[k];
    if (nr_iterations_ > max_iterations_ || (nr_iterations_ && (std::fabs(delta_p_norm) < transformation_epsilon_))) converged_ = true;
    nr_iterations_++;
  }
  trans_probability_ = score / (double)input_.size();
}

double NDT::derivativesAt(const double p[6], double g[6], double H[36], bool compute_hessian) {
  initGauss();
  M4f T = m4f_from_xyz_euler(p);
  Cloud tc;
  transform_cloud(input_, tc, T);
  V6 gv;
  M6 Hm;
  double s = computeDerivatives(gv, Hm, tc, p, compute_hessian);
  for (int k = 0; k < 6; ++k) g[k] = gv[k];
  for (int k = 0; k < 36; ++k) H[k] = Hm.m[k];
  return s;
}

// computeHessian at an arbitrary pose (the double-precision sweep that closes a line search upstream,
// A.4 "if step_iterations: computeHessian(H, trans_cloud, x_t)")
void NDT::hessianAt(const double p[6], double H[36]) {
  initGauss();
  M4f T = m4f_from_xyz_euler(p);
  Cloud tc;
  transform_cloud(input_, tc, T);
  computeAngleDerivatives(p);
  M6 Hm;
  computeHessian(Hm, tc, p);
  for (int k = 0; k < 36; ++k) H[k] = Hm.m[k];
}

}  // namespace orc

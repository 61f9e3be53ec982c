// ORACLE — test infrastructure only (see oracle.hpp).  C entry points for the
// ctypes harness in tests/ and bench.py's cpu_baseline / --impl reference legs.
// 4x4 transforms cross this boundary as 16 floats in Eigen's column-major order.
#include <omp.h>

#include <cmath>
#include <cstring>
#include <algorithm>
#include <array>
#include <limits>
#include <vector>
#include <string>

#include "../delta_graph_slam_b200/synth/synth_scene.h"
#include "oracle.hpp"

using namespace orc;

namespace {
struct Handle {
  int method;  // 0 NDT, 1 FAST_GICP
  NDT ndt;
  FastGICP gicp;
  Cloud aligned;
  Registration& reg() { return method == 0 ? (Registration&)ndt : (Registration&)gicp; }
};
M4f from_colmajor(const float* c) {
  M4f m;
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) m(r, k) = c[4 * k + r];
  return m;
}
void to_colmajor(const M4f& m, float* c) {
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 4; ++k) c[4 * k + r] = m(r, k);
}
}  // namespace

extern "C" {

int orc_max_threads() { return omp_get_max_threads(); }
// explicit thread count for every OpenMP region of the oracle (overrides an inherited OMP_NUM_THREADS: torchrun exports 1)
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

long long orc_voxelgrid(const float* in, long long n, float lx, float ly, float lz, unsigned min_pts, int is_dense, float* out, unsigned* voxel_id, unsigned* count, unsigned* key, int* grid6, int* overflow) {
  VoxelGridResult r = voxelgrid_filter((const Pt*)in, (size_t)n, lx, ly, lz, min_pts, is_dense != 0);
  if (out && !r.out.empty()) std::memcpy(out, r.out.data(), r.out.size() * sizeof(Pt));
  if (voxel_id && !r.voxel_id.empty()) std::memcpy(voxel_id, r.voxel_id.data(), r.voxel_id.size() * 4);
  if (count && !r.count.empty()) std::memcpy(count, r.count.data(), r.count.size() * 4);
  if (key && n) std::memcpy(key, r.key.data(), (size_t)n * 4);
  if (grid6) {
    for (int a = 0; a < 3; ++a) { grid6[a] = r.min_b[a]; grid6[3 + a] = r.div_b[a]; }
  }
  if (overflow) *overflow = r.overflow ? 1 : 0;
  return (long long)r.out.size();
}

void* orc_reg_create(int method) {
  Handle* h = new Handle();
  h->method = method;
  return h;
}
void orc_reg_destroy(void* hv) { delete (Handle*)hv; }

int orc_reg_set(void* hv, const char* name, double v) {
  Handle* h = (Handle*)hv;
  std::string s(name);
  if (s == "num_threads") { h->ndt.setNumThreads((int)v); h->gicp.setNumThreads((int)v); }
  else if (s == "trans_eps") { h->reg().setTransformationEpsilon(v); }
  else if (s == "max_iter") { h->reg().setMaximumIterations((int)v); }
  else if (s == "resolution") { h->ndt.setResolution((float)v); }
  else if (s == "nn_search") { h->ndt.setNeighborhoodSearchMethod((NeighborSearchMethod)(int)v); }
  else if (s == "step_size") { h->ndt.setStepSize(v); }
  else if (s == "outlier_ratio") { h->ndt.setOutlierRatio(v); }
  else if (s == "max_corr_dist") { h->gicp.setMaxCorrespondenceDistance(v); }
  else if (s == "k_corr") { h->gicp.setCorrespondenceRandomness((int)v); }
  else if (s == "regularization") { h->gicp.setRegularizationMethod((RegularizationMethod)(int)v); }
  else if (s == "rot_eps") { h->gicp.setRotationEpsilon(v); }
  else if (s == "lsq") { h->gicp.setLsqOptimizer((LsqOptimizer)(int)v); }
  else return -1;
  return 0;
}

void orc_reg_set_target(void* hv, const float* xyzw, long long n) { ((Handle*)hv)->reg().setInputTarget((const Pt*)xyzw, (size_t)n); }
void orc_reg_set_source(void* hv, const float* xyzw, long long n) { ((Handle*)hv)->reg().setInputSource((const Pt*)xyzw, (size_t)n); }

int orc_reg_align(void* hv, const float* guess_colmajor, float* aligned_out) {
  Handle* h = (Handle*)hv;
  h->reg().align(h->aligned, from_colmajor(guess_colmajor));
  if (aligned_out && !h->aligned.empty()) std::memcpy(aligned_out, h->aligned.data(), h->aligned.size() * sizeof(Pt));
  return h->reg().hasConverged() ? 1 : 0;
}

// info: [0] transformation probability (NDT), [1] derivative passes, [2] (point,voxel) hits;
// GICP: [1] linearize calls, [2] compute_error calls
void orc_reg_get_result(void* hv, float* T_colmajor, int* converged, int* iterations, double* info3) {
  Handle* h = (Handle*)hv;
  to_colmajor(h->reg().getFinalTransformation(), T_colmajor);
  if (converged) *converged = h->reg().hasConverged() ? 1 : 0;
  if (iterations) *iterations = h->reg().getFinalNumIteration();
  if (info3) {
    if (h->method == 0) {
      info3[0] = h->ndt.getTransformationProbability(); info3[1] = (double)h->ndt.n_eval; info3[2] = (double)h->ndt.n_hits;
    } else {
      info3[0] = 0.0; info3[1] = (double)h->gicp.n_linearize; info3[2] = (double)h->gicp.n_error;
    }
  }
}

double orc_reg_fitness(void* hv, double max_range) { return ((Handle*)hv)->reg().getFitnessScore(max_range); }
double orc_reg_inlier_fraction(void* hv, const float* aligned, long long n, double max_dist) {
  Cloud c((const Pt*)aligned, (const Pt*)aligned + n);
  return ((Handle*)hv)->reg().inlierFraction(c, max_dist);
}

// ---- NDT introspection -----------------------------------------------------
long long orc_ndt_num_leaves(void* hv) { return (long long)((Handle*)hv)->ndt.cells().leaves().size(); }
void orc_ndt_grid(void* hv, int* min_b3, int* div_b3) {
  const VoxelGridCovariance& c = ((Handle*)hv)->ndt.cells();
  for (int a = 0; a < 3; ++a) { min_b3[a] = c.min_b_[a]; div_b3[a] = c.div_b_[a]; }
}
void orc_ndt_get_leaves(void* hv, unsigned long long* idx, int* npts, double* mean3, double* cov9, double* icov9, float* centroid3) {
  long long k = 0;
  for (auto& kv : ((Handle*)hv)->ndt.cells().leaves()) {
    const Leaf& l = kv.second;
    idx[k] = kv.first;
    npts[k] = l.nr_points;
    for (int a = 0; a < 3; ++a) { mean3[3 * k + a] = l.mean[a]; centroid3[3 * k + a] = l.centroid[a]; }
    for (int a = 0; a < 9; ++a) { cov9[9 * k + a] = l.cov.m[a]; icov9[9 * k + a] = l.icov.m[a]; }
    ++k;
  }
}
double orc_ndt_derivatives(void* hv, const double* p6, double* g6, double* H36, int compute_hessian) {
  return ((Handle*)hv)->ndt.derivativesAt(p6, g6, H36, compute_hessian != 0);
}

// developer trace of the last NDT align (oracle.hpp NDT::trace): returns the number of 12-double records, copies up to cap
long long orc_ndt_trace(void* hv, double* out, long long cap) {
  const std::vector<double>& t = ((Handle*)hv)->ndt.trace;
  const long long n = (long long)t.size() / 12;
  for (long long i = 0; i < n && i < cap; ++i) std::memcpy(out + 12 * i, t.data() + 12 * i, 12 * sizeof(double));
  return n;
}
void orc_ndt_hessian(void* hv, const double* p6, double* H36) { ((Handle*)hv)->ndt.hessianAt(p6, H36); }

// ---- GICP introspection ----------------------------------------------------
void orc_gicp_covariances(void* hv, int which, double* out9) {
  Handle* h = (Handle*)hv;
  const std::vector<M3>& c = which == 0 ? h->gicp.sourceCovariances() : h->gicp.targetCovariances();
  for (size_t i = 0; i < c.size(); ++i) std::memcpy(out9 + 9 * i, c[i].m, 9 * sizeof(double));
}

// ---- search ----------------------------------------------------------------
void orc_knn(const float* pts, long long n, const float* queries, long long m, int k, int* idx, float* d2) {
  KdTree t;
  t.build(pts, (size_t)n);
#pragma omp parallel for schedule(guided, 64)
  for (long long i = 0; i < m; ++i) {
    int found = t.knn(queries + 4 * i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
    for (int j = found; j < k; ++j) { idx[(size_t)k * i + j] = -1; d2[(size_t)k * i + j] = 3.402823466e+38f; }
  }
}

// ---- the rest of the prefilter chain ----------------------
// PrefilteringNodelet::distance_filter [REF apps/prefiltering_nodelet.cpp:275-291]: keep p when
// near < |p| < far, |p| = Eigen's float norm() = sqrtf((x*x + y*y) + z*z) widened to double for
// the comparison with the double thresholds; order kept; NaN / inf never pass.
long long orc_distance_filter(const float* in, long long n, double near_thresh, double far_thresh, float* out) {
  long long m = 0;
  for (long long i = 0; i < n; ++i) {
    const float x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2];
    const float sq = (x * x + y * y) + z * z;
    const double d = (double)std::sqrt(sq);
    if (d > near_thresh && d < far_thresh) {
      std::memcpy(out + 4 * m, in + 4 * i, 16);
      ++m;
    }
  }
  return m;
}

// The base_link step in front of it [REF apps/prefiltering_nodelet.cpp:123-148]:
//   pcl::transformPointCloud(*src_cloud, *transformed, transform_isometry.matrix())
// with transform_isometry an Eigen::Isometry3d, i.e. pcl::transformPointCloud<PointT, double> [UPSTREAM-RECALLED,
// pcl/common/impl/transforms.hpp, PCL 1.8-1.12]: the output starts as a copy of the input; per point (all of them when
// is_dense, the finite ones otherwise) each coordinate is
//   static_cast<float>(m(r,0) * x + m(r,1) * y + m(r,2) * z + m(r,3))
// evaluated in double, left to right (1.8: written out on coeffRefs; 1.10+: detail::Transformer<double>::se3, the same
// expression), and w = 1.  m: 16 doubles, column-major (Eigen's storage).
void orc_transform_cloud_d(const float* in, long long n, const double* m, int is_dense, float* out) {
  for (long long i = 0; i < n; ++i) {
    const float* p = in + 4 * i;
    float* q = out + 4 * i;
    std::memcpy(q, p, 16);
    if (!is_dense && !(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
    const double x = p[0], y = p[1], z = p[2];
    for (int r = 0; r < 3; ++r) q[r] = static_cast<float>(m[r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r]);
    q[3] = 1.0f;
  }
}

// pcl::RadiusOutlierRemoval<PointXYZ>::applyFilterIndices [UPSTREAM-RECALLED, PCL 1.8-1.10] as set up at
// [REF apps/prefiltering_nodelet.cpp:88-96]: k = radiusSearch(p, radius) counts the points of the SAME
// cloud with d2 < radius^2 (the query itself included); the point is kept when k > min_neighbors;
// non-finite points are dropped; order kept.
long long orc_radius_outlier_removal(const float* in, long long n, double radius, int min_neighbors, float* out) {
  // pcl::KdTreeFLANN::setInputCloud leaves non-finite points out of the tree
  std::vector<float> finite;
  finite.reserve((size_t)n * 4);
  for (long long i = 0; i < n; ++i)
    if (std::isfinite(in[4 * i]) && std::isfinite(in[4 * i + 1]) && std::isfinite(in[4 * i + 2])) finite.insert(finite.end(), in + 4 * i, in + 4 * i + 4);
  KdTree t;
  t.build(finite.data(), finite.size() / 4);
  std::vector<unsigned char> keep((size_t)(n > 0 ? n : 1), 0);
  const int k = min_neighbors + 1;
  const float r2 = (float)(radius * radius);
#pragma omp parallel for schedule(guided, 64)
  for (long long i = 0; i < n; ++i) {
    const float* q = in + 4 * i;
    if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
    std::vector<int> idx((size_t)k);
    std::vector<float> d2((size_t)k);
    const int found = t.knn(q, k, idx.data(), d2.data());
    keep[(size_t)i] = (found >= k && d2[(size_t)k - 1] < r2) ? 1 : 0;  // at least k points inside the radius <=> count > min_neighbors
  }
  long long m = 0;
  for (long long i = 0; i < n; ++i)
    if (keep[(size_t)i]) { std::memcpy(out + 4 * m, in + 4 * i, 16); ++m; }
  return m;
}

// pcl::StatisticalOutlierRemoval<PointXYZ>::applyFilterIndices as configured at
// [REF apps/prefiltering_nodelet.cpp:77-87] (mean_k 20, stddev multiplier 1.0, keep inliers):
//   per point: k = mean_k + 1 nearest neighbours on a tree of the finite points (the first one is the
//   point itself), distance_i = float( sum_{j=1..mean_k} sqrtf(d2_j) [double sum] / mean_k );
//   a non-finite point, or one whose search returns fewer than k, keeps distance 0 and is not counted;
//   sum / sq_sum (float product, double sums, INDEX order), mean = sum / valid,
//   variance = (sq_sum - sum * sum / valid) / (valid - 1), threshold = mean + mul * sqrt(variance);
//   a point is removed when distance > threshold (so the uncounted points, at 0, stay).
// stats3 (optional) receives {mean, stddev, threshold}; dist (optional) the per-point distances.
long long orc_statistical_outlier_removal(const float* in, long long n, int mean_k, double stddev_mul, float* out, double* stats3, float* dist) {
  std::vector<float> finite;
  finite.reserve((size_t)n * 4);
  for (long long i = 0; i < n; ++i)
    if (std::isfinite(in[4 * i]) && std::isfinite(in[4 * i + 1]) && std::isfinite(in[4 * i + 2])) finite.insert(finite.end(), in + 4 * i, in + 4 * i + 4);
  KdTree t;
  t.build(finite.data(), finite.size() / 4);
  const int k = mean_k + 1;
  std::vector<float> distances((size_t)(n > 0 ? n : 1), 0.0f);
  std::vector<unsigned char> valid((size_t)(n > 0 ? n : 1), 0);
#pragma omp parallel for schedule(guided, 64)
  for (long long i = 0; i < n; ++i) {
    const float* q = in + 4 * i;
    if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
    std::vector<int> idx((size_t)k);
    std::vector<float> d2((size_t)k);
    if (t.knn(q, k, idx.data(), d2.data()) != k) continue;
    double dist_sum = 0.0;
    for (int j = 1; j < k; ++j) dist_sum += std::sqrt(d2[(size_t)j]);  // float sqrt, double sum
    distances[(size_t)i] = static_cast<float>(dist_sum / mean_k);
    valid[(size_t)i] = 1;
  }
  long long valid_distances = 0;
  double sum = 0, sq_sum = 0;
  for (long long i = 0; i < n; ++i) {
    valid_distances += valid[(size_t)i];
    const float d = distances[(size_t)i];
    sum += d;
    sq_sum += d * d;  // float product
  }
  const double mean = sum / static_cast<double>(valid_distances);
  const double variance = (sq_sum - sum * sum / static_cast<double>(valid_distances)) / (static_cast<double>(valid_distances) - 1);
  const double stddev = std::sqrt(variance);
  const double threshold = mean + stddev_mul * stddev;
  if (stats3) { stats3[0] = mean; stats3[1] = stddev; stats3[2] = threshold; }
  if (dist) std::memcpy(dist, distances.data(), (size_t)n * sizeof(float));
  long long m = 0;
  for (long long i = 0; i < n; ++i) {
    if (distances[(size_t)i] > threshold) continue;
    std::memcpy(out + 4 * m, in + 4 * i, 16);
    ++m;
  }
  return m;
}

// ---- filtered2D of PrefilteringNodelet::cloud_callback [REF apps/prefiltering_nodelet.cpp:155-158] -----------------
// height_filtering (:198-214, z > lidar z) -> normal_filtering (:222-251: pcl::NormalEstimation, k = 10, keep
// |normalized n_z| < 0.2) -> flatten (:166-183, z := 0).  pcl::NormalEstimation restated from PCL 1.8-1.10
// [UPSTREAM-RECALLED]: computeMeanAndCovarianceMatrix (single pass, nine FLOAT accumulators in neighbour order),
// solvePlaneParameters -> pcl::eigen33 (matrix scaled by its largest |entry|, closed-form roots, eigenvector of the
// smallest root as the longest cross product of two rows).  flipNormalTowardsViewpoint only changes the sign, which
// |n_z| does not see.
static void pcl_compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = b * b - 4.0f * c;
  if (d < 0.0f) d = 0.0f;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
static void pcl_compute_roots(const float m[3][3], float roots[3]) {
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    pcl_compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = std::sqrt(-a_over_3);
  // atan2f / cosf / sinf taken CORRECTLY ROUNDED (double evaluation, rounded once; what glibc >= 2.41's CORE-MATH
  // routines return).  Older libms differ in the last bit (glibc 2.39's atan2f in 16 % of calls, measured here), and a
  // near-degenerate neighbourhood amplifies that bit: upstream's own normals are not bit-reproducible across hosts.
  const float theta = static_cast<float>(std::atan2(static_cast<double>(std::sqrt(-q)), static_cast<double>(half_b))) * s_inv3;
  const float cos_theta = static_cast<float>(std::cos(static_cast<double>(theta)));
  const float sin_theta = static_cast<float>(std::sin(static_cast<double>(theta)));
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0.0f) pcl_compute_roots2(c2, c1, roots);
}
// normal of the neighbourhood idx[0..cnt) (finite points); false = NaN normal upstream
static bool pcl_point_normal(const float* pts, const int* idx, int cnt, float n[3]) {
  if (cnt < 3) return false;
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < cnt; ++j) {
    const float* p = pts + 4 * (size_t)idx[j];
    accu[0] += p[0] * p[0];
    accu[1] += p[0] * p[1];
    accu[2] += p[0] * p[2];
    accu[3] += p[1] * p[1];
    accu[4] += p[1] * p[2];
    accu[5] += p[2] * p[2];
    accu[6] += p[0];
    accu[7] += p[1];
    accu[8] += p[2];
  }
  for (float& a : accu) a /= static_cast<float>(cnt);
  float c[3][3];
  c[0][0] = accu[0] - accu[6] * accu[6];
  c[0][1] = accu[1] - accu[6] * accu[7];
  c[0][2] = accu[2] - accu[6] * accu[8];
  c[1][1] = accu[3] - accu[7] * accu[7];
  c[1][2] = accu[4] - accu[7] * accu[8];
  c[2][2] = accu[5] - accu[8] * accu[8];
  c[1][0] = c[0][1]; c[2][0] = c[0][2]; c[2][1] = c[1][2];
  float scale = 0.0f;
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) scale = std::max(scale, std::fabs(c[r][k]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float m[3][3];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) m[r][k] = c[r][k] / scale;
  float roots[3];
  pcl_compute_roots(m, roots);
  for (int r = 0; r < 3; ++r) m[r][r] -= roots[0];
  auto cross = [](const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(m[0], m[1], v1);
  cross(m[0], m[2], v2);
  cross(m[1], m[2], v3);
  auto sq = [](const float* v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; };
  const float l1 = sq(v1), l2 = sq(v2), l3 = sq(v3);
  const float* v = v3;
  float l = l3;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  const float s = std::sqrt(l);
  for (int a = 0; a < 3; ++a) n[a] = v[a] / s;
  return true;
}

// out: kept points with z = 0; nz_out (optional, n floats): |normalized n_z| per INPUT point, NaN where the point
// fell to the height filter or has no normal (tests use it to tell decisions that sit on the threshold)
long long orc_flat_filter(const float* in, long long n, double lidar_z, int k, float thresh, float* out, float* nz_out) {
  std::vector<float> high;   // height_filtering output (is_dense = false: non-finite points may remain)
  std::vector<long long> src;
  for (long long i = 0; i < n; ++i)
    if (in[4 * i + 2] > lidar_z) { high.insert(high.end(), in + 4 * i, in + 4 * i + 4); src.push_back(i); }
  const long long m = (long long)src.size();
  if (nz_out) for (long long i = 0; i < n; ++i) nz_out[i] = std::numeric_limits<float>::quiet_NaN();
  // the search tree holds the finite points; indices map back into `high`
  std::vector<float> finite;
  std::vector<int> fidx;
  for (long long i = 0; i < m; ++i)
    if (std::isfinite(high[4 * i]) && std::isfinite(high[4 * i + 1]) && std::isfinite(high[4 * i + 2])) { finite.insert(finite.end(), high.begin() + 4 * i, high.begin() + 4 * i + 4); fidx.push_back((int)i); }
  KdTree t;
  t.build(finite.data(), finite.size() / 4);
  std::vector<unsigned char> keep((size_t)(m > 0 ? m : 1), 0);
#pragma omp parallel for schedule(guided, 64)
  for (long long i = 0; i < m; ++i) {
    const float* q = high.data() + 4 * i;
    if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
    std::vector<int> idx((size_t)k);
    std::vector<float> d2((size_t)k);
    const int found = t.knn(q, k, idx.data(), d2.data());
    float nrm[3];
    if (found == 0 || !pcl_point_normal(finite.data(), idx.data(), found, nrm)) continue;
    const float z2 = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];  // Eigen normalized(): v / sqrt(squaredNorm) when > 0
    const float nz = z2 > 0.0f ? nrm[2] / std::sqrt(z2) : nrm[2];
    const float a = std::fabs(nz);
    if (nz_out) nz_out[src[(size_t)i]] = a;
    keep[(size_t)i] = a < thresh ? 1 : 0;
  }
  long long o = 0;
  for (long long i = 0; i < m; ++i)
    if (keep[(size_t)i]) {
      std::memcpy(out + 4 * o, high.data() + 4 * i, 16);
      out[4 * o + 2] = 0.0f;
      ++o;
    }
  return o;
}

// ---- MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49] (the other half) ----
// Every keyframe cloud transformed by its pose (Matrix4f * Vector4f, w = 1, accumulated column by column without
// contraction), concatenated; resolution <= 0 returns that cloud, otherwise the occupied voxel centres of a
// pcl::octree::OctreePointCloud(resolution) in the order getOccupiedVoxelCenters walks them.  PCL 1.8-1.10
// [UPSTREAM-RECALLED]: points are added in order, non-finite ones skipped; the FIRST point centres a box of one
// resolution which getKeyBitSize widens to depth 1 (two cells per axis); a later point outside [min, max) re-roots the
// tree one level up, moving min down by the old side length on every axis where the point did not violate the UPPER
// bound, until it fits; a key is (unsigned)((p - min) / resolution) with the min of the moment (earlier keys gain the
// top bit of each such move); the walk is depth-first with child index = x-bit << 2 | y-bit << 1 | z-bit, i.e. ascending
// Morton order (x most significant); a centre is float((key + 0.5f) * resolution + min).
// GPU counterpart: not built yet; this restatement and its property tests come first.
long long orc_map_cloud(const float* clouds, const long long* offsets, long long n_keyframes, const float* poses_colmajor, double resolution, float* out, double* min3_depth) {
  const long long n = offsets[n_keyframes];
  std::vector<float> world((size_t)(n > 0 ? n : 1) * 4);
  for (long long kf = 0; kf < n_keyframes; ++kf) {
    const float* T = poses_colmajor + 16 * kf;
    for (long long i = offsets[kf]; i < offsets[kf + 1]; ++i) {
      const float* p = clouds + 4 * i;
      float* d = world.data() + 4 * i;
      for (int r = 0; r < 4; ++r) d[r] = ((T[r] * p[0] + T[4 + r] * p[1]) + T[8 + r] * p[2]) + T[12 + r] * 1.0f;
    }
  }
  if (!(resolution > 0.0)) {
    std::memcpy(out, world.data(), (size_t)n * 16);
    return n;
  }
  const float min_value = std::numeric_limits<float>::epsilon();
  double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  bool defined = false;
  unsigned depth = 0;
  unsigned long long shift[3] = {0, 0, 0};  // cells the origin has moved down since the first point
  struct Entry { unsigned long long k[3]; unsigned long long at[3]; };
  std::vector<Entry> entries;
  for (long long i = 0; i < n; ++i) {
    const float* p = world.data() + 4 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    while (true) {  // adoptBoundingBoxToPoint
      bool lo[3], up[3], any = false;
      for (int a = 0; a < 3; ++a) { lo[a] = p[a] < mn[a]; up[a] = p[a] >= mx[a]; any = any || lo[a] || up[a]; }
      if (!any && defined) break;
      if (defined) {
        double side = static_cast<double>(1ull << depth) * resolution;
        for (int a = 0; a < 3; ++a)
          if (!up[a]) { mn[a] -= side; shift[a] += 1ull << depth; }
        ++depth;
        side = static_cast<double>(1ull << depth) * resolution - min_value;
        for (int a = 0; a < 3; ++a) mx[a] = mn[a] + side;
      } else {
        for (int a = 0; a < 3; ++a) { mn[a] = p[a] - resolution / 2; mx[a] = p[a] + resolution / 2; }
        // getKeyBitSize on the empty tree
        unsigned max_key = 2;
        for (int a = 0; a < 3; ++a) max_key = std::max(max_key, static_cast<unsigned>(std::ceil((mx[a] - mn[a] - min_value) / resolution)));
        depth = static_cast<unsigned>(std::ceil(std::log2(static_cast<double>(max_key)) - min_value));
        const double side = static_cast<double>(1ull << depth) * resolution;
        for (int a = 0; a < 3; ++a) {
          const double over = (side - (mx[a] - mn[a])) / 2.0;
          if (over > min_value) { mn[a] -= over; mx[a] += over; }
        }
        defined = true;
      }
    }
    Entry e;
    for (int a = 0; a < 3; ++a) { e.k[a] = static_cast<unsigned>((p[a] - mn[a]) / resolution); e.at[a] = shift[a]; }
    entries.push_back(e);
  }
  if (min3_depth) { min3_depth[0] = mn[0]; min3_depth[1] = mn[1]; min3_depth[2] = mn[2]; min3_depth[3] = depth; }
  // final keys, Morton codes (x most significant per level), unique in ascending order
  std::vector<std::array<unsigned long long, 3>> keys(entries.size());
  std::vector<std::pair<unsigned __int128, size_t>> order(entries.size());
  for (size_t j = 0; j < entries.size(); ++j) {
    unsigned __int128 code = 0;
    for (int a = 0; a < 3; ++a) keys[j][a] = entries[j].k[a] + (shift[a] - entries[j].at[a]);
    for (int b = (int)depth - 1; b >= 0; --b)
      for (int a = 0; a < 3; ++a) code = (code << 1) | ((keys[j][a] >> b) & 1ull);
    order[j] = {code, j};
  }
  std::sort(order.begin(), order.end());
  long long m = 0;
  for (size_t j = 0; j < order.size(); ++j) {
    if (j && order[j].first == order[j - 1].first) continue;
    const auto& k = keys[order[j].second];
    for (int a = 0; a < 3; ++a) out[4 * m + a] = static_cast<float>((static_cast<double>(k[a]) + 0.5f) * resolution + mn[a]);
    out[4 * m + 3] = 1.0f;
    ++m;
  }
  return m;
}

// ---- linear algebra known-answer hooks -------------------------------------
void orc_sym_eigen3(const double* a9, double* evals3, double* evecs9) {
  M3 a, v;
  std::memcpy(a.m, a9, sizeof a.m);
  m3_sym_eigen(a, evals3, v);
  std::memcpy(evecs9, v.m, sizeof v.m);
}
void orc_inverse3(const double* a9, double* out9) {
  M3 a;
  std::memcpy(a.m, a9, sizeof a.m);
  M3 r = m3_inverse(a);
  std::memcpy(out9, r.m, sizeof r.m);
}
void orc_svd_solve6(const double* A36, const double* b6, double* x6) {
  M6 A; V6 b;
  std::memcpy(A.m, A36, sizeof A.m);
  std::memcpy(b.v, b6, sizeof b.v);
  V6 x = m6_svd_solve(A, b);
  std::memcpy(x6, x.v, sizeof x.v);
}
void orc_ldlt_solve6(const double* A36, const double* b6, double* x6) {
  M6 A; V6 b;
  std::memcpy(A.m, A36, sizeof A.m);
  std::memcpy(b.v, b6, sizeof b.v);
  V6 x = m6_ldlt_solve(A, b);
  std::memcpy(x6, x.v, sizeof x.v);
}
void orc_euler_xyz(const float* T_colmajor, float* out3) { m4f_euler_xyz(from_colmajor(T_colmajor), out3); }
void orc_transform_from_p(const double* p6, float* T_colmajor) { to_colmajor(m4f_from_xyz_euler(p6), T_colmajor); }

// ---- synthetic scans (CPU build of delta_graph_slam_b200/synth/synth_scene.h) ---
long long orc_synth_scan(int sensor, unsigned long long scene_seed, unsigned long long noise_seed, const double* pose_rowmajor, float* out_xyzw) {
  synth::Sensor s = sensor == 0 ? synth::sensor_hdl64() : synth::sensor_dense128();
  const long rays = (long)s.beams * s.azimuth_steps;
  std::vector<unsigned char> ok(rays);
  std::vector<float> tmp((size_t)rays * 4);
#pragma omp parallel for schedule(dynamic, 1024)
  for (long r = 0; r < rays; ++r) ok[r] = synth::scan_ray(s, scene_seed, noise_seed, pose_rowmajor, r, &tmp[4 * (size_t)r]) ? 1 : 0;
  long long n = 0;
  for (long r = 0; r < rays; ++r)
    if (ok[r]) { std::memcpy(out_xyzw + 4 * n, &tmp[4 * (size_t)r], 16); ++n; }
  return n;
}
long long orc_synth_num_rays(int sensor) {
  synth::Sensor s = sensor == 0 ? synth::sensor_hdl64() : synth::sensor_dense128();
  return (long long)s.beams * s.azimuth_steps;
}
void orc_synth_traj(long long k, unsigned long long seed, double* T_rowmajor) { synth::traj_kitti_like((long)k, seed, T_rowmajor); }
void orc_synth_pose(const double* xyzrpy, double* T_rowmajor) { synth::pose_from_xyzrpy(xyzrpy, T_rowmajor); }

}  // extern "C"

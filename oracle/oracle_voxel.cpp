// ORACLE — test infrastructure only (see oracle.hpp).
// pcl::VoxelGrid (A.1), pcl::Registration base (A.2), pclomp::VoxelGridCovariance (A.3).
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <limits>

#include "oracle.hpp"

namespace orc {

static inline bool finite3(const Pt& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// getMinMax3D + overflow guard + min_b/div_b (A.1 steps 1-3; shared with A.3 step 1)
static bool grid_bounds(const Pt* in, size_t n, bool is_dense, const float inv[3], int min_b[3], int max_b[3], int div_b[3], bool* any) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  *any = false;
  for (size_t i = 0; i < n; ++i) {
    if (!is_dense && !finite3(in[i])) continue;
    const float v[3] = {in[i].x, in[i].y, in[i].z};
    for (int a = 0; a < 3; ++a) {
      if (v[a] < mn[a]) mn[a] = v[a];
      if (v[a] > mx[a]) mx[a] = v[a];
    }
    *any = true;
  }
  if (!*any) return true;
  int64_t d[3];
  for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv[a]) + 1;
  if (d[0] * d[1] * d[2] > (int64_t)std::numeric_limits<int32_t>::max()) return false;
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv[a]);
    max_b[a] = (int)std::floor(mx[a] * inv[a]);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  return true;
}

VoxelGridResult voxelgrid_filter(const Pt* in, size_t n, float lx, float ly, float lz, unsigned min_points_per_voxel, bool is_dense) {
  VoxelGridResult r;
  const float inv[3] = {1.0f / lx, 1.0f / ly, 1.0f / lz};
  int max_b[3] = {0, 0, 0};
  r.min_b[0] = r.min_b[1] = r.min_b[2] = 0;
  r.div_b[0] = r.div_b[1] = r.div_b[2] = 0;
  r.key.assign(n, 0xFFFFFFFFu);
  bool any;
  if (!grid_bounds(in, n, is_dense, inv, r.min_b, max_b, r.div_b, &any)) {
    r.overflow = true;  // PCL_WARN "Leaf size is too small" ; output = *input_
    r.out.assign(in, in + n);
    return r;
  }
  if (!any) return r;
  const int mul[3] = {1, r.div_b[0], r.div_b[0] * r.div_b[1]};
  std::vector<std::pair<uint32_t, uint32_t>> index_vector;
  index_vector.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    if (!is_dense && !finite3(in[i])) continue;
    int ijk0 = (int)(std::floor(in[i].x * inv[0]) - (float)r.min_b[0]);
    int ijk1 = (int)(std::floor(in[i].y * inv[1]) - (float)r.min_b[1]);
    int ijk2 = (int)(std::floor(in[i].z * inv[2]) - (float)r.min_b[2]);
    uint32_t idx = (uint32_t)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]);
    r.key[i] = idx;
    index_vector.emplace_back(idx, (uint32_t)i);
  }
  // PCL sorts by idx only with an unstable std::sort; the in-voxel order is therefore
  // implementation-defined upstream.  Oracle and engine both define it as ascending
  // input index (a stable sort), see SURVEY.md A.1 "Consequences".
  std::stable_sort(index_vector.begin(), index_vector.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t i = 0;
  while (i < index_vector.size()) {
    size_t j = i + 1;
    while (j < index_vector.size() && index_vector[j].first == index_vector[i].first) ++j;
    if (j - i >= min_points_per_voxel) {
      float acc[3] = {0.f, 0.f, 0.f};
      for (size_t k = i; k < j; ++k) {
        const Pt& p = in[index_vector[k].second];
        acc[0] += p.x; acc[1] += p.y; acc[2] += p.z;
      }
      float cnt = (float)(j - i);
      r.out.push_back(Pt{acc[0] / cnt, acc[1] / cnt, acc[2] / cnt, 1.0f});
      r.voxel_id.push_back(index_vector[i].first);
      r.count.push_back((uint32_t)(j - i));
    }
    i = j;
  }
  return r;
}

// ---------------------------------------------------------------------------
int Registration::threads() const { return num_threads_ > 0 ? num_threads_ : omp_get_max_threads(); }

void Registration::setInputSource(const Pt* pts, size_t n) {
  input_.assign(pts, pts + n);
  source_cloud_updated_ = true;
}
void Registration::setInputTarget(const Pt* pts, size_t n) {
  if (n == 0) {  // PCL_ERROR "Invalid or empty point cloud dataset given!"
    return;
  }
  target_.assign(pts, pts + n);
  target_cloud_updated_ = true;
}
bool Registration::initCompute() {
  if (target_.empty()) return false;  // "No input target dataset was given!"
  if (target_cloud_updated_) {
    tree_.build(&target_[0].x, target_.size());
    target_cloud_updated_ = false;
  }
  return !input_.empty();
}
void Registration::align(Cloud& output, const M4f& guess) {
  if (!initCompute()) return;
  output = input_;
  converged_ = false;
  final_transformation_ = transformation_ = previous_transformation_ = m4f_identity();
  for (auto& p : output) p.w = 1.0f;
  computeTransformation(output, guess);
}
// getFitnessScore: identical loop in-tree at
// [REF src/hdl_graph_slam/information_matrix_calculator.cpp:77-108]
double Registration::getFitnessScore(double max_range) {
  if (target_.empty() || input_.empty()) return DBL_MAX;
  if (target_cloud_updated_) {
    tree_.build(&target_[0].x, target_.size());
    target_cloud_updated_ = false;
  }
  double fitness_score = 0.0;
  int nr = 0;
  for (size_t i = 0; i < input_.size(); ++i) {
    float q[3];
    m4f_apply(final_transformation_, &input_[i].x, q);
    int idx;
    float d2;
    tree_.knn(q, 1, &idx, &d2);
    if (d2 <= max_range) {  // squared distance compared un-squared, as upstream
      fitness_score += d2;
      nr++;
    }
  }
  return nr > 0 ? fitness_score / nr : DBL_MAX;
}
double Registration::inlierFraction(const Cloud& aligned, double max_correspondence_dist) {
  if (aligned.empty() || target_.empty()) return 0.0;
  int num_inliers = 0;
  for (size_t i = 0; i < aligned.size(); ++i) {
    int idx;
    float d2;
    tree_.knn(&aligned[i].x, 1, &idx, &d2);
    if (d2 < max_correspondence_dist * max_correspondence_dist) num_inliers++;
  }
  return (double)((float)num_inliers / (float)aligned.size());
}

// ---------------------------------------------------------------------------
void VoxelGridCovariance::build(const Pt* pts, size_t n, bool is_dense) {
  leaves_.clear();
  voxel_centroids_.clear();
  voxel_centroids_leaf_indices_.clear();
  const float inv[3] = {inv_leaf_, inv_leaf_, inv_leaf_};
  bool any;
  if (!grid_bounds(pts, n, is_dense, inv, min_b_, max_b_, div_b_, &any) || !any) {
    div_b_[0] = div_b_[1] = div_b_[2] = 0;
    kdtree_.build(nullptr, 0);
    return;
  }
  divb_mul_[0] = 1;
  divb_mul_[1] = div_b_[0];
  divb_mul_[2] = div_b_[0] * div_b_[1];
  for (size_t i = 0; i < n; ++i) {
    if (!is_dense && !finite3(pts[i])) continue;
    int ijk0 = (int)(std::floor(pts[i].x * inv[0]) - (float)min_b_[0]);
    int ijk1 = (int)(std::floor(pts[i].y * inv[1]) - (float)min_b_[1]);
    int ijk2 = (int)(std::floor(pts[i].z * inv[2]) - (float)min_b_[2]);
    size_t idx = (size_t)(ijk0 * divb_mul_[0] + ijk1 * divb_mul_[1] + ijk2 * divb_mul_[2]);
    Leaf& leaf = leaves_[idx];
    const double p[3] = {(double)pts[i].x, (double)pts[i].y, (double)pts[i].z};
    for (int a = 0; a < 3; ++a) leaf.mean[a] += p[a];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) leaf.cov(a, b) += p[a] * p[b];
    leaf.centroid[0] += pts[i].x;
    leaf.centroid[1] += pts[i].y;
    leaf.centroid[2] += pts[i].z;
    ++leaf.nr_points;
  }
  for (auto& kv : leaves_) {
    Leaf& leaf = kv.second;
    const double n_pts = (double)leaf.nr_points;
    for (int a = 0; a < 3; ++a) leaf.centroid[a] /= (float)leaf.nr_points;
    double pt_sum[3] = {leaf.mean[0], leaf.mean[1], leaf.mean[2]};
    for (int a = 0; a < 3; ++a) leaf.mean[a] /= n_pts;
    if (leaf.nr_points < min_points_per_voxel_) continue;
    voxel_centroids_.push_back(Pt{leaf.centroid[0], leaf.centroid[1], leaf.centroid[2], 1.0f});
    voxel_centroids_leaf_indices_.push_back(kv.first);
    // cov = (cov - 2 (pt_sum mean^T)) / n + mean mean^T ; cov *= (n - 1) / n
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        leaf.cov(a, b) = (leaf.cov(a, b) - 2.0 * (pt_sum[a] * leaf.mean[b])) / n_pts + leaf.mean[a] * leaf.mean[b];
    for (int k = 0; k < 9; ++k) leaf.cov.m[k] *= (n_pts - 1.0) / n_pts;
    m3_sym_eigen(leaf.cov, leaf.evals, leaf.evecs);
    if (leaf.evals[0] < 0 || leaf.evals[1] < 0 || leaf.evals[2] <= 0) {
      leaf.nr_points = -1;
      continue;
    }
    double min_covar_eigvalue = min_covar_eigvalue_mult_ * leaf.evals[2];
    if (leaf.evals[0] < min_covar_eigvalue) {
      leaf.evals[0] = min_covar_eigvalue;
      if (leaf.evals[1] < min_covar_eigvalue) leaf.evals[1] = min_covar_eigvalue;
      // cov = evecs * diag(evals) * evecs^-1   (evecs orthonormal: inverse = transpose)
      M3 d = m3_zero();
      d(0, 0) = leaf.evals[0]; d(1, 1) = leaf.evals[1]; d(2, 2) = leaf.evals[2];
      leaf.cov = m3_mul(m3_mul(leaf.evecs, d), m3_inverse(leaf.evecs));
    }
    leaf.icov = m3_inverse(leaf.cov);
    double mx = -DBL_MAX, mn = DBL_MAX;
    for (int k = 0; k < 9; ++k) { mx = std::max(mx, leaf.icov.m[k]); mn = std::min(mn, leaf.icov.m[k]); }
    if (mx == std::numeric_limits<double>::infinity() || mn == -std::numeric_limits<double>::infinity()) leaf.nr_points = -1;
  }
  kdtree_.build(voxel_centroids_.empty() ? nullptr : &voxel_centroids_[0].x, voxel_centroids_.size());
}

int VoxelGridCovariance::neighborhood(const float p[3], int mode, const Leaf** out) const {
  if (div_b_[0] == 0) return 0;
  // getNeighborhoodAtPoint*: float DIVISION by the leaf size here (multiply-by-inverse in build)
  const int ijk[3] = {(int)std::floor(p[0] / leaf_), (int)std::floor(p[1] / leaf_), (int)std::floor(p[2] / leaf_)};
  static const int off7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  int cnt = 0;
  auto probe = [&](int dx, int dy, int dz) {
    const int c[3] = {ijk[0] + dx, ijk[1] + dy, ijk[2] + dz};
    for (int a = 0; a < 3; ++a)
      if (c[a] < min_b_[a] || c[a] > max_b_[a]) return;
    size_t idx = (size_t)((c[0] - min_b_[0]) * divb_mul_[0] + (c[1] - min_b_[1]) * divb_mul_[1] + (c[2] - min_b_[2]) * divb_mul_[2]);
    auto it = leaves_.find(idx);
    if (it != leaves_.end() && it->second.nr_points >= min_points_per_voxel_) out[cnt++] = &it->second;
  };
  if (mode == 1) {
    probe(0, 0, 0);
  } else if (mode == 7) {
    for (int k = 0; k < 7; ++k) probe(off7[k][0], off7[k][1], off7[k][2]);
  } else {
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) probe(dx, dy, dz);
  }
  return cnt;
}

int VoxelGridCovariance::radiusSearch(const float p[3], double radius, std::vector<const Leaf*>& out) const {
  out.clear();
  std::vector<std::pair<float, int>> hits;
  kdtree_.radius(p, (float)(radius * radius), hits);
  for (auto& h : hits) {
    auto it = leaves_.find(voxel_centroids_leaf_indices_[h.second]);
    out.push_back(&it->second);
  }
  return (int)out.size();
}

}  // namespace orc

// ORACLE — test infrastructure only (see oracle_linalg.hpp).
//
// Exact k-d tree standing in for pcl::search::KdTree / FLANN KDTreeSingleIndex
// (SURVEY.md A.2): exact search (eps = 0), float L2_Simple distance
// ((dx*dx) + dy*dy) + dz*dz, results ascending.  Where FLANN breaks distance
// ties by traversal order (unknowable), this oracle — and the CUDA engine —
// break them by ascending point index.  Used by Registration::getFitnessScore
// [REF src/hdl_graph_slam/information_matrix_calculator.cpp:77-108 is the in-tree
// copy of that loop], by NDT's KDTREE neighbourhood search and by FastGICP.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <utility>
#include <vector>

namespace orc {

inline float l2_simple(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

class KdTree {
 public:
  // pts: n points with a stride of 4 floats (PointXYZ layout)
  void build(const float* pts, size_t n) {
    pts_ = pts;
    n_ = n;
    idx_.resize(n);
    std::iota(idx_.begin(), idx_.end(), 0);
    nodes_.clear();
    nodes_.reserve(n / 4 + 16);
    if (n) build_rec(0, (int)n);
  }
  size_t size() const { return n_; }

  // k nearest, ascending (d2, index).  Returns the number found (min(k, n)).
  int knn(const float* q, int k, int* out_idx, float* out_d2) const {
    if (!n_ || k <= 0) return 0;
    Result r{k, 0, out_idx, out_d2};
    search(0, q, r);
    return r.count;
  }
  // all points with d2 < radius2 (strict, as FLANN's RadiusResultSet), ascending
  void radius(const float* q, float radius2, std::vector<std::pair<float, int>>& out) const {
    out.clear();
    if (n_) radius_rec(0, q, radius2, out);
    std::sort(out.begin(), out.end());
  }

 private:
  struct Node {
    int lo, hi;        // index range (leaf) [lo, hi)
    int left, right;   // children, -1 for a leaf
    int axis;
    float split_lo, split_hi;  // max of left child / min of right child along axis
  };
  struct Result {
    int k, count;
    int* idx;
    float* d2;
    float worst() const { return count < k ? 3.402823466e+38f : d2[k - 1]; }
    void offer(float d, int i) {
      if (count == k) {
        if (d > d2[k - 1] || (d == d2[k - 1] && i > idx[k - 1])) return;
      }
      int pos = count < k ? count : k - 1;
      while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
        d2[pos] = d2[pos - 1];
        idx[pos] = idx[pos - 1];
        --pos;
      }
      d2[pos] = d;
      idx[pos] = i;
      if (count < k) ++count;
    }
  };

  int build_rec(int lo, int hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.f, 0.f});
    if (hi - lo <= 15) return id;  // FLANN leaf_max_size 15 as configured by PCL
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = lo; i < hi; ++i)
      for (int a = 0; a < 3; ++a) {
        float v = pts_[4 * (size_t)idx_[i] + a];
        mn[a] = std::min(mn[a], v);
        mx[a] = std::max(mx[a], v);
      }
    int axis = 0;
    if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
    if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
    if (!(mx[axis] > mn[axis])) return id;  // all points identical: keep as a leaf
    int mid = (lo + hi) / 2;
    std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi, [&](int a, int b) {
      float va = pts_[4 * (size_t)a + axis], vb = pts_[4 * (size_t)b + axis];
      return va < vb || (va == vb && a < b);
    });
    float slo = -3.4e38f, shi = 3.4e38f;
    for (int i = lo; i < mid; ++i) slo = std::max(slo, pts_[4 * (size_t)idx_[i] + axis]);
    for (int i = mid; i < hi; ++i) shi = std::min(shi, pts_[4 * (size_t)idx_[i] + axis]);
    int l = build_rec(lo, mid);
    int r = build_rec(mid, hi);
    nodes_[id].left = l;
    nodes_[id].right = r;
    nodes_[id].axis = axis;
    nodes_[id].split_lo = slo;
    nodes_[id].split_hi = shi;
    return id;
  }

  void search(int id, const float* q, Result& r) const {
    const Node& n = nodes_[id];
    if (n.left < 0) {
      for (int i = n.lo; i < n.hi; ++i) r.offer(l2_simple(q, pts_ + 4 * (size_t)idx_[i]), idx_[i]);
      return;
    }
    float v = q[n.axis];
    // distance (squared, float, monotone in the coordinate gap) to each child's slab
    float dl = v > n.split_lo ? (v - n.split_lo) * (v - n.split_lo) : 0.f;
    float dr = v < n.split_hi ? (n.split_hi - v) * (n.split_hi - v) : 0.f;
    int first = dl <= dr ? n.left : n.right, second = dl <= dr ? n.right : n.left;
    float dfirst = dl <= dr ? dl : dr, dsecond = dl <= dr ? dr : dl;
    if (!(dfirst > r.worst())) search(first, q, r);
    if (!(dsecond > r.worst())) search(second, q, r);
  }

  void radius_rec(int id, const float* q, float r2, std::vector<std::pair<float, int>>& out) const {
    const Node& n = nodes_[id];
    if (n.left < 0) {
      for (int i = n.lo; i < n.hi; ++i) {
        float d = l2_simple(q, pts_ + 4 * (size_t)idx_[i]);
        if (d < r2) out.emplace_back(d, idx_[i]);
      }
      return;
    }
    float v = q[n.axis];
    float dl = v > n.split_lo ? (v - n.split_lo) * (v - n.split_lo) : 0.f;
    float dr = v < n.split_hi ? (n.split_hi - v) * (n.split_hi - v) : 0.f;
    if (dl < r2) radius_rec(n.left, q, r2, out);
    if (dr < r2) radius_rec(n.right, q, r2, out);
  }

  const float* pts_ = nullptr;
  size_t n_ = 0;
  std::vector<int> idx_;
  std::vector<Node> nodes_;
};

}  // namespace orc

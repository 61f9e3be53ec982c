// Minimal stand-in for pcl::search::KdTree: virtual setInputCloud that "builds" (counted, so the adapter
// check can prove no tree is built on the align path), brute-force nearestKSearch / radiusSearch.
#pragma once
#include <algorithm>
#include <memory>
#include <utility>
#include <vector>

#include <pcl/point_cloud.h>
namespace pcl {
using IndicesConstPtr = shared_ptr<const std::vector<int>>;
namespace search {
template <typename PointT>
class Search {
 public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using IndicesConstPtr = pcl::IndicesConstPtr;
  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) = 0;
  virtual PointCloudConstPtr getInputCloud() const { return input_; }
  virtual int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const = 0;
  virtual int radiusSearch(const PointT& point, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int max_nn = 0) const = 0;

 protected:
  PointCloudConstPtr input_;
};
inline int& mock_tree_builds() {
  static int n = 0;
  return n;
}
template <typename PointT>
class KdTree : public Search<PointT> {
 public:
  using Base = Search<PointT>;
  using PointCloudConstPtr = typename Base::PointCloudConstPtr;
  using IndicesConstPtr = typename Base::IndicesConstPtr;
  using Ptr = shared_ptr<KdTree<PointT>>;
  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) override {
    this->input_ = cloud;
    ++mock_tree_builds();  // the FLANN build of the real class
  }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& d2) const override {
    std::vector<std::pair<float, int>> all;
    for (std::size_t i = 0; this->input_ && i < this->input_->points.size(); ++i) {
      const PointT& p = this->input_->points[i];
      const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      all.emplace_back(dx * dx + dy * dy + dz * dz, (int)i);
    }
    const std::size_t kk = std::min<std::size_t>((std::size_t)k, all.size());
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    idx.resize(kk);
    d2.resize(kk);
    for (std::size_t i = 0; i < kk; ++i) { d2[i] = all[i].first; idx[i] = all[i].second; }
    return (int)kk;
  }
  int radiusSearch(const PointT& q, double radius, std::vector<int>& idx, std::vector<float>& d2, unsigned int max_nn = 0) const override {
    std::vector<int> i2;
    std::vector<float> dd;
    nearestKSearch(q, this->input_ ? (int)this->input_->points.size() : 0, i2, dd);
    idx.clear();
    d2.clear();
    for (std::size_t i = 0; i < i2.size() && dd[i] <= (float)(radius * radius) && (!max_nn || idx.size() < max_nn); ++i) { idx.push_back(i2[i]); d2.push_back(dd[i]); }
    return (int)idx.size();
  }
};
}  // namespace search
}  // namespace pcl

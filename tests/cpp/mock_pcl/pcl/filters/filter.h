#pragma once
#include <string>

#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
class Filter {
 public:
  using PointCloud = pcl::PointCloud<PointT>;
  using Ptr = shared_ptr<Filter<PointT>>;
  virtual ~Filter() {}
  void setInputCloud(const typename PointCloud::ConstPtr& c) { input_ = c; }
  void filter(PointCloud& output) { applyFilter(output); }

 protected:
  virtual void applyFilter(PointCloud& output) = 0;
  typename PointCloud::ConstPtr input_;
  std::string filter_name_;
};
}  // namespace pcl

#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
namespace pcl {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
};
struct Vec4 { float v[4] = {0, 0, 0, 0}; };
template <typename PointT>
struct PointCloud {
  using Ptr = shared_ptr<PointCloud<PointT>>;
  using ConstPtr = shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  Vec4 sensor_origin_, sensor_orientation_;
  std::size_t size() const { return points.size(); }
};
}  // namespace pcl

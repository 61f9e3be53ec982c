// Minimal stand-in for <pcl/point_types.h> — compile check of include/b200reg_pcl.hpp only
// (PCL is not installed in this image).  Mirrors the members the adapter touches, nothing more.
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad(1.0f) {}
};
}  // namespace pcl

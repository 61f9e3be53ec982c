// Minimal stand-in for pcl::Registration: the non-virtual align() wrapper calling the protected
// pure-virtual computeTransformation(), and the protected state the adapters must maintain.
#pragma once
#include <string>

#include <pcl/point_cloud.h>
namespace pcl {
struct Matrix4fMock {  // column-major 4x4 like Eigen::Matrix4f
  float m[16];
  Matrix4fMock() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
  float& operator()(int r, int c) { return m[4 * c + r]; }
  float operator()(int r, int c) const { return m[4 * c + r]; }
  const float* data() const { return m; }
  float* data() { return m; }
};
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  using Matrix4 = Matrix4fMock;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = shared_ptr<Registration>;
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; target_cloud_updated_ = true; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void align(PointCloudSource& output, const Matrix4& guess = Matrix4()) {
    if (!input_ || !target_) return;
    output.points = input_->points;  // PCL copies the source into the output before the virtual call
    output.width = (std::uint32_t)output.points.size();
    output.height = 1;
    converged_ = false;
    computeTransformation(output, guess);
  }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  bool target_cloud_updated_ = true, converged_ = false;
  int nr_iterations_ = 0, max_iterations_ = 10;
  double transformation_epsilon_ = 0.0, corr_dist_threshold_ = 0.0;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
};
}  // namespace pcl

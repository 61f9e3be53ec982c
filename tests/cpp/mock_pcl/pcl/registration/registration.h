// Minimal stand-in for pcl::Registration: the non-virtual align() wrapper calling initCompute() (which builds
// the target kd-tree unless setSearchMethodTarget(tree, force_no_recompute = true) said otherwise) and then the
// protected pure-virtual computeTransformation(); the non-virtual getFitnessScore() over tree_; and the
// protected state the adapters must maintain.
#pragma once
#include <cfloat>
#include <string>

#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>
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
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  Registration() : tree_(new KdTree) {}
  virtual ~Registration() {}
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) {
    tree_ = tree;
    force_no_recompute_ = force_no_recompute;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  // pcl::Registration::getFitnessScore(max_range): non-virtual, serial nearest-neighbour loop over tree_
  double getFitnessScore(double max_range = DBL_MAX) {
    if (!input_ || !target_) return DBL_MAX;
    double sum = 0.0;
    int nr = 0;
    std::vector<int> idx(1);
    std::vector<float> d2(1);
    for (const PointSource& p : input_->points) {
      PointSource q;
      const Matrix4& T = final_transformation_;
      q.x = T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3);
      q.y = T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3);
      q.z = T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3);
      tree_->nearestKSearch(q, 1, idx, d2);
      if (!d2.empty() && d2[0] <= max_range) { sum += d2[0]; ++nr; }
    }
    return nr > 0 ? sum / nr : DBL_MAX;
  }
  virtual void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; target_cloud_updated_ = true; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void align(PointCloudSource& output, const Matrix4& guess = Matrix4()) {
    if (!input_ || !target_) return;
    // initCompute(): the FLANN tree over the target, rebuilt whenever the target changed — unless the
    // caller installed its own search object with force_no_recompute
    if (target_cloud_updated_ && !force_no_recompute_) {
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
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
  KdTreePtr tree_;
  bool force_no_recompute_ = false;
  bool target_cloud_updated_ = true, converged_ = false;
  int nr_iterations_ = 0, max_iterations_ = 10;
  double transformation_epsilon_ = 0.0, corr_dist_threshold_ = 0.0;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
};
}  // namespace pcl

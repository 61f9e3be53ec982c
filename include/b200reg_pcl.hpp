// b200reg_pcl.hpp — header-only pcl::Registration / pcl::Filter adapters over the b200reg C ABI.
//
// This is the C++ side of the drop-in boundary: the object
// select_registration_method() returns [REF src/hdl_graph_slam/registrations.cpp:22-124] is a
// pcl::Registration<pcl::PointXYZ, pcl::PointXYZ>::Ptr and the nodelets' down-sampler is a
// pcl::Filter<pcl::PointXYZ>::Ptr [REF apps/prefiltering_nodelet.cpp:380;
// apps/scan_matching_odometry_nodelet.cpp:391], so a C-ABI engine needs this thin C++ shim,
// compiled by the maintainer against THEIR PCL (1.8 / 1.10 / 1.11+).  patches/registrations.cpp.patch
// adds the factory branches; INTEGRATION.md walks through it.  Here it is compile-checked against a
// small mock of the PCL base classes (tests/cpp/mock_pcl), since PCL is not installed in this image.
//
// What the adapters override — exactly the virtual surface the reference reaches:
//   setInputTarget / setInputSource  (virtual since PCL 1.7)            -> b200reg_set_target / _source
//   computeTransformation(output, guess)  (pure virtual, called by align) -> b200reg_align
// and what they leave to the base class, so existing call sites keep working unchanged:
//   align(), hasConverged(), getFinalTransformation(), getFitnessScore() (non-virtual: the base
//   class's CPU kd-tree loop), getSearchMethodTarget().
// Because getFitnessScore is non-virtual, a caller that wants the GPU fitness calls
// fitnessScoreGPU() / alignAndScore() found through dynamic_cast (optional patch).
#pragma once
#include <pcl/filters/filter.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
#include <pcl/search/kdtree.h>

#include <cfloat>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200reg.h"

namespace b200reg {

static_assert(sizeof(pcl::PointXYZ) == 16, "pcl::PointXYZ is a 16-byte float4 record");

enum NeighborSearchMethod { KDTREE = B200REG_KDTREE, DIRECT26 = B200REG_DIRECT26, DIRECT7 = B200REG_DIRECT7, DIRECT1 = B200REG_DIRECT1 };  // = pclomp::NeighborSearchMethod

namespace detail {
inline b200reg_handle* create(int method, int device) {
  b200reg_config cfg;
  b200reg_default_config(method, &cfg);
  cfg.device = device;
  b200reg_handle* h = nullptr;
  // the engine never falls back to the CPU: no usable B200 -> the constructor throws, the factory's
  // caller sees it at start-up (same moment a missing ndt_omp would fail to link)
  if (b200reg_create(&cfg, &h) != B200REG_OK || !h) throw std::runtime_error("b200reg_create failed: no usable sm_100 CUDA device");
  return h;
}
inline const float* xyz(const pcl::PointCloud<pcl::PointXYZ>& c) { return c.points.empty() ? nullptr : reinterpret_cast<const float*>(c.points.data()); }
}  // namespace detail

// The search object the adapters install with setSearchMethodTarget(tree, /*force_no_recompute=*/true).
//
// pcl::Registration::align() -> initCompute() rebuilds a FLANN kd-tree over the target every time the target changed
//  — 10-20 ms of single-thread CPU work per keyframe switch [REF apps/scan_matching_odometry_nodelet.cpp:180,254]
// and per loop-closure target [REF include/hdl_graph_slam/loop_detector.hpp:124] that the engine never uses (its own exact-NN
// structure lives on the GPU).  With force_no_recompute PCL leaves the tree alone, and this subclass only REMEMBERS the
// target.  The real tree is built on first use, so the base class's non-virtual getFitnessScore() and a caller that takes
// getSearchMethodTarget() [REF apps/scan_matching_odometry_nodelet.cpp:327] still get exact answers — they pay for the tree
// only if they ask for it.
class LazyKdTree : public pcl::search::KdTree<pcl::PointXYZ> {
 public:
  using Base = pcl::search::KdTree<pcl::PointXYZ>;
  using Ptr = pcl::shared_ptr<LazyKdTree>;
  using PointCloudConstPtr = Base::PointCloudConstPtr;
  using IndicesConstPtr = Base::IndicesConstPtr;

  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) override {
    std::lock_guard<std::mutex> lock(mutex_);
    pending_ = cloud;
    pending_indices_ = indices;
    dirty_ = true;
  }
  PointCloudConstPtr getInputCloud() const override {
    std::lock_guard<std::mutex> lock(mutex_);
    return dirty_ ? pending_ : Base::getInputCloud();
  }
  int nearestKSearch(const pcl::PointXYZ& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const override {
    ensure_built();
    return Base::nearestKSearch(point, k, k_indices, k_sqr_distances);
  }
  int radiusSearch(const pcl::PointXYZ& point, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int max_nn = 0) const override {
    ensure_built();
    return Base::radiusSearch(point, radius, k_indices, k_sqr_distances, max_nn);
  }
  int builds() const { return builds_; }  // how many real trees were built (adapter check)

 private:
  void ensure_built() const {
    std::lock_guard<std::mutex> lock(mutex_);
    if (!dirty_) return;
    if (pending_) const_cast<LazyKdTree*>(this)->Base::setInputCloud(pending_, pending_indices_);
    dirty_ = false;
    ++builds_;
  }
  mutable std::mutex mutex_;
  PointCloudConstPtr pending_;
  IndicesConstPtr pending_indices_;
  mutable bool dirty_ = false;
  mutable int builds_ = 0;
};

// Common part of the two registration adapters.
class RegistrationBase : public pcl::Registration<pcl::PointXYZ, pcl::PointXYZ, float> {
 public:
  using Base = pcl::Registration<pcl::PointXYZ, pcl::PointXYZ, float>;
  using PointCloudSource = Base::PointCloudSource;
  using PointCloudSourceConstPtr = Base::PointCloudSourceConstPtr;
  using PointCloudTargetConstPtr = Base::PointCloudTargetConstPtr;
  using Matrix4 = Base::Matrix4;

  ~RegistrationBase() override {
    if (h_) b200reg_destroy(h_);
  }
  RegistrationBase(const RegistrationBase&) = delete;
  RegistrationBase& operator=(const RegistrationBase&) = delete;

  void setNumThreads(int n) { num_threads_ = n; }  // reg_num_threads: the CUDA grid replaces OpenMP

  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    Base::setInputTarget(cloud);  // keeps target_ so the base getFitnessScore / getSearchMethodTarget still work
    if (!cloud || cloud->points.empty()) return;  // PCL_ERROR + return upstream; previous target stays
    lazy_tree_->setInputCloud(cloud);  // remembered, not built (initCompute no longer touches the tree)
    if (b200reg_set_target(h_, detail::xyz(*cloud), cloud->points.size(), sizeof(pcl::PointXYZ)) != B200REG_OK) PCL_ERROR("[b200reg::setInputTarget] %s\n", b200reg_last_error(h_));
  }
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    Base::setInputSource(cloud);
    if (!cloud) return;
    if (b200reg_set_source(h_, detail::xyz(*cloud), cloud->points.size(), sizeof(pcl::PointXYZ)) != B200REG_OK) PCL_ERROR("[b200reg::setInputSource] %s\n", b200reg_last_error(h_));
  }

  // getFitnessScore(max_range) on the GPU with the last final transformation
  double fitnessScoreGPU(double max_range = DBL_MAX) {
    double v = DBL_MAX;
    b200reg_get_fitness_score(h_, max_range, &v);
    return v;
  }
  // align + getFitnessScore in one call (what LoopDetector::matching does per candidate,
  // [REF include/hdl_graph_slam/loop_detector.hpp:145-148])
  double alignAndScore(PointCloudSource& output, const Matrix4& guess, double max_range = DBL_MAX) {
    this->align(output, guess);
    return fitnessScoreGPU(max_range);
  }
  int getFinalNumIteration() const { return this->nr_iterations_; }
  b200reg_handle* handle() { return h_; }
  const LazyKdTree& lazyTree() const { return *lazy_tree_; }

 protected:
  RegistrationBase(int method, int device) : h_(detail::create(method, device)), lazy_tree_(new LazyKdTree) {
    // no FLANN build inside align(): see LazyKdTree
    this->setSearchMethodTarget(lazy_tree_, /*force_no_recompute=*/true);
  }

  void push_common() {
    b200reg_set_transformation_epsilon(h_, this->transformation_epsilon_);
    b200reg_set_maximum_iterations(h_, this->max_iterations_);
  }

  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    push_common();
    push_params();
    this->converged_ = false;
    this->nr_iterations_ = 0;
    // Eigen::Matrix4f is column-major: guess.data() is the layout the C ABI takes
    float* out_xyz = output.points.empty() ? nullptr : reinterpret_cast<float*>(output.points.data());
    if (b200reg_align(h_, guess.data(), out_xyz) != B200REG_OK) {
      PCL_ERROR("[b200reg::computeTransformation] %s\n", b200reg_last_error(h_));
      return;
    }
    b200reg_result r;
    if (b200reg_get_result(h_, &r) != B200REG_OK) return;
    this->previous_transformation_ = this->transformation_;
    for (int c = 0; c < 4; ++c)
      for (int rr = 0; rr < 4; ++rr) this->final_transformation_(rr, c) = r.transformation[4 * c + rr];
    this->transformation_ = this->final_transformation_;
    this->converged_ = r.converged != 0;
    this->nr_iterations_ = r.iterations;
    last_ = r;
  }

  virtual void push_params() = 0;

  b200reg_handle* h_ = nullptr;
  LazyKdTree::Ptr lazy_tree_;
  b200reg_result last_{};
  int num_threads_ = 0;
};

// Replaces pclomp::NormalDistributionsTransform ("NDT_OMP") [REF src/hdl_graph_slam/registrations.cpp:105-119]
class NormalDistributionsTransform : public RegistrationBase {
 public:
  using Ptr = pcl::shared_ptr<NormalDistributionsTransform>;
  explicit NormalDistributionsTransform(int device = 0) : RegistrationBase(B200REG_METHOD_NDT, device) {
    this->reg_name_ = "b200reg::NormalDistributionsTransform";
    this->transformation_epsilon_ = 0.1;  // pclomp defaults
    this->max_iterations_ = 35;
  }
  void setResolution(float r) { resolution_ = r; b200reg_set_resolution(h_, r); }
  float getResolution() const { return resolution_; }
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { b200reg_set_nn_search(h_, (int)m); }
  void setNeighborhoodSearchMethod(int pclomp_enum_value) { b200reg_set_nn_search(h_, pclomp_enum_value); }
  double getTransformationProbability() const { return last_.score; }

 protected:
  void push_params() override {}
  float resolution_ = 1.0f;
};

// Replaces fast_gicp::FastGICP ("FAST_GICP") [REF src/hdl_graph_slam/registrations.cpp:27-36]
class FastGICP : public RegistrationBase {
 public:
  using Ptr = pcl::shared_ptr<FastGICP>;
  explicit FastGICP(int device = 0) : RegistrationBase(B200REG_METHOD_GICP, device) {
    this->reg_name_ = "b200reg::FastGICP";
    this->transformation_epsilon_ = 5e-4;  // fast_gicp::LsqRegistration defaults
    this->max_iterations_ = 64;
  }
  void setMaxCorrespondenceDistance(double d) {
    this->corr_dist_threshold_ = d;
    b200reg_set_max_correspondence_distance(h_, d);
  }
  void setCorrespondenceRandomness(int k) { b200reg_set_correspondence_randomness(h_, k); }

 protected:
  void push_params() override {}
};

// Replaces pcl::VoxelGrid<pcl::PointXYZ> behind pcl::Filter::Ptr
// [REF apps/prefiltering_nodelet.cpp:59-63,249-260; apps/scan_matching_odometry_nodelet.cpp:85-89,155-165]
class VoxelGrid : public pcl::Filter<pcl::PointXYZ> {
 public:
  using PointCloud = pcl::PointCloud<pcl::PointXYZ>;
  explicit VoxelGrid(int device = 0) : h_(detail::create(B200REG_METHOD_NONE, device)) { this->filter_name_ = "b200reg::VoxelGrid"; }
  ~VoxelGrid() override {
    if (h_) b200reg_destroy(h_);
  }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setMinimumPointsNumberPerVoxel(unsigned n) { min_points_ = n; }
  // PrefilteringNodelet::distance_filter fused into this filter [REF apps/prefiltering_nodelet.cpp:100-102,275-291]
  void setDistanceFilter(bool use, double near_thresh, double far_thresh) { b200reg_set_distance_filter(h_, use ? 1 : 0, near_thresh, far_thresh); }
  // the base_link step of cloud_callback in front of this filter [REF apps/prefiltering_nodelet.cpp:123-148]:
  // setInputTransform(transform_isometry.matrix().data()) — 16 doubles, column-major, x / y translation already zeroed —
  // replaces pcl::transformPointCloud(*src_cloud, *transformed, ...); nullptr switches it off
  void setInputTransform(const double* matrix4x4_colmajor) { b200reg_set_input_transform(h_, matrix4x4_colmajor); }
  b200reg_handle* handle() { return h_; }

 protected:
  void applyFilter(PointCloud& output) override {
    const PointCloud& in = *this->input_;
    output.header = in.header;
    output.sensor_origin_ = in.sensor_origin_;
    output.sensor_orientation_ = in.sensor_orientation_;
    output.points.resize(in.points.size());
    size_t n_out = 0;
    int rc = b200reg_voxelgrid_filter(h_, detail::xyz(in), in.points.size(), sizeof(pcl::PointXYZ), leaf_, min_points_, in.is_dense ? 1 : 0,
                                      output.points.empty() ? nullptr : reinterpret_cast<float*>(output.points.data()), output.points.size(), &n_out);
    if (rc != B200REG_OK) {
      PCL_ERROR("[b200reg::VoxelGrid] %s\n", b200reg_last_error(h_));
      n_out = 0;
    }
    output.points.resize(n_out);
    output.width = static_cast<uint32_t>(n_out);
    output.height = 1;
    output.is_dense = true;
  }
  b200reg_handle* h_ = nullptr;
  float leaf_[3] = {0.f, 0.f, 0.f};
  unsigned min_points_ = 0;
};

// Replaces pcl::RadiusOutlierRemoval<pcl::PointXYZ> behind pcl::Filter::Ptr (outlier_removal_method RADIUS)
// [REF apps/prefiltering_nodelet.cpp:88-96,262-273]
class RadiusOutlierRemoval : public pcl::Filter<pcl::PointXYZ> {
 public:
  using PointCloud = pcl::PointCloud<pcl::PointXYZ>;
  explicit RadiusOutlierRemoval(int device = 0) : h_(detail::create(B200REG_METHOD_NONE, device)) { this->filter_name_ = "b200reg::RadiusOutlierRemoval"; }
  ~RadiusOutlierRemoval() override {
    if (h_) b200reg_destroy(h_);
  }
  void setRadiusSearch(double radius) { radius_ = radius; }
  void setMinNeighborsInRadius(int n) { min_neighbors_ = n; }
  b200reg_handle* handle() { return h_; }

 protected:
  void applyFilter(PointCloud& output) override {
    const PointCloud& in = *this->input_;
    output.header = in.header;
    output.sensor_origin_ = in.sensor_origin_;
    output.sensor_orientation_ = in.sensor_orientation_;
    output.points.resize(in.points.size());
    size_t n_out = 0;
    int rc = b200reg_radius_outlier_removal(h_, detail::xyz(in), in.points.size(), sizeof(pcl::PointXYZ), radius_, min_neighbors_,
                                            output.points.empty() ? nullptr : reinterpret_cast<float*>(output.points.data()), output.points.size(), &n_out);
    if (rc != B200REG_OK) {
      PCL_ERROR("[b200reg::RadiusOutlierRemoval] %s\n", b200reg_last_error(h_));
      n_out = 0;
    }
    output.points.resize(n_out);
    output.width = static_cast<uint32_t>(n_out);
    output.height = 1;
    output.is_dense = true;
  }
  b200reg_handle* h_ = nullptr;
  double radius_ = 0.0;
  int min_neighbors_ = 1;
};

// Replaces pcl::StatisticalOutlierRemoval<pcl::PointXYZ> behind pcl::Filter::Ptr (outlier_removal_method STATISTICAL,
// the nodelet's default) [REF apps/prefiltering_nodelet.cpp:77-87,262-273]
class StatisticalOutlierRemoval : public pcl::Filter<pcl::PointXYZ> {
 public:
  using PointCloud = pcl::PointCloud<pcl::PointXYZ>;
  explicit StatisticalOutlierRemoval(int device = 0) : h_(detail::create(B200REG_METHOD_NONE, device)) { this->filter_name_ = "b200reg::StatisticalOutlierRemoval"; }
  ~StatisticalOutlierRemoval() override {
    if (h_) b200reg_destroy(h_);
  }
  void setMeanK(int k) { mean_k_ = k; }
  int getMeanK() const { return mean_k_; }
  void setStddevMulThresh(double m) { std_mul_ = m; }
  double getStddevMulThresh() const { return std_mul_; }
  b200reg_handle* handle() { return h_; }

 protected:
  void applyFilter(PointCloud& output) override {
    const PointCloud& in = *this->input_;
    output.header = in.header;
    output.sensor_origin_ = in.sensor_origin_;
    output.sensor_orientation_ = in.sensor_orientation_;
    output.points.resize(in.points.size());
    size_t n_out = 0;
    int rc = b200reg_statistical_outlier_removal(h_, detail::xyz(in), in.points.size(), sizeof(pcl::PointXYZ), mean_k_, std_mul_,
                                                 output.points.empty() ? nullptr : reinterpret_cast<float*>(output.points.data()), output.points.size(), &n_out);
    if (rc != B200REG_OK) {
      PCL_ERROR("[b200reg::StatisticalOutlierRemoval] %s\n", b200reg_last_error(h_));
      n_out = 0;
    }
    output.points.resize(n_out);
    output.width = static_cast<uint32_t>(n_out);
    output.height = 1;
    output.is_dense = in.is_dense;  // non-finite input points are kept, as upstream
  }
  b200reg_handle* h_ = nullptr;
  int mean_k_ = 1;       // pcl's defaults; the nodelet sets 20 / 1.0
  double std_mul_ = 0.0;
};

// filtered2D of PrefilteringNodelet::cloud_callback behind pcl::Filter::Ptr: height_filtering -> normal_filtering ->
// flatten [REF apps/prefiltering_nodelet.cpp:155-158, 166-251] in one engine call (b200reg_flat_filter)
class FlatFilter : public pcl::Filter<pcl::PointXYZ> {
 public:
  using PointCloud = pcl::PointCloud<pcl::PointXYZ>;
  explicit FlatFilter(int device = 0) : h_(detail::create(B200REG_METHOD_NONE, device)) { this->filter_name_ = "b200reg::FlatFilter"; }
  ~FlatFilter() override {
    if (h_) b200reg_destroy(h_);
  }
  void setLidarHeight(double lidar_z) { lidar_z_ = lidar_z; }       // lidar_position.z() (:143)
  void setKSearch(int k) { k_ = k; }                                 // ne.setKSearch(10) (:231)
  void setNormalThreshold(double t) { thresh_ = t; }                 // normal_filter_thresh = 0.2f (:241)
  b200reg_handle* handle() { return h_; }

 protected:
  void applyFilter(PointCloud& output) override {
    const PointCloud& in = *this->input_;
    output.header = in.header;
    output.sensor_origin_ = in.sensor_origin_;
    output.sensor_orientation_ = in.sensor_orientation_;
    output.points.resize(in.points.size());
    size_t n_out = 0;
    int rc = b200reg_flat_filter(h_, detail::xyz(in), in.points.size(), sizeof(pcl::PointXYZ), lidar_z_, k_, thresh_,
                                 output.points.empty() ? nullptr : reinterpret_cast<float*>(output.points.data()), output.points.size(), &n_out);
    if (rc != B200REG_OK) {
      PCL_ERROR("[b200reg::FlatFilter] %s\n", b200reg_last_error(h_));
      n_out = 0;
    }
    output.points.resize(n_out);
    output.width = static_cast<uint32_t>(n_out);
    output.height = 1;
    output.is_dense = false;  // as flatten() leaves it (:178)
  }
  b200reg_handle* h_ = nullptr;
  double lidar_z_ = 0.0;
  int k_ = 10;
  double thresh_ = 0.2;
};

}  // namespace b200reg

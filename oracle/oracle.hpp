// ORACLE — test infrastructure only (see oracle_linalg.hpp).
//
// CPU restatement of the scan-registration hot path of delta_graph_slam:
//   pcl::VoxelGrid                        via [REF apps/prefiltering_nodelet.cpp:59-63,249-260]
//   pcl::Registration align / fitness     via [REF apps/scan_matching_odometry_nodelet.cpp:180-228,318]
//   pclomp::VoxelGridCovariance           via ndt->setInputTarget
//   pclomp::NormalDistributionsTransform  via [REF src/hdl_graph_slam/registrations.cpp:93-120]
//   fast_gicp::FastGICP / LsqRegistration via [REF src/hdl_graph_slam/registrations.cpp:27-36]
//
// PARITY UNPINNED.  The arithmetic of this path lives in three third-party
// libraries (koide3/ndt_omp, SMRT-AIST/fast_gicp, PCL) that the reference pulls
// un-pinned at docker-build time [REF docker/noetic/Dockerfile:14-15], whose
// sources are not under /root/reference and are not installed in this image,
// and the reference holds no tests or golden vectors.  This is a restatement
// of their published algorithms from SURVEY.md Appendix A, pinned only by its
// own known-answer tests (tests/test_oracle_*.py) and by the one in-tree copy
// of the fitness loop [REF src/hdl_graph_slam/information_matrix_calculator.cpp:77-108].
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <map>
#include <vector>

#include "oracle_kdtree.hpp"
#include "oracle_linalg.hpp"

namespace orc {

struct Pt {
  float x, y, z, w;
};  // pcl::PointXYZ: 16 bytes, data[3] = 1.0f
using Cloud = std::vector<Pt>;

// ---------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZ>::applyFilter  (A.1)
// ---------------------------------------------------------------------------
struct VoxelGridResult {
  Cloud out;                      // centroids, ascending linear voxel index
  std::vector<uint32_t> voxel_id; // per OUTPUT voxel: linear index
  std::vector<uint32_t> count;    // per OUTPUT voxel: points inside
  std::vector<uint32_t> key;      // per INPUT point: linear index (0xFFFFFFFF if skipped)
  int min_b[3], div_b[3];
  bool overflow = false;          // leaf too small: output = input copy
};
VoxelGridResult voxelgrid_filter(const Pt* in, size_t n, float lx, float ly, float lz, unsigned min_points_per_voxel, bool is_dense);

// ---------------------------------------------------------------------------
// pcl::Registration<PointXYZ, PointXYZ, float>  (A.2)
// ---------------------------------------------------------------------------
class Registration {
 public:
  virtual ~Registration() {}
  virtual void setInputSource(const Pt* pts, size_t n);
  virtual void setInputTarget(const Pt* pts, size_t n);
  void align(Cloud& output, const M4f& guess);
  bool hasConverged() const { return converged_; }
  M4f getFinalTransformation() const { return final_transformation_; }
  double getFitnessScore(double max_range = DBL_MAX);
  // second half of publish_scan_matching_status [REF apps/scan_matching_odometry_nodelet.cpp:320-332]
  double inlierFraction(const Cloud& aligned, double max_correspondence_dist);
  int getFinalNumIteration() const { return nr_iterations_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  int num_threads_ = 0;  // 0 = omp_get_max_threads()
  void setNumThreads(int n) { num_threads_ = n; }
  int threads() const;
  const KdTree& searchMethodTarget() const { return tree_; }

 protected:
  virtual void computeTransformation(Cloud& output, const M4f& guess) = 0;
  bool initCompute();
  Cloud input_, target_;
  KdTree tree_;
  bool target_cloud_updated_ = true, source_cloud_updated_ = true;
  bool converged_ = false;
  int nr_iterations_ = 0, max_iterations_ = 10;
  double transformation_epsilon_ = 0.0;
  M4f final_transformation_ = m4f_identity(), transformation_ = m4f_identity(), previous_transformation_ = m4f_identity();
};

// ---------------------------------------------------------------------------
// pclomp::VoxelGridCovariance<PointXYZ>  (A.3)
// ---------------------------------------------------------------------------
struct Leaf {
  int nr_points = 0;
  double mean[3] = {0, 0, 0};
  float centroid[3] = {0, 0, 0};
  M3 cov = m3_zero(), icov = m3_zero(), evecs = m3_identity();
  double evals[3] = {0, 0, 0};
};
class VoxelGridCovariance {
 public:
  int min_points_per_voxel_ = 6;
  double min_covar_eigvalue_mult_ = 0.01;
  void setLeafSize(float l) { leaf_ = l; inv_leaf_ = 1.0f / l; }
  void build(const Pt* pts, size_t n, bool is_dense = true);  // setInputCloud + filter(true)
  // neighbourhood look-ups; all append valid Leaf* (nr_points >= min_points_per_voxel_)
  int neighborhood(const float p[3], int mode /*1,7,27*/, const Leaf** out) const;
  int radiusSearch(const float p[3], double radius, std::vector<const Leaf*>& out) const;
  const std::map<size_t, Leaf>& leaves() const { return leaves_; }
  int min_b_[3] = {0, 0, 0}, max_b_[3] = {0, 0, 0}, div_b_[3] = {0, 0, 0}, divb_mul_[3] = {0, 0, 0};
  float leaf_ = 1.f, inv_leaf_ = 1.f;
  Cloud voxel_centroids_;
  std::vector<size_t> voxel_centroids_leaf_indices_;

 private:
  std::map<size_t, Leaf> leaves_;
  KdTree kdtree_;
};

// ---------------------------------------------------------------------------
// pclomp::NormalDistributionsTransform  (A.4)
// ---------------------------------------------------------------------------
enum NeighborSearchMethod { KDTREE = 0, DIRECT26 = 1, DIRECT7 = 2, DIRECT1 = 3 };

class NDT : public Registration {
 public:
  NDT() { transformation_epsilon_ = 0.1; max_iterations_ = 35; }
  void setResolution(float r);
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { search_method_ = m; }
  void setStepSize(double s) { step_size_ = s; }
  void setOutlierRatio(double r) { outlier_ratio_ = r; }
  void setInputTarget(const Pt* pts, size_t n) override;
  double getTransformationProbability() const { return trans_probability_; }
  const VoxelGridCovariance& cells() const { return target_cells_; }
  // exposed for known-answer tests: score / gradient / Hessian of the source at pose p
  double derivativesAt(const double p[6], double g[6], double H[36], bool compute_hessian);
  void hessianAt(const double p[6], double H[36]);
  long n_eval = 0;  // derivative passes in the last align (roofline accounting)
  long n_hits = 0;  // (point, voxel) pairs visited in the last align
  // developer trace of the last align: one record of 12 doubles per line-search evaluation
  // {nr_iterations, step_iterations, a_t, score, phi_t, d_phi_t, psi_t, d_psi_t, open_interval, interval_converged, phi_0, d_phi_0}
  std::vector<double> trace;

 protected:
  void computeTransformation(Cloud& output, const M4f& guess) override;

 private:
  void init();
  void initGauss();
  void computeAngleDerivatives(const double p[6]);
  double computeDerivatives(V6& g, M6& H, const Cloud& trans_cloud, const double p[6], bool compute_hessian);
  void computeHessian(M6& H, const Cloud& trans_cloud, const double p[6]);
  double computeStepLengthMT(const double x[6], V6& step_dir, double step_init, double step_max, double step_min, double& score, V6& g, M6& H, Cloud& trans_cloud);
  float resolution_ = 1.0f;
  double step_size_ = 0.1, outlier_ratio_ = 0.55;
  double gauss_d1_ = 0, gauss_d2_ = 0, trans_probability_ = 0;
  NeighborSearchMethod search_method_ = DIRECT7;
  VoxelGridCovariance target_cells_;
  // angle-derivative tables (rows dotted with the original source point)
  float j_ang_[8][3];
  float h_ang_[15][3];
  double j_ang_d_[8][3], h_ang_d_[15][3];
};

// ---------------------------------------------------------------------------
// fast_gicp::FastGICP on fast_gicp::LsqRegistration  (A.5)
// ---------------------------------------------------------------------------
enum RegularizationMethod { REG_NONE = 0, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS };
enum LsqOptimizer { LSQ_GN = 0, LSQ_LM = 1 };

class FastGICP : public Registration {
 public:
  FastGICP() { max_iterations_ = 64; transformation_epsilon_ = 5e-4; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  void setCorrespondenceRandomness(int k) { k_correspondences_ = k; }
  void setRegularizationMethod(RegularizationMethod m) { regularization_method_ = m; }
  void setRotationEpsilon(double e) { rotation_epsilon_ = e; }
  void setLsqOptimizer(LsqOptimizer o) { lsq_ = o; }
  void setInputSource(const Pt* pts, size_t n) override;
  void setInputTarget(const Pt* pts, size_t n) override;
  // exposed for parity tests: 3x3 row-major covariance per point
  const std::vector<M3>& sourceCovariances();
  const std::vector<M3>& targetCovariances();
  long n_linearize = 0, n_error = 0;

 protected:
  void computeTransformation(Cloud& output, const M4f& guess) override;

 private:
  struct Iso {  // Eigen::Isometry3d
    double R[9];
    double t[3];
  };
  void calculate_covariances(const Cloud& cloud, const KdTree& tree, std::vector<M3>& covs);
  void update_correspondences(const Iso& trans);
  double linearize(const Iso& trans, M6* H, V6* b);
  double compute_error(const Iso& trans);
  bool step_lm(Iso& x0, Iso& delta);
  bool step_gn(Iso& x0, Iso& delta);
  bool is_converged(const Iso& delta) const;
  static Iso se3_exp(const V6& a);
  static Iso iso_mul(const Iso& a, const Iso& b);
  int k_correspondences_ = 20;
  double corr_dist_threshold_ = FLT_MAX;
  double rotation_epsilon_ = 2e-3;
  RegularizationMethod regularization_method_ = REG_PLANE;
  LsqOptimizer lsq_ = LSQ_LM;
  int lm_max_iterations_ = 10;
  double lm_init_lambda_factor_ = 1e-9, lm_lambda_ = -1.0;
  KdTree source_kdtree_, target_kdtree_;
  std::vector<M3> source_covs_, target_covs_;
  std::vector<int> correspondences_;
  std::vector<float> sq_distances_;
  std::vector<M3> mahalanobis_;
};

}  // namespace orc

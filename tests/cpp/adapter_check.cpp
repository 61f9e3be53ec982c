// Compile / link check of include/b200reg_pcl.hpp against the mock PCL, mirroring how
// select_registration_method() [REF src/hdl_graph_slam/registrations.cpp:22-124] and
// ScanMatchingOdometryNodelet::matching [REF apps/scan_matching_odometry_nodelet.cpp:173-270] use the
// objects.  Without a GPU the constructors must throw (no CPU fallback): exit code 3.  With a GPU it
// runs one tiny registration through the pcl::Registration surface: exit code 0.
#include <b200reg_pcl.hpp>

#include <cmath>
#include <cstdio>

using PointT = pcl::PointXYZ;

int main() {
  pcl::Registration<PointT, PointT>::Ptr registration;
  pcl::Filter<PointT>::Ptr downsample_filter;
  try {
    b200reg::NormalDistributionsTransform::Ptr ndt(new b200reg::NormalDistributionsTransform());
    ndt->setNumThreads(0);
    ndt->setTransformationEpsilon(0.01);
    ndt->setMaximumIterations(64);
    ndt->setResolution(1.0f);
    ndt->setNeighborhoodSearchMethod(b200reg::DIRECT7);
    registration = ndt;
    b200reg::FastGICP::Ptr gicp(new b200reg::FastGICP());
    gicp->setMaxCorrespondenceDistance(2.5);
    gicp->setCorrespondenceRandomness(20);
    auto vg = std::make_shared<b200reg::VoxelGrid>();
    vg->setLeafSize(0.1f, 0.1f, 0.1f);
    vg->setDistanceFilter(true, 0.1, 100.0);
    downsample_filter = vg;
    auto rad = std::make_shared<b200reg::RadiusOutlierRemoval>();
    rad->setRadiusSearch(0.5);
    rad->setMinNeighborsInRadius(2);
    pcl::Filter<PointT>::Ptr outlier_removal_filter = rad;
    auto sor = std::make_shared<b200reg::StatisticalOutlierRemoval>();
    sor->setMeanK(20);
    sor->setStddevMulThresh(1.0);
    outlier_removal_filter = sor;
    (void)outlier_removal_filter;
    auto flat = std::make_shared<b200reg::FlatFilter>();
    flat->setLidarHeight(0.0);
    flat->setKSearch(10);
    flat->setNormalThreshold(0.2);
    pcl::Filter<PointT>::Ptr flat_filter = flat;
    (void)flat_filter;
  } catch (const std::exception& e) {
    std::printf("no engine: %s\n", e.what());
    return 3;
  }
  // a wavy surface patch, shifted by 0.2 m
  auto make = [](float dx) {
    pcl::PointCloud<PointT>::Ptr c(new pcl::PointCloud<PointT>());
    for (int i = 0; i < 120; ++i)
      for (int j = 0; j < 120; ++j) {
        float x = 0.1f * i, y = 0.1f * j;
        c->points.emplace_back(x + dx, y, 0.4f * std::sin(x) * std::cos(0.7f * y));
        c->points.emplace_back(x + dx, 0.3f * std::sin(0.5f * x), y * 0.5f);
      }
    c->width = (std::uint32_t)c->points.size();
    c->height = 1;
    return c;
  };
  pcl::PointCloud<PointT>::Ptr target = make(0.f), source = make(0.2f), filtered(new pcl::PointCloud<PointT>()), aligned(new pcl::PointCloud<PointT>());
  downsample_filter->setInputCloud(source);
  downsample_filter->filter(*filtered);
  registration->setInputTarget(target);
  registration->setInputSource(filtered);
  registration->align(*aligned, pcl::Registration<PointT, PointT>::Matrix4());
  auto T = registration->getFinalTransformation();
  std::printf("converged=%d tx=%.4f (expect about -0.2) filtered=%zu aligned=%zu\n", (int)registration->hasConverged(), T(0, 3), filtered->size(), aligned->size());
  auto* b = dynamic_cast<b200reg::RegistrationBase*>(registration.get());
  std::printf("fitness(GPU)=%.6f\n", b ? b->fitnessScoreGPU() : -1.0);
  return (registration->hasConverged() && std::fabs(T(0, 3) + 0.2f) < 0.05f) ? 0 : 1;
}

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
  b200reg::FastGICP::Ptr gicp_keep;
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
    gicp->setTransformationEpsilon(0.01);
    gicp_keep = gicp;
    auto vg = std::make_shared<b200reg::VoxelGrid>();
    vg->setLeafSize(0.1f, 0.1f, 0.1f);
    vg->setDistanceFilter(true, 0.1, 100.0);
    {  // the base_link step [REF apps/prefiltering_nodelet.cpp:123-148]: set (identity here) and switched off again
      const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      vg->setInputTransform(I);
      vg->setInputTransform(nullptr);
    }
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
  int failures = 0;
  auto expect = [&](bool ok, const char* what) {
    if (!ok) { std::printf("FAILED: %s\n", what); ++failures; }
  };
  // the odometry's call sequence [REF apps/scan_matching_odometry_nodelet.cpp:180-228]; twice, with a target change in
  // between (a keyframe switch): pcl::Registration::initCompute must not build a kd-tree for either
  const int builds_before = pcl::search::mock_tree_builds();
  for (int round = 0; round < 2; ++round) {
    registration->setInputTarget(target);
    registration->setInputSource(filtered);
    registration->align(*aligned, pcl::Registration<PointT, PointT>::Matrix4());
  }
  expect(pcl::search::mock_tree_builds() == builds_before, "align() built a CPU kd-tree over the target (setSearchMethodTarget(tree, true) not in effect)");
  auto T = registration->getFinalTransformation();
  std::printf("converged=%d tx=%.4f (expect about -0.2) filtered=%zu aligned=%zu\n", (int)registration->hasConverged(), T(0, 3), filtered->size(), aligned->size());
  expect(registration->hasConverged() && std::fabs(T(0, 3) + 0.2f) < 0.05f, "registration did not recover the 0.2 m shift");
  // align(*aligned, guess) fills `aligned` with the transformed source [REF :217-218]
  expect(aligned->size() == filtered->size(), "aligned cloud has the wrong size");
  double worst = 0.0;
  for (std::size_t i = 0; i < aligned->size() && i < filtered->size(); ++i) {
    const PointT& p = filtered->points[i];
    const PointT& a = aligned->points[i];
    const float x = T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3), y = T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3),
                z = T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3);
    worst = std::fmax(worst, std::fmax(std::fabs(a.x - x), std::fmax(std::fabs(a.y - y), std::fabs(a.z - z))));
  }
  expect(worst < 1e-4, "aligned cloud is not the final transformation applied to the source");
  auto* b = dynamic_cast<b200reg::RegistrationBase*>(registration.get());
  expect(b != nullptr, "dynamic_cast to b200reg::RegistrationBase");
  const double fit_gpu = b ? b->fitnessScoreGPU() : -1.0;
  // the base class's own getFitnessScore (non-virtual, CPU tree) still works: the tree is built now, once, on demand
  const double fit_base = registration->getFitnessScore();
  const double fit_base2 = registration->getFitnessScore(4.0);
  std::printf("fitness(GPU)=%.9f fitness(base class, lazy tree)=%.9f worst aligned delta=%.2e lazy builds=%d\n", fit_gpu, fit_base, worst, b ? b->lazyTree().builds() : -1);
  expect(std::fabs(fit_gpu - fit_base) <= 1e-5 * fit_base, "GPU fitness differs from pcl's getFitnessScore over the same transform");
  expect(fit_base2 <= fit_base * (1.0 + 1e-12) || fit_base2 == fit_base, "getFitnessScore(max_range) inconsistent");
  expect(b && b->lazyTree().builds() == 1, "the lazy tree must be built exactly once, by the first search");
  expect(registration->getSearchMethodTarget() != nullptr, "getSearchMethodTarget() [REF :327] must keep returning a search object");
  // FAST_GICP through the same surface
  pcl::Registration<PointT, PointT>::Ptr reg2 = gicp_keep;
  reg2->setInputTarget(target);
  reg2->setInputSource(filtered);
  reg2->align(*aligned, pcl::Registration<PointT, PointT>::Matrix4());
  auto T2 = reg2->getFinalTransformation();
  std::printf("FAST_GICP: converged=%d tx=%.4f\n", (int)reg2->hasConverged(), T2(0, 3));
  expect(reg2->hasConverged() && std::fabs(T2(0, 3) + 0.2f) < 0.05f, "FAST_GICP did not recover the 0.2 m shift");
  expect(pcl::search::mock_tree_builds() == builds_before + 1, "only the one on-demand tree may have been built");
  return failures ? 1 : 0;
}

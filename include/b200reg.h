/* b200reg — C ABI of the B200-native scan-registration engine.
 *
 * Drop-in boundary for the data-parallel hot path of delta_graph_slam: the object
 * returned by select_registration_method() [REF src/hdl_graph_slam/registrations.cpp:22-124,
 * include/hdl_graph_slam/registrations.hpp:17] and the pcl::Filter used by
 * PrefilteringNodelet::downsample [REF apps/prefiltering_nodelet.cpp:249-260] and
 * ScanMatchingOdometryNodelet::downsample [REF apps/scan_matching_odometry_nodelet.cpp:155-165].
 * That boundary is a C++ virtual interface (pcl::Registration / pcl::Filter); this header
 * is the plain-C surface underneath it.  include/b200reg_pcl.hpp holds the header-only
 * pcl::Registration / pcl::Filter adapters a maintainer compiles against their PCL, and
 * INTEGRATION.md shows the factory patch.
 *
 * Conventions
 *  - every call returns 0 on success or a negative B200REG_E_* code; nothing throws or
 *    aborts (the reference's callers branch only on hasConverged() and the fitness
 *    value [REF apps/scan_matching_odometry_nodelet.cpp:222, include/hdl_graph_slam/loop_detector.hpp:149]);
 *  - clouds are arrays of pcl::PointXYZ: 16-byte records {float x, y, z, pad};
 *    `stride_bytes` is the record size (16 for PointXYZ), the pad is ignored;
 *  - 4x4 transforms are 16 floats in Eigen::Matrix4f storage order (column-major);
 *  - host pointers may be pageable; the library stages through its own pinned buffers
 *    on the handle's own CUDA stream.  Handles are independent (own stream, own device
 *    memory): the odometry object and the loop-detector object of one process may run
 *    concurrently from different threads; one handle is driven by one thread at a time.
 *  - the library never falls back to the CPU: without a usable CUDA device
 *    b200reg_create fails with B200REG_E_CUDA.
 */
#ifndef B200REG_H_
#define B200REG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REG_OK 0
#define B200REG_E_INVALID (-1)   /* bad argument */
#define B200REG_E_CUDA (-2)      /* CUDA runtime error (text in b200reg_last_error) */
#define B200REG_E_STATE (-3)     /* call order: no target / no source set */
#define B200REG_E_CAPACITY (-4)  /* caller's output buffer too small */

/* registration_method [REF src/hdl_graph_slam/registrations.cpp:26] */
enum b200reg_method {
  B200REG_METHOD_NONE = 0, /* filter-only handle (prefiltering nodelet) */
  B200REG_METHOD_NDT = 1,  /* replaces pclomp::NormalDistributionsTransform ("NDT_OMP") */
  B200REG_METHOD_GICP = 2  /* replaces fast_gicp::FastGICP ("FAST_GICP") */
};
/* reg_nn_search_method [REF src/hdl_graph_slam/registrations.cpp:103,112-118]; values follow pclomp::NeighborSearchMethod */
enum b200reg_nn_search { B200REG_KDTREE = 0, B200REG_DIRECT26 = 1, B200REG_DIRECT7 = 2, B200REG_DIRECT1 = 3 };
/* fast_gicp::RegularizationMethod */
enum b200reg_regularization { B200REG_REG_NONE = 0, B200REG_REG_MIN_EIG = 1, B200REG_REG_NORMALIZED_MIN_EIG = 2, B200REG_REG_PLANE = 3, B200REG_REG_FROBENIUS = 4 };
enum b200reg_lsq { B200REG_LSQ_GN = 0, B200REG_LSQ_LM = 1 };

/* Launch parameters of the path.  b200reg_default_config fills
 * the reference's code defaults for the chosen method. */
typedef struct b200reg_config {
  int device;                       /* CUDA device ordinal */
  int method;                       /* enum b200reg_method */
  double resolution;                /* reg_resolution              (NDT voxel size) */
  int nn_search;                    /* reg_nn_search_method */
  double transformation_epsilon;    /* reg_transformation_epsilon */
  int maximum_iterations;           /* reg_maximum_iterations */
  double step_size;                 /* NDT More-Thuente step_max (upstream 0.1, never set by the reference) */
  double outlier_ratio;             /* NDT (upstream 0.55) */
  double max_correspondence_distance; /* reg_max_correspondence_distance (GICP) */
  int correspondence_randomness;    /* reg_correspondence_randomness   (GICP k-NN size) */
  double rotation_epsilon;          /* GICP (upstream 2e-3) */
  int regularization;               /* GICP (upstream PLANE) */
  int lsq_optimizer;                /* GICP (upstream LM) */
  int num_threads;                  /* reg_num_threads: accepted and ignored (the CUDA grid replaces OpenMP) */
} b200reg_config;

typedef struct b200reg_handle b200reg_handle;

/* Result record of one registration (also the unit gathered across GPUs for loop batches). */
typedef struct b200reg_result {
  float transformation[16]; /* getFinalTransformation(), column-major */
  double fitness;           /* getFitnessScore(max_range) if requested, else 0 */
  double score;             /* NDT: getTransformationProbability(); GICP: last sum of errors */
  int32_t converged;        /* hasConverged() */
  int32_t iterations;       /* getFinalNumIteration() */
  int32_t evaluations;      /* evaluations of the reference's algorithm: computeDerivatives + computeHessian calls (NDT) / linearize + error passes (GICP) */
  int32_t passes;           /* passes over the source cloud the device actually ran (NDT: the computeHessian that closes a line search rides in its last trial pass, so passes <= evaluations) */
  int64_t hits;             /* (point, voxel) pairs visited (NDT) / correspondences evaluated (GICP) */
} b200reg_result;

void b200reg_default_config(int method, b200reg_config* out);
int b200reg_create(const b200reg_config* cfg, b200reg_handle** out);
int b200reg_destroy(b200reg_handle* h);
const char* b200reg_last_error(const b200reg_handle* h);
const char* b200reg_version(void);

/* setters used by the factory after construction [REF src/hdl_graph_slam/registrations.cpp:30-34,106-118] */
int b200reg_set_resolution(b200reg_handle* h, double resolution);
int b200reg_set_nn_search(b200reg_handle* h, int nn_search);
int b200reg_set_transformation_epsilon(b200reg_handle* h, double eps);
int b200reg_set_maximum_iterations(b200reg_handle* h, int n);
int b200reg_set_max_correspondence_distance(b200reg_handle* h, double d);
int b200reg_set_correspondence_randomness(b200reg_handle* h, int k);
/* fast_gicp setters the reference never calls (RegularizationMethod, LSQ optimiser, rotation epsilon); upstream defaults PLANE / LM / 2e-3 */
int b200reg_set_gicp_options(b200reg_handle* h, int regularization, int lsq_optimizer, double rotation_epsilon);

/* pcl::Registration::setInputTarget / setInputSource
 * [REF apps/scan_matching_odometry_nodelet.cpp:180,185,254; include/hdl_graph_slam/loop_detector.hpp:124,138].
 * An empty target is an error that leaves the previous target in place (PCL_ERROR + return). */
int b200reg_set_target(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes);
int b200reg_set_source(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes);
/* same, for clouds already resident on the handle's device (bench "value" leg, device pipelines) */
int b200reg_set_target_device(b200reg_handle* h, const float* d_xyzw, size_t n);
int b200reg_set_source_device(b200reg_handle* h, const float* d_xyzw, size_t n);
/* keyframe = filtered; registration->setInputTarget(keyframe)
 * [REF apps/scan_matching_odometry_nodelet.cpp:253-254]: the current source becomes the target
 * without leaving the device (same result as b200reg_set_target on the same cloud).  As in PCL the cloud also stays
 * the input source until b200reg_set_source replaces it (an align before that registers the cloud against itself). */
int b200reg_promote_source_to_target(b200reg_handle* h);

/* Hint: the current source will probably be promoted to target after its registration (the odometry's
 * keyframe switch, decided from the result of the align that is about to run).  The NDT grid of the
 * source is then built on a side stream WHILE that align runs, and b200reg_promote_source_to_target
 * takes it instead of building in front of the next registration.  Purely a scheduling hint: results
 * are bit-identical with or without it, a hint that does not come true only costs idle-SM time.
 * b200reg_set_side_budget: CTAs of the side build's persistent sort kernel (default 16) — together with
 * the budgets of b200reg_set_sm_budget the persistent kernels that may be resident at once must not
 * exceed the SM count. */
int b200reg_prepare_promotion(b200reg_handle* h);
int b200reg_set_side_budget(b200reg_handle* h, int n_sm);

/* pcl::Registration::align(output, guess) [REF apps/scan_matching_odometry_nodelet.cpp:218;
 * include/hdl_graph_slam/loop_detector.hpp:145].  guess == NULL means identity.
 * aligned_xyzw (optional, host, n_source * 16 bytes) receives the transformed source. */
int b200reg_align(b200reg_handle* h, const float* guess, float* aligned_xyzw);

int b200reg_has_converged(b200reg_handle* h, int* out);
int b200reg_get_final_transformation(b200reg_handle* h, float* out16);
int b200reg_get_num_iterations(b200reg_handle* h, int* out);
int b200reg_get_transformation_probability(b200reg_handle* h, double* out);
int b200reg_get_result(b200reg_handle* h, b200reg_result* out);
/* pcl::Registration::getFitnessScore(max_range) [REF include/hdl_graph_slam/loop_detector.hpp:148;
 * apps/scan_matching_odometry_nodelet.cpp:318; src/hdl_graph_slam/information_matrix_calculator.cpp:77-108] */
int b200reg_get_fitness_score(b200reg_handle* h, double max_range, double* out);
/* InformationMatrixCalculator::calc_fitness_score(cloud1 = target, cloud2 = source, relpose, max_range)
 * [REF src/hdl_graph_slam/information_matrix_calculator.cpp:77-108]: same loop as getFitnessScore with an
 * explicit transform (column-major float 4x4) instead of the last align result. */
int b200reg_calc_fitness_score(b200reg_handle* h, const float* relpose16, double max_range, double* out);
/* inlier fraction of publish_scan_matching_status [REF apps/scan_matching_odometry_nodelet.cpp:320-332] */
int b200reg_get_inlier_fraction(b200reg_handle* h, double max_correspondence_dist, double* out);

/* pcl::VoxelGrid<PointXYZ>::filter [REF apps/prefiltering_nodelet.cpp:59-63,249-260;
 * apps/scan_matching_odometry_nodelet.cpp:85-89,155-165].  Works on any handle (METHOD_NONE for a
 * filter-only object).  out has room for out_capacity points; *n_out receives the count.
 * PCL's "leaf size too small" overflow guard is reproduced: the output is then the input copy. */
int b200reg_voxelgrid_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, const float leaf[3], unsigned min_points_per_voxel,
                             int input_is_dense, float* out_xyzw, size_t out_capacity, size_t* n_out);
/* device-resident variant: d_out must hold n points; *n_out is written on the host after a stream sync */
int b200reg_voxelgrid_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, const float leaf[3], unsigned min_points_per_voxel, int input_is_dense,
                                    float* d_out_xyzw, size_t* n_out);
/* The same filter as two halves.  In the reference the prefiltering nodelet and the scan-matching
 * nodelet are separate nodelets of one manager, joined by the /filtered_points topic [REF
 * apps/prefiltering_nodelet.cpp:48,51; apps/scan_matching_odometry_nodelet.cpp:53; launch/delta_graph_slam.launch:26,46]: scan k+1 is down-sampled while scan k is being
 * matched.  _begin enqueues the whole filter (H2D of the raw scan, key / sort / centroid kernels) on
 * the handle's stream and returns; _end waits for it and returns the point count.  One filter call in
 * flight per handle (a second _begin is B200REG_E_STATE).  Page-locked `xyzw` / `out` are read by DMA
 * and written by the centroid kernel directly (mapped memory) and must stay untouched until _end;
 * pageable buffers are staged (input copied inside _begin, output copied inside _end). */
int b200reg_voxelgrid_filter_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, const float leaf[3], unsigned min_points_per_voxel,
                                   int input_is_dense, float* out_xyzw, size_t out_capacity);
int b200reg_voxelgrid_filter_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, const float leaf[3], unsigned min_points_per_voxel, int input_is_dense,
                                          float* d_out_xyzw);
/* host scan in, filtered cloud left on the device at d_out_xyzw (room for n points): what a front end that keeps
 * the prefilter and the registration in one process uses to hand the cloud over without a round trip through host memory */
int b200reg_voxelgrid_filter_host_to_device_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, const float leaf[3], unsigned min_points_per_voxel, int input_is_dense,
                                                  float* d_out_xyzw);
int b200reg_voxelgrid_filter_end(b200reg_handle* h, size_t* n_out);
/* Share of the GPU a handle's persistent (cooperative) kernels occupy: at most n_sm CTAs, one per SM.
 * Default: every SM.  A front end that overlaps the filter of scan k+1 with the registration of
 * scan k gives the filter handle a few SMs and the registration handle the rest, so both kernels
 * are resident at once (each is latency bound; neither needs the whole GPU).  Results do not depend
 * on the budget for the filter (bit-exact); NDT / GICP sums are taken in a budget-dependent order
 * (last-bit differences, as between OpenMP thread counts upstream). */
int b200reg_set_sm_budget(b200reg_handle* h, int n_sm);
/* The stages either side of VoxelGrid in PrefilteringNodelet::cloud_callback
 * [REF apps/prefiltering_nodelet.cpp:150-153: distance_filter -> downsample -> outlier_removal].
 *
 * distance_filter [REF :275-291]: with the gate on, every VoxelGrid call on this handle keeps only the
 * points with near < |p| < far (|p| in float, compared as double, as the reference does) — fused into
 * the key pipeline, so the copy_if pass and its intermediate cloud never exist.  Defaults of the
 * reference: use_distance_filter true, 1.0 / 100.0 [REF :100-102]; the launch files set 0.1 / 100.0. */
int b200reg_set_distance_filter(b200reg_handle* h, int use, double near_thresh, double far_thresh);
/* The base_link step in front of those stages [REF apps/prefiltering_nodelet.cpp:123-148]: when `base_link_frame` is set
 * the nodelet looks the sensor -> base_link transform up, zeroes its x / y translation and runs
 * pcl::transformPointCloud(*src_cloud, *transformed, transform_isometry.matrix()) — a Matrix4d, so PCL computes every
 * coordinate in double, left to right (m00*x + m01*y + m02*z + m03), and rounds to float once; w = 1; non-finite points
 * of a non-dense cloud stay as they are.  With a matrix set (16 doubles, column-major as Eigen stores it; the caller
 * zeroes the translation as the nodelet does) every VoxelGrid call — and b200reg_distance_filter, the first stage of a
 * prefilter without a down-sampler — on this handle starts with that transform, on the device, in front of the gate.
 * NULL switches it off (the reference's default: base_link_frame empty [REF :50]). */
int b200reg_set_input_transform(b200reg_handle* h, const double* matrix4x4_colmajor);
/* distance_filter as a call of its own [REF :150, :275-291] for a prefilter WITHOUT a VoxelGrid down-sampler
 * (downsample_method NONE [REF :70-75]: there is no key pass to fuse the gate into).  The reference applies the gate
 * to every scan whatever `use_distance_filter` says (it reads the flag [REF :100] and never tests it).  Order kept;
 * same conventions as the outlier filters (synchronous; host or device-resident; shares their in-flight slot). */
int b200reg_distance_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, double near_thresh, double far_thresh, float* out_xyzw, size_t out_capacity, size_t* n_out);
int b200reg_distance_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, double near_thresh, double far_thresh, float* d_out_xyzw, size_t* n_out);
/* pcl::RadiusOutlierRemoval (outlier_removal_method RADIUS) [REF :88-96,262-273]: a point stays when more
 * than min_neighbors points of the cloud (itself included) lie strictly inside `radius`; order kept.
 * Same calling conventions as the VoxelGrid filter: synchronous, device-resident, and begin / end halves
 * (one call in flight per handle, independent of a VoxelGrid call in flight on the same handle).
 * Either outlier filter (this one or the statistical one below) shares the one in-flight slot.  Both search on a
 * 0.5 m lattice over the cloud's bounding box: a cloud spanning more than 2^31 such cells (about 645 m cubed; the
 * reference's distance filter caps scans at 200 m) is refused by _end with B200REG_E_INVALID, never filtered wrongly. */
int b200reg_radius_outlier_removal(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, double radius, int min_neighbors, float* out_xyzw, size_t out_capacity,
                                   size_t* n_out);
int b200reg_radius_outlier_removal_device(b200reg_handle* h, const float* d_xyzw, size_t n, double radius, int min_neighbors, float* d_out_xyzw, size_t* n_out);
int b200reg_radius_outlier_removal_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, double radius, int min_neighbors, float* out_xyzw, size_t out_capacity);
int b200reg_radius_outlier_removal_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, double radius, int min_neighbors, float* d_out_xyzw);
int b200reg_radius_outlier_removal_end(b200reg_handle* h, size_t* n_out);
/* pcl::StatisticalOutlierRemoval (outlier_removal_method STATISTICAL, the nodelet's DEFAULT) [REF :77-87,262-273]:
 * per point the mean distance to its mean_k nearest neighbours (statistical_mean_k, default 20; 1..31 here),
 * over the cloud the mean and the sample standard deviation of those figures, and a point is dropped when
 * its figure exceeds mean + stddev_mul * stddev (statistical_stddev, default 1.0); order kept.  Non-finite
 * points and clouds of fewer than mean_k + 1 finite points behave as upstream (figure 0, not counted, kept).
 * Bit-exact against the CPU filter including its index-order double sums (csrc/sor.cuh).  Same calling
 * conventions as the radius filter; _end is the same call as b200reg_radius_outlier_removal_end. */
int b200reg_statistical_outlier_removal(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, int mean_k, double stddev_mul, float* out_xyzw, size_t out_capacity,
                                        size_t* n_out);
int b200reg_statistical_outlier_removal_device(b200reg_handle* h, const float* d_xyzw, size_t n, int mean_k, double stddev_mul, float* d_out_xyzw, size_t* n_out);
int b200reg_statistical_outlier_removal_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, int mean_k, double stddev_mul, float* out_xyzw, size_t out_capacity);
int b200reg_statistical_outlier_removal_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, int mean_k, double stddev_mul, float* d_out_xyzw);
int b200reg_statistical_outlier_removal_end(b200reg_handle* h, size_t* n_out);
/* introspection of the last statistical call, after its _end (parity tests): {mean, stddev, cut}, the number of
 * counted points, whether the index-order summation pass ran, and (dist != NULL) the n per-point figures. */
int b200reg_statistical_last_stats(b200reg_handle* h, double stats3[3], unsigned long long* valid, int* exact_pass, float* dist, size_t n);
/* filtered2D of PrefilteringNodelet::cloud_callback [REF apps/prefiltering_nodelet.cpp:155-158], one call:
 *   height_filtering (:198-214)  keep z > lidar_z (the lidar's z in base_link, :143)
 *   normal_filtering (:222-251)  pcl::NormalEstimation with setKSearch(normal_k = 10) on what is left, keep
 *                                |normalised n_z| < normal_thresh (0.2): the vertical structures
 *   flatten (:166-183)           z := 0
 * Order kept.  The height test gates the search-lattice build (no intermediate cloud), the normal is the third
 * tail of the warp-per-query k-NN kernel, the flatten happens in the order-preserving scatter.  The normal mirrors
 * PCL 1.8-1.10's float arithmetic operation by operation; its three libm calls (atan2f, cosf, sinf) are taken
 * correctly rounded (as glibc >= 2.41 returns them).  Against an older libm the last bit of a root can differ, and a
 * near-degenerate neighbourhood amplifies it: upstream's normals are not bit-reproducible across hosts either.
 * Same calling conventions and in-flight slot as the outlier filters; _end is b200reg_radius_outlier_removal_end. */
int b200reg_flat_filter(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, double lidar_z, int normal_k, double normal_thresh, float* out_xyzw, size_t out_capacity,
                        size_t* n_out);
int b200reg_flat_filter_device(b200reg_handle* h, const float* d_xyzw, size_t n, double lidar_z, int normal_k, double normal_thresh, float* d_out_xyzw, size_t* n_out);
int b200reg_flat_filter_begin(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, double lidar_z, int normal_k, double normal_thresh, float* out_xyzw, size_t out_capacity);
int b200reg_flat_filter_device_begin(b200reg_handle* h, const float* d_xyzw, size_t n, double lidar_z, int normal_k, double normal_thresh, float* d_out_xyzw);
int b200reg_flat_filter_end(b200reg_handle* h, size_t* n_out);
/* |n_z| per input point of the last flat-filter call, after its _end (NaN: below the height gate, or no normal) */
int b200reg_flat_filter_last_nz(b200reg_handle* h, float* nz, size_t n);
/* introspection of the last filter call (parity tests): per output voxel linear index and point
 * count, per input point key (0xFFFFFFFF = skipped), min_b[3] + div_b[3].  Any pointer may be NULL. */
int b200reg_voxelgrid_last_layout(b200reg_handle* h, uint32_t* voxel_id, uint32_t* count, size_t n_voxels, uint32_t* key, size_t n_points, int32_t* grid6, int* overflow);

/* ---- loop-closure batches: LoopDetector::matching [REF include/hdl_graph_slam/loop_detector.hpp:119-173] ----
 * The reference aligns every candidate keyframe against the new keyframe in a serial loop
 * (setInputTarget once :124; per candidate setInputSource :138, align :145, getFitnessScore :148).
 * Here keyframe clouds are cached on the device under the caller's keyframe id and a whole list of
 * (target, source, guess) pairs is registered in one call; pairs that share a target share its
 * NDT grid and exact-NN structure, built once.  results[i] belongs to pairs[i]:
 * transformation / converged / iterations as after align, fitness = getFitnessScore(max_range)
 * (DBL_MAX when with_fitness == 0).  A pair whose source cloud is empty comes back un-converged
 * with its guess; an unknown id or an empty target is B200REG_E_INVALID.  NDT handles only. */
typedef struct b200reg_pair {
  int64_t target_id; /* new_keyframe  (setInputTarget) */
  int64_t source_id; /* candidate     (setInputSource) */
  float guess[16];   /* column-major initial guess (transform2Dto3D of the 2-D relative pose, :139-143) */
} b200reg_pair;
/* b200reg_cloud_put: a PAGE-LOCKED keyframe cloud is read by DMA after the call has returned — keep it
 * unchanged until the next b200reg_align_batch / b200reg_calc_fitness_batch / b200reg_cloud_sync returns
 * (keyframe clouds are immutable for the whole run in the reference [REF include/hdl_graph_slam/keyframe.hpp:25-59]);
 * a pageable cloud is staged inside the call and may be reused at once. */
int b200reg_cloud_put(b200reg_handle* h, int64_t id, const float* xyzw, size_t n, size_t stride_bytes);
int b200reg_cloud_sync(b200reg_handle* h);
int b200reg_cloud_put_device(b200reg_handle* h, int64_t id, const float* d_xyzw, size_t n);
/* A cached cloud as the source / target of a plain b200reg_align: the serial loop LoopDetector::matching runs when its
 * registration object is not NDT — the launch file's loop detector uses FAST_GICP [REF launch/delta_graph_slam.launch:95;
 * include/hdl_graph_slam/loop_detector.hpp:124-156: setInputTarget(new keyframe) once, then setInputSource(candidate) +
 * align + getFitnessScore per candidate].  Equivalent to b200reg_set_source / _set_target with that keyframe's cloud (no
 * upload: a device copy); on a FAST_GICP handle the cloud's covariances (k nearest neighbours + regularisation, what
 * setInputSource / setInputTarget make fast_gicp compute lazily) are computed ONCE per cached keyframe and reused for
 * every pair it takes part in — same kernels on the same cloud, so the registration is bit-identical.  They are kept for
 * the handle's current correspondence_randomness / regularisation and recomputed when those change. */
int b200reg_set_source_cached(b200reg_handle* h, int64_t id);
int b200reg_set_target_cached(b200reg_handle* h, int64_t id);
int b200reg_cloud_drop(b200reg_handle* h, int64_t id);
int b200reg_cloud_clear(b200reg_handle* h);
int b200reg_cloud_count(b200reg_handle* h, size_t* out);
int b200reg_align_batch(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, b200reg_result* results);
/* InformationMatrixCalculator::calc_fitness_score for many keyframe pairs at once
 * [REF src/hdl_graph_slam/information_matrix_calculator.cpp:53-108; called per odometry edge and per
 * accepted loop, apps/delta_graph_slam_nodelet.cpp:572,820]: cloud1 = target_id (the exact-NN structure
 * is built once and kept with the cached cloud, where the reference builds a fresh kd-tree per call),
 * cloud2 = source_id, relpose = the pair's `guess` field (column-major float 4x4).  out[i] is the mean
 * squared nearest-neighbour distance over the points with d2 <= max_range, DBL_MAX when there is none.
 * No registration runs; works on a handle of any method. */
int b200reg_calc_fitness_batch(b200reg_handle* h, const b200reg_pair* pairs, size_t n_pairs, double max_range, double* out);
/* with timing on: CUDA-event durations of the last batch's align kernel and fitness kernels */
int b200reg_get_batch_timing(b200reg_handle* h, double* align_kernel_ms, double* fitness_ms);

/* ---- loop-closure batches over several GPUs of one box, single process ----
 * hdl_graph_slam::LoopDetector is one C++ object in one process [REF include/hdl_graph_slam/loop_detector.hpp:33-187;
 * apps/delta_graph_slam_nodelet.cpp:816-824], so the multi-GPU form of b200reg_align_batch lives behind one object too:
 * one engine handle per listed device, whole targets (a new keyframe with all of its candidates) dealt to the devices so
 * every target's structures are built on exactly one GPU, and — the path shards, there is no data-path collective — ONE
 * NCCL all-gather of the result records as the only exchange (ncclCommInitAll inside; NCCL is loaded at run time and is
 * required only for n_devices > 1).  results[i] belongs to pairs[i], exactly as b200reg_align_batch returns them.
 *
 * b200reg_batch_cloud_put registers a keyframe cloud by id and KEEPS THE POINTER: the cloud must stay valid and unchanged
 * until it is dropped or the object destroyed (keyframe clouds are immutable for the whole run in the reference
 * [REF include/hdl_graph_slam/keyframe.hpp:25-59]); it is uploaded once, to the device that first needs it.
 * `cfg->device` is ignored (the device list decides); `cfg->method` must be B200REG_METHOD_NDT. */
typedef struct b200reg_multi b200reg_multi;
int b200reg_batch_create(const b200reg_config* cfg, const int* devices, int n_devices, b200reg_multi** out);
int b200reg_batch_destroy(b200reg_multi* m);
const char* b200reg_batch_last_error(const b200reg_multi* m);
int b200reg_batch_cloud_put(b200reg_multi* m, int64_t id, const float* xyzw, size_t n, size_t stride_bytes);
int b200reg_batch_cloud_drop(b200reg_multi* m, int64_t id);
int b200reg_batch_run(b200reg_multi* m, const b200reg_pair* pairs, size_t n_pairs, int with_fitness, double fitness_max_range, b200reg_result* results);
/* after a run: device count, whether NCCL carried the gather (and its version code), pairs each device registered
 * (n_devices ints), CUDA-event duration of the all-gather on device 0.  Any pointer may be NULL. */
int b200reg_batch_get_info(b200reg_multi* m, int* n_devices, int* uses_nccl, int* nccl_version, int* pairs_per_device, double* gather_ms);

/* ---- MapCloudGenerator::generate [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49] ----
 * Every keyframe cloud transformed by its pose (poses16: n_keyframes column-major float 4x4, keyframe->pose.matrix().cast<float>())
 * and concatenated in keyframe order; with resolution > 0 reduced to the occupied voxel centres of a
 * pcl::octree::OctreePointCloud(resolution) over that cloud, in getOccupiedVoxelCenters' depth-first order (bit-identical:
 * the octree's bounding box grows with the points in insertion order, and the device path reproduces that growth); with
 * resolution <= 0 the unfiltered concatenation comes back [REF :33-34].  out needs room for every input point in the
 * worst case; B200REG_E_CAPACITY reports the needed count in *n_out.  min3_depth (optional, 4 doubles): origin of the final
 * octree box and its depth.  No keyframes is B200REG_E_INVALID (the reference warns and returns nullptr).
 * _cached takes the keyframe clouds from the handle's keyframe cache (b200reg_cloud_put) instead of host pointers. */
int b200reg_map_cloud(b200reg_handle* h, const float* const* clouds, const size_t* n_points, const float* poses16, size_t n_keyframes, double resolution, float* out_xyzw, size_t out_capacity,
                      size_t* n_out, double* min3_depth);
int b200reg_map_cloud_cached(b200reg_handle* h, const int64_t* ids, const float* poses16, size_t n_keyframes, double resolution, float* out_xyzw, size_t out_capacity, size_t* n_out,
                             double* min3_depth);

/* ---- the host side of the two front-end nodelets, in C++ above the calls of this header (csrc/b200reg_odometry.cu) ----
 * ScanMatchingOdometryNodelet::matching [REF apps/scan_matching_odometry_nodelet.cpp:173-270] as an object: the frame-to-
 * keyframe state machine (first scan -> keyframe; guess = prev_trans * delta; not converged / jump gate -> frame ignored;
 * keyframe switch on keyframe_delta_trans / _angle / _time, the aligned source promoted to target on the device) around
 * a registration handle the caller owns.  odom16 receives keyframe_pose * trans, column-major.  Parameters and defaults
 * are the nodelet's [REF :73-80]. */
typedef struct b200reg_odometry_config {
  double keyframe_delta_trans, keyframe_delta_angle, keyframe_delta_time; /* 0.25 m, 0.15 rad, 1.0 s */
  int transform_thresholding;                                             /* false */
  double max_acceptable_trans, max_acceptable_angle;                      /* 1.0 m, 1.0 rad */
} b200reg_odometry_config;
typedef struct b200reg_odometry b200reg_odometry;
void b200reg_odometry_default_config(b200reg_odometry_config* out);
int b200reg_odometry_create(b200reg_handle* registration, const b200reg_odometry_config* cfg, b200reg_odometry** out);
int b200reg_odometry_destroy(b200reg_odometry* o);
const char* b200reg_odometry_last_error(const b200reg_odometry* o);
int b200reg_odometry_reset(b200reg_odometry* o); /* forget the keyframe: the next scan starts a new sequence */
/* matching(stamp, cloud).  guess_delta16 (may be NULL = identity): the msf / robot-odometry delta multiplied onto prev_trans
 * [REF :190-214]; aligned_xyzw (may be NULL): the `aligned` cloud of registration->align(*aligned, guess) [REF :217-218]. */
int b200reg_odometry_matching(b200reg_odometry* o, double stamp, const float* xyzw, size_t n, size_t stride_bytes, const float* guess_delta16, float* aligned_xyzw, float* odom16);
int b200reg_odometry_matching_device(b200reg_odometry* o, double stamp, const float* d_xyzw, size_t n, const float* guess_delta16, float* odom16);
/* keyframes so far, whether the last scan converged / switched the keyframe, keyframe_pose, prev_trans, the last registration record; any pointer may be NULL */
int b200reg_odometry_get_state(b200reg_odometry* o, int* num_keyframes, int* last_converged, int* last_switched, float* keyframe_pose16, float* prev_trans16, b200reg_result* last_result);

/* prefiltering_nodelet -> /filtered_points -> scan_matching_odometry_nodelet as the pipeline it is in the reference
 * [REF launch/delta_graph_slam.launch:26,46; apps/prefiltering_nodelet.cpp:48,51,150-151; apps/scan_matching_odometry_nodelet.cpp:53]:
 * the distance gate + VoxelGrid of scan k+1 run on their own handle, stream and share of the SMs while scan k is matched.
 * _begin enqueues the prefilter of a scan; _step collects it, enqueues the prefilter of the NEXT scan (next_xyzw == NULL:
 * none) and matches the collected one.  Where a scan's filtered cloud goes is chosen at ITS begin:
 *   filtered_out != NULL  the reference's two nodelets: the filtered cloud lands in the caller's host cloud (the
 *                         /filtered_points message; a page-locked cloud is written by the filter kernel itself) and the
 *                         odometry side uploads it again, as setInputSource of a separate nodelet would;
 *                         the cloud must stay valid and untouched until the _step that MATCHES that scan has returned (a
 *                         page-locked cloud is read by DMA in front of the registration, without a host-side wait): hand
 *                         in at least three clouds in rotation, as a message queue would hold them;
 *   filtered_out == NULL  fused: the filtered cloud stays on the device and reaches the registration by a device copy.
 * Device-resident scans (_begin_device / _step_device / _run_device) always take the fused form. */
typedef struct b200reg_frontend_config {
  int device;
  b200reg_config registration;      /* the odometry nodelet's registration object (method, reg_* parameters) */
  b200reg_odometry_config odometry;
  double downsample_resolution;     /* prefilter VoxelGrid leaf, 0.1 [REF apps/prefiltering_nodelet.cpp:56-57] */
  int use_distance_filter;          /* 1: the gate runs on every scan, as the reference's cloud_callback does [REF :150] */
  double distance_near_thresh, distance_far_thresh; /* 1.0 / 100.0 [REF :101-102] */
  int filter_sms;                   /* SMs of the prefilter handle's persistent kernel (the registration takes the rest), 52; 0 = no split */
  int prepare_promotion;            /* scheduling only, results unchanged: 0 off; 1 build a scan's target structures on a side stream during
                                       its own registration when the motion so far says it will become the keyframe; 2 for every scan */
  int side_sms;                     /* SMs (out of filter_sms) of that side build's persistent sort kernel, 16 */
} b200reg_frontend_config;
typedef struct b200reg_frontend b200reg_frontend;
void b200reg_frontend_default_config(b200reg_frontend_config* out);
int b200reg_frontend_create(const b200reg_frontend_config* cfg, b200reg_frontend** out);
int b200reg_frontend_destroy(b200reg_frontend* fe);
const char* b200reg_frontend_last_error(const b200reg_frontend* fe);
int b200reg_frontend_reset(b200reg_frontend* fe);
b200reg_handle* b200reg_frontend_registration(b200reg_frontend* fe); /* borrowed: timing hooks, getFitnessScore, ... */
b200reg_handle* b200reg_frontend_filter(b200reg_frontend* fe);
b200reg_odometry* b200reg_frontend_odometry(b200reg_frontend* fe);
int b200reg_frontend_begin(b200reg_frontend* fe, double stamp, const float* xyzw, size_t n, size_t stride_bytes, float* filtered_out, size_t filtered_capacity);
int b200reg_frontend_begin_device(b200reg_frontend* fe, double stamp, const float* d_xyzw, size_t n);
int b200reg_frontend_step(b200reg_frontend* fe, double next_stamp, const float* next_xyzw, size_t next_n, size_t next_stride_bytes, float* next_filtered_out, size_t next_filtered_capacity,
                          size_t* n_filtered, float* aligned_out, float* odom16);
int b200reg_frontend_step_device(b200reg_frontend* fe, double next_stamp, const float* next_d_xyzw, size_t next_n, size_t* n_filtered, float* odom16);
/* a whole sequence of device-resident scans: odom16_out = frames x 16 floats; results (optional) the registration record of
 * every frame (zeros for frame 0); n_filtered (optional) the filtered cloud sizes; stamps == NULL: 0.1 s apart */
/* host wall clock per phase since the previous call, microseconds summed over the steps: out9 = {steps, wait for the
 * filter in flight, begin of the next filter, set source, align launch -> result, keyframe promotions, number of promotions,
 * whole steps, promotion hints given}; the counters restart */
int b200reg_frontend_get_timing(b200reg_frontend* fe, double* out9);
int b200reg_frontend_run_device(b200reg_frontend* fe, const float* const* d_scans, const size_t* n_points, const double* stamps, size_t frames, float* odom16_out, b200reg_result* results,
                                size_t* n_filtered, int* keyframes_out);

/* introspection of the NDT target grid (parity tests): number of occupied voxels, then per
 * voxel (ascending linear index): index, point count (-1 = rejected by the eigenvalue test),
 * mean[3], cov[9], icov[9] (row-major doubles), centroid[3] floats.  Any pointer may be NULL. */
int b200reg_ndt_num_leaves(b200reg_handle* h, size_t* out);
int b200reg_ndt_get_leaves(b200reg_handle* h, uint64_t* idx, int32_t* n, double* mean3, double* cov9, double* icov9, float* centroid3, int32_t* grid6);
/* introspection of FAST_GICP's per-point covariances (parity tests): which = 0 source, 1 target;
 * out9 receives n row-major 3x3 doubles in input order */
int b200reg_gicp_get_covariances(b200reg_handle* h, int which, double* out9, size_t n_points);
/* score / gradient / Hessian of the current source at pose p = [t, eulerXYZ] (one derivative pass) */
int b200reg_ndt_derivatives(b200reg_handle* h, const double p[6], double* score, double g[6], double H[36]);

/* Measurement hooks (bench.py).  With timing on, every align kernel launch is bracketed by CUDA
 * events on the handle's stream.  b200reg_get_counters: kernels launched by the library since it was
 * loaded (all handles), aligns timed on this handle, and their summed kernel duration in ms. */
int b200reg_set_timing(b200reg_handle* h, int on);
int b200reg_get_counters(b200reg_handle* h, long long* launches_total, long long* timed_aligns, double* align_kernel_ms);

/* developer counters of the last align (16 values): SM cycles spent by CTA 0 in {point pass, block reduce,
 * group barrier, partial sum, optimiser step}, the number of passes, the grid staging cycles, then the
 * optimiser step split into {state machine, pose trig, transform + angle tables} and the state machine
 * into {totals -> state, interval update, trial value, Newton end, 6x6 solve, Newton begin} */
int b200reg_get_profile(b200reg_handle* h, long long* out16);
/* the counters are compiled into a separate instantiation of the align kernels (they cost registers the
 * production kernel does not have to spare): off by default, single registrations only */
int b200reg_set_profile(b200reg_handle* h, int on);
/* developer trace of the last profiled NDT align: one record of 13 doubles per pass over the source
 * {nr_iterations, step_iterations, a_t, score, phi_t, d_phi_t, psi_t, d_psi_t, open_interval, interval_converged, phi_0, d_phi_0,
 *  solve_path: how H dp = -g was solved after this pass — 0 no solve, 1 closed-form 3x3 blocks, 2 pivoted elimination, 3 SVD}
 * (the More-Thuente quantities after the pass was digested), the counterpart of the oracle's NDT::trace */
int b200reg_get_trace(b200reg_handle* h, double* out, size_t cap_records, size_t* n_records);

/* A/B switch for tests (process-wide) of the voxel key / sort / segmentation pipeline: 0 = default (one cooperative
 * kernel for clouds up to 400 k points, the one-sweep chained-scan radix sort above that), 1 = always the three-launch-per-
 * digit path (<= 2 M points), 2 = always the one-sweep sort, 3 = the cooperative kernel whenever the cloud fits one tile
 * per SM.  All paths are stable sorts of the same keys: bit-identical output (tested). */
int b200reg_set_sort_path(int path);

/* developer counters of the last getFitnessScore / inlier-fraction search: {queries, queries the
 * thread-per-query near phase left open (warp-per-query far phase), queries finished by the brute-force pass} */
int b200reg_get_nn_stats(b200reg_handle* h, long long* out3);

/* raw CUDA stream of the handle (cudaStream_t) so a host can time or order work against it */
int b200reg_get_stream(b200reg_handle* h, void** out_stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REG_H_ */

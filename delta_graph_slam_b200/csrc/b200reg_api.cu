// b200reg — C ABI implementation (include/b200reg.h).  Host side is plain C++; the device work is
// the hand-written sm_100a kernels in the .cuh files next to this one.  There is no CPU
// fallback anywhere: a missing device or a CUDA error surfaces as B200REG_E_CUDA.
#include <immintrin.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include <new>
#include <string>
#include <vector>

#include "../../include/b200reg.h"
#include <map>

#include "gicp.cuh"
#include "host_copy.hpp"
#include "loop_batch.cuh"
#include "map_cloud.cuh"
#include "ndt_align.cuh"
#include "nn_grid.cuh"
#include "sor.cuh"
#include "voxel_coop.cuh"
#include "voxelgrid.cuh"

using namespace b200;

// page-locked host memory mapped into the device address space: results of single calls land here
// and the host polls the sequence flags instead of copying + synchronising the stream
struct HostMailbox {
  b200reg_result result;
  VgCounts vg;
  volatile unsigned int align_seq;
  volatile unsigned int vg_seq;
  RorCounts ror;
  volatile unsigned int ror_seq;
};

struct InputXf { double m[12]; };  // rows 0..2 of the base_link matrix, row-major (k_input_transform)

struct b200reg_handle {
  b200reg_config cfg;
  cudaStream_t stream = nullptr;
  std::string err;
  int num_sm = kNumSM;  // CTAs the cooperative kernels of this handle use (b200reg_set_sm_budget)
  int dev_sm = kNumSM;  // SMs of the device

  // clouds (device, float4)
  DevBuf<float4> src, tgt, stage_in, stage_out, aligned;
  PinnedBuf<float4> pin_in, pin_out;
  int n_src = 0, n_tgt = 0;
  bool have_src = false, have_tgt = false;
  bool src_is_tgt = false;  // the source was promoted to target and no new source has been set: pcl's input_ still is that cloud

  // VoxelGrid filter
  VoxelSort vg_sort;
  DevBuf<uint32_t> vg_id, vg_count;
  DevBuf<float4> vg_sorted;  // the scan's points in sorted (voxel, input) order: k_vg_gather -> k_vg_centroids
  bool defer_source_sync = false;  // b200reg_set_source of a page-locked cloud returns without waiting for the DMA (front end only)
  bool align_pending = false;  // b200reg_internal_align_begin .. _end
  float* align_aligned_out = nullptr;
  bool align_aligned_direct = false;
  DevBuf<float4> in_xf_buf;  // the scan in the base_link frame (b200reg_set_input_transform)
  bool in_xf_on = false;
  InputXf in_xf{};
  DevBuf<VgCounts> vg_counts;
  DevBuf<unsigned int> vg_done;
  int vg_last_n = 0, vg_last_out = 0;
  PointGate gate = kNoGate;  // distance_filter fused into the filter's key pipeline (b200reg_set_distance_filter)
  // pcl::RadiusOutlierRemoval: its own NN structure and staging, so a VoxelGrid call of the next scan
  // can be in flight on the same handle
  NnGrid nn_ror;
  DevBuf<float4> ror_in, ror_out;
  PinnedBuf<float4> ror_pin_in, ror_pin_out;
  DevBuf<unsigned char> ror_keep;
  DevBuf<uint32_t> ror_block_count;
  DevBuf<RorCounts> ror_counts;
  DevBuf<unsigned int> ror_done;
  DevBuf<float> sor_dist;       // pcl::StatisticalOutlierRemoval: mean neighbour distance per point
  DevBuf<SorStats> sor_stats;
  DevBuf<int> sor_pending;
  DevBuf<unsigned int> sor_n_pending;
  unsigned int ror_seq = 0;
  struct RorPending {
    bool active = false;
    float* host_out = nullptr;
    size_t cap = 0;
    bool zero_copy = false, device = false;
  } ror_pending;
  // filter call in flight (b200reg_voxelgrid_filter_begin .. _end)
  struct VgPending {
    bool active = false;
    float* host_out = nullptr;  // caller's output cloud (host variant), nullptr for the device variant
    size_t cap = 0;
    bool zero_copy = false;     // the kernel stores the centroids straight into host_out (page-locked)
  } vg_pending;

  // NDT
  NdtGrid grid;
  bool grid_stale = true;  // resolution changed since the grid was built
  DevBuf<NdtJob> jobs;
  DevBuf<b200reg_result> d_result;
  DevBuf<double> partials, deriv;
  DevBuf<long long> prof;
  DevBuf<double> trace;  // developer trace of the profiled align kernel (b200reg_get_trace)
  DevBuf<unsigned int> barriers;
  PinnedBuf<unsigned char> pin_small;  // results / jobs staging

  // Keyframe promotion prepared ahead (b200reg_prepare_promotion): the NDT grid of the CURRENT SOURCE is
  // built on a side stream while the registration of that source runs, so that promoting it to target
  // (the keyframe switch of the odometry) does not put a grid build in front of the next registration
  cudaStream_t side = nullptr;
  cudaEvent_t ev_src_ready = nullptr, ev_side_done = nullptr;
  NdtGrid grid_spec;
  bool spec_valid = false;      // grid_spec holds the grid of (spec_ptr, spec_n) at spec_res
  bool spec_in_flight = false;  // the side stream may still be reading the source buffer
  bool prepare_pending = false; // a hint was given: the side build is launched by the next align, behind its own kernel
  const float4* spec_ptr = nullptr;
  int spec_n = 0;
  float spec_res = 0.f;
  int spec_sms = 16;            // CTAs of the side build's cooperative sort

  // exact-NN structure on the target (fitness, inlier fraction, GICP)
  NnGrid nn;
  bool nn_stale = true;
  DevBuf<double> fit_partials;

  // FAST_GICP: the source's own NN structure, per-point covariances, correspondence scratch
  NnGrid nn_src;
  bool nn_src_stale = true;
  DevBuf<double> cov_src, cov_tgt, gicp_mahal;
  bool cov_src_ok = false, cov_tgt_ok = false;
  DevBuf<int> gicp_corr, gicp_pending;
  DevBuf<unsigned int> gicp_n_pending;
  DevBuf<GicpJob> gicp_jobs;

  // loop-closure batches: keyframe cloud cache + batch staging.  Target structures of different
  // keyframes are independent and each build is a short chain of latency-bound kernels, so they are
  // issued round-robin on kBuildLanes side streams, each with its own builders.
  static constexpr int kBuildLanes = 4;
  struct BuildLane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    NdtGrid grid;
    NnGrid nn;
  };
  BuildLane lanes[kBuildLanes];
  cudaEvent_t ev_fork = nullptr;
  std::map<long long, CachedCloud> cache;
  DevBuf<b200reg_result> batch_results;
  DevBuf<uint2> tq_runs;
  DevBuf<unsigned int> tq_next;
  PinnedBuf<uint2> pin_runs;
  DevBuf<FitJob> fit_jobs;
  DevBuf<float> batch_d2;
  DevBuf<uint2> batch_pending, batch_pending2;
  DevBuf<unsigned int> batch_n_pending;
  PinnedBuf<unsigned char> pin_batch;
  double batch_align_ms = 0.0;  // last batch: duration of the align kernel (timing on)
  double batch_fitness_ms = 0.0;

  // MapCloudGenerator::generate (map_cloud.cuh)
  DevBuf<float4> map_src, map_world, map_out;
  DevBuf<MapKeyframe> map_kf;
  DevBuf<MapEvents> map_events;
  DevBuf<unsigned long long> map_codes_a, map_codes_b, map_scalar;
  DevBuf<uint32_t> map_tile_heads;
  OneSweepScratch map_sort;

  HostMailbox* mail = nullptr;  // cudaHostAllocMapped
  unsigned int align_seq = 0, vg_seq = 0;
  size_t barriers_zeroed = 0;   // entries of `barriers` known to be zero (kernels restore them on exit)

  b200reg_result last;
  bool have_result = false;

  // optional device timing of the align kernel (bench.py roofline leg): CUDA events around every
  // launch, resolved lazily (when a pair is reused or the counters are read) so that timing never
  // puts an event synchronisation on the critical path of a call
  bool timing = false;
  bool profile = false;  // developer cycle counters (b200reg_set_profile): the PROF instantiation of the align kernels
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // aliases of the pair of the launch in flight
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<int> ev_pending;
  int ev_next = 0;
  double align_ms = 0.0;
  long long n_align = 0;

  void set_error(const std::string& s) { err = s; }
};

namespace {

const char* kVersion = "b200reg 0.1 (sm_100a)";

int set_device(b200reg_handle* h) {
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) {
    h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return B200REG_E_CUDA;
  }
  return B200REG_OK;
}

// true when p is page-locked host memory the copy engines can read or write directly
bool is_pinned_host(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();  // pageable memory is not an error for us
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

// host cloud (any stride) -> device float4 buffer on the handle's stream
int upload_cloud(b200reg_handle* h, const float* xyzw, size_t n, size_t stride_bytes, DevBuf<float4>& dst, bool defer_pinned = false) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (stride_bytes < 12 || (stride_bytes % 4) != 0) {
    h->err = "stride_bytes must be a multiple of 4 and at least 12";
    return B200REG_E_INVALID;
  }
  B200_CUDA_TRY(dst.reserve(n ? n : 1));
  if (!n) return B200REG_OK;
  if (stride_bytes == 16 && is_pinned_host(xyzw)) {
    // page-locked caller memory (cudaHostAlloc / cudaHostRegister): DMA straight from it
    B200_CUDA_TRY(cudaMemcpyAsync(dst.p, xyzw, n * 16, cudaMemcpyHostToDevice, h->stream));
    // the caller may reuse its buffer after the call returns — except for keyframe clouds put into the cache
    // (defer_pinned): those are immutable in the reference and are read by DMA until the next batch call
    if (!defer_pinned) B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return B200REG_OK;
  }
  B200_CUDA_TRY(h->pin_in.reserve(n));
  if (stride_bytes == 16) {
    host_copy(h->pin_in.p, xyzw, n * 16);  // a pageable cloud: split over the copy pool (host_copy.hpp)
  } else {
    const unsigned char* b = (const unsigned char*)xyzw;
    for (size_t i = 0; i < n; ++i) {
      const float* p = (const float*)(b + i * stride_bytes);
      h->pin_in.p[i] = make_float4(p[0], p[1], p[2], 1.0f);
    }
  }
  B200_CUDA_TRY(cudaMemcpyAsync(dst.p, h->pin_in.p, n * 16, cudaMemcpyHostToDevice, h->stream));
  // the pinned staging buffer is reused by the next upload
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return B200REG_OK;
}

// Eigen 3.3 Matrix3f::eulerAngles(0,1,2) of the rotation block; host libm floats
void euler_xyz_from_colmajor(const float* T, float out[3]) {
  auto M = [&](int r, int c) { return T[4 * c + r]; };
  const float pi = 3.14159265358979323846f;
  float r0 = atan2f(M(1, 2), M(2, 2));
  float c2 = sqrtf(M(0, 0) * M(0, 0) + M(0, 1) * M(0, 1));
  float r1;
  if (r0 > 0.f) {
    r0 -= pi;
    r1 = atan2f(-M(0, 2), -c2);
  } else {
    r1 = atan2f(-M(0, 2), c2);
  }
  float s1 = sinf(r0), c1 = cosf(r0);
  float r2 = atan2f(s1 * M(2, 0) - c1 * M(1, 0), c1 * M(1, 1) - s1 * M(2, 1));
  out[0] = -r0; out[1] = -r1; out[2] = -r2;
}

// pcl::transformPointCloud(in, out, Matrix4d) as PrefilteringNodelet::cloud_callback calls it for the base_link frame
// [REF apps/prefiltering_nodelet.cpp:137-147]: per point, in double, left to right
//   out.x = float(m00 * x + m01 * y + m02 * z + m03)   (pcl/common/impl/transforms.hpp, Scalar = double)
// and w = 1; a cloud that is not dense keeps its non-finite points as they are.  m: rows 0..2 of the matrix, row-major.
__global__ void __launch_bounds__(256) k_input_transform(const float4* __restrict__ in, int n, InputXf xf, int is_dense, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  float4 q = p;
  if (is_dense || finite3(p.x, p.y, p.z)) {
    const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
    q.x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf.m[0], x), __dmul_rn(xf.m[1], y)), __dmul_rn(xf.m[2], z)), xf.m[3]);
    q.y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf.m[4], x), __dmul_rn(xf.m[5], y)), __dmul_rn(xf.m[6], z)), xf.m[7]);
    q.z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf.m[8], x), __dmul_rn(xf.m[9], y)), __dmul_rn(xf.m[10], z)), xf.m[11]);
    q.w = 1.0f;
  }
  out[i] = q;
}

__global__ void k_transform_cloud(const float4* __restrict__ in, int n, const b200reg_result* __restrict__ res, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* T = res->transformation;  // column-major
  const float4 p = in[i];
  out[i] = make_float4(affine_row(T[0], T[4], T[8], T[12], p.x, p.y, p.z), affine_row(T[1], T[5], T[9], T[13], p.x, p.y, p.z),
                       affine_row(T[2], T[6], T[10], T[14], p.x, p.y, p.z), 1.0f);
}

int ensure_ndt_grid(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h->grid_stale && h->grid.built) return B200REG_OK;
  B200_CUDA_TRY(h->grid.build(h->stream, h->tgt.p, h->n_tgt, (float)h->cfg.resolution));
  h->grid_stale = false;
  return B200REG_OK;
}

int ensure_nn_grid(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h->nn_stale && h->nn.built) return B200REG_OK;
  B200_CUDA_TRY(h->nn.build(h->stream, h->tgt.p, h->n_tgt));
  h->nn_stale = false;
  return B200REG_OK;
}

// After b200reg_promote_source_to_target the cloud lives in the target buffer only.  pcl::Registration would
// still hold it as input_ as well (the reference never aligns in that state: the next matching() sets a new
// source first), so a call that needs the source re-creates it from the target on demand.
int materialize_source(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (h->have_src || !h->src_is_tgt || !h->have_tgt) return B200REG_OK;
  B200_CUDA_TRY(h->src.reserve(h->n_tgt ? h->n_tgt : 1));
  B200_CUDA_TRY(cudaMemcpyAsync(h->src.p, h->tgt.p, (size_t)h->n_tgt * 16, cudaMemcpyDeviceToDevice, h->stream));
  h->n_src = h->n_tgt;
  h->have_src = true;
  h->src_is_tgt = false;
  h->nn_src_stale = true;
  h->cov_src_ok = false;
  return B200REG_OK;
}

// Kernel attributes, once per device: the align kernels opt in to 192 KB of dynamic shared memory
// (the staged target grid), and EVERY kernel of the library asks for the maximum shared-memory
// carve-out, so that consecutive kernels of a frame never make the SMs re-partition L1 / shared
// memory between launches.
template <typename F>
cudaError_t prefer_shared(F* fn) { return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }

cudaError_t init_kernel_attributes(int device) {
  static std::atomic<unsigned long long> done{0};
  if (device < 64 && (done.load() >> device) & 1ull) return cudaSuccess;
  cudaError_t e;
#define B200_ATTR(expr) if ((e = (expr)) != cudaSuccess) return e
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<27, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<27, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_ndt_align<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes));
  B200_ATTR(prefer_shared(k_ndt_align<1, false>)); B200_ATTR(prefer_shared(k_ndt_align<7, false>)); B200_ATTR(prefer_shared(k_ndt_align<27, false>)); B200_ATTR(prefer_shared(k_ndt_align<0, false>));
  B200_ATTR(prefer_shared(k_ndt_align<1, true>)); B200_ATTR(prefer_shared(k_ndt_align<7, true>)); B200_ATTR(prefer_shared(k_ndt_align<27, true>)); B200_ATTR(prefer_shared(k_ndt_align<0, true>));
  B200_ATTR(prefer_shared(k_voxel_sort_coop<2>)); B200_ATTR(prefer_shared(k_voxel_sort_coop<4>)); B200_ATTR(prefer_shared(k_voxel_sort_coop<8>));
  B200_ATTR(prefer_shared(k_voxel_sort_coop<16>)); B200_ATTR(prefer_shared(k_voxel_sort_coop<32>));
  B200_ATTR(prefer_shared(k_os_histogram<uint32_t>)); B200_ATTR(prefer_shared(k_os_histogram<unsigned long long>)); B200_ATTR(prefer_shared((k_os_pass<unsigned long long, 8>))); B200_ATTR(prefer_shared((k_os_pass<uint32_t, 8>))); B200_ATTR(prefer_shared(k_vg_centroids)); B200_ATTR(prefer_shared(k_vg_gather)); B200_ATTR(prefer_shared(k_vg_compact)); B200_ATTR(prefer_shared(k_gate_copy)); B200_ATTR(prefer_shared(k_ror_flags)); B200_ATTR(prefer_shared(k_ror_scatter)); B200_ATTR(prefer_shared(k_nn_occ_clear)); B200_ATTR(prefer_shared(k_transform_cloud)); B200_ATTR(prefer_shared(k_input_transform));
  B200_ATTR(prefer_shared(k_nn_reorder)); B200_ATTR(prefer_shared(k_nn_insert)); B200_ATTR(prefer_shared(k_nn_search)); B200_ATTR(prefer_shared(k_nn_far));
  B200_ATTR(prefer_shared(k_nn_bruteforce)); B200_ATTR(prefer_shared(k_fitness_partial));
  B200_ATTR(prefer_shared(k_nn_search_batch)); B200_ATTR(prefer_shared(k_nn_far_batch)); B200_ATTR(prefer_shared(k_nn_bruteforce_batch)); B200_ATTR(prefer_shared(k_fitness_batch));
  B200_ATTR(prefer_shared(k_gicp_knn<kKnnCovariance>)); B200_ATTR(prefer_shared(k_gicp_knn_brute<kKnnCovariance>)); B200_ATTR(prefer_shared(k_gicp_knn<kKnnMeanDistance>));
  B200_ATTR(prefer_shared(k_gicp_knn_brute<kKnnMeanDistance>)); B200_ATTR(prefer_shared(k_sor_threshold)); B200_ATTR(prefer_shared(k_sor_flags));
  B200_ATTR(prefer_shared(k_gicp_knn<kKnnNormalNz>)); B200_ATTR(prefer_shared(k_gicp_knn_brute<kKnnNormalNz>)); B200_ATTR(prefer_shared(k_nz_flags)); B200_ATTR(prefer_shared(k_gate_flags)); B200_ATTR(prefer_shared(k_gicp_regularize));
  // k_gicp_align needs 17.5 KB of shared memory and keeps its 29 double accumulators + 3x3 temporaries in a
  // 1.4 KB per-thread stack frame (128-register cap at 512 threads): it wants the L1, not the carve-out
  B200_ATTR(cudaFuncSetAttribute((const void*)k_gicp_align<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 16));
  B200_ATTR(cudaFuncSetAttribute((const void*)k_gicp_align<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 16));
  B200_ATTR(ndt_leaf_prefer_shared());
#undef B200_ATTR
  if (device < 64) done.fetch_or(1ull << device);
  return cudaSuccess;
}

// ---- lazy event timing ---------------------------------------------------------------------------
int drain_events(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  for (int idx : h->ev_pending) {
    B200_CUDA_TRY(cudaEventSynchronize(h->ev_pool[idx].second));
    float ms = 0.f;
    B200_CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_pool[idx].first, h->ev_pool[idx].second));
    h->align_ms += (double)ms;
    h->n_align += 1;
  }
  h->ev_pending.clear();
  return B200REG_OK;
}

// next event pair for a launch (timing on); its elapsed time is added to the counters later
int begin_timed_launch(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h->timing) return B200REG_OK;
  if (h->ev_pool.empty()) {
    h->ev_pool.resize(64);
    for (auto& pr : h->ev_pool) {
      B200_CUDA_TRY(cudaEventCreate(&pr.first));
      B200_CUDA_TRY(cudaEventCreate(&pr.second));
    }
  }
  if (h->ev_pending.size() >= h->ev_pool.size()) {
    int rc = drain_events(h);
    if (rc) return rc;
  }
  const int idx = h->ev_next;
  h->ev_next = (h->ev_next + 1) % (int)h->ev_pool.size();
  h->ev0 = h->ev_pool[idx].first;
  h->ev1 = h->ev_pool[idx].second;
  h->ev_pending.push_back(idx);
  B200_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  return B200REG_OK;
}
int end_timed_launch(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (h->timing) B200_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  return B200REG_OK;
}

// barrier lines / job counters: zero when (re)allocated, afterwards every kernel restores them on exit
int ensure_barriers(b200reg_handle* h, size_t entries) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  const unsigned int* before = h->barriers.p;
  B200_CUDA_TRY(h->barriers.reserve(entries));
  if (h->barriers.p != before || h->barriers_zeroed < entries) {
    B200_CUDA_TRY(cudaMemsetAsync(h->barriers.p, 0, h->barriers.cap * sizeof(unsigned int), h->stream));
    h->barriers_zeroed = h->barriers.cap;
  }
  return B200REG_OK;
}

// wait for the kernel of a single call to publish its result in the mailbox
int wait_mail(b200reg_handle* h, volatile unsigned int* flag, unsigned int want) {
  unsigned long long spins = 0;
  while (*flag != want) {
    _mm_pause();
    if ((++spins & 0x3FFFull) == 0) {
      const cudaError_t q = cudaStreamQuery(h->stream);
      if (q == cudaErrorNotReady) continue;
      if (q != cudaSuccess) {
        h->err = std::string("kernel failed: ") + cudaGetErrorString(q);
        return B200REG_E_CUDA;
      }
      if (*flag != want) {  // stream drained without the flag: should not happen
        h->err = "kernel finished without publishing its result";
        return B200REG_E_CUDA;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return B200REG_OK;
}

template <int MODE>
cudaError_t launch_ndt(b200reg_handle* h, int n_jobs, int ctas_per_group, int n_groups, const NdtJob* single, NdtTargetQueue tq) {
  NdtParams prm;
  prm.search = h->cfg.nn_search;
  prm.resolution = h->cfg.resolution;
  prm.step_size = h->cfg.step_size;
  prm.outlier_ratio = h->cfg.outlier_ratio;
  prm.trans_eps = h->cfg.transformation_epsilon;
  prm.max_iterations = h->cfg.maximum_iterations;
  static const NdtJob kNoJob = {};
  const NdtJob* jobs = single ? nullptr : h->jobs.p;  // one registration: the job rides in the kernel parameters
  const NdtJob& sj = single ? *single : kNoJob;
  double* partials = h->partials.p;
  unsigned int* barriers = h->barriers.p;
  unsigned int* queue = h->barriers.p + (size_t)n_groups * 32;  // the job ticket counter sits behind the groups' barrier lines
  void* args[] = {(void*)&jobs, (void*)&n_jobs, (void*)&ctas_per_group, (void*)&tq, (void*)&prm, (void*)&partials, (void*)&barriers, (void*)&queue, (void*)&sj};
  const void* fn = (h->profile && single) ? (const void*)k_ndt_align<MODE, true> : (const void*)k_ndt_align<MODE, false>;
  return cudaLaunchCooperativeKernel(fn, dim3(ctas_per_group * n_groups), dim3(kAlignThreads), args, kStageBytes, h->stream);
}

cudaError_t launch_ndt_mode(b200reg_handle* h, int n_jobs, int G, int n_groups, const NdtJob* single = nullptr, NdtTargetQueue tq = NdtTargetQueue{nullptr, nullptr, 0}) {
  switch (h->cfg.nn_search) {
    case B200REG_DIRECT1: return launch_ndt<1>(h, n_jobs, G, n_groups, single, tq);
    case B200REG_DIRECT26: return launch_ndt<27>(h, n_jobs, G, n_groups, single, tq);
    case B200REG_KDTREE: return launch_ndt<0>(h, n_jobs, G, n_groups, single, tq);
    default: return launch_ndt<7>(h, n_jobs, G, n_groups, single, tq);
  }
}

// one NDT job on the whole GPU; result lands in h->d_result[0]
int run_ndt_single(b200reg_handle* h, const float* guess_colmajor, const double* p_eval) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  int rc = ensure_ndt_grid(h);
  if (rc) return rc;
  const int G = h->num_sm;
  B200_CUDA_TRY(h->d_result.reserve(1));
  B200_CUDA_TRY(h->deriv.reserve(64));
  B200_CUDA_TRY(h->prof.reserve(16));
  B200_CUDA_TRY(h->partials.reserve((size_t)2 * G * kAccStride));
  if ((rc = ensure_barriers(h, 64))) return rc;
  NdtJob jobv;
  NdtJob* job = &jobv;
  memset(job, 0, sizeof(NdtJob));
  job->src = h->src.p;
  job->n_src = h->n_src;
  job->grid = h->grid.view();
  job->result = h->d_result.p;
  job->result_host = &h->mail->result;
  job->done_flag = const_cast<unsigned int*>(&h->mail->align_seq);
  job->done_seq = ++h->align_seq;
  job->deriv_out = h->deriv.p;
  job->prof = h->prof.p;
  if (h->profile) {
    B200_CUDA_TRY(h->trace.reserve((size_t)(kTraceCap + 1) * kTraceDoubles));
    B200_CUDA_TRY(cudaMemsetAsync(h->trace.p, 0, (size_t)(kTraceCap + 1) * kTraceDoubles * sizeof(double), h->stream));
    job->trace = h->trace.p;
  }
  if (p_eval) {
    job->eval_only = 1;
    for (int i = 0; i < 6; ++i) job->p0[i] = p_eval[i];
  } else {
    float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const float* g = guess_colmajor ? guess_colmajor : I;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) job->guess[4 * r + c] = g[4 * c + r];
    float eul[3];
    euler_xyz_from_colmajor(g, eul);
    job->p0[0] = g[12]; job->p0[1] = g[13]; job->p0[2] = g[14];
    job->p0[3] = eul[0]; job->p0[4] = eul[1]; job->p0[5] = eul[2];
  }
  if ((rc = begin_timed_launch(h))) return rc;
  B200_CUDA_TRY(launch_ndt_mode(h, 1, G, 1, job));
  launch_counter() += 1;
  if ((rc = end_timed_launch(h))) return rc;
  return B200REG_OK;
}

// ---- FAST_GICP ---------------------------------------------------------------------------------
cudaError_t launch_gicp_covariances(b200reg_handle* h, const NnView& nnv, const float4* pts, int n, double* covs) {
  if (n <= 0) return cudaSuccess;
  const int k = h->cfg.correspondence_randomness, reg = h->cfg.regularization;
  cudaError_t e;
  if ((e = h->gicp_pending.reserve((size_t)n)) != cudaSuccess) return e;
  if ((e = h->gicp_n_pending.reserve(1)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(h->gicp_n_pending.p, 0, sizeof(unsigned int), h->stream)) != cudaSuccess) return e;
  const int blocks = (n + 7) / 8;  // one warp per query, 8 warps per CTA
  launch_counter() += 3;
  k_gicp_knn<kKnnCovariance><<<blocks, 256, 0, h->stream>>>(nnv, pts, n, k, covs, h->gicp_pending.p, h->gicp_n_pending.p, nullptr);
  k_gicp_knn_brute<kKnnCovariance><<<kNumSM * 2, kBruteWarps * 32, 0, h->stream>>>(nnv, pts, k, covs, h->gicp_pending.p, h->gicp_n_pending.p, nullptr);
  k_gicp_regularize<<<(n + 127) / 128, 128, 0, h->stream>>>(n, reg, covs);
  return cudaGetLastError();
}

// source_covs_ / target_covs_ are computed lazily at the first align after a cloud changed (A.5)
int ensure_gicp_structures(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  int rc = ensure_nn_grid(h);
  if (rc) return rc;
  if (h->cfg.correspondence_randomness > 32) { h->err = "reg_correspondence_randomness above 32 is not supported by the device k-NN (one neighbour per warp lane)"; return B200REG_E_INVALID; }
  if (!h->cov_tgt_ok) {
    B200_CUDA_TRY(h->cov_tgt.reserve((size_t)h->n_tgt * 6));
    B200_CUDA_TRY(launch_gicp_covariances(h, h->nn.view(), h->tgt.p, h->n_tgt, h->cov_tgt.p));
    h->cov_tgt_ok = true;
  }
  if (!h->cov_src_ok) {
    if (h->nn_src_stale || !h->nn_src.built) {
      B200_CUDA_TRY(h->nn_src.build(h->stream, h->src.p, h->n_src));
      h->nn_src_stale = false;
    }
    B200_CUDA_TRY(h->cov_src.reserve((size_t)(h->n_src > 0 ? h->n_src : 1) * 6));
    B200_CUDA_TRY(launch_gicp_covariances(h, h->nn_src.view(), h->src.p, h->n_src, h->cov_src.p));
    h->cov_src_ok = true;
  }
  return B200REG_OK;
}

int run_gicp_single(b200reg_handle* h, const float* guess_colmajor) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  int rc = ensure_gicp_structures(h);
  if (rc) return rc;
  const int G = h->num_sm;
  const size_t ns = (size_t)(h->n_src > 0 ? h->n_src : 1);
  B200_CUDA_TRY(h->d_result.reserve(1));
  B200_CUDA_TRY(h->gicp_corr.reserve(ns));
  B200_CUDA_TRY(h->gicp_mahal.reserve(ns * 6));
  B200_CUDA_TRY(h->partials.reserve((size_t)2 * G * kGicpStride));
  if ((rc = ensure_barriers(h, 64))) return rc;
  GicpJob jobv;
  GicpJob* job = &jobv;
  memset(job, 0, sizeof(GicpJob));
  job->result_host = &h->mail->result;
  job->done_flag = const_cast<unsigned int*>(&h->mail->align_seq);
  job->done_seq = ++h->align_seq;
  job->src = h->src.p;
  job->n_src = h->n_src;
  job->cov_src = h->cov_src.p;
  job->tgt = h->nn.view();
  job->tgt_pts = h->tgt.p;
  job->cov_tgt = h->cov_tgt.p;
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(job->guess, guess_colmajor ? guess_colmajor : I, 64);
  B200_CUDA_TRY(h->prof.reserve(16));
  job->prof = h->prof.p;
  job->corr = h->gicp_corr.p;
  job->mahal = h->gicp_mahal.p;
  job->result = h->d_result.p;
  GicpParams prm;
  const double d = h->cfg.max_correspondence_distance;
  prm.corr_dist2 = d * d;
  prm.search_d2 = prm.corr_dist2 >= 3.0e38 ? 3.402823466e+38f : (float)prm.corr_dist2 * 1.0001f + 1e-6f;
  const double rings = ceil(d / (double)kNnCell) + 1.0;
  prm.far_ring = rings > (double)kFarRing ? kFarRing : (int)rings;
  {
    // lanes per source point in the near search of a linearize pass: as many (4, 2, 1) as still let ONE sweep of the grid's
    // warps cover the cloud.  More lanes shorten a query's chain of dependent loads, but a second sweep costs more than
    // that gains: on the 48 k-point scans of the odometry leg four lanes (four sweeps on 108 SMs) measured 1456
    // registrations/s, two lanes 1470, one lane 1535.  B200_GICP_LANES overrides (A/B runs).
    static const int forced = getenv("B200_GICP_LANES") ? atoi(getenv("B200_GICP_LANES")) : 0;
    const long long warps = (long long)G * kGicpWarps;
    int ql = 1;
    if ((long long)h->n_src * 4 <= warps * 32) ql = 4;
    else if ((long long)h->n_src * 2 <= warps * 32) ql = 2;
    prm.query_lanes = (forced == 1 || forced == 2 || forced == 4) ? forced : ql;
  }
  prm.trans_eps = h->cfg.transformation_epsilon;
  prm.rot_eps = h->cfg.rotation_epsilon;
  prm.max_iterations = h->cfg.maximum_iterations;
  prm.lsq = h->cfg.lsq_optimizer;
  prm.lm_max_iterations = 10;
  prm.lm_init_lambda_factor = 1e-9;
  int g = G;
  double* partials = h->partials.p;
  unsigned int* barrier = h->barriers.p;
  void* args[] = {(void*)job, (void*)&g, (void*)&prm, (void*)&partials, (void*)&barrier};
  if ((rc = begin_timed_launch(h))) return rc;
  B200_CUDA_TRY(cudaLaunchCooperativeKernel(h->profile ? (const void*)k_gicp_align<true> : (const void*)k_gicp_align<false>, dim3(G), dim3(kGicpThreads), args, 0, h->stream));
  launch_counter() += 1;
  if ((rc = end_timed_launch(h))) return rc;
  return B200REG_OK;
}

// result of the single call in flight: polled from the mailbox the kernel writes through the mapped
// host pointer.  `synced` = the caller has already synchronised the stream (aligned-cloud path).
int fetch_result(b200reg_handle* h, bool synced = false) {
  if (!synced) {
    int rc = wait_mail(h, &h->mail->align_seq, h->align_seq);
    if (rc) return rc;
  } else if (h->mail->align_seq != h->align_seq) {
    h->err = "kernel finished without publishing its result";
    return B200REG_E_CUDA;
  }
  memcpy(&h->last, const_cast<b200reg_result*>(&h->mail->result), sizeof(b200reg_result));
  h->have_result = true;
  return B200REG_OK;
}

// Morton codes of the map cloud: digits that hold code bits, then ONE pass over the top digit (zero for every valid code,
// 0xFF for the all-ones code of a non-finite point), so invalid codes end up behind the valid ones without sorting the
// empty digits in between.  The one-sweep sort takes "bits to sort" from device memory; here each launch names its digit.
cudaError_t map_sort_codes(b200reg_handle* h, int n, int code_passes) {
  cudaError_t e;
  OneSweepScratch& sc = h->map_sort;
  const int max_passes = 8;
  constexpr int ITEMS = os_items<unsigned long long>();
  const int n_tiles = sc.tiles(n, ITEMS);
  const size_t words = sc.words(n, max_passes, ITEMS);
  if ((e = sc.buf.reserve(words)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(sc.buf.p, 0, words * sizeof(uint32_t), h->stream)) != cudaSuccess) return e;
  uint32_t* hist = sc.buf.p;
  unsigned int* tickets = sc.buf.p + (size_t)max_passes * kOsRadix;
  uint32_t* status = sc.buf.p + (size_t)max_passes * (kOsRadix + 32);
  const uint32_t* d_nbits = reinterpret_cast<const uint32_t*>(h->map_scalar.p + 1);  // holds 64: every digit "significant"
  int hb = (n + kOsThreads * 8 - 1) / (kOsThreads * 8);
  if (hb > kNumSM * 4) hb = kNumSM * 4;
  launch_counter() += 2 + code_passes;
  k_os_histogram<unsigned long long><<<hb, kOsThreads, 0, h->stream>>>(h->map_codes_a.p, n, d_nbits, max_passes, hist, kNoGather);
  unsigned long long* a = h->map_codes_a.p;
  unsigned long long* b = h->map_codes_b.p;
  int launch = 0;
  for (int p = 0; p < code_passes + 1; ++p) {
    const int digit = p < code_passes ? p : 7;
    k_os_pass<unsigned long long, ITEMS><<<n_tiles, kOsThreads, 0, h->stream>>>((launch & 1) ? b : a, nullptr, (launch & 1) ? a : b, nullptr, n, digit, d_nbits, hist,
                                                                          status + (size_t)digit * n_tiles * kOsRadix, tickets + digit * 32, kNoGather);
    ++launch;
  }
  return cudaGetLastError();
}

}  // namespace

extern "C" {

const char* b200reg_version(void) { return kVersion; }

void b200reg_default_config(int method, b200reg_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->device = 0;
  c->method = method;
  // defaults of select_registration_method [REF src/hdl_graph_slam/registrations.cpp:26-119]
  c->resolution = 0.5;                  // reg_resolution (NDT branch)
  c->nn_search = B200REG_DIRECT7;       // reg_nn_search_method
  c->transformation_epsilon = 0.01;     // reg_transformation_epsilon
  c->maximum_iterations = 64;           // reg_maximum_iterations
  c->step_size = 0.1;                   // pclomp default, never set by the reference
  c->outlier_ratio = 0.55;              // pclomp default
  c->max_correspondence_distance = 2.5; // reg_max_correspondence_distance
  c->correspondence_randomness = 20;    // reg_correspondence_randomness
  c->rotation_epsilon = 2e-3;           // fast_gicp default
  c->regularization = B200REG_REG_PLANE;
  c->lsq_optimizer = B200REG_LSQ_LM;
  c->num_threads = 0;
}

int b200reg_create(const b200reg_config* cfg, b200reg_handle** out) {
  if (!cfg || !out) return B200REG_E_INVALID;
  *out = nullptr;
  if (cfg->method < B200REG_METHOD_NONE || cfg->method > B200REG_METHOD_GICP) return B200REG_E_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || cfg->device < 0 || cfg->device >= count) return B200REG_E_CUDA;
  b200reg_handle* h = new (std::nothrow) b200reg_handle();
  if (!h) return B200REG_E_INVALID;
  h->cfg = *cfg;
  if (cudaSetDevice(cfg->device) != cudaSuccess) { delete h; return B200REG_E_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete h; return B200REG_E_CUDA; }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
  if (prop.major < 10 || !coop) { delete h; return B200REG_E_CUDA; }  // sm_100a only
  h->num_sm = h->dev_sm = prop.multiProcessorCount < kNumSM ? prop.multiProcessorCount : kNumSM;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return B200REG_E_CUDA; }
  if (init_kernel_attributes(cfg->device) != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return B200REG_E_CUDA; }
  if (cudaHostAlloc((void**)&h->mail, sizeof(HostMailbox), cudaHostAllocMapped) != cudaSuccess) { cudaStreamDestroy(h->stream); delete h; return B200REG_E_CUDA; }
  memset((void*)h->mail, 0, sizeof(HostMailbox));
  *out = h;
  return B200REG_OK;
}

int b200reg_destroy(b200reg_handle* h) {
  if (!h) return B200REG_E_INVALID;
  cudaSetDevice(h->cfg.device);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  for (auto& pr : h->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (h->mail) cudaFreeHost((void*)h->mail);
  h->src.release(); h->tgt.release(); h->stage_in.release(); h->stage_out.release(); h->aligned.release();
  h->pin_in.release(); h->pin_out.release(); h->vg_sort.release(); h->vg_id.release(); h->vg_count.release(); h->vg_sorted.release(); h->in_xf_buf.release(); h->vg_counts.release(); h->vg_done.release();
  h->grid.release(); h->jobs.release(); h->d_result.release(); h->partials.release(); h->deriv.release(); h->barriers.release(); h->pin_small.release(); h->prof.release(); h->trace.release();
  h->nn.release(); h->fit_partials.release();
  h->nn_src.release(); h->cov_src.release(); h->cov_tgt.release(); h->gicp_mahal.release(); h->nn_ror.release(); h->ror_in.release(); h->ror_out.release(); h->ror_pin_in.release(); h->ror_pin_out.release(); h->ror_keep.release(); h->ror_block_count.release(); h->ror_counts.release(); h->ror_done.release(); h->sor_dist.release(); h->sor_stats.release(); h->sor_pending.release(); h->sor_n_pending.release();
  h->gicp_corr.release(); h->gicp_pending.release(); h->gicp_n_pending.release(); h->gicp_jobs.release();
  for (auto& kv : h->cache) kv.second.release();
  h->cache.clear();
  for (auto& ln : h->lanes) {
    if (ln.st) { cudaStreamSynchronize(ln.st); cudaStreamDestroy(ln.st); }
    if (ln.done) cudaEventDestroy(ln.done);
    ln.grid.release();
    ln.nn.release();
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_src_ready) cudaEventDestroy(h->ev_src_ready);
  if (h->ev_side_done) cudaEventDestroy(h->ev_side_done);
  h->grid_spec.release();
  h->batch_results.release(); h->tq_runs.release(); h->tq_next.release(); h->pin_runs.release(); h->fit_jobs.release(); h->batch_d2.release(); h->batch_pending.release(); h->batch_pending2.release(); h->batch_n_pending.release(); h->pin_batch.release();
  h->map_src.release(); h->map_world.release(); h->map_out.release(); h->map_kf.release(); h->map_events.release(); h->map_codes_a.release(); h->map_codes_b.release(); h->map_scalar.release(); h->map_tile_heads.release(); h->map_sort.buf.release();
  delete h;
  return B200REG_OK;
}

const char* b200reg_last_error(const b200reg_handle* h) { return h ? h->err.c_str() : "null handle"; }

int b200reg_set_resolution(b200reg_handle* h, double r) {
  if (!h || !(r > 0)) return B200REG_E_INVALID;
  if ((float)r != (float)h->cfg.resolution) h->grid_stale = true;  // pclomp::setResolution re-inits the grid
  h->cfg.resolution = r;
  return B200REG_OK;
}
int b200reg_set_nn_search(b200reg_handle* h, int m) {
  if (!h || m < B200REG_KDTREE || m > B200REG_DIRECT1) return B200REG_E_INVALID;
  h->cfg.nn_search = m;
  return B200REG_OK;
}
int b200reg_set_transformation_epsilon(b200reg_handle* h, double e) { if (!h) return B200REG_E_INVALID; h->cfg.transformation_epsilon = e; return B200REG_OK; }
int b200reg_set_maximum_iterations(b200reg_handle* h, int n) { if (!h) return B200REG_E_INVALID; h->cfg.maximum_iterations = n; return B200REG_OK; }
int b200reg_set_max_correspondence_distance(b200reg_handle* h, double d) { if (!h) return B200REG_E_INVALID; h->cfg.max_correspondence_distance = d; return B200REG_OK; }
int b200reg_set_correspondence_randomness(b200reg_handle* h, int k) {
  if (!h || k < 1) return B200REG_E_INVALID;
  if (k != h->cfg.correspondence_randomness) h->cov_src_ok = h->cov_tgt_ok = false;
  h->cfg.correspondence_randomness = k;
  return B200REG_OK;
}
int b200reg_set_gicp_options(b200reg_handle* h, int regularization, int lsq_optimizer, double rotation_epsilon) {
  if (!h || regularization < B200REG_REG_NONE || regularization > B200REG_REG_FROBENIUS || lsq_optimizer < B200REG_LSQ_GN || lsq_optimizer > B200REG_LSQ_LM || !(rotation_epsilon > 0))
    return B200REG_E_INVALID;
  if (regularization != h->cfg.regularization) h->cov_src_ok = h->cov_tgt_ok = false;
  h->cfg.regularization = regularization;
  h->cfg.lsq_optimizer = lsq_optimizer;
  h->cfg.rotation_epsilon = rotation_epsilon;
  return B200REG_OK;
}

int b200reg_set_target(b200reg_handle* h, const float* xyzw, size_t n, size_t stride) {
  if (!h) return B200REG_E_INVALID;
  if (!n || !xyzw) { h->err = "Invalid or empty point cloud dataset given!"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = materialize_source(h))) return rc;  // the promoted cloud is still pcl's input_
  if ((rc = upload_cloud(h, xyzw, n, stride, h->tgt))) return rc;
  h->n_tgt = (int)n;
  h->have_tgt = true;
  h->grid_stale = true;
  h->nn_stale = true;
  h->cov_tgt_ok = false;
  if (h->cfg.method == B200REG_METHOD_NDT) return ensure_ndt_grid(h);  // ndt->setInputTarget builds the grid eagerly
  return B200REG_OK;
}

// the source buffer is about to change: a prepared promotion no longer applies, and a side build that
// may still be reading the buffer has to finish first (stream order, no host wait)
static int retire_spec(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (h->spec_in_flight) {
    B200_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_side_done, 0));
    h->spec_in_flight = false;
  }
  h->spec_valid = false;
  h->prepare_pending = false;
  return B200REG_OK;
}

int b200reg_set_source(b200reg_handle* h, const float* xyzw, size_t n, size_t stride) {
  if (!h || (n && !xyzw)) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = retire_spec(h))) return rc;
  // (defer_source_sync: the front end's filtered clouds stay valid until two scans later; the copy is ordered in front of the
  // registration on the stream, and the host goes on to launch it instead of waiting for the DMA)
  if ((rc = upload_cloud(h, xyzw, n, stride, h->src, /*defer_pinned=*/h->defer_source_sync))) return rc;
  h->n_src = (int)n;
  h->have_src = true;
  h->src_is_tgt = false;
  h->nn_src_stale = true;
  h->cov_src_ok = false;
  return B200REG_OK;
}

int b200reg_set_target_device(b200reg_handle* h, const float* d_xyzw, size_t n) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  if (!n || !d_xyzw) { h->err = "Invalid or empty point cloud dataset given!"; return B200REG_E_INVALID; }
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = materialize_source(h))) return rc;
  B200_CUDA_TRY(h->tgt.reserve(n));
  B200_CUDA_TRY(cudaMemcpyAsync(h->tgt.p, d_xyzw, n * 16, cudaMemcpyDeviceToDevice, h->stream));
  h->n_tgt = (int)n;
  h->have_tgt = true;
  h->grid_stale = true;
  h->nn_stale = true;
  h->cov_tgt_ok = false;
  if (h->cfg.method == B200REG_METHOD_NDT) return ensure_ndt_grid(h);
  return B200REG_OK;
}

int b200reg_set_source_device(b200reg_handle* h, const float* d_xyzw, size_t n) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || (n && !d_xyzw)) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = retire_spec(h))) return rc;
  B200_CUDA_TRY(h->src.reserve(n ? n : 1));
  if (n) B200_CUDA_TRY(cudaMemcpyAsync(h->src.p, d_xyzw, n * 16, cudaMemcpyDeviceToDevice, h->stream));
  h->n_src = (int)n;
  h->have_src = true;
  h->src_is_tgt = false;
  h->nn_src_stale = true;
  h->cov_src_ok = false;
  return B200REG_OK;
}

int b200reg_promote_source_to_target(b200reg_handle* h) {
  if (!h) return B200REG_E_INVALID;
  if (!h->have_src || h->n_src == 0) { h->err = "no source to promote"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  // swap the buffers: the old target storage becomes the (now unset) source storage
  std::swap(h->src, h->tgt);
  h->n_tgt = h->n_src;
  h->have_tgt = true;
  h->have_src = false;
  h->src_is_tgt = true;
  h->n_src = 0;
  h->grid_stale = true;
  // GICP: the cloud keeps its NN structure and covariances when it changes role (same cloud, same values)
  std::swap(h->nn, h->nn_src);
  h->nn_stale = h->nn_src_stale || !h->nn.built;
  h->nn_src_stale = true;
  std::swap(h->cov_tgt, h->cov_src);
  h->cov_tgt_ok = h->cov_src_ok;
  h->cov_src_ok = false;
  if (h->cfg.method == B200REG_METHOD_NDT) {
    if (h->spec_valid && h->spec_ptr == h->tgt.p && h->spec_n == h->n_tgt && h->spec_res == (float)h->cfg.resolution) {
      // the grid of this cloud was built ahead on the side stream: take it (same kernels, same bits)
      auto set_error = [&](const std::string& s) { h->err = s; };
      B200_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_side_done, 0));
      std::swap(h->grid, h->grid_spec);
      h->grid_stale = false;
      h->spec_valid = false;
      h->spec_in_flight = false;
      return B200REG_OK;
    }
    if ((rc = retire_spec(h))) return rc;
    return ensure_ndt_grid(h);
  }
  return B200REG_OK;
}

// The side build itself: enqueued on the side stream behind ev_src_ready (recorded when the hint was given, i.e. after
// the source cloud's copy and BEFORE the align kernel was launched on the main stream).
static int launch_side_build(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  h->prepare_pending = false;
  B200_CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_src_ready, 0));
  h->grid_spec.sort.max_ctas = h->spec_sms;
  B200_CUDA_TRY(h->grid_spec.build(h->side, h->src.p, h->n_src, (float)h->cfg.resolution));
  B200_CUDA_TRY(cudaEventRecord(h->ev_side_done, h->side));
  h->spec_valid = true;
  h->spec_in_flight = true;
  h->spec_ptr = h->src.p;
  h->spec_n = h->n_src;
  h->spec_res = (float)h->cfg.resolution;
  return B200REG_OK;
}

int b200reg_prepare_promotion(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  if (h->cfg.method != B200REG_METHOD_NDT) return B200REG_OK;  // FAST_GICP keeps the source's structures on promotion anyway
  if (!h->have_src || h->n_src == 0) { h->err = "no source to prepare"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  if (h->spec_valid && h->spec_ptr == h->src.p && h->spec_n == h->n_src && h->spec_res == (float)h->cfg.resolution) return B200REG_OK;
  if (!h->side) {
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_src_ready, cudaEventDisableTiming));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_side_done, cudaEventDisableTiming));
  }
  // the side stream starts after the source cloud has arrived (set_source's copy is on the main stream).  The four
  // launches of the build are issued by the NEXT b200reg_align right after its own kernel is on its way, so the host time
  // they cost runs under the registration instead of in front of it; without an align in between, the promotion builds
  // the grid as usual.
  B200_CUDA_TRY(cudaEventRecord(h->ev_src_ready, h->stream));
  h->prepare_pending = true;
  return B200REG_OK;
}

// align in two halves (internal, used by the front end of b200reg_odometry.cu): _begin enqueues the registration (and the
// kernel + copy that fill `output`) and returns; _end waits for the result.  Between the two the host is free — the
// front end enqueues the NEXT scan's filter there, under the registration that is running.
int b200reg_internal_align_begin(b200reg_handle* h, const float* guess, float* aligned_xyzw) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  h->have_result = false;
  h->align_pending = false;
  if (!h->have_tgt) { h->err = "No input target dataset was given!"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = materialize_source(h))) return rc;
  if (!h->have_src || h->n_src == 0) { h->err = "No input source dataset was given!"; return B200REG_E_STATE; }
  if (h->cfg.method == B200REG_METHOD_NDT) {
    if ((rc = run_ndt_single(h, guess, nullptr))) return rc;
  } else if (h->cfg.method == B200REG_METHOD_GICP) {
    if ((rc = run_gicp_single(h, guess))) return rc;
  } else {
    h->err = "align: registration method not available on this handle";
    return B200REG_E_STATE;
  }
  if (h->prepare_pending && (rc = launch_side_build(h))) return rc;  // under the registration that has just been launched
  h->align_aligned_out = aligned_xyzw;
  h->align_aligned_direct = false;
  if (aligned_xyzw) {
    // pcl::Registration::align fills `output` with the transformed source [REF apps/scan_matching_odometry_nodelet.cpp:217-218]:
    // one kernel behind the registration, then DMA straight into a page-locked caller cloud, or through the
    // handle's pinned staging buffer into a pageable one
    B200_CUDA_TRY(h->aligned.reserve(h->n_src));
    h->align_aligned_direct = is_pinned_host(aligned_xyzw);
    if (!h->align_aligned_direct) B200_CUDA_TRY(h->pin_out.reserve(h->n_src));
    launch_counter() += 1;
    k_transform_cloud<<<(h->n_src + 255) / 256, 256, 0, h->stream>>>(h->src.p, h->n_src, h->d_result.p, h->aligned.p);
    B200_CUDA_TRY(cudaMemcpyAsync(h->align_aligned_direct ? (void*)aligned_xyzw : (void*)h->pin_out.p, h->aligned.p, (size_t)h->n_src * 16, cudaMemcpyDeviceToHost, h->stream));
  }
  h->align_pending = true;
  return B200REG_OK;
}

int b200reg_internal_align_end(b200reg_handle* h) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h) return B200REG_E_INVALID;
  if (!h->align_pending) { h->err = "no registration in flight on this handle"; return B200REG_E_STATE; }
  h->align_pending = false;
  int rc = set_device(h);
  if (rc) return rc;
  float* aligned_xyzw = h->align_aligned_out;
  if (aligned_xyzw) B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if ((rc = fetch_result(h, aligned_xyzw != nullptr))) return rc;
  if (aligned_xyzw && !h->align_aligned_direct) host_copy(aligned_xyzw, h->pin_out.p, (size_t)h->n_src * 16);
  return B200REG_OK;
}

int b200reg_internal_set_defer_source_sync(b200reg_handle* h, int on) {
  if (!h) return B200REG_E_INVALID;
  h->defer_source_sync = on != 0;
  return B200REG_OK;
}

int b200reg_align(b200reg_handle* h, const float* guess, float* aligned_xyzw) {
  const int rc = b200reg_internal_align_begin(h, guess, aligned_xyzw);
  if (rc) return rc;
  return b200reg_internal_align_end(h);
}

int b200reg_has_converged(b200reg_handle* h, int* out) {
  if (!h || !out) return B200REG_E_INVALID;
  *out = h->have_result ? h->last.converged : 0;
  return B200REG_OK;
}
int b200reg_get_final_transformation(b200reg_handle* h, float* out16) {
  if (!h || !out16) return B200REG_E_INVALID;
  if (!h->have_result) {
    for (int i = 0; i < 16; ++i) out16[i] = (i % 5 == 0) ? 1.f : 0.f;
    return B200REG_OK;
  }
  memcpy(out16, h->last.transformation, 64);
  return B200REG_OK;
}
int b200reg_get_num_iterations(b200reg_handle* h, int* out) {
  if (!h || !out) return B200REG_E_INVALID;
  *out = h->have_result ? h->last.iterations : 0;
  return B200REG_OK;
}
int b200reg_get_transformation_probability(b200reg_handle* h, double* out) {
  if (!h || !out) return B200REG_E_INVALID;
  *out = h->have_result ? h->last.score : 0.0;
  return B200REG_OK;
}
int b200reg_get_result(b200reg_handle* h, b200reg_result* out) {
  if (!h || !out) return B200REG_E_INVALID;
  if (!h->have_result) return B200REG_E_STATE;
  *out = h->last;
  return B200REG_OK;
}

int b200reg_get_fitness_score(b200reg_handle* h, double max_range, double* out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out) return B200REG_E_INVALID;
  *out = 1.7976931348623157e308;
  if (h->src_is_tgt && set_device(h) == B200REG_OK) materialize_source(h);
  if (!h->have_tgt || !h->have_src || h->n_src == 0) return B200REG_OK;  // DBL_MAX, as upstream with no correspondences
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_nn_grid(h))) return rc;
  float T[16];
  b200reg_get_final_transformation(h, T);
  double sum = 0.0;
  long long cnt = 0;
  B200_CUDA_TRY(nn_fitness(h->stream, h->nn, h->src.p, h->n_src, T, max_range, /*strict_less=*/false, h->fit_partials, h->pin_small, &sum, &cnt));
  h->last.fitness = cnt > 0 ? sum / (double)cnt : 1.7976931348623157e308;
  *out = h->last.fitness;
  return B200REG_OK;
}

int b200reg_calc_fitness_score(b200reg_handle* h, const float* relpose16, double max_range, double* out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out || !relpose16) return B200REG_E_INVALID;
  *out = 1.7976931348623157e308;
  if (h->src_is_tgt && set_device(h) == B200REG_OK) materialize_source(h);
  if (!h->have_tgt || !h->have_src || h->n_src == 0) return B200REG_OK;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_nn_grid(h))) return rc;
  double sum = 0.0;
  long long cnt = 0;
  B200_CUDA_TRY(nn_fitness(h->stream, h->nn, h->src.p, h->n_src, relpose16, max_range, /*strict_less=*/false, h->fit_partials, h->pin_small, &sum, &cnt));
  *out = cnt > 0 ? sum / (double)cnt : 1.7976931348623157e308;
  return B200REG_OK;
}

int b200reg_get_inlier_fraction(b200reg_handle* h, double max_dist, double* out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out) return B200REG_E_INVALID;
  *out = 0.0;
  if (h->src_is_tgt && set_device(h) == B200REG_OK) materialize_source(h);
  if (!h->have_tgt || !h->have_src || h->n_src == 0) return B200REG_OK;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_nn_grid(h))) return rc;
  float T[16];
  b200reg_get_final_transformation(h, T);
  double sum = 0.0;
  long long cnt = 0;
  // k_sq_dists[0] < max_correspondence_dist^2 [REF apps/scan_matching_odometry_nodelet.cpp:328]
  B200_CUDA_TRY(nn_fitness(h->stream, h->nn, h->src.p, h->n_src, T, max_dist * max_dist, /*strict_less=*/true, h->fit_partials, h->pin_small, &sum, &cnt));
  *out = (double)((float)cnt / (float)h->n_src);
  return B200REG_OK;
}

#include "b200reg_api_filters.inl"

#include "b200reg_api_batch.inl"

#include "b200reg_api_map.inl"

// ---- GICP introspection ----------------------------------------------------------------------
int b200reg_gicp_get_covariances(b200reg_handle* h, int which, double* out9, size_t n_points) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out9) return B200REG_E_INVALID;
  if (h->cfg.method != B200REG_METHOD_GICP || !h->have_tgt || !h->have_src) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_gicp_structures(h))) return rc;
  const size_t n = which == 0 ? (size_t)h->n_src : (size_t)h->n_tgt;
  if (n_points < n) return B200REG_E_CAPACITY;
  std::vector<double> c6(n * 6 + 1);
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (n) B200_CUDA_TRY(cudaMemcpy(c6.data(), which == 0 ? h->cov_src.p : h->cov_tgt.p, n * 48, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    const double* c = &c6[6 * i];
    double* o = out9 + 9 * i;
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[1]; o[4] = c[3]; o[5] = c[4]; o[6] = c[2]; o[7] = c[4]; o[8] = c[5];
  }
  return B200REG_OK;
}

// ---- NDT introspection -----------------------------------------------------------------------
int b200reg_ndt_num_leaves(b200reg_handle* h, size_t* out) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out) return B200REG_E_INVALID;
  if (!h->have_tgt) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = ensure_ndt_grid(h))) return rc;
  SortMeta meta;
  B200_CUDA_TRY(cudaMemcpyAsync(&meta, h->grid.sort.meta.p, sizeof(meta), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = meta.n_vox;
  return B200REG_OK;
}

int b200reg_ndt_get_leaves(b200reg_handle* h, uint64_t* idx, int32_t* n, double* mean3, double* cov9, double* icov9, float* centroid3, int32_t* grid6) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  size_t nv = 0;
  int rc = b200reg_ndt_num_leaves(h, &nv);
  if (rc) return rc;
  SortMeta meta;
  B200_CUDA_TRY(cudaMemcpy(&meta, h->grid.sort.meta.p, sizeof(meta), cudaMemcpyDeviceToHost));
  if (grid6) for (int a = 0; a < 3; ++a) { grid6[a] = meta.grid.min_b[a]; grid6[3 + a] = meta.grid.div_b[a]; }
  if (!nv) return B200REG_OK;
  if (idx) {
    std::vector<uint32_t> k(nv);
    B200_CUDA_TRY(cudaMemcpy(k.data(), h->grid.sort.vox_key.p, nv * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nv; ++i) idx[i] = k[i];
  }
  if (n) B200_CUDA_TRY(cudaMemcpy(n, h->grid.leaf_n.p, nv * 4, cudaMemcpyDeviceToHost));
  if (mean3) B200_CUDA_TRY(cudaMemcpy(mean3, h->grid.leaf_mean.p, nv * 24, cudaMemcpyDeviceToHost));
  if (cov9) B200_CUDA_TRY(cudaMemcpy(cov9, h->grid.leaf_cov.p, nv * 72, cudaMemcpyDeviceToHost));
  if (icov9) B200_CUDA_TRY(cudaMemcpy(icov9, h->grid.leaf_icov.p, nv * 72, cudaMemcpyDeviceToHost));
  if (centroid3) {
    std::vector<float4> c(nv);
    B200_CUDA_TRY(cudaMemcpy(c.data(), h->grid.stage_cen.p, nv * 16, cudaMemcpyDeviceToHost));  // per OCCUPIED voxel (the compact array holds records only)
    for (size_t i = 0; i < nv; ++i) { centroid3[3 * i] = c[i].x; centroid3[3 * i + 1] = c[i].y; centroid3[3 * i + 2] = c[i].z; }
  }
  return B200REG_OK;
}

int b200reg_ndt_derivatives(b200reg_handle* h, const double p[6], double* score, double g[6], double H[36]) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !p) return B200REG_E_INVALID;
  if (h->cfg.method != B200REG_METHOD_NDT || !h->have_tgt || !h->have_src || h->n_src == 0) return B200REG_E_STATE;
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(h->pin_small.reserve(sizeof(NdtJob) + sizeof(GicpJob) + sizeof(b200reg_result) + 64 * sizeof(double)));
  if ((rc = run_ndt_single(h, nullptr, p))) return rc;
  double* pd = reinterpret_cast<double*>(h->pin_small.p + sizeof(NdtJob) + sizeof(b200reg_result));
  B200_CUDA_TRY(cudaMemcpyAsync(pd, h->deriv.p, 43 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (score) *score = pd[0];
  if (g) memcpy(g, pd + 1, 6 * sizeof(double));
  if (H) memcpy(H, pd + 7, 36 * sizeof(double));
  return B200REG_OK;
}

int b200reg_set_timing(b200reg_handle* h, int on) {
  if (!h) return B200REG_E_INVALID;
  int rc = set_device(h);
  if (rc) return rc;
  if ((rc = drain_events(h))) return rc;
  h->timing = on != 0;
  h->align_ms = 0.0;
  h->n_align = 0;
  return B200REG_OK;
}

int b200reg_get_counters(b200reg_handle* h, long long* launches_total, long long* timed_aligns, double* align_kernel_ms) {
  if (!h) return B200REG_E_INVALID;
  if (set_device(h) == B200REG_OK) drain_events(h);
  if (launches_total) *launches_total = launch_counter().load();
  if (timed_aligns) *timed_aligns = h->n_align;
  if (align_kernel_ms) *align_kernel_ms = h->align_ms;
  return B200REG_OK;
}

int b200reg_set_profile(b200reg_handle* h, int on) {
  if (!h) return B200REG_E_INVALID;
  h->profile = on != 0;
  return B200REG_OK;
}

int b200reg_get_profile(b200reg_handle* h, long long* out6) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out6) return B200REG_E_INVALID;
  if (!h->prof.p || !h->profile) { h->err = "no profile: b200reg_set_profile(h, 1) before the align"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  B200_CUDA_TRY(cudaMemcpy(out6, h->prof.p, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return B200REG_OK;
}

int b200reg_get_trace(b200reg_handle* h, double* out, size_t cap_records, size_t* n_records) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !n_records) return B200REG_E_INVALID;
  *n_records = 0;
  if (!h->trace.p || !h->profile) { h->err = "no trace: b200reg_set_profile(h, 1) before the align"; return B200REG_E_STATE; }
  int rc = set_device(h);
  if (rc) return rc;
  std::vector<double> t((size_t)(kTraceCap + 1) * kTraceDoubles);
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  B200_CUDA_TRY(cudaMemcpy(t.data(), h->trace.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
  size_t n = (size_t)t[0];
  if (n > (size_t)kTraceCap) n = kTraceCap;
  *n_records = n;
  // record k (1-based: pass k) starts at k * kTraceDoubles; record 0's first slot is the count
  for (size_t k = 0; k < n && k < cap_records && out; ++k) memcpy(out + k * kTraceDoubles, t.data() + (k + 1) * kTraceDoubles, kTraceDoubles * sizeof(double));
  return B200REG_OK;
}

int b200reg_set_sm_budget(b200reg_handle* h, int n_sm) {
  if (!h || n_sm < 1) return B200REG_E_INVALID;
  if (n_sm > h->dev_sm) n_sm = h->dev_sm;
  h->num_sm = n_sm;
  h->vg_sort.max_ctas = n_sm;
  h->grid.sort.max_ctas = n_sm;
  h->nn.sort.max_ctas = n_sm;
  h->nn_src.sort.max_ctas = n_sm;
  h->nn_ror.sort.max_ctas = n_sm;
  return B200REG_OK;
}

int b200reg_set_side_budget(b200reg_handle* h, int n_sm) {
  if (!h || n_sm < 1) return B200REG_E_INVALID;
  h->spec_sms = n_sm > h->dev_sm ? h->dev_sm : n_sm;
  return B200REG_OK;
}

int b200reg_set_sort_path(int path) {
  if (path < 0 || path > 3) return B200REG_E_INVALID;
  sort_path_override() = path;
  return B200REG_OK;
}

int b200reg_get_nn_stats(b200reg_handle* h, long long* out2 /*[3]*/) {
  auto set_error = [&](const std::string& s) { h->err = s; };
  if (!h || !out2) return B200REG_E_INVALID;
  out2[0] = out2[1] = out2[2] = 0;
  if (!h->nn.n_pending.p) return B200REG_OK;
  int rc = set_device(h);
  if (rc) return rc;
  unsigned int np[2] = {0, 0};
  B200_CUDA_TRY(cudaStreamSynchronize(h->stream));
  B200_CUDA_TRY(cudaMemcpy(np, h->nn.n_pending.p, sizeof(np), cudaMemcpyDeviceToHost));
  out2[0] = h->n_src;
  out2[1] = np[0];
  out2[2] = np[1];
  return B200REG_OK;
}

int b200reg_get_stream(b200reg_handle* h, void** out) {
  if (!h || !out) return B200REG_E_INVALID;
  *out = (void*)h->stream;
  return B200REG_OK;
}

}  // extern "C"

// b200reg — loop-closure batches: the B200 replacement of the serial candidate loop of
// LoopDetector::matching [REF include/hdl_graph_slam/loop_detector.hpp:119-173], where every
// candidate keyframe is aligned against the new keyframe (:138-145) and scored with
// getFitnessScore(fitness_score_max_range) (:148).
//
// Keyframe clouds live in a device-side cache keyed by the caller's keyframe id (the reference
// keeps every KeyFrame::cloud for the whole run [REF include/hdl_graph_slam/keyframe.hpp:25-59]).
// For a cloud used as a target the cache also keeps the two search structures derived from it:
//   NDT product : compact voxel records + hash table (what k_ndt_align stages in shared memory)
//   NN product  : cell-ordered copy of the points + cell hash (exact nearest neighbour)
// both built once per target — the hoisted setInputTarget of loop_detector.hpp:124 — by the same
// builders the single-registration path uses; the products are a few MB per keyframe.
//
// A batch is one launch of k_ndt_align (one CTA group per registration, jobs handed out from a
// device-side queue) followed by one exact-NN search + fitness reduction over all pairs; the
// transforms never leave the device between the two.
#pragma once
#include "ndt_grid.cuh"
#include "nn_grid.cuh"

#ifndef B200_FIT_DEPTH
#define B200_FIT_DEPTH 4
#endif
#ifndef B200_FIT_LANES
#define B200_FIT_LANES 1  // lanes per query of the batched near search.  Measured on the 4096-pair batch (fitness part of a step): 1 lane 65.3 ms, 2 lanes 67.4, 4 lanes 72.6 — unlike the GICP linearize pass the batch has queries for every thread, so sharing one only adds votes
#endif
#ifndef B200_FIT_COMPACT
#define B200_FIT_COMPACT 1  // ring 1 of the batched near search on a compacted queue (k_nn_search_batch)
#endif
#ifndef B200_FIT_MINB
#define B200_FIT_MINB 6  // 40 registers: six resident CTAs per SM (measured: 15.8 ms per 1024 pairs; with the 66 registers ptxas picks when left alone, 18.6)
#endif
namespace b200 {

struct CachedCloud {
  DevBuf<float4> pts;
  int n = 0;
  // NDT product
  bool has_ndt = false;
  float ndt_res = 0.f;
  DevBuf<NdtVoxel> voxels;
  DevBuf<float4> centroids;
  DevBuf<uint2> table;
  DevBuf<NdtGridMeta> gmeta;
  DevBuf<SortMeta> meta;
  // NN product
  bool has_nn = false;
  DevBuf<float4> nn_pts;
  DevBuf<uint4> nn_table;
  DevBuf<SortMeta> nn_meta;
  DevBuf<uint32_t> nn_occ;
  uint32_t nn_cap = 0;
  // FAST_GICP product: regularised covariances of the cloud's points (k nearest neighbours), 6 doubles per point
  bool has_cov = false;
  int cov_k = 0, cov_reg = 0;
  DevBuf<double> cov;

  void release() {
    cov.release();
    has_cov = false;
    pts.release(); voxels.release(); centroids.release(); table.release(); gmeta.release(); meta.release();
    nn_pts.release(); nn_table.release(); nn_meta.release(); nn_occ.release();
    has_ndt = has_nn = false;
    n = 0;
  }
  NdtGridView ndt_view() const {
    NdtGridView v;
    v.meta = meta.p; v.gmeta = gmeta.p; v.table = table.p; v.voxels = voxels.p; v.centroids = centroids.p;
    return v;
  }
  NnView nn_view() const {
    NnView v;
    v.meta = nn_meta.p; v.table = nn_table.p; v.table_mask = nn_cap - 1; v.table_shift = 32 - (int)__builtin_ctz(nn_cap);
    v.pts = nn_pts.p; v.n = n; v.occ = nn_occ.p;
    return v;
  }
};

// Build the products with the handle's builders, then swap the freshly written arrays into the
// cache entry (pointer swaps; the builder re-grows its own buffers on its next use).
inline cudaError_t cache_build_ndt(cudaStream_t st, NdtGrid& builder, CachedCloud& c, float resolution) {
  std::swap(builder.voxels, c.voxels); std::swap(builder.centroids, c.centroids); std::swap(builder.table, c.table);
  std::swap(builder.gmeta, c.gmeta); std::swap(builder.sort.meta, c.meta);
  cudaError_t e = builder.build(st, c.pts.p, c.n, resolution);
  std::swap(builder.voxels, c.voxels); std::swap(builder.centroids, c.centroids); std::swap(builder.table, c.table);
  std::swap(builder.gmeta, c.gmeta); std::swap(builder.sort.meta, c.meta);
  builder.built = false;  // the builder's own view is now stale
  if (e == cudaSuccess) { c.has_ndt = true; c.ndt_res = resolution; }
  return e;
}

inline cudaError_t cache_build_nn(cudaStream_t st, NnGrid& builder, CachedCloud& c) {
  std::swap(builder.pts, c.nn_pts); std::swap(builder.table, c.nn_table); std::swap(builder.sort.meta, c.nn_meta); std::swap(builder.occ, c.nn_occ);
  cudaError_t e = builder.build(st, c.pts.p, c.n);
  c.nn_cap = builder.table_cap;
  std::swap(builder.pts, c.nn_pts); std::swap(builder.table, c.nn_table); std::swap(builder.sort.meta, c.nn_meta); std::swap(builder.occ, c.nn_occ);
  builder.built = false;
  if (e == cudaSuccess) c.has_nn = true;
  return e;
}

// ---- batched getFitnessScore ------------------------------------------------------------------
struct FitJob {
  NnView view;          // exact-NN structure of the pair's target
  const float4* src;    // the pair's source cloud
  int n_src;
  int result;           // index into the result array (transform in, fitness out)
  long long d2_offset;  // first slot of this pair in the d2 array
};

constexpr float kNoNeighbour = __builtin_huge_valf();  // d2 of a query with nothing inside max_range

__device__ __forceinline__ void fit_transform(const float* T, const float4 p, float& qx, float& qy, float& qz) {
  qx = affine_row(T[0], T[4], T[8], T[12], p.x, p.y, p.z);
  qy = affine_row(T[1], T[5], T[9], T[13], p.x, p.y, p.z);
  qz = affine_row(T[2], T[6], T[10], T[14], p.x, p.y, p.z);
}

// near phase: blockIdx.y = job, blockIdx.x = 256-point slice of its source.  Transform by the pair's
// final transformation (float, pcl::transformPoint order) and search rings 0..1; unresolved queries
// go to `pending` with their best-so-far in d2_out.
__global__ void __launch_bounds__(256, B200_FIT_MINB) k_nn_search_batch(const FitJob* __restrict__ jobs, const b200reg_result* __restrict__ results, float max_d2, float* __restrict__ d2_out,
                                                         uint2* __restrict__ pending, unsigned int* __restrict__ n_pending) {
  const FitJob& job = jobs[blockIdx.y];
  constexpr int QL = B200_FIT_LANES;            // lanes that share one query (1 = a thread per query)
  constexpr int PER_BLOCK = 256 / QL;
  const int i = blockIdx.x * PER_BLOCK + threadIdx.x / QL;
  if (blockIdx.x * PER_BLOCK >= job.n_src) return;
  __shared__ float T[16];
  __shared__ GridParams s_gp;  // the target's cell lattice, shared by the CTA instead of 20 registers per thread
  if (threadIdx.x < 16) T[threadIdx.x] = results[job.result].transformation[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 32 + (int)(sizeof(GridParams) / 4))
    reinterpret_cast<uint32_t*>(&s_gp)[threadIdx.x - 32] = reinterpret_cast<const uint32_t*>(&job.view.meta->grid)[threadIdx.x - 32];
  __syncthreads();
  const bool active = i < job.n_src;
  if (QL == 1 && !B200_FIT_COMPACT && !active) return;
  const GridParams& gp = s_gp;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) fit_transform(T, __ldg(job.src + i), qx, qy, qz);
  float best = 3.402823466e+38f;
  int best_idx = kNoIndex;
  int st = kNnDone;
  const bool searchable = job.view.n > 0 && gp.any && !gp.overflow;
  if (QL == 1 && B200_FIT_COMPACT) {
    // ring 0 by every thread; the queries it leaves open are packed into a queue in shared memory and ring 1 is walked by
    // the first threads of the block, one open query each: full warps instead of scattered lanes
    __shared__ float4 s_q[256];   // transformed point, w = best so far
    __shared__ int2 s_qi[256];    // (best index, source index)
    __shared__ int s_wn[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool open = false;
    if (active && searchable) open = !nn_near_ring0<B200_FIT_DEPTH>(job.view, gp, nn_make_query(gp, qx, qy, qz), max_d2, best, best_idx);
    const unsigned m = __ballot_sync(0xffffffffu, open);
    if (lane == 0) s_wn[warp] = __popc(m);
    if (active && !open) d2_out[job.d2_offset + i] = best_idx != kNoIndex ? best : kNoNeighbour;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) base += s_wn[w];
      total += s_wn[w];
    }
    if (open) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      s_q[slot] = make_float4(qx, qy, qz, best);
      s_qi[slot] = make_int2(best_idx, i);
    }
    __syncthreads();
    if ((int)threadIdx.x >= total) return;
    const float4 e = s_q[threadIdx.x];
    const int2 ei = s_qi[threadIdx.x];
    best = e.w;
    best_idx = ei.x;
    st = nn_near_ring1<B200_FIT_DEPTH>(job.view, gp, nn_make_query(gp, e.x, e.y, e.z), max_d2, best, best_idx);
    d2_out[job.d2_offset + ei.y] = best_idx != kNoIndex ? best : kNoNeighbour;
    if (st != kNnDone) pending[atomicAdd(n_pending, 1u)] = make_uint2(blockIdx.y, (unsigned)ei.y);
    return;
  } else if (QL == 1) {
    if (searchable) st = nn_query_near<B200_FIT_DEPTH>(job.view, gp, nn_make_query(gp, qx, qy, qz), max_d2, best, best_idx);
  } else {
    // QL lanes walk one query together (nn_query_near_group): the per-thread loops of a thread-per-query search have very
    // different trip counts (ncu: 10 of 32 lanes active), a group's lanes share them
    st = nn_query_near_group<(QL > 1 ? QL : 2), false>(job.view, gp, nn_make_query(gp, qx, qy, qz), max_d2, (int)(threadIdx.x % QL), active && searchable, best, best_idx);
    if (!active || (threadIdx.x % QL) != 0) return;
  }
  d2_out[job.d2_offset + i] = best_idx != kNoIndex ? best : kNoNeighbour;
  if (st != kNnDone) pending[atomicAdd(n_pending, 1u)] = make_uint2(blockIdx.y, (unsigned)i | (st == kNnBail ? kBailFlag : 0u));
}

// far phase: one warp per pending (job, point); still-open queries move to pending2
__global__ void __launch_bounds__(256) k_nn_far_batch(const FitJob* __restrict__ jobs, const b200reg_result* __restrict__ results, const uint2* __restrict__ pending,
                                                      const unsigned int* __restrict__ n_pending, float max_d2, float* __restrict__ d2_out, uint2* __restrict__ pending2,
                                                      unsigned int* __restrict__ n_pending2) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int np = (int)*n_pending;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < np; w += warps) {
    uint2 pe = pending[w];
    const bool bailed = (pe.y & kBailFlag) != 0u;
    pe.y &= ~kBailFlag;
    const FitJob& job = jobs[pe.x];
    const GridParams gp = job.view.meta->grid;
    float qx, qy, qz;
    fit_transform(results[job.result].transformation, __ldg(job.src + pe.y), qx, qy, qz);
    const float prev = d2_out[job.d2_offset + pe.y];
    float best = prev == kNoNeighbour ? 3.402823466e+38f : prev;
    int best_idx = prev == kNoNeighbour ? kNoIndex : 0;  // the index itself is not needed for the score
    const bool ok = nn_query_far_warp(job.view, gp, nn_make_query(gp, qx, qy, qz), max_d2, kFarRing, lane, best, best_idx, bailed);
    __syncwarp();
    if (lane == 0) {
      d2_out[job.d2_offset + pe.y] = best_idx != kNoIndex ? best : kNoNeighbour;
      if (!ok) pending2[atomicAdd(n_pending2, 1u)] = pe;
    }
  }
}

// brute phase: one CTA per still-open (job, point) scans the pair's whole target
__global__ void __launch_bounds__(256) k_nn_bruteforce_batch(const FitJob* __restrict__ jobs, const b200reg_result* __restrict__ results, const uint2* __restrict__ pending,
                                                             const unsigned int* __restrict__ n_pending, float* __restrict__ d2_out) {
  const int np = (int)*n_pending;
  for (int w = blockIdx.x; w < np; w += gridDim.x) {
    const uint2 pe = pending[w];
    const FitJob& job = jobs[pe.x];
    float qx, qy, qz;
    fit_transform(results[job.result].transformation, __ldg(job.src + pe.y), qx, qy, qz);
    float best;
    int best_idx;
    nn_query_brute_block(job.view, qx, qy, qz, best, best_idx);
    if (threadIdx.x == 0) d2_out[job.d2_offset + pe.y] = best_idx != kNoIndex ? best : kNoNeighbour;
  }
}

// one CTA per pair: mean of the squared distances <= max_range over the points that have a
// neighbour, DBL_MAX when none [REF src/hdl_graph_slam/information_matrix_calculator.cpp:96-107];
// fixed summation order (thread-strided double sums, shuffle tree, 8 warps in order)
__global__ void __launch_bounds__(256) k_fitness_batch(const FitJob* __restrict__ jobs, const float* __restrict__ d2, double max_range, b200reg_result* __restrict__ results) {
  const FitJob& job = jobs[blockIdx.x];
  __shared__ double s_sum[8], s_cnt[8];
  double sum = 0.0, cnt = 0.0;
  const float* d = d2 + job.d2_offset;
  for (int i = threadIdx.x; i < job.n_src; i += 256) {
    const float v = d[i];
    if (v == kNoNeighbour) continue;
    const double dv = (double)v;
    if (dv <= max_range) { sum += dv; cnt += 1.0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += s_sum[w]; b += s_cnt[w]; }
    results[job.result].fitness = b > 0 ? a / b : 1.7976931348623157e308;
  }
}

}  // namespace b200

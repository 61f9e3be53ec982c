// b200reg — NDT target grid: the B200 replacement of pclomp::VoxelGridCovariance::filter
// (SURVEY.md A.3), reached through ndt->setInputTarget
// [REF apps/scan_matching_odometry_nodelet.cpp:180,254; include/hdl_graph_slam/loop_detector.hpp:124].
//
// Layout in HBM (all per target cloud):
//   leaf_*   : one entry per OCCUPIED voxel, ascending linear voxel index (introspection)
//   voxels   : compact array of 48-byte records {double mean[3]; float icov[6]}, one per voxel with
//              >= 6 points; centroids: float centroid of the same voxels (KDTREE radius search)
//   table    : open-addressing hash  linear voxel index -> compact record index  (voxels rejected
//              by the eigenvalue test carry kNdtRejected and serve the KDTREE mode only).  Its
//              capacity is chosen ON THE DEVICE from the number of records (load factor <= 1/2),
//              so no data-dependent size ever reaches the host.
// records + table of an HDL-64 scan at 1 m are ~100 KB: the align kernel stages them in the
// 227 KB shared memory of every SM once per registration and every voxel probe of every pass
// is then an LDS instead of an L2 round trip.
//
// The per-voxel statistics kernel lives in ndt_leaf.cu, a translation unit compiled with
// -fmad=false: sums run in input order and every double operation is rounded separately, so
// the eigenvalue test that decides whether a (near-)planar voxel is valid sees the same bits as
// the reference-order CPU evaluation instead of flipping on contraction noise.
#pragma once
#include "voxel_sort.cuh"

namespace b200 {

// mean as an unevaluated float pair (hi + lo carries 48 bits of the double mean): the pass forms
// q = (x' - hi) - lo in float, which equals the reference's float(double(x') - mean) up to a
// last-place double-rounding case, without any FP64 conversion in the per-hit path.
struct __align__(16) NdtVoxel {
  float mean_hi[3];
  float mean_lo[3];
  float icov[6];  // xx xy xz yy yz zz
};

struct NdtGridMeta {  // device-resident, written by the build kernels
  uint32_t n_records;  // voxels with >= 6 points (valid + eigen-rejected)
  uint32_t table_cap;  // power of two >= 2 * n_records (>= 16)
};

struct NdtGridView {  // what the align kernels need, by value
  const SortMeta* meta;
  const NdtGridMeta* gmeta;
  const uint2* table;  // (key, record | flag); key == kInvalidKey marks an empty cell
  const NdtVoxel* voxels;
  const float4* centroids;
};

constexpr uint32_t kNdtRejected = 0x40000000u;  // table value flag: >= 6 points but failed the eigenvalue test

__device__ __forceinline__ uint32_t ndt_hash(uint32_t key, uint32_t mask) {
  uint32_t h = key * 2654435761u;
  return (h ^ (h >> 15)) & mask;
}

struct NdtLeafArgs {
  const float4* pts;
  const uint32_t *vals_a, *vals_b;
  const SortMeta* meta;
  const uint32_t *vox_start, *vox_key;
  int min_points;
  double eig_mult;
  NdtGridMeta* gmeta;
  NdtVoxel* voxels;
  float4* centroids;
  uint32_t* rec_key;   // scratch, one entry per occupied slot: record index | kNdtRejected, or 0xFFFFFFFF (k_ndt_table)
  uint2* table;
  int32_t* leaf_n;
  double *leaf_mean, *leaf_cov, *leaf_icov;
};
// ndt_leaf.cu (-fmad=false): statistics -> records, then the hash over the records
cudaError_t ndt_leaf_prefer_shared();  // maximum shared-memory carve-out for the kernels of ndt_leaf.cu (see init_kernel_attributes)
cudaError_t launch_ndt_leaf_stats(cudaStream_t st, const NdtLeafArgs& a, NdtVoxel* stage_vox, float4* stage_cen, double* sums, float* csum);

struct NdtGrid {
  VoxelSort sort;
  DevBuf<double> leaf_mean, leaf_cov, leaf_icov, sums;
  DevBuf<float> csum;
  DevBuf<int32_t> leaf_n;
  DevBuf<NdtVoxel> voxels, stage_vox;
  DevBuf<float4> centroids, stage_cen;
  DevBuf<uint32_t> rec_key;
  DevBuf<uint2> table;
  DevBuf<NdtGridMeta> gmeta;
  int n_points = 0;
  bool built = false;

  void release() {
    sort.release(); sums.release(); csum.release(); leaf_mean.release(); leaf_cov.release(); leaf_icov.release(); leaf_n.release(); voxels.release(); centroids.release(); stage_vox.release(); stage_cen.release();
    rec_key.release(); table.release(); gmeta.release();
  }

  NdtGridView view() const {
    NdtGridView v;
    v.meta = sort.meta.p;
    v.gmeta = gmeta.p;
    v.table = table.p;
    v.voxels = voxels.p;
    v.centroids = centroids.p;
    return v;
  }

  cudaError_t build(cudaStream_t st, const float4* d_pts, int n, float resolution) {
    cudaError_t e;
    n_points = n;
    if ((e = sort.run(st, d_pts, n, /*is_dense=*/1, resolution, resolution, resolution, false)) != cudaSuccess) return e;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    const size_t max_rec = nn / 6 + 1;  // records hold >= 6 points each
    size_t cap = 16;
    while (cap < 2 * max_rec) cap <<= 1;
    if ((e = leaf_mean.reserve(nn * 3)) != cudaSuccess) return e;
    if ((e = sums.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = csum.reserve(nn * 3)) != cudaSuccess) return e;
    if ((e = leaf_cov.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = leaf_icov.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = leaf_n.reserve(nn)) != cudaSuccess) return e;
    if ((e = voxels.reserve(max_rec)) != cudaSuccess) return e;
    if ((e = stage_vox.reserve(nn)) != cudaSuccess) return e;
    if ((e = stage_cen.reserve(nn)) != cudaSuccess) return e;
    if ((e = centroids.reserve(max_rec)) != cudaSuccess) return e;
    if ((e = rec_key.reserve(nn)) != cudaSuccess) return e;
    if ((e = table.reserve(cap)) != cudaSuccess) return e;
    if ((e = gmeta.reserve(1)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(gmeta.p, 0, sizeof(NdtGridMeta), st)) != cudaSuccess) return e;
    NdtLeafArgs a;
    a.pts = d_pts; a.vals_a = sort.vals_a.p; a.vals_b = sort.vals_b.p; a.meta = sort.meta.p; a.vox_start = sort.vox_start.p; a.vox_key = sort.vox_key.p;
    a.min_points = 6; a.eig_mult = 0.01;
    a.gmeta = gmeta.p; a.voxels = voxels.p; a.centroids = centroids.p; a.rec_key = rec_key.p; a.table = table.p;
    a.leaf_n = leaf_n.p; a.leaf_mean = leaf_mean.p; a.leaf_cov = leaf_cov.p; a.leaf_icov = leaf_icov.p;
    if ((e = launch_ndt_leaf_stats(st, a, stage_vox.p, stage_cen.p, sums.p, csum.p)) != cudaSuccess) return e;
    built = true;
    return cudaGetLastError();
  }
};

}  // namespace b200

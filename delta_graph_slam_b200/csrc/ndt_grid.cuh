// b200reg — NDT target grid: the B200 replacement of pclomp::VoxelGridCovariance::filter
// (SURVEY.md A.3), reached through ndt->setInputTarget
// [REF apps/scan_matching_odometry_nodelet.cpp:180,254; include/hdl_graph_slam/loop_detector.hpp:124].
//
// Layout in HBM (all per target cloud):
//   leaf_*   : one entry per OCCUPIED voxel, ascending linear voxel index (introspection + KDTREE)
//   voxels   : 48-byte records {double mean[3]; float icov[6]} of the VALID voxels (>= 6 points,
//              eigenvalue test passed), indexed by the occupied-voxel slot
//   table    : open-addressing hash  linear voxel index -> slot  holding valid voxels only;
//              sized from the point count so no data-dependent size ever reaches the host
#pragma once
#include "small_solve.cuh"
#include "voxel_sort.cuh"

namespace b200 {

struct __align__(16) NdtVoxel {
  double mean[3];
  float icov[6];  // xx xy xz yy yz zz
};

struct NdtGridView {  // what the align kernels need, by value
  const SortMeta* meta;
  const uint2* table;  // (key, slot); key == kInvalidKey marks an empty cell
  uint32_t table_mask;
  int table_shift;     // 32 - log2(capacity)
  const NdtVoxel* voxels;
  const float4* centroids;  // per occupied slot: xyz = float centroid, w = (float)valid (KDTREE search)
};

__device__ __forceinline__ uint32_t ndt_hash(uint32_t key, int shift) { return (key * 2654435761u) >> shift; }

__device__ __forceinline__ int ndt_lookup(const NdtGridView& g, uint32_t key) {
  uint32_t h = ndt_hash(key, g.table_shift);
  while (true) {
    uint2 e = __ldg(g.table + h);
    if (e.x == key) return (int)e.y;
    if (e.x == kInvalidKey) return -1;
    h = (h + 1) & g.table_mask;
  }
}

// pass 1: one warp per occupied voxel: n, sum x, sum x x^T in double (lanes stride the run,
// fixed shuffle tree afterwards -> deterministic)
__global__ void __launch_bounds__(256) k_ndt_leaf_sums(const float4* __restrict__ pts, const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
                                                       const SortMeta* __restrict__ meta, const uint32_t* __restrict__ vox_start, double* __restrict__ sums /*[n_vox][9]*/) {
  const uint32_t* vals = sorted_in_b(meta) ? vals_b : vals_a;
  const int n_vox = (int)meta->n_vox;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < n_vox; slot += warps) {
    const uint32_t s = vox_start[slot], e = vox_start[slot + 1];
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t j = s + lane; j < e; j += 32) {
      float4 p = __ldg(pts + vals[j]);
      double x = (double)p.x, y = (double)p.y, z = (double)p.z;
      a[0] += x; a[1] += y; a[2] += z;
      a[3] += x * x; a[4] += x * y; a[5] += x * z; a[6] += y * y; a[7] += y * z; a[8] += z * z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 9; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane < 9) sums[(size_t)slot * 9 + lane] = a[lane];
  }
}

// pass 2: one thread per occupied voxel: mean, covariance, eigenvalue clamp, inverse, hash insert
__global__ void __launch_bounds__(128) k_ndt_leaf_finalize(const SortMeta* __restrict__ meta, const uint32_t* __restrict__ vox_start, const uint32_t* __restrict__ vox_key,
                                                           const double* __restrict__ sums, int min_points, double eig_mult, NdtVoxel* __restrict__ voxels,
                                                           uint2* __restrict__ table, uint32_t table_mask, int table_shift, int32_t* __restrict__ leaf_n,
                                                           double* __restrict__ leaf_mean, double* __restrict__ leaf_cov, double* __restrict__ leaf_icov, float4* __restrict__ centroids) {
  const int n_vox = (int)meta->n_vox;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_vox; slot += gridDim.x * blockDim.x) {
    const int n = (int)(vox_start[slot + 1] - vox_start[slot]);
    const double* s = sums + (size_t)slot * 9;
    const double nd = (double)n;
    const double pt_sum[3] = {s[0], s[1], s[2]};
    double mean[3] = {pt_sum[0] / nd, pt_sum[1] / nd, pt_sum[2] / nd};
    int nr_points = n;
    double cov[9] = {s[3], s[4], s[5], s[4], s[6], s[7], s[5], s[7], s[8]};
    double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool valid = false;
    if (n >= min_points) {
      // cov = (cov - 2 (pt_sum mean^T)) / n + mean mean^T ;  cov *= (n - 1) / n      (A.3)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) cov[3 * a + b] = (cov[3 * a + b] - 2.0 * (pt_sum[a] * mean[b])) / nd + mean[a] * mean[b];
#pragma unroll
      for (int k = 0; k < 9; ++k) cov[k] *= (nd - 1.0) / nd;
      double ev[3], V[9];
      sym_eigen3(cov, ev, V);
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        nr_points = -1;
      } else {
        const double min_ev = eig_mult * ev[2];
        if (ev[0] < min_ev) {
          ev[0] = min_ev;
          if (ev[1] < min_ev) ev[1] = min_ev;
          // cov = V diag(ev) V^-1, V orthonormal
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) cov[3 * a + b] = V[3 * a + 0] * ev[0] * V[3 * b + 0] + V[3 * a + 1] * ev[1] * V[3 * b + 1] + V[3 * a + 2] * ev[2] * V[3 * b + 2];
        }
        inverse3(cov, icov);
        double mx = -1.7e308, mn = 1.7e308;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mx = fmax(mx, icov[k]); mn = fmin(mn, icov[k]); }
        if (isinf(mx) || isinf(mn) || mx != mx || mn != mn) nr_points = -1;
        else valid = true;
      }
    }
    leaf_n[slot] = nr_points;
#pragma unroll
    for (int a = 0; a < 3; ++a) leaf_mean[(size_t)slot * 3 + a] = mean[a];
#pragma unroll
    for (int k = 0; k < 9; ++k) { leaf_cov[(size_t)slot * 9 + k] = cov[k]; leaf_icov[(size_t)slot * 9 + k] = icov[k]; }
    centroids[slot] = make_float4((float)mean[0], (float)mean[1], (float)mean[2], valid ? 1.f : 0.f);
    if (valid) {
      NdtVoxel v;
      v.mean[0] = mean[0]; v.mean[1] = mean[1]; v.mean[2] = mean[2];
      v.icov[0] = (float)icov[0]; v.icov[1] = (float)icov[1]; v.icov[2] = (float)icov[2];
      v.icov[3] = (float)icov[4]; v.icov[4] = (float)icov[5]; v.icov[5] = (float)icov[8];
      voxels[slot] = v;
      const uint32_t key = vox_key[slot];
      uint32_t h = ndt_hash(key, table_shift);
      while (true) {
        uint32_t old = atomicCAS(&table[h].x, kInvalidKey, key);
        if (old == kInvalidKey) { table[h].y = (uint32_t)slot; break; }
        h = (h + 1) & table_mask;
      }
    }
  }
}

struct NdtGrid {
  VoxelSort sort;
  DevBuf<double> sums, leaf_mean, leaf_cov, leaf_icov;
  DevBuf<int32_t> leaf_n;
  DevBuf<NdtVoxel> voxels;
  DevBuf<float4> centroids;
  DevBuf<uint2> table;
  uint32_t table_cap = 0;
  int n_points = 0;
  bool built = false;

  void release() {
    sort.release(); sums.release(); leaf_mean.release(); leaf_cov.release(); leaf_icov.release(); leaf_n.release(); voxels.release(); centroids.release(); table.release();
  }

  NdtGridView view() const {
    NdtGridView v;
    v.meta = sort.meta.p;
    v.table = table.p;
    v.table_mask = table_cap - 1;
    v.table_shift = 32 - (int)__builtin_ctz(table_cap);
    v.voxels = voxels.p;
    v.centroids = centroids.p;
    return v;
  }

  cudaError_t build(cudaStream_t st, const float4* d_pts, int n, float resolution) {
    cudaError_t e;
    n_points = n;
    if ((e = sort.run(st, d_pts, n, /*is_dense=*/1, resolution, resolution, resolution, false)) != cudaSuccess) return e;
    size_t nn = (size_t)(n > 0 ? n : 1);
    if ((e = sums.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = leaf_mean.reserve(nn * 3)) != cudaSuccess) return e;
    if ((e = leaf_cov.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = leaf_icov.reserve(nn * 9)) != cudaSuccess) return e;
    if ((e = leaf_n.reserve(nn)) != cudaSuccess) return e;
    if ((e = voxels.reserve(nn)) != cudaSuccess) return e;
    if ((e = centroids.reserve(nn)) != cudaSuccess) return e;
    // valid voxels hold >= 6 points each: at most n / 6 entries -> load factor <= 1/3
    uint32_t cap = 64;
    while (cap < (uint32_t)(n / 2 + 1)) cap <<= 1;
    table_cap = cap;
    if ((e = table.reserve(cap)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(table.p, 0xFF, (size_t)cap * sizeof(uint2), st)) != cudaSuccess) return e;
    k_ndt_leaf_sums<<<kNumSM * 2, 256, 0, st>>>(d_pts, sort.vals_a.p, sort.vals_b.p, sort.meta.p, sort.vox_start.p, sums.p);
    k_ndt_leaf_finalize<<<kNumSM, 128, 0, st>>>(sort.meta.p, sort.vox_start.p, sort.vox_key.p, sums.p, 6, 0.01, voxels.p, table.p, cap - 1, 32 - (int)__builtin_ctz(cap),
                                                 leaf_n.p, leaf_mean.p, leaf_cov.p, leaf_icov.p, centroids.p);
    built = true;
    return cudaGetLastError();
  }
};

}  // namespace b200

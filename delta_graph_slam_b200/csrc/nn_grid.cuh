// b200reg — exact nearest-neighbour search on the target cloud: the B200 replacement of the FLANN
// kd-tree that pcl::Registration::initCompute builds and getFitnessScore queries,
// [REF include/hdl_graph_slam/loop_detector.hpp:148; apps/scan_matching_odometry_nodelet.cpp:318-332;
//  src/hdl_graph_slam/information_matrix_calculator.cpp:77-108].
//
// Structure: a uniform cell grid over the target.  Points are re-ordered by cell with the same
// key / radix-sort / segmentation machinery as VoxelGrid (w of each re-ordered float4 carries the
// original index), and an open-addressing hash maps cell -> run.  A query walks Chebyshev rings
// of cells around its own cell, prunes cells by box distance, and stops when nothing outside the
// examined block can beat the best distance: the result is EXACT, with FLANN's float metric
// ((dx*dx)+dy*dy)+dz*dz and ties broken by lowest index (the oracle's convention).  Queries that
// are still open after kMaxRing rings (far outliers; max_range defaults to DBL_MAX upstream) are
// finished by a warp-per-query brute-force kernel, so the search is exact at any range.
#pragma once
#include "voxel_sort.cuh"

namespace b200 {

constexpr int kMaxRing = 4;
constexpr float kNnCell = 0.5f;

struct NnView {
  const SortMeta* meta;
  const uint2* table;
  uint32_t table_mask;
  int table_shift;
  const uint32_t* cell_start;  // [n_cells + 1]
  const float4* pts;           // re-ordered by cell; w = original index (int bits)
  int n;
};

__device__ __forceinline__ int nn_lookup(const NnView& g, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> g.table_shift;
  while (true) {
    uint2 e = __ldg(g.table + h);
    if (e.x == key) return (int)e.y;
    if (e.x == kInvalidKey) return -1;
    h = (h + 1) & g.table_mask;
  }
}

__global__ void __launch_bounds__(256) k_nn_reorder(const float4* __restrict__ pts, int n, const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
                                                    const SortMeta* __restrict__ meta, float4* __restrict__ out) {
  const uint32_t* vals = sorted_in_b(meta) ? vals_b : vals_a;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i >= (int)meta->n_valid) return;
  const uint32_t src = vals[i];
  float4 p = __ldg(pts + src);
  p.w = __int_as_float((int)src);
  out[i] = p;
}

__global__ void __launch_bounds__(256) k_nn_insert(const SortMeta* __restrict__ meta, const uint32_t* __restrict__ vox_key, uint2* __restrict__ table, uint32_t mask, int shift) {
  const int n_vox = (int)meta->n_vox;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_vox; slot += gridDim.x * blockDim.x) {
    const uint32_t key = vox_key[slot];
    uint32_t h = (key * 2654435761u) >> shift;
    while (true) {
      uint32_t old = atomicCAS(&table[h].x, kInvalidKey, key);
      if (old == kInvalidKey) { table[h].y = (uint32_t)slot; break; }
      h = (h + 1) & mask;
    }
  }
}

// exact 1-NN of q.  Returns true when resolved (best / best_idx final, best_idx = -1 if nothing
// lies within max_d2); false when the ring budget ran out (caller defers to the brute-force pass).
__device__ __forceinline__ bool nn_query(const NnView& g, const GridParams& gp, float qx, float qy, float qz, float max_d2, float& best, int& best_idx) {
  const float c = gp.leaf[0];
  const int cx = (int)floorf(__fmul_rn(qx, gp.inv_leaf[0])), cy = (int)floorf(__fmul_rn(qy, gp.inv_leaf[1])), cz = (int)floorf(__fmul_rn(qz, gp.inv_leaf[2]));
  const float margin = 1e-3f * c + 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz));
  best = 3.402823466e+38f;
  best_idx = -1;
  // true when every point outside the (2r-1)^3 block around the query cell is provably no better
  // than `best`, lies beyond max_d2, or does not exist (block covers the occupied lattice)
  auto settled = [&](int r) {
    const float gx = fminf(qx - (float)(cx - (r - 1)) * c, (float)(cx + r) * c - qx);
    const float gy = fminf(qy - (float)(cy - (r - 1)) * c, (float)(cy + r) * c - qy);
    const float gz = fminf(qz - (float)(cz - (r - 1)) * c, (float)(cz + r) * c - qz);
    const float gap = fmaxf(fminf(gx, fminf(gy, gz)) - margin, 0.f);
    const float gap2 = gap * gap;
    if (best <= gap2 || gap2 > max_d2) return true;
    return cx - (r - 1) <= gp.min_b[0] && cx + (r - 1) >= gp.max_b[0] && cy - (r - 1) <= gp.min_b[1] && cy + (r - 1) >= gp.max_b[1] && cz - (r - 1) <= gp.min_b[2] &&
           cz + (r - 1) >= gp.max_b[2];
  };
  for (int r = 0; r <= kMaxRing; ++r) {
    if (r >= 1 && settled(r)) return true;
    const int z0 = max(cz - r, gp.min_b[2]), z1 = min(cz + r, gp.max_b[2]);
    const int y0 = max(cy - r, gp.min_b[1]), y1 = min(cy + r, gp.max_b[1]);
    for (int iz = z0; iz <= z1; ++iz) {
      const float dzl = (float)iz * c - margin - qz, dzh = qz - ((float)(iz + 1) * c + margin);
      const float dz = fmaxf(fmaxf(dzl, dzh), 0.f);
      const bool zface = (iz == cz - r) || (iz == cz + r);
      for (int iy = y0; iy <= y1; ++iy) {
        const float dyl = (float)iy * c - margin - qy, dyh = qy - ((float)(iy + 1) * c + margin);
        const float dy = fmaxf(fmaxf(dyl, dyh), 0.f);
        const float dyz2 = dz * dz + dy * dy;
        if (dyz2 > best) continue;
        const bool face = zface || (iy == cy - r) || (iy == cy + r);
        const int xstep = face ? 1 : (r == 0 ? 1 : 2 * r);
        for (int ix = cx - r; ix <= cx + r; ix += xstep) {
          if (ix < gp.min_b[0] || ix > gp.max_b[0]) continue;
          const float dxl = (float)ix * c - margin - qx, dxh = qx - ((float)(ix + 1) * c + margin);
          const float dx = fmaxf(fmaxf(dxl, dxh), 0.f);
          if (dyz2 + dx * dx > best) continue;
          const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
          const int slot = nn_lookup(g, key);
          if (slot < 0) continue;
          const uint32_t s = __ldg(g.cell_start + slot), e = __ldg(g.cell_start + slot + 1);
          for (uint32_t j = s; j < e; ++j) {
            const float4 p = __ldg(g.pts + j);
            const float d = l2_simple(qx, qy, qz, p.x, p.y, p.z);
            const int idx = __float_as_int(p.w);
            if (d < best || (d == best && idx < best_idx)) { best = d; best_idx = idx; }
          }
        }
      }
    }
  }
  return settled(kMaxRing + 1);
}

// One thread per source point: transform by T (column-major 4x4, float, pcl::transformPoint order)
// and search.  d2_out[i] = squared NN distance (float), idx_out[i] = target index or -1; unresolved
// queries are appended to `pending`.
__global__ void __launch_bounds__(256) k_nn_search(NnView g, const float4* __restrict__ src, int n_src, const float* __restrict__ T16, int use_T, float max_d2,
                                                   float* __restrict__ d2_out, int* __restrict__ idx_out, float4* __restrict__ q_out, int* __restrict__ pending,
                                                   unsigned int* __restrict__ n_pending) {
  __shared__ float T[16];
  if (threadIdx.x < 16) T[threadIdx.x] = use_T ? T16[threadIdx.x] : ((threadIdx.x % 5 == 0) ? 1.f : 0.f);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_src) return;
  const GridParams gp = g.meta->grid;
  const float4 p = __ldg(src + i);
  float qx = p.x, qy = p.y, qz = p.z;
  if (use_T) {
    qx = affine_row(T[0], T[4], T[8], T[12], p.x, p.y, p.z);
    qy = affine_row(T[1], T[5], T[9], T[13], p.x, p.y, p.z);
    qz = affine_row(T[2], T[6], T[10], T[14], p.x, p.y, p.z);
  }
  if (q_out) q_out[i] = make_float4(qx, qy, qz, 1.f);
  float best = 3.402823466e+38f;
  int best_idx = -1;
  bool ok = true;
  if (g.n > 0 && gp.any && !gp.overflow) ok = nn_query(g, gp, qx, qy, qz, max_d2, best, best_idx);
  d2_out[i] = best;
  idx_out[i] = best_idx;
  if (!ok) pending[atomicAdd(n_pending, 1u)] = i;
}

// far outliers: one warp per pending query scans the whole target (exact, same metric and tie rule)
__global__ void __launch_bounds__(256) k_nn_bruteforce(NnView g, const float4* __restrict__ queries, const int* __restrict__ pending, const unsigned int* __restrict__ n_pending,
                                                       float* __restrict__ d2_out, int* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int np = (int)*n_pending;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < np; w += warps) {
    const int i = pending[w];
    const float4 q = queries[i];
    float best = 3.402823466e+38f;
    int best_idx = 0x7FFFFFFF;
    for (int j = lane; j < g.n; j += 32) {
      const float4 p = __ldg(g.pts + j);
      const float d = l2_simple(q.x, q.y, q.z, p.x, p.y, p.z);
      const int idx = __float_as_int(p.w);
      if (d < best || (d == best && idx < best_idx)) { best = d; best_idx = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
      if (ob < best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
    if (lane == 0) { d2_out[i] = best; idx_out[i] = best_idx == 0x7FFFFFFF ? -1 : best_idx; }
  }
}

// fitness reduction: sum of d2 (double) and count over d2 <= max_range (or < for the inlier test),
// fixed tree order -> deterministic.  partial[2*b] = sum, partial[2*b+1] = count
__global__ void __launch_bounds__(256) k_fitness_partial(const float* __restrict__ d2, const int* __restrict__ idx, int n, double max_range, int strict_less, double* __restrict__ partial) {
  __shared__ double s_sum[8], s_cnt[8];
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (idx[i] < 0) continue;
    const double d = (double)d2[i];
    const bool in = strict_less ? (d < max_range) : (d <= max_range);
    if (in) { sum += d; cnt += 1.0; }
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
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

struct NnGrid {
  VoxelSort sort;
  DevBuf<float4> pts, queries;
  DevBuf<uint2> table;
  DevBuf<float> d2;
  DevBuf<int> idx, pending;
  DevBuf<unsigned int> n_pending;
  DevBuf<float> T;
  uint32_t table_cap = 0;
  int n = 0;
  bool built = false;

  void release() {
    sort.release(); pts.release(); queries.release(); table.release(); d2.release(); idx.release(); pending.release(); n_pending.release(); T.release();
  }
  NnView view() const {
    NnView v;
    v.meta = sort.meta.p;
    v.table = table.p;
    v.table_mask = table_cap - 1;
    v.table_shift = 32 - (int)__builtin_ctz(table_cap);
    v.cell_start = sort.vox_start.p;
    v.pts = pts.p;
    v.n = n;
    return v;
  }
  cudaError_t build(cudaStream_t st, const float4* d_pts, int n_points) {
    cudaError_t e;
    n = n_points;
    if ((e = sort.run(st, d_pts, n, 1, kNnCell, kNnCell, kNnCell, false)) != cudaSuccess) return e;
    if ((e = pts.reserve(n > 0 ? n : 1)) != cudaSuccess) return e;
    uint32_t cap = 64;
    while (cap < (uint32_t)(2 * n + 1)) cap <<= 1;
    table_cap = cap;
    if ((e = table.reserve(cap)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(table.p, 0xFF, (size_t)cap * sizeof(uint2), st)) != cudaSuccess) return e;
    if (n > 0) {
      launch_counter() += 2;
      k_nn_reorder<<<(n + 255) / 256, 256, 0, st>>>(d_pts, n, sort.vals_a.p, sort.vals_b.p, sort.meta.p, pts.p);
      k_nn_insert<<<kNumSM * 2, 256, 0, st>>>(sort.meta.p, sort.vox_key.p, table.p, cap - 1, 32 - (int)__builtin_ctz(cap));
    }
    built = true;
    return cudaGetLastError();
  }
  // d2 / idx of every source point under T (nullptr = identity), on the stream
  cudaError_t search(cudaStream_t st, const float4* d_src, int n_src, const float* T_colmajor_host, float max_d2) {
    cudaError_t e;
    if ((e = d2.reserve(n_src > 0 ? n_src : 1)) != cudaSuccess) return e;
    if ((e = idx.reserve(n_src > 0 ? n_src : 1)) != cudaSuccess) return e;
    if ((e = pending.reserve(n_src > 0 ? n_src : 1)) != cudaSuccess) return e;
    if ((e = queries.reserve(n_src > 0 ? n_src : 1)) != cudaSuccess) return e;
    if ((e = n_pending.reserve(1)) != cudaSuccess) return e;
    if ((e = T.reserve(16)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(n_pending.p, 0, sizeof(unsigned int), st)) != cudaSuccess) return e;
    if (T_colmajor_host && (e = cudaMemcpyAsync(T.p, T_colmajor_host, 64, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if (n_src > 0) {
      launch_counter() += 2;
      k_nn_search<<<(n_src + 255) / 256, 256, 0, st>>>(view(), d_src, n_src, T.p, T_colmajor_host ? 1 : 0, max_d2, d2.p, idx.p, queries.p, pending.p, n_pending.p);
      k_nn_bruteforce<<<kNumSM * 2, 256, 0, st>>>(view(), queries.p, pending.p, n_pending.p, d2.p, idx.p);
    }
    return cudaGetLastError();
  }
};

constexpr int kFitBlocks = kNumSM;

// getFitnessScore on the device: NN search + thresholded sum.  Synchronises the stream once.
inline cudaError_t nn_fitness(cudaStream_t st, NnGrid& nn, const float4* d_src, int n_src, const float* T_colmajor, double max_range, bool strict_less, DevBuf<double>& partials,
                              PinnedBuf<unsigned char>& pin, double* sum, long long* cnt) {
  cudaError_t e;
  // the NN search itself can stop at max_range (squared distance compared un-squared, as upstream)
  const float max_d2 = max_range >= 3.0e38 ? 3.402823466e+38f : (float)max_range * 1.0001f + 1e-6f;
  // T_colmajor lives on the caller's stack: stage it through pinned memory so the async copy is safe
  if ((e = pin.reserve(2 * kFitBlocks * sizeof(double) + 64)) != cudaSuccess) return e;
  float* pinT = reinterpret_cast<float*>(pin.p + 2 * kFitBlocks * sizeof(double));
  for (int i = 0; i < 16; ++i) pinT[i] = T_colmajor[i];
  if ((e = nn.search(st, d_src, n_src, pinT, max_d2)) != cudaSuccess) return e;
  if ((e = partials.reserve(2 * kFitBlocks)) != cudaSuccess) return e;
  launch_counter() += 1;
  k_fitness_partial<<<kFitBlocks, 256, 0, st>>>(nn.d2.p, nn.idx.p, n_src, max_range, strict_less ? 1 : 0, partials.p);
  double* hp = reinterpret_cast<double*>(pin.p);
  if ((e = cudaMemcpyAsync(hp, partials.p, 2 * kFitBlocks * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  double s = 0.0, c = 0.0;
  for (int b = 0; b < kFitBlocks; ++b) { s += hp[2 * b]; c += hp[2 * b + 1]; }
  *sum = s;
  *cnt = (long long)c;
  return cudaSuccess;
}

}  // namespace b200

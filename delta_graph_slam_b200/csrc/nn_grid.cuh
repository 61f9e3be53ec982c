// b200reg — exact nearest-neighbour search on a cloud: the B200 replacement of the FLANN kd-tree
// that pcl::Registration::initCompute builds and getFitnessScore queries,
// [REF include/hdl_graph_slam/loop_detector.hpp:148; apps/scan_matching_odometry_nodelet.cpp:318-332;
//  src/hdl_graph_slam/information_matrix_calculator.cpp:77-108], and of fast_gicp's source / target
// kd-trees (1-NN correspondences, k-NN covariances).
//
// Structure: a uniform 0.5 m cell grid over the cloud.  Points are re-ordered by cell with the same
// key / radix-sort / segmentation machinery as VoxelGrid (w of each re-ordered float4 carries the
// original index) and an open-addressing hash maps cell -> (first point, end point), so a probe is
// ONE 16-byte load.  Results are EXACT, with FLANN's float metric ((dx*dx)+dy*dy)+dz*dz and ties
// broken by lowest index (the oracle's convention).
//
// A query runs in up to three phases so that the rare far queries cannot stretch the kernel:
//   near  : one thread per query walks the Chebyshev rings 0 and 1 (27 cells, pruned by box
//           distance); ~98 % of the queries of an aligned scan pair finish here
//   far   : one WARP per remaining query: the 32 lanes take the cells of ring 2, 3, ... in
//           parallel, merge their best after every ring and stop as soon as nothing outside the
//           examined block can win, lies beyond max_d2, or the block covers the occupied lattice
//   brute : (unbounded range only) queries still open after kFarRing rings are finished by a
//           CTA-per-query scan of the whole cloud, with several loads in flight per thread
#pragma once
#include "voxel_sort.cuh"

namespace b200 {

constexpr int kNearRing = 1;   // rings walked by the thread-per-query phase
#ifndef B200_NN_CELL
#define B200_NN_CELL 0.5f
#endif
#ifndef B200_NN_FAR_RING
#define B200_NN_FAR_RING 15
#endif
constexpr int kFarRing = B200_NN_FAR_RING;  // rings walked by the warp-per-query phase before the brute-force pass (7.5 m; rows of the occupancy bitmap, so a ring is cheap)
constexpr float kNnCell = B200_NN_CELL;

struct NnView {
  const SortMeta* meta;
  const uint4* table;  // (cell key, first point, end point, -)
  uint32_t table_mask;
  int table_shift;
  const float4* pts;  // re-ordered by cell; w = original index (int bits)
  int n;
  // occupancy bitmap over the cell lattice, bit index = cell key (x runs fastest, so one 32-bit word
  // covers 32 consecutive cells of a row): word 0 is 1 when the bitmap is valid (the lattice fits
  // kOccMaxCells), the bits start at word kOccHeader.  The far searches walk whole rows of a ring with
  // two word loads and probe the hash only where a bit is set; nullptr = no bitmap.
  const uint32_t* occ;
};
constexpr uint32_t kOccMaxCells = 32u << 20;  // 4 MB of bits: a 400 x 400 x 200 lattice of 0.5 m cells
constexpr int kOccHeader = 4;                 // words in front of the bits (keeps them 16-byte aligned)
constexpr size_t kOccWords = kOccMaxCells / 32 + kOccHeader + 4;

__device__ __forceinline__ bool nn_occ_valid(const NnView& g) { return g.occ != nullptr && __ldg(g.occ) != 0u; }
// bits [key0, key0 + nbits) of the bitmap, nbits <= 32 (the buffer is padded for the second word)
__device__ __forceinline__ uint32_t nn_occ_bits(const NnView& g, uint32_t key0, int nbits) {
  const uint32_t w = key0 >> 5, sh = key0 & 31u;
  const uint32_t lo = __ldg(g.occ + kOccHeader + w), hi = __ldg(g.occ + kOccHeader + 1 + w);
  const uint32_t v = __funnelshift_r(lo, hi, sh);
  return nbits >= 32 ? v : (v & ((1u << nbits) - 1u));
}
__device__ __forceinline__ bool nn_occ_bit(const NnView& g, uint32_t key) { return (__ldg(g.occ + kOccHeader + (key >> 5)) >> (key & 31u)) & 1u; }

// (first, end) of the cell's run, or first == end when the cell is empty
__device__ __forceinline__ uint2 nn_lookup(const NnView& g, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> g.table_shift;
  while (true) {
    const uint4 e = __ldg(g.table + h);
    if (e.x == key) return make_uint2(e.y, e.z);
    if (e.x == kInvalidKey) return make_uint2(0u, 0u);
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

// zero the words of the occupancy bitmap the lattice needs (its size is only known on the device)
// and mark it valid; lattices above kOccMaxCells go without (the searches fall back to hash probes)
__global__ void __launch_bounds__(256) k_nn_occ_clear(const SortMeta* __restrict__ meta, uint32_t* __restrict__ occ) {
  const GridParams& gp = meta->grid;
  const unsigned long long cells = (unsigned long long)gp.div_b[0] * (unsigned long long)gp.div_b[1] * (unsigned long long)gp.div_b[2];
  const bool ok = gp.any && !gp.overflow && cells <= (unsigned long long)kOccMaxCells;
  if (blockIdx.x == 0 && threadIdx.x == 0) occ[0] = ok ? 1u : 0u;
  if (!ok) return;
  const uint32_t words = (uint32_t)(cells / 32) + 3;
  uint4* o4 = reinterpret_cast<uint4*>(occ + kOccHeader);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (words + 3) / 4; i += gridDim.x * blockDim.x) o4[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256) k_nn_insert(const SortMeta* __restrict__ meta, const uint32_t* __restrict__ vox_key, const uint32_t* __restrict__ vox_start, uint4* __restrict__ table,
                                                   uint32_t mask, int shift, uint32_t* __restrict__ occ) {
  const int n_vox = (int)meta->n_vox;
  const bool use_occ = occ != nullptr && occ[0] != 0u;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_vox; slot += gridDim.x * blockDim.x) {
    const uint32_t key = vox_key[slot];
    if (use_occ) atomicOr(occ + kOccHeader + (key >> 5), 1u << (key & 31u));
    uint32_t h = (key * 2654435761u) >> shift;
    while (true) {
      uint32_t old = atomicCAS(&table[h].x, kInvalidKey, key);
      if (old == kInvalidKey) {
        table[h].y = vox_start[slot];
        table[h].z = vox_start[slot + 1];
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

// ---- the search --------------------------------------------------------------------------------
struct NnQuery {
  float qx, qy, qz;
  int cx, cy, cz;
  float margin;
};

__device__ __forceinline__ NnQuery nn_make_query(const GridParams& gp, float qx, float qy, float qz) {
  NnQuery q;
  q.qx = qx; q.qy = qy; q.qz = qz;
  q.cx = (int)floorf(__fmul_rn(qx, gp.inv_leaf[0]));
  q.cy = (int)floorf(__fmul_rn(qy, gp.inv_leaf[1]));
  q.cz = (int)floorf(__fmul_rn(qz, gp.inv_leaf[2]));
  q.margin = 1e-3f * gp.leaf[0] + 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz));
  return q;
}

// true when every point outside the (2r-1)^3 block of cells around the query cell is provably no
// better than `best`, lies beyond max_d2, or does not exist (the block covers the occupied lattice)
__device__ __forceinline__ bool nn_settled(const GridParams& gp, const NnQuery& q, int r, float best, float max_d2) {
  const float c = gp.leaf[0];
  const float gx = fminf(q.qx - (float)(q.cx - (r - 1)) * c, (float)(q.cx + r) * c - q.qx);
  const float gy = fminf(q.qy - (float)(q.cy - (r - 1)) * c, (float)(q.cy + r) * c - q.qy);
  const float gz = fminf(q.qz - (float)(q.cz - (r - 1)) * c, (float)(q.cz + r) * c - q.qz);
  const float gap = fmaxf(fminf(gx, fminf(gy, gz)) - q.margin, 0.f);
  const float gap2 = gap * gap;
  if (best <= gap2 || gap2 > max_d2) return true;
  return q.cx - (r - 1) <= gp.min_b[0] && q.cx + (r - 1) >= gp.max_b[0] && q.cy - (r - 1) <= gp.min_b[1] && q.cy + (r - 1) >= gp.max_b[1] && q.cz - (r - 1) <= gp.min_b[2] &&
         q.cz + (r - 1) >= gp.max_b[2];
}

// lower bound of the squared distance from the query to cell (ix, iy, iz)
__device__ __forceinline__ float nn_box_d2(const GridParams& gp, const NnQuery& q, int ix, int iy, int iz) {
  const float c = gp.leaf[0];
  const float dx = fmaxf(fmaxf((float)ix * c - q.margin - q.qx, q.qx - ((float)(ix + 1) * c + q.margin)), 0.f);
  const float dy = fmaxf(fmaxf((float)iy * c - q.margin - q.qy, q.qy - ((float)(iy + 1) * c + q.margin)), 0.f);
  const float dz = fmaxf(fmaxf((float)iz * c - q.margin - q.qz, q.qz - ((float)(iz + 1) * c + q.margin)), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ bool nn_better(float d, int idx, float best, int best_idx) { return d < best || (d == best && idx < best_idx); }

// scan one cell; best_idx = INT_MAX-style sentinel 0x7FFFFFFF while nothing was found.  DEPTH loads are
// in flight at a time: a thread-per-query scan is a chain of L2 round trips.  The fitness kernels keep
// DEPTH = 4 (measured: with 8 the 16 extra registers cost them a resident CTA per SM and the batched
// fitness went from 16.7 to 22.0 ms per 1024 pairs; in the GICP align kernel 8 was slower as well, 374 vs 318 us).
template <int DEPTH = 4>
__device__ __forceinline__ uint32_t nn_scan_cell(const NnView& g, const GridParams& gp, const NnQuery& q, int ix, int iy, int iz, float& best, int& best_idx) {
  const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
  const uint2 run = nn_lookup(g, key);
  if (DEPTH == 4) {
    // (d, index) as one 64-bit key: squared distances are >= +0, so their float bits order like the values,
    // and indices are non-negative — nn_better's "smaller d, then lower index" is an unsigned 64-bit minimum
    // (a NaN distance has the largest bits and never wins, as it never wins nn_better)
    unsigned long long bk = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(unsigned)best_idx;
#pragma unroll 4
    for (uint32_t j = run.x; j < run.y; ++j) {
      const float4 p = __ldg(g.pts + j);
      const float d = l2_simple(q.qx, q.qy, q.qz, p.x, p.y, p.z);
      const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)__float_as_uint(p.w);
      bk = k < bk ? k : bk;
    }
    best = __uint_as_float((unsigned)(bk >> 32));
    best_idx = (int)(unsigned)(bk & 0xffffffffull);
  } else {
    for (uint32_t j = run.x; j < run.y; j += DEPTH) {
      float4 p[DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) p[u] = j + u < run.y ? __ldg(g.pts + j + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) {
        if (j + u < run.y) {
          const float d = l2_simple(q.qx, q.qy, q.qz, p[u].x, p[u].y, p[u].z);
          const int idx = __float_as_int(p[u].w);
          if (nn_better(d, idx, best, best_idx)) { best = d; best_idx = idx; }
        }
      }
    }
  }
  return run.y - run.x;
}

constexpr int kNoIndex = 0x7FFFFFFF;

// near phase, one thread: rings 0..kNearRing.
//   kNnDone : resolved
//   kNnOpen : rings 0..kNearRing done, something farther out can still win -> far phase from ring kNearRing + 1
//   kNnBail : (only with -DB200_NN_NEAR_BUDGET=n) the thread has scanned more than n points without
//             settling -> the far phase takes the query over from ring 0, 32 lanes to a cell.  Built to test
//             whether single slow threads inside dense clumps hold their CTA at the block barrier;
//             measured with n = 128: exact (full GPU suite green) but no faster — batched fitness 18.9 vs
//             16.8 ms per 1024 pairs, GICP align 0.324 vs 0.289 ms — so the default is off.
// SEEDED: best / best_idx arrive holding a real point of the cloud (e.g. last iteration's correspondence),
// which only tightens the pruning — the result is the same exact nearest neighbour.
constexpr int kNnDone = 0, kNnOpen = 1, kNnBail = 2;
#ifndef B200_NN_NEAR_BUDGET
#define B200_NN_NEAR_BUDGET 0
#endif
constexpr uint32_t kNearBudget = B200_NN_NEAR_BUDGET;
constexpr uint32_t kBailFlag = 1u << 30;  // carried in the point index of a pending / queue entry

template <int DEPTH = 4, bool SEEDED = false>
__device__ __forceinline__ int nn_query_near(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, float& best, int& best_idx) {
  if (!SEEDED) {
    best = 3.402823466e+38f;
    best_idx = kNoIndex;
  }
  uint32_t scanned = 0;
  for (int r = 0; r <= kNearRing; ++r) {
    if (r >= 1 && nn_settled(gp, q, r, best, max_d2)) return kNnDone;
    const int z0 = max(q.cz - r, gp.min_b[2]), z1 = min(q.cz + r, gp.max_b[2]);
    const int y0 = max(q.cy - r, gp.min_b[1]), y1 = min(q.cy + r, gp.max_b[1]);
    for (int iz = z0; iz <= z1; ++iz) {
      const bool zface = (iz == q.cz - r) || (iz == q.cz + r);
      for (int iy = y0; iy <= y1; ++iy) {
        const bool face = zface || (iy == q.cy - r) || (iy == q.cy + r);
        const int xstep = face ? 1 : (r == 0 ? 1 : 2 * r);
        for (int ix = q.cx - r; ix <= q.cx + r; ix += xstep) {
          if (ix < gp.min_b[0] || ix > gp.max_b[0]) continue;
          if (nn_box_d2(gp, q, ix, iy, iz) > fminf(best, max_d2)) continue;  // beyond the range of interest: nothing in there can count
          if (kNearBudget != 0u && scanned > kNearBudget) return kNnBail;
          scanned += nn_scan_cell<DEPTH>(g, gp, q, ix, iy, iz, best, best_idx);
        }
      }
    }
  }
  return nn_settled(gp, q, kNearRing + 1, best, max_d2) ? kNnDone : kNnOpen;
}

// nn_query_near in two steps, for kernels that compact the queries ring 0 left open before walking ring 1 (the lanes of a
// thread-per-query warp otherwise idle while their neighbours walk on: ncu counted 10 of 32 lanes active in the batched
// fitness search).  nn_near_ring0: the query's own cell; true = settled.  nn_near_ring1: the 26 neighbours, then the
// same closing test as nn_query_near.  Same candidates, same (distance, index) order: the same exact neighbour.
template <int DEPTH = 4>
__device__ __forceinline__ bool nn_near_ring0(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, float& best, int& best_idx) {
  best = 3.402823466e+38f;
  best_idx = kNoIndex;
  if (q.cx >= gp.min_b[0] && q.cx <= gp.max_b[0] && q.cy >= gp.min_b[1] && q.cy <= gp.max_b[1] && q.cz >= gp.min_b[2] && q.cz <= gp.max_b[2] &&
      !(nn_box_d2(gp, q, q.cx, q.cy, q.cz) > fminf(best, max_d2)))
    nn_scan_cell<DEPTH>(g, gp, q, q.cx, q.cy, q.cz, best, best_idx);
  return nn_settled(gp, q, 1, best, max_d2);
}
template <int DEPTH = 4>
__device__ __forceinline__ int nn_near_ring1(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, float& best, int& best_idx) {
  static_assert(kNearRing == 1, "ring 1 is the last ring of the near phase");
  const int z0 = max(q.cz - 1, gp.min_b[2]), z1 = min(q.cz + 1, gp.max_b[2]);
  const int y0 = max(q.cy - 1, gp.min_b[1]), y1 = min(q.cy + 1, gp.max_b[1]);
  for (int iz = z0; iz <= z1; ++iz) {
    const bool zface = (iz == q.cz - 1) || (iz == q.cz + 1);
    for (int iy = y0; iy <= y1; ++iy) {
      const bool face = zface || (iy == q.cy - 1) || (iy == q.cy + 1);
      const int xstep = face ? 1 : 2;
      for (int ix = q.cx - 1; ix <= q.cx + 1; ix += xstep) {
        if (ix < gp.min_b[0] || ix > gp.max_b[0]) continue;
        if (nn_box_d2(gp, q, ix, iy, iz) > fminf(best, max_d2)) continue;
        nn_scan_cell<DEPTH>(g, gp, q, ix, iy, iz, best, best_idx);
      }
    }
  }
  return nn_settled(gp, q, kNearRing + 1, best, max_d2) ? kNnDone : kNnOpen;
}

// The near phase for ONE query by a group of LANES adjacent lanes (2, 4 or 8; all 32 lanes of the warp call it, `sub` =
// lane % LANES, `active` false for a group without a query).  A thread-per-query walk of rings 0..1 is a chain of
// dependent L2 round trips — hash slot, then the cell's points, cell after cell — and a pass of k_gicp_align over a
// 15 k-point scan gives each thread of the grid at most ONE query: the pass lasted as long as that chain while four
// fifths of the warps had nothing to do.  Here the group scans the query's own cell LANES points at a time and deals the
// 26 neighbours over its lanes; candidates are compared as (distance bits, index) keys, whose minimum does not depend on
// the order of the scan, so the result is nn_query_near's exact neighbour (same tie rule).  Group-uniform on return.
template <int LANES, bool SEEDED = false>
__device__ __forceinline__ int nn_query_near_group(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, int sub, bool active, float& best, int& best_idx) {
  static_assert(kNearRing == 1, "rings 0 and 1 are written out");
  if (!SEEDED) {
    best = 3.402823466e+38f;
    best_idx = kNoIndex;
  }
  unsigned long long bk = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(unsigned)best_idx;
  auto merge = [&]() {
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, bk, o);
      bk = other < bk ? other : bk;
    }
    best = __uint_as_float((unsigned)(bk >> 32));
    best_idx = (int)(unsigned)(bk & 0xffffffffull);
  };
  auto inside = [&](int ix, int iy, int iz) { return ix >= gp.min_b[0] && ix <= gp.max_b[0] && iy >= gp.min_b[1] && iy <= gp.max_b[1] && iz >= gp.min_b[2] && iz <= gp.max_b[2]; };
  auto offer = [&](uint32_t j) {
    const float4 p = __ldg(g.pts + j);
    const float d = l2_simple(q.qx, q.qy, q.qz, p.x, p.y, p.z);
    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)__float_as_uint(p.w);
    bk = k < bk ? k : bk;
  };
  // ring 0: the query's own cell, LANES points at a time
  if (active && inside(q.cx, q.cy, q.cz) && !(nn_box_d2(gp, q, q.cx, q.cy, q.cz) > fminf(best, max_d2))) {
    const uint32_t key = (uint32_t)((q.cx - gp.min_b[0]) * gp.mul[0] + (q.cy - gp.min_b[1]) * gp.mul[1] + (q.cz - gp.min_b[2]) * gp.mul[2]);
    const uint2 run = nn_lookup(g, key);
#pragma unroll 2
    for (uint32_t j = run.x + (uint32_t)sub; j < run.y; j += LANES) offer(j);
  }
  merge();
  const bool done = !active || nn_settled(gp, q, 1, best, max_d2);
  // ring 1: the 26 neighbours dealt over the lanes, each pruned against what the lane knows so far
  if (!done) {
    for (int c = sub; c < 27; c += LANES) {
      if (c == 13) continue;
      const int ix = q.cx + c % 3 - 1, iy = q.cy + (c / 3) % 3 - 1, iz = q.cz + c / 9 - 1;
      if (!inside(ix, iy, iz)) continue;
      if (nn_box_d2(gp, q, ix, iy, iz) > fminf(__uint_as_float((unsigned)(bk >> 32)), max_d2)) continue;
      const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
      const uint2 run = nn_lookup(g, key);
#pragma unroll 4
      for (uint32_t j = run.x; j < run.y; ++j) offer(j);
    }
  }
  merge();
  if (done) return kNnDone;
  return nn_settled(gp, q, kNearRing + 1, best, max_d2) ? kNnDone : kNnOpen;
}

__device__ __forceinline__ void nn_warp_merge(float& best, int& best_idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
    if (nn_better(ob, oi, best, best_idx)) { best = ob; best_idx = oi; }
  }
}

// one warp scans the runs its lanes hold (empty when run.x == run.y) for ONE query, flattened: the
// points of all runs are numbered consecutively and batch b takes points 32 b .. 32 b + 31 whatever
// run they belong to; the load of the next batch is issued before the current one is compared
__device__ __forceinline__ void nn_scan_runs_warp(const NnView& g, uint2 run, const NnQuery& q, int lane, float& best, int& best_idx) {
  const uint32_t len = run.y - run.x;
  uint32_t incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0u) return;
  const uint32_t first = run.x - (incl - len);
  auto fetch = [&](uint32_t t0, float4& p) -> bool {
    const uint32_t t = t0 + lane;
    int o = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const uint32_t v = __shfl_sync(0xffffffffu, incl, o + step - 1);
      if (v <= t) o += step;
    }
    const uint32_t f = __shfl_sync(0xffffffffu, first, o & 31);
    const bool ok = t < total;
    p = ok ? __ldg(g.pts + (f + t)) : make_float4(0.f, 0.f, 0.f, 0.f);
    return ok;
  };
  float4 p_cur, p_next = make_float4(0.f, 0.f, 0.f, 0.f);
  bool ok_cur = fetch(0u, p_cur), ok_next = false;
  for (uint32_t t0 = 0; t0 < total; t0 += 32) {
    if (t0 + 32 < total) ok_next = fetch(t0 + 32, p_next);
    if (ok_cur) {
      const float d = l2_simple(q.qx, q.qy, q.qz, p_cur.x, p_cur.y, p_cur.z);
      const int idx = __float_as_int(p_cur.w);
      if (nn_better(d, idx, best, best_idx)) { best = d; best_idx = idx; }
    }
    p_cur = p_next;
    ok_cur = ok_next;
    ok_next = false;
  }
}

// rings 0 and 1 for one query by a whole warp (a query the near phase gave up on): the query's own cell
// is scanned 32 points at a time, then the 26 neighbours are pruned against that result and what
// survives is scanned flattened
__device__ __forceinline__ void nn_rings01_warp(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, int lane, float& best, int& best_idx) {
  const bool use_occ = nn_occ_valid(g);
  auto cell_run = [&](int ix, int iy, int iz) -> uint2 {
    if (ix < gp.min_b[0] || ix > gp.max_b[0] || iy < gp.min_b[1] || iy > gp.max_b[1] || iz < gp.min_b[2] || iz > gp.max_b[2]) return make_uint2(0u, 0u);
    if (nn_box_d2(gp, q, ix, iy, iz) > fminf(best, max_d2)) return make_uint2(0u, 0u);
    const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
    if (use_occ && !nn_occ_bit(g, key)) return make_uint2(0u, 0u);
    return nn_lookup(g, key);
  };
  uint2 run = make_uint2(0u, 0u);
  if (lane == 0) run = cell_run(q.cx, q.cy, q.cz);
  nn_scan_runs_warp(g, run, q, lane, best, best_idx);
  nn_warp_merge(best, best_idx);
  run = make_uint2(0u, 0u);
  if (lane < 27 && lane != 13) run = cell_run(q.cx + lane % 3 - 1, q.cy + (lane / 3) % 3 - 1, q.cz + lane / 9 - 1);
  nn_scan_runs_warp(g, run, q, lane, best, best_idx);
  nn_warp_merge(best, best_idx);
}

// far phase, one warp (all 32 lanes call with the same query and the near phase's best): rings
// kNearRing+1 .. max_ring, preceded by rings 0..kNearRing when the near phase bailed out (from_ring0).
// Returns true when resolved; best / best_idx are warp-uniform on return.
__device__ __forceinline__ bool nn_query_far_warp(const NnView& g, const GridParams& gp, const NnQuery& q, float max_d2, int max_ring, int lane, float& best, int& best_idx,
                                                  bool from_ring0 = false) {
  const bool use_occ = nn_occ_valid(g);  // warp-uniform
  if (from_ring0) nn_rings01_warp(g, gp, q, max_d2, lane, best, best_idx);
  for (int r = kNearRing + 1; r <= max_ring; ++r) {
    if (nn_settled(gp, q, r, best, max_d2)) return true;
    const int side = 2 * r + 1, inner = 2 * r - 1;
    const int x0 = max(q.cx - r, gp.min_b[0]), x1 = min(q.cx + r, gp.max_b[0]);
    if (use_occ && side <= 32) {
      // the shell of ring r row by row: the (2r+1)^2 rows along x are dealt to the lanes; a row costs
      // two word loads of the occupancy bitmap, and only occupied cells of the shell (every cell of a
      // face row, the two end cells of an interior row) are probed and scanned
      if (x0 <= x1) {
        const int nb = x1 - x0 + 1;
        uint32_t ends = 0u;
        if (q.cx - r >= gp.min_b[0]) ends |= 1u;
        if (q.cx + r <= gp.max_b[0]) ends |= 1u << (nb - 1);
        const int xq = min(max(q.cx, gp.min_b[0]), gp.max_b[0]);
        for (int row = lane; row < side * side; row += 32) {
          const int dz = row / side - r, dy = row - (dz + r) * side - r;
          const int iy = q.cy + dy, iz = q.cz + dz;
          if (iy < gp.min_b[1] || iy > gp.max_b[1] || iz < gp.min_b[2] || iz > gp.max_b[2]) continue;
          if (nn_box_d2(gp, q, xq, iy, iz) > fminf(best, max_d2)) continue;  // no cell of this row can win or lies in range
          const bool face = dy == -r || dy == r || dz == -r || dz == r;
          const uint32_t key0 = (uint32_t)((x0 - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
          uint32_t bits = nn_occ_bits(g, key0, nb);
          if (!face) bits &= ends;
          while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int ix = x0 + b;
            if (nn_box_d2(gp, q, ix, iy, iz) > fminf(best, max_d2)) continue;
            nn_scan_cell(g, gp, q, ix, iy, iz, best, best_idx);
          }
        }
      }
    } else {
      // the shell of ring r: two z faces (side^2), two y faces (side x inner), two x faces (inner^2)
      const int nz = 2 * side * side, ny = 2 * side * inner, total = nz + ny + 2 * inner * inner;
      for (int c = lane; c < total; c += 32) {
        int dx, dy, dz;
        if (c < nz) {
          const int f = c / (side * side), rem = c - f * side * side;
          dz = f ? r : -r; dx = rem % side - r; dy = rem / side - r;
        } else if (c < nz + ny) {
          const int cc = c - nz, f = cc / (side * inner), rem = cc - f * side * inner;
          dy = f ? r : -r; dx = rem % side - r; dz = rem / side - (r - 1);
        } else {
          const int cc = c - nz - ny, f = cc / (inner * inner), rem = cc - f * inner * inner;
          dx = f ? r : -r; dy = rem % inner - (r - 1); dz = rem / inner - (r - 1);
        }
        const int ix = q.cx + dx, iy = q.cy + dy, iz = q.cz + dz;
        if (ix < gp.min_b[0] || ix > gp.max_b[0] || iy < gp.min_b[1] || iy > gp.max_b[1] || iz < gp.min_b[2] || iz > gp.max_b[2]) continue;
        if (nn_box_d2(gp, q, ix, iy, iz) > fminf(best, max_d2)) continue;
        nn_scan_cell(g, gp, q, ix, iy, iz, best, best_idx);
      }
    }
    nn_warp_merge(best, best_idx);
  }
  return nn_settled(gp, q, max_ring + 1, best, max_d2);
}

// brute phase, one warp: the whole cloud, four loads in flight per lane
__device__ __forceinline__ void nn_query_brute_warp(const NnView& g, float qx, float qy, float qz, int lane, float& best, int& best_idx) {
  best = 3.402823466e+38f;
  best_idx = kNoIndex;
  int j = lane;
  for (; j + 96 < g.n; j += 128) {
    const float4 p0 = __ldg(g.pts + j), p1 = __ldg(g.pts + j + 32), p2 = __ldg(g.pts + j + 64), p3 = __ldg(g.pts + j + 96);
    const float d0 = l2_simple(qx, qy, qz, p0.x, p0.y, p0.z), d1 = l2_simple(qx, qy, qz, p1.x, p1.y, p1.z);
    const float d2 = l2_simple(qx, qy, qz, p2.x, p2.y, p2.z), d3 = l2_simple(qx, qy, qz, p3.x, p3.y, p3.z);
    if (nn_better(d0, __float_as_int(p0.w), best, best_idx)) { best = d0; best_idx = __float_as_int(p0.w); }
    if (nn_better(d1, __float_as_int(p1.w), best, best_idx)) { best = d1; best_idx = __float_as_int(p1.w); }
    if (nn_better(d2, __float_as_int(p2.w), best, best_idx)) { best = d2; best_idx = __float_as_int(p2.w); }
    if (nn_better(d3, __float_as_int(p3.w), best, best_idx)) { best = d3; best_idx = __float_as_int(p3.w); }
  }
  for (; j < g.n; j += 32) {
    const float4 p = __ldg(g.pts + j);
    const float d = l2_simple(qx, qy, qz, p.x, p.y, p.z);
    if (nn_better(d, __float_as_int(p.w), best, best_idx)) { best = d; best_idx = __float_as_int(p.w); }
  }
  nn_warp_merge(best, best_idx);
}

// ---- single-pair kernels (getFitnessScore / inlier fraction on one handle) ----------------------
// near: one thread per source point: transform by T (column-major 4x4, float, pcl::transformPoint
// order) and search.  d2_out[i] = squared NN distance, idx_out[i] = target index or -1; unresolved
// queries are appended to `pending` with their best-so-far kept in d2_out / idx_out.
__global__ void __launch_bounds__(256) k_nn_search(NnView g, const float4* __restrict__ src, int n_src, const float* __restrict__ T16, int use_T, float max_d2,
                                                   float* __restrict__ d2_out, int* __restrict__ idx_out, float4* __restrict__ q_out, int* __restrict__ pending,
                                                   unsigned int* __restrict__ n_pending) {
  __shared__ float T[16];
  __shared__ GridParams s_gp;  // the target's cell lattice, shared by the CTA instead of 20 registers per thread
  if (threadIdx.x < 16) T[threadIdx.x] = use_T ? T16[threadIdx.x] : ((threadIdx.x % 5 == 0) ? 1.f : 0.f);
  if (threadIdx.x >= 32 && threadIdx.x < 32 + (int)(sizeof(GridParams) / 4))
    reinterpret_cast<uint32_t*>(&s_gp)[threadIdx.x - 32] = reinterpret_cast<const uint32_t*>(&g.meta->grid)[threadIdx.x - 32];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_src) return;
  const GridParams& gp = s_gp;
  const float4 p = __ldg(src + i);
  float qx = p.x, qy = p.y, qz = p.z;
  if (use_T) {
    qx = affine_row(T[0], T[4], T[8], T[12], p.x, p.y, p.z);
    qy = affine_row(T[1], T[5], T[9], T[13], p.x, p.y, p.z);
    qz = affine_row(T[2], T[6], T[10], T[14], p.x, p.y, p.z);
  }
  q_out[i] = make_float4(qx, qy, qz, 1.f);
  float best = 3.402823466e+38f;
  int best_idx = kNoIndex;
  int st = kNnDone;
  if (g.n > 0 && gp.any && !gp.overflow) st = nn_query_near(g, gp, nn_make_query(gp, qx, qy, qz), max_d2, best, best_idx);
  d2_out[i] = best;
  idx_out[i] = best_idx == kNoIndex ? -1 : best_idx;
  if (st != kNnDone) pending[atomicAdd(n_pending, 1u)] = i | (st == kNnBail ? (int)kBailFlag : 0);
}

// far: one warp per pending query; what is still open afterwards moves to pending2
__global__ void __launch_bounds__(256) k_nn_far(NnView g, const float4* __restrict__ queries, const int* __restrict__ pending, const unsigned int* __restrict__ n_pending, float max_d2,
                                                float* __restrict__ d2_out, int* __restrict__ idx_out, int* __restrict__ pending2, unsigned int* __restrict__ n_pending2) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int np = (int)*n_pending;
  const GridParams gp = g.meta->grid;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < np; w += warps) {
    const int pe = pending[w];
    const int i = pe & (int)~kBailFlag;
    const float4 q = queries[i];
    float best = d2_out[i];
    int best_idx = idx_out[i] < 0 ? kNoIndex : idx_out[i];
    const bool ok = nn_query_far_warp(g, gp, nn_make_query(gp, q.x, q.y, q.z), max_d2, kFarRing, lane, best, best_idx, (pe & (int)kBailFlag) != 0);
    __syncwarp();
    if (lane == 0) {
      d2_out[i] = best;
      idx_out[i] = best_idx == kNoIndex ? -1 : best_idx;
      if (!ok) pending2[atomicAdd(n_pending2, 1u)] = i;
    }
  }
}

// block-wide version for the stand-alone brute-force kernels: 256 threads scan the cloud, four loads
// in flight per thread; the result is valid in thread 0
__device__ __forceinline__ void nn_query_brute_block(const NnView& g, float qx, float qy, float qz, float& best, int& best_idx) {
  __shared__ float s_b[8];
  __shared__ int s_i[8];
  best = 3.402823466e+38f;
  best_idx = kNoIndex;
  int j = threadIdx.x;
  for (; j + 768 < g.n; j += 1024) {
    const float4 p0 = __ldg(g.pts + j), p1 = __ldg(g.pts + j + 256), p2 = __ldg(g.pts + j + 512), p3 = __ldg(g.pts + j + 768);
    const float d0 = l2_simple(qx, qy, qz, p0.x, p0.y, p0.z), d1 = l2_simple(qx, qy, qz, p1.x, p1.y, p1.z);
    const float d2 = l2_simple(qx, qy, qz, p2.x, p2.y, p2.z), d3 = l2_simple(qx, qy, qz, p3.x, p3.y, p3.z);
    if (nn_better(d0, __float_as_int(p0.w), best, best_idx)) { best = d0; best_idx = __float_as_int(p0.w); }
    if (nn_better(d1, __float_as_int(p1.w), best, best_idx)) { best = d1; best_idx = __float_as_int(p1.w); }
    if (nn_better(d2, __float_as_int(p2.w), best, best_idx)) { best = d2; best_idx = __float_as_int(p2.w); }
    if (nn_better(d3, __float_as_int(p3.w), best, best_idx)) { best = d3; best_idx = __float_as_int(p3.w); }
  }
  for (; j < g.n; j += 256) {
    const float4 p = __ldg(g.pts + j);
    const float d = l2_simple(qx, qy, qz, p.x, p.y, p.z);
    if (nn_better(d, __float_as_int(p.w), best, best_idx)) { best = d; best_idx = __float_as_int(p.w); }
  }
  nn_warp_merge(best, best_idx);
  __syncthreads();  // s_b / s_i may still be read by the previous query's thread 0
  if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = best_idx; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int w = 1; w < 8; ++w)
      if (nn_better(s_b[w], s_i[w], best, best_idx)) { best = s_b[w]; best_idx = s_i[w]; }
}

// brute: one CTA per still-open query scans the whole target (exact, same metric and tie rule)
__global__ void __launch_bounds__(256) k_nn_bruteforce(NnView g, const float4* __restrict__ queries, const int* __restrict__ pending, const unsigned int* __restrict__ n_pending,
                                                       float* __restrict__ d2_out, int* __restrict__ idx_out) {
  const int np = (int)*n_pending;
  for (int w = blockIdx.x; w < np; w += gridDim.x) {
    const int i = pending[w];
    const float4 q = queries[i];
    float best;
    int best_idx;
    nn_query_brute_block(g, q.x, q.y, q.z, best, best_idx);
    if (threadIdx.x == 0) { d2_out[i] = best; idx_out[i] = best_idx == kNoIndex ? -1 : best_idx; }
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

// ---- pcl::RadiusOutlierRemoval [REF apps/prefiltering_nodelet.cpp:88-96,262-273] ----------------
// keep[i] = 1 when more than min_neighbors points of the same cloud (the point itself included) lie
// strictly inside the radius.  One thread per point in INPUT order over the cells the radius can
// reach, pruned by box distance and the occupancy bitmap, stopping as soon as the count is reached.
// block_count[b] = points kept by block b (for the order-preserving scatter).
__global__ void __launch_bounds__(256) k_ror_flags(NnView g, const float4* __restrict__ pts, int n, float r2, int rings, int min_neighbors, unsigned char* __restrict__ keep,
                                                   uint32_t* __restrict__ block_count) {
  __shared__ uint32_t s_cnt[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int k = 0;
  if (i < n) {
    const float4 p = __ldg(pts + i);
    const GridParams gp = g.meta->grid;
    if (finite3(p.x, p.y, p.z) && gp.any && !gp.overflow) {
      const bool use_occ = nn_occ_valid(g);
      const NnQuery q = nn_make_query(gp, p.x, p.y, p.z);
      int count = 0;
      for (int dz = -rings; dz <= rings && count <= min_neighbors; ++dz)
        for (int dy = -rings; dy <= rings && count <= min_neighbors; ++dy)
          for (int dx = -rings; dx <= rings && count <= min_neighbors; ++dx) {
            const int ix = q.cx + dx, iy = q.cy + dy, iz = q.cz + dz;
            if (ix < gp.min_b[0] || ix > gp.max_b[0] || iy < gp.min_b[1] || iy > gp.max_b[1] || iz < gp.min_b[2] || iz > gp.max_b[2]) continue;
            if (!(nn_box_d2(gp, q, ix, iy, iz) < r2)) continue;
            const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
            if (use_occ && !nn_occ_bit(g, key)) continue;
            const uint2 run = nn_lookup(g, key);
            for (uint32_t j = run.x; j < run.y && count <= min_neighbors; ++j) {
              const float4 t = __ldg(g.pts + j);
              if (l2_simple(p.x, p.y, p.z, t.x, t.y, t.z) < r2) ++count;
            }
          }
      k = count > min_neighbors ? 1 : 0;
    }
    keep[i] = (unsigned char)k;
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, k);
  if (lane == 0) s_cnt[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t c = 0;
    for (int w = 0; w < 8; ++w) c += s_cnt[w];
    block_count[blockIdx.x] = c;
  }
}

struct RorCounts { uint32_t n_out; uint32_t overflow; };  // overflow: the cloud spans more 0.5 m cells than the lattice can index (no result)
// order-preserving scatter: every block sums the counts of the blocks before it (a few hundred words
// out of L2), scans its own flags, writes; the last block to finish publishes the total
__global__ void __launch_bounds__(256) k_ror_scatter(const float4* __restrict__ pts, int n, const unsigned char* __restrict__ keep, const uint32_t* __restrict__ block_count,
                                                     float4* __restrict__ out, float4* host_out, unsigned host_cap, RorCounts* __restrict__ counts, RorCounts* host_counts,
                                                     unsigned int* host_flag, unsigned int host_seq, unsigned int* done_blocks, const SortMeta* __restrict__ meta, int flatten) {
  __shared__ uint32_t s_part[8];
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t before = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += 256) before += __ldcg(block_count + b);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_part[warp] = before;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    s_base = t;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = i < n ? (int)keep[i] : 0;
  const uint32_t bal = __ballot_sync(0xffffffffu, k);
  __syncthreads();  // s_part is reused
  if (lane == 0) s_part[warp] = __popc(bal);
  __syncthreads();
  uint32_t off = s_base;
  for (int w = 0; w < warp; ++w) off += s_part[w];
  if (k) {
    const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
    float4 p = __ldg(pts + i);
    if (flatten) p.z = 0.f;  // PrefilteringNodelet::flatten [REF apps/prefiltering_nodelet.cpp:166-183]
    out[dst] = p;
    if (host_out && dst < host_cap) host_out[dst] = p;
  }
  if (host_out) __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(done_blocks, 1u) != gridDim.x - 1) return;
  *done_blocks = 0u;
  uint32_t total = 0;
  for (int b = 0; b < (int)gridDim.x; ++b) total += __ldcg(block_count + b);
  const uint32_t overflow = (meta && meta->grid.overflow) ? 1u : 0u;  // meta == nullptr: no search lattice behind this call (distance filter)
  counts->n_out = total;
  counts->overflow = overflow;
  if (host_counts) {
    host_counts->n_out = total;
    host_counts->overflow = overflow;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(host_flag) = host_seq;
  }
}

struct NnGrid {
  VoxelSort sort;
  DevBuf<float4> pts, queries;
  DevBuf<uint4> table;
  DevBuf<uint32_t> occ;
  DevBuf<float> d2;
  DevBuf<int> idx, pending, pending2;
  DevBuf<unsigned int> n_pending;  // [0] after the near phase, [1] after the far phase
  DevBuf<float> T;
  uint32_t table_cap = 0;
  int n = 0;
  bool built = false;

  void release() {
    sort.release(); pts.release(); queries.release(); table.release(); occ.release(); d2.release(); idx.release(); pending.release(); pending2.release(); n_pending.release(); T.release();
  }
  NnView view() const {
    NnView v;
    v.meta = sort.meta.p;
    v.table = table.p;
    v.table_mask = table_cap - 1;
    v.table_shift = 32 - (int)__builtin_ctz(table_cap);
    v.pts = pts.p;
    v.n = n;
    v.occ = occ.p;
    return v;
  }
  static uint32_t capacity_for(int n_points) {
    uint32_t cap = 64;
    while (cap < (uint32_t)(2 * n_points + 1)) cap <<= 1;
    return cap;
  }
  cudaError_t build(cudaStream_t st, const float4* d_pts, int n_points, int is_dense = 1, PointGate gate = kNoGate) {
    cudaError_t e;
    n = n_points;
    if ((e = sort.run(st, d_pts, n, is_dense, kNnCell, kNnCell, kNnCell, false, gate)) != cudaSuccess) return e;
    if ((e = pts.reserve(n > 0 ? n : 1)) != cudaSuccess) return e;
    table_cap = capacity_for(n);
    if ((e = table.reserve(table_cap)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(table.p, 0xFF, (size_t)table_cap * sizeof(uint4), st)) != cudaSuccess) return e;
    if (!occ.p) {
      if ((e = occ.reserve(kOccWords)) != cudaSuccess) return e;
      if ((e = cudaMemsetAsync(occ.p, 0, 16, st)) != cudaSuccess) return e;  // header: no bitmap until k_nn_occ_clear says so
    }
    if (n > 0) {
      launch_counter() += 3;
      k_nn_occ_clear<<<kNumSM, 256, 0, st>>>(sort.meta.p, occ.p);
      k_nn_reorder<<<(n + 255) / 256, 256, 0, st>>>(d_pts, n, sort.vals_a.p, sort.vals_b.p, sort.meta.p, pts.p);
      k_nn_insert<<<kNumSM * 2, 256, 0, st>>>(sort.meta.p, sort.vox_key.p, sort.vox_start.p, table.p, table_cap - 1, 32 - (int)__builtin_ctz(table_cap), occ.p);
    }
    built = true;
    return cudaGetLastError();
  }
  // d2 / idx of every source point under T (nullptr = identity), on the stream
  cudaError_t search(cudaStream_t st, const float4* d_src, int n_src, const float* T_colmajor_host, float max_d2) {
    cudaError_t e;
    const size_t m = n_src > 0 ? n_src : 1;
    if ((e = d2.reserve(m)) != cudaSuccess) return e;
    if ((e = idx.reserve(m)) != cudaSuccess) return e;
    if ((e = pending.reserve(m)) != cudaSuccess) return e;
    if ((e = pending2.reserve(m)) != cudaSuccess) return e;
    if ((e = queries.reserve(m)) != cudaSuccess) return e;
    if ((e = n_pending.reserve(2)) != cudaSuccess) return e;
    if ((e = T.reserve(16)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(n_pending.p, 0, 2 * sizeof(unsigned int), st)) != cudaSuccess) return e;
    if (T_colmajor_host && (e = cudaMemcpyAsync(T.p, T_colmajor_host, 64, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if (n_src > 0) {
      launch_counter() += 3;
      k_nn_search<<<(n_src + 255) / 256, 256, 0, st>>>(view(), d_src, n_src, T.p, T_colmajor_host ? 1 : 0, max_d2, d2.p, idx.p, queries.p, pending.p, n_pending.p);
      k_nn_far<<<kNumSM * 4, 256, 0, st>>>(view(), queries.p, pending.p, n_pending.p, max_d2, d2.p, idx.p, pending2.p, n_pending.p + 1);
      k_nn_bruteforce<<<kNumSM * 4, 256, 0, st>>>(view(), queries.p, pending2.p, n_pending.p + 1, d2.p, idx.p);
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

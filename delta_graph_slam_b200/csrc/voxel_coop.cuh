// b200reg — the voxel key / radix sort / run segmentation pipeline as ONE persistent cooperative
// kernel (one CTA per SM, grid barriers between phases).
//
// voxel_sort.cuh runs the same pipeline as 17 launches of ~5 us kernels that each touch ~2 MB: on a
// B200 that is pure launch / dependency latency (the data would stream through HBM in < 1 us).
// Here a scan lives in the registers of at most 148 CTAs for the whole pipeline:
//   A  per-CTA min / max of its points                                  -> barrier
//   B  every CTA rebuilds the lattice from the 148 partial extents (no atomics, no init kernel),
//      computes the keys of its tile straight into registers
//   per 8-bit pass:  per-warp digit counts -> tile histogram            -> barrier
//                    bases of this tile (redundant per CTA: every CTA reads the histograms of
//                    the lower tiles, a few KB out of L2) -> stable scatter    -> barrier
//                    (the next pass re-loads its tile from the scattered buffer)
//   C  run heads per tile                                               -> barrier
//      slots of the heads (exclusive count over the lower tiles), vox_start / vox_key, totals
// Bit-for-bit the same result as the multi-kernel path (same keys, same stable order, same runs);
// tests/test_gpu_parity.py compares both.  Used when the cloud fits one tile per CTA
// (n <= 148 * 256 * 32 = 1.2 M points, i.e. every BASELINE configuration); larger clouds take the
// multi-kernel path.
#pragma once
#include "voxel_sort.cuh"

namespace b200 {

constexpr int kCoopMaxItems = 32;

struct CoopSortArgs {
  const float4* pts;
  int n, is_dense;
  PointGate gate;
  float lx, ly, lz;
  SortMeta* meta;
  uint32_t *keys_a, *vals_a, *keys_b, *vals_b;
  uint32_t* hist;        // [n_tiles][256], tile-major
  uint32_t* tile_heads;  // [n_tiles][2]: run heads, valid points
  uint32_t* seg_first;   // [n / kSegFirstTile + 1]
  uint32_t *vox_start, *vox_key, *point_key;
  int* slots;            // [gridDim][8]: min xyz, max xyz (ordered ints), any
  unsigned int* barrier; // zeroed by the host before the launch
};

__device__ __forceinline__ void coop_barrier(unsigned int* counter, unsigned int& epoch) {
  epoch += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < epoch);
    __threadfence();
  }
  __syncthreads();
}

// the last CTA to leave zeroes the barrier counter (word 0) and the exit counter (word 1), so the
// next launch needs no memset in front of it
__device__ __forceinline__ void coop_exit(unsigned int* counter) {
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(counter + 1, 1u);
    if (prev == gridDim.x - 1) {
      counter[0] = 0u;
      counter[1] = 0u;
      __threadfence();
    }
  }
}

// exclusive scan of one value per thread over a 256-thread block; returns the exclusive prefix,
// *total receives the block sum
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const uint32_t c = s_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  if (total) *total = tot;
  return off + incl - v;
}

template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads, 1) k_voxel_sort_coop(CoopSortArgs a) {
  constexpr int WARPS = kSortThreads / 32;
  constexpr int TILE = kSortThreads * ITEMS;
  __shared__ uint32_t cnt[WARPS][kRadix];
  __shared__ GridParams g;
  __shared__ uint32_t s_skip, s_nbits;
  __shared__ uint32_t s_warp[8];
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ int s_any[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, n = a.n;
  unsigned int epoch = 0;

  // ---- A: extents of this CTA's points
  {
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    int any = 0;
    for (int i = blockIdx.x * kSortThreads + tid; i < n; i += gridDim.x * kSortThreads) {
      const float4 p = __ldg(a.pts + i);
      if (!point_takes_part(a.gate, a.is_dense, p.x, p.y, p.z)) continue;
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      any = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
        mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
      }
      any |= __shfl_xor_sync(0xffffffffu, any, o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { s_mn[warp][k] = mn[k]; s_mx[warp][k] = mx[k]; }
      s_any[warp] = any;
    }
    __syncthreads();
    if (tid < 3) {
      float lo = s_mn[0][tid], hi = s_mx[0][tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_mn[w][tid]); hi = fmaxf(hi, s_mx[w][tid]); }
      a.slots[blockIdx.x * 8 + tid] = float_to_ordered(lo);
      a.slots[blockIdx.x * 8 + 3 + tid] = float_to_ordered(hi);
    } else if (tid == 3) {
      int an = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) an |= s_any[w];
      a.slots[blockIdx.x * 8 + 6] = an;
    }
  }
  coop_barrier(a.barrier, epoch);

  // ---- B: the lattice (every CTA, identical), then the keys of this tile into registers
  if (warp == 0) {
    int mm[7];
#pragma unroll
    for (int k = 0; k < 3; ++k) { mm[k] = 0x7FFFFFFF; mm[3 + k] = (int)0x80000000; }
    mm[6] = 0;
    for (int c = lane; c < (int)gridDim.x; c += 32) {
      const int* sl = a.slots + c * 8;
      if (__ldcg(sl + 6)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { mm[k] = min(mm[k], __ldcg(sl + k)); mm[3 + k] = max(mm[3 + k], __ldcg(sl + 3 + k)); }
        mm[6] = 1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        mm[k] = min(mm[k], __shfl_xor_sync(0xffffffffu, mm[k], o));
        mm[3 + k] = max(mm[3 + k], __shfl_xor_sync(0xffffffffu, mm[3 + k], o));
      }
      mm[6] |= __shfl_xor_sync(0xffffffffu, mm[6], o);
    }
    if (lane == 0) {
      make_grid(mm, a.lx, a.ly, a.lz, g);
      unsigned long long cells = (unsigned long long)g.div_b[0] * (unsigned long long)g.div_b[1] * (unsigned long long)g.div_b[2];
      if (cells > 0xFFFFFFFEull) cells = 0xFFFFFFFEull;
      s_skip = (uint32_t)cells;
      s_nbits = g.overflow ? 0u : (uint32_t)(64 - __clzll((unsigned long long)s_skip));
      if (blockIdx.x == 0) {
        a.meta->grid = g;
        a.meta->skip_key = s_skip;
        a.meta->nbits = s_nbits;
        a.meta->n_vox = 0;
        a.meta->n_valid = 0;
      }
    }
  }
  __syncthreads();
  if (g.overflow) {  // "leaf size too small": the callers copy the input instead (uniform exit)
    coop_exit(a.barrier);
    return;
  }
  const uint32_t skip = s_skip, nbits = s_nbits;
  const int wbase = tile * TILE + warp * (32 * ITEMS);
  uint32_t k[ITEMS], v[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    k[r] = 0u;
    v[r] = (uint32_t)i;
    if (i < n) {
      const float4 p = __ldg(a.pts + i);
      k[r] = point_takes_part(a.gate, a.is_dense, p.x, p.y, p.z) ? voxel_key(g, p.x, p.y, p.z) : skip;
      if (a.point_key) a.point_key[i] = (k[r] == skip) ? kInvalidKey : k[r];
    }
  }

  // ---- radix passes
  const int n_pass = (int)((nbits + kRadixBits - 1) / kRadixBits);
  for (int pass = 0; pass < n_pass; ++pass) {
    const int shift = pass * kRadixBits;
    const uint32_t* ki = (pass & 1) ? a.keys_b : a.keys_a;
    const uint32_t* vi = (pass & 1) ? a.vals_b : a.vals_a;
    uint32_t* ko = (pass & 1) ? a.keys_a : a.keys_b;
    uint32_t* vo = (pass & 1) ? a.vals_a : a.vals_b;
    if (pass > 0) {
#pragma unroll
      for (int r = 0; r < ITEMS; ++r) {
        const int i = wbase + r * 32 + lane;
        const bool ok = i < n;
        k[r] = ok ? __ldcg(ki + i) : 0u;
        v[r] = ok ? __ldcg(vi + i) : 0u;
      }
    }
    for (int i = tid; i < WARPS * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();
    // per-warp digit counts
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = wbase + r * 32 + lane;
      const bool ok = i < n;
      const uint32_t dgt = ok ? ((k[r] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
      const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
      if (ok && (__ffs(peers) - 1) == lane) cnt[warp][dgt] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    {
      uint32_t c = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) c += cnt[w][tid];
      a.hist[tile * kRadix + tid] = c;
    }
    coop_barrier(a.barrier, epoch);
    // bases of this tile: digits below d over all tiles + digit d over the lower tiles
    {
      const int d = tid;
      uint32_t before = 0, tot = 0;
      const int n_tiles = (int)gridDim.x;
      int t = 0;
      for (; t + 4 <= n_tiles; t += 4) {
        const uint32_t c0 = __ldcg(a.hist + (t + 0) * kRadix + d), c1 = __ldcg(a.hist + (t + 1) * kRadix + d);
        const uint32_t c2 = __ldcg(a.hist + (t + 2) * kRadix + d), c3 = __ldcg(a.hist + (t + 3) * kRadix + d);
        before += (t + 0 < tile ? c0 : 0u) + (t + 1 < tile ? c1 : 0u) + (t + 2 < tile ? c2 : 0u) + (t + 3 < tile ? c3 : 0u);
        tot += c0 + c1 + c2 + c3;
      }
      for (; t < n_tiles; ++t) {
        const uint32_t c = __ldcg(a.hist + t * kRadix + d);
        before += t < tile ? c : 0u;
        tot += c;
      }
      uint32_t run = block_excl_scan256(tot, s_warp, nullptr) + before;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = cnt[w][d];
        cnt[w][d] = run;
        run += c;
      }
    }
    __syncthreads();
    // stable scatter, ranks inside the warp round in lane order
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = wbase + r * 32 + lane;
      const bool ok = i < n;
      const uint32_t dgt = ok ? ((k[r] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
      const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t dst = 0;
      if (ok) dst = cnt[warp][dgt] + rank;
      __syncwarp();
      if (ok && (__ffs(peers) - 1) == lane) cnt[warp][dgt] += __popc(peers);
      __syncwarp();
      if (ok) {
        __stcg(ko + dst, k[r]);
        __stcg(vo + dst, v[r]);
      }
    }
    coop_barrier(a.barrier, epoch);
  }

  // ---- C: run heads.  The sorted pairs sit in B after an odd number of passes, else in A.
  const uint32_t* keys = (n_pass & 1) ? a.keys_b : a.keys_a;
  uint32_t heads = 0, valid = 0;
  uint32_t hk[ITEMS];
  uint32_t hflag = 0;  // bit r: element r of this thread is a head
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = tile * TILE + r * kSortThreads + tid;  // linear order inside the tile
    hk[r] = 0;
    if (i < n) {
      hk[r] = __ldcg(keys + i);
      const bool vld = hk[r] != skip;
      const bool head = vld && (i == 0 || __ldcg(keys + i - 1) != hk[r]);
      valid += vld;
      heads += head;
      if (head) hflag |= 1u << r;
    }
  }
  {
    uint32_t th = 0, tv = 0;
    block_excl_scan256(heads, s_warp, &th);
    block_excl_scan256(valid, s_warp, &tv);
    if (tid == 0) { a.tile_heads[2 * tile] = th; a.tile_heads[2 * tile + 1] = tv; }
  }
  coop_barrier(a.barrier, epoch);
  {
    // heads in the lower tiles, and the totals
    uint32_t before = 0, tot_h = 0, tot_v = 0;
    for (int t = tid; t < (int)gridDim.x; t += kSortThreads) {
      const uint32_t h = __ldcg(a.tile_heads + 2 * t), vv = __ldcg(a.tile_heads + 2 * t + 1);
      if (t < tile) before += h;
      tot_h += h;
      tot_v += vv;
    }
    uint32_t base = 0, TH = 0, TV = 0;
    block_excl_scan256(before, s_warp, &base);
    block_excl_scan256(tot_h, s_warp, &TH);
    block_excl_scan256(tot_v, s_warp, &TV);
    // slot of every head: linear order = r-major over the tile (element r * 256 + tid)
    uint32_t running = base;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const uint32_t is_head = (hflag >> r) & 1u;
      uint32_t row_total = 0;
      const uint32_t excl = block_excl_scan256(is_head, s_warp, &row_total);
      // row r of the tile starts at a multiple of 256 sorted positions: the runs in front of it are k_vg_centroids' split
      if (tid == 0 && tile * TILE + r * kSortThreads < n) a.seg_first[(tile * TILE + r * kSortThreads) / kSegFirstTile] = running;
      if (is_head) {
        const uint32_t slot = running + excl;
        a.vox_start[slot] = (uint32_t)(tile * TILE + r * kSortThreads + tid);
        a.vox_key[slot] = hk[r];
      }
      running += row_total;
    }
    if (blockIdx.x == 0 && tid == 0) {
      a.meta->n_vox = TH;
      a.meta->n_valid = TV;
      a.vox_start[TH] = TV;
    }
  }
  coop_exit(a.barrier);
}

// host side: launch when the cloud fits one tile per CTA; returns false when the caller must take
// the multi-kernel path
inline bool coop_sort_items(int n, int num_sm, int* items_out) {
  int items = 2;
  while (items <= kCoopMaxItems && (long long)num_sm * kSortThreads * items < (long long)n) items *= 2;
  if (items > kCoopMaxItems) return false;
  *items_out = items;
  return true;
}

inline bool launch_voxel_sort_coop(VoxelSort& vs, cudaStream_t st, const float4* d_pts, int n, int is_dense, PointGate gate, float lx, float ly, float lz, bool keep_point_keys, cudaError_t* err) {
  int items = 2;
  if (!coop_sort_items(n, vs.max_ctas, &items)) return false;
  const int tile = kSortThreads * items;
  const int n_tiles = n > 0 ? (n + tile - 1) / tile : 1;
  cudaError_t e;
  *err = cudaSuccess;
  if ((e = vs.hist.reserve((size_t)kNumSM * kRadix)) != cudaSuccess) { *err = e; return true; }
  if ((e = vs.tile_heads.reserve((size_t)2 * kNumSM)) != cudaSuccess) { *err = e; return true; }
  if ((e = vs.mm.reserve((size_t)8 * kNumSM)) != cudaSuccess) { *err = e; return true; }
  if (!vs.coop_bar.p) {  // zeroed once; every kernel restores the counters on exit
    if ((e = vs.coop_bar.reserve(32)) != cudaSuccess) { *err = e; return true; }
    if ((e = cudaMemsetAsync(vs.coop_bar.p, 0, vs.coop_bar.cap * sizeof(unsigned int), st)) != cudaSuccess) { *err = e; return true; }
  }
  CoopSortArgs a;
  a.pts = d_pts; a.n = n; a.is_dense = is_dense; a.gate = gate; a.lx = lx; a.ly = ly; a.lz = lz;
  a.meta = vs.meta.p; a.keys_a = vs.keys_a.p; a.vals_a = vs.vals_a.p; a.keys_b = vs.keys_b.p; a.vals_b = vs.vals_b.p;
  a.hist = vs.hist.p; a.tile_heads = vs.tile_heads.p; a.vox_start = vs.vox_start.p; a.vox_key = vs.vox_key.p; a.seg_first = vs.seg_first.p;
  a.point_key = keep_point_keys ? vs.point_key.p : nullptr;
  a.slots = vs.mm.p; a.barrier = vs.coop_bar.p;
  void* args[] = {(void*)&a};
  const void* fn = nullptr;
  switch (items) {
    case 2: fn = (const void*)k_voxel_sort_coop<2>; break;
    case 4: fn = (const void*)k_voxel_sort_coop<4>; break;
    case 8: fn = (const void*)k_voxel_sort_coop<8>; break;
    case 16: fn = (const void*)k_voxel_sort_coop<16>; break;
    default: fn = (const void*)k_voxel_sort_coop<32>; break;
  }
  launch_counter() += 1;
  *err = cudaLaunchCooperativeKernel(fn, dim3(n_tiles), dim3(kSortThreads), args, 0, st);
  return true;
}

}  // namespace b200

// b200reg — voxel keys + stable LSD radix sort + run segmentation.
//
// The machinery shared by VoxelGrid down-sampling, the NDT target grid and the
// exact-NN cell grid: every point gets the linear index of its voxel (bit-exact with
// pcl::VoxelGrid), (key, point index) pairs are sorted by key
// with a hand-written stable radix sort (so points inside a voxel stay in input order),
// and runs of equal keys are numbered.  Nothing here synchronises with the host: sizes
// that depend on the data (bits to sort, number of voxels) stay in device memory and the
// kernels that need them read them there.
#pragma once
#include "common.cuh"
#include "radix_onesweep.cuh"

namespace b200 {

constexpr int kSortThreads = 256;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;

struct SortMeta {     // device-resident, written by k_grid_keys / k_seg_scan
  GridParams grid;
  uint32_t nbits;     // key bits that have to be sorted
  uint32_t skip_key;  // key given to skipped (non-finite) points: sorts after every voxel
  uint32_t n_vox;     // number of runs (occupied voxels), skipped points excluded
  uint32_t n_valid;   // points that received a voxel
};

// ---- which points take part ---------------------------------------------------
// PrefilteringNodelet::distance_filter [REF apps/prefiltering_nodelet.cpp:275-291] fused into the key
// pipeline: with the gate on a point takes part when near < |p| < far, |p| = Eigen's float norm()
// (sqrtf((x*x + y*y) + z*z)) widened to double for the comparison with the double thresholds — the
// copy_if pass and the cloud it materialises never exist.  NaN / inf fail both comparisons, which is
// also what the finite test of the is_dense = false cloud the reference hands to VoxelGrid would drop.
struct PointGate {
  int on;
  double near_thresh, far_thresh;
};
constexpr PointGate kNoGate = {0, 0.0, 0.0};
__device__ __forceinline__ bool point_takes_part(const PointGate& gate, int is_dense, float x, float y, float z) {
  if (gate.on) {
    // on == 2: PrefilteringNodelet::height_filtering [REF apps/prefiltering_nodelet.cpp:198-214] as a gate — the point
    // takes part when z > near_thresh (float widened to double, as the reference compares) and it is finite
    if (gate.on == 2) return finite3(x, y, z) && (double)z > gate.near_thresh;
    const float sq = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const double d = (double)__fsqrt_rn(sq);
    return d > gate.near_thresh && d < gate.far_thresh;
  }
  return is_dense || finite3(x, y, z);
}

// ---- min / max -------------------------------------------------------------
static __global__ void k_minmax_init(int* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = 0x7FFFFFFF;
  else if (threadIdx.x < 6) mm[threadIdx.x] = (int)0x80000000;
  else if (threadIdx.x == 6) mm[6] = 0;
}

static __global__ void __launch_bounds__(256) k_minmax(const float4* __restrict__ pts, int n, int is_dense, PointGate gate, int* mm) {
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  int any = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(pts + i);
    if (!point_takes_part(gate, is_dense, p.x, p.y, p.z)) continue;
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    any = 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    any |= __shfl_xor_sync(0xffffffffu, any, o);
  }
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ int s_any[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_mn[warp][a] = mn[a]; s_mx[warp][a] = mx[a]; }
    s_any[warp] = any;
  }
  __syncthreads();
  if (threadIdx.x < 3) {  // one atomic pair per axis per block
    const int a = threadIdx.x;
    float lo = s_mn[0][a], hi = s_mx[0][a];
    int an = s_any[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_mn[w][a]); hi = fmaxf(hi, s_mx[w][a]); an |= s_any[w]; }
    if (an) {
      atomicMin(mm + a, float_to_ordered(lo));
      atomicMax(mm + 3 + a, float_to_ordered(hi));
      if (a == 0) atomicOr(mm + 6, 1);
    }
  }
}

// A.1 steps 2-3 from the reduced min / max
__device__ __forceinline__ void make_grid(const int* mm, float lx, float ly, float lz, GridParams& g) {
  const float leaf[3] = {lx, ly, lz};
  g.any = mm[6];
  g.overflow = 0;
  long long cells = 1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    g.leaf[a] = leaf[a];
    g.inv_leaf[a] = __fdiv_rn(1.0f, leaf[a]);
    float mn = ordered_to_float(mm[a]), mx = ordered_to_float(mm[3 + a]);
    if (!g.any) { mn = 0.f; mx = 0.f; }
    long long d = (long long)__fmul_rn(__fsub_rn(mx, mn), g.inv_leaf[a]) + 1;
    cells *= d;
    g.min_b[a] = (int)floorf(__fmul_rn(mn, g.inv_leaf[a]));
    g.max_b[a] = (int)floorf(__fmul_rn(mx, g.inv_leaf[a]));
    g.div_b[a] = g.max_b[a] - g.min_b[a] + 1;
  }
  // the guard multiplies the three truncated extents exactly as PCL does (int64)
  {
    long long dx = (long long)__fmul_rn(__fsub_rn(ordered_to_float(mm[3]), ordered_to_float(mm[0])), g.inv_leaf[0]) + 1;
    long long dy = (long long)__fmul_rn(__fsub_rn(ordered_to_float(mm[4]), ordered_to_float(mm[1])), g.inv_leaf[1]) + 1;
    long long dz = (long long)__fmul_rn(__fsub_rn(ordered_to_float(mm[5]), ordered_to_float(mm[2])), g.inv_leaf[2]) + 1;
    if (g.any && dx * dy * dz > 2147483647ll) g.overflow = 1;
  }
  g.mul[0] = 1;
  g.mul[1] = g.div_b[0];
  g.mul[2] = g.div_b[0] * g.div_b[1];
}

__device__ __forceinline__ uint32_t voxel_key(const GridParams& g, float x, float y, float z) {
  int i0 = (int)__fsub_rn(floorf(__fmul_rn(x, g.inv_leaf[0])), (float)g.min_b[0]);
  int i1 = (int)__fsub_rn(floorf(__fmul_rn(y, g.inv_leaf[1])), (float)g.min_b[1]);
  int i2 = (int)__fsub_rn(floorf(__fmul_rn(z, g.inv_leaf[2])), (float)g.min_b[2]);
  return (uint32_t)(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
}

// keys (A.1 step 4) + pass-0 digit histogram per tile
static __global__ void __launch_bounds__(kSortThreads) k_grid_keys(const float4* __restrict__ pts, int n, int is_dense, PointGate gate, float lx, float ly, float lz, const int* __restrict__ mm,
                                                             SortMeta* meta, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ point_key) {
  __shared__ GridParams g;
  __shared__ uint32_t s_skip;
  if (threadIdx.x == 0) {
    make_grid(mm, lx, ly, lz, g);
    // cells < 2^31 when !overflow (div_b product can exceed the truncated-extent product by a
    // rounding cell per axis; clamp so the skip key still sorts last)
    unsigned long long cells = (unsigned long long)g.div_b[0] * (unsigned long long)g.div_b[1] * (unsigned long long)g.div_b[2];
    if (cells > 0xFFFFFFFEull) cells = 0xFFFFFFFEull;
    s_skip = (uint32_t)cells;
    if (blockIdx.x == 0) {
      meta->grid = g;
      meta->skip_key = s_skip;
      meta->nbits = g.overflow ? 0u : (uint32_t)(64 - __clzll((unsigned long long)s_skip));
      meta->n_vox = 0;
      meta->n_valid = 0;
    }
  }
  __syncthreads();
  if (g.overflow) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(pts + i);
    uint32_t k = s_skip;
    if (point_takes_part(gate, is_dense, p.x, p.y, p.z)) k = voxel_key(g, p.x, p.y, p.z);
    keys[i] = k;
    vals[i] = (uint32_t)i;
    if (point_key) point_key[i] = (k == s_skip) ? kInvalidKey : k;
  }
}

// ---- radix sort: one pass = hist -> scan -> scatter --------------------------
// hist[digit][tile]: digit-major, so the global order of the scatter (digit, then tile) is a flat
// exclusive scan of the array
template <int ITEMS>
static __global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t* __restrict__ keys, int n, int pass, const SortMeta* __restrict__ meta, uint32_t* __restrict__ hist) {
  if ((uint32_t)(pass * kRadixBits) >= meta->nbits) return;
  __shared__ uint32_t s[kRadix];
  s[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * (kSortThreads * ITEMS);
  const int shift = pass * kRadixBits;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    int i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&s[(keys[i] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = s[threadIdx.x];
}

// exclusive scan of the flat hist[digit][tile] array, in place; one block.  Each thread owns up to
// 16 consecutive uint4 (all loads issued up front, prefix in registers), the 1024 thread totals are
// scanned with shuffles.  The array (kRadix * n_tiles <= 65536 counters) is a multiple of 4 long.
static __global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* __restrict__ hist, int n_tiles, int pass, const SortMeta* __restrict__ meta) {
  if ((uint32_t)(pass * kRadixBits) >= meta->nbits) return;
  __shared__ uint32_t s_warp[32];
  const int N = kRadix * n_tiles;           // multiple of 4 (kRadix is)
  const int nvec = N >> 2;
  const int V = (nvec + 1023) / 1024;       // uint4 per thread, <= 16
  const int vbase = threadIdx.x * V;
  uint4* h4 = reinterpret_cast<uint4*>(hist);
  uint4 v[16];
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = make_uint4(0, 0, 0, 0);
    if (i < V && vbase + i < nvec) v[i] = h4[vbase + i];
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) sum += v[i].x + v[i].y + v[i].z + v[i].w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - sum;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint4 o;
    o.x = run; run += v[i].x;
    o.y = run; run += v[i].y;
    o.z = run; run += v[i].z;
    o.w = run; run += v[i].w;
    if (i < V && vbase + i < nvec) h4[vbase + i] = o;
  }
}

// stable scatter: each warp owns a contiguous slice of the tile and walks it in rounds of 32
template <int ITEMS>
static __global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out, int n, int pass, const SortMeta* __restrict__ meta, const uint32_t* __restrict__ hist) {
  if ((uint32_t)(pass * kRadixBits) >= meta->nbits) return;
  constexpr int WARPS = kSortThreads / 32;
  __shared__ uint32_t cnt[WARPS][kRadix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < WARPS * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int shift = pass * kRadixBits;
  const int wbase = blockIdx.x * (kSortThreads * ITEMS) + warp * (32 * ITEMS);
  uint32_t k[ITEMS], v[ITEMS];
  // phase A: per-warp digit counts
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    int i = wbase + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? keys_in[i] : 0u;
    v[r] = ok ? vals_in[i] : 0u;
    uint32_t dgt = ok ? ((k[r] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;  // kRadix = "no element"
    uint32_t peers = __match_any_sync(0xffffffffu, dgt);
    if (ok && (__ffs(peers) - 1) == lane) cnt[warp][dgt] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // phase B: digit bases for every warp = global base of (digit, tile) + counts of lower warps
  {
    const int d = threadIdx.x;  // kSortThreads == kRadix
    uint32_t run = hist[d * gridDim.x + blockIdx.x];
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase C: ranks inside the warp round, in lane order
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    int i = wbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t dgt = ok ? ((k[r] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
    uint32_t peers = __match_any_sync(0xffffffffu, dgt);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t dst = 0;
    if (ok) dst = cnt[warp][dgt] + rank;
    __syncwarp();
    if (ok && (__ffs(peers) - 1) == lane) cnt[warp][dgt] += __popc(peers);
    __syncwarp();
    if (ok) {
      keys_out[dst] = k[r];
      vals_out[dst] = v[r];
    }
  }
}

// ---- run segmentation --------------------------------------------------------
// sorted buffers live in A when an even number of passes ran, else in B
__device__ __forceinline__ bool sorted_in_b(const SortMeta* meta) { return (((meta->nbits + kRadixBits - 1) / kRadixBits) & 1u) != 0; }

// Segmentation tiles: kSegItems consecutive sorted positions per thread.  The second kernel adds up the head counts of
// all earlier tiles per block — with 256-position tiles that was 3641 x 3641 / 2 loads on a 1 M-point scan and a third of
// the kernel's stall samples (profiles/r02_ncu_voxelgrid_1m.txt); 2048-position tiles make it 455 x 455 / 2.
constexpr int kSegFirstTile = 256;  // seg_first[t] = runs that start in front of sorted position t * kSegFirstTile (k_vg_centroids' work split)
constexpr int kSegItems = 8;
constexpr int kSegTile = 256 * kSegItems;

// keys of this thread's kSegItems positions (kInvalid beyond n) and the key in front of the first one
__device__ __forceinline__ void seg_load(const uint32_t* __restrict__ keys, int n, int first, uint32_t skip, uint32_t (&k)[kSegItems], uint32_t& prev, bool& has_prev) {
  if (first + kSegItems <= n) {
    const uint4 a = *reinterpret_cast<const uint4*>(keys + first), b = *reinterpret_cast<const uint4*>(keys + first + 4);
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
  } else {
#pragma unroll
    for (int r = 0; r < kSegItems; ++r) k[r] = first + r < n ? keys[first + r] : skip;
  }
  has_prev = first > 0 && first < n;
  prev = has_prev ? keys[first - 1] : 0u;
}

// heads per tile
static __global__ void __launch_bounds__(256) k_seg_count(const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, int n, const SortMeta* __restrict__ meta,
                                                   uint32_t* __restrict__ tile_heads, uint32_t* __restrict__ tile_valid) {
  if (meta->grid.overflow) return;
  const uint32_t* keys = sorted_in_b(meta) ? keys_b : keys_a;
  const uint32_t skip = meta->skip_key;
  const int first = (blockIdx.x * 256 + threadIdx.x) * kSegItems;
  uint32_t k[kSegItems], prev;
  bool has_prev;
  seg_load(keys, n, first, skip, k, prev, has_prev);
  int head = 0, valid = 0;
#pragma unroll
  for (int r = 0; r < kSegItems; ++r) {
    const bool in = first + r < n, val = in && k[r] != skip;
    const bool differs = r == 0 ? (!has_prev || prev != k[0]) : (k[r - 1] != k[r]);
    valid += val;
    head += val && differs;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_h[8], s_v[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    head += __shfl_xor_sync(0xffffffffu, head, o);
    valid += __shfl_xor_sync(0xffffffffu, valid, o);
  }
  if (lane == 0) { s_h[warp] = head; s_v[warp] = valid; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int h = 0, v = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { h += s_h[w]; v += s_v[w]; }
    tile_heads[blockIdx.x] = (uint32_t)h;
    tile_valid[blockIdx.x] = (uint32_t)v;
  }
}

// slot of every head (= exclusive count of heads before it); vox_start[slot] = position of the
// head in the sorted order; vox_start[n_vox] = n_valid closes the last run
static __global__ void __launch_bounds__(256) k_seg_scan(const uint32_t* __restrict__ keys_a, const uint32_t* __restrict__ keys_b, int n, SortMeta* meta,
                                                  const uint32_t* __restrict__ tile_heads, const uint32_t* __restrict__ tile_valid, int n_tiles,
                                                  uint32_t* __restrict__ vox_start, uint32_t* __restrict__ vox_key, uint32_t* __restrict__ seg_first) {
  if (meta->grid.overflow) return;
  const uint32_t* keys = sorted_in_b(meta) ? keys_b : keys_a;
  const uint32_t skip = meta->skip_key;
  __shared__ uint32_t s_red[8];
  __shared__ uint32_t s_base, s_tot, s_valid;
  // this thread's positions first: the loads overlap the sums over the tiles below
  const int first = (blockIdx.x * 256 + threadIdx.x) * kSegItems;
  uint32_t k[kSegItems], prev;
  bool has_prev;
  seg_load(keys, n, first, skip, k, prev, has_prev);
  // base = heads in all earlier tiles (block-wide strided sum); block 0 also needs the totals
  const bool totals = blockIdx.x == 0;
  const int t_end = totals ? n_tiles : (int)blockIdx.x;
  uint32_t part = 0, tot = 0, val = 0;
  for (int t = threadIdx.x; t < t_end; t += blockDim.x) {
    const uint32_t h = tile_heads[t];
    if (t < (int)blockIdx.x) part += h;
    tot += h;
    if (totals) val += tile_valid[t];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_sum = [&](uint32_t x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += s_red[w];
    return r;
  };
  uint32_t b = block_sum(part);
  uint32_t T = 0, V = 0;
  if (totals) { T = block_sum(tot); V = block_sum(val); }  // block-uniform branch
  if (threadIdx.x == 0) { s_base = b; s_tot = T; s_valid = V; }
  uint32_t heads = 0;  // bit r: position first + r opens a run
#pragma unroll
  for (int r = 0; r < kSegItems; ++r) {
    const bool val_r = first + r < n && k[r] != skip;
    const bool differs = r == 0 ? (!has_prev || prev != k[0]) : (k[r - 1] != k[r]);
    if (val_r && differs) heads |= 1u << r;
  }
  // in-block exclusive scan of the per-thread head counts
  const uint32_t mine = __popc(heads);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) s_red[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) woff += s_red[w];
  uint32_t slot = s_base + woff + incl - mine;
  if (first < n && first % kSegFirstTile == 0) seg_first[first / kSegFirstTile] = slot;
#pragma unroll
  for (int r = 0; r < kSegItems; ++r) {
    if (heads & (1u << r)) {
      vox_start[slot] = (uint32_t)(first + r);
      vox_key[slot] = k[r];
      ++slot;
    }
  }
  if (totals && threadIdx.x == 0) {
    meta->n_vox = s_tot;
    meta->n_valid = s_valid;
    vox_start[s_tot] = s_valid;
  }
}

// ---- host driver -------------------------------------------------------------
struct VoxelSort;
// voxel_coop.cuh: the whole pipeline as one cooperative kernel; false = cloud too large, take the multi-kernel path
bool launch_voxel_sort_coop(VoxelSort& vs, cudaStream_t st, const float4* d_pts, int n, int is_dense, PointGate gate, float lx, float ly, float lz, bool keep_point_keys, cudaError_t* err);
// 0 = default: the cooperative single kernel for clouds up to kCoopDefaultMax points, the one-sweep sort above that;
// 1 = always the three-launch-per-digit path (A/B tests, <= 2 M points); 2 = always the one-sweep sort; 3 = always the
// cooperative kernel when the cloud fits it
constexpr int kCoopDefaultMax = 400000;
inline int& sort_path_override() {
  static int v = 0;
  return v;
}

struct VoxelSort {
  DevBuf<uint32_t> keys_a, keys_b, vals_a, vals_b, hist, tile_heads, tile_valid, vox_start, vox_key, point_key, seg_first;
  DevBuf<int> mm;
  DevBuf<unsigned int> coop_bar;
  DevBuf<SortMeta> meta;
  OneSweepScratch onesweep;
  int n = 0;
  int max_ctas = kNumSM;  // CTAs the cooperative path may occupy (b200reg_set_sm_budget)

  void release() {
    keys_a.release(); keys_b.release(); vals_a.release(); vals_b.release(); hist.release(); tile_heads.release(); tile_valid.release();
    vox_start.release(); vox_key.release(); point_key.release(); seg_first.release(); mm.release(); meta.release(); coop_bar.release(); onesweep.buf.release();
  }

  // enqueue: keys -> sort -> segmentation.  Afterwards (on the stream):
  //   meta->grid / n_vox / n_valid, sorted (key, point index) in A or B (sorted_in_b), vox_start[0..n_vox], vox_key[0..n_vox)
  // gather_dst (optional): the points in sorted order, written by the one-sweep sort's last pass; *gathered tells the
  // caller whether that happened (the other sort paths leave it to a k_vg_gather launch)
  cudaError_t run(cudaStream_t st, const float4* d_pts, int n_points, int is_dense, float lx, float ly, float lz, bool keep_point_keys, PointGate gate = kNoGate,
                  float4* gather_dst = nullptr, bool* gathered = nullptr) {
    if (gathered) *gathered = false;
    n = n_points;
    cudaError_t e;
    size_t nn = (size_t)(n > 0 ? n : 1);
    if ((e = keys_a.reserve(nn)) != cudaSuccess) return e;
    if ((e = keys_b.reserve(nn)) != cudaSuccess) return e;
    if ((e = vals_a.reserve(nn)) != cudaSuccess) return e;
    if ((e = vals_b.reserve(nn)) != cudaSuccess) return e;
    if ((e = vox_start.reserve(nn + 1)) != cudaSuccess) return e;
    if ((e = vox_key.reserve(nn)) != cudaSuccess) return e;
    if ((e = seg_first.reserve(nn / kSegFirstTile + 2)) != cudaSuccess) return e;
    if ((e = mm.reserve(8)) != cudaSuccess) return e;
    if ((e = meta.reserve(1)) != cudaSuccess) return e;
    if (keep_point_keys && (e = point_key.reserve(nn)) != cudaSuccess) return e;
    const int path = sort_path_override();
    if ((path == 0 && n <= kCoopDefaultMax) || path == 3) {
      cudaError_t ce = cudaSuccess;
      if (launch_voxel_sort_coop(*this, st, d_pts, n, is_dense, gate, lx, ly, lz, keep_point_keys, &ce)) return ce;
    }
    const bool one_sweep = path != 1;
    int items = 4;
    while (items < 32 && (n + kSortThreads * items - 1) / (kSortThreads * items) > 256) items *= 2;
    const int tile = kSortThreads * items;
    const int n_tiles = n > 0 ? (n + tile - 1) / tile : 1;
    const int n_seg_tiles = n > 0 ? (n + kSegTile - 1) / kSegTile : 1;
    if (!one_sweep && n_tiles > 256) return cudaErrorInvalidValue;  // the one-block histogram scan holds 65536 counters
    if ((e = hist.reserve((size_t)n_tiles * kRadix)) != cudaSuccess) return e;
    if ((e = tile_heads.reserve(n_seg_tiles)) != cudaSuccess) return e;
    if ((e = tile_valid.reserve(n_seg_tiles)) != cudaSuccess) return e;

    k_minmax_init<<<1, 32, 0, st>>>(mm.p);
    launch_counter() += 4 + (n > 0 ? 1 : 0) + (n > 0 ? 12 : 0);  // init, minmax, keys, 4 x (hist, scan, scatter), 2 x segmentation
    int blocks = n > 0 ? (n + 255) / 256 : 1;
    if (blocks > kNumSM * 4) blocks = kNumSM * 4;
    int mm_blocks = n > 0 ? (n + 1023) / 1024 : 1;
    if (mm_blocks > kNumSM * 2) mm_blocks = kNumSM * 2;
    if (n > 0) k_minmax<<<mm_blocks, 256, 0, st>>>(d_pts, n, is_dense, gate, mm.p);
    k_grid_keys<<<blocks, kSortThreads, 0, st>>>(d_pts, n, is_dense, gate, lx, ly, lz, mm.p, meta.p, keys_a.p, vals_a.p, keep_point_keys ? point_key.p : nullptr);
    if (n > 0 && one_sweep) {
      // one read of the keys for all digit histograms, then one launch per digit with chained tile prefixes
      const OsGather g = {d_pts, gather_dst, &meta.p->grid.overflow};
      if ((e = onesweep_sort<uint32_t>(st, onesweep, keys_a.p, vals_a.p, keys_b.p, vals_b.p, n, &meta.p->nbits, 4, gather_dst ? g : kNoGather)) != cudaSuccess) return e;
      if (gathered && gather_dst) *gathered = true;
    } else if (n > 0) {
      for (int pass = 0; pass < 4; ++pass) {
        const uint32_t* ki = (pass & 1) ? keys_b.p : keys_a.p;
        const uint32_t* vi = (pass & 1) ? vals_b.p : vals_a.p;
        uint32_t* ko = (pass & 1) ? keys_a.p : keys_b.p;
        uint32_t* vo = (pass & 1) ? vals_a.p : vals_b.p;
        switch (items) {
#define B200_SORT_PASS(IT)                                                                              \
  case IT:                                                                                              \
    k_sort_hist<IT><<<n_tiles, kSortThreads, 0, st>>>(ki, n, pass, meta.p, hist.p);                    \
    k_sort_scan<<<1, 1024, 0, st>>>(hist.p, n_tiles, pass, meta.p);                                  \
    k_sort_scatter<IT><<<n_tiles, kSortThreads, 0, st>>>(ki, vi, ko, vo, n, pass, meta.p, hist.p);     \
    break;
          B200_SORT_PASS(4)
          B200_SORT_PASS(8)
          B200_SORT_PASS(16)
          B200_SORT_PASS(32)
#undef B200_SORT_PASS
        }
      }
    }
    k_seg_count<<<n_seg_tiles, 256, 0, st>>>(keys_a.p, keys_b.p, n, meta.p, tile_heads.p, tile_valid.p);
    k_seg_scan<<<n_seg_tiles, 256, 0, st>>>(keys_a.p, keys_b.p, n, meta.p, tile_heads.p, tile_valid.p, n_seg_tiles, vox_start.p, vox_key.p, seg_first.p);
    return cudaGetLastError();
  }
};

}  // namespace b200

// b200reg — MapCloudGenerator::generate on the device [REF src/hdl_graph_slam/map_cloud_generator.cpp:13-49]:
// every keyframe cloud transformed by its pose and concatenated (:22-29), then reduced to the occupied voxel centres
// of a pcl::octree::OctreePointCloud at `resolution`, in the octree's own depth-first order (:36-42).
//
// What makes pcl's octree more than "floor(p / resolution), unique" is that its bounding box GROWS with the points in
// insertion order (adoptBoundingBoxToPoint): the box starts as one cell around the first point and doubles, towards the
// side the violating point lies on, until it holds it; a key is taken against the box of the moment, and the centres come
// out against the final box, so the lattice's anchor — and with it every centre — depends on the order of the points.
// The device path keeps exactly that:
//   k_map_transform     : pose * point per point (Eigen's Matrix4f * Vector4f operation order), one launch for all keyframes
//   k_map_first_outside : the next point (by index) that the current box does not hold — the only sequential part: the
//                         host grows the box for that point with the reference's scalar arithmetic and asks again;
//                         a map needs ~20-30 such rounds whatever its size (the box doubles every time)
//   k_map_keys          : per point the growth epoch it was inserted in (binary search over the <= 64 events), its key
//                         against THAT epoch's origin (the reference's double arithmetic), shifted into the final box, and
//                         the Morton code of the key (x most significant per level = the octree's child order)
//   one-sweep radix sort of the 64-bit codes (radix_onesweep.cuh, keys only, 3 * depth bits)
//   k_map_heads / k_map_scan / k_map_centres : distinct codes in ascending order -> voxel centres
//                         float((key + 0.5) * resolution + min), the reference's double expression
// Bit-identical to the CPU restatement of pcl's octree (oracle/oracle_capi.cpp orc_map_cloud), tests/test_gpu_map_cloud.py.
#pragma once
#include "common.cuh"
#include "radix_onesweep.cuh"

namespace b200 {

constexpr int kMapMaxEvents = 64;
constexpr int kMapMaxDepth = 21;  // 3 * 21 = 63 code bits
constexpr unsigned long long kMapInvalid = 0xFFFFFFFFFFFFFFFFull;

struct MapKeyframe {
  const float4* src;
  long long first;  // index of the keyframe's first point in the concatenated cloud
  float T[16];      // column-major pose
};

// dst = pose * src as Eigen evaluates Matrix4f * Vector4f: ((T_r0 x + T_r1 y) + T_r2 z) + T_r3 * 1
__global__ void __launch_bounds__(256) k_map_transform(const MapKeyframe* __restrict__ kfs, int n_kf, long long n, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_kf - 1;
  while (lo < hi) {  // last keyframe whose first point is <= i
    const int mid = (lo + hi + 1) >> 1;
    if (kfs[mid].first <= i) lo = mid;
    else hi = mid - 1;
  }
  const MapKeyframe& kf = kfs[lo];
  const float4 p = __ldg(kf.src + (i - kf.first));
  const float* T = kf.T;
  float4 q;
  q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], p.x), __fmul_rn(T[4], p.y)), __fmul_rn(T[8], p.z)), __fmul_rn(T[12], 1.0f));
  q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[1], p.x), __fmul_rn(T[5], p.y)), __fmul_rn(T[9], p.z)), __fmul_rn(T[13], 1.0f));
  q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[2], p.x), __fmul_rn(T[6], p.y)), __fmul_rn(T[10], p.z)), __fmul_rn(T[14], 1.0f));
  q.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[3], p.x), __fmul_rn(T[7], p.y)), __fmul_rn(T[11], p.z)), __fmul_rn(T[15], 1.0f));
  out[i] = q;
}

struct MapBox {
  double mn[3], mx[3];
  int defined;
};

// smallest index >= start whose (finite) point lies outside the box; *found starts at n
__global__ void __launch_bounds__(256) k_map_first_outside(const float4* __restrict__ pts, long long start, long long n, MapBox box, unsigned long long* found) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = start + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if ((unsigned long long)i >= *reinterpret_cast<volatile unsigned long long*>(found)) return;  // an earlier violation is already known
    const float4 p = __ldg(pts + i);
    if (!finite3(p.x, p.y, p.z)) continue;
    const bool out = !box.defined || (double)p.x < box.mn[0] || (double)p.x >= box.mx[0] || (double)p.y < box.mn[1] || (double)p.y >= box.mx[1] || (double)p.z < box.mn[2] ||
                     (double)p.z >= box.mx[2];
    if (out) {
      atomicMin(found, (unsigned long long)i);
      return;
    }
  }
}

struct MapEvents {
  int n;                              // epochs; epoch e holds the points with index in [first[e], first[e + 1])
  long long first[kMapMaxEvents + 1];
  double mn[kMapMaxEvents][3];
  unsigned long long shift[kMapMaxEvents][3];
  unsigned long long shift_final[3];
  double resolution;
  int depth;
};

__device__ __forceinline__ unsigned long long morton3(unsigned long long kx, unsigned long long ky, unsigned long long kz, int depth) {
  unsigned long long code = 0;
  for (int b = depth - 1; b >= 0; --b) code = (code << 3) | (((kx >> b) & 1ull) << 2) | (((ky >> b) & 1ull) << 1) | ((kz >> b) & 1ull);
  return code;
}

__global__ void __launch_bounds__(256) k_map_keys(const float4* __restrict__ pts, long long n, const MapEvents* __restrict__ evp, unsigned long long* __restrict__ codes) {
  __shared__ MapEvents ev;
  for (int w = threadIdx.x; w < (int)(sizeof(MapEvents) / 4); w += blockDim.x) reinterpret_cast<uint32_t*>(&ev)[w] = reinterpret_cast<const uint32_t*>(evp)[w];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(pts + i);
  if (!finite3(p.x, p.y, p.z)) {
    codes[i] = kMapInvalid;
    return;
  }
  int lo = 0, hi = ev.n - 1;
  while (lo < hi) {  // the epoch in force when point i was inserted
    const int mid = (lo + hi + 1) >> 1;
    if (ev.first[mid] <= i) lo = mid;
    else hi = mid - 1;
  }
  const float c[3] = {p.x, p.y, p.z};
  unsigned long long k[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // genOctreeKeyforPoint: static_cast<unsigned>((p - min_of_the_moment) / resolution), then the cells the origin moved since
    const unsigned int ka = (unsigned int)__ddiv_rn(__dsub_rn((double)c[a], ev.mn[lo][a]), ev.resolution);
    k[a] = (unsigned long long)ka + (ev.shift_final[a] - ev.shift[lo][a]);
  }
  codes[i] = morton3(k[0], k[1], k[2], ev.depth);
}

// heads of the runs of equal codes in the sorted order (invalid codes sort last and are never heads)
__global__ void __launch_bounds__(1024) k_map_heads(const unsigned long long* __restrict__ codes, long long n, uint32_t* __restrict__ tile_heads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0;
  if (i < n) {
    const unsigned long long c = codes[i];
    head = c != kMapInvalid && (i == 0 || codes[i - 1] != c);
  }
  const int h = __syncthreads_count(head);
  if (threadIdx.x == 0) tile_heads[blockIdx.x] = (uint32_t)h;
}

// exclusive scan of the per-tile head counts, one block walking the array in chunks; total[0] = number of voxels
__global__ void __launch_bounds__(1024) k_map_scan(uint32_t* __restrict__ tile_heads, long long n_tiles, unsigned long long* total) {
  __shared__ uint32_t s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < n_tiles; base += 1024) {
    const long long t = base + threadIdx.x;
    const uint32_t v = t < n_tiles ? tile_heads[t] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += u;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long carry = s_carry;
    const unsigned long long excl = carry + s_warp[warp] + incl - v;
    if (t < n_tiles) tile_heads[t] = (uint32_t)excl;  // maps are far below 2^32 voxels
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// genLeafNodeCenterFromOctreeKey: float((double(key) + 0.5f) * resolution + min)
__global__ void __launch_bounds__(1024) k_map_centres(const unsigned long long* __restrict__ codes, long long n, const uint32_t* __restrict__ tile_base, const MapEvents* __restrict__ evp,
                                                      float4* __restrict__ out, unsigned long long out_cap) {
  __shared__ uint32_t s_warp[32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long c = kMapInvalid;
  int head = 0;
  if (i < n) {
    c = codes[i];
    head = c != kMapInvalid && (i == 0 || codes[i - 1] != c);
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, head);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  uint32_t off = tile_base[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  if (!head) return;
  const unsigned long long dst = off + __popc(bal & ((1u << lane) - 1u));
  if (dst >= out_cap) return;
  const int depth = evp->depth;
  const double res = evp->resolution;
  unsigned long long k[3] = {0ull, 0ull, 0ull};
  for (int b = 0; b < depth; ++b) {
    k[2] |= ((c >> (3 * b)) & 1ull) << b;
    k[1] |= ((c >> (3 * b + 1)) & 1ull) << b;
    k[0] |= ((c >> (3 * b + 2)) & 1ull) << b;
  }
  const int last = evp->n - 1;
  float4 o;
  o.x = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k[0], 0.5), res), evp->mn[last][0]);
  o.y = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k[1], 0.5), res), evp->mn[last][1]);
  o.z = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)k[2], 0.5), res), evp->mn[last][2]);
  o.w = 1.0f;
  out[dst] = o;
}

}  // namespace b200

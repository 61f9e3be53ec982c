// b200reg — pcl::StatisticalOutlierRemoval on the device, the reference nodelet's DEFAULT outlier filter
// [REF apps/prefiltering_nodelet.cpp:77-87 (mean_k 20, stddev multiplier 1.0), applied :262-273].
//
//   k_gicp_knn<kKnnMeanDistance> (+ _brute) : per point the mean distance to its mean_k nearest neighbours (gicp.cuh:
//                        the same exact warp-per-query k-NN FAST_GICP's covariances use, k = mean_k + 1)
//   k_sor_threshold    : mean / standard deviation of those distances and the cut mean + mul * stddev
//   k_sor_flags        : keep[i] = !(distance_i > cut), per-block counts
//   k_ror_scatter      : the order-preserving scatter of nn_grid.cuh
//
// PCL sums the distances in INDEX order in double; a parallel sum rounds differently in the last bits,
// which could flip a point whose distance equals the cut to ~1e-11.  k_sor_threshold therefore sums in a
// fixed tree, bounds the difference to any other summation order (n * 2^-52 * conditioning), and only when
// some distance lies inside that band around the cut (practically never: the distances are floats, 6e-8
// apart) does one thread redo the sum in index order.  The filter is bit-exact either way.
#pragma once
#include "nn_grid.cuh"

namespace b200 {

struct SorStats {
  double mean, stddev, threshold;
  unsigned long long valid;   // points with a full neighbour list
  unsigned int exact_pass;    // 1 when the index-order sum had to run
  unsigned int ambiguous;     // distances found inside the band
};

__device__ __forceinline__ double sor_cut(double sum, double sq_sum, unsigned long long valid, double mul, double* mean_out, double* stddev_out, double* variance_out) {
  const double nv = (double)valid;
  const double mean = __ddiv_rn(sum, nv);
  const double variance = __ddiv_rn(__dsub_rn(sq_sum, __ddiv_rn(__dmul_rn(sum, sum), nv)), __dsub_rn(nv, 1.0));
  const double stddev = __dsqrt_rn(variance);
  *mean_out = mean; *stddev_out = stddev; *variance_out = variance;
  return __dadd_rn(mean, __dmul_rn(mul, stddev));
}

// dist[i] < 0: the point was not counted (non-finite, or fewer than k finite points) and stands at 0
__global__ void __launch_bounds__(1024) k_sor_threshold(const float* __restrict__ dist, int n, double stddev_mul, SorStats* __restrict__ out) {
  __shared__ double s_sum[32], s_sq[32];
  __shared__ unsigned long long s_cnt[32];
  __shared__ double s_cut, s_band;
  __shared__ unsigned int s_amb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0, sq = 0.0;
  unsigned long long cnt = 0;
  // one CTA, so the loads have to overlap inside each thread: 16-byte loads, four of them in flight
  const int n4 = n >> 2;
  const float4* __restrict__ dist4 = reinterpret_cast<const float4*>(dist);
  auto take = [&](float d) {
    if (d >= 0.f) { sum += (double)d; sq += (double)__fmul_rn(d, d); ++cnt; }
  };
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < n4 ? __ldg(dist4 + i) : make_float4(-1.f, -1.f, -1.f, -1.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { take(v[u].x); take(v[u].y); take(v[u].z); take(v[u].w); }
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) take(__ldg(dist + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { s_sum[warp] = sum; s_sq[warp] = sq; s_cnt[warp] = cnt; }
  if (threadIdx.x == 0) s_amb = 0u;
  __syncthreads();
  double mean = 0, stddev = 0, variance = 0;
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    unsigned long long c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_sum[w]; b += s_sq[w]; c += s_cnt[w]; }
    const double cut = sor_cut(a, b, c, stddev_mul, &mean, &stddev, &variance);
    // |cut - cut'| for any other summation order: both sums carry <= n u relative error (positive terms),
    // the variance amplifies it by (mean^2 + var) / var; a non-positive or non-finite variance has no bound
    double band = -1.0;  // < 0: undecidable here, take the exact pass
    if (c >= 2 && variance > 0.0 && isfinite(variance) && isfinite(cut)) {
      const double amp = (mean * mean + variance) / variance;
      band = 4.5e-16 * (double)n * (2.0 + amp) * (fabs(mean) + fabs(stddev_mul) * stddev);
    } else if (c < 2) {
      band = 0.0;  // 0 or 1 counted points: a single term, no order to differ in
    }
    s_cut = cut; s_band = band;
    out->mean = mean; out->stddev = stddev; out->threshold = cut; out->valid = c; out->exact_pass = 0u; out->ambiguous = 0u;
  }
  __syncthreads();
  const double cut = s_cut, band = s_band;
  unsigned int amb = 0;
  if (band > 0.0) {
    // the distances are floats: the band test runs on the two floats that bracket it (one compare each, the
    // double subtraction only decides for values that land exactly on a bracket)
    const float lo = __double2float_rd(cut - band), hi = __double2float_ru(cut + band);
    auto check = [&](float d) {
      d = fmaxf(d, 0.f);
      if (d >= lo && d <= hi && fabs((double)d - cut) <= band) ++amb;
    };
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < n4 ? __ldg(dist4 + i) : make_float4(-1.f, -1.f, -1.f, -1.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { check(v[u].x); check(v[u].y); check(v[u].z); check(v[u].w); }
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) check(__ldg(dist + i));
    if (amb) atomicAdd(&s_amb, amb);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (!(band < 0.0) && s_amb == 0u) return;
  // the reference's own loop, in index order
  double a = 0.0, b = 0.0;
  unsigned long long c = 0;
  for (int i = 0; i < n; ++i) {
    const float d = __ldg(dist + i);
    if (d >= 0.f) { a = __dadd_rn(a, (double)d); b = __dadd_rn(b, (double)__fmul_rn(d, d)); ++c; }
  }
  const double exact = sor_cut(a, b, c, stddev_mul, &mean, &stddev, &variance);
  out->mean = mean; out->stddev = stddev; out->threshold = exact; out->valid = c; out->exact_pass = 1u; out->ambiguous = s_amb;
}

__global__ void __launch_bounds__(256) k_sor_flags(const float* __restrict__ dist, int n, const SorStats* __restrict__ stats, unsigned char* __restrict__ keep, uint32_t* __restrict__ block_count) {
  __shared__ uint32_t s_cnt[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int k = 0;
  if (i < n) {
    const double cut = stats->threshold;
    const float d = fmaxf(__ldg(dist + i), 0.f);
    k = ((double)d > cut) ? 0 : 1;  // a NaN cut (no counted point) keeps everything, as upstream
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

// PrefilteringNodelet::normal_filtering's test [REF apps/prefiltering_nodelet.cpp:243-246]: keep when |n_z| < thresh
// (nz[i] = NaN for the points the height gate dropped and the points without a normal: never kept)
__global__ void __launch_bounds__(256) k_nz_flags(const float* __restrict__ nz, int n, float thresh, unsigned char* __restrict__ keep, uint32_t* __restrict__ block_count) {
  __shared__ uint32_t s_cnt[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int k = 0;
  if (i < n) {
    k = __ldg(nz + i) < thresh ? 1 : 0;
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

// PrefilteringNodelet::distance_filter on its own [REF apps/prefiltering_nodelet.cpp:275-291] (down-sampling NONE: there is
// no VoxelGrid key pass to fuse the gate into): keep when near < |p| < far, order kept by k_ror_scatter
__global__ void __launch_bounds__(256) k_gate_flags(const float4* __restrict__ pts, int n, PointGate gate, unsigned char* __restrict__ keep, uint32_t* __restrict__ block_count) {
  __shared__ uint32_t s_cnt[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int k = 0;
  if (i < n) {
    const float4 p = __ldg(pts + i);
    k = point_takes_part(gate, 0, p.x, p.y, p.z) ? 1 : 0;
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

}  // namespace b200

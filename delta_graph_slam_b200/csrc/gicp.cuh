// b200reg — FAST_GICP on the device: the B200 replacement of fast_gicp::FastGICP on
// fast_gicp::LsqRegistration as configured by the reference's factory
// [REF src/hdl_graph_slam/registrations.cpp:27-36] (k = 20 neighbours, PLANE regularisation,
// Levenberg-Marquardt, max correspondence distance 2.5 m / 2.0 m in the launch files).
//
//   k_gicp_knn / k_gicp_knn_brute / k_gicp_regularize : calculate_covariances — per point the k nearest neighbours (exact, found
//                        on the cloud's own cell grid), their covariance in double, and the
//                        regularised 3x3 (stored as 6 doubles)
//   k_gicp_align       : LsqRegistration::computeTransformation — one persistent cooperative
//                        kernel per registration.  A pass is either `linearize` (1-NN correspondence
//                        of every transformed source point inside the distance gate, fused
//                        Mahalanobis matrix (C_B + R C_A R^T)^-1, residual, H = sum J^T M J,
//                        b = sum J^T M e) or `compute_error` (residual sum with the stored
//                        correspondences and Mahalanobis matrices); between passes one lane of every
//                        CTA runs the LM / GN state machine redundantly on identical totals, exactly
//                        like the NDT kernel, so an iteration needs no host round trip.
// Arithmetic is double throughout except the point storage and the NN search (float), as upstream.
#pragma once
#include "../../include/b200reg.h"
#include "nn_grid.cuh"
#include "small_solve.cuh"

namespace b200 {

// ---- k nearest neighbours ---------------------------------------------------------------------
// One WARP per query.  The running k-best list (k <= 32) is held distributed over the lanes: lane l
// owns the l-th smallest (d2, index) so far.  The lanes probe 32 cells of the current Chebyshev ring
// at once (one 16-byte hash load each); the points of every non-empty cell are then taken 32 at a
// time, one per lane, and merged into the list with a bitonic network (sort the 32 candidates,
// min against the reversed list, re-sort: ~20 shuffle steps for 32 candidates).  Batches in which no
// candidate beats the current k-th are skipped with one ballot.  Rings grow until nothing outside
// the examined block can beat the k-th (or the block covers the lattice); a query still open after
// kKnnMaxRing rings (an isolated point) restarts as a linear scan of the whole cloud.  Exact; ties
// by lowest index, as the oracle's kd-tree.
constexpr int kKnnMaxRing = 15;  // 7.5 m; the rings from 2 on are walked through the occupancy bitmap (32 cells of a row per word)

// one compare-exchange step of a bitonic network over the lanes: partner = lane ^ j; in an ascending
// block the lower lane of the pair keeps the smaller entry.  (d, i) pairs are totally ordered.
__device__ __forceinline__ void bitonic_cmpex(float& d, int& i, int lane, int j, bool up) {
  // (d, i) as ONE 64-bit key: squared distances are >= +0, so their float bits order like the values, and
  // indices are non-negative — the lexicographic (d, then lowest index) order of nn_better is a single
  // unsigned compare.  The networks were half of the k-NN kernel's instructions with the two-field compare.
  const unsigned long long me = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)i;
  const unsigned long long other = __shfl_xor_sync(0xffffffffu, me, j);
  const bool keep_min = ((lane & j) == 0) == up;
  const bool take = keep_min ? (other < me) : (me < other);
  if (take) { d = __uint_as_float((unsigned)(other >> 32)); i = (int)(unsigned)(other & 0xffffffffull); }
}

// ascending bitonic sort of one (d, i) per lane
__device__ __forceinline__ void warp_sort32(float& d, int& i, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
    const bool up = (lane & k) == 0 || k == 32;
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) bitonic_cmpex(d, i, lane, j, up);
  }
}

// merge 32 candidates (one per lane, +inf when absent) into the distributed sorted list
__device__ __forceinline__ void knn_merge32(float& td, int& ti, float cd, int ci, int lane) {
  warp_sort32(cd, ci, lane);
  const float rd = __shfl_sync(0xffffffffu, cd, 31 - lane);
  const int ri = __shfl_sync(0xffffffffu, ci, 31 - lane);
  if (nn_better(rd, ri, td, ti)) { td = rd; ti = ri; }  // the 32 smallest of the 64, as a bitonic sequence
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) bitonic_cmpex(td, ti, lane, j, true);
}

// The k-best list of one warp plus a staging row in shared memory.  A merge costs ~20 shuffle steps,
// so candidates are FILTERED against the current k-th first (one ballot per 32 points) and the
// survivors are packed into the staging row; the network runs only when 32 of them have gathered or
// the caller needs the exact k-th (flush).  After the query's own cell has filled the list few points
// of the neighbouring cells survive, so a query costs ~2 merges instead of one per occupied cell.
struct KnnList {
  float td;  // lane l: the l-th smallest squared distance so far
  int ti;
  float kd;  // the current k-th (warp-uniform); (+inf, kNoIndex) while the list is not full
  int ki;
  int ns;    // staged candidates (warp-uniform, < 32 between calls)
  float* bd; // staging row of this warp, 64 entries
  int* bi;
};

__device__ __forceinline__ void knn_init(KnnList& L, float* bd, int* bi) {
  L.td = 3.402823466e+38f; L.ti = kNoIndex; L.kd = 3.402823466e+38f; L.ki = kNoIndex; L.ns = 0; L.bd = bd; L.bi = bi;
}
__device__ __forceinline__ void knn_merge_staged(KnnList& L, int km1, int lane) {
  __syncwarp();
  const float cd = lane < L.ns ? L.bd[lane] : 3.402823466e+38f;
  const int ci = lane < L.ns ? L.bi[lane] : kNoIndex;
  knn_merge32(L.td, L.ti, cd, ci, lane);
  L.kd = __shfl_sync(0xffffffffu, L.td, km1);
  L.ki = __shfl_sync(0xffffffffu, L.ti, km1);
  const int rest = L.ns > 32 ? L.ns - 32 : 0;
  float rd = 0.f;
  int ri = 0;
  if (lane < rest) { rd = L.bd[32 + lane]; ri = L.bi[32 + lane]; }
  __syncwarp();
  if (lane < rest) { L.bd[lane] = rd; L.bi[lane] = ri; }
  L.ns = rest;
  __syncwarp();
}
__device__ __forceinline__ void knn_flush(KnnList& L, int km1, int lane) {
  if (L.ns > 0) knn_merge_staged(L, km1, lane);
}
// take one candidate per lane (valid = false: none): filter against the k-th, pack, merge when 32 gathered
__device__ __forceinline__ void knn_take(KnnList& L, bool valid, float cd, int ci, int km1, int lane) {
  const bool pass = valid && nn_better(cd, ci, L.kd, L.ki);
  const unsigned mask = __ballot_sync(0xffffffffu, pass);
  if (mask == 0u) return;
  if (pass) {
    const int slot = L.ns + __popc(mask & ((1u << lane) - 1u));
    L.bd[slot] = cd;
    L.bi[slot] = ci;
  }
  L.ns += __popc(mask);
  if (L.ns >= 32) knn_merge_staged(L, km1, lane);
}

// offer the points [s, e) of the cell-ordered array (the brute-force slices): eight loads in flight per lane
// (a CTA works on one query and the kernel is pure load latency: ncu long_scoreboard 39 %, 12 % warps active)
__device__ __forceinline__ void knn_offer_range(const NnView& g, uint32_t s, uint32_t e, float qx, float qy, float qz, int km1, int lane, KnnList& L) {
  constexpr int kInFlight = 8;
  for (uint32_t j0 = s; j0 < e; j0 += 32u * kInFlight) {
    float4 p[kInFlight];
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t j = j0 + 32u * u + lane;
      p[u] = j < e ? __ldg(g.pts + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t j = j0 + 32u * u + lane;
      knn_take(L, j < e, l2_simple(qx, qy, qz, p[u].x, p[u].y, p[u].z), __float_as_int(p[u].w), km1, lane);
    }
  }
}

// offer the runs the lanes hold (run.x .. run.y, empty when equal), FLATTENED: the points of all runs
// are numbered consecutively, batch b takes points 32 b .. 32 b + 31 whatever run they belong to (the
// owner of a point is found by a 5-step search over the shuffled prefix sums), and the load of batch
// b + 1 is issued before batch b is filtered — full batches and two loads in flight instead of one
// short, dependent load per occupied cell.
__device__ __forceinline__ void knn_offer_runs(const NnView& g, uint2 run, float qx, float qy, float qz, int km1, int lane, KnnList& L) {
  const uint32_t len = run.y - run.x;
  uint32_t incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0u) return;
  const uint32_t first = run.x - (incl - len);  // point t of the flattened order owned by this lane sits at first + t
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
    knn_take(L, ok_cur, l2_simple(qx, qy, qz, p_cur.x, p_cur.y, p_cur.z), __float_as_int(p_cur.w), km1, lane);
    p_cur = p_next;
    ok_cur = ok_next;
    ok_next = false;
  }
}

// neighbours as columns, minus the row-wise mean over the k REQUESTED columns, cov = N N^T / k
// (calculate_covariances, A.5).  Lanes 0..5 return xx, xy, xz, yy, yz, zz in c.
__device__ __forceinline__ double knn_raw_covariance(const float4* __restrict__ pts, int k, int lane, float td, int ti) {
  (void)td;
  const bool have = lane < k && ti != kNoIndex;
  float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (have) nb = __ldg(pts + ti);
  const double kd = (double)k;
  double m0 = have ? (double)nb.x : 0.0, m1 = have ? (double)nb.y : 0.0, m2 = have ? (double)nb.z : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 += __shfl_xor_sync(0xffffffffu, m0, o);
    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
    m2 += __shfl_xor_sync(0xffffffffu, m2, o);
  }
  m0 /= kd; m1 /= kd; m2 /= kd;
  const double v0 = have ? (double)nb.x - m0 : 0.0, v1 = have ? (double)nb.y - m1 : 0.0, v2 = have ? (double)nb.z - m2 : 0.0;
  double c6[6] = {v0 * v0, v0 * v1, v0 * v2, v1 * v1, v1 * v2, v2 * v2};
  double mine = 0.0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c6[a] += __shfl_xor_sync(0xffffffffu, c6[a], o);
    c6[a] /= kd;
    if (lane == a) mine = c6[a];
  }
  return mine;
}

// pcl::StatisticalOutlierRemoval's per-point figure [REF apps/prefiltering_nodelet.cpp:77-87]: the list holds the
// k = mean_k + 1 nearest of the cloud itself (entry 0 is the point, or a duplicate of it, at distance 0);
// distance = float( (sqrtf(d2_1) + ... + sqrtf(d2_mean_k)) [double, nearest first] / mean_k ).  A list that
// never filled (fewer than k finite points in the cloud) leaves the point uncounted: -1.
__device__ __forceinline__ float knn_mean_distance(int k, int lane, float td, int ti) {
  const bool full = __shfl_sync(0xffffffffu, ti, k - 1) != kNoIndex;
  const float s = __fsqrt_rn(lane < k ? td : 0.f);
  double sum = 0.0;
  for (int j = 1; j < k; ++j) sum = __dadd_rn(sum, (double)__shfl_sync(0xffffffffu, s, j));
  return full ? (float)__ddiv_rn(sum, (double)(k - 1)) : -1.0f;
}

// pcl::NormalEstimation's per-point normal, reduced to what PrefilteringNodelet::normal_filtering keeps of it
// [REF apps/prefiltering_nodelet.cpp:222-251]: |n_z| of the normalised normal of the k nearest neighbours (the point
// included).  PCL 1.8-1.10 [UPSTREAM-RECALLED, mirrored operation by operation, no contraction]:
// computeMeanAndCovarianceMatrix = nine FLOAT accumulators in neighbour order, / count, cov = E[ab] - E[a]E[b];
// pcl::eigen33 = scale by the largest |entry|, closed-form roots (atan2 / cos / sin: computed in double and rounded,
// the correctly rounded float), eigenvector of the smallest root = the longest cross product of two rows of A - root I.
// flipNormalTowardsViewpoint only changes the sign.  Fewer than three neighbours: no normal (NaN), as upstream.
__device__ __forceinline__ void pcl_roots2(float b, float c, float* r) {
  r[0] = 0.0f;
  float d = __fsub_rn(__fmul_rn(b, b), __fmul_rn(4.0f, c));
  if (d < 0.0f) d = 0.0f;
  const float sd = __fsqrt_rn(d);
  r[2] = __fmul_rn(0.5f, __fadd_rn(b, sd));
  r[1] = __fmul_rn(0.5f, __fsub_rn(b, sd));
}
__device__ __forceinline__ float knn_normal_abs_nz(const float4* __restrict__ pts, int k, int lane, float td, int ti) {
  (void)td;
  const bool have = lane < k && ti != kNoIndex;
  const int cnt = __popc(__ballot_sync(0xffffffffu, have));  // the list is packed: entries 0 .. cnt-1
  float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (have) nb = __ldg(pts + ti);
  if (cnt < 3) return __int_as_float(0x7fc00000);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f, a8 = 0.f;
  for (int j = 0; j < cnt; ++j) {
    const float x = __shfl_sync(0xffffffffu, nb.x, j), y = __shfl_sync(0xffffffffu, nb.y, j), z = __shfl_sync(0xffffffffu, nb.z, j);
    a0 = __fadd_rn(a0, __fmul_rn(x, x)); a1 = __fadd_rn(a1, __fmul_rn(x, y)); a2 = __fadd_rn(a2, __fmul_rn(x, z));
    a3 = __fadd_rn(a3, __fmul_rn(y, y)); a4 = __fadd_rn(a4, __fmul_rn(y, z)); a5 = __fadd_rn(a5, __fmul_rn(z, z));
    a6 = __fadd_rn(a6, x); a7 = __fadd_rn(a7, y); a8 = __fadd_rn(a8, z);
  }
  const float fc = (float)cnt;
  a0 = __fdiv_rn(a0, fc); a1 = __fdiv_rn(a1, fc); a2 = __fdiv_rn(a2, fc); a3 = __fdiv_rn(a3, fc); a4 = __fdiv_rn(a4, fc);
  a5 = __fdiv_rn(a5, fc); a6 = __fdiv_rn(a6, fc); a7 = __fdiv_rn(a7, fc); a8 = __fdiv_rn(a8, fc);
  float c00 = __fsub_rn(a0, __fmul_rn(a6, a6)), c01 = __fsub_rn(a1, __fmul_rn(a6, a7)), c02 = __fsub_rn(a2, __fmul_rn(a6, a8));
  float c11 = __fsub_rn(a3, __fmul_rn(a7, a7)), c12 = __fsub_rn(a4, __fmul_rn(a7, a8)), c22 = __fsub_rn(a5, __fmul_rn(a8, a8));
  float scale = fmaxf(fmaxf(fmaxf(fabsf(c00), fabsf(c01)), fmaxf(fabsf(c02), fabsf(c11))), fmaxf(fabsf(c12), fabsf(c22)));
  if (!(scale == scale)) return __int_as_float(0x7fc00000);
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float m00 = __fdiv_rn(c00, scale), m01 = __fdiv_rn(c01, scale), m02 = __fdiv_rn(c02, scale), m11 = __fdiv_rn(c11, scale), m12 = __fdiv_rn(c12, scale), m22 = __fdiv_rn(c22, scale);
  // computeRoots
  float r[3];
  {
    const float t1 = __fmul_rn(__fmul_rn(m00, m11), m22), t2 = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, m01), m02), m12), t3 = __fmul_rn(__fmul_rn(m00, m12), m12);
    const float t4 = __fmul_rn(__fmul_rn(m11, m02), m02), t5 = __fmul_rn(__fmul_rn(m22, m01), m01);
    const float c0 = __fsub_rn(__fsub_rn(__fsub_rn(__fadd_rn(t1, t2), t3), t4), t5);
    const float c1 = __fsub_rn(__fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m01)), __fmul_rn(m00, m22)), __fmul_rn(m02, m02)), __fmul_rn(m11, m22)), __fmul_rn(m12, m12));
    const float c2 = __fadd_rn(__fadd_rn(m00, m11), m22);
    if (fabsf(c0) < 1.1920929e-07f) {
      pcl_roots2(c2, c1, r);
    } else {
      const float s_inv3 = 0.333333343f, s_sqrt3 = 1.73205078f;  // 1.0f / 3.0f, sqrtf(3.0f)
      const float c2_over_3 = __fmul_rn(c2, s_inv3);
      float a_over_3 = __fmul_rn(__fsub_rn(c1, __fmul_rn(c2, c2_over_3)), s_inv3);
      if (a_over_3 > 0.0f) a_over_3 = 0.0f;
      const float half_b = __fmul_rn(0.5f, __fadd_rn(c0, __fmul_rn(c2_over_3, __fsub_rn(__fmul_rn(__fmul_rn(2.0f, c2_over_3), c2_over_3), c1))));
      float q = __fadd_rn(__fmul_rn(half_b, half_b), __fmul_rn(__fmul_rn(a_over_3, a_over_3), a_over_3));
      if (q > 0.0f) q = 0.0f;
      const float rho = __fsqrt_rn(-a_over_3);
      const float theta = __fmul_rn((float)atan2((double)__fsqrt_rn(-q), (double)half_b), s_inv3);
      const float cos_theta = (float)cos((double)theta), sin_theta = (float)sin((double)theta);
      r[0] = __fadd_rn(c2_over_3, __fmul_rn(__fmul_rn(2.0f, rho), cos_theta));
      r[1] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fadd_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
      r[2] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fsub_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
      float t;
      if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
      if (r[1] >= r[2]) {
        t = r[1]; r[1] = r[2]; r[2] = t;
        if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
      }
      if (r[0] <= 0.0f) pcl_roots2(c2, c1, r);
    }
  }
  m00 = __fsub_rn(m00, r[0]); m11 = __fsub_rn(m11, r[0]); m22 = __fsub_rn(m22, r[0]);
  // rows: (m00 m01 m02) (m01 m11 m12) (m02 m12 m22)
#define B200_CROSS(ax, ay, az, bx, by, bz, ox, oy, oz) \
  const float ox = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by)), oy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz)), oz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
  B200_CROSS(m00, m01, m02, m01, m11, m12, u0, u1, u2)
  B200_CROSS(m00, m01, m02, m02, m12, m22, v0, v1, v2)
  B200_CROSS(m01, m11, m12, m02, m12, m22, w0, w1, w2)
#undef B200_CROSS
  const float l1 = __fadd_rn(__fadd_rn(__fmul_rn(u0, u0), __fmul_rn(u1, u1)), __fmul_rn(u2, u2));
  const float l2 = __fadd_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)), __fmul_rn(v2, v2));
  const float l3 = __fadd_rn(__fadd_rn(__fmul_rn(w0, w0), __fmul_rn(w1, w1)), __fmul_rn(w2, w2));
  float e0 = w0, e1 = w1, e2 = w2, l = l3;
  if (l1 >= l2 && l1 >= l3) { e0 = u0; e1 = u1; e2 = u2; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { e0 = v0; e1 = v1; e2 = v2; l = l2; }
  const float sl = __fsqrt_rn(l);
  e0 = __fdiv_rn(e0, sl); e1 = __fdiv_rn(e1, sl); e2 = __fdiv_rn(e2, sl);
  const float z2 = __fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2));  // Eigen normalized()
  const float nz = z2 > 0.0f ? __fdiv_rn(e2, __fsqrt_rn(z2)) : e2;
  return fabsf(nz);
}

// The two k-NN kernels serve two callers; kKnnCovariance is FAST_GICP's, kKnnMeanDistance the statistical
// outlier filter's (same search, `n` taken from the grid's own count of finite points, a float per point out).
// kKnnNormalNz the flat-cloud normal filter's (|n_z| per point into the same float array).
constexpr int kKnnCovariance = 0, kKnnMeanDistance = 1, kKnnNormalNz = 2;

// calculate_covariances, step 1: exact k nearest neighbours + raw covariance, one warp per point.
// covs[i] = {xx, xy, xz, yy, yz, zz} (not yet regularised).  A point whose k-th neighbour is not settled
// within kKnnMaxRing rings (an isolated return) goes to `pending` for the block-per-query scan.
template <int TAIL>
__global__ void __launch_bounds__(256) k_gicp_knn(NnView g, const float4* __restrict__ pts, int n, int k, double* __restrict__ covs, int* __restrict__ pending,
                                                  unsigned int* __restrict__ n_pending, float* __restrict__ mean_dist) {
  __shared__ float s_bd[8][64];
  __shared__ int s_bi[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ GridParams s_gp;  // the cloud's cell lattice, shared by the CTA instead of 20 registers per thread
  if (threadIdx.x < (int)(sizeof(GridParams) / 4)) reinterpret_cast<uint32_t*>(&s_gp)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&g.meta->grid)[threadIdx.x];
  __syncthreads();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // queries are taken in cell order: neighbouring warps touch neighbouring cells
  if (TAIL != kKnnCovariance) n = min(n, (int)g.meta->n_valid);  // non-finite points never entered the grid
  if (w >= n) return;
  const GridParams& gp = s_gp;
  const float4 qp = __ldg(g.pts + w);
  const int qi = __float_as_int(qp.w);
  const float qx = qp.x, qy = qp.y, qz = qp.z;
  const NnQuery q = nn_make_query(gp, qx, qy, qz);
  const int km1 = k - 1;
  KnnList L;
  knn_init(L, s_bd[warp], s_bi[warp]);
  bool open = true;
  const bool use_occ = nn_occ_valid(g);
  for (int r = 0; r <= kKnnMaxRing && open; ++r) {
    knn_flush(L, km1, lane);  // the settle test needs the exact k-th
    if (r >= 1 && nn_settled(gp, q, r, L.ki == kNoIndex ? 3.402823466e+38f : L.kd, 3.402823466e+38f)) { open = false; break; }
    const int side = 2 * r + 1, inner = 2 * r - 1;
    if (use_occ && r >= 2 && side <= 32) {
      // outer rings row by row through the occupancy bitmap (see nn_query_far_warp): a lane takes a row
      // of the shell, keeps the bits of its occupied shell cells, and the warp then consumes one cell
      // per lane and round until no lane has a bit left
      const int x0 = max(q.cx - r, gp.min_b[0]), x1 = min(q.cx + r, gp.max_b[0]);
      if (x0 > x1) continue;
      const int nb = x1 - x0 + 1;
      uint32_t ends = 0u;
      if (q.cx - r >= gp.min_b[0]) ends |= 1u;
      if (q.cx + r <= gp.max_b[0]) ends |= 1u << (nb - 1);
      const int xq = min(max(q.cx, gp.min_b[0]), gp.max_b[0]);
      for (int row0 = 0; row0 < side * side; row0 += 32) {
        const int row = row0 + lane;
        uint32_t bits = 0u, key0 = 0u;
        int iy = 0, iz = 0;
        if (row < side * side) {
          const int dz = row / side - r, dy = row - (dz + r) * side - r;
          iy = q.cy + dy; iz = q.cz + dz;
          const bool in = iy >= gp.min_b[1] && iy <= gp.max_b[1] && iz >= gp.min_b[2] && iz <= gp.max_b[2];
          if (in && (L.ki == kNoIndex || nn_box_d2(gp, q, xq, iy, iz) <= L.kd)) {
            key0 = (uint32_t)((x0 - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
            bits = nn_occ_bits(g, key0, nb);
            if (!(dy == -r || dy == r || dz == -r || dz == r)) bits &= ends;
          }
        }
        while (__any_sync(0xffffffffu, bits != 0u)) {
          uint2 run = make_uint2(0u, 0u);
          if (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1u;
            if (L.ki == kNoIndex || nn_box_d2(gp, q, x0 + b, iy, iz) <= L.kd) run = nn_lookup(g, key0 + (uint32_t)b * (uint32_t)gp.mul[0]);
          }
          knn_offer_runs(g, run, qx, qy, qz, km1, lane, L);
        }
      }
      continue;
    }
    const int nz = r == 0 ? 1 : 2 * side * side, ny = r == 0 ? 0 : 2 * side * inner, total = r == 0 ? 1 : nz + ny + 2 * inner * inner;
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      uint2 run = make_uint2(0u, 0u);
      if (c < total) {
        int dx = 0, dy = 0, dz = 0;
        if (r > 0) {
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
        }
        const int ix = q.cx + dx, iy = q.cy + dy, iz = q.cz + dz;
        const bool in = ix >= gp.min_b[0] && ix <= gp.max_b[0] && iy >= gp.min_b[1] && iy <= gp.max_b[1] && iz >= gp.min_b[2] && iz <= gp.max_b[2];
        if (in && (L.ki == kNoIndex || nn_box_d2(gp, q, ix, iy, iz) <= L.kd)) {
          const uint32_t key = (uint32_t)((ix - gp.min_b[0]) * gp.mul[0] + (iy - gp.min_b[1]) * gp.mul[1] + (iz - gp.min_b[2]) * gp.mul[2]);
          if (!use_occ || nn_occ_bit(g, key)) run = nn_lookup(g, key);  // most cells of a ring are empty: one bitmap bit instead of a hash walk
        }
      }
      knn_offer_runs(g, run, qx, qy, qz, km1, lane, L);
    }
  }
  if (open) {
    knn_flush(L, km1, lane);
    if (!nn_settled(gp, q, kKnnMaxRing + 1, L.ki == kNoIndex ? 3.402823466e+38f : L.kd, 3.402823466e+38f)) {
      if (lane == 0) pending[atomicAdd(n_pending, 1u)] = w;
      return;
    }
  }
  if (TAIL == kKnnMeanDistance) {
    const float d = knn_mean_distance(k, lane, L.td, L.ti);
    if (lane == 0) mean_dist[qi] = d;
    return;
  }
  if (TAIL == kKnnNormalNz) {
    const float a = knn_normal_abs_nz(pts, k, lane, L.td, L.ti);
    if (lane == 0) mean_dist[qi] = a;
    return;
  }
  const double c = knn_raw_covariance(pts, k, lane, L.td, L.ti);
  if (lane < 6) covs[(size_t)qi * 6 + lane] = c;
}

// step 1b: one CTA per isolated point: every one of its sixteen warps scans its own sixteenth of the cloud with the same
// filtered list, then the sixteen lists are merged pairwise in four rounds (a tree: warps 0, 2, 4, .. take their right
// neighbour's list, then 0, 4, 8, .. and so on).  Round 1 of this kernel ran eight warps and let warp 0 merge the other
// seven lists one after the other: one query per CTA is pure latency (ncu: long_scoreboard 39 %, 12 % warps active), so
// halving the scan per warp and turning seven serial merges into four rounds is worth ~2x on the ~100 isolated returns of
// a scan (65-85 us of a 0.41 ms statistical filter call, and the same for FAST_GICP's covariances and the flat filter).
constexpr int kBruteWarps = 16;
template <int TAIL>
__global__ void __launch_bounds__(kBruteWarps * 32) k_gicp_knn_brute(NnView g, const float4* __restrict__ pts, int k, double* __restrict__ covs, const int* __restrict__ pending,
                                                                     const unsigned int* __restrict__ n_pending, float* __restrict__ mean_dist) {
  __shared__ float s_bd[kBruteWarps][64];
  __shared__ int s_bi[kBruteWarps][64];
  __shared__ float s_ld[kBruteWarps][32];
  __shared__ int s_li[kBruteWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = (int)*n_pending;
  const int km1 = k - 1;
  const uint32_t gn = TAIL != kKnnCovariance ? min((uint32_t)g.n, g.meta->n_valid) : (uint32_t)g.n;
  const uint32_t chunk = ((gn + (uint32_t)kBruteWarps - 1u) / (uint32_t)kBruteWarps + 31u) & ~31u;
  for (int e = blockIdx.x; e < np; e += gridDim.x) {
    const int w = pending[e];
    const float4 qp = __ldg(g.pts + w);
    KnnList L;
    knn_init(L, s_bd[warp], s_bi[warp]);
    const uint32_t s0 = min(gn, chunk * (uint32_t)warp), s1 = min(gn, s0 + chunk);
    knn_offer_range(g, s0, s1, qp.x, qp.y, qp.z, km1, lane, L);
    knn_flush(L, km1, lane);
    __syncthreads();  // the previous query's lists have been consumed
    // pairwise merge tree; the k smallest of a union do not depend on the order of the merges (ties by index)
    for (int stride = 1; stride < kBruteWarps; stride <<= 1) {
      if ((warp & (2 * stride - 1)) == stride) {  // a right-hand list of this round: publish it
        s_ld[warp][lane] = L.td;
        s_li[warp][lane] = L.ti;
      }
      __syncthreads();
      if ((warp & (2 * stride - 1)) == 0) knn_merge32(L.td, L.ti, s_ld[warp + stride][lane], s_li[warp + stride][lane], lane);
      __syncthreads();
    }
    if (warp == 0) {
      if (TAIL == kKnnMeanDistance) {
        const float d = knn_mean_distance(k, lane, L.td, L.ti);
        if (lane == 0) mean_dist[__float_as_int(qp.w)] = d;
        continue;
      }
      if (TAIL == kKnnNormalNz) {
        const float a = knn_normal_abs_nz(pts, k, lane, L.td, L.ti);
        if (lane == 0) mean_dist[__float_as_int(qp.w)] = a;
        continue;
      }
      const double c = knn_raw_covariance(pts, k, lane, L.td, L.ti);
      if (lane < 6) covs[(size_t)__float_as_int(qp.w) * 6 + lane] = c;
    }
  }
}

// step 2: regularisation, one thread per point, in place (thirty-two eigen-decompositions per warp
// instead of one lane of a warp each)
__global__ void __launch_bounds__(128) k_gicp_regularize(int n, int reg_method, double* __restrict__ covs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* o = covs + (size_t)i * 6;
  const double c6[6] = {o[0], o[1], o[2], o[3], o[4], o[5]};
  const double c[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
  double out[9];
  if (reg_method == B200REG_REG_NONE) {
    return;
  } else if (reg_method == B200REG_REG_FROBENIUS) {
    const double lambda = 1e-3;
    double C[9], Ci[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) C[a] = c[a] + ((a % 4 == 0) ? lambda : 0.0);
    inverse3(C, Ci);
    double fro = 0;
#pragma unroll
    for (int a = 0; a < 9; ++a) fro += Ci[a] * Ci[a];
    fro = sqrt(fro);
#pragma unroll
    for (int a = 0; a < 9; ++a) Ci[a] /= fro;
    inverse3(Ci, out);
  } else {
    // JacobiSVD of a symmetric PSD 3x3: U = V = eigenvectors, singular values descending
    double ev[3], V[9];
    sym_eigen3(c, ev, V);  // ascending, eigenvectors in columns
    double values[3];
    if (reg_method == B200REG_REG_PLANE) {
      values[0] = 1.0; values[1] = 1.0; values[2] = 1e-3;
    } else if (reg_method == B200REG_REG_MIN_EIG) {
      values[0] = fmax(fabs(ev[2]), 1e-3); values[1] = fmax(fabs(ev[1]), 1e-3); values[2] = fmax(fabs(ev[0]), 1e-3);
    } else {
      const double smax = fabs(ev[2]);
      values[0] = fmax(fabs(ev[2]) / smax, 1e-3); values[1] = fmax(fabs(ev[1]) / smax, 1e-3); values[2] = fmax(fabs(ev[0]) / smax, 1e-3);
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) out[a] = 0.0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int col = 2 - s;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) out[3 * a + b] += values[s] * V[3 * a + col] * V[3 * b + col];
    }
  }
  o[0] = out[0]; o[1] = out[1]; o[2] = out[2]; o[3] = out[4]; o[4] = out[5]; o[5] = out[8];
}

// ---- the alignment kernel -------------------------------------------------------------------------
constexpr int kGicpThreads = 512;
constexpr int kGicpWarps = kGicpThreads / 32;
constexpr int kGicpAcc = 29;  // H upper triangle [21], b [6], sum of errors, correspondences
constexpr int kGicpStride = 32;

enum GicpPhase { GP_LINEARIZE = 0, GP_ERROR = 1, GP_DONE = 2 };

struct GicpParams {
  double corr_dist2;  // corr_dist_threshold_^2 (double)
  float search_d2;    // float bound handed to the NN search (>= corr_dist2)
  int far_ring;       // rings that cover the gate (capped at kFarRing; beyond that brute force)
  double trans_eps, rot_eps;
  int max_iterations, lsq, lm_max_iterations;
  double lm_init_lambda_factor;
  int query_lanes;    // lanes that share one source point's near search in a linearize pass: 1, 2 or 4
};

struct GicpJob {
  const float4* src;
  int n_src;
  const double* cov_src;   // [n_src][6]
  NnView tgt;              // exact-NN structure of the target
  const float4* tgt_pts;   // target in original order
  const double* cov_tgt;   // [n_tgt][6]
  float guess[16];         // column-major
  int* corr;               // [n_src] scratch: correspondences_
  double* mahal;           // [n_src][6] scratch: mahalanobis_
  b200reg_result* result;
  b200reg_result* result_host;  // mapped host copy + completion flag (see NdtJob)
  unsigned int* done_flag;
  unsigned int done_seq;
  long long* prof;         // optional developer counters of CTA 0: cycles in {near phase, far queue, error pass, block reduce, group barrier, row sum, step}, linearize passes, error passes, far queries of CTA 0
};

struct GicpShared {
  double R[9], t[3];    // pose of the current pass
  float Tf[12];         // the same, cast to float (row-major 3x4)
  double x0R[9], x0t[3];
  double H[36], b[6], d[6], dR[9], dt[3];
  double y0, lambda, nu, last_err;
  int phase, lm_k, iter, nr_iterations, converged, n_lin, n_err;
  double hits;
  double tot[kGicpStride];
  double red[kGicpWarps][kGicpStride];
};

__device__ inline void iso_mul(const double* aR, const double* at, const double* bR, const double* bt, double* cR, double* ct) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cR[3 * i + j] = aR[3 * i] * bR[j] + aR[3 * i + 1] * bR[3 + j] + aR[3 * i + 2] * bR[6 + j];
    ct[i] = aR[3 * i] * bt[0] + aR[3 * i + 1] * bt[1] + aR[3 * i + 2] * bt[2] + at[i];
  }
}

// so3_exp / se3_exp of fast_gicp/so3/so3.hpp (A.5): a = [rotation(3), translation(3)]
__device__ inline void se3_exp(const double a[6], double R[9], double t[3]) {
  const double w0 = a[0], w1 = a[1], w2 = a[2];
  const double theta_sq = w0 * w0 + w1 * w1 + w2 * w2;
  const double theta = sqrt(theta_sq);
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  const double qw = real, qx = imag * w0, qy = imag * w1, qz = imag * w2;
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
  double V[9];
  if (theta < 1e-10) {
    for (int k = 0; k < 9; ++k) V[k] = R[k];
  } else {
    const double O[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
    const double c1 = (1.0 - cos(theta)) / theta_sq, c2 = (theta - sin(theta)) / (theta_sq * theta);
    for (int k = 0; k < 9; ++k) V[k] = ((k % 4 == 0) ? 1.0 : 0.0) + c1 * O[k] + c2 * O2[k];
  }
  for (int i = 0; i < 3; ++i) t[i] = V[3 * i] * a[3] + V[3 * i + 1] * a[4] + V[3 * i + 2] * a[5];
}

__device__ inline bool gicp_is_converged(const double* dR, const double* dt, double rot_eps, double trans_eps) {
  double rmax = 0, tmax = 0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rmax = fmax(rmax, fabs(dR[3 * i + j] - (i == j ? 1.0 : 0.0)));
    tmax = fmax(tmax, fabs(dt[i]));
  }
  return fmax(rmax / rot_eps, tmax / trans_eps) < 1;
}

__device__ inline void gicp_set_pose(GicpShared& s, const double* R, const double* t) {
  for (int k = 0; k < 9; ++k) s.R[k] = R[k];
  for (int k = 0; k < 3; ++k) s.t[k] = t[k];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) s.Tf[4 * i + j] = (float)R[3 * i + j];
    s.Tf[4 * i + 3] = (float)t[i];
  }
}

__host__ __device__ constexpr int gicp_hidx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }  // i <= j, 0..20

// LM / GN state machine (step_lm / step_gn / the outer loop of computeTransformation), one lane
static __device__ __noinline__ void gicp_step(GicpShared& s, const GicpParams& prm) {
  const double* T = s.tot;
  bool step_done = false, step_ok = true, try_lm = false;
  if (s.phase == GP_LINEARIZE) {
    s.n_lin++;
    s.hits += T[28];
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) s.H[6 * i + j] = s.H[6 * j + i] = T[gicp_hidx(i, j)];
    for (int i = 0; i < 6; ++i) s.b[i] = T[21 + i];
    s.y0 = T[27];
    s.last_err = s.y0;
    if (prm.lsq == B200REG_LSQ_GN) {
      double nb[6];
      for (int k = 0; k < 6; ++k) nb[k] = -s.b[k];
      solve6(s.H, nb, s.d);
      se3_exp(s.d, s.dR, s.dt);
      double nR[9], nt[3];
      iso_mul(s.dR, s.dt, s.x0R, s.x0t, nR, nt);
      for (int k = 0; k < 9; ++k) s.x0R[k] = nR[k];
      for (int k = 0; k < 3; ++k) s.x0t[k] = nt[k];
      step_done = true;
    } else {
      if (s.lambda < 0.0) {
        double mx = 0;
        for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(s.H[6 * i + i]));
        s.lambda = prm.lm_init_lambda_factor * mx;
      }
      s.nu = 2.0;
      s.lm_k = 0;
      try_lm = true;
    }
  } else if (s.phase == GP_ERROR) {
    s.n_err++;
    const double yi = T[27];
    double denom = 0;
    for (int k = 0; k < 6; ++k) denom += s.d[k] * (s.lambda * s.d[k] - s.b[k]);
    const double rho = (s.y0 - yi) / denom;
    if (rho < 0) {
      if (gicp_is_converged(s.dR, s.dt, prm.rot_eps, prm.trans_eps)) {
        step_done = true;
      } else {
        s.lambda = s.nu * s.lambda;
        s.nu = 2 * s.nu;
        s.lm_k++;
        if (s.lm_k < prm.lm_max_iterations) try_lm = true;
        else { step_done = true; step_ok = false; }  // "lm not converged!!"
      }
    } else {
      for (int k = 0; k < 9; ++k) s.x0R[k] = s.R[k];
      for (int k = 0; k < 3; ++k) s.x0t[k] = s.t[k];
      const double c = 2 * rho - 1;
      const double f = 1 - c * c * c;
      s.lambda = s.lambda * ((1.0 / 3.0 < f) ? f : 1.0 / 3.0);  // std::max(1/3, f): NaN keeps 1/3
      s.last_err = yi;
      step_done = true;
    }
  }
  if (try_lm) {
    double A[36], nb[6];
    for (int k = 0; k < 36; ++k) A[k] = s.H[k];
    for (int k = 0; k < 6; ++k) { A[7 * k] += s.lambda; nb[k] = -s.b[k]; }
    solve6(A, nb, s.d);
    se3_exp(s.d, s.dR, s.dt);
    double nR[9], nt[3];
    iso_mul(s.dR, s.dt, s.x0R, s.x0t, nR, nt);
    gicp_set_pose(s, nR, nt);
    s.phase = GP_ERROR;
    return;
  }
  if (step_done) {
    if (!step_ok) { s.phase = GP_DONE; return; }
    s.converged = gicp_is_converged(s.dR, s.dt, prm.rot_eps, prm.trans_eps) ? 1 : 0;
    s.iter++;
    if (s.converged || s.iter >= prm.max_iterations) { s.phase = GP_DONE; return; }
    s.nr_iterations = s.iter;
    gicp_set_pose(s, s.x0R, s.x0t);
    s.phase = GP_LINEARIZE;
  }
}

// the residual of one correspondence; accumulates into acc[0..27] when LIN (H, b, error) else acc[27]
template <bool LIN>
__device__ __forceinline__ void gicp_residual(const GicpShared& s, const float4 p, const float4 tq, const double M[6], double (&acc)[kGicpAcc]) {
  const double a0 = (double)p.x, a1 = (double)p.y, a2 = (double)p.z;
  const double x = s.R[0] * a0 + s.R[1] * a1 + s.R[2] * a2 + s.t[0];
  const double y = s.R[3] * a0 + s.R[4] * a1 + s.R[5] * a2 + s.t[1];
  const double z = s.R[6] * a0 + s.R[7] * a1 + s.R[8] * a2 + s.t[2];
  const double e0 = (double)tq.x - x, e1 = (double)tq.y - y, e2 = (double)tq.z - z;
  const double m00 = M[0], m01 = M[1], m02 = M[2], m11 = M[3], m12 = M[4], m22 = M[5];
  const double Me0 = m00 * e0 + m01 * e1 + m02 * e2, Me1 = m01 * e0 + m11 * e1 + m12 * e2, Me2 = m02 * e0 + m12 * e1 + m22 * e2;
  acc[27] += e0 * Me0 + e1 * Me1 + e2 * Me2;
  if (!LIN) return;
  acc[28] += 1.0;
  // J = [ skew(ta) | -I ],  skew(ta) = S = [[0,-z,y],[z,0,-x],[-y,x,0]]
  // A = M S (columns of S):  S col0 = (0, z, -y), col1 = (-z, 0, x), col2 = (y, -x, 0)
  const double A00 = m01 * z - m02 * y, A01 = -m00 * z + m02 * x, A02 = m00 * y - m01 * x;
  const double A10 = m11 * z - m12 * y, A11 = -m01 * z + m12 * x, A12 = m01 * y - m11 * x;
  const double A20 = m12 * z - m22 * y, A21 = -m02 * z + m22 * x, A22 = m02 * y - m12 * x;
  // H_rr = S^T A ; rows of S^T = columns of S
  acc[gicp_hidx(0, 0)] += z * A10 - y * A20;
  acc[gicp_hidx(0, 1)] += z * A11 - y * A21;
  acc[gicp_hidx(0, 2)] += z * A12 - y * A22;
  acc[gicp_hidx(1, 1)] += -z * A01 + x * A21;
  acc[gicp_hidx(1, 2)] += -z * A02 + x * A22;
  acc[gicp_hidx(2, 2)] += y * A02 - x * A12;
  // H_rt = S^T M (-I) = -(M S)^T = -A^T
  acc[gicp_hidx(0, 3)] -= A00; acc[gicp_hidx(0, 4)] -= A10; acc[gicp_hidx(0, 5)] -= A20;
  acc[gicp_hidx(1, 3)] -= A01; acc[gicp_hidx(1, 4)] -= A11; acc[gicp_hidx(1, 5)] -= A21;
  acc[gicp_hidx(2, 3)] -= A02; acc[gicp_hidx(2, 4)] -= A12; acc[gicp_hidx(2, 5)] -= A22;
  // H_tt = M
  acc[gicp_hidx(3, 3)] += m00; acc[gicp_hidx(3, 4)] += m01; acc[gicp_hidx(3, 5)] += m02;
  acc[gicp_hidx(4, 4)] += m11; acc[gicp_hidx(4, 5)] += m12; acc[gicp_hidx(5, 5)] += m22;
  // b = J^T M e:  b_r = S^T Me, b_t = -Me
  acc[21] += z * Me1 - y * Me2;
  acc[22] += -z * Me0 + x * Me2;
  acc[23] += y * Me0 - x * Me1;
  acc[24] -= Me0; acc[25] -= Me1; acc[26] -= Me2;
}

// mahalanobis_[i] = (C_B + R C_A R^T)^-1 as 6 doubles
__device__ __forceinline__ void gicp_mahalanobis(const double* R, const double* CA, const double* CB, double M[6]) {
  const double a[9] = {CA[0], CA[1], CA[2], CA[1], CA[3], CA[4], CA[2], CA[4], CA[5]};
  double RA[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) RA[3 * i + j] = R[3 * i] * a[j] + R[3 * i + 1] * a[3 + j] + R[3 * i + 2] * a[6 + j];
  double S[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S[3 * i + j] = RA[3 * i] * R[3 * j] + RA[3 * i + 1] * R[3 * j + 1] + RA[3 * i + 2] * R[3 * j + 2];
  S[0] += CB[0]; S[1] += CB[1]; S[2] += CB[2]; S[3] += CB[1]; S[4] += CB[3]; S[5] += CB[4]; S[6] += CB[2]; S[7] += CB[4]; S[8] += CB[5];
  double Si[9];
  inverse3(S, Si);
  M[0] = Si[0]; M[1] = Si[1]; M[2] = Si[2]; M[3] = Si[4]; M[4] = Si[5]; M[5] = Si[8];
}

__device__ __forceinline__ void gicp_group_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

template <bool PROF>
__global__ void __launch_bounds__(kGicpThreads, 1) k_gicp_align(const __grid_constant__ GicpJob job, int G, GicpParams prm, double* partials, unsigned int* barrier) {
  __shared__ GicpShared s;
  __shared__ float4 s_q[kGicpThreads];   // far-query queue of the current slice: transformed point, w = source index
  __shared__ float2 s_qb[kGicpThreads];  // best so far (d2, index bits)
  __shared__ int s_wq[kGicpWarps];
  const int rank = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_src = job.n_src;
  const GridParams gp = job.tgt.meta->grid;
  const bool grid_ok = job.tgt.n > 0 && gp.any && !gp.overflow;
  unsigned int epoch = 0;
  int parity = 0;
  if (tid == 0) {
    // x0 = Isometry3d(guess.cast<double>())
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) s.x0R[3 * i + j] = (double)job.guess[4 * j + i];
      s.x0t[i] = (double)job.guess[12 + i];
    }
    for (int k = 0; k < 9; ++k) s.dR[k] = (k % 4 == 0) ? 1.0 : 0.0;
    s.dt[0] = s.dt[1] = s.dt[2] = 0.0;
    s.lambda = -1.0; s.nu = 2.0; s.y0 = 0.0; s.last_err = 0.0;
    s.lm_k = 0; s.iter = 0; s.nr_iterations = 0; s.converged = 0; s.n_lin = 0; s.n_err = 0; s.hits = 0.0;
    gicp_set_pose(s, s.x0R, s.x0t);
    s.phase = (prm.max_iterations > 0 && n_src > 0) ? GP_LINEARIZE : GP_DONE;
  }
  __syncthreads();
  long long pf[PROF ? 16 : 1] = {0};
  while (s.phase != GP_DONE) {
    const int phase = s.phase;
    const bool seed_from_corr = s.n_lin > 0;  // job.corr holds the correspondences of this align's previous linearize pass
    const long long tp0 = PROF ? clock64() : 0;
    long long tp_near = tp0;
    double acc[kGicpAcc];
#pragma unroll
    for (int k = 0; k < kGicpAcc; ++k) acc[k] = 0.0;
    // ---- pass over the source points.  linearize: every thread resolves its own query in the near
    // phase (rings 0-1) and finishes the point; the few queries left open are queued in shared memory
    // and, after a block barrier, dealt round-robin to the CTA's 16 warps for the cooperative far
    // search — so a warp that happened to draw several far points does not hold up the pass.
    const int stride = G * kGicpWarps;  // 32-point groups per sweep of the whole grid
    auto finish_point = [&](int i, const float4 p, float best, int best_idx) {
      int c = -1;
      if (best_idx != kNoIndex && (double)best < prm.corr_dist2) c = best_idx;
      job.corr[i] = c;
      if (c < 0) return;
      double CA[6], CB[6], M[6];
      const double* pa = job.cov_src + (size_t)i * 6;
      const double* pb = job.cov_tgt + (size_t)c * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) { CA[k] = pa[k]; CB[k] = __ldg(pb + k); }
      gicp_mahalanobis(s.R, CA, CB, M);
      double* pm = job.mahal + (size_t)i * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) pm[k] = M[k];
      gicp_residual<true>(s, p, __ldg(job.tgt_pts + c), M, acc);
    };
    // 32-point groups are dealt round-robin over the warps of all CTAs (like the NDT pass): source
    // points without a correspondence come in spatial clumps, and their far searches would otherwise
    // all land in a few CTAs while the rest of the grid waits at the group barrier
    // linearize: prm.query_lanes lanes share a query in the near phase (nn_query_near_group), so a warp takes 32 /
    // query_lanes points; the error passes keep a point per lane.  Either way every point is visited once.
    const int ql = prm.query_lanes;
    const int per_warp = phase == GP_LINEARIZE ? 32 / ql : 32;
    const int n_groups = (n_src + per_warp - 1) / per_warp;
    const int sub = lane % ql;
    const bool leader = phase != GP_LINEARIZE || sub == 0;
    for (int q0 = 0; q0 < n_groups; q0 += stride) {
      const int qg = q0 + warp * G + rank;
      const int i = qg * per_warp + (phase == GP_LINEARIZE ? lane / ql : lane);
      const bool active = qg < n_groups && i < n_src;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) p = __ldg(job.src + i);
      if (phase == GP_LINEARIZE) {
        // update_correspondences: float transform, exact 1-NN, gate on the squared distance
        const float qx = affine_row(s.Tf[0], s.Tf[1], s.Tf[2], s.Tf[3], p.x, p.y, p.z);
        const float qy = affine_row(s.Tf[4], s.Tf[5], s.Tf[6], s.Tf[7], p.x, p.y, p.z);
        const float qz = affine_row(s.Tf[8], s.Tf[9], s.Tf[10], s.Tf[11], p.x, p.y, p.z);
        float best = 3.402823466e+38f;
        int best_idx = kNoIndex;
        int st = kNnDone;
        // from the second linearize pass on, last pass's correspondence seeds the search: it is a real
        // target point, so the answer is the same exact nearest neighbour, but nearly every neighbour cell
        // is pruned at once and a point that used to need a far search is settled after a ring or two
        if (active && grid_ok && seed_from_corr) {
          const int cp = job.corr[i];
          if (cp >= 0) {
            const float4 tp = __ldg(job.tgt_pts + cp);
            best = l2_simple(qx, qy, qz, tp.x, tp.y, tp.z);
            best_idx = cp;
          }
        }
        if (ql == 4) st = nn_query_near_group<4, true>(job.tgt, gp, nn_make_query(gp, qx, qy, qz), prm.search_d2, sub, active && grid_ok, best, best_idx);
        else if (ql == 2) st = nn_query_near_group<2, true>(job.tgt, gp, nn_make_query(gp, qx, qy, qz), prm.search_d2, sub, active && grid_ok, best, best_idx);
        else if (active && grid_ok) st = nn_query_near<4, true>(job.tgt, gp, nn_make_query(gp, qx, qy, qz), prm.search_d2, best, best_idx);
        const bool ok = st == kNnDone;
        // queue slots in thread order (ballot compaction): the far points are always dealt to the same
        // warps, so the summation order — and with it the result — is reproducible bit for bit
        const unsigned open = __ballot_sync(0xffffffffu, active && !ok && leader);
        if (lane == 0) s_wq[warp] = __popc(open);
        if (active && ok && leader) finish_point(i, p, best, best_idx);
        __syncthreads();
        if (PROF) tp_near = clock64();
        int qbase = 0, nq = 0;
#pragma unroll
        for (int w = 0; w < kGicpWarps; ++w) {
          const int c = s_wq[w];
          if (w < warp) qbase += c;
          nq += c;
        }
        if (PROF) pf[9] += nq;
        if (active && !ok && leader) {
          const int slot = qbase + __popc(open & ((1u << lane) - 1u));
          s_q[slot] = make_float4(qx, qy, qz, __int_as_float(i | (st == kNnBail ? (int)kBailFlag : 0)));
          s_qb[slot] = make_float2(best, __int_as_float(best_idx));
        }
        __syncthreads();
        const long long tq0 = PROF ? clock64() : 0;
        for (int e = warp; e < nq; e += kGicpWarps) {
          const float4 q = s_q[e];
          const float2 qb = s_qb[e];
          float fb = qb.x;
          int fi = __float_as_int(qb.y);
          const long long tf0 = PROF ? clock64() : 0;
          const bool done = nn_query_far_warp(job.tgt, gp, nn_make_query(gp, q.x, q.y, q.z), prm.search_d2, prm.far_ring, lane, fb, fi, (__float_as_int(q.w) & (int)kBailFlag) != 0);
          const long long tf1 = PROF ? clock64() : 0;
          if (!done) nn_query_brute_warp(job.tgt, q.x, q.y, q.z, lane, fb, fi);
          if (PROF) { pf[10] += done ? 0 : 1; pf[11] += tf1 - tf0; pf[12] += clock64() - tf1; }
          if (lane == 0) s_qb[e] = make_float2(fb, __int_as_float(fi));
        }
        __syncthreads();
        const long long tq1 = PROF ? clock64() : 0;
        // the far queries are FINISHED thread-per-query (Mahalanobis matrix, residual) once all searches are
        // done — not by lane 0 of each searching warp, which serialised ~400 FP64 operations per query
        if (tid < nq) {
          const float4 q = s_q[tid];
          const float2 qb = s_qb[tid];
          const int qi = __float_as_int(q.w) & (int)~kBailFlag;
          finish_point(qi, __ldg(job.src + qi), qb.x, __float_as_int(qb.y));
        }
        __syncthreads();  // the queue is reused by the next slice
        if (PROF) { pf[13] += tq0 - tp_near; pf[14] += tq1 - tq0; pf[15] += clock64() - tq1; }
      } else {
        const int c = active ? job.corr[i] : -1;
        if (c >= 0) {
          double M[6];
          const double* pm = job.mahal + (size_t)i * 6;
#pragma unroll
          for (int k = 0; k < 6; ++k) M[k] = pm[k];
          gicp_residual<false>(s, p, __ldg(job.tgt_pts + c), M, acc);
        }
      }
    }
    const long long tp1 = PROF ? clock64() : 0;
    if (PROF) {
      if (phase == GP_LINEARIZE) { pf[0] += tp_near - tp0; pf[1] += tp1 - tp_near; pf[7] += 1; } else { pf[2] += tp1 - tp0; pf[8] += 1; }
    }
    // ---- reduce: warp shuffles -> shared memory -> one partial row per CTA -> group
#pragma unroll
    for (int k = 0; k < kGicpAcc; ++k) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s.red[warp][k] = v;
    }
    __syncthreads();
    if (tid < kGicpAcc) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kGicpWarps; ++w) v += s.red[w][tid];
      if (G == 1) s.tot[tid] = v;
      else partials[((size_t)parity * G + rank) * kGicpStride + tid] = v;
    }
    const long long tp2 = PROF ? clock64() : 0;
    long long tp3 = tp2;
    if (G > 1) {
      epoch += (unsigned)G;
      gicp_group_barrier(barrier, epoch);
      if (PROF) tp3 = clock64();
      const double* base = partials + (size_t)parity * G * kGicpStride + lane;
      double v = 0.0;
      if (lane < kGicpAcc)
        for (int r = warp; r < G; r += kGicpWarps) v += __ldcg(base + (size_t)r * kGicpStride);
      __syncthreads();
      s.red[warp][lane] = v;
      __syncthreads();
      if (tid < kGicpAcc) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kGicpWarps; ++w) t += s.red[w][tid];
        s.tot[tid] = t;
      }
    }
    parity ^= 1;
    __syncthreads();
    const long long tp4 = PROF ? clock64() : 0;
    if (tid == 0) gicp_step(s, prm);
    __syncthreads();
    if (PROF) {
      const long long tp5 = clock64();
      pf[3] += tp2 - tp1; pf[4] += tp3 - tp2; pf[5] += tp4 - tp3; pf[6] += tp5 - tp4;
    }
  }
  if (PROF && job.prof && rank == 0 && tid == 0)
    for (int k = 0; k < 16; ++k) job.prof[k] = pf[k];
  if (rank == 0 && tid == 0) {
    b200reg_result r;
    for (int c = 0; c < 3; ++c) {
      for (int rr = 0; rr < 3; ++rr) r.transformation[4 * c + rr] = (float)s.x0R[3 * rr + c];
      r.transformation[4 * c + 3] = 0.f;
    }
    for (int rr = 0; rr < 3; ++rr) r.transformation[12 + rr] = (float)s.x0t[rr];
    r.transformation[15] = 1.f;
    r.fitness = 0.0;
    r.score = s.last_err;
    r.converged = s.converged;
    r.iterations = s.nr_iterations;
    r.evaluations = s.n_lin + s.n_err;
    r.passes = s.n_lin + s.n_err;
    r.hits = (long long)s.hits;
    *job.result = r;
    if (job.result_host) {
      *job.result_host = r;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int*>(job.done_flag) = job.done_seq;
    }
  }
  // the last CTA to leave zeroes the barrier counter and the exit counter for the next launch
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(barrier + 1, 1u);
    if (prev == gridDim.x - 1) {
      barrier[0] = 0u;
      barrier[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace b200

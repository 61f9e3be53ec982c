// b200reg — per-voxel statistics of the NDT target grid (pclomp::VoxelGridCovariance::applyFilter
// steps 2-3).  COMPILED WITH -fmad=false (see Makefile): one thread per occupied
// voxel walks its points in input order and every double multiply / add is rounded on its own,
// the evaluation order of the reference's serial CPU loop.  The eigenvalue test that accepts or
// rejects a voxel (lambda_0 < 0 on near-planar voxels is decided by the last bits) and the
// clamped covariance / inverse therefore come out bit-identical to a CPU evaluation in that
// order, instead of depending on FMA contraction.
#include "ndt_grid.cuh"
#include "small_solve.cuh"

namespace b200 {

// pass 1: one WARP per occupied voxel (grid-stride over warps).  The warp gathers up to 128 points of
// the run at once (one memory latency, four independent loads per lane) into its own shared-memory
// row and requests the NEXT 128 before it starts on these; lane a in 0..8 then adds statistic a
// (sum x y z, sum xx xy xz yy yz zz) point by point IN INPUT ORDER — nine independent serial double
// chains, each bit-identical to the reference's per-leaf accumulation — while lanes 9..11 do the
// float centroid sums.  Eight points per step are prepared (LDS, F2F, DMUL) ahead of the dependent
// DADD chain, which is what bounds a voxel: ~11 cycles per point (the densest 1 m voxel of an HDL-64
// scan holds ~500 points).  Rows are zero padded to a multiple of eight: x + (+0) = x exactly.
constexpr int kLeafWarps = 8;
constexpr int kLeafChunk = 128;
static __global__ void __launch_bounds__(kLeafWarps * 32) k_ndt_leaf_sums(NdtLeafArgs a, double* __restrict__ sums /*[n_vox][9]*/, float* __restrict__ csum /*[n_vox][3]*/) {
  __shared__ __align__(16) float s_pts[kLeafWarps][kLeafChunk * 4];
  const uint32_t* vals = sorted_in_b(a.meta) ? a.vals_b : a.vals_a;
  const int n_vox = (int)a.meta->n_vox;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* row = s_pts[warp];
  // operand selectors of this lane's statistic (lanes 9..11: the float centroid sums of x, y, z)
  const int ia = lane < 3 ? lane : lane < 6 ? 0 : lane < 8 ? 1 : lane == 8 ? 2 : lane < 12 ? lane - 9 : 0;
  const int ib = lane == 3 ? 0 : lane == 4 ? 1 : lane == 5 ? 2 : lane == 6 ? 1 : 2;
  const bool is_prod = lane >= 3 && lane < 9;
  const int n_warps = gridDim.x * kLeafWarps;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int slot = blockIdx.x * kLeafWarps + warp; slot < n_vox; slot += n_warps) {
    const uint32_t s = a.vox_start[slot], e = a.vox_start[slot + 1];
    double acc = 0.0;
    float accf = 0.f;
    float4 pre[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t i = s + 32u * q + lane;
      pre[q] = i < e ? __ldg(a.pts + vals[i]) : zero4;
    }
    for (uint32_t c = s; c < e; c += kLeafChunk) {
      const int cnt = (int)min((uint32_t)kLeafChunk, e - c);
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(row)[32 * q + lane] = pre[q];
      __syncwarp();
      if (c + kLeafChunk < e) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t i = c + kLeafChunk + 32u * q + lane;
          pre[q] = i < e ? __ldg(a.pts + vals[i]) : zero4;
        }
      }
      const int cnt8 = (cnt + 7) & ~7;
      for (int j = 0; j < cnt8; j += 8) {
        double term[8];
        float fas[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float fa = row[4 * (j + u) + ia], fb = row[4 * (j + u) + ib];
          fas[u] = fa;
          term[u] = is_prod ? (double)fa * (double)fb : (double)fa;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc += term[u];
          accf += fas[u];
        }
      }
    }
    if (lane < 9) sums[(size_t)slot * 9 + lane] = acc;
    else if (lane < 12) csum[(size_t)slot * 3 + (lane - 9)] = accf;
  }
}

// pass 2: one thread per occupied voxel: mean, covariance, eigenvalue clamp, inverse
// stage: per occupied slot, record + centroid (w: 0 = fewer than min_points, 1 = valid, 2 = rejected)
static __global__ void __launch_bounds__(64) k_ndt_leaf_stats(NdtLeafArgs a, const double* __restrict__ sums, const float* __restrict__ csum, NdtVoxel* __restrict__ stage_vox,
                                                             float4* __restrict__ stage_cen) {
  const int n_vox = (int)a.meta->n_vox;
  for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_vox; slot += gridDim.x * blockDim.x) {
    const uint32_t s = a.vox_start[slot], e = a.vox_start[slot + 1];
    const int n = (int)(e - s);
    const double* S = sums + (size_t)slot * 9;
    const double sx = S[0], sy = S[1], sz = S[2], xx = S[3], xy = S[4], xz = S[5], yy = S[6], yz = S[7], zz = S[8];
    const float cx = csum[(size_t)slot * 3], cy = csum[(size_t)slot * 3 + 1], cz = csum[(size_t)slot * 3 + 2];
    const double nd = (double)n;
    const double pt_sum[3] = {sx, sy, sz};
    const double mean[3] = {sx / nd, sy / nd, sz / nd};
    const float fn = (float)n;
    int nr_points = n;
    double cov[9] = {xx, xy, xz, xy, yy, yz, xz, yz, zz};
    double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool valid = false;
    if (n >= a.min_points) {
      // cov = (cov - 2 (pt_sum mean^T)) / n + mean mean^T ;  cov *= (n - 1) / n
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) cov[3 * r + c] = (cov[3 * r + c] - 2.0 * (pt_sum[r] * mean[c])) / nd + mean[r] * mean[c];
      const double ratio = (nd - 1.0) / nd;
#pragma unroll
      for (int k = 0; k < 9; ++k) cov[k] *= ratio;
      double ev[3], V[9];
      sym_eigen3(cov, ev, V);
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        nr_points = -1;
      } else {
        const double min_ev = a.eig_mult * ev[2];
        if (ev[0] < min_ev) {
          ev[0] = min_ev;
          if (ev[1] < min_ev) ev[1] = min_ev;
          // cov = evecs * diag(evals) * evecs^-1, products accumulated left to right
          double Vi[9];
          inverse3(V, Vi);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
              cov[3 * r + c] = ((V[3 * r + 0] * ev[0]) * Vi[0 + c] + (V[3 * r + 1] * ev[1]) * Vi[3 + c]) + (V[3 * r + 2] * ev[2]) * Vi[6 + c];
        }
        inverse3(cov, icov);
        double mx = -1.7976931348623157e308, mn = 1.7976931348623157e308;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mx = fmax(mx, icov[k]); mn = fmin(mn, icov[k]); }
        if (isinf(mx) || isinf(mn) || mx != mx || mn != mn) nr_points = -1;
        else valid = true;
      }
    }
    a.leaf_n[slot] = nr_points;
#pragma unroll
    for (int r = 0; r < 3; ++r) a.leaf_mean[(size_t)slot * 3 + r] = mean[r];
#pragma unroll
    for (int k = 0; k < 9; ++k) { a.leaf_cov[(size_t)slot * 9 + k] = cov[k]; a.leaf_icov[(size_t)slot * 9 + k] = icov[k]; }
    // Voxels with >= min_points get a record.  Those rejected by the eigenvalue test stay in the
    // KDTREE centroid search upstream (their centroid is in the kd-tree cloud, there is no
    // nr_points re-check and icov is left at zero): flag 2, zero icov, KDTREE mode only.
    const float flag = n >= a.min_points ? (valid ? 1.f : 2.f) : 0.f;
    stage_cen[slot] = make_float4(cx / fn, cy / fn, cz / fn, flag);
    if (n >= a.min_points) {
      NdtVoxel v;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        v.mean_hi[r] = (float)mean[r];
        v.mean_lo[r] = (float)(mean[r] - (double)v.mean_hi[r]);
      }
      v.icov[0] = valid ? (float)icov[0] : 0.f; v.icov[1] = valid ? (float)icov[1] : 0.f; v.icov[2] = valid ? (float)icov[2] : 0.f;
      v.icov[3] = valid ? (float)icov[4] : 0.f; v.icov[4] = valid ? (float)icov[5] : 0.f; v.icov[5] = valid ? (float)icov[8] : 0.f;
      stage_vox[slot] = v;
    }
  }
}

// One CTA: compact the records in ascending voxel order, size the hash on the device, fill it.
//   1  thread t owns the contiguous slots [t K, (t + 1) K): their keep flags are loaded together and
//      ONE block scan gives the record index of every kept slot, written to a slot -> record map
//   2  the table is cleared (its capacity follows from the record count)
//   3  slot-strided: map, centroid, key and the 48-byte record of a slot are loaded together
//      (coalesced, one latency), stored to the record arrays, and the key is inserted
// — three memory latencies in all instead of a serial load / scan / store round per 1024 slots.
static __global__ void __launch_bounds__(1024) k_ndt_table(NdtLeafArgs a, const NdtVoxel* __restrict__ stage_vox, const float4* __restrict__ stage_cen) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  constexpr int kMaxOwn = 8;  // slots per thread whose flags stay in registers; beyond that they are re-read
  constexpr uint32_t kNone = 0xFFFFFFFFu;
  const int n_vox = (int)a.meta->n_vox;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = (n_vox + 1023) / 1024;
  const int first = tid * K, last = min(first + K, n_vox);
  uint32_t* slot_rec = a.rec_key;  // [n_vox] scratch: record index | kNdtRejected, or kNone
  float flag[kMaxOwn];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kMaxOwn; ++k) {
    flag[k] = 0.f;
    if (k < K && first + k < last) flag[k] = stage_cen[first + k].w;
  }
#pragma unroll
  for (int k = 0; k < kMaxOwn; ++k) mine += flag[k] != 0.f;
  for (int slot = first + kMaxOwn; slot < last; ++slot) mine += stage_cen[slot].w != 0.f;
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = s_warp[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  uint32_t dst = s_warp[warp] + incl - mine;
  const uint32_t n_rec = s_total;
  uint32_t cap = 16;
  while (cap < 2 * n_rec) cap <<= 1;
  if (tid == 0) { a.gmeta->n_records = n_rec; a.gmeta->table_cap = cap; }
#pragma unroll
  for (int k = 0; k < kMaxOwn; ++k) {
    if (k < K && first + k < last) {
      slot_rec[first + k] = flag[k] != 0.f ? (dst | (flag[k] == 2.f ? kNdtRejected : 0u)) : kNone;
      dst += flag[k] != 0.f;
    }
  }
  for (int slot = first + kMaxOwn; slot < last; ++slot) {
    const float w = stage_cen[slot].w;
    slot_rec[slot] = w != 0.f ? (dst | (w == 2.f ? kNdtRejected : 0u)) : kNone;
    dst += w != 0.f;
  }
  for (uint32_t i = tid; i < cap; i += 1024) a.table[i] = make_uint2(kInvalidKey, 0u);
  __threadfence_block();
  __syncthreads();
  for (int slot = tid; slot < n_vox; slot += 1024) {
    const uint32_t rec = __ldcg(slot_rec + slot);
    const float4 c = stage_cen[slot];
    const uint32_t key = a.vox_key[slot];
    const float4* vsrc = reinterpret_cast<const float4*>(stage_vox + slot);
    const float4 v0 = vsrc[0], v1 = vsrc[1], v2 = vsrc[2];
    if (rec == kNone) continue;
    const uint32_t d = rec & ~kNdtRejected;
    float4* vdst = reinterpret_cast<float4*>(a.voxels + d);
    vdst[0] = v0; vdst[1] = v1; vdst[2] = v2;
    a.centroids[d] = c;
    // bucketised insert (see lk_bucket in ndt_align.cuh): entry 0 of the bucket, then entry 1, then the next bucket
    uint32_t b = ndt_hash(key, cap / 2 - 1);
    while (true) {
      uint32_t old = atomicCAS(&a.table[2 * b].x, kInvalidKey, key);
      if (old == kInvalidKey) { a.table[2 * b].y = rec; break; }
      old = atomicCAS(&a.table[2 * b + 1].x, kInvalidKey, key);
      if (old == kInvalidKey) { a.table[2 * b + 1].y = rec; break; }
      b = (b + 1) & (cap / 2 - 1);
    }
  }
}

cudaError_t ndt_leaf_prefer_shared() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute((const void*)k_ndt_leaf_sums, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute((const void*)k_ndt_leaf_stats, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)k_ndt_table, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_ndt_leaf_stats(cudaStream_t st, const NdtLeafArgs& a, NdtVoxel* stage_vox, float4* stage_cen, double* sums, float* csum) {
  launch_counter() += 3;
  k_ndt_leaf_sums<<<kNumSM * 4, kLeafWarps * 32, 0, st>>>(a, sums, csum);
  k_ndt_leaf_stats<<<kNumSM, 64, 0, st>>>(a, sums, csum, stage_vox, stage_cen);
  k_ndt_table<<<1, 1024, 0, st>>>(a, stage_vox, stage_cen);
  return cudaGetLastError();
}

}  // namespace b200

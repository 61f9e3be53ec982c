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

// pass 1: one CTA per occupied voxel (grid-stride).  All 256 threads gather 256 points of the run
// into shared memory at once (one memory latency, zero padded), then lane a in 0..8 of warp 0 adds
// statistic a (sum x y z, sum xx xy xz yy yz zz) point by point IN INPUT ORDER — nine independent
// serial double chains, each bit-identical to the reference's per-leaf accumulation (adding the
// zero padding changes nothing) — while lanes 9..11 do the float centroid sums.
static __global__ void __launch_bounds__(256) k_ndt_leaf_sums(NdtLeafArgs a, double* __restrict__ sums /*[n_vox][9]*/, float* __restrict__ csum /*[n_vox][3]*/) {
  __shared__ float4 s_pts[256];
  const uint32_t* vals = sorted_in_b(a.meta) ? a.vals_b : a.vals_a;
  const int n_vox = (int)a.meta->n_vox;
  const int tid = threadIdx.x, lane = tid & 31;
  // operand selectors of this lane's statistic
  const int ia = lane < 3 ? lane : lane == 3 || lane == 4 || lane == 5 ? 0 : lane == 6 || lane == 7 ? 1 : 2;
  const int ib = lane == 3 ? 0 : lane == 4 ? 1 : lane == 5 ? 2 : lane == 6 ? 1 : lane == 7 ? 2 : 2;
  const bool is_prod = lane >= 3 && lane < 9;
  for (int slot = blockIdx.x; slot < n_vox; slot += gridDim.x) {
    const uint32_t s = a.vox_start[slot], e = a.vox_start[slot + 1];
    double acc = 0.0;
    float accf = 0.f;
    for (uint32_t c = s; c < e; c += 256) {
      const int cnt = (int)min(256u, e - c);
      __syncthreads();
      s_pts[tid] = tid < cnt ? __ldg(a.pts + vals[c + tid]) : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
      if (tid < 32) {
        const int rounds = (cnt + 31) >> 5;
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 p = s_pts[r * 32 + j];
            const float fa = ia == 0 ? p.x : ia == 1 ? p.y : p.z;
            const float fb = ib == 0 ? p.x : ib == 1 ? p.y : p.z;
            const double term = is_prod ? (double)fa * (double)fb : (double)fa;
            acc += term;
            accf += (lane == 9 ? p.x : lane == 10 ? p.y : p.z);
          }
        }
      }
    }
    if (tid < 9) sums[(size_t)slot * 9 + tid] = acc;
    else if (tid < 12) csum[(size_t)slot * 3 + (tid - 9)] = accf;
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
static __global__ void __launch_bounds__(1024) k_ndt_table(NdtLeafArgs a, const NdtVoxel* __restrict__ stage_vox, const float4* __restrict__ stage_cen) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  const int n_vox = (int)a.meta->n_vox;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n_vox; base += 1024) {
    const int slot = base + threadIdx.x;
    float4 c = make_float4(0, 0, 0, 0);
    if (slot < n_vox) c = stage_cen[slot];
    const int keep = c.w != 0.f;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      a.voxels[dst] = stage_vox[slot];
      a.centroids[dst] = c;
      a.rec_key[dst] = a.vox_key[slot];
      a.rec_flag[dst] = c.w == 2.f ? kNdtRejected : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_base = off + __popc(bal);
    __syncthreads();
  }
  const uint32_t n_rec = s_base;
  uint32_t cap = 16;
  while (cap < 2 * n_rec) cap <<= 1;
  if (threadIdx.x == 0) { a.gmeta->n_records = n_rec; a.gmeta->table_cap = cap; }
  for (uint32_t i = threadIdx.x; i < cap; i += 1024) a.table[i] = make_uint2(kInvalidKey, 0u);
  __threadfence_block();
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_rec; i += 1024) {
    const uint32_t key = a.rec_key[i];
    uint32_t h = ndt_hash(key, cap - 1);
    while (true) {
      const uint32_t old = atomicCAS(&a.table[h].x, kInvalidKey, key);
      if (old == kInvalidKey) { a.table[h].y = i | a.rec_flag[i]; break; }
      h = (h + 1) & (cap - 1);
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
  k_ndt_leaf_sums<<<kNumSM * 8, 256, 0, st>>>(a, sums, csum);
  k_ndt_leaf_stats<<<kNumSM, 64, 0, st>>>(a, sums, csum, stage_vox, stage_cen);
  k_ndt_table<<<1, 1024, 0, st>>>(a, stage_vox, stage_cen);
  return cudaGetLastError();
}

}  // namespace b200

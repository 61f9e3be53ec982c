// b200reg — pcl::VoxelGrid<PointXYZ>::applyFilter on the device, reached through
// PrefilteringNodelet::downsample [REF apps/prefiltering_nodelet.cpp:249-260] and
// ScanMatchingOdometryNodelet::downsample [REF apps/scan_matching_odometry_nodelet.cpp:155-165].
#pragma once
#include "voxel_sort.cuh"

namespace b200 {

struct VgCounts {  // device-resident summary of the last filter call
  uint32_t n_out;
  uint32_t overflow;
};

// The points in sorted (voxel, input) order, one thread per sorted position.  The centroid pass used to gather
// pts[vals[j]] itself, one thread per voxel: two dependent random loads per point on a serial chain as long as the voxel is
// populated — the densest voxel of a scan (tens of points at 0.1 m next to the sensor) set the kernel's duration
// (63 us of a 190 us filter call on a 1 M-point scan, 26 of 81 us on an HDL-64 scan: the largest single kernel of the filter).
// Here every gather is its own thread; the centroid pass then reads consecutive 16-byte records.
constexpr int kVgChunk = kSegFirstTile;  // sorted positions per warp of k_vg_centroids (and records it stages in shared memory at a time)

__global__ void __launch_bounds__(256) k_vg_gather(const float4* __restrict__ pts, int n, const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
                                                   const SortMeta* __restrict__ meta, float4* __restrict__ sorted_pts) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (meta->grid.overflow || j >= (int)meta->n_valid || j >= n) return;
  const uint32_t* vals = sorted_in_b(meta) ? vals_b : vals_a;
  sorted_pts[j] = __ldg(pts + vals[j]);
}

// One thread per occupied voxel: float centroid accumulated in ascending input order (the stable
// sort's in-voxel order), divided by float(count) — A.1 step 7.  sorted_pts (optional): the points already gathered
// into sorted order by k_vg_gather.
__global__ void __launch_bounds__(256) k_vg_centroids(const float4* __restrict__ pts, int n, const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
                                                      const float4* __restrict__ sorted_pts, const uint32_t* __restrict__ seg_first,
                                                      const SortMeta* __restrict__ meta, const uint32_t* __restrict__ vox_start, const uint32_t* __restrict__ vox_key,
                                                      unsigned min_points, float4* __restrict__ out, uint32_t* __restrict__ out_id, uint32_t* __restrict__ out_count,
                                                      VgCounts* __restrict__ counts, VgCounts* host_counts, unsigned int* host_flag, unsigned int host_seq,
                                                      unsigned int* done_blocks, int publish_here, float4* host_out, unsigned host_cap, int gated) {
  // host_out (optional): the caller's page-locked output cloud, mapped into the device address space.
  // The centroids are stored there as well (coalesced 16-byte stores over PCIe), so the host needs
  // no D2H copy after the count arrives — the stores are fenced before the flag below.
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool overflow = meta->grid.overflow != 0;  // "Leaf size is too small for the input dataset": output = *input_
  const int n_vox = (int)meta->n_vox;
  if (overflow && gated) {
    // "leaf size too small" with the distance gate on: the output is the GATED input in order, which
    // k_gate_copy (one block, launched behind this kernel) compacts and publishes
  } else if (overflow) {
    if (i < n) {
      const float4 p = pts[i];
      out[i] = p;
      if (host_out && (unsigned)i < host_cap) host_out[i] = p;
    }
  } else {
    // One warp per kVgChunk sorted positions: it owns the runs that START there (seg_first = number of runs in front of
    // the tile, from the segmentation) and follows the last one past the tile's end.  A warp's work is bounded by the
    // tile plus one run, however the points crowd into the voxels next to the sensor — one thread per voxel walking its
    // run through global memory made the densest voxel the kernel's duration (63 us of a 190 us call on a 1 M-point scan).
    // The tile's records are loaded (eight coalesced 16-byte loads per lane) without waiting for the run table; every
    // lane then adds up its runs from shared memory — one serial float sum per voxel in input order, so the bits stay.
    const uint32_t* vals = sorted_in_b(meta) ? vals_b : vals_a;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = (uint32_t)i >> 5, p0 = tile * (uint32_t)kVgChunk, n_valid = meta->n_valid;
    if (p0 < n_valid) {  // warp-uniform
      __shared__ float s_x[8][kVgChunk], s_y[8][kVgChunk], s_z[8][kVgChunk];
      auto stage = [&](uint32_t c0) {  // records [c0, c0 + kVgChunk) -> shared memory
        float4 p[kVgChunk / 32];
#pragma unroll
        for (int u = 0; u < kVgChunk / 32; ++u) {
          const uint32_t j = c0 + u * 32 + lane;
          if (sorted_pts) p[u] = j < n_valid ? __ldg(sorted_pts + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          else p[u] = j < n_valid ? __ldg(pts + vals[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kVgChunk / 32; ++u) {
          s_x[warp][u * 32 + lane] = p[u].x;
          s_y[warp][u * 32 + lane] = p[u].y;
          s_z[warp][u * 32 + lane] = p[u].z;
        }
        __syncwarp();
      };
      const uint32_t v0 = seg_first[tile];
      const uint32_t v1 = p0 + (uint32_t)kVgChunk < (uint32_t)n ? seg_first[tile + 1] : (uint32_t)n_vox;
      uint32_t s = 0, e = 0;
      if (v0 + lane < v1) { s = vox_start[v0 + lane]; e = vox_start[v0 + lane + 1]; }
      stage(p0);
      const uint32_t tile_end = p0 + (uint32_t)kVgChunk;
      for (uint32_t vb = v0; vb < v1; vb += 32) {
        const uint32_t v = vb + lane;
        const bool mine = v < v1;
        uint32_t s_next = 0, e_next = 0;  // the next batch's run bounds travel while this batch is summed
        if (v + 32 < v1) { s_next = vox_start[v + 32]; e_next = vox_start[v + 33]; }
        float ax = 0.f, ay = 0.f, az = 0.f;
        const uint32_t hi = min(e, tile_end);
        for (uint32_t j = s; j < hi; ++j) {
          ax = __fadd_rn(ax, s_x[warp][j - p0]);
          ay = __fadd_rn(ay, s_y[warp][j - p0]);
          az = __fadd_rn(az, s_z[warp][j - p0]);
        }
        // the tile's last run may go on past its end (only the last lane of the last batch): the warp streams the rest
        const uint32_t tail = __reduce_max_sync(0xffffffffu, mine ? e : 0u);
        if (tail > tile_end) {  // warp-uniform
          __syncwarp();
          for (uint32_t c0 = tile_end; c0 < tail; c0 += kVgChunk) {
            stage(c0);
            const uint32_t lo = max(max(s, c0), tile_end), h2 = min(e, c0 + (uint32_t)kVgChunk);
            for (uint32_t j = lo; j < h2; ++j) {
              ax = __fadd_rn(ax, s_x[warp][j - c0]);
              ay = __fadd_rn(ay, s_y[warp][j - c0]);
              az = __fadd_rn(az, s_z[warp][j - c0]);
            }
            __syncwarp();
          }
        }
        if (mine) {
          const float cnt = (float)(e - s);
          const float4 c = make_float4(__fdiv_rn(ax, cnt), __fdiv_rn(ay, cnt), __fdiv_rn(az, cnt), 1.0f);
          out[v] = c;
          if (host_out && v < host_cap) host_out[v] = c;
          if (out_id) out_id[v] = vox_key[v];
          if (out_count) out_count[v] = e - s;
        }
        s = s_next;
        e = e_next;
      }
    }
  }
  // The point count goes to the device summary and, for the caller waiting on the host, straight into
  // mapped page-locked memory followed by a sequence flag.  The LAST block to finish publishes it:
  // when the host sees the flag every centroid is written and fenced, so a consumer on another
  // stream (the registration handle copying this cloud) may read the output right away.
  if (host_out) __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(done_blocks, 1u) != gridDim.x - 1) return;
  *done_blocks = 0u;
  if ((!publish_here && !overflow) || (overflow && gated)) return;  // a compaction pass follows and publishes instead
  const uint32_t n_out = overflow ? (uint32_t)n : (uint32_t)n_vox;
  counts->n_out = n_out;
  counts->overflow = overflow ? 1u : 0u;
  if (host_counts) {
    host_counts->n_out = n_out;
    host_counts->overflow = overflow ? 1u : 0u;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(host_flag) = host_seq;
  }
}

// min_points_per_voxel > 1 (never set by the reference): drop sparse voxels, keeping the order.
// Single block; voxel counts are at most the point count.
__global__ void __launch_bounds__(1024) k_vg_compact(const SortMeta* __restrict__ meta, unsigned min_points, float4* __restrict__ out, uint32_t* __restrict__ out_id,
                                                     uint32_t* __restrict__ out_count, VgCounts* __restrict__ counts, VgCounts* host_counts, unsigned int* host_flag, unsigned int host_seq) {
  if (meta->grid.overflow) return;  // published by k_vg_centroids
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  const int n_vox = (int)meta->n_vox;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n_vox; base += 1024) {
    const int i = base + threadIdx.x;
    float4 c = make_float4(0, 0, 0, 0);
    uint32_t id = 0, cnt = 0;
    int keep = 0;
    if (i < n_vox) { c = out[i]; id = out_id[i]; cnt = out_count[i]; keep = cnt >= min_points; }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
    __syncthreads();  // all reads of out[] in this chunk are done before any write (dst <= i)
    if (keep) { out[dst] = c; out_id[dst] = id; out_count[dst] = cnt; }
    if (threadIdx.x == 1023) s_base = off + __popc(bal);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts->n_out = s_base;
    counts->overflow = 0;
    if (host_counts) {
      host_counts->n_out = s_base;
      host_counts->overflow = 0;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int*>(host_flag) = host_seq;
    }
  }
}

// Overflow case with the distance gate: output = the points that pass the gate, in input order (what
// pcl::VoxelGrid does with the cloud distance_filter handed it).  One block; never on a hot path.
__global__ void __launch_bounds__(1024) k_gate_copy(const float4* __restrict__ pts, int n, PointGate gate, const SortMeta* __restrict__ meta, float4* __restrict__ out,
                                                    float4* host_out, unsigned host_cap, VgCounts* __restrict__ counts, VgCounts* host_counts, unsigned int* host_flag, unsigned int host_seq) {
  if (!meta->grid.overflow) return;  // the regular path published
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    float4 p = make_float4(0, 0, 0, 0);
    int keep = 0;
    if (i < n) { p = pts[i]; keep = point_takes_part(gate, 0, p.x, p.y, p.z); }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      out[dst] = p;
      if (host_out && dst < host_cap) host_out[dst] = p;
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_base = off + __popc(bal);
    __syncthreads();
  }
  if (host_out) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    counts->n_out = s_base;
    counts->overflow = 1;
    if (host_counts) {
      host_counts->n_out = s_base;
      host_counts->overflow = 1;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int*>(host_flag) = host_seq;
    }
  }
}

}  // namespace b200

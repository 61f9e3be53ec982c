// b200reg — device build of the synthetic lidar generator (bench / test infrastructure, not
// part of the registration path).  bench.py uses it to produce the 1000-scan odometry sequence
// and the loop-closure keyframes without a CPU ray caster in the loop.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../synth/synth_scene.h"

namespace {

__global__ void __launch_bounds__(256) k_synth_rays(int sensor, unsigned long long scene_seed, unsigned long long noise_seed, const double* __restrict__ pose, long rays,
                                                    float4* __restrict__ tmp, unsigned char* __restrict__ ok) {
  long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rays) return;
  synth::Sensor s = sensor == 0 ? synth::sensor_hdl64() : synth::sensor_dense128();
  double P[16];
  for (int i = 0; i < 16; ++i) P[i] = pose[i];
  float o[4];
  bool hit = synth::scan_ray(s, scene_seed, noise_seed, P, r, o);
  ok[r] = hit ? 1 : 0;
  if (hit) tmp[r] = make_float4(o[0], o[1], o[2], o[3]);
}

// order-preserving compaction (firing order), single pass over per-block counts
__global__ void __launch_bounds__(256) k_synth_count(const unsigned char* __restrict__ ok, long rays, unsigned int* __restrict__ block_count) {
  long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int v = r < rays ? ok[r] : 0;
  int c = __syncthreads_count(v);
  if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned)c;
}
__global__ void __launch_bounds__(1024) k_synth_scan(unsigned int* __restrict__ block_count, int n_blocks, unsigned int* __restrict__ total) {
  __shared__ unsigned int s_warp[32];
  __shared__ unsigned int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    int i = base + threadIdx.x;
    unsigned int v = i < n_blocks ? block_count[i] : 0;
    unsigned int incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (i < n_blocks) block_count[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_base = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_base;
}
__global__ void __launch_bounds__(256) k_synth_compact(const unsigned char* __restrict__ ok, const float4* __restrict__ tmp, long rays, const unsigned int* __restrict__ block_off,
                                                       float4* __restrict__ out) {
  long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int v = r < rays ? ok[r] : 0;
  __shared__ unsigned int s_warp[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int bal = __ballot_sync(0xffffffffu, v);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  unsigned int off = block_off[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  if (v) out[off + __popc(bal & ((1u << lane) - 1u))] = tmp[r];
}

}  // namespace

extern "C" {

long long b200synth_num_rays(int sensor) {
  synth::Sensor s = sensor == 0 ? synth::sensor_hdl64() : synth::sensor_dense128();
  return (long long)s.beams * s.azimuth_steps;
}

// Generates one scan into d_out (device, capacity = num_rays float4).  pose_rowmajor: host, 16
// doubles sensor->world.  Returns the point count (>= 0) or -1 on a CUDA error.  Synchronous.
long long b200synth_scan_device(int device, int sensor, unsigned long long scene_seed, unsigned long long noise_seed, const double* pose_rowmajor, float* d_out) {
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  const long rays = (long)b200synth_num_rays(sensor);
  const int blocks = (int)((rays + 255) / 256);
  double* d_pose = nullptr;
  float4* d_tmp = nullptr;
  unsigned char* d_ok = nullptr;
  unsigned int* d_cnt = nullptr;
  long long result = -1;
  unsigned int total = 0;
  if (cudaMalloc(&d_pose, 16 * sizeof(double)) != cudaSuccess) goto done;
  if (cudaMalloc(&d_tmp, rays * sizeof(float4)) != cudaSuccess) goto done;
  if (cudaMalloc(&d_ok, rays) != cudaSuccess) goto done;
  if (cudaMalloc(&d_cnt, (blocks + 1) * sizeof(unsigned int)) != cudaSuccess) goto done;
  if (cudaMemcpy(d_pose, pose_rowmajor, 16 * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) goto done;
  k_synth_rays<<<blocks, 256>>>(sensor, scene_seed, noise_seed, d_pose, rays, d_tmp, d_ok);
  k_synth_count<<<blocks, 256>>>(d_ok, rays, d_cnt);
  k_synth_scan<<<1, 1024>>>(d_cnt, blocks, d_cnt + blocks);
  k_synth_compact<<<blocks, 256>>>(d_ok, d_tmp, rays, d_cnt, (float4*)d_out);
  if (cudaMemcpy(&total, d_cnt + blocks, sizeof(unsigned int), cudaMemcpyDeviceToHost) != cudaSuccess) goto done;
  if (cudaDeviceSynchronize() != cudaSuccess) goto done;
  result = (long long)total;
done:
  cudaFree(d_pose); cudaFree(d_tmp); cudaFree(d_ok); cudaFree(d_cnt);
  return result;
}

void b200synth_traj(long long k, unsigned long long seed, double* T_rowmajor) { synth::traj_kitti_like((long)k, seed, T_rowmajor); }
void b200synth_pose(const double* xyzrpy, double* T_rowmajor) { synth::pose_from_xyzrpy(xyzrpy, T_rowmajor); }

}  // extern "C"

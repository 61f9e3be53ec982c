// b200reg — shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

namespace b200 {

constexpr int kNumSM = 148;  // B200; grids are sized in multiples of the SM count

// kernels launched by this library since load (bench.py reports it as gpu_launches)
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}

#define B200_CUDA_TRY(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
      return B200REG_E_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

// growable device buffer; never shrinks (the odometry loop reuses the same sizes every frame)
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// float <-> order-preserving int (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

// pcl::transformPoint with a Matrix4f, evaluated left to right without contraction:
// ((m0*x + m1*y) + m2*z) + m3   — the order the oracle uses (oracle_linalg.hpp m4f_apply)
__device__ __forceinline__ float affine_row(float m0, float m1, float m2, float m3, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, x), __fmul_rn(m1, y)), __fmul_rn(m2, z)), m3);
}

// FLANN L2_Simple: ((dx*dx) + dy*dy) + dz*dz, no contraction
__device__ __forceinline__ float l2_simple(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Voxel lattice of a cloud: pcl::VoxelGrid / VoxelGridCovariance bookkeeping
struct GridParams {
  int min_b[3], max_b[3], div_b[3], mul[3];
  int overflow;  // dx*dy*dz > INT32_MAX ("leaf size too small")
  int any;       // at least one finite point
  float inv_leaf[3], leaf[3];
};

}  // namespace b200

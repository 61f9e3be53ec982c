// Developer microbenchmark: FP64 / conversion issue rate and dependent-chain latency on the box's GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k_rate(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = (float)a0, f1 = (float)a1;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c); }
    if (OP == 1) { a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c; }
    if (OP == 2) { a0 += (double)f0; a1 += (double)f1; f0 += 1.f; f1 += 1.f; a2 += (double)f0; a3 += (double)f1; a4 += (double)(f0 * 0.5f); a5 += (double)(f1 * 0.5f); a6 += (double)(f0 * 0.25f); a7 += (double)(f1 * 0.25f); }
    if (OP == 3) { a0 = fma(a0, m, c); }  // dependent chain
    if (OP == 4) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1;
}
template <int OP>
void run(const char* name, int ops_per_iter, int blocks, int threads) {
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int iters = 4096;
  k_rate<OP><<<blocks, threads>>>(out, iters, 1.0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k_rate<OP><<<blocks, threads>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)blocks * threads * iters * ops_per_iter;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s blocks %4d x %4d: %8.3f ms  %8.2f Gop/s  %.2f thread-ops/clk/SM  (cycles/iter/warp %.1f)\n", name, blocks, threads, ms, ops / ms * 1e-6, ops / cycles / 148.0, cycles / iters);
  cudaFree(out);
}
int main() {
  run<0>("DFMA x8 independent", 8, 148 * 4, 256);
  run<1>("DADD x8 independent", 8, 148 * 4, 256);
  run<2>("F2F.F64.F32 + DADD x8", 8, 148 * 4, 256);
  run<3>("DFMA dependent chain", 1, 148, 32);
  run<4>("FFMA x2 dependent chains", 2, 148, 32);
  run<0>("DFMA x8, one warp per SM", 8, 148, 32);
  return 0;
}

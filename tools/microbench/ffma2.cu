// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 / FMUL2 / FADD2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int MODE>
__global__ void k(float* out, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  float2 m = make_float2(a, a * 1.0001f), c = make_float2(b, b * 0.999f);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { x[i].x = __fmaf_rn(x[i].x, m.x, c.x); x[i].y = __fmaf_rn(x[i].y, m.y, c.y); }          // 2 scalar FFMA
      if (MODE == 1) { x[i] = __ffma2_rn(x[i], m, c); }                                                        // 1 FFMA2
      if (MODE == 2) { x[i] = __fmul2_rn(x[i], m); }
      if (MODE == 3) { x[i] = __fadd2_rn(x[i], c); }
      if (MODE == 4) { x[i].x = __fmul_rn(x[i].x, m.x); x[i].y = __fmul_rn(x[i].y, m.y); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = 148 * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(d, 1.0000001f, 1e-7f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, 1.0000001f, 1e-7f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double elem_ops = (double)blocks * threads * ITERS * 8 * 2;  // scalar-equivalent element operations
  printf("%-28s %8.3f ms  %8.2f T elem-op/s\n", name, ms, elem_ops / (ms * 1e-3) / 1e12);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("2x scalar FFMA", d); run<1>("1x FFMA2", d); run<4>("2x scalar FMUL", d); run<2>("1x FMUL2", d); run<3>("1x FADD2", d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

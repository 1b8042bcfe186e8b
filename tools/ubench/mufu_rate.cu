// Micro-benchmark: MUFU throughput per SM for tanh.approx.f32, tanh.approx.f16x2 (two MUFU.TANH.F16 in SASS),
// ex2.approx.f32 and ex2.approx.f16x2 - does a packed-half transcendental cost less XU time per ELEMENT on sm_100a?
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a mufu_rate.cu -o mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters) {
  float x[8];
  unsigned h[4];
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + 1) + 0.1f * i;
  for (int i = 0; i < 4; ++i) { __half2 v = __floats2half2_rn(x[2 * i], x[2 * i + 1]); h[i] = *reinterpret_cast<unsigned*>(&v); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i];
  for (int i = 0; i < 4; ++i) s += __half2float(reinterpret_cast<__half2*>(&h[i])->x);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4096;
  const char* names[5] = {"tanh.approx.f32", "tanh.approx.f16x2", "ex2.approx.ftz.f32", "ex2.approx.f16x2", "rcp.approx.ftz.f32"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 512>>>(out, cyc, iters);
      if (mode == 1) k<1><<<148, 512>>>(out, cyc, iters);
      if (mode == 2) k<2><<<148, 512>>>(out, cyc, iters);
      if (mode == 3) k<3><<<148, 512>>>(out, cyc, iters);
      if (mode == 4) k<4><<<148, 512>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 16 warps per SM, 8 elements per thread per iteration
    const double elems = 512.0 * 8 * iters;
    printf("%-20s %lld cycles for %d iterations: %.2f elements per clock per SM\n", names[mode], h, iters, elems / (double)h);
  }
  return 0;
}

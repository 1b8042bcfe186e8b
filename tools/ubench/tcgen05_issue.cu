// Micro-benchmark: issue cost and round-trip latency of small tcgen05.mma / commit / fence from one warp.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../speech_separation_b200/csrc tcgen05_issue.cu -o tcgen05_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vatss::ptx;

template <bool WARP>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (WARP) umma_f16_warp<1>(d, a, b, idesc, acc);
  else umma_f16<1>(d, a, b, idesc, acc);
}

template <bool WARP>
__global__ void __launch_bounds__(256) k_bench(long long* out, int busy) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar = base + 65536, slot = base + 65536 + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<1>(slot, 512); tmem_relinquish<1>(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 64);
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 0) {
    if (WARP || lane == 0) {
      const uint32_t tm = WARP ? __shfl_sync(0xffffffffu, tmem, 0) : tmem;
      const uint64_t a = smem_desc_sw128_kmajor(base), b = smem_desc_sw128_kmajor(base + 32768);
      int o = 0;
      uint32_t phase = 0;
      const int Ns[4] = {32, 64, 128, 192};
      for (int rep = 0; rep < 2; ++rep) {
        o = 0;
        for (int ni = 0; ni < 4; ++ni) {
          const uint32_t idesc = idesc_f16(128, Ns[ni], 0);
          // (1) 16 back-to-back MMAs: issue cost
          long long t0 = clock64();
#pragma unroll
          for (int i = 0; i < 16; ++i) mma<WARP>(tm, a + 2 * (i & 3), b + 2 * (i & 3), idesc, i > 0);
          long long t1 = clock64();
          if (WARP) umma_commit_warp(bar); else umma_commit(bar);
          long long t2 = clock64();
          mbar_wait(bar, phase); phase ^= 1;
          long long t3 = clock64();
          if (lane == 0) { out[o] = t1 - t0; out[o + 1] = t2 - t1; out[o + 2] = t3 - t2; }
          o += 3;
          // (2) round trip of a single MMA + commit + wait
          t0 = clock64();
          mma<WARP>(tm, a, b, idesc, 0);
          if (WARP) umma_commit_warp(bar); else umma_commit(bar);
          mbar_wait(bar, phase); phase ^= 1;
          t1 = clock64();
          if (lane == 0) out[o] = t1 - t0;
          o += 1;
        }
        // (3) 16 fences
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 16; ++i) tc_fence_after();
        long long t1 = clock64();
        if (lane == 0) out[o] = t1 - t0;
        o += 1;
        // (4) 4 MMAs + commit, repeated 8 times without waiting (8 arrivals on distinct phases are not waited: use count trick)
        t0 = clock64();
        for (int r = 0; r < 8; ++r) {
          const uint32_t idesc = idesc_f16(128, 32, 0);
#pragma unroll
          for (int i = 0; i < 4; ++i) mma<WARP>(tm, a + 2 * i, b + 2 * i, idesc, i > 0);
          if (WARP) umma_commit_warp(bar); else umma_commit(bar);
          mbar_wait(bar, phase); phase ^= 1;
        }
        t1 = clock64();
        if (lane == 0) out[o] = t1 - t0;
        o += 1;
      }
    }
    __syncwarp();
    if (lane == 0) stop = 1;
  } else if (busy) {
    // other warps spin (polling pressure on the schedulers), like the softmax warps waiting on mbarriers
    float x = threadIdx.x;
    while (!stop) {
      if (busy == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x = __expf(x) * 0.5f;
      }
    }
    if (x == 123.456f) out[63] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  const int smem = 65536 + 256;
  cudaFuncSetAttribute(k_bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int warpmode = 0; warpmode < 2; ++warpmode)
    for (int busy = 0; busy < 3; ++busy) {
      cudaMemset(d, 0, 64 * sizeof(long long));
      if (warpmode) k_bench<true><<<1, 256, smem>>>(d, busy);
      else k_bench<false><<<1, 256, smem>>>(d, busy);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[64];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("issuer=%s other warps=%s\n", warpmode ? "whole warp + elect" : "lane 0 only", busy == 0 ? "idle" : busy == 1 ? "spinning" : "MUFU loop");
      const int Ns[4] = {32, 64, 128, 192};
      for (int ni = 0; ni < 4; ++ni)
        printf("  N=%3d: 16 MMAs issue %lld clk (%.1f each), commit %lld, drain wait %lld | single MMA+commit round trip %lld\n", Ns[ni],
               h[ni * 4], h[ni * 4] / 16.0, h[ni * 4 + 1], h[ni * 4 + 2], h[ni * 4 + 3]);
      printf("  16 x tcgen05.fence::after_thread_sync: %lld clk; 8 x (4 MMA N=32 + commit + wait): %lld clk (%.0f per round)\n", h[16], h[17], h[17] / 8.0);
    }
  return 0;
}

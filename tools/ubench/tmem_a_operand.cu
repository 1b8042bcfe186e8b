// Probes for the round-2 attention kernel (P kept in TMEM, FA4 style):
//  (1) layout of a kind::f16 A operand read from TMEM (tcgen05.mma [d], [a_tmem], b_desc): which half of a 32-bit
//      TMEM cell is the even k, and how far a K = 16 step advances (8 columns).  A[i][k] small integers (exact), B
//      K-major SWIZZLE_128B in shared memory, D compared with the CPU product for both packing hypotheses.
//  (2) ceiling of the softmax inner loop on one SM: 8 warps (two warpgroups), each thread owns one row of an S tile
//      of NB fp32 columns in TMEM: max pass, then exp2(s - max) -> fp16 pairs written back in place (tcgen05.st),
//      fp32 row sum.  No MMAs, no barriers: cycles per (128 x NB) job per warpgroup = the MUFU / issue bound.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../speech_separation_b200/csrc tmem_a_operand.cu -o tmem_a_operand
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vatss::ptx;

constexpr int KDIM = 32, NCOL = 32;

__global__ void __launch_bounds__(128) k_probe_ts(const __half* A, const __half* B, float* out, int swap_halves) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar = base + 16384, slot = base + 16384 + 16;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // B[n][k] K-major, rows of 128 B
  for (int i = threadIdx.x; i < NCOL * KDIM; i += blockDim.x) {
    const int n = i / KDIM, k = i % KDIM;
    *reinterpret_cast<__half*>(smem + sw128_offset((uint32_t)n, (uint32_t)(k >> 3)) + (k & 7) * 2) = B[n * KDIM + k];
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc<1>(slot, 64); tmem_relinquish<1>(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 16384 + 16);
  const uint32_t lane_base = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
  {   // A row of this thread -> TMEM columns [32, 48): two halfs per cell
    uint32_t pk[16];
    for (int c = 0; c < 16; ++c) {
      const uint16_t lo = __half_as_ushort(A[threadIdx.x * KDIM + 2 * c]);
      const uint16_t hi = __half_as_ushort(A[threadIdx.x * KDIM + 2 * c + 1]);
      pk[c] = swap_halves ? ((uint32_t)lo << 16 | hi) : ((uint32_t)hi << 16 | lo);
    }
    tmem_st_32x32b_x16(lane_base + 32, pk);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_f16(128, NCOL, 0);
    for (int k16 = 0; k16 < KDIM / 16; ++k16)
      umma_f16_ts(tmem, tmem + 32 + 8 * k16, smem_desc_sw128_kmajor(base) + 2 * k16, idesc, k16 > 0 ? 1u : 0u);
    umma_commit(bar);
    mbar_wait(bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(lane_base, v);
  tmem_ld_wait();
  for (int i = 0; i < NCOL; ++i) out[threadIdx.x * NCOL + i] = __uint_as_float(v[i]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tmem, 64);
}

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y;
}

// VARIANT 0: two TMEM passes (max, then exp); 1: exp pass only (max known); 2: max pass only
template <int NB, int VARIANT>
__global__ void __launch_bounds__(256, 1) k_softmax_ceiling(float* sums, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&slot), 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int wg = warp >> 2, q = warp & 3;
  const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + wg * 192;
  {   // some finite data
    uint32_t z[16];
    for (int c0 = 0; c0 < NB; c0 += 16) {
      for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-0.01f * (float)((threadIdx.x * 7 + c0 + i) % 97));
      tmem_st_32x32b_x16(t_base + c0, z);
    }
    tmem_st_wait();
  }
  __syncthreads();
  float total = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float mx = -1e30f;
    if (VARIANT != 1) {
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_base + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      }
    } else {
      mx = (float)it * 1e-3f;
    }
    float sum = 0.f;
    if (VARIANT != 2) {
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(t_base, v[0]);
#pragma unroll
      for (int c = 0; c < NB / 32; ++c) {
        tmem_ld_wait();
        if (c + 1 < NB / 32) tmem_ld_32x32b_x32(t_base + (c + 1) * 32, v[(c + 1) & 1]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2f(__uint_as_float(v[c & 1][2 * i]) - mx);
          const float e1 = ex2f(__uint_as_float(v[c & 1][2 * i + 1]) - mx);
          sum += e0 + e1;
          const __half2 h2 = __floats2half2_rn(e0, e1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        // (benchmark only: P goes to a scratch region so that the S data stays valid for the next iteration)
        tmem_st_32x32b_x16(t_base + 160 + (c & 1) * 16, pk);
      }
      tmem_st_wait();
    }
    total += sum + mx;
  }
  const long long t1 = clock64();
  sums[blockIdx.x * 256 + threadIdx.x] = total;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

int main() {
  // ---- (1) A operand from TMEM
  __half hA[128 * KDIM], hB[NCOL * KDIM];
  srand(1);
  for (auto& x : hA) x = __float2half((float)(rand() % 9 - 4));
  for (auto& x : hB) x = __float2half((float)(rand() % 7 - 3));
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, 128 * NCOL * 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const int smem = 16384 + 64;
  cudaFuncSetAttribute(k_probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int swap = 0; swap < 2; ++swap) {
    k_probe_ts<<<1, 128, smem>>>(dA, dB, dO, swap);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("probe_ts failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float hO[128 * NCOL];
    cudaMemcpy(hO, dO, sizeof(hO), cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < NCOL; ++n) {
        double ref = 0;
        for (int k = 0; k < KDIM; ++k) ref += (double)__half2float(hA[i * KDIM + k]) * __half2float(hB[n * KDIM + k]);
        const double e = fabs(ref - hO[i * NCOL + n]);
        if (e > maxerr) maxerr = e;
      }
    printf("A from TMEM, %s: max |D - A B^T| = %g  %s\n", swap ? "even k in HIGH half" : "even k in LOW half", maxerr,
           maxerr == 0 ? "<- MATCH (K step = 8 columns)" : "");
  }
  // ---- (2) softmax loop ceiling
  float* dS; long long* dC;
  cudaMalloc(&dS, 148 * 256 * 4); cudaMalloc(&dC, 148 * 8);
  const int iters = 200;
  auto run = [&](auto kern, const char* name, int nb) {
    kern<<<148, 256>>>(dS, dC, iters);
    kern<<<148, 256>>>(dS, dC, iters);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); exit(1); }
    long long hC[148];
    cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto c : hC) avg += (double)c; avg /= 148;
    printf("%s NB=%d: %.0f cycles per iteration (two warpgroups, one 128 x NB job each) = %.2f cycles per exp-column per SM "
           "(MUFU floor at 16/clk: %.2f)\n", name, nb, avg / iters, avg / iters / (2.0 * nb), 128.0 / 16.0);
  };
  run(k_softmax_ceiling<160, 0>, "max pass + exp pass", 160);
  run(k_softmax_ceiling<160, 1>, "exp pass only      ", 160);
  run(k_softmax_ceiling<160, 2>, "max pass only      ", 160);
  run(k_softmax_ceiling<128, 0>, "max pass + exp pass", 128);
  return 0;
}

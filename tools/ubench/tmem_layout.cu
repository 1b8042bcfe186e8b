// Probe: where do the rows of a tcgen05.mma accumulator land in TMEM for M = 64 (cta_group::1) and
// M = 128 / 256 (cta_group::2)?  D[i][j] = (i + 1) + 256 (j + 1) is exact in fp32, so every (lane, column) of the
// 128 x N TMEM window can be decoded back to the (row, column) of D it holds.  Needed for interleaving two 64-row
// recurrences per CTA pair in the LSTM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../speech_separation_b200/csrc tmem_layout.cu -o tmem_layout
#include <cstdio>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vatss::ptx;

constexpr int N = 64;   // accumulator columns

// A: rows x 16 halfs used (K = 16), K-major SWIZZLE_128B rows of 128 B.  A[i][0] = row_id + 1, A[i][1] = 1.
// B: N_local rows: B[j][0] = 1, B[j][1] = 256 (col_id + 1).
__device__ void fill_tiles(unsigned char* smA, unsigned char* smB, int rowsA, int row0, int rowsB, int col0) {
  for (int i = threadIdx.x; i < rowsA * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    const float v = k == 0 ? (float)(row0 + r + 1) : (k == 1 ? 1.f : 0.f);
    *reinterpret_cast<__half*>(smA + sw128_offset((uint32_t)r, (uint32_t)(k >> 3)) + (k & 7) * 2) = __float2half(v);
  }
  for (int i = threadIdx.x; i < rowsB * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    const float v = k == 0 ? 1.f : (k == 1 ? 256.f * (float)(col0 + r + 1) : 0.f);
    *reinterpret_cast<__half*>(smB + sw128_offset((uint32_t)r, (uint32_t)(k >> 3)) + (k & 7) * 2) = __float2half(v);
  }
}

__device__ void dump_tmem(uint32_t tmem, float* out) {   // out[128][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
}

__global__ void __launch_bounds__(128) k_probe_cg1(float* out, int M) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  unsigned char* smA = smem;
  unsigned char* smB = smem + 16384;
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  fill_tiles(smA, smB, M, 0, N, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc<1>(slot, 64); tmem_relinquish<1>(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 16);
  // poison the window so untouched lanes are visible
  if (threadIdx.x < 128) {
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0xFFC00000u;   // NaN
    for (int c0 = 0; c0 < N; c0 += 16) tmem_st_32x32b_x16(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + c0, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    umma_f16<1>(tmem, smem_desc_sw128_kmajor(base), smem_desc_sw128_kmajor(base + 16384), idesc_f16(M, N, 0), 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  dump_tmem(tmem, out);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tmem, 64);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_probe_cg2(float* out, int M) {   // M = 128 or 256
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  unsigned char* smA = smem;
  unsigned char* smB = smem + 16384;
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  fill_tiles(smA, smB, M / 2, (int)rank * (M / 2), N / 2, (int)rank * (N / 2));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc<2>(slot, 64); tmem_relinquish<2>(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 32768 + 16);
  if (threadIdx.x < 128) {
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0xFFC00000u;
    for (int c0 = 0; c0 < N; c0 += 16) tmem_st_32x32b_x16(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + c0, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    umma_f16<2>(tmem, smem_desc_sw128_kmajor(base), smem_desc_sw128_kmajor(base + 16384), idesc_f16(M, N, 0), 0);
    umma_commit_cg2(bar, 3);
  }
  if (threadIdx.x == 0) mbar_wait(bar, 0);
  __syncthreads();
  tc_fence_after();
  dump_tmem(tmem, out + (size_t)rank * 128 * N);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) tmem_dealloc<2>(tmem, 64);
}

static void report(const char* name, const float* h, int ctas) {
  printf("%s\n", name);
  for (int c = 0; c < ctas; ++c) {
    printf("  CTA %d: ", c);
    // decode (row, first column) of every lane, then print maximal runs of consecutive rows / untouched lanes
    int row[128], col[128];
    for (int lane = 0; lane < 128; ++lane) {
      const float v = h[((size_t)c * 128 + lane) * N];
      if (v == v) { const int iv = (int)v - 1; row[lane] = iv % 256; col[lane] = iv / 256 - 1; }
      else { row[lane] = -1; col[lane] = -1; }
    }
    int s0 = 0;
    for (int lane = 1; lane <= 128; ++lane) {
      const bool cont = lane < 128 && ((row[lane] < 0 && row[lane - 1] < 0) ||
                                       (row[lane] >= 0 && row[lane] == row[lane - 1] + 1 && col[lane] == col[lane - 1]));
      if (cont) continue;
      if (row[s0] < 0) printf("lanes %d-%d untouched; ", s0, lane - 1);
      else printf("lanes %d-%d = rows %d-%d (TMEM column 0 = D column %d); ", s0, lane - 1, row[s0], row[lane - 1], col[s0]);
      s0 = lane;
    }
    // how many TMEM columns were written in the first lane
    int used = 0;
    for (int j = 0; j < N; ++j) { const float v = h[((size_t)c * 128) * N + j]; if (v == v) ++used; }
    printf("| %d of %d TMEM columns written\n", used, N);
  }
}

int main() {
  float* d;
  cudaMalloc(&d, 2 * 128 * N * sizeof(float));
  float* h = (float*)malloc(2 * 128 * N * sizeof(float));
  const int smem = 32768 + 64;
  cudaFuncSetAttribute(k_probe_cg1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_probe_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int M : {128, 64}) {
    cudaMemset(d, 0xFF, 2 * 128 * N * sizeof(float));
    k_probe_cg1<<<1, 128, smem>>>(d, M);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cg1 M=%d failed: %s\n", M, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
    char name[64]; snprintf(name, sizeof(name), "cta_group::1, M = %d, N = %d", M, N);
    report(name, h, 1);
  }
  for (int M : {256, 128}) {
    cudaMemset(d, 0xFF, 2 * 128 * N * sizeof(float));
    k_probe_cg2<<<2, 128, smem>>>(d, M);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cg2 M=%d failed: %s\n", M, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, 2 * 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
    char name[64]; snprintf(name, sizeof(name), "cta_group::2, M = %d, N = %d", M, N);
    report(name, h, 2);
  }
  return 0;
}

// TENSOR engine: multi-head self-attention core on fp16 Q/K/V.
//
// out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v   with q pre-scaled by log2(e)/sqrt(hd) at weight-pack time,
// so probabilities are exp2(s - max).  (nn.MultiheadAttention core, src/model/dptn.py:16-21,46.)
//
// k_attention_f16_simt: one CTA per (sequence, head), K/V of the head staged in shared memory as fp32,
// one query per thread, online softmax in fp32.  Handles any sequence length that fits shared memory.
#include "common.cuh"
#include "tc_kernels.cuh"

namespace vatss {

template <int HD>
__global__ void __launch_bounds__(128)
k_attention_f16_simt(const __half* __restrict__ qkv, __half* __restrict__ out, SeqMap map, int N) {
  extern __shared__ float smem[];
  const int g = blockIdx.x, h = blockIdx.y;
  const int len = map.len;
  float* sK = smem;
  float* sV = smem + len * HD;
  for (int i = threadIdx.x; i < len * (HD / 2); i += blockDim.x) {
    const int t = i / (HD / 2), d2 = i - t * (HD / 2);
    const long long r = map.row(g, t);
    const __half2 k2 = *reinterpret_cast<const __half2*>(qkv + r * 3 * N + N + h * HD + 2 * d2);
    const __half2 v2 = *reinterpret_cast<const __half2*>(qkv + r * 3 * N + 2 * N + h * HD + 2 * d2);
    const float2 kf = __half22float2(k2), vf = __half22float2(v2);
    sK[t * HD + 2 * d2] = kf.x; sK[t * HD + 2 * d2 + 1] = kf.y;
    sV[t * HD + 2 * d2] = vf.x; sV[t * HD + 2 * d2 + 1] = vf.y;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < len; t += blockDim.x) {
    const long long r = map.row(g, t);
    float q[HD], acc[HD];
#pragma unroll
    for (int d = 0; d < HD; d += 2) {
      const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(qkv + r * 3 * N + h * HD + d));
      q[d] = qf.x; q[d + 1] = qf.y;
      acc[d] = 0.f; acc[d + 1] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < len; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], sK[j * HD + d], s);
      if (s > m) {
        const float c = exp2f(m - s);
        l *= c;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= c;
        m = s;
      }
      const float p = exp2f(s - m);
      l += p;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, sV[j * HD + d], acc[d]);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; d += 2)
      *reinterpret_cast<__half2*>(out + r * N + h * HD + d) = __floats2half2_rn(acc[d] * inv, acc[d + 1] * inv);
  }
}

template <int HD>
static int attention_f16_launch(const __half* qkv, __half* out, SeqMap map, int N, int heads, cudaStream_t st) {
  const size_t smem = (size_t)2 * map.len * HD * sizeof(float);
  VATSS_CHECK_ARG(smem <= 200 * 1024, "attention: sequence length %d too long", map.len);
  VATSS_CUDA_OK(cudaFuncSetAttribute(k_attention_f16_simt<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(map.G, heads);
  k_attention_f16_simt<HD><<<grid, 128, smem, st>>>(qkv, out, map, N);
  VATSS_LAUNCH_OK();
  return 0;
}

int launch_attention_f16(const __half* qkv, __half* out, SeqMap map, int N, int heads, cudaStream_t st) {
  if (map.G == 0) return 0;
  switch (N / heads) {
    case 16: return attention_f16_launch<16>(qkv, out, map, N, heads, st);
    case 32: return attention_f16_launch<32>(qkv, out, map, N, heads, st);
    default: set_error("attention: head dim %d unsupported by the tensor engine", N / heads); return -1;
  }
}

}  // namespace vatss

// Waveform front end: visual compression, Conv1d encoder + tanh-gated AV fusion fused with the
// token-major chunk segmentation, and the stand-alone (reference-layout) segmentation /
// overlap-add kernels.  HBM-bound kernels: coalesced row-wise stores, shared-memory staging of
// the waveform frames.
//
// Reference semantics: src/model/dptn_wav.py:173-184 (encoder + fusion),
// src/model/dprnn.py:122-136 (SplitToFolds), src/model/dprnn.py:145-163 (OverlapAdd).
#include "common.cuh"

namespace vatss {

// ----------------------------------------------------------------------------------------
// visual compression: vis[b,t,j] = sum_e emb_{j/(N/2)}[b,e,t] * Wv[j%(N/2),e] + bv[j%(N/2)]
// (nn.Linear(E, N/2) applied to both lip-embedding streams, then concat; dptn_wav.py:173-179)
// ----------------------------------------------------------------------------------------
// One CTA = (utterance b, stream, 64 video frames) x 64 outputs, 256 threads with a 4 x 4 register tile each.  The
// embedding slab arrives frames-contiguous (coalesced), W is transposed into [e][j] on the way into shared memory, so
// both operands of the inner product are 16-byte shared-memory reads (2 reads per 16 FMAs; the round-1 kernel read
// one W value per FMA and was bound by the shared-memory pipe: 108 us for 0.4 GFLOP).
constexpr int VC_T = 64, VC_E = 32, VC_J = 64;
__global__ void __launch_bounds__(256)
k_visual_compress(const float* __restrict__ emb1, const float* __restrict__ emb2, const float* __restrict__ Wv,
                  const float* __restrict__ bv, int E, int Tv, int N, float* __restrict__ vis) {
  __shared__ __align__(16) float sE[VC_E][VC_T];
  __shared__ __align__(16) float sW[VC_E][VC_J + 4];
  const int half = N / 2;                    // outputs per stream
  const int t0 = blockIdx.x * VC_T, which = blockIdx.y, b = blockIdx.z;
  const float* emb = (which == 0 ? emb1 : emb2) + (size_t)b * E * Tv;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // frames 4 tx .. 4 tx + 3, outputs 4 ty .. 4 ty + 3
  for (int j0 = 0; j0 < half; j0 += VC_J) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    // global -> registers -> shared memory, the next slab's loads issued before this slab's arithmetic: every load of
    // a slab is in flight at once (a load feeding a shared-memory store inside the copy loop exposed one full memory
    // latency per element: that, not the arithmetic, was the 108 us)
    constexpr int RE = VC_E * VC_T / 256, RW = VC_J * VC_E / 256;
    float re[RE], rw[RW];
    auto gload = [&](int e0) {
#pragma unroll
      for (int k = 0; k < RE; ++k) {
        const int i = threadIdx.x + 256 * k, e = i / VC_T, t = i - e * VC_T;
        re[k] = (e0 + e < E && t0 + t < Tv) ? __ldg(emb + (size_t)(e0 + e) * Tv + t0 + t) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        const int i = threadIdx.x + 256 * k, j = i / VC_E, e = i - j * VC_E;
        rw[k] = (j0 + j < half && e0 + e < E) ? __ldg(Wv + (size_t)(j0 + j) * E + e0 + e) : 0.f;
      }
    };
    gload(0);
    for (int e0 = 0; e0 < E; e0 += VC_E) {
#pragma unroll
      for (int k = 0; k < RE; ++k) {
        const int i = threadIdx.x + 256 * k, e = i / VC_T, t = i - e * VC_T;
        sE[e][t] = re[k];
      }
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        const int i = threadIdx.x + 256 * k, j = i / VC_E, e = i - j * VC_E;
        sW[e][j] = rw[k];
      }
      __syncthreads();
      if (e0 + VC_E < E) gload(e0 + VC_E);
#pragma unroll 8
      for (int e = 0; e < VC_E; ++e) {
        const float4 x = *reinterpret_cast<const float4*>(&sE[e][4 * tx]);
        const float4 w = *reinterpret_cast<const float4*>(&sW[e][4 * ty]);
        const float xs[4] = {x.x, x.y, x.z, x.w}, ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(ws[k], xs[i], acc[i][k]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + 4 * tx + i, j = j0 + 4 * ty;
      if (t < Tv) {
        float* dst = vis + ((size_t)b * Tv + t) * N + which * half + j;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (j + k < half) dst[k] = acc[i][k] + bv[j + k];
      }
    }
  }
}

int launch_visual_compress(const float* emb1, const float* emb2, const float* Wv, const float* bv, int B,
                           int E, int Tv, int N, float* vis, cudaStream_t st) {
  dim3 grid(ceil_div(Tv, VC_T), 2, B);
  k_visual_compress<<<grid, 256, 0, st>>>(emb1, emb2, Wv, bv, E, Tv, N, vis);
  VATSS_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------
// encoder (+ fusion) (+ token-major segmentation)
//   enc[b,l,n]  = sum_k W[n,k] mix[b, st*l+k]                      (nn.Conv1d, no bias)
//   enc[b,l,:] += tanh(gate) * LN(lerp(vis[b,i0,:], vis[b,i1,:]))   (F.interpolate linear,
//                                                                   align_corners=False)
//   seg[b,s,l-P*s,:] = enc[b,l,:]  for every chunk s containing frame l
// One warp per frame; lane owns features lane, lane+32, ...
// ----------------------------------------------------------------------------------------
constexpr int ENC_FRAMES_PER_BLOCK = 64;
constexpr int ENC_MAX_NI = 8;  // N <= 256

__global__ void __launch_bounds__(256)
k_encoder(const float* __restrict__ mix, const float* __restrict__ Wenc, const float* __restrict__ vis,
          const float* __restrict__ gate, const float* __restrict__ vln_w, const float* __restrict__ vln_b,
          int T, int Tv, int N, int K, int L, int S, int C, int P, float* __restrict__ enc,
          float* __restrict__ seg, __half* __restrict__ seg16, __half* __restrict__ seg16lo) {
  extern __shared__ float smem[];
  const int st = K / 2;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * ENC_FRAMES_PER_BLOCK;
  const int nfr = min(ENC_FRAMES_PER_BLOCK, L - l0);
  const int nsamp = (nfr - 1) * st + K;
  float* s_w = smem;               // [K][N]
  float* s_mix = smem + K * N;     // [nsamp]
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
    int k = i / N, n = i % N;
    s_w[i] = Wenc[n * K + k];
  }
  const float* mrow = mix + (size_t)b * T + (size_t)l0 * st;
  for (int i = threadIdx.x; i < nsamp; i += blockDim.x) s_mix[i] = mrow[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int NI = (N + 31) / 32;
  const bool av = (vis != nullptr);
  float tg = 0.f, scale = 0.f;
  if (av) {
    tg = tanhf(gate[0]);
    scale = (float)Tv / (float)L;
  }
  for (int f = warp; f < nfr; f += nwarps) {
    const int l = l0 + f;
    float e[ENC_MAX_NI];
#pragma unroll
    for (int i = 0; i < ENC_MAX_NI; ++i) {
      e[i] = 0.f;
      const int n = lane + 32 * i;
      if (i < NI && n < N) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(s_w[k * N + n], s_mix[f * st + k], acc);
        e[i] = acc;
      }
    }
    if (av) {
      float src = scale * ((float)l + 0.5f) - 0.5f;
      src = src < 0.f ? 0.f : src;
      int i0 = (int)src;
      if (i0 > Tv - 1) i0 = Tv - 1;
      const int i1 = i0 + ((i0 < Tv - 1) ? 1 : 0);
      const float lam1 = src - (float)i0, lam0 = 1.f - lam1;
      const float* v0 = vis + ((size_t)b * Tv + i0) * N;
      const float* v1 = vis + ((size_t)b * Tv + i1) * N;
      float vi[ENC_MAX_NI];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < ENC_MAX_NI; ++i) {
        vi[i] = 0.f;
        const int n = lane + 32 * i;
        if (i < NI && n < N) {
          vi[i] = lam0 * v0[n] + lam1 * v1[n];
          sum += vi[i];
        }
      }
      const float mean = warp_sum(sum) / (float)N;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < ENC_MAX_NI; ++i) {
        const int n = lane + 32 * i;
        if (i < NI && n < N) {
          const float d = vi[i] - mean;
          sq += d * d;
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / (float)N + 1e-5f);
#pragma unroll
      for (int i = 0; i < ENC_MAX_NI; ++i) {
        const int n = lane + 32 * i;
        if (i < NI && n < N) e[i] += tg * ((vi[i] - mean) * rstd * vln_w[n] + vln_b[n]);
      }
    }
    float* erow = enc + ((size_t)b * L + l) * N;
#pragma unroll
    for (int i = 0; i < ENC_MAX_NI; ++i) {
      const int n = lane + 32 * i;
      if (i < NI && n < N) erow[n] = e[i];
    }
    if (seg != nullptr || seg16 != nullptr) {
      int s_lo = (l - C + 1 + P - 1);
      s_lo = s_lo <= 0 ? 0 : s_lo / P;
      int s_hi = l / P;
      if (s_hi > S - 1) s_hi = S - 1;
      for (int s = s_lo; s <= s_hi; ++s) {
        const size_t row = ((size_t)b * S + s) * C + (l - P * s);
#pragma unroll
        for (int i = 0; i < ENC_MAX_NI; ++i) {
          const int n = lane + 32 * i;
          if (i < NI && n < N) {
            if (seg) seg[row * N + n] = e[i];
            if (seg16) {
              const __half hi = __float2half_rn(e[i]);
              seg16[row * N + n] = hi;
              if (seg16lo) seg16lo[row * N + n] = __float2half_rn(e[i] - __half2float(hi));
            }
          }
        }
      }
    }
  }
}

// Same computation for N = 32 * CH (CH = 2, 4): every lane owns CH consecutive channels, so each warp store is one
// contiguous row (16-byte fp32 / 8-byte fp16 accesses per lane) instead of CH separate 128-byte instructions.
template <int CH>
__global__ void __launch_bounds__(256)
k_encoder_vec(const float* __restrict__ mix, const float* __restrict__ Wenc, const float* __restrict__ vis,
              const float* __restrict__ gate, const float* __restrict__ vln_w, const float* __restrict__ vln_b,
              int T, int Tv, int K, int L, int S, int C, int P, float* __restrict__ enc,
              float* __restrict__ seg, __half* __restrict__ seg16, __half* __restrict__ seg16lo) {
  constexpr int N = 32 * CH;
  extern __shared__ float smem[];
  const int st = K / 2;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * ENC_FRAMES_PER_BLOCK;
  const int nfr = min(ENC_FRAMES_PER_BLOCK, L - l0);
  const int nsamp = (nfr - 1) * st + K;
  float* s_w = smem;               // [K][N]
  float* s_mix = smem + K * N;     // [nsamp]
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
    int k = i / N, n = i % N;
    s_w[i] = Wenc[n * K + k];
  }
  const float* mrow = mix + (size_t)b * T + (size_t)l0 * st;
  for (int i = threadIdx.x; i < nsamp; i += blockDim.x) s_mix[i] = mrow[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n0 = lane * CH;
  const bool av = (vis != nullptr);
  float tg = 0.f, scale = 0.f, lw[CH], lb[CH];
  if (av) {
    tg = tanhf(gate[0]);
    scale = (float)Tv / (float)L;
#pragma unroll
    for (int i = 0; i < CH; ++i) { lw[i] = vln_w[n0 + i]; lb[i] = vln_b[n0 + i]; }
  }
  for (int f = warp; f < nfr; f += nwarps) {
    const int l = l0 + f;
    float e[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) e[i] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float x = s_mix[f * st + k];
#pragma unroll
      for (int i = 0; i < CH; ++i) e[i] = fmaf(s_w[k * N + n0 + i], x, e[i]);
    }
    if (av) {
      float src = scale * ((float)l + 0.5f) - 0.5f;
      src = src < 0.f ? 0.f : src;
      int i0 = (int)src;
      if (i0 > Tv - 1) i0 = Tv - 1;
      const int i1 = i0 + ((i0 < Tv - 1) ? 1 : 0);
      const float lam1 = src - (float)i0, lam0 = 1.f - lam1;
      const float* v0 = vis + ((size_t)b * Tv + i0) * N + n0;
      const float* v1 = vis + ((size_t)b * Tv + i1) * N + n0;
      float vi[CH];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        vi[i] = lam0 * v0[i] + lam1 * v1[i];
        sum += vi[i];
      }
      const float mean = warp_sum(sum) / (float)N;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const float d = vi[i] - mean;
        sq += d * d;
      }
      const float rstd = rsqrtf(warp_sum(sq) / (float)N + 1e-5f);
#pragma unroll
      for (int i = 0; i < CH; ++i) e[i] += tg * ((vi[i] - mean) * rstd * lw[i] + lb[i]);
    }
    auto store32 = [&](float* dst) {
      if constexpr (CH == 4) *reinterpret_cast<float4*>(dst) = make_float4(e[0], e[1], e[2], e[3]);
      else *reinterpret_cast<float2*>(dst) = make_float2(e[0], e[1]);
    };
    store32(enc + ((size_t)b * L + l) * N + n0);
    if (seg != nullptr || seg16 != nullptr) {
      uint32_t hi[CH / 2], lo[CH / 2];
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) {
        const __half2 h2 = __floats2half2_rn(e[2 * i], e[2 * i + 1]);
        const float2 back = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(e[2 * i] - back.x, e[2 * i + 1] - back.y);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      auto store16 = [&](__half* dst, const uint32_t* v) {
        if constexpr (CH == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(v[0], v[1]);
        else *reinterpret_cast<uint32_t*>(dst) = v[0];
      };
      int s_lo = (l - C + 1 + P - 1);
      s_lo = s_lo <= 0 ? 0 : s_lo / P;
      int s_hi = l / P;
      if (s_hi > S - 1) s_hi = S - 1;
      for (int s = s_lo; s <= s_hi; ++s) {
        const size_t row = ((size_t)b * S + s) * C + (l - P * s);
        if (seg) store32(seg + row * N + n0);
        if (seg16) store16(seg16 + row * N + n0, hi);
        if (seg16lo) store16(seg16lo + row * N + n0, lo);
      }
    }
  }
}

int launch_encoder(const float* mix, const float* Wenc, const float* vis, const float* gate,
                   const float* vln_w, const float* vln_b, int B, int T, int Tv, int N, int K, int L, int S,
                   int C, int P, float* enc, float* seg, __half* seg16, cudaStream_t st, __half* seg16lo) {
  VATSS_CHECK_ARG(N <= 32 * ENC_MAX_NI, "encoder: num_features %d > %d unsupported", N, 32 * ENC_MAX_NI);
  dim3 grid(ceil_div(L, ENC_FRAMES_PER_BLOCK), B);
  size_t smem = ((size_t)K * N + (size_t)(ENC_FRAMES_PER_BLOCK - 1) * (K / 2) + K) * sizeof(float);
  if (N == 128)
    k_encoder_vec<4><<<grid, 256, smem, st>>>(mix, Wenc, vis, gate, vln_w, vln_b, T, Tv, K, L, S, C, P, enc, seg, seg16,
                                              seg16lo);
  else if (N == 64)
    k_encoder_vec<2><<<grid, 256, smem, st>>>(mix, Wenc, vis, gate, vln_w, vln_b, T, Tv, K, L, S, C, P, enc, seg, seg16,
                                              seg16lo);
  else
    k_encoder<<<grid, 256, smem, st>>>(mix, Wenc, vis, gate, vln_w, vln_b, T, Tv, N, K, L, S, C, P, enc, seg,
                                       seg16, seg16lo);
  VATSS_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------
// reference-layout segmentation: out[b,n,s,k] = x[b,n,P*s+k]    (SplitToFolds, exact copy)
// ----------------------------------------------------------------------------------------
__global__ void k_segment_cm(const float* __restrict__ x, int L, int S, int C, int P, long long total,
                             float* __restrict__ out) {
  const long long SC = (long long)S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bn = i / SC;
    const int r = (int)(i - bn * SC);
    const int s = r / C, k = r - s * C;
    out[i] = x[bn * L + (long long)P * s + k];
  }
}

int launch_segment_cm(const float* x, int B, int N, int L, int C, int P, float* out, cudaStream_t st) {
  const int S = (L - C) / P + 1;
  const long long total = (long long)B * N * S * C;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  k_segment_cm<<<blocks, 256, 0, st>>>(x, L, S, C, P, total, out);
  VATSS_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------
// reference-layout overlap-add: out[b,n,t] = sum_{s: 0<=t-P*s<C} y[b,n,s,t-P*s]
// Addends are accumulated in increasing s, the order F.fold's col2im uses.
// ----------------------------------------------------------------------------------------
__global__ void k_overlap_add_cm(const float* __restrict__ y, int S, int C, int P, int Lo, long long total,
                                 float* __restrict__ out) {
  const long long SC = (long long)S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bn = i / Lo;
    const int t = (int)(i - bn * Lo);
    int s_lo = t - C + 1 + P - 1;
    s_lo = s_lo <= 0 ? 0 : s_lo / P;
    int s_hi = t / P;
    if (s_hi > S - 1) s_hi = S - 1;
    float acc = 0.f;
    for (int s = s_lo; s <= s_hi; ++s) acc += y[bn * SC + (long long)s * C + (t - P * s)];
    out[i] = acc;
  }
}

int launch_overlap_add_cm(const float* y, int B, int N, int S, int C, int P, float* out, cudaStream_t st) {
  const int Lo = (S - 1) * P + C;
  const long long total = (long long)B * N * Lo;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  k_overlap_add_cm<<<blocks, 256, 0, st>>>(y, S, C, P, Lo, total, out);
  VATSS_LAUNCH_OK();
  return 0;
}

}  // namespace vatss

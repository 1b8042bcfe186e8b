// TENSOR engine: multi-head self-attention core on tcgen05 tensor cores.
//
//   out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v     (nn.MultiheadAttention core, src/model/dptn.py:16-21,46)
//
// q arrives pre-scaled by log2(e)/sqrt(hd) (folded into the in-projection at weight-pack time), so the
// probabilities are exp2(s - rowmax).
//
// Work item = (sequence, 64-feature head group): the Q/K/V tiles of a head group are 128-byte rows, i.e. plain
// SWIZZLE_128B TMA tiles; a head is a 16/32-column K-slice of them.  One persistent CTA per SM:
//   warp 0        TMA producer (K and V of the item once, Q per 128-query tile, double buffered)
//   warp 1        MMA issuer: S = Q K^T (M=128, N=kv block, K=hd) and O += P V (M=128, N=hd, K=kv block)
//   warps 2..5    softmax warpgroup 0  (thread = query row; heads [0, HPT/2) of the group)
//   warps 6..9    softmax warpgroup 1  (heads [HPT/2, HPT))
// Each warpgroup owns an S slot and an O accumulator in TMEM and a P tile in shared memory, so the two
// heads' MMAs and exponentials overlap.  Sequences longer than one kv block use an exact two-pass softmax
// (pass A: row max over all blocks, pass B: exp / P V accumulation) - no accumulator rescaling.
//
// k_attention_f16_simt is the shape-agnostic fallback (sequence too long for the shared-memory plan).
#include "common.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

// ------------------------------------------------------------------------------------------
// SIMT fallback: one CTA per (sequence, head), K/V staged in shared memory as fp32
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
k_attention_f16_simt(const __half* __restrict__ qkv, __half* __restrict__ out, SeqMap map, int N) {
  extern __shared__ float smem_f[];
  const int g = blockIdx.x, h = blockIdx.y;
  const int len = map.len;
  float* sK = smem_f;
  float* sV = smem_f + len * HD;
  for (int i = threadIdx.x; i < len * (HD / 2); i += blockDim.x) {
    const int t = i / (HD / 2), d2 = i - t * (HD / 2);
    const long long r = map.row(g, t);
    const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(qkv + r * 3 * N + N + h * HD + 2 * d2));
    const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(qkv + r * 3 * N + 2 * N + h * HD + 2 * d2));
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
static int attention_simt_launch(const __half* qkv, __half* out, SeqMap map, int N, int heads, cudaStream_t st) {
  const size_t smem = (size_t)2 * map.len * HD * sizeof(float);
  VATSS_CHECK_ARG(smem <= 200 * 1024, "attention: sequence length %d too long", map.len);
  VATSS_CUDA_OK(cudaFuncSetAttribute(k_attention_f16_simt<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(map.G, heads);
  k_attention_f16_simt<HD><<<grid, 128, smem, st>>>(qkv, out, map, N);
  VATSS_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------
struct TcAttnArgs {
  int mode;       // 0 intra, 1 inter
  int len;        // tokens per sequence
  int N;          // features (row of qkv is 3N halfs)
  int groups;     // 64-feature head groups per sequence (N / 64)
  int nblk, NB;   // kv blocks per sequence and rows per block (NB % 16 == 0, NB <= 224)
  int mtiles;     // 128-query tiles per sequence
  int pkb;        // 64-column k-blocks of the P tile (ceil(NB / 64))
  int num_items;  // sequences * groups
  SeqMap map;
  __half* out;    // (tokens, N)
};

constexpr int ATT_THREADS = 320;
constexpr uint32_t ATT_S_COLS = 224;   // S slot (<= 224 columns) then O (<= 32 columns) per warpgroup

// MN-major (N contiguous) B operand with 128-byte rows, SWIZZLE_128B: 8-row (K) groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc_sw128_mnmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO: stride between 64-element MN atoms (single atom here)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: stride between groups of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
k_tc_attention(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapKV, TcAttnArgs p) {
  constexpr int HPT = 64 / HD;       // heads per 64-feature group
  constexpr int HPW = HPT / 2;       // heads per warpgroup
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const int kvrows = p.nblk * p.NB;
  const uint32_t KV_BYTES = (uint32_t)kvrows * 128u;
  const uint32_t P_BYTES = (uint32_t)p.pkb * 16384u;
  const uint32_t sQ = base;                        // [2][128 x 128 B]
  const uint32_t sK = sQ + 2 * 16384;
  const uint32_t sV = sK + KV_BYTES;
  const uint32_t sP = sV + KV_BYTES;               // [2 warpgroups]
  const uint32_t bars = sP + 2 * P_BYTES;
  const uint32_t bar_kfull = bars, bar_kfree = bars + 8, bar_vfull = bars + 16, bar_vfree = bars + 24;
  const uint32_t bar_qfull = bars + 32, bar_qfree = bars + 48;       // [2]
  const uint32_t bar_sfull = bars + 64, bar_sfree = bars + 80;       // [2 warpgroups]
  const uint32_t bar_pfull = bars + 96, bar_pfree = bars + 112;
  const uint32_t bar_ofull = bars + 128, bar_ofree = bars + 144;
  const uint32_t tmem_slot = bars + 160;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_kfull, 1); mbar_init(bar_kfree, 1); mbar_init(bar_vfull, 1); mbar_init(bar_vfree, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qfull + 8 * i, 1); mbar_init(bar_qfree + 8 * i, 1);
      mbar_init(bar_sfull + 8 * i, 1); mbar_init(bar_sfree + 8 * i, 128);
      mbar_init(bar_pfull + 8 * i, 128); mbar_init(bar_pfree + 8 * i, 1);
      mbar_init(bar_ofull + 8 * i, 1); mbar_init(bar_ofree + 8 * i, 128);
    }
    fence_mbar_init();
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapKV);
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0, qn = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int g = item / p.groups, grp = item - g * p.groups;
        const int colq = grp * 64, colk = p.N + grp * 64, colv = 2 * p.N + grp * 64;
        int cb = 0, ck = 0;         // inter: utterance and chunk position of the sequence
        long long row0 = 0;         // intra: first token row
        if (p.mode == 0) row0 = (long long)g * p.len;
        else { cb = g / p.map.J; ck = g - cb * p.map.J; }
        auto load_rows = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int col, int r0) {
          if (p.mode == 0) tma_load_2d(dst, tm, bar, col, (int)(row0 + r0));
          else tma_load_4d(dst, tm, bar, col, ck, r0, cb);
        };
        mbar_wait(bar_kfree, (it & 1) ^ 1);
        mbar_expect_tx(bar_kfull, KV_BYTES);
        for (int j = 0; j < p.nblk; ++j) load_rows(&tmapKV, sK + j * p.NB * 128, bar_kfull, colk, j * p.NB);
        mbar_wait(bar_vfree, (it & 1) ^ 1);
        mbar_expect_tx(bar_vfull, KV_BYTES);
        for (int j = 0; j < p.nblk; ++j) load_rows(&tmapKV, sV + j * p.NB * 128, bar_vfull, colv, j * p.NB);
        for (int m = 0; m < p.mtiles; ++m, ++qn) {
          const int s = qn & 1;
          mbar_wait(bar_qfree + 8 * s, ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(bar_qfull + 8 * s, 16384);
          load_rows(&tmapQ, sQ + s * 16384, bar_qfull + 8 * s, colq, m * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = idesc_f16(128, p.NB, 0);
      const uint32_t idesc_o = idesc_f16(128, HD, 0) | (1u << 16);   // B (= V) is MN-major
      int it = 0, qn = 0;
      uint32_t n_s[2] = {0, 0}, n_p[2] = {0, 0}, n_o[2] = {0, 0};   // barrier use counters per warpgroup
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        mbar_wait(bar_kfull, it & 1);
        bool v_ready = false;
        for (int m = 0; m < p.mtiles; ++m, ++qn) {
          const int s = qn & 1;
          mbar_wait(bar_qfull + 8 * s, (qn >> 1) & 1);
          tc_fence_after();
          for (int hh = 0; hh < HPW; ++hh) {
            const int npass = p.nblk > 1 ? 2 : 1;
            for (int pass = 0; pass < npass; ++pass) {
              const bool exp_pass = pass == npass - 1;
              for (int j = 0; j < p.nblk; ++j) {
                for (int w = 0; w < 2; ++w) {
                  const int hsel = w * HPW + hh;   // head inside the 64-feature group
                  mbar_wait(bar_sfree + 8 * w, (n_s[w] & 1) ^ 1);
                  tc_fence_after();
#pragma unroll
                  for (int k16 = 0; k16 < HD / 16; ++k16) {
                    const uint32_t koff = (uint32_t)(hsel * HD * 2 + k16 * 32) >> 4;
                    const uint64_t a = smem_desc_sw128_kmajor(sQ + s * 16384) + koff;
                    const uint64_t b = smem_desc_sw128_kmajor(sK + j * p.NB * 128) + koff;
                    umma_f16<1>(tmem + w * 256, a, b, idesc_s, k16 > 0 ? 1u : 0u);
                  }
                  umma_commit(bar_sfull + 8 * w);
                  ++n_s[w];
                  // last S MMA of the item: K may be refilled while the last softmax / P V still run
                  if (w == 1 && exp_pass && j == p.nblk - 1 && hh == HPW - 1 && m == p.mtiles - 1)
                    umma_commit(bar_kfree);
                }
                if (exp_pass) {
                  if (!v_ready) { mbar_wait(bar_vfull, it & 1); v_ready = true; }
                  for (int w = 0; w < 2; ++w) {
                    const int hsel = w * HPW + hh;
                    if (j == 0) {   // first P V of this (m, head): the previous O must have been read out
                      mbar_wait(bar_ofree + 8 * w, (n_o[w] & 1) ^ 1);
                    }
                    mbar_wait(bar_pfull + 8 * w, n_p[w] & 1);
                    tc_fence_after();
                    for (int k16 = 0; k16 < p.NB / 16; ++k16) {
                      const uint64_t a = smem_desc_sw128_kmajor(sP + w * P_BYTES + (k16 >> 2) * 16384) + (uint64_t)((k16 & 3) * 2);
                      const uint64_t b = smem_desc_sw128_mnmajor(sV + (j * p.NB + k16 * 16) * 128 + hsel * HD * 2);
                      umma_f16<1>(tmem + w * 256 + ATT_S_COLS, a, b, idesc_o, (j > 0 || k16 > 0) ? 1u : 0u);
                    }
                    umma_commit(bar_pfree + 8 * w);
                    ++n_p[w];
                    if (j == p.nblk - 1) {
                      umma_commit(bar_ofull + 8 * w);
                      ++n_o[w];
                    }
                  }
                }
              }
            }
          }
          umma_commit(bar_qfree + 8 * s);    // every S MMA that reads this Q tile has been issued
        }
        umma_commit(bar_vfree);              // ... and every P V MMA that reads V of this item
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int w = (warp - 2) >> 2;               // warpgroup
    const int q = warp & 3;                      // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;                 // query row inside the 128-row tile
    const uint32_t t_s = tmem + ((uint32_t)(q * 32) << 16) + w * 256;
    const uint32_t t_o = t_s + ATT_S_COLS;
    const uint32_t sPw = sP + w * P_BYTES;
    uint32_t n_s = 0, n_p = 0, n_o = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int g = item / p.groups, grp = item - g * p.groups;
      for (int m = 0; m < p.mtiles; ++m) {
        const int qi = m * 128 + r;
        for (int hh = 0; hh < HPW; ++hh) {
          const int head = grp * HPT + w * HPW + hh;
          float mx = -INFINITY, sum = 0.f;
          const int npass = p.nblk > 1 ? 2 : 1;
          for (int pass = 0; pass < npass; ++pass) {
            const bool exp_pass = pass == npass - 1;
            for (int j = 0; j < p.nblk; ++j) {
              mbar_wait(bar_sfull + 8 * w, n_s & 1);
              ++n_s;
              tc_fence_after();
              const int kv0 = j * p.NB;
              if (!exp_pass || p.nblk == 1) {
                // row maximum over the valid columns of this block
                for (int c0 = 0; c0 < p.NB; c0 += 16) {
                  uint32_t v[16];
                  tmem_ld_32x32b_x16(t_s + c0, v);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (kv0 + c0 + i < p.len) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
              }
              if (exp_pass) {
                // the P tile of the previous block / head must have been consumed by its P V MMAs
                mbar_wait(bar_pfree + 8 * w, (n_p & 1) ^ 1);
                for (int c0 = 0; c0 < p.NB; c0 += 16) {
                  uint32_t v[16];
                  tmem_ld_32x32b_x16(t_s + c0, v);
                  tmem_ld_wait();
                  uint32_t pk[8];
#pragma unroll
                  for (int i = 0; i < 16; i += 2) {
                    const float e0 = (kv0 + c0 + i < p.len) ? ex2_fast(__uint_as_float(v[i]) - mx) : 0.f;
                    const float e1 = (kv0 + c0 + i + 1 < p.len) ? ex2_fast(__uint_as_float(v[i + 1]) - mx) : 0.f;
                    const __half2 h2 = __floats2half2_rn(e0, e1);
                    // accumulate the ROUNDED probabilities so that numerator and denominator agree
                    const float2 f2 = __half22float2(h2);
                    sum += f2.x + f2.y;
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                  const uint32_t kb = (uint32_t)c0 >> 6, ch = ((uint32_t)c0 & 63u) >> 3;
                  const uint32_t a0 = sPw + kb * 16384 + sw128_offset((uint32_t)r, ch);
                  const uint32_t a1 = sPw + kb * 16384 + sw128_offset((uint32_t)r, ch + 1);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                               "r"(pk[3]) : "memory");
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]),
                               "r"(pk[7]) : "memory");
                }
              }
              // S slot drained
              tc_fence_before();
              mbar_arrive(bar_sfree + 8 * w);
              if (exp_pass) {
                fence_proxy_async();
                mbar_arrive(bar_pfull + 8 * w);
                ++n_p;
              }
            }
          }
          // O = sum_j P_j V_j complete
          mbar_wait(bar_ofull + 8 * w, n_o & 1);
          ++n_o;
          tc_fence_after();
          uint32_t o[HD];
          if constexpr (HD == 32) tmem_ld_32x32b_x32(t_o, o);
          else tmem_ld_32x32b_x16(t_o, o);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(bar_ofree + 8 * w);
          if (qi < p.len) {
            const float inv = 1.f / sum;
            uint4* dst = reinterpret_cast<uint4*>(p.out + p.map.row(g, qi) * p.N + head * HD);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              uint32_t wd[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(__uint_as_float(o[c * 8 + 2 * e]) * inv,
                                                     __uint_as_float(o[c * 8 + 2 * e + 1]) * inv);
                wd[e] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              dst[c] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

template <int HD>
static int tc_attention_launch(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N,
                               cudaStream_t st, bool* handled) {
  *handled = false;
  TcAttnArgs a;
  a.mode = mode; a.len = map.len; a.N = N; a.groups = N / 64; a.map = map; a.out = out;
  a.nblk = (a.len + 223) / 224;
  a.NB = (((a.len + a.nblk - 1) / a.nblk) + 15) / 16 * 16;
  a.mtiles = (a.len + 127) / 128;
  a.pkb = (a.NB + 63) / 64;
  a.num_items = map.G * a.groups;
  const size_t smem = 2 * 16384 + 2 * (size_t)a.nblk * a.NB * 128 + 2 * (size_t)a.pkb * 16384 + 256;
  if (smem > 227 * 1024 || a.NB > 224) return 0;   // not handled: caller falls back
  CUtensorMap tmQ, tmKV;
  const long long tok = (long long)B * S * C;
  if (mode == 0) {
    const uint64_t dims[2] = {(uint64_t)3 * N, (uint64_t)tok};
    const uint64_t str[1] = {(uint64_t)3 * N * 2};
    const uint32_t boxq[2] = {64, 128}, boxkv[2] = {64, (uint32_t)a.NB};
    if (make_tmap_f16(&tmQ, qkv, 2, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 2, dims, str, boxkv)) return -1;
  } else {
    const uint64_t dims[4] = {(uint64_t)3 * N, (uint64_t)C, (uint64_t)S, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)3 * N * 2, (uint64_t)C * 3 * N * 2, (uint64_t)S * C * 3 * N * 2};
    const uint32_t boxq[4] = {64, 1, 128, 1}, boxkv[4] = {64, 1, (uint32_t)a.NB, 1};
    if (make_tmap_f16(&tmQ, qkv, 4, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 4, dims, str, boxkv)) return -1;
  }
  auto kern = k_tc_attention<HD>;
  static size_t configured = 0;
  if (configured < smem) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    configured = 227 * 1024;
  }
  const int grid = a.num_items < num_sms() ? a.num_items : num_sms();
  kern<<<grid, ATT_THREADS, smem, st>>>(tmQ, tmKV, a);
  VATSS_LAUNCH_OK();
  *handled = true;
  return 0;
}

int launch_attention_f16(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                         int force_simt, cudaStream_t st) {
  if (map.G == 0) return 0;
  const int hd = N / heads;
  if (!force_simt && N % 64 == 0 && (hd == 16 || hd == 32)) {
    bool handled = false;
    int rc = hd == 32 ? tc_attention_launch<32>(qkv, out, map, mode, B, S, C, N, st, &handled)
                      : tc_attention_launch<16>(qkv, out, map, mode, B, S, C, N, st, &handled);
    if (rc) return rc;
    if (handled) return 0;
  }
  switch (hd) {
    case 16: return attention_simt_launch<16>(qkv, out, map, N, heads, st);
    case 32: return attention_simt_launch<32>(qkv, out, map, N, heads, st);
    default: set_error("attention: head dim %d unsupported by the tensor engine", hd); return -1;
  }
}

}  // namespace vatss

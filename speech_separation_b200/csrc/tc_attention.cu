// TENSOR engine: multi-head self-attention core on tcgen05 tensor cores.
//
//   out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v     (nn.MultiheadAttention core, src/model/dptn.py:16-21,46)
//
// q arrives pre-scaled by log2(e)/sqrt(hd) (folded into the in-projection at weight-pack time), so the
// probabilities are exp2(s - rowmax).
//
// Work item = (sequence, 64-feature head group): the Q/K/V tiles of a head group are 128-byte rows, i.e. plain
// SWIZZLE_128B TMA tiles; a head is a 16/32-column K-slice of them (start-address offset inside the swizzle atom) and
// V is consumed as an MN-major B operand (no transpose).  One persistent CTA per SM, 576 threads:
//   warp 0        TMA producer (K, V of the item, double buffered; Q per 128-query tile, double buffered)
//   warp 1        MMA issuer: S = Q K^T (M=128, N<=192 per MMA, K=hd) and O += P V (M=128, N=hd, K=64 per kv block)
//   warps 2..17   softmax: thread = one query row x one 16-column quarter of every 64-column kv block
// The unit of work is a group = (item, query tile, head).  S and O are double buffered in TMEM (two regions of
// 192 + 32 columns) and P is a 4-deep ring in shared memory, so S of group g+1 and P V of group g run on the tensor
// pipe while the softmax warps work on group g, and the O read-out of group g is deferred until after the softmax of
// group g+1: in steady state the softmax warps never wait for an MMA.  Row maxima / sums of the four column
// quarters are combined through shared memory.  Sequences longer than 192 use an exact two-pass softmax
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
  int nblk;       // 64-row kv blocks per sequence
  int mtiles;     // 128-query tiles per sequence
  int vbufs;      // V buffers (2 when shared memory allows)
  int num_items;  // sequences * groups
  SeqMap map;
  __half* out;    // (tokens, N)
};

constexpr int ATT_THREADS = 576;       // producer warp, MMA warp, 16 softmax warps
constexpr int ATT_NB = 64;             // kv rows per block = one SWIZZLE_128B k-block of the P tile
constexpr int ATT_SUPER = 3;           // kv blocks per S job: one N <= 192 MMA per 16-wide K step
constexpr int ATT_PRING = 4;           // P tiles in flight
constexpr uint32_t ATT_REGION = 256;   // TMEM columns per pipeline region: 192 (S) + 32 (O)
constexpr uint32_t ATT_O_COL = 192;

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

// One (item, query tile, head) unit of work; every role enumerates the same sequence.
struct AttGroup {
  int item, it, m, qn, h;
  bool valid;
};

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
k_tc_attention(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapKV, TcAttnArgs p) {
  constexpr int HPT = 64 / HD;       // heads per 64-feature group
  constexpr int OC = HD / 4;         // output features per softmax thread
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t KV_BYTES = (uint32_t)p.nblk * ATT_NB * 128u;
  const uint32_t sQ = base;                                   // [2][128 x 128 B]
  const uint32_t sP = sQ + 2 * 16384;                         // [ATT_PRING][128 x 128 B]
  const uint32_t sK = sP + ATT_PRING * 16384;                 // [2]
  const uint32_t sV = sK + 2 * KV_BYTES;                      // [vbufs]
  const uint32_t bars = sV + p.vbufs * KV_BYTES;
  // barrier map, 8 bytes each: k/v/q full+free [2 each], s full/free [2], p full/free [4], o full/free [2]
  const uint32_t b_kfull = bars, b_kfree = bars + 16, b_vfull = bars + 32, b_vfree = bars + 48;
  const uint32_t b_qfull = bars + 64, b_qfree = bars + 80, b_sfull = bars + 96, b_sfree = bars + 112;
  const uint32_t b_pfull = bars + 128, b_pfree = bars + 160, b_ofull = bars + 192, b_ofree = bars + 208;
  const uint32_t tmem_slot = bars + 224;
  float* s_ex = reinterpret_cast<float*>(smem + (bars + 256 - base));   // [2 kinds][2 parity][4 quarters][128 rows]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool resident = p.nblk <= ATT_SUPER;       // all of S of a (tile, head) fits one region: single S pass
  const int nsuper = (p.nblk + ATT_SUPER - 1) / ATT_SUPER;
  const int jobs_per_group = resident ? 1 : 2 * nsuper;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_kfull + 8 * i, 1); mbar_init(b_kfree + 8 * i, 1);
      mbar_init(b_vfull + 8 * i, 1); mbar_init(b_vfree + 8 * i, 1);
      mbar_init(b_qfull + 8 * i, 1); mbar_init(b_qfree + 8 * i, 1);
      mbar_init(b_sfull + 8 * i, 1); mbar_init(b_sfree + 8 * i, 16);   // one elected arrival per softmax warp
      mbar_init(b_ofull + 8 * i, 1); mbar_init(b_ofree + 8 * i, 16);
    }
    for (int i = 0; i < ATT_PRING; ++i) { mbar_init(b_pfull + 8 * i, 16); mbar_init(b_pfree + 8 * i, 1); }
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

  auto advance = [&](AttGroup& c) {
    if (++c.h == HPT) {
      c.h = 0; ++c.qn;
      if (++c.m == p.mtiles) { c.m = 0; c.item += gridDim.x; ++c.it; c.valid = c.item < p.num_items; }
    }
  };
  const AttGroup first = {(int)blockIdx.x, 0, 0, 0, 0, (int)blockIdx.x < p.num_items};

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0, qn = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int g = item / p.groups, grp = item - g * p.groups;
        const int colq = grp * 64, colk = p.N + grp * 64, colv = 2 * p.N + grp * 64;
        int cb = 0, ck = 0;
        long long row0 = 0;
        if (p.mode == 0) row0 = (long long)g * p.len;
        else { cb = g / p.map.J; ck = g - cb * p.map.J; }
        auto load_rows = [&](const CUtensorMap* tm, uint32_t dst, uint32_t bar, int col, int r0) {
          if (p.mode == 0) tma_load_2d(dst, tm, bar, col, (int)(row0 + r0));
          else tma_load_4d(dst, tm, bar, col, ck, r0, cb);
        };
        const int kb = it & 1;
        mbar_wait(b_kfree + 8 * kb, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(b_kfull + 8 * kb, KV_BYTES);
        for (int j = 0; j < p.nblk; ++j)
          load_rows(&tmapKV, sK + kb * KV_BYTES + j * ATT_NB * 128, b_kfull + 8 * kb, colk, j * ATT_NB);
        for (int m = 0; m < p.mtiles; ++m, ++qn) {
          const int s = qn & 1;
          mbar_wait(b_qfree + 8 * s, ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(b_qfull + 8 * s, 16384);
          load_rows(&tmapQ, sQ + s * 16384, b_qfull + 8 * s, colq, m * 128);
          if (m == 0) {
            const int vb = it % p.vbufs, vuse = it / p.vbufs;
            mbar_wait(b_vfree + 8 * vb, (vuse & 1) ^ 1);
            mbar_expect_tx(b_vfull + 8 * vb, KV_BYTES);
            for (int j = 0; j < p.nblk; ++j)
              load_rows(&tmapKV, sV + vb * KV_BYTES + j * ATT_NB * 128, b_vfull + 8 * vb, colv, j * ATT_NB);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // S jobs run one job ahead of the softmax (two S regions in TMEM); P V MMAs follow the P tiles; O is double
    // buffered as well, so the softmax warps never wait for the tensor pipe in steady state.
    if (lane == 0) {
      const uint32_t idesc_o = idesc_f16(128, HD, 0) | (1u << 16);     // B (= V) is MN-major
      uint32_t sjob = 0, pjob = 0, ogrp = 0;
      int k_seen = -1, q_seen = -1, v_seen = -1;
      // job cursor: (group, job index inside the group)
      AttGroup jg = first; int jj = 0;        // next S job to issue
      AttGroup cg = first; int cj = 0;        // job whose P V MMAs come next
      auto issue_s = [&]() {
        const AttGroup& c = jg;
        if (k_seen != c.it) { mbar_wait(b_kfull + 8 * (c.it & 1), (c.it >> 1) & 1); k_seen = c.it; }
        if (q_seen != c.qn) { mbar_wait(b_qfull + 8 * (c.qn & 1), (c.qn >> 1) & 1); q_seen = c.qn; }
        const int sb = resident ? 0 : jj % nsuper;
        const int nb = min(ATT_SUPER, p.nblk - sb * ATT_SUPER);
        const uint32_t idesc_s = idesc_f16(128, nb * ATT_NB, 0);
        const uint32_t reg = sjob & 1;
        mbar_wait(b_sfree + 8 * reg, ((sjob >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16) {
          const uint32_t koff = (uint32_t)(c.h * HD * 2 + k16 * 32) >> 4;
          const uint64_t a = smem_desc_sw128_kmajor(sQ + (c.qn & 1) * 16384) + koff;
          const uint64_t b = smem_desc_sw128_kmajor(sK + (c.it & 1) * KV_BYTES + sb * ATT_SUPER * ATT_NB * 128) + koff;
          umma_f16<1>(tmem + reg * ATT_REGION, a, b, idesc_s, k16 > 0 ? 1u : 0u);
        }
        umma_commit(b_sfull + 8 * reg);
        ++sjob;
        const bool last_job = jj == jobs_per_group - 1;
        if (last_job && c.h == HPT - 1) {
          umma_commit(b_qfree + 8 * (c.qn & 1));                          // last S MMA reading this Q tile
          if (c.m == p.mtiles - 1) umma_commit(b_kfree + 8 * (c.it & 1));  // ... and this K buffer
        }
        if (++jj == jobs_per_group) { jj = 0; advance(jg); }
      };
      auto issue_pv_job = [&]() {
        const AttGroup& c = cg;
        const bool exp_job = resident || cj >= nsuper;
        if (exp_job) {
          const int sb = resident ? 0 : cj - nsuper;
          const int j1 = min(p.nblk, (sb + 1) * ATT_SUPER);
          const int vb = c.it % p.vbufs;
          if (v_seen != c.it) { mbar_wait(b_vfull + 8 * vb, (c.it / p.vbufs) & 1); v_seen = c.it; }
          const uint32_t oreg = ogrp & 1;
          for (int j = sb * ATT_SUPER; j < j1; ++j, ++pjob) {
            const uint32_t pb = pjob % ATT_PRING;
            if (j == 0) mbar_wait(b_ofree + 8 * oreg, ((ogrp >> 1) & 1) ^ 1);   // O region read out two groups ago
            mbar_wait(b_pfull + 8 * pb, (pjob / ATT_PRING) & 1);
            tc_fence_after();
            const uint32_t d_o = tmem + oreg * ATT_REGION + ATT_O_COL;
#pragma unroll
            for (int k16 = 0; k16 < ATT_NB / 16; ++k16) {
              const uint64_t a = smem_desc_sw128_kmajor(sP + pb * 16384) + (uint64_t)(k16 * 2);
              const uint64_t bv = smem_desc_sw128_mnmajor(sV + vb * KV_BYTES + (j * ATT_NB + k16 * 16) * 128 + c.h * HD * 2);
              umma_f16<1>(d_o, a, bv, idesc_o, (j > 0 || k16 > 0) ? 1u : 0u);
            }
            umma_commit(b_pfree + 8 * pb);
          }
          if (j1 == p.nblk) {
            umma_commit(b_ofull + 8 * oreg);
            ++ogrp;
            if (c.h == HPT - 1 && c.m == p.mtiles - 1) umma_commit(b_vfree + 8 * vb);   // last P V reading this V
          }
        }
        if (++cj == jobs_per_group) { cj = 0; advance(cg); }
      };
      if (jg.valid) issue_s();
      while (cg.valid) {
        if (jg.valid) issue_s();      // keep S one job ahead
        issue_pv_job();
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax warps (16): thread = row x column quarter
    const int sw = warp - 2;
    const int q = warp & 3;                      // TMEM lane quadrant of this warp
    const int cq = sw >> 2;                      // which 16 columns of each 64-column kv block
    const int r = q * 32 + lane;                 // query row inside the 128-row tile
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const int cb = cq * 16;
    uint32_t sjob = 0, pjob = 0, gcount = 0;
    // deferred epilogue of the previous group (its O accumulates while this group's softmax runs)
    bool have_prev = false; float inv_prev = 0.f; long long orow_prev = -1; int ocol_prev = 0; uint32_t oreg_prev = 0;
    auto epilogue = [&]() {
      mbar_wait(b_ofull + 8 * oreg_prev, ((gcount - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t o[OC];
      if constexpr (OC == 8) tmem_ld_32x32b_x8(t_lane + oreg_prev * ATT_REGION + ATT_O_COL + cq * OC, o);
      else tmem_ld_32x32b_x4(t_lane + oreg_prev * ATT_REGION + ATT_O_COL + cq * OC, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_ofree + 8 * oreg_prev);   // 512 same-address arrivals would serialise
      if (orow_prev >= 0) {
        uint32_t wd[OC / 2];
#pragma unroll
        for (int e = 0; e < OC / 2; ++e) {
          const __half2 h2 = __floats2half2_rn(__uint_as_float(o[2 * e]) * inv_prev, __uint_as_float(o[2 * e + 1]) * inv_prev);
          wd[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        __half* dst = p.out + orow_prev * p.N + ocol_prev;
        if constexpr (OC == 8) *reinterpret_cast<uint4*>(dst) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        else *reinterpret_cast<uint2*>(dst) = make_uint2(wd[0], wd[1]);
      }
    };
    for (AttGroup c = first; c.valid; advance(c), ++gcount) {
      const int g = c.item / p.groups, grp = c.item - g * p.groups;
      const int qi = c.m * 128 + r;
      const bool warp_live = c.m * 128 + q * 32 < p.len;     // warp-uniform: any valid query row in this warp
      float* exm = s_ex + ((gcount & 1) * 4) * 128;          // row-max exchange  [4 quarters][128]
      float* exs = s_ex + (8 + (gcount & 1) * 4) * 128;      // row-sum exchange
      float mx = -INFINITY;
      // ---- row maximum over this thread's columns
      for (int sb = 0; sb < nsuper; ++sb) {
        const uint32_t reg = (sjob + sb) & 1;
        mbar_wait(b_sfull + 8 * reg, ((sjob + sb) >> 1) & 1);
        tc_fence_after();
        const int j1 = min(p.nblk, (sb + 1) * ATT_SUPER);
        if (warp_live) {
          for (int j = sb * ATT_SUPER; j < j1; ++j) {
            const int nv = min(ATT_NB, p.len - j * ATT_NB) - cb;   // valid columns of this thread's quarter
            if (nv <= 0) continue;
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_lane + reg * ATT_REGION + (j - sb * ATT_SUPER) * ATT_NB + cb, v);
            tmem_ld_wait();
            if (nv >= 16) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
          }
        }
        if (!resident) {   // the S region is recycled
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_sfree + 8 * reg);
        }
      }
      if (!resident) sjob += nsuper;
      exm[cq * 128 + r] = mx;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      mx = fmaxf(fmaxf(exm[r], exm[128 + r]), fmaxf(exm[256 + r], exm[384 + r]));
      // ---- probabilities -> P tiles (fp16, K-major SWIZZLE_128B), consumed by the P V MMAs
      float sum = 0.f;
      for (int sb = 0; sb < nsuper; ++sb, ++sjob) {
        const uint32_t reg = sjob & 1;
        if (!resident) {
          mbar_wait(b_sfull + 8 * reg, (sjob >> 1) & 1);
          tc_fence_after();
        }
        const int j1 = min(p.nblk, (sb + 1) * ATT_SUPER);
        for (int j = sb * ATT_SUPER; j < j1; ++j, ++pjob) {
          const uint32_t pb = pjob % ATT_PRING;
          if (warp_live) {
            const int nv = min(ATT_NB, p.len - j * ATT_NB) - cb;
            uint32_t pk[8];
            if (nv > 0) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(t_lane + reg * ATT_REGION + (j - sb * ATT_SUPER) * ATT_NB + cb, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float e0 = ex2_fast(__uint_as_float(v[2 * i]) - mx);
                float e1 = ex2_fast(__uint_as_float(v[2 * i + 1]) - mx);
                if (nv < 16) {
                  e0 = (2 * i < nv) ? e0 : 0.f;
                  e1 = (2 * i + 1 < nv) ? e1 : 0.f;
                }
                sum += e0 + e1;
                const __half2 h2 = __floats2half2_rn(e0, e1);
                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[i] = 0u;
            }
            mbar_wait(b_pfree + 8 * pb, ((pjob / ATT_PRING) & 1) ^ 1);   // P tile consumed by its P V MMAs
            const uint32_t sPw = sP + pb * 16384;
            const uint32_t a0 = sPw + sw128_offset((uint32_t)r, (uint32_t)(cb >> 3));
            const uint32_t a1 = sPw + sw128_offset((uint32_t)r, (uint32_t)(cb >> 3) + 1);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
          } else {
            mbar_wait(b_pfree + 8 * pb, ((pjob / ATT_PRING) & 1) ^ 1);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (j == j1 - 1) mbar_arrive(b_sfree + 8 * reg);   // last block of the S job: region may be overwritten
            mbar_arrive(b_pfull + 8 * pb);
          }
        }
      }
      // row sums: combine the four column quarters
      exs[cq * 128 + r] = sum;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      sum = (exs[r] + exs[128 + r]) + (exs[256 + r] + exs[384 + r]);
      // ---- epilogue of the previous group, then remember this one
      if (have_prev) epilogue();
      have_prev = true;
      inv_prev = 1.f / sum;
      orow_prev = qi < p.len ? p.map.row(g, qi) : -1;
      ocol_prev = (grp * HPT + c.h) * HD + cq * OC;
      oreg_prev = gcount & 1;
    }
    if (have_prev) { epilogue(); }
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
  a.nblk = (a.len + ATT_NB - 1) / ATT_NB;
  a.mtiles = (a.len + 127) / 128;
  a.num_items = map.G * a.groups;
  const size_t kv = (size_t)a.nblk * ATT_NB * 128;
  const size_t fixed = 2 * 16384 + ATT_PRING * 16384 + 256 + 16 * 128 * 4 + 256;
  a.vbufs = (fixed + 4 * kv <= 227 * 1024) ? 2 : 1;
  const size_t smem = fixed + (2 + a.vbufs) * kv;
  if (smem > 227 * 1024) return 0;   // not handled: caller falls back
  CUtensorMap tmQ, tmKV;
  const long long tok = (long long)B * S * C;
  if (mode == 0) {
    const uint64_t dims[2] = {(uint64_t)3 * N, (uint64_t)tok};
    const uint64_t str[1] = {(uint64_t)3 * N * 2};
    const uint32_t boxq[2] = {64, 128}, boxkv[2] = {64, ATT_NB};
    if (make_tmap_f16(&tmQ, qkv, 2, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 2, dims, str, boxkv)) return -1;
  } else {
    const uint64_t dims[4] = {(uint64_t)3 * N, (uint64_t)C, (uint64_t)S, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)3 * N * 2, (uint64_t)C * 3 * N * 2, (uint64_t)S * C * 3 * N * 2};
    const uint32_t boxq[4] = {64, 1, 128, 1}, boxkv[4] = {64, 1, ATT_NB, 1};
    if (make_tmap_f16(&tmQ, qkv, 4, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 4, dims, str, boxkv)) return -1;
  }
  auto kern = k_tc_attention<HD>;
  static bool configured = false;
  if (!configured) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    configured = true;
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

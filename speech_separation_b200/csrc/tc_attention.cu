// TENSOR engine: multi-head self-attention core on tcgen05 tensor cores.
//
//   out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v     (nn.MultiheadAttention core, src/model/dptn.py:16-21,46)
//
// q arrives pre-scaled by log2(e)/sqrt(hd) (folded into the in-projection at weight-pack time), so the
// probabilities are exp2(s - rowmax).
//
// Work item = (sequence, 64-feature head group): the Q/K/V tiles of a head group are 128-byte rows, i.e. plain
// SWIZZLE_128B TMA tiles; a head is a 16/32-column K-slice of them.  One persistent CTA per SM, 20 warps:
//   warp 0        TMA producer (K and V of the item once, Q per 128-query tile, double buffered)
//   warp 1        S issuer for both warpgroups: S block = Q K_j^T (M=128, N=64, K=hd) into a ring of three
//                 64-column TMEM slots per warpgroup, running up to three blocks ahead of the softmax warps
//   warps 2,3     P V issuers, one per warpgroup: O += P_j V_j (M=128, N=hd, K=16 steps of the kv block)
//   warps 4..11   softmax warpgroup 0  (heads [0, HPT/2) of the group): two warps per TMEM lane quadrant, each
//                 thread owns one query row and one 32-column half of every 64-column S block
//   warps 12..19  softmax warpgroup 1  (heads [HPT/2, HPT))
// Each warpgroup owns its S ring, two O accumulators in TMEM and two P tiles in shared memory.  The read-out of a
// group's O is deferred until the next group's probabilities are written, so S, the exponentials and P V of
// neighbouring groups overlap.  The row maximum is exact: resident mode (<= 3 kv blocks) keeps S in TMEM between
// the max and the exp pass; longer sequences recompute S in a second pass (no accumulator rescaling).
// A ragged last query tile (<= 32 rows: 150 = 128 + 22, 283 = 2 x 128 + 27) is loaded into all four lane quadrants
// and its rows are shared by the eight warps of the warpgroup in 8-column strips.
//
// Measured (tools/attn_trace.py, tools/ubench/tcgen05_issue.cu): the kernel is bound by the per-block
// synchronisation chain (mbarrier round trips ~100 cycles each, tcgen05 issue ~30-50 cycles per instruction in
// isolation and 2-3x that next to four busy softmax warps on the same scheduler), not by the MUFU or tensor pipes.
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
  int num_items;  // sequences * groups
  int stream;     // 1: K / V super-blocks stream through a shared-memory ring (sequence too long to keep resident)
  int rag;        // 1: the last query tile has <= 32 valid rows.  It is loaded four times, once per TMEM lane quadrant,
                  // so that all eight softmax warps of a warpgroup can share its rows (8-column strips each)
                  // instead of one warp per column half doing all of it on a single SM sub-partition
  SeqMap map;
  __half* out;    // (tokens, N)
  long long* trace;   // optional clock64 trace of CTA 0 (debug), NULL in production
};

constexpr int ATT_THREADS = 640;        // producer warp, S issuer warp, 2 P V issuer warps, 2 x 8 softmax warps
constexpr int ATT_NB = 64;             // kv rows per block = one SWIZZLE_128B k-block of the P tile
constexpr int ATT_SUPER = 3;           // 64-row kv blocks per S job: one N <= 192 MMA fills the whole S region
constexpr uint32_t ATT_WG_COLS = 256;  // TMEM columns per warpgroup: 3 x 64 (S slots) + 2 x 32 (O, double buffered)
constexpr uint32_t ATT_O_COL = 192;
constexpr int ATT_RING = 4;            // streaming mode: ring stages of one super-block (3 x 64 rows x 128 B = 24 KB)
constexpr uint32_t ATT_STAGE_BYTES = ATT_SUPER * ATT_NB * 128;

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
// p = 2^(a), 2^(b) as packed fp16 (one MUFU op for the pair)
__device__ __forceinline__ uint32_t ex2_f16x2(float a, float b) {
  uint32_t packed, y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(packed) : "f"(b), "f"(a));   // low half = a, high half = b
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(packed));
  return y;
}

struct AttBars {
  uint32_t kfull, kfree, vfull, vfree, qfull, qfree;   // kfull/kfree/qfull/qfree: [2]
  uint32_t sfull, sfree;                               // [2 wg]
  uint32_t pfull, pfree;                               // [2 wg][2 buffers]
  uint32_t ofull, ofree;                               // [2 wg]
};

// MODE: 0 = resident softmax (<= 3 kv blocks: S stays in TMEM between the max and the exp pass), 1 = two-pass,
// 2 = two-pass with K / V streamed through the shared-memory ring.  Compile-time so that the hot loops carry no mode
// branches (the kernel is issue- and instruction-fetch-limited: ncu shows 9 % branch-resolving and 5 % no-instruction
// stalls).
template <int HD, int MODE, bool TRACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
k_tc_attention(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapQ32,
               const __grid_constant__ CUtensorMap tmapKV, TcAttnArgs p) {
  constexpr int HPT = 64 / HD;       // heads per 64-feature group
  constexpr int HPW = HPT / 2;       // heads per warpgroup
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t KV_BYTES = (uint32_t)p.nblk * ATT_NB * 128u;
  const uint32_t sQ = base;                        // [2][128 x 128 B]
  const uint32_t sP = sQ + 2 * 16384;              // [2 wg][2][128 x 128 B]
  const uint32_t sK = sP + 4 * 16384;              // [2] double buffered across items   (resident mode)
  const uint32_t sV = sK + 2 * KV_BYTES;
  const uint32_t sRing = sP + 4 * 16384;           // [ATT_RING] K / V super-blocks      (streaming mode)
  constexpr bool stream = MODE == 2;
  const uint32_t bars = stream ? sRing + ATT_RING * ATT_STAGE_BYTES : sV + KV_BYTES;
  AttBars B;
  B.kfull = bars; B.kfree = bars + 16;             // 2 each
  B.vfull = bars + 32; B.vfree = bars + 40;
  B.qfull = bars + 48; B.qfree = bars + 64;        // 2 each
  B.sfull = bars + 80; B.sfree = bars + 128;       // [2 wg][3 slots]
  B.pfull = bars + 176; B.pfree = bars + 208;      // [2 wg][2 buffers]
  B.ofull = bars + 240; B.ofree = bars + 272;      // [2 wg][2 buffers]
  const uint32_t tmem_slot = bars + 304;
  const uint32_t b_rfull = bars + 320, b_rfree = bars + 320 + 8 * ATT_RING;
  float* s_mx = reinterpret_cast<float*>(smem + (bars + 512 - base));   // row max / row sum exchange, 8 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool resident = MODE == 0;             // all of S of a (tile, head) fits the S region: single S pass
  const int nsuper = (p.nblk + ATT_SUPER - 1) / ATT_SUPER;

  if (threadIdx.x == 0) {
    mbar_init(B.vfull, 1); mbar_init(B.vfree, 2);      // V: one commit from each P V issuer
    for (int i = 0; i < 2; ++i) {
      mbar_init(B.kfull + 8 * i, 1); mbar_init(B.kfree + 8 * i, 1);
      mbar_init(B.qfull + 8 * i, 1); mbar_init(B.qfree + 8 * i, 1);
    }
    for (int i = 0; i < 6; ++i) { mbar_init(B.sfull + 8 * i, 1); mbar_init(B.sfree + 8 * i, 256); }
    for (int i = 0; i < 4; ++i) { mbar_init(B.pfull + 8 * i, 256); mbar_init(B.pfree + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(B.ofull + 8 * i, 1); mbar_init(B.ofree + 8 * i, 256); }
    for (int i = 0; i < ATT_RING; ++i) { mbar_init(b_rfull + 8 * i, 1); mbar_init(b_rfree + 8 * i, 2); }
    fence_mbar_init();
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapQ32);
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
      uint32_t rseq = 0;
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
        if (!stream) {
          const int kb = it & 1;
          mbar_wait(B.kfree + 8 * kb, ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(B.kfull + 8 * kb, KV_BYTES);
          for (int j = 0; j < p.nblk; ++j)
            load_rows(&tmapKV, sK + kb * KV_BYTES + j * ATT_NB * 128, B.kfull + 8 * kb, colk, j * ATT_NB);
        }
        for (int m = 0; m < p.mtiles; ++m, ++qn) {
          const int s = qn & 1;
          mbar_wait(B.qfree + 8 * s, ((qn >> 1) & 1) ^ 1);
          if (p.rag && m == p.mtiles - 1) {
            mbar_expect_tx(B.qfull + 8 * s, 16384);
            for (int k = 0; k < 4; ++k)
              load_rows(&tmapQ32, sQ + s * 16384 + k * 32 * 128, B.qfull + 8 * s, colq, m * 128);
          } else {
            mbar_expect_tx(B.qfull + 8 * s, 16384);
            load_rows(&tmapQ, sQ + s * 16384, B.qfull + 8 * s, colq, m * 128);
          }
          if (!stream) {
            if (m == 0) {
              mbar_wait(B.vfree, (it & 1) ^ 1);
              mbar_expect_tx(B.vfull, KV_BYTES);
              for (int j = 0; j < p.nblk; ++j) load_rows(&tmapKV, sV + j * ATT_NB * 128, B.vfull, colv, j * ATT_NB);
            }
          } else {
            // ring order = consumption order of the MMA warps: per head slot, pass A: K_0..K_{n-1}; pass B: K_0,V_0,K_1,V_1,...
            auto push = [&](int col, int sb) {
              const uint32_t st = rseq % ATT_RING;
              mbar_wait(b_rfree + 8 * st, ((rseq / ATT_RING) & 1) ^ 1);
              const int nb = min(ATT_SUPER, p.nblk - sb * ATT_SUPER);
              mbar_expect_tx(b_rfull + 8 * st, nb * ATT_NB * 128);
              for (int j = 0; j < nb; ++j)
                load_rows(&tmapKV, sRing + st * ATT_STAGE_BYTES + j * ATT_NB * 128, b_rfull + 8 * st, col,
                          (sb * ATT_SUPER + j) * ATT_NB);
              ++rseq;
            };
            for (int hh = 0; hh < HPW; ++hh) {
              for (int sb = 0; sb < nsuper; ++sb) push(colk, sb);
              for (int sb = 0; sb < nsuper; ++sb) { push(colk, sb); push(colv, sb); }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- S issuer (both warpgroups)
    // The S region of a warpgroup is a ring of three 64-column slots: S block jobs run up to three blocks ahead of
    // the softmax warps (into the next group / item), so S is ready when the exponentials of the previous block end.
    // Every lane runs the warp-uniform loop; one elected lane issues (see umma_f16_warp).  The issuer warps are
    // the scarce resource of this kernel (each tcgen05 instruction costs ~20 issue slots of address arithmetic,
    // competing with four softmax warps on the same scheduler), hence one warp for S and one per warpgroup for P V.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc_s = idesc_f16(128, ATT_NB, 0);
    uint32_t slot = 0, sphase = 0, gidx = 0;
    int it = 0, qn = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      if (!stream) mbar_wait(B.kfull + 8 * (it & 1), (it >> 1) & 1);
      for (int m = 0; m < p.mtiles; ++m, ++qn) {
        mbar_wait(B.qfull + 8 * (qn & 1), (qn >> 1) & 1);
        const uint32_t qbase = sQ + (qn & 1) * 16384;
        for (int hh = 0; hh < HPW; ++hh, ++gidx) {
          for (int pass = resident ? 1 : 0; pass < 2; ++pass) {
            for (int j = 0; j < p.nblk; ++j) {
              uint32_t kbase, rst = 0;
              if (!stream) {
                kbase = sK + (it & 1) * KV_BYTES + j * ATT_NB * 128;
              } else {
                // ring sequence of the K super-block: pass A: base + sb, pass B: base + nsuper + 2 sb
                const uint32_t sb = j / ATT_SUPER;
                const uint32_t rs = gidx * 3u * nsuper + (pass == 0 ? sb : nsuper + 2u * sb);
                rst = rs % ATT_RING;
                if (j % ATT_SUPER == 0) mbar_wait(b_rfull + 8 * rst, (rs / ATT_RING) & 1);
                kbase = sRing + rst * ATT_STAGE_BYTES + (j % ATT_SUPER) * ATT_NB * 128;
              }
#pragma unroll
              for (int w = 0; w < 2; ++w) {
                const uint32_t koff = (uint32_t)((w * HPW + hh) * HD * 2) >> 4;
                mbar_wait(B.sfree + 8 * (w * 3 + slot), sphase ^ 1);
                tc_fence_after();
#pragma unroll
                for (int k16 = 0; k16 < HD / 16; ++k16)
                  umma_f16_warp<1>(tmem_u + w * ATT_WG_COLS + slot * ATT_NB, smem_desc_sw128_kmajor(qbase) + koff + 2 * k16,
                                   smem_desc_sw128_kmajor(kbase) + koff + 2 * k16, idesc_s, k16 > 0 ? 1u : 0u);
                umma_commit_warp(B.sfull + 8 * (w * 3 + slot));
              }
              if (stream && (j % ATT_SUPER == ATT_SUPER - 1 || j == p.nblk - 1)) {
                umma_commit_warp(b_rfree + 8 * rst);     // the ring barriers count two readers (the P V issuers
                umma_commit_warp(b_rfree + 8 * rst);     // of a V stage); a K stage has this warp only
              }
              if (++slot == 3) { slot = 0; sphase ^= 1; }
            }
          }
        }
        umma_commit_warp(B.qfree + 8 * (qn & 1));                      // last S MMA reading this Q tile
      }
      if (!stream) umma_commit_warp(B.kfree + 8 * (it & 1));           // ... and this K buffer
    }
  } else if (warp == 2 || warp == 3) {
    // ---------------------------------------------------------------- P V issuer of warpgroup (warp - 2)
    const int w = __shfl_sync(0xffffffffu, warp, 0) - 2;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc_o = idesc_f16(128, HD, 0) | (1u << 16);     // B (= V) is MN-major
    const int nk_last = (p.len - (p.nblk - 1) * ATT_NB + 15) >> 4;   // 16-row K steps of the (ragged) last kv block
    uint32_t pjob = 0, gidx = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      if (!stream) mbar_wait(B.vfull, it & 1);
      for (int m = 0; m < p.mtiles; ++m) {
        for (int hh = 0; hh < HPW; ++hh, ++gidx) {
          const uint32_t ob = gidx & 1;
          const uint32_t d_o = tmem_u + w * ATT_WG_COLS + ATT_O_COL + ob * 32;
          const uint32_t hoff = (uint32_t)((w * HPW + hh) * HD * 2);
          mbar_wait(B.ofree + 8 * (w * 2 + ob), ((gidx >> 1) & 1) ^ 1);   // O buffer read out by the softmax warps
          for (int j = 0; j < p.nblk; ++j, ++pjob) {
            uint32_t vbase, rst = 0;
            if (!stream) {
              vbase = sV + j * ATT_NB * 128;
            } else {
              const uint32_t sb = j / ATT_SUPER;
              const uint32_t rs = gidx * 3u * nsuper + nsuper + 2u * sb + 1u;
              rst = rs % ATT_RING;
              if (j % ATT_SUPER == 0) mbar_wait(b_rfull + 8 * rst, (rs / ATT_RING) & 1);
              vbase = sRing + rst * ATT_STAGE_BYTES + (j % ATT_SUPER) * ATT_NB * 128;
            }
            const uint32_t pb = pjob & 1;
            const uint64_t adesc = smem_desc_sw128_kmajor(sP + (w * 2 + pb) * 16384);
            const uint64_t bdesc = smem_desc_sw128_mnmajor(vbase + hoff);
            const int nk = j == p.nblk - 1 ? nk_last : ATT_NB / 16;
            mbar_wait(B.pfull + 8 * (w * 2 + pb), (pjob >> 1) & 1);
            tc_fence_after();
            for (int k16 = 0; k16 < nk; ++k16)
              umma_f16_warp<1>(d_o, adesc + 2 * k16, bdesc + ((k16 * 16 * 128) >> 4), idesc_o, (j > 0 || k16 > 0) ? 1u : 0u);
            umma_commit_warp(B.pfree + 8 * (w * 2 + pb));
            if (stream && (j % ATT_SUPER == ATT_SUPER - 1 || j == p.nblk - 1)) umma_commit_warp(b_rfree + 8 * rst);
          }
          umma_commit_warp(B.ofull + 8 * (w * 2 + ob));
        }
      }
      if (!stream) umma_commit_warp(B.vfree);                          // last P V reading this item's V
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int sw = warp - 4;
    const int w = sw >> 3;                       // warpgroup
    const int set = (sw >> 2) & 1;               // which 32-column half of each 64-column kv block this thread owns
    const int q = warp & 3;                      // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;                 // query row inside the 128-row tile
    const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + w * ATT_WG_COLS;
    const int cb = set * 32;
    float* ex_mx = s_mx + w * 512;               // [2 parities][2 sets][128 rows]
    float* ex_sum = s_mx + 1024 + w * 512;
    uint32_t sseq = 0, pjob = 0, gidx = 0;
    // the read-out of a group's O accumulator is deferred until the next group's probabilities are written, so the
    // P V MMAs of one group overlap the exponentials of the next
    struct Pending { bool valid, rag; float sum; long long row; int head; uint32_t gidx; } pend = {false, false, 0.f, 0, 0, 0u};
    auto read_out = [&](const Pending& pd) {
      const uint32_t ob = pd.gidx & 1;
      mbar_wait(B.ofull + 8 * (w * 2 + ob), (pd.gidx >> 1) & 1);
      tc_fence_after();
      uint32_t o[HD / 2];
      if constexpr (HD == 32) tmem_ld_32x32b_x16(t_base + ATT_O_COL + ob * 32 + set * 16, o);
      else tmem_ld_32x32b_x8(t_base + ATT_O_COL + ob * 32 + set * 8, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(B.ofree + 8 * (w * 2 + ob));
      if (pd.row >= 0) {
        float tot = pd.sum;
        if (pd.rag) {
          tot = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) tot += ex_sum[ob * 256 + i * 32 + lane];
        } else {
          tot += ex_sum[ob * 256 + (set ^ 1) * 128 + r];
        }
        const float inv = 1.f / tot;
        uint4* dst = reinterpret_cast<uint4*>(p.out + pd.row * p.N + pd.head * HD + set * (HD / 2));
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
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
    };
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int g = item / p.groups, grp = item - g * p.groups;
      for (int m = 0; m < p.mtiles; ++m) {
        // Ragged last tile (<= 32 queries, replicated in every lane quadrant): lane = query, and the eight warps of
        // the warpgroup split each 64-column block into 8-column strips (cc) instead of 32-column halves.
        const bool rag = p.rag && m == p.mtiles - 1;
        const int qi = m * 128 + (rag ? lane : r);
        const bool warp_live = rag || m * 128 + q * 32 < p.len;     // warp-uniform: any valid query row in this warp
        const int cc = rag ? cb + q * 8 : cb;                        // first column (inside a block) of this thread
        const int cw = rag ? 8 : 32;                                 // columns per block of this thread
        for (int hh = 0; hh < HPW; ++hh, ++gidx) {
          const int head = grp * HPT + w * HPW + hh;
          const uint32_t par = gidx & 1;
          long long* T = (TRACE && blockIdx.x == 0 && sw == 0 && lane == 0 && gidx >= 8 && gidx < 12)   // debug timeline
                             ? p.trace + (gidx - 8) * 32 : nullptr;
          if (T) T[0] = clock64();
          float mx = -INFINITY;
          // ---- row maximum over this thread's columns (resident: the blocks stay in their slots for the exp pass)
          for (int j = 0; j < p.nblk; ++j) {
            const uint32_t sq = sseq + j, slot = sq % 3u;
            mbar_wait(B.sfull + 8 * (w * 3 + slot), (sq / 3u) & 1);
            tc_fence_after();
            const int nv = min(ATT_NB, p.len - j * ATT_NB) - cc;   // valid columns of this thread's strip
            if (warp_live && nv > 0) {
              if (rag) {
                uint32_t v[8];
                tmem_ld_32x32b_x8(t_base + slot * ATT_NB + cc, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
              } else {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_base + slot * ATT_NB + cc, v);
                tmem_ld_wait();
                if (nv >= 32) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
              }
            }
            if (!resident) {   // pass A of the two-pass softmax: the slot is recycled right away
              tc_fence_before();
              mbar_arrive(B.sfree + 8 * (w * 3 + slot));
            }
          }
          if (!resident) sseq += p.nblk;
          if (T) T[1] = clock64();
          // combine the partial maxima of the row through shared memory (named barrier of this warpgroup)
          ex_mx[par * 256 + set * 128 + r] = mx;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
          if (rag) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, ex_mx[par * 256 + i * 32 + lane]);
          } else {
            mx = fmaxf(mx, ex_mx[par * 256 + (set ^ 1) * 128 + r]);
          }
          if (T) T[2] = clock64();
          // ---- probabilities -> P tiles (fp16, K-major SWIZZLE_128B), consumed by the P V MMAs
          float sum = 0.f;
          for (int j = 0; j < p.nblk; ++j, ++sseq, ++pjob) {
            const uint32_t slot = sseq % 3u, pb = pjob & 1;
            if (!resident) {
              mbar_wait(B.sfull + 8 * (w * 3 + slot), (sseq / 3u) & 1);
              tc_fence_after();
            }
            const int nv = min(ATT_NB, p.len - j * ATT_NB) - cc;
            const uint32_t sPw = sP + (w * 2 + pb) * 16384;
            if (rag) {
              uint32_t pk[4] = {0u, 0u, 0u, 0u};
              if (nv > 0) {
                uint32_t v[8];
                tmem_ld_32x32b_x8(t_base + slot * ATT_NB + cc, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float e0 = ex2_fast(__uint_as_float(v[2 * i]) - mx);
                  float e1 = ex2_fast(__uint_as_float(v[2 * i + 1]) - mx);
                  e0 = (2 * i < nv) ? e0 : 0.f;
                  e1 = (2 * i + 1 < nv) ? e1 : 0.f;
                  sum += e0 + e1;
                  const __half2 h2 = __floats2half2_rn(e0, e1);
                  pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              }
              tc_fence_before();
              mbar_arrive(B.sfree + 8 * (w * 3 + slot));
              mbar_wait(B.pfree + 8 * (w * 2 + pb), ((pjob >> 1) & 1) ^ 1);
              // P row = query = lane (tile rows 0..31); the read-out is done by the quadrant-0 warps
              const uint32_t a0 = sPw + sw128_offset((uint32_t)lane, (uint32_t)(cc >> 3));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                           "r"(pk[3]) : "memory");
              fence_proxy_async();
            } else {
              uint32_t pk[16];
              // (probe the P buffer's barrier now: its ~100-cycle round trip overlaps the exponentials)
              const bool p_free = mbar_try_wait(B.pfree + 8 * (w * 2 + pb), ((pjob >> 1) & 1) ^ 1);
              if (warp_live && nv > 0) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_base + slot * ATT_NB + cc, v);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(B.sfree + 8 * (w * 3 + slot));     // values are in registers: the slot may be refilled
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float e0 = ex2_fast(__uint_as_float(v[2 * i]) - mx);
                  float e1 = ex2_fast(__uint_as_float(v[2 * i + 1]) - mx);
                  if (nv < 32) {
                    e0 = (2 * i < nv) ? e0 : 0.f;
                    e1 = (2 * i + 1 < nv) ? e1 : 0.f;
                  }
                  sum += e0 + e1;
                  const __half2 h2 = __floats2half2_rn(e0, e1);
                  pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              } else {
                tc_fence_before();
                mbar_arrive(B.sfree + 8 * (w * 3 + slot));
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              }
              // the P buffer of two blocks ago must have been consumed by its P V MMAs
              if (!p_free) mbar_wait(B.pfree + 8 * (w * 2 + pb), ((pjob >> 1) & 1) ^ 1);
              if (warp_live) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint32_t a0 = sPw + sw128_offset((uint32_t)r, (uint32_t)(cb >> 3) + c);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                               "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                }
                fence_proxy_async();
              }
            }
            mbar_arrive(B.pfull + 8 * (w * 2 + pb));
            if (T && j < 5) T[8 + j] = clock64();
          }
          if (T) T[3] = clock64();
          ex_sum[par * 256 + set * 128 + r] = sum;     // partial row sums, combined at read-out time (after a barrier)
          if (pend.valid) read_out(pend);
          if (T) T[4] = clock64();
          pend.valid = true; pend.sum = sum; pend.head = head; pend.gidx = gidx; pend.rag = rag;
          // ragged tile: O rows 0..31 are the queries, read out by the quadrant-0 warps
          pend.row = (warp_live && qi < p.len && (!rag || q == 0)) ? p.map.row(g, qi) : -1;
        }
      }
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");
    if (pend.valid) read_out(pend);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

template <int HD, int MODE, bool TRACE = false>
static int tc_attention_launch_mode(const CUtensorMap& tmQ, const CUtensorMap& tmQ32, const CUtensorMap& tmKV,
                                    const TcAttnArgs& a, size_t smem, cudaStream_t st) {
  if (!TRACE && a.trace != nullptr) return tc_attention_launch_mode<HD, MODE, true>(tmQ, tmQ32, tmKV, a, smem, st);
  auto kern = k_tc_attention<HD, MODE, TRACE>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  }
  const int grid = a.num_items < grid_cap() ? a.num_items : grid_cap();
  kern<<<grid, ATT_THREADS, smem, st>>>(tmQ, tmQ32, tmKV, a);
  VATSS_LAUNCH_OK();
  return 0;
}

template <int HD>
static int tc_attention_launch(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N,
                               cudaStream_t st, bool* handled) {
  *handled = false;
  TcAttnArgs a;
  a.trace = g_lstm_trace;
  a.mode = mode; a.len = map.len; a.N = N; a.groups = N / 64; a.map = map; a.out = out;
  a.nblk = (a.len + ATT_NB - 1) / ATT_NB;
  a.mtiles = (a.len + 127) / 128;
  a.num_items = map.G * a.groups;
  const size_t fixed = 2 * 16384 + 4 * 16384 + 512 + 8192;
  size_t smem = fixed + 3 * (size_t)a.nblk * ATT_NB * 128;
  a.stream = 0;
  if (smem > 227 * 1024) {   // K / V do not fit: stream them through the ring (two-pass softmax re-reads K)
    a.stream = 1;
    smem = fixed + (size_t)ATT_RING * ATT_STAGE_BYTES;
  }
  a.rag = a.len - (a.mtiles - 1) * 128 <= 32;
  CUtensorMap tmQ, tmQ32, tmKV;
  const long long tok = (long long)B * S * C;
  if (mode == 0) {
    const uint64_t dims[2] = {(uint64_t)3 * N, (uint64_t)tok};
    const uint64_t str[1] = {(uint64_t)3 * N * 2};
    const uint32_t boxq[2] = {64, 128}, boxq32[2] = {64, 32}, boxkv[2] = {64, ATT_NB};
    if (make_tmap_f16(&tmQ, qkv, 2, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmQ32, qkv, 2, dims, str, boxq32)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 2, dims, str, boxkv)) return -1;
  } else {
    const uint64_t dims[4] = {(uint64_t)3 * N, (uint64_t)C, (uint64_t)S, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)3 * N * 2, (uint64_t)C * 3 * N * 2, (uint64_t)S * C * 3 * N * 2};
    const uint32_t boxq[4] = {64, 1, 128, 1}, boxq32[4] = {64, 1, 32, 1}, boxkv[4] = {64, 1, ATT_NB, 1};
    if (make_tmap_f16(&tmQ, qkv, 4, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmQ32, qkv, 4, dims, str, boxq32)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 4, dims, str, boxkv)) return -1;
  }
  int rc;
  if (a.stream) rc = tc_attention_launch_mode<HD, 2>(tmQ, tmQ32, tmKV, a, smem, st);
  else if (a.nblk <= ATT_SUPER) rc = tc_attention_launch_mode<HD, 0>(tmQ, tmQ32, tmKV, a, smem, st);
  else rc = tc_attention_launch_mode<HD, 1>(tmQ, tmQ32, tmKV, a, smem, st);
  if (rc) return rc;
  *handled = true;
  return 0;
}

int g_attention_version = 3;

int launch_attention_f16(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                         int force_simt, cudaStream_t st) {
  if (map.G == 0) return 0;
  const int hd = N / heads;
  if (!force_simt && N % 64 == 0 && (hd == 16 || hd == 32)) {
    if (g_attention_version == 3) return launch_attention_v3(qkv, out, map, mode, B, S, C, N, heads, st);
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

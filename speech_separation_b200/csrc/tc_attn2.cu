// TENSOR engine, attention core v2: softmax probabilities stay in TMEM (round 2).
//
//   out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v     (nn.MultiheadAttention core, src/model/dptn.py:16-21,46)
//
// q arrives pre-scaled by log2(e)/sqrt(hd) (folded into the in-projection at weight-pack time): P = exp2(s - max).
//
// Why a second kernel: v1 (tc_attention.cu) walks 64-column S blocks, hands P to the P V MMAs through shared memory
// (st.shared + fence.proxy.async + mbarrier per block) and is bound by that per-block synchronisation chain (0.31 /
// 0.41 of the MUFU floor).  Here a JOB is one (128-query tile, kv block of NB <= 96 keys, head): one S MMA group
// fills a TMEM slot, one thread per query row reduces its whole row (no cross-thread max / sum exchange), overwrites
// S in place with P as packed fp16 (tcgen05.st) and the P V MMAs read their A operand straight from TMEM.
// Jobs are self-contained: each carries its own row maximum and row sum, so the kv blocks of a sequence (150 ->
// 2 x 80, 283 -> 3 x 96, 710 -> 8 x 96) and the column strips of a ragged query tile are merged by the epilogue with
// O = sum_p O_p 2^(m_p - m) / sum_p l_p 2^(m_p - m): exact softmax, no accumulator rescaling, no second S pass, no
// resident / two-pass / streaming modes.  K / V blocks stream through a 4-stage TMA ring for every sequence length.
//
// Work item = (sequence, 64-feature head group); persistent, one CTA per SM, 12 warps:
//   warp 0        TMA producer;  warp 1  S issuer;  warps 2, 3  P V issuers, one per softmax warpgroup
//   warps 4..7    softmax warpgroup 0 (even heads of the group); warps 8..11 softmax warpgroup 1 (odd heads); each
//                 thread also reads out the O rows it produced (deferred behind the next job's softmax), merges
//                 kv blocks / strips, normalises and stores - no separate epilogue role, no statistics hand-off
// Few roles and compact (not unrolled) softmax loops on purpose: the 16-warp / 5-role version of this kernel spent
// a third of its issue-stall samples on instruction-cache misses (62 KB of SASS, profiles/r02_*).
// Each softmax warpgroup owns a ring of TWO S / P slots: S(i + 2) is issued as soon as P V(i) has completed, so the
// chain P(i) ready -> P V(i) -> S(i + 2) -> S full (measured with the first cut of this kernel, one slot per
// warpgroup: 2000 - 2400 cycles, as long as the softmax of a job itself) runs while the warpgroup works on job i + 1.
// TMEM (512 columns): S / P slot (w, r) at 96 (2 w + r), O accumulator (w, r) at 384 + 32 (2 w + r).
// A ragged last query tile (<= 32 rows) is loaded into all four lane quadrants; quadrant q handles a strip of the kv
// block (zeros elsewhere in its P rows) and the epilogue merges the four partial results through shared memory.
#include "common.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

struct Attn2Args {
  int mode;       // 0 intra, 1 inter
  int len;        // tokens per sequence
  int N;          // features (row of qkv is 3N halfs)
  int groups;     // 64-feature head groups per sequence
  int nblk;       // kv blocks per sequence
  int NB;         // keys per kv block (multiple of 16, <= 96)
  int mtiles;     // 128-query tiles per sequence
  int num_items;  // sequences * groups
  int rag;        // last query tile has <= 32 rows: replicated-quadrant strip mode
  SeqMap map;
  __half* out;    // (tokens, N)
  long long* trace;   // optional clock64 timeline of CTA 0 (debug), NULL in production
};

constexpr int A2_THREADS = 384;
constexpr int A2_NSTG = 6;
constexpr uint32_t A2_SLOT = 96;        // TMEM columns of one S / P slot
constexpr uint32_t A2_OCOL = 384;
constexpr uint32_t A2_QBYTES = 16384;

__device__ __forceinline__ uint64_t a2_desc_mnmajor(uint32_t smem_addr) {   // V: kv rows of 128 B, features contiguous
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float a2_max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ uint32_t a2_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Position in the job stream of one CTA.  STEP = 1: every job (S issuer); STEP = 2: the jobs of softmax warpgroup w
// (heads w, w + 2, ...).  qn / kvn count the Q tiles and kv stages consumed so far (ring positions and phases).
template <int HPG, int STEP>
struct A2Job {
  int item, m, j, hh, h0;
  uint32_t qn, kvn;
  bool valid;
  __device__ __forceinline__ void init(const Attn2Args& p, int first_head) {
    item = blockIdx.x; m = j = 0; hh = h0 = first_head; qn = kvn = 0; valid = item < p.num_items;
  }
  __device__ __forceinline__ void next(const Attn2Args& p) {
    hh += STEP;
    if (hh < HPG) return;
    hh = h0; ++kvn;
    if (++j < p.nblk) return;
    j = 0; ++qn;
    if (++m < p.mtiles) return;
    m = 0; item += gridDim.x;
    valid = item < p.num_items;
  }
};

// debug timeline: jobs [A2_T0, A2_T0 + 8) of softmax warpgroup 0 in CTA 0, 16 slots per job
constexpr uint32_t A2_T0 = 24;
#define A2_MARK(cond, i, k)                                                                    \
  do {                                                                                         \
    if (TRACE && blockIdx.x == 0 && (cond) && (i) >= A2_T0 && (i) < A2_T0 + 8 && lane == 0)    \
      p.trace[((i) - A2_T0) * 16 + (k)] = clock64();                                           \
  } while (0)

template <int HD, bool TRACE>
__global__ void __launch_bounds__(A2_THREADS, 1)
k_tc_attn2(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapQ32,
           const __grid_constant__ CUtensorMap tmapKV, Attn2Args p) {
  constexpr int HPG = 64 / HD;       // heads per 64-feature group (2 or 4: even, so head hh belongs to warpgroup hh & 1)
  constexpr int HPW = HPG / 2;       // heads per warpgroup
  constexpr int XW = (HD + 2) * 32;  // floats of one quadrant's partial in the ragged-tile merge buffer
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t KVB = (uint32_t)p.NB * 128u;                 // bytes of one K (or V) block
  const uint32_t sQ = base;                                    // [2][128 x 128 B]
  const uint32_t sKV = sQ + 2 * A2_QBYTES;                     // [A2_NSTG][K block | V block]
  const uint32_t xoff = 2 * A2_QBYTES + A2_NSTG * 2 * KVB;     // ragged-tile merge buffers [2 wg][4 quadrants][HD + 2][32] floats
  const uint32_t bars = base + xoff + 2 * 4 * XW * 4;
  const uint32_t q_full = bars, q_free = bars + 16;            // [2] each
  const uint32_t kv_full = bars + 32, kv_free = bars + 80;     // [A2_NSTG <= 6] each
  const uint32_t s_full = bars + 128, p_ready = bars + 160, o_full = bars + 192;   // [2 wg][2 slots] each
  const uint32_t tmem_slot = bars + 224;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(q_full + 8 * i, 1); mbar_init(q_free + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(s_full + 8 * i, 1); mbar_init(p_ready + 8 * i, 4);     // one arrival per softmax warp
      mbar_init(o_full + 8 * i, 1);
    }
    for (int i = 0; i < A2_NSTG; ++i) { mbar_init(kv_full + 8 * i, 1); mbar_init(kv_free + 8 * i, 2); }   // one commit per P V issuer
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

  // Register budget (launch: 384 x 168): the control warpgroup hands registers to the two softmax warpgroups, whose
  // threads keep a whole S row (<= 96 columns) in registers.  Each setmaxnreg sits at the top of its role's branch
  // (ptxas budgets the code that follows it and takes the minimum where branches with different budgets merge).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t qn = 0, kvn = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
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
        for (int m = 0; m < p.mtiles; ++m, ++qn) {
          const uint32_t qb = qn & 1;
          mbar_wait(q_free + 8 * qb, ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(q_full + 8 * qb, A2_QBYTES);
          if (p.rag && m == p.mtiles - 1) {
            for (int k = 0; k < 4; ++k) load_rows(&tmapQ32, sQ + qb * A2_QBYTES + k * 4096, q_full + 8 * qb, colq, m * 128);
          } else {
            load_rows(&tmapQ, sQ + qb * A2_QBYTES, q_full + 8 * qb, colq, m * 128);
          }
          for (int j = 0; j < p.nblk; ++j, ++kvn) {
            const uint32_t st = kvn % A2_NSTG;
            mbar_wait(kv_free + 8 * st, ((kvn / A2_NSTG) & 1) ^ 1);
            A2_MARK(HPG == 2, kvn, 14);
            mbar_expect_tx(kv_full + 8 * st, 2 * KVB);
            load_rows(&tmapKV, sKV + st * 2 * KVB, kv_full + 8 * st, colk, j * p.NB);
            load_rows(&tmapKV, sKV + st * 2 * KVB + KVB, kv_full + 8 * st, colv, j * p.NB);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- S issuer (both warpgroups)
    // S(i) of warpgroup w goes to slot (w, i & 1) once P V(i - 2), which read P from that slot, has completed.
    // Warp-uniform control flow, one elected lane issues (umma_*_warp).
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc_s = idesc_f16(128, p.NB, 0);
    A2Job<HPG, 1> it;
    it.init(p, 0);
    uint32_t iw0 = 0, iw1 = 0;
    while (it.valid) {
      const uint32_t w = it.hh & 1, i = w ? iw1 : iw0, r = i & 1;
      A2_MARK(w == 0, i, 10);
      if (it.j == 0 && it.hh == 0) mbar_wait_warp(q_full + 8 * (it.qn & 1), (it.qn >> 1) & 1);
      if (it.hh == 0) mbar_wait_warp(kv_full + 8 * (it.kvn % A2_NSTG), (it.kvn / A2_NSTG) & 1);
      A2_MARK(w == 0, i, 11);
      if (i >= 2) mbar_wait_warp(o_full + 8 * (2 * w + r), ((i >> 1) - 1) & 1);
      tc_fence_after();
      A2_MARK(w == 0, i, 12);
      const uint64_t qd = smem_desc_sw128_kmajor(sQ + (it.qn & 1) * A2_QBYTES) + ((uint32_t)(it.hh * HD * 2) >> 4);
      const uint64_t kd = smem_desc_sw128_kmajor(sKV + (it.kvn % A2_NSTG) * 2 * KVB) + ((uint32_t)(it.hh * HD * 2) >> 4);
#pragma unroll
      for (int k16 = 0; k16 < HD / 16; ++k16)
        umma_f16_warp<1>(tmem_u + (2 * w + r) * A2_SLOT, qd + 2 * k16, kd + 2 * k16, idesc_s, k16 > 0 ? 1u : 0u);
      umma_commit_warp(s_full + 8 * (2 * w + r));
      if (it.j == p.nblk - 1 && it.hh == HPG - 1) umma_commit_warp(q_free + 8 * (it.qn & 1));   // last S MMA on this Q tile
      A2_MARK(w == 0, i, 13);
      if (w) ++iw1; else ++iw0;
      it.next(p);
    }
  } else {
    // ---------------------------------------------------------------- P V issuer of softmax warpgroup w
    // O(i) goes to accumulator (w, i & 1): the read-out of O(i - 2) precedes the P(i) arrival in program order of
    // every softmax thread, so no separate "O free" barrier is needed.
    const int w = __shfl_sync(0xffffffffu, warp, 0) - 2;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc_o = idesc_f16(128, HD, 0) | (1u << 16);      // B (= V) is MN-major
    A2Job<HPG, 2> it;
    it.init(p, w);
    for (uint32_t i = 0; it.valid; ++i) {
      const uint32_t r = i & 1;
      mbar_wait_warp(p_ready + 8 * (2 * w + r), (i >> 1) & 1);
      tc_fence_after();
      A2_MARK(w == 0, i, 4);
      const uint32_t vbase = sKV + (it.kvn % A2_NSTG) * 2 * KVB + KVB + (uint32_t)(it.hh * HD * 2);
      const uint64_t vd = a2_desc_mnmajor(vbase);
      const int nv = min(p.NB, p.len - it.j * p.NB);
      const int nk = (nv + 15) >> 4;                            // P columns beyond the sequence are never multiplied
      for (int k16 = 0; k16 < nk; ++k16)
        umma_f16_ts_warp(tmem_u + A2_OCOL + (2 * w + r) * 32, tmem_u + (2 * w + r) * A2_SLOT + 8 * k16,
                         vd + (uint32_t)((k16 * 16 * 128) >> 4), idesc_o, k16 > 0 ? 1u : 0u);
      umma_commit_warp(o_full + 8 * (2 * w + r));
      if (it.hh == HPG - 2 + w) umma_commit_warp(kv_free + 8 * (it.kvn % A2_NSTG));   // this warp's last MMA on the stage
      A2_MARK(w == 0, i, 6);
      it.next(p);
    }
  }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups (+ their own read-out)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int w = (warp >> 2) - 1;
    const int q = warp & 3;                                     // TMEM lane quadrant
    const int r_tile = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* xbuf = reinterpret_cast<float*>(smem + xoff) + w * 4 * XW;
    // ragged strips: the kv block is NB / 8 eight-column units, shared out over the four quadrants (<= 3 each)
    const int nu8 = p.NB >> 3;
    const int u_cnt = nu8 / 4 + (q < (nu8 & 3) ? 1 : 0);
    const int u_first = q * (nu8 / 4) + min(q, nu8 & 3);
    // merge state of the heads of this warpgroup: running (max, sum, O) over the kv blocks of a (tile, head)
    float Mr[HPW], Lr[HPW], Or[HPW][HD];
    // the job whose O is still to be read out (deferred behind the next job's softmax)
    bool pend = false, p_rag = false, p_first = false, p_last = false;
    uint32_t p_i = 0;
    float p_mx = 0.f, p_sum = 0.f;
    long long p_off = -1;
    // read-out of job p_i: O accumulator (w, p_i & 1) -> merge -> (last kv block) normalise and store
    auto read_out = [&](float& M, float& L, float* O) {
      const uint32_t rb = p_i & 1;
      mbar_wait(o_full + 8 * (2 * w + rb), (p_i >> 1) & 1);
      tc_fence_after();
      A2_MARK(w == 0 && q == 0, p_i, 8);
      uint32_t o[HD];
      if constexpr (HD == 32) tmem_ld_32x32b_x32(t_lane + A2_OCOL + (2 * w + rb) * 32, o);
      else tmem_ld_32x32b_x16(t_lane + A2_OCOL + (2 * w + rb) * 32, o);
      tmem_ld_wait();
      if (p_first) {
        M = p_mx; L = p_sum;
#pragma unroll
        for (int c = 0; c < HD; ++c) O[c] = __uint_as_float(o[c]);
      } else {
        const float mn = fmaxf(M, p_mx);
        const float a = a2_ex2(M - mn), b = a2_ex2(p_mx - mn);
        M = mn;
        L = L * a + p_sum * b;
#pragma unroll
        for (int c = 0; c < HD; ++c) O[c] = O[c] * a + __uint_as_float(o[c]) * b;
      }
      if (p_last) {
        if (!p_rag) {
          if (p_off >= 0) {
            const float inv = 1.f / L;
            uint4* dst = reinterpret_cast<uint4*>(p.out + p_off);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              uint4 wd;
              wd.x = a2_pack(O[8 * c] * inv, O[8 * c + 1] * inv);
              wd.y = a2_pack(O[8 * c + 2] * inv, O[8 * c + 3] * inv);
              wd.z = a2_pack(O[8 * c + 4] * inv, O[8 * c + 5] * inv);
              wd.w = a2_pack(O[8 * c + 6] * inv, O[8 * c + 7] * inv);
              dst[c] = wd;
            }
          }
        } else {
          // four strip partials of row `lane` live in the four quadrants: merge through shared memory, then
          // quadrant q finishes features [q HD/4, (q+1) HD/4) of the row
#pragma unroll
          for (int c = 0; c < HD; ++c) xbuf[q * XW + c * 32 + lane] = O[c];
          xbuf[q * XW + HD * 32 + lane] = M;
          xbuf[q * XW + (HD + 1) * 32 + lane] = L;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
          float mt = -1e30f;
#pragma unroll
          for (int s = 0; s < 4; ++s) mt = fmaxf(mt, xbuf[s * XW + HD * 32 + lane]);
          float wt[4], lt = 0.f;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            wt[s] = a2_ex2(xbuf[s * XW + HD * 32 + lane] - mt);
            lt += xbuf[s * XW + (HD + 1) * 32 + lane] * wt[s];
          }
          const float inv = 1.f / lt;
          float f[HD / 4];
#pragma unroll
          for (int c = 0; c < HD / 4; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) acc += xbuf[s * XW + (q * (HD / 4) + c) * 32 + lane] * wt[s];
            f[c] = acc * inv;
          }
          if (p_off >= 0) {
            __half* dst = p.out + p_off;
            if constexpr (HD == 32) {
              uint4 wd;
              wd.x = a2_pack(f[0], f[1]); wd.y = a2_pack(f[2], f[3]); wd.z = a2_pack(f[4], f[5]); wd.w = a2_pack(f[6], f[7]);
              *reinterpret_cast<uint4*>(dst) = wd;
            } else {
              uint2 wd;
              wd.x = a2_pack(f[0], f[1]); wd.y = a2_pack(f[2], f[3]);
              *reinterpret_cast<uint2*>(dst) = wd;
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");   // xbuf is reused by the next ragged read-out
        }
      }
      A2_MARK(w == 0 && q == 0, p_i, 9);
    };
    // One flat loop over the jobs of this warpgroup plus a final flush iteration, so that the (large) read-out code
    // has a single call site per head state.
    // Ping-pong of the two warpgroups (FA3 style, named barriers 3 + w): a warpgroup enters the MUFU-heavy part of
    // a job only after the other one has left its own, so the exponentials of one overlap the loads, hand-offs and
    // read-outs of the other instead of both competing for the MUFU pipe and then idling together.  Both warpgroups
    // run the same number of jobs, and every job passes through the barrier pair exactly once.
    if (w == 1) asm volatile("bar.arrive 3, 256;" ::: "memory");   // warpgroup 0 goes first
    A2Job<HPG, 2> it;
    it.init(p, w);
    int cur_item = -1;
    long long item_off = 0;                                     // element offset of (row 0 of the sequence, head 0 of the group)
    const long long row_pitch = (long long)p.map.t_stride * p.N;
    for (uint32_t i = 0;; ++i) {
      const bool have = it.valid;
      float mx = -1e30f, sum = 0.f;
      bool rag = false;
      long long off = -1;
      if (have) {
        if (it.item != cur_item) {                              // (integer divisions: once per item, not per job)
          cur_item = it.item;
          const int g = it.item / p.groups, grp = it.item - g * p.groups;
          item_off = p.map.row(g, 0) * p.N + grp * 64;
        }
        rag = p.rag && it.m == p.mtiles - 1;
        const bool warp_live = rag || it.m * 128 + q * 32 < p.len;   // any valid query row in this warp
        const int qi = it.m * 128 + (rag ? lane : r_tile);
        if (qi < p.len) off = item_off + qi * row_pitch + (rag ? q * (HD / 4) : 0) + it.hh * HD;
        const int nv = min(p.NB, p.len - it.j * p.NB);           // valid keys of this block (>= 1)
        const uint32_t r = i & 1;
        const uint32_t t_s = t_lane + (2 * w + r) * A2_SLOT;
        mbar_wait(s_full + 8 * (2 * w + r), (i >> 1) & 1);
        tc_fence_after();
        A2_MARK(w == 0 && q == 0, i, 0);
        asm volatile("bar.sync %0, 256;" ::"r"(3 + w) : "memory");
        if (!rag && warp_live) {
          // ---- full tile: this thread owns query row q * 32 + lane.  The whole row (<= 96 columns) is loaded into
          // registers ONCE: one TMEM round trip per job instead of one per 32 columns and pass (those round trips,
          // ~200 cycles each, were most of a job's time).  Columns beyond the sequence become -inf (P = 0).
          uint32_t v[96];
          tmem_ld_32x32b_x32(t_s, v);
          if (nv > 32) tmem_ld_32x32b_x32(t_s + 32, v + 32);
          if (nv > 64) tmem_ld_32x32b_x32(t_s + 64, v + 64);
          tmem_ld_wait();
          const int hl = (nv - 1) >> 4, nvl = nv - 16 * hl;     // last 16-column group with valid keys, their count
          if (nvl < 16) {
#pragma unroll
            for (int h = 0; h < 6; ++h)
              if (h == hl) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (e >= nvl) v[16 * h + e] = 0xff800000u;   // -inf
              }
          }
#pragma unroll
          for (int h = 0; h < 6; ++h)
            if (h <= hl) {
#pragma unroll
              for (int e = 0; e < 16; e += 2) mx = a2_max3(mx, __uint_as_float(v[16 * h + e]), __uint_as_float(v[16 * h + e + 1]));
            }
          A2_MARK(w == 0 && q == 0, i, 1);
          // P = exp2(s - max) as packed fp16 over S in place (every S column is in registers already), fp32 row sum
          float sum1 = 0.f;
#pragma unroll
          for (int h = 0; h < 6; ++h)
            if (h <= hl) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float e0 = a2_ex2(__uint_as_float(v[16 * h + 2 * e]) - mx);
                const float e1 = a2_ex2(__uint_as_float(v[16 * h + 2 * e + 1]) - mx);
                sum += e0; sum1 += e1;
                pk[e] = a2_pack(e0, e1);
              }
              tmem_st_32x32b_x8(t_s + 8 * h, pk);
            }
          sum += sum1;
        } else if (rag) {
          // ---- ragged tile: row = lane (replicated in every quadrant), this warp owns columns
          // [8 u_first, 8 (u_first + u_cnt)) of the block; the rest of its P row is zero
          const int c0 = 8 * u_first;
          const int n_ok = max(0, min(nv - c0, 8 * u_cnt));   // valid columns of the strip
          uint32_t v[24];
#pragma unroll
          for (int k = 0; k < 3; ++k)
            if (k < u_cnt) tmem_ld_32x32b_x8(t_s + c0 + 8 * k, v + 8 * k);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 24; ++e) mx = fmaxf(mx, e < n_ok ? __uint_as_float(v[e]) : -1e30f);
          {
            uint32_t z[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) z[e] = 0u;
#pragma unroll
            for (int k = 0; k < 3; ++k) tmem_st_32x32b_x16(t_s + 16 * k, z);   // packed columns [0, 48)
            tmem_st_wait();
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (k < u_cnt && 8 * k < n_ok) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float e0 = a2_ex2(__uint_as_float(v[8 * k + 2 * e]) - mx);
                float e1 = a2_ex2(__uint_as_float(v[8 * k + 2 * e + 1]) - mx);
                e0 = (8 * k + 2 * e < n_ok) ? e0 : 0.f;
                e1 = (8 * k + 2 * e + 1 < n_ok) ? e1 : 0.f;
                sum += e0 + e1;
                pk[e] = a2_pack(e0, e1);
              }
              tmem_st_32x32b_x4(t_s + ((c0 + 8 * k) >> 1), pk);
            }
          }
        }
        A2_MARK(w == 0 && q == 0, i, 2);
        asm volatile("bar.arrive %0, 256;" ::"r"(4 - w) : "memory");
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready + 8 * (2 * w + r));
        A2_MARK(w == 0 && q == 0, i, 3);
      }
      // deferred read-out of the previous job of this warpgroup (its P V MMAs ran during this softmax)
      if (pend) {
        if constexpr (HPW == 1) read_out(Mr[0], Lr[0], Or[0]);
        else if (p_i & 1) read_out(Mr[1], Lr[1], Or[1]);         // (HPW == 2: job i of the warpgroup is head i & 1)
        else read_out(Mr[0], Lr[0], Or[0]);
      }
      if (!have) break;
      pend = true; p_i = i; p_mx = mx; p_sum = sum; p_rag = rag; p_first = it.j == 0; p_last = it.j == p.nblk - 1;
      p_off = off;
      it.next(p);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

template <int HD>
static int attn2_launch(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, cudaStream_t st) {
  Attn2Args a;
  a.mode = mode; a.len = map.len; a.N = N; a.groups = N / 64; a.map = map; a.out = out;
  a.trace = g_lstm_trace;
  a.nblk = (a.len + 95) / 96;
  a.NB = (((a.len + a.nblk - 1) / a.nblk) + 15) / 16 * 16;
  a.mtiles = (a.len + 127) / 128;
  a.num_items = map.G * a.groups;
  a.rag = a.len - (a.mtiles - 1) * 128 <= 32;
  const size_t smem = 2 * A2_QBYTES + (size_t)A2_NSTG * 2 * a.NB * 128 + 2 * 4 * (HD + 2) * 32 * 4 + 256;
  CUtensorMap tmQ, tmQ32, tmKV;
  const long long tok = (long long)B * S * C;
  if (mode == 0) {
    const uint64_t dims[2] = {(uint64_t)3 * N, (uint64_t)tok};
    const uint64_t str[1] = {(uint64_t)3 * N * 2};
    const uint32_t boxq[2] = {64, 128}, boxq32[2] = {64, 32}, boxkv[2] = {64, (uint32_t)a.NB};
    if (make_tmap_f16(&tmQ, qkv, 2, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmQ32, qkv, 2, dims, str, boxq32)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 2, dims, str, boxkv)) return -1;
  } else {
    const uint64_t dims[4] = {(uint64_t)3 * N, (uint64_t)C, (uint64_t)S, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)3 * N * 2, (uint64_t)C * 3 * N * 2, (uint64_t)S * C * 3 * N * 2};
    const uint32_t boxq[4] = {64, 1, 128, 1}, boxq32[4] = {64, 1, 32, 1}, boxkv[4] = {64, 1, (uint32_t)a.NB, 1};
    if (make_tmap_f16(&tmQ, qkv, 4, dims, str, boxq)) return -1;
    if (make_tmap_f16(&tmQ32, qkv, 4, dims, str, boxq32)) return -1;
    if (make_tmap_f16(&tmKV, qkv, 4, dims, str, boxkv)) return -1;
  }
  static PerDeviceOnce configured;
  if (configured.first()) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_attn2<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_attn2<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  }
  const int grid = a.num_items < grid_cap() ? a.num_items : grid_cap();
  if (a.trace) k_tc_attn2<HD, true><<<grid, A2_THREADS, smem, st>>>(tmQ, tmQ32, tmKV, a);
  else k_tc_attn2<HD, false><<<grid, A2_THREADS, smem, st>>>(tmQ, tmQ32, tmKV, a);
  VATSS_LAUNCH_OK();
  return 0;
}

// N % 64 == 0 and head dim 16 / 32 (checked by the caller)
int launch_attention_v2(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                        cudaStream_t st) {
  if (N / heads == 32) return attn2_launch<32>(qkv, out, map, mode, B, S, C, N, st);
  return attn2_launch<16>(qkv, out, map, mode, B, S, C, N, st);
}

}  // namespace vatss

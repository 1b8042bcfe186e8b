// TENSOR engine, attention core v3 (round 2): P in TMEM, O accumulated in TMEM, four lean softmax warpgroups.
//
//   out16[row, h*hd:(h+1)*hd] = softmax(q k^T) v     (nn.MultiheadAttention core, src/model/dptn.py:16-21,46)
//
// q arrives pre-scaled by log2(e)/sqrt(hd) (folded into the in-projection at weight-pack time): P = exp2(s - ref).
//
// Lineage (profiles/r02_attn*): v1 (tc_attention.cu, P through shared memory) and a first TMEM design with two fat
// softmax warpgroups that also read out O ("v2", removed) both sat at 0.22 - 0.32 of the MUFU floor.  Per
// (128-query x <= 96-key x head) job a softmax thread has ~640 cycles of MUFU work but ~2500 cycles of exposed latency
// (TMEM round trips ~200 cycles each, mbarrier hand-offs, O read-out) - with two warps per scheduler neither the MUFU
// pipe (40 %) nor the issue slots (45 %) are busy.  The cure is occupancy, and registers are what limits it:
//   * softmax threads are LEAN (<= 80 registers): one 32-column TMEM chunk at a time, loaded ONCE (no separate
//     maximum pass), no O state.  O is not read out per job: the P V MMAs of all kv blocks of a (query tile, head)
//     accumulate in TMEM.  The reference is the maximum of the first 32-column chunk of the row; a later chunk only
//     forces a rescale (of O - after the previous P V has completed -, of the row sum and of the P chunks already
//     written, all by the softmax thread that owns the row) when it exceeds the reference by more than 2^8: exact
//     softmax, the rescale is a rare slow path (FA4's conditional rescaling, at chunk granularity).
//   * FOUR softmax warpgroups = two independent pipelines per CTA (even / odd items of the CTA's list), each with its
//     own TMA ring and MMA issuer; inside a pipeline warpgroup w handles head w of the current head pair.
//   * one epilogue warpgroup reads O once per (query tile, head), normalises and stores.
// 24 warps: 0..15 softmax; 16..19 epilogue; 20, 21 TMA producers; 22, 23 MMA issuers (S(i + 1) right behind P V(i): the
// in-order tensor pipe protects the aliased S / P slot).  The issuers carry the highest warp ids: the scheduler favours
// the highest id among eligible warps, and next to four MUFU-bound softmax warps a low-id issuer starves (the
// P ready -> P V -> S -> S full round trip was 2300 cycles with the issuers as warps 2, 3).  TMEM (512 columns): S / P slot of softmax warpgroup sw at 96 sw,
// its O accumulator at 384 + 32 sw.  A ragged last query tile (<= 32 rows) is loaded into all four lane quadrants;
// quadrant q handles a strip of the kv block (zeros elsewhere in its P rows) and the epilogue merges the four partial
// results (each with its own reference and row sum) through shared memory.
// What still bounds it (tools/attn3_trace.py): S and P of a warpgroup share one TMEM slot, so S(i + 1) follows P V(i) and
// every job ends with a ~2000-cycle round trip through the issuer (tcgen05 instructions cost ~100 cycles each to issue).
// Tried and rejected: kv blocks of 64 keys with P in separate TMEM columns so that S(i + 1) is issued while job i is still
// being exponentiated ("v4": 0.74 / 0.99 ms against 0.62 / 0.81 - more, smaller jobs load the issuer further); ping-pong
// barriers between the two warpgroups of a pipeline (0.69 / 0.89); a second S slot per warpgroup needs 2 x 96 columns per
// warpgroup and leaves room for two warpgroups only (0.81 / 1.06).
#include "common.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

#ifndef A3_BACKOFF_NS
#define A3_BACKOFF_NS 40
#endif
#define a3_wait mbar_wait_backoff<A3_BACKOFF_NS>

struct Attn3Args {
  int mode;       // 0 intra, 1 inter
  int len;        // tokens per sequence
  int N;          // features (row of qkv is 3N halfs)
  int groups;     // 64-feature head groups per sequence
  int nblk;       // kv blocks per sequence
  int NB;         // keys per kv block (multiple of 16, <= 96)
  int mtiles;     // 128-query tiles per sequence
  int num_items;  // sequences * groups
  int rag;        // last query tile has <= 32 rows: replicated-quadrant strip mode
  int nstg;       // K / V ring stages per pipeline (2 or 3)
  SeqMap map;
  __half* out;    // (tokens, N)
  long long* trace;   // optional clock64 timeline of CTA 0 (debug), NULL in production
};

// debug timeline: jobs [A3_T0, A3_T0 + 8) of softmax warpgroup 0 in CTA 0, 16 slots per job
constexpr uint32_t A3_T0 = 24;
#define A3_MARK(cond, i, k)                                                                    \
  do {                                                                                         \
    if (TRACE && blockIdx.x == 0 && (cond) && (i) >= A3_T0 && (i) < A3_T0 + 8 && lane == 0)    \
      p.trace[((i) - A3_T0) * 16 + (k)] = clock64();                                           \
  } while (0)

constexpr int A3_THREADS = 768;
constexpr uint32_t A3_SLOT = 96;
constexpr uint32_t A3_OCOL = 384;
constexpr uint32_t A3_QBYTES = 16384;
constexpr float A3_RESCALE = 8.f;       // log2 headroom before O is rescaled (P <= 2^8 is exact enough in fp16)

__device__ __forceinline__ uint64_t a3_desc_mnmajor(uint32_t smem_addr) {   // V: kv rows of 128 B, features contiguous
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float a3_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float a3_max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ uint32_t a3_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Job stream of one pipeline: (item, query tile m, head pair hp, kv block j, warpgroup w), w fastest.  STEP = 1 walks
// every job (issuer), STEP = 2 the jobs of one warpgroup.  qn / kvn count Q tiles and kv stages (ring positions).
template <int HPG, int STEP>
struct A3Job {
  int item, m, hp, j, w, w0, stride;
  uint32_t qn, kst, kph;      // Q tiles consumed; K / V ring stage and phase of the current kv block (no divisions)
  bool valid;
  __device__ __forceinline__ void init(const Attn3Args& p, int pl, int first_w) {
    stride = 2 * gridDim.x;
    item = blockIdx.x + pl * gridDim.x;
    m = hp = j = 0; w = w0 = first_w; qn = kst = kph = 0; valid = item < p.num_items;
  }
  __device__ __forceinline__ void next(const Attn3Args& p) {
    w += STEP;
    if (w < 2) return;
    w = w0;
    if (++kst == (uint32_t)p.nstg) { kst = 0; kph ^= 1u; }
    if (++j < p.nblk) return;
    j = 0;
    if (++hp < HPG / 2) return;
    hp = 0; ++qn;
    if (++m < p.mtiles) return;
    m = 0; item += stride;
    valid = item < p.num_items;
  }
};

template <int HD, bool TRACE>
__global__ void __launch_bounds__(A3_THREADS, 1)
k_tc_attn3(const __grid_constant__ CUtensorMap tmapQ, const __grid_constant__ CUtensorMap tmapQ32,
           const __grid_constant__ CUtensorMap tmapKV, Attn3Args p) {
  constexpr int HPG = 64 / HD;       // heads per 64-feature group (2 or 4)
  constexpr int XW = (HD + 2) * 32;  // floats of one quadrant's partial in the ragged-tile merge buffer
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t KVB = (uint32_t)p.NB * 128u;                 // bytes of one K (or V) block
  const uint32_t PLB = 2 * A3_QBYTES + (uint32_t)p.nstg * 2 * KVB;   // bytes of one pipeline's tiles
  // per pipeline: Q [2][128 x 128 B], then the K / V ring [nstg][K block | V block]
  const uint32_t stat_off = 2 * PLB;                           // [4 sw][2 group parities][128 rows] (reference, row sum)
  const uint32_t x_off = stat_off + 4 * 2 * 128 * 8;           // ragged-tile merge buffer [4 quadrants][HD + 2][32] floats
  const uint32_t bars = base + x_off + 4 * XW * 4;
  // barriers: per pipeline q_full[2] q_free[2] kv_full[3] kv_free[3] (80 B), then per softmax warpgroup
  // s_full, p_ready, pv_done, g_full, o_free, st_ready[2 group parities] (56 B)
  constexpr uint32_t BSW = 56;
  const uint32_t bar_pl = bars, bar_sw = bars + 2 * 80;
  const uint32_t tmem_slot = bar_sw + 4 * BSW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int pl = 0; pl < 2; ++pl) {
      const uint32_t b = bar_pl + 80 * pl;
      for (int i = 0; i < 10; ++i) mbar_init(b + 8 * i, 1);
    }
    for (int sw = 0; sw < 4; ++sw) {
      const uint32_t b = bar_sw + BSW * sw;
      mbar_init(b, 1); mbar_init(b + 8, 4); mbar_init(b + 16, 1); mbar_init(b + 24, 1); mbar_init(b + 32, 4);
      mbar_init(b + 40, 4); mbar_init(b + 48, 4);
    }
    fence_mbar_init();
    prefetch_tmap(&tmapQ);
    prefetch_tmap(&tmapQ32);
    prefetch_tmap(&tmapKV);
  }
  if (warp == 22) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

  if (warp == 20 || warp == 21) {
    // ---------------------------------------------------------------- TMA producer of pipeline (warp - 20)
    if (lane == 0) {
      const int pl = warp - 20;
      const uint32_t sQ = base + pl * PLB, sKV = sQ + 2 * A3_QBYTES;
      const uint32_t q_full = bar_pl + 80 * pl, q_free = q_full + 16, kv_full = q_full + 32, kv_free = q_full + 56;
      uint32_t qn = 0, st = 0, ph = 0;
      for (int item = blockIdx.x + pl * gridDim.x; item < p.num_items; item += 2 * gridDim.x) {
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
          a3_wait(q_free + 8 * qb, ((qn >> 1) & 1) ^ 1);
          mbar_expect_tx(q_full + 8 * qb, A3_QBYTES);
          if (p.rag && m == p.mtiles - 1) {
            for (int k = 0; k < 4; ++k) load_rows(&tmapQ32, sQ + qb * A3_QBYTES + k * 4096, q_full + 8 * qb, colq, m * 128);
          } else {
            load_rows(&tmapQ, sQ + qb * A3_QBYTES, q_full + 8 * qb, colq, m * 128);
          }
          for (int hp = 0; hp < HPG / 2; ++hp)
            for (int j = 0; j < p.nblk; ++j) {
              a3_wait(kv_free + 8 * st, ph ^ 1);
              mbar_expect_tx(kv_full + 8 * st, 2 * KVB);
              load_rows(&tmapKV, sKV + st * 2 * KVB, kv_full + 8 * st, colk, j * p.NB);
              load_rows(&tmapKV, sKV + st * 2 * KVB + KVB, kv_full + 8 * st, colv, j * p.NB);
              if (++st == (uint32_t)p.nstg) { st = 0; ph ^= 1u; }
            }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 22) {
    // ---------------------------------------------------------------- MMA issuer of pipeline (warp - 22)
    // Warp-uniform control flow, one elected lane issues (umma_*_warp).  Jobs alternate between the two warpgroups
    // of the pipeline; the S of a warpgroup's next job is issued right behind the P V of its current one.
    const int pl = __shfl_sync(0xffffffffu, warp, 0) - 22;
    // The CTA owns all 512 TMEM columns, so the allocation starts at column 0, lane 0 (checked): with a literal base
    // every tcgen05 operand of the issuer is computed from warp-uniform values (kernel parameters, blockIdx, loop
    // counters) and ptxas can keep it in uniform registers instead of moving it over with R2UR before each MMA.
    if (tmem != 0) __trap();
    const uint32_t tmem_u = 0;
    const uint32_t sQ = base + pl * PLB, sKV = sQ + 2 * A3_QBYTES;
    const uint32_t q_full = bar_pl + 80 * pl, q_free = q_full + 16, kv_full = q_full + 32, kv_free = q_full + 56;
    const uint32_t idesc_s = idesc_f16(128, p.NB, 0);
    const uint32_t idesc_o = idesc_f16(128, HD, 0) | (1u << 16);      // B (= V) is MN-major
    A3Job<HPG, 1> si, pi;
    si.init(p, pl, 0); pi.init(p, pl, 0);
    uint32_t ip0 = 0, ip1 = 0;      // P V jobs issued per warpgroup
    uint32_t kg0 = 0, kg1 = 0;      // (tile, head) groups started per warpgroup
    auto issue_s = [&]() {
      const uint32_t w = si.w, sw = 2 * pl + w, bsw = bar_sw + BSW * sw;
      const int head = 2 * si.hp + w;
      if (si.hp == 0 && si.j == 0 && w == 0) mbar_wait_warp(q_full + 8 * (si.qn & 1), (si.qn >> 1) & 1);
      if (w == 0) mbar_wait_warp(kv_full + 8 * si.kst, si.kph);
      tc_fence_after();
      const uint64_t qd = smem_desc_sw128_kmajor(sQ + (si.qn & 1) * A3_QBYTES) + ((uint32_t)(head * HD * 2) >> 4);
      const uint64_t kd = smem_desc_sw128_kmajor(sKV + si.kst * 2 * KVB) + ((uint32_t)(head * HD * 2) >> 4);
#pragma unroll
      for (int k16 = 0; k16 < HD / 16; ++k16)
        umma_f16_warp<1>(tmem_u + sw * A3_SLOT, qd + 2 * k16, kd + 2 * k16, idesc_s, k16 > 0 ? 1u : 0u);
      umma_commit_warp(bsw);                                                            // s_full
      if (si.hp == HPG / 2 - 1 && si.j == p.nblk - 1 && w == 1) umma_commit_warp(q_free + 8 * (si.qn & 1));
      si.next(p);
    };
    if (si.valid) issue_s();
    if (si.valid) issue_s();
    while (pi.valid) {
      const uint32_t w = pi.w, sw = 2 * pl + w, bsw = bar_sw + BSW * sw;
      const uint32_t i = w ? ip1 : ip0, k = w ? kg1 : kg0;
      const int head = 2 * pi.hp + w;
      A3_MARK(sw == 0, i, 3);
      mbar_wait_warp(bsw + 8, i & 1);                                                   // p_ready
      A3_MARK(sw == 0, i, 4);
      if (pi.j == 0) mbar_wait_warp(bsw + 32, (k & 1) ^ 1);                             // o_free: previous group read out
      tc_fence_after();
      A3_MARK(sw == 0, i, 5);
      const uint32_t vbase = sKV + pi.kst * 2 * KVB + KVB + (uint32_t)(head * HD * 2);
      const uint64_t vd = a3_desc_mnmajor(vbase);
      const int nv = min(p.NB, p.len - pi.j * p.NB);
      const int nk = (nv + 15) >> 4;                            // P columns beyond the sequence are never multiplied
      for (int k16 = 0; k16 < nk; ++k16)
        umma_f16_ts_warp(tmem_u + A3_OCOL + sw * 32, tmem_u + sw * A3_SLOT + 8 * k16, vd + (uint32_t)((k16 * 16 * 128) >> 4),
                         idesc_o, (pi.j > 0 || k16 > 0) ? 1u : 0u);
      umma_commit_warp(bsw + 16);                                                       // pv_done
      if (pi.j == p.nblk - 1) {
        umma_commit_warp(bsw + 24);                                                     // g_full
        if (w) ++kg1; else ++kg0;
      }
      if (w == 1) umma_commit_warp(kv_free + 8 * pi.kst);                               // last MMA on this K / V stage
      A3_MARK(sw == 0, i, 6);
      if (w) ++ip1; else ++ip0;
      pi.next(p);
      if (si.valid) issue_s();
      A3_MARK(sw == 0, i, 7);
    }
  } else if (warp >= 16) {
    // ---------------------------------------------------------------- epilogue warpgroup
    // Fixed round-robin over the four softmax warpgroups (their group counts are known), one (tile, head) at a time.
    const int q = warp & 3;
    const int r_tile = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    float* xbuf = reinterpret_cast<float*>(smem + x_off);
    const float2* stat = reinterpret_cast<const float2*>(smem + stat_off);
    const long long row_pitch = (long long)p.map.t_stride * p.N;
    // Order: (item slot t of the CTA, (tile, head pair) group, pipeline, warpgroup) - the two pipelines alternate group by
    // group.  Runtime loops over pipeline and warpgroup keep a single copy of the read-out code.
    const int groups_per_item = p.mtiles * (HPG / 2);
#pragma unroll 1
    for (int t = 0;; ++t) {
      const int item0 = blockIdx.x + 2 * t * gridDim.x;
      if (item0 >= p.num_items) break;
#pragma unroll 1
      for (int gi = 0; gi < groups_per_item; ++gi) {
        const int m = gi / (HPG / 2), hp = gi - m * (HPG / 2);
        const bool rag = p.rag && m == p.mtiles - 1;
        const int qi = m * 128 + (rag ? lane : r_tile);
        const uint32_t k = (uint32_t)(t * groups_per_item + gi);        // group index inside each warpgroup's stream
#pragma unroll 1
        for (int pl = 0; pl < 2; ++pl) {
          const int item = item0 + pl * gridDim.x;
          if (item >= p.num_items) continue;
          const int g = item / p.groups, grp = item - g * p.groups;
          const long long off0 = p.map.row(g, 0) * p.N + grp * 64 + qi * row_pitch;
#pragma unroll 1
          for (int w = 0; w < 2; ++w) {
            const int head = 2 * hp + w;
            const uint32_t sw = 2 * pl + w, bsw = bar_sw + BSW * sw;
            a3_wait(bsw + 24, k & 1);                                                   // g_full
            tc_fence_after();
            uint32_t o[HD];
            if constexpr (HD == 32) tmem_ld_32x32b_x32(t_lane + A3_OCOL + sw * 32, o);
            else tmem_ld_32x32b_x16(t_lane + A3_OCOL + sw * 32, o);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bsw + 32);                                         // o_free
            a3_wait(bsw + 40 + 8 * (k & 1), (k >> 1) & 1);                              // st_ready: (reference, sum) written
            const float2 ml = stat[(sw * 2 + (k & 1)) * 128 + r_tile];
            if (!rag) {
              if (qi < p.len) {
                const float inv = 1.f / ml.y;
                uint4* dst = reinterpret_cast<uint4*>(p.out + off0 + head * HD);
#pragma unroll
                for (int c = 0; c < HD / 8; ++c) {
                  uint4 wd;
                  wd.x = a3_pack(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                  wd.y = a3_pack(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                  wd.z = a3_pack(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                  wd.w = a3_pack(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                  dst[c] = wd;
                }
              }
            } else {
              // four strip partials of row `lane` live in the four quadrants: merge through shared memory, then
              // quadrant q finishes features [q HD/4, (q+1) HD/4) of the row
#pragma unroll
              for (int c = 0; c < HD; ++c) xbuf[q * XW + c * 32 + lane] = __uint_as_float(o[c]);
              xbuf[q * XW + HD * 32 + lane] = ml.x;
              xbuf[q * XW + (HD + 1) * 32 + lane] = ml.y;
              asm volatile("bar.sync 1, 128;" ::: "memory");
              float mt = -1e30f;
#pragma unroll
              for (int s = 0; s < 4; ++s) mt = fmaxf(mt, xbuf[s * XW + HD * 32 + lane]);
              float wt[4], lt = 0.f;
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                wt[s] = a3_ex2(xbuf[s * XW + HD * 32 + lane] - mt);
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
              if (qi < p.len) {
                __half* dst = p.out + off0 + head * HD + q * (HD / 4);
                if constexpr (HD == 32) {
                  uint4 wd;
                  wd.x = a3_pack(f[0], f[1]); wd.y = a3_pack(f[2], f[3]); wd.z = a3_pack(f[4], f[5]); wd.w = a3_pack(f[6], f[7]);
                  *reinterpret_cast<uint4*>(dst) = wd;
                } else {
                  uint2 wd;
                  wd.x = a3_pack(f[0], f[1]); wd.y = a3_pack(f[2], f[3]);
                  *reinterpret_cast<uint2*>(dst) = wd;
                }
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");   // xbuf is reused by the next ragged read-out
            }
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    const int sw = warp >> 2;
    const int pl = sw >> 1, w = sw & 1;
    const int q = warp & 3;                                     // TMEM lane quadrant
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t t_s = t_lane + sw * A3_SLOT, t_o = t_lane + A3_OCOL + sw * 32;
    const uint32_t bsw = bar_sw + BSW * sw;
    float2* stat = reinterpret_cast<float2*>(smem + stat_off) + sw * 2 * 128 + q * 32 + lane;
    // ragged strips: the kv block is NB / 8 eight-column units, shared out over the four quadrants (<= 3 each)
    const int nu8 = p.NB >> 3;
    const int u_cnt = nu8 / 4 + (q < (nu8 & 3) ? 1 : 0);
    const int u_first = q * (nu8 / 4) + min(q, nu8 & 3);
    A3Job<HPG, 2> it;
    it.init(p, pl, w);
    float mref = 0.f, l = 0.f;                                  // reference (log2 domain) and row sum of the open group
    bool have_ref = false;
    uint32_t k = 0;                                             // (tile, head) group index of this warpgroup
#pragma unroll 1
    for (uint32_t i = 0; it.valid; ++i) {
      const bool rag = p.rag && it.m == p.mtiles - 1;
      const bool warp_live = rag || it.m * 128 + q * 32 < p.len;   // any valid query row in this warp
      const int nv = min(p.NB, p.len - it.j * p.NB);              // valid keys of this block (>= 1)
      // this thread's columns: the whole block, or its strip of the ragged tile
      const int c0 = rag ? 8 * u_first : 0;
      const int ncol = !warp_live ? 0 : (rag ? max(0, min(nv - c0, 8 * u_cnt)) : nv);
      A3_MARK(sw == 0 && q == 0, i, 8);
      a3_wait(bsw, i & 1);                                                            // s_full
      tc_fence_after();
      A3_MARK(sw == 0 && q == 0, i, 0);
      uint32_t v[32];
      const int nfull = ncol >> 5, rem = ncol & 31;
      // ---- ragged tile: the P row is zero outside this thread's strip.  The strip (<= 24 columns, one chunk) is
      // loaded first: the zero fill overwrites it in TMEM.
      if (rag) {
        if (rem) tmem_ld_32x32b_x32(t_s + c0, v);
        tmem_ld_wait();
        uint32_t z[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = 0u;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) tmem_st_32x32b_x16(t_s + 16 * kk, z);           // packed columns [0, 48)
        tmem_st_wait();
      }
      if (it.j == 0) { l = 0.f; have_ref = false; }
      // The reference of a (tile, head) is the maximum of the FIRST 32-column chunk this thread sees; every later chunk
      // (same or later kv block) is checked against it and only forces a rescale - of O, of the row sum and of the
      // P chunks of this job already written - when it exceeds the reference by more than 2^8.  One TMEM load per
      // chunk: no separate maximum pass.
      auto check_reference = [&](float cmax, int u_done) {
        if (!have_ref) { mref = cmax; have_ref = true; return; }
        const bool need = cmax > mref + A3_RESCALE;
        if (__any_sync(0xffffffffu, need)) {
          const float mnew = need ? cmax : mref;
          const float f = a3_ex2(mref - mnew);                  // 1 for the rows that keep their reference
          if (it.j > 0) {                                       // O holds the previous kv blocks of this group
            a3_wait(bsw + 16, (i - 1) & 1);                     // pv_done: the previous block's P V has completed
            tc_fence_after();
#pragma unroll 1
            for (int c0o = 0; c0o < HD; c0o += 16) {            // 16 columns at a time: v[] stays live around this
              uint32_t o[16];
              tmem_ld_32x32b_x16(t_o + c0o, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
              tmem_st_32x32b_x16(t_o + c0o, o);
            }
          }
          const __half2 f2 = __float2half2_rn(f);
#pragma unroll 1
          for (int uu = 0; uu < u_done; ++uu) {                 // P chunks of this job written with the old reference
            uint32_t pp[16];
            tmem_ld_32x32b_x16(t_s + ((c0 + 32 * uu) >> 1), pp);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&pp[e]), f2);
              pp[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            tmem_st_32x32b_x16(t_s + ((c0 + 32 * uu) >> 1), pp);
          }
          l *= f;
          mref = mnew;
        }
      };
      // ---- P = exp2(s - reference) as packed fp16 over S in place, fp32 row sum.  The P of columns [c, c + 32)
      // lands in packed columns [c / 2, c / 2 + 16): S columns that have been consumed already.  One loop over the
      // chunks (the last one may be partial) with a single call site of the reference check: code size matters here.
      const int nchunk = nfull + (rem ? 1 : 0);
#pragma unroll 1
      for (int u = 0; u < nchunk; ++u) {
        const int cnt = u < nfull ? 32 : rem;
        if (!rag) {
          tmem_ld_32x32b_x32(t_s + c0 + 32 * u, v);
          tmem_ld_wait();
        }
        float cmax = -1e30f;
        if (cnt == 32) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) cmax = a3_max3(cmax, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) cmax = fmaxf(cmax, e < cnt ? __uint_as_float(v[e]) : -1e30f);
        }
        check_reference(cmax, u);
        float sum = 0.f, sum1 = 0.f;
        if (cnt == 32) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float e0 = a3_ex2(__uint_as_float(v[2 * e]) - mref);
            const float e1 = a3_ex2(__uint_as_float(v[2 * e + 1]) - mref);
            sum += e0; sum1 += e1;
            pk[e] = a3_pack(e0, e1);
          }
          tmem_st_32x32b_x16(t_s + ((c0 + 32 * u) >> 1), pk);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (cnt > 16 * h) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float e0 = a3_ex2(__uint_as_float(v[16 * h + 2 * e]) - mref);
                float e1 = a3_ex2(__uint_as_float(v[16 * h + 2 * e + 1]) - mref);
                e0 = (16 * h + 2 * e < cnt) ? e0 : 0.f;
                e1 = (16 * h + 2 * e + 1 < cnt) ? e1 : 0.f;
                sum += e0; sum1 += e1;
                pk[e] = a3_pack(e0, e1);
              }
              tmem_st_32x32b_x8(t_s + ((c0 + 32 * u + 16 * h) >> 1), pk);
            }
          }
        }
        l += sum + sum1;
      }
      A3_MARK(sw == 0 && q == 0, i, 1);
      const bool last_blk = it.j == p.nblk - 1;
      if (last_blk) stat[(k & 1) * 128] = make_float2(have_ref ? mref : -1e30f, l);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (last_blk) mbar_arrive(bsw + 40 + 8 * (k & 1));                              // st_ready
        mbar_arrive(bsw + 8);                                                           // p_ready
      }
      A3_MARK(sw == 0 && q == 0, i, 2);
      if (last_blk) ++k;
      it.next(p);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 22) tmem_dealloc<1>(tmem, 512);
}

template <int HD>
static int attn3_launch(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, cudaStream_t st) {
  Attn3Args a;
  a.mode = mode; a.len = map.len; a.N = N; a.groups = N / 64; a.map = map; a.out = out;
  a.nblk = (a.len + 95) / 96;
  a.NB = (((a.len + a.nblk - 1) / a.nblk) + 15) / 16 * 16;
  {
    // The softmax walks a kv block in 32-column chunks: blocks of 96 keys (three full chunks) beat the even split when
    // they save a chunk (150 keys: 96 + 54 = 5 chunks against 80 + 70 = 6; measured 0.560 against 0.572 ms per launch).
    auto chunks = [&](int nb) {
      int c = 0;
      for (int k0 = 0; k0 < a.len; k0 += nb) c += ((a.len - k0 < nb ? a.len - k0 : nb) + 31) / 32;
      return c;
    };
    if (a.len > 96 && chunks(96) < chunks(a.NB)) a.NB = 96;
  }
  a.mtiles = (a.len + 127) / 128;
  a.num_items = map.G * a.groups;
  a.rag = a.len - (a.mtiles - 1) * 128 <= 32;
  a.trace = g_lstm_trace;
  const size_t fixed = 2 * 2 * A3_QBYTES + 4 * 2 * 128 * 8 + 4 * (HD + 2) * 32 * 4 + 1024;
  a.nstg = 3;
  size_t smem = fixed + 2 * (size_t)a.nstg * 2 * a.NB * 128;
  if (smem > 227 * 1024) {
    a.nstg = 2;
    smem = fixed + 2 * (size_t)a.nstg * 2 * a.NB * 128;
  }
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
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_attn3<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_attn3<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  }
  const int items2 = (a.num_items + 1) / 2;      // two pipelines per CTA
  const int grid = items2 < grid_cap() ? items2 : grid_cap();
  if (a.trace) k_tc_attn3<HD, true><<<grid, A3_THREADS, smem, st>>>(tmQ, tmQ32, tmKV, a);
  else k_tc_attn3<HD, false><<<grid, A3_THREADS, smem, st>>>(tmQ, tmQ32, tmKV, a);
  VATSS_LAUNCH_OK();
  return 0;
}

// N % 64 == 0 and head dim 16 / 32 (checked by the caller)
int launch_attention_v3(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                        cudaStream_t st) {
  if (N / heads == 32) return attn3_launch<32>(qkv, out, map, mode, B, S, C, N, st);
  return attn3_launch<16>(qkv, out, map, mode, B, S, C, N, st);
}

}  // namespace vatss

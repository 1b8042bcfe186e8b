// TENSOR engine: persistent LSTM recurrence on a CTA pair (tcgen05 cta_group::2).
//
// nn.LSTM(N -> H=128), one layer, h0 = c0 = 0, gate order i,f,g,o (src/model/dptn.py:23-29,49;
// src/model/dprnn.py:18,60).  One cluster of two CTAs owns 2 x 128 sequences of one direction for all
// time steps.  Per step and per sequence tile the gate pre-activations
//     G[256 seq, 512] = [x_t | h_{t-1}] [256, N+128] x [W_ih | W_hh]^T
// are one M=256 tensor-core contraction (input and recurrent halves fused, so the 1.36 M x 1024
// pre-activation tensor never exists in HBM).  The fp16 weights of a direction (256 KB) are split over
// the two SMs' shared memory (each CTA supplies half of the B rows of every MMA, hardware shares them),
// the fp32 accumulators fill each CTA's TMEM (128 lanes x 512 columns), x_t tiles arrive by TMA
// (4-D tensor map over the token-major activation: works for intra- and inter-chunk sequences), h_t is
// written back to shared memory as the next step's A operand and to HBM as the layer output.
//
// The 512 gate columns are processed as 4 chunks of 128 columns = 32 hidden units x (i,f,g,o), so that
// the gate math of chunk c overlaps the MMAs of the other chunks and the x-part of step t+1.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM allocator,
// warps 2..9 gate math (thread = sequence row, 16 hidden units per chunk).
//
// Two kernels share this scheme: k_tc_lstm (M = 256 MMAs, one 128-row tile per CTA) and k_tc_lstm_pp (M = 128 MMAs,
// two interleaved 64-row half tiles per CTA, further below), which hides the recurrence bubble and is the default.
// Both exist in the plain fp16 form and in the hi/lo split PRECISE form used for DPRNN.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

#ifndef LSTM_PREFETCH
#define LSTM_PREFETCH 1
#endif
constexpr int LSTM_H = 128;
constexpr int LSTM_THREADS = 320;
constexpr int LSTM_CHUNKS = 4;          // 4 x (32 units x 4 gates) = 512 accumulator columns
constexpr int LSTM_UNITS_PER_CHUNK = 32;

struct TcLstmArgs {
  int mode;        // 0 intra (sequences = flattened (b,s), time = k), 1 inter (sequences = (b,k), time = s)
  int len;         // time steps
  int ndir;
  int G;           // intra: number of sequences
  int B, S, C;     // geometry of the (B,S,C,N) activation
  int Kc, Bc;      // inter: tile = Kc chunk positions x Bc utterances (Kc*Bc <= 128)
  int kblocks;     // inter: ceil(C / Kc)
  int num_tiles;   // sequence tiles of 128 rows (per direction)
  int act;         // 1: store relu(h) (DPTN feeds the LSTM output through ReLU only), 0: store h
  const float* bias;   // [ndir][512] b_ih + b_hh in accumulator-column order
  __half* out;         // [tokens, ndir*128]
  long long* trace;    // optional clock64 trace of CTA 0 (debug), NULL in production
};

// PRECISE (N = 64 only): the input contraction runs on fp16 hi/lo splits of both x and W_ih
//   x W^T ~= x_hi W_hi^T + x_lo W_hi^T + x_hi W_lo^T       (lo = value - half(value), |lo| <= 2^-11 |value|)
// because DPRNN's un-normalised residual stream makes fp16-rounded W_ih / x miss the 1e-3 tolerance.
// Shared-memory layout is that of a 128-feature input: x tile = [x_hi | x_lo], weight rows = [W_hi | W_lo | W_hh].
template <int NFEAT, bool PRECISE = false>
struct TcLstmSmem {
  static constexpr int KBX = PRECISE ? 2 : NFEAT / 64;   // x k-blocks
  static constexpr int KBT = KBX + 2;                // + 2 h k-blocks
  static constexpr int W_BLOCK = 64 * 128;           // 64 B-rows x 128 B
  static constexpr int W_BYTES = LSTM_CHUNKS * KBT * W_BLOCK;
  static constexpr int X_STAGE = KBX * 16384;
  static constexpr int H_BYTES = 2 * 16384;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + W_BYTES;
  static constexpr int OFF_H = OFF_X + 2 * X_STAGE;
  static constexpr int OFF_BIAS = OFF_H + H_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 512 * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(2 x): the packed weights and biases of the i, f, o gates carry the factor 0.5 (exact in fp16 / fp32, see
// k_pack_lstm), so the accumulator already holds half the pre-activation and the pre-scaling multiply is gone
__device__ __forceinline__ float sigmoid_fast(float xh) { return fmaf(0.5f, tanh_fast(xh), 0.5f); }
// ~1e-6-accurate variants (two MUFU ops each: ex2 + rcp) for the PRECISE kernel
__device__ __forceinline__ float sigmoid_acc(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_acc(float x) { return fmaf(2.0f, sigmoid_acc(2.0f * x), -1.0f); }
template <bool PRECISE> __device__ __forceinline__ float act_sigmoid(float x) { return PRECISE ? sigmoid_acc(x) : sigmoid_fast(x); }
template <bool PRECISE> __device__ __forceinline__ float act_tanh(float x) { return PRECISE ? tanh_acc(x) : tanh_fast(x); }

// One LSTM cell from the four gate pre-activations (bias added); returns h, updates c.
//   plain fp16 path: five tanh.approx (the i / f / o pre-activations arrive halved, see sigmoid_fast).
//   PRECISE path: ex2 + rcp sigmoids would cost 10 MUFU operations per cell.  The four gate functions share ONE
//   reciprocal instead (1 / (d_i d_f d_g d_o), each factor recovered with two multiplies), so a cell costs 4 ex2 + 1 rcp
//   for the gates and ex2 + rcp for tanh(c): 7.  The exponents are clamped to 2^30 so that the product of four
//   denominators stays finite (a gate below 2^-30 is returned as 2^-30).
template <bool PRECISE>
__device__ __forceinline__ float lstm_cell(float xi, float xf, float xg, float xo, float& c) {
  if constexpr (!PRECISE) {
    const float ig = sigmoid_fast(xi), fg = sigmoid_fast(xf), g_ = tanh_fast(xg), og = sigmoid_fast(xo);
    const float cc = fmaf(fg, c, ig * g_);
    c = cc;
    return og * tanh_fast(cc);
  } else {
    constexpr float L2E = 1.4426950408889634f;
    auto ex2c = [](float a) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(a, 30.f)));
      return e;
    };
    const float di = 1.f + ex2c(-L2E * xi), df = 1.f + ex2c(-L2E * xf);
    const float dg = 1.f + ex2c(-2.f * L2E * xg), d_o = 1.f + ex2c(-L2E * xo);
    const float p1 = di * df, p2 = dg * d_o;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p1 * p2));
    const float r12 = r * p2, r34 = r * p1;                       // 1 / (d_i d_f), 1 / (d_g d_o)
    const float ig = r12 * df, fg = r12 * di, og = r34 * dg;
    const float g_ = fmaf(2.f, r34 * d_o, -1.f);                   // tanh(x_g) = 2 sigmoid(2 x_g) - 1
    const float cc = fmaf(fg, c, ig * g_);
    c = cc;
    return og * tanh_acc(cc);
  }
}

template <int NFEAT, bool PRECISE, bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
k_tc_lstm(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapXlo,
          const __grid_constant__ CUtensorMap tmapW, TcLstmArgs p) {
  static_assert(!PRECISE || NFEAT == 64, "the hi/lo split variant is built for 64 input features");
  using L = TcLstmSmem<NFEAT, PRECISE>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sW = base + L::OFF_W, sX = base + L::OFF_X, sH = base + L::OFF_H;
  float* sBias = reinterpret_cast<float*>(smem + L::OFF_BIAS);
  const uint32_t bars = base + L::OFF_BAR;
  // barrier map (8 bytes each)
  const uint32_t bar_w = bars;                 // local: weights landed
  const uint32_t bar_xfull = bars + 8;         // [2] leader: x tiles of both CTAs landed
  const uint32_t bar_xempty = bars + 24;       // [2] both: MMAs finished reading the x stage
  const uint32_t bar_accfull = bars + 40;      // [4] both: gate pre-activations of chunk c complete
  const uint32_t bar_accempty = bars + 72;     // [4] leader: both CTAs' gate warps drained chunk c
  const uint32_t bar_hfull = bars + 104;       // leader: h_t of both CTAs in shared memory
  const uint32_t tmem_slot = bars + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int dir = blockIdx.y;
  const int tile = (blockIdx.x >> 1) * 2 + (int)rank;   // this CTA's 128-sequence tile
  const int len = p.len;

  if ((base & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
    }
    for (int c = 0; c < LSTM_CHUNKS; ++c) {
      mbar_init(bar_accfull + 8 * c, 1);
      mbar_init(bar_accempty + 8 * c, 16);  // 8 gate warps x 2 CTAs
    }
    mbar_init(bar_hfull, 16);
    fence_mbar_init();
    prefetch_tmap(&tmapX);
    prefetch_tmap(&tmapW);
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sBias[i] = p.bias[dir * 512 + i];
  // clean operand buffers: rows that no TMA box covers must not hold NaN bit patterns
  for (int i = threadIdx.x; i < (2 * L::X_STAGE + L::H_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + L::OFF_X)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::OFF_BAR + 128);

  // this CTA's half of the weights: packed rows (dir, rank, chunk, 64) x K, resident for the whole kernel
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, L::W_BYTES);
    const int wrow0 = (dir * 2 + (int)rank) * (LSTM_CHUNKS * 64);
    for (int c = 0; c < LSTM_CHUNKS; ++c)
      for (int kb = 0; kb < L::KBT; ++kb)
        tma_load_2d(sW + (c * L::KBT + kb) * L::W_BLOCK, &tmapW, bar_w, kb * 64, wrow0 + c * 64);
  }
  mbar_wait(bar_w, 0);
  cluster_sync();   // both halves of the weights are in place before the leader issues any MMA

  // tile geometry
  int c0 = 0, c1 = 0, c2 = 0;  // TMA coordinates other than feature/time
  if (p.mode == 0) {
    c0 = tile * 128;           // g0
  } else {
    c1 = (tile % p.kblocks) * p.Kc;   // k0
    c2 = (tile / p.kblocks) * p.Bc;   // b0
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // x tiles: both CTAs' loads complete on the leader's barrier
      uint32_t xfull_leader[2];
      for (int s = 0; s < 2; ++s)
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(xfull_leader[s]) : "r"(bar_xfull + 8 * s));
      const uint32_t box_bytes = (p.mode == 0 ? 128 : p.Kc * p.Bc) * 128;
      for (int step = 0; step < len; ++step) {
        const int t = dir == 0 ? step : len - 1 - step;
        const int s = step & 1, n = step >> 1;
        mbar_wait(bar_xempty + 8 * s, (n & 1) ^ 1);
        if (leader) mbar_expect_tx(bar_xfull + 8 * s, 2 * L::KBX * box_bytes);
        for (int kb = 0; kb < L::KBX; ++kb) {
          const uint32_t dst = sX + s * L::X_STAGE + kb * 16384;
          const CUtensorMap* tm = (PRECISE && kb == 1) ? &tmapXlo : &tmapX;   // k-block 1 = x_lo in PRECISE mode
          const int f0 = PRECISE ? 0 : kb * 64;
          if (p.mode == 0) tma_load_4d_cg2(dst, tm, xfull_leader[s], f0, t, c0, 0);
          else tma_load_4d_cg2(dst, tm, xfull_leader[s], f0, c1, t, c2);
        }
      }
    }
    __syncwarp();
  }

  constexpr uint32_t IDESC = idesc_f16(256, 128, 0);

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      auto issue_x = [&](int c, int s) {
        if constexpr (!PRECISE) {
#pragma unroll
          for (int k16 = 0; k16 < NFEAT / 16; ++k16) {
            const int kb = k16 >> 2, kk = k16 & 3;
            const uint64_t a = smem_desc_sw128_kmajor(sX + s * L::X_STAGE + kb * 16384) + (uint64_t)(kk * 2);
            const uint64_t b = smem_desc_sw128_kmajor(sW + (c * L::KBT + kb) * L::W_BLOCK) + (uint64_t)(kk * 2);
            umma_f16<2>(tmem + c * 128, a, b, IDESC, k16 > 0 ? 1u : 0u);
          }
        } else {
          // x_hi W_hi + x_lo W_hi + x_hi W_lo: (A k-block, B k-block) = (0,0), (1,0), (0,1)
#pragma unroll
          for (int t3 = 0; t3 < 3; ++t3) {
            const int akb = t3 == 1 ? 1 : 0, bkb = t3 == 2 ? 1 : 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t a = smem_desc_sw128_kmajor(sX + s * L::X_STAGE + akb * 16384) + (uint64_t)(kk * 2);
              const uint64_t b = smem_desc_sw128_kmajor(sW + (c * L::KBT + bkb) * L::W_BLOCK) + (uint64_t)(kk * 2);
              umma_f16<2>(tmem + c * 128, a, b, IDESC, (t3 > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
      };
      auto issue_h = [&](int c) {
#pragma unroll
        for (int k16 = 0; k16 < LSTM_H / 16; ++k16) {
          const int kb = k16 >> 2, kk = k16 & 3;
          const uint64_t a = smem_desc_sw128_kmajor(sH + kb * 16384) + (uint64_t)(kk * 2);
          const uint64_t b = smem_desc_sw128_kmajor(sW + (c * L::KBT + L::KBX + kb) * L::W_BLOCK) + (uint64_t)(kk * 2);
          umma_f16<2>(tmem + c * 128, a, b, IDESC, 1u);
        }
      };
      // step 0: x-part only (h_{-1} = 0)
      mbar_wait(bar_xfull, 0);
      tc_fence_after();
      for (int c = 0; c < LSTM_CHUNKS; ++c) {
        issue_x(c, 0);
        umma_commit_cg2(bar_accfull + 8 * c, 3);
      }
      umma_commit_cg2(bar_xempty, 3);
      const bool tr = TRACE && blockIdx.x == 0 && blockIdx.y == 0;   // debug timeline (tools/lstm_trace.py)
      for (int step = 1; step < len; ++step) {
        const int s = step & 1;
        long long* T = (tr && step >= 8 && step < 12) ? p.trace + (step - 8) * 32 : nullptr;
        if (T) T[0] = clock64();
        mbar_wait(bar_xfull + 8 * s, (step >> 1) & 1);
        if (T) T[1] = clock64();
        tc_fence_after();
        // x-part of this step for chunks 0..2 as soon as the gate warps have drained them (step-1)
        for (int c = 0; c < LSTM_CHUNKS - 1; ++c) {
          mbar_wait(bar_accempty + 8 * c, (step - 1) & 1);
          tc_fence_after();
          issue_x(c, s);
        }
        // h_{step-1} complete in both CTAs
        if (T) T[2] = clock64();
        mbar_wait(bar_hfull, (step - 1) & 1);
        if (T) T[3] = clock64();
        tc_fence_after();
        for (int c = 0; c < LSTM_CHUNKS - 1; ++c) {
          issue_h(c);
          umma_commit_cg2(bar_accfull + 8 * c, 3);
        }
        if (T) T[4] = clock64();
        mbar_wait(bar_accempty + 8 * (LSTM_CHUNKS - 1), (step - 1) & 1);
        if (T) T[5] = clock64();
        tc_fence_after();
        issue_x(LSTM_CHUNKS - 1, s);
        umma_commit_cg2(bar_xempty + 8 * s, 3);
        issue_h(LSTM_CHUNKS - 1);
        umma_commit_cg2(bar_accfull + 8 * (LSTM_CHUNKS - 1), 3);
        if (T) T[6] = clock64();
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ gate math
    const int gw = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int half = gw >> 2;          // which 16 of the chunk's 32 units
    const int r = q * 32 + lane;       // sequence row inside the tile
    // global output row of this sequence at time t: row_base + t * row_tstride (or invalid)
    long long row_base = 0, row_tstride = 0;
    bool valid;
    if (p.mode == 0) {
      const long long g = (long long)c0 + r;
      valid = g < p.G;
      row_base = g * p.C;
      row_tstride = 1;
    } else {
      const int bl = r / p.Kc, kl = r - bl * p.Kc;
      valid = (r < p.Kc * p.Bc) && (c2 + bl < p.B) && (c1 + kl < p.C);
      row_base = (long long)(c2 + bl) * p.S * p.C + (c1 + kl);
      row_tstride = p.C;
    }
    const int ldo = p.ndir * LSTM_H;
    float cst[LSTM_CHUNKS][16];
#pragma unroll
    for (int c = 0; c < LSTM_CHUNKS; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) cst[c][j] = 0.f;
    uint32_t accempty_leader[LSTM_CHUNKS], hfull_leader;
#pragma unroll
    for (int c = 0; c < LSTM_CHUNKS; ++c)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(accempty_leader[c]) : "r"(bar_accempty + 8 * c));
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(hfull_leader) : "r"(bar_hfull));

    const bool trg = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && gw == 0 && lane == 0;
    for (int step = 0; step < len; ++step) {
      const int t = dir == 0 ? step : len - 1 - step;
      long long* T = (trg && step >= 8 && step < 12) ? p.trace + (step - 8) * 32 + 8 : nullptr;
      uint32_t hp[LSTM_CHUNKS][8];   // packed fp16 h of this step (kept until all h-part MMAs have read sH)
      bool next_ready = false;       // accfull of the next chunk, probed while this chunk's gate math runs
#pragma unroll
      for (int c = 0; c < LSTM_CHUNKS; ++c) {
        if (T) T[3 * c] = clock64(); else asm volatile("" ::: "memory");
        if (!next_ready) mbar_wait(bar_accfull + 8 * c, step & 1);
        if (T) T[3 * c + 1] = clock64(); else asm volatile("" ::: "memory");
        tc_fence_after();
        // the ~90-cycle round trip of a (normally already complete) barrier test overlaps the gate math below
        next_ready = (c + 1 < LSTM_CHUNKS) && mbar_try_wait(bar_accfull + 8 * (c + 1), step & 1);
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + c * 128 + half * 16;
        __half* orow = p.out + (row_base + (long long)t * row_tstride) * ldo + dir * LSTM_H + c * LSTM_UNITS_PER_CHUNK + half * 16;
        // two passes of 8 hidden units keep the live register set small (no spills of the packed h values)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t gi[8], gf[8], gg[8], go[8];
          tmem_ld_32x32b_x8(taddr + sub * 8 + 0, gi);
          tmem_ld_32x32b_x8(taddr + sub * 8 + 32, gf);
          tmem_ld_32x32b_x8(taddr + sub * 8 + 64, gg);
          tmem_ld_32x32b_x8(taddr + sub * 8 + 96, go);
          tmem_ld_wait();
          if (sub == 1) {
            // accumulator chunk drained -> the MMA warp may start the next step's x-part into it
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
              asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(accempty_leader[c]) : "memory");
          }
          // biases: broadcast 16-byte reads (i | f | g | o blocks of 32 units per chunk)
          const float4* b4 = reinterpret_cast<const float4*>(sBias + c * 128 + half * 16 + sub * 8);
          float bi[8], bf[8], bg[8], bo[8];
          {
            const float4 i0 = b4[0], i1 = b4[1], f0 = b4[8], f1 = b4[9], g0 = b4[16], g1 = b4[17], o0 = b4[24], o1 = b4[25];
            bi[0] = i0.x; bi[1] = i0.y; bi[2] = i0.z; bi[3] = i0.w; bi[4] = i1.x; bi[5] = i1.y; bi[6] = i1.z; bi[7] = i1.w;
            bf[0] = f0.x; bf[1] = f0.y; bf[2] = f0.z; bf[3] = f0.w; bf[4] = f1.x; bf[5] = f1.y; bf[6] = f1.z; bf[7] = f1.w;
            bg[0] = g0.x; bg[1] = g0.y; bg[2] = g0.z; bg[3] = g0.w; bg[4] = g1.x; bg[5] = g1.y; bg[6] = g1.z; bg[7] = g1.w;
            bo[0] = o0.x; bo[1] = o0.y; bo[2] = o0.z; bo[3] = o0.w; bo[4] = o1.x; bo[5] = o1.y; bo[6] = o1.z; bo[7] = o1.w;
          }
          float hv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hv[j] = lstm_cell<PRECISE>(__uint_as_float(gi[j]) + bi[j], __uint_as_float(gf[j]) + bf[j],
                                       __uint_as_float(gg[j]) + bg[j], __uint_as_float(go[j]) + bo[j], cst[c][sub * 8 + j]);
          }
          uint32_t ho[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __half2 a = __floats2half2_rn(hv[2 * j], hv[2 * j + 1]);
            hp[c][sub * 4 + j] = *reinterpret_cast<const uint32_t*>(&a);
            const __half2 o = p.act ? __floats2half2_rn(fmaxf(hv[2 * j], 0.f), fmaxf(hv[2 * j + 1], 0.f)) : a;
            ho[j] = *reinterpret_cast<const uint32_t*>(&o);
          }
          if (valid) *reinterpret_cast<uint4*>(orow + sub * 8) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
          asm volatile("" ::: "memory");   // keep the two 8-unit passes apart (see DESIGN.md 3.5 on scheduling)
        }
        if (T) T[3 * c + 2] = clock64(); else asm volatile("" ::: "memory");
      }
      // acc_full of the last chunk implies every h-part MMA of this step has finished reading sH
      if (step + 1 < len) {
#pragma unroll
        for (int c = 0; c < LSTM_CHUNKS; ++c) {
          const int k = c * LSTM_UNITS_PER_CHUNK + half * 16;   // hidden-unit index = K index of the h operand
          const int kb = k >> 6, ch = (k & 63) >> 3;
          const uint32_t a0 = sH + kb * 16384 + sw128_offset((uint32_t)r, (uint32_t)ch);
          const uint32_t a1 = sH + kb * 16384 + sw128_offset((uint32_t)r, (uint32_t)ch + 1);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(hp[c][0]), "r"(hp[c][1]),
                       "r"(hp[c][2]), "r"(hp[c][3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(hp[c][4]), "r"(hp[c][5]),
                       "r"(hp[c][6]), "r"(hp[c][7]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(hfull_leader) : "memory");
        if (T) T[12] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc<2>(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// Ping-pong variant: two interleaved 64-row recurrences per CTA.  Default (plain fp16 N = 64 / 128 and PRECISE N = 64);
// VATSS_LSTM_PINGPONG=0 or vatss_debug_lstm_pingpong(0) select k_tc_lstm instead (bit-identical results).
//
// tcgen05.mma.cta_group::2 with M = 128 takes 64 rows from each CTA and leaves, in each CTA's TMEM, lanes 0-63 with D
// columns [0, N/2) and lanes 64-127 with D columns [N/2, N), both in TMEM columns [0, N/2) (tools/ubench/tmem_layout.cu).
// With N = 256 (two 128-column gate chunks per MMA) a half tile of 64 sequences needs 2 x 128 = 256 TMEM columns, so
// both half tiles A (rows 0-63 of this CTA's 128-row tile) and B (rows 64-127) fit.  While the gate warps work on one
// half, the recurrent MMAs of the other half's next step run: the recurrence bubble (fence -> arrive -> 8 MMAs ->
// commit -> tcgen05.ld, ~1400 of 8650 cycles per step) is hidden.
//   CTA r keeps the complete chunks {r, 2 + r} (all 128 gate columns x K) in shared memory: B rows of pair p.
//   gate thread: lane quadrant q -> row (q & 1) * 32 + lane of the half tile and chunk 2 p + (q >> 1) of pair p;
//   the two warps of a quadrant split the chunk's 32 units.
// ------------------------------------------------------------------------------------------
template <int NFEAT, bool PRECISE = false>
struct TcLstmPpSmem {
  static constexpr int KBX = PRECISE ? 2 : NFEAT / 64;   // x k-blocks (PRECISE: [x_hi | x_lo], weights [W_hi | W_lo | W_hh])
  static constexpr int KBT = KBX + 2;
  static constexpr int W_TILE = 128 * 128;           // 128 B-rows (one chunk) x 128 B
  static constexpr int W_BYTES = 2 * KBT * W_TILE;   // two chunk pairs
  static constexpr int X_STAGE = KBX * 16384;
  static constexpr int H_BYTES = 2 * 16384;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_X = OFF_W + W_BYTES;
  static constexpr int OFF_H = OFF_X + 2 * X_STAGE;
  static constexpr int OFF_BIAS = OFF_H + H_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + 512 * 4;
  static constexpr int TOTAL = OFF_BAR + 256;
};

//
// GROUPS (row groups per CTA).  The gate math is MUFU-bound and a MUFU warp instruction costs the same whatever the
// number of active lanes, so the unit of gate work is a ROW GROUP: 32 sequences = one warp-wide TMEM lane quadrant.
// GROUPS = 4 is the tile above (slots (half, lane-half e) = four different groups).  When there are fewer groups than
// 4 x SMs (the inter-chunk phase: 150 chunk positions x B utterances x 2 directions), a CTA takes only 3 or 2 groups
// and a group occupies BOTH 32-row slots of a half tile: the TMA box is loaded twice, the recurrent state h is written
// to both rows, so lanes 0-31 and 32-63 (and 64-95 / 96-127 for the other chunk of the pair) carry the same
// pre-activations and the four lane quadrants = four SM sub-partitions share the group's hidden units (8 instead of 16
// per thread and chunk).  Gate time per step is then proportional to GROUPS (3: half A = two groups, half B = one
// duplicated group; 2: both halves duplicated) and the recurrence spreads over up to twice as many SMs.  Results are
// bit-identical for every GROUPS (a row's accumulator does not depend on its neighbours).
//   slot geometry for GROUPS < 4: slot (h, e) = rows h*64 + e*32 .. +31 of the CTA's operand tiles; a group is
//   intra: 32 consecutive sequences, inter: Kc <= 32 chunk positions of one utterance (one TMA box per slot).
template <int GROUPS>
__device__ __forceinline__ int lstm_slot_group(int tile, int h, int e) {
  if (GROUPS == 4) return tile * 4 + h * 2 + e;
  if (GROUPS == 3) return tile * 3 + (h == 0 ? e : 2);
  return tile * 2 + h;
}

template <int NFEAT, bool PRECISE, bool TRACE, int GROUPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LSTM_THREADS, 1)
k_tc_lstm_pp(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapXlo,
             const __grid_constant__ CUtensorMap tmapW, TcLstmArgs p) {
  static_assert(!PRECISE || NFEAT == 64, "the hi/lo split variant is built for 64 input features");
  using L = TcLstmPpSmem<NFEAT, PRECISE>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sW = base + L::OFF_W, sX = base + L::OFF_X, sH = base + L::OFF_H;
  float* sBias = reinterpret_cast<float*>(smem + L::OFF_BIAS);
  const uint32_t bars = base + L::OFF_BAR;
  const uint32_t bar_w = bars;                 // local: weights landed
  const uint32_t bar_xfull = bars + 8;         // [2] leader: x tiles of both CTAs landed
  const uint32_t bar_xempty = bars + 24;       // [2] both: MMAs finished reading the x stage
  const uint32_t bar_accfull = bars + 40;      // [half * 2 + pair] both: gate pre-activations complete
  const uint32_t bar_accempty = bars + 72;     // [half * 2 + pair] leader: both CTAs' gate warps drained it
  const uint32_t bar_hfull = bars + 104;       // [half] leader: h_t of that half in shared memory (both CTAs)
  const uint32_t tmem_slot = bars + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int dir = blockIdx.y;
  const int tile = (blockIdx.x >> 1) * 2 + (int)rank;
  const int len = p.len;
  if ((base & 1023u) != 0) __trap();

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
      mbar_init(bar_hfull + 8 * s, 16);
    }
    for (int c = 0; c < 4; ++c) {
      mbar_init(bar_accfull + 8 * c, 1);
      mbar_init(bar_accempty + 8 * c, 16);  // 8 gate warps x 2 CTAs
    }
    fence_mbar_init();
    prefetch_tmap(&tmapX);
    prefetch_tmap(&tmapW);
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sBias[i] = p.bias[dir * 512 + i];
  for (int i = threadIdx.x; i < (2 * L::X_STAGE + L::H_BYTES) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + L::OFF_X)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::OFF_BAR + 128);

  // weights: pair pr -> chunk 2 pr + rank, all 128 gate columns = packed rows (dir, 0, chunk, 64) and (dir, 1, chunk, 64)
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_w, L::W_BYTES);
    for (int pr = 0; pr < 2; ++pr) {
      const int chunk = 2 * pr + (int)rank;
      for (int kb = 0; kb < L::KBT; ++kb)
        for (int r2 = 0; r2 < 2; ++r2)
          tma_load_2d(sW + (pr * L::KBT + kb) * L::W_TILE + r2 * 8192, &tmapW, bar_w, kb * 64,
                      (dir * 2 + r2) * (LSTM_CHUNKS * 64) + chunk * 64);
    }
  }
  mbar_wait(bar_w, 0);
  cluster_sync();

  int c0 = 0, c1 = 0, c2 = 0;
  if (p.mode == 0) {
    c0 = tile * 128;
  } else {
    c1 = (tile % p.kblocks) * p.Kc;
    c2 = (tile / p.kblocks) * p.Bc;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (as in k_tc_lstm)
    if (lane == 0) {
      uint32_t xfull_leader[2];
      for (int s = 0; s < 2; ++s)
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(xfull_leader[s]) : "r"(bar_xfull + 8 * s));
      const uint32_t box_bytes = GROUPS == 4 ? (p.mode == 0 ? 128 : p.Kc * p.Bc) * 128
                                             : 4 * (p.mode == 0 ? 32 : p.Kc) * 128;   // four slot boxes
      int sa[4], sb[4];   // GROUPS < 4: per slot, intra: first sequence / inter: first chunk position and utterance
      if constexpr (GROUPS < 4) {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          const int gid = lstm_slot_group<GROUPS>(tile, sl >> 1, sl & 1);
          if (p.mode == 0) { sa[sl] = gid * 32; sb[sl] = 0; }
          else { sa[sl] = (gid % p.kblocks) * p.Kc; sb[sl] = gid / p.kblocks; }   // gid past the end: b >= B, zero fill
        }
      }
      for (int step = 0; step < len; ++step) {
        const int t = dir == 0 ? step : len - 1 - step;
        const int s = step & 1, n = step >> 1;
        mbar_wait(bar_xempty + 8 * s, (n & 1) ^ 1);
        if (leader) mbar_expect_tx(bar_xfull + 8 * s, 2 * L::KBX * box_bytes);
        for (int kb = 0; kb < L::KBX; ++kb) {
          const uint32_t dst = sX + s * L::X_STAGE + kb * 16384;
          const CUtensorMap* tm = (PRECISE && kb == 1) ? &tmapXlo : &tmapX;   // k-block 1 = x_lo in PRECISE mode
          const int f0 = PRECISE ? 0 : kb * 64;
          if constexpr (GROUPS == 4) {
            if (p.mode == 0) tma_load_4d_cg2(dst, tm, xfull_leader[s], f0, t, c0, 0);
            else tma_load_4d_cg2(dst, tm, xfull_leader[s], f0, c1, t, c2);
          } else {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              if (p.mode == 0) tma_load_4d_cg2(dst + sl * 4096, tm, xfull_leader[s], f0, t, sa[sl], 0);
              else tma_load_4d_cg2(dst + sl * 4096, tm, xfull_leader[s], f0, sa[sl], t, sb[sl]);
            }
          }
        }
      }
    }
    __syncwarp();
  }

  constexpr uint32_t IDESC = idesc_f16(128, 256, 0);

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      auto issue_x = [&](int h, int pr, int s) {
        if constexpr (!PRECISE) {
#pragma unroll
          for (int k16 = 0; k16 < NFEAT / 16; ++k16) {
            const int kb = k16 >> 2, kk = k16 & 3;
            const uint64_t a = smem_desc_sw128_kmajor(sX + s * L::X_STAGE + kb * 16384 + h * 8192) + (uint64_t)(kk * 2);
            const uint64_t b = smem_desc_sw128_kmajor(sW + (pr * L::KBT + kb) * L::W_TILE) + (uint64_t)(kk * 2);
            umma_f16<2>(tmem + h * 256 + pr * 128, a, b, IDESC, k16 > 0 ? 1u : 0u);
          }
        } else {
          // x_hi W_hi + x_lo W_hi + x_hi W_lo: (A k-block, B k-block) = (0,0), (1,0), (0,1)
#pragma unroll
          for (int t3 = 0; t3 < 3; ++t3) {
            const int akb = t3 == 1 ? 1 : 0, bkb = t3 == 2 ? 1 : 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t a = smem_desc_sw128_kmajor(sX + s * L::X_STAGE + akb * 16384 + h * 8192) + (uint64_t)(kk * 2);
              const uint64_t b = smem_desc_sw128_kmajor(sW + (pr * L::KBT + bkb) * L::W_TILE) + (uint64_t)(kk * 2);
              umma_f16<2>(tmem + h * 256 + pr * 128, a, b, IDESC, (t3 > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
      };
      auto issue_h = [&](int h, int pr) {
#pragma unroll
        for (int k16 = 0; k16 < LSTM_H / 16; ++k16) {
          const int kb = k16 >> 2, kk = k16 & 3;
          const uint64_t a = smem_desc_sw128_kmajor(sH + kb * 16384 + h * 8192) + (uint64_t)(kk * 2);
          const uint64_t b = smem_desc_sw128_kmajor(sW + (pr * L::KBT + L::KBX + kb) * L::W_TILE) + (uint64_t)(kk * 2);
          umma_f16<2>(tmem + h * 256 + pr * 128, a, b, IDESC, 1u);
        }
      };
      // step 0: x-part only (h_{-1} = 0)
      mbar_wait(bar_xfull, 0);
      tc_fence_after();
      for (int h = 0; h < 2; ++h)
        for (int pr = 0; pr < 2; ++pr) {
          issue_x(h, pr, 0);
          umma_commit_cg2(bar_accfull + 8 * (h * 2 + pr), 3);
        }
      umma_commit_cg2(bar_xempty, 3);
      const bool tr = TRACE && blockIdx.x == 0 && blockIdx.y == 0;
      for (int step = 1; step < len; ++step) {
        const int s = step & 1;
        const uint32_t prev = (uint32_t)(step - 1) & 1u;
        long long* T = (tr && step >= 8 && step < 12) ? p.trace + (step - 8) * 32 : nullptr;
        if (T) T[0] = clock64();
        mbar_wait(bar_xfull + 8 * s, (step >> 1) & 1);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
          for (int pr = 0; pr < 2; ++pr) {   // x-part as soon as the gate warps have drained the accumulator (step - 1)
            mbar_wait(bar_accempty + 8 * (h * 2 + pr), prev);
            tc_fence_after();
            issue_x(h, pr, s);
          }
          if (h == 1) umma_commit_cg2(bar_xempty + 8 * s, 3);
          if (T) T[1 + 3 * h] = clock64();
          mbar_wait(bar_hfull + 8 * h, prev);   // h_{step-1} of this half complete in both CTAs
          if (T) T[2 + 3 * h] = clock64();
          tc_fence_after();
          for (int pr = 0; pr < 2; ++pr) {
            issue_h(h, pr);
            umma_commit_cg2(bar_accfull + 8 * (h * 2 + pr), 3);
          }
          if (T) T[3 + 3 * h] = clock64();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ gate math
    const int gw = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int hs = gw >> 2;            // which 16 of the chunk's 32 units
    const int pc = q >> 1;             // which chunk of a pair lives in this lane half
    const int ldo = p.ndir * LSTM_H;
    // global output row of this thread's sequence in each half tile
    long long row_base[2], row_tstride = 0;
    bool valid[2];
    int rrow[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = h * 64 + (q & 1) * 32 + lane;   // row inside this CTA's 128-row tile
      rrow[h] = r;
      if constexpr (GROUPS == 4) {
        if (p.mode == 0) {
          const long long g = (long long)c0 + r;
          valid[h] = g < p.G;
          row_base[h] = g * p.C;
          row_tstride = 1;
        } else {
          const int bl = r / p.Kc, kl = r - bl * p.Kc;
          valid[h] = (r < p.Kc * p.Bc) && (c2 + bl < p.B) && (c1 + kl < p.C);
          row_base[h] = (long long)(c2 + bl) * p.S * p.C + (c1 + kl);
          row_tstride = p.C;
        }
      } else {
        const int gid = lstm_slot_group<GROUPS>(tile, h, q & 1);
        if (p.mode == 0) {
          const long long g = (long long)gid * 32 + lane;
          valid[h] = g < p.G;
          row_base[h] = g * p.C;
          row_tstride = 1;
        } else {
          const int bb = gid / p.kblocks, kk = (gid % p.kblocks) * p.Kc + lane;
          valid[h] = (lane < p.Kc) && (kk < p.C) && (bb < p.B);
          row_base[h] = (long long)bb * p.S * p.C + kk;
          row_tstride = p.C;
        }
      }
    }
    float cst[2][2][16];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int pr = 0; pr < 2; ++pr)
#pragma unroll
        for (int j = 0; j < 16; ++j) cst[h][pr][j] = 0.f;
    uint32_t accempty_leader[4], hfull_leader[2];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(accempty_leader[c]) : "r"(bar_accempty + 8 * c));
#pragma unroll
    for (int h = 0; h < 2; ++h)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(hfull_leader[h]) : "r"(bar_hfull + 8 * h));

    const bool trg = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && gw == 0 && lane == 0;
    for (int step = 0; step < len; ++step) {
      const int t = dir == 0 ? step : len - 1 - step;
      long long* T = (trg && step >= 8 && step < 12) ? p.trace + (step - 8) * 32 + 8 : nullptr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool dup = GROUPS == 2 || (GROUPS == 3 && h == 1);   // this half holds one group twice (constant after unrolling)
        uint32_t hp[2][8];   // packed fp16 h of this half (kept until both pairs' recurrent MMAs have read sH)
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int chunk = 2 * pr + pc;
          if (T) T[3 * (h * 2 + pr)] = clock64(); else asm volatile("" ::: "memory");
          mbar_wait(bar_accfull + 8 * (h * 2 + pr), step & 1);
          if (T) T[3 * (h * 2 + pr) + 1] = clock64(); else asm volatile("" ::: "memory");
          tc_fence_after();
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + h * 256 + pr * 128 + hs * 16;
          __half* orow = p.out + (row_base[h] + (long long)t * row_tstride) * ldo + dir * LSTM_H +
                         chunk * LSTM_UNITS_PER_CHUNK + hs * 16;
#if LSTM_PREFETCH
          // The second 8-unit pass's accumulator loads are issued before the first pass's arithmetic: the two gate warps
          // of a scheduler run in step (same barriers), so a tcgen05.ld round trip in front of every pass was exposed
          // (~200 of the ~860 cycles of a pass).
          uint32_t gq[2][4][8];
          {
            const int uo0 = dup ? (q & 1) * 8 : 0;
            tmem_ld_32x32b_x8(taddr + uo0 + 0, gq[0][0]);
            tmem_ld_32x32b_x8(taddr + uo0 + 32, gq[0][1]);
            tmem_ld_32x32b_x8(taddr + uo0 + 64, gq[0][2]);
            tmem_ld_32x32b_x8(taddr + uo0 + 96, gq[0][3]);
            tmem_ld_wait();
            if (!dup) {
              tmem_ld_32x32b_x8(taddr + 8 + 0, gq[1][0]);
              tmem_ld_32x32b_x8(taddr + 8 + 32, gq[1][1]);
              tmem_ld_32x32b_x8(taddr + 8 + 64, gq[1][2]);
              tmem_ld_32x32b_x8(taddr + 8 + 96, gq[1][3]);
            }
          }
#endif
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (dup && sub == 1) break;   // a duplicated group: this lane half owns 8 of the 16 units, one pass
            const int uo = dup ? (q & 1) * 8 : sub * 8;
#if LSTM_PREFETCH
            uint32_t (&gi)[8] = gq[sub][0], (&gf)[8] = gq[sub][1], (&gg)[8] = gq[sub][2], (&go)[8] = gq[sub][3];
            if (sub == 1) tmem_ld_wait();
#else
            uint32_t gi[8], gf[8], gg[8], go[8];
            tmem_ld_32x32b_x8(taddr + uo + 0, gi);
            tmem_ld_32x32b_x8(taddr + uo + 32, gf);
            tmem_ld_32x32b_x8(taddr + uo + 64, gg);
            tmem_ld_32x32b_x8(taddr + uo + 96, go);
            tmem_ld_wait();
#endif
            if (sub == 1 || dup) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(accempty_leader[h * 2 + pr]) : "memory");
            }
            const float4* b4 = reinterpret_cast<const float4*>(sBias + chunk * 128 + hs * 16 + uo);
            float bi[8], bf[8], bg[8], bo[8];
            {
              const float4 i0 = b4[0], i1 = b4[1], f0 = b4[8], f1 = b4[9], g0 = b4[16], g1 = b4[17], o0 = b4[24], o1 = b4[25];
              bi[0] = i0.x; bi[1] = i0.y; bi[2] = i0.z; bi[3] = i0.w; bi[4] = i1.x; bi[5] = i1.y; bi[6] = i1.z; bi[7] = i1.w;
              bf[0] = f0.x; bf[1] = f0.y; bf[2] = f0.z; bf[3] = f0.w; bf[4] = f1.x; bf[5] = f1.y; bf[6] = f1.z; bf[7] = f1.w;
              bg[0] = g0.x; bg[1] = g0.y; bg[2] = g0.z; bg[3] = g0.w; bg[4] = g1.x; bg[5] = g1.y; bg[6] = g1.z; bg[7] = g1.w;
              bo[0] = o0.x; bo[1] = o0.y; bo[2] = o0.z; bo[3] = o0.w; bo[4] = o1.x; bo[5] = o1.y; bo[6] = o1.z; bo[7] = o1.w;
            }
            float hv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hv[j] = lstm_cell<PRECISE>(__uint_as_float(gi[j]) + bi[j], __uint_as_float(gf[j]) + bf[j],
                                         __uint_as_float(gg[j]) + bg[j], __uint_as_float(go[j]) + bo[j], cst[h][pr][sub * 8 + j]);
            }
            uint32_t ho[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __half2 a = __floats2half2_rn(hv[2 * j], hv[2 * j + 1]);
              hp[pr][sub * 4 + j] = *reinterpret_cast<const uint32_t*>(&a);
              const __half2 o = p.act ? __floats2half2_rn(fmaxf(hv[2 * j], 0.f), fmaxf(hv[2 * j + 1], 0.f)) : a;
              ho[j] = *reinterpret_cast<const uint32_t*>(&o);
            }
            if (valid[h]) *reinterpret_cast<uint4*>(orow + uo) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
            asm volatile("" ::: "memory");
          }
          if (T) T[3 * (h * 2 + pr) + 2] = clock64(); else asm volatile("" ::: "memory");
        }
        // accfull of the second pair implies every recurrent MMA of this half and step has finished reading sH
        if (step + 1 < len) {
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int k = (2 * pr + pc) * LSTM_UNITS_PER_CHUNK + hs * 16;   // hidden-unit index = K index of the h operand
            const int kb = k >> 6, ch = (k & 63) >> 3;
            if (!dup) {
              const uint32_t a0 = sH + kb * 16384 + sw128_offset((uint32_t)rrow[h], (uint32_t)ch);
              const uint32_t a1 = sH + kb * 16384 + sw128_offset((uint32_t)rrow[h], (uint32_t)ch + 1);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(hp[pr][0]), "r"(hp[pr][1]),
                           "r"(hp[pr][2]), "r"(hp[pr][3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(hp[pr][4]), "r"(hp[pr][5]),
                           "r"(hp[pr][6]), "r"(hp[pr][7]) : "memory");
            } else {
              // this thread's 8 units of the sequence go to both copies of the row (slots e = 0 and e = 1)
              const uint32_t che = (uint32_t)ch + (uint32_t)(q & 1);
              const uint32_t a0 = sH + kb * 16384 + sw128_offset((uint32_t)(h * 64 + lane), che);
              const uint32_t a1 = sH + kb * 16384 + sw128_offset((uint32_t)(h * 64 + 32 + lane), che);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(hp[pr][0]), "r"(hp[pr][1]),
                           "r"(hp[pr][2]), "r"(hp[pr][3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(hp[pr][0]), "r"(hp[pr][1]),
                           "r"(hp[pr][2]), "r"(hp[pr][3]) : "memory");
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(hfull_leader[h]) : "memory");
          if (T) T[12 + h] = clock64();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc<2>(tmem, 512);
}

int g_lstm_pingpong = -1;   // -1: VATSS_LSTM_PINGPONG from the environment on first use (default 1); else 0 / 1
int g_lstm_groups = -1;     // -1: VATSS_LSTM_GROUPS from the environment (default 0 = automatic); 2 / 3 / 4 force it

// Row groups per CTA of k_tc_lstm_pp for `ngroups` groups per direction: the fewest machine waves x time per step.
// Time per step measured on B200 (tools/lstm_groups_bench.py, 283 steps): 4.02 / 3.32 / 2.95 us for 4 / 3 / 2 groups -
// about 0.54 us of gate math per group on top of the 64 M = 128 MMAs of a step, which cost the same for every GROUPS.
static int pick_lstm_groups(int ngroups, int ndir) {
  if (g_lstm_groups < 0) {
    const char* e = getenv("VATSS_LSTM_GROUPS");
    g_lstm_groups = e ? atoi(e) : 0;
  }
  if (g_lstm_groups >= 2 && g_lstm_groups <= 4) return g_lstm_groups;
  const int sms = num_sms();
  int best = 4;
  long long best_cost = -1;
  for (int g = 4; g >= 2; --g) {
    const int tiles = (ngroups + g - 1) / g;
    const int ctas = 2 * ((tiles + 1) / 2) * ndir;
    const long long waves = (ctas + sms - 1) / sms;
    const long long cost = waves * (g == 4 ? 4020 : g == 3 ? 3320 : 2950);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = g; }
  }
  return best;
}

template <int NFEAT, bool PRECISE, int GROUPS>
static int tc_lstm_pp_launch(const CUtensorMap& tmX, const CUtensorMap& tmXlo, const CUtensorMap& tmW,
                             const TcLstmArgs& a, cudaStream_t st) {
  using LP = TcLstmPpSmem<NFEAT, PRECISE>;
  static_assert(LP::TOTAL <= 227 * 1024, "shared memory budget");
  auto kpp = a.trace ? k_tc_lstm_pp<NFEAT, PRECISE, true, GROUPS> : k_tc_lstm_pp<NFEAT, PRECISE, false, GROUPS>;
  static PerDeviceOnce configured_pp;
  if (configured_pp.first()) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_lstm_pp<NFEAT, PRECISE, true, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::TOTAL));
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_tc_lstm_pp<NFEAT, PRECISE, false, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::TOTAL));
  }
  const int pairs_pp = (a.num_tiles + 1) / 2;
  dim3 grid_pp(2 * pairs_pp, a.ndir);
  kpp<<<grid_pp, LSTM_THREADS, LP::TOTAL, st>>>(tmX, tmXlo, tmW, a);
  VATSS_LAUNCH_OK();
  return 0;
}

template <int NFEAT, bool PRECISE, bool TRACE = false>
static int tc_lstm_launch(const __half* x16, const __half* x16lo, const __half* Wpack, TcLstmArgs a,
                          cudaStream_t st) {
  if (!TRACE && a.trace != nullptr) return tc_lstm_launch<NFEAT, PRECISE, true>(x16, x16lo, Wpack, a, st);
  using L = TcLstmSmem<NFEAT, PRECISE>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  if (g_lstm_pingpong < 0) {
    const char* e = getenv("VATSS_LSTM_PINGPONG");
    g_lstm_pingpong = e ? atoi(e) : 1;
  }
  // row groups per CTA (k_tc_lstm_pp only); groups < 4 use one TMA box per 32-row slot
  int groups = 4;
  if (g_lstm_pingpong) {
    const int kc_s = a.mode == 0 ? 32 : (a.C + (a.C + 31) / 32 - 1) / ((a.C + 31) / 32);   // <= 32 chunk positions
    const int kb_s = a.mode == 0 ? 1 : (a.C + kc_s - 1) / kc_s;
    const int ngroups = a.mode == 0 ? (a.G + 31) / 32 : kb_s * a.B;
    groups = pick_lstm_groups(ngroups, a.ndir);
    // the 4-group tile of the inter phase packs Kc x Bc rows densely; keep it unless spreading was chosen
    if (groups < 4) {
      if (a.mode == 1) { a.Kc = kc_s; a.Bc = 1; a.kblocks = kb_s; }
      a.num_tiles = (ngroups + groups - 1) / groups;
    }
  }
  CUtensorMap tmX, tmXlo, tmW;
  auto make_x = [&](CUtensorMap* tm, const __half* ptr) -> int {
    if (a.mode == 0) {
      const uint64_t dims[4] = {(uint64_t)NFEAT, (uint64_t)a.C, (uint64_t)a.G, 1};
      const uint64_t str[3] = {(uint64_t)NFEAT * 2, (uint64_t)a.C * NFEAT * 2, (uint64_t)a.G * a.C * NFEAT * 2};
      const uint32_t box[4] = {64, 1, groups < 4 ? 32u : 128u, 1};
      return make_tmap_f16(tm, ptr, 4, dims, str, box);
    }
    const uint64_t dims[4] = {(uint64_t)NFEAT, (uint64_t)a.C, (uint64_t)a.S, (uint64_t)a.B};
    const uint64_t str[3] = {(uint64_t)NFEAT * 2, (uint64_t)a.C * NFEAT * 2, (uint64_t)a.S * a.C * NFEAT * 2};
    const uint32_t box[4] = {64, (uint32_t)a.Kc, 1, (uint32_t)a.Bc};
    return make_tmap_f16(tm, ptr, 4, dims, str, box);
  };
  if (make_x(&tmX, x16)) return -1;
  if (make_x(&tmXlo, PRECISE ? x16lo : x16)) return -1;
  {
    const uint64_t ktot = (PRECISE ? 2 * NFEAT : NFEAT) + LSTM_H;
    const uint64_t dims[2] = {ktot, (uint64_t)a.ndir * 2 * LSTM_CHUNKS * 64};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t box[2] = {64, 64};
    if (make_tmap_f16(&tmW, Wpack, 2, dims, str, box)) return -1;
  }
  if (g_lstm_pingpong) {
    if (groups == 2) return tc_lstm_pp_launch<NFEAT, PRECISE, 2>(tmX, tmXlo, tmW, a, st);
    if (groups == 3) return tc_lstm_pp_launch<NFEAT, PRECISE, 3>(tmX, tmXlo, tmW, a, st);
    return tc_lstm_pp_launch<NFEAT, PRECISE, 4>(tmX, tmXlo, tmW, a, st);
  }
  auto kern = k_tc_lstm<NFEAT, PRECISE, TRACE>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  const int pairs = (a.num_tiles + 1) / 2;
  dim3 grid(2 * pairs, a.ndir);
  kern<<<grid, LSTM_THREADS, L::TOTAL, st>>>(tmX, tmXlo, tmW, a);
  VATSS_LAUNCH_OK();
  return 0;
}

// choose the inter-chunk tile shape (Kc chunk positions x Bc utterances, Kc*Bc <= 128) with the fewest tiles
static void pick_inter_tile(int C, int B, int* Kc, int* Bc) {
  long long best = -1;
  for (int k = 1; k <= 128 && k <= C; ++k) {
    int b = 128 / k;
    if (b > B) b = B;
    if (b < 1) continue;
    const long long tiles = (long long)((C + k - 1) / k) * ((B + b - 1) / b);
    if (best < 0 || tiles < best || (tiles == best && k * b > (*Kc) * (*Bc))) {
      best = tiles; *Kc = k; *Bc = b;
    }
  }
}

long long* g_lstm_trace = nullptr;   // debug: set through vatss_debug_lstm_trace

int launch_tc_lstm(const __half* x16, const __half* x16lo, const __half* Wpack, const float* bias_pack, __half* out16,
                   int mode, int B, int S, int C, int NFEAT, int ndir, int act, cudaStream_t st) {
  TcLstmArgs a;
  a.mode = mode; a.ndir = ndir; a.B = B; a.S = S; a.C = C; a.act = act; a.bias = bias_pack; a.out = out16;
  a.Kc = 0; a.Bc = 0; a.kblocks = 1; a.G = 0;
  a.trace = g_lstm_trace;
  if (mode == 0) {
    a.len = C;
    a.G = B * S;
    a.num_tiles = (a.G + 127) / 128;
  } else {
    a.len = S;
    pick_inter_tile(C, B, &a.Kc, &a.Bc);
    a.kblocks = (C + a.Kc - 1) / a.Kc;
    a.num_tiles = a.kblocks * ((B + a.Bc - 1) / a.Bc);
  }
  if (a.num_tiles == 0 || a.len == 0) return 0;
  if (x16lo != nullptr) {
    VATSS_CHECK_ARG(NFEAT == 64, "tc_lstm: the hi/lo split variant needs num_features = 64");
    return tc_lstm_launch<64, true>(x16, x16lo, Wpack, a, st);
  }
  if (NFEAT == 128) return tc_lstm_launch<128, false>(x16, nullptr, Wpack, a, st);
  if (NFEAT == 64) return tc_lstm_launch<64, false>(x16, nullptr, Wpack, a, st);
  set_error("tc_lstm: num_features %d unsupported (64 or 128)", NFEAT);
  return -1;
}

// ------------------------------------------------------------------------------------------
// weight packing: fp32 nn.LSTM parameters -> per (direction, CTA rank, chunk) B-operand rows
//   packed row (dir, rank, c, j): gate = 2*rank + j/32, unit = 32*c + j%32, source row = gate*128 + unit
//   columns [0,N) = W_ih row, [N, N+128) = W_hh row
//   bias_pack[dir][128*c + 32*gate + u] = b_ih + b_hh of (gate, unit 32*c+u)   (accumulator-column order)
//   plain fp16 path (precise = 0): rows and biases of the sigmoid gates i, f, o are stored times 0.5 (sigmoid_fast)
// ------------------------------------------------------------------------------------------
__global__ void k_pack_lstm(const float* __restrict__ Wih, const float* __restrict__ Whh,
                            const float* __restrict__ bih, const float* __restrict__ bhh, int N, int dir, int precise,
                            __half* __restrict__ Wpack, float* __restrict__ bias_pack) {
  const int nx = precise ? 2 * N : N;   // precise: [half(W) | half(W - half(W))]
  const int ktot = nx + LSTM_H;
  const int prow = blockIdx.x;  // 0..511 within this direction: (rank, c, j)
  const int rank = prow / 256, c = (prow % 256) / 64, j = prow % 64;
  const int gate = 2 * rank + j / 32, unit = 32 * c + j % 32;
  const int src = gate * LSTM_H + unit;
  // plain fp16 path: sigmoid(x) = 0.5 + 0.5 tanh(x / 2) - the halving of the i, f, o pre-activations is folded in here
  const float gscale = (!precise && gate != 2) ? 0.5f : 1.0f;
  __half* dst = Wpack + ((size_t)dir * 512 + prow) * ktot;
  for (int k = threadIdx.x; k < ktot; k += blockDim.x) {
    float v;
    if (k < N) v = Wih[(size_t)src * N + k];
    else if (k < nx) {
      const float w = Wih[(size_t)src * N + (k - N)];
      v = w - __half2float(__float2half_rn(w));
    } else v = Whh[(size_t)src * LSTM_H + (k - nx)];
    dst[k] = __float2half_rn(v * gscale);
  }
  if (threadIdx.x == 0) bias_pack[dir * 512 + 128 * c + 32 * gate + unit % 32] = (bih[src] + bhh[src]) * gscale;
}

int launch_pack_lstm(const float* Wih, const float* Whh, const float* bih, const float* bhh, int N, int dir,
                     int precise, __half* Wpack, float* bias_pack, cudaStream_t st) {
  k_pack_lstm<<<512, 128, 0, st>>>(Wih, Whh, bih, bhh, N, dir, precise, Wpack, bias_pack);
  VATSS_LAUNCH_OK();
  return 0;
}

}  // namespace vatss

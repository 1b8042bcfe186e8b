// Lipreader trunk convolutions on the tcgen05 tensor cores (reference: models/resnet.py:8-17,31-84 - the 3x3 / 1x1
// Conv2d + BatchNorm2d + activation (+ shortcut) of the ResNet-18 BasicBlocks).
//
// Implicit GEMM over pixel-major fp16 activations:
//   D[128 pixels, NT channels] += A_slab[128 pixels, 64 input channels of one tap] x W_slab[NT, 64]^T
// for the taps x (Cin / 64) K slabs of the convolution.  K = 9 Cin is 576 ... 4608, so - unlike the per-token
// projections of the separation path (tc_gemm.cu: K <= 256, W resident) - this is a K-pipelined kernel:
//   warps 0..7  A producers: the 128-byte channel vector of every output pixel's tap neighbour (zeros outside the
//               image) goes into the SWIZZLE_128B K-major operand layout with 16-byte shared-memory stores (eight lanes
//               per pixel: whole lines per warp instruction); any stride / padding / image size, every tile has 128
//               useful rows.  The loads of three slabs are in flight per thread.  Thread 0 also fetches the weight slab with one TMA box.
//               After the last slab the same warps are the epilogue: tcgen05.ld of their TMEM lane quadrant, folded
//               BatchNorm, shortcut add, activation, fp16 store.
//   warp 8      MMA issuer: 4 x tcgen05.mma (M = 128, N = NT, K = 16) per slab, accumulator in TMEM, tcgen05.commit
//               hands the stage back to the producers.
// One tile per CTA, two (128 output channels) or three (64) CTAs per SM: one CTA's epilogue runs under the others' main loops.
#include "lipreader.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

int g_lip_dbg = 0;
long long* g_lip_trace = nullptr;   // debug: per-CTA clock64 / globaltimer stamps of the next launches (NULL in production)
constexpr int LT_STAGES = 3;
constexpr int LT_A_BYTES = 128 * 128;

struct LipTcArgs {
  const __half* in;
  int F, H, W, Cin, Ho, Wo, Cout, ks, stride, pad;
  const float* scale; const float* shift; const float* slope;
  const __half* res;
  int act;
  __half* out;
  long long* trace;   // debug: 8 stamps per CTA for the first 1024 CTAs
  int dbg;   // experiments (vatss_debug_lipreader): bit 0 no gather loads, bit 1 no epilogue stores, bit 2 no weight TMA
};

__device__ __forceinline__ float lip_act_tc(float v, int act, float slope) {
  if (act == LIP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == LIP_ACT_PRELU) return v >= 0.f ? v : v * slope;
  // x sigmoid(x) with the approximate exponential and division (MUFU.EX2 + MUFU.RCP, no IEEE slow path): the result is
  // rounded to fp16 anyway; the IEEE division cost 1.0 of 9.7 ms per 3200 frames (tools/lipreader_ablate.py)
  if (act == LIP_ACT_SWISH) return __fdividef(v, 1.0f + __expf(-v));
  return v;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(288, MINB) k_lip_conv_tc(const __grid_constant__ CUtensorMap tmapW, const LipTcArgs a) {
  constexpr int B_BYTES = NT * 128;
  constexpr int STAGE_BYTES = LT_A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + LT_STAGES * STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * LT_STAGES, bar_acc = bars + 16 * LT_STAGES;
  const uint32_t tmem_slot_addr = bars + 16 * LT_STAGES + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + LT_STAGES * STAGE_BYTES + 16 * LT_STAGES + 8);
  float* sPar = reinterpret_cast<float*>(gen + LT_STAGES * STAGE_BYTES + 16 * LT_STAGES + 16);   // scale, shift, slope [NT]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool tr = a.trace != nullptr && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024;
  long long* trow = a.trace + (tr ? blockIdx.x * 8 : 0);
  if (tr) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trow[0] = (long long)gt;
    trow[1] = clock64();
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trow[7] = smid;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < LT_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 9);     // one arrival per producer warp + the expect_tx arrival of the weight TMA
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
    prefetch_tmap(&tmapW);
  }
  if (warp == 8) {
    tmem_alloc<1>(tmem_slot_addr, NT);
    tmem_relinquish<1>();
  }
  const int n0 = blockIdx.y * NT;
  // folded BatchNorm / PReLU parameters of this CTA's channels: read once (the streaming gathers evict them from L1)
  for (int i = threadIdx.x; i < NT; i += blockDim.x) {
    sPar[i] = a.scale[n0 + i];
    sPar[NT + i] = a.shift[n0 + i];
    sPar[2 * NT + i] = a.slope[n0 + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tr) trow[2] = clock64();   // prologue done

  const int M = a.F * a.Ho * a.Wo;            // < 2^31: at most LIP_CHUNK frames per launch (32-bit index maths)
  const int m0 = blockIdx.x * 128;
  const int cblocks = a.Cin / 64;
  const int nslabs = a.ks * a.ks * cblocks;

  if (warp < 8) {
    // ------------------------------------------------------------------ A producers
    // Gather mapping: thread = (16-byte chunk ch of the 128-byte channel vector, rows g, g + 32, g + 64, g + 96), so the
    // eight lanes of a row read one whole line and a warp instruction touches 4 lines (a thread-per-row mapping
    // touches 32 lines with 16 bytes each and is bound by the L1 tag rate).  The epilogue below is thread = row.
    const int r = threadIdx.x & 127, hsel = threadIdx.x >> 7;
    const int m = m0 + r;
    const bool row_ok = m < M;
    const int ch = threadIdx.x & 7, g = threadIdx.x >> 3;
    int gy[4], gx[4];          // stride * oy - pad, stride * ox - pad of the thread's four rows (gy = -2^20: row beyond M)
    long long gbase[4];        // frame offset (elements)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int mk = m0 + g + 32 * k;
      if (mk < M) {
        const int fr = mk / (a.Wo * a.Ho), rem = mk - fr * (a.Wo * a.Ho);
        const int oy = rem / a.Wo, ox = rem - oy * a.Wo;
        gy[k] = oy * a.stride - a.pad;
        gx[k] = ox * a.stride - a.pad;
        gbase[k] = (long long)fr * a.H * a.W * a.Cin + ch * 8;
      } else {
        gy[k] = -(1 << 20); gx[k] = 0; gbase[k] = 0;
      }
    }
    // Three slabs of loads are in flight per thread (registers) while earlier slabs are written to the ring: the
    // global / L2 latency of a gather (~1 us) is several times the tensor-core time of a slab.
    auto issue = [&](int i, uint4* v) {
      if (i >= nslabs) return;
      const int tap = i / cblocks, cb = i - tap * cblocks;
      const int ky = tap / a.ks, kx = tap - ky * a.ks;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int iy = gy[k] + ky, ix = gx[k] + kx;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W && !(a.dbg & 1))
          v[k] = __ldg(reinterpret_cast<const uint4*>(a.in + gbase[k] + ((long long)iy * a.W + ix) * a.Cin + cb * 64));
        else
          v[k] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto commit = [&](int i, const uint4* v) {
      if (i >= nslabs) return;
      const int s = i % LT_STAGES, ph = (i / LT_STAGES) & 1;
      if (lane == 0) mbar_wait(bar_empty + 8 * s, ph ^ 1);   // one poller per warp
      __syncwarp();
      const uint32_t sA = base + s * STAGE_BYTES;
      if (threadIdx.x == 0) {
        if (a.dbg & 4) mbar_arrive(bar_full + 8 * s);
        else {
          mbar_expect_tx(bar_full + 8 * s, B_BYTES);
          tma_load_2d(sA + LT_A_BYTES, &tmapW, bar_full + 8 * s, i * 64, n0);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t dst = sA + sw128_offset((uint32_t)(g + 32 * k), (uint32_t)ch);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[k].x), "r"(v[k].y), "r"(v[k].z), "r"(v[k].w)
                     : "memory");
      }
      if (!(a.dbg & 128)) fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy operand reads
      __syncwarp();
      if (lane == 0) {   // one arrival per warp (256 arrivals on one mbarrier serialise)
        if (a.dbg & 64) asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8 * s) : "memory");
        else mbar_arrive(bar_full + 8 * s);
      }
    };
    if constexpr (MINB == 2) {
      uint4 v0[4], v1[4], v2[4];
      issue(0, v0);
      issue(1, v1);
      for (int i = 0; i < nslabs; i += 3) {
        issue(i + 2, v2); commit(i, v0);
        issue(i + 3, v0); commit(i + 1, v1);
        issue(i + 4, v1); commit(i + 2, v2);
      }
    } else {   // three CTAs per SM (72 registers): two slabs of loads in flight
      uint4 v0[4], v1[4];
      issue(0, v0);
      for (int i = 0; i < nslabs; i += 2) {
        issue(i + 1, v1); commit(i, v0);
        issue(i + 2, v0); commit(i + 1, v1);
      }
    }
    // ------------------------------------------------------------------ epilogue (thread = pixel row; the two warps of a
    // TMEM lane quadrant take alternate 32-column pieces)
    if (tr) trow[3] = clock64();   // all slabs written
    if (lane == 0) mbar_wait(bar_acc, 0);
    __syncwarp();
    tc_fence_after();
    if (tr) trow[4] = clock64();   // accumulator complete
#pragma unroll 1
    for (int c0 = 32 * hsel; c0 < NT; c0 += 64) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, acc);
      tmem_ld_wait();
      if (row_ok && !(a.dbg & 2)) {
        const int cg = n0 + c0;
        const float* sSc = sPar + c0;
        const float* sSh = sPar + NT + c0;
        const float* sSl = sPar + 2 * NT + c0;
        uint32_t rr[16];
        if (a.res && !(a.dbg & 32)) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + (long long)m * a.Cout + cg);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 x = __ldg(rp + q);
            rr[4 * q] = x.x; rr[4 * q + 1] = x.y; rr[4 * q + 2] = x.z; rr[4 * q + 3] = x.w;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v0f = fmaf(__uint_as_float(acc[2 * j]), sSc[2 * j], sSh[2 * j]);
          float v1f = fmaf(__uint_as_float(acc[2 * j + 1]), sSc[2 * j + 1], sSh[2 * j + 1]);
          if (a.res && !(a.dbg & 32)) {
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rr[j]));
            v0f += rf.x; v1f += rf.y;
          }
          const float s0 = sSl[2 * j], s1 = sSl[2 * j + 1];
          const int actx = (a.dbg & 16) ? LIP_ACT_NONE : a.act;
          const __half2 h = __floats2half2_rn(lip_act_tc(v0f, actx, s0), lip_act_tc(v1f, actx, s1));
          pk[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* op = reinterpret_cast<uint4*>(a.out + (long long)m * a.Cout + cg);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (!(a.dbg & 8) || pk[4 * q] == 0x12345678u) op[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(128, NT, 0);
      for (int i = 0; i < nslabs; ++i) {
        const int s = i % LT_STAGES, ph = (i / LT_STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const uint32_t sA = base + s * STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t a_desc = smem_desc_sw128_kmajor(sA) + (uint64_t)(kk * 2);
          const uint64_t b_desc = smem_desc_sw128_kmajor(sA + LT_A_BYTES) + (uint64_t)(kk * 2);
          umma_f16<1>(tmem, a_desc, b_desc, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty + 8 * s);
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
  }
  if (tr) trow[5] = clock64();     // epilogue of warp 0 done
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<1>(tmem, NT);
  if (tr) trow[6] = clock64();
}

// ------------------------------------------------------------------------------------------
// Persistent version: one CTA per SM walks tiles (m tile, n tile); every role runs ahead across tile boundaries.
//   warps 0..7   A producers in four groups of two warps; group g owns ring stage g and the K slabs j = g (mod 4) of the
//                CTA's slab sequence (thread = 16-byte chunk ch of rows r, r + 8, ...): the stage turn-round of one
//                group (loads -> st.shared -> fence.proxy.async -> arrival -> MMA -> commit) overlaps the other three
//   warp 8       MMA issuer, accumulators double-buffered in TMEM (2 x NT columns)
//   warp 9       weight slabs by TMA into their own 8-stage ring (a slab ahead of the A ring by up to 8)
//   warps 10..13 epilogue (TMEM lane quadrant = warp & 3): runs under the next tile's main loop
// The one-tile-per-CTA kernel above paid prologue, pipeline fill, epilogue and teardown per tile (clock64 trace of a
// 9-slab tile: 13.4k cycles of slabs, 9k of epilogue, 3k of prologue / teardown; profiles/r02_lipreader_*).
// ------------------------------------------------------------------------------------------
constexpr int LP_SA = 4, LP_SB = 8;

template <int NT>
__global__ void __launch_bounds__(448, 1) k_lip_conv_tc2(const __grid_constant__ CUtensorMap tmapW, const LipTcArgs a) {
  constexpr int B_BYTES = NT * 128;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA0 = base, sB0 = base + LP_SA * LT_A_BYTES;
  constexpr int OFF_BAR = LP_SA * LT_A_BYTES + LP_SB * B_BYTES;
  const uint32_t bars = base + OFF_BAR;
  const uint32_t bar_fullA = bars, bar_emptyA = bars + 32, bar_fullB = bars + 64, bar_emptyB = bars + 128,
                 bar_accfull = bars + 192, bar_accempty = bars + 208, tmem_slot_addr = bars + 224;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + 224);
  float* sPar = reinterpret_cast<float*>(gen + OFF_BAR + 256);   // scale, shift, slope [Cout]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LP_SA; ++s) { mbar_init(bar_fullA + 8 * s, 2); mbar_init(bar_emptyA + 8 * s, 1); }
    for (int s = 0; s < LP_SB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, 4); }
    fence_mbar_init();
    prefetch_tmap(&tmapW);
  }
  if (warp == 8) {
    tmem_alloc<1>(tmem_slot_addr, 2 * NT);
    tmem_relinquish<1>();
  }
  for (int i = threadIdx.x; i < a.Cout; i += blockDim.x) {
    sPar[i] = a.scale[i];
    sPar[a.Cout + i] = a.shift[i];
    sPar[2 * a.Cout + i] = a.slope[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int M = a.F * a.Ho * a.Wo;
  const int mtiles = (M + 127) / 128, ntiles = a.Cout / NT, total = mtiles * ntiles;
  const int cblocks = a.Cin / 64;
  const int nslabs = a.ks * a.ks * cblocks;
  const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // grid <= total
  const long long nj = (long long)my_tiles * nslabs;   // K slabs this CTA streams, numbered across its tiles

  if (warp < 8) {
    // ------------------------------------------------------------------ A producers
    const int gi = warp >> 1, t64 = threadIdx.x & 63, ch = t64 & 7, g8 = t64 >> 3;
    const uint4* in4 = reinterpret_cast<const uint4*>(a.in);
    const int cin8 = a.Cin / 8;
    int gyx[16], gb[16];   // (stride oy - pad) << 16 | (stride ox - pad) & 0xffff; frame offset in 16-byte units + ch
    int cur_tk = -1;
    uint4 v[16];
    auto issue = [&](long long j) {
      const int tk = (int)(j / nslabs), i = (int)(j - (long long)tk * nslabs);
      if (tk != cur_tk) {
        cur_tk = tk;
        const int mt = ((int)blockIdx.x + tk * (int)gridDim.x) / ntiles;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int mk = mt * 128 + g8 + 8 * k;
          if (mk < M) {
            const int fr = mk / (a.Wo * a.Ho), rem = mk - fr * (a.Wo * a.Ho);
            const int oy = rem / a.Wo, ox = rem - oy * a.Wo;
            gyx[k] = ((oy * a.stride - a.pad) << 16) | ((ox * a.stride - a.pad) & 0xffff);
            gb[k] = fr * a.H * a.W * cin8 + ch;
          } else {
            gyx[k] = (int)0x80000000;   // gy = -32768: never inside the image
            gb[k] = 0;
          }
        }
      }
      const int tap = i / cblocks, cb = i - tap * cblocks;
      const int ky = tap / a.ks, kx = tap - ky * a.ks;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int iy = (gyx[k] >> 16) + ky, ix = (int)(short)(gyx[k] & 0xffff) + kx;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
          v[k] = __ldg(in4 + gb[k] + (iy * a.W + ix) * cin8 + cb * 8);
        else
          v[k] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    long long j = gi;
    if (j < nj) issue(j);
    const uint32_t sA = sA0 + gi * LT_A_BYTES;
    for (; j < nj; j += LP_SA) {
      const int ph = (int)((j >> 2) & 1);
      if (lane == 0) mbar_wait(bar_emptyA + 8 * gi, ph ^ 1);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t dst = sA + sw128_offset((uint32_t)(g8 + 8 * k), (uint32_t)ch);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[k].x), "r"(v[k].y), "r"(v[k].z), "r"(v[k].w)
                     : "memory");
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fullA + 8 * gi);
      if (j + LP_SA < nj) issue(j + LP_SA);   // in flight during the stage's turn-round
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(128, NT, 0);
      long long j = 0;
      for (int tk = 0; tk < my_tiles; ++tk) {
        const int slot = tk & 1;
        mbar_wait(bar_accempty + 8 * slot, ((tk >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int i = 0; i < nslabs; ++i, ++j) {
          const int sa = (int)(j & 3), sb = (int)(j & 7);
          mbar_wait(bar_fullA + 8 * sa, (int)((j >> 2) & 1));
          mbar_wait(bar_fullB + 8 * sb, (int)((j >> 3) & 1));
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t a_desc = smem_desc_sw128_kmajor(sA0 + sa * LT_A_BYTES) + (uint64_t)(kk * 2);
            const uint64_t b_desc = smem_desc_sw128_kmajor(sB0 + sb * B_BYTES) + (uint64_t)(kk * 2);
            umma_f16<1>(tmem + slot * NT, a_desc, b_desc, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(bar_emptyA + 8 * sa);
          umma_commit(bar_emptyB + 8 * sb);
        }
        umma_commit(bar_accfull + 8 * slot);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ weight slabs (TMA)
    if (lane == 0) {
      long long j = 0;
      for (int tk = 0; tk < my_tiles; ++tk) {
        const int nt = ((int)blockIdx.x + tk * (int)gridDim.x) % ntiles;
        for (int i = 0; i < nslabs; ++i, ++j) {
          const int sb = (int)(j & 7);
          mbar_wait(bar_emptyB + 8 * sb, (int)((j >> 3) & 1) ^ 1);
          mbar_expect_tx(bar_fullB + 8 * sb, B_BYTES);
          tma_load_2d(sB0 + sb * B_BYTES, &tmapW, bar_fullB + 8 * sb, i * 64, nt * NT);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (thread = pixel row)
    const int q = warp & 3;
    for (int tk = 0; tk < my_tiles; ++tk) {
      const int tile = (int)blockIdx.x + tk * (int)gridDim.x;
      const int mt = tile / ntiles, n0 = (tile - mt * ntiles) * NT, slot = tk & 1;
      const int m = mt * 128 + q * 32 + lane;
      const bool row_ok = m < M;
      if (lane == 0) mbar_wait(bar_accfull + 8 * slot, (tk >> 1) & 1);
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + slot * NT + c0, acc);
        tmem_ld_wait();
        if (c0 + 32 >= NT) {   // the accumulator slot is free for the tile after the next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_accempty + 8 * slot);
        }
        if (row_ok && !(a.dbg & 2)) {
          const int cg = n0 + c0;
          const float* sSc = sPar + cg;
          const float* sSh = sPar + a.Cout + cg;
          const float* sSl = sPar + 2 * a.Cout + cg;
          uint32_t rr[16];
          if (a.res) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.res + (long long)m * a.Cout + cg);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 x = __ldg(rp + u);
              rr[4 * u] = x.x; rr[4 * u + 1] = x.y; rr[4 * u + 2] = x.z; rr[4 * u + 3] = x.w;
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float v0f = fmaf(__uint_as_float(acc[2 * u]), sSc[2 * u], sSh[2 * u]);
            float v1f = fmaf(__uint_as_float(acc[2 * u + 1]), sSc[2 * u + 1], sSh[2 * u + 1]);
            if (a.res) {
              const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rr[u]));
              v0f += rf.x; v1f += rf.y;
            }
            const __half2 h = __floats2half2_rn(lip_act_tc(v0f, a.act, sSl[2 * u]), lip_act_tc(v1f, a.act, sSl[2 * u + 1]));
            pk[u] = *reinterpret_cast<const uint32_t*>(&h);
          }
          uint4* op = reinterpret_cast<uint4*>(a.out + (long long)m * a.Cout + cg);
#pragma unroll
          for (int u = 0; u < 4; ++u) op[u] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<1>(tmem, 2 * NT);
}

int g_lip_tc_version = 1;   // 1 = one tile per CTA, two CTAs per SM (default: faster), 2 = persistent kernel

template <int NT>
static int lip_conv_tc2_launch(const char* packed, const LipConv& c, const LipTcArgs& a, cudaStream_t st) {
  constexpr int SMEM = LP_SA * LT_A_BYTES + LP_SB * NT * 128 + 256 + 3 * 512 * 4 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory");
  static PerDeviceOnce configured;
  if (configured.first())
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_lip_conv_tc2<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  VATSS_CHECK_ARG(c.cout <= 512 && a.H < 16384 && a.W < 16384, "lipreader tensor engine: shape");
  CUtensorMap tmapW;
  const uint64_t K = (uint64_t)c.taps * c.cin;
  const uint64_t dims[2] = {K, (uint64_t)c.cout};
  const uint64_t strides[1] = {K * sizeof(__half)};
  const uint32_t box[2] = {64, (uint32_t)NT};
  if (int rc = make_tmap_f16(&tmapW, packed + c.off_w16, 2, dims, strides, box)) return rc;
  const long long M = (long long)a.F * a.Ho * a.Wo;
  VATSS_CHECK_ARG(M < (1ll << 31) - 256 && (long long)a.F * a.H * a.W * (a.Cin / 8) < (1ll << 31),
                  "lipreader tensor engine: %lld output pixels in one launch", M);
  const long long total = (long long)ceil_div(M, 128) * (c.cout / NT);
  const int grid = (int)(total < num_sms() ? total : num_sms());
  k_lip_conv_tc2<NT><<<grid, 448, SMEM, st>>>(tmapW, a);
  VATSS_LAUNCH_OK();
  return 0;
}

template <int NT, int MINB>
static int lip_conv_tc_launch(const char* packed, const LipConv& c, const LipTcArgs& a, cudaStream_t st) {
  constexpr int SMEM = LT_STAGES * (LT_A_BYTES + NT * 128) + 16 * LT_STAGES + 16 + 3 * NT * 4 + 1024;
  static PerDeviceOnce configured;
  if (configured.first())
    VATSS_CUDA_OK(cudaFuncSetAttribute(k_lip_conv_tc<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  CUtensorMap tmapW;
  const uint64_t K = (uint64_t)c.taps * c.cin;
  const uint64_t dims[2] = {K, (uint64_t)c.cout};
  const uint64_t strides[1] = {K * sizeof(__half)};
  const uint32_t box[2] = {64, (uint32_t)NT};
  if (int rc = make_tmap_f16(&tmapW, packed + c.off_w16, 2, dims, strides, box)) return rc;
  const long long M = (long long)a.F * a.Ho * a.Wo;
  VATSS_CHECK_ARG(M < (1ll << 31) - 256, "lipreader tensor engine: %lld output pixels in one launch", M);
  dim3 grid(ceil_div(M, 128), c.cout / NT);
  k_lip_conv_tc<NT, MINB><<<grid, 288, SMEM, st>>>(tmapW, a);
  VATSS_LAUNCH_OK();
  return 0;
}

int lip_conv_tc(const char* packed, const LipConv& c, const __half* in16, int F, int H, int W, int Ho, int Wo,
                const __half* res16, int act, __half* out16, cudaStream_t st) {
  VATSS_CHECK_ARG(c.cin % 64 == 0 && c.cout % 64 == 0, "lipreader tensor engine: channels %d -> %d must be multiples of 64",
                  c.cin, c.cout);
  LipTcArgs a;
  a.in = in16; a.F = F; a.H = H; a.W = W; a.Cin = c.cin; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout;
  a.ks = c.ks; a.stride = c.stride; a.pad = c.pad;
  a.scale = (const float*)(packed + c.off_scale); a.shift = (const float*)(packed + c.off_shift);
  a.slope = (const float*)(packed + c.off_slope);
  a.res = res16; a.act = act; a.out = out16; a.dbg = g_lip_dbg; a.trace = g_lip_trace;
  if (g_lip_tc_version == 2 && !g_lip_trace) {
    if (c.cout % 128 == 0) return lip_conv_tc2_launch<128>(packed, c, a, st);
    return lip_conv_tc2_launch<64>(packed, c, a, st);
  }
  if (c.cout % 128 == 0) return lip_conv_tc_launch<128, 2>(packed, c, a, st);
  // 64-channel layers (front-end GEMM, layer 1): three CTAs per SM (72 registers, two slabs of loads in flight) hide more
  // of the per-tile prologue / epilogue than two with three slabs in flight: 9.01 -> 8.47 ms per 3200 frames, bit-identical
  if (g_lip_dbg & 256) return lip_conv_tc_launch<64, 2>(packed, c, a, st);   // cross-check: the two-CTA variant
  return lip_conv_tc_launch<64, 3>(packed, c, a, st);
}

}  // namespace vatss

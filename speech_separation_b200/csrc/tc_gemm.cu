// TENSOR engine: per-token projections on tcgen05 tensor cores.
//
//   out[M, NOUT] = A[M, KDIM] (fp16, TMA, SWIZZLE_128B) x W[NOUT, KDIM]^T (fp16, smem-resident) + epilogue
//
// Persistent kernel, one CTA per SM, warp-specialised:
//   warp 0      TMA producer   (W once, then a 2-stage ring of 128-row A tiles)
//   warp 1      MMA issuer     (one thread; tcgen05.mma kind::f16, M=128, N<=256, fp32 accum in TMEM)
//   warps 2..5  epilogue       (tcgen05.ld, thread = output row; all global traffic is staged through a
//                               per-warp shared-memory tile so that every warp instruction touches whole lines)
// K is only 64..256 here, so a whole K extent of A and all of W fit in shared memory: no K pipeline,
// W is read from HBM/L2 once per CTA and A exactly once per forward.  These GEMMs are HBM-bound
// (AI ~ 96 flop/B < ridge ~ 255), hence everything that can be folded into the epilogue is.
//
// Used for: QKV in-projection, attention out-projection (+residual+LayerNorm), FFN (+residual+LN),
// DPRNN fc (+LN, +residual), speaker split, post-conv / mask heads
// (src/model/dptn.py:46-51, dprnn.py:39-45, dptn_wav.py:47,59).
#include "common.cuh"
#include "ptx.cuh"
#include "tc_kernels.cuh"

namespace vatss {

using namespace ptx;

// ------------------------------------------------------------------------------------------
// host: tensor-map construction through the driver entry point
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  VATSS_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VATSS_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rank %d)", (int)r, rank);
  return 0;
}

int g_cta_limit = 0;   // experiment knob: cap on persistent-kernel grids (0 = all SMs)
int grid_cap() { return (g_cta_limit > 0 && g_cta_limit < num_sms()) ? g_cta_limit : num_sms(); }

int num_sms() {
  static int cache[64] = {0};   // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[dev & 63];
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct TcGemmArgs {
  long long M;
  int num_tiles;
  const float* bias;
  const float* res;
  long long ldr;
  const __half* res16;   // fp16 residual (RES16 instances): the same tensor the next operand is read from
  const __half* res16lo; // RES16 = 2: second fp16 tensor, residual = res16 + res16lo (hi / lo split of an fp32 stream)
  long long ldr16;
  const float* ln_w;
  const float* ln_b;
  float* out32;
  long long ldo32;
  __half* out16;
  long long ldo16;
  __half* out16lo;   // optional: half(value - half(value)) (hi/lo split for the PRECISE LSTM), same pitch as out16
  int act16;  // activation applied to the fp16 copy only: 0 none, 1 relu, 2 prelu
  const float* prelu_a;
  int reverse;   // walk the tiles from the last to the first (L2 reuse of the producer's most recent output)
};

// TMA warp, MMA warp and 8 epilogue warps: two warps per TMEM lane quadrant split the columns of the 128-row tile
// (with one warp per scheduler every tcgen05.ld / shared / global latency of the store path was exposed).  Plain
// epilogues take alternate 32-column pieces; LayerNorm epilogues take half a row each and exchange the partial
// mean / variance sums through shared memory.  Falls back to 4 warps (thread = whole row) where the extra staging
// tiles do not fit next to W, A and the residual tile (K = 256 FFN).
constexpr int tcg_epilogue_warps(int fixed_bytes) { return fixed_bytes + 8 * 32 * 36 * 4 + 1024 + 2048 <= 227 * 1024 ? 8 : 4; }

// WSPLIT: W is stored as [half(W) | half(W - half(W))] (hi/lo split, 2 x KDIM columns) and both halves are
// contracted with the same A tile - used where an fp16-rounded weight misses the tolerance (DPRNN fc).
// RES16: 0 = fp32 residual, 1 = fp16 residual, 2 = fp16 hi + lo pair (the fp32 residual stream stored as
// half(x) + half(x - half(x)): same bytes to read as fp32, but the producer writes 4 instead of 6 bytes per element
// because the hi half IS the fp16 operand copy the next projection reads)
template <int NOUT, int KDIM, bool RES_TMA = false, bool WSPLIT = false, int RES16 = 0>
struct TcGemmSmem {
  static constexpr int KB = KDIM / 64;
  static constexpr int KBW = WSPLIT ? 2 * KB : KB;
  static constexpr int W_BYTES = KBW * NOUT * 128;
  static constexpr int A_STAGE_BYTES = KB * 128 * 128;
  // residual tile [128 rows x NOUT fp32] as NOUT/32 column blocks of 128-byte rows (SWIZZLE_128B), TMA-prefetched
  // (RES16: NOUT/64 column blocks of 64 fp16 columns instead)
  static constexpr int RES_BYTES = RES_TMA ? 128 * NOUT * (RES16 == 1 ? 2 : 4) : 0;
  static constexpr int A_STAGES = (W_BYTES + 2 * A_STAGE_BYTES + RES_BYTES + 24 * 1024 <= 227 * 1024) ? 2 : 1;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = OFF_W + W_BYTES;
  static constexpr int OFF_RES = OFF_A + A_STAGES * A_STAGE_BYTES;
  static constexpr int OFF_PAR = OFF_RES + RES_BYTES;  // bias, ln_w, ln_b
  static constexpr int OFF_BAR = OFF_PAR + 3 * NOUT * 4;
  static constexpr int OFF_XCH = OFF_BAR + 128;          // LayerNorm partial sums: [2 halves][128 rows] x (sum, sq)
  static constexpr int OFF_STAGE = OFF_XCH + 2048;       // per epilogue warp: [32 rows][36 floats] transpose tile
  static constexpr int EPW = tcg_epilogue_warps(OFF_STAGE);
  static constexpr int THREADS = 64 + 32 * EPW;
  static constexpr int TOTAL = OFF_STAGE + EPW * 32 * 36 * 4 + 1024;  // + alignment slack
  // Accumulators live in a ring of TMEM slots of CH columns.  A tile wider than 256 columns (QKV: 384) is issued
  // as NCH chunks of 128 columns, each with its own full/empty barrier, so the MMAs of the next chunk / tile overlap
  // the epilogue of the current one instead of waiting for the whole 384-column tile to drain.
  static constexpr int NCH = (NOUT > 256) ? NOUT / 128 : 1;
  static constexpr int CH = NOUT / NCH;
  static constexpr int ACC_SLOTS = (4 * CH <= 512 && NCH > 1) ? 4 : ((2 * CH <= 512) ? 2 : 1);
  static constexpr int TMEM_COLS = (ACC_SLOTS * CH <= 32)    ? 32
                                   : (ACC_SLOTS * CH <= 64)  ? 64
                                   : (ACC_SLOTS * CH <= 128) ? 128
                                   : (ACC_SLOTS * CH <= 256) ? 256
                                                             : 512;
  static_assert(NOUT % NCH == 0 && CH <= 256, "chunking");
};

// ---- warp-level staged global access: a warp owns 32 consecutive rows; thread = row in registers, but every
// global instruction is row-contiguous (4 rows x 128 B for fp32 chunks of 32 columns, 8 rows x 64 B for fp16).
constexpr int STG_LD = 36;   // padded row pitch (floats) of the staging tile: conflict-free for both access patterns

// v[32] += / = g[row lane][0..31]
template <bool ACCUM>
__device__ __forceinline__ void staged_load_f32(const float* __restrict__ g, long long ld, int rows_valid, float* stage,
                                                int lane, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);   // 8 lanes x 16 B = one 128-byte row segment, 4 rows per instruction
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) x = *reinterpret_cast<const float4*>(g + (long long)r * ld + (lane & 7) * 4);
    *reinterpret_cast<float4*>(stage + r * STG_LD + (lane & 7) * 4) = x;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 x = *reinterpret_cast<const float4*>(stage + lane * STG_LD + 4 * j);
    if (ACCUM) { v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w; }
    else { v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w; }
  }
  __syncwarp();
}
// g[row lane][0..31] = v[32]
__device__ __forceinline__ void staged_store_f32(float* __restrict__ g, long long ld, int rows_valid, float* stage, int lane,
                                                 const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * STG_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3);
    const float4 x = *reinterpret_cast<const float4*>(stage + r * STG_LD + (lane & 7) * 4);
    if (r < rows_valid) *reinterpret_cast<float4*>(g + (long long)r * ld + (lane & 7) * 4) = x;
  }
  __syncwarp();
}
// g16[row lane][0..31] = half(pk) where pk[16] holds the row's 32 values as packed half2 (64 B per row).
// Staging tile: 32 rows x 64 B, unpadded, 16-byte chunk j of row r stored at chunk j ^ ((r >> 1) & 3): both the
// row-per-lane writes and the 4-lanes-per-row reads are bank-conflict free (a padded pitch of 80 B made every read
// a 2-way conflict; the QKV kernel sat at 93 % of the L1/shared-memory pipe).
__device__ __forceinline__ void staged_store_f16(__half* __restrict__ g, long long ld, int rows_valid, float* stage, int lane,
                                                 const uint32_t* pk) {
  uint32_t* st = reinterpret_cast<uint32_t*>(stage);
  const int wsw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(st + lane * 16 + ((j ^ wsw) << 2)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), c = lane & 3;
    const uint4 x = *reinterpret_cast<const uint4*>(st + r * 16 + ((c ^ ((r >> 1) & 3)) << 2));
    if (r < rows_valid) *reinterpret_cast<uint4*>(g + (long long)r * ld + c * 8) = x;
  }
  __syncwarp();
}

// a += half(lo 16 bits of h2), b += half(hi 16 bits): sm_100a mixed-precision adds (SASS: FHADD R, R.H0 / R.H1, R)
__device__ __forceinline__ void add_half2(float& a, float& b, uint32_t h2) {
  asm("{\n\t.reg .b16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "add.rn.f32.f16 %0, l, %0;\n\t"
      "add.rn.f32.f16 %1, h, %1;\n\t}"
      : "+f"(a), "+f"(b)
      : "r"(h2));
}

template <int NOUT, int KDIM, int EPI, bool WSPLIT, int RES16>
__global__ void __launch_bounds__((TcGemmSmem<NOUT, KDIM, (EPI == TC_EPI_LN || EPI == TC_EPI_LN_POST), WSPLIT, RES16>::THREADS), 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapW,
          const __grid_constant__ CUtensorMap tmapR, const __grid_constant__ CUtensorMap tmapR2, TcGemmArgs p) {
  constexpr bool RES_TMA = (EPI == TC_EPI_LN || EPI == TC_EPI_LN_POST);   // residual tile prefetched by TMA
  using L = TcGemmSmem<NOUT, KDIM, RES_TMA, WSPLIT, RES16>;
  static_assert(NOUT % 16 == 0 && NOUT <= 512 && KDIM % 64 == 0, "shape");
  static_assert(!RES16 || (RES_TMA && NOUT % 64 == 0), "fp16 residual: LayerNorm epilogues, 64-column blocks");
  static_assert(EPI != TC_EPI_LN && EPI != TC_EPI_LN_POST || NOUT <= 128, "LayerNorm epilogue needs the row in regs");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sW = base + L::OFF_W;
  const uint32_t sA = base + L::OFF_A;
  float* sPar = reinterpret_cast<float*>(gen + L::OFF_PAR);
  const uint32_t bars = base + L::OFF_BAR;
  const uint32_t bar_w = bars, bar_afull = bars + 8, bar_aempty = bars + 24, bar_accfull = bars + 40,
                 bar_accempty = bars + 72, bar_rfull = bars + 104, bar_rempty = bars + 112;
  const uint32_t sR = base + L::OFF_RES;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + L::OFF_BAR + 120);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar_accfull + 8 * s, 1);
      mbar_init(bar_accempty + 8 * s, 32 * L::EPW);
    }
    mbar_init(bar_rfull, 1);
    mbar_init(bar_rempty, 32 * L::EPW);
    fence_mbar_init();
    prefetch_tmap(&tmapA);
    prefetch_tmap(&tmapW);
    if (RES_TMA) prefetch_tmap(&tmapR);
  }
  for (int i = threadIdx.x; i < NOUT; i += blockDim.x) {
    sPar[i] = p.bias ? p.bias[i] : 0.f;
    sPar[NOUT + i] = p.ln_w ? p.ln_w[i] : 1.f;
    sPar[2 * NOUT + i] = p.ln_b ? p.ln_b[i] : 0.f;
  }
  if (warp == 1) {
    tmem_alloc<1>(bars + 120, L::TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(bar_w, L::W_BYTES);
      for (int kb = 0; kb < L::KBW; ++kb)
        for (int n0 = 0; n0 < NOUT; n0 += 64) tma_load_2d(sW + kb * NOUT * 128 + n0 * 128, &tmapW, bar_w, kb * 64, n0);
      int i = 0;
      for (int tile_i = blockIdx.x; tile_i < p.num_tiles; tile_i += gridDim.x, ++i) {
        const int tile = p.reverse ? p.num_tiles - 1 - tile_i : tile_i;
        const int s = i % L::A_STAGES, ph = (i / L::A_STAGES) & 1;
        mbar_wait(bar_aempty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_afull + 8 * s, L::A_STAGE_BYTES);
        for (int kb = 0; kb < L::KB; ++kb)
          tma_load_2d(sA + s * L::A_STAGE_BYTES + kb * 16384, &tmapA, bar_afull + 8 * s, kb * 64, tile * 128);
        if constexpr (RES_TMA) {
          mbar_wait(bar_rempty, (i & 1) ^ 1);
          mbar_expect_tx(bar_rfull, L::RES_BYTES);
          constexpr int RCOLS = RES16 ? 64 : 32;   // columns per 128-byte-row block
          for (int cbk = 0; cbk < NOUT / RCOLS; ++cbk)
            tma_load_2d(sR + cbk * 16384, &tmapR, bar_rfull, cbk * RCOLS, tile * 128);
          if constexpr (RES16 == 2)                // lo tile behind the hi tile
            for (int cbk = 0; cbk < NOUT / 64; ++cbk)
              tma_load_2d(sR + (NOUT / 64 + cbk) * 16384, &tmapR2, bar_rfull, cbk * 64, tile * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      int i = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
        const int s = i % L::A_STAGES, ph = (i / L::A_STAGES) & 1;
        mbar_wait(bar_afull + 8 * s, ph);
#pragma unroll
        for (int ch = 0; ch < L::NCH; ++ch) {
          const int jj = i * L::NCH + ch, as = jj % L::ACC_SLOTS, aph = (jj / L::ACC_SLOTS) & 1;
          const int n0 = ch * L::CH;
          const uint32_t idesc = idesc_f16(128, L::CH, 0);
          mbar_wait(bar_accempty + 8 * as, aph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int part = 0; part < (WSPLIT ? 2 : 1); ++part) {
#pragma unroll
            for (int k16 = 0; k16 < KDIM / 16; ++k16) {
              const int kb = k16 >> 2, kk = k16 & 3;
              const uint64_t a_desc = smem_desc_sw128_kmajor(sA + s * L::A_STAGE_BYTES + kb * 16384) + (uint64_t)(kk * 2);
              const uint64_t b_desc = smem_desc_sw128_kmajor(sW + (part * L::KB + kb) * NOUT * 128 + n0 * 128) + (uint64_t)(kk * 2);
              umma_f16<1>(tmem + as * L::CH, a_desc, b_desc, idesc, (part > 0 || k16 > 0) ? 1u : 0u);
            }
          }
          if (ch == L::NCH - 1) umma_commit(bar_aempty + 8 * s);
          umma_commit(bar_accfull + 8 * as);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (thread = row)
    const int q = warp & 3;
    const float* sBias = sPar;
    const float* sLw = sPar + NOUT;
    const float* sLb = sPar + 2 * NOUT;
    float* stage = reinterpret_cast<float*>(gen + L::OFF_STAGE) + (warp - 2) * 32 * STG_LD;
    const float slope = (p.act16 == 2 && p.prelu_a) ? p.prelu_a[0] : 0.f;
    int i = 0;
    for (int tile_i = blockIdx.x; tile_i < p.num_tiles; tile_i += gridDim.x, ++i) {
      const int tile = p.reverse ? p.num_tiles - 1 - tile_i : tile_i;
      const long long row0 = (long long)tile * 128 + q * 32;          // first row of this warp
      const long long left = p.M - row0;
      const int rows_valid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
      // (LayerNorm epilogues have one chunk per tile: slot / phase of chunk 0)
      const int as = (i * L::NCH) % L::ACC_SLOTS, aph = ((i * L::NCH) / L::ACC_SLOTS) & 1;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + as * L::CH;
      if constexpr (EPI == TC_EPI_F16 && L::NCH > 1 && L::EPW == 8 && L::CH == 128) {
        // Wide fp16 tile (QKV, 3 chunks of 128 columns): the two 32-column pieces a warp owns in a chunk are handled
        // together - both tcgen05.ld in flight, one hand-back of the slot, both staging tiles (2 x 2 KB of this warp's
        // staging area) written before one __syncwarp - so a warp exposes half as many TMEM / shared-memory round trips.
        const int half = (warp - 2) >> 2;
        uint32_t* st = reinterpret_cast<uint32_t*>(stage);
        const int wsw = (lane >> 1) & 3;
#pragma unroll 1
        for (int ch = 0; ch < L::NCH; ++ch) {
          const int jj = i * L::NCH + ch, cs = jj % L::ACC_SLOTS;
          const int c0 = ch * L::CH + half * 32;          // pieces [c0, c0 + 32) and [c0 + 64, c0 + 96)
          mbar_wait(bar_accfull + 8 * cs, (jj / L::ACC_SLOTS) & 1);
          tc_fence_after();
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + cs * L::CH + half * 32, r0);
          tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + cs * L::CH + half * 32 + 64, r1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(bar_accempty + 8 * cs);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 8 * j + 2 * e;
              const __half2 ha = __floats2half2_rn(__uint_as_float(r0[c]) + sBias[c0 + c], __uint_as_float(r0[c + 1]) + sBias[c0 + c + 1]);
              const __half2 hb = __floats2half2_rn(__uint_as_float(r1[c]) + sBias[c0 + 64 + c],
                                                   __uint_as_float(r1[c + 1]) + sBias[c0 + 64 + c + 1]);
              a[e] = *reinterpret_cast<const uint32_t*>(&ha);
              b[e] = *reinterpret_cast<const uint32_t*>(&hb);
            }
            *reinterpret_cast<uint4*>(st + lane * 16 + ((j ^ wsw) << 2)) = make_uint4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<uint4*>(st + 512 + lane * 16 + ((j ^ wsw) << 2)) = make_uint4(b[0], b[1], b[2], b[3]);
          }
          __syncwarp();
          __half* g = p.out16 + row0 * p.ldo16 + c0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = k * 8 + (lane >> 2), c = lane & 3;
            const uint4 x = *reinterpret_cast<const uint4*>(st + r * 16 + ((c ^ ((r >> 1) & 3)) << 2));
            const uint4 y = *reinterpret_cast<const uint4*>(st + 512 + r * 16 + ((c ^ ((r >> 1) & 3)) << 2));
            if (r < rows_valid) {
              *reinterpret_cast<uint4*>(g + (long long)r * p.ldo16 + c * 8) = x;
              *reinterpret_cast<uint4*>(g + (long long)r * p.ldo16 + 64 + c * 8) = y;
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == TC_EPI_F16 || EPI == TC_EPI_F32) {
        const int half = (warp - 2) >> 2;       // which of the (EPW / 4) warps of this lane quadrant
        int cur = -1;
#pragma unroll 1
        for (int c0 = half * 32; c0 < NOUT; c0 += 8 * L::EPW) {
          const int jj = i * L::NCH + c0 / L::CH, cs = jj % L::ACC_SLOTS;
          if (c0 / L::CH != cur) {
            cur = c0 / L::CH;
            mbar_wait(bar_accfull + 8 * cs, (jj / L::ACC_SLOTS) & 1);
            tc_fence_after();
          }
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + cs * L::CH + c0 % L::CH, r);
          tmem_ld_wait();
          if (L::NCH > 1 && c0 % L::CH == L::CH - 8 * L::EPW + half * 32) {   // this warp's last piece of the chunk is in
            tc_fence_before();                                        // registers: hand the slot back
            mbar_arrive(bar_accempty + 8 * cs);
          }
          if constexpr (EPI == TC_EPI_F16) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __half2 h = __floats2half2_rn(__uint_as_float(r[2 * j]) + sBias[c0 + 2 * j],
                                                  __uint_as_float(r[2 * j + 1]) + sBias[c0 + 2 * j + 1]);
              pk[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            staged_store_f16(p.out16 + row0 * p.ldo16 + c0, p.ldo16, rows_valid, stage, lane, pk);
          } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sBias[c0 + j];
            if (p.res) staged_load_f32<true>(p.res + row0 * p.ldr + c0, p.ldr, rows_valid, stage, lane, v);
            staged_store_f32(p.out32 + row0 * p.ldo32 + c0, p.ldo32, rows_valid, stage, lane, v);
          }
        }
      } else {
        // LayerNorm over the row: each thread holds NOUT / NH columns (NH = warps per lane quadrant); the residual tile
        // was prefetched into shared memory by TMA
        constexpr int NH = L::EPW / 4;
        constexpr int NC = NOUT / NH;
        static_assert(NC % 32 == 0, "LayerNorm epilogue: column split");
        const int half = (warp - 2) >> 2;
        const int cbase = half * NC;
        float* xch = reinterpret_cast<float*>(gen + L::OFF_XCH);
        float v[NC];
        const int rr = q * 32 + lane;
        // v[c0 .. c0+32) += residual[row rr, cbase + c0 ...] from the TMA-prefetched tile
        auto add_residual = [&](int c0) {
          if constexpr (!RES16) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 x = *reinterpret_cast<const float4*>(gen + L::OFF_RES + ((cbase + c0) / 32) * 16384 +
                                                                sw128_offset((uint32_t)rr, (uint32_t)c));
              v[c0 + 4 * c] += x.x; v[c0 + 4 * c + 1] += x.y; v[c0 + 4 * c + 2] += x.z; v[c0 + 4 * c + 3] += x.w;
            }
          } else {
            const int col = cbase + c0;
#pragma unroll
            for (int part = 0; part < (RES16 == 2 ? 2 : 1); ++part)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 x = *reinterpret_cast<const uint4*>(gen + L::OFF_RES + (part * (NOUT / 64) + col / 64) * 16384 +
                                                              sw128_offset((uint32_t)rr, (uint32_t)((col % 64) / 8 + c)));
              // mixed-precision add (add.rn.f32.f16 = one FHADD per value: no conversion instructions)
              add_half2(v[c0 + 8 * c], v[c0 + 8 * c + 1], x.x);
              add_half2(v[c0 + 8 * c + 2], v[c0 + 8 * c + 3], x.y);
              add_half2(v[c0 + 8 * c + 4], v[c0 + 8 * c + 5], x.z);
              add_half2(v[c0 + 8 * c + 6], v[c0 + 8 * c + 7], x.w);
            }
          }
        };
        mbar_wait(bar_rfull, i & 1);
        mbar_wait(bar_accfull + 8 * as, aph);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + cbase + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c0 + j] = __uint_as_float(r[j]) + sBias[cbase + c0 + j];
          if constexpr (EPI == TC_EPI_LN) add_residual(c0);
        }
        // accumulator and (pre-LayerNorm) residual are in registers: hand both buffers back now, so the next tile's
        // MMAs and residual prefetch run under this tile's LayerNorm arithmetic and stores
        tc_fence_before();
        mbar_arrive(bar_accempty + 8 * as);
        if constexpr (EPI == TC_EPI_LN) {
          // The residual tile is written by TMA (async proxy) and read here with ordinary shared loads.  The proxy
          // fence orders those reads before the next TMA write that this arrival allows; without it the stress test
          // (tools/stress_outproj.py) sees rows of one tile normalised with the next tile's residual.
          fence_proxy_async();
          mbar_arrive(bar_rempty);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) sum += v[j];
        if constexpr (NH == 2) {   // the two warps of a lane quadrant meet at named barrier 1 + q
          xch[half * 256 + rr] = sum;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          sum += xch[(half ^ 1) * 256 + rr];
        }
        const float mean = sum * (1.f / NOUT);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const float d = v[j] - mean;
          sq = fmaf(d, d, sq);
        }
        if constexpr (NH == 2) {
          xch[half * 256 + 128 + rr] = sq;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          sq += xch[(half ^ 1) * 256 + 128 + rr];
        }
        const float rstd = rsqrtf(sq * (1.f / NOUT) + 1e-5f);
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = (v[j] - mean) * rstd * sLw[cbase + j] + sLb[cbase + j];
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 32) {
          if constexpr (EPI == TC_EPI_LN_POST) add_residual(c0);
          if (p.out32) staged_store_f32(p.out32 + row0 * p.ldo32 + cbase + c0, p.ldo32, rows_valid, stage, lane, v + c0);
          if (p.out16) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a = v[c0 + 2 * j], b = v[c0 + 2 * j + 1];
              if (p.act16 == 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              else if (p.act16 == 2) { a = a >= 0.f ? a : slope * a; b = b >= 0.f ? b : slope * b; }
              const __half2 h = __floats2half2_rn(a, b);
              pk[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            staged_store_f16(p.out16 + row0 * p.ldo16 + cbase + c0, p.ldo16, rows_valid, stage, lane, pk);
            if (p.out16lo) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const __half2 hi = *reinterpret_cast<const __half2*>(&pk[j]);
                const float2 hf = __half22float2(hi);
                float a = v[c0 + 2 * j], b = v[c0 + 2 * j + 1];
                if (p.act16 == 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                else if (p.act16 == 2) { a = a >= 0.f ? a : slope * a; b = b >= 0.f ? b : slope * b; }
                const __half2 lo = __floats2half2_rn(a - hf.x, b - hf.y);
                pk[j] = *reinterpret_cast<const uint32_t*>(&lo);
              }
              staged_store_f16(p.out16lo + row0 * p.ldo16 + cbase + c0, p.ldo16, rows_valid, stage, lane, pk);
            }
          }
        }
      }
      if constexpr (EPI == TC_EPI_LN_POST) {   // (post-LayerNorm residual: read during the stores)
        fence_proxy_async();
        mbar_arrive(bar_rempty);
      }
      if constexpr (L::NCH == 1 && !RES_TMA) {
        tc_fence_before();
        mbar_arrive(bar_accempty + 8 * as);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, L::TMEM_COLS);
}

template <int NOUT, int KDIM, int EPI, bool WSPLIT = false, int RES16 = 0>
static int tc_gemm_launch(const __half* A, long long lda, const __half* W, const TcGemmArgs& args, cudaStream_t st) {
  constexpr bool RES_TMA = (EPI == TC_EPI_LN || EPI == TC_EPI_LN_POST);
  using L = TcGemmSmem<NOUT, KDIM, RES_TMA, WSPLIT, RES16>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  CUtensorMap tmA, tmW, tmR, tmR2;
  tmR = CUtensorMap();
  tmR2 = CUtensorMap();
  if (RES_TMA && RES16) {
    const uint64_t dims[2] = {(uint64_t)NOUT, (uint64_t)args.M};
    const uint64_t str[1] = {(uint64_t)args.ldr16 * 2};
    const uint32_t box[2] = {64, 128};
    if (make_tmap_f16(&tmR, args.res16, 2, dims, str, box)) return -1;
    if (RES16 == 2 && make_tmap_f16(&tmR2, args.res16lo, 2, dims, str, box)) return -1;
  } else if (RES_TMA) {
    const uint64_t dims[2] = {(uint64_t)NOUT, (uint64_t)args.M};
    const uint64_t str[1] = {(uint64_t)args.ldr * 4};
    const uint32_t box[2] = {32, 128};
    if (make_tmap_f32(&tmR, args.res, 2, dims, str, box)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)KDIM, (uint64_t)args.M};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {64, 128};
    if (make_tmap_f16(&tmA, A, 2, dims, str, box)) return -1;
  }
  {
    const uint64_t kw = WSPLIT ? 2 * KDIM : KDIM;
    const uint64_t dims[2] = {kw, (uint64_t)NOUT};
    const uint64_t str[1] = {kw * 2};
    const uint32_t box[2] = {64, 64};
    if (make_tmap_f16(&tmW, W, 2, dims, str, box)) return -1;
  }
  auto kern = k_tc_gemm<NOUT, KDIM, EPI, WSPLIT, RES16>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VATSS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  const int grid = args.num_tiles < grid_cap() ? args.num_tiles : grid_cap();
  kern<<<grid, L::THREADS, L::TOTAL, st>>>(tmA, tmW, tmR, tmR2, args);
  VATSS_LAUNCH_OK();
  return 0;
}

int launch_tc_gemm(int epi, const __half* A, long long lda, const __half* W, const float* bias, const float* res,
                   long long ldr, const float* ln_w, const float* ln_b, float* out32, long long ldo32, __half* out16,
                   long long ldo16, int act16, const float* prelu_a, long long M, int NOUT, int KDIM,
                   cudaStream_t st, __half* out16lo, int wsplit, const __half* res16, long long ldr16, int reverse,
                   const __half* res16lo) {
  if (M == 0) return 0;
  VATSS_CHECK_ARG(((uintptr_t)A & 15) == 0 && (lda * 2) % 16 == 0, "tc_gemm: A must be 16-byte aligned with 16-byte row pitch");
  TcGemmArgs a;
  a.M = M;
  a.num_tiles = (int)((M + 127) / 128);
  a.bias = bias; a.res = res; a.ldr = ldr; a.res16 = res16; a.res16lo = res16lo; a.ldr16 = ldr16; a.ln_w = ln_w; a.ln_b = ln_b;
  a.out32 = out32; a.ldo32 = ldo32; a.out16 = out16; a.ldo16 = ldo16; a.act16 = act16; a.prelu_a = prelu_a;
  a.out16lo = out16lo;
  a.reverse = reverse;
  if (epi == TC_EPI_F16) VATSS_CHECK_ARG(out16 != nullptr, "tc_gemm: fp16 output missing");
  if (epi == TC_EPI_F32) VATSS_CHECK_ARG(out32 != nullptr, "tc_gemm: fp32 output missing");
  if (epi == TC_EPI_LN || epi == TC_EPI_LN_POST)
    VATSS_CHECK_ARG((out32 || out16) && (res || res16) && ln_w && ln_b,
                    "tc_gemm: LayerNorm epilogue needs an output, a residual and ln_w/ln_b");
  if (res16) {   // fp16 residual stream (DPTN sub-blocks)
    VATSS_CHECK_ARG(epi == TC_EPI_LN && !wsplit && ((uintptr_t)res16 & 15) == 0 && (ldr16 * 2) % 16 == 0,
                    "tc_gemm: fp16 residual needs the LayerNorm epilogue and 16-byte aligned rows");
    if (res16lo) {   // residual = res16 + res16lo (out-projection of the DPTN sub-blocks)
      VATSS_CHECK_ARG(((uintptr_t)res16lo & 15) == 0, "tc_gemm: lo residual must be 16-byte aligned");
      if (NOUT == 128 && KDIM == 128) return tc_gemm_launch<128, 128, TC_EPI_LN, false, 2>(A, lda, W, a, st);
      if (NOUT == 64 && KDIM == 64) return tc_gemm_launch<64, 64, TC_EPI_LN, false, 2>(A, lda, W, a, st);
      set_error("tc_gemm: no hi/lo-residual instantiation for NOUT=%d K=%d", NOUT, KDIM);
      return -1;
    }
#define TCG_CASE16(N_, K_) \
  if (NOUT == N_ && KDIM == K_) return tc_gemm_launch<N_, K_, TC_EPI_LN, false, 1>(A, lda, W, a, st);
    TCG_CASE16(128, 128)
    TCG_CASE16(128, 256)
    TCG_CASE16(64, 64)
    TCG_CASE16(64, 256)
    TCG_CASE16(64, 128)
#undef TCG_CASE16
    set_error("tc_gemm: no fp16-residual instantiation for NOUT=%d K=%d", NOUT, KDIM);
    return -1;
  }
  if (wsplit) {   // W = [hi | lo], 2 x KDIM columns
    if (NOUT == 64 && KDIM == 256 && epi == TC_EPI_LN_POST) return tc_gemm_launch<64, 256, TC_EPI_LN_POST, true>(A, lda, W, a, st);
    if (NOUT == 64 && KDIM == 128 && epi == TC_EPI_LN_POST) return tc_gemm_launch<64, 128, TC_EPI_LN_POST, true>(A, lda, W, a, st);
    set_error("tc_gemm: no hi/lo weight instantiation for NOUT=%d K=%d epilogue=%d", NOUT, KDIM, epi);
    return -1;
  }
#define TCG_CASE(N_, K_, E_) \
  if (NOUT == N_ && KDIM == K_ && epi == E_) return tc_gemm_launch<N_, K_, E_>(A, lda, W, a, st);
  // N = 128 models
  TCG_CASE(384, 128, TC_EPI_F16)
  TCG_CASE(128, 128, TC_EPI_LN)
  TCG_CASE(128, 256, TC_EPI_LN)
  TCG_CASE(128, 256, TC_EPI_LN_POST)
  TCG_CASE(128, 128, TC_EPI_LN_POST)
  TCG_CASE(256, 128, TC_EPI_F32)
  TCG_CASE(128, 128, TC_EPI_F32)
  // N = 64 models
  TCG_CASE(192, 64, TC_EPI_F16)
  TCG_CASE(64, 64, TC_EPI_LN)
  TCG_CASE(64, 256, TC_EPI_LN)
  TCG_CASE(64, 128, TC_EPI_LN)
  TCG_CASE(64, 256, TC_EPI_LN_POST)
  TCG_CASE(64, 128, TC_EPI_LN_POST)
  TCG_CASE(128, 64, TC_EPI_F32)
  TCG_CASE(64, 64, TC_EPI_F32)
#undef TCG_CASE
  set_error("tc_gemm: no tensor-core instantiation for NOUT=%d K=%d epilogue=%d", NOUT, KDIM, epi);
  return -1;
}

}  // namespace vatss

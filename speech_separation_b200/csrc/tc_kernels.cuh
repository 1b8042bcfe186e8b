// Launchers of the tcgen05 TENSOR-engine kernels (tc_gemm.cu, tc_lstm.cu, tc_attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace vatss {

enum { TC_EPI_F16 = 0, TC_EPI_F32 = 1, TC_EPI_LN = 2, TC_EPI_LN_POST = 3 };

int num_sms();
int grid_cap();            // num_sms() or the debug cap set through vatss_debug_cta_limit
extern int g_cta_limit;
extern int g_lstm_pingpong;
extern int g_lstm_groups;
extern int g_tail_staged;
extern int g_gemm_l2_order;
extern long long* g_lstm_trace;   // debug trace buffer of k_tc_lstm (NULL in production)

// out = A[M,K] (fp16, row pitch lda) x W[NOUT,K]^T (fp16) + bias, then
//   TC_EPI_F16     -> out16 (fp16)
//   TC_EPI_F32     -> out32 = . (+ res)
//   TC_EPI_LN      -> out32 = LN(. + res), out16 = act16(out32)       (out16 optional)
//   TC_EPI_LN_POST -> out32 = LN(.) + res, out16 = act16(out32)
int launch_tc_gemm(int epi, const __half* A, long long lda, const __half* W, const float* bias, const float* res,
                   long long ldr, const float* ln_w, const float* ln_b, float* out32, long long ldo32, __half* out16,
                   long long ldo16, int act16, const float* prelu_a, long long M, int NOUT, int KDIM, cudaStream_t st,
                   __half* out16lo = nullptr, int wsplit = 0, const __half* res16 = nullptr, long long ldr16 = 0,
                   int reverse = 0, const __half* res16lo = nullptr);
// reverse = 1: the persistent CTAs walk the row tiles from the last to the first.  A consumer that starts where its
// producer stopped finds the most recently written ~100 MB of its input still in the 126 MB L2 (FFN -> QKV reads x16,
// attention -> out-projection reads att16).

// Persistent LSTM recurrence on a CTA pair (tc_lstm.cu).  x16: token-major (B,S,C,NFEAT) fp16 activation,
// Wpack / bias_pack from launch_pack_lstm, out16: (tokens, ndir*128) fp16 (relu(h) when act=1).
// mode 0: intra-chunk sequences (time = chunk position), mode 1: inter-chunk sequences (time = chunk index).
// x16lo != NULL selects the PRECISE variant (N = 64): hi/lo fp16 splits of x and W_ih, accurate gate functions.
int launch_tc_lstm(const __half* x16, const __half* x16lo, const __half* Wpack, const float* bias_pack, __half* out16,
                   int mode, int B, int S, int C, int NFEAT, int ndir, int act, cudaStream_t st);
// Wpack: [ndir][512][(precise ? 2 : 1) * NFEAT + 128] fp16, bias_pack: [ndir][512] fp32
int launch_pack_lstm(const float* Wih, const float* Whh, const float* bih, const float* bhh, int N, int dir,
                     int precise, __half* Wpack, float* bias_pack, cudaStream_t st);

// attention core on fp16 packed qkv (tokens, 3N) -> out16 (tokens, N); q is pre-scaled by log2(e)/sqrt(hd)
// mode 0 intra / 1 inter (map must be intra_map / inter_map of (B,S,C)); force_simt selects the SIMT fallback
int launch_attention_f16(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                         int force_simt, cudaStream_t st);

// v3 of the tcgen05 attention core (P and O kept in TMEM, four lean softmax warpgroups, tc_attn3.cu); same contract,
// N % 64 == 0 and head dim 16 / 32 only
int launch_attention_v3(const __half* qkv, __half* out, SeqMap map, int mode, int B, int S, int C, int N, int heads,
                        cudaStream_t st);
extern int g_attention_version;   // 3 (default, tc_attn3.cu) or 1 (round-1 kernel, tc_attention.cu)

}  // namespace vatss

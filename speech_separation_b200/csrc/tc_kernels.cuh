// Launchers of the tcgen05 TENSOR-engine kernels (tc_gemm.cu, tc_lstm.cu, tc_attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace vatss {

enum { TC_EPI_F16 = 0, TC_EPI_F32 = 1, TC_EPI_LN = 2, TC_EPI_LN_POST = 3 };

int num_sms();

// out = A[M,K] (fp16, row pitch lda) x W[NOUT,K]^T (fp16) + bias, then
//   TC_EPI_F16     -> out16 (fp16)
//   TC_EPI_F32     -> out32 = . (+ res)
//   TC_EPI_LN      -> out32 = LN(. + res), out16 = act16(out32)       (out16 optional)
//   TC_EPI_LN_POST -> out32 = LN(.) + res, out16 = act16(out32)
int launch_tc_gemm(int epi, const __half* A, long long lda, const __half* W, const float* bias, const float* res,
                   long long ldr, const float* ln_w, const float* ln_b, float* out32, long long ldo32, __half* out16,
                   long long ldo16, int act16, const float* prelu_a, long long M, int NOUT, int KDIM, cudaStream_t st);

}  // namespace vatss

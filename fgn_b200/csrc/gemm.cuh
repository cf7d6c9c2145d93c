// gemm.cuh -- C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]), fp32 in/out, both operands K-major.
// The relation head's FC contraction (fgn_roi_head.py:272: Conv2d(2C->C, 1x1)).
#pragma once
#include "common.cuh"

namespace fgn {

// precision 0: fp32 parity (tcgen05 3xTF32 split when the shape qualifies, else fp32 SIMT)
// precision 1: bf16 operands, fp32 accumulate (tcgen05 kind::f16)
int gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
            int M, int N, int K, int precision, cudaStream_t st);

// always the SIMT fp32 kernel (exported for cross-checks)
int gemm_nt_simt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C,
                 int ldc, int M, int N, int K, cudaStream_t st);

}  // namespace fgn

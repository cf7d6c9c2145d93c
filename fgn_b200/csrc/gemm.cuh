// gemm.cuh -- C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]), fp32 in/out, both operands K-major.
// The relation head's FC contraction (fgn_roi_head.py:272: Conv2d(2C->C, 1x1)).
#pragma once
#include "common.cuh"

namespace fgn {

// precision 0: fp32 parity (tcgen05 3xTF32 error-compensated split when the shape qualifies, else
//              the fp32 SIMT kernel)
// precision 1: single-pass TF32 on tcgen05 (10-bit mantissa operands, fp32 accumulate)
// split_ws: gemm_tc_workspace_bytes(N, K) bytes of scratch for the TF32 split of B.
size_t gemm_tc_workspace_bytes(int N, int K);
int gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
            int M, int N, int K, int precision, float *split_ws, cudaStream_t st, bool presplit = false,
            const float *residual = nullptr, bool relu = false);      // epilogue: C = [relu](A B^T + bias [+ residual])
int gemm_split_weights(const float *B, int ldb, int N, int K, float *split_ws, cudaStream_t st);

// always the SIMT fp32 kernel (exported for cross-checks)
int gemm_nt_simt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C,
                 int ldc, int M, int N, int K, cudaStream_t st, const float *residual = nullptr, bool relu = false);

}  // namespace fgn

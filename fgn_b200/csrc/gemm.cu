// gemm.cu -- dispatcher for the relation head's contraction.
#include "gemm.cuh"

namespace fgn {

int gemm_nt_tc(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
               int M, int N, int K, int precision, cudaStream_t st, bool *taken);

int gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
            int M, int N, int K, int precision, cudaStream_t st)
{
    bool taken = false;
    const int rc = gemm_nt_tc(A, lda, B, ldb, bias, C, ldc, M, N, K, precision, st, &taken);
    if (rc) return rc;
    if (taken) return FGN_OK;
    if (precision != 0) {
        set_error("gemm: bf16 precision needs the tcgen05 path (M=%d N=%d K=%d not supported by it)", M, N, K);
        return FGN_ERR_UNSUPPORTED;
    }
    return gemm_nt_simt(A, lda, B, ldb, bias, C, ldc, M, N, K, st);
}

}  // namespace fgn

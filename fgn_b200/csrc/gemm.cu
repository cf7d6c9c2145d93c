// gemm.cu -- dispatcher for the relation head's contraction.
#include "gemm.cuh"
#include <stdlib.h>

namespace fgn {

int gemm_nt_tc2(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc, int M, int N, int K,
                int precision, float *split_ws, cudaStream_t st, bool presplit, const float *residual, bool relu, bool *taken);

int gemm_nt_tc_bf16(const uint16_t *A, int lda, const uint16_t *B, int ldb, const float *bias, float *C, int ldc,
                    int M, int N, int K, cudaStream_t st);

int gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
            int M, int N, int K, int precision, float *split_ws, cudaStream_t st, bool presplit,
            const float *residual, bool relu)
{
    bool taken = false;
    const char *force = getenv("FGN_GEMM_IMPL");          // "simt" forces the fp32 SIMT kernel (cross-checks)
    // A class-term contraction (M = B*N*49 rows: 49 at cfg3) on the tcgen05 kernel is ONE tile paying the whole
    // pipeline latency (TMEM allocation, 16 dependent k-blocks: ~15 us); the fp32 SIMT kernel does it in a few us, exactly.
    const bool simt_only = (force != nullptr && force[0] == 's') || (precision == 0 && M <= 512 && K <= 4096 && (force == nullptr || force[0] != 't'));
    // tcgen05 kernels of conv_tc.cu: CTA pairs (cta_group::2: half the B bytes per SM and k-block) where the shape allows,
    // else one CTA per tile
    if (!simt_only)
        if (int rc2 = gemm_nt_tc2(A, lda, B, ldb, bias, C, ldc, M, N, K, precision, split_ws, st, presplit, residual, relu, &taken)) return rc2;
    if (taken) return FGN_OK;
    if (precision != 0) {
        set_error("gemm: tf32 precision needs the tcgen05 path (M=%d N=%d K=%d not supported by it)", M, N, K);
        return FGN_ERR_UNSUPPORTED;
    }
    return gemm_nt_simt(A, lda, B, ldb, bias, C, ldc, M, N, K, st, residual, relu);
}

}  // namespace fgn

using namespace fgn;

extern "C" size_t fgn_gemm_workspace_bytes(int N, int K)
{
    return (N > 0 && K > 0) ? gemm_tc_workspace_bytes(N, K) : 0;
}

extern "C" int fgn_gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C,
                           int ldc, int M, int N, int K, int precision, void *workspace,
                           size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm dims M=%d N=%d K=%d", M, N, K);
    FGN_CHECK_ARG(precision == 0 || precision == 1, "precision=%d", precision);
    if (M == 0) return FGN_OK;
    FGN_CHECK_ARG(A && B && C, "NULL pointer");
    float *ws = workspace_bytes >= gemm_tc_workspace_bytes(N, K) ? (float *)workspace : nullptr;
    return gemm_nt(A, lda, B, ldb, bias, C, ldc, M, N, K, precision, ws, (cudaStream_t)stream);
}

// The fp32-parity contraction with the weights' TF32 split made beforehand (fgn_conv_split_weights(B, 1, N, K, ...) or a
// quarter of fgn_relation_split_weights' output): what the relation head runs per call once its weights are loaded.
extern "C" int fgn_gemm_nt_presplit(const float *A, int lda, const float *b_split, const float *bias, float *C, int ldc,
                                    int M, int N, int K, void *stream)
{
    FGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm dims M=%d N=%d K=%d", M, N, K);
    if (M == 0) return FGN_OK;
    FGN_CHECK_ARG(A && b_split && C, "NULL pointer");
    return gemm_nt(A, lda, b_split, K, bias, C, ldc, M, N, K, 0, const_cast<float *>(b_split), (cudaStream_t)stream, true);
}

extern "C" int fgn_gemm_nt_bf16(const uint16_t *A, int lda, const uint16_t *B, int ldb, const float *bias, float *C,
                                int ldc, int M, int N, int K, void *stream)
{
    FGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm dims M=%d N=%d K=%d", M, N, K);
    if (M == 0) return FGN_OK;
    FGN_CHECK_ARG(A && B && C, "NULL pointer");
    return gemm_nt_tc_bf16(A, lda, B, ldb, bias, C, ldc, M, N, K, (cudaStream_t)stream);
}

// Post-RoI head convolutions as contractions (first piece): a 1x1 convolution over NHWC RoI tiles is
// C[M = R*H*W, Cout] = X[M, Cin] W[Cout, Cin]^T; BatchNorm (eval) folded into W / bias by the caller, the bottleneck's
// identity branch and ReLU in the epilogue.  Same tcgen05 3xTF32 kernel as the relation conv.
extern "C" int fgn_conv1x1_nhwc(const float *x, const float *weight, const float *bias, const float *residual, int relu,
                                float *out, int M, int Cin, int Cout, int precision, void *workspace, size_t workspace_bytes,
                                void *stream)
{
    FGN_CHECK_ARG(precision == 0 || precision == 1, "precision=%d", precision);
    FGN_CHECK_ARG(M >= 0 && Cin > 0 && Cout > 0, "conv1x1 dims M=%d Cin=%d Cout=%d", M, Cin, Cout);
    if (M == 0) return FGN_OK;
    FGN_CHECK_ARG(x && weight && out, "NULL pointer");
    float *ws = workspace_bytes >= gemm_tc_workspace_bytes(Cout, Cin) ? (float *)workspace : nullptr;
    return gemm_nt(x, Cin, weight, Cin, bias, out, Cout, M, Cout, Cin, precision, ws, (cudaStream_t)stream, false, residual, relu != 0);
}

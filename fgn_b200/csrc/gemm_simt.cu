// gemm_simt.cu -- fp32 SIMT GEMM (exact fp32 FMA accumulation).  Fallback for shapes the tcgen05
// kernel does not take, and the in-library cross-check for it.
#include "gemm.cuh"

namespace fgn {

constexpr int BM = 128, BN = 128, BK = 16;

__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float *__restrict__ A, const int lda, const float *__restrict__ B, const int ldb,
                const float *__restrict__ bias, float *__restrict__ C, const int ldc,
                const int M, const int N, const int K, const float *__restrict__ residual, const int relu)
{
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;          // loader: 64 rows x 4 float4 per pass

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = lrow + 64 * h;
            const int k = k0 + lk;
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f); rb[h] = ra[h];
            if (m0 + row < M && k < K) ra[h] = ldg4(A + (size_t)(m0 + row) * lda + k);
            if (n0 + row < N && k < K) rb[h] = ldg4(B + (size_t)(n0 + row) * ldb + k);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = lrow + 64 * h;
            As[buf][lk + 0][row] = ra[h].x; As[buf][lk + 1][row] = ra[h].y;
            As[buf][lk + 2][row] = ra[h].z; As[buf][lk + 3][row] = ra[h].w;
            Bs[buf][lk + 0][row] = rb[h].x; Bs[buf][lk + 1][row] = rb[h].y;
            Bs[buf][lk + 2][row] = rb[h].z; Bs[buf][lk + 3][row] = rb[h].w;
        }
    };
    const int nk = (K + BK - 1) / BK;
    gload(0); sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) { sstore(buf ^ 1); }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n < N) {
                float o = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
                if (residual) o += __ldg(residual + (size_t)m * ldc + n);
                C[(size_t)m * ldc + n] = relu ? fmaxf(o, 0.f) : o;
            }
        }
    }
}

// Few rows (the class term of the relation conv: M = B*N*49, 49 at cfg3): one CTA per output row; a warp takes output
// columns n = warp, warp + 8, ...: its lanes stride over k with coalesced 128-bit loads of the weight row against the
// shared-memory copy of the A row, fp32 FMA, fixed-order shuffle reduction (deterministic).  The 128x128-tile kernel
// above would run this shape on two CTAs (34 us).
__global__ void __launch_bounds__(256)
sgemm_nt_skinny_kernel(const float *__restrict__ A, const int lda, const float *__restrict__ B, const int ldb,
                       const float *__restrict__ bias, float *__restrict__ C, const int ldc, const int N, const int K,
                       const float *__restrict__ residual, const int relu)
{
    extern __shared__ __align__(16) float a_row[];               // [K]
    const int m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4)
        *reinterpret_cast<float4 *>(a_row + k) = ldg4(A + (size_t)m * lda + k);
    __syncthreads();
    const int n_lo = blockIdx.y * 64;                            // 64 columns per CTA: 8 per warp
    for (int n = n_lo + warp; n < min(N, n_lo + 64); n += 8) {
        const float *b = B + (size_t)n * ldb;
        float acc = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 w = ldg4(b + k);
            const float4 a = *reinterpret_cast<const float4 *>(a_row + k);
            acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc);
            acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float o = acc + (bias ? __ldg(bias + n) : 0.f);
            if (residual) o += __ldg(residual + (size_t)m * ldc + n);
            C[(size_t)m * ldc + n] = relu ? fmaxf(o, 0.f) : o;
        }
    }
}

int gemm_nt_simt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C,
                 int ldc, int M, int N, int K, cudaStream_t st, const float *residual, bool relu)
{
    FGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm dims M=%d N=%d K=%d", M, N, K);
    FGN_CHECK_ARG((K & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0, "gemm needs K, lda, ldb multiples of 4 (K=%d lda=%d ldb=%d)", K, lda, ldb);
    FGN_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm operands must be 16-byte aligned");
    if (M == 0) return FGN_OK;
    if (M <= 512 && (size_t)K * 4 <= 48 * 1024) {
        sgemm_nt_skinny_kernel<<<dim3(M, ceil_div(N, 64)), 256, (size_t)K * 4, st>>>(A, lda, B, ldb, bias, C, ldc, N, K, residual, relu ? 1 : 0);
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    FGN_CHECK_ARG(grid.y <= 65535, "gemm N too large");
    sgemm_nt_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, residual, relu ? 1 : 0);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

}  // namespace fgn

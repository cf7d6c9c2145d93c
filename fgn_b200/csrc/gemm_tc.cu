// gemm_tc.cu -- the single-CTA tcgen05 contraction kernel of round 1.  Since round 2 the fp32 contractions (relation conv, its
// adjoint, the heads' convolutions) run on conv_tc.cu's kernels (CTA pairs where the shape allows); what is launched from
// this file is the bf16 variant (kind::f16) and the weights' TF32 split.  Original header:
// tcgen05 (5th-gen tensor core) contraction for the relation head's 1x1 conv
// (fgn_roi_head.py:272): C[M,N] = A[M,K] * B[N,K]^T (+ bias), fp32 in / fp32 out.
//
// fp32 parity on tensor cores: 3xTF32 error compensation.  Every operand is split as
//   x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly a TF32 value),
//                 lo = x - hi (exact in fp32; its own TF32 truncation error is ~2^-21 |x|)
// and D += A_hi B_hi + A_hi B_lo + A_lo B_hi accumulates in fp32 in tensor memory.  The dropped
// term A_lo B_lo is ~2^-22 relative.  precision=1 runs the single A_hi B_hi pass (plain TF32).
//
// Kernel anatomy (persistent, one CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) of A, B_hi, B_lo per
//               32-wide k-block into a 2-stage shared-memory ring, mbarrier complete_tx;
//   warps 8-11  operand splitters: rewrite the landed fp32 A tile as A_hi in place and A_lo beside
//               it (same swizzled positions), fence.proxy.async, signal the MMA warp;
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32
//               (M=128, N<=256, K=8 per instruction, 3 passes), tcgen05.commit frees ring slots and
//               publishes the accumulator;
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered);
//   warps 4-7   epilogue: tcgen05.ld 32x32b.x32, + bias, 128-bit global stores.
#include "gemm.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace fgn {

constexpr int TC_BM = 128;          // rows per tile (UMMA_M, cta_group::1)
constexpr int TC_BN_MAX = 256;      // UMMA_N max
constexpr int TC_THREADS = 384;

// BK = fp32 elements per k-block = one swizzle row: 32 (128-byte swizzle, 2-stage ring) or
// 16 (64-byte swizzle, 4-stage ring: same shared memory, twice the copies in flight).
template <int BK>
struct TcSmem {
    static constexpr int kStages = BK == 32 ? 2 : 4;
    // per stage: A (hi, in place) | A_lo | B_hi | B_lo ; every tile 1024-byte aligned
    static constexpr int kA = TC_BM * BK * 4;               // 16 / 8 KB
    static constexpr int kB = TC_BN_MAX * BK * 4;           // 32 / 16 KB
    static constexpr int kStage = 2 * kA + 2 * kB;          // 96 / 48 KB
    static constexpr int kTotal = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * 32 * 36 * 4 /*epilogue tiles*/;
};


// B [N,K] -> B_hi, B_lo (the TF32 split of the weights; tiny, once per call)
__global__ void split_tf32_kernel(const float *__restrict__ in, int rows, int cols, int ld,
                                  float *__restrict__ hi, float *__restrict__ lo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    const float x = in[(size_t)r * ld + c];
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    hi[i] = h;
    lo[i] = x - h;
}

// BF16 = true: operands are bf16 (kind::f16, K=16 per instruction, single pass); the byte geometry of
// the ring (row bytes = 4*BK) is unchanged, a k-block then holds 2*BK elements.
template <int PASSES, int BK, bool BF16 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                    const __grid_constant__ CUtensorMap map_blo, const float *__restrict__ bias,
                    float *__restrict__ C, const int ldc, const int M, const int N, const int K, const int BN,
                    const float *__restrict__ residual = nullptr, const int relu = 0)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    using Sm = TcSmem<BK>;
    constexpr int TC_STAGES = Sm::kStages;
    constexpr int TC_BK = BK;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * Sm::kStage);
    uint64_t *full_bar = bars, *conv_bar = bars + TC_STAGES, *empty_bar = bars + 2 * TC_STAGES;
    uint64_t *tmem_full = bars + 3 * TC_STAGES, *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = (N + BN - 1) / BN;
    constexpr int KE = BF16 ? 2 * BK : BK;                  // elements per k-block
    const int num_tiles = m_tiles * n_tiles, num_kb = K / KE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(&full_bar[s], 1);
            tc_mbar_init(&conv_bar[s], 4);                 // one arrive per splitter warp
            tc_mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) { tc_mbar_init(&tmem_full[a], 1); tc_mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====================================================================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    tc_mbar_wait(&empty_bar[s], ((it / TC_STAGES) & 1) ^ 1);
                    unsigned char *st = smem + (size_t)s * Sm::kStage;
                    const uint32_t bytes = TC_BM * TC_BK * 4 + (PASSES == 3 ? 2 : 1) * BN * TC_BK * 4;
                    tc_mbar_expect_tx(&full_bar[s], bytes);
                    tma_load_2d(st, &map_a, kb * KE, m0, &full_bar[s]);
                    tma_load_2d(st + 2 * Sm::kA, &map_bhi, kb * KE, n0, &full_bar[s]);
                    if (PASSES == 3) tma_load_2d(st + 2 * Sm::kA + Sm::kB, &map_blo, kb * KE, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ========================================================================
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2,
        // B=TF32 [10,13)=2, K-major A/B, N>>3 at [17,23), M>>4 at [24,29)
        constexpr uint32_t fmt = BF16 ? 1u : 2u;          // F16F32Format: BF16 = 1, TF32 = 2
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        int it = 0, local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int a = local_tile & 1;
            tc_mbar_wait(&tmem_empty[a], ((local_tile >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(a * TC_BN_MAX);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                const uint32_t par = (it / TC_STAGES) & 1;
                tc_mbar_wait(&full_bar[s], par);
                if (PASSES == 3) tc_mbar_wait(&conv_bar[s], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = s_u32(smem + (size_t)s * Sm::kStage);
                    const uint64_t a_hi = umma_desc_kmajor<BK>(st), a_lo = umma_desc_kmajor<BK>(st + Sm::kA);
                    const uint64_t b_hi = umma_desc_kmajor<BK>(st + 2 * Sm::kA);
                    const uint64_t b_lo = umma_desc_kmajor<BK>(st + 2 * Sm::kA + Sm::kB);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);      // +32 B inside the swizzle row
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        if (BF16) umma_bf16(tmem_d, a_hi + ko, b_hi + ko, idesc, acc);
                        else      umma_tf32(tmem_d, a_hi + ko, b_hi + ko, idesc, acc);
                        if (PASSES == 3) {
                            umma_tf32(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                            umma_tf32(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[s]);                               // frees the ring slot when the MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[a]);         // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        // ===== operand splitters (A tile -> A_hi in place, A_lo) ================================
        if (PASSES == 3) {
            const int tid = threadIdx.x - 256;                                // 0..127
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    tc_mbar_wait(&full_bar[s], (it / TC_STAGES) & 1);
                    float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * Sm::kStage);
                    float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * Sm::kStage + Sm::kA);
#pragma unroll
                    for (int j = 0; j < Sm::kA / 16 / 128; ++j) {         // 8 float4 per thread
                        const int i = j * 128 + tid;
                        const float4 x = hi[i];
                        float4 h;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                        hi[i] = h;
                        lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> tensor-core reads
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(&conv_bar[s]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias) -> global =================================
        const int ew = warp - 4;                                              // TMEM lanes 32*ew .. +31
        float *epi_tile = reinterpret_cast<float *>(smem + TC_STAGES * Sm::kStage + 256) + ew * 32 * kEpiPitch;
        int local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int a = local_tile & 1;
            const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
            tc_mbar_wait(&tmem_full[a], (local_tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + ew * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * TC_BN_MAX);
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + (uint32_t)c0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                store_chunk(r, epi_tile, lane, m0 + ew * 32, M, n0 + c0, N, bias, C, ldc, residual, relu != 0);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_mbar_arrive(&tmem_empty[a]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side --------------------------------------------------------------------------------

// 2D fp32 row-major [rows, cols] with row pitch ld (floats); box = bk cols x box_rows, 128B/64B swizzle.
static bool make_map(CUtensorMap *m, const void *base, int rows, int cols, int ld, int box_rows, int bk, bool bf16 = false)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 2 * bk : bk), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t gemm_tc_workspace_bytes(int N, int K) { return (size_t)2 * N * K * sizeof(float); }

// The TF32 split of a weight matrix B [N,K] (row pitch ldb) into split_ws = { B_hi [N,K], B_lo [N,K] }.
int gemm_split_weights(const float *B, int ldb, int N, int K, float *split_ws, cudaStream_t st)
{
    split_tf32_kernel<<<ceil_div(N * K, 256), 256, 0, st>>>(B, N, K, ldb, split_ws, split_ws + (size_t)N * K);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

// bf16 operands (A [M,K], B [N,K], both K-major bf16), fp32 accumulate and output.  The bf16 variant
// of the relation contraction: one pass at the bf16 tensor rate, a quarter of the 3xTF32 operand bytes.
int gemm_nt_tc_bf16(const uint16_t *A, int lda, const uint16_t *B, int ldb, const float *bias, float *C, int ldc,
                    int M, int N, int K, cudaStream_t st)
{
    FGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm dims M=%d N=%d K=%d", M, N, K);
    if (M == 0) return FGN_OK;
    if ((K % 64) != 0 || (N % 16) != 0 || (N > TC_BN_MAX && (N % TC_BN_MAX) != 0) || (lda & 7) || (ldb & 7) || (ldc & 3) ||
        ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15)) {
        set_error("bf16 contraction needs K%%64==0, N%%16==0 (N<=256 or N%%256==0) and 16-byte aligned rows (M=%d N=%d K=%d)", M, N, K);
        return FGN_ERR_UNSUPPORTED;
    }
    const int BN = N > TC_BN_MAX ? TC_BN_MAX : N;
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, M, K, lda, TC_BM, 32, true) || !make_map(&mb, B, N, K, ldb, BN, 32, true)) {
        set_error("cuTensorMapEncodeTiled unavailable or failed");
        return FGN_ERR_CUDA;
    }
    int sm_count = 0;
    if (int rc_sm = current_sm_count(&sm_count)) return rc_sm;
    const int m_tiles = ceil_div(M, TC_BM), n_tiles = ceil_div(N, BN);
    const int grid = min(sm_count, m_tiles * n_tiles);
    FGN_SMEM_OPTIN((gemm_tf32_tc_kernel<1, 32, true>), TcSmem<32>::kTotal);
    gemm_tf32_tc_kernel<1, 32, true><<<grid, TC_THREADS, TcSmem<32>::kTotal, st>>>(ma, mb, mb, bias, C, ldc, M, N, K, BN);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

}  // namespace fgn

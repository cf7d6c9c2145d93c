// gemm_tc.cu -- tcgen05 (5th-gen tensor core) contraction for the relation head's 1x1 conv
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
#include <cuda.h>
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

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(x), "r"(y) : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows x 128 B)
// | version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
// ... for BK=16 the rows are 64 bytes: SBO = 8 rows x 64 B, layout_type = SWIZZLE_64B (4).
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * BK * 4) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(BK == 32 ? 2 : 4) << 61;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// Epilogue store of one 32-row x 32-column accumulator chunk held row-per-lane (tcgen05.ld 32x32b.x32).
// Storing it straight from registers makes every warp-wide 128-bit store touch 32 different rows
// (32 half-written sectors); that alone bounded the K=256 contraction (epilogue ~12 us per tile vs
// 6.3 us of MMAs).  The chunk is transposed through a padded shared-memory tile instead, and each
// store instruction then writes 4 complete 128-byte row segments.
constexpr int kEpiPitch = 36;                                  // floats per staged row: 16-byte aligned, conflict-free
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;              // four epilogue warps

__device__ __forceinline__ void store_chunk(const uint32_t (&r)[32], float *tile, int lane, int row0, int M,
                                            int col0, int N, const float *__restrict__ bias,
                                            float *__restrict__ C, int ldc)
{
    float4 *mine = reinterpret_cast<float4 *>(tile + lane * kEpiPitch);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        mine[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                              __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const bool col_ok = col0 + c4 < N;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr && col_ok) b = ldg4(bias + col0 + c4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + sub;
        float4 o = *reinterpret_cast<const float4 *>(tile + rr * kEpiPitch + c4);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        if (row0 + rr < M && col_ok) *reinterpret_cast<float4 *>(C + (size_t)(row0 + rr) * ldc + col0 + c4) = o;
    }
    __syncwarp();
}

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
                    float *__restrict__ C, const int ldc, const int M, const int N, const int K, const int BN)
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
                store_chunk(r, epi_tile, lane, m0 + ew * 32, M, n0 + c0, N, bias, C, ldc);
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

// ---- relation fusion straight out of tensor memory --------------------------------------------------------
// count_one_roi_by_n_spp + BBoxHead.forward + count_modified_cls_bbox (fgn_roi_head.py:253-279,338,302-326) for the
// FPN shapes (C = 32 * CG <= 256, GroupNorm(32)): the query half of the 1x1 conv runs on tcgen05 exactly as above
// (3xTF32, TMA ring, double-buffered TMEM accumulator), but a tile is TWO RoIs -- rows [98 t, 98 t + 98) of the
// [R*49, C] RoI-feature matrix, the 30 remaining rows of the 128-row MMA are computed and ignored -- with the whole
// N = C of the conv output in tensor memory, so every (RoI, group) GroupNorm statistic is CTA-local.  The epilogue
// warps never write the conv output: per class they add the class term (Ys[b,n], L2-resident), take the GroupNorm
// statistics (per-thread mean / M2 over the CG channels of a group, combined over the 49 rows with Chan's formula),
// apply affine + ReLU, and reduce the 7x7 average pool and the two FC heads to six numbers per (RoI, class); the score
// re-assembly of count_modified_cls_bbox writes [R,N+1] / [R,4N] directly.  Gone: the 50 MB Yq round trip, the
// separate epilogue and finalize launches.
constexpr int kFusedRows = 98;                       // two RoIs x 49 positions per MMA tile
constexpr int kFusedNMax = 8;                        // classes per RoI (more: the separate epilogue kernel scales better)
constexpr int kFusedStages = 3;                      // ring stages (one fewer than the plain contraction: the class term lives in shared memory)
constexpr int kFusedYsPitch = 256 + 4;               // floats per row of the class-term copy (+4: conflict-free 128-bit row reads)
constexpr int kFusedEpiScratch = 20 * 1024;          // FC rows, scale/shift, statistics, per-warp partials (19.4 KB at C = 256)
constexpr int kFusedEpiBytes = kFusedEpiScratch + 49 * kFusedYsPitch * 4;

constexpr int kFusedThreads = 512;                   // 16 warps: the plain kernel's twelve + a second epilogue quartet

template <int CG>
__global__ void __launch_bounds__(kFusedThreads, 1)
relation_fused_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                         const __grid_constant__ CUtensorMap map_blo, const float *__restrict__ Ys,
                         const int32_t *__restrict__ roi_batch, const int R, const int B, const int N, const float eps,
                         const float *__restrict__ gn_w, const float *__restrict__ gn_b,
                         const float *__restrict__ fc_cls_w, const float *__restrict__ fc_cls_b,
                         const float *__restrict__ fc_reg_w, const float *__restrict__ fc_reg_b,
                         float *__restrict__ cls_out, float *__restrict__ reg_out,
                         float *__restrict__ raw_cls, float *__restrict__ raw_reg, const int debug_mode)
{
    // debug_mode (development only): bit 0 = the epilogue only recycles the accumulators (contraction pipeline alone)
    constexpr int BK = 16, C = 32 * CG, PP = 49;
    constexpr int G = 32 / CG;                       // GroupNorm groups per 32-column chunk
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    using Sm = TcSmem<BK>;
    constexpr int TC_STAGES = kFusedStages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * Sm::kStage);
    uint64_t *full_bar = bars, *conv_bar = bars + TC_STAGES, *empty_bar = bars + 2 * TC_STAGES;
    uint64_t *tmem_full = bars + 3 * TC_STAGES, *tmem_empty = tmem_full + 2;
    uint64_t *ys_bar = tmem_empty + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(ys_bar + 1);
    // epilogue shared memory
    float *e_wfc  = reinterpret_cast<float *>(smem + TC_STAGES * Sm::kStage + 256);   // [C][8]: fc_cls rows 0,1, fc_reg rows 0..3
    float *e_scsh = e_wfc + C * 8;                   // [2][C][2]: GroupNorm scale, shift per RoI and channel
    float *e_stat = e_scsh + 2 * C * 2;              // [2][32][2]: mean, rstd per RoI and group
    float *e_dots = e_stat + 2 * 32 * 2;             // [4 warps][2 RoIs][6] (+ slack): per-warp sums of the FC dot products
    float *e_res  = e_dots + 128 * 7;                // [2][kFusedNMax][6]: head outputs per RoI and class
    float *e_part = e_res + 2 * kFusedNMax * 6;      // [4 warps][2 RoIs][32 groups][3]: per-warp sums of the statistics
    // Ys[b,n] ([49, C], the class term of the tile's image and the current class), rows kFusedYsPitch floats apart
    float *e_ys = reinterpret_cast<float *>(smem + TC_STAGES * Sm::kStage + 256 + kFusedEpiScratch);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (R + 1) / 2, num_kb = C / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(&full_bar[s], 1);
            tc_mbar_init(&conv_bar[s], 4);
            tc_mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) { tc_mbar_init(&tmem_full[a], 1); tc_mbar_init(&tmem_empty[a], 256); }
        tc_mbar_init(ys_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 4 && warp < 8) {                     // FC rows, transposed, for broadcast reads in the epilogue
        for (int i = threadIdx.x - 128; i < C * 8; i += 128) {
            const int c = i >> 3, k = i & 7;
            e_wfc[i] = k < 2 ? fc_cls_w[(size_t)k * C + c] : (k < 6 ? fc_reg_w[(size_t)(k - 2) * C + c] : 0.f);
        }
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
                const int m0 = tile * kFusedRows;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    tc_mbar_wait(&empty_bar[s], ((it / TC_STAGES) & 1) ^ 1);
                    unsigned char *st = smem + (size_t)s * Sm::kStage;
                    tc_mbar_expect_tx(&full_bar[s], TC_BM * BK * 4 + 2 * C * BK * 4);
                    tma_load_2d(st, &map_a, kb * BK, m0, &full_bar[s]);
                    tma_load_2d(st + 2 * Sm::kA, &map_bhi, kb * BK, 0, &full_bar[s]);
                    tma_load_2d(st + 2 * Sm::kA + Sm::kB, &map_blo, kb * BK, 0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ========================================================================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
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
                tc_mbar_wait(&conv_bar[s], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = s_u32(smem + (size_t)s * Sm::kStage);
                    const uint64_t a_hi = umma_desc_kmajor<BK>(st), a_lo = umma_desc_kmajor<BK>(st + Sm::kA);
                    const uint64_t b_hi = umma_desc_kmajor<BK>(st + 2 * Sm::kA);
                    const uint64_t b_lo = umma_desc_kmajor<BK>(st + 2 * Sm::kA + Sm::kB);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);
                        umma_tf32(tmem_d, a_hi + ko, b_hi + ko, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                        umma_tf32(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full[a]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== operand splitters (A tile -> A_hi in place, A_lo) ================================
        const int tid = threadIdx.x - 256;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % TC_STAGES;
                tc_mbar_wait(&full_bar[s], (it / TC_STAGES) & 1);
                float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * Sm::kStage);
                float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * Sm::kStage + Sm::kA);
#pragma unroll
                for (int j = 0; j < Sm::kA / 16 / 128; ++j) {
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
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&conv_bar[s]);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = tile row = (RoI of the pair, position p); two warp quartets (warps 4-7 and 12-15: a
        // warp reads the TMEM lane quarter warp % 4) share the accumulator's columns, half of the 32-column chunks each
        const int ew = warp & 3, grp = warp >= 12 ? 1 : 0, te = grp * 128 + ew * 32 + lane;   // TMEM lanes 32*ew .. +31
        constexpr int NCH = C / 32, CH_PER = (NCH + 1) / 2;
        const int ch_lo = grp * CH_PER, ch_hi = min(NCH, ch_lo + CH_PER);
        const int row = ew * 32 + lane;
        const int e = row < PP ? 0 : 1, p = row - e * PP;                     // RoI of the pair, position in it
#define FGN_EPI_BAR() asm volatile("bar.sync 1, 256;" ::: "memory")
        int local_tile = 0;
        int ys_key = -1;                                                       // (image, class) whose class term is in shared memory
        uint32_t ys_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int a = local_tile & 1;
            const int r = 2 * tile + e;
            const bool valid = row < kFusedRows && r < R;
            int b = valid ? roi_batch[r] : 0;
            b = b < 0 ? 0 : (b >= B ? B - 1 : b);
            // the class term is staged in shared memory when both RoIs of the pair belong to one image (RoIs are grouped
            // by image: every pair but the ones that straddle an image boundary); otherwise it is read from L2
            int b0 = roi_batch[min(2 * tile, R - 1)], b1 = roi_batch[min(2 * tile + 1, R - 1)];
            b0 = b0 < 0 ? 0 : (b0 >= B ? B - 1 : b0);
            b1 = b1 < 0 ? 0 : (b1 >= B ? B - 1 : b1);
            const bool staged = b0 == b1;
            tc_mbar_wait(&tmem_full[a], (local_tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * TC_BN_MAX);
            if (debug_mode & 1) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tc_mbar_arrive(&tmem_empty[a]);
                continue;
            }
            for (int n = 0; n < N; ++n) {
                const float *ysp = Ys + ((size_t)(b * N + n) * PP + p) * C;
                if (staged) {
                    if (ys_key != b0 * N + n) {                                // (CTA-uniform)
                        ys_key = b0 * N + n;
                        if (te == 0) {
                            tc_mbar_expect_tx(ys_bar, PP * C * 4);
                            const float *src = Ys + (size_t)ys_key * PP * C;
                            for (int q = 0; q < PP; ++q)
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                             ::"r"(s_u32(e_ys + q * kFusedYsPitch)), "l"(src + (size_t)q * C), "r"(C * 4),
                                               "r"(s_u32(ys_bar)) : "memory");
                        }
                        tc_mbar_wait(ys_bar, ys_phase);
                        ys_phase ^= 1;
                    }
                    ysp = e_ys + p * kFusedYsPitch;
                }
                // ---- pass 1: GroupNorm statistics
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t rr[32];
                    tmem_ld32(taddr + (uint32_t)(ch * 32), rr);
                    float y[32];
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = *reinterpret_cast<const float4 *>(ysp + ch * 32 + 4 * j);
                            y[4 * j] = v.x; y[4 * j + 1] = v.y; y[4 * j + 2] = v.z; y[4 * j + 3] = v.w;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // per thread and group: mean m and M2 about it over the CG channels; over the RoI's 49 rows the plain sums
                    // S1 = sum m, S2 = sum m^2, S3 = sum M2 (M2_total = S3 + CG (S2 - S1^2 / 49): the cancellation-prone term
                    // only sees the row means).  Rows of a RoI are reduced by shuffles inside each warp, per RoI of the pair.
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        float s1 = 0.f;
#pragma unroll
                        for (int j = 0; j < CG; ++j) {
                            y[g * CG + j] = valid ? y[g * CG + j] + __uint_as_float(rr[g * CG + j]) : 0.f;
                            s1 += y[g * CG + j];
                        }
                        const float m = s1 * (1.0f / CG);
                        float m2 = 0.f;
#pragma unroll
                        for (int j = 0; j < CG; ++j) { const float d = y[g * CG + j] - m; m2 = fmaf(d, d, m2); }
#pragma unroll
                        for (int re = 0; re < 2; ++re) {
                            if (!((re == 0 && ew <= 1) || (re == 1 && ew >= 1))) continue;       // (warp-uniform: RoIs this warp holds rows of)
                            const bool in = valid && e == re;
                            float v1 = in ? m : 0.f, v2 = in ? m * m : 0.f, v3 = in ? m2 : 0.f;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                                v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                            }
                            if (lane == 0) {
                                float *dst = e_part + (((ew * 2 + re) * 32) + ch * G + g) * 3;
                                dst[0] = v1; dst[1] = v2; dst[2] = v3;
                            }
                        }
                    }
                }
                FGN_EPI_BAR();
                if (te < 64) {                                                 // one thread per (RoI of the pair, group)
                    const int re = te >> 5, g = te & 31;
                    if (g < C / CG) {
                        float S1 = 0.f, S2 = 0.f, S3 = 0.f;
                        for (int w = (re == 0 ? 0 : 1); w <= (re == 0 ? 1 : 3); ++w) {        // (the quartet that owns the group's chunk wrote it)
                            const float *src = e_part + (((w * 2 + re) * 32) + g) * 3;
                            S1 += src[0]; S2 += src[1]; S3 += src[2];
                        }
                        const float mean = S1 * (1.0f / PP);
                        const float m2 = fmaxf(S3 + (float)CG * (S2 - S1 * mean), 0.f);
                        e_stat[(re * 32 + g) * 2] = mean;
                        e_stat[(re * 32 + g) * 2 + 1] = 1.0f / sqrtf(m2 * (1.0f / (PP * CG)) + eps);
                    }
                }
                FGN_EPI_BAR();
                // GroupNorm affine as torch does it: y*scale + shift, scale = rstd*gamma
                for (int i = te; i < 2 * C; i += 256) {
                    const int re = i / C, c = i - re * C;
                    const float mean = e_stat[(re * 32 + c / CG) * 2], rstd = e_stat[(re * 32 + c / CG) * 2 + 1];
                    const float scale = rstd * gn_w[c];
                    e_scsh[2 * i] = scale;
                    e_scsh[2 * i + 1] = gn_b[c] - mean * scale;
                }
                FGN_EPI_BAR();
                // ---- pass 2: affine + ReLU, 7x7 average pool and the FC heads as per-row dot products
                float dot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int ch = ch_lo; ch < ch_hi; ++ch) {
                    uint32_t rr[32];
                    tmem_ld32(taddr + (uint32_t)(ch * 32), rr);
                    float y[32];
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = *reinterpret_cast<const float4 *>(ysp + ch * 32 + 4 * j);
                            y[4 * j] = v.x; y[4 * j + 1] = v.y; y[4 * j + 2] = v.z; y[4 * j + 3] = v.w;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (valid) {
                        const float2 *ss = reinterpret_cast<const float2 *>(e_scsh) + e * C + ch * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float2 sc = ss[j];
                            const float z = fmaxf(fmaf(y[j] + __uint_as_float(rr[j]), sc.x, sc.y), 0.f);
                            const float4 w0 = *reinterpret_cast<const float4 *>(e_wfc + (ch * 32 + j) * 8);
                            const float2 w1 = *reinterpret_cast<const float2 *>(e_wfc + (ch * 32 + j) * 8 + 4);
                            dot[0] = fmaf(w0.x, z, dot[0]); dot[1] = fmaf(w0.y, z, dot[1]); dot[2] = fmaf(w0.z, z, dot[2]);
                            dot[3] = fmaf(w0.w, z, dot[3]); dot[4] = fmaf(w1.x, z, dot[4]); dot[5] = fmaf(w1.y, z, dot[5]);
                        }
                    }
                }
                if (n == N - 1) {                                              // the accumulator is no longer needed
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    tc_mbar_arrive(&tmem_empty[a]);
                }
#pragma unroll
                for (int re = 0; re < 2; ++re) {
                    if (!((re == 0 && ew <= 1) || (re == 1 && ew >= 1))) continue;
                    const bool in = valid && e == re;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        float v = in ? dot[k] : 0.f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0) e_dots[((grp * 4 + ew) * 2 + re) * 6 + k] = v;
                    }
                }
                FGN_EPI_BAR();
                if (te < 12) {                                                 // (RoI of the pair, head row)
                    const int re = te / 6, k = te - re * 6;
                    float sm = 0.f;
                    for (int q = 0; q < 2; ++q)
                        for (int w = (re == 0 ? 0 : 1); w <= (re == 0 ? 1 : 3); ++w) sm += e_dots[((q * 4 + w) * 2 + re) * 6 + k];
                    e_res[(re * kFusedNMax + n) * 6 + k] = sm * (1.0f / PP) + (k < 2 ? fc_cls_b[k] : fc_reg_b[k - 2]);
                }
                FGN_EPI_BAR();
            }
            // ---- count_modified_cls_bbox (generalised to any N): fg scores, bg of the first-max fg class, deltas
            if (te < 2 && 2 * tile + te < R) {
                const int ro = 2 * tile + te;
                float best_fg = 0.f, best_bg = 0.f;
                for (int n = 0; n < N; ++n) {
                    const float *v = e_res + (te * kFusedNMax + n) * 6;
                    if (raw_cls) { raw_cls[((size_t)ro * N + n) * 2] = v[0]; raw_cls[((size_t)ro * N + n) * 2 + 1] = v[1]; }
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        reg_out[(size_t)ro * 4 * N + 4 * n + d] = v[2 + d];
                        if (raw_reg) raw_reg[((size_t)ro * N + n) * 4 + d] = v[2 + d];
                    }
                    cls_out[(size_t)ro * (N + 1) + n] = v[1];
                    const bool better = n == 0 || v[1] > best_fg || (v[1] != v[1] && best_fg == best_fg);
                    if (better) { best_fg = v[1]; best_bg = v[0]; }
                }
                cls_out[(size_t)ro * (N + 1) + N] = best_bg;
            }
            FGN_EPI_BAR();
        }
#undef FGN_EPI_BAR
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    // (a driver entry point is per process, not per device; C++11 makes the one-time initialisation race-free)
    static const EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)p;
        return nullptr;
    }();
    return fn;
}

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

int gemm_nt_tc(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
               int M, int N, int K, int precision, float *split_ws, cudaStream_t st, bool *taken)
{
    *taken = false;
    if (M <= 0) return FGN_OK;
    // shapes the kernel takes: K in whole 128-byte k-blocks, N tiles of <=256 that are UMMA_N-legal
    if ((K % 32) != 0 || (N % 16) != 0 || (N > TC_BN_MAX && (N % TC_BN_MAX) != 0)) return FGN_OK;
    if ((lda & 3) || (ldb & 3) || (ldc & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15)) return FGN_OK;
    if (split_ws == nullptr) return FGN_OK;
    const int BN = N > TC_BN_MAX ? TC_BN_MAX : N;
    const int passes = precision == 0 ? 3 : 1;
    const char *e = getenv("FGN_GEMM_BK");
    const int bk = (e != nullptr && atoi(e) == 32) ? 32 : 16;

    float *bhi = split_ws, *blo = split_ws + (size_t)N * K;
    split_tf32_kernel<<<ceil_div(N * K, 256), 256, 0, st>>>(B, N, K, ldb, bhi, blo);
    FGN_LAUNCH_OK();

    CUtensorMap ma, mbh, mbl;
    if (!make_map(&ma, A, M, K, lda, TC_BM, bk) || !make_map(&mbh, bhi, N, K, K, BN, bk) || !make_map(&mbl, blo, N, K, K, BN, bk)) {
        set_error("cuTensorMapEncodeTiled unavailable or failed");
        return FGN_ERR_CUDA;
    }
    int sm_count = 0;
    if (int rc_sm = current_sm_count(&sm_count)) return rc_sm;
    const int m_tiles = ceil_div(M, TC_BM), n_tiles = ceil_div(N, BN);
    const int grid = min(sm_count, m_tiles * n_tiles);
#define FGN_TC_LAUNCH(PS, BKV, IDX)                                                                                  \
    do {                                                                                                           \
        FGN_SMEM_OPTIN((gemm_tf32_tc_kernel<PS, BKV>), TcSmem<BKV>::kTotal);                                        \
        gemm_tf32_tc_kernel<PS, BKV><<<grid, TC_THREADS, TcSmem<BKV>::kTotal, st>>>(ma, mbh, mbl, bias, C, ldc, M, N, K, BN); \
    } while (0)
    if (passes == 3 && bk == 32) FGN_TC_LAUNCH(3, 32, 0);
    else if (passes == 3)        FGN_TC_LAUNCH(3, 16, 1);
    else if (bk == 32)           FGN_TC_LAUNCH(1, 32, 2);
    else                         FGN_TC_LAUNCH(1, 16, 3);
#undef FGN_TC_LAUNCH
    FGN_LAUNCH_OK();
    *taken = true;
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

// Fused relation head for C = 64 / 128 / 256, GroupNorm(32), N <= kFusedNMax (the FPN configs).  Xq [R*49, C] row-major
// (NHWC RoI features), Ys [B*N*49, C] (class term incl. conv bias), Wq = conv_w[:, :C] with row pitch ldw.
// split_ws: gemm_tc_workspace_bytes(C, C).  *taken = false: shape not served (caller runs the unfused path).
int relation_fused_tc(const float *Xq, const float *Wq, int ldw, const float *Ys, const int32_t *roi_batch, int R, int B,
                      int N, int C, float eps, const float *gn_w, const float *gn_b, const float *fc_cls_w,
                      const float *fc_cls_b, const float *fc_reg_w, const float *fc_reg_b, float *cls_out,
                      float *reg_out, float *raw_cls, float *raw_reg, float *split_ws, cudaStream_t st, bool *taken)
{
    *taken = false;
    if (R <= 0) return FGN_OK;
    if (!(C == 64 || C == 128 || C == 256) || N > kFusedNMax || split_ws == nullptr) return FGN_OK;
    if (((uintptr_t)Xq & 15) || ((uintptr_t)Wq & 15) || (ldw & 3) || ((uintptr_t)Ys & 15)) return FGN_OK;
    const char *e = getenv("FGN_REL_FUSED");                   // development knob: 0 = separate contraction + epilogue kernels
    if (e != nullptr && atoi(e) == 0) return FGN_OK;
    float *bhi = split_ws, *blo = split_ws + (size_t)C * C;
    split_tf32_kernel<<<ceil_div(C * C, 256), 256, 0, st>>>(Wq, C, C, ldw, bhi, blo);
    FGN_LAUNCH_OK();
    CUtensorMap ma, mbh, mbl;
    if (!make_map(&ma, Xq, R * 49, C, C, TC_BM, 16) || !make_map(&mbh, bhi, C, C, C, C, 16) || !make_map(&mbl, blo, C, C, C, C, 16)) {
        set_error("cuTensorMapEncodeTiled unavailable or failed");
        return FGN_ERR_CUDA;
    }
    int sm_count = 0;
    if (int rc_sm = current_sm_count(&sm_count)) return rc_sm;
    const int tiles = (R + 1) / 2, grid = min(sm_count, tiles);
    const char *ed = getenv("FGN_REL_DEBUG");
    const int dbg = ed != nullptr ? atoi(ed) : 0;
    constexpr int kSmem = kFusedStages * TcSmem<16>::kStage + 1024 + 256 + kFusedEpiBytes;
#define FGN_FUSED_LAUNCH(CGV)                                                                                       \
    do {                                                                                                          \
        FGN_SMEM_OPTIN(relation_fused_tc_kernel<CGV>, kSmem);                                                     \
        relation_fused_tc_kernel<CGV><<<grid, kFusedThreads, kSmem, st>>>(ma, mbh, mbl, Ys, roi_batch, R, B, N, eps, gn_w, gn_b, \
                                                                      fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b, cls_out,   \
                                                                      reg_out, raw_cls, raw_reg, dbg);                  \
    } while (0)
    if (C == 256) FGN_FUSED_LAUNCH(8);
    else if (C == 128) FGN_FUSED_LAUNCH(4);
    else FGN_FUSED_LAUNCH(2);
#undef FGN_FUSED_LAUNCH
    FGN_LAUNCH_OK();
    *taken = true;
    return FGN_OK;
}

}  // namespace fgn

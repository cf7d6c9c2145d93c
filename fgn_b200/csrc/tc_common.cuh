// tc_common.cuh -- tcgen05 / TMA / mbarrier device helpers shared by the tensor-core kernels (gemm_tc.cu, conv_tc.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace fgn {

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
// 6.3 us of MMAs).  The chunk is transposed through a 4 KB shared-memory tile instead (128-byte rows, the 16-byte
// chunk j of row r stored at chunk j ^ (r & 7): conflict-free both ways without padding), and each
// store instruction then writes 4 complete 128-byte row segments.
constexpr int kEpiPitch = 32;                                  // floats per staged row
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;              // four epilogue warps

__device__ __forceinline__ void store_chunk(const uint32_t (&r)[32], float *tile, int lane, int row0, int M,
                                            int col0, int N, const float *__restrict__ bias,
                                            float *__restrict__ C, int ldc,
                                            const float *__restrict__ residual = nullptr, const bool relu = false)
{
    float4 *mine = reinterpret_cast<float4 *>(tile + lane * kEpiPitch);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        mine[j ^ (lane & 7)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const bool col_ok = col0 + c4 < N;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr && col_ok) b = ldg4(bias + col0 + c4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + sub;
        float4 o = *reinterpret_cast<const float4 *>(tile + rr * kEpiPitch + (((lane & 7) ^ (rr & 7)) << 2));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        if (row0 + rr < M && col_ok) {
            if (residual != nullptr) {               // (+ identity branch of a bottleneck, same leading dimension as C)
                const float4 q = ldg4(residual + (size_t)(row0 + rr) * ldc + col0 + c4);
                o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4 *>(C + (size_t)(row0 + rr) * ldc + col0 + c4) = o;
        }
    }
    __syncwarp();
}

// ---- CTA-pair (cta_group::2) variants ---------------------------------------------------------------
// Two CTAs of a cluster on the two SMs of a TPC run ONE 256-row MMA: each holds its own 128 rows of A and HALF of the
// B tile, so a k-block costs each SM half the B bytes of the single-CTA kernel.  Barriers the MMA issuer (CTA rank 0)
// waits on live in rank 0's shared memory; rank 1 reaches them by clearing the rank bit of the shared::cluster address
// (cute::Sm100MmaPeerBitMask).
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on rank 0's copy of `bar`.  Default (cta-scope release) semantics as in cutlass::arch::umma_arrive_2x1SM_sm0: an
// explicit .release.cluster makes ptxas emit MEMBAR.ALL.GPU in front of every arrive (measured: the 3xTF32 pair kernel
// ran 1.6x SLOWER than the single-CTA one with it); the splitters' writes are ordered by their fence.proxy.async.
__device__ __forceinline__ void tc_mbar_arrive_leader(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(s_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// loads issued by either CTA of the pair whose bytes are counted on rank 0's barrier
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar) & kPeerMask), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar)            // arrives on `bar` in BOTH CTAs when the MMAs retire
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(s_u32(bar)), "h"((uint16_t)3) : "memory");
}

// ---- host side: the driver entry point that encodes TMA descriptors ---------------------------
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

}  // namespace fgn

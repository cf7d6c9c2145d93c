// conv_tc.cu -- every fp32 contraction of the library on tcgen05: the relation head's 1x1 conv (fgn_roi_head.py:272) and
// its adjoint's contractions, and the post-RoI heads' convolutions as implicit GEMMs:
//   * the 1x1 / 3x3 (pad 1) convolutions of the C4 res5 shared_head (fgn_roi_head.py:202-233, mmdet Bottleneck [3P]) and
//     of FCNMaskHead.convs (fgn_r50_c4_densecl.py:115-129, num_convs=4 [3P]) over NHWC RoI tiles [R,H,W,Cin];
//   * FCNMaskHead's tail -- ConvTranspose2d(k=2, s=2) + ReLU + conv_logits (1x1) -- as ONE contraction whose epilogue takes
//     the ReLU and the logits' dot product straight out of tensor memory: the [R,2H,2W,Cout] upsampled map (80 MB for 100
//     detections at 14x14 -> 28x28, 256 channels) is never written.
//
// Three kernels, one pipeline (TMA producer warp, MMA issuer warp, operand splitters for the 3xTF32 route, epilogue warps,
// accumulators in tensor memory):
//   conv_tc2_kernel   CTA pairs, tcgen05.mma.cta_group::2 (256-row MMAs, half a B tile per SM): the production kernel;
//   conv_tc2d_kernel  the same with two row tiles per CTA against one half B tile: 3x3 convolutions under one TF32 pass;
//   conv_tc_kernel    one CTA per tile: shapes the pair kernels decline (column tiles not a multiple of 32, one row tile).
//
// No im2col buffer: the A operand of tap (dy,dx) is the SAME activation tensor read through a 4D TMA descriptor
// (C, W, H, R) with the box origin shifted by (dx-1, dy-1); what falls outside the RoI tile is zero-filled by the TMA
// unit, which is exactly the convolution's zero padding.  A tile is a box of whole rows: RB whole RoIs (7x7: two RoIs = 98
// of the 128 MMA rows) or HB rows of one RoI (14x14: nine rows = 126); the MMA rows beyond the box hold stale shared
// memory and produce accumulator rows nobody reads.  The k loop runs over taps x Cin/16; weights are laid out
// [tap][Cout][Cin] (K-major rows, one 2D descriptor).  Precision: 3xTF32 (x = hi + lo with hi = x's top 19 bits;
// D += A_hi B_hi + A_hi B_lo + A_lo B_hi in fp32 tensor memory -- fp32 parity up to the accumulator's truncation, DESIGN.md
// section 2) or one TF32 pass.
#include "gemm.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace fgn {

constexpr int CV_BM = 128, CV_BN_MAX = 256, CV_THREADS = 384, CV_BK = 16, CV_STAGES = 4;
constexpr int CV_A = CV_BM * CV_BK * 4;                      // 8 KB
constexpr int CV_B = CV_BN_MAX * CV_BK * 4;                  // 16 KB
constexpr int CV_STAGE = 2 * CV_A + 2 * CV_B;                // A(hi) | A_lo | B_hi | B_lo = 48 KB
constexpr int CV_SMEM = CV_STAGES * CV_STAGE + 1024 + 256 + kEpiBytes;
constexpr int CV_MAX_CLS = 4;

struct ConvArgs {
    const float *bias, *residual;
    float *out;
    int R, H, W, Cin, Cout;        // activations [R,H,W,Cin]
    int N, BN;                     // GEMM columns per tap (conv: Cout; deconv: 4*Cout) and columns per tile
    int flat;                      // 1: rows are plain [M, Cin] (1x1-type contraction), 128 rows per tile
    int HB, RB, h_blocks;          // box of a tile: W x HB x RB rows
    int taps;                      // 9 (3x3, pad 1) or 1
    int relu;
    const float *w_l, *b_l;        // MODE 1: conv_logits [ncls, Cout], [ncls]
    int ncls;
    int lda, ldc;                  // flat mode: row pitch (floats) of A and of out / residual (conv modes: Cin, Cout)
    int ldb;                       // host side: row pitch of w_taps when it is read in place (one TF32 pass); 0 = Cin
    int debug;                     // FGN_TC_DEBUG (development): bit 0 = epilogue does not read / store, bit 1 = splitters do not split
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct ConvTile { int m0, nv, n0, c1, c2, c3; };

__device__ __forceinline__ ConvTile conv_tile(const ConvArgs &a, int tile, int n_tiles)
{
    ConvTile t;
    const int mt = tile / n_tiles;
    t.n0 = (tile % n_tiles) * a.BN;
    if (a.flat) {
        const int M = a.R * a.H * a.W;
        t.m0 = mt * CV_BM;
        t.nv = min(CV_BM, M - t.m0);
        t.c1 = t.m0; t.c2 = 0; t.c3 = 0;
    } else {
        const int r0 = (mt / a.h_blocks) * a.RB, h0 = (mt % a.h_blocks) * a.HB;
        t.m0 = (r0 * a.H + h0) * a.W;
        t.nv = r0 >= a.R ? 0 : (a.HB == a.H ? min(a.RB, a.R - r0) * a.H * a.W : min(a.HB, a.H - h0) * a.W);   // (padding tiles of a pair / quad: nothing to store)
        t.c1 = 0; t.c2 = h0; t.c3 = r0;
    }
    return t;
}

// One tile's epilogue for one epilogue warp (TMEM lanes 32*ew .. +31 of the accumulator at taddr).
// MODE 0 covers the tile's columns [cbeg, cend) (two warp groups can share a tile); MODE 1 reduces over all of them.
template <int MODE>
__device__ __forceinline__ void conv_epilogue_tile(const ConvArgs &args, const ConvTile &t, uint32_t taddr, float *epi_tile,
                                                   const float *epi_smem, int lane, int ew, int BN, int cbeg = 0, int cend = 1 << 30)
{
    if (MODE == 0) {
        for (int c0 = cbeg; c0 < min(BN, cend); c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            store_chunk(r, epi_tile, lane, t.m0 + ew * 32, t.m0 + t.nv, t.n0 + c0, args.N, args.bias, args.out, args.ldc,
                        args.residual, args.relu != 0);
        }
    } else {
        // deconv tail: this tile's columns are the Cout channels of output pixel (2h+i, 2w+j), ij = n-tile index
        const float *b_d = epi_smem, *w_l = epi_smem + args.Cout;
        float lg[CV_MAX_CLS];
#pragma unroll
        for (int c = 0; c < CV_MAX_CLS; ++c) lg[c] = 0.f;
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)c0, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (c0 + j >= BN) break;                                  // BN need not be a multiple of the chunk
                const float v = fmaxf(__uint_as_float(r[j]) + b_d[c0 + j], 0.f);
#pragma unroll
                for (int c = 0; c < CV_MAX_CLS; ++c)
                    if (c < args.ncls) lg[c] = fmaf(v, w_l[c * args.Cout + c0 + j], lg[c]);
            }
        }
        const int i = ew * 32 + lane;
        if (i < t.nv) {
            const int m = t.m0 + i, hw = args.H * args.W;
            const int rr = m / hw, h = (m % hw) / args.W, w = m % args.W;
            const int ij = t.n0 / BN, oy = 2 * h + (ij >> 1), ox = 2 * w + (ij & 1);
#pragma unroll
            for (int c = 0; c < CV_MAX_CLS; ++c)
                if (c < args.ncls)
                    args.out[(((size_t)rr * args.ncls + c) * (2 * args.H) + oy) * (2 * args.W) + ox] =
                        lg[c] + (args.b_l != nullptr ? __ldg(args.b_l + c) : 0.f);
        }
    }
}

// MODE 0: out = [relu](conv + bias [+ residual]) stored NHWC.  MODE 1: mask logits of the deconv tail.
template <int PASSES, int MODE>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const ConvArgs args)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CV_STAGES * CV_STAGE);
    uint64_t *full_bar = bars, *conv_bar = bars + CV_STAGES, *empty_bar = bars + 2 * CV_STAGES;
    uint64_t *tmem_full = bars + 3 * CV_STAGES, *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *epi_smem = reinterpret_cast<float *>(smem + CV_STAGES * CV_STAGE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = args.BN;
    const int m_tiles = args.flat ? (args.R * args.H * args.W + CV_BM - 1) / CV_BM
                                  : ((args.R + args.RB - 1) / args.RB) * args.h_blocks;
    const int n_tiles = args.N / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int kb_per_tap = args.Cin / CV_BK, num_kb = args.taps * kb_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CV_STAGES; ++s) {
            tc_mbar_init(&full_bar[s], 1);
            tc_mbar_init(&conv_bar[s], 4);
            tc_mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) { tc_mbar_init(&tmem_full[a], 1); tc_mbar_init(&tmem_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (MODE == 1) {
        // deconv bias [Cout] and logits weights [ncls, Cout] staged once: the epilogue reads them as broadcasts
        for (int i = threadIdx.x; i < args.Cout * (1 + args.ncls); i += blockDim.x)
            epi_smem[i] = i < args.Cout ? (args.bias != nullptr ? __ldg(args.bias + i) : 0.f) : __ldg(args.w_l + i - args.Cout);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer: per (tap, k-block) one shifted activation box + the tap's weight rows =====
        if (lane == 0) {
            const uint32_t a_bytes = (args.flat ? CV_BM : args.W * args.HB * args.RB) * CV_BK * 4;
            const uint32_t bytes = a_bytes + (PASSES == 3 ? 2 : 1) * BN * CV_BK * 4;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const ConvTile t = conv_tile(args, tile, n_tiles);
                for (int tap = 0; tap < args.taps; ++tap) {
                    const int dy = args.taps == 9 ? tap / 3 - 1 : 0, dx = args.taps == 9 ? tap % 3 - 1 : 0;
                    const int brow = tap * args.N + t.n0;
                    for (int kb = 0; kb < kb_per_tap; ++kb, ++it) {
                        const int s = it % CV_STAGES;
                        tc_mbar_wait(&empty_bar[s], ((it / CV_STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * CV_STAGE;
                        tc_mbar_expect_tx(&full_bar[s], bytes);
                        tma_load_4d(st, &map_a, kb * CV_BK, t.c1 + dx, t.c2 + dy, t.c3, &full_bar[s]);
                        tma_load_2d(st + 2 * CV_A, &map_bhi, kb * CV_BK, brow, &full_bar[s]);
                        if (PASSES == 3) tma_load_2d(st + 2 * CV_A + CV_B, &map_blo, kb * CV_BK, brow, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (instruction descriptor as in gemm_tc.cu: D=F32, A=B=TF32, K-major, N>>3, M>>4) =====
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CV_BM >> 4) << 24);
        int it = 0, local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int a = local_tile & 1;
            tc_mbar_wait(&tmem_empty[a], ((local_tile >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(a * CV_BN_MAX);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % CV_STAGES;
                const uint32_t par = (it / CV_STAGES) & 1;
                tc_mbar_wait(&full_bar[s], par);
                if (PASSES == 3) tc_mbar_wait(&conv_bar[s], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t st = s_u32(smem + (size_t)s * CV_STAGE);
                    const uint64_t a_hi = umma_desc_kmajor<CV_BK>(st), a_lo = umma_desc_kmajor<CV_BK>(st + CV_A);
                    const uint64_t b_hi = umma_desc_kmajor<CV_BK>(st + 2 * CV_A);
                    const uint64_t b_lo = umma_desc_kmajor<CV_BK>(st + 2 * CV_A + CV_B);
#pragma unroll
                    for (int k = 0; k < CV_BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);
                        umma_tf32(tmem_d, a_hi + ko, b_hi + ko, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        if (PASSES == 3) {
                            umma_tf32(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                            umma_tf32(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full[a]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 8) {
        // ===== operand splitters: landed fp32 A tile -> A_lo beside it (A_hi = A as loaded) =====
        if (PASSES == 3) {
            const int tid = threadIdx.x - 256;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % CV_STAGES;
                    tc_mbar_wait(&full_bar[s], (it / CV_STAGES) & 1);
                    float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * CV_STAGE);
                    float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * CV_STAGE + CV_A);
#pragma unroll
                    for (int j = 0; j < CV_A / 16 / 128; ++j) {
                        const int i = j * 128 + tid;
                        const float4 x = hi[i];
                        float4 h;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                        // (A itself serves as A_hi: the MMA truncates its operands to TF32, see the pair kernel)
                        lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(&conv_bar[s]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int ew = warp - 4;
        float *epi_tile = epi_smem + ew * 32 * kEpiPitch;
        int local_tile = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local_tile) {
            const int a = local_tile & 1;
            const ConvTile t = conv_tile(args, tile, n_tiles);
            tc_mbar_wait(&tmem_full[a], (local_tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * CV_BN_MAX);
            conv_epilogue_tile<MODE>(args, t, taddr, epi_tile, epi_smem, lane, ew, BN);
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

// ---- CTA-pair kernel (cta_group::2) ----------------------------------------------------------------------------------
// The single-CTA kernel above is bound by what the L2 can deliver: per 16-wide k-block an SM needs its 8 KB of A and the
// WHOLE 256-column B tile (16 KB, 32 KB with B_lo), and at ~68 GB/s per SM (148 SMs x that = the L2's ~10 TB/s) the operands
// arrive slower than the tensor core consumes them (profiles/r02_ncu_gemm_tcgen05.json: tensor pipe 57 % of active).  Here two
// CTAs on the SMs of one TPC run one 256-row MMA: each loads its own 128 rows of A and HALF of the B tile (rows
// n0 + rank*BN/2 ..), the tensor cores read both halves, and each SM's tensor memory receives its own 128 x BN block of D.
// Per SM and k-block that is 16 KB instead of 24 KB (one TF32 pass) or 24 KB instead of 40 KB (3xTF32), six ring stages deep.
// Rank 0 issues every MMA; its barriers are the ones the loads of both CTAs count on.
constexpr int C2_BH = (CV_BN_MAX / 2) * CV_BK * 4;           // half B tile: 8 KB
template <int PASSES>
struct C2Ring {                                               // 192 KB of ring either way
    static constexpr int kStage = PASSES == 3 ? 2 * CV_A + 2 * C2_BH : CV_A + C2_BH;   // A(hi) | A_lo | B_hi | B_lo = 32 KB, or A | B = 16 KB
    static constexpr int kStages = PASSES == 3 ? 6 : 12;     // the loop is latency-bound (~2.5 us from a freed slot to its MMAs): depth is throughput
    static constexpr int kOffB = PASSES == 3 ? 2 * CV_A : CV_A;
};
constexpr int C2_THREADS = 512;                               // + warps 12-15: a second epilogue group
constexpr int C2_RING = 6 * (2 * CV_A + 2 * C2_BH);
constexpr int C2_SMEM = C2_RING + 1024 + 512 + 2 * kEpiBytes;

template <int PASSES, int MODE>
__global__ void __launch_bounds__(C2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                const __grid_constant__ CUtensorMap map_blo, const ConvArgs args)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int C2_STAGES = C2Ring<PASSES>::kStages, C2_STAGE = C2Ring<PASSES>::kStage, C2_OFFB = C2Ring<PASSES>::kOffB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C2_RING);
    uint64_t *full_bar = bars;                          // rank 0's: B halves of both CTAs (and both A tiles when PASSES == 1)
    uint64_t *afull_bar = bars + C2_STAGES;             // local: this CTA's A tile landed (PASSES == 3: the splitters wait on it)
    uint64_t *conv_bar = bars + 2 * C2_STAGES;          // rank 0's: the eight splitter warps of the pair
    uint64_t *empty_bar = bars + 3 * C2_STAGES;         // local: the MMAs that read this stage retired (multicast commit)
    uint64_t *tmem_full = bars + 4 * C2_STAGES;         // local, multicast commit
    uint64_t *tmem_empty = tmem_full + 2;               // rank 0's: the eight epilogue warps of the pair
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *epi_smem = reinterpret_cast<float *>(smem + C2_RING + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int BN = args.BN, BH = BN >> 1;
    const int m_tiles = args.flat ? (args.R * args.H * args.W + CV_BM - 1) / CV_BM
                                  : ((args.R + args.RB - 1) / args.RB) * args.h_blocks;
    const int pm_tiles = (m_tiles + 1) >> 1;
    const int n_tiles = args.N / BN;
    const int num_ptiles = pm_tiles * n_tiles;
    const int kb_per_tap = args.Cin / CV_BK, num_kb = args.taps * kb_per_tap;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    // this CTA's tile of pair-tile pt: m-tile 2*pm + rank, same n-tile
#define C2_TILE(pt) conv_tile(args, (((pt) / n_tiles) * 2 + rank) * n_tiles + (pt) % n_tiles, n_tiles)

    if (threadIdx.x == 0) {
        for (int s = 0; s < C2_STAGES; ++s) {
            tc_mbar_init(&full_bar[s], 1);
            tc_mbar_init(&afull_bar[s], 1);
            tc_mbar_init(&conv_bar[s], 8);
            tc_mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) { tc_mbar_init(&tmem_full[a], 1); tc_mbar_init(&tmem_empty[a], MODE == 0 ? 16 : 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (MODE == 1) {
        for (int i = threadIdx.x; i < args.Cout * (1 + args.ncls); i += blockDim.x)
            epi_smem[i] = i < args.Cout ? (args.bias != nullptr ? __ldg(args.bias + i) : 0.f) : __ldg(args.w_l + i - args.Cout);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' barriers exist before anyone arrives on the peer's
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows, own half of the B tile =====
        if (lane == 0) {
            const uint32_t a_bytes = (args.flat ? CV_BM : args.W * args.HB * args.RB) * CV_BK * 4;
            const uint32_t b_bytes = (PASSES == 3 ? 2 : 1) * BH * CV_BK * 4;
            int it = 0;
            for (int pt = pair; pt < num_ptiles; pt += num_pairs) {
                const ConvTile t = C2_TILE(pt);
                for (int tap = 0; tap < args.taps; ++tap) {
                    const int dy = args.taps == 9 ? tap / 3 - 1 : 0, dx = args.taps == 9 ? tap % 3 - 1 : 0;
                    const int brow = tap * args.N + t.n0 + rank * BH;
                    for (int kb = 0; kb < kb_per_tap; ++kb, ++it) {
                        const int s = it % C2_STAGES;
                        tc_mbar_wait(&empty_bar[s], ((it / C2_STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * C2_STAGE;
                        if (PASSES == 3) {
                            tc_mbar_expect_tx(&afull_bar[s], a_bytes);
                            tma_load_4d(st, &map_a, kb * CV_BK, t.c1 + dx, t.c2 + dy, t.c3, &afull_bar[s]);
                            if (rank == 0) tc_mbar_expect_tx(&full_bar[s], 2 * b_bytes);
                            tma_load_2d_2sm(st + C2_OFFB, &map_bhi, kb * CV_BK, brow, &full_bar[s]);
                            tma_load_2d_2sm(st + C2_OFFB + C2_BH, &map_blo, kb * CV_BK, brow, &full_bar[s]);
                        } else {
                            if (rank == 0) tc_mbar_expect_tx(&full_bar[s], 2 * (a_bytes + b_bytes));
                            tma_load_4d_2sm(st, &map_a, kb * CV_BK, t.c1 + dx, t.c2 + dy, t.c3, &full_bar[s]);
                            tma_load_2d_2sm(st + C2_OFFB, &map_bhi, kb * CV_BK, brow, &full_bar[s]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: rank 0 only, one 256 x BN instruction per k step and pass =====
        if (rank == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * CV_BM) >> 4) << 24);
            int it = 0, local_tile = 0;
            for (int pt = pair; pt < num_ptiles; pt += num_pairs, ++local_tile) {
                const int a = local_tile & 1;
                tc_mbar_wait(&tmem_empty[a], ((local_tile >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(a * CV_BN_MAX);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % C2_STAGES;
                    const uint32_t par = (it / C2_STAGES) & 1;
                    tc_mbar_wait(&full_bar[s], par);
                    if (PASSES == 3) tc_mbar_wait(&conv_bar[s], par);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t st = s_u32(smem + (size_t)s * C2_STAGE);
                        const uint64_t a_hi = umma_desc_kmajor<CV_BK>(st), a_lo = umma_desc_kmajor<CV_BK>(st + CV_A);
                        const uint64_t b_hi = umma_desc_kmajor<CV_BK>(st + C2_OFFB);
                        const uint64_t b_lo = umma_desc_kmajor<CV_BK>(st + C2_OFFB + C2_BH);
#pragma unroll
                        for (int k = 0; k < CV_BK / 8; ++k) {
                            const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);
                            umma_tf32_2sm(tmem_d, a_hi + ko, b_hi + ko, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            if (PASSES == 3) {
                                umma_tf32_2sm(tmem_d, a_hi + ko, b_lo + ko, idesc, 1u);
                                umma_tf32_2sm(tmem_d, a_lo + ko, b_hi + ko, idesc, 1u);
                            }
                        }
                        umma_commit_2sm(&empty_bar[s]);                       // both CTAs' ring slots
                        if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[a]); // both CTAs' epilogues
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== operand splitters (both CTAs): own A tile -> A_lo (A_hi = A as loaded); arrive on rank 0's barrier =====
        if (PASSES == 3) {
            const int tid = threadIdx.x - 256;
            int it = 0;
            for (int pt = pair; pt < num_ptiles; pt += num_pairs) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % C2_STAGES;
                    tc_mbar_wait(&afull_bar[s], (it / C2_STAGES) & 1);
                    float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * C2_STAGE);
                    float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * C2_STAGE + CV_A);
                    if (!(args.debug & 2))
#pragma unroll
                    for (int j = 0; j < CV_A / 16 / 128; ++j) {
                        const int i = j * 128 + tid;
                        const float4 x = hi[i];
                        float4 h;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                        // (A itself serves as A_hi: kind::tf32 reads the top 19 bits of an fp32 operand -- clearing the low 13
                        //  here changed no result bit and cost 8 KB of shared-memory writes per k-block)
                        lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive_leader(&conv_bar[s]);
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 rows of D out of own tensor memory.  MODE 0: two groups of four warps (4-7 and
        // 12-15; a warp reaches the TMEM lane quarter warp % 4) take half of the tile's columns each -- one group needs
        // ~8.5 us per 128 x 256 tile, more than the 6.8 us of a K = 256 mainloop it is supposed to hide behind.
        const int ew = warp & 3, grp = warp >= 12 ? 1 : 0;
        if (MODE == 0 || grp == 0) {
            float *epi_tile = epi_smem + (grp * 4 + ew) * 32 * kEpiPitch;
            const int half = ((BN / 32 + 1) / 2) * 32;                      // column split between the groups (chunk-aligned)
            int local_tile = 0;
            for (int pt = pair; pt < num_ptiles; pt += num_pairs, ++local_tile) {
                const int a = local_tile & 1;
                const ConvTile t = C2_TILE(pt);
                tc_mbar_wait(&tmem_full[a], (local_tile >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * CV_BN_MAX);
                if (!(args.debug & 1)) {
                    if (MODE == 0) conv_epilogue_tile<MODE>(args, t, taddr, epi_tile, epi_smem, lane, ew, BN, grp ? half : 0, grp ? BN : half);
                    else           conv_epilogue_tile<MODE>(args, t, taddr, epi_tile, epi_smem, lane, ew, BN);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) tc_mbar_arrive_leader(&tmem_empty[a]);
            }
        }
    }
#undef C2_TILE

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                  // the peer may still be reading this CTA's shared / tensor memory
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- CTA-pair kernel with TWO row tiles per CTA sharing one B tile (3x3 convolutions, one TF32 pass) --------------------
// One TF32 pass is bound by what the L2 delivers per MMA (pair kernel: 6.3-8 KB of A + 8 KB of B per SM and k-block).  A 3x3
// convolution's k loop is long (9 x Cin / 16 k-blocks), so the epilogue need not hide behind the next tile: both 256-column
// accumulators serve ONE step -- two 128-row tiles per CTA against the same half B tile -- and a k-block costs each SM
// 2 x A + B for twice the tensor work: 10.3 instead of 14.3 KB per MMA at 7x7 tiles.  Eight 24 KB stages; no splitters.
constexpr int C2D_STAGES = 8;
constexpr int C2D_STAGE = 2 * CV_A + C2_BH;                  // A (tile 0) | A (tile 1) | B half

__global__ void __launch_bounds__(C2_THREADS, 1)
conv_tc2d_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const ConvArgs args)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C2_RING);
    uint64_t *full_bar = bars;                          // rank 0's: both CTAs' two A boxes and B halves
    uint64_t *empty_bar = bars + C2D_STAGES;            // local (multicast commit)
    uint64_t *tmem_full = bars + 2 * C2D_STAGES;        // local (multicast commit)
    uint64_t *tmem_empty = tmem_full + 1;               // rank 0's: the sixteen epilogue warps of the pair
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 1);
    float *epi_smem = reinterpret_cast<float *>(smem + C2_RING + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int BN = args.BN, BH = BN >> 1;
    const int m_tiles = args.flat ? (args.R * args.H * args.W + CV_BM - 1) / CV_BM
                                  : ((args.R + args.RB - 1) / args.RB) * args.h_blocks;
    const int qm_tiles = (m_tiles + 3) >> 2;
    const int n_tiles = args.N / BN;
    const int num_qtiles = qm_tiles * n_tiles;
    const int kb_per_tap = args.Cin / CV_BK, num_kb = args.taps * kb_per_tap;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    // this CTA's row tile j of step qt: m-tile 4*qm + 2*j + rank (accumulator j = the pair's 256 rows {4qm+2j, 4qm+2j+1})
#define C2D_TILE(qt, j) conv_tile(args, (((qt) / n_tiles) * 4 + 2 * (j) + rank) * n_tiles + (qt) % n_tiles, n_tiles)

    if (threadIdx.x == 0) {
        for (int s = 0; s < C2D_STAGES; ++s) { tc_mbar_init(&full_bar[s], 1); tc_mbar_init(&empty_bar[s], 1); }
        tc_mbar_init(tmem_full, 1);
        tc_mbar_init(tmem_empty, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_ptr)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t a_bytes = (args.flat ? CV_BM : args.W * args.HB * args.RB) * CV_BK * 4;
            const uint32_t b_bytes = BH * CV_BK * 4;
            int it = 0;
            for (int qt = pair; qt < num_qtiles; qt += num_pairs) {
                const ConvTile t0 = C2D_TILE(qt, 0), t1 = C2D_TILE(qt, 1);
                for (int tap = 0; tap < args.taps; ++tap) {
                    const int dy = args.taps == 9 ? tap / 3 - 1 : 0, dx = args.taps == 9 ? tap % 3 - 1 : 0;
                    const int brow = tap * args.N + t0.n0 + rank * BH;
                    for (int kb = 0; kb < kb_per_tap; ++kb, ++it) {
                        const int s = it % C2D_STAGES;
                        tc_mbar_wait(&empty_bar[s], ((it / C2D_STAGES) & 1) ^ 1);
                        unsigned char *st = smem + (size_t)s * C2D_STAGE;
                        if (rank == 0) tc_mbar_expect_tx(&full_bar[s], 2 * (2 * a_bytes + b_bytes));
                        tma_load_4d_2sm(st, &map_a, kb * CV_BK, t0.c1 + dx, t0.c2 + dy, t0.c3, &full_bar[s]);
                        tma_load_4d_2sm(st + CV_A, &map_a, kb * CV_BK, t1.c1 + dx, t1.c2 + dy, t1.c3, &full_bar[s]);
                        tma_load_2d_2sm(st + 2 * CV_A, &map_b, kb * CV_BK, brow, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * CV_BM) >> 4) << 24);
            int it = 0, local_tile = 0;
            for (int qt = pair; qt < num_qtiles; qt += num_pairs, ++local_tile) {
                tc_mbar_wait(tmem_empty, (local_tile & 1) ^ 1);           // both accumulators drained by every epilogue warp
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % C2D_STAGES;
                    tc_mbar_wait(&full_bar[s], (it / C2D_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint32_t st = s_u32(smem + (size_t)s * C2D_STAGE);
                        const uint64_t a0 = umma_desc_kmajor<CV_BK>(st), a1 = umma_desc_kmajor<CV_BK>(st + CV_A);
                        const uint64_t b = umma_desc_kmajor<CV_BK>(st + 2 * CV_A);
#pragma unroll
                        for (int k = 0; k < CV_BK / 8; ++k) {
                            const uint64_t ko = (uint64_t)((k * 8 * 4) >> 4);
                            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                            umma_tf32_2sm(tmem_base, a0 + ko, b + ko, idesc, acc);
                            umma_tf32_2sm(tmem_base + (uint32_t)CV_BN_MAX, a1 + ko, b + ko, idesc, acc);
                        }
                        umma_commit_2sm(&empty_bar[s]);
                        if (kb == num_kb - 1) umma_commit_2sm(tmem_full);
                    }
                    __syncwarp();
                }
            }
        }
    } else if ((warp >= 4 && warp < 8) || warp >= 12) {
        // epilogue: two warp groups split each accumulator's columns, accumulator 0 then 1
        const int ew = warp & 3, grp = warp >= 12 ? 1 : 0;
        float *epi_tile = epi_smem + (grp * 4 + ew) * 32 * kEpiPitch;
        const int half = ((BN / 32 + 1) / 2) * 32;
        int local_tile = 0;
        for (int qt = pair; qt < num_qtiles; qt += num_pairs, ++local_tile) {
            tc_mbar_wait(tmem_full, local_tile & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const ConvTile t = C2D_TILE(qt, j);
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(j * CV_BN_MAX);
                conv_epilogue_tile<0>(args, t, taddr, epi_tile, epi_smem, lane, ew, BN, grp ? half : 0, grp ? BN : half);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive_leader(tmem_empty);
        }
    }
#undef C2D_TILE

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side --------------------------------------------------------------------------------
static bool make_map_nd(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                        const cuuint32_t *box)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;               // out-of-bounds elements read as zero
}

// w_taps [taps, N, Cin] (rows K-major).  w_split = { hi [taps*N, Cin], lo [taps*N, Cin] } or NULL (then split into ws).
static int conv_tc_launch(int mode, const float *x, const float *w_taps, const float *w_split, ConvArgs a, int precision,
                          float *ws, size_t ws_bytes, cudaStream_t st)
{
    const int rows = a.taps * a.N;
    if (a.lda == 0) a.lda = a.Cin;
    if (a.ldc == 0) a.ldc = a.Cout;
    const float *bhi = w_taps, *blo = w_taps;
    if (precision == 0) {
        if (w_split != nullptr) { bhi = w_split; blo = w_split + (size_t)rows * a.Cin; }
        else {
            FGN_CHECK_ARG(ws != nullptr && ws_bytes >= (size_t)2 * rows * a.Cin * sizeof(float),
                          "conv: fp32 precision needs w_split or %zu bytes of workspace", (size_t)2 * rows * a.Cin * sizeof(float));
            if (int rc = gemm_split_weights(w_taps, a.Cin, rows, a.Cin, ws, st)) return rc;
            bhi = ws; blo = ws + (size_t)rows * a.Cin;
        }
    }
    const int m_tiles = a.flat ? ceil_div(a.R * a.H * a.W, CV_BM) : ceil_div(a.R, a.RB) * a.h_blocks;
    if (const char *ed = getenv("FGN_TC_DEBUG")) a.debug = atoi(ed);
    const char *e2 = getenv("FGN_TC_2SM");                       // development knob: 0 = single-CTA kernel
    const bool two_sm = (a.BN % 32) == 0 && m_tiles >= 2 && !(e2 != nullptr && e2[0] == '0');
    CUtensorMap ma, mbh, mbl;
    bool ok;
    {
        const cuuint64_t M = (cuuint64_t)a.R * a.H * a.W;
        cuuint64_t dims[4], strides[3];
        cuuint32_t box[4];
        if (a.flat) {
            dims[0] = a.Cin; dims[1] = M; dims[2] = 1; dims[3] = 1;
            strides[0] = (cuuint64_t)a.lda * 4; strides[1] = strides[2] = M * a.lda * 4;
            box[0] = CV_BK; box[1] = CV_BM; box[2] = 1; box[3] = 1;
        } else {
            dims[0] = a.Cin; dims[1] = a.W; dims[2] = a.H; dims[3] = a.R;
            strides[0] = (cuuint64_t)a.Cin * 4; strides[1] = (cuuint64_t)a.W * a.Cin * 4; strides[2] = (cuuint64_t)a.H * a.W * a.Cin * 4;
            box[0] = CV_BK; box[1] = a.W; box[2] = a.HB; box[3] = a.RB;
        }
        ok = make_map_nd(&ma, x, 4, dims, strides, box);
        const cuuint64_t b_pitch = (precision != 0 && a.ldb != 0) ? (cuuint64_t)a.ldb : (cuuint64_t)a.Cin;   // the split is dense
        cuuint64_t bd[2] = {(cuuint64_t)a.Cin, (cuuint64_t)rows}, bs[1] = {b_pitch * 4};
        cuuint32_t bb[2] = {CV_BK, (cuuint32_t)(two_sm ? a.BN / 2 : a.BN)};   // the pair kernel: half a B tile per CTA
        ok = ok && make_map_nd(&mbh, bhi, 2, bd, bs, bb) && make_map_nd(&mbl, blo, 2, bd, bs, bb);
    }
    if (!ok) { set_error("cuTensorMapEncodeTiled unavailable or failed (conv)"); return FGN_ERR_CUDA; }
    int sm_count = 0;
    if (int rc = current_sm_count(&sm_count)) return rc;
    const char *ed2 = getenv("FGN_TC_DUAL");                      // development knob: 0 = one row tile per CTA also for 3x3 / one pass
    const bool dual = two_sm && mode == 0 && precision != 0 && a.taps == 9 && m_tiles >= 4 && !(ed2 != nullptr && ed2[0] == '0');
    if (dual) {
        const int pairs = min(sm_count / 2, ceil_div(m_tiles, 4) * (a.N / a.BN));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(C2_THREADS); cfg.dynamicSmemBytes = C2_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        FGN_SMEM_OPTIN(conv_tc2d_kernel, C2_SMEM);
        FGN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc2d_kernel, ma, mbh, a));
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    if (two_sm) {
        const int pairs = min(sm_count / 2, ceil_div(m_tiles, 2) * (a.N / a.BN));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(C2_THREADS); cfg.dynamicSmemBytes = C2_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
#define FGN_C2_LAUNCH(PS, MD)                                                                        \
    do {                                                                                             \
        FGN_SMEM_OPTIN((conv_tc2_kernel<PS, MD>), C2_SMEM);                                          \
        FGN_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<PS, MD>, ma, mbh, mbl, a));             \
    } while (0)
        if (mode == 0) { if (precision == 0) FGN_C2_LAUNCH(3, 0); else FGN_C2_LAUNCH(1, 0); }
        else           { if (precision == 0) FGN_C2_LAUNCH(3, 1); else FGN_C2_LAUNCH(1, 1); }
#undef FGN_C2_LAUNCH
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    const int grid = min(sm_count, m_tiles * (a.N / a.BN));
#define FGN_CV_LAUNCH(PS, MD)                                                                        \
    do {                                                                                             \
        FGN_SMEM_OPTIN((conv_tc_kernel<PS, MD>), CV_SMEM);                                           \
        conv_tc_kernel<PS, MD><<<grid, CV_THREADS, CV_SMEM, st>>>(ma, mbh, mbl, a);                  \
    } while (0)
    if (mode == 0) { if (precision == 0) FGN_CV_LAUNCH(3, 0); else FGN_CV_LAUNCH(1, 0); }
    else           { if (precision == 0) FGN_CV_LAUNCH(3, 1); else FGN_CV_LAUNCH(1, 1); }
#undef FGN_CV_LAUNCH
    FGN_LAUNCH_OK();
    return FGN_OK;
}

// The plain contraction C[M,N] = A[M,K] B[N,K]^T (+ bias [+ residual], ReLU): the relation conv, its adjoint's contractions
// and the heads' 1x1 convolutions.  CTA-pair kernel when the column tile splits in two legal halves and there are at least
// two row tiles, else the single-CTA kernel.  *taken = false when the shape does not qualify for either.
int gemm_nt_tc2(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc, int M, int N, int K,
                int precision, float *split_ws, cudaStream_t st, bool presplit, const float *residual, bool relu, bool *taken)
{
    *taken = false;
    if ((K % CV_BK) != 0 || (N % 16) != 0 || (N > CV_BN_MAX && (N % CV_BN_MAX) != 0)) return FGN_OK;
    if ((lda & 3) || (ldb & 3) || (ldc & 3) || (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15)) return FGN_OK;
    if (precision == 0 && split_ws == nullptr) return FGN_OK;
    ConvArgs a = {};
    a.ldb = ldb;
    a.bias = bias; a.residual = residual; a.out = C;
    a.R = M; a.H = 1; a.W = 1; a.Cin = K; a.Cout = N;
    a.N = N; a.BN = N > CV_BN_MAX ? CV_BN_MAX : N;
    a.flat = 1; a.taps = 1; a.HB = 1; a.RB = 1; a.h_blocks = 1; a.relu = relu ? 1 : 0;
    a.lda = lda; a.ldc = ldc;
    if (precision == 0 && !presplit)
        if (int rcs = gemm_split_weights(B, ldb, N, K, split_ws, st)) return rcs;
    const int rc = conv_tc_launch(0, A, B, split_ws, a, precision, nullptr, 0, st);
    if (rc == FGN_OK) *taken = true;
    return rc;
}

}  // namespace fgn

using namespace fgn;

static int conv_common_checks(const void *x, const void *w, const void *out, int R, int H, int W, int Cin, int Cout, int precision)
{
    FGN_CHECK_ARG(precision == 0 || precision == 1, "precision=%d", precision);
    FGN_CHECK_ARG(R >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv dims R=%d H=%d W=%d Cin=%d Cout=%d", R, H, W, Cin, Cout);
    FGN_CHECK_ARG(x && w && out, "NULL pointer");
    FGN_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) == 0, "conv: pointers must be 16-byte aligned");
    if ((Cin % CV_BK) != 0 || (Cout % 16) != 0 || (Cout > CV_BN_MAX && (Cout % CV_BN_MAX) != 0)) {
        set_error("conv: the tcgen05 path needs Cin%%16==0, Cout%%16==0 and (Cout<=256 or Cout%%256==0) (Cin=%d Cout=%d)", Cin, Cout);
        return FGN_ERR_UNSUPPORTED;
    }
    return FGN_OK;
}

extern "C" size_t fgn_conv_split_weights_bytes(int taps, int Cout, int Cin)
{
    return (taps > 0 && Cout > 0 && Cin > 0) ? (size_t)2 * taps * Cout * Cin * sizeof(float) : 0;
}

extern "C" int fgn_conv_split_weights(const float *w_taps, int taps, int Cout, int Cin, float *out, void *stream)
{
    FGN_CHECK_ARG(w_taps && out && taps > 0 && Cout > 0 && Cin > 0, "conv_split_weights: bad arguments");
    return gemm_split_weights(w_taps, Cin, taps * Cout, Cin, out, (cudaStream_t)stream);
}

extern "C" int fgn_conv3x3_nhwc(const float *x, const float *w_taps, const float *w_split, const float *bias,
                                const float *residual, int relu, float *out, int R, int H, int W, int Cin, int Cout,
                                int precision, void *workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = conv_common_checks(x, w_taps, out, R, H, W, Cin, Cout, precision)) return rc;
    if (R == 0) return FGN_OK;
    if (W > CV_BM) { set_error("conv3x3: tiles wider than %d cells are not supported (W=%d)", CV_BM, W); return FGN_ERR_UNSUPPORTED; }
    ConvArgs a = {};
    a.bias = bias; a.residual = residual; a.out = out;
    a.R = R; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.N = Cout; a.BN = Cout > CV_BN_MAX ? CV_BN_MAX : Cout;
    a.flat = 0; a.taps = 9; a.relu = relu;
    if (H * W <= CV_BM) { a.HB = H; a.RB = CV_BM / (H * W); if (a.RB > 256) a.RB = 256; }
    else                { a.HB = CV_BM / W; a.RB = 1; }
    a.h_blocks = ceil_div(H, a.HB);
    return conv_tc_launch(0, x, w_taps, w_split, a, precision, (float *)workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int fgn_deconv2x2_logits_nhwc(const float *x, const float *w_taps, const float *w_split, const float *b_deconv,
                                         const float *w_logits, const float *b_logits, float *mask_pred, int R, int H, int W,
                                         int Cin, int Cout, int ncls, int precision, void *workspace, size_t workspace_bytes,
                                         void *stream)
{
    if (int rc = conv_common_checks(x, w_taps, mask_pred, R, H, W, Cin, Cout, precision)) return rc;
    FGN_CHECK_ARG(w_logits != nullptr && ncls >= 1, "deconv2x2_logits: w_logits NULL or ncls=%d", ncls);
    if (Cout > CV_BN_MAX || ncls > CV_MAX_CLS) {
        set_error("deconv2x2_logits: needs Cout<=256 and ncls<=%d (Cout=%d ncls=%d)", CV_MAX_CLS, Cout, ncls);
        return FGN_ERR_UNSUPPORTED;
    }
    if (R == 0) return FGN_OK;
    ConvArgs a = {};
    a.bias = b_deconv; a.out = mask_pred;
    a.R = R; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.N = 4 * Cout; a.BN = Cout;                  // one column tile per output sub-pixel (i,j)
    a.flat = 1; a.taps = 1; a.HB = 1; a.RB = 1; a.h_blocks = 1;
    a.w_l = w_logits; a.b_l = b_logits; a.ncls = ncls;
    return conv_tc_launch(1, x, w_taps, w_split, a, precision, (float *)workspace, workspace_bytes, (cudaStream_t)stream);
}

// roi_align_stream.cu -- row-streaming multi-level RoIAlign for sm_100a (the fast path).
//
// Same separable formulation and the same exact coordinate arithmetic as roi_align.cu
// (out[ph,pw] = 1/count * sum_y Ay[ph][y] * sum_x Ax[pw][x] * v[y,x]), but organised around the
// memory system instead of around bins:
//
//   * one CTA = one RoI x one block of CB channels (CB = 256 or 128);
//   * a PRODUCER warp walks the RoI's footprint rows once, top to bottom, and pulls each row
//     segment from the NHWC feature map into a shared-memory ring with bulk async copies
//     (cp.async.bulk ... mbarrier::complete_tx -- the TMA engine, SASS UBLKCP): no registers are
//     tied up by loads in flight, and several rows are in flight per CTA at all times;
//   * P CONSUMER warps, warp pw = bin column pw, read the staged row from shared memory
//     (conflict-free 128-bit LDS), form the row's weighted sum over their own cells, and fold it
//     into the (<=3) bin rows the footprint row belongs to.
//
// Every footprint cell therefore crosses L2->SM exactly once per (RoI, channel block): the
// re-reads that adjacent bins share (about 1.5x in the bin-centric kernel) are served from
// shared memory.  full/empty mbarriers form the usual producer/consumer pipeline.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace fgn {

constexpr int kStreamMaxP   = 14;
constexpr int kStageCells   = 32;     // cells (of CB channels) per ring stage

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct StreamPlan {              // written by warp 0, read by the other consumer warps
    int   level;
    float count;
    int   xlo[kStreamMaxP], xn[kStreamMaxP], xoff[kStreamMaxP];   // per bin column: cells + weights
    int   X0, Y0, ncols, nrows, nseg, rps, nstages;               // footprint + ring schedule
};

// Per-lane (lane < 2P: axis = lane / P, bin = lane % P) touched-cell range, then the warp-wide
// footprint and ring schedule via shuffles.  Executed by warp 0 and, redundantly, by the producer
// warp, so that neither has to wait for the other.
struct WarpPlan {
    RoiGeom g;
    int level, H, W;
    int lo, n;                    // this lane's bin
    int X0, Y0, ncols, nrows, nseg, rps, nstages;
};

template <int P>
__device__ __forceinline__ WarpPlan warp_plan(const Pyramid &pyr, const float *__restrict__ rois, int r,
                                              int sampling_ratio, int aligned, float finest_scale,
                                              int lane, int wx_cap, int wyd_rows)
{
    constexpr unsigned FULL = 0xffffffffu;
    WarpPlan wp;
    const float *roi = rois + 5 * (size_t)r;
    wp.level = roi_level(roi, pyr, finest_scale);
    wp.g = roi_geometry(roi, pyr.scale[wp.level], P, sampling_ratio, aligned, pyr.B);
    wp.H = pyr.H[wp.level]; wp.W = pyr.W[wp.level];
    const int axis = lane / P, p = lane % P;
    int lo = 0x7fffffff, hi = -1;
    if (lane < 2 * P) {
        const float start = axis ? wp.g.start_w : wp.g.start_h, bin = axis ? wp.g.bin_w : wp.g.bin_h;
        const int grid = axis ? wp.g.grid_w : wp.g.grid_h, size = axis ? wp.W : wp.H;
        for (int i = 0; i < grid; ++i) {
            const AxisSample s = axis_sample(start, bin, grid, size, p, i);
            if (s.valid) { lo = min(lo, s.low); hi = max(hi, s.high); }
        }
    }
    wp.n  = hi >= 0 ? hi - lo + 1 : 0;
    wp.lo = hi >= 0 ? lo : 0;
    const bool isx = lane >= P && lane < 2 * P, isy = lane < P;
    const int big = 0x7fffffff;
    int X0 = __reduce_min_sync(FULL, (isx && wp.n > 0) ? wp.lo : big);
    int X1 = __reduce_max_sync(FULL, (isx && wp.n > 0) ? wp.lo + wp.n : -1);
    int Y0 = __reduce_min_sync(FULL, (isy && wp.n > 0) ? wp.lo : big);
    int Y1 = __reduce_max_sync(FULL, (isy && wp.n > 0) ? wp.lo + wp.n : -1);
    const int xsum = __reduce_add_sync(FULL, isx ? wp.n : 0);
    if (X1 < 0 || Y1 < 0 || xsum > wx_cap || (Y1 - Y0) > wyd_rows) { X0 = X1 = Y0 = Y1 = 0; }
    wp.X0 = X0; wp.Y0 = Y0; wp.ncols = X1 - X0; wp.nrows = Y1 - Y0;
    if (wp.ncols <= kStageCells) {
        wp.nseg = 1;
        wp.rps = wp.ncols > 0 ? kStageCells / wp.ncols : 1;
        wp.nstages = (wp.nrows + wp.rps - 1) / wp.rps;
    } else {
        wp.nseg = (wp.ncols + kStageCells - 1) / kStageCells;
        wp.rps = 1;
        wp.nstages = wp.nrows * wp.nseg;
    }
    if (wp.ncols == 0 || wp.nrows == 0) wp.nstages = 0;
    return wp;
}

// CTA layout: warps 0..P*WS-1 consumers (warp = bin column x channel slice), warp P*WS = producer.
// Shared memory: ring[NS][kStageCells*CB] | wx[wx_cap] | wyd[wyd_rows][PP8] (+ NCHW staging tile)
template <int P, int VEC, int NS, int WS>
__global__ void __launch_bounds__((P * WS + 1) * 32, (P > 8 || WS * VEC > 2) ? 1 : (WS * VEC == 2 ? 2 : 3))
roi_align_stream_kernel(const Pyramid pyr, const int C, const float *__restrict__ rois, const int R,
                        const int sampling_ratio, const int aligned, const float finest_scale,
                        const float *__restrict__ chan_scale, const int32_t *__restrict__ scale_index,
                        float *__restrict__ out, const int out_layout, int32_t *__restrict__ lvl_out,
                        const int wx_cap, const int wyd_rows, const int size_classes, const int debug_mode)
{
    // debug_mode (development only): 1 = skip the copies (consumers run on stale shared memory),
    //                                2 = skip the consumer arithmetic (copies + pipeline only)
    constexpr int CB  = 128 * VEC * WS;
    constexpr int NCW = P * WS;                     // consumer warps
    constexpr int PP8 = (P + 3) & ~3;               // bin-row weights of one footprint row, padded
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ StreamPlan plan;
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];

    float *ring = reinterpret_cast<float *>(smem_raw);                       // 128 B aligned stages
    float *wx   = ring + (size_t)NS * kStageCells * CB;
    float *wyd  = wx + wx_cap;
    float *stage_out = wyd + (size_t)wyd_rows * PP8;                         // NCHW staging (optional)

    // Largest-first scheduling without a sort: the grid is `size_classes` copies of the work list;
    // copy k only keeps the RoIs of size class k (0 = largest footprints) and the rest of its CTAs
    // retire at once, so the hardware's in-order CTA dispatch starts the long RoIs first and the
    // short ones fill the tail.
    const int nblk  = (C + CB - 1) / CB;
    const int items = R * nblk;
    const int cls   = blockIdx.x / items, item = blockIdx.x - cls * items;
    const int r     = item / nblk;
    const int cb0   = (item % nblk) * CB;
    const int t     = threadIdx.x;
    const int warp  = t >> 5, lane = t & 31;
    const int cbn   = min(CB, C - cb0);                   // channels actually present in this block
    const bool contiguous = (cbn == C);                   // whole cells are adjacent in memory
    const int cstride = contiguous ? C : CB;              // floats between consecutive staged cells
    __shared__ int my_class;

    if (t == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int c = 0;
        if (size_classes > 1) {
            const float *roi = rois + 5 * (size_t)r;
            const float sc = pyr.scale[roi_level(roi, pyr, finest_scale)];
            const float cells = (roi[3] - roi[1]) * sc * (roi[4] - roi[2]) * sc;     // scheduling hint only
            c = cells >= 384.f ? 0 : (cells >= 96.f ? 1 : 2);
            c = min(c, size_classes - 1);
        }
        my_class = c;
    }
    __syncthreads();                                      // reached at once by every warp
    if (my_class != cls) return;

    if (warp == NCW) {
        // ===== producer: own copy of the plan, then bulk async copies of footprint rows ==========
        const WarpPlan wp = warp_plan<P>(pyr, rois, r, sampling_ratio, aligned, finest_scale, lane, wx_cap, wyd_rows);
        const float *fbase = pyr.feat[wp.level] + ((size_t)wp.g.batch * wp.H * wp.W) * C + cb0
                             + ((size_t)wp.Y0 * wp.W + wp.X0) * C;
        const size_t row_pitch = (size_t)wp.W * C;
        int s = 0, par = 1;                               // parity 1: first pass over a fresh barrier
        if (wp.nseg == 1) {
            int row0 = 0;
            for (int st = 0; st < wp.nstages; ++st) {
                const int nr = min(wp.rps, wp.nrows - row0);
                mbar_wait(&empty_bar[s], par);
                float *dst = ring + (size_t)s * kStageCells * CB;
                if (debug_mode == 1) { if (lane == 0) mbar_arrive(&full_bar[s]); row0 += nr; if (++s == NS) { s = 0; par ^= 1; } continue; }
                if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(nr * wp.ncols * cbn * 4));
                __syncwarp();
                if (contiguous) {
                    if (lane < nr)
                        bulk_g2s(dst + (size_t)lane * wp.ncols * cstride, fbase + (size_t)(row0 + lane) * row_pitch,
                                 (uint32_t)(wp.ncols * C * 4), &full_bar[s]);
                } else {
                    for (int cell = lane; cell < nr * wp.ncols; cell += 32) {
                        const int rr = cell / wp.ncols, cc = cell - rr * wp.ncols;
                        bulk_g2s(dst + (size_t)cell * cstride, fbase + (size_t)(row0 + rr) * row_pitch + (size_t)cc * C,
                                 (uint32_t)(cbn * 4), &full_bar[s]);
                    }
                }
                row0 += nr;
                if (++s == NS) { s = 0; par ^= 1; }
            }
        } else {
            for (int row = 0; row < wp.nrows; ++row)
                for (int col0 = 0; col0 < wp.ncols; col0 += kStageCells) {
                    const int nc = min(kStageCells, wp.ncols - col0);
                    mbar_wait(&empty_bar[s], par);
                    float *dst = ring + (size_t)s * kStageCells * CB;
                    if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(nc * cbn * 4));
                    __syncwarp();
                    const float *src = fbase + (size_t)row * row_pitch + (size_t)col0 * C;
                    if (contiguous) {
                        if (lane == 0) bulk_g2s(dst, src, (uint32_t)(nc * C * 4), &full_bar[s]);
                    } else if (lane < nc) {
                        bulk_g2s(dst + (size_t)lane * cstride, src + (size_t)lane * C, (uint32_t)(cbn * 4), &full_bar[s]);
                    }
                    if (++s == NS) { s = 0; par ^= 1; }
                }
        }
    } else {
        // ===== consumers ============================================================================
        if (warp == 0) {
            // plan + weights, warp-local (shuffles, no block barrier), published to the other consumers
            const WarpPlan wp = warp_plan<P>(pyr, rois, r, sampling_ratio, aligned, finest_scale, lane, wx_cap, wyd_rows);
            const int axis = lane / P, p = lane % P;
            int off = 0;                                   // exclusive scan of xn over the x lanes
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const int nq = __shfl_sync(FULL, wp.n, P + q);
                if (lane >= P && q < p) off += nq;
            }
            if (lane == 0) {
                plan.level = wp.level; plan.count = wp.g.count;
                plan.X0 = wp.X0; plan.Y0 = wp.Y0; plan.ncols = wp.ncols; plan.nrows = wp.nrows;
                plan.nseg = wp.nseg; plan.rps = wp.rps; plan.nstages = wp.nstages;
                if (lvl_out != nullptr && cb0 == 0) lvl_out[r] = wp.level;
            }
            if (lane >= P && lane < 2 * P) { plan.xlo[p] = wp.lo; plan.xn[p] = wp.n; plan.xoff[p] = off; }
            for (int i = lane; i < wp.nrows * (PP8 / 4); i += 32)
                reinterpret_cast<float4 *>(wyd)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            if (lane < 2 * P && wp.nstages > 0) {
                const float start = axis ? wp.g.start_w : wp.g.start_h, bin = axis ? wp.g.bin_w : wp.g.bin_h;
                const int grid = axis ? wp.g.grid_w : wp.g.grid_h, size = axis ? wp.W : wp.H;
                if (axis) {
                    float *w = wx + off;
                    for (int i = 0; i < wp.n; ++i) w[i] = 0.f;
                    for (int i = 0; i < grid; ++i) {
                        const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                        if (s.valid) { w[s.low - wp.lo] += s.h; w[s.high - wp.lo] += s.l; }
                    }
                } else {
                    float *w = wyd + p - (size_t)wp.Y0 * PP8;     // wyd[row - Y0][p]
                    for (int i = 0; i < grid; ++i) {
                        const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                        if (s.valid) { w[(size_t)s.low * PP8] += s.h; w[(size_t)s.high * PP8] += s.l; }
                    }
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");   // consumers only

        float4 acc[P][VEC];
#pragma unroll
        for (int ph = 0; ph < P; ++ph)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[ph][v] = make_float4(0.f, 0.f, 0.f, 0.f);

        const int pw = warp % P, slice = warp / P;
        const int nrows = plan.nrows, ncols = plan.ncols, nstages = plan.nstages, rps = plan.rps;
        const int xlo = plan.xlo[pw] - plan.X0, nx = plan.xn[pw];
        const float *wxp = wx + plan.xoff[pw];
        int loff[VEC];                                     // lanes past the channel count re-read channel 0
#pragma unroll
        for (int v = 0; v < VEC; ++v) loff[v] = ((v * WS + slice) * 128 + lane * 4 < cbn) ? (v * WS + slice) * 128 + lane * 4 : 0;

        auto fold = [&](int j, float4 (&racc)[VEC]) {      // footprint row j -> the bin rows it touches
            float wrow[PP8];
#pragma unroll
            for (int q = 0; q < PP8 / 4; ++q) {
                const float4 w4 = *reinterpret_cast<const float4 *>(wyd + (size_t)j * PP8 + 4 * q);
                wrow[4 * q] = w4.x; wrow[4 * q + 1] = w4.y; wrow[4 * q + 2] = w4.z; wrow[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                if (wrow[ph] != 0.f) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) fma4(acc[ph][v], wrow[ph], racc[v]);
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        };

        float4 racc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0, par = 0;
        if (plan.nseg == 1) {
            int row = 0;
            const size_t rstride = (size_t)ncols * cstride;
            for (int st = 0; st < nstages; ++st) {
                const int nr = min(rps, nrows - row);
                mbar_wait(&full_bar[s], par);
                const float *base = ring + (size_t)s * kStageCells * CB + (size_t)xlo * cstride;
                for (int rr = 0; rr < nr; ++rr, ++row) {
                    if (debug_mode == 2) continue;
                    const float *cp = base + rr * rstride;
#pragma unroll 2
                    for (int i = 0; i < nx; ++i, cp += cstride) {
                        const float w = wxp[i];
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            fma4(racc[v], w, *reinterpret_cast<const float4 *>(cp + loff[v]));
                    }
                    fold(row, racc);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                if (++s == NS) { s = 0; par ^= 1; }
            }
        } else {
            for (int row = 0; row < nrows; ++row)
                for (int col0 = 0; col0 < ncols; col0 += kStageCells) {
                    const int nc = min(kStageCells, ncols - col0);
                    mbar_wait(&full_bar[s], par);
                    const float *base = ring + (size_t)s * kStageCells * CB;
                    const int c_beg = max(xlo, col0), c_end = min(xlo + nx, col0 + nc);
                    for (int cx = c_beg; cx < c_end; ++cx) {
                        const float w = wxp[cx - xlo];
                        const float *cp = base + (size_t)(cx - col0) * cstride;
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            fma4(racc[v], w, *reinterpret_cast<const float4 *>(cp + loff[v]));
                    }
                    if (col0 + nc >= ncols) fold(row, racc);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    if (++s == NS) { s = 0; par ^= 1; }
                }
        }

        // ---- epilogue: 1/count, optional AG-FCN channel attention, store ----------------------
        const float inv = 1.0f / plan.count;       // count is a small exact integer; <= 1 ulp vs acc/count
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int cl = (v * WS + slice) * 128 + lane * 4;
            const int c = cb0 + cl;
            if (cl >= cbn) continue;
            float4 cs = make_float4(inv, inv, inv, inv);
            if (chan_scale != nullptr) {
                const int si = scale_index != nullptr ? scale_index[r] : r;
                const float4 a = ldg4(chan_scale + (size_t)si * C + c);
                cs = make_float4(inv * a.x, inv * a.y, inv * a.z, inv * a.w);
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                const float4 a = acc[ph][v];
                float4 o;
                if (chan_scale != nullptr) {               // (acc * 1/count) * vec, same rounding order as unfused
                    const int si = scale_index != nullptr ? scale_index[r] : r;
                    const float4 b = ldg4(chan_scale + (size_t)si * C + c);
                    o = make_float4(a.x * inv * b.x, a.y * inv * b.y, a.z * inv * b.z, a.w * inv * b.w);
                } else {
                    o = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
                }
                if (out_layout == FGN_LAYOUT_NHWC) {
                    *reinterpret_cast<float4 *>(out + (((size_t)r * P + ph) * P + pw) * C + c) = o;
                } else {
                    float *sdst = stage_out + (size_t)cl * (P * P) + ph * P + pw;
                    sdst[0] = o.x; sdst[P * P] = o.y; sdst[2 * P * P] = o.z; sdst[3 * P * P] = o.w;
                }
            }
            (void)cs;
        }
    }
    if (out_layout == FGN_LAYOUT_NCHW) {
        __syncthreads();
        const int n = cbn * P * P;
        float *o = out + ((size_t)r * C + cb0) * (P * P);
        const int n4 = n >> 2;
        for (int i = t; i < n4; i += blockDim.x)
            reinterpret_cast<float4 *>(o)[i] = reinterpret_cast<const float4 *>(stage_out)[i];
        for (int i = (n4 << 2) + t; i < n; i += blockDim.x) o[i] = stage_out[i];
    }
}

// ================================================================================================
// bf16 variant (reported separately from the fp32 contract): the pyramid is stored as NHWC bf16, so
// every footprint cell moves half the bytes through HBM / L2 / shared memory; weights, accumulation
// and the 1/count scaling stay fp32; the pooled features leave as bf16 (the A operand of the bf16
// contraction) or fp32.  One CTA = RoI x 256 channels, lane = 8 contiguous channels (one 128-bit LDS).
// ================================================================================================
template <int P, int NS>
__global__ void __launch_bounds__((P + 1) * 32, P > 8 ? 1 : 2)
roi_align_stream_bf16_kernel(const Pyramid pyr, const int C, const float *__restrict__ rois, const int R,
                             const int sampling_ratio, const int aligned, const float finest_scale,
                             const float *__restrict__ chan_scale, const int32_t *__restrict__ scale_index,
                             void *__restrict__ out, const int out_is_bf16, int32_t *__restrict__ lvl_out,
                             const int wx_cap, const int wyd_rows)
{
    constexpr int CB  = 256;
    constexpr int PP8 = (P + 3) & ~3;
    constexpr unsigned FULL = 0xffffffffu;
    typedef unsigned short bf16_t;                          // raw bf16 storage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ StreamPlan plan;
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];

    bf16_t *ring = reinterpret_cast<bf16_t *>(smem_raw);
    float *wx   = reinterpret_cast<float *>(ring + (size_t)NS * kStageCells * CB);
    float *wyd  = wx + wx_cap;

    const int nblk = (C + CB - 1) / CB;
    const int r    = blockIdx.x / nblk;
    const int cb0  = (blockIdx.x % nblk) * CB;
    const int t    = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cbn  = min(CB, C - cb0);
    const bool contiguous = (cbn == C);
    const int cstride = contiguous ? C : CB;

    if (t == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], P); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == P) {
        const WarpPlan wp = warp_plan<P>(pyr, rois, r, sampling_ratio, aligned, finest_scale, lane, wx_cap, wyd_rows);
        const bf16_t *fbase = reinterpret_cast<const bf16_t *>(pyr.feat[wp.level]) + ((size_t)wp.g.batch * wp.H * wp.W) * C + cb0
                              + ((size_t)wp.Y0 * wp.W + wp.X0) * C;
        const size_t row_pitch = (size_t)wp.W * C;
        int s = 0, par = 1, row0 = 0, col0 = 0;
        for (int st = 0; st < wp.nstages; ++st) {
            int nr, nc;
            if (wp.nseg == 1) { nr = min(wp.rps, wp.nrows - row0); nc = wp.ncols; }
            else              { nr = 1; nc = min(kStageCells, wp.ncols - col0); }
            mbar_wait(&empty_bar[s], par);
            bf16_t *dst = ring + (size_t)s * kStageCells * CB;
            if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(nr * nc * cbn * 2));
            __syncwarp();
            const bf16_t *src = fbase + (size_t)row0 * row_pitch + (size_t)col0 * C;
            if (contiguous) {
                if (lane < nr)
                    bulk_g2s(dst + (size_t)lane * nc * cstride, src + (size_t)lane * row_pitch, (uint32_t)(nc * C * 2), &full_bar[s]);
            } else {
                for (int cell = lane; cell < nr * nc; cell += 32) {
                    const int rr = cell / nc, cc = cell - rr * nc;
                    bulk_g2s(dst + (size_t)cell * cstride, src + (size_t)rr * row_pitch + (size_t)cc * C, (uint32_t)(cbn * 2), &full_bar[s]);
                }
            }
            if (wp.nseg == 1) row0 += nr;
            else { col0 += nc; if (col0 >= wp.ncols) { col0 = 0; ++row0; } }
            if (++s == NS) { s = 0; par ^= 1; }
        }
    } else {
        if (warp == 0) {
            const WarpPlan wp = warp_plan<P>(pyr, rois, r, sampling_ratio, aligned, finest_scale, lane, wx_cap, wyd_rows);
            const int axis = lane / P, p = lane % P;
            int off = 0;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const int nq = __shfl_sync(FULL, wp.n, P + q);
                if (lane >= P && q < p) off += nq;
            }
            if (lane == 0) {
                plan.level = wp.level; plan.count = wp.g.count;
                plan.X0 = wp.X0; plan.Y0 = wp.Y0; plan.ncols = wp.ncols; plan.nrows = wp.nrows;
                plan.nseg = wp.nseg; plan.rps = wp.rps; plan.nstages = wp.nstages;
                if (lvl_out != nullptr && cb0 == 0) lvl_out[r] = wp.level;
            }
            if (lane >= P && lane < 2 * P) { plan.xlo[p] = wp.lo; plan.xn[p] = wp.n; plan.xoff[p] = off; }
            for (int i = lane; i < wp.nrows * (PP8 / 4); i += 32)
                reinterpret_cast<float4 *>(wyd)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            if (lane < 2 * P && wp.nstages > 0) {
                const float start = axis ? wp.g.start_w : wp.g.start_h, bin = axis ? wp.g.bin_w : wp.g.bin_h;
                const int grid = axis ? wp.g.grid_w : wp.g.grid_h, size = axis ? wp.W : wp.H;
                if (axis) {
                    float *w = wx + off;
                    for (int i = 0; i < wp.n; ++i) w[i] = 0.f;
                    for (int i = 0; i < grid; ++i) {
                        const AxisSample sm = axis_sample(start, bin, grid, size, p, i);
                        if (sm.valid) { w[sm.low - wp.lo] += sm.h; w[sm.high - wp.lo] += sm.l; }
                    }
                } else {
                    float *w = wyd + p - (size_t)wp.Y0 * PP8;
                    for (int i = 0; i < grid; ++i) {
                        const AxisSample sm = axis_sample(start, bin, grid, size, p, i);
                        if (sm.valid) { w[(size_t)sm.low * PP8] += sm.h; w[(size_t)sm.high * PP8] += sm.l; }
                    }
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(P * 32) : "memory");

        float4 acc[P][2];
#pragma unroll
        for (int ph = 0; ph < P; ++ph) { acc[ph][0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[ph][1] = acc[ph][0]; }
        const int pw = warp;
        const int nrows = plan.nrows, ncols = plan.ncols, nstages = plan.nstages, rps = plan.rps, nseg = plan.nseg;
        const int xlo = plan.xlo[pw] - plan.X0, nx = plan.xn[pw];
        const float *wxp = wx + plan.xoff[pw];
        const int loff = (lane * 8 < cbn) ? lane * 8 : 0;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;

        auto cell_fma = [&](const bf16_t *cp, float w) {
            const uint4 q = *reinterpret_cast<const uint4 *>(cp + loff);
            ra.x = fmaf(w, __uint_as_float(q.x << 16), ra.x); ra.y = fmaf(w, __uint_as_float(q.x & 0xffff0000u), ra.y);
            ra.z = fmaf(w, __uint_as_float(q.y << 16), ra.z); ra.w = fmaf(w, __uint_as_float(q.y & 0xffff0000u), ra.w);
            rb.x = fmaf(w, __uint_as_float(q.z << 16), rb.x); rb.y = fmaf(w, __uint_as_float(q.z & 0xffff0000u), rb.y);
            rb.z = fmaf(w, __uint_as_float(q.w << 16), rb.z); rb.w = fmaf(w, __uint_as_float(q.w & 0xffff0000u), rb.w);
        };
        auto fold = [&](int j) {
            float wrow[PP8];
#pragma unroll
            for (int q = 0; q < PP8 / 4; ++q) {
                const float4 w4 = *reinterpret_cast<const float4 *>(wyd + (size_t)j * PP8 + 4 * q);
                wrow[4 * q] = w4.x; wrow[4 * q + 1] = w4.y; wrow[4 * q + 2] = w4.z; wrow[4 * q + 3] = w4.w;
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph)
                if (wrow[ph] != 0.f) { fma4(acc[ph][0], wrow[ph], ra); fma4(acc[ph][1], wrow[ph], rb); }
            ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = ra;
        };

        int s = 0, par = 0, row = 0, col0 = 0;
        for (int st = 0; st < nstages; ++st) {
            int nr, nc;
            if (nseg == 1) { nr = min(rps, nrows - row); nc = ncols; }
            else           { nr = 1; nc = min(kStageCells, ncols - col0); }
            mbar_wait(&full_bar[s], par);
            const bf16_t *base = ring + (size_t)s * kStageCells * CB;
            const int c_beg = max(xlo, col0), c_end = min(xlo + nx, col0 + nc);
            for (int rr = 0; rr < nr; ++rr) {
                const bf16_t *cp = base + (size_t)(rr * nc + c_beg - col0) * cstride;
#pragma unroll 4
                for (int cx = c_beg; cx < c_end; ++cx, cp += cstride) cell_fma(cp, wxp[cx - xlo]);
                if (nseg == 1) { fold(row); ++row; }
            }
            if (nseg != 1) { col0 += nc; if (col0 >= ncols) { fold(row); col0 = 0; ++row; } }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == NS) { s = 0; par ^= 1; }
        }

        if (lane * 8 < cbn) {
            const int c = cb0 + lane * 8;
            const float inv = 1.0f / plan.count;
            float4 sa = make_float4(inv, inv, inv, inv), sb = sa;
            if (chan_scale != nullptr) {
                const int si = scale_index != nullptr ? scale_index[r] : r;
                const float4 a = ldg4(chan_scale + (size_t)si * C + c), b = ldg4(chan_scale + (size_t)si * C + c + 4);
                sa = make_float4(inv * a.x, inv * a.y, inv * a.z, inv * a.w);
                sb = make_float4(inv * b.x, inv * b.y, inv * b.z, inv * b.w);
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                const float4 a = acc[ph][0], b = acc[ph][1];
                const size_t o = (((size_t)r * P + ph) * P + pw) * C + c;
                if (out_is_bf16) {
                    uint4 q;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x * sa.x, a.y * sa.y), t1 = __floats2bfloat162_rn(a.z * sa.z, a.w * sa.w);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x * sb.x, b.y * sb.y), t3 = __floats2bfloat162_rn(b.z * sb.z, b.w * sb.w);
                    q.x = *reinterpret_cast<unsigned *>(&t0); q.y = *reinterpret_cast<unsigned *>(&t1);
                    q.z = *reinterpret_cast<unsigned *>(&t2); q.w = *reinterpret_cast<unsigned *>(&t3);
                    *reinterpret_cast<uint4 *>(reinterpret_cast<bf16_t *>(out) + o) = q;
                } else {
                    float *of = reinterpret_cast<float *>(out) + o;
                    *reinterpret_cast<float4 *>(of) = make_float4(a.x * sa.x, a.y * sa.y, a.z * sa.z, a.w * sa.w);
                    *reinterpret_cast<float4 *>(of + 4) = make_float4(b.x * sb.x, b.y * sb.y, b.z * sb.z, b.w * sb.w);
                }
            }
        }
    }
}

int launch_roi_align_stream_bf16(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                                 int aligned, float finest_scale, const float *chan_scale,
                                 const int32_t *scale_index, void *out, int out_is_bf16, int32_t *lvl_out,
                                 cudaStream_t st)
{
    if ((C & 7) != 0 || (P != 7 && P != 14)) {
        set_error("bf16 RoIAlign needs C%%8==0 and P in {7,14} (C=%d P=%d)", C, P);
        return FGN_ERR_UNSUPPORTED;
    }
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int wx_cap = (maxW + 6 * P + 16 + 3) & ~3;
    const int wyd_rows = maxH;
    const int PP8 = (P + 3) & ~3;
    constexpr int NS = 4;
    const size_t smem = (size_t)NS * kStageCells * 256 * 2 + (size_t)wx_cap * 4 + (size_t)wyd_rows * PP8 * 4;
    if (smem > 200 * 1024) { set_error("bf16 RoIAlign: feature map too tall for the weight table (%d rows)", maxH); return FGN_ERR_UNSUPPORTED; }
    const int nblk = (C + 255) / 256;
    if (P == 7) {
        FGN_SMEM_OPTIN((roi_align_stream_bf16_kernel<7, NS>), smem);
        roi_align_stream_bf16_kernel<7, NS><<<R * nblk, 8 * 32, smem, st>>>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                                                         scale_index, out, out_is_bf16, lvl_out, wx_cap, wyd_rows);
    } else {
        FGN_SMEM_OPTIN((roi_align_stream_bf16_kernel<14, NS>), smem);
        roi_align_stream_bf16_kernel<14, NS><<<R * nblk, 15 * 32, smem, st>>>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                                                           scale_index, out, out_is_bf16, lvl_out, wx_cap, wyd_rows);
    }
    FGN_LAUNCH_OK();
    return FGN_OK;
}

template <int P, int VEC, int NS, int WS = 1>
static int launch_stream_cfg(const Pyramid &d, int C, const float *rois, int R, int sampling_ratio,
                             int aligned, float finest_scale, const float *chan_scale,
                             const int32_t *scale_index, float *out, int out_layout,
                             int32_t *lvl_out, cudaStream_t st, bool *taken)
{
    constexpr int CB = 128 * VEC * WS;
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int wx_cap = (maxW + 6 * P + 16 + 3) & ~3;
    const int wyd_rows = maxH;
    constexpr int PP8 = (P + 3) & ~3;
    size_t smem = (size_t)NS * kStageCells * CB * 4 + (size_t)wx_cap * 4 + (size_t)wyd_rows * PP8 * 4;
    if (out_layout == FGN_LAYOUT_NCHW) smem += (size_t)CB * P * P * 4;
    if (smem > 200 * 1024) { *taken = false; return FGN_OK; }
    auto kern = roi_align_stream_kernel<P, VEC, NS, WS>;
    FGN_SMEM_OPTIN(kern, smem);
    const int nblk = (C + CB - 1) / CB;
    const char *e = getenv("FGN_RA_CLASSES");
    // default 1: the 3-class order shortens the kernel alone by ~3% but its retiring CTAs cost more than that
    // when several episodes overlap on the GPU (bench.py); FGN_RA_CLASSES=3 turns it on
    const int size_classes = e != nullptr ? max(1, min(3, atoi(e))) : 1;
    const char *ed = getenv("FGN_RA_DEBUG");
    const int dbg = ed != nullptr ? atoi(ed) : 0;
    kern<<<R * nblk * size_classes, (P * WS + 1) * 32, smem, st>>>(d, C, rois, R, sampling_ratio, aligned, finest_scale,
                                                                    chan_scale, scale_index, out, out_layout, lvl_out,
                                                                    wx_cap, wyd_rows, size_classes, dbg);
    FGN_LAUNCH_OK();
    *taken = true;
    return FGN_OK;
}

// Dispatcher used by fgn_roi_align_ml_fwd.  Declines (taken=false) shapes it has no instantiation for.
int launch_roi_align_stream(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int out_layout, int32_t *lvl_out,
                            cudaStream_t st, int vec_pref, int ns_pref, bool *taken)
{
    *taken = false;
    if ((C & 3) != 0) return FGN_OK;
    if (vec_pref < 0) vec_pref = -vec_pref;
#define FGN_STREAM(PV, VV, NV) launch_stream_cfg<PV, VV, NV>(d, C, rois, R, sampling_ratio, aligned, finest_scale, \
                                                             chan_scale, scale_index, out, out_layout, lvl_out, st, taken)
    if (P == 7 && vec_pref == 3 && C > 128) {          // 2 warps per bin column, 4 channels per lane
        if (ns_pref == 2) return launch_stream_cfg<7, 1, 2, 2>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale, scale_index, out, out_layout, lvl_out, st, taken);
        return launch_stream_cfg<7, 1, 3, 2>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale, scale_index, out, out_layout, lvl_out, st, taken);
    }
    if (P == 7) {
        const int vec = (vec_pref == 1 || C <= 128) ? 1 : 2;
        if (vec == 2) {
            if (ns_pref == 2) return FGN_STREAM(7, 2, 2);
            if (ns_pref == 4) return FGN_STREAM(7, 2, 4);
            return FGN_STREAM(7, 2, 3);
        }
        if (ns_pref == 2) return FGN_STREAM(7, 1, 2);
        if (ns_pref == 3) return FGN_STREAM(7, 1, 3);
        if (ns_pref == 6) return FGN_STREAM(7, 1, 6);
        return FGN_STREAM(7, 1, 4);
    }
    if (P == 14) {
        if (ns_pref == 2) return FGN_STREAM(14, 1, 2);
        if (ns_pref == 4) return FGN_STREAM(14, 1, 4);
        return FGN_STREAM(14, 1, 3);
    }
#undef FGN_STREAM
    return FGN_OK;
}

}  // namespace fgn

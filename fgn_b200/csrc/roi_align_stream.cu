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

struct StreamPlan {
    int   level, batch, H, W;
    float count;
    int   xlo[kStreamMaxP], xn[kStreamMaxP], xoff[kStreamMaxP];   // per bin column: cells + weights
    int   ylo[kStreamMaxP], yn[kStreamMaxP];                      // per bin row: touched rows
    int   X0, X1, Y0, Y1;                                         // footprint [X0,X1) x [Y0,Y1)
    int   ncols, nrows, nseg, rps, nstages;                       // ring schedule
};

// CTA layout: warps 0..P-1 consumers (warp = bin column), warp P = producer.
// Shared memory: wx[wx_cap] | wyd[P][wyd_stride] | ring[NS][kStageCells*CB] (+ NCHW staging tile)
template <int P, int VEC, int NS>
__global__ void __launch_bounds__((P + 1) * 32)
roi_align_stream_kernel(const Pyramid pyr, const int C, const float *__restrict__ rois, const int R,
                        const int sampling_ratio, const int aligned, const float finest_scale,
                        const float *__restrict__ chan_scale, const int32_t *__restrict__ scale_index,
                        float *__restrict__ out, const int out_layout, int32_t *__restrict__ lvl_out,
                        const int wx_cap, const int wyd_stride)
{
    constexpr int CB = 128 * VEC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ StreamPlan plan;
    __shared__ RoiGeom g_s;
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];

    float *ring = reinterpret_cast<float *>(smem_raw);                       // 128 B aligned stages
    float *wx   = ring + (size_t)NS * kStageCells * CB;
    float *wyd  = wx + wx_cap;
    float *stage_out = wyd + (size_t)P * wyd_stride;                         // NCHW staging (optional)

    const int nblk = (C + CB - 1) / CB;
    const int r    = blockIdx.x / nblk;
    const int cb0  = (blockIdx.x % nblk) * CB;
    const int t    = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;

    // ---- plan ------------------------------------------------------------------------------
    if (t == 0) {
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr.L, finest_scale);
        g_s = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned);
        plan.level = lvl; plan.batch = g_s.batch;
        plan.H = pyr.H[lvl]; plan.W = pyr.W[lvl];
        plan.count = g_s.count;
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], P); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (lvl_out != nullptr && cb0 == 0) lvl_out[r] = lvl;
    }
    __syncthreads();
    const RoiGeom g = g_s;
    if (t < 2 * P) {                       // touched-cell range per (axis, bin)
        const int axis = t / P, p = t % P;
        const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
        const int grid = axis ? g.grid_w : g.grid_h, size = axis ? plan.W : plan.H;
        int lo = 0x7fffffff, hi = -1;
        for (int i = 0; i < grid; ++i) {
            const AxisSample s = axis_sample(start, bin, grid, size, p, i);
            if (s.valid) { lo = min(lo, s.low); hi = max(hi, s.high); }
        }
        const int n = hi >= 0 ? hi - lo + 1 : 0;
        if (axis) { plan.xlo[p] = hi >= 0 ? lo : 0; plan.xn[p] = n; }
        else      { plan.ylo[p] = hi >= 0 ? lo : 0; plan.yn[p] = n; }
    }
    __syncthreads();
    if (t == 0) {                          // footprint, x-weight offsets, ring schedule
        int X0 = 0x7fffffff, X1 = -1, Y0 = 0x7fffffff, Y1 = -1, off = 0;
        for (int p = 0; p < P; ++p) {
            plan.xoff[p] = off; off += plan.xn[p];
            if (plan.xn[p] > 0) { X0 = min(X0, plan.xlo[p]); X1 = max(X1, plan.xlo[p] + plan.xn[p]); }
            if (plan.yn[p] > 0) { Y0 = min(Y0, plan.ylo[p]); Y1 = max(Y1, plan.ylo[p] + plan.yn[p]); }
        }
        if (X1 < 0 || Y1 < 0 || off > wx_cap || (Y1 - Y0) > wyd_stride) { X0 = X1 = Y0 = Y1 = 0; }
        plan.X0 = X0; plan.X1 = X1; plan.Y0 = Y0; plan.Y1 = Y1;
        const int ncols = X1 - X0, nrows = Y1 - Y0;
        plan.ncols = ncols; plan.nrows = nrows;
        if (ncols <= kStageCells) {
            plan.nseg = 1;
            plan.rps = ncols > 0 ? max(1, kStageCells / ncols) : 1;
            plan.nstages = (nrows + plan.rps - 1) / plan.rps;
        } else {
            plan.nseg = (ncols + kStageCells - 1) / kStageCells;
            plan.rps = 1;
            plan.nstages = nrows * plan.nseg;
        }
        if (ncols == 0 || nrows == 0) plan.nstages = 0;
    }
    __syncthreads();
    const int nrows = plan.nrows, ncols = plan.ncols, nstages = plan.nstages;
    const int H = plan.H, W = plan.W;
    (void)H;

    // ===== producer (warp P): bulk async copies of footprint rows into the ring ==================
    const float *fbase = pyr.feat[plan.level] + (size_t)plan.batch * plan.H * W * C + cb0;
    const int cbn = min(CB, C - cb0);                     // channels actually present in this block
    const bool contiguous = (cbn == C);                   // whole cells are adjacent in memory
    const int cstride = contiguous ? C : CB;              // floats between consecutive staged cells
    auto produce = [&](int st) {
        const int s = st % NS;
        mbar_wait(&empty_bar[s], ((st / NS) & 1) ^ 1);    // returns at once for the first NS stages
        float *dst = ring + (size_t)s * kStageCells * CB;
        int row0, nr, col0, nc;
        if (plan.nseg == 1) { row0 = st * plan.rps; nr = min(plan.rps, nrows - row0); col0 = 0; nc = ncols; }
        else { row0 = st / plan.nseg; nr = 1; col0 = (st % plan.nseg) * kStageCells; nc = min(kStageCells, ncols - col0); }
        if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(nr * nc * cbn * 4));
        __syncwarp();
        if (contiguous) {
            for (int rr = lane; rr < nr; rr += 32) {
                const float *src = fbase + ((size_t)(plan.Y0 + row0 + rr) * W + plan.X0 + col0) * C;
                bulk_g2s(dst + (size_t)rr * nc * cstride, src, (uint32_t)(nc * C * 4), &full_bar[s]);
            }
        } else {
            for (int cell = lane; cell < nr * nc; cell += 32) {
                const int rr = cell / nc, cc = cell % nc;
                const float *src = fbase + ((size_t)(plan.Y0 + row0 + rr) * W + plan.X0 + col0 + cc) * C;
                bulk_g2s(dst + (size_t)cell * cstride, src, (uint32_t)(cbn * 4), &full_bar[s]);
            }
        }
    };
    // the first NS stages need no empty-wait: put them in flight before the weights are built
    if (warp == P)
        for (int st = 0; st < min(NS, nstages); ++st) produce(st);

    // ---- weights (all threads help; overlaps with the first copies already in flight) ---------
    for (int i = t; i < P * nrows; i += blockDim.x) wyd[(size_t)(i / nrows) * wyd_stride + (i % nrows)] = 0.f;
    __syncthreads();
    if (t < 2 * P && nstages > 0) {
        const int axis = t / P, p = t % P;
        const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
        const int grid = axis ? g.grid_w : g.grid_h, size = axis ? W : plan.H;
        if (axis) {
            float *w = wx + plan.xoff[p];
            const int n = plan.xn[p], lo = plan.xlo[p];
            for (int i = 0; i < n; ++i) w[i] = 0.f;
            for (int i = 0; i < grid; ++i) {
                const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                if (s.valid) { w[s.low - lo] += s.h; w[s.high - lo] += s.l; }
            }
        } else {
            float *w = wyd + (size_t)p * wyd_stride - plan.Y0;        // dense over footprint rows
            for (int i = 0; i < grid; ++i) {
                const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                if (s.valid) { w[s.low] += s.h; w[s.high] += s.l; }
            }
        }
    }
    __syncthreads();

    if (warp == P)
        for (int st = NS; st < nstages; ++st) produce(st);

    if (warp < P) {
        float4 acc[P][VEC];
#pragma unroll
        for (int ph = 0; ph < P; ++ph)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[ph][v] = make_float4(0.f, 0.f, 0.f, 0.f);

        // ===== consumers: warp = bin column pw ==================================================
        const int pw = warp;
        const int xlo = plan.xlo[pw] - plan.X0, nx = plan.xn[pw];
        const float *wxp = wx + plan.xoff[pw];
        float4 racc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int st = 0; st < nstages; ++st) {
            const int s = st % NS;
            mbar_wait(&full_bar[s], (st / NS) & 1);
            const float *src = ring + (size_t)s * kStageCells * CB + lane * 4;
            int row0, nr, col0, nc;
            if (plan.nseg == 1) { row0 = st * plan.rps; nr = min(plan.rps, nrows - row0); col0 = 0; nc = ncols; }
            else { row0 = st / plan.nseg; nr = 1; col0 = (st % plan.nseg) * kStageCells; nc = min(kStageCells, ncols - col0); }
            // my cells inside this stage's column window [col0, col0+nc)
            const int c_beg = max(xlo, col0), c_end = min(xlo + nx, col0 + nc);
            for (int rr = 0; rr < nr; ++rr) {
                const float *rowp = src + (size_t)(rr * nc - col0) * cstride;
#pragma unroll 4
                for (int cx = c_beg; cx < c_end; ++cx) {
                    const float w = wxp[cx - xlo];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (v * 128 + lane * 4 < cbn) {
                            const float4 val = *reinterpret_cast<const float4 *>(rowp + (size_t)cx * cstride + v * 128);
                            fma4(racc[v], w, val);
                        }
                    }
                }
                const bool row_done = (plan.nseg == 1) || (col0 + nc >= ncols);
                if (row_done) {
                    const int j = row0 + rr;                          // footprint row index
#pragma unroll
                    for (int ph = 0; ph < P; ++ph) {
                        const float wy = wyd[(size_t)ph * wyd_stride + j];
                        if (wy != 0.f) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) fma4(acc[ph][v], wy, racc[v]);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }

        // ---- epilogue: 1/count, optional AG-FCN channel attention, store ----------------------
        const float inv = 1.0f / plan.count;       // count is a small exact integer; <= 1 ulp vs acc/count
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int cl = v * 128 + lane * 4;
            const int c = cb0 + cl;
            if (c >= C) continue;
            float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
            if (chan_scale != nullptr) {
                const int si = scale_index != nullptr ? scale_index[r] : r;
                cs = ldg4(chan_scale + (size_t)si * C + c);
            }
#pragma unroll
            for (int ph = 0; ph < P; ++ph) {
                const float4 a = acc[ph][v];
                const float4 o = make_float4(a.x * inv * cs.x, a.y * inv * cs.y, a.z * inv * cs.z, a.w * inv * cs.w);
                if (out_layout == FGN_LAYOUT_NHWC) {
                    *reinterpret_cast<float4 *>(out + (((size_t)r * P + ph) * P + pw) * C + c) = o;
                } else {
                    float *sdst = stage_out + (size_t)cl * (P * P) + ph * P + pw;
                    sdst[0] = o.x; sdst[P * P] = o.y; sdst[2 * P * P] = o.z; sdst[3 * P * P] = o.w;
                }
            }
        }
    }
    if (out_layout == FGN_LAYOUT_NCHW) {
        __syncthreads();
        const int cb_n = min(CB, C - cb0);
        const int n = cb_n * P * P;
        float *o = out + ((size_t)r * C + cb0) * (P * P);
        const int n4 = n >> 2;
        for (int i = t; i < n4; i += blockDim.x)
            reinterpret_cast<float4 *>(o)[i] = reinterpret_cast<const float4 *>(stage_out)[i];
        for (int i = (n4 << 2) + t; i < n; i += blockDim.x) o[i] = stage_out[i];
    }
}

template <int P, int VEC, int NS>
static int launch_stream_cfg(const Pyramid &d, int C, const float *rois, int R, int sampling_ratio,
                             int aligned, float finest_scale, const float *chan_scale,
                             const int32_t *scale_index, float *out, int out_layout,
                             int32_t *lvl_out, cudaStream_t st, bool *taken)
{
    constexpr int CB = 128 * VEC;
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int wx_cap = (maxW + 6 * P + 16 + 3) & ~3;
    const int wyd_stride = (maxH + 3) & ~3;
    size_t smem = (size_t)NS * kStageCells * CB * 4 + (size_t)wx_cap * 4 + (size_t)P * wyd_stride * 4;
    if (out_layout == FGN_LAYOUT_NCHW) smem += (size_t)CB * P * P * 4;
    if (smem > 200 * 1024) { *taken = false; return FGN_OK; }
    auto kern = roi_align_stream_kernel<P, VEC, NS>;
    static int attr_set = 0;          // per instantiation
    if ((int)smem > attr_set) {
        FGN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = (int)smem;
    }
    const int nblk = (C + CB - 1) / CB;
    kern<<<R * nblk, (P + 1) * 32, smem, st>>>(d, C, rois, R, sampling_ratio, aligned, finest_scale,
                                                chan_scale, scale_index, out, out_layout, lvl_out,
                                                wx_cap, wyd_stride);
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
#define FGN_STREAM(PV, VV, NV) launch_stream_cfg<PV, VV, NV>(d, C, rois, R, sampling_ratio, aligned, finest_scale, \
                                                             chan_scale, scale_index, out, out_layout, lvl_out, st, taken)
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

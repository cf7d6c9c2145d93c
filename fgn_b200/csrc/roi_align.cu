// roi_align.cu -- multi-level RoIAlign forward for sm_100a.
//
// Replaces SingleRoIExtractor.forward -> map_roi_levels -> mmcv.ops.RoIAlign (reference call
// sites fgn_roi_head.py:331-332,366-367, config fgn_r50_c4_densecl.py:69-73) and
// torchvision.ops.roi_align (fgn_roi_head.py:429,432).
//
// Formulation.  The reference averages gh x gw bilinear samples per bin.  Because the sampling
// grid is a cartesian product and the bilinear weight is a product hy*hx, the bin value is
//     out[ph,pw] = 1/count * sum_y Ay[ph][y] * sum_x Ax[pw][x] * v[y,x]
// with per-axis cell weights  A[p][cell] = sum over valid samples i of (h if cell==low) +
// (l if cell==high).  Each CTA builds the two small tables for its RoI in shared memory from
// the exact reference coordinate arithmetic (common.cuh), then every footprint cell is loaded
// ONCE per bin-row as a 128-bit NHWC channel vector instead of 4 corner loads per sample.
// Sample indices are therefore identical to the reference by construction (and exported by
// roi_align_sample_indices_kernel from the same device functions); values differ only by fp32
// summation order (<= ~1e-6 relative).
#include "common.cuh"
#include <stdlib.h>

namespace fgn {

constexpr int kMaxP = 16;

struct RoiPlan {               // lives in shared memory, one per CTA
    int   level, batch, H, W;
    float count;
    int   lo[2][kMaxP];        // first touched cell per (axis, bin); axis 0 = y, 1 = x
    int   n[2][kMaxP];         // number of touched cells (0 = bin has no valid sample)
    int   off[2][kMaxP];       // offset of the bin's weights inside wtab
    int   overflow;            // weight table did not fit (never with a table sized by the host)
};

// Builds RoiPlan + weight table.  All threads of the CTA must call it; contains __syncthreads.
template <int P>
__device__ __forceinline__ void build_plan(const Pyramid &pyr, const float *rois, int r,
                                           int sampling_ratio, int aligned, float finest_scale,
                                           RoiPlan &plan, float *wtab, int wtab_cap,
                                           RoiGeom &g_out)
{
    __shared__ RoiGeom g_s;
    const int t = threadIdx.x;
    if (t == 0) {
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr, finest_scale);
        g_s = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned);
        plan.level = lvl; plan.batch = g_s.batch;
        plan.H = pyr.H[lvl]; plan.W = pyr.W[lvl];
        plan.count = g_s.count; plan.overflow = 0;
    }
    __syncthreads();
    const RoiGeom g = g_s;
    // phase A: touched-cell range per (axis, bin)
    if (t < 2 * P) {
        const int axis = t / P, p = t % P;
        const float start = axis ? g.start_w : g.start_h;
        const float bin   = axis ? g.bin_w : g.bin_h;
        const int   grid  = axis ? g.grid_w : g.grid_h;
        const int   size  = axis ? plan.W : plan.H;
        int lo = 0x7fffffff, hi = -1;
        for (int i = 0; i < grid; ++i) {
            const AxisSample s = axis_sample(start, bin, grid, size, p, i);
            if (s.valid) { lo = min(lo, s.low); hi = max(hi, s.high); }
        }
        plan.lo[axis][p] = hi >= 0 ? lo : 0;
        plan.n[axis][p]  = hi >= 0 ? hi - lo + 1 : 0;
    }
    __syncthreads();
    // phase B: offsets (tiny serial prefix sum) + phase C: weights
    if (t < 2 * P) {
        const int axis = t / P, p = t % P;
        int off = 0;
        for (int a = 0; a <= axis; ++a)
            for (int q = 0; q < (a == axis ? p : P); ++q) off += plan.n[a][q];
        plan.off[axis][p] = off;
        const int n = plan.n[axis][p];
        if (off + n > wtab_cap) { plan.overflow = 1; }
        else {
            float *w = wtab + off;
            for (int i = 0; i < n; ++i) w[i] = 0.f;
            const float start = axis ? g.start_w : g.start_h;
            const float bin   = axis ? g.bin_w : g.bin_h;
            const int   grid  = axis ? g.grid_w : g.grid_h;
            const int   size  = axis ? plan.W : plan.H;
            const int   lo    = plan.lo[axis][p];
            for (int i = 0; i < grid; ++i) {
                const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                if (s.valid) { w[s.low - lo] += s.h; w[s.high - lo] += s.l; }
            }
        }
    }
    __syncthreads();
    g_out = g;
}

// One CTA = one RoI x one block of CB channels; one warp = (bin-row ph, 128-channel chunk).
// NHWC input.  Output NHWC ([R,P,P,C], direct 512 B warp stores) or NCHW ([R,C,P,P], staged
// through shared memory and written as one contiguous CB*P*P*4-byte run).
template <int P, int NB>
__global__ void __launch_bounds__(448)
roi_align_sep_nhwc_kernel(const Pyramid pyr, const int C, const int CB,
                          const float *__restrict__ rois, const int R,
                          const int sampling_ratio, const int aligned, const float finest_scale,
                          const float *__restrict__ chan_scale,
                          const int32_t *__restrict__ scale_index,
                          float *__restrict__ out, const int out_layout,
                          int32_t *__restrict__ lvl_out, const int wtab_cap)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ RoiPlan plan;
    float *wtab  = smem;                      // [wtab_cap]
    float *stage = smem + wtab_cap;           // [CB][P*P] only for NCHW output

    const int nblk = (C + CB - 1) / CB;
    const int r    = blockIdx.x / nblk;
    const int cb0  = (blockIdx.x % nblk) * CB;
    RoiGeom g;
    build_plan<P>(pyr, rois, r, sampling_ratio, aligned, finest_scale, plan, wtab, wtab_cap, g);
    if (lvl_out != nullptr && cb0 == 0 && threadIdx.x == 0) lvl_out[r] = plan.level;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int chunks = (CB + 127) / 128;
    const int H = plan.H, W = plan.W;
    const float *fbase = pyr.feat[plan.level] + (size_t)plan.batch * H * W * C;

    for (int item = warp; item < P * chunks; item += nwarps) {
        const int ph = item % P, chunk = item / P;
        const int cl = chunk * 128 + lane * 4;            // channel inside the CTA's block
        const int c  = cb0 + cl;
        const bool active = (cl < CB) && (c < C);
        float4 acc[P];
#pragma unroll
        for (int pw = 0; pw < P; ++pw) acc[pw] = make_float4(0.f, 0.f, 0.f, 0.f);

        if (active && !plan.overflow) {
            const int ylo = plan.lo[0][ph], ny = plan.n[0][ph];
            const float *wy = wtab + plan.off[0][ph];
            for (int yi = 0; yi < ny; ++yi) {
                const float wyv = wy[yi];
                const float *row = fbase + ((size_t)(ylo + yi) * W) * C + c;
#pragma unroll
                for (int pw = 0; pw < P; ++pw) {
                    const int nx = plan.n[1][pw];
                    const float *wx = wtab + plan.off[1][pw];
                    const float *cell = row + (size_t)plan.lo[1][pw] * C;
                    float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);
                    // first NB cells of the bin as one batch of independent, predicated 128-bit
                    // loads (memory-level parallelism); adaptive grids rarely need more
                    float4 v[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        v[k] = k < nx ? ldg4(cell + (size_t)k * C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < NB; ++k) fma4(racc, k < nx ? wx[k] : 0.f, v[k]);
                    for (int xi = NB; xi < nx; ++xi) {
                        const float4 u = ldg4(cell + (size_t)xi * C);
                        fma4(racc, wx[xi], u);
                    }
                    fma4(acc[pw], wyv, racc);
                }
            }
        }
        // epilogue: divide by the sample count, optional AG-FCN channel attention
        float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
        if (active && chan_scale != nullptr) {
            const int si = scale_index != nullptr ? scale_index[r] : r;
            cs = ldg4(chan_scale + (size_t)si * C + c);
        }
#pragma unroll
        for (int pw = 0; pw < P; ++pw) {
            // acc/count as in the reference; count is an integer-valued float, and x*(1/count)
            // differs from x/count by at most 1 ulp -- use the true division to stay closest.
            acc[pw].x = __fdiv_rn(acc[pw].x, plan.count) * cs.x;
            acc[pw].y = __fdiv_rn(acc[pw].y, plan.count) * cs.y;
            acc[pw].z = __fdiv_rn(acc[pw].z, plan.count) * cs.z;
            acc[pw].w = __fdiv_rn(acc[pw].w, plan.count) * cs.w;
        }
        if (out_layout == FGN_LAYOUT_NHWC) {
            if (active) {
                float *o = out + (((size_t)r * P + ph) * P) * C + c;
#pragma unroll
                for (int pw = 0; pw < P; ++pw)
                    *reinterpret_cast<float4 *>(o + (size_t)pw * C) = acc[pw];
            }
        } else if (active) {
            float *s = stage + (size_t)cl * (P * P) + ph * P;
#pragma unroll
            for (int pw = 0; pw < P; ++pw) {
                s[pw]             = acc[pw].x;
                s[pw + P * P]     = acc[pw].y;
                s[pw + 2 * P * P] = acc[pw].z;
                s[pw + 3 * P * P] = acc[pw].w;
            }
        }
    }
    if (out_layout == FGN_LAYOUT_NCHW) {
        __syncthreads();
        const int cb_n = min(CB, C - cb0);
        const int n = cb_n * P * P;                       // contiguous run in out
        float *o = out + ((size_t)r * C + cb0) * (P * P);
        // ((r*C+cb0)*P*P) % 4 == 0 whenever C % 4 == 0 and cb0 % 4 == 0 -> 16 B aligned
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x)
            reinterpret_cast<float4 *>(o)[i] = reinterpret_cast<const float4 *>(stage)[i];
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) o[i] = stage[i];
    }
}

// Direct (one thread per output element) RoIAlign in the reference's own NCHW layout and
// summation order.  Used for NCHW inputs, channel counts that are not a multiple of 4, pooled
// sizes without a tuned instantiation, and as an in-library cross-check of the separable kernel.
__global__ void roi_align_direct_kernel(const Pyramid pyr, const int C, const int in_layout,
                                        const float *__restrict__ rois, const int R, const int P,
                                        const int sampling_ratio, const int aligned,
                                        const float finest_scale,
                                        const float *__restrict__ chan_scale,
                                        const int32_t *__restrict__ scale_index,
                                        float *__restrict__ out, const int out_layout,
                                        int32_t *__restrict__ lvl_out)
{
    const size_t total = (size_t)R * C * P * P;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        // NCHW order for idx so consecutive threads share (r,c) and walk the bins
        const int pw = idx % P, ph = (idx / P) % P;
        const int c = (idx / ((size_t)P * P)) % C;
        const int r = idx / ((size_t)P * P * C);
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr, finest_scale);
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned);
        const int H = pyr.H[lvl], W = pyr.W[lvl];
        if (lvl_out != nullptr && c == 0 && ph == 0 && pw == 0) lvl_out[r] = lvl;
        const float *f = pyr.feat[lvl];
        size_t base, sy, sx;
        if (in_layout == FGN_LAYOUT_NCHW) {
            base = ((size_t)g.batch * C + c) * (size_t)H * W; sy = W; sx = 1;
        } else {
            base = (size_t)g.batch * H * W * C + c; sy = (size_t)W * C; sx = C;
        }
        float acc = 0.f;
        for (int iy = 0; iy < g.grid_h; ++iy) {
            const AxisSample y = axis_sample(g.start_h, g.bin_h, g.grid_h, H, ph, iy);
            if (!y.valid) continue;
            for (int ix = 0; ix < g.grid_w; ++ix) {
                const AxisSample x = axis_sample(g.start_w, g.bin_w, g.grid_w, W, pw, ix);
                if (!x.valid) continue;
                const float w1 = __fmul_rn(y.h, x.h), w2 = __fmul_rn(y.h, x.l);
                const float w3 = __fmul_rn(y.l, x.h), w4 = __fmul_rn(y.l, x.l);
                const float v1 = __ldg(f + base + y.low * sy + x.low * sx);
                const float v2 = __ldg(f + base + y.low * sy + x.high * sx);
                const float v3 = __ldg(f + base + y.high * sy + x.low * sx);
                const float v4 = __ldg(f + base + y.high * sy + x.high * sx);
                // ((w1v1 + w2v2) + w3v3) + w4v4, then += : the reference's order, unfused
                const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)),
                                                      __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
                acc = __fadd_rn(acc, val);
            }
        }
        float o = __fdiv_rn(acc, g.count);
        if (chan_scale != nullptr) {
            const int si = scale_index != nullptr ? scale_index[r] : r;
            o *= __ldg(chan_scale + (size_t)si * C + c);
        }
        if (out_layout == FGN_LAYOUT_NCHW) out[idx] = o;
        else out[(((size_t)r * P + ph) * P + pw) * C + c] = o;
    }
}

__global__ void roi_align_sample_indices_kernel(const Pyramid pyr, const float *__restrict__ rois,
                                                const int R, const int P, const int sampling_ratio,
                                                const int aligned, const float finest_scale,
                                                const int max_grid, int32_t *__restrict__ lvl_out,
                                                int32_t *__restrict__ grid_out,
                                                int32_t *__restrict__ ytab, int32_t *__restrict__ xtab)
{
    const size_t per_roi = (size_t)2 * P * max_grid;
    const size_t total = (size_t)R * per_roi;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int r = idx / per_roi;
        const int rem = idx % per_roi;
        const int axis = rem / (P * max_grid);
        const int p = (rem / max_grid) % P, i = rem % max_grid;
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr, finest_scale);
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned);
        if (rem == 0) {
            if (lvl_out) lvl_out[r] = lvl;
            grid_out[2 * r] = g.grid_h; grid_out[2 * r + 1] = g.grid_w;
        }
        int32_t *t = (axis ? xtab : ytab) + (((size_t)r * P + p) * max_grid + i) * 3;
        const int grid = axis ? g.grid_w : g.grid_h;
        if (i < grid) {
            const AxisSample s = axis ? axis_sample(g.start_w, g.bin_w, g.grid_w, pyr.W[lvl], p, i)
                                      : axis_sample(g.start_h, g.bin_h, g.grid_h, pyr.H[lvl], p, i);
            t[0] = s.valid; t[1] = s.low; t[2] = s.high;
        } else { t[0] = t[1] = t[2] = -1; }
    }
}

__global__ void map_roi_levels_kernel(const Pyramid pyr, const float *__restrict__ rois, const int R,
                                      const float finest_scale, int32_t *__restrict__ lvl)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) lvl[r] = roi_level(rois + 5 * (size_t)r, pyr, finest_scale);
}

// ---- host side ------------------------------------------------------------------------------

static int validate_pyramid(const fgn_pyramid_t *pyr)
{
    FGN_CHECK_ARG(pyr != nullptr, "pyramid is NULL");
    FGN_CHECK_ARG(pyr->num_levels >= 1 && pyr->num_levels <= FGN_MAX_LEVELS,
                  "num_levels=%d outside [1,%d]", pyr->num_levels, FGN_MAX_LEVELS);
    for (int l = 0; l < pyr->num_levels; ++l) {
        FGN_CHECK_ARG(pyr->H[l] > 0 && pyr->W[l] > 0, "level %d has empty extent %dx%d", l,
                      pyr->H[l], pyr->W[l]);
    }
    return FGN_OK;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int P, int NB>
static int launch_sep_nb(const Pyramid &d, int C, int CB, const float *rois, int R, int sampling_ratio,
                         int aligned, float finest_scale, const float *chan_scale,
                         const int32_t *scale_index, float *out, int out_layout, int32_t *lvl_out,
                         int wtab_cap, cudaStream_t st)
{
    const int warps = P * (CB / 128);
    size_t smem = (size_t)wtab_cap * 4;
    if (out_layout == FGN_LAYOUT_NCHW) smem += (size_t)CB * P * P * 4;
    auto kern = roi_align_sep_nhwc_kernel<P, NB>;
    static int attr_set = 48 * 1024;      // per instantiation
    if ((int)smem > attr_set) {
        FGN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = (int)smem;
    }
    const int nblk = (C + CB - 1) / CB;
    kern<<<R * nblk, warps * 32, smem, st>>>(d, C, CB, rois, R, sampling_ratio, aligned,
                                              finest_scale, chan_scale, scale_index, out,
                                              out_layout, lvl_out, wtab_cap);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

template <int P>
static int launch_sep(const Pyramid &d, int C, const float *rois, int R, int sampling_ratio,
                      int aligned, float finest_scale, const float *chan_scale,
                      const int32_t *scale_index, float *out, int out_layout, int32_t *lvl_out,
                      cudaStream_t st)
{
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    // sum of touched cells over the bins of one axis <= extent + 2 per bin boundary
    int wtab_cap = maxH + maxW + 6 * P + 16;
    wtab_cap = (wtab_cap + 3) & ~3;
    // channel block per CTA: bounded by 32 warps (P * CB/128) and by the NCHW staging tile
    int CB = env_int("FGN_RA_CB", 256);
    if (P > 8) CB = 128;
    if (C < CB) CB = ((C + 127) / 128) * 128;
    const int nb = env_int("FGN_RA_NB", 4);
#define FGN_SEP(NBV) launch_sep_nb<P, NBV>(d, C, CB, rois, R, sampling_ratio, aligned, finest_scale, \
                                           chan_scale, scale_index, out, out_layout, lvl_out, wtab_cap, st)
    if (P == 7) {
        if (nb <= 1) return FGN_SEP(1);
        if (nb == 2) return FGN_SEP(2);
        if (nb == 3) return FGN_SEP(3);
        if (nb == 6) return FGN_SEP(6);
        if (nb >= 8) return FGN_SEP(8);
    }
    return FGN_SEP(4);
#undef FGN_SEP
}

int launch_roi_align_stream(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int out_layout, int32_t *lvl_out,
                            cudaStream_t st, int vec_pref, int ns_pref, bool *taken);

int launch_roi_align_stream_bf16(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                                 int aligned, float finest_scale, const float *chan_scale,
                                 const int32_t *scale_index, void *out, int out_is_bf16, int32_t *lvl_out,
                                 cudaStream_t st);

int launch_roi_align_window(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int32_t *lvl_out, void *workspace,
                            size_t workspace_bytes, cudaStream_t st, int ns_pref, bool *taken);
size_t roi_align_window_workspace_bytes(const Pyramid &d, int R, int P);
int launch_roi_align_gather(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int32_t *lvl_out, void *workspace,
                            size_t workspace_bytes, cudaStream_t st, bool *taken);
size_t roi_align_gather_workspace_bytes(const Pyramid &d, int R, int P);
unsigned int roi_align_gather_violations();
unsigned int roi_align_window_violations();
void roi_align_window_trace(unsigned long long *dst, int n);
void roi_align_window_trace_reset();

}  // namespace fgn

using namespace fgn;

extern "C" int fgn_map_roi_levels(const float *rois, int R, int num_levels, float finest_scale,
                                  int32_t *lvl_out, void *stream)
{
    FGN_CHECK_ARG(R >= 0, "R=%d", R);
    FGN_CHECK_ARG(num_levels >= 1 && num_levels <= FGN_MAX_LEVELS, "num_levels=%d", num_levels);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && lvl_out, "NULL pointer");
    Pyramid d;
    d.L = num_levels;
    for (int i = 0; i < FGN_MAX_LEVELS; ++i) { d.feat[i] = nullptr; d.H[i] = d.W[i] = 0; d.scale[i] = 0.f; }
    level_thresholds(d.lvl_thr);
    map_roi_levels_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(d, rois, R, finest_scale, lvl_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" size_t fgn_roi_align_ml_workspace_bytes(const fgn_pyramid_t *pyr, int R, int P)
{
    if (pyr == nullptr || validate_pyramid(pyr) != 0 || R <= 0 || P <= 0) return 0;
    const Pyramid d = to_device_pyramid(pyr);
    return max(roi_align_window_workspace_bytes(d, R, P), roi_align_gather_workspace_bytes(d, R, P));
}

extern "C" int fgn_roi_align_ml_fwd(const fgn_pyramid_t *pyr, int B, int C, int in_layout,
                                    const float *rois, int R, int P, int sampling_ratio,
                                    int aligned, float finest_scale, const float *chan_scale,
                                    const int32_t *scale_index, float *out, int out_layout,
                                    int32_t *lvl_out, void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = validate_pyramid(pyr);
    if (rc) return rc;
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && P > 0, "bad dims R=%d B=%d C=%d P=%d", R, B, C, P);
    FGN_CHECK_ARG(in_layout == FGN_LAYOUT_NCHW || in_layout == FGN_LAYOUT_NHWC, "in_layout=%d", in_layout);
    FGN_CHECK_ARG(out_layout == FGN_LAYOUT_NCHW || out_layout == FGN_LAYOUT_NHWC, "out_layout=%d", out_layout);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && out, "NULL pointer");
    for (int l = 0; l < pyr->num_levels; ++l) FGN_CHECK_ARG(pyr->feat[l], "level %d pointer is NULL", l);
    const Pyramid d = to_device_pyramid(pyr);
    cudaStream_t st = (cudaStream_t)stream;
    // FGN_RA_IMPL (development knob): 4 = persistent rotating-window kernel (default, NHWC out),
    // 3 / 2 = row-streaming kernel (one CTA per RoI), 1 = bin-centric, 0 = direct
    const int impl = env_int("FGN_RA_IMPL", 5);
    if (in_layout == FGN_LAYOUT_NHWC && out_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 5) {
        bool taken = false;
        rc = launch_roi_align_gather(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                     scale_index, out, lvl_out, workspace, workspace_bytes, st, &taken);
        if (rc || taken) return rc;
    }
    if (in_layout == FGN_LAYOUT_NHWC && out_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 4) {
        bool taken = false;
        rc = launch_roi_align_window(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                     scale_index, out, lvl_out, workspace, workspace_bytes, st,
                                     env_int("FGN_RA_NS", 0), &taken);
        if (rc || taken) return rc;
    }
    if (in_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 2) {
        bool taken = false;
        rc = launch_roi_align_stream(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                     scale_index, out, out_layout, lvl_out, st,
                                     impl == 3 ? env_int("FGN_RA_VEC", 2) : -env_int("FGN_RA_VEC", 2),
                                     env_int("FGN_RA_NS", 0), &taken);
        if (rc || taken) return rc;
    }
    if (in_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 1) {
        if (P == 7)  return launch_sep<7>(d, C, rois, R, sampling_ratio, aligned, finest_scale,
                                          chan_scale, scale_index, out, out_layout, lvl_out, st);
        if (P == 14) return launch_sep<14>(d, C, rois, R, sampling_ratio, aligned, finest_scale,
                                           chan_scale, scale_index, out, out_layout, lvl_out, st);
    }
    const size_t total = (size_t)R * C * P * P;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    roi_align_direct_kernel<<<blocks, 256, 0, st>>>(d, C, in_layout, rois, R, P, sampling_ratio,
                                                    aligned, finest_scale, chan_scale, scale_index,
                                                    out, out_layout, lvl_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

// Planner self-check of the rotating-window kernel (test/debug export): number of (row, bin) weights
// that fell outside the register window since the library was loaded.  Must be 0.
extern "C" unsigned int fgn_debug_roi_window_violations(void)
{
    return roi_align_window_violations() + roi_align_gather_violations();
}

// Development trace of the rotating-window kernel (FGN_RA_DEBUG bit 5); not declared in the public header.
extern "C" void fgn_debug_roi_window_trace(unsigned long long *dst, int n)
{
    if (dst == nullptr) roi_align_window_trace_reset();
    else roi_align_window_trace(dst, n);
}

// Same entry, forcing the direct kernel (exported for the in-library cross-check in tests).
extern "C" int fgn_roi_align_ml_fwd_direct(const fgn_pyramid_t *pyr, int B, int C, int in_layout,
                                           const float *rois, int R, int P, int sampling_ratio,
                                           int aligned, float finest_scale, const float *chan_scale,
                                           const int32_t *scale_index, float *out, int out_layout,
                                           int32_t *lvl_out, void *stream)
{
    int rc = validate_pyramid(pyr);
    if (rc) return rc;
    FGN_CHECK_ARG(R >= 0 && C > 0 && P > 0, "bad dims");
    (void)B;
    if (R == 0) return FGN_OK;
    const Pyramid d = to_device_pyramid(pyr);
    const size_t total = (size_t)R * C * P * P;
    const int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    roi_align_direct_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        d, C, in_layout, rois, R, P, sampling_ratio, aligned, finest_scale, chan_scale,
        scale_index, out, out_layout, lvl_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_roi_align_sample_indices(const fgn_pyramid_t *pyr, const float *rois, int R,
                                            int P, int sampling_ratio, int aligned,
                                            float finest_scale, int max_grid, int32_t *lvl_out,
                                            int32_t *grid_out, int32_t *ytab_out,
                                            int32_t *xtab_out, void *stream)
{
    int rc = validate_pyramid(pyr);
    if (rc) return rc;
    FGN_CHECK_ARG(R >= 0 && P > 0 && max_grid > 0, "bad dims");
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && grid_out && ytab_out && xtab_out, "NULL pointer");
    const Pyramid d = to_device_pyramid(pyr);
    const size_t total = (size_t)R * 2 * P * max_grid;
    const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    roi_align_sample_indices_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        d, rois, R, P, sampling_ratio, aligned, finest_scale, max_grid, lvl_out, grid_out,
        ytab_out, xtab_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_roi_align_ml_fwd_bf16(const fgn_pyramid_t *pyr, int B, int C, const float *rois, int R, int P,
                                         int sampling_ratio, int aligned, float finest_scale,
                                         const float *chan_scale, const int32_t *scale_index, void *out,
                                         int out_is_bf16, int32_t *lvl_out, void *stream)
{
    int rc = validate_pyramid(pyr);
    if (rc) return rc;
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && P > 0, "bad dims R=%d B=%d C=%d P=%d", R, B, C, P);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && out, "NULL pointer");
    for (int l = 0; l < pyr->num_levels; ++l) FGN_CHECK_ARG(pyr->feat[l], "level %d pointer is NULL", l);
    const Pyramid d = to_device_pyramid(pyr);
    return launch_roi_align_stream_bf16(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                        scale_index, out, out_is_bf16, lvl_out, (cudaStream_t)stream);
}

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
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned, pyr.B);
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

// Few RoIs (the support branch: one box per support image, fgn_roi_head.py:432): the persistent window kernel is
// one CTA per RoI chunk and a 26x26-cell support box then costs ~27 us of per-row latency on an otherwise idle GPU.
// Here a warp takes (RoI, bin, 128-channel block), lane = 4 channels: the samples of the bin are visited in the
// reference's own (iy, ix) order with four 128-bit corner loads each -- the corner cells of neighbouring samples
// hit in L1 -- so the whole launch is R * P * P * C/128 independent warps.  Same summation order as the direct kernel
// (bit-exact against torchvision), NHWC in, NHWC or NCHW out.
__global__ void __launch_bounds__(256)
roi_align_small_nhwc_kernel(const Pyramid pyr, const int C, const float *__restrict__ rois, const int R, const int P,
                            const int sampling_ratio, const int aligned, const float finest_scale,
                            const float *__restrict__ chan_scale, const int32_t *__restrict__ scale_index,
                            float *__restrict__ out, const int out_layout, int32_t *__restrict__ lvl_out)
{
    const int nblk = (C + 127) >> 7;
    const size_t warps = (size_t)R * P * P * nblk;
    const int lane = threadIdx.x & 31;
    for (size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; w < warps; w += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const int cb = (int)(w % nblk);
        const int pw = (int)((w / nblk) % P), ph = (int)((w / ((size_t)nblk * P)) % P);
        const int r = (int)(w / ((size_t)nblk * P * P));
        const int c = cb * 128 + lane * 4;
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr, finest_scale);
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned, pyr.B);
        const int H = pyr.H[lvl], W = pyr.W[lvl];
        if (lvl_out != nullptr && cb == 0 && ph == 0 && pw == 0 && lane == 0) lvl_out[r] = lvl;
        if (c >= C) continue;
        const float *f = pyr.feat[lvl] + (size_t)g.batch * H * W * C + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int iy = 0; iy < g.grid_h; ++iy) {
            const AxisSample y = axis_sample(g.start_h, g.bin_h, g.grid_h, H, ph, iy);
            if (!y.valid) continue;
            for (int ix = 0; ix < g.grid_w; ++ix) {
                const AxisSample x = axis_sample(g.start_w, g.bin_w, g.grid_w, W, pw, ix);
                if (!x.valid) continue;
                const float w1 = __fmul_rn(y.h, x.h), w2 = __fmul_rn(y.h, x.l);
                const float w3 = __fmul_rn(y.l, x.h), w4 = __fmul_rn(y.l, x.l);
                const float4 v1 = ldg4(f + ((size_t)y.low * W + x.low) * C), v2 = ldg4(f + ((size_t)y.low * W + x.high) * C);
                const float4 v3 = ldg4(f + ((size_t)y.high * W + x.low) * C), v4 = ldg4(f + ((size_t)y.high * W + x.high) * C);
                // ((w1v1 + w2v2) + w3v3) + w4v4, then += : the reference's order, unfused
#define FGN_S(q) acc.q = __fadd_rn(acc.q, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1.q), __fmul_rn(w2, v2.q)), \
                                                              __fmul_rn(w3, v3.q)), __fmul_rn(w4, v4.q)))
                FGN_S(x); FGN_S(y); FGN_S(z); FGN_S(w);
#undef FGN_S
            }
        }
        float4 o = make_float4(__fdiv_rn(acc.x, g.count), __fdiv_rn(acc.y, g.count), __fdiv_rn(acc.z, g.count), __fdiv_rn(acc.w, g.count));
        if (chan_scale != nullptr) {
            const int si = scale_index != nullptr ? scale_index[r] : r;
            const float4 cs = ldg4(chan_scale + (size_t)si * C + c);
            o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
        }
        if (out_layout == FGN_LAYOUT_NHWC)
            *reinterpret_cast<float4 *>(out + (((size_t)r * P + ph) * P + pw) * C + c) = o;
        else {
            float *q = out + (((size_t)r * C + c) * P + ph) * P + pw;
            const size_t pp = (size_t)P * P;
            q[0] = o.x; q[pp] = o.y; q[2 * pp] = o.z; q[3 * pp] = o.w;
        }
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
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned, pyr.B);
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
                            const int32_t *scale_index, float *out, int32_t *lvl_out, cudaStream_t st,
                            int ns_pref, bool *taken);
int launch_roi_align_window_bf16(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                                 int aligned, float finest_scale, const float *chan_scale,
                                 const int32_t *scale_index, void *out, int out_is_bf16, int32_t *lvl_out,
                                 cudaStream_t st, int ns_pref, bool *taken);
unsigned int roi_align_window_violations();
void roi_align_window_trace(unsigned long long *dst, int n);

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

extern "C" int fgn_roi_align_ml_fwd(const fgn_pyramid_t *pyr, int B, int C, int in_layout,
                                    const float *rois, int R, int P, int sampling_ratio,
                                    int aligned, float finest_scale, const float *chan_scale,
                                    const int32_t *scale_index, float *out, int out_layout,
                                    int32_t *lvl_out, void *stream)
{
    int rc = validate_pyramid(pyr);
    if (rc) return rc;
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && P > 0, "bad dims R=%d B=%d C=%d P=%d", R, B, C, P);
    FGN_CHECK_ARG(in_layout == FGN_LAYOUT_NCHW || in_layout == FGN_LAYOUT_NHWC, "in_layout=%d", in_layout);
    FGN_CHECK_ARG(out_layout == FGN_LAYOUT_NCHW || out_layout == FGN_LAYOUT_NHWC, "out_layout=%d", out_layout);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && out, "NULL pointer");
    for (int l = 0; l < pyr->num_levels; ++l) FGN_CHECK_ARG(pyr->feat[l], "level %d pointer is NULL", l);
    const Pyramid d = to_device_pyramid(pyr, B);
    cudaStream_t st = (cudaStream_t)stream;
    // FGN_RA_IMPL (development knob): 4 = persistent rotating-window kernel (default, NHWC out),
    // 2 = row-streaming kernel (one CTA per RoI: NCHW output and the shapes the window kernel declines), 0 = direct
    const int impl = env_int("FGN_RA_IMPL", 4);
    // a handful of RoIs (the support branch): the wide small-launch kernel, see roi_align_small_nhwc_kernel
    if (in_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 4 && R <= env_int("FGN_RA_SMALL", 16)) {
        const size_t warps = (size_t)R * P * P * ((C + 127) / 128);
        const int blocks = (int)min((size_t)148 * 8, (warps + 7) / 8);
        roi_align_small_nhwc_kernel<<<blocks, 256, 0, st>>>(d, C, rois, R, P, sampling_ratio, aligned, finest_scale,
                                                            chan_scale, scale_index, out, out_layout, lvl_out);
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    if (in_layout == FGN_LAYOUT_NHWC && out_layout == FGN_LAYOUT_NHWC && (C % 4) == 0 && impl >= 4) {
        bool taken = false;
        rc = launch_roi_align_window(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                     scale_index, out, lvl_out, st, env_int("FGN_RA_NS", 0), &taken);
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
    return roi_align_window_violations();
}

// Development trace of the rotating-window kernel (FGN_RA_DEBUG bit 5); not declared in the public header.
extern "C" void fgn_debug_roi_window_trace(unsigned long long *dst, int n)
{
    roi_align_window_trace(dst, n);
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
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && P > 0, "bad dims R=%d B=%d C=%d P=%d", R, B, C, P);
    FGN_CHECK_ARG(in_layout == FGN_LAYOUT_NCHW || in_layout == FGN_LAYOUT_NHWC, "in_layout=%d", in_layout);
    FGN_CHECK_ARG(out_layout == FGN_LAYOUT_NCHW || out_layout == FGN_LAYOUT_NHWC, "out_layout=%d", out_layout);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && out, "NULL pointer");
    for (int l = 0; l < pyr->num_levels; ++l) FGN_CHECK_ARG(pyr->feat[l], "level %d pointer is NULL", l);
    const Pyramid d = to_device_pyramid(pyr, B);
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
    const Pyramid d = to_device_pyramid(pyr, B);
    // the rotating-window kernel with half-width cells; a handful of RoIs (the support branch) and the shapes it
    // declines stay on the one-CTA-per-RoI kernel (FGN_RA_IMPL=2 forces that one)
    if (env_int("FGN_RA_IMPL", 4) >= 4 && R > env_int("FGN_RA_SMALL", 16)) {
        bool taken = false;
        rc = launch_roi_align_window_bf16(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                          scale_index, out, out_is_bf16, lvl_out, (cudaStream_t)stream,
                                          env_int("FGN_RA_NS", 0), &taken);
        if (rc || taken) return rc;
    }
    return launch_roi_align_stream_bf16(d, C, P, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                        scale_index, out, out_is_bf16, lvl_out, (cudaStream_t)stream);
}

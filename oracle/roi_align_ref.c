/*
 * oracle/roi_align_ref.c -- TEST INFRASTRUCTURE ONLY. Never imported by the product path
 * (fgn_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this.
 *
 * Plain-C, scalar, fp32 restatement of the third-party arithmetic the FGN hot path calls:
 *
 *   - RoIAlign, avg mode.  Call sites in the reference:
 *       subprojects/sp02_omniiseg_fgn_mmdet/fgn_r50_c4_densecl.py:69-73 (mmcv RoIAlign,
 *         output_size=7, sampling_ratio=0, aligned=True, via SingleRoIExtractor, used at
 *         fgn_roi_head.py:331-332 and :366-367)
 *       subprojects/sp02_omniiseg_fgn_mmdet/fgn_roi_head.py:429,432 (torchvision.ops.roi_align,
 *         output_size=7, spatial_scale=1, sampling_ratio=-1, aligned=False)
 *     The algorithm lives in third-party code absent from /root/reference (mmcv-full 1.3.16
 *     mmcv/ops/csrc/pytorch/cpu/roi_align.cpp; torchvision 0.10 csrc/ops/cpu/roi_align_kernel.cpp,
 *     requirements.txt:45,100-101).  Both are the Detectron ROIAlign: same operation order,
 *     restated in SURVEY.md appendix A.1.  This file follows that order exactly, in fp32,
 *     and must be compiled with -ffp-contract=off (no FMA), as the CPU wheels are.
 *   - map_roi_levels (mmdet 2.18 SingleRoIExtractor).
 *
 * Parity pin: tests/test_oracle.py checks this file bit-for-bit (values AND implied indices)
 * against torch.ops.torchvision.roi_align on CPU -- the very op fgn_roi_head.py:429 calls --
 * and against the committed fixtures in tests/golden/.  The mmcv op itself is not installable
 * here, so the aligned=True path is pinned through torchvision's aligned=True mode only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int valid;          /* 0: sample contributes exactly 0 (y<-1 || y>H || x<-1 || x>W) */
    int low, high;      /* clamped integer cell indices along this axis */
    float l, h;         /* l = coord - low ; h = 1 - l */
} axis_sample_t;

/* One axis of the Detectron bilinear pre-calc (SURVEY A.1 "bilinear(y,x)").  `start` is the
 * roi start along the axis, `bin` the bin size, `grid` the sampling grid, `size` the map
 * extent (H or W).  Coordinates are formed exactly as
 *   start + p * bin + (float)(i + .5f) * bin / (float)grid                (left to right)  */
static void axis_sample(float start, float bin, int grid, int size, int p, int i,
                        axis_sample_t *s)
{
    float c = start + (float)p * bin + ((float)i + .5f) * bin / (float)grid;
    s->valid = !(c < -1.0f || c > (float)size);
    if (!s->valid) { s->low = s->high = 0; s->l = s->h = 0.f; return; }   /* also NaN-safe */
    if (c <= 0.0f) c = 0.0f;
    int low = (int)c, high;
    if (low >= size - 1) { high = low = size - 1; c = (float)low; }
    else                 { high = low + 1; }
    s->low = low; s->high = high;
    s->l = c - (float)low;
    s->h = 1.0f - s->l;
}

typedef struct {
    int batch;
    float start_w, start_h, bin_w, bin_h;
    int grid_h, grid_w;
    float count;
} roi_geom_t;

static void roi_geometry(const float *roi, float spatial_scale, int PH, int PW,
                         int sampling_ratio, int aligned, roi_geom_t *g)
{
    float off = aligned ? 0.5f : 0.0f;
    g->batch = (int)roi[0];
    float sw = roi[1] * spatial_scale - off;
    float sh = roi[2] * spatial_scale - off;
    float ew = roi[3] * spatial_scale - off;
    float eh = roi[4] * spatial_scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; }
    g->start_w = sw; g->start_h = sh;
    g->bin_h = rh / (float)PH;
    g->bin_w = rw / (float)PW;
    g->grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    g->grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    int cnt = g->grid_h * g->grid_w;
    g->count = (float)(cnt > 1 ? cnt : 1);
}

/* feat: [B,C,H,W] NCHW fp32 contiguous; rois: [R,5]; out: [R,C,PH,PW]. */
void fgn_oracle_roi_align(const float *feat, int B, int C, int H, int W,
                          const float *rois, int R, float spatial_scale,
                          int PH, int PW, int sampling_ratio, int aligned, float *out)
{
    (void)B;
    for (int r = 0; r < R; ++r) {
        roi_geom_t g;
        roi_geometry(rois + 5 * r, spatial_scale, PH, PW, sampling_ratio, aligned, &g);
        int gh = g.grid_h > 0 ? g.grid_h : 0, gw = g.grid_w > 0 ? g.grid_w : 0;
        axis_sample_t *ys = (axis_sample_t *)malloc(sizeof(axis_sample_t) * (size_t)(PH * gh + 1));
        axis_sample_t *xs = (axis_sample_t *)malloc(sizeof(axis_sample_t) * (size_t)(PW * gw + 1));
        for (int ph = 0; ph < PH; ++ph)
            for (int iy = 0; iy < gh; ++iy)
                axis_sample(g.start_h, g.bin_h, g.grid_h, H, ph, iy, &ys[ph * gh + iy]);
        for (int pw = 0; pw < PW; ++pw)
            for (int ix = 0; ix < gw; ++ix)
                axis_sample(g.start_w, g.bin_w, g.grid_w, W, pw, ix, &xs[pw * gw + ix]);
        for (int c = 0; c < C; ++c) {
            const float *f = feat + ((size_t)g.batch * C + c) * (size_t)H * W;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.f;
                    for (int iy = 0; iy < gh; ++iy) {
                        const axis_sample_t *y = &ys[ph * gh + iy];
                        for (int ix = 0; ix < gw; ++ix) {
                            const axis_sample_t *x = &xs[pw * gw + ix];
                            if (!y->valid || !x->valid) continue;   /* adds exactly 0 */
                            float w1 = y->h * x->h, w2 = y->h * x->l;
                            float w3 = y->l * x->h, w4 = y->l * x->l;
                            float v1 = f[y->low * W + x->low], v2 = f[y->low * W + x->high];
                            float v3 = f[y->high * W + x->low], v4 = f[y->high * W + x->high];
                            acc += w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
                        }
                    }
                    out[(((size_t)r * C + c) * PH + ph) * PW + pw] = acc / g.count;
                }
        }
        free(ys); free(xs);
    }
}

/* Integer side of RoIAlign for the bit-exact check.  Because the sampling grid is a cartesian
 * product, the full (r,ph,pw,iy,ix) index tuple (yl,xl,yh,xh,zero_flag) is the product of the
 * per-axis tables written here:
 *   grid[r] = (gh, gw)
 *   ytab[r][ph][iy] = (valid, low, high)   for iy < min(gh, max_grid), padded with -1
 *   xtab[r][pw][ix] = (valid, low, high)
 * H/W are per-RoI here (hw[r] = (H,W) of the level the RoI is pooled from). */
void fgn_oracle_roi_align_indices(const float *rois, int R, const float *scale_per_roi,
                                  const int32_t *hw, int PH, int PW, int sampling_ratio,
                                  int aligned, int max_grid, int32_t *grid,
                                  int32_t *ytab, int32_t *xtab)
{
    for (int r = 0; r < R; ++r) {
        roi_geom_t g;
        roi_geometry(rois + 5 * r, scale_per_roi[r], PH, PW, sampling_ratio, aligned, &g);
        grid[2 * r] = g.grid_h; grid[2 * r + 1] = g.grid_w;
        int H = hw[2 * r], W = hw[2 * r + 1];
        for (int p = 0; p < PH; ++p)
            for (int i = 0; i < max_grid; ++i) {
                int32_t *t = ytab + (((size_t)r * PH + p) * max_grid + i) * 3;
                if (i < g.grid_h) {
                    axis_sample_t s; axis_sample(g.start_h, g.bin_h, g.grid_h, H, p, i, &s);
                    t[0] = s.valid; t[1] = s.low; t[2] = s.high;
                } else { t[0] = t[1] = t[2] = -1; }
            }
        for (int p = 0; p < PW; ++p)
            for (int i = 0; i < max_grid; ++i) {
                int32_t *t = xtab + (((size_t)r * PW + p) * max_grid + i) * 3;
                if (i < g.grid_w) {
                    axis_sample_t s; axis_sample(g.start_w, g.bin_w, g.grid_w, W, p, i, &s);
                    t[0] = s.valid; t[1] = s.low; t[2] = s.high;
                } else { t[0] = t[1] = t[2] = -1; }
            }
    }
}

/* mmdet 2.18 SingleRoIExtractor.map_roi_levels:
 *   scale = sqrt((x2-x1)*(y2-y1)); lvl = clamp(floor(log2(scale/finest + 1e-6)), 0, L-1)
 * all in fp32.  torch.log2 on fp32 is taken as the correctly rounded fp32 log2 (verified for
 * torch CPU on the values adjacent to 2^k in tests/test_oracle.py), which is computed here as
 * (float)log2((double)v).  NaN scale (negative area) -> floor(NaN) -> .long() is undefined in
 * torch; the contract here clamps it to level 0. */
void fgn_oracle_map_roi_levels(const float *rois, int R, int L, float finest_scale, int32_t *lvl)
{
    for (int r = 0; r < R; ++r) {
        const float *q = rois + 5 * r;
        float area = (q[3] - q[1]) * (q[4] - q[2]);
        float scale = sqrtf(area);
        float v = scale / finest_scale + 1e-6f;
        float lg = (float)log2((double)v);
        float fl = floorf(lg);
        int l;
        if (!(fl >= 0.f)) l = 0;                 /* also catches NaN and -inf */
        else if (fl > (float)(L - 1)) l = L - 1;
        else l = (int)fl;
        lvl[r] = l;
    }
}

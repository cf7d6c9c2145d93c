// support.cu -- support-branch kernels: mask pooling, class mean + masked GAP, AG-RPN class
// attention vectors.  Reference: FGNRoIHead.count_spp (fgn_roi_head.py:419-449) and
// AGRPNHead.forward_single (fgn_ag_rpn_head.py:37-41).
#include "common.cuh"

namespace fgn {

// ---- K2: roi_align(spp_isegmaps.float(), boxes, 7)  (fgn_roi_head.py:429) --------------------
// One CTA per support image.  spatial_scale=1, sampling_ratio=-1 (adaptive), aligned=False,
// one box per image, one channel.  The adaptive grid is ~ceil(0.8*S/7)^2 (~900 samples per bin
// for S=256), so the separable form matters most here:
//   1. every (axis, bin, sample) coordinate is evaluated by its own thread (exact reference
//      arithmetic) into shared memory;
//   2. one thread per (axis, bin) folds its samples, in sample order, into per-cell weights;
//   3. row pass: t[y][pw] = sum_x wx[pw][x] * mask[y][x], one (row, bin) pair per thread;
//   4. column pass: out[ph][pw] = sum_y wy[ph][y] * t[y][pw] / count, one bin per thread.
constexpr int kMaskThreads = 512;

template <int P>
__global__ void __launch_bounds__(kMaskThreads)
support_mask_pool_kernel(const uint8_t *__restrict__ mask, const float *__restrict__ boxes,
                         const int S_h, const int S_w, float *__restrict__ out, const int cap,
                         const int gcap)
{
    extern __shared__ __align__(16) float smem[];
    // layout: wy[cap] | wx[cap] | t[S_h * P] | samples: low[2P*gcap] high[2P*gcap] l[2P*gcap] h[2P*gcap]
    float *wy = smem, *wx = smem + cap, *t = smem + 2 * cap;
    int   *s_low  = reinterpret_cast<int *>(t + (size_t)S_h * P);
    int   *s_high = s_low + 2 * P * gcap;
    float *s_l    = reinterpret_cast<float *>(s_high + 2 * P * gcap);
    float *s_h    = s_l + 2 * P * gcap;
    __shared__ int lo[2][P], n[2][P], off[2][P];
    __shared__ RoiGeom g_s;
    const int m = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        float roi[5] = {0.f, boxes[4 * m], boxes[4 * m + 1], boxes[4 * m + 2], boxes[4 * m + 3]};
        g_s = roi_geometry(roi, 1.0f, P, -1, 0);
    }
    __syncthreads();
    const RoiGeom g = g_s;
    const bool staged = g.grid_h <= gcap && g.grid_w <= gcap;      // else: recompute samples on the fly
    if (staged) {
        for (int i = tid; i < 2 * P * gcap; i += blockDim.x) {
            const int axis = i / (P * gcap), p = (i / gcap) % P, k = i % gcap;
            const int grid = axis ? g.grid_w : g.grid_h;
            if (k < grid) {
                const AxisSample s = axis ? axis_sample(g.start_w, g.bin_w, g.grid_w, S_w, p, k)
                                          : axis_sample(g.start_h, g.bin_h, g.grid_h, S_h, p, k);
                s_low[i] = s.valid ? s.low : -1; s_high[i] = s.high; s_l[i] = s.l; s_h[i] = s.h;
            }
        }
    }
    __syncthreads();
    if (tid < 2 * P) {
        const int axis = tid / P, p = tid % P;
        const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
        const int grid = axis ? g.grid_w : g.grid_h, size = axis ? S_w : S_h;
        int l = 0x7fffffff, h = -1;
        for (int i = 0; i < grid; ++i) {
            int sl, sh;
            if (staged) { sl = s_low[tid * gcap + i]; sh = s_high[tid * gcap + i]; }
            else { const AxisSample s = axis_sample(start, bin, grid, size, p, i); sl = s.valid ? s.low : -1; sh = s.high; }
            if (sl >= 0) { l = min(l, sl); h = max(h, sh); }
        }
        lo[axis][p] = h >= 0 ? l : 0;
        n[axis][p]  = h >= 0 ? h - l + 1 : 0;
    }
    __syncthreads();
    if (tid < 2 * P) {
        const int axis = tid / P, p = tid % P;
        int o = 0;
        for (int q = 0; q < p; ++q) o += n[axis][q];
        off[axis][p] = o;
        float *w = (axis ? wx : wy) + o;
        const int cnt = min(n[axis][p], cap - o);
        for (int i = 0; i < cnt; ++i) w[i] = 0.f;
        const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
        const int grid = axis ? g.grid_w : g.grid_h, size = axis ? S_w : S_h;
        const int l = lo[axis][p];
        for (int i = 0; i < grid; ++i) {
            int sl, sh; float fl, fh;
            if (staged) { sl = s_low[tid * gcap + i]; sh = s_high[tid * gcap + i]; fl = s_l[tid * gcap + i]; fh = s_h[tid * gcap + i]; }
            else { const AxisSample s = axis_sample(start, bin, grid, size, p, i); sl = s.valid ? s.low : -1; sh = s.high; fl = s.l; fh = s.h; }
            if (sl >= 0 && sh - l < cnt) { w[sl - l] += fh; w[sh - l] += fl; }
        }
    }
    __syncthreads();
    int ymin = 0x7fffffff, ymax = 0;
    for (int p = 0; p < P; ++p)
        if (n[0][p] > 0) { ymin = min(ymin, lo[0][p]); ymax = max(ymax, lo[0][p] + n[0][p]); }
    if (ymin == 0x7fffffff) ymin = ymax = 0;
    const uint8_t *mk = mask + (size_t)m * S_h * S_w;
    // row pass: consecutive threads take consecutive bins of the same row -> neighbouring byte runs
    for (int i = tid; i < (ymax - ymin) * P; i += blockDim.x) {
        const int y = ymin + i / P, pw = i % P;
        const uint8_t *row = mk + (size_t)y * S_w + lo[1][pw];
        const float *w = wx + off[1][pw];
        const int nx = n[1][pw];
        float s0 = 0.f, s1 = 0.f;
        int xi = 0;
        for (; xi + 1 < nx; xi += 2) {
            s0 = fmaf(w[xi], row[xi] ? 1.f : 0.f, s0);
            s1 = fmaf(w[xi + 1], row[xi + 1] ? 1.f : 0.f, s1);
        }
        if (xi < nx) s0 = fmaf(w[xi], row[xi] ? 1.f : 0.f, s0);
        t[i] = s0 + s1;
    }
    __syncthreads();
    if (tid < P * P) {
        const int ph = tid / P, pw = tid % P;
        const int yl = lo[0][ph], ny = n[0][ph];
        const float *w = wy + off[0][ph];
        float s = 0.f;
        for (int yi = 0; yi < ny; ++yi) s = fmaf(w[yi], t[(size_t)(yl + yi - ymin) * P + pw], s);
        out[(size_t)m * P * P + tid] = __fdiv_rn(s, g.count);
    }
}

// ---- K4 + K5: class mean and masked GAP (fgn_roi_head.py:439-447) ----------------------------
// One warp per (bn, c): lanes cover the P*P cells, K shots looped; cat_mean is written per
// cell, the masked GAP is a warp-shuffle reduction over (k, cell) divided by K*P*P.
__global__ void __launch_bounds__(256)
support_pool_kernel(const float *__restrict__ f, const int f_layout, const float *__restrict__ m,
                    const int BN, const int K, const int C, const int PP,
                    float *__restrict__ cat_mean, const int out_layout,
                    float *__restrict__ masked_gap)
{
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= BN * C) return;
    int bn, c;
    if (f_layout == FGN_LAYOUT_NCHW) { bn = warp_global / C; c = warp_global % C; }
    else                             { bn = warp_global / C; c = warp_global % C; }
    const float invK = 1.0f / (float)K;
    float gap = 0.f;
    for (int p = lane; p < PP; p += 32) {
        float s = 0.f;
        for (int k = 0; k < K; ++k) {
            const size_t img = (size_t)bn * K + k;
            const float v = f_layout == FGN_LAYOUT_NCHW ? __ldg(f + (img * C + c) * PP + p)
                                                        : __ldg(f + (img * PP + p) * C + c);
            s += v;
            gap = fmaf(v, __ldg(m + img * PP + p), gap);
        }
        const float mean = K == 1 ? s : s * invK;
        if (out_layout == FGN_LAYOUT_NCHW) cat_mean[((size_t)bn * C + c) * PP + p] = mean;
        else                               cat_mean[((size_t)bn * PP + p) * C + c] = mean;
    }
    gap = warp_sum(gap);
    if (lane == 0) masked_gap[(size_t)bn * C + c] = gap / (float)(K * PP);
}

// NHWC variant: one CTA per (bn): threads = (C/4 channel vectors) x (cell groups); each thread walks
// its cells with coalesced 128-bit loads, the masked GAP is reduced over the cell groups in shared
// memory in a fixed order.
__global__ void __launch_bounds__(256)
support_pool_nhwc_kernel(const float *__restrict__ f, const float *__restrict__ m, const int BN,
                         const int K, const int C, const int PP, float *__restrict__ cat_mean,
                         const int out_layout, float *__restrict__ masked_gap)
{
    extern __shared__ __align__(16) float red[];          // [groups][C]
    const int bn = blockIdx.x;
    const int c4 = C >> 2;
    const int lanes = min(c4, (int)blockDim.x);            // channel vectors handled per pass
    const int groups = max(1, (int)blockDim.x / lanes);
    const int cvl = threadIdx.x % lanes, pg = threadIdx.x / lanes;
    const float invK = 1.0f / (float)K;
    for (int cv = cvl; cv < c4; cv += lanes) {
        const int c = cv * 4;
        float4 gap = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pg < groups) {
            for (int p = pg; p < PP; p += groups) {
                float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < K; ++k) {
                    const size_t img = (size_t)bn * K + k;
                    const float4 v = ldg4(f + (img * PP + p) * C + c);
                    const float mv = __ldg(m + img * PP + p);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    fma4(gap, mv, v);
                }
                if (K != 1) { s.x *= invK; s.y *= invK; s.z *= invK; s.w *= invK; }
                if (out_layout == FGN_LAYOUT_NHWC) {
                    *reinterpret_cast<float4 *>(cat_mean + ((size_t)bn * PP + p) * C + c) = s;
                } else {
                    cat_mean[((size_t)bn * C + c + 0) * PP + p] = s.x;
                    cat_mean[((size_t)bn * C + c + 1) * PP + p] = s.y;
                    cat_mean[((size_t)bn * C + c + 2) * PP + p] = s.z;
                    cat_mean[((size_t)bn * C + c + 3) * PP + p] = s.w;
                }
            }
            *reinterpret_cast<float4 *>(red + (size_t)pg * C + c) = gap;
        }
    }
    __syncthreads();
    const float d = (float)(K * PP);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < groups; ++g) s += red[(size_t)g * C + c];
        masked_gap[(size_t)bn * C + c] = s / d;
    }
}

// ---- the whole support branch in one launch (count_spp, fgn_roi_head.py:419-449, + the class term of :272) ----------
// One CTA per (class bn, bin p): everything count_spp produces for that bin is local to it --
//   K2  the mask's pooled value of bin p for each of the K shots: the bin's adaptive sampling grid (~30 x 30 samples of a
//       256-px support) spread over the CTA's threads, per-axis samples (exact reference arithmetic) staged once;
//   K3  the RoIAligned support features of bin p, all C channels, per shot: a warp per (shot, 128 channels) visiting the
//       samples in the reference's order (the arithmetic of roi_align_small_nhwc_kernel, bit-exact against torchvision);
//   K4  the class mean over the shots -> cat_mean[bn, p, :];  K5  this bin's share of the masked GAP;
//   and the row of the relation conv's class term Ys[bn*P*P + p, :] = cat_mean[bn, p, :] Ws^T + bias (exact fp32: a warp
//       per output channel, lanes striding k, fixed-order shuffle reduction) that the relation head otherwise computes
//       with a launch of its own.
// The masked GAP is the one cross-bin quantity: bins leave their shares in workspace and the LAST CTA of a class to
// arrive (a counter per class, zeroed by a memset node in front of the launch) adds them in bin order -- deterministic,
// no spinning, no co-residency assumption.  4 launches (16 + 14 + 11 + 10 us at cfg3) -> 1.
constexpr int kProThreads = 256;
constexpr int kProGcap = 64;                 // staged samples per axis; larger adaptive grids are evaluated on the fly

__global__ void __launch_bounds__(kProThreads)
support_prologue_kernel(const Pyramid pyr, const int C, const float *__restrict__ boxes, const uint8_t *__restrict__ mask,
                        const int S_h, const int S_w, const int K, const int P, const float finest_scale,
                        const float *__restrict__ conv_w, const float *__restrict__ conv_b,
                        float *__restrict__ cat_mean, float *__restrict__ masked_gap, float *__restrict__ class_term,
                        float *__restrict__ gap_part, unsigned int *__restrict__ counters)
{
    extern __shared__ __align__(16) float sm[];
    // layout: f_s[K][C] | cm_s[C] | m_s[K] | red[8] | axis tables: low[2][gcap] high[2][gcap] (int), l[2][gcap] h[2][gcap]
    float *f_s = sm, *cm_s = sm + (size_t)K * C, *m_s = cm_s + C, *red = m_s + K;
    int   *t_low = reinterpret_cast<int *>(red + 8), *t_high = t_low + 2 * kProGcap;
    float *t_l = reinterpret_cast<float *>(t_high + 2 * kProGcap), *t_h = t_l + 2 * kProGcap;
    __shared__ int s_last;
    const int PP = P * P;
    const int bn = blockIdx.x / PP, p = blockIdx.x % PP, ph = p / P, pw = p % P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- K2: pooled mask value of bin (ph, pw) for each shot (spatial_scale 1, adaptive grid, aligned=False) ----
    for (int k = 0; k < K; ++k) {
        const int m = bn * K + k;
        const float roi[5] = {0.f, boxes[4 * m], boxes[4 * m + 1], boxes[4 * m + 2], boxes[4 * m + 3]};
        const RoiGeom g = roi_geometry(roi, 1.0f, P, -1, 0);
        const bool staged = g.grid_h <= kProGcap && g.grid_w <= kProGcap;
        if (staged) {
            for (int i = tid; i < 2 * kProGcap; i += kProThreads) {
                const int axis = i / kProGcap, j = i % kProGcap;
                if (j < (axis ? g.grid_w : g.grid_h)) {
                    const AxisSample a = axis ? axis_sample(g.start_w, g.bin_w, g.grid_w, S_w, pw, j)
                                              : axis_sample(g.start_h, g.bin_h, g.grid_h, S_h, ph, j);
                    t_low[i] = a.valid ? a.low : -1; t_high[i] = a.high; t_l[i] = a.l; t_h[i] = a.h;
                }
            }
        }
        __syncthreads();
        const uint8_t *mk = mask + (size_t)m * S_h * S_w;
        float acc = 0.f;
        const int ns = g.grid_h * g.grid_w;
        for (int i = tid; i < ns; i += kProThreads) {
            const int iy = i / g.grid_w, ix = i % g.grid_w;
            int yl, yh, xl, xh; float fyl, fyh, fxl, fxh;
            if (staged) {
                yl = t_low[iy]; yh = t_high[iy]; fyl = t_l[iy]; fyh = t_h[iy];
                xl = t_low[kProGcap + ix]; xh = t_high[kProGcap + ix]; fxl = t_l[kProGcap + ix]; fxh = t_h[kProGcap + ix];
            } else {
                const AxisSample y = axis_sample(g.start_h, g.bin_h, g.grid_h, S_h, ph, iy);
                const AxisSample x = axis_sample(g.start_w, g.bin_w, g.grid_w, S_w, pw, ix);
                yl = y.valid ? y.low : -1; yh = y.high; fyl = y.l; fyh = y.h;
                xl = x.valid ? x.low : -1; xh = x.high; fxl = x.l; fxh = x.h;
            }
            if (yl < 0 || xl < 0) continue;
            const float v1 = mk[(size_t)yl * S_w + xl] ? 1.f : 0.f, v2 = mk[(size_t)yl * S_w + xh] ? 1.f : 0.f;
            const float v3 = mk[(size_t)yh * S_w + xl] ? 1.f : 0.f, v4 = mk[(size_t)yh * S_w + xh] ? 1.f : 0.f;
            acc += __fmul_rn(fyh, fxh) * v1 + __fmul_rn(fyh, fxl) * v2 + __fmul_rn(fyl, fxh) * v3 + __fmul_rn(fyl, fxl) * v4;
        }
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < kProThreads / 32; ++w) s += red[w];
            m_s[k] = __fdiv_rn(s, g.count);
        }
        __syncthreads();
    }

    // ---- K3: RoIAligned support features of this bin, per shot, all channels ----
    // The bin's per-axis samples are evaluated one per lane (exact reference arithmetic) and handed round by shuffles, so
    // the (iy, ix) loop is nothing but independent 128-bit corner loads -- visited and added in the reference's order.
    const int nblk = (C + 127) >> 7;
    for (int t = warp; t < K * nblk; t += kProThreads / 32) {
        const int k = t / nblk, c = min((t % nblk) * 128 + lane * 4, C - 4);   // (lanes past C recompute the last vector)
        const int m = bn * K + k;
        const float roi[5] = {(float)m, boxes[4 * m], boxes[4 * m + 1], boxes[4 * m + 2], boxes[4 * m + 3]};
        const int lvl = roi_level(roi, pyr, finest_scale);
        const RoiGeom g = roi_geometry(roi, pyr.scale[lvl], P, -1, 0, pyr.B);
        const int H = pyr.H[lvl], W = pyr.W[lvl];
        const float *f = pyr.feat[lvl] + (size_t)g.batch * H * W * C + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#define FGN_S(q) acc.q = __fadd_rn(acc.q, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1.q), __fmul_rn(w2, v2.q)), \
                                                              __fmul_rn(w3, v3.q)), __fmul_rn(w4, v4.q)))
        if (g.grid_h <= 32 && g.grid_w <= 32) {
            const AxisSample ys = axis_sample(g.start_h, g.bin_h, g.grid_h, H, ph, min(lane, max(g.grid_h - 1, 0)));
            const AxisSample xs = axis_sample(g.start_w, g.bin_w, g.grid_w, W, pw, min(lane, max(g.grid_w - 1, 0)));
            for (int iy = 0; iy < g.grid_h; ++iy) {
                const int yv = __shfl_sync(0xffffffffu, ys.valid, iy), ylo = __shfl_sync(0xffffffffu, ys.low, iy);
                const int yhi = __shfl_sync(0xffffffffu, ys.high, iy);
                const float yl = __shfl_sync(0xffffffffu, ys.l, iy), yh = __shfl_sync(0xffffffffu, ys.h, iy);
                if (!yv) continue;
#pragma unroll 4
                for (int ix = 0; ix < g.grid_w; ++ix) {
                    const int xv = __shfl_sync(0xffffffffu, xs.valid, ix), xlo = __shfl_sync(0xffffffffu, xs.low, ix);
                    const int xhi = __shfl_sync(0xffffffffu, xs.high, ix);
                    const float xl = __shfl_sync(0xffffffffu, xs.l, ix), xh = __shfl_sync(0xffffffffu, xs.h, ix);
                    if (!xv) continue;
                    const float w1 = __fmul_rn(yh, xh), w2 = __fmul_rn(yh, xl), w3 = __fmul_rn(yl, xh), w4 = __fmul_rn(yl, xl);
                    const float4 v1 = ldg4(f + ((size_t)ylo * W + xlo) * C), v2 = ldg4(f + ((size_t)ylo * W + xhi) * C);
                    const float4 v3 = ldg4(f + ((size_t)yhi * W + xlo) * C), v4 = ldg4(f + ((size_t)yhi * W + xhi) * C);
                    FGN_S(x); FGN_S(y); FGN_S(z); FGN_S(w);
                }
            }
        } else {
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
                    FGN_S(x); FGN_S(y); FGN_S(z); FGN_S(w);
                }
            }
        }
#undef FGN_S
        *reinterpret_cast<float4 *>(f_s + (size_t)k * C + c) =
            make_float4(__fdiv_rn(acc.x, g.count), __fdiv_rn(acc.y, g.count), __fdiv_rn(acc.z, g.count), __fdiv_rn(acc.w, g.count));
    }
    __syncthreads();

    // ---- K4 + K5: class mean of the bin, and the bin's share of the masked GAP ----
    const float invK = 1.0f / (float)K;
    const size_t row = (size_t)bn * PP + p;
    for (int c = tid; c < C; c += kProThreads) {
        float s = 0.f, gap = 0.f;
        for (int k = 0; k < K; ++k) {
            const float v = f_s[(size_t)k * C + c];
            s += v;
            gap = fmaf(v, m_s[k], gap);
        }
        const float mean = K == 1 ? s : s * invK;
        cm_s[c] = mean;
        cat_mean[row * C + c] = mean;
        gap_part[row * C + c] = gap;
    }
    __syncthreads();

    // ---- the class term's row: Ys[row, n] = sum_c cat_mean[row, c] * conv_w[n, C + c] + conv_b[n] ----
    if (class_term != nullptr) {
        // a warp takes 8 output channels at a time: their weight vectors are all in flight before the first FMA
        constexpr int NB = 8, NW = kProThreads / 32;
        for (int n0 = warp * NB; n0 < C; n0 += NW * NB) {
            float acc[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int k4 = lane * 4; k4 < C; k4 += 256) {              // two 128-wide k steps per trip: 16 vectors in flight
                float4 w[2][NB];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        w[h][j] = (n0 + j < C && k4 + 128 * h < C) ? ldg4(conv_w + (size_t)(n0 + j) * 2 * C + C + k4 + 128 * h)
                                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (k4 + 128 * h >= C) break;
                    const float4 a = *reinterpret_cast<const float4 *>(cm_s + k4 + 128 * h);
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        acc[j] = fmaf(a.x, w[h][j].x, acc[j]); acc[j] = fmaf(a.y, w[h][j].y, acc[j]);
                        acc[j] = fmaf(a.z, w[h][j].z, acc[j]); acc[j] = fmaf(a.w, w[h][j].w, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = warp_sum(acc[j]);
            if (lane < NB && n0 + lane < C) {
                float o = acc[0];
#pragma unroll
                for (int j = 1; j < NB; ++j) o = lane == j ? acc[j] : o;
                class_term[row * C + n0 + lane] = o + __ldg(conv_b + n0 + lane);
            }
        }
    }

    // ---- masked GAP: the last CTA of the class adds the bins' shares in bin order ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&counters[bn], 1u) == (unsigned)(PP - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float d = (float)(K * PP);
        for (int c = tid; c < C; c += kProThreads) {
            float s = 0.f;
            for (int q0 = 0; q0 < PP; q0 += 8) {                  // eight loads in flight, added in bin order
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = q0 + j < PP ? __ldcg(gap_part + ((size_t)bn * PP + q0 + j) * C + c) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (q0 + j < PP) s += v[j];
            }
            masked_gap[(size_t)bn * C + c] = s / d;
        }
    }
}

// ---- K6: AG-RPN class attention vector (fgn_ag_rpn_head.py:37-41) ----------------------------
// vec[bn,c] = mean over (k,h,w).  NCHW: one warp per (bn,c) streams K contiguous h*w planes
// with 128-bit loads where alignment allows and shuffle-reduces.  NHWC: two deterministic
// stages (per-slab partial sums, then a finalize pass) so no atomics are needed.
__global__ void __launch_bounds__(256)
attention_vec_nchw_kernel(const float *__restrict__ x, const int BN, const int K, const int C,
                          const int HW, float *__restrict__ vec)
{
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= BN * C) return;
    const int bn = warp_global / C, c = warp_global % C;
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
        const float *p = x + (((size_t)bn * K + k) * C + c) * HW;
        if ((HW & 3) == 0) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = lane * 4; i < HW; i += 128) {
                const float4 v = ldg4(p + i);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            s += (a.x + a.y) + (a.z + a.w);
        } else {
            for (int i = lane; i < HW; i += 32) s += __ldg(p + i);
        }
    }
    s = warp_sum(s);
    if (lane == 0) vec[(size_t)bn * C + c] = s / (float)((size_t)K * HW);
}

constexpr int kVecSlab = 64;   // pixels per partial-sum CTA in the NHWC path

__global__ void __launch_bounds__(256)
attention_vec_nhwc_partial_kernel(const float *__restrict__ x, const int BN, const int K,
                                  const int C, const int HW, const int slabs,
                                  float *__restrict__ partial)
{
    // grid: (slab, bn); block: 256 threads = (C/4 channel vectors) x pixel lanes
    const int bn = blockIdx.y, slab = blockIdx.x;
    const int c4 = C >> 2;
    const int total = K * HW;                    // K consecutive NHWC images form one run
    const int p0 = slab * kVecSlab, p1 = min(total, p0 + kVecSlab);
    extern __shared__ __align__(16) float red[];   // [rows][C]
    const int rows = blockDim.x / c4 > 0 ? blockDim.x / c4 : 1;
    const float *base = x + (size_t)bn * K * HW * C;
    for (int cv = threadIdx.x % c4, rowi = threadIdx.x / c4; cv < c4 && rowi < rows; cv += c4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0 + rowi; p < p1; p += rows) {
            const float4 v = ldg4(base + (size_t)p * C + cv * 4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        *reinterpret_cast<float4 *>(red + (size_t)rowi * C + cv * 4) = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s += red[(size_t)rr * C + c];
        partial[((size_t)bn * slabs + slab) * C + c] = s;
    }
}

__global__ void attention_vec_finalize_kernel(const float *__restrict__ partial, const int BN,
                                              const int C, const int slabs, const float inv_n,
                                              float *__restrict__ vec)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN * C) return;
    const int bn = idx / C, c = idx % C;
    float s = 0.f;
    for (int i = 0; i < slabs; ++i) s += partial[((size_t)bn * slabs + i) * C + c];
    vec[idx] = s * inv_n;
}

}  // namespace fgn

using namespace fgn;

extern "C" int fgn_support_mask_pool(const uint8_t *mask, const float *boxes, int M, int S_h,
                                     int S_w, int P, float *out, void *stream)
{
    FGN_CHECK_ARG(M >= 0 && S_h > 0 && S_w > 0, "bad dims M=%d S=%dx%d", M, S_h, S_w);
    if (M == 0) return FGN_OK;
    FGN_CHECK_ARG(mask && boxes && out, "NULL pointer");
    if (P != 7 && P != 14) { set_error("support_mask_pool: P=%d not instantiated (7, 14)", P); return FGN_ERR_UNSUPPORTED; }
    const int cap = ((max(S_h, S_w) + 6 * P + 16) + 3) & ~3;
    const int gcap = 64;                                          // staged samples per bin (adaptive grid <= 64)
    const size_t smem = ((size_t)2 * cap + (size_t)S_h * P + (size_t)4 * 2 * P * gcap) * 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 7) {
        FGN_SMEM_OPTIN(support_mask_pool_kernel<7>, smem);
        support_mask_pool_kernel<7><<<M, kMaskThreads, smem, st>>>(mask, boxes, S_h, S_w, out, cap, gcap);
    } else {
        FGN_SMEM_OPTIN(support_mask_pool_kernel<14>, smem);
        support_mask_pool_kernel<14><<<M, kMaskThreads, smem, st>>>(mask, boxes, S_h, S_w, out, cap, gcap);
    }
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_support_pool(const float *f, int f_layout, const float *m, int BN, int K, int C,
                                int P, float *cat_mean, int out_layout, float *masked_gap,
                                void *stream)
{
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0 && P > 0, "bad dims BN=%d K=%d C=%d P=%d", BN, K, C, P);
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(f && m && cat_mean && masked_gap, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (f_layout == FGN_LAYOUT_NHWC && (C & 3) == 0) {
        const int c4 = C >> 2, lanes = min(c4, 256), groups = max(1, 256 / lanes);
        const size_t smem = (size_t)groups * C * 4;
        if (smem > 48 * 1024) { set_error("support_pool: C=%d too large for the NHWC kernel", C); return FGN_ERR_UNSUPPORTED; }
        support_pool_nhwc_kernel<<<BN, 256, smem, st>>>(f, m, BN, K, C, P * P, cat_mean, out_layout, masked_gap);
    } else {
        const long warps = (long)BN * C;
        support_pool_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(
            f, f_layout, m, BN, K, C, P * P, cat_mean, out_layout, masked_gap);
    }
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" size_t fgn_attention_vectors_workspace_bytes(int BN, int K, int C, int H, int W, int layout)
{
    if (layout != FGN_LAYOUT_NHWC) return 0;
    const long total = (long)K * H * W;
    const long slabs = (total + kVecSlab - 1) / kVecSlab;
    return (size_t)BN * slabs * C * sizeof(float);
}

extern "C" int fgn_attention_vectors(const float *x, int layout, int BN, int K, int C, int H, int W,
                                     float *vec, void *workspace, size_t workspace_bytes,
                                     void *stream)
{
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0 && H > 0 && W > 0, "bad dims");
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(x && vec, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    if (layout == FGN_LAYOUT_NCHW || (C & 3) != 0 || C > 1024) {
        if (layout == FGN_LAYOUT_NHWC) { set_error("attention_vectors NHWC needs C%%4==0 and C<=1024 (C=%d)", C); return FGN_ERR_UNSUPPORTED; }
        const long warps = (long)BN * C;
        attention_vec_nchw_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(x, BN, K, C, HW, vec);
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    const size_t need = fgn_attention_vectors_workspace_bytes(BN, K, C, H, W, layout);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("attention_vectors: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    const int total = K * HW;
    const int slabs = ceil_div(total, kVecSlab);
    const int c4 = C >> 2;
    const int threads = 256;
    const int rows = max(1, threads / c4);
    const size_t smem = (size_t)rows * C * 4;
    dim3 grid(slabs, BN);
    attention_vec_nhwc_partial_kernel<<<grid, threads, smem, st>>>(x, BN, K, C, HW, slabs,
                                                                  (float *)workspace);
    FGN_LAUNCH_OK();
    attention_vec_finalize_kernel<<<ceil_div(BN * C, 256), 256, 0, st>>>(
        (const float *)workspace, BN, C, slabs, 1.0f / (float)total, vec);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

static size_t align256s(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t fgn_support_prologue_workspace_bytes(int BN, int C, int P)
{
    if (BN <= 0 || C <= 0 || P <= 0) return 0;
    return align256s((size_t)BN * P * P * C * sizeof(float)) + align256s((size_t)BN * sizeof(unsigned int));
}

extern "C" int fgn_support_prologue_fwd(const fgn_pyramid_t *spp, int C, const float *boxes, const uint8_t *masks,
                                        int S_h, int S_w, int BN, int K, int P, float finest_scale,
                                        const float *conv_w, const float *conv_b, float *cat_mean, float *masked_gap,
                                        float *class_term, void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(spp != nullptr && spp->num_levels >= 1 && spp->num_levels <= FGN_MAX_LEVELS, "bad pyramid");
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0 && P > 0 && S_h > 0 && S_w > 0, "bad dims BN=%d K=%d C=%d P=%d S=%dx%d", BN, K, C, P, S_h, S_w);
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(boxes && masks && cat_mean && masked_gap, "NULL pointer");
    FGN_CHECK_ARG(class_term == nullptr || (conv_w != nullptr && conv_b != nullptr), "class_term needs conv_w and conv_b");
    for (int l = 0; l < spp->num_levels; ++l) FGN_CHECK_ARG(spp->feat[l], "level %d pointer is NULL", l);
    if ((C & 3) != 0) { set_error("support_prologue: C=%d must be a multiple of 4", C); return FGN_ERR_UNSUPPORTED; }
    const size_t smem = ((size_t)K * C + C + K + 8 + 8 * kProGcap) * sizeof(float);
    if (smem > 200 * 1024) { set_error("support_prologue: K*C = %d*%d does not fit shared memory", K, C); return FGN_ERR_UNSUPPORTED; }
    const size_t need = fgn_support_prologue_workspace_bytes(BN, C, P);
    if (!workspace || workspace_bytes < need) {
        set_error("support_prologue: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    float *gap_part = (float *)workspace;
    unsigned int *counters = (unsigned int *)((char *)workspace + align256s((size_t)BN * P * P * C * sizeof(float)));
    cudaStream_t st = (cudaStream_t)stream;
    FGN_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)BN * sizeof(unsigned int), st));
    const Pyramid d = to_device_pyramid(spp, BN * K);
    FGN_SMEM_OPTIN(support_prologue_kernel, smem);
    support_prologue_kernel<<<BN * P * P, kProThreads, smem, st>>>(d, C, boxes, masks, S_h, S_w, K, P, finest_scale, conv_w, conv_b,
                                                                 cat_mean, masked_gap, class_term, gap_part, counters);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

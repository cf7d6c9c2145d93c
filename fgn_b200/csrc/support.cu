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

// backward.cu -- gradient kernels of the hot path: what the training
// configs need so that the guided heads stay a drop-in under autograd.  The reference gets these
// gradients from autograd over mmcv/torchvision/ATen ops (fgn_roi_head.py:344-358,451-529); here each
// forward entry point has a hand-written adjoint.  All feature tensors are NHWC (channels_last).
#include "common.cuh"

namespace fgn {

// ---- RoIAlign backward ----------------------------------------------------------------------------
// Adjoint of out[r,ph,pw,c] = cs[c]/count * sum_y Ay[ph][y] sum_x Ax[pw][x] v[y,x,c]:
//   grad_v[y,x,c] += cs[c]/count * Ay[ph][y] * Ax[pw][x] * g[r,ph,pw,c]
// with exactly the forward's per-axis weights (same device functions, same sample indices).  One CTA
// per (RoI, 128-channel block); warp = bin row ph; lanes hold 4 channels; 128-bit vector atomics
// (the reference's mmcv backward is an atomicAdd scatter as well, so summation order is unspecified
// on both sides).
constexpr int kBwdMaxP = 16;

struct BwdPlan {
    int   level, batch, H, W;
    float count;
    int   lo[2][kBwdMaxP], n[2][kBwdMaxP], off[2][kBwdMaxP];
    int   overflow;
};

__global__ void absmax_bits_kernel(const float *__restrict__ x, const size_t n, unsigned int *__restrict__ out)
{
    unsigned int m = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(x[i])) & 0x7fffffffu);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(out, m);       // max of non-negative float bit patterns: order-independent
}

// DET: deterministic mode.  Float atomics make the sum depend on the order the RoIs' CTAs happen to run in; here every
// contribution is rounded once to a 64-bit fixed-point number (value * 2^e, e chosen from max |grad_out| and max |chan_scale|
// so that the largest term sits near 2^40: 2^-40 of it in resolution, 2^22 terms of headroom) and added with integer atomics
// -- integer addition is associative, so the result is bit-identical from run to run whatever the schedule.  feat[l] then
// points at the level's int64 accumulators; fgn_roi_align_ml_bwd_det converts and adds them to the gradient maps.
struct DetScale { unsigned int max_g, max_cs; };           // float bits of the maxima (non-negative floats order like uints)

__device__ __forceinline__ double det_multiplier(const DetScale *det)
{
    // 2^e with 2^e * (max|g| * max(1, max|cs|)) in [2^39, 2^40]: exponent arithmetic only, the same on every thread
    float top = __uint_as_float(det->max_g);
    const float cs = __uint_as_float(det->max_cs);
    if (cs > 1.f) top *= cs;
    if (!(top > 0.f) || !(top < 3.0e38f)) return 1.0;                // all-zero or non-finite gradients: nothing to scale
    int ex;
    frexpf(top, &ex);                                                // top = m * 2^ex, m in [0.5, 1)
    return ldexp(1.0, 40 - ex);
}

__global__ void det_finish_kernel(const long long *__restrict__ acc, float *__restrict__ grad, const size_t n, const DetScale *__restrict__ det)
{
    const double inv = 1.0 / det_multiplier(det);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const long long a = acc[i];
        if (a != 0) grad[i] += (float)((double)a * inv);
    }
}

template <int P, bool DET>
__global__ void __launch_bounds__(P * 32)
roi_align_bwd_kernel(const Pyramid pyr /* feat[l] = gradient buffer of level l, accumulated into */, const int C,
                     const float *__restrict__ rois, const int R, const int sampling_ratio, const int aligned,
                     const float finest_scale, const float *__restrict__ chan_scale,
                     const int32_t *__restrict__ scale_index, const float *__restrict__ grad_out,
                     const int wtab_cap, const DetScale *__restrict__ det = nullptr)
{
    extern __shared__ __align__(16) float wtab[];
    __shared__ BwdPlan plan;
    __shared__ RoiGeom g_s;
    const int nblk = (C + 127) / 128;
    const int r = blockIdx.x / nblk, cb0 = (blockIdx.x % nblk) * 128;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
        const float *roi = rois + 5 * (size_t)r;
        const int lvl = roi_level(roi, pyr, finest_scale);
        g_s = roi_geometry(roi, pyr.scale[lvl], P, sampling_ratio, aligned, pyr.B);
        plan.level = lvl; plan.batch = g_s.batch; plan.H = pyr.H[lvl]; plan.W = pyr.W[lvl];
        plan.count = g_s.count; plan.overflow = 0;
    }
    __syncthreads();
    const RoiGeom g = g_s;
    if (t < 2 * P) {
        const int axis = t / P, p = t % P;
        const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
        const int grid = axis ? g.grid_w : g.grid_h, size = axis ? plan.W : plan.H;
        int lo = 0x7fffffff, hi = -1;
        for (int i = 0; i < grid; ++i) {
            const AxisSample s = axis_sample(start, bin, grid, size, p, i);
            if (s.valid) { lo = min(lo, s.low); hi = max(hi, s.high); }
        }
        plan.lo[axis][p] = hi >= 0 ? lo : 0;
        plan.n[axis][p]  = hi >= 0 ? hi - lo + 1 : 0;
    }
    __syncthreads();
    if (t < 2 * P) {
        const int axis = t / P, p = t % P;
        int off = 0;
        for (int a = 0; a <= axis; ++a)
            for (int q = 0; q < (a == axis ? p : P); ++q) off += plan.n[a][q];
        plan.off[axis][p] = off;
        const int n = plan.n[axis][p];
        if (off + n > wtab_cap) plan.overflow = 1;
        else {
            float *w = wtab + off;
            for (int i = 0; i < n; ++i) w[i] = 0.f;
            const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
            const int grid = axis ? g.grid_w : g.grid_h, size = axis ? plan.W : plan.H;
            const int lo = plan.lo[axis][p];
            for (int i = 0; i < grid; ++i) {
                const AxisSample s = axis_sample(start, bin, grid, size, p, i);
                if (s.valid) { w[s.low - lo] += s.h; w[s.high - lo] += s.l; }
            }
        }
    }
    __syncthreads();
    const int c = cb0 + lane * 4;
    if (c >= C || plan.overflow) return;
    const int ph = warp;
    const int W = plan.W;
    float *gbase = const_cast<float *>(pyr.feat[plan.level]) + (size_t)plan.batch * plan.H * W * C + c;
    float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
    if (chan_scale != nullptr) {
        const int si = scale_index != nullptr ? scale_index[r] : r;
        cs = ldg4(chan_scale + (size_t)si * C + c);
    }
    const float inv = 1.0f / plan.count;
    double det_mul = 1.0;
    if (DET) det_mul = det_multiplier(det);
    float4 gv[P];
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
        const float4 x = ldg4(grad_out + (((size_t)r * P + ph) * P + pw) * C + c);
        gv[pw] = make_float4(x.x * inv * cs.x, x.y * inv * cs.y, x.z * inv * cs.z, x.w * inv * cs.w);
    }
    const int ylo = plan.lo[0][ph], ny = plan.n[0][ph];
    const float *wy = wtab + plan.off[0][ph];
    for (int yi = 0; yi < ny; ++yi) {
        const float wyv = wy[yi];
        if (wyv == 0.f) continue;
        float *row = gbase + ((size_t)(ylo + yi) * W) * C;
#pragma unroll
        for (int pw = 0; pw < P; ++pw) {
            const int xlo = plan.lo[1][pw], nx = plan.n[1][pw];
            const float *wx = wtab + plan.off[1][pw];
            for (int xi = 0; xi < nx; ++xi) {
                const float w = wyv * wx[xi];
                if (w == 0.f) continue;
                if (DET) {
                    unsigned long long *acc = reinterpret_cast<unsigned long long *>(const_cast<float *>(pyr.feat[plan.level])) +
                                              ((size_t)plan.batch * plan.H * W + (size_t)(ylo + yi) * W + (xlo + xi)) * C + c;
                    atomicAdd(acc + 0, (unsigned long long)__double2ll_rn((double)(w * gv[pw].x) * det_mul));
                    atomicAdd(acc + 1, (unsigned long long)__double2ll_rn((double)(w * gv[pw].y) * det_mul));
                    atomicAdd(acc + 2, (unsigned long long)__double2ll_rn((double)(w * gv[pw].z) * det_mul));
                    atomicAdd(acc + 3, (unsigned long long)__double2ll_rn((double)(w * gv[pw].w) * det_mul));
                } else
                atomicAdd(reinterpret_cast<float4 *>(row + (size_t)(xlo + xi) * C),
                          make_float4(w * gv[pw].x, w * gv[pw].y, w * gv[pw].z, w * gv[pw].w));
            }
        }
    }
}

// ---- channel attention backward -----------------------------------------------------------------
// out[bn,p,c] = q[b,p,c] * v[bn,c]  =>  grad_q[b,p,c] = sum_n g[bn,p,c] v[bn,c] ;  grad_v[bn,c] = sum_p g[bn,p,c] q[b,p,c]
__global__ void __launch_bounds__(256)
channel_attention_bwd_q_kernel(const float *__restrict__ g, const float *__restrict__ vec, const int B, const int N,
                               const int C, const size_t HW, float *__restrict__ grad_q)
{
    const int c4 = C >> 2;
    const size_t per_img = HW * c4, total = (size_t)B * per_img;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / per_img;
        const size_t rem = i % per_img;
        const int cv = rem % c4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N; ++n) {
            const float4 gg = ldg4(g + (((size_t)b * N + n) * per_img + rem) * 4);
            const float4 s = ldg4(vec + ((size_t)b * N + n) * C + cv * 4);
            a.x = fmaf(gg.x, s.x, a.x); a.y = fmaf(gg.y, s.y, a.y); a.z = fmaf(gg.z, s.z, a.z); a.w = fmaf(gg.w, s.w, a.w);
        }
        *reinterpret_cast<float4 *>(grad_q + 4 * i) = a;
    }
}

constexpr int kBwdSlab = 64;

// partial[(bn*slabs + slab)][C] = sum over the slab's pixels of g*q ; finalized in slab order
__global__ void __launch_bounds__(256)
channel_attention_bwd_v_partial_kernel(const float *__restrict__ g, const float *__restrict__ q, const int N,
                                       const int C, const int HW, const int slabs, float *__restrict__ partial)
{
    const int bn = blockIdx.y, slab = blockIdx.x, b = bn / N;
    const int c4 = C >> 2;
    const int p0 = slab * kBwdSlab, p1 = min(HW, p0 + kBwdSlab);
    extern __shared__ __align__(16) float red[];
    const int rows = max(1, (int)blockDim.x / c4);
    const int cv = threadIdx.x % c4, rowi = threadIdx.x / c4;
    if (rowi < rows && threadIdx.x < rows * c4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0 + rowi; p < p1; p += rows) {
            const float4 gg = ldg4(g + ((size_t)bn * HW + p) * C + cv * 4);
            const float4 qq = ldg4(q + ((size_t)b * HW + p) * C + cv * 4);
            a.x = fmaf(gg.x, qq.x, a.x); a.y = fmaf(gg.y, qq.y, a.y); a.z = fmaf(gg.z, qq.z, a.z); a.w = fmaf(gg.w, qq.w, a.w);
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

__global__ void sum_slabs_kernel(const float *__restrict__ partial, const int BN, const int C, const int slabs,
                                 const float scale, float *__restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN * C) return;
    const int bn = idx / C, c = idx % C;
    float s = 0.f;
    for (int i = 0; i < slabs; ++i) s += partial[((size_t)bn * slabs + i) * C + c];
    out[idx] = s * scale;
}

// ---- attention vectors backward: vec = mean_{k,p} x  =>  grad_x[(bn,k),p,c] = grad_vec[bn,c] / (K*HW)
__global__ void __launch_bounds__(256)
attention_vectors_bwd_kernel(const float *__restrict__ grad_vec, const int BN, const int K, const int C,
                             const size_t HW, float *__restrict__ grad_x)
{
    const int c4 = C >> 2;
    const size_t per_bn = (size_t)K * HW * c4, total = (size_t)BN * per_bn;
    const float s = 1.0f / (float)((size_t)K * HW);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int bn = i / per_bn, cv = i % c4;
        const float4 gvv = ldg4(grad_vec + (size_t)bn * C + cv * 4);
        __stcs(reinterpret_cast<float4 *>(grad_x) + i, make_float4(gvv.x * s, gvv.y * s, gvv.z * s, gvv.w * s));
    }
}

// ---- support pool backward: cat = mean_k f ; gap = sum_{k,p} f*m / (K*PP)
//   grad_f[(bn,k),p,c] = grad_cat[bn,p,c]/K + grad_gap[bn,c] * m[(bn,k),p] / (K*PP)
__global__ void __launch_bounds__(256)
support_pool_bwd_kernel(const float *__restrict__ grad_cat, const float *__restrict__ grad_gap,
                        const float *__restrict__ m, const int BN, const int K, const int C, const int PP,
                        float *__restrict__ grad_f)
{
    const int c4 = C >> 2;
    const size_t total = (size_t)BN * K * PP * c4;
    const float invK = 1.0f / (float)K, invKP = 1.0f / (float)(K * PP);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = i % c4;
        const size_t rest = i / c4;
        const int p = rest % PP;
        const size_t img = rest / PP;                // bn*K + k
        const int bn = img / K;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grad_cat != nullptr) {
            const float4 gc = ldg4(grad_cat + ((size_t)bn * PP + p) * C + cv * 4);
            o = make_float4(gc.x * invK, gc.y * invK, gc.z * invK, gc.w * invK);
        }
        if (grad_gap != nullptr) {
            const float4 gg = ldg4(grad_gap + (size_t)bn * C + cv * 4);
            const float mv = __ldg(m + img * PP + p) * invKP;
            o.x = fmaf(gg.x, mv, o.x); o.y = fmaf(gg.y, mv, o.y); o.z = fmaf(gg.z, mv, o.z); o.w = fmaf(gg.w, mv, o.w);
        }
        *reinterpret_cast<float4 *>(grad_f + 4 * i) = o;
    }
}

}  // namespace fgn

using namespace fgn;

extern "C" int fgn_roi_align_ml_bwd(const fgn_pyramid_t *grad_pyr, int B, int C,
                                    const float *rois, int R, int P, int sampling_ratio, int aligned,
                                    float finest_scale, const float *chan_scale, const int32_t *scale_index,
                                    const float *grad_out, void *stream)
{
    FGN_CHECK_ARG(grad_pyr != nullptr && grad_pyr->num_levels >= 1 && grad_pyr->num_levels <= FGN_MAX_LEVELS, "bad pyramid");
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && (C & 3) == 0, "roi_align_bwd needs C%%4==0 (C=%d)", C);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && grad_out, "NULL pointer");
    for (int l = 0; l < grad_pyr->num_levels; ++l) FGN_CHECK_ARG(grad_pyr->feat[l], "gradient level %d is NULL", l);
    if (P != 7 && P != 14) { set_error("roi_align_bwd: P=%d not instantiated (7, 14)", P); return FGN_ERR_UNSUPPORTED; }
    const Pyramid d = to_device_pyramid(grad_pyr, B);
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int cap = (maxH + maxW + 6 * P + 16 + 3) & ~3;
    const int nblk = (C + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 7) roi_align_bwd_kernel<7, false><<<R * nblk, 7 * 32, (size_t)cap * 4, st>>>(d, C, rois, R, sampling_ratio, aligned,
                                                                                        finest_scale, chan_scale, scale_index, grad_out, cap);
    else        roi_align_bwd_kernel<14, false><<<R * nblk, 14 * 32, (size_t)cap * 4, st>>>(d, C, rois, R, sampling_ratio, aligned,
                                                                                          finest_scale, chan_scale, scale_index, grad_out, cap);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

static size_t det_level_elems(const fgn_pyramid_t *p, int B, int C, int l) { return (size_t)B * p->H[l] * p->W[l] * C; }

extern "C" size_t fgn_roi_align_ml_bwd_det_workspace_bytes(const fgn_pyramid_t *grad_pyr, int B, int C)
{
    if (!grad_pyr || B <= 0 || C <= 0) return 0;
    size_t n = 0;
    for (int l = 0; l < grad_pyr->num_levels && l < FGN_MAX_LEVELS; ++l) n += det_level_elems(grad_pyr, B, C, l);
    return 256 + n * sizeof(long long);
}

// Deterministic form of fgn_roi_align_ml_bwd (same arguments + workspace): bit-identical gradients from run to run.
// chan_scale_rows: rows of chan_scale ([rows, C]; 0 when chan_scale is NULL).
extern "C" int fgn_roi_align_ml_bwd_det(const fgn_pyramid_t *grad_pyr, int B, int C, const float *rois, int R, int P,
                                        int sampling_ratio, int aligned, float finest_scale, const float *chan_scale,
                                        int chan_scale_rows, const int32_t *scale_index, const float *grad_out,
                                        void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(grad_pyr != nullptr && grad_pyr->num_levels >= 1 && grad_pyr->num_levels <= FGN_MAX_LEVELS, "bad pyramid");
    FGN_CHECK_ARG(R >= 0 && B >= 0 && C > 0 && (C & 3) == 0, "roi_align_bwd needs C%%4==0 (C=%d)", C);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && grad_out, "NULL pointer");
    FGN_CHECK_ARG(chan_scale == nullptr || chan_scale_rows > 0, "chan_scale_rows");
    for (int l = 0; l < grad_pyr->num_levels; ++l) FGN_CHECK_ARG(grad_pyr->feat[l], "gradient level %d is NULL", l);
    if (P != 7 && P != 14) { set_error("roi_align_bwd: P=%d not instantiated (7, 14)", P); return FGN_ERR_UNSUPPORTED; }
    const size_t need = fgn_roi_align_ml_bwd_det_workspace_bytes(grad_pyr, B, C);
    if (!workspace || workspace_bytes < need) {
        set_error("roi_align_bwd_det: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FGN_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
    DetScale *det = (DetScale *)workspace;
    long long *acc0 = (long long *)((char *)workspace + 256);
    absmax_bits_kernel<<<296, 256, 0, st>>>(grad_out, (size_t)R * P * P * C, &det->max_g);
    FGN_LAUNCH_OK();
    if (chan_scale != nullptr) {
        absmax_bits_kernel<<<64, 256, 0, st>>>(chan_scale, (size_t)chan_scale_rows * C, &det->max_cs);
        FGN_LAUNCH_OK();
    }
    fgn_pyramid_t accp = *grad_pyr;                                  // same geometry, feat[l] = the level's int64 accumulators
    size_t off = 0;
    for (int l = 0; l < grad_pyr->num_levels; ++l) { accp.feat[l] = (float *)(acc0 + off); off += det_level_elems(grad_pyr, B, C, l); }
    const Pyramid d = to_device_pyramid(&accp, B);
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int cap = (maxH + maxW + 6 * P + 16 + 3) & ~3;
    const int nblk = (C + 127) / 128;
    if (P == 7) roi_align_bwd_kernel<7, true><<<R * nblk, 7 * 32, (size_t)cap * 4, st>>>(d, C, rois, R, sampling_ratio, aligned,
                                                                                       finest_scale, chan_scale, scale_index, grad_out, cap, det);
    else        roi_align_bwd_kernel<14, true><<<R * nblk, 14 * 32, (size_t)cap * 4, st>>>(d, C, rois, R, sampling_ratio, aligned,
                                                                                         finest_scale, chan_scale, scale_index, grad_out, cap, det);
    FGN_LAUNCH_OK();
    off = 0;
    for (int l = 0; l < grad_pyr->num_levels; ++l) {
        const size_t n = det_level_elems(grad_pyr, B, C, l);
        det_finish_kernel<<<(int)min((size_t)148 * 8, (n + 255) / 256), 256, 0, st>>>(acc0 + off, const_cast<float *>(grad_pyr->feat[l]), n, det);
        FGN_LAUNCH_OK();
        off += n;
    }
    return FGN_OK;
}

extern "C" size_t fgn_channel_attention_bwd_workspace_bytes(int B, int N, int C, int H, int W)
{
    if (B <= 0 || N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * N * ceil_div(H * W, kBwdSlab) * C * sizeof(float);
}

extern "C" int fgn_channel_attention_bwd(const float *qry, const float *vec, const float *grad_out, int B, int N, int C,
                                         int H, int W, float *grad_qry, float *grad_vec, void *workspace,
                                         size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(B >= 0 && N > 0 && C > 0 && (C & 3) == 0 && C <= 1024 && H > 0 && W > 0, "channel_attention_bwd needs C%%4==0, C<=1024");
    if (B == 0) return FGN_OK;
    FGN_CHECK_ARG(grad_out != nullptr, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    if (grad_qry != nullptr) {
        FGN_CHECK_ARG(vec != nullptr, "vec is NULL");
        const size_t total = (size_t)B * HW * (C >> 2);
        channel_attention_bwd_q_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, st>>>(grad_out, vec, B, N, C, (size_t)HW, grad_qry);
        FGN_LAUNCH_OK();
    }
    if (grad_vec != nullptr) {
        FGN_CHECK_ARG(qry != nullptr, "qry is NULL");
        const size_t need = fgn_channel_attention_bwd_workspace_bytes(B, N, C, H, W);
        if (!workspace || workspace_bytes < need) { set_error("channel_attention_bwd: workspace %zu B < %zu B", workspace_bytes, need); return FGN_ERR_WORKSPACE; }
        const int slabs = ceil_div(HW, kBwdSlab), c4 = C >> 2, rows = max(1, 256 / c4);
        FGN_CHECK_ARG(B * N <= 65535, "B*N too large");
        channel_attention_bwd_v_partial_kernel<<<dim3(slabs, B * N), 256, (size_t)rows * C * 4, st>>>(grad_out, qry, N, C, HW, slabs, (float *)workspace);
        FGN_LAUNCH_OK();
        sum_slabs_kernel<<<ceil_div(B * N * C, 256), 256, 0, st>>>((const float *)workspace, B * N, C, slabs, 1.0f, grad_vec);
        FGN_LAUNCH_OK();
    }
    return FGN_OK;
}

extern "C" int fgn_attention_vectors_bwd(const float *grad_vec, int BN, int K, int C, int H, int W, float *grad_spp, void *stream)
{
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0 && (C & 3) == 0 && H > 0 && W > 0, "attention_vectors_bwd needs C%%4==0");
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(grad_vec && grad_spp, "NULL pointer");
    const size_t total = (size_t)BN * K * H * W * (C >> 2);
    attention_vectors_bwd_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_vec, BN, K, C, (size_t)H * W, grad_spp);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_support_pool_bwd(const float *grad_cat, const float *grad_gap, const float *m, int BN, int K, int C, int P,
                                    float *grad_f, void *stream)
{
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0 && (C & 3) == 0 && P > 0, "support_pool_bwd needs C%%4==0");
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(grad_f && (grad_cat || grad_gap) && (grad_gap == nullptr || m != nullptr), "NULL pointer");
    const size_t total = (size_t)BN * K * P * P * (C >> 2);
    support_pool_bwd_kernel<<<(int)min((size_t)148 * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_cat, grad_gap, m, BN, K, C, P * P, grad_f);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

// attention.cu -- AG-RPN channel attention, best-class selection and layout repacks.
// Reference: AGRPNHead.forward_single (fgn_ag_rpn_head.py:44-46 and :87-108).
#include "common.cuh"
#include <algorithm>

namespace fgn {

// out[b*N+n, c, :, :] = qry[b, c, :, :] * vec[b*N+n, c]     (fgn_ag_rpn_head.py:44-46)
// Pure streaming: every query element is read once and written N times.  Stores use the
// streaming (evict-first) hint so the N-fold write stream does not push the query out of L2.
__global__ void __launch_bounds__(256)
channel_attention_nchw_kernel(const float *__restrict__ qry, const float *__restrict__ vec,
                              const int B, const int N, const int C, const int HW,
                              float *__restrict__ out)
{
    // grid.y strides over the planes b*C + c, grid.x over one plane
    for (int plane = blockIdx.y; plane < B * C; plane += gridDim.y) {
    const int b = plane / C, c = plane % C;
    const float *q = qry + (size_t)plane * HW;
    const size_t plane_elems = (size_t)HW;
    if ((HW & 3) == 0) {
        const int n4 = HW >> 2;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            const float4 v = ldg4(q + 4 * (size_t)i);
            for (int n = 0; n < N; ++n) {
                const float s = __ldg(vec + ((size_t)b * N + n) * C + c);
                float4 o = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
                __stcs(reinterpret_cast<float4 *>(out + (((size_t)b * N + n) * C + c) * plane_elems) + i, o);
            }
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            const float v = __ldg(q + i);
            for (int n = 0; n < N; ++n) {
                const float s = __ldg(vec + ((size_t)b * N + n) * C + c);
                __stcs(out + (((size_t)b * N + n) * C + c) * plane_elems + i, v * s);
            }
        }
    }
    }
}

// Small planes (RoI features, HW = 49 / 196): flat index over [B*C*HW], one element per thread step.
__global__ void __launch_bounds__(256)
channel_attention_nchw_small_kernel(const float *__restrict__ qry, const float *__restrict__ vec,
                                    const int B, const int N, const int C, const int HW,
                                    float *__restrict__ out)
{
    const size_t total = (size_t)B * C * HW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t plane = i / HW;
        const int p = (int)(i - plane * HW), b = (int)(plane / C), c = (int)(plane - (size_t)b * C);
        const float v = __ldg(qry + i);
        for (int n = 0; n < N; ++n) {
            const float s = __ldg(vec + ((size_t)b * N + n) * C + c);
            __stcs(out + (((size_t)b * N + n) * C + c) * HW + p, v * s);
        }
    }
}

__global__ void __launch_bounds__(256)
channel_attention_nhwc_kernel(const float *__restrict__ qry, const float *__restrict__ vec,
                              const int B, const int N, const int C, const size_t HW,
                              float *__restrict__ out)
{
    // element index over [B, HW, C/4]
    const int c4 = C >> 2;
    const size_t per_img = HW * c4;
    const size_t total = (size_t)B * per_img;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int b = i / per_img;
        const size_t rem = i % per_img;
        const int cv = rem % c4;
        const float4 v = ldg4(qry + 4 * i);
        for (int n = 0; n < N; ++n) {
            const float4 s = ldg4(vec + ((size_t)b * N + n) * C + cv * 4);
            float4 o = make_float4(v.x * s.x, v.y * s.y, v.z * s.z, v.w * s.w);
            __stcs(reinterpret_cast<float4 *>(out + ((size_t)b * N + n) * HW * C) + rem, o);
        }
    }
}

// Per anchor position: argmax over the N classes of the objectness score (first maximum wins,
// like torch.argmax), then gather that class's score and 4 deltas (fgn_ag_rpn_head.py:87-108).
// cls [B*N,A,H,W]; reg [B*N,4A,H,W] with delta d of anchor a at channel 4a+d.
__global__ void best_class_select_kernel(const float *__restrict__ cls, const float *__restrict__ reg,
                                         const int B, const int N, const int A, const int HW,
                                         float *__restrict__ cls_out, float *__restrict__ reg_out)
{
    const size_t total = (size_t)B * A * HW;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int p = i % HW, a = (i / HW) % A, b = i / ((size_t)HW * A);
        int best = 0;
        float bv = __ldg(cls + (((size_t)b * N) * A + a) * HW + p);
        for (int n = 1; n < N; ++n) {
            const float v = __ldg(cls + (((size_t)b * N + n) * A + a) * HW + p);
            // torch.argmax: NaN counts as maximal, first occurrence wins
            if ((v > bv) || (v != v && bv == bv)) { bv = v; best = n; }
        }
        cls_out[i] = bv;
#pragma unroll
        for (int d = 0; d < 4; ++d)
            reg_out[(((size_t)b * 4 * A) + 4 * a + d) * HW + p] =
                __ldg(reg + (((size_t)b * N + best) * 4 * A + 4 * a + d) * HW + p);
    }
}

// ---- layout repacks -----------------------------------------------------------------------------
// [B,C,HW] <-> [B,HW,C] through a 32x33 shared tile; both sides coalesced.
__global__ void __launch_bounds__(256)
transpose_kernel(const float *__restrict__ in, const int rows, const int cols, float *__restrict__ out)
{
    // per batch item: in [rows, cols] -> out [cols, rows]
    __shared__ float tile[32][33];
    const size_t boff = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) tile[j][tx] = __ldg(in + boff + (size_t)r * cols + c);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[tx][j];
    }
}

}  // namespace fgn

using namespace fgn;

extern "C" int fgn_channel_attention(const float *qry, const float *vec, int B, int N, int C, int H,
                                     int W, int layout, float *out, void *stream)
{
    FGN_CHECK_ARG(B >= 0 && N > 0 && C > 0 && H > 0 && W > 0, "bad dims");
    if (B == 0) return FGN_OK;
    FGN_CHECK_ARG(qry && vec && out, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    if (layout == FGN_LAYOUT_NCHW) {
        if (HW < 1024) {
            const size_t total = (size_t)B * C * HW;
            const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
            channel_attention_nchw_small_kernel<<<blocks, 256, 0, st>>>(qry, vec, B, N, C, HW, out);
        } else {
            const int per = (HW & 3) == 0 ? HW >> 2 : HW;
            dim3 grid(max(1, min(ceil_div(per, 256), 64)), (unsigned)min((long)B * C, 65535L));
            channel_attention_nchw_kernel<<<grid, 256, 0, st>>>(qry, vec, B, N, C, HW, out);
        }
    } else if (layout == FGN_LAYOUT_NHWC) {
        if (C & 3) { set_error("channel_attention NHWC needs C%%4==0 (C=%d)", C); return FGN_ERR_UNSUPPORTED; }
        const size_t total = (size_t)B * HW * (C >> 2);
        const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
        channel_attention_nhwc_kernel<<<blocks, 256, 0, st>>>(qry, vec, B, N, C, (size_t)HW, out);
    } else { set_error("layout=%d", layout); return FGN_ERR_INVALID_ARG; }
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_best_class_select(const float *cls, const float *reg, int B, int N, int A, int H,
                                     int W, float *cls_out, float *reg_out, void *stream)
{
    FGN_CHECK_ARG(B >= 0 && N > 0 && A > 0 && H > 0 && W > 0, "bad dims");
    if (B == 0) return FGN_OK;
    FGN_CHECK_ARG(cls && reg && cls_out && reg_out, "NULL pointer");
    const size_t total = (size_t)B * A * H * W;
    const int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    best_class_select_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cls, reg, B, N, A, H * W,
                                                                       cls_out, reg_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

static int launch_transpose(const float *in, int batch, int rows, int cols, float *out, cudaStream_t st)
{
    FGN_CHECK_ARG(batch >= 0 && rows > 0 && cols > 0, "bad dims");
    if (batch == 0) return FGN_OK;
    FGN_CHECK_ARG(in && out, "NULL pointer");
    FGN_CHECK_ARG(batch <= 65535 && ceil_div(rows, 32) <= 65535, "transpose grid too large");
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
    transpose_kernel<<<grid, 256, 0, st>>>(in, rows, cols, out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_nchw_to_nhwc(const float *in, int B, int C, int H, int W, float *out, void *stream)
{
    return launch_transpose(in, B, C, H * W, out, (cudaStream_t)stream);   // [C,HW] -> [HW,C]
}

extern "C" int fgn_nhwc_to_nchw(const float *in, int B, int C, int H, int W, float *out, void *stream)
{
    return launch_transpose(in, B, H * W, C, out, (cudaStream_t)stream);   // [HW,C] -> [C,HW]
}

// ---- channel attention folded into the consumer's weights ----------------------------------------------------
// conv(qry * vec[bn]) == conv'(qry) with w'[bn,o,c,k] = w[o,c,k] * vec[bn,c]: the [B*N,C,H,W] product of
// fgn_ag_rpn_head.py:44-46 is never written or re-read; the RPN conv (mmdet RPNHead [3P], cuDNN) runs on the
// unmodified query map with B*N weight sets (one grouped conv).
namespace fgn {
namespace {
__global__ void __launch_bounds__(256)
fold_attention_weights_kernel(const float *__restrict__ w, const float *__restrict__ vec, const int Co, const int Ci,
                              const int KK, const size_t total, float *__restrict__ out)
{
    const size_t per = (size_t)Co * Ci * KK;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bn = i / per, r = i - bn * per;
        const int c = (int)((r / KK) % Ci);
        out[i] = __fmul_rn(w[r], vec[bn * Ci + c]);
    }
}
}  // namespace
}  // namespace fgn

extern "C" int fgn_fold_attention_weights(const float *weight, const float *vec, int BN, int Co, int Ci, int KK,
                                          float *out, void *stream)
{
    FGN_CHECK_ARG(BN >= 0 && Co >= 1 && Ci >= 1 && KK >= 1, "bad dims BN=%d Co=%d Ci=%d KK=%d", BN, Co, Ci, KK);
    if (BN == 0) return FGN_OK;
    FGN_CHECK_ARG(weight && vec && out, "NULL pointer");
    const size_t total = (size_t)BN * Co * Ci * KK;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
    fgn::fold_attention_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, vec, Co, Ci, KK, total, out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

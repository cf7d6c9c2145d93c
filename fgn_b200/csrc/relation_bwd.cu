// relation_bwd.cu -- adjoint of the Relation-Guided Detector fusion (fgn_relation_fusion_fwd).
//
// Reference: the training path of FGNRoIHead (fgn_roi_head.py:344-358,451-529) obtains these gradients from autograd
// over count_one_roi_by_n_spp (:253-279), BBoxHead.forward [3P] (:338) and count_modified_cls_bbox (:302-326).
// Forward, per (RoI r, class n), NHWC rows [49, C]:
//     y = yq[r] + ys[b(r), n]                     (yq = Xq Wq^T, ys = Xs Ws^T + bias: the split 1x1 conv)
//     g = (y - mean_G) * rstd_G                   (GroupNorm(32): statistics over cg channels x 49 positions)
//     a = relu(g * gamma + beta) ; z = mean_p a ; raw = Wfc z + bfc   (2 cls + 4 reg rows)
// Backward:
//     dz = Wfc^T draw ; da = dz / 49 * [a > 0] ; dgamma += sum_p da g ; dbeta += sum_p da ; dg = da gamma
//     dy = rstd (dg - mean_G(dg) - g mean_G(dg g))                       (GroupNorm adjoint)
//     dyq[r] = sum_n dy[r,n] ; dys[b,n] = sum_{r in b} dy[r,n] ; dWfc += draw z^T
// Three kernels: relation_bwd_roi_kernel (thread = channel, one CTA per RoI and channel block: dyq, the per-(r,n)
// group terms, per-RoI partials of the parameter gradients), relation_bwd_class_kernel (dys: re-derives dy pointwise from
// the stored group terms), column_sum_kernel (partials -> parameter gradients).  The four contractions of the conv
// adjoint (dXq = dyq Wq, dXs = dys Ws, dWq = dyq^T Xq, dWs = dys^T Xs) run on the forward's contraction kernels.
#include "common.cuh"
#include "gemm.cuh"

namespace fgn {

namespace {

constexpr int kBwdThreads = 256;
constexpr int kPP = 49;

__device__ __forceinline__ float group_sum(float v, int cg)
{
    for (int o = cg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// count_modified_cls_bbox adjoint: (dcls [R,N+1], dreg [R,4N], fg scores of the forward) -> draw [R*N, 6] (bg, fg, 4 deltas)
__global__ void reassemble_bwd_kernel(const float *__restrict__ dcls, const float *__restrict__ dreg,
                                      const float *__restrict__ cls_fwd, const int R, const int N,
                                      float *__restrict__ draw)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int best = 0;
    float bv = cls_fwd[(size_t)r * (N + 1)];
    for (int n = 1; n < N; ++n) {                      // first maximum, NaN counts as maximal (torch.argmax)
        const float v = cls_fwd[(size_t)r * (N + 1) + n];
        if (v > bv || (v != v && bv == bv)) { bv = v; best = n; }
    }
    for (int n = 0; n < N; ++n) {
        float *o = draw + ((size_t)r * N + n) * 6;
        o[0] = n == best ? dcls[(size_t)r * (N + 1) + N] : 0.f;      // bg logit feeds the last column through the argmax class
        o[1] = dcls[(size_t)r * (N + 1) + n];
#pragma unroll
        for (int d = 0; d < 4; ++d) o[2 + d] = dreg[(size_t)r * 4 * N + 4 * n + d];
    }
}

// grid (R, nblk), thread = channel.  Outputs: dyq [R,49,C]; gterm [R,N,32,4] = (mean, rstd, mean_G(dg), mean_G(dg g));
// part [R, 8, C] = per-RoI partial sums of (dgamma, dbeta, dWfc rows 0..5).
__global__ void __launch_bounds__(kBwdThreads)
relation_bwd_roi_kernel(const float *__restrict__ Yq, const float *__restrict__ Ys, const int32_t *__restrict__ roi_batch,
                        const float *__restrict__ draw, const int R, const int B, const int N, const int C, const int cg,
                        const float eps, const float *__restrict__ gn_w, const float *__restrict__ gn_b,
                        const float *__restrict__ fc_cls_w, const float *__restrict__ fc_reg_w,
                        float *__restrict__ dYq, float *__restrict__ gterm, float *__restrict__ part)
{
    const int r = blockIdx.x, c = blockIdx.y * kBwdThreads + threadIdx.x;
    const bool active = c < C;
    int b = roi_batch[r];
    b = b < 0 ? 0 : (b >= B ? B - 1 : b);
    const int groups = C / cg;
    float yq[kPP], dyq[kPP];
#pragma unroll
    for (int p = 0; p < kPP; ++p) { yq[p] = active ? __ldg(Yq + ((size_t)r * kPP + p) * C + c) : 0.f; dyq[p] = 0.f; }
    const float gamma = active ? __ldg(gn_w + c) : 0.f, beta = active ? __ldg(gn_b + c) : 0.f;
    float wfc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j)
        wfc[j] = active ? (j < 2 ? __ldg(fc_cls_w + (size_t)j * C + c) : __ldg(fc_reg_w + (size_t)(j - 2) * C + c)) : 0.f;
    const float inv_m = 1.0f / (float)(cg * kPP);
    float dgamma = 0.f, dbeta = 0.f, dw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    for (int n = 0; n < N; ++n) {
        const float *ys = Ys + ((size_t)(b * N + n) * kPP) * C + c;
        const float *dr = draw + ((size_t)r * N + n) * 6;
        float y[kPP];
        float s1 = 0.f;
#pragma unroll
        for (int p = 0; p < kPP; ++p) { y[p] = active ? yq[p] + __ldg(ys + (size_t)p * C) : 0.f; s1 += y[p]; }
        const float mean = group_sum(s1, cg) * inv_m;
        float s2 = 0.f;
#pragma unroll
        for (int p = 0; p < kPP; ++p) { const float d = y[p] - mean; s2 = fmaf(d, d, s2); }
        if (!active) s2 = 0.f;
        const float rstd = 1.0f / sqrtf(group_sum(s2, cg) * inv_m + eps);
        float dz = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) dz = fmaf(wfc[j], dr[j], dz);
        const float da0 = dz * (1.0f / kPP);
        // first sweep: z (for dWfc), dgamma / dbeta, the two group terms
        float z = 0.f, t1 = 0.f, t2 = 0.f, dgam = 0.f, dbet = 0.f;
#pragma unroll
        for (int p = 0; p < kPP; ++p) {
            const float g = (y[p] - mean) * rstd;
            const float a = fmaf(g, gamma, beta);
            const float da = a > 0.f ? da0 : 0.f;
            z += fmaxf(a, 0.f);
            dgam = fmaf(da, g, dgam);
            dbet += da;
            const float dg = da * gamma;
            t1 += dg;
            t2 = fmaf(dg, g, t2);
        }
        if (!active) { z = 0.f; t1 = 0.f; t2 = 0.f; dgam = 0.f; dbet = 0.f; }
        z *= (1.0f / kPP);
        dgamma += dgam; dbeta += dbet;
#pragma unroll
        for (int j = 0; j < 6; ++j) dw[j] = fmaf(dr[j], z, dw[j]);
        const float m1 = group_sum(t1, cg) * inv_m, m2 = group_sum(t2, cg) * inv_m;
        if (active && (c % cg) == 0) {
            float *gt = gterm + (((size_t)r * N + n) * groups + c / cg) * 4;
            gt[0] = mean; gt[1] = rstd; gt[2] = m1; gt[3] = m2;
        }
        // second sweep: dy, accumulated over the classes
#pragma unroll
        for (int p = 0; p < kPP; ++p) {
            const float g = (y[p] - mean) * rstd;
            const float a = fmaf(g, gamma, beta);
            const float dg = a > 0.f ? da0 * gamma : 0.f;
            dyq[p] += rstd * (dg - m1 - g * m2);
        }
    }
    if (!active) return;
#pragma unroll
    for (int p = 0; p < kPP; ++p) dYq[((size_t)r * kPP + p) * C + c] = dyq[p];
    float *pr = part + (size_t)r * 8 * C + c;
    pr[0] = dgamma; pr[C] = dbeta;
#pragma unroll
    for (int j = 0; j < 6; ++j) pr[(size_t)(2 + j) * C] = dw[j];
}

// dYs[b,n][p][c] = sum over the RoIs of image b of dy[r,n][p][c], re-derived pointwise from the stored group terms.
// grid (B*N, 49), thread = channel (loops over channel blocks).
__global__ void __launch_bounds__(kBwdThreads)
relation_bwd_class_kernel(const float *__restrict__ Yq, const float *__restrict__ Ys, const int32_t *__restrict__ roi_batch,
                          const float *__restrict__ draw, const float *__restrict__ gterm, const int R, const int B,
                          const int N, const int C, const int cg, const float *__restrict__ gn_w,
                          const float *__restrict__ gn_b, const float *__restrict__ fc_cls_w,
                          const float *__restrict__ fc_reg_w, float *__restrict__ dYs)
{
    const int bn = blockIdx.x, p = blockIdx.y, b = bn / N, n = bn - b * N;
    const int groups = C / cg;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float gamma = __ldg(gn_w + c), beta = __ldg(gn_b + c);
        float wfc[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) wfc[j] = j < 2 ? __ldg(fc_cls_w + (size_t)j * C + c) : __ldg(fc_reg_w + (size_t)(j - 2) * C + c);
        const float ys = __ldg(Ys + ((size_t)bn * kPP + p) * C + c);
        float acc = 0.f;
        for (int r = 0; r < R; ++r) {
            int rb = roi_batch[r];
            rb = rb < 0 ? 0 : (rb >= B ? B - 1 : rb);
            if (rb != b) continue;                                   // (warp-uniform)
            const float4 gt = *reinterpret_cast<const float4 *>(gterm + (((size_t)r * N + n) * groups + c / cg) * 4);
            const float *dr = draw + ((size_t)r * N + n) * 6;
            float dz = 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j) dz = fmaf(wfc[j], dr[j], dz);
            const float y = __ldg(Yq + ((size_t)r * kPP + p) * C + c) + ys;
            const float g = (y - gt.x) * gt.y;
            const float a = fmaf(g, gamma, beta);
            const float dg = a > 0.f ? dz * (1.0f / kPP) * gamma : 0.f;
            acc += gt.y * (dg - gt.z - g * gt.w);
        }
        dYs[((size_t)bn * kPP + p) * C + c] = acc;
    }
}

// out[j][c] = sum_i in[i][j][c]   (in [rows, J, C]) -- fixed order, deterministic
__global__ void column_sum_kernel(const float *__restrict__ in, const int rows, const int J, const int C,
                                  float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J * C) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += in[(size_t)r * J * C + i];
    out[i] = s;
}

// [rows, cols] -> [cols, rows_pad] (rows_pad >= rows, the tail zero-filled): K-major operands for the weight-gradient
// contractions, K = rows padded to the contraction kernel's k-block
__global__ void transpose_pad_kernel(const float *__restrict__ in, const int rows, const int cols, const int rows_pad,
                                     float *__restrict__ out)
{
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < rows && c < cols) ? __ldg(in + (size_t)r * cols + c) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (c < cols && r < rows_pad) out[(size_t)c * rows_pad + r] = tile[tx][j];
    }
}

size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BwdWs {
    float *draw, *dyq, *dys, *gterm, *part, *t_dy, *t_x, *t_w, *split;
    size_t bytes;
};

BwdWs carve_bwd(void *base, int R, int BN, int N, int C)
{
    BwdWs w;
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { char *p = b ? b + off : nullptr; off += a256(bytes); return (float *)p; };
    const size_t Mq = (size_t)R * kPP, Ms = (size_t)BN * kPP, Mpad = ((Mq > Ms ? Mq : Ms) + 31) & ~(size_t)31;
    w.draw  = take((size_t)R * N * 6 * 4);
    w.dyq   = take(Mq * C * 4);
    w.dys   = take(Ms * C * 4);
    w.gterm = take((size_t)R * N * 32 * 4 * 4);
    w.part  = take((size_t)R * 8 * C * 4);
    w.t_dy  = take((size_t)C * Mpad * 4);
    w.t_x   = take((size_t)C * Mpad * 4);
    w.t_w   = take((size_t)C * C * 4);
    w.split = take(gemm_tc_workspace_bytes(C, (int)Mpad));      // TF32 split of the [C, Mpad] B operand
    w.bytes = off;
    return w;
}

}  // namespace

}  // namespace fgn

using namespace fgn;

extern "C" size_t fgn_relation_fusion_bwd_workspace_bytes(int R, int BN, int N, int C)
{
    if (R <= 0 || BN <= 0 || N <= 0 || C <= 0) return 0;
    return carve_bwd(nullptr, R, BN, N, C).bytes;
}

// Gradients of fgn_relation_fusion_fwd.  All feature tensors NHWC.  Yq [R*49, C] and Ys [B*N*49, C] are the forward's
// split-conv outputs (the forward's workspace; re-derive them with fgn_gemm_nt if they were not kept);
// cls_fwd [R,N+1] = the forward's cls_out (its argmax decides where the background gradient goes).
// Outputs (each may be NULL): d_roi_feat [R,49,C], d_spp [B*N,49,C], d_conv_w [C,2C], d_conv_b [C], d_gn_w, d_gn_b [C],
// d_fc_cls_w [2,C], d_fc_cls_b [2], d_fc_reg_w [4,C], d_fc_reg_b [4].
extern "C" int fgn_relation_fusion_bwd(const float *roi_feat, const float *spp_cat_mean, const int32_t *roi_batch,
                                       const float *Yq, const float *Ys, const float *cls_fwd,
                                       const float *d_cls, const float *d_reg, int R, int B, int N, int C, int P,
                                       const float *conv_w, const float *gn_w, const float *gn_b, int gn_groups,
                                       float gn_eps, const float *fc_cls_w, const float *fc_reg_w,
                                       float *d_roi_feat, float *d_spp, float *d_conv_w, float *d_conv_b,
                                       float *d_gn_w, float *d_gn_b, float *d_fc_cls_w, float *d_fc_cls_b,
                                       float *d_fc_reg_w, float *d_fc_reg_b,
                                       void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(R >= 0 && B > 0 && N > 0 && C > 0, "bad dims R=%d B=%d N=%d C=%d", R, B, N, C);
    if (R == 0) return FGN_OK;
    if (P * P != kPP) { set_error("relation_fusion_bwd: P=%d not instantiated (7)", P); return FGN_ERR_UNSUPPORTED; }
    FGN_CHECK_ARG(gn_groups > 0 && C % gn_groups == 0, "GroupNorm groups=%d does not divide C=%d", gn_groups, C);
    const int cg = C / gn_groups;
    if ((cg & (cg - 1)) != 0 || cg > 32 || gn_groups > 32 || (C & 3)) {
        set_error("relation_fusion_bwd: needs <= 32 groups of a power-of-two <= 32 channels and C %% 4 == 0 (C=%d, groups=%d)", C, gn_groups);
        return FGN_ERR_UNSUPPORTED;
    }
    FGN_CHECK_ARG(roi_feat && spp_cat_mean && roi_batch && Yq && Ys && cls_fwd && d_cls && d_reg && conv_w && gn_w && gn_b &&
                  fc_cls_w && fc_reg_w, "NULL pointer");
    const int BN = B * N;
    const BwdWs w = carve_bwd(workspace, R, BN, N, C);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("relation_fusion_bwd: workspace %zu B < required %zu B", workspace_bytes, w.bytes);
        return FGN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int Mq = R * kPP, Ms = BN * kPP;

    reassemble_bwd_kernel<<<ceil_div(R, 128), 128, 0, st>>>(d_cls, d_reg, cls_fwd, R, N, w.draw);
    FGN_LAUNCH_OK();
    relation_bwd_roi_kernel<<<dim3(R, ceil_div(C, kBwdThreads)), kBwdThreads, 0, st>>>(
        Yq, Ys, roi_batch, w.draw, R, B, N, C, cg, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, w.dyq, w.gterm, w.part);
    FGN_LAUNCH_OK();
    relation_bwd_class_kernel<<<dim3(BN, kPP), kBwdThreads, 0, st>>>(Yq, Ys, roi_batch, w.draw, w.gterm, R, B, N, C, cg,
                                                                    gn_w, gn_b, fc_cls_w, fc_reg_w, w.dys);
    FGN_LAUNCH_OK();
    // parameter gradients of GroupNorm and the FC heads: column sums of the per-RoI partials
    if (d_gn_w || d_gn_b || d_fc_cls_w || d_fc_reg_w) {
        float *sums = w.t_w;                                     // [8, C] (t_w is free until the conv adjoint below)
        column_sum_kernel<<<ceil_div(8 * C, 256), 256, 0, st>>>(w.part, R, 8, C, sums);
        FGN_LAUNCH_OK();
        if (d_gn_w) FGN_CUDA_OK(cudaMemcpyAsync(d_gn_w, sums, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
        if (d_gn_b) FGN_CUDA_OK(cudaMemcpyAsync(d_gn_b, sums + C, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
        if (d_fc_cls_w) FGN_CUDA_OK(cudaMemcpyAsync(d_fc_cls_w, sums + 2 * (size_t)C, (size_t)2 * C * 4, cudaMemcpyDeviceToDevice, st));
        if (d_fc_reg_w) FGN_CUDA_OK(cudaMemcpyAsync(d_fc_reg_w, sums + 4 * (size_t)C, (size_t)4 * C * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (d_fc_cls_b || d_fc_reg_b) {                              // column sums of draw [R*N, 6]
        float *sums = w.gterm;                                   // (its first 6 floats; gterm is no longer needed)
        column_sum_kernel<<<1, 32, 0, st>>>(w.draw, R * N, 1, 6, sums);
        FGN_LAUNCH_OK();
        if (d_fc_cls_b) FGN_CUDA_OK(cudaMemcpyAsync(d_fc_cls_b, sums, 2 * 4, cudaMemcpyDeviceToDevice, st));
        if (d_fc_reg_b) FGN_CUDA_OK(cudaMemcpyAsync(d_fc_reg_b, sums + 2, 4 * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (d_conv_b) {                                              // bias rides in ys: d_bias = column sum of dys
        column_sum_kernel<<<ceil_div(C, 256), 256, 0, st>>>(w.dys, Ms, 1, C, d_conv_b);
        FGN_LAUNCH_OK();
    }
    // conv adjoint, data side: dXq = dyq Wq, dXs = dys Ws   (gemm_nt wants B as [N_out = C_in, K = C_out]: W^T)
    const dim3 tgrid_w(ceil_div(C, 32), ceil_div(C, 32));
    int rc;
    // (transpose_pad reads a densely packed [rows, cols] matrix; the halves of conv_w have row pitch 2C: copy them out first)
    for (int half = 0; half < 2; ++half) {
        float *dX = half == 0 ? d_roi_feat : d_spp;
        if (!dX) continue;
        const float *dY = half == 0 ? w.dyq : w.dys;
        const int M = half == 0 ? Mq : Ms;
        // Wh^T: [C_in, C_out] from conv_w[:, half*C : (half+1)*C] (row pitch 2C)
        FGN_CUDA_OK(cudaMemcpy2DAsync(w.t_x, (size_t)C * 4, conv_w + (size_t)half * C, (size_t)2 * C * 4, (size_t)C * 4, C,
                                      cudaMemcpyDeviceToDevice, st));
        transpose_pad_kernel<<<tgrid_w, 256, 0, st>>>(w.t_x, C, C, C, w.t_w);
        FGN_LAUNCH_OK();
        rc = gemm_nt(dY, C, w.t_w, C, nullptr, dX, C, M, C, C, 0, w.split, st);
        if (rc) return rc;
    }
    // conv adjoint, weight side: dWq = dyq^T Xq, dWs = dys^T Xs  ->  d_conv_w[:, half*C : (half+1)*C]
    if (d_conv_w) {
        for (int half = 0; half < 2; ++half) {
            const float *dY = half == 0 ? w.dyq : w.dys, *X = half == 0 ? roi_feat : spp_cat_mean;
            const int M = half == 0 ? Mq : Ms, Mpad = (M + 31) & ~31;
            const dim3 tg(ceil_div(C, 32), ceil_div(Mpad, 32));
            transpose_pad_kernel<<<tg, 256, 0, st>>>(dY, M, C, Mpad, w.t_dy);
            FGN_LAUNCH_OK();
            transpose_pad_kernel<<<tg, 256, 0, st>>>(X, M, C, Mpad, w.t_x);
            FGN_LAUNCH_OK();
            rc = gemm_nt(w.t_dy, Mpad, w.t_x, Mpad, nullptr, d_conv_w + (size_t)half * C, 2 * C, C, C, Mpad, 0, w.split, st);
            if (rc) return rc;
        }
    }
    return FGN_OK;
}

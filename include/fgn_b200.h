/*
 * fgn_b200.h -- C ABI of libfgn_b200.so: the B200 (sm_100a) implementation of FGN's guided
 * RoIAlign + support-guided fusion hot path.
 *
 * The reference (tooHotSpot/FGN) has no native code and no FFI; the path sits behind Python
 * modules registered with mmdet.  Each entry point below names the
 * reference call site (file:line under /root/reference/subprojects/sp02_omniiseg_fgn_mmdet/)
 * whose device work it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - all tensors are fp32 and densely packed in the stated layout; int tensors are int32;
 *   - outputs are caller-allocated; no entry point allocates, frees or keeps device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host;
 *   - return value: 0 on success, <0 = FGN_ERR_* ; fgn_last_error_string() (thread-local)
 *     describes the last failure on the calling thread;
 *   - zero-size problems (R == 0 ...) return 0 without launching anything
 *     (reference early-outs: fgn_roi_head.py:558-567,596-603,638-640).
 */
#ifndef FGN_B200_H_
#define FGN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGN_ABI_VERSION 3

#define FGN_OK                 0
#define FGN_ERR_INVALID_ARG   -1
#define FGN_ERR_UNSUPPORTED   -2
#define FGN_ERR_CUDA          -3
#define FGN_ERR_WORKSPACE     -4

/* feature-map / RoI-feature memory layouts */
#define FGN_LAYOUT_NCHW 0      /* [B,C,H,W]  the reference's layout                      */
#define FGN_LAYOUT_NHWC 1      /* [B,H,W,C]  torch.channels_last storage; the fast layout */

#define FGN_MAX_LEVELS 8

int         fgn_abi_version(void);
const char *fgn_last_error_string(void);
/* number of SMs / compute capability of the current device (host query, for tests/bench) */
int         fgn_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* mmdet SingleRoIExtractor.map_roi_levels [3P] (config fgn_r50_c4_densecl.py:69-73, reached
 * from fgn_roi_head.py:331-332,366-367):
 *   lvl = clamp(floor(log2(sqrt((x2-x1)(y2-y1)) / finest_scale + 1e-6)), 0, L-1)
 * rois [R,5] (batch_idx,x1,y1,x2,y2) -> lvl_out [R] int32. */
int fgn_map_roi_levels(const float *rois, int R, int num_levels, float finest_scale,
                       int32_t *lvl_out, void *stream);

/* Pyramid descriptor passed by value (host struct holding device pointers). */
typedef struct {
    int          num_levels;                 /* 1 = the reference's C4 single-level mode */
    const float *feat[FGN_MAX_LEVELS];       /* level l: [B,C,H_l,W_l] in `layout`       */
    int          H[FGN_MAX_LEVELS];
    int          W[FGN_MAX_LEVELS];
    float        spatial_scale[FGN_MAX_LEVELS];   /* 1/stride_l                          */
} fgn_pyramid_t;

/* SingleRoIExtractor.forward + mmcv.ops.RoIAlign (avg) [3P], fgn_roi_head.py:331-332,366-367,
 * and torchvision.ops.roi_align, fgn_roi_head.py:429,432.  Level assignment, the adaptive
 * sampling grid and bilinear pooling run in one kernel; with num_levels == 1 the level step is
 * skipped exactly like mmdet's short-circuit.
 *   rois [R,5]; out [R,C,P,P] in out_layout (NCHW = [R,C,P,P], NHWC = [R,P,P,C]).
 *   sampling_ratio <= 0 -> adaptive grid ceil(roi/P); aligned: 1 = mmcv default, 0 = torchvision default.
 *   chan_scale (optional, may be NULL): [S,C] vectors; out[r,c,:,:] *= chan_scale[scale_index[r],c]
 *     (the AG-FCN attention multiply fgn_roi_head.py:379 with the gather :707-714 fused in);
 *   lvl_out (optional): [R] int32 level each RoI was pooled from.
 * in_layout NHWC is the fast path (128-bit channel-vector loads); NCHW input is served by a
 * direct kernel that keeps the reference layout. */
int fgn_roi_align_ml_fwd(const fgn_pyramid_t *pyr, int B, int C, int in_layout,
                         const float *rois, int R, int P, int sampling_ratio, int aligned,
                         float finest_scale,
                         const float *chan_scale, const int32_t *scale_index,
                         float *out, int out_layout, int32_t *lvl_out, void *stream);

/* Integer side of the same kernel for the bit-exact check (test/debug export).  For RoI r,
 * pooled from the level fgn_roi_align_ml_fwd would choose:
 *   grid_out [R,2]              = (grid_h, grid_w)
 *   ytab_out [R,P,max_grid,3]   = (valid, y_low, y_high) per (ph, iy), -1 padded past grid_h
 *   xtab_out [R,P,max_grid,3]   = (valid, x_low, x_high) per (pw, ix)
 * The (r,ph,pw,iy,ix) sample tuple of the reference is the cartesian product of the two. */
int fgn_roi_align_sample_indices(const fgn_pyramid_t *pyr, const float *rois, int R, int P,
                                 int sampling_ratio, int aligned, float finest_scale,
                                 int max_grid, int32_t *lvl_out, int32_t *grid_out,
                                 int32_t *ytab_out, int32_t *xtab_out, void *stream);

/* NCHW -> NHWC repack of a feature map (one read, one write); used by the host shim when a
 * caller hands over reference-layout tensors and wants the fast path. */
int fgn_nchw_to_nhwc(const float *in, int B, int C, int H, int W, float *out, void *stream);
int fgn_nhwc_to_nchw(const float *in, int B, int C, int H, int W, float *out, void *stream);

/* count_spp, mask part (fgn_roi_head.py:429): roi_align(spp_isegmaps.float(), boxes, 7) with
 * spatial_scale=1, sampling_ratio=-1, aligned=False, one box per support image.
 *   mask [M,S_h,S_w] uint8 (torch.bool storage); boxes [M,4] XYXY px -> out [M,P,P]. */
int fgn_support_mask_pool(const uint8_t *mask, const float *boxes, int M, int S_h, int S_w,
                          int P, float *out, void *stream);

/* count_spp, reduction part (fgn_roi_head.py:439-447):
 *   cat_mean[b,n,c,p] = mean_k f[(b*N+n)*K+k, c, p]
 *   masked_gap[b,n,c] = mean_{k,p} f[..]*m[(b*N+n)*K+k, p]          (divides by K*P*P)
 *   f [B*N*K,C,P,P] in f_layout; m [B*N*K,P,P]; cat_mean [B*N,C,P,P] in out_layout;
 *   masked_gap [B*N,C]. */
int fgn_support_pool(const float *f, int f_layout, const float *m, int BN, int K, int C, int P,
                     float *cat_mean, int out_layout, float *masked_gap, void *stream);

/* The whole of count_spp (fgn_roi_head.py:419-449) in ONE launch, for heads without a shared_head between RoIAlign and
 * the class mean (FPN mode): mask pooling (:429), support RoIAlign with level assignment (:432; aligned=False,
 * sampling_ratio=-1), class mean (:439-442), masked GAP (:444-447) -- and, when conv_w is given, the class half of the
 * relation convolution (:272) Ys = cat_mean Ws^T + conv_b, which fgn_relation_fusion_fwd / fgn_guided_roi_fused_fwd accept
 * as class_term.  One CTA per (class, bin); the masked GAP's cross-bin sum is finished by the last CTA of a class.
 *   spp: NHWC support maps per level, [M = BN*K, H_l, W_l, C] with spatial_scale[l] (C4: one level, scale 1/16 -- the same
 *        cells as the reference's boxes /= 16 with scale 1, a power of two); boxes [M,4] XYXY px; masks [M,S_h,S_w] uint8;
 *   cat_mean [BN,P,P,C] (NHWC), masked_gap [BN,C], class_term [BN*P*P,C] (optional, needs conv_w [C,2C] and conv_b [C]).
 * workspace: fgn_support_prologue_workspace_bytes(BN, C, P) bytes. */
size_t fgn_support_prologue_workspace_bytes(int BN, int C, int P);
int fgn_support_prologue_fwd(const fgn_pyramid_t *spp, int C, const float *boxes, const uint8_t *masks,
                             int S_h, int S_w, int BN, int K, int P, float finest_scale,
                             const float *conv_w, const float *conv_b, float *cat_mean, float *masked_gap,
                             float *class_term, void *workspace, size_t workspace_bytes, void *stream);

/* AGRPNHead class attention vector (fgn_ag_rpn_head.py:37-41):
 *   vec[b,n,c] = mean_{k,h,w} spp_fmaps[(b*N+n)*K+k, c, h, w]      -> vec [B*N,C]
 * workspace: fgn_attention_vectors_workspace_bytes(...) bytes (may be 0). */
size_t fgn_attention_vectors_workspace_bytes(int BN, int K, int C, int H, int W, int layout);
int fgn_attention_vectors(const float *spp_fmaps, int layout, int BN, int K, int C, int H, int W,
                          float *vec, void *workspace, size_t workspace_bytes, void *stream);

/* AGRPNHead channel attention (fgn_ag_rpn_head.py:44-46):
 *   out[b*N+n, c, h, w] = qry[b,c,h,w] * vec[b*N+n, c]
 *   qry [B,C,H,W], out [B*N,C,H,W], both in `layout`. */
int fgn_channel_attention(const float *qry, const float *vec, int B, int N, int C, int H, int W,
                          int layout, float *out, void *stream);

/* The same attention without its [B*N,C,H,W] output: conv(qry * vec[bn]) == conv'(qry) with
 *   out[bn, o, c, k] = weight[o, c, k] * vec[bn, c]      (weight [Co,Ci,KK] = rpn_conv.weight, KK = kh*kw)
 * so the RPN conv that consumes qry_fmap_mod (fgn_ag_rpn_head.py:48, mmdet RPNHead [3P]) runs on the unmodified
 * query map with B*N weight sets; (1+N) * |qry| bytes of traffic per level disappear.  vec may hold the vectors
 * of several levels back to back (BN = L*B*N). */
int fgn_fold_attention_weights(const float *weight, const float *vec, int BN, int Co, int Ci, int KK,
                               float *out, void *stream);

/* The two calls above for a whole pyramid (FPN mode, SURVEY A.9: per level, the vector comes from
 * the same level's support maps) in three launches.  NHWC storage only.
 *   spp: level l = [B*N*K,h_l,w_l,C]; vec [L,B*N,C];  qry: level l = [B,H_l,W_l,C];
 *   out_host: HOST array of L device pointers, level l = [B*N,H_l,W_l,C]. */
size_t fgn_attention_vectors_ml_workspace_bytes(const fgn_pyramid_t *spp, int BN, int K, int C);
int fgn_attention_vectors_ml(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec,
                             void *workspace, size_t workspace_bytes, void *stream);
int fgn_channel_attention_ml(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                             float *const *out_host, void *stream);

/* AGRPNHead best-class selection (fgn_ag_rpn_head.py:87-108): per anchor position take the
 * score and the 4 deltas of the class with the largest score (first max wins).
 *   cls [B*N,A,H,W], reg [B*N,4A,H,W] (NCHW) -> cls_out [B,A,H,W], reg_out [B,4A,H,W]. */
int fgn_best_class_select(const float *cls, const float *reg, int B, int N, int A, int H, int W,
                          float *cls_out, float *reg_out, void *stream);

/* Relation-Guided Detector: count_one_roi_by_n_spp (fgn_roi_head.py:253-279) + BBoxHead.forward
 * with_avg_pool [3P] (:338) + count_modified_cls_bbox (:302-326), fused.  The [R*N,2C,P,P]
 * concat is never built: conv1x1(cat(q,s)) = Wq q + Ws s + b.
 *   roi_feat [R,C,P,P] in feat_layout; roi_batch [R] int32 (rois[:,0]);
 *   spp_cat_mean [B*N,C,P,P] in feat_layout (count_spp's first output);
 *   conv_w [C,2C] (Conv2d(2C->C,1x1).weight), conv_b [C]; gn_w, gn_b [C] (GroupNorm(32,C));
 *   fc_cls_w [2,C], fc_cls_b [2]; fc_reg_w [4,C], fc_reg_b [4];
 *   cls_out [R,N+1], reg_out [R,4N];  raw_cls_out/raw_reg_out optional ([R*N,2]/[R*N,4]).
 *   precision: 0 = fp32 parity (3xTF32 error-compensated tcgen05 contraction),
 *              1 = single-pass TF32 on tcgen05 (reduced precision, reported separately).
 *   class_term (optional) [B*N*P*P, C]: the class half of the split convolution, Ys = spp_cat_mean Ws^T + conv_b, as
 *              fgn_support_prologue_fwd leaves it; when given, spp_cat_mean may be NULL and one launch is saved.
 * workspace: fgn_relation_fusion_workspace_bytes(R, BN, C, P) bytes. */
size_t fgn_relation_fusion_workspace_bytes(int R, int BN, int C, int P);
/* Load-time preparation of the relation conv's weights: the TF32 hi/lo split of Wq = conv_w[:, :C] and
 * Ws = conv_w[:, C:] that the 3xTF32 contraction consumes (out: fgn_relation_split_weights_bytes(C) bytes, valid as
 * long as conv_w is unchanged).  Passing it as conv_w_split below removes two launches per call; NULL = split per call. */
size_t fgn_relation_split_weights_bytes(int C);
int fgn_relation_split_weights(const float *conv_w, int C, float *out, void *stream);
int fgn_relation_fusion_fwd(const float *roi_feat, int feat_layout, const int32_t *roi_batch,
                            const float *spp_cat_mean, const float *class_term /* optional */, int R, int B, int N, int C, int P,
                            const float *conv_w, const float *conv_w_split /* optional */, const float *conv_b,
                            const float *gn_w, const float *gn_b, int gn_groups, float gn_eps,
                            const float *fc_cls_w, const float *fc_cls_b,
                            const float *fc_reg_w, const float *fc_reg_b,
                            float *cls_out, float *reg_out, float *raw_cls_out, float *raw_reg_out,
                            int precision, void *workspace, size_t workspace_bytes, void *stream);

/* The relation head's contraction alone (Conv2d(2C->C,1x1) as a GEMM, fgn_roi_head.py:272):
 *   C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]); row-major fp32 with leading dimensions lda/ldb/ldc.
 * precision as in fgn_relation_fusion_fwd.  workspace: fgn_gemm_workspace_bytes(N, K) bytes;
 * without it (or for shapes outside K%32==0, N%16==0) the fp32 SIMT kernel is used. */
size_t fgn_gemm_workspace_bytes(int N, int K);
int fgn_gemm_nt(const float *A, int lda, const float *B, int ldb, const float *bias, float *C, int ldc,
                int M, int N, int K, int precision, void *workspace, size_t workspace_bytes, void *stream);

/* Post-RoI head 1x1 convolutions on NHWC RoI tiles (the C4 res5 shared_head's conv1 / conv3,
 * fgn_roi_head.py:202-238, and FCNMaskHead's conv_logits [3P], fgn_r50_c4_densecl.py:115-129) through the same tcgen05
 * contraction:  out[M,Cout] = [relu]( x[M,Cin] weight[Cout,Cin]^T + bias [+ residual[M,Cout]] ),  M = R*H*W rows.
 * BatchNorm (eval) is folded into weight / bias by the caller.  precision as in fgn_gemm_nt (0 = 3xTF32 fp32 parity,
 * 1 = single-pass TF32: what cuDNN runs the neighbouring 3x3 convolution in when torch.backends.cudnn.allow_tf32 is set).
 * workspace: fgn_gemm_workspace_bytes(Cout, Cin). */
int fgn_conv1x1_nhwc(const float *x, const float *weight, const float *bias, const float *residual, int relu,
                     float *out, int M, int Cin, int Cout, int precision, void *workspace, size_t workspace_bytes,
                     void *stream);

/* fgn_gemm_nt at precision 0 with the TF32 hi/lo split of B made beforehand: b_split = { B_hi [N,K], B_lo [N,K] } as
 * fgn_conv_split_weights(B, 1, N, K, b_split) writes it (B dense).  No per-call split launch, no workspace. */
int fgn_gemm_nt_presplit(const float *A, int lda, const float *b_split, const float *bias, float *C, int ldc,
                         int M, int N, int K, void *stream);

/* Post-RoI head 3x3 / pad 1 convolution over NHWC RoI tiles as an implicit GEMM on tcgen05 (no im2col buffer: tap
 * (dy,dx) reads the activation tensor through a 4D TMA descriptor shifted by (dx-1, dy-1), out-of-tile cells are
 * zero-filled by the TMA unit).  Replaces mmdet Bottleneck.conv2 of the C4 shared_head (fgn_roi_head.py:202-233) and
 * FCNMaskHead.convs[i] (fgn_r50_c4_densecl.py:115-129) at inference.
 *   x [R,H,W,Cin] NHWC; w_taps [9,Cout,Cin] = weight.permute(2,3,0,1) with any BatchNorm folded in (tap = ky*3+kx);
 *   w_split optional: fgn_conv_split_weights(w_taps, 9, Cout, Cin) (else split per call into workspace, which must then
 *   hold fgn_conv_split_weights_bytes(9, Cout, Cin)); bias [Cout] / residual [R,H,W,Cout] optional; out [R,H,W,Cout].
 *   precision 0 = 3xTF32 (fp32 parity), 1 = one TF32 pass (what cuDNN does under allow_tf32).
 * Needs Cin%16==0, Cout%16==0, (Cout<=256 or Cout%256==0), W<=128; else FGN_ERR_UNSUPPORTED. */
size_t fgn_conv_split_weights_bytes(int taps, int Cout, int Cin);
int fgn_conv_split_weights(const float *w_taps, int taps, int Cout, int Cin, float *out, void *stream);
int fgn_conv3x3_nhwc(const float *x, const float *w_taps, const float *w_split, const float *bias,
                     const float *residual, int relu, float *out, int R, int H, int W, int Cin, int Cout,
                     int precision, void *workspace, size_t workspace_bytes, void *stream);

/* FCNMaskHead's tail in one launch (mmdet FCNMaskHead.forward [3P]: upsample = ConvTranspose2d(Cin, Cout, 2, stride=2),
 * ReLU, conv_logits = Conv2d(Cout, ncls, 1); called at fgn_roi_head.py:380): a [R*H*W, Cin] x [4*Cout, Cin]^T contraction
 * whose epilogue applies the deconv bias, the ReLU and the logits' dot product straight out of tensor memory -- the
 * upsampled [R,2H,2W,Cout] map is never written.
 *   x [R,H,W,Cin] NHWC; w_taps [4,Cout,Cin] = upsample.weight.permute(2,3,1,0) (tap = i*2+j, output pixel (2h+i, 2w+j));
 *   b_deconv [Cout]; w_logits [ncls,Cout]; b_logits [ncls]; mask_pred [R,ncls,2H,2W] (NCHW, the reference's layout).
 * Needs Cin%16==0, Cout%16==0, Cout<=256, ncls<=4. */
int fgn_deconv2x2_logits_nhwc(const float *x, const float *w_taps, const float *w_split, const float *b_deconv,
                              const float *w_logits, const float *b_logits, float *mask_pred, int R, int H, int W,
                              int Cin, int Cout, int ncls, int precision, void *workspace, size_t workspace_bytes,
                              void *stream);

/* bf16 variant of the contraction (reported separately, never the fp32 default): A [M,K] and B [N,K]
 * hold bf16 (uint16 storage), accumulation and C are fp32; tcgen05 kind::f16, one pass.
 * Needs K%64==0, N%16==0 (N<=256 or N%256==0), 16-byte aligned rows. */
int fgn_gemm_nt_bf16(const uint16_t *A, int lda, const uint16_t *B, int ldb, const float *bias, float *C,
                     int ldc, int M, int N, int K, void *stream);

/* count_modified_cls_bbox alone (fgn_roi_head.py:302-326), generalised from N in {1,3} to any N:
 *   raw_cls [R*N,2] (bg,fg), raw_reg [R*N,4] -> cls_out [R,N+1] = (fg_0..fg_{N-1}, bg of the
 *   first-max fg class), reg_out [R,4N]. */
int fgn_cls_bbox_reassemble(const float *raw_cls, const float *raw_reg, int R, int N,
                            float *cls_out, float *reg_out, void *stream);

/* FPN-mode single call (no shared_head between RoIAlign and the relation conv): level assignment + RoIAlign +
 * relation fusion + heads, three kernels (RoIAlign, RoI contraction, epilogue; a fourth, the class-term contraction, when
 * class_term is not handed over).  The RoI features
 * (NHWC, R*49*C*4 bytes) and the conv output travel through the caller's workspace between them.  Measured with the
 * kernels chained and the caches left alone (profiles/r02_chained_fused_call_ncu.txt): per 1000 RoIs the contraction
 * re-reads ~30 of the 50 MB of RoI features from DRAM (the rest is still in L2), the epilogue finds the conv output in L2.  A single kernel
 * that keeps them on chip was built and measured slower.  Same arguments as the two calls it replaces. */
size_t fgn_guided_roi_fused_workspace_bytes(int R, int BN, int C, int P);
int fgn_guided_roi_fused_fwd(const fgn_pyramid_t *pyr, int B, int C, const float *rois, int R,
                             int P, int sampling_ratio, int aligned, float finest_scale,
                             const float *spp_cat_mean /* [B*N,P,P,C] NHWC */, const float *class_term /* optional */, int N,
                             const float *conv_w, const float *conv_w_split /* optional */, const float *conv_b,
                             const float *gn_w, const float *gn_b, int gn_groups, float gn_eps,
                             const float *fc_cls_w, const float *fc_cls_b,
                             const float *fc_reg_w, const float *fc_reg_b,
                             float *cls_out, float *reg_out, int32_t *lvl_out,
                             int precision, void *workspace, size_t workspace_bytes, void *stream);

/* ---- bf16 variant (reported separately; the fp32 entry points above are the parity contract) ----
 * Pyramid levels hold bf16 in NHWC storage (fgn_pyramid_t.feat reinterpreted as uint16 pointers);
 * coordinates, weights and accumulation stay fp32, so level assignment and sample indices are the
 * same bit-exact contract.  out is bf16 (out_is_bf16 = 1) or fp32, NHWC [R,P,P,C].  Needs C%8==0. */
int fgn_roi_align_ml_fwd_bf16(const fgn_pyramid_t *pyr, int B, int C, const float *rois, int R, int P,
                              int sampling_ratio, int aligned, float finest_scale,
                              const float *chan_scale, const int32_t *scale_index, void *out,
                              int out_is_bf16, int32_t *lvl_out, void *stream);
int fgn_attention_vectors_ml_bf16(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec,
                                  void *workspace, size_t workspace_bytes, void *stream);   /* vec stays fp32 */
int fgn_channel_attention_ml_bf16(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                                  void *const *out_host, void *stream);                     /* bf16 in, bf16 out */
size_t fgn_guided_roi_fused_bf16_workspace_bytes(int R, int BN, int C, int P);
int fgn_guided_roi_fused_fwd_bf16(const fgn_pyramid_t *pyr, int B, int C, const float *rois, int R,
                                  int P, int sampling_ratio, int aligned, float finest_scale,
                                  const float *spp_cat_mean /* [B*N,P,P,C] NHWC fp32 */, int N,
                                  const float *conv_w, const float *conv_b,
                                  const float *gn_w, const float *gn_b, int gn_groups, float gn_eps,
                                  const float *fc_cls_w, const float *fc_cls_b,
                                  const float *fc_reg_w, const float *fc_reg_b,
                                  float *cls_out, float *reg_out, int32_t *lvl_out,
                                  void *workspace, size_t workspace_bytes, void *stream);

/* ---- gradients (what the training configs need under autograd) ----------
 * All feature tensors NHWC fp32.  The reference obtains these through autograd over mmcv /
 * torchvision / ATen ops (fgn_roi_head.py:344-358,451-529). */

/* Adjoint of fgn_roi_align_ml_fwd w.r.t. the feature maps: grad_pyr->feat[l] is the gradient buffer of
 * level l ([B,H_l,W_l,C], zero-initialised or holding a running sum: the kernel accumulates with
 * vector atomics); grad_out [R,P,P,C].  Same level assignment / sample indices as the forward. */
int fgn_roi_align_ml_bwd(const fgn_pyramid_t *grad_pyr, int B, int C, const float *rois, int R, int P,
                         int sampling_ratio, int aligned, float finest_scale,
                         const float *chan_scale, const int32_t *scale_index,
                         const float *grad_out, void *stream);
/* Deterministic form of fgn_roi_align_ml_bwd: every contribution is rounded once to 64-bit fixed point (scale chosen from
 * max |grad_out| and max |chan_scale|: 2^-40 of the largest term in resolution) and added with integer atomics, so the
 * gradient maps are bit-identical from run to run whatever order the RoIs' CTAs run in (float atomics -- this library's
 * default and mmcv's own backward -- are not).  chan_scale_rows: rows of chan_scale (0 if NULL).
 * workspace: fgn_roi_align_ml_bwd_det_workspace_bytes(grad_pyr, B, C) bytes. */
size_t fgn_roi_align_ml_bwd_det_workspace_bytes(const fgn_pyramid_t *grad_pyr, int B, int C);
int fgn_roi_align_ml_bwd_det(const fgn_pyramid_t *grad_pyr, int B, int C, const float *rois, int R, int P,
                             int sampling_ratio, int aligned, float finest_scale, const float *chan_scale,
                             int chan_scale_rows, const int32_t *scale_index, const float *grad_out,
                             void *workspace, size_t workspace_bytes, void *stream);
/* Adjoint of fgn_channel_attention: grad_qry [B,H,W,C] = sum_n g*vec, grad_vec [B*N,C] = sum_hw g*qry;
 * either output may be NULL. */
size_t fgn_channel_attention_bwd_workspace_bytes(int B, int N, int C, int H, int W);
int fgn_channel_attention_bwd(const float *qry, const float *vec, const float *grad_out, int B, int N, int C,
                              int H, int W, float *grad_qry, float *grad_vec, void *workspace,
                              size_t workspace_bytes, void *stream);
/* Adjoint of fgn_attention_vectors: grad_spp [B*N*K,H,W,C] = grad_vec / (K*H*W). */
int fgn_attention_vectors_bwd(const float *grad_vec, int BN, int K, int C, int H, int W, float *grad_spp, void *stream);
/* Adjoint of fgn_support_pool: grad_f [B*N*K,P,P,C] = grad_cat/K + grad_gap*m/(K*P*P); grad_cat [B*N,P,P,C]
 * and grad_gap [B*N,C] may each be NULL. */
int fgn_support_pool_bwd(const float *grad_cat, const float *grad_gap, const float *m, int BN, int K, int C, int P,
                         float *grad_f, void *stream);

/* Adjoint of fgn_relation_fusion_fwd (the reference: autograd over fgn_roi_head.py:253-279,338,302-326 in forward_train,
 * :344-358).  All feature tensors NHWC.  Yq [R*49,C] = roi_feat Wq^T and Ys [B*N*49,C] = spp_cat_mean Ws^T + conv_b are
 * the forward's split-conv outputs (re-derive them with fgn_gemm_nt); cls_fwd [R,N+1] = the forward's cls_out (its
 * first-max foreground class receives the background column's gradient).  Every output may be NULL:
 * d_roi_feat [R,49,C], d_spp [B*N,49,C], d_conv_w [C,2C], d_conv_b [C], d_gn_w / d_gn_b [C], d_fc_cls_w [2,C],
 * d_fc_cls_b [2], d_fc_reg_w [4,C], d_fc_reg_b [4].  GroupNorm of <= 32 groups of a power-of-two (<= 32) channels. */
size_t fgn_relation_fusion_bwd_workspace_bytes(int R, int BN, int N, int C);
int fgn_relation_fusion_bwd(const float *roi_feat, const float *spp_cat_mean, const int32_t *roi_batch,
                            const float *Yq, const float *Ys, const float *cls_fwd,
                            const float *d_cls, const float *d_reg, int R, int B, int N, int C, int P,
                            const float *conv_w, const float *gn_w, const float *gn_b, int gn_groups, float gn_eps,
                            const float *fc_cls_w, const float *fc_reg_w,
                            float *d_roi_feat, float *d_spp, float *d_conv_w, float *d_conv_b,
                            float *d_gn_w, float *d_gn_b, float *d_fc_cls_w, float *d_fc_cls_b,
                            float *d_fc_reg_w, float *d_fc_reg_b,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Test-time step between the relation head's outputs and the mask branch: BBoxHead.get_bboxes [3P, mmdet 2.18]
 * as called from fgn_roi_head.py:606-613 with test_cfg.rcnn (fgn_r50_c4_densecl.py:181-185) -- softmax over the
 * N+1 scores (background last), DeltaXYWHBBoxCoder.decode of the per-class deltas (bbox_coder, :91-94) against the
 * proposals, clip to img_shape, optional division by scale_factor (rescale=True), multiclass_nms: score > score_thr,
 * class-aware NMS through mmcv batched_nms' coordinate-offset trick, the max_per_img best by score.
 *   rois [R,5] grouped by image (bbox2roi), cls_score [R,N+1], bbox_pred [R,4N];
 *   img_offsets [B+1] int32 (device): first RoI of every image; Rmax = largest per-image RoI count;
 *   img_hw [B,2] (h,w) or NULL (no clipping); scale_factor [B,4] or NULL (no rescale);
 *   means/stds: HOST pointers to 4 floats; wh_ratio_clip = 16/1000 in mmdet.
 *   det_out [B,max_per_img,5] (x1,y1,x2,y2,score), label_out [B,max_per_img], count_out [B] valid rows per image.
 * Integer contract: which (RoI, class) pairs are kept and in what order. */
size_t fgn_det_postprocess_workspace_bytes(int R, int N, int B, int Rmax);
int fgn_det_postprocess(const float *rois, const float *cls_score, const float *bbox_pred,
                        const int32_t *img_offsets, int R, int N, int B, int Rmax,
                        const float *img_hw, const float *scale_factor,
                        const float *means, const float *stds, float wh_ratio_clip,
                        float score_thr, float iou_thr, int max_per_img,
                        float *det_out, int32_t *label_out, int32_t *count_out,
                        void *workspace, size_t workspace_bytes, void *stream);

/* RPN proposals: RPNHead.get_bboxes / _get_bboxes_single [3P, mmdet 2.18] as called from fgn.py:229-235 on the
 * (best-class-selected, fgn_ag_rpn_head.py:87-113) outputs of AGRPNHead with test_cfg.rpn
 * (fgn_r50_c4_densecl.py:175-180) -- per level: sigmoid scores in (H,W,A) order, the nms_pre best, DeltaXYWHBBoxCoder
 * decode against the grid anchors (base_anchors[l,a] + (x,y,x,y)*stride_l), clip to img_hw, w/h > min_bbox_size
 * (negative: no filter), per-level NMS (mmcv batched_nms with ids = level), the max_per_img best by score.
 *   cls[l] [B,A,H_l,W_l], reg[l] [B,4A,H_l,W_l] (device pointers in HOST arrays; H, W, strides HOST int arrays);
 *   base_anchors [L,A,4] device (mmdet AnchorGenerator.base_anchors); means/stds HOST float[4];
 *   prop_out [B,max_per_img,5] (x1,y1,x2,y2,score), level_out [B,max_per_img], count_out [B].
 * The pre-NMS selection is a hand-written two-level radix select + shared-memory bitonic sort for nms_pre <= 8192
 * (every test config); larger nms_pre falls back to cub::DeviceSegmentedRadixSort. */
size_t fgn_rpn_proposals_workspace_bytes(const int *H, const int *W, int L, int A, int B, int nms_pre);
int fgn_rpn_proposals(const float *const *cls, const float *const *reg, const int *H, const int *W,
                      const int *strides, int L, int A, int B, const float *base_anchors,
                      const float *img_hw, const float *means, const float *stds, float wh_ratio_clip,
                      int nms_pre, float iou_thr, int max_per_img, float min_bbox_size,
                      float *prop_out, int32_t *level_out, int32_t *count_out,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Test-time mask pasting fused with the result encoding (second part):
 * FCNMaskHead.get_seg_masks / _do_paste_mask [3P, mmdet 2.18] as called from fgn_roi_head.py:668-671 (sigmoid,
 * F.grid_sample(bilinear, zeros padding, align_corners=False) of every [M,M] mask over the whole image,
 * ">= mask_thr_binary", test_cfg.rcnn fgn_r50_c4_densecl.py:186), then mmdet.core.encode_mask_results ->
 * pycocotools mask.encode [3P] as called from fgn.py:281 (column-major run lengths + their compressed string).
 * The dense [D,img_h,img_w] masks are never materialised: the pixels a box can touch are walked in column-major
 * order by as many CTAs as the box needs (32 rows of a column per thread and step) and the runs are emitted directly;
 * a second kernel orders the pieces, takes differences and writes the string.  Hmax, Wmax: the largest image in
 * img_hw (sizes the launch and the workspace).
 *   mask_pred [D,M,M] logits of a class-agnostic mask head (fgn_r50_c4_densecl.py:123,127; labels forced to 0,
 *   fgn_roi_head.py:716); boxes: D rows of box_stride floats starting with x1,y1,x2,y2 (box_stride 5 takes
 *   det_bboxes as they are); det_img [D] int32 image of every detection or NULL (all image 0); img_hw [B,2] int32
 *   (h,w) of the pasted masks (ori_shape when rescale, else the scaled shape).
 *   counts_out [D,cap] int32 run lengths (first run = zeros, may be 0); ncounts_out [D] runs, or -(runs needed)
 *   when cap is too small; str_out [D,cap_bytes] the pycocotools "counts" string (NULL: skip), strlen_out [D]
 *   its length, or -(bytes needed).
 * Integer contract: identical runs to the reference wherever no pixel value lies within fp32 rounding of the
 * threshold. */
size_t fgn_mask_paste_rle_workspace_bytes(int D, int cap, int Hmax, int Wmax);
int fgn_mask_paste_rle(const float *mask_pred, const float *boxes, int box_stride,
                       const int32_t *det_img, const int32_t *img_hw, int D, int M, float mask_thr,
                       int Hmax, int Wmax, int32_t *counts_out, int32_t *ncounts_out, unsigned char *str_out,
                       int32_t *strlen_out, int cap, int cap_bytes,
                       void *workspace, size_t workspace_bytes, void *stream);

/* encode_mask_results of GIVEN masks (fgn.py:296-298: the ground-truth qry_isegmaps travel in the result dict as COCO
 * RLE, pycocotools mask.encode [3P]): masks [D,H,W] bytes (0 / non-zero) -> the same outputs as fgn_mask_paste_rle,
 * through the same run-length / string encoder.  workspace: fgn_mask_paste_rle_workspace_bytes(D, cap, H, W). */
int fgn_mask_rle_encode(const unsigned char *masks, int D, int H, int W, int32_t *counts_out, int32_t *ncounts_out,
                        unsigned char *str_out, int32_t *strlen_out, int cap, int cap_bytes,
                        void *workspace, size_t workspace_bytes, void *stream);

/* get_seg_masks' own return value for one image: out [D,img_h,img_w] bytes (0/1). */
int fgn_mask_paste(const float *mask_pred, const float *boxes, int box_stride, int D, int M, int img_h,
                   int img_w, float mask_thr, unsigned char *out, void *stream);

/* Number of kernels this library has launched in the calling process since load
 * (bench.py's gpu_launches). */
uint64_t fgn_launch_count(void);

/* Test/debug export: self-check counter of the rotating-window RoIAlign planner -- the number of
 * (footprint row, bin row) weights that did not fit the register window.  Always 0 unless the
 * chunking rule in roi_align_window.cu is broken. */
unsigned int fgn_debug_roi_window_violations(void);

#ifdef __cplusplus
}
#endif
#endif /* FGN_B200_H_ */

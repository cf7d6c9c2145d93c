// detect.cu -- the step between the relation head's outputs (a8) and the mask branch (a9) at test time:
// BBoxHead.get_bboxes [3P, mmdet 2.18] as called from fgn_roi_head.py:606-613 -- softmax over the N+1 scores,
// DeltaXYWHBBoxCoder.decode of the per-class deltas against the proposals (clip to the image, optional rescale),
// multiclass_nms (score threshold, class-aware NMS through mmcv's coordinate-offset trick, top max_per_img).
//
// Integer contract: the kept (RoI, class) pairs and their order.  The box / IoU arithmetic below is written
// operation by operation in the order the reference's torch expressions evaluate (explicitly rounded, no FMA
// contraction), so that IoU-vs-threshold decisions agree with the CPU restatement in oracle/fgn_oracle.py; the
// only library difference left is expf (softmax, exp of the size deltas), 1-2 ulp.
//
//   D1 decode   thread = RoI: softmax, N boxes, validity (score > score_thr), per-image max coordinate
//   D2 rank     per (image, class): rank of every valid candidate by (score desc, RoI index asc) -> sorted list
//   D3 mask     per (image, class): 64x64 tiles of the IoU > thr bit matrix on the OFFSET boxes
//   D4 reduce   per (image, class): one warp walks the sorted list and ORs the rows of kept boxes
//   D5 merge    thread = kept box: global rank = sum over classes of a binary search; writes the top max_per_img
#include "common.cuh"

namespace fgn {

namespace {

struct DetWs {                  // workspace carve-up (all offsets in bytes from the workspace base)
    size_t hdr, cand_box, cand_score, sorted, kept, mask, total;
    int    words;               // 64-bit mask words per row
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

DetWs det_layout(int R, int N, int B, int Rmax)
{
    DetWs w;
    w.words = (Rmax + 63) / 64;
    size_t o = 0;
    w.hdr = o;        o = align256(o + (size_t)B * (2 + 2 * N) * sizeof(unsigned int));     // maxc, total kept, count[N], kept[N]
    w.cand_box = o;   o = align256(o + (size_t)R * N * 4 * sizeof(float));
    w.cand_score = o; o = align256(o + (size_t)R * N * sizeof(float));
    w.sorted = o;     o = align256(o + (size_t)B * N * Rmax * sizeof(int));                 // RoI index (image-local), score order
    w.kept = o;       o = align256(o + (size_t)B * N * Rmax * sizeof(int));                 // positions in `sorted` that survive
    w.mask = o;       o = align256(o + (size_t)B * N * Rmax * w.words * sizeof(unsigned long long));
    w.total = o;
    return w;
}

// order-preserving float <-> uint map (for atomicMax on floats of either sign)
__device__ __forceinline__ unsigned int f2ord(float f)
{
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ unsigned int *hdr_of(unsigned char *ws, const DetWs &L, int b, int N)
{
    return reinterpret_cast<unsigned int *>(ws + L.hdr) + (size_t)b * (2 + 2 * N);
}

struct DetParams {
    const float *rois, *cls, *reg;      // [R,5], [R,N+1], [R,4N]
    const int   *img_off;               // [B+1] first RoI of every image
    const float *img_hw;                // [B,2] (h, w) clip bounds, or NULL: no clipping
    const float *scale;                 // [B,4] rescale divisors, or NULL
    int   R, N, B, Rmax, max_per_img;
    float means[4], stds[4], score_thr, iou_thr, max_ratio;
};

__device__ __forceinline__ int image_of(const DetParams &p, int r)
{
    int b = 0;
    while (b + 1 < p.B && r >= p.img_off[b + 1]) ++b;
    return b;
}

// ---- D1 -----------------------------------------------------------------------------------------------------
__global__ void det_decode_kernel(const DetParams p, unsigned char *ws, const DetWs L)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    const int b = image_of(p, r);
    const int N = p.N;
    // softmax over the N+1 scores, background last (torch: x - max, exp, sum in index order, divide)
    const float *x = p.cls + (size_t)r * (N + 1);
    float m = x[0];
    for (int i = 1; i <= N; ++i) m = fmaxf(m, x[i]);
    float sum = 0.f;
    for (int i = 0; i <= N; ++i) sum = __fadd_rn(sum, expf(__fsub_rn(x[i], m)));
    const float *roi = p.rois + (size_t)r * 5;
    const float px = __fmul_rn(__fadd_rn(roi[1], roi[3]), 0.5f), py = __fmul_rn(__fadd_rn(roi[2], roi[4]), 0.5f);
    const float pw = __fsub_rn(roi[3], roi[1]), ph = __fsub_rn(roi[4], roi[2]);
    float *cbox = reinterpret_cast<float *>(ws + L.cand_box);
    float *cscore = reinterpret_cast<float *>(ws + L.cand_score);
    unsigned int *hdr = hdr_of(ws, L, b, N);
    float mx = -3.0e38f;
    bool any = false;
    for (int n = 0; n < N; ++n) {
        const float score = __fdiv_rn(expf(__fsub_rn(x[n], m)), sum);
        const float *d = p.reg + (size_t)r * 4 * N + 4 * n;
        const float dx = __fadd_rn(__fmul_rn(d[0], p.stds[0]), p.means[0]);
        const float dy = __fadd_rn(__fmul_rn(d[1], p.stds[1]), p.means[1]);
        float dw = __fadd_rn(__fmul_rn(d[2], p.stds[2]), p.means[2]);
        float dh = __fadd_rn(__fmul_rn(d[3], p.stds[3]), p.means[3]);
        dw = fminf(fmaxf(dw, -p.max_ratio), p.max_ratio);
        dh = fminf(fmaxf(dh, -p.max_ratio), p.max_ratio);
        const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
        const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
        float x1 = __fsub_rn(gx, __fmul_rn(gw, 0.5f)), y1 = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
        float x2 = __fadd_rn(gx, __fmul_rn(gw, 0.5f)), y2 = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
        if (p.img_hw != nullptr) {
            const float H = p.img_hw[2 * b], W = p.img_hw[2 * b + 1];
            x1 = fminf(fmaxf(x1, 0.f), W); x2 = fminf(fmaxf(x2, 0.f), W);
            y1 = fminf(fmaxf(y1, 0.f), H); y2 = fminf(fmaxf(y2, 0.f), H);
        }
        if (p.scale != nullptr) {
            const float *s = p.scale + 4 * b;
            x1 = __fdiv_rn(x1, s[0]); y1 = __fdiv_rn(y1, s[1]); x2 = __fdiv_rn(x2, s[2]); y2 = __fdiv_rn(y2, s[3]);
        }
        const size_t c = (size_t)r * N + n;
        reinterpret_cast<float4 *>(cbox)[c] = make_float4(x1, y1, x2, y2);
        const bool valid = score > p.score_thr;
        cscore[c] = valid ? score : -1.f;
        if (valid) {
            any = true;
            mx = fmaxf(mx, fmaxf(fmaxf(x1, y1), fmaxf(x2, y2)));
            atomicAdd(&hdr[2 + n], 1u);                          // candidates of class n in image b
        }
    }
    if (any) atomicMax(&hdr[0], f2ord(mx));                      // boxes.max() of mmcv's batched_nms
}

// ---- D2 -----------------------------------------------------------------------------------------------------
// grid (ceil(Rmax / 256), N, B); rank of candidate i of class n among the valid candidates of its image and class
__global__ void det_rank_kernel(const DetParams p, unsigned char *ws, const DetWs L)
{
    __shared__ float tile[256];
    const int n = blockIdx.y, b = blockIdx.z, N = p.N;
    const int r0 = p.img_off[b], nr = p.img_off[b + 1] - r0;
    if ((int)(blockIdx.x * 256) >= nr) return;
    const float *cscore = reinterpret_cast<const float *>(ws + L.cand_score);
    const int i = blockIdx.x * 256 + threadIdx.x;
    const float si = i < nr ? cscore[(size_t)(r0 + i) * N + n] : -1.f;
    int rank = 0;
    for (int j0 = 0; j0 < nr; j0 += 256) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < nr ? cscore[(size_t)(r0 + j) * N + n] : -1.f;
        __syncthreads();
        const int lim = min(256, nr - j0);
        for (int t = 0; t < lim; ++t) {
            const float sj = tile[t];
            rank += (sj > si || (sj == si && j0 + t < i)) ? 1 : 0;    // invalid (-1) never outranks a valid score
        }
        __syncthreads();
    }
    if (i < nr && si >= 0.f)
        reinterpret_cast<int *>(ws + L.sorted)[((size_t)b * N + n) * p.Rmax + rank] = i;
}

// the box NMS sees: mmcv batched_nms adds label * (max_coordinate + 1) to all four coordinates
__device__ __forceinline__ float4 offset_box(const float4 bx, int n, float maxc)
{
    const float off = __fmul_rn((float)n, __fadd_rn(maxc, 1.f));
    return make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
}
__device__ __forceinline__ bool iou_over(const float4 a, const float4 c, float thr)
{
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_c = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
    const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_c), inter));
    return ovr > thr;                                             // NaN (0/0) compares false, as on the CPU
}

// ---- D3 -----------------------------------------------------------------------------------------------------
// grid (words, words, B*N), 64 threads: tile (row block, column word) of the IoU > thr matrix, columns > row
__global__ void det_nms_mask_kernel(const DetParams p, unsigned char *ws, const DetWs L)
{
    __shared__ float4 cols[64];
    const int n = blockIdx.z % p.N, b = blockIdx.z / p.N, N = p.N;
    const unsigned int *hdr = hdr_of(ws, L, b, N);
    const int m = (int)hdr[2 + n];
    const int rb = blockIdx.y, cw = blockIdx.x;
    if (rb * 64 >= m || cw * 64 >= m || cw < rb) return;
    const float maxc = ord2f(hdr[0]);
    const int r0 = p.img_off[b];
    const int *sorted = reinterpret_cast<const int *>(ws + L.sorted) + ((size_t)b * N + n) * p.Rmax;
    const float4 *cbox = reinterpret_cast<const float4 *>(ws + L.cand_box);
    const int t = threadIdx.x;
    const int cj = cw * 64 + t;
    if (cj < m) cols[t] = offset_box(cbox[(size_t)(r0 + sorted[cj]) * N + n], n, maxc);
    __syncthreads();
    const int ri = rb * 64 + t;
    if (ri >= m) return;
    const float4 me = offset_box(cbox[(size_t)(r0 + sorted[ri]) * N + n], n, maxc);
    unsigned long long bits = 0ull;
    const int lim = min(64, m - cw * 64);
    for (int j = (cw == rb ? t + 1 : 0); j < lim; ++j)
        if (iou_over(me, cols[j], p.iou_thr)) bits |= 1ull << j;
    reinterpret_cast<unsigned long long *>(ws + L.mask)[(((size_t)b * N + n) * p.Rmax + ri) * L.words + cw] = bits;
}

// ---- D4 -----------------------------------------------------------------------------------------------------
// grid (N, B), 256 threads: the class's bit matrix is staged in shared memory (as many rows as fit), then warp 0
// makes the greedy pass over the score-sorted list, ORing the rows of the boxes it keeps
__global__ void det_nms_reduce_kernel(const DetParams p, unsigned char *ws, const DetWs L, const int smem_rows)
{
    extern __shared__ unsigned long long sm[];                   // remv[words] | rows[smem_rows][words]
    const int n = blockIdx.x, b = blockIdx.y, N = p.N, t = threadIdx.x, lane = t & 31;
    unsigned int *hdr = hdr_of(ws, L, b, N);
    const int m = (int)hdr[2 + n];
    const int words = (m + 63) / 64, W = L.words;
    unsigned long long *remv = sm, *rows = sm + W;
    const unsigned long long *mask = reinterpret_cast<const unsigned long long *>(ws + L.mask) + ((size_t)b * N + n) * p.Rmax * W;
    for (int w = t; w < words; w += blockDim.x) remv[w] = 0ull;
    const int staged = min(m, smem_rows);
    for (int e = t; e < staged * words; e += blockDim.x) {
        const int i = e / words, w = e - i * words;
        if (w >= (i >> 6)) rows[(size_t)i * W + w] = mask[(size_t)i * W + w];     // words below the row's block were never written
    }
    __syncthreads();
    if (t >= 32) return;
    int *kept = reinterpret_cast<int *>(ws + L.kept) + ((size_t)b * N + n) * p.Rmax;
    int nk = 0;
    for (int wb = 0; wb < words; ++wb) {                         // 64 boxes at a time: their removed-bits word lives in a register
        unsigned long long cur = remv[wb];
        const int base = wb * 64, lim = min(64, m - base);
        for (int bit = 0; bit < lim; ++bit) {
            if ((cur >> bit) & 1ull) continue;
            const int i = base + bit;
            if (lane == 0) kept[nk] = i;
            ++nk;
            const unsigned long long *row = i < staged ? rows + (size_t)i * W : mask + (size_t)i * W;
            cur |= row[wb];                                      // the only load on the serial chain (uniform)
            for (int w = wb + 1 + lane; w < words; w += 32) remv[w] |= row[w];    // each lane owns its words within a block
        }
        __syncwarp();
    }
    if (lane == 0) { hdr[2 + N + n] = (unsigned)nk; atomicAdd(&hdr[1], (unsigned)nk); }
}

// ---- D5 -----------------------------------------------------------------------------------------------------
// grid (ceil(Rmax / 128), N, B): thread = kept box k of class n; its rank among all kept boxes of the image by
// (score desc, flat candidate index r*N+n asc) is a sum of binary searches over the classes' kept lists
__global__ void det_merge_kernel(const DetParams p, unsigned char *ws, const DetWs L, float *det_out,
                                 int32_t *label_out, int32_t *count_out)
{
    const int n = blockIdx.y, b = blockIdx.z, N = p.N;
    const unsigned int *hdr = hdr_of(ws, L, b, N);
    if (blockIdx.x == 0 && n == 0 && threadIdx.x == 0) count_out[b] = min((int)hdr[1], p.max_per_img);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int)hdr[2 + N + n]) return;
    const int r0 = p.img_off[b];
    const float *cscore = reinterpret_cast<const float *>(ws + L.cand_score);
    const int *sorted_all = reinterpret_cast<const int *>(ws + L.sorted) + (size_t)b * N * p.Rmax;
    const int *kept_all = reinterpret_cast<const int *>(ws + L.kept) + (size_t)b * N * p.Rmax;
    const int me = sorted_all[(size_t)n * p.Rmax + kept_all[(size_t)n * p.Rmax + k]];      // image-local RoI
    const float ms = cscore[(size_t)(r0 + me) * N + n];
    int rank = 0;
    for (int c = 0; c < N; ++c) {
        const int *ks = kept_all + (size_t)c * p.Rmax, *so = sorted_all + (size_t)c * p.Rmax;
        int lo = 0, hi = (int)hdr[2 + N + c];                   // first position in class c's list that does not beat me
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int rj = so[ks[mid]];
            const float sj = cscore[(size_t)(r0 + rj) * N + c];
            const bool better = sj > ms || (sj == ms && ((size_t)rj * N + c) < ((size_t)me * N + n));
            if (better) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    if (rank < p.max_per_img) {
        const float4 bx = reinterpret_cast<const float4 *>(ws + L.cand_box)[(size_t)(r0 + me) * N + n];
        float *o = det_out + ((size_t)b * p.max_per_img + rank) * 5;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = ms;
        label_out[(size_t)b * p.max_per_img + rank] = n;
    }
}

}  // namespace
}  // namespace fgn

using namespace fgn;

extern "C" size_t fgn_det_postprocess_workspace_bytes(int R, int N, int B, int Rmax)
{
    if (R <= 0 || N <= 0 || B <= 0 || Rmax <= 0) return 256;
    return det_layout(R, N, B, Rmax).total;
}

extern "C" int fgn_det_postprocess(const float *rois, const float *cls_score, const float *bbox_pred,
                                   const int32_t *img_offsets, int R, int N, int B, int Rmax,
                                   const float *img_hw, const float *scale_factor,
                                   const float *means, const float *stds, float wh_ratio_clip,
                                   float score_thr, float iou_thr, int max_per_img,
                                   float *det_out, int32_t *label_out, int32_t *count_out,
                                   void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(R >= 0 && N >= 1 && B >= 1 && Rmax >= 0 && max_per_img >= 1, "bad dims R=%d N=%d B=%d Rmax=%d max_per_img=%d",
                  R, N, B, Rmax, max_per_img);
    FGN_CHECK_ARG(count_out && det_out && label_out, "NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    FGN_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(int32_t) * B, st));
    if (R == 0 || Rmax == 0) return FGN_OK;
    FGN_CHECK_ARG(rois && cls_score && bbox_pred && img_offsets && means && stds, "NULL input");
    FGN_CHECK_ARG(wh_ratio_clip > 0.f, "wh_ratio_clip=%f", wh_ratio_clip);
    const DetWs L = det_layout(R, N, B, Rmax);
    FGN_CHECK_ARG(workspace && workspace_bytes >= L.total, "workspace %zu < %zu bytes", workspace_bytes, L.total);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    DetParams p;
    p.rois = rois; p.cls = cls_score; p.reg = bbox_pred; p.img_off = img_offsets; p.img_hw = img_hw; p.scale = scale_factor;
    p.R = R; p.N = N; p.B = B; p.Rmax = Rmax; p.max_per_img = max_per_img;
    for (int i = 0; i < 4; ++i) { p.means[i] = means[i]; p.stds[i] = stds[i]; }
    p.score_thr = score_thr; p.iou_thr = iou_thr;
    p.max_ratio = fabsf(logf(wh_ratio_clip));
    FGN_CUDA_OK(cudaMemsetAsync(ws + L.hdr, 0, (size_t)B * (2 + 2 * N) * sizeof(unsigned int), st));
    det_decode_kernel<<<ceil_div(R, 128), 128, 0, st>>>(p, ws, L);
    FGN_LAUNCH_OK();
    det_rank_kernel<<<dim3(ceil_div(Rmax, 256), N, B), 256, 0, st>>>(p, ws, L);
    FGN_LAUNCH_OK();
    det_nms_mask_kernel<<<dim3(L.words, L.words, B * N), 64, 0, st>>>(p, ws, L);
    FGN_LAUNCH_OK();
    {
        // stage as many mask rows as fit in 200 KB of shared memory next to the removed-bits vector
        const size_t row_bytes = (size_t)L.words * sizeof(unsigned long long);
        const int smem_rows = (int)min((size_t)Rmax, (200 * 1024 - row_bytes) / row_bytes);
        const size_t smem = row_bytes * (1 + (size_t)smem_rows);
        static size_t attr = 48 * 1024;
        if (smem > attr) {
            FGN_CUDA_OK(cudaFuncSetAttribute(det_nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = smem;
        }
        det_nms_reduce_kernel<<<dim3(N, B), 256, smem, st>>>(p, ws, L, smem_rows);
        FGN_LAUNCH_OK();
    }
    det_merge_kernel<<<dim3(ceil_div(Rmax, 128), N, B), 128, 0, st>>>(p, ws, L, det_out, label_out, count_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

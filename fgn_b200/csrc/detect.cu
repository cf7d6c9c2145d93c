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
#include <cub/device/device_segmented_radix_sort.cuh>

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
    int   class_major_ties;             // equal scores: class (level) index first, then row (RPN); else row first
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
    extern __shared__ unsigned long long sm[];                   // remv[W] | diag[Rmax] | rows[smem_rows][W]
    const int n = blockIdx.x, b = blockIdx.y, N = p.N, t = threadIdx.x, lane = t & 31;
    unsigned int *hdr = hdr_of(ws, L, b, N);
    const int m = (int)hdr[2 + n];
    const int words = (m + 63) / 64, W = L.words;
    unsigned long long *remv = sm, *diag = sm + W, *rows = diag + p.Rmax;
    const unsigned long long *mask = reinterpret_cast<const unsigned long long *>(ws + L.mask) + ((size_t)b * N + n) * p.Rmax * W;
    for (int w = t; w < words; w += blockDim.x) remv[w] = 0ull;
    // the word a box needs at once when it is kept (its own 64-box block) is always served from shared memory;
    // the rest of a kept box's row is ORed in off the serial chain (shared memory for the first rows, global after)
    for (int i = t; i < m; i += blockDim.x) diag[i] = mask[(size_t)i * W + (i >> 6)];
    const int staged = min(m, smem_rows);
    for (int e = t; e < staged * words; e += blockDim.x) {
        const int i = e / words, w = e - i * words;
        if (w > (i >> 6)) rows[(size_t)i * W + w] = mask[(size_t)i * W + w];      // words below the row's block were never written
    }
    __syncthreads();
    if (t >= 32) return;
    int *kept = reinterpret_cast<int *>(ws + L.kept) + ((size_t)b * N + n) * p.Rmax;
    int nk = 0;
    for (int wb = 0; wb < words && nk < p.max_per_img; ++wb) {   // 64 boxes at a time: their removed-bits word lives in a register
        unsigned long long cur = remv[wb];
        const int base = wb * 64, lim = min(64, m - base);
        for (int bit = 0; bit < lim && nk < p.max_per_img; ++bit) {
            if ((cur >> bit) & 1ull) continue;
            const int i = base + bit;
            if (lane == 0) kept[nk] = i;
            ++nk;
            cur |= diag[i];                                      // the only load on the serial chain (uniform, shared)
            const unsigned long long *row = i < staged ? rows + (size_t)i * W : mask + (size_t)i * W;
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
            const bool before = p.class_major_ties ? (c < n || (c == n && rj < me)) : (((size_t)rj * N + c) < ((size_t)me * N + n));
            const bool better = sj > ms || (sj == ms && before);
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


// ================================================================================================================
// RPN proposals: RPNHead._get_bboxes_single [3P, mmdet 2.18] as called from fgn.py:229-235 on the outputs of
// AGRPNHead.forward_single -- per level: sigmoid scores in (H, W, A) order, the nms_pre best (stable descending),
// DeltaXYWHBBoxCoder.decode against the level's grid anchors, clip, min size filter; then NMS per level through
// the same coordinate-offset trick (ids = level), and the max_per_img best by score.  The candidates are laid out
// like the detections above with "class" = level and "RoI" = position in the level's sorted list, so kernels
// D3 - D5 are reused as they are.
struct RpnParams {
    const float *cls[FGN_MAX_LEVELS], *reg[FGN_MAX_LEVELS];   // [B,A,H,W], [B,4A,H,W]
    int   H[FGN_MAX_LEVELS], W[FGN_MAX_LEVELS], stride[FGN_MAX_LEVELS];
    int   seg_off[FGN_MAX_LEVELS + 1];                         // first sort item of level l within one image
    const float *base_anchors;                                 // [L,A,4]
    int   L, A, B, K;                                          // K = candidates kept per level (<= nms_pre)
    float min_size;                                            // < 0: no filter
};

// R1: sort keys (logits: monotone in the sigmoid score and free of its rounding ties) and anchor indices
__global__ void rpn_keys_kernel(const RpnParams q, float *keys, int *vals, int *seg_begin, int *seg_end, int *img_off)
{
    const int per_img = q.seg_off[q.L];
    const size_t total = (size_t)q.B * per_img;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; keys != nullptr && i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_img), j = (int)(i - (size_t)b * per_img);
        int l = 0;
        while (l + 1 < q.L && j >= q.seg_off[l + 1]) ++l;
        const int idx = j - q.seg_off[l];                      // (y*W + x)*A + a
        const int a = idx % q.A, cell = idx / q.A;
        keys[i] = q.cls[l][((size_t)b * q.A + a) * q.H[l] * q.W[l] + cell];
        vals[i] = idx;
    }
    if (blockIdx.x == 0) {
        for (int sgm = threadIdx.x; sgm < q.B * q.L; sgm += blockDim.x) {
            const int b = sgm / q.L, l = sgm % q.L;
            seg_begin[sgm] = b * per_img + q.seg_off[l];
            seg_end[sgm] = b * per_img + q.seg_off[l + 1];
        }
        for (int b = threadIdx.x; b <= q.B; b += blockDim.x) img_off[b] = b * q.K;
    }
}


// ---- hand-written segmented top-K (K <= kSelCap) for the pre-NMS selection ----------------------------------------
// One segment = one (image, level).  The K best logits by (value desc, anchor index asc) are found with a two-level
// radix select on the order-preserving 32-bit image of the float (16 + 16 bits, histograms in global memory), the
// winners are collected in anchor-index order by one block per segment (ties at the threshold value resolve to the
// lowest indices, exactly like a stable sort) and sorted in shared memory (bitonic).  No library call.
constexpr int kSelCap = 8192;                      // candidates one block sorts in shared memory
constexpr int kSelBins = 65536;

__device__ __forceinline__ float rpn_logit(const RpnParams &q, int b, int l, int idx)
{
    const int a = idx % q.A, cell = idx / q.A;
    return q.cls[l][((size_t)b * q.A + a) * q.H[l] * q.W[l] + cell];
}

// S1: histogram of the high 16 bits; grid (blocks, L, B)
__global__ void rpn_hist_hi_kernel(const RpnParams q, unsigned int *hist)
{
    const int l = blockIdx.y, b = blockIdx.z, m = q.seg_off[l + 1] - q.seg_off[l];
    unsigned int *h = hist + ((size_t)b * q.L + l) * kSelBins;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m; idx += gridDim.x * blockDim.x)
        atomicAdd(&h[f2ord(rpn_logit(q, b, l, idx)) >> 16], 1u);
}

// block-wide: largest bin T with  sum_{bin > T} h[bin] < want <= sum_{bin >= T} h[bin];  returns T and the count above it
__device__ void select_bin(const unsigned int *h, unsigned int want, unsigned int *sm /*[1024 + 2]*/, unsigned int &T, unsigned int &above)
{
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;   // blockDim.x == 1024: thread t owns bins [lo, lo + 64), t = 0 at the top
    const int lo = kSelBins - 64 * (t + 1);
    unsigned int mine = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint4 u = reinterpret_cast<const uint4 *>(h + lo)[i];
        mine += u.x + u.y + u.z + u.w;
    }
    // exclusive prefix over threads (top bins first): warp scan, then the 32 warp totals
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = sm[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
        sm[32 + lane] = wi - w;                                   // exclusive warp offsets
    }
    __syncthreads();
    const unsigned int before = sm[32 + wid] + incl - mine;      // items in bins above this thread's range
    __syncthreads();
    if (before < want && before + mine >= want) {                // exactly one thread
        unsigned int cum = before;
        int bin = lo + 63;
        for (int i = 63; i >= 0; --i) { const unsigned int c = h[lo + i]; if (cum + c >= want) { bin = lo + i; break; } cum += c; }
        sm[1024] = (unsigned)bin; sm[1025] = cum;
    }
    __syncthreads();
    T = sm[1024]; above = sm[1025];
    __syncthreads();
}

// S2: every block finds the segment's high bin T (redundantly), then histograms the low 16 bits inside it
__global__ void __launch_bounds__(1024) rpn_hist_lo_kernel(const RpnParams q, const unsigned int *hist_hi, unsigned int *hist_lo)
{
    __shared__ unsigned int sm[1026];
    const int l = blockIdx.y, b = blockIdx.z, m = q.seg_off[l + 1] - q.seg_off[l];
    if (m <= q.K) return;                           // everything is kept: no threshold
    const size_t sgm = (size_t)b * q.L + l;
    unsigned int T, above;
    select_bin(hist_hi + sgm * kSelBins, (unsigned)q.K, sm, T, above);
    unsigned int *h = hist_lo + sgm * kSelBins;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m; idx += gridDim.x * blockDim.x) {
        const unsigned int o = f2ord(rpn_logit(q, b, l, idx));
        if ((o >> 16) == T) atomicAdd(&h[o & 0xffffu], 1u);
    }
}

// S3: one block per segment: exact threshold value, winners collected in index order, bitonic sort, output
__global__ void __launch_bounds__(1024) rpn_select_sort_kernel(const RpnParams q, const unsigned int *hist_hi,
                                                              const unsigned int *hist_lo, float *keys_out, int *vals_out)
{
    extern __shared__ unsigned int dyn[];           // ord[kSelCap] | idx[kSelCap]
    __shared__ unsigned int sm[1026];
    __shared__ int s_warp[32];
    __shared__ int s_base, s_eq_taken;
    unsigned int *sord = dyn;
    int *sidx = reinterpret_cast<int *>(dyn + kSelCap);
    const int l = blockIdx.x, b = blockIdx.y, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int m = q.seg_off[l + 1] - q.seg_off[l];
    const int K = min(q.K, m);
    const size_t sgm = (size_t)b * q.L + l;
    unsigned int thr = 0, need_eq = 0;              // keep ord > thr, and the first need_eq (by index) with ord == thr
    if (m > q.K) {
        unsigned int T, above, Tl, above_l;
        select_bin(hist_hi + sgm * kSelBins, (unsigned)q.K, sm, T, above);
        select_bin(hist_lo + sgm * kSelBins, (unsigned)q.K - above, sm, Tl, above_l);
        thr = (T << 16) | Tl;
        need_eq = (unsigned)q.K - above - above_l;
    }
    if (t == 0) { s_base = 0; s_eq_taken = 0; }
    __syncthreads();
    // block-wide exclusive scan of one int per thread (returns the prefix; *total = block sum)
    auto block_scan = [&](int v, int *total) {
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) { const int c = s_warp[w]; if (w < wid) woff += c; tot += c; }
        __syncthreads();
        *total = tot;
        return woff + incl - v;
    };
    constexpr int kPer = 8;                          // consecutive anchor indices per thread and step
    for (int i0 = 0; i0 < m; i0 += 1024 * kPer) {   // index order: the equal-valued winners are the lowest indices
        const int first = i0 + t * kPer;
        unsigned int o[kPer];
        int ce = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int idx = first + j;
            o[j] = idx < m ? f2ord(rpn_logit(q, b, l, idx)) : 0u;
            ce += (idx < m && m > q.K && o[j] == thr) ? 1 : 0;
        }
        int eq_total, tk_total;
        int eq_rank = s_eq_taken + block_scan(ce, &eq_total);
        unsigned takes = 0;
        int ct = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int idx = first + j;
            const bool gt = idx < m && (m <= q.K || o[j] > thr), eq = idx < m && m > q.K && o[j] == thr;
            const bool take = gt || (eq && (unsigned)eq_rank < need_eq);
            eq_rank += eq ? 1 : 0;
            if (take) { takes |= 1u << j; ++ct; }
        }
        int pos = s_base + block_scan(ct, &tk_total);
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            if ((takes >> j) & 1u) { if (pos < kSelCap) { sord[pos] = o[j]; sidx[pos] = first + j; } ++pos; }
        __syncthreads();
        if (t == 0) { s_base += tk_total; s_eq_taken += eq_total; }
        __syncthreads();
    }
    // pad to a power of two with sentinels that sort last, then bitonic sort by (ord desc, idx asc)
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    for (int i = K + t; i < n2; i += 1024) { sord[i] = 0u; sidx[i] = 0x7fffffff; }
    __syncthreads();
    for (int k2 = 2; k2 <= n2; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = t; i < n2; i += 1024) {
                const int p2 = i ^ j;
                if (p2 > i) {
                    const unsigned int oa = sord[i], ob = sord[p2];
                    const int ia = sidx[i], ib = sidx[p2];
                    const bool a_first = oa > ob || (oa == ob && ia < ib);      // a belongs before b
                    const bool up = (i & k2) == 0;
                    if (up ? !a_first : a_first) { sord[i] = ob; sord[p2] = oa; sidx[i] = ib; sidx[p2] = ia; }
                }
            }
            __syncthreads();
        }
    const size_t out0 = (size_t)b * q.seg_off[q.L] + q.seg_off[l];
    for (int i = t; i < K; i += 1024) { keys_out[out0 + i] = ord2f(sord[i]); vals_out[out0 + i] = sidx[i]; }
}

// R2: decode the K best of every (image, level); grid (ceil(K/128), L, B)
__global__ void rpn_decode_kernel(const RpnParams q, const DetParams p, const float *keys, const int *vals,
                                  unsigned char *ws, const DetWs Lw)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, b = blockIdx.z;
    if (k >= q.K) return;
    const int per_img = q.seg_off[q.L], m_l = q.seg_off[l + 1] - q.seg_off[l];
    float *cscore = reinterpret_cast<float *>(ws + Lw.cand_score);
    const size_t c = ((size_t)b * q.K + k) * q.L + l;
    if (k >= m_l) { cscore[c] = -1.f; return; }
    const size_t it = (size_t)b * per_img + q.seg_off[l] + k;
    const int idx = vals[it];
    const int a = idx % q.A, cell = idx / q.A, y = cell / q.W[l], x = cell % q.W[l];
    const float score = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-keys[it])));                    // torch sigmoid
    const float *ba = q.base_anchors + ((size_t)l * q.A + a) * 4;
    const float sx = (float)(x * q.stride[l]), sy = (float)(y * q.stride[l]);
    const float ax1 = __fadd_rn(ba[0], sx), ay1 = __fadd_rn(ba[1], sy), ax2 = __fadd_rn(ba[2], sx), ay2 = __fadd_rn(ba[3], sy);
    const size_t hw = (size_t)q.H[l] * q.W[l];
    const float *d = q.reg[l] + ((size_t)b * 4 * q.A + 4 * a) * hw + cell;
    const float dx = __fadd_rn(__fmul_rn(d[0], p.stds[0]), p.means[0]);
    const float dy = __fadd_rn(__fmul_rn(d[hw], p.stds[1]), p.means[1]);
    float dw = __fadd_rn(__fmul_rn(d[2 * hw], p.stds[2]), p.means[2]);
    float dh = __fadd_rn(__fmul_rn(d[3 * hw], p.stds[3]), p.means[3]);
    dw = fminf(fmaxf(dw, -p.max_ratio), p.max_ratio);
    dh = fminf(fmaxf(dh, -p.max_ratio), p.max_ratio);
    const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
    const float pw = __fsub_rn(ax2, ax1), ph = __fsub_rn(ay2, ay1);
    const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
    const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
    float x1 = __fsub_rn(gx, __fmul_rn(gw, 0.5f)), y1 = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
    float x2 = __fadd_rn(gx, __fmul_rn(gw, 0.5f)), y2 = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
    if (p.img_hw != nullptr) {
        const float H = p.img_hw[2 * b], W = p.img_hw[2 * b + 1];
        x1 = fminf(fmaxf(x1, 0.f), W); x2 = fminf(fmaxf(x2, 0.f), W);
        y1 = fminf(fmaxf(y1, 0.f), H); y2 = fminf(fmaxf(y2, 0.f), H);
    }
    reinterpret_cast<float4 *>(ws + Lw.cand_box)[c] = make_float4(x1, y1, x2, y2);
    const bool valid = q.min_size < 0.f || (__fsub_rn(x2, x1) > q.min_size && __fsub_rn(y2, y1) > q.min_size);
    cscore[c] = valid ? score : -1.f;
    if (valid) atomicMax(&hdr_of(ws, Lw, b, q.L)[0], f2ord(fmaxf(fmaxf(x1, y1), fmaxf(x2, y2))));
}

// R3: per (image, level) the list of valid positions in score order; grid (L, B), 1024 threads
__global__ void rpn_compact_kernel(const RpnParams q, unsigned char *ws, const DetWs Lw)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int l = blockIdx.x, b = blockIdx.y, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const float *cscore = reinterpret_cast<const float *>(ws + Lw.cand_score);
    int *sorted = reinterpret_cast<int *>(ws + Lw.sorted) + ((size_t)b * q.L + l) * q.K;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int k0 = 0; k0 < q.K; k0 += blockDim.x) {
        const int k = k0 + t;
        const bool v = k < q.K && cscore[((size_t)b * q.K + k) * q.L + l] >= 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = carry;
        for (int w = 0; w < wid; ++w) before += warp_tot[w];
        if (v) sorted[before + within] = k;
        __syncthreads();
        if (t == 0) { int s = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_tot[w]; carry += s; }
        __syncthreads();
    }
    if (t == 0) hdr_of(ws, Lw, b, q.L)[2 + l] = (unsigned)carry;
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
    p.score_thr = score_thr; p.iou_thr = iou_thr; p.class_major_ties = 0;
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
        const size_t diag_bytes = (size_t)Rmax * sizeof(unsigned long long);
        const int smem_rows = (int)min((size_t)Rmax, (200 * 1024 - row_bytes - diag_bytes) / row_bytes);
        const size_t smem = row_bytes * (1 + (size_t)smem_rows) + diag_bytes;
        FGN_SMEM_OPTIN(det_nms_reduce_kernel, smem);
        det_nms_reduce_kernel<<<dim3(N, B), 256, smem, st>>>(p, ws, L, smem_rows);
        FGN_LAUNCH_OK();
    }
    det_merge_kernel<<<dim3(ceil_div(Rmax, 128), N, B), 128, 0, st>>>(p, ws, L, det_out, label_out, count_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}


namespace {
struct RpnWs { size_t det, keys_in, vals_in, keys_out, vals_out, seg_b, seg_e, img_off, hist, cub, total; size_t cub_bytes; };

RpnWs rpn_layout(const int *H, const int *W, int L, int A, int B, int K)
{
    RpnWs w;
    size_t per_img = 0;
    for (int l = 0; l < L; ++l) per_img += (size_t)H[l] * W[l] * A;
    const size_t items = per_img * B;
    size_t o = 0;
    w.det = o;      o = align256(o + det_layout(B * K, L, B, K).total);
    w.keys_in = o;  o = align256(o + items * 4);
    w.vals_in = o;  o = align256(o + items * 4);
    w.keys_out = o; o = align256(o + items * 4);
    w.vals_out = o; o = align256(o + items * 4);
    w.seg_b = o;    o = align256(o + (size_t)B * L * 4);
    w.seg_e = o;    o = align256(o + (size_t)B * L * 4);
    w.img_off = o;  o = align256(o + (size_t)(B + 1) * 4);
    w.hist = o;     o = align256(o + (size_t)2 * B * L * kSelBins * 4);    // high-16 and low-16 histograms per segment
    w.cub_bytes = 0;
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, w.cub_bytes, (const float *)nullptr, (float *)nullptr,
                                                       (const int *)nullptr, (int *)nullptr, (int)items, B * L,
                                                       (const int *)nullptr, (const int *)nullptr);
    w.cub = o;      o = align256(o + w.cub_bytes);
    w.total = o;
    return w;
}
}  // namespace

extern "C" size_t fgn_rpn_proposals_workspace_bytes(const int *H, const int *W, int L, int A, int B, int nms_pre)
{
    if (L < 1 || L > FGN_MAX_LEVELS || A < 1 || B < 1 || nms_pre < 1) return 256;
    int maxm = 0;
    for (int l = 0; l < L; ++l) maxm = max(maxm, H[l] * W[l] * A);
    return rpn_layout(H, W, L, A, B, min(nms_pre, maxm)).total;
}

extern "C" int fgn_rpn_proposals(const float *const *cls, const float *const *reg, const int *H, const int *W,
                                 const int *strides, int L, int A, int B, const float *base_anchors,
                                 const float *img_hw, const float *means, const float *stds, float wh_ratio_clip,
                                 int nms_pre, float iou_thr, int max_per_img, float min_bbox_size,
                                 float *prop_out, int32_t *level_out, int32_t *count_out,
                                 void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(L >= 1 && L <= FGN_MAX_LEVELS && A >= 1 && B >= 1 && nms_pre >= 1 && max_per_img >= 1,
                  "bad dims L=%d A=%d B=%d nms_pre=%d max_per_img=%d", L, A, B, nms_pre, max_per_img);
    FGN_CHECK_ARG(cls && reg && H && W && strides && base_anchors && means && stds && prop_out && level_out && count_out,
                  "NULL pointer");
    FGN_CHECK_ARG(wh_ratio_clip > 0.f, "wh_ratio_clip=%f", wh_ratio_clip);
    cudaStream_t st = (cudaStream_t)stream;
    RpnParams q;
    q.L = L; q.A = A; q.B = B; q.base_anchors = base_anchors; q.min_size = min_bbox_size;
    int maxm = 0, off = 0;
    for (int l = 0; l < FGN_MAX_LEVELS; ++l) {
        q.cls[l] = l < L ? cls[l] : nullptr; q.reg[l] = l < L ? reg[l] : nullptr;
        q.H[l] = l < L ? H[l] : 0; q.W[l] = l < L ? W[l] : 0; q.stride[l] = l < L ? strides[l] : 0;
        q.seg_off[l] = off;
        if (l < L) {
            FGN_CHECK_ARG(cls[l] && reg[l] && H[l] > 0 && W[l] > 0, "level %d", l);
            off += H[l] * W[l] * A; maxm = max(maxm, H[l] * W[l] * A);
        }
    }
    for (int l = L; l <= FGN_MAX_LEVELS; ++l) q.seg_off[l] = off;
    const int K = min(nms_pre, maxm);
    q.K = K;
    const RpnWs Wl = rpn_layout(H, W, L, A, B, K);
    FGN_CHECK_ARG(workspace && workspace_bytes >= Wl.total, "workspace %zu < %zu bytes", workspace_bytes, Wl.total);
    unsigned char *base = static_cast<unsigned char *>(workspace);
    unsigned char *ws = base + Wl.det;
    const DetWs Ld = det_layout(B * K, L, B, K);
    float *keys_in = (float *)(base + Wl.keys_in), *keys_out = (float *)(base + Wl.keys_out);
    int *vals_in = (int *)(base + Wl.vals_in), *vals_out = (int *)(base + Wl.vals_out);
    int *seg_b = (int *)(base + Wl.seg_b), *seg_e = (int *)(base + Wl.seg_e), *img_off = (int *)(base + Wl.img_off);
    const size_t items = (size_t)off * B;

    DetParams p;
    p.rois = nullptr; p.cls = nullptr; p.reg = nullptr; p.img_off = img_off; p.img_hw = img_hw; p.scale = nullptr;
    p.R = B * K; p.N = L; p.B = B; p.Rmax = K; p.max_per_img = max_per_img; p.class_major_ties = 1;
    for (int i = 0; i < 4; ++i) { p.means[i] = means[i]; p.stds[i] = stds[i]; }
    p.score_thr = 0.f; p.iou_thr = iou_thr; p.max_ratio = fabsf(logf(wh_ratio_clip));

    FGN_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(int32_t) * B, st));
    FGN_CUDA_OK(cudaMemsetAsync(ws + Ld.hdr, 0, (size_t)B * (2 + 2 * L) * sizeof(unsigned int), st));
    if (K <= kSelCap) {
        // hand-written segmented top-K: two histograms + one select-and-sort block per (image, level)
        unsigned int *hist_hi = (unsigned int *)(base + Wl.hist), *hist_lo = hist_hi + (size_t)B * L * kSelBins;
        FGN_CUDA_OK(cudaMemsetAsync(hist_hi, 0, (size_t)2 * B * L * kSelBins * 4, st));
        rpn_keys_kernel<<<1, 256, 0, st>>>(q, nullptr, nullptr, seg_b, seg_e, img_off);     // only the small index tables
        FGN_LAUNCH_OK();
        const int hb = max(1, min(64, ceil_div(maxm, 4096)));
        rpn_hist_hi_kernel<<<dim3(hb, L, B), 256, 0, st>>>(q, hist_hi);
        FGN_LAUNCH_OK();
        rpn_hist_lo_kernel<<<dim3(hb, L, B), 1024, 0, st>>>(q, hist_hi, hist_lo);
        FGN_LAUNCH_OK();
        FGN_SMEM_OPTIN(rpn_select_sort_kernel, kSelCap * 8);
        rpn_select_sort_kernel<<<dim3(L, B), 1024, kSelCap * 8, st>>>(q, hist_hi, hist_lo, keys_out, vals_out);
        FGN_LAUNCH_OK();
    } else {
        // more candidates than one block sorts in shared memory (nms_pre > 8192): library segmented sort
        rpn_keys_kernel<<<(int)min((size_t)1184, (items + 255) / 256), 256, 0, st>>>(q, keys_in, vals_in, seg_b, seg_e, img_off);
        FGN_LAUNCH_OK();
        size_t cub_bytes = Wl.cub_bytes;
        FGN_CUDA_OK(cub::DeviceSegmentedRadixSort::SortPairsDescending(base + Wl.cub, cub_bytes, keys_in, keys_out, vals_in, vals_out,
                                                                       (int)items, B * L, seg_b, seg_e, 0, 32, st));
        count_launch(2);                                         // (library sort passes, not counted exactly)
    }
    rpn_decode_kernel<<<dim3(ceil_div(K, 128), L, B), 128, 0, st>>>(q, p, keys_out, vals_out, ws, Ld);
    FGN_LAUNCH_OK();
    rpn_compact_kernel<<<dim3(L, B), 1024, 0, st>>>(q, ws, Ld);
    FGN_LAUNCH_OK();
    det_nms_mask_kernel<<<dim3(Ld.words, Ld.words, B * L), 64, 0, st>>>(p, ws, Ld);
    FGN_LAUNCH_OK();
    {
        const size_t row_bytes = (size_t)Ld.words * sizeof(unsigned long long);
        const size_t diag_bytes = (size_t)K * sizeof(unsigned long long);
        const int smem_rows = (int)min((size_t)K, (200 * 1024 - row_bytes - diag_bytes) / row_bytes);
        const size_t smem = row_bytes * (1 + (size_t)smem_rows) + diag_bytes;
        FGN_SMEM_OPTIN(det_nms_reduce_kernel, smem);
        det_nms_reduce_kernel<<<dim3(L, B), 256, smem, st>>>(p, ws, Ld, smem_rows);
        FGN_LAUNCH_OK();
    }
    det_merge_kernel<<<dim3(ceil_div(K, 128), L, B), 128, 0, st>>>(p, ws, Ld, prop_out, level_out, count_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

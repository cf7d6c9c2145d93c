// mask_paste.cu -- test-time mask pasting fused with COCO run-length encoding (sm_100a).
//
// Reference chain (second part): fgn_roi_head.py:668-671 -> FCNMaskHead.get_seg_masks
// [3P, mmdet 2.18]: sigmoid, _do_paste_mask = F.grid_sample(bilinear, zeros, align_corners=False) of every
// [M,M] mask onto the whole image, ">= mask_thr_binary"; then fgn.py:281 -> mmdet.core.encode_mask_results ->
// pycocotools mask.encode [3P]: run lengths of the column-major scan and their compressed string.
//
// The reference materialises a [D, img_h, img_w] bool tensor on the device (107 MB for 100 detections at
// 800x1333), copies it to the host and encodes it there.  Here the walk kernel visits only the part of the image a
// box can touch, in column-major order and in pieces of 512 32-row words (as many CTAs as the box needs), and finds
// the run boundaries with bit tricks and a block scan; the encode kernel (one CTA per detection) puts the pieces in
// order, turns run starts into run lengths in place and writes the compressed string.  The dense mask never exists
// and the device->host copy is the RLE itself (a few KB per detection).  fgn_mask_paste writes the dense masks
// for callers that want get_seg_masks' own return value.
//
// Arithmetic: every operation of the coordinate chain is an explicitly rounded fp32 op in the order of
// _do_paste_mask and of torch's CUDA grid sampler (oracle/fgn_oracle.py::paste_values restates the same chain),
// so a pixel can differ from the reference only where sigmoid/expf rounding moves a value across the threshold.
#include "common.cuh"
#include <algorithm>

namespace fgn {

namespace {

constexpr int kPasteThreads = 512;
constexpr int kPasteIters = 1;                         // scan steps (of kPasteThreads 32-row words) per CTA of the walk kernel:
                                                       // one step is ~18 K issue cycles of an SM, coarser pieces leave SMs idle
constexpr int kPasteMaxChunks = 4096;                  // (CTA, step) chunks per detection the encode kernel can order

struct PasteBox {
    float x0, y0, x1, y1;
    int   H, W;
};

// _do_paste_mask: ((p + 0.5) - b0) / (b1 - b0) * 2 - 1, inf -> 0
__device__ __forceinline__ float paste_coord(int p, float b0, float b1)
{
    float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)p, 0.5f), b0), __fsub_rn(b1, b0)), 2.f), 1.f);
    if (isinf(g)) g = 0.f;
    return g;
}

// grid sampler, align_corners=False: ((g + 1) * M - 1) / 2
__device__ __forceinline__ float paste_unnormalize(float g, int M)
{
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)M), 1.f), 0.5f);   // (/ 2 is exact either way)
}

// bilinear sample with zero padding of the [M,M] probability map in shared memory
__device__ __forceinline__ float paste_sample(const float *sm, int M, float ix, float iy)
{
    if (!(ix > -1.f && ix < (float)M && iy > -1.f && iy < (float)M)) return 0.f;   // all four corners outside (or NaN)
    const float xf = floorf(ix), yf = floorf(iy);
    const float wx1 = __fsub_rn(ix, xf), wx0 = __fsub_rn(__fadd_rn(xf, 1.f), ix);
    const float wy1 = __fsub_rn(iy, yf), wy0 = __fsub_rn(__fadd_rn(yf, 1.f), iy);
    const int xa = (int)xf, ya = (int)yf, xb = xa + 1, yb = ya + 1;
    const bool xa_ok = xa >= 0, xb_ok = xb < M, ya_ok = ya >= 0, yb_ok = yb < M;
    const float nw = (xa_ok && ya_ok) ? sm[ya * M + xa] : 0.f;
    const float ne = (xb_ok && ya_ok) ? sm[ya * M + xb] : 0.f;
    const float sw = (xa_ok && yb_ok) ? sm[yb * M + xa] : 0.f;
    const float se = (xb_ok && yb_ok) ? sm[yb * M + xb] : 0.f;
    float v = __fmul_rn(nw, __fmul_rn(wy0, wx0));
    v = __fadd_rn(v, __fmul_rn(ne, __fmul_rn(wy0, wx1)));
    v = __fadd_rn(v, __fmul_rn(sw, __fmul_rn(wy1, wx0)));
    v = __fadd_rn(v, __fmul_rn(se, __fmul_rn(wy1, wx1)));
    return v;
}

__device__ __forceinline__ PasteBox load_box(const float *boxes, int box_stride, const int32_t *det_img,
                                             const int32_t *img_hw, int d)
{
    PasteBox b;
    const float *p = boxes + (size_t)d * box_stride;
    b.x0 = p[0]; b.y0 = p[1]; b.x1 = p[2]; b.y1 = p[3];
    const int im = det_img != nullptr ? det_img[d] : 0;
    b.H = img_hw[2 * im]; b.W = img_hw[2 * im + 1];
    return b;
}

__device__ __forceinline__ void load_probabilities(float *sm, const float *mask_pred, int d, int M)
{
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
        const float x = mask_pred[(size_t)d * M * M + i];
        sm[i] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
    }
}

// Pixels a box can influence along one axis: the sampler returns 0 unless -1 < i < M, i.e. within half a mask
// cell of the box.  The range is widened by a whole cell plus two pixels, so rounding never matters; a
// degenerate or non-finite extent (every pixel samples the same mask coordinate) takes the whole axis.
__device__ __forceinline__ void paste_range(float b0, float b1, int M, int size, int &lo, int &hi)
{
    const float ext = fabsf(b1 - b0);
    lo = 0; hi = size;
    if (!(ext > 0.f) || !isfinite(ext) || !isfinite(b0) || !isfinite(b1)) return;
    const float m = ext / (float)M + 2.f;
    const float a = fminf(b0, b1) - m, c = fmaxf(b0, b1) + m;
    if (a > 0.f) lo = a >= (float)size ? size : (int)floorf(a);
    if (c < (float)size) hi = c <= 0.f ? 0 : (int)ceilf(c);
    if (hi < lo) hi = lo;
}

// exclusive block scan of one int per thread (kPasteThreads threads); returns the block total in `total`
__device__ __forceinline__ int block_scan_excl(int v, int *warp_sums, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kPasteThreads / 32; ++w) {
        const int s = warp_sums[w];
        if (w < warp) base += s;
        total += s;
    }
    __syncthreads();                                   // warp_sums may be reused by the next call
    return base + incl - v;
}

// characters rleToString spends on one value (5 bits each, sign-extended stop rule)
__device__ __forceinline__ int rle_chars(long long x, unsigned char *dst)
{
    int n = 0;
    bool more = true;
    while (more) {
        int c = (int)(x & 0x1f);
        x >>= 5;
        more = (c & 0x10) ? (x != -1) : (x != 0);
        if (more) c |= 0x20;
        if (dst != nullptr) dst[n] = (unsigned char)(c + 48);
        ++n;
    }
    return n;
}

// Workspace per call: cursor[D] | chunk_off[D][G] | chunk_cnt[D][G] | starts[D][cap]   (G = segs * kPasteIters)
// Walk kernel, grid (segs, D): CTA (s, d) walks kPasteIters * kPasteThreads words of detection d's region and appends
// the run starts of every scan step as one chunk to starts[d] (atomic cursor; the chunk table keeps the order).
__global__ void __launch_bounds__(kPasteThreads)
mask_paste_walk_kernel(const float *__restrict__ mask_pred, const float *__restrict__ boxes, const int box_stride,
                       const int32_t *__restrict__ det_img, const int32_t *__restrict__ img_hw, const int M,
                       const float thr, int32_t *__restrict__ cursor, int32_t *__restrict__ chunk_off,
                       int32_t *__restrict__ chunk_cnt, int32_t *__restrict__ starts_ws, const int cap)
{
    extern __shared__ float sm[];                      // [M*M] probabilities
    __shared__ int warp_sums[kPasteThreads / 32];
    __shared__ int chunk_base;
    const int d = blockIdx.y, seg = blockIdx.x, G = gridDim.x * kPasteIters, tid = threadIdx.x;
    const PasteBox b = load_box(boxes, box_stride, det_img, img_hw, d);
    const int H = b.H, W = b.W;
    int32_t *cnt = starts_ws + (size_t)d * cap;
    if (H <= 0 || W <= 0) return;

    const bool zero_bit = 0.f >= thr;                  // what a pixel no mask cell reaches compares to
    int cx0, cx1, ry0, ry1;
    paste_range(b.x0, b.x1, M, W, cx0, cx1);
    paste_range(b.y0, b.y1, M, H, ry0, ry1);
    if (zero_bit) { cx0 = 0; cx1 = W; ry0 = 0; ry1 = H; }
    // x half of the sample (shared by every pixel of a column) and the pixel test
    struct ColX { float wx0, wx1; int xa, xb; bool any; };
    auto column = [&](int x) {
        ColX c;
        c.any = false; c.wx0 = c.wx1 = 0.f; c.xa = c.xb = 0;
        if (x < cx0 || x >= cx1) return c;
        const float ix = paste_unnormalize(paste_coord(x, b.x0, b.x1), M);
        if (!(ix > -1.f && ix < (float)M)) return c;
        const float xf = floorf(ix);
        c.wx1 = __fsub_rn(ix, xf); c.wx0 = __fsub_rn(__fadd_rn(xf, 1.f), ix);
        c.xa = (int)xf; c.xb = c.xa + 1; c.any = true;
        return c;
    };
    auto bit = [&](const ColX &c, int y) -> bool {                     // same op order as paste_sample
        if (y < ry0 || y >= ry1) return zero_bit;
        float v = 0.f;
        const float iy = paste_unnormalize(paste_coord(y, b.y0, b.y1), M);
        if (c.any && iy > -1.f && iy < (float)M) {
            const float yf = floorf(iy);
            const float wy1 = __fsub_rn(iy, yf), wy0 = __fsub_rn(__fadd_rn(yf, 1.f), iy);
            const int ya = (int)yf, yb = ya + 1;
            const bool xa_ok = c.xa >= 0, xb_ok = c.xb < M, ya_ok = ya >= 0, yb_ok = yb < M;
            const float nw = (xa_ok && ya_ok) ? sm[ya * M + c.xa] : 0.f;
            const float ne = (xb_ok && ya_ok) ? sm[ya * M + c.xb] : 0.f;
            const float sw = (xa_ok && yb_ok) ? sm[yb * M + c.xa] : 0.f;
            const float se = (xb_ok && yb_ok) ? sm[yb * M + c.xb] : 0.f;
            v = __fmul_rn(nw, __fmul_rn(wy0, c.wx0));
            v = __fadd_rn(v, __fmul_rn(ne, __fmul_rn(wy0, c.wx1)));
            v = __fadd_rn(v, __fmul_rn(sw, __fmul_rn(wy1, c.wx0)));
            v = __fadd_rn(v, __fmul_rn(se, __fmul_rn(wy1, c.wx1)));
        }
        return v >= thr;
    };

    // A thread takes 32 consecutive rows of one column (a "word") and marks where a run starts:
    //   start[k] = bit[k] != bit[k-1].  The rows walked per column are the region's [ry0, ry1) plus the row just
    // below (it reads zero_bit, so a run that reaches the region's bottom edge ends there); the bit before a
    // column's first walked row is zero_bit, or -- when the region starts at row 0 -- the bottom pixel of the
    // column to the left.  When the region reaches the bottom row but not the top one, the top pixel of the next
    // column (outside the region) can start a run too: the column's first word carries that extra position.
    const int xe = min(cx1, W - 1);                    // last examined column (cx1 itself: runs ending at its top)
    const int ncx = xe - cx0 + 1;
    const int re = min(ry1 + 1, H);                    // walked rows [ry0, re)
    const int nwc = max(1, (re - ry0 + 31) >> 5);      // words per column
    const bool top_extra = ry0 > 0 && ry1 == H;
    const long long U = (long long)ncx * nwc;
    const long long u_begin = (long long)seg * (kPasteIters * kPasteThreads);
    if (u_begin >= U) return;                          // (before the probabilities are needed: most CTAs of a small box)
    load_probabilities(sm, mask_pred, d, M);
    __syncthreads();
    for (int it = 0; it < kPasteIters; ++it) {
        const long long u0 = u_begin + (long long)it * kPasteThreads;
        if (u0 >= U) break;
        const long long u = u0 + tid;
        unsigned starts = 0u;
        bool pre = false;
        int x = 0, yw = 0;
        if (u < U) {
            const int xi = (int)(u / nwc), wi = (int)(u - (long long)xi * nwc);
            x = cx0 + xi; yw = ry0 + 32 * wi;
            const ColX c = column(x);
            bool prev;
            if (wi > 0) prev = bit(c, yw - 1);
            else {
                const bool left_bottom = x > 0 ? bit(column(x - 1), H - 1) : false;
                prev = ry0 > 0 ? zero_bit : left_bottom;
                pre = top_extra && x > 0 && left_bottom != zero_bit;     // run starting at (x, 0)
            }
            unsigned bits = 0u;
            const int n = min(32, re - yw);
            for (int k = 0; k < n; ++k) bits |= (bit(c, yw + k) ? 1u : 0u) << k;
            starts = bits ^ ((bits << 1) | (prev ? 1u : 0u));
            if (n < 32) starts &= (1u << max(n, 0)) - 1u;
        }
        int total;
        int pos = block_scan_excl((pre ? 1 : 0) + __popc(starts), warp_sums, total);
        if (tid == 0) {
            const int at = total > 0 ? atomicAdd(&cursor[d], total) : 0;
            chunk_base = at;
            chunk_off[(size_t)d * G + seg * kPasteIters + it] = at;
            chunk_cnt[(size_t)d * G + seg * kPasteIters + it] = total;
        }
        __syncthreads();
        pos += chunk_base;
        if (pre) { if (pos < cap - 1) cnt[pos] = (int32_t)((long long)x * H); ++pos; }
        while (starts != 0u) {
            const int k = __ffs(starts) - 1;
            starts &= starts - 1u;
            if (pos < cap - 1) cnt[pos] = (int32_t)((long long)x * H + yw + k);      // run starts, for now
            ++pos;
        }
        __syncthreads();                               // chunk_base is rewritten by the next step
    }
}

// The same walk over GIVEN masks (encode_mask_results of the ground-truth qry_isegmaps, fgn.py:296-298): masks [D,H,W]
// bytes (0 / non-zero), every pixel is walked, same chunk protocol, same encode kernel.
__global__ void __launch_bounds__(kPasteThreads)
mask_rle_walk_dense_kernel(const unsigned char *__restrict__ masks, const int H, const int W,
                           int32_t *__restrict__ cursor, int32_t *__restrict__ chunk_off,
                           int32_t *__restrict__ chunk_cnt, int32_t *__restrict__ starts_ws, const int cap)
{
    __shared__ int warp_sums[kPasteThreads / 32];
    __shared__ int chunk_base;
    const int d = blockIdx.y, seg = blockIdx.x, G = gridDim.x * kPasteIters, tid = threadIdx.x;
    const unsigned char *m = masks + (size_t)d * H * W;
    int32_t *cnt = starts_ws + (size_t)d * cap;
    const int nwc = (H + 31) >> 5;                     // words per column
    const long long U = (long long)W * nwc;
    const long long u_begin = (long long)seg * (kPasteIters * kPasteThreads);
    if (u_begin >= U) return;
    for (int it = 0; it < kPasteIters; ++it) {
        const long long u0 = u_begin + (long long)it * kPasteThreads;
        if (u0 >= U) break;
        const long long u = u0 + tid;
        unsigned starts = 0u;
        int x = 0, yw = 0;
        if (u < U) {
            x = (int)(u / nwc);
            yw = 32 * (int)(u - (long long)x * nwc);
            const bool prev = yw > 0 ? m[(size_t)(yw - 1) * W + x] != 0 : (x > 0 ? m[(size_t)(H - 1) * W + x - 1] != 0 : false);
            unsigned bits = 0u;
            const int n = min(32, H - yw);
            for (int k = 0; k < n; ++k) bits |= (m[(size_t)(yw + k) * W + x] != 0 ? 1u : 0u) << k;
            starts = bits ^ ((bits << 1) | (prev ? 1u : 0u));
            if (n < 32) starts &= (1u << max(n, 0)) - 1u;
        }
        int total;
        int pos = block_scan_excl(__popc(starts), warp_sums, total);
        if (tid == 0) {
            const int at = total > 0 ? atomicAdd(&cursor[d], total) : 0;
            chunk_base = at;
            chunk_off[(size_t)d * G + seg * kPasteIters + it] = at;
            chunk_cnt[(size_t)d * G + seg * kPasteIters + it] = total;
        }
        __syncthreads();
        pos += chunk_base;
        while (starts != 0u) {
            const int k = __ffs(starts) - 1;
            starts &= starts - 1u;
            if (pos < cap - 1) cnt[pos] = (int32_t)((long long)x * H + yw + k);
            ++pos;
        }
        __syncthreads();
    }
}

// Encode kernel, grid D: orders the chunks of a detection, turns run starts into run lengths and writes the string.
__global__ void __launch_bounds__(kPasteThreads)
mask_paste_encode_kernel(const int32_t *__restrict__ det_img, const int32_t *__restrict__ img_hw,
                         const int32_t *__restrict__ chunk_off, const int32_t *__restrict__ chunk_cnt,
                         const int32_t *__restrict__ starts_ws, const int G, int32_t *__restrict__ counts_out,
                         int32_t *__restrict__ ncounts_out, unsigned char *__restrict__ str_out,
                         int32_t *__restrict__ strlen_out, const int cap, const int cap_bytes,
                         const int fixed_h, const int fixed_w)
{
    __shared__ int warp_sums[kPasteThreads / 32];
    __shared__ int prefix[kPasteMaxChunks];            // exclusive prefix of the chunk sizes, chunk order
    const int d = blockIdx.x, tid = threadIdx.x;
    const int im = det_img != nullptr ? det_img[d] : 0;
    const int H = img_hw != nullptr ? img_hw[2 * im] : fixed_h, W = img_hw != nullptr ? img_hw[2 * im + 1] : fixed_w;
    int32_t *cnt = counts_out + (size_t)d * cap;
    if (H <= 0 || W <= 0) {                            // empty image: pycocotools emits no run
        if (tid == 0) { ncounts_out[d] = 0; if (strlen_out != nullptr) strlen_out[d] = 0; }
        return;
    }
    int base = 0;
    for (int g0 = 0; g0 < G; g0 += kPasteThreads) {
        const int g = g0 + tid;
        const int n = g < G ? chunk_cnt[(size_t)d * G + g] : 0;
        int total;
        const int ex = base + block_scan_excl(n, warp_sums, total);
        if (g < G) prefix[g] = ex;
        base += total;
    }
    const int ntr = base, m = ntr + 1;                 // runs = boundaries + 1 (the first run may be empty: t = 0)
    if (m > cap) {
        if (tid == 0) { ncounts_out[d] = -m; if (strlen_out != nullptr) strlen_out[d] = 0; }
        return;
    }
    __syncthreads();
    // run starts in image order: entry i lives in the last chunk whose prefix is <= i
    const int32_t *src = starts_ws + (size_t)d * cap;
    for (int i = tid; i < ntr; i += kPasteThreads) {
        int lo = 0, hi = G - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (prefix[mid] <= i) lo = mid; else hi = mid - 1;
        }
        cnt[i] = src[chunk_off[(size_t)d * G + lo] + (i - prefix[lo])];
    }
    __syncthreads();
    // run starts -> run lengths, in place, from the back (a chunk only reads entries no later chunk has rewritten)
    const long long HW = (long long)H * W;
    for (int hi = m; hi > 0; hi -= kPasteThreads) {
        const int i = hi - 1 - tid;
        int32_t v = 0;
        if (i >= 0) {
            const long long end = i < ntr ? (long long)cnt[i] : HW;
            const long long start = i > 0 ? (long long)cnt[i - 1] : 0;
            v = (int32_t)(end - start);
        }
        __syncthreads();
        if (i >= 0) cnt[i] = v;
        __syncthreads();
    }
    if (tid == 0) ncounts_out[d] = m;
    if (str_out == nullptr) return;

    // rleToString: value i > 2 is stored as the difference to value i - 2
    unsigned char *str = str_out + (size_t)d * cap_bytes;
    int sbase = 0;
    bool overflow = false;
    for (int i0 = 0; i0 < m; i0 += kPasteThreads) {
        const int i = i0 + tid;
        long long x = 0;
        int n = 0;
        if (i < m) {
            x = cnt[i];
            if (i > 2) x -= cnt[i - 2];
            n = rle_chars(x, nullptr);
        }
        int total;
        const int off = sbase + block_scan_excl(n, warp_sums, total);
        if (i < m) {
            if (off + n <= cap_bytes) rle_chars(x, str + off);
            else overflow = true;
        }
        sbase += total;
    }
    (void)overflow;
    if (tid == 0) strlen_out[d] = sbase <= cap_bytes ? sbase : -sbase;
}

// dense [D, H, W] (0/1 bytes), the tensor get_seg_masks returns
__global__ void __launch_bounds__(256)
mask_paste_dense_kernel(const float *__restrict__ mask_pred, const float *__restrict__ boxes, const int box_stride,
                        const int M, const int H, const int W, const float thr, unsigned char *__restrict__ out)
{
    extern __shared__ float sm[];
    const int d = blockIdx.y;
    load_probabilities(sm, mask_pred, d, M);
    __syncthreads();
    const float *p = boxes + (size_t)d * box_stride;
    const float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
    const size_t HW = (size_t)H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const float ix = paste_unnormalize(paste_coord(x, x0, x1), M);
        const float iy = paste_unnormalize(paste_coord(y, y0, y1), M);
        out[(size_t)d * HW + i] = paste_sample(sm, M, ix, iy) >= thr ? 1 : 0;
    }
}

}  // namespace

}  // namespace fgn

using namespace fgn;

static int paste_segments(int Hmax, int Wmax)
{
    const long long words = (long long)Wmax * ((Hmax + 1 + 31) / 32 + 1);
    const long long per = (long long)kPasteIters * kPasteThreads;
    return (int)std::max<long long>(1, (words + per - 1) / per);
}

extern "C" size_t fgn_mask_paste_rle_workspace_bytes(int D, int cap, int Hmax, int Wmax)
{
    if (D <= 0 || cap <= 0 || Hmax <= 0 || Wmax <= 0) return 0;
    const size_t G = (size_t)paste_segments(Hmax, Wmax) * kPasteIters;
    return ((size_t)D * (1 + 2 * G) + (size_t)D * cap) * sizeof(int32_t);
}

extern "C" int fgn_mask_paste_rle(const float *mask_pred, const float *boxes, int box_stride,
                                  const int32_t *det_img, const int32_t *img_hw, int D, int M, float mask_thr,
                                  int Hmax, int Wmax, int32_t *counts_out, int32_t *ncounts_out,
                                  unsigned char *str_out, int32_t *strlen_out, int cap, int cap_bytes,
                                  void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(D >= 0 && M >= 1 && M <= 112 && box_stride >= 4, "bad dims D=%d M=%d box_stride=%d", D, M, box_stride);
    if (D == 0) return FGN_OK;
    FGN_CHECK_ARG(mask_pred && boxes && img_hw && counts_out && ncounts_out, "NULL pointer");
    FGN_CHECK_ARG(cap >= 2 && Hmax >= 1 && Wmax >= 1 && D <= 65535, "cap=%d Hmax=%d Wmax=%d D=%d", cap, Hmax, Wmax, D);
    FGN_CHECK_ARG(str_out == nullptr || (strlen_out != nullptr && cap_bytes >= 1), "str_out needs strlen_out and cap_bytes");
    const int segs = paste_segments(Hmax, Wmax), G = segs * kPasteIters;
    FGN_CHECK_ARG(G <= kPasteMaxChunks, "image %dx%d needs %d chunks per detection (max %d)", Hmax, Wmax, G, kPasteMaxChunks);
    const size_t need = fgn_mask_paste_rle_workspace_bytes(D, cap, Hmax, Wmax);
    FGN_CHECK_ARG(workspace && workspace_bytes >= need, "workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *cursor = static_cast<int32_t *>(workspace);
    int32_t *chunk_off = cursor + D, *chunk_cnt = chunk_off + (size_t)D * G, *starts = chunk_cnt + (size_t)D * G;
    FGN_CUDA_OK(cudaMemsetAsync(cursor, 0, (size_t)D * (1 + 2 * (size_t)G) * sizeof(int32_t), st));
    const size_t smem = (size_t)M * M * sizeof(float);
    FGN_SMEM_OPTIN(mask_paste_walk_kernel, smem);
    mask_paste_walk_kernel<<<dim3(segs, D), kPasteThreads, smem, st>>>(mask_pred, boxes, box_stride, det_img, img_hw, M,
                                                                      mask_thr, cursor, chunk_off, chunk_cnt, starts, cap);
    FGN_LAUNCH_OK();
    mask_paste_encode_kernel<<<D, kPasteThreads, 0, st>>>(det_img, img_hw, chunk_off, chunk_cnt, starts, G, counts_out,
                                                          ncounts_out, str_out, strlen_out, cap, cap_bytes, 0, 0);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_mask_rle_encode(const unsigned char *masks, int D, int H, int W, int32_t *counts_out,
                                   int32_t *ncounts_out, unsigned char *str_out, int32_t *strlen_out, int cap,
                                   int cap_bytes, void *workspace, size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(D >= 0 && H >= 0 && W >= 0, "bad dims D=%d H=%d W=%d", D, H, W);
    if (D == 0) return FGN_OK;
    FGN_CHECK_ARG(masks && counts_out && ncounts_out, "NULL pointer");
    FGN_CHECK_ARG(cap >= 2 && D <= 65535, "cap=%d D=%d", cap, D);
    FGN_CHECK_ARG(str_out == nullptr || (strlen_out != nullptr && cap_bytes >= 1), "str_out needs strlen_out and cap_bytes");
    const int segs = paste_segments(max(H, 1), max(W, 1)), G = segs * kPasteIters;
    FGN_CHECK_ARG(G <= kPasteMaxChunks, "image %dx%d needs %d chunks per mask (max %d)", H, W, G, kPasteMaxChunks);
    const size_t need = fgn_mask_paste_rle_workspace_bytes(D, cap, max(H, 1), max(W, 1));
    FGN_CHECK_ARG(workspace && workspace_bytes >= need, "workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *cursor = static_cast<int32_t *>(workspace);
    int32_t *chunk_off = cursor + D, *chunk_cnt = chunk_off + (size_t)D * G, *starts = chunk_cnt + (size_t)D * G;
    FGN_CUDA_OK(cudaMemsetAsync(cursor, 0, (size_t)D * (1 + 2 * (size_t)G) * sizeof(int32_t), st));
    if (H > 0 && W > 0) {
        mask_rle_walk_dense_kernel<<<dim3(segs, D), kPasteThreads, 0, st>>>(masks, H, W, cursor, chunk_off, chunk_cnt, starts, cap);
        FGN_LAUNCH_OK();
    }
    mask_paste_encode_kernel<<<D, kPasteThreads, 0, st>>>(nullptr, nullptr, chunk_off, chunk_cnt, starts, G, counts_out,
                                                          ncounts_out, str_out, strlen_out, cap, cap_bytes, H, W);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" int fgn_mask_paste(const float *mask_pred, const float *boxes, int box_stride, int D, int M, int img_h,
                              int img_w, float mask_thr, unsigned char *out, void *stream)
{
    FGN_CHECK_ARG(D >= 0 && M >= 1 && M <= 112 && box_stride >= 4 && img_h >= 0 && img_w >= 0,
                  "bad dims D=%d M=%d box_stride=%d img %dx%d", D, M, box_stride, img_h, img_w);
    if (D == 0 || img_h == 0 || img_w == 0) return FGN_OK;
    FGN_CHECK_ARG(mask_pred && boxes && out, "NULL pointer");
    FGN_CHECK_ARG(D <= 65535, "D=%d > 65535", D);
    const size_t smem = (size_t)M * M * sizeof(float);
    FGN_SMEM_OPTIN(mask_paste_dense_kernel, smem);
    const size_t HW = (size_t)img_h * img_w;
    const int gx = (int)min((HW + 255) / 256, (size_t)2048);
    mask_paste_dense_kernel<<<dim3(gx, D), 256, smem, (cudaStream_t)stream>>>(mask_pred, boxes, box_stride, M, img_h, img_w,
                                                                             mask_thr, out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

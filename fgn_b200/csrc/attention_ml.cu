// attention_ml.cu -- AG-RPN attention (fgn_ag_rpn_head.py:37-46) for a whole pyramid in two
// launches (class vectors; multiply) instead of three per level: the coarse levels (P5, P6) are a few KB and purely
// launch-latency bound when done one by one.  NHWC (channels_last) only; other layouts go through
// the per-level entry points.
#include "common.cuh"
#include <stdlib.h>

namespace fgn {

constexpr int kMlSlab = 64;       // pixels per partial-sum block
constexpr int kMlChunk = 4;       // float4 per thread per block in the multiply kernel

struct MlLevels {
    int          L;
    const float *in[FGN_MAX_LEVELS];
    float       *out[FGN_MAX_LEVELS];
    int          HW[FGN_MAX_LEVELS];
    int          blk_off[FGN_MAX_LEVELS + 1];   // block ranges per level (grid.x)
    int          stream_loads;                  // multiply kernel: evict-first loads of the query maps (FGN_ATT_LDCS)
};

__device__ __forceinline__ int ml_level_of(const MlLevels &lv, int blk)
{
    int l = 0;
#pragma unroll
    for (int i = 1; i < FGN_MAX_LEVELS; ++i) l += (i < lv.L && blk >= lv.blk_off[i]) ? 1 : 0;
    return l;
}

// The last slab block of a (level, class) to arrive adds that pair's partial sums in slab order (the order of the old
// finalize launch: bitwise the same vectors) -- one launch instead of two; the counters are zeroed by a memset node.
__device__ __forceinline__ void ml_finish(const MlLevels &lv, const int l, const int bn, const int BN, const int K, const int C,
                                          const float *partial, float *__restrict__ vec, unsigned int *counters)
{
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&counters[l * BN + bn], 1u) == (unsigned)(lv.blk_off[l + 1] - lv.blk_off[l] - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float inv = 1.0f / (float)(K * lv.HW[l]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int b0 = lv.blk_off[l]; b0 < lv.blk_off[l + 1]; b0 += 8) {   // eight loads in flight, added in slab order
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = b0 + j < lv.blk_off[l + 1] ? __ldcg(partial + ((size_t)(b0 + j) * BN + bn) * C + c) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (b0 + j < lv.blk_off[l + 1]) s += v[j];
        }
        vec[((size_t)l * BN + bn) * C + c] = s * inv;
    }
}

// partial[(blk_off[l] + slab) * BN + bn][C] = sum over the slab's pixels (K consecutive images = one run)
__global__ void __launch_bounds__(256)
attention_vec_ml_partial_kernel(const MlLevels lv, const int BN, const int K, const int C,
                                float *partial, float *__restrict__ vec, unsigned int *counters)
{
    const int bn = blockIdx.y;
    const int l = ml_level_of(lv, blockIdx.x), slab = blockIdx.x - lv.blk_off[l];
    const int c4 = C >> 2;
    const int total = K * lv.HW[l];
    const int p0 = slab * kMlSlab, p1 = min(total, p0 + kMlSlab);
    extern __shared__ __align__(16) float red[];   // [rows][C]
    const int rows = max(1, (int)blockDim.x / c4);
    const float *base = lv.in[l] + (size_t)bn * total * C;
    const int cv = threadIdx.x % c4, rowi = threadIdx.x / c4;
    if (rowi < rows) {
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
        partial[((size_t)blockIdx.x * BN + bn) * C + c] = s;
    }
    ml_finish(lv, l, bn, BN, K, C, partial, vec, counters);
}

// out_l[b*N+n, :, :, c] = qry_l[b, :, :, c] * vec[l][b*N+n][c]; one read, N streaming writes.
// CH float4 per thread; MINB resident CTAs per SM asked of ptxas.  <4,5> (48 registers) is the production instantiation.
// The lean ones (<4,6>: 40 registers, <2,8>: 32) and the persistent form (FGN_ATT_GRID CTAs per SM, grid-stride) were built to
// let this HBM-bound kernel run UNDER the L2- / tensor-bound persistent kernels of other episodes -- a 256-thread CTA of 32
// registers fits in what two RoIAlign CTAs or the contraction's CTA leave free.  Measured (tools/overlap_probe.py,
// profiles/r02_overlap_probe.jsonl): it does not happen -- RoIAlign + attention on two streams take 81 us against 100 us back
// to back whatever the form (a one-CTA-per-SM persistent attention of 110 us + RoIAlign: 142 us), the overlapped bench step
// is 7.19-7.23 M RoIs/s for every variant; streaming (evict-first) loads of the query maps change nothing either.
template <int CH, int MINB>
__global__ void __launch_bounds__(256, MINB)
channel_attention_ml_kernel(const MlLevels lv, const float *__restrict__ vec, const int B, const int N, const int C,
                            const int nblocks)
{
  for (int gb = blockIdx.x; gb < nblocks; gb += gridDim.x) {         // (grid < nblocks: the persistent form, FGN_ATT_GRID)
    const int l = ml_level_of(lv, gb), blk = gb - lv.blk_off[l];
    const int c4 = C >> 2;
    const size_t per_img = (size_t)lv.HW[l] * c4, total = (size_t)B * per_img;
    const float *q = lv.in[l];
    float *o = lv.out[l];
    const float *v = vec + (size_t)l * B * N * C;
    const size_t i0 = (size_t)blk * (256 * CH) + threadIdx.x;
    float4 x[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const size_t i = i0 + (size_t)j * 256;
        x[j] = i < total ? (lv.stream_loads ? __ldcs(reinterpret_cast<const float4 *>(q) + i) : ldg4(q + 4 * i))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const size_t i = i0 + (size_t)j * 256;
        if (i >= total) continue;
        const int b = i / per_img;
        const size_t rem = i - (size_t)b * per_img;
        const int cv = rem % c4;
        for (int n = 0; n < N; ++n) {
            const float4 s = ldg4(v + ((size_t)b * N + n) * C + cv * 4);
            __stcs(reinterpret_cast<float4 *>(o + ((size_t)b * N + n) * per_img * 4) + rem,
                   make_float4(x[j].x * s.x, x[j].y * s.y, x[j].z * s.z, x[j].w * s.w));
        }
    }
  }
}

// ---- bf16 variants: 8 channels per 128-bit access, fp32 accumulation / multiply ----------------
__device__ __forceinline__ void unpack8(const uint4 &q, float (&f)[8])
{
    f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
    f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
    f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xffff0000u);
    f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ unsigned pack2(float a, float b)       // round to nearest even
{
    unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
    ua += 0x7fffu + ((ua >> 16) & 1u);
    ub += 0x7fffu + ((ub >> 16) & 1u);
    return (ua >> 16) | (ub & 0xffff0000u);
}

__global__ void __launch_bounds__(256)
attention_vec_ml_partial_bf16_kernel(const MlLevels lv, const int BN, const int K, const int C,
                                     float *partial, float *__restrict__ vec, unsigned int *counters)
{
    const int bn = blockIdx.y;
    const int l = ml_level_of(lv, blockIdx.x), slab = blockIdx.x - lv.blk_off[l];
    const int c8 = C >> 3;
    const int total = K * lv.HW[l];
    const int p0 = slab * kMlSlab, p1 = min(total, p0 + kMlSlab);
    extern __shared__ __align__(16) float red[];   // [rows][C]
    const int rows = max(1, (int)blockDim.x / c8);
    const unsigned short *base = reinterpret_cast<const unsigned short *>(lv.in[l]) + (size_t)bn * total * C;
    const int cv = threadIdx.x % c8, rowi = threadIdx.x / c8;
    if (rowi < rows) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int p = p0 + rowi; p < p1; p += rows) {
            float f[8];
            unpack8(__ldg(reinterpret_cast<const uint4 *>(base + (size_t)p * C + cv * 8)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(size_t)rowi * C + cv * 8 + j] = a[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s += red[(size_t)rr * C + c];
        partial[((size_t)blockIdx.x * BN + bn) * C + c] = s;
    }
    ml_finish(lv, l, bn, BN, K, C, partial, vec, counters);
}

__global__ void __launch_bounds__(256)
channel_attention_ml_bf16_kernel(const MlLevels lv, const float *__restrict__ vec, const int B, const int N, const int C)
{
    const int l = ml_level_of(lv, blockIdx.x), blk = blockIdx.x - lv.blk_off[l];
    const int c8 = C >> 3;
    const size_t per_img = (size_t)lv.HW[l] * c8, total = (size_t)B * per_img;
    const uint4 *q = reinterpret_cast<const uint4 *>(lv.in[l]);
    uint4 *o = reinterpret_cast<uint4 *>(lv.out[l]);
    const float *v = vec + (size_t)l * B * N * C;
    const size_t i0 = (size_t)blk * (256 * kMlChunk) + threadIdx.x;
    uint4 x[kMlChunk];
#pragma unroll
    for (int j = 0; j < kMlChunk; ++j) {
        const size_t i = i0 + (size_t)j * 256;
        x[j] = i < total ? __ldg(q + i) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < kMlChunk; ++j) {
        const size_t i = i0 + (size_t)j * 256;
        if (i >= total) continue;
        const int b = i / per_img;
        const size_t rem = i - (size_t)b * per_img;
        const int cv = rem % c8;
        float f[8];
        unpack8(x[j], f);
        for (int n = 0; n < N; ++n) {
            const float *s = v + ((size_t)b * N + n) * C + cv * 8;
            const float4 s0 = ldg4(s), s1 = ldg4(s + 4);
            uint4 r;
            r.x = pack2(f[0] * s0.x, f[1] * s0.y); r.y = pack2(f[2] * s0.z, f[3] * s0.w);
            r.z = pack2(f[4] * s1.x, f[5] * s1.y); r.w = pack2(f[6] * s1.z, f[7] * s1.w);
            __stcs(o + ((size_t)b * N + n) * per_img + rem, r);
        }
    }
}

static int fill_levels(const fgn_pyramid_t *p, MlLevels &lv)
{
    FGN_CHECK_ARG(p != nullptr && p->num_levels >= 1 && p->num_levels <= FGN_MAX_LEVELS, "bad pyramid");
    lv.L = p->num_levels;
    for (int l = 0; l < FGN_MAX_LEVELS; ++l) {
        lv.in[l] = l < lv.L ? p->feat[l] : nullptr;
        lv.out[l] = nullptr;
        lv.HW[l] = l < lv.L ? p->H[l] * p->W[l] : 0;
        FGN_CHECK_ARG(l >= lv.L || (lv.HW[l] > 0 && lv.in[l] != nullptr), "level %d is empty", l);
    }
    return FGN_OK;
}

}  // namespace fgn

using namespace fgn;

extern "C" size_t fgn_attention_vectors_ml_workspace_bytes(const fgn_pyramid_t *spp, int BN, int K, int C)
{
    if (!spp || BN <= 0 || K <= 0 || C <= 0) return 0;
    size_t blocks = 0;
    for (int l = 0; l < spp->num_levels && l < FGN_MAX_LEVELS; ++l)
        blocks += (size_t)ceil_div(K * spp->H[l] * spp->W[l], kMlSlab);
    return ((blocks * BN * C * sizeof(float) + 255) & ~(size_t)255) + (size_t)FGN_MAX_LEVELS * BN * sizeof(unsigned int);
}

static int attention_vectors_ml_impl(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec, void *workspace,
                                     size_t workspace_bytes, void *stream, bool bf16);

extern "C" int fgn_attention_vectors_ml(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec,
                                        void *workspace, size_t workspace_bytes, void *stream)
{
    return attention_vectors_ml_impl(spp, BN, K, C, vec, workspace, workspace_bytes, stream, false);
}

extern "C" int fgn_attention_vectors_ml_bf16(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec,
                                             void *workspace, size_t workspace_bytes, void *stream)
{
    return attention_vectors_ml_impl(spp, BN, K, C, vec, workspace, workspace_bytes, stream, true);
}

static int attention_vectors_ml_impl(const fgn_pyramid_t *spp, int BN, int K, int C, float *vec, void *workspace,
                                     size_t workspace_bytes, void *stream, bool bf16)
{
    FGN_CHECK_ARG(BN >= 0 && K > 0 && C > 0, "bad dims");
    if (BN == 0) return FGN_OK;
    if ((C & (bf16 ? 7 : 3)) || C > 1024) { set_error("attention_vectors_ml needs C%%4==0 (bf16: C%%8==0) and C<=1024 (C=%d)", C); return FGN_ERR_UNSUPPORTED; }
    MlLevels lv;
    int rc = fill_levels(spp, lv);
    if (rc) return rc;
    FGN_CHECK_ARG(vec != nullptr, "NULL pointer");
    lv.blk_off[0] = 0;
    for (int l = 0; l < lv.L; ++l) lv.blk_off[l + 1] = lv.blk_off[l] + ceil_div(K * lv.HW[l], kMlSlab);
    for (int l = lv.L + 1; l <= FGN_MAX_LEVELS; ++l) lv.blk_off[l] = lv.blk_off[lv.L];
    const size_t need = fgn_attention_vectors_ml_workspace_bytes(spp, BN, K, C);
    if (!workspace || workspace_bytes < need) {
        set_error("attention_vectors_ml: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    FGN_CHECK_ARG(BN <= 65535, "BN too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int cvec = bf16 ? C >> 3 : C >> 2, rows = max(1, 256 / cvec);
    unsigned int *counters = (unsigned int *)((char *)workspace + (((size_t)lv.blk_off[lv.L] * BN * C * sizeof(float) + 255) & ~(size_t)255));
    FGN_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)lv.L * BN * sizeof(unsigned int), st));
    if (bf16) attention_vec_ml_partial_bf16_kernel<<<dim3(lv.blk_off[lv.L], BN), 256, (size_t)rows * C * 4, st>>>(
                  lv, BN, K, C, (float *)workspace, vec, counters);
    else      attention_vec_ml_partial_kernel<<<dim3(lv.blk_off[lv.L], BN), 256, (size_t)rows * C * 4, st>>>(
                  lv, BN, K, C, (float *)workspace, vec, counters);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

static int channel_attention_ml_impl(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                                     float *const *out_host, void *stream, bool bf16);

extern "C" int fgn_channel_attention_ml(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                                        float *const *out_host, void *stream)
{
    return channel_attention_ml_impl(qry, vec, B, N, C, out_host, stream, false);
}

extern "C" int fgn_channel_attention_ml_bf16(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                                             void *const *out_host, void *stream)
{
    return channel_attention_ml_impl(qry, vec, B, N, C, (float *const *)out_host, stream, true);
}

static int channel_attention_ml_impl(const fgn_pyramid_t *qry, const float *vec, int B, int N, int C,
                                     float *const *out_host, void *stream, bool bf16)
{
    FGN_CHECK_ARG(B >= 0 && N > 0 && C > 0, "bad dims");
    if (B == 0) return FGN_OK;
    if (C & (bf16 ? 7 : 3)) { set_error("channel_attention_ml needs C%%4==0 (bf16: C%%8==0) (C=%d)", C); return FGN_ERR_UNSUPPORTED; }
    MlLevels lv;
    int rc = fill_levels(qry, lv);
    if (rc) return rc;
    FGN_CHECK_ARG(vec && out_host, "NULL pointer");
    // FGN_ATT_LEAN (development knob): 0 = 4 float4 per thread, 48 registers (round 1); 1 = same at <= 40 registers; 2 = lean
    const char *el = getenv("FGN_ATT_LEAN");
    const int lean = el != nullptr ? atoi(el) : 0;
    const int chunk = bf16 ? kMlChunk : (lean >= 2 ? 2 : 4);
    lv.blk_off[0] = 0;
    for (int l = 0; l < lv.L; ++l) {
        FGN_CHECK_ARG(out_host[l] != nullptr, "output level %d is NULL", l);
        lv.out[l] = out_host[l];
        const size_t elems = (size_t)B * lv.HW[l] * (bf16 ? C >> 3 : C >> 2);
        lv.blk_off[l + 1] = lv.blk_off[l] + (int)((elems + 256 * chunk - 1) / (256 * chunk));
    }
    for (int l = lv.L + 1; l <= FGN_MAX_LEVELS; ++l) lv.blk_off[l] = lv.blk_off[lv.L];
    const int nblocks = lv.blk_off[lv.L];
    { const char *es = getenv("FGN_ATT_LDCS"); lv.stream_loads = es != nullptr ? atoi(es) : 0; }
    int grid = nblocks;
    if (const char *eg = getenv("FGN_ATT_GRID")) {                  // development knob: persistent form, CTAs per SM
        int sm = 0;
        if (int rcs = current_sm_count(&sm)) return rcs;
        if (atoi(eg) > 0) grid = min(nblocks, atoi(eg) * sm);
    }
    if (bf16)            channel_attention_ml_bf16_kernel<<<lv.blk_off[lv.L], 256, 0, (cudaStream_t)stream>>>(lv, vec, B, N, C);
    else if (lean == 0)  channel_attention_ml_kernel<4, 5><<<grid, 256, 0, (cudaStream_t)stream>>>(lv, vec, B, N, C, nblocks);
    else if (lean == 1)  channel_attention_ml_kernel<4, 6><<<grid, 256, 0, (cudaStream_t)stream>>>(lv, vec, B, N, C, nblocks);
    else                 channel_attention_ml_kernel<2, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(lv, vec, B, N, C, nblocks);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

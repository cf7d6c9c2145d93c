// relation.cu -- Relation-Guided Detector fusion + box head, fused.
//
// Reference: FGNRoIHead.count_one_roi_by_n_spp (fgn_roi_head.py:253-279), BBoxHead.forward with
// with_avg_pool=True [3P] (called at fgn_roi_head.py:338, config fgn_r50_c4_densecl.py:76-93) and
// count_modified_cls_bbox (fgn_roi_head.py:302-326).
//
// The reference materialises X[r*N+n] = cat(q[r], s[b(r),n]) ([R*N,2C,P,P]) and runs a 1x1 conv
// over it.  Here the conv is split by linearity,
//     conv(cat(q,s)) = Wq q[r] + (Ws s[b,n] + bias)            Wq = W[:, :C], Ws = W[:, C:]
// so the contraction runs once per RoI (M = R*P*P rows) and once per class (M = B*N*P*P rows),
// not once per (RoI, class) pair.  The (RoI, class) work that remains is elementwise: add,
// GroupNorm(32) statistics, affine + ReLU, PxP average pool and the two tiny FCs, all done in
// registers by relation_epilogue_kernel; the re-assembly to [R,N+1]/[R,4N] rides in the finalize
// kernel.
#include "common.cuh"
#include <stdlib.h>
#include "gemm.cuh"

namespace fgn {

constexpr int kEpiThreads = 256;
constexpr int kMaxPP = 49;          // epilogue register tile is P*P = 49 (P = 7)

__device__ __forceinline__ float ld_nc_ordered(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float group_sum_shfl(float v, int cg)
{
    for (int o = cg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (R, nblk); block kEpiThreads; thread = one channel of the block.
// Yq [R*PP, C] and Ys [BN*PP, C] are the split-conv outputs (Ys already carries the conv bias).
// partial [R, N, 6, nblk]: per-channel-block partial dot products of the pooled vector with the
// 2 cls rows and 4 reg rows.
// FAST: every thread owns a channel (C is a multiple of the block's channel count) and a GroupNorm group lies inside
// one warp -- the production shapes (C = 256 / 1024, 32 groups); no per-element predication, reductions by shuffle.
// ONE: a single class (N = 1, the headline configuration): the RoI term is added to the class term as it arrives instead
// of being held across a class loop -- 49 fewer live registers, three CTAs per SM instead of two, i.e. half as many
// bytes again in flight for a kernel that is one read of the conv output.  Same operations in the same order.
// (C is a compile-time 256 there: with a run-time row pitch ptxas materialises the 98 load addresses and spills them.)
template <int PP, bool FAST, bool ONE = false>
__global__ void __launch_bounds__(kEpiThreads, ONE ? 3 : 2)
relation_epilogue_kernel(const float *__restrict__ Yq, const float *__restrict__ Ys,
                         const int32_t *__restrict__ roi_batch, const int R, const int B, const int N,
                         const int C_rt, const int cblk, const int cg, const float eps,
                         const float *__restrict__ gn_w, const float *__restrict__ gn_b,
                         const float *__restrict__ fc_cls_w, const float *__restrict__ fc_reg_w,
                         float *__restrict__ partial,
                         const float *__restrict__ rois5, const float *__restrict__ fc_cls_b,
                         const float *__restrict__ fc_reg_b, float *__restrict__ cls_out, float *__restrict__ reg_out,
                         float *__restrict__ raw_cls, float *__restrict__ raw_reg)
{
    // rois5 (optional): the [R,5] RoI tensor itself -- the batch index is read from its first column, no separate
    // roi_batch launch.  cls_out (optional, one channel block only): the FC biases and count_modified_cls_bbox are
    // applied here and the final [R,N+1] / [R,4N] rows written, no partial buffer and no finalize launch.
    const int C = ONE ? kEpiThreads : C_rt;
    extern __shared__ float sm[];                 // [N][6][nwarps] + [blockDim] scratch
    const int r = blockIdx.x, blk = blockIdx.y, nblk = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float *fc_part = sm;                          // [N*6*nwarps]
    float *scratch = sm + (size_t)N * 6 * nwarps; // [blockDim.x]
    const int c = blk * cblk + tid;
    const bool active = FAST || (tid < cblk && c < C);
    const bool shfl_ok = FAST || ((cg & (cg - 1)) == 0 && cg <= 32);
    int b = rois5 != nullptr ? (int)rois5[5 * (size_t)r] : roi_batch[r];
    b = b < 0 ? 0 : (b >= B ? B - 1 : b);

    float yq[ONE ? 1 : PP];
    if (!ONE) {
#pragma unroll
        for (int p = 0; p < PP; ++p) yq[ONE ? 0 : p] = active ? __ldg(Yq + ((size_t)r * PP + p) * C + c) : 0.f;
    }
    const float gamma = active ? __ldg(gn_w + c) : 0.f, beta = active ? __ldg(gn_b + c) : 0.f;
    float wfc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j)
        wfc[j] = active ? (j < 2 ? __ldg(fc_cls_w + (size_t)j * C + c) : __ldg(fc_reg_w + (size_t)(j - 2) * C + c)) : 0.f;
    const float inv_cnt = 1.0f / (float)(cg * PP);

    for (int n = 0; n < N; ++n) {
        const float *ys = Ys + ((size_t)(b * N + n) * PP) * C + c;
        float y[PP];
        float s1 = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            if (ONE) y[p] = ld_nc_ordered(Yq + ((size_t)r * PP + p) * C + c);
            else     y[p] = active ? yq[ONE ? 0 : p] + __ldg(ys + (size_t)p * C) : 0.f;
        }
        if (ONE) {                                  // the class term (L2-resident) behind the 49 loads above, in program order
#pragma unroll                                      // (volatile: left to itself nvcc hoists all 98 loads and spills at 80 registers)
            for (int p = 0; p < PP; ++p) y[p] = y[p] + ld_nc_ordered(ys + (size_t)p * C);
        }
#pragma unroll
        for (int p = 0; p < PP; ++p) s1 += y[p];
        float gs;
        if (shfl_ok) gs = group_sum_shfl(s1, cg);
        else {
            __syncthreads(); scratch[tid] = s1; __syncthreads();
            const int g0 = (tid / cg) * cg; gs = 0.f;
            for (int i = 0; i < cg; ++i) gs += scratch[g0 + i];
        }
        const float mean = gs * inv_cnt;
        float s2 = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) { const float d = y[p] - mean; s2 = fmaf(d, d, s2); }
        if (!active) s2 = 0.f;
        if (shfl_ok) gs = group_sum_shfl(s2, cg);
        else {
            __syncthreads(); scratch[tid] = s2; __syncthreads();
            const int g0 = (tid / cg) * cg; gs = 0.f;
            for (int i = 0; i < cg; ++i) gs += scratch[g0 + i];
        }
        const float rstd = 1.0f / sqrtf(gs * inv_cnt + eps);
        // GroupNorm affine as torch does it: y*scale + shift, scale = rstd*gamma
        const float scale = rstd * gamma, shift = beta - mean * scale;
        float z = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) z += fmaxf(fmaf(y[p], scale, shift), 0.f);
        z = active ? z / (float)PP : 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = warp_sum(z * wfc[j]);
            if (lane == 0) fc_part[((size_t)n * 6 + j) * nwarps + warp] = v;
        }
    }
    __syncthreads();
    if (cls_out == nullptr || nblk != 1) {
        for (int i = tid; i < N * 6; i += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w) s += fc_part[(size_t)i * nwarps + w];
            partial[((size_t)r * N * 6 + i) * nblk + blk] = s;
        }
        return;
    }
    // one channel block: this CTA holds the RoI's complete head outputs (same order of additions as the finalize kernel)
    for (int i = tid; i < N * 6; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += fc_part[(size_t)i * nwarps + w];
        const int j = i % 6;
        fc_part[(size_t)i * nwarps] = (0.f + s) + (j == 0 ? fc_cls_b[0] : j == 1 ? fc_cls_b[1] : fc_reg_b[j - 2]);
    }
    __syncthreads();
    if (tid == 0) {
        float best_fg = 0.f, best_bg = 0.f;
        for (int n = 0; n < N; ++n) {
            float v[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = fc_part[((size_t)n * 6 + j) * nwarps];
            if (raw_cls) { raw_cls[((size_t)r * N + n) * 2] = v[0]; raw_cls[((size_t)r * N + n) * 2 + 1] = v[1]; }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                reg_out[(size_t)r * 4 * N + 4 * n + d] = v[2 + d];
                if (raw_reg) raw_reg[((size_t)r * N + n) * 4 + d] = v[2 + d];
            }
            cls_out[(size_t)r * (N + 1) + n] = v[1];
            const bool better = n == 0 || v[1] > best_fg || (v[1] != v[1] && best_fg == best_fg);
            if (better) { best_fg = v[1]; best_bg = v[0]; }
        }
        cls_out[(size_t)r * (N + 1) + N] = best_bg;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The same epilogue for the headline shape (N = 1, C = 256, P = 7) as a persistent, bulk-copy-fed kernel.  The
// one-CTA-per-RoI kernel above is a read of the conv output with its loads in flight only during the first third of
// every CTA's life (3.8 TB/s).  Here one CTA per SM owns a contiguous range of RoIs; a producer thread keeps a ring of
// three 50 KB RoI tiles ([49,256] fp32, contiguous in the conv output) filled with cp.async.bulk, so ~150 KB per SM are
// always in flight; two teams of 256 consumer threads (thread = channel) take alternate RoIs: a tile goes into registers,
// the stage is freed at once, and the arithmetic is exactly relation_epilogue_kernel's, in the same order (one team
// alone is latency-bound on its shuffle chains: 1.5 us per RoI).  The class term of the RoI's image (also one 50 KB
// tile) sits in shared memory and is reloaded, behind a barrier of both teams, when the image changes (RoIs are grouped
// by image: once or twice per CTA).
// ------------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t ep_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ep_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ep_s32(bar)), "r"(count));
}
__device__ __forceinline__ void ep_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ep_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ep_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ep_s32(bar)) : "memory");
}
__device__ __forceinline__ void ep_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ep_s32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void ep_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ep_s32(dst)), "l"(src), "r"(bytes), "r"(ep_s32(bar)) : "memory");
}
constexpr int kRingStages = 3;
constexpr int kRingTeams = 2;
}  // namespace

// grid = persistent CTAs, block = 2 x 256 consumers + 1 producer warp; dynamic smem = (kRingStages + 1) tiles of PP*256 floats
template <int PP, int TEAMS>
__global__ void __launch_bounds__(TEAMS * kEpiThreads + 32, 1)
relation_epilogue_ring_kernel(const float *__restrict__ Yq, const float *__restrict__ Ys,
                              const int32_t *__restrict__ roi_batch, const float *__restrict__ rois5,
                              const int R, const int B, const float eps,
                              const float *__restrict__ gn_w, const float *__restrict__ gn_b,
                              const float *__restrict__ fc_cls_w, const float *__restrict__ fc_reg_w,
                              const float *__restrict__ fc_cls_b, const float *__restrict__ fc_reg_b,
                              float *__restrict__ cls_out, float *__restrict__ reg_out,
                              float *__restrict__ raw_cls, float *__restrict__ raw_reg)
{
    constexpr int C = kEpiThreads, cg = C / 32, nwarps = kEpiThreads / 32;
    constexpr int kTile = PP * C;                                   // floats per tile
    constexpr uint32_t kTileBytes = kTile * 4;
    extern __shared__ __align__(128) float ring[];                  // [kRingStages][kTile] | ys[kTile]
    float *ys_tile = ring + (size_t)kRingStages * kTile;
    constexpr int kFull = kRingStages * TEAMS;                      // (stage, team) pairs: RoI k signals full_bar[k % kFull]
    static_assert(TEAMS == 1 || (kRingStages % TEAMS) != 0, "stages and teams must interleave");
    __shared__ __align__(8) uint64_t full_bar[kFull], empty_bar[kRingStages], ys_bar;
    __shared__ float fc_part[TEAMS][2][6 * nwarps];
    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid / kEpiThreads, warp = (tid >> 5) % nwarps;     // team TEAMS = the producer warp
    const int per = (R + gridDim.x - 1) / gridDim.x;
    const int r_beg = blockIdx.x * per, r_end = min(R, r_beg + per);
    if (tid == 0) {
        for (int s = 0; s < kRingStages; ++s) ep_mbar_init(&empty_bar[s], nwarps);
        for (int s = 0; s < kFull; ++s) ep_mbar_init(&full_bar[s], 1);
        ep_mbar_init(&ys_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (r_beg >= r_end) return;

    if (team == TEAMS) {
        // ===== producer: one RoI tile per stage ==========================================================
        if (lane == 0) {
            int s = 0, f = 0;
            unsigned par = 1;                                        // first pass over fresh barriers
            for (int r = r_beg; r < r_end; ++r) {
                ep_mbar_wait(&empty_bar[s], par);
                ep_mbar_expect_tx(&full_bar[f], kTileBytes);
                ep_bulk_g2s(ring + (size_t)s * kTile, Yq + (size_t)r * kTile, kTileBytes, &full_bar[f]);
                if (++s == kRingStages) { s = 0; par ^= 1; }
                if (++f == kFull) f = 0;
            }
        }
        return;
    }

    // ===== consumers: thread = channel, team = RoI parity ================================================
    const int c = tid % kEpiThreads;
    const float gamma = __ldg(gn_w + c), beta = __ldg(gn_b + c);
    float wfc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) wfc[j] = j < 2 ? __ldg(fc_cls_w + (size_t)j * C + c) : __ldg(fc_reg_w + (size_t)(j - 2) * C + c);
    const float inv_cnt = 1.0f / (float)(cg * PP);
    float bias = 0.f;                                                // lanes 0..5 of warp 0 finish output j = lane
    if (c < 6) bias = c == 0 ? fc_cls_b[0] : c == 1 ? fc_cls_b[1] : fc_reg_b[c - 2];
    int cur_b = -1;
    unsigned ys_par = 0;
    for (int r = r_beg, k = 0; r < r_end; ++r, ++k) {
        int b = rois5 != nullptr ? (int)rois5[5 * (size_t)r] : roi_batch[r];
        b = b < 0 ? 0 : (b >= B ? B - 1 : b);
        if (b != cur_b) {
            // both teams walk the same RoI sequence, so both arrive here: everybody is done with the old class term
            asm volatile("bar.sync 3, %0;" ::"n"(TEAMS * kEpiThreads) : "memory");
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                ep_mbar_expect_tx(&ys_bar, kTileBytes);
                ep_bulk_g2s(ys_tile, Ys + (size_t)b * kTile, kTileBytes, &ys_bar);
            }
            ep_mbar_wait(&ys_bar, ys_par);
            ys_par ^= 1;
            cur_b = b;
        }
        if ((k % TEAMS) != team) continue;
        const int s = k % kRingStages;
        const unsigned par = (unsigned)(k / kFull) & 1u;
        // A parity wait only tells the current phase from the one before it, and with two teams a stage's previous use
        // belongs to the other team: on a per-stage barrier a team that runs ahead would find "its" parity already
        // satisfied by the use before that one.  So a "full" barrier per (stage, team): each is waited on by one team, in
        // order, phase by phase.
        ep_mbar_wait(&full_bar[k % kFull], par);
        const float *tq = ring + (size_t)s * kTile + c;
        float y[PP];
#pragma unroll
        for (int p = 0; p < PP; ++p) y[p] = tq[p * C];
        __syncwarp();
        if (lane == 0) ep_mbar_arrive(&empty_bar[s]);               // the tile is in registers: refill the stage
        float s1 = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) y[p] = y[p] + ys_tile[p * C + c];
#pragma unroll
        for (int p = 0; p < PP; ++p) s1 += y[p];
        float gs = group_sum_shfl(s1, cg);
        const float mean = gs * inv_cnt;
        float s2 = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) { const float d = y[p] - mean; s2 = fmaf(d, d, s2); }
        gs = group_sum_shfl(s2, cg);
        const float rstd = 1.0f / sqrtf(gs * inv_cnt + eps);
        const float scale = rstd * gamma, shift = beta - mean * scale;
        float z = 0.f;
#pragma unroll
        for (int p = 0; p < PP; ++p) z += fmaxf(fmaf(y[p], scale, shift), 0.f);
        z = z / (float)PP;
        float *fp = fc_part[team][(k / TEAMS) & 1];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = warp_sum(z * wfc[j]);
            if (lane == 0) fp[j * nwarps + warp] = v;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kEpiThreads) : "memory");
        // (fc_part is double-buffered per team: nobody writes this half again before the team's warp 0 has passed its next barrier)
        if (c < 6) {
            float sum = 0.f;
            for (int w = 0; w < nwarps; ++w) sum += fp[c * nwarps + w];
            const float v = (0.f + sum) + bias;
            if (c < 2) {                                             // N = 1: cls_out = (fg, bg) = raw[:, [1, 0]]
                cls_out[(size_t)r * 2 + (1 - c)] = v;
                if (raw_cls) raw_cls[(size_t)r * 2 + c] = v;
            } else {
                reg_out[(size_t)r * 4 + (c - 2)] = v;
                if (raw_reg) raw_reg[(size_t)r * 4 + (c - 2)] = v;
            }
        }
    }
}

// One thread per RoI: sum channel-block partials, add FC biases, write raw [R*N,2]/[R*N,4] if
// asked, and re-assemble (count_modified_cls_bbox, generalised from N in {1,3} to any N):
//   N == 1: cls_out = raw[:, [1,0]]
//   N  > 1: fg_n = raw[n,1]; j = argmax_n fg_n (first max); cls_out = (fg_0..fg_{N-1}, raw[j,0])
//   reg_out = raw_reg.view(R, 4N)
__global__ void relation_finalize_kernel(const float *__restrict__ partial, const int R, const int N,
                                         const int nblk, const float *__restrict__ fc_cls_b,
                                         const float *__restrict__ fc_reg_b,
                                         float *__restrict__ cls_out, float *__restrict__ reg_out,
                                         float *__restrict__ raw_cls, float *__restrict__ raw_reg)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float bc0 = fc_cls_b[0], bc1 = fc_cls_b[1];
    float best_fg = 0.f, best_bg = 0.f;
    for (int n = 0; n < N; ++n) {
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float s = 0.f;
            const float *p = partial + ((size_t)r * N * 6 + (size_t)n * 6 + j) * nblk;
            for (int k = 0; k < nblk; ++k) s += p[k];
            v[j] = s + (j == 0 ? bc0 : j == 1 ? bc1 : fc_reg_b[j - 2]);
        }
        if (raw_cls) { raw_cls[((size_t)r * N + n) * 2] = v[0]; raw_cls[((size_t)r * N + n) * 2 + 1] = v[1]; }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            reg_out[(size_t)r * 4 * N + 4 * n + d] = v[2 + d];
            if (raw_reg) raw_reg[((size_t)r * N + n) * 4 + d] = v[2 + d];
        }
        cls_out[(size_t)r * (N + 1) + n] = v[1];
        const bool better = n == 0 || v[1] > best_fg || (v[1] != v[1] && best_fg == best_fg);
        if (better) { best_fg = v[1]; best_bg = v[0]; }
    }
    cls_out[(size_t)r * (N + 1) + N] = best_bg;
}

// Stand-alone count_modified_cls_bbox (fgn_roi_head.py:302-326) on already computed head outputs.
__global__ void cls_bbox_reassemble_kernel(const float *__restrict__ raw_cls, const float *__restrict__ raw_reg,
                                           const int R, const int N, float *__restrict__ cls_out,
                                           float *__restrict__ reg_out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float best_fg = 0.f, best_bg = 0.f;
    for (int n = 0; n < N; ++n) {
        const float bg = raw_cls[((size_t)r * N + n) * 2], fg = raw_cls[((size_t)r * N + n) * 2 + 1];
        cls_out[(size_t)r * (N + 1) + n] = fg;
        if (n == 0 || fg > best_fg || (fg != fg && best_fg == best_fg)) { best_fg = fg; best_bg = bg; }
#pragma unroll
        for (int d = 0; d < 4; ++d) reg_out[(size_t)r * 4 * N + 4 * n + d] = raw_reg[((size_t)r * N + n) * 4 + d];
    }
    cls_out[(size_t)r * (N + 1) + N] = best_bg;
}

__global__ void roi_batch_kernel(const float *__restrict__ rois, int R, int32_t *__restrict__ out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) out[r] = (int)rois[5 * (size_t)r];
}

struct RelationWs {
    float *yq, *ys, *partial, *xq_nhwc, *xs_nhwc, *split_q, *split_s;
    size_t bytes;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static RelationWs carve(void *base, int R, int BN, int C, int P, int N_for_partial)
{
    RelationWs w;
    const size_t PP = (size_t)P * P;
    size_t off = 0;
    char *b = (char *)base;
    auto take = [&](size_t bytes) { char *p = b ? b + off : nullptr; off += align256(bytes); return (float *)p; };
    w.yq      = take((size_t)R * PP * C * 4);
    w.ys      = take((size_t)BN * PP * C * 4);
    w.xq_nhwc = take((size_t)R * PP * C * 4);     // only used when roi_feat arrives NCHW
    w.xs_nhwc = take((size_t)BN * PP * C * 4);
    w.split_q = take(gemm_tc_workspace_bytes(C, C));
    w.split_s = take(gemm_tc_workspace_bytes(C, C));
    const int nblk_max = ceil_div(C, 32);
    w.partial = take((size_t)R * (size_t)N_for_partial * 6 * nblk_max * 4);
    w.bytes = off;
    return w;
}

}  // namespace fgn

using namespace fgn;

extern "C" size_t fgn_relation_fusion_workspace_bytes(int R, int BN, int C, int P)
{
    if (R < 0 || BN <= 0 || C <= 0 || P <= 0) return 0;
    // partial is sized for N <= BN classes per RoI
    return carve(nullptr, R, BN, C, P, BN).bytes;
}

extern "C" int fgn_nchw_to_nhwc(const float *, int, int, int, int, float *, void *);

// conv_w_split (optional): fgn_relation_split_weights' output for this conv_w -- the TF32 split of Wq and Ws made once
// when the weights were loaded instead of by two launches per call.  rois5 (optional, instead of roi_batch): the [R,5]
// RoI tensor, whose first column is the batch index.
static int relation_fusion_impl(const float *roi_feat, int feat_layout, const int32_t *roi_batch, const float *rois5,
                                const float *spp_cat_mean, const float *class_term, int R, int B, int N, int C, int P, const float *conv_w,
                                const float *conv_w_split, const float *conv_b, const float *gn_w, const float *gn_b,
                                int gn_groups, float gn_eps, const float *fc_cls_w, const float *fc_cls_b,
                                const float *fc_reg_w, const float *fc_reg_b, float *cls_out, float *reg_out,
                                float *raw_cls_out, float *raw_reg_out, int precision, void *workspace,
                                size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(R >= 0 && B > 0 && N > 0 && C > 0 && P > 0, "bad dims R=%d B=%d N=%d C=%d P=%d", R, B, N, C, P);
    FGN_CHECK_ARG(gn_groups > 0 && C % gn_groups == 0, "GroupNorm groups=%d does not divide C=%d", gn_groups, C);
    FGN_CHECK_ARG(precision == 0 || precision == 1, "precision=%d", precision);
    if (R == 0) return FGN_OK;
    if (P * P != kMaxPP) { set_error("relation_fusion: P=%d not instantiated (7)", P); return FGN_ERR_UNSUPPORTED; }
    FGN_CHECK_ARG(roi_feat && (roi_batch || rois5) && (spp_cat_mean || class_term) && conv_w && conv_b && gn_w && gn_b &&
                  fc_cls_w && fc_cls_b && fc_reg_w && fc_reg_b && cls_out && reg_out, "NULL pointer");
    const int BN = B * N, PP = P * P;
    const size_t need = fgn_relation_fusion_workspace_bytes(R, BN, C, P);
    if (!workspace || workspace_bytes < need) {
        set_error("relation_fusion: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RelationWs w = carve(workspace, R, BN, C, P, BN);

    // operands as row-major [rows, C] (NHWC); repack when the caller hands over NCHW
    const float *xq = roi_feat, *xs = spp_cat_mean;
    if (feat_layout == FGN_LAYOUT_NCHW) {
        int rc = fgn_nchw_to_nhwc(roi_feat, R, C, P, P, w.xq_nhwc, stream);
        if (rc) return rc;
        if (class_term == nullptr) {
            rc = fgn_nchw_to_nhwc(spp_cat_mean, BN, C, P, P, w.xs_nhwc, stream);
            if (rc) return rc;
        }
        xq = w.xq_nhwc; xs = w.xs_nhwc;
    }
    // Yq = Xq Wq^T ; Ys = Xs Ws^T + bias        (conv_w is [C, 2C] row-major)
    float *split_q = conv_w_split ? const_cast<float *>(conv_w_split) : w.split_q;
    float *split_s = conv_w_split ? const_cast<float *>(conv_w_split) + (size_t)2 * C * C : w.split_s;
    int rc = gemm_nt(xq, C, conv_w, 2 * C, nullptr, w.yq, C, R * PP, C, C, precision, split_q, st, conv_w_split != nullptr);
    if (rc) return rc;
    // class_term: Ys precomputed by fgn_support_prologue_fwd (exact fp32) -- one launch less per call
    const float *ys = class_term;
    if (ys == nullptr) {
        rc = gemm_nt(xs, C, conv_w + C, 2 * C, conv_b, w.ys, C, BN * PP, C, C, precision, split_s, st, conv_w_split != nullptr);
        if (rc) return rc;
        ys = w.ys;
    }

    const int cg = C / gn_groups;
    FGN_CHECK_ARG(cg <= kEpiThreads, "channels per group %d > %d", cg, kEpiThreads);
    const int cblk = (kEpiThreads / cg) * cg;
    const int nblk = ceil_div(C, cblk);
    const int nwarps = kEpiThreads / 32;
    const size_t smem = ((size_t)N * 6 * nwarps + kEpiThreads) * 4;
    FGN_CHECK_ARG(nblk <= 65535, "nblk");
    // production shapes (every thread owns a channel, a GroupNorm group inside one warp) take the unpredicated kernel
    const bool fast = cblk == kEpiThreads && (C % cblk) == 0 && (cg & (cg - 1)) == 0 && cg <= 32;
    const char *eo = getenv("FGN_EPI_ONE");                   // development knob: 0 = the general kernel for N = 1 too
    const bool one = fast && N == 1 && C == kEpiThreads && !(eo != nullptr && atoi(eo) == 0);
    if (one) FGN_SMEM_OPTIN((relation_epilogue_kernel<kMaxPP, true, true>), smem);
    else if (fast) FGN_SMEM_OPTIN((relation_epilogue_kernel<kMaxPP, true>), smem);
    else      FGN_SMEM_OPTIN((relation_epilogue_kernel<kMaxPP, false>), smem);
    // one channel block (C <= 256): the epilogue writes the final rows itself, no partial buffer, no finalize launch
    float *direct = nblk == 1 ? cls_out : nullptr;
    const char *er = getenv("FGN_EPI_RING");                  // development knob: 0 = one CTA per RoI for the headline shape too
    // (bulk copies need 16-byte aligned tiles: a caller's class term at an odd offset takes the one-CTA-per-RoI kernel)
    const bool tiles_aligned = (((uintptr_t)w.yq | (uintptr_t)ys) & 15u) == 0;
    if (one && direct != nullptr && R >= 512 && tiles_aligned && !(er != nullptr && atoi(er) == 0)) {
        int sm_count = 0, devi = 0;
        FGN_CUDA_OK(cudaGetDevice(&devi));
        FGN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, devi));
        const size_t ring_smem = (size_t)(kRingStages + 1) * kMaxPP * kEpiThreads * sizeof(float);
        const char *et = getenv("FGN_EPI_TEAMS");             // development knob: consumer teams per CTA (1 or 2)
        if (et != nullptr && atoi(et) == 1) {
            FGN_CUDA_OK(cudaFuncSetAttribute(relation_epilogue_ring_kernel<kMaxPP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
            relation_epilogue_ring_kernel<kMaxPP, 1><<<sm_count, kEpiThreads + 32, ring_smem, st>>>(
                w.yq, ys, roi_batch, rois5, R, B, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, fc_cls_b, fc_reg_b,
                cls_out, reg_out, raw_cls_out, raw_reg_out);
        } else {
            FGN_CUDA_OK(cudaFuncSetAttribute(relation_epilogue_ring_kernel<kMaxPP, kRingTeams>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
            relation_epilogue_ring_kernel<kMaxPP, kRingTeams><<<sm_count, kRingTeams * kEpiThreads + 32, ring_smem, st>>>(
                w.yq, ys, roi_batch, rois5, R, B, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, fc_cls_b, fc_reg_b,
                cls_out, reg_out, raw_cls_out, raw_reg_out);
        }
        FGN_LAUNCH_OK();
        return FGN_OK;
    }
    if (one)
        relation_epilogue_kernel<kMaxPP, true, true><<<dim3(R, nblk), kEpiThreads, smem, st>>>(
            w.yq, ys, roi_batch, R, B, N, C, cblk, cg, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, w.partial,
            rois5, fc_cls_b, fc_reg_b, direct, reg_out, raw_cls_out, raw_reg_out);
    else if (fast)
        relation_epilogue_kernel<kMaxPP, true><<<dim3(R, nblk), kEpiThreads, smem, st>>>(
            w.yq, ys, roi_batch, R, B, N, C, cblk, cg, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, w.partial,
            rois5, fc_cls_b, fc_reg_b, direct, reg_out, raw_cls_out, raw_reg_out);
    else
        relation_epilogue_kernel<kMaxPP, false><<<dim3(R, nblk), kEpiThreads, smem, st>>>(
            w.yq, ys, roi_batch, R, B, N, C, cblk, cg, gn_eps, gn_w, gn_b, fc_cls_w, fc_reg_w, w.partial,
            rois5, fc_cls_b, fc_reg_b, direct, reg_out, raw_cls_out, raw_reg_out);
    FGN_LAUNCH_OK();
    if (direct == nullptr) {
        relation_finalize_kernel<<<ceil_div(R, 128), 128, 0, st>>>(w.partial, R, N, nblk, fc_cls_b, fc_reg_b,
                                                                  cls_out, reg_out, raw_cls_out, raw_reg_out);
        FGN_LAUNCH_OK();
    }
    return FGN_OK;
}

extern "C" size_t fgn_relation_split_weights_bytes(int C)
{
    return C > 0 ? (size_t)4 * C * C * sizeof(float) : 0;       // { Wq_hi, Wq_lo, Ws_hi, Ws_lo }, each [C,C]
}

extern "C" int fgn_relation_split_weights(const float *conv_w, int C, float *out, void *stream)
{
    FGN_CHECK_ARG(conv_w && out && C > 0, "bad arguments");
    int rc = gemm_split_weights(conv_w, 2 * C, C, C, out, (cudaStream_t)stream);
    if (rc) return rc;
    return gemm_split_weights(conv_w + C, 2 * C, C, C, out + (size_t)2 * C * C, (cudaStream_t)stream);
}

extern "C" int fgn_relation_fusion_fwd(const float *roi_feat, int feat_layout,
                                       const int32_t *roi_batch, const float *spp_cat_mean, const float *class_term, int R,
                                       int B, int N, int C, int P, const float *conv_w, const float *conv_w_split,
                                       const float *conv_b, const float *gn_w, const float *gn_b,
                                       int gn_groups, float gn_eps, const float *fc_cls_w,
                                       const float *fc_cls_b, const float *fc_reg_w,
                                       const float *fc_reg_b, float *cls_out, float *reg_out,
                                       float *raw_cls_out, float *raw_reg_out, int precision,
                                       void *workspace, size_t workspace_bytes, void *stream)
{
    return relation_fusion_impl(roi_feat, feat_layout, roi_batch, nullptr, spp_cat_mean, class_term, R, B, N, C, P, conv_w, conv_w_split,
                                conv_b, gn_w, gn_b, gn_groups, gn_eps, fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b, cls_out,
                                reg_out, raw_cls_out, raw_reg_out, precision, workspace, workspace_bytes, stream);
}

extern "C" int fgn_cls_bbox_reassemble(const float *raw_cls, const float *raw_reg, int R, int N,
                                       float *cls_out, float *reg_out, void *stream)
{
    FGN_CHECK_ARG(R >= 0 && N > 0, "bad dims R=%d N=%d", R, N);
    if (R == 0) return FGN_OK;
    FGN_CHECK_ARG(raw_cls && raw_reg && cls_out && reg_out, "NULL pointer");
    cls_bbox_reassemble_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(raw_cls, raw_reg, R, N,
                                                                                cls_out, reg_out);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

// fp32 [rows, cols] (row pitch ld) -> dense bf16 [rows, cols], round to nearest even
__global__ void f32_to_bf16_kernel(const float *__restrict__ in, int rows, int cols, int ld, uint16_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const float x = in[(size_t)(i / cols) * ld + (i % cols)];
    unsigned u = __float_as_uint(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) { out[i] = (uint16_t)((u >> 16) | 0x40u); return; }   // NaN stays NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    out[i] = (uint16_t)(u >> 16);
}

extern "C" int fgn_roi_align_ml_fwd_bf16(const fgn_pyramid_t *, int, int, const float *, int, int, int, int, float,
                                         const float *, const int32_t *, void *, int, int32_t *, void *);
extern "C" int fgn_gemm_nt_bf16(const uint16_t *, int, const uint16_t *, int, const float *, float *, int, int, int, int, void *);

extern "C" size_t fgn_guided_roi_fused_bf16_workspace_bytes(int R, int BN, int C, int P)
{
    if (R < 0 || BN <= 0 || C <= 0 || P <= 0) return 0;
    const size_t PP = (size_t)P * P;
    return align256((size_t)R * PP * C * 2) + align256((size_t)R * 4) + align256((size_t)BN * PP * C * 2) +
           2 * align256((size_t)C * C * 2) + align256((size_t)R * PP * C * 4) + align256((size_t)BN * PP * C * 4) +
           align256((size_t)R * BN * 6 * ceil_div(C, 32) * 4);
}

// bf16 variant of fgn_guided_roi_fused_fwd: bf16 NHWC pyramid -> bf16 RoI features -> bf16 tcgen05
// contraction (fp32 accumulate) -> the fp32 GroupNorm/ReLU/pool/FC epilogue.  Weights and the class
// maps arrive in fp32 and are rounded to bf16 here.
extern "C" int fgn_guided_roi_fused_fwd_bf16(const fgn_pyramid_t *pyr, int B, int C, const float *rois, int R,
                                             int P, int sampling_ratio, int aligned, float finest_scale,
                                             const float *spp_cat_mean /* [B*N,P,P,C] NHWC fp32 */, int N,
                                             const float *conv_w, const float *conv_b, const float *gn_w,
                                             const float *gn_b, int gn_groups, float gn_eps,
                                             const float *fc_cls_w, const float *fc_cls_b,
                                             const float *fc_reg_w, const float *fc_reg_b, float *cls_out,
                                             float *reg_out, int32_t *lvl_out, void *workspace,
                                             size_t workspace_bytes, void *stream)
{
    FGN_CHECK_ARG(R >= 0 && B > 0 && N > 0 && C > 0 && P > 0, "bad dims");
    FGN_CHECK_ARG(gn_groups > 0 && C % gn_groups == 0, "GroupNorm groups=%d does not divide C=%d", gn_groups, C);
    if (R == 0) return FGN_OK;
    if (P * P != kMaxPP) { set_error("guided_roi_fused_bf16: P=%d not instantiated (7)", P); return FGN_ERR_UNSUPPORTED; }
    const int BN = B * N, PP = P * P;
    const size_t need = fgn_guided_roi_fused_bf16_workspace_bytes(R, BN, C, P);
    if (!workspace || workspace_bytes < need) {
        set_error("guided_roi_fused_bf16: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    uint16_t *feat = (uint16_t *)ws;  ws += align256((size_t)R * PP * C * 2);
    int32_t *rb = (int32_t *)ws;      ws += align256((size_t)R * 4);
    uint16_t *spp16 = (uint16_t *)ws; ws += align256((size_t)BN * PP * C * 2);
    uint16_t *wq16 = (uint16_t *)ws;  ws += align256((size_t)C * C * 2);
    uint16_t *ws16 = (uint16_t *)ws;  ws += align256((size_t)C * C * 2);
    float *yq = (float *)ws;          ws += align256((size_t)R * PP * C * 4);
    float *ys = (float *)ws;          ws += align256((size_t)BN * PP * C * 4);
    float *partial = (float *)ws;

    int rc = fgn_roi_align_ml_fwd_bf16(pyr, B, C, rois, R, P, sampling_ratio, aligned, finest_scale, nullptr, nullptr,
                                       feat, 1, lvl_out, stream);
    if (rc) return rc;
    roi_batch_kernel<<<ceil_div(R, 128), 128, 0, st>>>(rois, R, rb);
    FGN_LAUNCH_OK();
    f32_to_bf16_kernel<<<ceil_div(C * C, 256), 256, 0, st>>>(conv_w, C, C, 2 * C, wq16);
    FGN_LAUNCH_OK();
    f32_to_bf16_kernel<<<ceil_div(C * C, 256), 256, 0, st>>>(conv_w + C, C, C, 2 * C, ws16);
    FGN_LAUNCH_OK();
    f32_to_bf16_kernel<<<ceil_div(BN * PP * C, 256), 256, 0, st>>>(spp_cat_mean, BN * PP, C, C, spp16);
    FGN_LAUNCH_OK();
    rc = fgn_gemm_nt_bf16(feat, C, wq16, C, nullptr, yq, C, R * PP, C, C, stream);
    if (rc) return rc;
    rc = fgn_gemm_nt_bf16(spp16, C, ws16, C, conv_b, ys, C, BN * PP, C, C, stream);
    if (rc) return rc;
    const int cg = C / gn_groups;
    FGN_CHECK_ARG(cg <= kEpiThreads, "channels per group %d > %d", cg, kEpiThreads);
    const int cblk = (kEpiThreads / cg) * cg;
    const int nblk = ceil_div(C, cblk);
    const int nwarps = kEpiThreads / 32;
    const size_t smem = ((size_t)N * 6 * nwarps + kEpiThreads) * 4;
    const bool fast = cblk == kEpiThreads && (C % cblk) == 0 && (cg & (cg - 1)) == 0 && cg <= 32;
    if (fast) FGN_SMEM_OPTIN((relation_epilogue_kernel<kMaxPP, true>), smem);
    else      FGN_SMEM_OPTIN((relation_epilogue_kernel<kMaxPP, false>), smem);
    if (fast)
        relation_epilogue_kernel<kMaxPP, true><<<dim3(R, nblk), kEpiThreads, smem, st>>>(yq, ys, rb, R, B, N, C, cblk, cg, gn_eps, gn_w, gn_b,
                                                                                        fc_cls_w, fc_reg_w, partial,
                                                                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
        relation_epilogue_kernel<kMaxPP, false><<<dim3(R, nblk), kEpiThreads, smem, st>>>(yq, ys, rb, R, B, N, C, cblk, cg, gn_eps, gn_w, gn_b,
                                                                                         fc_cls_w, fc_reg_w, partial,
                                                                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    FGN_LAUNCH_OK();
    relation_finalize_kernel<<<ceil_div(R, 128), 128, 0, st>>>(partial, R, N, nblk, fc_cls_b, fc_reg_b, cls_out, reg_out,
                                                              nullptr, nullptr);
    FGN_LAUNCH_OK();
    return FGN_OK;
}

extern "C" size_t fgn_guided_roi_fused_workspace_bytes(int R, int BN, int C, int P)
{
    if (R < 0 || BN <= 0 || C <= 0 || P <= 0) return 0;
    return align256((size_t)R * P * P * C * 4) + align256((size_t)R * 4) +
           fgn_relation_fusion_workspace_bytes(R, BN, C, P);
}


extern "C" int fgn_guided_roi_fused_fwd(const fgn_pyramid_t *pyr, int B, int C, const float *rois,
                                        int R, int P, int sampling_ratio, int aligned,
                                        float finest_scale, const float *spp_cat_mean, const float *class_term, int N,
                                        const float *conv_w, const float *conv_w_split, const float *conv_b,
                                        const float *gn_w, const float *gn_b, int gn_groups,
                                        float gn_eps, const float *fc_cls_w, const float *fc_cls_b,
                                        const float *fc_reg_w, const float *fc_reg_b,
                                        float *cls_out, float *reg_out, int32_t *lvl_out,
                                        int precision, void *workspace, size_t workspace_bytes,
                                        void *stream)
{
    FGN_CHECK_ARG(R >= 0 && B > 0 && N > 0 && C > 0 && P > 0, "bad dims");
    if (R == 0) return FGN_OK;
    const size_t need = fgn_guided_roi_fused_workspace_bytes(R, B * N, C, P);
    if (!workspace || workspace_bytes < need) {
        set_error("guided_roi_fused: workspace %zu B < required %zu B", workspace_bytes, need);
        return FGN_ERR_WORKSPACE;
    }
    char *ws = (char *)workspace;
    float *feat = (float *)ws;                ws += align256((size_t)R * P * P * C * 4);
    int32_t *rb = (int32_t *)ws;              ws += align256((size_t)R * 4);
    int rc = fgn_roi_align_ml_fwd(pyr, B, C, FGN_LAYOUT_NHWC, rois, R, P, sampling_ratio, aligned,
                                  finest_scale, nullptr, nullptr, feat, FGN_LAYOUT_NHWC, lvl_out, stream);
    if (rc) return rc;
    (void)rb;                                 // the epilogue reads the batch index from rois[:,0] itself
    return relation_fusion_impl(feat, FGN_LAYOUT_NHWC, nullptr, rois, spp_cat_mean, class_term, R, B, N, C, P, conv_w, conv_w_split,
                                conv_b, gn_w, gn_b, gn_groups, gn_eps, fc_cls_w, fc_cls_b,
                                fc_reg_w, fc_reg_b, cls_out, reg_out, nullptr, nullptr, precision,
                                ws, workspace_bytes - (size_t)(ws - (char *)workspace), stream);
}

// roi_align_window.cu -- persistent, planner-fed, rotating-window multi-level RoIAlign (sm_100a).
//
// Same separable formulation and exact coordinate arithmetic as roi_align_stream.cu
//   out[ph,pw] = 1/count * sum_y Ay[ph][y] * sum_x Ax[pw][x] * v[y,x]
// reorganised so that the SMs spend their issue slots on the row pass and nothing else:
//
//   * PERSISTENT CTAs (2 per SM at P=7) pull work items from a ticket counter.  An item is
//     (RoI, range of bin rows, channel block).  With one channel block the tickets are handed out
//     largest-footprint-class first (the warps that idle during the first plan classify all RoIs into
//     ballot masks; every CTA derives the same order), big RoIs as 2 or 4 items that different CTAs
//     take, and CTA i starts on RoI i without a ticket when it is small.
//   * a PLANNER warp works ahead of everyone else: it takes the ticket, assigns the FPN level,
//     evaluates the reference coordinate arithmetic once per (axis, bin, sample) and leaves the
//     footprint, ring schedule and weight tables in one of three shared-memory plan slots.  The
//     per-RoI prologue is therefore off the critical path of the copy and the math.
//   * a PRODUCER warp streams the footprint rows NHWC -> shared-memory ring with bulk async copies
//     (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP); the ring keeps running across items.
//   * P CONSUMER warps (warp = bin column) do the row pass out of shared memory with 128-bit loads
//     at compile-time offsets, then fold the row into a WINDOW of kWin bin rows held in registers.
//     Footprint rows are visited top to bottom and the set of bin rows a footprint row touches is a
//     contiguous range whose first member never decreases, so the window only ever slides down:
//     when the row passes the last row of the window's first bin, that bin is scaled by 1/count
//     (and the AG-FCN channel attention) and stored, and the window rotates.  The accumulators of
//     one RoI shrink from P to kWin rows of registers, the fold loses its P compare-and-branch
//     pairs, and the output stores are spread over the item instead of bursting at its end.
//   * a footprint row can touch more than kWin bin rows only when a bin is shorter than 2/3 of a
//     cell ((dph - 1 + 1/g) * bin_h < 2); such RoIs are planned as ceil(P / kWin) items of kWin
//     bin rows each, for which the bound holds trivially.
// What bounds it on B200 (measurements in DESIGN.md section 5): the ring (bytes in flight / copy latency),
// the planner and the consumers' per-row instruction chains all sit at 40-48 us per 1000 cfg3 RoIs.
#include "common.cuh"
#include <stdlib.h>
#include <atomic>
#include <type_traits>
#include <cuda_bf16.h>

namespace fgn {

namespace {

constexpr int kWin        = 4;      // bin rows held in registers per consumer warp
constexpr int kStageCells = 32;     // cells (of CB channels) per ring stage
constexpr unsigned kTicketSlots = 65536;   // ticket counters: [0, half) for launches baked into CUDA graphs, [half, end) eager

__device__ unsigned int g_window_ticket[kTicketSlots];
// Slot allocators shared by EVERY instantiation of the launcher (P=7 and P=14 launches in flight at the same time on
// different streams must never share a counter) and safe against concurrent host threads.
std::atomic<unsigned int> g_eager_seq{0}, g_captured_seq{0};
__device__ unsigned int g_window_violation;           // planner self-check (must stay 0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int N> struct IntTag { static constexpr int value = N; };

// packed fp32 FMA / MUL (Blackwell FFMA2 / FMUL2: two IEEE fp32 results per lane per issue slot)
__device__ __forceinline__ void fma4x2(float4 &a, const float w, const float4 v)
{
    unsigned long long a0, a1, v0, v1, ww;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(a.z), "f"(a.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v0) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v1) : "f"(v.z), "f"(v.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(ww), "l"(v0));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(ww), "l"(v1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.z), "=f"(a.w) : "l"(a1));
}
__device__ __forceinline__ float4 mul4x2(const float w, const float4 v)
{
    unsigned long long r0, r1, v0, v1, ww;
    float4 o;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v0) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v1) : "f"(v.z), "f"(v.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r0) : "l"(ww), "l"(v0));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r1) : "l"(ww), "l"(v1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.z), "=f"(o.w) : "l"(r1));
    return o;
}

constexpr int kSortCap = 2048;      // most RoIs per launch the sorted ticket scheme handles
constexpr int kPlanSlots = 3;       // plans in flight per CTA (the planner runs up to two items ahead)

// mbarrier ops on precomputed 32-bit shared addresses (the generic->shared conversion is done once per kernel)
__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive32(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx32(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s32(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int P>
struct WinSlot {
    int   r, cb0, level, batch, H, W;
    int   pa, pb;                                  // bin rows [pa, pb) of this item
    float count;
    int   X0, Y0, ncols, nrows, nseg, rps, nstages;
    int   xlo[P], xn[P], xoff[P];
    int   hi[P + 1];                               // last footprint row (relative to Y0) of bins <= ph
};

// development trace (debug_mode bit 5): per CTA, per item (first 16), 12 globaltimer stamps
constexpr int kTraceItems = 16, kTraceEvents = 12, kTraceCtas = 296;
__device__ unsigned long long g_window_trace[kTraceCtas * kTraceItems * kTraceEvents];
__device__ __forceinline__ void trace(int debug_mode, int item, int ev, int lane)
{
    if ((debug_mode & 32) && lane == 0 && item < kTraceItems && blockIdx.x < kTraceCtas) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_window_trace[((size_t)blockIdx.x * kTraceItems + item) * kTraceEvents + ev] = t;
    }
}

}  // namespace

// bf16 cells -> fp32 (exact: a bf16 is the upper half of an fp32); half 0 = channels 0..3 of the lane's 8, half 1 = 4..7
__device__ __forceinline__ float4 bf16x4_lo(const uint4 q)
{
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u),
                       __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
}
__device__ __forceinline__ float4 bf16x4_hi(const uint4 q)
{
    return make_float4(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u),
                       __uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
}
__device__ __forceinline__ float4 bf16x4(const uint2 q)
{
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u),
                       __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
}

// Warps: 0..P-1 consumers (warp = bin column), P = producer, P+1 = planner.
// Dynamic shared memory: ring[NS][SC*CB] | kPlanSlots x { wx[wx_cap] | wrow[wyd_rows] float4 }
// H16: the pyramid holds bf16 (NHWC, uint16 storage): a cell moves half the bytes, a stage holds twice the cells, a lane
// owns 8 contiguous channels (one 128-bit LDS per cell); weights, accumulation and scaling stay fp32.  O16: bf16 output
// (round to nearest even), H16 kernels only.
template <int P, int VEC, int NS, int MINB, bool SCALED, bool H16 = false, bool O16 = false>
__global__ void __launch_bounds__((P + 2) * 32, MINB)
roi_align_window_kernel(const Pyramid pyr, const int C, const float *__restrict__ rois, const int R,
                        const int sampling_ratio, const int aligned, const float finest_scale,
                        const float *__restrict__ chan_scale, const int32_t *__restrict__ scale_index,
                        void *__restrict__ out_v, int32_t *__restrict__ lvl_out,
                        const int wx_cap, const int wyd_rows, const int ticket_slot, const float split_cells,
                        const float3 thr, const int debug_mode)
{
    // debug_mode (development only): bit 0 = no copies (producer only signals), bit 1 = no row math,
    // bit 5 = record the per-CTA timeline
    constexpr int CB = 128 * VEC;                   // channels per item; cell stride in the ring
    constexpr int S  = (P + kWin - 1) / kWin;       // bin-row chunks of a split RoI
    static_assert(!H16 || VEC == 2, "bf16 cells: a lane owns 8 channels of a 256-channel block");
    static_assert(!O16 || H16, "bf16 output comes with bf16 input");
    typedef typename std::conditional<H16, unsigned short, float>::type in_t;
    constexpr int ES = (int)sizeof(in_t);
    constexpr int SC = kStageCells * (4 / ES);      // cells per ring stage (a stage is 32 fp32 cells' worth of bytes)
    constexpr int kStageFloats = SC * CB;           // elements per stage
    constexpr int V1 = H16 ? 4 : 128;               // channel offset of a lane's second float4
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ WinSlot<P> slot[kPlanSlots];
    __shared__ __align__(8) uint64_t bars[2 * NS + 2 * kPlanSlots];
    uint64_t *const full_bar = bars, *const empty_bar = bars + NS, *const plan_full = bars + 2 * NS,
                   *const plan_empty = bars + 2 * NS + kPlanSlots;
    __shared__ unsigned int cls_mask[4][kSortCap / 32];   // size-class membership of every RoI (sorted ticket scheme)

    in_t *ring = reinterpret_cast<in_t *>(smem_raw);
    float *wtab = reinterpret_cast<float *>(smem_raw + (size_t)NS * kStageFloats * ES);
    const int wslot = wx_cap + 4 * wyd_rows;        // floats per plan slot
    // one opaque register holds the barriers' shared address: left to itself ptxas re-derives it before every wait
    // (S2UR SR_CgaCtaId + ULEA, ~40 cycles of latency per ring stage)
    uint32_t bar32;
    asm volatile("mov.u32 %0, %1;" : "=r"(bar32) : "r"(smem_u32(bars)));
    const uint32_t full32 = bar32, empty32 = bar32 + 8 * NS;
    const uint32_t pfull32 = bar32 + 16 * NS, pempty32 = bar32 + 16 * NS + 8 * kPlanSlots;

    const int nblk  = (C + CB - 1) / CB;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], P); }
        for (int b = 0; b < kPlanSlots; ++b) { mbar_init(&plan_full[b], 1); mbar_init(&plan_empty[b], P + 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Sorted ticket scheme (see the planner): while the planner works on the CTA's first RoI, the other warps
    // classify every RoI by footprint size -- 32 RoIs per warp step, one ballot mask per class and block.
    const bool sorted = (nblk == 1) && (R <= kSortCap) && !(debug_mode & 64);
    const int nstatic = sorted ? min(R, (int)gridDim.x) : 0;             // CTAs past R start on a ticket (small launches)
    const int nsb = (R + 31) >> 5;                                       // 32-RoI blocks
    // Plain tickets are (RoI, bin-row chunk, channel block), unused chunks skipped, so that different CTAs can take the
    // chunks of a big RoI and the launch's tail stays short.  A launch with many items per CTA has no tail to balance and
    // a skipped ticket is a planner round trip (geometry + an exposed atomic, ~2 us of a 7 us plan period at 12 000
    // RoIs): there a ticket is (RoI, channel block) and the planner walks the chunks of the few RoIs that need them.
    // (32 items per CTA: cfg5's 8192-RoI launch, 28 per CTA and HBM-bound, still gains 2.6 % from the chunk tickets)
    const bool roi_tickets = !sorted && (long long)R * nblk >= 32ll * gridDim.x;
    const int items = roi_tickets ? R * nblk : R * S * nblk;
    auto size_class = [&](const float *roi) {                            // 0 = largest footprints ... 3 = smallest (NaN -> 3)
        const int lv = roi_level(roi, pyr, finest_scale);
        const RoiGeom gg = roi_geometry(roi, pyr.scale[lv], P, sampling_ratio, aligned, pyr.B);
        const float est = (gg.bin_h * (float)P + 2.f) * (gg.bin_w * (float)P + 2.f);   // cells, from the box alone
        return est > thr.x ? 0 : (est > thr.y ? 1 : (est > thr.z ? 2 : 3));
    };
    if (sorted && warp <= P) {
        for (int blk = warp; blk < nsb; blk += P + 1) {
            const int idx = blk * 32 + lane;
            int c = -1;
            if (idx < R) {
                c = size_class(rois + 5 * (size_t)idx);
                if (idx < nstatic && c >= 2) c = -1;                     // CTA idx starts on it without a ticket
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned m = __ballot_sync(FULL, c == q);
                if (lane == q) cls_mask[q][blk] = m;
            }
        }
        // announce the masks to the planner without waiting for it (it reads them before its first lookup;
        // blocking here could deadlock: a first RoI split into more items than plan slots needs the consumers)
        __threadfence_block();
        asm volatile("bar.arrive 1, %0;" ::"n"((P + 2) * 32) : "memory");
    }

    if (warp == P + 1) {
        // ===== planner ===================================================================================
        // Two ticket schemes.
        //  * sorted (one channel block, R <= kSortCap): CTA i starts on RoI i with no ticket at all; while its
        //    consumers work on that, the planner classifies every other RoI by footprint size into four classes
        //    (ballot masks in shared memory, identical in every CTA because every CTA computes them from the
        //    same rois) and ticket t then means "the t-th RoI in largest-class-first order".  Big footprints
        //    start early, the launch ends on small ones, and nobody has to be split to balance the tail.
        //  * plain (otherwise): ticket = (RoI, bin-row chunk, channel block) in index order; chunks a RoI does
        //    not use are skipped.
        // The planner always holds one prefetched ticket, so a launch draws ntickets + 2 per CTA in all; the
        // last one resets the counter for the next launch that uses this slot.
        constexpr int S0 = (P + 1) / 2, S1 = (P + 3) / 4;                // chunks (of 2 / 4 bin rows) of a class 0 / 1 RoI
        int ntickets = sorted ? 0x3fffffff : items;                      // sorted: known once the class totals are
        unsigned last_ticket = sorted ? 0xffffffffu : (unsigned)items + 2u * gridDim.x - 1u;
        auto take = [&]() {
            unsigned int tk = 0;
            if (lane == 0) {
                tk = atomicAdd(&g_window_ticket[ticket_slot], 1u);
                if (tk == last_ticket) g_window_ticket[ticket_slot] = 0u;
            }
            return tk;                                                   // valid in lane 0; broadcast at use
        };
        int cls_total = 0;                                               // lane c < 4: RoIs in size class c
        auto class_totals = [&]() {                                      // once the other warps have built the masks
            asm volatile("bar.sync 1, %0;" ::"n"((P + 2) * 32) : "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int n = 0;
                for (int i = lane; i < nsb; i += 32) n += __popc(cls_mask[q][i]);
                n = __reduce_add_sync(FULL, n);
                if (lane == q) cls_total = n;
            }
            ntickets = __shfl_sync(FULL, cls_total, 0) * S0 + __shfl_sync(FULL, cls_total, 1) * S1
                       + __shfl_sync(FULL, cls_total, 2) + __shfl_sync(FULL, cls_total, 3);
            // (the one ticket this planner drew before it knew the totals is < 2 * gridDim.x - 1 <= last_ticket,
            //  and the draw that returns exactly last_ticket comes after every planner's first, so none is missed)
            last_ticket = (unsigned)ntickets + 2u * gridDim.x - 1u;
        };
        auto lookup = [&](int tk, int &ba, int &bb) {                    // ticket -> (RoI, bin rows [ba, bb)), largest class first
            const int t0 = __shfl_sync(FULL, cls_total, 0), t1 = __shfl_sync(FULL, cls_total, 1),
                      t2 = __shfl_sync(FULL, cls_total, 2);
            int c, krem, chunk = 0, cw = P;
            if (tk < t0 * S0)                     { c = 0; krem = tk / S0; chunk = tk - krem * S0; cw = 2; }
            else if ((tk -= t0 * S0) < t1 * S1)   { c = 1; krem = tk / S1; chunk = tk - krem * S1; cw = 4; }
            else if ((tk -= t1 * S1) < t2)        { c = 2; krem = tk; }
            else                                  { c = 3; krem = tk - t2; }
            ba = chunk * cw; bb = min(P, ba + cw);
            int r = 0;
            for (int b0 = 0; b0 < nsb; b0 += 32) {                       // 32 blocks per round
                const unsigned m = (b0 + lane < nsb) ? cls_mask[c][b0 + lane] : 0u;
                const int cnt = __popc(m);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += up;
                }
                const unsigned hit = __ballot_sync(FULL, incl > krem);
                if (hit != 0u) {
                    const int src = __ffs(hit) - 1;
                    const unsigned mm = __shfl_sync(FULL, m, src);
                    const int before = __shfl_sync(FULL, incl - cnt, src);
                    r = (b0 + src) * 32 + (int)__fns(mm, 0, krem - before + 1);
                    break;
                }
                krem -= __shfl_sync(FULL, incl, 31);
            }
            return r;
        };

        unsigned int tnext = take();
        int k = 0, b = 0;                                                // items published so far, their slot
        unsigned eparity = 1;                                            // plan_empty: item k - kPlanSlots consumed
        bool totals_known = false;
        for (unsigned int it = 0;; ++it) {
            // the ticket's RoI, channel block and bin rows [ba, bb); window chunks [chunk_lo, chunk_hi) of them
            int r = 0, cbi = 0, ba = 0, bb = P, chunk_lo = 0, chunk_hi = S;
            bool have = false;
            if (sorted && it == 0 && (int)blockIdx.x < nstatic) {        // small first RoI: start on it right away
                r = (int)blockIdx.x;
                have = size_class(rois + 5 * (size_t)r) >= 2;
            }
            if (!have) {
                const unsigned int ticket = __shfl_sync(FULL, tnext, 0);
                if (sorted && !totals_known) {
                    class_totals();
                    totals_known = true;
                    // `ticket` was drawn before last_ticket was known (a CTA that starts late can draw the last one)
                    if (lane == 0 && ticket == last_ticket) g_window_ticket[ticket_slot] = 0u;
                }
                tnext = take();                                          // in flight while this item is planned
                if ((int)ticket >= ntickets) {
                    mbar_wait32(pempty32 + 8 * b, eparity);
                    if (lane == 0) slot[b].r = -1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive32(pfull32 + 8 * b);
                    break;
                }
                if (sorted) r = lookup((int)ticket, ba, bb);
                else if (roi_tickets) { cbi = (int)ticket % nblk; r = (int)ticket / nblk; }
                else {
                    cbi = (int)ticket % nblk; chunk_lo = ((int)ticket / nblk) % S; chunk_hi = chunk_lo + 1;
                    r = (int)ticket / (nblk * S);
                }
            }
            trace(debug_mode, k, 0, lane);
            const float *roi = rois + 5 * (size_t)r;
            const int level = roi_level(roi, pyr, finest_scale);
            const RoiGeom g = roi_geometry(roi, pyr.scale[level], P, sampling_ratio, aligned, pyr.B);
            const int H = pyr.H[level], W = pyr.W[level];
            // bins shorter than 3/4 cell can put > kWin bin rows on one footprint row: such RoIs are planned
            // as items of kWin bin rows each (plain scheme: big footprints too, to shorten the launch's tail)
            const float est = (g.bin_h * (float)P + 2.f) * (g.bin_w * (float)P + 2.f);
            const bool split = S > 1 && (g.bin_h < 0.75f || (!sorted && !roi_tickets && est > split_cells));
            if (!split) { if (chunk_lo > 0) continue; chunk_hi = 1; }
            else if (sorted) chunk_hi = (bb - ba + kWin - 1) / kWin;
            if (lane == 0 && lvl_out != nullptr && cbi == 0 && chunk_lo == 0 && ba == 0) lvl_out[r] = level;

            for (int chunk = chunk_lo; chunk < chunk_hi; ++chunk) {
            WinSlot<P> &ps = slot[b];
            const int pa = split ? ba + chunk * kWin : ba, pb = split ? min(bb, pa + kWin) : bb;

            // per-lane bin: lanes [0,P) = bin rows, [P,2P) = bin columns.  Sample coordinates are monotone in
            // the sample index, so when the first and last sample of a bin are valid they bound its cells.
            const int axis = lane >= P ? 1 : 0, p = lane - axis * P;
            const bool isx = lane >= P && lane < 2 * P, isy = lane < P && p >= pa && p < pb;
            const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
            const int grid = axis ? g.grid_w : g.grid_h, size = axis ? W : H;
            int lo = 0x7fffffff, hi = -1;
            if ((isx || isy) && grid > 0) {
                const AxisSample s0 = axis_sample(start, bin, grid, size, p, 0);
                const AxisSample s1 = axis_sample(start, bin, grid, size, p, grid - 1);
                if (s0.valid && s1.valid) { lo = min(s0.low, s1.low); hi = max(s0.high, s1.high); }   // either direction (x2 < x1)
                else {
                    for (int i = 0; i < grid; ++i) {
                        const AxisSample sm = axis_sample(start, bin, grid, size, p, i);
                        if (sm.valid) { lo = min(lo, sm.low); hi = max(hi, sm.high); }
                    }
                }
            }
            int n = hi >= 0 ? hi - lo + 1 : 0;
            if (hi < 0) lo = 0;
            const int big = 0x7fffffff;
            int X0 = __reduce_min_sync(FULL, (isx && n > 0) ? lo : big);
            int X1 = __reduce_max_sync(FULL, (isx && n > 0) ? lo + n : -1);
            int Y0 = __reduce_min_sync(FULL, (isy && n > 0) ? lo : big);
            int Y1 = __reduce_max_sync(FULL, (isy && n > 0) ? lo + n : -1);
            const int n4 = (n + 3) & ~3;                                 // weight runs start 16 B aligned
            const int xsum = __reduce_add_sync(FULL, isx ? n4 : 0);
            if (X1 < 0 || Y1 < 0 || xsum + 8 > wx_cap || (Y1 - Y0) > wyd_rows) { X0 = X1 = Y0 = Y1 = 0; n = 0; }
            const int ncols = X1 - X0, nrows = Y1 - Y0;
            int nseg, rps, nstages;
            if (ncols <= SC) {
                nseg = 1; rps = ncols > 0 ? min(32, SC / ncols) : 1; nstages = (nrows + rps - 1) / rps;   // a producer lane per row
            } else {
                nseg = (ncols + SC - 1) / SC; rps = 1; nstages = nrows * nseg;
            }
            if (ncols == 0 || nrows == 0) nstages = 0;
            int off = 0;                                                 // exclusive scan of the padded x runs
            int hiall[P];                                                // running max of the bin rows' last footprint row
            int him = -1, myhi = -1;
#pragma unroll
            for (int qq = 0; qq < P; ++qq) {
                const int nq = __shfl_sync(FULL, n4, P + qq);
                if (isx && qq < p) off += nq;
                const int hq = __shfl_sync(FULL, (isy && n > 0) ? lo + n - 1 - Y0 : -1, qq);
                him = max(him, hq);
                hiall[qq] = him;
                if (qq == lane) myhi = him;
            }
            float *wx = wtab + (size_t)b * wslot;
            float *wrow = wx + wx_cap;                                   // [nrows][4]
            trace(debug_mode, k, 1, lane);
            mbar_wait32(pempty32 + 8 * b, eparity);
            trace(debug_mode, k, 2, lane);
            if (lane == 0) {
                ps.r = r; ps.cb0 = cbi * CB; ps.level = level; ps.batch = g.batch; ps.H = H; ps.W = W;
                ps.pa = pa; ps.pb = pb; ps.count = g.count;
                ps.X0 = X0; ps.Y0 = Y0; ps.ncols = ncols; ps.nrows = nrows;
                ps.nseg = nseg; ps.rps = rps; ps.nstages = nstages;
            }
            if (isx) { ps.xlo[p] = lo; ps.xn[p] = n; ps.xoff[p] = off; }
            if (lane < P) ps.hi[lane] = myhi;
            for (int i = lane; i < nrows; i += 32)
                reinterpret_cast<float4 *>(wrow)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = lane; i < (xsum >> 2); i += 32)
                reinterpret_cast<float4 *>(wx)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            if (nstages > 0 && n > 0) {
                // Both axes run ONE instruction stream: a sample adds its two bilinear weights to two entries of
                // the slot's table.  x: entry = run offset + cell.  y: footprint row j lives in window slot
                // (p - base_j) of wrow[j], base_j = first bin row of the item whose (running-max) last row is >= j.
                auto entry = [&](int cell) {
                    if (isx) return off + cell - lo;
                    const int j = cell - Y0;
                    int base = pa;
#pragma unroll
                    for (int qq = 0; qq < P; ++qq) base += (qq >= pa && qq < pb && hiall[qq] < j) ? 1 : 0;
                    const int comp = p - base;
                    if (comp < 0 || comp >= kWin) { atomicAdd(&g_window_violation, 1u); return -1; }
                    return wx_cap + 4 * j + comp;
                };
                for (int i = 0; i < grid; ++i) {
                    const AxisSample sm = axis_sample(start, bin, grid, size, p, i);
                    if (sm.valid) {
                        const int e0 = entry(sm.low), e1 = entry(sm.high);
                        if (e0 >= 0) wx[e0] += sm.h;
                        if (e1 >= 0) wx[e1] += sm.l;
                    }
                }
            }
            __syncwarp();
            trace(debug_mode, k, 3, lane);
            if (lane == 0) mbar_arrive32(pfull32 + 8 * b);
            ++k;
            if (++b == kPlanSlots) { b = 0; eparity ^= 1; }
            }   // chunks of this RoI
        }
    } else if (warp == P) {
        // ===== producer: bulk async copies of footprint rows, ring runs across items =======================
        const uint32_t ring32 = smem_u32(ring);
        int s = 0;
        unsigned par = 1;                                                // parity 1: first pass over a fresh barrier
        int b = 0;
        unsigned fparity = 0;
        for (int k = 0;; ++k) {
            mbar_wait32(pfull32 + 8 * b, fparity);
            const WinSlot<P> &ps = slot[b];
            if (ps.r < 0) break;
            const int cbn = min(CB, C - ps.cb0);
            const bool rowcopy = (C == CB);                              // a row segment is one contiguous run
            const int nrows = ps.nrows, ncols = ps.ncols, nstages = ps.nstages, rps = ps.rps, nseg = ps.nseg;
            const in_t *fbase = reinterpret_cast<const in_t *>(pyr.feat[ps.level]) + ((size_t)ps.batch * ps.H * ps.W) * C + ps.cb0
                                + ((size_t)ps.Y0 * ps.W + ps.X0) * C;
            const size_t row_pitch = (size_t)ps.W * C;
            __syncwarp();
            if (lane == 0) mbar_arrive32(pempty32 + 8 * b);              // everything needed is in registers now
            if (++b == kPlanSlots) { b = 0; fparity ^= 1; }
            trace(debug_mode, k, 6, lane);
            if (rowcopy && nseg == 1 && !(debug_mode & 1)) {
                // Common case, kept short because this warp's latency per stage bounds the whole pipeline: a stage
                // is rps whole rows; lane i owns row i of every stage, so its source and destination only advance
                // by constants.  (No ordering is needed between lane 0's expect_tx and the other lanes' copies:
                // the phase cannot complete before the expect_tx arrival.)
                const uint32_t row_bytes = (uint32_t)ncols * (CB * ES);
                const in_t *src = fbase + (size_t)lane * row_pitch;
                const size_t src_step = (size_t)rps * row_pitch;
                const uint32_t dst_lane = ring32 + (uint32_t)lane * row_bytes;
                int rows_left = nrows;
                for (int st = 0; st < nstages; ++st) {
                    const int nr = min(rps, rows_left);
                    rows_left -= nr;
                    mbar_wait32(empty32 + 8 * s, par);
                    const uint32_t fb = full32 + 8 * s;
                    if (lane == 0) mbar_expect_tx32(fb, (uint32_t)nr * row_bytes);
                    if (lane < nr) bulk_g2s32(dst_lane + (uint32_t)s * (kStageFloats * ES), src, row_bytes, fb);
                    src += src_step;
                    if (++s == NS) { s = 0; par ^= 1; }
                }
                trace(debug_mode, k, 7, lane);
                continue;
            }
            int row0 = 0, col0 = 0;
            for (int st = 0; st < nstages; ++st) {
                int nr, nc;
                if (nseg == 1) { nr = min(rps, nrows - row0); nc = ncols; }
                else           { nr = 1; nc = min(SC, ncols - col0); }
                mbar_wait32(empty32 + 8 * s, par);
                const uint32_t dst = ring32 + (uint32_t)s * (kStageFloats * ES), fb = full32 + 8 * s;
                if (debug_mode & 1) {
                    if (lane == 0) mbar_arrive32(fb);
                } else {
                    if (lane == 0) mbar_expect_tx32(fb, (uint32_t)(nr * nc * cbn * ES));
                    __syncwarp();
                    const in_t *src = fbase + (size_t)row0 * row_pitch + (size_t)col0 * C;
                    if (rowcopy) {
                        if (lane < nr)
                            bulk_g2s32(dst + (uint32_t)(lane * nc) * (CB * ES), src + (size_t)lane * row_pitch,
                                       (uint32_t)(nc * CB * ES), fb);
                    } else {
                        for (int cell = lane; cell < nr * nc; cell += 32) {
                            const int rr = cell / nc, cc = cell - rr * nc;
                            bulk_g2s32(dst + (uint32_t)cell * (CB * ES), src + (size_t)rr * row_pitch + (size_t)cc * C,
                                       (uint32_t)(cbn * ES), fb);
                        }
                    }
                }
                if (nseg == 1) row0 += nr;
                else { col0 += nc; if (col0 >= ncols) { col0 = 0; ++row0; } }
                if (++s == NS) { s = 0; par ^= 1; }
            }
            trace(debug_mode, k, 7, lane);
        }
    } else {
        // ===== consumers: warp = bin column pw ==============================================================
        const int pw = warp;
        const int lch = H16 ? lane * 8 : lane * 4;       // lane -> channels [4*lane, 4*lane+4) + 128*v of the block (bf16 cells: 8 contiguous)
        int s = 0;
        unsigned par = 0;
        int b = 0;
        unsigned fparity = 0;
        for (int k = 0;; ++k) {
            mbar_wait32(pfull32 + 8 * b, fparity);
            const WinSlot<P> &ps = slot[b];
            const int r = ps.r;
            if (r < 0) break;
            if (pw == 0) trace(debug_mode, k, 8, lane);
            const float *wx = wtab + (size_t)b * wslot;
            const float4 *wr = reinterpret_cast<const float4 *>(wx + wx_cap);   // y weights of the next footprint row
            const int cb0 = ps.cb0, cbn = min(CB, C - cb0);
            const int nrows = ps.nrows, ncols = ps.ncols, nstages = ps.nstages, rps = ps.rps, nseg = ps.nseg;
            const int xlo = ps.xlo[pw] - ps.X0, nx = ps.xn[pw];
            const float *wxp = wx + ps.xoff[pw];
            const float inv = 1.0f / ps.count;          // count is a small exact integer; <= 1 ulp vs acc/count
            const int *hip = &ps.hi[ps.pa];              // last footprint row of the window's first bin, next bins
            int bins_left = ps.pb - ps.pa;               // bin rows of the item not stored yet
            int rows_left = *hip + 1;                    // footprint rows before the window's first bin is complete
            int prev_hi = *hip;
            // Lanes past the channel count of a ragged last block read stale ring bytes (the cell stride is CB)
            // and never store.
            float4 cs[SCALED ? VEC : 1];                 // AG-FCN channel attention of this RoI (SCALED kernels only)
#pragma unroll
            for (int v = 0; v < (SCALED ? VEC : 1); ++v) cs[v] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (SCALED) {
                const int si = scale_index != nullptr ? scale_index[r] : r;
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (v * V1 + lch < cbn) cs[v] = ldg4(chan_scale + (size_t)si * C + cb0 + v * V1 + lch);
            }
            const size_t o0 = ((size_t)(r * P + ps.pa) * P + pw) * C + cb0 + lch;     // output of the window's first bin
            float *op = reinterpret_cast<float *>(out_v) + o0;
            unsigned short *op16 = reinterpret_cast<unsigned short *>(out_v) + o0;
            // the bin column's weights live in registers for the whole item (runs are padded to 4 floats and
            // the table has 8 floats of slack, so reading 8 is always in bounds; entries past nx are unused)
            float wreg[8];
            {
                const float4 w0 = *reinterpret_cast<const float4 *>(wxp), w1 = *reinterpret_cast<const float4 *>(wxp + 4);
                wreg[0] = w0.x; wreg[1] = w0.y; wreg[2] = w0.z; wreg[3] = w0.w;
                wreg[4] = w1.x; wreg[5] = w1.y; wreg[6] = w1.z; wreg[7] = w1.w;
            }
            __syncwarp();
            // (the slot's header is in registers; its tables are read until the last row)

            float4 a[kWin][VEC], racc[VEC];
#pragma unroll
            for (int qq = 0; qq < kWin; ++qq)
#pragma unroll
                for (int v = 0; v < VEC; ++v) a[qq][v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

            // store the window's first bin row and slide the window down by one
            auto rotate = [&]() {
                unsigned o16[2 * VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float4 qv = a[0][v];
                    float4 o;
                    if (SCALED) {                       // (acc * 1/count) * vec, same rounding order as unfused
                        const float4 c4 = cs[SCALED ? v : 0];
                        o = make_float4(qv.x * inv * c4.x, qv.y * inv * c4.y, qv.z * inv * c4.z, qv.w * inv * c4.w);
                    }
                    else
                        o = make_float4(qv.x * inv, qv.y * inv, qv.z * inv, qv.w * inv);
                    if (O16) {                          // round to nearest even; the lane's 8 channels leave as one 128-bit store
                        const __nv_bfloat162 t0 = __floats2bfloat162_rn(o.x, o.y), t1 = __floats2bfloat162_rn(o.z, o.w);
                        o16[2 * v] = *reinterpret_cast<const unsigned *>(&t0);
                        o16[2 * v + 1] = *reinterpret_cast<const unsigned *>(&t1);
                    }
                    else if (v * V1 + lch < cbn) *reinterpret_cast<float4 *>(op + v * V1) = o;
#pragma unroll
                    for (int w = 0; w + 1 < kWin; ++w) a[w][v] = a[w + 1][v];
                    a[kWin - 1][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (O16 && lch < cbn) *reinterpret_cast<uint4 *>(op16) = make_uint4(o16[0], o16[1], o16[2 * VEC - 2], o16[2 * VEC - 1]);
                op += (size_t)P * C;
                op16 += (size_t)P * C;
                --bins_left;
            };
            // the footprint row in racc is complete: fold it into the window
            auto fold = [&]() {
                while (rows_left <= 0) {                 // the row lies past the window's first bin: store it, slide
                    rotate();
                    ++hip;
                    const int h = bins_left > 0 ? *hip : 0x3fffffff;
                    rows_left += h - prev_hi;
                    prev_hi = h;
                }
                --rows_left;
                const float4 w4 = *wr++;
#pragma unroll
                for (int v = 0; v < VEC; ++v) fma4x2(a[0][v], w4.x, racc[v]);
                if (w4.y != 0.f) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) fma4x2(a[1][v], w4.y, racc[v]);
                }
                if (w4.z != 0.f || w4.w != 0.f) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { fma4x2(a[2][v], w4.z, racc[v]); fma4x2(a[3][v], w4.w, racc[v]); }
                }
            };
            // Row pass over whole-row stages, specialised on the number of cells NX of this warp's bin column
            // (constant over the item; NX = 0 is the generic loop): straight-line LDS / packed-FMA code per row.
            auto run_rows = [&](auto nx_tag) {
                constexpr int NX = decltype(nx_tag)::value;
                int rows_todo = nrows;
                const int rstride = ncols * CB;
                const in_t *sp = ring + (size_t)s * kStageFloats + xlo * CB + lch;   // this warp's first cell in stage s
                for (int st = 0; st < nstages; ++st) {
                    const int nr = min(rps, rows_todo);
                    rows_todo -= nr;
                    mbar_wait32(full32 + 8 * s, par);
                    const in_t *rp = sp;
                    for (int rr = 0; rr < nr; ++rr, rp += rstride) {
                        if (!(debug_mode & 2)) {
                            if constexpr (H16) {
                                if (NX > 0) {
                                    // all NX 128-bit loads of the row first (8 channels each), then the two packed-FMA chains
                                    uint4 cell[NX > 0 ? NX : 1];
#pragma unroll
                                    for (int i = 0; i < NX; ++i) cell[i] = *reinterpret_cast<const uint4 *>(rp + i * CB);
                                    racc[0] = mul4x2(wreg[0], bf16x4_lo(cell[0]));
                                    racc[1] = mul4x2(wreg[0], bf16x4_hi(cell[0]));
#pragma unroll
                                    for (int i = 1; i < NX; ++i) {
                                        fma4x2(racc[0], wreg[i], bf16x4_lo(cell[i]));
                                        fma4x2(racc[1], wreg[i], bf16x4_hi(cell[i]));
                                    }
                                } else {
                                    racc[0] = make_float4(0.f, 0.f, 0.f, 0.f); racc[1] = racc[0];
                                    const in_t *cp = rp;
                                    for (int i = 0; i < nx; ++i, cp += CB) {
                                        const float w = wxp[i];
                                        const uint4 q = *reinterpret_cast<const uint4 *>(cp);
                                        fma4x2(racc[0], w, bf16x4_lo(q));
                                        fma4x2(racc[1], w, bf16x4_hi(q));
                                    }
                                }
                            } else
                            if (NX > 0) {
                                // one channel half at a time: all NX loads of the half are issued back to back
                                // into their own registers, then the packed-FMA chain (first cell: MUL)
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    float4 cell[NX > 0 ? NX : 1];
#pragma unroll
                                    for (int i = 0; i < NX; ++i) cell[i] = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(rp) + i * CB + v * 128);
                                    racc[v] = mul4x2(wreg[0], cell[0]);
#pragma unroll
                                    for (int i = 1; i < NX; ++i) fma4x2(racc[v], wreg[i], cell[i]);
                                }
                            } else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                                const float *cp = reinterpret_cast<const float *>(rp);
                                for (int i = 0; i < nx; ++i, cp += CB) {
                                    const float w = wxp[i];
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) fma4x2(racc[v], w, *reinterpret_cast<const float4 *>(cp + v * 128));
                                }
                            }
                        }
                        fold();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive32(empty32 + 8 * s);
                    sp += kStageFloats;
                    if (++s == NS) { s = 0; par ^= 1; sp -= (size_t)NS * kStageFloats; }
                }
            };
            if (nseg == 1) {
                switch (nx) {
                case 1: run_rows(IntTag<1>{}); break;
                case 2: run_rows(IntTag<2>{}); break;
                case 3: run_rows(IntTag<3>{}); break;
                case 4: run_rows(IntTag<4>{}); break;
                case 5: run_rows(IntTag<5>{}); break;
                case 6: run_rows(IntTag<6>{}); break;
                default: run_rows(IntTag<0>{}); break;
                }
            } else {
                for (int row = 0; row < nrows; ++row)
                    for (int col0 = 0; col0 < ncols; col0 += SC) {
                        const int nc = min(SC, ncols - col0);
                        mbar_wait32(full32 + 8 * s, par);
                        const in_t *sb = ring + (size_t)s * kStageFloats + lch;
                        const int c_beg = max(xlo, col0), c_end = min(xlo + nx, col0 + nc);
                        for (int cx = c_beg; cx < c_end; ++cx) {
                            const float w = wxp[cx - xlo];
                            if constexpr (H16) {
                                const uint4 q = *reinterpret_cast<const uint4 *>(sb + (cx - col0) * CB);
                                fma4x2(racc[0], w, bf16x4_lo(q));
                                fma4x2(racc[1], w, bf16x4_hi(q));
                            } else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v)
                                    fma4x2(racc[v], w, *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(sb) + (cx - col0) * CB + v * 128));
                            }
                        }
                        if (col0 + nc >= ncols) {
                            fold();
#pragma unroll
                            for (int v = 0; v < VEC; ++v) racc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive32(empty32 + 8 * s);
                        if (++s == NS) { s = 0; par ^= 1; }
                    }
            }
            if (pw == 0) trace(debug_mode, k, 9, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive32(pempty32 + 8 * b);  // the slot's tables are no longer needed
            if (++b == kPlanSlots) { b = 0; fparity ^= 1; }
            while (bins_left > 0) rotate();              // bins below the last footprint row (or with no samples)
            if (pw == 0) trace(debug_mode, k, 10, lane);
        }
    }
}

template <int P, int VEC, int NS, int MINB, bool H16 = false>
static int launch_window_cfg(const Pyramid &d, int C, const float *rois, int R, int sampling_ratio,
                             int aligned, float finest_scale, const float *chan_scale,
                             const int32_t *scale_index, void *out, bool out16, int32_t *lvl_out, cudaStream_t st,
                             bool *taken)
{
    constexpr int CB = 128 * VEC;
    constexpr int S  = (P + kWin - 1) / kWin;
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    const int wx_cap = (maxW + 9 * P + 24 + 3) & ~3;           // touched cells <= extent + 2 per bin boundary, runs padded to 4, 8 slack
    const int wyd_rows = maxH;
    const size_t smem = (size_t)NS * kStageCells * CB * 4 + (size_t)kPlanSlots * (wx_cap + 4 * wyd_rows) * 4;
    const size_t smem_cap = MINB == 2 ? 115200 : 230000;           // MINB CTAs (+1 KB reserved each) must fit one SM's 228 KB
    if (smem > smem_cap) { *taken = false; return FGN_OK; }
    auto kern = chan_scale != nullptr ? roi_align_window_kernel<P, VEC, NS, MINB, true, H16, false>
                                      : roi_align_window_kernel<P, VEC, NS, MINB, false, H16, false>;
    if (H16 && out16)
        kern = chan_scale != nullptr ? roi_align_window_kernel<P, VEC, NS, MINB, true, H16, H16>
                                     : roi_align_window_kernel<P, VEC, NS, MINB, false, H16, H16>;
    // cudaFuncSetAttribute applies to the current device only and a process may drive several GPUs: set it on every
    // call (a host-side write) and size the persistent grid for the current device
    FGN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sm_count = 0, dev = 0;
    FGN_CUDA_OK(cudaGetDevice(&dev));
    FGN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    // One ticket counter per launch that can be in flight.  A launch captured into a CUDA graph bakes its
    // counter in and may run at any later time, so captured launches draw from a range that is never
    // recycled (when it is exhausted the caller falls back to the non-persistent kernel); eager launches
    // cycle through the other half, far more slots than launches can be in flight at once.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    FGN_CUDA_OK(cudaStreamIsCapturing(st, &cap));
    int slot;
    if (cap != cudaStreamCaptureStatusNone) {
        const unsigned int seq = g_captured_seq.fetch_add(1u, std::memory_order_relaxed);
        if (seq >= kTicketSlots / 2) { *taken = false; return FGN_OK; }
        slot = (int)seq;
    } else {
        slot = (int)(kTicketSlots / 2 + (g_eager_seq.fetch_add(1u, std::memory_order_relaxed) % (kTicketSlots / 2)));
    }
    const int nblk = (C + CB - 1) / CB;
    const char *e = getenv("FGN_RA_SPLIT");
    const float split_cells = e != nullptr ? (float)atof(e) : 512.f;
    const char *ed = getenv("FGN_RA_DEBUG");
    const int dbg = ed != nullptr ? atoi(ed) : 0;
    const char *eg = getenv("FGN_RA_CTAS");                   // development knob: persistent CTAs per SM (<= MINB)
    const int per_sm = eg != nullptr ? max(1, min(MINB, atoi(eg))) : MINB;
    // launches with fewer RoIs than resident CTAs still fill the machine: the chunks of their big RoIs (up to
    // ceil(P/2) per RoI in the sorted ticket scheme) are taken by the extra CTAs
    const bool sorted_scheme = nblk == 1 && R <= kSortCap && !(dbg & 64);
    const int grid = min(per_sm * sm_count, sorted_scheme ? R * ((P + 1) / 2) : R * nblk);
    // size classes of the sorted ticket scheme (footprint cells): > x: 4 chunks, > y: 2 chunks, > z / rest: whole
    float3 thr = make_float3(1000.f, 500.f, 250.f);
    // A launch with few RoIs per resident CTA (the mask branch: 100 detections on 296 CTAs) is as long as its
    // largest item -- one CTA streams about 33 GB/s -- so its RoIs are cut finer: the thresholds shrink with
    // RoIs per 2 CTAs, down to 1/8.
    if (sorted_scheme) {
        const char *ea = getenv("FGN_RA_ADAPT");              // development knob: 0 = fixed thresholds
        const float sc = (ea != nullptr && atoi(ea) == 0) ? 1.f
                         : fminf(1.f, fmaxf(0.125f, (float)R / (2.f * (float)(per_sm * sm_count))));
        thr = make_float3(thr.x * sc, thr.y * sc, thr.z * sc);
    }
    if (const char *et = getenv("FGN_RA_THR")) sscanf(et, "%f,%f,%f", &thr.x, &thr.y, &thr.z);
    (void)CB;
    kern<<<grid, (P + 2) * 32, smem, st>>>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale,
                                           scale_index, out, lvl_out, wx_cap, wyd_rows, slot,
                                           split_cells > 0.f ? split_cells : 3.0e38f, thr, dbg);
    FGN_LAUNCH_OK();
    (void)S;
    *taken = true;
    return FGN_OK;
}

// NHWC in, NHWC out.  Declines (taken=false) shapes it has no instantiation for.
int launch_roi_align_window(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int32_t *lvl_out, cudaStream_t st,
                            int ns_pref, bool *taken)
{
    *taken = false;
    if ((C & 3) != 0) return FGN_OK;
#define FGN_WIN(PV, VV, NV, MB) launch_window_cfg<PV, VV, NV, MB>(d, C, rois, R, sampling_ratio, aligned, finest_scale, \
                                                                  chan_scale, scale_index, out, false, lvl_out, st, taken)
    if (P == 7) {
        if (C > 128) return ns_pref == 2 ? FGN_WIN(7, 2, 2, 2) : FGN_WIN(7, 2, 3, 2);
        return ns_pref == 3 ? FGN_WIN(7, 1, 3, 2) : FGN_WIN(7, 1, 4, 2);
    }
    if (P == 14) {
        if (C > 128) return ns_pref == 3 ? FGN_WIN(14, 2, 3, 1) : FGN_WIN(14, 2, 5, 1);
        return FGN_WIN(14, 1, 6, 1);
    }
#undef FGN_WIN
    return FGN_OK;
}

// bf16 NHWC pyramid (uint16 storage behind the Pyramid's pointers) -> bf16 or fp32 NHWC RoI features: the same kernel
// with half-width cells (template H16).  Channel blocks are 256 wide whatever C is; lanes past C idle.
int launch_roi_align_window_bf16(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                                 int aligned, float finest_scale, const float *chan_scale,
                                 const int32_t *scale_index, void *out, int out_is_bf16, int32_t *lvl_out,
                                 cudaStream_t st, int ns_pref, bool *taken)
{
    *taken = false;
    if ((C & 7) != 0) return FGN_OK;
#define FGN_WIN16(PV, NV, MB) launch_window_cfg<PV, 2, NV, MB, true>(d, C, rois, R, sampling_ratio, aligned, finest_scale, \
                                                                     chan_scale, scale_index, out, out_is_bf16 != 0, lvl_out, st, taken)
    if (P == 7) return ns_pref == 2 ? FGN_WIN16(7, 2, 2) : (ns_pref == 4 ? FGN_WIN16(7, 4, 2) : FGN_WIN16(7, 3, 2));
    if (P == 14) return ns_pref == 3 ? FGN_WIN16(14, 3, 1) : FGN_WIN16(14, 5, 1);
#undef FGN_WIN16
    return FGN_OK;
}

void roi_align_window_trace(unsigned long long *dst, int n)
{
    const int cap = kTraceCtas * kTraceItems * kTraceEvents;
    cudaMemcpyFromSymbol(dst, g_window_trace, sizeof(unsigned long long) * (n < cap ? n : cap));
}

unsigned int roi_align_window_violations()
{
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_window_violation, sizeof(v));
    return v;
}

}  // namespace fgn

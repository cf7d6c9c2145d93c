// roi_align_gather.cu -- plan pre-pass + register-gather multi-level RoIAlign (sm_100a).
//
// Separable formulation with the reference's exact coordinate arithmetic (common.cuh):
//   out[ph,pw] = 1/count * sum_y Ay[ph][y] * sum_x Ax[pw][x] * v[y,x]
// Two kernels per call, all state in the caller's workspace:
//
//   roi_plan_kernel    one warp per (RoI, item slot).  Assigns the FPN level, evaluates the reference coordinate
//                      arithmetic once per (axis, bin, sample) and appends a compact PLAN RECORD per work item
//                      (footprint, per-bin-column x weights, per-footprint-row y weights) to the item list.  An item is
//                      (RoI, range of bin rows); RoIs with big footprints are cut into 2 / 4 (7 at P=14) items so that
//                      no item is much longer than the average -- the launch then needs no size ordering -- and RoIs
//                      whose bins are shorter than two cells into items of two bin rows (see the window below).
//                      Depends on the RoIs only, not on the feature maps.
//   roi_align_gather_kernel   persistent CTAs (4 per SM at P=7) of P warps, warp = bin column, lane = 4 channels of a
//                      128-channel block.  A CTA draws (item, channel block) tickets; the next ticket and its plan
//                      record are fetched while the current item is processed.  Per footprint row a warp loads the
//                      cells of its bin column STRAIGHT FROM L2/L1 INTO REGISTERS with 128-bit read-only loads (D rows
//                      in flight per warp, straight-line code specialised on the cells per bin column), does the x pass
//                      with packed FFMA2, and folds the row into a WINDOW of two bin rows held in registers: footprint
//                      rows are visited top to bottom, a row touches the current bin row and at most the next one (bins
//                      of >= 2 cells: an open interval of length 2 meets at most two of them; shorter bins: the item has
//                      only two bin rows), so when the row index passes the current bin's last row that bin is scaled
//                      by 1/count (and the AG-FCN channel attention), stored, and the next bin takes its place.
// Why no shared-memory staging (round-1 design: bulk copies into a ring, consumers out of shared memory): on cfg3 the
// footprints are 355-415 MB per 1000 RoIs; through a ring every byte is written to shared memory once and read ~1.5
// times (neighbouring bin columns share their border cells) = ~950 MB of shared-memory traffic, a 26 us floor at
// 128 B/clk/SM, before any of the mbarrier hand-offs; and both designs turned out to be bound by the instructions a
// warp spends per footprint row, so the lean one wins.  Measurements: DESIGN.md section 5.
#include "common.cuh"
#include <stdlib.h>
#include <stddef.h>

namespace fgn {

namespace {

constexpr int kWin       = 2;      // bin rows held in registers per warp
constexpr int kPlanWarps = 8;      // warps (= item slots) per CTA of the plan kernel
constexpr int kCB        = 128;    // channels per item ticket (lane = 4 channels)

__device__ unsigned int g_window_violation;           // planner self-check (must stay 0)

__device__ __forceinline__ void fma4x2(float4 &a, const float w, const float4 v)
{
    // packed fp32 FMA (Blackwell FFMA2: two IEEE fp32 results per lane per issue slot)
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

template <int N> struct IntTag { static constexpr int value = N; };

// Predicated 128-bit read-only load: dst keeps its value when pred is false.  (Written in PTX because a conditional
// assignment to a register array in C++ makes nvcc demote the whole array to local memory.)
__device__ __forceinline__ void ldg4_if(float4 &dst, const float *p, const bool pred)
{
    asm("{\n\t.reg .pred q;\n\t"
        "setp.ne.s32 q, %5, 0;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(dst.x), "+f"(dst.y), "+f"(dst.z), "+f"(dst.w) : "l"(p), "r"((int)pred));
}

// One work item: bin rows [pa, pb) of RoI r.  Header of the item's plan record in the workspace (and of its copy in
// shared memory); followed by the x-weight runs (wx_used floats) and one float4 of window y weights per footprint row.
template <int P>
struct alignas(16) PlanRec {
    int   rec_bytes;                               // header + x runs + nrows float4 window weights
    int   wx_used;                                 // floats of the x-weight runs (multiple of 4, 8 floats of read slack included)
    int   r, level;
    int   batch, H, W, pa;
    int   pb, X0, Y0, ncols;
    int   nrows;
    float count;
    int   pad0, pad1;
    int   xlo[P], xn[P], xoff[P];
    int   hi[P + 1];                               // last footprint row (relative to Y0) of bins <= ph
};

// workspace: [0,256) counters {items, ticket} | records [R * S0] of rec_stride bytes
struct GatherWs {
    unsigned int  *counters;
    unsigned char *recs;
    size_t         bytes;
};
inline GatherWs carve_gather_ws(void *base, int R, int S0, int rec_stride)
{
    GatherWs w;
    unsigned char *p = (unsigned char *)base;
    w.counters = (unsigned int *)p;
    w.recs     = p + 256;
    w.bytes    = 256 + (size_t)R * S0 * rec_stride;
    return w;
}

template <int P> __host__ __device__ constexpr int slots_per_roi() { return (P + 1) / 2; }

// bin rows per item when a RoI is cut into nch items
template <int P> __device__ __forceinline__ int rows_per_chunk(int nch)
{
    return (P + nch - 1) / nch;
}

}  // namespace

// ---- plan pre-pass: one warp per (RoI, item slot) ------------------------------------------------------------
// A RoI becomes nch items of ceil(P/nch) bin rows each: nch grows with the footprint (est > cut -> 2, > 2 cut -> 4,
// P=14: > 4 cut -> 7) and is ceil(P/2) when a bin is shorter than two cells (a footprint row of such a RoI can touch
// more than two bin rows of the whole RoI, but never more than the two rows of a two-row item).
template <int P>
__global__ void __launch_bounds__(kPlanWarps * 32)
roi_plan_kernel(const Pyramid pyr, const float *__restrict__ rois, const int R, const int sampling_ratio,
                const int aligned, const float finest_scale, int32_t *__restrict__ lvl_out,
                unsigned int *__restrict__ counters, unsigned char *__restrict__ recs, const int rec_stride,
                const int wx_cap, const int wyd_rows, const float cut_cells)
{
    constexpr int S0 = slots_per_roi<P>();
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char stage_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kPlanWarps + warp;
    const int r = wg / S0, j = wg - r * S0;
    if (r >= R) return;

    const float *roi = rois + 5 * (size_t)r;
    const int level = roi_level(roi, pyr, finest_scale);
    const RoiGeom g = roi_geometry(roi, pyr.scale[level], P, sampling_ratio, aligned);
    const int H = pyr.H[level], W = pyr.W[level];
    if (j == 0 && lane == 0 && lvl_out != nullptr) lvl_out[r] = level;
    const float est = (g.bin_h * (float)P + 2.f) * (g.bin_w * (float)P + 2.f);      // cells, from the box alone
    int nch = 1;
    if (est > cut_cells) nch = 2;
    if (est > 2.f * cut_cells) nch = 4;
    if (P > 8 && est > 4.f * cut_cells) nch = S0;
    if (!(g.bin_h >= 2.0f)) nch = S0;                            // (NaN geometry too)
    nch = min(nch, S0);
    const int rpc = rows_per_chunk<P>(nch);
    const int pa = j * rpc, pb = min(P, pa + rpc);
    if (pa >= P) return;                                         // this slot has no item
    unsigned int pos = 0;
    if (lane == 0) pos = atomicAdd(&counters[0], 1u);            // (result needed only when the record is written out)

    PlanRec<P> &ps = *reinterpret_cast<PlanRec<P> *>(stage_raw + (size_t)warp * rec_stride);
    float *wx = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(&ps) + sizeof(PlanRec<P>));

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
    int n4 = (n + 3) & ~3;                                       // weight runs start 16 B aligned
    int xsum = __reduce_add_sync(FULL, isx ? n4 : 0);
    if (X1 < 0 || Y1 < 0 || xsum + 8 > wx_cap || (Y1 - Y0) > wyd_rows) { X0 = X1 = Y0 = Y1 = 0; n = 0; n4 = 0; xsum = 0; }
    const int ncols = X1 - X0, nrows = (ncols > 0) ? Y1 - Y0 : 0;
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
    const int wx_used = xsum + 8;                                // warps read 8 weights per bin column whatever nx is
    if (lane == 0) {
        ps.rec_bytes = ((int)sizeof(PlanRec<P>) + 4 * (wx_used + kWin * nrows) + 15) & ~15;
        ps.wx_used = wx_used; ps.r = r; ps.level = level;
        ps.batch = g.batch; ps.H = H; ps.W = W; ps.pa = pa;
        ps.pb = pb; ps.X0 = X0; ps.Y0 = Y0; ps.ncols = ncols;
        ps.nrows = nrows; ps.count = g.count; ps.pad0 = 0; ps.pad1 = 0;
    }
    if (isx) { ps.xlo[p] = lo; ps.xn[p] = n; ps.xoff[p] = off; }
    if (lane < P) ps.hi[lane] = myhi;
    if (lane == P) ps.hi[P] = him;
    for (int i = lane; i < ((wx_used + kWin * nrows + 3) >> 2); i += 32)
        reinterpret_cast<float4 *>(wx)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    if (nrows > 0 && n > 0) {
        // Both axes run ONE instruction stream: a sample adds its two bilinear weights to two entries of
        // the record's table.  x: entry = run offset + cell.  y: footprint row j lives in window slot
        // (p - base_j) of wrow[j], base_j = first bin row of the item whose (running-max) last row is >= j.
        auto entry = [&](int cell) {
            if (isx) return off + cell - lo;
            const int jj = cell - Y0;
            int base = pa;
#pragma unroll
            for (int qq = 0; qq < P; ++qq) base += (qq >= pa && qq < pb && hiall[qq] < jj) ? 1 : 0;
            const int comp = p - base;
            if (comp < 0 || comp >= kWin) { atomicAdd(&g_window_violation, 1u); return -1; }
            return wx_used + kWin * jj + comp;
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
    pos = __shfl_sync(FULL, pos, 0);
    const int nvec = ps.rec_bytes >> 4;
    const uint4 *srcv = reinterpret_cast<const uint4 *>(&ps);
    uint4 *dstv = reinterpret_cast<uint4 *>(recs + (size_t)pos * rec_stride);
    for (int i = lane; i < nvec; i += 32) dstv[i] = srcv[i];
}

// Warp = bin column pw; lane = channels [4*lane, 4*lane+4) of the ticket's 128-channel block.
// Dynamic shared memory: 2 x rec_stride bytes (plan record of the current and of the next ticket).
template <int P, int MINB, bool SCALED, int CC>
__global__ void __launch_bounds__(P * 32, MINB)
roi_align_gather_kernel(const Pyramid pyr, const int C_rt, const float *__restrict__ chan_scale,
                        const int32_t *__restrict__ scale_index, float *__restrict__ out,
                        unsigned int *__restrict__ counters, const unsigned char *__restrict__ recs,
                        const int rec_stride, const int debug_mode)
{
    // CC: compile-time channel count (0 = run-time): cell offsets become load immediates instead of registers
    const int C = CC > 0 ? CC : C_rt;
    constexpr int LIF = 8;                           // 128-bit loads a warp keeps in flight (registers: 4 * LIF)
    extern __shared__ __align__(16) unsigned char rec_smem[];
    __shared__ unsigned int s_ticket[2];
    const int t = threadIdx.x, pw = t >> 5, lane = t & 31;
    const int nblk = (C + kCB - 1) / kCB;
    const unsigned int total = counters[0] * (unsigned)nblk;     // tickets = (item, channel block)
    const int lch = lane * 4;

    // The first NT*16 bytes of a record (3.5 KB at P=7: nearly every record entirely) are copied by one 16-byte
    // asynchronous copy per thread; the rare longer record fetches its rest directly.
    constexpr int NT = P * 32;
    auto rec_of = [&](unsigned int tk) { return recs + (size_t)(tk / (unsigned)nblk) * rec_stride; };
    unsigned int tk = blockIdx.x;
    if (tk < total && t * 16 < rec_stride)
        *reinterpret_cast<uint4 *>(rec_smem + t * 16) = *reinterpret_cast<const uint4 *>(rec_of(tk) + t * 16);
    if (t == 0) s_ticket[0] = gridDim.x + atomicAdd(&counters[1], 1u);
    int cur = 0;

    for (;;) {
        __syncthreads();                                         // record[cur] and s_ticket[cur] are in place
        if (tk >= total) break;
        const PlanRec<P> &ps = *reinterpret_cast<const PlanRec<P> *>(rec_smem + (size_t)cur * rec_stride);
        if (ps.rec_bytes > NT * 16) {                            // (CTA-uniform)
            const unsigned char *rec = rec_of(tk);
            for (int o = NT * 16 + t * 16; o < ps.rec_bytes; o += NT * 16)
                *reinterpret_cast<uint4 *>(rec_smem + (size_t)cur * rec_stride + o) = *reinterpret_cast<const uint4 *>(rec + o);
            __syncthreads();
        }
        // ---- the next ticket's record (and the ticket after it) are fetched while this item is processed
        const unsigned int tk_next = s_ticket[cur];
        if (tk_next < total && t * 16 < rec_stride) {            // asynchronous global -> shared copy (LDGSTS), no register
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rec_smem + (size_t)(cur ^ 1) * rec_stride + t * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rec_of(tk_next) + t * 16) : "memory");
        }
        // (the ticket after the next: drawn now, its value is first touched at the end of the item -- a warp that stores
        //  the result of an atomic right away waits the atomic's whole round trip)
        unsigned int tk_after = 0;
        if (t == 0) tk_after = atomicAdd(&counters[1], 1u);

        // ---- this item
        const float *wx = reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(&ps) + sizeof(PlanRec<P>));
        const float2 *wr = reinterpret_cast<const float2 *>(wx + ps.wx_used);   // y weights of the next footprint row
        const int cb0 = (int)(tk % (unsigned)nblk) * kCB;
        const bool chan_ok = cb0 + lch < C;
        const int nrows = ps.nrows;
        const int nx = ps.xn[pw];
        const float *wxp = wx + ps.xoff[pw];
        const float inv = 1.0f / ps.count;           // count is a small exact integer; <= 1 ulp vs acc/count
        const int *hip = &ps.hi[ps.pa];               // last footprint row of the current bin, of the next bins
        int bins_left = ps.pb - ps.pa;                // bin rows of the item not stored yet
        int rows_left = *hip + 1;                     // footprint rows before the current bin is complete
        float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);  // AG-FCN channel attention of this RoI (SCALED kernels only)
        if (SCALED && chan_ok) {
            const int si = scale_index != nullptr ? scale_index[ps.r] : ps.r;
            cs = ldg4(chan_scale + (size_t)si * C + cb0 + lch);
        }
        float *op = out + ((size_t)(ps.r * P + ps.pa) * P + pw) * C + cb0 + lch;   // output of the current bin
        const size_t row_pitch = (size_t)ps.W * C;
        // lanes past the channel count of a ragged last block load the block's first channels instead and never store
        const float *pn = pyr.feat[ps.level] + ((size_t)ps.batch * ps.H + ps.Y0) * row_pitch
                          + (size_t)(ps.xlo[pw]) * C + cb0 + (chan_ok ? lch : 0);        // next row to load
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, racc = a0;               // current / next bin row, x pass

        // the current bin row is complete: store it, the next one takes its place
        auto rotate = [&]() {
            float4 o;
            if (SCALED) o = make_float4(a0.x * inv * cs.x, a0.y * inv * cs.y, a0.z * inv * cs.z, a0.w * inv * cs.w);
            else        o = make_float4(a0.x * inv, a0.y * inv, a0.z * inv, a0.w * inv);   // (acc * 1/count) [* vec], as unfused
            if (chan_ok) *reinterpret_cast<float4 *>(op) = o;
            a0 = a1;
            a1 = make_float4(0.f, 0.f, 0.f, 0.f);
            op += (size_t)P * C;
            --bins_left;
        };
        // the footprint row in racc is complete: fold it into the window
        auto fold = [&]() {
            while (rows_left <= 0) {                 // the row lies past the current bin: store it, move on
                rotate();
                const int h0 = *hip++;
                rows_left += (bins_left > 0 ? *hip : 0x3fffffff) - h0;
            }
            --rows_left;
            const float2 w2 = *wr++;
            fma4x2(a0, w2.x, racc);
            if (w2.y != 0.f) fma4x2(a1, w2.y, racc);
        };

        // Row pass specialised on the cells NX of this warp's bin column (constant over the item).  A warp keeps D
        // footprint rows of its bin column in registers: row j is consumed (x pass into racc), the registers it occupied
        // immediately take the loads of row j + D, then the row is folded -- so D - 1 rows of NX 128-bit loads are in
        // flight whenever the warp waits.  MASKED: one instantiation serves every nx <= NX (cells past nx are neither
        // loaded nor accumulated).
        auto run_rows = [&](auto nx_tag, auto masked_tag) {
            constexpr int NX = decltype(nx_tag)::value;
            constexpr bool MASKED = decltype(masked_tag)::value != 0;
            constexpr int D0 = LIF / NX, D = D0 > 4 ? 4 : (D0 < 1 ? 1 : D0);
            float wreg[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) wreg[i] = wxp[i];       // (runs are padded: reading up to 8 weights is in bounds)
            float4 buf[D][NX];
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int i = 0; i < NX; ++i) buf[d][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            int jn = 0;                                          // index of the next row to load
            // (a macro, not a lambda: the buffers must stay in registers, i.e. every index a compile-time constant)
#define FGN_LOAD_ROW(d)                                                                                         \
    do {                                                                                                        \
        const bool more = jn < nrows;                                                                           \
        _Pragma("unroll") for (int i = 0; i < NX; ++i)                                                          \
            ldg4_if(buf[d][i], pn + i * C, more && (!MASKED || i < nx));                                        \
        pn += row_pitch;                                                                                        \
        ++jn;                                                                                                   \
    } while (0)
#pragma unroll
            for (int d = 0; d < D; ++d) FGN_LOAD_ROW(d);
            for (int j = 0; j < nrows; j += D) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    if (j + d < nrows) {                         // (warp-uniform)
                        racc = mul4x2(wreg[0], buf[d][0]);
#pragma unroll
                        for (int i = 1; i < NX; ++i)
                            if (!MASKED || i < nx) fma4x2(racc, wreg[i], buf[d][i]);
                        FGN_LOAD_ROW(d);
                        fold();
                    }
                }
            }
#undef FGN_LOAD_ROW
        };
        if (!(debug_mode & 2)) {
            if (nx == 0) {                                       // bin column without a valid sample: zeros
                for (int j = 0; j < nrows; ++j) fold();
            } else if (nx <= 4) {
                switch (nx) {
                case 1: run_rows(IntTag<1>{}, IntTag<0>{}); break;
                case 2: run_rows(IntTag<2>{}, IntTag<0>{}); break;
                case 3: run_rows(IntTag<3>{}, IntTag<0>{}); break;
                default: run_rows(IntTag<4>{}, IntTag<0>{}); break;
                }
            } else if (nx <= 8) {
                run_rows(IntTag<8>{}, IntTag<1>{});
            } else {
                // very wide bin columns (single-level maps with boxes wider than 50 cells): one row at a time
                for (int j = 0; j < nrows; ++j, pn += row_pitch) {
                    racc = make_float4(0.f, 0.f, 0.f, 0.f);
                    int i = 0;
                    for (; i + 4 <= nx; i += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wxp + i);
                        const float4 c0 = ldg4(pn + (size_t)i * C), c1 = ldg4(pn + (size_t)(i + 1) * C),
                                     c2 = ldg4(pn + (size_t)(i + 2) * C), c3 = ldg4(pn + (size_t)(i + 3) * C);
                        fma4x2(racc, w.x, c0); fma4x2(racc, w.y, c1); fma4x2(racc, w.z, c2); fma4x2(racc, w.w, c3);
                    }
                    for (; i < nx; ++i) fma4x2(racc, wxp[i], ldg4(pn + (size_t)i * C));
                    fold();
                }
            }
        }
        while (bins_left > 0) rotate();              // bins below the last footprint row (or with no samples)

        // ---- the prefetched record becomes the next item
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (t == 0) s_ticket[cur ^ 1] = gridDim.x + tk_after;
        tk = tk_next;
        cur ^= 1;
    }
}

template <int P> static int gather_rec_stride(int wx_cap, int wyd_rows)
{
    return ((int)sizeof(PlanRec<P>) + 4 * (wx_cap + kWin * wyd_rows) + 31) & ~15;
}
static void gather_table_caps(const Pyramid &d, int P, int *wx_cap, int *wyd_rows)
{
    int maxH = 0, maxW = 0;
    for (int l = 0; l < d.L; ++l) { maxH = max(maxH, d.H[l]); maxW = max(maxW, d.W[l]); }
    *wx_cap = (maxW + 9 * P + 24 + 3) & ~3;                    // touched cells <= extent + 2 per bin boundary, runs padded to 4, 8 slack
    *wyd_rows = maxH;
}

// Bytes of caller workspace one call needs (counters + plan records); 0 when there is no instantiation for P.
size_t roi_align_gather_workspace_bytes(const Pyramid &d, int R, int P)
{
    int wx_cap, wyd_rows;
    gather_table_caps(d, P, &wx_cap, &wyd_rows);
    if (P == 7)  return carve_gather_ws(nullptr, R, slots_per_roi<7>(), gather_rec_stride<7>(wx_cap, wyd_rows)).bytes;
    if (P == 14) return carve_gather_ws(nullptr, R, slots_per_roi<14>(), gather_rec_stride<14>(wx_cap, wyd_rows)).bytes;
    return 0;
}

template <int P, int MINB>
static int launch_gather_cfg(const Pyramid &d, int C, const float *rois, int R, int sampling_ratio,
                             int aligned, float finest_scale, const float *chan_scale,
                             const int32_t *scale_index, float *out, int32_t *lvl_out, void *workspace,
                             size_t workspace_bytes, cudaStream_t st, bool *taken)
{
    constexpr int S0 = slots_per_roi<P>();
    int wx_cap, wyd_rows;
    gather_table_caps(d, P, &wx_cap, &wyd_rows);
    const int rec_stride = gather_rec_stride<P>(wx_cap, wyd_rows);
    const size_t smem = 2 * (size_t)rec_stride;
    const size_t plan_smem = (size_t)kPlanWarps * rec_stride;
    if (smem * MINB > 200 * 1024 || plan_smem > 200 * 1024) { *taken = false; return FGN_OK; }
    const GatherWs ws = carve_gather_ws(workspace, R, S0, rec_stride);
    if (workspace == nullptr || workspace_bytes < ws.bytes) {
        set_error("roi_align: workspace %zu B < required %zu B (fgn_roi_align_ml_workspace_bytes)", workspace_bytes, ws.bytes);
        return FGN_ERR_WORKSPACE;
    }
    auto kern = C == 256 ? (chan_scale != nullptr ? roi_align_gather_kernel<P, MINB, true, 256> : roi_align_gather_kernel<P, MINB, false, 256>)
                         : (chan_scale != nullptr ? roi_align_gather_kernel<P, MINB, true, 0> : roi_align_gather_kernel<P, MINB, false, 0>);
    // (cudaFuncSetAttribute applies to the current device only: set it on every call, it is a cheap host-side write)
    if (smem > 48 * 1024) FGN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (plan_smem > 48 * 1024)
        FGN_CUDA_OK(cudaFuncSetAttribute(roi_plan_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem));
    int dev = 0, sm_count = 0;
    FGN_CUDA_OK(cudaGetDevice(&dev));
    FGN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const char *ed = getenv("FGN_RA_DEBUG");
    const int dbg = ed != nullptr ? atoi(ed) : 0;
    const char *eg = getenv("FGN_RA_CTAS");                   // development knob: persistent CTAs per SM (<= MINB)
    const int per_sm = eg != nullptr ? max(1, min(MINB, atoi(eg))) : MINB;
    const int nblk = (C + kCB - 1) / kCB;
    // items longer than `cut` cells are halved / quartered; a launch with few RoIs per resident CTA (the mask branch:
    // 100 detections) is as long as its largest item, so its cut shrinks with the work per CTA
    const char *ec = getenv("FGN_RA_CUT");
    float cut = ec != nullptr ? (float)atof(ec) : 256.f;
    const float per_cta = (float)R * (float)nblk / (float)(per_sm * sm_count);
    if (ec == nullptr && per_cta < 2.f) cut = fmaxf(32.f, cut * per_cta * 0.5f);
    const long max_tickets = (long)R * S0 * nblk;
    const int grid = (int)min((long)per_sm * sm_count, max_tickets);
    FGN_CUDA_OK(cudaMemsetAsync(ws.counters, 0, 16, st));
    const int plan_ctas = (R * S0 + kPlanWarps - 1) / kPlanWarps;
    roi_plan_kernel<P><<<plan_ctas, kPlanWarps * 32, plan_smem, st>>>(
        d, rois, R, sampling_ratio, aligned, finest_scale, lvl_out, ws.counters, ws.recs, rec_stride, wx_cap, wyd_rows, cut);
    FGN_LAUNCH_OK();
    if (!(dbg & 128)) {                                       // (development: bit 7 = plan pre-pass only)
        kern<<<grid, P * 32, smem, st>>>(d, C, chan_scale, scale_index, out, ws.counters, ws.recs, rec_stride, dbg);
        FGN_LAUNCH_OK();
    }
    *taken = true;
    return FGN_OK;
}

// NHWC in, NHWC out.  Declines (taken=false) shapes it has no instantiation for.
int launch_roi_align_gather(const Pyramid &d, int C, int P, const float *rois, int R, int sampling_ratio,
                            int aligned, float finest_scale, const float *chan_scale,
                            const int32_t *scale_index, float *out, int32_t *lvl_out, void *workspace,
                            size_t workspace_bytes, cudaStream_t st, bool *taken)
{
    *taken = false;
    if ((C & 3) != 0) return FGN_OK;
    if (P == 7)
        return launch_gather_cfg<7, 4>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale, scale_index, out,
                                       lvl_out, workspace, workspace_bytes, st, taken);
    if (P == 14)
        return launch_gather_cfg<14, 2>(d, C, rois, R, sampling_ratio, aligned, finest_scale, chan_scale, scale_index, out,
                                        lvl_out, workspace, workspace_bytes, st, taken);
    return FGN_OK;
}

unsigned int roi_align_gather_violations()
{
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_window_violation, sizeof(v));
    return v;
}

}  // namespace fgn

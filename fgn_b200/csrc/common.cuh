// common.cuh -- shared host/device helpers for libfgn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/fgn_b200.h"

namespace fgn {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define FGN_CHECK_ARG(cond, ...)                                                        \
    do { if (!(cond)) { fgn::set_error(__VA_ARGS__); return FGN_ERR_INVALID_ARG; } } while (0)

#define FGN_CUDA_OK(expr)                                                               \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) {                            \
        fgn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),         \
                       __FILE__, __LINE__); return FGN_ERR_CUDA; } } while (0)

#define FGN_LAUNCH_OK()                                                                 \
    do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) {                \
        fgn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),     \
                       __FILE__, __LINE__); return FGN_ERR_CUDA; }                      \
        fgn::count_launch(); } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute applies to the CURRENT device only and a process may drive several GPUs, so nothing about a
// device is cached in process-wide statics: launches that need more than the default 48 KB of dynamic shared memory opt
// in on every call (a host-side write, no synchronisation), and persistent grids are sized from the current device.
#define FGN_SMEM_OPTIN(kernel, bytes)                                                                   \
    do { if ((size_t)(bytes) > 48 * 1024)                                                               \
        FGN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); } while (0)

static inline int current_sm_count(int *out)
{
    int dev = 0;
    FGN_CUDA_OK(cudaGetDevice(&dev));
    FGN_CUDA_OK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return FGN_OK;
}

// ---- RoI geometry: the integer/coordinate contract -----------------------------------------
// Mirrors oracle/roi_align_ref.c (= mmcv RoIAlign / torchvision roi_align CPU).
// Every coordinate operation is an explicitly rounded fp32 op (__f*_rn are never contracted
// into FMAs), so the (int) truncations below land on the same cells as the CPU reference.

struct RoiGeom {
    int   batch;
    float start_w, start_h, bin_w, bin_h;
    int   grid_h, grid_w;
    float count;
};

// `B` (images in the feature maps): a RoI whose batch index lies outside [0, B) -- a stale or corrupt rois[:,0] -- is
// given an empty sampling grid: every forward kernel then writes zeros for it and the backward kernels add nothing,
// instead of reading or writing out of bounds.  (The reference would raise an index error on the host.)
__device__ __forceinline__ RoiGeom roi_geometry(const float *roi, float spatial_scale, int P,
                                                int sampling_ratio, int aligned, int B = 0x7fffffff)
{
    RoiGeom g;
    const float off = aligned ? 0.5f : 0.0f;
    g.batch = (int)roi[0];
    const float sw = __fsub_rn(__fmul_rn(roi[1], spatial_scale), off);
    const float sh = __fsub_rn(__fmul_rn(roi[2], spatial_scale), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], spatial_scale), off);
    const float eh = __fsub_rn(__fmul_rn(roi[4], spatial_scale), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!aligned) { rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; }
    g.start_w = sw; g.start_h = sh;
    g.bin_h = __fdiv_rn(rh, (float)P);
    g.bin_w = __fdiv_rn(rw, (float)P);
    g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)P));
    g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)P));
    if ((unsigned)g.batch >= (unsigned)B) { g.batch = 0; g.grid_h = g.grid_w = 0; }
    const int cnt = g.grid_h * g.grid_w;
    g.count = (float)(cnt > 1 ? cnt : 1);
    return g;
}

struct AxisSample { int valid, low, high; float l, h; };

// coordinate = start + p*bin + (i+.5)*bin/grid, evaluated left to right, unfused.
__device__ __forceinline__ AxisSample axis_sample(float start, float bin, int grid, int size,
                                                  int p, int i)
{
    AxisSample s;
    float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                        __fdiv_rn(__fmul_rn(__fadd_rn((float)i, .5f), bin), (float)grid));
    s.valid = !(c < -1.0f || c > (float)size);
    if (!s.valid) { s.low = s.high = 0; s.l = s.h = 0.f; return s; }
    if (c <= 0.0f) c = 0.0f;
    int low = (int)c, high;
    if (low >= size - 1) { high = low = size - 1; c = (float)low; }
    else                 { high = low + 1; }
    s.low = low; s.high = high;
    s.l = __fsub_rn(c, (float)low);
    s.h = __fsub_rn(1.0f, s.l);
    return s;
}

// Pyramid by value in kernel parameter space.
struct Pyramid {
    int          L;
    int          B;                          // images per level (batch indices outside [0,B) pool zeros)
    const float *feat[FGN_MAX_LEVELS];
    int          H[FGN_MAX_LEVELS];
    int          W[FGN_MAX_LEVELS];
    float        scale[FGN_MAX_LEVELS];
    float        lvl_thr[FGN_MAX_LEVELS];   // lvl_thr[k]: smallest fp32 v with floor(log2f(v)) >= k
};

// mmdet map_roi_levels: lvl = clamp(floor(log2(sqrt(w*h)/finest + 1e-6)), 0, L-1), all fp32.
// floor(log2f(v)) is taken on the correctly rounded fp32 logarithm (what torch CPU returns for the
// values next to powers of two, tests/test_oracle.py).  Because v -> floor(fl(log2 v)) is monotone,
// the level is the number of thresholds T_k <= v, with T_k the smallest fp32 whose rounded log2 is
// >= k; the host derives T_k from (float)log2((double)v) (level_thresholds below), the same
// expression oracle/roi_align_ref.c evaluates per RoI.  NaN (negative area) compares false -> 0.
__device__ __forceinline__ int roi_level(const float *roi, const Pyramid &pyr, float finest_scale)
{
    if (pyr.L <= 1) return 0;
    const float area  = __fmul_rn(__fsub_rn(roi[3], roi[1]), __fsub_rn(roi[4], roi[2]));
    const float scale = __fsqrt_rn(area);
    const float v     = __fadd_rn(__fdiv_rn(scale, finest_scale), 1e-6f);
    int lvl = 0;
#pragma unroll
    for (int k = 1; k < FGN_MAX_LEVELS; ++k) lvl += (k < pyr.L && v >= pyr.lvl_thr[k]) ? 1 : 0;
    return lvl;
}

void level_thresholds(float *thr /* [FGN_MAX_LEVELS] */);

static inline Pyramid to_device_pyramid(const fgn_pyramid_t *p, int B = 0x7fffffff)
{
    Pyramid d;
    d.L = p->num_levels;
    d.B = B > 0 ? B : 0x7fffffff;
    for (int i = 0; i < FGN_MAX_LEVELS; ++i) {
        d.feat[i]  = i < p->num_levels ? p->feat[i] : nullptr;
        d.H[i]     = i < p->num_levels ? p->H[i] : 0;
        d.W[i]     = i < p->num_levels ? p->W[i] : 0;
        d.scale[i] = i < p->num_levels ? p->spatial_scale[i] : 0.f;
    }
    level_thresholds(d.lvl_thr);
    return d;
}

// ---- small device utilities -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float *p)
{
    return __ldg(reinterpret_cast<const float4 *>(p));
}

__device__ __forceinline__ void fma4(float4 &a, float w, const float4 &v)
{
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y);
    a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

}  // namespace fgn

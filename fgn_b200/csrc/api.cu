// api.cu -- library-wide C-ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include <atomic>
#include <string.h>
#include <math.h>

namespace fgn {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// T_k = smallest fp32 v with floorf((float)log2((double)v)) >= k (see roi_level in common.cuh).
void level_thresholds(float *thr)
{
    struct Table {                                    // built once; C++11 guarantees a race-free initialisation
        float v[FGN_MAX_LEVELS];
        Table()
        {
            v[0] = 0.f;
            for (int k = 1; k < FGN_MAX_LEVELS; ++k) {
                float x = ldexpf(1.0f, k);
                for (;;) {
                    const float pred = nextafterf(x, 0.0f);
                    if (floorf((float)log2((double)pred)) >= (float)k) x = pred; else break;
                }
                v[k] = x;
            }
        }
    };
    static const Table table;
    for (int k = 0; k < FGN_MAX_LEVELS; ++k) thr[k] = table.v[k];
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace fgn

extern "C" int fgn_abi_version(void) { return FGN_ABI_VERSION; }

extern "C" const char *fgn_last_error_string(void) { return fgn::g_err; }

extern "C" uint64_t fgn_launch_count(void) { return fgn::g_launches.load(std::memory_order_relaxed); }

extern "C" int fgn_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    FGN_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    FGN_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return FGN_OK;
}

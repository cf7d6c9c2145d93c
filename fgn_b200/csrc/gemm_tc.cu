// gemm_tc.cu -- tcgen05 (5th-gen tensor core) contraction for the relation head.
#include "gemm.cuh"

namespace fgn {

// Placeholder until the tcgen05 kernel lands: declines every shape so the dispatcher uses the
// fp32 SIMT kernel.
int gemm_nt_tc(const float *, int, const float *, int, const float *, float *, int, int, int, int,
               int, cudaStream_t, bool *taken)
{
    *taken = false;
    return FGN_OK;
}

}  // namespace fgn

// l2_fabric_bench.cu -- how fast can the SMs pull L2-resident bytes into shared memory / registers?
// Standalone (nvcc -arch=sm_100a tools/l2_fabric_bench.cu -o tools/_build/l2_fabric_bench).
// Sets the ceiling the RoIAlign producers (cp.async.bulk rows -> smem ring) can reach.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One issuing thread per CTA; stage = nrow copies of row_bytes each (rows `pitch` bytes apart).
template <int NS>
__global__ void bulk_kernel(const char *__restrict__ buf, size_t buf_bytes, int row_bytes, int nrow, size_t pitch,
                            int stages_per_cta)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[NS];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int stage_bytes = row_bytes * nrow;
    size_t off = ((size_t)blockIdx.x * 2654435761u * 4096) % (buf_bytes - (size_t)nrow * pitch - row_bytes);
    off &= ~(size_t)1023;
    int par[NS];
    for (int s = 0; s < NS; ++s) par[s] = 0;
    for (int it = 0; it < stages_per_cta + NS; ++it) {
        const int s = it % NS;
        if (it >= NS) { mbar_wait(&bar[s], par[s]); par[s] ^= 1; }
        if (it < stages_per_cta) {
            mbar_expect_tx(&bar[s], stage_bytes);
            for (int r = 0; r < nrow; ++r)
                bulk_g2s(smem + (size_t)s * stage_bytes + (size_t)r * row_bytes, buf + off + (size_t)r * pitch, row_bytes, &bar[s]);
            off += (size_t)nrow * pitch + 7168;
            if (off + (size_t)nrow * pitch + row_bytes >= buf_bytes) off = (off * 7) % (buf_bytes / 2) & ~(size_t)1023;
        }
    }
}

// every thread: LDG.128 stream, `unroll` loads in flight
__global__ void ldg_kernel(const float4 *__restrict__ buf, size_t n4, int iters, float *sink)
{
    float4 acc = make_float4(0, 0, 0, 0);
    size_t i = ((size_t)blockIdx.x * blockDim.x * 8 * 977 + threadIdx.x) % n4;
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { size_t j = i + (size_t)u * blockDim.x; if (j >= n4) j -= n4; v[u] = __ldcg(buf + j); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        i += (size_t)blockDim.x * 8; if (i >= n4) i -= n4;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

int main()
{
    const size_t MB = 1 << 20;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    char *buf; float *sink;
    CK(cudaMalloc(&buf, 512 * MB)); CK(cudaMemset(buf, 1, 512 * MB)); CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t sizes[] = {32 * MB, 92 * MB, 512 * MB};
    for (size_t bytes : sizes) {
        // ---- bulk copies ----
        for (int per_sm = 1; per_sm <= 4; ++per_sm)
            for (int cfg = 0; cfg < 4; ++cfg) {
                const int row_bytes = cfg == 0 ? 8192 : (cfg == 1 ? 32768 : (cfg == 2 ? 2048 : 16384));
                const int nrow = 32768 / row_bytes;
                const size_t pitch = cfg == 1 ? 32768 : 336 * 1024;
                const int NS = 3;
                const size_t smem = (size_t)NS * 32768;
                if (per_sm * (smem + 1024) > 227 * 1024) continue;
                CK(cudaFuncSetAttribute(bulk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const int grid = sms * per_sm, stages = 600 / per_sm;
                for (int w = 0; w < 2; ++w) bulk_kernel<3><<<grid, 32, smem>>>(buf, bytes, row_bytes, nrow, pitch, stages);
                CK(cudaEventRecord(e0));
                for (int w = 0; w < 5; ++w) bulk_kernel<3><<<grid, 32, smem>>>(buf, bytes, row_bytes, nrow, pitch, stages);
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double moved = 5.0 * grid * (double)stages * 32768;
                printf("{\"kind\":\"bulk\",\"buf_MB\":%zu,\"cta_per_sm\":%d,\"row_bytes\":%d,\"rows_per_stage\":%d,\"NS\":%d,\"GB/s\":%.0f}\n",
                       bytes / MB, per_sm, row_bytes, nrow, NS, moved / ms / 1e6);
            }
        // ---- LDG.128 ----
        for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
            const int grid = sms * per_sm, iters = 4096 / per_sm;
            for (int w = 0; w < 2; ++w) ldg_kernel<<<grid, 256>>>((const float4 *)buf, bytes / 16, iters, sink);
            CK(cudaEventRecord(e0));
            for (int w = 0; w < 5; ++w) ldg_kernel<<<grid, 256>>>((const float4 *)buf, bytes / 16, iters, sink);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double moved = 5.0 * grid * 256.0 * iters * 8 * 16;
            printf("{\"kind\":\"ldg128\",\"buf_MB\":%zu,\"cta_per_sm\":%d,\"GB/s\":%.0f}\n", bytes / MB, per_sm, moved / ms / 1e6);
        }
    }
    return 0;
}

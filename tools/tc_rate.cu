// tc_rate.cu -- how long does one tcgen05.mma (kind::f16, M = 128, K = 16) take as a function of N?
// One CTA per SM issues `iters` MMAs back to back (same operands, alternating TMEM halves) and
// commits once; clock64 around issue + completion.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long *out, int vary)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *tmemPtr = reinterpret_cast<uint32_t *>(smem + 65536 + 16);
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmemPtr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmemPtr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da = make_desc(smem_u32(smem)), db = make_desc(smem_u32(smem) + 4096);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (uint32_t)((i & 1) * 256);
            if (vary) {
                da = make_desc(smem_u32(smem) + (uint32_t)(i & 3) * 4096u);
                db = make_desc(smem_u32(smem) + 16384u + (uint32_t)((i >> 2) % 6) * 6144u + (uint32_t)((i >> 1) & 1) * 1024u);
            }
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        } while (!done);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66048);
    const int Ns[5] = {32, 64, 128, 256, 16};
    for (int vary : {0, 1})
        for (int n = 0; n < 4; ++n)
            for (int iters : {64, 1024}) {
                const int grid = 148;
                rate_kernel<<<grid, 128, 66048>>>(Ns[n], iters, d, vary);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("vary %d N %3d iters %4d: issue %7lld clk, issue+complete %7lld clk  (%.1f clk/MMA)\n",
                       vary, Ns[n], iters, h[0], h[1], (double)h[1] / iters);
            }
    return 0;
}

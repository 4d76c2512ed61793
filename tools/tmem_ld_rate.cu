// tmem_ld_rate.cu -- how fast can the epilogue warps of one CTA read TMEM?  W warps (warp w reads lanes
// 32*(w%4)...) loop over tcgen05.ld.32x32b.x32 with and without .pack::16b, wait::ld after every load or
// after every second one; no MMA runs.  Prints TMEM cells (lane x column) per clock per SM: the bound of a
// kernel whose tensor-core results are consumed once each (K = 16: one MMA fills 128 x 256 cells in 128 clk).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(PACK, r, addr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32" PACK ".b32 "                                                       \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr) : "memory")

template <int kPack, int kDepth>
__global__ void __launch_bounds__(512, 1) ld_kernel(int iters, long long *out, unsigned *sink)
{
    __shared__ uint32_t tmemPtr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmemPtr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemPtr + ((uint32_t)(warp & 3) * 32u << 16);
    const int colsPerLd = kPack ? 64 : 32;
    uint32_t acc = 0;
    uint32_t col = (uint32_t)((warp >> 2) * colsPerLd) & 511u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32], q[32];
        if (kPack) LD32(".pack::16b", r, tmem + col); else LD32("", r, tmem + col);
        col = (col + 4 * colsPerLd) & 511u;
        if (kDepth == 2) {
            if (kPack) LD32(".pack::16b", q, tmem + col); else LD32("", q, tmem + col);
            col = (col + 4 * colsPerLd) & 511u;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        #pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= r[k];
        if (kDepth == 2) {
            #pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= q[k];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmemPtr) : "memory");
}

template <int kPack, int kDepth>
static void run(int warps, long long *d_out, unsigned *d_sink)
{
    const int iters = 4096;
    ld_kernel<kPack, kDepth><<<148, warps * 32>>>(iters, d_out, d_sink);
    ld_kernel<kPack, kDepth><<<148, warps * 32>>>(iters, d_out, d_sink);
    long long clk = 0;
    cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const double cells = (double)warps * iters * kDepth * 32.0 * (kPack ? 64 : 32);
    printf("warps %2d  %s  loads in flight %d : %8lld clk  %6.1f cells/clk/SM  (%.0f clk per load per warp)%s\n", warps,
           kPack ? "x32.pack::16b (64 col)" : "x32           (32 col)", kDepth, clk, cells / clk, (double)clk / (iters * kDepth),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    long long *d_out;
    unsigned *d_sink;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&d_sink, 4096);
    for (int warps : {4, 8, 16}) {
        run<0, 1>(warps, d_out, d_sink);
        run<0, 2>(warps, d_out, d_sink);
        run<1, 1>(warps, d_out, d_sink);
        run<1, 2>(warps, d_out, d_sink);
    }
    return 0;
}

// tc_probe.cu -- pins the tcgen05 conventions the tensor-core pair filter relies on (sm_100a):
//   * shared-memory operand layout: 128-row blocks [kchunk(2)][row(128)][8 halves], K-major,
//     no swizzle; descriptor LBO = 2048 B (k-chunk stride), SBO = 128 B (8-row group stride)
//   * kind::f16 MMA, M = 128, N = 64, K = 16, fp32 accumulators: D row -> TMEM lane, D col -> column
//   * tcgen05.ld.32x32b.x64 per warp (lanes 32*(warp%4)..)
// Prints the max |D_gpu - D_cpu| for the layout; a wrong convention shows up as O(1) errors.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe.bin tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lboBytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sboBytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    return d;                                     // base offset 0, layout type 0 = no swizzle
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done, addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// A: 512 rows (4 blocks), B: 128 rows (1 block), both [row][16] halves in global memory
__global__ void __launch_bounds__(544, 1)
probe_kernel(const __half *A, const __half *B, float *D, uint32_t lbo, uint32_t sbo, uint32_t idesc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __half *sA = reinterpret_cast<__half *>(smem);                  // 4 x 4 KB
    __half *sB = reinterpret_cast<__half *>(smem + 16384);          // 4 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 20480);
    uint32_t *tmemPtr = reinterpret_cast<uint32_t *>(smem + 20480 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // block layout: [kc][row][8]
    for (int i = tid; i < 512 * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16, blk = row / 128, r = row % 128, kc = k / 8, kk = k % 8;
        sA[blk * 2048 + kc * 1024 + r * 8 + kk] = A[i];
    }
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16, kc = k / 8, kk = k % 8;
        sB[kc * 1024 + row * 8 + kk] = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> async proxy (tensor core)
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmemPtr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmemPtr;

    if (warp == 16 && lane == 0) {
        for (int h = 0; h < 2; ++h)
            for (int t = 0; t < 4; ++t) {
                const uint64_t da = make_desc(smem_u32(sA) + t * 4096, lbo, sbo);
                const uint64_t db = make_desc(smem_u32(sB) + h * 1024, lbo, sbo);
                const uint32_t dcol = tmem + h * 256 + t * 64;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     :: "r"(smem_u32(bar)) : "memory");
    }
    if (warp < 16) {
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, t = warp >> 2;
        for (int h = 0; h < 2; ++h) {
            uint32_t v[64];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + t * 64);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                         "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                         "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
                           "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
                           "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
                           "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
                           "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = t * 128 + q * 32 + lane;
            for (int n = 0; n < 64; ++n) D[(size_t)row * 128 + h * 64 + n] = __uint_as_float(v[n]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

int main()
{
    std::vector<__half> hA(512 * 16), hB(128 * 16);
    std::vector<float> fA(512 * 16), fB(128 * 16);
    srand(7);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 64.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 64.0f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 512 * 128 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    // instruction descriptor: c_format F32 (1 << 4), a/b F16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t combos[3][2] = {{2048, 128}, {128, 2048}, {2048, 256}};
    for (int c = 0; c < 3; ++c) {
        cudaMemset(dD, 0xff, 512 * 128 * 4);
        probe_kernel<<<1, 544, 32768>>>(dA, dB, dD, combos[c][0], combos[c][1], idesc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo %u sbo %u: CUDA error %s\n", combos[c][0], combos[c][1], cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(512 * 128);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double maxErr = 0, maxRef = 0; int bad = 0;
        for (int r = 0; r < 512; ++r)
            for (int n = 0; n < 128; ++n) {
                double ref = 0;
                for (int k = 0; k < 16; ++k) ref += (double)fA[r * 16 + k] * (double)fB[n * 16 + k];
                const double err = fabs(ref - (double)hD[r * 128 + n]);
                if (!(err <= 1e-2)) ++bad;
                if (err > maxErr || err != err) maxErr = err;
                if (fabs(ref) > maxRef) maxRef = fabs(ref);
            }
        printf("lbo %4u sbo %4u: max |err| %.6g (max |ref| %.1f), %d of %d entries off by > 1e-2\n",
               combos[c][0], combos[c][1], maxErr, maxRef, bad, 512 * 128);
    }
    return 0;
}

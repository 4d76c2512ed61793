// tc_probe16.cu -- can the MMA itself subtract the threshold and hand back PACKED fp16 accumulators?
//   D = A.B^T with K slots 6/7 carrying (1, 1) x (-thr_hi, -thr_lo), D format F16, read back with
//   tcgen05.ld ... .pack::16b (two columns per 32-bit register).
// Checks (a) the column -> register/half mapping, (b) that the fp16 result is the correctly rounded
// value of the exact sum (internal accumulation wider than fp16), so the sign of D decides "candidate".
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done, addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// A: 128 rows, B: 64 rows, [row][16] halves.  out[row][32] packed registers
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half *A, const __half *B, uint32_t *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __half *sA = reinterpret_cast<__half *>(smem);
    __half *sB = reinterpret_cast<__half *>(smem + 4096);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192);
    uint32_t *tmemPtr = reinterpret_cast<uint32_t *>(smem + 8192 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16, kc = k / 8, kk = k % 8;
        sA[kc * 1024 + row * 8 + kk] = A[i];
    }
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16, kc = k / 8, kk = k % 8;
        sB[kc * 1024 + row * 8 + kk] = (row < 64) ? B[i] : __float2half(0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(tmemPtr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmemPtr;
    if (tid == 0) {
        const uint32_t idesc = (0u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);     // D = F16
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(make_desc(smem_u32(sA))), "l"(make_desc(smem_u32(sB))), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; ++n) out[(size_t)(warp * 32 + lane) * 32 + n] = v[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem) : "memory");
}

int main()
{
    const double PI = 3.14159265358979323846, c = 63.0, rho = c / (2 * PI), P = 3 * rho * rho, thr = 9.4;
    const double thrDot = P - thr / 2;
    std::vector<__half> hA(128 * 16), hB(64 * 16);
    std::vector<double> xa(128 * 3), xb(64 * 3);
    srand(3);
    for (auto &x : xa) x = c * (rand() / (double)RAND_MAX);
    // B atoms: close to some A atom so that many D are near zero
    for (int j = 0; j < 64; ++j)
        for (int k = 0; k < 3; ++k) xb[j * 3 + k] = xa[(j * 2) * 3 + k] + 3.2 * (rand() / (double)RAND_MAX - 0.5);
    auto emb = [&](double x, double *o) { o[0] = rho * cos(2 * PI * x / c); o[1] = rho * sin(2 * PI * x / c); };
    const __half thi = __float2half((float)thrDot);
    const __half tlo = __float2half((float)(thrDot - (double)__half2float(thi)));
    std::vector<double> ea(128 * 16, 0.0), eb(64 * 16, 0.0);
    for (int i = 0; i < 128; ++i) {
        double e[6];
        for (int k = 0; k < 3; ++k) emb(xa[i * 3 + k], e + 2 * k);
        for (int k = 0; k < 16; ++k) hA[i * 16 + k] = __float2half(0.f);
        for (int k = 0; k < 6; ++k) {
            const __half h = __float2half((float)e[k]);
            hA[i * 16 + k] = h;
            hA[i * 16 + 8 + k] = __float2half((float)(e[k] - (double)__half2float(h)));
        }
        hA[i * 16 + 6] = __float2half(1.f); hA[i * 16 + 7] = __float2half(1.f);
        for (int k = 0; k < 16; ++k) ea[i * 16 + k] = __half2float(hA[i * 16 + k]);
    }
    for (int j = 0; j < 64; ++j) {
        double e[6];
        for (int k = 0; k < 3; ++k) emb(xb[j * 3 + k], e + 2 * k);
        for (int k = 0; k < 16; ++k) hB[j * 16 + k] = __float2half(0.f);
        for (int k = 0; k < 6; ++k) { hB[j * 16 + k] = __float2half((float)e[k]); hB[j * 16 + 8 + k] = hB[j * 16 + k]; }
        hB[j * 16 + 6] = __float2half(-__half2float(thi)); hB[j * 16 + 7] = __float2half(-__half2float(tlo));
        for (int k = 0; k < 16; ++k) eb[j * 16 + k] = __half2float(hB[j * 16 + k]);
    }
    __half *dA, *dB; uint32_t *dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 32 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    probe_kernel<<<1, 128, 16384>>>(dA, dB, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> hO(128 * 32);
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    // mapping hypothesis: register n holds columns 2n (low half) and 2n+1 (high half)
    double maxRel = 0, maxAbsSmall = 0; int signBad = 0, nSmall = 0, bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 64; ++j) {
            double ref = 0;
            for (int k = 0; k < 16; ++k) ref += ea[i * 16 + k] * eb[j * 16 + k];
            const uint32_t w = hO[i * 32 + j / 2];
            const __half_raw hr = { (unsigned short)((j & 1) ? (w >> 16) : (w & 0xffff)) };
            const double got = __half2float(__half(hr));
            const double err = fabs(got - ref);
            const double tol = fabs(ref) * 4.9e-4 + 1e-6;            // one fp16 rounding
            if (err > tol * 1.01) ++bad;
            if (fabs(ref) > 1e-3) maxRel = fmax(maxRel, err / fabs(ref));
            if (fabs(ref) < 2.0) { ++nSmall; maxAbsSmall = fmax(maxAbsSmall, err); }
            if ((got >= 0) != (ref >= 0) && fabs(ref) > 1e-3) ++signBad;
        }
    printf("pack::16b mapping (reg n = cols 2n | 2n+1): %d of %d entries outside one fp16 rounding; max rel err %.3g; "
           "%d entries with |D| < 2: max abs err %.3g; sign mismatches (|D| > 1e-3): %d\n",
           bad, 128 * 64, maxRel, nSmall, maxAbsSmall, signBad);
    return 0;
}

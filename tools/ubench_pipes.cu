// ubench_pipes.cu -- issue-rate probe for the instructions the pair filters are built from (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes tools/ubench_pipes.cu
// Prints warp-instructions per clock per SM for each variant (4 = one per scheduler per clock).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int NV = 16;   // independent chains per thread

template <int kV>
__global__ void __launch_bounds__(256) probe(unsigned *out, int iters, unsigned s, unsigned t)
{
    unsigned v[NV];
    #pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = threadIdx.x * 0x01010101u + i * 0x3c003c00u;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (kV == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(s), "r"(t));
            if (kV == 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(s), "r"(t));
            if (kV == 2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(s));
            if (kV == 3) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(s));
            if (kV == 4) asm volatile("min.f32 %0, %0, %1;" : "+r"(v[i]) : "r"(s));
            if (kV == 5) asm volatile("{.reg .f16x2 a; abs.f16x2 a, %0; sub.rn.f16x2 %0, %1, a;}" : "+r"(v[i]) : "r"(s));
            if (kV == 6) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(s));
            if (kV == 7) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(s), "r"(t));
            if (kV == 8) asm volatile("set.lt.u32.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(s));
            if (kV == 9) asm volatile("{.reg .f16x2 a; abs.f16x2 a, %0; min.f16x2 %0, a, %1;}" : "+r"(v[i]) : "r"(s));
        }
        if (kV == 10) {   // filter-shaped mix per 2 chains: 3 x (sub, sub|.|, min|.|, fma)
            #pragma unroll
            for (int i = 0; i < NV; i += 2) {
                asm volatile("{.reg .f16x2 d, e, w, a;\n\t"
                             "sub.rn.f16x2 d, %0, %2;\n\tabs.f16x2 a, d;\n\tsub.rn.f16x2 e, %3, a;\n\tmin.f16x2 w, a, e;\n\t"
                             "fma.rn.f16x2 %1, w, w, %1;\n\t}"
                             : "+r"(v[i]), "+r"(v[i + 1]) : "r"(s), "r"(t));
            }
        }
        if (kV == 11) {   // same in the add form: all on the fma pipe
            #pragma unroll
            for (int i = 0; i < NV; i += 2) {
                asm volatile("{.reg .f16x2 d, e, w, a, b;\n\t"
                             "sub.rn.f16x2 d, %0, %2;\n\tabs.f16x2 a, d;\n\tsub.rn.f16x2 e, a, %3;\n\tabs.f16x2 b, e;\n\t"
                             "sub.rn.f16x2 w, %3, b;\n\t"
                             "fma.rn.f16x2 %1, w, w, %1;\n\t}"
                             : "+r"(v[i]), "+r"(v[i + 1]) : "r"(s), "r"(t));
            }
        }
    }
    unsigned acc = 0;
    #pragma unroll
    for (int i = 0; i < NV; ++i) acc ^= v[i];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int kV>
static void run(const char *name, double instrPerIter, unsigned *d, int nSM, double clkGHz)
{
    const int iters = 4096, blocks = nSM * 8;
    probe<kV><<<blocks, 256>>>(d, 16, 0x3c003c00u, 0x00010001u);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<kV><<<blocks, 256>>>(d, iters, 0x3c003c00u, 0x00010001u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warpInstr = (double)blocks * 8 * iters * instrPerIter;
    const double perClkSM = warpInstr / (best * 1e-3) / (clkGHz * 1e9) / nSM;
    printf("%-34s %8.3f ms   %6.3f warp-instr/clk/SM\n", name, best, perClkSM);
}

int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("%s  SMs %d  clock %.3f GHz (attribute; rates assume it)\n", pr.name, pr.multiProcessorCount, ghz);
    unsigned *d; cudaMalloc(&d, 64);
    const int n = pr.multiProcessorCount;
    run<0>("fma.f32", NV, d, n, ghz);
    run<1>("fma.f16x2", NV, d, n, ghz);
    run<7>("fma.bf16x2", NV, d, n, ghz);
    run<2>("add.f16x2", NV, d, n, ghz);
    run<6>("mul.f16x2", NV, d, n, ghz);
    run<5>("sub.f16x2 c - |x|", NV, d, n, ghz);
    run<3>("min.f16x2", NV, d, n, ghz);
    run<9>("min.f16x2 |x|", NV, d, n, ghz);
    run<4>("min.f32", NV, d, n, ghz);
    run<8>("set.lt.u32.f16x2", NV, d, n, ghz);
    run<10>("mix sub,sub,min,fma (per instr)", NV / 2 * 4, d, n, ghz);
    run<11>("mix sub,sub,sub,fma (per instr)", NV / 2 * 4, d, n, ghz);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

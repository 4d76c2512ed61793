// Host gather of the pageable upload path under realistic conditions: rows of 2.4 KB pieces 12 KB apart
// copied by OpenMP threads into (a) malloc'ed or (b) cudaMallocHost staging, with or without a concurrent H2D DMA of the
// other staging buffer -- to see what the 37 GB/s of the library's gather (against 77 GB/s for plain memcpy) comes from.
//   nvcc -O2 -Xcompiler -fopenmp -o /tmp/hgb2 tools/host_gather_bench.cu && /tmp/hgb2 <threads>
#include <cuda_runtime.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/mman.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 16;
    const long nA = 25000, nF = 1000, nRows = 9154, wf = 200;
    const size_t w = wf * 12, bytes = nRows * w;
    float *src = (float *)malloc(nA * nF * 12);
    memset(src, 1, nA * nF * 12);
    long *rows = (long *)malloc(nRows * 8);
    for (long r = 0; r < nRows; ++r) rows[r] = r < 1231 ? r : 1231 + 3 * (r - 1231);
    char *dstM = (char *)aligned_alloc(4096, bytes), *dstP[2], *dstR[2], *dev;
    cudaMallocHost(&dstP[0], bytes); cudaMallocHost(&dstP[1], bytes); cudaMalloc(&dev, bytes);
    memset(dstM, 0, bytes); memset(dstP[0], 0, bytes); memset(dstP[1], 0, bytes);
    // (c) ordinary anonymous memory, 2 MB aligned, transparent huge pages requested, touched, THEN page-locked
    const size_t rb = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    for (int i = 0; i < 2; ++i) {
        dstR[i] = (char *)aligned_alloc(2u << 20, rb);
        madvise(dstR[i], rb, MADV_HUGEPAGE);
        memset(dstR[i], 0, rb);
        cudaError_t e = cudaHostRegister(dstR[i], rb, cudaHostRegisterPortable);
        if (e != cudaSuccess) printf("cudaHostRegister: %s\n", cudaGetErrorString(e));
    }
    cudaStream_t cs; cudaStreamCreate(&cs);
    for (int mode = 0; mode < 6; ++mode) {                  // bit 0: pinned destination, bit 1: concurrent DMA; 4/5: registered THP
        double sum = 0; int cnt = 0;
        for (int rep = 0; rep < 24; ++rep) {
            const long a0 = (rep % 5) * wf;
            char *dst = mode >= 4 ? dstR[rep & 1] : (mode & 1) ? dstP[rep & 1] : dstM;
            const bool dma = (mode & 2) || mode == 5;
            if (dma) cudaMemcpyAsync(dev, mode >= 4 ? dstR[(rep & 1) ^ 1] : dstP[(rep & 1) ^ 1], bytes, cudaMemcpyHostToDevice, cs);
            const double t0 = now();
            #pragma omp parallel for schedule(static) num_threads(T)
            for (long r = 0; r < nRows; ++r) memcpy(dst + r * w, src + (rows[r] * nF + a0) * 3, w);
            const double dt = now() - t0;
            if (dma) cudaStreamSynchronize(cs);
            if (rep > 3) { sum += dt; ++cnt; }
        }
        printf("T=%d %s destination, %s: %.3f ms per 22 MB = %.1f GB/s\n", T,
               mode >= 4 ? "registered huge-page" : (mode & 1) ? "cudaMallocHost" : "malloc",
               ((mode & 2) || mode == 5) ? "H2D DMA of the other buffer running" : "no DMA", sum / cnt * 1e3, bytes / (sum / cnt) / 1e9);
    }
    // H2D rate out of the two kinds of page-locked memory
    for (int kind = 0; kind < 2; ++kind) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        char *h = kind ? dstR[0] : dstP[0];
        cudaMemcpyAsync(dev, h, bytes, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(e0, cs);
        for (int i = 0; i < 20; ++i) cudaMemcpyAsync(dev, h, bytes, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(e1, cs); cudaStreamSynchronize(cs);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("H2D from %s: %.1f GB/s\n", kind ? "registered huge-page memory" : "cudaMallocHost memory", bytes * 20.0 / (ms * 1e-3) / 1e9);
    }
    return 0;
}

// Host-side microbenchmark of the row gather of the pageable upload path: rows of 2.4 KB pieces, 12 KB apart.
//   gcc -O2 -mavx2 -fopenmp tools/host_gather_bench.c -o /tmp/hgb && /tmp/hgb <threads> <variant 0 memcpy | 1 + prefetch | 2 NT stores | 3 both> <frames per piece>
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include <time.h>
#include <immintrin.h>
#include <stdint.h>
static double now(){struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec+1e-9*t.tv_nsec;}
// variant 0: memcpy; 1: memcpy + prefetch next row piece; 2: AVX2 loads + NT stores; 3: 2 + prefetch
static void copy_nt(char*d,const char*s,size_t w){
  size_t i=0;
  // align dst to 32
  while(((uintptr_t)(d+i)&31) && i<w){d[i]=s[i];++i;}
  for(;i+32<=w;i+=32){ __m256i v=_mm256_loadu_si256((const __m256i*)(s+i)); _mm256_stream_si256((__m256i*)(d+i),v);}
  for(;i<w;++i)d[i]=s[i];
}
int main(int argc,char**argv){
  int T=argc>1?atoi(argv[1]):8; int variant=argc>2?atoi(argv[2]):0; int wf=argc>3?atoi(argv[3]):200;
  const long nA=25000,nF=1000; const long nRows=9154;
  float*src=malloc(nA*nF*12); for(long i=0;i<nA*nF*3;i+=64)src[i]=i;
  memset(src,1,nA*nF*12);
  size_t w=wf*12; char*dst=aligned_alloc(4096,nRows*w+4096); memset(dst,0,nRows*w);
  long*rows=malloc(nRows*8); for(long r=0;r<nRows;++r) rows[r]= r<1231? r : 1231+3*(r-1231);
  double best=1e9, sum=0; int cnt=0;
  for(int rep=0;rep<22;++rep){
    long a0=(rep%5)*wf;
    double t0=now();
    #pragma omp parallel for schedule(static) num_threads(T)
    for(long r=0;r<nRows;++r){
      const char*s=(const char*)(src+(rows[r]*nF+a0)*3);
      if(variant&1){ if(r+1<nRows){const char*n=(const char*)(src+(rows[r+1]*nF+a0)*3); for(size_t o=0;o<w;o+=64)_mm_prefetch(n+o,_MM_HINT_T0);} }
      if(variant&2) copy_nt(dst+r*w,s,w); else memcpy(dst+r*w,s,w);
    }
    if(variant&2)_mm_sfence();
    double dt=now()-t0; if(rep>1){ if(dt<best)best=dt; sum+=dt; cnt++; }
  }
  printf("T=%d variant=%d w=%zu: best %.3f ms mean %.3f ms  %.1f GB/s (mean)\n",T,variant,w,best*1e3,sum/cnt*1e3,nRows*w/(sum/cnt)/1e9);
  return 0;
}

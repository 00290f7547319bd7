// FP64 DMMA throughput microbenchmark (scratch)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, const double *gin) {
  double a0 = gin[threadIdx.x], a1 = gin[threadIdx.x + 1], a2 = gin[threadIdx.x + 2], a3 = gin[threadIdx.x + 3];
  double b0 = gin[threadIdx.x + 4], b1 = gin[threadIdx.x + 5];
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // m8n8k4: 8 independent accumulators
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[2 * j]), "+d"(c[2 * j + 1]) : "d"(a0), "d"(b0));
    } else if (MODE == 1) {  // m16n8k8: A 4 regs, B 2 regs, C 4 regs; 4 independent accumulators
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[4 * j]), "+d"(c[4 * j + 1]), "+d"(c[4 * j + 2]), "+d"(c[4 * j + 3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    } else if (MODE == 2) {  // m16n8k4: A 2 regs, B 1 reg, C 4 regs
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[4 * j]), "+d"(c[4 * j + 1]), "+d"(c[4 * j + 2]), "+d"(c[4 * j + 3])
                     : "d"(a0), "d"(a1), "d"(b0));
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int warps_per_sm, double fma_per_warp_iter, double *out, double *gin) {
  int iters = 20000, threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, 100, gin);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters, gin);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double fma = (double)blocks * threads / 32 * iters * fma_per_warp_iter;
  printf("%-12s warps/SM=%2d  %.3f ms  %.2f TFLOP/s (err=%s)\n", name, warps_per_sm, ms, 2 * fma / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double *out, *gin; cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&gin, 4096 * 8);
  double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1e-3 * (i % 7); 
  cudaMemcpy(gin, h, sizeof h, cudaMemcpyHostToDevice);
  for (int w : {8, 16, 32}) {
    run<0>("m8n8k4", w, 8 * 256.0, out, gin);
    run<1>("m16n8k8", w, 4 * 1024.0, out, gin);
    run<2>("m16n8k4", w, 4 * 512.0, out, gin);
  }
  return 0;
}

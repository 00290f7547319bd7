// FP64 issue-rate microbenchmarks for B200 (scratch; not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
struct Coef { double c[12]; };
#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, Coef cf, const double *gin) {
  double x[CHAINS];
  for (int i = 0; i < CHAINS; ++i) x[i] = gin[threadIdx.x + i];
  double ra = gin[100], rb = gin[101];
  double rc[12];
  for (int i = 0; i < 12; ++i) rc[i] = gin[200 + i];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // DFMA reg operands, 8 chains x 12
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], ra, rb);
    } else if (MODE == 1) {  // DFMA with 12 distinct kernel-param coefficients (c-bank / UR operands)
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], ra, cf.c[j]);
    } else if (MODE == 2) {  // 12 distinct register coefficients
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], ra, rc[j]);
    } else if (MODE == 3) {  // DMUL
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = x[i] * ra;
    } else if (MODE == 4) {  // DADD
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = x[i] + ra;
    } else if (MODE == 5) {  // rcp64h + 2 DFMA newton, 8 chains x 4
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          double y;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
          const double e = fma(-x[i], y, 1.0);
          x[i] = fma(y, e, y);
        }
    } else if (MODE == 6) {  // DFMA 8 chains x 12 interleaved with 1 integer op per DFMA
      int acc = it;
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) { x[i] = fma(x[i], ra, rb); acc = acc * 3 + i; }
      if (acc == 123456789) x[0] += 1.0;
    } else if (MODE == 7) {  // DFMA x12x8 with 2 integer ops per DFMA
      int acc = it, acc2 = it;
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) { x[i] = fma(x[i], ra, rb); acc = acc * 3 + i; acc2 = (acc2 ^ acc) + j; }
      if (acc == 123456789 || acc2 == 987654321) x[0] += 1.0;
    } else if (MODE == 8) {  // DSETP + FSEL pairs
#pragma unroll
      for (int j = 0; j < 12; ++j)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = (x[i] < rc[j]) ? x[i] + ra : rb;
    } else if (MODE == 9) {  // single dependent chain per warp (latency)
#pragma unroll
      for (int j = 0; j < 96; ++j) x[0] = fma(x[0], ra, rb);
    }
  }
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int warps_per_sm, double ops_per_iter, double *out, double *gin) {
  int dev_sms = 148;
  Coef cf; for (int i = 0; i < 12; ++i) cf.c[i] = 1.0 + 1e-9 * i;
  int iters = 20000;
  int threads = 256, blocks = dev_sms * warps_per_sm * 32 / threads;
  if (warps_per_sm * 32 < threads) { threads = warps_per_sm * 32; blocks = dev_sms; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, 100, cf, gin);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters, cf, gin);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double warp_instr = (double)blocks * threads / 32 * iters * ops_per_iter;
  double per_smsp_cyc = ms * 1e-3 * 1.965e9;  // cycles
  double instr_per_smsp = warp_instr / (148.0 * 4);
  printf("%-34s warps/SM=%2d  %.3f ms  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_sm, ms, per_smsp_cyc / instr_per_smsp);
}
int main() {
  double *out, *gin; cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&gin, 4096 * 8);
  double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0 + 1e-7 * i; h[100] = 0.999999; h[101] = 1e-7;
  cudaMemcpy(gin, h, sizeof h, cudaMemcpyHostToDevice);
  for (int w : {4, 16, 32}) {
    if (w == 4) { run<0>("DFMA reg", 4, 96, out, gin); run<1>("DFMA const coef", 4, 96, out, gin); run<2>("DFMA 12 reg coef", 4, 96, out, gin);
      run<3>("DMUL", 4, 96, out, gin); run<4>("DADD", 4, 96, out, gin); run<5>("rcp64h+2DFMA (per 3 instr group)", 4, 32, out, gin);
      run<6>("DFMA + 1 int (per DFMA)", 4, 96, out, gin); run<7>("DFMA + 2 int (per DFMA)", 4, 96, out, gin); run<8>("DSETP+FSEL+DADD (per group)", 4, 96, out, gin);
      run<9>("DFMA dependent chain", 4, 96, out, gin); }
    if (w == 16) { run<0>("DFMA reg", 16, 96, out, gin); run<1>("DFMA const coef", 16, 96, out, gin); run<2>("DFMA 12 reg coef", 16, 96, out, gin);
      run<3>("DMUL", 16, 96, out, gin); run<4>("DADD", 16, 96, out, gin); run<5>("rcp64h+2DFMA (per 3 instr group)", 16, 32, out, gin);
      run<6>("DFMA + 1 int (per DFMA)", 16, 96, out, gin); run<7>("DFMA + 2 int (per DFMA)", 16, 96, out, gin); run<8>("DSETP+FSEL+DADD (per group)", 16, 96, out, gin);
      run<9>("DFMA dependent chain", 16, 96, out, gin); }
    if (w == 32) { run<0>("DFMA reg", 32, 96, out, gin); run<1>("DFMA const coef", 32, 96, out, gin); run<5>("rcp64h+2DFMA (per 3 instr group)", 32, 32, out, gin); run<7>("DFMA + 2 int (per DFMA)", 32, 96, out, gin); }
  }
  return 0;
}

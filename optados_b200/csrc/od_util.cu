// od_util.cu -- measurement helpers: device stopwatch, FP64 peak microbenchmark, contribution counter,
// pinned host memory.
#include <algorithm>
#include <cmath>

#include "od_internal.h"
#include "od_spectral.h"

namespace {
// 8 independent FMA chains per thread: enough ILP to saturate the FP64 pipe at 4+ warps per SMSP
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

// The same pipe through the FP64 tensor-core instruction (DMMA m16n8k8, 4 independent accumulator sets per warp): 1024
// FMAs per instruction.  On B200 this reaches the pipe's 64 FMA / clock / SM (37.1 TFLOP/s at 1965 MHz); the DFMA loop
// above issues every 2.18 cycles instead of 2 (register-operand read ports) and reads 34.2.
__global__ void __launch_bounds__(256) fp64_dmma_peak_kernel(double *out, int iters, const double *gin) {
  const double a0 = gin[threadIdx.x & 7], a1 = gin[(threadIdx.x + 1) & 7], a2 = gin[(threadIdx.x + 2) & 7], a3 = gin[(threadIdx.x + 3) & 7];
  const double b0 = gin[(threadIdx.x + 4) & 7], b1 = gin[(threadIdx.x + 5) & 7];
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[4 * j]), "+d"(c[4 * j + 1]), "+d"(c[4 * j + 2]), "+d"(c[4 * j + 3])
                   : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 12345.678) out[8] = s;
}

__global__ void count_contrib_kernel(DosProducer prod, GridSpec grid, double *__restrict__ out /* [gridDim.x] */) {
  double acc = 0.0;
  const long long total = prod.units_per_z * prod.ns;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int is = (int)(t / prod.units_per_z);
    const long long u = t % prod.units_per_z;
    RecOut r;
    prod(u, is, r);
    double lo, hi;
    if (r.mode == 0) { lo = r.e0 - OD_SQRT60 * r.p1; hi = r.e0 + OD_SQRT60 * r.p1; }
    else { lo = r.p4; hi = 2.0 * r.e0 - r.p4; }
    double jlo = ceil((lo - grid.emin) / grid.delta), jhi = floor((hi - grid.emin) / grid.delta);
    jlo = fmax(jlo, 0.0);
    jhi = fmin(jhi, (double)(grid.nbins - 1));
    if (jhi >= jlo) acc += jhi - jlo + 1.0;
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}
}  // namespace

extern "C" int od_timer_start(od_ctx *ctx) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->t0) { OD_CUDA(ctx, cudaEventCreate(&ctx->t0)); OD_CUDA(ctx, cudaEventCreate(&ctx->t1)); }
  OD_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
  return OD_OK;
}

extern "C" int od_timer_stop(od_ctx *ctx, double *elapsed_ms) {
  if (!ctx || !elapsed_ms || !ctx->t0) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
  OD_CUDA(ctx, cudaEventSynchronize(ctx->t1));
  float ms = 0.f;
  OD_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
  *elapsed_ms = ms;
  return OD_OK;
}

// mode 0: the larger of the two loops (the pipe peak); 1: the DFMA loop only; 2: the DMMA loop only
static int measure_fp64(od_ctx *ctx, double seconds, int mode, double *tflops) {
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CHECK(od_buf_ensure(ctx, ctx->small, 256));
  OD_CUDA(ctx, cudaMemsetAsync(ctx->small.p, 0, 256, ctx->stream));
  const int blocks = ctx->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int which = 0; which < 2; ++which) {
    if (mode == 1 && which == 1) continue;
    if (mode == 2 && which == 0) continue;
    int iters = 2000;
    double elapsed_total = 0.0;
    for (int rep = 0; rep < 64 && elapsed_total < seconds * 0.5e3; ++rep) {
      OD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
      if (which == 0) fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->small.as<double>(), iters, 0.999999, 1e-7);
      else fp64_dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->small.as<double>() + 8, iters, ctx->small.as<double>());
      OD_LAUNCH_CHECK(ctx);
      OD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
      OD_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
      elapsed_total += ms;
      // FMAs: 64 per thread and iteration in the DFMA loop; 4 x 1024 per warp and iteration in the DMMA loop
      const double fma = which == 0 ? 64.0 * (double)iters * blocks * threads : 4096.0 * (double)iters * blocks * (threads / 32);
      if (rep > 0) best = std::max(best, 2.0 * fma / (ms * 1e-3) / 1e12);
      if (ms < 20.f) iters *= 2;
    }
  }
  *tflops = best;
  return OD_OK;
}

extern "C" int od_measure_fp64_peak(od_ctx *ctx, double seconds, double *tflops) {
  if (!ctx || !tflops) return OD_ERR_ARG;
  return measure_fp64(ctx, seconds, 0, tflops);
}

extern "C" int od_measure_fp64_peak_mode(od_ctx *ctx, double seconds, int mode, double *tflops) {
  if (!ctx || !tflops || mode < 0 || mode > 2) return OD_ERR_ARG;
  return measure_fp64(ctx, seconds, mode, tflops);
}

extern "C" int od_count_contributions(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins,
                                      const od_broadening *br, double *contributions) {
  if (!ctx || !contributions) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p || !ctx->kweight.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CHECK(od_grads_resident(ctx));
  if (dos_type != 'f' && !ctx->grads.p) return od_fail(ctx, OD_ERR_STATE, "band_gradient not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  DosProducer p;
  p.units_per_z = (long long)ctx->nb * ctx->nk;
  p.k_first = 0; p.nk_g = ctx->nk;
  p.nb = ctx->nb; p.ns = ctx->ns; p.nk = ctx->nk;
  p.bands = ctx->bands.as<double>(); p.grads = ctx->grads.as<double>(); p.kw = ctx->kweight.as<double>();
  p.eps = ctx->eps;
  OD_CHECK(od_make_broadspec(ctx, dos_type, br, delta_bins, false, &p.b));
  p.mw = nullptr; p.norb = 0; p.mw_nb = 0; p.mw_nk = 0; p.og = 0;
  GridSpec grid{emin, delta_bins, nbins};
  const int blocks = ctx->sm_count * 8;
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(double) * (size_t)blocks));
  count_contrib_kernel<<<blocks, 256, 0, ctx->stream>>>(p, grid, ctx->small.as<double>());
  OD_LAUNCH_CHECK(ctx);
  std::vector<double> h(blocks);
  OD_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->small.p, sizeof(double) * (size_t)blocks, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double s = 0.0;
  for (double v : h) s += v;
  *contributions = s;
  return OD_OK;
}

extern "C" int od_host_alloc_pinned(size_t nbytes, void **hptr) {
  if (!hptr) return OD_ERR_ARG;
  return cudaHostAlloc(hptr, nbytes ? nbytes : 8, cudaHostAllocDefault) == cudaSuccess ? OD_OK : OD_ERR_CUDA;
}
// page-lock / release memory the CALLER allocated (a Fortran array): what cudaHostRegister does, so that OD_MEM_HOST /
// OD_MEM_HOST_DEFERRED uploads from it run at the PCIe rate and overlap with kernels
extern "C" int od_host_register(void *hptr, size_t nbytes) {
  if (!hptr || nbytes == 0) return OD_ERR_ARG;
  return cudaHostRegister(hptr, nbytes, cudaHostRegisterDefault) == cudaSuccess ? OD_OK : OD_ERR_CUDA;
}
extern "C" int od_host_unregister(void *hptr) { return cudaHostUnregister(hptr) == cudaSuccess ? OD_OK : OD_ERR_CUDA; }
extern "C" int od_host_free_pinned(void *hptr) { return cudaFreeHost(hptr) == cudaSuccess ? OD_OK : OD_ERR_CUDA; }

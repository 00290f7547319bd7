// od_spectral.cu -- C-ABI entry points of the hot loops: calculate_dos (dos_utils.f90:1140),
// calculate_dos_at_e (:1651), calculate_jdos (jdos_utils.f90:280) and their merges.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "od_internal.h"
#include "od_producers.cuh"
#include "od_spectral.h"

// ------------------------------------------------------------------ helpers
static int type_index(char t) { return t == 'f' ? 0 : t == 'a' ? 1 : t == 'l' ? 2 : -1; }

int od_make_broadspec(od_ctx *ctx, char type, const od_broadening *br, double delta_bins, bool honor_hybrid, BroadSpec *out) {
  const int ti = type_index(type);
  if (ti < 0) return od_fail(ctx, OD_ERR_ARG, "unknown dos_type '%c' (dos_utils.f90:1205)", type);
  if (!br) return od_fail(ctx, OD_ERR_ARG, "null od_broadening");
  BroadSpec b;
  const GridGeom g = od_grid_geom(ctx, br->adaptive_smearing);
  b.type = ti;
  b.fixed_w = br->fixed_smearing;
  b.adapt_temp = g.adaptive_smearing_temp;
  b.delta_bins = delta_bins;
  b.fbc = br->finite_bin_correction ? 1 : 0;
  b.honor_hybrid = honor_hybrid ? 1 : 0;
  b.hybrid = br->hybrid_linear ? 1 : 0;
  b.hybrid_tol = br->hybrid_linear_grad_tol;
  b.fast_corners = 0;
  for (int i = 0; i < 3; ++i) b.step[i] = g.step[i];
  for (int i = 0; i < 9; ++i) b.rs[i] = g.recip[i];
  *out = b;
  return OD_OK;
}

static int check_ready(od_ctx *ctx, bool need_grad) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  if (!ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "band_energy not set (od_set_bands)");
  if (!ctx->kweight.p) return od_fail(ctx, OD_ERR_STATE, "kpoint_weight not set");
  if (need_grad && !ctx->grads.p && !ctx->grads_host) return od_fail(ctx, OD_ERR_STATE, "band_gradient not set (od_set_gradients)");
  return OD_OK;
}

namespace {
// numerical_intdos (dos_utils.f90:1256-1258, :1315): intdos(j) = delta * sum_{i<=j} dos(i)
__global__ void numerical_intdos_kernel(const double *__restrict__ dos, double *__restrict__ intdos, int nbins, int ns, double delta) {
  const int is = blockIdx.x * blockDim.x + threadIdx.x;
  if (is >= ns) return;
  double run = 0.0;
  for (int j = 0; j < nbins; ++j) {
    run += dos[(size_t)is * nbins + j];
    intdos[(size_t)is * nbins + j] = run * delta;
  }
}

// linear_smearing post-smear (dos_utils.f90:1282-1309)
__global__ void linear_smear_kernel(const double *__restrict__ dos, double *__restrict__ out, int nbins, int ns, double emin,
                                    double delta, double smear) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int is = blockIdx.y;
  if (j >= nbins) return;
  const double Ej = emin + (double)j * delta;
  double s = 0.0;
  for (int i = 0; i < nbins; ++i) {
    const double Ei = emin + (double)i * delta;
    s += dos[(size_t)is * nbins + i] * od_gaussian(Ej, smear, Ei) * delta;
  }
  out[(size_t)is * nbins + j] = s;
}
}  // namespace

// ------------------------------------------------------------------ calculate_dos
// k-point slabs (measured on B200, config 2, see profiles/r02_summary.md): the accumulation kernels like long record
// arrays (fewer launch tails, fuller buckets), the scatter of the records likes short ones (its frontier of half-written
// lines has to stay mergeable in the L2), the 64-byte linear records more so than the 32-byte Gaussian ones.  Step time with
// the final kernels, (slab, linear sub-slab) in units: (48M, 24M) 117.8 ms, (64M, 32M) 116.0, (96M, 32M) 117.4, (128M, 32M)
// 118.7, (96M, 16M) 121.8, (96M, 48M) 120.6, (96M, 96M) 127.1 (prepare 47.5 instead of 34.8 ms).
#define OD_SLAB_UNITS_DEFAULT (64ll << 20)
#define OD_LINEAR_SLAB_UNITS_DEFAULT (32ll << 20)
// streamed (host) gradients: the copy of the first slab and the kernels of the last one are not overlapped -- short slabs
#define OD_STREAM_SLAB_UNITS (32ll << 20)
// one k-point slab of one broadening scheme, ADDED into accum = [dos | intdos | weighted_dos]
static int dos_producer(od_ctx *ctx, char dos_type, const od_broadening *br, const GridSpec &grid, const double *mw_dev, int norb,
                        int mw_nb, int mw_nk, long long k_first, int nk_slab, const double *grads_dev, int nk_g, DosProducer *out) {
  DosProducer p;
  p.units_per_z = (long long)ctx->nb * nk_slab;
  p.k_first = k_first;
  p.nb = ctx->nb; p.ns = ctx->ns; p.nk = ctx->nk; p.nk_g = nk_g;
  p.bands = ctx->bands.as<double>(); p.grads = grads_dev; p.kw = ctx->kweight.as<double>();
  p.eps = ctx->eps;
  OD_CHECK(od_make_broadspec(ctx, dos_type, br, grid.delta, /*honor_hybrid=*/false, &p.b));
  p.mw = mw_dev; p.norb = norb; p.mw_nb = mw_nb; p.mw_nk = mw_nk; p.og = OD_MAXW;
  p.nb_div = FastDiv((unsigned)ctx->nb);
  *out = p;
  return OD_OK;
}

static int dos_slab(od_ctx *ctx, char dos_type, const od_broadening *br, const GridSpec &grid, const double *mw_dev, int norb,
                    int mw_nb, int mw_nk, long long k_first, int nk_slab, const double *grads_dev, int nk_g, double *accum) {
  const int ns = ctx->ns, nbins = grid.nbins;
  DosProducer p;
  OD_CHECK(dos_producer(ctx, dos_type, br, grid, mw_dev, norb, mw_nb, mw_nk, k_first, nk_slab, grads_dev, nk_g, &p));
  const size_t n1 = (size_t)nbins * ns;
  if (ctx->engine != OD_ENGINE_DENSE && od_narrow_grid_fits(nbins, ns)) return od_narrow_dos(ctx, p, grid, dos_type == 'l', accum);
  if (!mw_dev) {
    std::vector<long long> dest(ns * 2);
    for (int is = 0; is < ns; ++is) {
      dest[is * 2 + 0] = (long long)is * nbins;
      dest[is * 2 + 1] = (long long)(n1 + (size_t)is * nbins);
    }
    return od_run_dense<0, true, DosProducer>(ctx, p, grid, ns, dest, accum);
  }
  const int ngroups = (norb + OD_MAXW - 1) / OD_MAXW;
  const int nz = ns * ngroups;
  constexpr int NACC = 2 + OD_MAXW;
  std::vector<long long> dest((size_t)nz * NACC, -1);
  for (int g = 0; g < ngroups; ++g)
    for (int is = 0; is < ns; ++is) {
      const int z = is + ns * g;
      if (g == 0) {
        dest[(size_t)z * NACC + 0] = (long long)is * nbins;
        dest[(size_t)z * NACC + 1] = (long long)(n1 + (size_t)is * nbins);
      }
      for (int a = 0; a < OD_MAXW; ++a) {
        const int o = g * OD_MAXW + a;
        if (o < norb) dest[(size_t)z * NACC + 2 + a] = (long long)(2 * n1 + (size_t)nbins * (is + (size_t)ns * o));
      }
    }
  return od_run_dense<OD_MAXW, true, DosProducer>(ctx, p, grid, nz, dest, accum);
}

// Several broadening schemes over the rank's k-points in ONE sweep of k-point slabs.  With deferred (host)
// gradients the slabs are uploaded on a second stream into two buffers while the kernels work on the
// previous slab; otherwise there is a single slab.  accum_t = accum_base + t * stride:
// [dos(B,ns) | intdos(B,ns) | weighted_dos(B,ns,norb)], all schemes contiguous, followed by one status word and a
// scratch array.  merge: ONE sum-allreduce over [all schemes | status] (dos_utils_merge, dos_utils.f90:1013-1053, for
// every scheme at once); the status word makes a rank-local failure an error on EVERY rank instead of a hang in the
// collective.  host_out (may be null): the merged block is copied there (ntypes * stride doubles) in one transfer,
// with the synchronisation and the deferred-error check of the call.
int od_dos_accumulate(od_ctx *ctx, int ntypes, const char *types, const bool *weighted, double emin, double delta_bins, int nbins,
                      const od_broadening *br, const double *mw_dev, int norb, int mw_nb, int mw_nk, int merge, double **accum_base,
                      size_t *accum_stride, double *host_out) {
  bool need_grad = false;
  for (int t = 0; t < ntypes; ++t) need_grad = need_grad || types[t] != 'f';
  OD_CHECK(check_ready(ctx, need_grad));
  if (nbins < 2) return od_fail(ctx, OD_ERR_ARG, "nbins = %d", nbins);
  if (!(delta_bins > 0.0)) return od_fail(ctx, OD_ERR_ARG, "delta_bins must be positive");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ns = ctx->ns, nb = ctx->nb, nk = ctx->nk;
  if (mw_dev && (norb < 1 || mw_nb < 1 || mw_nk < 1)) return od_fail(ctx, OD_ERR_ARG, "bad matrix_weights bounds");
  if (mw_dev && mw_nk != nk) return od_fail(ctx, OD_ERR_REF, "ERROR : DOS : mw%%nkpoints not equal to nkpoints. (dos_utils.f90:159)");
  if (!mw_dev) norb = 0;
  GridSpec grid{emin, delta_bins, nbins};
  std::vector<BroadSpec> bspec((size_t)ntypes);
  std::vector<char> fixedwide((size_t)ntypes, 0);  // fixed broadening with very wide windows: moment expansion
  for (int t = 0; t < ntypes; ++t) {
    OD_CHECK(od_make_broadspec(ctx, types[t], br, delta_bins, false, &bspec[t]));  // validates dos_type
    fixedwide[t] = od_fixedwide_applies(ctx, bspec[t], grid, OD_ERF_CUT * OD_SQRT_TWO);
  }
  const size_t n1 = (size_t)nbins * ns;
  const size_t stride = 2 * n1 + n1 * norb;
  const size_t nall = stride * ntypes;  // + 1 status word + n1 scratch
  OD_CHECK(od_buf_ensure(ctx, ctx->accum, sizeof(double) * (nall + 1 + n1)));
  double *accum = ctx->accum.as<double>();
  OD_CUDA(ctx, cudaMemsetAsync(accum, 0, sizeof(double) * (nall + 1), ctx->stream));

  // everything rank-local runs inside this lambda: its return code travels in the status word of the merge
  auto local = [&]() -> int {
  for (int t = 0; t < ntypes; ++t)
    if (fixedwide[t]) OD_CHECK(od_fixedwide_begin(ctx, bspec[t], grid, 1 + (weighted[t] ? norb : 0)));
  auto one_slab = [&](int t, long long k0, int nks, const double *grads_dev, int nk_g) -> int {
    if (fixedwide[t]) {
      DosProducer p;
      OD_CHECK(dos_producer(ctx, types[t], br, grid, weighted[t] ? mw_dev : nullptr, norb, mw_nb, mw_nk, k0, nks, grads_dev, nk_g, &p));
      return od_fixedwide_accumulate_dos(ctx, p, grid, weighted[t] ? norb : 0);
    }
    if (types[t] == 'l') {
      const long long lin_units = getenv("OD_LINEAR_SLAB_UNITS") ? atoll(getenv("OD_LINEAR_SLAB_UNITS")) : OD_LINEAR_SLAB_UNITS_DEFAULT;
      const int sub_nk = (int)std::max<long long>(1, lin_units / ((long long)nb * ns));
      for (int s0 = 0; s0 < nks; s0 += sub_nk)
        OD_CHECK(dos_slab(ctx, types[t], br, grid, weighted[t] ? mw_dev : nullptr, norb, mw_nb, mw_nk, k0 + s0, std::min(sub_nk, nks - s0),
                          grads_dev ? grads_dev + (size_t)nb * 3 * s0 : nullptr, nk_g, accum + stride * t));
      return OD_OK;
    }
    return dos_slab(ctx, types[t], br, grid, weighted[t] ? mw_dev : nullptr, norb, mw_nb, mw_nk, k0, nks, grads_dev, nk_g, accum + stride * t);
  };
  const bool streamed = ctx->grads_host != nullptr && need_grad;
  if (!streamed) {
    // resident inputs: k-point slabs (see OD_SLAB_UNITS_DEFAULT above); OD_SLAB_UNITS overrides for experiments.
    long long slab_units = getenv("OD_SLAB_UNITS") ? atoll(getenv("OD_SLAB_UNITS")) : OD_SLAB_UNITS_DEFAULT;
    const int slab_nk = (int)std::max<long long>(1, std::min<long long>(nk, slab_units / ((long long)nb * ns)));
    for (long long k0 = 0; k0 < nk; k0 += slab_nk) {
      const int nks = (int)std::min<long long>(slab_nk, nk - k0);
      for (int t = 0; t < ntypes; ++t)
        OD_CHECK(one_slab(t, k0, nks, ctx->grads.p ? ctx->grads.as<double>() + (size_t)nb * 3 * k0 : nullptr, nk));
    }
  } else {
    if (!ctx->copy_stream) {
      OD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        OD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        OD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
      }
    }
    const long long units = (long long)nb * ns * nk;
    const int nslabs = (int)std::max<long long>(1, std::min<long long>({64, (units + OD_STREAM_SLAB_UNITS - 1) / OD_STREAM_SLAB_UNITS,
                                                                      (long long)nk / 1024}));
    const int slab_nk = (nk + nslabs - 1) / nslabs;
    const size_t slab_elems = (size_t)nb * 3 * slab_nk * ns;
    for (int i = 0; i < 2; ++i) OD_CHECK(od_buf_ensure(ctx, ctx->gslab[i], sizeof(double) * slab_elems));
    auto issue_copy = [&](int s) -> int {
      const int b = s & 1;
      const long long k0 = (long long)s * slab_nk;
      const int nks = (int)std::min<long long>(slab_nk, nk - k0);
      if (s >= 2) OD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));  // buffer free again?
      for (int is = 0; is < ns; ++is)  // band_gradient(nb,3,nk,ns) -> slab buffer (nb,3,nks,ns)
        OD_CUDA(ctx, cudaMemcpyAsync(ctx->gslab[b].as<double>() + (size_t)nb * 3 * nks * is,
                                     ctx->grads_host + (size_t)nb * 3 * (k0 + (size_t)nk * is), sizeof(double) * (size_t)nb * 3 * nks,
                                     cudaMemcpyHostToDevice, ctx->copy_stream));
      OD_CUDA(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
      return OD_OK;
    };
    OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // band_energy upload / memset ordered before the copy stream starts
    OD_CHECK(issue_copy(0));
    for (int s = 0; (long long)s * slab_nk < nk; ++s) {
      const int b = s & 1;
      const long long k0 = (long long)s * slab_nk;
      const int nks = (int)std::min<long long>(slab_nk, nk - k0);
      if ((long long)(s + 1) * slab_nk < nk) OD_CHECK(issue_copy(s + 1));
      OD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
      for (int t = 0; t < ntypes; ++t) OD_CHECK(one_slab(t, k0, nks, ctx->gslab[b].as<double>(), nks));
      OD_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], ctx->stream));
    }
  }

  for (int t = 0; t < ntypes; ++t) {
    double *acc = accum + stride * t;
      if (fixedwide[t]) {
      const int nw = weighted[t] ? norb : 0;
      std::vector<long long> out_off((size_t)(1 + nw) * ns), int_off((size_t)ns);
      for (int is = 0; is < ns; ++is) {
        out_off[is] = (long long)is * nbins;
        int_off[is] = (long long)(n1 + (size_t)is * nbins);
        for (int o = 0; o < nw; ++o) out_off[(size_t)(1 + o) * ns + is] = (long long)(2 * n1 + (size_t)nbins * (is + (size_t)ns * o));
      }
      OD_CHECK(od_fixedwide_finish(ctx, bspec[t], grid, 1 + nw, out_off.data(), int_off.data(), acc));
    }
    // rank-local epilogue, before the merge like the reference (quirk Q16)
    if (types[t] == 'l' && br->linear_smearing > 0.0) {
      double *tmp = accum + nall + 1;
      dim3 g((nbins + 127) / 128, ns);
      linear_smear_kernel<<<g, 128, 0, ctx->stream>>>(acc, tmp, nbins, ns, emin, delta_bins, br->linear_smearing);
      OD_LAUNCH_CHECK(ctx);
      OD_CUDA(ctx, cudaMemcpyAsync(acc, tmp, sizeof(double) * n1, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (types[t] != 'l' && br->numerical_intdos) {
      numerical_intdos_kernel<<<1, 32, 0, ctx->stream>>>(acc, acc + n1, nbins, ns, delta_bins);
      OD_LAUNCH_CHECK(ctx);
    }
  }
  return OD_OK;
  };
  const int rc_local = local();
  const std::string err_local = ctx->err;
  if (merge && ctx->nranks > 1) {
    if (rc_local != 0) {
      const double one = 1.0;
      cudaMemcpyAsync(accum + nall, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    }
    OD_CHECK(od_allreduce_device(ctx, accum, nall + 1, 's'));  // dos_utils_merge (dos_utils.f90:1041-1043), all schemes at once
  }
  if (rc_local != 0 && !(merge && ctx->nranks > 1 && host_out)) return rc_local;
  *accum_base = accum;
  *accum_stride = stride;
  if (host_out) {
    OD_CUDA(ctx, cudaMemcpyAsync(host_out, accum, sizeof(double) * nall, cudaMemcpyDeviceToHost, ctx->stream));
    double status = 0.0;
    OD_CUDA(ctx, cudaMemcpyAsync(&status, accum + nall, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    OD_CHECK(od_check_deferred(ctx));  // (synchronises)
    if (rc_local != 0) { ctx->err = err_local; return rc_local; }
    if (status != 0.0) return od_fail(ctx, OD_ERR_STATE, "calculate_dos failed on %d other rank(s) of the communicator", (int)status);
  }
  return OD_OK;
}

int od_calculate_dos_device(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins, const od_broadening *br,
                            const double *mw_dev, int norb, int mw_nb, int mw_nk, int merge, double **accum_out) {
  const bool weighted = mw_dev != nullptr;
  size_t stride = 0;
  return od_dos_accumulate(ctx, 1, &dos_type, &weighted, emin, delta_bins, nbins, br, mw_dev, norb, mw_nb, mw_nk, merge, accum_out, &stride, nullptr);
}

extern "C" int od_calculate_dos(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins, const od_broadening *br,
                                const double *matrix_weights, int mw_norb, int mw_nb, int mw_nk, int mw_mem, int merge,
                                double *dos, double *intdos, double *weighted_dos) {
  if (!ctx) return OD_ERR_ARG;
  if (!dos || !intdos) return od_fail(ctx, OD_ERR_ARG, "null output array");
  if (matrix_weights && !weighted_dos) return od_fail(ctx, OD_ERR_ARG, "matrix_weights given without weighted_dos");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const double *mw_dev = nullptr;
  if (matrix_weights) {
    if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
    const size_t bytes = sizeof(double) * (size_t)mw_norb * mw_nb * mw_nk * ctx->ns;
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, matrix_weights, bytes, mw_mem, ctx->mw_tmp, &d));
    mw_dev = static_cast<const double *>(d);
  }
  double *accum = nullptr;
  OD_CHECK(od_calculate_dos_device(ctx, dos_type, emin, delta_bins, nbins, br, mw_dev, mw_norb, mw_nb, mw_nk, merge, &accum));
  const size_t n1 = (size_t)nbins * ctx->ns;
  OD_CUDA(ctx, cudaMemcpyAsync(dos, accum, sizeof(double) * n1, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaMemcpyAsync(intdos, accum + n1, sizeof(double) * n1, cudaMemcpyDeviceToHost, ctx->stream));
  if (matrix_weights)
    OD_CUDA(ctx, cudaMemcpyAsync(weighted_dos, accum + 2 * n1, sizeof(double) * n1 * mw_norb, cudaMemcpyDeviceToHost, ctx->stream));
  return od_check_deferred(ctx);  // synchronises; doslin `stop` conditions met by the wide-window pass surface here
}

// ------------------------------------------------------------------ calculate_dos_at_e
namespace {
// One warp walks a contiguous slice of units; lane 0..31 own orbitals o = lane, lane+32, ... (<= 8 each)
// and every lane carries the same dos_at_e partial.  Per-warp partials are summed on the host in order.
__global__ void __launch_bounds__(128)
dos_at_e_kernel(DosProducer prod, double energy, long long units_per_warp, double *__restrict__ part /* [nwarps][ns][1+norb] */,
                int *__restrict__ bad_flag) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int is = blockIdx.y;
  const long long u0 = warp * units_per_warp;
  long long u1 = u0 + units_per_warp;
  if (u1 > prod.units_per_z) u1 = prod.units_per_z;
  double acc = 0.0, accw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int bad = 0;
  const int norb = prod.mw ? prod.norb : 0;
  for (long long base = u0; base < u1; base += 32) {
    // each lane builds one record, then the warp evaluates the 32 of them in turn
    RecOut r;
    r.mode = -1;
    DosProducer p2 = prod;
    p2.mw = nullptr;
    if (base + lane < u1) p2(base + lane, is, r);
    double d = 0.0, cum;
    if (r.mode == 0) d = od_gaussian(r.e0, r.p1, energy) * r.omega;                       // dos_utils.f90:1777
    else if (r.mode == 1) d = od_doslin(r.e0, r.p1, r.p2, r.p3, r.p4, energy, &cum, &bad) * r.omega;  // :1774
    // lane-ordered sum (fixed order -> reproducible)
    for (int l = 0; l < 32; ++l) {
      const double dl = __shfl_sync(0xffffffffu, d, l);
      acc += dl;
      if (norb > 0 && dl != 0.0) {
        const long long u = base + l;
        if (u < u1) {
          const int ib = (int)(u % prod.nb);
          const long long ik = u / prod.nb;
          if (ik < prod.mw_nk && ib < prod.mw_nb) {
            const size_t row = (size_t)norb * (ib + (size_t)prod.mw_nb * (ik + (size_t)prod.mw_nk * is));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int o = lane + 32 * k;
              if (o < norb) accw[k] += dl * prod.mw[row + o];                              // :1787
            }
          }
        }
      }
    }
  }
  if (bad) atomicOr(bad_flag, bad);
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  (void)nwarps;
  double *out = part + ((size_t)warp * gridDim.y + is) * (1 + norb);
  if (lane == 0) out[0] = acc;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int o = lane + 32 * k;
    if (o < norb) out[1 + o] = accw[k];
  }
}
}  // namespace

extern "C" int od_calculate_dos_at_e(od_ctx *ctx, char dos_type, double energy, double delta_bins, const od_broadening *br,
                                     const double *matrix_weights, int mw_norb, int mw_nb, int mw_nk, int mw_mem, int merge,
                                     double *dos_at_e, double *weighted_dos_at_e) {
  if (ctx) OD_CHECK(od_grads_resident(ctx));
  OD_CHECK(check_ready(ctx, dos_type != 'f'));
  if (!dos_at_e) return od_fail(ctx, OD_ERR_ARG, "null output array");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ns = ctx->ns;
  int norb = 0;
  const double *mw_dev = nullptr;
  if (matrix_weights && weighted_dos_at_e) {
    if (mw_norb < 1 || mw_norb > 256) return od_fail(ctx, OD_ERR_ARG, "calculate_dos_at_e supports 1..256 weight columns, got %d", mw_norb);
    if (mw_nk != ctx->nk) return od_fail(ctx, OD_ERR_REF, "ERROR : DOS : mw%%nkpoints not equal to nkpoints. (dos_utils.f90:1587)");
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, matrix_weights, sizeof(double) * (size_t)mw_norb * mw_nb * mw_nk * ns, mw_mem, ctx->mw_tmp, &d));
    mw_dev = static_cast<const double *>(d);
    norb = mw_norb;
  }
  DosProducer p;
  p.units_per_z = (long long)ctx->nb * ctx->nk;
  p.k_first = 0;
  p.nb = ctx->nb; p.ns = ns; p.nk = ctx->nk; p.nk_g = ctx->nk;
  p.bands = ctx->bands.as<double>(); p.grads = ctx->grads.as<double>(); p.kw = ctx->kweight.as<double>();
  p.eps = ctx->eps;
  OD_CHECK(od_make_broadspec(ctx, dos_type, br, delta_bins, /*honor_hybrid=*/true, &p.b));
  p.mw = mw_dev; p.norb = norb; p.mw_nb = mw_nb; p.mw_nk = mw_nk; p.og = 0;

  const long long units = p.units_per_z;
  long long nwarps = std::min<long long>((units + 31) / 32, (long long)ctx->sm_count * 16);
  if (nwarps < 1) nwarps = 1;
  long long upw = (units + nwarps - 1) / nwarps;
  upw = (upw + 31) / 32 * 32;
  nwarps = (units + upw - 1) / upw;
  const int blocks = (int)((nwarps + 3) / 4);
  const long long nwarps_launched = (long long)blocks * 4;
  const size_t per = (size_t)(1 + norb);
  const size_t nvals = (size_t)nwarps_launched * ns * per;
  OD_CHECK(od_buf_ensure(ctx, ctx->partial, sizeof(double) * nvals + 16));
  OD_CUDA(ctx, cudaMemsetAsync(ctx->partial.p, 0, sizeof(double) * nvals + 16, ctx->stream));
  int *bad = reinterpret_cast<int *>(ctx->partial.as<char>() + sizeof(double) * nvals);
  dim3 g(blocks, ns);
  dos_at_e_kernel<<<g, 128, 0, ctx->stream>>>(p, energy, upw, ctx->partial.as<double>(), bad);
  OD_LAUNCH_CHECK(ctx);
  std::vector<double> h(nvals);
  int hbad = 0;
  OD_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->partial.p, sizeof(double) * nvals, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (hbad) return od_fail(ctx, OD_ERR_DOSLIN, "doslin: input arguments incorrect (dos_utils.f90:1431/1514)");
  std::vector<double> res((size_t)ns * per, 0.0);
  for (long long w = 0; w < nwarps_launched; ++w)
    for (int is = 0; is < ns; ++is)
      for (size_t k = 0; k < per; ++k) res[(size_t)is * per + k] += h[((size_t)w * ns + is) * per + k];
  if (merge) OD_CHECK(od_allreduce(ctx, res.data(), res.size(), 's'));  // dos_utils_merge_at_e (dos_utils.f90:1830-1832)
  for (int is = 0; is < ns; ++is) {
    dos_at_e[is] = res[(size_t)is * per];
    for (int o = 0; o < norb; ++o) weighted_dos_at_e[is + (size_t)ns * o] = res[(size_t)is * per + 1 + o];
  }
  return OD_OK;
}

// ------------------------------------------------------------------ calculate_jdos
namespace {
// flags[ib] = band ib lies below efermi at some (k, spin); flags[nb + ib] = at or above it somewhere
__global__ void band_occupancy_kernel(const double *__restrict__ bands, int nb, size_t nstates, double efermi, int *__restrict__ flags) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nstates; i += (size_t)gridDim.x * blockDim.x) {
    const int ib = (int)(i % (size_t)nb);
    if (bands[i] < efermi) { if (!flags[ib]) flags[ib] = 1; }
    else if (!flags[nb + ib]) flags[nb + ib] = 1;
  }
}
}  // namespace

int od_calculate_jdos_device(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br, double efermi,
                             double scissor_op, const int *exclude_bands, int num_exclude_bands, const double *mw_dev,
                             const double *ome_dev, const OpticsCoef *coef, int n_geom, int merge, double **accum_out) {
  if (ctx) OD_CHECK(od_grads_resident(ctx));
  OD_CHECK(check_ready(ctx, jdos_type != 'f'));
  if (nbins < 2) return od_fail(ctx, OD_ERR_ARG, "nbins = %d", nbins);
  if (!(delta_bins > 0.0)) return od_fail(ctx, OD_ERR_ARG, "delta_bins must be positive");
  if ((mw_dev || ome_dev) && n_geom != 1 && n_geom != 6) return od_fail(ctx, OD_ERR_ARG, "N_geom must be 1 or 6, got %d", n_geom);
  if (!mw_dev && !ome_dev) n_geom = 0;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ns = ctx->ns;

  JdosProducer p;
  p.units_per_z = (long long)ctx->nb * ctx->nb * ctx->nk;
  p.k_first = 0;
  p.nb = ctx->nb; p.ns = ns; p.nk = ctx->nk;
  p.bands = ctx->bands.as<double>(); p.grads = ctx->grads.as<double>(); p.kw = ctx->kweight.as<double>();
  p.eps = ctx->eps; p.efermi = efermi; p.scissor = scissor_op;
  OD_CHECK(od_make_broadspec(ctx, jdos_type, br, delta_bins, /*honor_hybrid=*/true, &p.b));
  p.exclude = nullptr; p.n_exclude = 0;
  if (num_exclude_bands > 0) {
    if (!exclude_bands) return od_fail(ctx, OD_ERR_ARG, "null exclude_bands");
    OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(int) * (size_t)num_exclude_bands));
    OD_CUDA(ctx, cudaMemcpyAsync(ctx->small.p, exclude_bands, sizeof(int) * (size_t)num_exclude_bands, cudaMemcpyHostToDevice, ctx->stream));
    p.exclude = ctx->small.as<int>();
    p.n_exclude = num_exclude_bands;
  }
  p.mw = mw_dev; p.ome = ome_dev; p.n_geom = n_geom;
  p.occ_list = p.unocc_list = nullptr; p.n_occ = p.n_unocc = 0; p.eager = 0;
  if (coef) p.coef = *coef; else memset(&p.coef, 0, sizeof(p.coef));

  GridSpec grid{0.0, delta_bins, nbins};  // E(i) = (i-1)*delta_bins (jdos_utils.f90:217)
  const size_t n1 = (size_t)nbins * ns;
  const size_t ntot = n1 + n1 * n_geom;
  OD_CHECK(od_buf_ensure(ctx, ctx->accum, sizeof(double) * (ntot + 1)));  // + the status word of the merge
  double *accum = ctx->accum.as<double>();
  OD_CUDA(ctx, cudaMemsetAsync(accum, 0, sizeof(double) * (ntot + 1), ctx->stream));
  ctx->merge_status_dev = nullptr;
  // everything rank-local runs inside this lambda: a failure on one rank travels in the status word of the allreduce and
  // becomes an error on every rank (od_merge_status_check) instead of a hang in the collective
  auto local = [&]() -> int {
  const int nacc = 1 + n_geom;
  std::vector<long long> dest((size_t)ns * nacc);
  for (int is = 0; is < ns; ++is) {
    dest[(size_t)is * nacc] = (long long)is * nbins;
    for (int a = 0; a < n_geom; ++a) dest[(size_t)is * nacc + 1 + a] = (long long)(n1 + (size_t)nbins * (is + (size_t)ns * a));
  }
  // the exclude list lives in ctx->small, which the narrow engine also uses: it gets its own copy.  ONE allocation of
  // the final size first (od_buf_ensure does not preserve contents when it grows a buffer), then the copy.
  if (ctx->engine != OD_ENGINE_DENSE) {
    const int nb = ctx->nb;
    const size_t excl_bytes = (sizeof(int) * (size_t)num_exclude_bands + 15) & ~(size_t)15;
    OD_CHECK(od_buf_ensure(ctx, ctx->in_tmp2, excl_bytes + sizeof(int) * 4 * (size_t)nb));
    if (num_exclude_bands > 0) {
      OD_CUDA(ctx, cudaMemcpyAsync(ctx->in_tmp2.p, exclude_bands, sizeof(int) * (size_t)num_exclude_bands, cudaMemcpyHostToDevice, ctx->stream));
      p.exclude = ctx->in_tmp2.as<int>();
    }
    // candidate band lists: a pair (ib, jb) needs E_ib < efermi and E_jb >= efermi at the same (k, spin)
    int *flags = reinterpret_cast<int *>(ctx->in_tmp2.as<char>() + excl_bytes);
    OD_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int) * 2 * (size_t)nb, ctx->stream));
    const size_t nstates = (size_t)nb * ns * ctx->nk;
    if (nstates > 0) {
      band_occupancy_kernel<<<(unsigned)std::min<size_t>((nstates + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(p.bands, nb, nstates, efermi, flags);
      OD_LAUNCH_CHECK(ctx);
    }
    std::vector<int> hflags(2 * (size_t)nb), lists;
    OD_CUDA(ctx, cudaMemcpyAsync(hflags.data(), flags, sizeof(int) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
    OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int ib = 0; ib < nb; ++ib) {
      bool excluded = false;
      for (int x = 0; x < num_exclude_bands; ++x) excluded = excluded || (exclude_bands[x] == ib + 1);  // jdos_utils.f90:355
      if (hflags[ib] && !excluded) lists.push_back(ib);
    }
    p.n_occ = (int)lists.size();
    for (int jb = 0; jb < nb; ++jb)
      if (hflags[nb + jb]) lists.push_back(jb);
    p.n_unocc = (int)lists.size() - p.n_occ;
    if (p.n_occ > 0 && p.n_unocc > 0) {
      int *dl = flags + 2 * (size_t)nb;
      OD_CUDA(ctx, cudaMemcpyAsync(dl, lists.data(), sizeof(int) * lists.size(), cudaMemcpyHostToDevice, ctx->stream));
      OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // lists is a host vector
      p.occ_list = dl;
      p.unocc_list = dl + p.n_occ;
      p.units_per_z = p.pairs_per_k() * ctx->nk;
      if (od_fixedwide_applies(ctx, p.b, grid, OD_SQRT60)) {
        // fixed broadening with windows beyond the narrow engine: moment expansion (od_fixedwide.cu), in k-point
        // slabs of < 2^30 pair slots
        OD_CHECK(od_fixedwide_begin(ctx, p.b, grid, 1 + n_geom));
        const long long ppk = p.pairs_per_k(), slab_k = std::max<long long>(1, (1ll << 30) / ppk);
        for (long long k0 = 0; k0 < ctx->nk; k0 += slab_k) {
          JdosProducer ps = p;
          ps.k_first = k0;
          ps.units_per_z = std::min<long long>(slab_k, ctx->nk - k0) * ppk;
          OD_CHECK(od_fixedwide_accumulate_jdos(ctx, ps, grid, n_geom));
        }
        std::vector<long long> out_off((size_t)(1 + n_geom) * ns);
        for (int is = 0; is < ns; ++is) {
          out_off[is] = (long long)is * nbins;
          for (int a = 0; a < n_geom; ++a) out_off[(size_t)(1 + a) * ns + is] = (long long)(n1 + (size_t)nbins * (is + (size_t)ns * a));
        }
        OD_CHECK(od_fixedwide_finish(ctx, p.b, grid, 1 + n_geom, out_off.data(), nullptr, accum));
      } else if (od_narrow_grid_fits(nbins, ns)) {
        OD_CHECK(od_narrow_jdos(ctx, p, grid, jdos_type == 'l', n_geom, accum));
      } else {  // a grid beyond the narrow engine's bucket table: the dense engine over all nb x nb slots
        p.occ_list = p.unocc_list = nullptr;
        p.units_per_z = (long long)nb * nb * ctx->nk;
        if (n_geom == 0) OD_CHECK((od_run_dense<0, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
        else if (n_geom == 1) OD_CHECK((od_run_dense<1, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
        else OD_CHECK((od_run_dense<6, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
      }
    }  // else: no pair anywhere, the spectra stay zero
  } else if (n_geom == 0) OD_CHECK((od_run_dense<0, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
  else if (n_geom == 1) OD_CHECK((od_run_dense<1, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
  else OD_CHECK((od_run_dense<6, false, JdosProducer>(ctx, p, grid, ns, dest, accum)));
  return OD_OK;
  };
  const int rc_local = local();
  if (merge && ctx->nranks > 1) {
    const std::string err_local = ctx->err;
    if (rc_local != 0) {
      const double one = 1.0;
      cudaMemcpyAsync(accum + ntot, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    }
    OD_CHECK(od_allreduce_device(ctx, accum, ntot + 1, 's'));  // jdos_utils_merge (jdos_utils.f90:435-437)
    ctx->merge_status_dev = accum + ntot;
    ctx->err = err_local;
  }
  if (rc_local != 0) return rc_local;
  *accum_out = accum;
  return OD_OK;
}

// after a merge: non-zero if the calculation failed on another rank (synchronises)
int od_merge_status_check(od_ctx *ctx) {
  if (!ctx->merge_status_dev) return OD_OK;
  double status = 0.0;
  OD_CUDA(ctx, cudaMemcpyAsync(&status, ctx->merge_status_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->merge_status_dev = nullptr;
  if (status != 0.0) return od_fail(ctx, OD_ERR_STATE, "the calculation failed on %d other rank(s) of the communicator", (int)status);
  return OD_OK;
}

static int jdos_copy_out(od_ctx *ctx, const double *accum, int nbins, int n_geom, double *jdos, double *wjdos) {
  const size_t n1 = (size_t)nbins * ctx->ns;
  OD_CUDA(ctx, cudaMemcpyAsync(jdos, accum, sizeof(double) * n1, cudaMemcpyDeviceToHost, ctx->stream));
  if (wjdos && n_geom > 0)
    OD_CUDA(ctx, cudaMemcpyAsync(wjdos, accum + n1, sizeof(double) * n1 * n_geom, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CHECK(od_merge_status_check(ctx));
  return od_check_deferred(ctx);  // synchronises
}

extern "C" int od_calculate_jdos(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br, double efermi,
                                 double scissor_op, const int *exclude_bands, int num_exclude_bands, const double *matrix_weights,
                                 int n_geom, int mw_mem, int merge, double *jdos, double *weighted_jdos) {
  if (!ctx) return OD_ERR_ARG;
  if (!jdos) return od_fail(ctx, OD_ERR_ARG, "null output array");
  if (matrix_weights && !weighted_jdos) return od_fail(ctx, OD_ERR_ARG, "matrix_weights given without weighted_jdos");
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const double *mw_dev = nullptr;
  if (matrix_weights) {
    const size_t bytes = sizeof(double) * (size_t)ctx->nb * ctx->nb * ctx->nk * ctx->ns * n_geom;
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, matrix_weights, bytes, mw_mem, ctx->mw_tmp, &d));
    mw_dev = static_cast<const double *>(d);
  }
  double *accum = nullptr;
  OD_CHECK(od_calculate_jdos_device(ctx, jdos_type, delta_bins, nbins, br, efermi, scissor_op, exclude_bands, num_exclude_bands,
                                    mw_dev, nullptr, nullptr, n_geom, merge, &accum));
  return jdos_copy_out(ctx, accum, nbins, matrix_weights ? n_geom : 0, jdos, weighted_jdos);
}

extern "C" int od_calculate_jdos_optics(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br,
                                        double efermi, double scissor_op, const int *exclude_bands, int num_exclude_bands,
                                        const double *optical_mat, int ome_mem, int geom, const double qdir[3],
                                        const double *symops, int num_symm, int merge, double *jdos, double *weighted_jdos) {
  if (!ctx) return OD_ERR_ARG;
  if (!jdos || !weighted_jdos || !optical_mat) return od_fail(ctx, OD_ERR_ARG, "null array");
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OpticsCoef coef;
  OD_CHECK(od_optics_coefficients(ctx, geom, qdir, symops, num_symm, &coef));
  const size_t bytes = sizeof(double) * 2 * (size_t)ctx->nb * ctx->nb * 3 * ctx->nk * ctx->ns;
  const void *d = nullptr;
  OD_CHECK(od_stage_in(ctx, optical_mat, bytes, ome_mem, ctx->in_tmp, &d));
  double *accum = nullptr;
  OD_CHECK(od_calculate_jdos_device(ctx, jdos_type, delta_bins, nbins, br, efermi, scissor_op, exclude_bands, num_exclude_bands,
                                    nullptr, static_cast<const double *>(d), &coef, coef.n_geom, merge, &accum));
  return jdos_copy_out(ctx, accum, nbins, coef.n_geom, jdos, weighted_jdos);
}

// od_analysis.cu -- the analysis routines of od_dos_utils that follow the Fermi level:
//   dos_utils_compute_bandgap (dos_utils.f90:457-750) and dos_utils_compute_band_energies (:853-924, with
//   calc_band_energies :753-795).  The O(nk nb) passes over band_energy run on the GPU; what the reference does
//   on the root node alone (a sequential scan of the (vbm, cbm) slices whose tie counting depends on the order of
//   the k-points, :606-646) stays a host loop over 2 ns nk doubles so that the multiplicities and the k-point of
//   the gap come out exactly as the reference prints them.
#include <cfloat>
#include <cmath>

#include "od_internal.h"

namespace {
// One warp per (k, spin): the first band above efermi is the CBM, the one below it the VBM (dos_utils.f90:507-519).
// 32 consecutive bands per load (coalesced), ballot for the first hit.  slice(1:2, is, ik) as in the reference,
// -200 where no band lies above efermi (:504).
__global__ void __launch_bounds__(256) band_edges_kernel(const double *__restrict__ bands, int nb, int ns, long long nk, double efermi,
                                                         double *__restrict__ slice) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < nk * ns; t += nwarps) {
    const double *col = bands + (size_t)nb * t;  // band_energy(:, is, ik), t = is + ns * ik
    double vbm = -200.0, cbm = -200.0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int ib = b0 + lane;
      const double e = ib < nb ? col[ib] : 0.0;
      const unsigned hit = __ballot_sync(0xffffffffu, ib < nb && e > efermi);
      if (hit) {
        const int first = b0 + __ffs(hit) - 1;
        cbm = col[first];
        if (first > 0) vbm = col[first - 1];  // first == 0: band_energy(0, ...) in the reference (out of bounds)
        break;
      }
    }
    if (lane == 0) { slice[2 * t] = vbm; slice[2 * t + 1] = cbm; }
  }
}

// sum over the rank's states with E <= efermi of E * electrons_per_state * kpoint_weight(ik) (dos_utils.f90:900-908);
// per-CTA partial sums in a fixed order -> reproducible
__global__ void __launch_bounds__(256) band_energy_sum_kernel(const double *__restrict__ bands, const double *__restrict__ kw, int nb,
                                                              int ns, long long nk, double efermi, double eps,
                                                              double *__restrict__ part) {
  const size_t total = (size_t)nb * ns * nk, chunk = (total + gridDim.x - 1) / gridDim.x;
  const size_t i0 = chunk * blockIdx.x, i1 = i0 + chunk < total ? i0 + chunk : total;
  double s = 0.0;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const double e = bands[i];
    if (e <= efermi) s += e * eps * kw[i / ((size_t)nb * ns)];
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
}  // namespace

extern "C" int od_dos_utils_compute_bandgap(od_ctx *ctx, double efermi, const double *num_electrons, long long nk_global,
                                            long long k_first, od_bandgap_result *res) {
  if (!ctx || !res) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ns = ctx->ns, nb = ctx->nb;
  const long long nk = ctx->nk;
  if (ctx->nranks == 1) { nk_global = nk; k_first = 0; }
  if (nk_global < nk || k_first < 0 || k_first + nk > nk_global)
    return od_fail(ctx, OD_ERR_ARG, "compute_bandgap: k-points [%lld, %lld) outside the global set of %lld", k_first, k_first + nk, nk_global);
  if (ns > 1 && !num_electrons) return od_fail(ctx, OD_ERR_ARG, "compute_bandgap: num_electrons needed for the weighted average");
  // global_bandgap(1:2, 1:nspins, 1:nkpoints) (:522-544): every rank writes its block, zeros elsewhere, sum-allreduce
  // (adding zeros is exact) -- the same k-point numbering whatever the number of ranks, as the reference insists
  const size_t n = 2 * (size_t)ns * nk_global;
  OD_CHECK(od_buf_ensure(ctx, ctx->partial, sizeof(double) * n));
  double *slice = ctx->partial.as<double>();
  if (ctx->nranks > 1) OD_CUDA(ctx, cudaMemsetAsync(slice, 0, sizeof(double) * n, ctx->stream));
  const int blocks = (int)std::min<long long>((long long)ctx->sm_count * 8, (nk * ns + 7) / 8);
  band_edges_kernel<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(ctx->bands.as<double>(), nb, ns, nk, efermi,
                                                                   slice + 2 * (size_t)ns * k_first);
  OD_LAUNCH_CHECK(ctx);
  OD_CHECK(od_allreduce_device(ctx, slice, n, 's'));
  std::vector<double> g(n);
  OD_CUDA(ctx, cudaMemcpyAsync(g.data(), slice, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

  // :590-653, the root's scan: spins outermost, ONE running thermal_vbm / thermal_cbm shared by both spins
  memset(res, 0, sizeof(*res));
  double thermal_cbm = DBL_MAX, thermal_vbm = -DBL_MAX;
  res->thermal_vbm_k = res->thermal_cbm_k = -1;
  for (int is = 0; is < 2; ++is) {
    res->vbm_multiplicity[is] = res->cbm_multiplicity[is] = res->optical_multiplicity[is] = 1;
    res->optical_bandgap[is] = is < ns ? DBL_MAX : 0.0;
  }
  for (int is = 0; is < ns; ++is)
    for (long long ik = 0; ik < nk_global; ++ik) {
      const double v = g[2 * ((size_t)is + (size_t)ns * ik)], c = g[2 * ((size_t)is + (size_t)ns * ik) + 1];
      if (v >= thermal_vbm) {
        if (std::fabs(v - thermal_vbm) < DBL_EPSILON) res->vbm_multiplicity[is] += 1;
        else { thermal_vbm = v; res->vbm_multiplicity[is] = 1; res->thermal_vbm_k = ik + 1; }
      }
      if (c <= thermal_cbm) {
        if (std::fabs(c - thermal_cbm) < DBL_EPSILON) res->cbm_multiplicity[is] += 1;
        else { thermal_cbm = c; res->cbm_multiplicity[is] = 1; res->thermal_cbm_k = ik + 1; }
      }
      const double kgap = c - v;
      if (kgap <= res->optical_bandgap[is]) {
        if (std::fabs(kgap - res->optical_bandgap[is]) < DBL_EPSILON) res->optical_multiplicity[is] += 1;
        else { res->optical_bandgap[is] = kgap; res->optical_multiplicity[is] = 1; }
      }
      res->average_bandgap[is] = res->average_bandgap[is] + (c - v);
    }
  for (int is = 0; is < ns; ++is) res->average_bandgap[is] = res->average_bandgap[is] / (double)nk_global;
  res->vbm_energy = thermal_vbm;
  res->cbm_energy = thermal_cbm;
  res->thermal_bandgap = thermal_cbm - thermal_vbm;
  for (int is = 0; is < ns; ++is)
    if (res->vbm_multiplicity[is] != 1 || res->cbm_multiplicity[is] != 1) res->thermal_multiplicity = 1;
  if (ns > 1)
    res->weighted_average = (res->average_bandgap[0] * num_electrons[0] + res->average_bandgap[1] * num_electrons[1]) /
                            (num_electrons[0] + num_electrons[1]);
  ctx->cbm_energy = thermal_cbm;  // the module variables of od_dos_utils (dos_utils.f90:55-56) that write_core reads
  ctx->vbm_energy = thermal_vbm;
  return OD_OK;
}

extern "C" int od_dos_utils_compute_band_energies(od_ctx *ctx, double efermi, double band_energy_dos[3], double *band_energy_castep) {
  if (!ctx || !band_energy_dos || !band_energy_castep) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p || !ctx->kweight.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const Spectra &S = ctx->dos_res;
  // calc_band_energies (:753-795) of the spectra of the last od_dos_utils_calculate (zero for schemes not enabled)
  for (int t = 0; t < 3; ++t) {
    double gband = 0.0;
    if (S.have[t])
      for (int is = 0; is < S.ns; ++is)
        for (int i = 0; i < S.nbins; ++i) {
          const double E = S.emin + (double)i * S.delta;
          if (E <= efermi) gband = gband + S.dos[t][(size_t)i + (size_t)S.nbins * is] * E * S.delta;
          else break;
        }
    band_energy_dos[t] = gband;
  }
  const int blocks = ctx->sm_count * 4;
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(double) * (size_t)blocks));
  band_energy_sum_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bands.as<double>(), ctx->kweight.as<double>(), ctx->nb, ctx->ns, ctx->nk,
                                                          efermi, ctx->eps, ctx->small.as<double>());
  OD_LAUNCH_CHECK(ctx);
  std::vector<double> h((size_t)blocks);
  OD_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->small.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double eband = 0.0;
  for (double v : h) eband += v;
  OD_CHECK(od_allreduce(ctx, &eband, 1, 's'));  // comms_reduce(eband, 1, 'SUM') (:909)
  *band_energy_castep = eband;
  return OD_OK;
}

// od_drivers.cu -- host-side mirror of the reference's L3 task drivers above the kernels:
//   dos_utils_calculate (dos_utils.f90:92-293), setup_energy_scale (:927-1010),
//   calc_efermi_from_intdos (:798-850), dos_utils_set_efermi (:296-388),
//   jdos_utils_calculate (jdos_utils.f90:61-174) and its setup_energy_scale (:177-231).
// Same control flow, flags and quirks (sticky dos_nbins, "most complex" weighted broadening, ...);
// the module-global result arrays live in the ctx and are read back with od_dos_get / od_jdos_get.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "od_internal.h"
#include "od_spectral.h"

extern "C" void od_dos_params_defaults(od_dos_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->adaptive = 1;  // parameters.f90:232-234 (broadening : adaptive)
  od_broadening_defaults(&p->br);
  p->dos_spacing = (double)0.005f;  // parameters.f90:459-461: single-precision literal
  p->dos_nbins = -1;                // parameters.f90:307
  p->efermi_choice = OD_EFERMI_OPTADOS;  // parameters.f90:257
  p->nspins_electrons = 1;
}

extern "C" void od_jdos_params_defaults(od_jdos_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->adaptive = 1;
  od_broadening_defaults(&p->br);
  p->jdos_spacing = 0.01;      // parameters.f90:321
  p->jdos_max_energy = -1.0;   // parameters.f90:318
  p->scissor_op = 0.0;         // parameters.f90:351
}

// dos_utils.f90:798-850 (E(i) = emin + (i-1)*delta_bins)
extern "C" double od_calc_efermi_from_intdos(const double *intdos, double emin, double delta_bins, int nbins, int ns,
                                             const double *num_electrons) {
  const double tolerance = 0.0001;
  double nel = 0.0;
  for (int is = 0; is < ns; ++is) nel += num_electrons[is];
  auto E = [&](int i1) { return emin + (double)(i1 - 1) * delta_bins; };
  auto S = [&](int i1) {
    double s = 0.0;
    for (int is = 0; is < ns; ++is) s += intdos[(size_t)(i1 - 1) + (size_t)nbins * is];
    return s;
  };
  int idos;
  for (idos = 1; idos <= nbins; ++idos)
    if (S(idos) > (nel + (tolerance / 2.0))) break;
  if (idos > nbins) idos = nbins;  // the reference would read E(nbins+1); clamp
  const double efermi = E(idos);
  const double sref = (idos >= 2) ? S(idos - 1) : 0.0;
  int i;
  for (i = idos; i >= 1; --i)
    if (S(i) < (sref - (tolerance / 2.0))) break;
  if (i < 1) i = 1;  // E(0) is out of bounds in the reference; clamp
  return (efermi + E(i)) / 2.0;
}

// dos_utils.f90:927-1010
static int setup_energy_scale(od_ctx *ctx, od_dos_params *p, double *emin, double *delta, int *nbins) {
  double mn, mx;
  OD_CHECK(od_band_minmax(ctx, &mn, &mx));
  const double min_band_energy = p->have_dos_min_energy ? p->dos_min_energy : mn - 5.0;
  double max_band_energy = p->have_dos_max_energy ? p->dos_max_energy : mx + 5.0;
  if (p->dos_nbins < 0) {
    if (!(p->dos_spacing > 0.0)) return od_fail(ctx, OD_ERR_ARG, "dos_spacing must be positive");
    p->dos_nbins = std::abs((int)std::ceil((max_band_energy - min_band_energy) / p->dos_spacing));
    max_band_energy = min_band_energy + p->dos_nbins * p->dos_spacing;
  }
  if (p->dos_nbins < 2) return od_fail(ctx, OD_ERR_ARG, "dos_nbins = %d", p->dos_nbins);
  *delta = (max_band_energy - min_band_energy) / (double)(p->dos_nbins - 1);
  *emin = min_band_energy;
  *nbins = p->dos_nbins;
  return OD_OK;
}

static int pull(od_ctx *ctx, std::vector<double> &dst, const double *dev, size_t n) {
  dst.resize(n);
  OD_CUDA(ctx, cudaMemcpyAsync(dst.data(), dev, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

extern "C" int od_dos_utils_calculate(od_ctx *ctx, od_dos_params *p, const double *matrix_weights, int mw_norb, int mw_nb,
                                      int mw_nk, int mw_mem, od_dos_result *res) {
  if (!ctx || !p || !res) return OD_ERR_ARG;
  if (!(p->linear || p->adaptive || p->fixed)) return od_fail(ctx, OD_ERR_REF, " DOS: No Broadening Set (dos_utils.f90:136)");
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool calc_weighted = matrix_weights != nullptr;
  if (calc_weighted && mw_nk != ctx->nk)
    return od_fail(ctx, OD_ERR_REF, "ERROR : DOS : mw%%nkpoints not equal to nkpoints. (dos_utils.f90:159)");
  const double *mw_dev = nullptr;
  if (calc_weighted) {
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, matrix_weights, sizeof(double) * (size_t)mw_norb * mw_nb * mw_nk * ctx->ns, mw_mem, ctx->mw_tmp, &d));
    mw_dev = static_cast<const double *>(d);
  }
  memset(res, 0, sizeof(*res));
  double emin, delta;
  int nbins;
  OD_CHECK(setup_energy_scale(ctx, p, &emin, &delta, &nbins));
  const int ns = ctx->ns;
  const size_t n1 = (size_t)nbins * ns;
  Spectra &S = ctx->dos_res;
  S = Spectra();
  S.nbins = nbins;
  S.ns = ns;
  S.emin = emin;
  S.delta = delta;
  const char types[3] = {'f', 'a', 'l'};
  const int enabled[3] = {p->fixed, p->adaptive, p->linear};
  // weighted spectrum only with the most complex enabled scheme (dos_utils.f90:179-208)
  const bool wt[3] = {calc_weighted && !p->adaptive && !p->linear, calc_weighted && !p->linear, calc_weighted};
  // all enabled schemes in one sweep over the k-points (one upload of the gradients when they are streamed)
  char tl[3];
  bool wl[3];
  int idx[3], nt = 0;
  for (int t = 0; t < 3; ++t)
    if (enabled[t]) { tl[nt] = types[t]; wl[nt] = wt[t]; idx[nt] = t; ++nt; }
  double *accum = nullptr;
  size_t stride = 0;
  // one merge, one device -> host transfer and one synchronisation for all schemes
  const size_t per = 2 * n1 + (calc_weighted ? n1 * (size_t)mw_norb : 0);
  if (ctx->host_stage_bytes < sizeof(double) * per * nt) {
    if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
    ctx->host_stage = nullptr;
    ctx->host_stage_bytes = 0;
    OD_CUDA(ctx, cudaHostAlloc(&ctx->host_stage, sizeof(double) * per * nt, cudaHostAllocDefault));
    ctx->host_stage_bytes = sizeof(double) * per * nt;
  }
  double *hs = static_cast<double *>(ctx->host_stage);
  OD_CHECK(od_dos_accumulate(ctx, nt, tl, wl, emin, delta, nbins, &p->br, mw_dev, mw_norb, mw_nb, mw_nk, /*merge=*/1, &accum, &stride, hs));
  if (stride != per) return od_fail(ctx, OD_ERR_STATE, "internal: spectrum block of %zu doubles, expected %zu", stride, per);
  for (int i = 0; i < nt; ++i) {
    const int t = idx[i];
    const double *acc = hs + stride * i;
    S.dos[t].assign(acc, acc + n1);
    S.intdos[t].assign(acc + n1, acc + 2 * n1);
    S.have[t] = true;
    if (wt[t]) {
      S.weighted.assign(acc + 2 * n1, acc + 2 * n1 + n1 * mw_norb);
      S.have_weighted = true;
      S.norb = mw_norb;
    }
  }
  res->nbins = nbins;
  res->emin = emin;
  res->delta_bins = delta;
  res->have_fixed = S.have[0];
  res->have_adaptive = S.have[1];
  res->have_linear = S.have[2];
  res->have_weighted = S.have_weighted;
  res->weighted_norb = S.norb;
  // Fermi energy analysis (dos_utils.f90:231-263)
  if (p->efermi_choice == OD_EFERMI_OPTADOS) {
    const int nse = p->nspins_electrons > 0 ? p->nspins_electrons : ns;
    double nel[2] = {p->num_electrons[0], nse > 1 ? p->num_electrons[1] : 0.0};
    if (nse < ns) nel[1] = 0.0;
    if (S.have[0]) res->efermi_fixed = od_calc_efermi_from_intdos(S.intdos[0].data(), emin, delta, nbins, ns, nel);
    if (S.have[1]) res->efermi_adaptive = od_calc_efermi_from_intdos(S.intdos[1].data(), emin, delta, nbins, ns, nel);
    if (S.have[2]) res->efermi_linear = od_calc_efermi_from_intdos(S.intdos[2].data(), emin, delta, nbins, ns, nel);
  }
  if (p->dos_per_volume) {  // dos_utils.f90:269-291
    if (!(p->cell_volume > 0.0)) return od_fail(ctx, OD_ERR_ARG, "dos_per_volume needs cell_volume");
    for (int t = 0; t < 3; ++t)
      if (S.have[t]) {
        for (double &v : S.dos[t]) v = v / p->cell_volume;
        for (double &v : S.intdos[t]) v = v / p->cell_volume;
      }
    if (S.have_weighted)
      for (double &v : S.weighted) v = v / p->cell_volume;
  }
  return OD_OK;
}

static int tidx(char t) { return t == 'f' ? 0 : t == 'a' ? 1 : t == 'l' ? 2 : -1; }

extern "C" int od_dos_get(od_ctx *ctx, char dos_type, double *dos, double *intdos) {
  if (!ctx) return OD_ERR_ARG;
  const int t = tidx(dos_type);
  if (t < 0 || !ctx->dos_res.have[t]) return od_fail(ctx, OD_ERR_STATE, "dos_%c has not been calculated", dos_type);
  const size_t n = (size_t)ctx->dos_res.nbins * ctx->dos_res.ns;
  if (dos) memcpy(dos, ctx->dos_res.dos[t].data(), sizeof(double) * n);
  if (intdos) memcpy(intdos, ctx->dos_res.intdos[t].data(), sizeof(double) * n);
  return OD_OK;
}

extern "C" int od_dos_get_weighted(od_ctx *ctx, double *weighted_dos) {
  if (!ctx || !weighted_dos) return OD_ERR_ARG;
  if (!ctx->dos_res.have_weighted) return od_fail(ctx, OD_ERR_STATE, "weighted_dos has not been calculated");
  memcpy(weighted_dos, ctx->dos_res.weighted.data(), sizeof(double) * ctx->dos_res.weighted.size());
  return OD_OK;
}

// ------------------------------------------------------------------ efermi : insulator
namespace {
__global__ void vbm_cbm_kernel(const double *__restrict__ bands, int nb, int ns, int nk, int top0, int top1, int has_c0, int has_c1,
                               double *__restrict__ out /* [2*gridDim.x]: vbm, cbm */) {
  double vbm = -DBL_MAX, cbm = DBL_MAX;
  const size_t total = (size_t)nk * ns;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int is = (int)(t % ns);
    const size_t ik = t / ns;
    const int top = is == 0 ? top0 : top1;  // 1-based top occupied band
    const double *col = bands + (size_t)nb * (is + (size_t)ns * ik);
    vbm = fmax(vbm, col[top - 1]);
    if (is == 0 ? has_c0 : has_c1) cbm = fmin(cbm, col[top]);
  }
  __shared__ double sv[256], sc[256];
  sv[threadIdx.x] = vbm;
  sc[threadIdx.x] = cbm;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sv[threadIdx.x] = fmax(sv[threadIdx.x], sv[threadIdx.x + s]);
      sc[threadIdx.x] = fmin(sc[threadIdx.x], sc[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = sv[0]; out[2 * blockIdx.x + 1] = sc[0]; }
}
}  // namespace

static int efermi_insulator(od_ctx *ctx, const od_dos_params *p, double *efermi) {
  const int ns = ctx->ns, nb = ctx->nb;
  int top[2] = {1, 1}, has_c[2] = {0, 0};
  for (int is = 0; is < ns; ++is) {
    top[is] = (int)std::ceil(p->num_electrons[is] / ctx->eps);  // dos_utils.f90:339
    if (top[is] < 1 || top[is] > nb) return od_fail(ctx, OD_ERR_ARG, "top occupied band %d outside 1..%d", top[is], nb);
    has_c[is] = (p->num_electrons[is] + 1 <= nb) && (top[is] + 1 <= nb);  // :344 (+ bounds)
  }
  const int blocks = std::min(ctx->sm_count * 4, (int)(((size_t)ctx->nk * ns + 255) / 256));
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(double) * 2 * (size_t)blocks));
  vbm_cbm_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bands.as<double>(), nb, ns, ctx->nk, top[0], top[1], has_c[0], has_c[1],
                                                   ctx->small.as<double>());
  OD_LAUNCH_CHECK(ctx);
  std::vector<double> h(2 * (size_t)blocks);
  OD_CUDA(ctx, cudaMemcpyAsync(h.data(), ctx->small.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double vbm = -DBL_MAX, cbm = DBL_MAX;
  for (int b = 0; b < blocks; ++b) { vbm = std::max(vbm, h[2 * b]); cbm = std::min(cbm, h[2 * b + 1]); }
  if (ctx->nranks > 1) {  // dos_utils.f90:353-356
    double v[2] = {cbm, -vbm};
    OD_CHECK(od_allreduce(ctx, v, 2, 'm'));
    cbm = v[0];
    vbm = -v[1];
  }
  *efermi = (cbm == DBL_MAX) ? vbm + 0.5 : vbm + 0.5 * (cbm - vbm);  // :361-365
  return OD_OK;
}

extern "C" int od_dos_utils_set_efermi(od_ctx *ctx, od_dos_params *p, double *efermi) {
  if (!ctx || !p || !efermi) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  switch (p->efermi_choice) {
    case OD_EFERMI_FILE: *efermi = p->efermi_castep; return OD_OK;
    case OD_EFERMI_USER: *efermi = p->efermi_user; return OD_OK;
    case OD_EFERMI_INSULATOR: return efermi_insulator(ctx, p, efermi);
    case OD_EFERMI_OPTADOS: {
      od_dos_result r;
      OD_CHECK(od_dos_utils_calculate(ctx, p, nullptr, 0, 0, 0, OD_MEM_HOST, &r));
      if (p->fixed) *efermi = r.efermi_fixed;      // dos_utils.f90:373-375
      if (p->linear) *efermi = r.efermi_linear;
      if (p->adaptive) *efermi = r.efermi_adaptive;
      return OD_OK;
    }
    default: return od_fail(ctx, OD_ERR_REF, "Error in dos_utils_set_efermi: unknown efermi choice this is a bug (dos_utils.f90:379)");
  }
}

// ------------------------------------------------------------------ dos_utils_calculate_at_e
// dos_utils.f90:1531-1648: every enabled scheme fills its row of dos_at_e(1:3, ns) = (fixed, adaptive, linear);
// the weighted values come from the most complex enabled scheme only (same rule as dos_utils_calculate, Q14).
extern "C" int od_dos_utils_calculate_at_e(od_ctx *ctx, const od_dos_params *p, double energy, double delta_bins,
                                           const double *matrix_weights, int mw_norb, int mw_nb, int mw_nk, int mw_mem,
                                           double *dos_at_e, double *weighted_dos_at_e) {
  if (!ctx || !p || !dos_at_e) return OD_ERR_ARG;
  if (!(p->linear || p->adaptive || p->fixed)) return od_fail(ctx, OD_ERR_REF, " DOS: No Broadening Set (dos_utils.f90:1545)");
  if (matrix_weights && !weighted_dos_at_e) return od_fail(ctx, OD_ERR_ARG, "matrix_weights given without weighted_dos_at_e");
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  const int ns = ctx->ns;
  for (int i = 0; i < 3 * ns; ++i) dos_at_e[i] = 0.0;
  const char types[3] = {'f', 'a', 'l'};
  const int enabled[3] = {p->fixed, p->adaptive, p->linear};
  const bool with_w[3] = {matrix_weights && !p->adaptive && !p->linear, matrix_weights && !p->linear, matrix_weights != nullptr};
  std::vector<double> row((size_t)ns);
  for (int t = 0; t < 3; ++t) {
    if (!enabled[t]) continue;
    OD_CHECK(od_calculate_dos_at_e(ctx, types[t], energy, delta_bins, &p->br, with_w[t] ? matrix_weights : nullptr, mw_norb, mw_nb, mw_nk,
                                   mw_mem, /*merge=*/1, row.data(), with_w[t] ? weighted_dos_at_e : nullptr));
    for (int is = 0; is < ns; ++is) dos_at_e[t + 3 * is] = row[is];  // dos_at_e(t, is)
  }
  return OD_OK;
}

// ------------------------------------------------------------------ jdos_utils_calculate
extern "C" int od_jdos_utils_calculate(od_ctx *ctx, od_jdos_params *p, const double *matrix_weights, int n_geom, int mw_mem,
                                       const double *optical_mat, int ome_mem, int geom, const double qdir[3], const double *symops,
                                       int num_symm, od_jdos_result *res) {
  if (!ctx || !p || !res) return OD_ERR_ARG;
  if (!(p->linear || p->adaptive || p->fixed)) return od_fail(ctx, OD_ERR_REF, "JDOS: No Broadening Set");
  if (matrix_weights && optical_mat) return od_fail(ctx, OD_ERR_ARG, "give matrix_weights or optical_mat, not both");
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  memset(res, 0, sizeof(*res));
  // setup_energy_scale (jdos_utils.f90:177-231)
  if (p->jdos_max_energy < 0.0) {
    double mn, mx;
    OD_CHECK(od_band_minmax(ctx, &mn, &mx));
    p->jdos_max_energy = p->efermi - mx;  // :197 (sic)
  }
  if (!(p->jdos_spacing > 0.0)) return od_fail(ctx, OD_ERR_ARG, "jdos_spacing must be positive");
  const int nbins = std::abs((int)std::ceil(p->jdos_max_energy / p->jdos_spacing));
  if (nbins < 2) return od_fail(ctx, OD_ERR_ARG, "jdos_nbins = %d", nbins);
  p->jdos_max_energy = nbins * p->jdos_spacing;
  const double delta = p->jdos_max_energy / (double)(nbins - 1);

  const int ns = ctx->ns;
  const size_t n1 = (size_t)nbins * ns;
  const double *mw_dev = nullptr, *ome_dev = nullptr;
  OpticsCoef coef;
  memset(&coef, 0, sizeof(coef));
  int ng = 0;
  if (matrix_weights) {
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, matrix_weights, sizeof(double) * (size_t)ctx->nb * ctx->nb * ctx->nk * ns * n_geom, mw_mem, ctx->mw_tmp, &d));
    mw_dev = static_cast<const double *>(d);
    ng = n_geom;
  } else if (optical_mat) {
    OD_CHECK(od_optics_coefficients(ctx, geom, qdir, symops, num_symm, &coef));
    const void *d = nullptr;
    OD_CHECK(od_stage_in(ctx, optical_mat, sizeof(double) * 2 * (size_t)ctx->nb * ctx->nb * 3 * ctx->nk * ns, ome_mem, ctx->in_tmp, &d));
    ome_dev = static_cast<const double *>(d);
    ng = coef.n_geom;
  }
  Spectra &S = ctx->jdos_res;
  S = Spectra();
  S.nbins = nbins;
  S.ns = ns;
  const char types[3] = {'f', 'a', 'l'};
  const int enabled[3] = {p->fixed, p->adaptive, p->linear};
  for (int t = 0; t < 3; ++t) {
    if (!enabled[t]) continue;
    double *accum = nullptr;
    OD_CHECK(od_calculate_jdos_device(ctx, types[t], delta, nbins, &p->br, p->efermi, p->scissor_op, p->exclude_bands,
                                      p->num_exclude_bands, mw_dev, ome_dev, ome_dev ? &coef : nullptr, ng, /*merge=*/1, &accum));
    OD_CHECK(od_merge_status_check(ctx));
    OD_CHECK(pull(ctx, S.dos[t], accum, n1));
    S.have[t] = true;
    if (ng > 0) {  // every enabled scheme recomputes weighted_jdos; the last one survives (jdos_utils.f90:109-136)
      OD_CHECK(pull(ctx, S.weighted, accum + n1, n1 * ng));
      S.have_weighted = true;
      S.norb = ng;
    }
  }
  if (p->dos_per_volume) {  // jdos_utils.f90:157-172 (jdos only)
    if (!(p->cell_volume > 0.0)) return od_fail(ctx, OD_ERR_ARG, "dos_per_volume needs cell_volume");
    for (int t = 0; t < 3; ++t)
      if (S.have[t])
        for (double &v : S.dos[t]) v = v / p->cell_volume;
  }
  res->nbins = nbins;
  res->delta_bins = delta;
  res->have_fixed = S.have[0];
  res->have_adaptive = S.have[1];
  res->have_linear = S.have[2];
  res->have_weighted = S.have_weighted;
  res->n_geom = ng;
  return OD_OK;
}

extern "C" int od_jdos_get(od_ctx *ctx, char jdos_type, double *jdos) {
  if (!ctx || !jdos) return OD_ERR_ARG;
  const int t = tidx(jdos_type);
  if (t < 0 || !ctx->jdos_res.have[t]) return od_fail(ctx, OD_ERR_STATE, "jdos_%c has not been calculated", jdos_type);
  memcpy(jdos, ctx->jdos_res.dos[t].data(), sizeof(double) * (size_t)ctx->jdos_res.nbins * ctx->jdos_res.ns);
  return OD_OK;
}

extern "C" int od_jdos_get_weighted(od_ctx *ctx, double *weighted_jdos) {
  if (!ctx || !weighted_jdos) return OD_ERR_ARG;
  if (!ctx->jdos_res.have_weighted) return od_fail(ctx, OD_ERR_STATE, "weighted_jdos has not been calculated");
  memcpy(weighted_jdos, ctx->jdos_res.weighted.data(), sizeof(double) * ctx->jdos_res.weighted.size());
  return OD_OK;
}

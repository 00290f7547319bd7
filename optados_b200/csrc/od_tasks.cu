// od_tasks.cu -- host-side mirrors of two of the reference's tasks, built from the entry points of this library:
//   optics_calculate  optics.f90:78-155   (interband): make_weights fused into the JDOS pass -> epsilon_2 ->
//                                         Kramers-Kronig epsilon_1 -> conductivity, refractive index, loss function,
//                                         absorption, reflection (the last five not for optics_geom = tensor)
//   core_calculate    core.f90:38-83      Fermi level -> ELNES weights -> weighted DOS -> LAI broadening -> the unit
//                                         factors of write_core (core.f90:317-323)
// Same order of operations, same flags, same quirks (sticky dos_nbins Q4, weighted spectrum of the most complex
// scheme Q14, single-precision literals Q11), including the intraband (Drude) terms (optics_intraband).
#include <cmath>
#include <cstring>
#include <vector>

#include "od_internal.h"
#include "od_spectral.h"

namespace {
// optics.f90:68-72
constexpr double kEps0 = 8.8541878176E-12, kECharge = 1.602176487E-19, kHbar = 1.054571628E-34, kCSpeed = 299792458.0;
constexpr double kEMass = 9.10938215E-31;
constexpr double kBohr2Ang = 0.52917720859;
constexpr double kPi = 3.141592653589793238462643383279502884197;
}  // namespace

// ------------------------------------------------------------------ optics
extern "C" void od_optics_params_defaults(od_optics_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  od_jdos_params_defaults(&p->jdos);
  p->geom = OD_GEOM_POLYCRYS;       // parameters.f90: optics_geom default 'polycrystalline'
  p->qdir[0] = 1.0; p->qdir[1] = 1.0; p->qdir[2] = 1.0;  // parameters.f90: optics_qdir default
  p->lossfn_gaussian = 0.0;
  p->intraband = 0;               // parameters.f90:361
  p->drude_broadening = 1.0e14;   // parameters.f90:364
  p->dos_delta_bins = 0.0;
}

// optics_intraband = T: the weighted DOS at the Fermi level of the diagonal matrix elements (optics.f90:110-121) and the
// Drude terms it feeds in calc_epsilon_2 (:517-549), calc_epsilon_1 (:586-617), calc_loss_fn (:632-719), calc_conduct
// (:736-743) and calc_refract (:769-778).  wj = weighted_jdos(nbins, ns, ng), e1 = the interband epsilon_1 (Kramers-Kronig
// of the interband epsilon_2, which the Drude terms do not enter).  Root-only O(B) arithmetic in the reference: host code.
static int optics_intraband(od_ctx *ctx, od_optics_params *p, const double *optical_mat, int ome_mem, const double *symops, int num_symm,
                            const std::vector<double> &wj, const std::vector<double> &e1, od_optics_result *res) {
  OpticsRes &O = ctx->optics_res;
  const int nbins = O.nbins, ng = O.n_geom, ns = ctx->ns, nb = ctx->nb, nk = ctx->nk;
  const double vol = p->jdos.cell_volume, gam = p->drude_broadening, dE = O.E[1] - O.E[0];
  // dos_matrix_weights(N, ib, ik, is) = matrix_weights(ib, ib, ik, is, N), formed from the matrix elements on the device
  OpticsCoef coef;
  OD_CHECK(od_optics_coefficients(ctx, p->geom, p->qdir, symops, num_symm, &coef));
  const double *ome_dev = ome_mem == OD_MEM_DEVICE ? optical_mat : ctx->in_tmp.as<double>();  // staged by od_jdos_utils_calculate
  OD_CHECK(od_buf_ensure(ctx, ctx->task_mw, sizeof(double) * (size_t)ng * nb * nk * ns));
  OD_CHECK(od_optics_diag_weights(ctx, ome_dev, coef, ctx->task_mw.as<double>()));
  od_dos_params dp;
  od_dos_params_defaults(&dp);
  dp.fixed = p->jdos.fixed; dp.adaptive = p->jdos.adaptive; dp.linear = p->jdos.linear; dp.br = p->jdos.br;
  std::vector<double> dos_at_e(3 * (size_t)ns), wdae((size_t)ns * ng, 0.0);
  OD_CHECK(od_dos_utils_calculate_at_e(ctx, &dp, p->jdos.efermi, p->dos_delta_bins, ctx->task_mw.as<double>(), ng, nb, nk, OD_MEM_DEVICE,
                                       dos_at_e.data(), wdae.data()));
  const double c20 = (double)1E-20f, c30 = (double)1E-30f, c10 = (double)1E-10f;  // single-precision literals (Q11)
  const double eps2c = (kECharge * kPi * c20) / (vol * c30 * kEps0);
  std::vector<double> intra((size_t)ng, 0.0);
  for (int g = 0; g < ng; ++g) {
    for (int is = 0; is < ns; ++is) intra[g] = intra[g] + wdae[is + (size_t)ns * g];
    intra[g] = intra[g] * kECharge / (vol * c10 * kEps0);
  }
  res->n_parts = 3;
  O.epsilon.assign((size_t)nbins * 2 * ng * 3, 0.0);  // epsilon(nbins, 2, n_geom, 3)
  auto EP = [&](int i, int c, int g, int t) -> double & { return O.epsilon[(size_t)i + (size_t)nbins * (c + 2 * (g + (size_t)ng * t))]; };
  for (int g = 0; g < ng; ++g) {
    for (int is = 0; is < ns; ++is)  // the spin loop encloses the Drude term too (:536-547)
      for (int i = 1; i < nbins; ++i) {
        const double E = O.E[i];
        EP(i, 1, g, 0) = EP(i, 1, g, 0) + eps2c * wj[i + (size_t)nbins * (is + (size_t)ns * g)];
        EP(i, 1, g, 1) = EP(i, 1, g, 1) + ((intra[g] * (kECharge * kECharge) * kHbar * gam) / (((E * kECharge) * (E * kECharge)) + ((gam * kHbar) * (gam * kHbar))));
        EP(i, 1, g, 2) = EP(i, 1, g, 2) + EP(i, 1, g, 1) + EP(i, 1, g, 0) * E * kECharge;
      }
    for (int i = 0; i < nbins; ++i) {
      const double E = O.E[i];
      EP(i, 0, g, 0) = e1[i + (size_t)nbins * g];
      EP(i, 0, g, 1) = 1.0 - (intra[g] / ((E * E) + (((gam * kHbar) / kECharge) * ((gam * kHbar) / kECharge))));
      EP(i, 0, g, 2) = EP(i, 0, g, 0) + EP(i, 0, g, 1) - 1.0;
    }
  }
  if (ng == 1) {  // :552-563
    double x = 0.0;
    for (int N = 2; N <= nbins; ++N) x = x + ((N * (dE * dE) * EP(N - 1, 1, 0, 2)) / ((kHbar * kHbar) * O.E[N - 1] * kECharge));
    res->n_eff = (x * kEMass * vol * c30 * kEps0 * 2) / (kPi);
  }
  if (p->geom == OD_GEOM_TENSOR) return OD_OK;
  const bool lossb = fabs(p->lossfn_gaussian) >= 1.0e-6;
  const int ncol = lossb ? 4 : 3;
  O.loss_fn.assign((size_t)nbins * ncol, 0.0);
  for (int i = 0; i < nbins; ++i) {
    double a = EP(i, 0, 0, 0), b = EP(i, 1, 0, 0);
    O.loss_fn[i] = b / (a * a + b * b);
    if (i > 0) {
      a = EP(i, 0, 0, 1); b = EP(i, 1, 0, 1) / (O.E[i] * kECharge);
      O.loss_fn[i + (size_t)nbins] = b / (a * a + b * b);
      a = EP(i, 0, 0, 2); b = EP(i, 1, 0, 2) / (O.E[i] * kECharge);
      O.loss_fn[i + 2 * (size_t)nbins] = b / (a * a + b * b);
    }
  }
  if (lossb) {  // the Gaussian broadening of the TOTAL loss function (:683-697) is the O(B^2) pass of od_calc_loss_fn (GPU), fed
                // with a spectrum whose -Im(1 / eps) is column 3: eps = (0, 1 / L) gives b / (a a + b b) = L; eps = (1, 0) gives 0
    std::vector<double> a((size_t)nbins), b((size_t)nbins), tmp((size_t)nbins);
    for (int i = 0; i < nbins; ++i) {
      const double L = O.loss_fn[i + 2 * (size_t)nbins];
      a[i] = L != 0.0 ? 0.0 : 1.0;
      b[i] = L != 0.0 ? 1.0 / L : 0.0;
    }
    OD_CHECK(od_calc_loss_fn(ctx, O.E.data(), nbins, a.data(), b.data(), p->lossfn_gaussian, tmp.data(), O.loss_fn.data() + 3 * (size_t)nbins));
  }
  {
    double x = 0.0;
    for (int N = 2; N <= nbins; ++N) x = x + (N * (dE * dE) * O.loss_fn[N - 1 + 2 * (size_t)nbins]);
    res->n_eff2 = x * (kEMass * vol * c30 * kEps0 * 2) / (kPi * (kHbar * kHbar));
    x = 0;
    for (int N = 2; N <= nbins; ++N) x = x + (O.loss_fn[N - 1 + 2 * (size_t)nbins] / N);
    res->n_eff3 = x;
  }
  O.conduct.resize((size_t)nbins * 2); O.refract.resize((size_t)nbins * 2); O.absorp.resize(nbins); O.reflect.resize(nbins);
  for (int i = 0; i < nbins; ++i) {
    const double E = O.E[i], e1i = EP(i, 0, 0, 0), e1t = EP(i, 0, 0, 2), e2t = EP(i, 1, 0, 2);
    O.conduct[i] = kEps0 * e2t / kHbar;
    O.conduct[i + (size_t)nbins] = (E * kECharge / kHbar) * kEps0 * (1.0 - e1t);
    const double r = e2t / (E * kECharge);  // 0 / 0 at E = 0, as in the reference
    O.refract[i] = pow(0.5 * (pow((e1t * e1t) + (r * r), 0.5) + e1i), 0.5);
    O.refract[i + (size_t)nbins] = pow(0.5 * (pow((e1i * e1i) + (r * r), 0.5) - e1i), 0.5);
    O.absorp[i] = 2 * O.refract[i + (size_t)nbins] * E * kECharge / (kHbar * kCSpeed);
    const double n = O.refract[i], k = O.refract[i + (size_t)nbins];
    O.reflect[i] = (((n - 1) * (n - 1)) + (k * k)) / (((n + 1) * (n + 1)) + (k * k));
  }
  res->have_derived = 1;
  res->have_loss_fn_broadened = lossb ? 1 : 0;
  res->n_loss_fn = ncol;
  return OD_OK;
}

extern "C" int od_optics_calculate(od_ctx *ctx, od_optics_params *p, const double *optical_mat, int ome_mem, const double *symops,
                                   int num_symm, od_optics_result *res) {
  if (!ctx || !p || !res || !optical_mat) return ctx ? od_fail(ctx, OD_ERR_ARG, "od_optics_calculate: null argument") : OD_ERR_ARG;
  if (!(p->jdos.cell_volume > 0.0)) return od_fail(ctx, OD_ERR_ARG, "od_optics_calculate needs cell_volume");
  memset(res, 0, sizeof(*res));
  od_jdos_result jr;
  // make_weights + jdos_utils_calculate(matrix_weights, weighted_jdos)   (optics.f90:104-107)
  OD_CHECK(od_jdos_utils_calculate(ctx, &p->jdos, nullptr, 0, OD_MEM_HOST, optical_mat, ome_mem, p->geom, p->qdir, symops, num_symm, &jr));
  const int nbins = jr.nbins, ng = jr.n_geom, ns = ctx->ns;
  OpticsRes &O = ctx->optics_res;
  O = OpticsRes();
  O.nbins = nbins; O.n_geom = ng; O.delta = jr.delta_bins;
  O.E.resize(nbins);
  for (int i = 0; i < nbins; ++i) O.E[i] = (double)i * jr.delta_bins;  // jdos_utils.f90:217
  std::vector<double> wj((size_t)nbins * ns * ng), e2((size_t)nbins * ng), e1((size_t)nbins * ng);
  OD_CHECK(od_jdos_get_weighted(ctx, wj.data()));
  OD_CHECK(od_calc_epsilon_2(ctx, wj.data(), nbins, ns, ng, p->jdos.cell_volume, e2.data()));   // optics.f90:495-551
  OD_CHECK(od_calc_epsilon_1(ctx, O.E.data(), nbins, ng, e2.data(), e1.data()));                // :569-621
  O.epsilon.assign((size_t)nbins * 2 * ng, 0.0);  // epsilon(nbins, 2, n_geom): (:,1,:) real part, (:,2,:) imaginary part
  for (int g = 0; g < ng; ++g)
    for (int i = 0; i < nbins; ++i) {
      O.epsilon[i + (size_t)nbins * (0 + 2 * g)] = e1[i + (size_t)nbins * g];
      O.epsilon[i + (size_t)nbins * (1 + 2 * g)] = e2[i + (size_t)nbins * g];
    }
  res->nbins = nbins; res->n_geom = ng; res->delta_bins = jr.delta_bins;
  res->n_parts = 1;
  const double dE = O.E[1] - O.E[0], vol = p->jdos.cell_volume;
  if (p->intraband) return optics_intraband(ctx, p, optical_mat, ome_mem, symops, num_symm, wj, e1, res);
  if (ng == 1) {  // sum rule of calc_epsilon_2 (optics.f90:552-563): the bin index, not the energy, multiplies dE**2
    double x = 0.0;
    for (int N = 2; N <= nbins; ++N) x = x + ((N * (dE * dE) * e2[N - 1]) / (kHbar * kHbar));
    res->n_eff = (x * kEMass * vol * (double)1E-30f * kEps0 * 2) / (kPi);
  }
  if (p->geom == OD_GEOM_TENSOR) return OD_OK;  // optics.f90:130
  // ---- derived quantities of component 1 (optics.f90:722-825)
  O.conduct.resize((size_t)nbins * 2); O.refract.resize((size_t)nbins * 2); O.absorp.resize(nbins); O.reflect.resize(nbins);
  for (int i = 0; i < nbins; ++i) {
    const double E = O.E[i], a = e1[i], b = e2[i];
    O.conduct[i] = (E * kECharge / kHbar) * kEps0 * b;                      // calc_conduct
    O.conduct[i + (size_t)nbins] = (E * kECharge / kHbar) * kEps0 * (1.0 - a);
    const double mod = pow((a * a) + (b * b), 0.5);
    O.refract[i] = pow(0.5 * (mod + a), 0.5);                               // calc_refract
    O.refract[i + (size_t)nbins] = pow(0.5 * (mod - a), 0.5);
    O.absorp[i] = 2 * O.refract[i + (size_t)nbins] * E * kECharge / (kHbar * kCSpeed);  // calc_absorp
    const double n = O.refract[i], k = O.refract[i + (size_t)nbins];
    O.reflect[i] = (((n - 1) * (n - 1)) + (k * k)) / (((n + 1) * (n + 1)) + (k * k));     // calc_reflect
  }
  const bool lossb = fabs(p->lossfn_gaussian) >= 1.0e-6;  // parameters.f90:371
  O.loss_fn.assign((size_t)nbins * (lossb ? 2 : 1), 0.0);
  OD_CHECK(od_calc_loss_fn(ctx, O.E.data(), nbins, e1.data(), e2.data(), p->lossfn_gaussian, O.loss_fn.data(),
                           lossb ? O.loss_fn.data() + nbins : nullptr));
  {  // sum rules of calc_loss_fn (optics.f90:700-719)
    double x = 0.0;
    for (int N = 2; N <= nbins; ++N) x = x + (N * (dE * dE) * O.loss_fn[N - 1]);
    res->n_eff2 = x * (kEMass * vol * (double)1E-30f * kEps0 * 2) / (kPi * (kHbar * kHbar));
    x = 0;
    for (int N = 2; N <= nbins; ++N) x = x + (O.loss_fn[N - 1] / N);
    res->n_eff3 = x;
  }
  res->have_derived = 1;
  res->have_loss_fn_broadened = lossb ? 1 : 0;
  res->n_loss_fn = lossb ? 2 : 1;
  return OD_OK;
}

static void copy_out(const std::vector<double> &v, double *dst) {
  if (dst && !v.empty()) memcpy(dst, v.data(), sizeof(double) * v.size());
}

extern "C" int od_optics_get(od_ctx *ctx, double *E, double *epsilon, double *conduct, double *refract, double *loss_fn, double *absorp,
                             double *reflect) {
  if (!ctx) return OD_ERR_ARG;
  const OpticsRes &O = ctx->optics_res;
  if (O.nbins == 0) return od_fail(ctx, OD_ERR_STATE, "od_optics_calculate has not been called");
  if ((conduct || refract || loss_fn || absorp || reflect) && O.conduct.empty())
    return od_fail(ctx, OD_ERR_STATE, "no derived optical properties for optics_geom = tensor (optics.f90:130)");
  copy_out(O.E, E); copy_out(O.epsilon, epsilon); copy_out(O.conduct, conduct); copy_out(O.refract, refract);
  copy_out(O.loss_fn, loss_fn); copy_out(O.absorp, absorp); copy_out(O.reflect, reflect);
  return OD_OK;
}

// ------------------------------------------------------------------ core
extern "C" void od_core_params_defaults(od_core_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  od_dos_params_defaults(&p->dos);
  p->core_type = OD_CORE_ABSORPTION;  // parameters.f90: core_type default 'absorption'
  p->polar = 0;                       // core_geom default 'polycrystalline'
  p->qdir[0] = 0.0; p->qdir[1] = 0.0; p->qdir[2] = 1.0;
  p->lai_broadening = 0;
  p->set_efermi_zero = 0;
  p->core_chemical_shift = -1.0;  // parameters.f90:267
}

extern "C" int od_core_calculate(od_ctx *ctx, od_core_params *p, const double *elnes_mat, int elnes_mem, int norb, const double *symops,
                                 int num_sym, od_core_result *res) {
  if (!ctx || !p || !res || !elnes_mat) return ctx ? od_fail(ctx, OD_ERR_ARG, "od_core_calculate: null argument") : OD_ERR_ARG;
  if (norb < 1) return od_fail(ctx, OD_ERR_ARG, "od_core_calculate: norb = %d", norb);
  if (!(p->dos.cell_volume > 0.0)) return od_fail(ctx, OD_ERR_ARG, "od_core_calculate needs cell_volume");
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  memset(res, 0, sizeof(*res));
  const int ns = ctx->ns, nb = ctx->nb, nk = ctx->nk;
  // if (.not. efermi_set) call dos_utils_set_efermi   (core.f90:62)
  double efermi = 0.0;
  OD_CHECK(od_dos_utils_set_efermi(ctx, &p->dos, &efermi));
  // core_prepare_matrix_elements (core.f90:64), weights stay on the device
  const size_t nmw = (size_t)norb * nb * nk * ns;
  OD_CHECK(od_buf_ensure(ctx, ctx->task_mw, sizeof(double) * nmw));
  OD_CHECK(od_core_prepare_matrix_elements(ctx, elnes_mat, elnes_mem, norb, p->polar, p->qdir, symops, num_sym, p->core_type, efermi,
                                           ctx->task_mw.as<double>(), OD_MEM_DEVICE));
  // dos_utils_calculate(matrix_weights, weighted_dos) (core.f90:66); dos_nbins is sticky by now (quirk Q4)
  od_dos_result dr;
  OD_CHECK(od_dos_utils_calculate(ctx, &p->dos, ctx->task_mw.as<double>(), norb, nb, nk, OD_MEM_DEVICE, &dr));
  if (!dr.have_weighted) return od_fail(ctx, OD_ERR_STATE, "no weighted DOS came back");
  const int nbins = dr.nbins;
  CoreRes &R = ctx->core_res;
  R = CoreRes();
  R.nbins = nbins; R.ns = ns; R.norb = norb; R.efermi = efermi;
  R.E.resize(nbins);
  for (int i = 0; i < nbins; ++i) R.E[i] = dr.emin + (double)i * dr.delta_bins;
  R.E_shift = R.E;  // write_core's abscissa (core.f90:490-494, :533-535); the LAI broadening below works on E itself
  if (p->set_efermi_zero)
    for (double &e : R.E_shift) e = e - efermi;
  if (p->core_chemical_shift != -1.0)
    for (int i = 0; i < nbins; ++i) R.E_shift[i] = R.E[i] + p->core_chemical_shift - ctx->cbm_energy;
  const size_t n = (size_t)nbins * ns * norb;
  R.wdos.resize(n);
  OD_CHECK(od_dos_get_weighted(ctx, R.wdos.data()));
  if (p->lai_broadening) {  // core.f90:70-76
    R.wdos_broadened.resize(n);
    OD_CHECK(od_core_lai_broadening(ctx, R.E.data(), nbins, ns * norb, efermi, p->lai_gaussian_width, p->lai_lorentzian_width,
                                    p->lai_lorentzian_scale, p->lai_lorentzian_offset, R.wdos.data(), R.wdos_broadened.data()));
  }
  // write_core (core.f90:317-323): Ang^2 and the dimensionless-epsilon_2 factor, single-precision literals (Q11)
  const double b2 = kBohr2Ang * kBohr2Ang;  // bohr2ang**2
  const double c = (kECharge * kPi * (double)1E-20f) / (p->dos.cell_volume * (double)1E-30f * kEps0);
  for (double &v : R.wdos) { v = v * b2; v = v * c; }
  for (double &v : R.wdos_broadened) { v = v * b2; v = v * c; }
  res->nbins = nbins; res->emin = dr.emin; res->delta_bins = dr.delta_bins; res->efermi = efermi; res->norb = norb;
  res->have_broadened = p->lai_broadening ? 1 : 0;
  return OD_OK;
}

extern "C" int od_core_get(od_ctx *ctx, double *E, double *weighted_dos, double *weighted_dos_broadened) {
  if (!ctx) return OD_ERR_ARG;
  const CoreRes &R = ctx->core_res;
  if (R.nbins == 0) return od_fail(ctx, OD_ERR_STATE, "od_core_calculate has not been called");
  if (weighted_dos_broadened && R.wdos_broadened.empty()) return od_fail(ctx, OD_ERR_STATE, "core_LAI_broadening was off");
  copy_out(R.E_shift, E); copy_out(R.wdos, weighted_dos); copy_out(R.wdos_broadened, weighted_dos_broadened);
  return OD_OK;
}

// od_post.cu -- the reference's root-only O(B^2) post-processing passes on the GPU.
//
// Once the spectral integration takes milliseconds, these serial double loops over the energy grid (4e8 exp or
// divide evaluations per column at 20 000 bins) are what a run waits for:
//   calc_epsilon_1   optics.f90:569-621   Kramers-Kronig transform of epsilon_2 (interband)
//   calc_loss_fn     optics.f90:622-720   -Im(1/eps) and its Gaussian broadening (optics_lossfn_broadening)
//   core_lorentzian  core.f90:692-732     life-time broadening with an energy-dependent width
//   core_gaussian    core.f90:655-690     instrumental broadening (FWHM)
// One pattern: out(j, col) = sum_i term(E_j; E_i, v_i, p_i).  A thread owns an output bin, the sources of its
// segment stream through shared memory, segment partials are summed in a fixed order (deterministic).  The
// per-pair arithmetic is the reference's (same divisions, same operation order inside a term); only the
// association of the sum over i differs (segments), i.e. rounding at the 1e-16 B level.
#include <algorithm>
#include <cmath>
#include <vector>

#include "od_internal.h"

namespace {

constexpr int PP_THREADS = 128, PP_TILE = 256;
constexpr double PP_PI = 3.141592653589793238462643383279502884197;

struct KKOp {  // q = q + (((energy2*eps2(i))/((energy2**2) - (energy1**2)))*dE), i /= j
  double dE;
  __device__ __forceinline__ double operator()(int j, int i, double Ej, double Ei, double v, double) const {
    if (i == j) return 0.0;
    return ((Ei * v) / ((Ei * Ei) - (Ej * Ej))) * dE;
  }
};
struct GaussOp {  // g = sqrt(4 ln2 / pi) (1/W) exp(-4 ln2 ((E_j - E_i)/W)^2);  out += g v dE
  double pref, width, c4ln2, dE;
  __device__ __forceinline__ double operator()(int, int, double Ej, double Ei, double v, double) const {
    const double x = (Ej - Ei) / width;
    const double g = pref * exp(-c4ln2 * (x * x));
    return g * v * dE;
  }
};
struct LorentzOp {  // l = v L/(pi ((E_j - E_i)^2 + L^2));  out += l dE   (p = L, the half width of source bin i)
  double dE;
  __device__ __forceinline__ double operator()(int, int, double Ej, double Ei, double v, double L) const {
    const double d = Ej - Ei;
    const double l = v * L / (PP_PI * ((d * d) + (L * L)));
    return l * dE;
  }
};

// partial[(seg * ncols + col) * nbins + j] = sum over the sources of segment seg
template <class Op>
__global__ void __launch_bounds__(PP_THREADS) pair_sum_kernel(Op op, const double *__restrict__ E, const double *__restrict__ val,
                                                              const double *__restrict__ par, int nbins, int ncols, int seg_len,
                                                              double *__restrict__ partial) {
  __shared__ double sE[PP_TILE], sV[PP_TILE], sP[PP_TILE];
  const int j = blockIdx.x * PP_THREADS + threadIdx.x, col = blockIdx.y, seg = blockIdx.z;
  const int i0 = seg * seg_len, i1 = min(i0 + seg_len, nbins);
  const double Ej = j < nbins ? E[j] : 0.0;
  double q = 0.0;
  for (int t0 = i0; t0 < i1; t0 += PP_TILE) {
    __syncthreads();
    for (int k = threadIdx.x; k < PP_TILE; k += PP_THREADS) {
      const int i = t0 + k;
      const bool in = i < i1;
      sE[k] = in ? E[i] : 0.0;
      sV[k] = in ? val[(size_t)col * nbins + i] : 0.0;
      sP[k] = (in && par) ? par[i] : 1.0;
    }
    __syncthreads();
    const int n = min(PP_TILE, i1 - t0);
    if (j < nbins)
      for (int k = 0; k < n; ++k) q += op(j, t0 + k, Ej, sE[k], sV[k], sP[k]);
  }
  if (j < nbins) partial[((size_t)seg * ncols + col) * nbins + j] = q;
}

__global__ void pair_reduce_kernel(const double *__restrict__ partial, int nseg, size_t n, double *__restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double s = 0.0;
  for (int g = 0; g < nseg; ++g) s += partial[(size_t)g * n + t];
  out[t] = s;
}

// d_E, d_val (nbins x ncols), d_par (nbins or null) on the device -> d_out (nbins x ncols)
template <class Op>
int pair_sum(od_ctx *ctx, const Op &op, const double *d_E, const double *d_val, const double *d_par, int nbins, int ncols, double *d_out) {
  const int jb = (nbins + PP_THREADS - 1) / PP_THREADS;
  int nseg = (16 * ctx->sm_count + jb * ncols - 1) / (jb * ncols);  // enough CTAs of 128 threads to fill every SM
  nseg = std::max(1, std::min(nseg, 64));
  int seg_len = (nbins + nseg - 1) / nseg;
  seg_len = ((seg_len + PP_TILE - 1) / PP_TILE) * PP_TILE;
  nseg = (nbins + seg_len - 1) / seg_len;
  const size_t n = (size_t)nbins * ncols;
  OD_CHECK(od_buf_ensure(ctx, ctx->partial, sizeof(double) * n * nseg));
  dim3 grid(jb, ncols, nseg);
  pair_sum_kernel<Op><<<grid, PP_THREADS, 0, ctx->stream>>>(op, d_E, d_val, d_par, nbins, ncols, seg_len, ctx->partial.as<double>());
  OD_LAUNCH_CHECK(ctx);
  pair_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->partial.as<double>(), nseg, n, d_out);
  OD_LAUNCH_CHECK(ctx);
  return OD_OK;
}

// device staging: [E | a (n) | b (n) | par (nbins)] in scratch_a
struct Staged {
  double *E, *a, *b, *par;
};
int stage(od_ctx *ctx, const double *E, int nbins, size_t n, Staged *s) {
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_a, sizeof(double) * (2 * (size_t)nbins + 2 * n)));
  s->E = ctx->scratch_a.as<double>();
  s->a = s->E + nbins;
  s->b = s->a + n;
  s->par = s->b + n;
  OD_CUDA(ctx, cudaMemcpyAsync(s->E, E, sizeof(double) * nbins, cudaMemcpyHostToDevice, ctx->stream));
  return OD_OK;
}

int check_grid(od_ctx *ctx, const double *E, int nbins) {
  if (!ctx) return OD_ERR_ARG;
  if (!E || nbins < 2) return od_fail(ctx, OD_ERR_ARG, "post-processing needs an energy grid of at least 2 bins");
  return OD_OK;
}

}  // namespace

// ------------------------------------------------------------------ optics.f90:569-621
extern "C" int od_calc_epsilon_1(od_ctx *ctx, const double *E, int nbins, int n_geom, const double *epsilon_2, double *epsilon_1) {
  OD_CHECK(check_grid(ctx, E, nbins));
  if (!epsilon_2 || !epsilon_1 || n_geom < 1) return od_fail(ctx, OD_ERR_ARG, "od_calc_epsilon_1: null array or N_geom < 1");
  const size_t n = (size_t)nbins * n_geom;
  Staged s;
  OD_CHECK(stage(ctx, E, nbins, n, &s));
  OD_CUDA(ctx, cudaMemcpyAsync(s.a, epsilon_2, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  KKOp op{E[1] - E[0]};
  OD_CHECK(pair_sum(ctx, op, s.E, s.a, nullptr, nbins, n_geom, s.b));
  OD_CUDA(ctx, cudaMemcpyAsync(epsilon_1, s.b, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int g = 0; g < n_geom; ++g)  // diagonal components get the +1 (:603-607)
    for (int i = 0; i < nbins; ++i) {
      double &v = epsilon_1[i + (size_t)nbins * g];
      v = (g < 3) ? ((2.0 / PP_PI) * v) + 1.0 : ((2.0 / PP_PI) * v);
    }
  return OD_OK;
}

// ------------------------------------------------------------------ optics.f90:622-720 (interband)
extern "C" int od_calc_loss_fn(od_ctx *ctx, const double *E, int nbins, const double *epsilon_1, const double *epsilon_2,
                               double lossfn_gaussian, double *loss_fn, double *loss_fn_broadened) {
  OD_CHECK(check_grid(ctx, E, nbins));
  if (!epsilon_1 || !epsilon_2 || !loss_fn) return od_fail(ctx, OD_ERR_ARG, "od_calc_loss_fn: null array");
  if (lossfn_gaussian < 0.0) return od_fail(ctx, OD_ERR_ARG, "optics_lossfn_broadening must be positive (parameters.f90:370)");
  for (int i = 0; i < nbins; ++i) {  // -aimag(1/(eps1 + i eps2))
    const double a = epsilon_1[i], b = epsilon_2[i];
    loss_fn[i] = b / (a * a + b * b);
  }
  if (!loss_fn_broadened || fabs(lossfn_gaussian) < 1.0e-6) return OD_OK;  // parameters.f90:371
  Staged s;
  OD_CHECK(stage(ctx, E, nbins, (size_t)nbins, &s));
  OD_CUDA(ctx, cudaMemcpyAsync(s.a, loss_fn, sizeof(double) * nbins, cudaMemcpyHostToDevice, ctx->stream));
  GaussOp op{pow((4.0 * log(2.0)) / PP_PI, 0.5) * (1.0 / lossfn_gaussian), lossfn_gaussian, 4.0 * log(2.0), E[1] - E[0]};
  OD_CHECK(pair_sum(ctx, op, s.E, s.a, nullptr, nbins, 1, s.b));
  OD_CUDA(ctx, cudaMemcpyAsync(loss_fn_broadened, s.b, sizeof(double) * nbins, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

// ------------------------------------------------------------------ core.f90:70-76, 655-732
extern "C" int od_core_lai_broadening(od_ctx *ctx, const double *E, int nbins, int ncols, double efermi, double gaussian_width,
                                      double lorentzian_width, double lorentzian_scale, double lorentzian_offset,
                                      const double *weighted_dos, double *weighted_dos_broadened) {
  OD_CHECK(check_grid(ctx, E, nbins));
  if (!weighted_dos || !weighted_dos_broadened || ncols < 1) return od_fail(ctx, OD_ERR_ARG, "od_core_lai_broadening: null array or no columns");
  if (lorentzian_offset < 0.0) return od_fail(ctx, OD_ERR_ARG, "LAI_lorentzian_offset must be positive (parameters.f90:412)");
  const size_t n = (size_t)nbins * ncols;
  const bool lorentzian = lorentzian_width > 1E-14 || lorentzian_scale > 0.00001;  // parameters.f90:402, core.f90:74
  const bool gaussian = gaussian_width > 1E-14;                                     // parameters.f90:396
  if (!lorentzian && !gaussian) {  // core.f90:73: allocated and zeroed, nothing added
    std::fill(weighted_dos_broadened, weighted_dos_broadened + n, 0.0);
    return OD_OK;
  }
  const double dE = E[1] - E[0];
  Staged s;
  OD_CHECK(stage(ctx, E, nbins, n, &s));
  OD_CUDA(ctx, cudaMemcpyAsync(s.a, weighted_dos, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  double *src = s.a, *dst = s.b;
  if (lorentzian) {
    std::vector<double> L((size_t)nbins);
    for (int i = 0; i < nbins; ++i) {  // core.f90:711-719
      double w;
      if (E[i] >= (lorentzian_offset + efermi)) w = 0.5 * (lorentzian_width + ((E[i] - efermi - lorentzian_offset) * lorentzian_scale));
      else w = 0.5 * lorentzian_width;
      if ((w * PP_PI) < dE) w = dE / PP_PI;
      L[i] = w;
    }
    OD_CUDA(ctx, cudaMemcpyAsync(s.par, L.data(), sizeof(double) * nbins, cudaMemcpyHostToDevice, ctx->stream));
    OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // L is a host vector
    LorentzOp op{dE};
    OD_CHECK(pair_sum(ctx, op, s.E, src, s.par, nbins, ncols, dst));
    std::swap(src, dst);  // the Gaussian, if any, acts on the Lorentzian-broadened spectrum (core.f90:671-673)
  }
  if (gaussian) {
    GaussOp op{pow((4.0 * log(2.0)) / PP_PI, 0.5) * (1 / gaussian_width), gaussian_width, 4.0 * log(2.0), dE};
    OD_CHECK(pair_sum(ctx, op, s.E, src, nullptr, nbins, ncols, dst));
    std::swap(src, dst);
  }
  OD_CUDA(ctx, cudaMemcpyAsync(weighted_dos_broadened, src, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

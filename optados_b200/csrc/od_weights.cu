// od_weights.cu -- HBM-bound weight builders and small maps:
//   make_weights (optics.f90:157-435), core_prepare_matrix_elements (core.f90:84-168),
//   projection_merge (projection_utils.f90:64-99), the diag-of-OME band gradient
//   (electronic.f90:276-291), calc_epsilon_2 (optics.f90:495-551) and the synthetic band generator.
//
// B200-first restatement of the optical / core weights: for a REAL direction d,
//   |d.M|^2 = sum_ij d_i d_j Re(M_i conj(M_j)),   Re((a.M) conj(b.M)) = sum_ij a_i b_j Re(M_i conj(M_j)),
// so the whole symmetry x inversion average of the reference (up to 96 complex dot products per
// band pair) contracts on the host into one symmetric 3x3 real matrix per geometry component; the
// kernels then do 6 real products P_ij = Re(M_i conj(M_j)) and N_geom 6-term dot products per
// element -- the streams of complex matrix elements are read exactly once, coalesced.
#include <cmath>

#include "od_internal.h"
#include "od_spectral.h"

namespace {
inline void add_outer(double c[6], const double a[3], const double b[3], double coef) {
  c[0] += coef * a[0] * b[0];
  c[1] += coef * a[1] * b[1];
  c[2] += coef * a[2] * b[2];
  c[3] += coef * (a[0] * b[1] + a[1] * b[0]);
  c[4] += coef * (a[0] * b[2] + a[2] * b[0]);
  c[5] += coef * (a[1] * b[2] + a[2] * b[1]);
}
inline double SYM(const double *s, int a, int b, int n) { return s[a + 3 * (b + 3 * (size_t)n)]; }
// qdir(i) = sum_j sgn * ops(j,i,N2) * q(j)   (optics.f90:283-289, core.f90:133-139)
inline void rotate_q(const double *symops, int n2, double sgn, const double q[3], double out[3]) {
  for (int i = 0; i < 3; ++i) {
    out[i] = 0.0;
    for (int j = 0; j < 3; ++j) out[i] = out[i] + sgn * (SYM(symops, j, i, n2) * q[j]);
  }
}
}  // namespace

int od_optics_coefficients(od_ctx *ctx, int geom, const double optics_qdir[3], const double *symops, int num_symm, OpticsCoef *coef) {
  memset(coef, 0, sizeof(*coef));
  if (geom < 0 || geom > 3) return od_fail(ctx, OD_ERR_ARG, "unknown optics_geom %d", geom);
  if (num_symm < 0 || (num_symm > 0 && !symops)) return od_fail(ctx, OD_ERR_ARG, "bad symmetry operations");
  coef->n_geom = (geom == OD_GEOM_TENSOR) ? 6 : 1;
  const int N_in = 1;  // optics.f90:246
  const double nq = (double)(num_symm * (N_in + 1));
  double q[3] = {0, 0, 1}, q1[3] = {0, 0, 0}, q2[3] = {0, 0, 0}, qw = 1, qw1 = 1, qw2 = 1;
  if (geom == OD_GEOM_POLAR || geom == OD_GEOM_UNPOLAR) {
    if (!optics_qdir) return od_fail(ctx, OD_ERR_ARG, "optics_qdir missing");
    for (int i = 0; i < 3; ++i) q[i] = optics_qdir[i];
    qw = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (qw < 0.001) return od_fail(ctx, OD_ERR_REF, "Error:  please check optics_qdir, norm close to zero (optics.f90:225)");
  }
  if (geom == OD_GEOM_UNPOLAR) {  // optics.f90:229-244
    if (q[2] == 0) { q1[0] = 0.0; q1[1] = 0.0; q1[2] = 1.0; }
    else { q1[0] = 1.0; q1[1] = 1.0; q1[2] = -(q[0] + q[1]) / q[2]; }
    q2[0] = (q[1] * q1[2]) - (q[2] * q1[1]);
    q2[1] = (q[2] * q1[0]) - (q[0] * q1[2]);
    q2[2] = (q[0] * q1[1]) - (q[1] * q1[0]);
    qw1 = std::sqrt(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2]);
    qw2 = std::sqrt(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2]);
  }
  auto scaled = [](const double v[3], double w, double o[3]) { for (int i = 0; i < 3; ++i) o[i] = v[i] / w; };
  double a[3], b[3], d[3];
  if (geom == OD_GEOM_UNPOLAR) {
    if (num_symm == 0) {
      scaled(q1, qw1, a); add_outer(coef->c[0], a, a, 0.5);
      scaled(q2, qw2, b); add_outer(coef->c[0], b, b, 0.5);
    } else {
      for (int n2 = 0; n2 < num_symm; ++n2)
        for (int n3 = 1; n3 <= 1 + N_in; ++n3) {
          const double sgn = (n3 == 1) ? 1.0 : -1.0;
          rotate_q(symops, n2, sgn, q1, d); scaled(d, qw1, a); add_outer(coef->c[0], a, a, 0.5 / nq);
          rotate_q(symops, n2, sgn, q2, d); scaled(d, qw2, b); add_outer(coef->c[0], b, b, 0.5 / nq);
        }
    }
  } else if (geom == OD_GEOM_POLAR) {
    if (num_symm == 0) {
      scaled(q, qw, a); add_outer(coef->c[0], a, a, 1.0);
    } else {
      for (int n2 = 0; n2 < num_symm; ++n2)
        for (int n3 = 1; n3 <= 1 + N_in; ++n3) {
          rotate_q(symops, n2, (n3 == 1) ? 1.0 : -1.0, q, d);
          scaled(d, qw, a);
          add_outer(coef->c[0], a, a, 1.0 / nq);
        }
    }
  } else if (geom == OD_GEOM_POLYCRYS) {
    if (num_symm == 0) {
      coef->c[0][0] = coef->c[0][1] = coef->c[0][2] = 1.0 / 3.0;
    } else {
      for (int n2 = 0; n2 < num_symm; ++n2)
        for (int n3 = 1; n3 <= 1 + N_in; ++n3) {
          const double sgn = (n3 == 1) ? 1.0 : -1.0;
          for (int r = 0; r < 3; ++r) {  // optics.f90:365-369: rows ops(r,:,N2)
            for (int i = 0; i < 3; ++i) a[i] = sgn * SYM(symops, r, i, n2);
            add_outer(coef->c[0], a, a, (1.0 / nq) / 3.0);
          }
        }
    }
  } else {  // tensor
    static const int pa[6] = {0, 1, 2, 0, 0, 1}, pb[6] = {0, 1, 2, 1, 2, 2};
    if (num_symm == 0) {
      for (int g = 0; g < 6; ++g) {
        double ea[3] = {0, 0, 0}, eb[3] = {0, 0, 0};
        ea[pa[g]] = 1.0; eb[pb[g]] = 1.0;
        double tmp[6] = {0, 0, 0, 0, 0, 0};
        add_outer(tmp, ea, eb, 1.0);
        for (int k = 0; k < 6; ++k) coef->c[g][k] = tmp[k];
      }
    } else {
      for (int n2 = 0; n2 < num_symm; ++n2)
        for (int n3 = 1; n3 <= 1 + N_in; ++n3) {
          const double sgn = (n3 == 1) ? 1.0 : -1.0;
          double r[3][3];  // g(i) = sum_j sgn*ops(i,j,N2)*M_j  (optics.f90:403-408)
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r[i][j] = sgn * SYM(symops, i, j, n2);
          for (int g = 0; g < 6; ++g) add_outer(coef->c[g], r[pa[g]], r[pb[g]], 1.0 / nq);
        }
    }
  }
  return OD_OK;
}

namespace {
// one thread per (n1, n2, ik, is); writes all N_geom components of matrix_weights(n1,n2,ik,is,:)
__global__ void make_weights_kernel(const double *__restrict__ ome, const double *__restrict__ bands, int nb, int nk, int ns,
                                    double efermi, double scissor, OpticsCoef coef, double *__restrict__ mw) {
  const size_t nb2 = (size_t)nb * nb;
  const size_t total = nb2 * nk * ns;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int n1 = (int)(t % nb);
    const int n2 = (int)((t / nb) % nb);
    const size_t ks = t / nb2;  // ik + nk*is
    const size_t ik = ks % nk, is = ks / nk;
    double out[6] = {0, 0, 0, 0, 0, 0};
    if (n2 >= n1) {
      const double e1 = bands[n1 + (size_t)nb * (is + (size_t)ns * ik)], e2 = bands[n2 + (size_t)nb * (is + (size_t)ns * ik)];
      const bool skip = (e1 > efermi && n1 != n2) || (e2 < efermi && n1 != n2);  // optics.f90:261-262
      if (!skip) {
        double factor = 0.0;
        if (n2 == n1) factor = 1.0;
        else if (e2 > efermi) factor = 1.0 / ((e2 - e1 + scissor) * (e2 - e1 + scissor));  // :267
        double P[6];
        od_ome_products(ome, n1 + (size_t)nb * n2 + nb2 * 3 * ks, nb2, P);
#pragma unroll
        for (int g = 0; g < 6; ++g)
          if (g < coef.n_geom)
            out[g] = factor * (coef.c[g][0] * P[0] + coef.c[g][1] * P[1] + coef.c[g][2] * P[2] + coef.c[g][3] * P[3] +
                               coef.c[g][4] * P[4] + coef.c[g][5] * P[5]);
      }
    }
#pragma unroll
    for (int g = 0; g < 6; ++g)
      if (g < coef.n_geom) mw[t + total * g] = out[g];
  }
}

// the diagonal of make_weights as dos_matrix_weights(N_geom, nb, nk, ns) (optics.f90:110-119): factor = 1 on the diagonal
// whatever the occupation (:261-266)
__global__ void optics_diag_weights_kernel(const double *__restrict__ ome, int nb, int nk, int ns, OpticsCoef coef, double *__restrict__ out) {
  const size_t nb2 = (size_t)nb * nb;
  const size_t total = (size_t)nb * nk * ns;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(t % nb);
    const size_t ks = t / nb;  // ik + nk*is
    double P[6];
    od_ome_products(ome, n + (size_t)nb * n + nb2 * 3 * ks, nb2, P);
    for (int g = 0; g < coef.n_geom; ++g)
      out[g + (size_t)coef.n_geom * t] = coef.c[g][0] * P[0] + coef.c[g][1] * P[1] + coef.c[g][2] * P[2] + coef.c[g][3] * P[3] +
                                         coef.c[g][4] * P[4] + coef.c[g][5] * P[5];
  }
}

// one thread per (orb, ib, ik, is)
__global__ void core_weights_kernel(const double *__restrict__ elnes, const double *__restrict__ bands, int norb, int nb, int nk,
                                    int ns, double efermi, int core_type, OpticsCoef coef, double *__restrict__ mw) {
  const size_t stride = (size_t)norb * nb;
  const size_t total = stride * nk * ns;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ib = (int)((t / norb) % nb);
    const size_t ks = t / stride;
    const size_t ik = ks % nk, is = ks / nk;
    const double e = bands[ib + (size_t)nb * (is + (size_t)ns * ik)];
    double out = 0.0;
    const bool skip = (e < efermi && core_type == OD_CORE_ABSORPTION) || (e > efermi && core_type == OD_CORE_EMISSION);  // core.f90:125-126
    if (!skip) {
      double P[6];
      od_ome_products(elnes, (t % stride) + stride * 3 * ks, stride, P);
      out = coef.c[0][0] * P[0] + coef.c[0][1] * P[1] + coef.c[0][2] * P[2] + coef.c[0][3] * P[3] + coef.c[0][4] * P[4] +
            coef.c[0][5] * P[5];
    }
    mw[t] = out;
  }
}

// matrix_weights(p, state) = sum_orb mask(orb,p) * pdos_weights(orb, state)
__global__ void projection_merge_kernel(const double *__restrict__ pw, const int *__restrict__ mask, int norb_in, int nproj,
                                        size_t nstates, double *__restrict__ mw) {
  const size_t total = nstates * nproj;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(t % nproj);
    const size_t s = t / nproj;
    double acc = 0.0;
    for (int o = 0; o < norb_in; ++o)
      if (mask[o + (size_t)norb_in * p] == 1) acc = acc + pw[o + (size_t)norb_in * s];
    mw[t] = acc;
  }
}

__global__ void gradient_from_ome_kernel(const double *__restrict__ ome, int nb, size_t nks /* 3*nk*ns */, double *__restrict__ grad) {
  const size_t total = (size_t)nb * nks;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t ib = t % nb, rest = t / nb;
    grad[t] = ome[2 * (ib + (size_t)nb * (ib + (size_t)nb * rest))];  // Re(optical_mat(ib,ib,i,ik,is))
  }
}
}  // namespace

static int grid_for(const od_ctx *ctx, size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)ctx->sm_count * 32;
  return (int)std::max<size_t>(1, std::min(g, cap));
}

static int finish_out(od_ctx *ctx, double *dev_out, double *user_out, size_t n, int out_mem) {
  if (out_mem == OD_MEM_HOST) {
    OD_CUDA(ctx, cudaMemcpyAsync(user_out, dev_out, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

extern "C" int od_make_weights(od_ctx *ctx, const double *optical_mat, int ome_mem, int geom, const double qdir[3],
                               const double *symops, int num_symm, double efermi, double scissor_op, double *matrix_weights,
                               int out_mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  if (!optical_mat || !matrix_weights) return od_fail(ctx, OD_ERR_ARG, "null array");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OpticsCoef coef;
  OD_CHECK(od_optics_coefficients(ctx, geom, qdir, symops, num_symm, &coef));
  const size_t nel = (size_t)ctx->nb * ctx->nb * ctx->nk * ctx->ns;
  const void *d = nullptr;
  OD_CHECK(od_stage_in(ctx, optical_mat, sizeof(double) * 2 * 3 * nel, ome_mem, ctx->in_tmp, &d));
  double *out = matrix_weights;
  if (out_mem == OD_MEM_HOST) {
    OD_CHECK(od_buf_ensure(ctx, ctx->mw_tmp, sizeof(double) * nel * coef.n_geom));
    out = ctx->mw_tmp.as<double>();
  }
  make_weights_kernel<<<grid_for(ctx, nel, 256), 256, 0, ctx->stream>>>(static_cast<const double *>(d), ctx->bands.as<double>(), ctx->nb,
                                                                       ctx->nk, ctx->ns, efermi, scissor_op, coef, out);
  OD_LAUNCH_CHECK(ctx);
  return finish_out(ctx, out, matrix_weights, nel * coef.n_geom, out_mem);
}

int od_optics_diag_weights(od_ctx *ctx, const double *ome_dev, const OpticsCoef &coef, double *out_dev) {
  const size_t nel = (size_t)ctx->nb * ctx->nk * ctx->ns;
  optics_diag_weights_kernel<<<grid_for(ctx, nel, 256), 256, 0, ctx->stream>>>(ome_dev, ctx->nb, ctx->nk, ctx->ns, coef, out_dev);
  OD_LAUNCH_CHECK(ctx);
  return OD_OK;
}

extern "C" int od_core_prepare_matrix_elements(od_ctx *ctx, const double *elnes_mat, int elnes_mem, int norb, int polar,
                                               const double qdir[3], const double *symops, int num_sym, int core_type,
                                               double efermi, double *matrix_weights, int out_mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system || !ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "system / band_energy not set");
  if (!elnes_mat || !matrix_weights || norb < 1) return od_fail(ctx, OD_ERR_ARG, "bad arguments");
  if (num_sym > 1)
    return od_fail(ctx, OD_ERR_REF, "Error: Core loss not currently able to deal with symmetry. Please re-run CASTEP without symmetry. (core.f90:106)");
  if (core_type < 0 || core_type > 2) return od_fail(ctx, OD_ERR_ARG, "unknown core_type %d", core_type);
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OpticsCoef coef;
  if (polar) {
    if (!qdir) return od_fail(ctx, OD_ERR_ARG, "core_qdir missing");
    const double qw = std::sqrt(std::pow(qdir[0], 2.0) + std::pow(qdir[1], 2.0) + std::pow(qdir[2], 2.0));
    if (qw < 0.001)
      return od_fail(ctx, OD_ERR_REF, "Error: core_prepare_matrix_elements.  please check core_qdir, norm close to zero (core.f90:117)");
    OD_CHECK(od_optics_coefficients(ctx, OD_GEOM_POLAR, qdir, symops, num_sym, &coef));  // same +-q average (core.f90:127-150)
  } else {
    OD_CHECK(od_optics_coefficients(ctx, OD_GEOM_POLYCRYS, nullptr, nullptr, 0, &coef));  // (|Mx|^2+|My|^2+|Mz|^2)/3 (:152-158)
  }
  const size_t nel = (size_t)norb * ctx->nb * ctx->nk * ctx->ns;
  const void *d = nullptr;
  OD_CHECK(od_stage_in(ctx, elnes_mat, sizeof(double) * 2 * 3 * nel, elnes_mem, ctx->in_tmp, &d));
  double *out = matrix_weights;
  if (out_mem == OD_MEM_HOST) {
    OD_CHECK(od_buf_ensure(ctx, ctx->mw_tmp, sizeof(double) * nel));
    out = ctx->mw_tmp.as<double>();
  }
  core_weights_kernel<<<grid_for(ctx, nel, 256), 256, 0, ctx->stream>>>(static_cast<const double *>(d), ctx->bands.as<double>(), norb,
                                                                       ctx->nb, ctx->nk, ctx->ns, efermi, core_type, coef, out);
  OD_LAUNCH_CHECK(ctx);
  return finish_out(ctx, out, matrix_weights, nel, out_mem);
}

extern "C" int od_projection_merge(od_ctx *ctx, const double *pdos_weights, int pw_mem, const int *mask, int norb_in, int pw_nb,
                                   int nproj, double *matrix_weights, int out_mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system has not been called");
  if (!pdos_weights || !mask || !matrix_weights || norb_in < 1 || nproj < 1 || pw_nb < 1) return od_fail(ctx, OD_ERR_ARG, "bad arguments");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  // pdos_mwab%nbands of the .pdos_bin file, which may be smaller than nbands (projection_utils.f90:84; the result is then
  // consumed with mw%nbands = pw_nb by calculate_dos, dos_utils.f90:1268)
  const size_t nstates = (size_t)pw_nb * ctx->nk * ctx->ns;
  const void *d = nullptr;
  OD_CHECK(od_stage_in(ctx, pdos_weights, sizeof(double) * nstates * norb_in, pw_mem, ctx->in_tmp, &d));
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(int) * (size_t)norb_in * nproj));
  OD_CUDA(ctx, cudaMemcpyAsync(ctx->small.p, mask, sizeof(int) * (size_t)norb_in * nproj, cudaMemcpyHostToDevice, ctx->stream));
  double *out = matrix_weights;
  if (out_mem == OD_MEM_HOST) {
    OD_CHECK(od_buf_ensure(ctx, ctx->mw_tmp, sizeof(double) * nstates * nproj));
    out = ctx->mw_tmp.as<double>();
  }
  projection_merge_kernel<<<grid_for(ctx, nstates * nproj, 256), 256, 0, ctx->stream>>>(static_cast<const double *>(d), ctx->small.as<int>(),
                                                                                     norb_in, nproj, nstates, out);
  OD_LAUNCH_CHECK(ctx);
  return finish_out(ctx, out, matrix_weights, nstates * nproj, out_mem);
}

extern "C" int od_set_gradients_from_ome(od_ctx *ctx, const double *optical_mat, int mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system must be called first");
  if (!optical_mat) return od_fail(ctx, OD_ERR_ARG, "null optical_mat");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t nks = (size_t)3 * ctx->nk * ctx->ns;
  const void *d = nullptr;
  OD_CHECK(od_stage_in(ctx, optical_mat, sizeof(double) * 2 * (size_t)ctx->nb * ctx->nb * nks, mem, ctx->in_tmp, &d));
  ctx->grads_host = nullptr;
  if (!ctx->grads.owned) od_buf_release(ctx->grads);
  OD_CHECK(od_buf_ensure(ctx, ctx->grads, sizeof(double) * (size_t)ctx->nb * nks));
  gradient_from_ome_kernel<<<grid_for(ctx, (size_t)ctx->nb * nks, 256), 256, 0, ctx->stream>>>(static_cast<const double *>(d), ctx->nb, nks,
                                                                                            ctx->grads.as<double>());
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

// optics.f90:495-551 (interband).  The 1E-20 / 1E-30 literals are single precision in the reference
// (quirk Q11); physical constants optics.f90:68-72.  Root-only O(B) work in the reference: host.
extern "C" int od_calc_epsilon_2(od_ctx *ctx, const double *wjdos, int nbins, int ns, int n_geom, double cell_volume, double *eps2) {
  if (!wjdos || !eps2 || nbins < 1 || ns < 1 || n_geom < 1) return od_fail(ctx, OD_ERR_ARG, "bad arguments");
  const double epsilon_0 = 8.8541878176E-12, e_charge = 1.602176487E-19, pi = 3.141592653589793238462643383279502884197;
  const double c20 = (double)1E-20f, c30 = (double)1E-30f;
  const double epsilon2_const = (e_charge * pi * c20) / (cell_volume * c30 * epsilon_0);
  for (int g = 0; g < n_geom; ++g) {
    for (int ie = 0; ie < nbins; ++ie) eps2[ie + (size_t)nbins * g] = 0.0;
    for (int is = 0; is < ns; ++is)
      for (int ie = 1; ie < nbins; ++ie)  // N_energy = 2, jdos_nbins (quirk Q12)
        eps2[ie + (size_t)nbins * g] += epsilon2_const * wjdos[ie + (size_t)nbins * (is + (size_t)ns * g)];
  }
  return OD_OK;
}

// ------------------------------------------------------------------ synthetic band structures
namespace {
// one CTA per (k-point, spin): E_n = c_n + t_n * sum_i cos(2 pi k_i) (+0.3 eV for spin 2),
// dE/dk_i = -a t_n sin(2 pi k_i); (E, grad) pairs sorted ascending in E with a bitonic network.
__global__ void synth_bands_kernel(const double *__restrict__ cn, const double *__restrict__ tn, int nb, int ns, int nk_local,
                                   long long k_first, long long k_stride, int n1, int n2, int n3, double a,
                                   double *__restrict__ bands, double *__restrict__ grads) {
  extern __shared__ double sm[];
  const int P = blockDim.x;  // power of two >= nb
  double *sE = sm, *sT = sm + P;
  int *sIdx = reinterpret_cast<int *>(sm + 2 * P);
  const int ik = blockIdx.x, is = blockIdx.y, t = threadIdx.x;
  const long long kg = k_first + (long long)ik * k_stride;
  const int i3 = (int)(kg % n3), i2 = (int)((kg / n3) % n2), i1 = (int)(kg / ((long long)n3 * n2));
  const double twopi = 6.283185307179586476925286766559005768394;
  const double k1 = (2.0 * i1 - n1 + 1.0) / (2.0 * n1), k2 = (2.0 * i2 - n2 + 1.0) / (2.0 * n2), k3 = (2.0 * i3 - n3 + 1.0) / (2.0 * n3);
  const double cs = cos(twopi * k1) + cos(twopi * k2) + cos(twopi * k3);
  if (t < nb) { sE[t] = cn[t] + tn[t] * cs + (is == 1 ? 0.3 : 0.0); sT[t] = tn[t]; sIdx[t] = t; }
  else { sE[t] = 1.0e300; sT[t] = 0.0; sIdx[t] = t; }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = t ^ j;
      if (ixj > t) {
        const bool up = ((t & k) == 0);
        const double a0 = sE[t], a1 = sE[ixj];
        const bool sw = up ? (a0 > a1 || (a0 == a1 && sIdx[t] > sIdx[ixj])) : (a0 < a1 || (a0 == a1 && sIdx[t] < sIdx[ixj]));
        if (sw) {
          sE[t] = a1; sE[ixj] = a0;
          const double tt = sT[t]; sT[t] = sT[ixj]; sT[ixj] = tt;
          const int ii = sIdx[t]; sIdx[t] = sIdx[ixj]; sIdx[ixj] = ii;
        }
      }
      __syncthreads();
    }
  if (t < nb) {
    bands[t + (size_t)nb * (is + (size_t)ns * ik)] = sE[t];
    const size_t gi = t + (size_t)nb * (3 * (ik + (size_t)nk_local * is));
    grads[gi] = -a * sT[t] * sin(twopi * k1);
    grads[gi + nb] = -a * sT[t] * sin(twopi * k2);
    grads[gi + 2 * (size_t)nb] = -a * sT[t] * sin(twopi * k3);
  }
}

__global__ void fill_kernel(double *p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

inline uint64_t splitmix64_next(uint64_t &x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
}  // namespace

extern "C" int od_synth_bands(od_ctx *ctx, const int grid[3], long long k_first, int nk_local, int ns, int nb, double lattice_a,
                              uint64_t seed) {
  return od_synth_bands_strided(ctx, grid, k_first, 1, nk_local, ns, nb, lattice_a, seed);
}

extern "C" int od_synth_bands_strided(od_ctx *ctx, const int grid[3], long long k_first, long long k_stride, int nk_local, int ns, int nb,
                                      double lattice_a, uint64_t seed) {
  if (!ctx || !grid) return OD_ERR_ARG;
  if (nb < 1 || nb > 1024 || ns < 1 || ns > 2 || nk_local < 1 || k_stride < 1) return od_fail(ctx, OD_ERR_ARG, "bad synthetic dimensions");
  const long long nk_total = (long long)grid[0] * grid[1] * grid[2];
  if (k_first < 0 || k_first + (long long)(nk_local - 1) * k_stride >= nk_total) return od_fail(ctx, OD_ERR_ARG, "k-point range outside the grid");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const double twopi = 6.283185307179586476925286766559005768394;
  double recip[9] = {0};
  recip[0] = recip[4] = recip[8] = twopi / lattice_a;  // cubic cell, recip_lattice = (2 pi / a) I
  OD_CHECK(od_set_system(ctx, nk_local, ns, nb, recip, grid, nullptr, ns == 1 ? 2.0 : 1.0));
  std::vector<double> ct(2 * (size_t)nb);
  uint64_t st = seed;
  for (int n = 0; n < nb; ++n) {
    ct[n] = -10.0 + 30.0 * ((double)n + 0.5) / (double)nb;                                          // c_n
    ct[nb + n] = 0.2 + 1.3 * ((double)(splitmix64_next(st) >> 11) * (1.0 / 9007199254740992.0));    // t_n
  }
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(double) * ct.size()));
  OD_CUDA(ctx, cudaMemcpyAsync(ctx->small.p, ct.data(), sizeof(double) * ct.size(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->grads_host = nullptr;
  if (!ctx->bands.owned) od_buf_release(ctx->bands);
  if (!ctx->grads.owned) od_buf_release(ctx->grads);
  OD_CHECK(od_buf_ensure(ctx, ctx->bands, sizeof(double) * (size_t)nb * ns * nk_local));
  OD_CHECK(od_buf_ensure(ctx, ctx->grads, sizeof(double) * (size_t)nb * 3 * nk_local * ns));
  OD_CHECK(od_buf_ensure(ctx, ctx->kweight, sizeof(double) * (size_t)nk_local));
  int P = 32;
  while (P < nb) P <<= 1;
  dim3 g(nk_local, ns);
  synth_bands_kernel<<<g, P, (2 * sizeof(double) + sizeof(int)) * P, ctx->stream>>>(ctx->small.as<double>(), ctx->small.as<double>() + nb, nb, ns,
                                                                                   nk_local, k_first, k_stride, grid[0], grid[1], grid[2], lattice_a,
                                                                                   ctx->bands.as<double>(), ctx->grads.as<double>());
  OD_LAUNCH_CHECK(ctx);
  fill_kernel<<<grid_for(ctx, nk_local, 256), 256, 0, ctx->stream>>>(ctx->kweight.as<double>(), (size_t)nk_local, 1.0 / (double)nk_total);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

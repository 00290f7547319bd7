// od_fixedwide.cu -- fixed broadening with windows wider than the narrow engine takes (fixed_smearing >> dE).
//
// With one width w for every state, dos(E_j) = sum_s omega_s G(E_j - e_s) is a convolution of the state density
// with one Gaussian; the only thing that keeps it from being a convolution ON THE GRID is the offset
// delta_s = e_s - E_c of a state from its nearest grid point c.  That offset is small against w exactly when the
// windows are wide (|delta|/w <= dE/2w < 0.02 once a window exceeds 480 bins), so the Taylor series in
// eta = delta/w converges in a handful of terms:
//   G(x - delta)        = G(x) sum_p eta^p/p! He_p(z),                         z = x/w, He_p = Hermite (probabilists')
//   Phi((x - delta)/w)  = Phi(z) - phi(z) sum_{p>=1} eta^p/p! He_{p-1}(z)
// Pass 1 (HBM-bound, reads band_energy once): moments m_p[c] = sum_{s -> c} omega_s eta_s^p/p!, p = 0..P, by FP64
// reductions in L2 (neighbouring bands of a k-point land in different bins, so collisions are rare).
// Pass 2: 2 (P+1) correlations of the moment arrays with tabulated kernels T_p[d] = G(d dE) He_p(z_d) and
// U_p[d] (the reference's gaussian cut-off z^2/2 > 30 and NSWC erf, algorithms.f90:73-91, 252-306, evaluated on
// the grid offsets d); states whose erf has saturated enter through a prefix sum of m_0.
// P is chosen so that the truncation 1.09 eta_max^(P+1)/sqrt((P+1)!) is below 1e-13 of the peak (P <= 7).
// Weighted spectra (PDOS, core-loss, optics) get one more set of moment arrays per weight column, filled with
// omega_s mw(o, s) eta^p/p!, and JDOS treats a band pair as the "state"; the integrated spectrum exists for the DOS
// channel only.  Units come from the same producers as the other engines (DosProducer / JdosProducer).
// Differences from the reference: that truncation, and the cut-off being applied at |E_j - E_c| instead of
// |E_j - e_s| (a term of exp(-30) = 9.4e-14 of one state's peak).  Work: O(S) + O(B W P) instead of O(S W).
#include <algorithm>
#include <cmath>
#include <vector>

#include "od_internal.h"
#include "od_math.cuh"
#include "od_spectral.h"

namespace {

constexpr int FW_PMAX = 7;
#define FW_ERFWIN (OD_ERF_CUT * OD_SQRT_TWO)  // |z| beyond which the NSWC erf has saturated

struct FwPlan {
  int P;         // highest moment
  int pad;       // grid points added on either side of the spectrum grid
  int next;      // nbins + 2 pad
  int hg, he;    // half windows (bins) of the Gaussian cut-off and of the unsaturated erf
};

// moments: M[((ch * (P+1) + p) * ns + is) * next + (c + pad)], channel 0 = the spectrum itself, channel 1 + o =
// weight column o;  below[is] = weight of the states below the padded grid (their step covers the whole grid)
template <class Producer>
__global__ void __launch_bounds__(256) fw_moments_kernel(Producer prod, int ns, int og, int nch, GridSpec grid, double invw, FwPlan pl,
                                                         double *__restrict__ M, double *__restrict__ below) {
  const double inv_delta = 1.0 / grid.delta;
  const int z = blockIdx.y, is = z % ns, grp = z / ns;
  const size_t cstride = (size_t)(pl.P + 1) * ns * pl.next, pstride = (size_t)ns * pl.next;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < prod.units_per_z; u += (long long)gridDim.x * blockDim.x) {
    RecOut r;
    r.mode = -1;
#pragma unroll
    for (int a = 0; a < OD_MAXW; ++a) r.wt[a] = 0.0;
    prod(u, z, r);
    if (r.mode != 0) continue;  // null unit (fixed broadening only produces Gaussian records)
    const double e = r.e0, om = r.omega;
    const double cf = rint((e - grid.emin) * inv_delta);
    if (cf < -(double)pl.pad) { if (grp == 0 && below) atomicAdd(&below[is], om); continue; }
    if (cf >= (double)(grid.nbins + pl.pad)) continue;  // above everything: no contribution
    const int c = (int)cf;
    const double eta = (e - (grid.emin + (double)c * grid.delta)) * invw;
    double *m = M + (size_t)is * pl.next + (c + pl.pad);
    double pw[FW_PMAX + 1];  // omega eta^p / p!
    pw[0] = om;
    for (int p = 1; p <= pl.P; ++p) pw[p] = pw[p - 1] * (eta / (double)p);
    if (grp == 0)
      for (int p = 0; p <= pl.P; ++p) atomicAdd(m + p * pstride, pw[p]);
    for (int a = 0; a < OD_MAXW; ++a) {
      const int ch = 1 + grp * og + a;
      if (a >= og || ch >= nch) break;
      const double wa = r.wt[a];
      if (wa == 0.0) continue;
      double *mc = m + (size_t)ch * cstride;
      for (int p = 0; p <= pl.P; ++p) atomicAdd(mc + p * pstride, pw[p] * wa);
    }
  }
}

// T[p][d + he] (DOS kernels, zero beyond the Gaussian cut-off) and U[p][d + he] (integrated-DOS kernels)
__global__ void fw_tables_kernel(double w, double delta, FwPlan pl, double *__restrict__ T, double *__restrict__ U) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, nd = 2 * pl.he + 1;
  if (i >= nd) return;
  const int d = i - pl.he;
  const double x = (double)d * delta;      // E_j - E_c
  const double z = x / w;
  double he[FW_PMAX + 1];                   // He_0 .. He_P
  he[0] = 1.0;
  if (pl.P >= 1) he[1] = z;
  for (int p = 2; p <= pl.P; ++p) he[p] = z * he[p - 1] - (double)(p - 1) * he[p - 2];
  const double hz = 0.5 * (z * z);
  const double g = (hz > 30.0) ? 0.0 : OD_INV_SQRT_TWO_PI * exp(-hz) / w;          // algorithms.f90:83-89
  const double phi = OD_INV_SQRT_TWO_PI * exp(-hz);
  const double Phi = 0.5 * (1.0 + od_erf_nswc(z / OD_SQRT_TWO));                   // dos_utils.f90:1261
  for (int p = 0; p <= pl.P; ++p) {
    T[(size_t)p * nd + i] = g * he[p];
    U[(size_t)p * nd + i] = (p == 0) ? Phi : -phi * he[p - 1];
  }
}

// inclusive prefix of m_0 per spin: pre[is][c] = sum_{c' <= c} m_0[is][c']  (one warp per spin; tiny)
__global__ void fw_prefix_kernel(const double *__restrict__ M, int ns, int next, double *__restrict__ pre) {
  const int is = blockIdx.x;
  if (threadIdx.x != 0) return;
  double run = 0.0;
  for (int c = 0; c < next; ++c) {
    run += M[(size_t)is * next + c];
    pre[(size_t)is * next + c] = run;
  }
}

// channel ch of spin is:  out[j] += sum_{|d| <= hw} sum_p m_p[j - d] T_p[d];  for channel 0 with an integrated
// spectrum also  intdos[j] += below + pre[j - he - 1] + sum_{|d| <= he} sum_p m_p[j - d] U_p[d].
// grid: (bins / 128, ns * nch, nseg): the offsets d are split into segments; results are reduced with FP64 atomics.
// out_off[ch * ns + is] = offset of that spectrum in `accum`; int_off[is] likewise (or < 0: no integrated spectrum)
__global__ void __launch_bounds__(128) fw_convolve_kernel(const double *__restrict__ M, const double *__restrict__ pre,
                                                          const double *__restrict__ below, const double *__restrict__ T,
                                                          const double *__restrict__ U, int nbins, int ns, FwPlan pl, int seg_len,
                                                          const long long *__restrict__ out_off, const long long *__restrict__ int_off,
                                                          double *__restrict__ accum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, ch = blockIdx.y / ns, is = blockIdx.y % ns, seg = blockIdx.z;
  if (j >= nbins) return;
  const long long ioff = (ch == 0) ? int_off[is] : -1;
  const bool with_int = ioff >= 0;
  const int nd = 2 * pl.he + 1;
  const int i0 = seg * seg_len, i1 = min(i0 + seg_len, nd);
  const size_t pstride = (size_t)ns * pl.next;
  const double *m = M + (size_t)ch * (pl.P + 1) * pstride + (size_t)is * pl.next + pl.pad + j;  // m[p * pstride - d] = m_p[j - d]
  double sd = 0.0, si = 0.0;
  for (int i = i0; i < i1; ++i) {
    const int d = i - pl.he;
    if (!with_int && (d < -pl.hg - 1 || d > pl.hg + 1)) continue;  // beyond the Gaussian cut-off T is zero
    double a = 0.0, b = 0.0;
    for (int p = 0; p <= pl.P; ++p) {
      const double mv = m[(long long)p * (long long)pstride - d];
      a = fma(mv, T[(size_t)p * nd + i], a);
      if (with_int) b = fma(mv, U[(size_t)p * nd + i], b);
    }
    sd += a;
    si += b;
  }
  if (sd != 0.0) atomicAdd(&accum[out_off[ch * ns + is] + j], sd);
  if (with_int) {
    if (seg == 0) si += below[is] + pre[(size_t)is * pl.next + pl.pad + j - pl.he - 1];  // pad >= he + 1
    atomicAdd(&accum[ioff + j], si);
  }
}

FwPlan make_plan(double w, const GridSpec &grid) {
  FwPlan pl;
  const double eta_max = 0.5 * grid.delta / w;
  int P = 1;
  for (; P < FW_PMAX; ++P) {  // truncation after eta^P: 1.09 eta^(P+1) / sqrt((P+1)!)
    double f = 1.0;
    for (int i = 2; i <= P + 1; ++i) f *= (double)i;
    if (1.09 * pow(eta_max, P + 1) / sqrt(f) < 1e-13) break;
  }
  pl.P = P;
  pl.hg = (int)floor(OD_SQRT60 * w / grid.delta);
  pl.he = (int)ceil(FW_ERFWIN * w / grid.delta);
  pl.pad = pl.he + 2;
  pl.next = grid.nbins + 2 * pl.pad;
  return pl;
}

}  // namespace

// window (bins) of a fixed-width state; the narrow engine takes up to OD_NARROW_WMAX
// gwin: half width of a state's window in units of w (the erf window for DOS, the Gaussian cut-off for JDOS)
bool od_fixedwide_applies(const od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, double gwin) {
  if (ctx->engine == OD_ENGINE_DENSE || b.type != 0) return false;
  double w = b.fixed_w;
  if (b.fbc && w < b.delta_bins) w = b.delta_bins;
  if (!(w > 0.0)) return false;
  return 2.0 * gwin * w / grid.delta + 2.0 > (double)OD_NARROW_WMAX;
}

static double fw_width(const BroadSpec &b) { return (b.fbc && b.fixed_w < b.delta_bins) ? b.delta_bins : b.fixed_w; }

static size_t fw_moment_count(const od_ctx *ctx, const FwPlan &pl, int nch) { return (size_t)nch * (pl.P + 1) * ctx->ns * pl.next; }

// nch = 1 + number of weight columns
int od_fixedwide_begin(od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, int nch) {
  const FwPlan pl = make_plan(fw_width(b), grid);
  const size_t n = fw_moment_count(ctx, pl, nch) + 2;
  OD_CHECK(od_buf_ensure(ctx, ctx->moments, sizeof(double) * n));
  OD_CUDA(ctx, cudaMemsetAsync(ctx->moments.p, 0, sizeof(double) * n, ctx->stream));
  return OD_OK;
}

// units of `prod` (nz = ns * weight groups of `og` columns each for the DOS producer, ns for the JDOS producer)
template <class Producer>
static int fw_accumulate(od_ctx *ctx, const Producer &prod, int nz, int og, const BroadSpec &b, const GridSpec &grid, int nch, bool with_int) {
  const double w = fw_width(b);
  const FwPlan pl = make_plan(w, grid);
  if (prod.units_per_z <= 0) return OD_OK;
  double *M = ctx->moments.as<double>();
  double *below = with_int ? M + fw_moment_count(ctx, pl, nch) : nullptr;
  const int bx = (int)std::min<long long>((prod.units_per_z + 255) / 256, (long long)ctx->sm_count * 16);
  dim3 g(bx, nz);
  OD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  fw_moments_kernel<Producer><<<g, 256, 0, ctx->stream>>>(prod, ctx->ns, og, nch, grid, 1.0 / w, pl, M, below);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  OD_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->prof.prepare_ms += ms;
  return OD_OK;
}

int od_fixedwide_accumulate_dos(od_ctx *ctx, const DosProducer &prod, const GridSpec &grid, int norb) {
  const int ngroups = norb > 0 ? (norb + OD_MAXW - 1) / OD_MAXW : 1;
  return fw_accumulate(ctx, prod, ctx->ns * ngroups, OD_MAXW, prod.b, grid, 1 + norb, true);
}

int od_fixedwide_accumulate_jdos(od_ctx *ctx, const JdosProducer &prod, const GridSpec &grid, int n_geom) {
  return fw_accumulate(ctx, prod, ctx->ns, OD_MAXW, prod.b, grid, 1 + n_geom, false);
}

// out_off[ch * ns + is]: offset of the spectrum of channel ch / spin is in accum; int_off[is]: the integrated spectrum
// of channel 0, or null
int od_fixedwide_finish(od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, int nch, const long long *out_off, const long long *int_off,
                        double *accum) {
  const double w = fw_width(b);
  const FwPlan pl = make_plan(w, grid);
  const int ns = ctx->ns, nbins = grid.nbins, nd = 2 * pl.he + 1;
  double *M = ctx->moments.as<double>();
  double *below = M + fw_moment_count(ctx, pl, nch);
  // tables, prefix and the offset lists live in scratch_a: [T (P+1) nd | U (P+1) nd | pre ns next | offsets]
  const size_t ndbl = 2 * (size_t)(pl.P + 1) * nd + (size_t)ns * pl.next;
  const size_t noff = (size_t)nch * ns + ns;
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_a, sizeof(double) * ndbl + sizeof(long long) * noff));
  double *T = ctx->scratch_a.as<double>(), *U = T + (size_t)(pl.P + 1) * nd, *pre = U + (size_t)(pl.P + 1) * nd;
  long long *d_off = reinterpret_cast<long long *>(pre + (size_t)ns * pl.next);
  std::vector<long long> h_off(noff);
  for (size_t i = 0; i < (size_t)nch * ns; ++i) h_off[i] = out_off[i];
  for (int is = 0; is < ns; ++is) h_off[(size_t)nch * ns + is] = int_off ? int_off[is] : -1;
  OD_CUDA(ctx, cudaMemcpyAsync(d_off, h_off.data(), sizeof(long long) * noff, cudaMemcpyHostToDevice, ctx->stream));
  OD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  fw_tables_kernel<<<(nd + 127) / 128, 128, 0, ctx->stream>>>(w, grid.delta, pl, T, U);
  OD_LAUNCH_CHECK(ctx);
  if (int_off) {
    fw_prefix_kernel<<<ns, 32, 0, ctx->stream>>>(M, ns, pl.next, pre);
    OD_LAUNCH_CHECK(ctx);
  }
  const int jb = (nbins + 127) / 128;
  int nseg = std::max(1, std::min(64, (4 * ctx->sm_count + jb * ns * nch - 1) / (jb * ns * nch)));
  const int seg_len = (nd + nseg - 1) / nseg;
  nseg = (nd + seg_len - 1) / seg_len;
  dim3 g(jb, ns * nch, nseg);
  fw_convolve_kernel<<<g, 128, 0, ctx->stream>>>(M, pre, below, T, U, nbins, ns, pl, seg_len, d_off, d_off + (size_t)nch * ns, accum);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  OD_CUDA(ctx, cudaEventSynchronize(ctx->ev1));  // (h_off is a host vector)
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->prof.accumulate_ms += ms;
  ctx->prof.accumulate_launches += 1;
  return OD_OK;
}

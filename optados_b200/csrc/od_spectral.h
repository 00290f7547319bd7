// od_spectral.h -- internal (device-pointer) forms of the hot-loop entry points.
#pragma once
#include "od_internal.h"
#include <algorithm>
#include <vector>

#include "od_producers.cuh"

int od_make_broadspec(od_ctx *ctx, char type, const od_broadening *br, double delta_bins, bool honor_hybrid, BroadSpec *out);

// results stay on the device: *accum_out -> [dos(B,ns) | intdos(B,ns) | weighted_dos(B,ns,norb)]
int od_calculate_dos_device(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins, const od_broadening *br,
                            const double *mw_dev, int norb, int mw_nb, int mw_nk, int merge, double **accum_out);
// several broadening schemes in one sweep over k-point slabs (streams deferred gradients); see od_spectral.cu
int od_dos_accumulate(od_ctx *ctx, int ntypes, const char *types, const bool *weighted, double emin, double delta_bins, int nbins,
                      const od_broadening *br, const double *mw_dev, int norb, int mw_nb, int mw_nk, int merge, double **accum_base,
                      size_t *accum_stride, double *host_out);
// *accum_out -> [jdos(B,ns) | weighted_jdos(B,ns,n_geom)]
int od_calculate_jdos_device(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br, double efermi,
                             double scissor_op, const int *exclude_bands, int num_exclude_bands, const double *mw_dev,
                             const double *ome_dev, const OpticsCoef *coef, int n_geom, int merge, double **accum_out);

// contracted optics coefficients (optics.f90:213-435), see od_weights.cu
int od_optics_coefficients(od_ctx *ctx, int geom, const double qdir[3], const double *symops, int num_symm, OpticsCoef *coef);

// dos_matrix_weights(N_geom, nb, nk, ns) = the diagonal of make_weights (optics.f90:110-119), device in / device out
int od_optics_diag_weights(od_ctx *ctx, const double *ome_dev, const OpticsCoef &coef, double *out_dev);

// widest window (bins) the narrow engine takes; wider ones go to the dense engine or, for fixed broadening
// without weights, to the moment expansion of od_fixedwide.cu
#define OD_NARROW_WMAX 480
bool od_fixedwide_applies(const od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, double gwin);
int od_fixedwide_begin(od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, int nch);
int od_fixedwide_accumulate_dos(od_ctx *ctx, const DosProducer &prod, const GridSpec &grid, int norb);
int od_fixedwide_accumulate_jdos(od_ctx *ctx, const JdosProducer &prod, const GridSpec &grid, int n_geom);
int od_fixedwide_finish(od_ctx *ctx, const BroadSpec &b, const GridSpec &grid, int nch, const long long *out_off, const long long *int_off,
                        double *accum);

// The same launch for a unit list whose length is only known on the device (dev_units[0 / 1] = units of spin 1 / 2; the
// wide list of the narrow engine): no host synchronisation; a doslin `stop` condition is OR-ed into ctx->flags and
// reported by od_check_deferred after the caller's next synchronisation.  prod.units_per_z must be an UPPER BOUND.
template <int NW, bool INT, class Producer>
int od_run_dense_dev(od_ctx *ctx, const Producer &prod, const GridSpec &grid, int nz, const std::vector<long long> &dest, double *accum,
                     const unsigned long long *dev_units) {
  constexpr int NACC = 1 + (INT ? 1 : 0) + NW;
  const int ntiles = (grid.nbins + OD_TILE - 1) / OD_TILE;
  const long long nchunks = std::max<long long>(1, std::min<long long>(256, ((long long)ctx->sm_count * 16 + (long long)ntiles * nz - 1) / ((long long)ntiles * nz)));
  const size_t pbytes = sizeof(double) * (size_t)nchunks * nz * NACC * grid.nbins;
  OD_CHECK(od_buf_ensure(ctx, ctx->partial, pbytes));
  OD_CHECK(od_buf_ensure(ctx, ctx->dest_dev, sizeof(long long) * dest.size()));
  OD_CUDA(ctx, cudaMemcpyAsync(ctx->dest_dev.p, dest.data(), sizeof(long long) * dest.size(), cudaMemcpyHostToDevice, ctx->stream));
  OD_CHECK(od_flags_ready(ctx));
  dim3 g(ntiles, (unsigned)nchunks, nz);
  od_dense_kernel<NW, INT, Producer><<<g, OD_TILE, 0, ctx->stream>>>(prod, grid, 0, ctx->partial.as<double>(), ctx->flags.as<int>(), dev_units);
  OD_LAUNCH_CHECK(ctx);
  dim3 rg((grid.nbins + 255) / 256, nz * NACC);
  od_reduce_partials<<<rg, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), (int)nchunks, nz, NACC, grid.nbins,
                                                  ctx->dest_dev.as<long long>(), accum, dev_units);
  OD_LAUNCH_CHECK(ctx);
  return OD_OK;
}

// does the bucket table of the narrow engine fit for this grid?  (~650k bins per spin channel at most; finer grids run on
// the dense engine, as the reference accepts any dos_nbins / jdos_nbins)
bool od_narrow_grid_fits(int nbins, int ns);
// narrow (sorted, recurrence) engine for calculate_dos without weights (od_narrow.cu)
int od_narrow_dos(od_ctx *ctx, const DosProducer &prod, const GridSpec &grid, bool linear, double *accum);
// narrow engine for calculate_jdos: accum = [jdos(B,ns) | weighted_jdos(B,ns,n_geom)]
int od_narrow_jdos(od_ctx *ctx, const JdosProducer &prod, const GridSpec &grid, bool linear, int n_geom, double *accum);

// Launch the dense engine for a producer and reduce the chunk partials into `accum` (device).
// dest: nz*nacc offsets into accum (or -1).
template <int NW, bool INT, class Producer>
int od_run_dense(od_ctx *ctx, const Producer &prod, const GridSpec &grid, int nz, const std::vector<long long> &dest,
                     double *accum) {
  constexpr int NACC = 1 + (INT ? 1 : 0) + NW;
  const int ntiles = (grid.nbins + OD_TILE - 1) / OD_TILE;
  const long long units = prod.units_per_z;
  long long max_chunks = (units + OD_TILE - 1) / OD_TILE;
  long long want = ((long long)ctx->sm_count * 16 + (long long)ntiles * nz - 1) / ((long long)ntiles * nz);
  long long nchunks = std::max<long long>(1, std::min<long long>({want, max_chunks, 65535}));
  long long chunk_units = (units + nchunks - 1) / nchunks;
  chunk_units = (chunk_units + OD_TILE - 1) / OD_TILE * OD_TILE;
  nchunks = (units + chunk_units - 1) / chunk_units;
  if (nchunks < 1) nchunks = 1;
  const size_t pbytes = sizeof(double) * (size_t)nchunks * nz * NACC * grid.nbins;
  OD_CHECK(od_buf_ensure(ctx, ctx->partial, pbytes));
  OD_CHECK(od_buf_ensure(ctx, ctx->keys, sizeof(long long) * dest.size() + sizeof(int) * 4));
  int *bad = reinterpret_cast<int *>(ctx->keys.as<char>() + sizeof(long long) * dest.size());
  OD_CUDA(ctx, cudaMemcpyAsync(ctx->keys.p, dest.data(), sizeof(long long) * dest.size(), cudaMemcpyHostToDevice, ctx->stream));
  OD_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  dim3 g(ntiles, (unsigned)nchunks, nz);
  OD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  od_dense_kernel<NW, INT, Producer><<<g, OD_TILE, 0, ctx->stream>>>(prod, grid, chunk_units, ctx->partial.as<double>(), bad);
  OD_LAUNCH_CHECK(ctx);
  dim3 rg((grid.nbins + 255) / 256, nz * NACC);
  od_reduce_partials<<<rg, 256, 0, ctx->stream>>>(ctx->partial.as<double>(), (int)nchunks, nz, NACC, grid.nbins,
                                                  ctx->keys.as<long long>(), accum);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  int hbad = 0;
  OD_CUDA(ctx, cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->prof.accumulate_ms += ms;
  ctx->prof.accumulate_launches += 1;
  if (hbad) return od_fail(ctx, OD_ERR_DOSLIN, "doslin: input arguments incorrect (dos_utils.f90:1431/1514)");
  return OD_OK;
}


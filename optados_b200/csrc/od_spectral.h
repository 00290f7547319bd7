// od_spectral.h -- internal (device-pointer) forms of the hot-loop entry points.
#pragma once
#include "od_internal.h"
#include "od_producers.cuh"

int od_make_broadspec(od_ctx *ctx, char type, const od_broadening *br, double delta_bins, bool honor_hybrid, BroadSpec *out);

// results stay on the device: *accum_out -> [dos(B,ns) | intdos(B,ns) | weighted_dos(B,ns,norb)]
int od_calculate_dos_device(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins, const od_broadening *br,
                            const double *mw_dev, int norb, int mw_nb, int mw_nk, int merge, double **accum_out);
// *accum_out -> [jdos(B,ns) | weighted_jdos(B,ns,n_geom)]
int od_calculate_jdos_device(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br, double efermi,
                             double scissor_op, const int *exclude_bands, int num_exclude_bands, const double *mw_dev,
                             const double *ome_dev, const OpticsCoef *coef, int n_geom, int merge, double **accum_out);

// contracted optics coefficients (optics.f90:213-435), see od_weights.cu
int od_optics_coefficients(od_ctx *ctx, int geom, const double qdir[3], const double *symops, int num_symm, OpticsCoef *coef);

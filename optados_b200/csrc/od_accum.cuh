// od_accum.cuh -- the "dense" (bin-owner) accumulation engine.
//
// One *record* is one broadened unit: a (k, spin, band) state for DOS/PDOS/core, an occupied ->
// unoccupied band pair for JDOS/optics.  A CTA owns OD_TILE_BINS consecutive energy bins of one
// spin channel and one chunk of the unit list; every thread owns ONE bin and keeps its partial
// sums in registers, so the inner loop has no atomics and no shared-memory accumulation at all.
// Units are produced 128 at a time into shared memory by a Producer functor (one unit per thread,
// null units compacted away with a ballot scan), then every warp walks the staged records and
// skips those whose window does not reach its 32 bins (warp-uniform test).  Per-bin evaluation is
// the reference's arithmetic (od_math.cuh), including the exact Gaussian cut-off and the NSWC erf.
// Chunk partials are written to HBM and summed in fixed order by od_reduce_partials, so results
// are bit-reproducible run to run.
//
// This engine is exact for any window width (it is what the reference does, minus the bins that
// are provably zero) and is used for wide windows and small inputs; the sorted narrow-window
// engine (od_narrow.cu) takes over when windows are short compared with the grid.
#pragma once
#include "od_math.cuh"

#define OD_TILE 128       // records staged per iteration == threads per CTA == bins per CTA
#define OD_MAXW 8         // max weight columns carried per record

struct RecOut {
  int mode;  // -1 null, 0 gaussian, 1 linear
  double e0, p1, p2, p3, p4;  // gaussian: p1 = width;  linear: p1..p4 = e1..e4
  double omega;
  double wt[OD_MAXW];
};

struct GridSpec {
  double emin, delta;
  int nbins;
};

// window of a record on the grid (inclusive bin indices, padded by one bin on each side; the exact
// tests are repeated per bin by the faithful evaluators)
__device__ __forceinline__ void od_rec_window(const RecOut &r, const GridSpec &g, int &jlo, int &jhi, int &klo, int &khi) {
  double dlo, dhi, clo, chi;
  if (r.mode == 0) {
    dlo = r.e0 - OD_SQRT60 * r.p1;
    dhi = r.e0 + OD_SQRT60 * r.p1;
    clo = r.e0 - (OD_ERF_CUT * OD_SQRT_TWO) * r.p1;
    chi = r.e0 + (OD_ERF_CUT * OD_SQRT_TWO) * r.p1;
  } else {
    dlo = clo = r.p4;
    dhi = chi = 2.0 * r.e0 - r.p4;
  }
  const double inv = 1.0 / g.delta;
  const double big = 2.0e9;
  double a = floor((dlo - g.emin) * inv) - 1.0, b = ceil((dhi - g.emin) * inv) + 1.0;
  double c = floor((clo - g.emin) * inv) - 1.0, d = ceil((chi - g.emin) * inv) + 1.0;
  a = fmin(fmax(a, -big), big); b = fmin(fmax(b, -big), big);
  c = fmin(fmax(c, -big), big); d = fmin(fmax(d, -big), big);
  jlo = (int)a; jhi = (int)b; klo = (int)c; khi = (int)d;
}

// NW  : weight columns accumulated per record (0..OD_MAXW)
// INT : also accumulate the integrated spectrum
// Producer: struct with  __device__ void operator()(long long unit, int z, RecOut &r) const
//           and members  long long units_per_z;
template <int NW, bool INT, class Producer>
__global__ void __launch_bounds__(OD_TILE)
od_dense_kernel(Producer prod, GridSpec grid, long long chunk_units, double *__restrict__ partial,
                int *__restrict__ bad_flag, const unsigned long long *__restrict__ dev_units = nullptr) {
  constexpr int NACC = 1 + (INT ? 1 : 0) + NW;
  __shared__ double s_e0[OD_TILE], s_p1[OD_TILE], s_p2[OD_TILE], s_p3[OD_TILE], s_p4[OD_TILE], s_om[OD_TILE];
  __shared__ double s_wt[NW > 0 ? NW : 1][OD_TILE];
  __shared__ int s_jlo[OD_TILE], s_jhi[OD_TILE], s_klo[OD_TILE], s_khi[OD_TILE], s_mode[OD_TILE];
  __shared__ int s_warp_cnt[OD_TILE / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int z = blockIdx.z;
  const int j = blockIdx.x * OD_TILE + tid;                 // my bin
  const int wlo = blockIdx.x * OD_TILE + warp * 32, whi = wlo + 31;  // my warp's bins
  const double Ej = grid.emin + (double)j * grid.delta;      // dos_utils.f90:996
  // dev_units: the unit count is only known on the device (the wide list of the narrow engine, [0] / [1] per spin): the
  // chunking is then derived here, and a launch over an empty list returns at once (od_reduce_partials skips it too)
  long long upz = prod.units_per_z;
  if (dev_units) {
    upz = (long long)max(dev_units[0], dev_units[1]);
    if (upz == 0) return;
    chunk_units = ((upz + gridDim.y - 1) / gridDim.y + OD_TILE - 1) / OD_TILE * OD_TILE;
  }
  const long long u0 = (long long)blockIdx.y * chunk_units;
  long long u1 = u0 + chunk_units;
  if (u1 > upz) u1 = upz;

  double acc_d = 0.0, acc_i = 0.0;
  double acc_w[NW > 0 ? NW : 1];
#pragma unroll
  for (int a = 0; a < (NW > 0 ? NW : 1); ++a) acc_w[a] = 0.0;
  int bad = 0;

  for (long long base = u0; base < u1; base += OD_TILE) {
    // ---- produce one record per thread, compact the valid ones into shared memory
    RecOut r;
    r.mode = -1;
    const long long u = base + tid;
    if (u < u1) prod(u, z, r);
    int jlo = 0, jhi = -1, klo = 0, khi = -1;
    if (r.mode >= 0) {
      od_rec_window(r, grid, jlo, jhi, klo, khi);
      // a record entirely above the grid contributes nothing; one entirely below still adds omega to
      // every bin of the integrated spectrum, so it is only dropped when INT is off
      if (klo > grid.nbins - 1) r.mode = -1;
      if (!INT && khi < 0) r.mode = -1;
    }
    const unsigned valid = __ballot_sync(0xffffffffu, r.mode >= 0);
    if (lane == 0) s_warp_cnt[warp] = __popc(valid);
    __syncthreads();  // (also protects the previous tile's records until everyone has consumed them)
    int off = __popc(valid & ((1u << lane) - 1u));
    int total = 0;
#pragma unroll
    for (int w = 0; w < OD_TILE / 32; ++w) {
      const int c = s_warp_cnt[w];
      if (w < warp) off += c;
      total += c;
    }
    if (r.mode >= 0) {
      s_e0[off] = r.e0; s_p1[off] = r.p1; s_p2[off] = r.p2; s_p3[off] = r.p3; s_p4[off] = r.p4;
      s_om[off] = r.omega; s_mode[off] = r.mode;
      s_jlo[off] = jlo; s_jhi[off] = jhi; s_klo[off] = klo; s_khi[off] = khi;
#pragma unroll
      for (int a = 0; a < NW; ++a) s_wt[a][off] = r.wt[a];
    }
    __syncthreads();

    // ---- consume: every warp walks the staged records
    for (int i = 0; i < total; ++i) {
      const int rk_lo = s_klo[i], rk_hi = s_khi[i];
      if (whi < rk_lo) continue;                     // warp entirely below the record
      if (wlo > rk_hi) {                             // entirely above: cumulative == 1 exactly
        if (INT) acc_i += s_om[i];
        continue;
      }
      const double e0 = s_e0[i], om = s_om[i];
      double dos_temp, cum;
      if (s_mode[i] == 0) {
        const double w = s_p1[i];
        dos_temp = od_gaussian(e0, w, Ej) * om;                                   // dos_utils.f90:1255
        if (INT) cum = 0.5 * (1.0 + od_erf_nswc((Ej - e0) / (OD_SQRT_TWO * w)));  // dos_utils.f90:1260
      } else {
        dos_temp = od_doslin(e0, s_p1[i], s_p2[i], s_p3[i], s_p4[i], Ej, &cum, &bad) * om;  // :1252
      }
      acc_d += dos_temp;
      if (INT) acc_i += om * cum;
#pragma unroll
      for (int a = 0; a < NW; ++a) acc_w[a] += dos_temp * s_wt[a][i];             // :1271 / jdos_utils.f90:389
    }
  }

  if (bad) atomicOr(bad_flag, bad);
  if (j < grid.nbins) {
    const size_t slot = ((size_t)blockIdx.y * gridDim.z + z) * NACC;
    partial[(slot + 0) * grid.nbins + j] = acc_d;
    if (INT) partial[(slot + 1) * grid.nbins + j] = acc_i;
#pragma unroll
    for (int a = 0; a < NW; ++a) partial[(slot + 1 + (INT ? 1 : 0) + a) * grid.nbins + j] = acc_w[a];
  }
}

// out[dest[z*nacc+a] + j] (+)= sum_c partial[((c*nz+z)*nacc+a)*nbins + j], chunks summed in order.
static __global__ void od_reduce_partials(const double *__restrict__ partial, int nchunks, int nz, int nacc, int nbins,
                                   const long long *__restrict__ dest, double *__restrict__ out,
                                   const unsigned long long *__restrict__ dev_units = nullptr) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int za = blockIdx.y;  // z*nacc + a
  if (j >= nbins) return;
  if (dev_units && max(dev_units[0], dev_units[1]) == 0) return;
  const long long d = dest[za];
  if (d < 0) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += partial[((size_t)c * nz * nacc + za) * nbins + j];
  out[d + j] += s;
}

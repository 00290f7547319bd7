// od_narrow.cu -- the sorted "narrow-window" accumulation engine (the fast path of calculate_dos).
//
// Regime: many units (5e8 states at BASELINE config 2) whose broadening windows (15..250 bins) are
// short compared with the energy grid (20 000 bins).  Every unit then touches ~0.5 % of the grid,
// so the reference's dense loop does 200x more work than needed, and a scatter into the grid with
// atomics would serialise on hot bins.  This engine instead
//   1. PREPARES (no global atomics): pass A classifies every unit -- window on the grid, step bin,
//      narrow / wide / off-grid -- and histograms the narrow ones in shared memory by
//      (width class, first cell of the window); each CTA writes its histogram as one row of a
//      [CTA][bucket] table.  A column scan turns the table into per-CTA write cursors, and pass B
//      (same unit -> CTA mapping) scatters 32-byte (Gaussian) / 64-byte (linear) records into
//      bucket order using shared-memory cursors only.  A CTA walks the k-points of one group of 32
//      bands (PrepMap): it then fills few buckets, and its scattered records complete their lines
//      in L2.  Sorting by width class as well as by position makes the records of a warp-group
//      nearly congruent, which is what lets a lane run a record without idle lanes.
//   2. ACCUMULATES: warps pull chunks of 256 consecutive records and regroup them by a counting
//      sort on (window start, length).  Within a group of 64 records LANE = TWO RECORDS: every
//      lane walks all bins of the group's common window, the lanes starting at staggered bins and
//      wrapping round, so that in any one step they touch DIFFERENT bins of a warp-private
//      shared-memory tile -- plain LDS/STS read-modify-write, no atomics, no reductions.  Windows
//      that fit half a tile run on two copies of it (one per half-warp: starts staggered over 16
//      bins, shortest walk 16 steps).  Walking consecutive bins lets a lane advance its Gaussian by
//      the recurrence g(j+1) = g(j) r(j), r(j+1) = r(j) q  (q = exp(-dE^2/w^2)): two multiplies
//      per bin instead of an exp.  The integrated spectrum:
//        - windows under 41 bins (dE/w > 0.4): Phi(z) = H(z >= 0) + [Phi(z) - H]; the bracket is as
//          local as the DOS and shares its Gaussian, -+ (1/2) g' R(|z|), R = P5/Q6 ~ erfcx(|z|/sqrt 2)
//          (tools/fit_erfcx.py, 5.4e-15), one reciprocal per bin; the steps H go to a warp-private
//          step tile once per record and become a prefix sum at the very end;
//        - wider windows (nearly all of a fine grid): NO per-bin work -- the kernel deposits only the
//          DOS of those groups (two bins per step), and their integral follows from the summed DOS by
//          a prefix sum and one 40-tap Euler-Maclaurin filter on the grid (c_em_fir, combine_kernel).
//      The linear scheme evaluates doslin as two smooth forms with clamped arguments and one select.
//      JDOS / optics / weighted DOS run the same loops in channel passes (JD), many projectors through
//      a banded GEMM on the FP64 tensor cores (pdos_gemm_kernel).  Tiles are flushed to HBM with FP64
//      reductions only when a warp's window slides off its tile or its chunk ends.
//   3. Units whose window is wider than 480 bins (and degenerate ones) go to the dense bin-owner
//      engine (od_accum.cuh) through a compacted list; the partial spectra are then combined.
//
// Numerics vs the reference (tolerances: 1e-9 of peak on spectra, 1e-10 relative on intDOS):
//   * the recurrence accumulates ~n^2/2 ulp over n <= 512 steps: < 2e-11 relative on a value that
//     is itself <= one state's peak;  R carries 5.4e-15 absolute on Phi;
//   * the reference truncates the Gaussian at z^2/2 > 30 and the erf at |x| > 5.8; here a record
//     contributes its true (untruncated) tails over the common window of its group: a difference
//     below exp(-30) = 9.4e-14 of that state's own peak per bin;
//   * summation order inside a bucket follows the shared-memory cursor order, so results are
//     reproducible to rounding (~1e-16 relative), not bit for bit.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <type_traits>

#include "od_internal.h"
#include "od_spectral.h"

namespace {

#ifndef OD_MAIN_UNROLL
#define OD_MAIN_UNROLL 2
#endif
constexpr int NR_MAIN_UNROLL = OD_MAIN_UNROLL;  // main (wrap-free) walks
constexpr int NR_T = 512;        // bins per warp tile
constexpr int NR_WMAX = OD_NARROW_WMAX;     // widest window (bins) the narrow engine takes (Wc + 1 <= NR_T - cell)
constexpr int NR_CHUNK = 256;   // records per work item (512: +1 %, 1024: +3 % on config 2)
constexpr int NR_WARPS = 8;      // warps per CTA
constexpr int NR_CTAS_PER_SM = 2; // accumulate kernels: 2 x 8 warps per SM (registers and shared memory both allow no more)
constexpr int NR_NCLS = 10;      // width classes
constexpr int NR_MAXBUCKETS = 52224;  // x 4 B = 204 KB of shared memory in the prepare passes
#ifndef OD_PREP_THREADS
#define OD_PREP_THREADS 512
#endif
constexpr int NR_PREP_THREADS = OD_PREP_THREADS;   // x PREP_U units per thread in flight (the 204 KB bucket table allows one CTA per SM)
constexpr unsigned KEY_WIDE = 0xFFFFFFFEu, KEY_DROP = 0xFFFFFFFFu;

// With PrepSpec.want_unit (PDOS / core: the weights are gathered from matrix_weights by unit index) a record carries its
// unit index instead of a separate sorted index array (one scattered store request less per unit): a linear record in its
// padding, a Gaussian one in the high word of its tail, whose low word then packs klo (23 bits) | wlen << 23 -- the step
// position cpos is recomputed from e by the one consumer that needs it.
struct __align__(32) GRec {  // Gaussian record
  double e, w, om;
  int klo;                    // first bin of the window, global numbering (spin * bpad + bin)
  unsigned short cpos, wlen;  // step bin relative to klo; window length
};
struct __align__(64) LRec {  // linear record
  double e0, e1, e2, e3, e4, om;
  int klo;
  unsigned short cpos, wlen;
  double pad;
};

static_assert(sizeof(GRec) == 32 && offsetof(GRec, klo) == 24 && sizeof(LRec) == 64 && offsetof(LRec, klo) == 48, "record layouts the 32-byte stores of the scatter pass write");

struct PrepSpec {
  GridSpec grid;
  int bpad;      // padded bins per spin (>= nbins + NR_T)
  int cell;      // bins per position cell (power of two)
  int cshift;    // log2(cell)
  int ncells;    // ns * bpad / cell
  int nbuckets;  // NR_NCLS * ncells; bucket = class * ncells + cell
  double gwin;      // half-width of a Gaussian window in units of w (erf window for DOS, cut-off for JDOS)
  int with_int;     // integrated spectrum wanted (DOS): off-grid units still contribute their step
  int linear_pass;  // a linear pass sends Gaussian records (hybrid_linear fall-back) to the dense engine
  int ng;           // weight columns carried beside each record (JDOS / optics)
  int wstride;      // doubles per weight row: 0 (none), 1, 4 or 8
  int want_unit;    // the records carry their unit index (PDOS / core: weights are gathered from matrix_weights by it; see GRec)
  int wmax;         // widest window the tiles take with this cell: min(NR_WMAX, NR_T - 2 * cell)
  double inv_delta; // 1 / grid.delta (IEEE, host): a per-thread division is ~30 instructions of a 200-instruction pass
};

struct Cls {
  int klo, wlen, cpos, cstep;  // cstep: absolute step bin in [0, nbins]
  int kind;                    // 0 narrow, 1 wide, 2 off the grid (step only)
};

// unit -> (CTA, step, thread) mapping of the prepare passes (pass A and pass B MUST share it: the write cursors of pass B are
// per CTA).  The lanes of a warp take 32 CONSECUTIVE units of ONE k-point (consecutive bands / pair slots: 256 B per input
// array and load), a "row"; rows are numbered band-group-major, row = group * nk + k, and a CTA takes one contiguous range of
// rows: it walks the k-points of (mostly) one group of 32 bands.  A band lives in a small part of the energy grid, so a CTA
// then fills a few hundred buckets instead of all of them, and along a regular k-mesh consecutive records of a lane fall
// into the same bucket: the 32-byte records of pass B complete their 128-byte lines in L2 within microseconds.  With the
// k-major mapping of round 1 (a CTA = consecutive units = all bands of a few k-points) every CTA had every bucket open
// (148 x 17 000 half-written lines = 320 MB against 126 MB of L2) and the scatter ran at the rate of random 32-byte DRAM
// writes (40-50 G sectors/s, a quarter of the bandwidth: the row-activation limit).
struct PrepMap {
  unsigned W, nk;            // units per k-point, k-points of the slab
  unsigned ngroups;          // ceil(W / 32)
  int kmajor;                // rows numbered k-major, row = k * ngroups + group (JDOS: the 4096 pair slots of a k-point share its
                             // band energies and gradients, which stay in L1 while a CTA walks them; group-major rows cost
                             // config 4 a quarter more in both passes)
  FastDiv nk_div;            // / nk (group-major) or / ngroups (k-major)
  unsigned long long nrows;  // ceil(W / 32) * nk
  unsigned long long per;    // rows per CTA
};
constexpr unsigned NR_PREP_ROWS = NR_PREP_THREADS / 32;  // rows a CTA takes per step and unit slot

__device__ __forceinline__ bool prep_unit(const PrepMap &pm, unsigned long long row, unsigned long long row_end, unsigned lane, unsigned &u);
__device__ __forceinline__ bool prep_unit(const PrepMap &pm, unsigned long long row, unsigned long long row_end, unsigned lane, unsigned &u,
                                          unsigned &k, unsigned &w) {
  if (row >= row_end) return false;
  unsigned g;
  if (pm.kmajor) {
    k = pm.nk_div.div((unsigned)row);
    g = (unsigned)row - k * pm.ngroups;
  } else {
    g = pm.nk_div.div((unsigned)row);
    k = (unsigned)row - g * pm.nk;
  }
  w = g * 32u + lane;
  u = k * pm.W + w;
  return w < pm.W;
}
__device__ __forceinline__ bool prep_unit(const PrepMap &pm, unsigned long long row, unsigned long long row_end, unsigned lane, unsigned &u) {
  unsigned k, w;
  return prep_unit(pm, row, row_end, lane, u, k, w);
}

#define NR_ERFWIN (OD_ERF_CUT * OD_SQRT_TWO)  // 8.2024...: |z| beyond which the NSWC erf saturates

__device__ __forceinline__ int width_class(int wlen) {
  return (wlen > 24) + (wlen > 32) + (wlen > 44) + (wlen > 60) + (wlen > 84) + (wlen > 116) + (wlen > 160) + (wlen > 224) + (wlen > 320);
}

__device__ __forceinline__ Cls classify(const RecOut &r, const PrepSpec &ps) {
  const GridSpec &g = ps.grid;
  Cls c;
  double lo, hi, centre;
  const double inv = ps.inv_delta;
  bool degenerate = false;
  if (r.mode == 0) {
    lo = r.e0 - ps.gwin * r.p1;
    hi = r.e0 + ps.gwin * r.p1;
    centre = ceil((r.e0 - g.emin) * inv);  // first bin with E_j >= e
    // Gaussians narrower than dE/8 (only possible with finite_bin_correction off) would overflow the
    // recurrence ratio: leave them to the exact dense engine
    degenerate = !(r.p1 * 8.0 >= g.delta) || ps.linear_pass;
  } else {
    lo = r.p4;
    hi = 2.0 * r.e0 - r.p4;
    centre = floor((r.e0 - g.emin) * inv) + 1.0;  // first bin with E_j > e0
    // "flat" cells (e1 = e2 = e3 = e4, dos_utils.f90:1458/1505) have a DOS that jumps at the window
    // edge; they only occur for gradients with two exactly-zero components: dense engine
    degenerate = !((r.p1 - r.p4) + (r.p2 - r.p4) - (r.p3 - r.p4) > 0.0);
  }
  double a = floor((lo - g.emin) * inv), b = ceil((hi - g.emin) * inv);
  if (!(hi > lo)) b = a - 1.0;  // zero-width window: only the step survives
  const double nb = (double)g.nbins;
  centre = fmin(fmax(centre, 0.0), nb);
  c.cstep = (int)centre;
  if (b < 0.0 || a > nb - 1.0 || !(b >= a)) {  // nothing of the window on the grid
    c.kind = 2; c.klo = 0; c.wlen = 0; c.cpos = 0;
    return c;
  }
  a = fmax(a, 0.0);
  b = fmin(b, nb - 1.0);
  c.klo = (int)a;
  const int khi = (int)b;
  c.wlen = khi - c.klo + 1;
  c.kind = (c.wlen <= ps.wmax && !degenerate) ? 0 : 1;
  int cp = c.cstep - c.klo;
  cp = cp < 0 ? 0 : (cp > c.wlen ? c.wlen : cp);
  c.cpos = cp;
  return c;
}

// ---------------------------------------------------------------- pass A: keys + per-CTA histogram row
// Both prepare passes walk the units in tiles of 16 rows x U (PrepMap; pass A and pass B MUST use the same unit -> CTA
// mapping: the write cursors of pass B are per CTA).  A thread issues the loads of its U units of a tile before it touches
// the first of them.  Pass A is bound by its issue slots (8 warp-instructions per unit at 52-57 % issue utilisation, 16 warps
// per SM next to a 204 KB bucket table), pass B by the number of scattered store requests (see below), neither by DRAM
// bandwidth (40-45 %).
//
// (Measured and dropped, profiles/r02_summary.md: pass A handing its records (32-byte Gaussian ones, staged in unit order) or
// their tails to pass B so that pass B need not rebuild them -- once the band-group-major rows had made pass B a matter of
// scattered store REQUESTS (~56 G/s for 32-byte ones) its instruction count no longer mattered, and the staged traffic cost
// more than the arithmetic it saved (config 2: 119.7 -> 117.7 ms without staging); pass B sorting only 4-byte indices with the
// accumulation kernel gathering staged records; ranks from pass A's atomics instead of cursor atomics in pass B.)
template <class Producer>
__global__ void __launch_bounds__(NR_PREP_THREADS)
prep_hist_kernel(Producer prod, PrepSpec ps, unsigned *__restrict__ keys, unsigned *__restrict__ table /* [gridDim.x][nbuckets] */,
                 unsigned long long *__restrict__ wide_count /* [ns] */, PrepMap pm) {
  extern __shared__ unsigned sh_hist[];
  __shared__ unsigned char s_cls[512];  // width class of a window length (9 compares + adds per unit otherwise)
  constexpr int U = Producer::PREP_U;
  static_assert(NR_WMAX < 512, "class table");
  for (int i = threadIdx.x; i < ps.nbuckets; i += blockDim.x) sh_hist[i] = 0;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s_cls[i] = (unsigned char)width_class(i);
  __syncthreads();
  unsigned long long wide0 = 0, wide1 = 0;
  const unsigned upz = (unsigned)prod.units_per_z, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long r_begin = (unsigned long long)blockIdx.x * pm.per, r_end = min(pm.nrows, r_begin + pm.per);
  for (int is = 0; is < prod.ns; ++is)
  for (unsigned long long t0 = r_begin; t0 < r_end; t0 += NR_PREP_ROWS * U) {
    typename Producer::Raw raw[U];
    unsigned uu[U];
    bool ok[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      unsigned kk, ww;
      ok[j] = prep_unit(pm, t0 + (unsigned)j * NR_PREP_ROWS + warp, r_end, lane, uu[j], kk, ww);
      if (ok[j]) prod.template load_kw<false>(kk, ww, is, raw[j]);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (!ok[j]) continue;
      const unsigned u = uu[j];
      RecOut r;
      r.mode = -1;
      prod.template make<false>((unsigned)u, is, raw[j], r);
      Cls c;
      c.kind = 2; c.klo = 0; c.wlen = 0; c.cpos = 0; c.cstep = ps.grid.nbins;
      if (r.mode >= 0) c = classify(r, ps);
      unsigned key;
      if (c.kind == 0) {
        // key = cell index * 16 + width class (bucket = class * ncells + cell index: no division in either pass)
        const unsigned cls = s_cls[c.wlen], ci = (unsigned)(is * ps.bpad + c.klo) >> ps.cshift;
        key = ci * 16u + cls;
        atomicAdd(&sh_hist[cls * (unsigned)ps.ncells + ci], 1u);
      } else if (c.kind == 1) {
        key = KEY_WIDE;
        if (is == 0) wide0++; else wide1++;
      } else {
        key = KEY_DROP;
      }
      keys[(size_t)is * upz + u] = key;
    }
  }
  __syncthreads();
  unsigned *row = table + (size_t)blockIdx.x * ps.nbuckets;
  for (int i = threadIdx.x; i < ps.nbuckets; i += blockDim.x) row[i] = sh_hist[i];
  for (int o = 16; o > 0; o >>= 1) {
    wide0 += __shfl_xor_sync(0xffffffffu, wide0, o);
    wide1 += __shfl_xor_sync(0xffffffffu, wide1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (wide0) atomicAdd(&wide_count[0], wide0);
    if (wide1) atomicAdd(&wide_count[1], wide1);
  }
}

// column scan: table[c][b] <- sum_{c' < c} table[c'][b];  colsum[b] <- sum_c table[c][b]
// Four threads share a column (a quarter of the rows each, 16 independent loads in flight per thread): the column is
// short, the round trips are what costs -- this kernel and the scan below run once per slab and scheme, between the two
// prepare passes, and were 84 us of fixed cost each time (0.25 of the 0.68 ms a config-2 step does not scale by at N = 8).
__global__ void __launch_bounds__(256) prep_colscan_kernel(unsigned *__restrict__ table, int nrows, int nbuckets,
                                                          unsigned *__restrict__ colsum) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = tid >> 2, part = tid & 3;  // (the four threads of a column are neighbours in a warp)
  const bool on = b < nbuckets;
  const int per = (nrows + 3) / 4, r0 = min(part * per, nrows), r1 = min(r0 + per, nrows);
  unsigned sum = 0;
  if (on)
    for (int c0 = r0; c0 < r1; c0 += 16) {
      unsigned v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (c0 + i < r1) ? table[(size_t)(c0 + i) * nbuckets + b] : 0u;
#pragma unroll
      for (int i = 0; i < 16; ++i) sum += v[i];
    }
  // exclusive prefix over the four parts of the column
  const unsigned s1 = __shfl_up_sync(0xffffffffu, sum, 1), s2 = __shfl_up_sync(0xffffffffu, sum, 2), s3 = __shfl_up_sync(0xffffffffu, sum, 3);
  unsigned run = (part >= 1 ? s1 : 0u) + (part >= 2 ? s2 : 0u) + (part >= 3 ? s3 : 0u);
  if (on && part == 3) colsum[b] = run + sum;
  if (!on) return;
  for (int c0 = r0; c0 < r1; c0 += 16) {
    unsigned v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (c0 + i < r1) ? table[(size_t)(c0 + i) * nbuckets + b] : 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < r1) {
        table[(size_t)(c0 + i) * nbuckets + b] = run;
        run += v[i];
      }
  }
}

// exclusive scan of the bucket totals (single CTA; tiles of 1024 consecutive totals, the next tile's load in flight while
// the current one is scanned with shuffles); base[n] = number of narrow records
__global__ void __launch_bounds__(1024) scan_kernel(const unsigned *__restrict__ hist, int n, unsigned *__restrict__ base) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned carry = 0;
  unsigned next = threadIdx.x < n ? hist[threadIdx.x] : 0u;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    const unsigned v = next;
    next = (i + 1024 < n) ? hist[i + 1024] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = warp_tot[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive over the warps
      if (lane == 31) carry_s = wi;
    }
    __syncthreads();
    if (i < n) base[i] = carry + warp_tot[warp] + incl - v;
    carry += carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) base[n] = carry;
}

// A record leaves as 32-byte stores (st.global.v4.f64 = STG.E.ENL2.256, sm_100): the scatter pass is bound by the number of
// store transactions in flight per SM -- every record lands in a different line -- so 8-byte stores (4 / 8 per record),
// 16-byte stores (2 / 4) and 32-byte stores (1 / 2) differ in proportion.
__device__ __forceinline__ void st256(void *dst, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// (klo, cpos, wlen) as the last 8 bytes of a record: int klo | unsigned short cpos | unsigned short wlen
__device__ __forceinline__ double rec_tail(int klo, int cpos, int wlen) {
  return __hiloint2double((int)((unsigned)cpos | ((unsigned)wlen << 16)), klo);
}

// ---------------------------------------------------------------- pass B: scatter (shared-memory cursors)
// The cursors of a tile's units are taken (shared-memory atomics with return) right after its keys are known, before the
// records are built: a store that waits for its own cursor was 16 % of this pass's stall samples.  What bounds the pass is
// the NUMBER of scattered store requests it issues -- one per 32 bytes of record, ~56 G/s chip-wide whatever else the kernel
// does (a pass that only permutes staged records: 0.57 ms per 32M Gaussian records; this one, which rebuilds them: 0.59 ms;
// one that scatters 4-byte indices: 0.36 ms with or without cursor atomics, at 512 or 1024 threads; neighbouring lanes swapping
// halves so that every store instruction writes whole 64-byte linear records: 0.87 -> 0.90 ms, so the unit is the 32-byte
// sector, not the instruction).
template <bool LINEAR, class Producer>
__global__ void __launch_bounds__(NR_PREP_THREADS)
prep_scatter_kernel(Producer prod, PrepSpec ps, const unsigned *__restrict__ keys, const unsigned *__restrict__ table,
                    const unsigned *__restrict__ base, void *__restrict__ recs_v, double *__restrict__ wts,
                    double *__restrict__ stephist,
                    unsigned *__restrict__ widelist, unsigned long long *__restrict__ wide_cursor, long long wide_stride,
                    PrepMap pm) {
  extern __shared__ unsigned sh_cur[];
  constexpr int U = Producer::PREP_U;
  const unsigned *row = table + (size_t)blockIdx.x * ps.nbuckets;
  for (int i = threadIdx.x; i < ps.nbuckets; i += blockDim.x) sh_cur[i] = base[i] + row[i];
  __syncthreads();
  const unsigned upz = (unsigned)prod.units_per_z, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long r_begin = (unsigned long long)blockIdx.x * pm.per, r_end = min(pm.nrows, r_begin + pm.per);
  for (int is = 0; is < prod.ns; ++is) {
  unsigned key_next[U], k_next[U], w_next[U];  // (the unit index is k * W + w)
  bool ok_next[U];
  auto load_keys = [&](unsigned long long t0) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      unsigned un;
      ok_next[j] = prep_unit(pm, t0 + (unsigned)j * NR_PREP_ROWS + warp, r_end, lane, un, k_next[j], w_next[j]);
      key_next[j] = ok_next[j] ? keys[(size_t)is * upz + un] : KEY_WIDE;  // (KEY_WIDE: nothing to load; the slot is skipped below)
    }
  };
  load_keys(r_begin);
  for (unsigned long long t0 = r_begin; t0 < r_end; t0 += NR_PREP_ROWS * U) {
    typename Producer::Raw raw[U];
    unsigned key[U], pos[U], uu[U];
    bool ok[U];
    // the keys of this tile were loaded one iteration ago; issue every input load of the tile, then the keys of the
    // next tile, before the first use: two levels of loads in flight per thread
#pragma unroll
    for (int j = 0; j < U; ++j) {
      key[j] = key_next[j]; uu[j] = k_next[j] * pm.W + w_next[j]; ok[j] = ok_next[j];
      if (key[j] != KEY_WIDE && (key[j] != KEY_DROP || ps.with_int)) prod.template load_kw<true>(k_next[j], w_next[j], is, raw[j]);
    }
    load_keys(t0 + NR_PREP_ROWS * U);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      pos[j] = 0;
      if (key[j] < KEY_WIDE) pos[j] = atomicAdd(&sh_cur[(key[j] & 15u) * (unsigned)ps.ncells + (key[j] >> 4)], 1u);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (!ok[j]) continue;
      const unsigned u = uu[j];
      if (key[j] == KEY_WIDE) {
        const unsigned long long wp = atomicAdd(&wide_cursor[is], 1ull);
        widelist[(size_t)is * wide_stride + wp] = u;
        continue;
      }
      if (key[j] == KEY_DROP && !ps.with_int) continue;  // null pair / off-grid unit without a step
      RecOut r;
      r.mode = -1;
      prod.template make<true>(u, is, raw[j], r);
      if (r.mode < 0) continue;
      const Cls c = classify(r, ps);
      if (key[j] == KEY_DROP) {  // off the grid: only its step (rare)
        if (c.cstep < ps.grid.nbins) atomicAdd(&stephist[(size_t)is * ps.bpad + c.cstep], r.omega);
        continue;
      }
      // keep the record inside the cell pass A counted it in, whatever this pass recomputed
      const unsigned ci = key[j] >> 4;
      const int cell_lo = (int)(ci << ps.cshift) - is * ps.bpad;
      const int kl = min(max(c.klo, cell_lo), cell_lo + ps.cell - 1);
      const int wlen = min(max(c.wlen - (kl - c.klo), 1), ps.wmax);
      const int cpos = min(max(c.cstep - kl, 0), wlen);
      const int klo = is * ps.bpad + kl;  // global numbering (spin * bpad + bin)
      const unsigned p = pos[j];
      if (LINEAR) {
        LRec *o = reinterpret_cast<LRec *>(recs_v) + p;
        st256(o, r.e0, r.p1, r.p2, r.p3);
        st256(reinterpret_cast<char *>(o) + 32, r.p4, r.omega, rec_tail(klo, cpos, wlen), ps.want_unit ? __hiloint2double(0, (int)u) : 0.0);
      } else {
        const double tail = ps.want_unit ? __hiloint2double((int)u, klo | (wlen << 23)) : rec_tail(klo, cpos, wlen);
        st256(reinterpret_cast<GRec *>(recs_v) + p, r.e0, r.p1, r.omega, tail);
      }
      // weight columns ride beside the record in rows of 1, 4 or 8 doubles: one / two 32-byte stores instead of ng 8-byte ones
      // (static indices: a loop over ps.ng would put the whole record into local memory)
      if (ps.wstride == 1) wts[p] = r.wt[0];
      else if (ps.wstride == 4) st256(wts + (size_t)p * 4, r.wt[0], r.wt[1], r.wt[2], r.wt[3]);
      else if (ps.wstride == 8) {
        st256(wts + (size_t)p * 8, r.wt[0], r.wt[1], r.wt[2], r.wt[3]);
        st256(wts + (size_t)p * 8 + 4, r.wt[4], r.wt[5], 0.0, 0.0);
      }
    }
  }
  }
}

// ---------------------------------------------------------------- accumulate
// R(u) = P5(u)/Q6(u) ~ erfcx(u / sqrt 2), u in [0, 8.45]; all coefficients positive (tools/fit_erfcx.py).
// Kept in constant memory so that they enter the FMAs as constant-bank operands (no register, no UMOV).
static const double h_P[6] = {913.67854424431436856, 802.96638201765452004, 355.14175547719036648,
                              89.133859365727141543, 12.593079456648224345, 0.7978004363933496607};
static const double h_Q[6] = {913.67854424432351188, 1531.976386005259472, 1120.6427893189240945,
                              460.29258012431025649, 112.76898276075761223, 15.779603335732765349};

__device__ __forceinline__ double rcp_newton(double q) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  const double e = fma(-q, y, 1.0);
  return fma(y, e, y);  // ~2^-46 relative
}

// 1/x to ~1 ulp for normal x well inside the exponent range (per-record set-up: windows, widths, doslin
// normalisation); an IEEE division costs ~25 instructions and a slow-path branch
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

struct ErfcxCoef {
  double P[6], Q[6];
};

// cf points into the kernel parameter block, so the coefficients enter the FMAs as c[0x0][..] operands
__device__ __forceinline__ double erfcx_ratio(double u, const ErfcxCoef &cf) {
  double p = cf.P[5];
  p = fma(p, u, cf.P[4]);
  p = fma(p, u, cf.P[3]);
  p = fma(p, u, cf.P[2]);
  p = fma(p, u, cf.P[1]);
  p = fma(p, u, cf.P[0]);
  double q = u + cf.Q[5];
  q = fma(q, u, cf.Q[4]);
  q = fma(q, u, cf.Q[3]);
  q = fma(q, u, cf.Q[2]);
  q = fma(q, u, cf.Q[1]);
  q = fma(q, u, cf.Q[0]);
  return p * rcp_newton(q);
}

__device__ __forceinline__ double flip_sign_if(double v, bool flip) {
  int hi = __double2hiint(v);
  hi ^= flip ? 0x80000000 : 0;
  return __hiloint2double(hi, __double2loint(v));
}

struct NarrowArgs {
  const void *recs;
  const unsigned *nrec_dev;  // number of records, left on the device by the prepare passes
  GridSpec grid;
  int bpad, cell;
  unsigned long long *next_chunk;
  unsigned long long *stats;             // [0] lane-steps spent, [1] in-window (record, bin) pairs
  ErfcxCoef cf;
  double *out_dos, *out_int, *out_step;  // [ns * bpad + NR_T]; JD mode: out_dos / out_int are the two channels
  double *out_dos_em;                    // Gaussian DOS, Euler-Maclaurin groups: their DOS (the integral follows from it, c_em_fir)
  int no_em;                             // OD_NO_EM=1: rational erfc for every group (tests)
  int debug;                             // OD_NARROW_DEBUG bit 0: every group through the record-by-record path
  int *flags;                            // ctx->flags: [0] |= doslin `stop` conditions (reported by od_check_deferred)
  // JD (JDOS / optics) mode: channel pass `chan` accumulates K*m0 and K*m1 with (m0, m1) =
  // (1, wt[0]) for chan 0 and (wt[2 chan - 1], wt[2 chan]) after that; wts[rec * ng + a]
  const double *wts;
  int ng, wstride;
  int col0, col1;  // weight column feeding m0 / m1; -1 = the constant 1, -2 = zero
  // four-channel passes (JD == 2: the tile holds 32 bytes per bin over NR_T / 2 bins): two more columns and outputs
  int col2, col3;
  double *out2, *out3;
  // PDOS / core: weights gathered from matrix_weights(norb, mw_nb, mw_nk, ns) by the unit index the records carry
  int unit_in_rec;       // records were built with want_unit (see GRec)
  const double *mw;
  int norb, mw_nb, mw_nk, nb, mw_ns;
  long long k_first;
};

// tile -> HBM.  mode 0: (dos, local part of the integrated DOS, steps) / the two JD channels; mode 2: the same in two copies
// of NR_T / 2 bins (one per half-warp, summed here); mode 1: the DOS of the Euler-Maclaurin groups alone, the tile then
// being two arrays of NR_T doubles (one per half-warp); mode 3: four JD channels per bin in two copies of NR_T / 4 bins
__device__ __forceinline__ void flush_tile(double2 *tile, double *stile, int base, int used, int mode, const NarrowArgs &A, int lane) {
  if (mode == 1) {  // two copies of NR_T doubles, one per half-warp
    double *t1 = reinterpret_cast<double *>(tile);
    for (int i = lane; i < used; i += 32) {
      const double v = t1[i] + t1[i + NR_T];
      if (v != 0.0) atomicAdd(&A.out_dos_em[base + i], v);
      t1[i] = 0.0;
      t1[i + NR_T] = 0.0;
    }
    __syncwarp();
    return;
  }
  if (mode == 3) {  // four JD channels per bin, two copies of NR_T / 4 bins (one per half-warp)
    for (int i = lane; i < used; i += 32) {
      double2 a = tile[2 * i], b = tile[2 * i + 1];
      const double2 a2 = tile[NR_T / 2 + 2 * i], b2 = tile[NR_T / 2 + 2 * i + 1];
      a.x += a2.x; a.y += a2.y; b.x += b2.x; b.y += b2.y;
      tile[NR_T / 2 + 2 * i] = make_double2(0.0, 0.0);
      tile[NR_T / 2 + 2 * i + 1] = make_double2(0.0, 0.0);
      if (a.x != 0.0) atomicAdd(&A.out_dos[base + i], a.x);
      if (a.y != 0.0) atomicAdd(&A.out_int[base + i], a.y);
      if (b.x != 0.0) atomicAdd(&A.out2[base + i], b.x);
      if (b.y != 0.0) atomicAdd(&A.out3[base + i], b.y);
      tile[2 * i] = make_double2(0.0, 0.0);
      tile[2 * i + 1] = make_double2(0.0, 0.0);
    }
    __syncwarp();
    return;
  }
  for (int i = lane; i < used; i += 32) {
    double2 v = tile[i];
    if (mode == 2) {  // two copies of NR_T / 2 bins, one per half-warp
      const double2 v2 = tile[i + NR_T / 2];
      v.x += v2.x; v.y += v2.y;
      tile[i + NR_T / 2] = make_double2(0.0, 0.0);
    }
    if (v.x != 0.0) atomicAdd(&A.out_dos[base + i], v.x);
    if (v.y != 0.0) atomicAdd(&A.out_int[base + i], v.y);
    tile[i] = make_double2(0.0, 0.0);
    const double s = stile[i];
    if (s != 0.0) atomicAdd(&A.out_step[base + i], s);
    stile[i] = 0.0;
  }
  __syncwarp();
}

// direct (non-recurrence) evaluation of one Gaussian record at one bin: used by the rare irregular groups
__device__ __forceinline__ void gauss_direct(double e, double w, double om, double Ej, int j, int cabs, const ErfcxCoef &cf, double &d,
                                             double &s) {
  const double z = (Ej - e) / w;
  const double gp = exp(-0.5 * z * z);
  d = OD_INV_SQRT_TWO_PI * gp / w * om;
  const double h = 0.5 * om * gp * erfcx_ratio(fabs(z), cf);
  s = (j < cabs) ? h : -h;
}

// Integrated DOS of the wide Gaussians without any per-bin work (tools/em_fir.py).  Euler-Maclaurin: for f sampled with
// spacing dE,   dE sum_{n<=j} f_n = F(E_j) + dE [ f_j / 2 + sum_k B_2k dE^(2k-1)/(2k)! f^(2k-1)(E_j) ],  F the running integral.
// On f = exp(i k E) the bracket is 1/2 + i phi(theta), phi = 1/theta - cot(theta/2)/2, theta = k dE.  A sum of Gaussians of
// width w >= dE / 0.4, each deposited over at least +-8 w, has the spectrum exp(-(theta w/dE)^2/2): 4e-14 at the Nyquist
// frequency, so the sampled DOS of those records determines its own integral, and the derivative series can be applied to the
// SUMMED spectrum as one antisymmetric 40-tap filter,  sum_k ... = sum_n a_n (f_{j+n} - f_{j-n})  (least-squares fit of
// 2 sum_n a_n sin(n theta) to phi under the weight of the widest spectrum; 1.3e-15 of a record's weight over all widths and
// grid offsets).  The accumulation kernel therefore deposits only g (two multiplies and an add per record and bin) for groups
// with dE/w <= EM_DELMAX whose windows are not clipped by the grid, and combine_kernel forms
//   intdos_j += dE (prefix(dos_em)_j - dos_em_j / 2 - sum_n a_n (dos_em_{j+n} - dos_em_{j-n}))
// once per call.  (Round 2 first evaluated the series per record and bin as a polynomial in z: 10-13 FP64 operations.)
constexpr double EM_DELMAX = 0.40;
constexpr int EM_NTAPS = 40;
__constant__ double c_em_fir[EM_NTAPS] = {
    8.91334410894595863e-02, -4.78780582194980112e-02, 3.20342800797231569e-02, -2.36303978263413123e-02,
    1.83764680682251304e-02, -1.47459748485493179e-02, 1.20640884105840951e-02, -9.98802551019282656e-03,
    8.32593542078903688e-03, -6.96245923776002765e-03, 5.82427128159195874e-03, -4.86261777000064187e-03,
    4.04381183373205087e-03, -3.34374882732787777e-03, 2.74458465447531335e-03, -2.23264161192572412e-03,
    1.79704460262508072e-03, -1.42881152412623579e-03, 1.12023855146349289e-03, -8.64485364005221487e-04,
    6.55301778524417411e-04, -4.86858148529509716e-04, 3.53653802422202303e-04, -2.50484339417797615e-04,
    1.72451929225559677e-04, -1.15004265215845567e-04, 7.39885074993904500e-05, -4.57072027603903677e-05,
    2.69643832116552231e-05, -1.50922236765937371e-05, 7.95190288326125929e-06, -3.90647835159152347e-06,
    1.76811678738557096e-06, -7.26140962748321804e-07, 2.65201468100874104e-07, -8.37836256336142419e-08,
    2.19918392195048603e-08, -4.49858390838473364e-09, 6.37782080464529024e-10, -4.70258149603027033e-11};

constexpr int NR_NLV = 64;                                    // levels of the in-chunk counting sort
constexpr int NR_SORT_INTS = 72;                              // per warp: 65 counters (+ padding) ...
constexpr int NR_SORT_BYTES = NR_SORT_INTS * 4 + NR_CHUNK;    // ... and the sorted order (one byte per record)
static_assert(NR_CHUNK == 256, "the in-chunk sort keeps 8 records per lane and one-byte local indices");

// Accumulation.  A warp takes a chunk of 256 consecutive records, regroups them (below) and walks them in groups of
// 64: LANE = TWO RECORDS (A = sorted position g0 + lane, B = g0 + 32 + lane).  Every lane walks all Wc bins of the
// group's common window, lane l starting l bins in and wrapping round, so that in any one step the 32 lanes touch 32
// DIFFERENT bins of the warp-private tile: one LDS.128 / STS.128 read-modify-write per lane and step carries the
// contributions of both of its records.  (One record per lane, the round-1 layout, moved 32 B of shared memory per
// contribution: with the integrated spectrum down to 10 FP64 operations per bin the kernel then ran at 72 % of the
// shared-memory wavefront peak and 53 % of the FP64 pipe; two records per lane halve the wavefronts and give every lane
// two independent recurrences, which is also what the short windows of the linear scheme were missing.)
// JD: 0 = DOS + integrated DOS; 1 = two weight channels per pass (K m0, K m1); 2 = four channels per pass on two copies
// (one per half-warp) of a tile of NR_T / 4 bins -- the optics tensor (JDOS + 6 components) then walks its records twice
// instead of four times.
template <bool LINEAR, int JD>
__global__ void __launch_bounds__(NR_WARPS * 32, NR_CTAS_PER_SM) narrow_kernel(NarrowArgs A) {
  constexpr bool JD4 = JD == 2;
  constexpr int CAP = JD4 ? NR_T / 4 : NR_T;  // bins of the tile (of one copy)
  extern __shared__ double2 sh_tiles[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 *tile = sh_tiles + (size_t)warp * NR_T;
  double *stile = reinterpret_cast<double *>(sh_tiles + (size_t)NR_WARPS * NR_T) + (size_t)warp * NR_T;
  unsigned char *sort_area = reinterpret_cast<unsigned char *>(reinterpret_cast<double *>(sh_tiles + (size_t)NR_WARPS * NR_T) + (size_t)NR_WARPS * NR_T) +
                             (size_t)warp * NR_SORT_BYTES;
  int *cnt = reinterpret_cast<int *>(sort_area);
  unsigned char *order = sort_area + NR_SORT_INTS * 4;
  for (int i = lane; i < NR_T; i += 32) { tile[i] = make_double2(0.0, 0.0); stile[i] = 0.0; }
  __syncwarp();
  int base = 0, used = 0, tmode = 0;
  const double delta = A.grid.delta, emin = A.grid.emin;
  const long long nrec = (long long)*A.nrec_dev;
  const long long nchunks = (nrec + NR_CHUNK - 1) / NR_CHUNK;
  // Small jobs (fewer than two chunks per warp of the grid: 600k records) are bound by the time ONE warp needs for its
  // chunk, not by throughput (a 2000-k-point step spent 0.76 of its 0.94 ms of kernels in the two accumulation launches at a
  // quarter of their large-job rate): four warps then share a chunk -- each regroups all of it (4 % of the work) and walks
  // one of its four groups of 64 records.
  const int split = nchunks < 2ll * gridDim.x * NR_WARPS ? 4 : 1;
  const long long nitems = nchunks * split;
  constexpr size_t RECB = LINEAR ? sizeof(LRec) : sizeof(GRec);
  constexpr size_t TAILB = LINEAR ? 48 : 24;  // byte offset of (klo, cpos | wlen << 16) in a record
  constexpr unsigned FULL = 0xffffffffu;

  for (;;) {
    unsigned long long ch = 0;
    if (lane == 0) ch = atomicAdd(A.next_chunk, 1ull);
    ch = __shfl_sync(FULL, ch, 0);
    if ((long long)ch >= nitems) break;
    const int quarter = split == 1 ? 0 : (int)(ch & 3ull);
    const long long c0 = (long long)(split == 1 ? ch : ch >> 2) * NR_CHUNK;
    const int n = (int)min((long long)NR_CHUNK, nrec - c0);
    unsigned long long st_steps = 0, st_useful = 0;

    // ---- regroup the chunk.  The records of a chunk come from one or a few (width class, 8-bin cell) buckets, in no
    //      particular order inside a bucket, and every group walks the UNION of its records' windows: a counting sort on
    //      (window start, window length), quantised to 64 levels -- nlk start levels x 64 / nlk length levels, nlk chosen
    //      so that the groups tile the chunk's (start, length) rectangle in roughly square pieces -- makes the windows of
    //      a group nearly congruent (lane efficiency 0.80 -> 0.9 at config 2).
    {
      const char *rb = reinterpret_cast<const char *>(A.recs) + (size_t)c0 * RECB + TAILB;
      int klo_j[8], wl_j[8];
      int kmin = 0x7fffffff, kmax = -0x7fffffff, wmin = 0x7fffffff, wmax = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int li = j * 32 + lane;
        int2 t = make_int2(0, 0);
        if (li < n) t = *reinterpret_cast<const int2 *>(rb + (size_t)li * RECB);
        const bool packed = !LINEAR && A.unit_in_rec;  // tail = klo | wlen << 23, unit
        klo_j[j] = packed ? (t.x & 0x7fffff) : t.x;
        wl_j[j] = packed ? (int)((unsigned)t.x >> 23) : (int)((unsigned)t.y >> 16);
        if (li < n) { kmin = min(kmin, klo_j[j]); kmax = max(kmax, klo_j[j]); wmin = min(wmin, wl_j[j]); wmax = max(wmax, wl_j[j]); }
      }
      kmin = __reduce_min_sync(FULL, kmin); kmax = __reduce_max_sync(FULL, kmax);
      wmin = __reduce_min_sync(FULL, wmin); wmax = __reduce_max_sync(FULL, wmax);
      const int dk = kmax - kmin + 1, dw = wmax - wmin + 1;
      const long long r4 = 4ll * dk;  // 4 groups of 64 per chunk
      const int nlk = r4 < 2ll * dw ? 1 : (r4 < 8ll * dw ? 2 : 4);
      const int nlw = NR_NLV / nlk;
      const float fk = (float)nlk / (float)dk, fw = (float)nlw / (float)dw;
      cnt[lane] = 0; cnt[lane + 32] = 0;
      if (lane == 0) cnt[64] = 0;
      __syncwarp();
      int lvl_j[8], pos_j[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kl = min(nlk - 1, (int)((float)(klo_j[j] - kmin) * fk)), wl = min(nlw - 1, (int)((float)(wl_j[j] - wmin) * fw));
        lvl_j[j] = kl * nlw + wl;
        pos_j[j] = 0;
        if (j * 32 + lane < n) pos_j[j] = atomicAdd(&cnt[lvl_j[j]], 1);
      }
      __syncwarp();
      {
        const int a0 = cnt[2 * lane], a1 = cnt[2 * lane + 1];
        int incl = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        __syncwarp();
        cnt[2 * lane] = incl - a0 - a1;
        cnt[2 * lane + 1] = incl - a1;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j * 32 + lane < n) order[cnt[lvl_j[j]] + pos_j[j]] = (unsigned char)(j * 32 + lane);
      __syncwarp();
    }

    // A group whose records do not fit one tile together (a chunk that straddles two width classes or the two spin channels,
    // sparse buckets of a small job) is walked in several passes over the SAME g0: every pass takes the records that fit a
    // tile with the lowest pending one and defers the others (`defer`, `adv` = 0).  The record-by-record path below is
    // ~100x slower per record, and a single such group used to be the tail of every launch (0.5 ms).
    bool defer[2] = {false, false};
    int adv = 64, npass = 0;
    const int g_end = split == 1 ? n : min(n, 64 * quarter + 64);
    for (int g0 = 64 * quarter; g0 < g_end; g0 += adv) {
      // ---- my two records (and a prefetch of the two this lane takes in the next group)
      bool valid[2];
      long long idx[2];
      double e0[2], p1[2], p2[2], p3[2], p4[2], om[2];
      int klo[2], wlen[2], cpos[2];
      unsigned unit[2] = {0u, 0u};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int pos = g0 + 32 * s + lane;
        valid[s] = pos < n && (adv != 0 || defer[s]);
        idx[s] = c0 + (valid[s] ? (int)order[pos] : 0);
        if (pos + 64 < n) {
          const char *nx = reinterpret_cast<const char *>(A.recs) + (size_t)(c0 + order[pos + 64]) * RECB;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
        }
        e0[s] = 0; p1[s] = 1; p2[s] = p3[s] = p4[s] = 0; om[s] = 0;
        klo[s] = 0x3fffffff; wlen[s] = 0; cpos[s] = 0;
        if (valid[s]) {
          if (LINEAR) {
            const LRec r = reinterpret_cast<const LRec *>(A.recs)[idx[s]];
            e0[s] = r.e0; p1[s] = r.e1; p2[s] = r.e2; p3[s] = r.e3; p4[s] = r.e4; om[s] = r.om; klo[s] = r.klo; wlen[s] = r.wlen; cpos[s] = r.cpos;
            unit[s] = (unsigned)__double2loint(r.pad);
          } else {
            const GRec r = reinterpret_cast<const GRec *>(A.recs)[idx[s]];
            e0[s] = r.e; p1[s] = r.w; om[s] = r.om; klo[s] = r.klo; wlen[s] = r.wlen; cpos[s] = r.cpos;
            if (A.unit_in_rec) {  // packed tail (see GRec): the step position follows from e as in classify()
              unit[s] = (unsigned)r.cpos | ((unsigned)r.wlen << 16);
              wlen[s] = (int)((unsigned)r.klo >> 23);
              klo[s] = r.klo & 0x7fffff;
              const int kl = klo[s] - (klo[s] >= A.bpad ? A.bpad : 0);
              const int cstep = (int)fmin(fmax(ceil((r.e - emin) * (1.0 / delta)), 0.0), (double)A.grid.nbins);
              cpos[s] = min(max(cstep - kl, 0), wlen[s]);
            }
          }
        }
      }
      // ---- which of the pending records go together in this pass: those whose window ends inside a tile that starts at the
      //      lowest pending window, and (Gaussian) whose recurrence may start that far below their own window
      int gl = __reduce_min_sync(FULL, min(klo[0], klo[1]));
      if (!LINEAR && !JD) gl &= ~1;  // (the two-bin walk of the Euler-Maclaurin groups addresses the tile in aligned pairs)
      {
        ++npass;
        bool any_defer = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          bool sel = valid[s] && (klo[s] + wlen[s] - gl <= CAP - A.cell - 1);
          if (!LINEAR) sel = sel && ((double)(klo[s] - gl) * delta <= 25.0 * p1[s]);
          if (npass > 80) sel = valid[s];  // (cannot happen: the lowest record is always taken; then the record-by-record path)
          defer[s] = valid[s] && !sel;
          if (defer[s]) {
            valid[s] = false;
            e0[s] = 0; p1[s] = 1; p2[s] = p3[s] = p4[s] = 0; om[s] = 0;
            klo[s] = 0x3fffffff; wlen[s] = 0; cpos[s] = 0;
          }
          any_defer = any_defer || defer[s];
        }
        any_defer = __any_sync(FULL, any_defer);
        adv = any_defer ? 0 : 64;
        if (!any_defer) npass = 0;
      }
      double m0[2] = {1.0, 1.0}, m1[2] = {0.0, 0.0}, m2[2] = {0.0, 0.0}, m3[2] = {0.0, 0.0};  // JD channel multipliers
      if (JD) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!valid[s]) continue;
          const double *w = nullptr;
          if (A.mw) {
            const unsigned u = unit[s];
            const unsigned ikl = u / (unsigned)A.nb, ib = u - ikl * (unsigned)A.nb;
            const long long ik = A.k_first + ikl;
            const int is = klo[s] >= A.bpad ? 1 : 0;
            if (ik < A.mw_nk && (int)ib < A.mw_nb)  // dos_utils.f90:1268-1269
              w = A.mw + (size_t)A.norb * (ib + (size_t)A.mw_nb * (ik + (size_t)A.mw_nk * is));
          } else if (A.wts) {
            w = A.wts + (size_t)idx[s] * A.wstride;
          }
          m0[s] = A.col0 == -1 ? 1.0 : ((A.col0 >= 0 && w) ? w[A.col0] : 0.0);
          m1[s] = A.col1 == -1 ? 1.0 : ((A.col1 >= 0 && w) ? w[A.col1] : 0.0);
          if (JD4) {
            m2[s] = (A.col2 >= 0 && w) ? w[A.col2] : 0.0;
            m3[s] = (A.col3 >= 0 && w) ? w[A.col3] : 0.0;
          }
        }
      }
      // ---- the common window [gl, gl + Wc) of this pass
      const int gh = __reduce_max_sync(FULL, max(valid[0] ? klo[0] + wlen[0] - 1 : -0x3fffffff, valid[1] ? klo[1] + wlen[1] - 1 : -0x3fffffff));
      int Wc = gh - gl + 1;
      // A window that fits half a tile is walked on TWO copies of the tile, one per half-warp (summed at the flush): the
      // lanes' starts are then staggered over 16 bins instead of 32, so the shortest walk is 16 steps instead of 32 (a third
      // of the linear and JDOS groups of the BASELINE configurations have windows under 32 bins) and 15 steps of a walk
      // carry the wrap test instead of 31.
      const bool half = JD4 || Wc + 2 <= NR_T / 2 - A.cell;
      const int wfloor = half ? 16 : 32;
      if (Wc < wfloor) Wc = wfloor;
      if (!LINEAR && !JD) Wc += Wc & 1;
      int jb[2], a[2], cabs[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int spin_off = (klo[s] >= A.bpad && valid[s]) ? A.bpad : 0;  // ns <= 2
        jb[s] = valid[s] ? klo[s] - spin_off : 0;                           // my first bin on the grid
        a[s] = valid[s] ? klo[s] - gl : 0;                                  // my window starts at relative bin a
        cabs[s] = a[s] + cpos[s];                                           // first relative bin at / above my centre
      }
      // The fast loop starts every lane's recurrence a few bins BELOW its own window (relative bin
      // `lane`, and bin 0 after the wrap): harmless while that start is not so deep in the tail that
      // exp underflows.  Gaussian only; del <= 8 by construction, so this only trips on sparse data.
      bool irregular = Wc + 1 > CAP - A.cell;  // (+1: the step of a centre just above the window)
      if (!LINEAR) irregular = irregular || (valid[0] && (double)a[0] * delta > 25.0 * p1[0]) || (valid[1] && (double)a[1] * delta > 25.0 * p1[1]);
      irregular = __any_sync(FULL, irregular) || (A.debug & 1);

      if (irregular && (A.debug & 4)) continue;  // (timing experiment: drop the record-by-record groups)
      if (irregular) {
        // sparse data, a group that straddles the two spin channels, ...: record by record, lanes over
        // bins, straight to HBM.
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll 1
          for (int l = 0; l < 32; ++l) {
            const int k_l = __shfl_sync(FULL, klo[s], l), w_l = __shfl_sync(FULL, wlen[s], l), c_l = __shfl_sync(FULL, cpos[s], l);
            const int jb_l = __shfl_sync(FULL, jb[s], l);
            const double a0 = __shfl_sync(FULL, e0[s], l), a1 = __shfl_sync(FULL, p1[s], l), a2 = __shfl_sync(FULL, p2[s], l);
            const double a3 = __shfl_sync(FULL, p3[s], l), a4 = __shfl_sync(FULL, p4[s], l), ao = __shfl_sync(FULL, om[s], l);
            const double b0 = __shfl_sync(FULL, m0[s], l), b1 = __shfl_sync(FULL, m1[s], l);
            const double b2 = JD4 ? __shfl_sync(FULL, m2[s], l) : 0.0, b3 = JD4 ? __shfl_sync(FULL, m3[s], l) : 0.0;
            if (w_l == 0) continue;
            if (!JD && lane == 0 && jb_l + c_l < A.grid.nbins) atomicAdd(&A.out_step[k_l + c_l], ao);
            for (int r = lane; r < w_l; r += 32) {
              const double Ej = emin + (double)(jb_l + r) * delta;
              double d, sv;
              if (LINEAR) {
                double cum; int bad = 0;
                d = od_doslin(a0, a1, a2, a3, a4, Ej, &cum, &bad) * ao;
                sv = ((r < c_l) ? cum : cum - 1.0) * ao;
                if (bad) atomicOr(&A.flags[0], bad);
              } else {
                gauss_direct(a0, a1, ao, Ej, r, c_l, A.cf, d, sv);
              }
              if (JD4) {
                if (d * b2 != 0.0) atomicAdd(&A.out2[k_l + r], d * b2);
                if (d * b3 != 0.0) atomicAdd(&A.out3[k_l + r], d * b3);
              }
              if (JD) { sv = d * b1; d = d * b0; }
              if (d != 0.0) atomicAdd(&A.out_dos[k_l + r], d);
              if (sv != 0.0) atomicAdd(&A.out_int[k_l + r], sv);
            }
          }
        continue;
      }
      // energy of relative bin 0 (the same for both records up to rounding; each record uses its own copy so that an
      // invalid slot cannot disturb the valid one)
      double E0[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) E0[s] = emin + (double)(jb[s] - a[s]) * delta;
      if (!valid[0]) E0[0] = E0[1];
      if (!valid[1]) E0[1] = E0[0];

      // ---- Gaussian DOS: which form of the integrated spectrum does this group take?
      double invw[2] = {1.0, 1.0}, del[2] = {0.0, 0.0}, zB[2] = {0.0, 0.0};
      int emode = 0;  // 0: rational erfc (bracket + steps); 1: Euler-Maclaurin group -- only the DOS is deposited (c_em_fir)
      if (!LINEAR) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          invw[s] = 1.0 / p1[s];  // (the reciprocal by MUFU + Newton measured slower here)
          del[s] = delta * invw[s];
          zB[s] = (E0[s] - e0[s]) * invw[s];
        }
        if (!JD && !A.no_em) {
          // every record must deposit its whole Gaussian: the walk starts at or below -8 w (a window clipped by the bottom
          // of the grid does not), the record's own window ends at or above +8 w (one clipped by the top does not), and the
          // widths stay inside the filter's band
          const double dl = fmax(valid[0] ? del[0] : 0.0, valid[1] ? del[1] : 0.0);
          const double dmax = __hiloint2double((int)__reduce_max_sync(FULL, (unsigned)__double2hiint(dl)), 0);
          bool ok = true;
#pragma unroll
          for (int s = 0; s < 2; ++s)
            ok = ok && (!valid[s] || (zB[s] <= -8.0 && fma((double)(a[s] + wlen[s] - 1), del[s], zB[s]) >= 8.0));
          ok = __all_sync(FULL, ok);
          // dmax is del rounded DOWN to its high word: compare against a threshold lowered by one part in 10^6
          if (ok && dmax <= EM_DELMAX * (1.0 - 1e-6)) emode = 1;
        }
      }
      if (emode && Wc < 32) Wc = 32;  // (cannot happen: dE/w <= 0.4 means windows of >= 41 bins)
      const int want_mode = JD4 ? 3 : (emode ? 1 : (half ? 2 : 0));
      if (gl < base || gl + Wc + 1 > base + (want_mode == 3 ? NR_T / 4 : want_mode == 2 ? NR_T / 2 : NR_T) || want_mode != tmode) {
        flush_tile(tile, stile, base, used, tmode, A, lane);
        base = gl & ~(A.cell - 1);
        used = 0;
        tmode = want_mode;
      }
      const int off = gl - base;
      used = max(used, off + Wc + 1);
      st_steps += (unsigned)Wc * ((valid[0] ? 1u : 0u) + (valid[1] ? 1u : 0u));
      st_useful += (unsigned)(wlen[0] + wlen[1]);
      double2 *tw = JD4 ? tile + (lane >> 4) * (NR_T / 2) + 2 * off : tile + off + (want_mode == 2 ? (lane >> 4) * (NR_T / 2) : 0);
      const int ntail = want_mode >= 2 ? 15 : 31;  // wrap steps of the one-bin walks

      // ---- the steps H: one add per record into the warp-private step tile.  Several records of a
      //      group usually share their centre bin, so this one place uses shared-memory atomics.
      if (!JD && !emode) {
        if (valid[0]) atomicAdd(stile + off + cabs[0], om[0]);
        if (valid[1]) atomicAdd(stile + off + cabs[1], om[1]);
      }
      __syncwarp();

      // Euler-Maclaurin groups walk two bins per step: one LDS.128 / STS.128 and one warp synchronisation per four (record, bin)
      // pairs -- the one-bin walk was bound by the shared-memory round trip.  The two half-warps deposit into their own copies
      // of the tile (it holds 8 bytes per bin in this mode, so there is room for two), lane l starting at bin 2 (l mod 16): the
      // staggered starts then span 30 bins instead of 62 and only 15 steps of a walk carry the wrap test and its 8 selects
      // (with one copy the 31 wrap steps were 83 % of the instructions of an 84-bin walk).
      const bool pairwalk = !LINEAR && !JD && emode;  // (Wc is even and >= 32)
      int r = pairwalk ? 2 * (lane & 15) : (want_mode >= 2 ? (lane & 15) : lane);
      if (!LINEAR) {
        double amp[2], q[2], gB[2], rB[2], z[2], g[2], rr[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          amp[s] = om[s] * invw[s] * OD_INV_SQRT_TWO_PI;
          const double hd2 = 0.5 * del[s] * del[s];
          q[s] = exp(-del[s] * del[s]);
          gB[s] = amp[s] * exp(-0.5 * zB[s] * zB[s]);
          rB[s] = exp(-fma(zB[s], del[s], hd2));
          z[s] = zB[s] + (double)r * del[s];
          g[s] = amp[s] * exp(-0.5 * z[s] * z[s]);
          rr[s] = exp(-fma(z[s], del[s], hd2));
        }
        // Every walk below is split in two: Wc - 31 steps in which no lane can reach the end of the window (lane l starts
        // at bin l), then 31 steps in which exactly one lane per step wraps round to bin 0.  The restart values are 6
        // doubles; as predicated moves inside one loop they were 13 of 47 instructions of EVERY step.
        const int nmain = Wc - ntail;
        if (pairwalk) {
          // (same multiplications in the same order as the one-bin walk: g_{j+1} = g_j r_j, r_{j+1} = r_j q)
          double2 *tw2 = reinterpret_cast<double2 *>(reinterpret_cast<double *>(tile) + (lane >> 4) * NR_T + off);  // off is even
          auto step2 = [&](auto WRAP) {
            if (decltype(WRAP)::value) {
              if (r == Wc) { r = 0; g[0] = gB[0]; rr[0] = rB[0]; g[1] = gB[1]; rr[1] = rB[1]; }
            }
            const double a0 = g[0], b0 = g[1];
            const double a1 = a0 * rr[0], b1 = b0 * rr[1];
            const double ra = rr[0] * q[0], rb = rr[1] * q[1];
            double2 *qp = tw2 + (r >> 1);
            g[0] = a1 * ra; g[1] = b1 * rb;
            rr[0] = ra * q[0]; rr[1] = rb * q[1];
            r += 2;
            double2 acc = *qp;
            acc.x += a0 + b0;
            acc.y += a1 + b1;
            *qp = acc;
            __syncwarp();
          };
          const int nmain2 = (Wc >> 1) - 15;
#pragma unroll 2
          for (int t = 0; t < nmain2; ++t) step2(std::false_type{});
#pragma unroll 1
          for (int t = 0; t < 15; ++t) step2(std::true_type{});
          continue;
        }
        const double clA = p1[0] * 1.2533141373155002512, clB = p1[1] * 1.2533141373155002512;  // w sqrt(pi/2):  (omega/2) g' = g cl
        auto step = [&](auto WRAP) {
          if (decltype(WRAP)::value) {
            if (r == Wc) { r = 0; g[0] = gB[0]; rr[0] = rB[0]; z[0] = zB[0]; g[1] = gB[1]; rr[1] = rB[1]; z[1] = zB[1]; }
          }
          // capture this step's inputs, then advance the recurrences at once so that the next step's
          // operands are ready long before the rational chains below have drained
          const double uA = fabs(z[0]), uB = fabs(z[1]), gA = g[0], gb = g[1];
          double2 *qp = JD4 ? tw + 2 * r : tw + r;
          const bool fA = r >= cabs[0], fB = r >= cabs[1];
          g[0] *= rr[0]; rr[0] *= q[0]; z[0] += del[0];
          g[1] *= rr[1]; rr[1] *= q[1]; z[1] += del[1];
          ++r;
          double2 acc = *qp;
          if (JD4) {
            double2 acc2 = qp[1];
            acc2.x += fma(gA, m2[0], gb * m2[1]);
            acc2.y += fma(gA, m3[0], gb * m3[1]);
            qp[1] = acc2;
          }
          if (JD) {
            acc.x += fma(gA, m0[0], gb * m0[1]);
            acc.y += fma(gA, m1[0], gb * m1[1]);
          } else {
            const double hA = (gA * clA) * erfcx_ratio(uA, A.cf);
            const double hB = (gb * clB) * erfcx_ratio(uB, A.cf);
            acc.x += gA + gb;
            acc.y += flip_sign_if(hA, fA) + flip_sign_if(hB, fB);
          }
          *qp = acc;
          __syncwarp();
        };
#pragma unroll NR_MAIN_UNROLL
        for (int t = 0; t < nmain; ++t) step(std::false_type{});
#pragma unroll 1
        for (int t = 0; t < ntail; ++t) step(std::true_type{});
      } else {
        // doslin (dos_utils.f90:1391-1528) in x = e_use - e4 = dd - |E_j - e0| (the mirror of :1469), with
        // aa = e3-e4 <= bb = e2-e4 <= cc = e1-e4 <= dd = e0-e4.  The five branches collapse into two smooth
        // forms with clamped arguments (p = max(aa - x, 0), y = max(cc - x, 0)):
        //   x <  bb:  D = x - aa/2 + p^2/(2 aa),          I = x^2/2 - aa x/2 + aa^2/6 - p^3/(6 aa)
        //   x >= bb:  D = hh - y^2/(2 (cc-bb)),           I = hh x + k30 + y^3/(6 (cc-bb))
        // (pieces 1|2 and 3|4 of the reference; identical polynomials, continuity at aa / cc used once).
        // (Measured and dropped: coefficients that carry the record's weight alpha * omega, so that the walk adds D and I as
        // they come and flips the sign of I with an XOR -- 2 FP64 operations less per dual step, 3.91 -> 3.78 ms per 32M units,
        // but the golden linear DOS of the reference's test-suite then differs by 4e-12 of peak instead of < 2e-12.)
        constexpr double third = 1.0 / 3.0, sixth = 1.0 / 6.0, tiny = 1e-280;
        double aa[2], bb[2], cc[2], dd[2], hh[2], haa[2], inv2a[2], inv2cb[2], i6a[2], i6cb[2], a2_6[2], k30[2], sc[2], tB[2], tt[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          aa[s] = p3[s] - p4[s]; bb[s] = p2[s] - p4[s]; cc[s] = p1[s] - p4[s]; dd[s] = e0[s] - p4[s];
          hh[s] = 0.5 * (cc[s] + bb[s] - aa[s]); haa[s] = 0.5 * aa[s];
          const double cb = cc[s] - bb[s];
          const double full = (cc[s] - aa[s]) * cc[s] + (aa[s] * aa[s] - cb * cb) * third;
          sc[s] = valid[s] ? om[s] * fast_rcp(full + (cc[s] + bb[s] - aa[s]) * (dd[s] - cc[s])) : 0.0;  // alpha * omega
          inv2a[s] = aa[s] > tiny ? 0.5 * fast_rcp(aa[s]) : 0.0;
          inv2cb[s] = cb > tiny ? 0.5 * fast_rcp(cb) : 0.0;
          i6a[s] = inv2a[s] * third; i6cb[s] = inv2cb[s] * third;
          a2_6[s] = aa[s] * aa[s] * sixth;
          k30[s] = 0.5 * (aa[s] * aa[s] * third - cc[s] * bb[s]) - cb * cb * sixth;
          tB[s] = E0[s] - e0[s];  // E_j - e0 at relative bin 0
          tt[s] = tB[s] + (double)r * delta;
        }
        auto clamp0 = [](double v) {  // max(v, +0) on the sign / exponent word: negative -> < 2^-1022
          return __hiloint2double(max(__double2hiint(v), 0), __double2loint(v));
        };
        auto lin_eval = [&](int s, double t, double &D, double &I) {
          const double x = clamp0(dd[s] - fabs(t));  // bins outside (e4, 2 e0 - e4): x -> +0, D = I = 0
          const double p = clamp0(aa[s] - x), y = clamp0(cc[s] - x);
          const double pp = p * p, yy = y * y;
          const double Dlo = fma(pp, inv2a[s], x - haa[s]), Dhi = fma(-yy, inv2cb[s], hh[s]);
          const bool lo = x < bb[s];
          D = lo ? Dlo : Dhi;
          if (!JD) {
            const double Ilo = fma(-(pp * p), i6a[s], fma(x, fma(x, 0.5, -haa[s]), a2_6[s]));
            const double Ihi = fma(yy * y, i6cb[s], fma(hh[s], x, k30[s]));
            I = lo ? Ilo : Ihi;
          }
        };
        const double scm0A = sc[0] * m0[0], scm1A = sc[0] * m1[0], scm0B = sc[1] * m0[1], scm1B = sc[1] * m1[1];
        // the sign of the local part flips where the walk passes the record's centre: d = cabs - 1 - r goes negative there,
        // so the flip is the sign bit of a counter (one add and one logic operation per record and step)
        int dA = cabs[0] - 1 - r, dB = cabs[1] - 1 - r;
        const int nmain = Wc - ntail;
        auto step = [&](auto WRAP) {
          if (decltype(WRAP)::value) {
            if (r == Wc) { r = 0; tt[0] = tB[0]; tt[1] = tB[1]; dA = cabs[0] - 1; dB = cabs[1] - 1; }
          }
          double2 *qp = JD4 ? tw + 2 * r : tw + r;
          double2 acc = *qp;
          double DA, IA = 0, DB, IB = 0;
          lin_eval(0, tt[0], DA, IA);
          lin_eval(1, tt[1], DB, IB);
          if (JD4) {
            double2 acc2 = qp[1];
            acc2.x += fma(DA, sc[0] * m2[0], DB * (sc[1] * m2[1]));
            acc2.y += fma(DA, sc[0] * m3[0], DB * (sc[1] * m3[1]));
            qp[1] = acc2;
          }
          if (JD) {
            acc.x += fma(DA, scm0A, DB * scm0B);
            acc.y += fma(DA, scm1A, DB * scm1B);
          } else {
            const double sA = __hiloint2double(__double2hiint(sc[0]) ^ (dA & 0x80000000), __double2loint(sc[0]));
            const double sB = __hiloint2double(__double2hiint(sc[1]) ^ (dB & 0x80000000), __double2loint(sc[1]));
            acc.x += fma(DA, sc[0], DB * sc[1]);
            acc.y += fma(IA, sA, IB * sB);
          }
          *qp = acc;
          tt[0] += delta; tt[1] += delta;
          ++r; --dA; --dB;
          __syncwarp();
        };
#pragma unroll NR_MAIN_UNROLL
        for (int t = 0; t < nmain; ++t) step(std::false_type{});
#pragma unroll 1
        for (int t = 0; t < ntail; ++t) step(std::true_type{});
      }
    }
    for (int o = 16; o > 0; o >>= 1) st_useful += __shfl_xor_sync(FULL, st_useful, o);
    for (int o = 16; o > 0; o >>= 1) st_steps += __shfl_xor_sync(FULL, st_steps, o);
    if (lane == 0) { atomicAdd(&A.stats[0], st_steps); atomicAdd(&A.stats[1], st_useful); }
    // consecutive chunks of one warp are usually far apart on the grid: flush now
    flush_tile(tile, stile, base, used, tmode, A, lane);
    used = 0;
  }
}

// ---------------------------------------------------------------- many-projector contraction (PDOS / core)
// weighted_dos(bin, is, o) += sum_records K(bin; record) * matrix_weights(o, record)  is a banded GEMM: for a
// tile of 64 bins only the records whose window reaches the tile contribute, and in the sorted record array
// those are 8 contiguous ranges (one per width class).  One CTA = one work item (bin tile, orbital tile, record
// segment): K-chunks of 32 records are expanded into a dense G[32][64] tile in shared memory (Gaussian by the
// same two-multiply recurrence along bins, linear by the branch-free doslin pieces), their weight rows are
// gathered (coalesced, 8 * BN bytes per record) into W[32][BN], and the product runs on the FP64 tensor cores
// (DMMA m16n8k8, see pdos_gemm_kernel below).
constexpr int PG_BM = 64, PG_KC = 32, PG_SEG = 4096;
constexpr int PG_THREADS = 288;  // 8 DMMA warps + 1 producer warp (record parameters of the chunk after next)


struct PgItem {
  int is, b0, otile;
  unsigned rec0, rec1;
};

// The work items are enumerated on the device: a "triple" (spin, 64-bin tile, width class) owns the contiguous range of sorted
// records whose windows can reach the tile, cut into segments of PG_SEG records x orbital tiles; pg_count_kernel counts the
// items of every triple from the bucket bases, scan_kernel turns the counts into offsets, and a CTA maps an item number back
// to its triple by bisection.  (Round 1 built the list on the host: a copy of the bucket table and two stream
// synchronisations per slab.)
struct PgPlan {
  const unsigned *base;   // bucket bases of the prepare passes (S.base)
  const unsigned *offs;   // [ntriples + 1] exclusive scan of the item counts; offs[ntriples] = number of items
  int ntriples, ntiles, otiles, nbins, bpad, cell, ncells;
  int wmax[NR_NCLS];
};

__device__ __forceinline__ void pg_triple_range(const PgPlan &P, int tr, int &is, int &b0, unsigned &r0, unsigned &r1) {
  const int c = tr % NR_NCLS, tt = tr / NR_NCLS;
  const int tile = tt % P.ntiles;
  is = tt / P.ntiles;
  b0 = tile * PG_BM;
  const int g0 = is * P.bpad + b0;
  const int glo = max(is * P.bpad, g0 - P.wmax[c] + 1), ghi = min(is * P.bpad + P.nbins - 1, g0 + PG_BM - 1);
  r0 = P.base[(size_t)c * P.ncells + glo / P.cell];
  r1 = P.base[(size_t)c * P.ncells + ghi / P.cell + 1];
}

__global__ void pg_count_kernel(PgPlan P, unsigned *__restrict__ counts) {
  const int tr = blockIdx.x * blockDim.x + threadIdx.x;
  if (tr >= P.ntriples) return;
  int is, b0;
  unsigned r0, r1;
  pg_triple_range(P, tr, is, b0, r0, r1);
  counts[tr] = r1 > r0 ? ((r1 - r0 + PG_SEG - 1) / PG_SEG) * (unsigned)P.otiles : 0u;
}

struct PgArgs {
  const void *recs;
  PgPlan plan;
  unsigned long long *next_item;
  GridSpec grid;
  int bpad, ns;
  const double *mw;
  int norb, mw_nb, mw_nk, nb;
  long long k_first;
  double *wdos;  // weighted_dos(nbins, ns, norb)
};

// Shared-memory plan of one CTA (dynamic, two stages): G[2][BM][KS] (bin-major, KS = KC + 4), W[2][KC][BN + 4],
// P[2][KC][12], row[2][KC], win[2][KC][2].  The +4 paddings make every fragment load below conflict-free.
constexpr int PG_KS = PG_KC + 4;
template <int BN>
constexpr size_t pg_smem_bytes() {
  return sizeof(double) * (2 * PG_BM * PG_KS + 2 * PG_KC * (BN + 4) + 2 * PG_KC * 12) + sizeof(long long) * 2 * PG_KC +
         sizeof(int) * 4 * PG_KC + sizeof(unsigned) * 2 * (PG_BM / 8);
}

// D(16x8) += A(16x8) B(8x8) in FP64 on the tensor cores.  Fragments (g = lane / 4, t = lane % 4):
//   a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4];  b0 = B[t][g], b1 = B[t+4][g];
//   c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], double a0, double a1, double a2, double a3, double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// The contraction runs on the FP64 tensor cores (DMMA m16n8k8): with a 4x4 register tile per thread the CUDA-core
// version of this loop was bound by shared-memory wavefronts (ncu: 90 % of the LSU wavefront peak at 35 % FP64-pipe
// utilisation, profiles/), a DMMA needs 12 operand loads for 128 FMAs per thread instead of 64.  Measured DMMA peak
// on B200: 37.1 TFLOP/s (the same pipe as DFMA).
// The chunk loop is software-pipelined over three chunks with one barrier per chunk:
//   chunk c+2: its 32 records (+ unit indices) are loaded and turned into parameters (one exp, two reciprocals, one
//              division per record) by a ninth, producer warp -- as a side job of warp 0 that chain ended every chunk
//              after the other warps had reached the barrier (34 % of the kernel's stall samples);
//   chunk c+1: weight rows are loaded (coalesced) into registers before the MMA loop and parked in shared memory
//              after it; its G tile is expanded before the MMA loop;
//   chunk c  : 4 k-steps of NB DMMAs per warp (warp tile 16 bins x 8 NB orbitals).
template <int BN, bool LINEAR>
__global__ void __launch_bounds__(PG_THREADS, 2) pdos_gemm_kernel(PgArgs A) {
  constexpr int WS = BN + 4;             // row stride of the W tile
  constexpr int NB = BN / 16;            // 8-orbital blocks per warp (8 warps = 4 bin blocks x 2 orbital halves)
  constexpr int EPT = PG_KC * BN / 256;  // weight elements per thread and chunk
  extern __shared__ __align__(16) unsigned char pg_smem[];
  double *sG = reinterpret_cast<double *>(pg_smem);            // [2][BM][KS]
  double *sW = sG + 2 * PG_BM * PG_KS;                         // [2][KC][WS]
  double *sP = sW + 2 * PG_KC * WS;                            // [2][KC][12]
  long long *sRow = reinterpret_cast<long long *>(sP + 2 * PG_KC * 12);  // [2][KC]
  int *sI = reinterpret_cast<int *>(sRow + 2 * PG_KC);         // [2][KC][2]
  unsigned *sMask = reinterpret_cast<unsigned *>(sI + 4 * PG_KC);  // [2][BM / 8]: records with a non-zero entry in each 8-bin run
  __shared__ PgItem s_item;
  __shared__ int s_more;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int fg = lane >> 2, ft = lane & 3;      // fragment coordinates
  const int wm = warp >> 1, wn = warp & 1;      // warp tile: bins wm*16 .. +15, orbitals wn*8*NB .. +8*NB-1
  const double delta = A.grid.delta, emin = A.grid.emin;

  struct RecRegs {
    double e0, p1, p2, p3, p4, om;
    int klo, wlen;
    unsigned unit;
    bool ok;
  };

  for (;;) {
    __syncthreads();
    if (t == 0) {
      const PgPlan &P = A.plan;
      const unsigned it = (unsigned)atomicAdd(A.next_item, 1ull);
      s_more = it < P.offs[P.ntriples];
      if (s_more) {
        int lo = 0, hi = P.ntriples;  // the last triple with offs[tr] <= it (empty triples repeat an offset: skipped by "<=")
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (P.offs[mid] <= it) lo = mid; else hi = mid;
        }
        int is, b0;
        unsigned r0, r1;
        pg_triple_range(P, lo, is, b0, r0, r1);
        const unsigned local = it - P.offs[lo], seg = local / (unsigned)P.otiles;
        PgItem q;
        q.is = is; q.b0 = b0; q.otile = (int)(local - seg * (unsigned)P.otiles);
        q.rec0 = r0 + seg * PG_SEG;
        q.rec1 = min(r1, q.rec0 + (unsigned)PG_SEG);
        s_item = q;
      }
    }
    __syncthreads();
    if (!s_more) break;
    const PgItem item = s_item;
    const int b0 = item.b0, is = item.is, obase = item.otile * BN;
    const int nch = (int)((item.rec1 - item.rec0 + PG_KC - 1) / PG_KC);

    auto load_rec = [&](int c, RecRegs &R) {  // warp 0: record `lane` of chunk c
      const unsigned idx = item.rec0 + (unsigned)c * PG_KC + lane;
      R.ok = c < nch && idx < item.rec1;
      if (!R.ok) return;
      if (LINEAR) {
        const LRec r = reinterpret_cast<const LRec *>(A.recs)[idx];
        R.e0 = r.e0; R.p1 = r.e1; R.p2 = r.e2; R.p3 = r.e3; R.p4 = r.e4; R.om = r.om; R.klo = r.klo; R.wlen = r.wlen;
        R.unit = (unsigned)__double2loint(r.pad);
      } else {  // packed tail (see GRec)
        const GRec r = reinterpret_cast<const GRec *>(A.recs)[idx];
        R.e0 = r.e; R.p1 = r.w; R.om = r.om; R.klo = r.klo & 0x7fffff; R.wlen = (int)((unsigned)r.klo >> 23);
        R.unit = (unsigned)r.cpos | ((unsigned)r.wlen << 16);
      }
    };
    auto store_params = [&](int buf, const RecRegs &R) {  // warp 0
      double *P = sP + ((size_t)buf * PG_KC + lane) * 12;
      int klo = 0, wlen = 0;
      long long row = -1;
      if (R.ok) {
        klo = R.klo - is * A.bpad; wlen = R.wlen;
        if (LINEAR) {
          const double aa = R.p3 - R.p4, bb = R.p2 - R.p4, cc = R.p1 - R.p4, dd = R.e0 - R.p4, cb = cc - bb;
          const double full = (cc - aa) * cc + (aa * aa - cb * cb) * (1.0 / 3.0);
          P[0] = R.e0; P[1] = dd; P[2] = aa; P[3] = bb; P[4] = cc;
          P[5] = 0.5 * (cc + bb - aa);                                  // hh
          P[6] = aa > 1e-280 ? 0.5 * fast_rcp(aa) : 0.0;                // inv2a
          P[7] = cb > 1e-280 ? 0.5 * fast_rcp(cb) : 0.0;                // inv2cb
          P[8] = R.om * fast_rcp(full + (cc + bb - aa) * (dd - cc));    // alpha * omega
        } else {
          const double invw = fast_rcp(R.p1), del = delta * invw;
          P[0] = R.e0; P[1] = invw; P[2] = R.om * invw * OD_INV_SQRT_TWO_PI; P[3] = del; P[4] = exp(-del * del);
        }
        const unsigned ikl = R.unit / (unsigned)A.nb, ib = R.unit - ikl * (unsigned)A.nb;
        const long long ik = A.k_first + ikl;
        if (ik < A.mw_nk && (int)ib < A.mw_nb) row = (long long)A.norb * (ib + (long long)A.mw_nb * (ik + (long long)A.mw_nk * is));
      }
      sI[(buf * PG_KC + lane) * 2] = klo; sI[(buf * PG_KC + lane) * 2 + 1] = wlen;
      sRow[buf * PG_KC + lane] = row;
    };
    // weight rows of a chunk, global -> registers -> shared.  With an even number of orbitals (rows then start on 16-byte
    // boundaries) a thread moves pairs: half the row look-ups, bounds tests, loads and stores (they were 30 % of the kernel's
    // instructions element by element).
    const bool wpairs = BN == 64 && !LINEAR && (A.norb & 1) == 0 && (reinterpret_cast<size_t>(A.mw) & 15) == 0;  // (no gain at 16 / 32 orbitals per tile; spills in the linear variant)
    auto load_w = [&](int buf, double (&wr)[EPT]) {  // rows of the chunk whose parameters sit in `buf`
      if (wpairs) {
#pragma unroll
        for (int i = 0; i < EPT / 2; ++i) {
          const int e = i * 256 + t, k = e / (BN / 2), oo = 2 * (e % (BN / 2));
          const long long row = sRow[buf * PG_KC + k];
          const int o = obase + oo;
          double2 v = make_double2(0.0, 0.0);
          if (row >= 0 && o < A.norb) v = *reinterpret_cast<const double2 *>(A.mw + row + o);  // (o + 1 < norb: both even)
          wr[2 * i] = v.x; wr[2 * i + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int e = i * 256 + t, k = e / BN, oo = e % BN;
          const long long row = sRow[buf * PG_KC + k];
          const int o = obase + oo;
          wr[i] = (row >= 0 && o < A.norb) ? A.mw[row + o] : 0.0;
        }
      }
    };
    auto store_w = [&](int buf, const double (&wr)[EPT]) {
      if (wpairs) {
#pragma unroll
        for (int i = 0; i < EPT / 2; ++i) {
          const int e = i * 256 + t, k = e / (BN / 2), oo = 2 * (e % (BN / 2));
          *reinterpret_cast<double2 *>(sW + ((size_t)buf * PG_KC + k) * WS + oo) = make_double2(wr[2 * i], wr[2 * i + 1]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int e = i * 256 + t, k = e / BN, oo = e % BN;
          sW[((size_t)buf * PG_KC + k) * WS + oo] = wr[i];
        }
      }
    };
    auto fill_g = [&](int buf) {  // thread -> record k = lane, run of 8 bins = warp
      const int k = lane, run = warp;
      const int klo = sI[(buf * PG_KC + k) * 2], wlen = sI[(buf * PG_KC + k) * 2 + 1];
      const double *P = sP + ((size_t)buf * PG_KC + k) * 12;
      const int j0 = b0 + run * 8;
      const int lo = max(j0, klo), hi = min(j0 + 7, klo + wlen - 1);
      double v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.0;
      if (lo <= hi) {
        if (!LINEAR) {
          const double e0 = P[0], invw = P[1], amp = P[2], del = P[3], q = P[4];
          const double z = (emin + (double)lo * delta - e0) * invw;
          double g = amp * exp(-0.5 * z * z), rr = exp(-fma(z, del, 0.5 * del * del));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = j0 + i;
            if (j >= lo && j <= hi) { v[i] = g; g *= rr; rr *= q; }
          }
        } else {
          const double e0 = P[0], dd = P[1], aa = P[2], bb = P[3], cc = P[4], hh = P[5], inv2a = P[6], inv2cb = P[7], sc = P[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = j0 + i;
            if (j >= lo && j <= hi) {
              const double x = fmax(dd - fabs(emin + (double)j * delta - e0), 0.0);
              const double pq = fmax(aa - x, 0.0), y = fmax(cc - x, 0.0);
              const double D = x < bb ? fma(pq * pq, inv2a, x - 0.5 * aa) : fma(-y * y, inv2cb, hh);
              v[i] = D * sc;
            }
          }
        }
      }
      double *dst = sG + ((size_t)buf * PG_BM + run * 8) * PG_KS + k;
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i * PG_KS] = v[i];
      // which records reach this run of 8 bins: the MMA loop skips the (16 bins x 8 records) blocks that are all zero --
      // the records of a chunk are neighbours in (class, window start), so the zeros of a tile are whole bands of it
      const unsigned m = __ballot_sync(0xffffffffu, lo <= hi);
      if (lane == 0) sMask[buf * (PG_BM / 8) + run] = m;
    };

    double acc[NB][4];
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    // ---- prologue: parameters of chunks 0 and 1, G and W of chunk 0
    RecRegs R;
    R.ok = false;
    double wr[EPT];
    if (warp == 8) { load_rec(0, R); store_params(0, R); load_rec(1, R); }
    __syncthreads();
    if (warp < 8) {
      load_w(0, wr);
      fill_g(0);
      store_w(0, wr);
    } else {
      store_params(1, R);
    }
    __syncthreads();

    for (int c = 0; c < nch; ++c) {
      const int cur = c & 1, nxt = cur ^ 1;
      const bool more = c + 1 < nch;
      if (warp == 8) {  // cur's parameters were last read before the previous barrier
        load_rec(c + 2, R);
        if (c + 2 < nch) store_params(cur, R);
        __syncthreads();
        continue;
      }
      if (more) load_w(nxt, wr);
      if (more) fill_g(nxt);
      const double *gp = sG + ((size_t)cur * PG_BM + wm * 16 + fg) * PG_KS + ft;
      const double *wp = sW + ((size_t)cur * PG_KC + ft) * WS + wn * 8 * NB + fg;
      const unsigned blk = sMask[cur * (PG_BM / 8) + 2 * wm] | sMask[cur * (PG_BM / 8) + 2 * wm + 1];  // my 16 bins
#pragma unroll
      for (int kk = 0; kk < PG_KC; kk += 8) {
        if (((blk >> kk) & 0xffu) == 0) continue;  // (warp-uniform)
        const double a0 = gp[kk], a1 = gp[8 * PG_KS + kk], a2 = gp[kk + 4], a3 = gp[8 * PG_KS + kk + 4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const double b0v = wp[(size_t)kk * WS + nb * 8], b1v = wp[(size_t)(kk + 4) * WS + nb * 8];
          dmma_16x8x8(acc[nb], a0, a1, a2, a3, b0v, b1v);
        }
      }
      if (more) store_w(nxt, wr);
      __syncthreads();
    }
    if (warp < 8)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bin = b0 + wm * 16 + fg + (j >> 1) * 8;
        const int o = obase + wn * 8 * NB + nb * 8 + ft * 2 + (j & 1);
        if (bin < A.grid.nbins && o < A.norb && acc[nb][j] != 0.0)
          atomicAdd(&A.wdos[(size_t)bin + (size_t)A.grid.nbins * (is + (size_t)A.ns * o)], acc[nb][j]);
      }
  }
}

// dense-engine producer over the compacted list of wide units
template <class Inner>
struct WideListProducer {
  long long units_per_z;  // max over spins of the wide count
  Inner inner;
  const unsigned *list;
  long long stride;
  const unsigned long long *count;  // device: wide units of spin 1 / 2
  __device__ __forceinline__ void operator()(long long u, int z, RecOut &r) const {
    r.mode = -1;
    const int is = z % inner.ns;  // z may also carry an orbital group (dense weighted DOS)
    if ((unsigned long long)u >= count[is]) return;
    inner((long long)list[(size_t)is * stride + u], z, r);
  }
};

// out(j, is): dos += narrow dos (both families of groups); intdos += the local part of the rational-erfc groups + the
// inclusive prefix sum of their step histogram + the Euler-Maclaurin groups' dE (prefix(dos_em)_j - dos_em_j / 2 - fir_j)
__global__ void __launch_bounds__(1024) combine_kernel(const double *__restrict__ n_dos, const double *__restrict__ n_int,
                                                       const double *__restrict__ stephist, const double *__restrict__ n_dos_em,
                                                       double delta, int nbins, int bpad,
                                                       double *__restrict__ dos, double *__restrict__ intdos) {
  // grid (blocks of 1024 bins, spin channels).  A CTA first sums the prefix terms of all bins below its block (independent,
  // coalesced loads: at most 20 per thread on a 20 000-bin grid), then scans its own 1024 bins with shuffles.
  __shared__ double warp_tot[32], below_w[32];
  __shared__ double below_s;
  const int is = blockIdx.y, b0 = blockIdx.x * 1024, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *dem_p = n_dos_em + (size_t)is * bpad, *step_p = stephist + (size_t)is * bpad;
  double below = 0.0;
  for (int i = threadIdx.x; i < b0; i += 1024) below += fma(delta, dem_p[i], step_p[i]);
  for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
  const int j = b0 + threadIdx.x;
  const double dem = (j < nbins) ? dem_p[j] : 0.0;
  double v = (j < nbins) ? fma(delta, dem, step_p[j]) : 0.0;
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_tot[warp] = v;
  if (lane == 0) below_w[warp] = below;
  __syncthreads();
  if (warp == 0) {
    double w = warp_tot[lane], bsum = below_w[lane];
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_tot[lane] = w;  // inclusive over the warps of this block
    if (lane == 0) below_s = bsum;
  }
  __syncthreads();
  if (j < nbins) {
    const double before = below_s + (warp > 0 ? warp_tot[warp - 1] : 0.0);
    dos[(size_t)is * nbins + j] += n_dos[(size_t)is * bpad + j] + dem;
    // end correction of the Euler-Maclaurin groups (their Gaussians lie wholly inside the grid: zero beyond its ends)
    double fir = 0.0;
#pragma unroll 8
    for (int n = 1; n <= EM_NTAPS; ++n) {
      const double up = (j + n < nbins) ? dem_p[j + n] : 0.0, dn = (j - n >= 0) ? dem_p[j - n] : 0.0;
      fir = fma(c_em_fir[n - 1], up - dn, fir);
    }
    intdos[(size_t)is * nbins + j] += n_int[(size_t)is * bpad + j] - delta * (0.5 * dem + fir) + (before + v);
  }
}

// JD mode: out[dest(c) + is*nbins + j] += channel array c
struct ChanDest {  // destination offsets of up to 8 channel arrays, passed by value (no staging copy, no synchronisation)
  long long d[8];
};
__global__ void combine_channels_kernel(const double *__restrict__ chan, size_t chan_stride, int nchan, int nbins, int bpad, int ns,
                                        ChanDest dest, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y / ns, is = blockIdx.y % ns;
  if (j >= nbins || c >= nchan || dest.d[c] < 0) return;
  out[dest.d[c] + (size_t)is * nbins + j] += chan[(size_t)c * chan_stride + (size_t)is * bpad + j];
}

struct Scratch {
  PrepSpec ps;
  size_t nout;
  unsigned *keys, *table, *colsum, *base;
  unsigned long long *wide;  // [0..1] wide counts, [2..3] wide cursors, [4] next_chunk
  double *stephist, *chan;   // chan: 2 * npass arrays of nout doubles (DOS: dos, int)
};

inline void strip_weights(DosProducer &p) { p.mw = nullptr; }
inline void strip_weights(JdosProducer &p) { p.mw = nullptr; p.ome = nullptr; }
inline void set_eager(DosProducer &) {}
inline void set_eager(JdosProducer &p) { p.eager = 1; }
inline void set_divs(DosProducer &) {}
inline void set_divs(JdosProducer &p) { p.nocc_div = FastDiv((unsigned)(p.occ_list ? p.n_occ : p.nb)); }

// Both prepare passes, without a host synchronisation: the number of records (S.base[nbuckets]) and of wide units
// (S.wide[0..1]) stay on the device, the record buffers are sized for the worst case (every unit narrow / every unit wide).
template <class Producer>
int narrow_prepare(od_ctx *ctx, const Producer &prod_in, Scratch &S, bool linear, long long *wide_stride) {
  Producer prod = prod_in;
  prod.b.fast_corners = 1;  // both passes; wide units are re-made by the dense engine with the reference's arithmetic
  set_divs(prod);
  Producer prod_keys = prod;  // pass A only needs the windows: no weight columns, no optical matrix elements
  strip_weights(prod_keys);
  set_eager(prod);            // pass B only runs on slots pass A accepted: loads first, tests after
  const PrepSpec &ps = S.ps;
  // one prepare CTA per SM, fewer for small jobs (every CTA zeroes, dumps and re-reads a row of the bucket table); pass A and
  // pass B MUST use the same grid
  PrepMap pm;
  pm.W = (unsigned)prod.inner();
  pm.nk = (unsigned)(prod.units_per_z / prod.inner());
  pm.ngroups = (pm.W + 31) / 32;
  pm.kmajor = Producer::PREP_KMAJOR ? 1 : 0;
  pm.nk_div = FastDiv(pm.kmajor ? pm.ngroups : pm.nk);
  pm.nrows = (unsigned long long)pm.ngroups * pm.nk;
  const long long per_cta = (long long)NR_PREP_ROWS * Producer::PREP_U;  // at least one tile of rows per CTA
  const int prows = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count, ((long long)pm.nrows + per_cta - 1) / per_cta));
  pm.per = (pm.nrows + prows - 1) / prows;
  const size_t hist_smem = sizeof(unsigned) * ps.nbuckets;
  const size_t units = (size_t)prod.units_per_z * prod.ns;
  *wide_stride = prod.units_per_z;
  const size_t rec_bytes = linear ? sizeof(LRec) : sizeof(GRec);
  OD_CHECK(od_buf_ensure(ctx, ctx->recs, rec_bytes * units + 256));
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_b, sizeof(unsigned) * units + 256));
  if (ps.ng > 0) OD_CHECK(od_buf_ensure(ctx, ctx->scratch_d, sizeof(double) * units * ps.wstride + 256));
  OD_CUDA(ctx, cudaMemsetAsync(S.wide, 0, sizeof(unsigned long long) * 5, ctx->stream));
  double *wts = ps.ng > 0 ? ctx->scratch_d.as<double>() : nullptr;
  OD_CUDA(ctx, cudaFuncSetAttribute(prep_hist_kernel<Producer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
  prep_hist_kernel<Producer><<<prows, NR_PREP_THREADS, hist_smem, ctx->stream>>>(prod_keys, ps, S.keys, S.table, S.wide, pm);
  OD_LAUNCH_CHECK(ctx);
  prep_colscan_kernel<<<(4 * ps.nbuckets + 255) / 256, 256, 0, ctx->stream>>>(S.table, prows, ps.nbuckets, S.colsum);
  OD_LAUNCH_CHECK(ctx);
  scan_kernel<<<1, 1024, 0, ctx->stream>>>(S.colsum, ps.nbuckets, S.base);
  OD_LAUNCH_CHECK(ctx);
#define SC_LAUNCH(LIN)                                                                                                               \
  do {                                                                                                                               \
    OD_CUDA(ctx, cudaFuncSetAttribute(prep_scatter_kernel<LIN, Producer>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem)); \
    prep_scatter_kernel<LIN, Producer><<<prows, NR_PREP_THREADS, hist_smem, ctx->stream>>>(                                         \
        prod, ps, S.keys, S.table, S.base, ctx->recs.p, wts, S.stephist, ctx->scratch_b.as<unsigned>(), S.wide + 2, *wide_stride, pm); \
  } while (0)
  if (linear) SC_LAUNCH(true); else SC_LAUNCH(false);
#undef SC_LAUNCH
  OD_LAUNCH_CHECK(ctx);
  return OD_OK;
}

struct MwGather {
  const double *mw = nullptr;
  int norb = 0, mw_nb = 0, mw_nk = 0, nb = 0;
  long long k_first = 0;
};

template <int JD>
int narrow_launch(od_ctx *ctx, const Scratch &S, bool linear, int col0, int col1, double *out0, double *out1,
                  const MwGather *gather = nullptr, int col2 = -2, int col3 = -2, double *out2 = nullptr, double *out3 = nullptr) {
  NarrowArgs A;
  A.recs = ctx->recs.p;
  A.nrec_dev = S.base + S.ps.nbuckets;
  A.grid = S.ps.grid;
  A.bpad = S.ps.bpad;
  A.cell = S.ps.cell;
  A.next_chunk = S.wide + 4;
  A.stats = ctx->stats_dev.as<unsigned long long>();
  for (int i = 0; i < 6; ++i) { A.cf.P[i] = h_P[i]; A.cf.Q[i] = h_Q[i]; }
  A.out_dos = out0;
  A.out_int = out1;
  A.out_step = S.stephist;
  A.out_dos_em = JD ? nullptr : S.chan + 2 * S.nout;  // DOS: chan = [dos | int | dos_em]
  A.no_em = getenv("OD_NO_EM") ? 1 : 0;
  A.debug = getenv("OD_NARROW_DEBUG") ? atoi(getenv("OD_NARROW_DEBUG")) : 0;
  A.flags = ctx->flags.as<int>();
  A.wts = S.ps.ng > 0 ? ctx->scratch_d.as<double>() : nullptr;
  A.ng = S.ps.ng;
  A.wstride = S.ps.wstride;
  A.col0 = col0;
  A.col1 = col1;
  A.col2 = col2; A.col3 = col3; A.out2 = out2; A.out3 = out3;
  A.mw = nullptr; A.norb = A.mw_nb = A.mw_nk = A.nb = A.mw_ns = 0; A.k_first = 0;
  A.unit_in_rec = S.ps.want_unit;
  if (gather) {
    A.mw = gather->mw; A.norb = gather->norb; A.mw_nb = gather->mw_nb; A.mw_nk = gather->mw_nk; A.nb = gather->nb;
    A.k_first = gather->k_first;
  }
  const size_t smem = (sizeof(double2) + sizeof(double)) * NR_T * NR_WARPS + (size_t)NR_SORT_BYTES * NR_WARPS;
  OD_CUDA(ctx, cudaMemsetAsync(S.wide + 4, 0, sizeof(unsigned long long), ctx->stream));
  const int blocks = ctx->sm_count * NR_CTAS_PER_SM;  // persistent; the record count is read on the device
  if (linear) {
    OD_CUDA(ctx, cudaFuncSetAttribute(narrow_kernel<true, JD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    narrow_kernel<true, JD><<<blocks, NR_WARPS * 32, smem, ctx->stream>>>(A);
  } else {
    OD_CUDA(ctx, cudaFuncSetAttribute(narrow_kernel<false, JD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    narrow_kernel<false, JD><<<blocks, NR_WARPS * 32, smem, ctx->stream>>>(A);
  }
  OD_LAUNCH_CHECK(ctx);
  return OD_OK;
}

// bins per position cell: the smallest power of two >= 8 whose bucket table fits the prepare passes' shared memory -- and,
// for small jobs, large enough that a bucket holds of the order of a chunk of records (a chunk of 256 consecutive records
// that spans a dozen sparse buckets walks the union of their windows: a 1M-unit job on a 20 000-bin grid ran 6x slower per
// record than a 500M-unit one); 0 = the grid is too large for the narrow engine (the callers then use the dense engine).
// total_units < 0: the grid test alone.
int narrow_cell_for(int nbins, int ns, long long total_units = -1) {
  for (int cell = 8; cell <= 128; cell *= 2) {
    const int bpad = ((nbins + NR_T + cell - 1) / cell) * cell;
    const long long nbuckets = (long long)NR_NCLS * ns * (bpad / cell);
    if (nbuckets > NR_MAXBUCKETS) continue;
    // (units spread over about a third of the buckets: the band structure does not cover the padded grid, and a width
    // class holds a part of the units)
    if (total_units >= 0 && cell < 64 && total_units < 64 * nbuckets) continue;
    return cell;
  }
  return 0;
}

// carve the scratch buffers for a grid / unit count / number of channel arrays
int narrow_scratch(od_ctx *ctx, const GridSpec &grid, int ns, long long total_units, int nchan_arrays, Scratch &S) {
  PrepSpec &ps = S.ps;
  ps.grid = grid;
  ps.inv_delta = 1.0 / grid.delta;
  ps.cell = narrow_cell_for(grid.nbins, ns, total_units);
  if (ps.cell == 0) return od_fail(ctx, OD_ERR_ARG, "narrow engine: energy grid too large (%d bins)", grid.nbins);
  ps.wmax = std::min(NR_WMAX, NR_T - 2 * ps.cell);
  ps.cshift = 0;
  while ((1 << ps.cshift) < ps.cell) ++ps.cshift;
  ps.bpad = ((grid.nbins + NR_T + ps.cell - 1) / ps.cell) * ps.cell;
  ps.ncells = ns * ps.bpad / ps.cell;
  ps.nbuckets = NR_NCLS * ps.ncells;
  S.nout = (size_t)ns * ps.bpad + NR_T;
  OD_CHECK(od_flags_ready(ctx));
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_csum = take(sizeof(unsigned) * ps.nbuckets);
  const size_t o_base = take(sizeof(unsigned) * (ps.nbuckets + 1));
  const size_t o_wide = take(sizeof(unsigned long long) * 8);
  const size_t o_step = take(sizeof(double) * S.nout);
  const size_t o_chan = take(sizeof(double) * S.nout * nchan_arrays);
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_a, off));
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_c, sizeof(unsigned) * (size_t)ctx->sm_count * ps.nbuckets));
  OD_CHECK(od_buf_ensure(ctx, ctx->keys, sizeof(unsigned) * (size_t)total_units + 1024));
  char *sa = ctx->scratch_a.as<char>();
  OD_CUDA(ctx, cudaMemsetAsync(sa, 0, off, ctx->stream));
  S.keys = ctx->keys.as<unsigned>();
  S.table = ctx->scratch_c.as<unsigned>();
  S.colsum = reinterpret_cast<unsigned *>(sa + o_csum);
  S.base = reinterpret_cast<unsigned *>(sa + o_base);
  S.wide = reinterpret_cast<unsigned long long *>(sa + o_wide);
  S.stephist = reinterpret_cast<double *>(sa + o_step);
  S.chan = reinterpret_cast<double *>(sa + o_chan);
  return OD_OK;
}

}  // namespace

bool od_narrow_grid_fits(int nbins, int ns) { return narrow_cell_for(nbins, ns) != 0; }

// Narrow pipeline for calculate_dos.  accum = [dos(B,ns) | intdos(B,ns) | weighted_dos(B,ns,norb)], zeroed by the
// caller.  With matrix_weights (prod_in.mw != null) the sorted records are walked once more per pair of weight
// columns (the Gaussian recurrence / doslin evaluation is cheap; the weights are gathered by unit index).
// No host synchronisation on the path without weights: record and wide-unit counts stay on the device.
int od_narrow_dos(od_ctx *ctx, const DosProducer &prod_in, const GridSpec &grid, bool linear, double *accum) {
  DosProducer prod = prod_in;
  prod.mw = nullptr;  // the prepare passes and the DOS pass do not need the weights
  const int norb = prod_in.mw ? prod_in.norb : 0;
  const int ns = prod.ns, nbins = grid.nbins;
  const long long total = prod.units_per_z * ns;
  if (total >= (1ll << 32) || prod.units_per_z >= (1ll << 31))
    return od_fail(ctx, OD_ERR_ARG, "narrow engine: more than 2^31 units per spin channel in one call");
  Scratch S;
  OD_CHECK(narrow_scratch(ctx, grid, ns, total, 3, S));  // [dos | local intdos | dos of the Euler-Maclaurin groups]
  S.ps.gwin = NR_ERFWIN;
  S.ps.with_int = 1;
  S.ps.linear_pass = 0;  // calculate_dos ignores hybrid_linear (quirk Q8): a linear pass has no Gaussian records
  S.ps.ng = 0;
  S.ps.wstride = 0;
  S.ps.want_unit = norb > 0;
  double *ndos = S.chan, *nint = S.chan + S.nout;

  long long wide_stride = 0;
  size_t tp = 0, ta = 0;
  OD_CHECK(od_time_begin(ctx, 0, &tp));
  OD_CHECK(narrow_prepare(ctx, prod, S, linear, &wide_stride));
  OD_CHECK(od_time_end(ctx, tp));
  OD_CHECK(od_time_begin(ctx, 1, &ta));
  OD_CHECK(narrow_launch<0>(ctx, S, linear, -2, -2, ndos, nint, nullptr));
  const size_t n1 = (size_t)nbins * ns;
  combine_kernel<<<dim3((nbins + 1023) / 1024, ns), 1024, 0, ctx->stream>>>(ndos, nint, S.stephist, S.chan + 2 * S.nout, grid.delta, nbins, S.ps.bpad,
                                                accum, accum + n1);
  OD_LAUNCH_CHECK(ctx);
  int wpasses = 0;
  if (norb >= 8) {
    // many projectors: banded GEMM over (bin tile, orbital tile, record segment) work items, enumerated on the device
    const PrepSpec &ps = S.ps;
    static const int wmax[NR_NCLS] = {24, 32, 44, 60, 84, 116, 160, 224, 320, NR_WMAX};
    const int BN = norb > 32 ? 64 : (norb > 16 ? 32 : 16);
    PgArgs A;
    PgPlan &P = A.plan;
    P.ntiles = (nbins + PG_BM - 1) / PG_BM;
    P.ntriples = ns * P.ntiles * NR_NCLS;
    P.otiles = (norb + BN - 1) / BN;
    P.nbins = nbins; P.bpad = ps.bpad; P.cell = ps.cell; P.ncells = ps.ncells;
    for (int c = 0; c < NR_NCLS; ++c) P.wmax[c] = wmax[c];
    OD_CHECK(od_buf_ensure(ctx, ctx->scratch_d, sizeof(unsigned) * (2 * (size_t)P.ntriples + 2) + 64));
    unsigned *counts = ctx->scratch_d.as<unsigned>(), *offs = counts + P.ntriples;
    P.base = S.base; P.offs = offs;
    pg_count_kernel<<<(P.ntriples + 255) / 256, 256, 0, ctx->stream>>>(P, counts);
    OD_LAUNCH_CHECK(ctx);
    scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, P.ntriples, offs);
    OD_LAUNCH_CHECK(ctx);
    OD_CUDA(ctx, cudaMemsetAsync(S.wide + 4, 0, sizeof(unsigned long long), ctx->stream));
    A.recs = ctx->recs.p;
    A.next_item = S.wide + 4; A.grid = grid; A.bpad = ps.bpad; A.ns = ns;
    A.mw = prod_in.mw; A.norb = norb; A.mw_nb = prod_in.mw_nb; A.mw_nk = prod_in.mw_nk; A.nb = prod_in.nb; A.k_first = prod_in.k_first;
    A.wdos = accum + 2 * n1;
    const int blocks = ctx->sm_count * 2;  // persistent; the number of items is read on the device
#define PG_LAUNCH(BNV)                                                                                                   \
  do {                                                                                                                   \
    const int smem = (int)pg_smem_bytes<BNV>();                                                                          \
    if (linear) {                                                                                                        \
      OD_CUDA(ctx, cudaFuncSetAttribute(pdos_gemm_kernel<BNV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      pdos_gemm_kernel<BNV, true><<<blocks, PG_THREADS, smem, ctx->stream>>>(A);                                          \
    } else {                                                                                                             \
      OD_CUDA(ctx, cudaFuncSetAttribute(pdos_gemm_kernel<BNV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      pdos_gemm_kernel<BNV, false><<<blocks, PG_THREADS, smem, ctx->stream>>>(A);                                         \
    }                                                                                                                    \
  } while (0)
    if (BN == 64) PG_LAUNCH(64); else if (BN == 32) PG_LAUNCH(32); else PG_LAUNCH(16);
#undef PG_LAUNCH
    OD_LAUNCH_CHECK(ctx);
    wpasses = 1;
  } else if (norb > 0) {
    // weighted spectrum: columns (2c, 2c+1) per pass, K * matrix_weights(o, ib, ik, is) (dos_utils.f90:1271)
    MwGather gth;
    gth.mw = prod_in.mw; gth.norb = norb; gth.mw_nb = prod_in.mw_nb; gth.mw_nk = prod_in.mw_nk; gth.nb = prod_in.nb;
    gth.k_first = prod_in.k_first;
    for (int c = 0; 2 * c < norb; ++c, ++wpasses) {
      OD_CUDA(ctx, cudaMemsetAsync(S.chan, 0, sizeof(double) * 2 * S.nout, ctx->stream));
      OD_CHECK(narrow_launch<1>(ctx, S, linear, 2 * c, 2 * c + 1 < norb ? 2 * c + 1 : -2, ndos, nint, &gth));
      ChanDest dest;
      for (long long &d : dest.d) d = -1;
      dest.d[0] = (long long)(2 * n1 + n1 * (size_t)(2 * c));
      if (2 * c + 1 < norb) dest.d[1] = (long long)(2 * n1 + n1 * (size_t)(2 * c + 1));
      dim3 g((nbins + 255) / 256, 2 * ns);
      combine_channels_kernel<<<g, 256, 0, ctx->stream>>>(S.chan, S.nout, 2, nbins, S.ps.bpad, ns, dest, accum);
      OD_LAUNCH_CHECK(ctx);
    }
  }
  OD_CHECK(od_time_end(ctx, ta));

  // ---- the wide units through the dense engine (adds into accum); their count is only known on the device
  {
    WideListProducer<DosProducer> wp;
    wp.units_per_z = wide_stride;
    wp.inner = prod_in;
    wp.list = ctx->scratch_b.as<unsigned>();
    wp.stride = wide_stride;
    wp.count = S.wide;
    if (norb == 0) {
      std::vector<long long> dest(ns * 2);
      for (int is = 0; is < ns; ++is) {
        dest[is * 2 + 0] = (long long)is * nbins;
        dest[is * 2 + 1] = (long long)(n1 + (size_t)is * nbins);
      }
      OD_CHECK((od_run_dense_dev<0, true, WideListProducer<DosProducer>>(ctx, wp, grid, ns, dest, accum, S.wide)));
    } else {
      const int ngroups = (norb + OD_MAXW - 1) / OD_MAXW, nz = ns * ngroups;
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
      OD_CHECK((od_run_dense_dev<OD_MAXW, true, WideListProducer<DosProducer>>(ctx, wp, grid, nz, dest, accum, S.wide)));
    }
  }
  ctx->prof.accumulate_launches += 1 + wpasses;
  if (ctx->pending_times.size() > 4096) OD_CHECK(od_times_resolve(ctx));
  return OD_OK;
}

// Narrow pipeline for calculate_jdos (optionally with n_geom weight columns, materialised or formed on chip
// from the optical matrix elements by the producer).  accum = [jdos(B,ns) | weighted_jdos(B,ns,n_geom)], zeroed
// by the caller.  Pair slots are processed in k-point slabs of < 2^30 slots.
int od_narrow_jdos(od_ctx *ctx, const JdosProducer &prod_in, const GridSpec &grid, bool linear, int n_geom, double *accum) {
  const int ns = prod_in.ns, nbins = grid.nbins;
  const long long ppk = prod_in.pairs_per_k();  // pair slots per k-point
  const long long nk = prod_in.units_per_z / ppk;
  // pair slots per k-slab: small enough that the bucket-order scatter of one slab stays L2-friendly (see the DOS
  // slabs in od_spectral.cu), and < 2^30 for the 32-bit unit arithmetic; OD_JDOS_SLAB_SLOTS overrides for experiments
  const long long slab_slots = std::min<long long>(1ll << 30, getenv("OD_JDOS_SLAB_SLOTS") ? atoll(getenv("OD_JDOS_SLAB_SLOTS")) : (16ll << 20));
  const long long slab_k = std::max<long long>(1, slab_slots / ppk);
  // channels: 0 = JDOS, a >= 1 = geometry component a.  Two per pass (the tile in two copies: short walks), or -- from three
  // channels on -- four per pass on a tile of half the bins: the tensor's 7 channels take 2 passes over the records, not 4
  const bool quad = n_geom >= 2 && narrow_cell_for(grid.nbins, ns, std::min(nk, slab_k) * ppk * ns) <= 16;  // (room for windows of 96 bins)
  const int cpp = quad ? 4 : 2;                  // channels per pass
  const int npass = (1 + n_geom + cpp - 1) / cpp;
  Scratch S;
  OD_CHECK(narrow_scratch(ctx, grid, ns, std::min(nk, slab_k) * ppk * ns, cpp * npass, S));
  if (quad) S.ps.wmax = std::min(S.ps.wmax, NR_T / 4 - 2 * S.ps.cell);  // (wider windows: dense engine)
  S.ps.gwin = OD_SQRT60;   // the Gaussian cut-off (algorithms.f90:83); no erf window in JDOS
  S.ps.with_int = 0;
  S.ps.linear_pass = linear ? 1 : 0;  // hybrid_linear fall-back records (jdos_utils.f90:365) -> dense engine
  S.ps.ng = n_geom;
  S.ps.wstride = n_geom <= 0 ? 0 : (n_geom == 1 ? 1 : (n_geom <= 4 ? 4 : 8));
  S.ps.want_unit = 0;
  const size_t n1 = (size_t)nbins * ns;
  for (long long k0 = 0; k0 < nk; k0 += slab_k) {
    JdosProducer prod = prod_in;
    prod.k_first = prod_in.k_first + k0;
    prod.units_per_z = std::min(slab_k, nk - k0) * ppk;
    long long wide_stride = 0;
    size_t tp = 0, ta = 0;
    OD_CHECK(od_time_begin(ctx, 0, &tp));
    OD_CHECK(narrow_prepare(ctx, prod, S, linear, &wide_stride));
    OD_CHECK(od_time_end(ctx, tp));
    OD_CHECK(od_time_begin(ctx, 1, &ta));
    for (int c = 0; c < npass; ++c) {
      auto col = [&](int ch) { return ch == 0 ? -1 : (ch <= n_geom ? ch - 1 : -2); };  // weight column of channel ch
      double *o = S.chan + (size_t)(cpp * c) * S.nout;
      if (quad)
        OD_CHECK(narrow_launch<2>(ctx, S, linear, col(4 * c), col(4 * c + 1), o, o + S.nout, nullptr, col(4 * c + 2), col(4 * c + 3),
                                  o + 2 * S.nout, o + 3 * S.nout));
      else
        OD_CHECK(narrow_launch<1>(ctx, S, linear, col(2 * c), col(2 * c + 1), o, o + S.nout));
    }
    OD_CHECK(od_time_end(ctx, ta));
    {
      WideListProducer<JdosProducer> wp;
      wp.units_per_z = wide_stride;
      wp.inner = prod;
      wp.list = ctx->scratch_b.as<unsigned>();
      wp.stride = wide_stride;
      wp.count = S.wide;
      const int nacc = 1 + n_geom;
      std::vector<long long> dest((size_t)ns * nacc);
      for (int is = 0; is < ns; ++is) {
        dest[(size_t)is * nacc] = (long long)is * nbins;
        for (int a = 0; a < n_geom; ++a) dest[(size_t)is * nacc + 1 + a] = (long long)(n1 + (size_t)nbins * (is + (size_t)ns * a));
      }
      if (n_geom == 0) OD_CHECK((od_run_dense_dev<0, false, WideListProducer<JdosProducer>>(ctx, wp, grid, ns, dest, accum, S.wide)));
      else if (n_geom == 1) OD_CHECK((od_run_dense_dev<1, false, WideListProducer<JdosProducer>>(ctx, wp, grid, ns, dest, accum, S.wide)));
      else OD_CHECK((od_run_dense_dev<6, false, WideListProducer<JdosProducer>>(ctx, wp, grid, ns, dest, accum, S.wide)));
    }
    ctx->prof.accumulate_launches += npass;
    // the next slab's prepare resets S.wide and overwrites the lists: stream order keeps the dense pass above ahead of it
  }
  // channel arrays -> [jdos | weighted_jdos]: array 0 = jdos, array a (a >= 1) = geometry component a
  ChanDest dest;
  for (long long &d : dest.d) d = -1;
  dest.d[0] = 0;
  for (int a = 1; a <= n_geom; ++a) dest.d[a] = (long long)(n1 + (size_t)nbins * ns * (a - 1));
  dim3 g((nbins + 255) / 256, cpp * npass * ns);
  combine_channels_kernel<<<g, 256, 0, ctx->stream>>>(S.chan, S.nout, cpp * npass, nbins, S.ps.bpad, ns, dest, accum);
  OD_LAUNCH_CHECK(ctx);
  if (ctx->pending_times.size() > 4096) OD_CHECK(od_times_resolve(ctx));
  return OD_OK;
}

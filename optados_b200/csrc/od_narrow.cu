// od_narrow.cu -- the sorted "narrow-window" accumulation engine (the fast path of calculate_dos).
//
// Regime: many units (5e8 states at BASELINE config 2) whose broadening windows (15..250 bins) are
// short compared with the energy grid (20 000 bins).  Every unit then touches ~0.5 % of the grid,
// so the reference's dense loop does 200x more work than needed, and a scatter into the grid with
// atomics would serialise on hot bins.  This engine instead
//   1. PREPARES: one pass classifies every unit (window on the grid, step bin, narrow / wide) and
//      histograms the narrow ones by the first bin of their window; an exclusive scan turns the
//      histogram into bucket offsets; a second pass scatters 32-byte (Gaussian) or 64-byte (linear)
//      records into bucket order.  The semi-infinite part of the integrated spectrum (every state
//      adds its full weight to all bins above it) is split off as a STEP at the state's centre bin,
//      histogrammed here and prefix-summed at the end, so what remains to accumulate for intDOS,
//      Phi(z) - H(z >= 0), is as local as the DOS itself.
//   2. ACCUMULATES: warps pull chunks of 2048 consecutive records.  Within a group of 32 records
//      LANE = RECORD: each lane walks the bins of the group's common window, lane l starting l bins
//      in and wrapping round, so that in any one step the 32 lanes touch 32 DIFFERENT bins of a
//      warp-private shared-memory tile -- plain LDS/STS read-modify-write, no atomics, no
//      reductions.  Walking consecutive bins lets a lane advance its Gaussian by the recurrence
//      g(j+1) = g(j) r(j), r(j+1) = r(j) q  (q = exp(-dE^2/w^2)): two multiplies per bin instead of
//      an exp.  The cumulative part shares that Gaussian: Phi(z) - H = -+ (1/2) g' R(|z|) with
//      R = P5/Q6 ~ erfcx(|z|/sqrt 2) (tools/fit_erfcx.py, 5.4e-15 absolute), one reciprocal per bin.
//      The linear scheme evaluates the four doslin pieces branch-free and selects.
//      Tiles are flushed to HBM with FP64 reductions only when a warp's window slides off its
//      512-bin tile (about once per chunk).
//   3. Units whose window is wider than 256 bins go to the dense bin-owner engine (od_accum.cuh)
//      through a compacted list; the two partial spectra and the step prefix sum are combined.
//
// Numerics vs the reference (tolerances: 1e-9 of peak on spectra, 1e-10 relative on intDOS): the
// recurrence accumulates ~n^2/2 ulp over n <= 256 steps (< 4e-12 relative on a value that is itself
// <= the state's peak); R carries 5.4e-15; the exact cut-offs of the reference (gaussian == 0 for
// z^2/2 > 30, erf == +-1 for |x| > 5.8) are applied on bin indices, a difference of at most 1e-13 of
// one state's peak on the edge bin.  Summation order inside a bucket follows the scatter order
// (atomic cursors), so results are reproducible to rounding (~1e-16 relative), not bit for bit.
#include <algorithm>
#include <cmath>

#include "od_internal.h"
#include "od_spectral.h"

namespace {

constexpr int NR_T = 512;       // bins per warp tile
constexpr int NR_WMAX = 256;    // widest window (bins) the narrow engine takes
constexpr int NR_CHUNK = 2048;  // records per work item
constexpr int NR_WARPS = 8;     // warps per CTA
constexpr int NR_MAXBUCKETS = 49152;
constexpr unsigned KEY_WIDE = 0xFFFFFFFEu, KEY_DROP = 0xFFFFFFFFu;

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

struct PrepSpec {
  GridSpec grid;
  int bpad;     // padded bins per spin (>= nbins + NR_T)
  int cell;     // bins per sort bucket (power of two)
  int nbuckets; // ns * bpad / cell
};

struct Cls {
  int klo, wlen, cpos, cstep;  // cstep: absolute step bin in [0, nbins]
  int kind;                    // 0 narrow, 1 wide, 2 dropped
};

#define NR_ERFWIN (OD_ERF_CUT * OD_SQRT_TWO)  // 8.2024...: |z| beyond which the NSWC erf saturates

__device__ __forceinline__ Cls classify(const RecOut &r, const GridSpec &g) {
  Cls c;
  double lo, hi, centre;
  const double inv = 1.0 / g.delta;
  if (r.mode == 0) {
    lo = r.e0 - NR_ERFWIN * r.p1;
    hi = r.e0 + NR_ERFWIN * r.p1;
    centre = ceil((r.e0 - g.emin) * inv);  // first bin with E_j >= e
  } else {
    lo = r.p4;
    hi = 2.0 * r.e0 - r.p4;
    centre = floor((r.e0 - g.emin) * inv) + 1.0;  // first bin with E_j > e0
  }
  double a = floor((lo - g.emin) * inv), b = ceil((hi - g.emin) * inv);
  if (!(hi > lo)) b = a - 1.0;  // zero-width window (flat linear cell): only the step survives
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
  c.kind = (c.wlen <= NR_WMAX) ? 0 : 1;
  // Gaussians narrower than dE/8 (only possible with finite_bin_correction off) would overflow the
  // recurrence ratio: leave them to the exact dense engine
  if (r.mode == 0 && !(r.p1 * 8.0 >= g.delta)) c.kind = 1;
  int cp = c.cstep - c.klo;
  cp = cp < 0 ? 0 : (cp > c.wlen ? c.wlen : cp);
  c.cpos = cp;
  return c;
}

// ---------------------------------------------------------------- pass A: keys + histogram
__global__ void __launch_bounds__(256)
prep_hist_kernel(DosProducer prod, PrepSpec ps, unsigned *__restrict__ keys, unsigned *__restrict__ hist,
                 unsigned long long *__restrict__ wide_count /* [ns] */) {
  extern __shared__ unsigned sh_hist[];
  for (int i = threadIdx.x; i < ps.nbuckets; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const long long total = prod.units_per_z * prod.ns;
  unsigned long long wide_local[2] = {0, 0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int is = (int)(t / prod.units_per_z);
    const long long u = t - (long long)is * prod.units_per_z;
    RecOut r;
    prod(u, is, r);
    const Cls c = classify(r, ps.grid);
    unsigned key;
    if (c.kind == 0) {
      key = (unsigned)((is * ps.bpad + c.klo) / ps.cell);
      atomicAdd(&sh_hist[key], 1u);
    } else if (c.kind == 1) {
      key = KEY_WIDE;
      wide_local[is]++;
    } else {
      key = KEY_DROP;
    }
    keys[t] = key;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ps.nbuckets; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&hist[i], sh_hist[i]);
  for (int is = 0; is < prod.ns; ++is) {
    unsigned long long v = wide_local[is];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&wide_count[is], v);
  }
}

// exclusive scan of the bucket histogram (single CTA); cursor[b] = offsets[b]; total -> offsets[nbuckets]
__global__ void __launch_bounds__(1024) scan_kernel(const unsigned *__restrict__ hist, int n, unsigned long long *__restrict__ offsets,
                                                    unsigned long long *__restrict__ cursor) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = (i < n) ? hist[i] : 0ull;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      unsigned long long t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0ull;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const unsigned long long excl = carry + sh[threadIdx.x] - v;
    if (i < n) { offsets[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// ---------------------------------------------------------------- pass B: scatter
template <bool LINEAR>
__global__ void __launch_bounds__(256)
prep_scatter_kernel(DosProducer prod, PrepSpec ps, const unsigned *__restrict__ keys, unsigned long long *__restrict__ cursor,
                    void *__restrict__ recs_v, double *__restrict__ stephist, unsigned *__restrict__ widelist,
                    unsigned long long *__restrict__ wide_cursor, long long wide_stride) {
  const long long total = prod.units_per_z * prod.ns;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const unsigned key = keys[t];
    const int is = (int)(t / prod.units_per_z);
    const long long u = t - (long long)is * prod.units_per_z;
    if (key == KEY_WIDE) {
      const unsigned long long pos = atomicAdd(&wide_cursor[is], 1ull);
      widelist[(size_t)is * wide_stride + pos] = (unsigned)u;
      continue;
    }
    RecOut r;
    prod(u, is, r);
    const Cls c = classify(r, ps.grid);
    // the step of every unit the dense engine does not see
    if (c.cstep < ps.grid.nbins) atomicAdd(&stephist[(size_t)is * ps.bpad + c.cstep], r.omega);
    if (key == KEY_DROP) continue;
    const unsigned long long pos = atomicAdd(&cursor[key], 1ull);
    // keep the record inside the bucket pass A counted it in, whatever this pass recomputed
    int klo = c.klo;
    const int bucket_lo = (int)key * ps.cell - is * ps.bpad;
    klo = max(klo, bucket_lo);
    klo = min(klo, bucket_lo + ps.cell - 1);
    int wlen = min(max(c.wlen - (klo - c.klo), 1), NR_WMAX);
    int cpos = min(max(c.cstep - klo, 0), wlen);
    if (LINEAR) {
      LRec o;
      o.e0 = r.e0; o.e1 = r.p1; o.e2 = r.p2; o.e3 = r.p3; o.e4 = r.p4; o.om = r.omega;
      o.klo = is * ps.bpad + klo; o.cpos = (unsigned short)cpos; o.wlen = (unsigned short)wlen; o.pad = 0.0;
      reinterpret_cast<LRec *>(recs_v)[pos] = o;
    } else {
      GRec o;
      o.e = r.e0; o.w = r.p1; o.om = r.omega;
      o.klo = is * ps.bpad + klo; o.cpos = (unsigned short)cpos; o.wlen = (unsigned short)wlen;
      reinterpret_cast<GRec *>(recs_v)[pos] = o;
    }
  }
}

// ---------------------------------------------------------------- accumulate
__device__ __forceinline__ double rcp_newton(double q) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  const double e = fma(-q, y, 1.0);
  y = fma(y, e, y);            // ~2^-46
  return y;
}

// R(u) = P5(u)/Q6(u) ~ erfcx(u / sqrt 2), u in [0, 8.45]; all coefficients positive (tools/fit_erfcx.py)
__device__ __forceinline__ double erfcx_ratio(double u) {
  double p = 0.7978004363933496607;
  p = fma(p, u, 12.593079456648224345);
  p = fma(p, u, 89.133859365727141543);
  p = fma(p, u, 355.14175547719036648);
  p = fma(p, u, 802.96638201765452004);
  p = fma(p, u, 913.67854424431436856);
  double q = u + 15.779603335732765349;
  q = fma(q, u, 112.76898276075761223);
  q = fma(q, u, 460.29258012431025649);
  q = fma(q, u, 1120.6427893189240945);
  q = fma(q, u, 1531.976386005259472);
  q = fma(q, u, 913.67854424432351188);
  return p * rcp_newton(q);
}

__device__ __forceinline__ void flush_tile(double2 *tile, int base, int used, double *__restrict__ out_dos,
                                           double *__restrict__ out_int, int lane) {
  for (int i = lane; i < used; i += 32) {
    const double2 v = tile[i];
    if (v.x != 0.0) atomicAdd(&out_dos[base + i], v.x);
    if (v.y != 0.0) atomicAdd(&out_int[base + i], v.y);
    tile[i] = make_double2(0.0, 0.0);
  }
  __syncwarp();
}

struct NarrowArgs {
  const void *recs;
  long long nrec;
  GridSpec grid;
  int bpad, cell;
  unsigned long long *next_chunk;
  double *out_dos, *out_int;  // [ns * bpad + NR_T]
};

// direct (non-recurrence) evaluation of one Gaussian record at one bin: used by the rare oversize groups
__device__ __forceinline__ void gauss_direct(double e, double w, double om, double Ej, int j, int cabs, double &d, double &s) {
  const double z = (Ej - e) / w;
  const double gp = exp(-0.5 * z * z);
  d = (0.5 * z * z > 30.0) ? 0.0 : OD_INV_SQRT_TWO_PI * gp / w * om;
  const double h = 0.5 * om * gp * erfcx_ratio(fabs(z));
  s = (j < cabs) ? h : -h;
}

template <bool LINEAR>
__global__ void __launch_bounds__(NR_WARPS * 32, 2) narrow_kernel(NarrowArgs A) {
  extern __shared__ double2 sh_tiles[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2 *tile = sh_tiles + (size_t)warp * NR_T;
  for (int i = lane; i < NR_T; i += 32) tile[i] = make_double2(0.0, 0.0);
  __syncwarp();
  int base = 0, used = 0;
  const double delta = A.grid.delta, emin = A.grid.emin;
  const long long nchunks = (A.nrec + NR_CHUNK - 1) / NR_CHUNK;

  for (;;) {
    unsigned long long ch = 0;
    if (lane == 0) ch = atomicAdd(A.next_chunk, 1ull);
    ch = __shfl_sync(0xffffffffu, ch, 0);
    if ((long long)ch >= nchunks) break;
    const long long c0 = (long long)ch * NR_CHUNK;
    const long long c1 = min(c0 + (long long)NR_CHUNK, A.nrec);

    for (long long g0 = c0; g0 < c1; g0 += 32) {
      const long long idx = g0 + lane;
      const bool valid = idx < c1;
      // ---- my record
      double e0 = 0, p1 = 1, p2 = 0, p3 = 0, p4 = 0, om = 0;
      int klo = 0x3fffffff, wlen = 0, cpos = 0;
      if (valid) {
        if (LINEAR) {
          const LRec r = reinterpret_cast<const LRec *>(A.recs)[idx];
          e0 = r.e0; p1 = r.e1; p2 = r.e2; p3 = r.e3; p4 = r.e4; om = r.om; klo = r.klo; wlen = r.wlen; cpos = r.cpos;
        } else {
          const GRec r = reinterpret_cast<const GRec *>(A.recs)[idx];
          e0 = r.e; p1 = r.w; om = r.om; klo = r.klo; wlen = r.wlen; cpos = r.cpos;
        }
      }
      // ---- the group's common window [gl, gl + Wc)
      int gl = klo, gh = valid ? klo + wlen - 1 : -0x3fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        gl = min(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        gh = max(gh, __shfl_xor_sync(0xffffffffu, gh, o));
      }
      int Wc = gh - gl + 1;
      if (Wc < 32) Wc = 32;
      const int spin_off = (klo >= A.bpad && valid) ? A.bpad : 0;   // ns <= 2
      const int jb = klo - spin_off;                                 // my first bin on the grid

      if (Wc > NR_T - A.cell) {
        // oversize group (sparse data, or it straddles the two spin channels): record by record,
        // lanes over bins, straight to HBM.
        for (int l = 0; l < 32; ++l) {
          const int k_l = __shfl_sync(0xffffffffu, klo, l), w_l = __shfl_sync(0xffffffffu, wlen, l), c_l = __shfl_sync(0xffffffffu, cpos, l);
          const int jb_l = __shfl_sync(0xffffffffu, jb, l);
          const double a0 = __shfl_sync(0xffffffffu, e0, l), a1 = __shfl_sync(0xffffffffu, p1, l), a2 = __shfl_sync(0xffffffffu, p2, l);
          const double a3 = __shfl_sync(0xffffffffu, p3, l), a4 = __shfl_sync(0xffffffffu, p4, l), ao = __shfl_sync(0xffffffffu, om, l);
          for (int r = lane; r < w_l; r += 32) {
            const double Ej = emin + (double)(jb_l + r) * delta;
            double d, s;
            if (LINEAR) {
              double cum; int bad = 0;
              d = od_doslin(a0, a1, a2, a3, a4, Ej, &cum, &bad) * ao;
              s = ((r < c_l) ? cum : cum - 1.0) * ao;
            } else {
              gauss_direct(a0, a1, ao, Ej, r, c_l, d, s);
            }
            if (d != 0.0) atomicAdd(&A.out_dos[k_l + r], d);
            if (s != 0.0) atomicAdd(&A.out_int[k_l + r], s);
          }
        }
        continue;
      }
      if (gl < base || gl + Wc > base + NR_T) {
        flush_tile(tile, base, used, A.out_dos, A.out_int, lane);
        base = gl & ~(A.cell - 1);
        used = 0;
      }
      const int off = gl - base;
      used = max(used, off + Wc);
      double2 *tw = tile + off;

      // ---- my window in group-relative bins [a, b]; the two visiting segments start at sA and sB
      const int a = valid ? klo - gl : 0x3fffffff, b = valid ? a + wlen - 1 : -1;
      const int cabs = a + cpos;                       // first relative bin at / above my centre
      const int sA = max(a, lane);                     // first segment [sA, b]   (visited before the wrap)
      const int sB = a;                                // second segment [a, min(b, lane-1)] (after the wrap)
      const double EA = emin + (double)(jb + (sA - a)) * delta;
      const double EB = emin + (double)jb * delta;

      if (!LINEAR) {
        const double invw = 1.0 / p1, del = delta * invw;
        const double amp = om * invw * OD_INV_SQRT_TWO_PI;
        const double cl = p1 * 1.2533141373155002512;  // w sqrt(pi/2): omega/2 * g' = g * cl
        const double q = exp(-del * del), hd2 = 0.5 * del * del;
        const double zA = (EA - e0) * invw, zB = (EB - e0) * invw;
        const double gA = amp * exp(-0.5 * zA * zA), rA = exp(-fma(zA, del, hd2));
        const double gB = amp * exp(-0.5 * zB * zB), rB = exp(-fma(zB, del, hd2));
        // DOS cut-off window (0.5 z^2 > 30  <=>  |E - e| > sqrt(60) w), relative bins
        const int dlo = a + (int)ceil(((e0 - OD_SQRT60 * p1) - EB) / delta);
        const int dhi = a + (int)floor(((e0 + OD_SQRT60 * p1) - EB) / delta);
        double g = 0.0, rr = 1.0, z = 0.0;
        int r = lane;
#pragma unroll 1
        for (int t = 0; t < Wc; ++t) {
          if (r == sA) { g = gA; rr = rA; z = zA; }
          if (r == sB) { g = gB; rr = rB; z = zB; }
          const bool in = (r >= a) && (r <= b);
          double2 acc = tw[r];
          const double h = g * erfcx_ratio(fabs(z)) * cl;
          const double hs = (r < cabs) ? h : -h;
          acc.x += (in && r >= dlo && r <= dhi) ? g : 0.0;
          acc.y += in ? hs : 0.0;
          tw[r] = acc;
          g *= rr; rr *= q; z += del;
          ++r;
          if (r >= Wc) r -= Wc;
          __syncwarp();
        }
      } else {
        // doslin (dos_utils.f90:1391-1528) in the variable x = e_use - e4 with a = e3-e4 <= bb = e2-e4 <= c = e1-e4 <= d = e0-e4
        const double aa = p3 - p4, bb = p2 - p4, cc = p1 - p4, dd = e0 - p4;
        const bool flat = !(cc + bb - aa > 0.0);
        const double hh = flat ? 0.5 / dd : 0.5 * (cc + bb - aa);
        const double full = (cc - aa) * cc + aa * aa / 3.0 - (cc - bb) * (cc - bb) / 3.0;
        const double alpha = flat ? 1.0 : 1.0 / (full + (cc + bb - aa) * (dd - cc));
        const double sc = alpha * om;
        const double inv2a = aa > 0.0 ? 0.5 / aa : 0.0, inv2cb = cc > bb ? 0.5 / (cc - bb) : 0.0;
        const double a2_6 = aa * aa / 6.0;
        const double k30 = 0.5 * (aa * aa / 3.0 - cc * bb) - (cc - bb) * (cc - bb) / 6.0;
        const double k40 = flat ? 0.0 : 0.5 * full - hh * cc;
        const double two_d = 2.0 * dd;
        const double xA = EA - p4, xB = EB - p4;
        const long long ia = __double_as_longlong(aa), ib = __double_as_longlong(bb), ic = __double_as_longlong(cc);
        double xl = 0.0;
        int r = lane;
#pragma unroll 1
        for (int t = 0; t < Wc; ++t) {
          if (r == sA) xl = xA;
          if (r == sB) xl = xB;
          const bool in = (r >= a) && (r <= b);
          const bool below = r < cabs;               // E_j <= e0
          double x = below ? xl : two_d - xl;         // e_use - e4
          long long ix = __double_as_longlong(x);
          if (ix < 0) { ix = 0; x = 0.0; }            // edge bins just outside (e4, 2 e0 - e4)
          const double x2 = x * x;
          const double D1 = x2 * inv2a, I1 = D1 * x * (1.0 / 3.0);
          const double xa = x - aa;
          const double D2 = x - 0.5 * aa, I2 = fma(0.5 * x, xa, a2_6);
          const double y = cc - x, Y = y * y * inv2cb;
          const double D3 = hh - Y, I3 = fma(Y * y, 1.0 / 3.0, fma(hh, x, k30));
          const double D4 = hh, I4 = fma(hh, x, k40);
          const bool s1 = ix < ia, s2 = ix < ib, s3 = ix < ic;
          const double D = s1 ? D1 : (s2 ? D2 : (s3 ? D3 : D4));
          const double I = s1 ? I1 : (s2 ? I2 : (s3 ? I3 : I4));
          double2 acc = tw[r];
          const bool on = in && ix > 0;                // E_j inside (e4, 2 e0 - e4)
          acc.x += on ? D * sc : 0.0;
          acc.y += on ? (below ? I * sc : -(I * sc)) : 0.0;
          tw[r] = acc;
          xl += delta;
          ++r;
          if (r >= Wc) r -= Wc;
          __syncwarp();
        }
      }
    }
    // keep the tile across chunks: consecutive chunks of one warp are usually far apart, so flush now
    flush_tile(tile, base, used, A.out_dos, A.out_int, lane);
    used = 0;
  }
}

// dense-engine producer over the compacted list of wide units
struct WideListProducer {
  long long units_per_z;  // max over spins of the wide count
  DosProducer inner;
  const unsigned *list;
  long long stride;
  unsigned long long count[2];
  __device__ __forceinline__ void operator()(long long u, int z, RecOut &r) const {
    r.mode = -1;
    if ((unsigned long long)u >= count[z]) return;
    inner((long long)list[(size_t)z * stride + u], z, r);
  }
};

// out(j, is): dos += narrow dos; intdos += narrow smooth part + inclusive prefix sum of the step histogram
__global__ void __launch_bounds__(1024) combine_kernel(const double *__restrict__ n_dos, const double *__restrict__ n_int,
                                                       const double *__restrict__ stephist, int nbins, int bpad,
                                                       double *__restrict__ dos, double *__restrict__ intdos) {
  __shared__ double sh[1024];
  __shared__ double carry;
  const int is = blockIdx.x;
  if (threadIdx.x == 0) carry = 0.0;
  __syncthreads();
  for (int base = 0; base < nbins; base += 1024) {
    const int j = base + threadIdx.x;
    const double v = (j < nbins) ? stephist[(size_t)is * bpad + j] : 0.0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const double t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0.0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (j < nbins) {
      dos[(size_t)is * nbins + j] += n_dos[(size_t)is * bpad + j];
      intdos[(size_t)is * nbins + j] += n_int[(size_t)is * bpad + j] + (carry + sh[threadIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}

}  // namespace

// Narrow pipeline for calculate_dos without weights.  accum = [dos(B,ns) | intdos(B,ns)], zeroed by the caller.
int od_narrow_dos(od_ctx *ctx, const DosProducer &prod, const GridSpec &grid, bool linear, double *accum) {
  const int ns = prod.ns, nbins = grid.nbins;
  const long long total = prod.units_per_z * ns;
  if (total >= (1ll << 32)) return od_fail(ctx, OD_ERR_ARG, "narrow engine: more than 2^32 units per call");
  PrepSpec ps;
  ps.grid = grid;
  ps.cell = 4;
  for (;;) {
    ps.bpad = ((nbins + NR_T + ps.cell - 1) / ps.cell) * ps.cell;
    ps.nbuckets = ns * ps.bpad / ps.cell;
    if (ps.nbuckets <= NR_MAXBUCKETS) break;
    ps.cell *= 2;
  }
  if (ps.cell > 64) return od_fail(ctx, OD_ERR_ARG, "narrow engine: energy grid too large (%d bins)", nbins);

  // scratch: keys | hist | offsets | cursor | wide counters | next_chunk | stephist | narrow dos | narrow int
  const size_t nout = (size_t)ns * ps.bpad + NR_T;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_hist = take(sizeof(unsigned) * ps.nbuckets);
  const size_t o_offs = take(sizeof(unsigned long long) * (ps.nbuckets + 1));
  const size_t o_curs = take(sizeof(unsigned long long) * ps.nbuckets);
  const size_t o_wide = take(sizeof(unsigned long long) * 8);
  const size_t o_step = take(sizeof(double) * nout);
  const size_t o_ndos = take(sizeof(double) * nout);
  const size_t o_nint = take(sizeof(double) * nout);
  const size_t zero_bytes = off;
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_a, off));
  OD_CHECK(od_buf_ensure(ctx, ctx->keys, sizeof(unsigned) * (size_t)total + 1024));
  char *sa = ctx->scratch_a.as<char>();
  OD_CUDA(ctx, cudaMemsetAsync(sa, 0, zero_bytes, ctx->stream));
  unsigned *keys = ctx->keys.as<unsigned>();
  unsigned *hist = reinterpret_cast<unsigned *>(sa + o_hist);
  unsigned long long *offsets = reinterpret_cast<unsigned long long *>(sa + o_offs);
  unsigned long long *cursor = reinterpret_cast<unsigned long long *>(sa + o_curs);
  unsigned long long *wide = reinterpret_cast<unsigned long long *>(sa + o_wide);  // [0..1] counts, [2..3] cursors, [4] next_chunk
  double *stephist = reinterpret_cast<double *>(sa + o_step);
  double *ndos = reinterpret_cast<double *>(sa + o_ndos);
  double *nint = reinterpret_cast<double *>(sa + o_nint);

  const int pblocks = ctx->sm_count * 8;
  const size_t hist_smem = sizeof(unsigned) * ps.nbuckets;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
  OD_CUDA(ctx, cudaEventCreate(&e0));
  OD_CUDA(ctx, cudaEventCreate(&e1));
  OD_CUDA(ctx, cudaEventCreate(&e2));
  OD_CUDA(ctx, cudaEventCreate(&e3));
  OD_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  OD_CUDA(ctx, cudaFuncSetAttribute(prep_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
  prep_hist_kernel<<<pblocks, 256, hist_smem, ctx->stream>>>(prod, ps, keys, hist, wide);
  OD_LAUNCH_CHECK(ctx);
  scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, ps.nbuckets, offsets, cursor);
  OD_LAUNCH_CHECK(ctx);
  unsigned long long h_wide[2] = {0, 0}, h_nrec = 0;
  OD_CUDA(ctx, cudaMemcpyAsync(h_wide, wide, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaMemcpyAsync(&h_nrec, offsets + ps.nbuckets, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const long long wide_stride = (long long)std::max(h_wide[0], h_wide[1]);
  const size_t rec_bytes = linear ? sizeof(LRec) : sizeof(GRec);
  OD_CHECK(od_buf_ensure(ctx, ctx->recs, rec_bytes * (size_t)h_nrec + 256));
  OD_CHECK(od_buf_ensure(ctx, ctx->scratch_b, sizeof(unsigned) * (size_t)wide_stride * ns + 256));
  if (linear)
    prep_scatter_kernel<true><<<pblocks, 256, 0, ctx->stream>>>(prod, ps, keys, cursor, ctx->recs.p, stephist, ctx->scratch_b.as<unsigned>(),
                                                               wide + 2, wide_stride);
  else
    prep_scatter_kernel<false><<<pblocks, 256, 0, ctx->stream>>>(prod, ps, keys, cursor, ctx->recs.p, stephist, ctx->scratch_b.as<unsigned>(),
                                                                wide + 2, wide_stride);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaEventRecord(e1, ctx->stream));

  // ---- accumulate the narrow records
  OD_CUDA(ctx, cudaEventRecord(e2, ctx->stream));
  if (h_nrec > 0) {
    NarrowArgs A;
    A.recs = ctx->recs.p;
    A.nrec = (long long)h_nrec;
    A.grid = grid;
    A.bpad = ps.bpad;
    A.cell = ps.cell;
    A.next_chunk = wide + 4;
    A.out_dos = ndos;
    A.out_int = nint;
    const size_t smem = sizeof(double2) * NR_T * NR_WARPS;
    const long long nchunks = ((long long)h_nrec + NR_CHUNK - 1) / NR_CHUNK;
    const int blocks = (int)std::min<long long>((long long)ctx->sm_count * 2, (nchunks + NR_WARPS - 1) / NR_WARPS);
    if (linear) {
      OD_CUDA(ctx, cudaFuncSetAttribute(narrow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      narrow_kernel<true><<<blocks, NR_WARPS * 32, smem, ctx->stream>>>(A);
    } else {
      OD_CUDA(ctx, cudaFuncSetAttribute(narrow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      narrow_kernel<false><<<blocks, NR_WARPS * 32, smem, ctx->stream>>>(A);
    }
    OD_LAUNCH_CHECK(ctx);
  }
  OD_CUDA(ctx, cudaEventRecord(e3, ctx->stream));

  // ---- the wide units through the dense engine (adds into accum)
  if (wide_stride > 0) {
    WideListProducer wp;
    wp.units_per_z = wide_stride;
    wp.inner = prod;
    wp.list = ctx->scratch_b.as<unsigned>();
    wp.stride = wide_stride;
    wp.count[0] = h_wide[0];
    wp.count[1] = h_wide[1];
    const size_t n1 = (size_t)nbins * ns;
    std::vector<long long> dest(ns * 2);
    for (int is = 0; is < ns; ++is) {
      dest[is * 2 + 0] = (long long)is * nbins;
      dest[is * 2 + 1] = (long long)(n1 + (size_t)is * nbins);
    }
    OD_CHECK((od_run_dense<0, true, WideListProducer>(ctx, wp, grid, ns, dest, accum)));
  }
  combine_kernel<<<ns, 1024, 0, ctx->stream>>>(ndos, nint, stephist, nbins, ps.bpad, accum, accum + (size_t)nbins * ns);
  OD_LAUNCH_CHECK(ctx);
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms_prep = 0.f, ms_acc = 0.f;
  cudaEventElapsedTime(&ms_prep, e0, e1);
  cudaEventElapsedTime(&ms_acc, e2, e3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  cudaEventDestroy(e3);
  ctx->prof.prepare_ms += ms_prep;
  ctx->prof.accumulate_ms += ms_acc;
  ctx->prof.accumulate_launches += (h_nrec > 0) ? 1 : 0;
  ctx->prof.contributions += 0.0;
  return OD_OK;
}

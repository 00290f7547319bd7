// od_producers.cuh -- unit -> record producers shared by the dense and the narrow engines.
#pragma once
#include "od_accum.cuh"

// n / d for 32-bit n by multiplication (Granlund-Montgomery round-up method, exact for every n < 2^32): the prepare passes
// of the narrow engine split a unit index into (k-point, band) once per unit, and an integer division is ~20 instructions
struct FastDiv {
  unsigned d = 1, m = 1, s1 = 0, s2 = 0;
  FastDiv() = default;
  explicit FastDiv(unsigned div) : d(div) {
    unsigned l = 0;
    while ((1ull << l) < div) ++l;  // ceil(log2 d)
    m = (unsigned)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
    s1 = l < 1 ? l : 1;
    s2 = l > 1 ? l - 1 : 0;
  }
  __host__ __device__ __forceinline__ unsigned div(unsigned n) const {
#ifdef __CUDA_ARCH__
    const unsigned t = __umulhi(m, n);
#else
    const unsigned t = (unsigned)(((unsigned long long)m * n) >> 32);
#endif
    return (t + ((n - t) >> s1)) >> s2;
  }
};

// sqrt(x) to ~1 ulp for finite x >= 0 (rsqrt.approx.ftz.f64 carries ~22 bits; two Goldschmidt steps and one residual
// correction); tiny and huge arguments take the IEEE path
__device__ __forceinline__ double od_fast_sqrt(double x) {
  if (!(x > 1e-280 && x < 1e280)) return sqrt(x);
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double g = x * r, h = 0.5 * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double d = fma(-g, g, x);
  return fma(d, h, g);
}

// Broadening set-up common to calculate_dos / calculate_dos_at_e / calculate_jdos
// (dos_utils.f90:1208-1218, jdos_utils.f90:324-334)
struct BroadSpec {
  int type;  // 0 fixed, 1 adaptive, 2 linear
  double fixed_w, adapt_temp, delta_bins;
  int fbc;
  int honor_hybrid;  // calculate_dos ignores hybrid_linear (Q8); at_e and jdos honour it
  int hybrid;
  double hybrid_tol;
  double step[3];
  double rs[9];  // recip_lattice, column-major
  int fast_corners;  // narrow engine: closed-form sub-cell corners (od_sub_cell_corners_fast)
};

// width / corners of one unit from its centre energy and gradient
__device__ __forceinline__ void od_make_record(const BroadSpec &b, double centre, const double grad[3], double omega, RecOut &r) {
  r.e0 = centre;
  r.omega = omega;
  double width = 0.0;
  bool force_adaptive = false;
  double gnorm = 0.0;
  // sqrt(dot_product(grad,grad)): only the adaptive width and the hybrid_linear test use it (an IEEE square root is ~30
  // instructions, a seventh of the narrow engine's key pass); the narrow engine (fast_corners) takes it to ~1 ulp by
  // MUFU.RSQ64H + Goldschmidt, the exact engines keep the IEEE one
  if (b.type == 1 || (b.type == 2 && b.honor_hybrid && b.hybrid)) {
    const double g2 = grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2];
    gnorm = b.fast_corners ? od_fast_sqrt(g2) : sqrt(g2);
  }
  // hybrid_linear only changes anything for the linear scheme (for 'f' the reference tests an
  // unset gradient); dos_utils.f90:1758, jdos_utils.f90:365
  if (b.type == 2 && b.honor_hybrid && b.hybrid && (b.hybrid_tol > gnorm)) force_adaptive = true;
  if (b.type == 2 && !force_adaptive) {
    double EV[5];
    if (b.fast_corners) od_sub_cell_corners_fast(grad, b.step, b.rs, centre, EV);
    else od_sub_cell_corners(grad, b.step, b.rs, centre, EV);                        // dos_utils.f90:1240
    r.mode = 1;
    r.p1 = EV[1]; r.p2 = EV[2]; r.p3 = EV[3]; r.p4 = EV[4];
    return;
  }
  if (b.type == 0) width = b.fixed_w;
  // NB for type 'l' with force_adaptive the reference starts from width = 0 (jdos_utils.f90:324,368)
  if (b.type == 1 || force_adaptive) width = gnorm * b.adapt_temp;                   // dos_utils.f90:1241
  if (b.fbc && (width < b.delta_bins)) width = b.delta_bins;                          // dos_utils.f90:1244
  r.mode = 0;
  r.p1 = width; r.p2 = r.p3 = r.p4 = 0.0;
}

// ------------------------------------------------------------------ DOS / PDOS / core
struct DosProducer {
  long long units_per_z;  // nb * (k-points of this slab)
  long long k_first;      // first k-point of the slab: band_energy / kpoint_weight / matrix_weights use k_first + ik
  int nb, ns, nk;
  int nk_g;               // k-extent of the array `grads` points to (the whole shard, or one streamed slab)
  const double *bands, *grads, *kw;
  double eps;
  BroadSpec b;
  const double *mw;  // matrix_weights(norb, mw_nb, mw_nk, ns) or null
  int norb, mw_nb, mw_nk;
  int og;  // orbitals per z-slice group (== NW of the kernel)
  FastDiv nb_div;  // / nb (set by dos_producer)

  // The prepare passes of the narrow engine split a unit into its loads and the arithmetic on them, so that a thread can
  // have the loads of several units in flight at once (PREP_U units per thread and iteration; the passes are bound by
  // memory latency, not by bandwidth).  No weights here: those passes run with mw == nullptr.
  struct Raw {
    double e, g0, g1, g2, kw;
  };
#ifndef OD_PREP_U_DOS
#define OD_PREP_U_DOS 4
#endif
  static constexpr int PREP_U = OD_PREP_U_DOS;
  static constexpr bool PREP_KMAJOR = false;  // band-group-major rows: a CTA fills few buckets (see PrepMap, od_narrow.cu)
  __host__ __device__ long long inner() const { return nb; }  // units per k-point
  template <bool WT>
  __device__ __forceinline__ void load(unsigned u, int is, Raw &w) const {
    const unsigned ikl = nb_div.div(u);
    load_kw<WT>(ikl, u - ikl * (unsigned)nb, is, w);
  }
  // (k-point of the slab, band): the prepare passes know both, no division
  template <bool WT>
  __device__ __forceinline__ void load_kw(unsigned ikl, unsigned ib, int is, Raw &w) const {
    const long long ik = k_first + ikl;
    w.e = bands[ib + (size_t)nb * (is + (size_t)ns * ik)];
    w.g0 = w.g1 = w.g2 = 0.0;
    if (b.type != 0) {
      const size_t gi = ib + (size_t)nb * (3 * (ikl + (size_t)nk_g * is));
      w.g0 = grads[gi];
      w.g1 = grads[gi + nb];
      w.g2 = grads[gi + 2 * (size_t)nb];
    }
    w.kw = kw[ik];
  }
  template <bool WT>
  __device__ __forceinline__ void make(unsigned, int, const Raw &w, RecOut &r) const {
    const double g[3] = {w.g0, w.g1, w.g2};
    od_make_record(b, w.e, g, eps * w.kw, r);
  }

  __device__ __forceinline__ void operator()(long long u, int z, RecOut &r) const {
    const int is = z % ns, grp = z / ns;
    // units_per_z = nb * nk < 2^32 whenever the engines accept the job: 32-bit division
    const unsigned uu = (unsigned)u;
    const long long ikl = (units_per_z <= 0xffffffffll) ? (long long)(uu / (unsigned)nb) : u / nb;  // within the slab
    const int ib = (int)(u - ikl * nb);
    const long long ik = k_first + ikl;
    const double e = bands[ib + (size_t)nb * (is + (size_t)ns * ik)];
    double g[3] = {0.0, 0.0, 0.0};
    if (b.type != 0) {
      const size_t gi = ib + (size_t)nb * (3 * (ikl + (size_t)nk_g * is));
      g[0] = grads[gi];
      g[1] = grads[gi + nb];
      g[2] = grads[gi + 2 * (size_t)nb];
    }
    od_make_record(b, e, g, eps * kw[ik], r);
    if (mw) {
      const bool in = (ik < mw_nk) && (ib < mw_nb);  // dos_utils.f90:1268-1269
      const size_t row = (size_t)norb * (ib + (size_t)mw_nb * (ik + (size_t)mw_nk * is));
#pragma unroll
      for (int a = 0; a < OD_MAXW; ++a) {
        const int o = grp * og + a;
        r.wt[a] = (in && a < og && o < norb) ? mw[row + o] : 0.0;
      }
    }
  }
};

// ------------------------------------------------------------------ JDOS / optics
// optics weights in contracted form: weight_g = factor * sum_{i<=j} C[g][ij] * Re(M_i conj(M_j)),
// C built on the host from optics_geom, q and the symmetry operations (see od_weights.cu).
struct OpticsCoef {
  int n_geom;
  double c[6][6];  // [geom component][xx, yy, zz, xy, xz, yz]  (off-diagonals already doubled)
};

__device__ __forceinline__ void od_ome_products(const double2 a, const double2 b, const double2 c, double P[6]) {
  // a, b, c: the three cartesian components (re, im) of one optical matrix element
  P[0] = a.x * a.x + a.y * a.y;
  P[1] = b.x * b.x + b.y * b.y;
  P[2] = c.x * c.x + c.y * c.y;
  P[3] = a.x * b.x + a.y * b.y;
  P[4] = a.x * c.x + a.y * c.y;
  P[5] = b.x * c.x + b.y * c.y;
}

__device__ __forceinline__ void od_ome_products(const double *om, size_t idx, size_t comp_stride, double P[6]) {
  // om: complex (re,im) pairs; element (n1,n2,i) at idx + i*comp_stride
  const double2 *m = reinterpret_cast<const double2 *>(om);
  od_ome_products(m[idx], m[idx + comp_stride], m[idx + 2 * comp_stride], P);
}

struct JdosProducer {
  long long units_per_z;  // pairs_per_k() * (k-points of this slab)
  long long k_first;      // first k-point of the slab (the arrays are indexed with the rank's full nk)
  int nb, ns, nk;
  // candidate bands (narrow engine): only bands that are occupied at some (k, spin) can be the lower band of a
  // pair, only bands that are unoccupied somewhere the upper one; null = all nb x nb slots
  const int *occ_list, *unocc_list;
  int n_occ, n_unocc;
  FastDiv nocc_div;  // / n_occ (candidate lists) or / nb (all slots); set by the narrow engine before its prepare passes
  int eager;  // issue all loads of a slot before testing it (scatter pass of the narrow engine)
  __host__ __device__ long long pairs_per_k() const { return occ_list ? (long long)n_occ * n_unocc : (long long)nb * nb; }
  const double *bands, *grads, *kw;
  double eps, efermi, scissor;
  BroadSpec b;
  const int *exclude;  // device, 1-based band indices
  int n_exclude;
  const double *mw;    // matrix_weights(nb,nb,nk,ns,n_geom) or null
  const double *ome;   // optical_mat(nb,nb,3,nk,ns) complex or null (fused make_weights)
  int n_geom;
  OpticsCoef coef;

  // load / make split for the prepare passes of the narrow engine (see DosProducer): every load of a slot is issued
  // before the first test; WT = also the weight columns (scatter pass).  Slabs hold < 2^30 pair slots.
  struct Raw {
    double ei, ej, g0, g1, g2, kw;
    double2 m0, m1, m2;
    double wt[6];
    int ib, jb;
  };
  static constexpr int PREP_U = 2;
  static constexpr bool PREP_KMAJOR = true;   // see PrepMap (od_narrow.cu)
  __host__ __device__ long long inner() const { return pairs_per_k(); }  // units (pair slots) per k-point
  __device__ __forceinline__ void slot(unsigned uu, int &ib, int &jb, long long &ik) const {
    if (occ_list) {
      const unsigned t = uu / (unsigned)n_occ, tj = t / (unsigned)n_unocc;
      ib = occ_list[uu - t * (unsigned)n_occ];
      jb = unocc_list[t - tj * (unsigned)n_unocc];
      ik = k_first + tj;
    } else {
      const unsigned t = uu / (unsigned)nb, tj = t / (unsigned)nb;
      ib = (int)(uu - t * (unsigned)nb);
      jb = (int)(t - tj * (unsigned)nb);
      ik = k_first + tj;
    }
  }
  template <bool WT>
  __device__ __forceinline__ void load(unsigned uu, int is, Raw &w) const {
    int ib, jb;
    long long ik;
    slot(uu, ib, jb, ik);
    load_at<WT>(ib, jb, ik, is, w);
  }
  // (k-point of the slab, pair slot of that k-point): one multiply-shift division instead of two hardware ones
  template <bool WT>
  __device__ __forceinline__ void load_kw(unsigned ikl, unsigned wslot, int is, Raw &w) const {
    int ib, jb;
    if (occ_list) {
      const unsigned t = nocc_div.div(wslot);
      ib = occ_list[wslot - t * (unsigned)n_occ];
      jb = unocc_list[t];
    } else {
      const unsigned t = nocc_div.div(wslot);  // (set to / nb without candidate lists)
      ib = (int)(wslot - t * (unsigned)nb);
      jb = (int)t;
    }
    load_at<WT>(ib, jb, k_first + ikl, is, w);
  }
  template <bool WT>
  __device__ __forceinline__ void load_at(int ib, int jb, long long ik, int is, Raw &w) const {
    w.ib = ib; w.jb = jb;
    const size_t eb = (size_t)nb * (is + (size_t)ns * ik);
    w.ei = bands[eb + ib]; w.ej = bands[eb + jb];
    w.g0 = w.g1 = w.g2 = 0.0;
    if (b.type != 0) {
      const size_t gi = (size_t)nb * (3 * (ik + (size_t)nk * is));
      const double a0 = grads[gi + jb], a1 = grads[gi + ib], b0 = grads[gi + nb + jb], b1 = grads[gi + nb + ib];
      const double c0 = grads[gi + 2 * (size_t)nb + jb], c1 = grads[gi + 2 * (size_t)nb + ib];
      w.g0 = a0 - a1; w.g1 = b0 - b1; w.g2 = c0 - c1;
    }
    w.kw = kw[ik];
    if (WT) {
      const size_t nb2 = (size_t)nb * nb;
      if (ome) {
        const size_t oi = ib + (size_t)nb * jb + nb2 * 3 * (ik + (size_t)nk * is);
        w.m0 = reinterpret_cast<const double2 *>(ome)[oi];
        w.m1 = reinterpret_cast<const double2 *>(ome)[oi + nb2];
        w.m2 = reinterpret_cast<const double2 *>(ome)[oi + 2 * nb2];
      } else if (mw) {
#pragma unroll
        for (int a = 0; a < 6; ++a)
          if (a < n_geom) w.wt[a] = mw[ib + (size_t)nb * (jb + (size_t)nb * (ik + (size_t)nk * (is + (size_t)ns * a)))];
      }
    }
  }
  template <bool WT>
  __device__ __forceinline__ void make(unsigned, int, const Raw &w, RecOut &r) const {
    r.mode = -1;
    const int ib = w.ib, jb = w.jb;
    for (int x = 0; x < n_exclude; ++x)
      if (exclude[x] == ib + 1) return;                         // jdos_utils.f90:355-357
    const double ei = w.ei, ej = w.ej;
    if (ei >= efermi) return;                                   // :358
    if (ej < efermi) return;                                    // :360
    const double g[3] = {w.g0, w.g1, w.g2};
    od_make_record(b, ej - ei + scissor, g, eps * w.kw, r);   // :366-372, :383
    if (!WT) return;
    if (mw) {
#pragma unroll
      for (int a = 0; a < 6; ++a) r.wt[a] = a < n_geom ? w.wt[a] : 0.0;
    } else if (ome) {
      // make_weights fused (optics.f90:259-269): only n2 >= n1 is ever filled; strict >/< tests
      double wt[6] = {0, 0, 0, 0, 0, 0};
      if (jb >= ib && !(ei > efermi && ib != jb) && !(ej < efermi && ib != jb)) {
        double factor = 0.0;
        if (jb == ib) factor = 1.0;
        else if (ej > efermi) factor = 1.0 / ((ej - ei + scissor) * (ej - ei + scissor));
        double P[6];
        od_ome_products(w.m0, w.m1, w.m2, P);
#pragma unroll
        for (int a = 0; a < 6; ++a)
          if (a < n_geom)
            wt[a] = factor * (coef.c[a][0] * P[0] + coef.c[a][1] * P[1] + coef.c[a][2] * P[2] + coef.c[a][3] * P[3] +
                              coef.c[a][4] * P[4] + coef.c[a][5] * P[5]);
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) r.wt[a] = wt[a];
    }
  }

  __device__ __forceinline__ void operator()(long long u, int z, RecOut &r) const {
    const int is = z;
    int ib, jb;
    long long ik;
    if (occ_list) {
      const unsigned uu = (unsigned)u;  // slabs hold < 2^30 pair slots
      const unsigned t = uu / (unsigned)n_occ, tj = t / (unsigned)n_unocc;
      ib = occ_list[uu - t * (unsigned)n_occ];
      jb = unocc_list[t - tj * (unsigned)n_unocc];
      ik = k_first + tj;
    } else {
      ib = (int)(u % nb);
      const long long t = u / nb;
      jb = (int)(t % nb);
      ik = k_first + t / nb;
    }
    r.mode = -1;
    const size_t eb = (size_t)nb * (is + (size_t)ns * ik);
    const size_t gi = (size_t)nb * (3 * (ik + (size_t)nk * is));
    const size_t nb2 = (size_t)nb * nb;
    const size_t oi = ib + (size_t)nb * jb + nb2 * 3 * (ik + (size_t)nk * is);
    double g[3] = {0.0, 0.0, 0.0};
    double2 m0 = make_double2(0.0, 0.0), m1 = m0, m2 = m0;
    double ei, ej;
    if (eager) {
      // scatter pass: the key already says this slot is a pair, so issue every load before the first use
      ei = bands[eb + ib]; ej = bands[eb + jb];
      if (b.type != 0) {
        const double a0 = grads[gi + jb], a1 = grads[gi + ib], b0 = grads[gi + nb + jb], b1 = grads[gi + nb + ib];
        const double c0 = grads[gi + 2 * (size_t)nb + jb], c1 = grads[gi + 2 * (size_t)nb + ib];
        g[0] = a0 - a1; g[1] = b0 - b1; g[2] = c0 - c1;
      }
      if (ome) {
        m0 = reinterpret_cast<const double2 *>(ome)[oi];
        m1 = reinterpret_cast<const double2 *>(ome)[oi + nb2];
        m2 = reinterpret_cast<const double2 *>(ome)[oi + 2 * nb2];
      }
    }
    for (int x = 0; x < n_exclude; ++x)
      if (exclude[x] == ib + 1) return;                         // jdos_utils.f90:355-357
    if (!eager) { ei = bands[eb + ib]; ej = bands[eb + jb]; }
    if (ei >= efermi) return;                                   // :358
    if (ej < efermi) return;                                    // :360
    if (!eager && b.type != 0) {
      g[0] = grads[gi + jb] - grads[gi + ib];
      g[1] = grads[gi + nb + jb] - grads[gi + nb + ib];
      g[2] = grads[gi + 2 * (size_t)nb + jb] - grads[gi + 2 * (size_t)nb + ib];
    }
    od_make_record(b, ej - ei + scissor, g, eps * kw[ik], r);   // :366-372, :383
    if (mw) {
#pragma unroll
      for (int a = 0; a < 6; ++a)
        if (a < n_geom) r.wt[a] = mw[ib + (size_t)nb * (jb + (size_t)nb * (ik + (size_t)nk * (is + (size_t)ns * a)))];
    } else if (ome) {
      // make_weights fused (optics.f90:259-269): only n2 >= n1 is ever filled; strict >/< tests
      double wt[6] = {0, 0, 0, 0, 0, 0};
      if (jb >= ib && !(ei > efermi && ib != jb) && !(ej < efermi && ib != jb)) {
        double factor = 0.0;
        if (jb == ib) factor = 1.0;
        else if (ej > efermi) factor = 1.0 / ((ej - ei + scissor) * (ej - ei + scissor));
        double P[6];
        if (!eager) {
          m0 = reinterpret_cast<const double2 *>(ome)[oi];
          m1 = reinterpret_cast<const double2 *>(ome)[oi + nb2];
          m2 = reinterpret_cast<const double2 *>(ome)[oi + 2 * nb2];
        }
        od_ome_products(m0, m1, m2, P);
#pragma unroll
        for (int a = 0; a < 6; ++a)
          if (a < n_geom)
            wt[a] = factor * (coef.c[a][0] * P[0] + coef.c[a][1] * P[1] + coef.c[a][2] * P[2] + coef.c[a][3] * P[3] +
                              coef.c[a][4] * P[4] + coef.c[a][5] * P[5]);
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) r.wt[a] = wt[a];
    }
  }
};

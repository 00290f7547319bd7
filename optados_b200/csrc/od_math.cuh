// od_math.cuh -- scalar device math of the spectral-integration kernels.
// "faithful" routines follow the reference operation for operation (citations: optados/src/).
#pragma once
#include <cuda_runtime.h>

#define OD_INV_SQRT_TWO_PI 0.3989422804014326779399460599  // constants.f90:44
#define OD_SQRT_TWO 1.414213562373095048801688724209698079 // constants.f90:46
#define OD_SQRT60 7.745966692414834                        // gaussian cutoff 0.5 z^2 > 30 (algorithms.f90:83)
#define OD_ERF_CUT 5.8                                     // NSWC erf saturation (algorithms.f90:292)

// algorithms.f90:73-91.  __dmul_rn/__dadd_rn are not needed: nvcc contracts a*b+c into FMA, which
// only changes the last bit; the cutoff test itself is evaluated on the same t as the exponent.
__device__ __forceinline__ double od_gaussian(double m, double w, double x) {
  const double t = (x - m) / w;
  const double h = 0.5 * (t * t);
  if (h > 30.0) return 0.0;
  return OD_INV_SQRT_TWO_PI * exp(-h) / w;
}

// algorithms.f90:252-306 (NSWC erf), restated.
__device__ __forceinline__ double od_erf_nswc(double x) {
  const double ax = fabs(x);
  double r;
  if (ax <= 0.5) {
    const double t = x * x;
    const double top = ((((0.771058495001320E-04 * t + -0.133733772997339E-02) * t + 0.323076579225834E-01) * t +
                         0.479137145607681E-01) * t + 0.128379167095513E+00) + 1.0;
    const double bot = ((0.301048631703895E-02 * t + 0.538971687740286E-01) * t + 0.375795757275549E+00) * t + 1.0;
    return x * (top / bot);
  } else if (ax <= 4.0) {
    const double top = ((((((-1.36864857382717E-07 * ax + 5.64195517478974E-01) * ax + 7.21175825088309E+00) * ax +
                           4.31622272220567E+01) * ax + 1.52989285046940E+02) * ax + 3.39320816734344E+02) * ax +
                        4.51918953711873E+02) * ax + 3.00459261020162E+02;
    const double bot = ((((((1.00000000000000E+00 * ax + 1.27827273196294E+01) * ax + 7.70001529352295E+01) * ax +
                           2.77585444743988E+02) * ax + 6.38980264465631E+02) * ax + 9.31354094850610E+02) * ax +
                        7.90950925327898E+02) * ax + 3.00459260956983E+02;
    r = 0.5 + (0.5 - exp(-x * x) * top / bot);
  } else if (ax <= OD_ERF_CUT) {
    const double x2 = x * x;
    const double t = 1.0 / x2;
    const double top = (((2.10144126479064E+00 * t + 2.62370141675169E+01) * t + 2.13688200555087E+01) * t +
                        4.65807828718470E+00) * t + 2.82094791773523E-01;
    const double bot = (((9.41537750555460E+01 * t + 1.87114811799590E+02) * t + 9.90191814623914E+01) * t +
                        1.80124575948747E+01) * t + 1.0;
    r = (0.564189583547756 - top / (x2 * bot)) / ax;
    r = 0.5 + (0.5 - exp(-x2) * r);
  } else {
    r = 1.0;
  }
  return (x < 0.0) ? -r : r;
}

// dos_utils.f90:1391-1528.  *bad is set (never cleared) on the two `stop` conditions.
__device__ __forceinline__ double od_doslin(double e0, double e1, double e2, double e3, double e4, double e,
                                            double *cuml, int *bad) {
  double doslin, in, alpha, e_use;
  bool shift;
  if (!((e4 <= e3) && (e3 <= e2) && (e2 <= e1) && (e1 <= e0))) {
    *bad = 1;
    *cuml = 0.0;
    return 0.0;
  }
  if (e <= e4) {
    *cuml = 0.0;
    return 0.0;
  }
  if (e >= (2.0 * e0 - e4)) {
    *cuml = 1.0;
    return 0.0;
  }
  if ((e1 == e2) && (e1 == e3) && (e3 == e4)) {
    alpha = 1.0;
  } else {
    alpha = 1.0 / ((e1 - e3) * (e1 - e4) + (e3 - e4) * (e3 - e4) / 3.0 - (e1 - e2) * (e1 - e2) / 3.0 +
                   (e1 + e2 - e3 - e4) * (e0 - e1));
  }
  if (e > e0) {
    e_use = 2.0 * e0 - e;
    shift = true;
  } else {
    e_use = e;
    shift = false;
  }
  if (e_use <= e4) {
    doslin = 0.0;
    in = 0.0;
  } else if (e_use < e3) {
    doslin = (e_use - e4) * (e_use - e4) / (e3 - e4) / 2.0;
    in = (e_use - e4) * (e_use - e4) * (e_use - e4) / (e3 - e4) / 6.0;
  } else if (e_use < e2) {
    doslin = (e_use - (e3 + e4) / 2.0);
    in = ((e3 - e4) * (e3 - e4) / 3.0 + (e_use - e3) * (e_use - e4)) / 2.0;
  } else if (e_use < e1) {
    doslin = (e1 + e2 - e3 - e4) / 2.0;
    if (e1 != e2) doslin = doslin - (e1 - e_use) * (e1 - e_use) / (e1 - e2) / 2.0;
    in = ((e2 - e4) * (e_use - e3) + (e1 - e3) * (e_use - e2) + (e3 - e4) * (e3 - e4) / 3.0 +
          ((e1 - e_use) * (e1 - e_use) * (e1 - e_use) - (e1 - e2) * (e1 - e2) * (e1 - e2)) / (e1 - e2) / 3.0) / 2.0;
  } else if (e_use <= e0) {
    if ((e1 + e2 - e3 - e4) > 0.0) {
      doslin = (e1 + e2 - e3 - e4) / 2.0;
      in = ((e1 - e3) * (e1 - e4) + (e3 - e4) * (e3 - e4) / 3.0 - (e1 - e2) * (e1 - e2) / 3.0 +
            (e1 + e2 - e3 - e4) * (e_use - e1)) / 2.0;
    } else {
      doslin = 1.0 / (e0 - e4) / 2.0;
      in = (e_use - e4) / (e0 - e4) / 2.0;
    }
  } else {
    *bad = 2;
    *cuml = 0.0;
    return 0.0;
  }
  doslin = doslin * alpha;
  in = shift ? 1.0 - in * alpha : in * alpha;
  *cuml = in;
  return doslin;
}

__device__ __forceinline__ void od_cswap_desc(double &a, double &b) {
  // after: a >= b
  const double hi = fmax(a, b), lo = fmin(a, b);
  a = hi;
  b = lo;
}

// dos_utils.f90:1320-1388.  rs = recip_lattice (column-major 3x3), step = 1/(2 N_i).
// The eight DE values are formed exactly as the reference does (matmul then dot_product), sorted
// descending (any correct sort yields the same multiset -> the same four lowest values), and
// EV(1..4) = E + DE(5..8).
// The same four lowest corners in closed form (narrow engine only; the dense engine keeps the corner-by-corner
// restatement above).  A corner offset is sum_i sigma_i a_i with a_i = step_i * (grad . recip_lattice(:, i)), so with
// the magnitudes sorted A >= B >= C the eight values are +-A +-B +-C and the four lowest are -A-B-C, -A-B+C, -A+B-C
// and min(-A+B+C, A-B-C): 3 dot products and 3 compare-exchanges instead of 8 mat-vecs and a 19-comparator
// network.  Differs from the reference arithmetic by rounding only (~1e-16 of |a|).
__device__ __forceinline__ void od_sub_cell_corners_fast(const double grad[3], const double step[3], const double *rs,
                                                         double energy, double EV[5]) {
  double A = fabs(step[0] * (grad[0] * rs[0] + grad[1] * rs[1] + grad[2] * rs[2]));
  double B = fabs(step[1] * (grad[0] * rs[3] + grad[1] * rs[4] + grad[2] * rs[5]));
  double C = fabs(step[2] * (grad[0] * rs[6] + grad[1] * rs[7] + grad[2] * rs[8]));
  double t;
  if (A < B) { t = A; A = B; B = t; }
  if (B < C) { t = B; B = C; C = t; }
  if (A < B) { t = A; A = B; B = t; }
  const double ab = A + B, amb = A - B;
  // (A - B) + C <= (A + B) - C exactly, but not after rounding when B == C: keep EV[3] <= EV[2] (doslin `stop`s on
  // unordered corners, dos_utils.f90:1430)
  const double x2 = amb + C, x3 = fmax(ab - C, x2);
  EV[0] = energy;
  EV[1] = energy + fmin(C - amb, amb - C);
  EV[2] = energy - x2;
  EV[3] = energy - x3;
  EV[4] = energy - fmax(ab + C, x3);
}

__device__ __forceinline__ void od_sub_cell_corners(const double grad[3], const double step[3], const double *rs,
                                                    double energy, double EV[5]) {
  double DE[8];
  int nn = 0;
#pragma unroll
  for (int m = -1; m <= 1; m += 2)
#pragma unroll
    for (int n = -1; n <= 1; n += 2)
#pragma unroll
      for (int o = -1; o <= 1; o += 2) {
        const double s0 = step[0] * m, s1 = step[1] * n, s2 = step[2] * o;
        const double t0 = rs[0] * s0 + rs[3] * s1 + rs[6] * s2;
        const double t1 = rs[1] * s0 + rs[4] * s1 + rs[7] * s2;
        const double t2 = rs[2] * s0 + rs[5] * s1 + rs[8] * s2;
        DE[nn++] = grad[0] * t0 + grad[1] * t1 + grad[2] * t2;
      }
  // 19-comparator sorting network for 8 keys (descending)
#define CS(i, j) od_cswap_desc(DE[i], DE[j])
  CS(0, 1); CS(2, 3); CS(4, 5); CS(6, 7);
  CS(0, 2); CS(1, 3); CS(4, 6); CS(5, 7);
  CS(1, 2); CS(5, 6); CS(0, 4); CS(3, 7);
  CS(1, 5); CS(2, 6);
  CS(1, 4); CS(3, 6);
  CS(2, 4); CS(3, 5);
  CS(3, 4);
#undef CS
  EV[0] = energy;
  EV[1] = energy + DE[4];
  EV[2] = energy + DE[5];
  EV[3] = energy + DE[6];
  EV[4] = energy + DE[7];
}

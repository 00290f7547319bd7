// od_io.cpp -- host-side decoders of the reference's input files.
//
// What the Fortran readers of the reference produce, produced here from the same files so that a host program
// without the Fortran front end (tests, tools, a C/C++ caller) can feed the C ABI:
//   <seed>.bands                       elec_read_band_energy   electronic.f90:445-618
//   <seed>.ome_bin / .cst_ome / .ome_fmt      elec_read_optical_mat   electronic.f90:302-442, od2od.f90:110-158
//   <seed>.dome_bin / .cst_vel / .dome_fmt    elec_read_band_gradient electronic.f90:170-299, od2od.f90:254-295
//   <seed>.pdos_bin / .pdos_weights / .pdos_fmt   elec_pdos_read      electronic.f90:1122-1276, od2od.f90:468-560
//   <seed>.elnes_bin / .eels_mat / .elnes_fmt     elec_read_elnes_mat electronic.f90:811-992, od2od.f90:639-706
//   <seed>-out.cell (symmetry_ops)     cell_get_symmetry       cell.f90:405-696
//   cell_calc_lattice cell.f90:924-980, cell_find_MP_grid cell.f90:87-318
// Binary files are Fortran sequential unformatted records (4-byte markers, sub-records for > 2 GiB); the byte
// order (big-endian under the reference's gfortran flags, make.system:38) is detected from the first marker.
// All arrays are returned in the reference's layouts (column-major) and units (eV, eV Ang).
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/optados_b200.h"

namespace {

thread_local std::string g_io_err;

int io_fail(int code, const char *fmt, const char *a = "", long long b = 0) {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a, b);
  g_io_err = buf;
  return code;
}

constexpr double kH2eV = 27.21138342902473;   // constants.f90:46
constexpr double kBohr2Ang = 0.52917720859;   // constants.f90:40

// ------------------------------------------------------------------ text
struct Text {
  std::string s;
  size_t p = 0;
  bool load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    s.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(&s[0], 1, (size_t)n, f) : 0;
    fclose(f);
    s.resize(got);
    return true;
  }
  bool eof() const { return p >= s.size(); }
  std::string line() {  // next line without its terminator
    const size_t e = s.find('\n', p);
    std::string r = s.substr(p, (e == std::string::npos ? s.size() : e) - p);
    p = (e == std::string::npos) ? s.size() : e + 1;
    if (!r.empty() && r.back() == '\r') r.pop_back();
    return r;
  }
  // next number of a free-format stream (Fortran list-directed / fixed-width reals: blanks, newlines and commas
  // separate; D exponents accepted)
  bool number(double *v) {
    while (p < s.size() && (isspace((unsigned char)s[p]) || s[p] == ',')) ++p;
    if (p >= s.size()) return false;
    char tmp[64];
    size_t n = 0;
    while (p < s.size() && n < sizeof tmp - 1 && !isspace((unsigned char)s[p]) && s[p] != ',') {
      char c = s[p++];
      if (c == 'D' || c == 'd') c = 'e';
      tmp[n++] = c;
    }
    tmp[n] = 0;
    char *end = nullptr;
    *v = strtod(tmp, &end);
    return end != tmp;
  }
};

bool numbers_in(const std::string &str, std::vector<double> *out) {
  Text t;
  t.s = str;
  double v;
  out->clear();
  while (t.number(&v)) out->push_back(v);
  return true;
}

// ------------------------------------------------------------------ Fortran sequential unformatted
struct FortranFile {
  FILE *f = nullptr;
  bool swap = false;
  ~FortranFile() { if (f) fclose(f); }
  static uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
  bool open(const char *path, long long first_record_bytes_hint) {
    f = fopen(path, "rb");
    if (!f) return false;
    uint32_t m = 0;
    if (fread(&m, 4, 1, f) != 1) return false;
    fseek(f, 0, SEEK_SET);
    // the first marker is the byte length of the first record: pick the byte order in which it is the expected
    // length (or, without a hint, the smaller one)
    const uint32_t sw = bswap32(m);
    if (first_record_bytes_hint > 0) {
      if (m == (uint32_t)first_record_bytes_hint) swap = false;
      else if (sw == (uint32_t)first_record_bytes_hint) swap = true;
      else swap = sw < m;
    } else {
      swap = sw < m;
    }
    return true;
  }
  bool marker(int32_t *v) {
    uint32_t m;
    if (fread(&m, 4, 1, f) != 1) return false;
    if (swap) m = bswap32(m);
    *v = (int32_t)m;
    return true;
  }
  // read one logical record of exactly `bytes` bytes made of elements of `elem` bytes (byte-swapped per element)
  bool record(void *dst, size_t bytes, size_t elem) {
    size_t got = 0;
    char *d = static_cast<char *>(dst);
    for (;;) {
      int32_t head;
      if (!marker(&head)) return false;
      const size_t len = (size_t)(head < 0 ? -(long long)head : head);  // negative: the record continues
      if (got + len > bytes) return false;
      if (len && fread(d + got, 1, len, f) != len) return false;
      got += len;
      int32_t tail;
      if (!marker(&tail)) return false;
      if (head >= 0) break;
    }
    if (got != bytes) return false;
    if (swap && elem > 1) {
      for (size_t i = 0; i + elem <= bytes; i += elem)
        for (size_t a = 0, b = elem - 1; a < b; ++a, --b) { const char t = d[i + a]; d[i + a] = d[i + b]; d[i + b] = t; }
    }
    return true;
  }
  bool skip_record() {
    for (;;) {
      int32_t head;
      if (!marker(&head)) return false;
      const long long len = head < 0 ? -(long long)head : head;
      if (fseek(f, (long)len, SEEK_CUR) != 0) return false;
      int32_t tail;
      if (!marker(&tail)) return false;
      if (head >= 0) return true;
    }
  }
  // new-format files start with a version record (real(dp)) and an 80-character header (electronic.f90:355-361)
  int header(const char *what) {
    double ver = 0.0;
    if (!record(&ver, 8, 8)) return io_fail(OD_ERR_IO, "cannot read the file version of %s", what);
    if (ver - 1.0 > 0.001) return io_fail(OD_ERR_IO, "Trying to read newer version of %s file", what);
    if (!skip_record()) return io_fail(OD_ERR_IO, "cannot read the header of %s", what);
    return OD_OK;
  }
};

inline double to_ev_ang(double x) { return x * kBohr2Ang * kH2eV; }  // electronic.f90:271 / :427, left to right
// *_fmt -> *_bin -> read, as the reference's test-suite does it (od2od.f90:152, :229; electronic.f90:427)
inline double fmt_round_trip(double x) {
  const double c = kBohr2Ang * kH2eV;
  double v = x * c;
  v = v / c;
  return v * kBohr2Ang * kH2eV;
}

bool skip_lines(Text &t, int n) {
  for (int i = 0; i < n; ++i) {
    if (t.eof()) return false;
    t.line();
  }
  return true;
}

int int_at(const std::string &l, size_t pos, size_t width) {
  if (l.size() <= pos) return 0;
  return atoi(l.substr(pos, width).c_str());
}

}  // namespace

extern "C" {

const char *od_io_last_error(void) { return g_io_err.c_str(); }

// ------------------------------------------------------------------ <seed>.bands
int od_read_bands_info(const char *path, od_bands_info *info) {
  if (!path || !info) return io_fail(OD_ERR_ARG, "null argument to od_read_bands_info%s");
  Text t;
  if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening bands file %s", path);
  memset(info, 0, sizeof *info);
  std::vector<double> v;
  std::string l = t.line();
  size_t pos = l.find("k-points");
  if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: no 'k-points' on line 1", path);
  info->nkpoints = atoi(l.c_str() + pos + 8);
  l = t.line();
  pos = l.find("components");
  if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: no 'components' on line 2", path);
  info->nspins = atoi(l.c_str() + pos + 10);
  if (info->nspins < 1 || info->nspins > 2) return io_fail(OD_ERR_IO, "%s: bad number of spin components", path);
  l = t.line();
  pos = l.find("electrons");
  if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: no 'electrons' on line 3", path);
  numbers_in(l.substr(pos + 10), &v);
  if ((int)v.size() < info->nspins) return io_fail(OD_ERR_IO, "%s: too few electron counts", path);
  for (int i = 0; i < info->nspins; ++i) info->num_electrons[i] = v[i];
  l = t.line();
  pos = l.find("eigenvalues");
  if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: no 'eigenvalues' on line 4", path);
  info->nbands = atoi(l.c_str() + pos + 11);
  l = t.line();
  pos = l.find("units)");
  if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: no 'units)' on line 5", path);
  info->efermi_castep = atof(l.substr(pos + 6, 12).c_str()) * kH2eV;  // '(f12.4)' (electronic.f90:511-512), then :602
  t.line();
  for (int i = 0; i < 3; ++i) {  // real_lattice(:, i) = line i, in Bohr
    numbers_in(t.line(), &v);
    if (v.size() < 3) return io_fail(OD_ERR_IO, "%s: bad lattice vector", path);
    for (int c = 0; c < 3; ++c) info->real_lattice[c + 3 * i] = v[c];
  }
  info->electrons_per_state = info->nspins < 2 ? 2.0 : 1.0;  // electronic.f90:604-610
  if (info->nkpoints < 1 || info->nbands < 1) return io_fail(OD_ERR_IO, "%s: bad k-point / band count", path);
  return OD_OK;
}

int od_read_bands(const char *path, const od_bands_info *info, double *band_energy, double *kpoint_r, double *kpoint_weight) {
  if (!path || !info || !band_energy) return io_fail(OD_ERR_ARG, "null argument to od_read_bands%s");
  Text t;
  if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening bands file %s", path);
  if (!skip_lines(t, 9)) return io_fail(OD_ERR_IO, "%s: truncated header", path);
  const int nk = info->nkpoints, ns = info->nspins, nb = info->nbands;
  std::vector<double> v;
  for (int ik = 0; ik < nk; ++ik) {
    const std::string l = t.line();
    const size_t pos = l.find("K-point");
    if (pos == std::string::npos) return io_fail(OD_ERR_IO, "%s: expected a 'K-point' line (k-point %lld)", path, ik + 1);
    numbers_in(l.substr(pos + 7), &v);
    if (v.size() < 5) return io_fail(OD_ERR_IO, "%s: bad 'K-point' line (k-point %lld)", path, ik + 1);
    if (kpoint_r) { kpoint_r[3 * (size_t)ik] = v[1]; kpoint_r[3 * (size_t)ik + 1] = v[2]; kpoint_r[3 * (size_t)ik + 2] = v[3]; }
    if (kpoint_weight) kpoint_weight[ik] = v[4];
    for (int is = 0; is < ns; ++is) {
      t.line();  // 'Spin component'
      for (int ib = 0; ib < nb; ++ib) {
        double e;
        if (!t.number(&e)) return io_fail(OD_ERR_IO, "%s: truncated eigenvalue list (k-point %lld)", path, ik + 1);
        band_energy[ib + (size_t)nb * (is + (size_t)ns * ik)] = e * kH2eV;  // band_energy(ib, is, ik), electronic.f90:601
      }
      t.line();  // rest of the last eigenvalue line
    }
  }
  return OD_OK;
}

// ------------------------------------------------------------------ cell
int od_cell_calc_lattice(const double *real_lattice_bohr, double *real_lattice, double *recip_lattice, double *cell_volume) {
  if (!real_lattice_bohr || !recip_lattice) return io_fail(OD_ERR_ARG, "null argument to od_cell_calc_lattice%s");
  double rl[9], rc[9];
  for (int i = 0; i < 9; ++i) rl[i] = real_lattice_bohr[i] * kBohr2Ang;  // electronic.f90 -> cell.f90:924
#define RL(r, c) rl[(r) + 3 * (c)]
#define RC(r, c) rc[(r) + 3 * (c)]
  RC(0, 0) = RL(1, 1) * RL(2, 2) - RL(2, 1) * RL(1, 2);
  RC(1, 0) = RL(1, 2) * RL(2, 0) - RL(2, 2) * RL(1, 0);
  RC(2, 0) = RL(1, 0) * RL(2, 1) - RL(2, 0) * RL(1, 1);
  RC(0, 1) = RL(2, 1) * RL(0, 2) - RL(0, 1) * RL(2, 2);
  RC(1, 1) = RL(2, 2) * RL(0, 0) - RL(0, 2) * RL(2, 0);
  RC(2, 1) = RL(2, 0) * RL(0, 1) - RL(0, 0) * RL(2, 1);
  RC(0, 2) = RL(0, 1) * RL(1, 2) - RL(1, 1) * RL(0, 2);
  RC(1, 2) = RL(0, 2) * RL(1, 0) - RL(1, 2) * RL(0, 0);
  RC(2, 2) = RL(0, 0) * RL(1, 1) - RL(1, 0) * RL(0, 1);
  double vol = RL(0, 0) * RC(0, 0) + RL(1, 0) * RC(0, 1) + RL(2, 0) * RC(0, 2);
#undef RL
#undef RC
  if (vol < 0.0) vol = -vol;
  if (!(vol > 0.0)) return io_fail(OD_ERR_ARG, "cell volume is zero%s");
  const double pi = 3.141592653589793238462643383279502884197;
  for (int i = 0; i < 9; ++i) recip_lattice[i] = rc[i] * pi * 2.0 / vol;
  if (real_lattice) for (int i = 0; i < 9; ++i) real_lattice[i] = rl[i];
  if (cell_volume) *cell_volume = vol;
  return OD_OK;
}

int od_cell_find_mp_grid(const double *kpoints, int nk, int *grid) {
  if (!kpoints || !grid || nk < 1) return io_fail(OD_ERR_ARG, "bad argument to od_cell_find_mp_grid%s");
  const double sub_tol = std::numeric_limits<double>::epsilon(), tol = 0.000001, big = std::numeric_limits<double>::max();
  std::vector<double> ktr(2 * (size_t)nk), uniq;
  for (int d = 0; d < 3; ++d) {
    for (int ik = 0; ik < nk; ++ik) {  // k and -k (time reversal), folded into (-0.5, 0.5]
      const double a = kpoints[d + 3 * (size_t)ik];
      ktr[ik] = a - floor(a + 0.5);
      ktr[nk + ik] = -a - floor(-a + 0.5);
    }
    uniq.clear();
    uniq.push_back(ktr[0]);
    for (size_t i = 1; i < ktr.size(); ++i) {
      bool seen = false;
      for (double u : uniq)
        if (fabs(u - ktr[i]) <= sub_tol) { seen = true; break; }
      if (!seen) uniq.push_back(ktr[i]);
    }
    const size_t n = uniq.size();
    if (n == 1) { grid[d] = 1; continue; }
    if (n == 2) { grid[d] = fabs(fabs(uniq[0] - uniq[1]) - 0.5) <= tol ? 2 : 1; continue; }
    double m1 = big, m2 = big, m3 = big;
    for (size_t i = 0; i < n; ++i)
      for (size_t j = i + 1; j < n; ++j) m1 = std::fmin(m1, fabs(uniq[i] - uniq[j]));
    for (size_t i = 0; i < n; ++i)
      for (size_t j = i; j < n; ++j) {
        const double im = fabs(uniq[i] - uniq[j]);
        if (im < m2 && im > m1 + tol) m2 = im;
      }
    for (size_t i = 0; i < n; ++i)
      for (size_t j = i; j < n; ++j) {
        const double im = fabs(uniq[i] - uniq[j]);
        if (im < m3 && im > m2 + tol) m3 = im;
      }
    grid[d] = fabs(2.0 * m1 - m2) < tol ? (int)(1.0 / m1) : (int)(1.0 / m3);
  }
  return OD_OK;
}

int od_read_symmetry_ops(const char *cell_path, int max_ops, double *ops, int *n_ops) {
  if (!cell_path || !n_ops) return io_fail(OD_ERR_ARG, "null argument to od_read_symmetry_ops%s");
  Text t;
  if (!t.load(cell_path)) return io_fail(OD_ERR_IO, "Problem opening cell file %s", cell_path);
  std::vector<std::string> rows;
  bool inside = false, found = false;
  while (!t.eof()) {
    std::string l = t.line();
    for (char &c : l) c = (char)tolower((unsigned char)c);
    const size_t cm = l.find_first_of("#!");
    if (cm != std::string::npos) l.erase(cm);
    if (l.find("%block") != std::string::npos && l.find("symmetry_ops") != std::string::npos) { inside = true; found = true; continue; }
    if (l.find("%endblock") != std::string::npos && l.find("symmetry_ops") != std::string::npos) { inside = false; continue; }
    if (inside && l.find_first_not_of(" \t") != std::string::npos) rows.push_back(l);
  }
  *n_ops = found ? (int)(rows.size() / 4) : 0;
  if (!ops) return OD_OK;
  if (*n_ops > max_ops) return io_fail(OD_ERR_ARG, "%s: more symmetry operations than max_ops", cell_path);
  std::vector<double> v;
  for (int n = 0; n < *n_ops; ++n)
    for (int r = 0; r < 3; ++r) {  // row r of operation n -> crystal_symmetry_operations(1:3, r, n)  (cell.f90:675-682)
      numbers_in(rows[4 * (size_t)n + r], &v);
      if (v.size() < 3) return io_fail(OD_ERR_IO, "%s: bad symmetry_ops row", cell_path);
      for (int c = 0; c < 3; ++c) ops[c + 3 * (r + 3 * (size_t)n)] = v[c];
    }
  return OD_OK;
}

// ------------------------------------------------------------------ optical matrix elements
int od_read_ome(const char *path, char fmt, int nb, int nk, int ns, double *optical_mat) {
  if (!path || !optical_mat || nb < 1 || nk < 0 || ns < 1) return io_fail(OD_ERR_ARG, "bad argument to od_read_ome%s");
  const size_t per = (size_t)nb * nb * 3;  // complex numbers per (ik, is) block, (ib, jb, i) with ib fastest
  if (fmt == 't') {
    Text t;
    if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening ome_fmt file %s", path);
    skip_lines(t, 2);
    for (size_t blk = 0; blk < (size_t)nk * ns; ++blk) {
      const size_t ik = blk / ns, is = blk % ns;
      double *dst = optical_mat + 2 * per * (ik + (size_t)nk * is);  // optical_mat(:, :, :, ik, is)
      for (size_t e = 0; e < 2 * per; ++e) {
        double x;
        if (!t.number(&x)) return io_fail(OD_ERR_IO, "%s: truncated (block %lld)", path, (long long)blk);
        dst[e] = fmt_round_trip(x);
      }
    }
    return OD_OK;
  }
  if (fmt != 'b' && fmt != 'l') return io_fail(OD_ERR_ARG, "od_read_ome: fmt must be 'b', 'l' or 't'%s");
  FortranFile f;
  if (!f.open(path, fmt == 'b' ? 8 : 16)) return io_fail(OD_ERR_IO, "Problem opening %s", path);
  if (fmt == 'b') {
    const int rc = f.header("ome_bin");
    if (rc) return rc;
  }
  for (size_t ik = 0; ik < (size_t)nk; ++ik)
    for (size_t is = 0; is < (size_t)ns; ++is) {
      double *dst = optical_mat + 2 * per * (ik + (size_t)nk * is);
      if (fmt == 'b') {
        if (!f.record(dst, 16 * per, 8)) return io_fail(OD_ERR_IO, "%s: short record (k-point %lld)", path, (long long)ik + 1);
        for (size_t e = 0; e < 2 * per; ++e) dst[e] = to_ev_ang(dst[e]);
      } else {  // legacy: one complex per record, extra factor bohr2ang (electronic.f90:423-425)
        for (size_t e = 0; e < per; ++e) {
          if (!f.record(dst + 2 * e, 16, 8)) return io_fail(OD_ERR_IO, "%s: short record (k-point %lld)", path, (long long)ik + 1);
          dst[2 * e] = dst[2 * e] * kBohr2Ang * kBohr2Ang * kH2eV;
          dst[2 * e + 1] = dst[2 * e + 1] * kBohr2Ang * kBohr2Ang * kH2eV;
        }
      }
    }
  return OD_OK;
}

// ------------------------------------------------------------------ band gradients
int od_read_dome(const char *path, char fmt, int nb, int nk, int ns, double *band_gradient) {
  if (!path || !band_gradient || nb < 1 || nk < 0 || ns < 1) return io_fail(OD_ERR_ARG, "bad argument to od_read_dome%s");
  const size_t per = (size_t)nb * 3;  // (ib, i) with ib fastest
  if (fmt == 't') {
    Text t;
    if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening dome_fmt file %s", path);
    skip_lines(t, 2);
    for (size_t blk = 0; blk < (size_t)nk * ns; ++blk) {
      const size_t ik = blk / ns, is = blk % ns;
      double *dst = band_gradient + per * (ik + (size_t)nk * is);  // band_gradient(:, :, ik, is)
      for (size_t e = 0; e < per; ++e) {
        double x;
        if (!t.number(&x)) return io_fail(OD_ERR_IO, "%s: truncated (block %lld)", path, (long long)blk);
        dst[e] = fmt_round_trip(x);
      }
    }
    return OD_OK;
  }
  if (fmt != 'b' && fmt != 'l') return io_fail(OD_ERR_ARG, "od_read_dome: fmt must be 'b', 'l' or 't'%s");
  FortranFile f;
  if (!f.open(path, fmt == 'b' ? 8 : (long long)(8 * per))) return io_fail(OD_ERR_IO, "Problem opening %s", path);
  if (fmt == 'b') {
    const int rc = f.header("dome_bin");
    if (rc) return rc;
  }
  for (size_t ik = 0; ik < (size_t)nk; ++ik)
    for (size_t is = 0; is < (size_t)ns; ++is) {
      double *dst = band_gradient + per * (ik + (size_t)nk * is);
      if (!f.record(dst, 8 * per, 8)) return io_fail(OD_ERR_IO, "%s: short record (k-point %lld)", path, (long long)ik + 1);
      for (size_t e = 0; e < per; ++e) dst[e] = to_ev_ang(dst[e]);  // electronic.f90:271
    }
  return OD_OK;
}

// ------------------------------------------------------------------ projector weights
static int pdos_open_bin(FortranFile &f, const char *path, char fmt, od_pdos_info *info) {
  if (!f.open(path, fmt == 'b' ? 8 : 4)) return io_fail(OD_ERR_IO, "Problem opening %s", path);
  if (fmt == 'b') {
    const int rc = f.header("pdos_bin");
    if (rc) return rc;
  }
  int32_t v[4];
  for (int i = 0; i < 4; ++i)
    if (!f.record(&v[i], 4, 4)) return io_fail(OD_ERR_IO, "%s: cannot read the dimensions", path);
  info->nkpoints = v[0]; info->nspins = v[1]; info->norbitals = v[2]; info->nbands = v[3];  // electronic.f90:1165-1168
  return OD_OK;
}

static int pdos_open_fmt(Text &t, const char *path, od_pdos_info *info) {
  if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening pdos_fmt file %s", path);
  skip_lines(t, 2);
  info->nkpoints = int_at(t.line(), 10, 6);   // '(a10, i6)'  od2od.f90:489-492
  info->nspins = int_at(t.line(), 10, 6);
  info->norbitals = int_at(t.line(), 10, 6);
  info->nbands = int_at(t.line(), 10, 6);
  return OD_OK;
}

int od_read_pdos_info(const char *path, char fmt, od_pdos_info *info) {
  if (!path || !info) return io_fail(OD_ERR_ARG, "null argument to od_read_pdos_info%s");
  int rc;
  if (fmt == 't') {
    Text t;
    rc = pdos_open_fmt(t, path, info);
  } else if (fmt == 'b' || fmt == 'l') {
    FortranFile f;
    rc = pdos_open_bin(f, path, fmt, info);
  } else {
    return io_fail(OD_ERR_ARG, "od_read_pdos_info: fmt must be 'b', 'l' or 't'%s");
  }
  if (rc) return rc;
  if (info->nkpoints < 1 || info->nspins < 1 || info->norbitals < 1 || info->nbands < 1) return io_fail(OD_ERR_IO, "%s: bad dimensions", path);
  return OD_OK;
}

int od_read_pdos(const char *path, char fmt, const od_pdos_info *info, int *species_no, int *rank_in_species, int *am_channel,
                 int *nbands_occ, double *pdos_weights) {
  if (!path || !info || !pdos_weights) return io_fail(OD_ERR_ARG, "null argument to od_read_pdos%s");
  const int nk = info->nkpoints, ns = info->nspins, norb = info->norbitals, nb = info->nbands;
  memset(pdos_weights, 0, sizeof(double) * (size_t)norb * nb * nk * ns);  // bands above nbands_occ stay zero
  std::vector<int32_t> meta((size_t)norb);
  int *dst_meta[3] = {species_no, rank_in_species, am_channel};
  od_pdos_info chk;
  if (fmt == 't') {
    Text t;
    int rc = pdos_open_fmt(t, path, &chk);
    if (rc) return rc;
    for (int m = 0; m < 3; ++m) {
      t.line();  // caption
      for (int o = 0; o < norb; ++o) {
        double x;
        if (!t.number(&x)) return io_fail(OD_ERR_IO, "%s: truncated orbital table", path);
        if (dst_meta[m]) dst_meta[m][o] = (int)x;
      }
      t.line();
    }
    for (int ik = 0; ik < nk; ++ik) {
      t.line();  // index and k-point
      for (int is = 0; is < ns; ++is) {
        t.line();  // spin number
        const int nocc = atoi(t.line().c_str());
        if (nocc < 0 || nocc > nb) return io_fail(OD_ERR_IO, "%s: bad nbands_occ (k-point %lld)", path, ik + 1);
        if (nbands_occ) nbands_occ[ik + (size_t)nk * is] = nocc;  // nbands_occ(ik, is)
        for (int ib = 0; ib < nocc; ++ib) {
          double *w = pdos_weights + (size_t)norb * (ib + (size_t)nb * (ik + (size_t)nk * is));
          for (int o = 0; o < norb; ++o)
            if (!t.number(&w[o])) return io_fail(OD_ERR_IO, "%s: truncated weights (k-point %lld)", path, ik + 1);
        }
        if (nocc > 0) t.line();
      }
    }
    return OD_OK;
  }
  if (fmt != 'b' && fmt != 'l') return io_fail(OD_ERR_ARG, "od_read_pdos: fmt must be 'b', 'l' or 't'%s");
  FortranFile f;
  int rc = pdos_open_bin(f, path, fmt, &chk);
  if (rc) return rc;
  if (chk.nkpoints != nk || chk.nspins != ns || chk.norbitals != norb || chk.nbands != nb) return io_fail(OD_ERR_ARG, "%s: dimensions differ from info", path);
  for (int m = 0; m < 3; ++m) {
    if (!f.record(meta.data(), 4 * (size_t)norb, 4)) return io_fail(OD_ERR_IO, "%s: cannot read the orbital table", path);
    if (dst_meta[m]) for (int o = 0; o < norb; ++o) dst_meta[m][o] = meta[o];
  }
  for (int ik = 0; ik < nk; ++ik) {
    if (!f.skip_record()) return io_fail(OD_ERR_IO, "%s: truncated (k-point %lld)", path, ik + 1);  // index + k-point
    for (int is = 0; is < ns; ++is) {
      int32_t spin, nocc;
      if (!f.record(&spin, 4, 4) || !f.record(&nocc, 4, 4)) return io_fail(OD_ERR_IO, "%s: truncated (k-point %lld)", path, ik + 1);
      if (nocc < 0 || nocc > nb) return io_fail(OD_ERR_IO, "%s: bad nbands_occ (k-point %lld)", path, ik + 1);
      if (nbands_occ) nbands_occ[ik + (size_t)nk * is] = nocc;
      for (int ib = 0; ib < nocc; ++ib)
        if (!f.record(pdos_weights + (size_t)norb * (ib + (size_t)nb * (ik + (size_t)nk * is)), 8 * (size_t)norb, 8))
          return io_fail(OD_ERR_IO, "%s: short weight record (k-point %lld)", path, ik + 1);
    }
  }
  return OD_OK;
}

// ------------------------------------------------------------------ ELNES matrix elements
static int elnes_open_bin(FortranFile &f, const char *path, char fmt, od_elnes_info *info) {
  if (!f.open(path, fmt == 'b' ? 8 : 4)) return io_fail(OD_ERR_IO, "Problem opening %s", path);
  if (fmt == 'b') {
    const int rc = f.header("elnes_bin");
    if (rc) return rc;
  }
  int32_t v[4];
  for (int i = 0; i < 4; ++i)
    if (!f.record(&v[i], 4, 4)) return io_fail(OD_ERR_IO, "%s: cannot read the dimensions", path);
  info->norbitals = v[0]; info->nbands = v[1]; info->nkpoints = v[2]; info->nspins = v[3];  // electronic.f90:866-869
  return OD_OK;
}

static int elnes_open_fmt(Text &t, const char *path, od_elnes_info *info) {
  if (!t.load(path)) return io_fail(OD_ERR_IO, "Problem opening elnes_fmt file %s", path);
  skip_lines(t, 2);
  info->norbitals = int_at(t.line(), 21, 5);  // '(a20,1x,i5)'  od2od.f90:662-665
  info->nbands = int_at(t.line(), 21, 5);
  info->nkpoints = int_at(t.line(), 21, 5);
  info->nspins = int_at(t.line(), 21, 5);
  return OD_OK;
}

int od_read_elnes_info(const char *path, char fmt, od_elnes_info *info) {
  if (!path || !info) return io_fail(OD_ERR_ARG, "null argument to od_read_elnes_info%s");
  int rc;
  if (fmt == 't') {
    Text t;
    rc = elnes_open_fmt(t, path, info);
  } else if (fmt == 'b' || fmt == 'l') {
    FortranFile f;
    rc = elnes_open_bin(f, path, fmt, info);
  } else {
    return io_fail(OD_ERR_ARG, "od_read_elnes_info: fmt must be 'b', 'l' or 't'%s");
  }
  if (rc) return rc;
  if (info->nkpoints < 1 || info->nspins < 1 || info->norbitals < 1 || info->nbands < 1) return io_fail(OD_ERR_IO, "%s: bad dimensions", path);
  return OD_OK;
}

int od_read_elnes(const char *path, char fmt, const od_elnes_info *info, int *species_no, int *rank_in_species, int *shell,
                  int *am_channel, double *elnes_mat) {
  if (!path || !info || !elnes_mat) return io_fail(OD_ERR_ARG, "null argument to od_read_elnes%s");
  const int nk = info->nkpoints, ns = info->nspins, norb = info->norbitals, nb = info->nbands;
  const size_t per = (size_t)norb * nb * 3;  // complex numbers per (ik, is): (orb, ib, indx), orb fastest
  int *dst_meta[4] = {species_no, rank_in_species, shell, am_channel};
  od_elnes_info chk;
  if (fmt == 't') {
    Text t;
    int rc = elnes_open_fmt(t, path, &chk);
    if (rc) return rc;
    for (int m = 0; m < 4; ++m) {  // '(a10, norb(1x,i5))'
      std::string l = t.line();
      std::vector<double> v;
      numbers_in(l.size() > 10 ? l.substr(10) : std::string(), &v);
      if ((int)v.size() < norb) return io_fail(OD_ERR_IO, "%s: truncated orbital table", path);
      if (dst_meta[m]) for (int o = 0; o < norb; ++o) dst_meta[m][o] = (int)v[o];
    }
    for (size_t blk = 0; blk < (size_t)nk * ns; ++blk) {
      const size_t ik = blk / ns, is = blk % ns;
      double *dst = elnes_mat + 2 * per * (ik + (size_t)nk * is);
      for (size_t e = 0; e < 2 * per; ++e)
        if (!t.number(&dst[e])) return io_fail(OD_ERR_IO, "%s: truncated (block %lld)", path, (long long)blk);
    }
    return OD_OK;
  }
  if (fmt != 'b' && fmt != 'l') return io_fail(OD_ERR_ARG, "od_read_elnes: fmt must be 'b', 'l' or 't'%s");
  FortranFile f;
  int rc = elnes_open_bin(f, path, fmt, &chk);
  if (rc) return rc;
  if (chk.nkpoints != nk || chk.nspins != ns || chk.norbitals != norb || chk.nbands != nb) return io_fail(OD_ERR_ARG, "%s: dimensions differ from info", path);
  std::vector<int32_t> meta((size_t)norb);
  for (int m = 0; m < 4; ++m) {
    if (!f.record(meta.data(), 4 * (size_t)norb, 4)) return io_fail(OD_ERR_IO, "%s: cannot read the orbital table", path);
    if (dst_meta[m]) for (int o = 0; o < norb; ++o) dst_meta[m][o] = meta[o];
  }
  for (size_t ik = 0; ik < (size_t)nk; ++ik)
    for (size_t is = 0; is < (size_t)ns; ++is) {
      double *dst = elnes_mat + 2 * per * (ik + (size_t)nk * is);
      if (fmt == 'b') {
        if (!f.record(dst, 16 * per, 8)) return io_fail(OD_ERR_IO, "%s: short record (k-point %lld)", path, (long long)ik + 1);
      } else {  // legacy: one record of 3 complex numbers per (orb, band), orb outer (electronic.f90:912-920)
        double c[6];
        for (int orb = 0; orb < norb; ++orb)
          for (int ib = 0; ib < nb; ++ib) {
            if (!f.record(c, 48, 8)) return io_fail(OD_ERR_IO, "%s: short record (k-point %lld)", path, (long long)ik + 1);
            for (int x = 0; x < 3; ++x) {
              const size_t e = orb + (size_t)norb * (ib + (size_t)nb * x);
              dst[2 * e] = c[2 * x];
              dst[2 * e + 1] = c[2 * x + 1];
            }
          }
      }
    }
  return OD_OK;
}

}  // extern "C"

/*
 * od_oracle.h -- CPU restatement of OptaDOS's Brillouin-zone spectral integration
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle of optados-b200: a plain-C, double precision,
 * loop-for-loop restatement of the reference algorithm (reference =
 * optados-developers/optados, Fortran, cited below as <file>:<line> relative to
 * optados/src/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (liboptados_b200.so) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED.  The oracle is checked against the reference's own
 * committed golden outputs (test-suite/tests/<t>/benchmark.out.default.inp=*.odi) by
 * tests/test_oracle_golden.py using fixtures extracted by tests/golden/make_golden.py.
 *
 * Array layouts are the reference's (Fortran column-major, first index fastest):
 *   band_energy(nb,ns,nk)            electronic.f90:38
 *   band_gradient(nb,3,nk,ns)        electronic.f90:39
 *   kpoint_weight(nk)                cell.f90:43
 *   recip_lattice(3,3)               cell.f90:924-980  [i + 3*j]
 *   matrix_weights(norb,nb,nk,ns)    dos_utils.f90:131
 *   dos(B,ns), intdos(B,ns), weighted_dos(B,ns,norb)   dos_utils.f90:1188,1185
 *   jdos matrix_weights(nb,nb,nk,ns,N_geom)            jdos_utils.f90:79
 *   optical_mat(nb,nb,3,nk,ns) complex                 electronic.f90:40
 *   elnes_mat(norb,nb,3,nk,ns) complex                 electronic.f90:41
 *   crystal_symmetry_operations(3,3,nsym)              cell.f90:675-682
 */
#ifndef OD_ORACLE_H
#define OD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* constants.f90:40-47 */
#define ODO_PI 3.141592653589793238462643383279502884197
#define ODO_H2EV 27.21138342902473
#define ODO_BOHR2ANG 0.52917720859
#define ODO_INV_SQRT_TWO_PI 0.3989422804014326779399460599
#define ODO_SQRT_TWO 1.414213562373095048801688724209698079

typedef struct odo_system {
  int nk, ns, nb;
  const double *band_energy;   /* (nb,ns,nk) eV */
  const double *band_gradient; /* (nb,3,nk,ns) eV Ang; may be NULL for fixed */
  const double *kpoint_weight; /* (nk) */
  double recip_lattice[9];     /* (3,3) column-major, 1/Ang incl. 2 pi */
  int kpoint_grid_dim[3];
  double electrons_per_state; /* electronic.f90:605-611 */
} odo_system;

typedef struct odo_broadening {
  double adaptive_smearing;     /* parameters.f90:246 default 0.4 */
  double fixed_smearing;        /* parameters.f90:249 default 0.3 */
  double linear_smearing;       /* parameters.f90:252 default 0.0 */
  int finite_bin_correction;    /* parameters.f90:275 default T */
  int numerical_intdos;         /* parameters.f90:280 default F */
  int hybrid_linear;            /* parameters.f90:286 default F */
  double hybrid_linear_grad_tol; /* parameters.f90:289 default 0.01 */
} odo_broadening;

/* algorithms.f90:73 */
double odo_gaussian(double m, double w, double x);
/* algorithms.f90:252 */
double odo_erf(double x);
/* algorithms.f90:94 (descending order) */
void odo_heap_sort(int num_items, double *weight);
/* algorithms.f90:309; returns 0 ok, 1 if fewer elements than nodes */
int odo_dist_array(int num_elements, int num_nodes, int *elements_per_node);

/* dos_utils.f90:1320 */
void odo_doslin_sub_cell_corners(const double grad[3], const double step[3], double energy,
                                 const double recip_lattice[9], double EV[5]);
/* dos_utils.f90:1391 ; returns dos, *cuml = integrated; *err set to 1 on the two `stop`s */
double odo_doslin(double e0, double e1, double e2, double e3, double e4, double e, double *cuml,
                  int *err);

/* dos_utils.f90:927  (min/max given so the caller can do the "reduce").
 * dos_nbins is in/out (sticky, Q4).  Writes E[0..nbins) if E != NULL. */
void odo_dos_energy_scale(double min_band_energy_in, double max_band_energy_in, double dos_min_energy,
                          double dos_max_energy, int have_min, int have_max, double dos_spacing,
                          int *dos_nbins, double *emin_out, double *delta_bins_out);
void odo_minmax(const double *a, long n, double *mn, double *mx);

/* dos_utils.f90:1140.  dos_type in {'f','a','l'}.  mw may be NULL.  Outputs are zeroed here
 * (allocate_dos).  Returns 0, or nonzero on doslin `stop`. */
int odo_calculate_dos(char dos_type, const odo_system *sys, const odo_broadening *br, const double *E,
                      int nbins, double delta_bins, const double *matrix_weights, int mw_norb,
                      int mw_nb, int mw_nk, double *dos, double *intdos, double *weighted_dos);

/* Same loop, k-points block-partitioned over `nthreads` host threads exactly as
 * algor_dist_array would over MPI ranks, private partial spectra, final sum
 * (dos_utils_merge -> comms_reduce, comms.F90:556).  Used as the timed CPU baseline. */
int odo_calculate_dos_mt(int nthreads, char dos_type, const odo_system *sys, const odo_broadening *br,
                         const double *E, int nbins, double delta_bins, const double *matrix_weights,
                         int mw_norb, int mw_nb, int mw_nk, double *dos, double *intdos,
                         double *weighted_dos);

/* dos_utils.f90:1651.  dos_at_e(ns); weighted_dos_at_e(ns,norb) may be NULL */
int odo_calculate_dos_at_e(char dos_type, const odo_system *sys, const odo_broadening *br,
                           double delta_bins, double energy, const double *matrix_weights, int mw_norb,
                           int mw_nb, int mw_nk, double *dos_at_e, double *weighted_dos_at_e);

/* dos_utils.f90:798 ; num_electrons(ns) */
double odo_calc_efermi_from_intdos(const double *intdos, const double *E, int nbins, int ns,
                                   const double *num_electrons);
/* dos_utils.f90:326-365 (efermi : insulator) */
double odo_efermi_insulator(const odo_system *sys, const double *num_electrons);

/* jdos_utils.f90:177 ; jdos_max_energy in/out (<0: derive from efermi - max_band_energy) */
void odo_jdos_energy_scale(double max_band_energy, double efermi, double jdos_spacing,
                           double *jdos_max_energy, int *jdos_nbins, double *delta_bins);

/* jdos_utils.f90:280.  matrix_weights(nb,nb,nk,ns,N_geom) may be NULL. */
int odo_calculate_jdos(char jdos_type, const odo_system *sys, const odo_broadening *br, const double *E,
                       int nbins, double delta_bins, double efermi, double scissor_op,
                       const int *exclude_bands, int num_exclude_bands, const double *matrix_weights,
                       int n_geom, double *jdos, double *weighted_jdos);
int odo_calculate_jdos_mt(int nthreads, char jdos_type, const odo_system *sys, const odo_broadening *br,
                          const double *E, int nbins, double delta_bins, double efermi,
                          double scissor_op, const int *exclude_bands, int num_exclude_bands,
                          const double *matrix_weights, int n_geom, double *jdos,
                          double *weighted_jdos);

/* optics.f90:157.  geom: 0 polycrys, 1 polar, 2 unpolar, 3 tensor.
 * optical_mat complex (re,im) pairs.  out matrix_weights(nb,nb,nk,ns,N_geom), zero-filled here. */
enum { ODO_GEOM_POLY = 0, ODO_GEOM_POLAR = 1, ODO_GEOM_UNPOLAR = 2, ODO_GEOM_TENSOR = 3 };
int odo_make_weights(const odo_system *sys, const double *optical_mat, int geom, const double qdir_in[3],
                     const double *symops, int num_symm, double efermi, double scissor_op,
                     double *matrix_weights);

/* optics.f90:495 (interband part): epsilon2(B, N_geom) */
void odo_calc_epsilon_2(const double *weighted_jdos, int nbins, int ns, int n_geom, double cell_volume,
                        double *epsilon2);
/* optics.f90:569 (no intraband): epsilon1(B, N_geom) */
void odo_calc_epsilon_1(const double *epsilon2, const double *E, int nbins, int n_geom,
                        double *epsilon1);
void odo_calc_loss_fn(const double *epsilon1, const double *epsilon2, const double *E, int nbins, int broadening,
                      double lossfn_gaussian, double *loss_fn, double *loss_fn_broadened);
void odo_optics_sum_rules(const double *E, const double *epsilon2, const double *loss_fn, int nbins, double cell_volume,
                          double *N_eff, double *N_eff2, double *N_eff3);
void odo_calc_optics_derived(const double *E, const double *epsilon1, const double *epsilon2, int nbins, double *conduct,
                             double *refract, double *absorp, double *reflect);
void odo_core_lorentzian(const double *wdos, const double *E, int nbins, int ncols, double efermi, double lorentzian_width,
                         double lorentzian_scale, double lorentzian_offset, double *out);
void odo_core_gaussian(const double *in, const double *E, int nbins, int ncols, double gaussian_width, double *out);

/* core.f90:84.  core_type: 0 absorption, 1 emission, 2 all.  polar: use qdir.  out (norb,nb,nk,ns).
 * returns 1 if num_sym > 1 (reference io_error) */
int odo_core_prepare_matrix_elements(const odo_system *sys, const double *elnes_mat, int norb,
                                     int polar, const double qdir_in[3], const double *symops,
                                     int num_sym, int core_type, double efermi,
                                     double *matrix_weights);

/* projection_utils.f90:64 with the 4-index 0/1 lookup flattened by the caller into
 * mask(norb_in, nproj) = projection_array(species(orb), rank(orb), am(orb)+1, proj).
 * pdos_weights(norb_in, nb, nk, ns) -> matrix_weights(nproj, nb, nk, ns) */
void odo_projection_merge(const double *pdos_weights, const int *mask, int norb_in, int nproj, int nb,
                          int nk, int ns, double *matrix_weights);

/* electronic.f90:280-291 : band_gradient(ib,:,ik,is) = Re(optical_mat(ib,ib,:,ik,is)) */
void odo_gradient_from_ome(const double *optical_mat, int nb, int nk, int ns, double *band_gradient);

/* dos_utils_compute_bandgap (dos_utils.f90:457-750): root's analysis of the (vbm, cbm) slice of every (k, spin).
 * k indices are 1-based like the reference's printout; -1 = never set. */
typedef struct odo_bandgap {
  double thermal_vbm, thermal_cbm, thermal_bandgap; /* vbm_energy, cbm_energy (:651-653) */
  int thermal_vbm_k, thermal_cbm_k;
  int vbm_multiplicity[2], cbm_multiplicity[2], optical_multiplicity[2];
  double optical_bandgap[2], average_bandgap[2];
  double weighted_average;  /* :713-715, nspins > 1 only */
  int thermal_multiplicity; /* :664-667 */
} odo_bandgap;
int odo_compute_bandgap(const odo_system *sys, double efermi, const double *num_electrons, odo_bandgap *out);
/* calc_band_energies (dos_utils.f90:753-795) and the "From CASTEP" sum of dos_utils_compute_band_energies (:900-908) */
double odo_calc_band_energies(const double *dos, const double *E, int nbins, int ns, double delta_bins, double efermi);
double odo_band_energy_from_bands(const odo_system *sys, double efermi);

/* the intraband (Drude) branches of optics.f90 (calc_epsilon_2 / _1, calc_loss_fn, sum rules, calc_conduct, calc_refract,
 * calc_absorp, calc_reflect); see od_oracle.c.  Unpinned: the Al4 fixtures of the reference's suite are incomplete. */
void odo_optics_intraband(const double *wjdos, const double *weighted_dos_at_e, const double *E, int nbins, int ns, int n_geom,
                          double cell_volume, double drude_broadening, int lossfn_broadening, double lossfn_gaussian, int derived,
                          double *epsilon, double *loss_fn, double *conduct, double *refract, double *absorp, double *reflect,
                          double *N_eff, double *N_eff2, double *N_eff3);

#ifdef __cplusplus
}
#endif
#endif

/*
 * optados_b200.h -- C ABI of liboptados_b200.so
 *
 * B200-native (sm_100a, FP64 CUDA) replacement for the Brillouin-zone spectral-integration hot
 * path of OptaDOS.  The reference (optados-developers/optados, Fortran) has no plugin / FFI API:
 * its tasks call module procedures that read module-global arrays.  This header is the boundary
 * a maintainer binds with ISO_C_BINDING (see INTEGRATION.md and optados_b200/fortran/): every
 * entry point below names the reference routine whose body it replaces (paths relative to
 * optados/src/ in the reference tree).
 *
 * Conventions
 *   - all functions return int: 0 = OK, non-zero = error; od_last_error(ctx) gives the message the
 *     Fortran shim forwards to io_error (io.f90:91).  No exceptions cross the ABI.
 *   - arrays are the reference's: IEEE double, Fortran column-major, first index fastest; complex
 *     arrays are (re,im) pairs.  Layouts are repeated at each entry point.
 *   - the caller owns every array it passes; `mem` says where an input lives (OD_MEM_HOST: the
 *     library copies it to HBM; OD_MEM_DEVICE: a device pointer the library uses in place and
 *     that must stay valid until the next od_set_* of the same array or od_ctx_destroy).
 *     Output arrays are always host memory unless the name says _device.
 *   - OD_MEM_HOST inputs are copied with cudaMemcpyAsync on the ctx stream.  From PAGEABLE memory (an ordinary
 *     Fortran array) the source has been read when the call returns; from PINNED memory the copy may still be in
 *     flight, so a pinned input must stay valid and untouched until the next blocking call on the ctx (any
 *     calculate / get call, or od_synchronize).
 *   - no global state: everything lives in the ctx, with one exception -- the table of NCCL entry points bound with
 *     dlopen on first use (process-wide, read-only afterwards).
 *   - calls are blocking and not re-entrant per ctx.  There is no CPU fallback: without a CUDA
 *     device od_ctx_create fails.
 *   - one ctx = one GPU = one k-point shard (the reference's "node", algorithms.f90:309).  With
 *     od_ctx_create_rank the ctx also owns an NCCL communicator and the *_merge / od_allreduce_*
 *     calls combine the shards (replacing comms_reduce, comms.F90:556); results are then valid on
 *     every rank (allreduce, a superset of the reference's reduce-to-root).
 */
#ifndef OPTADOS_B200_H
#define OPTADOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct od_ctx od_ctx;
typedef struct od_broadening od_broadening; /* defined below */

/* OD_MEM_HOST_DEFERRED (od_set_gradients only): the library keeps the HOST pointer (pinned memory advised; it must
 * stay valid until the next od_set_gradients) and uploads the array slab by slab INSIDE calculate_dos /
 * dos_utils_calculate, overlapped with the kernels working on the previous slab */
enum { OD_MEM_HOST = 0, OD_MEM_DEVICE = 1, OD_MEM_HOST_DEFERRED = 2 };
enum { OD_OK = 0, OD_ERR_ARG = 1, OD_ERR_CUDA = 2, OD_ERR_STATE = 3, OD_ERR_NCCL = 4, OD_ERR_DOSLIN = 5, OD_ERR_REF = 6, OD_ERR_IO = 7 };

/* optics_geom (parameters.f90:346; optics.f90:213-244) and core_type (parameters.f90:373-378) */
enum { OD_GEOM_POLYCRYS = 0, OD_GEOM_POLAR = 1, OD_GEOM_UNPOLAR = 2, OD_GEOM_TENSOR = 3 };
enum { OD_CORE_ABSORPTION = 0, OD_CORE_EMISSION = 1, OD_CORE_ALL = 2 };
/* efermi_choice (parameters.f90:257; dos_utils.f90:317-380) */
enum { OD_EFERMI_OPTADOS = 0, OD_EFERMI_FILE = 1, OD_EFERMI_USER = 2, OD_EFERMI_INSULATOR = 3 };

/* ---------------------------------------------------------------- context ------------------ */
const char *od_version(void);
/* sizeof(struct <name>) for the public structures of this header ("od_dos_params", ...), -1 for an unknown name: lets a
 * binding in another language (the Fortran shim's derived types, ctypes mirrors) verify its layout at start-up */
int od_abi_sizeof(const char *struct_name);
/* single GPU, no communicator (the reference's serial build: comms.F90 stubs) */
int od_ctx_create(int device, od_ctx **ctx);
/* one rank per GPU (the reference's MPI build: comms_setup, comms.F90:85).  id128 is the 128-byte
 * NCCL unique id produced by od_comm_get_unique_id on rank 0 and broadcast by the host program. */
int od_comm_get_unique_id(void *id128);
int od_ctx_create_rank(int device, int rank, int nranks, const void *id128, od_ctx **ctx);
int od_ctx_destroy(od_ctx *ctx);
const char *od_last_error(const od_ctx *ctx);
int od_ctx_rank(const od_ctx *ctx);
int od_ctx_nranks(const od_ctx *ctx);
int od_synchronize(od_ctx *ctx);
/* engine selection for calculate_dos: OD_ENGINE_AUTO (narrow-window engine with the dense engine for
 * wide windows) or OD_ENGINE_DENSE (dense bin-owner engine only: the reference's arithmetic, slow) */
enum { OD_ENGINE_AUTO = 0, OD_ENGINE_DENSE = 1 };
int od_set_engine(od_ctx *ctx, int engine);
/* number of CUDA kernels this ctx has launched so far (bench.py's gpu_launches) */
long long od_kernel_launches(const od_ctx *ctx);
/* device-side milliseconds (CUDA events on the ctx stream) spent in the accumulation kernels and in the
 * prepare (classify + sort) passes since the last reset; the in-window (record, bin) pairs the narrow
 * engine evaluated (contributions) and the lane-steps it spent on them (lane_steps >= contributions:
 * the ratio is the lane efficiency of the record groups) */
int od_profile_reset(od_ctx *ctx);
int od_profile_get(od_ctx *ctx, double *accumulate_ms, double *prepare_ms, double *contributions, double *lane_steps,
                   long long *accumulate_launches);

/* device-side stopwatch on the ctx stream (CUDA events): start, ... library calls ..., stop -> ms */
int od_timer_start(od_ctx *ctx);
int od_timer_stop(od_ctx *ctx, double *elapsed_ms);
/* sustained FP64 throughput of this GPU's FP64 pipe (all SMs, about `seconds` of dependent-chain-free DFMA and DMMA
 * loops, the larger of the two), in TFLOP/s (2 flops per FMA) -- the denominator of the FP64 roofline, measured in
 * place because MEASURED_PEAKS.json carries no FP64 entry */
int od_measure_fp64_peak(od_ctx *ctx, double seconds, double *tflops);
/* mode 0: the larger of a DFMA loop and a DMMA (mma.sync m16n8k8 f64) loop = the FP64 pipe peak (what
 * od_measure_fp64_peak returns); 1: the DFMA loop alone; 2: the DMMA loop alone */
int od_measure_fp64_peak_mode(od_ctx *ctx, double seconds, int mode, double *tflops);
/* in-window (unit, bin) contributions of a calculate_dos pass on this rank, the unit of the
 * throughput metric: gaussian W = floor(2 sqrt(60) w / dE) + 1,
 * linear W = floor(2 (e0 - e4) / dE) + 1, each clipped to the grid */
int od_count_contributions(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins,
                           const od_broadening *br, double *contributions);
/* pinned host memory for the end-to-end path (cudaHostAlloc) */
int od_host_alloc_pinned(size_t nbytes, void **hptr);
int od_host_free_pinned(void *hptr);
/* page-lock / release host memory the caller owns (cudaHostRegister): a Fortran array handed over with OD_MEM_HOST or
 * OD_MEM_HOST_DEFERRED is then copied at the PCIe rate, asynchronously, instead of through a pageable bounce buffer */
int od_host_register(void *hptr, size_t nbytes);
int od_host_unregister(void *hptr);

/* device memory helpers (so weight arrays can stay resident between calls) */
int od_device_malloc(od_ctx *ctx, size_t nbytes, void **dptr);
int od_device_free(od_ctx *ctx, void *dptr);
int od_memcpy_h2d(od_ctx *ctx, void *dst_device, const void *src_host, size_t nbytes);
int od_memcpy_d2h(od_ctx *ctx, void *dst_host, const void *src_device, size_t nbytes);

/* algor_dist_array (algorithms.f90:309): k-points per rank, remainder to the lowest ranks */
int od_dist_array(int num_elements, int num_nodes, int *elements_per_node);

/* comms_reduce_real (comms.F90:556) on a HOST buffer, result valid on every rank. op: 's','m'(in),'M'(ax) */
int od_allreduce(od_ctx *ctx, double *host_buf, size_t n, char op);

/* ---------------------------------------------------------------- module data --------------- */
/* od_cell / od_electronic data the hot loops read (cell.f90:43-56, electronic.f90:38-41,605-611):
 *   recip_lattice(3,3) [i+3*j], kpoint_grid_dim(3), kpoint_weight(nk_local), electrons_per_state */
int od_set_system(od_ctx *ctx, int nk_local, int ns, int nb, const double recip_lattice[9],
                  const int kpoint_grid_dim[3], const double *kpoint_weight, double electrons_per_state);
/* band_energy(nb,ns,nk_local) in eV (electronic.f90:38, filled :554-601) */
int od_set_bands(od_ctx *ctx, const double *band_energy, int mem);
/* band_gradient(nb,3,nk_local,ns) in eV Ang (electronic.f90:39; elec_read_band_gradient :170-299) */
int od_set_gradients(od_ctx *ctx, const double *band_gradient, int mem);
/* the "diag of OME" branch of elec_read_band_gradient (electronic.f90:276-291):
 * band_gradient(ib,:,ik,is) = Re(optical_mat(ib,ib,:,ik,is)); optical_mat(nb,nb,3,nk,ns) complex */
int od_set_gradients_from_ome(od_ctx *ctx, const double *optical_mat, int mem);
/* copy the resident gradients back (tests) */
int od_get_gradients(od_ctx *ctx, double *band_gradient_host);
/* minval / maxval(band_energy) over this rank, then MIN/MAX-allreduced if the ctx has a communicator
 * (dos_utils.f90:964-977, jdos_utils.f90:194-196) */
int od_band_minmax(od_ctx *ctx, double *min_band_energy, double *max_band_energy);

/* broadening flags read by the hot loops (parameters.f90:246-289) */
typedef struct od_broadening {
  double adaptive_smearing;      /* default 0.4 */
  double fixed_smearing;         /* default 0.3 */
  double linear_smearing;        /* default 0.0 */
  int finite_bin_correction;     /* default 1 */
  int numerical_intdos;          /* default 0 */
  int hybrid_linear;             /* default 0 */
  double hybrid_linear_grad_tol; /* default 0.01 */
} od_broadening;
void od_broadening_defaults(od_broadening *b);

/* ---------------------------------------------------------------- hot loops ------------------ */
/* calculate_dos (dos_utils.f90:1140-1317): rank-local partial spectra on the grid
 * E(i) = emin + (i-1)*delta_bins, i = 1..nbins (dos_utils.f90:994-997).
 *   dos_type 'f' | 'a' | 'l'
 *   matrix_weights(mw_norb, mw_nb, mw_nk, ns) or NULL (dos_utils.f90:1267-1276)
 *   out: dos(nbins,ns), intdos(nbins,ns), weighted_dos(nbins,ns,mw_norb) (NULL if no weights)
 * With merge != 0 the partial spectra are sum-allreduced over the communicator on the device before
 * they are copied out (= calculate_dos followed by dos_utils_merge, dos_utils.f90:1013-1053). */
int od_calculate_dos(od_ctx *ctx, char dos_type, double emin, double delta_bins, int nbins,
                     const od_broadening *br, const double *matrix_weights, int mw_norb, int mw_nb,
                     int mw_nk, int mw_mem, int merge, double *dos, double *intdos, double *weighted_dos);

/* calculate_dos_at_e (dos_utils.f90:1651-1800) [+ dos_utils_merge_at_e :1803 when merge != 0]
 *   out: dos_at_e(ns), weighted_dos_at_e(ns, mw_norb) or NULL */
int od_calculate_dos_at_e(od_ctx *ctx, char dos_type, double energy, double delta_bins,
                          const od_broadening *br, const double *matrix_weights, int mw_norb, int mw_nb,
                          int mw_nk, int mw_mem, int merge, double *dos_at_e, double *weighted_dos_at_e);

/* calculate_jdos (jdos_utils.f90:280-404) on E(i) = (i-1)*delta_bins [+ jdos_utils_merge :407].
 *   matrix_weights(nb,nb,nk,ns,n_geom) or NULL;  exclude_bands are 1-based band indices
 *   out: jdos(nbins,ns), weighted_jdos(nbins,ns,n_geom) or NULL
 * NB merge reduces the whole weighted_jdos, not only its first N_geom slice (jdos_utils.f90:437 is
 * an MPI bug for optics_geom = tensor; the serial reference is unaffected). */
int od_calculate_jdos(od_ctx *ctx, char jdos_type, double delta_bins, int nbins, const od_broadening *br,
                      double efermi, double scissor_op, const int *exclude_bands, int num_exclude_bands,
                      const double *matrix_weights, int n_geom, int mw_mem, int merge, double *jdos,
                      double *weighted_jdos);

/* calculate_jdos with make_weights (optics.f90:157-435) fused in: the optical weights are formed on
 * chip from optical_mat(nb,nb,3,nk,ns) (complex, eV Ang) and matrix_weights(nb,nb,nk,ns,N_geom) is
 * never materialised.  symops(3,3,num_symm) = crystal_symmetry_operations (cell.f90:675-682). */
int od_calculate_jdos_optics(od_ctx *ctx, char jdos_type, double delta_bins, int nbins,
                             const od_broadening *br, double efermi, double scissor_op,
                             const int *exclude_bands, int num_exclude_bands, const double *optical_mat,
                             int ome_mem, int geom, const double qdir[3], const double *symops, int num_symm,
                             int merge, double *jdos, double *weighted_jdos);

/* ---------------------------------------------------------------- weight builders ------------ */
/* make_weights (optics.f90:157-435): out matrix_weights(nb,nb,nk,ns,N_geom), N_geom = 6 for tensor
 * else 1.  out_mem selects a host or (caller-allocated) device output. */
int od_make_weights(od_ctx *ctx, const double *optical_mat, int ome_mem, int geom, const double qdir[3],
                    const double *symops, int num_symm, double efermi, double scissor_op,
                    double *matrix_weights, int out_mem);
/* core_prepare_matrix_elements (core.f90:84-168): elnes_mat(norb,nb,3,nk,ns) complex ->
 * matrix_weights(norb,nb,nk,ns).  Returns OD_ERR_REF if num_sym > 1 (core.f90:106). */
int od_core_prepare_matrix_elements(od_ctx *ctx, const double *elnes_mat, int elnes_mem, int norb, int polar,
                                    const double qdir[3], const double *symops, int num_sym, int core_type,
                                    double efermi, double *matrix_weights, int out_mem);
/* projection_merge (projection_utils.f90:64-99) with projection_array flattened by the caller to
 * mask(norb_in, nproj) in {0,1}: pdos_weights(norb_in,pw_nb,nk,ns) -> matrix_weights(nproj,pw_nb,nk,ns), pw_nb =
 * pdos_mwab%nbands (the bands the .pdos_bin file holds, <= nbands; pass the result on with mw_nb = pw_nb) */
int od_projection_merge(od_ctx *ctx, const double *pdos_weights, int pw_mem, const int *mask, int norb_in, int pw_nb,
                        int nproj, double *matrix_weights, int out_mem);
/* calc_epsilon_2 interband part (optics.f90:495-551): epsilon2(nbins, n_geom) */
int od_calc_epsilon_2(od_ctx *ctx, const double *weighted_jdos, int nbins, int ns, int n_geom,
                      double cell_volume, double *epsilon2);

/* ---------------------------------------------------------------- task drivers --------------- */
/* Host-side mirror of the reference's L3 drivers: same control flow, flags and side effects, with
 * the module-global results kept inside the ctx and read back with the od_*_get calls. */
typedef struct od_dos_params {
  int fixed, adaptive, linear;   /* broadening switches (parameters.f90:232-234, compare_dos) */
  od_broadening br;
  double dos_spacing;            /* parameters.f90:459-461: default is the SINGLE literal 0.005 */
  int dos_nbins;                 /* in/out: < 0 = derive; sticky afterwards (quirk Q4, dos_utils.f90:980) */
  int have_dos_min_energy, have_dos_max_energy;
  double dos_min_energy, dos_max_energy;
  int efermi_choice;             /* OD_EFERMI_* */
  double efermi_user, efermi_castep;
  int nspins_electrons;          /* entries of num_electrons used */
  double num_electrons[2];       /* electronic.f90: num_electrons(nspins) */
  int dos_per_volume;            /* parameters.f90: dos_per_volume */
  double cell_volume;
} od_dos_params;
void od_dos_params_defaults(od_dos_params *p);

typedef struct od_dos_result {
  int nbins;
  double emin, delta_bins;                             /* E(i) = emin + (i-1)*delta_bins */
  double efermi_fixed, efermi_adaptive, efermi_linear; /* calc_efermi_from_intdos (dos_utils.f90:798) */
  int have_fixed, have_adaptive, have_linear, have_weighted;
  int weighted_norb;
} od_dos_result;

/* dos_utils_calculate (dos_utils.f90:92-293): energy scale, one calculate_dos + merge per enabled
 * broadening (weighted spectrum only with the "most complex" one, :179-208), Fermi analysis. */
int od_dos_utils_calculate(od_ctx *ctx, od_dos_params *p, const double *matrix_weights, int mw_norb,
                           int mw_nb, int mw_nk, int mw_mem, od_dos_result *res);
/* read back dos_<type>, intdos_<type> (nbins,ns) / weighted_dos (nbins,ns,norb) of the last call */
int od_dos_get(od_ctx *ctx, char dos_type, double *dos, double *intdos);
int od_dos_get_weighted(od_ctx *ctx, double *weighted_dos);
/* calc_efermi_from_intdos (dos_utils.f90:798-850) on host arrays */
double od_calc_efermi_from_intdos(const double *intdos, double emin, double delta_bins, int nbins, int ns,
                                  const double *num_electrons);
/* dos_utils_calculate_at_e (dos_utils.f90:1531-1648): dos_at_e(3, ns) = rows (fixed, adaptive, linear) of the enabled
 * schemes (others zero); weighted_dos_at_e(ns, mw_norb) from the most complex enabled scheme; delta_bins is the
 * reference's module variable of the last energy scale (it only enters through finite_bin_correction). */
int od_dos_utils_calculate_at_e(od_ctx *ctx, const od_dos_params *p, double energy, double delta_bins,
                                const double *matrix_weights, int mw_norb, int mw_nb, int mw_nk, int mw_mem,
                                double *dos_at_e, double *weighted_dos_at_e);
/* dos_utils_set_efermi (dos_utils.f90:296-388).  For OD_EFERMI_OPTADOS runs od_dos_utils_calculate. */
int od_dos_utils_set_efermi(od_ctx *ctx, od_dos_params *p, double *efermi);

/* dos_utils_compute_bandgap (dos_utils.f90:457-750).  Every rank passes its block [k_first, k_first + nk_local) of the
 * nk_global k-points (ignored without a communicator); the (vbm, cbm) slices are gathered and scanned in the
 * reference's order, so multiplicities and k-point numbers (1-based, -1 = never set) do not depend on the number of
 * ranks.  num_electrons(nspins) only enters the weighted average.  Also sets the vbm_energy / cbm_energy the
 * core-loss chemical shift uses (core.f90:533-535). */
typedef struct od_bandgap_result {
  double vbm_energy, cbm_energy, thermal_bandgap;   /* <- TBg */
  long long thermal_vbm_k, thermal_cbm_k;           /* equal = direct gap (only meaningful without multiplicity) */
  int vbm_multiplicity[2], cbm_multiplicity[2], optical_multiplicity[2];
  double optical_bandgap[2], average_bandgap[2];    /* <- OBg, <- ABg */
  double weighted_average;                          /* <- wAB (nspins > 1) */
  int thermal_multiplicity;                         /* "Multiple Band Minima and Maxima -- Gap unknown" */
} od_bandgap_result;
int od_dos_utils_compute_bandgap(od_ctx *ctx, double efermi, const double *num_electrons, long long nk_global,
                                 long long k_first, od_bandgap_result *res);
/* dos_utils_compute_band_energies (dos_utils.f90:853-924): band_energy_dos(3) = calc_band_energies (:753-795) of the
 * fixed / adaptive / linear DOS of the last od_dos_utils_calculate (0 where not computed; <- BEF/BEA/BEL),
 * band_energy_castep = sum of the occupied eigenvalues, SUM-reduced over the ranks (<- BEC) */
int od_dos_utils_compute_band_energies(od_ctx *ctx, double efermi, double band_energy_dos[3], double *band_energy_castep);

typedef struct od_jdos_params {
  int fixed, adaptive, linear;
  od_broadening br;
  double jdos_spacing;    /* parameters.f90:321 default 0.01 */
  double jdos_max_energy; /* in/out: < 0 = derive (jdos_utils.f90:193-197) */
  double scissor_op;      /* parameters.f90:351 */
  int num_exclude_bands;
  const int *exclude_bands;
  double efermi;
  int dos_per_volume;
  double cell_volume;
} od_jdos_params;
void od_jdos_params_defaults(od_jdos_params *p);
typedef struct od_jdos_result {
  int nbins;
  double delta_bins; /* E(i) = (i-1)*delta_bins */
  int have_fixed, have_adaptive, have_linear, have_weighted, n_geom;
} od_jdos_result;
/* jdos_utils_calculate (jdos_utils.f90:61-174).  Exactly one of matrix_weights / optical_mat may be
 * non-NULL; with optical_mat the optical weights are fused (see od_calculate_jdos_optics). */
int od_jdos_utils_calculate(od_ctx *ctx, od_jdos_params *p, const double *matrix_weights, int n_geom, int mw_mem,
                            const double *optical_mat, int ome_mem, int geom, const double qdir[3],
                            const double *symops, int num_symm, od_jdos_result *res);
int od_jdos_get(od_ctx *ctx, char jdos_type, double *jdos);
int od_jdos_get_weighted(od_ctx *ctx, double *weighted_jdos);

/* ---------------------------------------------------------------- synthetic inputs ----------- */
/* cosine-band synthetic band structure generated directly in HBM
 * (E_n(k) = c_n + t_n (cos 2pi k1 + cos 2pi k2 + cos 2pi k3), analytic gradients, per-(k,s) sorted).
 * Fills band_energy / band_gradient / kpoint_weight of the ctx for the k-points [k_first,
 * k_first + nk_local) of a grid[0] x grid[1] x grid[2] MP mesh (k3 fastest) and calls od_set_system. */
int od_synth_bands(od_ctx *ctx, const int grid[3], long long k_first, int nk_local, int ns, int nb,
                   double lattice_a, uint64_t seed);
/* the same for the k-points k_first, k_first + k_stride, ... (a cyclic partition over ranks: the synthetic mesh is
 * ordered by k1, so contiguous blocks differ by +-12 % in work at 8 ranks; sums commute, any assignment is valid) */
int od_synth_bands_strided(od_ctx *ctx, const int grid[3], long long k_first, long long k_stride, int nk_local, int ns,
                           int nb, double lattice_a, uint64_t seed);
/* copy the resident arrays of the ctx back to the host (for parity tests against the oracle) */
int od_get_bands(od_ctx *ctx, double *band_energy_host);
int od_get_kweights(od_ctx *ctx, double *kpoint_weight_host);

/* ---------------------------------------------------------------- task mirrors --------------- */
/* two of the reference's tasks end to end, built from the entry points above.
 * optics_calculate (optics.f90:78-155, interband): make_weights fused into the JDOS pass -> epsilon_2 ->
 * Kramers-Kronig epsilon_1 -> (not for optics_geom = tensor) conductivity, refractive index, loss function,
 * absorption, reflection of component 1 (optics.f90:722-825).  jdos.cell_volume is required. */
typedef struct od_optics_params {
  od_jdos_params jdos;    /* broadening switches, grid, efermi, scissor_op, excluded bands, cell_volume */
  int geom;               /* OD_GEOM_* */
  double qdir[3];         /* optics_qdir */
  double lossfn_gaussian; /* optics_lossfn_broadening: FWHM in eV, < 1e-6 = off (parameters.f90:367-371) */
  int intraband;          /* optics_intraband (parameters.f90:361): add the Drude term of the bands crossing the Fermi level */
  double drude_broadening;/* optics_drude_broadening in 1/s, default 1e14 (parameters.f90:364) */
  double dos_delta_bins;  /* intraband only: the od_dos_utils module variable delta_bins that calculate_dos_at_e applies its
                           * finite_bin_correction with (dos_utils.f90:1746) -- the grid spacing of the last DOS calculation of
                           * the run (od_dos_result.delta_bins), 0 if there was none */
} od_optics_params;
void od_optics_params_defaults(od_optics_params *p);
typedef struct od_optics_result {
  int nbins, n_geom;
  double delta_bins;      /* E(i) = (i-1)*delta_bins */
  int have_derived;       /* 0 for optics_geom = tensor */
  int have_loss_fn_broadened;
  double n_eff;           /* sum rule of calc_epsilon_2 (optics.f90:552-563; N_geom = 1 only, else 0) */
  double n_eff2, n_eff3;  /* first / second sum rule of calc_loss_fn (optics.f90:700-719; with have_derived) */
  int n_parts;            /* 1, or 3 with optics_intraband: epsilon(nbins, 2, n_geom, n_parts) = interband, intraband, total */
  int n_loss_fn;          /* columns of loss_fn: 1 (+1 broadened), with intraband 3 (+1 broadened total) (optics.f90:632-644) */
} od_optics_result;
int od_optics_calculate(od_ctx *ctx, od_optics_params *p, const double *optical_mat, int ome_mem, const double *symops,
                        int num_symm, od_optics_result *res);
/* any pointer may be NULL: E(nbins), epsilon(nbins, 2, n_geom, n_parts) with (:,1,:,:) = eps1 and (:,2,:,:) = eps2,
 * conduct(nbins,2), refract(nbins,2), loss_fn(nbins, n_loss_fn), absorp(nbins), reflect(nbins).
 * The intraband branches (optics.f90:517-549, :586-617, :632-697, :736-743, :769-778) are restated with their quirks
 * (the Drude term added once per spin, E e carried by the total epsilon_2, 0/0 at E = 0 in calc_refract). */
int od_optics_get(od_ctx *ctx, double *E, double *epsilon, double *conduct, double *refract, double *loss_fn,
                  double *absorp, double *reflect);

/* core_calculate (core.f90:38-83) + the unit factors of write_core (core.f90:317-323): dos_utils_set_efermi ->
 * core_prepare_matrix_elements -> dos_utils_calculate(matrix_weights, weighted_dos) on the now-sticky grid (Q4) ->
 * LAI broadening.  dos.cell_volume is required.  Edge bookkeeping (which orbitals form an edge) stays with the caller. */
typedef struct od_core_params {
  od_dos_params dos;
  int core_type;          /* OD_CORE_* */
  int polar;              /* core_geom = polarised */
  double qdir[3];         /* core_qdir */
  int lai_broadening;     /* core_LAI_broadening */
  double lai_gaussian_width, lai_lorentzian_width, lai_lorentzian_scale, lai_lorentzian_offset;
  int set_efermi_zero;          /* write_core: E_shift = E - efermi (core.f90:490-494) */
  double core_chemical_shift;   /* Mizoguchi shift; -1 = off (parameters.f90:267).  write_core prints E + shift - cbm_energy
                                 * (core.f90:533-535) with the MODULE variable cbm_energy of od_dos_utils: core_calculate never
                                 * calls dos_utils_compute_bandgap, so it is 0 unless od_dos_utils_compute_bandgap ran on this
                                 * ctx before (a dos task in the same run) -- reproduced as is */
} od_core_params;
void od_core_params_defaults(od_core_params *p);
typedef struct od_core_result {
  int nbins, norb;
  double emin, delta_bins, efermi;
  int have_broadened;
} od_core_result;
int od_core_calculate(od_ctx *ctx, od_core_params *p, const double *elnes_mat, int elnes_mem, int norb, const double *symops,
                      int num_sym, od_core_result *res);
/* E(nbins) as write_core prints it (E_shift: set_efermi_zero, core_chemical_shift), weighted_dos(nbins, ns, norb) and its
 * LAI-broadened copy, both in the units write_core prints */
int od_core_get(od_ctx *ctx, double *E, double *weighted_dos, double *weighted_dos_broadened);

/* ---------------------------------------------------------------- O(B^2) post-processing ------ */
/* the reference's root-only double loops over the energy grid, on the GPU.
 * Host arrays in and out (a few B x N doubles); E(nbins) is the reference's grid array. */
/* calc_epsilon_1 (optics.f90:569-621, interband): Kramers-Kronig transform; epsilon_2, epsilon_1 (nbins, N_geom);
 * components 1-3 get the +1 of the diagonal, 4-6 do not */
int od_calc_epsilon_1(od_ctx *ctx, const double *E, int nbins, int n_geom, const double *epsilon_2, double *epsilon_1);
/* calc_loss_fn (optics.f90:622-720, interband, component 1): loss_fn = -Im(1/(eps1 + i eps2)); if
 * loss_fn_broadened != NULL and optics_lossfn_broadening >= 1e-6 it receives the Gaussian-broadened (FWHM) copy */
int od_calc_loss_fn(od_ctx *ctx, const double *E, int nbins, const double *epsilon_1, const double *epsilon_2,
                    double lossfn_gaussian, double *loss_fn, double *loss_fn_broadened);
/* core_lorentzian + core_gaussian (core.f90:70-76, 655-732): life-time and instrumental broadening of
 * weighted_dos(nbins, ncols = ns * norb) -> weighted_dos_broadened; Lorentzian if width > 1e-14 or scale > 1e-5,
 * Gaussian (applied to the Lorentzian result when both are on) if width > 1e-14 */
int od_core_lai_broadening(od_ctx *ctx, const double *E, int nbins, int ncols, double efermi, double gaussian_width,
                           double lorentzian_width, double lorentzian_scale, double lorentzian_offset,
                           const double *weighted_dos, double *weighted_dos_broadened);

/* ---------------------------------------------------------------- input decoders (host) ------ */
/* SURVEY.md section 8(f) rank 1.  What the reference's Fortran readers produce, from the same files, for host
 * programs without the Fortran front end.  No ctx, no GPU: plain host functions, caller-allocated outputs in the
 * reference's layouts (column-major) and units (eV, eV Ang); a failing call returns non-zero and leaves a message
 * in od_io_last_error() (thread-local).  `fmt`: 'b' = the *_bin files (Fortran sequential unformatted; byte order
 * detected, big-endian under the reference's build flags), 'l' = the legacy files (.cst_ome, .cst_vel,
 * .pdos_weights, .eels_mat), 't' = the formatted *_fmt twins od2od writes (read as the test-suite reads them:
 * fmt -> bin -> read, including the unit round trip of od2od.f90:152/229). */
const char *od_io_last_error(void);

/* <seed>.bands -- elec_read_band_energy, electronic.f90:445-618 */
typedef struct od_bands_info {
  int nkpoints, nspins, nbands;
  double num_electrons[2];
  double efermi_castep;       /* eV (read as f12.4 in Hartree, electronic.f90:511-512) */
  double real_lattice[9];     /* Bohr, column-major: real_lattice(:, i) = lattice-vector line i */
  double electrons_per_state; /* 2 without spin polarisation, else 1 (electronic.f90:604-610) */
} od_bands_info;
int od_read_bands_info(const char *path, od_bands_info *info);
/* band_energy(nb, ns, nk) in eV, kpoint_r(3, nk) (may be NULL), kpoint_weight(nk) (may be NULL) */
int od_read_bands(const char *path, const od_bands_info *info, double *band_energy, double *kpoint_r, double *kpoint_weight);

/* cell_calc_lattice (cell.f90:924-980): Bohr lattice -> real_lattice (Ang, may be NULL), recip_lattice (1/Ang,
 * incl. 2 pi), cell_volume (Ang^3, may be NULL); all 3x3 column-major */
int od_cell_calc_lattice(const double *real_lattice_bohr, double *real_lattice, double *recip_lattice, double *cell_volume);
/* cell_find_MP_grid (cell.f90:87-318): kpoints(3, nk) fractional -> kpoint_grid_dim(3) */
int od_cell_find_mp_grid(const double *kpoints, int nk, int *grid);
/* %block symmetry_ops of <seed>-out.cell (cell.f90:405-696) -> crystal_symmetry_operations(3, 3, n); with
 * ops == NULL only counts */
int od_read_symmetry_ops(const char *cell_path, int max_ops, double *ops, int *n_ops);

/* optical_mat(nb, nb, 3, nk, ns) complex, eV Ang -- elec_read_optical_mat, electronic.f90:302-442 */
int od_read_ome(const char *path, char fmt, int nb, int nk, int ns, double *optical_mat);
/* band_gradient(nb, 3, nk, ns), eV Ang -- elec_read_band_gradient, electronic.f90:170-299 */
int od_read_dome(const char *path, char fmt, int nb, int nk, int ns, double *band_gradient);

/* pdos_weights(norb, nb, nk, ns) + orbital table + nbands_occ(nk, ns) -- elec_pdos_read, electronic.f90:1122-1276 */
typedef struct od_pdos_info { int nkpoints, nspins, norbitals, nbands; } od_pdos_info;
int od_read_pdos_info(const char *path, char fmt, od_pdos_info *info);
int od_read_pdos(const char *path, char fmt, const od_pdos_info *info, int *species_no, int *rank_in_species, int *am_channel,
                 int *nbands_occ, double *pdos_weights);

/* elnes_mat(norb, nb, 3, nk, ns) complex + orbital table -- elec_read_elnes_mat, electronic.f90:811-992 */
typedef struct od_elnes_info { int norbitals, nbands, nkpoints, nspins; } od_elnes_info;
int od_read_elnes_info(const char *path, char fmt, od_elnes_info *info);
int od_read_elnes(const char *path, char fmt, const od_elnes_info *info, int *species_no, int *rank_in_species, int *shell,
                  int *am_channel, double *elnes_mat);

/* ---------------------------------------------------------------- the file contract (host) ---- */
/* SURVEY.md section 8(f) rank 4.  od_param_read = param_read (parameters.f90:138-465) with its tokenizer (:837-1470):
 * lower-cased lines, `!` / `#` comments, keyword at column 1 followed by `=`, `:` or a blank, every line must be
 * consumed ("Unrecognised keyword(s) in input file").  Defaults as the reference sets them. */
typedef struct od_odi {
  int iprint, legacy_file_format;
  int dos, pdos, pdis, jdos, optics, core, compare_dos, compare_jdos;      /* task (parameters.f90:164-197) */
  int fixed, adaptive, linear;                                             /* broadening (:205-224) */
  double adaptive_smearing, fixed_smearing, linear_smearing;
  int efermi_choice;                                                       /* OD_EFERMI_* (:254-259, param_get_efermi) */
  double efermi_user;
  int compute_band_energy, compute_band_gap;
  double core_chemical_shift;                                              /* -1 = off */
  int finite_bin_correction, numerical_intdos, hybrid_linear;
  double hybrid_linear_grad_tol;
  int set_efermi_zero, dos_per_volume;
  int have_dos_min_energy, have_dos_max_energy;
  double dos_min_energy, dos_max_energy, dos_spacing;
  int dos_nbins;
  char projectors_string[256];                                             /* pdos : ... */
  double jdos_max_energy, jdos_spacing;
  int num_exclude_bands;
  int exclude_bands[256];
  char optics_geom[32];
  double scissor_op, optics_qdir[3];
  int optics_intraband;
  double optics_drude_broadening;
  int optics_lossfn_broadening;
  double optics_lossfn_gaussian;
  char core_geom[32], core_type[32];
  double core_qdir[3];
  int core_lai_broadening;
  double lai_gaussian_width, lai_lorentzian_width, lai_lorentzian_scale, lai_lorentzian_offset;
} od_odi;
void od_odi_defaults(od_odi *p);
int od_param_read(const char *odi_path, od_odi *p);
/* message of the last failing od_param_read / od_run_seed of this thread */
const char *od_run_last_error(void);
/* <dir>/<seed>.odi + the CASTEP outputs it asks for (<seed>.bands; dome / ome / pdos / elnes files as *_bin, legacy or
 * formatted *_fmt; <seed>-out.cell) -> <seed>.odo and the reference's .dat files in <dir>, tasks in the reference's order
 * (optados.f90:120-222) on the GPU of ctx.  pdos_write's projector headers, write_dos / write_jdos E21.13 tables,
 * write_epsilon / write_loss_fn / ... and write_core blocks as the reference lays them out; .agr files are not written. */
int od_run_seed(od_ctx *ctx, const char *dir, const char *seedname);

#ifdef __cplusplus
}
#endif
#endif /* OPTADOS_B200_H */

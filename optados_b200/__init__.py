"""optados_b200 -- thin ctypes binding of liboptados_b200.so (the C ABI in include/optados_b200.h).

The product is the CUDA library; this module only loads it and marshals numpy arrays, so that the
parity tests and bench.py can call the same entry points a Fortran host binds with ISO_C_BINDING
(optados_b200/fortran/od_b200.f90).  There is no CPU fallback: if the library is missing or no
CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OD_B200_LIB") or os.path.join(_HERE, "lib", "liboptados_b200.so")   # (override: kernel-variant experiments only)

OD_MEM_HOST, OD_MEM_DEVICE, OD_MEM_HOST_DEFERRED = 0, 1, 2
GEOM = {"polycrys": 0, "polar": 1, "unpolar": 2, "tensor": 3}
CORE_TYPE = {"absorption": 0, "emission": 1, "all": 2}
EFERMI = {"optados": 0, "file": 1, "user": 2, "insulator": 3}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class Broadening(C.Structure):
    _fields_ = [("adaptive_smearing", C.c_double), ("fixed_smearing", C.c_double), ("linear_smearing", C.c_double),
                ("finite_bin_correction", C.c_int), ("numerical_intdos", C.c_int), ("hybrid_linear", C.c_int),
                ("hybrid_linear_grad_tol", C.c_double)]


class DosParams(C.Structure):
    _fields_ = [("fixed", C.c_int), ("adaptive", C.c_int), ("linear", C.c_int), ("br", Broadening),
                ("dos_spacing", C.c_double), ("dos_nbins", C.c_int),
                ("have_dos_min_energy", C.c_int), ("have_dos_max_energy", C.c_int),
                ("dos_min_energy", C.c_double), ("dos_max_energy", C.c_double),
                ("efermi_choice", C.c_int), ("efermi_user", C.c_double), ("efermi_castep", C.c_double),
                ("nspins_electrons", C.c_int), ("num_electrons", C.c_double * 2),
                ("dos_per_volume", C.c_int), ("cell_volume", C.c_double)]


class DosResult(C.Structure):
    _fields_ = [("nbins", C.c_int), ("emin", C.c_double), ("delta_bins", C.c_double),
                ("efermi_fixed", C.c_double), ("efermi_adaptive", C.c_double), ("efermi_linear", C.c_double),
                ("have_fixed", C.c_int), ("have_adaptive", C.c_int), ("have_linear", C.c_int),
                ("have_weighted", C.c_int), ("weighted_norb", C.c_int)]


class JdosParams(C.Structure):
    _fields_ = [("fixed", C.c_int), ("adaptive", C.c_int), ("linear", C.c_int), ("br", Broadening),
                ("jdos_spacing", C.c_double), ("jdos_max_energy", C.c_double), ("scissor_op", C.c_double),
                ("num_exclude_bands", C.c_int), ("exclude_bands", c_ip), ("efermi", C.c_double),
                ("dos_per_volume", C.c_int), ("cell_volume", C.c_double)]


class JdosResult(C.Structure):
    _fields_ = [("nbins", C.c_int), ("delta_bins", C.c_double), ("have_fixed", C.c_int), ("have_adaptive", C.c_int),
                ("have_linear", C.c_int), ("have_weighted", C.c_int), ("n_geom", C.c_int)]


class OpticsParams(C.Structure):
    _fields_ = [("jdos", JdosParams), ("geom", C.c_int), ("qdir", C.c_double * 3), ("lossfn_gaussian", C.c_double),
                ("intraband", C.c_int), ("drude_broadening", C.c_double), ("dos_delta_bins", C.c_double)]


class OpticsResult(C.Structure):
    _fields_ = [("nbins", C.c_int), ("n_geom", C.c_int), ("delta_bins", C.c_double), ("have_derived", C.c_int),
                ("have_loss_fn_broadened", C.c_int), ("n_eff", C.c_double), ("n_eff2", C.c_double), ("n_eff3", C.c_double),
                ("n_parts", C.c_int), ("n_loss_fn", C.c_int)]


class CoreParams(C.Structure):
    _fields_ = [("dos", DosParams), ("core_type", C.c_int), ("polar", C.c_int), ("qdir", C.c_double * 3),
                ("lai_broadening", C.c_int), ("lai_gaussian_width", C.c_double), ("lai_lorentzian_width", C.c_double),
                ("lai_lorentzian_scale", C.c_double), ("lai_lorentzian_offset", C.c_double),
                ("set_efermi_zero", C.c_int), ("core_chemical_shift", C.c_double)]


class Odi(C.Structure):
    """od_odi: the parsed .odi (param_read, parameters.f90:138-465)"""
    _fields_ = [("iprint", C.c_int), ("legacy_file_format", C.c_int),
                ("dos", C.c_int), ("pdos", C.c_int), ("pdis", C.c_int), ("jdos", C.c_int), ("optics", C.c_int), ("core", C.c_int),
                ("compare_dos", C.c_int), ("compare_jdos", C.c_int),
                ("fixed", C.c_int), ("adaptive", C.c_int), ("linear", C.c_int),
                ("adaptive_smearing", C.c_double), ("fixed_smearing", C.c_double), ("linear_smearing", C.c_double),
                ("efermi_choice", C.c_int), ("efermi_user", C.c_double),
                ("compute_band_energy", C.c_int), ("compute_band_gap", C.c_int), ("core_chemical_shift", C.c_double),
                ("finite_bin_correction", C.c_int), ("numerical_intdos", C.c_int), ("hybrid_linear", C.c_int),
                ("hybrid_linear_grad_tol", C.c_double), ("set_efermi_zero", C.c_int), ("dos_per_volume", C.c_int),
                ("have_dos_min_energy", C.c_int), ("have_dos_max_energy", C.c_int),
                ("dos_min_energy", C.c_double), ("dos_max_energy", C.c_double), ("dos_spacing", C.c_double), ("dos_nbins", C.c_int),
                ("projectors_string", C.c_char * 256), ("jdos_max_energy", C.c_double), ("jdos_spacing", C.c_double),
                ("num_exclude_bands", C.c_int), ("exclude_bands", C.c_int * 256),
                ("optics_geom", C.c_char * 32), ("scissor_op", C.c_double), ("optics_qdir", C.c_double * 3),
                ("optics_intraband", C.c_int), ("optics_drude_broadening", C.c_double),
                ("optics_lossfn_broadening", C.c_int), ("optics_lossfn_gaussian", C.c_double),
                ("core_geom", C.c_char * 32), ("core_type", C.c_char * 32), ("core_qdir", C.c_double * 3),
                ("core_lai_broadening", C.c_int), ("lai_gaussian_width", C.c_double), ("lai_lorentzian_width", C.c_double),
                ("lai_lorentzian_scale", C.c_double), ("lai_lorentzian_offset", C.c_double)]


class BandgapResult(C.Structure):
    _fields_ = [("vbm_energy", C.c_double), ("cbm_energy", C.c_double), ("thermal_bandgap", C.c_double),
                ("thermal_vbm_k", C.c_longlong), ("thermal_cbm_k", C.c_longlong),
                ("vbm_multiplicity", C.c_int * 2), ("cbm_multiplicity", C.c_int * 2), ("optical_multiplicity", C.c_int * 2),
                ("optical_bandgap", C.c_double * 2), ("average_bandgap", C.c_double * 2),
                ("weighted_average", C.c_double), ("thermal_multiplicity", C.c_int)]


class CoreResult(C.Structure):
    _fields_ = [("nbins", C.c_int), ("norb", C.c_int), ("emin", C.c_double), ("delta_bins", C.c_double), ("efermi", C.c_double),
                ("have_broadened", C.c_int)]


class OptadosB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"optados_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load liboptados_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                          "optados_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.od_version.restype = C.c_char_p
    L.od_last_error.restype = C.c_char_p
    L.od_last_error.argtypes = [C.c_void_p]
    L.od_kernel_launches.restype = C.c_longlong
    L.od_kernel_launches.argtypes = [C.c_void_p]
    L.od_calc_efermi_from_intdos.restype = C.c_double
    L.od_calc_efermi_from_intdos.argtypes = [c_dp, C.c_double, C.c_double, C.c_int, C.c_int, c_dp]
    _lib = L
    return L


EXPORTS = [
    "od_version", "od_ctx_create", "od_comm_get_unique_id", "od_ctx_create_rank", "od_ctx_destroy", "od_last_error",
    "od_ctx_rank", "od_ctx_nranks", "od_synchronize", "od_kernel_launches", "od_profile_reset", "od_profile_get",
    "od_device_malloc", "od_device_free", "od_memcpy_h2d", "od_memcpy_d2h", "od_dist_array", "od_allreduce",
    "od_set_system", "od_set_bands", "od_set_gradients", "od_set_gradients_from_ome", "od_get_gradients",
    "od_band_minmax", "od_broadening_defaults", "od_calculate_dos", "od_calculate_dos_at_e", "od_calculate_jdos",
    "od_calculate_jdos_optics", "od_make_weights", "od_core_prepare_matrix_elements", "od_projection_merge",
    "od_calc_epsilon_2", "od_dos_params_defaults", "od_dos_utils_calculate", "od_dos_get", "od_dos_get_weighted",
    "od_calc_efermi_from_intdos", "od_dos_utils_set_efermi", "od_jdos_params_defaults", "od_jdos_utils_calculate",
    "od_jdos_get", "od_jdos_get_weighted", "od_synth_bands", "od_synth_bands_strided", "od_get_bands", "od_get_kweights",
    "od_timer_start", "od_timer_stop", "od_measure_fp64_peak", "od_measure_fp64_peak_mode", "od_count_contributions", "od_host_alloc_pinned",
    "od_host_free_pinned", "od_set_engine",
    "od_calc_epsilon_1", "od_calc_loss_fn", "od_core_lai_broadening",
    "od_dos_utils_calculate_at_e", "od_optics_params_defaults", "od_optics_calculate", "od_optics_get", "od_core_params_defaults", "od_core_calculate",
    "od_core_get", "od_dos_utils_compute_bandgap", "od_dos_utils_compute_band_energies",
    "od_odi_defaults", "od_param_read", "od_run_last_error", "od_run_seed", "od_abi_sizeof", "od_host_register", "od_host_unregister",
    "od_io_last_error", "od_read_bands_info", "od_read_bands", "od_cell_calc_lattice", "od_cell_find_mp_grid",
    "od_read_symmetry_ops", "od_read_ome", "od_read_dome", "od_read_pdos_info", "od_read_pdos", "od_read_elnes_info",
    "od_read_elnes",
]


def f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def broadening(adaptive_smearing=0.4, fixed_smearing=0.3, linear_smearing=0.0, finite_bin_correction=True,
               numerical_intdos=False, hybrid_linear=False, hybrid_linear_grad_tol=0.01):
    return Broadening(adaptive_smearing, fixed_smearing, linear_smearing, int(finite_bin_correction),
                      int(numerical_intdos), int(hybrid_linear), hybrid_linear_grad_tol)


def dist_array(num_elements, num_nodes):
    per = (C.c_int * num_nodes)()
    rc = load().od_dist_array(C.c_int(num_elements), C.c_int(num_nodes), per)
    if rc:
        raise OptadosB200Error(rc, "Fewer kpoints than nodes. Reduce the number of nodes used!")
    return list(per)


def param_read(path):
    """param_read (parameters.f90:138-465) of one .odi file -> dict (host only: no GPU needed)"""
    L = load()
    L.od_run_last_error.restype = C.c_char_p
    p = Odi()
    rc = L.od_param_read(str(path).encode(), C.byref(p))
    if rc:
        raise OptadosB200Error(rc, L.od_run_last_error().decode())
    out = {}
    for name, ctype in Odi._fields_:
        v = getattr(p, name)
        if isinstance(v, bytes):
            v = v.decode()
        elif hasattr(v, "__len__"):
            v = list(v)
        out[name] = v
    out["exclude_bands"] = out["exclude_bands"][:out["num_exclude_bands"]]
    return out


def get_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().od_comm_get_unique_id(buf)
    if rc:
        raise OptadosB200Error(rc, "od_comm_get_unique_id failed")
    return buf.raw


class Context:
    """One GPU / one k-point shard.  Methods map 1:1 onto the C ABI."""

    def __init__(self, device=0, rank=0, nranks=1, unique_id=None):
        self.L = load()
        self.h = C.c_void_p()
        if nranks > 1:
            rc = self.L.od_ctx_create_rank(C.c_int(device), C.c_int(rank), C.c_int(nranks), unique_id, C.byref(self.h))
        else:
            rc = self.L.od_ctx_create(C.c_int(device), C.byref(self.h))
        if rc:
            raise OptadosB200Error(rc, "od_ctx_create failed (no CUDA device? optados_b200 has no CPU fallback)")
        self.nk = self.ns = self.nb = 0
        self._keep = {}

    def close(self):
        if self.h:
            self.L.od_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise OptadosB200Error(rc, self.L.od_last_error(self.h).decode())

    # ---- module data
    def set_system(self, nk, ns, nb, recip_lattice, kpoint_grid_dim, kpoint_weight, electrons_per_state):
        r = f64(recip_lattice).ravel(order="F")
        g = np.asarray(kpoint_grid_dim, dtype=np.int32)
        kw = None if kpoint_weight is None else f64(kpoint_weight)
        self._ck(self.L.od_set_system(self.h, C.c_int(nk), C.c_int(ns), C.c_int(nb), _p(r), g.ctypes.data_as(c_ip), _p(kw),
                                      C.c_double(electrons_per_state)))
        self.nk, self.ns, self.nb = nk, ns, nb

    def set_bands(self, band_energy):
        a = f64(band_energy)
        self._ck(self.L.od_set_bands(self.h, _p(a), OD_MEM_HOST))
        self._ck(self.L.od_synchronize(self.h))

    def set_gradients(self, band_gradient, deferred=False):
        """deferred=True: the library keeps the host pointer and uploads the array slab by slab inside
        calculate_dos / dos_utils_calculate, overlapped with the kernels (the array is kept alive here)."""
        a = f64(band_gradient)
        if deferred:
            self._keep["grads"] = a
            self._ck(self.L.od_set_gradients(self.h, _p(a), OD_MEM_HOST_DEFERRED))
            return
        self._ck(self.L.od_set_gradients(self.h, _p(a), OD_MEM_HOST))
        self._ck(self.L.od_synchronize(self.h))

    def set_gradients_from_ome(self, optical_mat):
        a = np.asfortranarray(optical_mat, dtype=np.complex128)
        self._ck(self.L.od_set_gradients_from_ome(self.h, a.ctypes.data_as(c_dp), OD_MEM_HOST))

    def synth_bands(self, grid, k_first, nk_local, ns, nb, lattice_a=5.43, seed=2, k_stride=1):
        g = np.asarray(grid, dtype=np.int32)
        self._ck(self.L.od_synth_bands_strided(self.h, g.ctypes.data_as(c_ip), C.c_longlong(k_first), C.c_longlong(k_stride),
                                               C.c_int(nk_local), C.c_int(ns), C.c_int(nb), C.c_double(lattice_a), C.c_uint64(seed)))
        self.nk, self.ns, self.nb = nk_local, ns, nb

    def get_bands(self):
        a = np.zeros((self.nb, self.ns, self.nk), order="F")
        self._ck(self.L.od_get_bands(self.h, _p(a)))
        return a

    def get_gradients(self):
        a = np.zeros((self.nb, 3, self.nk, self.ns), order="F")
        self._ck(self.L.od_get_gradients(self.h, _p(a)))
        return a

    def get_kweights(self):
        a = np.zeros(self.nk)
        self._ck(self.L.od_get_kweights(self.h, _p(a)))
        return a

    def band_minmax(self):
        mn, mx = C.c_double(), C.c_double()
        self._ck(self.L.od_band_minmax(self.h, C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def allreduce(self, a, op="s"):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._ck(self.L.od_allreduce(self.h, _p(a), C.c_size_t(a.size), C.c_char(op.encode())))
        return a

    # ---- hot loops
    def calculate_dos(self, dos_type, emin, delta_bins, nbins, br=None, mw=None, merge=False):
        br = br or broadening()
        dos = np.zeros((nbins, self.ns), order="F")
        intdos = np.zeros((nbins, self.ns), order="F")
        wdos = None
        norb = mwnb = mwnk = 0
        if mw is not None:
            mw = f64(mw)
            norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
            wdos = np.zeros((nbins, self.ns, norb), order="F")
        self._ck(self.L.od_calculate_dos(self.h, C.c_char(dos_type.encode()), C.c_double(emin), C.c_double(delta_bins),
                                         C.c_int(nbins), C.byref(br), _p(mw), C.c_int(norb), C.c_int(mwnb), C.c_int(mwnk),
                                         C.c_int(OD_MEM_HOST), C.c_int(int(merge)), _p(dos), _p(intdos), _p(wdos)))
        return dos, intdos, wdos

    def calculate_dos_at_e(self, dos_type, energy, delta_bins, br=None, mw=None, merge=False):
        br = br or broadening()
        d = np.zeros(self.ns)
        w = None
        norb = mwnb = mwnk = 0
        if mw is not None:
            mw = f64(mw)
            norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
            w = np.zeros((self.ns, norb), order="F")
        self._ck(self.L.od_calculate_dos_at_e(self.h, C.c_char(dos_type.encode()), C.c_double(energy), C.c_double(delta_bins),
                                              C.byref(br), _p(mw), C.c_int(norb), C.c_int(mwnb), C.c_int(mwnk),
                                              C.c_int(OD_MEM_HOST), C.c_int(int(merge)), _p(d), _p(w)))
        return d, w

    def calculate_jdos(self, jdos_type, delta_bins, nbins, efermi, br=None, scissor_op=0.0, exclude_bands=(), mw=None,
                       merge=False):
        br = br or broadening()
        jdos = np.zeros((nbins, self.ns), order="F")
        wj = None
        ng = 0
        if mw is not None:
            mw = f64(mw)
            ng = mw.shape[4]
            wj = np.zeros((nbins, self.ns, ng), order="F")
        ex = np.asarray(exclude_bands, dtype=np.int32)
        self._ck(self.L.od_calculate_jdos(self.h, C.c_char(jdos_type.encode()), C.c_double(delta_bins), C.c_int(nbins),
                                          C.byref(br), C.c_double(efermi), C.c_double(scissor_op), ex.ctypes.data_as(c_ip),
                                          C.c_int(ex.size), _p(mw), C.c_int(ng), C.c_int(OD_MEM_HOST), C.c_int(int(merge)),
                                          _p(jdos), _p(wj)))
        return jdos, wj

    def calculate_jdos_optics(self, jdos_type, delta_bins, nbins, efermi, optical_mat, geom, qdir=(0, 0, 1), symops=None,
                              br=None, scissor_op=0.0, exclude_bands=(), merge=False):
        br = br or broadening()
        ng = 6 if geom == "tensor" else 1
        jdos = np.zeros((nbins, self.ns), order="F")
        wj = np.zeros((nbins, self.ns, ng), order="F")
        om = np.asfortranarray(optical_mat, dtype=np.complex128)
        q = f64(np.asarray(qdir, dtype=np.float64))
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        ex = np.asarray(exclude_bands, dtype=np.int32)
        self._ck(self.L.od_calculate_jdos_optics(self.h, C.c_char(jdos_type.encode()), C.c_double(delta_bins), C.c_int(nbins),
                                                 C.byref(br), C.c_double(efermi), C.c_double(scissor_op),
                                                 ex.ctypes.data_as(c_ip), C.c_int(ex.size), om.ctypes.data_as(c_dp),
                                                 C.c_int(OD_MEM_HOST), C.c_int(GEOM[geom]), _p(q), _p(so), C.c_int(nsym),
                                                 C.c_int(int(merge)), _p(jdos), _p(wj)))
        return jdos, wj

    # ---- raw-pointer variants (inputs already resident in HBM: OD_MEM_DEVICE)
    def calculate_dos_ptr(self, dos_type, emin, delta_bins, nbins, mw_ptr, norb, mw_nb, mw_nk, mw_mem=OD_MEM_DEVICE, br=None,
                          merge=False):
        br = br or broadening()
        dos = np.zeros((nbins, self.ns), order="F")
        intdos = np.zeros((nbins, self.ns), order="F")
        wdos = np.zeros((nbins, self.ns, norb), order="F")
        self._ck(self.L.od_calculate_dos(self.h, C.c_char(dos_type.encode()), C.c_double(emin), C.c_double(delta_bins),
                                         C.c_int(nbins), C.byref(br), C.c_void_p(mw_ptr), C.c_int(norb), C.c_int(mw_nb),
                                         C.c_int(mw_nk), C.c_int(mw_mem), C.c_int(int(merge)), _p(dos), _p(intdos), _p(wdos)))
        return dos, intdos, wdos

    def calculate_jdos_optics_ptr(self, jdos_type, delta_bins, nbins, efermi, ome_ptr, geom, qdir=(0, 0, 1), br=None,
                                  scissor_op=0.0, ome_mem=OD_MEM_DEVICE, merge=False):
        br = br or broadening()
        ng = 6 if geom == "tensor" else 1
        jdos = np.zeros((nbins, self.ns), order="F")
        wj = np.zeros((nbins, self.ns, ng), order="F")
        q = f64(np.asarray(qdir, dtype=np.float64))
        ex = np.zeros(1, dtype=np.int32)
        self._ck(self.L.od_calculate_jdos_optics(self.h, C.c_char(jdos_type.encode()), C.c_double(delta_bins), C.c_int(nbins),
                                                 C.byref(br), C.c_double(efermi), C.c_double(scissor_op),
                                                 ex.ctypes.data_as(c_ip), C.c_int(0), C.c_void_p(ome_ptr), C.c_int(ome_mem),
                                                 C.c_int(GEOM[geom]), _p(q), None, C.c_int(0), C.c_int(int(merge)), _p(jdos), _p(wj)))
        return jdos, wj

    def core_prepare_matrix_elements_ptr(self, elnes_ptr, norb, core_type, efermi, out_ptr, polar=False, qdir=(0, 0, 1)):
        q = f64(np.asarray(qdir, dtype=np.float64))
        self._ck(self.L.od_core_prepare_matrix_elements(self.h, C.c_void_p(elnes_ptr), C.c_int(OD_MEM_DEVICE), C.c_int(norb),
                                                        C.c_int(int(polar)), _p(q), None, C.c_int(0), C.c_int(CORE_TYPE[core_type]),
                                                        C.c_double(efermi), C.c_void_p(out_ptr), C.c_int(OD_MEM_DEVICE)))

    # ---- weight builders
    def make_weights(self, optical_mat, geom, efermi, qdir=(0, 0, 1), symops=None, scissor_op=0.0):
        om = np.asfortranarray(optical_mat, dtype=np.complex128)
        ng = 6 if geom == "tensor" else 1
        mw = np.zeros((self.nb, self.nb, self.nk, self.ns, ng), order="F")
        q = f64(np.asarray(qdir, dtype=np.float64))
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        self._ck(self.L.od_make_weights(self.h, om.ctypes.data_as(c_dp), C.c_int(OD_MEM_HOST), C.c_int(GEOM[geom]), _p(q), _p(so),
                                        C.c_int(nsym), C.c_double(efermi), C.c_double(scissor_op), _p(mw), C.c_int(OD_MEM_HOST)))
        return mw

    def core_prepare_matrix_elements(self, elnes_mat, core_type, efermi, polar=False, qdir=(0, 0, 1), symops=None):
        em = np.asfortranarray(elnes_mat, dtype=np.complex128)
        norb = em.shape[0]
        mw = np.zeros((norb, self.nb, self.nk, self.ns), order="F")
        q = f64(np.asarray(qdir, dtype=np.float64))
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        self._ck(self.L.od_core_prepare_matrix_elements(self.h, em.ctypes.data_as(c_dp), C.c_int(OD_MEM_HOST), C.c_int(norb),
                                                        C.c_int(int(polar)), _p(q), _p(so), C.c_int(nsym),
                                                        C.c_int(CORE_TYPE[core_type]), C.c_double(efermi), _p(mw),
                                                        C.c_int(OD_MEM_HOST)))
        return mw

    def projection_merge(self, pdos_weights, mask):
        pw = f64(pdos_weights)
        m = np.asfortranarray(mask, dtype=np.int32)
        norb, nproj = m.shape
        pw_nb = pw.shape[1]              # pdos_mwab%nbands: may be smaller than nbands
        mw = np.zeros((nproj, pw_nb, self.nk, self.ns), order="F")
        self._ck(self.L.od_projection_merge(self.h, _p(pw), C.c_int(OD_MEM_HOST), m.ctypes.data_as(c_ip), C.c_int(norb),
                                            C.c_int(pw_nb), C.c_int(nproj), _p(mw), C.c_int(OD_MEM_HOST)))
        return mw

    def calc_epsilon_2(self, wjdos, cell_volume):
        wj = f64(wjdos)
        B, ns, ng = wj.shape
        e2 = np.zeros((B, ng), order="F")
        self._ck(self.L.od_calc_epsilon_2(self.h, _p(wj), C.c_int(B), C.c_int(ns), C.c_int(ng), C.c_double(cell_volume), _p(e2)))
        return e2

    # ---- task drivers
    def dos_params(self, broadening_type="adaptive", dos_spacing=None, dos_nbins=-1, num_electrons=(0.0,), efermi="optados",
                   br=None, **kw):
        p = DosParams()
        self.L.od_dos_params_defaults(C.byref(p))
        p.fixed = p.adaptive = p.linear = 0
        for t in ([broadening_type] if isinstance(broadening_type, str) else broadening_type):
            setattr(p, t, 1)
        if dos_spacing is not None:
            p.dos_spacing = dos_spacing
        p.dos_nbins = dos_nbins
        p.efermi_choice = EFERMI[efermi]
        ne = list(num_electrons)
        p.nspins_electrons = len(ne)
        for i, v in enumerate(ne[:2]):
            p.num_electrons[i] = v
        if br is not None:
            p.br = br
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def dos_utils_calculate(self, p, mw=None):
        res = DosResult()
        norb = mwnb = mwnk = 0
        if mw is not None:
            mw = f64(mw)
            norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
        self._ck(self.L.od_dos_utils_calculate(self.h, C.byref(p), _p(mw), C.c_int(norb), C.c_int(mwnb), C.c_int(mwnk),
                                               C.c_int(OD_MEM_HOST), C.byref(res)))
        out = dict(nbins=res.nbins, emin=res.emin, delta=res.delta_bins,
                   E=res.emin + np.arange(res.nbins, dtype=np.float64) * res.delta_bins,
                   efermi_fixed=res.efermi_fixed, efermi_adaptive=res.efermi_adaptive, efermi_linear=res.efermi_linear)
        for t, have in (("fixed", res.have_fixed), ("adaptive", res.have_adaptive), ("linear", res.have_linear)):
            if have:
                d = np.zeros((res.nbins, self.ns), order="F")
                i = np.zeros((res.nbins, self.ns), order="F")
                self._ck(self.L.od_dos_get(self.h, C.c_char(t[0].encode()), _p(d), _p(i)))
                out["dos_" + t], out["intdos_" + t] = d, i
        if res.have_weighted:
            w = np.zeros((res.nbins, self.ns, res.weighted_norb), order="F")
            self._ck(self.L.od_dos_get_weighted(self.h, _p(w)))
            out["weighted_dos"] = w
        return out

    def dos_utils_calculate_at_e(self, p, energy, delta_bins, mw=None):
        """dos_utils_calculate_at_e (dos_utils.f90:1531): -> dos_at_e(3, ns), weighted_dos_at_e(ns, norb) or None"""
        d = np.zeros((3, self.ns), order="F")
        w = None
        norb = mwnb = mwnk = 0
        if mw is not None:
            mw = f64(mw)
            norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
            w = np.zeros((self.ns, norb), order="F")
        self._ck(self.L.od_dos_utils_calculate_at_e(self.h, C.byref(p), C.c_double(energy), C.c_double(delta_bins), _p(mw), C.c_int(norb),
                                                    C.c_int(mwnb), C.c_int(mwnk), C.c_int(OD_MEM_HOST), _p(d), _p(w)))
        return d, w

    def dos_utils_set_efermi(self, p):
        ef = C.c_double()
        self._ck(self.L.od_dos_utils_set_efermi(self.h, C.byref(p), C.byref(ef)))
        return ef.value

    def jdos_utils_calculate(self, efermi, broadening_type="adaptive", jdos_spacing=0.01, jdos_max_energy=-1.0, br=None,
                             scissor_op=0.0, exclude_bands=(), mw=None, optical_mat=None, geom="polycrys", qdir=(0, 0, 1),
                             symops=None):
        p = JdosParams()
        self.L.od_jdos_params_defaults(C.byref(p))
        p.fixed = p.adaptive = p.linear = 0
        for t in ([broadening_type] if isinstance(broadening_type, str) else broadening_type):
            setattr(p, t, 1)
        p.jdos_spacing, p.jdos_max_energy, p.scissor_op, p.efermi = jdos_spacing, jdos_max_energy, scissor_op, efermi
        if br is not None:
            p.br = br
        ex = np.asarray(exclude_bands, dtype=np.int32)
        p.num_exclude_bands = ex.size
        p.exclude_bands = ex.ctypes.data_as(c_ip)
        ng = 0
        if mw is not None:
            mw = f64(mw)
            ng = mw.shape[4]
        om = None if optical_mat is None else np.asfortranarray(optical_mat, dtype=np.complex128)
        q = f64(np.asarray(qdir, dtype=np.float64))
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        res = JdosResult()
        self._ck(self.L.od_jdos_utils_calculate(self.h, C.byref(p), _p(mw), C.c_int(ng), C.c_int(OD_MEM_HOST),
                                                None if om is None else om.ctypes.data_as(c_dp), C.c_int(OD_MEM_HOST),
                                                C.c_int(GEOM[geom]), _p(q), _p(so), C.c_int(nsym), C.byref(res)))
        out = dict(nbins=res.nbins, delta=res.delta_bins, E=np.arange(res.nbins, dtype=np.float64) * res.delta_bins,
                   jdos_max_energy=p.jdos_max_energy)
        for t, have in (("fixed", res.have_fixed), ("adaptive", res.have_adaptive), ("linear", res.have_linear)):
            if have:
                d = np.zeros((res.nbins, self.ns), order="F")
                self._ck(self.L.od_jdos_get(self.h, C.c_char(t[0].encode()), _p(d)))
                out["jdos_" + t] = d
        if res.have_weighted:
            w = np.zeros((res.nbins, self.ns, res.n_geom), order="F")
            self._ck(self.L.od_jdos_get_weighted(self.h, _p(w)))
            out["weighted_jdos"] = w
        return out

    # ---- task mirrors (optics_calculate, core_calculate)
    def optics_calculate(self, optical_mat, efermi, cell_volume, broadening_type="adaptive", geom="polycrys", qdir=(1, 1, 1),
                         symops=None, jdos_spacing=0.01, jdos_max_energy=-1.0, scissor_op=0.0, lossfn_gaussian=0.0, br=None,
                         intraband=False, drude_broadening=1.0e14, dos_delta_bins=0.0):
        """optics_calculate (optics.f90:78-155).  With intraband=True epsilon comes back as (nbins, 2, n_geom, 3) =
        interband / intraband / total and loss_fn has 3 (+1 broadened) columns."""
        p = OpticsParams()
        self.L.od_optics_params_defaults(C.byref(p))
        p.jdos.fixed = p.jdos.adaptive = p.jdos.linear = 0
        for t in ([broadening_type] if isinstance(broadening_type, str) else broadening_type):
            setattr(p.jdos, t, 1)
        p.jdos.jdos_spacing, p.jdos.jdos_max_energy, p.jdos.scissor_op = jdos_spacing, jdos_max_energy, scissor_op
        p.jdos.efermi, p.jdos.cell_volume = efermi, cell_volume
        if br is not None:
            p.jdos.br = br
        p.geom, p.lossfn_gaussian = GEOM[geom], lossfn_gaussian
        p.intraband, p.drude_broadening, p.dos_delta_bins = int(intraband), float(drude_broadening), float(dos_delta_bins)
        for i in range(3):
            p.qdir[i] = float(qdir[i])
        om = np.asfortranarray(optical_mat, dtype=np.complex128)
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        res = OpticsResult()
        self._ck(self.L.od_optics_calculate(self.h, C.byref(p), om.ctypes.data_as(c_dp), C.c_int(OD_MEM_HOST), _p(so), C.c_int(nsym),
                                            C.byref(res)))
        n, ng = res.nbins, res.n_geom
        eps_shape = (n, 2, ng) if res.n_parts == 1 else (n, 2, ng, res.n_parts)
        out = dict(nbins=n, n_geom=ng, delta=res.delta_bins, E=np.zeros(n), epsilon=np.zeros(eps_shape, order="F"),
                   n_eff=res.n_eff, n_eff2=res.n_eff2, n_eff3=res.n_eff3)
        args = [_p(out["E"]), _p(out["epsilon"])]
        if res.have_derived:
            out.update(conduct=np.zeros((n, 2), order="F"), refract=np.zeros((n, 2), order="F"),
                       loss_fn=np.zeros((n, res.n_loss_fn), order="F"), absorp=np.zeros(n), reflect=np.zeros(n))
            args += [_p(out[k]) for k in ("conduct", "refract", "loss_fn", "absorp", "reflect")]
        else:
            args += [None] * 5
        self._ck(self.L.od_optics_get(self.h, *args))
        return out

    def compute_bandgap(self, efermi, num_electrons=None, nk_global=0, k_first=0):
        """dos_utils_compute_bandgap (dos_utils.f90:457-750) -> dict"""
        b = BandgapResult()
        ne = None if num_electrons is None else f64(np.asarray(num_electrons, dtype=np.float64))
        self._ck(self.L.od_dos_utils_compute_bandgap(self.h, C.c_double(efermi), _p(ne), C.c_longlong(nk_global),
                                                     C.c_longlong(k_first), C.byref(b)))
        ns = self.ns
        return dict(thermal_vbm=b.vbm_energy, thermal_cbm=b.cbm_energy, thermal_bandgap=b.thermal_bandgap,
                    thermal_vbm_k=b.thermal_vbm_k, thermal_cbm_k=b.thermal_cbm_k,
                    vbm_multiplicity=list(b.vbm_multiplicity)[:ns], cbm_multiplicity=list(b.cbm_multiplicity)[:ns],
                    optical_multiplicity=list(b.optical_multiplicity)[:ns], optical_bandgap=list(b.optical_bandgap)[:ns],
                    average_bandgap=list(b.average_bandgap)[:ns], weighted_average=b.weighted_average,
                    thermal_multiplicity=bool(b.thermal_multiplicity))

    def compute_band_energies(self, efermi):
        """dos_utils_compute_band_energies (dos_utils.f90:853-924) -> (dict(fixed=, adaptive=, linear=), from_bands)"""
        be = (C.c_double * 3)()
        ec = C.c_double()
        self._ck(self.L.od_dos_utils_compute_band_energies(self.h, C.c_double(efermi), be, C.byref(ec)))
        return dict(fixed=be[0], adaptive=be[1], linear=be[2]), ec.value

    def core_calculate(self, elnes_mat, dos_params, core_type="absorption", polar=False, qdir=(0, 0, 1), symops=None,
                       lai=None, set_efermi_zero=False, core_chemical_shift=-1.0):
        """lai: None or dict(gaussian_width=, lorentzian_width=, lorentzian_scale=, lorentzian_offset=)"""
        p = CoreParams()
        self.L.od_core_params_defaults(C.byref(p))
        p.dos = dos_params
        p.set_efermi_zero, p.core_chemical_shift = int(set_efermi_zero), float(core_chemical_shift)
        p.core_type, p.polar = CORE_TYPE[core_type], int(polar)
        for i in range(3):
            p.qdir[i] = float(qdir[i])
        if lai is not None:
            p.lai_broadening = 1
            p.lai_gaussian_width = lai.get("gaussian_width", 0.0)
            p.lai_lorentzian_width = lai.get("lorentzian_width", 0.0)
            p.lai_lorentzian_scale = lai.get("lorentzian_scale", 0.0)
            p.lai_lorentzian_offset = lai.get("lorentzian_offset", 0.0)
        em = np.asfortranarray(elnes_mat, dtype=np.complex128)
        norb = em.shape[0]
        nsym = 0 if symops is None else symops.shape[2]
        so = None if symops is None else f64(symops)
        res = CoreResult()
        self._ck(self.L.od_core_calculate(self.h, C.byref(p), em.ctypes.data_as(c_dp), C.c_int(OD_MEM_HOST), C.c_int(norb), _p(so),
                                          C.c_int(nsym), C.byref(res)))
        n = res.nbins
        out = dict(nbins=n, efermi=res.efermi, emin=res.emin, delta=res.delta_bins, E=np.zeros(n),
                   weighted_dos=np.zeros((n, self.ns, norb), order="F"))
        wb = None
        if res.have_broadened:
            out["weighted_dos_broadened"] = np.zeros((n, self.ns, norb), order="F")
            wb = _p(out["weighted_dos_broadened"])
        self._ck(self.L.od_core_get(self.h, _p(out["E"]), _p(out["weighted_dos"]), wb))
        return out

    def run_seed(self, directory, seedname):
        """the reference's file contract: <directory>/<seedname>.odi + CASTEP outputs -> .odo and .dat files (od_run_seed)"""
        self.L.od_run_last_error.restype = C.c_char_p
        rc = self.L.od_run_seed(self.h, str(directory).encode(), str(seedname).encode())
        if rc:
            raise OptadosB200Error(rc, self.L.od_run_last_error().decode() or self.L.od_last_error(self.h).decode())

    # ---- bookkeeping
    def kernel_launches(self):
        return int(self.L.od_kernel_launches(self.h))

    def profile_reset(self):
        self._ck(self.L.od_profile_reset(self.h))

    def profile_get(self):
        a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        n = C.c_longlong()
        self._ck(self.L.od_profile_get(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(n)))
        return dict(accumulate_ms=a.value, prepare_ms=b.value, contributions=c.value, lane_steps=d.value,
                    accumulate_launches=n.value)

    def calc_epsilon_1(self, E, eps2):
        """Kramers-Kronig (optics.f90:569-621): eps2 (nbins, N_geom) -> eps1"""
        E = f64(E)
        e2 = f64(np.asarray(eps2).reshape(E.size, -1))
        e1 = np.zeros_like(e2, order="F")
        self._ck(self.L.od_calc_epsilon_1(self.h, _p(E), C.c_int(E.size), C.c_int(e2.shape[1]), _p(e2), _p(e1)))
        return e1

    def calc_loss_fn(self, E, eps1, eps2, lossfn_gaussian=0.0):
        E = f64(E)
        a, b = f64(np.asarray(eps1).reshape(-1)[:E.size]), f64(np.asarray(eps2).reshape(-1)[:E.size])
        lf, lb = np.zeros(E.size), np.zeros(E.size)
        self._ck(self.L.od_calc_loss_fn(self.h, _p(E), C.c_int(E.size), _p(a), _p(b), C.c_double(lossfn_gaussian), _p(lf), _p(lb)))
        return lf, (lb if abs(lossfn_gaussian) >= 1.0e-6 else None)

    def core_lai_broadening(self, E, wdos, efermi, gaussian_width=0.0, lorentzian_width=0.0, lorentzian_scale=0.0,
                            lorentzian_offset=0.0):
        E = f64(E)
        w = f64(wdos)
        out = np.zeros_like(w, order="F")
        self._ck(self.L.od_core_lai_broadening(self.h, _p(E), C.c_int(E.size), C.c_int(w.size // E.size), C.c_double(efermi),
                                               C.c_double(gaussian_width), C.c_double(lorentzian_width),
                                               C.c_double(lorentzian_scale), C.c_double(lorentzian_offset), _p(w), _p(out)))
        return out

    def synchronize(self):
        self._ck(self.L.od_synchronize(self.h))

    def set_engine(self, engine):
        """'auto' (narrow-window engine + dense for wide windows) or 'dense'"""
        self._ck(self.L.od_set_engine(self.h, C.c_int({"auto": 0, "dense": 1}[engine])))

    def timer_start(self):
        self._ck(self.L.od_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.L.od_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def measure_fp64_peak(self, seconds=0.3, mode=0):
        """mode 0: the FP64 pipe peak (max of the DFMA and the DMMA loop); 1: DFMA loop; 2: DMMA loop"""
        t = C.c_double()
        self._ck(self.L.od_measure_fp64_peak_mode(self.h, C.c_double(seconds), C.c_int(mode), C.byref(t)))
        return t.value

    def count_contributions(self, dos_type, emin, delta_bins, nbins, br=None):
        br = br or broadening()
        n = C.c_double()
        self._ck(self.L.od_count_contributions(self.h, C.c_char(dos_type.encode()), C.c_double(emin), C.c_double(delta_bins),
                                               C.c_int(nbins), C.byref(br), C.byref(n)))
        return n.value

    def set_bands_ptr(self, host_ptr):
        """od_set_bands from a raw (e.g. pinned) host pointer; asynchronous on the ctx stream."""
        self._ck(self.L.od_set_bands(self.h, C.c_void_p(host_ptr), OD_MEM_HOST))

    def set_gradients_ptr(self, host_ptr, deferred=False):
        self._ck(self.L.od_set_gradients(self.h, C.c_void_p(host_ptr), OD_MEM_HOST_DEFERRED if deferred else OD_MEM_HOST))

    def get_bands_ptr(self, host_ptr):
        self._ck(self.L.od_get_bands(self.h, C.c_void_p(host_ptr)))

    def get_gradients_ptr(self, host_ptr):
        self._ck(self.L.od_get_gradients(self.h, C.c_void_p(host_ptr)))


def host_alloc_pinned(nbytes):
    p = C.c_void_p()
    rc = load().od_host_alloc_pinned(C.c_size_t(nbytes), C.byref(p))
    if rc:
        raise OptadosB200Error(rc, f"cudaHostAlloc({nbytes}) failed")
    return p.value


def host_free_pinned(ptr):
    load().od_host_free_pinned(C.c_void_p(ptr))


# ---------------------------------------------------------------------------------------------------------------
# input decoders (include/optados_b200.h, "input decoders"): host functions, no Context needed
class BandsInfo(C.Structure):
    _fields_ = [("nkpoints", C.c_int), ("nspins", C.c_int), ("nbands", C.c_int), ("num_electrons", C.c_double * 2),
                ("efermi_castep", C.c_double), ("real_lattice", C.c_double * 9), ("electrons_per_state", C.c_double)]


class PdosInfo(C.Structure):
    _fields_ = [("nkpoints", C.c_int), ("nspins", C.c_int), ("norbitals", C.c_int), ("nbands", C.c_int)]


class ElnesInfo(C.Structure):
    _fields_ = [("norbitals", C.c_int), ("nbands", C.c_int), ("nkpoints", C.c_int), ("nspins", C.c_int)]


def _io_ck(rc):
    if rc:
        L = load()
        L.od_io_last_error.restype = C.c_char_p
        raise OptadosB200Error(rc, L.od_io_last_error().decode())


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def read_bands(path):
    """<seed>.bands -> dict in the reference's layouts (band_energy(nb, ns, nk) in eV, ...)"""
    L = load()
    info = BandsInfo()
    _io_ck(L.od_read_bands_info(str(path).encode(), C.byref(info)))
    nk, ns, nb = info.nkpoints, info.nspins, info.nbands
    E = np.zeros((nb, ns, nk), order="F")
    kr = np.zeros((3, nk), order="F")
    kw = np.zeros(nk)
    _io_ck(L.od_read_bands(str(path).encode(), C.byref(info), _p(E), _p(kr), _p(kw)))
    return dict(nk=nk, ns=ns, nb=nb, num_electrons=np.array(info.num_electrons[:ns]), efermi_castep=info.efermi_castep,
                real_lattice_bohr=np.array(info.real_lattice[:]).reshape((3, 3), order="F"), band_energy=E, kpoint_r=kr,
                kpoint_weight=kw, electrons_per_state=info.electrons_per_state)


def cell_calc_lattice(real_lattice_bohr):
    rl = np.zeros((3, 3), order="F")
    rc = np.zeros((3, 3), order="F")
    vol = C.c_double()
    _io_ck(load().od_cell_calc_lattice(_p(f64(real_lattice_bohr)), _p(rl), _p(rc), C.byref(vol)))
    return rl, rc, vol.value


def cell_find_mp_grid(kpoints):
    k = f64(kpoints)
    grid = np.zeros(3, dtype=np.int32)
    _io_ck(load().od_cell_find_mp_grid(_p(k), C.c_int(k.shape[1]), _ip(grid)))
    return grid


def read_symmetry_ops(path):
    L = load()
    n = C.c_int()
    _io_ck(L.od_read_symmetry_ops(str(path).encode(), C.c_int(0), None, C.byref(n)))
    ops = np.zeros((3, 3, max(n.value, 1)), order="F")
    _io_ck(L.od_read_symmetry_ops(str(path).encode(), C.c_int(n.value), _p(ops), C.byref(n)))
    return ops[:, :, :n.value]


def read_ome(path, fmt, nb, nk, ns):
    """optical_mat(nb, nb, 3, nk, ns) complex, eV Ang; fmt 'b' (.ome_bin), 'l' (.cst_ome) or 't' (.ome_fmt)"""
    om = np.zeros((nb, nb, 3, nk, ns), dtype=np.complex128, order="F")
    _io_ck(load().od_read_ome(str(path).encode(), C.c_char(fmt.encode()), C.c_int(nb), C.c_int(nk), C.c_int(ns),
                              om.ctypes.data_as(c_dp)))
    return om


def read_dome(path, fmt, nb, nk, ns):
    g = np.zeros((nb, 3, nk, ns), order="F")
    _io_ck(load().od_read_dome(str(path).encode(), C.c_char(fmt.encode()), C.c_int(nb), C.c_int(nk), C.c_int(ns), _p(g)))
    return g


def read_pdos(path, fmt):
    L = load()
    info = PdosInfo()
    _io_ck(L.od_read_pdos_info(str(path).encode(), C.c_char(fmt.encode()), C.byref(info)))
    nk, ns, norb, nb = info.nkpoints, info.nspins, info.norbitals, info.nbands
    sp, rk, am = (np.zeros(norb, dtype=np.int32) for _ in range(3))
    nocc = np.zeros((nk, ns), dtype=np.int32, order="F")
    w = np.zeros((norb, nb, nk, ns), order="F")
    _io_ck(L.od_read_pdos(str(path).encode(), C.c_char(fmt.encode()), C.byref(info), _ip(sp), _ip(rk), _ip(am), _ip(nocc), _p(w)))
    return dict(pdos_weights=w, pdos_species=sp, pdos_rank=rk, pdos_am=am, nbands_occ=nocc)


def read_elnes(path, fmt):
    L = load()
    info = ElnesInfo()
    _io_ck(L.od_read_elnes_info(str(path).encode(), C.c_char(fmt.encode()), C.byref(info)))
    norb, nb, nk, ns = info.norbitals, info.nbands, info.nkpoints, info.nspins
    sp, rk, sh, am = (np.zeros(norb, dtype=np.int32) for _ in range(4))
    m = np.zeros((norb, nb, 3, nk, ns), dtype=np.complex128, order="F")
    _io_ck(L.od_read_elnes(str(path).encode(), C.c_char(fmt.encode()), C.byref(info), _ip(sp), _ip(rk), _ip(sh), _ip(am),
                           m.ctypes.data_as(c_dp)))
    return dict(elnes_mat=m, elnes_species=sp, elnes_rank=rk, elnes_shell=sh, elnes_am=am)

"""ctypes front-end of the CPU oracle (oracle/libod_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It mirrors the reference's task drivers on top of the C restatement:
  dos_utils_calculate   dos_utils.f90:92-293      jdos_utils_calculate  jdos_utils.f90:61-174
  optics (eps2 / eps1)  optics.f90:76-152         core                  core.f90:38-80
  pdos                  pdos.F90:48-92
Arrays are numpy, Fortran-ordered, in the reference's layouts (see od_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class System(C.Structure):
    _fields_ = [("nk", C.c_int), ("ns", C.c_int), ("nb", C.c_int),
                ("band_energy", c_dp), ("band_gradient", c_dp), ("kpoint_weight", c_dp),
                ("recip_lattice", C.c_double * 9), ("kpoint_grid_dim", C.c_int * 3),
                ("electrons_per_state", C.c_double)]


class Bandgap(C.Structure):
    _fields_ = [("thermal_vbm", C.c_double), ("thermal_cbm", C.c_double), ("thermal_bandgap", C.c_double),
                ("thermal_vbm_k", C.c_int), ("thermal_cbm_k", C.c_int),
                ("vbm_multiplicity", C.c_int * 2), ("cbm_multiplicity", C.c_int * 2), ("optical_multiplicity", C.c_int * 2),
                ("optical_bandgap", C.c_double * 2), ("average_bandgap", C.c_double * 2),
                ("weighted_average", C.c_double), ("thermal_multiplicity", C.c_int)]


class Broadening(C.Structure):
    _fields_ = [("adaptive_smearing", C.c_double), ("fixed_smearing", C.c_double),
                ("linear_smearing", C.c_double), ("finite_bin_correction", C.c_int),
                ("numerical_intdos", C.c_int), ("hybrid_linear", C.c_int),
                ("hybrid_linear_grad_tol", C.c_double)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libod_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.odo_gaussian.restype = C.c_double
        L.odo_gaussian.argtypes = [C.c_double] * 3
        L.odo_erf.restype = C.c_double
        L.odo_erf.argtypes = [C.c_double]
        L.odo_doslin.restype = C.c_double
        L.odo_doslin.argtypes = [C.c_double] * 6 + [c_dp, c_ip]
        L.odo_calc_efermi_from_intdos.restype = C.c_double
        L.odo_efermi_insulator.restype = C.c_double
        L.odo_calc_band_energies.restype = C.c_double
        L.odo_band_energy_from_bands.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def f64(a):
    return np.asfortranarray(a, dtype=np.float64)


class Sys:
    """Holds the arrays of one (k-shard of a) band structure and the ctypes view of it."""

    def __init__(self, band_energy, kpoint_weight, recip_lattice, kpoint_grid_dim, band_gradient=None,
                 electrons_per_state=None, num_electrons=None, cell_volume=None):
        self.band_energy = f64(band_energy)
        self.nb, self.ns, self.nk = self.band_energy.shape
        self.kpoint_weight = f64(kpoint_weight)
        self.recip_lattice = f64(recip_lattice)
        self.kpoint_grid_dim = np.asarray(kpoint_grid_dim, dtype=np.int32)
        self.band_gradient = None if band_gradient is None else f64(band_gradient)
        self.electrons_per_state = float(electrons_per_state if electrons_per_state is not None
                                         else (2.0 if self.ns < 2 else 1.0))
        self.num_electrons = None if num_electrons is None else f64(num_electrons)
        self.cell_volume = cell_volume
        s = System()
        s.nk, s.ns, s.nb = self.nk, self.ns, self.nb
        s.band_energy = _p(self.band_energy)
        s.band_gradient = _p(self.band_gradient)
        s.kpoint_weight = _p(self.kpoint_weight)
        for j in range(3):
            for i in range(3):
                s.recip_lattice[i + 3 * j] = self.recip_lattice[i, j]
            s.kpoint_grid_dim[j] = int(self.kpoint_grid_dim[j])
        s.electrons_per_state = self.electrons_per_state
        self.c = s

    def kslice(self, k0, k1):
        g = None if self.band_gradient is None else self.band_gradient[:, :, k0:k1, :]
        return Sys(self.band_energy[:, :, k0:k1], self.kpoint_weight[k0:k1], self.recip_lattice,
                   self.kpoint_grid_dim, g, self.electrons_per_state, self.num_electrons, self.cell_volume)


def make_broadening(adaptive_smearing=0.4, fixed_smearing=0.3, linear_smearing=0.0,
                    finite_bin_correction=True, numerical_intdos=False, hybrid_linear=False,
                    hybrid_linear_grad_tol=0.01):
    return Broadening(adaptive_smearing, fixed_smearing, linear_smearing, int(finite_bin_correction),
                      int(numerical_intdos), int(hybrid_linear), hybrid_linear_grad_tol)


def gradient_from_ome(optical_mat):
    """electronic.f90:280-291"""
    nb = optical_mat.shape[0]
    idx = np.arange(nb)
    return f64(optical_mat[idx, idx, :, :, :].real)


def dos_energy_scale(sys, dos_spacing, dos_nbins=-1, dos_min_energy=None, dos_max_energy=None,
                     minmax=None):
    """dos_utils.f90:927-1010.  Returns (E, delta_bins, nbins).  `minmax` overrides the local
    minval/maxval (what the MIN/MAX comms_reduce would deliver on a shard)."""
    L = lib()
    mn, mx = (float(sys.band_energy.min()), float(sys.band_energy.max())) if minmax is None else minmax
    nb = C.c_int(dos_nbins)
    emin = C.c_double()
    delta = C.c_double()
    L.odo_dos_energy_scale(C.c_double(mn), C.c_double(mx),
                           C.c_double(0.0 if dos_min_energy is None else dos_min_energy),
                           C.c_double(0.0 if dos_max_energy is None else dos_max_energy),
                           C.c_int(dos_min_energy is not None), C.c_int(dos_max_energy is not None),
                           C.c_double(dos_spacing), C.byref(nb), C.byref(emin), C.byref(delta))
    n = nb.value
    E = emin.value + np.arange(n, dtype=np.float64) * delta.value  # :996 real(idos-1,dp)*delta_bins
    return E, delta.value, n


def calculate_dos(dos_type, sys, br, E, delta_bins, mw=None, nthreads=1):
    L = lib()
    B = E.size
    dos = np.zeros((B, sys.ns), order="F")
    intdos = np.zeros((B, sys.ns), order="F")
    wdos = None
    norb = mwnb = mwnk = 0
    if mw is not None:
        mw = f64(mw)
        norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
        wdos = np.zeros((B, sys.ns, norb), order="F")
    E = f64(E)
    if nthreads > 1:
        rc = L.odo_calculate_dos_mt(C.c_int(nthreads), C.c_char(dos_type.encode()), C.byref(sys.c), C.byref(br),
                                    _p(E), C.c_int(B), C.c_double(delta_bins), _p(mw), C.c_int(norb),
                                    C.c_int(mwnb), C.c_int(mwnk), _p(dos), _p(intdos), _p(wdos))
    else:
        rc = L.odo_calculate_dos(C.c_char(dos_type.encode()), C.byref(sys.c), C.byref(br), _p(E), C.c_int(B),
                                 C.c_double(delta_bins), _p(mw), C.c_int(norb), C.c_int(mwnb), C.c_int(mwnk),
                                 _p(dos), _p(intdos), _p(wdos))
    if rc:
        raise RuntimeError(f"oracle calculate_dos failed rc={rc}")
    return dos, intdos, wdos


def calculate_dos_at_e(dos_type, sys, br, delta_bins, energy, mw=None):
    L = lib()
    d = np.zeros(sys.ns)
    w = None
    norb = mwnb = mwnk = 0
    if mw is not None:
        mw = f64(mw)
        norb, mwnb, mwnk = mw.shape[0], mw.shape[1], mw.shape[2]
        w = np.zeros((sys.ns, norb), order="F")
    rc = L.odo_calculate_dos_at_e(C.c_char(dos_type.encode()), C.byref(sys.c), C.byref(br),
                                  C.c_double(delta_bins), C.c_double(energy), _p(mw), C.c_int(norb),
                                  C.c_int(mwnb), C.c_int(mwnk), _p(d), _p(w))
    if rc:
        raise RuntimeError(f"oracle calculate_dos_at_e failed rc={rc}")
    return d, w


def calc_efermi_from_intdos(intdos, E, num_electrons):
    L = lib()
    intdos = f64(intdos)
    E = f64(E)
    ne = f64(num_electrons)
    return L.odo_calc_efermi_from_intdos(_p(intdos), _p(E), C.c_int(E.size), C.c_int(intdos.shape[1]), _p(ne))


def efermi_insulator(sys):
    return lib().odo_efermi_insulator(C.byref(sys.c), _p(sys.num_electrons))


def compute_bandgap(sys, efermi):
    """dos_utils_compute_bandgap (dos_utils.f90:457-750) -> dict"""
    b = Bandgap()
    ne = f64(sys.num_electrons if sys.num_electrons is not None else np.ones(sys.ns))
    rc = lib().odo_compute_bandgap(C.byref(sys.c), C.c_double(efermi), _p(ne), C.byref(b))
    if rc:
        raise RuntimeError("oracle compute_bandgap failed")
    ns = sys.ns
    return dict(thermal_vbm=b.thermal_vbm, thermal_cbm=b.thermal_cbm, thermal_bandgap=b.thermal_bandgap,
                thermal_vbm_k=b.thermal_vbm_k, thermal_cbm_k=b.thermal_cbm_k,
                vbm_multiplicity=list(b.vbm_multiplicity)[:ns], cbm_multiplicity=list(b.cbm_multiplicity)[:ns],
                optical_multiplicity=list(b.optical_multiplicity)[:ns], optical_bandgap=list(b.optical_bandgap)[:ns],
                average_bandgap=list(b.average_bandgap)[:ns], weighted_average=b.weighted_average,
                thermal_multiplicity=bool(b.thermal_multiplicity))


def calc_band_energies(dos, E, delta_bins, efermi):
    """calc_band_energies (dos_utils.f90:753-795)"""
    dos = f64(dos)
    E = f64(E)
    return lib().odo_calc_band_energies(_p(dos), _p(E), C.c_int(E.size), C.c_int(dos.shape[1]), C.c_double(delta_bins),
                                        C.c_double(efermi))


def band_energy_from_bands(sys, efermi):
    """the 'From CASTEP' band energy of dos_utils_compute_band_energies (dos_utils.f90:900-908)"""
    return lib().odo_band_energy_from_bands(C.byref(sys.c), C.c_double(efermi))


def jdos_energy_scale(sys, efermi, jdos_spacing=0.01, jdos_max_energy=-1.0, max_band_energy=None):
    L = lib()
    mx = float(sys.band_energy.max()) if max_band_energy is None else max_band_energy
    jm = C.c_double(jdos_max_energy)
    nb = C.c_int(0)
    delta = C.c_double()
    L.odo_jdos_energy_scale(C.c_double(mx), C.c_double(efermi), C.c_double(jdos_spacing), C.byref(jm),
                            C.byref(nb), C.byref(delta))
    E = np.arange(nb.value, dtype=np.float64) * delta.value
    return E, delta.value, nb.value


def calculate_jdos(jdos_type, sys, br, E, delta_bins, efermi, scissor_op=0.0, exclude_bands=(), mw=None,
                   nthreads=1):
    L = lib()
    B = E.size
    E = f64(E)
    jdos = np.zeros((B, sys.ns), order="F")
    wj = None
    ng = 0
    if mw is not None:
        mw = f64(mw)
        ng = mw.shape[4]
        wj = np.zeros((B, sys.ns, ng), order="F")
    ex = np.asarray(exclude_bands, dtype=np.int32)
    args = [C.c_char(jdos_type.encode()), C.byref(sys.c), C.byref(br), _p(E), C.c_int(B), C.c_double(delta_bins),
            C.c_double(efermi), C.c_double(scissor_op), ex.ctypes.data_as(c_ip), C.c_int(ex.size), _p(mw),
            C.c_int(ng), _p(jdos), _p(wj)]
    rc = L.odo_calculate_jdos_mt(C.c_int(nthreads), *args) if nthreads > 1 else L.odo_calculate_jdos(*args)
    if rc:
        raise RuntimeError(f"oracle calculate_jdos failed rc={rc}")
    return jdos, wj


GEOM = {"polycrys": 0, "polar": 1, "unpolar": 2, "tensor": 3}


def make_weights(sys, optical_mat, geom, efermi, qdir=(0, 0, 1), symops=None, scissor_op=0.0):
    L = lib()
    om = np.asfortranarray(optical_mat, dtype=np.complex128)
    ng = 6 if geom == "tensor" else 1
    mw = np.zeros((sys.nb, sys.nb, sys.nk, sys.ns, ng), order="F")
    q = f64(np.asarray(qdir, dtype=np.float64))
    nsym = 0 if symops is None else symops.shape[2]
    so = None if symops is None else f64(symops)
    rc = L.odo_make_weights(C.byref(sys.c), om.ctypes.data_as(c_dp), C.c_int(GEOM[geom]), _p(q), _p(so),
                            C.c_int(nsym), C.c_double(efermi), C.c_double(scissor_op), _p(mw))
    if rc:
        raise RuntimeError(f"oracle make_weights failed rc={rc}")
    return mw


def calc_epsilon_2(wjdos, cell_volume):
    L = lib()
    wj = f64(wjdos)
    B, ns, ng = wj.shape
    e2 = np.zeros((B, ng), order="F")
    L.odo_calc_epsilon_2(_p(wj), C.c_int(B), C.c_int(ns), C.c_int(ng), C.c_double(cell_volume), _p(e2))
    return e2


def calc_epsilon_1(eps2, E):
    L = lib()
    e2 = f64(eps2)
    E = f64(E)
    e1 = np.zeros_like(e2, order="F")
    L.odo_calc_epsilon_1(_p(e2), _p(E), C.c_int(E.size), C.c_int(e2.shape[1]), _p(e1))
    return e1


def calc_loss_fn(eps1, eps2, E, lossfn_gaussian=0.0):
    """optics.f90:622-720, component 1, no intraband -> (loss_fn, broadened or None)"""
    L = lib()
    E = f64(E)
    a, b = f64(np.asarray(eps1).reshape(-1)[:E.size]), f64(np.asarray(eps2).reshape(-1)[:E.size])
    lf, lb = np.zeros(E.size), np.zeros(E.size)
    br = abs(lossfn_gaussian) >= 1.0e-6          # parameters.f90:371
    L.odo_calc_loss_fn(_p(a), _p(b), _p(E), C.c_int(E.size), C.c_int(int(br)), C.c_double(lossfn_gaussian), _p(lf), _p(lb))
    return lf, (lb if br else None)


def optics_sum_rules(E, eps2, loss_fn, cell_volume):
    """N_eff (optics.f90:552-563), N_eff2, N_eff3 (:700-719); eps2 / loss_fn: component 1, no intraband"""
    L = lib()
    E = f64(E)
    a, b = f64(np.asarray(eps2).reshape(-1)[:E.size]), f64(np.asarray(loss_fn).reshape(-1)[:E.size])
    n1, n2, n3 = C.c_double(), C.c_double(), C.c_double()
    L.odo_optics_sum_rules(_p(E), _p(a), _p(b), C.c_int(E.size), C.c_double(cell_volume), C.byref(n1), C.byref(n2), C.byref(n3))
    return n1.value, n2.value, n3.value


def calc_optics_derived(E, eps1, eps2):
    """optics.f90:722-825, component 1, no intraband -> dict(conduct (B,2), refract (B,2), absorp, reflect)"""
    L = lib()
    E = f64(E)
    a, b = f64(np.asarray(eps1).reshape(-1)[:E.size]), f64(np.asarray(eps2).reshape(-1)[:E.size])
    n = E.size
    out = dict(conduct=np.zeros((n, 2), order="F"), refract=np.zeros((n, 2), order="F"), absorp=np.zeros(n), reflect=np.zeros(n))
    L.odo_calc_optics_derived(_p(E), _p(a), _p(b), C.c_int(n), _p(out["conduct"]), _p(out["refract"]), _p(out["absorp"]),
                              _p(out["reflect"]))
    return out


def optics_intraband(wjdos, weighted_dos_at_e, E, cell_volume, drude_broadening=1.0e14, lossfn_gaussian=0.0, derived=True):
    """the intraband branches of optics.f90 (see odo_optics_intraband in od_oracle.c; unpinned: no complete Al4 fixture)"""
    L = lib()
    wj = f64(wjdos)
    n, ns, ng = wj.shape
    wd = f64(weighted_dos_at_e)
    E = f64(E)
    lossb = abs(lossfn_gaussian) >= 1e-6
    out = dict(epsilon=np.zeros((n, 2, ng, 3), order="F"), loss_fn=np.zeros((n, 4 if lossb else 3), order="F"),
               conduct=np.zeros((n, 2), order="F"), refract=np.zeros((n, 2), order="F"), absorp=np.zeros(n), reflect=np.zeros(n))
    ne, ne2, ne3 = C.c_double(), C.c_double(), C.c_double()
    L.odo_optics_intraband(_p(wj), _p(wd), _p(E), C.c_int(n), C.c_int(ns), C.c_int(ng), C.c_double(cell_volume),
                           C.c_double(drude_broadening), C.c_int(int(lossb)), C.c_double(lossfn_gaussian), C.c_int(int(derived)),
                           _p(out["epsilon"]), _p(out["loss_fn"]), _p(out["conduct"]), _p(out["refract"]), _p(out["absorp"]),
                           _p(out["reflect"]), C.byref(ne), C.byref(ne2), C.byref(ne3))
    out.update(n_eff=ne.value, n_eff2=ne2.value, n_eff3=ne3.value)
    return out


def core_lai_broadening(wdos, E, efermi, gaussian_width=0.0, lorentzian_width=0.0, lorentzian_scale=0.0, lorentzian_offset=0.0):
    """core.f90:70-76: weighted_dos(B, ns, norb) -> weighted_dos_broadened"""
    L = lib()
    E = f64(E)
    w = f64(wdos)
    ncols = w.size // E.size
    out = np.zeros_like(w, order="F")
    if lorentzian_width > 1e-14 or lorentzian_scale > 0.00001:          # parameters.f90:402, core.f90:74
        L.odo_core_lorentzian(_p(w), _p(E), C.c_int(E.size), C.c_int(ncols), C.c_double(efermi), C.c_double(lorentzian_width),
                              C.c_double(lorentzian_scale), C.c_double(lorentzian_offset), _p(out))
        if gaussian_width > 1e-14:
            tmp = out.copy(order="F")
            out[...] = 0.0
            L.odo_core_gaussian(_p(tmp), _p(E), C.c_int(E.size), C.c_int(ncols), C.c_double(gaussian_width), _p(out))
    elif gaussian_width > 1e-14:
        L.odo_core_gaussian(_p(w), _p(E), C.c_int(E.size), C.c_int(ncols), C.c_double(gaussian_width), _p(out))
    return out


CORE_TYPE = {"absorption": 0, "emission": 1, "all": 2}


def core_prepare_matrix_elements(sys, elnes_mat, core_type, efermi, polar=False, qdir=(0, 0, 1), symops=None):
    L = lib()
    em = np.asfortranarray(elnes_mat, dtype=np.complex128)
    norb = em.shape[0]
    mw = np.zeros((norb, sys.nb, sys.nk, sys.ns), order="F")
    q = f64(np.asarray(qdir, dtype=np.float64))
    nsym = 0 if symops is None else symops.shape[2]
    so = None if symops is None else f64(symops)
    rc = L.odo_core_prepare_matrix_elements(C.byref(sys.c), em.ctypes.data_as(c_dp), C.c_int(norb),
                                            C.c_int(int(polar)), _p(q), _p(so), C.c_int(nsym),
                                            C.c_int(CORE_TYPE[core_type]), C.c_double(efermi), _p(mw))
    if rc:
        raise RuntimeError(f"oracle core_prepare_matrix_elements failed rc={rc}")
    return mw


def projection_merge(pdos_weights, mask):
    """mask: (norb_in, nproj) int 0/1"""
    L = lib()
    pw = f64(pdos_weights)
    norb, nb, nk, ns = pw.shape
    m = np.asfortranarray(mask, dtype=np.int32)
    nproj = m.shape[1]
    mw = np.zeros((nproj, nb, nk, ns), order="F")
    L.odo_projection_merge(_p(pw), m.ctypes.data_as(c_ip), C.c_int(norb), C.c_int(nproj), C.c_int(nb),
                           C.c_int(nk), C.c_int(ns), _p(mw))
    return mw


# ----------------------------------------------------------------------------------------------
# Task drivers (serial restatement of the reference's control flow around the hot loops)
# ----------------------------------------------------------------------------------------------
def dos_utils_calculate(sys, broadening="adaptive", dos_spacing=0.1, dos_nbins=-1, br=None, mw=None,
                        efermi_choice="optados", nthreads=1):
    """dos_utils.f90:92-293 for ONE broadening type.  Returns dict(E, delta, dos, intdos, wdos, efermi)."""
    br = br or make_broadening()
    # parameters.f90:459-461: unset dos_spacing defaults to the SINGLE precision literal 0.005
    E, delta, n = dos_energy_scale(sys, dos_spacing, dos_nbins)
    t = broadening[0]
    dos, intdos, wdos = calculate_dos(t, sys, br, E, delta, mw, nthreads)
    out = dict(E=E, delta=delta, nbins=n, dos=dos, intdos=intdos, wdos=wdos)
    if efermi_choice == "optados":
        out["efermi"] = calc_efermi_from_intdos(intdos, E, sys.num_electrons)
    return out


def jdos_utils_calculate(sys, efermi, broadening="adaptive", jdos_spacing=0.01, jdos_max_energy=-1.0,
                         br=None, mw=None, scissor_op=0.0, exclude_bands=(), nthreads=1):
    br = br or make_broadening()
    E, delta, n = jdos_energy_scale(sys, efermi, jdos_spacing, jdos_max_energy)
    jdos, wj = calculate_jdos(broadening[0], sys, br, E, delta, efermi, scissor_op, exclude_bands, mw, nthreads)
    return dict(E=E, delta=delta, nbins=n, jdos=jdos, wjdos=wj)

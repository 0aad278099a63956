"""ctypes binding of libpcfd_b200.so (include/pcfd.h) -- the only compute path.

There is deliberately no fallback: if the CUDA library is missing or no B200 is
visible, loading / `pcfd_create` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCFD_B200_LIB") or os.path.join(_HERE, "libpcfd_b200.so")   # override: tuning builds

NEQN, NVARS, NTERMS = 5, 10, 9
EQNSET_COMPRESSIBLE_EULER_FR = 0
EQNSET_COMPRESSIBLE_NS_FR = 1
EQNSET_COMPRESSIBLE_EULER = 2
EQNSET_COMPRESSIBLE_NS = 3

BC_PARALLEL, BC_DIRICHLET, BC_NEUMANN, BC_IMPERMEABLE_WALL, BC_NOSLIP = 0, 1, 2, 3, 4
BC_FARFIELD_VISCOUS, BC_FARFIELD, BC_SONIC_INFLOW, BC_SONIC_OUTFLOW, BC_SYMMETRY = 5, 6, 7, 8, 9

F_Q, F_QGRAD, F_LIMITER, F_B, F_X, F_TIMESTEP, F_BETA, F_LSQ_S, F_LSQ_SW, F_A, F_MUT = range(11)
F_TVAR, F_TGRAD, F_WALLDIST, F_TURB_B, F_TURB_X, F_TURB_A = range(11, 17)
F_QOLD, F_QOLDM1 = 17, 18     # conservative variables at t^n, t^{n-1} (nnode*nvars; unsteady runs)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# every symbol include/pcfd.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pcfd_abi_version", "pcfd_last_error", "pcfd_create", "pcfd_destroy", "pcfd_set_stream", "pcfd_synchronize",
    "pcfd_set_cfl", "pcfd_field_size", "pcfd_set_field", "pcfd_get_field", "pcfd_field_device_ptr", "pcfd_crs_sizes",
    "pcfd_get_crs", "pcfd_lsq_coefficients", "pcfd_update_bcs", "pcfd_gradient", "pcfd_limiter", "pcfd_residual",
    "pcfd_timestep", "pcfd_explicit_solve", "pcfd_jacobian", "pcfd_prepare_sgs", "pcfd_blank_x", "pcfd_sgs",
    "pcfd_apply_dq", "pcfd_explicit_iterate", "pcfd_implicit_iterate", "pcfd_launch_count",
    "pcfd_profile_enable", "pcfd_profile_reset", "pcfd_profile_count", "pcfd_profile_get",
    "pcfd_ipc_export", "pcfd_ipc_open", "pcfd_ipc_close",
    "pcfd_turb_compute", "pcfd_turb_phase", "pcfd_halo_configure",
    "pcfd_chem_create", "pcfd_chem_destroy", "pcfd_chem_last_error", "pcfd_chem_mass_production",
    "pcfd_create_fr", "pcfd_widths", "pcfd_limiter_raw", "pcfd_residual_fused", "pcfd_clip_fallbacks",
    "pcfd_set_time_integration", "pcfd_set_gradient_type", "pcfd_set_jacobian_type",
    "pcfd_chem_source_term", "pcfd_chem_source_term_device", "pcfd_halo_width", "pcfd_halo_send_total", "pcfd_halo_pack", "pcfd_halo_recv_ptr",
    "pcfd_comm_blob_size", "pcfd_comm_export", "pcfd_comm_connect", "pcfd_comm_disconnect", "pcfd_comm_connected",
    "pcfd_comm_post", "pcfd_comm_wait", "pcfd_comm_update", "pcfd_comm_allgather", "pcfd_comm_debug_flags", "pcfd_gmres",
    "pcfd_forces_configure", "pcfd_forces_areas", "pcfd_forces_compute", "pcfd_forces_get", "pcfd_zeroed_updates",
    "pcfd_wall_distance", "pcfd_crs_transpose", "pcfd_crs_ghost_blocks",
]


class ForcesDesc(C.Structure):
    """pcfd_forces_desc (include/pcfd.h)"""
    _fields_ = [("nbodies", C.c_int), ("num_bcs", C.c_int), ("body_offsets", _ip), ("body_factags", _ip),
                ("moment_pt", _dp), ("moment_axis", _dp), ("bedges_factag", _ip), ("cg", _dp),
                ("liftdir", C.c_double * 3), ("dragdir", C.c_double * 3), ("velocity", C.c_double)]


SURF_CP, SURF_YPLUS, SURF_CF = 0, 1, 2


class MeshDesc(C.Structure):
    _fields_ = [("nnode", C.c_int), ("gnode", C.c_int), ("nbnode", C.c_int),
                ("nedge", C.c_int), ("nbedge", C.c_int), ("ngedge", C.c_int),
                ("edges_n", _ip), ("edges_a", _dp), ("bedges_n", _ip), ("bedges_a", _dp), ("bedges_bctype", _ip),
                ("xyz", _dp), ("vol", _dp), ("ipsp", _ip), ("psp", _ip), ("bedges_twall", _dp)]


class Params(C.Structure):
    _fields_ = [("eqnset", C.c_int), ("sorder", C.c_int), ("limiter", C.c_int), ("no_cvbc", C.c_int),
                ("gamma", C.c_double), ("chi", C.c_double), ("cfl", C.c_double), ("qinf", C.c_double * NVARS),
                ("enable_vnn", C.c_int), ("vnn", C.c_double), ("Re", C.c_double), ("Pr", C.c_double),
                ("PrT", C.c_double), ("tref", C.c_double), ("mach", C.c_double), ("turb_model", C.c_int)]


CHEM_MAX_SPECIES, CHEM_MAX_REACTIONS = 16, 32


class ChemModelDesc(C.Structure):
    """pcfd_chem_model (include/pcfd.h)."""
    _S, _R = CHEM_MAX_SPECIES, CHEM_MAX_REACTIONS
    _fields_ = [("nspecies", C.c_int), ("nreactions", C.c_int),
                ("mw", C.c_double * _S), ("nasa7", C.c_double * 7 * 2 * _S),
                ("rxn_type", C.c_int * _R), ("third_body", C.c_int * _R), ("backward_given", C.c_int * _R),
                ("rxn_type_b", C.c_int * _R), ("nsp", C.c_int * _R), ("species", C.c_int * _S * _R),
                ("A", C.c_double * _R), ("EA", C.c_double * _R), ("n", C.c_double * _R),
                ("Ab", C.c_double * _R), ("EAb", C.c_double * _R), ("nb", C.c_double * _R),
                ("nup", C.c_double * _S * _R), ("nupp", C.c_double * _S * _R), ("tbeff", C.c_double * _S * _R)]


class TransportDesc(C.Structure):
    """pcfd_transport_model (include/pcfd.h)."""
    _S = CHEM_MAX_SPECIES
    _fields_ = [("nmu", C.c_int * _S), ("nk", C.c_int * _S),
                ("mu_fit", C.c_double * 6 * 3 * _S), ("k_fit", C.c_double * 6 * 3 * _S),
                ("mu_white", C.c_double * 4 * _S), ("k_white", C.c_double * 4 * _S)]


def fill_transport(t, g):
    """Fill a pcfd_transport_model-shaped ctypes struct from the flat species transport tables of a fixture / a reference
    ChemModel: species_mu_fit, species_k_fit [ns,3,6] (rows [Tlo, Thi, A, B, C, D]); species_white [ns,8] (viscosity then
    conductivity: value, T0, S, transition T); species_fit_counts [ns,2] (viscosity, conductivity ranges)."""
    cnt = np.asarray(g["species_fit_counts"]).reshape(-1, 2)
    ns = len(cnt)
    mu, k, wh = (np.asarray(g["species_mu_fit"]).reshape(ns, 3, 6), np.asarray(g["species_k_fit"]).reshape(ns, 3, 6),
                 np.asarray(g["species_white"]).reshape(ns, 8))
    for i in range(ns):
        t.nmu[i], t.nk[i] = int(cnt[i, 0]), int(cnt[i, 1])
        for r in range(3):
            for j in range(6):
                t.mu_fit[i][r][j] = float(mu[i, r, j])
                t.k_fit[i][r][j] = float(k[i, r, j])
        for j in range(4):
            t.mu_white[i][j] = float(wh[i, j])
            t.k_white[i][j] = float(wh[i, 4 + j])
    return t


class FrParams(C.Structure):
    """pcfd_fr_params (include/pcfd.h): the reacting eqnset's chemistry tables, reference values and free stream."""
    _fields_ = [("chem", ChemModelDesc),
                ("ref_density", C.c_double), ("ref_velocity", C.c_double), ("ref_temperature", C.c_double),
                ("ref_pressure", C.c_double), ("ref_time", C.c_double), ("ref_specific_enthalpy", C.c_double),
                ("pref", C.c_double), ("dt", C.c_double), ("use_local_dt", C.c_int), ("rxn_on", C.c_int),
                ("qinf", C.c_double * (3 * CHEM_MAX_SPECIES + 6)),
                ("transport", TransportDesc), ("ref_viscosity", C.c_double), ("ref_k", C.c_double)]


def fill_chem_model(md, t):
    """Fill a pcfd_chem_model-shaped ctypes struct from the flat tables of a fixture / a reference ChemModel:
    t = dict(dims=[ns, nr], species_mw, species_nasa7 [ns*14], rxn_A_EA_n [nr*3], rxn_flags [nr*4: type, third body,
    backward given, nsp], rxn_species [nr*ns local->global], rxn_nup, rxn_nupp, rxn_tbeff [nr*ns by local index])."""
    ns, nr = int(t["dims"][0]), int(t["dims"][1])
    md.nspecies, md.nreactions = ns, nr
    coeff = np.asarray(t["species_nasa7"]).reshape(ns, 2, 7)
    for i in range(ns):
        md.mw[i] = float(t["species_mw"][i])
        for r in range(2):
            for k in range(7):
                md.nasa7[i][r][k] = float(coeff[i, r, k])
    rk = np.asarray(t["rxn_A_EA_n"]).reshape(nr, 3)
    fl = np.asarray(t["rxn_flags"]).reshape(nr, 4)
    sp = np.asarray(t["rxn_species"]).reshape(nr, ns)
    nup, nupp, tb = (np.asarray(t[k]).reshape(nr, ns) for k in ("rxn_nup", "rxn_nupp", "rxn_tbeff"))
    for j in range(nr):
        md.rxn_type[j], md.third_body[j], md.backward_given[j], md.nsp[j] = (int(v) for v in fl[j])
        md.A[j], md.EA[j], md.n[j] = (float(v) for v in rk[j])
        for k in range(int(fl[j, 3])):
            md.species[j][k] = int(sp[j, k])
            md.nup[j][k], md.nupp[j][k], md.tbeff[j][k] = float(nup[j, k]), float(nupp[j, k]), float(tb[j, k])
    return md


_lib = None


_libs = {}
FMA_LIB_PATH = os.path.join(_HERE, "libpcfd_b200_fma.so")   # the same sources built with --fmad=true (bench.py: price of bit-exactness)


def load_library(path=LIB_PATH):
    """dlopen the CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if path == LIB_PATH and _lib is not None:
        return _lib
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.pcfd_last_error.restype = C.c_char_p
    lib.pcfd_last_error.argtypes = [C.c_void_p]
    lib.pcfd_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]
    lib.pcfd_field_size.restype = C.c_size_t
    lib.pcfd_field_size.argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_set_field.argtypes = [C.c_void_p, C.c_int, _dp, C.c_size_t]
    lib.pcfd_get_field.argtypes = [C.c_void_p, C.c_int, _dp, C.c_size_t]
    lib.pcfd_field_device_ptr.restype = C.c_void_p
    lib.pcfd_field_device_ptr.argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.pcfd_set_cfl.argtypes = [C.c_void_p, C.c_double]
    lib.pcfd_set_time_integration.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.pcfd_set_gradient_type.argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_set_jacobian_type.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.pcfd_crs_sizes.argtypes = [C.c_void_p, _ip, _ip]
    lib.pcfd_get_crs.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    lib.pcfd_residual.argtypes = [C.c_void_p, _dp]
    lib.pcfd_timestep.argtypes = [C.c_void_p, _dp]
    lib.pcfd_sgs.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.pcfd_turb_compute.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.pcfd_chem_create.argtypes = [C.POINTER(ChemModelDesc), C.c_int, C.POINTER(C.c_void_p)]
    lib.pcfd_chem_destroy.argtypes = [C.c_void_p]
    lib.pcfd_chem_last_error.restype = C.c_char_p
    lib.pcfd_chem_last_error.argtypes = [C.c_void_p]
    lib.pcfd_chem_mass_production.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.pcfd_chem_source_term.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
    lib.pcfd_chem_source_term_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                                 C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.pcfd_create_fr.argtypes = [C.POINTER(MeshDesc), C.POINTER(Params), C.POINTER(FrParams), C.c_int,
                                   C.POINTER(C.c_void_p)]
    lib.pcfd_widths.argtypes = [C.c_void_p, _ip, _ip, _ip]
    lib.pcfd_limiter_raw.argtypes = [C.c_void_p]
    lib.pcfd_residual_fused.argtypes = [C.c_void_p, _dp, _ip]
    lib.pcfd_explicit_iterate.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.pcfd_implicit_iterate.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp]
    lib.pcfd_launch_count.restype = C.c_longlong
    lib.pcfd_launch_count.argtypes = [C.c_void_p]
    lib.pcfd_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_profile_reset.argtypes = [C.c_void_p]
    lib.pcfd_profile_count.argtypes = [C.c_void_p]
    lib.pcfd_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _dp, C.POINTER(C.c_longlong)]
    lib.pcfd_ipc_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.pcfd_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.pcfd_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    lib.pcfd_halo_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _ip]
    lib.pcfd_halo_width.argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_halo_send_total.argtypes = [C.c_void_p]
    lib.pcfd_halo_pack.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.pcfd_halo_recv_ptr.restype = C.c_void_p
    lib.pcfd_halo_recv_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.pcfd_comm_blob_size.restype = C.c_size_t
    lib.pcfd_comm_blob_size.argtypes = []
    lib.pcfd_comm_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.pcfd_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.pcfd_comm_disconnect.argtypes = [C.c_void_p]
    lib.pcfd_comm_connected.argtypes = [C.c_void_p]
    for name in ("pcfd_comm_post", "pcfd_comm_wait", "pcfd_comm_update"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
    lib.pcfd_comm_allgather.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
    lib.pcfd_turb_phase.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.pcfd_gmres.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp]
    lib.pcfd_wall_distance.argtypes = [C.c_void_p, _dp, C.c_int]
    lib.pcfd_crs_transpose.argtypes = [C.c_void_p]
    lib.pcfd_crs_ghost_blocks.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.pcfd_zeroed_updates.argtypes = [C.c_void_p]
    lib.pcfd_zeroed_updates.restype = C.c_longlong
    lib.pcfd_forces_configure.argtypes = [C.c_void_p, C.POINTER(ForcesDesc)]
    lib.pcfd_forces_areas.argtypes = [C.c_void_p, _dp, _dp]
    lib.pcfd_forces_compute.argtypes = [C.c_void_p, _dp, _dp]
    lib.pcfd_forces_get.argtypes = [C.c_void_p, C.c_int, _dp]
    for name in ("pcfd_destroy", "pcfd_synchronize", "pcfd_lsq_coefficients", "pcfd_update_bcs", "pcfd_gradient",
                 "pcfd_limiter", "pcfd_explicit_solve", "pcfd_jacobian", "pcfd_prepare_sgs", "pcfd_blank_x",
                 "pcfd_apply_dq"):
        getattr(lib, name).argtypes = [C.c_void_p]
    _libs[path] = lib
    if path == LIB_PATH:
        _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class PcfdError(RuntimeError):
    pass


class Context:
    """One hot-path context on one GPU (thin, 1:1 with the C ABI)."""

    def __init__(self, mesh, params, device=0, lib_path=None):
        """mesh: dict with the pcfd_mesh_desc arrays (+ counts); params: dict with the pcfd_params fields."""
        self.lib = load_library(lib_path) if lib_path else load_library()
        self._keep = {}
        md = MeshDesc()
        for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
            setattr(md, k, int(mesh[k]))
        for k, ctype in (("edges_n", np.int32), ("edges_a", np.float64), ("bedges_n", np.int32),
                         ("bedges_a", np.float64), ("bedges_bctype", np.int32), ("xyz", np.float64),
                         ("vol", np.float64), ("ipsp", np.int32), ("psp", np.int32)):
            a = np.ascontiguousarray(np.asarray(mesh[k]).reshape(-1), dtype=ctype)
            self._keep[k] = a
            setattr(md, k, _i(a) if ctype == np.int32 else _d(a))
        if mesh.get("bedges_twall") is not None:
            a = np.ascontiguousarray(np.asarray(mesh["bedges_twall"]).reshape(-1), dtype=np.float64)
            if a.size != md.nbedge:
                raise ValueError("bedges_twall must hold one value per BC half-edge")
            self._keep["bedges_twall"] = a
            md.bedges_twall = _d(a)
        pr = Params()
        pr.eqnset = int(params.get("eqnset", EQNSET_COMPRESSIBLE_EULER))
        pr.sorder, pr.limiter, pr.no_cvbc = int(params["sorder"]), int(params["limiter"]), int(params.get("no_cvbc", 0))
        pr.gamma, pr.chi, pr.cfl = float(params["gamma"]), float(params.get("chi", 0.0)), float(params["cfl"])
        fr = params.get("fr")      # reacting eqnset: dict(chem tables, reference values, qinf [nvars], ...)
        if fr is None:
            for j in range(NVARS):
                pr.qinf[j] = float(params["qinf"][j])
        pr.enable_vnn, pr.vnn = int(params.get("enable_vnn", 0)), float(params.get("vnn", 20.0))
        pr.Re, pr.Pr, pr.PrT = float(params.get("Re", 0.0)), float(params.get("Pr", 0.72)), float(params.get("PrT", 0.85))
        pr.tref, pr.mach = float(params.get("tref", 0.0)), float(params.get("mach", 0.0))
        pr.turb_model = int(params.get("turb_model", 0))
        self.nnode, self.gnode, self.nbnode = md.nnode, md.gnode, md.nbnode
        self.nedge, self.nbedge, self.ngedge = md.nedge, md.nbedge, md.ngedge
        h = C.c_void_p()
        if fr is not None:
            fp = FrParams()
            fill_chem_model(fp.chem, fr["chem"])
            for k in ("ref_density", "ref_velocity", "ref_temperature", "ref_pressure", "ref_time", "ref_specific_enthalpy",
                      "pref", "dt"):
                setattr(fp, k, float(fr[k]))
            fp.use_local_dt, fp.rxn_on = int(fr.get("use_local_dt", 1)), int(fr.get("rxn_on", 1))
            for j, v in enumerate(np.asarray(fr["qinf"], dtype=np.float64).reshape(-1)):
                fp.qinf[j] = float(v)
            pr.eqnset = EQNSET_COMPRESSIBLE_EULER_FR
            if fr.get("transport") is not None:     # compressibleNSFR: viscous terms with Wilke-mixed species transport
                fill_transport(fp.transport, fr["transport"])
                fp.ref_viscosity, fp.ref_k = float(fr["ref_viscosity"]), float(fr["ref_k"])
                pr.eqnset = EQNSET_COMPRESSIBLE_NS_FR
            rc = self.lib.pcfd_create_fr(C.byref(md), C.byref(pr), C.byref(fp), int(device), C.byref(h))
        else:
            rc = self.lib.pcfd_create(C.byref(md), C.byref(pr), int(device), C.byref(h))
        if rc != 0:
            raise PcfdError(self.lib.pcfd_last_error(None).decode())
        self.h = h
        ne, nv, nt = C.c_int(), C.c_int(), C.c_int()
        self.lib.pcfd_widths(self.h, C.byref(ne), C.byref(nv), C.byref(nt))
        self.neqn, self.nvars, self.nterms = ne.value, nv.value, nt.value
        self._keep.clear()   # the library has copied everything it needs

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcfd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PcfdError(self.lib.pcfd_last_error(self.h).decode())

    # -- data movement
    def field_size(self, field):
        return int(self.lib.pcfd_field_size(self.h, field))

    def set_field(self, field, host):
        host = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
        self._ck(self.lib.pcfd_set_field(self.h, field, _d(host), host.size))

    def get_field(self, field, out=None):
        if out is None:
            out = np.empty(self.field_size(field), dtype=np.float64)
        self._ck(self.lib.pcfd_get_field(self.h, field, _d(out), out.size))
        return out

    def device_ptr(self, field):
        return self.lib.pcfd_field_device_ptr(self.h, field)

    def get_crs(self):
        nrows, nblocks = C.c_int(), C.c_int()
        self._ck(self.lib.pcfd_crs_sizes(self.h, C.byref(nrows), C.byref(nblocks)))
        ia = np.empty(nrows.value + 1, np.int32)
        ja = np.empty(nblocks.value, np.int32)
        iau = np.empty(nrows.value, np.int32)
        pv = np.empty(nrows.value * self.neqn, np.int32)
        self._ck(self.lib.pcfd_get_crs(self.h, _i(ia), _i(ja), _i(iau), _i(pv)))
        return ia, ja, iau, pv

    def set_stream(self, cuda_stream):
        self._ck(self.lib.pcfd_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.lib.pcfd_synchronize(self.h))

    def set_cfl(self, cfl):
        self._ck(self.lib.pcfd_set_cfl(self.h, float(cfl)))

    def set_time_integration(self, dt, use_local_time_stepping=1, torder=1, it=1):
        """Param::dt / useLocalTimeStepping / torder and SolutionSpace::iter (see pcfd_set_time_integration)."""
        self._ck(self.lib.pcfd_set_time_integration(self.h, C.c_double(float(dt)), int(use_local_time_stepping), int(torder),
                                                    int(it)))

    def set_gradient_type(self, grad_type):
        """Param::gradType: 0 weighted least squares, 1 Green-Gauss (gradient.tcc:68-90)."""
        self._ck(self.lib.pcfd_set_gradient_type(self.h, int(grad_type)))

    def set_jacobian_type(self, field_type, boundary_type):
        """Param::fieldJacType / boundaryJacType: 0 one-sided, 1 central differences (jacobian.tcc:140-176)."""
        self._ck(self.lib.pcfd_set_jacobian_type(self.h, int(field_type), int(boundary_type)))

    def clip_fallbacks(self):
        """times the fused limiter / residual pair fell back to the ordered pressure-clip path"""
        self.lib.pcfd_clip_fallbacks.restype = C.c_longlong
        return int(self.lib.pcfd_clip_fallbacks(self.h))

    def launch_count(self):
        return int(self.lib.pcfd_launch_count(self.h))

    # -- halo (PObj)
    def halo_configure(self, rank, nranks, send_counts, send_list, recv_counts):
        sc = np.ascontiguousarray(send_counts, dtype=np.int32)
        sl = np.ascontiguousarray(send_list, dtype=np.int32)
        rc = np.ascontiguousarray(recv_counts, dtype=np.int32)
        if sl.size == 0:
            sl = np.zeros(1, np.int32)
        self._ck(self.lib.pcfd_halo_configure(self.h, int(rank), int(nranks), _i(sc), _i(sl), _i(rc)))

    def halo_width(self, field):
        return int(self.lib.pcfd_halo_width(self.h, field))

    def halo_pack(self, field, dst_ptr, peer=-1):
        self._ck(self.lib.pcfd_halo_pack(self.h, field, int(peer), C.c_void_p(dst_ptr)))

    def halo_recv_ptr(self, field, peer=-1):
        return self.lib.pcfd_halo_recv_ptr(self.h, field, int(peer))

    def ipc_export(self, field):
        buf = C.create_string_buffer(64)
        self._ck(self.lib.pcfd_ipc_export(self.h, field, buf))
        return buf.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._ck(self.lib.pcfd_ipc_open(self.h, C.c_char_p(handle), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._ck(self.lib.pcfd_ipc_close(self.h, C.c_void_p(ptr)))

    # -- collective-free exchange (pcfd_comm_*): blobs travel through whatever the host has
    def comm_export(self):
        buf = C.create_string_buffer(int(self.lib.pcfd_comm_blob_size()))
        self._ck(self.lib.pcfd_comm_export(self.h, buf))
        return buf.raw

    def comm_connect(self, blobs):
        """blobs: every rank's comm_export() in rank order"""
        joined = b"".join(blobs)
        self._ck(self.lib.pcfd_comm_connect(self.h, C.c_char_p(joined)))

    def comm_debug_flags(self, nranks):
        out = (C.c_ulonglong * (2 * nranks + 1))()
        self.lib.pcfd_comm_debug_flags.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.pcfd_comm_debug_flags(self.h, out))
        v = list(out)
        return dict(ready=v[:nranks], done=v[nranks:2 * nranks], err=v[2 * nranks])

    def comm_disconnect(self):
        self._ck(self.lib.pcfd_comm_disconnect(self.h))

    def comm_connected(self):
        return bool(self.lib.pcfd_comm_connected(self.h))

    def comm_post(self, field):
        self._ck(self.lib.pcfd_comm_post(self.h, field))

    def comm_wait(self, field):
        self._ck(self.lib.pcfd_comm_wait(self.h, field))

    def comm_update(self, field):
        self._ck(self.lib.pcfd_comm_update(self.h, field))

    def comm_allgather(self, vals, nranks):
        v = np.ascontiguousarray(vals, dtype=np.float64).reshape(-1)
        out = np.zeros(nranks * v.size)
        self._ck(self.lib.pcfd_comm_allgather(self.h, _d(v), int(v.size), _d(out)))
        return out.reshape(nranks, v.size)

    def profile(self, on=True, reset=False):
        if reset:
            self._ck(self.lib.pcfd_profile_reset(self.h))
        self._ck(self.lib.pcfd_profile_enable(self.h, int(on)))

    def profile_table(self):
        """{kernel name: (total ms, launches)} accumulated since the last reset."""
        out = {}
        for i in range(self.lib.pcfd_profile_count(self.h)):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_longlong()
            self._ck(self.lib.pcfd_profile_get(self.h, i, C.byref(name), C.byref(ms), C.byref(n)))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    # -- phases (names follow the reference functions they replace)
    def lsq_coefficients(self):
        self._ck(self.lib.pcfd_lsq_coefficients(self.h))

    def update_bcs(self):
        self._ck(self.lib.pcfd_update_bcs(self.h))

    def gradient(self):
        self._ck(self.lib.pcfd_gradient(self.h))

    def limiter(self):
        self._ck(self.lib.pcfd_limiter(self.h))

    def limiter_raw(self):
        self._ck(self.lib.pcfd_limiter_raw(self.h))

    def residual_fused(self, want_norms=False):
        """(norms or None, clip_hit): see pcfd_residual_fused."""
        hit = C.c_int(0)
        s = np.zeros(1 + self.neqn) if want_norms else None
        self._ck(self.lib.pcfd_residual_fused(self.h, _d(s) if want_norms else None, C.byref(hit)))
        return s, bool(hit.value)

    def residual(self, want_norms=False):
        if not want_norms:
            self._ck(self.lib.pcfd_residual(self.h, None))
            return None
        s = np.zeros(1 + self.neqn)
        self._ck(self.lib.pcfd_residual(self.h, _d(s)))
        return s

    def timestep(self, want_min=True):
        if not want_min:
            self._ck(self.lib.pcfd_timestep(self.h, None))
            return None
        d = C.c_double()
        self._ck(self.lib.pcfd_timestep(self.h, C.byref(d)))
        return d.value

    def explicit_solve(self):
        self._ck(self.lib.pcfd_explicit_solve(self.h))

    def jacobian(self):
        self._ck(self.lib.pcfd_jacobian(self.h))

    def prepare_sgs(self):
        self._ck(self.lib.pcfd_prepare_sgs(self.h))

    def blank_x(self):
        self._ck(self.lib.pcfd_blank_x(self.h))

    def sgs(self, nsgs, want_ddq=True):
        if not want_ddq:
            self._ck(self.lib.pcfd_sgs(self.h, int(nsgs), None))
            return None
        d = C.c_double()
        self._ck(self.lib.pcfd_sgs(self.h, int(nsgs), C.byref(d)))
        return d.value

    def gmres(self, restarts, nsearch, precond_type):
        """CRS::GMRES on the context's A (assembled, not factored), b and x; returns the reference's dqNorm"""
        d = C.c_double()
        self._ck(self.lib.pcfd_gmres(self.h, int(restarts), int(nsearch), int(precond_type), C.byref(d)))
        return d.value

    def crs_transpose(self):
        """CRSMatrix::CRSTranspose on the assembled A: the local part (blocks transposed, mirror blocks of local node pairs
        swapped, ghost-column blocks transposed in place); on a partition follow with parallel.crs_transpose_ghosts"""
        self._ck(self.lib.pcfd_crs_transpose(self.h))

    def get_ghost_blocks(self):
        """the blocks of the ghost columns, [ngedge, neqn, neqn], in the order of the parallel half-edges"""
        neqn = self.neqn
        out = np.zeros((max(self.ngedge, 1), neqn, neqn))
        self._ck(self.lib.pcfd_crs_ghost_blocks(self.h, 0, out.ctypes.data_as(_dp)))
        return out[: self.ngedge]

    def set_ghost_blocks(self, blocks):
        neqn = self.neqn
        b = np.ascontiguousarray(blocks, dtype=np.float64)
        if b.size != self.ngedge * neqn * neqn:
            raise ValueError("set_ghost_blocks: ngedge * neqn^2 doubles expected")
        self._ck(self.lib.pcfd_crs_ghost_blocks(self.h, 1, b.ctypes.data_as(_dp)))

    def wall_distance(self, points):
        """ComputeWallDistOct: field F_WALLDIST = distance of every local node to the nearest of `points` [n, 3] (the viscous
        wall nodes of all ranks)"""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.pcfd_wall_distance(self.h, pts.ctypes.data_as(_dp), int(pts.shape[0])))

    def zeroed_updates(self):
        """nodes whose NaN / Inf update apply_dq zeroed (NewtonIterate, solutionSpace.tcc:771-796)"""
        return int(self.lib.pcfd_zeroed_updates(self.h))

    def forces_configure(self, body_offsets, body_factags, moment_pt, moment_axis, bedges_factag, cg, liftdir, dragdir,
                         velocity, num_bcs):
        """Composite bodies + surface data for Forces (pcfd_forces_configure); also evaluates ComputeSurfaceAreas."""
        keep = [np.ascontiguousarray(body_offsets, dtype=np.int32), np.ascontiguousarray(body_factags, dtype=np.int32),
                np.ascontiguousarray(moment_pt, dtype=np.float64), np.ascontiguousarray(moment_axis, dtype=np.float64),
                np.ascontiguousarray(bedges_factag, dtype=np.int32), np.ascontiguousarray(cg, dtype=np.float64)]
        d = ForcesDesc()
        d.nbodies, d.num_bcs = keep[0].size - 1, int(num_bcs)
        d.body_offsets, d.body_factags = keep[0].ctypes.data_as(_ip), keep[1].ctypes.data_as(_ip)
        d.moment_pt, d.moment_axis = keep[2].ctypes.data_as(_dp), keep[3].ctypes.data_as(_dp)
        d.bedges_factag, d.cg = keep[4].ctypes.data_as(_ip), keep[5].ctypes.data_as(_dp)
        for j in range(3):
            d.liftdir[j], d.dragdir[j] = float(liftdir[j]), float(dragdir[j])
        d.velocity = float(velocity)
        if keep[4].size < self.nbedge or keep[5].size < 3 * (self.nnode + self.gnode + self.nbnode):
            raise ValueError("forces_configure: bedges_factag needs nbedge entries, cg (nnode+gnode+nbnode)*3")
        self._ck(self.lib.pcfd_forces_configure(self.h, C.byref(d)))
        self._forces_shape = (d.nbodies, d.num_bcs)

    def forces_areas(self):
        nb, nbc = self._forces_shape
        sa, ba = np.zeros(3 * (nbc + 1)), np.zeros(3 * nb)
        self._ck(self.lib.pcfd_forces_areas(self.h, sa.ctypes.data_as(_dp), ba.ctypes.data_as(_dp)))
        return sa, ba

    def forces_compute(self):
        """Forces::Compute: (body [nbodies, 12] = forces, vforces, moments, vmoments; coef [nbodies, 3] = cl, cd, cm)"""
        nb, _ = self._forces_shape
        body, coef = np.zeros(12 * nb), np.zeros(3 * nb)
        self._ck(self.lib.pcfd_forces_compute(self.h, body.ctypes.data_as(_dp), coef.ctypes.data_as(_dp)))
        return body.reshape(nb, 12), coef.reshape(nb, 3)

    def forces_get(self, which):
        out = np.zeros(max(self.nbedge, 1))
        self._ck(self.lib.pcfd_forces_get(self.h, int(which), out.ctypes.data_as(_dp)))
        return out[: self.nbedge]

    def turb_compute(self, nsgs, want_norm=False):
        """TurbulenceModel::Compute (Spalart-Allmaras); returns sum(b^2) of the turbulence residual when asked."""
        if not want_norm:
            self._ck(self.lib.pcfd_turb_compute(self.h, int(nsgs), None))
            return None
        d = C.c_double()
        self._ck(self.lib.pcfd_turb_compute(self.h, int(nsgs), C.byref(d)))
        return d.value

    def turb_phase(self, phase, want_norm=False):
        """one phase of TurbulenceModel::Compute between the reference's exchange points (see pcfd_turb_phase)"""
        if not want_norm:
            self._ck(self.lib.pcfd_turb_phase(self.h, int(phase), None))
            return None
        d = C.c_double()
        self._ck(self.lib.pcfd_turb_phase(self.h, int(phase), C.byref(d)))
        return d.value

    def apply_dq(self):
        self._ck(self.lib.pcfd_apply_dq(self.h))

    def explicit_iterate(self, refresh_dt=True, want_norms=False):
        s = np.zeros(1 + self.neqn) if want_norms else None
        self._ck(self.lib.pcfd_explicit_iterate(self.h, int(refresh_dt), _d(s) if want_norms else None))
        return s

    def implicit_iterate(self, nsgs, refresh_jac=True, want_norms=False):
        s = np.zeros(1 + self.neqn) if want_norms else None
        self._ck(self.lib.pcfd_implicit_iterate(self.h, int(refresh_jac), int(nsgs), _d(s) if want_norms else None, None))
        return s


class Chem:
    """Finite-rate chemistry handle (pcfd_chem): the compressibleFR source term."""

    def __init__(self, tables, device=0):
        self.lib = load_library()
        md = fill_chem_model(ChemModelDesc(), tables)
        self.ns, self.nr = md.nspecies, md.nreactions
        h = C.c_void_p()
        if self.lib.pcfd_chem_create(C.byref(md), int(device), C.byref(h)) != 0:
            raise PcfdError(self.lib.pcfd_chem_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcfd_chem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise PcfdError(self.lib.pcfd_chem_last_error(self.h).decode())

    def mass_production(self, rhoi, T):
        """ChemModel::GetMassProductionRates: rhoi [n, ns] kg/m^3, T [n] K -> wdot [n, ns] kg/(m^3 s)."""
        rhoi = np.ascontiguousarray(rhoi, dtype=np.float64).reshape(-1, self.ns)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(-1)
        w = np.empty_like(rhoi)
        self._ck(self.lib.pcfd_chem_mass_production(self.h, len(T), _d(rhoi), _d(T), _d(w)))
        return w

    def source_term(self, Q, vol, ref_density, ref_time, ref_temperature):
        """CompressibleFREqnSet::SourceTerm on host arrays: Q [n, stride], vol [n] -> source [n, ns + 4]."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        vol = np.ascontiguousarray(vol, dtype=np.float64).reshape(-1)
        src = np.empty((len(vol), self.ns + 4))
        self._ck(self.lib.pcfd_chem_source_term(self.h, len(vol), Q.shape[1], _d(Q), _d(vol), float(ref_density),
                                                float(ref_time), float(ref_temperature), _d(src)))
        return src

    def source_term_device(self, n, stride, d_Q, d_vol, ref_density, ref_time, ref_temperature, d_source, stream=0):
        self._ck(self.lib.pcfd_chem_source_term_device(self.h, int(n), int(stride), C.c_void_p(d_Q), C.c_void_p(d_vol),
                                                       float(ref_density), float(ref_time), float(ref_temperature),
                                                       C.c_void_p(d_source), C.c_void_p(stream)))

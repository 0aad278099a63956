"""ctypes wrapper around oracle/libpcfd_oracle.so (the plain-C restatement).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

NEQN, NVARS, NTERMS = 5, 10, 9
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcCase(C.Structure):
    _fields_ = [("nnode", C.c_int), ("gnode", C.c_int), ("nbnode", C.c_int),
                ("nedge", C.c_int), ("nbedge", C.c_int), ("ngedge", C.c_int),
                ("edges_n", _ip), ("edges_a", _dp), ("bedges_n", _ip), ("bedges_a", _dp),
                ("bedges_bctype", _ip), ("xyz", _dp), ("vol", _dp), ("ipsp", _ip), ("psp", _ip),
                ("gamma", C.c_double), ("chi", C.c_double), ("cfl", C.c_double),
                ("limiter", C.c_int), ("sorder", C.c_int), ("no_cvbc", C.c_int),
                ("qinf", C.c_double * NVARS),
                ("viscous", C.c_int), ("enable_vnn", C.c_int),
                ("Re", C.c_double), ("Pr", C.c_double), ("PrT", C.c_double), ("tref", C.c_double),
                ("mach", C.c_double), ("vnn", C.c_double), ("bedges_twall", _dp), ("mut", _dp),
                ("dt_param", C.c_double), ("use_local_dt", C.c_int), ("torder", C.c_int), ("iter", C.c_int),
                ("qold", _dp), ("qoldm1", _dp), ("walldist", _dp), ("field_jac_type", C.c_int), ("boundary_jac_type", C.c_int),
                ("grad_type", C.c_int)]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    meta = dict(zip([str(k) for k in d.pop("meta_keys")], d.pop("meta_vals")))
    return d, meta


class OrcForcesDesc(C.Structure):
    """orc_forces_desc (oracle/pcfd_oracle.h)"""
    _fields_ = [("nbodies", C.c_int), ("num_bcs", C.c_int), ("body_offsets", _ip), ("body_factags", _ip),
                ("moment_pt", _dp), ("moment_axis", _dp), ("bedges_factag", _ip), ("cg", _dp),
                ("liftdir", C.c_double * 3), ("dragdir", C.c_double * 3)]


def bodies_from_fixture(g):
    """The composite bodies of a forces fixture in the flat form of the C ABI: (offsets, factags, moment_pt, moment_axis)."""
    lists = g["forces_body_lists"]
    offs, tags, i = [0], [], 0
    while i < lists.size:
        n = int(lists[i])
        tags += [int(t) for t in lists[i + 1: i + 1 + n]]
        offs.append(len(tags))
        i += 1 + n
    geom = g["forces_body_geom"].reshape(-1, 6)
    return (np.array(offs, dtype=np.int32), np.array(tags, dtype=np.int32), np.ascontiguousarray(geom[:, :3]).ravel(),
            np.ascontiguousarray(geom[:, 3:]).ravel())


class ForcesMixin:
    """orc_surface_areas / orc_forces* on a forces fixture (keys forces_*)."""

    def forces_desc(self, g):
        offs, tags, mpt, max_ = bodies_from_fixture(g)
        keep = dict(offs=offs, tags=tags, mpt=mpt, max=max_, factag=np.ascontiguousarray(g["bedges_factag"], dtype=np.int32),
                    cg=np.ascontiguousarray(g["forces_cg"]))
        d = OrcForcesDesc()
        d.nbodies, d.num_bcs = offs.size - 1, g["forces_surfArea"].size // 3 - 1
        d.body_offsets, d.body_factags = _i(offs), _i(tags)
        d.moment_pt, d.moment_axis = _d(mpt), _d(max_)
        d.bedges_factag, d.cg = _i(keep["factag"]), _d(keep["cg"])
        for j in range(3):
            d.liftdir[j], d.dragdir[j] = g["forces_dirs"][j], g["forces_dirs"][3 + j]
        self._forces_keep = keep
        return d

    def surface_areas(self, d):
        sa, ba = np.zeros(3 * (d.num_bcs + 1)), np.zeros(3 * d.nbodies)
        self.lib.orc_surface_areas(C.byref(self.c), C.byref(d), _d(sa), _d(ba))
        return sa, ba

    def _forces_out(self, d):
        nbe = self.c.nbedge
        return dict(cp=np.zeros(nbe), yp=np.zeros(nbe), cf=np.zeros(nbe), body=np.zeros(12 * d.nbodies), coef=np.zeros(3 * d.nbodies))


class Oracle(ForcesMixin):
    """One mesh + parameter set bound to the C oracle."""

    def forces(self, d, q, qgrad, body_area):
        o = self._forces_out(d)
        self.lib.orc_forces(C.byref(self.c), C.byref(d), _d(q), _d(qgrad), _d(body_area), _d(o["cp"]), _d(o["yp"]), _d(o["cf"]),
                            _d(o["body"]), _d(o["coef"]))
        return o

    def __init__(self, lib, g, meta):
        self.lib = lib
        self.keep = {k: np.ascontiguousarray(v) for k, v in g.items()}
        k = self.keep
        c = OrcCase()
        for f in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge", "limiter", "sorder", "no_cvbc"):
            setattr(c, f, int(meta[f]))
        c.gamma, c.chi, c.cfl = meta["gamma"], meta["chi"], meta["cfl"]
        c.edges_n, c.edges_a = _i(k["edges_n"]), _d(k["edges_a"])
        c.bedges_n, c.bedges_a, c.bedges_bctype = _i(k["bedges_n"]), _d(k["bedges_a"]), _i(k["bedges_bctype"])
        c.xyz, c.vol, c.ipsp, c.psp = _d(k["xyz"]), _d(k["vol"]), _i(k["ipsp"]), _i(k["psp"])
        for j in range(NVARS):
            c.qinf[j] = k["qinf"][j]
        # viscous terms (compressibleNS); absent keys mean an inviscid case
        c.viscous, c.enable_vnn = int(meta.get("viscous", 0)), int(meta.get("enableVNN", 0))
        c.Re, c.Pr, c.PrT = meta.get("Re", 1.0), meta.get("Pr", 0.72), meta.get("PrT", 0.85)
        c.tref, c.mach, c.vnn = meta.get("ref_temperature", 300.0), meta.get("velocity", 0.0), meta.get("VNN", 20.0)
        c.bedges_twall = _d(k["bedges_twall"]) if "bedges_twall" in k and k["bedges_twall"].size else None
        c.mut = _d(k["mut"]) if "mut" in k else None
        # time integration: the BDF terms are live only when the fixture / caller carries a q^n that differs from q
        c.dt_param, c.use_local_dt = float(meta.get("dt", -1.0)), int(meta.get("useLocalTimeStepping", 1))
        c.torder, c.iter = int(meta.get("torder", 1)), int(meta.get("iter", 1))
        unsteady = float(meta.get("dt", -1.0)) > 0.0 and "qold" in k
        c.qold = _d(k["qold"]) if unsteady else None
        c.qoldm1 = _d(k["qoldm1"]) if unsteady else None
        c.walldist = _d(k["wallDistance"]) if "wallDistance" in k else None
        c.field_jac_type, c.boundary_jac_type = int(meta.get("fieldJacType", 0)), int(meta.get("boundaryJacType", 0))
        c.grad_type = int(meta.get("gradType", 0))
        self.c = c
        self.nn = c.nnode + c.gnode
        self.nnode = c.nnode
        self.nblocks = int(k["ipsp"][c.nnode]) + c.nnode

    def lsq(self):
        s, sw = np.zeros(self.nn * 6), np.zeros(self.nn * 6)
        self.lib.orc_lsq_coefficients(C.byref(self.c), _d(s), _d(sw))
        return s, sw

    def gradient(self, q, sw):
        g = np.zeros(self.nn * NTERMS * 3)
        self.lib.orc_gradient(C.byref(self.c), _d(q), _d(sw), _d(g))
        return g

    def limiter(self, q, grad):
        lim = np.zeros(self.nn * NEQN)
        self.lib.orc_limiter(C.byref(self.c), _d(q), _d(grad), _d(lim))
        return lim

    def update_bcs(self, q, beta):
        self.lib.orc_update_bcs(C.byref(self.c), _d(q), _d(beta))

    def residual(self, q, grad, lim, beta):
        b = np.zeros(self.nnode * NEQN)
        self.lib.orc_residual(C.byref(self.c), _d(q), _d(grad), _d(lim), _d(beta), _d(b))
        return b

    def timestep(self, q, beta):
        dt = np.zeros(self.nnode)
        self.lib.orc_timestep.restype = C.c_double
        dtmin = self.lib.orc_timestep(C.byref(self.c), _d(q), _d(beta), _d(dt))
        return dt, dtmin

    def explicit_solve(self, q, b, dt):
        x = np.zeros(self.nnode * NEQN)
        self.lib.orc_explicit_solve(C.byref(self.c), _d(q), _d(b), _d(dt), _d(x))
        return x

    def apply_dq(self, q, x):
        self.lib.orc_apply_dq(C.byref(self.c), _d(q), _d(x))

    def crs_init(self):
        ia = np.zeros(self.nnode + 1, dtype=np.int32)
        ja = np.zeros(self.nblocks, dtype=np.int32)
        iau = np.zeros(self.nnode, dtype=np.int32)
        self.lib.orc_crs_init(C.byref(self.c), _i(ia), _i(ja), _i(iau))
        return ia, ja, iau

    def jacobian(self, q, beta, dt, ia, ja, iau):
        A = np.zeros(self.nblocks * 25)
        self.lib.orc_jacobian(C.byref(self.c), _d(q), _d(beta), _d(dt), _i(ia), _i(ja), _i(iau), _d(A))
        return A

    def prepare_sgs(self, iau, A):
        pv = np.zeros(self.nnode * NEQN, dtype=np.int32)
        self.lib.orc_prepare_sgs(C.byref(self.c), _i(iau), _d(A), _i(pv))
        return pv

    def sgs(self, nsgs, ia, ja, iau, A, pv, b):
        x = np.zeros(self.nn * NEQN)
        self.lib.orc_sgs.restype = C.c_double
        d = self.lib.orc_sgs(C.byref(self.c), nsgs, _i(ia), _i(ja), _i(iau), _d(A), _i(pv), _d(b), _d(x))
        return x, d


    def turb_sa(self, nsgs, q, qgrad, s, dist, dt, ia, ja, iau, tvar):
        """One TurbulenceModel::Compute of the Spalart-Allmaras model; tvar is updated in place."""
        out = dict(tgrad=np.zeros(self.nn * 3), b=np.zeros(self.nnode), A=np.zeros(self.nblocks), x=np.zeros(self.nn),
                   mut=np.zeros(self.nn))
        self.lib.orc_turb_sa.restype = C.c_double
        out["res"] = self.lib.orc_turb_sa(C.byref(self.c), int(nsgs), _d(q), _d(qgrad), _d(s), _d(dist), _d(dt), _i(ia), _i(ja),
                                          _i(iau), _d(tvar), _d(out["tgrad"]), _d(out["b"]), _d(out["A"]), _d(out["x"]),
                                          _d(out["mut"]))
        return out


    def turb_sa_phase(self, phase, nsgs, q, qgrad, s, dist, dt, ia, ja, iau, st):
        """One phase (0..5) of the same update on the state dict `st` (tvar, tgrad, b, A, x, mut: allocated by
        turb_sa_state); phase 2 returns the sum of b^2."""
        self.lib.orc_turb_sa_phase.restype = C.c_double
        return self.lib.orc_turb_sa_phase(C.byref(self.c), int(phase), int(nsgs), _d(q), _d(qgrad), _d(s), _d(dist), _d(dt),
                                          _i(ia), _i(ja), _i(iau), _d(st["tvar"]), _d(st["tgrad"]), _d(st["b"]), _d(st["A"]),
                                          _d(st["x"]), _d(st["mut"]))

    def turb_sa_state(self, tvar):
        return dict(tvar=tvar, tgrad=np.zeros(self.nn * 3), b=np.zeros(self.nnode), A=np.zeros(self.nblocks),
                    x=np.zeros(self.nn), mut=np.zeros(self.nn))


def oracle_for(lib, mesh, params):
    """Bind the C oracle to a mesh/params pair in the C-ABI dict form (proteuscfd_b200.cases)."""
    g = {k: np.asarray(mesh[k]) for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol",
                                          "ipsp", "psp")}
    g["qinf"] = np.asarray(params["qinf"])
    if mesh.get("bedges_twall") is not None:
        g["bedges_twall"] = np.asarray(mesh["bedges_twall"], dtype=np.float64)
    meta = {k: mesh[k] for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge")}
    meta.update(limiter=params["limiter"], sorder=params["sorder"], no_cvbc=params["no_cvbc"], gamma=params["gamma"],
                chi=params["chi"], cfl=params["cfl"])
    if params.get("viscous"):
        meta.update(viscous=1, Re=params["Re"], Pr=params["Pr"], PrT=params["PrT"], ref_temperature=params["tref"],
                    velocity=params["mach"], enableVNN=params.get("enable_vnn", 0), VNN=params.get("vnn", 20.0))
    return Oracle(lib, g, meta)


def load_oracle():
    so = os.path.join(ORACLE_DIR, "libpcfd_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("pcfd_oracle.c", "pcfd_oracle_fr.c", "pcfd_oracle.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return C.CDLL(so)


class OrcChemModel(C.Structure):
    """orc_chem_model (oracle/pcfd_oracle.h) -- same layout as the product's pcfd_chem_model."""
    _S, _R = 16, 32
    _fields_ = [("nspecies", C.c_int), ("nreactions", C.c_int),
                ("mw", C.c_double * _S), ("nasa7", C.c_double * 7 * 2 * _S),
                ("rxn_type", C.c_int * _R), ("third_body", C.c_int * _R), ("backward_given", C.c_int * _R),
                ("rxn_type_b", C.c_int * _R), ("nsp", C.c_int * _R), ("species", C.c_int * _S * _R),
                ("A", C.c_double * _R), ("EA", C.c_double * _R), ("n", C.c_double * _R),
                ("Ab", C.c_double * _R), ("EAb", C.c_double * _R), ("nb", C.c_double * _R),
                ("nup", C.c_double * _S * _R), ("nupp", C.c_double * _S * _R), ("tbeff", C.c_double * _S * _R)]


class ChemOracle:
    def __init__(self, lib, tables):
        from proteuscfd_b200.capi import fill_chem_model   # a pure-Python table filler, no compute
        self.lib = lib
        self.m = fill_chem_model(OrcChemModel(), tables)
        self.ns, self.nr = self.m.nspecies, self.m.nreactions

    def mass_production(self, rhoi, T):
        rhoi = np.ascontiguousarray(rhoi, dtype=np.float64).reshape(-1, self.ns)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(-1)
        n = len(T)
        w, sc = np.empty_like(rhoi), np.empty_like(rhoi)
        kf, kb = np.empty((n, self.nr)), np.empty((n, self.nr))
        self.lib.orc_chem_mass_production(C.byref(self.m), n, _d(rhoi), _d(T), _d(w), _d(sc), _d(kf), _d(kb))
        return w, sc, kf, kb

    def source_term(self, Q, vol, ref_density, ref_time, ref_temperature):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        vol = np.ascontiguousarray(vol, dtype=np.float64).reshape(-1)
        src = np.empty((len(vol), self.ns + 4))
        self.lib.orc_chem_source_term.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double,
                                                  C.c_double, _dp]
        self.lib.orc_chem_source_term(C.byref(self.m), len(vol), Q.shape[1], _d(Q), _d(vol), ref_density, ref_time,
                                      ref_temperature, _d(src))
        return src


class OrcTransport(C.Structure):
    """orc_transport (oracle/pcfd_oracle.h)."""
    _S = 16
    _fields_ = [("nmu", C.c_int * _S), ("nk", C.c_int * _S),
                ("mu_fit", C.c_double * 6 * 3 * _S), ("k_fit", C.c_double * 6 * 3 * _S),
                ("mu_white", C.c_double * 4 * _S), ("k_white", C.c_double * 4 * _S)]


class OrcFrParams(C.Structure):
    """orc_fr_params (oracle/pcfd_oracle.h)."""
    _fields_ = [("chem", C.POINTER(OrcChemModel)),
                ("ref_density", C.c_double), ("ref_velocity", C.c_double), ("ref_temperature", C.c_double),
                ("ref_pressure", C.c_double), ("ref_time", C.c_double), ("ref_specific_enthalpy", C.c_double),
                ("Pref", C.c_double), ("dt_param", C.c_double), ("use_local_dt", C.c_int), ("rxn_on", C.c_int),
                ("qinf", C.c_double * (3 * 16 + 6)),
                ("transport", C.POINTER(OrcTransport)), ("ref_viscosity", C.c_double), ("ref_k", C.c_double)]


def chem_tables(g):
    """The chemistry tables of an FR fixture in fill_chem_model's form."""
    t = {k: g[k] for k in ("species_mw", "species_nasa7", "rxn_A_EA_n", "rxn_flags", "rxn_species", "rxn_nup", "rxn_nupp",
                           "rxn_tbeff")}
    t["dims"] = g["chem_dims"]
    return t


class FrOracle(Oracle):
    """The reacting eqnset (oracle/pcfd_oracle_fr.c) on one mesh + parameter set."""

    def __init__(self, lib, g, meta):
        g = dict(g)
        qinf = np.asarray(g["qinf"], dtype=np.float64)
        g["qinf"] = np.zeros(NVARS)     # orc_case.qinf belongs to the perfect-gas eqnset
        super().__init__(lib, g, meta)
        from proteuscfd_b200.capi import fill_chem_model
        self.chem = fill_chem_model(OrcChemModel(), chem_tables(g))
        self.ns = self.chem.nspecies
        self.neqn, self.nvars, self.nterms = self.ns + 4, 3 * self.ns + 6, 2 * self.ns + 4
        p = OrcFrParams()
        p.chem = C.pointer(self.chem)
        for f in ("ref_density", "ref_velocity", "ref_temperature", "ref_pressure", "ref_time", "ref_specific_enthalpy",
                  "Pref"):
            setattr(p, f, float(meta[f]))
        p.dt_param, p.use_local_dt, p.rxn_on = float(meta["dt"]), int(meta["useLocalTimeStepping"]), int(meta["rxnOn"])
        for j in range(self.nvars):
            p.qinf[j] = qinf[j]
        if int(meta.get("viscous", 0)):
            from proteuscfd_b200.capi import fill_transport   # a pure-Python table filler, no compute
            self.transport = fill_transport(OrcTransport(), g)
            p.transport = C.pointer(self.transport)
            p.ref_viscosity, p.ref_k = float(meta["ref_viscosity"]), float(meta["ref_k"])
        self.p = p
        lib.orc_fr_timestep.restype = C.c_double
        lib.orc_fr_sgs.restype = C.c_double

    def gradient(self, q, sw):
        g = np.zeros(self.nn * self.nterms * 3)
        self.lib.orc_fr_gradient(C.byref(self.c), C.byref(self.p), _d(q), _d(sw), _d(g))
        return g

    def limiter(self, q, grad):
        lim = np.zeros(self.nn * self.neqn)
        self.lib.orc_fr_limiter(C.byref(self.c), C.byref(self.p), _d(q), _d(grad), _d(lim))
        return lim

    def update_bcs(self, q, beta):
        self.lib.orc_fr_update_bcs(C.byref(self.c), C.byref(self.p), _d(q), _d(beta))

    def residual(self, q, grad, lim, beta):
        b = np.zeros(self.nnode * self.neqn)
        self.lib.orc_fr_residual(C.byref(self.c), C.byref(self.p), _d(q), _d(grad), _d(lim), _d(beta), _d(b))
        return b

    def timestep(self, q, beta):
        dt = np.zeros(self.nnode)
        dtmin = self.lib.orc_fr_timestep(C.byref(self.c), C.byref(self.p), _d(q), _d(beta), _d(dt))
        return dt, dtmin

    def explicit_solve(self, q, b, dt):
        x = np.zeros(self.nnode * self.neqn)
        bad = self.lib.orc_fr_explicit_solve(C.byref(self.c), C.byref(self.p), _d(q), _d(b), _d(dt), _d(x))
        assert bad == 0, "ConservativeToNative did not converge"
        return x

    def apply_dq(self, q, x):
        self.lib.orc_fr_apply_dq(C.byref(self.c), C.byref(self.p), _d(q), _d(x))

    def jacobian(self, q, beta, dt, ia, ja, iau):
        A = np.zeros(self.nblocks * self.neqn * self.neqn)
        self.lib.orc_fr_jacobian(C.byref(self.c), C.byref(self.p), _d(q), _d(beta), _d(dt), _i(ia), _i(ja), _i(iau), _d(A))
        return A

    def prepare_sgs(self, iau, A):
        pv = np.zeros(self.nnode * self.neqn, dtype=np.int32)
        self.lib.orc_fr_prepare_sgs(C.byref(self.c), C.byref(self.p), _i(iau), _d(A), _i(pv))
        return pv

    def forces(self, d, q, qgrad, body_area, V):
        o = self._forces_out(d)
        self.lib.orc_fr_forces(C.byref(self.c), C.byref(self.p), C.byref(d), C.c_double(V), _d(q), _d(qgrad), _d(body_area),
                               _d(o["cp"]), _d(o["yp"]), _d(o["cf"]), _d(o["body"]), _d(o["coef"]))
        return o

    def turb_sa(self, nsgs, q, qgrad, s, dist, dt, ia, ja, iau, tvar):
        """TurbulenceModel::Compute of the Spalart-Allmaras model under compressibleNSFR; tvar is updated in place."""
        out = dict(tgrad=np.zeros(self.nn * 3), b=np.zeros(self.nnode), A=np.zeros(self.nblocks), x=np.zeros(self.nn),
                   mut=np.zeros(self.nn))
        self.lib.orc_fr_turb_sa.restype = C.c_double
        out["res"] = self.lib.orc_fr_turb_sa(C.byref(self.c), C.byref(self.p), int(nsgs), _d(q), _d(qgrad), _d(s), _d(dist),
                                             _d(dt), _i(ia), _i(ja), _i(iau), _d(tvar), _d(out["tgrad"]), _d(out["b"]),
                                             _d(out["A"]), _d(out["x"]), _d(out["mut"]))
        return out

    def sgs(self, nsgs, ia, ja, iau, A, pv, b):
        x = np.zeros(self.nn * self.neqn)
        d = self.lib.orc_fr_sgs(C.byref(self.c), C.byref(self.p), nsgs, _i(ia), _i(ja), _i(iau), _d(A), _i(pv), _d(b), _d(x))
        return x, d

"""The C oracle (oracle/pcfd_oracle.c) against fixtures produced by the reference itself.

tests/golden/*.npz were written by tools/make_golden.py from runs of the
unmodified reference (oracle/_ref/ref_harness).  The oracle restates the
reference's loops in the reference's order and is compiled without FMA
contraction like the reference, so the bar here is BIT-EXACT equality.
"""
import numpy as np
import pytest

from tests.oracle_lib import Oracle, load_golden

# box6_unsteady_bdf2: dual time stepping, third step of a BDF2 run (TemporalResidual with q^n, q^{n-1}; cnp1 V/dt + V/dtau)
INVISCID = ["box8_explicit_venkat", "box8_explicit_barth", "box8_explicit_venkatmod", "box6_implicit_sgs", "box6c_implicit_sgs",
            "ramp15_implicit", "cube_LowFi", "box6_unsteady_bdf2"]
# laminar Navier-Stokes (compressibleNS): viscous flux + analytic viscous Jacobian + no-slip wall hooks
NS = ["box6_ns_implicit", "box6_ns_adiabatic", "box6_sa_implicit"]      # also the list tests/test_gpu_viscous.py runs on the GPU
# farFieldViscous side faces (power-law scaled free stream, bc.tcc:1092-1108) next to the no-slip floor; on the GPU since
# round 2 (tests/test_gpu_viscous.py runs NS + NS_FFV); the name of the list is historical
NS_ORACLE_ONLY = ["box6_ns_ffv"]
NS_FFV = NS_ORACLE_ONLY
# oracle only so far: central-difference flux Jacobians (jacobianFieldType = jacobianBoundaryType = 1)
JAC_ORACLE_ONLY = ["box6_implicit_central",
                   # complex-step field Jacobians (jacobianFieldType = 2, jacobian.tcc:370-433; oracle/pcfd_oracle_cs.c)
                   "box6_implicit_complex"]
# oracle only so far: Green-Gauss gradients (gradientType = 1, gradient.tcc:170-248)
GG_ORACLE_ONLY = ["box8_explicit_gg"]
# general elements (hexes, prisms, pyramids, tets, quadrilateral boundary faces: boxmesh.mixed_box through the reference's
# UGRID reader); the hot path is edge-based and sees them only through the edge / half-edge lists
GENERAL = ["elem_mixed", "elem_pyramid"]
ALL = INVISCID + NS + NS_ORACLE_ONLY + JAC_ORACLE_ONLY + GG_ORACLE_ONLY + GENERAL
INVISCID_IMPLICIT = ["box6_implicit_sgs", "box6c_implicit_sgs", "ramp15_implicit", "cube_LowFi", "box6_unsteady_bdf2"]
IMPLICIT = INVISCID_IMPLICIT + NS + NS_ORACLE_ONLY + JAC_ORACLE_ONLY
EXPLICIT = ["box8_explicit_venkat", "box8_explicit_barth", "box8_explicit_venkatmod"]      # also run on the GPU (test_gpu_parity)
EXPLICIT_ORACLE = EXPLICIT + GG_ORACLE_ONLY + GENERAL


def exact(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    if bad.any():
        i = int(np.argmax(bad))
        scale = np.abs(b).max()
        raise AssertionError(f"{what}: {int(bad.sum())}/{a.size} differ; first at {i}: {a.flat[i]!r} vs "
                             f"{b.flat[i]!r}; max |diff|/max|ref| = {np.abs(a - b).max() / scale:.3e}")


@pytest.mark.parametrize("name", ALL)
def test_lsq_coefficients(oracle, name):
    g, meta = load_golden(name)
    o = Oracle(oracle, g, meta)
    s, sw = o.lsq()
    exact(s, g["lsq_s"], "s")
    exact(sw, g["lsq_sw"], "sw")


@pytest.mark.parametrize("name", ALL)
def test_update_bcs(oracle, name):
    # q_pre -> (one UpdateBCs, bc.tcc:1399-1457) -> q0 in the reference run
    g, meta = load_golden(name)
    o = Oracle(oracle, g, meta)
    q = g["q_pre"].copy()
    o.update_bcs(q, g["beta"])
    exact(q, g["q0"], "q after BC update")


@pytest.mark.parametrize("name", ALL)
def test_gradient_limiter_residual_timestep(oracle, name):
    g, meta = load_golden(name)
    o = Oracle(oracle, g, meta)
    q = g["q0"].copy()
    grad = o.gradient(q, g["lsq_sw"])
    exact(grad, g["qgrad"], "qgrad")
    lim = o.limiter(q, grad)
    exact(lim, g["limiter"], "limiter")
    b = o.residual(q, grad, lim, g["beta"])
    exact(b, g["b"], "b")
    dt, dtmin = o.timestep(q, g["beta"])
    exact(dt, g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]


@pytest.mark.parametrize("name", EXPLICIT_ORACLE)
def test_explicit_update(oracle, name):
    g, meta = load_golden(name)
    o = Oracle(oracle, g, meta)
    q = g["q0"].copy()
    x = o.explicit_solve(q, g["b"], g["timestep"])
    exact(x, g["x"], "x")
    exact(q, g["q1"], "q1")


@pytest.mark.parametrize("name", IMPLICIT)
def test_jacobian_lu_sgs(oracle, name):
    g, meta = load_golden(name)
    o = Oracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    exact(ia, g["ia"], "ia")
    exact(ja, g["ja"], "ja")
    exact(iau, g["iau"], "iau")
    q = g["q0"].copy()
    A = o.jacobian(q, g["beta"], g["timestep"], ia, ja, iau)
    exact(A, g["A"], "A")
    pv = o.prepare_sgs(iau, A)
    exact(A, g["A_lu"], "A_lu")
    exact(pv, g["pv"], "pv")
    x, ddq = o.sgs(int(meta["nSgs"]), ia, ja, iau, A, pv, g["b"])
    exact(x, g["x"], "x")
    assert ddq == g["sgs_ddq"][0]
    o.apply_dq(q, x)
    exact(q, g["q1"], "q1")


def test_spalart_allmaras_compute(oracle):
    """One TurbulenceModel::Compute (turb.tcc:163-339) of the SA model on the reference's own state: BCs, unweighted
    LSQ gradient, first-order convection + diffusion with inline Jacobians, source (exp / pow), scalar SGS, update,
    eddy viscosity.  The oracle calls the same libm, so this too is bit-exact."""
    g, meta = load_golden("box6_sa_implicit")
    o = Oracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    tvar = g["turb_tvar0"].copy()
    out = o.turb_sa(int(meta["nSgs"]), g["turb_q"], g["turb_qgrad"], g["lsq_s"], g["wallDistance"], g["turb_dt"], ia, ja,
                    iau, tvar)
    for k in ("tgrad", "b", "A", "x", "mut"):
        exact(out[k], g["turb_" + k], "turb " + k)
    exact(tvar, g["turb_tvar1"], "tvar after the update")
    assert out["res"] == g["turb_res"][0]


def test_resnorm_matches_reference_definition():
    # ParallelL2Norm (ucs/parallel.h:160-181): sqrt(sum x^2)/N
    g, meta = load_golden("box8_explicit_venkat")
    b = g["b"]
    assert np.isclose(np.sqrt(np.sum(b * b)) / b.size, g["resnorm"][0], rtol=1e-13)


def test_unsteady_fixture_exercises_the_bdf_terms(oracle):
    g, meta = load_golden("box6_unsteady_bdf2")
    assert meta["dt"] == 0.02 and int(meta["torder"]) == 2 and int(meta["iter"]) == 3
    o = Oracle(oracle, g, meta)
    b = o.residual(g["q0"].copy(), g["qgrad"], g["limiter"], g["beta"])
    exact(b, g["b"], "b")
    o.c.qold = None     # steady form: no temporal residual
    b0 = o.residual(g["q0"].copy(), g["qgrad"], g["limiter"], g["beta"])
    assert np.abs(b - b0).max() > 0.1 * np.abs(b).max()


def test_far_field_viscous_fixture_scales_the_free_stream(oracle):
    from tests.oracle_lib import C
    g, meta = load_golden("box6_ns_ffv")
    assert (g["bedges_bctype"] == 5).any() and (g["bedges_bctype"] == 4).any()
    oracle.orc_power_law_u.restype = C.c_double
    oracle.orc_power_law_u.argtypes = [C.c_double, C.c_double, C.c_double]
    d = g["wallDistance"][: int(meta["nnode"])]
    ubar = np.array([oracle.orc_power_law_u(1.0, float(x), float(meta["Re"])) for x in d])
    assert ubar.min() == 0.0 and 0.5 < ubar.max() < 1.0          # the whole box sits inside the power-law layer
    # with the plain far-field BC on the same faces the phantom states differ: the scaling is live
    o = Oracle(oracle, g, meta)
    q = g["q_pre"].copy()
    o.update_bcs(q, g["beta"])
    exact(q, g["q0"], "q after BC update")
    g2 = dict(g, bedges_bctype=np.where(g["bedges_bctype"] == 5, 6, g["bedges_bctype"]).astype(np.int32))
    o2 = Oracle(oracle, g2, meta)
    q2 = g["q_pre"].copy()
    o2.update_bcs(q2, g["beta"])
    assert np.abs(q2 - g["q0"]).max() > 1e-3


def test_central_jacobian_fixture_differs_from_the_one_sided_one(oracle):
    g, meta = load_golden("box6_implicit_central")
    assert int(meta["fieldJacType"]) == 1 and int(meta["boundaryJacType"]) == 1
    o = Oracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    A = o.jacobian(g["q0"].copy(), g["beta"], g["timestep"], ia, ja, iau)
    exact(A, g["A"], "A (central differences)")
    o.c.field_jac_type = o.c.boundary_jac_type = 0
    A0 = o.jacobian(g["q0"].copy(), g["beta"], g["timestep"], ia, ja, iau)
    offd = np.ones(len(ja), bool)
    offd[iau] = False
    A, A0 = A.reshape(-1, 25), A0.reshape(-1, 25)
    d = np.abs(A[offd] - A0[offd]).max() / np.abs(A0[offd]).max()
    assert 1e-12 < d < 1e-5          # interior edges: same derivative, different truncation / rounding error
    # boundary blocks: the reference re-evaluates the BC for the +h state only (jacobian.tcc:581-612), so the "central"
    # boundary Jacobian carries (F(BC(q+h)) - F(q-h, frozen BC)) / 2h -- reproduced as it is, far from the one-sided one
    assert np.abs(A[iau] - A0[iau]).max() > 1e-3 * np.abs(A0[iau]).max()


def test_green_gauss_fixture_differs_from_least_squares(oracle):
    g, meta = load_golden("box8_explicit_gg")
    assert int(meta["gradType"]) == 1
    o = Oracle(oracle, g, meta)
    grad = o.gradient(g["q0"].copy(), g["lsq_sw"])
    exact(grad, g["qgrad"], "qgrad (Green-Gauss)")
    o.c.grad_type = 0
    lsq = o.gradient(g["q0"].copy(), g["lsq_sw"])
    nn = int(meta["nnode"])
    d = np.abs(grad - lsq).reshape(-1, 27)[:nn].max() / np.abs(lsq).max()
    assert 1e-3 < d        # two discretisations of the same derivative (and GG is only first-order at boundary nodes)

"""The viscous reacting eqnset (compressibleNSFR) on the GPU, through the C ABI, against the fixture written by the
reference itself (tests/golden/box4_nsfr_implicit.npz: Re = 111, temperatures on both sides of the 1000 K switch between
the Sutherland law and the NASA RP-1311 fits).

Bars, and why:
 * BIT-EXACT: BC states, gradient, limiter, time step, LU and SGS given the same matrix, the species rows of the
   off-diagonal Jacobian blocks (no viscous part).
 * 1e-12 relative: the momentum / energy rows of the residual and of the off-diagonal blocks.  They carry the viscous
   flux / analytic viscous Jacobian, whose mixture viscosity and conductivity call pow / log / exp (species.tcc:393-479):
   CUDA libm here, glibc in the reference, 1-2 ulp apart.  The two state-independent powers of Wilke's rule
   (chem.tcc:908-909) are evaluated on the host with the C library, like the reference.
 * species rows of b, diagonal blocks, the implicit update: as for compressibleEulerFR (tests/test_gpu_fr.py).
"""
import numpy as np
import pytest

from tests.oracle_lib import FrOracle, load_golden
from tests.test_gpu_fr import NEQ, NS, NV, fixture_fr_params, fr_ctx, oracle_for_fr, source_scale
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu

NAME = "box4_nsfr_implicit"


def test_nsfr_update_bcs_gradient_limiter_timestep():
    from proteuscfd_b200 import capi
    ctx, g, _ = fr_ctx(NAME)
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.update_bcs()
    exact(ctx.get_field(capi.F_Q), g["q0"], "q after UpdateBCs")
    ctx.lsq_coefficients()
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")
    dtmin = ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]


def test_nsfr_residual(oracle):
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(NAME)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_QGRAD, g["qgrad"])
    ctx.set_field(capi.F_LIMITER, g["limiter"])
    s = ctx.residual(want_norms=True)
    b = ctx.get_field(capi.F_B).reshape(-1, NEQ)
    bref = g["b"].reshape(-1, NEQ)
    # momentum / energy rows: inviscid part exact, viscous part to libm rounding -> 1e-12 of the row's magnitude
    scale = np.abs(bref[:, NS:]).max(axis=0)
    err = np.abs(b[:, NS:] - bref[:, NS:]) / scale
    assert err.max() <= 1e-12, f"momentum / energy rows off by {err.max():.3e} relative"
    sc = source_scale(oracle, g, meta, g["q0"], g["vol"])
    errs = np.abs(b[:, :NS] - bref[:, :NS])
    assert np.all(errs <= 1e-12 * sc + 1e-300)
    assert np.isclose(np.sqrt(s[0]) / b.size, g["resnorm"][0], rtol=1e-12)
    # the viscous flux is really there: the oracle without it differs at the percent level
    o = FrOracle(oracle, g, meta)
    o.c.viscous = 0
    b0 = o.residual(g["q0"].copy(), g["qgrad"], g["limiter"], g["beta"]).reshape(-1, NEQ)
    assert np.abs(b0[:, NS:] - bref[:, NS:]).max() > 1e-3 * scale.max()
    assert np.abs(b[:, NS:] - bref[:, NS:]).max() < 1e-9 * np.abs(b0[:, NS:] - bref[:, NS:]).max()


def test_nsfr_jacobian_lu_sgs():
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(NAME)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    ia, ja, iau, _ = ctx.get_crs()
    exact(ia, g["ia"], "ia")
    exact(ja, g["ja"], "ja")
    A = ctx.get_field(capi.F_A).reshape(-1, NEQ, NEQ)
    Aref = g["A"].reshape(-1, NEQ, NEQ)
    offd = np.ones(len(A), bool)
    offd[iau] = False
    exact(A[offd][:, :NS, :], Aref[offd][:, :NS, :], "species rows of the off-diagonal blocks (no viscous part)")
    blk = np.abs(Aref[offd]).reshape(-1, NEQ * NEQ).max(axis=1)[:, None, None]
    err = np.abs(A[offd][:, NS:, :] - Aref[offd][:, NS:, :]) / blk
    assert err.max() <= 1e-12, f"momentum / energy rows of the off-diagonal blocks off by {err.max():.3e} of the block"
    scale = np.abs(Aref[iau]).reshape(-1, NEQ * NEQ).max(axis=1)[:, None, None]
    assert np.all(np.abs(A[iau] - Aref[iau]) <= 2e-6 * scale)
    # LU + SGS on the reference's own matrix and right-hand side: exact
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A after LU")
    exact(ctx.get_crs()[3], g["pv"], "pv")
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


def test_nsfr_implicit_iteration_close():
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(NAME)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.implicit_iterate(int(meta["nSgs"]), refresh_jac=True)
    x = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xref = g["x"].reshape(-1, NEQ)
    err = np.abs(x - xref).max(axis=0) / np.abs(xref).max(axis=0)
    assert np.all(err <= 1e-5), f"relative error of the update per equation: {err}"


def test_nsfr_frozen_implicit_vs_oracle(oracle):
    """reactionsOn = 0 leaves the transport fits as the only libm calls: the whole implicit iteration then agrees with
    the oracle to 1e-10 relative on the update (the analytic viscous Jacobian is not finite-differenced, so nothing
    amplifies the 1-2 ulp of pow / log / exp)."""
    from proteuscfd_b200 import capi
    g, meta = load_golden(NAME)
    meta = dict(meta, rxnOn=0.0)
    o = FrOracle(oracle, g, meta)
    ctx, _, _ = fr_ctx(NAME, rxn_on=0)
    q = g["q_pre"].copy()
    beta, sw = g["beta"], g["lsq_sw"]
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, beta)
    A = o.jacobian(q, beta, dt, ia, ja, iau)
    o.update_bcs(q, beta)
    grad = o.gradient(q, sw)
    lim = o.limiter(q, grad)
    b = o.residual(q, grad, lim, beta)
    pv = o.prepare_sgs(iau, A)
    x, _ = o.sgs(3, ia, ja, iau, A, pv, b)
    o.apply_dq(q, x)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.implicit_iterate(3, refresh_jac=True)
    xg = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xo = x.reshape(-1, NEQ)
    err = np.abs(xg - xo).max(axis=0) / np.abs(xo).max(axis=0)
    assert np.all(err <= 1e-10), f"relative error of the update per equation: {err}"
    qg = ctx.get_field(capi.F_Q).reshape(-1, NV)
    assert np.allclose(qg, q.reshape(-1, NV), rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("colored", [True, False])
def test_nsfr_seeded_box_vs_oracle(oracle, colored):
    """A 10^3 box in the SURVEY 8d state with a smooth, non-zero eddy-viscosity field (so the (mu + mut) and cp mut / PrT
    terms are live), frozen chemistry: two implicit iterations (the second re-using the LU'd Jacobian), GPU vs oracle.
    Gradient and limiter are bit-exact; residual and update to 1e-11 of the equation's magnitude (transport libm)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    fr, g, meta = fixture_fr_params(NAME, rxn_on=0)
    mesh, params, q0, beta = fr_box_case(10, fr, colored=colored)
    ntot = mesh["nnode"] + mesh["gnode"] + mesh["nbnode"]
    X = np.asarray(mesh["xyz"]).reshape(-1, 3)
    mut = np.zeros(ntot)
    mut[: len(X)] = 0.2 * (1.0 + 0.5 * np.sin(2 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 2]))
    mesh = dict(mesh, mut=mut)
    o = oracle_for_fr(oracle, mesh, params, g, meta)
    assert o.c.viscous == 1
    ctx = capi.Context(mesh, params)
    ctx.set_field(capi.F_BETA, beta)
    ctx.set_field(capi.F_MUT, mut)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q0)
    q = q0.copy()
    bo = beta[: mesh["nnode"] + mesh["gnode"]].copy()
    _, sw = o.lsq()
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, bo)
    A = o.jacobian(q, bo, dt, ia, ja, iau)
    pv = None

    def close(a, ref, what, tol):
        a, ref = a.reshape(-1, NEQ), ref.reshape(-1, NEQ)
        err = np.abs(a - ref).max(axis=0) / np.abs(ref).max(axis=0)
        assert np.all(err <= tol), f"{what}: relative error per equation {err}"

    for it in range(2):
        o.update_bcs(q, bo)
        grad = o.gradient(q, sw)
        lim = o.limiter(q, grad)
        b = o.residual(q, grad, lim, bo)
        if pv is None:
            pv = o.prepare_sgs(iau, A)
        x, _ = o.sgs(2, ia, ja, iau, A, pv, b)
        o.apply_dq(q, x)
        ctx.implicit_iterate(2, refresh_jac=(it == 0))
        if it == 0:
            exact(ctx.get_field(capi.F_QGRAD), grad, "qgrad")
            exact(ctx.get_field(capi.F_LIMITER), lim, "limiter")
        close(ctx.get_field(capi.F_B), b, f"b, iteration {it}", 1e-11)
        close(ctx.get_field(capi.F_X), x, f"x, iteration {it}", 1e-10)
    assert np.allclose(ctx.get_field(capi.F_Q).reshape(-1, NV), q.reshape(-1, NV), rtol=1e-10, atol=1e-14)
    # the eddy viscosity matters: the laminar oracle residual differs visibly
    o.c.mut = None
    b_lam = o.residual(q, grad, lim, bo).reshape(-1, NEQ)
    assert np.abs(b_lam[:, NS:] - b.reshape(-1, NEQ)[:, NS:]).max() > 1e-4 * np.abs(b).max()


# box4_nsfr_ffv: farFieldViscous side faces (free stream scaled by the power-law profile of the wall distance,
# bc.tcc:1092-1108, with the compounding copy of Bkernel_NumJac) next to the no-slip floor
WALL = ["box4_nsfr_wall", "box4_nsfr_adiabatic", "box4_nsfr_ffv"]


@pytest.mark.parametrize("name", WALL)
def test_nsfr_noslip_wall(oracle, name):
    """No-slip floor under the viscous reacting eqnset (isothermal 900 K / adiabatic): hard-set wall state from the
    most-normal neighbour (bc.tcc:1182-1291, compressibleFR.tcc:2048-2070) by the sequential per-node BC walk, wall rows
    of the residual (:2101-2114) and of the Jacobian (:2072-2099), against the reference's fixtures."""
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(name)
    nn = int(meta["nnode"])
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.update_bcs()
    exact(ctx.get_field(capi.F_Q), g["q0"], "q after UpdateBCs (wall nodes hard-set)")
    ctx.lsq_coefficients()
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")
    ctx.residual()
    b = ctx.get_field(capi.F_B).reshape(-1, NEQ)
    bref = g["b"].reshape(-1, NEQ)
    wall = np.zeros(nn, bool)
    wall[g["bedges_n"].reshape(-1, 2)[: int(meta["nbedge"]), 0][g["bedges_bctype"][: int(meta["nbedge"])] == 4]] = True
    assert wall.any() and np.all(bref[wall, NS:] == 0.0)
    assert np.all(b[wall, NS:] == 0.0), "momentum / energy rows of wall nodes must be hard zeros"
    scale = np.abs(bref[:, NS:]).max(axis=0)
    assert (np.abs(b[:, NS:] - bref[:, NS:]) / scale).max() <= 1e-12
    sc = source_scale(oracle, g, meta, g["q0"], g["vol"])
    assert np.all(np.abs(b[:, :NS] - bref[:, :NS]) <= 1e-12 * sc + 1e-300)
    dtmin = ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]
    # Jacobian: boundary pass by node walk, wall rows blanked with a unit diagonal (and -1 towards the normal node)
    ctx.jacobian()
    ia, ja, iau, _ = ctx.get_crs()
    A = ctx.get_field(capi.F_A).reshape(-1, NEQ, NEQ)
    Aref = g["A"].reshape(-1, NEQ, NEQ)
    rows = np.repeat(np.arange(nn), np.diff(ia))
    wblk = wall[rows]
    exact(A[wblk][:, NS:, :], Aref[wblk][:, NS:, :], "wall rows of every block of a wall node")
    offd = np.ones(len(A), bool)
    offd[iau] = False
    exact(A[offd][:, :NS, :], Aref[offd][:, :NS, :], "species rows of the off-diagonal blocks")
    blk = np.abs(Aref[offd]).reshape(-1, NEQ * NEQ).max(axis=1)[:, None, None]
    assert (np.abs(A[offd][:, NS:, :] - Aref[offd][:, NS:, :]) / blk).max() <= 1e-12
    scale = np.abs(Aref[iau]).reshape(-1, NEQ * NEQ).max(axis=1)[:, None, None]
    assert np.all(np.abs(A[iau] - Aref[iau]) <= 2e-6 * scale)
    nloc = nn + int(meta["gnode"])
    qa = ctx.get_field(capi.F_Q).reshape(-1, NV)
    exact(qa[:nloc], g["q0"].reshape(-1, NV)[:nloc], "interior rows after the boundary Jacobian pass")
    exact(qa[nloc:], g["q1"].reshape(-1, NV)[nloc:], "phantom rows after the boundary Jacobian pass")
    # LU + SGS on the reference's own matrix and right-hand side: exact
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A after LU")
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


@pytest.mark.parametrize("name", WALL)
def test_nsfr_noslip_implicit_iteration_close(oracle, name):
    """The whole implicit iteration through pcfd_implicit_iterate against the ORACLE stepped in the same order
    (NewtonIterate takes the time step and the Jacobian before UpdateBCs; the harness that wrote the fixture after it, and
    with hard-set walls that moves dt by a few percent -- so the fixture's x is not the reference point here)."""
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(name)
    o = FrOracle(oracle, g, meta)
    q = g["q_pre"].copy()
    beta, sw = g["beta"], g["lsq_sw"]
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, beta)
    A = o.jacobian(q, beta, dt, ia, ja, iau)
    o.update_bcs(q, beta)
    grad = o.gradient(q, sw)
    lim = o.limiter(q, grad)
    b = o.residual(q, grad, lim, beta)
    pv = o.prepare_sgs(iau, A)
    xo, _ = o.sgs(int(meta["nSgs"]), ia, ja, iau, A, pv, b)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.implicit_iterate(int(meta["nSgs"]), refresh_jac=True)
    exact(ctx.get_field(capi.F_TIMESTEP), dt, "timestep")
    exact(ctx.get_field(capi.F_LIMITER), lim, "limiter")
    x = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xo = xo.reshape(-1, NEQ)
    err = np.abs(x - xo).max(axis=0) / np.abs(xo).max(axis=0)
    assert np.all(err <= 1e-5), f"relative error of the update per equation: {err}"


def _sa_fr_ctx():
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx("box4_nsfr_sa")
    assert int(meta["turbModel"]) == 1
    ctx.set_field(capi.F_Q, g["turb_q"])
    ctx.set_field(capi.F_QGRAD, g["turb_qgrad"])
    ctx.set_field(capi.F_TIMESTEP, g["turb_dt"])
    ctx.set_field(capi.F_LSQ_S, g["lsq_s"])
    ctx.set_field(capi.F_TVAR, g["turb_tvar0"])
    return ctx, g, meta


def test_spalart_allmaras_under_nsfr_vs_reference():
    """turbulenceModel = 1 with compressibleNSFR: pcfd_turb_compute against the reference's own TurbulenceModel::Compute
    run under the reacting eqnset (tests/golden/box4_nsfr_sa.npz: Wilke-mixed molecular viscosity, density from the aux
    variables, native velocities, Re = Param::Re, velocity gradient at row nspecies).  The LSQ gradient of nu~ is
    bit-exact; everything downstream of the species viscosity fits and the source term's exp / pow carries 1e-12."""
    from proteuscfd_b200 import capi
    from tests.test_gpu_viscous import close_per_node
    ctx, g, meta = _sa_fr_ctx()
    ss = ctx.turb_compute(int(meta["nSgs"]), want_norm=True)
    exact(ctx.get_field(capi.F_TGRAD), g["turb_tgrad"], "tgrad")
    close_per_node(ctx.get_field(capi.F_TURB_B), g["turb_b"], 1, "turbulence residual b")
    close_per_node(ctx.get_field(capi.F_TURB_A), g["turb_A"], 1, "turbulence matrix (inverted diagonal)")
    close_per_node(ctx.get_field(capi.F_TURB_X), g["turb_x"], 1, "turbulence update x")
    close_per_node(ctx.get_field(capi.F_TVAR), g["turb_tvar1"], 1, "nu~ after the update")
    nn = g["turb_mut"].size
    close_per_node(ctx.get_field(capi.F_MUT)[:nn], g["turb_mut"], 1, "eddy viscosity")
    assert np.isclose(np.sqrt(ss) / ctx.nnode, g["turb_res"][0], rtol=1e-12)
    assert np.abs(g["turb_x"]).max() > 0.1


def test_spalart_allmaras_under_nsfr_phases_equal_the_monolithic_call():
    """pcfd_turb_phase 0..5 (the cut at the reference's exchange points) is the same kernels in the same order"""
    from proteuscfd_b200 import capi
    ctx, g, meta = _sa_fr_ctx()
    ctx.turb_compute(int(meta["nSgs"]))
    ref = {f: ctx.get_field(f).copy() for f in (capi.F_TVAR, capi.F_TGRAD, capi.F_TURB_B, capi.F_TURB_A, capi.F_TURB_X, capi.F_MUT)}
    ctx2, _, _ = _sa_fr_ctx()
    for ph in (0, 1, 2):
        ctx2.turb_phase(ph)
    for _ in range(int(meta["nSgs"])):
        ctx2.turb_phase(3)
    ctx2.turb_phase(4)
    ctx2.turb_phase(5)
    for f, v in ref.items():
        exact(ctx2.get_field(f), v, f"field {f}")


def test_spalart_allmaras_needs_a_viscous_reacting_eqnset():
    from proteuscfd_b200 import capi
    fr, g, meta = fixture_fr_params("box4_fr_implicit")
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    params = dict(sorder=2, limiter=0, no_cvbc=0, gamma=0.0, chi=0.0, cfl=1.0, turb_model=1, fr=fr)
    with pytest.raises(capi.PcfdError, match="Spalart-Allmaras needs compressibleNSFR"):
        capi.Context(mesh, params)


def test_nsfr_sa_implicit_iteration_updates_the_model(oracle):
    """pcfd_implicit_iterate on a compressibleNSFR + SA context runs TurbulenceModel::Compute after the flow update
    (solutionSpace.tcc:862-866): nu~ and mu_t afterwards equal a separate pcfd_turb_compute on the same state, and the
    oracle's model update from the GPU's own flow state agrees to 1e-12."""
    from proteuscfd_b200 import capi
    from tests.test_gpu_viscous import close_per_node
    ctx, g, meta = fr_ctx("box4_nsfr_sa")
    nsgs = int(meta["nSgs"])
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.set_field(capi.F_TVAR, g["turb_tvar0"])
    mut0 = np.zeros(ctx.field_size(capi.F_MUT)); mut0[: g["mut"].size] = g["mut"]
    ctx.set_field(capi.F_MUT, mut0)
    ctx.implicit_iterate(nsgs, refresh_jac=True)
    q, qgrad, dt = ctx.get_field(capi.F_Q), ctx.get_field(capi.F_QGRAD), ctx.get_field(capi.F_TIMESTEP)
    o = FrOracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    tvar = g["turb_tvar0"].copy()
    out = o.turb_sa(nsgs, q[: g["turb_q"].size], qgrad, ctx.get_field(capi.F_LSQ_S), g["wallDistance"], dt, ia, ja, iau, tvar)
    close_per_node(ctx.get_field(capi.F_TVAR)[: tvar.size], tvar, 1, "nu~ after the iteration")
    close_per_node(ctx.get_field(capi.F_MUT)[: out["mut"].size], out["mut"], 1, "eddy viscosity after the iteration")

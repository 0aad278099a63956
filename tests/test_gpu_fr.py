"""The reacting eqnset (compressibleEulerFR: HLLC flux, NASA-7 thermodynamics, characteristic BCs, finite-rate source
term, 9x9 block Jacobian + SGS) on the GPU, through the C ABI, against the fixtures written by the reference itself
(tests/golden/box5_fr_explicit.npz, box4_fr_implicit.npz) and against the C oracle on a larger seeded box.

Bars, and why:
 * BIT-EXACT: BC states, gradient, limiter, time step, the momentum and energy rows of the residual, the flux part of
   the Jacobian, LU and SGS given the same matrix -- only +, -, *, /, sqrt, in the reference's order (--fmad=false).
 * 1e-12 of the rate scale: the species rows of the residual.  They carry the chemistry source term, which calls
   exp / log / pow: CUDA libm there, glibc in the reference (1-2 ulp apart), and the net production rate cancels.
 * FD-amplified: the reference forms the source-term Jacobian by one-sided finite differences with h = 1e-8
   (eqnset.tcc:163-187), so a 1-ulp difference of the source becomes ~1e-8 of |S|/h-sized entries in the diagonal
   blocks.  Diagonal blocks are compared to 2e-6 of the block's largest entry, the implicit update to 1e-5 relative
   (measured: 1.3e-6 -- the level at which the reference's own FD Jacobian changes with the libm it is linked to);
   with reactionsOn = 0 the whole implicit iteration is bit-exact (test_fr_frozen_implicit_bit_exact).
"""
import numpy as np
import pytest

from tests.oracle_lib import FrOracle, chem_tables, load_golden
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu

NS, NEQ, NV, NT = 5, 9, 21, 14


def fr_ctx(name, rxn_on=None):
    from proteuscfd_b200 import capi
    g, meta = load_golden(name)
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    if (g["bedges_bctype"] == 4).any():      # no-slip walls: wall temperature per half-edge (< 0: adiabatic)
        mesh["bedges_twall"] = g["bedges_twall"]
    fr = dict(chem=chem_tables(g), ref_density=meta["ref_density"], ref_velocity=meta["ref_velocity"],
              ref_temperature=meta["ref_temperature"], ref_pressure=meta["ref_pressure"], ref_time=meta["ref_time"],
              ref_specific_enthalpy=meta["ref_specific_enthalpy"], pref=meta["Pref"], dt=meta["dt"],
              use_local_dt=int(meta["useLocalTimeStepping"]), rxn_on=int(meta["rxnOn"]) if rxn_on is None else rxn_on,
              qinf=g["qinf"])
    params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]), gamma=0.0,
                  chi=meta["chi"], cfl=meta["cfl"], fr=fr)
    if int(meta.get("viscous", 0)):      # compressibleNSFR: species transport tables, Re, PrT
        fr.update(transport={k: g[k] for k in ("species_mu_fit", "species_k_fit", "species_white", "species_fit_counts")},
                  ref_viscosity=meta["ref_viscosity"], ref_k=meta["ref_k"])
        params.update(Re=meta["Re"], PrT=meta["PrT"], turb_model=int(meta.get("turbModel", 0)))
    ctx = capi.Context(mesh, params)
    assert (ctx.neqn, ctx.nvars, ctx.nterms) == (NEQ, NV, NT)
    beta = np.ones(ctx.field_size(capi.F_BETA))
    beta[: g["beta"].size] = g["beta"]
    ctx.set_field(capi.F_BETA, beta)
    if "wallDistance" in g and ctx.field_size(capi.F_WALLDIST) > 0:      # viscous far-field BC: power-law profile
        wd = np.zeros(ctx.field_size(capi.F_WALLDIST))
        wd[: min(wd.size, g["wallDistance"].size)] = g["wallDistance"][: wd.size]
        ctx.set_field(capi.F_WALLDIST, wd)
    return ctx, g, meta


def species_rows_close(b, bref, src_scale, what):
    """momentum / energy rows exact; species rows within 1e-12 of the production-rate scale of the node."""
    b, bref = b.reshape(-1, NEQ), bref.reshape(-1, NEQ)
    exact(b[:, NS:], bref[:, NS:], what + " (momentum, energy rows)")
    err = np.abs(b[:, :NS] - bref[:, :NS])
    assert np.all(err <= 1e-12 * src_scale + 1e-300), f"{what}: species rows off by {np.max(err / (src_scale + 1e-300)):.3e} of scale"


def source_scale(oracle, g, meta, q, vol):
    """vol * (rate magnitude the net production cancels from), non-dimensional, per node and species."""
    from tests.oracle_lib import ChemOracle
    t = chem_tables(g)
    o = ChemOracle(oracle, t)
    n = vol.size
    Q = q.reshape(-1, NV)[:n]
    _, sc, _, _ = o.mass_production(Q[:, :NS] * meta["ref_density"], Q[:, NS + 3] * meta["ref_temperature"])
    return vol[:, None] * sc / (meta["ref_density"] / meta["ref_time"])


@pytest.mark.parametrize("name", ["box5_fr_explicit", "box4_fr_implicit"])
def test_fr_update_bcs(name):
    from proteuscfd_b200 import capi
    ctx, g, _ = fr_ctx(name)
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.update_bcs()
    exact(ctx.get_field(capi.F_Q), g["q0"], "q after UpdateBCs")


@pytest.mark.parametrize("name", ["box5_fr_explicit", "box4_fr_implicit"])
def test_fr_gradient_limiter_residual_timestep(oracle, name):
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx(name)
    ctx.lsq_coefficients()
    exact(ctx.get_field(capi.F_LSQ_SW), g["lsq_sw"], "sw")
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")
    s = ctx.residual(want_norms=True)
    b = ctx.get_field(capi.F_B)
    species_rows_close(b, g["b"], source_scale(oracle, g, meta, g["q0"], g["vol"]), "b")
    assert np.isclose(np.sqrt(s[0]) / b.size, g["resnorm"][0], rtol=1e-12)
    dtmin = ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]


def test_fr_explicit_update():
    """ExplicitSolve's native -> conservative -> Newton-on-T -> native update (solve.tcc:112-130) and ApplyDQ, fed the
    reference's own residual so that the comparison is exact."""
    from proteuscfd_b200 import capi
    ctx, g, _ = fr_ctx("box5_fr_explicit")
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.explicit_solve()
    exact(ctx.get_field(capi.F_X)[: g["x"].size], g["x"], "x")
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


def test_fr_jacobian_lu_sgs():
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx("box4_fr_implicit")
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    ia, ja, iau, _ = ctx.get_crs()
    exact(ia, g["ia"], "ia")
    exact(ja, g["ja"], "ja")
    A = ctx.get_field(capi.F_A).reshape(-1, NEQ * NEQ)
    Aref = g["A"].reshape(-1, NEQ * NEQ)
    offd = np.ones(len(A), bool)
    offd[iau] = False
    exact(A[offd], Aref[offd], "off-diagonal blocks (flux Jacobian)")
    # diagonal blocks carry the FD source-term Jacobian: libm rounding / h (see the module docstring)
    scale = np.abs(Aref[iau]).max(axis=1, keepdims=True)
    assert np.all(np.abs(A[iau] - Aref[iau]) <= 2e-6 * scale)
    # Bkernel_NumJac re-runs the (iterated) BC map on the phantom states: the reference's q1 holds them
    nloc = int(meta["nnode"]) + int(meta["gnode"])
    qa = ctx.get_field(capi.F_Q).reshape(-1, NV)
    exact(qa[:nloc], g["q0"].reshape(-1, NV)[:nloc], "interior rows after the boundary Jacobian pass")
    exact(qa[nloc:], g["q1"].reshape(-1, NV)[nloc:], "phantom rows after the boundary Jacobian pass")
    # LU + SGS on the reference's own matrix and right-hand side: exact
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A after LU")
    exact(ctx.get_crs()[3], g["pv"], "pv")
    ctx.blank_x()
    ddq = ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")
    assert np.isclose(ddq, g["sgs_ddq"][0], rtol=1e-10)
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


def test_fr_reactions_on_update_with_reference_matrix():
    """Pins the split claimed in the module docstring: with reactions ON, the only thing that keeps the implicit update
    away from the north-star 1e-12 is the reference's finite-difference source Jacobian (libm rounding / h) in the
    diagonal blocks.  Here the fixture's A is injected and everything else is ours (BCs, gradient, limiter, HLLC residual
    + finite-rate source from q_pre, LU, SGS): x must then agree to 1e-12 of each equation's largest update."""
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx("box4_fr_implicit")
    assert int(meta["rxnOn"]) == 1
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.update_bcs()
    ctx.gradient()
    ctx.limiter()
    ctx.residual()
    ctx.set_field(capi.F_A, g["A"])
    ctx.prepare_sgs()
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]), want_ddq=False)
    x = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xref = g["x"].reshape(-1, NEQ)
    err = np.abs(x - xref).max(axis=0) / np.abs(xref).max(axis=0)
    assert np.all(err <= 1e-12), f"relative error of the update per equation with the reference's A: {err}"


def test_fr_implicit_iteration_close():
    """the whole implicit iteration from q_pre, own residual and own Jacobian: 1e-5 relative on the update."""
    from proteuscfd_b200 import capi
    ctx, g, meta = fr_ctx("box4_fr_implicit")
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.implicit_iterate(int(meta["nSgs"]), refresh_jac=True)
    x = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xref = g["x"].reshape(-1, NEQ)
    err = np.abs(x - xref).max(axis=0) / np.abs(xref).max(axis=0)
    assert np.all(err <= 1e-5), f"relative error of the update per equation: {err}"   # measured on B200: <= 1.3e-6


def test_fr_frozen_implicit_bit_exact(oracle):
    """reactionsOn = 0 removes the only libm calls: the implicit iteration (HLLC residual, 19-flux FD Jacobian, boundary
    Jacobian, dense temporal terms, LU, SGS, ApplyDQ) is then bit-identical to the oracle, here on a 10^3 box in the
    reference's state."""
    from proteuscfd_b200 import capi
    g, meta = load_golden("box4_fr_implicit")
    meta = dict(meta, rxnOn=0.0)
    o = FrOracle(oracle, g, meta)
    ctx, _, _ = fr_ctx("box4_fr_implicit", rxn_on=0)
    q = g["q_pre"].copy()
    beta = g["beta"]
    sw = g["lsq_sw"]
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
    exact(ctx.get_field(capi.F_B), b, "b")
    exact(ctx.get_field(capi.F_A), A, "A (LU form)")
    exact(ctx.get_field(capi.F_X), x, "x")
    exact(ctx.get_field(capi.F_Q), q, "q")


def fixture_fr_params(name="box4_fr_implicit", rxn_on=None):
    g, meta = load_golden(name)
    extra = {}
    if int(meta.get("viscous", 0)):      # compressibleNSFR: species transport tables, Re, PrT
        extra = dict(transport={k: g[k] for k in ("species_mu_fit", "species_k_fit", "species_white", "species_fit_counts")},
                     ref_viscosity=meta["ref_viscosity"], ref_k=meta["ref_k"], Re=meta["Re"], PrT=meta["PrT"])
    return dict(chem=chem_tables(g), **extra, ref_density=meta["ref_density"], ref_velocity=meta["ref_velocity"],
                ref_temperature=meta["ref_temperature"], ref_pressure=meta["ref_pressure"], ref_time=meta["ref_time"],
                ref_specific_enthalpy=meta["ref_specific_enthalpy"], pref=meta["Pref"], dt=meta["dt"],
                use_local_dt=int(meta["useLocalTimeStepping"]), rxn_on=int(meta["rxnOn"]) if rxn_on is None else rxn_on,
                qinf=g["qinf"]), g, meta


def oracle_for_fr(lib, mesh, params, g, meta):
    """FrOracle on a generated mesh: fixture tables / reference values, the mesh and numerics of `mesh`, `params`."""
    gg = dict(g)
    gg.pop("mut", None)      # the fixture's eddy-viscosity field belongs to the fixture's mesh
    if mesh.get("mut") is not None:
        gg["mut"] = np.asarray(mesh["mut"], dtype=np.float64)
    for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp"):
        gg[k] = np.asarray(mesh[k])
    mm = dict(meta)
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mm[k] = mesh[k]
    fr = params["fr"]
    mm.update(limiter=params["limiter"], sorder=params["sorder"], no_cvbc=params["no_cvbc"], chi=params["chi"],
              cfl=params["cfl"], rxnOn=fr["rxn_on"], gamma=1.4)
    return FrOracle(lib, gg, mm)


@pytest.mark.parametrize("colored,limiter", [(True, 2), (False, 2), (True, 3), (False, 1)])
def test_fr_seeded_box_vs_oracle(oracle, colored, limiter):
    """A 10^3 box (1331 nodes, 8.6 k edges) in the SURVEY 8d state, frozen chemistry: two implicit iterations (the
    second re-using the LU'd Jacobian) then an explicit one, every field bit-identical to the oracle.  colored = the
    colour-sorted numbering used at scale (8 SGS levels per direction), else lexicographic (level schedule)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    fr, g, meta = fixture_fr_params(rxn_on=0)
    mesh, params, q0, beta = fr_box_case(10, fr, colored=colored, limiter=limiter)   # 2 Venkatakrishnan, 3 modified, 1 Barth
    o = oracle_for_fr(oracle, mesh, params, g, meta)
    ctx = capi.Context(mesh, params)
    ctx.set_field(capi.F_BETA, beta)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q0)
    q = q0.copy()
    bo = beta[: mesh["nnode"] + mesh["gnode"]].copy()
    _, sw = o.lsq()
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, bo)
    A = o.jacobian(q, bo, dt, ia, ja, iau)
    pv = None
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
        exact(ctx.get_field(capi.F_QGRAD), grad, f"qgrad, iteration {it}")
        exact(ctx.get_field(capi.F_LIMITER), lim, f"limiter, iteration {it}")
        exact(ctx.get_field(capi.F_B), b, f"b, iteration {it}")
        exact(ctx.get_field(capi.F_X), x, f"x, iteration {it}")
        exact(ctx.get_field(capi.F_Q), q, f"q, iteration {it}")
    o.c.cfl = 0.05
    ctx.set_cfl(0.05)
    dt, _ = o.timestep(q, bo)       # ComputeTimesteps runs before UpdateBCs (PreTimeAdvance, solutionSpace.tcc:629-634)
    o.update_bcs(q, bo)
    grad = o.gradient(q, sw)
    lim = o.limiter(q, grad)
    b = o.residual(q, grad, lim, bo)
    x = o.explicit_solve(q, b, dt)
    o.apply_dq(q, x)
    ctx.explicit_iterate(refresh_dt=True)
    exact(ctx.get_field(capi.F_TIMESTEP), dt, "timestep")
    exact(ctx.get_field(capi.F_X), x, "explicit x")
    exact(ctx.get_field(capi.F_Q), q, "q after the explicit iteration")


def test_fr_reacting_box_close_to_oracle(oracle):
    """the same box with the reactions on: residual species rows to 1e-12 of the rate scale, update to 1e-5."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    fr, g, meta = fixture_fr_params()
    mesh, params, q0, beta = fr_box_case(10, fr)
    o = oracle_for_fr(oracle, mesh, params, g, meta)
    ctx = capi.Context(mesh, params)
    ctx.set_field(capi.F_BETA, beta)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q0)
    q = q0.copy()
    bo = beta[: mesh["nnode"] + mesh["gnode"]].copy()
    _, sw = o.lsq()
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, bo)
    A = o.jacobian(q, bo, dt, ia, ja, iau)
    o.update_bcs(q, bo)
    grad = o.gradient(q, sw)
    lim = o.limiter(q, grad)
    b = o.residual(q, grad, lim, bo)
    pv = o.prepare_sgs(iau, A)
    x, _ = o.sgs(3, ia, ja, iau, A, pv, b)
    ctx.implicit_iterate(3, refresh_jac=True)
    scale = source_scale(oracle, g, meta, q, np.asarray(mesh["vol"]))
    species_rows_close(ctx.get_field(capi.F_B), b, scale, "b")
    xg = ctx.get_field(capi.F_X).reshape(-1, NEQ)
    xo = x.reshape(-1, NEQ)
    err = np.abs(xg - xo).max(axis=0) / np.abs(xo).max(axis=0)
    assert np.all(err <= 1e-5), f"relative error of the update per equation: {err}"

"""The CUDA hot path (through the C ABI) against (1) the golden fixtures written by the
reference itself and (2) the C oracle on larger seeded cases.

Bar: BIT-EXACT (integer and floating point alike).  The kernels gather each node's
contributions in the reference's edge order and are compiled with --fmad=false, so
there is nothing to tolerate; the only tolerance in this file is on the SGS
convergence monitor |xOld - xNorm| (a parallel sum of squares), stated where used.
"""
import numpy as np
import pytest

from tests.oracle_lib import Oracle, load_golden, oracle_for
from tests.test_oracle import EXPLICIT, INVISCID as ALL, INVISCID_IMPLICIT as IMPLICIT, exact

pytestmark = pytest.mark.gpu


def golden_ctx(name):
    from proteuscfd_b200 import capi
    g, meta = load_golden(name)
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]),
                  gamma=meta["gamma"], chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
    if int(meta.get("viscous", 0)):
        mesh["bedges_twall"] = g["bedges_twall"]
        params.update(eqnset=capi.EQNSET_COMPRESSIBLE_NS, Re=meta["Re"], Pr=meta["Pr"], PrT=meta["PrT"],
                      tref=meta["ref_temperature"], mach=meta["velocity"], enable_vnn=int(meta["enableVNN"]),
                      vnn=meta["VNN"], turb_model=int(meta.get("turbModel", 0)))
    ctx = capi.Context(mesh, params)
    if "wallDistance" in g and ctx.field_size(capi.F_WALLDIST) > 0:   # SA model / viscous far-field BC (power-law profile)
        wd = np.zeros(ctx.field_size(capi.F_WALLDIST))
        wd[: min(wd.size, g["wallDistance"].size)] = g["wallDistance"][: wd.size]
        ctx.set_field(capi.F_WALLDIST, wd)
    if "mut" in g:   # eddy viscosity the flow's viscous terms saw in the reference run
        mut = np.zeros(ctx.field_size(capi.F_MUT))
        mut[: g["mut"].size] = g["mut"]
        ctx.set_field(capi.F_MUT, mut)
    if float(meta.get("dt", -1.0)) > 0.0 and "qold" in g:   # unsteady fixture: BDF terms live (tests/test_gpu_unsteady.py)
        ctx.set_time_integration(meta["dt"], int(meta["useLocalTimeStepping"]), int(meta["torder"]), int(meta["iter"]))
        ctx.set_field(capi.F_QOLD, g["qold"])
        ctx.set_field(capi.F_QOLDM1, g["qoldm1"])
    return ctx, g, meta


@pytest.mark.parametrize("name", ALL)
def test_lsq_coefficients(name):
    from proteuscfd_b200 import capi
    ctx, g, _ = golden_ctx(name)
    ctx.lsq_coefficients()
    exact(ctx.get_field(capi.F_LSQ_S), g["lsq_s"], "s")
    exact(ctx.get_field(capi.F_LSQ_SW), g["lsq_sw"], "sw")


@pytest.mark.parametrize("name", ALL)
def test_update_bcs(name):
    from proteuscfd_b200 import capi
    ctx, g, _ = golden_ctx(name)
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.update_bcs()
    exact(ctx.get_field(capi.F_Q), g["q0"], "q after UpdateBCs")


@pytest.mark.parametrize("name", ALL)
def test_gradient_limiter_residual_timestep(name):
    from proteuscfd_b200 import capi
    ctx, g, _ = golden_ctx(name)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")
    s = ctx.residual(want_norms=True)
    b = ctx.get_field(capi.F_B)
    exact(b, g["b"], "b")
    # ParallelL2Norm (parallel.h:160-181): sqrt(sum)/N -- parallel sum, so 1e-13 relative
    assert np.isclose(np.sqrt(s[0]) / b.size, g["resnorm"][0], rtol=1e-13)
    dtmin = ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]


@pytest.mark.parametrize("name", EXPLICIT)
def test_explicit_update(name):
    from proteuscfd_b200 import capi
    ctx, g, _ = golden_ctx(name)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.explicit_solve()
    exact(ctx.get_field(capi.F_X)[: g["x"].size], g["x"], "x")
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


@pytest.mark.parametrize("name", EXPLICIT)
def test_explicit_iterate_composite(name):
    """pcfd_explicit_iterate == the phase-by-phase sequence (NewtonIterate with nSgs == 0)."""
    from proteuscfd_b200 import capi
    ctx, g, _ = golden_ctx(name)
    ref, _, _ = golden_ctx(name)
    for c in (ctx, ref):
        c.lsq_coefficients()
        c.set_field(capi.F_Q, g["q_pre"])
    ctx.explicit_iterate(refresh_dt=True)
    ref.timestep(want_min=False); ref.update_bcs(); ref.gradient(); ref.limiter(); ref.residual(); ref.explicit_solve()
    exact(ctx.get_field(capi.F_B), g["b"], "b")      # b does not depend on dt
    exact(ctx.get_field(capi.F_Q), ref.get_field(capi.F_Q), "q1")


@pytest.mark.parametrize("name", IMPLICIT)
def test_jacobian_lu_sgs(name):
    from proteuscfd_b200 import capi
    ctx, g, meta = golden_ctx(name)
    ia, ja, iau, _ = ctx.get_crs()
    exact(ia, g["ia"], "ia")
    exact(ja, g["ja"], "ja")
    exact(iau, g["iau"], "iau")
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), g["A"], "A")
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A_lu")
    exact(ctx.get_crs()[3], g["pv"], "pv")
    ctx.set_field(capi.F_B, g["b"])
    ctx.blank_x()
    ddq = ctx.sgs(int(meta["nSgs"]))
    x = ctx.get_field(capi.F_X)
    exact(x, g["x"], "x")
    # |xOld - xNorm| is a difference of two parallel sums: 1e-12 of xNorm
    xnorm = np.sqrt(np.sum(x * x)) / (ctx.nnode * 5)
    assert abs(ddq - g["sgs_ddq"][0]) <= 1e-12 * xnorm
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


# ---------------------------------------------------------------- larger seeded cases vs the C oracle
def run_both_explicit(oracle, mesh, params, q, steps=2):
    from proteuscfd_b200 import capi
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ctx.lsq_coefficients()
    s, sw = o.lsq()
    exact(ctx.get_field(capi.F_LSQ_SW), sw, "sw")
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    beta = np.zeros(1)
    for it in range(steps):
        dt, dtmin = o.timestep(qo, beta)
        o.update_bcs(qo, beta)
        grad = o.gradient(qo, sw)
        lim = o.limiter(qo, grad)
        b = o.residual(qo, grad, lim, beta)
        o.explicit_solve(qo, b, dt)
        ctx.explicit_iterate(refresh_dt=True)
        exact(ctx.get_field(capi.F_QGRAD), grad, f"qgrad it{it}")
        exact(ctx.get_field(capi.F_LIMITER), lim, f"limiter it{it}")
        exact(ctx.get_field(capi.F_B), b, f"b it{it}")
        exact(ctx.get_field(capi.F_Q), qo, f"q it{it}")
    return lim


@pytest.mark.parametrize("limiter", [1, 2])
def test_box24_explicit_vs_oracle(oracle, limiter):
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(24, limiter=limiter)
    run_both_explicit(oracle, mesh, params, q)


def test_ramp_explicit_vs_oracle(oracle):
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(16, ramp_deg=15.0, mach=2.0, jitter=0.1)
    run_both_explicit(oracle, mesh, params, q)


def test_dirichlet_type_bcs_vs_oracle(oracle):
    """Dirichlet / sonic-inflow half-edges overwrite the LEFT node state (bc.tcc:1058-1120), so the nodes that own
    one are walked sequentially; mixed with far-field half-edges at the same nodes (box edges and corners)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    bc = {1: capi.BC_SONIC_INFLOW, 2: capi.BC_SONIC_OUTFLOW, 3: capi.BC_SYMMETRY, 4: capi.BC_DIRICHLET,
          5: capi.BC_FARFIELD, 6: capi.BC_NEUMANN}
    mesh, params, q = box_case(10, bc=bc, cfl=5.0)
    run_both_explicit(oracle, mesh, params, q)
    # and through the boundary Jacobian
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ia, ja, iau = o.crs_init()
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    dt, _ = o.timestep(qo, np.zeros(1))
    A = o.jacobian(qo, np.zeros(1), dt, ia, ja, iau)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), A, "A")
    exact(ctx.get_field(capi.F_Q), qo, "q after the boundary Jacobian")


def test_pressure_clip_sequential_semantics(oracle):
    """A rough state makes Kernel_PressureClip (limiters.tcc:737-815) fire on many edges, including
    chains where a later edge sees an earlier clip; the fixed-point iteration must reproduce the
    reference's sequential result exactly."""
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(12, limiter=2)
    rng = np.random.default_rng(7)
    nn = mesh["nnode"]
    Q = q.reshape(-1, 10)
    # random strong pressure/density jumps
    Q[:nn, 0] *= rng.choice([0.05, 1.0, 4.0], size=nn)
    Q[:nn, 4] *= rng.choice([0.3, 1.0, 6.0], size=nn)
    from proteuscfd_b200.cases import aux_vars
    aux_vars(Q, params["gamma"])
    lim = run_both_explicit(oracle, mesh, params, Q.reshape(-1), steps=1)
    nz = int((lim.reshape(-1, 5)[:nn].max(axis=1) == 0.0).sum())
    assert nz > 10, "test state did not trigger the pressure clip"


@pytest.mark.parametrize("colored", [False, True])
def test_box16_implicit_vs_oracle(oracle, colored):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(16, cfl=5.0, colored=colored)
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ctx.lsq_coefficients()
    _, sw = o.lsq()
    ia, ja, iau = o.crs_init()
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    beta = np.zeros(1)
    nsgs = 4
    for it in range(2):
        dt, _ = o.timestep(qo, beta)
        A = o.jacobian(qo, beta, dt, ia, ja, iau)
        o.update_bcs(qo, beta)
        grad = o.gradient(qo, sw)
        lim = o.limiter(qo, grad)
        b = o.residual(qo, grad, lim, beta)
        A0 = A.copy()
        pv = o.prepare_sgs(iau, A)
        x, _ = o.sgs(nsgs, ia, ja, iau, A, pv, b)
        o.apply_dq(qo, x)

        ctx.timestep(want_min=False)
        ctx.jacobian()
        exact(ctx.get_field(capi.F_A), A0, f"A it{it}")
        ctx.update_bcs()
        ctx.gradient()
        ctx.limiter()
        ctx.residual()
        exact(ctx.get_field(capi.F_B), b, f"b it{it}")
        ctx.prepare_sgs()
        exact(ctx.get_field(capi.F_A), A, f"A_lu it{it}")
        ctx.blank_x()
        ctx.sgs(nsgs)
        exact(ctx.get_field(capi.F_X), x, f"x it{it}")
        ctx.apply_dq()
        exact(ctx.get_field(capi.F_Q), qo, f"q it{it}")


def test_implicit_iterate_composite(oracle):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(10, cfl=5.0, colored=True)
    a = capi.Context(mesh, params)
    b = capi.Context(mesh, params)
    for c in (a, b):
        c.lsq_coefficients()
        c.set_field(capi.F_Q, q)
    a.implicit_iterate(3, refresh_jac=True)
    b.timestep(want_min=False); b.jacobian(); b.update_bcs(); b.gradient(); b.limiter(); b.residual()
    b.prepare_sgs(); b.blank_x(); b.sgs(3); b.apply_dq()
    exact(a.get_field(capi.F_Q), b.get_field(capi.F_Q), "q")


def test_run_to_run_determinism():
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(20)
    outs = []
    for _ in range(3):
        ctx = capi.Context(mesh, params)
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, q)
        for _ in range(3):
            ctx.explicit_iterate()
        outs.append(ctx.get_field(capi.F_Q))
        ctx.close()
    exact(outs[0], outs[1], "run 0 vs 1")
    exact(outs[0], outs[2], "run 0 vs 2")


def test_create_rejects_bad_input():
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(4)
    bad = dict(mesh)
    bad["edges_n"] = mesh["edges_n"].copy()
    bad["edges_n"][0] = 10 ** 6
    with pytest.raises(capi.PcfdError):
        capi.Context(bad, params)
    p2 = dict(params, eqnset=99)
    with pytest.raises(capi.PcfdError):
        capi.Context(mesh, p2)
    ctx = capi.Context(mesh, params)
    with pytest.raises(capi.PcfdError):
        ctx.set_field(capi.F_Q, np.zeros(3))

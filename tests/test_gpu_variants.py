"""GPU parity tests of the variants of ABI v6: Green-Gauss gradients (Param::gradType = 1, gradient.tcc:170-248) for both
eqnset families and central-difference flux Jacobians (fieldJacType = boundaryJacType = 1, jacobian.tcc:306-366, 546-640)
for both families, their end-to-end drop-in runs through include/pcfd_host.hpp, the error behaviour of the two setters,
and pcfd_turb_phase (Spalart-Allmaras cut at the reference's exchange points) against the monolithic pcfd_turb_compute.

First B200 run: the driver's round-1 GPU test pass (GPUTEST_r01.json, 13 xpassed); the xfail marker the file carried
until then is gone, so a regression now fails the suite.

Bar: BIT-EXACT against the fixtures the reference wrote, like tests/test_gpu_parity.py (ordered gathers, --fmad=false).
"""
import numpy as np
import pytest

from tests.oracle_lib import oracle_for
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu


def test_green_gauss_gradient_fixture():
    """qgrad, then limiter / residual / time step computed from it, against the reference's box8_explicit_gg dump."""
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box8_explicit_gg")
    assert int(meta["gradType"]) == 1
    ctx.set_gradient_type(1)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad (Green-Gauss)")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")
    ctx.residual()
    exact(ctx.get_field(capi.F_B), g["b"], "b")
    ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    ctx.explicit_solve()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


def test_green_gauss_explicit_iterate_vs_oracle(oracle):
    """two explicit iterations on a seeded 16^3 box with symmetry planes (the symmetry fix follows the volume division)"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    bc = {1: capi.BC_FARFIELD, 2: capi.BC_FARFIELD, 3: capi.BC_SYMMETRY, 4: capi.BC_SYMMETRY, 5: capi.BC_FARFIELD,
          6: capi.BC_IMPERMEABLE_WALL}
    mesh, params, q = box_case(16, bc=bc, limiter=2)
    o = oracle_for(oracle, mesh, params)
    o.c.grad_type = 1
    ctx = capi.Context(mesh, params)
    ctx.set_gradient_type(1)
    ctx.lsq_coefficients()
    _, sw = o.lsq()
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    beta = np.zeros(1)
    for it in range(2):
        dt, _ = o.timestep(qo, beta)
        o.update_bcs(qo, beta)
        grad = o.gradient(qo, sw)
        lim = o.limiter(qo, grad)
        b = o.residual(qo, grad, lim, beta)
        o.explicit_solve(qo, b, dt)
        ctx.explicit_iterate(refresh_dt=True)
        exact(ctx.get_field(capi.F_QGRAD), grad, f"qgrad it{it}")
        exact(ctx.get_field(capi.F_B), b, f"b it{it}")
        exact(ctx.get_field(capi.F_Q), qo, f"q it{it}")


def test_green_gauss_gradient_fr_fixture():
    """the reacting eqnset (14 gradient terms) against the reference's box4_fr_gg dump"""
    from proteuscfd_b200 import capi
    from tests.test_gpu_fr import fr_ctx
    ctx, g, meta = fr_ctx("box4_fr_gg")
    assert int(meta["gradType"]) == 1
    ctx.set_gradient_type(1)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.gradient()
    exact(ctx.get_field(capi.F_QGRAD), g["qgrad"], "qgrad (Green-Gauss, reacting)")
    ctx.limiter()
    exact(ctx.get_field(capi.F_LIMITER), g["limiter"], "limiter")


def test_central_difference_jacobian_fixture():
    """A, its LU, the SGS update and q1 against the reference's box6_implicit_central dump"""
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box6_implicit_central")
    assert int(meta["fieldJacType"]) == 1 and int(meta["boundaryJacType"]) == 1
    ctx.set_jacobian_type(1, 1)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), g["A"], "A (central differences)")
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A_lu")
    ctx.set_field(capi.F_B, g["b"])
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


@pytest.mark.parametrize("types", [(1, 1), (1, 0), (0, 1)])
def test_central_difference_jacobian_vs_oracle(oracle, types):
    """field and boundary types independently, Dirichlet-type half-edges (node walk) mixed with far field / symmetry"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    bc = {1: capi.BC_SONIC_INFLOW, 2: capi.BC_SONIC_OUTFLOW, 3: capi.BC_SYMMETRY, 4: capi.BC_DIRICHLET,
          5: capi.BC_FARFIELD, 6: capi.BC_NEUMANN}
    mesh, params, q = box_case(10, bc=bc, cfl=5.0)
    o = oracle_for(oracle, mesh, params)
    o.c.field_jac_type, o.c.boundary_jac_type = types
    ctx = capi.Context(mesh, params)
    ctx.set_jacobian_type(*types)
    ia, ja, iau = o.crs_init()
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    dt, _ = o.timestep(qo, np.zeros(1))
    A = o.jacobian(qo, np.zeros(1), dt, ia, ja, iau)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), A, "A")
    exact(ctx.get_field(capi.F_Q), qo, "q after the boundary Jacobian")


def test_central_difference_jacobian_fr_fixture():
    """the reacting eqnset (9x9 blocks, HLLC) against the reference's box4_fr_central dump: off-diagonal blocks (pure
    flux Jacobian) bit-exact; diagonal blocks carry the one-sided FD source-term Jacobian (libm rounding / h, see
    tests/test_gpu_fr.py): 2e-6 of the block's largest entry, the bar of test_fr_jacobian_lu_sgs"""
    from proteuscfd_b200 import capi
    from tests.test_gpu_fr import NEQ, NV, fr_ctx
    ctx, g, meta = fr_ctx("box4_fr_central")
    assert int(meta["fieldJacType"]) == 1 and int(meta["boundaryJacType"]) == 1
    ctx.set_jacobian_type(1, 1)
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    _, _, iau, _ = ctx.get_crs()
    A = ctx.get_field(capi.F_A).reshape(-1, NEQ * NEQ)
    Aref = g["A"].reshape(-1, NEQ * NEQ)
    offd = np.ones(len(A), bool)
    offd[iau] = False
    exact(A[offd], Aref[offd], "off-diagonal blocks (central flux Jacobian)")
    scale = np.abs(Aref[iau]).max(axis=1, keepdims=True)
    assert np.all(np.abs(A[iau] - Aref[iau]) <= 2e-6 * scale)
    nloc = int(meta["nnode"]) + int(meta["gnode"])
    qa = ctx.get_field(capi.F_Q).reshape(-1, NV)
    exact(qa[nloc:], g["q1"].reshape(-1, NV)[nloc:], "phantom rows after the boundary Jacobian pass")


def test_central_difference_jacobian_fr_frozen_vs_oracle(oracle):
    """reactionsOn = 0 removes the libm calls: the whole central-difference Jacobian is then bit-identical to the oracle"""
    from proteuscfd_b200 import capi
    from tests.oracle_lib import FrOracle, load_golden
    from tests.test_gpu_fr import fr_ctx
    g, meta = load_golden("box4_fr_central")
    meta = dict(meta, rxnOn=0.0)
    o = FrOracle(oracle, g, meta)
    ctx, _, _ = fr_ctx("box4_fr_central", rxn_on=0)
    ctx.set_jacobian_type(1, 1)
    q = g["q_pre"].copy()
    ia, ja, iau = o.crs_init()
    dt, _ = o.timestep(q, g["beta"])
    A = o.jacobian(q, g["beta"], dt, ia, ja, iau)
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.timestep(want_min=False)
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), A, "A")
    exact(ctx.get_field(capi.F_Q)[: q.size], q, "q after the boundary Jacobian")


@pytest.mark.parametrize("kind", ["explicit_green_gauss", "implicit_central"])
def test_reference_with_dropin_matches_reference_variants(kind):
    """end to end through include/pcfd_host.hpp (which hands Param::gradType / fieldJacType / boundaryJacType over): the
    unmodified reference with gradientType = 1 resp. jacobianFieldType = jacobianBoundaryType = 1 against the same
    harness with the GPU phases, every dumped array bit-identical (tests/test_dropin_reference.py for the defaults)"""
    from oracle import ref_bench
    if not ref_bench.available(dropin=True):
        pytest.skip("oracle/_ref binaries not built (needs /root/reference at build time)")
    if kind == "explicit_green_gauss":
        case = ref_bench.ReferenceCase(10, 1, limiter=2, nsgs=0, cfl=0.5, gradtype=1)
    else:
        case = ref_bench.ReferenceCase(8, 1, limiter=2, nsgs=3, cfl=5.0, colored=True, jactype=1)
    try:
        cpu = case.dump(dropin=False)[0]
        gpu = case.dump(dropin=True)[0]
    finally:
        case.close()
    assert set(cpu) == set(gpu)
    for name in sorted(cpu):
        if name == "resnorm":
            assert np.allclose(gpu[name], cpu[name], rtol=1e-13, atol=0), name
        elif name == "sgs_ddq":
            xn = np.sqrt(np.sum(cpu["x"] ** 2)) / max(cpu["b"].size, 1)
            assert abs(gpu[name][0] - cpu[name][0]) <= 1e-12 * xn
        else:
            exact(gpu[name], cpu[name], name)


def test_variant_setters_reject_what_is_not_built():
    """the complex-step BOUNDARY Jacobian (boundary type 2, jacobian.tcc:170-172), unknown types and unknown gradient types
    are refused with an error, never silently mapped to another kernel (the complex-step FIELD Jacobian, field type 2, is
    built for the perfect-gas eqnsets: tests/test_zzz_gpu_complex_step.py)"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(4)
    ctx = capi.Context(mesh, params)
    with pytest.raises(capi.PcfdError, match="complex"):
        ctx.set_jacobian_type(0, 2)
    with pytest.raises(capi.PcfdError, match="complex"):
        ctx.set_jacobian_type(3, 0)
    with pytest.raises(capi.PcfdError, match="Green-Gauss"):
        ctx.set_gradient_type(2)
    ctx.set_jacobian_type(1, 0)
    ctx.set_gradient_type(1)


def test_turb_phases_equal_turb_compute():
    """pcfd_turb_phase 0..5 (nSgs repeats of phase 3, no exchange) relaunch the kernels of pcfd_turb_compute in the same
    order: on one partition every turbulence field and the residual sum must come out bit-identical"""
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    res = []
    for phased in (False, True):
        ctx, g, meta = golden_ctx("box6_sa_implicit")
        ctx.set_field(capi.F_Q, g["turb_q"])             # the set-up of test_spalart_allmaras_compute_vs_reference
        ctx.set_field(capi.F_QGRAD, g["turb_qgrad"])
        ctx.set_field(capi.F_TIMESTEP, g["turb_dt"])
        ctx.set_field(capi.F_LSQ_S, g["lsq_s"])
        ctx.set_field(capi.F_WALLDIST, g["wallDistance"])
        ctx.set_field(capi.F_TVAR, g["turb_tvar0"])
        nsgs = 3
        if phased:
            ctx.turb_phase(0)
            ctx.turb_phase(1)
            s = ctx.turb_phase(2, want_norm=True)
            for _ in range(nsgs):
                ctx.turb_phase(3)
            ctx.turb_phase(4)
            ctx.turb_phase(5)
        else:
            s = ctx.turb_compute(nsgs, want_norm=True)
        res.append((s, ctx.get_field(capi.F_TVAR), ctx.get_field(capi.F_MUT), ctx.get_field(capi.F_TURB_X),
                    ctx.get_field(capi.F_TGRAD)))
    assert res[0][0] == res[1][0]
    for a, b, what in zip(res[0][1:], res[1][1:], ("tvar", "mut", "turb_x", "tgrad")):
        exact(b, a, what)
    assert np.abs(res[0][3]).max() > 0


@pytest.mark.parametrize("family", ["perfect_gas", "reacting"])
def test_apply_dq_zeroes_the_update_of_a_node_with_a_non_finite_component(family):
    """NewtonIterate (solutionSpace.tcc:771-796): a node whose update has a NaN / Inf component gets its WHOLE update
    zeroed (in crs->x too) before ApplyDQ; the other nodes are updated as usual."""
    from proteuscfd_b200 import capi
    if family == "perfect_gas":
        from tests.test_gpu_parity import golden_ctx
        ctx, g, meta = golden_ctx("box6_implicit_sgs")
    else:
        from tests.test_gpu_fr import fr_ctx
        ctx, g, meta = fr_ctx("box4_fr_implicit")
    neqn, nvars = ctx.neqn, ctx.nvars
    x = g["x"].copy()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_X, x)
    ctx.apply_dq()
    qref = ctx.get_field(capi.F_Q).reshape(-1, nvars)
    assert ctx.zeroed_updates() == 0
    bad = x.reshape(-1, neqn).copy()
    bad[3, 1] = np.nan
    bad[7, neqn - 1] = np.inf
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_X, bad.reshape(-1))
    ctx.apply_dq()
    q = ctx.get_field(capi.F_Q).reshape(-1, nvars)
    q0 = g["q0"].reshape(-1, nvars)
    xs = ctx.get_field(capi.F_X).reshape(-1, neqn)
    assert ctx.zeroed_updates() == 2
    for n in (3, 7):
        assert np.all(xs[n] == 0.0)
        assert np.array_equal(q[n, :neqn], q0[n, :neqn])
    keep = np.ones(ctx.nnode, dtype=bool)
    keep[[3, 7]] = False
    assert np.array_equal(q[: ctx.nnode][keep], qref[: ctx.nnode][keep])
    assert np.array_equal(xs[: ctx.nnode][keep], x.reshape(-1, neqn)[: ctx.nnode][keep])


@pytest.mark.parametrize("name", ["box6_implicit_sgs", "box6_implicit_central", "box6_ns_implicit", "box6_sa_implicit",
                                  "box9_3rank_implicit_r1of3", "cube_LowFi"])
def test_jacobian_overwrites_every_block_perfect_gas(name):
    """pcfd_jacobian does not blank the matrix first (CRSMatrix::Blank would stream 5-16 GB for nothing): every block is
    written in full before anything is added to it.  Poison the matrix with NaN, refresh, compare with the reference's A."""
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx(name)
    ctx.set_jacobian_type(int(meta.get("fieldJacType", 0)), int(meta.get("boundaryJacType", 0)))
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.set_field(capi.F_A, np.full(g["A"].size, np.nan))
    ctx.jacobian()
    A = ctx.get_field(capi.F_A)
    assert np.isfinite(A).all(), f"{int((~np.isfinite(A)).sum())} entries of A were not overwritten"
    exact(A, g["A"], "A after a refresh on a poisoned matrix")


@pytest.mark.parametrize("name", ["box4_fr_implicit", "box4_nsfr_wall", "box4_fr_central"])
def test_jacobian_overwrites_every_block_reacting(name):
    from proteuscfd_b200 import capi
    from tests.test_gpu_fr import fr_ctx
    ctx, g, meta = fr_ctx(name)
    ctx.set_jacobian_type(int(meta.get("fieldJacType", 0)), int(meta.get("boundaryJacType", 0)))
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    ref = ctx.get_field(capi.F_A).copy()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_field(capi.F_A, np.full(ref.size, np.nan))
    ctx.jacobian()
    A = ctx.get_field(capi.F_A)
    assert np.isfinite(A).all(), f"{int((~np.isfinite(A)).sum())} entries of A were not overwritten"
    exact(A, ref, "A after a refresh on a poisoned matrix")

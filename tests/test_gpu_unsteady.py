"""Unsteady (dual time stepping) path on the GPU: TemporalResidual (residual.tcc:125-179) with q^n, q^{n-1} and the
BDF2 coefficients of the third step, and the cnp1 V/dt + V/dtau diagonal terms (eqnset.tcc:195-208; reacting:
compressibleFR.tcc:1326-1331), through the C ABI (pcfd_set_time_integration, fields PCFD_F_QOLD / PCFD_F_QOLDM1) against
fixtures written by the reference itself (tests/golden/box6_unsteady_bdf2.npz, box4_fr_unsteady.npz).

Perfect gas: bit-exact everywhere.  Reacting: as tests/test_gpu_fr.py (species rows of b to 1e-12 of the rate scale,
diagonal blocks to 2e-6 of the block, everything else bit-exact)."""
import numpy as np
import pytest

from tests.oracle_lib import load_golden
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu


def pg_ctx(name):
    from proteuscfd_b200 import capi
    g, meta = load_golden(name)
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]), gamma=meta["gamma"],
                  chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
    return capi.Context(mesh, params), g, meta


def arm(ctx, g, meta):
    from proteuscfd_b200 import capi
    ctx.set_time_integration(meta["dt"], int(meta["useLocalTimeStepping"]), int(meta["torder"]), int(meta["iter"]))
    ctx.set_field(capi.F_QOLD, g["qold"])
    ctx.set_field(capi.F_QOLDM1, g["qoldm1"])


def test_unsteady_perfect_gas_bit_exact():
    from proteuscfd_b200 import capi
    ctx, g, meta = pg_ctx("box6_unsteady_bdf2")
    assert ctx.field_size(capi.F_QOLD) == g["qold"].size
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.gradient()
    ctx.limiter()
    # steady form first: without q^n the temporal residual is absent
    ctx.residual()
    b_steady = ctx.get_field(capi.F_B)
    arm(ctx, g, meta)
    s = ctx.residual(want_norms=True)
    b = ctx.get_field(capi.F_B)
    exact(b, g["b"], "b with the BDF2 terms")
    assert np.abs(b - b_steady).max() > 0.1 * np.abs(b).max()
    assert np.isclose(np.sqrt(s[0]) / b.size, g["resnorm"][0], rtol=1e-12)
    dtmin = ctx.timestep()
    exact(ctx.get_field(capi.F_TIMESTEP), g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), g["A"], "A (cnp1 V/dt + V/dtau on the diagonal)")
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A after LU")
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")
    ctx.apply_dq()
    exact(ctx.get_field(capi.F_Q), g["q1"], "q1")


def test_unsteady_composite_iteration_bit_exact():
    """pcfd_implicit_iterate (fused limiter / residual pair) from q_pre == the phase-by-phase sequence in the same order
    (NewtonIterate takes the time step and the Jacobian before UpdateBCs, the harness that wrote the fixture after it, so
    x is compared with a second context, b -- which does not depend on the time step -- with the fixture as well)."""
    from proteuscfd_b200 import capi
    ctx, g, meta = pg_ctx("box6_unsteady_bdf2")
    ref, _, _ = pg_ctx("box6_unsteady_bdf2")
    nsgs = int(meta["nSgs"])
    for c in (ctx, ref):
        c.lsq_coefficients()
        c.set_field(capi.F_Q, g["q_pre"])
        arm(c, g, meta)
    ctx.implicit_iterate(nsgs, refresh_jac=True)
    ref.timestep(want_min=False); ref.jacobian(); ref.update_bcs(); ref.gradient(); ref.limiter(); ref.residual()
    ref.prepare_sgs(); ref.blank_x(); ref.sgs(nsgs, want_ddq=False); ref.apply_dq()
    exact(ctx.get_field(capi.F_B), g["b"], "b")
    exact(ctx.get_field(capi.F_B), ref.get_field(capi.F_B), "b (phase by phase)")
    exact(ctx.get_field(capi.F_X), ref.get_field(capi.F_X), "x")
    exact(ctx.get_field(capi.F_Q), ref.get_field(capi.F_Q), "q1")
    x, xref = ctx.get_field(capi.F_X), g["x"]
    assert np.abs(x - xref).max() <= 1e-10 * np.abs(xref).max()


def test_unsteady_reacting(oracle):
    from proteuscfd_b200 import capi
    from tests.test_gpu_fr import NEQ, fr_ctx, source_scale, species_rows_close
    ctx, g, meta = fr_ctx("box4_fr_unsteady")
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q0"])
    arm(ctx, g, meta)
    ctx.gradient()
    ctx.limiter()
    ctx.residual()
    b = ctx.get_field(capi.F_B)
    # the temporal terms are exact arithmetic on top of the spatial residual: same bars as the steady fixture
    species_rows_close(b, g["b"], source_scale(oracle, g, meta, g["q0"], g["vol"]), "b")
    ctx.set_field(capi.F_TIMESTEP, g["timestep"])
    ctx.jacobian()
    ia, ja, iau, _ = ctx.get_crs()
    A = ctx.get_field(capi.F_A).reshape(-1, NEQ * NEQ)
    Aref = g["A"].reshape(-1, NEQ * NEQ)
    offd = np.ones(len(A), bool)
    offd[iau] = False
    exact(A[offd], Aref[offd], "off-diagonal blocks")
    scale = np.abs(Aref[iau]).max(axis=1, keepdims=True)
    assert np.all(np.abs(A[iau] - Aref[iau]) <= 2e-6 * scale)
    # LU + SGS on the reference's own matrix and right-hand side: exact
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.prepare_sgs()
    exact(ctx.get_field(capi.F_A), g["A_lu"], "A after LU")
    ctx.blank_x()
    ctx.sgs(int(meta["nSgs"]))
    exact(ctx.get_field(capi.F_X), g["x"], "x")


@pytest.mark.parametrize("family", ["pg", "fr"])
def test_global_time_step_branches(family):
    """ComputeTimesteps without local time stepping (timestep.tcc:47-74): Param::dt everywhere when it is positive; the
    rank's smallest CFL-limited step everywhere otherwise.  min() is exact, so the second branch is compared with the
    minimum of the local-time-stepping field, and the diagonal of A with the field actually used."""
    from proteuscfd_b200 import capi
    if family == "pg":
        ctx, g, meta = pg_ctx("box6_unsteady_bdf2")
    else:
        from tests.test_gpu_fr import fr_ctx
        ctx, g, meta = fr_ctx("box4_fr_unsteady")
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q0"])
    ctx.set_time_integration(-1.0, 1, 1, 1)
    dtmin_local = ctx.timestep()
    local = ctx.get_field(capi.F_TIMESTEP)
    assert dtmin_local == local.min() and local.max() > local.min()
    ctx.set_time_integration(-1.0, 0, 1, 1)          # steady, global step: min over the rank's cells
    assert ctx.timestep() == dtmin_local
    exact(ctx.get_field(capi.F_TIMESTEP), np.full_like(local, dtmin_local), "timestep (global minimum)")
    ctx.set_time_integration(0.0125, 0, 1, 1)        # unsteady, prescribed global step
    assert ctx.timestep() == 0.0125
    exact(ctx.get_field(capi.F_TIMESTEP), np.full_like(local, 0.0125), "timestep (Param::dt)")
    if family == "pg":
        # explicit update advances every cell with that step (solve.tcc:71-98)
        ctx.update_bcs(); ctx.gradient(); ctx.limiter(); ctx.residual()
        b = ctx.get_field(capi.F_B).reshape(-1, 5)
        ctx.explicit_solve()
        x = ctx.get_field(capi.F_X).reshape(-1, 5)[: b.shape[0]]
        exact(x, b * 0.0125 / g["vol"][:, None], "explicit dq with the global step")

"""Parity at BENCHMARK SCALE (BASELINE configs[1] size: Kuhn box n = 118, 1 685 159 nodes, 11 626 894 edges): the paths
that only engage on big meshes -- many-tile k_sgs_tile_t levels with their L2 prefetch distance, the RED_BLOCKS
reductions, 5 GB (5x5) / 16 GB (9x9) matrices with nblocks*81 close to 2^31 -- against the C oracle on the SAME mesh and
state, bit for bit.  The oracle (single core) needs a few seconds per phase at this size; its 56 s finite-difference
Jacobian is run at n = 64 only, the n = 118 LU / SGS comparison takes the GPU's own matrix as input on both sides.

Bar: `==` on every double (perfect gas, frozen reacting chemistry); parallel sums (residual norm) 1e-13."""
import numpy as np
import pytest

from tests.oracle_lib import oracle_for
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu

N_BENCH = 118


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def test_explicit_iteration_at_bench_scale_vs_oracle(oracle):
    """bench.py's timed step on bench.py's mesh (lexicographic numbering): two explicit iterations through
    pcfd_explicit_iterate against the oracle phase by phase -- time step (and its global minimum through the
    fixed-tree reduction), gradient, limiter, residual (and its sum of squares), update."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(N_BENCH, device="cuda")
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ctx.lsq_coefficients()
    _, sw = o.lsq()
    exact(ctx.get_field(capi.F_LSQ_SW), sw, "sw")
    ctx.set_field(capi.F_Q, q)
    qo = q.copy()
    beta = np.zeros(1)
    for it in range(2):
        dt, dtmin = o.timestep(qo, beta)
        assert ctx.timestep() == dtmin
        exact(ctx.get_field(capi.F_TIMESTEP), dt, f"timestep it{it}")
        o.update_bcs(qo, beta)
        grad = o.gradient(qo, sw)
        lim = o.limiter(qo, grad)
        b = o.residual(qo, grad, lim, beta)
        o.explicit_solve(qo, b, dt)
        s = ctx.explicit_iterate(refresh_dt=True, want_norms=True)
        exact(ctx.get_field(capi.F_QGRAD), grad, f"qgrad it{it}")
        exact(ctx.get_field(capi.F_LIMITER), lim, f"limiter it{it}")
        exact(ctx.get_field(capi.F_B), b, f"b it{it}")
        exact(ctx.get_field(capi.F_Q), qo, f"q it{it}")
        assert np.isclose(s[0], float(np.dot(b, b)), rtol=1e-13)
    assert ctx.clip_fallbacks() == 0


def test_implicit_iteration_n64_vs_oracle(oracle):
    """the whole implicit iteration incl. the 11-flux finite-difference Jacobian on a 274 625-node box (colour-sorted
    numbering: the bulk-copy SGS tiles, several hundred tiles per level)"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(64, cfl=5.0, colored=True, device="cuda")
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    _, sw = o.lsq()
    ia, ja, iau = o.crs_init()
    qo = q.copy()
    beta = np.zeros(1)
    dt, _ = o.timestep(qo, beta)
    A = o.jacobian(qo, beta, dt, ia, ja, iau)
    o.update_bcs(qo, beta)
    grad = o.gradient(qo, sw)
    lim = o.limiter(qo, grad)
    b = o.residual(qo, grad, lim, beta)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    exact(ctx.get_field(capi.F_A), A, "A")
    pv = o.prepare_sgs(iau, A)
    x, _ = o.sgs(3, ia, ja, iau, A, pv, b)
    o.apply_dq(qo, x)
    ctx.implicit_iterate(3, refresh_jac=False)
    exact(ctx.get_field(capi.F_B), b, "b")
    exact(ctx.get_field(capi.F_A), A, "A after LU")
    exact(ctx.get_crs()[3], pv, "pv")
    exact(ctx.get_field(capi.F_X), x, "x after 3 sweeps")
    exact(ctx.get_field(capi.F_Q), qo, "q after ApplyDQ")


def test_sgs_at_bench_scale_vs_oracle(oracle):
    """LU of the diagonal blocks and two SGS sweeps on the n = 118 matrix (24.9 M 5x5 blocks, 5 GB; 16 levels of
    ~4400 bulk-copy tiles each): the oracle factors and sweeps the matrix the GPU assembled."""
    if _host_ram_gb() < 24:
        pytest.skip("needs 24 GB of host memory for the matrix copies")
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(N_BENCH, cfl=5.0, colored=True, device="cuda")
    o = oracle_for(oracle, mesh, params)
    ctx = capi.Context(mesh, params)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    ctx.update_bcs()
    ctx.gradient()
    ctx.limiter()
    ctx.residual()
    A = ctx.get_field(capi.F_A)
    b = ctx.get_field(capi.F_B)
    assert np.isfinite(A).all() and np.abs(b).max() > 0
    ia, ja, iau = o.crs_init()
    gia, gja, giau, _ = ctx.get_crs()
    exact(gia, ia, "ia")
    exact(gja, ja, "ja")
    exact(giau, iau, "iau")
    pv = o.prepare_sgs(iau, A)          # factors A in place
    ctx.prepare_sgs()
    exact(ctx.get_crs()[3], pv, "pv")
    exact(ctx.get_field(capi.F_A), A, "A after LU")
    x, ddq = o.sgs(2, ia, ja, iau, A, pv, b)
    ctx.blank_x()
    g_ddq = ctx.sgs(2)
    exact(ctx.get_field(capi.F_X), x, "x after 2 sweeps")
    assert np.isclose(g_ddq, ddq, rtol=1e-10)


def test_reacting_sgs_at_bench_scale_vs_oracle(oracle):
    """9x9 blocks at n = 118: 24.9 M blocks = 2.02e9 doubles (16.2 GB), within 6 % of 2^31 entries -- index arithmetic
    of the Jacobian scatter, the LU and the tiles at that size.  Frozen chemistry (no libm on the path): LU and two
    sweeps bit-exact against the oracle working on the GPU's own matrix."""
    if _host_ram_gb() < 60:
        pytest.skip("needs 60 GB of host memory for the matrix copies")
    import torch
    if torch.cuda.mem_get_info()[0] < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of device memory")
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_box_case
    from tests.test_gpu_fr import fixture_fr_params, oracle_for_fr
    fr, g, meta = fixture_fr_params("box4_fr_implicit", rxn_on=0)
    mesh, params, q, beta = fr_box_case(N_BENCH, fr, device="cuda")
    o = oracle_for_fr(oracle, mesh, params, g, meta)
    ctx = capi.Context(mesh, params)
    ctx.set_field(capi.F_BETA, beta)
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, q)
    ctx.timestep(want_min=False)
    ctx.jacobian()
    ctx.update_bcs()
    ctx.gradient()
    ctx.limiter()
    ctx.residual()
    A = ctx.get_field(capi.F_A)
    assert A.size == ctx.get_crs()[1].size * 81 and A.size > 2.0e9
    assert np.isfinite(A).all()
    b = ctx.get_field(capi.F_B)
    ia, ja, iau = o.crs_init()
    pv = o.prepare_sgs(iau, A)
    ctx.prepare_sgs()
    exact(ctx.get_crs()[3], pv, "pv")
    exact(ctx.get_field(capi.F_A), A, "A after LU")
    x, _ = o.sgs(2, ia, ja, iau, A, pv, b)
    ctx.blank_x()
    ctx.sgs(2, want_ddq=False)
    exact(ctx.get_field(capi.F_X), x, "x after 2 sweeps")

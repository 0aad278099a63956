"""B200 run of the local ILU0 preconditioner of CRS::GMRES (precondType 3) through the C ABI (pcfd_gmres), against the
REFERENCE's own solutions (tests/golden/box6_gmres_ilu0, box4_fr_gmres_ilu0, box8_2rank_gmres_ilu0_r*of2) and the C oracle.
tests/test_ilu0.py runs the kernels' source text on the host bit-exactly against the oracle; this file is their B200 run
(profiles/r2_gpu_tests_ilu0_call145.log).  Bars as in
tests/test_gmres.py: the factorisation and the two sweeps keep the reference's arithmetic per entry, the dot products of
GMRES are fixed-tree sums, so 1e-12 of the solution's scale on one rank and 1e-10 across ranks."""
import numpy as np
import pytest

from tests.oracle_lib import load_golden, load_oracle
from tests.test_gmres import gpu_ctx, run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,neqn,bar", [("box6_gmres_ilu0", 5, 1e-12), ("box4_fr_gmres_ilu0", 9, 1e-7)])
def test_gpu_ilu0_gmres_vs_reference(name, neqn, bar):
    """First B200 run (gpurun_out/s6_ilu0.log): 5x5 inside 1e-12; 9x9 at 4.2e-9 -- the summation order of the dot products
    alone moves the reference's own 9x9 solution by 3.9e-9 (tests/test_ilu0.py::test_what_the_summation_order_...: the
    Hessenberg diagonal of the ILU0-preconditioned reacting system spans six decades), hence 1e-7 there; with ONE search
    direction there is nothing to amplify and the 9x9 path is held to 1e-12 below."""
    from proteuscfd_b200 import capi
    ctx, g, meta = gpu_ctx(name, neqn)
    pt, nd, nr = [int(v) for v in g["gmres_cfg"]]
    assert pt == 3
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.blank_x()
    dq = ctx.gmres(nr, nd, pt)
    x = ctx.get_field(capi.F_X)
    ref = g["gmres_x"]
    assert np.abs(x - ref).max() <= bar * np.abs(ref).max(), np.abs(x - ref).max() / np.abs(ref).max()
    assert np.isclose(dq, g["gmres_dq"][0], rtol=1e-9 if neqn == 5 else 1e-5)
    assert np.array_equal(ctx.get_field(capi.F_A), g["A"])        # the factorisation works on a copy
    # one search direction: x = N^-1 v0 g0 / h00, no ill-conditioned back substitution between the sums and x
    ref1, dq1 = run_oracle(load_oracle(), g, meta, neqn, cfg=(3, 1, 1))
    ctx.blank_x()
    dq = ctx.gmres(1, 1, 3)
    x = ctx.get_field(capi.F_X)
    assert np.abs(x - ref1).max() <= 1e-12 * np.abs(ref1).max(), np.abs(x - ref1).max() / np.abs(ref1).max()
    assert np.isclose(dq, dq1, rtol=1e-9)


def test_gpu_ilu0_gmres_variants_vs_oracle():
    """other restart / direction counts and a non-zero initial guess against the oracle; the context's matrix is left
    as assembled, so the SGS path still factors and solves it afterwards"""
    from proteuscfd_b200 import capi
    ctx, g, meta = gpu_ctx("box6_gmres_ilu0", 5)
    lib = load_oracle()
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    for cfg in ((3, 3, 1), (3, 10, 2)):
        ref, dq_ref = run_oracle(lib, g, meta, 5, cfg=cfg)
        ctx.blank_x()
        dq = ctx.gmres(cfg[2], cfg[1], cfg[0])
        x = ctx.get_field(capi.F_X)
        assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max(), (cfg, np.abs(x - ref).max() / np.abs(ref).max())
        assert np.isclose(dq, dq_ref, rtol=1e-8)
    x0 = 0.5 * g["gmres_x"]
    ref, _ = run_oracle(lib, g, meta, 5, cfg=(3, 4, 1), x0=x0)
    ctx.set_field(capi.F_X, x0)
    ctx.gmres(1, 4, 3)
    assert np.abs(ctx.get_field(capi.F_X) - ref).max() <= 1e-11 * np.abs(ref).max()
    ctx.prepare_sgs()
    ctx.blank_x()
    ctx.sgs(3)
    assert np.array_equal(ctx.get_field(capi.F_X), g["x"])


def test_gpu_ilu0_gmres_two_ranks_vs_reference():
    """two thread ranks against TWO reference processes: the factorisation skips the ghost columns and blanks their blocks,
    the preconditioned vector is exchanged before every product (crs.tcc:300), dot products in rank order"""
    from proteuscfd_b200 import capi
    from tests.test_gpu_comm import run_threads
    parts = []
    for r in (0, 1):
        g, meta = load_golden(f"box8_2rank_gmres_ilu0_r{r}of2")
        mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp",
                                  "gNodeOwner", "gNodeLocalId")}
        for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
            mesh[k] = int(meta[k])
        params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]),
                      gamma=meta["gamma"], chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
        parts.append((mesh, params, g))
    pt, nd, nr = [int(v) for v in parts[0][2]["gmres_cfg"]]
    assert pt == 3

    def prepare(ctx):      # Krylov scratch and the preconditioner's copy of the matrix: allocated before the ranks connect
        ctx.gmres(1, nd, 0)
        ctx.gmres(1, nd, 3)

    def body(rank, ctx, x):
        g = parts[rank][2]
        ctx.set_field(capi.F_A, g["A"])
        ctx.set_field(capi.F_B, g["b"])
        ctx.blank_x()
        dq = ctx.gmres(nr, nd, pt)
        return ctx.get_field(capi.F_X), dq

    got = run_threads(parts, body, prepare=prepare)
    for r in (0, 1):
        ref = parts[r][2]["gmres_x"]
        x, dq = got[r]
        assert np.abs(x - ref).max() <= 1e-10 * np.abs(ref).max(), (r, np.abs(x - ref).max() / np.abs(ref).max())
        assert np.isclose(dq, parts[r][2]["gmres_dq"][0], rtol=1e-8)
    assert got[0][1] == got[1][1]

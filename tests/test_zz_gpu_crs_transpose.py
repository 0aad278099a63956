"""B200 run of CRSMatrix::CRSTranspose through the C ABI (pcfd_crs_transpose, pcfd_crs_ghost_blocks) against the REFERENCE's
own transposed Jacobians (tests/golden/*_transpose*.npz): one rank for both block sizes, two and four thread ranks with the
ghost-column blocks routed by proteuscfd_b200.parallel.crs_transpose.  Pure data movement: bit for bit.
tests/test_crs_transpose.py runs the same kernels' source text on the host."""
import numpy as np
import pytest

from tests.oracle_lib import load_golden
from tests.test_gmres import gpu_ctx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,neqn", [("box5_transpose", 5), ("box4_fr_transpose", 9)])
def test_gpu_crs_transpose_vs_reference(name, neqn):
    from proteuscfd_b200 import capi
    ctx, g, meta = gpu_ctx(name, neqn)
    ctx.set_field(capi.F_A, g["A"])
    ctx.crs_transpose()
    assert np.array_equal(ctx.get_field(capi.F_A), g["A_T"])
    assert ctx.get_ghost_blocks().shape == (0, neqn, neqn)
    ctx.crs_transpose()
    assert np.array_equal(ctx.get_field(capi.F_A), g["A"])
    # the transposed matrix is a matrix like any other: the context's own Jacobian (bit-exact for the perfect gas),
    # transposed, equals the reference's; and a factored matrix is refused
    if neqn == 5:
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, g["q0"])
        ctx.set_field(capi.F_TIMESTEP, g["timestep"])
        ctx.jacobian()
        ctx.crs_transpose()
        assert np.array_equal(ctx.get_field(capi.F_A), g["A_T"])
    ctx.prepare_sgs()
    with pytest.raises(capi.PcfdError):
        ctx.crs_transpose()


@pytest.mark.parametrize("name,nranks", [("box6_2rank_transpose", 2), ("box6_4rank_transpose", 4)])
def test_gpu_crs_transpose_on_partitions_vs_reference(name, nranks):
    """thread ranks, one context each; ghost-column blocks through pcfd_crs_ghost_blocks and the group's all-gather"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.parallel import ThreadGroup, crs_transpose
    parts = []
    for r in range(nranks):
        g, meta = load_golden(f"{name}_r{r}of{nranks}")
        mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp",
                                  "gNodeOwner", "gNodeLocalId")}
        for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
            mesh[k] = int(meta[k])
        params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]),
                      gamma=meta["gamma"], chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
        parts.append((mesh, params, g))
    out = [None] * nranks

    def fn(rank, group):
        mesh, params, g = parts[rank]
        ctx = capi.Context(mesh, params)
        try:
            ctx.set_field(capi.F_A, g["A"])
            crs_transpose(ctx, mesh, group)
            once = ctx.get_field(capi.F_A)
            crs_transpose(ctx, mesh, group)
            out[rank] = (once, ctx.get_field(capi.F_A))
        finally:
            group.allgather(None)
            ctx.close()

    ThreadGroup(nranks).run(fn)
    for r in range(nranks):
        assert np.array_equal(out[r][0], parts[r][2]["A_T"]), r
        assert np.array_equal(out[r][1], parts[r][2]["A"]), r

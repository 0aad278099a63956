"""CRS::GMRES (ucs/crs.tcc:176-415): restarted GMRES with right preconditioning on the block-CRS system.

CPU: the C restatement (oracle/pcfd_oracle.c: orc_gmres) against the solution the REFERENCE's own CRS::GMRES produced
(tests/golden/box6_gmres.npz: 5x5 blocks, block-diagonal LU preconditioner, 8 directions, 2 restarts;
box4_fr_gmres.npz: 9x9 blocks, diagonal preconditioner, 6 directions; box6_gmres_sgs / box4_fr_gmres_sgs: the SGS
preconditioner, six sweeps on a copy of the matrix per application, two directions: beyond that the residual is at round-off
and the reference's GMRES keeps adding normalised round-off directions) -- bit-exact, x and the returned norm.
GPU: pcfd_gmres through the C ABI with the fixture's A and b against the same vectors.  The matrix-vector product,
the preconditioner and the vector updates keep the reference's arithmetic per entry; the dot products are fixed-tree
parallel sums instead of the reference's sequential ones, so the bar is the north star's 1e-12 relative (of the largest
update), not bit-exactness."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_lib import _d, _i, load_golden, load_oracle

CASES = [("box6_gmres", 5), ("box4_fr_gmres", 9), ("box6_gmres_sgs", 5), ("box4_fr_gmres_sgs", 9)]


def run_oracle(lib, g, meta, neqn, cfg=None, x0=None):
    lib.orc_gmres.restype = C.c_double
    nnode, gnode = int(meta["nnode"]), int(meta["gnode"])
    pt, nd, nr = cfg or [int(v) for v in g["gmres_cfg"]]
    x = np.zeros((nnode + gnode) * neqn) if x0 is None else x0.copy()
    dq = lib.orc_gmres(nnode, gnode, neqn, nr, nd, pt, _i(g["ia"]), _i(g["ja"]), _i(g["iau"]), _d(g["A"].copy()),
                       _d(g["b"].copy()), _d(x))
    return x, dq


@pytest.mark.parametrize("name,neqn", CASES)
def test_oracle_gmres_equals_the_reference(name, neqn):
    g, meta = load_golden(name)
    x, dq = run_oracle(load_oracle(), g, meta, neqn)
    assert np.array_equal(x, g["gmres_x"]) and dq == g["gmres_dq"][0]
    assert np.abs(x).max() > 0


def test_oracle_gmres_converges_and_agrees_with_sgs():
    """block-diagonal preconditioned GMRES drives the linear residual down and lands where the reference's SGS lands"""
    g, meta = load_golden("box6_gmres")
    lib = load_oracle()
    x, dq = run_oracle(lib, g, meta, 5, cfg=(2, 20, 3))
    n = int(meta["nnode"]) * 5
    ia, ja = g["ia"], g["ja"]
    A = g["A"].reshape(-1, 5, 5)
    r = g["b"].copy().reshape(-1, 5)
    X = x.reshape(-1, 5)
    for i in range(int(meta["nnode"])):
        for k in range(ia[i], ia[i + 1]):
            r[i] -= A[k] @ X[ja[k]]
    assert np.linalg.norm(r) < 1e-8 * np.linalg.norm(g["b"])
    assert dq < 1e-7 * np.linalg.norm(g["b"]) + 1e-8
    assert np.abs(x[:n] - g["x"][:n]).max() < 5e-3 * np.abs(x).max()      # 3 SGS sweeps are not converged; same ballpark


def gpu_ctx(name, neqn):
    if neqn == 5:
        from tests.test_gpu_parity import golden_ctx
        return golden_ctx(name)
    from tests.test_gpu_fr import fr_ctx
    return fr_ctx(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,neqn", CASES)
def test_gpu_gmres_vs_reference(name, neqn):
    from proteuscfd_b200 import capi
    ctx, g, meta = gpu_ctx(name, neqn)
    pt, nd, nr = [int(v) for v in g["gmres_cfg"]]
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    ctx.blank_x()
    dq = ctx.gmres(nr, nd, pt)
    x = ctx.get_field(capi.F_X)
    ref = g["gmres_x"]
    assert np.abs(x - ref).max() <= 1e-12 * np.abs(ref).max(), np.abs(x - ref).max() / np.abs(ref).max()
    # (SGS-preconditioned: the returned |g| of the last restart is round-off of round-off)
    assert np.isclose(dq, g["gmres_dq"][0], rtol=1e-9, atol=1e-15 if pt == 4 else 1e-18)
    # own Jacobian (perfect gas: bit-exact A) gives the same answer through the whole path
    if neqn == 5:
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, g["q0"])
        ctx.set_field(capi.F_TIMESTEP, g["timestep"])
        ctx.jacobian()
        ctx.set_field(capi.F_B, g["b"])
        ctx.blank_x()
        ctx.gmres(nr, nd, pt)
        assert np.abs(ctx.get_field(capi.F_X) - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.gpu
def test_gpu_gmres_variants_and_errors():
    from proteuscfd_b200 import capi
    ctx, g, meta = gpu_ctx("box6_gmres", 5)
    lib = load_oracle()
    ctx.set_field(capi.F_A, g["A"])
    ctx.set_field(capi.F_B, g["b"])
    for cfg in ((0, 5, 1), (1, 7, 2), (2, 12, 1), (4, 2, 1)):
        ref, dq_ref = run_oracle(lib, g, meta, 5, cfg=cfg)
        ctx.blank_x()
        dq = ctx.gmres(cfg[2], cfg[1], cfg[0])
        x = ctx.get_field(capi.F_X)
        assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max(), (cfg, np.abs(x - ref).max() / np.abs(ref).max())
        assert np.isclose(dq, dq_ref, rtol=1e-6 if cfg[0] == 4 else 1e-8, atol=1e-16)
    # a non-zero initial guess is honoured (crs.tcc:246-256)
    x0 = 0.5 * g["gmres_x"]
    ref, _ = run_oracle(lib, g, meta, 5, cfg=(2, 6, 1), x0=x0)
    ctx.set_field(capi.F_X, x0)
    ctx.gmres(1, 6, 2)
    assert np.abs(ctx.get_field(capi.F_X) - ref).max() <= 1e-11 * np.abs(ref).max()
    with pytest.raises(capi.PcfdError):
        ctx.gmres(1, 5, 5)              # no such preconditioner (crs.tcc:582-585 warns and carries on with garbage)
    ctx.prepare_sgs()
    with pytest.raises(capi.PcfdError):
        ctx.gmres(1, 5, 2)              # diagonal already factored in place


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(3, 20, 2)])
def test_gpu_gmres_across_ranks_vs_oracle(oracle, cfg):
    """three slabs as thread ranks: the halo of the preconditioned vector before every product (crs.tcc:300) and the
    rank-ordered sums of the dot products through the library exchange.  Checked through the algebra: every rank sees
    the same (global) Hessenberg system, A x = b holds on every rank's rows -- ghost columns included -- to the GMRES
    tolerance, and the ghost rows of x equal the owners' rows."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    from tests.test_gpu_comm import run_threads
    nr = 3
    parts = [slab_case(6, r, nr, colored=True, cfl=5.0) for r in range(nr)]

    def body(rank, ctx, x):
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, parts[rank][2])
        ctx.timestep(want_min=False)
        ctx.jacobian()
        ctx.update_bcs()
        x.update(capi.F_Q)
        ctx.gradient()
        x.update(capi.F_QGRAD)
        ctx.limiter()
        x.update(capi.F_LIMITER)
        ctx.residual()
        ctx.blank_x()
        dq = ctx.gmres(*cfg)
        ia, ja, iau, _ = ctx.get_crs()
        return dict(dq=dq, x=ctx.get_field(capi.F_X), A=ctx.get_field(capi.F_A), b=ctx.get_field(capi.F_B), ia=ia, ja=ja)

    # the Krylov scratch is allocated on first use: do that before the ranks connect (a cudaMalloc synchronises the device,
    # which the thread ranks share)
    def prepare(ctx):
        ctx.gmres(1, 20, 0)
        if cfg[2] == 4:
            ctx.gmres(1, 2, 4)      # the preconditioner's copy of the matrix as well

    got = run_threads(parts, body, prepare=prepare)
    assert len({g["dq"] for g in got}) == 1, "every rank must see the same (global) Hessenberg system"
    from proteuscfd_b200.parallel import build_local_group_maps
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts])
    bn = np.sqrt(sum(np.dot(g["b"], g["b"]) for g in got))
    rn = 0.0
    for r, g in enumerate(got):
        nn = parts[r][0]["nnode"]
        A = g["A"].reshape(-1, 5, 5)
        X = g["x"].reshape(-1, 5)
        res = g["b"].copy().reshape(-1, 5)
        for i in range(nn):
            for k in range(g["ia"][i], g["ia"][i + 1]):
                res[i] -= A[k] @ X[g["ja"][k]]
        rn += float(np.sum(res * res))
        # ghost rows of x hold the owners' values
        packed = [pobjs[p].pack_numpy(got[p]["x"], 5) for p in range(nr)]
        chk = g["x"].copy()
        pobjs[r].unpack_numpy(chk, 5, nn, [packed[p][r] for p in range(nr)])
        assert np.array_equal(chk, g["x"])
    assert np.sqrt(rn) < 1e-8 * bn


@pytest.mark.gpu
def test_gpu_gmres_sgs_two_ranks_vs_reference():
    """The SGS-preconditioned GMRES across ranks against TWO reference processes (tests/golden/box8_2rank_gmres_sgs_r*of2:
    the reference's own CRS::GMRES with its MPI halos and all-reduces): block-Jacobi sweeps with a halo of the iterate after
    every sweep, the preconditioned vector exchanged before every product, rank-ordered dot products -- every rank's x,
    ghost rows included, to 1e-10 of its scale (parallel sums in a different order than MPI's)."""
    from proteuscfd_b200 import capi
    from tests.test_gpu_comm import run_threads
    parts = []
    for r in (0, 1):
        g, meta = load_golden(f"box8_2rank_gmres_sgs_r{r}of2")
        mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp",
                                  "gNodeOwner", "gNodeLocalId")}
        for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
            mesh[k] = int(meta[k])
        params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]),
                      gamma=meta["gamma"], chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
        parts.append((mesh, params, g))
    pt, nd, nr = [int(v) for v in parts[0][2]["gmres_cfg"]]
    assert pt == 4

    def prepare(ctx):      # Krylov scratch and the preconditioner's copy of the matrix: allocated before the ranks connect
        ctx.gmres(1, nd, 0)
        ctx.gmres(1, nd, 4)

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
    assert got[0][1] == got[1][1]

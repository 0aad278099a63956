"""Multi-rank numerics on the GPU against fixtures written by MULTI-RANK runs of the reference (2 and 3 ranks over the
process-based MPI shim).  All ranks' contexts live on one GPU and exchange halos by direct puts into each other's
ghost segments (LoopbackExchange: the same k_halo_pack-into-peer-memory path the IPC exchange uses).  Bit-exact,
ghost rows included."""
import numpy as np
import pytest

from tests.oracle_lib import load_golden
from tests.test_oracle import exact
from tests.test_parallel_maps import CASES, load_ranks

pytestmark = pytest.mark.gpu


def make_ranks(name):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.parallel import LoopbackExchange, build_local_group_maps
    ranks = load_ranks(name)
    ctxs = []
    for g, meta in ranks:
        mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol", "ipsp", "psp")}
        for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
            mesh[k] = int(meta[k])
        params = dict(sorder=int(meta["sorder"]), limiter=int(meta["limiter"]), no_cvbc=int(meta["no_cvbc"]),
                      gamma=meta["gamma"], chi=meta["chi"], cfl=meta["cfl"], qinf=g["qinf"])
        ctxs.append(capi.Context(mesh, params))
    pobjs = build_local_group_maps([(g["gNodeOwner"], g["gNodeLocalId"]) for g, _ in ranks])
    return ranks, ctxs, LoopbackExchange(ctxs, pobjs)


def each(ctxs, fn):
    for c in ctxs:
        fn(c)


def check(ranks, ctxs, field, key, what=None):
    for r, ((g, _), c) in enumerate(zip(ranks, ctxs)):
        got = c.get_field(field)
        exact(got[: g[key].size], g[key], f"{what or key} rank {r}")


@pytest.mark.parametrize("name", sorted(CASES))
def test_multirank_iteration_matches_reference(name):
    from proteuscfd_b200 import capi
    ranks, ctxs, x = make_ranks(name)
    nsgs = int(ranks[0][1]["nSgs"])
    # ComputeNodeLSQCoefficients + halos of s, sw (gradient.tcc:115-138)
    each(ctxs, lambda c: c.lsq_coefficients())
    x.update(capi.F_LSQ_S)
    x.update(capi.F_LSQ_SW)
    check(ranks, ctxs, capi.F_LSQ_S, "lsq_s")
    check(ranks, ctxs, capi.F_LSQ_SW, "lsq_sw")
    # NewtonIterate head: UpdateBCs + halo(q)
    for (g, _), c in zip(ranks, ctxs):
        c.set_field(capi.F_Q, g["q_pre"])
    each(ctxs, lambda c: c.update_bcs())
    x.update(capi.F_Q)
    check(ranks, ctxs, capi.F_Q, "q0")
    for (g, _), c in zip(ranks, ctxs):
        assert c.timestep() == float(g["dtmin"][0])      # ComputeTimesteps returns the rank-local minimum
    check(ranks, ctxs, capi.F_TIMESTEP, "timestep")
    each(ctxs, lambda c: c.gradient())
    x.update(capi.F_QGRAD)
    check(ranks, ctxs, capi.F_QGRAD, "qgrad")
    each(ctxs, lambda c: c.limiter())
    x.update(capi.F_LIMITER)
    check(ranks, ctxs, capi.F_LIMITER, "limiter")
    sums = [c.residual(want_norms=True) for c in ctxs]
    check(ranks, ctxs, capi.F_B, "b")
    # ParallelL2Norm across ranks (parallel.h:160-181): sqrt(sum over ranks)/N_global
    ntot = sum(int(m["nnode"]) for _, m in ranks) * 5
    res = np.sqrt(sum(s[0] for s in sums)) / ntot
    assert np.isclose(res, ranks[0][0]["resnorm"][0], rtol=1e-13)
    if nsgs > 0:
        each(ctxs, lambda c: c.jacobian())
        check(ranks, ctxs, capi.F_A, "A")
        each(ctxs, lambda c: c.prepare_sgs())
        check(ranks, ctxs, capi.F_A, "A_lu")
        for (g, _), c in zip(ranks, ctxs):
            exact(c.get_crs()[3], g["pv"], "pv")
        each(ctxs, lambda c: c.blank_x())
        x.update(capi.F_X)                       # crs.tcc:88
        for _ in range(nsgs):                    # block-Jacobi across ranks: ghosts frozen during a sweep
            each(ctxs, lambda c: c.sgs(1, want_ddq=False))
            x.update(capi.F_X)                   # crs.tcc:146
        check(ranks, ctxs, capi.F_X, "x")
        each(ctxs, lambda c: c.apply_dq())
    else:
        each(ctxs, lambda c: c.explicit_solve())
        check(ranks, ctxs, capi.F_X, "x")
    x.update(capi.F_Q)
    check(ranks, ctxs, capi.F_Q, "q1")


@pytest.mark.parametrize("colored,implicit,fused", [(False, False, False), (True, True, False), (False, False, True),
                                                    (True, True, True)])
def test_slab_partitions_vs_oracle(oracle, colored, implicit, fused):
    """Partitions generated in memory (cases.slab_case, the bench's multi-GPU input): three ranks on one GPU with
    direct-put halos against the C oracle run per rank with a numpy halo exchange through the same maps."""
    from proteuscfd_b200.cases import slab_case
    nr = 3
    parts = [slab_case(7, r, nr, colored=colored, cfl=5.0 if implicit else 0.5) for r in range(nr)]
    run_partitions_vs_oracle(oracle, parts, implicit, fused)


@pytest.mark.parametrize("nr,implicit,fused", [(3, True, True), (5, False, False)])
def test_rcb_partitions_vs_oracle(oracle, nr, implicit, fused):
    """General (non-slab) partitions: recursive-bisection node partition of the box laid out as udecomp does
    (proteuscfd_b200/partition.py), ranks with several neighbours each; same bit-exact bar."""
    from proteuscfd_b200.cases import partitioned_box_case
    parts = partitioned_box_case(7, nr, cfl=5.0 if implicit else 0.5)
    assert max(len(np.unique(m["gNodeOwner"])) for m, _, _ in parts) >= 2
    run_partitions_vs_oracle(oracle, parts, implicit, fused)


def run_partitions_vs_oracle(oracle, parts, implicit, fused):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.parallel import LoopbackExchange, build_local_group_maps
    from tests.oracle_lib import oracle_for
    nr = len(parts)
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts])
    orcs = [oracle_for(oracle, m, p) for m, p, _ in parts]
    ctxs = [capi.Context(m, p) for m, p, _ in parts]
    x = LoopbackExchange(ctxs, build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts]))
    nn = [m["nnode"] for m, _, _ in parts]

    def halo(arrs, w):
        packed = [pobjs[r].pack_numpy(arrs[r], w) for r in range(nr)]
        for r in range(nr):
            pobjs[r].unpack_numpy(arrs[r], w, nn[r], [packed[p][r] for p in range(nr)])

    beta = np.zeros(1)
    qs = [q.copy() for _, _, q in parts]
    sws = []
    ss = []
    for o in orcs:
        s_, sw_ = o.lsq()
        ss.append(s_); sws.append(sw_)
    halo(sws, 6)
    each(ctxs, lambda c: c.lsq_coefficients())
    x.update(capi.F_LSQ_SW)
    for r in range(nr):
        ctxs[r].set_field(capi.F_Q, qs[r])
        exact(ctxs[r].get_field(capi.F_LSQ_SW), sws[r], f"sw rank {r}")
    for it in range(2):
        dts = [orcs[r].timestep(qs[r], beta)[0] for r in range(nr)]
        if implicit:
            crs = [o.crs_init() for o in orcs]
            As = [orcs[r].jacobian(qs[r], beta, dts[r], *crs[r]) for r in range(nr)]
        for r in range(nr):
            orcs[r].update_bcs(qs[r], beta)
        halo(qs, 10)
        grads = [orcs[r].gradient(qs[r], sws[r]) for r in range(nr)]
        halo(grads, 27)
        lims = [orcs[r].limiter(qs[r], grads[r]) for r in range(nr)]
        halo(lims, 5)
        bs = [orcs[r].residual(qs[r], grads[r], lims[r], beta) for r in range(nr)]
        each(ctxs, lambda c: c.timestep(want_min=False))
        if implicit:
            each(ctxs, lambda c: c.jacobian())
        each(ctxs, lambda c: c.update_bcs())
        x.update(capi.F_Q)
        each(ctxs, lambda c: c.gradient())
        x.update(capi.F_QGRAD)
        if fused:
            # pcfd_limiter_raw -> halo of the raw limiter -> pcfd_residual_fused (clamp + residual + clip test in one pass)
            each(ctxs, lambda c: c.limiter_raw())
            x.update(capi.F_LIMITER)
            hits = [c.residual_fused()[1] for c in ctxs]
            assert not any(hits), "the smooth state must not trigger the pressure clip"
        else:
            each(ctxs, lambda c: c.limiter())
            x.update(capi.F_LIMITER)
            each(ctxs, lambda c: c.residual())
        for r in range(nr):
            exact(ctxs[r].get_field(capi.F_QGRAD), grads[r], f"qgrad rank {r} it {it}")
            exact(ctxs[r].get_field(capi.F_LIMITER), lims[r], f"limiter rank {r} it {it}")
            exact(ctxs[r].get_field(capi.F_B), bs[r], f"b rank {r} it {it}")
        if implicit:
            xs = []
            for r in range(nr):
                exact(ctxs[r].get_field(capi.F_A), As[r], f"A rank {r}")
                pv = orcs[r].prepare_sgs(crs[r][2], As[r])
                xs.append((pv, np.zeros((nn[r] + parts[r][0]["gnode"]) * 5)))
            each(ctxs, lambda c: c.prepare_sgs())
            each(ctxs, lambda c: c.blank_x())
            x.update(capi.F_X)
            for sweep in range(3):
                for r in range(nr):
                    # one sweep of the oracle continuing from the current x (ghost values frozen)
                    xr = xs[r][1]
                    orcs[r].lib.orc_sgs.restype = __import__("ctypes").c_double
                    from tests.oracle_lib import _d, _i
                    import ctypes as C
                    orcs[r].lib.orc_sgs(C.byref(orcs[r].c), 1, _i(crs[r][0]), _i(crs[r][1]), _i(crs[r][2]), _d(As[r]),
                                        _i(xs[r][0]), _d(bs[r]), _d(xr))
                halo([xx[1] for xx in xs], 5)
                each(ctxs, lambda c: c.sgs(1, want_ddq=False))
                x.update(capi.F_X)
            for r in range(nr):
                exact(ctxs[r].get_field(capi.F_X), xs[r][1], f"x rank {r} it {it}")
                orcs[r].apply_dq(qs[r], xs[r][1])
            each(ctxs, lambda c: c.apply_dq())
        else:
            for r in range(nr):
                orcs[r].explicit_solve(qs[r], bs[r], dts[r])
            each(ctxs, lambda c: c.explicit_solve())
        halo(qs, 10)
        x.update(capi.F_Q)
        for r in range(nr):
            exact(ctxs[r].get_field(capi.F_Q), qs[r], f"q rank {r} it {it}")


@pytest.mark.parametrize("colored,viscous,fused", [(True, False, False), (False, False, False), (True, True, False),
                                                   (True, False, True), (True, True, True)])
def test_fr_slab_partitions_vs_oracle(oracle, colored, viscous, fused):
    """The reacting eqnset on partitions (BASELINE configs[4] is an 8-partition case): two z-slabs on one GPU with
    direct-put halos of q (21 wide), qgrad (42), limiter and x (9), stepped as DistributedHotPath does, against
    the FR oracle run per rank with a numpy halo exchange through the same maps.  Frozen chemistry: one implicit
    iteration (2 sweeps, block-Jacobi across partitions) and one explicit iteration, bit-exact; with the viscous terms
    (compressibleNSFR) to 1e-11 (transport libm)."""
    import ctypes as C
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_slab_case
    from proteuscfd_b200.parallel import LoopbackExchange, build_local_group_maps
    from tests.oracle_lib import _d, _i
    from tests.test_gpu_fr import fixture_fr_params, oracle_for_fr
    nr, NEQ, NV, NT = 2, 9, 21, 14
    fr, g, meta = fixture_fr_params("box4_nsfr_implicit" if viscous else "box4_fr_implicit", rxn_on=0)
    parts = [fr_slab_case(6, r, nr, fr, colored=colored) for r in range(nr)]
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _, _ in parts])
    orcs = [oracle_for_fr(oracle, m, p, g, meta) for m, p, _, _ in parts]
    ctxs = [capi.Context(m, p) for m, p, _, _ in parts]
    x = LoopbackExchange(ctxs, build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _, _ in parts]))
    nn = [m["nnode"] for m, _, _, _ in parts]
    nl = [m["nnode"] + m["gnode"] for m, _, _, _ in parts]

    def halo(arrs, w):
        packed = [pobjs[r].pack_numpy(arrs[r], w) for r in range(nr)]
        for r in range(nr):
            pobjs[r].unpack_numpy(arrs[r], w, nn[r], [packed[p][r] for p in range(nr)])

    def same(a, ref, what):
        if not viscous:
            return exact(a, ref, what)
        a, ref = np.asarray(a).reshape(-1, NEQ), np.asarray(ref).reshape(-1, NEQ)
        err = np.abs(a - ref).max(axis=0) / np.abs(ref).max(axis=0)
        assert np.all(err <= 1e-10), f"{what}: relative error per equation {err}"

    qs = [q.copy() for _, _, q, _ in parts]
    betas = [b[: nl[r]].copy() for r, (_, _, _, b) in enumerate(parts)]
    sws = [o.lsq()[1] for o in orcs]
    halo(sws, 6)
    for r in range(nr):
        ctxs[r].set_field(capi.F_BETA, parts[r][3])
        ctxs[r].lsq_coefficients()
    x.update(capi.F_LSQ_S)
    x.update(capi.F_LSQ_SW)
    for r in range(nr):
        ctxs[r].set_field(capi.F_Q, qs[r])
        exact(ctxs[r].get_field(capi.F_LSQ_SW), sws[r], f"sw rank {r}")

    def oracle_head():
        for r in range(nr):
            orcs[r].update_bcs(qs[r], betas[r])
        halo(qs, NV)
        grads = [orcs[r].gradient(qs[r], sws[r]) for r in range(nr)]
        halo(grads, NT * 3)
        lims = [orcs[r].limiter(qs[r], grads[r]) for r in range(nr)]
        halo(lims, NEQ)
        return grads, lims, [orcs[r].residual(qs[r], grads[r], lims[r], betas[r]) for r in range(nr)]

    # ---- implicit iteration
    dts = [orcs[r].timestep(qs[r], betas[r])[0] for r in range(nr)]
    crs = [o.crs_init() for o in orcs]
    As = [orcs[r].jacobian(qs[r], betas[r], dts[r], *crs[r]) for r in range(nr)]
    grads, lims, bs = oracle_head()
    xs = []
    for r in range(nr):
        pv = orcs[r].prepare_sgs(crs[r][2], As[r])
        xs.append((pv, np.zeros(nl[r] * NEQ)))
    for sweep in range(2):
        for r in range(nr):
            orcs[r].lib.orc_fr_sgs.restype = C.c_double
            orcs[r].lib.orc_fr_sgs(C.byref(orcs[r].c), C.byref(orcs[r].p), 1, _i(crs[r][0]), _i(crs[r][1]), _i(crs[r][2]),
                                   _d(As[r]), _i(xs[r][0]), _d(bs[r]), _d(xs[r][1]))
        halo([xx[1] for xx in xs], NEQ)
    for r in range(nr):
        orcs[r].apply_dq(qs[r], xs[r][1])
    halo(qs, NV)
    # the hot-path driver calls its exchange per rank; with all ranks in one process the loopback exchange moves every
    # rank's rows at once, so the phases are stepped in lock-step here exactly as DistributedHotPath.implicit_iterate does
    for c in ctxs:
        c.timestep(want_min=False)
        c.jacobian()
        c.update_bcs()
    x.update(capi.F_Q)
    each(ctxs, lambda c: c.gradient())
    x.update(capi.F_QGRAD)
    if fused:
        # pcfd_limiter_raw -> halo of the raw limiter -> pcfd_residual_fused (clamp + residual + clip test in one pass)
        each(ctxs, lambda c: c.limiter_raw())
        x.update(capi.F_LIMITER)
        hits = [c.residual_fused()[1] for c in ctxs]
        assert not any(hits), "the smooth state must not trigger the pressure clip"
        assert all(c.clip_fallbacks() == 0 for c in ctxs)
    else:
        each(ctxs, lambda c: c.limiter())
        x.update(capi.F_LIMITER)
        each(ctxs, lambda c: c.residual())
    each(ctxs, lambda c: c.prepare_sgs())
    each(ctxs, lambda c: c.blank_x())
    x.update(capi.F_X)
    for sweep in range(2):
        each(ctxs, lambda c: c.sgs(1, want_ddq=False))
        x.update(capi.F_X)
    each(ctxs, lambda c: c.apply_dq())
    x.update(capi.F_Q)
    for r in range(nr):
        exact(ctxs[r].get_field(capi.F_QGRAD), grads[r], f"qgrad rank {r}")
        exact(ctxs[r].get_field(capi.F_LIMITER), lims[r], f"limiter rank {r}")
        same(ctxs[r].get_field(capi.F_B), bs[r], f"b rank {r}")
        same(ctxs[r].get_field(capi.F_X), xs[r][1], f"x rank {r}")
        if viscous:
            assert np.allclose(ctxs[r].get_field(capi.F_Q), qs[r], rtol=1e-10, atol=1e-14)
        else:
            exact(ctxs[r].get_field(capi.F_Q), qs[r], f"q rank {r} after the implicit iteration")
    if viscous:
        return
    # ---- explicit iteration
    for r in range(nr):
        orcs[r].c.cfl = 0.05
        ctxs[r].set_cfl(0.05)
    dts = [orcs[r].timestep(qs[r], betas[r])[0] for r in range(nr)]
    grads, lims, bs = oracle_head()
    for r in range(nr):
        xe = orcs[r].explicit_solve(qs[r], bs[r], dts[r])
        orcs[r].apply_dq(qs[r], xe)
    halo(qs, NV)
    each(ctxs, lambda c: c.timestep(want_min=False))
    each(ctxs, lambda c: c.update_bcs())
    x.update(capi.F_Q)
    each(ctxs, lambda c: c.gradient())
    x.update(capi.F_QGRAD)
    each(ctxs, lambda c: c.limiter())
    x.update(capi.F_LIMITER)
    each(ctxs, lambda c: c.residual())
    each(ctxs, lambda c: c.explicit_solve())
    x.update(capi.F_Q)
    for r in range(nr):
        exact(ctxs[r].get_field(capi.F_B), bs[r], f"explicit b rank {r}")
        exact(ctxs[r].get_field(capi.F_Q), qs[r], f"q rank {r} after the explicit iteration")


def test_nccl_two_process_exchange():
    """The real thing: one process per GPU, NCCL send/recv into the ghost segments and the direct-put exchange,
    launched with torchrun when the box has >= 2 GPUs (gpurun --gpus 2)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29713", os.path.join(root, "tests", "nccl_worker.py"),
                        "box8_2rank_explicit"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK") == 2, r.stdout[-3000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29714", os.path.join(root, "tests", "nccl_worker.py"),
                        "box8_2rank_explicit", "put"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK") == 2, r.stdout[-3000:]


@pytest.mark.parametrize("nr,wall_tag", [(2, 5), (3, 3)])
def test_sa_slab_partitions_vs_oracle(oracle, nr, wall_tag):
    """Spalart-Allmaras across partitions (BASELINE configs[3]; TurbulenceModel::Compute, ucs/turb.tcc:163-339 with its
    exchanges at :185, gradient.tcc:98, crs.tcc:146, :325): z-slabs of a laminar-NS + SA box on one GPU, stepped through
    DistributedHotPath.turb_compute's phase / halo order with direct-put halos, against the oracle's
    orc_turb_sa_phase replayed per rank with a numpy halo through the same maps.  wall_tag 5: the no-slip floor lives
    in rank 0 only (the other ranks' wall distance comes from rank 0's wall nodes); 3: every slab owns a strip of wall.
    tgrad and the matrix pattern bit-exact, everything downstream of Sutherland's law / the source term's libm calls to
    1e-12 per node (tests/test_gpu_viscous.py)."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import NS_BC, slab_case
    from proteuscfd_b200.parallel import LoopbackExchange, build_local_group_maps
    from proteuscfd_b200.walldist import nearest_distance, wall_points
    from tests.oracle_lib import oracle_for
    from tests.test_gpu_viscous import close_per_node
    nsgs = 3
    bc = dict(NS_BC)
    bc[5] = capi.BC_SYMMETRY
    bc[wall_tag] = capi.BC_NOSLIP
    parts = [slab_case(6, r, nr, viscous=True, turb=True, colored=True, cfl=5.0, bc=bc) for r in range(nr)]
    maps = [(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts]
    pobjs = build_local_group_maps(maps)
    orcs = [oracle_for(oracle, m, p) for m, p, _ in parts]
    ctxs = [capi.Context(m, p) for m, p, _ in parts]
    x = LoopbackExchange(ctxs, build_local_group_maps(maps))
    nn = [m["nnode"] for m, _, _ in parts]
    nl = [m["nnode"] + m["gnode"] for m, _, _ in parts]

    def halo(arrs, w):
        packed = [pobjs[r].pack_numpy(arrs[r], w) for r in range(nr)]
        for r in range(nr):
            pobjs[r].unpack_numpy(arrs[r], w, nn[r], [packed[p][r] for p in range(nr)])

    # wall distance: nearest viscous wall node of ANY rank (walldist.tcc:24-113)
    pts = np.concatenate([wall_points(m) for m, _, _ in parts])
    assert len(pts) > 0
    dist = [nearest_distance(m["xyz"].reshape(-1, 3)[: nl[r]], pts) for r, (m, _, _) in enumerate(parts)]
    # the flow side of the model's inputs from the oracle (bit-exact against the GPU elsewhere in this file)
    beta = np.zeros(1)
    qs = [q.copy() for _, _, q in parts]
    ss_, sws = zip(*[o.lsq() for o in orcs])
    ss_, sws = list(ss_), list(sws)
    halo(ss_, 6)
    halo(sws, 6)
    for r in range(nr):
        orcs[r].update_bcs(qs[r], beta)
    halo(qs, 10)
    dts = [orcs[r].timestep(qs[r], beta)[0] for r in range(nr)]
    grads = [orcs[r].gradient(qs[r], sws[r]) for r in range(nr)]
    halo(grads, 27)
    crs = [o.crs_init() for o in orcs]
    rng = np.random.default_rng(7)
    sts = []
    for r, (m, _, _) in enumerate(parts):
        # nu~ = free-stream value modulated smoothly in space (consistent across ranks: a function of the coordinates)
        X = m["xyz"].reshape(-1, 3)[: nl[r]]
        tv = np.zeros(nl[r] + m["nbnode"])
        tv[: nl[r]] = 1.341946 * (1.0 + 0.3 * np.sin(2 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 2]) + 0.2 * X[:, 1])
        sts.append(orcs[r].turb_sa_state(tv))
        c = ctxs[r]
        c.set_field(capi.F_Q, qs[r])
        c.set_field(capi.F_QGRAD, grads[r])
        c.set_field(capi.F_TIMESTEP, dts[r])
        c.set_field(capi.F_LSQ_S, ss_[r])
        c.set_field(capi.F_WALLDIST, dist[r])
        c.set_field(capi.F_TVAR, tv.copy())

    def ophase(ph):
        return [orcs[r].turb_sa_phase(ph, nsgs, qs[r], grads[r], ss_[r], dist[r], dts[r], *crs[r], sts[r]) for r in range(nr)]

    # ---- oracle, rank by rank, halos at the reference's exchange points
    ophase(0)
    halo([s["tvar"] for s in sts], 1)
    ophase(1)
    halo([s["tgrad"] for s in sts], 3)
    osum = sum(ophase(2))
    for _ in range(nsgs):
        ophase(3)
        halo([s["x"] for s in sts], 1)
    ophase(4)
    halo([s["tvar"] for s in sts], 1)
    ophase(5)
    # ---- GPU: the order of DistributedHotPath.turb_compute, all ranks in lock-step
    each(ctxs, lambda c: c.turb_phase(0))
    x.update(capi.F_TVAR)
    each(ctxs, lambda c: c.turb_phase(1))
    x.update(capi.F_TGRAD)
    gsum = sum(c.turb_phase(2, want_norm=True) for c in ctxs)
    for _ in range(nsgs):
        each(ctxs, lambda c: c.turb_phase(3))
        x.update(capi.F_TURB_X)
    each(ctxs, lambda c: c.turb_phase(4))
    x.update(capi.F_TVAR)
    each(ctxs, lambda c: c.turb_phase(5))
    assert np.isclose(gsum, osum, rtol=1e-11)
    moved = 0.0
    for r in range(nr):
        c, s = ctxs[r], sts[r]
        exact(c.get_field(capi.F_TGRAD), s["tgrad"], f"tgrad rank {r} (ghost rows included)")
        close_per_node(c.get_field(capi.F_TURB_B), s["b"], 1, f"turbulence residual rank {r}")
        close_per_node(c.get_field(capi.F_TURB_A), s["A"], 1, f"turbulence matrix rank {r} (ghost columns included)")
        close_per_node(c.get_field(capi.F_TURB_X), s["x"], 1, f"turbulence update rank {r} (ghost rows included)")
        close_per_node(c.get_field(capi.F_TVAR)[: nl[r]], s["tvar"][: nl[r]], 1, f"nu~ rank {r}")
        close_per_node(c.get_field(capi.F_MUT)[: nl[r]], s["mut"], 1, f"eddy viscosity rank {r}")
        moved = max(moved, np.abs(s["x"][: nn[r]]).max())
        # the ghost columns of the scalar matrix are live: the test would not notice a dropped halo otherwise
        ia, ja, _ = crs[r]
        assert np.abs(s["A"][ja >= nn[r]]).max() > 0
    assert moved > 1e-6

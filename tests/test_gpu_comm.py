"""The library's own multi-rank path (csrc/pcfd_comm.cuh; ABI v7): flag-based direct puts into peer ghost segments and
the composite entry points (pcfd_explicit_iterate / pcfd_implicit_iterate / pcfd_turb_compute / pcfd_lsq_coefficients)
running the reference's multi-rank sequence themselves once pcfd_comm_connect has been called.

Single-GPU box: the ranks are THREADS of this process, one context each, all on cuda:0 -- the same kernels, flags and
epochs as one process per GPU, with plain pointers where the other case goes through CUDA IPC; a second test runs TWO
PROCESSES on the one GPU (gloo control plane) so that the IPC mapping itself is exercised.  With two or more GPUs the
processes spread over them (tests/comm_worker.py uses LOCAL_RANK % device count).

Bar: bit-exact against the C oracle replayed rank by rank with a numpy halo through the reference's maps
(tests/partition_oracle.py) -- ghost rows included."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.test_oracle import exact

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_threads(parts, body, prepare=None):
    """one thread per rank: context, halo maps over the thread group, CommExchange; body(rank, ctx, x) -> result.
    prepare(ctx) or prepare(rank, ctx): anything that allocates device memory on first use, run BEFORE the ranks connect"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.parallel import CommExchange, PObj, ThreadGroup
    nr = len(parts)
    out = [None] * nr

    def fn(rank, group):
        mesh, params = parts[rank][0], parts[rank][1]
        ctx = capi.Context(mesh, params)
        # allocate-on-first-use fields now: a cudaMalloc inside an iteration synchronises the DEVICE, i.e. (ranks as
        # threads sharing one GPU) it would wait for a peer's put kernel that is waiting for this very rank
        ctx.device_ptr(capi.F_A)
        if prepare is not None:
            prepare(ctx) if prepare.__code__.co_argcount == 1 else prepare(rank, ctx)
        pobj = PObj(rank, nr).BuildCommMaps(mesh["gNodeOwner"], mesh["gNodeLocalId"], group)
        x = CommExchange(ctx, pobj, group)
        try:
            out[rank] = body(rank, ctx, x)
            ctx.synchronize()
        finally:
            group.allgather(None)        # nobody unmaps while a peer may still be writing
            x.close()
            ctx.close()

    ThreadGroup(nr).run(fn)
    return out


@pytest.mark.parametrize("kind,nr,implicit", [("slab", 3, False), ("slab", 3, True), ("rcb", 5, False), ("rcb", 4, True)])
def test_comm_composite_iterations_vs_oracle(oracle, kind, nr, implicit):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import partitioned_box_case, slab_case
    from tests.partition_oracle import replay_perfect_gas
    cfl = 5.0 if implicit else 0.5
    if kind == "slab":
        parts = [slab_case(7, r, nr, colored=implicit, cfl=cfl) for r in range(nr)]
    else:
        parts = partitioned_box_case(7, nr, cfl=cfl)
    nsw = 3
    sws, ref = replay_perfect_gas(oracle, parts, implicit, iters=2, nsweeps=nsw)

    def body(rank, ctx, x):
        ctx.lsq_coefficients()                     # halos of s and sw inside (gradient.tcc:131-134)
        recs = [dict(sw=ctx.get_field(capi.F_LSQ_SW))]
        ctx.set_field(capi.F_Q, parts[rank][2])
        for _ in range(2):
            if implicit:
                ctx.implicit_iterate(nsw, refresh_jac=True)
            else:
                ctx.explicit_iterate(refresh_dt=True)
            recs.append({k: ctx.get_field(f) for k, f in (("qgrad", capi.F_QGRAD), ("limiter", capi.F_LIMITER), ("b", capi.F_B),
                                                           ("x", capi.F_X), ("q", capi.F_Q))})
        assert ctx.clip_fallbacks() == 0
        return recs

    got = run_threads(parts, body)
    for r in range(nr):
        exact(got[r][0]["sw"], sws[r], f"sw rank {r}")
        for it in range(2):
            g, o = got[r][it + 1], ref[it]
            exact(g["qgrad"], o["qgrad"][r], f"qgrad rank {r} it {it}")
            exact(g["limiter"], o["limiter"][r], f"limiter rank {r} it {it}")
            exact(g["b"], o["b"][r], f"b rank {r} it {it}")
            if implicit:
                exact(g["x"], o["x"][r], f"x rank {r} it {it}")
            exact(g["q"], o["q"][r], f"q rank {r} it {it}")


def test_comm_pressure_clip_is_a_global_decision(oracle):
    """a rough state on ONE rank makes its fused flux kernel raise the clip flag: every rank must fall back to the
    ordered clip path (the flag words go round through the peers' flag pages) and the result must still be the
    oracle's"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    from tests.partition_oracle import replay_perfect_gas
    nr = 3
    parts = [list(slab_case(6, r, nr, cfl=0.5)) for r in range(nr)]
    # roughen a patch of owned nodes of rank 1 (consistent ghost copies follow from the first halo of q)
    m, _, q = parts[1]
    Q = q.reshape(-1, 10)
    rng = np.random.default_rng(7)
    nn = m["nnode"]
    # only the two middle planes of rank 1's slab: ranks 0 and 2 never see a rough value, their own flags stay down
    mid = np.isin(m["gid"][:nn] // 49, (8, 9))
    Q[:nn, 0] *= np.where(mid, rng.choice([0.05, 1.0, 4.0], size=nn), 1.0)   # the state of test_pressure_clip_sequential_semantics
    Q[:nn, 4] *= np.where(mid, rng.choice([0.3, 1.0, 6.0], size=nn), 1.0)
    from proteuscfd_b200.cases import aux_vars
    aux_vars(Q, parts[1][1]["gamma"])
    parts = [tuple(p) for p in parts]
    sws, ref = replay_perfect_gas(oracle, parts, False, iters=1)

    def body(rank, ctx, x):
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, parts[rank][2])
        ctx.explicit_iterate(refresh_dt=True)
        return dict(fallbacks=ctx.clip_fallbacks(), limiter=ctx.get_field(capi.F_LIMITER), b=ctx.get_field(capi.F_B),
                    q=ctx.get_field(capi.F_Q))

    got = run_threads(parts, body)
    assert [g["fallbacks"] for g in got] == [1] * nr, "the clip fallback must be taken by every rank or by none"
    for r in range(nr):
        exact(got[r]["limiter"], ref[0]["limiter"][r], f"limiter rank {r}")
        exact(got[r]["b"], ref[0]["b"][r], f"b rank {r}")
        exact(got[r]["q"], ref[0]["q"][r], f"q rank {r}")
    zeros = [int((ref[0]["limiter"][r].reshape(-1, 5).max(axis=1) == 0.0).sum()) for r in range(nr)]
    assert zeros[1] > 10, zeros       # (a few ghost rows of the neighbours inherit zeros through the limiter halo)


def test_comm_allgather_and_error_paths():
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    nr = 4
    parts = [slab_case(4, r, nr) for r in range(nr)]

    def body(rank, ctx, x):
        a = x.allgather([rank + 0.5, -rank])
        b = x.allgather([10.0 * rank])            # a second round uses the other parity slot
        c = x.allgather([7.0])
        with pytest.raises(capi.PcfdError):
            ctx.comm_update(capi.F_B)             # b has no ghost rows
        return a, b, c

    for a, b, c in run_threads(parts, body):
        assert np.array_equal(a, np.array([[r + 0.5, -r] for r in range(nr)]))
        assert np.array_equal(b[:, 0], 10.0 * np.arange(nr))
        assert np.array_equal(c[:, 0], np.full(nr, 7.0))
    ctx = capi.Context(*slab_case(4, 0, 2)[:2])
    with pytest.raises(capi.PcfdError):
        ctx.comm_update(capi.F_Q)                 # not connected
    with pytest.raises(capi.PcfdError):
        ctx.comm_export()                         # pcfd_halo_configure first


def test_comm_spalart_allmaras_composite_equals_lockstep_phases():
    """pcfd_implicit_iterate on a connected laminar-NS + SA context (flow iteration, then TurbulenceModel::Compute with the
    halos of tvar / tgrad / turb_x / tvar inside the library) == the same phases stepped in lock-step over the
    host-synchronised loopback exchange, which tests/test_gpu_multirank.py holds against the oracle"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import NS_BC, slab_case
    from proteuscfd_b200.parallel import DistributedHotPath, LoopbackExchange, build_local_group_maps
    from proteuscfd_b200.walldist import nearest_distance, wall_points
    nr, nsgs = 2, 3
    bc = dict(NS_BC)
    bc[5] = capi.BC_SYMMETRY
    bc[3] = capi.BC_NOSLIP
    parts = [slab_case(6, r, nr, viscous=True, turb=True, colored=True, cfl=5.0, bc=bc) for r in range(nr)]
    pts = np.concatenate([wall_points(m) for m, _, _ in parts])
    nl = [m["nnode"] + m["gnode"] for m, _, _ in parts]
    dist = [nearest_distance(m["xyz"].reshape(-1, 3)[: nl[r]], pts) for r, (m, _, _) in enumerate(parts)]
    tv0 = []
    for r, (m, _, _) in enumerate(parts):
        X = m["xyz"].reshape(-1, 3)[: nl[r]]
        tv = np.zeros(nl[r] + m["nbnode"])
        tv[: nl[r]] = 1.341946 * (1.0 + 0.3 * np.sin(2 * np.pi * X[:, 0]) + 0.2 * X[:, 1])
        tv0.append(tv)
    fields = (("q", capi.F_Q), ("b", capi.F_B), ("x", capi.F_X), ("tvar", capi.F_TVAR), ("mut", capi.F_MUT),
              ("tx", capi.F_TURB_X), ("tgrad", capi.F_TGRAD))

    def arm(r, c):
        c.set_field(capi.F_WALLDIST, dist[r])
        c.set_field(capi.F_TVAR, tv0[r])
        c.set_field(capi.F_Q, parts[r][2])

    def body(rank, ctx, x):
        ctx.lsq_coefficients()
        arm(rank, ctx)
        for _ in range(2):
            ctx.implicit_iterate(nsgs, refresh_jac=True)
        return {k: ctx.get_field(f) for k, f in fields}

    got = run_threads(parts, body)
    # lock-step reference on the loopback exchange
    ctxs = [capi.Context(m, p) for m, p, _ in parts]
    x = LoopbackExchange(ctxs, build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts]))
    for c in ctxs:
        c.lsq_coefficients()
    x.update(capi.F_LSQ_S)
    x.update(capi.F_LSQ_SW)
    for r, c in enumerate(ctxs):
        arm(r, c)
    each = lambda fn: [fn(c) for c in ctxs]
    for _ in range(2):
        each(lambda c: (c.timestep(want_min=False), c.jacobian(), c.update_bcs()))
        x.update(capi.F_Q)
        each(lambda c: c.gradient())
        x.update(capi.F_QGRAD)
        each(lambda c: c.limiter())
        x.update(capi.F_LIMITER)
        each(lambda c: (c.residual(), c.prepare_sgs(), c.blank_x()))
        x.update(capi.F_X)
        for _s in range(nsgs):
            each(lambda c: c.sgs(1, want_ddq=False))
            x.update(capi.F_X)
        each(lambda c: c.apply_dq())
        x.update(capi.F_Q)
        for ph, fld in ((0, capi.F_TVAR), (1, capi.F_TGRAD), (2, None)):
            each(lambda c: c.turb_phase(ph))
            if fld is not None:
                x.update(fld)
        for _s in range(nsgs):
            each(lambda c: c.turb_phase(3))
            x.update(capi.F_TURB_X)
        each(lambda c: c.turb_phase(4))
        x.update(capi.F_TVAR)
        each(lambda c: c.turb_phase(5))
    for r, c in enumerate(ctxs):
        for k, f in fields:
            exact(got[r][k], c.get_field(f), f"{k} rank {r}")
        assert np.abs(got[r]["tx"]).max() > 0 and np.abs(got[r]["mut"]).max() > 0


def test_comm_reacting_composite_vs_lockstep():
    """the reacting eqnset (9 equations, 21 variables, 42-wide gradient rows) through the library's exchange: composite
    implicit iterations on two connected slabs == lock-step phases over the loopback exchange (held against the FR
    oracle in tests/test_gpu_multirank.py)"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_slab_case
    from proteuscfd_b200.parallel import LoopbackExchange, build_local_group_maps
    from tests.test_gpu_fr import fixture_fr_params
    nr, nsgs = 2, 2
    fr, g, meta = fixture_fr_params("box4_fr_implicit", rxn_on=1)
    parts = [fr_slab_case(6, r, nr, fr) for r in range(nr)]
    fields = (("qgrad", capi.F_QGRAD), ("limiter", capi.F_LIMITER), ("b", capi.F_B), ("x", capi.F_X), ("q", capi.F_Q))

    def body(rank, ctx, x):
        ctx.set_field(capi.F_BETA, parts[rank][3])
        ctx.lsq_coefficients()
        ctx.set_field(capi.F_Q, parts[rank][2])
        ctx.implicit_iterate(nsgs, refresh_jac=True)
        ctx.implicit_iterate(nsgs, refresh_jac=False)
        return {k: ctx.get_field(f) for k, f in fields}

    got = run_threads(parts, body)
    ctxs = [capi.Context(m, p) for m, p, _, _ in parts]
    x = LoopbackExchange(ctxs, build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _, _ in parts]))
    for r, c in enumerate(ctxs):
        c.set_field(capi.F_BETA, parts[r][3])
        c.lsq_coefficients()
    x.update(capi.F_LSQ_S)
    x.update(capi.F_LSQ_SW)
    for r, c in enumerate(ctxs):
        c.set_field(capi.F_Q, parts[r][2])
    each = lambda fn: [fn(c) for c in ctxs]
    for it in range(2):
        if it == 0:
            each(lambda c: (c.timestep(want_min=False), c.jacobian()))
        each(lambda c: c.update_bcs())
        x.update(capi.F_Q)
        each(lambda c: c.gradient())
        x.update(capi.F_QGRAD)
        each(lambda c: c.limiter())
        x.update(capi.F_LIMITER)
        each(lambda c: (c.residual(), c.prepare_sgs(), c.blank_x()))
        x.update(capi.F_X)
        for _s in range(nsgs):
            each(lambda c: c.sgs(1, want_ddq=False))
            x.update(capi.F_X)
        each(lambda c: c.apply_dq())
        x.update(capi.F_Q)
    for r, c in enumerate(ctxs):
        for k, f in fields:
            exact(got[r][k], c.get_field(f), f"{k} rank {r}")


@pytest.mark.parametrize("implicit", [False, True])
def test_comm_two_processes_over_cuda_ipc(implicit):
    """one PROCESS per rank (gloo control plane, CUDA-IPC mapped fields and flag pages): both on cuda:0 on a single-GPU
    box, one GPU each when there are two"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29721" if implicit else "29720",
                        os.path.join(ROOT, "tests", "comm_worker.py"), "implicit" if implicit else "explicit"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK") == 2, r.stdout[-3000:]


def test_comm_forces_are_summed_over_the_ranks(oracle):
    """Forces::Compute on partitions: every rank integrates its own BC half-edges, the body sums and the projected body
    areas are added over the ranks (forces.tcc:236-243, 383-388: MPI_Allreduce; here pcfd_comm in rank order) and every
    rank ends up with the same coefficients.  Against the oracle's per-rank sums added in the same order."""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import slab_case
    from tests.oracle_lib import oracle_for
    nr = 3
    parts = [slab_case(6, r, nr, colored=True, cfl=5.0, viscous=True) for r in range(nr)]
    tags = [capi.BC_NOSLIP, capi.BC_FARFIELD]       # the surrogate "factag" of a half-edge is its BC type here
    V = 0.5
    fix = []
    for mesh, params, q in parts:
        nloc, nbe = mesh["nnode"] + mesh["gnode"], mesh["nbedge"]
        bn = mesh["bedges_n"].reshape(-1, 2)[:nbe]
        cg = np.zeros((nloc + mesh["nbnode"], 3))
        cg[bn[:, 1]] = mesh["xyz"].reshape(-1, 3)[bn[:, 0]] + 0.01
        fix.append(dict(forces_body_lists=np.array([1, tags[0], 1, tags[1]], dtype=np.int32),
                        forces_body_geom=np.array([0.2, 0.1, 0.0, 0, 0, 1, 0.0, 0.0, 0.0, 0, 1, 0], dtype=np.float64),
                        bedges_factag=mesh["bedges_bctype"].astype(np.int32), forces_cg=cg.reshape(-1),
                        forces_surfArea=np.zeros(3 * 10), forces_dirs=np.array([0.3, 0.2, -1.0, 1.0, 0.1, 0.0])))
    rng = np.random.default_rng(3)
    grads = [rng.standard_normal((m["nnode"] + m["gnode"]) * 27) * 0.1 for m, _, _ in parts]

    def prepare(rank, ctx):
        from tests.oracle_lib import bodies_from_fixture
        g = fix[rank]
        offs, tg, mpt, max_ = bodies_from_fixture(g)
        ctx.forces_configure(offs, tg, mpt, max_, g["bedges_factag"], g["forces_cg"], g["forces_dirs"][:3], g["forces_dirs"][3:], V, 9)

    def body(rank, ctx, x):
        ctx.set_field(capi.F_Q, parts[rank][2])
        ctx.set_field(capi.F_QGRAD, grads[rank])
        b, c = ctx.forces_compute()
        return b, c, ctx.forces_areas()[1]

    got = run_threads(parts, body, prepare=prepare)
    sums, areas, rho_inf = np.zeros(24), np.zeros(6), None
    scale = 0.0
    for r, (mesh, params, q) in enumerate(parts):
        o = oracle_for(oracle, mesh, dict(params, viscous=1))
        d = o.forces_desc(fix[r])
        _, ba = o.surface_areas(d)
        exact(got[r][2], ba, f"rank {r}: local projected areas")
        o.c.mach = V
        out = o.forces(d, q, grads[r], ba)
        sums += out["body"]
        areas += ba
        scale += np.abs(q).max() * np.abs(mesh["bedges_a"].reshape(-1, 4)[: mesh["nbedge"], 3]).sum()
    for r in range(nr):
        assert np.all(np.abs(got[r][0].ravel() - sums) <= 1e-12 * scale)
        exact(got[r][0], got[0][0], "every rank holds the same sums")
        exact(got[r][1], got[0][1], "every rank holds the same coefficients")
    assert np.abs(sums[3:6]).max() > 0 and np.isfinite(got[0][1]).all()

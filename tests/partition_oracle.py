"""Per-rank replay of the C oracle on a partitioned case with a numpy halo exchange through the reference's maps: what a
multi-rank run of the reference computes (block-Jacobi SGS across partitions, crs.tcc:88,146).  Test infrastructure
shared by the multi-rank GPU tests; not a test file."""
import ctypes as C

import numpy as np

from proteuscfd_b200.parallel import build_local_group_maps
from tests.oracle_lib import _d, _i, oracle_for


def numpy_halo(pobjs, nn):
    nr = len(pobjs)

    def halo(arrs, w):
        packed = [pobjs[r].pack_numpy(arrs[r], w) for r in range(nr)]
        for r in range(nr):
            pobjs[r].unpack_numpy(arrs[r], w, nn[r], [packed[p][r] for p in range(nr)])
    return halo


def replay_perfect_gas(oracle, parts, implicit, iters=2, nsweeps=3):
    """parts: [(mesh, params, q)] per rank.  Returns (sws, per-iteration list of dicts of per-rank arrays): the order of
    NewtonIterate -- time step (+ Jacobian), UpdateBCs, halo q, gradient, halo, limiter, halo, residual, [LU, nsweeps
    sweeps with a halo of x each, ApplyDQ | ExplicitSolve], halo q."""
    nr = len(parts)
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m, _, _ in parts])
    orcs = [oracle_for(oracle, m, p) for m, p, _ in parts]
    nn = [m["nnode"] for m, _, _ in parts]
    halo = numpy_halo(pobjs, nn)
    beta = np.zeros(1)
    qs = [q.copy() for _, _, q in parts]
    sws = [o.lsq()[1] for o in orcs]
    halo(sws, 6)
    out = []
    for _ in range(iters):
        rec = {}
        dts = [orcs[r].timestep(qs[r], beta)[0] for r in range(nr)]
        if implicit:
            crs = [o.crs_init() for o in orcs]
            As = [orcs[r].jacobian(qs[r], beta, dts[r], *crs[r]) for r in range(nr)]
            rec["A"] = [a.copy() for a in As]
        for r in range(nr):
            orcs[r].update_bcs(qs[r], beta)
        halo(qs, 10)
        grads = [orcs[r].gradient(qs[r], sws[r]) for r in range(nr)]
        halo(grads, 27)
        lims = [orcs[r].limiter(qs[r], grads[r]) for r in range(nr)]
        halo(lims, 5)
        bs = [orcs[r].residual(qs[r], grads[r], lims[r], beta) for r in range(nr)]
        rec.update(qgrad=grads, limiter=lims, b=bs, timestep=dts)
        if implicit:
            xs = []
            for r in range(nr):
                pv = orcs[r].prepare_sgs(crs[r][2], As[r])
                xs.append((pv, np.zeros((nn[r] + parts[r][0]["gnode"]) * 5)))
            for _sweep in range(nsweeps):
                for r in range(nr):      # one sweep continuing from the current x, ghost values frozen
                    orcs[r].lib.orc_sgs.restype = C.c_double
                    orcs[r].lib.orc_sgs(C.byref(orcs[r].c), 1, _i(crs[r][0]), _i(crs[r][1]), _i(crs[r][2]), _d(As[r]),
                                        _i(xs[r][0]), _d(bs[r]), _d(xs[r][1]))
                halo([xx[1] for xx in xs], 5)
            for r in range(nr):
                orcs[r].apply_dq(qs[r], xs[r][1])
            rec["x"] = [xx[1].copy() for xx in xs]
        else:
            for r in range(nr):
                orcs[r].explicit_solve(qs[r], bs[r], dts[r])
        halo(qs, 10)
        rec["q"] = [q.copy() for q in qs]
        out.append(rec)
    return sws, out

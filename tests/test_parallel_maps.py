"""Host logic of the halo exchange (proteuscfd_b200/parallel.py) on the CPU: the maps must be BIT-EXACT the
reference's (PObj::BuildCommMaps, ucs/parallel.tcc:461-554; dumped from the reference by tools/make_golden.py),
and an exchange driven by them must reproduce the ghost rows of the reference's own arrays."""
import os
import sys

import numpy as np
import pytest

from tests.oracle_lib import load_golden
from tests.test_oracle import exact

CASES = {"box8_2rank_explicit": 2, "box9_3rank_implicit": 3}


def load_ranks(name):
    n = CASES[name]
    return [load_golden(f"{name}_r{r}of{n}") for r in range(n)]


@pytest.mark.parametrize("name", sorted(CASES))
def test_comm_maps_match_reference(name):
    from proteuscfd_b200.parallel import build_local_group_maps
    ranks = load_ranks(name)
    pobjs = build_local_group_maps([(g["gNodeOwner"], g["gNodeLocalId"]) for g, _ in ranks])
    for (g, meta), p in zip(ranks, pobjs):
        exact(p.commCountsRecv, g["commCountsRecv"], "commCountsRecv")
        exact(p.commOffsetsRecv, g["commOffsetsRecv"], "commOffsetsRecv")
        exact(p.commCountsSend, g["commCountsSend"], "commCountsSend")
        exact(p.nodePackingList, g["nodePackingList"], "nodePackingList")
        assert int(p.commCountsRecv.sum()) == int(meta["gnode"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_exchange_reproduces_reference_ghost_rows(name):
    from proteuscfd_b200.parallel import build_local_group_maps
    ranks = load_ranks(name)
    n = len(ranks)
    pobjs = build_local_group_maps([(g["gNodeOwner"], g["gNodeLocalId"]) for g, _ in ranks])
    fields = [("q0", 10), ("qgrad", 27), ("limiter", 5), ("lsq_sw", 6)]
    if "x" in ranks[0][0] and ranks[0][0]["x"].size == (int(ranks[0][1]["nnode"]) + int(ranks[0][1]["gnode"])) * 5:
        fields.append(("x", 5))
    for fname, w in fields:
        packed = [pobjs[r].pack_numpy(ranks[r][0][fname], w) for r in range(n)]
        for r in range(n):
            g, meta = ranks[r]
            nnode, gnode = int(meta["nnode"]), int(meta["gnode"])
            v = g[fname].copy().reshape(-1, w)
            v[nnode:nnode + gnode] = np.nan          # wipe the ghost rows, then refill them through the maps
            pobjs[r].unpack_numpy(v, w, nnode, [packed[p][r] for p in range(n)])
            exact(v.reshape(-1), g[fname], f"{fname} rank {r}")


def _gloo_worker(rank, world, name, port, q):
    try:
        import torch.distributed as dist
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from proteuscfd_b200.parallel import PObj, TorchGroup
        import torch
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        g, meta = load_golden(f"{name}_r{rank}of{world}")
        p = PObj(rank, world).BuildCommMaps(g["gNodeOwner"], g["gNodeLocalId"], TorchGroup(dist))
        ok = (np.array_equal(p.commCountsSend, g["commCountsSend"]) and np.array_equal(p.nodePackingList, g["nodePackingList"])
              and np.array_equal(p.commCountsRecv, g["commCountsRecv"]))
        # a real exchange of q0 over gloo with the persistent lists
        nnode, gnode = int(meta["nnode"]), int(meta["gnode"])
        v = g["q0"].copy().reshape(-1, 10)
        v[nnode:nnode + gnode] = np.nan
        send = p.pack_numpy(v, 10)
        reqs, recvs = [], {}
        for peer in range(world):
            if peer == rank:
                continue
            if p.commCountsSend[peer]:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send[peer])), peer))
            if p.commCountsRecv[peer]:
                recvs[peer] = torch.empty((int(p.commCountsRecv[peer]), 10), dtype=torch.float64)
                reqs.append(dist.irecv(recvs[peer], peer))
        for r in reqs:
            r.wait()
        p.unpack_numpy(v, 10, nnode, [recvs[k].numpy() if k in recvs else None for k in range(world)])
        ok = ok and np.array_equal(v.reshape(-1), g["q0"])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok)))
    except Exception as e:   # pragma: no cover
        q.put((rank, f"error: {e!r}"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_gloo_multiprocess_maps_and_exchange(name):
    import torch.multiprocessing as mp
    world = CASES[name]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (0 if world == 2 else 301)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, name, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)], res

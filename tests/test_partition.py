"""proteuscfd_b200/partition.py against the partitions the REFERENCE's udecomp wrote for the same mesh and partition
vector (2- and 3-rank fixtures, tests/golden/box8_2rank_explicit_r*.npz, box9_3rank_implicit_r*.npz): local numbering,
ghost order, gNodeOwner / gNodeLocalId are bit-exact; edge / half-edge sets coincide, metrics to 1e-12 (same dual
faces, summed in a different order).  Plus the invariants of an arbitrary (recursive-bisection) partition."""
import numpy as np
import pytest

from proteuscfd_b200.boxmesh import kuhn_box
from proteuscfd_b200.parallel import build_local_group_maps
from proteuscfd_b200.partition import rcb_partition, udecomp_partition
from tests.oracle_lib import load_golden


def slab(xyz, ranks, axis):
    return (np.clip(xyz[:, axis], 0.0, 1.0 - 1e-12) * ranks).astype(np.int64)


CASES = {"box8_2rank_explicit": (8, 2, 2), "box9_3rank_implicit": (9, 3, 0)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_partition_layout_matches_udecomp(name):
    n, nr, axis = CASES[name]
    xyz, tets, tris, tags = kuhn_box(n, jitter=0.15)
    parts = udecomp_partition(xyz, tets, tris, tags, slab(xyz, nr, axis), nr)
    for r, m in enumerate(parts):
        g, meta = load_golden(f"{name}_r{r}of{nr}")
        nn, gn = int(meta["nnode"]), int(meta["gnode"])
        assert (m["nnode"], m["gnode"], m["nedge"], m["nbedge"], m["ngedge"]) == \
               (nn, gn, int(meta["nedge"]), int(meta["nbedge"]), int(meta["ngedge"]))
        # numbering: owned nodes and ghosts in udecomp's order, ghost tables bit-exact
        assert np.array_equal(m["xyz"].reshape(-1, 3), g["xyz"].reshape(-1, 3)[: nn + gn])
        assert np.array_equal(m["gNodeOwner"], g["gNodeOwner"])
        assert np.array_equal(m["gNodeLocalId"], g["gNodeLocalId"])
        assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)
        # interior edges as a set, with their metrics
        ours = {(int(a), int(b)): v for (a, b), v in zip(m["edges_n"].reshape(-1, 2), m["edges_a"].reshape(-1, 4))}
        for (a, b), v in zip(g["edges_n"].reshape(-1, 2), g["edges_a"].reshape(-1, 4)):
            w = ours.pop((int(a), int(b)))
            assert abs(w[3] - v[3]) <= 1e-12 * v[3] and np.abs(w[:3] - v[:3]).max() <= 1e-11
        assert not ours
        # ghost half-edges (behind the nbedge boundary half-edges): (owned, ghost) pairs with the whole dual face
        nb = m["nbedge"]
        ours = {(int(a), int(b)): v for (a, b), v in zip(m["bedges_n"].reshape(-1, 2)[nb:], m["bedges_a"].reshape(-1, 4)[nb:])}
        for (a, b), v in zip(g["bedges_n"].reshape(-1, 2)[nb:], g["bedges_a"].reshape(-1, 4)[nb:]):
            w = ours.pop((int(a), int(b)))
            assert abs(w[3] - v[3]) <= 1e-12 * v[3] and np.abs(w[:3] - v[:3]).max() <= 1e-11
        assert not ours
        # boundary half-edges: same left nodes with the same multiplicity
        assert sorted(m["bedges_n"].reshape(-1, 2)[:nb, 0]) == sorted(g["bedges_n"].reshape(-1, 2)[:nb, 0])
    # the halo maps PObj derives from these tables are the reference's
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m in parts])
    for r, p in enumerate(pobjs):
        g, _ = load_golden(f"{name}_r{r}of{nr}")
        assert np.array_equal(p.commCountsSend, g["commCountsSend"])
        assert np.array_equal(p.commCountsRecv, g["commCountsRecv"])
        assert np.array_equal(p.nodePackingList, g["nodePackingList"])


@pytest.mark.parametrize("nr", [2, 3, 5, 8])
def test_rcb_partition_invariants(nr):
    xyz, tets, tris, tags = kuhn_box(5, jitter=0.15)
    part = rcb_partition(xyz, nr)
    counts = np.bincount(part, minlength=nr)
    assert counts.min() > 0 and counts.max() - counts.min() <= int(np.ceil(np.log2(nr))) + 1
    parts = udecomp_partition(xyz, tets, tris, tags, part, nr)
    assert sum(m["nnode"] for m in parts) == len(xyz)
    assert np.isclose(sum(m["vol"].sum() for m in parts), 1.0, rtol=1e-13)
    # ghosts receive their owners' coordinates through the maps
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m in parts])
    arrs = []
    for m in parts:
        a = m["xyz"].copy()
        a[3 * m["nnode"]:] = -777.0
        arrs.append(a)
    packed = [pobjs[r].pack_numpy(arrs[r], 3) for r in range(nr)]
    for r, m in enumerate(parts):
        pobjs[r].unpack_numpy(arrs[r], 3, m["nnode"], [packed[p][r] for p in range(nr)])
        assert np.array_equal(arrs[r], m["xyz"])
    # every cut edge is a ghost half-edge on both sides with opposite normals; interior + cut = all edges
    cut = {}
    for m in parts:
        nb = m["nbedge"]
        for (l, gno), a in zip(m["bedges_n"].reshape(-1, 2)[nb:], m["bedges_a"].reshape(-1, 4)[nb:]):
            cut[(int(m["gid"][l]), int(m["gid"][gno]))] = a
    for (a, b), v in cut.items():
        w = cut[(b, a)]
        assert np.array_equal(v[:3], -w[:3]) and v[3] == w[3]
    from proteuscfd_b200.dualmesh import median_dual
    assert sum(m["nedge"] for m in parts) + len(cut) // 2 == median_dual(xyz, tets, tris, tags)["nedge"]

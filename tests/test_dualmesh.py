"""Median-dual metrics (proteuscfd_b200/dualmesh.py: the set-up plumbing that feeds the synthetic bench / parity boxes)
against the metrics the REFERENCE computed for the same tetrahedra (Mesh::BuildMaps + CalcMetrics,
ucs/mesh.tcc:517-713, 1056-1167, 1653-2333), as dumped into the golden fixtures by the reference harness.

The reference orders edges and half-edges by its own neighbour lists (first-occurrence order), the
generator by ascending node id, so the comparison is by edge key; areas / volumes are sums of the same face pieces in a
different order, hence a 1e-12 bar, not bit equality.  (The hot path itself takes `ipsp/psp` and the edge arrays from
the host, never rebuilds them: pcfd_mesh_desc.)"""
import numpy as np
import pytest

from proteuscfd_b200.boxmesh import kuhn_box
from proteuscfd_b200.dualmesh import closure_defect, median_dual
from tests.oracle_lib import load_golden

CASES = {"box8_explicit_venkat": dict(n=8, jitter=0.15), "box6_implicit_sgs": dict(n=6, jitter=0.15),
         "ramp15_implicit": dict(n=8, jitter=0.1, ramp_deg=15.0)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_median_dual_matches_reference_metrics(name):
    g, meta = load_golden(name)
    xyz, tets, tris, tags = kuhn_box(**CASES[name])
    m = median_dual(xyz, tets, tris, tags)
    nn = int(meta["nnode"])
    assert m["nnode"] == nn and m["nedge"] == int(meta["nedge"]) and m["nbedge"] == int(meta["nbedge"])
    # same coordinates (the reference read them from the .ugrid this generator wrote)
    assert np.allclose(m["xyz"].reshape(-1, 3), g["xyz"].reshape(-1, 3)[:nn], rtol=0, atol=1e-15)
    # dual volumes
    assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)
    # interior edges by key: unit normal (n0 -> n1, n0 < n1) and dual face area
    en, ea = m["edges_n"].reshape(-1, 2), m["edges_a"].reshape(-1, 4)
    rn, ra = g["edges_n"].reshape(-1, 2), g["edges_a"].reshape(-1, 4)
    assert (en[:, 0] < en[:, 1]).all() and (rn[:, 0] < rn[:, 1]).all()
    ours = {(int(a), int(b)): v for (a, b), v in zip(en, ea)}
    assert len(ours) == len(en)
    for (a, b), v in zip(rn, ra):
        w = ours[(int(a), int(b))]
        assert abs(w[3] - v[3]) <= 1e-12 * v[3]
        assert np.abs(w[:3] - v[:3]).max() <= 1e-11
    # the reference's neighbour lists hold the same neighbours (its order is first-occurrence, ours ascending)
    ip, ps = g["ipsp"], g["psp"]
    for n in range(nn):
        assert sorted(ps[ip[n]:ip[n + 1]]) == list(m["psp"][m["ipsp"][n]:m["ipsp"][n + 1]])
    # boundary half-edges: one per (surface triangle, node); per (node, surface tag) the area vectors add up the same
    def by_node_tag(bn, ba, tag):
        acc = {}
        for (l, _), a, t in zip(bn.reshape(-1, 2), ba.reshape(-1, 4), tag):
            acc.setdefault((int(l), int(t)), np.zeros(3))
            acc[(int(l), int(t))] += a[:3] * a[3]
        return acc
    nb = int(meta["nbedge"])
    A = by_node_tag(m["bedges_n"], m["bedges_a"], m["bedges_factag"])
    B = by_node_tag(g["bedges_n"][: 2 * nb], g["bedges_a"][: 4 * nb], g["bedges_factag"][:nb])
    assert A.keys() == B.keys()
    for k in A:
        assert np.abs(A[k] - B[k]).max() <= 1e-13
    # and every dual cell is closed
    assert closure_defect(m) < 1e-15

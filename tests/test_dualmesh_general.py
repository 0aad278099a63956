"""Median-dual metrics for general elements (proteuscfd_b200/dualmesh.py: median_dual_general -- tets, pyramids, prisms,
hexes, triangular and quadrilateral boundary faces) against the metrics the REFERENCE computed (Mesh::CalcAreasVolumes,
ucs/mesh.tcc:1653-2218) for boxes of hexes, of prisms, of pyramids, for a box with all four volume element types
(tests/golden/elem_*.npz: boxmesh.mixed_box written as .ugrid, read and partitioned by the reference) and for the
reference's own unit-test mesh (cube_LowFi: unitTest/meshResources, prisms with triangular and quadrilateral faces through its HDF5 reader).  The fixtures carry
the reference's element list in its own winding (elem_type / elem_nodes / elem_factag).  Set-up plumbing (SURVEY.md 8f row
2), off the hot path; sums of the same face pieces in another order, hence 1e-12, not bit equality."""
import numpy as np
import pytest

from proteuscfd_b200.boxmesh import mixed_box
from proteuscfd_b200.dualmesh import closure_defect, median_dual_general, ugrid_to_reference_winding
from tests.oracle_lib import load_golden

CASES = ["elem_hex", "elem_prism", "elem_pyramid", "elem_mixed", "cube_LowFi"]


@pytest.mark.parametrize("name", CASES)
def test_general_median_dual_matches_reference_metrics(name):
    g, meta = load_golden(name)
    nn, nb = int(meta["nnode"]), int(meta["nbedge"])
    m = median_dual_general(g["xyz"].reshape(-1, 3)[:nn], g["elem_type"], g["elem_nodes"], g["elem_factag"])
    assert m["nnode"] == nn and m["nedge"] == int(meta["nedge"]) and m["nbedge"] == nb
    assert (g["vol"] > 0).all()
    assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)
    en, ea = m["edges_n"].reshape(-1, 2), m["edges_a"].reshape(-1, 4)
    rn, ra = g["edges_n"].reshape(-1, 2), g["edges_a"].reshape(-1, 4)
    assert (rn[:, 0] < rn[:, 1]).all()
    ours = {(int(a), int(b)): v for (a, b), v in zip(en, ea)}
    assert len(ours) == len(en) == len(rn)
    for (a, b), v in zip(rn, ra):
        w = ours[(int(a), int(b))]
        assert abs(w[3] - v[3]) <= 1e-12 * v[3]
        assert np.abs(w[:3] - v[:3]).max() <= 1e-11
    ip, ps = g["ipsp"], g["psp"]
    for n in range(nn):
        assert sorted(ps[ip[n]:ip[n + 1]]) == list(m["psp"][m["ipsp"][n]:m["ipsp"][n + 1]])

    # boundary half-edges: one per (boundary face, node); per (node, surface tag) the area vectors add up the same
    def by_node_tag(bn, ba, tag):
        acc = {}
        for (l, _), a, t in zip(bn.reshape(-1, 2), ba.reshape(-1, 4), tag):
            acc.setdefault((int(l), int(t)), np.zeros(3))
            acc[(int(l), int(t))] += a[:3] * a[3]
        return acc
    A = by_node_tag(m["bedges_n"], m["bedges_a"], m["bedges_factag"])
    B = by_node_tag(g["bedges_n"][: 2 * nb], g["bedges_a"][: 4 * nb], g["bedges_factag"][:nb])
    assert A.keys() == B.keys()
    for k in A:
        assert np.abs(A[k] - B[k]).max() <= 1e-13
    assert closure_defect(m) < 1e-15


@pytest.mark.parametrize("kind", ["hex", "prism", "pyramid", "mixed"])
def test_generator_winding_is_the_references(kind):
    """what the reference's UGRID reader made of the generator's file is what ugrid_to_reference_winding says: same
    elements, same winding (the reference keeps file order within a type)"""
    g, meta = load_golden(f"elem_{kind}")
    xyz, el, tris, tt, quads, qt = mixed_box(4, kind, jitter=0.12)
    et, en, ef = ugrid_to_reference_winding(el, tris, tt, quads, qt)
    ref = sorted(zip(g["elem_type"].tolist(), map(tuple, g["elem_nodes"].reshape(-1, 8).tolist()), g["elem_factag"].tolist()))
    ours = sorted(zip(et.tolist(), map(tuple, en.tolist()), np.where(et <= 1, ef, 0).tolist()))
    ref = [(t, n, f if t <= 1 else 0) for t, n, f in ref]
    assert ours == ref
    assert np.allclose(xyz, g["xyz"].reshape(-1, 3)[: len(xyz)], rtol=0, atol=1e-15)


MAPS = ["elem_hex", "elem_prism", "elem_pyramid", "elem_mixed", "rcm_2rank_r0of2", "rcm_2rank_r1of2"]


@pytest.mark.parametrize("name", MAPS)
def test_connectivity_maps_in_the_references_order(name):
    """build_maps == Mesh::BuildPsp / BuildEdges on the reference's element list: neighbour lists, interior edges, boundary
    and ghost half-edges, phantom nodes and surface tags, entry for entry -- also on both ranks of a partition (ghost nodes,
    surface elements with ghost corners).  (cube_LowFi is left out: a pre-partitioned HDF5 mesh brings its own maps.)"""
    from proteuscfd_b200.dualmesh import build_maps
    g, meta = load_golden(name)
    nn, gn, nb = int(meta["nnode"]), int(meta["gnode"]), int(meta["nbedge"])
    m = build_maps(nn, gn, g["elem_type"], g["elem_nodes"], g["elem_factag"])
    assert (m["nedge"], m["nbedge"], m["ngedge"]) == (int(meta["nedge"]), nb, int(meta["ngedge"]))
    assert np.array_equal(m["ipsp"][: nn + 1], g["ipsp"])
    assert np.array_equal(m["psp"][: g["psp"].size], g["psp"])
    assert np.array_equal(m["edges_n"], g["edges_n"])
    assert np.array_equal(m["bedges_n"], g["bedges_n"])
    assert np.array_equal(m["bedges_factag"][:nb], g["bedges_factag"][:nb])


@pytest.mark.parametrize("name", ["elem_mixed", "elem_pyramid", "elem_hex"])
def test_reference_ordered_mesh_description(name):
    """median_dual_general(reference_order=True) is the reference's mesh description position by position: index arrays
    equal, metrics at 1e-12 -- a pcfd_mesh_desc built here from an element list feeds the hot path the reference's sums in
    the reference's order"""
    g, meta = load_golden(name)
    nn, nb = int(meta["nnode"]), int(meta["nbedge"])
    m = median_dual_general(g["xyz"].reshape(-1, 3)[:nn], g["elem_type"], g["elem_nodes"], g["elem_factag"], reference_order=True)
    for k in ("edges_n", "bedges_n", "ipsp", "psp"):
        assert np.array_equal(m[k], g[k][: m[k].size]), k
    assert np.array_equal(m["bedges_factag"], g["bedges_factag"][:nb])
    ea, ra = m["edges_a"].reshape(-1, 4), g["edges_a"].reshape(-1, 4)
    assert np.abs(ea - ra).max() <= 1e-12
    ba, rb = m["bedges_a"].reshape(-1, 4), g["bedges_a"].reshape(-1, 4)[:nb]
    assert np.abs(ba - rb).max() <= 1e-12
    assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)


def test_hot_path_on_a_mesh_built_here_reproduces_the_reference(oracle):
    """element list -> build_maps + median_dual_general -> the C oracle's gradient / limiter / residual, against the
    REFERENCE's own arrays for that mesh: same order of every sum, metrics within 1e-15, hence results within 1e-11 of
    their scale (not bit equality: the dual-face pieces are added in another order)"""
    from tests.oracle_lib import Oracle
    g, meta = load_golden("elem_mixed")
    nn, nb = int(meta["nnode"]), int(meta["nbedge"])
    m = median_dual_general(g["xyz"].reshape(-1, 3)[:nn], g["elem_type"], g["elem_nodes"], g["elem_factag"], reference_order=True)
    g2 = dict(g)
    for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "vol", "ipsp", "psp"):
        g2[k] = m[k]
    tag2type = {int(t): int(b) for t, b in zip(g["bedges_factag"][:nb], g["bedges_bctype"][:nb])}
    g2["bedges_bctype"] = np.array([tag2type[int(t)] for t in m["bedges_factag"]], dtype=np.int32)
    assert np.array_equal(g2["bedges_bctype"], g["bedges_bctype"][:nb])
    o = Oracle(oracle, g2, meta)
    _, sw = o.lsq()
    qgrad = o.gradient(g["q0"].copy(), sw)
    assert np.abs(qgrad - g["qgrad"]).max() <= 1e-11 * np.abs(g["qgrad"]).max()
    assert np.abs(sw - g["lsq_sw"]).max() <= 1e-11 * np.abs(g["lsq_sw"]).max()


@pytest.mark.parametrize("kind", ["mixed", "pyramid"])
def test_mesh_file_to_mesh_description(tmp_path, kind):
    """.ugrid -> mesh_from_ugrid: from the FILE alone the reference's mesh description, position by position (the same
    file went through the reference's reader and decomposer for the fixture)"""
    from proteuscfd_b200.boxmesh import read_ugrid, write_ugrid_general
    from proteuscfd_b200.dualmesh import mesh_from_ugrid
    g, meta = load_golden(f"elem_{kind}")
    nn, nb = int(meta["nnode"]), int(meta["nbedge"])
    path = str(tmp_path / "m.ugrid")
    gen = mixed_box(4, kind, jitter=0.12)
    write_ugrid_general(path, *gen)
    back = read_ugrid(path)
    assert np.array_equal(back[0], gen[0])
    for k in gen[1]:
        assert np.array_equal(back[1][k], gen[1][k]), k
    for a, b in zip(back[2:], gen[2:]):
        assert np.array_equal(a, b)
    m = mesh_from_ugrid(path)
    assert np.array_equal(m["elem_type"], g["elem_type"]) and np.array_equal(m["elem_nodes"].reshape(-1), g["elem_nodes"])
    for k in ("edges_n", "bedges_n", "ipsp", "psp"):
        assert np.array_equal(m[k], g[k][: m[k].size]), k
    assert np.array_equal(m["bedges_factag"], g["bedges_factag"][:nb])
    assert np.abs(m["edges_a"] - g["edges_a"]).max() <= 1e-12
    assert np.abs(m["bedges_a"] - g["bedges_a"][: 4 * nb]).max() <= 1e-12
    assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)
    assert np.array_equal(m["xyz"], g["xyz"][: 3 * nn])


def test_mesh_file_to_reordered_mesh_description(tmp_path):
    """the solver's DEFAULT start-up (reorderMesh = 1): .ugrid -> reverse Cuthill-McKee -> ReorderC2nMap -> maps and
    metrics, against the reference run with reordering on (elem_mixed_rcm): the renumbered element list, coordinates and
    every index array equal, metrics at 1e-12"""
    from proteuscfd_b200.boxmesh import write_ugrid_general
    from proteuscfd_b200.dualmesh import mesh_from_ugrid
    g, meta = load_golden("elem_mixed_rcm")
    plain, _ = load_golden("elem_mixed")
    assert not np.array_equal(g["elem_nodes"], plain["elem_nodes"])          # the reference did renumber
    nn, nb = int(meta["nnode"]), int(meta["nbedge"])
    path = str(tmp_path / "m.ugrid")
    write_ugrid_general(path, *mixed_box(4, "mixed", jitter=0.12))
    m = mesh_from_ugrid(path, reorder=True)
    assert np.array_equal(m["elem_nodes"].reshape(-1), g["elem_nodes"])
    assert np.array_equal(m["xyz"], g["xyz"][: 3 * nn])
    for k in ("edges_n", "bedges_n", "ipsp", "psp"):
        assert np.array_equal(m[k], g[k][: m[k].size]), k
    assert np.array_equal(m["bedges_factag"], g["bedges_factag"][:nb])
    assert np.abs(m["edges_a"] - g["edges_a"]).max() <= 1e-12
    assert np.abs(m["bedges_a"] - g["bedges_a"][: 4 * nb]).max() <= 1e-12
    assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)


def test_partitioned_mesh_description_from_the_global_element_list():
    """global element list + node partition -> per rank what udecomp writes and the solver builds from it
    (partition.udecomp_elements: local then split elements, owned nodes then ghosts grouped by owner; build_maps; metrics
    with the cut edges as ghost half-edges), against both ranks of a reference run (rcm_2rank_r*of2): element lists, ghost
    tables, coordinates and every index array entry for entry, metrics at 1e-12"""
    import sys
    from proteuscfd_b200.boxmesh import kuhn_box
    from proteuscfd_b200.partition import udecomp_elements
    xyz, tets, tris, tags = kuhn_box(6, jitter=0.15)
    part = (np.clip(xyz[:, 2], 0.0, 1.0 - 1e-12) * 2).astype(np.int64)          # tools/make_golden.py: slab_part
    none = {"pyramid": np.zeros((0, 5), int), "prism": np.zeros((0, 6), int), "hex": np.zeros((0, 8), int)}
    et, en, ef = ugrid_to_reference_winding(dict(none, tet=tets), tris, tags, np.zeros((0, 4), int), np.zeros(0, int))
    ranks = udecomp_elements(et, en, ef, part, 2)
    for r in (0, 1):
        g, meta = load_golden(f"rcm_2rank_r{r}of2")
        R = ranks[r]
        assert np.array_equal(R["elem_type"], g["elem_type"]) and np.array_equal(R["elem_nodes"].reshape(-1), g["elem_nodes"])
        assert np.array_equal(R["gNodeOwner"], g["gNodeOwner"]) and np.array_equal(R["gNodeLocalId"], g["gNodeLocalId"])
        gid = np.concatenate([R["owned"], R["ghosts"]])
        m = median_dual_general(xyz[gid], R["elem_type"], R["elem_nodes"], R["elem_factag"], reference_order=True,
                                nnode=R["owned"].size)
        assert (m["nnode"], m["gnode"], m["nedge"], m["nbedge"], m["ngedge"]) == tuple(
            int(meta[k]) for k in ("nnode", "gnode", "nedge", "nbedge", "ngedge"))
        for k in ("edges_n", "bedges_n", "ipsp", "psp", "xyz"):
            assert np.array_equal(m[k], g[k]), (r, k)
        nb = m["nbedge"]
        assert np.array_equal(m["bedges_factag"][:nb], g["bedges_factag"][:nb])
        assert np.abs(m["edges_a"] - g["edges_a"]).max() <= 1e-12
        assert np.abs(m["bedges_a"] - g["bedges_a"]).max() <= 1e-12
        assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)


def test_partitioned_general_element_mesh_from_the_mesh_file(tmp_path):
    """the same for the all-types box cut into two y-slabs through prisms, hexes, pyramids and tets, starting from the
    .ugrid FILE (elem_mixed_2rank_r*of2)"""
    from proteuscfd_b200.boxmesh import read_ugrid, write_ugrid_general
    from proteuscfd_b200.partition import udecomp_elements
    path = str(tmp_path / "m.ugrid")
    write_ugrid_general(path, *mixed_box(4, "mixed", jitter=0.12))
    xyz, el, tris, tt, quads, qt = read_ugrid(path)
    et, en, ef = ugrid_to_reference_winding(el, tris, tt, quads, qt)
    part = (np.clip(xyz[:, 1], 0.0, 1.0 - 1e-12) * 2).astype(np.int64)
    ranks = udecomp_elements(et, en, ef, part, 2)
    for r in (0, 1):
        g, meta = load_golden(f"elem_mixed_2rank_r{r}of2")
        R = ranks[r]
        assert np.array_equal(R["elem_type"], g["elem_type"]) and np.array_equal(R["elem_nodes"].reshape(-1), g["elem_nodes"])
        assert np.array_equal(R["gNodeOwner"], g["gNodeOwner"]) and np.array_equal(R["gNodeLocalId"], g["gNodeLocalId"])
        gid = np.concatenate([R["owned"], R["ghosts"]])
        m = median_dual_general(xyz[gid], R["elem_type"], R["elem_nodes"], R["elem_factag"], reference_order=True,
                                nnode=R["owned"].size)
        for k in ("edges_n", "bedges_n", "ipsp", "psp", "xyz"):
            assert np.array_equal(m[k], g[k]), (r, k)
        assert np.abs(m["edges_a"] - g["edges_a"]).max() <= 1e-12
        assert np.abs(m["bedges_a"] - g["bedges_a"]).max() <= 1e-12
        assert np.allclose(m["vol"], g["vol"], rtol=1e-12, atol=0)

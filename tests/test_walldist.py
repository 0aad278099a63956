"""proteuscfd_b200/walldist.py against the `wallDistance` field the reference computed (ComputeWallDistOct,
ucs/walldist.tcc:116-199) for its Spalart-Allmaras fixture: bit-exact on every node."""
import numpy as np

from proteuscfd_b200.walldist import nearest_distance, wall_distance, wall_points
from tests.oracle_lib import load_golden


def sa_mesh():
    g, meta = load_golden("box6_sa_implicit")
    mesh = {k: g[k] for k in ("bedges_n", "bedges_bctype", "xyz")}
    for k in ("nnode", "gnode", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    return mesh, g


def test_wall_distance_matches_reference_field():
    mesh, g = sa_mesh()
    pts = wall_points(mesh)
    assert len(pts) > 0 and np.all(pts[:, 2] == pts[0, 2])          # the no-slip floor of ns_bc (tools/make_golden.py)
    d = wall_distance(mesh)
    ref = g["wallDistance"][: mesh["nnode"] + mesh["gnode"]]
    assert np.array_equal(d, ref)
    assert (d[:mesh["nnode"]] == 0.0).sum() == len(np.unique(pts, axis=0))   # exactly the wall nodes sit at distance zero


def test_chunking_does_not_change_the_result():
    mesh, _ = sa_mesh()
    xyz = mesh["xyz"].reshape(-1, 3)[: mesh["nnode"]]
    pts = wall_points(mesh)
    a = nearest_distance(xyz, pts)
    b = nearest_distance(xyz, pts, chunk=997)
    assert np.array_equal(a, b)
    assert np.isinf(nearest_distance(xyz, np.zeros((0, 3)))).all()            # no viscous wall anywhere


def test_partitioned_wall_distance_equals_the_serial_one():
    # wall nodes of every rank are gathered (SyncParallelPoint, walldist.tcc:24-113): the field of a partition equals the
    # serial field on its nodes
    from proteuscfd_b200.boxmesh import kuhn_box
    from proteuscfd_b200.dualmesh import median_dual
    from proteuscfd_b200.partition import rcb_partition, udecomp_partition
    xyz, tets, tris, tags = kuhn_box(5, jitter=0.15)
    lut = np.array([0, 6, 6, 9, 9, 4, 6], dtype=np.int32)      # tag 5 (zmin) = no-slip floor
    full = median_dual(xyz, tets, tris, tags)
    serial = dict(full, bedges_bctype=lut[full["bedges_factag"]], gnode=0, ngedge=0)
    dser = wall_distance(serial)
    nr = 3
    parts = udecomp_partition(xyz, tets, tris, tags, rcb_partition(xyz, nr), nr, bc_lut=lut, full=full)

    class Group:      # every rank's allgather returns all ranks' wall points
        def __init__(self, items):
            self.items = items

        def allgather(self, _):
            return self.items
    grp = Group([wall_points(m) for m in parts])
    for m in parts:
        d = wall_distance(m, group=grp)
        assert np.array_equal(d, dser[m["gid"]])

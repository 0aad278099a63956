"""Median-dual metrics of a tetrahedral mesh, as flat arrays in the reference's layout.

This is set-up plumbing for the synthetic benchmark / parity meshes:
it produces what the reference's `Mesh::BuildMaps` + `CalcMetrics`
(ucs/mesh.tcc:517-713, 1056-1167, 1653-2333) hand to the hot path -- `edges`,
`bedges` (one half-edge per (surface triangle, node), each with its own phantom
node), `vol`, `ipsp/psp` -- for a mesh generated in memory.  It runs on whichever
torch device it is given (the B200 for the 10 M-cell box, the CPU in tests).

Edge `a[0:3]` is the unit normal pointing from n0 to n1 (n0 < n1), `a[3]` the dual
face area (mesh.tcc:2320-2325); half-edge normals point out of the domain.
Neighbour lists are sorted ascending, so edges are sorted by (n0, n1).
"""
import numpy as np
import torch

# local edge (i, j) and the remaining vertices (k, l) with (i, j, k, l) an even permutation
_LOCAL_EDGES = [(0, 1, 2, 3), (0, 2, 3, 1), (0, 3, 1, 2), (1, 2, 0, 3), (1, 3, 2, 0), (2, 3, 0, 1)]


def median_dual(xyz, tets, tris, tags, device="cpu"):
    """Return a dict of numpy arrays: edges_n, edges_a, bedges_n, bedges_a, bedges_factag, vol, ipsp, psp
    plus the counts nnode, nedge, nbedge, nbnode (gnode = ngedge = 0: one partition)."""
    dev = torch.device(device)
    X = torch.as_tensor(np.ascontiguousarray(xyz), dtype=torch.float64, device=dev)
    T = torch.as_tensor(np.ascontiguousarray(tets).astype(np.int64), device=dev)
    nn = X.shape[0]
    P = [X[T[:, v]] for v in range(4)]
    G = (P[0] + P[1] + P[2] + P[3]) / 4.0

    keys, normals = [], []
    for (i, j, k, l) in _LOCAL_EDGES:
        a, b = T[:, i], T[:, j]
        M = (P[i] + P[j]) / 2.0
        F1 = (P[i] + P[j] + P[k]) / 3.0
        F2 = (P[i] + P[j] + P[l]) / 3.0
        GM = G - M
        # two triangles (M, F1, G) and (M, G, F2) of the dual face; for a positively
        # oriented tet this sum points from vertex i to vertex j
        n = 0.5 * (torch.linalg.cross(F1 - M, GM) + torch.linalg.cross(GM, F2 - M))
        swap = a > b
        lo = torch.where(swap, b, a)
        hi = torch.where(swap, a, b)
        n = torch.where(swap[:, None], -n, n)
        keys.append(lo * nn + hi)
        normals.append(n)
    keys = torch.cat(keys)
    normals = torch.cat(normals)
    # deterministic accumulation: sort contributions by edge key (stable) and segment-sum
    order = torch.argsort(keys, stable=True)
    keys = keys[order]
    normals = normals[order]
    uniq, counts = torch.unique_consecutive(keys, return_counts=True)
    ne = uniq.shape[0]
    seg = torch.repeat_interleave(torch.arange(ne, device=dev), counts)
    avec = torch.zeros((ne, 3), dtype=torch.float64, device=dev)
    avec.index_add_(0, seg, normals)
    del normals, keys, order, seg
    n0 = uniq // nn
    n1 = uniq % nn
    area = torch.linalg.norm(avec, dim=1)
    edges_a = torch.cat([avec / area[:, None], area[:, None]], dim=1)

    # dual volumes: a quarter of each tet to each of its nodes
    tv = (torch.linalg.cross(P[1] - P[0], P[2] - P[0]) * (P[3] - P[0])).sum(dim=1) / 6.0
    vol = torch.zeros(nn, dtype=torch.float64, device=dev)
    for v in range(4):
        vol.index_add_(0, T[:, v], tv / 4.0)

    # boundary half-edges: a third of each surface triangle to each of its nodes, outward
    F = torch.as_tensor(np.ascontiguousarray(tris).astype(np.int64), device=dev)
    nf = F.shape[0]
    p0, p1, p2 = X[F[:, 0]], X[F[:, 1]], X[F[:, 2]]
    fn = -0.5 * torch.linalg.cross(p1 - p0, p2 - p0) / 3.0     # triangles are wound with the normal inward
    fa = torch.linalg.norm(fn, dim=1)
    ba = torch.cat([fn / fa[:, None], fa[:, None]], dim=1)      # [nf,4]
    bedges_a = ba[:, None, :].expand(nf, 3, 4).reshape(nf * 3, 4)
    left = F.reshape(-1)
    right = nn + torch.arange(nf * 3, device=dev)
    bedges_n = torch.stack([left, right], dim=1)
    factag = torch.as_tensor(np.ascontiguousarray(tags).astype(np.int64), device=dev)[:, None].expand(nf, 3).reshape(-1)

    # point-surrounding-point lists, ascending
    a = torch.cat([n0, n1])
    b = torch.cat([n1, n0])
    o = torch.argsort(a * nn + b)
    psp = b[o]
    deg = torch.bincount(a, minlength=nn)
    ipsp = torch.zeros(nn + 1, dtype=torch.int64, device=dev)
    ipsp[1:] = torch.cumsum(deg, 0)

    def npi(t):
        return t.to(torch.int32).cpu().numpy()

    return dict(
        nnode=nn, gnode=0, nbnode=nf * 3, nedge=ne, nbedge=nf * 3, ngedge=0,
        edges_n=np.ascontiguousarray(npi(torch.stack([n0, n1], dim=1)).reshape(-1)),
        edges_a=np.ascontiguousarray(edges_a.cpu().numpy().reshape(-1)),
        bedges_n=np.ascontiguousarray(npi(bedges_n).reshape(-1)),
        bedges_a=np.ascontiguousarray(bedges_a.cpu().numpy().reshape(-1)),
        bedges_factag=npi(factag),
        xyz=np.ascontiguousarray(X.cpu().numpy().reshape(-1)),
        vol=vol.cpu().numpy(), ipsp=npi(ipsp), psp=npi(psp))


def closure_defect(mesh):
    """max over nodes of |sum of outward dual-face area vectors| (zero for a closed dual)."""
    nn = mesh["nnode"]
    en = mesh["edges_n"].reshape(-1, 2)
    ea = mesh["edges_a"].reshape(-1, 4)
    bn = mesh["bedges_n"].reshape(-1, 2)
    ba = mesh["bedges_a"].reshape(-1, 4)
    s = np.zeros((nn, 3))
    v = ea[:, :3] * ea[:, 3:4]
    np.add.at(s, en[:, 0], v)
    np.add.at(s, en[:, 1], -v)
    np.add.at(s, bn[:, 0], ba[:, :3] * ba[:, 3:4])
    return np.abs(s).max()


# ---- general elements (Mesh::CalcAreasVolumes, ucs/mesh.tcc:1653-2218).  The reference's element conventions as data:
# element type ids (etypes.h), local edges with their preferred orientation, the two faces of every edge (the first one to
# the left of the oriented edge seen from inside), and the nodes of every face, all in the reference's own winding.
TRI, QUAD, TET, PYRAMID, PRISM, HEX = range(6)
_NV = {TRI: 3, QUAD: 4, TET: 4, PYRAMID: 5, PRISM: 6, HEX: 8}
_EDGES = {
    TET: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)],
    PYRAMID: [(0, 1), (0, 3), (0, 4), (1, 2), (1, 4), (2, 3), (2, 4), (3, 4)],
    PRISM: [(0, 1), (0, 2), (0, 3), (1, 2), (1, 4), (2, 5), (3, 4), (3, 5), (4, 5)],
    HEX: [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)],
}
_EDGE_FACES = {
    TET: [(0, 1), (2, 0), (1, 2), (0, 3), (3, 1), (2, 3)],
    PYRAMID: [(0, 1), (2, 0), (1, 2), (0, 3), (3, 1), (0, 4), (4, 3), (2, 4)],
    PRISM: [(0, 1), (2, 0), (1, 2), (0, 3), (3, 1), (2, 3), (1, 4), (4, 2), (3, 4)],
    HEX: [(0, 1), (2, 0), (1, 2), (0, 3), (3, 1), (0, 4), (4, 3), (2, 4), (1, 5), (5, 2), (3, 5), (4, 5)],
}
_FACES = {
    TET: [(0, 1, 2), (0, 3, 1), (0, 2, 3), (1, 3, 2)],
    PYRAMID: [(0, 1, 2, 3), (0, 4, 1), (0, 3, 4), (1, 4, 2), (2, 4, 3)],
    PRISM: [(0, 1, 2), (0, 3, 4, 1), (0, 2, 5, 3), (1, 4, 5, 2), (3, 5, 4)],
    HEX: [(0, 1, 2, 3), (0, 4, 5, 1), (0, 3, 7, 4), (1, 5, 6, 2), (2, 6, 7, 3), (4, 7, 6, 5)],
}
# the two neighbours of every node inside a boundary face (mesh.tcc:2127-2131)
_FACE_NEIGHBOURS = {TRI: [(1, 2), (2, 0), (0, 1)], QUAD: [(1, 3), (2, 0), (3, 1), (0, 2)]}


def _tri_area(p1, p2, p3):
    return 0.5 * torch.linalg.cross(p2 - p1, p3 - p1)


def _tet_vol(p1, p2, p3, p4):
    return (torch.linalg.cross(p2 - p1, p3 - p1) * (p4 - p1)).sum(dim=1) / 6.0


def median_dual_general(xyz, elem_type, elem_nodes, elem_factag, device="cpu", reference_order=False, nnode=None):
    """Median-dual metrics of a mesh of tets, pyramids, prisms and hexes with triangular / quadrilateral boundary faces,
    elements in the reference's winding (elem_nodes [nelem, 8], -1 padded; elem_type as etypes.h).  Per volume element and
    local edge the dual face is two triangles (element centroid, face centroid, edge midpoint), signed by whether the
    mesh edge n0 < n1 runs along the element's preferred edge orientation, and four tets for the two nodes' volumes
    (mesh.tcc:1812-1945); per boundary face and node one half-edge of two triangles (:2142-2187).  Centroids are node
    averages (geometry.h:217-249).  Same result layout as median_dual; edges sorted by (n0, n1), half-edges by face --
    or, with reference_order, everything arranged as build_maps (the reference's BuildPsp / BuildEdges) orders it, which
    makes the dict a pcfd_mesh_desc whose results equal ucs.x's on the same element list.
    nnode (with reference_order): a partition -- the first nnode nodes are local, the rest ghosts; elements are the rank's
    local + split elements (partition.udecomp_elements).  Edges between two local nodes stay edges, (local, ghost) edges
    become the ghost half-edges behind the boundary ones (whole dual face: every element around a cut edge is on the rank,
    mesh.tcc:1948-2120), volumes and boundary half-edges are the local nodes' only."""
    if nnode is not None and not reference_order:
        raise ValueError("median_dual_general: a partition (nnode) is laid out in the reference's order only")
    dev = torch.device(device)
    X = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3), device=dev)
    et = np.asarray(elem_type).astype(np.int64)
    en = np.asarray(elem_nodes).astype(np.int64).reshape(-1, 8)
    ef = np.asarray(elem_factag).astype(np.int64)
    nn = X.shape[0]
    keys, avecs, vol = [], [], torch.zeros(nn, dtype=torch.float64, device=dev)
    for t in (TET, PYRAMID, PRISM, HEX):
        E = torch.as_tensor(en[et == t][:, :_NV[t]], device=dev)
        if E.shape[0] == 0:
            continue
        P = X[E]                                    # [ne, nv, 3]
        ctr = P.mean(dim=1)
        fc = [P[:, list(f)].mean(dim=1) for f in _FACES[t]]
        for (la, lb), (f1, f2) in zip(_EDGES[t], _EDGE_FACES[t]):
            a, b = E[:, la], E[:, lb]
            fwd = a < b                              # the mesh edge (n0 < n1) runs la -> lb
            o = torch.where(fwd, 1.0, -1.0).to(torch.float64)
            n0, n1 = torch.where(fwd, a, b), torch.where(fwd, b, a)
            X0, X1 = X[n0], X[n1]
            mid = (X0 + X1) / 2.0
            area = (_tri_area(ctr, fc[f2], mid) + _tri_area(mid, fc[f1], ctr)) * o[:, None]
            keys.append(n0 * nn + n1)
            avecs.append(area)
            vol.index_add_(0, n0, o * (_tet_vol(X0, ctr, fc[f2], mid) + _tet_vol(X0, fc[f1], ctr, mid)))
            vol.index_add_(0, n1, o * (_tet_vol(X1, ctr, fc[f1], mid) + _tet_vol(X1, fc[f2], ctr, mid)))
    keys = torch.cat(keys)
    avecs = torch.cat(avecs)
    order = torch.argsort(keys, stable=True)
    keys, avecs = keys[order], avecs[order]
    uniq, counts = torch.unique_consecutive(keys, return_counts=True)
    ne = uniq.shape[0]
    seg = torch.repeat_interleave(torch.arange(ne, device=dev), counts)
    avec = torch.zeros((ne, 3), dtype=torch.float64, device=dev)
    avec.index_add_(0, seg, avecs)
    e0, e1 = uniq // nn, uniq % nn
    area = torch.linalg.norm(avec, dim=1)
    edges_a = torch.cat([avec / area[:, None], area[:, None]], dim=1)

    bl, ba, bt, bk = [], [], [], []
    for t in (TRI, QUAD):
        sel = et == t
        F = torch.as_tensor(en[sel][:, :_NV[t]], device=dev)
        if F.shape[0] == 0:
            continue
        P = X[F]
        ctr = P.mean(dim=1)
        tag = torch.as_tensor(ef[sel], device=dev)
        eid = torch.as_tensor(np.nonzero(sel)[0], device=dev)
        for k, (a1, a2) in enumerate(_FACE_NEIGHBOURS[t]):
            p = P[:, k]
            v = _tri_area(ctr, p, (p + P[:, a1]) / 2.0) + _tri_area((p + P[:, a2]) / 2.0, p, ctr)
            bl.append(F[:, k])
            ba.append(v)
            bt.append(tag)
            bk.append(eid * 4 + k)
    left = torch.cat(bl)
    bvec = torch.cat(ba)
    btag = torch.cat(bt)
    if reference_order:      # surface elements in list order, their nodes in element order (BuildEdges, mesh.tcc:1103-1123)
        o = torch.argsort(torch.cat(bk))
        left, bvec, btag = left[o], bvec[o], btag[o]
        if nnode is not None:
            keep = left < nnode
            left, bvec, btag = left[keep], bvec[keep], btag[keep]
    barea = torch.linalg.norm(bvec, dim=1)
    bedges_a = torch.cat([bvec / barea[:, None], barea[:, None]], dim=1)
    nbe = left.shape[0]
    bedges_n = torch.stack([left, nn + torch.arange(nbe, device=dev)], dim=1)

    a = torch.cat([e0, e1])
    b = torch.cat([e1, e0])
    o = torch.argsort(a * nn + b)
    psp = b[o]
    ipsp = torch.zeros(nn + 1, dtype=torch.int64, device=dev)
    ipsp[1:] = torch.cumsum(torch.bincount(a, minlength=nn), 0)

    def npi(v):
        return v.to(torch.int32).cpu().numpy()

    out = dict(
        nnode=nn, gnode=0, nbnode=nbe, nedge=ne, nbedge=nbe, ngedge=0,
        edges_n=np.ascontiguousarray(npi(torch.stack([e0, e1], dim=1)).reshape(-1)),
        edges_a=np.ascontiguousarray(edges_a.cpu().numpy().reshape(-1)),
        bedges_n=np.ascontiguousarray(npi(bedges_n).reshape(-1)),
        bedges_a=np.ascontiguousarray(bedges_a.cpu().numpy().reshape(-1)),
        bedges_factag=npi(btag), xyz=np.ascontiguousarray(X.cpu().numpy().reshape(-1)),
        vol=vol.cpu().numpy(), ipsp=npi(ipsp), psp=npi(psp))
    if reference_order:
        nloc = nn if nnode is None else int(nnode)
        maps = build_maps(nloc, nn - nloc, et, en, ef)
        assert maps["nbedge"] == nbe and (nnode is not None or maps["nedge"] == ne)
        keys_sorted = uniq.cpu().numpy()                                               # ascending edge keys n0 * nn + n1
        all_a = out["edges_a"].reshape(-1, 4)

        def rows(pairs):
            pairs = pairs.reshape(-1, 2).astype(np.int64)
            pos = np.searchsorted(keys_sorted, pairs[:, 0] * nn + pairs[:, 1])
            assert np.array_equal(keys_sorted[pos], pairs[:, 0] * nn + pairs[:, 1])
            return all_a[pos]

        ghost_n = maps["bedges_n"].reshape(-1, 2)[nbe:]
        assert np.array_equal(out["bedges_n"], maps["bedges_n"][: 2 * nbe])
        assert np.array_equal(out["bedges_factag"], maps["bedges_factag"][:nbe])
        out.update(nnode=nloc, gnode=nn - nloc, nedge=maps["nedge"], ngedge=maps["ngedge"], edges_n=maps["edges_n"],
                   edges_a=np.ascontiguousarray(rows(maps["edges_n"]).reshape(-1)), bedges_n=maps["bedges_n"],
                   bedges_a=np.ascontiguousarray(np.concatenate([out["bedges_a"].reshape(-1, 4), rows(ghost_n)]).reshape(-1)),
                   bedges_factag=maps["bedges_factag"], vol=out["vol"][:nloc], ipsp=maps["ipsp"][: nloc + 1],
                   psp=maps["psp"][: int(maps["ipsp"][nloc])])
    return out


def ugrid_to_reference_winding(el, tris, tri_tags, quads, quad_tags):
    """element list (type, 8 node slots, factag) in the reference's winding from UGRID-wound arrays, in the order the
    reference's reader stores them (ReadUGRID_Ascii, mesh.tcc:6740-6935: boundary faces reversed, pyramids through
    {0,3,4,1,2} -- boxmesh.mixed_box already hands pyramids over as (base, apex) --, the rest unchanged)"""
    types, nodes, tags = [], [], []

    def put(t, arr, tg=None):
        arr = np.asarray(arr, dtype=np.int64).reshape(-1, _NV[t])
        pad = -np.ones((arr.shape[0], 8), dtype=np.int64)
        pad[:, :_NV[t]] = arr
        nodes.append(pad)
        types.append(np.full(arr.shape[0], t, dtype=np.int64))
        tags.append(np.asarray(tg, dtype=np.int64) if tg is not None else np.zeros(arr.shape[0], dtype=np.int64))

    put(TRI, np.asarray(tris).reshape(-1, 3)[:, ::-1], tri_tags)
    put(QUAD, np.asarray(quads).reshape(-1, 4)[:, ::-1], quad_tags)
    put(TET, el["tet"])
    put(PYRAMID, el["pyramid"])
    put(PRISM, el["prism"])
    put(HEX, el["hex"])
    return np.concatenate(types), np.concatenate(nodes), np.concatenate(tags)


# nodes connected to each node of an element, in the order Mesh::BuildPsp walks them (mesh.tcc:541-548)
_NODE_NEIGHBOURS = {
    TRI: [(1, 2), (2, 0), (0, 1)],
    QUAD: [(1, 3), (2, 0), (3, 1), (0, 2)],
    TET: [(1, 2, 3), (2, 0, 3), (0, 1, 3), (0, 1, 2)],
    PYRAMID: [(1, 3, 4), (2, 0, 4), (3, 1, 4), (0, 2, 4), (0, 1, 2, 3)],
    PRISM: [(1, 2, 3), (2, 0, 4), (0, 1, 5), (4, 5, 0), (5, 3, 1), (3, 4, 2)],
    HEX: [(1, 3, 4), (2, 0, 5), (3, 1, 6), (0, 2, 7), (5, 7, 0), (6, 4, 1), (7, 5, 2), (4, 6, 3)],
}


def build_maps(nnode, gnode, elem_type, elem_nodes, elem_factag):
    """The reference's connectivity maps IN THE REFERENCE'S ORDER (Mesh::BuildElsp / BuildPsp / BuildEdges,
    ucs/mesh.tcc:397-515, 517-713, 1056-1167) from the element list in its winding and order: `psp` of a node lists its
    neighbours by first occurrence while walking the node's elements in list order and each element's connected nodes in
    table order (surface elements included); interior edges are (i, psp entry > i) in that order; one boundary half-edge
    per (surface element, local node) with its own phantom node, then one per (local node, ghost neighbour).  With these
    arrays the hot path's sums run in the reference's order, i.e. results on a mesh built here equal ucs.x's on the same
    element list bit for bit.  Vectorised numpy (sorts + first-occurrence unique), set-up."""
    et = np.asarray(elem_type).astype(np.int64)
    en = np.asarray(elem_nodes).astype(np.int64).reshape(-1, 8)
    ef = np.asarray(elem_factag).astype(np.int64)
    ntot = nnode + gnode
    owner, cand, eidx, lidx = [], [], [], []
    for t, table in _NODE_NEIGHBOURS.items():
        sel = np.nonzero(et == t)[0]
        if sel.size == 0:
            continue
        E = en[sel]
        for k, nbrs in enumerate(table):
            for pos, l in enumerate(nbrs):
                owner.append(E[:, k])
                cand.append(E[:, l])
                eidx.append(sel)
                lidx.append(np.full(sel.size, pos))
    owner, cand, eidx, lidx = (np.concatenate(a) for a in (owner, cand, eidx, lidx))
    # walk order: node, then its elements in list order (BuildElsp), then the table order inside the element
    order = np.lexsort((lidx, eidx, owner))
    owner, cand = owner[order], cand[order]
    # EliminateRepeats (mesh.tcc:1341-1375): first occurrences, order kept
    key = owner * ntot + cand
    _, first = np.unique(key, return_index=True)
    first.sort()
    owner, cand = owner[first], cand[first]
    ipsp = np.zeros(ntot + 1, dtype=np.int64)
    np.add.at(ipsp, owner + 1, 1)
    ipsp = np.cumsum(ipsp)
    psp = cand
    loc = owner < nnode
    interior = loc & (cand < nnode) & (cand > owner)
    edges_n = np.stack([owner[interior], cand[interior]], axis=1)
    # boundary half-edges: surface elements in list order, their nodes in element order, interior (local) nodes only
    surf = np.nonzero(et <= QUAD)[0]
    bn, bf, be = [], [], []
    for t in (TRI, QUAD):
        sel = surf[et[surf] == t]
        for k in range(_NV[t]):
            bn.append(en[sel, k])
            bf.append(ef[sel])
            be.append(sel * 4 + k)
    bn, bf, be = (np.concatenate(a) if a else np.zeros(0, np.int64) for a in (bn, bf, be))
    o = np.argsort(be, kind="stable")
    bn, bf, be = bn[o], bf[o], be[o]
    keep = bn < nnode
    bn, bf, be = bn[keep], bf[keep], be[keep]
    nbedge = bn.size
    ghost = loc & (cand >= nnode)
    bedges_n = np.concatenate([np.stack([bn, ntot + np.arange(nbedge)], axis=1), np.stack([owner[ghost], cand[ghost]], axis=1)])
    return dict(ipsp=ipsp.astype(np.int32), psp=psp.astype(np.int32), edges_n=edges_n.astype(np.int32).reshape(-1),
                bedges_n=bedges_n.astype(np.int32).reshape(-1),
                bedges_factag=np.concatenate([bf, np.zeros(int(ghost.sum()), np.int64)]).astype(np.int32),
                bedges_elem=np.concatenate([be // 4, -np.ones(int(ghost.sum()), np.int64)]).astype(np.int32),
                bedges_local=(be % 4).astype(np.int32), nedge=int(interior.sum()), nbedge=int(nbedge), nbnode=int(nbedge),
                ngedge=int(ghost.sum()))


def mesh_from_ugrid(path, device="cpu", reorder=False):
    """.ugrid file -> the mesh description ucs.x builds for it on one partition (element list in the reference's winding
    and order, BuildPsp / BuildEdges order, median-dual metrics): what pcfd_mesh_desc takes, plus the element list.
    reorder: the solver's default start-up (reorderMesh = 1, solutionSpace.tcc:61-74) -- reverse Cuthill-McKee on the
    first maps, then Mesh::ReorderC2nMap (mesh.tcc:224-262), which reads the permutation as NEW id of each OLD node
    (nodes[i] = ordering[nodes[i]], xyz[ordering[i]] = xyz[i]), then maps and metrics on the renumbered mesh."""
    from .boxmesh import read_ugrid
    from .ordering import cuthill_mckee
    xyz, el, tris, tri_tags, quads, quad_tags = read_ugrid(path)
    et, en, ef = ugrid_to_reference_winding(el, tris, tri_tags, quads, quad_tags)
    if reorder:
        nn = xyz.shape[0]
        first = build_maps(nn, 0, et, en, ef)
        ordering = cuthill_mckee(nn, first["ipsp"], first["psp"], reverse=True).astype(np.int64)
        en = np.where(en >= 0, ordering[np.clip(en, 0, None)], -1)
        moved = np.empty_like(xyz)
        moved[ordering] = xyz
        xyz = moved
    m = median_dual_general(xyz, et, en, ef, device=device, reference_order=True)
    m.update(elem_type=et.astype(np.int32), elem_nodes=en.astype(np.int32), elem_factag=ef.astype(np.int32))
    return m

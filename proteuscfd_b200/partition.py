"""Node partitions of a mesh in the layout `udecomp` writes (ucs/decomp.cpp:20-635) -- the input format of
the multi-rank hot path.

`udecomp_partition` restates the part of udecomp that fixes the numbering a rank sees:

  * owned nodes keep their global order (decomp.cpp:121-131);
  * ghost nodes are collected from the elements split across partitions, in element order and, inside an element, in
    connectivity order, first occurrence wins (:215-239); the list is then stably regrouped by owning rank (:241-268).
    The element list is the one the reference's .ugrid reader builds: boundary triangles first -- with their winding
    reversed -- then the tetrahedra (checked against the 2- and 3-rank fixtures the reference wrote,
    tests/test_partition.py);
  * `gNodeOwner` / `gNodeLocalId` are the owner rank and the owner's local id of every ghost (mesh.h:206-210).

The dual metrics are taken from the unpartitioned median dual: an edge cut by the partition keeps its whole dual face
(every element around it touches the owned end, so it is in the rank's local + split element set) and becomes a ghost
half-edge on both sides; interior edges are those with both ends owned.  Neighbour lists are ascending here (the
reference's are in first-occurrence order; the hot path takes whatever `ipsp/psp` the host hands over).

`rcb_partition` is a recursive-coordinate-bisection stand-in for METIS (the reference keeps a `Coordpartition`
alternative next to `METISpartition`, decomp.cpp:111-112): balanced, deterministic, no external library.
"""
import numpy as np

from .dualmesh import median_dual


def rcb_partition(xyz, nranks):
    """Partition id per node by recursive coordinate bisection (longest extent first; sizes differ by at most one per
    split)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    part = np.zeros(len(xyz), dtype=np.int32)

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        left = count // 2
        ext = xyz[idx].max(axis=0) - xyz[idx].min(axis=0)
        axis = int(np.argmax(ext))
        order = idx[np.argsort(xyz[idx, axis], kind="stable")]
        cut = (len(order) * left) // count
        split(order[:cut], first, left)
        split(order[cut:], first + left, count - left)

    split(np.arange(len(xyz)), 0, int(nranks))
    return part


def ghost_nodes(tets, tris, part, rank):
    """Global ids of the ghost nodes of `rank` in udecomp's order."""
    part = np.asarray(part)
    chunks = []
    for E in (np.asarray(tris)[:, ::-1], np.asarray(tets)):      # the reference's element list: reversed triangles, then tets
        dom = part[E]
        sel = (dom != dom[:, :1]).any(axis=1) & (dom == rank).any(axis=1)
        chunks.append(E[sel].reshape(-1))
    nodes = np.concatenate(chunks)
    nodes = nodes[part[nodes] != rank]
    _, first = np.unique(nodes, return_index=True)
    ghosts = nodes[np.sort(first)]                                # first occurrence wins
    return ghosts[np.argsort(part[ghosts], kind="stable")]        # regrouped by owner


def udecomp_partition(xyz, tets, tris, tags, part, nranks, bc_lut=None, device="cpu", full=None):
    """One mesh dict per rank (the pcfd_mesh_desc arrays + gNodeOwner / gNodeLocalId + gid) for the node partition `part`.
    bc_lut: array mapping a surface tag to a BC type (adds `bedges_bctype`)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    part = np.asarray(part)
    if full is None:
        full = median_dual(xyz, tets, tris, tags, device=device)
    en = full["edges_n"].reshape(-1, 2).astype(np.int64)
    ea = full["edges_a"].reshape(-1, 4)
    bn = full["bedges_n"].reshape(-1, 2).astype(np.int64)
    ba = full["bedges_a"].reshape(-1, 4)
    owned_of = [np.nonzero(part == r)[0] for r in range(nranks)]
    out = []
    for r in range(nranks):
        owned = owned_of[r]
        ghosts = ghost_nodes(tets, tris, part, r)
        nn, gn = len(owned), len(ghosts)
        new = np.full(len(xyz), -1, dtype=np.int64)
        new[owned] = np.arange(nn)
        new[ghosts] = nn + np.arange(gn)
        g_owner = part[ghosts].astype(np.int32)
        g_local = np.zeros(gn, dtype=np.int32)
        for o in np.unique(g_owner):
            sel = g_owner == o
            g_local[sel] = np.searchsorted(owned_of[o], ghosts[sel])
        own0, own1 = part[en[:, 0]] == r, part[en[:, 1]] == r
        # interior edges: both ends owned; n0 < n1 in the local numbering
        m = own0 & own1
        a, b, av = new[en[m, 0]], new[en[m, 1]], ea[m].copy()
        sw = a > b
        a, b = np.where(sw, b, a), np.where(sw, a, b)
        av[sw, :3] *= -1.0
        o = np.lexsort((b, a))
        edges_n, edges_a = np.stack([a[o], b[o]], axis=1), av[o]
        # cut edges: owned -> ghost half-edges, normal pointing away from the owned node, whole dual face
        m = own0 ^ own1
        ga, gb, gv = en[m, 0], en[m, 1], ea[m].copy()
        sw = ~own0[m]
        ga, gb = np.where(sw, gb, ga), np.where(sw, ga, gb)
        gv[sw, :3] *= -1.0
        ga, gb = new[ga], new[gb]
        assert (gb >= nn).all(), "a cut edge ends in a node that is neither owned nor a ghost"
        o = np.lexsort((gb, ga))
        gh_n, gh_a = np.stack([ga[o], gb[o]], axis=1), gv[o]
        # boundary half-edges of the owned nodes, phantom nodes numbered behind the ghosts
        keep = part[bn[:, 0]] == r
        nbedge = int(keep.sum())
        b_n = np.stack([new[bn[keep, 0]], nn + gn + np.arange(nbedge)], axis=1)
        b_a = ba[keep]
        factag = full["bedges_factag"][keep]
        # neighbour lists of the owned nodes (ghost neighbours included), ascending
        pa = np.concatenate([edges_n[:, 0], edges_n[:, 1], gh_n[:, 0]])
        pb = np.concatenate([edges_n[:, 1], edges_n[:, 0], gh_n[:, 1]])
        o = np.lexsort((pb, pa))
        ipsp = np.zeros(nn + 1, dtype=np.int64)
        ipsp[1:] = np.cumsum(np.bincount(pa, minlength=nn))
        local_gid = np.concatenate([owned, ghosts])
        mesh = dict(
            nnode=nn, gnode=gn, nbnode=nbedge, nedge=len(edges_n), nbedge=nbedge, ngedge=len(gh_n),
            edges_n=edges_n.astype(np.int32).reshape(-1), edges_a=np.ascontiguousarray(edges_a).reshape(-1),
            bedges_n=np.concatenate([b_n, gh_n]).astype(np.int32).reshape(-1),
            bedges_a=np.ascontiguousarray(np.concatenate([b_a, gh_a])).reshape(-1),
            bedges_factag=np.concatenate([factag, np.zeros(len(gh_n), dtype=factag.dtype)]),
            xyz=np.ascontiguousarray(xyz[local_gid]).reshape(-1), vol=np.ascontiguousarray(full["vol"][owned]),
            ipsp=ipsp.astype(np.int32), psp=pb[o].astype(np.int32), gNodeOwner=g_owner, gNodeLocalId=g_local,
            gid=local_gid)
        if bc_lut is not None:
            lut = np.asarray(bc_lut, dtype=np.int32)
            mesh["bedges_bctype"] = np.concatenate([lut[factag], np.zeros(len(gh_n), dtype=np.int32)]).astype(np.int32)
        out.append(mesh)
    return out


def udecomp_elements(elem_type, elem_nodes, elem_factag, part, nranks):
    """Per rank the element list udecomp writes (ucs/decomp.cpp:155-210, 414-490) and the solver reads back: the elements
    wholly inside the rank in global list order, then the elements split across ranks that touch it, in global list
    order; nodes in the rank's numbering -- owned nodes in global order, ghosts behind them in first-occurrence order over
    the split elements, regrouped by owner (:215-268).  Any element types (8 node slots, -1 padded).
    Returns per rank: dict(elem_type, elem_nodes, elem_factag, owned, ghosts, gNodeOwner, gNodeLocalId)."""
    et = np.asarray(elem_type).astype(np.int64)
    en = np.asarray(elem_nodes).astype(np.int64).reshape(-1, 8)
    ef = np.asarray(elem_factag).astype(np.int64)
    part = np.asarray(part).astype(np.int64)
    valid = en >= 0
    dom = np.where(valid, part[np.clip(en, 0, None)], -1)
    first = dom[:, 0]
    whole = ((dom == first[:, None]) | ~valid).all(axis=1)
    owned_of = [np.nonzero(part == r)[0] for r in range(nranks)]
    out = []
    for r in range(nranks):
        local = np.nonzero(whole & (first == r))[0]
        split = np.nonzero(~whole & (dom == r).any(axis=1))[0]
        nodes = en[split][valid[split]]
        nodes = nodes[part[nodes] != r]
        _, fi = np.unique(nodes, return_index=True)
        ghosts = nodes[np.sort(fi)]
        ghosts = ghosts[np.argsort(part[ghosts], kind="stable")]
        owned = owned_of[r]
        new = np.full(part.size, -1, dtype=np.int64)
        new[owned] = np.arange(owned.size)
        new[ghosts] = owned.size + np.arange(ghosts.size)
        sel = np.concatenate([local, split])
        loc_nodes = np.where(valid[sel], new[np.clip(en[sel], 0, None)], -1)
        g_owner = part[ghosts].astype(np.int32)
        g_local = np.zeros(ghosts.size, dtype=np.int32)
        for o in np.unique(g_owner):
            s = g_owner == o
            g_local[s] = np.searchsorted(owned_of[o], ghosts[s])
        out.append(dict(elem_type=et[sel].astype(np.int32), elem_nodes=loc_nodes.astype(np.int32), elem_factag=ef[sel].astype(np.int32),
                        owned=owned, ghosts=ghosts, gNodeOwner=g_owner, gNodeLocalId=g_local))
    return out

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

"""Wall distance for the Spalart-Allmaras model: the reference's octree method (ComputeWallDistOct, ucs/walldist.tcc:116-199)
returns, for every local node (owned and ghost), the distance to the nearest VISCOUS WALL NODE -- the left nodes of the
no-slip boundary half-edges of all ranks (SyncParallelPoint :24-113) -- found through an octree (ucs/octree.tcc).  This is
set-up work (once per static mesh), done here as an exact chunked nearest-point search with torch on whichever device it
is given; the arithmetic per pair is the reference's `Distance` (ucs/geometry.h:30-37: sqrt(dx*dx + dy*dy + dz*dz)), so the
field equals the reference's bit for bit wherever its octree search returns the true nearest point
(tests/test_walldist.py: the whole `wallDistance` field of the reference's SA fixture).

The result is what the host hands to the hot path as field PCFD_F_WALLDIST (SolutionSpace field "wallDistance").
"""
import numpy as np
import torch

BC_NOSLIP = 4   # ucs/bc_defines.h (Proteus_NoSlip)


def wall_points(mesh):
    """Coordinates of this partition's viscous wall nodes, one per no-slip boundary half-edge in half-edge order
    (duplicates kept, as the reference does; walldist.tcc:146-173)."""
    nb = int(mesh["nbedge"]) + int(mesh.get("ngedge", 0))
    bn = np.asarray(mesh["bedges_n"]).reshape(-1, 2)[:nb]
    bt = np.asarray(mesh["bedges_bctype"])[:nb]
    xyz = np.asarray(mesh["xyz"], dtype=np.float64).reshape(-1, 3)
    return np.ascontiguousarray(xyz[bn[bt == BC_NOSLIP, 0]])


def nearest_distance(xyz, points, device="cpu", chunk=1 << 22):
    """min over `points` of |x - p| for every row x of `xyz` (float64).  Work is chunked so that at most `chunk` pairs are
    in flight at a time (10 M-cell meshes: 1.7 M nodes x 14 k wall nodes on the GPU)."""
    dev = torch.device(device)
    X = torch.as_tensor(np.ascontiguousarray(xyz), dtype=torch.float64, device=dev).reshape(-1, 3)
    P = torch.as_tensor(np.ascontiguousarray(points), dtype=torch.float64, device=dev).reshape(-1, 3)
    out = torch.full((X.shape[0],), float("inf"), dtype=torch.float64, device=dev)
    if P.shape[0] == 0 or X.shape[0] == 0:
        return out.cpu().numpy()
    rows = max(1, chunk // P.shape[0])
    for i in range(0, X.shape[0], rows):
        x = X[i:i + rows]
        dx = x[:, None, 0] - P[None, :, 0]
        dy = x[:, None, 1] - P[None, :, 1]
        dz = x[:, None, 2] - P[None, :, 2]
        out[i:i + rows] = (dx * dx + dy * dy + dz * dz).min(dim=1).values
    # the square root is monotone and correctly rounded, so sqrt(min s) == min sqrt(s) bit for bit; it is taken on the
    # host with numpy (torch's vectorised CPU sqrt is not correctly rounded in float64)
    return np.sqrt(out.cpu().numpy())


def wall_distance(mesh, group=None, device="cpu"):
    """The "wallDistance" field of one partition: [nnode + gnode].  group: an object with allgather(obj) (parallel.TorchGroup
    / LocalGroup) when the mesh is partitioned -- every rank needs every rank's wall nodes."""
    pts = wall_points(mesh)
    if group is not None:
        pts = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in group.allgather(pts)])
    nl = int(mesh["nnode"]) + int(mesh.get("gnode", 0))
    xyz = np.asarray(mesh["xyz"], dtype=np.float64).reshape(-1, 3)[:nl]
    return nearest_distance(xyz, pts, device=device)

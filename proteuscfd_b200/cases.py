"""Synthetic cases for parity tests and the benchmark.

`box_case(n, ...)` builds the tetrahedralised box, its median-dual metrics, the
reference-style boundary-condition table and the smooth non-uniform state the
survey prescribes; everything comes back as the flat arrays the C ABI takes.
"""
import numpy as np

from . import capi
from .boxmesh import kuhn_box, renumber
from .dualmesh import median_dual
from .ordering import color_order, kuhn_box_brick_order, kuhn_box_colors

# tag -> BC type (the layout tools/make_golden.py uses for the box fixtures)
BOX_BC = {1: capi.BC_FARFIELD, 2: capi.BC_FARFIELD, 3: capi.BC_SYMMETRY, 4: capi.BC_IMPERMEABLE_WALL,
          5: capi.BC_FARFIELD, 6: capi.BC_FARFIELD}
# docs/master.bc:5-10 layout of the 15-degree ramp
# laminar Navier-Stokes box (tools/make_golden.py ns_bc): no-slip floor (zmin), symmetry side walls
NS_BC = {1: capi.BC_FARFIELD, 2: capi.BC_FARFIELD, 3: capi.BC_SYMMETRY, 4: capi.BC_SYMMETRY,
         5: capi.BC_NOSLIP, 6: capi.BC_FARFIELD}
RAMP_BC = {1: capi.BC_FARFIELD, 2: capi.BC_FARFIELD, 3: capi.BC_IMPERMEABLE_WALL, 4: capi.BC_FARFIELD,
           5: capi.BC_SYMMETRY, 6: capi.BC_SYMMETRY}


def aux_vars(Q, gamma):
    """ComputeAuxiliaryVariables (ucs/compressible.tcc:1230-1243) on rows of [rho,ru,rv,rw,rE | T,P,u,v,w]."""
    gm1 = gamma - 1.0
    u, v, w = Q[:, 1] / Q[:, 0], Q[:, 2] / Q[:, 0], Q[:, 3] / Q[:, 0]
    P = gm1 * (Q[:, 4] - 0.5 * Q[:, 0] * (u * u + v * v + w * w))
    Q[:, 6] = P
    Q[:, 5] = gamma * P / Q[:, 0]
    Q[:, 7], Q[:, 8], Q[:, 9] = u, v, w
    return Q


def freestream(mach, gamma, direction=(1.0, 0.0, 0.0)):
    """Non-dimensional free stream: rho = 1, c = 1, p = 1/gamma (param.tcc:355-437)."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    q = np.zeros((1, 10))
    q[0, 0] = 1.0
    q[0, 1:4] = mach * d
    q[0, 4] = (1.0 / gamma) / (gamma - 1.0) + 0.5 * mach * mach
    return aux_vars(q, gamma)[0]


def smooth_state(xyz, mach, gamma, amp=1.0):
    """rho = 1 + .1 sin(2 pi x) cos(2 pi y), u = M(1,0,0) + .05 (sin 2 pi y, sin 2 pi z, sin 2 pi x),
    p = (1 + .1 cos(2 pi z))/gamma."""
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    tp = 2.0 * np.pi
    rho = 1.0 + 0.1 * amp * np.sin(tp * x) * np.cos(tp * y)
    u = mach + 0.05 * amp * np.sin(tp * y)
    v = 0.05 * amp * np.sin(tp * z)
    w = 0.05 * amp * np.sin(tp * x)
    p = (1.0 + 0.1 * amp * np.cos(tp * z)) / gamma
    Q = np.zeros((len(xyz), 10))
    Q[:, 0] = rho
    Q[:, 1], Q[:, 2], Q[:, 3] = rho * u, rho * v, rho * w
    Q[:, 4] = p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v + w * w)
    return aux_vars(Q, gamma)


def box_case(n, jitter=0.15, mach=0.5, gamma=1.4, cfl=0.5, limiter=2, sorder=2, colored=False, ramp_deg=0.0,
             bc=None, device="cpu", amp=1.0, seed=1234, viscous=False, reynolds=400.0, twall=1.1, tref=300.0,
             enable_vnn=0, turb=False, brick=None):
    """Return (mesh dict, params dict, q [(nnode+nbnode)*10]) for an n^3-hex Kuhn box.  viscous=True selects the
    compressibleNS eqnset with a no-slip floor (wall temperature `twall`, non-dimensional; < 0: adiabatic)."""
    xyz, tets, tris, tags = kuhn_box(n, jitter=jitter, seed=seed, ramp_deg=ramp_deg)
    if colored:
        # colour-sorted numbering: sequential SGS == multicolour SGS (ordering.py)
        xyz, tets, tris = renumber(xyz, tets, tris, color_order(kuhn_box_colors(n)))
    elif brick:
        # brick-wise numbering: locality for the neighbour-row gathers (ordering.kuhn_box_brick_order)
        xyz, tets, tris = renumber(xyz, tets, tris, kuhn_box_brick_order(n, brick))
    mesh = median_dual(xyz, tets, tris, tags, device=device)
    table = bc or (RAMP_BC if ramp_deg else (NS_BC if viscous else BOX_BC))
    lut = np.zeros(max(table) + 1, dtype=np.int32)
    for t, b in table.items():
        lut[t] = b
    mesh["bedges_bctype"] = lut[mesh["bedges_factag"]]
    qinf = freestream(mach, gamma)
    params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=gamma,
                  chi=0.0, cfl=cfl, qinf=qinf)
    if viscous:
        params.update(eqnset=capi.EQNSET_COMPRESSIBLE_NS, viscous=1, Re=reynolds, Pr=0.72, PrT=0.85, tref=tref, mach=mach,
                      enable_vnn=enable_vnn, vnn=20.0, turb_model=1 if turb else 0)
        mesh["bedges_twall"] = np.where(mesh["bedges_bctype"] == capi.BC_NOSLIP, twall, 1.0 / tref)
    nn, nb = mesh["nnode"], mesh["nbnode"]
    q = np.zeros((nn + nb, 10))
    q[:nn] = smooth_state(mesh["xyz"].reshape(-1, 3), mach, gamma, amp)
    # phantom nodes start from the state of their left node (SetInitialConditions copies q everywhere;
    # UpdateBCs overwrites them before first use)
    q[nn:] = q[mesh["bedges_n"].reshape(-1, 2)[:, 0]]
    return mesh, params, q.reshape(-1)


# ------------------------------------------------------------------------------------------ partitioned boxes
def _slab_ranges(n, nranks, nz=None):
    """Owned planes of every rank for a box of n x n x nz hexes (default nz = n*nranks: every rank a cube, weak
    scaling): the nz + 1 node planes are dealt out in contiguous runs, as evenly as they divide (the last rank takes
    the closing plane of the default layout)."""
    if nz is None:
        nz = n * nranks
        return [(r * n, (r + 1) * n - 1 if r < nranks - 1 else nz) for r in range(nranks)], nz
    base, extra = divmod(nz + 1, nranks)
    out, k = [], 0
    for r in range(nranks):
        cnt = base + (1 if r < extra else 0)
        out.append((k, k + cnt - 1))
        k += cnt
    return out, nz


def _owned_order(n, k0, k1, colored):
    """Local id of every owned node (planes k0..k1, ascending global id, optionally colour-sorted)."""
    np1 = n + 1
    cnt = (k1 - k0 + 1) * np1 * np1
    if not colored:
        return np.arange(cnt, dtype=np.int64)
    loc = np.arange(cnt, dtype=np.int64)
    i, j, k = loc % np1, (loc // np1) % np1, loc // (np1 * np1) + k0
    color = (i + 2 * j + 4 * k) % 8
    old_of_new = np.argsort(color, kind="stable")
    new_of_old = np.empty_like(old_of_new)
    new_of_old[old_of_new] = np.arange(cnt)
    return new_of_old


def slab_case(n, rank, nranks, jitter=0.15, mach=0.5, gamma=1.4, cfl=0.5, limiter=2, sorder=2, colored=False,
              device="cpu", seed=1234, bc=None, viscous=False, reynolds=400.0, twall=1.1, tref=300.0, turb=False,
              enable_vnn=0, nz=None):
    """Partition `rank` of a box of n x n x nz hexes (default nz = n*nranks) cut into z-slabs, in the layout udecomp writes
    (ucs/decomp.cpp:122-273): owned nodes first, ghost nodes grouped by owning rank, cut edges as ghost half-edges
    carrying the full dual face, `gNodeOwner` / `gNodeLocalId` for the halo maps.  Every rank builds only its own
    slab plus one ghost plane per side; shared nodes get identical coordinates (hash jitter)."""
    from .boxmesh import kuhn_slab
    ranges, nz = _slab_ranges(n, nranks, nz)
    k0, k1 = ranges[rank]
    k_lo, k_hi = max(k0 - 1, 0), min(k1 + 1, nz)
    np1 = n + 1
    xyz, tets, tris, tags, gid = kuhn_slab(n, k_lo, k_hi, nz, jitter=jitter, seed=seed)
    full = median_dual(xyz, tets, tris, tags, device=device)
    K = gid // (np1 * np1)
    owned = (K >= k0) & (K <= k1)
    nnode = int(owned.sum())
    # new local ids: owned (ascending gid or colour-sorted), then ghosts by (owner, gid)
    new = np.full(len(gid), -1, dtype=np.int64)
    new[owned] = _owned_order(n, k0, k1, colored)
    owner = np.where(K < k0, rank - 1, rank + 1)
    gsel = np.nonzero(~owned)[0]
    gsel = gsel[np.lexsort((gid[gsel], owner[gsel]))]
    gnode = len(gsel)
    new[gsel] = nnode + np.arange(gnode)
    g_owner = owner[gsel].astype(np.int32)
    g_local = np.zeros(gnode, dtype=np.int32)
    for peer in np.unique(g_owner):
        pk0, pk1 = ranges[peer]
        order = _owned_order(n, pk0, pk1, colored)
        sel = g_owner == peer
        g_local[sel] = order[gid[gsel[sel]] - pk0 * np1 * np1]

    en = full["edges_n"].reshape(-1, 2).astype(np.int64)
    ea = full["edges_a"].reshape(-1, 4).copy()
    a, b = new[en[:, 0]], new[en[:, 1]]
    oa, ob = owned[en[:, 0]], owned[en[:, 1]]
    # interior edges: both ends owned; orient n0 < n1 in the new numbering
    m = oa & ob
    ia_, ib_, iav = a[m], b[m], ea[m]
    sw = ia_ > ib_
    ia_, ib_ = np.where(sw, ib_, ia_), np.where(sw, ia_, ib_)
    iav[sw, :3] *= -1.0
    o = np.lexsort((ib_, ia_))
    edges_n = np.stack([ia_[o], ib_[o]], axis=1)
    edges_a = iav[o]
    # ghost half-edges: owned -> ghost, normal pointing away from the owned node
    m = oa ^ ob
    ga, gb, gav = a[m], b[m], ea[m]
    sw = ~oa[m]
    ga, gb = np.where(sw, gb, ga), np.where(sw, ga, gb)
    gav[sw, :3] *= -1.0
    o = np.lexsort((gb, ga))
    gh_n = np.stack([ga[o], gb[o]], axis=1)
    gh_a = gav[o]
    # BC half-edges of owned nodes, phantom nodes numbered behind the ghosts
    bn = full["bedges_n"].reshape(-1, 2).astype(np.int64)
    keep = owned[bn[:, 0]]
    bleft = new[bn[keep, 0]]
    nbedge = int(keep.sum())
    b_n = np.stack([bleft, nnode + gnode + np.arange(nbedge)], axis=1)
    b_a = full["bedges_a"].reshape(-1, 4)[keep]
    factag = full["bedges_factag"][keep]
    # psp of owned nodes (neighbours may be ghosts), ascending
    pa = np.concatenate([edges_n[:, 0], edges_n[:, 1], gh_n[:, 0]])
    pb = np.concatenate([edges_n[:, 1], edges_n[:, 0], gh_n[:, 1]])
    o = np.lexsort((pb, pa))
    psp = pb[o]
    ipsp = np.zeros(nnode + 1, dtype=np.int64)
    ipsp[1:] = np.cumsum(np.bincount(pa, minlength=nnode))
    inv = np.empty(nnode + gnode, dtype=np.int64)
    inv[new[new >= 0]] = np.nonzero(new >= 0)[0]
    table = bc or (NS_BC if viscous else BOX_BC)
    lut = np.zeros(max(table) + 1, dtype=np.int32)
    for t, b_ in table.items():
        lut[t] = b_
    mesh = dict(
        nnode=nnode, gnode=gnode, nbnode=nbedge, nedge=len(edges_n), nbedge=nbedge, ngedge=len(gh_n),
        edges_n=edges_n.astype(np.int32).reshape(-1), edges_a=np.ascontiguousarray(edges_a).reshape(-1),
        bedges_n=np.concatenate([b_n, gh_n]).astype(np.int32).reshape(-1),
        bedges_a=np.ascontiguousarray(np.concatenate([b_a, gh_a])).reshape(-1),
        bedges_bctype=np.concatenate([lut[factag], np.zeros(len(gh_n), dtype=np.int32)]).astype(np.int32),
        xyz=np.ascontiguousarray(full["xyz"].reshape(-1, 3)[inv]).reshape(-1),
        vol=np.ascontiguousarray(full["vol"][inv[:nnode]]), ipsp=ipsp.astype(np.int32), psp=psp.astype(np.int32),
        gNodeOwner=g_owner, gNodeLocalId=g_local, gid=gid[inv])
    qinf = freestream(mach, gamma)
    params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=gamma,
                  chi=0.0, cfl=cfl, qinf=qinf)
    if viscous:     # compressibleNS (+ Spalart-Allmaras): as box_case
        params.update(eqnset=capi.EQNSET_COMPRESSIBLE_NS, viscous=1, Re=reynolds, Pr=0.72, PrT=0.85, tref=tref, mach=mach,
                      enable_vnn=enable_vnn, vnn=20.0, turb_model=1 if turb else 0)
        mesh["bedges_twall"] = np.where(mesh["bedges_bctype"][:nbedge] == capi.BC_NOSLIP, twall, 1.0 / tref)
    q = np.zeros((nnode + gnode + nbedge, 10))
    q[: nnode + gnode] = smooth_state(mesh["xyz"].reshape(-1, 3), mach, gamma)
    q[nnode + gnode:] = q[b_n[:, 0]]
    return mesh, params, q.reshape(-1)


def partitioned_box_case(n, nranks, part=None, jitter=0.15, mach=0.5, gamma=1.4, cfl=0.5, limiter=2, sorder=2, seed=1234):
    """The n^3 Kuhn box cut into `nranks` node partitions in udecomp's layout (partition.py): a list of
    (mesh, params, q) per rank.  part: partition id per node (default: recursive coordinate bisection)."""
    from .partition import rcb_partition, udecomp_partition
    xyz, tets, tris, tags = kuhn_box(n, jitter=jitter, seed=seed)
    if part is None:
        part = rcb_partition(xyz, nranks)
    lut = np.zeros(max(BOX_BC) + 1, dtype=np.int32)
    for t, bc in BOX_BC.items():
        lut[t] = bc
    out = []
    qinf = freestream(mach, gamma)
    for mesh in udecomp_partition(xyz, tets, tris, tags, part, nranks, bc_lut=lut):
        params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=gamma,
                      chi=0.0, cfl=cfl, qinf=qinf)
        nl, nb = mesh["nnode"] + mesh["gnode"], mesh["nbedge"]
        q = np.zeros((nl + nb, 10))
        q[:nl] = smooth_state(mesh["xyz"].reshape(-1, 3), mach, gamma)
        q[nl:] = q[mesh["bedges_n"].reshape(-1, 2)[:nb, 0]]
        out.append((mesh, params, q.reshape(-1)))
    return out


# ------------------------------------------------------------------------------------------ reacting eqnset
UNIV_R = 8.31447215   # chem_constants.h:5


def fr_aux_vars(Q, fr):
    """ComputeAuxiliaryVariables of the reacting eqnset (ucs/compressibleFR.tcc:755-814) on rows of
    [rho_i | u v w | T | P | rho | cv_i | mol_i] (set-up only: the solver recomputes them with its own arithmetic)."""
    ch = fr["chem"]
    ns = int(ch["dims"][0])
    mw = np.asarray(ch["species_mw"], dtype=np.float64)
    Rs = UNIV_R / mw
    a = np.asarray(ch["species_nasa7"], dtype=np.float64).reshape(ns, 2, 7)
    T = Q[:, ns + 3] * fr["ref_temperature"]
    Q[:, ns + 5] = Q[:, :ns].sum(axis=1)
    Q[:, ns + 4] = ((Q[:, :ns] * fr["ref_density"]) * Rs * T[:, None]).sum(axis=1) / fr["ref_pressure"]
    s_ref = fr["ref_velocity"] ** 2 / fr["ref_temperature"]
    hi = T > 1000.0
    for i in range(ns):
        c = np.where(hi[:, None], a[i, 1], a[i, 0])
        cp = (c[:, 0] + T * (c[:, 1] + T * (c[:, 2] + T * (c[:, 3] + T * c[:, 4])))) * Rs[i]
        Q[:, ns + 6 + i] = (cp - Rs[i]) / s_ref
        Q[:, 2 * ns + 6 + i] = Q[:, i] / mw[i] / 1000.0
    return Q


def fr_box_case(n, fr, mach=0.5, jitter=0.15, cfl=5.0, limiter=2, sorder=2, colored=True, device="cpu", seed=1234,
                amp=1.0):
    """Reacting-eqnset counterpart of box_case: (mesh, params, q, beta).  `fr` carries the chemistry tables, the
    reference values and the free stream (dict: chem, ref_*, pref, dt, use_local_dt, rxn_on, qinf) -- the caller
    takes them from a case set-up or a fixture; the state is the free stream perturbed as SURVEY.md 8d prescribes."""
    xyz, tets, tris, tags = kuhn_box(n, jitter=jitter, seed=seed)
    if colored:
        xyz, tets, tris = renumber(xyz, tets, tris, color_order(kuhn_box_colors(n)))
    mesh = median_dual(xyz, tets, tris, tags, device=device)
    lut = np.zeros(max(BOX_BC) + 1, dtype=np.int32)
    for t, b in BOX_BC.items():
        lut[t] = b
    mesh["bedges_bctype"] = lut[mesh["bedges_factag"]]
    ns = int(fr["chem"]["dims"][0])
    nv = 3 * ns + 6
    qinf = np.asarray(fr["qinf"], dtype=np.float64)
    params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER_FR, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=0.0, chi=0.0,
                  cfl=cfl, fr=fr)
    if fr.get("transport") is not None:     # compressibleNSFR: Re and PrT travel in pcfd_params
        params.update(eqnset=capi.EQNSET_COMPRESSIBLE_NS_FR, Re=fr["Re"], PrT=fr.get("PrT", 0.85))
    nn, nb = mesh["nnode"], mesh["nbnode"]
    X = mesh["xyz"].reshape(-1, 3)
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    tp = 2.0 * np.pi
    q = np.zeros((nn + nb, nv))
    for k in range(ns):
        q[:nn, k] = qinf[k] * (1.0 + 0.1 * amp * np.sin(tp * x) * np.cos(tp * y)) * (1.0 + 0.05 * amp * np.sin(tp * (z + 0.17 * k)))
    q[:nn, ns + 0] = mach + 0.05 * amp * np.sin(tp * y)
    q[:nn, ns + 1] = 0.05 * amp * np.sin(tp * z)
    q[:nn, ns + 2] = 0.05 * amp * np.sin(tp * x)
    q[:nn, ns + 3] = qinf[ns + 3] * (1.0 + 0.1 * amp * np.cos(tp * z))
    fr_aux_vars(q[:nn], fr)
    q[nn:] = q[mesh["bedges_n"].reshape(-1, 2)[:, 0]]
    # preconditioning field: beta = max(betaMin, Mach^2) below sonic, 1 otherwise (solutionSpace.tcc:235-247)
    beta = np.full(nn + nb, mach * mach if mach < 1.0 else 1.0)
    return mesh, params, q.reshape(-1), beta


def fr_slab_case(n, rank, nranks, fr, mach=0.5, jitter=0.15, cfl=5.0, limiter=2, sorder=2, colored=True, device="cpu",
                 seed=1234, amp=1.0):
    """Reacting-eqnset counterpart of slab_case: partition `rank` of the n x n x (n*nranks) box in udecomp's layout with
    the fr_box_case state (owned and ghost nodes from their coordinates, phantom nodes from their left node):
    (mesh, params, q, beta)."""
    mesh, _, _ = slab_case(n, rank, nranks, jitter=jitter, mach=mach, cfl=cfl, limiter=limiter, sorder=sorder, colored=colored,
                           device=device, seed=seed)
    ns = int(fr["chem"]["dims"][0])
    nv = 3 * ns + 6
    qinf = np.asarray(fr["qinf"], dtype=np.float64)
    params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER_FR, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=0.0, chi=0.0,
                  cfl=cfl, fr=fr)
    if fr.get("transport") is not None:
        params.update(eqnset=capi.EQNSET_COMPRESSIBLE_NS_FR, Re=fr["Re"], PrT=fr.get("PrT", 0.85))
    nn, nb = mesh["nnode"] + mesh["gnode"], mesh["nbnode"]
    X = mesh["xyz"].reshape(-1, 3)
    # the slab cases stretch the box along z (n*nranks hexes of the same size): keep the state periodic per unit length
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    tp = 2.0 * np.pi
    q = np.zeros((nn + nb, nv))
    for k in range(ns):
        q[:nn, k] = qinf[k] * (1.0 + 0.1 * amp * np.sin(tp * x) * np.cos(tp * y)) * (1.0 + 0.05 * amp * np.sin(tp * (z + 0.17 * k)))
    q[:nn, ns + 0] = mach + 0.05 * amp * np.sin(tp * y)
    q[:nn, ns + 1] = 0.05 * amp * np.sin(tp * z)
    q[:nn, ns + 2] = 0.05 * amp * np.sin(tp * x)
    q[:nn, ns + 3] = qinf[ns + 3] * (1.0 + 0.1 * amp * np.cos(tp * z))
    fr_aux_vars(q[:nn], fr)
    q[nn:] = q[mesh["bedges_n"].reshape(-1, 2)[: mesh["nbedge"], 0]]
    beta = np.full(nn + nb, mach * mach if mach < 1.0 else 1.0)
    return mesh, params, q.reshape(-1), beta

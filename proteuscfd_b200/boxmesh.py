"""Synthetic Kuhn 6-tet box meshes (the benchmark/parity input).

`kuhn_box(n)` returns node coordinates, tetrahedra and tagged boundary
triangles of an n x n x n hex box split into 6 tets per hex (Kuhn/Freudenthal
triangulation), optionally with jittered interior nodes.  `write_ugrid`
writes the AFLR3 ASCII layout that the reference's `udecomp` reads
(ucs/mesh.tcc:6740-6920, ReadUGRID_Ascii) so the same mesh can be fed to the
reference build used for parity fixtures.
"""
import itertools

import numpy as np

# the six axis permutations = six monotone paths (0,0,0)->(1,1,1)
_PERMS = list(itertools.permutations(range(3)))


def kuhn_box(n, lengths=(1.0, 1.0, 1.0), jitter=0.0, seed=1234, ramp_deg=0.0, ramp_x0=0.3):
    """Return (xyz[nn,3] f64, tets[nt,4] i32, tris[nf,3] i32, tags[nf] i32).

    Node id = i + (n+1)*(j + (n+1)*k).  Boundary tags: x-min 1, x-max 2,
    y-min 3, y-max 4, z-min 5, z-max 6.  Triangles are wound so the
    right-hand normal points INTO the domain and tets so that nodes 0-1-2 see
    node 3 on their right-hand side (UGRID convention).
    `ramp_deg` shears the floor (y-min) up by that angle for x > ramp_x0,
    blending to zero at y-max: the 15-degree supersonic ramp of BASELINE.json.
    """
    np1 = n + 1
    g = np.arange(np1, dtype=np.float64) / n
    k, j, i = np.meshgrid(g, g, g, indexing="ij")
    xyz = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter / n, jitter / n, size=xyz.shape)
        interior = np.all((xyz > 1e-12) & (xyz < 1 - 1e-12), axis=1)
        xyz[interior] += d[interior]
    if ramp_deg != 0.0:
        t = np.tan(np.deg2rad(ramp_deg))
        lift = np.maximum(xyz[:, 0] - ramp_x0, 0.0) * t
        xyz[:, 1] = xyz[:, 1] + lift * (1.0 - xyz[:, 1])
    xyz = xyz * np.asarray(lengths, dtype=np.float64)

    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    # hex order: k slowest, i fastest (same as nodes)
    ci, cj, ck = (a.transpose(2, 1, 0).ravel() for a in (ci, cj, ck))

    def nid(a, b, c):
        return (a + np1 * (b + np1 * c)).astype(np.int64)

    tets = []
    for perm in _PERMS:
        off = np.zeros(3, dtype=np.int64)
        verts = [nid(ci, cj, ck)]
        for ax in perm:
            off = off.copy()
            off[ax] += 1
            verts.append(nid(ci + off[0], cj + off[1], ck + off[2]))
        t = np.stack(verts, axis=1)
        # parity of the permutation decides orientation; swap to make it positive
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b])
        if inv % 2 == 1:
            t = t[:, [0, 2, 1, 3]]
        tets.append(t)
    # interleave so the 6 tets of a hex are consecutive
    tets = np.stack(tets, axis=1).reshape(-1, 4)

    # boundary triangles: tet faces whose 3 nodes lie on one box plane
    faces = np.concatenate([tets[:, [1, 2, 3]], tets[:, [0, 3, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 1]]])
    opp = np.concatenate([tets[:, 0], tets[:, 1], tets[:, 2], tets[:, 3]])
    ijk = np.stack([faces % np1, (faces // np1) % np1, faces // (np1 * np1)], axis=2)  # [nf,3,3]
    tris, tags = [], []
    tag = 0
    for ax in range(3):
        for val in (0, n):
            tag += 1
            m = np.all(ijk[:, :, ax] == val, axis=1)
            tris.append(faces[m])
            tags.append(np.full(int(m.sum()), tag, dtype=np.int32))
            opp_sel = opp[m]
            _ = opp_sel
    tris = np.concatenate(tris)
    tags = np.concatenate(tags)
    # orient: right-hand normal must point into the domain (towards the box centre side)
    p0, p1, p2 = xyz[tris[:, 0]], xyz[tris[:, 1]], xyz[tris[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0)
    centre = xyz.mean(axis=0)
    inward = np.einsum("ij,ij->i", nrm, centre - (p0 + p1 + p2) / 3.0) > 0
    tris[~inward] = tris[~inward][:, [0, 2, 1]]
    vol = np.einsum("ij,ij->i", np.cross(xyz[tets[:, 1]] - xyz[tets[:, 0]], xyz[tets[:, 2]] - xyz[tets[:, 0]]),
                    xyz[tets[:, 3]] - xyz[tets[:, 0]])
    assert np.all(vol > 0), "negative tet volume (jitter too large?)"
    return xyz, tets.astype(np.int32), tris.astype(np.int32), tags


def write_ugrid(path, xyz, tets, tris, tags):
    """AFLR3 ASCII .ugrid (1-based), 17 significant digits so doubles round-trip."""
    with open(path, "w") as f:
        f.write(f"{len(xyz)} {len(tris)} 0 {len(tets)} 0 0 0\n")
        np.savetxt(f, xyz, fmt="%.17g")
        np.savetxt(f, tris + 1, fmt="%d")
        np.savetxt(f, tags, fmt="%d")
        np.savetxt(f, tets + 1, fmt="%d")


def renumber(xyz, tets, tris, new_of_old):
    """Apply a node permutation (new id of each old node)."""
    new_of_old = np.asarray(new_of_old)
    old_of_new = np.empty_like(new_of_old)
    old_of_new[new_of_old] = np.arange(len(new_of_old))
    return xyz[old_of_new], new_of_old[tets].astype(np.int32), new_of_old[tris].astype(np.int32)


def _hash_uniform(gid, seed, salt):
    """Counter-based uniform [0,1) per global node id (splitmix64), so every rank jitters a shared node identically."""
    with np.errstate(over="ignore"):
        x = (gid.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(salt) * np.uint64(0xD1B54A32D192ED03))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def kuhn_slab(n, k_lo, k_hi, nz_total, jitter=0.0, seed=1234):
    """Planes k_lo..k_hi (inclusive, global indices) of the Kuhn box with n x n x nz_total hexes of size 1/n.

    Returns (xyz, tets, tris, tags, gid): local node id = i + (n+1)*(j + (n+1)*(k - k_lo)), gid = the global id
    i + (n+1)*(j + (n+1)*k).  Only faces of the GLOBAL box are boundary triangles (tags as kuhn_box); the cut planes
    k_lo / k_hi carry none.  Jitter is a hash of the global id, identical on every rank that holds the node.
    """
    np1 = n + 1
    nzl = k_hi - k_lo
    ii = np.arange(np1, dtype=np.int64)
    kk = np.arange(k_lo, k_hi + 1, dtype=np.int64)
    K, J, I = np.meshgrid(kk, ii, ii, indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    gid = I + np1 * (J + np1 * K)
    xyz = np.stack([I, J, K], axis=1).astype(np.float64) / n
    if jitter > 0.0:
        interior = (I > 0) & (I < n) & (J > 0) & (J < n) & (K > 0) & (K < nz_total)
        d = np.stack([_hash_uniform(gid, seed, s) for s in (1, 2, 3)], axis=1)
        d = (2.0 * d - 1.0) * (jitter / n)
        xyz[interior] += d[interior]
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(nzl), indexing="ij")
    ci, cj, ck = (a.transpose(2, 1, 0).ravel() for a in (ci, cj, ck))

    def nid(a, b, c):
        return (a + np1 * (b + np1 * c)).astype(np.int64)

    tets = []
    for perm in _PERMS:
        off = np.zeros(3, dtype=np.int64)
        verts = [nid(ci, cj, ck)]
        for ax in perm:
            off = off.copy()
            off[ax] += 1
            verts.append(nid(ci + off[0], cj + off[1], ck + off[2]))
        t = np.stack(verts, axis=1)
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b])
        if inv % 2 == 1:
            t = t[:, [0, 2, 1, 3]]
        tets.append(t)
    tets = np.stack(tets, axis=1).reshape(-1, 4)
    vol = np.einsum("ij,ij->i", np.cross(xyz[tets[:, 1]] - xyz[tets[:, 0]], xyz[tets[:, 2]] - xyz[tets[:, 0]]),
                    xyz[tets[:, 3]] - xyz[tets[:, 0]])
    assert np.all(vol > 0), "negative tet volume"

    faces = np.concatenate([tets[:, [1, 2, 3]], tets[:, [0, 3, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 1]]])
    opp = np.concatenate([tets[:, 0], tets[:, 1], tets[:, 2], tets[:, 3]])
    coords = (I, J, K)
    limits = ((0, n), (0, n), (0, nz_total))
    tris, tags = [], []
    tag = 0
    for ax in range(3):
        for val in limits[ax]:
            tag += 1
            m = np.all(coords[ax][faces] == val, axis=1)
            f = faces[m]
            if len(f):
                # wind so the right-hand normal points into the domain: towards the opposite tet vertex
                p0, p1, p2 = xyz[f[:, 0]], xyz[f[:, 1]], xyz[f[:, 2]]
                nrm = np.cross(p1 - p0, p2 - p0)
                inward = np.einsum("ij,ij->i", nrm, xyz[opp[m]] - p0) > 0
                f[~inward] = f[~inward][:, [0, 2, 1]]
            tris.append(f)
            tags.append(np.full(len(f), tag, dtype=np.int32))
    return xyz, tets.astype(np.int32), np.concatenate(tris).astype(np.int32), np.concatenate(tags), gid


def mixed_box(n, kind="mixed", jitter=0.0, seed=4321):
    """A unit box of n^3 cells with general elements, in UGRID winding (hex / prism: bottom face counter-clockwise seen
    from the top face, then the top face; UGRID pyramids: see `_PYR_UGRID`; boundary faces with the right-hand normal
    INTO the domain).  kind: "hex", "prism" (every cell cut into two prisms along the vertical 0-2 diagonal plane),
    "pyramid" (six pyramids per cell around an extra centre node), "mixed" (columns i < n/2: prisms; the other columns:
    hexes below n/2, pyramid-split cells above, and in the top layer the pyramid under the lid cut into two tets -- all
    four element types, triangular and quadrilateral boundary faces, conforming).
    Returns xyz, {"tet", "pyramid", "prism", "hex"} -> node arrays, tris, tri_tags, quads, quad_tags (tags as kuhn_box)."""
    np1 = n + 1
    g = np.arange(np1, dtype=np.float64) / n
    kk, jj, ii = np.meshgrid(g, g, g, indexing="ij")
    xyz = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter / n, jitter / n, size=xyz.shape)
        interior = np.all((xyz > 1e-12) & (xyz < 1 - 1e-12), axis=1)
        xyz[interior] += d[interior]

    def nid(a, b, c):
        return a + np1 * (b + np1 * c)

    extra = []          # cell-centre nodes of the pyramid-split cells
    el = {"tet": [], "pyramid": [], "prism": [], "hex": []}
    tris, ttags, quads, qtags = [], [], [], []

    def bface(nodes, tag):   # nodes: right-hand normal into the domain
        (quads if len(nodes) == 4 else tris).append(nodes)
        (qtags if len(nodes) == 4 else ttags).append(tag)

    for c in range(n):
        for b in range(n):
            for a in range(n):
                v = [nid(a, b, c), nid(a + 1, b, c), nid(a + 1, b + 1, c), nid(a, b + 1, c),
                     nid(a, b, c + 1), nid(a + 1, b, c + 1), nid(a + 1, b + 1, c + 1), nid(a, b + 1, c + 1)]
                # the six faces of the cell, right-hand normal towards the cell centre, and the tag if on the box
                faces = [((v[0], v[3], v[7], v[4]), 1 if a == 0 else 0), ((v[1], v[5], v[6], v[2]), 2 if a == n - 1 else 0),
                         ((v[0], v[4], v[5], v[1]), 3 if b == 0 else 0), ((v[3], v[2], v[6], v[7]), 4 if b == n - 1 else 0),
                         ((v[0], v[1], v[2], v[3]), 5 if c == 0 else 0), ((v[4], v[7], v[6], v[5]), 6 if c == n - 1 else 0)]
                what = kind
                if kind == "mixed":
                    what = "prism" if a < n // 2 else ("hex" if c < n // 2 else "pyramid")
                if what == "hex":
                    el["hex"].append(v)
                    for f, t in faces:
                        if t:
                            bface(f, t)
                elif what == "prism":
                    el["prism"].append([v[0], v[1], v[2], v[4], v[5], v[6]])
                    el["prism"].append([v[0], v[2], v[3], v[4], v[6], v[7]])
                    for f, t in faces[:4]:
                        if t:
                            bface(f, t)
                    if c == 0:
                        bface((v[0], v[1], v[2]), 5)
                        bface((v[0], v[2], v[3]), 5)
                    if c == n - 1:
                        bface((v[4], v[6], v[5]), 6)
                        bface((v[4], v[7], v[6]), 6)
                else:
                    ctr = np1 ** 3 + len(extra)
                    extra.append(xyz[v].mean(axis=0))
                    for f, t in faces:
                        lid = kind == "mixed" and t == 6
                        if lid:      # the pyramid under the lid as two tets; the lid face as two triangles
                            el["tet"].append([f[0], f[1], f[2], ctr])
                            el["tet"].append([f[0], f[2], f[3], ctr])
                            bface((f[0], f[1], f[2]), t)
                            bface((f[0], f[2], f[3]), t)
                        else:
                            el["pyramid"].append([f[0], f[1], f[2], f[3], ctr])
                            if t:
                                bface(f, t)
    if extra:
        xyz = np.concatenate([xyz, np.array(extra)])
    width = {"tet": 4, "pyramid": 5, "prism": 6, "hex": 8}
    el = {k: np.array(v, dtype=np.int32).reshape(-1, width[k]) for k, v in el.items()}
    return (xyz, el, np.array(tris, dtype=np.int32).reshape(-1, 3), np.array(ttags, dtype=np.int32),
            np.array(quads, dtype=np.int32).reshape(-1, 4), np.array(qtags, dtype=np.int32))


# UGRID file order of a pyramid with base (b0, b1, b2, b3) (right-hand normal towards the apex) and apex p: the
# reference's reader maps file slot translation[k] to its own slot k (ReadUGRID_Ascii, ucs/mesh.tcc:6751-6758:
# {0, 3, 4, 1, 2}), so that its own winding is base 0-3, apex 4
_PYR_UGRID = [0, 3, 4, 1, 2]


def write_ugrid_general(path, xyz, el, tris, tri_tags, quads, quad_tags):
    """AFLR3 ASCII .ugrid with all element types (pyramids: internal (b0..b3, apex) -> file order)."""
    pyr = el["pyramid"]
    if len(pyr):
        filed = np.empty_like(pyr)
        filed[:, _PYR_UGRID] = pyr
        pyr = filed
    with open(path, "w") as f:
        f.write(f"{len(xyz)} {len(tris)} {len(quads)} {len(el['tet'])} {len(pyr)} {len(el['prism'])} {len(el['hex'])}\n")
        np.savetxt(f, xyz, fmt="%.17g")
        for arr in (tris, quads):
            if len(arr):
                np.savetxt(f, arr + 1, fmt="%d")
        for arr in (tri_tags, quad_tags):
            if len(arr):
                np.savetxt(f, arr, fmt="%d")
        for arr in (el["tet"], pyr, el["prism"], el["hex"]):
            if len(arr):
                np.savetxt(f, arr + 1, fmt="%d")


def read_ugrid(path):
    """AFLR3 ASCII .ugrid (the mesh file ucs.x's decomposer reads, Mesh::ReadUGRID_Ascii, ucs/mesh.tcc:6740-6935) ->
    xyz, {"tet", "pyramid", "prism", "hex"} (pyramids as (base 0-3, apex): the file's slots 0,3,4,1,2), tris, tri_tags,
    quads, quad_tags, all 0-based, file winding otherwise."""
    with open(path) as f:
        tok = f.read().split()
    nn, ntri, nquad, ntet, npyr, npri, nhex = (int(v) for v in tok[:7])
    pos = 7

    def take(count, width, dtype):
        nonlocal pos
        a = np.array(tok[pos: pos + count * width], dtype=dtype).reshape(count, width)
        pos += count * width
        return a

    xyz = take(nn, 3, np.float64)
    tris = take(ntri, 3, np.int64) - 1
    quads = take(nquad, 4, np.int64) - 1
    tri_tags = take(ntri, 1, np.int64).reshape(-1)
    quad_tags = take(nquad, 1, np.int64).reshape(-1)
    tets = take(ntet, 4, np.int64) - 1
    pyr = take(npyr, 5, np.int64) - 1
    prisms = take(npri, 6, np.int64) - 1
    hexes = take(nhex, 8, np.int64) - 1
    el = {"tet": tets.astype(np.int32), "pyramid": pyr[:, _PYR_UGRID].astype(np.int32), "prism": prisms.astype(np.int32),
          "hex": hexes.astype(np.int32)}
    return xyz, el, tris.astype(np.int32), tri_tags.astype(np.int32), quads.astype(np.int32), quad_tags.astype(np.int32)

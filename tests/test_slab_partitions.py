"""The in-memory z-slab partitioner (proteuscfd_b200/cases.py: slab_case / fr_slab_case -- the multi-GPU bench input, laid out
as udecomp writes partitions, ucs/decomp.cpp:122-273) checked on the CPU through the halo maps it implies: exchanging
any owner-defined node field through PObj's pack / unpack must give every ghost node exactly its owner's value, cut
edges must appear on both sides with opposite normals, and the union of the partitions must be the unpartitioned box."""
import numpy as np
import pytest

from proteuscfd_b200.cases import slab_case
from proteuscfd_b200.parallel import build_local_group_maps


@pytest.mark.parametrize("nr,colored,nz", [(2, False, None), (3, True, None), (4, True, None), (3, True, 7), (4, False, 5)])
def test_slab_halo_maps_deliver_owner_values(nr, colored, nz):
    """nz: a FIXED box of n x n x nz hexes dealt out to the ranks (strong scaling; uneven slabs), else one cube per rank"""
    n = 4
    parts = [slab_case(n, r, nr, colored=colored, nz=nz)[0] for r in range(nr)]
    assert sum(m["nnode"] for m in parts) == (n + 1) * (n + 1) * ((nz or n * nr) + 1)
    pobjs = build_local_group_maps([(m["gNodeOwner"], m["gNodeLocalId"]) for m in parts])
    nn = [m["nnode"] for m in parts]
    # field = global node id (as a double) and the coordinates: owners fill their rows, ghosts start poisoned
    for w, key in ((1, "gid"), (3, "xyz")):
        arrs = []
        for m in parts:
            a = np.asarray(m[key], dtype=np.float64).reshape(-1, w).copy()
            ref = a.copy()
            a[m["nnode"]:] = -777.0
            arrs.append((a.reshape(-1), ref.reshape(-1)))
        packed = [pobjs[r].pack_numpy(arrs[r][0], w) for r in range(nr)]
        for r in range(nr):
            pobjs[r].unpack_numpy(arrs[r][0], w, nn[r], [packed[p][r] for p in range(nr)])
            assert np.array_equal(arrs[r][0], arrs[r][1]), f"{key}: rank {r} ghosts differ from their owners' rows"
    # ghost tables: owner rank and the owner's LOCAL id really name the node with the same global id
    for r, m in enumerate(parts):
        for g in range(m["gnode"]):
            o, lid = int(m["gNodeOwner"][g]), int(m["gNodeLocalId"][g])
            assert o != r and parts[o]["gid"][lid] == m["gid"][m["nnode"] + g]


def test_slab_partitions_tile_the_box():
    n, nr = 3, 3
    parts = [slab_case(n, r, nr)[0] for r in range(nr)]
    # every global node is owned exactly once
    owned = np.concatenate([m["gid"][: m["nnode"]] for m in parts])
    assert len(owned) == len(np.unique(owned)) == (n + 1) * (n + 1) * (n * nr + 1)
    # the dual volumes of the owned nodes add up to the box volume (1 x 1 x nr)
    assert np.isclose(sum(m["vol"].sum() for m in parts), float(nr), rtol=1e-13)
    # a cut edge is a ghost half-edge on both sides: same area, opposite normal
    cut = {}
    for r, m in enumerate(parts):
        bn = m["bedges_n"].reshape(-1, 2)[m["nbedge"]:]
        ba = m["bedges_a"].reshape(-1, 4)[m["nbedge"]:]
        for (l, g), a in zip(bn, ba):
            key = (int(m["gid"][l]), int(m["gid"][g]))
            cut[key] = a
    assert cut
    for (a, b), v in cut.items():
        w = cut[(b, a)]
        assert np.array_equal(v[:3], -w[:3]) and v[3] == w[3]
    # interior + cut edges = the edges of the unpartitioned box
    total = sum(m["nedge"] for m in parts) + len(cut) // 2
    # Kuhn box of a x b x c hexes: edges = axis (3 families) + face diagonals (3) + body diagonal
    a, b, c = n, n, n * nr
    expect = (a * (b + 1) * (c + 1) + (a + 1) * b * (c + 1) + (a + 1) * (b + 1) * c
              + a * b * (c + 1) + a * (b + 1) * c + (a + 1) * b * c + a * b * c)
    assert total == expect


def test_fr_slab_case_shares_the_mesh_and_fills_the_state():
    import os
    golden = os.path.join(os.path.dirname(__file__), "golden", "box4_fr_implicit.npz")
    d = dict(np.load(golden))
    meta = dict(zip([str(k) for k in d["meta_keys"]], d["meta_vals"]))
    chem = {k: d[k] for k in ("species_mw", "species_nasa7", "rxn_A_EA_n", "rxn_flags", "rxn_species", "rxn_nup", "rxn_nupp",
                              "rxn_tbeff")}
    chem["dims"] = d["chem_dims"]
    fr = dict(chem=chem, ref_density=meta["ref_density"], ref_velocity=meta["ref_velocity"], ref_temperature=meta["ref_temperature"],
              ref_pressure=meta["ref_pressure"], ref_time=meta["ref_time"], ref_specific_enthalpy=meta["ref_specific_enthalpy"],
              pref=meta["Pref"], dt=meta["dt"], use_local_dt=1, rxn_on=1, qinf=d["qinf"])
    from proteuscfd_b200.cases import fr_slab_case
    for r in range(2):
        mesh, params, q, beta = fr_slab_case(4, r, 2, fr)
        ref_mesh = slab_case(4, r, 2, colored=True, cfl=5.0)[0]
        for k in ("edges_n", "bedges_n", "gNodeOwner", "gNodeLocalId", "ipsp", "psp"):
            assert np.array_equal(mesh[k], ref_mesh[k])
        ntot = mesh["nnode"] + mesh["gnode"] + mesh["nbnode"]
        Q = q.reshape(ntot, 21)
        assert np.isfinite(Q).all() and (Q[:, :5] > 0).all() and (Q[:, 8] > 0).all() and beta.shape == (ntot,)
        # rho is the sum of the species densities, P follows Dalton's law (ComputeAuxiliaryVariables)
        assert np.allclose(Q[: mesh["nnode"], 10], Q[: mesh["nnode"], :5].sum(axis=1), rtol=1e-15)
    # ghost rows of rank 0 carry the state rank 1 computed for the same nodes (state is a function of x only)
    m0, _, q0, _ = fr_slab_case(4, 0, 2, fr)
    m1, _, q1, _ = fr_slab_case(4, 1, 2, fr)
    Q0, Q1 = q0.reshape(-1, 21), q1.reshape(-1, 21)
    for g in range(m0["gnode"]):
        lid = int(m0["gNodeLocalId"][g])
        assert np.array_equal(Q0[m0["nnode"] + g], Q1[lid])

"""Synthetic cases for parity tests and the benchmark.

`box_case(n, ...)` builds the tetrahedralised box, its median-dual metrics, the
reference-style boundary-condition table and the smooth non-uniform state the
survey prescribes; everything comes back as the flat arrays the C ABI takes.
"""
import numpy as np

from . import capi
from .boxmesh import kuhn_box, renumber
from .dualmesh import median_dual
from .ordering import color_order, kuhn_box_colors

# tag -> BC type (the layout tools/make_golden.py uses for the box fixtures)
BOX_BC = {1: capi.BC_FARFIELD, 2: capi.BC_FARFIELD, 3: capi.BC_SYMMETRY, 4: capi.BC_IMPERMEABLE_WALL,
          5: capi.BC_FARFIELD, 6: capi.BC_FARFIELD}
# docs/master.bc:5-10 layout of the 15-degree ramp
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
             bc=None, device="cpu", amp=1.0, seed=1234):
    """Return (mesh dict, params dict, q [(nnode+nbnode)*10]) for an n^3-hex Kuhn box."""
    xyz, tets, tris, tags = kuhn_box(n, jitter=jitter, seed=seed, ramp_deg=ramp_deg)
    if colored:
        # colour-sorted numbering: sequential SGS == multicolour SGS (ordering.py)
        xyz, tets, tris = renumber(xyz, tets, tris, color_order(kuhn_box_colors(n)))
    mesh = median_dual(xyz, tets, tris, tags, device=device)
    table = bc or (RAMP_BC if ramp_deg else BOX_BC)
    lut = np.zeros(max(table) + 1, dtype=np.int32)
    for t, b in table.items():
        lut[t] = b
    mesh["bedges_bctype"] = lut[mesh["bedges_factag"]]
    qinf = freestream(mach, gamma)
    params = dict(eqnset=capi.EQNSET_COMPRESSIBLE_EULER, sorder=sorder, limiter=limiter, no_cvbc=0, gamma=gamma,
                  chi=0.0, cfl=cfl, qinf=qinf)
    nn, nb = mesh["nnode"], mesh["nbnode"]
    q = np.zeros((nn + nb, 10))
    q[:nn] = smooth_state(mesh["xyz"].reshape(-1, 3), mach, gamma, amp)
    # phantom nodes start from the state of their left node (SetInitialConditions copies q everywhere;
    # UpdateBCs overwrites them before first use)
    q[nn:] = q[mesh["bedges_n"].reshape(-1, 2)[:, 0]]
    return mesh, params, q.reshape(-1)

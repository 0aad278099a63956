"""Surface forces on the GPU (pcfd_forces_*): ComputeSurfaceAreas and Forces::Compute of the reference
(ucs/forces.tcc) against its own run (tests/golden/box6_ns_forces.npz, box4_nsfr_forces.npz).

Per half-edge (cp, y+, cf, pressure / viscous force terms) the kernels follow the reference's arithmetic: bit-exact for the
perfect gas where no libm call is involved (cp), 1e-12 where Sutherland's law / the species fits feed in.  The body sums are
fixed-shape tree sums, the reference's a sequential +=: 1e-12 of the sum of magnitudes."""
import numpy as np
import pytest

from tests.oracle_lib import bodies_from_fixture, load_golden
from tests.test_oracle import exact

pytestmark = pytest.mark.gpu

RTOL = 1.0e-12


def configure(ctx, g, meta):
    offs, tags, mpt, max_ = bodies_from_fixture(g)
    d = g["forces_dirs"]
    ctx.forces_configure(offs, tags, mpt, max_, g["bedges_factag"], g["forces_cg"], d[:3], d[3:], float(meta["velocity"]),
                         g["forces_surfArea"].size // 3 - 1)


def close(a, b, scale, what):
    err = np.abs(np.asarray(a) - np.asarray(b))
    assert np.all(err <= RTOL * scale + 1e-300), f"{what}: worst {np.max(err / (scale + 1e-300)):.3e} of scale"


def body_scales(g, body):
    """Sum of magnitudes of the terms behind every body sum, from the reference's own per-body results as a lower bound
    plus the surface integral of |p| A (pressure) and |tau| A (viscous) -- generous only where the sum cancels."""
    area = np.abs(g["bedges_a"].reshape(-1, 4)[:, 3]).sum()
    pmax = 10.0 * np.abs(g["forces_q"]).max()
    return np.maximum(np.abs(body), 1e-3 * pmax * area)


def check(ctx, g, meta, per_edge_exact):
    from proteuscfd_b200 import capi
    ref = g["forces_body"].reshape(-1, 18)
    sa, ba = ctx.forces_areas()
    exact(sa, g["forces_surfArea"], "surface areas per factag")
    exact(ba, np.ascontiguousarray(ref[:, 12:15]).ravel(), "projected body areas")
    body, coef = ctx.forces_compute()
    cp, yp, cf = (ctx.forces_get(w) for w in (capi.SURF_CP, capi.SURF_YPLUS, capi.SURF_CF))
    if per_edge_exact:
        exact(cp, g["forces_cp"], "cp per half-edge")
    else:
        close(cp, g["forces_cp"], np.abs(g["forces_cp"]).max(), "cp per half-edge")
    assert np.array_equal(yp == 0.0, g["forces_yp"] == 0.0) and np.abs(g["forces_yp"]).max() > 0
    close(yp, g["forces_yp"], np.abs(g["forces_yp"]), "y+ per half-edge")
    close(cf, g["forces_cf"], np.abs(g["forces_cf"]), "cf per half-edge")
    # sum-of-magnitude scales from the per-half-edge terms themselves: |p| A and |tau| A summed over the body's surfaces
    close(body, ref[:, :12], body_scales(g, ref[:, :12]), "forces, viscous forces, moments, viscous moments")
    assert np.allclose(coef, ref[:, 15:18], rtol=1e-10, atol=0.0)
    assert np.abs(ref[:, 3:6]).max() > 0


def test_forces_perfect_gas():
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box6_ns_forces")
    ctx.set_field(capi.F_Q, g["forces_q"])
    ctx.set_field(capi.F_QGRAD, g["forces_qgrad"])
    configure(ctx, g, meta)
    check(ctx, g, meta, per_edge_exact=True)


def test_forces_reacting():
    from proteuscfd_b200 import capi
    from tests.test_gpu_fr import fr_ctx
    ctx, g, meta = fr_ctx("box4_nsfr_forces")
    ctx.set_field(capi.F_Q, g["forces_q"])
    ctx.set_field(capi.F_QGRAD, g["forces_qgrad"])
    configure(ctx, g, meta)
    check(ctx, g, meta, per_edge_exact=True)


def test_forces_after_an_iteration_equal_the_oracle(oracle):
    """the path a host takes: iterate, then Forces::Compute on the device-resident state -- against the oracle fed with
    the GPU's own q and qgrad"""
    from proteuscfd_b200 import capi
    from tests.oracle_lib import Oracle
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box6_ns_forces")
    ctx.lsq_coefficients()
    ctx.set_field(capi.F_Q, g["q_pre"])
    ctx.implicit_iterate(int(meta["nSgs"]), refresh_jac=True)
    configure(ctx, g, meta)
    body, coef = ctx.forces_compute()
    o = Oracle(oracle, g, meta)
    d = o.forces_desc(g)
    _, ba = o.surface_areas(d)
    out = o.forces(d, ctx.get_field(capi.F_Q), ctx.get_field(capi.F_QGRAD), ba)
    exact(ctx.forces_get(capi.SURF_CP), out["cp"], "cp")
    close(ctx.forces_get(capi.SURF_YPLUS), out["yp"], np.abs(out["yp"]), "y+")
    close(body.ravel(), out["body"], body_scales(g, out["body"].reshape(-1, 12)).ravel(), "body sums")
    assert np.allclose(coef.ravel(), out["coef"], rtol=1e-10)


def test_forces_need_configuration():
    from proteuscfd_b200 import capi
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box6_ns_forces")
    with pytest.raises(capi.PcfdError, match="pcfd_forces_configure has not been called"):
        ctx._forces_shape = (1, 1)
        ctx.forces_compute()


def test_wall_distance_on_the_device_vs_reference_field():
    """pcfd_wall_distance (ComputeWallDistOct, ucs/walldist.tcc:116-199) against the `wallDistance` field of the reference's
    Spalart-Allmaras fixture, bit-exact, and against the torch search of walldist.py on a larger seeded box"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import NS_BC, box_case
    from proteuscfd_b200.walldist import wall_distance, wall_points
    from tests.test_gpu_parity import golden_ctx
    ctx, g, meta = golden_ctx("box6_sa_implicit")
    mesh = {k: g[k] for k in ("bedges_n", "bedges_bctype", "xyz")}
    for k in ("nnode", "gnode", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    ctx.wall_distance(wall_points(mesh))
    nn = mesh["nnode"] + mesh["gnode"]
    exact(ctx.get_field(capi.F_WALLDIST)[:nn], g["wallDistance"][:nn], "wallDistance")
    mesh2, params2, q2 = box_case(20, cfl=5.0, colored=True, viscous=True, turb=True, bc=dict(NS_BC))
    ctx2 = capi.Context(mesh2, params2)
    ctx2.wall_distance(wall_points(mesh2))
    exact(ctx2.get_field(capi.F_WALLDIST)[: mesh2["nnode"]], wall_distance(mesh2)[: mesh2["nnode"]], "wallDistance, n = 20")
    ctx2.wall_distance(np.zeros((0, 3)))
    assert np.isinf(ctx2.get_field(capi.F_WALLDIST)).all()
    ctx3, _, _ = golden_ctx("box6_implicit_sgs")
    with pytest.raises(capi.PcfdError, match="no wall-distance field"):
        ctx3.wall_distance(np.zeros((1, 3)))

"""Surface forces: the C restatement of ComputeSurfaceAreas / Forces::Compute (forces.tcc) against the
reference's own run (tests/golden/box6_ns_forces.npz, box4_nsfr_forces.npz: pressure and viscous forces and moments of two
composite bodies, cp / y+ / cf per half-edge, lift / drag / moment coefficients).  Same compiler, same libm, same
summation order: bit-exact."""
import numpy as np
import pytest

from tests.oracle_lib import FrOracle, Oracle, load_golden
from tests.test_oracle import exact


def check(o, out, g, sa, ba):
    body = g["forces_body"].reshape(-1, 18)
    exact(sa, g["forces_surfArea"], "surface areas per factag")
    exact(ba, np.ascontiguousarray(body[:, 12:15]).ravel(), "projected body areas")
    exact(out["cp"], g["forces_cp"], "cp per half-edge")
    exact(out["yp"], g["forces_yp"], "y+ per half-edge")
    exact(out["cf"], g["forces_cf"], "cf per half-edge")
    exact(out["body"], np.ascontiguousarray(body[:, :12]).ravel(), "forces, viscous forces, moments, viscous moments")
    exact(out["coef"], np.ascontiguousarray(body[:, 15:18]).ravel(), "cl, cd, cm")
    assert np.abs(body[:, 3:6]).max() > 0 and np.abs(g["forces_yp"]).max() > 0 and np.isfinite(body).all()


def test_forces_perfect_gas(oracle):
    g, meta = load_golden("box6_ns_forces")
    o = Oracle(oracle, g, meta)
    d = o.forces_desc(g)
    sa, ba = o.surface_areas(d)
    out = o.forces(d, g["forces_q"], g["forces_qgrad"], ba)
    check(o, out, g, sa, ba)


def test_forces_reacting(oracle):
    g, meta = load_golden("box4_nsfr_forces")
    o = FrOracle(oracle, g, meta)
    d = o.forces_desc(g)
    sa, ba = o.surface_areas(d)
    out = o.forces(d, g["forces_q"], g["forces_qgrad"], ba, float(meta["velocity"]))
    check(o, out, g, sa, ba)

"""The reacting-eqnset oracle (oracle/pcfd_oracle_fr.c) against fixtures produced by the reference itself.

tests/golden/box5_fr_explicit.npz and box4_fr_implicit.npz were written by tools/make_golden.py from runs of the
unmodified reference (oracle/_ref/ref_harness, equationSet = compressibleEulerFR on the reference's own
chemModels/5speciesAir.rxn).  Same loops, same order, same libm, no FMA contraction: the bar is BIT-EXACT equality.
"""
import numpy as np
import pytest

from tests.oracle_lib import FrOracle, load_golden
from tests.test_oracle import exact

# box4_nsfr_wall / box4_nsfr_adiabatic: no-slip floor (isothermal 900 K / adiabatic) under the viscous reacting eqnset
# box4_nsfr_ffv: farFieldViscous side faces next to the no-slip floor
# box4_fr_central: central-difference flux Jacobians (jacobianFieldType = jacobianBoundaryType = 1)
FR = ["box5_fr_explicit", "box4_fr_implicit", "box4_nsfr_implicit", "box4_fr_unsteady", "box4_nsfr_wall", "box4_nsfr_adiabatic",
      "box4_nsfr_ffv", "box4_fr_central", "box4_fr_gg",      # box4_fr_gg: Green-Gauss gradients (gradientType = 1)
      "box4_nsfr_sa"]                                        # Spalart-Allmaras under compressibleNSFR
IMPLICIT = ["box4_fr_implicit", "box4_nsfr_implicit", "box4_fr_unsteady", "box4_nsfr_wall", "box4_nsfr_adiabatic", "box4_nsfr_ffv",
            "box4_fr_central", "box4_fr_gg"]


@pytest.mark.parametrize("name", FR)
def test_fr_fixture_is_the_reacting_eqnset(name):
    g, meta = load_golden(name)
    assert int(meta["nspecies"]) == 5 and int(meta["neqn"]) == 9 and int(meta["nvars"]) == 21 and int(meta["nterms"]) == 14
    assert int(meta["rxnOn"]) == 1 and g["beta"][0] == 0.25       # preconditioned: beta = Mach^2 (solutionSpace.tcc:238-243)
    if "wall" in name or "adiabatic" in name or "ffv" in name:
        noslip = g["bedges_bctype"][: int(meta["nbedge"])] == 4
        tw = g["bedges_twall"][noslip]
        assert noslip.any() and np.all(tw == tw[0]) and (tw[0] < 0.0 if "adiabatic" in name else np.isclose(tw[0], 900.0 / 950.0))
    assert np.array_equal(g["species_R"], 8.31447215 / g["species_mw"])


@pytest.mark.parametrize("name", FR)
def test_fr_update_bcs(oracle, name):
    # far field (characteristic, 9x9 eigensystem + Newton on T), slip wall / symmetry (9x9 LU), 10 sub-iterations
    g, meta = load_golden(name)
    o = FrOracle(oracle, g, meta)
    q = g["q_pre"].copy()
    o.update_bcs(q, g["beta"])
    exact(q, g["q0"], "q after BC update")


@pytest.mark.parametrize("name", FR)
def test_fr_gradient_limiter_residual_timestep(oracle, name):
    g, meta = load_golden(name)
    o = FrOracle(oracle, g, meta)
    q = g["q0"].copy()
    grad = o.gradient(q, g["lsq_sw"])
    exact(grad, g["qgrad"], "qgrad")
    lim = o.limiter(q, grad)
    exact(lim, g["limiter"], "limiter")
    b = o.residual(q, grad, lim, g["beta"])
    exact(b, g["b"], "b")
    dt, dtmin = o.timestep(q, g["beta"])
    exact(dt, g["timestep"], "timestep")
    assert dtmin == g["dtmin"][0]


def test_fr_explicit_update(oracle):
    # native -> conservative, update, Newton back to temperature (solve.tcc:112-130), then ApplyDQ
    g, meta = load_golden("box5_fr_explicit")
    o = FrOracle(oracle, g, meta)
    q = g["q0"].copy()
    x = o.explicit_solve(q, g["b"], g["timestep"])
    exact(x, g["x"], "x")
    exact(q, g["q0"], "q untouched by ExplicitSolve")
    o.apply_dq(q, x)
    exact(q, g["q1"], "q1")


@pytest.mark.parametrize("name", IMPLICIT)
def test_fr_jacobian_lu_sgs(oracle, name):
    # box4_nsfr_implicit adds the analytic viscous Jacobian (compressibleFR.tcc:1713-2040) to the off-diagonal blocks
    g, meta = load_golden(name)
    o = FrOracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    exact(ia, g["ia"], "ia")
    exact(ja, g["ja"], "ja")
    q = g["q0"].copy()
    A = o.jacobian(q, g["beta"], g["timestep"], ia, ja, iau)
    exact(A, g["A"], "A")
    pv = o.prepare_sgs(iau, A)
    exact(A, g["A_lu"], "A_lu")
    exact(pv, g["pv"], "pv")
    x, ddq = o.sgs(int(meta["nSgs"]), ia, ja, iau, A, pv, g["b"])
    exact(x, g["x"], "x")
    assert ddq == g["sgs_ddq"][0]
    o.apply_dq(q, x)
    exact(q, g["q1"], "q1")


def test_nsfr_fixture_is_viscous():
    # compressibleNSFR at Re = 111: the viscous flux is a visible part of b, temperatures straddle the 1000 K switch
    # between the Sutherland law and the NASA RP-1311 fits (species.tcc:393-479)
    g, meta = load_golden("box4_nsfr_implicit")
    assert int(meta["viscous"]) == 1 and int(meta["eqnset_id"]) == 1 and 100.0 < meta["Re"] < 125.0
    T = g["q0"].reshape(-1, 21)[: int(meta["nnode"]), 8] * meta["ref_temperature"]
    assert T.min() < 1000.0 < T.max()
    assert list(g["species_fit_counts"]) == [3, 3, 2, 2, 2, 2, 3, 3, 3, 3]      # O2, O, N, N2, NO: (mu, k) ranges


def test_nsfr_viscous_part_of_residual(oracle):
    # the viscous flux really is in b: the same state without it differs at the 1e-2 level of |b|
    g, meta = load_golden("box4_nsfr_implicit")
    o = FrOracle(oracle, g, meta)
    b = o.residual(g["q0"].copy(), g["qgrad"], g["limiter"], g["beta"])
    exact(b, g["b"], "b")
    o.c.viscous = 0
    b0 = o.residual(g["q0"].copy(), g["qgrad"], g["limiter"], g["beta"])
    assert np.abs(b - b0).max() > 1e-3 * np.abs(b).max()


def test_spalart_allmaras_under_the_reacting_eqnset(oracle):
    """turbulenceModel = 1 with compressibleNSFR: the eqnset-agnostic TurbulenceModel::Compute (turb.tcc:163-339) with
    the reacting eqnset's accessors (Wilke-mixed molecular viscosity, density from the aux variables, native velocities,
    Re = Param::Re), against the reference's own run (tests/golden/box4_nsfr_sa.npz).  Same libm: bit-exact."""
    g, meta = load_golden("box4_nsfr_sa")
    assert int(meta["turbModel"]) == 1 and int(meta["viscous"]) == 1
    o = FrOracle(oracle, g, meta)
    ia, ja, iau = o.crs_init()
    tvar = g["turb_tvar0"].copy()
    out = o.turb_sa(int(meta["nSgs"]), g["turb_q"], g["turb_qgrad"], g["lsq_s"], g["wallDistance"], g["turb_dt"], ia, ja,
                    iau, tvar)
    for k in ("tgrad", "b", "A", "x", "mut"):
        exact(out[k], g["turb_" + k], "turb " + k)
    exact(tvar, g["turb_tvar1"], "tvar after the update")
    assert out["res"] == g["turb_res"][0]
    assert np.abs(g["turb_x"]).max() > 0.1 and np.abs(g["turb_mut"]).max() > 0

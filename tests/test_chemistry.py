"""Finite-rate chemistry source term (compressibleFR): the C oracle and the CUDA kernel against a fixture written by
the reference's own ChemModel (tools/make_chem_golden.py: chemModels/5speciesAir.rxn, NASA-7 data from the
reference's chemdata/BURCAT_FIXED.THR; 4096 states, 600-5800 K, both thermo ranges, all six reactions active).

Oracle: same libm as the reference, same operation order -> BIT-EXACT (rate constants and wdot).
GPU: CUDA exp / log / pow are 1-2 ulp from glibc's, and net = k_f prod_f - k_b prod_b cancels near equilibrium, so
the per-species error is measured against the magnitude the net rate cancels FROM (the oracle's wscale): 1e-12.
"""
import numpy as np
import pytest

from tests.oracle_lib import ChemOracle, load_golden
from tests.test_oracle import exact


def fixture():
    d = dict(np.load(__import__("os").path.join(__import__("tests.oracle_lib").oracle_lib.GOLDEN_DIR, "chem_5species_air.npz")))
    ns = int(d["dims"][0])
    st = d["states"].reshape(-1, ns + 1)
    return d, st[:, :ns].copy(), st[:, ns].copy()


def test_tables_are_the_reference_models():
    d, rhoi, T = fixture()
    assert list(d["species"]) == ["O2", "O", "N", "N2", "NO"]       # order of first appearance in the .rxn file
    assert d["dims"].tolist() == [5, 6]
    fl = d["rxn_flags"].reshape(6, 4)
    assert fl[:, 0].tolist() == [2] * 6                              # GuptaModArrhenius throughout (5speciesAir.rxn)
    assert fl[:, 1].tolist() == [1, 1, 0, 1, 0, 0]                   # third bodies in reactions 1, 2, 4
    assert T.min() < 1000.0 < T.max()


def test_oracle_chemistry_bit_exact(oracle):
    d, rhoi, T = fixture()
    o = ChemOracle(oracle, d)
    w, sc, kf, kb = o.mass_production(rhoi, T)
    exact(kf.reshape(-1), d["kf"], "forward rate constants")
    exact(kb.reshape(-1), d["kb"], "backward rate constants (k_f / K_c)")
    exact(w.reshape(-1), d["wdot"], "mass production rates")
    # mass is conserved up to the tabulated molecular weights (2 x 14.00674 vs 28.0134: 3e-6)
    assert np.all(np.abs(w.sum(axis=1)) <= 1e-5 * sc.sum(axis=1) + 1e-300)


def test_oracle_source_term_is_scaled_wdot(oracle):
    d, rhoi, T = fixture()
    o = ChemOracle(oracle, d)
    ns = o.ns
    n = 64
    ref_density, ref_time, ref_temperature = 1.17, 2.9e-3, 300.0
    Q = np.zeros((n, ns + 4 + ns + 8))
    Q[:, :ns] = rhoi[:n] / ref_density
    Q[:, ns + 3] = T[:n] / ref_temperature
    vol = np.linspace(1e-6, 2e-6, n)
    src = o.source_term(Q, vol, ref_density, ref_time, ref_temperature)
    assert np.all(src[:, ns:] == 0.0)
    w, _, _, _ = o.mass_production(Q[:, :ns] * ref_density, Q[:, ns + 3] * ref_temperature)
    exact(src[:, :ns], vol[:, None] * (w / (ref_density / ref_time)), "source = vol * wdot / (rho_ref / t_ref)")


@pytest.mark.gpu
def test_gpu_mass_production_vs_reference(oracle):
    from proteuscfd_b200 import capi
    d, rhoi, T = fixture()
    o = ChemOracle(oracle, d)
    _, sc, _, _ = o.mass_production(rhoi, T)
    chem = capi.Chem(d)
    w = chem.mass_production(rhoi, T)
    ref = d["wdot"].reshape(w.shape)
    err = np.abs(w - ref)
    bad = err > 1e-12 * sc
    assert not bad.any(), f"{int(bad.sum())} of {w.size} outside 1e-12 of the cancellation scale; worst {np.max(err / np.maximum(sc, 1e-300)):.3e}"
    # and most values agree to a few ulp outright
    rel = err / np.maximum(np.abs(ref), 1e-300)
    assert np.median(rel) < 1e-14


@pytest.mark.gpu
def test_gpu_source_term_device_and_host(oracle):
    import torch
    from proteuscfd_b200 import capi
    d, rhoi, T = fixture()
    o = ChemOracle(oracle, d)
    ns = o.ns
    n = rhoi.shape[0]
    ref_density, ref_time, ref_temperature = 1.17, 2.9e-3, 300.0
    stride = ns + 4 + ns + 8                      # neqn + nauxvars of the 5-species FR eqnset (compressibleFR.tcc:43-44)
    Q = np.zeros((n, stride))
    Q[:, :ns] = rhoi / ref_density
    Q[:, ns + 3] = T / ref_temperature
    vol = np.linspace(1e-6, 2e-6, n)
    ref = o.source_term(Q, vol, ref_density, ref_time, ref_temperature)
    _, sc, _, _ = o.mass_production(Q[:, :ns] * ref_density, Q[:, ns + 3] * ref_temperature)
    scale = vol[:, None] * sc / (ref_density / ref_time)
    chem = capi.Chem(d)
    host = chem.source_term(Q, vol, ref_density, ref_time, ref_temperature)
    assert np.all(np.abs(host[:, :ns] - ref[:, :ns]) <= 1e-12 * scale)
    assert np.all(host[:, ns:] == 0.0)
    dQ, dv = torch.from_numpy(Q).cuda(), torch.from_numpy(vol).cuda()
    ds = torch.empty((n, ns + 4), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    chem.source_term_device(n, stride, dQ.data_ptr(), dv.data_ptr(), ref_density, ref_time, ref_temperature, ds.data_ptr())
    torch.cuda.synchronize()
    exact(ds.cpu().numpy(), host, "device-resident entry point == host entry point")


@pytest.mark.gpu
def test_chem_create_rejects_bad_model():
    from proteuscfd_b200 import capi
    d, _, _ = fixture()
    bad = dict(d)
    bad["rxn_species"] = d["rxn_species"].copy()
    bad["rxn_species"][0] = 99
    with pytest.raises(capi.PcfdError):
        capi.Chem(bad)

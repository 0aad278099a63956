"""include/pcfd_host.hpp, reacting eqnsets: the pcfd_fr_params the C++ shim flattens from the reference's own ChemModel /
Species / Reaction / Param objects (DropIn::FillFrParams) must equal, byte for byte, the struct the Python host builds
from the harness's dump of the same objects (the struct every FR GPU parity test runs on).

CPU test: the drop-in harness (oracle/_ref/ref_harness_gpu = the unmodified reference + the shim) is run up to
pcfd_create_fr, which writes the struct out (env PCFD_HOST_DUMP_FR_PARAMS) and then fails for want of a GPU here -- on a
GPU box the run simply carries on.  Needs /root/reference (the 5-species air model and the NASA-7 data the case reads),
so it is skipped where that is absent.
"""
import ctypes as C
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not (os.path.exists("/root/reference/chemModels/5speciesAir.rxn")
                         and os.path.exists(os.path.join(REFBIN, "ref_harness_gpu"))),
                    reason="needs /root/reference and the oracle/_ref binaries")
def test_shim_fr_params_match_python_host(tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_golden as mg
    from proteuscfd_b200 import capi
    from proteuscfd_b200.boxmesh import kuhn_box
    from tests.oracle_lib import chem_tables, load_golden

    monkeypatch.setattr(mg, "GOLDEN", str(tmp_path / "golden"))      # never touch the committed fixtures
    monkeypatch.setattr(mg.tempfile, "tempdir", str(tmp_path))
    monkeypatch.setenv("PCFD_KEEP", "1")
    name = "shim_nsfr"
    mg.make_case(name, mesh=kuhn_box(3, jitter=0.15), eqnset="compressibleNSFR", nsgs=3, cfl=5.0, refvisc=2.0e-4,
                 extra=mg.FR_EXTRA.format(temp=950, pres=2000, rxn=1) + "refThermalConductivity = 0.05\nrefLength = 0.01\n")
    work = glob.glob(str(tmp_path / "pcfd_golden_*"))[0]
    dump = str(tmp_path / "fr_params.bin")
    env = dict(os.environ, HOME=work, PCFD_MPI_NP="1", PCFD_HOST_DUMP_FR_PARAMS=dump)
    subprocess.run([os.path.join(REFBIN, "ref_harness_gpu"), os.path.join(work, name), os.path.join(work, "out_gpu"), "dump"],
                   cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
    assert os.path.exists(dump), "the shim did not reach pcfd_create_fr"
    raw = open(dump, "rb").read()
    assert len(raw) == C.sizeof(capi.FrParams), "pcfd_fr_params layout differs between pcfd.h and capi.FrParams"
    shim = capi.FrParams.from_buffer_copy(raw)

    # the Python host's struct from the CPU harness's dump of the same case (what capi.Context hands to pcfd_create_fr)
    mg_golden = os.path.join(str(tmp_path / "golden"), name + ".npz")
    d = dict(np.load(mg_golden))
    meta = dict(zip([str(k) for k in d.pop("meta_keys")], d.pop("meta_vals")))
    ref = capi.FrParams()
    capi.fill_chem_model(ref.chem, chem_tables(d))
    capi.fill_transport(ref.transport, d)
    for k, m in (("ref_density", "ref_density"), ("ref_velocity", "ref_velocity"), ("ref_temperature", "ref_temperature"),
                 ("ref_pressure", "ref_pressure"), ("ref_time", "ref_time"), ("ref_specific_enthalpy", "ref_specific_enthalpy"),
                 ("pref", "Pref"), ("dt", "dt"), ("ref_viscosity", "ref_viscosity"), ("ref_k", "ref_k")):
        setattr(ref, k, float(meta[m]))
    ref.use_local_dt, ref.rxn_on = int(meta["useLocalTimeStepping"]), int(meta["rxnOn"])
    for j, v in enumerate(d["qinf"]):
        ref.qinf[j] = float(v)

    ns, nr = ref.chem.nspecies, ref.chem.nreactions
    assert (shim.chem.nspecies, shim.chem.nreactions) == (ns, nr) and ns == 5 and nr > 0
    for field, _ in capi.FrParams._fields_:
        a, b = getattr(shim, field), getattr(ref, field)
        if isinstance(a, (int, float)):
            assert a == b, field
        else:
            assert bytes(a) == bytes(b), f"pcfd_fr_params.{field} differs between the C++ shim and the Python host"
    shutil.rmtree(work, ignore_errors=True)

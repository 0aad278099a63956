#!/usr/bin/env python
"""Prepare the reacting drop-in case directories under oracle/_ref/cases/ (git-ignored, travel to the GPU box like the
oracle/_ref binaries): everything the reference harness reads for a small compressibleEulerFR / compressibleNSFR run -- the
partitioned mesh udecomp wrote, .param, .bc, the reference's own 5-species air model and the chemistry database generated
from the reference's NASA-7 records (oracle/_ref/ref_chem).  Needs /root/reference and `make -C oracle ref`; run by
__graft_entry__.build() here, read by tests/test_dropin_reference.py on the GPU box.

    python tools/make_dropin_cases.py
"""
import glob
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
CASES = os.path.join(ROOT, "oracle", "_ref", "cases")


def main():
    import make_golden as mg
    from proteuscfd_b200.boxmesh import kuhn_box
    os.makedirs(CASES, exist_ok=True)
    specs = {
        # frozen chemistry: the update agrees to ~1e-10 between glibc and CUDA libm (no finite-differenced source term)
        "fr_box4_frozen": dict(eqnset="compressibleEulerFR", nsgs=3, cfl=5.0, extra=mg.FR_EXTRA.format(temp=3000, pres=101325, rxn=0)),
        "fr_box4": dict(eqnset="compressibleEulerFR", nsgs=3, cfl=5.0, extra=mg.FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
        "nsfr_box4_frozen": dict(eqnset="compressibleNSFR", nsgs=3, cfl=5.0, refvisc=2.0e-4, bc=mg.ns_bc(900.0),
                                 extra=mg.FR_EXTRA.format(temp=950, pres=2000, rxn=0)
                                 + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    }
    scratch = tempfile.mkdtemp(prefix="pcfd_dropin_")
    golden_save, tmp_save = mg.GOLDEN, mg.tempfile.tempdir
    os.environ["PCFD_KEEP"] = "1"
    try:
        mg.GOLDEN = os.path.join(scratch, "golden")          # never touch the committed fixtures
        mg.tempfile.tempdir = scratch
        for name, kw in specs.items():
            before = set(glob.glob(os.path.join(scratch, "pcfd_golden_*")))
            mg.make_case(name, mesh=kuhn_box(4, jitter=0.15), **kw)
            work = (set(glob.glob(os.path.join(scratch, "pcfd_golden_*"))) - before).pop()
            dst = os.path.join(CASES, name)
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(work, dst, ignore=shutil.ignore_patterns("out", "*.ugrid", "chemout", "states.bin", "species.txt"))
            print("prepared", dst, sorted(os.listdir(dst)))
    finally:
        mg.GOLDEN, mg.tempfile.tempdir = golden_save, tmp_save
        os.environ.pop("PCFD_KEEP", None)
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()

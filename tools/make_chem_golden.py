#!/usr/bin/env python
"""Golden fixture for the finite-rate chemistry source term, produced by RUNNING THE REFERENCE
(oracle/_ref/ref_chem: the reference's ChemModel on the reference's chemModels/5speciesAir.rxn).

The species thermo tables come from the reference's own chemdata/BURCAT_FIXED.THR (NASA 7-coefficient records);
the reference's chemdb.hdf5 is not shipped and its generator needs h5py, so ref_chem writes the database itself
through the reference's HDF layer from the table this script extracts.

    python tools/make_chem_golden.py     ->  tests/golden/chem_5species_air.npz
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
SPECIES = ["N", "N2", "O2", "NO", "O"]
# first-line tags of the records to take (the file holds several entries per formula)
TAGS = {"N": "L 6/88N", "N2": "N2  REF ELEMENT", "O2": "O2 REF ELEMENT", "NO": "RUS 89N  1.O  1.", "O": "L 1/90O  1."}


def nasa7_record(lines, i):
    def nums(s):
        return [float(s[k:k + 15]) for k in range(0, 75, 15) if s[k:k + 15].strip()]
    l2, l3, l4 = nums(lines[i + 1]), nums(lines[i + 2]), nums(lines[i + 3])
    high = l2 + l3[:2]
    low = l3[2:] + l4[:4]
    hf298_over_R = l4[4] if len(l4) > 4 else 0.0
    mw = float(lines[i].split()[-2])   # "... A  28.01340 1" (some records use tabs: no fixed columns)
    return mw, hf298_over_R, low, high


def transport_fits(sym):
    """NASA RP-1311 viscosity (V) and conductivity (C) fits [Tlo, Thi, A, B, C, D] of a pure species from the reference's
    chemdata/trans.inp (the tables its chemDB.py stores as /species/<sym>/mu and /species/<sym>/k)."""
    lines = open(os.path.join(REFERENCE, "chemdata", "trans.inp"), errors="replace").read().splitlines()
    for i, ln in enumerate(lines):
        if ln[:16].strip() == sym and ln[16:32].strip() == "" and ln[32:36].strip().startswith("V"):
            tag = ln[34:38]
            nv, nc = int(tag[1]), int(tag[3])
            mu, k = [], []
            for l in lines[i + 1: i + 1 + nv + nc]:
                row = [float(l[2:9]), float(l[9:19])] + [float(l[20 + 15 * j: 35 + 15 * j].replace("E ", "E+")) for j in range(4)]
                (mu if l[1] == "V" else k).append(row)
            return mu, k
    raise RuntimeError(f"no transport record for {sym}")


def write_species_table(path, tab):
    """One line per species: symbol MW hf298/R low[7] high[7] nk k[nk][6] nmu mu[nmu][6] (read by oracle/harness/ref_chem.cpp)."""
    with open(path, "w") as f:
        for s, mw, hf, low, high in tab:
            mu, k = transport_fits(s)
            f.write(f"{s} {mw!r} {hf!r} " + " ".join(repr(v) for v in low) + " " + " ".join(repr(v) for v in high))
            f.write(f" {len(k)} " + " ".join(repr(v) for row in k for v in row))
            f.write(f" {len(mu)} " + " ".join(repr(v) for row in mu for v in row) + "\n")


def species_table():
    lines = open(os.path.join(REFERENCE, "chemdata", "BURCAT_FIXED.THR"), errors="replace").read().splitlines()
    out = []
    for s in SPECIES:
        idx = [i for i, ln in enumerate(lines) if ln.split()[:1] == [s] and TAGS[s] in ln and ln.rstrip().endswith("1")]
        if not idx:
            raise RuntimeError(f"no NASA-7 record for {s}")
        mw, hf, low, high = nasa7_record(lines, idx[0])
        assert len(low) == 7 and len(high) == 7, (s, low, high)
        out.append((s, mw, hf, low, high))
    return out


def main():
    tab = species_table()
    rng = np.random.default_rng(20261017)
    n = 4096
    ns = len(SPECIES)
    # mass fractions around SURVEY.md 8d's [N, N2, O2, NO, O] = [.01, .75, .20, .02, .02], density 0.02..1.5 kg/m^3,
    # T from 600 K to 5800 K so that both NASA ranges and all six reactions are exercised; a few pure-air states
    Y = np.array([0.01, 0.75, 0.20, 0.02, 0.02]) * (1.0 + 0.6 * (rng.random((n, ns)) - 0.5))
    Y[: n // 16, [0, 3, 4]] *= 1e-6
    Y /= Y.sum(axis=1, keepdims=True)
    rho = np.exp(rng.uniform(np.log(0.02), np.log(1.5), n))
    T = rng.uniform(600.0, 5800.0, n)
    T[:8] = [1000.0, 1000.0000001, 999.9999999, 200.5, 5999.0, 3000.0, 2000.0, 4000.0]
    states = np.concatenate([rho[:, None] * Y, T[:, None]], axis=1)
    work = tempfile.mkdtemp(prefix="pcfd_chem_")
    write_species_table(os.path.join(work, "species.txt"), tab)
    out = os.path.join(work, "out")
    # the model orders its species by first appearance in the reactions, not by the speciesInModel list: ask it
    np.zeros(0).tofile(os.path.join(work, "states.bin"))
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_chem"), os.path.join(REFERENCE, "chemModels", "5speciesAir"),
                    os.path.join(work, "species.txt"), os.path.join(work, "states.bin"), out],
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(os.environ, HOME=work), check=True)
    names = open(os.path.join(out, "species_names.txt")).read().split()
    perm = [SPECIES.index(nm) for nm in names]
    states = np.concatenate([states[:, perm], states[:, -1:]], axis=1)
    states.astype(np.float64).tofile(os.path.join(work, "states.bin"))
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_chem"), os.path.join(REFERENCE, "chemModels", "5speciesAir"),
                        os.path.join(work, "species.txt"), os.path.join(work, "states.bin"), out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, HOME=work))
    if r.returncode != 0:
        print(r.stdout[-3000:])
        raise SystemExit("ref_chem failed")
    ld = lambda name, dt=np.float64: np.fromfile(os.path.join(out, name + ".bin"), dtype=dt)
    dims = ld("dims", np.int32)
    d = dict(dims=dims, species_mw=ld("species_mw"), species_nasa7=ld("species_nasa7"), rxn_A_EA_n=ld("rxn_A_EA_n"),
             rxn_flags=ld("rxn_flags", np.int32), rxn_species=ld("rxn_species", np.int32), rxn_nup=ld("rxn_nup"),
             rxn_nupp=ld("rxn_nupp"), rxn_tbeff=ld("rxn_tbeff"), states=states.reshape(-1), wdot=ld("wdot"), kf=ld("kf"),
             kb=ld("kb"), species=np.array(names))
    path = os.path.join(ROOT, "tests", "golden", "chem_5species_air.npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB; ns={dims[0]} nr={dims[1]} states={n}")
    print("MW", d["species_mw"], "flags", d["rxn_flags"].reshape(-1, 4).tolist())
    print("wdot sample", d["wdot"][:10])


if __name__ == "__main__":
    main()

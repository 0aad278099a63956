"""Time the UNMODIFIED reference (oracle/_ref/ref_harness, built by oracle/Makefile from
/root/reference) on a synthetic Kuhn box.  TEST/BENCH INFRASTRUCTURE ONLY: used by
bench.py's `cpu_baseline` leg and `--impl reference`; never by the product.

The reference runs one MPI rank per partition (ucs/main.cpp:70-75); ranks are processes
of oracle/mpi_shim, partitions are slabs written through the METIS stub
(oracle/harness/metis_stub) and cut by the reference's own udecomp (ucs/decomp.cpp).
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFBIN = os.path.join(ROOT, "oracle", "_ref")

PARAM_TMPL = """<<<BEGIN TEMPORAL CONTROL>>>
numTimeSteps = 1
newtonIterations = 1
<<<END TEMPORAL CONTROL>>>

<<<BEGIN SOLUTION ORDERING>>>
Iterate {name}
<<<END SOLUTION ORDERING>>>

<<<BEGIN SPACE {name}>>>
equationSet = compressibleEuler
fluxType = roeFlux
spatialOrder = {sorder}
limiter = {limiter}
numberSGS = {nsgs}
reorderMesh = 0
refPressure = 101325
velocity = {mach}
CFL = {cfl}
flowDirection = [1.0, 0.0, 0.0]
jacobianFieldType = {jactype}
jacobianBoundaryType = {jactype}
gradientType = {gradtype}
{extra}<<<END SPACE>>>
"""

BOX_BC = """surface #1 = farField "xmin"
surface #2 = farField "xmax"
surface #3 = symmetry "ymin"
surface #4 = impermeableWall "ymax"
surface #5 = farField "zmin"
surface #6 = farField "zmax"
"""


def available(dropin=False):
    bins = ("ref_harness", "udecomp_ref") + (("ref_harness_gpu",) if dropin else ())
    return all(os.access(os.path.join(REFBIN, b), os.X_OK) for b in bins)


def _run(cmd, cwd, env=None, timeout=None):
    e = dict(os.environ)
    e["HOME"] = cwd
    if env:
        e.update(env)
    r = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed with {r.returncode}:\n{r.stdout[-3000:]}")
    return r.stdout


def block_factors(ranks):
    """(px, py, pz) with px*py*pz == ranks, as cubical as possible (small cut surface: the reference's
    udecomp aborts when a part has more split elements than nelem/np + 35, ucs/decomp.cpp:145-205)."""
    f = min(((px, py, ranks // px // py) for px in range(1, ranks + 1) if ranks % px == 0
             for py in range(1, ranks // px + 1) if (ranks // px) % py == 0), key=lambda t: (sum(t), t))
    return f


def block_partition(xyz, ranks):
    px, py, pz = block_factors(ranks)
    x = np.clip(xyz, 0.0, 1.0 - 1e-12)
    ix = (x[:, 0] * px).astype(np.int64)
    iy = (x[:, 1] * py).astype(np.int64)
    iz = (x[:, 2] * pz).astype(np.int64)
    return ix + px * (iy + py * iz)


class ReferenceCase:
    """A box case decomposed for `ranks` reference processes, kept on disk so it can be timed repeatedly."""

    def __init__(self, n, ranks, limiter=2, nsgs=0, sorder=2, mach=0.5, cfl=0.5, jitter=0.15, colored=False, jactype=0,
                 gradtype=0, forces=False):
        from proteuscfd_b200.boxmesh import kuhn_box, renumber, write_ugrid
        from proteuscfd_b200.ordering import color_order, kuhn_box_colors
        self.work = tempfile.mkdtemp(prefix="pcfd_refbench_")
        self.name = "box"
        self.ranks = int(ranks)
        xyz, tets, tris, tags = kuhn_box(n, jitter=jitter)
        part = block_partition(xyz, self.ranks)
        if colored:
            new_of_old = color_order(kuhn_box_colors(n))
            xyz, tets, tris = renumber(xyz, tets, tris, new_of_old)
            p2 = np.empty_like(part)
            p2[new_of_old] = part
            part = p2
        with open(os.path.join(self.work, "box.param"), "w") as f:
            f.write(PARAM_TMPL.format(name=self.name, sorder=sorder, limiter=limiter, nsgs=nsgs, mach=mach, cfl=cfl,
                                     jactype=int(jactype), gradtype=int(gradtype),
                                     extra="liftDirection = [0.3, 1.0, 0.2]\ndragDirection = [1.0, 0.1, 0.0]\n" if forces else ""))
        self.forces = bool(forces)
        with open(os.path.join(self.work, "box.bc"), "w") as f:
            # forces: two composite bodies (the impermeable wall; three far-field faces) for Forces::Compute
            f.write(BOX_BC + ("\nbody #1 = [4]\nbody #2 = [1,2,6]\n" if forces else ""))
        write_ugrid(os.path.join(self.work, "box.ugrid"), xyz, tets, tris, tags)
        env = {}
        if self.ranks > 1:
            np.savetxt(os.path.join(self.work, "part.txt"), part, fmt="%d")
            env["PCFD_PARTITION_FILE"] = os.path.join(self.work, "part.txt")
        _run([os.path.join(REFBIN, "udecomp_ref"), "box.ugrid", str(self.ranks)], self.work, env)
        os.remove(os.path.join(self.work, "box.ugrid"))

    def time(self, reps, timeout=1500):
        """Run `reps` iterations; return the reference's per-phase seconds per iteration (max over ranks)."""
        out = os.path.join(self.work, "out")
        _run([os.path.join(REFBIN, "ref_harness"), os.path.join(self.work, self.name), out, "time", str(int(reps))],
             self.work, {"PCFD_MPI_NP": str(self.ranks)}, timeout=timeout)
        with open(os.path.join(out, "timing.json")) as f:
            return json.load(f)

    INT_ARRAYS = {"edges_n", "bedges_n", "bedges_factag", "bedges_bctype", "ipsp", "psp", "gNodeOwner", "gNodeLocalId",
                  "commCountsSend", "commCountsRecv", "commOffsetsRecv", "nodePackingList", "ia", "ja", "iau", "pv",
                  "forces_body_lists"}

    def dump(self, dropin=False, timeout=600):
        """One pass with every intermediate array dumped; dropin=True runs oracle/_ref/ref_harness_gpu, the same
        harness with the phase calls replaced by include/pcfd_host.hpp (needs a B200).  Returns {rank: {name: array}}."""
        binary = "ref_harness_gpu" if dropin else "ref_harness"
        out = os.path.join(self.work, "out_gpu" if dropin else "out_cpu")
        env = {"PCFD_MPI_NP": str(self.ranks)}
        if self.forces:
            env["PCFD_FORCES"] = "1"
        _run([os.path.join(REFBIN, binary), os.path.join(self.work, self.name), out, "dump"], self.work, env, timeout=timeout)
        res = {}
        for r in range(self.ranks):
            d = {}
            for fn in sorted(os.listdir(out)):
                if fn.endswith(f".{r}.bin"):
                    name = fn[: -len(f".{r}.bin")]
                    d[name] = np.fromfile(os.path.join(out, fn), dtype=np.int32 if name in self.INT_ARRAYS else np.float64)
            res[r] = d
        return res

    def close(self):
        shutil.rmtree(self.work, ignore_errors=True)


if __name__ == "__main__":
    n, ranks, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    nsgs = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    c = ReferenceCase(n, ranks, nsgs=nsgs, cfl=5.0 if nsgs else 0.5)
    try:
        print(json.dumps(c.time(reps)))
    finally:
        c.close()

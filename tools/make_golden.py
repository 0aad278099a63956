#!/usr/bin/env python
"""Generate the golden parity fixtures under tests/golden/ by RUNNING THE REFERENCE.

Needs /root/reference and the binaries built by `make -C oracle ref`
(oracle/_ref/udecomp_ref, oracle/_ref/ref_harness).  For each case it writes a
synthetic mesh (or copies one of the reference's own unit-test fixtures),
partitions it with the reference's udecomp, runs the reference harness (which
calls the reference's own Gradient::Compute / Limiter::Compute /
ComputeResiduals / ComputeJacobians / CRS::SGS ...) and packs every dumped
array into tests/golden/<case>.npz.  The fixtures are small and committed; this
script is how they were made.

    python tools/make_golden.py [case ...]
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proteuscfd_b200.boxmesh import kuhn_box, mixed_box, renumber, write_ugrid, write_ugrid_general  # noqa: E402
from proteuscfd_b200.ordering import greedy_color_order  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

PARAM_TMPL = """<<<BEGIN TEMPORAL CONTROL>>>
numTimeSteps = 1
newtonIterations = 1
<<<END TEMPORAL CONTROL>>>

<<<BEGIN SOLUTION ORDERING>>>
Iterate {name}
<<<END SOLUTION ORDERING>>>

<<<BEGIN SPACE {name}>>>
equationSet = {eqnset}
fluxType = roeFlux
spatialOrder = {sorder}
limiter = {limiter}
numberSGS = {nsgs}
reorderMesh = {reorder}
refPressure = 101325
velocity = {mach}
CFL = {cfl}
flowDirection = [{fx}, {fy}, {fz}]
jacobianFieldType = {jactype}
jacobianBoundaryType = {jactype_b}
refViscosity = {refvisc}
enableVNN = {vnn}
turbulenceModel = {turb}
turbulenceSpatialOrder = 1
{extra}<<<END SPACE>>>
"""

# reacting eqnset (compressibleEulerFR): 5-species air (chemModels/5speciesAir.rxn), free stream at 3000 K so that all
# six reactions are active; mass fractions in the MODEL's species order [O2, O, N, N2, NO]
FR_EXTRA = """initialPressure = {pres}
initialTemperature = {temp}
chemicalDatabase = chemdb.hdf5
massFractions = [0.20, 0.02, 0.01, 0.75, 0.02]
reactionsOn = {rxn}
"""

INT_ARRAYS = {"rcm_ordering", "elem_type", "elem_factag", "elem_nodes", "forces_body_lists", "species_fit_counts", "rxn_flags", "rxn_species", "chem_dims", "edges_n", "bedges_n", "bedges_factag", "bedges_bctype", "ipsp", "psp", "gNodeOwner",
              "gNodeLocalId", "commCountsSend", "commCountsRecv", "commOffsetsRecv", "nodePackingList",
              "ia", "ja", "iau", "pv"}

BOX_BC = """surface #1 = farField "xmin"
surface #2 = farField "xmax"
surface #3 = symmetry "ymin"
surface #4 = impermeableWall "ymax"
surface #5 = farField "zmin"
surface #6 = farField "zmax"
"""

# docs/master.bc:5-10 layout of the 15-degree ramp: symmetry side walls,
# farField inflow/outlet/top, impermeableWall ramp floor
RAMP_BC = """surface #1 = farField "inflow"
surface #2 = farField "outlet"
surface #3 = impermeableWall "ramp"
surface #4 = farField "top"
surface #5 = symmetry "wall0"
surface #6 = symmetry "wall1"
"""


# laminar Navier-Stokes box: no-slip floor (isothermal or adiabatic through twall), far field elsewhere
def ns_bc(twall):
    return f"""surface #1 = farField "xmin"
surface #2 = farField "xmax"
surface #3 = symmetry "ymin"
surface #4 = symmetry "ymax"
surface #5 = noSlip "floor" twall = [{twall}]
surface #6 = farField "zmax"
"""


def run(cmd, cwd, env=None):
    e = dict(os.environ)
    e["HOME"] = cwd
    if env:
        e.update(env)
    r = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout[-4000:])
        raise RuntimeError(f"{cmd} failed with {r.returncode}")
    return r.stdout


def collect(outdir, rank):
    d = {}
    meta = {}
    with open(os.path.join(outdir, f"meta.{rank}.txt")) as f:
        for line in f:
            k, v = line.split()
            meta[k] = float(v)
    for fn in sorted(os.listdir(outdir)):
        if not fn.endswith(f".{rank}.bin"):
            continue
        name = fn[: -len(f".{rank}.bin")]
        dt = np.int32 if name in INT_ARRAYS else np.float64
        d[name] = np.fromfile(os.path.join(outdir, fn), dtype=dt)
    d["meta_keys"] = np.array(sorted(meta.keys()))
    d["meta_vals"] = np.array([meta[k] for k in sorted(meta.keys())], dtype=np.float64)
    return d


def make_case(name, mesh=None, h5=None, bc=BOX_BC, np_ranks=1, part=None, unsteady=False, gmres=None, forces=False, transpose=False, elements=False, ugrid=None, rcm=0, **kw):
    opts = dict(eqnset="compressibleEuler", sorder=2, limiter=2, nsgs=0, mach=0.5, cfl=0.5,
                fx=1.0, fy=0.0, fz=0.0, jactype=0, refvisc=1.0, vnn=0, turb=0, reorder=0, extra="")
    opts.update(kw)
    opts.setdefault("jactype_b", opts["jactype"])
    work = tempfile.mkdtemp(prefix="pcfd_golden_")
    try:
        if opts["eqnset"].endswith("FR"):
            fr_inputs(work, name)
        with open(os.path.join(work, f"{name}.param"), "w") as f:
            f.write(PARAM_TMPL.format(name=name, **opts))
        with open(os.path.join(work, f"{name}.bc"), "w") as f:
            f.write(bc)
        if h5 is not None:
            shutil.copy(h5, os.path.join(work, f"{name}.0.h5"))
        else:
            if ugrid is not None:      # a general-element mesh: the caller writes the .ugrid
                ugrid(os.path.join(work, f"{name}.ugrid"))
            else:
                xyz, tets, tris, tags = mesh
                write_ugrid(os.path.join(work, f"{name}.ugrid"), xyz, tets, tris, tags)
            env = {}
            if np_ranks > 1:
                np.savetxt(os.path.join(work, "part.txt"), part, fmt="%d")
                env["PCFD_PARTITION_FILE"] = os.path.join(work, "part.txt")
            run([os.path.join(REFBIN, "udecomp_ref"), f"{name}.ugrid", str(np_ranks)], work, env)
        henv = {"PCFD_MPI_NP": str(np_ranks)}
        if unsteady:
            henv["PCFD_UNSTEADY"] = "1"
        if forces:                 # Forces::Compute on the bodies the .bc file declares
            henv["PCFD_FORCES"] = "1"
        if rcm:                    # Mesh::ReorderMeshCuthillMcKee's permutation (1: Cuthill-McKee, 2: reversed)
            henv["PCFD_RCM"] = str(rcm)
        if elements:               # the element list in the reference's internal winding
            henv["PCFD_DUMP_ELEMENTS"] = "1"
        if transpose:              # CRSMatrix::CRSTranspose on the assembled matrix -> A_T
            henv["PCFD_TRANSPOSE"] = "1"
        if gmres is not None:      # (precondType, search directions, restarts): also run CRS::GMRES on the assembled system
            henv.update(PCFD_GMRES=str(gmres[0]), PCFD_GMRES_NDIR=str(gmres[1]), PCFD_GMRES_RESTARTS=str(gmres[2]))
        run([os.path.join(REFBIN, "ref_harness"), os.path.join(work, name), os.path.join(work, "out"), "dump"], work, henv)
        os.makedirs(GOLDEN, exist_ok=True)
        for r in range(np_ranks):
            d = collect(os.path.join(work, "out"), r)
            suffix = "" if np_ranks == 1 else f"_r{r}of{np_ranks}"
            path = os.path.join(GOLDEN, f"{name}{suffix}.npz")
            np.savez_compressed(path, **d)
            print(f"wrote {path}: {os.path.getsize(path)/1024:.0f} KiB, nnode={int(d['vol'].size)}")
    finally:
        if not os.environ.get('PCFD_KEEP'): shutil.rmtree(work, ignore_errors=True)
        else: print('kept', work)


def fr_inputs(work, name):
    """<case>.rxn (the reference's own 5speciesAir model) and chemdb.hdf5 (written by oracle/_ref/ref_chem through the
    reference's HDF layer from the NASA-7 records of the reference's chemdata/BURCAT_FIXED.THR)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_chem_golden as mc
    mc.write_species_table(os.path.join(work, "species.txt"), mc.species_table())
    np.zeros(0).tofile(os.path.join(work, "states.bin"))
    run([os.path.join(REFBIN, "ref_chem"), os.path.join(REFERENCE, "chemModels", "5speciesAir"),
         os.path.join(work, "species.txt"), os.path.join(work, "states.bin"), os.path.join(work, "chemout")], work)
    shutil.copy(os.path.join(work, "chemout", "chemdb.hdf5"), os.path.join(work, "chemdb.hdf5"))
    shutil.copy(os.path.join(REFERENCE, "chemModels", "5speciesAir.rxn"), os.path.join(work, f"{name}.rxn"))
    os.chmod(os.path.join(work, f"{name}.rxn"), 0o644)


def ffv_bc(twall):
    """no-slip floor with farFieldViscous side faces (free-stream momentum scaled by a 1/7 power-law profile of the wall
    distance, bc.tcc:1092-1108), far field on top"""
    return f"""surface #1 = farFieldViscous "xmin"
surface #2 = farFieldViscous "xmax"
surface #3 = symmetry "ymin"
surface #4 = symmetry "ymax"
surface #5 = noSlip "floor" twall = [{twall}]
surface #6 = farField "zmax"
"""


def slab_part(xyz, ranks, axis=2):
    """Partition id per node: equal slabs along one axis."""
    x = np.clip(xyz[:, axis], 0.0, 1.0 - 1e-12)
    return (x * ranks).astype(np.int64)


def quadrant_part(xyz):
    """Partition id per node: four columns cut at x = 0.5 and y = 0.5 (every part touches the other three)."""
    return 2 * (xyz[:, 0] > 0.5).astype(np.int64) + (xyz[:, 1] > 0.5).astype(np.int64)


def colored_box(n, **kw):
    """Box whose node numbering is colour-sorted (multicolour SGS == sequential SGS)."""
    xyz, tets, tris, tags = kuhn_box(n, **kw)
    new_of_old, _ = greedy_color_order(len(xyz), tets)
    xyz, tets, tris = renumber(xyz, tets, tris, new_of_old)
    return xyz, tets, tris, tags


CASES = {
    # config[1] in miniature: explicit Euler, Roe, 2nd order, Venkatakrishnan limiter
    "box8_explicit_venkat": lambda: make_case("box8_explicit_venkat", mesh=kuhn_box(8, jitter=0.15)),
    # Barth limiter on the same mesh
    "box8_explicit_barth": lambda: make_case("box8_explicit_barth", mesh=kuhn_box(8, jitter=0.15), limiter=1),
    # modified Venkatakrishnan limiter (limiter = 3, eps^2 = 6 pi V), limiters.tcc:534-735
    "box8_explicit_venkatmod": lambda: make_case("box8_explicit_venkatmod", mesh=kuhn_box(8, jitter=0.15), limiter=3),
    # implicit: FD Jacobian, LU, 3 SGS sweeps on natural (lexicographic) numbering
    "box6_implicit_sgs": lambda: make_case("box6_implicit_sgs", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0),
    # implicit on a colour-sorted numbering (the multicolour schedule used at scale)
    "box6c_implicit_sgs": lambda: make_case("box6c_implicit_sgs", mesh=colored_box(6, jitter=0.15), nsgs=3, cfl=5.0),
    # config[0]: 15-degree ramp, supersonic inviscid Euler, 1 partition, Roe + LSQ, 5 SGS
    "ramp15_implicit": lambda: make_case("ramp15_implicit", mesh=kuhn_box(8, ramp_deg=15.0, jitter=0.1), bc=RAMP_BC,
                                         mach=2.0, nsgs=5, cfl=5.0),
    # two reference ranks (process-based MPI shim): halo maps + multi-rank numerics, explicit and implicit
    "box8_2rank_explicit": lambda: make_case("box8_2rank_explicit", mesh=kuhn_box(8, jitter=0.15), np_ranks=2,
                                             part=slab_part(kuhn_box(8, jitter=0.15)[0], 2)),
    "box9_3rank_implicit": lambda: make_case("box9_3rank_implicit", mesh=kuhn_box(9, jitter=0.15), np_ranks=3,
                                             part=slab_part(kuhn_box(9, jitter=0.15)[0], 3, axis=0), nsgs=3, cfl=5.0),
    # config[2] in miniature: laminar Navier-Stokes (compressibleNS), viscous flux + analytic viscous Jacobian,
    # isothermal no-slip wall, implicit; Re = 146 / refViscosity
    "box6_ns_implicit": lambda: make_case("box6_ns_implicit", mesh=kuhn_box(6, jitter=0.15), bc=ns_bc(330.0),
                                          eqnset="compressibleNS", nsgs=3, cfl=5.0, refvisc=0.5),
    # adiabatic wall, explicit, Von Neumann time-step limit on
    "box6_ns_adiabatic": lambda: make_case("box6_ns_adiabatic", mesh=kuhn_box(6, jitter=0.15), bc=ns_bc(-1.0),
                                           eqnset="compressibleNS", nsgs=2, cfl=5.0, refvisc=0.25, vnn=1),
    # config[3] in miniature: RANS, Spalart-Allmaras one-equation model (segregated scalar system, first-order
    # convection), no-slip floor; Re = 146 / refViscosity
    "box6_sa_implicit": lambda: make_case("box6_sa_implicit", mesh=kuhn_box(6, jitter=0.15), bc=ns_bc(330.0),
                                          eqnset="compressibleNS", nsgs=3, cfl=5.0, refvisc=0.01, turb=1),
    # config[4] in miniature: reacting 5-species air (compressibleEulerFR), HLLC flux with preconditioned wave speeds,
    # finite-rate source term; explicit (native <-> conservative Newton solve) and implicit (9x9 blocks, FD flux and
    # source Jacobians, dense temporal terms, 3 SGS sweeps)
    "box5_fr_explicit": lambda: make_case("box5_fr_explicit", mesh=kuhn_box(5, jitter=0.15), eqnset="compressibleEulerFR",
                                          cfl=0.05, extra=FR_EXTRA.format(temp=2500, pres=2000, rxn=1)),
    "box4_fr_implicit": lambda: make_case("box4_fr_implicit", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                          nsgs=3, cfl=5.0, extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
    # the viscous reacting eqnset (compressibleNSFR): Wilke-mixed species transport (Sutherland below 1000 K, NASA
    # RP-1311 fits of the reference's chemdata/trans.inp above), viscous flux + analytic viscous Jacobian, implicit;
    # T_inf = 900 K with a +-10 % field so both transport branches are hit
    "box4_nsfr_implicit": lambda: make_case("box4_nsfr_implicit", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleNSFR",
                                            nsgs=3, cfl=5.0, refvisc=2.0e-4,   # Re = 111 (refLength 1 cm, 2 kPa)
                                            extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                            + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    # viscous reacting eqnset with a no-slip floor (isothermal 900 K / adiabatic): hard-set wall state from the
    # most-normal neighbour (bc.tcc:1182-1291, compressibleFR.tcc:2048-2070), wall rows of residual and Jacobian (:2072-2114)
    "box4_nsfr_wall": lambda: make_case("box4_nsfr_wall", mesh=kuhn_box(4, jitter=0.15), bc=ns_bc(900.0), eqnset="compressibleNSFR",
                                        nsgs=3, cfl=5.0, refvisc=2.0e-4,
                                        extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                        + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    "box4_nsfr_adiabatic": lambda: make_case("box4_nsfr_adiabatic", mesh=kuhn_box(4, jitter=0.15), bc=ns_bc(-1.0),
                                             eqnset="compressibleNSFR", nsgs=3, cfl=5.0, refvisc=2.0e-4,
                                             extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                             + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    # farFieldViscous boundaries next to a no-slip floor, laminar NS and viscous reacting
    "box6_ns_ffv": lambda: make_case("box6_ns_ffv", mesh=kuhn_box(6, jitter=0.15), bc=ffv_bc(330.0), eqnset="compressibleNS",
                                     nsgs=3, cfl=5.0, refvisc=0.5),
    "box4_nsfr_ffv": lambda: make_case("box4_nsfr_ffv", mesh=kuhn_box(4, jitter=0.15), bc=ffv_bc(900.0), eqnset="compressibleNSFR",
                                       nsgs=3, cfl=5.0, refvisc=2.0e-4,
                                       extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                       + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    # central-difference flux Jacobians (jacobianFieldType = jacobianBoundaryType = 1, jacobian.tcc:306-366, 546-640)
    "box6_implicit_central": lambda: make_case("box6_implicit_central", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0, jactype=1),
    "box4_fr_central": lambda: make_case("box4_fr_central", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                         nsgs=3, cfl=5.0, jactype=1, extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
    # Green-Gauss gradients (gradientType = 1, gradient.tcc:170-248) under both eqnset families
    "box8_explicit_gg": lambda: make_case("box8_explicit_gg", mesh=kuhn_box(8, jitter=0.15), extra="gradientType = 1\n"),
    "box4_fr_gg": lambda: make_case("box4_fr_gg", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR", nsgs=3, cfl=5.0,
                                    extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1) + "gradientType = 1\n"),
    # unsteady (dual time stepping): physical time step 0.02, BDF2 at the third step -- TemporalResidual with
    # q^n, q^{n-1} and the cnp1 V/dt + V/dtau diagonal (perfect gas: diagonal; reacting: dense dQ/dq blocks)
    "box6_unsteady_bdf2": lambda: make_case("box6_unsteady_bdf2", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0, unsteady=True,
                                            extra="timeStep = 0.02\ntimeOrder = 2\n"),
    "box4_fr_unsteady": lambda: make_case("box4_fr_unsteady", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                          nsgs=3, cfl=5.0, unsteady=True,
                                          extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1) + "timeStep = 0.02\ntimeOrder = 2\n"),
    # Spalart-Allmaras under the viscous reacting eqnset (turbulenceModel = 1 with compressibleNSFR): the eqnset-agnostic
    # TurbulenceModel::Compute with Wilke-mixed molecular viscosity, no-slip floor
    "box4_nsfr_sa": lambda: make_case("box4_nsfr_sa", mesh=kuhn_box(4, jitter=0.15), bc=ns_bc(900.0), eqnset="compressibleNSFR",
                                      nsgs=3, cfl=5.0, refvisc=2.0e-4, turb=1,
                                      extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                      + "refThermalConductivity = 0.05\nrefLength = 0.01\n"),
    # CRS::GMRES (crs.tcc:176-415) on the assembled implicit system: block-diagonal (LU) right preconditioner, 8 search
    # directions, 2 restarts (perfect gas 5x5); diagonal preconditioner, 6 directions, 1 restart (reacting 9x9)
    "box6_gmres": lambda: make_case("box6_gmres", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0, gmres=(2, 8, 2)),
    # SGS-preconditioned GMRES (precondType 4: six sweeps on a copy of the matrix per application, crs.tcc:577-581, 629-632).
    # Two directions: with this preconditioner the residual is at round-off after two, and every further direction of
    # the reference's GMRES is normalised round-off (its breakdown test is |h| < 1e-15) -- x then moves by 1e-9 .. 1e-5
    # with the summation order of a dot product, which pins nothing
    "box6_gmres_sgs": lambda: make_case("box6_gmres_sgs", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0, gmres=(4, 2, 2)),
    # the same on two reference ranks: block-Jacobi sweeps with a halo of the iterate after every sweep (crs.tcc:88,146), the
    # preconditioned vector exchanged before every product (:300), dot products through MPI_Allreduce
    "box8_2rank_gmres_sgs": lambda: make_case("box8_2rank_gmres_sgs", mesh=kuhn_box(8, jitter=0.15), np_ranks=2,
                                              part=slab_part(kuhn_box(8, jitter=0.15)[0], 2), nsgs=3, cfl=5.0, gmres=(4, 2, 2)),
    "box4_fr_gmres_sgs": lambda: make_case("box4_fr_gmres_sgs", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                           nsgs=3, cfl=5.0, gmres=(4, 2, 1), extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
    "box4_fr_gmres": lambda: make_case("box4_fr_gmres", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                       nsgs=3, cfl=5.0, gmres=(1, 6, 1), extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
    # GMRES with the local ILU0 right preconditioner (precondType 3: BuildILU0Local / ILU0BackSub, crsmatrix.tcc:276-507):
    # both block sizes, and two reference ranks (ghost columns skipped by the factorisation and blanked afterwards).  The 9x9
    # case freezes the chemistry: with the source Jacobian in the diagonal blocks the reference's pivot-free factorisation
    # divides by zero and its GMRES returns NaN, which pins nothing.
    "box6_gmres_ilu0": lambda: make_case("box6_gmres_ilu0", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0, gmres=(3, 6, 2)),
    "box4_fr_gmres_ilu0": lambda: make_case("box4_fr_gmres_ilu0", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                            nsgs=3, cfl=5.0, gmres=(3, 5, 1), extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=0)),
    "box8_2rank_gmres_ilu0": lambda: make_case("box8_2rank_gmres_ilu0", mesh=kuhn_box(8, jitter=0.15), np_ranks=2,
                                               part=slab_part(kuhn_box(8, jitter=0.15)[0], 2), nsgs=3, cfl=5.0, gmres=(3, 6, 2)),
    # general elements: the reference's median-dual metrics (Mesh::CalcAreasVolumes, mesh.tcc:1653-2218) and its element list
    # in its own winding for boxes of hexes, of prisms, of pyramids, and one with all four volume element types and both
    # boundary face types (boxmesh.mixed_box)
    **{f"elem_{kind}": (lambda kind=kind: make_case(
        f"elem_{kind}", ugrid=lambda path: write_ugrid_general(path, *mixed_box(4, kind, jitter=0.12)), elements=True))
       for kind in ("hex", "prism", "pyramid", "mixed")},
    # Mesh::ReorderMeshCuthillMcKee (mesh.tcc:2412-2494): the permutation only (rcm_ordering), plain and reversed, on a Kuhn
    # box, on the pyramid box (valences 8 .. 26: fronts beyond std::sort's insertion-sort threshold) and on a partition
    "rcm_box6": lambda: make_case("rcm_box6", mesh=kuhn_box(6, jitter=0.15), rcm=2),
    "rcm_pyramid": lambda: make_case("rcm_pyramid", ugrid=lambda path: write_ugrid_general(path, *mixed_box(5, "pyramid", jitter=0.1)), rcm=1),
    "rcm_2rank": lambda: make_case("rcm_2rank", mesh=kuhn_box(6, jitter=0.15), np_ranks=2,
                                   part=slab_part(kuhn_box(6, jitter=0.15)[0], 2), rcm=2, elements=True),
    # the solver's default start-up (reorderMesh = 1, solutionSpace.tcc:61-74): reverse Cuthill-McKee, ReorderC2nMap (which
    # takes ordering[] as new-of-old), maps and metrics rebuilt on the renumbered mesh
    "elem_mixed_rcm": lambda: make_case(
        "elem_mixed_rcm", ugrid=lambda path: write_ugrid_general(path, *mixed_box(4, "mixed", jitter=0.12)), elements=True, reorder=1),
    # the all-types box cut into two y-slabs (through prisms, hexes, pyramids and tets): what udecomp writes per rank
    "elem_mixed_2rank": lambda: make_case(
        "elem_mixed_2rank", ugrid=lambda path: write_ugrid_general(path, *mixed_box(4, "mixed", jitter=0.12)), elements=True,
        np_ranks=2, part=slab_part(mixed_box(4, "mixed", jitter=0.12)[0], 2, axis=1)),
    # complex-step field Jacobians (jacobianFieldType = 2: Kernel_NumJac_Complex, jacobian.tcc:370-433) with the one-sided
    # boundary Jacobian
    "box6_implicit_complex": lambda: make_case("box6_implicit_complex", mesh=kuhn_box(6, jitter=0.15), nsgs=3, cfl=5.0,
                                               jactype=2, jactype_b=0),
    # CRSMatrix::CRSTranspose (crsmatrix.tcc:568-599) of the assembled Jacobian: blocks transposed in place, local mirror
    # blocks swapped, ghost-column blocks replaced by the owner's through PObj::TransposeCommCRS (parallel.tcc:54-338) --
    # one rank for both block sizes, two slabs, four quadrant columns (every rank has three neighbours)
    "box5_transpose": lambda: make_case("box5_transpose", mesh=kuhn_box(5, jitter=0.15), nsgs=1, cfl=5.0, transpose=True),
    "box4_fr_transpose": lambda: make_case("box4_fr_transpose", mesh=kuhn_box(4, jitter=0.15), eqnset="compressibleEulerFR",
                                           nsgs=1, cfl=5.0, transpose=True, extra=FR_EXTRA.format(temp=3000, pres=101325, rxn=1)),
    "box6_2rank_transpose": lambda: make_case("box6_2rank_transpose", mesh=kuhn_box(6, jitter=0.15), np_ranks=2,
                                              part=slab_part(kuhn_box(6, jitter=0.15)[0], 2), nsgs=1, cfl=5.0, transpose=True),
    "box6_4rank_transpose": lambda: make_case("box6_4rank_transpose", mesh=kuhn_box(6, jitter=0.15), np_ranks=4,
                                              part=quadrant_part(kuhn_box(6, jitter=0.15)[0]), nsgs=1, cfl=5.0, transpose=True),
    # Forces::Compute / ComputeSurfaceAreas (forces.tcc): pressure and viscous forces, moments, cp / y+ / cf per
    # half-edge, lift / drag / moment coefficients of two composite bodies (the no-slip floor; three far-field faces)
    "box6_ns_forces": lambda: make_case("box6_ns_forces", mesh=kuhn_box(6, jitter=0.15),
                                        bc=ns_bc(330.0) + "\nbody #1 = [5]\nbody #2 = [1,2,6]\n", eqnset="compressibleNS",
                                        nsgs=3, cfl=5.0, refvisc=0.5, forces=True,
                                        extra="liftDirection = [0.3, 0.2, -1.0]\ndragDirection = [1.0, 0.1, 0.0]\n"),
    "box4_nsfr_forces": lambda: make_case("box4_nsfr_forces", mesh=kuhn_box(4, jitter=0.15),
                                          bc=ns_bc(900.0) + "\nbody #1 = [5]\nbody #2 = [1,2,6]\n", eqnset="compressibleNSFR",
                                          nsgs=3, cfl=5.0, refvisc=2.0e-4, forces=True,
                                          extra=FR_EXTRA.format(temp=950, pres=2000, rxn=1)
                                          + "refThermalConductivity = 0.05\nrefLength = 0.01\n"
                                          + "liftDirection = [0.3, 0.2, -1.0]\ndragDirection = [1.0, 0.1, 0.0]\n"),
    # the reference's own unit-test fixture (unitTest/gradientTest.h:20-232): prism cube, 216 nodes
    "cube_LowFi": lambda: make_case(
        "cube_LowFi", h5=os.path.join(REFERENCE, "unitTest/meshResources/cubeStructuredSeries/cube_LowFi.0.h5"),
        bc="".join(f"surface #{i} = farField\n" for i in range(1, 27)), nsgs=2, cfl=5.0, elements=True),
}


def main():
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()


if __name__ == "__main__":
    main()

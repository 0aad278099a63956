"""Host emulation of selected CUDA kernels: the kernel SOURCE TEXT of proteuscfd_b200/csrc (the device functions of
eqnset_compressible.cuh, the helpers of pcfd_internal.cuh and the named kernels of pcfd_kernels.cu) is extracted, compiled
for the host with g++ (-ffp-contract=off, the counterpart of nvcc's --fmad=false; `__global__` / `__device__` defined away,
`threadIdx` / `blockIdx` set by a loop) and run against the C oracle.

Why: Green-Gauss gradients and central-difference Jacobians were written after the round's GPU minutes were spent
(tests/test_zz_gpu_pending.py holds their B200 tests).  This test executes the very same kernel source on the CPU, so a
logic error in them cannot wait for the next GPU run to show.  It is TEST INFRASTRUCTURE: nothing here is reachable from
the product, which still has no CPU path.  As a control the same emulation runs two GPU-verified kernels (k_gradient,
k_jac_edges + jac_half_edge) and must reproduce the oracle with them too.

What it cannot check: launch configuration, the half-edge work-list split (bnodes / blist / bfirst -- shared with the
verified kernels), nvcc's code generation.  Bars: bit-exact for the gradient and for the off-diagonal Jacobian blocks;
the diagonal blocks are summed here in numpy in a different order than k_jac_diag does it, so 1e-10 of the block scale.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import oracle_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "proteuscfd_b200", "csrc")

PRELUDE = r"""
#include <cmath>
#include <cstddef>
#include <cstring>
struct int2 { int x, y; };
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::isnan;
using std::isfinite;
#define __noinline__
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, sizeof d); return d; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }   // explicit, exact on both sides
struct idx3 { unsigned x, y, z; };
static idx3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1};
#include "pcfd.h"
#include "eqnset_compressible.cuh"
#define NEQN PCFD_NEQN
#define NVARS PCFD_NVARS
#define NTERMS PCFD_NTERMS
#define NEQN2 (NEQN * NEQN)
"""

DRIVER = r"""
extern "C" {
struct emu_mesh {
  int nnode, gnode, nbnode, nedge, nbedge, ngedge;
  const int* en; const double* ea; const int* ben; const double* bea; const int* bctype; const double* xyz;
  const double* vol; const int* adjp; const int* adj; const int* bnormal; const double* btwall;
};
static DevMesh dev(const emu_mesh* m) {
  DevMesh d;
  d.nnode = m->nnode; d.gnode = m->gnode; d.nbnode = m->nbnode; d.nedge = m->nedge; d.nbedge = m->nbedge; d.ngedge = m->ngedge;
  d.en = (const int2*)m->en; d.ea = m->ea; d.ben = (const int2*)m->ben; d.bea = m->bea; d.bctype = m->bctype; d.xyz = m->xyz;
  d.vol = m->vol; d.adjp = m->adjp; d.adj = (const int2*)m->adj; d.bnormal = m->bnormal; d.btwall = m->btwall;
  return d;
}
#define FOR_THREADS(n) for (blockIdx.x = 0; blockIdx.x < (unsigned)(n); blockIdx.x++)
void emu_gradient(const emu_mesh* m, int gg, const double* q, const double* sw, double* qgrad) {
  DevMesh d = dev(m);
  FOR_THREADS(m->nnode) { if (gg) k_gradient_gg(d, q, qgrad); else k_gradient(d, q, sw, qgrad); }
}
// posLR / posRL: slot e and nedge + e of a per-edge block array
void emu_jac_edges(const emu_mesh* m, int central, double gamma, const double* q, const int* posLR, const int* posRL, double* A) {
  DevMesh d = dev(m);
  FOR_THREADS(m->nedge) { if (central) k_jac_edges_central(d, gamma, q, posLR, posRL, A); else k_jac_edges(d, gamma, q, posLR, posRL, A); }
}
// all half-edges in the reference's order, the interior state read from / written to q directly (Bdriver semantics)
void emu_jac_bedges(const emu_mesh* m, int central, double gamma, int no_cvbc, const double* qinf, double* q, double* bdiag) {
  DevMesh d = dev(m);
  eq::BcParams bp;
  bp.gamma = gamma; bp.no_cvbc = no_cvbc;
  for (int k = 0; k < NVARS; k++) bp.qinf[k] = qinf[k];
  for (int be = 0; be < m->nbedge + m->ngedge; be++) {
    double* QL = q + (size_t)d.ben[be].x * NVARS;
    if (central) jac_half_edge_central(d, bp, be, QL, q, (const int*)0, bdiag, (double*)0);
    else jac_half_edge(d, bp, be, QL, q, (const int*)0, bdiag, (double*)0);
  }
}
// the same with the block positions of A(l, ghost) for the parallel half-edges of a partition
void emu_jac_bedges_pos(const emu_mesh* m, int central, double gamma, int no_cvbc, const double* qinf, double* q, const int* bpos,
                        double* bdiag, double* A) {
  DevMesh d = dev(m);
  eq::BcParams bp;
  bp.gamma = gamma; bp.no_cvbc = no_cvbc;
  for (int k = 0; k < NVARS; k++) bp.qinf[k] = qinf[k];
  for (int be = 0; be < m->nbedge + m->ngedge; be++) {
    double* QL = q + (size_t)d.ben[be].x * NVARS;
    if (central) jac_half_edge_central(d, bp, be, QL, q, bpos, bdiag, A);
    else jac_half_edge(d, bp, be, QL, q, bpos, bdiag, A);
  }
}
}
"""


def extract(src, name):
    """the definition of function / struct `name`: from the start of its line group to the matching closing brace"""
    m = re.search(r"^[^\n/]*\b" + re.escape(name) + r"\s*\(", src, re.M) if not name.startswith("struct ") else \
        re.search(r"^" + re.escape(name) + r"\s*\{", src, re.M)
    assert m, f"{name} not found"
    start = m.start()
    prev = src.rfind("\n", 0, start - 1) + 1
    if src[prev:start].startswith("template"):
        start = prev
    i = src.index("{", m.end() - 1)
    depth = 0
    while True:
        c = src[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    end = i + 1
    if name.startswith("struct "):
        end = src.index(";", end) + 1
    return src[start:end] + "\n"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    work = tmp_path_factory.mktemp("host_emul")
    internal = open(os.path.join(CSRC, "pcfd_internal.cuh")).read()
    kernels = open(os.path.join(CSRC, "pcfd_kernels.cu")).read()
    parts = [PRELUDE]
    for n in ("struct DevMesh", "is_ghost", "load_avec", "lsq_weights"):
        parts.append(extract(internal, n))
    for n in ("load_q5", "load_q10", "store_q10", "k_gradient", "k_gradient_gg", "k_jac_edges", "k_jac_edges_central",
              "jac_half_edge", "jac_half_edge_central"):
        parts.append(extract(kernels, n))
    parts.append(DRIVER)
    cpp = work / "emul.cpp"
    cpp.write_text("".join(parts))
    so = work / "libemul.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    return C.CDLL(str(so))


class EmuMesh(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge")] + \
               [(k, C.c_void_p) for k in ("en", "ea", "ben", "bea", "bctype", "xyz", "vol", "adjp", "adj", "bnormal", "btwall")]


def build_mesh(mesh):
    """emu_mesh + the node -> edge lists the library builds in pcfd_create: per node its edges sorted by edge id (interior
    edges, then half-edges), .x = other node | 1 << 31 if the node is the edge's RIGHT node, .y = edge id"""
    keep = {}
    nnode, nedge = int(mesh["nnode"]), int(mesh["nedge"])
    en = np.ascontiguousarray(mesh["edges_n"], dtype=np.int32).reshape(-1, 2)
    ben = np.ascontiguousarray(mesh["bedges_n"], dtype=np.int32).reshape(-1, 2)
    node = np.concatenate([en[:, 0], en[:, 1], ben[:, 0]]).astype(np.int64)
    other = np.concatenate([en[:, 1].astype(np.int64), en[:, 0].astype(np.int64) | (1 << 31), ben[:, 1].astype(np.int64)])
    eid = np.concatenate([np.arange(nedge), np.arange(nedge), nedge + np.arange(len(ben))]).astype(np.int64)
    order = np.lexsort((eid, node))
    node, other, eid = node[order], other[order], eid[order]
    adjp = np.zeros(nnode + 1, dtype=np.int32)
    np.add.at(adjp, node + 1, 1)
    adjp = np.cumsum(adjp).astype(np.int32)
    adj = np.stack([(other & 0xFFFFFFFF).astype(np.uint32).view(np.int32), eid.astype(np.int32)], axis=1)
    m = EmuMesh()
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        setattr(m, k, int(mesh[k]))
    arrays = dict(en=en, ea=np.ascontiguousarray(mesh["edges_a"], dtype=np.float64), ben=ben,
                  bea=np.ascontiguousarray(mesh["bedges_a"], dtype=np.float64),
                  bctype=np.ascontiguousarray(mesh["bedges_bctype"], dtype=np.int32),
                  xyz=np.ascontiguousarray(mesh["xyz"], dtype=np.float64), vol=np.ascontiguousarray(mesh["vol"], dtype=np.float64),
                  adjp=adjp, adj=np.ascontiguousarray(adj), bnormal=np.full(len(ben), -1, dtype=np.int32),
                  btwall=np.zeros(len(ben)))
    for k, a in arrays.items():
        keep[k] = a
        setattr(m, k, a.ctypes.data)
    return m, keep


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


BC_SETS = {
    "farfield_symmetry_wall": None,     # the default box: far field, symmetry, impermeable wall
    "dirichlet_types": "dirichlet",
}


def case(kind):
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    if kind == "dirichlet":
        bc = {1: capi.BC_SONIC_INFLOW, 2: capi.BC_SONIC_OUTFLOW, 3: capi.BC_SYMMETRY, 4: capi.BC_DIRICHLET,
              5: capi.BC_FARFIELD, 6: capi.BC_NEUMANN}
        return box_case(8, bc=bc, cfl=5.0)
    return box_case(8, cfl=5.0)


@pytest.mark.parametrize("gg", [0, 1])
@pytest.mark.parametrize("kind", list(BC_SETS.values()))
def test_gradient_kernels_on_host(emu, oracle, gg, kind):
    mesh, params, q = case(kind)
    o = oracle_for(oracle, mesh, params)
    o.c.grad_type = gg
    qo = q.copy()
    o.update_bcs(qo, np.zeros(1))
    _, sw = o.lsq()
    ref = o.gradient(qo, sw)
    m, keep = build_mesh(mesh)
    out = np.zeros_like(ref)
    emu.emu_gradient(C.byref(m), gg, _p(qo), _p(sw), _p(out))
    nloc = int(mesh["nnode"]) * 27
    assert np.abs(ref[:nloc]).max() > 0
    assert np.array_equal(out[:nloc], ref[:nloc]), f"max diff {np.abs(out[:nloc] - ref[:nloc]).max():.3e}"


@pytest.mark.parametrize("types", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("kind", list(BC_SETS.values()))
def test_jacobian_kernels_on_host(emu, oracle, types, kind):
    mesh, params, q = case(kind)
    o = oracle_for(oracle, mesh, params)
    o.c.field_jac_type, o.c.boundary_jac_type = types
    ia, ja, iau = o.crs_init()
    qo = q.copy()
    dt, _ = o.timestep(qo, np.zeros(1))
    A = o.jacobian(qo, np.zeros(1), dt, ia, ja, iau).reshape(-1, 25)

    m, keep = build_mesh(mesh)
    nedge, nnode = int(mesh["nedge"]), int(mesh["nnode"])
    nb = int(mesh["nbedge"]) + int(mesh["ngedge"])
    qe = q.copy()
    # field kernel: block slots e (row l, column r) and nedge + e (row r, column l)
    posLR = np.arange(nedge, dtype=np.int32)
    posRL = (nedge + np.arange(nedge)).astype(np.int32)
    E = np.full((2 * nedge, 25), np.nan)
    emu.emu_jac_edges(C.byref(m), types[0], C.c_double(params["gamma"]), _p(qe), _p(posLR), _p(posRL), _p(E))
    # boundary kernel, reference order
    bd = np.full((nb, 25), np.nan)
    qinf = np.ascontiguousarray(params["qinf"], dtype=np.float64)
    emu.emu_jac_bedges(C.byref(m), types[1], C.c_double(params["gamma"]), int(params["no_cvbc"]), _p(qinf), _p(qe), _p(bd))
    assert np.isfinite(E).all() and np.isfinite(bd).all()
    assert np.array_equal(qe, qo), "q after the boundary Jacobian pass (phantom states, aux of the boundary nodes)"

    def block(row, col):
        k = ia[row] + np.nonzero(ja[ia[row]:ia[row + 1]] == col)[0][0]
        return A[k]

    en = keep["en"]
    for e in range(nedge):      # off-diagonal blocks: bit-exact
        l, r = en[e]
        assert np.array_equal(E[e], block(l, r)), f"A(l,r) of edge {e}"
        assert np.array_equal(E[nedge + e], block(r, l)), f"A(r,l) of edge {e}"
    # diagonal blocks: Kernel_Diag_NumJac (-sum of the column's off-diagonal blocks) + boundary terms + V/dt on the diagonal
    D = np.zeros((nnode, 25))
    np.add.at(D, en[:, 0], -E[nedge:])      # dL += -jacL = -A(r,l)
    np.add.at(D, en[:, 1], -E[:nedge])      # dR += -jacR = -A(l,r)
    np.add.at(D, keep["ben"][:, 0], bd)
    vol = keep["vol"]
    for i in range(5):
        D[:, 6 * i] += vol / dt
    ref = A[iau]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(D - ref) <= 1e-10 * scale), f"diagonal blocks off by {np.max(np.abs(D - ref) / scale):.3e} of scale"


# ------------------------------------------------------------------------------------------------ reacting eqnset
FR_DRIVER = r"""
extern "C" {
struct emu_mesh {
  int nnode, gnode, nbnode, nedge, nbedge, ngedge;
  const int* en; const double* ea; const int* ben; const double* bea; const int* bctype; const double* xyz;
  const double* vol; const int* adjp; const int* adj; const int* bnormal; const double* btwall;
};
static DevMesh dev(const emu_mesh* m) {
  DevMesh d;
  d.nnode = m->nnode; d.gnode = m->gnode; d.nbnode = m->nbnode; d.nedge = m->nedge; d.nbedge = m->nbedge; d.ngedge = m->ngedge;
  d.en = (const int2*)m->en; d.ea = m->ea; d.ben = (const int2*)m->ben; d.bea = m->bea; d.bctype = m->bctype; d.xyz = m->xyz;
  d.vol = m->vol; d.adjp = m->adjp; d.adj = (const int2*)m->adj; d.bnormal = m->bnormal; d.btwall = m->btwall;
  return d;
}
// make_params of pcfd_fr.cu, from the same pcfd_fr_params / pcfd_params values
static fr::Params<5> params(const pcfd_fr_params* h, double chi, double cfl, int no_cvbc, int sorder, int limiter) {
  fr::Params<5> p;
  std::memset(&p, 0, sizeof(p));
  for (int i = 0; i < 5; i++) {
    p.mw[i] = h->chem.mw[i];
    p.Rs[i] = chemdev::UNIV_R / h->chem.mw[i];
    for (int r = 0; r < 2; r++) for (int k = 0; k < 7; k++) p.nasa[i][r][k] = h->chem.nasa7[i][r][k];
  }
  p.ref_density = h->ref_density; p.ref_velocity = h->ref_velocity; p.ref_temperature = h->ref_temperature;
  p.ref_pressure = h->ref_pressure; p.ref_time = h->ref_time; p.ref_specific_enthalpy = h->ref_specific_enthalpy;
  p.Pref = h->pref; p.dt = h->dt; p.chi = chi; p.cfl = cfl;
  p.use_local_dt = h->use_local_dt; p.rxn_on = h->rxn_on; p.no_cvbc = no_cvbc; p.sorder = sorder; p.limiter = limiter;
  for (int i = 0; i < 21; i++) p.qinf[i] = h->qinf[i];
  return p;
}
// HLLCFlux as the residual kernels call it (the composition of hllc_side_thermo / hllc_roe_c2 / hllc_assemble), and the same
// flux put together the way kfr_jac_edges does for a perturbed left state: the right state's and -- for a velocity
// perturbation -- every thermodynamic piece taken from the UNPERTURBED evaluation
void emu_fr_flux(const pcfd_fr_params* h, const double* QL, const double* QR, const double* av, double vdotn, double beta,
                 double* flux) {
  fr::Params<5> p = params(h, 0.0, 1.0, 0, 2, 2);
  fr::numerical_flux<5>(p, QL, QR, av, vdotn, beta, flux);
}
void emu_fr_flux_shared(const pcfd_fr_params* h, const double* QL0, const double* QR0, int column, double hstep, const double* av,
                        double beta, double* direct, double* shared) {
  fr::Params<5> p = params(h, 0.0, 1.0, 0, 2, 2);
  double QL[11], QR[11], QP[11];
  for (int k = 0; k < 11; k++) { QL[k] = QL0[k]; QR[k] = QR0[k]; QP[k] = QL0[k]; }
  QP[column] += hstep;
  fr::aux_pr(p, QP);
  fr::numerical_flux<5>(p, QP, QR, av, 0.0, beta, direct);
  double c2L, hrL, c2R, hrR, c2Lp, hrLp;
  fr::hllc_side_thermo(p, QL, c2L, hrL);
  fr::hllc_side_thermo(p, QR, c2R, hrR);
  fr::hllc_side_thermo(p, QP, c2Lp, hrLp);
  const bool thermo = column < 5 || column == 8;      // rho_i or T: new one-state and Roe pieces; u, v, w: the old ones
  const double roe = thermo ? fr::hllc_roe_c2(p, QP, QR) : fr::hllc_roe_c2(p, QL, QR);
  fr::hllc_assemble(p, QP, QR, av, 0.0, beta, thermo ? c2Lp : c2L, thermo ? hrLp : hrL, c2R, hrR, roe, shared);
}
void emu_fr_gradient(const emu_mesh* m, int gg, const double* q, const double* sw, double* qgrad) {
  DevMesh d = dev(m);
  for (blockIdx.x = 0; blockIdx.x < (unsigned)m->nnode; blockIdx.x++) {
    if (gg) kfr_gradient_gg<5>(d, q, qgrad); else kfr_gradient<5>(d, q, sw, qgrad);
  }
}
void emu_fr_jac_edges_central(const emu_mesh* m, const pcfd_fr_params* h, double chi, double cfl, int no_cvbc, const double* q,
                              const double* beta, const int* posLR, const int* posRL, double* A) {
  DevMesh d = dev(m);
  fr::Params<5> p = params(h, chi, cfl, no_cvbc, 2, 2);
  blockDim.x = 2 * 9 * 4;
  for (blockIdx.x = 0; blockIdx.x < (unsigned)((m->nedge + 3) / 4); blockIdx.x++)
    for (threadIdx.x = 0; threadIdx.x < blockDim.x; threadIdx.x++) kfr_jac_edges_central<5, 4>(d, p, q, beta, posLR, posRL, A);
  threadIdx.x = 0; blockDim.x = 1;
}
void emu_fr_jac_bnodes_central(const emu_mesh* m, const pcfd_fr_params* h, double chi, double cfl, int no_cvbc, const int* bnodes,
                               int n, const double* beta, double* q, double* bdiag) {
  DevMesh d = dev(m);
  fr::Params<5> p = params(h, chi, cfl, no_cvbc, 2, 2);
  for (blockIdx.x = 0; blockIdx.x < (unsigned)n; blockIdx.x++)
    kfr_jac_bnodes_central<5>(d, p, bnodes, n, beta, q, (const int*)0, bdiag, (double*)0);
}
void emu_fr_jac_bedges_pos(const emu_mesh* m, const pcfd_fr_params* h, double chi, double cfl, int no_cvbc, int central,
                           const int* list, int n, const unsigned char* bfirst, const double* beta, double* q, const int* bpos,
                           double* bdiag, double* A) {
  DevMesh d = dev(m);
  fr::Params<5> p = params(h, chi, cfl, no_cvbc, 2, 2);
  for (blockIdx.x = 0; blockIdx.x < (unsigned)n; blockIdx.x++) {
    if (central) kfr_jac_bedges_central<5>(d, p, list, n, bfirst, beta, q, bpos, bdiag, A);
    else kfr_jac_bedges<5>(d, p, list, n, bfirst, beta, q, bpos, bdiag, A);
  }
}
void emu_fr_jac_bedges(const emu_mesh* m, const pcfd_fr_params* h, double chi, double cfl, int no_cvbc, int central, const int* list,
                       int n, const unsigned char* bfirst, const double* beta, double* q, double* bdiag) {
  DevMesh d = dev(m);
  fr::Params<5> p = params(h, chi, cfl, no_cvbc, 2, 2);
  for (blockIdx.x = 0; blockIdx.x < (unsigned)n; blockIdx.x++) {
    if (central) kfr_jac_bedges_central<5>(d, p, list, n, bfirst, beta, q, (const int*)0, bdiag, (double*)0);
    else kfr_jac_bedges<5>(d, p, list, n, bfirst, beta, q, (const int*)0, bdiag, (double*)0);
  }
}
}
"""


@pytest.fixture(scope="module")
def emu_fr(tmp_path_factory):
    work = tmp_path_factory.mktemp("host_emul_fr")
    (work / "cuda_runtime.h").write_text("/* stub: the host emulation defines what the kernels use */\n")
    internal = open(os.path.join(CSRC, "pcfd_internal.cuh")).read()
    kernels = open(os.path.join(CSRC, "pcfd_fr.cu")).read()
    prelude = PRELUDE.replace('#include "eqnset_compressible.cuh"', '#include "eqnset_compressible.cuh"\n#include "eqnset_fr.cuh"')
    parts = [prelude]
    for n in ("struct DevMesh", "is_ghost", "load_avec", "lsq_weights"):
        parts.append(extract(internal, n))
    for n in ("struct W", "gradloc", "load_row", "kfr_gradient", "kfr_gradient_gg", "kfr_jac_edges_central", "kfr_jac_bedges",
              "kfr_jac_bedges_central", "kfr_jac_bnodes_central"):
        parts.append(extract(kernels, n))
    parts.append(FR_DRIVER)
    cpp = work / "emul_fr.cpp"
    cpp.write_text("".join(parts))
    so = work / "libemul_fr.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", str(work), "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    return C.CDLL(str(so))


def fr_fixture(name, rxn_on=0):
    from proteuscfd_b200 import capi
    from tests.oracle_lib import chem_tables, load_golden
    g, meta = load_golden(name)
    meta = dict(meta, rxnOn=float(rxn_on))
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    fp = capi.FrParams()
    capi.fill_chem_model(fp.chem, chem_tables(g))
    for k, mk in (("ref_density", "ref_density"), ("ref_velocity", "ref_velocity"), ("ref_temperature", "ref_temperature"),
                  ("ref_pressure", "ref_pressure"), ("ref_time", "ref_time"), ("ref_specific_enthalpy", "ref_specific_enthalpy"),
                  ("pref", "Pref"), ("dt", "dt")):
        setattr(fp, k, float(meta[mk]))
    fp.use_local_dt, fp.rxn_on = int(meta["useLocalTimeStepping"]), int(rxn_on)
    for j, v in enumerate(g["qinf"]):
        fp.qinf[j] = float(v)
    return g, meta, mesh, fp


@pytest.mark.parametrize("name,gg", [("box4_fr_implicit", 0), ("box4_fr_gg", 1)])
def test_fr_gradient_kernels_on_host(emu_fr, name, gg):
    """kfr_gradient (GPU-verified, control) and kfr_gradient_gg against the qgrad the REFERENCE dumped"""
    g, meta, mesh, _ = fr_fixture(name)
    assert int(meta["gradType"]) == gg
    m, keep = build_mesh(mesh)
    q0 = np.ascontiguousarray(g["q0"])
    sw = np.ascontiguousarray(g["lsq_sw"])
    out = np.zeros_like(g["qgrad"])
    emu_fr.emu_fr_gradient(C.byref(m), gg, _p(q0), _p(sw), _p(out))
    nloc = int(meta["nnode"]) * 14 * 3
    assert np.array_equal(out[:nloc], g["qgrad"][:nloc]), f"max diff {np.abs(out[:nloc] - g['qgrad'][:nloc]).max():.3e}"


def test_fr_central_jacobian_kernels_on_host(emu_fr, oracle):
    """kfr_jac_edges_central: off-diagonal blocks bit-exact against the oracle (which is bit-exact on the reference's
    box4_fr_central dump).  kfr_jac_bedges_central: the change of the diagonal blocks between the central and the
    one-sided boundary type must equal the change of the per-node sums of the half-edge blocks, the one-sided ones coming
    from the GPU-verified kfr_jac_bedges run through the same emulation (1e-9 of the block scale: summation order)."""
    from tests.oracle_lib import FrOracle
    g, meta, mesh, fp = fr_fixture("box4_fr_central", rxn_on=0)
    NEQ, NV = 9, 21
    nedge, nnode = int(meta["nedge"]), int(meta["nnode"])
    nb = int(meta["nbedge"]) + int(meta["ngedge"])
    beta = np.ascontiguousarray(g["beta"])
    res = {}
    for btype in (0, 1):
        o = FrOracle(oracle, g, meta)
        o.c.field_jac_type, o.c.boundary_jac_type = 1, btype
        ia, ja, iau = o.crs_init()
        q = g["q_pre"].copy()
        dt, _ = o.timestep(q, beta)
        res[btype] = (o.jacobian(q, beta, dt, ia, ja, iau).reshape(-1, NEQ * NEQ), q)
    A1, q1 = res[1]
    A0, q0 = res[0]

    m, keep = build_mesh(mesh)
    chi, cfl, no_cvbc = float(meta["chi"]), float(meta["cfl"]), int(meta["no_cvbc"])
    posLR = np.arange(nedge, dtype=np.int32)
    posRL = (nedge + np.arange(nedge)).astype(np.int32)
    E = np.full((2 * nedge, NEQ * NEQ), np.nan)
    qe = np.ascontiguousarray(g["q_pre"].copy())
    emu_fr.emu_fr_jac_edges_central(C.byref(m), C.byref(fp), C.c_double(chi), C.c_double(cfl), no_cvbc, _p(qe), _p(beta),
                                    _p(posLR), _p(posRL), _p(E))
    assert np.isfinite(E).all()
    en = keep["en"]

    def block(A, row, col):
        k = ia[row] + np.nonzero(ja[ia[row]:ia[row + 1]] == col)[0][0]
        return A[k]

    for e in range(nedge):
        l, r = en[e]
        assert np.array_equal(E[e], block(A1, l, r)), f"A(l,r) of edge {e}"
        assert np.array_equal(E[nedge + e], block(A1, r, l)), f"A(r,l) of edge {e}"

    # boundary kernels: every half-edge in order; bfirst = first half-edge of its node
    ben = keep["ben"]
    blist = np.arange(nb, dtype=np.int32)
    bfirst = np.zeros(nb, dtype=np.uint8)
    seen = set()
    for be in range(nb):
        if int(ben[be, 0]) not in seen:
            bfirst[be] = 1
            seen.add(int(ben[be, 0]))
    sums = {}
    for central in (0, 1):
        bd = np.full((nb, NEQ * NEQ), np.nan)
        qb = np.ascontiguousarray(g["q_pre"].copy())
        emu_fr.emu_fr_jac_bedges(C.byref(m), C.byref(fp), C.c_double(chi), C.c_double(cfl), no_cvbc, central, _p(blist), nb,
                                 _p(bfirst), _p(beta), _p(qb), _p(bd))
        assert np.isfinite(bd).all()
        assert np.array_equal(qb, q1 if central else q0), "q after the boundary Jacobian pass"
        S = np.zeros((nnode, NEQ * NEQ))
        np.add.at(S, ben[:, 0], bd)
        sums[central] = S
        if central:
            # the node-walk variant (nodes owning a Dirichlet-type half-edge on the GPU; here: every boundary node) must
            # produce the same half-edge blocks and the same state
            bnodes = np.unique(ben[:, 0]).astype(np.int32)
            bd2 = np.full((nb, NEQ * NEQ), np.nan)
            qn = np.ascontiguousarray(g["q_pre"].copy())
            emu_fr.emu_fr_jac_bnodes_central(C.byref(m), C.byref(fp), C.c_double(chi), C.c_double(cfl), no_cvbc, _p(bnodes),
                                             len(bnodes), _p(beta), _p(qn), _p(bd2))
            assert np.array_equal(bd2, bd), "kfr_jac_bnodes_central vs kfr_jac_bedges_central"
            assert np.array_equal(qn, q1)
    dref = A1[iau] - A0[iau]
    demu = sums[1] - sums[0]
    scale = np.maximum(np.abs(A1[iau]).max(axis=1, keepdims=True), np.abs(sums[1]).max(axis=1, keepdims=True))
    assert np.abs(dref).max() > 1e-3 * scale.max(), "the two boundary types do not differ: the test would be vacuous"
    assert np.all(np.abs(demu - dref) <= 1e-9 * scale), f"off by {np.max(np.abs(demu - dref) / scale):.3e} of scale"


# ------------------------------------------------------------------- perfect gas, straight against the reference's dumps
def golden_mesh(name):
    from tests.oracle_lib import load_golden
    g, meta = load_golden(name)
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    return g, meta, mesh


def test_green_gauss_kernel_on_host_vs_reference_dump(emu):
    g, meta, mesh = golden_mesh("box8_explicit_gg")
    assert int(meta["gradType"]) == 1
    m, keep = build_mesh(mesh)
    q0 = np.ascontiguousarray(g["q0"])
    out = np.zeros_like(g["qgrad"])
    emu.emu_gradient(C.byref(m), 1, _p(q0), _p(np.zeros(6)), _p(out))
    nloc = int(meta["nnode"]) * 27
    assert np.array_equal(out[:nloc], g["qgrad"][:nloc])


def test_central_jacobian_kernels_on_host_vs_reference_dump(emu):
    """off-diagonal blocks and the state after the boundary pass bit-exact against the reference's box6_implicit_central
    dump; diagonal blocks to 1e-10 of their scale (numpy sums in another order than k_jac_diag)"""
    g, meta, mesh = golden_mesh("box6_implicit_central")
    assert int(meta["fieldJacType"]) == 1 and int(meta["boundaryJacType"]) == 1
    m, keep = build_mesh(mesh)
    nedge, nnode = int(meta["nedge"]), int(meta["nnode"])
    nb = int(meta["nbedge"]) + int(meta["ngedge"])
    ia, ja, iau = g["ia"], g["ja"], g["iau"]
    A = g["A"].reshape(-1, 25)
    qe = np.ascontiguousarray(g["q0"].copy())
    posLR = np.arange(nedge, dtype=np.int32)
    posRL = (nedge + np.arange(nedge)).astype(np.int32)
    E = np.full((2 * nedge, 25), np.nan)
    gamma = float(meta["gamma"])
    emu.emu_jac_edges(C.byref(m), 1, C.c_double(gamma), _p(qe), _p(posLR), _p(posRL), _p(E))
    bd = np.full((nb, 25), np.nan)
    qinf = np.ascontiguousarray(g["qinf"], dtype=np.float64)
    emu.emu_jac_bedges(C.byref(m), 1, C.c_double(gamma), int(meta["no_cvbc"]), _p(qinf), _p(qe), _p(bd))
    en = keep["en"]
    for e in range(nedge):
        l, r = en[e]
        kLR = ia[l] + np.nonzero(ja[ia[l]:ia[l + 1]] == r)[0][0]
        kRL = ia[r] + np.nonzero(ja[ia[r]:ia[r + 1]] == l)[0][0]
        assert np.array_equal(E[e], A[kLR]) and np.array_equal(E[nedge + e], A[kRL]), f"edge {e}"
    D = np.zeros((nnode, 25))
    np.add.at(D, en[:, 0], -E[nedge:])
    np.add.at(D, en[:, 1], -E[:nedge])
    np.add.at(D, keep["ben"][:, 0], bd)
    for i in range(5):
        D[:, 6 * i] += g["vol"] / g["timestep"]
    ref = A[iau]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(D - ref) <= 1e-10 * scale), f"diagonal blocks off by {np.max(np.abs(D - ref) / scale):.3e} of scale"


# ------------------------------------------------------------------------- a partition: ghost nodes and ghost half-edges
@pytest.mark.parametrize("name", ["box9_3rank_implicit_r1of3", "box8_2rank_explicit_r0of2"])
def test_variants_on_a_partition_on_host(emu, oracle, name):
    """rank-local mesh of a reference multi-rank run: Green-Gauss over the parallel half-edges, and the ghost branch of the
    central boundary Jacobian (A(l, ghost) blocks), against the oracle with the variant switched on"""
    from tests.oracle_lib import Oracle, load_golden
    g, meta = load_golden(name)
    assert int(meta["gnode"]) > 0 and int(meta["ngedge"]) > 0
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    m, keep = build_mesh(mesh)
    nnode, nedge = mesh["nnode"], mesh["nedge"]
    nb = mesh["nbedge"] + mesh["ngedge"]

    o = Oracle(oracle, g, meta)
    o.c.grad_type = 1
    q0 = np.ascontiguousarray(g["q0"].copy())
    ref = o.gradient(q0, np.ascontiguousarray(g["lsq_sw"]))
    out = np.zeros_like(ref)
    emu.emu_gradient(C.byref(m), 1, _p(q0), _p(np.zeros(6)), _p(out))
    assert np.array_equal(out[: nnode * 27], ref[: nnode * 27]), "Green-Gauss on a partition"

    o = Oracle(oracle, g, meta)
    o.c.field_jac_type = o.c.boundary_jac_type = 1
    ia, ja, iau = o.crs_init()
    qo = g["q0"].copy()
    dt, _ = o.timestep(qo, np.zeros(1))
    A = o.jacobian(qo, np.zeros(1), dt, ia, ja, iau).reshape(-1, 25)
    ben = keep["ben"]
    bpos = np.full(nb, -1, dtype=np.int32)
    nghost = 0
    for be in range(nb):
        l, r = int(ben[be, 0]), int(ben[be, 1])
        if nnode <= r < nnode + mesh["gnode"]:
            bpos[be] = ia[l] + np.nonzero(ja[ia[l]:ia[l + 1]] == r)[0][0]
            nghost += 1
    assert nghost > 0
    Ae = np.zeros_like(A)
    bd = np.full((nb, 25), np.nan)
    qe = np.ascontiguousarray(g["q0"].copy())
    qinf = np.ascontiguousarray(g["qinf"], dtype=np.float64)
    emu.emu_jac_bedges_pos(C.byref(m), 1, C.c_double(float(meta["gamma"])), int(meta["no_cvbc"]), _p(qinf), _p(qe), _p(bpos),
                           _p(bd), _p(Ae))
    assert np.array_equal(qe, qo)
    for be in range(nb):
        if bpos[be] >= 0:
            assert np.array_equal(Ae[bpos[be]], A[bpos[be]]), f"A(l, ghost) of half-edge {be}"


@pytest.mark.parametrize("rank", [0, 1])
def test_fr_variants_on_a_partition_on_host(emu_fr, oracle, rank):
    """the reacting eqnset on one of two z-slab partitions (udecomp layout, ghost nodes and parallel half-edges):
    kfr_gradient_gg over the parallel half-edges and the ghost branch of kfr_jac_bedges_central (A(l, ghost) blocks, with
    the GPU-verified one-sided kernel as control) against the FR oracle with the variant switched on"""
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import fr_slab_case
    from tests.test_gpu_fr import fixture_fr_params, oracle_for_fr
    NEQ = 9
    fr, g, meta = fixture_fr_params("box4_fr_implicit", rxn_on=0)
    mesh, params, q, beta = fr_slab_case(4, rank, 2, fr, colored=False)
    assert mesh["gnode"] > 0 and mesh["ngedge"] > 0
    nnode = mesh["nnode"]
    nb = mesh["nbedge"] + mesh["ngedge"]
    m, keep = build_mesh(mesh)
    fp = capi.FrParams()
    capi.fill_chem_model(fp.chem, fr["chem"])
    for k in ("ref_density", "ref_velocity", "ref_temperature", "ref_pressure", "ref_time", "ref_specific_enthalpy", "pref", "dt"):
        setattr(fp, k, float(fr[k]))
    fp.use_local_dt, fp.rxn_on = int(fr.get("use_local_dt", 1)), 0
    for j, v in enumerate(np.asarray(fr["qinf"]).reshape(-1)):
        fp.qinf[j] = float(v)
    beta = np.ascontiguousarray(beta)

    o = oracle_for_fr(oracle, mesh, params, g, meta)
    o.c.grad_type = 1
    q0 = np.ascontiguousarray(q.copy())
    o.update_bcs(q0, beta)
    ref = o.gradient(q0, np.zeros((nnode + mesh["gnode"]) * 6))
    out = np.zeros_like(ref)
    emu_fr.emu_fr_gradient(C.byref(m), 1, _p(q0), _p(np.zeros(6)), _p(out))
    assert np.abs(ref[: nnode * 42]).max() > 0
    assert np.array_equal(out[: nnode * 42], ref[: nnode * 42]), "Green-Gauss (reacting) on a partition"

    ben = keep["ben"]
    blist = np.arange(nb, dtype=np.int32)
    bfirst = np.zeros(nb, dtype=np.uint8)
    seen = set()
    for be in range(nb):
        if int(ben[be, 0]) not in seen:
            bfirst[be] = 1
            seen.add(int(ben[be, 0]))
    for central in (0, 1):
        o = oracle_for_fr(oracle, mesh, params, g, meta)
        o.c.field_jac_type, o.c.boundary_jac_type = 0, central
        ia, ja, iau = o.crs_init()
        qo = q.copy()
        dt, _ = o.timestep(qo, beta)
        A = o.jacobian(qo, beta, dt, ia, ja, iau).reshape(-1, NEQ * NEQ)
        bpos = np.full(nb, -1, dtype=np.int32)
        for be in range(nb):
            l, r = int(ben[be, 0]), int(ben[be, 1])
            if nnode <= r < nnode + mesh["gnode"]:
                bpos[be] = ia[l] + np.nonzero(ja[ia[l]:ia[l + 1]] == r)[0][0]
        assert (bpos >= 0).sum() > 0
        Ae = np.zeros_like(A)
        bd = np.full((nb, NEQ * NEQ), np.nan)
        qe = np.ascontiguousarray(q.copy())
        emu_fr.emu_fr_jac_bedges_pos(C.byref(m), C.byref(fp), C.c_double(params["chi"]), C.c_double(params["cfl"]),
                                     int(params["no_cvbc"]), central, _p(blist), nb, _p(bfirst), _p(beta), _p(qe), _p(bpos),
                                     _p(bd), _p(Ae))
        assert np.array_equal(qe, qo), f"q after the boundary Jacobian pass (central={central})"
        for be in range(nb):
            if bpos[be] >= 0:
                assert np.array_equal(Ae[bpos[be]], A[bpos[be]]), f"A(l, ghost) of half-edge {be} (central={central})"


# ------------------------------------------------------------------------------------------------ surface forces
FORCES_DRIVER = r"""
extern "C" {
void emu_forces(const emu_mesh* m, double gamma, double Re, double Pr, double PrT, double tref, double mach, double V,
                int viscous, const double* q, const double* qgrad, const int* ia, const int* ja, double* props,
                double* terms, double* cp, double* yp, double* cf) {
  DevMesh d = dev(m);
  eq::ViscParams vp{gamma, Re, Pr, PrT, tref, mach};
  FOR_THREADS(m->nbedge) k_surface_props(d, vp, gamma, V, viscous != 0, q, props);
  FOR_THREADS(m->nbedge) k_forces_bedges(d, props, qgrad, NTERMS * 3, 3, viscous ? Re / mach : 1.0, V, viscous != 0, ia, ja,
                                         terms, cp, yp, cf);
}
}
"""


@pytest.fixture(scope="module")
def emu_forces(tmp_path_factory):
    work = tmp_path_factory.mktemp("host_emul_forces")
    internal = open(os.path.join(CSRC, "pcfd_internal.cuh")).read()
    forces = open(os.path.join(CSRC, "pcfd_forces.cuh")).read()
    parts = [PRELUDE]
    for n in ("struct DevMesh", "is_ghost", "load_avec"):
        parts.append(extract(internal, n))
    for n in ("k_surface_props", "stress_vector", "k_forces_bedges"):
        parts.append(extract(forces, n))
    # the emu_mesh / dev() / FOR_THREADS part of the common driver, then the forces entry
    head = DRIVER[: DRIVER.index("void emu_gradient")] + "}\n"
    parts.append(head)
    parts.append(FORCES_DRIVER)
    cpp = work / "emul_forces.cpp"
    cpp.write_text("".join(parts))
    so = work / "libemul_forces.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    return C.CDLL(str(so))


def test_forces_kernels_on_host_vs_reference_dump(emu_forces):
    """k_surface_props + k_forces_bedges (csrc/pcfd_forces.cuh), executed from their source text on the host, against the
    reference's own Forces::Compute (tests/golden/box6_ns_forces.npz).  With glibc's pow behind Sutherland's law on both
    sides: cp, y+, cf bit-exact, and the per-half-edge force terms summed in half-edge order reproduce the reference's body
    sums bit for bit (on the GPU that sum is a tree: tests/test_gpu_forces.py)."""
    from tests.oracle_lib import bodies_from_fixture, load_golden
    g, meta = load_golden("box6_ns_forces")
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    m, keep = build_mesh(mesh)
    nbe = mesh["nbedge"]
    q, qgrad = np.ascontiguousarray(g["forces_q"]), np.ascontiguousarray(g["forces_qgrad"])
    ia, ja = np.ascontiguousarray(g["ia"], dtype=np.int32), np.ascontiguousarray(g["ja"], dtype=np.int32)
    props, terms = np.zeros(4 * nbe), np.zeros(6 * nbe)
    cp, yp, cf = np.zeros(nbe), np.zeros(nbe), np.zeros(nbe)
    V = float(meta["velocity"])
    emu_forces.emu_forces(C.byref(m), C.c_double(meta["gamma"]), C.c_double(meta["Re"]), C.c_double(meta["Pr"]),
                          C.c_double(meta["PrT"]), C.c_double(meta["ref_temperature"]), C.c_double(V), C.c_double(V), 1,
                          _p(q), _p(qgrad), _p(ia), _p(ja), _p(props), _p(terms), _p(cp), _p(yp), _p(cf))
    assert np.array_equal(cp, g["forces_cp"]), "cp"
    assert np.array_equal(yp, g["forces_yp"]), "y+"
    assert np.array_equal(cf, g["forces_cf"]), "cf"
    assert np.abs(yp).max() > 0
    # FORCE_Kernel's sequential += over the half-edges of each body
    offs, tags, mpt, _ = bodies_from_fixture(g)
    ref = g["forces_body"].reshape(-1, 18)
    T = terms.reshape(-1, 6)
    right = keep["ben"][:nbe, 1]
    cg = g["forces_cg"].reshape(-1, 3)
    for b in range(offs.size - 1):
        mine = set(int(t) for t in tags[offs[b]: offs[b + 1]])
        acc = np.zeros(12)
        noslip = g["bedges_bctype"][:nbe] == 4
        for e in range(nbe):
            if int(g["bedges_factag"][e]) not in mine:
                continue
            r = cg[right[e]] - mpt[3 * b: 3 * b + 3]
            for s in range(2):
                if s == 1 and not noslip[e]:
                    continue
                f = T[e, 3 * s: 3 * s + 3]
                acc[3 * s: 3 * s + 3] += f
                acc[6 + 3 * s] += r[1] * f[2] - f[1] * r[2]
                acc[6 + 3 * s + 1] += r[2] * f[0] - f[2] * r[0]
                acc[6 + 3 * s + 2] += r[0] * f[1] - f[0] * r[1]
        assert np.array_equal(acc, ref[b, :12]), f"body {b}: {acc - ref[b, :12]}"


def test_wall_distance_kernel_on_host_vs_reference_field(tmp_path):
    """k_wall_distance (csrc/pcfd_forces.cuh) from its source text on the host against the `wallDistance` field the reference's
    octree search produced for its Spalart-Allmaras fixture (ComputeWallDistOct, ucs/walldist.tcc:116-199): bit-exact."""
    from proteuscfd_b200.walldist import wall_points
    from tests.oracle_lib import load_golden
    forces = open(os.path.join(CSRC, "pcfd_forces.cuh")).read()
    src = PRELUDE + extract(forces, "k_wall_distance") + r"""
extern "C" void emu_wall_distance(int nn, const double* xyz, int npts, const double* pts, double* dist) {
  for (blockIdx.x = 0; blockIdx.x < (unsigned)nn; blockIdx.x++) k_wall_distance(nn, xyz, npts, pts, dist);
}
"""
    cpp = tmp_path / "emul_wd.cpp"
    cpp.write_text(src)
    so = tmp_path / "libemul_wd.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))
    g, meta = load_golden("box6_sa_implicit")
    mesh = {k: g[k] for k in ("bedges_n", "bedges_bctype", "xyz")}
    for k in ("nnode", "gnode", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    pts = wall_points(mesh)
    nn = mesh["nnode"] + mesh["gnode"]
    xyz = np.ascontiguousarray(g["xyz"][: 3 * nn])
    out = np.zeros(nn)
    lib.emu_wall_distance(nn, _p(xyz), int(pts.shape[0]), _p(np.ascontiguousarray(pts)), _p(out))
    assert np.array_equal(out, g["wallDistance"][:nn])
    lib.emu_wall_distance(nn, _p(xyz), 0, _p(np.zeros(3)), _p(out))
    assert np.isinf(out).all()


# ------------------------------------------------------------------------------------------------ Spalart-Allmaras
SA_DRIVER = r"""
extern "C" {
// TurbulenceModel::Compute as pcfd_turb_compute runs it (perfect gas): the kernels in the library's order; the scalar
// Gauss-Seidel sweeps are walked row by row here (for any numbering the level schedule of the library equals that order)
void emu_turb_sa(const emu_mesh* m, double gamma, double Re, double Pr, double PrT, double tref, double mach, int nsgs,
                 const double* q, const double* qgrad, const double* s, const double* dist, const double* dt,
                 const int* ia, const int* ja, const int* iau, const int* posLR, const int* posRL, const int* bpos,
                 const int* tbnodes, int ntb, const int* wnodes, int nw, double* tvar, double* tgrad, double* b, double* A,
                 double* x, double* mut, double* slots, double* bslots) {
  DevMesh d = dev(m);
  eq::ViscParams vp{gamma, Re, Pr, PrT, tref, mach};
  TurbGas g;
  g.Re = Re / mach; g.gstride = NTERMS * 3; g.goff = 3; g.pe = g.pb = g.pn = 0;
  const int nn = m->nnode + m->gnode, nb = m->nbedge + m->ngedge;
  for (int k = 0; k < ia[m->nnode]; k++) A[k] = 0.0;
  for (int k = 0; k < nn; k++) x[k] = 0.0;
  FOR_THREADS(ntb) k_turb_bcs(d, tbnodes, ntb, tvar);
  FOR_THREADS(m->nnode) k_turb_gradient(d, tvar, s, tgrad);
  FOR_THREADS(m->nedge) k_turb_edges<false>(d, vp, g, q, tvar, tgrad, posLR, posRL, slots, A);
  FOR_THREADS(nb) k_turb_bedges<false>(d, vp, g, q, tvar, tgrad, bpos, bslots, A);
  FOR_THREADS(m->nnode) k_turb_node<false>(d, vp, g, q, qgrad, tvar, dist, dt, slots, bslots, iau, posLR, posRL, b, A);
  FOR_THREADS(nw) k_turb_wall(wnodes, nw, ia, iau, b, x, A);
  FOR_THREADS(m->nnode) k_turb_invdiag(m->nnode, iau, A, b);
  for (int sweep = 0; sweep < nsgs; sweep++)
    for (int dir = 0; dir < 2; dir++)
      for (int kk = 0; kk < m->nnode; kk++) {
        const int row = dir ? m->nnode - 1 - kk : kk;
        double rhs = b[row];
        for (int k = ia[row] + 1; k < ia[row + 1]; k++) { const double vout = A[k] * x[ja[k]]; rhs -= vout; }
        x[row] = A[ia[row]] * rhs;
      }
  FOR_THREADS(m->nnode) k_turb_update(m->nnode, x, tvar);
  FOR_THREADS(nn) k_turb_mut<false>(nn, vp, g, q, tvar, mut);
}
}
"""


def test_spalart_allmaras_kernels_on_host_vs_reference_dump(tmp_path):
    """The Spalart-Allmaras kernels of pcfd_kernels.cu (k_turb_*: BCs, unweighted LSQ gradient, convective / diffusive edge
    terms, node assembly with the source term, wall rows, eddy viscosity), executed from their source text on the host in
    the library's order, against the reference's own TurbulenceModel::Compute (tests/golden/box6_sa_implicit.npz).  With
    glibc's exp / pow on both sides tgrad, b, A, x and nu~ are bit-exact, mu_t to the last bit (on the GPU: 1e-12,
    tests/test_gpu_viscous.py)."""
    from tests.oracle_lib import load_golden
    internal = open(os.path.join(CSRC, "pcfd_internal.cuh")).read()
    kernels = open(os.path.join(CSRC, "pcfd_kernels.cu")).read()
    parts = [PRELUDE]
    for n in ("struct DevMesh", "is_ghost", "load_avec", "lsq_weights"):
        parts.append(extract(internal, n))
    for n in ("load_q5", "load_q10"):
        parts.append(extract(kernels, n))
    i0 = kernels.index("namespace sa {")
    parts.append(kernels[i0: kernels.index("}  // namespace sa", i0) + len("}  // namespace sa")] + "\n")
    for n in ("k_turb_bcs", "k_turb_gradient", "struct TurbGas", "k_turb_edges", "k_turb_bedges", "k_turb_node", "k_turb_wall",
              "k_turb_invdiag", "k_turb_update", "k_turb_mut"):
        parts.append(extract(kernels, n))
    parts.append(DRIVER[: DRIVER.index("void emu_gradient")] + "}\n")
    parts.append(SA_DRIVER)
    cpp = tmp_path / "emul_sa.cpp"
    cpp.write_text("".join(parts))
    so = tmp_path / "libemul_sa.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))
    g, meta = load_golden("box6_sa_implicit")
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    m, keep = build_mesh(mesh)
    nnode, nedge, nb = mesh["nnode"], mesh["nedge"], mesh["nbedge"] + mesh["ngedge"]
    ia, ja, iau = (np.ascontiguousarray(g[k], dtype=np.int32) for k in ("ia", "ja", "iau"))

    def pos(row, col):
        return ia[row] + int(np.nonzero(ja[ia[row]: ia[row + 1]] == col)[0][0])
    en, ben = keep["en"], keep["ben"]
    posLR = np.array([pos(l, r) for l, r in en], dtype=np.int32)
    posRL = np.array([pos(r, l) for l, r in en], dtype=np.int32)
    bpos = np.full(nb, -1, dtype=np.int32)
    bct = g["bedges_bctype"][:nb]
    tbnodes = np.unique(ben[bct != 0, 0]).astype(np.int32)
    wnodes = np.unique(ben[bct == 4, 0]).astype(np.int32)
    tvar = g["turb_tvar0"].copy()
    nn = nnode + mesh["gnode"]
    tgrad, b, A, x, mut = np.zeros(nn * 3), np.zeros(nnode), np.zeros(ja.size), np.zeros(nn), np.zeros(nn)
    slots, bslots = np.zeros(nedge * 3), np.zeros(nb * 4)
    lib.emu_turb_sa(C.byref(m), C.c_double(meta["gamma"]), C.c_double(meta["Re"]), C.c_double(meta["Pr"]), C.c_double(meta["PrT"]),
                    C.c_double(meta["ref_temperature"]), C.c_double(meta["velocity"]), int(meta["nSgs"]),
                    _p(np.ascontiguousarray(g["turb_q"])), _p(np.ascontiguousarray(g["turb_qgrad"])),
                    _p(np.ascontiguousarray(g["lsq_s"])), _p(np.ascontiguousarray(g["wallDistance"])),
                    _p(np.ascontiguousarray(g["turb_dt"])), _p(ia), _p(ja), _p(iau), _p(posLR), _p(posRL), _p(bpos),
                    _p(tbnodes), int(tbnodes.size), _p(wnodes), int(wnodes.size), _p(tvar), _p(tgrad), _p(b), _p(A), _p(x),
                    _p(mut), _p(slots), _p(bslots))
    for name, arr in (("tgrad", tgrad), ("b", b), ("A", A), ("x", x)):
        ref = g["turb_" + name]
        assert np.array_equal(arr[: ref.size], ref), f"turb {name}: max diff {np.abs(arr[: ref.size] - ref).max():.3e}"
    # the library's Sutherland law forms T^1.5 as T * sqrt(T) (eq::pow15), the reference calls pow(T, 1.5): the last bit
    # of the molecular viscosity may differ, and with it the last bit of mu_t
    ref = g["turb_mut"]
    print("mu_t entries that differ in the last bit:", int((mut[: ref.size] != ref).sum()), "of", ref.size)
    assert np.all(np.abs(mut[: ref.size] - ref) <= 1.0e-14 * np.abs(ref))      # fv1 = chi^3 / (chi^3 + cv1^3) amplifies it ~4x
    assert np.array_equal(tvar, g["turb_tvar1"])
    assert np.abs(g["turb_x"]).max() > 0


def test_fr_hllc_pieces_on_host_vs_oracle(emu_fr, oracle):
    """fr::numerical_flux -- since round 2 the composition of hllc_side_thermo / hllc_roe_c2 / hllc_assemble -- against the
    oracle's HLLC (bit-exact) on every interior edge of the reacting fixture, and the flux of a perturbed left state put
    together from SHARED pieces the way kfr_jac_edges does (the right state's pieces, and for a velocity perturbation all
    of them, taken from the unperturbed evaluation) against the direct evaluation: identical bits for all nine columns."""
    from tests.oracle_lib import FrOracle
    g, meta, mesh, fp = fr_fixture("box4_fr_implicit")
    o = FrOracle(oracle, g, meta)
    q = g["q0"].reshape(-1, 21)
    en = g["edges_n"].reshape(-1, 2)
    ea = np.ascontiguousarray(g["edges_a"].reshape(-1, 4))
    oracle.orc_fr_hllc_flux.restype = None
    nchk = 0
    for e in range(0, en.shape[0], 3):
        QL, QR = np.ascontiguousarray(q[en[e, 0]]), np.ascontiguousarray(q[en[e, 1]])
        av = np.ascontiguousarray(ea[e])
        beta = 0.5 * (g["beta"][en[e, 0]] + g["beta"][en[e, 1]])
        f, fo = np.zeros(9), np.zeros(9)
        emu_fr.emu_fr_flux(C.byref(fp), _p(QL), _p(QR), _p(av), C.c_double(0.0), C.c_double(beta), _p(f))
        oracle.orc_fr_hllc_flux(C.byref(o.p), _p(QL), _p(QR), _p(av), C.c_double(0.0), C.c_double(beta), _p(fo))
        assert np.array_equal(f, fo), f"edge {e}: {f - fo}"
        for col in range(9):
            d, s = np.zeros(9), np.zeros(9)
            emu_fr.emu_fr_flux_shared(C.byref(fp), _p(QL), _p(QR), col, C.c_double(1.0e-8), _p(av), C.c_double(beta), _p(d), _p(s))
            assert np.array_equal(d, s), f"edge {e}, column {col}: {d - s}"
        nchk += 1
    assert nchk > 50


@pytest.mark.parametrize("name", ["elem_mixed", "elem_pyramid"])
def test_gradient_kernel_on_general_elements_vs_reference(emu, name):
    """k_gradient on meshes of hexes / prisms / pyramids / tets with quadrilateral boundary faces (node valences from 6 to
    beyond 14), against the REFERENCE's own qgrad of that fixture, bit for bit: the edge-based kernels see the element
    types only through the edge and half-edge lists (B200: tests/test_zzz_gpu_general_elements.py)"""
    from tests.oracle_lib import load_golden
    g, meta = load_golden(name)
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    m, keep = build_mesh(mesh)
    q0 = np.ascontiguousarray(g["q0"])
    sw = np.ascontiguousarray(g["lsq_sw"])
    out = np.zeros_like(g["qgrad"])
    emu.emu_gradient(C.byref(m), 0, _p(q0), _p(sw), _p(out))
    assert np.abs(g["qgrad"]).max() > 0
    assert np.array_equal(out, g["qgrad"]), f"max diff {np.abs(out - g['qgrad']).max():.3e}"

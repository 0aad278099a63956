"""Complex-step field Jacobians (Kernel_NumJac_Complex, ucs/jacobian.tcc:370-433; jacobianFieldType = 2).

Fixture: tests/golden/box6_implicit_complex.npz -- the reference run with jacobianFieldType = 2 and the one-sided boundary
Jacobian.  The C oracle (oracle/pcfd_oracle_cs.c: the Roe flux on C99 complex, the same libgcc / glibc routines
std::complex uses) meets it bit for bit in tests/test_oracle.py.  Here the DEVICE side on the host: the kernel source
text (k_jac_edges_complex, csrc/pcfd_kernels.cu) with the templated flux of csrc/eqnset_compressible_cs.cuh, compiled
with g++ -ffp-contract=off, against the reference's off-diagonal blocks (bit for bit); the cplx arithmetic against C99
complex; and the templated flux instantiated on double
against eq::roe_flux (the hot path's flux) bit for bit -- the template is generated from that text and must stay it.
B200: tests/test_zzz_gpu_complex_step.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import load_golden
from tests.test_host_emulation import CSRC, PRELUDE, ROOT, EmuMesh, _p, build_mesh, extract

DRIVER = r"""
extern "C" {
struct emu_mesh {
  int nnode, gnode, nbnode, nedge, nbedge, ngedge;
  const int* en; const double* ea; const int* ben; const double* bea; const int* bctype; const double* xyz;
  const double* vol; const int* adjp; const int* adj; const int* bnormal; const double* btwall;
};
void emu_jac_edges_complex(const emu_mesh* m, double gamma, const double* q, const int* posLR, const int* posRL, double* A) {
  DevMesh d;
  d.nnode = m->nnode; d.gnode = m->gnode; d.nbnode = m->nbnode; d.nedge = m->nedge; d.nbedge = m->nbedge; d.ngedge = m->ngedge;
  d.en = (const int2*)m->en; d.ea = m->ea;
  for (blockIdx.x = 0; blockIdx.x < (unsigned)m->nedge; blockIdx.x++) k_jac_edges_complex(d, gamma, q, posLR, posRL, A);
}
// the templated flux on double next to the hot path's own
void emu_flux_pair(const double* QL, const double* QR, const double* n, double gamma, double* f_eq, double* f_cs) {
  eq::roe_flux(QL, QR, n, 0.0, gamma, f_eq);
  eqcs::roe_flux<double>(QL, QR, n, 0.0, gamma, f_cs);
}
// one complex flux evaluation: imaginary parts out
void emu_flux_complex(const double* QL, const double* QR, int which, int comp, double h, const double* n, double gamma,
                      double* re, double* im) {
  eqcs::cplx L[5], R[5], f[5];
  for (int j = 0; j < 5; j++) { L[j] = eqcs::cplx(QL[j]); R[j] = eqcs::cplx(QR[j]); }
  (which ? R : L)[comp].im += h;
  eqcs::roe_flux(L, R, n, 0.0, gamma, f);
  for (int j = 0; j < 5; j++) { re[j] = f[j].re; im[j] = f[j].im; }
}
}
"""


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    work = tmp_path_factory.mktemp("cs_emul")
    internal = open(os.path.join(CSRC, "pcfd_internal.cuh")).read()
    kernels = open(os.path.join(CSRC, "pcfd_kernels.cu")).read()
    parts = [PRELUDE, '#include "eqnset_compressible_cs.cuh"\n']
    for n in ("struct DevMesh", "load_avec"):
        parts.append(extract(internal, n))
    for n in ("load_q5", "k_jac_edges_complex"):
        parts.append(extract(kernels, n))
    parts.append(DRIVER)
    cpp = work / "emul.cpp"
    cpp.write_text("".join(parts))
    so = work / "libemul.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp)])
    return C.CDLL(str(so))


def fixture_mesh():
    g, meta = load_golden("box6_implicit_complex")
    assert int(meta["fieldJacType"]) == 2 and int(meta["boundaryJacType"]) == 0
    mesh = {k: g[k] for k in ("edges_n", "edges_a", "bedges_n", "bedges_a", "bedges_bctype", "xyz", "vol")}
    for k in ("nnode", "gnode", "nbnode", "nedge", "nbedge", "ngedge"):
        mesh[k] = int(meta[k])
    return g, meta, mesh


def positions(g, nedge):
    ia, ja = g["ia"], g["ja"]

    def find(row, col):
        for k in range(ia[row], ia[row + 1]):
            if ja[k] == col:
                return k
        raise KeyError((row, col))
    en = g["edges_n"].reshape(-1, 2)[:nedge]
    return (np.array([find(l, r) for l, r in en], dtype=np.int32), np.array([find(r, l) for l, r in en], dtype=np.int32))


def test_templated_flux_on_double_is_the_hot_paths_flux(emu):
    g, meta, mesh = fixture_mesh()
    q = g["q0"].reshape(-1, 10)
    en, ea = g["edges_n"].reshape(-1, 2), g["edges_a"].reshape(-1, 4)
    f1, f2 = np.zeros(5), np.zeros(5)
    for e in range(0, en.shape[0], 7):
        QL, QR, n = np.ascontiguousarray(q[en[e, 0], :5]), np.ascontiguousarray(q[en[e, 1], :5]), np.ascontiguousarray(ea[e])
        emu.emu_flux_pair(_p(QL), _p(QR), _p(n), C.c_double(meta["gamma"]), _p(f1), _p(f2))
        assert np.array_equal(f1, f2) and np.abs(f1).max() > 0


def test_complex_step_kernel_vs_reference_blocks(emu):
    """off-diagonal blocks of the reference's A (written by Kernel_NumJac_Complex alone: the viscous / boundary / diagonal
    passes do not touch them for an Euler run) against the emulated device kernel"""
    g, meta, mesh = fixture_mesh()
    m, keep = build_mesh(mesh)
    nedge = int(meta["nedge"])
    posLR, posRL = positions(g, nedge)
    A = np.zeros_like(g["A"])
    q0 = np.ascontiguousarray(g["q0"])
    emu.emu_jac_edges_complex(C.byref(m), C.c_double(meta["gamma"]), _p(q0), _p(posLR), _p(posRL), _p(A))
    ours, ref = A.reshape(-1, 25), g["A"].reshape(-1, 25)
    pos = np.concatenate([posLR, posRL])
    assert np.abs(ref[pos]).max() > 0
    assert np.array_equal(ours[pos], ref[pos])          # bit for bit: the device's cplx arithmetic is libgcc's / glibc's here


def test_complex_step_is_the_derivative_the_differences_approximate(emu):
    """imag(F(q + ih)) / h against a central difference of the real flux: agreement at the central difference's own
    accuracy (1e-7 of the block), and the real part of the complex evaluation is the unperturbed flux"""
    g, meta, mesh = fixture_mesh()
    q = g["q0"].reshape(-1, 10)
    en, ea = g["edges_n"].reshape(-1, 2), g["edges_a"].reshape(-1, 4)
    e = en.shape[0] // 2
    QL, QR, n = np.ascontiguousarray(q[en[e, 0], :5]), np.ascontiguousarray(q[en[e, 1], :5]), np.ascontiguousarray(ea[e])
    gam = C.c_double(meta["gamma"])
    f0, tmp, re, im = np.zeros(5), np.zeros(5), np.zeros(5), np.zeros(5)
    emu.emu_flux_pair(_p(QL), _p(QR), _p(n), gam, _p(f0), _p(tmp))
    for comp in range(5):
        emu.emu_flux_complex(_p(QL), _p(QR), 1, comp, C.c_double(1e-11), _p(n), gam, _p(re), _p(im))
        assert np.abs(re - f0).max() <= 1e-15 * np.abs(f0).max()
        h = 1e-6
        Qp, Qm = QR.copy(), QR.copy()
        Qp[comp] += h
        Qm[comp] -= h
        fp, fm = np.zeros(5), np.zeros(5)
        emu.emu_flux_pair(_p(QL), _p(Qp), _p(n), gam, _p(fp), _p(tmp))
        emu.emu_flux_pair(_p(QL), _p(Qm), _p(n), gam, _p(fm), _p(tmp))
        cd = (fp - fm) / (2 * h)
        assert np.abs(im / 1e-11 - cd).max() <= 1e-6 * max(np.abs(cd).max(), 1e-3)


@pytest.mark.parametrize("kind", [None, "dirichlet"])
def test_complex_step_kernel_vs_oracle_on_other_states(emu, oracle, kind):
    """the same kernel against the C oracle (field type 2) on the seeded boxes of tests/test_host_emulation.py -- other
    states and BC sets than the reference fixture: off-diagonal blocks bit for bit"""
    from tests.oracle_lib import oracle_for
    from tests.test_host_emulation import case
    mesh, params, q = case(kind)
    o = oracle_for(oracle, mesh, params)
    o.c.field_jac_type, o.c.boundary_jac_type = 2, 0
    ia, ja, iau = o.crs_init()
    qo = q.copy()
    dt, _ = o.timestep(qo, np.zeros(1))
    A = o.jacobian(qo, np.zeros(1), dt, ia, ja, iau).reshape(-1, 25)
    m, keep = build_mesh(mesh)
    nedge = int(mesh["nedge"])
    posLR = np.arange(nedge, dtype=np.int32)
    posRL = (nedge + np.arange(nedge)).astype(np.int32)
    E = np.full((2 * nedge, 25), np.nan)
    qe = q.copy()      # (the oracle's boundary pass hard-sets the Dirichlet nodes of qo AFTER its field pass)
    emu.emu_jac_edges_complex(C.byref(m), C.c_double(params["gamma"]), _p(qe), _p(posLR), _p(posRL), _p(E))
    assert np.isfinite(E).all()
    en = keep["en"]

    def pos(row, col):
        return ia[row] + np.nonzero(ja[ia[row]:ia[row + 1]] == col)[0][0]
    ref = np.array([A[pos(l, r)] for l, r in en] + [A[pos(r, l)] for l, r in en])
    assert np.abs(ref).max() > 0
    assert np.array_equal(E, ref), np.abs(E - ref).max() / np.abs(ref).max()


ARITH_C = r"""
#include <complex.h>
void c_ops(int n, const double* a, const double* b, double* out) {
  for (int i = 0; i < n; i++) {
    double complex x = a[2*i] + a[2*i+1]*_Complex_I, y = b[2*i] + b[2*i+1]*_Complex_I;
    double complex m = x*y, d = x/y, s = csqrt(x), r = 0.5/y;
    double* o = out + 8*i;
    o[0] = creal(m); o[1] = cimag(m); o[2] = creal(d); o[3] = cimag(d); o[4] = creal(s); o[5] = cimag(s); o[6] = creal(r); o[7] = cimag(r);
  }
}
"""
ARITH_CPP = r"""
extern "C" void d_ops(int n, const double* a, const double* b, double* out) {
  for (int i = 0; i < n; i++) {
    eqcs::cplx x(a[2*i], a[2*i+1]), y(b[2*i], b[2*i+1]);
    eqcs::cplx m = x*y, d = x/y, s = eqcs::sqrt(x), r = 0.5/y;
    double* o = out + 8*i;
    o[0] = m.re; o[1] = m.im; o[2] = d.re; o[3] = d.im; o[4] = s.re; o[5] = s.im; o[6] = r.re; o[7] = r.im;
  }
}
"""


def test_device_complex_arithmetic_is_libgccs_in_the_complex_step_regime(tmp_path):
    """eqcs::cplx (product, quotient, real / complex, square root) against C99 `double complex` -- the libgcc / glibc
    routines the reference's std::complex<double> resolves to -- for numbers as the complex step produces them (real part
    of order one, imaginary part zero or ~1e-11 of it): bit for bit"""
    c = tmp_path / "a.c"
    c.write_text(ARITH_C)
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(tmp_path / "a.so"), str(c), "-lm"])
    cpp = tmp_path / "d.cpp"
    cpp.write_text(PRELUDE + '#include "eqnset_compressible_cs.cuh"\n' + ARITH_CPP)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", CSRC,
                           "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "d.so"), str(cpp)])
    la, ld = C.CDLL(str(tmp_path / "a.so")), C.CDLL(str(tmp_path / "d.so"))
    rng = np.random.default_rng(11)
    n = 20000
    a = np.stack([rng.uniform(0.05, 30.0, n), rng.uniform(-1, 1, n) * 1e-11 * rng.uniform(0.01, 30.0, n)], axis=1)
    b = np.stack([rng.uniform(0.05, 30.0, n) * rng.choice([-1.0, 1.0], n), rng.uniform(-1, 1, n) * 1e-11 * rng.uniform(0.01, 30.0, n)], axis=1)
    a[::5, 1] = 0.0          # unperturbed quantities: exactly real
    b[::3, 1] = 0.0
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    oc, od = np.zeros((n, 8)), np.zeros((n, 8))
    la.c_ops(n, _p(a), _p(b), _p(oc))
    ld.d_ops(n, _p(a), _p(b), _p(od))
    assert np.array_equal(oc, od), (np.argwhere(oc != od)[:5], np.abs(oc - od).max())

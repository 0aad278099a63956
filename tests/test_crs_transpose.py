"""CRSMatrix::CRSTranspose (ucs/crsmatrix.tcc:568-599) with PObj::TransposeCommCRS (ucs/parallel.tcc:54-338): the transposed
Jacobian of the adjoint path (Compute_dRdQ_Transpose, ucs/jacobian.tcc:121-127).

Fixtures: the REFERENCE's own CRSTranspose of its assembled Jacobian (A -> A_T) on one rank for both block sizes
(box5_transpose, box4_fr_transpose), on two slabs (box6_2rank_transpose_r*of2) and on four quadrant columns where every
rank has three neighbours (box6_4rank_transpose_r*of4), written by tools/make_golden.py through oracle/_ref/ref_harness.
CPU: (1) the C restatement of the local part + the host routing of the ghost-column blocks (proteuscfd_b200/parallel.py:
TransposeMaps, the product's host logic) against A_T on every rank, bit for bit; (2) the device kernels' source text
(csrc/pcfd_crsmatrix.cuh) run on the host, threads in reversed order, against the same; (3) transposing twice is the
identity.  The B200 run through the C ABI is tests/test_zz_gpu_crs_transpose.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.oracle_lib import _d, _i, load_golden, load_oracle
from tests.test_host_emulation import CSRC, extract

SINGLE = [("box5_transpose", 5), ("box4_fr_transpose", 9)]
MULTI = [("box6_2rank_transpose", 2), ("box6_4rank_transpose", 4)]

PRELUDE = r"""
#include <cstddef>
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
struct idx3 { unsigned x, y, z; };
static idx3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1};
"""

DRIVER = r"""
extern "C" {
#define FOR_THREADS_REV(n) for (long long t_ = (long long)(n) - 1; t_ >= 0 && ((blockIdx.x = (unsigned)t_), true); t_--)
void emu_pairs(int N, int nedge, const int* posLR, const int* posRL, double* A) {
  FOR_THREADS_REV((long long)nedge * N * N) { if (N == 5) k_crs_transpose_pairs<5>(nedge, posLR, posRL, A); else k_crs_transpose_pairs<9>(nedge, posLR, posRL, A); }
}
void emu_inplace(int N, int count, const int* pos, double* A) {
  FOR_THREADS_REV((long long)count * N * N) { if (N == 5) k_crs_transpose_inplace<5>(count, pos, A); else k_crs_transpose_inplace<9>(count, pos, A); }
}
void emu_ghost_blocks(int N, int count, const int* pos, int set, double* A, double* buf) {
  FOR_THREADS_REV((long long)count * N * N) { if (N == 5) k_crs_ghost_blocks<5>(count, pos, set, A, buf); else k_crs_ghost_blocks<9>(count, pos, set, A, buf); }
}
}
"""


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    work = tmp_path_factory.mktemp("crs_transpose_emul")
    src = open(os.path.join(CSRC, "pcfd_crsmatrix.cuh")).read()
    parts = [PRELUDE] + [extract(src, n) for n in ("k_crs_transpose_pairs", "k_crs_transpose_inplace", "k_crs_ghost_blocks")]
    parts.append(DRIVER)
    cpp = os.path.join(work, "crs_emu.cpp")
    with open(cpp, "w") as f:
        f.write("\n".join(parts))
    so = os.path.join(work, "crs_emu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, cpp], check=True)
    return C.CDLL(so)


def positions(g, meta):
    """what pcfd_create keeps: block positions of both directions of every interior edge, of the parallel half-edges"""
    ia, ja = g["ia"], g["ja"]

    def find(row, col):
        for k in range(ia[row], ia[row + 1]):
            if ja[k] == col:
                return k
        raise KeyError((row, col))

    nedge, nb, ng = int(meta["nedge"]), int(meta["nbedge"]), int(meta["ngedge"])
    en = g["edges_n"].reshape(-1, 2)
    posLR = np.array([find(l, r) for l, r in en[:nedge]], dtype=np.int32)
    posRL = np.array([find(r, l) for l, r in en[:nedge]], dtype=np.int32)
    ge = g["bedges_n"].reshape(-1, 2)[nb: nb + ng]
    bpos = np.array([find(l, r) for l, r in ge], dtype=np.int32)
    return posLR, posRL, bpos, ge


def oracle_local(g, meta, neqn):
    A = g["A"].copy()
    load_oracle().orc_crs_transpose_local(int(meta["nnode"]), neqn, _i(g["ia"]), _i(g["ja"]), _d(A))
    return A


def emulated_local(emu, g, meta, neqn):
    posLR, posRL, bpos, _ = positions(g, meta)
    A = g["A"].copy()
    emu.emu_pairs(neqn, posLR.size, _i(posLR), _i(posRL), _d(A))
    iau = np.ascontiguousarray(g["iau"], dtype=np.int32)
    emu.emu_inplace(neqn, int(meta["nnode"]), _i(iau), _d(A))
    if bpos.size:
        emu.emu_inplace(neqn, bpos.size, _i(bpos), _d(A))
    return A


def route(parts, locals_, neqn):
    """the ghost-column blocks of every rank through the product's TransposeMaps"""
    from proteuscfd_b200.parallel import TransposeMaps, transpose_ghost_blocks_local
    maps, blocks, pos = [], [], []
    for r, (g, meta) in enumerate(parts):
        _, _, bpos, ge = positions(g, meta)
        maps.append(TransposeMaps(r, len(parts), int(meta["nnode"]), ge, g["gNodeOwner"], g["gNodeLocalId"]))
        blocks.append(locals_[r].reshape(-1, neqn, neqn)[bpos])
        pos.append(bpos)
    new = transpose_ghost_blocks_local(maps, blocks)
    out = []
    for r in range(len(parts)):
        A = locals_[r].copy().reshape(-1, neqn, neqn)
        A[pos[r]] = new[r]
        out.append(A.reshape(-1))
    return out


@pytest.mark.parametrize("name,neqn", SINGLE)
def test_single_rank_transpose_vs_reference(emu, name, neqn):
    g, meta = load_golden(name)
    assert not np.array_equal(g["A"], g["A_T"])
    A = oracle_local(g, meta, neqn)
    assert np.array_equal(A, g["A_T"])
    assert np.array_equal(emulated_local(emu, g, meta, neqn), g["A_T"])
    # and it is the transpose: block (i, j) of A_T is block (j, i) of A, transposed
    ia, ja = g["ia"], g["ja"]
    B, T = g["A"].reshape(-1, neqn, neqn), g["A_T"].reshape(-1, neqn, neqn)
    i = int(meta["nnode"]) // 2
    for k in range(ia[i], ia[i + 1]):
        j = ja[k]
        m = [q for q in range(ia[j], ia[j + 1]) if ja[q] == i][0]
        assert np.array_equal(T[k], B[m].T)


@pytest.mark.parametrize("name,nranks", MULTI)
@pytest.mark.parametrize("how", ["oracle", "emulated"])
def test_partitioned_transpose_vs_reference(emu, name, nranks, how):
    parts = [load_golden(f"{name}_r{r}of{nranks}") for r in range(nranks)]
    if how == "oracle":
        locals_ = [oracle_local(g, m, 5) for g, m in parts]
    else:
        locals_ = [emulated_local(emu, g, m, 5) for g, m in parts]
    # the local part alone is NOT the reference's result: the ghost-column blocks still hold this rank's own values
    assert any(not np.array_equal(locals_[r], parts[r][0]["A_T"]) for r in range(nranks))
    out = route(parts, locals_, 5)
    for r in range(nranks):
        assert np.array_equal(out[r], parts[r][0]["A_T"]), r
    if nranks == 4:   # every rank of the quadrant cut talks to the three others
        from proteuscfd_b200.parallel import TransposeMaps
        g, meta = parts[0]
        _, _, _, ge = positions(g, meta)
        m = TransposeMaps(0, 4, int(meta["nnode"]), ge, g["gNodeOwner"], g["gNodeLocalId"])
        assert [len(q) > 0 for q in m.requests] == [False, True, True, True]


def test_ghost_block_pack_kernel_round_trip(emu):
    g, meta = load_golden("box6_4rank_transpose_r1of4")
    _, _, bpos, _ = positions(g, meta)
    A = g["A"].copy()
    buf = np.zeros(bpos.size * 25)
    emu.emu_ghost_blocks(5, bpos.size, _i(bpos), 0, _d(A), _d(buf))
    assert np.array_equal(buf.reshape(-1, 25), g["A"].reshape(-1, 25)[bpos])
    A2 = np.zeros_like(A)
    emu.emu_ghost_blocks(5, bpos.size, _i(bpos), 1, _d(A2), _d(buf))
    ref = np.zeros_like(A).reshape(-1, 25)
    ref[bpos] = g["A"].reshape(-1, 25)[bpos]
    assert np.array_equal(A2, ref.reshape(-1))


@pytest.mark.parametrize("name,nranks", MULTI)
def test_transposing_twice_is_the_identity(emu, name, nranks):
    parts = [load_golden(f"{name}_r{r}of{nranks}") for r in range(nranks)]
    once = [(dict(g, A=g["A_T"]), m) for g, m in parts]
    locals_ = [emulated_local(emu, g, m, 5) for g, m in once]
    out = route(once, locals_, 5)
    for r in range(nranks):
        assert np.array_equal(out[r], parts[r][0]["A"])


class HostCtx:
    """stands in for capi.Context in parallel.crs_transpose: the matrix in numpy, the local part by the oracle"""

    def __init__(self, g, meta):
        self.g, self.meta = g, meta
        self.A = g["A"].copy()
        self.bpos = positions(g, meta)[2]

    def crs_transpose(self):
        load_oracle().orc_crs_transpose_local(int(self.meta["nnode"]), 5, _i(self.g["ia"]), _i(self.g["ja"]), _d(self.A))

    def get_ghost_blocks(self):
        return self.A.reshape(-1, 5, 5)[self.bpos].copy()

    def set_ghost_blocks(self, blocks):
        self.A.reshape(-1, 5, 5)[self.bpos] = blocks


@pytest.mark.parametrize("name,nranks", MULTI)
def test_crs_transpose_entry_point_over_a_thread_group(name, nranks):
    """parallel.crs_transpose as the GPU test calls it (one thread per rank, the group's all-gather as the transport)"""
    from proteuscfd_b200.parallel import ThreadGroup, crs_transpose
    parts = [load_golden(f"{name}_r{r}of{nranks}") for r in range(nranks)]
    out = [None] * nranks

    def fn(rank, group):
        g, meta = parts[rank]
        mesh = dict(nnode=int(meta["nnode"]), nbedge=int(meta["nbedge"]), ngedge=int(meta["ngedge"]), bedges_n=g["bedges_n"],
                    gNodeOwner=g["gNodeOwner"], gNodeLocalId=g["gNodeLocalId"])
        ctx = HostCtx(g, meta)
        crs_transpose(ctx, mesh, group)
        out[rank] = ctx.A

    ThreadGroup(nranks).run(fn)
    for r in range(nranks):
        assert np.array_equal(out[r], parts[r][0]["A_T"]), r


@pytest.fixture(scope="module")
def route_bin(tmp_path_factory):
    """oracle/harness/transpose_route_test.cpp + the process-based MPI shim: pcfd::RouteTransposedGhostBlocks of
    include/pcfd_host.hpp (what DropIn::CRSTranspose runs between the two pcfd_crs_ghost_blocks calls), no device needed"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    work = tmp_path_factory.mktemp("route_test")
    exe = os.path.join(work, "transpose_route_test")
    subprocess.run(["g++", "-O1", "-std=c++11", "-I" + os.path.join(root, "oracle", "mpi_shim"), "-I" + os.path.join(root, "include"),
                    os.path.join(root, "oracle", "harness", "transpose_route_test.cpp"),
                    os.path.join(root, "oracle", "mpi_shim", "mpi_shim.cpp"), "-o", exe, "-lpthread"], check=True)
    return exe


@pytest.mark.parametrize("name,nranks", MULTI)
def test_cxx_mpi_routing_of_the_dropin_vs_reference(route_bin, tmp_path, name, nranks):
    """the C++ side of the multi-rank transpose: MPI_Alltoall of the counts, requests and blocks point to point, as
    separate PROCESSES over the MPI shim, on the fixtures' ghost tables -- every rank's matrix equals the reference's A_T"""
    parts = [load_golden(f"{name}_r{r}of{nranks}") for r in range(nranks)]
    locals_, pos = [], []
    for r, (g, meta) in enumerate(parts):
        A = oracle_local(g, meta, 5)
        _, _, bpos, ge = positions(g, meta)
        locals_.append(A)
        pos.append(bpos)
        with open(tmp_path / f"route_in.{r}.bin", "wb") as f:
            np.array([int(meta["nnode"]), bpos.size, 25, int(meta["gnode"])], dtype=np.int32).tofile(f)
            np.ascontiguousarray(ge, dtype=np.int32).tofile(f)
            np.ascontiguousarray(g["gNodeOwner"], dtype=np.int32).tofile(f)
            np.ascontiguousarray(g["gNodeLocalId"], dtype=np.int32).tofile(f)
            np.ascontiguousarray(A.reshape(-1, 25)[bpos]).tofile(f)
    env = dict(os.environ, PCFD_MPI_NP=str(nranks))
    subprocess.run([route_bin, str(tmp_path)], check=True, env=env, timeout=120)
    for r in range(nranks):
        routed = np.fromfile(tmp_path / f"route_out.{r}.bin", dtype=np.float64).reshape(-1, 25)
        A = locals_[r].copy().reshape(-1, 25)
        A[pos[r]] = routed
        assert np.array_equal(A.reshape(-1), parts[r][0]["A_T"]), r

"""The reference's node reordering (Mesh::ReorderMeshCuthillMcKee, ucs/mesh.tcc:2412-2494; ucs.x reorders by default,
solutionSpace.tcc:61-74) restated on the host (proteuscfd_b200/ordering.py: cuthill_mckee), against the permutation the
REFERENCE computed (tests/golden/rcm_*.npz: `rcm_ordering`, dumped by the harness from the unmodified routine) for a Kuhn
box (reversed), a box of pyramids (plain) and both ranks of a partition (ghost neighbours count towards the degree but are
never visited) -- exact.  The reference orders a front by degree only, with the UNSTABLE std::sort: the order of equal
degrees is libstdc++'s, so the restatement carries that algorithm (_std_sort), pinned here against the real std::sort
(compiled on the spot) for fronts far beyond the 16-element insertion-sort threshold the meshes never reach."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from proteuscfd_b200.ordering import _std_sort, cuthill_mckee
from tests.oracle_lib import load_golden


@pytest.mark.parametrize("name,reverse", [("rcm_box6", True), ("rcm_pyramid", False), ("rcm_2rank_r0of2", True), ("rcm_2rank_r1of2", True)])
def test_cuthill_mckee_equals_the_reference(name, reverse):
    g, meta = load_golden(name)
    nn = int(meta["nnode"])
    o = cuthill_mckee(nn, g["ipsp"], g["psp"], reverse=reverse)
    assert np.array_equal(o, g["rcm_ordering"])
    assert sorted(o.tolist()) == list(range(nn))
    if reverse:
        assert o[-1] == 0           # the seed ends up last
    # what it is for: the bandwidth of the local graph under the new numbering shrinks (here: against a random numbering)
    new_of_old = np.empty(nn, dtype=np.int64)
    new_of_old[o] = np.arange(nn)
    ip, ps = g["ipsp"], g["psp"]
    rows = np.repeat(np.arange(nn), np.diff(ip[: nn + 1]))
    cols = ps[: ip[nn]]
    loc = cols < nn
    bw = np.abs(new_of_old[rows[loc]] - new_of_old[cols[loc]]).max()
    rnd = np.random.default_rng(1).permutation(nn)
    assert bw < 0.6 * np.abs(rnd[rows[loc]] - rnd[cols[loc]]).max()


def test_disconnected_graph_is_refused():
    # two triangles that do not touch: the reference would spin in its re-seeding loop (mesh.tcc:2470-2479)
    ipsp = np.array([0, 2, 4, 6, 8, 10, 12])
    psp = np.array([1, 2, 0, 2, 0, 1, 4, 5, 3, 5, 3, 4])
    with pytest.raises(ValueError):
        cuthill_mckee(6, ipsp, psp)


SORT_SRC = r"""
#include <algorithm>
#include <deque>
struct IntInt { int a, b; };
static bool DegreeCompare(IntInt i, IntInt j) { return i.b < j.b; }
extern "C" void sort_by_degree(int n, const int* ids, const int* deg, int* out) {
  std::deque<IntInt> R;
  for (int i = 0; i < n; i++) { IntInt q; q.a = ids[i]; q.b = deg[i]; R.push_back(q); }
  std::sort(R.begin(), R.end(), DegreeCompare);
  for (int i = 0; i < n; i++) out[i] = R[i].a;
}
"""


def test_std_sort_restatement_equals_libstdcxx(tmp_path):
    src = tmp_path / "s.cpp"
    src.write_text(SORT_SRC)
    so = tmp_path / "s.so"
    subprocess.run(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-o", str(so), str(src)], check=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(7)
    ip = C.POINTER(C.c_int)
    beyond = 0
    for n in list(range(1, 70)) + [100, 257, 1000]:
        for distinct in (1, 2, 3, 6, 40):
            deg = rng.integers(0, distinct, size=n).astype(np.int32)
            ids = np.arange(n, dtype=np.int32)
            out = np.zeros(n, dtype=np.int32)
            lib.sort_by_degree(n, ids.ctypes.data_as(ip), deg.ctypes.data_as(ip), out.ctypes.data_as(ip))
            got = _std_sort(deg.tolist(), ids.tolist())
            assert got == out.tolist(), (n, distinct)
            stable = [int(i) for i in np.argsort(deg, kind="stable")]
            beyond += got != stable
    assert beyond > 50      # std::sort really is not stable beyond the threshold: the restatement has to carry it

"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/pcfd.h declares; creating a context without a GPU fails loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from proteuscfd_b200 import capi
    return capi.load_library()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pcfd.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pcfd_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_all_exported(lib):
    from proteuscfd_b200 import capi
    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(capi.SYMBOLS) == syms, "capi.SYMBOLS out of sync with include/pcfd.h"
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pcfd.h but not exported"
    assert lib.pcfd_abi_version() == 9


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from proteuscfd_b200 import capi
    from proteuscfd_b200.cases import box_case
    mesh, params, q = box_case(3)
    with pytest.raises(capi.PcfdError, match="no CUDA device|CUDA"):
        capi.Context(mesh, params)


def test_product_does_not_touch_oracle():
    """The product package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "proteuscfd_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn)).read()
                for needle in ("oracle/", "pcfd_oracle", "oracle_lib", "orc_", "import oracle", "from oracle"):
                    assert needle not in txt, f"{fn} references the oracle ({needle})"

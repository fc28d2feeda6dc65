"""CPU-side checks of the drop-in boundary: the library builds, loads without a GPU, and exports exactly
the entry points include/aigw_b200.h declares.  No compute calls here (those are `-m gpu`)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from aigw_b200 import build as B
    return B.build()


def test_header_functions_exported(so_path):
    hdr = open(os.path.join(ROOT, "include", "aigw_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(aigw_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    L = C.CDLL(so_path)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/aigw_b200.h but not exported"


def test_struct_sizes_match_header():
    import aigw_b200 as A
    assert A.DocResult.itemsize == 32 and A.SseResult.itemsize == 48


def test_no_gpu_means_loud_failure(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import aigw_b200 as A
    with pytest.raises(RuntimeError):
        A.Context(0)


def test_product_does_not_touch_oracle():
    """The package must not import, link or read anything under oracle/."""
    pkg = os.path.join(ROOT, "aigw_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "oracle/" not in txt and "_oracle" not in txt, f

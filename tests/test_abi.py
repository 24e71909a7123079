"""The C-ABI library loads on a CPU-only box and exports every symbol include/monovo_b200.h declares."""
import ctypes as C
import os
import re

from conftest import ROOT
from ros2_mono_vo_b200 import _lib


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "monovo_b200.h")).read()
    return sorted(set(re.findall(r"MVO_API\s+[\w\s\*]+?\b(mvo_\w+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in monovo_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"


def test_struct_layouts():
    assert C.sizeof(_lib.MvoKeypoint) == 28       # the fields of cv::KeyPoint
    assert C.sizeof(_lib.MvoDMatch) == 16         # cv::DMatch
    assert C.sizeof(_lib.MvoFrameResult) == 8 * 4 + 12 * 8


def test_no_cpu_fallback(lib):
    """Without a CUDA device mvo_create must fail loudly (MVO_ERR_CUDA), never silently run on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    cfg = _lib.MvoConfig(0, 640, 480, 1000, 1, 0, 0, None)
    h = C.c_void_p()
    rc = lib.mvo_create(C.byref(h), C.byref(cfg))
    assert rc == _lib.MVO_ERR_CUDA
    assert b"no CPU fallback" in lib.mvo_last_error(None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ros2_mono_vo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src, f

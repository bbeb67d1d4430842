"""CPU: the C-ABI library loads and exports every symbol include/dem_b200.h declares; the
product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    import dem_b200
    lib = ctypes.CDLL(dem_b200.library_path())
    hdr = open(os.path.join(ROOT, "include", "dem_b200.h")).read()
    declared = set(re.findall(r"\b(dem_[a-z_]+)\s*\(", hdr))
    assert declared == set(dem_b200.ABI_SYMBOLS), declared ^ set(dem_b200.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    lib.dem_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.dem_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import dem_b200
    with pytest.raises(dem_b200.DemError):
        dem_b200.Engine(device=0)

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


def test_lmp_b200_driver_is_built_and_refuses_cleanly():
    """the command-line driver exists next to the library; without arguments it prints its usage (exit 2); without a GPU it
    reports the missing device and exits 1 -- there is no CPU path to fall back to"""
    import os, subprocess, dem_b200
    exe = os.path.join(os.path.dirname(dem_b200.library_path()), "lmp_b200")
    assert os.path.exists(exe), "lmp_b200 not built (python -c 'import __graft_entry__ as g; g.build()')"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-in", "/dev/null"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr

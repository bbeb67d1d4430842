#!/usr/bin/env python
"""Builds a variant library next to libdem_b200.so with extra nvcc flags (measurement aid).
usage: tools/build_variant.py <name> "<flags>"   ->  liggghts-inl_b200/libv_<name>.so"""
import os, sys, importlib.util
os.environ["DEM_NVCC_EXTRA"] = sys.argv[2]
here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "liggghts-inl_b200")
spec = importlib.util.spec_from_file_location("dem_build", os.path.join(here, "build.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(m.build(force=True, verbose="-v" in sys.argv, out=os.path.join(m.HERE, "libv_%s.so" % sys.argv[1])))

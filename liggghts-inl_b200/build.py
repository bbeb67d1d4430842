"""Builds libdem_b200.so (hand-written CUDA for sm_100a + C ABI) in-tree with nvcc.
Translation units: dem_engine.cu (host orchestration + particle kernels), dem_mesh.cu (triangle-mesh kernels,
--fmad=false so that its geometric predicates round like the reference's C++) and dem_deck.cpp (input-script front end,
host only, on top of the public C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = [("dem_engine.cu", []), ("dem_mesh.cu", ["--fmad=false"]), ("dem_deck.cpp", [])]
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "dem_b200.h")]
OUT = os.path.join(HERE, "libdem_b200.so")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-diag-suppress", "550"]
EXTRA = os.environ.get("DEM_NVCC_EXTRA", "").split()


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdem_b200.so")
    if verbose:
        print(r.stderr)


def build(force=False, verbose=False, out=OUT):
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    os.makedirs(OBJDIR, exist_ok=True)
    import time
    t0 = time.time()  # (a source edited while the compilers run must make the next call build again: the library is stamped with the start time)
    procs, objs = [], []
    for src, extra in UNITS:
        obj = os.path.join(OBJDIR, os.path.basename(out) + "." + src.replace(".cu", ".o").replace(".cpp", ".o"))
        objs.append(obj)
        cmd = [NVCC] + FLAGS + EXTRA + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, p in procs:
        so, se = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(so + se)
            raise RuntimeError("nvcc failed building libdem_b200.so")
        if verbose:
            print(se)
    _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-o", out] + objs, verbose)
    os.utime(out, (t0, t0))
    if out == OUT:  # command-line driver `lmp_b200 -in deck` (the reference's lmp_<machine>), linked against the library next to it
        _run(["g++", "-O2", "-std=c++17", "-o", os.path.join(HERE, "lmp_b200"), os.path.join(CSRC, "lmp_b200_main.cpp"),
              "-L" + HERE, "-ldem_b200", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))

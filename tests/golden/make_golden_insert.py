"""Golden vectors of the insertion decks: tests/golden/in.insert_pack_* run by the UNMODIFIED reference
(oracle/_ref, fix insert/pack + particletemplate/sphere + particledistribution/discrete), state of all particles by tag at the
checkpoints.  Run in the build container only (needs /root/reference compiled by oracle/Makefile.ref):
    python tests/golden/make_golden_insert.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_driver  # noqa: E402
import cases  # noqa: E402

def one(name, cps):
    r = ref_driver.Ref()
    r.cmd(open(os.path.join(HERE, "in." + name)).read())
    out = {}
    for cp in cps:
        r.cmd("run %d upto" % cp)
        a = r.atoms()
        for k in ("tag", "type", "x", "v", "omega", "radius", "rmass", "f", "torque"):
            out["s%d_%s" % (cp, k)] = a[k]
        out["s%d_nbuilds" % cp] = np.int64(r.neigh_builds)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {cp: len(out["s%d_tag" % cp]) for cp in cps}, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1], cases.INSERT_DECKS[sys.argv[1]])
    else:  # (the reference registers its styles in static tables: one instance per process)
        import subprocess
        for name in cases.INSERT_DECKS:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)

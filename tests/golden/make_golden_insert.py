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


TUTORIAL = "/root/reference/examples/LIGGGHTS/INL_tutorials/t01a_static_angle_of_repose_monosphere"


def tutorial_text():
    """the reference's own tutorial deck t01a (read where it lies, never copied into the repo) with its two long runs cut to
    300 and 200 steps -- everything else, the insertion included, unchanged"""
    text = open(os.path.join(TUTORIAL, "in.staticAOR_MonoSphere")).read()
    a, b = "run\t\t1000000 upto", "run\t\t9000000"
    assert a in text and b in text
    return text.replace(a, "run\t\t300 upto").replace(b, "run\t\t200")


def tutorial():
    os.chdir(TUTORIAL)  # (the deck names its STL file relative to itself)
    r = ref_driver.Ref()
    out = {}
    text = tutorial_text().replace("&\n", " ")
    head, tail = text.split("unfix", 1)
    for cp, part in ((1, head), (500, "unfix" + tail)):
        r.cmd("\n".join(l for l in part.splitlines() if not l.strip().startswith("dump")))  # (no output files into the read-only tree)
        a = r.atoms()
        out["s%d_n" % cp] = np.int64(len(a["tag"]))
        for k in ("x", "v", "f", "radius", "rmass"):
            out["s%d_%s" % (cp, k)] = a[k][::16]
            out["s%d_sum_%s" % (cp, k)] = a[k].sum(axis=0)
    np.savez_compressed(os.path.join(HERE, "tutorial_t01a.npz"), **out)
    print("tutorial_t01a", int(out["s1_n"]), flush=True)


def example(rel):
    os.chdir(os.path.dirname(os.path.join(cases.INL_EXAMPLES, rel)))  # (mesh files are named relative to the deck; dump lines are dropped)
    r = ref_driver.Ref()
    r.cmd(cases.example_deck_text(rel, cases.ALL_EXAMPLE_DECKS[rel]))
    a = r.atoms()
    path = os.path.join(HERE, "inl_examples.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    key = rel.replace("/", "|")
    for k in ("tag", "x", "v", "omega", "f", "torque", "radius", "rmass"):
        out[key + ":" + k] = a[k]
    np.savez_compressed(path, **out)
    print(rel, len(a["tag"]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tutorial":
        tutorial()
    elif len(sys.argv) > 2 and sys.argv[1] == "example":
        example(sys.argv[2])
    elif len(sys.argv) > 1:
        one(sys.argv[1], cases.INSERT_DECKS[sys.argv[1]])
    else:  # (the reference registers its styles in static tables: one instance per process)
        import subprocess
        for name in list(cases.INSERT_DECKS) + ["tutorial"]:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)
        for rel in cases.ALL_EXAMPLE_DECKS:
            subprocess.run([sys.executable, os.path.abspath(__file__), "example", rel], check=True)

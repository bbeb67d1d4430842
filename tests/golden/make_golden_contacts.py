"""tests/golden/make_golden_contacts.py -- per-contact output of the UNMODIFIED reference (`compute pair/gran/local id force
torque`, read back through `dump local` with %.17g) for seeded cases of tests/cases.py: tests/golden/contacts_<case>.npz.
Each file holds, for the setup evaluation after S steps (`run S` then `run 0`), the rows (id1, id2, force on id1, torque on
id1).  Build container only:  make -C oracle ref && python tests/golden/make_golden_contacts.py"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import cases  # noqa: E402
import ref_driver  # noqa: E402

CASES = {"box_hertz_cdt": 400, "periodic_epsd2": 2500, "poly_hooke_epsd_cyl": 2500}


def run(name, S):
    c = cases.make_case(name)
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
    open(os.path.join(tmp, "case.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    r.cmd("compute cpl all pair/gran/local id force torque")
    cols = " ".join("c_cpl[%d]" % k for k in range(1, 10))
    r.cmd("dump dl all local %d %s/cpl.*.dump %s" % (S, tmp, cols))
    r.cmd('dump_modify dl format "%s"' % " ".join(["%.17g"] * 9))
    r.cmd("run %d" % S)
    r.cmd("run 0")
    rows = []
    lines = open(os.path.join(tmp, "cpl.%d.dump" % S)).read().splitlines()
    k = lines.index([l for l in lines if l.startswith("ITEM: ENTRIES")][0])
    for l in lines[k + 1:]:
        rows.append([float(v) for v in l.split()])
    a = np.asarray(rows).reshape(-1, 9)
    r.close()
    np.savez_compressed(os.path.join(HERE, "contacts_%s.npz" % name), steps=np.array(S), id1=a[:, 0].astype(np.int32), id2=a[:, 1].astype(np.int32),
                        periodic=a[:, 2].astype(np.int32), force=a[:, 3:6], torque=a[:, 6:9])
    print(name, "rows", len(a))


if __name__ == "__main__":
    if len(sys.argv) == 2:
        run(sys.argv[1], CASES[sys.argv[1]])
    else:
        import subprocess
        for name in CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)

"""tests/golden/make_golden_dump.py -- the reference's `dump custom` text and thermo lines for one seeded case
(tests/golden/dump_box_hertz_cdt.txt, tests/golden/thermo_box_hertz_cdt.txt):
the UNMODIFIED reference (oracle/_ref) runs the deck of cases.make_case("box_hertz_cdt") with
    dump d1 all custom 100 <file> id type x y z vx vy vz fx fy fz omegax omegay omegaz radius
    dump_modify d1 sort id
for 250 steps (snapshots at 0, 100, 200).  Run in the build container only; the fixture is committed."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import cases  # noqa: E402
import ref_driver  # noqa: E402

FIELDS = "id type x y z vx vy vz fx fy fz omegax omegay omegaz radius"


def main():
    c = cases.make_case("box_hertz_cdt")
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
    open(os.path.join(tmp, "case.data"), "w").write(data)
    out = os.path.join(tmp, "dump.txt")
    r = ref_driver.Ref()
    r.cmd(deck)
    r.cmd("dump d1 all custom 100 %s %s" % (out, FIELDS))
    r.cmd("dump_modify d1 sort id")
    # thermo lines of the same run (thermo.cpp): step atoms ke erotate, every 100 steps and on the last step
    log = os.path.join(tmp, "log.txt")
    r.cmd("thermo_style custom step atoms ke erotate")
    r.cmd("thermo 100")
    r.cmd("log " + log)
    r.cmd("run 250")
    r.cmd("log none")
    r.close()
    lines = open(log).read().splitlines()
    k0 = [k for k, ln in enumerate(lines) if ln.split()[:2] == ["Step", "Atoms"]][0]
    block = []
    for ln in lines[k0:]:
        if ln.startswith("Loop time"):
            break
        block.append(ln)
    open(os.path.join(HERE, "thermo_box_hertz_cdt.txt"), "w").write("\n".join(block) + "\n")
    print("thermo lines", len(block))
    txt = open(out).read()
    open(os.path.join(HERE, "dump_box_hertz_cdt.txt"), "w").write(txt)
    print("snapshots", txt.count("ITEM: TIMESTEP"), "bytes", len(txt))


if __name__ == "__main__":
    main()

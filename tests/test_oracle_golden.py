"""CPU: the oracle restatement (oracle/dem_oracle.c) against golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import cases
import parity


@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_oracle_matches_reference_golden(name):
    c = cases.make_case(name)
    g = parity.golden(name)
    e = cases.apply(c, parity.oracle_engine())
    done = 0
    for cp in cases.GOLDEN_CASES[name]["checkpoints"]:
        cases.apply_late(c, e, cp)
        e.setup()  # every `run` command of the deck starts with Verlet::setup (verlet.cpp:134)
        if done == 0:
            parity.compare_topology(e, c, g)  # mesh cases: active edges / corners as the reference derived them
        e.run(cp - done); done = cp
        snap = cases.snapshot(e, c)
        ref = parity.golden_at(g, cp)
        # trajectories are chaotic: rounding-level differences grow with step count
        tol = parity.tol_for(c, cp)
        errs = parity.compare_snapshot(snap, ref, g["rmass"], tol=tol, label="%s@%d" % (name, cp))
        assert e.stats().nbuilds == int(ref["nbuilds"]) , "rebuild cadence differs at %d" % cp
    e.close()

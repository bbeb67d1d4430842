"""bench_data/make_tile.py -- settles the periodic 16 384-sphere tile that bench.py replicates
16x16 into the 4 194 304-sphere bed of BASELINE.json configs[1].  The settling run itself is done
with the UNMODIFIED reference (oracle/_ref), in the build container:
    python bench_data/make_tile.py [nsteps]
Output: bench_data/tile16k.npz (committed; ~1 MB)."""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import ref_driver  # noqa: E402

SEED = 20261017
NX, NY, NZ = 16, 16, 64
RMIN, RMAX = 0.0015, 0.003
PITCH = 2.05 * RMAX


def tile_case():
    rng = np.random.default_rng(SEED)
    n = NX * NY * NZ
    x = cases.lattice((NX, NY, NZ), PITCH, [0.5 * PITCH, 0.5 * PITCH, 0.55 * PITCH], 0.02 * RMAX, rng)
    radius = rng.uniform(RMIN, RMAX, n)
    L = NX * PITCH
    model = "model hertz tangential history rolling_friction cdt"
    return dict(name="tile16k", lo=[0.0, 0.0, 0.0], hi=[L, NY * PITCH, NZ * PITCH + 0.05], periodic=[1, 1, 0], ntypes=1,
                skin=0.001, dt=1e-5,
                props=[("youngsModulus", "peratomtype", [5e6]), ("poissonsRatio", "peratomtype", [0.45]),
                       ("coefficientRestitution", "peratomtypepair", [0.3]), ("coefficientFriction", "peratomtypepair", [0.5]),
                       ("coefficientRollingFriction", "peratomtypepair", [0.1])],
                pair=model, walls=[("floor", model + " primitive type 1 zplane 0.0")], gravity=(9.81, [0.0, 0.0, -1.0]), freeze=0,
                tag=np.arange(1, n + 1, dtype=np.int32), type=np.ones(n, np.int32), mask=np.ones(n, np.int32), x=x,
                v=np.zeros((n, 3)), omega=np.zeros((n, 3)), radius=radius, density=np.full(n, 2500.0))


if __name__ == "__main__":
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    c = tile_case()
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "tile.data"))
    open(os.path.join(tmp, "tile.data"), "w").write(data)
    r = ref_driver.Ref(log=os.path.join(tmp, "log.liggghts"))
    r.cmd(deck)
    r.cmd("thermo 5000")
    r.cmd("run %d" % nsteps)
    a = r.atoms()
    p = r.pairs()
    ke = 0.5 * (a["rmass"] * (a["v"] ** 2).sum(1)).sum()
    print("settled: zmax=%.4f ke=%.3e pairs=%d contacts=%d" % (a["x"][:, 2].max(), ke, len(p["lo"]), int((p["flag"] != 0).sum())))
    np.savez_compressed(os.path.join(HERE, "tile16k.npz"), x=a["x"], v=a["v"], omega=a["omega"], radius=a["radius"],
                        density=a["density"], lo=np.array(c["lo"]), hi=np.array(c["hi"]), nsteps=np.array(nsteps))
    print(open(os.path.join(tmp, "log.liggghts")).read()[-1500:])

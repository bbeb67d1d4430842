"""GPU, BASELINE.json's full size (configs[1], the bench bed of 4,194,304 spheres): the bed is 16 x 16 periodic replicas of
one settled 16,384-sphere tile, so the full-size run must reproduce, tile by tile, the CPU oracle stepping ONE periodic
tile -- a size-independent property that checks the 4M-particle path (Morton sort, cell grid, ELLPACK lists, history
remap, fused step kernel) against the oracle without the oracle ever touching 4M particles."""
import numpy as np
import pytest
import cases
import parity

pytestmark = pytest.mark.gpu
N0 = 16384


def rel(a, b, floor):
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


def test_full_size_bed_reproduces_oracle_tile_by_tile():
    import bench
    import dem_b200
    tiles, steps = 16, 10
    c = bench.bed_case(tiles, tiles)
    n = len(c["tag"])
    assert n == 4194304
    eng = cases.apply(c, dem_b200.Engine(device=0))
    eng.setup(); eng.run(steps)
    got = {k: eng.download(k) for k in ("x", "v", "f", "omega", "torque")}
    tags = eng.download("tag")
    assert np.array_equal(tags, np.arange(1, n + 1)), "particles lost or duplicated"
    st = eng.stats()
    c1 = bench.bed_case(1, 1)
    ref = cases.apply(c1, parity.oracle_engine())
    ref.setup(); ref.run(steps)
    want = {k: ref.download(k) for k in got}
    rmass = 4.0 * np.pi / 3.0 * c1["radius"] ** 3 * c1["density"]
    # the bed is settled: a particle's net force is the small remainder of contact forces of many times its weight, so the
    # error is measured against max(|f|, m g) (the particle's weight), not against the vanishing net force itself
    # Tolerance 1e-7 of the weight (weight x radius for torques): replicating the tile adds offsets of up to 1.5 m to the coordinates, which rounds every
    # position by up to 2e-16 m; through the Hertz stiffness (~5e2 N/m at these overlaps) that alone moves a contact force by
    # ~1e-13 N ~ 1e-9 of a particle's weight (8e-4 N) before the engine has done anything.
    fl = (rmass * 9.81)[:, None]
    Lx, Ly = c1["hi"][0] - c1["lo"][0], c1["hi"][1] - c1["lo"][1]
    worst = 0.0
    for tile in (0, 1, 17, 100, 255):  # corner, edge and interior replicas
        ix, iy = divmod(tile, tiles)
        s = slice(tile * N0, (tile + 1) * N0)
        # positions: same tile-local coordinates (a particle that left its tile through a periodic face of the unit cell
        # re-enters it there, so compare modulo the tile period)
        dx = got["x"][s] - np.array([ix * Lx, iy * Ly, 0.0]) - want["x"]
        dx[:, 0] -= Lx * np.round(dx[:, 0] / Lx); dx[:, 1] -= Ly * np.round(dx[:, 1] / Ly)
        assert np.abs(dx).max() < 1e-13, "tile %d: positions differ by %.2e" % (tile, np.abs(dx).max())
        for k in ("f", "torque"):
            e = rel(got[k][s], want[k], fl if k == "f" else fl * c1["radius"][:, None])
            worst = max(worst, e)
            assert e < 1e-7, "tile %d: %s differs from the oracle's periodic tile by %.2e" % (tile, k, e)
        # a 1e-13 N force difference integrates to ~1e-14 m/s and (through 0.4 m r^2) ~1e-11 rad/s per step
        for k, tol in (("v", 1e-11), ("omega", 1e-8)):
            e = float(np.abs(got[k][s] - want[k]).max())
            assert e < tol, "tile %d: %s differs by %.2e" % (tile, k, e)
    # bookkeeping at full size: list entries and touching entries are 256 x the unit tile's
    p = ref.pairs()
    lo, hi, flag, _ = parity.unique_pairs(p["lo"], p["hi"], p["flag"], p["hist"])
    assert st.npairs_full == 2 * len(lo) * tiles * tiles, (st.npairs_full, len(lo))
    assert st.ncontacts_full == 2 * int((flag != 0).sum()) * tiles * tiles, (st.ncontacts_full, int((flag != 0).sum()))
    assert st.nbuilds == ref.stats().nbuilds
    print("full size vs oracle tile: worst rel. force/torque error %.2e" % worst)
    eng.close(); ref.close()


def test_full_size_run_is_bit_reproducible():
    import bench
    import dem_b200
    c = bench.bed_case(16, 16)
    out = []
    for rep in range(2):
        eng = cases.apply(c, dem_b200.Engine(device=0))
        eng.setup(); eng.run(30)
        out.append({k: eng.download(k) for k in ("x", "v", "omega", "f", "torque")})
        eng.close()
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k]), "run-to-run difference in " + k

"""GPU: size-independent properties of the DEM step that need no oracle -- they hold for any correct implementation of the
reference's pair loop (newton off: every rank computes both halves of a pair itself, forces equal and opposite)."""
import numpy as np
import pytest
import cases

pytestmark = pytest.mark.gpu


def periodic_gas(n3=(16, 16, 16), seed=5, model="model hertz tangential history rolling_friction epsd2"):
    """dense polydisperse cloud in a fully periodic box, no walls, no gravity: only pair forces act"""
    c = cases.case_box(n3=n3, poly=True, periodic=(1, 1, 1), ntypes=2, model=model, seed=seed, name="gas")
    c["walls"] = []; c["gravity"] = None
    L = c["hi"][0]
    c["lo"] = [0.0, 0.0, 0.0]; c["hi"] = [L, L, n3[2] * 2.05 * 0.003]
    rng = np.random.default_rng(seed)
    c["v"] = rng.uniform(-1.0, 1.0, c["v"].shape)      # collisions from the first steps on
    c["omega"] = rng.uniform(-50.0, 50.0, c["omega"].shape)
    return c


def run(c, steps):
    import dem_b200
    e = cases.apply(c, dem_b200.Engine(device=0))
    e.setup(); e.run(steps)
    out = {k: e.download(k) for k in ("x", "v", "omega", "f", "torque", "rmass", "radius")}
    st = e.stats()
    e.close()
    return out, st


def test_pair_forces_conserve_linear_momentum():
    """sum of m v is constant: every pair force is applied equal and opposite (bit-exact mirror evaluation), so the drift is
    pure summation rounding"""
    c = periodic_gas()
    m = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    p0 = (m[:, None] * c["v"]).sum(0)
    out, st = run(c, 400)
    assert st.ncontacts_full > 1000, "the cloud did not collide"
    p1 = (out["rmass"][:, None] * out["v"]).sum(0)
    scale = np.abs(out["rmass"][:, None] * out["v"]).sum()
    assert np.abs(p1 - p0).max() < 1e-12 * scale, (p0, p1)
    f = out["f"].sum(0)
    assert np.abs(f).max() < 1e-12 * np.abs(out["f"]).sum(), "pair forces do not sum to zero: %s" % f


def test_result_does_not_depend_on_upload_order():
    """the same particles uploaded in another order (same tags) give the same trajectory: storage order is the engine's own
    (Morton sort), summation order follows the cell walk, so the two runs differ by rounding only"""
    c = periodic_gas(n3=(10, 10, 10), seed=9)
    a, _ = run(c, 150)
    perm = np.random.default_rng(1).permutation(len(c["tag"]))
    c2 = dict(c)
    for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density"):
        c2[k] = c[k][perm]
    b, _ = run(c2, 150)   # downloads come back ordered by tag: directly comparable
    for k, tol in (("x", 1e-11), ("v", 1e-8), ("omega", 1e-6)):
        assert np.abs(a[k] - b[k]).max() < tol * max(1.0, np.abs(a[k]).max()), (k, np.abs(a[k] - b[k]).max())


def test_rigid_translation_of_the_whole_system():
    """shifting every coordinate (and the box) by a constant leaves forces unchanged to rounding: only differences enter"""
    c = periodic_gas(n3=(10, 10, 10), seed=11)
    a, _ = run(c, 50)
    c2 = dict(c)
    s = np.array([0.5, -0.25, 0.125])   # exactly representable shifts
    c2["x"] = c["x"] + s; c2["lo"] = list(np.array(c["lo"]) + s); c2["hi"] = list(np.array(c["hi"]) + s)
    b, _ = run(c2, 50)
    scale = np.abs(a["f"]).max()
    assert np.abs(a["f"] - b["f"]).max() < 1e-7 * scale   # coordinates of order 0.5 m round 1e-16 m; Hertz stiffness turns that into ~1e-9 relative
    assert np.abs(a["v"] - b["v"]).max() < 1e-9

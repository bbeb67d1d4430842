"""Multi-GPU parity driver (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi.py
Every rank feeds the same seeded case to its engine; the bricks exchange halos and migrate particles
over NCCL; rank 0 gathers the result and compares it with the single-process CPU oracle."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "liggghts-inl_b200"))
import cases  # noqa: E402
import parity  # noqa: E402
import dem_b200  # noqa: E402


def make_engine(rank, world, local):
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(dem_b200.Engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    e = dem_b200.Engine(device=local, rank=rank, nranks=world, nccl_id=bytes(buf.cpu().numpy().tobytes()))
    if os.environ.get("DEM_TEST_OWNER_LIST"):
        e.option("owner_list", 1)
    return e


def gather_snapshot(eng, c, rank, world):
    snap = cases.snapshot(eng, c)
    snap["tag"] = eng.download("tag")
    out = [None] * world
    dist.all_gather_object(out, snap)
    if rank != 0:
        return None
    return merge_snapshots(out, c)


def merge_snapshots(out, c):
    """per-rank snapshots (each with its owned particles' tags) -> one snapshot ordered like a single-process run"""
    tag = np.concatenate([o["tag"] for o in out]); order = np.argsort(tag, kind="stable")
    assert np.array_equal(tag[order], np.sort(c["tag"])), "particles lost or duplicated across ranks"
    merged = {}
    for k in ("x", "v", "f", "omega", "torque"):
        merged[k] = np.concatenate([o[k] for o in out])[order]
    for k in out[0]:
        if k.startswith("wall_"):
            merged[k] = np.concatenate([o[k] for o in out])[order]
    lo = np.concatenate([o["pair_lo"] for o in out]); hi = np.concatenate([o["pair_hi"] for o in out])
    fl = np.concatenate([o["pair_flag"] for o in out]); hs = np.concatenate([o["pair_hist"] for o in out])
    o2 = np.lexsort((hi, lo))
    merged.update(pair_lo=lo[o2], pair_hi=hi[o2], pair_flag=fl[o2], pair_hist=hs[o2])
    for mid, mtype, nodes in c.get("meshes", []):  # mesh contact rows live with the owning rank: merge, order by (tag, triangle)
        mt = np.concatenate([o["mesh_%s_tag" % mid] for o in out]); mi = np.concatenate([o["mesh_%s_tri" % mid] for o in out])
        mh = np.concatenate([o["mesh_%s_hist" % mid] for o in out])
        o3 = np.lexsort((mi, mt))
        merged["mesh_%s_tag" % mid] = mt[o3]; merged["mesh_%s_tri" % mid] = mi[o3]; merged["mesh_%s_hist" % mid] = mh[o3]
    return merged


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ok = True
    for kw, steps in ((dict(n3=(16, 8, 8), poly=True, periodic=(1, 1, 0), model="model hertz tangential history rolling_friction epsd2", ntypes=2), (0, 1, 10, 300, 900)),
                      (dict(n3=(14, 6, 6), model="model hertz tangential history rolling_friction cdt"), (0, 1, 10, 400)),
                      # bonded spheres across the brick boundary (k_step_bond): bonds form at step 2 between particles of different ranks
                      (dict(n3=(12, 5, 4), model="model hertz tangential history", poly=True, bond=dict(kind="bond/nonlinear")), (0, 1, 2, 3, 10, 100)),
                      # triangle-mesh walls on several GPUs: triangles replicated, mesh contact rows migrate with their particle
                      (dict(mesh="box", n3=(14, 6, 4)), (0, 1, 10, 400, 1200)),
                      (dict(mesh="plate", n3=(12, 6, 4), model="model hertz tangential history rolling_friction epsd"), (0, 1, 10, 400, 1500)),
                      (dict(mesh="drum", n3=(10, 10, 5), move=0.2), (0, 1, 10, 600, 1800))):
        kw = dict(kw)
        mesh = kw.pop("mesh", None)
        if world > 2:  # keep every brick wider than two neighbour cutoffs: stretch the bed along x with the rank count
            kw["n3"] = (kw["n3"][0] * world // 2,) + tuple(kw["n3"][1:])
        c = cases.case_mesh(kind=mesh, name="multi_" + mesh, seed=11, **kw) if mesh else cases.case_box(name="multi", seed=11, **kw)
        # give the particles a drift along x so that they migrate between the bricks
        c["v"][:, 0] += (2.5 if not mesh else 0.8) if "bond" not in kw else 0.3
        rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
        eng = make_engine(rank, world, local)
        eng.box(c["lo"], c["hi"], c["periodic"])
        eng.processors(world, 1, 1)  # slabs along x, the drift direction
        eng = cases.apply(c, eng)
        nl0 = None
        ref = cases.apply(c, parity.oracle_engine()) if rank == 0 else None
        done = 0
        for cp in steps:
            if cp == 400 and not mesh and not kw.get("periodic"):
                # particles added between two runs (dem_insert_particles): a row of newcomers above the bed, spread over all the
                # bricks; every rank is handed the whole set and keeps its own share
                k = 6 * world
                n0 = len(c["tag"])
                L0 = c["hi"][0] - c["lo"][0]
                xs = np.stack([c["lo"][0] + L0 * (np.arange(k) + 0.5) / k, np.full(k, 0.5 * (c["lo"][1] + c["hi"][1])), np.full(k, 0.5 * c["hi"][2])], 1)
                new = dict(tag=np.arange(n0 + 1, n0 + k + 1, dtype=np.int32), type=np.ones(k, np.int32), x=xs, radius=np.full(k, 0.0025), density=np.full(k, 2500.0))
                eng.insert(new["tag"], new["type"], new["x"], new["radius"], new["density"])
                if rank == 0:
                    ref.insert(new["tag"], new["type"], new["x"], new["radius"], new["density"])
                c["tag"] = np.concatenate([c["tag"], new["tag"]])
                rmass = np.concatenate([rmass, 4.0 * np.pi / 3.0 * new["radius"] ** 3 * new["density"]])
            eng.setup(); eng.run(cp - done)
            snap = gather_snapshot(eng, c, rank, world)
            nls = [None] * world
            dist.all_gather_object(nls, int(eng.nlocal))
            nl0 = nl0 or nls
            if rank == 0:
                ref.setup(); ref.run(cp - done)
                tol = 1e-10 if cp <= 10 else (1e-6 if cp <= 400 else (1e-4 if not mesh else 1e-3))
                if "bond" in kw: tol = parity.tol_for(c, cp, gpu=True)
                errs = parity.compare_snapshot(snap, cases.snapshot(ref, c), rmass, tol=tol, label="multi@%d" % cp)
                print("world %d step %4d ok: nlocal per rank %s f err %.2e" % (world, cp, nls, errs["f"]), flush=True)
            done = cp
        if not kw.get("periodic") and "bond" not in kw and rank == 0:
            assert nls != nl0, "no particle migrated between the bricks in the drift case"
        if mesh and rank == 0:
            assert sum(len(snap["mesh_%s_tag" % m[0]]) for m in c["meshes"]) > 0, "no mesh contact was exercised"
        eng.close()
    dist.barrier()
    if rank == 0:
        print("MULTI-GPU PARITY OK" if ok else "FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

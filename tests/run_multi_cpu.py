"""CPU driver of the multi-rank HOST logic (run under torchrun with the gloo backend, no GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/run_multi_cpu.py
Every rank asks the library for its brick (dem_brick_layout: the decomposition, neighbour and ownership code the engine
uses at upload / halo setup, callable without a device), the ranks exchange the results over gloo and check that the
bricks tile the box, that the neighbour relation is mutual and that every particle has exactly one owner; then the
snapshot merge of the multi-GPU parity driver (run_multi.merge_snapshots) is fed per-rank slices of a single-process
oracle run and must reproduce it."""
import os
import sys
import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "liggghts-inl_b200"))
import cases  # noqa: E402
import parity  # noqa: E402
import dem_b200  # noqa: E402
import run_multi  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for periodic, grid in (((1, 1, 0), (world, 1, 1)), ((0, 0, 0), (1, world, 1)), ((1, 0, 0), None)):
        c = cases.case_box(n3=(10, 6, 5), poly=True, periodic=periodic, name="cpu_multi", seed=23,
                           model="model hertz tangential history rolling_friction epsd2", ntypes=2)
        x = c["x"].copy()
        if periodic[0]:
            x[::7, 0] += c["hi"][0] - c["lo"][0]   # some particles start outside the periodic box: ownership is decided on the wrapped position
            x[::11, 0] -= c["hi"][0] - c["lo"][0]
        lay = dem_b200.brick_layout(world, rank, c["lo"], c["hi"], periodic, procgrid=grid, x=x)
        lays = [None] * world
        dist.all_gather_object(lays, lay)
        # bricks tile the box along the decomposed dimension, bit for bit
        d = int(np.argmax(lays[0]["pgrid"]))
        assert int(np.prod(lays[0]["pgrid"])) == world and all(l["pgrid"] == lays[0]["pgrid"] for l in lays)
        order = sorted(range(world), key=lambda r: lays[r]["myloc"][d])
        assert lays[order[0]]["sublo"][d] == c["lo"][d] and lays[order[-1]]["subhi"][d] == c["hi"][d]
        for a, b in zip(order[:-1], order[1:]):
            assert lays[a]["subhi"][d] == lays[b]["sublo"][d], "gap or overlap between bricks %d and %d" % (a, b)
            assert lays[a]["neigh"][2 * d + 1] == b and lays[b]["neigh"][2 * d] == a, "neighbour relation is not mutual"
        if periodic[d]:
            assert lays[order[-1]]["neigh"][2 * d + 1] == order[0] and lays[order[0]]["neigh"][2 * d] == order[-1]
        else:
            assert lays[order[-1]]["neigh"][2 * d + 1] == -1 and lays[order[0]]["neigh"][2 * d] == -1
        owners = np.stack([l["mine"] for l in lays])
        assert (owners.sum(0) == 1).all(), "a particle has %s owners" % sorted(set(owners.sum(0).tolist()))
        # snapshot merge of the parity driver: slice a single-process oracle run by owner, merge, compare
        c["x"] = x
        ref = cases.apply(c, parity.oracle_engine())
        ref.setup(); ref.run(40)
        full = cases.snapshot(ref, c)
        tags = np.sort(c["tag"])
        xw = ref.download("x")
        mine = dem_b200.brick_layout(world, rank, c["lo"], c["hi"], periodic, procgrid=grid, x=xw)["mine"].astype(bool)
        own = set(tags[mine].tolist())
        part = {k: full[k][mine] for k in ("x", "v", "f", "omega", "torque")}
        part["tag"] = tags[mine]
        for k in full:
            if k.startswith("wall_"):
                part[k] = full[k][mine]
        # a pair is reported by every rank that owns one of its particles (the engine keeps owned/ghost pairs on both sides)
        pm = np.array([(int(a) in own) or (int(b) in own) for a, b in zip(full["pair_lo"], full["pair_hi"])], bool)
        for k in ("pair_lo", "pair_hi", "pair_flag", "pair_hist"):
            part[k] = full[k][pm]
        parts = [None] * world
        dist.all_gather_object(parts, part)
        merged = run_multi.merge_snapshots(parts, c)
        rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
        parity.compare_snapshot(merged, full, rmass, tol=0.0 + 1e-300, label="cpu merge")
        ref.close()
        if rank == 0:
            print("world %d periodic %s grid %s ok: owned per rank %s" % (world, periodic, lays[0]["pgrid"], owners.sum(1).tolist()), flush=True)
    dist.barrier()
    if rank == 0:
        print("MULTI-RANK HOST LOGIC OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

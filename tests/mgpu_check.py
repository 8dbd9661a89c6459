"""Multi-GPU (slab) parity check, run under torchrun on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py [particles] [steps]

Every rank runs its slab; rank 0 additionally runs the same scene on ONE GPU and compares the merged multi-rank
state with it after every block of steps (positions / velocities / forces of the owners, vein vertices of their
owners), and checks that blood-cell ownership stays a partition while cells migrate and respawn.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200")
capi = importlib.import_module("simulation-server_b200.capi")
wl = importlib.import_module("simulation-server_b200.workloads")
dd = importlib.import_module("simulation-server_b200.distributed")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc, st, info = wl.long_vein(n)
    # wider lateral spread so that the wall is hit within the run
    st = pkg.make_initial_state(sc, seed=77, xz_half_width=44.0, y_range=(-20.0, -(sc.vein_pos[:, 1].min() * -1 - 40.0)))
    lay = sc.layout()
    uid = dd.broadcast_unique_id(rank)
    planes = dd.slab_boundaries(sc, st, world, lay)
    sim = dd.create_slab_sim(sc, st, rank, world, local, uid, planes)
    ref = None
    if rank == 0:
        ref = capi.Sim(sc, device=local)
        ref.upload_state(st)
        print(f"scene: {info['particles']} particles, planes {planes}", flush=True)
    ok = True
    block = 10
    for done in range(block, steps + 1, block):
        sim.step(block)
        sim.synchronize()
        mine = {k: np.stack(sim.download(w), 1) for k, w in (("pos", capi.PARTICLE_POS), ("vel", capi.PARTICLE_VEL), ("frc", capi.PARTICLE_FRC),
                                                           ("vpos", capi.VEIN_POS))}
        mine["owned"] = sim.ownership()
        mine["counts"] = sim.slab_counts()
        mine["stats"] = sim.stats()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            ref.step(block)
            owned = [g["owned"] for g in gathered]
            msg = f"step {done}: " + " | ".join(f"r{r}: {g['counts']}" for r, g in enumerate(gathered))
            # the N-rank run must reproduce the one-GPU run BIT FOR BIT (order-independent sums everywhere): tolerance 0
            for key, which in (("pos", capi.PARTICLE_POS), ("vel", capi.PARTICLE_VEL), ("frc", capi.PARTICLE_FRC)):
                merged = dd.merge_owned([g[key] for g in gathered], owned, lay)
                single = np.stack(ref.download(which), 1)
                err = np.abs(merged.astype(np.float64) - single).max(axis=1)
                bad = int((err > 0).sum())
                msg += f"  {key}: max {err.max():.2e} differing rows {bad}"
                ok = ok and bad == 0
            # vein vertices: take each from the rank whose slab holds its rest position
            y0 = sc.vein_pos[:, 1]
            vmerged = np.empty_like(gathered[0]["vpos"])
            for r in range(world):
                sel = (y0 >= planes[r + 1]) & (y0 < planes[r])
                vmerged[sel] = gathered[r]["vpos"][sel]
            verr = np.abs(vmerged - np.stack(ref.download(capi.VEIN_POS), 1)).max()
            tele = sum(g["stats"]["teleported_cells"] for g in gathered)
            hits = sum(g["stats"]["vein_hits"] for g in gathered)
            msg += f"  vein max {verr:.2e}  teleported {tele} (single {ref.stats()['teleported_cells']}) vein_hits {hits} (single {ref.stats()['vein_hits']})"
            ok = ok and verr == 0.0
            print(msg, flush=True)
    # owned-only transfers: every rank fills its blood cells' entries of a zeroed array - the sum over ranks is the state;
    # uploading it back (no change) makes the next step refresh the halos first, and the run must go on bit for bit
    parts = {w: np.stack(sim.download_owned(w), 1) for w in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC)}
    for w, arr in parts.items():
        cols = [np.ascontiguousarray(arr[:, k]) for k in range(3)]
        if w == capi.PARTICLE_VEL:   # one array through the pinned (zero-copy) path
            cols = [torch.from_numpy(c).pin_memory().numpy() for c in cols]
        sim.upload_owned(w, *cols)
        sim.synchronize()   # a pinned upload is asynchronous: the arrays must outlive it
    sim.step(block)
    sim.synchronize()
    # (this download through pinned arrays: the kernel writes the host memory in place, walking the owned-cell lists)
    pinned = tuple(torch.zeros(sim.n_particles, dtype=torch.float32).pin_memory().numpy() for _ in range(3))
    after = np.stack(sim.download_owned(capi.PARTICLE_POS, out=pinned), 1)
    gathered = [None] * world
    dist.all_gather_object(gathered, (parts, after))
    if rank == 0:
        for w in parts:
            total = sum(g[0][w].astype(np.float64) for g in gathered)
            bad = int((total != np.stack(ref.download(w), 1)).any(axis=1).sum())
            print(f"owned-only download, array {w}: differing rows {bad}", flush=True)
            ok = ok and bad == 0
        ref.step(block)
        total = sum(g[1].astype(np.float64) for g in gathered)
        bad = int((total != np.stack(ref.download(capi.PARTICLE_POS), 1)).any(axis=1).sum())
        print(f"after owned-only upload + {block} steps: differing rows {bad}", flush=True)
        ok = ok and bad == 0
        print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
        ref.close()
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a rank that raises must take the job down: its peers would otherwise wait in NCCL for ever
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)

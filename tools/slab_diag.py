"""Multi-rank diagnostic (torchrun): the same slab run under several environment switches in ONE process group; every rank
reports whether its handle raised and with what message.  usage: slab_diag.py [particles] [steps]"""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200"); capi = importlib.import_module("simulation-server_b200.capi")
wl = importlib.import_module("simulation-server_b200.workloads"); dd = importlib.import_module("simulation-server_b200.distributed")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc, st, info = wl.long_vein(n)
if os.environ.get("DIAG_WIDE", "1") == "1":
    st = pkg.make_initial_state(sc, seed=77, xz_half_width=44.0, y_range=(-20.0, -(sc.vein_pos[:, 1].min() * -1 - 40.0)))
lay = sc.layout()
planes = dd.slab_boundaries(sc, st, world, lay)
if rank == 0:
    print("planes", planes, "grid min", list(lay.grid_min), "cell", list(sc.cell_size), flush=True)
configs = (("fused", {}), ("nofuse", {"BCS_SLAB_NO_FUSE": "1"}), ("global_rows", {"BCS_ROWS_GLOBAL": "1"}))
if os.environ.get("DIAG_ONLY"):
    configs = tuple(c for c in configs if c[0] in os.environ["DIAG_ONLY"].split(","))
for name, env in configs:
    for k in ("BCS_SLAB_NO_FUSE", "BCS_ROWS_GLOBAL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    uid = dd.broadcast_unique_id(rank)
    kw = {}
    if os.environ.get('DIAG_CAPMIG'):
        kw['migration_capacity'] = int(os.environ['DIAG_CAPMIG'])
    sim = dd.create_slab_sim(sc, st, rank, world, local, uid, planes, **kw)
    msg = "ok"
    for blk in range(steps):
        try:
            sim.step(1)
            sim.synchronize()
        except Exception as e:   # the flag is sticky: report the first step it shows at
            msg = f"step {blk + 1}: {e}"
            break
    # every rank keeps stepping to the same count so that nobody waits in NCCL
    for _ in range(steps - (blk + 1)):
        try:
            sim.step(1)
        except Exception:
            pass
    try:
        cnt = sim.slab_counts()
    except Exception as e:
        cnt = str(e)[:60]
    out = [None] * world
    dist.all_gather_object(out, (msg, cnt))
    if rank == 0:
        for r, o in enumerate(out):
            print(f"[{name}] rank {r}: {o[0]}  {o[1]}", flush=True)
    sim.close()
dist.barrier()
dist.destroy_process_group()

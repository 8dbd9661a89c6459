"""One rank of an N-rank slab decomposition alone on one GPU (BCS_SLAB_NO_COMM=1: no NCCL, no halos): graph-replayed
ms/step of that rank's work.  usage: slab_rank_bench.py <rank> <world> [particles] [steps]"""
import importlib, sys, os
os.environ["BCS_SLAB_NO_COMM"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
capi = importlib.import_module("simulation-server_b200.capi"); wl = importlib.import_module("simulation-server_b200.workloads")
dd = importlib.import_module("simulation-server_b200.distributed")
rank, world = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
sc, st, info = wl.long_vein(n)
planes = dd.slab_boundaries(sc, st, world)
sim = dd.create_slab_sim(sc, st, rank, world, 0, bytes(128), planes)
sim.step(10); sim.synchronize()
stream = torch.cuda.ExternalStream(sim.device_view().stream, device=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); sim.step(steps); e1.record(stream); sim.synchronize()
print(f"rank {rank}/{world} of {n}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step (no communication)  {sim.slab_counts()}")
if os.environ.get("SLAB_PROFILE", "1") != "0":
    prof = sim.profile_steps(5)
    tot = sum(v[0] for v in prof.values())
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        print(f"   {k:24s} {v[0] / 5 * 1e3:8.1f} us/step  {v[1] / 5:4.1f} launches/step  {v[0] / tot * 100:5.1f}%")
    print(f"   serialised sum {tot / 5 * 1e3:.1f} us/step, {sum(v[1] for v in prof.values()) / 5:.0f} launches/step")

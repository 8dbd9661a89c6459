"""Profiling aid: run ONE rank of an N-rank slab decomposition alone on one GPU (BCS_SLAB_NO_COMM=1: no NCCL, no
halos - the physics near the faces is wrong, the kernel workloads are right).  usage: slab_emulate.py <rank> <world> [particles]"""
import importlib, sys, os
os.environ["BCS_SLAB_NO_COMM"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("simulation-server_b200.capi"); wl = importlib.import_module("simulation-server_b200.workloads")
dd = importlib.import_module("simulation-server_b200.distributed")
rank, world = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
sc, st, info = wl.long_vein(n)
planes = dd.slab_boundaries(sc, st, world)
sim = dd.create_slab_sim(sc, st, rank, world, 0, bytes(128), planes, use_graph=False)
sim.step(10); sim.synchronize()
print(sim.slab_counts())
prof = sim.profile_steps(5)
tot = sum(v[0] for v in prof.values()) / 5
print(f"rank {rank}/{world} sum {tot*1e3:.0f} us:", "  ".join(f"{k} {v[0]/5*1e3:.1f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])))

"""Single-rank slab mode (world = 1: every blood cell owned, list-driven kernels, no NCCL) next to the plain
single-GPU mode: isolates the cost of the slab-mode code paths from the communication."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("simulation-server_b200.capi"); wl = importlib.import_module("simulation-server_b200.workloads")
dd = importlib.import_module("simulation-server_b200.distributed")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc, st, info = wl.long_vein(n)
for mode in ("plain", "slab1"):
    if mode == "plain":
        sim = capi.Sim(sc, use_graph=False); sim.upload_state(st)
    else:
        sim = dd.create_slab_sim(sc, st, 0, 1, 0, bytes(128), [float("inf"), float("-inf")])
    sim.step(10); sim.synchronize()
    prof = sim.profile_steps(5)
    tot = sum(v[0] for v in prof.values()) / 5
    print(mode, f"sum {tot*1e3:.0f} us:", "  ".join(f"{k} {v[0]/5*1e3:.0f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:14]))
    sim.close()

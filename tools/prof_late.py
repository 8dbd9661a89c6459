"""Advance the 1 M-particle long-vein workload to a late step (wall contacts everywhere), then run a few more
steps - the window an ncu capture is taken from (tools/gpu_ncu_late.sh)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("simulation-server_b200"); capi = importlib.import_module("simulation-server_b200.capi"); wl = importlib.import_module("simulation-server_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
sc, st, info = wl.long_vein(n)
sim = capi.Sim(sc, use_graph=False)
sim.upload_state(st)
sim.step(steps)
sim.synchronize()
print("late window")
sim.step(3)
sim.synchronize()

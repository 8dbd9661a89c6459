#!/usr/bin/env python
"""Small run for compute-sanitizer (racecheck / initcheck / memcheck): a fused run, a staged step and the debug views.
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py [particles]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200")
capi = importlib.import_module("simulation-server_b200.capi")
workloads = importlib.import_module("simulation-server_b200.workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
sc, st, _ = workloads.long_vein(n)
st = pkg.make_initial_state(sc, seed=3, xz_half_width=46.0, y_range=(-25.0, float(sc.vein_pos[:, 1].min()) + 40.0))
with capi.Sim(sc, device=0) as sim:
    sim.upload_state(st)
    sim.step(4)
    for stage in range(9):
        sim.run_stage(stage)
    sim.debug_candidates()
    sim.step(1)
    sim.synchronize()
    print("ok", sim.stats())

#!/usr/bin/env python
"""Developer loop on the GPU box: time ONE staged entry point (bcs_run_stage) back to back on the bench workload.
usage: bench_stage.py <stage: springs|collide|grid|finish> [reps] [--particles N]   (BCS_* switches via the environment)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
capi = importlib.import_module("simulation-server_b200.capi")
workloads = importlib.import_module("simulation-server_b200.workloads")
stage = {"springs": capi.STAGE_SPRINGS, "collide": capi.STAGE_PARTICLE_COLLISIONS, "grid": capi.STAGE_GRID_PARTICLES,
         "finish": capi.STAGE_INTEGRATE_PARTICLES}[sys.argv[1]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sc, st, info = workloads.long_vein(1_000_000)
sim = capi.Sim(sc, device=0)
sim.upload_state(st)
sim.step(30)
sim.build_grid()
stream = torch.cuda.ExternalStream(sim.device_view().stream, device=0)
for _ in range(5):
    sim.run_stage(stage)
sim.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    sim.run_stage(stage)
e1.record(stream)
sim.synchronize()
env = {k: v for k, v in os.environ.items() if k.startswith("BCS_")}
print(f"{sys.argv[1]:8s} {env}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call")

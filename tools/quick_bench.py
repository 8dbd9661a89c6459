#!/usr/bin/env python
"""Developer loop on the GPU box: ms/step (graph replay, CUDA events) + per-kernel table of the bench workload, without
the parity / e2e / CPU legs of bench.py.  Environment switches (BCS_*) are read at handle creation, so variants are run
as `env BCS_X=1 python tools/quick_bench.py`.

usage: quick_bench.py [--workload long_vein] [--particles 1000000] [--steps 100] [--warmup 10] [--tag name]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("simulation-server_b200.capi")
workloads = importlib.import_module("simulation-server_b200.workloads")


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="long_vein")
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--tag", default="")
    ap.add_argument("--semantics", default="clean")
    a = ap.parse_args()
    sc, st, info = workloads.by_name(a.workload, a.particles)
    sim = capi.Sim(sc, semantics=capi.SEM_REFERENCE if a.semantics == "reference" else capi.SEM_CLEAN, device=0)
    sim.upload_state(st)
    stream = torch.cuda.ExternalStream(sim.device_view().stream, device=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.step(a.warmup)
    sim.synchronize()
    e0.record(stream)
    sim.step(a.steps)
    e1.record(stream)
    sim.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    prof = sim.profile_steps(5)
    env = {k: v for k, v in os.environ.items() if k.startswith("BCS_")}
    print(f"== {a.tag or info['workload']} {env}: {ms * 1e3:.1f} us/step  ({sim.n_particles / ms / 1e6:.2f} G particle-steps/s)")
    tot = sum(v[0] for v in prof.values())
    line = "   " + "  ".join(f"{k}={v[0] / max(v[1], 1) * 1e3:.1f}" + (f"x{v[1] / 5:.1f}" if v[1] != 5 else "") for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]))
    print(line + f"   [sum {tot / 5 * 1e3:.0f} us/step]")
    sim.close()


if __name__ == "__main__":
    main()

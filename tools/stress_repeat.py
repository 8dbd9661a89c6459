#!/usr/bin/env python
"""Bitwise repeatability stress at bench scale: the same state advanced by two handles must give the same bits.
usage: stress_repeat.py [trials] [steps] [particles]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
capi = importlib.import_module("simulation-server_b200.capi")
workloads = importlib.import_module("simulation-server_b200.workloads")
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 10
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
sc, st, _ = workloads.long_vein(n)
names = {capi.PARTICLE_POS: "pos", capi.PARTICLE_VEL: "vel", capi.PARTICLE_FRC: "frc", capi.VEIN_POS: "vpos", capi.VEIN_VEL: "vvel"}

def run():
    out = []
    with capi.Sim(sc, device=0) as sim:
        sim.upload_state(st)
        for s in range(steps):
            sim.step(1 if os.environ.get("STRESS_SINGLE") else steps)
            out.append({k: np.stack(sim.download(k), 1) for k in names})
            if not os.environ.get("STRESS_SINGLE"):
                break
        grid = sim.grid(0)
    return out, grid

def dirty(seed):
    """another scene state through the same allocations: stale device memory then differs from what a fresh handle expects"""
    sc2, st2, _ = workloads.long_vein(n, seed=seed)
    with capi.Sim(sc2, device=0) as sim:
        sim.upload_state(st2)
        sim.step(3)
        sim.synchronize()

ref, gref = run()
bad = 0
for t in range(trials):
    if os.environ.get("STRESS_DIRTY"):
        dirty(100 + t)
    got, g = run()
    for s, (a, b) in enumerate(zip(ref, got)):
        for k, nm in names.items():
            d = (a[k] != b[k]).any(axis=1)
            if d.any():
                bad += 1
                idx = np.nonzero(d)[0]
                print(f"trial {t} step {s}: {nm} differs for {d.sum()} rows, first {idx[:5]}, max |d| {np.abs(a[k][idx] - b[k][idx]).max():.3e}", flush=True)
    if not (np.array_equal(g[0], gref[0]) and np.array_equal(g[1], gref[1])):
        print(f"trial {t}: sorted grid differs", flush=True)
print("STRESS", "FAIL" if bad else "PASS", trials, "trials")

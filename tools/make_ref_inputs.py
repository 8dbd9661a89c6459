#!/usr/bin/env python
"""Write the seeded initial states the reference harness / oracle / product all start from.

usage: python tools/make_ref_inputs.py            (writes oracle/_ref/state_<cfg>_<variant>.bcsd)
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200")

OUT = os.path.join(ROOT, "oracle", "_ref")
# (cfg, variant, kwargs): "spawn" = the reference's own spawn box; "wide" = cells spread to and beyond
# the vein wall so that vein-wall collisions happen within a few steps.
CASES = [
    ("cfg1", "spawn", dict(seed=1234)),
    ("cfg1", "wide", dict(seed=1234, xz_half_width=52.0, y_range=(-30.0, -380.0))),
    ("mini3", "spawn", dict(seed=1234, y_range=(-20.0, -60.0))),
    ("mini3", "wide", dict(seed=4321, xz_half_width=52.0, y_range=(-30.0, -120.0))),
    ("cfg2", "spawn", dict(seed=1234)),
    ("cfg2", "wide", dict(seed=1234, xz_half_width=50.0, y_range=(-30.0, -400.0))),
    ("cfg3", "vein", dict(seed=1234, xz_half_width=34.0, y_range=(-30.0, -400.0))),
]

for cfg, variant, kw in CASES:
    scene_path = os.path.join(OUT, f"scene_{cfg}.bcsd")
    if not os.path.exists(scene_path):
        print("skip", cfg, "(no scene file)")
        continue
    sc = pkg.Scene.load(scene_path)
    st = pkg.make_initial_state(sc, **kw)
    path = os.path.join(OUT, f"state_{cfg}_{variant}.bcsd")
    pkg.bcsd.write(path, st)
    print(path, st["pos_x"].size)

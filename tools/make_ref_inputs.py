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
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES  # noqa: E402

for (cfg, variant), kw in CASES.items():
    scene_path = os.path.join(OUT, f"scene_{cfg}.bcsd")
    if not os.path.exists(scene_path):
        print("skip", cfg, "(no scene file)")
        continue
    sc = pkg.Scene.load(scene_path)
    st = pkg.make_initial_state(sc, **kw)
    path = os.path.join(OUT, f"state_{cfg}_{variant}.bcsd")
    pkg.bcsd.write(path, st)
    print(path, st["pos_x"].size)

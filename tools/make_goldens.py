#!/usr/bin/env python
"""Turn the raw dumps of the headless reference run (tools/gpu_ref_goldens.sh -> gpurun_out/ref/*.tgz)
into the small committed fixtures under tests/golden/.

    python tools/make_goldens.py [raw_dir=/tmp/ref]

Fixtures (all numpy .npz, compressed):
  scene_<cfg>.npz              scene file content (user-level definition + tables derived by the REAL
                               reference headers, see oracle/ref_harness/ref_scene_dump.cpp); the vein
                               mesh is stored once (vein_default.npz)
  setup_<cfg>.npz              tables the reference derives on the host at construction (radii, centres)
  ref_<cfg>_<variant>_stepNNNNN.npz   stage-by-stage dump of step N (staged mode of ref_headless)
  ref_<cfg>_<variant>_final.npz       state after 100 steps
"""
import glob
import importlib
import os
import sys
import tarfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200")
bcsd = pkg.bcsd

RAW = sys.argv[1] if len(sys.argv) > 1 else "/tmp/ref"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

if not os.path.isdir(RAW) or not os.listdir(RAW):
    os.makedirs(RAW, exist_ok=True)
    for tgz in glob.glob(os.path.join(ROOT, "gpurun_out", "ref", "*.tgz")):
        with tarfile.open(tgz) as t:
            t.extractall(RAW)

VEIN_KEYS = ["vein_x", "vein_y", "vein_z", "vein_indices", "vein_nbr_ids", "vein_nbr_len"]


def save(name, arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"{name:44s} {os.path.getsize(path) / 1024:9.1f} KiB")


# ---- scenes
vein_saved = False
for cfg in ["cfg1", "mini3", "cfg2"]:
    a = bcsd.read(os.path.join(ROOT, "oracle", "_ref", f"scene_{cfg}.bcsd"))
    if not vein_saved:
        save("vein_default.npz", {k: a[k] for k in VEIN_KEYS})
        vein_saved = True
    save(f"scene_{cfg}.npz", {k: v for k, v in a.items() if k not in VEIN_KEYS})

# ---- which steps / arrays to keep (keeps the committed size small)
PARTICLE = ["begin.pos", "begin.vel", "begin.frc", "springs.frc", "pcoll.frc", "vcoll.frc", "vcoll.vel",
            "integrate.pos", "integrate.vel", "end.pos", "end.vel"]
VEIN = ["begin.vein_pos", "begin.vein_vel", "vein_gather.vein_frc", "vcoll.vein_frc", "integrate.vein_pos",
        "integrate.vein_vel"]
KEEP = {
    ("mini3", "wide"): {1: True, 2: True, 3: False, 10: True},      # step -> keep vein arrays?
    ("mini3", "spawn"): {1: False, 2: False},
    ("cfg1", "spawn"): {1: False, 2: False},
    ("cfg1", "wide"): {1: True, 2: False},
}


def vec_keys(prefix):
    return [prefix + s for s in ("_x", "_y", "_z")]


for (cfg, var), steps in KEEP.items():
    d = os.path.join(RAW, f"{cfg}_{var}_staged")
    setup = bcsd.read(os.path.join(d, "setup.bcsd"))
    save(f"setup_{cfg}.npz", {k: setup[k] for k in ["bounding_spheres", "initial_radiuses", "models_x", "models_y",
                                                     "models_z", "tri_centers_x", "tri_centers_y", "tri_centers_z",
                                                     "smallest_radius_in_type"]})
    for step, keep_vein in steps.items():
        a = bcsd.read(os.path.join(d, f"step{step:05d}.bcsd"))
        out = {}
        for p in PARTICLE:
            for k in vec_keys(p):
                out[k] = a[k]
        for k in vec_keys("springs.centers"):
            out[k] = a[k]
        for k in ["pgrid.keys", "pgrid.ids", "pgrid.table_cells", "pgrid.table_starts", "pgrid.table_ends", "pgrid.dims"]:
            out[k] = a[k]
        if step == 1:
            for k in ["tgrid.keys", "tgrid.ids", "tgrid.table_cells", "tgrid.table_starts", "tgrid.table_ends", "tgrid.dims"]:
                out[k] = a[k]
        if keep_vein:
            for p in VEIN:
                for k in vec_keys(p):
                    out[k] = a[k]
        save(f"ref_{cfg}_{var}_step{step:05d}.npz", out)
    fin = bcsd.read(os.path.join(d, "final.bcsd"))
    save(f"ref_{cfg}_{var}_final.npz", {k: fin[k] for p in ["final.pos", "final.vel", "final.frc"] for k in vec_keys(p)})

# plain (unstaged) run of the reference: must agree with the staged run of the same binary
for cfg, var in [("mini3", "wide"), ("cfg1", "spawn")]:
    fin = bcsd.read(os.path.join(RAW, f"{cfg}_{var}_plain", "final.bcsd"))
    save(f"ref_{cfg}_{var}_plain_final.npz", {k: fin[k] for p in ["final.pos", "final.vel", "final.frc"] for k in vec_keys(p)})

# checksums of the seeded input states (the states themselves are regenerated from the seed)
import hashlib
sums = {}
for f in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "state_*.bcsd"))):
    a = bcsd.read(f)
    h = hashlib.sha256()
    for k in sorted(a):
        h.update(a[k].tobytes())
    sums[os.path.basename(f)] = h.hexdigest()
with open(os.path.join(OUT, "state_checksums.txt"), "w") as f:
    for k, v in sums.items():
        f.write(f"{k} {v}\n")
print(open(os.path.join(OUT, "state_checksums.txt")).read())

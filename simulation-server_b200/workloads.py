"""Synthetic benchmark / test workloads (BASELINE.json configs), generated on the host.

No reference source is read at run time: the two blood-cell presets of the reference's default config
(White_blood_cell_One, Blood_dust_One - spring lists and model vertices) come from the committed fixture
tests/golden/scene_cfg1.npz, which tools/make_goldens.py extracted through the real reference headers.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np

from .scene import CellDef, Scene, make_cylinder_vein
from .state import make_initial_state

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRESETS = os.path.join(_ROOT, "tests", "golden", "scene_cfg1.npz")

# number density of the reference's default scene: 600 blood cells spawned over a 180-unit stretch of vein
# (simulation_controller.cu:205-207 with particleCount >= 1000)
DEFAULT_CELLS_PER_UNIT_LENGTH = 600.0 / 180.0


def reference_presets(path: str = PRESETS):
    """[(name, CellDef-with-count-0)]: White_blood_cell_One and Blood_dust_One in the user's order."""
    a = np.load(path)
    ut = a["user_types"].reshape(-1, 3)
    se = a["user_spring_se"].reshape(-1, 2)
    sl = a["user_spring_len"]
    uv = a["user_vertices"].reshape(-1, 3)
    out, so, vo = [], 0, 0
    for (cnt, p, ns), name in zip(ut, ["White_blood_cell_One", "Blood_dust_One"]):
        out.append((name, CellDef(0, int(p), se[so:so + ns].copy(), sl[so:so + ns].copy(), uv[vo:vo + p].copy())))
        so += ns
        vo += p
    return out


def long_vein(n_particles: int = 1_000_000, cells_per_unit_length: float = DEFAULT_CELLS_PER_UNIT_LENGTH,
              seed: int = 1234, use_blood_flow: int = 1) -> Tuple[Scene, Dict[str, np.ndarray], Dict[str, object]]:
    """BASELINE.json configs[2]/[3]: mixed blood-cell types (the reference's two presets, half and half) in a
    generated straight vein (radius 50, rings of 100 vertices every 5 units, like the trunk of the default
    mesh) whose length keeps the blood-cell number density of the reference's default scene; full vein-wall
    triangle collision.  Blood cells are spread uniformly along the whole vein."""
    presets = reference_presets()
    ppc = presets[0][1].particles_in_cell
    n_cells = max(2, n_particles // ppc)
    per_type = n_cells // 2
    length = float(np.ceil(n_cells / cells_per_unit_length / 5.0) * 5.0) + 60.0
    vp, vi, ec, er = make_cylinder_vein(length=length)
    defs = [CellDef(per_type, d.particles_in_cell, d.springs, d.spring_lengths, d.vertices) for _, d in presets]
    sc = Scene(user_defs=defs, vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er)
    sc.flags["use_blood_flow"] = int(use_blood_flow)
    st = make_initial_state(sc, seed=seed, y_range=(-20.0, -(length - 40.0)))
    info = {
        "workload": f"long_vein_{2 * per_type * ppc}",
        "particles": 2 * per_type * ppc,
        "blood_cells": 2 * per_type,
        "cell_types": "Blood_dust_One x%d + White_blood_cell_One x%d (20 particles each)" % (per_type, per_type),
        "vein": "straight cylinder r=50 length=%d: %d vertices, %d triangles" % (length, vp.shape[0], vi.shape[0]),
        "vein_wall_collision": "full",
        "density": "reference default (600 blood cells per 180 units of vein)",
    }
    return sc, st, info

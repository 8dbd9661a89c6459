"""Synthetic benchmark / test workloads (BASELINE.json configs[0..4]), generated on the host.

No reference source is read at run time.  The two blood-cell presets of the reference's default config
(White_blood_cell_One, Blood_dust_One - spring lists and model vertices, src/config/blood_cell_presets.hpp:13,174)
and its default vein mesh (src/config/vein_definition.hpp:12) ship as package data (``data/presets.npz``,
``data/vein_default.npz``); tools/make_goldens.py extracted them through the real reference headers.

    cfg1        configs[0]  the reference's default scene: 100 White_blood_cell_One + 500 Blood_dust_One, default vein
    cfg2        configs[1]  default vein, one type (White_blood_cell_One) x 5 000 = 100 k particles
    cfg3        configs[2]  default vein, both types x 25 000 = 1 M particles (dense: ~330 candidates per particle)
    long_vein   configs[3]  1 M particles at the default scene's number density in a generated straight vein
    cfg5        configs[4]  high-hematocrit stress scene: 10 M particles, 3x the default density, generated straight vein
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np

from .scene import CellDef, Scene, make_cylinder_vein
from .state import make_initial_state

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
PRESETS = os.path.join(_DATA, "presets.npz")
DEFAULT_VEIN = os.path.join(_DATA, "vein_default.npz")

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


def default_vein():
    """(vein_pos, vein_indices, ending_centers, ending_radii) of the reference's default config."""
    v = np.load(DEFAULT_VEIN)
    p = np.load(PRESETS)
    pos = np.stack([v["vein_x"], v["vein_y"], v["vein_z"]], axis=1).astype(np.float32)
    idx = v["vein_indices"].reshape(-1, 3).astype(np.uint32)
    return pos, idx, p["ending_centers"].reshape(-1, 3).astype(np.float32), p["ending_radii"].astype(np.float32)


def _with_counts(counts):
    return [CellDef(int(c), d.particles_in_cell, d.springs, d.spring_lengths, d.vertices)
            for c, (_, d) in zip(counts, reference_presets()) if c > 0]


def default_vein_scene(n_white: int, n_dust: int, seed: int = 1234, use_blood_flow: int = 1, **state_kw):
    """Blood cells of the two presets in the reference's default vein, spawned by the reference's rule
    (simulation_controller.cu:199-243) unless ``state_kw`` widens the spawn box."""
    vp, vi, ec, er = default_vein()
    sc = Scene(user_defs=_with_counts((n_white, n_dust)), vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er)
    sc.flags["use_blood_flow"] = int(use_blood_flow)
    st = make_initial_state(sc, seed=seed, **state_kw)
    n = 20 * (n_white + n_dust)
    info = {
        "particles": n, "blood_cells": n_white + n_dust,
        "cell_types": "White_blood_cell_One x%d + Blood_dust_One x%d (20 particles each)" % (n_white, n_dust),
        "vein": "reference default mesh: %d vertices, %d triangles" % (vp.shape[0], vi.shape[0]),
        "vein_wall_collision": "full",
    }
    return sc, st, info


def cfg1(seed: int = 1234):
    sc, st, info = default_vein_scene(100, 500, seed)
    info.update(workload="cfg1_default_scene_12000", density="reference spawn box (simulation_controller.cu:199-243)")
    return sc, st, info


def cfg2(seed: int = 1234):
    sc, st, info = default_vein_scene(5000, 0, seed, xz_half_width=50.0, y_range=(-30.0, -400.0))
    info.update(workload="cfg2_default_vein_100000", density="5 000 blood cells over 370 units of the default vein's trunk")
    return sc, st, info


def cfg3(seed: int = 1234):
    sc, st, info = default_vein_scene(25000, 25000, seed, xz_half_width=34.0, y_range=(-30.0, -400.0))
    info.update(workload="cfg3_default_vein_1000000", density="50 000 blood cells over 370 units of the default vein's trunk (dense)")
    return sc, st, info


def long_vein(n_particles: int = 1_000_000, cells_per_unit_length: float = DEFAULT_CELLS_PER_UNIT_LENGTH,
              seed: int = 1234, use_blood_flow: int = 1, name: str = "long_vein") -> Tuple[Scene, Dict[str, np.ndarray], Dict[str, object]]:
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
        "workload": f"{name}_{2 * per_type * ppc}",
        "particles": 2 * per_type * ppc,
        "blood_cells": 2 * per_type,
        "cell_types": "Blood_dust_One x%d + White_blood_cell_One x%d (20 particles each)" % (per_type, per_type),
        "vein": "straight cylinder r=50 length=%d: %d vertices, %d triangles" % (length, vp.shape[0], vi.shape[0]),
        "vein_wall_collision": "full",
        "density": "%.2f blood cells per unit of vein length (reference default scene: %.2f)" % (cells_per_unit_length, DEFAULT_CELLS_PER_UNIT_LENGTH),
    }
    return sc, st, info


def cfg5(n_particles: int = 10_000_000, seed: int = 1234):
    """configs[4]: high-hematocrit stress scene - three times the default number density (dense collisions)."""
    return long_vein(n_particles, 3.0 * DEFAULT_CELLS_PER_UNIT_LENGTH, seed, name="cfg5_high_hematocrit")


def by_name(name: str, particles: int = 1_000_000, seed: int = 1234):
    """The workload ``bench.py --workload`` names."""
    if name == "cfg1":
        return cfg1(seed)
    if name == "cfg2":
        return cfg2(seed)
    if name == "cfg3":
        return cfg3(seed)
    if name == "long_vein":
        return long_vein(particles, seed=seed)
    if name == "cfg5":
        return cfg5(particles, seed)
    raise ValueError(f"unknown workload {name!r}")

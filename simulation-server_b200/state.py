"""Host-side synthetic initial states (seeded, identical bytes for product, oracle and reference harness).

Follows the reference's spawn rule (simulation_controller.cu:199-243): blood-cell centres uniform in
x,z in (U-0.5)*1.2*cylinderRadius, y in [minSpawnY - 180*min(N/1000,1), minSpawnY]; velocity
(U*2c - c, 0.894*initVelocityY, (-id%2)*sqrt(c^2 - vx^2)) with c = |0.5*vy|; every particle = centre +
model vertex, cell velocity, zero force.  The RNG is numpy PCG64 instead of cuRAND XORWOW (the reference
seeds from time(0), so its stream is not reproducible anyway).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .scene import Layout, Scene


def make_initial_state(scene: Scene, seed: int = 1234, y_range: Optional[Tuple[float, float]] = None,
                       xz_half_width: Optional[float] = None, layout: Optional[Layout] = None) -> Dict[str, np.ndarray]:
    """Returns {"pos_x","pos_y","pos_z","vel_*","frc_*"} float32 arrays of length N (final particle order)."""
    lay = layout or scene.layout()
    ph = scene.physics
    n_cells, n = lay.n_cells, lay.n_particles
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random((n_cells, 4), dtype=np.float32)
    radius = np.float32(ph["cylinder_radius"])
    half = np.float32(1.2) * radius if xz_half_width is None else np.float32(2.0 * xz_half_width)
    cx = (u[:, 0] - np.float32(0.5)) * half
    cz = (u[:, 2] - np.float32(0.5)) * half
    if y_range is None:
        depth = np.float32(180.0) * np.float32(min(n / 1000.0, 1.0))
        cy = np.float32(ph["min_spawn_y"]) - depth * u[:, 1]
    else:
        y_hi, y_lo = np.float32(max(y_range)), np.float32(min(y_range))
        cy = y_hi - (y_hi - y_lo) * u[:, 1]
    vy = np.float32(ph["random_velocity_modifier"]) * np.float32(ph["init_velocity_y"])
    c = np.abs(np.float32(0.5) * vy)
    vx = u[:, 3] * np.float32(2.0) * c - c
    sign = -(np.arange(n_cells) % 2).astype(np.float32)         # (-1*id % 2) in C++: 0 for even ids, -1 for odd
    vz = sign * np.sqrt(np.maximum(c * c - vx * vx, np.float32(0.0))).astype(np.float32)

    pos = np.empty((n, 3), np.float32)
    vel = np.empty((n, 3), np.float32)
    for t in range(lay.n_types):
        cnt, p = int(lay.counts[t]), int(lay.particles_in_cell[t])
        ps, cs, ms = int(lay.particle_starts[t]), int(lay.cell_starts[t]), int(lay.model_starts[t])
        centre = np.stack([cx[cs:cs + cnt], cy[cs:cs + cnt], cz[cs:cs + cnt]], axis=1)            # (cnt,3)
        cvel = np.stack([vx[cs:cs + cnt], np.full(cnt, vy, np.float32), vz[cs:cs + cnt]], axis=1)
        model = lay.model[ms:ms + p]                                                             # (p,3)
        pos[ps:ps + cnt * p] = (centre[:, None, :] + model[None, :, :]).reshape(-1, 3)
        vel[ps:ps + cnt * p] = np.repeat(cvel, p, axis=0)
    z = np.zeros(n, np.float32)
    return {
        "pos_x": np.ascontiguousarray(pos[:, 0]), "pos_y": np.ascontiguousarray(pos[:, 1]),
        "pos_z": np.ascontiguousarray(pos[:, 2]),
        "vel_x": np.ascontiguousarray(vel[:, 0]), "vel_y": np.ascontiguousarray(vel[:, 1]),
        "vel_z": np.ascontiguousarray(vel[:, 2]),
        "frc_x": z.copy(), "frc_y": z.copy(), "frc_z": z.copy(),
    }

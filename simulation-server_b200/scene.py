"""Scene description: the reference's compile-time ``src/config`` content as run-time data.

A :class:`Scene` carries exactly what a user writes into the reference's config headers

* ``config/blood_cells_definition.hpp``  -> ``user_defs`` (count, particles per cell, spring list, model vertices),
  in the USER's order (the fold / unique / sort of ``meta_factory/blood_cell_factory.hpp:60-162`` is applied by
  the library, and mirrored by :func:`derive_layout` for host-side helpers),
* ``config/vein_definition.hpp``         -> vein vertices, triangle indices, vein endings,
* ``config/simulation.hpp`` / ``physics.hpp`` -> ``physics`` + ``flags`` + cell sizes + margins.

Scenes are stored as BCSD files (``bcsd.py``).  Files written by ``oracle/ref_harness/ref_scene_dump.cpp``
additionally hold the tables the REAL reference headers derived (``types``, ``type_starts``, ``spring_graph``,
``grid_min`` ...); those are kept in ``Scene.expected`` and are only used by tests as goldens.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import bcsd

# Order of the "physics" vector in scene files (twin: ref_scene_dump.cpp; names follow config/physics.hpp).
PHYSICS_FIELDS = [
    "dt", "velocity_collision_damping", "particle_k_sniff", "vein_k_sniff", "particle_d_fact", "vein_d_fact",
    "vein_boundaries_velocity_damping", "vein_collision_force_intensity", "viscous_damping",
    "collision_spring_coeff", "collision_damping_coeff", "collision_shear_coeff",
    "max_cell_size_factor_before_brake", "big_particle_braking_intensity",
    "init_velocity_x", "init_velocity_y", "init_velocity_z", "random_velocity_modifier",
    "vein_impact_distance", "vein_impact_minimal_force_distance", "gx", "gy", "gz",
    "grid_y_margin", "grid_xz_margin", "min_spawn_y", "cylinder_radius",
]
# config/physics.hpp:4-34 + config/simulation.hpp:5,19-21 + vein_factory.hpp:88 (reference defaults)
DEFAULT_PHYSICS = dict(zip(PHYSICS_FIELDS, [
    0.008, 0.96, 25.0, 0.1, 10.0, 0.7, 0.65, 0.005, 0.39, 6.0, 6.0, 4.0, 1.5, 0.8,
    0.0, -80.0, 0.0, 0.894, 6.0, 0.001, 0.0, -25.0, 0.0, 40.0, 40.0, -20.0, 50.0]))
FLAG_FIELDS = ["use_blood_flow", "enable_reaction_force", "enable_big_cells_brake", "bounding_spheres_coeff",
               "max_frames", "gpu_count"]
DEFAULT_FLAGS = dict(zip(FLAG_FIELDS, [1, 1, 1, 3, 600, 1]))


@dataclass
class CellDef:
    """One ``BloodCellDef<count, particlesInCell, ..., Springs, Vertices, ...>`` (blood_cells_def_type.hpp:21-31)."""
    count: int
    particles_in_cell: int
    springs: np.ndarray          # (S, 2) int32  start, end
    spring_lengths: np.ndarray   # (S,)  float32
    vertices: np.ndarray         # (P, 3) float32

    def same_type(self, other: "CellDef") -> bool:
        """``IsDuplicate`` (blood_cell_factory.hpp:52-56): same P and the same spring list."""
        return (self.particles_in_cell == other.particles_in_cell
                and self.springs.shape == other.springs.shape
                and np.array_equal(self.springs, other.springs)
                and np.array_equal(self.spring_lengths, other.spring_lengths))


@dataclass
class Scene:
    user_defs: List[CellDef]
    vein_pos: np.ndarray         # (V, 3) float32
    vein_indices: np.ndarray     # (T, 3) uint32
    ending_centers: np.ndarray   # (E, 3) float32
    ending_radii: np.ndarray     # (E,)  float32
    cell_size: Sequence[int] = (2, 2, 2)
    tri_cell_size: Sequence[int] = (25, 25, 25)
    physics: Dict[str, float] = field(default_factory=lambda: dict(DEFAULT_PHYSICS))
    flags: Dict[str, int] = field(default_factory=lambda: dict(DEFAULT_FLAGS))
    expected: Optional[Dict[str, np.ndarray]] = None   # reference-derived goldens, if the file had them

    # ------------------------------------------------------------------ io
    @staticmethod
    def load(path: str) -> "Scene":
        a = bcsd.read(path)
        ut = a["user_types"].reshape(-1, 3)
        se = a["user_spring_se"].reshape(-1, 2)
        sl = a["user_spring_len"]
        uv = a["user_vertices"].reshape(-1, 3)
        defs, so, vo = [], 0, 0
        for cnt, p, ns in ut:
            defs.append(CellDef(int(cnt), int(p), se[so:so + ns].copy(), sl[so:so + ns].copy(), uv[vo:vo + p].copy()))
            so += ns
            vo += p
        sc = Scene(
            user_defs=defs,
            vein_pos=np.stack([a["vein_x"], a["vein_y"], a["vein_z"]], axis=1).astype(np.float32),
            vein_indices=a["vein_indices"].reshape(-1, 3).astype(np.uint32),
            ending_centers=a["ending_centers"].reshape(-1, 3).astype(np.float32),
            ending_radii=a["ending_radii"].astype(np.float32),
            cell_size=tuple(int(v) for v in a["cell_size"]),
            tri_cell_size=tuple(int(v) for v in a["tri_cell_size"]),
            physics={k: float(np.float32(v)) for k, v in zip(PHYSICS_FIELDS, a["physics"])},
            flags={k: int(v) for k, v in zip(FLAG_FIELDS, a["flags"])},
        )
        golden_keys = ["types", "type_starts", "spring_graph", "model_x", "model_y", "model_z", "totals",
                       "grid_min", "grid_max", "grid_whd", "vein_nbr_ids", "vein_nbr_len", "spring_counts"]
        if all(k in a for k in golden_keys):
            sc.expected = {k: a[k] for k in golden_keys}
        return sc

    def save(self, path: str) -> None:
        ut, se, sl, uv = [], [], [], []
        for d in self.user_defs:
            ut += [d.count, d.particles_in_cell, len(d.spring_lengths)]
            se.append(np.asarray(d.springs, np.int32).reshape(-1, 2))
            sl.append(np.asarray(d.spring_lengths, np.float32))
            uv.append(np.asarray(d.vertices, np.float32).reshape(-1, 3))
        arrays = {
            "user_types": np.asarray(ut, np.int32),
            "user_spring_se": np.concatenate(se).reshape(-1).astype(np.int32),
            "user_spring_len": np.concatenate(sl).astype(np.float32),
            "user_vertices": np.concatenate(uv).reshape(-1).astype(np.float32),
            "vein_x": np.ascontiguousarray(self.vein_pos[:, 0], np.float32),
            "vein_y": np.ascontiguousarray(self.vein_pos[:, 1], np.float32),
            "vein_z": np.ascontiguousarray(self.vein_pos[:, 2], np.float32),
            "vein_indices": np.ascontiguousarray(self.vein_indices, np.uint32).reshape(-1),
            "ending_centers": np.ascontiguousarray(self.ending_centers, np.float32).reshape(-1),
            "ending_radii": np.ascontiguousarray(self.ending_radii, np.float32),
            "cell_size": np.asarray(self.cell_size, np.int32),
            "tri_cell_size": np.asarray(self.tri_cell_size, np.int32),
            "physics": np.asarray([self.physics[k] for k in PHYSICS_FIELDS], np.float32),
            "flags": np.asarray([self.flags[k] for k in FLAG_FIELDS], np.int32),
        }
        if self.expected:
            arrays.update(self.expected)
        bcsd.write(path, arrays)

    # ------------------------------------------------------------------ derived quantities
    def layout(self) -> "Layout":
        return derive_layout(self)

    @property
    def n_vertices(self) -> int:
        return int(self.vein_pos.shape[0])

    @property
    def n_triangles(self) -> int:
        return int(self.vein_indices.shape[0])


@dataclass
class Layout:
    """Host-side mirror of the tables ``meta_factory`` derives (final type order and offsets)."""
    counts: np.ndarray           # per final type
    particles_in_cell: np.ndarray
    particle_starts: np.ndarray
    cell_starts: np.ndarray
    model_starts: np.ndarray
    graph_starts: np.ndarray
    model: np.ndarray            # (sum P, 3) float32
    spring_graph: np.ndarray     # (sum P^2,) float32, dense, graph[start + a*P + b]
    src_def: List[int]           # index of the user definition each final type was folded from
    grid_min: np.ndarray
    grid_max: np.ndarray

    @property
    def n_particles(self) -> int:
        return int((self.counts * self.particles_in_cell).sum())

    @property
    def n_cells(self) -> int:
        return int(self.counts.sum())

    @property
    def n_types(self) -> int:
        return int(self.counts.size)


def _not_power_of_two(n: int) -> bool:
    # the reference's `isPowerOfTwo` returns n & (n-1), i.e. true for NON powers of two (blood_cell_factory.hpp:119-122)
    return (n & (n - 1)) != 0


def _order_blood_cells(p1: int, p2: int) -> bool:
    """``orderBloodCells`` (blood_cell_factory.hpp:130-147), including its dead `& 0` branches."""
    if _not_power_of_two(p1) and not _not_power_of_two(p2):
        return True
    if not _not_power_of_two(p1) and _not_power_of_two(p2):
        return False
    if (p1 & 0) and (p2 & 1):
        return True
    if (p1 & 1) and (p2 & 0):
        return False
    return True


def _mp_sort(items: List[int], p_of) -> List[int]:
    """boost::mp11::mp_sort: quicksort, pivot = first element, stable partition by P<U, pivot>."""
    if len(items) <= 1:
        return list(items)
    pivot, rest = items[0], items[1:]
    first = [u for u in rest if _order_blood_cells(p_of(u), p_of(pivot))]
    second = [u for u in rest if not _order_blood_cells(p_of(u), p_of(pivot))]
    return _mp_sort(first, p_of) + [pivot] + _mp_sort(second, p_of)


def derive_layout(scene: Scene) -> Layout:
    defs = scene.user_defs
    # fold: for every user def, count = sum over all duplicates (blood_cell_factory.hpp:64-106)
    folded = [sum(e.count for e in defs if d.same_type(e)) for d in defs]
    # mp_unique_if: keep first occurrence of each duplicate class (:115)
    uniq: List[int] = []
    for i, d in enumerate(defs):
        if not any(defs[j].same_type(d) for j in uniq):
            uniq.append(i)
    order = _mp_sort(uniq, lambda i: defs[i].particles_in_cell)

    counts = np.asarray([folded[i] for i in order], np.int32)
    ppc = np.asarray([defs[i].particles_in_cell for i in order], np.int32)
    pstart = np.concatenate([[0], np.cumsum(counts * ppc)[:-1]]).astype(np.int32)
    cstart = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    mstart = np.concatenate([[0], np.cumsum(ppc)[:-1]]).astype(np.int32)
    gstart = np.concatenate([[0], np.cumsum(ppc * ppc)[:-1]]).astype(np.int32)
    model = np.concatenate([np.asarray(defs[i].vertices, np.float32).reshape(-1, 3) for i in order])
    graph = np.zeros(int((ppc * ppc).sum()), np.float32)
    for t, i in enumerate(order):
        d, p = defs[i], int(ppc[t])
        for (a, b), ln in zip(d.springs, d.spring_lengths):   # later entries overwrite (springGraphGenerator :292-328)
            graph[gstart[t] + a * p + b] = ln
            graph[gstart[t] + b * p + a] = ln
    ymargin = np.float32(scene.physics["grid_y_margin"])
    xzmargin = np.float32(scene.physics["grid_xz_margin"])
    margin = np.asarray([xzmargin, ymargin, xzmargin], np.float32)
    gmin = (scene.vein_pos.min(axis=0).astype(np.float32) - margin).astype(np.float32)
    gmax = (scene.vein_pos.max(axis=0).astype(np.float32) + margin).astype(np.float32)
    return Layout(counts, ppc, pstart, cstart, mstart, gstart, model, graph, order, gmin, gmax)


# ---------------------------------------------------------------------- generators
def make_cylinder_vein(length: float, radius: float = 50.0, ring_step: float = 5.0, ring_vertices: int = 100,
                       y_top: float = 0.0):
    """A straight tube along -y with the tessellation of the reference mesh's trunk.

    Rings of ``ring_vertices`` vertices every ``ring_step`` from ``y_top`` downward; ring r holds vertices
    [r*n, (r+1)*n); quads between ring r (upper) and ring r+1 (lower) are split as
    (lower_k, upper_k, upper_k+1), (lower_k, upper_k+1, lower_k+1) - the index pattern of
    config/vein_definition.hpp:15965 (``100,0,1, 100,1,101, ...``).
    Returns (vein_pos (V,3) f32, vein_indices (T,3) u32, ending_centers (1,3), ending_radii (1,)).
    """
    n = ring_vertices
    rings = int(round(length / ring_step)) + 1
    ang = (2.0 * np.pi * np.arange(n) / n)
    ring_x = (radius * np.cos(ang)).astype(np.float32)
    ring_z = (radius * np.sin(ang)).astype(np.float32)
    pos = np.empty((rings, n, 3), np.float32)
    pos[:, :, 0] = ring_x[None, :]
    pos[:, :, 2] = ring_z[None, :]
    pos[:, :, 1] = (np.float32(y_top) - np.float32(ring_step) * np.arange(rings, dtype=np.float32))[:, None]
    k = np.arange(n)
    k1 = (k + 1) % n
    tris = []
    for r in range(rings - 1):
        up, lo = r * n, (r + 1) * n
        t = np.empty((n, 2, 3), np.uint32)
        t[:, 0, 0] = lo + k
        t[:, 0, 1] = up + k
        t[:, 0, 2] = up + k1
        t[:, 1, 0] = lo + k
        t[:, 1, 1] = up + k1
        t[:, 1, 2] = lo + k1
        tris.append(t.reshape(-1, 3))
    idx = np.concatenate(tris).astype(np.uint32)
    y_bottom = float(pos[-1, 0, 1])
    end_c = np.asarray([[0.0, y_bottom, 0.0]], np.float32)
    end_r = np.asarray([radius * 0.7], np.float32)
    return pos.reshape(-1, 3), idx, end_c, end_r


def make_bifurcated_vein(trunk_length: float = 150.0, branch_length: float = 150.0, radius: float = 50.0, branch_radius: float = 36.0,
                         half_angle_deg: float = 25.0, ring_step: float = 5.0, ring_vertices: int = 100):
    """A Y-shaped vein in the tessellation of :func:`make_cylinder_vein`: a trunk along -y that splits into two branches
    leaning +-``half_angle_deg`` in x (the reference's default mesh, config/vein_definition.hpp:12, bifurcates the same way).
    The three tubes are separate ring stacks; the branches start one ring inside the trunk's end so that the walls
    overlap at the junction instead of being stitched.  One vein ending per branch.
    Returns (vein_pos (V,3) f32, vein_indices (T,3) u32, ending_centers (2,3), ending_radii (2,))."""
    n = ring_vertices
    ang = 2.0 * np.pi * np.arange(n) / n

    def tube(origin, direction, length, r):
        d = np.asarray(direction, np.float64)
        d /= np.linalg.norm(d)
        # orthonormal frame: u in the x-y plane, w = z axis (branches lean in x only)
        w = np.array([0.0, 0.0, 1.0])
        u = np.cross(d, w)
        u /= np.linalg.norm(u)
        rings = int(round(length / ring_step)) + 1
        t = (np.arange(rings) * ring_step)[:, None, None]
        circle = (np.cos(ang)[:, None] * u[None, :] + np.sin(ang)[:, None] * w[None, :]) * r
        return (np.asarray(origin, np.float64)[None, None, :] + t * d[None, None, :] + circle[None, :, :]).astype(np.float32)

    def stitch(rings, base):
        k = np.arange(n)
        k1 = (k + 1) % n
        tris = []
        for r in range(rings - 1):
            up, lo = base + r * n, base + (r + 1) * n
            t = np.empty((n, 2, 3), np.uint32)
            t[:, 0, 0] = lo + k; t[:, 0, 1] = up + k; t[:, 0, 2] = up + k1
            t[:, 1, 0] = lo + k; t[:, 1, 1] = up + k1; t[:, 1, 2] = lo + k1
            tris.append(t.reshape(-1, 3))
        return np.concatenate(tris)

    a = np.deg2rad(half_angle_deg)
    split = np.array([0.0, -trunk_length + ring_step, 0.0])
    tubes = [tube((0.0, 0.0, 0.0), (0.0, -1.0, 0.0), trunk_length, radius),
             tube(split + np.array([-0.5 * radius, 0.0, 0.0]), (-np.sin(a), -np.cos(a), 0.0), branch_length, branch_radius),
             tube(split + np.array([0.5 * radius, 0.0, 0.0]), (np.sin(a), -np.cos(a), 0.0), branch_length, branch_radius)]
    pos, idx, base = [], [], 0
    for tb in tubes:
        pos.append(tb.reshape(-1, 3))
        idx.append(stitch(tb.shape[0], base))
        base += tb.shape[0] * n
    ends = np.stack([tubes[1][-1].mean(axis=0), tubes[2][-1].mean(axis=0)]).astype(np.float32)
    # three decimals: every coordinate survives the reference's fixed-point header format (headers.py)
    return (np.round(np.concatenate(pos), 3).astype(np.float32), np.concatenate(idx).astype(np.uint32), np.round(ends, 3).astype(np.float32),
            np.asarray([branch_radius * 0.7] * 2, np.float32))


def make_rbc_celldef(count: int, radius: float = 3.9, thickness: float = 1.2) -> CellDef:
    """A red-blood-cell preset in the reference's BloodCellDef form (blood_cells_def_type.hpp:21-31; the reference ships
    White_blood_cell_One and Blood_dust_One only, blood_cell_presets.hpp:13,174): a biconcave disc sampled with 26
    particles - an 8-vertex equator, an 8-vertex ring at 0.55 R on either face (thick), a centre vertex on either face
    (thin: the dimple) - and springs along the rings, between neighbouring rings, across the disc and through it."""
    pts = []
    ring = lambda r, z, phase: [(r * np.cos(2 * np.pi * (k + phase) / 8), z, r * np.sin(2 * np.pi * (k + phase) / 8)) for k in range(8)]
    pts += ring(radius, 0.0, 0.0)                                   # 0..7   equator
    pts += ring(0.55 * radius, 0.5 * thickness, 0.5)                # 8..15  upper ring
    pts += ring(0.55 * radius, -0.5 * thickness, 0.5)               # 16..23 lower ring
    pts += [(0.0, 0.2 * thickness, 0.0), (0.0, -0.2 * thickness, 0.0)]   # 24, 25 dimple centres
    v = np.round(np.asarray(pts, np.float64), 3)
    springs = []
    for k in range(8):
        k1 = (k + 1) % 8
        springs += [(k, k1), (8 + k, 8 + k1), (16 + k, 16 + k1)]                  # along the rings
        springs += [(k, 8 + k), (k1, 8 + k), (k, 16 + k), (k1, 16 + k)]           # equator <-> rings
        springs += [(8 + k, 16 + k), (8 + k, 24), (16 + k, 25)]                   # through the disc, to the dimple
    springs += [(k, k + 4) for k in range(4)] + [(24, 25)]                          # across the disc
    se = np.asarray(springs, np.int32)
    length = np.round(np.linalg.norm(v[se[:, 0]] - v[se[:, 1]], axis=1), 4).astype(np.float32)
    return CellDef(int(count), 26, se, length, v.astype(np.float32))

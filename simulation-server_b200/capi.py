"""ctypes binding of the C ABI declared in ``include/bcs.h``.

``Sim`` is generic over (library, symbol prefix): the product always uses ``libbcs.so`` / ``bcs_``
(see ``load_library``); the tests reuse the same thin wrapper for the CPU oracle (``orc_``), which
lives under ``oracle/`` and is never loaded from here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Mapping, Dict, Optional, Tuple

import numpy as np

from .scene import PHYSICS_FIELDS, Scene

BCS_MAX_TYPES = 16
BCS_VEIN_MAX_NEIGHBORS = 9

SEM_CLEAN, SEM_REFERENCE = 0, 1

(PARTICLE_POS, PARTICLE_VEL, PARTICLE_FRC, VEIN_POS, VEIN_VEL, VEIN_FRC, CELL_CENTERS) = range(7)
(STAGE_GRID_PARTICLES, STAGE_GRID_TRIANGLES, STAGE_VEIN_GATHER, STAGE_SPRINGS, STAGE_PARTICLE_COLLISIONS,
 STAGE_VEIN_COLLISIONS, STAGE_INTEGRATE_PARTICLES, STAGE_INTEGRATE_VEIN, STAGE_VEIN_END) = range(9)
(TABLE_SPRING_GRAPH, TABLE_MODEL_X, TABLE_MODEL_Y, TABLE_MODEL_Z, TABLE_COLLISION_RADII, TABLE_INITIAL_RADII,
 TABLE_VEIN_NBR_IDS, TABLE_VEIN_NBR_LEN, TABLE_TRI_CENTERS_X, TABLE_TRI_CENTERS_Y, TABLE_TRI_CENTERS_Z) = range(11)


class Spring(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("length", C.c_float)]


class CellDefC(C.Structure):
    _fields_ = [("count", C.c_int32), ("particles_in_cell", C.c_int32), ("n_springs", C.c_int32),
                ("springs", C.POINTER(Spring)), ("vertices", C.POINTER(C.c_float))]


class Physics(C.Structure):
    _fields_ = [
        ("dt", C.c_float), ("velocity_collision_damping", C.c_float), ("particle_k_sniff", C.c_float),
        ("vein_k_sniff", C.c_float), ("particle_d_fact", C.c_float), ("vein_d_fact", C.c_float),
        ("vein_boundaries_velocity_damping", C.c_float), ("vein_collision_force_intensity", C.c_float),
        ("viscous_damping", C.c_float), ("collision_spring_coeff", C.c_float),
        ("collision_damping_coeff", C.c_float), ("collision_shear_coeff", C.c_float),
        ("max_cell_size_factor_before_brake", C.c_float), ("big_particle_braking_intensity", C.c_float),
        ("init_velocity", C.c_float * 3), ("random_velocity_modifier", C.c_float),
        ("vein_impact_distance", C.c_float), ("vein_impact_minimal_force_distance", C.c_float),
        ("gravity", C.c_float * 3), ("grid_y_margin", C.c_float), ("grid_xz_margin", C.c_float),
        ("min_spawn_y", C.c_float), ("cylinder_radius", C.c_float),
    ]


class SceneC(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_defs", C.c_int32), ("defs", C.POINTER(CellDefC)),
        ("n_vertices", C.c_int32), ("vein_x", C.POINTER(C.c_float)), ("vein_y", C.POINTER(C.c_float)),
        ("vein_z", C.POINTER(C.c_float)), ("n_triangles", C.c_int32), ("vein_indices", C.POINTER(C.c_uint32)),
        ("n_endings", C.c_int32), ("ending_centers", C.POINTER(C.c_float)), ("ending_radii", C.POINTER(C.c_float)),
        ("cell_size", C.c_int32 * 3), ("tri_cell_size", C.c_int32 * 3),
        ("use_blood_flow", C.c_int32), ("enable_reaction_force", C.c_int32), ("enable_big_cells_brake", C.c_int32),
        ("bounding_spheres_coeff", C.c_int32), ("physics", Physics),
    ]


class Opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("semantics", C.c_int32),
                ("use_graph", C.c_int32), ("collect_stats", C.c_int32), ("exhaustive_vein_traversal", C.c_int32),
                ("seed", C.c_uint64), ("stream", C.c_void_p)]


class SlabOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_int32), ("world", C.c_int32), ("spawn_rank", C.c_int32),
                ("y_lo", C.c_float), ("y_hi", C.c_float), ("halo_width", C.c_float), ("vertex_halo", C.c_float),
                ("migration_capacity", C.c_int32), ("halo_capacity", C.c_int32), ("nccl_unique_id", C.c_ubyte * 128)]


class TypeInfo(C.Structure):
    _fields_ = [("count", C.c_int32), ("particles_in_cell", C.c_int32), ("particle_start", C.c_int32),
                ("cell_start", C.c_int32), ("model_start", C.c_int32), ("graph_start", C.c_int32),
                ("src_def", C.c_int32), ("vein_end_warp_sync", C.c_int32), ("smallest_radius", C.c_float)]


class LayoutC(C.Structure):
    _fields_ = [("n_types", C.c_int32), ("n_particles", C.c_int32), ("n_cells", C.c_int32), ("n_model", C.c_int32),
                ("n_graph", C.c_int32), ("n_vertices", C.c_int32), ("n_triangles", C.c_int32),
                ("grid_dims", C.c_int32 * 3), ("grid_cells", C.c_int32), ("tri_grid_dims", C.c_int32 * 3),
                ("tri_grid_cells", C.c_int32), ("grid_min", C.c_float * 3), ("grid_max", C.c_float * 3),
                ("grid_size", C.c_float * 3), ("types", TypeInfo * BCS_MAX_TYPES)]


class DeviceView(C.Structure):
    _fields_ = [("particle_pos4", C.c_void_p), ("particle_vel4", C.c_void_p), ("particle_frc4", C.c_void_p),
                ("vein_pos4", C.c_void_p), ("vein_vel4", C.c_void_p), ("vein_frc4", C.c_void_p), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("pair_tests", C.c_uint64), ("pair_hits", C.c_uint64), ("triangle_tests", C.c_uint64),
                ("vein_hits", C.c_uint64), ("teleported_cells", C.c_uint64), ("out_of_bounds", C.c_uint64),
                ("wall_rebuilds", C.c_uint64)]


class BcsError(RuntimeError):
    pass


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libbcs.so")
_lib_cache: Dict[str, C.CDLL] = {}


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads the product library.  Fails loudly if it has not been built (no fallback of any kind)."""
    path = path or os.environ.get("BCS_LIBRARY") or LIB_PATH   # BCS_LIBRARY: another BUILD of libbcs (A/B measurements)
    if path not in _lib_cache:
        if not os.path.exists(path):
            raise BcsError(f"{path} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback")
        _lib_cache[path] = C.CDLL(path)
    return _lib_cache[path]


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class SceneHandle:
    """Owns the numpy buffers a ``bcs_scene`` points into."""

    def __init__(self, scene: Scene):
        self._keep = []
        defs = (CellDefC * len(scene.user_defs))()
        for i, d in enumerate(scene.user_defs):
            springs = (Spring * max(1, len(d.spring_lengths)))()
            for k, ((a, b), ln) in enumerate(zip(d.springs, d.spring_lengths)):
                springs[k] = Spring(int(a), int(b), float(ln))
            verts = np.ascontiguousarray(d.vertices, np.float32).reshape(-1)
            self._keep += [springs, verts]
            defs[i] = CellDefC(d.count, d.particles_in_cell, len(d.spring_lengths), springs, _fp(verts))
        vx = np.ascontiguousarray(scene.vein_pos[:, 0], np.float32)
        vy = np.ascontiguousarray(scene.vein_pos[:, 1], np.float32)
        vz = np.ascontiguousarray(scene.vein_pos[:, 2], np.float32)
        idx = np.ascontiguousarray(scene.vein_indices, np.uint32).reshape(-1)
        ec = np.ascontiguousarray(scene.ending_centers, np.float32).reshape(-1)
        er = np.ascontiguousarray(scene.ending_radii, np.float32).reshape(-1)
        self._keep += [defs, vx, vy, vz, idx, ec, er]
        ph = Physics()
        p = scene.physics
        for name in PHYSICS_FIELDS:
            if name.startswith("init_velocity_") or name in ("gx", "gy", "gz"):
                continue
            setattr(ph, name, float(p[name]))
        ph.init_velocity = (C.c_float * 3)(p["init_velocity_x"], p["init_velocity_y"], p["init_velocity_z"])
        ph.gravity = (C.c_float * 3)(p["gx"], p["gy"], p["gz"])
        s = SceneC()
        s.struct_size = C.sizeof(SceneC)
        s.n_defs = len(scene.user_defs)
        s.defs = defs
        s.n_vertices = vx.size
        s.vein_x, s.vein_y, s.vein_z = _fp(vx), _fp(vy), _fp(vz)
        s.n_triangles = idx.size // 3
        s.vein_indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
        s.n_endings = er.size
        s.ending_centers, s.ending_radii = _fp(ec), _fp(er)
        s.cell_size = (C.c_int32 * 3)(*scene.cell_size)
        s.tri_cell_size = (C.c_int32 * 3)(*scene.tri_cell_size)
        s.use_blood_flow = scene.flags["use_blood_flow"]
        s.enable_reaction_force = scene.flags["enable_reaction_force"]
        s.enable_big_cells_brake = scene.flags["enable_big_cells_brake"]
        s.bounding_spheres_coeff = scene.flags["bounding_spheres_coeff"]
        s.physics = ph
        self.c = s


class Sim:
    """One simulation handle (``bcs_sim*``).  Mirrors the reference loop's vocabulary:
    ``build_grid`` = ``calculateGrid`` x2, ``compute_forces`` = ``calculateNextFrame``,
    ``integrate`` = ``propagateAll`` (main.cu:175-176,199,208)."""

    def __init__(self, scene: Scene, semantics: int = SEM_CLEAN, device: int = 0, use_graph: bool = True,
                 collect_stats: bool = False, seed: int = 1234, lib: Optional[C.CDLL] = None, prefix: str = "bcs_",
                 exhaustive_vein_traversal: bool = False, slab: Optional[dict] = None):
        self.lib = lib if lib is not None else load_library()
        self.prefix = prefix
        self.scene = scene
        self._sh = SceneHandle(scene)
        opts = Opts(C.sizeof(Opts), device, semantics, 1 if use_graph else 0, 1 if collect_stats else 0,
                    1 if exhaustive_vein_traversal else 0, seed, None)
        self._h = C.c_void_p()
        if slab is None:
            self._call("create", C.byref(self._sh.c), C.byref(opts), C.byref(self._h))
        else:
            so = SlabOpts()
            so.struct_size = C.sizeof(SlabOpts)
            so.rank, so.world, so.spawn_rank = slab["rank"], slab["world"], slab.get("spawn_rank", 0)
            so.y_lo, so.y_hi = slab["y_lo"], slab["y_hi"]
            so.halo_width, so.vertex_halo = slab.get("halo_width", 0.0), slab.get("vertex_halo", 0.0)
            so.migration_capacity, so.halo_capacity = slab.get("migration_capacity", 0), slab.get("halo_capacity", 0)
            uid = bytes(slab["nccl_unique_id"])
            assert len(uid) == 128
            so.nccl_unique_id = (C.c_ubyte * 128).from_buffer_copy(uid)   # (a c_char array would stop at the first NUL)
            self._call("create_slab", C.byref(self._sh.c), C.byref(opts), C.byref(so), C.byref(self._h))
        self.slab = slab
        lay = LayoutC()
        self._call("get_layout", self._h, C.byref(lay))
        self.layout = lay
        self.n_particles, self.n_cells = lay.n_particles, lay.n_cells
        self.n_vertices, self.n_triangles = lay.n_vertices, lay.n_triangles

    # -- plumbing
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _call(self, name, *args):
        rc = self._fn(name)(*args)
        if rc != 0:
            err = getattr(self.lib, self.prefix + "last_error")
            err.restype = C.c_char_p
            raise BcsError(f"{self.prefix}{name} failed ({rc}): {err().decode(errors='replace')}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            f = getattr(self.lib, self.prefix + "destroy")
            f.restype = None
            f(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state
    def _len(self, which: int) -> int:
        return {PARTICLE_POS: self.n_particles, PARTICLE_VEL: self.n_particles, PARTICLE_FRC: self.n_particles,
                VEIN_POS: self.n_vertices, VEIN_VEL: self.n_vertices, VEIN_FRC: self.n_vertices,
                CELL_CENTERS: self.n_cells}[which]

    def upload(self, which: int, x, y, z):
        x, y, z = (np.ascontiguousarray(a, np.float32) for a in (x, y, z))
        self._call("upload", self._h, which, _fp(x), _fp(y), _fp(z), C.c_int32(x.size))

    def download(self, which: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        n = self._len(which)
        x, y, z = (np.empty(n, np.float32) for _ in range(3))
        self._call("download", self._h, which, _fp(x), _fp(y), _fp(z), C.c_int32(n))
        return x, y, z

    def upload_owned(self, which: int, x, y, z):
        """Slab mode: only the particles of the blood cells this rank owns cross the bus (bcs_upload_owned)."""
        x, y, z = (np.ascontiguousarray(a, np.float32) for a in (x, y, z))
        self._call("upload_owned", self._h, which, _fp(x), _fp(y), _fp(z), C.c_int32(x.size))

    def download_owned(self, which: int, out=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Slab mode: fills only the entries of the blood cells this rank owns (bcs_download_owned); `out` = three
        full-length float32 arrays to fill in place (other entries are left alone)."""
        n = self._len(which)
        x, y, z = out if out is not None else tuple(np.zeros(n, np.float32) for _ in range(3))
        self._call("download_owned", self._h, which, _fp(x), _fp(y), _fp(z), C.c_int32(n))
        return x, y, z

    def upload_state(self, st: Dict[str, np.ndarray]):
        self.upload(PARTICLE_POS, st["pos_x"], st["pos_y"], st["pos_z"])
        self.upload(PARTICLE_VEL, st["vel_x"], st["vel_y"], st["vel_z"])
        if "frc_x" in st:
            self.upload(PARTICLE_FRC, st["frc_x"], st["frc_y"], st["frc_z"])

    def download_vec(self, which: int) -> np.ndarray:
        return np.stack(self.download(which), axis=1)

    # -- step
    def build_grid(self):
        self._call("build_grid", self._h)

    def compute_forces(self):
        self._call("compute_forces", self._h)

    def integrate(self):
        self._call("integrate", self._h)

    def step(self, n: int = 1):
        self._call("step", self._h, C.c_int32(n))

    def run_stage(self, stage: int):
        self._call("run_stage", self._h, stage)

    def synchronize(self):
        self._call("synchronize", self._h)

    def step_count(self) -> int:
        v = C.c_int64()
        self._call("get_step_count", self._h, C.byref(v))
        return v.value

    def set_step_count(self, steps: int):
        self._call("set_step_count", self._h, C.c_int64(steps))

    def export_frame(self, cell_vertices6: int = 0, offsets3: int = 0, vein_vertices6: int = 0):
        """Device pointers (ints, e.g. torch tensor .data_ptr() or mapped GL buffers) of the renderer's interleaved buffers."""
        self._call("export_frame", self._h, C.c_void_p(cell_vertices6 or None), C.c_void_p(offsets3 or None), C.c_void_p(vein_vertices6 or None))

    # -- checkpoint / restart: the six state arrays + the step count are the whole dynamic state
    _CHECKPOINT_ARRAYS = (("pos", PARTICLE_POS), ("vel", PARTICLE_VEL), ("frc", PARTICLE_FRC),
                          ("vein_pos", VEIN_POS), ("vein_vel", VEIN_VEL), ("vein_frc", VEIN_FRC))

    def checkpoint(self) -> Dict[str, np.ndarray]:
        """Snapshot of the dynamic state as a dict of float32 arrays (+ 'step'); write it with bcsd.write()."""
        out: Dict[str, np.ndarray] = {}
        for name, which in self._CHECKPOINT_ARRAYS:
            x, y, z = self.download(which)
            out[name + "_x"], out[name + "_y"], out[name + "_z"] = x, y, z
        out["step"] = np.array([self.step_count()], dtype=np.int64)
        return out

    def restore(self, ck: Mapping[str, np.ndarray]):
        """Inverse of checkpoint(): the run continues bit-identically (the respawn RNG is keyed by the step count)."""
        for name, which in self._CHECKPOINT_ARRAYS:
            self.upload(which, ck[name + "_x"], ck[name + "_y"], ck[name + "_z"])
        self.set_step_count(int(np.asarray(ck["step"]).ravel()[0]))

    # -- inspection
    def table(self, which: int) -> np.ndarray:
        lay = self.layout
        n, dt = {
            TABLE_SPRING_GRAPH: (lay.n_graph, np.float32), TABLE_MODEL_X: (lay.n_model, np.float32),
            TABLE_MODEL_Y: (lay.n_model, np.float32), TABLE_MODEL_Z: (lay.n_model, np.float32),
            TABLE_COLLISION_RADII: (lay.n_model, np.float32), TABLE_INITIAL_RADII: (lay.n_model, np.float32),
            TABLE_VEIN_NBR_IDS: (BCS_VEIN_MAX_NEIGHBORS * lay.n_vertices, np.int32),
            TABLE_VEIN_NBR_LEN: (BCS_VEIN_MAX_NEIGHBORS * lay.n_vertices, np.float32),
            TABLE_TRI_CENTERS_X: (lay.n_triangles, np.float32), TABLE_TRI_CENTERS_Y: (lay.n_triangles, np.float32),
            TABLE_TRI_CENTERS_Z: (lay.n_triangles, np.float32),
        }[which]
        out = np.empty(n, dt)
        self._call("get_table", self._h, which, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes))
        return out

    def grid(self, which: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        n = self.n_triangles if which else self.n_particles
        keys, ids = np.empty(n, np.int32), np.empty(n, np.int32)
        self._call("download_grid", self._h, which, _ip(keys), _ip(ids), C.c_int32(n))
        return keys, ids

    def cell_table(self, which: int = 0):
        cap = (self.layout.tri_grid_cells if which else min(self.layout.grid_cells, 4 * self.n_particles + 1024))
        while True:
            cells, starts, ends = (np.empty(cap, np.int32) for _ in range(3))
            cnt = C.c_int32()
            rc = self._fn("download_cell_table")(self._h, which, C.c_int32(cap), _ip(cells), _ip(starts), _ip(ends),
                                                C.byref(cnt))
            if rc == 0:
                k = cnt.value
                return cells[:k].copy(), starts[:k].copy(), ends[:k].copy()
            if cnt.value > cap:
                cap = cnt.value
                continue
            self._call("download_cell_table", self._h, which, C.c_int32(cap), _ip(cells), _ip(starts), _ip(ends),
                       C.byref(cnt))

    def debug_candidates(self):
        n = self.n_particles
        counts, hits = np.empty(n, np.int32), np.empty(n, np.int32)
        sums = np.empty(n, np.uint64)
        self._call("debug_candidates", self._h, _ip(counts), sums.ctypes.data_as(C.POINTER(C.c_uint64)), _ip(hits),
                   C.c_int32(n))
        return counts, sums, hits

    def debug_vein_hits(self):
        n = self.n_particles
        tri, t = np.empty(n, np.int32), np.empty(n, np.float32)
        self._call("debug_vein_hits", self._h, _ip(tri), _fp(t), C.c_int32(n))
        return tri, t

    def stats(self) -> Dict[str, int]:
        s = Stats()
        self._call("get_stats", self._h, C.byref(s))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def ownership(self) -> np.ndarray:
        """owned[c] for every blood cell (all ones without slab decomposition)"""
        out = np.empty(self.n_cells, np.uint8)
        self._call("download_ownership", self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int32(self.n_cells))
        return out

    def slab_counts(self):
        a, g, o = C.c_int32(), C.c_int32(), C.c_int32()
        self._call("slab_counts", self._h, C.byref(a), C.byref(g), C.byref(o))
        return {"active_particles": a.value, "ghost_particles": g.value, "owned_cells": o.value}

    def launch_count(self) -> int:
        v = C.c_uint64()
        self._call("get_launch_count", self._h, C.byref(v))
        return v.value

    def profile_steps(self, nsteps: int = 1, capacity: int = 64):
        """Per-kernel device time of ``nsteps`` plain-launch steps: {name: (ms_total, launches)}."""
        names = ((C.c_char * 48) * capacity)()
        ms = (C.c_float * capacity)()
        launches = (C.c_int32 * capacity)()
        cnt = C.c_int32()
        self._call("profile_steps", self._h, C.c_int32(nsteps), C.c_int32(capacity), names, ms, launches, C.byref(cnt))
        return {names[i].value.decode(): (float(ms[i]), int(launches[i])) for i in range(cnt.value)}

    def device_view(self) -> DeviceView:
        v = DeviceView()
        self._call("device_ptrs", self._h, C.byref(v))
        return v

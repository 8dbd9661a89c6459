"""Host plumbing of the multi-GPU (slab) mode: one process per GPU, `torch.distributed` for rendezvous only.

The data path (ghost particles, migrating blood cells, vein-vertex halo) is NCCL point-to-point inside libbcs
(csrc/slab.cu); torch.distributed just carries the 128-byte ncclUniqueId to all ranks and gathers results for
checks.  The pure-host parts (slab boundaries, ownership merge) run on CPU tensors, so they are testable with the
gloo backend.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi
from .scene import Layout, Scene


def cell_centres_y(scene: Scene, state: Dict[str, np.ndarray], layout: Optional[Layout] = None) -> np.ndarray:
    lay = layout or scene.layout()
    out = np.empty(lay.n_cells, np.float32)
    for t in range(lay.n_types):
        cnt, p, ps, cs = int(lay.counts[t]), int(lay.particles_in_cell[t]), int(lay.particle_starts[t]), int(lay.cell_starts[t])
        out[cs:cs + cnt] = state["pos_y"][ps:ps + cnt * p].reshape(cnt, p).mean(axis=1)
    return out


def slab_boundaries(scene: Scene, state: Dict[str, np.ndarray], world: int, layout: Optional[Layout] = None) -> List[float]:
    """world+1 descending y planes: rank r owns blood cells with centre in [b[r+1], b[r]).  Interior planes sit on
    grid-cell planes (multiples of the cell height above the grid origin) and split the blood cells evenly; the
    outermost planes are +-inf (flow is towards -y: config/physics.hpp:25,33)."""
    lay = layout or scene.layout()
    cy = np.sort(cell_centres_y(scene, state, lay))[::-1]
    h = float(scene.cell_size[1])
    y0 = float(lay.grid_min[1])
    planes = [float("inf")]
    for r in range(1, world):
        q = cy[min(len(cy) - 1, (len(cy) * r) // world)]
        planes.append(y0 + round((float(q) - y0) / h) * h)
    planes.append(float("-inf"))
    for a, b in zip(planes[:-1], planes[1:]):
        if not a > b:
            raise ValueError("degenerate slab decomposition (too many ranks for this scene)")
    return planes


def spawn_rank(scene: Scene, planes: Sequence[float]) -> int:
    y = float(scene.physics["min_spawn_y"])
    for r in range(len(planes) - 1):
        if planes[r + 1] <= y < planes[r]:
            return r
    return 0


def broadcast_unique_id(rank: int, lib=None) -> bytes:
    """ncclUniqueId from rank 0 to everyone over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        lib = lib or capi.load_library()
        raw = (C.c_char * 128)()
        rc = lib.bcs_nccl_unique_id(raw)
        if rc != 0:
            raise capi.BcsError("bcs_nccl_unique_id failed")
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = buf.to(dev)
        dist.broadcast(t, src=0)
        buf = t.cpu()
    else:
        dist.broadcast(buf, src=0)
    return bytes(buf.numpy().tobytes())


def create_slab_sim(scene: Scene, state: Dict[str, np.ndarray], rank: int, world: int, device: int, unique_id: bytes,
                    planes: Optional[Sequence[float]] = None, **kw) -> capi.Sim:
    planes = list(planes) if planes is not None else slab_boundaries(scene, state, world)
    slab = dict(rank=rank, world=world, spawn_rank=spawn_rank(scene, planes), y_lo=planes[rank + 1], y_hi=planes[rank],
                nccl_unique_id=unique_id)
    slab.update({k: kw.pop(k) for k in ("halo_width", "vertex_halo", "migration_capacity", "halo_capacity") if k in kw})
    sim = capi.Sim(scene, device=device, slab=slab, **kw)
    sim.upload_state(state)
    return sim


def merge_owned(per_rank_arrays: Sequence[np.ndarray], per_rank_owned: Sequence[np.ndarray], layout: Layout) -> np.ndarray:
    """Assemble a global per-particle array from per-rank copies: every blood cell is taken from the rank that owns it.
    Raises if a blood cell is owned by no rank or by several."""
    owners = np.stack([np.asarray(o, np.uint8) for o in per_rank_owned])
    count = owners.sum(axis=0)
    if not np.all(count == 1):
        raise ValueError(f"ownership is not a partition: {int((count == 0).sum())} orphan and {int((count > 1).sum())} duplicated blood cells")
    who = owners.argmax(axis=0)
    out = np.empty_like(per_rank_arrays[0])
    for t in range(layout.n_types):
        cnt, p, ps, cs = int(layout.counts[t]), int(layout.particles_in_cell[t]), int(layout.particle_starts[t]), int(layout.cell_starts[t])
        w = np.repeat(who[cs:cs + cnt], p)
        for r in range(len(per_rank_arrays)):
            sel = np.where(w == r)[0] + ps
            out[sel] = per_rank_arrays[r][sel]
    return out

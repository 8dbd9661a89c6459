"""CPU, world_size 2, gloo backend: the host-side plumbing of the slab (multi-GPU) mode - slab boundaries, spawn rank,
ncclUniqueId hand-off, ownership merge.  The data path itself (NCCL p2p inside libbcs) needs GPUs: tests/mgpu_check.py."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, golden_scene, pkg, seeded_state

WORKER = textwrap.dedent('''
    import importlib, os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["BCS_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BCS_ROOT"], "tests"))
    from conftest import golden_scene, seeded_state, pkg
    dd = importlib.import_module("simulation-server_b200.distributed")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sc = golden_scene("cfg1"); st, _ = seeded_state("cfg1", "spawn"); lay = sc.layout()
    planes = dd.slab_boundaries(sc, st, world, lay)
    all_planes = [None] * world
    dist.all_gather_object(all_planes, planes)
    assert all(p == planes for p in all_planes), "ranks disagree on the slab planes"
    assert planes[0] == float("inf") and planes[-1] == float("-inf") and all(a > b for a, b in zip(planes, planes[1:]))
    h, y0 = sc.cell_size[1], float(lay.grid_min[1])
    for p in planes[1:-1]:
        assert abs((p - y0) / h - round((p - y0) / h)) < 1e-4, "interior planes must sit on grid-cell planes"
    assert dd.spawn_rank(sc, planes) == 0            # minSpawnY = -20 is at the top of the vein
    # every rank "simulates" its slab: here, moves its own cells and leaves the others untouched
    cy = dd.cell_centres_y(sc, st, lay)
    owned = ((cy >= planes[rank + 1]) & (cy < planes[rank])).astype(np.uint8)
    counts = [None] * world
    dist.all_gather_object(counts, int(owned.sum()))
    assert sum(counts) == lay.n_cells and min(counts) > 0.3 * lay.n_cells / world, counts
    pos = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1).copy()
    pmask = np.zeros(lay.n_particles, bool)
    for t in range(lay.n_types):
        cnt, p, ps, cs = int(lay.counts[t]), int(lay.particles_in_cell[t]), int(lay.particle_starts[t]), int(lay.cell_starts[t])
        pmask[ps:ps + cnt * p] = np.repeat(owned[cs:cs + cnt].astype(bool), p)
    pos[pmask] += 1.0 + rank
    pos[~pmask] = np.nan                              # stale / foreign data must never be picked up by the merge
    gathered = [None] * world
    dist.all_gather_object(gathered, (pos, owned))
    if rank == 0:
        merged = dd.merge_owned([g[0] for g in gathered], [g[1] for g in gathered], lay)
        assert np.isfinite(merged).all()
        orig = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1)
        shift = (merged - orig)[:, 0]
        assert set(np.round(shift).astype(int).tolist()) == {1, 2}
        try:
            dd.merge_owned([g[0] for g in gathered], [gathered[0][1], gathered[0][1]], lay)
            raise SystemExit("merge_owned accepted an ownership that is not a partition")
        except ValueError:
            pass
    # the 128-byte ncclUniqueId travels intact (it contains NUL bytes) from rank 0 to everyone
    try:
        uid = dd.broadcast_unique_id(rank)
    except Exception as e:                            # no usable NCCL bootstrap on this machine
        uid = None
        print("unique id not available:", e)
    if uid is not None:
        ids = [None] * world
        dist.all_gather_object(ids, uid)
        assert len(uid) == 128 and all(i == uid for i in ids) and any(b != 0 for b in uid)
    dist.barrier()
    dist.destroy_process_group()
    print("WORKER_OK", rank)
''')


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_slab_host_logic_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BCS_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("WORKER_OK") == 2


def test_slab_boundaries_balance_and_errors():
    sc = golden_scene("cfg1")
    st, _ = seeded_state("cfg1", "spawn")
    dd = __import__("importlib").import_module("simulation-server_b200.distributed")
    for world in (1, 2, 4, 8):
        planes = dd.slab_boundaries(sc, st, world)
        assert len(planes) == world + 1
        cy = dd.cell_centres_y(sc, st)
        counts = [int(((cy >= planes[r + 1]) & (cy < planes[r])).sum()) for r in range(world)]
        assert sum(counts) == 600 and max(counts) - min(counts) <= 0.2 * 600 / world + 8, counts
    try:
        dd.slab_boundaries(sc, st, 512)   # 180 units of vein cannot be cut into 512 slabs on 2-unit planes
        assert False
    except ValueError:
        pass

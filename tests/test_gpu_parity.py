"""GPU (B200): parity of libbcs.so - called through its C ABI - with (a) dumps of the reference CUDA build
(reference-compatible semantics) and (b) the CPU oracle (clean and reference-compatible semantics)."""
import importlib

import numpy as np
import pytest

import refcheck
import stagecheck
from conftest import (capi, golden_dump, golden_scene, make_bcs, make_oracle, pkg, seeded_state, small_cylinder_scene)
from test_oracle_vs_reference import CASES, TRAJECTORY_CASES, check_trajectory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["cfg1", "mini3", "cfg2"])
def test_derived_tables(bcs_lib, oracle_lib, cfg):
    sc = golden_scene(cfg)
    with make_bcs(sc) as sim, make_oracle(oracle_lib, sc) as orc:
        a, b = sim.layout, orc.layout
        assert bytes(a) == bytes(b)
        for t in range(11):
            assert np.array_equal(sim.table(t), orc.table(t)), f"table {t}"


@pytest.mark.parametrize("cfg,variant,steps,vein_steps", CASES)
def test_stage_by_stage_vs_reference(bcs_lib, cfg, variant, steps, vein_steps):
    sc = golden_scene(cfg)
    with make_bcs(sc, capi.SEM_REFERENCE) as sim:
        summary = refcheck.replay_steps(sim, cfg, variant, steps, golden_dump, sc, sc.physics, vein_steps)
    if variant == "wide":
        assert summary[1]["vein_hits"] > 20


def test_candidate_sets_vs_reference_grid(bcs_lib):
    sc = golden_scene("mini3")
    with make_bcs(sc, capi.SEM_REFERENCE) as sim:
        for step in (1, 2, 3):
            d = golden_dump("mini3", "wide", step)
            refcheck.up(sim, capi.PARTICLE_POS, refcheck.vec(d, "begin.pos"))
            refcheck.up(sim, capi.PARTICLE_VEL, refcheck.vec(d, "begin.vel"))
            sim.run_stage(capi.STAGE_GRID_PARTICLES)
            cnt, chk, _ = sim.debug_candidates()
            ecnt, echk = refcheck.expected_candidates(d, sim.layout)
            assert np.array_equal(cnt, ecnt)
            assert np.array_equal(chk, echk)


@pytest.mark.parametrize("cfg,variant", TRAJECTORY_CASES)
def test_100_step_trajectory_vs_reference(bcs_lib, cfg, variant):
    sc = golden_scene(cfg)
    st, _ = seeded_state(cfg, variant)
    with make_bcs(sc, capi.SEM_REFERENCE) as sim:
        sim.upload_state(st)
        sim.step(100)
        pos = refcheck.down(sim, capi.PARTICLE_POS)
        assert sim.step_count() == 100
    check_trajectory(pos, cfg, variant, st)


@pytest.mark.parametrize("semantics", [capi.SEM_CLEAN, capi.SEM_REFERENCE])
@pytest.mark.parametrize("cfg,variant", [("mini3", "wide"), ("cfg1", "wide"), ("cfg1", "spawn")])
def test_vs_oracle_default_vein(bcs_lib, oracle_lib, cfg, variant, semantics):
    sc = golden_scene(cfg)
    st, _ = seeded_state(cfg, variant)
    with make_bcs(sc, semantics) as sim, make_oracle(oracle_lib, sc, semantics) as orc:
        orc.upload_state(st)
        stagecheck.compare_step(sim, orc, sc, 6, f"{cfg}/{variant}/sem{semantics}")


def test_vs_oracle_cylinder_scene(bcs_lib, oracle_lib):
    sc = small_cylinder_scene()
    st = pkg.make_initial_state(sc, seed=7, xz_half_width=49.0, y_range=(-25.0, -110.0))
    with make_bcs(sc) as sim, make_oracle(oracle_lib, sc) as orc:
        orc.upload_state(st)
        stagecheck.compare_step(sim, orc, sc, 8, "cylinder")


@pytest.mark.parametrize("semantics,wall_margin", [(capi.SEM_CLEAN, None), (capi.SEM_CLEAN, "0.0002"), (capi.SEM_REFERENCE, None)])
@pytest.mark.parametrize("which", ["cfg1_wide", "cylinder"])
def test_culled_vein_search_equals_exhaustive_traversal(bcs_lib, which, semantics, wall_margin, monkeypatch):
    """The production vein-collision kernels (clean semantics: lazily rebuilt wall grid, wall.cu; reference
    semantics: segment / half-line culling over the per-step refitted box hierarchy, vein.cu) must act on exactly
    the particles, with exactly the triangle, the reference's exhaustive in-order traversal ends on - including
    far triangles that mask a near hit.  Bitwise equal particle outputs.  A tiny wall margin forces
    the wall grid through its rebuild path every few steps."""
    if wall_margin:
        monkeypatch.setenv("BCS_WALL_MARGIN", wall_margin)
    if which == "cfg1_wide":
        sc = golden_scene("cfg1")
        st, _ = seeded_state("cfg1", "wide")
    else:
        sc = small_cylinder_scene(300, 300, 400.0)
        st = pkg.make_initial_state(sc, seed=11, xz_half_width=50.0, y_range=(-25.0, -360.0))
    total_hits = 0
    with make_bcs(sc, semantics) as fast, make_bcs(sc, semantics, exhaustive_vein_traversal=True) as slow:
        fast.upload_state(st)
        for step in range(25):
            for which_arr in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.VEIN_FRC):
                refcheck.up(slow, which_arr, refcheck.down(fast, which_arr))
            tri, t = slow.debug_vein_hits()
            total_hits += int(((tri >= 0) & (t <= 6.0)).sum())
            fast.run_stage(capi.STAGE_VEIN_COLLISIONS)
            slow.run_stage(capi.STAGE_VEIN_COLLISIONS)
            for which_arr in (capi.PARTICLE_FRC, capi.PARTICLE_VEL):
                a, b = refcheck.down(fast, which_arr), refcheck.down(slow, which_arr)
                assert np.array_equal(a, b), f"step {step}: culled and exhaustive vein collision differ for {(a != b).any(axis=1).sum()} particles"
            refcheck.assert_close(refcheck.down(fast, capi.VEIN_FRC), refcheck.down(slow, capi.VEIN_FRC), "vein force splats", rtol=1e-5,
                                  scale=1.0)
            fast.step(1)
        if semantics == capi.SEM_CLEAN:
            builds = fast.stats()["wall_rebuilds"]
            assert builds >= 1 and (not wall_margin or builds > 3), builds
    assert total_hits > 100


@pytest.mark.parametrize("semantics", [capi.SEM_CLEAN, capi.SEM_REFERENCE])
def test_fused_step_equals_staged_step(bcs_lib, semantics):
    """bcs_step (fused tail kernel, culled vein search, graph replay) vs the nine staged entry points in the
    reference's order, incl. vein-end teleports (a short vein so that blood cells reach the ending)."""
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=5, xz_half_width=40.0, y_range=(-25.0, -95.0))
    with make_bcs(sc, semantics) as a, make_bcs(sc, semantics) as b:
        a.upload_state(st)
        b.upload_state(st)
        for step in range(60):
            a.step(1)
            for stage in range(9):
                b.run_stage(stage)
            if step % 10 == 9:
                for which, name in ((capi.PARTICLE_POS, "pos"), (capi.PARTICLE_VEL, "vel"), (capi.PARTICLE_FRC, "frc"), (capi.VEIN_POS, "vein pos")):
                    # the fused tail evaluates the same expressions as the staged kernels: bitwise equal state
                    assert np.array_equal(refcheck.down(a, which), refcheck.down(b, which)), f"step {step} {name}: fused and staged step differ"
        assert a.step_count() == b.step_count() == 60
        assert a.stats()["teleported_cells"] == b.stats()["teleported_cells"] > 0


def test_graph_and_plain_launch_agree(bcs_lib):
    sc = golden_scene("mini3")
    st, _ = seeded_state("mini3", "wide")
    out = []
    for use_graph in (True, False):
        with make_bcs(sc, use_graph=use_graph) as sim:
            sim.upload_state(st)
            sim.step(20)
            out.append(refcheck.down(sim, capi.PARTICLE_POS))
    assert np.array_equal(out[0], out[1]), "graph replay and plain launches must give the same bits"


def test_headless_cpp_driver_matches_python_path(bcs_lib, tmp_path):
    """The C++ host mirror (same call names and order as main.cu:175-208) driven by bcs_headless vs bcs_step from Python."""
    import os
    import subprocess
    from conftest import ROOT
    sc = golden_scene("mini3")
    st, _ = seeded_state("mini3", "wide")
    scene_path, state_path, out_path = (str(tmp_path / n) for n in ("scene.bcsd", "state.bcsd", "out.bcsd"))
    sc.save(scene_path)
    pkg.bcsd.write(state_path, st)
    exe = os.path.join(ROOT, "simulation-server_b200", "bcs_headless")
    r = subprocess.run([exe, scene_path, state_path, "25", out_path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Average framerate" in r.stdout
    out = pkg.bcsd.read(out_path)
    with make_bcs(sc) as sim:
        sim.upload_state(st)
        sim.step(25)
        pos = refcheck.down(sim, capi.PARTICLE_POS)
    assert np.array_equal(np.stack([out["pos_x"], out["pos_y"], out["pos_z"]], 1), pos), "C++ headless loop vs bcs_step: same library, same bits"


def test_checkpoint_restart_is_bit_identical(bcs_lib, tmp_path):
    """state arrays + step count are the whole dynamic state (counter-based respawn RNG) - a run restored
    from a BCSD checkpoint continues bit-identically, vein-end teleports included."""
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=5, xz_half_width=40.0, y_range=(-25.0, -95.0))
    path = str(tmp_path / "ck.bcsd")
    with make_bcs(sc) as a:
        a.upload_state(st)
        a.step(30)
        pkg.bcsd.write(path, a.checkpoint())
        a.step(30)
        ref = {w: refcheck.down(a, w) for w in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS)}
        tele = a.stats()["teleported_cells"]
    with make_bcs(sc) as b:
        b.restore(pkg.bcsd.read(path))
        assert b.step_count() == 30
        b.step(30)
        for w, want in ref.items():
            assert np.array_equal(refcheck.down(b, w), want), f"array {w} differs after restart"
    assert tele > 0


def test_free_run_is_bitwise_repeatable(bcs_lib):
    """Wall splats are summed in fixed point (order-independent), every other stage is free of float atomics: the same
    initial state gives the same bits, run after run, graph replay with forked branches included."""
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=5, xz_half_width=40.0, y_range=(-25.0, -95.0))
    arrays = (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.VEIN_FRC)
    want = None
    for run in range(12):
        with make_bcs(sc) as a:
            a.upload_state(st)
            a.step(60)
            got = [refcheck.down(a, w) for w in arrays]
            hits = a.stats()["vein_hits"]
        assert hits > 50
        if want is None:
            want = got
        for w, x, y in zip(arrays, got, want):
            assert np.array_equal(x, y), f"run {run}: array {w} differs from run 0"


@pytest.mark.parametrize("switch", ["BCS_GRID=radix", "BCS_GRID=cells", "BCS_GRID=cells,BCS_COLLIDE=rows", "BCS_SCAN=fused", "BCS_COLLIDE=walk",
                                    "BCS_NO_NEAR_PROBE=1", "BCS_NO_OVERLAP=1", "BCS_SORT=classic", "BCS_NO_FUSE=1", "BCS_NO_FUSE=1,BCS_NO_OVERLAP=1",
                                    "BCS_SPRING_G=3", "BCS_SPRING_WARPS=2"])
def test_alternative_paths_equal_the_default_bitwise(bcs_lib, monkeypatch, switch):
    """every A/B switch of DESIGN.md section 7 selects another route to the SAME result: compact cell index instead of the
    row directory (counting or radix sorted), per-slot stencil walks instead of the symmetric pair search, single-launch
    scans, wall filter without the near list, unforked step, unfused run, other blood-cell group sizes, directed springs"""
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=8, xz_half_width=44.0, y_range=(-25.0, -95.0))
    arrays = (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL)

    def run():
        with make_bcs(sc) as sim:
            sim.upload_state(st)
            sim.step(25)
            return [refcheck.down(sim, w) for w in arrays], sim.stats()

    want, stats = run()
    assert stats["vein_hits"] > 20
    for one in switch.split(","):
        name, value = one.split("=")
        monkeypatch.setenv(name, value)
    got, _ = run()
    for w, x, y in zip(arrays, got, want):
        assert np.array_equal(x, y), f"{switch}: array {w} differs from the default path"


def test_reaction_force_disabled_vs_oracle(bcs_lib, oracle_lib):
    """the `enableReactionForce = false` branch of the wall collision (vein_collisions.cu:248-252) - the
    particle keeps its force and only its velocity is reflected.  libbcs vs oracle stage by stage, and the branch must
    actually change the forces of the particles that hit the wall."""
    sc = small_cylinder_scene()
    sc.flags["enable_reaction_force"] = 0
    st = pkg.make_initial_state(sc, seed=7, xz_half_width=49.0, y_range=(-25.0, -110.0))
    with make_bcs(sc) as sim, make_oracle(oracle_lib, sc) as orc:
        orc.upload_state(st)
        summary = stagecheck.compare_step(sim, orc, sc, 6, "no reaction force")
        assert summary["vein_hits"] > 20, summary
        off = refcheck.down(sim, capi.PARTICLE_FRC)
    sc.flags["enable_reaction_force"] = 1
    with make_bcs(sc) as sim:
        sim.upload_state(st)
        sim.step(6)
        on = refcheck.down(sim, capi.PARTICLE_FRC)
    assert (np.abs(on - off).max(axis=1) > 1e-3).sum() > 10, "the reaction-force switch changed nothing"


def test_fused_run_equals_single_steps(bcs_lib):
    """bcs_step(n) in row-directory mode runs the end of step k and the springs / row count of step k + 1 as one pass over
    the particle state (cellpass.cu: launch_advance); n calls of bcs_step(1) never fuse.  Same bits, teleports included."""
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=5, xz_half_width=40.0, y_range=(-25.0, -95.0))
    arrays = (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.CELL_CENTERS)
    with make_bcs(sc) as a, make_bcs(sc) as b:
        a.upload_state(st)
        b.upload_state(st)
        for block in (1, 2, 7, 30, 20):
            a.step(block)
            for _ in range(block):
                b.step(1)
            for w in arrays:
                assert np.array_equal(refcheck.down(a, w), refcheck.down(b, w)), f"after a fused run of {block}: array {w} differs"
        assert a.step_count() == b.step_count() == 60
        assert a.stats()["teleported_cells"] == b.stats()["teleported_cells"] > 0


@pytest.mark.parametrize("fuse", [True, False])
def test_single_rank_slab_handle_equals_plain_handle(bcs_lib, monkeypatch, fuse):
    """A slab-decomposition handle with world size 1 owns every blood cell: it runs the slab code paths (owned-cell lists,
    row directory over a window of cell rows, deferred fold, the fused run with the end-of-step chores riding on the cell
    pass) without a neighbour, and must reproduce the plain handle bit for bit - respawns at the vein end included."""
    if not fuse:
        monkeypatch.setenv("BCS_SLAB_NO_FUSE", "1")
    dd = importlib.import_module("simulation-server_b200.distributed")
    sc = small_cylinder_scene(120, 100, 120.0)
    st = pkg.make_initial_state(sc, seed=5, xz_half_width=40.0, y_range=(-25.0, -95.0))
    arrays = (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.CELL_CENTERS)
    with make_bcs(sc) as a:
        b = dd.create_slab_sim(sc, st, 0, 1, 0, bytes(128), seed=1234)
        try:
            a.upload_state(st)
            for block in (1, 2, 7, 30, 20):
                a.step(block)
                b.step(block)
                for w in arrays:
                    assert np.array_equal(refcheck.down(a, w), refcheck.down(b, w)), f"after a run of {block}: array {w} differs"
            assert a.stats()["teleported_cells"] == b.stats()["teleported_cells"] > 0
            assert b.slab_counts()["owned_cells"] == a.n_cells
            # owned-only transfers (bcs_upload_owned / bcs_download_owned): this rank owns everything, so they move everything;
            # an upload of positions makes the next step refresh the halo first - with the same state the run must not change
            import torch
            for w in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC):
                got = b.download_owned(w)   # pageable arrays: host-side gather + staging buffer
                assert all(np.array_equal(g, f) for g, f in zip(got, b.download(w)))
                # pinned arrays: the kernels read / write the host arrays in place
                pinned = tuple(torch.zeros(b.n_particles, dtype=torch.float32).pin_memory().numpy() for _ in range(3))
                b.download_owned(w, out=pinned)
                assert all(np.array_equal(g, f) for g, f in zip(got, pinned))
                b.upload_owned(w, *(pinned if w != capi.PARTICLE_VEL else got))
                b.synchronize()   # a pinned upload is asynchronous: the arrays must stay untouched until the stream has read them
                a.upload(w, *got)
            a.step(12)
            b.step(12)
            for w in arrays:
                assert np.array_equal(refcheck.down(a, w), refcheck.down(b, w)), f"after owned-only transfers: array {w} differs"
        finally:
            b.close()


def test_particles_outside_the_grid_take_the_full_walk(bcs_lib, oracle_lib, monkeypatch):
    """A particle outside the grid has a one-sided stencil (particle_collisions.cuh:126-268 trims by the unclamped index),
    so the symmetric pair search hands the whole build to the per-slot walk over the row directory.  Candidate sets and
    forces must equal the oracle's and the compact-cell-index path's."""
    sc = small_cylinder_scene(60, 40, 150.0)
    sc.flags["use_blood_flow"] = 0
    st = pkg.make_initial_state(sc, seed=7, xz_half_width=49.0, y_range=(-25.0, -110.0))
    lay = sc.layout()
    # push three blood cells out of the grid on different sides (whole cells, so that their springs stay sane)
    for cell, (dx, dy, dz) in ((3, (400.0, 0.0, 0.0)), (47, (0.0, 900.0, 0.0)), (80, (-300.0, 0.0, -300.0))):
        t = int(np.searchsorted(np.asarray(lay.cell_starts[:lay.n_types]), cell, side="right") - 1)
        p, ps, cs = int(lay.particles_in_cell[t]), int(lay.particle_starts[t]), int(lay.cell_starts[t])
        sl = slice(ps + (cell - cs) * p, ps + (cell - cs + 1) * p)
        st["pos_x"][sl] += np.float32(dx); st["pos_y"][sl] += np.float32(dy); st["pos_z"][sl] += np.float32(dz)
    results = []
    for mode in ("rows", "cells"):
        monkeypatch.setenv("BCS_GRID", mode)
        with make_bcs(sc) as sim, make_oracle(oracle_lib, sc) as orc:
            sim.upload_state(st)
            orc.upload_state(st)
            for step in range(3):
                sim.run_stage(capi.STAGE_GRID_PARTICLES); orc.run_stage(capi.STAGE_GRID_PARTICLES)
                (ka, ia), (kb, ib) = sim.grid(0), orc.grid(0)
                assert np.array_equal(ka, kb) and np.array_equal(ia, ib), f"{mode} step {step}: grid"
                for x, y in zip(sim.debug_candidates(), orc.debug_candidates()):
                    assert np.array_equal(x, y), f"{mode} step {step}: candidate sets"
                sim.step(1); orc.step(1)
            assert sim.stats()["out_of_bounds"] > 0
            results.append([refcheck.down(sim, w) for w in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC)])
            refcheck.assert_close(results[-1][0], refcheck.down(orc, capi.PARTICLE_POS), f"{mode}: positions vs oracle", scale=0.0)
    for x, y in zip(*results):
        assert np.array_equal(x, y), "row-directory walk and compact-cell-index walk differ"


def test_tiled_collision_kernel_equals_index_walk(bcs_lib, monkeypatch):
    """BCS_COLLIDE=tiled (neighbour windows staged in shared memory per tile of sorted slots) visits the same candidates
    in the same order as the index walk: bitwise equal forces, equal debug candidate sets."""
    sc = small_cylinder_scene(300, 300, 400.0)
    st = pkg.make_initial_state(sc, seed=11, xz_half_width=45.0, y_range=(-25.0, -360.0))
    out = []
    monkeypatch.setenv("BCS_GRID", "cells")   # the per-slot walks live on the compact cell index
    for mode in (None, "tiled"):
        if mode:
            monkeypatch.setenv("BCS_COLLIDE", mode)
        with make_bcs(sc) as sim:
            sim.upload_state(st)
            sim.step(12)
            sim.build_grid()
            out.append((refcheck.down(sim, capi.PARTICLE_FRC), sim.debug_candidates()))
    assert np.array_equal(out[0][0], out[1][0])
    for x, y in zip(out[0][1], out[1][1]):
        assert np.array_equal(x, y)
    assert out[0][1][2].sum() > 0   # some pairs actually collide


def test_frame_export_layouts(bcs_lib):
    """bcs_export_frame fills the renderer's interleaved buffers (glcontroller.cu:23-50): stride 6 with the normal slots
    untouched, stride 3 offsets, stride 6 vein vertices."""
    import torch
    sc = golden_scene("mini3")
    st, _ = seeded_state("mini3", "wide")
    with make_bcs(sc) as sim:
        sim.upload_state(st)
        sim.step(3)
        n, v = sim.n_particles, sim.n_vertices
        dev = torch.device("cuda", 0)
        a = torch.full((6 * n,), -7.0, device=dev)
        b = torch.full((3 * n,), -7.0, device=dev)
        c = torch.full((6 * v,), -7.0, device=dev)
        sim.export_frame(a.data_ptr(), b.data_ptr(), c.data_ptr())
        sim.synchronize()
        pos, vpos = refcheck.down(sim, capi.PARTICLE_POS), refcheck.down(sim, capi.VEIN_POS)
        a, b, c = a.cpu().numpy().reshape(n, 6), b.cpu().numpy().reshape(n, 3), c.cpu().numpy().reshape(v, 6)
        assert np.array_equal(a[:, :3], pos) and np.all(a[:, 3:] == -7.0)
        assert np.array_equal(b, pos)
        assert np.array_equal(c[:, :3], vpos) and np.all(c[:, 3:] == -7.0)

"""CPU: the C-ABI surface of libbcs.so (loads, exports every symbol of include/bcs.h, struct layouts agree with
the ctypes binding, fails loudly without a GPU) and the host-side helpers (BCSD files, scene layout, generators).
No compute call is made on the product library here."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, capi, golden_scene, has_gpu, make_oracle, pkg, small_cylinder_scene

HEADER = os.path.join(ROOT, "include", "bcs.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcs_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load_library()
    names = declared_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"libbcs.so does not export {n} declared in include/bcs.h"
    lib.bcs_abi_version.restype = ctypes.c_int
    assert lib.bcs_abi_version() == 1


def test_ctypes_structs_match_the_c_header():
    """sizeof/offsetof as the C compiler sees include/bcs.h vs the ctypes mirror in capi.py"""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "bcs.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(bcs_spring), sizeof(bcs_cell_def), sizeof(bcs_physics), sizeof(bcs_scene),
               sizeof(bcs_opts), sizeof(bcs_type_info), sizeof(bcs_layout), sizeof(bcs_device_view), sizeof(bcs_stats));
        printf("%zu %zu %zu %zu\n", offsetof(bcs_scene, physics), offsetof(bcs_scene, cell_size), offsetof(bcs_opts, seed), offsetof(bcs_layout, types));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(v) for v in out]
    mirror = [capi.Spring, capi.CellDefC, capi.Physics, capi.SceneC, capi.Opts, capi.TypeInfo, capi.LayoutC, capi.DeviceView, capi.Stats]
    assert sizes[:9] == [ctypes.sizeof(m) for m in mirror]
    assert sizes[9:] == [capi.SceneC.physics.offset, capi.SceneC.cell_size.offset, capi.Opts.seed.offset, capi.LayoutC.types.offset]


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_create_fails_loudly_without_a_gpu():
    with pytest.raises(capi.BcsError) as e:
        capi.Sim(small_cylinder_scene())
    assert "no CPU fallback" in str(e.value)


def test_missing_library_is_an_error_not_a_fallback():
    with pytest.raises(capi.BcsError):
        capi.load_library("/nonexistent/libbcs.so")


def test_bcsd_roundtrip(tmp_path):
    a = {"f": np.arange(5, dtype=np.float32), "i": np.arange(7, dtype=np.int32), "u": np.arange(3, dtype=np.uint32),
         "d": np.arange(2, dtype=np.float64), "l": np.arange(4, dtype=np.int64), "empty": np.zeros(0, np.float32)}
    p = str(tmp_path / "x.bcsd")
    pkg.bcsd.write(p, a)
    b = pkg.bcsd.read(p)
    assert list(b) == list(a)
    for k in a:
        assert b[k].dtype == a[k].dtype and np.array_equal(a[k], b[k])


def test_scene_save_load_roundtrip(tmp_path):
    sc = golden_scene("mini3")
    p = str(tmp_path / "s.bcsd")
    sc.save(p)
    sc2 = pkg.Scene.load(p)
    assert len(sc2.user_defs) == len(sc.user_defs)
    for a, b in zip(sc.user_defs, sc2.user_defs):
        assert a.count == b.count and a.same_type(b) and np.array_equal(a.vertices, b.vertices)
    assert np.array_equal(sc.vein_pos, sc2.vein_pos) and np.array_equal(sc.vein_indices, sc2.vein_indices)
    assert sc.physics == sc2.physics and sc.flags == sc2.flags


def test_cylinder_vein_uses_the_reference_tessellation():
    vp, vi, ec, er = pkg.make_cylinder_vein(length=50.0)
    assert vp.shape == (1100, 3) and vi.shape == (2000, 3)
    # config/vein_definition.hpp:15965: "100, 0, 1, 100, 1, 101, 101, 1, 2, 101, 2, 102, ..."
    assert vi[:4].reshape(-1).tolist() == [100, 0, 1, 100, 1, 101, 101, 1, 2, 101, 2, 102]
    assert vi[198:200].reshape(-1).tolist() == [199, 99, 0, 199, 0, 100]          # ring wrap-around
    assert np.allclose(np.hypot(vp[:, 0], vp[:, 2]), 50.0, atol=1e-4) and vp[:, 1].min() == -50.0


@pytest.mark.parametrize("seed", range(6))
def test_layout_mirror_agrees_with_oracle_on_random_type_lists(oracle_lib, seed):
    """fold / unique / mp_sort emulation (scene.derive_layout) vs the oracle's, on random user lists that mix
    power-of-two and other cell sizes and repeat definitions."""
    rng = np.random.default_rng(seed)
    base = golden_scene("mini3")
    protos = []
    for p in (4, 8, 5, 12, 16, 20):
        verts = rng.normal(size=(p, 3)).astype(np.float32) * 3
        pairs = [(a, b) for a in range(p) for b in range(a + 1, p) if rng.random() < 0.6] or [(0, 1)]
        springs = np.asarray(pairs, np.int32)
        lens = np.linalg.norm(verts[springs[:, 0]] - verts[springs[:, 1]], axis=1).astype(np.float32)
        protos.append((p, springs, lens, verts))
    defs = []
    for _ in range(int(rng.integers(2, 7))):
        p, springs, lens, verts = protos[int(rng.integers(len(protos)))]
        defs.append(pkg.CellDef(int(rng.integers(1, 9)), p, springs, lens, verts))
    sc = pkg.Scene(user_defs=defs, vein_pos=base.vein_pos, vein_indices=base.vein_indices, ending_centers=base.ending_centers,
                   ending_radii=base.ending_radii)
    lay = sc.layout()
    with make_oracle(oracle_lib, sc) as orc:
        L = orc.layout
        assert L.n_types == lay.n_types and L.n_particles == lay.n_particles
        got = [(t.count, t.particles_in_cell, t.particle_start, t.cell_start, t.model_start, t.graph_start, t.src_def) for t in L.types[:L.n_types]]
        exp = list(zip(lay.counts.tolist(), lay.particles_in_cell.tolist(), lay.particle_starts.tolist(), lay.cell_starts.tolist(),
                       lay.model_starts.tolist(), lay.graph_starts.tolist(), lay.src_def))
        assert got == exp
        assert np.array_equal(orc.table(capi.TABLE_SPRING_GRAPH), lay.spring_graph)
        # the reference would pick the warp-sync vein-end kernel for power-of-two cell sizes (vein_end.cu:23-30)
        for t in L.types[:L.n_types]:
            assert t.vein_end_warp_sync == int(t.count * t.particles_in_cell <= 32 or 32 % t.particles_in_cell == 0)


def test_oracle_clean_semantics_properties(oracle_lib):
    """Clean semantics on a self-contained scene: sorted grid, exact tables, symmetric candidate sets (every
    pair is seen from both sides exactly once - no stale-table double counting), and reference-compatible
    semantics only ever ADD candidates."""
    sc = small_cylinder_scene()
    st = pkg.make_initial_state(sc, seed=9, xz_half_width=45.0, y_range=(-25.0, -110.0))
    with make_oracle(oracle_lib, sc, capi.SEM_CLEAN) as a, make_oracle(oracle_lib, sc, capi.SEM_REFERENCE) as b:
        a.upload_state(st)
        b.upload_state(st)
        for step in range(4):
            a.run_stage(capi.STAGE_GRID_PARTICLES)
            b.run_stage(capi.STAGE_GRID_PARTICLES)
            keys, ids = a.grid(0)
            assert np.all(np.diff(keys) >= 0) and sorted(ids.tolist()) == list(range(a.n_particles))
            same = np.diff(keys) == 0
            assert np.all(np.diff(ids)[same] > 0), "ties must keep particle-id order (stable sort)"
            cells, starts, ends = a.cell_table(0)
            uniq, first, counts = np.unique(keys, return_index=True, return_counts=True)
            assert np.array_equal(cells, uniq) and np.array_equal(starts, first) and np.array_equal(ends, first + counts - 1)
            cnt_a, sum_a, _ = a.debug_candidates()
            cnt_b, _, _ = b.debug_candidates()
            assert int(cnt_a.sum()) % 2 == 0
            assert np.all(cnt_b >= cnt_a) or step == 0
            # candidate relation is symmetric: sum over particles of the checksum of their candidates equals the sum of
            # (own id hash) x (number of particles that see them)
            K = np.uint64(0x9E3779B97F4A7C15)
            with np.errstate(over="ignore"):
                lhs = sum_a.sum(dtype=np.uint64)
                rhs = ((np.arange(a.n_particles, dtype=np.uint64) + np.uint64(1)) * K * cnt_a.astype(np.uint64)).sum(dtype=np.uint64)
            assert lhs == rhs
            for sim in (a, b):
                for stage in range(2, 9):
                    sim.run_stage(stage)


def test_reference_shim_reproduces_the_scene_dump():
    """include/bcs_reference_shim.hpp compiled inside the reference's header tree (oracle/ref_harness/shim_check.cpp,
    built by oracle/build_ref.sh where /root/reference exists) yields the same user-level scene as the independent dump."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    found = 0
    for cfg in ("cfg1", "mini3"):
        a_path, b_path = os.path.join(ref, f"shim_scene_{cfg}.bcsd"), os.path.join(ref, f"scene_{cfg}.bcsd")
        if not (os.path.exists(a_path) and os.path.exists(b_path)):
            continue
        a, b = pkg.bcsd.read(a_path), pkg.bcsd.read(b_path)
        for k in a:
            assert np.array_equal(a[k], b[k]), (cfg, k)
        found += 1
    if not found:
        pytest.skip("oracle/_ref not built on this machine")


def test_headless_driver_is_built():
    exe = os.path.join(ROOT, "simulation-server_b200", "bcs_headless")
    assert os.path.exists(exe), "build() must produce the headless C++ driver"
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr

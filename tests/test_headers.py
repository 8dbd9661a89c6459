"""CPU: the header writers (simulation-server_b200/headers.py) emit the reference's compile-time config
format.  Where the reference tree is present (this container), the emitted headers are compiled inside the reference's
own meta-factory (oracle/ref_harness/ref_scene_dump.cpp, g++) and the scene it sees is compared bit for bit."""
import importlib
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pkg

headers = importlib.import_module("simulation-server_b200.headers")
workloads = importlib.import_module("simulation-server_b200.workloads")
REF = "/root/reference"


def _scene():
    presets = workloads.reference_presets()
    defs = [pkg.CellDef(3, presets[0][1].particles_in_cell, presets[0][1].springs, presets[0][1].spring_lengths, presets[0][1].vertices),
            pkg.CellDef(5, presets[1][1].particles_in_cell, presets[1][1].springs, presets[1][1].spring_lengths, presets[1][1].vertices)]
    vp, vi, ec, er = pkg.make_cylinder_vein(length=60.0, ring_vertices=24)
    return pkg.Scene(user_defs=defs, vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er), [n for n, _ in presets]


def _rbc_scene():
    """an RBC preset (26 particles, not in the reference's preset file) in a bifurcated vein"""
    presets = workloads.reference_presets()
    defs = [pkg.make_rbc_celldef(4), pkg.CellDef(2, presets[0][1].particles_in_cell, presets[0][1].springs, presets[0][1].spring_lengths, presets[0][1].vertices)]
    vp, vi, ec, er = pkg.make_bifurcated_vein(trunk_length=40.0, branch_length=40.0, ring_vertices=16)
    return pkg.Scene(user_defs=defs, vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er), ["Red_blood_cell_One", presets[0][0]]


def test_fixed_point_round_trip():
    for v in (3.849, -3.849, 7.697985, 4.757623, 50.0, -481.949, 34.946, 0.0, 1e-3):
        for p in (4, 7):
            try:
                i = headers._fixed(v, p)
            except ValueError:
                continue
            assert np.float32(i) / np.float32(10 ** (p - 1)) == np.float32(v)
    with pytest.raises(ValueError):
        headers._fixed(np.float32(1.0) / np.float32(3.0), 4)


def test_headers_are_written_in_the_reference_format(tmp_path):
    sc, names = _scene()
    headers.write_config(str(tmp_path), sc, names)
    vein = (tmp_path / "vein_definition.hpp").read_text()
    assert f"inline constexpr std::array<cvec, {len(sc.vein_pos)}> veinPositions {{" in vein
    assert f"inline constexpr std::array<unsigned int, {sc.vein_indices.size}> veinIndices {{" in vein
    assert "using VeinEndingCenters = mp_list<" in vein and "using VeinEndingRadii = mp_list<" in vein
    defs = (tmp_path / "blood_cells_definition.hpp").read_text()
    assert "using UserDefinedBloodCellList = mp_list<" in defs and "BloodCellDef<3, 20," in defs and "BloodCellDef<5, 20," in defs
    presets = (tmp_path / "blood_cell_presets.hpp").read_text()
    for n in names:
        assert f"using {n}_Springs = mp_list<" in presets and f"using {n}_Vertices = mp_list<" in presets


def test_rbc_preset_and_bifurcated_vein_run_on_the_oracle(oracle_lib):
    """an RBC preset and a Y-shaped vein as scene data - the host-core port accepts them and steps"""
    from conftest import capi, make_oracle
    sc, names = _rbc_scene()
    assert sc.vein_indices.max() < len(sc.vein_pos) and len(sc.ending_radii) == 2
    st = pkg.make_initial_state(sc, seed=3, xz_half_width=10.0, y_range=(-10.0, -25.0))
    with make_oracle(oracle_lib, sc) as orc:
        assert orc.n_particles == 4 * 26 + 2 * 20
        orc.upload_state(st)
        orc.step(5)
        pos = np.stack(orc.download(capi.PARTICLE_POS), 1)
        assert np.isfinite(pos).all() and np.abs(pos - np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1)).max() > 0.5


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present on this machine")
@pytest.mark.parametrize("which", ["presets", "rbc_bifurcated"])
def test_written_headers_compile_in_the_reference_meta_factory(tmp_path, which):
    sc, names = _scene() if which == "presets" else _rbc_scene()
    cfg = tmp_path / "cfg"
    headers.write_config(str(cfg), sc, names)
    src = tmp_path / "src"
    shutil.copytree(os.path.join(REF, "src"), src)
    for f in os.listdir(cfg):
        dst = src / "config" / f
        os.chmod(dst, 0o644)
        shutil.copy(cfg / f, dst)
    exe, out = tmp_path / "dump", tmp_path / "scene.bcsd"
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-w", f"-I{src}", f"-I{REF}/Libraries/include", f"-I{ROOT}/include",
                        os.path.join(ROOT, "oracle", "ref_harness", "ref_scene_dump.cpp"), "-o", str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    subprocess.run([str(exe), str(out)], check=True, timeout=120)
    got = pkg.Scene.load(str(out))
    assert np.array_equal(got.vein_pos, sc.vein_pos) and np.array_equal(got.vein_indices, sc.vein_indices)
    assert np.array_equal(got.ending_centers, sc.ending_centers) and np.array_equal(got.ending_radii, sc.ending_radii)
    assert len(got.user_defs) == len(sc.user_defs)
    for a, b in zip(got.user_defs, sc.user_defs):
        assert a.count == b.count and a.particles_in_cell == b.particles_in_cell
        assert np.array_equal(a.springs, b.springs) and np.array_equal(a.spring_lengths, b.spring_lengths)
        assert np.array_equal(a.vertices, b.vertices)

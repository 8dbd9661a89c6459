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


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present on this machine")
def test_written_headers_compile_in_the_reference_meta_factory(tmp_path):
    sc, names = _scene()
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

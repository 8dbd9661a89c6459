"""pytest configuration.

Markers: ``gpu`` = needs a CUDA device (the parity tests proper, run through the C ABI of libbcs.so).
Everything else runs on CPU: the oracle against the committed reference dumps, host logic, ABI surface.

Nothing here (or in any ``-m gpu`` test) reads /root/reference: scenes and reference dumps come from the
committed fixtures under tests/golden/ (made by tools/gpu_ref_goldens.sh + tools/make_goldens.py).
"""
import ctypes
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pkg = importlib.import_module("simulation-server_b200")
capi = importlib.import_module("simulation-server_b200.capi")

from cases import CASES  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ------------------------------------------------------------------------------------------ scenes / states
_scene_cache = {}


def golden_scene(cfg: str):
    """Scene rebuilt from the committed fixtures (bytes identical to what the reference harness dumped)."""
    if cfg not in _scene_cache:
        a = dict(np.load(os.path.join(GOLDEN, f"scene_{cfg}.npz")))
        a.update(dict(np.load(os.path.join(GOLDEN, "vein_default.npz"))))
        tmp = os.path.join("/tmp", f"bcs_test_scene_{cfg}_{os.getpid()}.bcsd")
        pkg.bcsd.write(tmp, a)
        _scene_cache[cfg] = pkg.Scene.load(tmp)
        os.remove(tmp)
    return _scene_cache[cfg]


def seeded_state(cfg: str, variant: str):
    st = pkg.make_initial_state(golden_scene(cfg), **CASES[(cfg, variant)])
    h = hashlib.sha256()
    for k in sorted(st):
        h.update(st[k].tobytes())
    return st, h.hexdigest()


def golden_dump(cfg, variant, step):
    return dict(np.load(os.path.join(GOLDEN, f"ref_{cfg}_{variant}_step{step:05d}.npz")))


def golden_file(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def state_checksums():
    out = {}
    with open(os.path.join(GOLDEN, "state_checksums.txt")) as f:
        for line in f:
            k, v = line.split()
            out[k] = v
    return out


def small_cylinder_scene(n_cells_a=60, n_cells_b=40, length=150.0):
    """A self-contained scene (no reference data): short straight vein + the two 20-particle presets of the
    cfg1 fixture + an 8-particle box type.  Used for clean-semantics / property tests."""
    base = golden_scene("mini3")
    vp, vi, ec, er = pkg.make_cylinder_vein(length=length)
    defs = []
    counts = {20: [n_cells_a, n_cells_b], 8: [30]}
    seen20 = 0
    for d in base.user_defs[:3]:
        if d.particles_in_cell == 20:
            c = counts[20][seen20]
            seen20 += 1
        else:
            c = counts[8][0]
        defs.append(pkg.CellDef(c, d.particles_in_cell, d.springs.copy(), d.spring_lengths.copy(), d.vertices.copy()))
    return pkg.Scene(user_defs=defs, vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er)


# ------------------------------------------------------------------------------------------ libraries
@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure).  Built on demand with `make -C oracle`."""
    path = os.path.join(ROOT, "oracle", "libbcs_oracle.so")
    src = os.path.join(ROOT, "oracle", "bcs_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return ctypes.CDLL(path)


@pytest.fixture(scope="session")
def bcs_lib():
    """The product library; GPU tests fail (not skip) if it is missing."""
    return capi.load_library()


def make_oracle(oracle_lib, scene, semantics=capi.SEM_CLEAN, seed=1234):
    return capi.Sim(scene, semantics=semantics, lib=oracle_lib, prefix="orc_", seed=seed)


def make_bcs(scene, semantics=capi.SEM_CLEAN, seed=1234, **kw):
    return capi.Sim(scene, semantics=semantics, seed=seed, **kw)

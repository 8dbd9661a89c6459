"""CPU: the host side of bench.py and the workload table (no GPU, no CUDA library calls)."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
workloads = importlib.import_module("simulation-server_b200.workloads")


@pytest.mark.parametrize("name,particles,expect", [("cfg1", None, 12_000), ("cfg2", None, 100_000), ("long_vein", 40_000, 40_000)])
def test_workloads_by_name(name, particles, expect):
    sc, st, info = workloads.by_name(name, particles)
    lay = sc.layout()
    assert info["particles"] == expect == int(lay.n_particles) == st["pos_x"].size
    assert info["workload"].startswith(name)
    # the seeded state lies inside the particle grid (clean semantics never see an out-of-grid particle at step 0)
    for ax, k in enumerate(("pos_x", "pos_y", "pos_z")):
        assert float(st[k].min()) > float(lay.grid_min[ax]) and float(st[k].max()) < float(lay.grid_max[ax])
    # same seed, same bytes
    sc2, st2, _ = workloads.by_name(name, particles)
    assert all(np.array_equal(st[k], st2[k]) for k in st)


def test_algorithmic_bytes_of_the_fused_pass_is_less_than_the_stages_it_replaces():
    bench = importlib.import_module("bench")
    N, B = 1_000_000, 50_000
    fused = bench.algorithmic_bytes("advance", N, B, 0, 0, 0)
    stages = sum(bench.algorithmic_bytes(k, N, B, 0, 0, 0) for k in ("finish_step", "springs", "cell_keys"))
    assert fused == 80 * N + 12 * B and stages == 140 * N + 12 * B


def test_reference_arm_prints_one_line_on_rank_0_only():
    """`bench.py --impl reference` under a multi-rank launch: rank 0 times the host-core port, the other ranks leave at once."""
    env = dict(os.environ, BCS_REF_THREADS="2")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "3", "--warmup", "1", "--gpus", "2"]
    out1 = subprocess.run(cmd, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
    out0 = subprocess.run(cmd, env=dict(env, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"), capture_output=True, text=True, timeout=300)
    assert out0.returncode == 0, out0.stderr[-2000:]
    line = json.loads(out0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["config"]["same_config"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2
    assert line["config"]["steps_timed"] == 3 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

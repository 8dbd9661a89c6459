#!/usr/bin/env python
"""Benchmark of the per-step physics hot path (BASELINE.json: particle-steps/s and ms/step at 1 M particles).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|long_vein|cfg5] [--particles P]
                    [--semantics clean|reference] [--impl reference]

One JSON line on stdout (rank 0).  A "step" = grid build + forces + integration of ALL particles of the
workload (bcs_step).  `value` is measured with the state resident in HBM (CUDA events on the library's
stream around K graph-replayed steps); `e2e` is the same metric through the C ABI with HOST buffers:
every step uploads positions/velocities/forces from pinned host memory and downloads the positions.
`--impl reference` times the host-core port of the reference step (oracle/, OpenMP, all host threads) -
upstream has no CPU path of its own - on the SAME workload (same_config).  The line carries a `parity` block: at N = 1
the first steps of the workload against the CPU oracle, at N > 1 the merged N-rank state against a one-GPU replay.
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_steps_per_s"
UNIT = "particle-steps/s"
CONTRACT_KERNELS = ("springs", "particle_collisions")


def _pkg():
    return importlib.import_module("simulation-server_b200"), importlib.import_module("simulation-server_b200.capi"), \
        importlib.import_module("simulation-server_b200.workloads")


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi needs a moment to start (longer on an 8-GPU box): wait for its first line, so that the samples
            # that follow fall into the warm-up / timed region
            t0 = time.time()
            while not self.rows and time.time() - t0 < 4.0:
                time.sleep(0.02)
            self.first = len(self.rows)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.03)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            if len(self.rows) > getattr(self, "first", 0):
                self.rows = self.rows[self.first:]   # drop the idle sample(s) taken before the region started

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- roofline
def algorithmic_bytes(kernel: str, N: int, B: int, V: int, T: int, c_occ: int, hits: int = 0, rows: int = 0):
    """Compulsory bytes of one launch (fp32 vec3 = 12 B, ids = 4 B); DESIGN.md section 'Kernels and rooflines'.
    The spring and collision figures are SURVEY.md 8(d)'s contract numbers."""
    table = {
        "cell_keys": 20 * N,                       # R pos 12N, W key+id 8N
        "slab_count_active": N,
        "radix_onesweep": 16 * N,                  # R key+id 8N, W key+id 8N (one launch = one 8-bit pass)
        "radix_tile_hist": 4 * N,                  # (classic 3-kernel passes, BCS_SORT=classic)
        "radix_scan": 0,
        "radix_scatter": 16 * N,
        "clear_cells": 4 * N + 8 * c_occ,
        "count_cell_starts": 4 * N,
        # compact cell index: R key+id; W start+key (+mask bit) per occupied cell; reorder pos+vel R+W.
        # row directory (rows > 0): R (key, id) 8N + pos 12N, W key+id 8N + sorted pos 12N
        "finalize_grid": 40 * N if rows else 8 * N + 12 * c_occ + 48 * N,
        "row_start_totals": 4 * rows, "row_start_scan": 8 * rows,      # R counts (twice), W starts
        "row_scatter": 20 * N,                     # R (key, place) 8N + row start 4N, W (key, id) 8N
        "pair_walk": 0,                            # fall-back of the pair search: returns at once
        "pair_force": 0, "pair_fold": 0,           # forces of the touching pairs (a few % of the particles): part of the collision stage
        # end of step k + springs / row count of step k+1 in one pass: R pos,vel,frc 36N, W pos,vel,frc 36N, W (key, place) 8N,
        # W centres 12B.  (The three stages it replaces: finish_step 72N + springs 48N + 12B + cell_keys 20N.)
        "advance": 80 * N + 12 * B,
        "vein_gather": 120 * V,
        "springs": 48 * N + 12 * B,                # R pos,vel,frc 36N, W frc 12N, W centres 12B
        "springs_count": 56 * N + 12 * B,          # head of a fused run: springs + the row count of the first grid build (W (key, place) 8N)
        "particle_collisions": 56 * N + 8 * c_occ,
        "tri_refit": 96 * T,                       # R 3 idx + 3 vertices (48), W packed triangle (48)
        "cell_box": 48 * T,                        # R packed triangles
        "vein_cull_cells": 12 * N,                 # R positions
        "vein_filter": 24 * N,                     # wall-grid path: R pos+vel of every particle
        "vein_collisions": 24 * N + 120 * hits,    # (wall-grid path: the triangle tests of the few candidates; same contract figure)
        "vein_ghost_splat": 0,
        "vein_masking": 0,                         # phase B of the wall-grid path: a few thousand particles
        "vein_apply": 0,                           # the stage's effect for the unmasked near hits (thread per listed particle)
        "wall_rebuild": 0,                         # returns at once unless a vertex left its margin
        "finish_step": 72 * N,                     # integrate (R 36N, W 24N) + vein-end test (12N)
        "integrate_particles": 60 * N,
        "vein_integrate": 72 * V,
        "vein_end": 12 * N,
        "advance_step": 0,
    }
    return table.get(kernel, 0)


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels captured with `ncu --set full`
    on this workload; profiles/traffic.json is written by tools/ncu_traffic.py from the committed captures."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------- oracle (CPU) leg
def oracle_library():
    path = os.path.join(ROOT, "oracle", "libbcs_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(path)
    lib.orc_threads.restype = ctypes.c_int
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to every rank; only rank 0 runs this leg)
    if "BCS_REF_THREADS" in os.environ:
        lib.orc_set_threads(ctypes.c_int(int(os.environ["BCS_REF_THREADS"])))
    elif hasattr(lib, "orc_set_threads"):
        lib.orc_set_threads(ctypes.c_int(len(os.sched_getaffinity(0))))
    return lib


def semantics_of(args, capi):
    return capi.SEM_REFERENCE if args.semantics == "reference" else capi.SEM_CLEAN


def time_oracle(args, steps: int, warmup: int):
    """Host-core port of the reference step (oracle/, OpenMP) on the WHOLE workload the product arm times."""
    pkg, capi, workloads = _pkg()
    lib = oracle_library()
    cores = int(lib.orc_threads())
    sc, st, info = workloads.by_name(args.workload, args.particles)
    sim = capi.Sim(sc, semantics=semantics_of(args, capi), lib=lib, prefix="orc_")
    sim.upload_state(st)
    # bounded: the CPU step costs ~0.27 s per million particles, so a long `--steps K` is cut after `budget_s` of timed work
    # (the per-step cost does not drift over a few hundred steps; the line reports how many steps were timed)
    budget_s = float(os.environ.get("BCS_REF_BUDGET_S", "150"))
    t0 = time.perf_counter()
    for w in range(warmup):
        sim.step(1)
        if w >= 2 and time.perf_counter() - t0 > 0.2 * budget_s:
            break
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        sim.step(1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    n = sim.n_particles
    sim.close()
    info = dict(info)
    info["steps_timed"] = done
    return n * done / dt, dt / done * 1e3, cores, n, info


def run_reference_arm(args):
    """`--impl reference`: upstream has no CPU path, so the reference arm is the host-core port of the same step
    (cpu_baseline.kind "port") on ALL host threads, same workload, same steps (same_config: true)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, n, info = time_oracle(args, args.steps, args.warmup)
    config = dict(info)
    config.update({"semantics": args.semantics, "same_config": True, "timed_particles": n})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"the whole {info['workload']} workload ({n} particles), {info['steps_timed']} steps after warm-up; "
                                   f"upstream has no CPU path, this is the host-core port under oracle/ (OpenMP, {cores} threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_cuda_baseline(workload: str):
    """The reference's OWN CUDA build timed on a B200 of this pool for the same scene (profiles/reference_cuda.json, written
    from tools/gpu_ref_bench.sh runs of the unmodified reference sources; oracle/build_ref.sh).  Reported, not a target."""
    path = os.path.join(ROOT, "profiles", "reference_cuda.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


# ----------------------------------------------------------------------------------------------- parity self-check
def parity_vs_oracle(args, sc, st, device):
    """N = 1: the first steps of the SAME workload on a second handle against the CPU oracle: integer outputs of step 0
    (sorted cell ids, sorted order, candidate counts + checksums, touching pairs) bit-exact, positions after `psteps` free
    steps within 1e-5 of the travelled distance scale."""
    pkg, capi, workloads = _pkg()
    lib = oracle_library()
    sem = semantics_of(args, capi)
    psteps = 2
    with capi.Sim(sc, semantics=sem, device=device) as sim, capi.Sim(sc, semantics=sem, lib=lib, prefix="orc_") as orc:
        sim.upload_state(st)
        orc.upload_state(st)
        sim.run_stage(capi.STAGE_GRID_PARTICLES)
        orc.run_stage(capi.STAGE_GRID_PARTICLES)
        (ka, ia), (kb, ib) = sim.grid(0), orc.grid(0)
        ca, cb = sim.debug_candidates(), orc.debug_candidates()
        sim.step(psteps)
        orc.step(psteps)
        a = np.stack(sim.download(capi.PARTICLE_POS), 1).astype(np.float64)
        b = np.stack(orc.download(capi.PARTICLE_POS), 1).astype(np.float64)
        f_a = np.stack(sim.download(capi.PARTICLE_FRC), 1).astype(np.float64)
        f_b = np.stack(orc.download(capi.PARTICLE_FRC), 1).astype(np.float64)
    scale = float(np.abs(b).max())
    fscale = float(np.percentile(np.abs(f_b).max(axis=1), 99))
    out = {
        "against": "CPU oracle (oracle/, host-core port pinned on reference-run fixtures)", "steps": psteps,
        "sorted_cell_ids_equal": bool(np.array_equal(ka, kb)), "sorted_order_equal": bool(np.array_equal(ia, ib)),
        "candidate_counts_equal": bool(np.array_equal(ca[0], cb[0])), "candidate_checksums_equal": bool(np.array_equal(ca[1], cb[1])),
        "touching_pairs_equal": bool(np.array_equal(ca[2], cb[2])),
        "pos_max_abs_diff": float(np.abs(a - b).max()), "pos_max_rel_diff": float(np.abs(a - b).max() / scale),
        "frc_max_abs_diff": float(np.abs(f_a - f_b).max()), "frc_p99_scale": fscale, "tolerance_rel": 1e-5,
    }
    out["ok"] = bool(out["sorted_cell_ids_equal"] and out["sorted_order_equal"] and out["candidate_counts_equal"] and
                     out["candidate_checksums_equal"] and out["touching_pairs_equal"] and out["pos_max_rel_diff"] <= 1e-5)
    return out


def parity_vs_single_gpu(sim, sc, st, total_steps, rank, world, local_rank, planes, dist, dd, capi):
    """N > 1: the merged state of the N-rank run (every blood cell from its owner, every vein vertex from the rank whose
    slab holds its rest position) against the SAME number of steps on one GPU (rank 0 replays): max |diff| must be 0."""
    lay = sc.layout()
    mine = {"pos": np.stack(sim.download(capi.PARTICLE_POS), 1), "vel": np.stack(sim.download(capi.PARTICLE_VEL), 1),
            "frc": np.stack(sim.download(capi.PARTICLE_FRC), 1), "vpos": np.stack(sim.download(capi.VEIN_POS), 1),
            "owned": sim.ownership()}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    out = None
    if rank == 0:
        with capi.Sim(sc, device=local_rank) as ref:
            ref.upload_state(st)
            ref.step(total_steps)
            single = {"pos": np.stack(ref.download(capi.PARTICLE_POS), 1), "vel": np.stack(ref.download(capi.PARTICLE_VEL), 1),
                      "frc": np.stack(ref.download(capi.PARTICLE_FRC), 1), "vpos": np.stack(ref.download(capi.VEIN_POS), 1)}
        owned = [g["owned"] for g in gathered]
        out = {"against": f"the same {total_steps} steps on one GPU (rank 0 replay)", "steps": total_steps}
        worst = 0.0
        for key in ("pos", "vel", "frc"):
            merged = dd.merge_owned([g[key] for g in gathered], owned, lay)
            d = float(np.abs(merged.astype(np.float64) - single[key]).max())
            out[f"{key}_max_abs_diff"] = d
            worst = max(worst, d)
        y0 = sc.vein_pos[:, 1]
        vm = np.empty_like(single["vpos"])
        for r in range(world):
            sel = (y0 >= planes[r + 1]) & (y0 < planes[r])
            vm[sel] = gathered[r]["vpos"][sel]
        out["vein_pos_max_abs_diff"] = float(np.abs(vm.astype(np.float64) - single["vpos"]).max())
        out["max_abs_diff"] = max(worst, out["vein_pos_max_abs_diff"])
        out["ok"] = out["max_abs_diff"] == 0.0
    return out


# ----------------------------------------------------------------------------------------------- product arm
def run_product(args):
    import torch
    import torch.distributed as dist

    pkg, capi, workloads = _pkg()
    dd = importlib.import_module("simulation-server_b200.distributed")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # slab decomposition along the vein axis.  Default: STRONG scaling, the same scene split over the ranks
    # (BASELINE.json's metric); --scaling weak keeps --particles PER RANK (a vein N times as long, same density)
    n_particles = args.particles * (world if args.scaling == "weak" else 1)
    sc, st, info = workloads.by_name(args.workload, n_particles)
    sem = semantics_of(args, capi)
    planes = None
    parity = None
    if world == 1 and not args.no_parity:
        parity = parity_vs_oracle(args, sc, st, local_rank)
    if world > 1:
        planes = dd.slab_boundaries(sc, st, world)
        sim = dd.create_slab_sim(sc, st, rank, world, local_rank, dd.broadcast_unique_id(rank), planes)
    else:
        sim = capi.Sim(sc, semantics=sem, device=local_rank, use_graph=True)
        sim.upload_state(st)
    N, B, V, T = sim.n_particles, sim.n_cells, sim.n_vertices, sim.n_triangles
    view = sim.device_view()
    stream = torch.cuda.ExternalStream(view.stream, device=local_rank)

    # ---- device-resident throughput: W warm-up steps, then exactly K steps between two events (max over ranks)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:   # samples from the warm-up on: clocks under load around the timed region
        sim.step(args.warmup)
        sim.synchronize()
        launches0 = sim.launch_count()
        barrier()
        torch.cuda.synchronize()
        start.record(stream)
        sim.step(args.steps)
        end.record(stream)
        sim.synchronize()
        torch.cuda.synchronize()
        barrier()
    ms = max_over_ranks(start.elapsed_time(end) / args.steps)
    launches = sim.launch_count() - launches0
    value = N / (ms * 1e-3)
    slab_info = sim.slab_counts() if world > 1 else None
    if world > 1 and not args.no_parity:
        parity = parity_vs_single_gpu(sim, sc, st, args.warmup + args.steps, rank, world, local_rank, planes, dist, dd, capi)

    # ---- per-kernel times (CUDA events around every launch, plain launches) and the roofline of the dominant kernel
    if world > 1:
        active = sim.slab_counts()["active_particles"]
        keys = sim.grid(0)[0][:active]
    else:
        keys, _ = sim.grid(0)
    c_occ = int(np.unique(keys).size)
    hits0 = sim.stats()["vein_hits"]
    prof_steps = 5
    prof = sim.profile_steps(prof_steps)
    hits = (sim.stats()["vein_hits"] - hits0) // prof_steps
    if world > 1:
        N_alg, B_alg = sim.slab_counts()["active_particles"], sim.slab_counts()["owned_cells"]   # this rank's share
    else:
        N_alg, B_alg = N, B
    total_ms = sum(v[0] for v in prof.values())
    kernels = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps, "share": v[0] / total_ms}
               for k, v in prof.items()}
    dominant = max(prof, key=lambda k: prof[k][0])
    peak, peak_src = measured_peaks()
    traffic = ncu_traffic()

    rows = int(sim.layout.grid_dims[1]) * int(sim.layout.grid_dims[2]) if "row_scatter" in prof else 0

    # In a fused run the spring stage lives inside `advance`; the contract figure for the spring kernel ALONE comes from its
    # staged entry point (bcs_run_stage), one launch per event pair, with a 256 MB fill between launches so that the 48 MB
    # it reads are not L2 hits left by the previous launch.
    prof_step = prof
    if world == 1 and "springs" not in prof:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
        reps, pairs = 12, []
        with torch.cuda.stream(stream):
            for r in range(reps):
                flush.fill_(r)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sim.run_stage(capi.STAGE_SPRINGS)
                e1.record(stream)
                pairs.append((e0, e1))
        sim.synchronize()
        torch.cuda.synchronize()
        t = [a.elapsed_time(b) for a, b in pairs[2:]]
        prof = dict(prof)
        prof["springs"] = (float(sum(t)), len(t))
        del flush

    def roof(name):
        per_launch_ms = prof[name][0] / prof[name][1]
        b = algorithmic_bytes(name, N_alg, B_alg, V, T, c_occ, hits, rows)
        gbs = b / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        t = traffic.get(name) if world == 1 and args.workload == "long_vein" and args.particles == 1_000_000 else None
        return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": t["bytes"] if t else None, "traffic_source": t["source"] if t else None,
                "algorithmic_bytes_per_launch": b, "ms_per_launch": per_launch_ms, "peak_source": peak_src}

    roofline = roof(dominant)
    roofline["contract_kernels"] = {k: roof(k) for k in CONTRACT_KERNELS if k in prof}
    if "pair_force" in prof:
        # the collision STAGE is pair search + (idle) fall-back + pair force + fold: SURVEY 8(d)'s stage figure over their summed time
        c = roofline["contract_kernels"]["particle_collisions"]
        stage_ms = sum(prof[k][0] / prof[k][1] for k in ("particle_collisions", "pair_walk", "pair_force", "pair_fold") if k in prof)
        c.update({"kernel": "particle_collisions + pair_walk + pair_force + pair_fold (the stage)", "search_ms_per_launch": c["ms_per_launch"], "ms_per_launch": stage_ms,
                  "achieved": c["algorithmic_bytes_per_launch"] / (stage_ms * 1e-3) / 1e9})
        c["frac"] = c["achieved"] / peak
    if "advance" in prof:
        roofline["contract_kernels"]["advance"] = roof("advance")
        roofline["contract_kernels"]["advance"]["stages_replaced_bytes"] = 140 * N_alg + 12 * B_alg
    step_bytes = sum(algorithmic_bytes(k, N_alg, B_alg, V, T, c_occ, hits, rows) * v[1] / prof_steps for k, v in prof_step.items())
    roofline["whole_step"] = {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms * 1e-3) / 1e9, "frac": step_bytes / (ms * 1e-3) / 1e9 / peak}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    host = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in st.items()}
    out = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(3)]
    hp = {k: ctypes.cast(v.data_ptr(), ctypes.POINTER(ctypes.c_float)) for k, v in host.items()}
    op = [ctypes.cast(v.data_ptr(), ctypes.POINTER(ctypes.c_float)) for v in out]
    n32 = ctypes.c_int32(N)

    # N > 1: every rank moves only the particles of the blood cells it owns (bcs_upload_owned / bcs_download_owned) between
    # the device and full-length host arrays in the reference's layout.  A blood cell that changes owner during the step
    # reaches its new owner's host arrays through that rank's download, so the loop carries the whole particle state
    # (positions, velocities, forces) both ways; at N = 1 the step's result read back is the positions.
    arrays = {capi.PARTICLE_POS: (hp["pos_x"], hp["pos_y"], hp["pos_z"]), capi.PARTICLE_VEL: (hp["vel_x"], hp["vel_y"], hp["vel_z"]),
              capi.PARTICLE_FRC: (hp["frc_x"], hp["frc_y"], hp["frc_z"])}

    def e2e_step():
        if world > 1:
            for w, (ax, ay, az) in arrays.items():
                sim._call("upload_owned", sim._h, w, ax, ay, az, n32)
            sim._call("step", sim._h, ctypes.c_int32(1))
            for w, (ax, ay, az) in arrays.items():
                sim._call("download_owned", sim._h, w, ax, ay, az, n32)
            return
        for w, (ax, ay, az) in arrays.items():
            sim._call("upload", sim._h, w, ax, ay, az, n32)
        sim._call("step", sim._h, ctypes.c_int32(1))
        sim._call("download", sim._h, capi.PARTICLE_POS, op[0], op[1], op[2], n32)

    if world > 1:
        # the host arrays start as the CURRENT state of the blood cells this rank owns (an owned-only upload does not move
        # ownership: what is uploaded has to be the state of the cells the rank holds)
        for w, (ax, ay, az) in arrays.items():
            sim._call("download_owned", sim._h, w, ax, ay, az, n32)

    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(3):
        e2e_step()
    sim.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sim.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    barrier()
    if world > 1:
        # bytes over all ranks: 3 arrays x 12 B + the 4 B slot index per owned particle up; 3 arrays x 12 B per owned particle +
        # the ownership flags (1 B per blood cell and rank) down
        cnts = sim.slab_counts()
        own = torch.tensor([float(cnts["active_particles"] - cnts["ghost_particles"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(own)
        owned_particles = int(own.item())
        h2d, d2h = 40 * owned_particles, 36 * owned_particles + B * world
    else:
        h2d, d2h = 36 * N, 12 * N
    e2e = {"value": N / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "transfers": "owned blood cells only, per rank (bcs_upload_owned / bcs_download_owned)" if world > 1 else "whole arrays"}

    # ---- CPU baseline (rank 0, N=1): the host-core port on the whole workload, a few steps
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cv, cms, cores, cn, _ = time_oracle(args, 3, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": cms,
               "sample": f"the whole workload ({cn} particles), 3 steps after 1 warm-up, OpenMP on {cores} threads"}

    ws_mb = (16 * 3 * N + 16 * 2 * N + 16 * N + 64 * c_occ + 48 * V + 48 * T) / 1e6
    config = dict(info)
    if world > 1:
        config["parallelism"] = f"y-slab decomposition over {world} ranks, NCCL halo exchange + blood-cell migration (one grouped send/recv per neighbour and step)"
        config["rank0_slab"] = slab_info
    config.update({"semantics": args.semantics, "launch": "CUDA graph replay" if world == 1 else "CUDA graph replay, NCCL send/recv captured in the graph",
                   "timed_window": f"steps {args.warmup}..{args.warmup + args.steps} after the seeded initial state (the workload is not stationary: "
                                   "blood cells drift towards the wall)",
                   "l2": f"no flush: per-step working set ~{ws_mb:.0f} MB exceeds the 126 MB L2 (inputs larger than L2)" if ws_mb > 126 else
                         f"no flush: per-step working set ~{ws_mb:.0f} MB is L2 resident (this scene is smaller than the 126 MB L2)",
                   "occupied_grid_cells": c_occ, "grid_cells": int(sim.layout.grid_cells)})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config, "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": reference_cuda_baseline(info["workload"]), "parity": parity,
        "kernels": kernels,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)   # before the teardown: a line is never lost to a close that does not return
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="long_vein", choices=["cfg1", "cfg2", "cfg3", "long_vein", "cfg5"],
                    help="BASELINE.json configs: cfg1 default scene, cfg2 100 k default vein, cfg3 1 M default vein (dense), "
                         "long_vein 1 M at default density (default; the metric's config), cfg5 10 M high hematocrit")
    ap.add_argument("--particles", type=int, default=None, help="long_vein / cfg5 only (default 1 M / 10 M)")
    ap.add_argument("--semantics", default="clean", choices=["clean", "reference"],
                    help="clean (default) or the reference's quirks bit for bit (stale cell tables, slice radii): single GPU only")
    ap.add_argument("--impl", default="bcs", choices=["bcs", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the --particles scene split over the ranks (default, BASELINE metric); weak = --particles per rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.particles is None:
        args.particles = 10_000_000 if args.workload == "cfg5" else 1_000_000
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product(args)


if __name__ == "__main__":
    try:
        main()
    except BaseException as e:
        if isinstance(e, SystemExit) and e.code in (0, None):
            raise
        # a rank that raises must take the job down: its peers would otherwise wait in NCCL for ever
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)

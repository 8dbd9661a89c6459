#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE: the UNMODIFIED reference CUDA sources built for the bench workload.

Writes the long-vein scene of `workloads.long_vein(N)` in the reference's own config-header format
(simulation-server_b200/headers.py), builds the headless reference binary from the reference tree with those headers
overlaid (oracle/build_ref.sh: scratch copy under /tmp, outputs only in oracle/_ref/) and writes the seeded state file
next to it.  `tools/gpu_ref_bench.sh` then times it on the GPU box: the "reference's own CUDA build on the same box"
baseline of BASELINE.json for the SAME scene bench.py times.

usage: python tools/build_ref_long_vein.py [particles=1000000]
"""
import importlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simulation-server_b200")
headers = importlib.import_module("simulation-server_b200.headers")
workloads = importlib.import_module("simulation-server_b200.workloads")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    name = f"longvein{n}"
    sc, st, info = workloads.long_vein(n)
    overlay = f"/tmp/bcs_overlay_{name}"
    headers.write_config(overlay, sc, [nm for nm, _ in workloads.reference_presets()])
    out = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out, exist_ok=True)
    pkg.bcsd.write(os.path.join(out, f"state_{name}_seed.bcsd"), st)
    t0 = time.time()
    env = dict(os.environ)
    env.setdefault("REF_CONSTEXPR_LIMIT", "400000000")   # vein_factory.hpp loops over every vertex / triangle in constexpr context
    r = subprocess.run([os.path.join(ROOT, "oracle", "build_ref.sh"), name, overlay], capture_output=True, text=True, env=env)
    print(r.stdout[-2000:], r.stderr[-4000:])
    print(f"build_ref {name}: exit {r.returncode} after {time.time() - t0:.0f} s")
    sys.exit(r.returncode)


if __name__ == "__main__":
    main()

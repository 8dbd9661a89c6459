"""ON THE GPU BOX: replays tests/test_gpu_parity.py::_compare_step for one case and prints which of the candidate
arrays (count, checksum, hits) differ from the oracle, and how close the offending pairs are to the touch threshold."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import refcheck
import ctypes
from conftest import ROOT, capi, golden_scene, make_bcs, make_oracle, seeded_state  # noqa

cfg, variant, sem = sys.argv[1], sys.argv[2], int(sys.argv[3])
sc = golden_scene(cfg)
st, _ = seeded_state(cfg, variant)
orc_lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libbcs_oracle.so"))
with make_bcs(sc, sem) as sim, make_oracle(orc_lib, sc, sem) as orc:
    orc.upload_state(st)
    for step in range(6):
        for which in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.VEIN_FRC):
            refcheck.up(sim, which, refcheck.down(orc, which))
        sim.run_stage(capi.STAGE_GRID_PARTICLES); orc.run_stage(capi.STAGE_GRID_PARTICLES)
        ca, cb = sim.debug_candidates(), orc.debug_candidates()
        for name, x, y in zip(("count", "sum", "hits"), ca, cb):
            bad = np.nonzero(x != y)[0]
            print(f"step {step} {name}: {len(bad)} differ", bad[:10], x[bad[:10]], y[bad[:10]])
        bad = np.nonzero(ca[2] != cb[2])[0]
        if len(bad):
            pos = refcheck.down(orc, capi.PARTICLE_POS).astype(np.float32)
            R = sim.table(capi.TABLE_COLL_RADII) if hasattr(capi, "TABLE_COLL_RADII") else None
            for p in bad[:6]:
                d = pos - pos[p]
                d2 = (d.astype(np.float64) ** 2).sum(1)
                near = np.argsort(d2)[1:6]
                print("  pid", p, "nearest", near, "dist", np.sqrt(d2[near]))
        for s in (capi.STAGE_VEIN_GATHER, capi.STAGE_SPRINGS, capi.STAGE_PARTICLE_COLLISIONS, capi.STAGE_VEIN_COLLISIONS,
                  capi.STAGE_INTEGRATE_PARTICLES, capi.STAGE_INTEGRATE_VEIN, capi.STAGE_VEIN_END):
            sim.run_stage(s); orc.run_stage(s)

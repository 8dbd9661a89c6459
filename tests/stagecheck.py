"""libbcs vs the CPU oracle, stage by stage from identical inputs (shared by test_gpu_parity.py and test_gpu_scale.py).

Integer outputs (cell ids, sorted order, cell tables, candidate counts / checksums / hit counts, first-hit triangles)
must be equal.  The one place where two correct float32 builds may legitimately pick different integers is a
Moeller-Trumbore decision that sits ON a threshold (vein_collisions.cu:11-45: |a| < EPS, u in [0,1], v >= 0,
u + v <= 1, t > EPS): the oracle is compiled without FMA contraction, the kernel with it.  Every disagreeing particle is
therefore re-evaluated in float64 and must be shown to be such a borderline case; anything else fails.
"""
import numpy as np

import refcheck
from conftest import capi

EPS = 1e-6
SPLAT_RTOL = 1e-4


def moller_trumbore_margin(origin, direction, v0, v1, v2):
    """float64 evaluation of realCollisionDetection; returns (accepted, margin) where margin is the distance of the
    closest decision quantity to its threshold (0 = exactly on it)."""
    e1, e2 = v1 - v0, v2 - v0
    h = np.cross(direction, e2)
    a = float(np.dot(e1, h))
    scale = float(np.linalg.norm(e1) * np.linalg.norm(e2)) + 1e-30
    if abs(a) < EPS:
        return False, abs(abs(a) - EPS) / scale
    f = 1.0 / a
    s = origin - v0
    u = f * float(np.dot(s, h))
    q = np.cross(s, e1)
    v = f * float(np.dot(direction, q))
    t = f * float(np.dot(e2, q))
    # rounding of u, v grows like 1/|a| (a is a cancelled determinant): margins are taken relative to that
    amp = max(1.0, scale / abs(a))
    margins = [abs(u) / amp, abs(1.0 - u) / amp, abs(v) / amp, abs(1.0 - u - v) / amp, abs(t - EPS) / (amp * max(1.0, abs(t))),
               abs(abs(a) - EPS) / scale]
    ok = (0.0 <= u <= 1.0) and v >= 0.0 and u + v <= 1.0 and t > EPS
    return ok, min(margins)


def explain_first_hit_disagreements(tri_a, t_a, tri_b, t_b, pos, vel, vpos, vein_indices, tol=2e-5, reach=None):
    """Particles whose first-hit triangle differs between the two sides.  Returns (explained, unexplained) index lists:
    explained = at least one of the two triangles is a borderline Moeller-Trumbore decision in float64."""
    diff = np.nonzero(tri_a != tri_b)[0]
    explained, unexplained = [], []
    for p in diff:
        o = pos[p].astype(np.float64)
        d = vel[p].astype(np.float64)
        n = np.linalg.norm(d)
        d = d / n if n > 0 else d
        best = np.inf
        for tri in (int(tri_a[p]), int(tri_b[p])):
            if tri < 0:
                continue
            i0, i1, i2 = (int(k) for k in vein_indices[tri])
            _, m = moller_trumbore_margin(o, d, vpos[i0].astype(np.float64), vpos[i1].astype(np.float64), vpos[i2].astype(np.float64))
            best = min(best, m)
        (explained if best <= tol else unexplained).append((int(p), float(best)))
    return explained, unexplained


def sync_inputs(sim, orc):
    for which in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC, capi.VEIN_POS, capi.VEIN_VEL, capi.VEIN_FRC):
        refcheck.up(sim, which, refcheck.down(orc, which))


def compare_step(sim, orc, scene, nsteps, tag, check_tables=True):
    """one step at a time: stage outputs of libbcs vs oracle from identical inputs.  Returns a summary dict."""
    names = ("count", "checksum", "hits")
    summary = {"borderline_first_hits": 0, "pair_hits": 0, "vein_hits": 0}
    for step in range(nsteps):
        sync_inputs(sim, orc)
        sim.run_stage(capi.STAGE_GRID_PARTICLES); orc.run_stage(capi.STAGE_GRID_PARTICLES)
        for which in (0, 1):
            ka, ia = sim.grid(which); kb, ib = orc.grid(which)
            assert np.array_equal(ka, kb) and np.array_equal(ia, ib), f"{tag} step {step}: grid {which}"
            if check_tables:
                ta, tb = sim.cell_table(which), orc.cell_table(which)
                assert all(np.array_equal(x, y) for x, y in zip(ta, tb)), f"{tag} step {step}: cell table {which}"
        ca, cb = sim.debug_candidates(), orc.debug_candidates()
        for name, x, y in zip(names, ca, cb):
            bad = np.nonzero(x != y)[0]
            assert len(bad) == 0, (f"{tag} step {step}: candidate {name} differs for {len(bad)} particles, first {bad[:6]}: "
                                   f"libbcs {x[bad[:6]]} oracle {y[bad[:6]]}")
        summary["pair_hits"] += int(ca[2].sum())
        for st in (capi.STAGE_VEIN_GATHER, capi.STAGE_SPRINGS, capi.STAGE_PARTICLE_COLLISIONS):
            sim.run_stage(st); orc.run_stage(st)
            refcheck.assert_close(refcheck.down(sim, capi.PARTICLE_FRC), refcheck.down(orc, capi.PARTICLE_FRC), f"{tag} step {step} stage {st} forces")
        refcheck.assert_close(refcheck.down(sim, capi.VEIN_FRC), refcheck.down(orc, capi.VEIN_FRC), f"{tag} step {step} vein spring forces",
                              scale=max(0.5, float(np.abs(refcheck.down(orc, capi.VEIN_FRC)).max())))
        # first-hit triangles (Q8): equal, or a float64-proven borderline Moeller-Trumbore decision
        ha, hb = sim.debug_vein_hits(), orc.debug_vein_hits()
        pos, vel, vpos = refcheck.down(orc, capi.PARTICLE_POS), refcheck.down(orc, capi.PARTICLE_VEL), refcheck.down(orc, capi.VEIN_POS)
        explained, unexplained = explain_first_hit_disagreements(ha[0], ha[1], hb[0], hb[1], pos, vel, vpos, scene.vein_indices)
        assert not unexplained, (f"{tag} step {step}: first-hit triangle differs for {len(unexplained)} particles that are NOT on a "
                                 f"Moeller-Trumbore threshold: {unexplained[:6]} (libbcs {ha[0][[p for p, _ in unexplained[:6]]]}, "
                                 f"oracle {hb[0][[p for p, _ in unexplained[:6]]]})")
        assert len(explained) <= max(3, len(pos) // 20000), f"{tag} step {step}: {len(explained)} borderline first hits - too many to be rounding"
        summary["borderline_first_hits"] += len(explained)
        summary["vein_hits"] += int(((hb[0] >= 0) & (hb[1] <= 6.0)).sum())
        ok = np.ones(len(pos), bool)
        ok[[p for p, _ in explained]] = False
        # vertices that receive a splat from a borderline particle are excluded from the splat comparison
        vok = np.ones(len(vpos), bool)
        for p, _ in explained:
            for tri in (int(ha[0][p]), int(hb[0][p])):
                if tri >= 0:
                    vok[scene.vein_indices[tri].astype(np.int64)] = False
        sim.run_stage(capi.STAGE_VEIN_COLLISIONS); orc.run_stage(capi.STAGE_VEIN_COLLISIONS)
        refcheck.assert_close(refcheck.down(sim, capi.PARTICLE_FRC)[ok], refcheck.down(orc, capi.PARTICLE_FRC)[ok], f"{tag} step {step} forces after vein collisions")
        refcheck.assert_close(refcheck.down(sim, capi.PARTICLE_VEL)[ok], refcheck.down(orc, capi.PARTICLE_VEL)[ok], f"{tag} step {step} velocities after vein collisions")
        vf = refcheck.down(orc, capi.VEIN_FRC)
        # barycentric weights divide by d00*d11 - d01^2 (a cancelled determinant): the contraction of that expression
        # moves a weight by up to ~1e-5 relative, so splats are compared at SPLAT_RTOL of the largest vertex force
        refcheck.assert_close(refcheck.down(sim, capi.VEIN_FRC)[vok], vf[vok], f"{tag} step {step} vein forces", rtol=SPLAT_RTOL,
                              scale=max(0.5, float(np.abs(vf).max())))
        for st in (capi.STAGE_INTEGRATE_PARTICLES, capi.STAGE_INTEGRATE_VEIN, capi.STAGE_VEIN_END):
            sim.run_stage(st); orc.run_stage(st)
        refcheck.assert_close(refcheck.down(sim, capi.PARTICLE_POS)[ok], refcheck.down(orc, capi.PARTICLE_POS)[ok], f"{tag} step {step} positions", scale=0.0)
        refcheck.assert_close(refcheck.down(sim, capi.VEIN_POS)[vok], refcheck.down(orc, capi.VEIN_POS)[vok], f"{tag} step {step} vein positions", scale=0.0)
    return summary

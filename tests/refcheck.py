"""Stage-by-stage comparison of a simulation handle (CPU oracle or libbcs) with dumps of the reference.

Every stage is replayed from the REFERENCE's dump of the previous stage, so that the racy parts of the
reference cannot leak from one check into the next.

Float tolerance ("1e-5 relative", BASELINE.json north_star): a vector passes if
    |a - b|_inf <= RTOL * max(|a|_inf, |b|_inf)  +  RTOL * S
with RTOL = 1e-5 and S the magnitude of the terms the stage sums (taken as the 99th percentile of |b|_inf
over the array): forces are differences of large spring/damping terms, so a particle whose net force is
near zero is compared relative to the size of its summands, not of the cancelled result.
"""
import itertools

import numpy as np

from conftest import capi

RTOL = 1e-5


def vec(d, name):
    return np.stack([d[name + "_x"], d[name + "_y"], d[name + "_z"]], axis=1)


def up(sim, which, a):
    sim.upload(which, a[:, 0], a[:, 1], a[:, 2])


def down(sim, which):
    return np.stack(sim.download(which), axis=1)


def mismatch(a, b, rtol=RTOL, scale=None):
    """indices of rows outside the tolerance"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b).max(axis=1)
    mag = np.maximum(np.abs(a).max(axis=1), np.abs(b).max(axis=1))
    if scale is None:
        scale = np.percentile(np.abs(b).max(axis=1), 99) if len(b) else 0.0
    return np.where(d > rtol * mag + rtol * scale)[0], d, scale


def assert_close(a, b, what, rtol=RTOL, scale=None, allowed=0):
    bad, d, scale = mismatch(a, b, rtol, scale)
    assert len(bad) <= allowed, (f"{what}: {len(bad)} of {len(d)} rows outside tolerance (rtol={rtol}, scale={scale:.3g}); "
                                 f"worst |d|={d.max():.3e} at row {int(d.argmax())}")


# ---------------------------------------------------------------------------------------------------
def check_grid(sim, d, which=0, tables=True):
    """keys / sorted ids (and tables) must be bit-identical to the reference."""
    name = "tgrid" if which else "pgrid"
    keys, ids = sim.grid(which)
    assert np.array_equal(keys, d[name + ".keys"]), f"{name}: sorted cell ids differ"
    assert np.array_equal(ids, d[name + ".ids"]), f"{name}: sorted object order differs"
    if tables:
        c, s, e = sim.cell_table(which)
        assert np.array_equal(c, d[name + ".table_cells"]), f"{name}: set of written cells differs"
        assert np.array_equal(s, d[name + ".table_starts"]), f"{name}: cell starts differ"
        assert np.array_equal(e, d[name + ".table_ends"]), f"{name}: cell ends differ"


def expected_candidates(d, layout):
    """Candidate sets of the particle-collision stage computed in numpy from the REFERENCE's own grid dump
    (keys, ids, persistent tables), following particle_collisions.cuh:104-269.  Returns per-particle
    (count, checksum) in the encoding of bcs_debug_candidates."""
    keys, ids = d["pgrid.keys"], d["pgrid.ids"]
    nx, ny, nz, cells = (int(v) for v in d["pgrid.dims"])
    starts = np.zeros(cells, np.int64)
    ends = np.zeros(cells, np.int64)
    starts[d["pgrid.table_cells"]] = d["pgrid.table_starts"]
    ends[d["pgrid.table_cells"]] = d["pgrid.table_ends"]
    pos = vec(d, "begin.pos")
    gmin = np.asarray(list(layout.grid_min), np.float32)
    n = len(keys)
    cnt = np.zeros(n, np.int32)
    chk = np.zeros(n, np.uint64)
    K = np.uint64(0x9E3779B97F4A7C15)

    def rng(i, c):
        if i < 1:
            return (0, 1)
        if i > c - 2:
            return (-1, 0)
        return (-1, 1)

    with np.errstate(over="ignore"):
        for slot in range(n):
            pid = int(ids[slot])
            p = pos[pid]
            xi = int(np.float32(p[0] - gmin[0]) / np.float32(2))
            yi = int(np.float32(p[1] - gmin[1]) / np.float32(2))
            zi = int(np.float32(p[2] - gmin[2]) / np.float32(2))
            (x0, x1), (y0, y1), (z0, z1) = rng(xi, nx), rng(yi, ny), rng(zi, nz)
            c0 = int(keys[slot])
            tot, acc = 0, np.uint64(0)
            for x in range(x0, x1 + 1):
                for y in range(y0, y1 + 1):
                    for z in range(z0, z1 + 1):
                        c = c0 + z * nx * ny + y * nx + x
                        if c < 0 or c >= cells:
                            continue
                        s, e = int(starts[c]), int(ends[c])
                        if e < s:
                            continue
                        q = ids[s:e + 1]
                        q = q[q != pid]
                        tot += len(q)
                        if len(q):
                            acc = acc + ((q.astype(np.uint64) + np.uint64(1)) * K).sum(dtype=np.uint64)
            cnt[pid] = tot
            chk[pid] = acc
    return cnt, chk


def check_springs(sim, d, dt, d_fact, layout):
    """Springs from the reference's begin state.  The reference kernel reads mates' forces while other
    threads overwrite theirs (blood_cells.cu:99 vs :118): rows outside the tolerance must be
    explainable by that race, i.e. |diff| <= 0.5 * d_fact * dt * sum_mates |F_new - F_old|."""
    begin_f = vec(d, "begin.frc")
    up(sim, capi.PARTICLE_POS, vec(d, "begin.pos"))
    up(sim, capi.PARTICLE_VEL, vec(d, "begin.vel"))
    up(sim, capi.PARTICLE_FRC, begin_f)
    sim.run_stage(capi.STAGE_SPRINGS)
    got, ref = down(sim, capi.PARTICLE_FRC), vec(d, "springs.frc")
    assert_close(down(sim, capi.CELL_CENTERS), vec(d, "springs.centers"), "blood cell centres")
    bad, diff, scale = mismatch(got, ref)
    if len(bad):
        change = np.abs(ref - begin_f).sum(axis=1)          # |F_new - F_old| per particle (L1)
        env = np.zeros(len(ref))
        for t in layout.types[:layout.n_types]:
            P, p0, cnt = t.particles_in_cell, t.particle_start, t.count
            per_cell = change[p0:p0 + cnt * P].reshape(cnt, P).sum(axis=1)
            env[p0:p0 + cnt * P] = np.repeat(per_cell, P)
        bound = 0.5 * d_fact * dt * env * 1.01 + RTOL * scale
        unexplained = [int(i) for i in bad if diff[i] > bound[i]]
        assert not unexplained, f"spring forces: rows {unexplained[:8]} differ beyond the reference's own race envelope"
    frac = len(bad) / len(ref)
    assert frac < 0.05, f"spring forces: {frac:.1%} of the particles hit the reference race, expected a small minority"
    return len(bad)


def check_particle_collisions(sim, d):
    """grid must already be built from begin.pos (with the table history of the run)"""
    up(sim, capi.PARTICLE_FRC, vec(d, "springs.frc"))
    sim.run_stage(capi.STAGE_PARTICLE_COLLISIONS)
    assert_close(down(sim, capi.PARTICLE_FRC), vec(d, "pcoll.frc"), "forces after particle collisions")


def check_vein_collisions(sim, d, scene, have_vein):
    up(sim, capi.PARTICLE_POS, vec(d, "begin.pos"))
    up(sim, capi.PARTICLE_VEL, vec(d, "begin.vel"))
    up(sim, capi.PARTICLE_FRC, vec(d, "pcoll.frc"))
    if have_vein:
        up(sim, capi.VEIN_POS, vec(d, "begin.vein_pos"))
        up(sim, capi.VEIN_VEL, vec(d, "begin.vein_vel"))
        up(sim, capi.VEIN_FRC, vec(d, "vein_gather.vein_frc"))
    tri, t = sim.debug_vein_hits()
    sim.run_stage(capi.STAGE_VEIN_COLLISIONS)
    assert_close(down(sim, capi.PARTICLE_FRC), vec(d, "vcoll.frc"), "forces after vein collisions")
    assert_close(down(sim, capi.PARTICLE_VEL), vec(d, "vcoll.vel"), "velocities after vein collisions")
    n_hits = int(((tri >= 0) & (t <= 6.0)).sum())
    if have_vein:
        # vein force splats: the reference's plain `+=` loses updates when several particles hit triangles that
        # share a vertex in the same step (vein_collisions.cu:272-274).  Vertices with a single
        # contribution must match; for shared vertices the reference must equal a partial sum of ours.
        before, ref = vec(d, "vein_gather.vein_frc"), vec(d, "vcoll.vein_frc")
        got = down(sim, capi.VEIN_FRC)
        vel = vec(d, "begin.vel")
        hits = np.where((tri >= 0) & (t <= 6.0))[0]
        contrib = {}
        for p in hits:
            for k in range(3):
                contrib.setdefault(int(scene.vein_indices[tri[p], k]), []).append(int(p))
        bad, _, scale = mismatch(got, ref, scale=np.abs(ref).max())
        for v in bad:
            ps = contrib.get(int(v), [])
            assert len(ps) > 1, f"vein vertex {v}: force differs although at most one particle contributed"
            assert len(ps) <= 12
            # our total minus the reference = sum of the contributions the reference lost; every contribution is
            # b_k * 0.005 * v_p, so the lost part must be a non-negative combination of the contributors' velocities
            lost = got[v] - ref[v]
            basis = np.stack([0.005 * vel[p] for p in ps], axis=1)          # 3 x k
            coef, *_ = np.linalg.lstsq(basis, lost, rcond=None)
            resid = np.abs(basis @ coef - lost).max()
            assert resid <= 1e-4 * max(1.0, np.abs(lost).max()), f"vein vertex {v}: difference is not a lost update"
    return n_hits


def check_integration(sim, d, have_vein):
    up(sim, capi.PARTICLE_POS, vec(d, "begin.pos"))
    up(sim, capi.PARTICLE_VEL, vec(d, "vcoll.vel"))
    up(sim, capi.PARTICLE_FRC, vec(d, "vcoll.frc"))
    sim.run_stage(capi.STAGE_INTEGRATE_PARTICLES)
    assert_close(down(sim, capi.PARTICLE_POS), vec(d, "integrate.pos"), "positions after integration", scale=0.0)
    assert_close(down(sim, capi.PARTICLE_VEL), vec(d, "integrate.vel"), "velocities after integration", scale=0.0)
    if have_vein:
        up(sim, capi.VEIN_POS, vec(d, "begin.vein_pos"))
        up(sim, capi.VEIN_VEL, vec(d, "begin.vein_vel"))
        up(sim, capi.VEIN_FRC, vec(d, "vcoll.vein_frc"))
        sim.run_stage(capi.STAGE_INTEGRATE_VEIN)
        assert_close(down(sim, capi.VEIN_POS), vec(d, "integrate.vein_pos"), "vein positions after integration", scale=0.0)
        assert_close(down(sim, capi.VEIN_VEL), vec(d, "integrate.vein_vel"), "vein velocities after integration",
                     scale=np.abs(vec(d, "integrate.vein_vel")).max())
        assert np.abs(down(sim, capi.VEIN_FRC)).max() == 0.0, "vein forces must be cleared by the vein integrator"
    sim.run_stage(capi.STAGE_VEIN_END)
    assert_close(down(sim, capi.PARTICLE_POS), vec(d, "end.pos"), "positions after vein-end handling", scale=0.0)


def check_vein_gather(sim, d):
    up(sim, capi.VEIN_POS, vec(d, "begin.vein_pos"))
    up(sim, capi.VEIN_VEL, vec(d, "begin.vein_vel"))
    z = np.zeros_like(vec(d, "begin.vein_pos"))
    up(sim, capi.VEIN_FRC, z)
    sim.run_stage(capi.STAGE_VEIN_GATHER)
    ref = vec(d, "vein_gather.vein_frc")
    # spring terms are (|p-q| - L)*0.1 with |p-q| ~ 4: absolute scale of the summands
    assert_close(down(sim, capi.VEIN_FRC), ref, "vein spring forces", scale=max(0.5, np.abs(ref).max()))


def replay_steps(sim, cfg, variant, steps, loader, scene, physics, vein_steps=()):
    """Full stage-by-stage check over consecutive dumped steps (table history is replayed through begin.pos)."""
    summary = {}
    for step in steps:
        d = loader(cfg, variant, step)
        have_vein = step in vein_steps
        up(sim, capi.PARTICLE_POS, vec(d, "begin.pos"))
        up(sim, capi.PARTICLE_VEL, vec(d, "begin.vel"))
        sim.run_stage(capi.STAGE_GRID_PARTICLES)
        check_grid(sim, d, 0)
        if "tgrid.keys" in d:
            sim.run_stage(capi.STAGE_GRID_TRIANGLES)
            check_grid(sim, d, 1)
        if have_vein:
            check_vein_gather(sim, d)
        raced = check_springs(sim, d, physics["dt"], physics["particle_d_fact"], sim.layout)
        check_particle_collisions(sim, d)
        hits = None
        if have_vein or step == 1:
            hits = check_vein_collisions(sim, d, scene, have_vein)
            check_integration(sim, d, have_vein)
        summary[step] = dict(raced=raced, vein_hits=hits)
    return summary

"""Hand-built micro-scenes: small enough to state the expected result in closed form.

CPU (`-m "not gpu"`): the oracle against numpy restatements of the reference formulas - one blood cell at rest
(blood_cells.cu:66-120), two particles touching across a grid-cell boundary (physics.cuh:133-145), particles at the
faces / edges / corners of the grid for the 27 stencil specialisations (particle_collisions.cuh:126-268), empty and
single-occupant neighbourhoods.  GPU (`-m gpu`): libbcs, through its C ABI, against the oracle on the same scenes, plus
ray / wall cases (towards the wall, parallel to it, away from it, outside the vein)."""
import itertools

import numpy as np
import pytest

import refcheck
from conftest import capi, make_bcs, make_oracle, pkg

EDGE = 3.0                       # tetrahedron edge = rest length of its six springs; collision radius = EDGE / 6 = 0.5
RADIUS = EDGE / 6.0


def tetra_def(count):
    v = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], np.float64) * (EDGE / (2.0 * np.sqrt(2.0)))
    springs = np.array(list(itertools.combinations(range(4), 2)), np.int32)
    return pkg.CellDef(count, 4, springs, np.full(len(springs), EDGE, np.float32), v.astype(np.float32))


def micro_scene(count, use_blood_flow=0):
    vp, vi, ec, er = pkg.make_cylinder_vein(length=150.0)
    sc = pkg.Scene(user_defs=[tetra_def(count)], vein_pos=vp, vein_indices=vi, ending_centers=ec, ending_radii=er)
    sc.flags["use_blood_flow"] = use_blood_flow
    return sc


def state_from(sc, centres, velocities=None):
    """every blood cell = its model translated to `centres[c]`, moving with `velocities[c]`, zero force"""
    lay = sc.layout()
    model = lay.model[:4]
    centres = np.asarray(centres, np.float32)
    vel = np.zeros_like(centres) if velocities is None else np.asarray(velocities, np.float32)
    pos = (centres[:, None, :] + model[None, :, :]).reshape(-1, 3)
    v = np.repeat(vel, 4, axis=0)
    z = np.zeros(len(pos), np.float32)
    return {"pos_x": pos[:, 0].copy(), "pos_y": pos[:, 1].copy(), "pos_z": pos[:, 2].copy(),
            "vel_x": v[:, 0].copy(), "vel_y": v[:, 1].copy(), "vel_z": v[:, 2].copy(),
            "frc_x": z.copy(), "frc_y": z.copy(), "frc_z": z.copy()}


def run_stages(sim, st, stages):
    sim.upload_state(st)
    for s in stages:
        sim.run_stage(s)


# ------------------------------------------------------------------------------------------------ closed forms
def expected_collision_force(ph, p1, v1, r1, p2, v2, r2):
    """physics::addResilientForceOnCollision with intensity 0.5 for the particle at p1 (physics.cuh:133-145), float64"""
    rel = p1 - p2
    d = np.linalg.norm(rel)
    assert 1e-2 <= d <= r1 + r2
    dirv = rel / d
    rv = v1 - v2
    tang = rv - np.dot(rv, dirv) * dirv
    return 0.5 * (-ph["collision_spring_coeff"] * (2.0 * r1 - d) * dirv + ph["collision_damping_coeff"] * rv
                  + ph["collision_shear_coeff"] * tang)


def brute_force_candidates(pos, lay_min, cell_size, dims):
    """candidate count per particle by the reference's stencil rule (particle_collisions.cuh:117-268): per axis the
    neighbours are {0,+1} if the (unclamped) cell index is < 1, {-1,0} if it is > count-2, else {-1,0,+1}"""
    idx = np.floor((pos.astype(np.float32) - lay_min.astype(np.float32)) / np.float32(cell_size)).astype(np.int64)
    n = len(pos)
    cnt = np.zeros(n, np.int64)
    for i in range(n):
        ok = np.ones(n, bool)
        for a in range(3):
            c = idx[i, a]
            lo, hi = (0, 1) if c < 1 else ((-1, 0) if c > dims[a] - 2 else (-1, 1))
            d = idx[:, a] - c
            ok &= (d >= lo) & (d <= hi)
        ok[i] = False
        cnt[i] = ok.sum()
    return cnt


# ------------------------------------------------------------------------------------------------ scenes
def scene_cell_at_rest():
    sc = micro_scene(1)
    return sc, state_from(sc, [[3.0, -61.0, -7.0]])


def scene_touching_pair():
    """two blood cells; particle 1 of the second one sits 0.8 from particle 0 of the first, on the other side of a
    grid-cell boundary (cells are 2 units wide: x = 0 is a cell face of the r = 50 vein's grid, margins are even); all
    other pairs of the two translated tetrahedra are >= 3 apart"""
    sc = micro_scene(2)
    lay = sc.layout()
    m0, m1 = lay.model[0].astype(np.float64), lay.model[1].astype(np.float64)
    a = np.array([-0.3, -60.5, 0.5]) - m0            # particle 0 of cell 0 lands at (-0.3, -60.5, 0.5)
    b = np.array([0.5, -60.5, 0.5]) - m1             # particle 1 of cell 1 lands at ( 0.5, -60.5, 0.5): 0.8 apart
    return sc, state_from(sc, [a, b], [[4.0, -70.0, 1.0], [-3.0, -64.0, 0.0]])


def scene_grid_boundaries():
    """blood cells centred in the corner / edge / face cells of the particle grid and one deep inside; the last two
    share a neighbourhood so that non-empty candidate sets appear as well"""
    sc = micro_scene(10)
    lay = sc.layout()
    lo, hi = np.asarray(lay.grid_min, np.float64), np.asarray(lay.grid_max, np.float64)
    mid = 0.5 * (lo + hi)
    inset = 1.4          # model half-extent is 1.06: every particle stays inside the grid, some in the outermost cells
    pts = [
        [lo[0] + inset, lo[1] + inset, lo[2] + inset],      # corner (min, min, min)
        [hi[0] - inset, hi[1] - inset, hi[2] - inset],      # corner (max, max, max)
        [lo[0] + inset, hi[1] - inset, mid[2]],             # edge
        [mid[0], lo[1] + inset, hi[2] - inset],             # edge
        [lo[0] + inset, mid[1], mid[2]],                    # face x-
        [hi[0] - inset, mid[1], mid[2]],                    # face x+
        [mid[0], hi[1] - inset, mid[2]],                    # face y+
        [mid[0], mid[1], lo[2] + inset],                    # face z-
        [mid[0], mid[1], mid[2]],                           # interior
        [mid[0] + 1.7, mid[1] + 0.4, mid[2] - 0.9],         # interior, interleaved with the previous one
    ]
    return sc, state_from(sc, pts)


# ------------------------------------------------------------------------------------------------ CPU: oracle vs closed form
def test_oracle_cell_at_rest_feels_only_gravity(oracle_lib):
    sc, st = scene_cell_at_rest()
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_GRID_PARTICLES, capi.STAGE_SPRINGS])
        F = refcheck.down(orc, capi.PARTICLE_FRC)
        ph = sc.physics
        want = 0.5 * np.array([ph["gx"], ph["gy"], ph["gz"]])       # F <- (F_old + F_new) / 2, springs at rest, v = 0
        assert np.abs(F - want).max() < 2e-3, F                      # rest lengths are met to float rounding: |(len-L) k| ~ 1e-4
        c = np.stack(orc.download(capi.CELL_CENTERS), 1)
        assert np.allclose(c[0], [3.0, -61.0, -7.0], atol=1e-5)


def test_oracle_touching_pair_matches_the_closed_form(oracle_lib):
    sc, st = scene_touching_pair()
    pos = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1).astype(np.float64)
    vel = np.stack([st["vel_x"], st["vel_y"], st["vel_z"]], 1).astype(np.float64)
    d = np.linalg.norm(pos[:, None] - pos[None], axis=2) + 10.0 * np.eye(8)
    touching = np.argwhere(d <= 2 * RADIUS)
    assert sorted(map(tuple, touching)) == [(0, 5), (5, 0)], "exactly one pair within reach, by construction"
    assert int(np.floor(pos[0, 0] / 2.0)) != int(np.floor(pos[5, 0] / 2.0)), "the pair straddles a grid-cell face"
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_GRID_PARTICLES])
        cnt, _, hits = orc.debug_candidates()
        assert hits.tolist() == [1, 0, 0, 0, 0, 1, 0, 0]
        assert cnt[0] >= 4 and cnt[5] >= 4                           # own blood cell (3) + at least the partner
        orc.run_stage(capi.STAGE_PARTICLE_COLLISIONS)
        F = refcheck.down(orc, capi.PARTICLE_FRC).astype(np.float64)
    r = float(RADIUS)
    want0 = expected_collision_force(sc.physics, pos[0], vel[0], r, pos[5], vel[5], r)
    want5 = expected_collision_force(sc.physics, pos[5], vel[5], r, pos[0], vel[0], r)
    assert np.abs(F[0] - want0).max() < 1e-4 * np.abs(want0).max(), (F[0], want0)
    assert np.abs(F[5] - want5).max() < 1e-4 * np.abs(want5).max(), (F[5], want5)
    assert np.abs(F[[1, 2, 3, 4, 6, 7]]).max() == 0.0               # nobody else is touched (forces start at zero)


def test_oracle_stencils_at_the_grid_boundaries(oracle_lib):
    sc, st = scene_grid_boundaries()
    pos = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1)
    with make_oracle(oracle_lib, sc) as orc:
        lay = orc.layout
        dims = list(lay.grid_dims)
        run_stages(orc, st, [capi.STAGE_GRID_PARTICLES])
        keys, ids = orc.grid(0)
        cnt, _, hits = orc.debug_candidates()
        idx = np.floor((pos - np.asarray(lay.grid_min, np.float32)) / np.float32(2.0)).astype(np.int64)
        assert idx.min() >= 0 and np.all(idx.max(0) <= np.asarray(dims) - 1)
        assert (idx.min(0) == 0).all() and (idx.max(0) == np.asarray(dims) - 1).all(), "outermost cells are occupied on every axis"
        want_key = (idx[:, 2] * dims[1] + idx[:, 1]) * dims[0] + idx[:, 0]
        assert np.array_equal(keys, want_key[ids])
        want = brute_force_candidates(pos, np.asarray(lay.grid_min), 2.0, dims)
        assert np.array_equal(cnt, want), (cnt, want)
        assert cnt[:32].min() >= 1 and cnt[32:].max() >= 4           # isolated cells see their mates; the interleaved pair more
        assert hits.sum() == 0 or hits.sum() % 2 == 0


def test_oracle_springs_closed_form(oracle_lib):
    """gatherForcesKernel (blood_cells.cu:66-120) + physics.cuh:24-27,53-78,102-120 restated in float64 numpy on three
    tetrahedra: one stretched by 10 % at rest (pure spring pull towards the centroid + gravity), one perturbed with
    velocities and old forces, one blown up by 60 % so that the big-cell brake (ratio > 1.5) is active"""
    sc = micro_scene(3)
    ph = sc.physics
    lay = sc.layout()
    model = lay.model[:4].astype(np.float64)
    centres = np.array([[0.0, -50.0, 0.0], [10.0, -70.0, -5.0], [-12.0, -90.0, 6.0]])
    rng = np.random.default_rng(11)
    pos = np.concatenate([centres[0] + 1.1 * model, centres[1] + model + rng.normal(0, 0.15, (4, 3)), centres[2] + 1.6 * model])
    vel = np.concatenate([np.zeros((4, 3)), rng.normal(0, 5.0, (4, 3)) + [0, -70, 0], rng.normal(0, 2.0, (4, 3)) + [0, -60, 0]])
    frc = np.concatenate([np.zeros((4, 3)), rng.normal(0, 40.0, (4, 3)), rng.normal(0, 10.0, (4, 3))])
    st = {f"{a}_{c}": arr[:, i].astype(np.float32).copy() for a, arr in (("pos", pos), ("vel", vel), ("frc", frc)) for i, c in enumerate("xyz")}
    pos, vel, frc = (np.stack([st[f"{a}_x"], st[f"{a}_y"], st[f"{a}_z"]], 1).astype(np.float64) for a in ("pos", "vel", "frc"))
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_SPRINGS])
        got = refcheck.down(orc, capi.PARTICLE_FRC).astype(np.float64)
    dt, k, d = ph["dt"], ph["particle_k_sniff"], ph["particle_d_fact"]
    G = np.array([ph["gx"], ph["gy"], ph["gz"]])
    init_r = np.linalg.norm(model - model.mean(0), axis=1)
    want = np.empty_like(frc)
    braked = 0
    for c in range(3):
        P = slice(4 * c, 4 * c + 4)
        p, v, f = pos[P], vel[P], frc[P]
        centre = p.sum(0) / 4.0
        for i in range(4):
            new = np.zeros(3)
            for j in range(4):
                if j == i:
                    continue
                dP = p[i] - p[j]
                n = dP / np.linalg.norm(dP)
                dv = v[i] - v[j] + dt * (f[i] - f[j])
                new += ((np.linalg.norm(dP) - EDGE) * k + np.dot(n, dv) * d) * (-n)
            ratio = np.linalg.norm(p[i] - centre) / init_r[i]
            brake = ratio * ph["big_particle_braking_intensity"] if ratio > ph["max_cell_size_factor_before_brake"] else 1.0
            braked += ratio > ph["max_cell_size_factor_before_brake"]
            new += G - ph["viscous_damping"] * brake * v[i]
            want[4 * c + i] = (f[i] + new) / 2.0
    assert braked == 4, "exactly the blown-up cell brakes"
    assert np.abs(got - want).max() < 2e-5 * np.abs(want).max(), np.abs(got - want).max()
    # the stretched cell at rest: 0.3 * k towards each mate = sqrt(6) * 0.3 * k towards the centroid, plus gravity, halved
    inward = (centres[0] - pos[:4]) / np.linalg.norm(centres[0] - pos[:4], axis=1)[:, None]
    assert np.abs(got[:4] - 0.5 * (np.sqrt(6.0) * 0.1 * EDGE * k * inward + G)).max() < 1e-3


def test_oracle_vein_springs_closed_form(oracle_lib):
    """VeinTriangles::gatherForcesFromNeighbors (vein_triangles.cu:126-154, physics.cuh:38-41) + the vertex integrator
    (:88-117): the neighbour slots are rebuilt here from the triangle list by the reference's rule (vein_factory.hpp:
    130-174: every triangle pushes both other vertices - each mesh edge twice -, the list is sorted and cut to 9, SURVEY
    Q10) and the force on a displaced patch of the wall is restated in float64"""
    sc = micro_scene(1)
    ph = sc.physics
    vp0, vi = sc.vein_pos.astype(np.float64), sc.vein_indices.astype(np.int64)
    V = len(vp0)
    nbrs = [[] for _ in range(V)]
    for a, b, c in vi:
        nbrs[a] += [b, c]; nbrs[b] += [a, c]; nbrs[c] += [a, b]
    slots = np.full((9, V), -1, np.int64)
    for v in range(V):
        lst = sorted(nbrs[v])[:9]
        slots[:len(lst), v] = lst
    with make_oracle(oracle_lib, sc) as orc:
        ids = orc.table(capi.TABLE_VEIN_NBR_IDS).reshape(9, V)
        rest = orc.table(capi.TABLE_VEIN_NBR_LEN).reshape(9, V).astype(np.float64)
        assert np.array_equal(ids, slots), "neighbour slots: duplicates kept, sorted, cut to 9"
        want_rest = np.where(slots >= 0, np.linalg.norm(vp0[np.maximum(slots, 0)] - vp0[None, :, :], axis=2), 0.0)
        assert np.abs(rest - want_rest)[slots >= 0].max() < 1e-4
        # displace a patch of the wall and give it velocities
        rng = np.random.default_rng(2)
        patch = np.arange(1500, 1530)
        vp = sc.vein_pos.copy(); vv = np.zeros_like(vp)
        vp[patch] += rng.normal(0, 0.3, (len(patch), 3)).astype(np.float32)
        vv[patch] = rng.normal(0, 2.0, (len(patch), 3)).astype(np.float32)
        orc.upload_state(state_from(sc, [[0.0, -60.0, 0.0]]))
        refcheck.up(orc, capi.VEIN_POS, vp)
        refcheck.up(orc, capi.VEIN_VEL, vv)
        orc.run_stage(capi.STAGE_VEIN_GATHER)
        F = refcheck.down(orc, capi.VEIN_FRC).astype(np.float64)
        orc.run_stage(capi.STAGE_INTEGRATE_VEIN)
        x1, v1, F1 = (refcheck.down(orc, w).astype(np.float64) for w in (capi.VEIN_POS, capi.VEIN_VEL, capi.VEIN_FRC))
    p, w = vp.astype(np.float64), vv.astype(np.float64)
    want = np.zeros_like(p)
    # every vertex that moved or LISTS a vertex that moved (the cut to 9 slots makes the lists asymmetric)
    touched = set(patch.tolist()) | set(np.nonzero(np.isin(slots, patch).any(0))[0].tolist())
    for v in touched:
        for s_ in range(9):
            q = slots[s_, v]
            if q < 0:
                continue
            dP = p[v] - p[q]
            n = dP / np.linalg.norm(dP)
            f = (np.linalg.norm(dP) - rest[s_, v]) * ph["vein_k_sniff"] + np.dot(n, w[v] - w[q]) * ph["vein_d_fact"]
            want[v] += f * (-n)
    far = np.setdiff1d(np.arange(V), np.fromiter(touched, int))
    assert np.abs(F[far]).max() < 1e-4, "the undisturbed wall is at rest"
    assert np.abs(F - want).max() < 1e-4 * np.abs(want).max() + 1e-5, np.abs(F - want).max()
    dt = ph["dt"]
    assert np.abs(v1 - (w + dt * F)).max() < 1e-6 and np.abs(x1 - (p + dt * (w + dt * F))).max() < 1e-4
    assert np.abs(F1).max() == 0.0, "the vertex integrator clears the forces"


def _first_hit_restatement(sc, lay, pos, vel):
    """calculateSideCollisions (vein_collisions.cuh:60-93) in numpy: the triangle grid is the uniform grid of the triangle
    centres (cell size 25, same bounds, stable order by (cell, triangle id)); a particle walks the <= 27 cells around its
    own (x outer, y, z inner) and returns the FIRST triangle its ray hits (Moeller-Trumbore, vein_collisions.cu:11-45),
    near or far.  float64 arithmetic: the caller keeps borderline rays out."""
    vp, vi = sc.vein_pos.astype(np.float64), sc.vein_indices.astype(np.int64)
    gmin = np.asarray(lay.grid_min, np.float64)
    dims = np.asarray(lay.tri_grid_dims, np.int64)
    cs = float(sc.tri_cell_size[0])
    cent = (vp[vi[:, 0]] + vp[vi[:, 1]] + vp[vi[:, 2]]) / 3.0
    tc = np.floor((cent - gmin) / cs).astype(np.int64)
    tkey = (tc[:, 2] * dims[1] + tc[:, 1]) * dims[0] + tc[:, 0]
    order = np.argsort(tkey, kind="stable")
    skey = tkey[order]
    out = np.full(len(pos), -1, np.int64)
    tt = np.full(len(pos), np.inf)
    for i, (p, v) in enumerate(zip(pos.astype(np.float64), vel.astype(np.float64))):
        d = v / np.linalg.norm(v)
        c = np.floor((p - gmin) / cs).astype(np.int64)
        rng_ = [(0, 1) if c[a] < 1 else ((-1, 0) if c[a] > dims[a] - 2 else (-1, 1)) for a in range(3)]
        cell = (c[2] * dims[1] + c[1]) * dims[0] + c[0]
        done = False
        for x in range(rng_[0][0], rng_[0][1] + 1):
            for y in range(rng_[1][0], rng_[1][1] + 1):
                for z in range(rng_[2][0], rng_[2][1] + 1):
                    nbr = cell + z * dims[0] * dims[1] + y * dims[0] + x
                    lo, hi = np.searchsorted(skey, nbr, "left"), np.searchsorted(skey, nbr, "right")
                    for tri in order[lo:hi]:
                        a, b, cc = vp[vi[tri]]
                        e1, e2 = b - a, cc - a
                        h = np.cross(d, e2)
                        det = e1 @ h
                        if abs(det) < 1e-6:
                            continue
                        f = 1.0 / det
                        sv = p - a
                        u = f * (sv @ h)
                        if u < 0 or u > 1:
                            continue
                        q = np.cross(sv, e1)
                        w = f * (d @ q)
                        if w < 0 or u + w > 1:
                            continue
                        t = f * (e2 @ q)
                        if t > 1e-6:
                            out[i], tt[i], done = tri, t, True
                            break
                    if done: break
                if done: break
            if done: break
    return out, tt


def test_oracle_first_hit_traversal_restated(oracle_lib):
    """the oracle's vein search against the numpy restatement above: 60 blood cells scattered through the lumen with
    random directions (most rays leave through a triangle of the 27-cell neighbourhood, some find none)"""
    sc = micro_scene(60)
    rng = np.random.default_rng(21)
    ang, rad = rng.uniform(0, 2 * np.pi, 60), 48.0 * np.sqrt(rng.uniform(0, 1, 60))
    centres = np.stack([rad * np.cos(ang), rng.uniform(-140.0, -10.0, 60), rad * np.sin(ang)], 1)
    dirs = rng.normal(0, 1, (60, 3))
    st = state_from(sc, centres, 70.0 * dirs / np.linalg.norm(dirs, axis=1)[:, None])
    pos = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1)
    vel = np.stack([st["vel_x"], st["vel_y"], st["vel_z"]], 1)
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_GRID_PARTICLES])
        tri, t = orc.debug_vein_hits()
        want, want_t = _first_hit_restatement(sc, orc.layout, pos, vel)
    same = tri == want
    assert same.mean() >= 0.98, f"{(~same).sum()} of {len(tri)} first-hit triangles differ"   # float32 vs float64 on edge-grazing rays
    assert (want >= 0).sum() >= 60 and (want < 0).sum() >= 20, "both outcomes are exercised"
    hit = same & (want >= 0)
    assert np.abs(t[hit] - want_t[hit]).max() < 1e-3


def test_oracle_integration_closed_form(oracle_lib):
    """propagateParticleForcesKernel (blood_cells.cu:155-179): v1 = v0 + dt F, x += dt/2 (v1 + v0), F untouched"""
    sc = micro_scene(2)
    st = state_from(sc, [[1.0, -60.0, 2.0], [-9.0, -75.0, 4.0]], [[3.0, -70.0, 1.0], [-2.0, -64.0, 5.0]])
    rng = np.random.default_rng(3)
    F = rng.normal(0.0, 200.0, (8, 3)).astype(np.float32)
    st["frc_x"], st["frc_y"], st["frc_z"] = F[:, 0].copy(), F[:, 1].copy(), F[:, 2].copy()
    x0 = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1).astype(np.float64)
    v0 = np.stack([st["vel_x"], st["vel_y"], st["vel_z"]], 1).astype(np.float64)
    dt = sc.physics["dt"]
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_INTEGRATE_PARTICLES])
        x, v, f = (refcheck.down(orc, w).astype(np.float64) for w in (capi.PARTICLE_POS, capi.PARTICLE_VEL, capi.PARTICLE_FRC))
    v1 = v0 + dt * F
    assert np.abs(v - v1).max() < 1e-5 * np.abs(v1).max()
    assert np.abs(x - (x0 + 0.5 * dt * (v1 + v0))).max() < 1e-5 * np.abs(x0).max()
    assert np.array_equal(f.astype(np.float32), F), "the integrator leaves the force alone (SURVEY Q6)"


def test_oracle_vein_end_respawn(oracle_lib):
    """HandleVeinEnd (vein_end.cu:12-173): a blood cell with ANY particle beyond a threshold is respawned as a whole at
    y = minSpawnY + (model_k - model_0).y with the initial velocity; its x / z offsets keep the model's shape; the
    other cell is left alone"""
    sc = micro_scene(2, use_blood_flow=1)
    lay = sc.layout()
    lower = float(lay.grid_min[1]) + sc.physics["grid_y_margin"] / 2.0
    st = state_from(sc, [[0.0, lower + 0.5, 0.0], [5.0, -60.0, 5.0]], [[1.0, -70.0, 0.0], [0.0, -64.0, 0.0]])
    y = st["pos_y"][:4]
    assert (y <= lower).any() and (y > lower).any(), "only some particles of the first cell are past the threshold"
    before = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1)
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_VEIN_END])
        x, v = refcheck.down(orc, capi.PARTICLE_POS), refcheck.down(orc, capi.PARTICLE_VEL)
        assert orc.stats()["teleported_cells"] == 1
    model = lay.model[:4]
    ph = sc.physics
    assert np.allclose(x[:4, 1], ph["min_spawn_y"] + model[:, 1] - model[0, 1], atol=1e-5)
    assert np.allclose(x[:4] - x[0], model - model[0], atol=1e-5), "the respawned cell keeps the model's shape"
    assert np.abs(x[0, [0, 2]]).max() <= 0.6 * ph["cylinder_radius"] + 1e-4, "x, z = (U - 0.5) * 1.2 * cylinderRadius"
    assert np.allclose(v[:4], [ph["init_velocity_x"], ph["init_velocity_y"], ph["init_velocity_z"]])
    assert np.array_equal(x[4:], before[4:]) and np.allclose(v[4:], [0.0, -64.0, 0.0])


def test_oracle_wall_hit_closed_form(oracle_lib):
    """detectVeinCollisions (vein_collisions.cu:234-276) for one blood cell flying at the wall: given the triangle and
    the distance the traversal reports, the effect is closed form - reaction force F -= (F.n) n / (n.n), velocity
    v <- 0.96 |v| reflect(dir, n), wall splat 0.005 v spread over the triangle's vertices by barycentric weights"""
    sc = micro_scene(1)
    st = state_from(sc, [[46.5, -60.0, 0.0]], [[80.0, -5.0, 3.0]])
    rng = np.random.default_rng(5)
    F0 = rng.normal(0.0, 50.0, (4, 3)).astype(np.float32)
    st["frc_x"], st["frc_y"], st["frc_z"] = F0[:, 0].copy(), F0[:, 1].copy(), F0[:, 2].copy()
    pos = np.stack([st["pos_x"], st["pos_y"], st["pos_z"]], 1).astype(np.float64)
    vel = np.stack([st["vel_x"], st["vel_y"], st["vel_z"]], 1).astype(np.float64)
    ph = sc.physics
    with make_oracle(oracle_lib, sc) as orc:
        run_stages(orc, st, [capi.STAGE_GRID_PARTICLES])
        tri, t = orc.debug_vein_hits()
        orc.run_stage(capi.STAGE_VEIN_COLLISIONS)
        v = refcheck.down(orc, capi.PARTICLE_VEL).astype(np.float64)
        F = refcheck.down(orc, capi.PARTICLE_FRC).astype(np.float64)
        vf = refcheck.down(orc, capi.VEIN_FRC).astype(np.float64)
        assert orc.stats()["vein_hits"] == int((t <= ph["vein_impact_distance"]).sum())
    vp, vi = sc.vein_pos.astype(np.float64), sc.vein_indices
    want_vf = np.zeros_like(vf)
    acted = 0
    for k in range(4):
        assert tri[k] >= 0, "every particle of the cell flies at the wall"
        a, b, c = vp[vi[tri[k]]]
        d = vel[k] / np.linalg.norm(vel[k])
        hit = pos[k] + float(t[k]) * d
        # the reported hit point lies in the reported triangle's plane, inside it
        n = np.cross(c - a, b - a)
        n /= np.linalg.norm(n)
        assert abs(np.dot(hit - a, n)) < 1e-3
        if t[k] > ph["vein_impact_distance"]:
            assert np.allclose(v[k], vel[k]) and np.allclose(F[k], F0[k])
            continue
        acted += 1
        refl = d - 2.0 * np.dot(d, n) * n
        want_v = ph["velocity_collision_damping"] * np.linalg.norm(vel[k]) * refl
        assert np.abs(v[k] - want_v).max() < 1e-4 * np.linalg.norm(vel[k]), (v[k], want_v)
        want_F = F0[k] - np.dot(F0[k], n) * n
        assert np.abs(F[k] - want_F).max() < 1e-4 * np.abs(F0[k]).max(), (F[k], want_F)
        # the hit point is inside the triangle (true barycentric coordinates) ...
        T = np.stack([a, b, c], 1)
        w = np.linalg.lstsq(np.vstack([T, np.ones(3)]), np.append(hit, 1.0), rcond=None)[0]
        assert w.min() > -1e-3 and abs(w.sum() - 1.0) < 1e-6
        # ... but the splat weights are calculateBaricentric's (vein_collisions.cu:47-61), whose second edge is v2 - v1,
        # not v2 - v0: weights (bx, by, 1 - bx - by) on (v0, v1, v2)
        e0, e1, e2 = b - a, c - b, hit - a
        d00, d01, d11, d20, d21 = e0 @ e0, e0 @ e1, e1 @ e1, e2 @ e0, e2 @ e1
        den = d00 * d11 - d01 * d01
        bx, by = (d11 * d20 - d01 * d21) / den, (d00 * d21 - d01 * d20) / den
        for wgt, vid in zip((bx, by, 1.0 - bx - by), vi[tri[k]]):
            want_vf[vid] += wgt * ph["vein_collision_force_intensity"] * vel[k]
    assert acted >= 2
    assert np.abs(vf - want_vf).max() < 2e-3 * np.abs(want_vf).max(), np.abs(vf - want_vf).max()


# ------------------------------------------------------------------------------------------------ GPU: libbcs vs oracle
ALL_STAGES = [capi.STAGE_GRID_PARTICLES, capi.STAGE_VEIN_GATHER, capi.STAGE_SPRINGS, capi.STAGE_PARTICLE_COLLISIONS,
              capi.STAGE_VEIN_COLLISIONS, capi.STAGE_INTEGRATE_PARTICLES, capi.STAGE_INTEGRATE_VEIN, capi.STAGE_VEIN_END]


def _gpu_vs_oracle(bcs_lib, oracle_lib, sc, st, semantics, steps=2):
    with make_bcs(sc, semantics) as sim, make_oracle(oracle_lib, sc, semantics) as orc:
        sim.upload_state(st)
        orc.upload_state(st)
        for step in range(steps):
            sim.run_stage(capi.STAGE_GRID_PARTICLES)
            orc.run_stage(capi.STAGE_GRID_PARTICLES)
            ka, ia = sim.grid(0)
            kb, ib = orc.grid(0)
            assert np.array_equal(ka, kb) and np.array_equal(ia, ib), f"step {step}: sorted grid"
            for name, x, y in zip(("count", "checksum", "hits"), sim.debug_candidates(), orc.debug_candidates()):
                assert np.array_equal(x, y), f"step {step}: candidate {name}"
            ta, tb = sim.debug_vein_hits(), orc.debug_vein_hits()
            assert np.array_equal(ta[0], tb[0]), f"step {step}: first-hit triangles"
            for s in ALL_STAGES[1:]:
                sim.run_stage(s)
                orc.run_stage(s)
            for which in (capi.PARTICLE_FRC, capi.PARTICLE_VEL, capi.PARTICLE_POS):
                refcheck.assert_close(refcheck.down(sim, which), refcheck.down(orc, which), f"step {step} array {which}")
        return sim.stats()


@pytest.mark.gpu
@pytest.mark.parametrize("semantics", [capi.SEM_CLEAN, capi.SEM_REFERENCE])
@pytest.mark.parametrize("scene", [scene_cell_at_rest, scene_touching_pair, scene_grid_boundaries])
def test_micro_scenes_gpu_vs_oracle(bcs_lib, oracle_lib, scene, semantics):
    sc, st = scene()
    _gpu_vs_oracle(bcs_lib, oracle_lib, sc, st, semantics)


@pytest.mark.gpu
def test_ray_cases_against_the_wall(bcs_lib, oracle_lib):
    """one blood cell per case, 2 units from the r = 50 wall (inside veinImpactDistance = 6) or elsewhere: flying at the
    wall, along it, away from it, through the lumen's axis, and outside the vein altogether"""
    sc = micro_scene(6)
    centres = [[47.0, -60.0, 0.0], [47.0, -70.0, 0.0], [47.0, -80.0, 0.0], [0.0, -90.0, 0.0], [0.0, -100.0, 46.5], [70.0, -60.0, 0.0]]
    vels = [[80.0, -5.0, 0.0], [0.0, -80.0, 0.0], [-80.0, -5.0, 0.0], [0.0, -80.0, 0.0], [3.0, -10.0, 70.0], [-80.0, 0.0, 0.0]]
    st = state_from(sc, centres, vels)
    stats = _gpu_vs_oracle(bcs_lib, oracle_lib, sc, st, capi.SEM_CLEAN, steps=3)
    assert stats["vein_hits"] >= 4, "the cells flying at the wall must have hit it"

// Device-side building blocks of the particle vs. vein-wall collision stage, shared by vein.cu (per-step refit,
// reference-compatible and exhaustive paths) and wall.cu (lazily rebuilt wall grid, production path).
#pragma once
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "pair_device.cuh"
#include "kernels.cuh"

namespace bcs {

constexpr float BOX_PAD = 0.05f;   // covers float error of a Moeller-Trumbore "hit" that lies marginally outside its triangle

struct RayHit {
    float t;
    float3 normal;
    float3 refl;
    int tri;
};

// realCollisionDetection (vein_collisions.cu:11-45) on a packed triangle
__device__ __forceinline__ bool ray_triangle(const float3 origin, const float3 dir, const TriPacked& tp, RayHit& h)
{
    constexpr float EPS = 0.000001f;
    const float3 v0 = f3(tp.a.x, tp.a.y, tp.a.z);
    const float3 edge1 = f3(tp.a.w, tp.b.x, tp.b.y);
    const float3 edge2 = f3(tp.b.z, tp.b.w, tp.c.x);
    const float3 hh = cross(dir, edge2);
    const float a = dot(edge1, hh);
    if (a > -EPS && a < EPS) return false;
    const float f = 1 / a;
    const float3 s = origin - v0;
    const float u = f * dot(s, hh);
    if (u < 0 || u > 1) return false;
    const float3 q = cross(s, edge1);
    const float v = f * dot(dir, q);
    if (v < 0 || u + v > 1) return false;
    const float t = f * dot(edge2, q);
    if (t > EPS) {
        h.t = t;
        h.normal = normalize(cross(edge2, edge1));
        h.refl = dir - (2 * dot(dir, h.normal)) * h.normal;
        h.tri = __float_as_int(tp.c.y);
        return true;
    }
    return false;
}

// calculateBaricentric (vein_collisions.cu:47-61); note e1 = v2 - v1 there
__device__ __forceinline__ float3 barycentric(float3 point, float3 v0, float3 v1, float3 v2)
{
    const float3 e0 = v1 - v0, e1 = v2 - v1, e2 = point - v0;
    const float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(e2, e0), d21 = dot(e2, e1);
    const float denom = d00 * d11 - d01 * d01;
    float3 b;
    b.x = (d11 * d20 - d01 * d21) / denom;
    b.y = (d00 * d21 - d01 * d20) / denom;
    b.z = 1.0f - b.x - b.y;
    return b;
}

__device__ __forceinline__ void tri_stencil_range(unsigned id, int count, int& lo, int& hi)
{
    // vein_collisions.cu:82-230: ids are unsigned, `id > count - 2` is an unsigned comparison
    if (id < 1u) { lo = 0; hi = 1; }
    else if (id > (unsigned)(count - 2)) { lo = -1; hi = 0; }
    else { lo = -1; hi = 1; }
}

// packed (v0, e1, e2) of the triangle in sorted slot i: from the per-step refit, or - wall-grid path, where nothing
// is repacked per step - gathered from the live vertices (same float operations, same values)
__device__ __forceinline__ TriPacked load_tri(const VeinCollideArgs& a, int i)
{
    TriPacked tp;
    if (a.liveTris) {
        const int4 sv = __ldg(a.wall.slotVerts + i);   // vertex ids + triangle id of the slot: one load level before the vertices
        const int tri = sv.w;
        const float3 v0 = xyz(a.vpos[sv.x]), v1 = xyz(a.vpos[sv.y]), v2 = xyz(a.vpos[sv.z]);
        const float3 e1 = v1 - v0, e2 = v2 - v0;
        tp.a = make_float4(v0.x, v0.y, v0.z, e1.x);
        tp.b = make_float4(e1.y, e1.z, e2.x, e2.y);
        tp.c = make_float4(e2.z, __int_as_float(tri), 0.f, 0.f);
    } else {
        tp.a = a.tris[i].a; tp.b = a.tris[i].b; tp.c = a.tris[i].c;
    }
    return tp;
}

// does the box [lo,hi] overlap the box q?
__device__ __forceinline__ bool box_overlap(const Aabb& q, float lox, float loy, float loz, float hix, float hiy, float hiz)
{
    return q.lox <= hix && q.hix >= lox && q.loy <= hiy && q.hiy >= loy && q.loz <= hiz && q.hiz >= loz;
}

// does the half-line o + t*d, t >= 0, touch the (already padded) box?  Conservative slab test.
__device__ __forceinline__ bool ray_box(const Aabb& q, const float3 o, const float3 d)
{
    if (q.lox > q.hix) return false;   // empty (never fitted) box
    float tn = 0.f, tf = 3e38f;
    const float lo[3] = {q.lox, q.loy, q.loz}, hi[3] = {q.hix, q.hiy, q.hiz};
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (fabsf(dd[k]) < 1e-12f) {
            if (oo[k] < lo[k] || oo[k] > hi[k]) return false;
        } else {
            const float inv = 1.0f / dd[k];
            const float t1 = (lo[k] - oo[k]) * inv, t2 = (hi[k] - oo[k]) * inv;
            tn = fmaxf(tn, fminf(t1, t2));
            tf = fminf(tf, fmaxf(t1, t2));
        }
    }
    return tn <= tf * 1.00001f + 1e-4f;
}

// ---- first hit in the reference's traversal order -------------------------------------------------------
// Straight traversal: x outer, y, z inner, sorted triangles inside a cell; the FIRST accepted triangle
// wins, however far away it is (vein_collisions.cuh:66-91).
template <bool STATS>
__device__ bool first_hit_naive(const VeinCollideArgs& a, const float3 pos, const float3 dir, int cell, int x0, int x1, int y0, int y1,
                                int z0, int z1, RayHit& h, unsigned long long& tests)
{
    const GridDev& g = a.tgrid;
    const int plane = g.nx * g.ny;
    for (int x = x0; x <= x1; ++x)
        for (int y = y0; y <= y1; ++y)
            for (int z = z0; z <= z1; ++z) {
                const int c = cell + z * plane + y * g.nx + x;
                if (c < 0 || c >= g.cells) continue;
                const int s = a.cellStart[c], e = a.cellEnd[c];
                for (int i = s; i <= e; ++i) {
                    if (STATS) ++tests;
                    if (ray_triangle(pos, dir, load_tri(a, i), h)) return true;
                }
            }
    return false;
}

// Same RESULT for everything the stage can observe, without testing ~800 triangles per particle.
// The stage only acts when the first hit in traversal order lies within veinImpactDistance (d^2 <= 36,
// vein_collisions.cu:237); a far first hit does nothing, exactly like no hit.  So:
//   phase A  find the first triangle in traversal order that the ray hits with t <= 6 (+margin): only
//            cells / slot groups whose box overlaps the box of that short segment are touched; for a
//            particle in the bulk of the vein that is 27 box tests and nothing else;
//   phase B  (rare: only particles about to touch the wall) make sure no EARLIER triangle in traversal
//            order is hit further away - such a far hit would have been returned first by the reference
//            and would mask the near one.  Uses half-line vs box culling.
// Returns true iff the reference's traversal ends on a triangle within reach; h = that hit.
template <bool STATS>
__device__ bool first_hit_fast(const VeinCollideArgs& a, const float3 pos, const float3 dir, int cell, int x0, int x1, int y0, int y1,
                               int z0, int z1, RayHit& h, unsigned long long& tests)
{
    const GridDev& g = a.tgrid;
    const int plane = g.nx * g.ny;
    const float reach = a.phys.impactNear;
    const float3 tip = pos + reach * dir;
    const float slx = fminf(pos.x, tip.x), shx = fmaxf(pos.x, tip.x);
    const float sly = fminf(pos.y, tip.y), shy = fmaxf(pos.y, tip.y);
    const float slz = fminf(pos.z, tip.z), shz = fmaxf(pos.z, tip.z);
    // ---- phase A
    int hitOrder = -1, hitSlot = -1;   // position of the near hit in traversal order: (cell ordinal, slot)
    int ord = 0;
    for (int x = x0; x <= x1 && hitOrder < 0; ++x)
        for (int y = y0; y <= y1 && hitOrder < 0; ++y)
            for (int z = z0; z <= z1 && hitOrder < 0; ++z, ++ord) {
                const int c = cell + z * plane + y * g.nx + x;
                if (c < 0 || c >= g.cells) continue;
                if (!box_overlap(a.cellBox[c], slx, sly, slz, shx, shy, shz)) continue;
                const int s = a.cellStart[c], e = a.cellEnd[c];
                for (int gi = s >> 3; gi <= (e >> 3) && hitOrder < 0; ++gi) {
                    if (!box_overlap(a.groupBox[gi], slx, sly, slz, shx, shy, shz)) continue;
                    const int i0 = max(s, gi << 3), i1 = min(e, (gi << 3) + 7);
                    for (int i = i0; i <= i1; ++i) {
                        if (STATS) ++tests;
                        RayHit cand;
                        if (ray_triangle(pos, dir, load_tri(a, i), cand) && cand.t <= reach) {
                            h = cand; hitOrder = ord; hitSlot = i;
                            break;
                        }
                    }
                }
            }
    if (hitOrder < 0) return false;
    // ---- phase B: any hit (necessarily beyond reach) earlier in traversal order?
    ord = 0;
    for (int x = x0; x <= x1; ++x)
        for (int y = y0; y <= y1; ++y)
            for (int z = z0; z <= z1; ++z, ++ord) {
                if (ord > hitOrder) return true;
                const int c = cell + z * plane + y * g.nx + x;
                if (c < 0 || c >= g.cells) continue;
                if (!ray_box(a.cellBox[c], pos, dir)) continue;
                const int s = a.cellStart[c];
                const int e = (ord == hitOrder) ? hitSlot - 1 : a.cellEnd[c];
                for (int gi = s >> 3; gi <= (e >> 3); ++gi) {
                    if (e < s) break;
                    if (!ray_box(a.groupBox[gi], pos, dir)) continue;
                    const int i0 = max(s, gi << 3), i1 = min(e, (gi << 3) + 7);
                    for (int i = i0; i <= i1; ++i) {
                        if (STATS) ++tests;
                        RayHit far;
                        if (ray_triangle(pos, dir, load_tri(a, i), far)) return false;   // masked by an earlier (far) triangle
                    }
                }
            }
    return true;
}

// Wall splats are summed ORDER-INDEPENDENTLY: a float atomicAdd makes the vertex force depend on which particle's add
// lands first (two splats on top of the spring force already differ in the last bit), which breaks bit-identical
// restarts and the N-rank == 1-rank equality.  Each contribution is therefore added as a 64-bit fixed-point integer
// (unit 2^-40: exact for every float of magnitude >= 2^-16, range +-8.3e6) and folded into the float force once, by
// the vertex integrator (or vein_fold_splats before a read-back).  vfrc.w flags vertices with parked splats.
constexpr float SPLAT_SCALE = 1099511627776.0f;        // 2^40
constexpr float SPLAT_UNSCALE = 1.0f / 1099511627776.0f;
__device__ __forceinline__ void splat_add(const VeinCollideArgs& a, unsigned v, float w, const float3 ds)
{
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(a.vsplat) + 3 * (size_t)v;
    atomicAdd(acc, (unsigned long long)__float2ll_rn(w * ds.x * SPLAT_SCALE));
    atomicAdd(acc + 1, (unsigned long long)__float2ll_rn(w * ds.y * SPLAT_SCALE));
    atomicAdd(acc + 2, (unsigned long long)__float2ll_rn(w * ds.z * SPLAT_SCALE));
    a.vfrc[v].w = 1.0f;
}
// the parked splats of vertex v as floats; clears the accumulators
__device__ __forceinline__ float3 splat_take(long long* vsplat, int v)
{
    long long* acc = vsplat + 3 * (size_t)v;
    const float3 r = make_float3(__ll2float_rn(acc[0]) * SPLAT_UNSCALE, __ll2float_rn(acc[1]) * SPLAT_UNSCALE, __ll2float_rn(acc[2]) * SPLAT_UNSCALE);
    acc[0] = 0; acc[1] = 0; acc[2] = 0;
    return r;
}

// what the stage does once the traversal has ended on triangle h (vein_collisions.cu:234-276)
__device__ __forceinline__ void vein_apply_hit(const VeinCollideArgs& a, int pid, const float4 p4, const float4 v4, const float3 dir,
                                               const RayHit& h, bool splatOnly = false)
{
    const PhysDev& ph = a.phys;
    const float3 pos = xyz(p4), velocity = xyz(v4);
    const bool hit = true;
    // relativePosition = pos - (pos + t*dir), evaluated literally (vein_collisions.cu:234)
    const float3 rel = pos - (pos + h.t * dir);
    const float d2 = length_squared(rel);
    if (a.apply && hit && d2 <= ph.impact2) {
        if (!splatOnly && d2 > ph.minForce2) {
            float4 F4 = a.frc[pid];
            if (a.pairAcc) {
                // the collision pass parked this particle's pair forces in fixed point (pairs.cu, deferred fold): fold them
                // before the reaction force reads the accumulated force, exactly as the stand-alone fold pass would have
                long long* acc = a.pairAcc + 3 * (size_t)pid;
                const long long sx = acc[0], sy = acc[1], sz = acc[2];
                if ((sx | sy | sz) != 0) {
                    F4.x += fx_value(sx); F4.y += fx_value(sy); F4.z += fx_value(sz);
                    acc[0] = 0; acc[1] = 0; acc[2] = 0;
                }
            }
            const float3 F = xyz(F4);
            float3 add;
            if (ph.reactionForce) {
                add = ((-1.0f * dot(F, h.normal)) * h.normal) / dot(h.normal, h.normal);
            } else {
                int t = 0;
                while (t + 1 < a.types.n && pid >= a.types.t[t + 1].pStart) ++t;
                const float radius = __ldg(a.collR + a.types.t[t].mStart + (pid - a.types.t[t].pStart) % a.types.t[t].P);
                // physics::addResilientForceOnCollision(relativePosition, velocity, d2, radius, id, 0.5f, forces)
                const float3 rdir = normalize(rel);
                const float3 tang = velocity - dot(velocity, rdir) * rdir;
                const float3 spring = (-ph.coll_spring * (radius * 2 - sqrtf(d2))) * rdir;
                add = 0.5f * (spring + ph.coll_damping * velocity + ph.coll_shear * tang);
            }
            a.frc[pid] = make_float4(F.x + add.x, F.y + add.y, F.z + add.z, F4.w);
        }
        if (!splatOnly) {
            const float speed = length(velocity);
            const float3 dv = 1.0f * ((ph.velocity_collision_damping * speed) * h.refl - velocity);   // gpuCount = 1
            a.vel[pid] = make_float4(velocity.x + dv.x, velocity.y + dv.y, velocity.z + dv.z, v4.w);
        }
        const float3 ds = ph.vein_collision_force_intensity * velocity;
        const unsigned i0 = a.vidx[3 * h.tri], i1 = a.vidx[3 * h.tri + 1], i2 = a.vidx[3 * h.tri + 2];
        const float3 b = barycentric(pos + h.t * dir, xyz(a.vpos[i0]), xyz(a.vpos[i1]), xyz(a.vpos[i2]));
        // the reference uses plain += here and loses updates when two particles share a vertex
        splat_add(a, i0, b.x, ds);
        splat_add(a, i1, b.y, ds);
        splat_add(a, i2, b.z, ds);
        if (!splatOnly) atomicAdd(&a.counters->veinHits, 1ull);
    }
}

// one particle of the vein-collision stage (vein_collisions.cu:63-277)
template <bool FAST, bool STATS>
__device__ __forceinline__ void vein_collide_particle(const VeinCollideArgs& a, int pid, unsigned long long& myTests, bool splatOnly = false)
{
    const GridDev& g = a.tgrid;
    const PhysDev& ph = a.phys;
    const float4 p4 = a.pos[pid], v4 = a.vel[pid];
    const float3 pos = xyz(p4), velocity = xyz(v4);
    const float3 dir = normalize(velocity);
    const int cell = axis_cell(pos.z, g.minz, g.lenz, g.csz) * g.nx * g.ny + axis_cell(pos.y, g.miny, g.leny, g.csy) * g.nx +
                     axis_cell(pos.x, g.minx, g.lenx, g.csx);
    int x0, x1, y0, y1, z0, z1;
    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.x - g.minx, (float)g.csx)), g.nx, x0, x1);
    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.y - g.miny, (float)g.csy)), g.ny, y0, y1);
    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.z - g.minz, (float)g.csz)), g.nz, z0, z1);
    RayHit h;
    h.t = 1e10f; h.normal = f3(0.f, 0.f, 0.f); h.refl = f3(0.f, 0.f, 0.f); h.tri = 0;
    const bool hit = FAST ? first_hit_fast<STATS>(a, pos, dir, cell, x0, x1, y0, y1, z0, z1, h, myTests)
                          : first_hit_naive<STATS>(a, pos, dir, cell, x0, x1, y0, y1, z0, z1, h, myTests);
    if (a.dbgTri) {
        a.dbgTri[pid] = hit ? h.tri : -1;
        a.dbgT[pid] = hit ? h.t : 1e10f;
    }
    if (hit) vein_apply_hit(a, pid, p4, v4, dir, h, splatOnly);
}

// point (or ball) vs slab widened by `reach`
__device__ __forceinline__ bool slab_near(const CellSlab& sl, float3 p, float reach)
{
    const float d = sl.nx * p.x + sl.ny * p.y + sl.nz * p.z;
    return d + reach >= sl.dmin && d - reach <= sl.dmax;
}

// segment p .. p + reach*dir vs slab
__device__ __forceinline__ bool slab_segment(const CellSlab& sl, float3 p, float3 dir, float reach)
{
    const float d0 = sl.nx * p.x + sl.ny * p.y + sl.nz * p.z;
    const float d1 = d0 + reach * (sl.nx * dir.x + sl.ny * dir.y + sl.nz * dir.z);
    return fmaxf(d0, d1) + 1e-3f >= sl.dmin && fminf(d0, d1) - 1e-3f <= sl.dmax;
}

}  // namespace bcs

// Pair test and pair force of the particle-collision stage, shared by the per-slot stencil walk (collide.cu) and the
// symmetric pair search (pairs.cu).  detectCollision (simulation/particle_collisions.cuh:26-38) and
// physics::addResilientForceOnCollision (simulation/physics.cuh:133-145).
#pragma once
#include "bcs_internal.cuh"
#include "device_math.cuh"

namespace bcs {

__device__ __forceinline__ void stencil_range(int id, int count, int& lo, int& hi)
{
    // particle_collisions.cuh:126-268: `id < 1` / `id > count - 2` / else
    if (id < 1) { lo = 0; hi = 1; }
    else if (id > count - 2) { lo = -1; hi = 0; }
    else { lo = -1; hi = 1; }
}

// Collision forces are summed ORDER-INDEPENDENTLY: every pair contribution is rounded to a 64-bit fixed-point value
// (unit 2^-40: exact for every float of magnitude >= 2^-16, range +-8.3e6 - the scheme of the wall splats,
// vein_device.cuh) and the integer sum is folded into the float force once.  A particle's force therefore does not
// depend on the order in which its touching partners are found - by the per-slot stencil walk (collide.cu), by the
// symmetric pair search (pairs.cu), on one GPU or on eight.
constexpr float FX_SCALE = 1099511627776.0f;        // 2^40
constexpr float FX_UNSCALE = 1.0f / 1099511627776.0f;
__device__ __forceinline__ long long fx_of(float v) { return __float2ll_rn(v * FX_SCALE); }
__device__ __forceinline__ float fx_value(long long v) { return __ll2float_rn(v) * FX_UNSCALE; }

struct PairAccum {
    long long x, y, z;   // fixed-point sums of the pair contributions
    int hits;
};
__device__ __forceinline__ PairAccum pair_accum_zero() { return PairAccum{0ll, 0ll, 0ll, 0}; }

// detectCollision (particle_collisions.cuh:26-38): the distance test
__device__ __forceinline__ bool pair_touches(const float3 p1, const float r1, const float4 q4, const float r2)
{
    // The touch decision is a threshold on d2, so its rounding sequence is pinned (the compiler is otherwise free to
    // contract x*x + y*y + z*z in either association): the FMA chain nvcc emits for the reference's length_squared.
    // The oracle evaluates the same chain (std::fmaf), which makes hit sets bit-identical, not just the candidate sets.
    const float3 rel = p1 - xyz(q4);
    const float d2 = __fmaf_rn(rel.z, rel.z, __fmaf_rn(rel.y, rel.y, __fmul_rn(rel.x, rel.x)));
    const float minD = r1 + r2;
    return d2 <= __fmul_rn(minD, minD) && d2 >= 0.0001f;
}

// addResilientForceOnCollision with intensityCoefficient 0.5 (physics.cuh:133-145): what a touching partner at q4 with
// velocity v2 adds to the force on the particle at (p1, v1, r1)
__device__ __forceinline__ float3 pair_contribution(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4, const float3 v2)
{
    const float3 rel = p1 - xyz(q4);
    const float d2 = length_squared(rel);
    const float3 rv = v1 - v2;
    const float3 dir = normalize(rel);
    const float3 tang = rv - dot(rv, dir) * dir;
    const float3 spring = (-ph.coll_spring * (r1 * 2 - sqrtf(d2))) * dir;
    const float3 damp = ph.coll_damping * rv;
    const float3 shear = ph.coll_shear * tang;
    return 0.5f * (spring + damp + shear);
}

__device__ __forceinline__ void pair_force(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4,
                                           const float4* __restrict__ svel, int j, PairAccum& acc)
{
    const float3 c = pair_contribution(ph, p1, v1, r1, q4, xyz(svel[j]));
    acc.x += fx_of(c.x); acc.y += fx_of(c.y); acc.z += fx_of(c.z);
    ++acc.hits;
}

}  // namespace bcs

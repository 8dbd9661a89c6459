// Vein wall: vertex springs, vertex integration, triangle refit and particle vs. triangle collisions.
//
// Stands in for
//   calculateCentersKernel                  objects/vein_triangles.cu:14-27
//   VeinTriangles::gatherForcesFromNeighbors objects/vein_triangles.cu:126-163 (+ physics.cuh:38-41)
//   VeinTriangles::propagateForcesIntoPositions objects/vein_triangles.cu:88-117 (kernel + 3 memsets)
//   sim::detectVeinCollisions<UniformGrid>   simulation/vein_collisions.cu:63-277
//     -> calculateSideCollisions             simulation/vein_collisions.cuh:60-93
//     -> realCollisionDetection / calculateBaricentric  simulation/vein_collisions.cu:11-61
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

namespace bcs {

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tri_centers_kernel(const float4* __restrict__ vpos, const unsigned* __restrict__ vidx, int T,
                                                          float4* __restrict__ centers)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float4 a = vpos[vidx[3 * t]], b = vpos[vidx[3 * t + 1]], c = vpos[vidx[3 * t + 2]];
    centers[t] = make_float4(__fdiv_rn(a.x + b.x + c.x, 3.0f), __fdiv_rn(a.y + b.y + c.y, 3.0f), __fdiv_rn(a.z + b.z + c.z, 3.0f), 0.f);
}

void launch_tri_centers(const VeinArgs& a, float4* centers, cudaStream_t st)
{
    BCS_LAUNCH("tri_centers", st, tri_centers_kernel<<<(a.T + 255) / 256, 256, 0, st>>>(a.vpos, a.vidx, a.T, centers));
    BCS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vein_gather_kernel(const VeinArgs a)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.V) return;
    const float3 p = xyz(a.vpos[id]), v = xyz(a.vvel[id]);
    float3 F = f3(0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < BCS_VEIN_MAX_NEIGHBORS; ++s) {
        const int nb = __ldg(a.nbrIds + (size_t)s * a.V + id);
        if (nb != -1) {
            const float L = __ldg(a.nbrLen + (size_t)s * a.V + id);
            const float3 q = xyz(a.vpos[nb]);
            const float sf = (length(p - q) - L) * a.phys.vein_k_sniff + dot(normalize(p - q), (v - xyz(a.vvel[nb]))) * a.phys.vein_d_fact;
            F = F + sf * normalize(q - p);
        }
    }
    float4 f = a.vfrc[id];
    f.x += F.x; f.y += F.y; f.z += F.z;
    a.vfrc[id] = f;
}

void launch_vein_gather(const VeinArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("vein_gather", st, vein_gather_kernel<<<(a.V + 255) / 256, 256, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

// v += dt*F; x += dt*v; F = 0   (semi-implicit Euler + the three cudaMemsets of the reference, fused)
__global__ void __launch_bounds__(256) vein_integrate_kernel(const VeinArgs a)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.V) return;
    float4 x = a.vpos[id], v = a.vvel[id];
    const float4 F = a.vfrc[id];
    const float dt = a.phys.dt;
    v.x += dt * F.x; v.y += dt * F.y; v.z += dt * F.z;
    x.x += dt * v.x; x.y += dt * v.y; x.z += dt * v.z;
    a.vvel[id] = v;
    a.vpos[id] = x;
    a.vfrc[id] = make_float4(0.f, 0.f, 0.f, 0.f);
}

void launch_vein_integrate(const VeinArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("vein_integrate", st, vein_integrate_kernel<<<(a.V + 255) / 256, 256, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Triangles are re-packed every step in sorted-slot order (the wall moves): v0, e1 = v1-v0, e2 = v2-v0.
// The Moeller-Trumbore test only ever uses these three vectors, so the per-test gathers of the reference
// (3 index loads + 9 coordinate loads through two indirections) become three aligned float4 loads.
__global__ void __launch_bounds__(256) tri_refit_kernel(const float4* __restrict__ vpos, const unsigned* __restrict__ vidx,
                                                        const int* __restrict__ triIds, int T, TriPacked* __restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T) return;
    const int tri = triIds[s];
    const float3 v0 = xyz(vpos[vidx[3 * tri]]), v1 = xyz(vpos[vidx[3 * tri + 1]]), v2 = xyz(vpos[vidx[3 * tri + 2]]);
    const float3 e1 = v1 - v0, e2 = v2 - v0;
    TriPacked p;
    p.a = make_float4(v0.x, v0.y, v0.z, e1.x);
    p.b = make_float4(e1.y, e1.z, e2.x, e2.y);
    p.c = make_float4(e2.z, __int_as_float(tri), 0.f, 0.f);
    out[s] = p;
}

void launch_tri_refit(const VeinCollideArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("tri_refit", st, tri_refit_kernel<<<(a.T + 255) / 256, 256, 0, st>>>(a.vpos, a.vidx, a.triIds, a.T, a.tris));
    BCS_CUDA(cudaGetLastError());
}

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

// Straight traversal in the reference's order: x outer, y, z inner, sorted triangles inside a cell;
// the FIRST accepted triangle wins (vein_collisions.cuh:66-91).
template <bool STATS>
__global__ void __launch_bounds__(128) vein_collisions_kernel(const VeinCollideArgs a)
{
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long myTests = 0;
    if (pid < a.n) {
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
        bool hit = false;
        const int plane = g.nx * g.ny;
        for (int x = x0; x <= x1 && !hit; ++x)
            for (int y = y0; y <= y1 && !hit; ++y)
                for (int z = z0; z <= z1 && !hit; ++z) {
                    const int c = cell + z * plane + y * g.nx + x;
                    if (c < 0 || c >= g.cells) continue;
                    const int s = a.cellStart[c], e = a.cellEnd[c];
                    for (int i = s; i <= e; ++i) {
                        TriPacked tp;
                        tp.a = a.tris[i].a; tp.b = a.tris[i].b; tp.c = a.tris[i].c;
                        if (STATS) ++myTests;
                        if (ray_triangle(pos, dir, tp, h)) { hit = true; break; }
                    }
                }
        if (a.dbgTri) {
            a.dbgTri[pid] = hit ? h.tri : -1;
            a.dbgT[pid] = h.t;
        }
        // relativePosition = pos - (pos + t*dir), evaluated literally (vein_collisions.cu:234)
        const float3 rel = pos - (pos + h.t * dir);
        const float d2 = length_squared(rel);
        if (a.apply && hit && d2 <= ph.impact2) {
            if (d2 > ph.minForce2) {
                const float4 F4 = a.frc[pid];
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
            const float speed = length(velocity);
            const float3 dv = 1.0f * ((ph.velocity_collision_damping * speed) * h.refl - velocity);   // gpuCount = 1
            a.vel[pid] = make_float4(velocity.x + dv.x, velocity.y + dv.y, velocity.z + dv.z, v4.w);
            const float3 ds = ph.vein_collision_force_intensity * velocity;
            const unsigned i0 = a.vidx[3 * h.tri], i1 = a.vidx[3 * h.tri + 1], i2 = a.vidx[3 * h.tri + 2];
            const float3 b = barycentric(pos + h.t * dir, xyz(a.vpos[i0]), xyz(a.vpos[i1]), xyz(a.vpos[i2]));
            // the reference uses plain += here and loses updates when two particles share a vertex
            atomicAdd(&a.vfrc[i0].x, b.x * ds.x); atomicAdd(&a.vfrc[i0].y, b.x * ds.y); atomicAdd(&a.vfrc[i0].z, b.x * ds.z);
            atomicAdd(&a.vfrc[i1].x, b.y * ds.x); atomicAdd(&a.vfrc[i1].y, b.y * ds.y); atomicAdd(&a.vfrc[i1].z, b.y * ds.z);
            atomicAdd(&a.vfrc[i2].x, b.z * ds.x); atomicAdd(&a.vfrc[i2].y, b.z * ds.y); atomicAdd(&a.vfrc[i2].z, b.z * ds.z);
            atomicAdd(&a.counters->veinHits, 1ull);
        }
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if ((threadIdx.x & 31) == 0 && a.apply) atomicAdd(&a.counters->triTests, myTests);
    }
}

void launch_vein_collisions(const VeinCollideArgs& a, cudaStream_t st)
{
    const int threads = 128, blocks = (a.n + threads - 1) / threads;
    if (a.stats) BCS_LAUNCH("vein_collisions", st, vein_collisions_kernel<true><<<blocks, threads, 0, st>>>(a));
    else BCS_LAUNCH("vein_collisions", st, vein_collisions_kernel<false><<<blocks, threads, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

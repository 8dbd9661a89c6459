// Particle-particle collision / repulsion pass over the 27-cell neighbourhood.
//
// Stands in for sim::calculateParticleCollisions<UniformGrid> (simulation/particle_collisions.cuh:104-269)
// -> detectCollisionsInNeighborCells (:53-83) -> detectCollision (:26-38)
// -> physics::addResilientForceOnCollision (simulation/physics.cuh:133-145).
//
// B200 mapping: one thread per SORTED slot (as the reference), but all candidate data comes from the
// sorted-order float4 copies written by the grid build, so a neighbour cell is one contiguous, 16-byte
// aligned run instead of a chain particleIds[i] -> positions.{x,y,z}[id] of four dependent gathers.
// Because the cell id is x-fastest (uniform_grid.cu:33-35), the three x-neighbours of a (y,z) row are
// adjacent cell ids and - in clean semantics - their particles form ONE contiguous slot range, so the
// stencil is walked as 9 rows instead of 27 cells.  Neighbouring threads of a warp sit in the same or
// adjacent cells and therefore read the same runs (L1 broadcast).  The force is accumulated in registers
// and written once (read-modify-write by particle id) only when a collision happened.
//
// The candidate SET of every particle is identical to the reference's; the visiting order differs
// (rows: z,y outer, x inner vs the reference's x outer), which only permutes a float sum.
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

namespace bcs {

__device__ __forceinline__ void stencil_range(int id, int count, int& lo, int& hi)
{
    // particle_collisions.cuh:126-268: `id < 1` / `id > count - 2` / else
    if (id < 1) { lo = 0; hi = 1; }
    else if (id > count - 2) { lo = -1; hi = 0; }
    else { lo = -1; hi = 1; }
}

struct PairAccum {
    float3 F;
    int hits;
};

// detectCollision + addResilientForceOnCollision with intensityCoefficient 0.5
__device__ __forceinline__ void test_pair(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4,
                                          const float r2, const float4* __restrict__ svel, int j, PairAccum& acc)
{
    const float3 rel = p1 - xyz(q4);
    const float d2 = length_squared(rel);
    const float minD = r1 + r2;
    if (d2 <= minD * minD && d2 >= 0.0001f) {
        const float3 rv = v1 - xyz(svel[j]);
        const float3 dir = normalize(rel);
        const float3 tang = rv - dot(rv, dir) * dir;
        const float3 spring = (-ph.coll_spring * (r1 * 2 - sqrtf(d2))) * dir;
        const float3 damp = ph.coll_damping * rv;
        const float3 shear = ph.coll_shear * tang;
        acc.F = acc.F + 0.5f * (spring + damp + shear);
        ++acc.hits;
    }
}

template <bool REFERENCE, bool DEBUG, bool STATS>
__global__ void __launch_bounds__(128) particle_collisions_kernel(const CollideArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long myTests = 0;
    int myHits = 0;
    // slab mode: the active count lives on the device and ghost slots (bit 31 of the id) are candidates only
    const int nActive = a.nDev ? *a.nDev : a.n;
    if (slot < nActive && __float_as_int(a.svel[slot].w) >= 0) {
        const GridDev& g = a.grid;
        const float4 p4 = a.spos[slot];
        const float4 v4 = a.svel[slot];
        const int tag = __float_as_int(v4.w);
        const int pid = tag & 0x7fffffff;
        const float3 p1 = xyz(p4), v1 = xyz(v4);
        const int cell = a.keys[slot];
        int x0, x1, y0, y1, z0, z1;
        stencil_range(axis_cell_raw(p1.x, g.minx, g.csx), g.nx, x0, x1);
        stencil_range(axis_cell_raw(p1.y, g.miny, g.csy), g.ny, y0, y1);
        stencil_range(axis_cell_raw(p1.z, g.minz, g.csz), g.nz, z0, z1);

        // radius lookup.  Reference semantics: the type whose launch slice contains this SLOT supplies
        // (modelStart, particlesStart, P) for BOTH particles (particle_collisions.cuh:76,124).
        int sM = 0, sP0 = 0, sPP = 1;
        float r1;
        if (REFERENCE) {
            int t = 0;
            while (t + 1 < a.types.n && slot >= a.types.t[t + 1].pStart) ++t;
            sM = a.types.t[t].mStart; sP0 = a.types.t[t].pStart; sPP = a.types.t[t].P;
            r1 = __ldg(a.collR + max(0, sM + (pid - sP0) % sPP));
        } else {
            r1 = p4.w;
        }

        PairAccum acc{f3(0.f, 0.f, 0.f), 0};
        int cnt = 0;
        unsigned long long sum = 0;
        const int plane = g.nx * g.ny;
        if (REFERENCE) {
            for (int z = z0; z <= z1; ++z) {
                for (int y = y0; y <= y1; ++y) {
                    const int row = cell + z * plane + y * g.nx;
                    // tables may hold stale ranges: every cell is scanned on its own, exactly as stored
                    for (int x = x0; x <= x1; ++x) {
                        const int c = row + x;
                        if (c < 0 || c >= g.cells) continue;
                        const int s = a.cellStart[c], e = a.cellEnd[c];
                        for (int j = s; j <= e; ++j) {
                            const float4 q4 = a.spos[j];
                            const int qid = __float_as_int(q4.w);
                            if (qid == pid) continue;
                            if (DEBUG) { ++cnt; sum += (unsigned long long)(qid + 1) * 0x9E3779B97F4A7C15ull; }
                            const float r2 = __ldg(a.collR + max(0, sM + (qid - sP0) % sPP));
                            test_pair(a.phys, p1, v1, r1, q4, r2, a.svel, j, acc);
                            if (STATS) ++myTests;
                        }
                    }
                }
            }
        } else {
            // pass 1 (branch-free, all lanes alike): occupancy bits of the 9 stencil rows.  A row is the <= 3
            // x-adjacent cells [row+x0, row+x1]: adjacent cell ids, i.e. adjacent bits of the occupancy mask.
            const int nb = x1 - x0 + 1;
            const unsigned nbm = (1u << nb) - 1u;
            unsigned occ = 0;   // 3 bits per row, row r = (dz+1)*3 + (dy+1)
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                const int dz = r / 3 - 1, dy = r % 3 - 1;
                const int c0 = cell + dz * plane + dy * g.nx + x0;
                const bool in = dz >= z0 && dz <= z1 && dy >= y0 && dy <= y1 && c0 >= 0 && c0 + nb <= g.cells;
                unsigned bits = 0;
                if (in) {
                    const int w0 = c0 >> 5;
                    bits = __funnelshift_r(__ldg(a.cellMask + w0), __ldg(a.cellMask + w0 + 1), c0 & 31) & nbm;
                }
                occ |= bits << (3 * r);
            }
            // pass 2: only the non-empty rows.  Occupied cells of a row have consecutive ranks and ONE contiguous
            // slot range [occStart[rank(first)], occStart[rank(first) + popc(bits)]).
            while (occ) {
                const int r = (__ffs(occ) - 1) / 3;
                const unsigned bits = (occ >> (3 * r)) & 7u;
                occ &= ~(7u << (3 * r));
                const int first = cell + (r / 3 - 1) * plane + (r % 3 - 1) * g.nx + x0 + __ffs(bits) - 1;
                const int fw = first >> 5;
                const int rank = __ldg(a.cellRank + fw) + __popc(__ldg(a.cellMask + fw) & ((1u << (first & 31)) - 1u));
                const int lo = __ldg(a.occStart + rank), hi = __ldg(a.occStart + rank + __popc(bits)) - 1;
                for (int j = lo; j <= hi; ++j) {
                    if (j == slot) continue;
                    const float4 q4 = a.spos[j];
                    if (DEBUG) {
                        const int qid = __float_as_int(a.svel[j].w) & 0x7fffffff;
                        ++cnt; sum += (unsigned long long)(qid + 1) * 0x9E3779B97F4A7C15ull;
                    }
                    test_pair(a.phys, p1, v1, r1, q4, q4.w, a.svel, j, acc);
                    if (STATS) ++myTests;
                }
            }
        }
        if (DEBUG) {
            a.dbgCount[pid] = cnt;
            a.dbgSum[pid] = sum;
            a.dbgHits[pid] = acc.hits;
        } else if (acc.hits) {
            float4 f = a.frc[pid];
            f.x += acc.F.x; f.y += acc.F.y; f.z += acc.F.z;
            a.frc[pid] = f;
        }
        myHits = acc.hits;
    }
    if (STATS && !DEBUG) {
        // warp-aggregated counters
        for (int o = 16; o; o >>= 1) {
            myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
            myHits += __shfl_xor_sync(0xffffffffu, myHits, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&a.counters->pairTests, myTests);
            atomicAdd(&a.counters->pairHits, (unsigned long long)myHits);
        }
    }
}

void launch_particle_collisions(const CollideArgs& a, cudaStream_t st)
{
    const int threads = 128, blocks = (a.n + threads - 1) / threads;
    const bool dbg = a.dbgCount != nullptr;
    if (a.reference) {
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, true, false><<<blocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, false, false><<<blocks, threads, 0, st>>>(a));
    } else {
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, true, false><<<blocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, false><<<blocks, threads, 0, st>>>(a));
    }
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

// Particle-particle collision stage over the ROW DIRECTORY (RowsGrid, kernels.cu): symmetric pair search.
//
// Stands in for sim::calculateParticleCollisions<UniformGrid> (simulation/particle_collisions.cuh:104-269)
// -> detectCollisionsInNeighborCells (:53-83) -> detectCollision (:26-38)
// -> physics::addResilientForceOnCollision (simulation/physics.cuh:133-145), like collide.cu.
//
// Why another mapping.  At the bench's density a particle has ~4 candidates in its 27 cells, but the per-slot walk pays
// nine index resolutions (occupancy bits -> rank -> slot range) to find them: 37 M warp instructions, 21 of 32 lanes
// active, 0.17 of the HBM roofline (round 1).  Here:
//   * the candidate relation is symmetric (j is in i's stencil <=> i is in j's, for particles inside the grid), so a slot
//     only looks FORWARD in the sorted order - the rest of its own row, the row above it, and the three rows of the next
//     z layer - and every touching pair is recorded once, for both ends;
//   * cell ids are x-fastest, so "own row forward + next row" is the run of slots that starts right after the slot
//     itself (no lookup at all), and the three rows of the next layer are ONE contiguous run found with one load from
//     the row directory.  Inside a run the stencil test is integer work on the sorted keys;
//   * candidates are only queued during the scan and tested together afterwards (all lanes test their k-th candidate in
//     the same iteration), touching pairs go to per-slot lists;
//   * pair_apply walks each slot's list in ascending partner order - the encounter order of the per-slot walk - so the
//     sums, and therefore the forces, are bit-identical to collide.cu's.
// A build in which some particle sits outside the grid (one-sided stencils: the relation is no longer symmetric), or
// whose pair lists overflow, is handled by pair_walk: every slot scans its whole stencil through the row directory.
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "pair_device.cuh"
#include "rows_device.cuh"

#include <algorithm>

namespace bcs {

namespace {

constexpr int PS_THREADS = 128;
constexpr int PS_QCAP = 16;   // queued candidates per slot before they are tested (8 KB of shared memory per CTA)
constexpr unsigned long long DBG_MIX = 0x9E3779B97F4A7C15ull;

// Half-hit entries are owned by the FORWARD end of a pair: slot i keeps entries 2 * (i * PS_FWD + h) and + 1 for its h-th
// touching forward partner, so nothing is allocated at run time and the only atomics are the exchanges on the two list
// heads (distinct addresses).  A slot with more than PS_FWD touching forward partners takes entries from a small shared
// pool; only when that runs dry is the overflow flag raised and the stage redone by pair_walk (dense clusters; the
// row-directory mode is meant for sparse scenes).
constexpr int PS_FWD = 4;

__device__ __forceinline__ void record_pair(const PairLists& L, int i, int j, int h)
{
    int e = 2 * (i * PS_FWD + h);
    if (h >= PS_FWD) {
        // beyond the slot's own entries: the shared pool behind them (rare, so its counter is not contended)
        const int k = atomicAdd(&L.ctl[2], 2);
        if (k + 2 > L.pool) {
            L.ctl[0] = 1;
            return;
        }
        e = L.poolStart + k;
    }
    L.entries[e] = make_int2(j, atomicExch(&L.head[i], e));
    L.entries[e + 1] = make_int2(i, atomicExch(&L.head[j], e + 1));
}

template <bool DEBUG, bool SLAB>
__device__ __forceinline__ void test_queued(const CollideArgs& a, int (*q)[PS_THREADS], int cnt, int slot, int tag, const float3 p1, const float r1,
                                            unsigned long long& tests, int& hits)
{
    const int tid = threadIdx.x;
    for (int t = 0; t < cnt; ++t) {
        const int j = q[t][tid];
        const float4 q4 = a.spos[j];
        int tagJ = 0;
        if (SLAB || DEBUG) tagJ = a.ids[j];
        if (SLAB && tag < 0 && tagJ < 0) continue;   // two ghosts: nobody here owns either end
        const bool touch = pair_touches(p1, r1, q4, q4.w);
        if (DEBUG) {
            const int pi = tag & 0x7fffffff, pj = tagJ & 0x7fffffff;
            if (tag >= 0) {
                atomicAdd(&a.dbgCount[pi], 1);
                atomicAdd(&a.dbgSum[pi], (unsigned long long)(pj + 1) * DBG_MIX);
                if (touch) atomicAdd(&a.dbgHits[pi], 1);
            }
            if (tagJ >= 0) {
                atomicAdd(&a.dbgCount[pj], 1);
                atomicAdd(&a.dbgSum[pj], (unsigned long long)(pi + 1) * DBG_MIX);
                if (touch) atomicAdd(&a.dbgHits[pj], 1);
            }
        } else if (touch) {
            record_pair(a.pairs, slot, j, hits);
            ++hits;
        }
        ++tests;
    }
}

template <bool DEBUG, bool SLAB, bool STATS>
__global__ void __launch_bounds__(PS_THREADS) pair_search_kernel(const CollideArgs a)
{
    __shared__ int q[PS_QCAP][PS_THREADS];
    if (*a.irregular) return;   // pair_walk takes the build
    const GridDev& g = a.grid;
    const int n = a.nDev ? *a.nDev : a.n;
    const int tid = threadIdx.x;
    const int plane = g.nx * g.ny;
    unsigned long long tests = 0;
    int hitsAll = 0;
    for (int slot = blockIdx.x * blockDim.x + tid; slot < n; slot += gridDim.x * blockDim.x) {
        int hits = 0;   // touching forward partners of this slot so far
        const float4 p4 = a.spos[slot];
        const int c = a.keys[slot];
        int tag = 0;
        if (SLAB || DEBUG) tag = a.ids[slot];
        const float3 p1 = xyz(p4);
        const float r1 = p4.w;
        int x0, x1, y0, y1, z0, z1;
        stencil_range(axis_cell_raw(p1.x, g.minx, g.csx), g.nx, x0, x1);
        stencil_range(axis_cell_raw(p1.y, g.miny, g.csy), g.ny, y0, y1);
        stencil_range(axis_cell_raw(p1.z, g.minz, g.csz), g.nz, z0, z1);
        const unsigned nb = (unsigned)(x1 - x0 + 1);
        int cnt = 0;

        // run A: the slots right after this one - the rest of the own row (cells c .. c + x1) and the row above
        {
            const int wRow = c + g.nx + x0;                               // window start in row y + 1
            const int hi = y1 > 0 ? c + g.nx + x1 : c + x1;
            int j = slot + 1;
            int k = a.keys[j];                                            // keys[n .. n + 3] are sentinels
            while (k <= hi) {
                const int kn = a.keys[j + 1];
                const bool in = (unsigned)(k - c) <= (unsigned)x1 || (y1 > 0 && (unsigned)(k - wRow) < nb);
                if (in) {
                    q[cnt][tid] = j;
                    if (++cnt == PS_QCAP) {
                        test_queued<DEBUG, SLAB>(a, q, cnt, slot, tag, p1, r1, tests, hits);
                        cnt = 0;
                    }
                }
                ++j;
                k = kn;
            }
        }
        // run B: rows y + y0 .. y + y1 of the next z layer, one contiguous run of slots
        if (z1 > 0) {
            const int w0 = c + plane + y0 * g.nx + x0;                    // window starts of the (up to) three rows
            const int nr = y1 - y0 + 1;
            const int w1 = nr > 1 ? w0 + g.nx : (int)0x80000000;          // absent row: (unsigned)(k - w) is never below nb
            const int w2 = nr > 2 ? w0 + 2 * g.nx : (int)0x80000000;
            const int hi = c + plane + y1 * g.nx + x1;
            int j = a.rowStart[rows_div((unsigned)w0, a.nxMagic, a.nxShift)];
            int k = a.keys[j];
            while (k <= hi) {
                const int kn = a.keys[j + 1];
                const bool in = (unsigned)(k - w0) < nb || (unsigned)(k - w1) < nb || (unsigned)(k - w2) < nb;
                if (in) {
                    q[cnt][tid] = j;
                    if (++cnt == PS_QCAP) {
                        test_queued<DEBUG, SLAB>(a, q, cnt, slot, tag, p1, r1, tests, hits);
                        cnt = 0;
                    }
                }
                ++j;
                k = kn;
            }
        }
        test_queued<DEBUG, SLAB>(a, q, cnt, slot, tag, p1, r1, tests, hits);
        hitsAll += hits;
    }
    if (STATS && !DEBUG) {
        for (int o = 16; o; o >>= 1) {
            tests += __shfl_xor_sync(0xffffffffu, tests, o);
            hitsAll += __shfl_xor_sync(0xffffffffu, hitsAll, o);
        }
        if ((threadIdx.x & 31) == 0) {
            // both ends of a pair test / feel each other (the per-slot walk counts a pair from either side)
            atomicAdd(&a.counters->pairTests, 2ull * tests);
            atomicAdd(&a.counters->pairHits, 2ull * (unsigned long long)hitsAll);
        }
    }
}

// force of one touching pair on the particle at (p1, v1, r1): addResilientForceOnCollision, same expression as
// pair_force (pair_device.cuh) with the partner's velocity gathered by particle id
__device__ __forceinline__ void pair_force_v(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4, const float3 v2,
                                             PairAccum& acc)
{
    const float3 rel = p1 - xyz(q4);
    const float d2 = length_squared(rel);
    const float3 rv = v1 - v2;
    const float3 dir = normalize(rel);
    const float3 tang = rv - dot(rv, dir) * dir;
    const float3 spring = (-ph.coll_spring * (r1 * 2 - sqrtf(d2))) * dir;
    const float3 damp = ph.coll_damping * rv;
    const float3 shear = ph.coll_shear * tang;
    acc.F = acc.F + 0.5f * (spring + damp + shear);
    ++acc.hits;
}

// one thread per sorted slot: a slot with half-hits takes its partners in ascending slot order, F[pid] += sum, and clears
// its list head
__global__ void __launch_bounds__(256) pair_apply_kernel(const CollideArgs a)
{
    const PairLists& L = a.pairs;
    const int n = a.nDev ? *a.nDev : a.n;
    const bool skip = L.ctl[0] != 0 || *a.irregular != 0;   // pair_walk did the stage: only clear the lists
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int first = L.head[slot];
        if (first < 0) continue;
        L.head[slot] = -1;
        const int tag = a.ids[slot];
        if (skip || tag < 0) continue;   // ghosts are partners only (forces go to owned particles, particle_collisions.cuh:36)
        const int pid = tag;
        const float4 p4 = a.spos[slot];
        const float3 p1 = xyz(p4), v1 = xyz(a.vel[pid]);
        PairAccum acc{f3(0.f, 0.f, 0.f), 0};
        int last = -1;
        while (true) {
            // next partner in ascending slot order (lists hold a handful of entries)
            int best = 0x7fffffff;
            for (int e = first; e >= 0;) {
                const int2 en = L.entries[e];
                if (en.x > last && en.x < best) best = en.x;
                e = en.y;
            }
            if (best == 0x7fffffff) break;
            const float3 v2 = xyz(a.vel[a.ids[best] & 0x7fffffff]);
            pair_force_v(a.phys, p1, v1, p4.w, a.spos[best], v2, acc);
            last = best;
        }
        float4 f = a.frc[pid];
        f.x += acc.F.x; f.y += acc.F.y; f.z += acc.F.z;
        a.frc[pid] = f;
    }
    // the last CTA to leave lowers the overflow flag for the next search
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&L.ctl[1], 1) == (int)gridDim.x - 1) {
            L.ctl[0] = 0; L.ctl[1] = 0; L.ctl[2] = 0;
            __threadfence();
        }
    }
}

// Every slot scans its WHOLE stencil through the row directory (the per-slot walk of collide.cu without the compact cell
// index): the fallback for builds with particles outside the grid or overflowing pair lists, the A/B partner of the
// symmetric search (BCS_COLLIDE=walk), and - windows being plain linear cell ranges [c0, c0 + nb) exactly as
// particle_collisions.cuh:53-83 forms them - correct for any key, clamped ones included.
template <bool DEBUG, bool ALWAYS, bool STATS>
__global__ void __launch_bounds__(PS_THREADS) pair_walk_kernel(const CollideArgs a)
{
    if (!ALWAYS && !(*a.irregular != 0 || a.pairs.ctl[0] != 0)) return;
    const GridDev& g = a.grid;
    const int n = a.nDev ? *a.nDev : a.n;
    const int plane = g.nx * g.ny;
    unsigned long long tests = 0;
    int hitsAll = 0;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
        const int tag = a.ids[slot];
        if (tag < 0) continue;
        const int pid = tag;
        const float4 p4 = a.spos[slot];
        const int c = a.keys[slot];
        const float3 p1 = xyz(p4), v1 = xyz(a.vel[pid]);
        const float r1 = p4.w;
        int x0, x1, y0, y1, z0, z1;
        stencil_range(axis_cell_raw(p1.x, g.minx, g.csx), g.nx, x0, x1);
        stencil_range(axis_cell_raw(p1.y, g.miny, g.csy), g.ny, y0, y1);
        stencil_range(axis_cell_raw(p1.z, g.minz, g.csz), g.nz, z0, z1);
        const int nbi = x1 - x0 + 1;
        const unsigned nb = (unsigned)nbi;
        PairAccum acc{f3(0.f, 0.f, 0.f), 0};
        int cnt = 0;
        unsigned long long sum = 0;
        for (int dz = z0; dz <= z1; ++dz) {
            int w[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
            int lo = 0x7fffffff, hi = -1;
            for (int dy = y0; dy <= y1; ++dy) {
                const int c0 = c + dz * plane + dy * g.nx + x0;
                if (c0 < 0 || c0 + nbi > g.cells) continue;   // the row lies beyond the grid
                w[dy - y0] = c0;
                lo = min(lo, c0);
                hi = max(hi, c0 + nbi - 1);
            }
            if (hi < 0) continue;
            int j = a.rowStart[rows_div((unsigned)lo, a.nxMagic, a.nxShift)];
            for (int k = a.keys[j]; k <= hi; k = a.keys[++j]) {
                if (j >= n) break;
                const bool in = (unsigned)(k - w[0]) < nb || (unsigned)(k - w[1]) < nb || (unsigned)(k - w[2]) < nb;
                if (!in || j == slot) continue;
                const float4 q4 = a.spos[j];
                if (DEBUG) {
                    const int qid = a.ids[j] & 0x7fffffff;
                    ++cnt; sum += (unsigned long long)(qid + 1) * DBG_MIX;
                }
                if (pair_touches(p1, r1, q4, q4.w)) pair_force_v(a.phys, p1, v1, r1, q4, xyz(a.vel[a.ids[j] & 0x7fffffff]), acc);
                if (STATS) ++tests;
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
        hitsAll += acc.hits;
    }
    if (STATS && !DEBUG) {
        for (int o = 16; o; o >>= 1) {
            tests += __shfl_xor_sync(0xffffffffu, tests, o);
            hitsAll += __shfl_xor_sync(0xffffffffu, hitsAll, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&a.counters->pairTests, tests);
            atomicAdd(&a.counters->pairHits, (unsigned long long)hitsAll);
        }
    }
}

}  // namespace

void launch_particle_collisions_rows(const CollideArgs& a, cudaStream_t st)
{
    const int threads = PS_THREADS, allBlocks = (a.n + threads - 1) / threads;
    const int blocks = a.nDev ? std::min(allBlocks, 2 * BOUNDED_BLOCKS) : allBlocks;
    const int walkBlocks = std::min(allBlocks, 2 * BOUNDED_BLOCKS);
    const bool dbg = a.dbgCount != nullptr, slab = a.nDev != nullptr;
    if (a.fullWalk) {
        if (dbg) BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<true, true, false><<<walkBlocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<false, true, true><<<walkBlocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<false, true, false><<<walkBlocks, threads, 0, st>>>(a));
        BCS_CUDA(cudaGetLastError());
        return;
    }
    if (dbg) {
        // candidate counts / checksums / touching pairs per particle (test instrumentation): both ends of a pair are
        // credited with integer atomics by the same search the production path runs
        BCS_CUDA(cudaMemsetAsync(a.dbgCount, 0, (size_t)a.n * sizeof(int), st));
        BCS_CUDA(cudaMemsetAsync(a.dbgSum, 0, (size_t)a.n * sizeof(unsigned long long), st));
        BCS_CUDA(cudaMemsetAsync(a.dbgHits, 0, (size_t)a.n * sizeof(int), st));
        if (slab) BCS_LAUNCH("particle_collisions", st, pair_search_kernel<true, true, false><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, pair_search_kernel<true, false, false><<<blocks, threads, 0, st>>>(a));
        BCS_LAUNCH("pair_walk", st, pair_walk_kernel<true, false, false><<<walkBlocks, threads, 0, st>>>(a));
        BCS_CUDA(cudaGetLastError());
        return;
    }
    if (slab) {
        if (a.stats) BCS_LAUNCH("particle_collisions", st, pair_search_kernel<false, true, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, pair_search_kernel<false, true, false><<<blocks, threads, 0, st>>>(a));
    } else {
        if (a.stats) BCS_LAUNCH("particle_collisions", st, pair_search_kernel<false, false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, pair_search_kernel<false, false, false><<<blocks, threads, 0, st>>>(a));
    }
    if (a.stats) BCS_LAUNCH("pair_walk", st, pair_walk_kernel<false, false, true><<<walkBlocks, threads, 0, st>>>(a));
    else BCS_LAUNCH("pair_walk", st, pair_walk_kernel<false, false, false><<<walkBlocks, threads, 0, st>>>(a));
    BCS_LAUNCH("pair_apply", st, pair_apply_kernel<<<std::min((a.n + 255) / 256, 2 * BOUNDED_BLOCKS), 256, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

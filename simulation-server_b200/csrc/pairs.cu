// Particle-particle collision stage over the ROW DIRECTORY (RowsGrid, kernels.cuh): symmetric pair search.
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
//     z layer - and every touching pair is found once, for both ends;
//   * cell ids are x-fastest, so "own row forward + next row" is the run of slots that starts right after the slot
//     itself (no lookup at all), and the three rows of the next layer are ONE contiguous run found with one load from
//     the row directory.  Inside a run the stencil test is integer work on the sorted keys;
//   * candidates are only queued during the scan and tested together afterwards (all lanes test their k-th candidate in
//     the same iteration); touching pairs are collected per CTA in shared memory and appended to a global pair list
//     with ONE atomic per CTA;
//   * pair_force takes a thread per touching pair (dense: no lane waits for a neighbour's hit), evaluates both ends and
//     adds the two contributions to per-particle FIXED-POINT accumulators (pair_device.cuh): integer sums do not depend
//     on the order of arrival, so the force is bit-identical to the per-slot walk's, run after run;
//   * pair_fold streams the accumulators in particle order and adds the non-zero sums to the float forces.
// A build in which some particle sits outside the grid (one-sided stencils: the relation is no longer symmetric), or
// whose pair list overflows, is handled by pair_walk: every slot scans its whole stencil through the row directory.
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "pair_device.cuh"
#include "rows_device.cuh"

#include <algorithm>
#include <type_traits>

namespace bcs {

namespace {

constexpr int PS_THREADS = 128;
constexpr int PS_QCAP = 16;      // queued candidates per slot before they are tested (8 KB of shared memory per CTA)
constexpr int PS_PAIRS = 384;    // touching pairs a CTA collects before it spills to the global list directly
constexpr unsigned long long DBG_MIX = 0x9E3779B97F4A7C15ull;

// ctl words of PairLists
enum { CTL_COUNT = 0, CTL_OVERFLOW = 1, CTL_DONE = 2 };

__device__ __forceinline__ void emit_pair(const PairLists& L, int2* sPairs, int* sCount, int i, int j)
{
    const int k = atomicAdd(sCount, 1);
    if (k < PS_PAIRS) {
        sPairs[k] = make_int2(i, j);
    } else {
        // the CTA's buffer is full (a dense cluster): straight to the global list
        const int g = atomicAdd(&L.ctl[CTL_COUNT], 1);
        if (g < L.cap) L.pairs[g] = make_int2(i, j);
        else L.ctl[CTL_OVERFLOW] = 1;   // pair_walk redoes the stage
    }
}

template <bool DEBUG, bool SLAB>
__device__ __forceinline__ void test_queued(const CollideArgs& a, int (*q)[PS_THREADS], int cnt, int slot, int tag, const float3 p1, const float r1,
                                            unsigned long long& tests, int& hits, int2* sPairs, int* sCount)
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
            emit_pair(a.pairs, sPairs, sCount, slot, j);
            ++hits;
        }
        ++tests;
    }
}

#define BCS_QUEUE_PUSH(J)                                                                               \
    {                                                                                                   \
        q[cnt][tid] = (J);                                                                              \
        if (++cnt == PS_QCAP) {                                                                         \
            test_queued<DEBUG, SLAB>(a, q, cnt, slot, tag, p1, r1, tests, hits, sPairs, &sCount);       \
            cnt = 0;                                                                                    \
        }                                                                                               \
    }

template <bool DEBUG, bool SLAB, bool STATS>
__global__ void __launch_bounds__(PS_THREADS) pair_search_kernel(const CollideArgs a)
{
    __shared__ int q[PS_QCAP][PS_THREADS];
    __shared__ int2 sPairs[PS_PAIRS];
    __shared__ int sCount, sBase;
    if (*a.irregular) return;   // pair_walk takes the build
    const GridDev& g = a.grid;
    const int n = a.nDev ? *a.nDev : a.n;
    const int tid = threadIdx.x;
    const int plane = g.nx * g.ny;
    unsigned long long tests = 0;
    int hits = 0;
    if (tid == 0) sCount = 0;
    __syncthreads();
    for (int slot = blockIdx.x * blockDim.x + tid; slot < n; slot += gridDim.x * blockDim.x) {
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
                if (in) BCS_QUEUE_PUSH(j)
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
            int j = a.rowStart[rows_of_key(a.rowsGrid, w0)];
            int k = a.keys[j];
            while (k <= hi) {
                const int kn = a.keys[j + 1];
                const bool in = (unsigned)(k - w0) < nb || (unsigned)(k - w1) < nb || (unsigned)(k - w2) < nb;
                if (in) BCS_QUEUE_PUSH(j)
                ++j;
                k = kn;
            }
        }
        test_queued<DEBUG, SLAB>(a, q, cnt, slot, tag, p1, r1, tests, hits, sPairs, &sCount);
    }
    if (!DEBUG) {
        // the CTA's touching pairs -> global list: one atomic per CTA
        __syncthreads();
        const int mine = min(sCount, PS_PAIRS);
        if (tid == 0 && mine) sBase = atomicAdd(&a.pairs.ctl[CTL_COUNT], mine);
        __syncthreads();
        if (mine) {
            const int base = sBase;
            if (base + mine <= a.pairs.cap) {
                for (int k = tid; k < mine; k += PS_THREADS) a.pairs.pairs[base + k] = sPairs[k];
            } else if (tid == 0) {
                a.pairs.ctl[CTL_OVERFLOW] = 1;
            }
        }
    }
    if (STATS && !DEBUG) {
        for (int o = 16; o; o >>= 1) {
            tests += __shfl_xor_sync(0xffffffffu, tests, o);
            hits += __shfl_xor_sync(0xffffffffu, hits, o);
        }
        if ((threadIdx.x & 31) == 0) {
            // both ends of a pair test / feel each other (the per-slot walk counts a pair from either side)
            atomicAdd(&a.counters->pairTests, 2ull * tests);
            atomicAdd(&a.counters->pairHits, 2ull * (unsigned long long)hits);
        }
    }
}

template <bool DEBUG, bool STATS>
__device__ __forceinline__ void walk_slots(const CollideArgs& a)
{
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
        PairAccum acc = pair_accum_zero();
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
            int j = a.rowStart[rows_of_key(a.rowsGrid, lo)];
            for (int k = a.keys[j]; k <= hi; k = a.keys[++j]) {
                if (j >= n) break;
                const bool in = (unsigned)(k - w[0]) < nb || (unsigned)(k - w[1]) < nb || (unsigned)(k - w[2]) < nb;
                if (!in || j == slot) continue;
                const float4 q4 = a.spos[j];
                if (DEBUG) {
                    const int qid = a.ids[j] & 0x7fffffff;
                    ++cnt; sum += (unsigned long long)(qid + 1) * DBG_MIX;
                }
                if (pair_touches(p1, r1, q4, q4.w)) {
                    const float3 f = pair_contribution(a.phys, p1, v1, r1, q4, xyz(a.vel[a.ids[j] & 0x7fffffff]));
                    acc.x += fx_of(f.x); acc.y += fx_of(f.y); acc.z += fx_of(f.z);
                    ++acc.hits;
                }
                if (STATS) ++tests;
            }
        }
        if (DEBUG) {
            a.dbgCount[pid] = cnt;
            a.dbgSum[pid] = sum;
            a.dbgHits[pid] = acc.hits;
        } else if (acc.hits) {
            float4 f = a.frc[pid];
            f.x += fx_value(acc.x); f.y += fx_value(acc.y); f.z += fx_value(acc.z);
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


template <bool DEBUG, bool STATS>
__global__ void __launch_bounds__(256) pair_walk_kernel(const CollideArgs a)
{
    walk_slots<DEBUG, STATS>(a);
}
// debug view of a build the symmetric search steps aside for
__global__ void __launch_bounds__(256) pair_walk_debug_kernel(const CollideArgs a)
{
    if (*a.irregular != 0) walk_slots<true, false>(a);
}

// one thread per touching pair: both contributions, added to the ends' fixed-point accumulators
template <bool STATS>
__global__ void __launch_bounds__(256) pair_force_kernel(const CollideArgs a)
{
    const PairLists& L = a.pairs;
    const bool fallback = L.ctl[CTL_OVERFLOW] != 0 || *a.irregular != 0;
    const int count = fallback ? 0 : min(L.ctl[CTL_COUNT], L.cap);
    // a build with particles outside the grid, or one whose pair list overflowed: every slot walks its whole stencil
    if (fallback) walk_slots<false, STATS>(a);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const int2 pr = L.pairs[t];
        const int ti = a.ids[pr.x], tj = a.ids[pr.y];
        const float4 pi = a.spos[pr.x], pj = a.spos[pr.y];
        const float3 vi = xyz(a.vel[ti & 0x7fffffff]), vj = xyz(a.vel[tj & 0x7fffffff]);
        if (ti >= 0) {   // ghosts are partners only (forces go to owned particles, particle_collisions.cuh:36)
            const float3 c = pair_contribution(a.phys, xyz(pi), vi, pi.w, pj, vj);
            unsigned long long* acc = reinterpret_cast<unsigned long long*>(L.acc) + 3 * (size_t)ti;
            atomicAdd(acc, (unsigned long long)fx_of(c.x)); atomicAdd(acc + 1, (unsigned long long)fx_of(c.y)); atomicAdd(acc + 2, (unsigned long long)fx_of(c.z));
        }
        if (tj >= 0) {
            const float3 c = pair_contribution(a.phys, xyz(pj), vj, pj.w, pi, vi);
            unsigned long long* acc = reinterpret_cast<unsigned long long*>(L.acc) + 3 * (size_t)tj;
            atomicAdd(acc, (unsigned long long)fx_of(c.x)); atomicAdd(acc + 1, (unsigned long long)fx_of(c.y)); atomicAdd(acc + 2, (unsigned long long)fx_of(c.z));
        }
    }
    // the last CTA to leave rewinds the pair list for the next search
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&L.ctl[CTL_DONE], 1) == (int)gridDim.x - 1) {
            L.ctl[CTL_COUNT] = 0; L.ctl[CTL_OVERFLOW] = 0; L.ctl[CTL_DONE] = 0;
        }
    }
}

// one thread per particle, in particle order (coalesced): non-zero sums are folded into the float force and cleared.
// Staged entry points only: inside bcs_step the fold rides in the passes that read the forces next (wall apply, cell pass)
__global__ void __launch_bounds__(256) pair_fold_kernel(const CollideArgs a, int nParticles)
{
    const PairLists& L = a.pairs;
    for (int pid = blockIdx.x * blockDim.x + threadIdx.x; pid < nParticles; pid += gridDim.x * blockDim.x) {
        long long* acc = L.acc + 3 * (size_t)pid;
        const long long sx = acc[0], sy = acc[1], sz = acc[2];
        if ((sx | sy | sz) == 0) continue;   // a sum of exactly zero adds nothing
        float4 f = a.frc[pid];
        f.x += fx_value(sx); f.y += fx_value(sy); f.z += fx_value(sz);
        acc[0] = 0; acc[1] = 0; acc[2] = 0;
        a.frc[pid] = f;
    }
}

// Every slot scans its WHOLE stencil through the row directory (the per-slot walk of collide.cu without the compact cell
// index): the fallback for builds with particles outside the grid or an overflowing pair list, the A/B partner of the
// symmetric search (BCS_COLLIDE=walk), and - windows being plain linear cell ranges [c0, c0 + nb) exactly as
// particle_collisions.cuh:53-83 forms them - correct for any key, clamped ones included.
}  // namespace

void launch_particle_collisions_rows(const CollideArgs& a, cudaStream_t st)
{
    const int threads = PS_THREADS, allBlocks = (a.n + threads - 1) / threads;
    const int blocks = a.nDev ? std::min(allBlocks, 2 * BOUNDED_BLOCKS) : allBlocks;
    const int walkBlocks = std::max(1, std::min((a.n + 255) / 256, BOUNDED_BLOCKS));
    const bool dbg = a.dbgCount != nullptr, slab = a.nDev != nullptr;
    const int foldBlocks = std::min((a.n + 255) / 256, 2 * BOUNDED_BLOCKS);
    if (a.fullWalk) {
        if (dbg) BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<true, false><<<walkBlocks, 256, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<false, true><<<walkBlocks, 256, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, pair_walk_kernel<false, false><<<walkBlocks, 256, 0, st>>>(a));
        BCS_CUDA(cudaGetLastError());
        return;
    }
    // template dispatch over (debug, slab, stats)
    auto search = [&](auto D, auto SL, auto ST) {
        BCS_LAUNCH("particle_collisions", st, pair_search_kernel<decltype(D)::value, decltype(SL)::value, decltype(ST)::value><<<blocks, threads, 0, st>>>(a));
    };
    using T = std::true_type;
    using F = std::false_type;
    if (dbg) {
        // candidate counts / checksums / touching pairs per particle (test instrumentation): both ends of a pair are
        // credited with integer atomics by the same search the production path runs; a build with particles outside the
        // grid takes the per-slot walk instead (the search steps aside)
        BCS_CUDA(cudaMemsetAsync(a.dbgCount, 0, (size_t)a.n * sizeof(int), st));
        BCS_CUDA(cudaMemsetAsync(a.dbgSum, 0, (size_t)a.n * sizeof(unsigned long long), st));
        BCS_CUDA(cudaMemsetAsync(a.dbgHits, 0, (size_t)a.n * sizeof(int), st));
        if (slab) search(T{}, T{}, F{});
        else search(T{}, F{}, F{});
        BCS_LAUNCH("pair_walk", st, pair_walk_debug_kernel<<<walkBlocks, 256, 0, st>>>(a));
        BCS_CUDA(cudaGetLastError());
        return;
    }
    if (slab) {
        if (a.stats) search(F{}, T{}, T{});
        else search(F{}, T{}, F{});
    } else {
        if (a.stats) search(F{}, F{}, T{});
        else search(F{}, F{}, F{});
    }
    // sized for a few touching pairs per ten particles; strides over the device-side pair count (or, in the fall-back,
    // over the sorted slots)
    const int pairBlocks = std::max(1, std::min((a.n / 4 + 255) / 256, BOUNDED_BLOCKS));
    if (a.stats) BCS_LAUNCH("pair_force", st, pair_force_kernel<true><<<pairBlocks, 256, 0, st>>>(a));
    else BCS_LAUNCH("pair_force", st, pair_force_kernel<false><<<pairBlocks, 256, 0, st>>>(a));
    if (!a.deferFold) BCS_LAUNCH("pair_fold", st, pair_fold_kernel<<<foldBlocks, 256, 0, st>>>(a, a.n));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

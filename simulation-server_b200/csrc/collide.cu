// Particle-particle collision / repulsion pass over the 27-cell neighbourhood - the paths that walk the per-cell index
// (reference semantics; clean semantics on dense scenes).  Sparse scenes (the 1 M-particle bench workload) build the
// row-directory grid instead and collide through the symmetric pair search of pairs.cu.
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
#include "pair_device.cuh"

#include <algorithm>

namespace bcs {

// detectCollision + addResilientForceOnCollision with intensityCoefficient 0.5
__device__ __forceinline__ void test_pair(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4,
                                          const float r2, const float4* __restrict__ svel, int j, PairAccum& acc)
{
    if (pair_touches(p1, r1, q4, r2)) pair_force(ph, p1, v1, r1, q4, svel, j, acc);
}

// Hits are rare (~7 % of the particles per step) and land in different loop iterations of different lanes: evaluated on
// the spot, each costs the whole warp a ~120-instruction pass with one lane active.  The walk therefore only RECORDS the
// touching candidates (up to DEFER per particle, in encounter order); afterwards all lanes evaluate their first hit in
// one common pass, then their second ...  Same forces, same summation order per particle.
constexpr int DEFER = 3;
struct Deferred {
    int j[DEFER];
    int n;
};
__device__ __forceinline__ void defer_or_evaluate(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, const float4 q4,
                                                  const float4* __restrict__ svel, int j, PairAccum& acc, Deferred& df,
                                                  const float4* __restrict__ spos)
{
    if (df.n < DEFER) {
#pragma unroll
        for (int k = 0; k < DEFER; ++k)
            if (k == df.n) df.j[k] = j;
        ++df.n;
    } else {
        // more touching candidates than slots (dense clusters): flush in order, keep the order
#pragma unroll
        for (int k = 0; k < DEFER; ++k) pair_force(ph, p1, v1, r1, spos[df.j[k]], svel, df.j[k], acc);
        df.n = 1;
        df.j[0] = j;
    }
}
__device__ __forceinline__ void evaluate_deferred(const PhysDev& ph, const float3 p1, const float3 v1, const float r1, PairAccum& acc,
                                                  const Deferred& df, const float4* __restrict__ spos, const float4* __restrict__ svel)
{
#pragma unroll
    for (int k = 0; k < DEFER; ++k)
        if (k < df.n) pair_force(ph, p1, v1, r1, spos[df.j[k]], svel, df.j[k], acc);
}

// clean-semantics walk of one sorted slot through the compact cell index (global memory): the default of this file
// (dense scenes, where the grid build is the compact cell index; sparse scenes take the row directory + pairs.cu)
constexpr int WALK_THREADS = 128;
template <bool DEBUG, bool STATS, bool FLAT>
__device__ __forceinline__ void clean_slot_walk(const CollideArgs& a, int slot, const float3 p1, const float3 v1, const float r1, int cell, int x0,
                                                int x1, int y0, int y1, int z0, int z1, PairAccum& acc, int& cnt, unsigned long long& sum,
                                                unsigned long long& myTests, int2 (*seg)[WALK_THREADS] = nullptr)
{
    const GridDev& g = a.grid;
    const int plane = g.nx * g.ny;
    // The 9 stencil rows are resolved level by level, branch-free and all lanes alike, so that the dependent loads
    // of ALL rows are in flight together (a row-after-row walk chains 3 L2 round trips per non-empty row):
    //   level 1  occupancy bits.  A row is the <= 3 x-adjacent cells [row+x0, row+x1]: adjacent cell ids, i.e.
    //            adjacent bits of the occupancy mask (two words, funnel-shifted)
    //   level 2  rank of the row's first occupied cell = cellRank[word] + popc(bits below it)
    //   level 3  occupied cells of a row have consecutive ranks and ONE contiguous slot range
    //            [occStart[rank], occStart[rank + popc(bits)])
    //   level 4  the candidates
    const int nb = x1 - x0 + 1;
    const unsigned nbm = (1u << nb) - 1u;
    int first[9];      // first occupied cell of the row (or -1)
    int nocc[9];
    int rank[9];
    {
        unsigned below[9]; // occupancy bits below it inside its mask word
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int dz = r / 3 - 1, dy = r % 3 - 1;
            const int c0 = cell + dz * plane + dy * g.nx + x0;
            const bool in = dz >= z0 && dz <= z1 && dy >= y0 && dy <= y1 && c0 >= 0 && c0 + nb <= g.cells;
            const int w0 = in ? c0 >> 5 : 0;
            const unsigned lo = __ldg(a.cellMask + w0), hi = __ldg(a.cellMask + w0 + 1);
            const unsigned bits = in ? __funnelshift_r(lo, hi, c0 & 31) & nbm : 0u;
            const int f = c0 + __ffs(bits) - 1;
            first[r] = bits ? f : -1;
            nocc[r] = __popc(bits);
            below[r] = ((f >> 5) == w0 ? lo : hi) & ((1u << (f & 31)) - 1u);
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) rank[r] = __ldg(a.cellRank + (first[r] >= 0 ? first[r] >> 5 : 0)) + __popc(below[r]);
    }
    int lo[9], hi[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int rk = first[r] >= 0 ? rank[r] : 0;
        lo[r] = __ldg(a.occStart + rk);
        hi[r] = first[r] >= 0 ? __ldg(a.occStart + rk + nocc[r]) : lo[r];
    }
    Deferred df;
    df.n = 0;
    if (FLAT) {
        // Level 4, flattened.  Walking row after row costs a warp sum_r max_lanes(len_r) iterations; at ~4 candidates per
        // particle spread over 9 rows nearly every row is non-empty for SOME lane, so most iterations run a few lanes.
        // The non-empty slot ranges are therefore parked compactly in shared memory and every lane walks its own list
        // end to end: iteration k of the warp is candidate k of every lane (max_lanes(sum_r len_r) iterations), and the
        // next candidate's load is issued before the current one is tested.  Same candidates, same order per particle.
        const int tid = threadIdx.x;
        int nseg = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r)
            if (hi[r] > lo[r]) seg[nseg++][tid] = make_int2(lo[r], hi[r]);
        int s = 0, j = 0, e = 0;
        bool live = nseg > 0;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            const int2 g0 = seg[0][tid];
            j = g0.x; e = g0.y;
            q = a.spos[j];
        }
        while (live) {
            const int jc = j;
            const float4 q4 = q;
            if (++j == e) {
                if (++s < nseg) {
                    const int2 g1 = seg[s][tid];
                    j = g1.x; e = g1.y;
                } else {
                    live = false;
                }
            }
            if (live) q = a.spos[j];
            if (jc == slot) continue;
            if (DEBUG) {
                const int qid = __float_as_int(a.svel[jc].w) & 0x7fffffff;
                ++cnt; sum += (unsigned long long)(qid + 1) * 0x9E3779B97F4A7C15ull;
            }
            if (pair_touches(p1, r1, q4, q4.w)) defer_or_evaluate(a.phys, p1, v1, r1, q4, a.svel, jc, acc, df, a.spos);
            if (STATS) ++myTests;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            for (int j = lo[r]; j < hi[r]; ++j) {
                if (j == slot) continue;
                const float4 q4 = a.spos[j];
                if (DEBUG) {
                    const int qid = __float_as_int(a.svel[j].w) & 0x7fffffff;
                    ++cnt; sum += (unsigned long long)(qid + 1) * 0x9E3779B97F4A7C15ull;
                }
                if (pair_touches(p1, r1, q4, q4.w)) defer_or_evaluate(a.phys, p1, v1, r1, q4, a.svel, j, acc, df, a.spos);
                if (STATS) ++myTests;
            }
        }
    }
    evaluate_deferred(a.phys, p1, v1, r1, acc, df, a.spos, a.svel);
}

// ------------------------------------------------------------------------------------------------------------
// Clean semantics, A/B alternative (BCS_COLLIDE=tiled; measured slower than the walk above on the bench scenes, kept
// for comparison): the neighbour runs are STAGED IN SHARED MEMORY per tile of sorted slots.
//
// Cell ids are z-major (uniform_grid.cu:33-35), so a tile of 256 consecutive sorted slots covers a run of x-rows
// (y,z) of one z-layer, and everything its particles can collide with lies in THREE contiguous slot windows - the
// rows [first-1, last+1] of the layers z-1, z, z+1.  The CTA finds the six window bounds through the compact cell
// index (six lookups instead of ~36 per particle), copies the three windows (positions + keys, coalesced) into
// shared memory and builds a row-start table there.  After that a particle resolves its 9 stencil rows with two
// shared-memory reads each and scans only shared memory: all lanes run the same short loops, where the per-thread
// walk through the global index ran with 13 of 32 lanes active and ~4 dependent L2 round trips per row.
// Candidate sets (checked in debug mode against the reference's), hits and forces are identical to the cell-index
// walk; a tile whose windows or row span exceed the shared-memory capacity (layer ends, pathological clustering)
// takes that walk instead.
// ------------------------------------------------------------------------------------------------------------
constexpr int TILE = 256;        // sorted slots per CTA
constexpr int WCAP = 512;        // window capacity (slots)
constexpr int RCAP = 256;        // window capacity (x-rows)

__device__ __forceinline__ int fast_div(unsigned n, unsigned long long magic, int shift)
{
    return (int)(((unsigned long long)n * magic) >> shift);   // exact for n < 2^31 (magic = ceil(2^shift / d), shift = 32 + ceil(log2 d))
}

// first sorted slot whose key is >= K, through the compact cell index (cellRank is a full exclusive prefix in the
// counting-sort grid build)
__device__ __forceinline__ int first_slot_at_or_after(const CollideArgs& a, int K, int n)
{
    if (K >= a.grid.cells) return n;
    if (K < 0) K = 0;
    const int w = K >> 5;
    const int rank = __ldg(a.cellRank + w) + __popc(__ldg(a.cellMask + w) & ((1u << (K & 31)) - 1u));
    return __ldg(a.occStart + rank);
}

template <bool DEBUG, bool STATS>
__global__ void __launch_bounds__(TILE) particle_collisions_tiled_kernel(const CollideArgs a)
{
    __shared__ float4 wpos[3][WCAP];
    __shared__ int wkey[3][WCAP];
    __shared__ unsigned short wrow[3][RCAP + 2];
    __shared__ int wlo[3], wcnt[3], wrowBase[3], wrows[3];
    __shared__ int sFallback;

    const GridDev& g = a.grid;
    const int tid = threadIdx.x;
    const int n = a.nDev ? *a.nDev : a.n;
    const int tileFirst = blockIdx.x * TILE;
    if (tileFirst >= n) return;
    const int tileLast = min(n, tileFirst + TILE) - 1;
    const int nRows = g.ny * g.nz;

    if (tid == 0) sFallback = 0;
    if (tid < 3) {
        const int dz = tid - 1;
        const int rowF = fast_div((unsigned)a.keys[tileFirst], a.nxMagic, a.nxShift), rowL = fast_div((unsigned)a.keys[tileLast], a.nxMagic, a.nxShift);
        const int rLo = max(0, min(nRows - 1, rowF + dz * g.ny - 1)), rHi = max(0, min(nRows - 1, rowL + dz * g.ny + 1));
        const int lo = first_slot_at_or_after(a, rLo * g.nx, n);
        const int hi = rHi + 1 >= nRows ? n : first_slot_at_or_after(a, (rHi + 1) * g.nx, n);
        wlo[tid] = lo; wcnt[tid] = max(0, hi - lo); wrowBase[tid] = rLo; wrows[tid] = rHi - rLo + 1;
    }
    __syncthreads();
    if (tid < 3 && (wcnt[tid] > WCAP || wrows[tid] > RCAP)) sFallback = 1;
    __syncthreads();
    const bool fallback = sFallback != 0;

    if (!fallback) {
        // stage the three windows and clear their row tables
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int lo = wlo[w], cnt = wcnt[w];
            for (int j = tid; j < cnt; j += TILE) {
                wpos[w][j] = a.spos[lo + j];
                wkey[w][j] = a.keys[lo + j];
            }
        }
        __syncthreads();
        // row-start table: window slot j opens every row in (row(j-1), row(j)]; the tail rows end at cnt
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int cnt = wcnt[w], base = wrowBase[w], rows = wrows[w];
            for (int j = tid; j <= cnt; j += TILE) {
                const int rPrev = j == 0 ? -1 : fast_div((unsigned)wkey[w][j - 1], a.nxMagic, a.nxShift) - base;
                const int rThis = j == cnt ? rows : fast_div((unsigned)wkey[w][j], a.nxMagic, a.nxShift) - base;
                for (int q = rPrev + 1; q <= min(rThis, rows); ++q) wrow[w][q] = (unsigned short)j;
            }
        }
        __syncthreads();
    }

    const int slot = tileFirst + tid;
    unsigned long long myTests = 0;
    int myHits = 0;
    if (slot < n && __float_as_int(a.svel[slot].w) >= 0) {
        const float4 p4 = a.spos[slot];
        const float4 v4 = a.svel[slot];
        const int pid = __float_as_int(v4.w) & 0x7fffffff;
        const float3 p1 = xyz(p4), v1 = xyz(v4);
        const float r1 = p4.w;
        const int cell = a.keys[slot];
        int x0, x1, y0, y1, z0, z1;
        stencil_range(axis_cell_raw(p1.x, g.minx, g.csx), g.nx, x0, x1);
        stencil_range(axis_cell_raw(p1.y, g.miny, g.csy), g.ny, y0, y1);
        stencil_range(axis_cell_raw(p1.z, g.minz, g.csz), g.nz, z0, z1);
        PairAccum acc = pair_accum_zero();
        int cnt = 0;
        unsigned long long sum = 0;
        if (fallback) {
            clean_slot_walk<DEBUG, STATS, false>(a, slot, p1, v1, r1, cell, x0, x1, y0, y1, z0, z1, acc, cnt, sum, myTests);
        } else {
            const int row0 = fast_div((unsigned)cell, a.nxMagic, a.nxShift);
            const int own = slot - wlo[1];
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const int dz = w - 1;
                if (dz < z0 || dz > z1) continue;
                const int base = wrowBase[w], rows = wrows[w], glo = wlo[w];
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    if (dy < y0 || dy > y1) continue;
                    const int row = row0 + dz * g.ny + dy, r = row - base;
                    if (r < 0 || r >= rows) continue;   // beyond the grid: no such cells
                    const int e = wrow[w][r + 1];
                    const int kLo = cell + (dz * g.ny + dy) * g.nx + x0, kHi = kLo + (x1 - x0);
                    // first slot of the row inside the x window (a row holds up to a chord of cells: binary search)
                    int s = wrow[w][r], hi = e;
                    while (s < hi) {
                        const int mid = (s + hi) >> 1;
                        if (wkey[w][mid] < kLo) s = mid + 1; else hi = mid;
                    }
                    for (int j = s; j < e; ++j) {
                        if (wkey[w][j] > kHi) break;
                        if (w == 1 && j == own) continue;
                        const float4 q4 = wpos[w][j];
                        if (DEBUG) {
                            const int qid = __float_as_int(a.svel[glo + j].w) & 0x7fffffff;
                            ++cnt; sum += (unsigned long long)(qid + 1) * 0x9E3779B97F4A7C15ull;
                        }
                        test_pair(a.phys, p1, v1, r1, q4, q4.w, a.svel, glo + j, acc);
                        if (STATS) ++myTests;
                    }
                }
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
        myHits = acc.hits;
    }
    if (STATS && !DEBUG) {
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

template <bool REFERENCE, bool DEBUG, bool STATS, bool FLAT = false, int MINB = 10>
__global__ void __launch_bounds__(WALK_THREADS, MINB) particle_collisions_kernel(const CollideArgs a)
{
    __shared__ int2 seg[FLAT ? 9 : 1][WALK_THREADS];
    unsigned long long myTests = 0;
    int myHits = 0;
    // slab mode: the active count lives on the device and ghost slots (bit 31 of the id) are candidates only; the launch
    // is a bounded grid that strides over the count (one round without slabs: the grid covers every slot)
    const int nActive = a.nDev ? *a.nDev : a.n;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nActive; slot += gridDim.x * blockDim.x) {
        // the slot's own record: all three loads issued together, before the activity test
        const float4 v4 = a.svel[slot];
        const float4 p4 = a.spos[slot];
        const int cell = a.keys[slot];
        if (__float_as_int(v4.w) >= 0) {
            const GridDev& g = a.grid;
            const int tag = __float_as_int(v4.w);
            const int pid = tag & 0x7fffffff;
            const float3 p1 = xyz(p4), v1 = xyz(v4);
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

            PairAccum acc = pair_accum_zero();
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
                clean_slot_walk<DEBUG, STATS, FLAT>(a, slot, p1, v1, r1, cell, x0, x1, y0, y1, z0, z1, acc, cnt, sum, myTests, seg);
            }
            if (DEBUG) {
                a.dbgCount[pid] = cnt;
                a.dbgSum[pid] = sum;
                a.dbgHits[pid] = acc.hits;
            } else if (acc.hits) {
                // F[pid] += acc: only this thread updates this particle, so the three reductions (performed in L2, no value
                // returned - the warp does not wait for a load at its very end) give the same IEEE sum as load-add-store
                float* f = reinterpret_cast<float*>(a.frc + pid);
                atomicAdd(f, fx_value(acc.x)); atomicAdd(f + 1, fx_value(acc.y)); atomicAdd(f + 2, fx_value(acc.z));
            }
            myHits += acc.hits;
        }
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
    const int threads = WALK_THREADS, allBlocks = (a.n + threads - 1) / threads;
    const int blocks = a.nDev ? std::min(allBlocks, 2 * BOUNDED_BLOCKS) : allBlocks;
    const bool dbg = a.dbgCount != nullptr;
    if (!a.reference && a.tiled) {
        const int tiles = (a.n + TILE - 1) / TILE;
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_tiled_kernel<true, false><<<tiles, TILE, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_tiled_kernel<false, true><<<tiles, TILE, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_tiled_kernel<false, false><<<tiles, TILE, 0, st>>>(a));
    } else if (a.reference) {
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, true, false><<<blocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<true, false, false><<<blocks, threads, 0, st>>>(a));
    } else if (!a.rows) {
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, true, false, true><<<blocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, true, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, false, true><<<blocks, threads, 0, st>>>(a));
    } else {
        if (dbg) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, true, false><<<blocks, threads, 0, st>>>(a));
        else if (a.stats) BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("particle_collisions", st, particle_collisions_kernel<false, false, false><<<blocks, threads, 0, st>>>(a));
    }
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

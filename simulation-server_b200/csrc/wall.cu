// Particle vs. vein-wall collisions through a lazily rebuilt wall grid (production path, clean semantics).
//
// Stands in for sim::detectVeinCollisions<UniformGrid> (simulation/vein_collisions.cu:63-277)
//   -> calculateSideCollisions (simulation/vein_collisions.cuh:60-93) -> realCollisionDetection (:11-45).
//
// What the stage can observe: the FIRST triangle in the reference's traversal order
// (27 triangle-grid cells, x outer / y / z inner, sorted slots inside a cell) that the ray of the particle hits,
// and only if that hit lies within veinImpactDistance.  So per particle:
//   phase A  the first triangle in traversal order with a hit at t <= reach.  Such a hit point lies on the segment
//            pos .. pos + reach*dir, so the triangle is listed in one of the wall-grid cells that segment's box
//            overlaps.  A particle in the bulk of the lumen reads a few (empty) cell ranges and is done; near the
//            wall the cell's slab (mean normal of its triangles) rejects rays that run alongside the wall.
//   phase B  (the few particles with a near hit) one warp checks that no EARLIER triangle in traversal order is
//            hit farther away - the reference would have returned that one and done nothing.
//
// The wall moves (vertex springs + collision splats), but slowly: the grid, its slabs and the box hierarchy of
// phase B are padded by `margin` and rebuilt only when a vertex has left that margin (flag raised by the vertex
// integrator).  Triangle tests always gather the LIVE vertices, so culling stays conservative and the result is
// identical to the exhaustive traversal (tests/test_gpu_parity.py compares them bitwise).
//
// B200 mapping: one thread per particle for phase A (pos/vel read once, coalesced; everything else is L2 resident:
// 4 B per wall-grid cell, the lists, the vertices), a persistent one-CTA-per-SM kernel with software grid barriers
// for the rare rebuild, so that a step costs ONE empty launch for the structure instead of refitting 600 k
// triangles and their boxes every step.
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "vein_device.cuh"

#include <algorithm>

namespace bcs {

namespace {

constexpr int REBUILD_THREADS = 1024;

__device__ __forceinline__ int wall_axis(float p, float o, float invh, int n)
{
    // monotone in p and clamped: overlapping intervals map to overlapping cell ranges, also outside the grid
    const float q = floorf((p - o) * invh);
    return (int)fminf(fmaxf(q, 0.f), (float)(n - 1));
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, unsigned nBlocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nBlocks;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < target) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void slot_vertices(const unsigned* __restrict__ vidx, const int* __restrict__ triIds, const float4* __restrict__ vpos,
                                              int slot, float3& v0, float3& v1, float3& v2)
{
    const int tri = triIds[slot];
    v0 = xyz(vpos[vidx[3 * tri]]); v1 = xyz(vpos[vidx[3 * tri + 1]]); v2 = xyz(vpos[vidx[3 * tri + 2]]);
}

// ------------------------------------------------------------------------------------------------------------
// rebuild: count -> scan -> fill -> slabs + box hierarchy, one persistent kernel (grid = one CTA per SM)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REBUILD_THREADS, 1)
wall_rebuild_kernel(const WallGridDev w, const GridDev tg, int V, int T, const float4* __restrict__ vpos, const unsigned* __restrict__ vidx,
                    const int* __restrict__ triIds, const int* __restrict__ tcellStart, const int* __restrict__ tcellEnd,
                    const unsigned char* __restrict__ groupLocal, const unsigned char* __restrict__ triCellLocal)
{
    if (*w.dirty == 0) return;   // nobody writes the flag while this kernel runs: the whole grid takes the same branch
    __shared__ int sScan[REBUILD_THREADS / 32];
    __shared__ int sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = gridDim.x, gsize = nb * REBUILD_THREADS, gtid = blockIdx.x * REBUILD_THREADS + tid;
    unsigned target = 0;
    const float pad = BOX_PAD + w.margin;

    // ---- P1: clear the counts, remember the vertex positions this build is made from
    for (int i = gtid; i < w.cells; i += gsize) w.cursor[i] = 0;
    for (int i = gtid; i < V; i += gsize) w.vposBuilt[i] = vpos[i];
    grid_barrier(w.barrier, target, nb);

    // ---- P2: count the cells every triangle's padded box overlaps; boxes of the groups of 8 sorted slots
    const int T8 = (T + 7) & ~7;
    for (int s = gtid; s < ((T8 + 31) & ~31); s += gsize) {
        float lox = 3e38f, loy = 3e38f, loz = 3e38f, hix = -3e38f, hiy = -3e38f, hiz = -3e38f;
        const bool local = s < T && (!groupLocal || groupLocal[s >> 3]);
        if (local) {
            float3 v0, v1, v2;
            slot_vertices(vidx, triIds, vpos, s, v0, v1, v2);
            lox = fminf(v0.x, fminf(v1.x, v2.x)) - pad; hix = fmaxf(v0.x, fmaxf(v1.x, v2.x)) + pad;
            loy = fminf(v0.y, fminf(v1.y, v2.y)) - pad; hiy = fmaxf(v0.y, fmaxf(v1.y, v2.y)) + pad;
            loz = fminf(v0.z, fminf(v1.z, v2.z)) - pad; hiz = fmaxf(v0.z, fmaxf(v1.z, v2.z)) + pad;
            const int x0 = wall_axis(lox, w.ox, w.invh, w.nx), x1 = wall_axis(hix, w.ox, w.invh, w.nx);
            const int y0 = wall_axis(loy, w.oy, w.invh, w.ny), y1 = wall_axis(hiy, w.oy, w.invh, w.ny);
            const int z0 = wall_axis(loz, w.oz, w.invh, w.nz), z1 = wall_axis(hiz, w.oz, w.invh, w.nz);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y)
                    for (int x = x0; x <= x1; ++x) atomicAdd(&w.cursor[(z * w.ny + y) * w.nx + x], 1);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
            loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
            loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
        }
        if ((s & 7) == 0 && local) w.groupBox[s >> 3] = Aabb{lox, loy, loz, hix, hiy, hiz};
    }
    grid_barrier(w.barrier, target, nb);

    // ---- P3: exclusive scan of the counts over the cells: per-block chunk totals, then the chunk itself
    const int chunk = (w.cells + nb - 1) / nb;
    const int c0 = min(w.cells, (int)blockIdx.x * chunk), c1 = min(w.cells, c0 + chunk);
    {
        int sum = 0;
        for (int c = c0 + tid; c < c1; c += REBUILD_THREADS) {
            const int n = w.cursor[c];
            sum += n;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sScan[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int k = 0; k < REBUILD_THREADS / 32; ++k) t += sScan[k];
            w.blockSums[blockIdx.x] = t;
        }
    }
    grid_barrier(w.barrier, target, nb);
    {
        int before = 0;
        for (int b = tid; b < (int)blockIdx.x; b += REBUILD_THREADS) before += w.blockSums[b];
#pragma unroll
        for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        __syncthreads();
        if (lane == 0) sScan[warp] = before;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int k = 0; k < REBUILD_THREADS / 32; ++k) t += sScan[k];
            sBase = t;
        }
        __syncthreads();
        for (int base = c0; base < c1; base += REBUILD_THREADS) {
            const int c = base + tid;
            const int n = c < c1 ? w.cursor[c] : 0;
            const int mine = n;
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            __syncthreads();
            if (lane == 31) sScan[warp] = incl;
            __syncthreads();
            int wbase = 0;
            for (int k = 0; k < warp; ++k) wbase += sScan[k];
            const int start = sBase + wbase + incl - mine;
            if (c < c1) {
                w.start[c] = start;
                w.cursor[c] = start;
            }
            __syncthreads();
            if (tid == REBUILD_THREADS - 1) sBase = start + mine;
            __syncthreads();
        }
        if (blockIdx.x == nb - 1 && tid == 0) {
            w.start[w.cells] = sBase;
            if (sBase > w.cap) *w.overflow = 1;
            *w.builds += 1ull;
        }
    }
    grid_barrier(w.barrier, target, nb);

    // ---- P4: fill the lists (order inside a cell is irrelevant: phase A takes a minimum over traversal keys)
    for (int s = gtid; s < T; s += gsize) {
        if (groupLocal && !groupLocal[s >> 3]) continue;
        float3 v0, v1, v2;
        slot_vertices(vidx, triIds, vpos, s, v0, v1, v2);
        const int x0 = wall_axis(fminf(v0.x, fminf(v1.x, v2.x)) - pad, w.ox, w.invh, w.nx), x1 = wall_axis(fmaxf(v0.x, fmaxf(v1.x, v2.x)) + pad, w.ox, w.invh, w.nx);
        const int y0 = wall_axis(fminf(v0.y, fminf(v1.y, v2.y)) - pad, w.oy, w.invh, w.ny), y1 = wall_axis(fmaxf(v0.y, fmaxf(v1.y, v2.y)) + pad, w.oy, w.invh, w.ny);
        const int z0 = wall_axis(fminf(v0.z, fminf(v1.z, v2.z)) - pad, w.oz, w.invh, w.nz), z1 = wall_axis(fmaxf(v0.z, fmaxf(v1.z, v2.z)) + pad, w.oz, w.invh, w.nz);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y)
                for (int x = x0; x <= x1; ++x) {
                    const int p = atomicAdd(&w.cursor[(z * w.ny + y) * w.nx + x], 1);
                    if (p < w.cap) w.list[p] = s;
                }
    }
    grid_barrier(w.barrier, target, nb);

    // ---- P5a: 32-byte record of every cell (one warp per cell): list range + slab along the mean normal
    const int nWarps = gsize >> 5, gwarp = gtid >> 5;
    for (int c = gwarp; c < w.cells; c += nWarps) {
        const int s = w.start[c], e = min(w.start[c + 1], w.cap);
        if (e <= s) {
            if (lane == 0) { w.rec[2 * c] = make_int4(s, 0, 0, 0); w.occ[c] = 0; }
            continue;
        }
        float3 nsum = f3(0.f, 0.f, 0.f);
        for (int i = s + lane; i < e; i += 32) {
            float3 v0, v1, v2;
            slot_vertices(vidx, triIds, vpos, w.list[i], v0, v1, v2);
            nsum = nsum + cross(v1 - v0, v2 - v0);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            nsum.x += __shfl_xor_sync(0xffffffffu, nsum.x, o); nsum.y += __shfl_xor_sync(0xffffffffu, nsum.y, o);
            nsum.z += __shfl_xor_sync(0xffffffffu, nsum.z, o);
        }
        float3 n = normalize(nsum);
        if (n.x == 0.f && n.y == 0.f && n.z == 0.f) n = f3(1.f, 0.f, 0.f);   // degenerate: any unit axis keeps the slab valid
        float dmin = 3e38f, dmax = -3e38f;
        for (int i = s + lane; i < e; i += 32) {
            float3 v0, v1, v2;
            slot_vertices(vidx, triIds, vpos, w.list[i], v0, v1, v2);
            const float d0 = dot(n, v0), d1 = dot(n, v1), d2 = dot(n, v2);
            dmin = fminf(dmin, fminf(d0, fminf(d1, d2)));
            dmax = fmaxf(dmax, fmaxf(d0, fmaxf(d1, d2)));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
        if (lane == 0) {
            w.rec[2 * c] = make_int4(s, e - s, __float_as_int(n.x), __float_as_int(n.y));
            w.rec[2 * c + 1] = make_int4(__float_as_int(n.z), __float_as_int(dmin - pad), __float_as_int(dmax + pad), 0);
            w.occ[c] = 1;
        }
    }
    // ---- P5b: boxes of the triangle-grid cells (phase B): union of the group boxes the cell's slot range touches
    for (int c = gwarp; c < tg.cells; c += nWarps) {
        if (triCellLocal && !triCellLocal[c]) continue;
        const int s = tcellStart[c], e = tcellEnd[c];
        Aabb b{3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
        if (e >= s)
            for (int g = (s >> 3) + lane; g <= (e >> 3); g += 32) {
                const Aabb q = w.groupBox[g];
                b.lox = fminf(b.lox, q.lox); b.loy = fminf(b.loy, q.loy); b.loz = fminf(b.loz, q.loz);
                b.hix = fmaxf(b.hix, q.hix); b.hiy = fmaxf(b.hiy, q.hiy); b.hiz = fmaxf(b.hiz, q.hiz);
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            b.lox = fminf(b.lox, __shfl_xor_sync(0xffffffffu, b.lox, o)); b.hix = fmaxf(b.hix, __shfl_xor_sync(0xffffffffu, b.hix, o));
            b.loy = fminf(b.loy, __shfl_xor_sync(0xffffffffu, b.loy, o)); b.hiy = fmaxf(b.hiy, __shfl_xor_sync(0xffffffffu, b.hiy, o));
            b.loz = fminf(b.loz, __shfl_xor_sync(0xffffffffu, b.loz, o)); b.hiz = fmaxf(b.hiz, __shfl_xor_sync(0xffffffffu, b.hiz, o));
        }
        if (lane == 0) w.cellBox[c] = b;
    }

    // ---- P6: near mask = occupancy dilated by 2 cells per axis (a segment of length reach <= 2 cells starts at most
    // 2 cells away from every cell it touches), separable: x, then y, then z
    grid_barrier(w.barrier, target, nb);
    unsigned char* t0 = w.nearTmp;
    unsigned char* t1 = w.nearTmp + w.cells;
    for (int c = gtid; c < w.cells; c += gsize) {
        const int x = c % w.nx;
        unsigned char v = 0;
        for (int d = -2; d <= 2; ++d)
            if (x + d >= 0 && x + d < w.nx) v |= w.occ[c + d];
        t0[c] = v;
        // occupancy of the three x-adjacent cells starting here, one byte: a (y,z) row of a segment's box is ONE load
        w.occ3[c] = (unsigned char)(w.occ[c] | (x + 1 < w.nx ? w.occ[c + 1] << 1 : 0) | (x + 2 < w.nx ? w.occ[c + 2] << 2 : 0));
    }
    grid_barrier(w.barrier, target, nb);
    for (int c = gtid; c < w.cells; c += gsize) {
        const int y = (c / w.nx) % w.ny;
        unsigned char v = 0;
        for (int d = -2; d <= 2; ++d)
            if (y + d >= 0 && y + d < w.ny) v |= t0[c + d * w.nx];
        t1[c] = v;
    }
    grid_barrier(w.barrier, target, nb);
    for (int c = gtid; c < w.cells; c += gsize) {
        const int z = c / (w.nx * w.ny);
        unsigned char v = 0;
        for (int d = -2; d <= 2; ++d)
            if (z + d >= 0 && z + d < w.nz) v |= t1[c + d * w.nx * w.ny];
        w.near[c] = v;
    }

    // ---- leave: the last block out re-arms the barrier for the next rebuild
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(w.barrier + 1, 1u) == (unsigned)nb - 1u) {
            w.barrier[0] = 0u;
            w.barrier[1] = 0u;
        }
    }
}

// static per sorted slot: triangle id and the triangle-grid cell the slot belongs to (the triangle grid is built
// once from the initial centres)
__global__ void __launch_bounds__(256) wall_slot_info_kernel(const int* __restrict__ keys, const int* __restrict__ triIds, const unsigned* __restrict__ vidx,
                                                             int T, GridDev g, int4* __restrict__ info, int4* __restrict__ verts)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T) return;
    const int key = keys[s], plane = g.nx * g.ny;
    const int z = key / plane, y = (key - z * plane) / g.nx, x = key - z * plane - y * g.nx;
    const int tri = triIds[s];
    info[s] = make_int4(tri, x, y, z);
    verts[s] = make_int4((int)vidx[3 * tri], (int)vidx[3 * tri + 1], (int)vidx[3 * tri + 2], tri);
}

// ------------------------------------------------------------------------------------------------------------
// Phase A in two kernels, so that every pass runs with full warps and nothing waits at a CTA barrier:
//   A1  thread per particle (streaming, light): near-mask byte; occupancy of the <= 3x3x3 wall-grid cells under the
//       segment's box (branch-free: all loads in flight together); slab test per occupied cell.  Survivors are
//       appended to a global queue of (particle, cell) entries - one atomicAdd per WARP - and the particle to the
//       candidate list, its best[] word reset.
//   A2  8 lanes per queue entry, entries spread evenly over a persistent grid: one lane per listed triangle,
//       stencil membership from the static slot table, Moeller-Trumbore on the live vertices; near hits race with a
//       64-bit atomicMin on best[particle] = (traversal key, slot).
// Particles outside the triangle grid and queue overflow (pathological clustering) take the sequential search
// (vein_device.cuh) in phase B.
// ------------------------------------------------------------------------------------------------------------
constexpr unsigned long long NO_HIT = ~0ull, SEQUENTIAL = ~0ull - 1ull;

struct ParticleFrame {     // what the triangle tests need to know about a particle; recomputed identically in A2 and B
    float3 pos, dir;
    int pcx, pcy, pcz;     // its triangle-grid cell
    int x0, x1, y0, y1, z0, z1;   // trimmed stencil (vein_collisions.cu:76-230)
};

__device__ __forceinline__ ParticleFrame particle_frame(const VeinCollideArgs& a, int pid)
{
    const GridDev& g = a.tgrid;
    ParticleFrame f;
    f.pos = xyz(a.pos[pid]);
    f.dir = normalize(xyz(a.vel[pid]));
    f.pcx = axis_cell(f.pos.x, g.minx, g.lenx, g.csx); f.pcy = axis_cell(f.pos.y, g.miny, g.leny, g.csy);
    f.pcz = axis_cell(f.pos.z, g.minz, g.lenz, g.csz);
    tri_stencil_range(__float2uint_rz(__fdiv_rn(f.pos.x - g.minx, (float)g.csx)), g.nx, f.x0, f.x1);
    tri_stencil_range(__float2uint_rz(__fdiv_rn(f.pos.y - g.miny, (float)g.csy)), g.ny, f.y0, f.y1);
    tri_stencil_range(__float2uint_rz(__fdiv_rn(f.pos.z - g.minz, (float)g.csz)), g.nz, f.z0, f.z1);
    return f;
}

template <bool SLAB>
__global__ void __launch_bounds__(256) wall_filter_kernel(const VeinCollideArgs a)
{
    const WallGridDev& w = a.wall;
    const GridDev& g = a.tgrid;
    const float reach = a.phys.impactNear;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *w.dirty = 0;   // consumed by this step's rebuild; the vertex integrator raises it again if needed
    // Which particles: (a) the near-wall list left by the spring kernel's probe (+ the ghosts in slab mode) when the step
    // was enqueued with the probe - the filter then never streams the bulk of the lumen; (b) slab mode: the rank's owned
    // particles and ghosts, enumerated through its lists; (c) every particle.  Bounded grid striding over the count.
    const int nNear = w.useNearList ? *w.nearCount : 0;
    const int total = w.useNearList ? nNear + (SLAB ? *a.ghostCount : 0) : (SLAB ? item_total(a.items) : a.n);
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        int pid = -1, ghost = 0;
        if (w.useNearList) {
            if (i < nNear) pid = w.nearList[i];
            else if (i < total) { pid = a.ghostList[i - nNear]; ghost = 1; }
        } else if (SLAB) {
            if (i < total) {
                int fl = 0;
                pid = active_item(a.items, i, fl);
                ghost = fl == 2;   // ghosts only deposit their wall-force splat (owner updates the particle)
            }
        } else if (i < a.n) {
            pid = i;
        }
        unsigned pass = 0;          // bit b: cell (b%3, (b/3)%3, b/9) of the segment's box passed the slab test
        int wx0 = 0, wy0 = 0, wz0 = 0;
        bool sequential = false;
        if (pid >= 0) {
            const float4 p4 = a.pos[pid], v4 = a.vel[pid];
            const float3 pos = xyz(p4);
            // one byte answers "can this particle reach the wall at all": its own cell +-2 holds no triangle -> done
            const int hx = wall_axis(pos.x, w.ox, w.invh, w.nx), hy = wall_axis(pos.y, w.oy, w.invh, w.ny), hz = wall_axis(pos.z, w.oz, w.invh, w.nz);
            if (__ldg(w.near + (hz * w.ny + hy) * w.nx + hx)) {
                const float3 dir = normalize(xyz(v4));
                const float3 tip = pos + reach * dir;
                constexpr float EPSB = 1e-3f;
                wx0 = wall_axis(fminf(pos.x, tip.x) - EPSB, w.ox, w.invh, w.nx);
                wy0 = wall_axis(fminf(pos.y, tip.y) - EPSB, w.oy, w.invh, w.ny);
                wz0 = wall_axis(fminf(pos.z, tip.z) - EPSB, w.oz, w.invh, w.nz);
                const int wx1 = wall_axis(fmaxf(pos.x, tip.x) + EPSB, w.ox, w.invh, w.nx);
                const int wy1 = wall_axis(fmaxf(pos.y, tip.y) + EPSB, w.oy, w.invh, w.ny);
                const int wz1 = wall_axis(fmaxf(pos.z, tip.z) + EPSB, w.oz, w.invh, w.nz);
                // reach <= 2 cells: the box spans at most 3 cells per axis.  Occupancy of all of them, branch-free: one byte
                // per (y,z) row holds the bits of its three x-adjacent cells.
                unsigned occ = 0;
                const unsigned xmask = (1u << (wx1 - wx0 + 1)) - 1u;
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    const int y = wy0 + r % 3, z = wz0 + r / 3;
                    const bool in = y <= wy1 && z <= wz1;
                    const unsigned bits = __ldg(w.occ3 + (in ? (z * w.ny + y) * w.nx + wx0 : 0));
                    occ |= in ? (bits & xmask) << (3 * r) : 0u;
                }
                while (occ) {
                    const int b = __ffs(occ) - 1;
                    occ &= occ - 1;
                    const int c = ((wz0 + b / 9) * w.ny + wy0 + (b / 3) % 3) * w.nx + wx0 + b % 3;
                    const int4 r0 = __ldg(w.rec + 2 * c), r1 = __ldg(w.rec + 2 * c + 1);
                    const CellSlab sl{__int_as_float(r0.z), __int_as_float(r0.w), __int_as_float(r1.x), __int_as_float(r1.y), __int_as_float(r1.z)};
                    if (slab_segment(sl, pos, dir, reach)) pass |= 1u << b;
                }
                if (pass) {
                    const int pcx = axis_cell(pos.x, g.minx, g.lenx, g.csx), pcy = axis_cell(pos.y, g.miny, g.leny, g.csy),
                              pcz = axis_cell(pos.z, g.minz, g.lenz, g.csz);
                    sequential = pcx >= g.nx || pcy >= g.ny || pcz >= g.nz;   // outside the triangle grid
                }
            }
        }
        // warp-aggregated append: entries, then candidates
        const int mine = sequential ? 0 : __popc(pass);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned candMask = __ballot_sync(0xffffffffu, pass != 0);
        if (candMask == 0) continue;
        int ebase = 0;
        if (lane == 0 && total) ebase = atomicAdd(w.entryCount, total);
        ebase = __shfl_sync(0xffffffffu, ebase, 0);
        if (pass) {
            int at = ebase + incl - mine;
            if (!sequential && at + mine > w.entryCap) sequential = true;   // queue full: this particle searches sequentially
            if (!sequential) {
                unsigned m = pass;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    w.entries[at++] = make_int2(pid, ((wz0 + b / 9) * w.ny + wy0 + (b / 3) % 3) * w.nx + wx0 + b % 3);
                }
            }
            w.best[pid] = sequential ? SEQUENTIAL : NO_HIT;
            w.ghostFlag[pid] = (unsigned char)ghost;
            if (sequential) w.queue[atomicAdd(w.queueCount, 1)] = pid;   // rare: straight to phase B
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(256) wall_triangles_kernel(const VeinCollideArgs a)
{
    const WallGridDev& w = a.wall;
    const float reach = a.phys.impactNear;
    const int n = min(*w.entryCount, w.entryCap);
    const int sub = threadIdx.x & 7;
    const int groups = (gridDim.x * blockDim.x) >> 3;
    unsigned long long myTests = 0;
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; e < n; e += groups) {
        const int2 it = w.entries[e];
        const int pid = it.x;
        if (w.best[pid] == SEQUENTIAL) continue;
        const int4 r0 = __ldg(w.rec + 2 * it.y);
        const ParticleFrame f = particle_frame(a, pid);
        for (int k = sub; k < r0.y; k += 8) {
            const int slot = __ldg(w.list + r0.x + k);
            const int4 si = __ldg(w.slotInfo + slot);
            const int dx = si.y - f.pcx, dy = si.z - f.pcy, dz = si.w - f.pcz;
            if (dx < f.x0 || dx > f.x1 || dy < f.y0 || dy > f.y1 || dz < f.z0 || dz > f.z1) continue;   // outside this particle's stencil
            const unsigned long long cand = ((unsigned long long)(((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1)) << 32) | (unsigned)slot;
            if (cand >= w.best[pid]) continue;   // a filter only (racy read): the atomicMin decides
            if (STATS) ++myTests;
            RayHit h;
            if (ray_triangle(f.pos, f.dir, load_tri(a, slot), h) && h.t <= reach) {
                // whoever replaces NO_HIT is the particle's first near hit to land: it enters the phase-B list exactly once
                if (atomicMin(w.best + pid, cand) == NO_HIT) w.queue[atomicAdd(w.queueCount, 1)] = pid;
            }
        }
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if ((threadIdx.x & 31) == 0 && myTests) atomicAdd(&a.counters->triTests, myTests);
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase B: a warp per particle with a near hit (listed by A2 when its first near hit landed) - is it masked by an earlier (far) hit?
// The (stencil cell, slot group) pairs up to the near hit are flattened over the lanes, so the box tests and the
// triangle tests run 32 wide instead of cell after cell (the search is a chain of dependent L2 reads).
// The check reads positions, velocities and the wall only, so it belongs to the SEARCH (beside the grid build and
// the particle collisions); a masked entry is struck from the list (queue[q] = ~pid).  What has to wait for the
// particle collisions - the stage's effect - is wall_apply_kernel below: a thread per surviving entry.
// ------------------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) wall_masking_kernel(const VeinCollideArgs a)
{
    const WallGridDev& w = a.wall;
    const GridDev& g = a.tgrid;
    const int lane = threadIdx.x & 31;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const int plane = g.nx * g.ny;
    const int n = *w.queueCount;
    unsigned long long myTests = 0;
    for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += nWarps) {
        {
            const int pid = w.queue[q];
            const unsigned long long best = w.best[pid];
            if (best == SEQUENTIAL) continue;   // searched (and applied) by wall_apply_kernel
            const ParticleFrame f = particle_frame(a, pid);
            const int bestKey = (int)(best >> 32), bestSlot = (int)(best & 0xffffffffu);
            const int cell = (f.pcz * g.ny + f.pcy) * g.nx + f.pcx;
            // lane == traversal key of one stencil cell (x outer, y, z inner): its slot-group range up to the near hit
            const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
            const int c = cell + dz * plane + dy * g.nx + dx;
            int s = 0, e = -1;
            if (lane < 27 && lane <= bestKey && dx >= f.x0 && dx <= f.x1 && dy >= f.y0 && dy <= f.y1 && dz >= f.z0 && dz <= f.z1 && c >= 0 && c < g.cells) {
                s = a.cellStart[c];
                e = (lane == bestKey) ? bestSlot - 1 : a.cellEnd[c];
                if (e >= s && !ray_box(w.cellBox[c], f.pos, f.dir)) e = s - 1;
            }
            const int g0 = s >> 3, ng = e >= s ? (e >> 3) - g0 + 1 : 0;
            int incl = ng;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const int G = __shfl_sync(0xffffffffu, incl, 31);
            bool masked = false;
            for (int fb = 0; fb < G && !masked; fb += 32) {
                // flattened group index -> owning stencil cell (lane L with incl[L-1] <= fi < incl[L])
                const int fi = fb + lane;
                int L = 0;
#pragma unroll
                for (int l = 0; l < 27; ++l) L += __shfl_sync(0xffffffffu, incl, l) <= fi;
                const int Ls = min(L, 26);
                const int os = __shfl_sync(0xffffffffu, s, Ls), oe = __shfl_sync(0xffffffffu, e, Ls);
                const int og0 = __shfl_sync(0xffffffffu, g0, Ls), oincl = __shfl_sync(0xffffffffu, incl, Ls), ong = __shfl_sync(0xffffffffu, ng, Ls);
                const int gi = og0 + (fi - (oincl - ong));
                const bool ok = fi < G && ray_box(w.groupBox[gi], f.pos, f.dir);
                unsigned gm = __ballot_sync(0xffffffffu, ok);
                while (gm && !masked) {
                    // four slot groups (32 triangles) per round
                    int from = -1;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int sel = gm ? __ffs(gm) - 1 : -1;
                        gm &= gm - 1;
                        if ((lane >> 3) == r) from = sel;
                    }
                    const int fromS = max(from, 0);
                    const int mg = __shfl_sync(0xffffffffu, gi, fromS), ms = __shfl_sync(0xffffffffu, os, fromS), me = __shfl_sync(0xffffffffu, oe, fromS);
                    bool hitFar = false;
                    if (from >= 0) {
                        const int slot = (mg << 3) + (lane & 7);
                        if (slot >= ms && slot <= me) {
                            if (STATS) ++myTests;
                            RayHit far;
                            hitFar = ray_triangle(f.pos, f.dir, load_tri(a, slot), far);
                        }
                    }
                    masked = __any_sync(0xffffffffu, hitFar);
                }
            }
            if (masked && lane == 0) w.queue[q] = ~pid;
        }
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if (lane == 0 && myTests) atomicAdd(&a.counters->triTests, myTests);
    }
}

// the stage's effect (vein_collisions.cu:234-276) for every listed particle whose near hit was not masked: reaction
// force, velocity reflection, wall-force splats.  Particles the grid search could not take (outside the triangle grid,
// queue overflow) run the sequential search here.
template <bool STATS>
__global__ void __launch_bounds__(128) wall_apply_kernel(const VeinCollideArgs a)
{
    const WallGridDev& w = a.wall;
    const int n = *w.queueCount;
    unsigned long long myTests = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int pid = w.queue[q];
        if (pid < 0) continue;   // masked
        const unsigned long long best = w.best[pid];
        const bool splatOnly = w.ghostFlag[pid] != 0;
        if (best == SEQUENTIAL) {
            vein_collide_particle<true, STATS>(a, pid, myTests, splatOnly);
            continue;
        }
        const float4 p4 = a.pos[pid], v4 = a.vel[pid];
        const float3 dir = normalize(xyz(v4));
        RayHit h;
        ray_triangle(xyz(p4), dir, load_tri(a, (int)(best & 0xffffffffu)), h);
        vein_apply_hit(a, pid, p4, v4, dir, h, splatOnly);
    }
    if (STATS && myTests) atomicAdd(&a.counters->triTests, myTests);
}

}  // namespace

void launch_wall_slot_info(const int* sortedTriKeys, const int* triIds, const unsigned* vidx, int T, GridDev tgrid, int4* slotInfo, int4* slotVerts,
                           cudaStream_t st)
{
    BCS_LAUNCH("wall_slot_info", st, wall_slot_info_kernel<<<(T + 255) / 256, 256, 0, st>>>(sortedTriKeys, triIds, vidx, T, tgrid, slotInfo, slotVerts));
    BCS_CUDA(cudaGetLastError());
}

void launch_wall_rebuild(const VeinCollideArgs& a, int V, int numSMs, cudaStream_t st)
{
    // The rebuild synchronises its CTAs with a spinning grid barrier, so ALL of them must be able to be resident at once:
    // one CTA per SM, checked against what the device can actually hold (a register count that grew with a compiler
    // change, or an SM partition, would otherwise make the first real rebuild spin for ever).  CTAs of other kernels
    // running beside it in the forked graph retire on their own; only this kernel's own grid has to fit.
    {
        int perSM = 0;
        BCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, wall_rebuild_kernel, REBUILD_THREADS, 0));
        BCS_REQUIRE(perSM >= 1, BCS_ERR_UNSUPPORTED, "the wall-grid rebuild kernel does not fit an SM on this device");
    }
    BCS_LAUNCH("wall_rebuild", st,
               wall_rebuild_kernel<<<numSMs, REBUILD_THREADS, 0, st>>>(a.wall, a.tgrid, V, a.T, a.vpos, a.vidx, a.triIds, a.cellStart, a.cellEnd,
                                                                        a.groupLocal, a.triCellLocal));
    BCS_CUDA(cudaGetLastError());
}

static VeinCollideArgs wall_args(const VeinCollideArgs& a0)
{
    VeinCollideArgs a = a0;
    a.liveTris = 1;                 // nothing is repacked per step on this path
    a.cellBox = a.wall.cellBox;     // the sequential fallback (particles outside the triangle grid) culls with the lazy boxes
    a.groupBox = a.wall.groupBox;
    return a;
}

void launch_wall_reset(const VeinCollideArgs& a, cudaStream_t st)
{
    BCS_CUDA(cudaMemsetAsync(a.wall.queueCount, 0, 3 * sizeof(int), st));
}

// phase A (filter + triangle tests): reads particle positions / velocities and the wall only - may run beside the
// spring and particle-collision kernels
void launch_wall_search(const VeinCollideArgs& a0, cudaStream_t st)
{
    const VeinCollideArgs a = wall_args(a0);
    // (with the near list the three counters were cleared before the spring kernel filled the list: launch_wall_reset)
    if (!a.wall.useNearList) BCS_CUDA(cudaMemsetAsync(a.wall.queueCount, 0, 3 * sizeof(int), st));   // queueCount, entryCount, nearCount (adjacent)
    const int blocks = (a.n + 255) / 256;
    if (a.pflag) BCS_LAUNCH("vein_filter", st, wall_filter_kernel<true><<<std::min(blocks, BOUNDED_BLOCKS), 256, 0, st>>>(a));
    else BCS_LAUNCH("vein_filter", st, wall_filter_kernel<false><<<a.wall.useNearList ? std::min(blocks, BOUNDED_BLOCKS) : blocks, 256, 0, st>>>(a));
    if (a.stats) BCS_LAUNCH("vein_collisions", st, wall_triangles_kernel<true><<<148 * 8, 256, 0, st>>>(a));
    else BCS_LAUNCH("vein_collisions", st, wall_triangles_kernel<false><<<148 * 8, 256, 0, st>>>(a));
    if (a.stats) BCS_LAUNCH("vein_masking", st, wall_masking_kernel<true><<<148 * 8, 128, 0, st>>>(a));
    else BCS_LAUNCH("vein_masking", st, wall_masking_kernel<false><<<148 * 8, 128, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

// the stage's effect on particle force / velocity and the wall-force splats: after the particle collisions (it reads
// the accumulated force) and the vein spring gather (both add to the vertex forces)
void launch_wall_apply(const VeinCollideArgs& a0, cudaStream_t st)
{
    const VeinCollideArgs a = wall_args(a0);
    if (a.stats) BCS_LAUNCH("vein_apply", st, wall_apply_kernel<true><<<148, 128, 0, st>>>(a));
    else BCS_LAUNCH("vein_apply", st, wall_apply_kernel<false><<<148, 128, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

void launch_wall_collisions(const VeinCollideArgs& a, cudaStream_t st)
{
    launch_wall_search(a, st);
    launch_wall_apply(a, st);
}

}  // namespace bcs

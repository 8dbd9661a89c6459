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
#include "vein_device.cuh"

#include <cstdlib>

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
// Mesh vertices are numbered ring by ring (config/vein_definition.hpp:12 and the generated cylinders alike), so the nine
// neighbours of a vertex lie within ~a ring of it in the arrays.  A CTA therefore stages a WINDOW of positions and
// velocities around its 256 vertices in shared memory with coalesced loads and gathers from there; a neighbour outside
// the window (any mesh is legal) is read from global memory.
constexpr int VG_THREADS = 256;
constexpr int VG_HALO = 128;
constexpr int VG_WINDOW = VG_THREADS + 2 * VG_HALO;

__global__ void __launch_bounds__(VG_THREADS) vein_gather_kernel(const VeinArgs a)
{
    __shared__ float4 wp[VG_WINDOW], wv[VG_WINDOW];
    const int base = a.vFirst + blockIdx.x * VG_THREADS - VG_HALO;
    for (int k = threadIdx.x; k < VG_WINDOW; k += VG_THREADS) {
        const int v = base + k;
        if (v >= 0 && v < a.V) { wp[k] = a.vpos[v]; wv[k] = a.vvel[v]; }
    }
    __syncthreads();
    const int id = a.vFirst + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.vFirst + a.vCount) return;
    if (a.vOwned && !a.vOwned[id]) return;   // slab mode: only vertices of this rank's slab
    const float3 p = xyz(wp[threadIdx.x + VG_HALO]), v = xyz(wv[threadIdx.x + VG_HALO]);
    float3 F = f3(0.f, 0.f, 0.f);
    // all eighteen table loads first: they are independent, and a vertex has nothing else to do while they fly
    int nbr[BCS_VEIN_MAX_NEIGHBORS];
    float len0[BCS_VEIN_MAX_NEIGHBORS];
#pragma unroll
    for (int s = 0; s < BCS_VEIN_MAX_NEIGHBORS; ++s) {
        nbr[s] = __ldg(a.nbrIds + (size_t)s * a.V + id);
        len0[s] = __ldg(a.nbrLen + (size_t)s * a.V + id);
    }
#pragma unroll
    for (int s = 0; s < BCS_VEIN_MAX_NEIGHBORS; ++s) {
        // branch-free so that the nine neighbour gathers are in flight together: an absent slot (-1) reads the
        // vertex itself, whose zero separation normalises to the zero vector and contributes exactly +0
        const int nb = nbr[s] < 0 ? id : nbr[s];
        const float L = len0[s];
        const unsigned w = (unsigned)(nb - base);
        float4 q4, qv4;
        if (w < (unsigned)VG_WINDOW) { q4 = wp[w]; qv4 = wv[w]; }
        else { q4 = a.vpos[nb]; qv4 = a.vvel[nb]; }
        const float3 q = xyz(q4);
        // q - p == -(p - q) exactly, so normalize(q-p) == -normalize(p-q).  length() and normalize() share one reciprocal
        // square root refined to <= 1 ulp (as in the blood-cell spring kernel): the IEEE sqrt + three IEEE divisions per
        // neighbour were ~60 % of this kernel's instructions.  Deviation <= 2 ulp per component.
        const float3 d = p - q;
        const float d2 = dot(d, d);
        float inv = rsqrtf(d2);
        float len = d2 * inv;
        len = fmaf(fmaf(-len, len, d2), 0.5f * inv, len);
        inv = fmaf(fmaf(-len, inv, 1.0f), inv, inv);
        if (!(d2 > 0.f)) { inv = 0.f; len = 0.f; }   // absent slot / coincident vertices: normalize() yields the zero vector
        const float3 n = f3(d.x * inv, d.y * inv, d.z * inv);
        const float sf = (len - L) * a.phys.vein_k_sniff + dot(n, (v - xyz(qv4))) * a.phys.vein_d_fact;
        F = F + sf * f3(-n.x, -n.y, -n.z);
    }
    float4 f = a.vfrc[id];
    f.x += F.x; f.y += F.y; f.z += F.z;
    a.vfrc[id] = f;
}

void launch_vein_gather(const VeinArgs& a, cudaStream_t st)
{
    if (a.vCount <= 0) return;
    BCS_LAUNCH("vein_gather", st, vein_gather_kernel<<<(a.vCount + VG_THREADS - 1) / VG_THREADS, VG_THREADS, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

// v += dt*F; x += dt*v; F = 0   (semi-implicit Euler + the three cudaMemsets of the reference, fused)
__global__ void __launch_bounds__(256) vein_integrate_kernel(const VeinArgs a)
{
    const int id = a.vFirst + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.vFirst + a.vCount) return;
    if (a.vOwned && !a.vOwned[id]) {
        // slab mode: not ours - state arrives with the vertex halo; forget the partial splats we accumulated on it
        if (a.vfrc[id].w != 0.f) {
            splat_take(a.vsplat, id);
            a.vfrc[id] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    float4 x = a.vpos[id], v = a.vvel[id];
    float4 F = a.vfrc[id];
    if (F.w != 0.f) {
        const float3 sp = splat_take(a.vsplat, id);
        F.x += sp.x; F.y += sp.y; F.z += sp.z;
    }
    const float dt = a.phys.dt;
    v.x += dt * F.x; v.y += dt * F.y; v.z += dt * F.z;
    x.x += dt * v.x; x.y += dt * v.y; x.z += dt * v.z;
    a.vvel[id] = v;
    a.vpos[id] = x;
    a.vfrc[id] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.vposBuilt) {
        // wall grid (wall.cu): a vertex that left its margin invalidates the padded structure
        const float4 b = a.vposBuilt[id];
        const float dx = x.x - b.x, dy = x.y - b.y, dz = x.z - b.z;
        if (dx * dx + dy * dy + dz * dz > a.wallMargin * a.wallMargin) *a.wallDirty = 1;
    }
}

// vfrc += parked wall splats (idempotent): run before vfrc is read back between the collision stage and the integrator
__global__ void __launch_bounds__(256) vein_fold_splats_kernel(const VeinArgs a)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= a.V) return;
    float4 F = a.vfrc[id];
    if (F.w == 0.f) return;
    const float3 sp = splat_take(a.vsplat, id);
    a.vfrc[id] = make_float4(F.x + sp.x, F.y + sp.y, F.z + sp.z, 0.f);
}

void launch_vein_fold_splats(const VeinArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("vein_fold_splats", st, vein_fold_splats_kernel<<<(a.V + 255) / 256, 256, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

void launch_vein_integrate(const VeinArgs& a, cudaStream_t st)
{
    if (a.vCount <= 0) return;
    BCS_LAUNCH("vein_integrate", st, vein_integrate_kernel<<<(a.vCount + 255) / 256, 256, 0, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Triangles are re-packed every step in sorted-slot order (the wall moves): v0, e1 = v1-v0, e2 = v2-v0.
// The Moeller-Trumbore test only ever uses these three vectors, so the per-test gathers of the reference
// (3 index loads + 9 coordinate loads through two indirections) become three aligned float4 loads.
// The same pass refits the culling hierarchy: one AABB per group of 8 consecutive sorted slots.

__global__ void __launch_bounds__(256) tri_refit_kernel(const float4* __restrict__ vpos, const unsigned* __restrict__ vidx,
                                                        const int* __restrict__ triIds, int T, TriPacked* __restrict__ out,
                                                        Aabb* __restrict__ groupBox, CellSlab* __restrict__ groupSlab,
                                                        const unsigned char* __restrict__ groupLocal)
{
    // slab mode: only slot groups near this rank's slab are refitted (whole groups of 8 lanes leave together)
    if (groupLocal && (int)(blockIdx.x * blockDim.x + threadIdx.x) < T && !groupLocal[(blockIdx.x * blockDim.x + threadIdx.x) >> 3]) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    float lox = 3e38f, loy = 3e38f, loz = 3e38f, hix = -3e38f, hiy = -3e38f, hiz = -3e38f;
    float3 nrm = f3(0.f, 0.f, 0.f), c0 = nrm, c1 = nrm, c2 = nrm;
    if (s < T) {
        const int tri = triIds[s];
        const float3 v0 = xyz(vpos[vidx[3 * tri]]), v1 = xyz(vpos[vidx[3 * tri + 1]]), v2 = xyz(vpos[vidx[3 * tri + 2]]);
        const float3 e1 = v1 - v0, e2 = v2 - v0;
        TriPacked p;
        p.a = make_float4(v0.x, v0.y, v0.z, e1.x);
        p.b = make_float4(e1.y, e1.z, e2.x, e2.y);
        p.c = make_float4(e2.z, __int_as_float(tri), 0.f, 0.f);
        out[s] = p;
        lox = fminf(v0.x, fminf(v1.x, v2.x)); hix = fmaxf(v0.x, fmaxf(v1.x, v2.x));
        loy = fminf(v0.y, fminf(v1.y, v2.y)); hiy = fmaxf(v0.y, fmaxf(v1.y, v2.y));
        loz = fminf(v0.z, fminf(v1.z, v2.z)); hiz = fmaxf(v0.z, fmaxf(v1.z, v2.z));
        nrm = cross(e1, e2);   // area-weighted normal
        c0 = v0; c1 = v1; c2 = v2;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    // slab of the group along its mean normal (a wall patch of 8 triangles is nearly planar)
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        nrm.x += __shfl_xor_sync(0xffffffffu, nrm.x, o); nrm.y += __shfl_xor_sync(0xffffffffu, nrm.y, o);
        nrm.z += __shfl_xor_sync(0xffffffffu, nrm.z, o);
    }
    float3 n = normalize(nrm);
    if (n.x == 0.f && n.y == 0.f && n.z == 0.f) n = f3(1.f, 0.f, 0.f);
    float dmin = 3e38f, dmax = -3e38f;
    if (s < T) {
        const float d0 = dot(n, c0), d1 = dot(n, c1), d2 = dot(n, c2);
        dmin = fminf(d0, fminf(d1, d2)); dmax = fmaxf(d0, fmaxf(d1, d2));
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if ((threadIdx.x & 7) == 0 && s < T) {
        groupBox[s >> 3] = Aabb{lox - BOX_PAD, loy - BOX_PAD, loz - BOX_PAD, hix + BOX_PAD, hiy + BOX_PAD, hiz + BOX_PAD};
        groupSlab[s >> 3] = CellSlab{n.x, n.y, n.z, dmin - BOX_PAD, dmax + BOX_PAD};
    }
}

// Bounds of everything a cell's table range [start,end] can reach: an AABB (union of the slot groups the range
// overlaps - a superset, which is all the culling needs) and a SLAB along the mean normal of the cell's
// triangles.  The wall patch of a 25-unit cell is nearly planar, so the slab is a few units thick where the
// AABB of a diagonal patch reaches far into the lumen; together they reject almost every particle that is
// not genuinely within reach of the wall.  Works for both table semantics, including the stale /
// zero-initialised ranges of the reference-compatible mode.  One warp per cell.
__global__ void __launch_bounds__(128) cell_box_kernel(const int* __restrict__ cellStart, const int* __restrict__ cellEnd, int cells,
                                                       const Aabb* __restrict__ groupBox, const TriPacked* __restrict__ tris,
                                                       Aabb* __restrict__ cellBox, CellSlab* __restrict__ cellSlab,
                                                       const unsigned char* __restrict__ cellLocal)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= cells) return;
    if (cellLocal && !cellLocal[c]) return;
    const int s = cellStart[c], e = cellEnd[c];
    Aabb b{3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
    float3 nsum = f3(0.f, 0.f, 0.f);
    if (e >= s) {
        for (int g = (s >> 3) + lane; g <= (e >> 3); g += 32) {
            const Aabb q = groupBox[g];
            b.lox = fminf(b.lox, q.lox); b.loy = fminf(b.loy, q.loy); b.loz = fminf(b.loz, q.loz);
            b.hix = fmaxf(b.hix, q.hix); b.hiy = fmaxf(b.hiy, q.hiy); b.hiz = fmaxf(b.hiz, q.hiz);
        }
        for (int i = s + lane; i <= e; i += 32) {
            const TriPacked tp = tris[i];
            nsum = nsum + cross(f3(tp.a.w, tp.b.x, tp.b.y), f3(tp.b.z, tp.b.w, tp.c.x));   // area-weighted normal
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        b.lox = fminf(b.lox, __shfl_xor_sync(0xffffffffu, b.lox, o)); b.hix = fmaxf(b.hix, __shfl_xor_sync(0xffffffffu, b.hix, o));
        b.loy = fminf(b.loy, __shfl_xor_sync(0xffffffffu, b.loy, o)); b.hiy = fmaxf(b.hiy, __shfl_xor_sync(0xffffffffu, b.hiy, o));
        b.loz = fminf(b.loz, __shfl_xor_sync(0xffffffffu, b.loz, o)); b.hiz = fmaxf(b.hiz, __shfl_xor_sync(0xffffffffu, b.hiz, o));
        nsum.x += __shfl_xor_sync(0xffffffffu, nsum.x, o); nsum.y += __shfl_xor_sync(0xffffffffu, nsum.y, o);
        nsum.z += __shfl_xor_sync(0xffffffffu, nsum.z, o);
    }
    float3 n = normalize(nsum);
    if (n.x == 0.f && n.y == 0.f && n.z == 0.f) n = f3(1.f, 0.f, 0.f);   // degenerate: any unit axis keeps the slab valid
    float dmin = 3e38f, dmax = -3e38f;
    if (e >= s) {
        for (int i = s + lane; i <= e; i += 32) {
            const TriPacked tp = tris[i];
            const float3 v0 = f3(tp.a.x, tp.a.y, tp.a.z);
            const float d0 = dot(n, v0), d1 = dot(n, v0 + f3(tp.a.w, tp.b.x, tp.b.y)), d2 = dot(n, v0 + f3(tp.b.z, tp.b.w, tp.c.x));
            dmin = fminf(dmin, fminf(d0, fminf(d1, d2)));
            dmax = fmaxf(dmax, fmaxf(d0, fmaxf(d1, d2)));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0) {
        cellBox[c] = b;
        cellSlab[c] = CellSlab{n.x, n.y, n.z, dmin - BOX_PAD, dmax + BOX_PAD};
    }
}

void launch_tri_refit(const VeinCollideArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("tri_refit", st, tri_refit_kernel<<<(a.T + 255) / 256, 256, 0, st>>>(a.vpos, a.vidx, a.triIds, a.T, a.tris, a.groupBox, a.groupSlab, a.groupLocal));
    BCS_LAUNCH("cell_box", st,
               cell_box_kernel<<<(a.tgrid.cells * 32 + 127) / 128, 128, 0, st>>>(a.cellStart, a.cellEnd, a.tgrid.cells, a.groupBox, a.tris,
                                                                                   a.cellBox, a.cellSlab, a.triCellLocal));
    BCS_CUDA(cudaGetLastError());
}


// every particle (exhaustive cross-check mode and the debug view)
template <bool FAST, bool STATS>
__global__ void __launch_bounds__(128) vein_collisions_kernel(const VeinCollideArgs a)
{
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long myTests = 0;
    if (pid < a.n) vein_collide_particle<FAST, STATS>(a, pid, myTests);
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if ((threadIdx.x & 31) == 0 && a.apply) atomicAdd(&a.counters->triTests, myTests);
    }
}


// Production path, step 1: one thread per BLOOD CELL.  The stage can only act on a particle that has a wall
// triangle within veinImpactDistance along its ray, so a blood cell is skipped as a whole unless its bounding
// box, widened by that reach, meets the box AND the slab of one of the triangle-grid cells its particles can
// visit.  Survivors are appended to a work list (order irrelevant: particles are independent) together with
// the bit set of those triangle cells.
__global__ void __launch_bounds__(128) vein_cull_cells_kernel(const VeinCollideArgs a, int nCells, CullEntry* __restrict__ list,
                                                             int* __restrict__ listCount)
{
    // one WARP per blood cell: lanes = particles for the bounding box, then lanes = candidate triangle cells
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int c = w, t = 0;
    if (a.lists.cells) {
        // slab mode: the w-th owned blood cell
        if (w >= a.lists.cellPrefix[a.types.n]) return;
        while (t + 1 < a.types.n && w >= a.lists.cellPrefix[t + 1]) ++t;
        c = a.lists.cells[a.lists.typeFirst[t] + (w - a.lists.cellPrefix[t])];
    } else {
        if (c >= nCells) return;
        while (t + 1 < a.types.n && c >= a.types.t[t + 1].cStart) ++t;
    }
    const TypeDev ty = a.types.t[t];
    const int first = ty.pStart + (c - ty.cStart) * ty.P;
    float lox = 3e38f, loy = 3e38f, loz = 3e38f, hix = -3e38f, hiy = -3e38f, hiz = -3e38f;
    for (int k = lane; k < ty.P; k += 32) {
        const float4 p = a.pos[first + k];
        lox = fminf(lox, p.x); hix = fmaxf(hix, p.x);
        loy = fminf(loy, p.y); hiy = fmaxf(hiy, p.y);
        loz = fminf(loz, p.z); hiz = fmaxf(hiz, p.z);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    const GridDev& g = a.tgrid;
    // triangle-grid cells any particle of the blood cell can visit: its own cell +-1 per axis (superset of
    // the trimmed stencils of vein_collisions.cu:86-230)
    const int cx0 = max(0, axis_cell(lox, g.minx, g.lenx, g.csx) - 1), cx1 = min(g.nx - 1, axis_cell(hix, g.minx, g.lenx, g.csx) + 1);
    const int cy0 = max(0, axis_cell(loy, g.miny, g.leny, g.csy) - 1), cy1 = min(g.ny - 1, axis_cell(hiy, g.miny, g.leny, g.csy) + 1);
    const int cz0 = max(0, axis_cell(loz, g.minz, g.lenz, g.csz) - 1), cz1 = min(g.nz - 1, axis_cell(hiz, g.minz, g.lenz, g.csz) + 1);
    CullEntry ent{c, cx0, cy0, cz0, 0ull};
    if (cx1 - cx0 > 3 || cy1 - cy0 > 3 || cz1 - cz0 > 3) {
        // stretched blood cell (> 4 triangle cells per axis): no cell-level culling, its particles search their full stencil
        ent.mask = ~0ull;
        ent.cx0 = -1;
    } else {
        const float r = a.phys.impactNear;
        const float3 ctr = f3(0.5f * (lox + hix), 0.5f * (loy + hiy), 0.5f * (loz + hiz));
        const float rad = 0.5f * sqrtf((hix - lox) * (hix - lox) + (hiy - loy) * (hiy - loy) + (hiz - loz) * (hiz - loz)) + r;
        lox -= r; loy -= r; loz -= r; hix += r; hiy += r; hiz += r;
        unsigned half[2];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            const int bit = lane + 32 * hb;
            const int x = cx0 + (bit & 3), y = cy0 + ((bit >> 2) & 3), z = cz0 + (bit >> 4);
            bool pass = x <= cx1 && y <= cy1 && z <= cz1;
            if (pass) {
                const int tc = (z * g.ny + y) * g.nx + x;
                pass = box_overlap(a.cellBox[tc], lox, loy, loz, hix, hiy, hiz) && slab_near(a.cellSlab[tc], ctr, rad);
            }
            half[hb] = __ballot_sync(0xffffffffu, pass);
        }
        ent.mask = (unsigned long long)half[0] | ((unsigned long long)half[1] << 32);
    }
    if (lane == 0 && ent.mask != 0ull) list[atomicAdd(listCount, 1)] = ent;
}

// slab mode: ghost particles (owned by a neighbouring rank) enter the same work list as single-particle entries
// (cell = -(particle id + 1)); they only deposit their wall-force splat, so that every rank sees all contributions
// to the vein vertices it integrates, while the particle itself is updated by its owner.
__global__ void __launch_bounds__(128) vein_cull_ghosts_kernel(const VeinCollideArgs a, CullEntry* __restrict__ list, int* __restrict__ listCount)
{
    // one warp per ghost: lane = one of the 27 triangle cells around it
    const int n = *a.ghostCount;
    const GridDev& g = a.tgrid;
    const float r = a.phys.impactNear;
    const int lane = threadIdx.x & 31, warpsTotal = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warpsTotal) {
        const int pid = a.ghostList[k];
        const float4 p = a.pos[pid];
        const int pcx = axis_cell(p.x, g.minx, g.lenx, g.csx), pcy = axis_cell(p.y, g.miny, g.leny, g.csy), pcz = axis_cell(p.z, g.minz, g.lenz, g.csz);
        if (pcx >= g.nx || pcy >= g.ny || pcz >= g.nz) continue;
        const int cx0 = max(0, pcx - 1), cy0 = max(0, pcy - 1), cz0 = max(0, pcz - 1);
        const int x = cx0 + lane % 3, y = cy0 + (lane / 3) % 3, z = cz0 + lane / 9;
        bool pass = lane < 27 && x <= min(g.nx - 1, pcx + 1) && y <= min(g.ny - 1, pcy + 1) && z <= min(g.nz - 1, pcz + 1);
        if (pass) {
            const int tc = (z * g.ny + y) * g.nx + x;
            pass = box_overlap(a.cellBox[tc], p.x - r, p.y - r, p.z - r, p.x + r, p.y + r, p.z + r) && slab_near(a.cellSlab[tc], xyz(p), r);
        }
        // bit layout of CullEntry::mask: ((dz*4 + dy)*4 + dx)
        unsigned long long mine = pass ? 1ull << (((z - cz0) * 4 + (y - cy0)) * 4 + (x - cx0)) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) list[atomicAdd(listCount, 1)] = CullEntry{-(pid + 1), cx0, cy0, cz0, mine};
    }
}

// Phase A restricted to the triangle cells the blood-cell cull marked (any visiting order; the winner is the
// near hit with the smallest (traversal ordinal, slot)), then the unchanged phase B of first_hit_fast.
template <bool STATS>
__device__ bool first_hit_marked(const VeinCollideArgs& a, const CullEntry& ent, const float3 pos, const float3 dir, int pcx, int pcy,
                                 int pcz, int x0, int x1, int y0, int y1, int z0, int z1, RayHit& h, unsigned long long& tests)
{
    const GridDev& g = a.tgrid;
    const float reach = a.phys.impactNear;
    const float3 tip = pos + reach * dir;
    const float slx = fminf(pos.x, tip.x), shx = fmaxf(pos.x, tip.x);
    const float sly = fminf(pos.y, tip.y), shy = fmaxf(pos.y, tip.y);
    const float slz = fminf(pos.z, tip.z), shz = fmaxf(pos.z, tip.z);
    int bestKey = 1 << 30, bestSlot = -1;
    unsigned long long m = ent.mask;
    while (m) {
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int dx = ent.cx0 + (bit & 3) - pcx, dy = ent.cy0 + ((bit >> 2) & 3) - pcy, dz = ent.cz0 + (bit >> 4) - pcz;
        if (dx < x0 || dx > x1 || dy < y0 || dy > y1 || dz < z0 || dz > z1) continue;   // outside this particle's stencil
        const int key = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);                      // x outer, y, z inner
        if (key > bestKey) continue;
        const int c = ((pcz + dz) * g.ny + (pcy + dy)) * g.nx + (pcx + dx);
        if (!box_overlap(a.cellBox[c], slx, sly, slz, shx, shy, shz) || !slab_near(a.cellSlab[c], pos, reach)) continue;
        const int s = a.cellStart[c], e = a.cellEnd[c];
        bool found = false;
        for (int gi = s >> 3; gi <= (e >> 3) && !found; ++gi) {
            if (!box_overlap(a.groupBox[gi], slx, sly, slz, shx, shy, shz)) continue;
            const int i0 = max(s, gi << 3), i1 = min(e, (gi << 3) + 7);
            for (int i = i0; i <= i1; ++i) {
                if (STATS) ++tests;
                RayHit cand;
                if (ray_triangle(pos, dir, load_tri(a, i), cand) && cand.t <= reach) {
                    h = cand; bestKey = key; bestSlot = i; found = true;
                    break;
                }
            }
        }
    }
    if (bestSlot < 0) return false;
    // phase B: any hit (necessarily beyond reach) earlier in traversal order?
    const int plane = g.nx * g.ny;
    const int cell = (pcz * g.ny + pcy) * g.nx + pcx;
    for (int x = x0; x <= x1; ++x)
        for (int y = y0; y <= y1; ++y)
            for (int z = z0; z <= z1; ++z) {
                const int key = ((x + 1) * 3 + (y + 1)) * 3 + (z + 1);
                if (key > bestKey) return true;
                const int c = cell + z * plane + y * g.nx + x;
                if (c < 0 || c >= g.cells) continue;
                if (!ray_box(a.cellBox[c], pos, dir)) continue;
                const int s = a.cellStart[c];
                const int e = (key == bestKey) ? bestSlot - 1 : a.cellEnd[c];
                if (e < s) continue;
                for (int gi = s >> 3; gi <= (e >> 3); ++gi) {
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

// step 2: the particles of the listed blood cells (grid-stride over list entries x particles per cell)
template <bool STATS>
__global__ void __launch_bounds__(128) vein_collisions_listed_kernel(const VeinCollideArgs a, const CullEntry* __restrict__ list,
                                                                    const int* __restrict__ listCount, int maxP)
{
    unsigned long long myTests = 0;
    const long long items = (long long)(*listCount) * maxP;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += (long long)gridDim.x * blockDim.x) {
        const CullEntry ent = list[w / maxP];
        const int c = ent.cell, k = (int)(w % maxP);
        int t = 0;
        while (t + 1 < a.types.n && c >= a.types.t[t + 1].cStart) ++t;
        if (k >= a.types.t[t].P) continue;
        const int pid = a.types.t[t].pStart + (c - a.types.t[t].cStart) * a.types.t[t].P + k;
        if (ent.cx0 < 0) { vein_collide_particle<true, STATS>(a, pid, myTests); continue; }
        const GridDev& g = a.tgrid;
        const float4 p4 = a.pos[pid], v4 = a.vel[pid];
        const float3 pos = xyz(p4), velocity = xyz(v4);
        const float3 dir = normalize(velocity);
        const int pcx = axis_cell(pos.x, g.minx, g.lenx, g.csx), pcy = axis_cell(pos.y, g.miny, g.leny, g.csy),
                  pcz = axis_cell(pos.z, g.minz, g.lenz, g.csz);
        if (pcx >= g.nx || pcy >= g.ny || pcz >= g.nz) { vein_collide_particle<true, STATS>(a, pid, myTests); continue; }
        int x0, x1, y0, y1, z0, z1;
        tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.x - g.minx, (float)g.csx)), g.nx, x0, x1);
        tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.y - g.miny, (float)g.csy)), g.ny, y0, y1);
        tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.z - g.minz, (float)g.csz)), g.nz, z0, z1);
        RayHit h;
        h.t = 1e10f; h.normal = f3(0.f, 0.f, 0.f); h.refl = f3(0.f, 0.f, 0.f); h.tri = 0;
        if (first_hit_marked<STATS>(a, ent, pos, dir, pcx, pcy, pcz, x0, x1, y0, y1, z0, z1, h, myTests))
            vein_apply_hit(a, pid, p4, v4, dir, h);
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if ((threadIdx.x & 31) == 0 && a.apply) atomicAdd(&a.counters->triTests, myTests);
    }
}

// step 2, block-cooperative form.  The thread-per-particle search above spends > 80 % of its issue slots with
// 5 of 32 lanes active (measured: every particle walks a different number of cells, slot groups and
// triangles).  Here a CTA takes 128 work items at a time and turns the nested search into three flat,
// uniformly executed passes over shared-memory queues:
//   Q1 (particle, triangle cell)   filled per thread from the blood cell's cell mask + particle-level box/slab test
//   Q2 (particle, slot group)      one warp per Q1 entry, one lane per slot group: segment-box vs group-box
//   MT tests                        one lane per (Q2 entry, triangle); near hits race with a 64-bit atomicMin on
//                                   (traversal key, slot) = the first near hit in the reference's order
//   phase B                         one warp per particle with a near hit: lanes = the 27 stencil cells, then slot
//                                   groups, then triangles - is there an earlier (far) hit that masks it?
// Queue overflow (pathological clustering) falls back to the sequential search for that particle.
constexpr int COOP_THREADS = 128;
constexpr int Q1_CAP = 768;
constexpr int Q2_CAP = 1024;

template <bool STATS>
__global__ void __launch_bounds__(COOP_THREADS) vein_collisions_coop_kernel(const VeinCollideArgs a, const CullEntry* __restrict__ list,
                                                                           const int* __restrict__ listCount, int maxP)
{
    __shared__ float4 sPos[COOP_THREADS], sDir[COOP_THREADS], sLo[COOP_THREADS], sHi[COOP_THREADS];
    __shared__ int4 sInfo[COOP_THREADS];                       // pcx, pcy, pcz, packed stencil ranges
    __shared__ unsigned long long sBest[COOP_THREADS];
    __shared__ int sPid[COOP_THREADS];
    __shared__ int sFallback[COOP_THREADS];
    __shared__ unsigned char sSplatOnly[COOP_THREADS];   // ghost particle: wall-force splat only
    __shared__ int2 q1[Q1_CAP];
    __shared__ int4 q2[Q2_CAP];
    __shared__ int q3[COOP_THREADS];
    __shared__ int q1n, q2n, q3n, q1Total;
    __shared__ int q1Off[Q1_CAP];
    __shared__ int sWarpTot[COOP_THREADS / 32];

    const GridDev& g = a.tgrid;
    const float reach = a.phys.impactNear;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int plane = g.nx * g.ny;
    unsigned long long myTests = 0;
    const long long items = (long long)(*listCount) * maxP;

    for (long long batch = (long long)blockIdx.x * COOP_THREADS; batch < items; batch += (long long)gridDim.x * COOP_THREADS) {
        if (tid == 0) { q1n = 0; q2n = 0; q3n = 0; }
        sBest[tid] = ~0ull;
        sFallback[tid] = 0;
        sPid[tid] = -1;
        __syncthreads();

        // ---- pass 0/1: load the particle, enqueue the triangle cells that can hold a near hit
        const long long w = batch + tid;
        if (w < items) {
            const CullEntry ent = list[w / maxP];
            const int c = ent.cell, k = (int)(w % maxP);
            int t = 0;
            while (c >= 0 && t + 1 < a.types.n && c >= a.types.t[t + 1].cStart) ++t;
            if (c < 0 ? k == 0 : k < a.types.t[t].P) {
                const int pid = c < 0 ? -c - 1 : a.types.t[t].pStart + (c - a.types.t[t].cStart) * a.types.t[t].P + k;
                sPid[tid] = pid;
                sSplatOnly[tid] = c < 0;
                const float4 p4 = a.pos[pid], v4 = a.vel[pid];
                const float3 pos = xyz(p4);
                const float3 dir = normalize(xyz(v4));
                const int pcx = axis_cell(pos.x, g.minx, g.lenx, g.csx), pcy = axis_cell(pos.y, g.miny, g.leny, g.csy),
                          pcz = axis_cell(pos.z, g.minz, g.lenz, g.csz);
                if (ent.cx0 < 0 || pcx >= g.nx || pcy >= g.ny || pcz >= g.nz) {
                    sFallback[tid] = 1;
                } else {
                    int x0, x1, y0, y1, z0, z1;
                    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.x - g.minx, (float)g.csx)), g.nx, x0, x1);
                    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.y - g.miny, (float)g.csy)), g.ny, y0, y1);
                    tri_stencil_range(__float2uint_rz(__fdiv_rn(pos.z - g.minz, (float)g.csz)), g.nz, z0, z1);
                    const float3 tip = pos + reach * dir;
                    const float slx = fminf(pos.x, tip.x), shx = fmaxf(pos.x, tip.x);
                    const float sly = fminf(pos.y, tip.y), shy = fmaxf(pos.y, tip.y);
                    const float slz = fminf(pos.z, tip.z), shz = fmaxf(pos.z, tip.z);
                    sPos[tid] = make_float4(pos.x, pos.y, pos.z, 0.f);
                    sDir[tid] = make_float4(dir.x, dir.y, dir.z, 0.f);
                    sLo[tid] = make_float4(slx, sly, slz, 0.f);
                    sHi[tid] = make_float4(shx, shy, shz, 0.f);
                    sInfo[tid] = make_int4(pcx, pcy, pcz, (x0 + 1) | ((x1 + 1) << 2) | ((y0 + 1) << 4) | ((y1 + 1) << 6) | ((z0 + 1) << 8) | ((z1 + 1) << 10));
                    unsigned long long m = ent.mask;
                    while (m) {
                        const int bit = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        const int dx = ent.cx0 + (bit & 3) - pcx, dy = ent.cy0 + ((bit >> 2) & 3) - pcy, dz = ent.cz0 + (bit >> 4) - pcz;
                        if (dx < x0 || dx > x1 || dy < y0 || dy > y1 || dz < z0 || dz > z1) continue;
                        const int tc = ((pcz + dz) * g.ny + (pcy + dy)) * g.nx + (pcx + dx);
                        if (!box_overlap(a.cellBox[tc], slx, sly, slz, shx, shy, shz) || !slab_segment(a.cellSlab[tc], pos, dir, reach)) continue;
                        if (a.cellEnd[tc] < a.cellStart[tc]) continue;
                        const int key = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
                        const int idx = atomicAdd(&q1n, 1);
                        if (idx < Q1_CAP) q1[idx] = make_int2(tid | (key << 8), tc);
                        else sFallback[tid] = 1;
                    }
                }
            }
        }
        __syncthreads();

        // ---- pass 2: one lane per (particle, cell, slot group).  The (entry, group) pairs are flattened with an
        // exclusive scan over the entries' group counts, so every lane has a box test to do (a warp per entry
        // left 2/3 of the lanes idle: a 25-unit cell holds ~10 groups).
        const int n1 = min(q1n, Q1_CAP);
        {
            constexpr int PER = (Q1_CAP + COOP_THREADS - 1) / COOP_THREADS;
            int cnt[PER], sum = 0;
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int e1 = tid * PER + r;
                cnt[r] = 0;
                if (e1 < n1) {
                    const int tc = q1[e1].y;
                    cnt[r] = (a.cellEnd[tc] >> 3) - (a.cellStart[tc] >> 3) + 1;
                }
                sum += cnt[r];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) sWarpTot[warp] = incl;
            __syncthreads();
            int run = incl - sum;
            for (int w2 = 0; w2 < warp; ++w2) run += sWarpTot[w2];
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const int e1 = tid * PER + r;
                if (e1 < n1) q1Off[e1] = run;
                run += cnt[r];
            }
            if (tid == COOP_THREADS - 1) q1Total = run;
            __syncthreads();
        }
        for (int f = tid; f < q1Total; f += COOP_THREADS) {
            // entry whose range [off, off + groups) contains f
            int lo1 = 0, hi1 = n1 - 1;
            while (lo1 < hi1) {
                const int mid = (lo1 + hi1 + 1) >> 1;
                if (q1Off[mid] <= f) lo1 = mid; else hi1 = mid - 1;
            }
            const int2 it = q1[lo1];
            const int pl = it.x & 255;
            const int s = a.cellStart[it.y], e = a.cellEnd[it.y];
            const int gi = (s >> 3) + (f - q1Off[lo1]);
            const float4 lo = sLo[pl], hi = sHi[pl];
            if (box_overlap(a.groupBox[gi], lo.x, lo.y, lo.z, hi.x, hi.y, hi.z) &&
                slab_segment(a.groupSlab[gi], xyz(sPos[pl]), xyz(sDir[pl]), reach)) {
                const int idx = atomicAdd(&q2n, 1);
                if (idx < Q2_CAP) q2[idx] = make_int4(it.x, max(s, gi << 3), min(e, (gi << 3) + 7), 0);
                else sFallback[pl] = 1;
            }
        }
        __syncthreads();

        // ---- pass 3: one lane per (slot group, triangle): Moeller-Trumbore, near hits race for the first place
        const int n2 = min(q2n, Q2_CAP);
        for (int it3 = tid; it3 < n2 * 8; it3 += COOP_THREADS) {
            const int4 q = q2[it3 >> 3];
            const int slot = q.y + (it3 & 7);
            if (slot > q.z) continue;
            const int pl = q.x & 255, key = q.x >> 8;
            if (STATS) ++myTests;
            RayHit cand;
            if (ray_triangle(xyz(sPos[pl]), xyz(sDir[pl]), load_tri(a, slot), cand) && cand.t <= reach)
                atomicMin(&sBest[pl], ((unsigned long long)key << 32) | (unsigned)slot);
        }
        __syncthreads();

        // ---- pass 4: particles with a near hit go to phase B; overflowed particles take the sequential path
        if (sPid[tid] >= 0) {
            if (sFallback[tid]) vein_collide_particle<true, STATS>(a, sPid[tid], myTests, sSplatOnly[tid] != 0);
            else if (sBest[tid] != ~0ull) q3[atomicAdd(&q3n, 1)] = tid;
        }
        __syncthreads();

        // ---- phase B: one warp per particle with a near hit
        for (int e3 = warp; e3 < q3n; e3 += COOP_THREADS / 32) {
            const int pl = q3[e3];
            const float3 pos = xyz(sPos[pl]), dir = xyz(sDir[pl]);
            const int4 info = sInfo[pl];
            const int bestKey = (int)(sBest[pl] >> 32), bestSlot = (int)(sBest[pl] & 0xffffffffu);
            const int x0 = (info.w & 3) - 1, x1 = ((info.w >> 2) & 3) - 1, y0 = ((info.w >> 4) & 3) - 1, y1 = ((info.w >> 6) & 3) - 1,
                      z0 = ((info.w >> 8) & 3) - 1, z1 = ((info.w >> 10) & 3) - 1;
            const int cell = (info.z * g.ny + info.y) * g.nx + info.x;
            // lane == traversal key of one stencil cell (x outer, y, z inner)
            const int dx = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dz = lane % 3 - 1;
            const int c = cell + dz * plane + dy * g.nx + dx;
            bool visit = lane < 27 && lane <= bestKey && dx >= x0 && dx <= x1 && dy >= y0 && dy <= y1 && dz >= z0 && dz <= z1 && c >= 0 &&
                         c < g.cells;
            if (visit) visit = ray_box(a.cellBox[c], pos, dir);
            unsigned cm = __ballot_sync(0xffffffffu, visit);
            bool masked = false;
            while (cm && !masked) {
                const int key = __ffs(cm) - 1;
                cm &= cm - 1;
                const int cc = cell + (key % 3 - 1) * plane + ((key / 3) % 3 - 1) * g.nx + (key / 9 - 1);
                const int s = a.cellStart[cc];
                const int e = (key == bestKey) ? bestSlot - 1 : a.cellEnd[cc];
                if (e < s) continue;
                for (int g0 = s >> 3; g0 <= (e >> 3) && !masked; g0 += 32) {
                    const int gi = g0 + lane;
                    const bool ok = gi <= (e >> 3) && ray_box(a.groupBox[gi], pos, dir);
                    unsigned gm = __ballot_sync(0xffffffffu, ok);
                    while (gm && !masked) {
                        // four slot groups (32 triangles) per round
                        int mine = -1;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int gsel = gm ? g0 + __ffs(gm) - 1 : -1;
                            gm &= gm - 1;
                            if ((lane >> 3) == r) mine = gsel;
                        }
                        bool hitFar = false;
                        if (mine >= 0) {
                            const int slot = (mine << 3) + (lane & 7);
                            if (slot >= s && slot <= e) {
                                if (STATS) ++myTests;
                                RayHit far;
                                hitFar = ray_triangle(pos, dir, load_tri(a, slot), far);
                            }
                        }
                        masked = __any_sync(0xffffffffu, hitFar);
                    }
                }
            }
            if (!masked && lane == 0) {
                RayHit h;
                ray_triangle(pos, dir, load_tri(a, bestSlot), h);
                const int pid = sPid[pl];
                vein_apply_hit(a, pid, a.pos[pid], a.vel[pid], dir, h, sSplatOnly[pl] != 0);
            }
        }
        __syncthreads();
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) myTests += __shfl_xor_sync(0xffffffffu, myTests, o);
        if (lane == 0 && a.apply) atomicAdd(&a.counters->triTests, myTests);
    }
}

// slab mode: ghost particles (owned by a neighbouring rank) only deposit their wall-force splat here, so that every
// rank sees all contributions to the vein vertices it integrates; the particle itself is updated by its owner
__global__ void __launch_bounds__(128) vein_ghost_splat_kernel(const VeinCollideArgs a)
{
    unsigned long long tests = 0;
    const int n = *a.ghostCount;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        vein_collide_particle<true, false>(a, a.ghostList[k], tests, true);
}

void launch_vein_collisions(const VeinCollideArgs& a, cudaStream_t st)
{
    static const bool sequentialEnv = getenv("BCS_VEIN_SEQUENTIAL") != nullptr;
    const bool coop = a.fast && a.cullList && !a.dbgTri && !sequentialEnv;
    if (a.ghostList && a.apply && !a.dbgTri && !coop) BCS_LAUNCH("vein_ghost_splat", st, vein_ghost_splat_kernel<<<64, 128, 0, st>>>(a));
    const int threads = 128, blocks = (a.n + threads - 1) / threads;
    if (a.fast && a.cullList && !a.dbgTri) {
        BCS_CUDA(cudaMemsetAsync(a.cullCount, 0, sizeof(int), st));
        BCS_LAUNCH("vein_cull_cells", st, vein_cull_cells_kernel<<<(a.nCells * 32 + 127) / 128, 128, 0, st>>>(a, a.nCells, a.cullList, a.cullCount));
        if (a.ghostList && a.apply && coop) BCS_LAUNCH("vein_cull_ghosts", st, vein_cull_ghosts_kernel<<<148, 128, 0, st>>>(a, a.cullList, a.cullCount));
        const int grid = min(blocks, 148 * 16);
        static const bool sequential = getenv("BCS_VEIN_SEQUENTIAL") != nullptr;
        if (sequential) {
            if (a.stats) BCS_LAUNCH("vein_collisions", st, vein_collisions_listed_kernel<true><<<grid, threads, 0, st>>>(a, a.cullList, a.cullCount, a.maxP));
            else BCS_LAUNCH("vein_collisions", st, vein_collisions_listed_kernel<false><<<grid, threads, 0, st>>>(a, a.cullList, a.cullCount, a.maxP));
        } else {
            if (a.stats) BCS_LAUNCH("vein_collisions", st, vein_collisions_coop_kernel<true><<<grid, COOP_THREADS, 0, st>>>(a, a.cullList, a.cullCount, a.maxP));
            else BCS_LAUNCH("vein_collisions", st, vein_collisions_coop_kernel<false><<<grid, COOP_THREADS, 0, st>>>(a, a.cullList, a.cullCount, a.maxP));
        }
    } else if (a.fast) {
        if (a.stats) BCS_LAUNCH("vein_collisions_all", st, vein_collisions_kernel<true, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("vein_collisions_all", st, vein_collisions_kernel<true, false><<<blocks, threads, 0, st>>>(a));
    } else {
        if (a.stats) BCS_LAUNCH("vein_collisions_exhaustive", st, vein_collisions_kernel<false, true><<<blocks, threads, 0, st>>>(a));
        else BCS_LAUNCH("vein_collisions_exhaustive", st, vein_collisions_kernel<false, false><<<blocks, threads, 0, st>>>(a));
    }
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

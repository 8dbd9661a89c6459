// The per-blood-cell pass: intra-cell spring forces fused with the blood-cell centres, and - when a step of bcs_step is
// followed by another - with the END of the previous step (integration, vein-end respawn, step counter) and the row
// count of the next grid build, so that the particle state crosses HBM once per step instead of three times.
//
// Stands in for
//   BloodCells::gatherForcesFromNeighbors (objects/blood_cells.cu:122-153): calculateBloodCellsCenters (:44-61) +
//     gatherForcesKernel (:66-120) with physics::calculateParticlesSpringForceComponent (simulation/physics.cuh:53-78,
//     Heun branch), springMassForceWithDampingForParticle (:24-27), accumulateEnvironmentForcesForParticles (:102-120);
//   propagateParticleForcesKernel (objects/blood_cells.cu:155-179, Heun branch) and HandleVeinEnd
//     (simulation/vein_end.cu:12-173);
//   calculateCellIdKernel (grids/uniform_grid.cu:38-49) for the next step's grid.
//
// B200 mapping.  A work unit is a GROUP of G whole blood cells of one type (G = 8 for the reference's 20-particle
// presets), owned by ONE WARP; nothing is ever synchronised across warps (the round-1 kernel spent 29 % of its stall
// samples at two CTA barriers, and measured shared-memory bound - 40 % of its wavefronts were bank-conflict replays).
//   * cp.async (16 B per lane, coalesced in global memory) brings the group's pos / vel / frc into the warp's private
//     tiles CELL-INTERLEAVED: particle k of cell c sits at tile[k * (G + 1) + c].
//   * a lane is (slot, cell) = (lane / G, lane % G): the G lanes of a slot handle the SAME particle index of the G cells.
//     For the spring stage they therefore read the SAME mate index m of their own cell - tile[m * (G + 1) + c], G
//     consecutive float4: conflict free - and share one adjacency entry (rest length, mate).  Every particle sums its
//     springs itself, in ascending mate order (the reference's order); the pair term is evaluated from both ends
//     (twice the arithmetic of an undirected evaluation, half its shared-memory traffic - the binding resource).
//   * the particle indices are dealt to the slots SORTED BY DEGREE, so a round of 32 lanes runs as many iterations as
//     its particles have springs (5-7), not as the worst particle of the type has (13).
//   * a last sweep in memory order adds gravity / viscous damping / the big-cell brake, F <- (F_old + F_new) / 2, and
//     writes back with coalesced stores (plus the row count of the next grid build from the new positions).
// Mate forces are read from the staged tile: the update is a snapshot (the reference races here).
// One launch covers all types (the reference launches per type on separate streams).
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "pair_device.cuh"
#include "rows_device.cuh"
#include "slab.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace bcs {

#ifndef BCS_CP_UNROLL
#define BCS_CP_UNROLL 2
#endif
constexpr int CP_UNROLL = BCS_CP_UNROLL;   // mates of a particle in flight together

constexpr int CP_MAX_CELL = 256;     // particles per blood cell supported by the pass
constexpr int CP_MAX_ROUNDS = 8;     // a group is at most 8 warp rounds (256 particles)

SpringPlan make_spring_plan(const TypesDev& types, const HostScene& hs, SpringTables& tb, int world)
{
    SpringPlan p{};
    tb.slot.clear(); tb.adj.clear();
    const char* forceG = getenv("BCS_SPRING_G");
    int acc = 0, tileMax = 32;
    for (int t = 0; t < types.n; ++t) {
        const TypeDev& ty = types.t[t];
        BCS_REQUIRE(ty.P <= CP_MAX_CELL, BCS_ERR_UNSUPPORTED, "more than 256 particles per blood cell");
        const int P = ty.P;
        // cells per group: a power of two, at most 8 (the lanes of a quarter warp), the group at most 8 warp rounds
        int G = 8;
        while (G > 1 && G * P > 32 * CP_MAX_ROUNDS) G >>= 1;
        // ... and small enough that the machine is not left idle: a group is one warp's work, a B200 holds ~2400 of these
        // warps, and a rank of a slab decomposition only sees its share of the blood cells
        const long long share = std::max<long long>(1, (long long)hs.B / std::max(1, world));
        while (G > 1 && share / G < 2 * 2368) G >>= 1;
        if (forceG) {
            int g = 1;
            while (2 * g <= std::min(atoi(forceG), G)) g <<= 1;
            G = g;
        }
        p.cellsPerBlock[t] = G;
        p.blockStart[t] = acc;
        acc += (ty.count + G - 1) / G;
        tileMax = std::max(tileMax, P * (G + 1));
        p.cellMagic[t] = ((1u << 20) + (unsigned)P - 1u) / (unsigned)P;

        // adjacency per particle of a cell: mates in ascending order (the order blood_cells.cu:88-110 visits them in),
        // springs from a particle to itself dropped (they contribute nothing: normalize(0) = 0, as in the reference)
        std::vector<std::vector<std::pair<int, float>>> adj(P);
        int maxDeg = 0;
        for (int i = 0; i < P; ++i) {
            for (int d = 0; d < ty.maxDeg; ++d) {
                const int j = hs.adjJ[ty.adjStart + d * P + i];
                if (j < 0) break;
                if (j != i) adj[i].push_back({j, hs.adjL[ty.adjStart + d * P + i]});
            }
            maxDeg = std::max(maxDeg, (int)adj[i].size());
        }
        // slots: particle indices sorted by degree (stable), so that a warp round holds similar degrees
        std::vector<int> order(P);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return adj[x].size() < adj[y].size(); });
        p.slotOff[t] = (int)tb.slot.size();
        for (int s = 0; s < P; ++s) tb.slot.push_back(order[s] | ((int)adj[order[s]].size() << 16));
        p.adjOff[t] = (int)tb.adj.size();
        p.adjDeg[t] = maxDeg;
        for (int d = 0; d < maxDeg; ++d)
            for (int s = 0; s < P; ++s) {
                const auto& a = adj[order[s]];
                int m = order[s];
                float L = 0.f;
                if (d < (int)a.size()) { m = a[d].first; L = a[d].second; }
                int Lbits;
                memcpy(&Lbits, &L, 4);
                tb.adj.push_back(make_int2(m, Lbits));
            }
    }
    for (int t = types.n; t <= BCS_MAX_TYPES; ++t) p.blockStart[t] = acc;
    p.totalBlocks = acc;
    p.tileMax = (tileMax + 7) & ~7;
    // per warp: three state tiles (float4), the spring sums (3 float arrays), 8 cell centres, 8 blood-cell ids (slab mode)
    p.warpBytes = 3 * p.tileMax * (int)sizeof(float4) + 3 * p.tileMax * (int)sizeof(float) + 8 * (int)sizeof(float4) + 8 * (int)sizeof(int);
    p.warpBytes = (p.warpBytes + 127) & ~127;
    const char* forceW = getenv("BCS_SPRING_WARPS");
    p.warps = forceW ? std::max(1, std::min(4, atoi(forceW))) : 4;
    while (p.warps > 1 && p.warps * p.warpBytes > 200 * 1024) --p.warps;
    if (tb.slot.empty()) tb.slot.push_back(0);
    if (tb.adj.empty()) tb.adj.push_back(make_int2(0, 0));
    // Per-type tables (adjacency, slots, rest radii) + the vein endings are staged once per CTA behind the warps' regions:
    // with most of the SM's unified memory carved out as shared memory, L1 is too small to keep them (measured: 8 % hit
    // rate, every adjacency load an L2 round trip).  Oversized tables stay in global memory.
    p.tabAdj = (int)tb.adj.size(); p.tabSlot = (int)tb.slot.size(); p.tabModel = hs.nModel; p.tabEnd = std::min((int)hs.endR.size(), 16);
    p.tabBytes = ((2 * p.tabAdj + p.tabSlot + p.tabModel + 4 * p.tabEnd) * 4 + 127) & ~127;
    if (p.tabBytes > 40 * 1024 || getenv("BCS_SPRING_NO_STAGING")) p.tabBytes = 0;
    p.sharedBytes = p.warps * p.warpBytes + p.tabBytes;
    BCS_REQUIRE(p.sharedBytes <= 200 * 1024, BCS_ERR_UNSUPPORTED, "a blood-cell type needs more shared memory than a CTA can have");
    return p;
}

struct CellPassArgs {
    SpringArgs s;
    IntegrateArgs g;
    GridDev grid;
    RowsGrid rows;
    unsigned* doneBlocks;
};

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only): every lane names its own destination, which is what
// lets the tiles be written cell-interleaved while the global side stays coalesced
__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// spring term of physics.cuh:24-27,53-78 for the pair (i <- j), added to `sum`.
// length(dP), normalize(dP) and normalize(-1*dP) of the reference share one reciprocal square root (MUFU.RSQ, <= 2 ulp):
// |dP| = d2 * rsqrt(d2), n = dP * rsqrt(d2); normalize(-dP) = -n exactly.  Deviation from the divided IEEE form: a few
// ulp per component, far inside the 1e-5 contract.
__device__ __forceinline__ void spring_accumulate(float dt, float kSniff, float dFact, float3 pi, float3 vi, float3 fi, float3 pj, float3 vj, float3 fj, float L,
                                                  float3& sum)
{
    const float3 dP = pi - pj;
    const float d2 = dot(dP, dP);
    float inv = rsqrtf(d2);
    if (!(d2 > 0.f)) inv = 0.f;   // coincident particles: normalize() of the reference yields the zero vector (and |dP| = 0)
    const float len = d2 * inv;
    const float3 n = f3(dP.x * inv, dP.y * inv, dP.z * inv);
    const float3 dv2 = (vi - vj) + dt * (fi - fj);
    const float s = (len - L) * kSniff + dot(n, dv2) * dFact;
    sum.x = fmaf(-s, n.x, sum.x); sum.y = fmaf(-s, n.y, sum.y); sum.z = fmaf(-s, n.z, sum.z);
}

template <bool INTEGRATE, bool SPRINGS, bool COUNT, bool LISTS, int MB = 4>
__global__ void __launch_bounds__(128, MB) cell_pass_kernel(const CellPassArgs a)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const SpringPlan& plan = a.s.plan;
    const PhysDev& ph = a.s.phys;
    const OwnedLists& lists = a.s.lists;
    const TypesDev* types = &a.s.types;   // kernel parameter space (constant bank): indexed loads, no trip to L2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* base = smemRaw + (size_t)warp * plan.warpBytes;
    float4* tp = reinterpret_cast<float4*>(base);             // tiles: particle k of cell c at [k * (G + 1) + c]
    float4* tv = tp + plan.tileMax;
    float4* tf = tv + plan.tileMax;
    float* sx = reinterpret_cast<float*>(tf + plan.tileMax);  // spring sums, same indexing
    float* sy = sx + plan.tileMax;
    float* sz = sy + plan.tileMax;
    float4* sc = reinterpret_cast<float4*>(sz + plan.tileMax);   // [8] blood-cell centres of the group
    int* sCid = reinterpret_cast<int*>(sc + 8);                  // [8] slab mode: ids of the group's blood cells
    float4* const gpos = const_cast<float4*>(a.s.pos);
    float4* const gvel = const_cast<float4*>(a.s.vel);

    // per-type tables and vein endings, staged once per CTA (plan.tabBytes != 0)
    const bool staged = plan.tabBytes != 0;
    int* const tab = reinterpret_cast<int*>(smemRaw + (size_t)plan.warps * plan.warpBytes);
    const int2* const sAdj = reinterpret_cast<const int2*>(tab);
    const int* const sSlot = tab + 2 * plan.tabAdj;
    const float* const sInitR = reinterpret_cast<const float*>(sSlot + plan.tabSlot);
    const float* const sEnd = sInitR + plan.tabModel;   // [tabEnd] x (cx, cy, cz, r)
    if (staged) {
        for (int k = threadIdx.x; k < 2 * plan.tabAdj; k += blockDim.x) tab[k] = reinterpret_cast<const int*>(a.s.adjTab)[k];
        for (int k = threadIdx.x; k < plan.tabSlot; k += blockDim.x) tab[2 * plan.tabAdj + k] = a.s.slotTab[k];
        if (SPRINGS)
            for (int k = threadIdx.x; k < plan.tabModel; k += blockDim.x) tab[2 * plan.tabAdj + plan.tabSlot + k] = __float_as_int(a.s.initR[k]);
        if (INTEGRATE)
            for (int k = threadIdx.x; k < plan.tabEnd; k += blockDim.x) {
                float* e = const_cast<float*>(sEnd) + 4 * k;
                e[0] = a.g.endC[3 * k]; e[1] = a.g.endC[3 * k + 1]; e[2] = a.g.endC[3 * k + 2]; e[3] = a.g.endR[k];
            }
        __syncthreads();
    }

    // the step this pass finishes (INTEGRATE): read before any CTA can advance it, see the end of the kernel
    unsigned long long step = 0;
    if (INTEGRATE) step = *reinterpret_cast<const volatile unsigned long long*>(&a.g.counters->step);

    if (INTEGRATE && LISTS && a.g.tail.ghostList) {
        // slab mode: last step's ghosts lose their flag (the collision and wall stages of this step are through with them)
        const int ng = *reinterpret_cast<const volatile int*>(a.g.tail.ghostCount);
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ng; k += gridDim.x * blockDim.x) {
            const int pid = a.g.tail.ghostList[k];
            if (!(a.g.tail.pflag[pid] & 1)) a.g.tail.pflag[pid] = 0;
        }
    }

    const int nTypes = types->n;
    const int totalGroups = LISTS ? lists.blockStart[nTypes] : plan.totalBlocks;
    const int nw = gridDim.x * plan.warps;

    long long tick = a.s.phaseClock ? clock64() : 0;
    auto TICK = [&](int k) {
        if (a.s.phaseClock) {
            const long long now = clock64();
            if (lane == 0) atomicAdd(a.s.phaseClock + k, (unsigned long long)(now - tick));
            tick = now;
        }
    };

    for (int grp = blockIdx.x * plan.warps + warp; grp < totalGroups; grp += nw) {
        int t = 0;
        while (t + 1 < nTypes && grp >= (LISTS ? lists.blockStart[t + 1] : plan.blockStart[t + 1])) ++t;
        const TypeDev ty = types->t[t];
        const int G = plan.cellsPerBlock[t], P = ty.P, stride = G + 1;
        const int firstIdx = (grp - (LISTS ? lists.blockStart[t] : plan.blockStart[t])) * G;
        const int nCells = min(G, (LISTS ? lists.count[t] : ty.count) - firstIdx);
        const int nPart = nCells * P;
        const int g0 = ty.pStart + firstIdx * P;   // first particle of the group (identity cell order, !LISTS)
        const unsigned cellMagic = plan.cellMagic[t];
        const int cell = lane & (G - 1);           // this lane's blood cell within the group
        const int slot0 = lane / G, slotsPerRound = 32 / G;
        if (LISTS) {
            // slab mode: the group's cells are entries of the rank's owned-cell list, anywhere in the arrays
            if (lane < nCells) sCid[lane] = lists.cells[lists.typeFirst[t] + firstIdx + lane];
            __syncwarp();
        }
        auto cell_id = [&](int c) { return LISTS ? sCid[c] : ty.cStart + firstIdx + c; };
        // memory order within the group: element e = c * P + k
        auto split = [&](int e, int& c, int& k) { c = (int)(((unsigned)e * cellMagic) >> 20); k = e - c * P; };
        auto particle_of = [&](int e, int c, int k) { return LISTS ? ty.pStart + (cell_id(c) - ty.cStart) * P + k : g0 + e; };

        // ---- stage the group: coalesced 16-byte async copies, cell-interleaved in shared memory
        for (int e = lane; e < nPart; e += 32) {
            int c, k;
            split(e, c, k);
            const int gi = particle_of(e, c, k), ti = k * stride + c;
            cp_async16(tp + ti, gpos + gi);
            cp_async16(tv + ti, gvel + gi);
            cp_async16(tf + ti, a.s.frc + gi);
        }
        if (!LISTS && lane == 0) {
            // the group this warp takes next: warm L2 while this one is computed
            const int nxt = grp + nw;
            if (nxt < totalGroups) {
                int t2 = t;
                while (t2 + 1 < nTypes && nxt >= plan.blockStart[t2 + 1]) ++t2;
                const TypeDev* y2 = &types->t[t2];
                const int f2 = (nxt - plan.blockStart[t2]) * plan.cellsPerBlock[t2];
                const int n2 = min(plan.cellsPerBlock[t2], y2->count - f2) * y2->P;
                const int q0 = y2->pStart + f2 * y2->P;
                const uint32_t b2 = (uint32_t)(n2 * sizeof(float4));
                l2_prefetch(gpos + q0, b2); l2_prefetch(gvel + q0, b2); l2_prefetch(a.s.frc + q0, b2);
            }
        }
        // deferred fold (pairs.cu): pair forces of the collision pass still parked in fixed point are added to the force
        // before anything reads it.  Their loads are issued here, in memory order (coalesced), so that they fly together
        // with the tile copies; the (few) non-zero sums are folded once the tiles have landed.
        long long ax[CP_MAX_ROUNDS], ay[CP_MAX_ROUNDS], az[CP_MAX_ROUNDS];
        if (INTEGRATE && a.g.pairAcc) {
#pragma unroll
            for (int r = 0; r < CP_MAX_ROUNDS; ++r) {
                const int e = lane + 32 * r;
                ax[r] = 0; ay[r] = 0; az[r] = 0;
                if (e < nPart) {
                    int c, k;
                    split(e, c, k);
                    const long long* acc = a.g.pairAcc + 3 * (size_t)particle_of(e, c, k);
                    ax[r] = acc[0]; ay[r] = acc[1]; az[r] = acc[2];
                }
            }
        }
        TICK(0);   // group set-up + issue
        cp_async_wait_all();
        __syncwarp();
        TICK(1);   // waiting for the tiles
        if (INTEGRATE && a.g.pairAcc) {
#pragma unroll
            for (int r = 0; r < CP_MAX_ROUNDS; ++r) {
                if ((ax[r] | ay[r] | az[r]) != 0) {
                    const int e = lane + 32 * r;
                    int c, k;
                    split(e, c, k);
                    float4& F = tf[k * stride + c];
                    F.x += fx_value(ax[r]); F.y += fx_value(ay[r]); F.z += fx_value(az[r]);
                    long long* acc = a.g.pairAcc + 3 * (size_t)particle_of(e, c, k);
                    acc[0] = 0; acc[1] = 0; acc[2] = 0;
                }
            }
            __syncwarp();
        }

        // ---- end of the previous step: integration, vein end, respawn (blood_cells.cu:155-179, vein_end.cu:57-138)
        unsigned leaving = 0u;   // slab mode: blood cells of the group that change owner after this step
        int target = -1;         // ... and where to (lanes 0 .. G-1: the group's cells)
        if (INTEGRATE) {
            const PhysDev& pg = a.g.phys;
            bool out = false;
            if (cell < nCells) {
                for (int k = slot0; k < P; k += slotsPerRound) {
                    const int ti = k * stride + cell;
                    const float4 F = tf[ti];
                    float4 v = tv[ti], x = tp[ti];
                    const float3 v0 = f3(v.x, v.y, v.z);
                    const float3 v1 = v0 + pg.dt * xyz(F);
                    const float3 dx = (0.5f * pg.dt) * (v1 + v0);
                    v = make_float4(v1.x, v1.y, v1.z, v.w);
                    x = make_float4(x.x + dx.x, x.y + dx.y, x.z + dx.z, x.w);
                    tv[ti] = v;
                    tp[ti] = x;
                    if (pg.useBloodFlow) {
                        if (!ty.warpSync) {
                            const int nStaged = staged ? plan.tabEnd : 0;
                            for (int e = 0; e < nStaged; ++e) {
                                const float4 en = *reinterpret_cast<const float4*>(sEnd + 4 * e);
                                out = out || length_squared(f3(x.x - en.x, x.y - en.y, x.z - en.z)) <= en.w * en.w;
                            }
                            for (int e = nStaged; e < pg.nEndings; ++e) {
                                const float r = __ldg(a.g.endR + e);
                                out = out || length_squared(f3(x.x - __ldg(a.g.endC + 3 * e), x.y - __ldg(a.g.endC + 3 * e + 1), x.z - __ldg(a.g.endC + 3 * e + 2))) <= r * r;
                            }
                        }
                        out = out || x.y <= pg.lowerY || x.y >= pg.upperY || x.x <= pg.leftX || x.x >= pg.rightX || x.z <= pg.backZ || x.z >= pg.frontZ;
                    }
                }
            }
            // "any particle of the cell": the lanes of a cell are those with lane % G == cell
            const unsigned outBits = __ballot_sync(0xffffffffu, out);
            bool flag = false;
            for (int o = cell; o < 32; o += G) flag = flag || ((outBits >> o) & 1u);
            if (__ballot_sync(0xffffffffu, flag)) {
                // a particle of the cell reached a vein end: the WHOLE cell is respawned at the top (Philox per (cell, step))
                if (flag && cell < nCells) {
                    if (slot0 == 0) atomicAdd(&a.g.counters->teleported, 1ull);
                    unsigned ctr[4] = {(unsigned)cell_id(cell), (unsigned)step, (unsigned)(step >> 32), 0u};
                    philox4x32_10(ctr, (unsigned)a.g.seed, (unsigned)(a.g.seed >> 32));
                    const float u1 = u01(ctr[0]), u2 = u01(ctr[1]);
                    const float bx = (u1 - 0.5f) * 1.2f * pg.cylinder_radius, bz = (u2 - 0.5f) * 1.2f * pg.cylinder_radius;
                    for (int k = slot0; k < P; k += slotsPerRound) {
                        const int ti = k * stride + cell;
                        tp[ti] = make_float4(bx + a.g.mx[ty.mStart + k] - a.g.mx[ty.mStart], pg.min_spawn_y + a.g.my[ty.mStart + k] - a.g.my[ty.mStart],
                                             bz + a.g.mz[ty.mStart + k] - a.g.mz[ty.mStart], tp[ti].w);
                        tv[ti] = make_float4(pg.initvx, pg.initvy, pg.initvz, tv[ti].w);
                    }
                }
            }
            __syncwarp();
            if (a.g.slab.enabled) {
                // ownership follows the blood cell's centre: which slab does it lie in after this step?
                if (slot0 == 0 && cell < nCells) {
                    float cy = 0.f;
                    for (int k = 0; k < P; ++k) cy += tp[k * stride + cell].y;
                    cy /= (float)P;
                    if (flag) target = a.g.slab.spawnRank;                               // respawned at the top of the vein
                    else if (cy >= a.g.slab.yHi && a.g.slab.rank > 0) target = a.g.slab.rank - 1;
                    else if (cy < a.g.slab.yLo && a.g.slab.rank < a.g.slab.world - 1) target = a.g.slab.rank + 1;
                    if (target == a.g.slab.rank) target = -1;
                    a.g.moveTo[cell_id(cell)] = (signed char)target;
                }
                // bit c: blood cell c of the group leaves this rank (slot 0's lanes are lanes 0 .. G-1 = the cells)
                leaving = __ballot_sync(0xffffffffu, target >= 0);
            }
        }
        TICK(2);   // integration

        if (SPRINGS) {
            // ---- centre = (p0 + p1 + ...) / P in index order (blood_cells.cu:54-60): a serial sum per cell
            if (slot0 == 0 && cell < nCells) {
                float3 c = f3(0.f, 0.f, 0.f);
                for (int k = 0; k < P; ++k) c = c + xyz(tp[k * stride + cell]);
                c = c / (float)P;
                const float4 c4 = make_float4(c.x, c.y, c.z, 0.f);
                sc[cell] = c4;
                a.s.centers[cell_id(cell)] = c4;
            }
            TICK(3);   // centres

            // ---- springs: slot s of the round is particle index slotTab[s] (sorted by degree); the G lanes of a slot share
            // the adjacency entry and read G consecutive float4 of the mate row
            const int* st = (staged ? sSlot : a.s.slotTab) + plan.slotOff[t];
            const int2* at = (staged ? sAdj : a.s.adjTab) + plan.adjOff[t];
            const int sEnd = ((P + slotsPerRound - 1) / slotsPerRound) * slotsPerRound;
            for (int s = slot0; s < sEnd; s += slotsPerRound) {
                const bool valid = s < P && cell < nCells;
                const int sd = s < P ? st[s] : 0;
                const int k = sd & 0xffff, deg = valid ? sd >> 16 : 0;
                const int ti = k * stride + cell;
                float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = p4, f4 = p4;
                if (valid) { p4 = tp[ti]; v4 = tv[ti]; f4 = tf[ti]; }
                const float3 pi = xyz(p4), vi = xyz(v4), fi = xyz(f4);
                const int maxd = __reduce_max_sync(0xffffffffu, deg);
                float3 sum = f3(0.f, 0.f, 0.f);
                const int2* ap = at + (s < P ? s : 0);
#pragma unroll CP_UNROLL
                for (int d = 0; d < maxd; ++d) {
                    if (d < deg) {
                        const int2 e = ap[d * P];
                        const int tj = e.x * stride + cell;
                        spring_accumulate(ph.dt, ph.particle_k_sniff, ph.particle_d_fact, pi, vi, fi, xyz(tp[tj]), xyz(tv[tj]), xyz(tf[tj]),
                                          __int_as_float(e.y), sum);
                    }
                }
                if (valid) { sx[ti] = sum.x; sy[ti] = sum.y; sz[ti] = sum.z; }
            }
            __syncwarp();
            TICK(4);   // springs
        }

        // ---- memory-order sweep: environment forces, F <- (F_old + F_new) / 2, near-wall probe, row count of the next grid
        // build (a reduction per particle), coalesced write back.
        for (int e0 = 0; e0 < nPart; e0 += 32) {
            const int e = e0 + lane;
            const bool on = e < nPart;
            int c = 0, k = 0;
            if (on) split(e, c, k);
            const int ti = k * stride + c;
            const int gidx = on ? particle_of(e, c, k) : 0;
            float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) p4 = tp[ti];
            // (a blood cell that leaves the rank is counted where it arrives - and here by the pack kernel, slab.cu, if it
            // stays around as a ghost - its springs are still this pass's job: the new owner receives the finished force)
            if (COUNT && on && !((leaving >> c) & 1u)) rows_count_particle(a.grid, a.rows, p4, gidx, 1, a.g.counters);
            bool isNear = false;
            float3 fOut = f3(0.f, 0.f, 0.f);   // the force the particle ends the pass with (slab mode: travels with a leaving cell)
            if (INTEGRATE && !SPRINGS && on) fOut = xyz(tf[ti]);
            if (SPRINGS && on) {
                const float4 v4 = tv[ti], f4 = tf[ti];
                const float3 position = xyz(p4), velocity = xyz(v4), initialForce = xyz(f4);
                float3 newForce = f3(sx[ti], sy[ti], sz[ti]);
                // gravity + viscous damping (+ brake for over-stretched cells)
                const float ratio = length(position - xyz(sc[c])) / (staged ? sInitR[ty.mStart + k] : __ldg(a.s.initR + ty.mStart + k));
                const float3 G3 = f3(ph.gx, ph.gy, ph.gz);
                float3 env;
                if (ph.bigBrake && ratio > ph.max_cell_size_factor) env = G3 - (ph.viscous_damping * ratio * ph.big_brake_intensity) * velocity;
                else env = G3 - ph.viscous_damping * velocity;
                newForce = newForce + env;
                const float3 out = (initialForce + newForce) / 2.0f;
                a.s.frc[gidx] = make_float4(out.x, out.y, out.z, 0.f);
                fOut = out;
                if (a.s.probe.near) {
                    // near-wall probe (NearProbe, kernels.cuh): same cell arithmetic as the wall filter (wall.cu: wall_axis)
                    const NearProbe& probe = a.s.probe;
                    const int hx = (int)fminf(fmaxf(floorf((position.x - probe.ox) * probe.invh), 0.f), (float)(probe.nx - 1));
                    const int hy = (int)fminf(fmaxf(floorf((position.y - probe.oy) * probe.invh), 0.f), (float)(probe.ny - 1));
                    const int hz = (int)fminf(fmaxf(floorf((position.z - probe.oz) * probe.invh), 0.f), (float)(probe.nz - 1));
                    isNear = __ldg(probe.near + (hz * probe.ny + hy) * probe.nx + hx) != 0;
                }
            }
            if (SPRINGS && a.s.probe.near) {
                const unsigned m = __ballot_sync(0xffffffffu, isNear);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    int b = 0;
                    if (lane == leader) b = atomicAdd(a.s.probe.count, __popc(m));
                    b = __shfl_sync(0xffffffffu, b, leader);
                    if (isNear) a.s.probe.list[b + __popc(m & ((1u << lane) - 1u))] = gidx;
                }
            }
            if (INTEGRATE && on) {
                gpos[gidx] = p4; gvel[gidx] = tv[ti];
                if (!SPRINGS && a.g.pairAcc) a.s.frc[gidx] = tf[ti];   // the folded force (the spring stage writes it otherwise)
            }
            if (INTEGRATE && LISTS) {
                if (a.g.tail.ghostList) {
                    // ---- slab mode: pack for the exchange that follows the step (SlabTail, kernels.cuh)
                    const SlabTail& tl = a.g.tail;
                    const SlabDev& sl = a.g.slab;
                    const int tgt = __shfl_sync(0xffffffffu, target, c);
                    const bool leave = on && tgt >= 0;
                    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (on) v4 = tv[ti];
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        // stays: mirrored on the neighbour whose slab it is close to (warp-aggregated append)
                        const bool nearFace = on && !leave && (d == 0 ? (sl.rank > 0 && p4.y >= sl.yHi - sl.haloWidth)
                                                                      : (sl.rank < sl.world - 1 && p4.y < sl.yLo + sl.haloWidth));
                        const unsigned m = __ballot_sync(0xffffffffu, nearFace);
                        if (m) {
                            const int leader = __ffs(m) - 1;
                            int b = 0;
                            if (lane == leader) b = atomicAdd(&tl.sendHdr[d][1], __popc(m));
                            b = __shfl_sync(0xffffffffu, b, leader);
                            if (nearFace) {
                                const int kq = b + __popc(m & ((1u << lane) - 1u));
                                if (kq < tl.capHalo) {
                                    HaloRecord r;
                                    r.id = gidx; r.px = p4.x; r.py = p4.y; r.pz = p4.z; r.vx = v4.x; r.vy = v4.y; r.vz = v4.z; r.pad = 0.f;
                                    reinterpret_cast<HaloRecord*>(tl.halo[d])[kq] = r;
                                } else {
                                    atomicExch(tl.errorFlag, 1);
                                }
                            }
                        }
                    }
                    if (leave) {
                        // the blood cell leaves: full state (finished force, centre) goes to its new owner
                        const int d = tgt == sl.rank - 1 ? 0 : tgt == sl.rank + 1 ? 1 : 2;
                        const int kq = atomicAdd(&tl.sendHdr[d][0], 1);
                        if (kq < tl.capMig) {
                            const float4 ctr = SPRINGS ? sc[c] : a.s.centers[cell_id(c)];
                            MigRecord r;
                            r.id = gidx; r.px = p4.x; r.py = p4.y; r.pz = p4.z; r.vx = v4.x; r.vy = v4.y; r.vz = v4.z;
                            r.fx = fOut.x; r.fy = fOut.y; r.fz = fOut.z; r.cx = ctr.x; r.cy = ctr.y; r.cz = ctr.z;
                            reinterpret_cast<MigRecord*>(tl.mig[d])[kq] = r;
                        } else {
                            atomicExch(tl.errorFlag, 1);
                        }
                        // on this side its particles stay around as ghosts for the next step if they are near the face they
                        // crossed (a respawned cell goes to the top of the vein, far from any face of this slab)
                        const bool keep = (d == 0 && p4.y >= sl.yHi - sl.haloWidth && p4.y < sl.yHi + sl.haloWidth) ||
                                          (d == 1 && p4.y < sl.yLo + sl.haloWidth && p4.y >= sl.yLo - sl.haloWidth);
                        if (keep) tl.keepList[atomicAdd(tl.keepCount, 1)] = gidx;   // flagged + counted by the unpack kernel
                        else tl.pflag[gidx] = 0;
                        if (k == 0) tl.ownedCell[cell_id(c)] = 0;
                    }
                }
            }
        }
        __syncwarp();   // the tiles are free for the next group's copies
        TICK(5);   // environment + count + write back
    }

    if (INTEGRATE) {
        // the last CTA to finish advances the step counter: by then every CTA has read `step`.  No fence: the only thing
        // that must be ordered is this CTA's read of `step` before its own arrival, and the arrival's operand carries a
        // (value-neutral) dependence on the value read, so the atomic cannot issue until the load has returned.
        __syncthreads();
        if (threadIdx.x == 0) {
            if (atomicAdd(a.doneBlocks, 1u + (unsigned)(step >> 63)) == gridDim.x - 1) {
                *a.doneBlocks = 0;
                a.g.counters->step = step + 1;
                // slab mode: every CTA has walked its share of the old ghost list - rewind it for the unpack kernel
                if (LISTS && a.g.tail.ghostList) *a.g.tail.ghostCount = 0;
            }
        }
    }
}

template <bool I, bool S, bool C>
void launch_variant(const CellPassArgs& a, const char* name, cudaStream_t st)
{
    const SpringPlan& plan = a.s.plan;
    int dev = 0, sms = 0;
    BCS_CUDA(cudaGetDevice(&dev));
    BCS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int perSM = std::max(1, std::min(8, (200 * 1024) / (plan.sharedBytes + 1024)));
    const int ctas = std::max(1, std::min((plan.totalBlocks + plan.warps - 1) / plan.warps, sms * perSM));
    const int threads = 32 * plan.warps;
    if (a.s.lists.cells)
        BCS_LAUNCH(name, st, cell_pass_kernel<I, S, C, true><<<ctas, threads, plan.sharedBytes, st>>>(a));
    else
        BCS_LAUNCH(name, st, cell_pass_kernel<I, S, C, false><<<ctas, threads, plan.sharedBytes, st>>>(a));
    BCS_CUDA(cudaGetLastError());
}

template <bool I, bool S, bool C>
void prepare_variant()
{
    BCS_CUDA(cudaFuncSetAttribute(cell_pass_kernel<I, S, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BCS_CUDA(cudaFuncSetAttribute(cell_pass_kernel<I, S, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

}  // namespace

// Function attributes are per device: called once per handle on the handle's device (bcs_create).
void cell_pass_prepare(const SpringPlan&)
{
    prepare_variant<false, true, false>();
    prepare_variant<false, true, true>();
    prepare_variant<true, false, false>();
    prepare_variant<true, true, true>();
    prepare_variant<true, true, false>();
}

void launch_springs(const SpringArgs& s, cudaStream_t st)
{
    CellPassArgs a{};
    a.s = s;
    launch_variant<false, true, false>(a, "springs", st);
}

void launch_springs_count(const SpringArgs& s, const GridDev& grid, const RowsGrid& rows, Counters* counters, cudaStream_t st)
{
    CellPassArgs a{};
    a.s = s; a.grid = grid; a.rows = rows;
    a.g.counters = counters;
    launch_variant<false, true, true>(a, "springs_count", st);
}

void launch_finish_step(const IntegrateArgs& g, const SpringArgs& s, unsigned* doneBlocks, cudaStream_t st)
{
    CellPassArgs a{};
    a.s = s; a.g = g; a.doneBlocks = doneBlocks;
    a.s.probe = NearProbe{};
    launch_variant<true, false, false>(a, "finish_step", st);
}

void launch_advance(const IntegrateArgs& g, const SpringArgs& s, const GridDev& grid, const RowsGrid& rows, unsigned* doneBlocks, cudaStream_t st)
{
    CellPassArgs a{};
    a.s = s; a.g = g; a.grid = grid; a.rows = rows; a.doneBlocks = doneBlocks;
    if (rows.enabled) launch_variant<true, true, true>(a, "advance", st);
    else launch_variant<true, true, false>(a, "advance", st);
}

}  // namespace bcs

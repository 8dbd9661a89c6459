// Intra-cell spring forces, fused with the blood-cell centre computation.
//
// Stands in for BloodCells::gatherForcesFromNeighbors (objects/blood_cells.cu:122-153):
// calculateBloodCellsCenters (:44-61) + gatherForcesKernel (:66-120) with
// physics::calculateParticlesSpringForceComponent (simulation/physics.cuh:53-78, Heun branch),
// springMassForceWithDampingForParticle (:24-27) and accumulateEnvironmentForcesForParticles (:102-120).
//
// B200 mapping.  A work group is a run of whole blood cells of one type (<= 256 particles, contiguous in the
// particle arrays).  The kernel is PERSISTENT (a few CTAs per SM, each looping over groups) and double
// buffered: while a group is being evaluated, one thread has already issued the TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx) that bring the next group's pos/vel/force tiles into the other
// shared-memory stage, so no warp ever waits on HBM.  Inside a group every mate access is a shared-memory read:
//   * the dense PxP spring matrix of the reference is replaced by the type's UNDIRECTED spring list: each spring is
//     evaluated once (the force on its b end is the exact negative of the force on its a end - dP, dv and f_a - f_b
//     all change sign exactly; the reference's directed evaluation does every spring twice) and parked in shared
//     memory;
//   * each particle then sums its incident springs in ascending mate order (the reference's summation order), adds
//     gravity / viscous damping / the big-cell brake and writes F <- (F_old + F_new) / 2;
//   * the type's spring list, incidence table, degrees and rest radii are staged in shared memory once per type.
// Staging makes the update a snapshot: every mate force read is the pre-stage value (the reference races here,
// SURVEY Q7).  One launch covers all types (the reference launches per type on separate streams).
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

#include <cstdint>

namespace bcs {

constexpr int SPRING_THREADS = 256;
constexpr int SPRING_F_CAP = 2048;   // per-CTA capacity of the shared spring-force array (entries)
constexpr int SPRING_STAGES = 2;

SpringPlan make_spring_plan(const TypesDev& types)
{
    SpringPlan p{};
    int acc = 0, sfCap = 0, tabInts = 0;
    for (int t = 0; t < types.n; ++t) {
        const TypeDev& ty = types.t[t];
        int g = SPRING_THREADS / ty.P;
        if (g < 1) g = 1;
        // pairwise evaluation needs cellsPerBlock * springsPerCell entries of shared memory; a type whose single
        // cell does not fit falls back to the directed (per-particle) evaluation
        p.pairwise[t] = ty.nSpr > 0 && ty.nSpr <= SPRING_F_CAP;
        if (p.pairwise[t]) g = min(g, SPRING_F_CAP / ty.nSpr);
        p.cellsPerBlock[t] = g;
        p.blockStart[t] = acc;
        acc += (ty.count + g - 1) / g;
        if (p.pairwise[t]) sfCap = max(sfCap, g * ty.nSpr);
        // per-type tables: spring ends + lengths (pairwise) and the ELL incidence / mate table + lengths, degrees, radii
        tabInts = max(tabInts, 2 * ty.nSpr + 2 * ty.P * ty.maxDeg + 2 * ty.P);
    }
    for (int t = types.n; t <= BCS_MAX_TYPES; ++t) p.blockStart[t] = acc;
    p.totalBlocks = acc;
    p.sfCap = (sfCap + 3) & ~3;
    p.tableInts = (tabInts + 3) & ~3;
    p.sharedBytes = SPRING_STAGES * 3 * SPRING_THREADS * (int)sizeof(float4)   // pos / vel / frc tiles
                    + 3 * p.sfCap * (int)sizeof(float)                          // parked spring forces
                    + SPRING_THREADS * (int)sizeof(float4)                      // blood-cell centres of the group
                    + p.tableInts * (int)sizeof(int)                            // per-type tables
                    + 64;                                                       // mbarriers + group bookkeeping
    return p;
}

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (bytes: multiple of 16, both 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// spring term of physics.cuh:24-27,53-78 for the pair (i <- j): returns the force on i.
// length(dP), normalize(dP) and normalize(-1*dP) of the reference share one reciprocal square root refined to <= 1 ulp
// (|dP| = d2 * rsqrt(d2) with one Newton step; n = dP / |dP| as a product); normalize(-dP) = -n exactly.  Deviation
// from the divided IEEE form: <= 2 ulp per component, far inside the 1e-5 contract.
__device__ __forceinline__ float3 spring_force(float dt, float kSniff, float dFact, float3 pi, float3 vi, float3 fi, float3 pj, float3 vj, float3 fj, float L)
{
    const float3 dP = pi - pj;
    const float d2 = dot(dP, dP);
    float inv = rsqrtf(d2);
    float len = d2 * inv;
    len = fmaf(fmaf(-len, len, d2), 0.5f * inv, len);
    inv = fmaf(fmaf(-len, inv, 1.0f), inv, inv);
    if (!(d2 > 0.f)) { inv = 0.f; len = 0.f; }   // coincident particles: normalize() of the reference yields the zero vector
    const float3 n = f3(dP.x * inv, dP.y * inv, dP.z * inv);
    const float3 dv2 = (vi - vj) + dt * (fi - fj);
    const float s = (len - L) * kSniff + dot(n, dv2) * dFact;
    return f3(-s * n.x, -s * n.y, -s * n.z);
}

struct GroupInfo {
    int t;          // type
    int firstIdx;   // first cell of the group within its type's (owned-)cell sequence
    int nCells;
};

template <bool LISTS>
__global__ void __launch_bounds__(SPRING_THREADS, 5)
springs_kernel(const TypesDev* __restrict__ typesDev, const SpringPlan plan, const PhysDev ph, const float4* __restrict__ pos,
               const float4* __restrict__ vel, float4* __restrict__ frc, float4* __restrict__ centers, const int* __restrict__ adjJ,
               const float* __restrict__ adjL, const int* __restrict__ adjS, const int* __restrict__ sprAB, const float* __restrict__ sprL,
               const float* __restrict__ initR, const OwnedLists lists, const NearProbe probe)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    float4* bpos = reinterpret_cast<float4*>(smemRaw);                     // [STAGES][THREADS]
    float4* bvel = bpos + SPRING_STAGES * SPRING_THREADS;
    float4* bfrc = bvel + SPRING_STAGES * SPRING_THREADS;
    float* sF = reinterpret_cast<float*>(bfrc + SPRING_STAGES * SPRING_THREADS);   // [3 * sfCap]
    float4* sc = reinterpret_cast<float4*>(sF + 3 * plan.sfCap);                   // [THREADS] centres
    int* tab = reinterpret_cast<int*>(sc + SPRING_THREADS);                         // per-type tables
    uint64_t* bars = reinterpret_cast<uint64_t*>(tab + plan.tableInts);             // [STAGES]
    __shared__ int sBlockStart[BCS_MAX_TYPES + 1];
    __shared__ int sCount[BCS_MAX_TYPES];
    __shared__ TypeDev sTy;

    const int tid = threadIdx.x;
    const int nTypes = typesDev->n;
    if (tid <= nTypes) sBlockStart[tid] = LISTS ? lists.blockStart[tid] : plan.blockStart[tid];
    if (tid < nTypes) sCount[tid] = LISTS ? lists.count[tid] : typesDev->t[tid].count;
    if (tid == 0) {
        for (int s = 0; s < SPRING_STAGES; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int totalGroups = sBlockStart[nTypes];

    auto group_info = [&](int grp) {
        GroupInfo gi;
        gi.t = 0;
        while (gi.t + 1 < nTypes && grp >= sBlockStart[gi.t + 1]) ++gi.t;
        const int cpb = plan.cellsPerBlock[gi.t];
        gi.firstIdx = (grp - sBlockStart[gi.t]) * cpb;
        gi.nCells = min(cpb, sCount[gi.t] - gi.firstIdx);
        return gi;
    };
    // thread 0: bring the tiles of group grp into stage s
    auto issue = [&](int grp, int s) {
        const GroupInfo gi = group_info(grp);
        const TypeDev& ty = typesDev->t[gi.t];
        const int P = ty.P;
        float4* dp = bpos + s * SPRING_THREADS;
        float4* dv = bvel + s * SPRING_THREADS;
        float4* df = bfrc + s * SPRING_THREADS;
        mbar_expect_tx(bars + s, (uint32_t)(3 * gi.nCells * P * sizeof(float4)));
        if (!LISTS) {
            const int g0 = ty.pStart + gi.firstIdx * P;
            const uint32_t bytes = (uint32_t)(gi.nCells * P * sizeof(float4));
            tma_load_1d(dp, pos + g0, bytes, bars + s);
            tma_load_1d(dv, vel + g0, bytes, bars + s);
            tma_load_1d(df, frc + g0, bytes, bars + s);
        } else {
            // slab mode: the group's cells are entries of the rank's owned-cell list, anywhere in the arrays
            const uint32_t bytes = (uint32_t)(P * sizeof(float4));
            for (int c = 0; c < gi.nCells; ++c) {
                const int cell = lists.cells[lists.typeFirst[gi.t] + gi.firstIdx + c];
                const int g0 = ty.pStart + (cell - ty.cStart) * P;
                tma_load_1d(dp + c * P, pos + g0, bytes, bars + s);
                tma_load_1d(dv + c * P, vel + g0, bytes, bars + s);
                tma_load_1d(df + c * P, frc + g0, bytes, bars + s);
            }
        }
    };

    int grp = blockIdx.x;
    if (grp >= totalGroups) return;
    if (tid == 0) issue(grp, 0);
    int curType = -1;
    // per-type tables (offsets inside tab)
    int* tAB = tab; float* tL = nullptr; int* tAdj = nullptr; float* tAdjL = nullptr; int* tDeg = nullptr; float* tR = nullptr;
    bool pairwise = false;
    float invNspr = 0.f, invP = 0.f;

    for (int it = 0; grp < totalGroups; grp += gridDim.x, ++it) {
        const int s = it & 1;
        const GroupInfo gi = group_info(grp);
        // prefetch the next group into the other stage (its previous contents were consumed before the barrier that
        // ended the previous iteration)
        if (tid == 0 && grp + (int)gridDim.x < totalGroups) issue(grp + gridDim.x, s ^ 1);

        if (gi.t != curType) {
            // stage this type's tables (a CTA meets each type once: groups are ordered by type)
            __syncthreads();
            curType = gi.t;
            if (tid == 0) sTy = typesDev->t[gi.t];
            __syncthreads();
            const TypeDev ty = sTy;
            pairwise = plan.pairwise[gi.t];
            const int nS = pairwise ? ty.nSpr : 0, nA = ty.P * ty.maxDeg;
            tAB = tab; tL = reinterpret_cast<float*>(tab + nS);
            tAdj = tab + 2 * nS; tAdjL = reinterpret_cast<float*>(tab + 2 * nS + nA);
            tDeg = tab + 2 * nS + 2 * nA; tR = reinterpret_cast<float*>(tab + 2 * nS + 2 * nA + ty.P);
            for (int k = tid; k < nS; k += SPRING_THREADS) { tAB[k] = sprAB[ty.sprStart + k]; tL[k] = sprL[ty.sprStart + k]; }
            for (int k = tid; k < nA; k += SPRING_THREADS) {
                // pairwise: incidence (spring index | sign); directed: mate index.  Absent slots are trimmed by tDeg.
                tAdj[k] = pairwise ? adjS[ty.adjStart + k] : adjJ[ty.adjStart + k];
                tAdjL[k] = adjL[ty.adjStart + k];
            }
            for (int k = tid; k < ty.P; k += SPRING_THREADS) {
                int d = 0;
                while (d < ty.maxDeg && adjJ[ty.adjStart + d * ty.P + k] >= 0) ++d;
                tDeg[k] = d;
                tR[k] = initR[ty.mStart + k];
            }
            invNspr = nS ? 1.0f / (float)nS : 0.f;
            invP = 1.0f / (float)ty.P;
            __syncthreads();
        }
        const int P = sTy.P, nSpr = sTy.nSpr;
        const int nPart = gi.nCells * P;
        const float4* sp = bpos + s * SPRING_THREADS;
        const float4* sv = bvel + s * SPRING_THREADS;
        const float4* sf = bfrc + s * SPRING_THREADS;

        mbar_wait(bars + s, (uint32_t)((it >> 1) & 1));

        if (tid < 32) {
            // centre = (p0 + p1 + ... ) / P in index order (blood_cells.cu:54-60): a serial sum per cell.  All cells of
            // the group are taken by the lanes of ONE warp (dealt over the warps, every warp would issue the whole
            // P-iteration loop for one or two active lanes); the other warps are already evaluating springs, which do
            // not need the centres.
            for (int myCell = tid; myCell < gi.nCells; myCell += 32) {
                float3 c = f3(0.f, 0.f, 0.f);
                for (int k = 0; k < P; ++k) c = c + xyz(sp[myCell * P + k]);
                c = c / (float)P;
                sc[myCell] = make_float4(c.x, c.y, c.z, 0.f);
                const int cellId = LISTS ? lists.cells[lists.typeFirst[gi.t] + gi.firstIdx + myCell] : sTy.cStart + gi.firstIdx + myCell;
                centers[cellId] = make_float4(c.x, c.y, c.z, 0.f);
            }
        }
        if (pairwise) {
            // every undirected spring once; independent iterations, unrolled so that their latencies overlap
            const int total = gi.nCells * nSpr;
#pragma unroll 2
            for (int idx = tid; idx < total; idx += SPRING_THREADS) {
                const int cell = __float2int_rz(((float)idx + 0.5f) * invNspr), k = idx - cell * nSpr;
                const int ab = tAB[k];
                const int ia = cell * P + (ab & 0xffff), ib = cell * P + (ab >> 16);
                const float3 F = spring_force(ph.dt, ph.particle_k_sniff, ph.particle_d_fact, xyz(sp[ia]), xyz(sv[ia]), xyz(sf[ia]), xyz(sp[ib]),
                                              xyz(sv[ib]), xyz(sf[ib]), tL[k]);
                sF[3 * idx] = F.x; sF[3 * idx + 1] = F.y; sF[3 * idx + 2] = F.z;
            }
        }
        __syncthreads();

        if (tid < nPart) {
            const int cell = __float2int_rz(((float)tid + 0.5f) * invP), inCell = tid - cell * P;
            const float4 p4 = sp[tid], v4 = sv[tid], f4 = sf[tid];
            const float3 position = xyz(p4), velocity = xyz(v4), initialForce = xyz(f4);
            float3 newForce = f3(0.f, 0.f, 0.f);
            const int deg = tDeg[inCell];
            if (pairwise) {
                // sum in ascending mate order (the reference's summation order)
                const float* cellF = sF + 3 * cell * nSpr;
                for (int d = 0; d < deg; ++d) {
                    const int e = tAdj[d * P + inCell];
                    if (e == -1) continue;   // a spring from a particle to itself contributes nothing
                    const float* F = cellF + 3 * (e & 0x7fffffff);
                    const float sgn = e < 0 ? -1.0f : 1.0f;   // exact: x + (-1)*F == x - F
                    newForce.x = fmaf(sgn, F[0], newForce.x); newForce.y = fmaf(sgn, F[1], newForce.y); newForce.z = fmaf(sgn, F[2], newForce.z);
                }
            } else {
                const int cellBase = cell * P;
                for (int d = 0; d < deg; ++d) {
                    const int m = cellBase + tAdj[d * P + inCell];
                    newForce = newForce + spring_force(ph.dt, ph.particle_k_sniff, ph.particle_d_fact, position, velocity, initialForce, xyz(sp[m]),
                                                       xyz(sv[m]), xyz(sf[m]), tAdjL[d * P + inCell]);
                }
            }
            // gravity + viscous damping (+ brake for over-stretched cells)
            const float ratio = length(position - xyz(sc[cell])) / tR[inCell];
            const float3 G3 = f3(ph.gx, ph.gy, ph.gz);
            float3 env;
            if (ph.bigBrake && ratio > ph.max_cell_size_factor) env = G3 - (ph.viscous_damping * ratio * ph.big_brake_intensity) * velocity;
            else env = G3 - ph.viscous_damping * velocity;
            newForce = newForce + env;
            const float3 out = (initialForce + newForce) / 2.0f;
            int gidx;
            if (LISTS) {
                const int cellId = lists.cells[lists.typeFirst[gi.t] + gi.firstIdx + cell];
                gidx = sTy.pStart + (cellId - sTy.cStart) * P + inCell;
            } else {
                gidx = sTy.pStart + gi.firstIdx * P + tid;
            }
            frc[gidx] = make_float4(out.x, out.y, out.z, 0.f);
            if (probe.near) {
                // near-wall probe (NearProbe, kernels.cuh): same cell arithmetic as the wall filter (wall.cu: wall_axis)
                const int hx = (int)fminf(fmaxf(floorf((position.x - probe.ox) * probe.invh), 0.f), (float)(probe.nx - 1));
                const int hy = (int)fminf(fmaxf(floorf((position.y - probe.oy) * probe.invh), 0.f), (float)(probe.ny - 1));
                const int hz = (int)fminf(fmaxf(floorf((position.z - probe.oz) * probe.invh), 0.f), (float)(probe.nz - 1));
                const bool isNear = __ldg(probe.near + (hz * probe.ny + hy) * probe.nx + hx) != 0;
                // warp-aggregated append
                const unsigned active = __activemask();
                const unsigned m = __ballot_sync(active, isNear);
                if (m) {
                    const int leader = __ffs(m) - 1, lane = tid & 31;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(probe.count, __popc(m));
                    base = __shfl_sync(active, base, leader);
                    if (isNear) probe.list[base + __popc(m & ((1u << lane) - 1u))] = gidx;
                }
            }
        }
        __syncthreads();   // both the tiles of stage s and sF / sc are free again
    }
}

}  // namespace

void launch_springs(const SpringArgs& a, cudaStream_t st)
{
    static int smCount = 0;
    if (!smCount) {
        int dev = 0;
        BCS_CUDA(cudaGetDevice(&dev));
        BCS_CUDA(cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, dev));
    }
    const int shared = a.plan.sharedBytes;
    static bool attrSet[2] = {false, false};
    const bool lists = a.lists.cells != nullptr;
    if (!attrSet[lists]) {
        if (lists) BCS_CUDA(cudaFuncSetAttribute(springs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        else BCS_CUDA(cudaFuncSetAttribute(springs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attrSet[lists] = true;
    }
    // persistent: as many CTAs as fit (bounded by the group count)
    const int perSM = max(1, min(5, (220 * 1024) / (shared + 1024)));
    const int grid = max(1, min(a.plan.totalBlocks, smCount * perSM));
    if (lists)
        BCS_LAUNCH("springs", st,
                   springs_kernel<true><<<grid, SPRING_THREADS, shared, st>>>(a.typesDev, a.plan, a.phys, a.pos, a.vel, a.frc, a.centers, a.adjJ,
                                                                             a.adjL, a.adjS, a.sprAB, a.sprL, a.initR, a.lists, a.probe));
    else
        BCS_LAUNCH("springs", st,
                   springs_kernel<false><<<grid, SPRING_THREADS, shared, st>>>(a.typesDev, a.plan, a.phys, a.pos, a.vel, a.frc, a.centers, a.adjJ,
                                                                              a.adjL, a.adjS, a.sprAB, a.sprL, a.initR, a.lists, a.probe));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

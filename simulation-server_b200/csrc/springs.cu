// Intra-cell spring forces, fused with the blood-cell centre computation.
//
// Stands in for BloodCells::gatherForcesFromNeighbors (objects/blood_cells.cu:122-153):
// calculateBloodCellsCenters (:44-61) + gatherForcesKernel (:66-120) with
// physics::calculateParticlesSpringForceComponent (simulation/physics.cuh:53-78, Heun branch),
// springMassForceWithDampingForParticle (:24-27) and accumulateEnvironmentForcesForParticles (:102-120).
//
// B200 mapping: a CTA owns a group of whole blood cells (contiguous particle range, so the three float4
// streams are perfectly coalesced), stages their pos/vel/force once in shared memory and resolves every
// mate access there.  The dense PxP spring matrix of the reference is replaced by a per-type ELL adjacency
// (mates in ascending order = the reference's summation order), so no lane iterates over absent springs.
// Staging makes the update a snapshot: every mate force read is the pre-stage value (the reference races
// here).  One launch covers all types (the reference launches per type on separate streams).
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

namespace bcs {

constexpr int SPRING_THREADS = 256;
constexpr int SPRING_F_CAP = 2048;   // per-CTA capacity of the shared spring-force array (float3 entries)

SpringPlan make_spring_plan(const TypesDev& types)
{
    SpringPlan p{};
    int acc = 0;
    for (int t = 0; t < types.n; ++t) {
        int g = SPRING_THREADS / types.t[t].P;
        if (g < 1) g = 1;
        // pairwise evaluation needs cellsPerBlock * springsPerCell entries of shared memory; a type whose single
        // cell does not fit falls back to the directed (per-particle) evaluation
        p.pairwise[t] = types.t[t].nSpr > 0 && types.t[t].nSpr <= SPRING_F_CAP;
        if (p.pairwise[t]) g = min(g, SPRING_F_CAP / types.t[t].nSpr);
        p.cellsPerBlock[t] = g;
        p.blockStart[t] = acc;
        acc += (types.t[t].count + g - 1) / g;
    }
    for (int t = types.n; t <= BCS_MAX_TYPES; ++t) p.blockStart[t] = acc;
    p.totalBlocks = acc;
    p.sharedBytes = 0;
    for (int t = 0; t < types.n; ++t)
        if (p.pairwise[t]) p.sharedBytes = max(p.sharedBytes, (int)(p.cellsPerBlock[t] * types.t[t].nSpr * sizeof(float3)));
    return p;
}

// spring term of physics.cuh:24-27,53-78 for the pair (i <- j): returns the force on i
__device__ __forceinline__ float3 spring_force(const PhysDev& ph, float3 pi, float3 vi, float3 fi, float3 pj, float3 vj, float3 fj, float L)
{
    const float3 dP = pi - pj;
    const float3 dv = vi - vj;
    // length(dP), normalize(dP) and normalize(-1*dP) of the reference share one sqrt: |dP| = sqrtf(dot(dP,dP)),
    // n = dP/|dP| (NaN -> 0), normalize(-dP) = -n exactly.  The three divisions are one reciprocal and three
    // products (<= 1 ulp from the divided form, far inside the 1e-5 contract).
    const float len = sqrtf(dot(dP, dP));
    const float inv = 1.0f / len;
    float3 n = f3(dP.x * inv, dP.y * inv, dP.z * inv);
    if (isnan(n.x) || isnan(n.y) || isnan(n.z)) n = f3(0.f, 0.f, 0.f);
    const float3 dv2 = dv + ph.dt * (fi - fj);
    const float s = (len - L) * ph.particle_k_sniff + dot(n, dv2) * ph.particle_d_fact;
    return s * f3(-n.x, -n.y, -n.z);
}

__global__ void __launch_bounds__(SPRING_THREADS, 5)
springs_kernel(const TypesDev types, const SpringPlan plan, const PhysDev ph, const float4* __restrict__ pos,
               const float4* __restrict__ vel, float4* __restrict__ frc, float4* __restrict__ centers,
               const int* __restrict__ adjJ, const float* __restrict__ adjL, const int* __restrict__ adjS,
               const int* __restrict__ sprAB, const float* __restrict__ sprL, const float* __restrict__ initR,
               const OwnedLists lists)
{
    __shared__ float4 sp[SPRING_THREADS], sv[SPRING_THREADS], sf[SPRING_THREADS];
    __shared__ float3 sc[SPRING_THREADS];
    extern __shared__ float3 sF[];   // cellsPerBlock * springsPerCell entries (plan.sharedBytes)

    // which blood cells does this CTA own?  Whole-scene mode: consecutive cells of one type.  Slab mode: a group of
    // entries of the rank's owned-cell list of one type (device-side counts: surplus CTAs leave at once).
    __shared__ int sCellId[SPRING_THREADS];
    int t = 0, firstIdx, nCells;
    if (lists.cells) {
        if ((int)blockIdx.x >= lists.blockStart[types.n]) return;
        while (t + 1 < types.n && (int)blockIdx.x >= lists.blockStart[t + 1]) ++t;
        firstIdx = ((int)blockIdx.x - lists.blockStart[t]) * plan.cellsPerBlock[t];
        nCells = min(plan.cellsPerBlock[t], lists.count[t] - firstIdx);
    } else {
        while (t + 1 < types.n && (int)blockIdx.x >= plan.blockStart[t + 1]) ++t;
        firstIdx = ((int)blockIdx.x - plan.blockStart[t]) * plan.cellsPerBlock[t];
        nCells = min(plan.cellsPerBlock[t], types.t[t].count - firstIdx);
    }
    const TypeDev ty = types.t[t];
    const int nPart = nCells * ty.P;
    const int tid = threadIdx.x;
    if (tid < nCells) sCellId[tid] = lists.cells ? lists.cells[lists.typeFirst[t] + firstIdx + tid] : ty.cStart + firstIdx + tid;
    __syncthreads();
    // global particle index of this thread's particle
    const int myCell = tid < nPart ? tid / ty.P : 0;
    const int gidx = ty.pStart + (sCellId[myCell] - ty.cStart) * ty.P + (tid - myCell * ty.P);

    float4 p4 = make_float4(0, 0, 0, 0), v4 = p4, f4 = p4;
    if (tid < nPart) {
        p4 = pos[gidx];
        v4 = vel[gidx];
        f4 = frc[gidx];
        sp[tid] = p4; sv[tid] = v4; sf[tid] = f4;
    }
    __syncthreads();
    if (tid < nCells) {
        // centre = (p0 + p1 + ... ) / P in index order (blood_cells.cu:54-60)
        float3 c = f3(0.f, 0.f, 0.f);
        for (int k = 0; k < ty.P; ++k) c = c + xyz(sp[tid * ty.P + k]);
        c = c / (float)ty.P;
        sc[tid] = c;
        centers[sCellId[tid]] = make_float4(c.x, c.y, c.z, 0.f);
    }
    const bool pairwise = plan.pairwise[t];
    if (pairwise) {
        // every undirected spring once: the force on its b end is the exact negative of the force on its a end
        // (dP, dv and f_a - f_b all change sign exactly), so the directed evaluation of the reference does each
        // of these twice
        const int total = nCells * ty.nSpr;
        // independent springs: unrolled so that their sqrt / reciprocal latencies overlap
#pragma unroll 4
        for (int idx = tid; idx < total; idx += SPRING_THREADS) {
            const int cell = idx / ty.nSpr, k = idx - cell * ty.nSpr;
            const int ab = __ldg(sprAB + ty.sprStart + k);
            const int ia = cell * ty.P + (ab & 0xffff), ib = cell * ty.P + (ab >> 16);
            sF[idx] = spring_force(ph, xyz(sp[ia]), xyz(sv[ia]), xyz(sf[ia]), xyz(sp[ib]), xyz(sv[ib]), xyz(sf[ib]), __ldg(sprL + ty.sprStart + k));
        }
    }
    __syncthreads();
    if (tid >= nPart) return;

    const int cell = tid / ty.P, inCell = tid - cell * ty.P, cellBase = cell * ty.P;
    const float3 position = xyz(p4), velocity = xyz(v4), initialForce = xyz(f4);
    float3 newForce = f3(0.f, 0.f, 0.f);
    if (pairwise) {
        // sum in ascending mate order (the reference's summation order)
        const int* as = adjS + ty.adjStart + inCell;
        const int* aj = adjJ + ty.adjStart + inCell;
        for (int d = 0; d < ty.maxDeg; ++d) {
            if (__ldg(aj + d * ty.P) < 0) break;
            const int e = __ldg(as + d * ty.P);
            if (e == -1) continue;
            const float3 F = sF[cell * ty.nSpr + (e & 0x7fffffff)];
            newForce = (e < 0) ? newForce - F : newForce + F;
        }
    } else {
        const int* aj = adjJ + ty.adjStart + inCell;
        const float* al = adjL + ty.adjStart + inCell;
        for (int d = 0; d < ty.maxDeg; ++d) {
            const int j = __ldg(aj + d * ty.P);
            if (j < 0) break;
            const int m = cellBase + j;
            newForce = newForce + spring_force(ph, position, velocity, initialForce, xyz(sp[m]), xyz(sv[m]), xyz(sf[m]), __ldg(al + d * ty.P));
        }
    }
    // gravity + viscous damping (+ brake for over-stretched cells)
    const float ratio = length(position - sc[cell]) / __ldg(initR + ty.mStart + inCell);
    const float3 G3 = f3(ph.gx, ph.gy, ph.gz);
    float3 env;
    if (ph.bigBrake && ratio > ph.max_cell_size_factor) env = G3 - (ph.viscous_damping * ratio * ph.big_brake_intensity) * velocity;
    else env = G3 - ph.viscous_damping * velocity;
    newForce = newForce + env;
    const float3 out = (initialForce + newForce) / 2.0f;
    frc[gidx] = make_float4(out.x, out.y, out.z, 0.f);
}

void launch_springs(const SpringArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("springs", st,
               springs_kernel<<<a.plan.totalBlocks, SPRING_THREADS, a.plan.sharedBytes, st>>>(a.types, a.plan, a.phys, a.pos, a.vel, a.frc, a.centers,
                                                                             a.adjJ, a.adjL, a.adjS, a.sprAB, a.sprL, a.initR, a.lists));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

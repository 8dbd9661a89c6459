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

SpringPlan make_spring_plan(const TypesDev& types)
{
    SpringPlan p{};
    int acc = 0;
    for (int t = 0; t < types.n; ++t) {
        const int g = SPRING_THREADS / types.t[t].P;
        p.cellsPerBlock[t] = g < 1 ? 1 : g;
        p.blockStart[t] = acc;
        acc += (types.t[t].count + p.cellsPerBlock[t] - 1) / p.cellsPerBlock[t];
    }
    for (int t = types.n; t <= BCS_MAX_TYPES; ++t) p.blockStart[t] = acc;
    p.totalBlocks = acc;
    return p;
}

__global__ void __launch_bounds__(SPRING_THREADS)
springs_kernel(const TypesDev types, const SpringPlan plan, const PhysDev ph, const float4* __restrict__ pos,
               const float4* __restrict__ vel, float4* __restrict__ frc, float4* __restrict__ centers,
               const int* __restrict__ adjJ, const float* __restrict__ adjL, const float* __restrict__ initR)
{
    __shared__ float4 sp[SPRING_THREADS], sv[SPRING_THREADS], sf[SPRING_THREADS];
    __shared__ float3 sc[SPRING_THREADS];

    int t = 0;
    while (t + 1 < types.n && (int)blockIdx.x >= plan.blockStart[t + 1]) ++t;
    const TypeDev ty = types.t[t];
    const int G = plan.cellsPerBlock[t];
    const int firstCell = ((int)blockIdx.x - plan.blockStart[t]) * G;
    const int nCells = min(G, ty.count - firstCell);
    const int nPart = nCells * ty.P;
    const int basePart = ty.pStart + firstCell * ty.P;
    const int tid = threadIdx.x;

    float4 p4 = make_float4(0, 0, 0, 0), v4 = p4, f4 = p4;
    if (tid < nPart) {
        p4 = pos[basePart + tid];
        v4 = vel[basePart + tid];
        f4 = frc[basePart + tid];
        sp[tid] = p4; sv[tid] = v4; sf[tid] = f4;
    }
    __syncthreads();
    if (tid < nCells) {
        // centre = (p0 + p1 + ... ) / P in index order (blood_cells.cu:54-60)
        float3 c = f3(0.f, 0.f, 0.f);
        for (int k = 0; k < ty.P; ++k) c = c + xyz(sp[tid * ty.P + k]);
        c = c / (float)ty.P;
        sc[tid] = c;
        centers[ty.cStart + firstCell + tid] = make_float4(c.x, c.y, c.z, 0.f);
    }
    __syncthreads();
    if (tid >= nPart) return;

    const int cell = tid / ty.P, inCell = tid - cell * ty.P, cellBase = cell * ty.P;
    const float3 position = xyz(p4), velocity = xyz(v4), initialForce = xyz(f4);
    float3 newForce = f3(0.f, 0.f, 0.f);
    const int* aj = adjJ + ty.adjStart + inCell;
    const float* al = adjL + ty.adjStart + inCell;
    for (int d = 0; d < ty.maxDeg; ++d) {
        const int j = __ldg(aj + d * ty.P);
        if (j < 0) break;
        const float L = __ldg(al + d * ty.P);
        const int m = cellBase + j;
        const float3 dP = position - xyz(sp[m]);
        const float3 dv = velocity - xyz(sv[m]);
        // length(dP), normalize(dP) and normalize(-1*dP) of the reference share one sqrt and one set of
        // IEEE divisions: |dP| = sqrtf(dot(dP,dP)), n = dP/|dP| (NaN -> 0), normalize(-dP) = -n exactly
        const float len = sqrtf(dot(dP, dP));
        float3 n = dP / len;
        if (isnan(n.x) || isnan(n.y) || isnan(n.z)) n = f3(0.f, 0.f, 0.f);
        const float3 dv2 = dv + ph.dt * (initialForce - xyz(sf[m]));
        const float s = (len - L) * ph.particle_k_sniff + dot(n, dv2) * ph.particle_d_fact;
        newForce = newForce + s * f3(-n.x, -n.y, -n.z);
    }
    // gravity + viscous damping (+ brake for over-stretched cells)
    const float ratio = length(position - sc[cell]) / __ldg(initR + ty.mStart + inCell);
    const float3 G3 = f3(ph.gx, ph.gy, ph.gz);
    float3 env;
    if (ph.bigBrake && ratio > ph.max_cell_size_factor) env = G3 - (ph.viscous_damping * ratio * ph.big_brake_intensity) * velocity;
    else env = G3 - ph.viscous_damping * velocity;
    newForce = newForce + env;
    const float3 out = (initialForce + newForce) / 2.0f;
    frc[basePart + tid] = make_float4(out.x, out.y, out.z, 0.f);
}

void launch_springs(const SpringArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("springs", st,
               springs_kernel<<<a.plan.totalBlocks, SPRING_THREADS, 0, st>>>(a.types, a.plan, a.phys, a.pos, a.vel, a.frc, a.centers,
                                                                             a.adjJ, a.adjL, a.initR));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

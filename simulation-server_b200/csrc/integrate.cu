// Particle integration and vein-end respawn.
//
// Stands in for propagateParticleForcesKernel (objects/blood_cells.cu:155-179, Heun branch) and
// HandleVeinEnd (simulation/vein_end.cu:12-173: handleVeinEndsBlockSync :57-108, handleVeinEndsWarpSync :111-138).
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

#include <algorithm>

namespace bcs {

// v0 = v / gpuCount (gpuCount = 1); v1 = v0 + dt*F; x += 0.5*dt*(v1 + v0); F is left untouched
__global__ void __launch_bounds__(256) integrate_particles_kernel(float4* __restrict__ pos, float4* __restrict__ vel,
                                                                  const float4* __restrict__ frc, int n, float dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 F = frc[i];
    float4 v = vel[i], x = pos[i];
    const float3 v0 = f3(v.x, v.y, v.z);
    const float3 v1 = v0 + dt * xyz(F);
    const float3 dx = (0.5f * dt) * (v1 + v0);
    vel[i] = make_float4(v1.x, v1.y, v1.z, v.w);
    pos[i] = make_float4(x.x + dx.x, x.y + dx.y, x.z + dx.z, x.w);
}

void launch_integrate_particles(const IntegrateArgs& a, cudaStream_t st)
{
    BCS_LAUNCH("integrate_particles", st, integrate_particles_kernel<<<(a.n + 255) / 256, 256, 0, st>>>(a.pos, a.vel, a.frc, a.n, a.phys.dt));
    BCS_CUDA(cudaGetLastError());
}

// One thread per blood cell: any particle inside a vein-ending sphere (only for types the reference would
// run through the block-sync kernel) or outside the safety box teleports the WHOLE cell back to the top.
// The respawn draw is counter-based: Philox4x32-10(key = seed, counter = (cell index, step)); the
// reference shares one racing cuRAND XORWOW state per cell seeded from time(0).
__global__ void __launch_bounds__(128) vein_end_kernel(const IntegrateArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < a.nCells && a.phys.useBloodFlow) {
        const PhysDev& ph = a.phys;
        int t = 0;
        while (t + 1 < a.types.n && c >= a.types.t[t + 1].cStart) ++t;
        const TypeDev ty = a.types.t[t];
        const int first = ty.pStart + (c - ty.cStart) * ty.P;
        bool tp = false;
        for (int k = 0; k < ty.P; ++k) {
            const float4 p = a.pos[first + k];
            bool one = false;
            if (!ty.warpSync)
                for (int e = 0; e < ph.nEndings; ++e) {
                    const float r = a.endR[e];
                    one = one || length_squared(f3(p.x - a.endC[3 * e], p.y - a.endC[3 * e + 1], p.z - a.endC[3 * e + 2])) <= r * r;
                }
            one = one || p.y <= ph.lowerY || p.y >= ph.upperY || p.x <= ph.leftX || p.x >= ph.rightX || p.z <= ph.backZ ||
                  p.z >= ph.frontZ;
            tp = tp || one;
        }
        if (tp) {
            const unsigned long long step = a.counters->step;
            unsigned ctr[4] = {(unsigned)c, (unsigned)step, (unsigned)(step >> 32), 0u};
            philox4x32_10(ctr, (unsigned)a.seed, (unsigned)(a.seed >> 32));
            const float u1 = u01(ctr[0]), u2 = u01(ctr[1]);
            const float bx = (u1 - 0.5f) * 1.2f * ph.cylinder_radius, bz = (u2 - 0.5f) * 1.2f * ph.cylinder_radius;
            const float m0x = a.mx[ty.mStart], m0y = a.my[ty.mStart], m0z = a.mz[ty.mStart];
            for (int k = 0; k < ty.P; ++k) {
                const float w = a.pos[first + k].w;
                a.pos[first + k] = make_float4(bx + a.mx[ty.mStart + k] - m0x, ph.min_spawn_y + a.my[ty.mStart + k] - m0y,
                                               bz + a.mz[ty.mStart + k] - m0z, w);
                a.vel[first + k] = make_float4(ph.initvx, ph.initvy, ph.initvz, a.vel[first + k].w);
            }
            atomicAdd(&a.counters->teleported, 1ull);
        }
    }
}

// Fused tail of the step used by bcs_step: integrate + vein-end test + respawn + step counter in ONE pass
// over the particle state (the staged entry points above are kept for stage-by-stage parity tests).
// A CTA owns whole blood cells (same mapping as the spring kernel), so the "any particle of the cell"
// reduction of handleVeinEndsBlockSync/WarpSync is a shared-memory OR.
constexpr int FINISH_THREADS = 256;
constexpr int FINISH_MAX_ENDINGS = 16;   // vein endings staged in shared memory (more: read from global)

__global__ void __launch_bounds__(FINISH_THREADS) finish_step_kernel(const IntegrateArgs a, const SpringPlan plan, unsigned* __restrict__ doneBlocks)
{
    __shared__ int sOut[FINISH_THREADS];
    __shared__ int sCell[FINISH_THREADS];
    const PhysDev& ph = a.phys;
    __shared__ int sCellId[FINISH_THREADS];
    __shared__ float sY[FINISH_THREADS];
    const OwnedLists& lists = a.lists;
    const int tid = threadIdx.x;
    const unsigned long long step = *reinterpret_cast<const volatile unsigned long long*>(&a.counters->step);   // read before any block can advance it (see below)
    // vein endings: staged once per CTA so that their loads overlap the particle loads
    __shared__ float sEnd[4 * FINISH_MAX_ENDINGS];
    const int nEnd = min(ph.nEndings, FINISH_MAX_ENDINGS);
    if (ph.useBloodFlow && threadIdx.x < nEnd) {
        sEnd[4 * threadIdx.x] = a.endC[3 * threadIdx.x]; sEnd[4 * threadIdx.x + 1] = a.endC[3 * threadIdx.x + 1];
        sEnd[4 * threadIdx.x + 2] = a.endC[3 * threadIdx.x + 2]; sEnd[4 * threadIdx.x + 3] = a.endR[threadIdx.x];
    }
    // slab mode: groups of this rank's owned blood cells, taken by a bounded grid striding over the device-side group
    // count; without slabs the grid covers every group (one round)
    const int totalGroups = lists.cells ? lists.blockStart[a.types.n] : plan.totalBlocks;
    for (int vb = blockIdx.x; vb < totalGroups; vb += gridDim.x) {
        int t = 0, firstIdx, nCells;
        if (lists.cells) {
            while (t + 1 < a.types.n && vb >= lists.blockStart[t + 1]) ++t;
            firstIdx = (vb - lists.blockStart[t]) * plan.cellsPerBlock[t];
            nCells = min(plan.cellsPerBlock[t], lists.count[t] - firstIdx);
        } else {
            while (t + 1 < a.types.n && vb >= plan.blockStart[t + 1]) ++t;
            firstIdx = (vb - plan.blockStart[t]) * plan.cellsPerBlock[t];
            nCells = min(plan.cellsPerBlock[t], a.types.t[t].count - firstIdx);
        }
        const TypeDev ty = a.types.t[t];
        const int nPart = nCells * ty.P;
        __syncthreads();   // shared arrays of the previous round are free (and the staged endings are visible)
        if (tid < nCells) sCellId[tid] = lists.cells ? lists.cells[lists.typeFirst[t] + firstIdx + tid] : ty.cStart + firstIdx + tid;
        __syncthreads();
        const bool mine = tid < nPart;
        const int myCell = mine ? tid / ty.P : 0;
        const int gidx = ty.pStart + (sCellId[myCell] - ty.cStart) * ty.P + (tid - myCell * ty.P);

        float4 x = make_float4(0, 0, 0, 0), v = x;
        bool out = false;
        if (mine) {
            const float4 F = a.frc[gidx];
            v = a.vel[gidx];
            x = a.pos[gidx];
            const float3 v0 = f3(v.x, v.y, v.z);
            const float3 v1 = v0 + ph.dt * xyz(F);
            const float3 dx = (0.5f * ph.dt) * (v1 + v0);
            v = make_float4(v1.x, v1.y, v1.z, v.w);
            x = make_float4(x.x + dx.x, x.y + dx.y, x.z + dx.z, x.w);
            if (ph.useBloodFlow) {
                if (!ty.warpSync) {
                    for (int e = 0; e < nEnd; ++e) {
                        const float r = sEnd[4 * e + 3];
                        out = out || length_squared(f3(x.x - sEnd[4 * e], x.y - sEnd[4 * e + 1], x.z - sEnd[4 * e + 2])) <= r * r;
                    }
                    for (int e = nEnd; e < ph.nEndings; ++e) {
                        const float r = a.endR[e];
                        out = out || length_squared(f3(x.x - a.endC[3 * e], x.y - a.endC[3 * e + 1], x.z - a.endC[3 * e + 2])) <= r * r;
                    }
                }
                out = out || x.y <= ph.lowerY || x.y >= ph.upperY || x.x <= ph.leftX || x.x >= ph.rightX || x.z <= ph.backZ || x.z >= ph.frontZ;
            }
        }
        sOut[tid] = out ? 1 : 0;
        __syncthreads();
        if (tid < nCells) {
            int any = 0;
            for (int k = 0; k < ty.P; ++k) any |= sOut[tid * ty.P + k];
            sCell[tid] = any;
            if (any) atomicAdd(&a.counters->teleported, 1ull);
        }
        __syncthreads();
        if (mine) {
            const int cell = tid / ty.P, k = tid - cell * ty.P;
            if (sCell[cell]) {
                unsigned ctr[4] = {(unsigned)sCellId[cell], (unsigned)step, (unsigned)(step >> 32), 0u};
                philox4x32_10(ctr, (unsigned)a.seed, (unsigned)(a.seed >> 32));
                const float u1 = u01(ctr[0]), u2 = u01(ctr[1]);
                const float bx = (u1 - 0.5f) * 1.2f * ph.cylinder_radius, bz = (u2 - 0.5f) * 1.2f * ph.cylinder_radius;
                x = make_float4(bx + a.mx[ty.mStart + k] - a.mx[ty.mStart], ph.min_spawn_y + a.my[ty.mStart + k] - a.my[ty.mStart],
                                bz + a.mz[ty.mStart + k] - a.mz[ty.mStart], x.w);
                v = make_float4(ph.initvx, ph.initvy, ph.initvz, v.w);
            }
            a.pos[gidx] = x;
            a.vel[gidx] = v;
        }
        if (a.slab.enabled) {
            // ownership follows the blood cell's centre: which slab does it lie in after this step?
            sY[tid] = x.y;
            __syncthreads();
            if (tid < nCells) {
                float cy = 0.f;
                for (int k = 0; k < ty.P; ++k) cy += sY[tid * ty.P + k];
                cy /= (float)ty.P;
                int target = -1;
                if (sCell[tid]) target = a.slab.spawnRank;                               // respawned at the top of the vein
                else if (cy >= a.slab.yHi && a.slab.rank > 0) target = a.slab.rank - 1;
                else if (cy < a.slab.yLo && a.slab.rank < a.slab.world - 1) target = a.slab.rank + 1;
                if (target == a.slab.rank) target = -1;
                a.moveTo[sCellId[tid]] = (signed char)target;
            }
        }
    }
    // the last CTA to finish advances the step counter: by then every CTA has read `step`.  No fence: the only thing
    // that must be ordered is this CTA's read of `step` before its own arrival, and the arrival's operand carries a
    // (value-neutral) dependence on the value read, so the atomic cannot issue until the load has returned.
    if (tid == 0) {
        if (atomicAdd(doneBlocks, 1u + (unsigned)(step >> 63)) == gridDim.x - 1) {
            *doneBlocks = 0;
            a.counters->step = step + 1;
        }
    }
}

void launch_finish_step(const IntegrateArgs& a, const SpringPlan& plan, unsigned* doneBlocks, cudaStream_t st)
{
    BCS_LAUNCH("finish_step", st, finish_step_kernel<<<a.lists.cells ? std::min(plan.totalBlocks, BOUNDED_BLOCKS) : plan.totalBlocks, FINISH_THREADS, 0, st>>>(a, plan, doneBlocks));
    BCS_CUDA(cudaGetLastError());
}

__global__ void advance_step_kernel(Counters* c) { c->step += 1; }

void launch_vein_end(const IntegrateArgs& a, cudaStream_t st)
{
    if (a.phys.useBloodFlow) BCS_LAUNCH("vein_end", st, vein_end_kernel<<<(a.nCells + 127) / 128, 128, 0, st>>>(a));
    BCS_LAUNCH("advance_step", st, advance_step_kernel<<<1, 1, 0, st>>>(a.counters));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

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

// (the fused tail of bcs_step - integrate + vein end + step counter in one pass - is the cell pass, cellpass.cu)

__global__ void advance_step_kernel(Counters* c) { c->step += 1; }

void launch_vein_end(const IntegrateArgs& a, cudaStream_t st)
{
    if (a.phys.useBloodFlow) BCS_LAUNCH("vein_end", st, vein_end_kernel<<<(a.nCells + 127) / 128, 128, 0, st>>>(a));
    BCS_LAUNCH("advance_step", st, advance_step_kernel<<<1, 1, 0, st>>>(a.counters));
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

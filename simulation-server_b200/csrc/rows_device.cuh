// Device helpers of the row-directory grid (RowsGrid, kernels.cuh) shared by the grid build (grid.cu), the cell pass
// (cellpass.cu: it counts for the next build while the new position is still in registers) and the pair search (pairs.cu).
#pragma once
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

namespace bcs {

__device__ __forceinline__ int rows_div(unsigned n, unsigned long long magic, int shift)
{
    return (int)(((unsigned long long)n * magic) >> shift);   // exact for n < 2^31 (magic = ceil(2^shift / d), shift = 32 + ceil(log2 d))
}

// directory row of a cell id: key / nx, re-based to the rank's window of cell rows under slab decomposition
__device__ __forceinline__ int rows_local(const RowsGrid& R, int cy, int cz)
{
    int y = cy - R.y0;
    if ((unsigned)y >= (unsigned)R.nyL) {
        // an active particle outside slab + halo: the halo width no longer fits the run.  [1], [2]: its cell row / layer
        if (atomicCAS(R.error, 0, 1) == 0) { R.error[1] = cy; R.error[2] = cz; }
        y = max(0, min(y, R.nyL - 1));
    }
    return cz * R.nyL + y;
}
__device__ __forceinline__ int rows_of_key(const RowsGrid& R, int key)
{
    const int grow = rows_div((unsigned)key, R.nxMagic, R.nxShift);
    if (!R.local) return grow;
    const int cz = rows_div((unsigned)grow, R.nyMagic, R.nyShift);
    return rows_local(R, grow - cz * R.ny, cz);
}

// Cell id, row and "irregular" flag of a position (calculateIdForCell, grids/uniform_grid.cu:24-36, plus the unclamped
// per-axis index particle_collisions.cuh:117-119 trims the stencil by).  One quotient per axis serves both.
struct RowKey {
    int key, row;
    bool oob, irregular;
};
__device__ __forceinline__ RowKey rows_key(const GridDev& g, const RowsGrid& R, const float4 p)
{
    RowKey o;
    o.oob = p.x < g.minx || p.x > g.maxx || p.y < g.miny || p.y > g.maxy || p.z < g.minz || p.z > g.maxz;
    float qx, qy, qz;
    if (g.pow2) { qx = (p.x - g.minx) * g.icsx; qy = (p.y - g.miny) * g.icsy; qz = (p.z - g.minz) * g.icsz; }
    else { qx = __fdiv_rn(p.x - g.minx, (float)g.csx); qy = __fdiv_rn(p.y - g.miny, (float)g.csy); qz = __fdiv_rn(p.z - g.minz, (float)g.csz); }
    // axis_cell: (int)min(len, max(0, q)) with the reference's macro forms; axis_cell_raw: (int)q
    const float mx = (0.f > qx) ? 0.f : qx, my = (0.f > qy) ? 0.f : qy, mz = (0.f > qz) ? 0.f : qz;
    const int cx = (int)((g.lenx > mx) ? mx : g.lenx), cy = (int)((g.leny > my) ? my : g.leny), cz = (int)((g.lenz > mz) ? mz : g.lenz);
    const int rx = (int)qx, ry = (int)qy, rz = (int)qz;
    int key = cz * g.nx * g.ny + cy * g.nx + cx;
    if (o.oob) key = max(0, min(key, g.cells - 1));
    else if (key >= g.cells) key = g.cells - 1;
    // a particle whose cell coordinates are not those of an interior grid cell has a one-sided stencil
    // (particle_collisions.cuh:126-268 trims by the UNclamped index): the symmetric pair search steps aside for the build
    o.irregular = o.oob || rx != cx || ry != cy || rz != cz || cx >= g.nx || cy >= g.ny || cz >= g.nz;
    o.key = key;
    o.row = o.irregular ? rows_of_key(R, key) : (R.local ? rows_local(R, cy, cz) : cz * g.ny + cy);
    return o;
}

// key of one particle, its row counted (a reduction: no value comes back, nothing waits).  flag bit 0: owned (only owned
// particles are counted as out of bounds).
__device__ __forceinline__ void rows_count_particle(const GridDev& g, const RowsGrid& R, const float4 p, int i, int flag, Counters* __restrict__ counters)
{
    const RowKey k = rows_key(g, R, p);
    if (k.oob && (flag & 1)) atomicAdd(&counters->oob, 1ull);
    if (k.irregular) R.irregular[0] = 1;
    atomicAdd(&R.rowCount[k.row], 1u);
    R.keyOf[i] = k.key;
}

}  // namespace bcs

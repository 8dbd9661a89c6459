// Uniform-grid build: cell keys -> stable LSD radix sort by cell key -> cell start/end tables
// (+ reordering of particle state into sorted-slot order for the collision pass).
//
// Stands in for UniformGrid::calculateGrid (grids/uniform_grid.cu:129-155): calculateCellIdKernel (:38-49),
// thrust::stable_sort_by_key (:144-147) and calculateStartAndEndOfCellKernel (:51-80).  Results are
// bit-identical by construction: same key formula, and a stable sort of (key, id) is unique.
#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "rows_device.cuh"

#include <algorithm>
#include <cstdlib>
#include <string>

namespace bcs {

// ------------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cell_keys_kernel(const float4* __restrict__ pos, GridDev g, int* __restrict__ keys,
                                                        int* __restrict__ ids, unsigned* __restrict__ digitTotals, int passes,
                                                        Counters* __restrict__ counters)
{
    __shared__ unsigned hist[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += stride) {
        const float4 p = pos[i];
        const bool oob = p.x < g.minx || p.x > g.maxx || p.y < g.miny || p.y > g.maxy || p.z < g.minz || p.z > g.maxz;
        int key = axis_cell(p.z, g.minz, g.lenz, g.csz) * g.nx * g.ny + axis_cell(p.y, g.miny, g.leny, g.csy) * g.nx +
                  axis_cell(p.x, g.minx, g.lenx, g.csx);
        if (oob) {
            // The reference printf()s and then indexes out of its tables; clamp for memory safety and count.
            atomicAdd(&counters->oob, 1ull);
            key = max(0, min(key, g.cells - 1));
        } else if (key >= g.cells) {
            key = g.cells - 1;   // position exactly on the upper bound
        }
        keys[i] = key;
        ids[i] = i;
        for (int p2 = 0; p2 < passes; ++p2) atomicAdd(&hist[p2][(key >> (8 * p2)) & 255], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
        const unsigned v = (&hist[0][0])[i];
        if (v) atomicAdd(&digitTotals[i], v);
    }
}

// ------------------------------------------------------------------------------------------------
// radix sort, one 8-bit digit per pass: per-tile histograms -> per-bin scan -> stable scatter
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_THREADS / 32;

__global__ void __launch_bounds__(SORT_THREADS) radix_tile_hist_kernel(const int* __restrict__ keys, int n, int shift,
                                                                       unsigned* __restrict__ tileHist)
{
    __shared__ unsigned hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int idx = base + i * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & 255], 1u);
    }
    __syncthreads();
    tileHist[(size_t)blockIdx.x * 256 + threadIdx.x] = hist[threadIdx.x];
}

// one block per digit value: global base of the bin + exclusive prefix of the bin's counts over tiles
__global__ void __launch_bounds__(256) radix_scan_kernel(unsigned* __restrict__ tileHist, int numTiles,
                                                         const unsigned* __restrict__ digitTotals)
{
    __shared__ unsigned warpSums[8];
    __shared__ unsigned carry;
    const int bin = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // bin base = sum of totals of smaller digits
    unsigned v = (threadIdx.x < bin) ? digitTotals[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) warpSums[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int w = 0; w < 8; ++w) s += warpSums[w];
        carry = s;
    }
    __syncthreads();
    for (int t0 = 0; t0 < numTiles; t0 += 256) {
        const int t = t0 + threadIdx.x;
        const unsigned c = (t < numTiles) ? tileHist[(size_t)t * 256 + bin] : 0u;
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        unsigned wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warpSums[w];
        const unsigned base = carry;
        if (t < numTiles) tileHist[(size_t)t * 256 + bin] = base + wbase + incl - c;
        __syncthreads();
        if (threadIdx.x == 255) carry = base + wbase + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const int* __restrict__ keysIn, const int* __restrict__ valsIn,
                                                                     int* __restrict__ keysOut, int* __restrict__ valsOut, int n,
                                                                     int shift, const unsigned* __restrict__ tileBase)
{
    __shared__ unsigned warpCnt[SORT_WARPS][256];   // per-warp digit counters, later exclusive offsets
    __shared__ unsigned binStart[256];              // tile-local start of each digit
    __shared__ unsigned scanTmp[8];
    __shared__ int exKeys[SORT_TILE];
    __shared__ int exVals[SORT_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tileStart = blockIdx.x * SORT_TILE;
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&warpCnt[0][0])[i] = 0;
    __syncthreads();

    int key[SORT_ITEMS], val[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
    const unsigned ltMask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        // warp-striped: warp w owns [w*256, (w+1)*256) of the tile; element order inside the tile is
        // (warp, item, lane), which is also ascending input order -> stability
        const int idx = tileStart + warp * (32 * SORT_ITEMS) + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? keysIn[idx] : 0;
        val[i] = valid ? valsIn[idx] : 0;
        const unsigned digit = valid ? ((unsigned)(key[i] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (valid && lane == leader) {
            base = warpCnt[warp][digit];
            warpCnt[warp][digit] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[i] = base + __popc(peers & ltMask);
        __syncwarp();
    }
    __syncthreads();

    // per digit (thread == digit): exclusive prefix over warps, then exclusive scan over digits
    unsigned total = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
        const unsigned c = warpCnt[w][tid];
        warpCnt[w][tid] = total;
        total += c;
    }
    unsigned incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) scanTmp[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += scanTmp[w];
    binStart[tid] = wbase + incl - total;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int idx = tileStart + warp * (32 * SORT_ITEMS) + i * 32 + lane;
        if (idx < n) {
            const unsigned digit = (unsigned)(key[i] >> shift) & 255u;
            const unsigned pos = binStart[digit] + warpCnt[warp][digit] + rank[i];
            exKeys[pos] = key[i];
            exVals[pos] = val[i];
        }
    }
    __syncthreads();

    const int tileCount = min(SORT_TILE, n - tileStart);
    const unsigned* base = tileBase + (size_t)blockIdx.x * 256;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int j = i * SORT_THREADS + tid;
        if (j < tileCount) {
            const int k = exKeys[j];
            const unsigned digit = (unsigned)(k >> shift) & 255u;
            const unsigned dst = base[digit] + (unsigned)j - binStart[digit];
            keysOut[dst] = k;
            valsOut[dst] = exVals[j];
        }
    }
}

// Single-kernel pass ("onesweep"): the same stable in-tile ranking, but the global offset of a tile's share of
// every digit comes from a decoupled look-back over per-tile status words instead of two extra kernels and
// an extra read of the keys.  status[tile][digit] = (flag << 30) | count, flag 1 = tile-local count,
// flag 2 = inclusive prefix over all earlier tiles; one 32-bit word carries flag and value, so no fences
// are needed.  Tiles are handed out by an atomic ticket, which guarantees that every tile a block waits
// for has already started.
#define STATUS_LOCAL (1u << 30)
#define STATUS_INCL (2u << 30)
#define STATUS_VALUE ((1u << 30) - 1u)

__global__ void __launch_bounds__(SORT_THREADS) radix_onesweep_kernel(const int* __restrict__ keysIn, const int* __restrict__ valsIn,
                                                                      int* __restrict__ keysOut, int* __restrict__ valsOut, int nArg,
                                                                      const int* __restrict__ nDev, int shift,
                                                                      const unsigned* __restrict__ digitTotals, volatile unsigned* status,
                                                                      unsigned* ticket)
{
    const int n = nDev ? *nDev : nArg;   // slab mode: the number of active (owned + ghost) particles lives on the device
    __shared__ unsigned warpCnt[SORT_WARPS][256];
    __shared__ unsigned binStart[256];
    __shared__ unsigned globalBase[256];
    __shared__ unsigned scanTmp[8];
    __shared__ unsigned sTile;
    __shared__ int exKeys[SORT_TILE];
    __shared__ int exVals[SORT_TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&warpCnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)sTile;
    const int tileStart = tile * SORT_TILE;
    if (tileStart >= n) return;   // launched for capacity; tiles past the data are never waited for

    int key[SORT_ITEMS], val[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
    const unsigned ltMask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int idx = tileStart + warp * (32 * SORT_ITEMS) + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? keysIn[idx] : 0;
        val[i] = valid ? valsIn[idx] : 0;
        const unsigned digit = valid ? ((unsigned)(key[i] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (valid && lane == leader) {
            base = warpCnt[warp][digit];
            warpCnt[warp][digit] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[i] = base + __popc(peers & ltMask);
        __syncwarp();
    }
    __syncthreads();

    // thread == digit: count of the digit in this tile, exclusive prefix over warps
    unsigned total = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
        const unsigned c = warpCnt[w][tid];
        warpCnt[w][tid] = total;
        total += c;
    }
    // publish the tile-local count as early as possible
    status[(size_t)tile * 256 + tid] = STATUS_LOCAL | total;

    // two block-wide exclusive scans over the 256 digits: tile-local starts and global bin bases
    unsigned inclA = total, inclB = digitTotals[tid];
    const unsigned myTotalB = inclB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned ya = __shfl_up_sync(0xffffffffu, inclA, o), yb = __shfl_up_sync(0xffffffffu, inclB, o);
        if (lane >= o) { inclA += ya; inclB += yb; }
    }
    __shared__ unsigned scanTmpB[8];
    if (lane == 31) { scanTmp[warp] = inclA; scanTmpB[warp] = inclB; }
    __syncthreads();
    unsigned wa = 0, wb = 0;
    for (int w = 0; w < warp; ++w) { wa += scanTmp[w]; wb += scanTmpB[w]; }
    binStart[tid] = wa + inclA - total;
    const unsigned binBase = wb + inclB - myTotalB;

    // decoupled look-back over earlier tiles for this thread's digit
    // (eight predecessors are fetched per round so that the walk costs one L2 round trip per eight tiles)
    unsigned excl = 0;
    for (int t = tile - 1; t >= 0;) {
        constexpr int LOOK = 8;
        unsigned v[LOOK];
#pragma unroll
        for (int k = 0; k < LOOK; ++k) v[k] = (t - k >= 0) ? status[(size_t)(t - k) * 256 + tid] : STATUS_INCL;
        bool done = false;
#pragma unroll
        for (int k = 0; k < LOOK; ++k) {
            if (done) break;
            unsigned x = v[k];
            while ((x >> 30) == 0u) x = status[(size_t)(t - k) * 256 + tid];
            excl += x & STATUS_VALUE;
            done = (x >> 30) == 2u;
        }
        if (done) break;
        t -= LOOK;
    }
    status[(size_t)tile * 256 + tid] = STATUS_INCL | (excl + total);
    globalBase[tid] = binBase + excl;
    __syncthreads();

#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int idx = tileStart + warp * (32 * SORT_ITEMS) + i * 32 + lane;
        if (idx < n) {
            const unsigned digit = (unsigned)(key[i] >> shift) & 255u;
            const unsigned pos = binStart[digit] + warpCnt[warp][digit] + rank[i];
            exKeys[pos] = key[i];
            exVals[pos] = val[i];
        }
    }
    __syncthreads();

    const int tileCount = min(SORT_TILE, n - tileStart);
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int j = i * SORT_THREADS + tid;
        if (j < tileCount) {
            const int k = exKeys[j];
            const unsigned digit = (unsigned)(k >> shift) & 255u;
            const unsigned dst = globalBase[digit] + (unsigned)j - binStart[digit];
            keysOut[dst] = k;
            valsOut[dst] = exVals[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cell tables + reorder
// ------------------------------------------------------------------------------------------------
// DENSE tables (reference-compatible semantics and the small static triangle grid): cellStart[c] /
// cellEnd[c] for every cell of the grid, as in the reference (uniform_grid.cu:51-80).
__global__ void __launch_bounds__(256) clear_cells_kernel(const int* __restrict__ sortedKeys, int n, int* __restrict__ cellStart,
                                                          int* __restrict__ cellEnd)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int key = sortedKeys[slot];
    if (slot == 0 || key != sortedKeys[slot - 1]) {
        cellStart[key] = 0;
        cellEnd[key] = -1;
    }
}

__device__ __forceinline__ void reorder_slot(int slot, int id, bool reference, const float4* __restrict__ pos,
                                             const float4* __restrict__ vel, float4* __restrict__ spos, float4* __restrict__ svel)
{
    // slab mode tags ghost particles in bit 31 of the sorted value; the tag travels on in svel.w
    const int pid = id & 0x7fffffff;
    float4 p = pos[pid];
    float4 v = vel[pid];
    // canonical pos4.w already carries the particle's own collision radius (set at upload, preserved by
    // the integrator); the reference-compatible radius lookup needs the particle id instead
    if (reference) p.w = __int_as_float(id);
    v.w = __int_as_float(id);
    spos[slot] = p;
    svel[slot] = v;
}

template <bool REFERENCE, bool REORDER>
__global__ void __launch_bounds__(256) finalize_grid_kernel(const int* __restrict__ keys, const int* __restrict__ ids, int n,
                                                            int* __restrict__ cellStart, int* __restrict__ cellEnd,
                                                            const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                            float4* __restrict__ spos, float4* __restrict__ svel)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int key = keys[slot];
    const int prev = slot > 0 ? keys[slot - 1] : -1;
    const int next = slot < n - 1 ? keys[slot + 1] : -1;
    if (REFERENCE) {
        // uniform_grid.cu:51-80 as written, including thread N-1's `cellStarts[...] = N-1` (:76-79).  That
        // stray write races with the true start of the last occupied cell; it is issued later in program
        // order, so it is made the winner here deterministically.
        const int lastKey = keys[n - 1];
        if (slot > 0 && key != prev && key != lastKey) cellStart[key] = slot;
        if (slot < n - 1 && key != next) cellEnd[key] = slot;
        if (slot == 0 && key != lastKey) cellStart[key] = 0;
        if (slot == n - 1) cellStart[key] = n - 1;
    } else {
        if (key != prev) cellStart[key] = slot;
        if (key != next) cellEnd[key] = slot;
    }
    if (REORDER) reorder_slot(slot, ids[slot], REFERENCE, pos, vel, spos, svel);
}

// COMPACT cell index (clean semantics, particle grid).  A blood-cell scene occupies ~1-15 % of its grid
// cells, and the reference's dense tables (2 ints per cell: 55 MB for the default vein, 490 MB for the
// 1 M-particle long vein) would make every neighbour lookup a DRAM sector fetch.  Instead:
//   cellMask[c/32]   bit c%32 set  <=> cell c holds particles            (1 bit per cell, L2 resident)
//   cellRank[c/32]   number of occupied cells before word c/32           (valid where the word is non-zero)
//   occStart[r]      first sorted slot of the r-th occupied cell; occStart[numOcc] = N
//   occKey[r]        cell id of the r-th occupied cell
// Cell ids are x-fastest, so the <= 3 cells of one stencil row are adjacent bits, their occupied cells
// have consecutive ranks and their particles form one contiguous slot range:
//   [ occStart[rank(first)], occStart[rank(first) + popc(bits)] ).
// Same cell ids, same sorted order, same candidate sets as the dense tables.
constexpr int FIN_THREADS = 256;
constexpr int FIN_ITEMS = 4;
constexpr int FIN_TILE = FIN_THREADS * FIN_ITEMS;

__global__ void __launch_bounds__(FIN_THREADS) count_cell_starts_kernel(const int* __restrict__ keys, int nArg, const int* __restrict__ nDev,
                                                                        int* __restrict__ tileCount)
{
    const int n = nDev ? *nDev : nArg;
    __shared__ int warpSum[FIN_THREADS / 32];
    const int base = blockIdx.x * FIN_TILE + threadIdx.x * FIN_ITEMS;
    int c = 0;
    int prev = (base > 0 && base - 1 < n) ? keys[base - 1] : -1;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k) {
        const int slot = base + k;
        if (slot < n) {
            const int key = keys[slot];
            c += (slot == 0 || key != prev);
            prev = key;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) s += warpSum[w];
        tileCount[blockIdx.x] = s;
    }
}

template <bool REORDER>
__global__ void __launch_bounds__(FIN_THREADS) finalize_compact_kernel(const int* __restrict__ keys, const int* __restrict__ ids, int nArg,
                                                                       const int* __restrict__ nDev, const int* __restrict__ tileCount, unsigned* __restrict__ cellMask,
                                                                       int* __restrict__ cellRank, int* __restrict__ occStart,
                                                                       int* __restrict__ occKey, int* __restrict__ numOcc,
                                                                       const float4* __restrict__ pos, const float4* __restrict__ vel,
                                                                       float4* __restrict__ spos, float4* __restrict__ svel)
{
    __shared__ int warpSum[FIN_THREADS / 32];
    __shared__ int tileBase;
    const int n = nDev ? *nDev : nArg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n == 0) {
        if (blockIdx.x == 0 && tid == 0) { occStart[0] = 0; *numOcc = 0; }
        return;
    }
    if ((int)blockIdx.x * FIN_TILE >= n) return;
    // occupied cells before this tile
    int acc = 0;
    for (int b = tid; b < (int)blockIdx.x; b += FIN_THREADS) acc += tileCount[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) warpSum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) s += warpSum[w];
        tileBase = s;
    }
    __syncthreads();

    const int base = blockIdx.x * FIN_TILE + tid * FIN_ITEMS;
    int key[FIN_ITEMS];
    bool start[FIN_ITEMS];
    int mine = 0;
    int prev = (base > 0 && base - 1 < n) ? keys[base - 1] : -1;
    const int prevOfFirst = prev;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k) {
        const int slot = base + k;
        key[k] = slot < n ? keys[slot] : -1;
        start[k] = slot < n && (slot == 0 || key[k] != prev);
        mine += start[k];
        if (slot < n) prev = key[k];
    }
    // exclusive scan of `mine` over the block (thread order == slot order)
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warpSum[w];
    int rank = tileBase + wbase + incl - mine;   // rank of the first cell STARTING in this thread's slots

    prev = prevOfFirst;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k) {
        const int slot = base + k;
        if (slot >= n) break;
        if (start[k]) {
            const int c = key[k];
            atomicOr(&cellMask[c >> 5], 1u << (c & 31));
            occStart[rank] = slot;
            occKey[rank] = c;
            if (slot == 0 || (prev >> 5) != (c >> 5)) cellRank[c >> 5] = rank;   // first occupied cell of its word
            ++rank;
        }
        if (slot == n - 1) {
            occStart[rank] = n;     // rank == number of occupied cells here
            *numOcc = rank;
        }
        prev = key[k];
        if (REORDER) reorder_slot(slot, ids[slot], false, pos, vel, spos, svel);
    }
}

// ------------------------------------------------------------------------------------------------
// Occupancy-rank counting sort (clean semantics, particle grid; BCS_GRID=radix selects the radix path above).
// The compact cell index already knows the sorted order of the occupied cells: rank(cell) = number of set
// occupancy bits before it.  So the (cell id, particle id) order needs no key sort at all:
//   1 mark      key per particle, occupancy bit set (atomicOr)
//   2 rank      exclusive prefix popcount over the mask words            -> cellRank, numOcc
//   3 count     per particle: r = rank(cell); place = atomicAdd(count[r]) (arbitrary order inside a cell)
//   4 starts    exclusive scan of count[0..numOcc)                       -> occStart
//   5 scatter   tmp[occStart[r] + place] = particle                      (cells contiguous, unordered inside)
//   6 order     per slot: position inside its cell = #ids of the cell smaller than its own (cells hold ~1-12
//               particles) -> final slot; write keys / ids / occKey and gather pos+vel into slot order
// Every pass is a plain streaming kernel (no look-back chains, no tile tickets), which is what matters when a
// rank of the slab decomposition holds only ~100 k particles.  Outputs are identical to the radix path.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_sum(int v, int* warpSum)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = v;
    __syncthreads();
    int s = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) s += warpSum[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(256) cs_mark_kernel(const float4* __restrict__ pos, const unsigned char* __restrict__ pflag, GridDev g,
                                                      int* __restrict__ keyOf, unsigned* __restrict__ cellMask, Counters* __restrict__ counters,
                                                      const ActiveItems act)
{
    // slab mode: the active particles are enumerated through the rank's lists; the launch is a bounded grid that strides
    // over the device-side item count, so the pass costs O(local particles) whatever the global count
    const int total = act.cells ? item_total(act) : g.n;
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x, f = 1;
        if (act.cells) {
            if (i >= total) continue;
            i = active_item(act, i, f);
            if (i < 0) continue;
        } else {
            if (i >= g.n) continue;
            f = pflag ? pflag[i] : 1;
            if (!f) continue;
        }
        const float4 p = pos[i];
        const bool oob = p.x < g.minx || p.x > g.maxx || p.y < g.miny || p.y > g.maxy || p.z < g.minz || p.z > g.maxz;
        int key = axis_cell(p.z, g.minz, g.lenz, g.csz) * g.nx * g.ny + axis_cell(p.y, g.miny, g.leny, g.csy) * g.nx +
                  axis_cell(p.x, g.minx, g.lenx, g.csx);
        if (oob) {
            if (f & 1) atomicAdd(&counters->oob, 1ull);
            key = max(0, min(key, g.cells - 1));
        } else if (key >= g.cells) {
            key = g.cells - 1;
        }
        keyOf[i] = key;
        const unsigned bit = 1u << (key & 31);
        if (!(cellMask[key >> 5] & bit)) atomicOr(&cellMask[key >> 5], bit);
    }
}

// per-tile totals of popc(mask word) (MODE 0) or of the per-cell counts (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) cs_tile_totals_kernel(const unsigned* __restrict__ in, int nArg, const int* __restrict__ nDev,
                                                                      int* __restrict__ tileTotal)
{
    __shared__ int warpSum[SCAN_THREADS / 32];
    const int n = nDev ? *nDev : nArg;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int c = 0;
    if (base + SCAN_ITEMS <= n) {
        const uint4* in4 = reinterpret_cast<const uint4*>(in + base);   // base is a multiple of 16 words
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            const uint4 q = in4[k];
            c += MODE == 0 ? __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w) : (int)(q.x + q.y + q.z + q.w);
        }
    } else if (base < n) {
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < n) c += MODE == 0 ? __popc(in[base + k]) : (int)in[base + k];
    }
    const int s = block_sum(c, warpSum);
    if (threadIdx.x == 0) tileTotal[blockIdx.x] = s;
}

// exclusive scan: out[i] = sum of f(in[j]) for j < i; out[n] (MODE 1) / *total = grand total
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) cs_scan_kernel(const unsigned* __restrict__ in, int nArg, const int* __restrict__ nDev,
                                                               const int* __restrict__ tileTotal, int* __restrict__ out, int* __restrict__ total,
                                                               int finalValue, const int* __restrict__ finalValueDev)
{
    __shared__ int warpSum[SCAN_THREADS / 32];
    __shared__ int sBase;
    const int n = nDev ? *nDev : nArg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int)blockIdx.x * SCAN_TILE >= n && !(blockIdx.x == 0)) return;
    int acc = 0;
    for (int b = tid; b < (int)blockIdx.x; b += SCAN_THREADS) acc += tileTotal[b];
    const int tileBase = block_sum(acc, warpSum);
    const int base = blockIdx.x * SCAN_TILE + tid * SCAN_ITEMS;
    int v[SCAN_ITEMS], mine = 0;
    if (base + SCAN_ITEMS <= n) {
        const uint4* in4 = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            const uint4 q = in4[k];
            v[4 * k] = MODE == 0 ? __popc(q.x) : (int)q.x; v[4 * k + 1] = MODE == 0 ? __popc(q.y) : (int)q.y;
            v[4 * k + 2] = MODE == 0 ? __popc(q.z) : (int)q.z; v[4 * k + 3] = MODE == 0 ? __popc(q.w) : (int)q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? (MODE == 0 ? __popc(in[base + k]) : (int)in[base + k]) : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) mine += v[k];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warpSum[w];
    int run = tileBase + wbase + incl - mine;
    if (base + SCAN_ITEMS <= n) {
        int4* out4 = reinterpret_cast<int4*>(out + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
            int4 o;
            o.x = run; run += v[4 * k]; o.y = run; run += v[4 * k + 1]; o.z = run; run += v[4 * k + 2]; o.w = run; run += v[4 * k + 3];
            out4[k] = o;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < n) out[base + k] = run;
            run += v[k];
        }
    }
    // the thread holding element n-1 (or thread 0 of block 0 when n == 0) publishes the grand total
    const bool last = (n > 0) ? (base <= n - 1 && n - 1 < base + SCAN_ITEMS) : (blockIdx.x == 0 && tid == 0);
    if (last) {
        if (total) *total = run;
        if (MODE == 1) out[n] = total ? run : (finalValueDev ? *finalValueDev : finalValue);   // occStart[numOcc] = number of sorted slots (= the grand total when that is asked for)
    }
    (void)sBase;
}

// Single-launch exclusive scan (same outputs as cs_tile_totals + cs_scan).  Tiles are handed out by ticket, so every
// tile a block waits for has already started; a block publishes its tile total BEFORE it looks back, so the totals
// do not chain: the wait is one publish + one read whatever the tile number.  Status words carry (epoch << 32 | total);
// the last block to leave bumps the epoch and rewinds the ticket, so nothing is cleared between launches (graph replay).
struct ScanCtl {
    unsigned ticket, done, epoch, pad;
};
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) cs_scan_fused_kernel(const unsigned* __restrict__ in, int nArg, const int* __restrict__ nDev,
                                                                     unsigned long long* __restrict__ status, ScanCtl* __restrict__ ctl,
                                                                     int* __restrict__ out, int* __restrict__ total, int finalValue,
                                                                     const int* __restrict__ finalValueDev)
{
    __shared__ int warpSum[SCAN_THREADS / 32];
    __shared__ unsigned sTile, sEpoch;
    const int n = nDev ? *nDev : nArg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        sEpoch = *reinterpret_cast<volatile unsigned*>(&ctl->epoch) + 1u;   // epochs start at 1: zero-filled status = never published
        sTile = atomicAdd(&ctl->ticket, 1u);
    }
    __syncthreads();
    const int tile = (int)sTile;
    const unsigned epoch = sEpoch;
    const bool active = tile * SCAN_TILE < n || tile == 0;
    if (active) {
        const int base = tile * SCAN_TILE + tid * SCAN_ITEMS;
        int v[SCAN_ITEMS], mine = 0;
        if (base + SCAN_ITEMS <= n) {
            const uint4* in4 = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
                const uint4 q = in4[k];
                v[4 * k] = MODE == 0 ? __popc(q.x) : (int)q.x; v[4 * k + 1] = MODE == 0 ? __popc(q.y) : (int)q.y;
                v[4 * k + 2] = MODE == 0 ? __popc(q.z) : (int)q.z; v[4 * k + 3] = MODE == 0 ? __popc(q.w) : (int)q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? (MODE == 0 ? __popc(in[base + k]) : (int)in[base + k]) : 0;
        }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) mine += v[k];
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        int wbase = 0, tileTotal = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            if (w < warp) wbase += warpSum[w];
            tileTotal += warpSum[w];
        }
        if (tid == 0) {
            const unsigned long long word = ((unsigned long long)epoch << 32) | (unsigned)tileTotal;
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(status + tile), "l"(word) : "memory");
        }
        // look back: totals of all earlier tiles (each published without waiting for anybody)
        int acc = 0;
        for (int b = tid; b < tile; b += SCAN_THREADS) {
            unsigned long long word;
            do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(word) : "l"(status + b) : "memory");
            } while ((unsigned)(word >> 32) != epoch);
            acc += (int)(unsigned)word;
        }
        __syncthreads();   // warpSum is reused by block_sum
        const int tileBase = block_sum(acc, warpSum);
        int run = tileBase + wbase + incl - mine;
        if (base + SCAN_ITEMS <= n) {
            int4* out4 = reinterpret_cast<int4*>(out + base);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS / 4; ++k) {
                int4 o;
                o.x = run; run += v[4 * k]; o.y = run; run += v[4 * k + 1]; o.z = run; run += v[4 * k + 2]; o.w = run; run += v[4 * k + 3];
                out4[k] = o;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (base + k < n) out[base + k] = run;
                run += v[k];
            }
        }
        const bool last = (n > 0) ? (base <= n - 1 && n - 1 < base + SCAN_ITEMS) : (tile == 0 && tid == 0);
        if (last) {
            if (total) *total = run;
            if (MODE == 1) out[n] = total ? run : (finalValueDev ? *finalValueDev : finalValue);   // occStart[numOcc] = number of sorted slots (= the grand total when that is asked for)
        }
    }
    // the last block to leave prepares the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&ctl->done, 1u) == gridDim.x - 1u) {
            ctl->ticket = 0u;
            ctl->done = 0u;
            ctl->epoch = epoch;
            __threadfence();
        }
    }
}

__global__ void __launch_bounds__(256) cs_count_kernel(const unsigned char* __restrict__ pflag, int n, const int* __restrict__ keyOf,
                                                       const unsigned* __restrict__ cellMask, const int* __restrict__ cellRank,
                                                       unsigned* __restrict__ cellCount, int* __restrict__ rankOf, int* __restrict__ placeOf,
                                                       int* __restrict__ nActive, const ActiveItems items)
{
    const int total = items.cells ? item_total(items) : n;
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x;
        bool act;
        if (items.cells) {
            int f;
            i = i < total ? active_item(items, i, f) : -1;
            act = i >= 0;
        } else {
            act = i < n && (pflag ? pflag[i] != 0 : true);
        }
        if (act) {
            const int key = keyOf[i];
            const int r = cellRank[key >> 5] + __popc(cellMask[key >> 5] & ((1u << (key & 31)) - 1u));
            rankOf[i] = r;
            placeOf[i] = (int)atomicAdd(&cellCount[r], 1u);
        }
        if (nActive) {
            // number of active particles (slab mode): one atomic per CTA and round
            const int blockActive = __syncthreads_count(act);
            if (threadIdx.x == 0 && blockActive) atomicAdd(nActive, blockActive);
        }
    }
}

__global__ void __launch_bounds__(256) cs_scatter_kernel(const unsigned char* __restrict__ pflag, int n, const int* __restrict__ rankOf,
                                                         const int* __restrict__ placeOf, const int* __restrict__ occStart,
                                                         int* __restrict__ tmpIds, int* __restrict__ tmpRank, const ActiveItems items)
{
    const int total = items.cells ? item_total(items) : n;
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x, f = 1;
        if (items.cells) {
            if (i >= total) continue;
            i = active_item(items, i, f);
            if (i < 0) continue;
        } else {
            if (i >= n) continue;
            f = pflag ? pflag[i] : 1;
            if (!f) continue;
        }
        const int r = rankOf[i];
        const int slot = occStart[r] + placeOf[i];
        tmpIds[slot] = (f & 1) ? i : (i | (int)0x80000000);   // ghost tag (slab mode)
        tmpRank[slot] = r;
    }
}

template <bool REORDER>
__global__ void __launch_bounds__(256) cs_order_kernel(int nArg, const int* __restrict__ nDev, const int* __restrict__ tmpIds,
                                                       const int* __restrict__ tmpRank, const int* __restrict__ occStart, const int* __restrict__ keyOf,
                                                       int* __restrict__ keys, int* __restrict__ ids, int* __restrict__ occKey,
                                                       const float4* __restrict__ pos, const float4* __restrict__ vel, float4* __restrict__ spos,
                                                       float4* __restrict__ svel, unsigned* __restrict__ cellCount)
{
    const int n = nDev ? *nDev : nArg;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int r = tmpRank[j];
        const int s = occStart[r], e = occStart[r + 1];
        const int tag = tmpIds[j], pid = tag & 0x7fffffff;
        int before = 0;
        for (int k = s; k < e; ++k) before += (tmpIds[k] & 0x7fffffff) < pid;
        const int slot = s + before;
        const int key = keyOf[pid];
        keys[slot] = key;
        ids[slot] = tag;
        if (before == 0) {
            occKey[r] = key;
            cellCount[r] = 0u;   // consumed by the start scan; clean for the next build (no per-step memset)
        }
        if (REORDER) reorder_slot(slot, tag, false, pos, vel, spos, svel);
    }
}

// ------------------------------------------------------------------------------------------------
// Row-directory grid build (RowsGrid, kernels.cuh): count -> scan -> scatter -> order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_count_kernel(const float4* __restrict__ pos, const unsigned char* __restrict__ pflag, const GridDev g,
                                                        const RowsGrid R, Counters* __restrict__ counters, int* __restrict__ nActive,
                                                        const ActiveItems act)
{
    const int total = act.cells ? item_total(act) : g.n;
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x, f = 1;
        bool on;
        if (act.cells) {
            i = i < total ? active_item(act, i, f) : -1;
            on = i >= 0;
        } else {
            on = i < g.n;
            if (on && pflag) { f = pflag[i]; on = f != 0; }
        }
        if (on) rows_count_particle(g, R, pos[i], i, f, counters);
        if (nActive) {
            const int blockActive = __syncthreads_count(on);
            if (threadIdx.x == 0 && blockActive) atomicAdd(nActive, blockActive);
        }
    }
}

__global__ void __launch_bounds__(256) row_scatter_kernel(const unsigned char* __restrict__ pflag, int n, const RowsGrid R, const ActiveItems act)
{
    const int total = act.cells ? item_total(act) : n;
    for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        int i = base + threadIdx.x, f = 1;
        if (act.cells) {
            if (i >= total) continue;
            i = active_item(act, i, f);
            if (i < 0) continue;
        } else {
            if (i >= n) continue;
            if (pflag) { f = pflag[i]; if (!f) continue; }
        }
        const int key = R.keyOf[i];
        const int row = rows_of_key(R, key);
        // the row's cursor is the scan's output for it (rowStart[row + 1]); bumped once per member it ends as the row's end
        const int slot = atomicAdd(&R.rowStart[row + 1], 1);
        R.tmp[slot] = make_int2(key, (f & 1) ? i : (i | (int)0x80000000));
    }
}

// order: final slot = row start + number of row mates that sort before (cell id, particle id).  Also the near-wall probe
// of the wall search (NearProbe, kernels.cuh) when the step is fused: every position passes through here anyway.
template <bool PROBE>
__global__ void __launch_bounds__(256) row_order_kernel(int nArg, const int* __restrict__ nDev, const RowsGrid R, int* __restrict__ keys,
                                                        int* __restrict__ ids, const float4* __restrict__ pos, float4* __restrict__ spos,
                                                        const NearProbe probe)
{
    __shared__ int sNear, sBase;
    if (PROBE) {
        if (threadIdx.x == 0) sNear = 0;
        __syncthreads();
    }
    const int n = nDev ? *nDev : nArg;
    // latch the "irregular particle" flag of this build (the count pass of the NEXT build may already run - fused into
    // the cell pass - before this build's collision stages are through with it)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        R.irregular[1] = R.irregular[0]; R.irregular[0] = 0;
        // sentinels: the pair search scans keys until one exceeds its window
        keys[n] = 0x7fffffff; keys[n + 1] = 0x7fffffff; keys[n + 2] = 0x7fffffff; keys[n + 3] = 0x7fffffff;
    }
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int j = base + threadIdx.x;
        bool isNear = false;
        int pidOut = 0;
        if (j < n) {
            const int2 me = R.tmp[j];
            const int key = me.x, tag = me.y, pid = tag & 0x7fffffff;
            const float4 p = pos[pid];
            const int row = rows_of_key(R, key);
            const int s = R.rowStart[row], e = R.rowStart[row + 1];
            int before = 0;
            for (int k = s; k < e; ++k) {
                const int2 o = R.tmp[k];
                before += (o.x < key) || (o.x == key && (o.y & 0x7fffffff) < pid);
            }
            const int slot = s + before;
            keys[slot] = key;
            ids[slot] = tag;
            spos[slot] = p;
            if (j == s) R.rowCount[row] = 0u;   // consumed by the scan; clean for the next count pass
            if (PROBE && tag >= 0) {
                const int hx = (int)fminf(fmaxf(floorf((p.x - probe.ox) * probe.invh), 0.f), (float)(probe.nx - 1));
                const int hy = (int)fminf(fmaxf(floorf((p.y - probe.oy) * probe.invh), 0.f), (float)(probe.ny - 1));
                const int hz = (int)fminf(fmaxf(floorf((p.z - probe.oz) * probe.invh), 0.f), (float)(probe.nz - 1));
                isNear = __ldg(probe.near + (hz * probe.ny + hy) * probe.nx + hx) != 0;
                pidOut = pid;
            }
        }
        if (PROBE) {
            // CTA-aggregated append: one global atomic per CTA and round (a third of the particles are near the wall, so a
            // per-warp atomic would put ~30 000 operations per step on one address)
            const unsigned m = __ballot_sync(0xffffffffu, isNear);
            const int lane = threadIdx.x & 31;
            int local = 0;
            if (m && lane == 0) local = atomicAdd(&sNear, __popc(m));
            local = __shfl_sync(0xffffffffu, local, 0);
            __syncthreads();
            if (threadIdx.x == 0) {
                sBase = sNear ? atomicAdd(probe.count, sNear) : 0;
                sNear = 0;
            }
            __syncthreads();
            if (isNear) probe.list[sBase + local + __popc(m & ((1u << lane) - 1u))] = pidOut;
            __syncthreads();   // sBase is rewritten in the next round
        }
    }
}

// ------------------------------------------------------------------------------------------------
// slab mode: cell keys of the active particles (pflag bit 0 = owned, bit 1 = ghost), compacted in ascending
// particle id so that the stable sort still ends in (cell id, particle id) order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FIN_THREADS) slab_count_active_kernel(const unsigned char* __restrict__ pflag, int n, int* __restrict__ tileCount)
{
    __shared__ int warpSum[FIN_THREADS / 32];
    const int base = blockIdx.x * FIN_TILE + threadIdx.x * FIN_ITEMS;
    int c = 0;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k)
        if (base + k < n) c += pflag[base + k] != 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) s += warpSum[w];
        tileCount[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(FIN_THREADS) slab_keys_kernel(const float4* __restrict__ pos, const unsigned char* __restrict__ pflag, GridDev g,
                                                                const int* __restrict__ tileCount, int* __restrict__ keys, int* __restrict__ ids,
                                                                unsigned* __restrict__ digitTotals, int passes, int* __restrict__ nActive,
                                                                Counters* __restrict__ counters)
{
    __shared__ unsigned hist[4][256];
    __shared__ int warpSum[FIN_THREADS / 32];
    __shared__ int tileBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4 * 256; i += FIN_THREADS) (&hist[0][0])[i] = 0;
    int acc = 0;
    for (int b = tid; b < (int)blockIdx.x; b += FIN_THREADS) acc += tileCount[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) warpSum[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) s += warpSum[w];
        tileBase = s;
    }
    __syncthreads();
    const int base = blockIdx.x * FIN_TILE + tid * FIN_ITEMS;
    unsigned char f[FIN_ITEMS];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k) {
        f[k] = (base + k < g.n) ? pflag[base + k] : 0;
        mine += f[k] != 0;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads();
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warpSum[w];
    int out = tileBase + wbase + incl - mine;
#pragma unroll
    for (int k = 0; k < FIN_ITEMS; ++k) {
        if (!f[k]) continue;
        const int i = base + k;
        const float4 p = pos[i];
        const bool oob = p.x < g.minx || p.x > g.maxx || p.y < g.miny || p.y > g.maxy || p.z < g.minz || p.z > g.maxz;
        int key = axis_cell(p.z, g.minz, g.lenz, g.csz) * g.nx * g.ny + axis_cell(p.y, g.miny, g.leny, g.csy) * g.nx +
                  axis_cell(p.x, g.minx, g.lenx, g.csx);
        if (oob) {
            if (f[k] & 1) atomicAdd(&counters->oob, 1ull);
            key = max(0, min(key, g.cells - 1));
        } else if (key >= g.cells) {
            key = g.cells - 1;
        }
        keys[out] = key;
        ids[out] = (f[k] & 1) ? i : (i | (int)0x80000000);   // ghost tag
        ++out;
        for (int p2 = 0; p2 < passes; ++p2) atomicAdd(&hist[p2][(key >> (8 * p2)) & 255], 1u);
    }
    if (blockIdx.x == gridDim.x - 1 && tid == FIN_THREADS - 1) *nActive = out;
    __syncthreads();
    for (int i = tid; i < passes * 256; i += FIN_THREADS) {
        const unsigned v = (&hist[0][0])[i];
        if (v) atomicAdd(&digitTotals[i], v);
    }
}

// ------------------------------------------------------------------------------------------------
// host-side driver
// ------------------------------------------------------------------------------------------------
void SortScratch::allocate(int n, int maskWordsHint)
{
    numTiles = (n + SORT_TILE - 1) / SORT_TILE;
    BCS_CUDA(cudaMalloc(&tileHist, (size_t)numTiles * 256 * sizeof(unsigned)));
    BCS_CUDA(cudaMalloc(&digitTotals, (4 * 256 + 4) * sizeof(unsigned)));   // + one tile ticket per pass
    BCS_CUDA(cudaMalloc(&status, (size_t)4 * numTiles * 256 * sizeof(unsigned)));
    const char* mode = getenv("BCS_SORT");
    classic = mode && std::string(mode) == "classic";
    const char* gridMode = getenv("BCS_GRID");
    radixForCompact = gridMode && std::string(gridMode) == "radix";
    // counting-sort scratch
    BCS_CUDA(cudaMalloc(&keyOf, (size_t)n * sizeof(int)));
    BCS_CUDA(cudaMalloc(&rankOf, (size_t)n * sizeof(int)));
    BCS_CUDA(cudaMalloc(&placeOf, (size_t)n * sizeof(int)));
    BCS_CUDA(cudaMalloc(&tmpIds, (size_t)n * sizeof(int)));
    BCS_CUDA(cudaMalloc(&tmpRank, (size_t)n * sizeof(int)));
    BCS_CUDA(cudaMalloc(&cellCount, ((size_t)n + 1) * sizeof(unsigned)));
    BCS_CUDA(cudaMemset(cellCount, 0, ((size_t)n + 1) * sizeof(unsigned)));   // kept clean by cs_order_kernel from then on
    {
        const size_t tiles = (size_t)(std::max(n, maskWordsHint) / SCAN_TILE + 2);
        BCS_CUDA(cudaMalloc(&scanStatus, tiles * sizeof(unsigned long long)));
        BCS_CUDA(cudaMemset(scanStatus, 0, tiles * sizeof(unsigned long long)));
        BCS_CUDA(cudaMalloc(&scanCtl, 4 * sizeof(unsigned)));
        BCS_CUDA(cudaMemset(scanCtl, 0, 4 * sizeof(unsigned)));
        BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));   // legacy-stream fills vs the handle's non-blocking stream
        const char* scanMode = getenv("BCS_SCAN");
        twoPassScan = !(scanMode && std::string(scanMode) == "fused");   // measured: 2 x (totals + scan) 18 us vs 22 us fused at 1 M particles
    }
    BCS_CUDA(cudaMalloc(&scanTotals, (size_t)(std::max(n, maskWordsHint) / SCAN_TILE + 2) * sizeof(int)));
    BCS_CUDA(cudaMalloc(&finTileCount, (size_t)((n + FIN_TILE - 1) / FIN_TILE + 1) * sizeof(int)));
}
void SortScratch::release()
{
    cudaFree(tileHist);
    cudaFree(digitTotals);
    cudaFree(finTileCount);
    cudaFree(status);
    cudaFree(keyOf); cudaFree(rankOf); cudaFree(placeOf); cudaFree(tmpIds); cudaFree(tmpRank); cudaFree(cellCount); cudaFree(scanTotals); cudaFree(scanStatus); cudaFree(scanCtl);
    scanStatus = nullptr; scanCtl = nullptr;
    keyOf = rankOf = placeOf = tmpIds = tmpRank = scanTotals = nullptr;
    cellCount = nullptr;
    status = nullptr;
    tileHist = digitTotals = nullptr;
    finTileCount = nullptr;
}

void launch_row_count(const GridBuildArgs& a, cudaStream_t st)
{
    const GridDev& g = a.grid;
    const int blocks = (g.n + 255) / 256;
    const int itemBlocks = a.items.cells ? (int)std::min<long long>(((long long)a.itemCapacity + 255) / 256, BOUNDED_BLOCKS) : blocks;
    if (a.nDevOut) BCS_CUDA(cudaMemsetAsync(a.nDevOut, 0, sizeof(int), st));
    BCS_LAUNCH("cell_keys", st, row_count_kernel<<<itemBlocks, 256, 0, st>>>(a.objPos, a.pflag, g, a.rows, a.counters, a.nDevOut, a.items));
    BCS_CUDA(cudaGetLastError());
}

static void launch_grid_build_rows(const GridBuildArgs& a, cudaStream_t st)
{
    const GridDev& g = a.grid;
    const RowsGrid& R = a.rows;
    const int n = g.n, blocks = (n + 255) / 256;
    const int itemBlocks = a.items.cells ? (int)std::min<long long>(((long long)a.itemCapacity + 255) / 256, BOUNDED_BLOCKS) : blocks;
    const int orderBlocks = std::max(1, a.nDev ? std::min(blocks, BOUNDED_BLOCKS) : blocks);
    SortScratch* sc = a.scratch;
    if (!R.countDone) launch_row_count(a, st);
    // rowStart = exclusive scan of the per-row counts; rowStart[nRows] = number of sorted slots
    // (slab mode, count fused into the previous step: nobody has counted the active particles - the scan's total is it)
    int* const total = (R.countDone && a.nDevOut) ? a.nDevOut : nullptr;
    const int tiles = (R.nRows + SCAN_TILE - 1) / SCAN_TILE;
    ScanCtl* ctl = reinterpret_cast<ScanCtl*>(sc->scanCtl);
    if (sc->twoPassScan) {
        BCS_LAUNCH("row_start_totals", st, cs_tile_totals_kernel<1><<<tiles, SCAN_THREADS, 0, st>>>(R.rowCount, R.nRows, nullptr, sc->scanTotals));
        BCS_LAUNCH("row_start_scan", st,
                   cs_scan_kernel<1><<<tiles, SCAN_THREADS, 0, st>>>(R.rowCount, R.nRows, nullptr, sc->scanTotals, R.rowStart + 1, total, n, a.nDev));
    } else {
        BCS_LAUNCH("row_start_scan", st,
                   cs_scan_fused_kernel<1><<<tiles, SCAN_THREADS, 0, st>>>(R.rowCount, R.nRows, nullptr, sc->scanStatus, ctl, R.rowStart + 1, total, n, a.nDev));
    }
    BCS_LAUNCH("row_scatter", st, row_scatter_kernel<<<itemBlocks, 256, 0, st>>>(a.pflag, n, R, a.items));
    if (a.probe && a.probe->near)
        BCS_LAUNCH("finalize_grid", st, row_order_kernel<true><<<orderBlocks, 256, 0, st>>>(n, a.nDev, R, a.keys[1], a.ids[1], a.pos, a.spos, *a.probe));
    else
        BCS_LAUNCH("finalize_grid", st, row_order_kernel<false><<<orderBlocks, 256, 0, st>>>(n, a.nDev, R, a.keys[1], a.ids[1], a.pos, a.spos, NearProbe{}));
    BCS_CUDA(cudaGetLastError());
}

static void launch_grid_build_counting(const GridBuildArgs& a, cudaStream_t st)
{
    const GridDev& g = a.grid;
    const int n = g.n, blocks = (n + 255) / 256;
    // slab mode enumerates owned cells x maxP (padding included) + ghosts: launched for that capacity, CTAs past the
    // device-side item count leave at once
    const int itemBlocks = a.items.cells ? (int)std::min<long long>(((long long)a.itemCapacity + 255) / 256, BOUNDED_BLOCKS) : blocks;
    const int orderBlocks = a.nDev ? std::min(blocks, BOUNDED_BLOCKS) : blocks;
    SortScratch* sc = a.scratch;
    BCS_CUDA(cudaMemsetAsync(a.cellMask, 0, (size_t)a.maskWords * sizeof(unsigned), st));
    if (a.nDevOut) BCS_CUDA(cudaMemsetAsync(a.nDevOut, 0, sizeof(int), st));
    BCS_LAUNCH("cell_keys", st, cs_mark_kernel<<<itemBlocks, 256, 0, st>>>(a.objPos, a.pflag, g, sc->keyOf, a.cellMask, a.counters, a.items));
    const int maskTiles = (a.maskWords + SCAN_TILE - 1) / SCAN_TILE;
    ScanCtl* ctl = reinterpret_cast<ScanCtl*>(sc->scanCtl);
    if (sc->twoPassScan) {
        BCS_LAUNCH("cell_rank_totals", st, cs_tile_totals_kernel<0><<<maskTiles, SCAN_THREADS, 0, st>>>(a.cellMask, a.maskWords, nullptr, sc->scanTotals));
        BCS_LAUNCH("cell_rank_scan", st,
                   cs_scan_kernel<0><<<maskTiles, SCAN_THREADS, 0, st>>>(a.cellMask, a.maskWords, nullptr, sc->scanTotals, a.cellRank, a.numOcc, 0, nullptr));
    } else {
        BCS_LAUNCH("cell_rank_scan", st,
                   cs_scan_fused_kernel<0><<<maskTiles, SCAN_THREADS, 0, st>>>(a.cellMask, a.maskWords, nullptr, sc->scanStatus, ctl, a.cellRank, a.numOcc, 0,
                                                                              nullptr));
    }
    BCS_LAUNCH("cell_count", st,
               cs_count_kernel<<<itemBlocks, 256, 0, st>>>(a.pflag, n, sc->keyOf, a.cellMask, a.cellRank, sc->cellCount, sc->rankOf, sc->placeOf, a.nDevOut,
                                                       a.items));
    const int cntTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (sc->twoPassScan) {
        BCS_LAUNCH("cell_start_totals", st, cs_tile_totals_kernel<1><<<cntTiles, SCAN_THREADS, 0, st>>>(sc->cellCount, n, a.numOcc, sc->scanTotals));
        BCS_LAUNCH("cell_start_scan", st,
                   cs_scan_kernel<1><<<cntTiles, SCAN_THREADS, 0, st>>>(sc->cellCount, n, a.numOcc, sc->scanTotals, a.occStart, nullptr, n, a.nDev));
    } else {
        BCS_LAUNCH("cell_start_scan", st,
                   cs_scan_fused_kernel<1><<<cntTiles, SCAN_THREADS, 0, st>>>(sc->cellCount, n, a.numOcc, sc->scanStatus, ctl, a.occStart, nullptr, n, a.nDev));
    }
    BCS_LAUNCH("cell_scatter", st, cs_scatter_kernel<<<itemBlocks, 256, 0, st>>>(a.pflag, n, sc->rankOf, sc->placeOf, a.occStart, sc->tmpIds, sc->tmpRank, a.items));
    if (a.reorder)
        BCS_LAUNCH("finalize_grid", st,
                   cs_order_kernel<true><<<orderBlocks, 256, 0, st>>>(n, a.nDev, sc->tmpIds, sc->tmpRank, a.occStart, sc->keyOf, a.keys[1], a.ids[1], a.occKey,
                                                                 a.pos, a.vel, a.spos, a.svel, sc->cellCount));
    else
        BCS_LAUNCH("finalize_grid", st,
                   cs_order_kernel<false><<<orderBlocks, 256, 0, st>>>(n, a.nDev, sc->tmpIds, sc->tmpRank, a.occStart, sc->keyOf, a.keys[1], a.ids[1], a.occKey,
                                                                  nullptr, nullptr, nullptr, nullptr, sc->cellCount));
    BCS_CUDA(cudaGetLastError());
}

void launch_grid_build(const GridBuildArgs& a, cudaStream_t st)
{
    if (a.rows.enabled) {
        launch_grid_build_rows(a, st);
        return;
    }
    if (a.compact && !a.scratch->radixForCompact) {
        launch_grid_build_counting(a, st);
        return;
    }
    const GridDev& g = a.grid;
    const int n = g.n;
    const int passes = (g.keyBits + 7) / 8;
    const int blocks = (n + 255) / 256;
    // buffer schedule: the last pass must land in buffer 1
    int cur = (passes & 1) ? 0 : 1;
    if (a.compact) BCS_CUDA(cudaMemsetAsync(a.cellMask, 0, (size_t)a.maskWords * sizeof(unsigned), st));
    else if (!a.reference) BCS_LAUNCH("clear_cells", st, clear_cells_kernel<<<blocks, 256, 0, st>>>(a.keys[1], n, a.cellStart, a.cellEnd));
    BCS_CUDA(cudaMemsetAsync(a.scratch->digitTotals, 0, (4 * 256 + 4) * sizeof(unsigned), st));
    if (!a.scratch->classic)
        BCS_CUDA(cudaMemsetAsync(a.scratch->status, 0, (size_t)passes * a.scratch->numTiles * 256 * sizeof(unsigned), st));
    const int keyBlocks = min(blocks, 148 * 8);
    if (a.pflag) {
        // slab mode: keys of the ACTIVE (owned + ghost) particles only, compacted in ascending particle id
        const int tiles = (n + FIN_TILE - 1) / FIN_TILE;
        BCS_LAUNCH("slab_count_active", st, slab_count_active_kernel<<<tiles, FIN_THREADS, 0, st>>>(a.pflag, n, a.scratch->finTileCount));
        BCS_LAUNCH("cell_keys", st,
                   slab_keys_kernel<<<tiles, FIN_THREADS, 0, st>>>(a.objPos, a.pflag, g, a.scratch->finTileCount, a.keys[cur], a.ids[cur],
                                                                    a.scratch->digitTotals, passes, a.nDevOut, a.counters));
    } else
    BCS_LAUNCH("cell_keys", st,
               cell_keys_kernel<<<keyBlocks, 256, 0, st>>>(a.objPos, g, a.keys[cur], a.ids[cur], a.scratch->digitTotals, passes, a.counters));
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        if (!a.scratch->classic) {
            BCS_LAUNCH("radix_onesweep", st,
                       radix_onesweep_kernel<<<a.scratch->numTiles, SORT_THREADS, 0, st>>>(
                           a.keys[cur], a.ids[cur], a.keys[cur ^ 1], a.ids[cur ^ 1], n, a.nDev, shift, a.scratch->digitTotals + 256 * p,
                           a.scratch->status + (size_t)p * a.scratch->numTiles * 256, a.scratch->digitTotals + 4 * 256 + p));
            cur ^= 1;
            continue;
        }
        BCS_LAUNCH("radix_tile_hist", st,
                   radix_tile_hist_kernel<<<a.scratch->numTiles, SORT_THREADS, 0, st>>>(a.keys[cur], n, shift, a.scratch->tileHist));
        BCS_LAUNCH("radix_scan", st,
                   radix_scan_kernel<<<256, 256, 0, st>>>(a.scratch->tileHist, a.scratch->numTiles, a.scratch->digitTotals + 256 * p));
        BCS_LAUNCH("radix_scatter", st,
                   radix_scatter_kernel<<<a.scratch->numTiles, SORT_THREADS, 0, st>>>(a.keys[cur], a.ids[cur], a.keys[cur ^ 1],
                                                                                      a.ids[cur ^ 1], n, shift, a.scratch->tileHist));
        cur ^= 1;
    }
    // cur == 1 here
    if (a.compact) {
        const int tiles = (n + FIN_TILE - 1) / FIN_TILE;
        BCS_LAUNCH("count_cell_starts", st, count_cell_starts_kernel<<<tiles, FIN_THREADS, 0, st>>>(a.keys[1], n, a.nDev, a.scratch->finTileCount));
        if (a.reorder)
            BCS_LAUNCH("finalize_grid", st,
                       finalize_compact_kernel<true><<<tiles, FIN_THREADS, 0, st>>>(a.keys[1], a.ids[1], n, a.nDev, a.scratch->finTileCount, a.cellMask,
                                                                                     a.cellRank, a.occStart, a.occKey, a.numOcc, a.pos, a.vel,
                                                                                     a.spos, a.svel));
        else
            BCS_LAUNCH("finalize_grid", st,
                       finalize_compact_kernel<false><<<tiles, FIN_THREADS, 0, st>>>(a.keys[1], a.ids[1], n, a.nDev, a.scratch->finTileCount, a.cellMask,
                                                                                      a.cellRank, a.occStart, a.occKey, a.numOcc, nullptr,
                                                                                      nullptr, nullptr, nullptr));
    } else if (a.reference) {
        if (a.reorder)
            BCS_LAUNCH("finalize_grid", st,
                       finalize_grid_kernel<true, true><<<blocks, 256, 0, st>>>(a.keys[1], a.ids[1], n, a.cellStart, a.cellEnd, a.pos, a.vel, a.spos, a.svel));
        else
            BCS_LAUNCH("finalize_grid", st,
                       finalize_grid_kernel<true, false><<<blocks, 256, 0, st>>>(a.keys[1], a.ids[1], n, a.cellStart, a.cellEnd, nullptr, nullptr, nullptr, nullptr));
    } else {
        if (a.reorder)
            BCS_LAUNCH("finalize_grid", st,
                       finalize_grid_kernel<false, true><<<blocks, 256, 0, st>>>(a.keys[1], a.ids[1], n, a.cellStart, a.cellEnd, a.pos, a.vel, a.spos, a.svel));
        else
            BCS_LAUNCH("finalize_grid", st,
                       finalize_grid_kernel<false, false><<<blocks, 256, 0, st>>>(a.keys[1], a.ids[1], n, a.cellStart, a.cellEnd, nullptr, nullptr, nullptr, nullptr));
    }
    BCS_CUDA(cudaGetLastError());
}

}  // namespace bcs

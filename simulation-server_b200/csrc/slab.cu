// Multi-GPU mode: slab decomposition along the vein axis (y) with ghost-particle halo exchange and whole-blood-cell
// migration over NCCL.  One process / one bcs_sim per GPU.
//
// Replaces the reference's -DMULTI_GPU scheme (replicate every array on 4 GPUs, split kernels by INDEX range,
// ncclBroadcast the full state + the 55 MB grid tables and ncclReduce the force arrays every frame:
// main.cu:162-204, blood_cells.cu:190-219, uniform_grid.cu:162-175, vein_triangles.cu:169-195), whose
// per-step traffic grows with the whole state.  Here:
//   * every rank allocates arrays indexed by GLOBAL particle / vertex id (1 M particles = 48 MB: trivial on
//     180 GB), but only advances the blood cells whose centre lies in its y-slab and the vein vertices whose
//     rest position lies in it;
//   * after each step ONE grouped ncclSend/ncclRecv per neighbour carries (a) the blood cells that crossed the
//     face (full state, ownership moves with them), (b) the owned particles within haloWidth of the face
//     (position + velocity: they become ghosts = collision candidates and wall-splat sources on the other
//     side; forces are only ever applied to owned particles, as in particle_collisions.cuh:36), (c) the vein
//     vertices within the vertex halo (position + velocity);  respawned cells (vein_end.cu) are routed to the
//     rank that owns the top of the vein;
//   * messages have a fixed capacity (header + records), so no size handshake and no host synchronisation is
//     needed per step; an overflow raises a sticky error flag checked by bcs_synchronize/bcs_get_stats.
// Cell ids, the (cell id, particle id) order of every rank's sorted list and the candidate sets are those of the
// single-GPU run restricted to the rank's active particles.
#include <nccl.h>

#include "rows_device.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "slab.cuh"

namespace bcs {

#define BCS_NCCL(expr)                                                                                              \
    do {                                                                                                            \
        ncclResult_t _r = (expr);                                                                                   \
        if (_r != ncclSuccess)                                                                                      \
            throw ::bcs::Error{BCS_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_r) + " (" + __FILE__ + ":" + \
                                                 std::to_string(__LINE__) + ")"};                                   \
    } while (0)

__device__ __forceinline__ int cell_of_particle(const TypesDev* __restrict__ types, int i, int* typeOut = nullptr)
{
    int t = 0;
    const int n = types->n;
    while (t + 1 < n && i >= types->t[t + 1].pStart) ++t;
    if (typeOut) *typeOut = t;
    return types->t[t].cStart + (i - types->t[t].pStart) / types->t[t].P;
}

// ownership from the current positions (after the initial upload): a blood cell belongs to the slab its centre is in
__global__ void __launch_bounds__(128) slab_init_ownership_kernel(const TypesDev* __restrict__ types, int nCells, SlabDev slab, const float4* __restrict__ pos,
                                                                 unsigned char* __restrict__ ownedCell, unsigned char* __restrict__ pflag,
                                                                 signed char* __restrict__ moveTo)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    int t = 0;
    while (t + 1 < types->n && c >= types->t[t + 1].cStart) ++t;
    const TypeDev ty = types->t[t];
    const int first = ty.pStart + (c - ty.cStart) * ty.P;
    float cy = 0.f;
    for (int k = 0; k < ty.P; ++k) cy += pos[first + k].y;
    cy /= (float)ty.P;
    const bool mine = cy >= slab.yLo && cy < slab.yHi;
    ownedCell[c] = mine ? 1 : 0;
    moveTo[c] = -1;
    for (int k = 0; k < ty.P; ++k) pflag[first + k] = mine ? 1 : 0;
}

__global__ void __launch_bounds__(256) slab_init_vertices_kernel(int V, SlabDev slab, const float4* __restrict__ vpos,
                                                                unsigned char* __restrict__ vOwned)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float y = vpos[v].y;
    vOwned[v] = (y >= slab.yLo && y < slab.yHi) ? 1 : 0;
}

// ---- owned-cell lists (rebuilt every step after the exchange) ---------------------------------------------------
constexpr int LIST_THREADS = 512;
__global__ void __launch_bounds__(LIST_THREADS) slab_list_cells_kernel(const TypesDev* __restrict__ types, int nCells, const unsigned char* __restrict__ ownedCell,
                                                                      int* __restrict__ cells, int* __restrict__ count, SpringPlan plan,
                                                                      int* __restrict__ blockStart, int* __restrict__ cellPrefix, unsigned* __restrict__ done,
                                                                      SlabBuffers buf, int* __restrict__ keepCount, int nv0, int nv1)
{
    __shared__ int sCount[BCS_MAX_TYPES], sBase[BCS_MAX_TYPES];
    if (threadIdx.x < BCS_MAX_TYPES) sCount[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // this step's messages have been sent and unpacked: rewind the send headers and the keep list for the next pack
        for (int d = 0; d < 3; ++d) { buf.send[d]->nMig = 0; buf.send[d]->nHalo = 0; buf.send[d]->nVerts = d == 0 ? nv0 : d == 1 ? nv1 : 0; }
        *keepCount = 0;
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool own = c < nCells && ownedCell[c];
    int t = 0;
    while (own && t + 1 < types->n && c >= types->t[t + 1].cStart) ++t;
    // append, aggregated twice: one shared-memory atomic per (warp, type), one global atomic per (CTA, type)
    const unsigned active = __ballot_sync(0xffffffffu, own);
    int local = 0;
    if (own) {
        const unsigned peers = __match_any_sync(active, t);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&sCount[t], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        local = base + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x < BCS_MAX_TYPES && sCount[threadIdx.x]) sBase[threadIdx.x] = atomicAdd(&count[threadIdx.x], sCount[threadIdx.x]);
    __syncthreads();
    if (own) cells[types->t[t].cStart + sBase[t] + local] = c;   // type t's segment starts at cStart_t
    // the last CTA to arrive turns the per-type counts into the prefixes the cell-group kernels index by
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        const int nTypes = types->n;
        int bsum = 0, k = 0;
        for (int ty = 0; ty < nTypes; ++ty) {
            const int cnt = atomicAdd(&count[ty], 0);   // the other CTAs' atomics, read where they were performed
            blockStart[ty] = bsum; cellPrefix[ty] = k;
            bsum += (cnt + plan.cellsPerBlock[ty] - 1) / plan.cellsPerBlock[ty];
            k += cnt;
        }
        blockStart[nTypes] = bsum; cellPrefix[nTypes] = k;
    }
}

// ---- pack ---------------------------------------------------------------------------------------------------
// destination index: 0 = up (rank-1), 1 = down (rank+1), 2 = spawn rank (when it is not a neighbour)
__device__ __forceinline__ int dest_of(const SlabDev& s, int target)
{
    if (target == s.rank - 1) return 0;
    if (target == s.rank + 1) return 1;
    return 2;
}

struct VertexPack {          // the vertex part of the pack: static lists -> the vertex regions of the two neighbour messages
    const int* list[2];
    int count[2];
    VertexRecord* out[2];
    const float4 *vpos, *vvel;
};

__device__ __forceinline__ void pack_vertices(const VertexPack& vp, int first, int stride)
{
    // vein vertices of the static halo lists (owned vertices near a face), both directions
    for (int k = first; k < vp.count[0] + vp.count[1]; k += stride) {
        const int d = k >= vp.count[0] ? 1 : 0, kk = d ? k - vp.count[0] : k;
        const int v = vp.list[d][kk];
        const float4 p = vp.vpos[v], w = vp.vvel[v];
        VertexRecord r;
        r.id = v; r.px = p.x; r.py = p.y; r.pz = p.z; r.vx = w.x; r.vy = w.y; r.vz = w.z; r.pad = 0.f;
        vp.out[d][kk] = r;
    }
}

__global__ void __launch_bounds__(256) slab_pack_vertices_kernel(const VertexPack vp)
{
    pack_vertices(vp, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) slab_pack_kernel(const TypesDev* __restrict__ types, int n, SlabDev slab, const float4* __restrict__ pos,
                                                       const float4* __restrict__ vel, const float4* __restrict__ frc,
                                                       unsigned char* __restrict__ ownedCell, unsigned char* __restrict__ pflag,
                                                       const signed char* __restrict__ moveTo, SlabBuffers buf, int* __restrict__ ghostList,
                                                       int* __restrict__ ghostCount, int* __restrict__ errorFlag, const ActiveItems items,
                                                       const float4* __restrict__ centers, const SlabCount cnt, const VertexPack vp)
{
    // bounded grid striding over the rank's owned items (or, before the lists exist, over all particles)
    const int total = items.cells ? items.cellPrefix[types->n] * items.maxP : n;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        int i = item;
        if (items.cells) {
            // enumerate the owned particles through the owned-cell lists (ghost flags were expired by slab_expire_ghosts)
            int fl = 0;
            i = active_item(items, i, fl);
            if (i < 0 || fl != 1) continue;
        } else {
            const unsigned char f = pflag[i];
            if (!(f & 1)) {
                if (f) pflag[i] = 0;   // last step's ghost expires
                continue;
            }
        }
        int t;
        const int c = cell_of_particle(types, i, &t);
        const int target = moveTo[c];
        const float4 p = pos[i], v = vel[i];
        if (target >= 0) {
            // the blood cell leaves: full state goes to its new owner
            const int d = dest_of(slab, target);
            const int k = atomicAdd(&buf.send[d]->nMig, 1);
            if (k < buf.capMig) {
                MigRecord r;
                r.id = i; r.px = p.x; r.py = p.y; r.pz = p.z; r.vx = v.x; r.vy = v.y; r.vz = v.z;
                const float4 F = frc[i];
                r.fx = F.x; r.fy = F.y; r.fz = F.z;
                const float4 ctr = centers[c];
                r.cx = ctr.x; r.cy = ctr.y; r.cz = ctr.z;
                buf.mig[d][k] = r;
            } else {
                atomicExch(errorFlag, 1);
            }
            // on this side its particles stay around as ghosts for the next step if they are near the face they crossed
            // (a respawned cell goes to the top of the vein, far from any face of this slab, whoever receives it)
            const bool keep = (d == 0 && p.y >= slab.yHi - slab.haloWidth && p.y < slab.yHi + slab.haloWidth) ||
                              (d == 1 && p.y < slab.yLo + slab.haloWidth && p.y >= slab.yLo - slab.haloWidth);
            pflag[i] = keep ? 2 : 0;
            if (keep) {
                ghostList[atomicAdd(ghostCount, 1)] = i;
                if (cnt.enabled) rows_count_particle(cnt.grid, cnt.rows, p, i, 2, cnt.counters);
            }
            if (i == types->t[t].pStart + (c - types->t[t].cStart) * types->t[t].P) ownedCell[c] = 0;
            continue;
        }
        // stays: mirror it on the neighbours whose slab it is close to
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const bool near = d == 0 ? (slab.rank > 0 && p.y >= slab.yHi - slab.haloWidth) : (slab.rank < slab.world - 1 && p.y < slab.yLo + slab.haloWidth);
            if (!near) continue;
            const int k = atomicAdd(&buf.send[d]->nHalo, 1);
            if (k < buf.capHalo) {
                HaloRecord r;
                r.id = i; r.px = p.x; r.py = p.y; r.pz = p.z; r.vx = v.x; r.vy = v.y; r.vz = v.z; r.pad = 0.f;
                buf.halo[d][k] = r;
            } else {
                atomicExch(errorFlag, 1);
            }
        }
    }
    pack_vertices(vp, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// last step's ghosts lose their flag, then (same single CTA, after a barrier) the send headers and the ghost count are
// reset for this step's pack - one launch instead of two on the per-step chain
__global__ void __launch_bounds__(1024) slab_expire_reset_kernel(const int* __restrict__ ghostList, int* __restrict__ ghostCount,
                                                                unsigned char* __restrict__ pflag, SlabBuffers buf, int nv0, int nv1)
{
    const int n = *ghostCount;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int pid = ghostList[k];
        if (!(pflag[pid] & 1)) pflag[pid] = 0;
    }
    __syncthreads();
    if (threadIdx.x < 3) { buf.send[threadIdx.x]->nMig = 0; buf.send[threadIdx.x]->nHalo = 0; buf.send[threadIdx.x]->nVerts = 0; }
    __syncthreads();
    if (threadIdx.x == 0) { *ghostCount = 0; buf.send[0]->nVerts = nv0; buf.send[1]->nVerts = nv1; }
}

__global__ void slab_reset_headers_kernel(SlabBuffers buf, int* ghostCount, int nv0, int nv1)
{
    if (threadIdx.x < 3) { buf.send[threadIdx.x]->nMig = 0; buf.send[threadIdx.x]->nHalo = 0; buf.send[threadIdx.x]->nVerts = 0; }
    if (threadIdx.x == 0) { *ghostCount = 0; buf.send[0]->nVerts = nv0; buf.send[1]->nVerts = nv1; }
}

// ---- unpack -------------------------------------------------------------------------------------------------
constexpr int MAX_UNPACK_SOURCES = 16;
struct UnpackSources {        // every message received this step: blockIdx.y selects one
    const char* raw[MAX_UNPACK_SOURCES];
    unsigned char full[MAX_UNPACK_SOURCES];   // 1: neighbour message (migration + halo + vertices), 0: spawn message (migration only)
    int n;
    size_t haloOffset, vertOffset;            // byte offsets of the halo / vertex regions inside a message
};

__global__ void __launch_bounds__(256) slab_unpack_kernel(const TypesDev* __restrict__ types, const UnpackSources src, int capVert,
                                                         int capMig, int capHaloFull, float4* __restrict__ pos, float4* __restrict__ vel,
                                                         float4* __restrict__ frc, float4* __restrict__ vpos, float4* __restrict__ vvel,
                                                         unsigned char* __restrict__ ownedCell, unsigned char* __restrict__ pflag,
                                                         int* __restrict__ ghostList, int* __restrict__ ghostCount,
                                                         const float4* __restrict__ wallBuilt, float wallMargin, int* __restrict__ wallDirty,
                                                         float4* __restrict__ centers, const SlabCount cnt, int* __restrict__ listCount,
                                                         const int* __restrict__ keepList, const int* __restrict__ keepCount)
{
    // the owned-cell lists are rebuilt by the next kernel: its per-type counters start from zero
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < BCS_MAX_TYPES) listCount[threadIdx.x] = 0;
    if (blockIdx.y == 0) {
        // particles of the blood cells that just left and stay around as ghosts (packed by the cell pass, SlabTail)
        const int nk = *keepCount;
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nk; k += gridDim.x * blockDim.x) {
            const int pid = keepList[k];
            pflag[pid] = 2;
            ghostList[atomicAdd(ghostCount, 1)] = pid;
            if (cnt.enabled) rows_count_particle(cnt.grid, cnt.rows, pos[pid], pid, 2, cnt.counters);
        }
    }
    const char* raw = src.raw[blockIdx.y];
    const bool full = src.full[blockIdx.y] != 0;
    const SlabHeader* hdr = (const SlabHeader*)raw;
    const MigRecord* mig = (const MigRecord*)(raw + sizeof(SlabHeader));
    const HaloRecord* halo = (const HaloRecord*)(raw + src.haloOffset);
    const VertexRecord* verts = (const VertexRecord*)(raw + src.vertOffset);
    const int capHalo = full ? capHaloFull : 0;
    const int nMig = min(hdr->nMig, capMig), nHalo = min(hdr->nHalo, capHalo);
    const int nVerts = full ? min(capVert, hdr->nVerts) : 0;
    const int total = nMig + nHalo + nVerts;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        if (k < nMig) {
            const MigRecord r = mig[k];
            const float4 p = make_float4(r.px, r.py, r.pz, pos[r.id].w);   // .w = collision radius, static per particle
            pos[r.id] = p;
            vel[r.id] = make_float4(r.vx, r.vy, r.vz, 0.f);
            frc[r.id] = make_float4(r.fx, r.fy, r.fz, 0.f);
            pflag[r.id] = 1;
            const int c = cell_of_particle(types, r.id);
            ownedCell[c] = 1;
            centers[c] = make_float4(r.cx, r.cy, r.cz, 0.f);   // every particle of the cell carries the same value
            if (cnt.enabled) rows_count_particle(cnt.grid, cnt.rows, p, r.id, 1, cnt.counters);
        } else if (k < nMig + nHalo) {
            const HaloRecord r = halo[k - nMig];
            const float4 p = make_float4(r.px, r.py, r.pz, pos[r.id].w);
            pos[r.id] = p;
            vel[r.id] = make_float4(r.vx, r.vy, r.vz, 0.f);
            pflag[r.id] = 2;
            ghostList[atomicAdd(ghostCount, 1)] = r.id;
            if (cnt.enabled) rows_count_particle(cnt.grid, cnt.rows, p, r.id, 2, cnt.counters);
        } else {
            const VertexRecord r = verts[k - nMig - nHalo];
            vpos[r.id] = make_float4(r.px, r.py, r.pz, 0.f);
            vvel[r.id] = make_float4(r.vx, r.vy, r.vz, 0.f);
            if (wallBuilt) {
                // wall grid (wall.cu): a halo vertex that left its margin invalidates the padded structure
                const float4 b = wallBuilt[r.id];
                const float dx = r.px - b.x, dy = r.py - b.y, dz = r.pz - b.z;
                if (dx * dx + dy * dy + dz * dz > wallMargin * wallMargin) *wallDirty = 1;
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
void SlabState::exchange(cudaStream_t st)
{
    ncclComm_t c = (ncclComm_t)comm;
    const SlabDev& s = dev;
    BCS_NCCL(ncclGroupStart());
    if (s.rank > 0) {
        BCS_NCCL(ncclSend(sendRaw[0], msgBytes, ncclChar, s.rank - 1, c, st));
        BCS_NCCL(ncclRecv(recvRaw[0], msgBytes, ncclChar, s.rank - 1, c, st));
    }
    if (s.rank < s.world - 1) {
        BCS_NCCL(ncclSend(sendRaw[1], msgBytes, ncclChar, s.rank + 1, c, st));
        BCS_NCCL(ncclRecv(recvRaw[1], msgBytes, ncclChar, s.rank + 1, c, st));
    }
    // respawned cells from ranks that are not neighbours of the spawn rank
    if (s.rank == s.spawnRank) {
        for (int r = 0; r < s.world; ++r)
            if (r != s.rank && r != s.rank - 1 && r != s.rank + 1) BCS_NCCL(ncclRecv(spawnRecvRaw[r], spawnBytes, ncclChar, r, c, st));
    } else if (s.spawnRank != s.rank - 1 && s.spawnRank != s.rank + 1) {
        BCS_NCCL(ncclSend(sendRaw[2], spawnBytes, ncclChar, s.spawnRank, c, st));
    }
    BCS_NCCL(ncclGroupEnd());
}

template <class T>
static T* salloc(SlabState* s, size_t count)
{
    T* p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) throw Error{BCS_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e)};
    BCS_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));   // legacy-stream fill vs the handle's non-blocking stream
    s->owned.push_back((void*)p);
    return p;
}

void slab_unique_id(char out[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    BCS_NCCL(ncclGetUniqueId(&id));
    std::memcpy(out, &id, 128);
}

SlabState* slab_create(const SlabInit& init, const HostScene& hs, const GridDev& tg, const int* dTriIds, const int* dTriCellStart,
                       const int* dTriCellEnd, const SlabCtx& ctx)
{
    BCS_REQUIRE(init.world >= 1 && init.rank >= 0 && init.rank < init.world, BCS_ERR_INVALID, "bad rank / world size");
    BCS_REQUIRE(init.spawnRank >= 0 && init.spawnRank < init.world, BCS_ERR_INVALID, "bad spawn rank");
    BCS_REQUIRE(init.yLo < init.yHi, BCS_ERR_INVALID, "empty slab");
    auto* s = new SlabState();
    try {
        s->dev = SlabDev{1, init.rank, init.world, init.spawnRank, init.yLo, init.yHi, init.haloWidth};
        s->capMig = init.capMig; s->capHalo = init.capHalo;
        s->ownedCell = salloc<unsigned char>(s, ctx.B);
        s->pflag = salloc<unsigned char>(s, ctx.N);
        s->vOwned = salloc<unsigned char>(s, ctx.V);
        s->moveTo = salloc<signed char>(s, ctx.B);
        s->ghostList = salloc<int>(s, ctx.N);
        s->ghostCount = salloc<int>(s, 1);
        s->nActive = salloc<int>(s, 1);
        s->errorFlag = salloc<int>(s, 1);
        s->listCells = salloc<int>(s, ctx.B);
        s->listCount = salloc<int>(s, BCS_MAX_TYPES);
        s->listBlockStart = salloc<int>(s, BCS_MAX_TYPES + 1);
        s->listCellPrefix = salloc<int>(s, BCS_MAX_TYPES + 1);
        s->listDone = salloc<unsigned>(s, 1);
        s->keepList = salloc<int>(s, ctx.N);
        s->keepCount = salloc<int>(s, 1);

        // static vein decomposition from the rest positions: owned vertices, halo lists, triangles to refit
        std::vector<unsigned char> vOwned(ctx.V);
        std::vector<int> vlist[2];
        for (int v = 0; v < ctx.V; ++v) {
            const float y = hs.vy[v];
            vOwned[v] = (y >= init.yLo && y < init.yHi) ? 1 : 0;
            if (!vOwned[v]) continue;
            if (init.rank > 0 && y >= init.yHi - init.vertexHalo) vlist[0].push_back(v);
            if (init.rank < init.world - 1 && y < init.yLo + init.vertexHalo) vlist[1].push_back(v);
        }
        BCS_CUDA(cudaMemcpy(s->vOwned, vOwned.data(), ctx.V, cudaMemcpyHostToDevice));
        {
            // id range of everything this rank integrates or may splat on: rest position inside slab + vertex halo
            int lo = ctx.V, hi = -1;
            for (int v = 0; v < ctx.V; ++v) {
                const float y = hs.vy[v];
                if (y >= init.yLo - init.vertexHalo && y < init.yHi + init.vertexHalo) { lo = std::min(lo, v); hi = std::max(hi, v); }
            }
            s->vFirst = hi >= lo ? lo : 0;
            s->vCount = hi >= lo ? hi - lo + 1 : 0;
        }
        for (int d = 0; d < 2; ++d) {
            s->vertCount[d] = (int)vlist[d].size();
            s->vertList[d] = salloc<int>(s, vlist[d].size());
            if (!vlist[d].empty()) BCS_CUDA(cudaMemcpy(s->vertList[d], vlist[d].data(), vlist[d].size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        // every rank must be able to hold its neighbours' vertex halo: capacity = the largest list over all faces,
        // computed identically everywhere from the rest positions would need the other ranks' bounds; a vertex halo
        // cannot exceed the vertices within vertexHalo of one plane, so bound it by scanning all planes we know of
        s->capVert = std::max(s->vertCount[0], s->vertCount[1]);
        {
            // the neighbours' lists toward us: vertices just outside our faces
            int up = 0, down = 0;
            for (int v = 0; v < ctx.V; ++v) {
                const float y = hs.vy[v];
                if (y >= init.yHi && y < init.yHi + init.vertexHalo) ++up;
                if (y < init.yLo && y >= init.yLo - init.vertexHalo) ++down;
            }
            s->capVert = std::max(s->capVert, std::max(up, down));
        }
        // triangles (sorted slots) whose boxes this rank keeps fresh: everything a local or ghost particle can reach
        {
            std::vector<int> triIds(ctx.T), cs(tg.cells), ce(tg.cells);
            BCS_CUDA(cudaMemcpy(triIds.data(), dTriIds, ctx.T * sizeof(int), cudaMemcpyDeviceToHost));
            BCS_CUDA(cudaMemcpy(cs.data(), dTriCellStart, tg.cells * sizeof(int), cudaMemcpyDeviceToHost));
            BCS_CUDA(cudaMemcpy(ce.data(), dTriCellEnd, tg.cells * sizeof(int), cudaMemcpyDeviceToHost));
            const int nGroups = (ctx.T + 7) / 8;
            std::vector<unsigned char> gl(nGroups, 0), cl(tg.cells, 0);
            const float lo = init.yLo - init.vertexHalo, hi = init.yHi + init.vertexHalo;
            for (int slot = 0; slot < ctx.T; ++slot) {
                const int tri = triIds[slot];
                float ymin = 3e38f, ymax = -3e38f;
                for (int k = 0; k < 3; ++k) {
                    const float y = hs.vy[hs.vidx[3 * tri + k]];
                    ymin = std::min(ymin, y); ymax = std::max(ymax, y);
                }
                if (ymax >= lo && ymin <= hi) gl[slot >> 3] = 1;
            }
            for (int c = 0; c < tg.cells; ++c) {
                if (ce[c] < cs[c]) { cl[c] = 1; continue; }   // empty cells keep an (empty) box: trivial
                for (int g = cs[c] >> 3; g <= (ce[c] >> 3); ++g)
                    if (gl[g]) { cl[c] = 1; break; }
            }
            s->groupLocal = salloc<unsigned char>(s, nGroups);
            s->triCellLocal = salloc<unsigned char>(s, tg.cells);
            BCS_CUDA(cudaMemcpy(s->groupLocal, gl.data(), nGroups, cudaMemcpyHostToDevice));
            BCS_CUDA(cudaMemcpy(s->triCellLocal, cl.data(), tg.cells, cudaMemcpyHostToDevice));
            BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));   // pageable sources: the DMA may outlive the call (capi.cu: dev_upload)
        }

        // BCS_SLAB_NO_COMM=1: profiling aid - one rank of an N-rank decomposition runs alone (no NCCL, no halos)
        if (init.world > 1 && !getenv("BCS_SLAB_NO_COMM")) {
            ncclUniqueId id;
            std::memcpy(&id, init.ncclId, 128);
            ncclComm_t comm;
            BCS_NCCL(ncclCommInitRank(&comm, init.world, id, init.rank));
            s->comm = (void*)comm;
            // Every message of the exchange has the SAME size on every rank (a send and the receive it pairs with must agree,
            // and the vertex halo of a face holds a ring of vertices more or less depending on where the plane falls between
            // two rings): the vertex capacity is the largest any rank needs.
            int* d = salloc<int>(s, 1);
            BCS_CUDA(cudaMemcpy(d, &s->capVert, sizeof(int), cudaMemcpyHostToDevice));
            cudaStream_t q = nullptr;
            BCS_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
            BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
            ncclResult_t r = ncclAllReduce(d, d, 1, ncclInt, ncclMax, comm, q);
            cudaError_t e = cudaStreamSynchronize(q);
            cudaStreamDestroy(q);
            BCS_NCCL(r);
            BCS_CUDA(e);
            BCS_CUDA(cudaMemcpy(&s->capVert, d, sizeof(int), cudaMemcpyDeviceToHost));
        }

        // message buffers
        s->msgBytes = sizeof(SlabHeader) + (size_t)s->capMig * sizeof(MigRecord) + (size_t)s->capHalo * sizeof(HaloRecord) +
                      (size_t)s->capVert * sizeof(VertexRecord);
        s->spawnBytes = sizeof(SlabHeader) + (size_t)s->capMig * sizeof(MigRecord);
        for (int d = 0; d < 3; ++d) s->sendRaw[d] = salloc<char>(s, d < 2 ? s->msgBytes : s->spawnBytes);
        for (int d = 0; d < 2; ++d) s->recvRaw[d] = salloc<char>(s, s->msgBytes);
        s->spawnRecvRaw.assign(init.world, nullptr);
        if (init.rank == init.spawnRank)
            for (int r = 0; r < init.world; ++r)
                if (r != init.rank && r != init.rank - 1 && r != init.rank + 1) s->spawnRecvRaw[r] = salloc<char>(s, s->spawnBytes);
        for (int d = 0; d < 3; ++d) {
            s->buf.send[d] = (SlabHeader*)s->sendRaw[d];
            s->buf.mig[d] = (MigRecord*)(s->sendRaw[d] + sizeof(SlabHeader));
            if (d < 2) s->buf.halo[d] = (HaloRecord*)(s->sendRaw[d] + sizeof(SlabHeader) + (size_t)s->capMig * sizeof(MigRecord));
        }
        s->buf.capMig = s->capMig; s->buf.capHalo = s->capHalo;

        return s;
    } catch (...) {
        slab_destroy(s);
        throw;
    }
}

void slab_destroy(SlabState* s)
{
    if (!s) return;
    if (s->comm) ncclCommDestroy((ncclComm_t)s->comm);
    for (void* p : s->owned) cudaFree(p);
    delete s;
}

static VertexRecord* vertex_region(const SlabState* s, char* raw)
{
    return (VertexRecord*)(raw + sizeof(SlabHeader) + (size_t)s->capMig * sizeof(MigRecord) + (size_t)s->capHalo * sizeof(HaloRecord));
}

static void unpack_all(SlabState* s, const SlabCtx& ctx, const UnpackSources& src, const SlabCount& cnt)
{
    BCS_LAUNCH("slab_unpack", ctx.stream,
               slab_unpack_kernel<<<dim3(32, src.n), 256, 0, ctx.stream>>>(ctx.typesDev, src, s->capVert, s->capMig, s->capHalo, ctx.pos, ctx.vel,
                                                                          ctx.frc, ctx.vpos, ctx.vvel, s->ownedCell, s->pflag, s->ghostList, s->ghostCount,
                                                                          ctx.wallBuilt, ctx.wallMargin, ctx.wallDirty, ctx.centers, cnt, s->listCount,
                                                                          s->keepList, s->keepCount));
}

SlabTail slab_tail(const SlabState* s)
{
    SlabTail t{};
    if (!s->listsValid) return t;   // before the first exchange the pack kernel sweeps the flags itself
    t.ghostList = s->ghostList; t.ghostCount = s->ghostCount; t.pflag = s->pflag; t.ownedCell = s->ownedCell;
    for (int d = 0; d < 3; ++d) { t.sendHdr[d] = reinterpret_cast<int*>(s->buf.send[d]); t.mig[d] = reinterpret_cast<char*>(s->buf.mig[d]); }
    for (int d = 0; d < 2; ++d) t.halo[d] = reinterpret_cast<char*>(s->buf.halo[d]);
    t.capMig = s->capMig; t.capHalo = s->capHalo;
    t.keepList = s->keepList; t.keepCount = s->keepCount; t.errorFlag = s->errorFlag;
    return t;
}

static VertexPack vertex_pack(SlabState* s, const SlabCtx& ctx)
{
    VertexPack vp{};
    for (int d = 0; d < 2; ++d) { vp.list[d] = s->vertList[d]; vp.count[d] = s->vertCount[d]; vp.out[d] = vertex_region(s, s->sendRaw[d]); }
    vp.vpos = ctx.vpos; vp.vvel = ctx.vvel;
    return vp;
}

// fused run: the particle part of the pack rides on the cell pass (SlabTail); the vertex part runs behind the vein integrator
void slab_pack_vertices(SlabState* s, const SlabCtx& ctx, cudaStream_t st)
{
    const int nv = s->vertCount[0] + s->vertCount[1];
    if (!nv) return;
    BCS_LAUNCH("slab_pack_vertices", st, slab_pack_vertices_kernel<<<(nv + 255) / 256, 256, 0, st>>>(vertex_pack(s, ctx)));
    BCS_CUDA(cudaGetLastError());
}

void slab_end_of_step(SlabState* s, const SlabCtx& ctx, bool packed, const SlabCount* count)
{
    cudaStream_t st = ctx.stream;
    SlabCount cnt{};
    if (count) cnt = *count;
    if (!packed) {
        ActiveItems items{};
        if (s->listsValid) {
            items.cells = s->listCells; items.cellPrefix = s->listCellPrefix; items.ghostList = s->ghostList; items.ghostCount = s->ghostCount;
            items.types = ctx.typesDev; items.maxP = ctx.maxP;
            BCS_LAUNCH("slab_expire_reset", st,
                       slab_expire_reset_kernel<<<1, 1024, 0, st>>>(s->ghostList, s->ghostCount, s->pflag, s->buf, s->vertCount[0], s->vertCount[1]));
        } else {
            BCS_LAUNCH("slab_reset", st, slab_reset_headers_kernel<<<1, 32, 0, st>>>(s->buf, s->ghostCount, s->vertCount[0], s->vertCount[1]));
        }
        const long long packItems = std::max<long long>(s->listsValid ? (long long)ctx.B * ctx.maxP : ctx.N, s->vertCount[0] + s->vertCount[1]);
        BCS_LAUNCH("slab_pack", st,
                   slab_pack_kernel<<<(int)std::max<long long>(1, std::min<long long>((packItems + 255) / 256, BOUNDED_BLOCKS)), 256, 0, st>>>(
                       ctx.typesDev, ctx.N, s->dev, ctx.pos, ctx.vel, ctx.frc, s->ownedCell, s->pflag, s->moveTo, s->buf, s->ghostList, s->ghostCount,
                       s->errorFlag, items, ctx.centers, cnt, vertex_pack(s, ctx)));
        BCS_CUDA(cudaGetLastError());
    } else {
        BCS_REQUIRE(s->listsValid, BCS_ERR_STATE, "internal: the cell pass cannot have packed before the first exchange");
    }
    bool unpacked = false;
    if (s->dev.world > 1 && s->comm) {
        s->exchange(st);
        UnpackSources src{};
        src.haloOffset = sizeof(SlabHeader) + (size_t)s->capMig * sizeof(MigRecord);
        src.vertOffset = src.haloOffset + (size_t)s->capHalo * sizeof(HaloRecord);
        auto add = [&](const char* raw, bool full) {
            BCS_REQUIRE(src.n < MAX_UNPACK_SOURCES, BCS_ERR_UNSUPPORTED, "too many ranks for one unpack launch");
            src.raw[src.n] = raw; src.full[src.n] = full ? 1 : 0; ++src.n;
        };
        if (s->dev.rank > 0) add(s->recvRaw[0], true);
        if (s->dev.rank < s->dev.world - 1) add(s->recvRaw[1], true);
        for (char* raw : s->spawnRecvRaw)
            if (raw) add(raw, false);
        if (src.n) {
            unpack_all(s, ctx, src, cnt);
            unpacked = true;
        }
    }
    if (!unpacked) BCS_CUDA(cudaMemsetAsync(s->listCount, 0, BCS_MAX_TYPES * sizeof(int), st));   // (the unpack kernel clears them otherwise)
    slab_build_lists(s, ctx);
    s->listsValid = true;
    BCS_CUDA(cudaGetLastError());
}

// expects the per-type counters (listCount) zeroed; also rewinds the send headers for the next pack
void slab_build_lists(SlabState* s, const SlabCtx& ctx)
{
    cudaStream_t st = ctx.stream;
    BCS_LAUNCH("slab_list_cells", st,
               slab_list_cells_kernel<<<(ctx.B + LIST_THREADS - 1) / LIST_THREADS, LIST_THREADS, 0, st>>>(
                   ctx.typesDev, ctx.B, s->ownedCell, s->listCells, s->listCount, ctx.plan, s->listBlockStart, s->listCellPrefix, s->listDone, s->buf,
                   s->keepCount, s->vertCount[0], s->vertCount[1]));
    BCS_CUDA(cudaGetLastError());
}

OwnedLists slab_lists(const SlabState* s, const TypesDev& types)
{
    OwnedLists l{};
    l.cells = s->listCells; l.count = s->listCount; l.blockStart = s->listBlockStart; l.cellPrefix = s->listCellPrefix;
    for (int t = 0; t < types.n; ++t) l.typeFirst[t] = types.t[t].cStart;
    return l;
}

void slab_prime(SlabState* s, const SlabCtx& ctx)
{
    cudaStream_t st = ctx.stream;
    s->listsValid = false;
    BCS_LAUNCH("slab_init_ownership", st,
               slab_init_ownership_kernel<<<(ctx.B + 127) / 128, 128, 0, st>>>(ctx.typesDev, ctx.B, s->dev, ctx.pos, s->ownedCell, s->pflag, s->moveTo));
    slab_end_of_step(s, ctx);   // nothing migrates (moveTo = -1): plain halo exchange (+ owned-cell lists)
    s->primed = true;
}

int slab_check_error(SlabState* s, cudaStream_t st)
{
    int flag = 0;
    BCS_CUDA(cudaMemcpyAsync(&flag, s->errorFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BCS_CUDA(cudaStreamSynchronize(st));
    return flag;
}

}  // namespace bcs

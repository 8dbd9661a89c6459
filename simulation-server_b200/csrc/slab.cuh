// Slab (multi-GPU) mode: message formats and host-side state.  See slab.cu.
#pragma once
#include "bcs_internal.cuh"
#include "kernels.cuh"

namespace bcs {

struct SlabHeader {          // first 32 bytes of every message
    int nMig;                // particle records of blood cells changing owner
    int nHalo;               // ghost particle records
    int nVerts;              // vein vertex records
    int pad[5];
};
struct MigRecord {           // 52 B: full state of one particle of a migrating blood cell (+ the cell's centre, an output array)
    int id;
    float px, py, pz, vx, vy, vz, fx, fy, fz;
    float cx, cy, cz;
};
struct HaloRecord {          // 32 B
    int id;
    float px, py, pz, vx, vy, vz, pad;
};
using VertexRecord = HaloRecord;

struct SlabBuffers {         // device view of the three send buffers (up, down, spawn rank)
    SlabHeader* send[3];
    MigRecord* mig[3];
    HaloRecord* halo[2];
    int capMig, capHalo;
};

struct SlabCount {           // fused run (capi.cu: run_fused): particles that arrive or stay around as ghosts are counted into
    int enabled;             // the row directory of the NEXT grid build as they are packed / unpacked (the cell pass that
    GridDev grid;            // ended the step has counted the particles that stay owned)
    RowsGrid rows;
    Counters* counters;
};

struct SlabCtx {             // what the slab code needs from the simulation handle
    TypesDev types;
    int N, B, V, T;
    float4 *pos, *vel, *frc, *vpos, *vvel, *centers;
    SpringPlan plan;         // cells per CTA of the cell-group kernels
    int maxP;                // largest particles-per-cell
    const TypesDev* typesDev;   // device copy of the type table
    const float4* wallBuilt;    // wall grid (wall.cu): vertex positions at build time, margin, flag to raise; null = none
    float wallMargin;
    int* wallDirty;
    cudaStream_t stream;
};

struct SlabInit {
    int rank, world, spawnRank;
    float yLo, yHi;
    float haloWidth;         // particle halo
    float vertexHalo;        // vein vertex / triangle halo
    int capMig, capHalo;     // message capacities (particle records)
    const char* ncclId;      // 128-byte ncclUniqueId
};

struct SlabState {
    SlabDev dev{};
    void* comm = nullptr;    // ncclComm_t
    unsigned char *ownedCell = nullptr, *pflag = nullptr, *vOwned = nullptr, *groupLocal = nullptr, *triCellLocal = nullptr;
    signed char* moveTo = nullptr;
    int *ghostList = nullptr, *ghostCount = nullptr, *nActive = nullptr, *errorFlag = nullptr;
    int *listCells = nullptr, *listCount = nullptr, *listBlockStart = nullptr, *listCellPrefix = nullptr;   // owned-cell lists
    unsigned* listDone = nullptr;   // CTA arrival counter of the list kernel (the last one writes the prefixes)
    int *keepList = nullptr, *keepCount = nullptr;   // fused run: particles of cells that just left and stay around as ghosts
    char* sendRaw[3] = {nullptr, nullptr, nullptr};
    char* recvRaw[2] = {nullptr, nullptr};
    std::vector<char*> spawnRecvRaw;
    size_t msgBytes = 0, spawnBytes = 0;
    SlabBuffers buf{};
    int capMig = 0, capHalo = 0, capVert = 0;
    int* vertList[2] = {nullptr, nullptr};
    int vertCount[2] = {0, 0};
    int vFirst = 0, vCount = 0;   // id range of the vertices within slab + vertex halo (what the vertex kernels sweep)
    bool primed = false;
    bool listsValid = false;   // owned-cell lists describe the current ownership
    std::vector<void*> owned;

    void exchange(cudaStream_t st);
};

// host-side entry points (slab.cu)
SlabState* slab_create(const SlabInit& init, const HostScene& hs, const GridDev& tgrid, const int* dTriIds, const int* dTriCellStart,
                       const int* dTriCellEnd, const SlabCtx& ctx);
void slab_destroy(SlabState* s);
void slab_prime(SlabState* s, const SlabCtx& ctx);          // ownership from the uploaded state + first halo exchange
// pack -> exchange -> unpack -> owned-cell lists.  packed: the cell pass that ended the step has expired the old ghosts and
// packed the particle records (SlabTail, kernels.cuh) and slab_pack_vertices has run; count: see SlabCount
void slab_end_of_step(SlabState* s, const SlabCtx& ctx, bool packed = false, const SlabCount* count = nullptr);
SlabTail slab_tail(const SlabState* s);
void slab_pack_vertices(SlabState* s, const SlabCtx& ctx, cudaStream_t st);
OwnedLists slab_lists(const SlabState* s, const TypesDev& types);
void slab_build_lists(SlabState* s, const SlabCtx& ctx);       // compacted owned blood cells for the cell-group kernels
void slab_unique_id(char out[128]);
int slab_check_error(SlabState* s, cudaStream_t st);        // 1 if a message overflowed since creation

}  // namespace bcs

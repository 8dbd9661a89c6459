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
struct MigRecord {           // 40 B: full state of one particle of a migrating blood cell
    int id;
    float px, py, pz, vx, vy, vz, fx, fy, fz;
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

struct SlabCtx {             // what the slab code needs from the simulation handle
    TypesDev types;
    int N, B, V, T;
    float4 *pos, *vel, *frc, *vpos, *vvel;
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
void slab_end_of_step(SlabState* s, const SlabCtx& ctx);    // pack -> exchange -> unpack
OwnedLists slab_lists(const SlabState* s, const TypesDev& types);
void slab_build_lists(SlabState* s, const SlabCtx& ctx);       // compacted owned blood cells for the cell-group kernels
void slab_unique_id(char out[128]);
int slab_check_error(SlabState* s, cudaStream_t st);        // 1 if a message overflowed since creation

}  // namespace bcs

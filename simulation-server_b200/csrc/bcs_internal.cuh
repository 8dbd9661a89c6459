// Internal declarations of libbcs (sm_100a).  Nothing here crosses the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "bcs.h"

namespace bcs {

// ---------------------------------------------------------------------------------------------
// error plumbing: C ABI functions return codes, never exit (the reference printf+exit()s,
// utilities/cuda_handle_error.cuh:18-25)
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
struct Error {
    int code;
    std::string msg;
};
#define BCS_CUDA(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            throw ::bcs::Error{BCS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                                 ":" + std::to_string(__LINE__) + ")"};                        \
    } while (0)
#define BCS_REQUIRE(cond, code, msg)                       \
    do {                                                   \
        if (!(cond)) throw ::bcs::Error{(code), (msg)};    \
    } while (0)

// ---------------------------------------------------------------------------------------------
// launch bookkeeping: every kernel launch goes through BCS_LAUNCH so that the library can (a) report how
// many of its kernels ran (bcs_get_launch_count) and (b) time each kernel with CUDA events on the launching
// stream when profiling is on (bcs_profile_steps).
// ---------------------------------------------------------------------------------------------
struct LaunchRecord {
    const char* name;
    cudaEvent_t start, stop;
};
struct LaunchCtx {
    unsigned long long launches = 0;
    bool timing = false;
    std::vector<LaunchRecord> records;
};
LaunchCtx* current_launch_ctx();
void set_launch_ctx(LaunchCtx* c);
void launch_begin(const char* name, cudaStream_t st);
void launch_end(cudaStream_t st);
#define BCS_LAUNCH(name, st, ...)            \
    do {                                     \
        ::bcs::launch_begin(name, st);       \
        __VA_ARGS__;                         \
        ::bcs::launch_end(st);               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device-side parameter blocks (passed by value to kernels / kept in __constant__-like structs)
// ---------------------------------------------------------------------------------------------
struct TypeDev {            // one blood-cell type in final (meta-factory) order
    int count;
    int P;                  // particles per cell
    int pStart;             // particleStarts
    int cStart;             // bloodCellTypesStarts
    int mStart;             // bloodCellModelStarts
    int warpSync;           // reference would use handleVeinEndsWarpSync (no ending-sphere test)
    int adjStart;           // offset of this type's adjacency (ELL) in adjJ/adjL (and of the incidence list adjS)
    int maxDeg;             // ELL width
    int sprStart;           // offset of this type's undirected spring list in sprAB/sprL
    int nSpr;               // undirected springs per cell
};

struct TypesDev {
    int n;
    TypeDev t[BCS_MAX_TYPES];
};

struct GridDev {
    float minx, miny, minz;     // grid origin (minX,minY,minZ)
    float lenx, leny, lenz;     // width,height,depth (the clamp bound of calculateIdForCell)
    float maxx, maxy, maxz;
    int csx, csy, csz;          // cell size (ints, converted to float at use like the reference)
    int nx, ny, nz;             // cell counts
    int cells;
    int n;                      // objects
    int keyBits;                // bits needed for a cell id
    float icsx, icsy, icsz;     // 1 / cell size, exact when pow2 != 0 (the reference's 2-unit particle grid)
    int pow2;                   // all three cell sizes are powers of two: (p - min) / cs == (p - min) * ics bit for bit
};

struct PhysDev {
    float dt;
    float velocity_collision_damping;
    float particle_k_sniff, vein_k_sniff, particle_d_fact, vein_d_fact;
    float vein_collision_force_intensity;
    float viscous_damping;
    float coll_spring, coll_damping, coll_shear;
    float max_cell_size_factor, big_brake_intensity;
    float initvx, initvy, initvz;
    float impact2;              // veinImpactDistance^2 (float product)
    float impactNear;           // veinImpactDistance + margin, used only to cull
    float minForce2;            // veinImpactMinimalForceDistance^2
    float gx, gy, gz;
    float min_spawn_y, cylinder_radius;
    // vein-end thresholds, vein_end.cu:12-18
    float upperY, lowerY, rightX, leftX, frontZ, backZ;
    int useBloodFlow, reactionForce, bigBrake;
    int nEndings;
};

// slab decomposition along y (multi-GPU): this rank owns the blood cells whose centre lies in [yLo, yHi)
struct SlabDev {
    int enabled;
    int rank, world, spawnRank;   // spawnRank: the rank whose slab contains minSpawnY (respawned cells go there)
    float yLo, yHi;               // the top slab has yHi = +inf, the bottom slab yLo = -inf
    float haloWidth;              // particles within this distance of a face are mirrored on the neighbour
};

// slab mode: the blood cells this rank owns, compacted per type every step (device resident).  Type t's cells sit in
// cells[typeFirst[t] .. typeFirst[t] + count[t]) in no particular order; blockStart is the exclusive prefix of
// ceil(count[t] / cellsPerBlock[t]) used by the kernels that give a CTA a group of whole blood cells.
struct OwnedLists {
    const int* cells;          // null = no slab decomposition (all cells, identity order)
    const int* count;          // [n_types]
    const int* blockStart;     // [n_types + 1]
    const int* cellPrefix;     // [n_types + 1] exclusive prefix of count
    int typeFirst[BCS_MAX_TYPES];
};

// slab mode: enumeration of this rank's ACTIVE particles (owned blood cells x particles, then the ghosts) so that the
// per-particle kernels of the grid build and of the exchange touch N_local instead of N entries
// Slab mode: kernels over device-side counts are launched with at most this many CTAs (8 per SM of a B200) and stride
// over the count, so a rank's step costs O(local particles) however large the global scene is.
constexpr int BOUNDED_BLOCKS = 148 * 8;

struct ActiveItems {           // small on purpose: passed by value to streaming kernels
    const int* cells;          // null: not in slab mode (kernels index all particles)
    const int* cellPrefix;     // [n_types + 1] exclusive prefix of the owned-cell counts
    const int* ghostList;
    const int* ghostCount;
    const TypesDev* types;     // device copy of the type table
    int maxP;
};

struct Counters {               // device-resident, see bcs_stats
    unsigned long long pairTests, pairHits, triTests, veinHits, teleported, oob;
    unsigned long long step;    // completed steps (drives the respawn RNG counter)
};

// packed triangle in sorted-slot order, refreshed every step from the moving vein vertices
struct TriPacked {
    float4 a;   // v0.xyz, e1.x
    float4 b;   // e1.yz, e2.xy
    float4 c;   // e2.z, triangle id (int bits), unused, unused
};

struct Aabb {
    float lox, loy, loz, hix, hiy, hiz;
};

// slab { x : dmin <= n.x <= dmax } along a unit normal n
struct CellSlab {
    float nx, ny, nz, dmin, dmax;
};

// work item of the vein-collision stage: a blood cell that may be within reach of the wall, with the
// triangle-grid cells (a <= 4x4x4 block starting at cx0,cy0,cz0; bit = (dz*4 + dy)*4 + dx) that may matter
struct CullEntry {
    int cell;
    int cx0, cy0, cz0;
    unsigned long long mask;
};

// ---------------------------------------------------------------------------------------------
// host-side derived scene (meta_factory + generateBoundingSpheres equivalents), scene_host.cpp
// ---------------------------------------------------------------------------------------------
struct HostType {
    int count, P, pStart, cStart, mStart, gStart, srcDef, warpSync;
    float smallestRadius;
};

struct HostScene {
    std::vector<HostType> types;
    int N = 0, B = 0, nModel = 0, nGraph = 0, V = 0, T = 0;
    std::vector<float> graph;               // dense spring lengths
    std::vector<float> mx, my, mz;          // model vertices in final order
    std::vector<float> collR, initR;        // collision radius, distance from model centroid
    std::vector<int32_t> adjJ;              // ELL adjacency per type: [adjStart + d*P + i] = mate index j (or -1)
    std::vector<float> adjL;                // spring length
    std::vector<int> adjStart, maxDeg;
    // undirected spring list per type (a < b) and, parallel to adjJ, the spring each adjacency entry refers to
    // (bit 31 set: the particle is the b end, i.e. the stored force enters with a minus sign)
    std::vector<int32_t> sprAB;             // a | b << 16
    std::vector<float> sprL;
    std::vector<int32_t> adjS;
    std::vector<int> sprStart, nSpr;
    float gmin[3], gmax[3], gsize[3];
    std::vector<float> vx, vy, vz;
    std::vector<uint32_t> vidx;
    std::vector<int32_t> nbrIds;            // [slot][vertex]
    std::vector<float> nbrLen;
    std::vector<float> tcx, tcy, tcz;       // triangle centres (initial)
    std::vector<float> endC, endR;
    int cellSize[3], triCellSize[3];
    int gdims[3], tdims[3];
    bcs_physics ph;
    int useBloodFlow, reactionForce, bigBrake, bsCoeff;
};

void derive_scene(const bcs_scene& in, HostScene& out);   // throws bcs::Error

}  // namespace bcs

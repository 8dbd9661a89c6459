// Launchers of the hot-path kernels (one .cu per stage family).  All launches are asynchronous on
// the stream given; errors are thrown as bcs::Error.
#pragma once
#include "bcs_internal.cuh"

namespace bcs {

// ---- grid.cu -------------------------------------------------------------------------------------------
struct SortScratch {
    unsigned* tileHist = nullptr;      // [numTiles][256] per-tile digit counts -> global offsets
    unsigned* digitTotals = nullptr;   // [4][256] whole-array digit counts per pass
    int numTiles = 0;
    // occupancy-rank counting sort (compact cell index)
    bool radixForCompact = false;      // BCS_GRID=radix: build the compact index from a radix sort instead
    int *keyOf = nullptr, *rankOf = nullptr, *placeOf = nullptr, *tmpIds = nullptr, *tmpRank = nullptr, *scanTotals = nullptr;
    unsigned* cellCount = nullptr;
    unsigned long long* scanStatus = nullptr;   // single-launch scans: (epoch << 32 | tile total) per tile
    unsigned* scanCtl = nullptr;                //                      ticket, done, epoch
    bool twoPassScan = true;                    // tile totals + scan as two launches; BCS_SCAN=fused: one ticketed launch per scan
    int* finTileCount = nullptr;       // occupied-cell starts per finalize tile (compact index build)
    unsigned* status = nullptr;        // [4][numTiles][256] look-back status words of the onesweep passes
    bool classic = false;              // BCS_SORT=classic: 3-kernel passes (histogram, scan, scatter) instead of onesweep
    void allocate(int n, int maskWordsHint = 0);
    void release();
};

// Row-directory grid build (clean semantics, sparse scenes; BCS_GRID=rows).  A "row" is the run of nx x-adjacent cells
// of one (y, z): cell ids are x-fastest (uniform_grid.cu:33-35), so row r = id / nx and the sorted list is row-major.
//   count    per particle: key; rowCount[row] += 1 (a reduction: nothing waits for it; rides in the cell pass of the previous step)
//   scan     rowStart[row + 1] = exclusive scan of rowCount = first slot of the row   (0.7 M rows for the 1 M-particle vein)
//   scatter  tmp[atomicAdd(rowStart[row + 1], 1)] = (key, id)               rows contiguous, unordered inside; the bumps turn
//                                                                           rowStart[row + 1] into the row's END = start of row + 1,
//                                                                           i.e. rowStart[] is now the start table itself
//   order    per slot: rank of (key, id) inside its row (rows hold ~5 particles) -> keys / ids / sorted positions,
//            rowCount cleared for the next build
// The (cell id, particle id) order is unique, so keys / ids are bit-identical to the other grid builds.
struct RowsGrid {
    int enabled;
    int nRows;
    unsigned* rowCount;       // [nRows + 1] all zero between builds
    int* rowStart;            // [nRows + 2] after the scatter pass: first sorted slot of each row, rowStart[nRows] = number of
                              // sorted slots; rowStart[0] is 0 for ever (the scan writes from [1] on)
    int* keyOf;               // [N] by particle id: cell id
    int2* tmp;                // [N] row-contiguous (cell id, id | ghost tag)
    int* irregular;           // [2] device flags: a particle sits outside the grid (its stencil is not symmetric).  [0] raised by
                              // the count pass, latched into [1] (and cleared) by the order pass; the collision stage reads [1]
    unsigned long long nxMagic;   // exact division by nx: q = (n * nxMagic) >> nxShift
    int nxShift;
    int countDone;            // 1: the count pass of THIS build already ran (fused into the previous step's cell pass)
    int nx;                   // cells per row
    // slab decomposition: the directory only spans the cell rows y0 .. y0 + nyL - 1 of every z layer (the rank's slab plus its
    // halo), in the same z-major order, so that its scan costs O(local rows) however long the vein is.  local == 0: all rows.
    int local, ny, nyL, y0;
    unsigned long long nyMagic;   // exact division by ny
    int nyShift;
    int* error;               // sticky device flag raised when an active particle lies outside the local window
};

struct NearProbe;

struct GridBuildArgs {
    GridDev grid;
    const float4* objPos;     // positions the keys are computed from (particles or triangle centres)
    int* keys[2];             // ping-pong; the sorted result always lands in [1]
    int* ids[2];
    int* cellStart;
    int* cellEnd;
    SortScratch* scratch;
    Counters* counters;
    bool reference;           // BCS_SEM_REFERENCE table semantics
    bool tablesValid;         // clean semantics: tables hold last step's ranges (to be un-written)
    bool compact;             // clean semantics, particle grid: build the compact cell index instead of dense tables
    unsigned* cellMask;       // [maskWords] occupancy bits
    int maskWords;
    int* cellRank;            // [maskWords]
    int* occStart;            // [n + 1]
    int* occKey;              // [n]
    int* numOcc;              // device scalar
    const unsigned char* pflag;  // slab mode: per-particle flags (bit 0 owned, bit 1 ghost); null otherwise
    ActiveItems items;        // slab mode: enumeration of the active particles (items.lists.cells == null otherwise)
    long long itemCapacity;   //            upper bound of the item count (launch size)
    const int* nDev;          // slab mode: number of active particles (device scalar read by the sort / finalize kernels)
    int* nDevOut;             //            ... written by the key kernel
    bool reorder;             // also write sorted-order copies of pos/vel
    const float4* pos;
    const float4* vel;
    float4* spos;
    float4* svel;
    RowsGrid rows;            // rows.enabled: row-directory build (no compact cell index, no svel)
    const NearProbe* probe;   // rows mode: near-wall probe carried by the order pass (null: none)
};
void launch_grid_build(const GridBuildArgs& a, cudaStream_t st);
void launch_row_count(const GridBuildArgs& a, cudaStream_t st);   // the count pass alone (head of a bcs_step run)

// ---- cellpass.cu: the per-blood-cell pass (springs, integration, vein end, row count) ------------------------------
// Work unit: a GROUP of whole blood cells of one type handled by ONE WARP out of warp-private shared memory (no CTA
// barrier anywhere).  The plan fixes the group size per type and the offsets of the per-type tables.
struct SpringPlan {
    int blockStart[BCS_MAX_TYPES + 1];   // first group of each type
    int cellsPerBlock[BCS_MAX_TYPES];    // blood cells per group: 1, 2, 4 or 8
    int totalBlocks;                     // groups over all types
    int sharedBytes;                     // dynamic shared memory per CTA
    int warpBytes;                       // ... of which per warp
    int warps;                           // warps per CTA
    int tileMax;                         // tile capacity per warp: P * (cellsPerBlock + 1) float4, maximum over the types
    int slotOff[BCS_MAX_TYPES];          // slot table: [P] particle index | degree << 16, sorted by degree
    int adjOff[BCS_MAX_TYPES];           // adjacency: [adjDeg][P] (mate particle index, rest length bits) per slot, ascending mates
    int adjDeg[BCS_MAX_TYPES];
    unsigned cellMagic[BCS_MAX_TYPES];   // i / P == (i * cellMagic) >> 20 for i < 512
    int tabAdj, tabSlot, tabModel, tabEnd;   // sizes (entries) of the tables staged in shared memory per CTA
    int tabBytes;                        // ... in bytes, 0 = tables are read from global memory
};
struct SpringTables {                    // host copies of the per-type tables (uploaded once per handle)
    std::vector<int> slot;
    std::vector<int2> adj;
};
SpringPlan make_spring_plan(const TypesDev& types, const HostScene& hs, SpringTables& tables, int world = 1);
// Near-wall probe carried by the spring kernel (every particle's position passes through its registers anyway): one byte
// of the wall grid's dilated occupancy answers "can this particle reach the wall at all"; the few that can are listed
// for the wall filter, which then never streams the other ~97 % of the particles.  near == null: disabled.
struct NearProbe {
    const unsigned char* near;  // [wall-grid cells]
    int* list;                  // [N] particle ids
    int* count;
    float ox, oy, oz, invh;
    int nx, ny, nz;
};
struct SpringArgs {
    TypesDev types;
    const TypesDev* typesDev;   // device copy of `types`
    SpringPlan plan;
    PhysDev phys;
    const float4* pos;
    const float4* vel;
    float4* frc;
    float4* centers;            // [B]
    const int* slotTab;         // per-type slot tables (SpringPlan::slotOff)
    const int2* adjTab;         // per-type adjacency in slot order (SpringPlan::adjOff)
    const float* initR;         // [nModel]
    OwnedLists lists;           // slab mode: owned blood cells (lists.cells == null otherwise)
    NearProbe probe;            // probe.near == null: no probe
    unsigned long long* phaseClock;   // developer aid (BCS_CP_CLOCK): summed SM cycles per phase of the cell pass; null in production
};
void launch_springs(const SpringArgs& a, cudaStream_t st);
void cell_pass_prepare(const SpringPlan& plan);   // per device: opt in to the dynamic shared memory of every variant

// ---- collide.cu / pairs.cu ------------------------------------------------------------------------------
// Touching pairs found by the symmetric search (pairs.cu): a global list of (slot i < slot j), filled with one atomic per
// CTA, and per-particle fixed-point force accumulators (pair_device.cuh) that make the sums independent of arrival order.
struct PairLists {
    int2* pairs;                // [cap] touching pairs of this step (sorted slots)
    int cap;
    int* ctl;                   // [4] pairs listed, overflow flag (pair_walk redoes the stage), CTA counter of the fold pass
    long long* acc;             // [3 N] by particle id, unit 2^-40; zero between steps
};

struct CollideArgs {
    GridDev grid;
    TypesDev types;
    PhysDev phys;
    int n;
    const int* nDev;            // slab mode: number of active sorted slots (device scalar); null otherwise
    const int* keys;            // sorted cell ids
    const float4* spos;         // sorted positions  (w: radius | particle id bits in reference mode)
    const float4* svel;         // sorted velocities (w: particle id bits)
    const int* cellStart;       // dense tables (reference-compatible semantics)
    const int* cellEnd;
    const unsigned* cellMask;   // compact cell index (clean semantics), see grid.cu
    const int* cellRank;
    const int* occStart;
    bool rows;                  // BCS_COLLIDE=rows: row-after-row candidate walk instead of the flattened one
    bool tiled;                 // shared-memory tiled kernel (needs the counting-sort grid: cellRank is a full prefix there)
    unsigned long long nxMagic; // exact division by grid.nx: q = (n * nxMagic) >> nxShift
    int nxShift;
    const float* collR;         // [nModel] (reference-mode lookup)
    float4* frc;
    Counters* counters;
    bool reference;
    bool stats;
    // debug outputs (indexed by particle id), all null in production
    int* dbgCount;
    unsigned long long* dbgSum;
    int* dbgHits;
    // row-directory mode (pairs.cu): symmetric pair search over the sorted keys + per-slot hit lists
    bool rowsMode;
    bool fullWalk;              // BCS_COLLIDE=walk: every slot scans its whole stencil (the fallback path) instead
    bool deferFold;             // the sums stay parked in PairLists::acc: the wall apply and the cell pass fold them (bcs_step)
    const int* rowStart;
    RowsGrid rowsGrid;          // the directory's geometry (row of a cell id: rows_of_key, rows_device.cuh)
    int nRows;
    const int* ids;             // sorted particle ids (bit 31: ghost)
    const float4* vel;          // velocities by particle id
    const int* irregular;       // raised by the grid build when a particle sits outside the grid
    PairLists pairs;
};
void launch_particle_collisions(const CollideArgs& a, cudaStream_t st);
void launch_particle_collisions_rows(const CollideArgs& a, cudaStream_t st);   // pairs.cu

// ---- wall.cu: lazily rebuilt wall grid (production path of the vein-collision stage, clean semantics) ----
// A uniform grid of `h`-unit cells over the vein's bounding box.  Every cell lists the sorted triangle slots whose
// AABB (padded by BOX_PAD + margin) overlaps it, preceded by a 5-word header: the slab {n, dmin, dmax} along the
// mean normal of those triangles (padded likewise).  The structure stays valid while no vertex has moved more than
// `margin` from its position at build time (vein_integrate / the vertex-halo unpack raise `dirty` otherwise), so a
// step normally rebuilds nothing: the rebuild kernel returns at once.
struct WallGridDev {
    int enabled;
    float ox, oy, oz, h, invh;
    int nx, ny, nz, cells;
    int* start;               // [cells + 1] offsets into list
    int4* rec;                // [2 * cells] 32-byte record per cell: {list start, entries, n.x, n.y | n.z, dmin, dmax, -}
    unsigned char* occ;       // [cells] 1 = the cell lists at least one triangle
    unsigned char* occ3;      // [cells] bits 0..2 = occ of the cell and its two +x neighbours
    unsigned char* near;      // [cells] 1 = a non-empty cell lies within +-2 cells (a particle here can reach the wall)
    unsigned char* nearTmp;   // [2 * cells] scratch of the separable dilation
    int* cursor;              // [cells]     counts during a rebuild, then fill cursors
    int* list;                // [cap]
    int cap;
    float4* vposBuilt;        // [V] vertex positions the structure was built from
    const int4* slotInfo;     // [T] static: triangle id and triangle-grid cell (x,y,z) of every sorted slot
    const int4* slotVerts;    // [T] static: the three vertex ids (and the triangle id) of every sorted slot
    Aabb* groupBox;           // [(T+7)/8] AABB of 8 consecutive sorted slots, padded by BOX_PAD + margin
    Aabb* cellBox;            // [triangle-grid cells] union of the group boxes a cell's slot range touches
    int* dirty;               // device flag: a vertex left its margin -> rebuild at the start of the next step
    int* overflow;            // sticky: list capacity exceeded
    unsigned* barrier;        // [2] grid barrier of the rebuild kernel
    int* blockSums;           // [rebuild grid]
    unsigned long long* builds;   // number of rebuilds so far
    float margin;
    int* queue;               // [N] particles with a near hit (phase B work list)
    unsigned char* ghostFlag; // [N] slab mode: 1 = ghost (splat only); valid for candidates
    int* queueCount;          // [3] {phase-B particles, entries, near-wall particles}: adjacent, cleared together
    int* entryCount;          // = queueCount + 1
    int* nearCount;           // = queueCount + 2
    int* nearList;            // [N] particles the spring kernel's probe found near the wall (this step)
    int useNearList;          // 1: the filter reads nearList (+ the ghost list in slab mode) instead of every particle
    int2* entries;            // [entryCap] (particle, wall-grid cell) pairs whose triangles are to be tested
    int entryCap;
    unsigned long long* best; // [N] (traversal key << 32 | slot) of the first near hit; valid for candidates only
};

// ---- vein.cu -------------------------------------------------------------------------------------------
struct VeinArgs {
    int V, T;
    PhysDev phys;
    float4* vpos;
    float4* vvel;
    float4* vfrc;               // .w != 0: wall splats are parked in vsplat (see vein_device.cuh: splat_add)
    long long* vsplat;          // [3V] fixed-point (2^-40) sums of this step's wall-collision splats
    const int* nbrIds;          // [9][V]
    const float* nbrLen;
    const unsigned* vidx;       // [3T]
    const unsigned char* vOwned;   // slab mode: [V] 1 = vertex integrated by this rank; null = all
    int vFirst, vCount;            // the vertex kernels cover ids [vFirst, vFirst + vCount): everything, or (slab mode) the id range
                                   // of the rank's slab + vertex halo - meshes are numbered ring by ring, so that is O(local)
    const float4* vposBuilt;       // wall grid: positions at build time, margin and the flag to raise (null: no tracking)
    float wallMargin;
    int* wallDirty;
};
void launch_tri_centers(const VeinArgs& a, float4* centers, cudaStream_t st);
void launch_vein_gather(const VeinArgs& a, cudaStream_t st);
void launch_vein_integrate(const VeinArgs& a, cudaStream_t st);
void launch_vein_fold_splats(const VeinArgs& a, cudaStream_t st);   // vfrc += parked splats (before vfrc is read back)

struct VeinCollideArgs {
    GridDev tgrid;
    TypesDev types;
    PhysDev phys;
    int n;                      // particles
    int T;
    float4* pos;
    float4* vel;
    float4* frc;
    const float4* vpos;
    float4* vfrc;
    long long* vsplat;          // [3V] fixed-point splat accumulators (order-independent sums)
    const unsigned* vidx;
    const int* triIds;          // sorted triangle ids
    const int* cellStart;
    const int* cellEnd;
    TriPacked* tris;            // [T] packed triangles in sorted-slot order (refit each step)
    Aabb* groupBox;             // [(T+7)/8] padded AABB of each group of 8 sorted slots (refit each step)
    CellSlab* groupSlab;        // [(T+7)/8] padded slab of the same groups along their mean normal
    Aabb* cellBox;              // [cells]   padded AABB of everything a cell's table range reaches
    CellSlab* cellSlab;         // [cells]   padded slab along the mean triangle normal of the same range
    const unsigned char* groupLocal;    // slab mode: [(T+7)/8] slot groups refitted by this rank; null = all
    const unsigned char* triCellLocal;  // slab mode: [cells] triangle-grid cells refitted by this rank; null = all
    OwnedLists lists;                   // slab mode: owned blood cells (lists.cells == null otherwise)
    const int* ghostList;               // slab mode: ghost particle ids (splat-only pass) and their count
    const int* ghostCount;
    ActiveItems items;                  // slab mode: enumeration of the rank's owned particles + ghosts (wall filter)
    bool fast;                  // culled two-phase search (default) vs exhaustive reference-order traversal
    int liveTris;               // 1: triangles are gathered from the live vertices (tris is not refreshed per step)
    WallGridDev wall;           // wall.enabled: production path (clean semantics)
    const unsigned char* pflag; // slab mode: per-particle flags (bit 0 owned, bit 1 ghost); null otherwise
    long long* pairAcc;         // deferred fold (pairs.cu): fixed-point pair forces still parked per particle; null = already folded
    int nCells;                 // blood cells
    int maxP;                   // largest particles-per-cell over the types
    CullEntry* cullList;        // [nCells] blood cells that may touch the wall this step
    int* cullCount;             // device scalar
    const float* collR;
    Counters* counters;
    bool stats;
    bool apply;                 // false: only fill the debug outputs
    int* dbgTri;
    float* dbgT;
};
void launch_tri_refit(const VeinCollideArgs& a, cudaStream_t st);
void launch_vein_collisions(const VeinCollideArgs& a, cudaStream_t st);
// wall.cu
void launch_wall_rebuild(const VeinCollideArgs& a, int V, int numSMs, cudaStream_t st);   // returns at once unless wall.dirty
void launch_wall_collisions(const VeinCollideArgs& a, cudaStream_t st);   // = search + apply
void launch_wall_reset(const VeinCollideArgs& a, cudaStream_t st);   // clears the search counters (before the spring kernel's near-wall probe)
void launch_wall_search(const VeinCollideArgs& a, cudaStream_t st);
void launch_wall_apply(const VeinCollideArgs& a, cudaStream_t st);
void launch_wall_slot_info(const int* sortedTriKeys, const int* triIds, const unsigned* vidx, int T, GridDev tgrid, int4* slotInfo, int4* slotVerts,
                           cudaStream_t st);

// ---- integrate.cu --------------------------------------------------------------------------------------
// Slab mode, fused run: the end-of-step chores of the halo exchange ride on the cell pass that ends the step (cellpass.cu)
// instead of launches of their own:
//   * on its way in, every CTA walks a share of last step's ghost list and clears the flags (the collision and wall stages
//     are through with them); the last CTA to leave rewinds the ghost count;
//   * the write-back sweep appends a halo record for every particle near a slab face and, for a blood cell that changes
//     owner, its migration records (finished force, centre) to the send buffers (formats: slab.cuh) - the state is in
//     registers there anyway, the stand-alone pack kernel re-read 32 B per owned particle.
// Particles of a leaving cell that stay around as ghosts go to keepList; the unpack kernel flags and counts them.
struct SlabTail {
    const int* ghostList;   // null = not used
    int* ghostCount;
    unsigned char* pflag;
    unsigned char* ownedCell;
    int* sendHdr[3];        // SlabHeader of the up / down / spawn-rank message: {nMig, nHalo, nVerts, ...}
    char* mig[3];           // MigRecord regions of the three messages
    char* halo[2];          // HaloRecord regions of the two neighbour messages
    int capMig, capHalo;
    int* keepList;
    int* keepCount;
    int* errorFlag;         // sticky: a message overflowed
};

struct IntegrateArgs {
    TypesDev types;
    PhysDev phys;
    int n;
    int nCells;
    float4* pos;
    float4* vel;
    const float4* frc;
    const float* mx;
    const float* my;
    const float* mz;
    const float* endC;
    const float* endR;
    Counters* counters;
    unsigned long long seed;
    // slab mode
    SlabDev slab;
    OwnedLists lists;                 // owned blood cells
    signed char* moveTo;              // [B] out: rank the blood cell migrates to after this step, -1 = stays
    long long* pairAcc;               // deferred fold (pairs.cu): parked pair forces are added to frc before it is used; null = none
    SlabTail tail;                    // slab mode, fused run
};
void launch_integrate_particles(const IntegrateArgs& a, cudaStream_t st);
void launch_vein_end(const IntegrateArgs& a, cudaStream_t st);   // also advances the device step counter
// integrate_particles + vein_end + step counter fused (bcs_step): the cell pass without its spring stage
void launch_finish_step(const IntegrateArgs& a, const SpringArgs& sp, unsigned* doneBlocks, cudaStream_t st);
// end of step k fused with the head of step k + 1: integrate + vein end + step counter, then - on the new state - blood
// cell centres + springs, and the row count of the next grid build
void launch_advance(const IntegrateArgs& a, const SpringArgs& sp, const GridDev& grid, const RowsGrid& rows, unsigned* doneBlocks, cudaStream_t st);
// springs + row count (head of a bcs_step run)
void launch_springs_count(const SpringArgs& sp, const GridDev& grid, const RowsGrid& rows, Counters* counters, cudaStream_t st);

}  // namespace bcs

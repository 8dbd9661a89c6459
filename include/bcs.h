/* bcs.h - C ABI of the B200-native blood-cell simulation step ("libbcs").
 *
 * The reference (GPU-Blood-Cell-Simulation/Simulation-Server) has no FFI seam; its seam is the set of
 * C++ calls programLoop() makes on its state objects (src/main.cu:122-138,175-176,199,208).  Every entry
 * point below names the reference interface it stands in for.  Plain pointers and sizes only; no C++,
 * CUDA or torch types cross this boundary.  INTEGRATION.md shows the binding a reference maintainer
 * would add (include/bcs_reference_shim.hpp fills bcs_scene from the reference's own config headers).
 *
 * Conventions
 *   - every function returns bcs_status (0 = ok); bcs_last_error() gives the message of the last failure
 *     on the calling thread.  The library never exits or throws across the boundary (the reference
 *     printf+exit()s: src/utilities/cuda_handle_error.cuh:18-25).
 *   - one bcs_sim per GPU (per rank); a handle is not re-entrant, distinct handles are independent.
 *   - all device work of a handle is enqueued on ONE stream (bcs_opts.stream or a library-owned one);
 *     calls are asynchronous unless they copy to host memory; bcs_synchronize() waits.
 *   - host arrays are SoA float x[],y[],z[] exactly like the reference's cudaVec3
 *     (src/utilities/cuda_vec3.cuh:11-73).
 *   - there is NO CPU fallback: without a CUDA device bcs_create fails with BCS_ERR_CUDA.
 */
#ifndef BCS_H
#define BCS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCS_ABI_VERSION 1
#define BCS_MAX_TYPES 16          /* maxCudaStreams, src/config/simulation.hpp:4 + static_assert blood_cell_factory.hpp:167 */
#define BCS_VEIN_MAX_NEIGHBORS 9  /* veinVertexMaxNeighbors, src/meta_factory/vein_factory.hpp:125 */

typedef enum bcs_status {
    BCS_OK = 0,
    BCS_ERR_INVALID = 1,      /* bad argument / malformed scene */
    BCS_ERR_CUDA = 2,         /* CUDA runtime error (message in bcs_last_error) */
    BCS_ERR_NOMEM = 3,
    BCS_ERR_UNSUPPORTED = 4,
    BCS_ERR_NCCL = 5,
    BCS_ERR_STATE = 6         /* call not valid in the current state (e.g. stage before grid build) */
} bcs_status;

/* ---------------------------------------------------------------------------------------------
 * Scene = the reference's compile-time src/config content as run-time data
 * ------------------------------------------------------------------------------------------- */

/* Spring<Start, End, Length> - src/meta_factory/blood_cells_def_type.hpp:39-45 */
typedef struct bcs_spring {
    int32_t start;
    int32_t end;
    float length;
} bcs_spring;

/* BloodCellDef<Count, ParticlesInCell, ..., Springs, Vertices, ...> - blood_cells_def_type.hpp:21-31.
 * One entry per element of UserDefinedBloodCellList (src/config/blood_cells_definition.hpp:13-25), in the
 * USER's order.  The library applies the fold / unique / sort of blood_cell_factory.hpp:60-162 itself. */
typedef struct bcs_cell_def {
    int32_t count;              /* blood cells of this definition */
    int32_t particles_in_cell;  /* P */
    int32_t n_springs;
    const bcs_spring* springs;  /* n_springs entries */
    const float* vertices;      /* 3*P floats, x,y,z per model vertex */
} bcs_cell_def;

/* src/config/physics.hpp:4-34, src/config/simulation.hpp:5-21, vein_factory.hpp:88 */
typedef struct bcs_physics {
    float dt;
    float velocity_collision_damping;
    float particle_k_sniff;
    float vein_k_sniff;
    float particle_d_fact;
    float vein_d_fact;
    float vein_boundaries_velocity_damping;   /* unused by the hot path (as in the reference) */
    float vein_collision_force_intensity;
    float viscous_damping;
    float collision_spring_coeff;
    float collision_damping_coeff;
    float collision_shear_coeff;
    float max_cell_size_factor_before_brake;
    float big_particle_braking_intensity;
    float init_velocity[3];
    float random_velocity_modifier;            /* initial-state generation only */
    float vein_impact_distance;
    float vein_impact_minimal_force_distance;
    float gravity[3];
    float grid_y_margin;
    float grid_xz_margin;
    float min_spawn_y;
    float cylinder_radius;
} bcs_physics;

typedef struct bcs_scene {
    uint32_t struct_size;       /* = sizeof(bcs_scene) */
    int32_t n_defs;
    const bcs_cell_def* defs;
    /* src/config/vein_definition.hpp: veinPositions, veinIndices, VeinEndingCenters, VeinEndingRadii */
    int32_t n_vertices;
    const float* vein_x;
    const float* vein_y;
    const float* vein_z;
    int32_t n_triangles;
    const uint32_t* vein_indices;   /* 3*n_triangles */
    int32_t n_endings;
    const float* ending_centers;    /* 3*n_endings */
    const float* ending_radii;      /* n_endings */
    /* src/config/simulation.hpp:10-21 */
    int32_t cell_size[3];           /* cellWidth, cellHeight, cellDepth */
    int32_t tri_cell_size[3];       /* cellWidthTriangles, ... */
    int32_t use_blood_flow;
    int32_t enable_reaction_force;
    int32_t enable_big_cells_brake;
    int32_t bounding_spheres_coeff;
    bcs_physics physics;
} bcs_scene;

/* Which observable behaviour the step reproduces (SURVEY.md 8(a) quirk register). */
typedef enum bcs_semantics {
    /* empty grid cells are empty, the last occupied cell is complete, every particle uses its own
     * type's collision radius.  Well defined under slab decomposition; the default. */
    BCS_SEM_CLEAN = 0,
    /* bit-compatible with the reference CUDA build: persistent never-cleared zero-initialised cell
     * tables (uniform_grid.cu:61-70,98-99), the last-cell write (uniform_grid.cu:76-79) and the
     * per-launch-slice radius lookup (particle_collisions.cuh:76,124).  Single GPU only. */
    BCS_SEM_REFERENCE = 1
} bcs_semantics;

typedef struct bcs_opts {
    uint32_t struct_size;       /* = sizeof(bcs_opts) */
    int32_t device;             /* CUDA device ordinal */
    int32_t semantics;          /* bcs_semantics */
    int32_t use_graph;          /* 1: bcs_step replays a captured CUDA graph (default); 0: plain launches */
    int32_t collect_stats;      /* 1: count pair tests / hits / triangle tests (slower) */
    int32_t exhaustive_vein_traversal; /* 1: test every triangle of the 27 cells in the reference's order (slow; for
                                          cross-checking the culled search, which yields the same result) */
    uint64_t seed;              /* counter-based respawn RNG seed (vein_end.cu:103-105 uses cuRAND seeded from time(0)) */
    void* stream;               /* cudaStream_t to enqueue on, or NULL for a library-owned stream */
} bcs_opts;

typedef struct bcs_sim bcs_sim;

/* ---------------------------------------------------------------------------------------------
 * Lifetime.  Stands in for the constructors of BloodCells, VeinTriangles, UniformGrid x2 and
 * sim::SimulationController (main.cu:122-138; blood_cells.cu:16-26; vein_triangles.cu:37-69;
 * uniform_grid.cu:83-104; simulation_controller.cu:42-76 incl. generateBoundingSpheres :93-153).
 * Unlike the reference, every device buffer is zero-initialised and NO random initial state is
 * generated: upload one with bcs_upload().
 * ------------------------------------------------------------------------------------------- */
int bcs_create(const bcs_scene* scene, const bcs_opts* opts, bcs_sim** out);
void bcs_destroy(bcs_sim* sim);
const char* bcs_last_error(void);
int bcs_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Derived tables (what src/meta_factory computes at compile time) - for verification and for the
 * renderer-side consumers (glcontroller.cu:72-73,110-111 read smallestRadiusInType).
 * ------------------------------------------------------------------------------------------- */
typedef struct bcs_type_info {
    int32_t count;              /* BloodCellDef::count after folding duplicates */
    int32_t particles_in_cell;
    int32_t particle_start;     /* particleStarts[i]          blood_cell_factory.hpp:197-216 */
    int32_t cell_start;         /* bloodCellTypesStarts[i]    :221-238 */
    int32_t model_start;        /* bloodCellModelStarts[i]    :243-260 */
    int32_t graph_start;        /* accumulatedGraphSizes[i]   :265-285 */
    int32_t src_def;            /* index into bcs_scene.defs this type took its springs/vertices from */
    int32_t vein_end_warp_sync; /* 1 if the reference would pick handleVeinEndsWarpSync (vein_end.cu:23-30) */
    float smallest_radius;      /* smallestRadiusInType[i]    simulation_controller.cu:126-136 */
} bcs_type_info;

typedef struct bcs_layout {
    int32_t n_types;
    int32_t n_particles;        /* particleCount */
    int32_t n_cells;            /* bloodCellCount */
    int32_t n_model;            /* particleDistinctCellsCount */
    int32_t n_graph;            /* totalGraphSize */
    int32_t n_vertices;
    int32_t n_triangles;
    int32_t grid_dims[3];       /* cellCountX/Y/Z of the particle grid   uniform_grid.cu:86-89 */
    int32_t grid_cells;
    int32_t tri_grid_dims[3];
    int32_t tri_grid_cells;
    float grid_min[3];          /* minX,minY,minZ                        vein_factory.hpp:46-76 */
    float grid_max[3];
    float grid_size[3];         /* width,height,depth                    vein_factory.hpp:78-86 */
    bcs_type_info types[BCS_MAX_TYPES];
} bcs_layout;

int bcs_get_layout(const bcs_sim* sim, bcs_layout* out);

typedef enum bcs_table {
    BCS_TABLE_SPRING_GRAPH = 0,   /* float[n_graph]   springGraph            blood_cell_factory.hpp:292-333 */
    BCS_TABLE_MODEL_X = 1,        /* float[n_model]   bloodCellModels        simulation_controller.cu:99-117 */
    BCS_TABLE_MODEL_Y = 2,
    BCS_TABLE_MODEL_Z = 3,
    BCS_TABLE_COLLISION_RADII = 4,/* float[n_model]   cellModelsBoundingSpheres  :119-145 */
    BCS_TABLE_INITIAL_RADII = 5,  /* float[n_model]   BloodCells::initialRadiuses :142-144 */
    BCS_TABLE_VEIN_NBR_IDS = 6,   /* int32[9*V]       [slot][vertex]         vein_factory.hpp:130-174 */
    BCS_TABLE_VEIN_NBR_LEN = 7,   /* float[9*V] */
    BCS_TABLE_TRI_CENTERS_X = 8,  /* float[T]         VeinTriangles::centers vein_triangles.cu:14-27 */
    BCS_TABLE_TRI_CENTERS_Y = 9,
    BCS_TABLE_TRI_CENTERS_Z = 10
} bcs_table;

/* Copies a derived table to host memory; dst_bytes must be at least the table size. */
int bcs_get_table(bcs_sim* sim, int table, void* dst, size_t dst_bytes);

/* ---------------------------------------------------------------------------------------------
 * State.  Stands in for direct access to the public device members
 * bloodCells.particles.{positions,velocities,forces}[0], triangles.{positions,velocities,forces}[0],
 * bloodCells.particleCenters[0] (particles.cuh:10-35, vein_triangles.cuh:53-55, blood_cells.cuh:30).
 * ------------------------------------------------------------------------------------------- */
typedef enum bcs_array {
    BCS_PARTICLE_POS = 0,
    BCS_PARTICLE_VEL = 1,
    BCS_PARTICLE_FRC = 2,
    BCS_VEIN_POS = 3,
    BCS_VEIN_VEL = 4,
    BCS_VEIN_FRC = 5,
    BCS_CELL_CENTERS = 6      /* download only */
} bcs_array;

/* Host SoA -> device (n = element count of that array).  Pinned host memory (bcs_host_alloc) makes the copy asynchronous:
 * the buffers must stay untouched until the stream has passed it; pageable buffers may be reused as soon as this returns. */
int bcs_upload(bcs_sim* sim, int array, const float* x, const float* y, const float* z, int32_t n);
/* Device -> host SoA; returns after the data has landed. */
int bcs_download(bcs_sim* sim, int array, float* x, float* y, float* z, int32_t n);

/* Device-resident view for zero-copy consumers (what the renderer maps: main.cu:183-184,
 * glcontroller.cu:168-202).  Particle and vein state is stored as float4 {x,y,z,w}; w is private. */
typedef struct bcs_device_view {
    void* particle_pos4;   /* float4[n_particles] */
    void* particle_vel4;
    void* particle_frc4;
    void* vein_pos4;       /* float4[n_vertices] */
    void* vein_vel4;
    void* vein_frc4;       /* between the vein-collision stage and the vein integrator of a step the wall splats are parked in
                            * fixed point and NOT yet part of this array (bcs_download folds them in); whole steps: complete */
    void* stream;          /* cudaStream_t all work of this handle is ordered on */
} bcs_device_view;
int bcs_device_ptrs(bcs_sim* sim, bcs_device_view* out);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers without a CUDA runtime binding. */
/* Headless frame export into the interleaved layouts the reference's renderer fills from the physics
 * state (graphics/glcontroller.cu:23-50: calculatePositionsKernel, calculateTriangleVerticesKernel): DEVICE pointers,
 * e.g. mapped GL buffers; any of them may be null.
 *   cell_vertices6  [6 * n_particles] xyz of particle i at [6 i .. 6 i + 2] (floats 3..5 - the normals - are left alone);
 *                   a per-type VBO of the reference is this buffer offset by 6 * particle_start of the type
 *   offsets3        [3 * n_particles] xyz of particle i at [3 i ..]
 *   vein_vertices6  [6 * n_vertices]  xyz of vein vertex v at [6 v .. 6 v + 2]
 * Asynchronous on the simulation's stream (bcs_device_ptrs().stream). */
int bcs_export_frame(bcs_sim* sim, float* cell_vertices6, float* offsets3, float* vein_vertices6);

int bcs_host_alloc(void** out, size_t bytes);
int bcs_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * The step.  Names follow the reference loop body (main.cu:175-176,199,208).
 * ------------------------------------------------------------------------------------------- */
/* particleGrid.calculateGrid(positions, particleCount) + triangleCentersGrid.calculateGrid(centers, T)
 * (uniform_grid.cu:129-155).  The triangle grid is built from the INITIAL centres, as in the
 * single-GPU reference (vein_triangles.cu:68), so it is computed once and reused. */
int bcs_build_grid(bcs_sim* sim);
/* sim::SimulationController::calculateNextFrame() (simulation_controller.cu:246-313) */
int bcs_compute_forces(bcs_sim* sim);
/* sim::SimulationController::propagateAll() (simulation_controller.cu:315-331) */
int bcs_integrate(bcs_sim* sim);
/* nsteps x { build_grid; compute_forces; integrate } without host round trips (CUDA graph replay).  Inside a run of
 * nsteps > 1 the end of step k and the spring stage of step k + 1 share one pass over the particle state; the state a
 * caller sees after the call is exactly that of nsteps single steps (bit for bit). */
int bcs_step(bcs_sim* sim, int32_t nsteps);
/* Waits for the handle's stream and reports sticky device-side errors with BCS_ERR_STATE (bcs_last_error names them): a
 * halo / migration message that overflowed its capacity (slab mode: records were dropped, the run is invalid - raise
 * bcs_slab_opts.migration_capacity / halo_capacity), an active particle outside the rank's window of grid rows (raise
 * halo_width), a wall grid that outgrew its lists.  bcs_download and bcs_get_stats report the same conditions. */
int bcs_synchronize(bcs_sim* sim);
/* number of completed steps (drives the respawn RNG counter) */
int bcs_get_step_count(const bcs_sim* sim, int64_t* out);
/* Checkpoint / restart (the reference has none): the complete dynamic state of a simulation is the six
 * state arrays (bcs_download / bcs_upload) plus the step count - the respawn RNG is counter based, keyed by
 * (seed, blood cell, step), so restoring the count restores the random stream.  In slab mode upload the merged global
 * state on every rank: ownership is re-derived from the positions.
 * BCS_SEM_CLEAN only: under BCS_SEM_REFERENCE the never-cleared particle-grid tables (stale ranges that
 * double-count pairs) are part of the dynamic state too, so a restored run does not continue bit-identically - restore
 * is still allowed there (the reference itself has no notion of it), the continuation just starts from clean tables. */
int bcs_set_step_count(bcs_sim* sim, int64_t steps);

/* Single stages, in the reference's order, for stage-by-stage parity tests. */
typedef enum bcs_stage {
    BCS_STAGE_GRID_PARTICLES = 0,      /* uniform_grid.cu:129-155 on particle positions */
    BCS_STAGE_GRID_TRIANGLES = 1,      /* same on triangle centres */
    BCS_STAGE_VEIN_GATHER = 2,         /* vein_triangles.cu:126-163 */
    BCS_STAGE_SPRINGS = 3,             /* blood_cells.cu:44-153 (centres + intra-cell springs) */
    BCS_STAGE_PARTICLE_COLLISIONS = 4, /* particle_collisions.cuh:104-269 */
    BCS_STAGE_VEIN_COLLISIONS = 5,     /* vein_collisions.cu:63-277 */
    BCS_STAGE_INTEGRATE_PARTICLES = 6, /* blood_cells.cu:155-186 */
    BCS_STAGE_INTEGRATE_VEIN = 7,      /* vein_triangles.cu:88-117 */
    BCS_STAGE_VEIN_END = 8             /* vein_end.cu:141-173 */
} bcs_stage;
int bcs_run_stage(bcs_sim* sim, int stage);

/* ---------------------------------------------------------------------------------------------
 * Grid inspection (bit-exact checks).  which_grid: 0 = particle grid, 1 = triangle grid.
 * Stands in for reading UniformGrid::{gridCellIds,particleIds,gridCellStarts,gridCellEnds}[0]
 * (uniform_grid.cuh:32-35).
 * ------------------------------------------------------------------------------------------- */
/* keys[n] = sorted cell ids, ids[n] = object ids in sorted order */
int bcs_download_grid(bcs_sim* sim, int which_grid, int32_t* keys, int32_t* ids, int32_t n);
/* Non-empty table entries in ascending cell order.  BCS_SEM_CLEAN: cells with end >= start.
 * BCS_SEM_REFERENCE: cells whose (start,end) != (0,0), i.e. everything ever written. */
int bcs_download_cell_table(bcs_sim* sim, int which_grid, int32_t capacity, int32_t* cells, int32_t* starts,
                            int32_t* ends, int32_t* out_count);
/* Per particle (indexed by particle id): number of collision candidates visited by the particle
 * collision stage for the CURRENT grid, a 64-bit order-independent checksum of the candidate particle
 * ids (sum of (id+1)*0x9E3779B97F4A7C15 mod 2^64), and the number of accepted collisions. */
int bcs_debug_candidates(bcs_sim* sim, int32_t* counts, uint64_t* checksums, int32_t* hits, int32_t n);
/* Per particle: id of the triangle the vein-collision traversal returns (-1: none) and its ray parameter t. */
int bcs_debug_vein_hits(bcs_sim* sim, int32_t* triangle, float* t, int32_t n);

typedef struct bcs_stats {
    uint64_t pair_tests;      /* candidate distance tests in the particle collision stage */
    uint64_t pair_hits;
    uint64_t triangle_tests;  /* ray/triangle tests in the vein collision stage */
    uint64_t vein_hits;
    uint64_t teleported_cells;
    uint64_t out_of_bounds;   /* positions outside the grid bounds seen by the cell-id stage */
    uint64_t wall_rebuilds;   /* rebuilds of the lazily maintained wall grid (vein-collision culling structure) */
} bcs_stats;
/* Totals since creation (pair/triangle counters only advance when bcs_opts.collect_stats = 1). */
int bcs_get_stats(bcs_sim* sim, bcs_stats* out);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: slab decomposition along the vein axis (y).  Stands in for the reference's -DMULTI_GPU build
 * (main.cu:105-117,162-173,194-204: ncclCommInitAll over 4 devices in ONE process, every array replicated,
 * ncclBroadcast of the state and of both grids + ncclReduce of the forces every frame).  Here: one process
 * and one bcs_sim per GPU; every rank is created from the same scene, uploads the same initial state and
 * then advances only the blood cells whose centre lies in its slab [y_lo, y_hi); ghost particles, migrating
 * blood cells and the vein-vertex halo travel in one grouped ncclSend/ncclRecv per neighbour and step.
 * Clean semantics only.  bcs_step() is the entry point (the staged calls do not exchange halos).
 * ------------------------------------------------------------------------------------------- */
typedef struct bcs_slab_opts {
    uint32_t struct_size;          /* = sizeof(bcs_slab_opts) */
    int32_t rank;
    int32_t world;
    int32_t spawn_rank;            /* rank whose slab contains min_spawn_y (respawned blood cells are sent there) */
    float y_lo, y_hi;              /* this rank's slab; use -INFINITY / +INFINITY for the outermost faces */
    float halo_width;              /* particle halo (<= 0: default 32) */
    float vertex_halo;             /* vein vertex / triangle halo (<= 0: default 2 triangle cells + halo_width + 20) */
    int32_t migration_capacity;    /* particle records per message (<= 0: default) */
    int32_t halo_capacity;         /* ghost records per message (<= 0: default) */
    char nccl_unique_id[128];      /* from bcs_nccl_unique_id() on rank 0, distributed by the caller */
} bcs_slab_opts;

/* ncclGetUniqueId(); the caller broadcasts the 128 bytes to all ranks (e.g. with torch.distributed). */
int bcs_nccl_unique_id(char out[128]);
/* Collective over all ranks (ncclCommInitRank inside). */
int bcs_create_slab(const bcs_scene* scene, const bcs_opts* opts, const bcs_slab_opts* slab, bcs_sim** out);
/* owned[c] = 1 if this rank currently owns blood cell c (n_cells entries); state arrays downloaded from a rank are
 * only meaningful for the blood cells it owns. */
int bcs_download_ownership(bcs_sim* sim, uint8_t* owned, int32_t n_cells);
/* Owned-only transfers: bcs_upload / bcs_download for the particle arrays (positions, velocities, forces) that move only
 * the particles of the blood cells this rank currently OWNS across the bus.  x, y, z are full-length arrays
 * (n = n_particles) in the reference's layout (cudaVec3, utilities/cuda_vec3.cuh); entries of other ranks' blood cells
 * are not read (upload) / not written (download), so N ranks working on the same host arrays fill them completely.
 * Ownership does not change in an upload (the next step decides about migration from the new positions); after an
 * upload of positions or velocities the next bcs_step first refreshes the neighbours' ghosts with one halo exchange -
 * every rank of the decomposition must therefore make the same sequence of calls.  With PINNED arrays (cudaHostAlloc /
 * cudaHostRegister) the kernels read and write the host arrays in place over the bus, walking the rank's owned-cell
 * lists, and the upload is asynchronous like bcs_upload; pageable arrays go through a host-side gather and a staging
 * buffer (correct, several times slower).  On a handle without slab decomposition these are bcs_upload / bcs_download. */
int bcs_upload_owned(bcs_sim* sim, int array, const float* x, const float* y, const float* z, int32_t n);
int bcs_download_owned(bcs_sim* sim, int array, float* x, float* y, float* z, int32_t n);
/* Number of active (owned + ghost) particles of the last grid build and of ghosts in the last exchange. */
int bcs_slab_counts(bcs_sim* sim, int32_t* active_particles, int32_t* ghost_particles, int32_t* owned_cells);

/* ---------------------------------------------------------------------------------------------
 * Measurement support (no reference counterpart: the reference only prints a wall-clock average at exit,
 * main.cu:248-253).
 * ------------------------------------------------------------------------------------------- */
/* Number of libbcs kernels enqueued on the handle's stream since creation (CUDA-graph replays count
 * their kernel nodes). */
int bcs_get_launch_count(bcs_sim* sim, uint64_t* kernels);

#define BCS_KERNEL_NAME_LEN 48
/* Runs nsteps full steps with plain launches, bracketing EVERY kernel with CUDA events on the handle's
 * stream, and returns per kernel name the summed device time (ms) and launch count, in first-launch
 * order.  Synchronous.  The simulation state advances by nsteps. */
int bcs_profile_steps(bcs_sim* sim, int32_t nsteps, int32_t capacity, char (*names)[BCS_KERNEL_NAME_LEN], float* ms_total,
                      int32_t* launches, int32_t* out_count);

#ifdef __cplusplus
}
#endif
#endif /* BCS_H */

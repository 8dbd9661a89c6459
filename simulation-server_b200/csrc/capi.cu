// C ABI of libbcs (include/bcs.h): handle lifetime, state transfer, stage sequencing, CUDA-graph replay.
//
// Stage order of one step (main.cu:175-208 + simulation_controller.cu:246-331):
//   grid(particles) [grid(triangles) is static] -> vein springs -> cell springs -> particle collisions ->
//   vein collisions -> integrate particles -> integrate vein (+clear forces) -> vein end
// The reference separates the stages with 11 device-wide cudaDeviceSynchronize calls and launches each
// particle stage once per blood-cell type; here everything is ordered on one stream, one launch per stage
// covers all types, and bcs_step replays the whole step as a captured CUDA graph.
#include <cstring>
#include <memory>
#include <mutex>
#include <string>

#include "bcs_internal.cuh"
#include "device_math.cuh"
#include "kernels.cuh"
#include "slab.cuh"

namespace bcs {

static thread_local std::string g_lastError;
void set_error(const std::string& msg) { g_lastError = msg; }

static thread_local LaunchCtx* g_launchCtx = nullptr;
LaunchCtx* current_launch_ctx() { return g_launchCtx; }
void set_launch_ctx(LaunchCtx* c) { g_launchCtx = c; }
void launch_begin(const char* name, cudaStream_t st)
{
    LaunchCtx* c = g_launchCtx;
    if (!c) return;
    ++c->launches;
    if (c->timing) {
        LaunchRecord r{name, nullptr, nullptr};
        cudaEventCreate(&r.start);
        cudaEventCreate(&r.stop);
        cudaEventRecord(r.start, st);
        c->records.push_back(r);
    }
}
void launch_end(cudaStream_t st)
{
    LaunchCtx* c = g_launchCtx;
    if (c && c->timing && !c->records.empty()) cudaEventRecord(c->records.back().stop, st);
}
struct CtxScope {
    explicit CtxScope(LaunchCtx* c) { set_launch_ctx(c); }
    ~CtxScope() { set_launch_ctx(nullptr); }
};

template <class T>
static T* dev_alloc(size_t count, bool zero = true)
{
    T* p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) throw Error{BCS_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e)};
    if (zero) {
        // cudaMemset runs on the legacy default stream and is asynchronous for device memory, while all work of a handle
        // runs on a NON-BLOCKING stream, which does not wait for it: without this drain a kernel could write into the
        // buffer before the zero fill lands (seen as rare zeroed debug outputs on a cold first run)
        BCS_CUDA(cudaMemset(p, 0, count * sizeof(T)));
        BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    return p;
}
template <class T>
static T* dev_upload(const std::vector<T>& v)
{
    T* p = dev_alloc<T>(v.size(), false);
    if (!v.empty()) {
        // A cudaMemcpy from PAGEABLE host memory returns once the source has been staged; the DMA into device memory may
        // still be in flight on the legacy stream, which the handle's non-blocking stream does not wait for.  Drain it:
        // a kernel launched right after must see the whole table.
        BCS_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    return p;
}

// ---- SoA <-> float4 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                                   float4* __restrict__ out, int n, const TypesDev types, const float* __restrict__ collR,
                                                   int setRadius)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float w = 0.f;
    if (setRadius) {
        // pos4.w = the particle's own collision radius (bounding sphere of its model vertex)
        int t = 0;
        while (t + 1 < types.n && i >= types.t[t + 1].pStart) ++t;
        w = collR[types.t[t].mStart + (i - types.t[t].pStart) % types.t[t].P];
    }
    out[i] = make_float4(x[i], y[i], z[i], w);
}
__global__ void __launch_bounds__(256) unpack_kernel(const float4* __restrict__ in, float* __restrict__ x, float* __restrict__ y,
                                                     float* __restrict__ z, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    x[i] = v.x; y[i] = v.y; z[i] = v.z;
}
// owned-only transfers (slab mode): compacted SoA <-> float4 through the list of owned particle ids
__global__ void __launch_bounds__(256) pack_owned_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                                         const int* __restrict__ idx, int n, float4* __restrict__ out, const TypesDev types,
                                                         const float* __restrict__ collR, int setRadius)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    float w = 0.f;
    if (setRadius) {
        int t = 0;
        while (t + 1 < types.n && i >= types.t[t + 1].pStart) ++t;
        w = collR[types.t[t].mStart + (i - types.t[t].pStart) % types.t[t].P];
    }
    out[i] = make_float4(x[k], y[k], z[k], w);
}
__global__ void __launch_bounds__(256) unpack_owned_kernel(const float4* __restrict__ in, const int* __restrict__ idx, int n, float* __restrict__ x,
                                                           float* __restrict__ y, float* __restrict__ z)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 v = in[idx[k]];
    x[k] = v.x; y[k] = v.y; z[k] = v.z;
}
// ... and, when the caller's arrays are pinned host memory, straight between them and the device arrays (zero copy over the
// bus: the kernel walks the rank's owned-cell lists - the 20-odd particles of a blood cell are one contiguous 80-byte run
// per component - so nothing is gathered on the host and no ownership table has to be read back first)
__global__ void __launch_bounds__(256) upload_owned_direct_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                                                  const ActiveItems items, float4* __restrict__ out, const float* __restrict__ collR,
                                                                  int setRadius)
{
    const TypesDev* types = items.types;
    const int total = items.cellPrefix[types->n] * items.maxP;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        int fl = 0;
        const int i = active_item(items, item, fl);
        if (i < 0) continue;
        float w = 0.f;
        if (setRadius) {
            int t = 0;
            while (t + 1 < types->n && i >= types->t[t + 1].pStart) ++t;
            w = collR[types->t[t].mStart + (i - types->t[t].pStart) % types->t[t].P];
        }
        out[i] = make_float4(x[i], y[i], z[i], w);
    }
}
__global__ void __launch_bounds__(256) download_owned_direct_kernel(const float4* __restrict__ in, const ActiveItems items, float* __restrict__ x,
                                                                    float* __restrict__ y, float* __restrict__ z)
{
    const int total = items.cellPrefix[items.types->n] * items.maxP;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        int fl = 0;
        const int i = active_item(items, item, fl);
        if (i < 0) continue;
        const float4 v = in[i];
        x[i] = v.x; y[i] = v.y; z[i] = v.z;
    }
}
// graphics/glcontroller.cu:23-50 equivalents: xyz into a strided float buffer (stride 6: interleaved with normals; 3: offsets)
__global__ void __launch_bounds__(256) export_xyz_kernel(const float4* __restrict__ in, int n, float* __restrict__ out6, float* __restrict__ out3)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    if (out6) { out6[6 * i] = p.x; out6[6 * i + 1] = p.y; out6[6 * i + 2] = p.z; }
    if (out3) { out3[3 * i] = p.x; out3[3 * i + 1] = p.y; out3[3 * i + 2] = p.z; }
}

__global__ void fill_int_kernel(int* p, int v, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
// sparse view of a dense cell table, ascending cell order (single block per chunk keeps it simple: debug path)
__global__ void mark_cells_kernel(const int* __restrict__ cs, const int* __restrict__ ce, int cells, int reference, int* __restrict__ flags)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int s = cs[c], e = ce[c];
    flags[c] = reference ? (s != 0 || e != 0) : (e >= s);
}

}  // namespace bcs

using namespace bcs;

struct bcs_sim {
    HostScene hs;
    int device = 0;
    int semantics = BCS_SEM_CLEAN;
    bool useGraph = true, stats = false, ownStream = false;
    unsigned long long seed = 0;
    cudaStream_t stream = nullptr;
    // particles
    float4 *pos = nullptr, *vel = nullptr, *frc = nullptr, *spos = nullptr, *svel = nullptr, *centers = nullptr;
    int *keys[2] = {nullptr, nullptr}, *ids[2] = {nullptr, nullptr}, *cellStart = nullptr, *cellEnd = nullptr;
    // compact cell index of the particle grid (clean semantics; replaces cellStart/cellEnd)
    unsigned* cellMask = nullptr;
    int *cellRank = nullptr, *occStart = nullptr, *occKey = nullptr, *numOcc = nullptr;
    int maskWords = 0;
    // vein
    float4 *vpos = nullptr, *vvel = nullptr, *vfrc = nullptr, *tcent = nullptr;
    long long* vsplat = nullptr;   // [3V] fixed-point wall-splat accumulators (vein_device.cuh: splat_add)
    int *tkeys[2] = {nullptr, nullptr}, *tids[2] = {nullptr, nullptr}, *tcellStart = nullptr, *tcellEnd = nullptr;
    TriPacked* tris = nullptr;
    Aabb *groupBox = nullptr, *cellBox = nullptr;
    CullEntry* cullList = nullptr;
    CellSlab *cellSlab = nullptr, *groupSlab = nullptr;
    int* cullCount = nullptr;
    unsigned* doneBlocks = nullptr;
    TypesDev* typesDev = nullptr;   // device copy of `types` for kernels that index it through a pointer
    int maxP = 1;
    bool exhaustiveVein = false;
    WallGridDev wall{};             // lazily rebuilt wall grid (clean semantics; wall.enabled = 0 otherwise)
    // independent stages of a step run on forked streams (graph branches when captured): springs | wall search | vein gather
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    bool nearProbe = true;      // BCS_NO_NEAR_PROBE: the wall filter streams every particle itself
    cudaEvent_t evFork = nullptr, evSprings = nullptr, evWall = nullptr, evGather = nullptr, evMasked = nullptr, evVein = nullptr;
    bool overlap = true;
    int numSMs = 148;
    unsigned* vidx = nullptr;
    int* nbrIds = nullptr;
    float* nbrLen = nullptr;
    // tables
    float *collR = nullptr, *initR = nullptr, *mx = nullptr, *my = nullptr, *mz = nullptr, *endC = nullptr, *endR = nullptr;
    int *adjJ = nullptr, *adjS = nullptr, *sprAB = nullptr;
    float *adjL = nullptr, *sprL = nullptr;
    Counters* counters = nullptr;
    float* staging = nullptr;   // 3 * maxLen floats
    int stagingLen = 0;
    SortScratch sortP, sortT;
    TypesDev types{};
    GridDev pg{}, tg{};
    PhysDev phys{};
    SpringPlan plan{};
    int* slotTab = nullptr;         // per-type tables of the cell pass (cellpass.cu)
    int2* adjTab = nullptr;
    // row-directory grid + symmetric pair search (clean semantics, sparse scenes; grid.cu / pairs.cu)
    RowsGrid rows{};
    PairLists pairs{};
    unsigned long long* phaseClock = nullptr;   // BCS_CP_CLOCK: per-phase cycle sums of the cell pass (developer aid)
    bool collideWalk = false;       // BCS_COLLIDE=walk: every slot scans its whole stencil (A/B partner of the pair search)
    bool inStep = false;            // inside bcs_step: the fold of the parked pair forces rides in the wall apply / cell pass
    bool fuseSteps = false;         // bcs_step(n): end of step k + springs / row count of step k + 1 in one pass (cellpass.cu)
    bool gridBuilt = false;
    cudaGraphExec_t graphExec = nullptr;
    cudaGraphExec_t graphMid = nullptr, graphLast = nullptr;   // fused run: a step that hands over to the next one / the last step
    bool gridModeSettled = false;
    // owned-only transfers (bcs_upload_owned / bcs_download_owned, slab mode)
    std::vector<unsigned char> ownedHost;       // ownership flags as of the last refresh
    std::vector<int2> ownedRuns;                // (first particle, count): maximal runs of owned particles in id order
    int nOwnedParticles = 0;
    bool ownedHostValid = false, haloStale = false;
    int* ownedIdxDev = nullptr;                 // [N] particle id of every compacted slot
    int* ownedIdxHost = nullptr;                // pinned
    float* hostStage = nullptr;                 // pinned, 3 x N floats   // the row directory has been confirmed (or dropped) against the first uploaded positions
    unsigned long long kernelsMid = 0, kernelsLast = 0;
    cudaEvent_t evRebuilt = nullptr, evGrid = nullptr;
    LaunchCtx ctx;
    unsigned long long kernelsPerGraph = 0;
    SlabState* slab = nullptr;   // multi-GPU slab mode
    std::vector<void*> owned;

    template <class T> T* track(T* p) { owned.push_back((void*)p); return p; }
};

namespace {

int bits_for(int cells)
{
    int b = 1;
    while ((1ll << b) < cells) ++b;
    return b;
}

GridDev make_grid(const HostScene& hs, const int cs[3], const int dims[3], int n)
{
    GridDev g{};
    g.minx = hs.gmin[0]; g.miny = hs.gmin[1]; g.minz = hs.gmin[2];
    g.maxx = hs.gmax[0]; g.maxy = hs.gmax[1]; g.maxz = hs.gmax[2];
    g.lenx = hs.gsize[0]; g.leny = hs.gsize[1]; g.lenz = hs.gsize[2];
    g.csx = cs[0]; g.csy = cs[1]; g.csz = cs[2];
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    g.cells = dims[0] * dims[1] * dims[2];
    g.n = n;
    g.keyBits = bits_for(g.cells);
    g.icsx = 1.0f / (float)cs[0]; g.icsy = 1.0f / (float)cs[1]; g.icsz = 1.0f / (float)cs[2];
    g.pow2 = ((cs[0] & (cs[0] - 1)) == 0 && (cs[1] & (cs[1] - 1)) == 0 && (cs[2] & (cs[2] - 1)) == 0) ? 1 : 0;
    return g;
}

void fill_phys(const HostScene& hs, PhysDev& p)
{
    const bcs_physics& s = hs.ph;
    p.dt = s.dt;
    p.velocity_collision_damping = s.velocity_collision_damping;
    p.particle_k_sniff = s.particle_k_sniff; p.vein_k_sniff = s.vein_k_sniff;
    p.particle_d_fact = s.particle_d_fact; p.vein_d_fact = s.vein_d_fact;
    p.vein_collision_force_intensity = s.vein_collision_force_intensity;
    p.viscous_damping = s.viscous_damping;
    p.coll_spring = s.collision_spring_coeff; p.coll_damping = s.collision_damping_coeff; p.coll_shear = s.collision_shear_coeff;
    p.max_cell_size_factor = s.max_cell_size_factor_before_brake; p.big_brake_intensity = s.big_particle_braking_intensity;
    p.initvx = s.init_velocity[0]; p.initvy = s.init_velocity[1]; p.initvz = s.init_velocity[2];
    p.impact2 = s.vein_impact_distance * s.vein_impact_distance;
    p.impactNear = s.vein_impact_distance * 1.001f + 0.01f;
    p.minForce2 = s.vein_impact_minimal_force_distance * s.vein_impact_minimal_force_distance;
    p.gx = s.gravity[0]; p.gy = s.gravity[1]; p.gz = s.gravity[2];
    p.min_spawn_y = s.min_spawn_y; p.cylinder_radius = s.cylinder_radius;
    // vein_end.cu:12-18
    p.upperY = hs.gmax[1] - 3 * s.grid_y_margin / 4;
    p.lowerY = hs.gmin[1] + s.grid_y_margin / 2;
    p.rightX = hs.gmax[0] - s.grid_xz_margin / 2;
    p.leftX = hs.gmin[0] + s.grid_xz_margin / 2;
    p.frontZ = hs.gmax[2] - s.grid_xz_margin / 2;
    p.backZ = hs.gmin[2] + s.grid_xz_margin / 2;
    p.useBloodFlow = hs.useBloodFlow; p.reactionForce = hs.reactionForce; p.bigBrake = hs.bigBrake;
    p.nEndings = (int)hs.endR.size();
}

// the parked pair forces (pairs.cu) are folded by the passes that read the forces next instead of by a pass of their own
bool defer_fold(const bcs_sim* s)
{
    if (!s->rows.enabled || !s->inStep || s->collideWalk) return false;
    const char* e = getenv("BCS_DEFER_FOLD");
    return e ? atoi(e) != 0 : s->slab != nullptr;
}

// ---- stage launchers ------------------------------------------------------------------------------------
GridBuildArgs particle_grid_args(bcs_sim* s)
{
    GridBuildArgs a{};
    a.grid = s->pg; a.objPos = s->pos;
    a.keys[0] = s->keys[0]; a.keys[1] = s->keys[1]; a.ids[0] = s->ids[0]; a.ids[1] = s->ids[1];
    a.cellStart = s->cellStart; a.cellEnd = s->cellEnd;
    a.scratch = &s->sortP; a.counters = s->counters;
    a.reference = s->semantics == BCS_SEM_REFERENCE;
    a.tablesValid = true;
    a.compact = !a.reference;
    a.cellMask = s->cellMask; a.maskWords = s->maskWords; a.cellRank = s->cellRank;
    a.occStart = s->occStart; a.occKey = s->occKey; a.numOcc = s->numOcc;
    a.reorder = true;
    a.pos = s->pos; a.vel = s->vel; a.spos = s->spos; a.svel = s->svel;
    a.rows = s->rows;
    if (s->slab) {
        a.pflag = s->slab->pflag; a.nDev = s->slab->nActive; a.nDevOut = s->slab->nActive;
        a.items.cells = s->slab->listCells; a.items.cellPrefix = s->slab->listCellPrefix;
        a.items.ghostList = s->slab->ghostList; a.items.ghostCount = s->slab->ghostCount;
        a.items.types = s->typesDev; a.items.maxP = s->maxP;
        a.itemCapacity = (long long)s->hs.B * s->maxP + s->hs.N;
    }
    return a;
}

void build_triangle_grid(bcs_sim* s)
{
    GridBuildArgs a{};
    a.grid = s->tg; a.objPos = s->tcent;
    a.keys[0] = s->tkeys[0]; a.keys[1] = s->tkeys[1]; a.ids[0] = s->tids[0]; a.ids[1] = s->tids[1];
    a.cellStart = s->tcellStart; a.cellEnd = s->tcellEnd;
    a.scratch = &s->sortT; a.counters = s->counters;
    a.reference = s->semantics == BCS_SEM_REFERENCE;
    a.tablesValid = true;
    a.compact = false;
    a.reorder = false;
    launch_grid_build(a, s->stream);
}

// Lazily rebuilt wall grid (wall.cu): geometry from the rest positions of the vein, list capacity from a host-side
// count of the (triangle, cell) overlaps with the same padding the device uses.
void setup_wall(bcs_sim* s)
{
    const HostScene& hs = s->hs;
    const int V = hs.V, T = hs.T;
    WallGridDev& w = s->wall;
    const char* em = getenv("BCS_WALL_MARGIN");
    const char* eh = getenv("BCS_WALL_CELL");
    w.margin = em ? (float)atof(em) : 0.25f;
    w.h = eh ? (float)atof(eh) : 4.0f;
    BCS_REQUIRE(w.margin > 0.f && w.h >= 1.f, BCS_ERR_INVALID, "bad BCS_WALL_MARGIN / BCS_WALL_CELL");
    w.h = std::max(w.h, 0.51f * s->phys.impactNear + 0.01f);   // phase A relies on reach <= 2 cells
    w.invh = 1.0f / w.h;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int v = 0; v < V; ++v) {
        const float p[3] = {hs.vx[v], hs.vy[v], hs.vz[v]};
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    const float slack = 2.0f + w.margin;   // beyond it cells are clamped (still conservative, only slower)
    int dims[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] -= slack; hi[k] += slack;
        dims[k] = std::max(1, (int)std::ceil((hi[k] - lo[k]) * w.invh));
    }
    w.ox = lo[0]; w.oy = lo[1]; w.oz = lo[2];
    w.nx = dims[0]; w.ny = dims[1]; w.nz = dims[2];
    const long long cells = (long long)dims[0] * dims[1] * dims[2];
    BCS_REQUIRE(cells < (1ll << 30), BCS_ERR_UNSUPPORTED, "wall grid too large (raise BCS_WALL_CELL)");
    w.cells = (int)cells;
    // capacity: exact overlap count at rest x 1.5 (+ header words)
    const float pad = 0.05f + w.margin;
    auto axis = [&](float p, int k) { float q = std::floor((p - lo[k]) * w.invh); return (int)std::min(std::max(q, 0.f), (float)(dims[k] - 1)); };
    long long entries = 0;
    std::vector<unsigned char> seen((size_t)cells, 0);
    long long nonEmpty = 0;
    for (int t = 0; t < T; ++t) {
        int c0[3], c1[3];
        for (int k = 0; k < 3; ++k) {
            const std::vector<float>& a = k == 0 ? hs.vx : k == 1 ? hs.vy : hs.vz;
            const float p0 = a[hs.vidx[3 * t]], p1 = a[hs.vidx[3 * t + 1]], p2 = a[hs.vidx[3 * t + 2]];
            c0[k] = axis(std::min(p0, std::min(p1, p2)) - pad, k);
            c1[k] = axis(std::max(p0, std::max(p1, p2)) + pad, k);
        }
        for (int z = c0[2]; z <= c1[2]; ++z)
            for (int y = c0[1]; y <= c1[1]; ++y)
                for (int x = c0[0]; x <= c1[0]; ++x) {
                    ++entries;
                    unsigned char& f = seen[((size_t)z * dims[1] + y) * dims[0] + x];
                    if (!f) { f = 1; ++nonEmpty; }
                }
    }
    const long long cap = entries * 3 / 2 + 4096;
    (void)nonEmpty;
    BCS_REQUIRE(cap < (1ll << 31), BCS_ERR_UNSUPPORTED, "wall grid lists too large (raise BCS_WALL_CELL)");
    w.cap = (int)cap;
    w.start = s->track(dev_alloc<int>((size_t)w.cells + 1));
    w.cursor = s->track(dev_alloc<int>((size_t)w.cells));
    w.rec = s->track(dev_alloc<int4>(2 * (size_t)w.cells));
    w.near = s->track(dev_alloc<unsigned char>((size_t)w.cells));
    w.occ = s->track(dev_alloc<unsigned char>((size_t)w.cells));
    w.occ3 = s->track(dev_alloc<unsigned char>((size_t)w.cells));
    w.nearTmp = s->track(dev_alloc<unsigned char>(2 * (size_t)w.cells));
    BCS_CUDA(cudaMemset(w.near, 0, (size_t)w.cells));
    BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    w.list = s->track(dev_alloc<int>((size_t)w.cap));
    w.vposBuilt = s->track(dev_alloc<float4>(V));
    int4* info = s->track(dev_alloc<int4>(T));
    int4* verts = s->track(dev_alloc<int4>(T));
    w.slotInfo = info;
    w.slotVerts = verts;
    std::vector<Aabb> inv((size_t)std::max((T + 7) / 8, s->tg.cells), Aabb{3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f});
    w.groupBox = s->track(dev_alloc<Aabb>((T + 7) / 8));
    w.cellBox = s->track(dev_alloc<Aabb>(s->tg.cells));
    BCS_CUDA(cudaMemcpy(w.groupBox, inv.data(), (size_t)((T + 7) / 8) * sizeof(Aabb), cudaMemcpyHostToDevice));
    BCS_CUDA(cudaMemcpy(w.cellBox, inv.data(), (size_t)s->tg.cells * sizeof(Aabb), cudaMemcpyHostToDevice));
    BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));   // pageable source: see dev_upload
    w.dirty = s->track(dev_alloc<int>(1));
    w.overflow = s->track(dev_alloc<int>(1));
    w.barrier = s->track(dev_alloc<unsigned>(2));
    w.builds = s->track(dev_alloc<unsigned long long>(1));
    w.queueCount = s->track(dev_alloc<int>(3));
    w.entryCount = w.queueCount + 1;
    w.nearCount = w.queueCount + 2;
    w.nearList = s->track(dev_alloc<int>((size_t)s->hs.N, false));
    cudaDeviceProp prop{};
    BCS_CUDA(cudaGetDeviceProperties(&prop, s->device));
    s->numSMs = prop.multiProcessorCount;
    w.blockSums = s->track(dev_alloc<int>(s->numSMs));
    w.queue = s->track(dev_alloc<int>(hs.N));
    w.best = s->track(dev_alloc<unsigned long long>(hs.N));
    w.ghostFlag = s->track(dev_alloc<unsigned char>(hs.N));
    w.entryCap = 4 * hs.N + 1024;
    w.entries = s->track(dev_alloc<int2>((size_t)w.entryCap));
    BCS_CUDA(cudaMemset(w.start, 0, ((size_t)w.cells + 1) * sizeof(int)));
    BCS_CUDA(cudaMemset(w.overflow, 0, sizeof(int)));
    BCS_CUDA(cudaMemset(w.barrier, 0, 2 * sizeof(unsigned)));
    BCS_CUDA(cudaMemset(w.builds, 0, sizeof(unsigned long long)));
    BCS_CUDA(cudaMemset(w.dirty, 1, sizeof(int)));
    BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));   // the fills above vs this handle's non-blocking stream
    launch_wall_slot_info(s->tkeys[1], s->tids[1], s->vidx, T, s->tg, info, verts, s->stream);
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    w.enabled = 1;
}

VeinArgs vein_args(bcs_sim* s)
{
    VeinArgs a{};
    a.V = s->hs.V; a.T = s->hs.T; a.phys = s->phys;
    a.vpos = s->vpos; a.vvel = s->vvel; a.vfrc = s->vfrc; a.vsplat = s->vsplat;
    a.nbrIds = s->nbrIds; a.nbrLen = s->nbrLen; a.vidx = s->vidx;
    a.vOwned = s->slab ? s->slab->vOwned : nullptr;
    a.vFirst = s->slab ? s->slab->vFirst : 0;
    a.vCount = s->slab ? s->slab->vCount : s->hs.V;
    if (s->wall.enabled) { a.vposBuilt = s->wall.vposBuilt; a.wallMargin = s->wall.margin; a.wallDirty = s->wall.dirty; }
    return a;
}

VeinCollideArgs vein_collide_args(bcs_sim* s)
{
    VeinCollideArgs a{};
    a.tgrid = s->tg; a.types = s->types; a.phys = s->phys;
    a.n = s->hs.N; a.T = s->hs.T;
    a.pos = s->pos; a.vel = s->vel; a.frc = s->frc;
    a.vpos = s->vpos; a.vfrc = s->vfrc; a.vsplat = s->vsplat; a.vidx = s->vidx;
    a.triIds = s->tids[1]; a.cellStart = s->tcellStart; a.cellEnd = s->tcellEnd;
    a.tris = s->tris; a.groupBox = s->groupBox; a.cellBox = s->cellBox; a.cellSlab = s->cellSlab; a.groupSlab = s->groupSlab; a.fast = !s->exhaustiveVein;
    a.nCells = s->hs.B; a.maxP = s->maxP; a.cullList = s->cullList; a.cullCount = s->cullCount;
    a.collR = s->collR; a.counters = s->counters;
    a.stats = s->stats; a.apply = true; a.dbgTri = nullptr; a.dbgT = nullptr;
    a.wall = s->wall;
    if (s->exhaustiveVein) a.wall.enabled = 0;
    if (defer_fold(s)) a.pairAcc = s->pairs.acc;
    if (s->slab) {
        a.pflag = s->slab->pflag;
        a.groupLocal = s->slab->groupLocal; a.triCellLocal = s->slab->triCellLocal; a.lists = slab_lists(s->slab, s->types);
        a.ghostList = s->slab->ghostList; a.ghostCount = s->slab->ghostCount;
        a.items = particle_grid_args(s).items;
    }
    return a;
}

CollideArgs collide_args(bcs_sim* s)
{
    CollideArgs a{};
    a.grid = s->pg; a.types = s->types; a.phys = s->phys; a.n = s->hs.N;
    a.keys = s->keys[1]; a.spos = s->spos; a.svel = s->svel;
    a.cellStart = s->cellStart; a.cellEnd = s->cellEnd; a.collR = s->collR;
    a.cellMask = s->cellMask; a.cellRank = s->cellRank; a.occStart = s->occStart;
    // BCS_COLLIDE=tiled: shared-memory staged variant (collide.cu).  Measured slower than the index walk at the bench's
    // ~1 particle per occupied cell (112 vs 70 us at 1 M particles), so it is opt-in.
    {
        const char* mode = getenv("BCS_COLLIDE");
        a.tiled = mode && std::string(mode) == "tiled" && !s->sortP.radixForCompact;
        a.rows = mode && std::string(mode) == "rows";
    }
    {
        int l = 0;
        while ((1ll << l) < s->pg.nx) ++l;
        a.nxShift = 32 + l;
        a.nxMagic = ((1ull << a.nxShift) + (unsigned long long)s->pg.nx - 1ull) / (unsigned long long)s->pg.nx;
    }
    a.frc = s->frc; a.counters = s->counters;
    a.reference = s->semantics == BCS_SEM_REFERENCE; a.stats = s->stats;
    a.dbgCount = nullptr; a.dbgSum = nullptr; a.dbgHits = nullptr;
    a.nDev = s->slab ? s->slab->nActive : nullptr;
    a.rowsMode = s->rows.enabled != 0; a.fullWalk = s->collideWalk;
    // deferred fold: in slab mode (the stand-alone fold sweeps all N particles of the global scene); on one GPU the fold
    // kernel measured cheaper than the extra loads in the cell pass.  BCS_DEFER_FOLD=1 / 0 overrides.
    a.deferFold = defer_fold(s);
    a.rowStart = s->rows.rowStart; a.rowsGrid = s->rows; a.nRows = s->rows.nRows; a.ids = s->ids[1]; a.vel = s->vel;
    a.irregular = s->rows.irregular ? s->rows.irregular + 1 : nullptr;   // latched copy (row_order_kernel)
    a.pairs = s->pairs;
    return a;
}

IntegrateArgs integrate_args(bcs_sim* s)
{
    IntegrateArgs a{};
    a.types = s->types; a.phys = s->phys; a.n = s->hs.N; a.nCells = s->hs.B;
    a.pos = s->pos; a.vel = s->vel; a.frc = s->frc;
    a.mx = s->mx; a.my = s->my; a.mz = s->mz; a.endC = s->endC; a.endR = s->endR;
    a.counters = s->counters; a.seed = s->seed;
    if (s->slab) { a.slab = s->slab->dev; a.lists = slab_lists(s->slab, s->types); a.moveTo = s->slab->moveTo; }
    if (defer_fold(s)) a.pairAcc = s->pairs.acc;
    return a;
}

SpringArgs spring_args(bcs_sim* s, bool withProbe = false);

void stage(bcs_sim* s, int st)
{
    switch (st) {
    case BCS_STAGE_GRID_PARTICLES:
        launch_grid_build(particle_grid_args(s), s->stream);
        s->gridBuilt = true;
        break;
    case BCS_STAGE_GRID_TRIANGLES:
        break;   // static: built at creation from the initial centres
    case BCS_STAGE_VEIN_GATHER: launch_vein_gather(vein_args(s), s->stream); break;
    case BCS_STAGE_SPRINGS: {
        launch_springs(spring_args(s), s->stream);
        break;
    }
    case BCS_STAGE_PARTICLE_COLLISIONS:
        BCS_REQUIRE(s->gridBuilt, BCS_ERR_STATE, "particle collisions need a built grid (bcs_build_grid)");
        if (s->rows.enabled) launch_particle_collisions_rows(collide_args(s), s->stream);
        else launch_particle_collisions(collide_args(s), s->stream);
        break;
    case BCS_STAGE_VEIN_COLLISIONS: {
        VeinCollideArgs a = vein_collide_args(s);
        if (a.wall.enabled) {
            launch_wall_rebuild(a, s->hs.V, s->numSMs, s->stream);
            launch_wall_collisions(a, s->stream);
        } else {
            launch_tri_refit(a, s->stream);
            launch_vein_collisions(a, s->stream);
        }
        break;
    }
    case BCS_STAGE_INTEGRATE_PARTICLES: launch_integrate_particles(integrate_args(s), s->stream); break;
    case BCS_STAGE_INTEGRATE_VEIN: launch_vein_integrate(vein_args(s), s->stream); break;
    case BCS_STAGE_VEIN_END: launch_vein_end(integrate_args(s), s->stream); break;
    default: throw Error{BCS_ERR_INVALID, "unknown stage"};
    }
}

SlabCtx slab_ctx(bcs_sim* s)
{
    SlabCtx c{};
    c.types = s->types; c.N = s->hs.N; c.B = s->hs.B; c.V = s->hs.V; c.T = s->hs.T;
    c.pos = s->pos; c.vel = s->vel; c.frc = s->frc; c.vpos = s->vpos; c.vvel = s->vvel; c.centers = s->centers;
    c.plan = s->plan;
    c.maxP = s->maxP;
    c.typesDev = s->typesDev;
    if (s->wall.enabled) { c.wallBuilt = s->wall.vposBuilt; c.wallMargin = s->wall.margin; c.wallDirty = s->wall.dirty; }
    c.stream = s->stream;
    return c;
}

SpringArgs spring_args(bcs_sim* s, bool withProbe)
{
    SpringArgs a{};
    a.types = s->types; a.typesDev = s->typesDev; a.plan = s->plan; a.phys = s->phys;
    a.pos = s->pos; a.vel = s->vel; a.frc = s->frc; a.centers = s->centers;
    a.slotTab = s->slotTab; a.adjTab = s->adjTab; a.initR = s->initR;
    if (s->slab) a.lists = slab_lists(s->slab, s->types);
    a.phaseClock = s->phaseClock;
    if (withProbe && s->wall.enabled) {
        const WallGridDev& w = s->wall;
        a.probe = NearProbe{w.near, w.nearList, w.nearCount, w.ox, w.oy, w.oz, w.invh, w.nx, w.ny, w.nz};
    }
    return a;
}

struct StepScope {   // marks "inside bcs_step" for the argument builders (deferred fold of the parked pair forces)
    bcs_sim* s;
    explicit StepScope(bcs_sim* sim) : s(sim) { s->inStep = true; }
    ~StepScope() { s->inStep = false; }
};

void enqueue_step(bcs_sim* s)
{
    if (s->slab && !s->slab->primed) slab_prime(s->slab, slab_ctx(s));
    StepScope scope(s);
    cudaStream_t m = s->stream;
    const bool fork = s->overlap && !s->ctx.timing && s->side[0];
    if (!fork) {
        // same stage order as the staged entry points; the tail (integrate particles, vein end, step counter) is one
        // fused kernel, and the vein integrator - independent of it - follows
        VeinCollideArgs va = vein_collide_args(s);
        if (va.wall.enabled && s->nearProbe) {
            // wall-grid path: the spring kernel carries the near-wall probe, so the structure is brought up to date first
            stage(s, BCS_STAGE_GRID_PARTICLES);
            stage(s, BCS_STAGE_GRID_TRIANGLES);
            stage(s, BCS_STAGE_VEIN_GATHER);
            launch_wall_reset(va, m);
            launch_wall_rebuild(va, s->hs.V, s->numSMs, m);
            launch_springs(spring_args(s, true), m);
            stage(s, BCS_STAGE_PARTICLE_COLLISIONS);
            va.wall.useNearList = 1;
            launch_wall_search(va, m);
            launch_wall_apply(va, m);
        } else {
            for (int st = BCS_STAGE_GRID_PARTICLES; st <= BCS_STAGE_VEIN_COLLISIONS; ++st) stage(s, st);
        }
        launch_finish_step(integrate_args(s), spring_args(s), s->doneBlocks, m);
        stage(s, BCS_STAGE_INTEGRATE_VEIN);
    } else {
        // Data flow of a step (reference order main.cu:175-208 + simulation_controller.cu:246-331, which serialises every
        // stage with a device-wide sync):
        //   grid build         reads pos, vel                    -> sorted copies, cell index
        //   springs            reads pos, vel, frc               -> frc, centres                (independent of the grid)
        //   vein gather        reads vpos, vvel                  -> vfrc                         (independent of particles)
        //   wall search        reads pos, vel, vpos, wall grid   -> near-hit list               (independent of frc)
        //   particle collisions need grid + springs; wall apply needs collisions + search + gather; the particle tail and
        //   the vein integrator both need wall apply and are independent of each other.
        VeinCollideArgs va = vein_collide_args(s);
        const bool wall = va.wall.enabled != 0;
        BCS_CUDA(cudaEventRecord(s->evFork, m));
        for (int k = 0; k < 3; ++k) BCS_CUDA(cudaStreamWaitEvent(s->side[k], s->evFork, 0));
        if (wall && s->nearProbe) {
            // the spring kernel carries the near-wall probe of the wall search (NearProbe, kernels.cuh): structure first,
            // then springs + probe, then the filter over the short near list, triangle tests and masking check
            launch_wall_reset(va, s->side[0]);
            launch_wall_rebuild(va, s->hs.V, s->numSMs, s->side[0]);
            launch_springs(spring_args(s, true), s->side[0]);
            BCS_CUDA(cudaEventRecord(s->evSprings, s->side[0]));
            va.wall.useNearList = 1;
            launch_wall_search(va, s->side[0]);
            BCS_CUDA(cudaEventRecord(s->evWall, s->side[0]));
        } else {
            launch_springs(spring_args(s), s->side[0]);
            BCS_CUDA(cudaEventRecord(s->evSprings, s->side[0]));
            if (wall) {
                launch_wall_rebuild(va, s->hs.V, s->numSMs, s->side[1]);
                launch_wall_search(va, s->side[1]);
            }
            BCS_CUDA(cudaEventRecord(s->evWall, s->side[1]));
        }
        launch_vein_gather(vein_args(s), s->side[2]);
        BCS_CUDA(cudaEventRecord(s->evGather, s->side[2]));
        stage(s, BCS_STAGE_GRID_PARTICLES);
        BCS_CUDA(cudaStreamWaitEvent(m, s->evSprings, 0));
        stage(s, BCS_STAGE_PARTICLE_COLLISIONS);
        BCS_CUDA(cudaStreamWaitEvent(m, s->evWall, 0));
        BCS_CUDA(cudaStreamWaitEvent(m, s->evGather, 0));
        if (wall) {
            launch_wall_apply(va, m);
        } else {
            launch_tri_refit(va, m);
            launch_vein_collisions(va, m);
        }
        BCS_CUDA(cudaEventRecord(s->evMasked, m));
        BCS_CUDA(cudaStreamWaitEvent(s->side[2], s->evMasked, 0));
        launch_vein_integrate(vein_args(s), s->side[2]);
        BCS_CUDA(cudaEventRecord(s->evVein, s->side[2]));
        launch_finish_step(integrate_args(s), spring_args(s), s->doneBlocks, m);
        BCS_CUDA(cudaStreamWaitEvent(m, s->evVein, 0));
    }
    if (s->slab) slab_end_of_step(s->slab, slab_ctx(s));   // migration + halo exchange for the next step
}

// ---- fused run (bcs_step(n) in row-directory mode) ---------------------------------------------------------------------
// The spring stage of step k + 1 reads exactly what the tail of step k writes (positions, velocities; forces are never
// zeroed), group by group of whole blood cells, so both run as ONE pass over the particle state (launch_advance,
// cellpass.cu) that also counts the new positions into the row directory of the next grid build.  A run of n steps is
//   head   springs + row count                                        (one launch)
//   body   row scan, scatter, order (+ near-wall probe) | pair search + apply | wall search + apply | vein
//   end    advance = integrate + vein end + springs + row count (steps 1 .. n-1) / finish_step (step n)
// with the same results, bit for bit, as n unfused steps (test_fused_run_equals_single_steps).
void enqueue_head(bcs_sim* s)
{
    if (s->slab) {
        // slab mode: springs over the owned blood cells, row count over owned + ghost particles (once per run; inside the
        // run the ghosts are counted as they are unpacked, slab.cu)
        launch_springs(spring_args(s), s->stream);
        launch_row_count(particle_grid_args(s), s->stream);
        return;
    }
    launch_springs_count(spring_args(s), s->pg, s->rows, s->counters, s->stream);
}

void enqueue_body(bcs_sim* s, bool last)
{
    StepScope scope(s);
    cudaStream_t m = s->stream;
    const bool fork = s->overlap && !s->ctx.timing && s->side[0];
    VeinCollideArgs va = vein_collide_args(s);
    const bool wall = va.wall.enabled != 0;
    const bool probeOn = wall && s->nearProbe;
    GridBuildArgs ga = particle_grid_args(s);
    ga.rows.countDone = 1;
    NearProbe probe{};
    if (probeOn) {
        const WallGridDev& w = s->wall;
        probe = NearProbe{w.near, w.nearList, w.nearCount, w.ox, w.oy, w.oz, w.invh, w.nx, w.ny, w.nz};
        ga.probe = &probe;
        va.wall.useNearList = 1;
    }
    cudaStream_t sWall = fork ? s->side[0] : m, sVein = fork ? s->side[2] : m;
    if (fork) {
        BCS_CUDA(cudaEventRecord(s->evFork, m));
        BCS_CUDA(cudaStreamWaitEvent(sWall, s->evFork, 0));
        BCS_CUDA(cudaStreamWaitEvent(sVein, s->evFork, 0));
    }
    // wall structure up to date before anything probes it
    if (wall) {
        launch_wall_reset(va, sWall);
        launch_wall_rebuild(va, s->hs.V, s->numSMs, sWall);
        if (fork) BCS_CUDA(cudaEventRecord(s->evRebuilt, sWall));
    }
    launch_vein_gather(vein_args(s), sVein);
    if (fork) BCS_CUDA(cudaEventRecord(s->evGather, sVein));
    if (fork && wall) BCS_CUDA(cudaStreamWaitEvent(m, s->evRebuilt, 0));
    launch_grid_build(ga, m);
    s->gridBuilt = true;
    if (wall) {
        if (fork) {
            BCS_CUDA(cudaEventRecord(s->evGrid, m));
            BCS_CUDA(cudaStreamWaitEvent(sWall, s->evGrid, 0));
        }
        launch_wall_search(va, sWall);
        if (fork) BCS_CUDA(cudaEventRecord(s->evWall, sWall));
    }
    launch_particle_collisions_rows(collide_args(s), m);
    if (fork) {
        if (wall) BCS_CUDA(cudaStreamWaitEvent(m, s->evWall, 0));
        BCS_CUDA(cudaStreamWaitEvent(m, s->evGather, 0));
    }
    if (wall) {
        launch_wall_apply(va, m);
    } else {
        launch_tri_refit(va, m);
        launch_vein_collisions(va, m);
    }
    if (fork) {
        BCS_CUDA(cudaEventRecord(s->evMasked, m));
        BCS_CUDA(cudaStreamWaitEvent(sVein, s->evMasked, 0));
    }
    launch_vein_integrate(vein_args(s), sVein);
    IntegrateArgs ia = integrate_args(s);
    if (s->slab) ia.tail = slab_tail(s->slab);   // ghost expiry + the particle part of the pack ride on the cell pass
    if (ia.tail.ghostList) slab_pack_vertices(s->slab, slab_ctx(s), sVein);
    if (fork) BCS_CUDA(cudaEventRecord(s->evVein, sVein));
    if (last) launch_finish_step(ia, spring_args(s), s->doneBlocks, m);
    else launch_advance(ia, spring_args(s), s->pg, s->rows, s->doneBlocks, m);
    if (fork) BCS_CUDA(cudaStreamWaitEvent(m, s->evVein, 0));
    if (s->slab) {
        // migration + halo exchange for the next step; inside a run the arrivals are counted into its row directory
        SlabCount cnt{};
        if (!last) { cnt.enabled = 1; cnt.grid = s->pg; cnt.rows = s->rows; cnt.counters = s->counters; }
        slab_end_of_step(s->slab, slab_ctx(s), ia.tail.ghostList != nullptr, &cnt);
    }
}

cudaGraphExec_t capture_body(bcs_sim* s, bool last, unsigned long long* kernels)
{
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    BCS_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    const unsigned long long before = s->ctx.launches;
    try {
        enqueue_body(s, last);
        *kernels = s->ctx.launches - before;
        s->ctx.launches = before;   // captured, not executed
    } catch (...) {
        cudaStreamEndCapture(s->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    BCS_CUDA(cudaStreamEndCapture(s->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    BCS_CUDA(e);
    return exec;
}

void run_fused(bcs_sim* s, int nsteps)
{
    if (nsteps <= 0) return;
    enqueue_head(s);
    if (!s->useGraph || s->ctx.timing) {
        for (int i = 0; i < nsteps; ++i) enqueue_body(s, i == nsteps - 1);
        return;
    }
    if (nsteps > 1 && !s->graphMid) s->graphMid = capture_body(s, false, &s->kernelsMid);
    if (!s->graphLast) s->graphLast = capture_body(s, true, &s->kernelsLast);
    for (int i = 0; i + 1 < nsteps; ++i) {
        BCS_CUDA(cudaGraphLaunch(s->graphMid, s->stream));
        s->ctx.launches += s->kernelsMid;
    }
    BCS_CUDA(cudaGraphLaunch(s->graphLast, s->stream));
    s->ctx.launches += s->kernelsLast;
}

struct Array {
    float4* ptr;
    int n;
    bool isParticlePos;
};
Array array_of(bcs_sim* s, int which)
{
    switch (which) {
    case BCS_PARTICLE_POS: return {s->pos, s->hs.N, true};
    case BCS_PARTICLE_VEL: return {s->vel, s->hs.N, false};
    case BCS_PARTICLE_FRC: return {s->frc, s->hs.N, false};
    case BCS_VEIN_POS: return {s->vpos, s->hs.V, false};
    case BCS_VEIN_VEL: return {s->vvel, s->hs.V, false};
    case BCS_VEIN_FRC: return {s->vfrc, s->hs.V, false};
    case BCS_CELL_CENTERS: return {s->centers, s->hs.B, false};
    }
    throw Error{BCS_ERR_INVALID, "unknown array id"};
}

void destroy(bcs_sim* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->phaseClock) {
        unsigned long long c[8] = {};
        cudaMemcpy(c, s->phaseClock, sizeof c, cudaMemcpyDeviceToHost);
        unsigned long long tot = 0;
        for (int k = 0; k < 7; ++k) tot += c[k];
        static const char* names[7] = {"setup+issue", "tile wait", "integrate", "centres", "springs", "gather+env", "count+writeback"};
        fprintf(stderr, "cell pass phases (share of summed warp cycles):");
        for (int k = 0; k < 7; ++k) fprintf(stderr, "  %s %.1f%%", names[k], tot ? 100.0 * (double)c[k] / (double)tot : 0.0);
        fprintf(stderr, "\n");
    }
    // graphs first: ncclCommDestroy (slab_destroy) waits for every graph that holds captured NCCL operations to be gone
    if (s->graphExec) cudaGraphExecDestroy(s->graphExec);
    if (s->graphMid) cudaGraphExecDestroy(s->graphMid);
    if (s->graphLast) cudaGraphExecDestroy(s->graphLast);
    slab_destroy(s->slab);
    for (void* p : s->owned) cudaFree(p);
    if (s->ownedIdxHost) cudaFreeHost(s->ownedIdxHost);
    if (s->hostStage) cudaFreeHost(s->hostStage);
    s->sortP.release();
    s->sortT.release();
    for (cudaStream_t q : s->side) if (q) cudaStreamDestroy(q);
    for (cudaEvent_t e : {s->evFork, s->evSprings, s->evWall, s->evGather, s->evMasked, s->evVein, s->evRebuilt, s->evGrid}) if (e) cudaEventDestroy(e);
    if (s->ownStream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

}  // namespace

#define BCS_API_BEGIN try {
#define BCS_API_END                                   \
    return BCS_OK;                                    \
    }                                                 \
    catch (const bcs::Error& e) {                     \
        bcs::set_error(e.msg);                        \
        return e.code;                                \
    }                                                 \
    catch (const std::exception& e) {                 \
        bcs::set_error(e.what());                     \
        return BCS_ERR_INVALID;                       \
    }

extern "C" {

const char* bcs_last_error(void) { return g_lastError.c_str(); }
int bcs_abi_version(void) { return BCS_ABI_VERSION; }

static int create_impl(const bcs_scene* scene, const bcs_opts* opts, const bcs_slab_opts* slabOpts, bcs_sim** out)
{
    bcs_sim* s = nullptr;
    try {
        BCS_REQUIRE(scene && out, BCS_ERR_INVALID, "null argument");
        BCS_REQUIRE(!opts || opts->struct_size == sizeof(bcs_opts), BCS_ERR_INVALID, "bcs_opts.struct_size mismatch");
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0)
            throw Error{BCS_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); libbcs has no CPU fallback"};
        s = new bcs_sim();
        s->device = opts ? opts->device : 0;
        BCS_REQUIRE(s->device >= 0 && s->device < ndev, BCS_ERR_INVALID, "device ordinal out of range");
        BCS_CUDA(cudaSetDevice(s->device));
        s->semantics = opts ? opts->semantics : BCS_SEM_CLEAN;
        BCS_REQUIRE(s->semantics == BCS_SEM_CLEAN || s->semantics == BCS_SEM_REFERENCE, BCS_ERR_INVALID, "unknown semantics");
        s->useGraph = opts ? opts->use_graph != 0 : true;
        s->stats = opts ? opts->collect_stats != 0 : false;
        s->seed = opts ? opts->seed : 0;
        s->exhaustiveVein = opts ? opts->exhaustive_vein_traversal != 0 : false;
        if (opts && opts->stream) s->stream = (cudaStream_t)opts->stream;
        else { BCS_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->ownStream = true; }

        derive_scene(*scene, s->hs);
        const HostScene& hs = s->hs;
        const int N = hs.N, B = hs.B, V = hs.V, T = hs.T;

        s->types.n = (int)hs.types.size();
        for (int i = 0; i < s->types.n; ++i) {
            const HostType& h = hs.types[i];
            s->types.t[i] = TypeDev{h.count, h.P, h.pStart, h.cStart, h.mStart, h.warpSync, hs.adjStart[i], hs.maxDeg[i], hs.sprStart[i], hs.nSpr[i]};
        }
        SpringTables springTables;
        s->plan = make_spring_plan(s->types, hs, springTables, slabOpts ? std::max(1, slabOpts->world) : 1);
        s->pg = make_grid(hs, hs.cellSize, hs.gdims, N);
        s->tg = make_grid(hs, hs.triCellSize, hs.tdims, T);
        fill_phys(hs, s->phys);

        s->pos = s->track(dev_alloc<float4>(N)); s->vel = s->track(dev_alloc<float4>(N)); s->frc = s->track(dev_alloc<float4>(N));
        s->spos = s->track(dev_alloc<float4>(N)); s->svel = s->track(dev_alloc<float4>(N));
        s->centers = s->track(dev_alloc<float4>(B));
        for (int k = 0; k < 2; ++k) {
            s->keys[k] = s->track(dev_alloc<int>((size_t)N + 4)); s->ids[k] = s->track(dev_alloc<int>(N));   // + key sentinels (pairs.cu)
            s->tkeys[k] = s->track(dev_alloc<int>(T)); s->tids[k] = s->track(dev_alloc<int>(T));
        }
        s->tcellStart = s->track(dev_alloc<int>(s->tg.cells)); s->tcellEnd = s->track(dev_alloc<int>(s->tg.cells));
        if (s->semantics == BCS_SEM_CLEAN) {
            // particle grid: compact cell index (1 bit + 1 rank word per 32 cells, starts per occupied cell)
            s->maskWords = s->pg.cells / 32 + 2;
            s->cellMask = s->track(dev_alloc<unsigned>(s->maskWords));
            s->cellRank = s->track(dev_alloc<int>(s->maskWords));
            s->occStart = s->track(dev_alloc<int>((size_t)N + 1));
            s->occKey = s->track(dev_alloc<int>(N));
            s->numOcc = s->track(dev_alloc<int>(1));
            // Row-directory grid + symmetric pair search: the default where rows are short (sparse scenes such as the long
            // vein: ~5 particles per occupied row); dense scenes (hundreds of particles per row) keep the compact cell index.
            // BCS_GRID=rows / cells / radix overrides.
            {
                const char* gm = getenv("BCS_GRID");
                // slab decomposition: only the rank's window of cell rows (slab + halo + slack) enters the row directory
                int rowY0 = 0, rowNyL = s->pg.ny;
                if (slabOpts && !getenv("BCS_ROWS_GLOBAL")) {
                    const float halo = (slabOpts->halo_width > 0 ? slabOpts->halo_width : 32.0f) + 16.0f;
                    const float lo = std::max(hs.gmin[1], slabOpts->y_lo - halo), hi = std::min(hs.gmax[1], slabOpts->y_hi + halo);
                    const int c0 = (int)std::floor((lo - hs.gmin[1]) / (float)hs.cellSize[1]) - 1, c1 = (int)std::floor((hi - hs.gmin[1]) / (float)hs.cellSize[1]) + 1;
                    rowY0 = std::max(0, std::min(c0, s->pg.ny - 1));
                    rowNyL = std::max(1, std::min(c1, s->pg.ny - 1) - rowY0 + 1);
                }
                const long long nRows = (long long)rowNyL * s->pg.nz;
                const bool sparse = (double)N / (slabOpts ? std::max(1, slabOpts->world) : 1) <= 4.0 * (double)nRows;
                const bool want = gm ? std::string(gm) == "rows" : sparse;
                if (want && nRows < (1ll << 30)) {
                    RowsGrid& R = s->rows;
                    R.enabled = 1;
                    R.nRows = (int)nRows;
                    R.rowCount = s->track(dev_alloc<unsigned>((size_t)nRows + 16));
                    // + 3: the scan writes from rowStart[1] on with 16-byte vector stores
                    R.rowStart = s->track(dev_alloc<int>((size_t)nRows + 24)) + 3;
                    R.keyOf = s->track(dev_alloc<int>(N));
                    R.tmp = s->track(dev_alloc<int2>(N));
                    R.irregular = s->track(dev_alloc<int>(2));
                    R.error = s->track(dev_alloc<int>(4));
                    R.nx = s->pg.nx; R.ny = s->pg.ny; R.nyL = rowNyL; R.y0 = rowY0;
                    R.local = (rowNyL != s->pg.ny) ? 1 : 0;
                    {
                        int l2 = 0;
                        while ((1ll << l2) < s->pg.ny) ++l2;
                        R.nyShift = 32 + l2;
                        R.nyMagic = ((1ull << R.nyShift) + (unsigned long long)s->pg.ny - 1ull) / (unsigned long long)s->pg.ny;
                    }
                    int l = 0;
                    while ((1ll << l) < s->pg.nx) ++l;
                    R.nxShift = 32 + l;
                    R.nxMagic = ((1ull << R.nxShift) + (unsigned long long)s->pg.nx - 1ull) / (unsigned long long)s->pg.nx;
                    PairLists& L = s->pairs;
                    L.cap = 4 * N + 4096;
                    L.pairs = s->track(dev_alloc<int2>((size_t)L.cap, false));
                    L.ctl = s->track(dev_alloc<int>(4));
                    L.acc = s->track(dev_alloc<long long>(3 * (size_t)N));
                    const char* cm = getenv("BCS_COLLIDE");
                    s->collideWalk = cm && std::string(cm) == "walk";
                }
            }
            // triangle grid: dense tables, empty cell = (start 0, end -1)
            BCS_CUDA(cudaMemset(s->tcellEnd, 0xFF, (size_t)s->tg.cells * sizeof(int)));
            BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        } else {
            // reference semantics: dense persistent tables, zero fill = (0,0)
            s->cellStart = s->track(dev_alloc<int>(s->pg.cells)); s->cellEnd = s->track(dev_alloc<int>(s->pg.cells));
        }
        s->vpos = s->track(dev_alloc<float4>(V)); s->vvel = s->track(dev_alloc<float4>(V)); s->vfrc = s->track(dev_alloc<float4>(V));
        s->vsplat = s->track(dev_alloc<long long>(3 * (size_t)V));
        s->tcent = s->track(dev_alloc<float4>(T));
        s->tris = s->track(dev_alloc<TriPacked>(T));
        s->groupBox = s->track(dev_alloc<Aabb>((T + 7) / 8));
        s->cellBox = s->track(dev_alloc<Aabb>(s->tg.cells));
        s->cullList = s->track(dev_alloc<CullEntry>((size_t)B + N));   // blood cells + (slab mode) ghost particles
        s->cellSlab = s->track(dev_alloc<CellSlab>(s->tg.cells));
        s->groupSlab = s->track(dev_alloc<CellSlab>((T + 7) / 8));
        s->cullCount = s->track(dev_alloc<int>(1));
        s->doneBlocks = s->track(dev_alloc<unsigned>(1));
        s->typesDev = s->track(dev_alloc<TypesDev>(1));
        BCS_CUDA(cudaMemcpy(s->typesDev, &s->types, sizeof(TypesDev), cudaMemcpyHostToDevice));
        BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        for (const HostType& h : hs.types) s->maxP = std::max(s->maxP, h.P);
        s->vidx = s->track(dev_upload(hs.vidx));
        s->nbrIds = s->track(dev_upload(hs.nbrIds)); s->nbrLen = s->track(dev_upload(hs.nbrLen));
        s->collR = s->track(dev_upload(hs.collR)); s->initR = s->track(dev_upload(hs.initR));
        s->mx = s->track(dev_upload(hs.mx)); s->my = s->track(dev_upload(hs.my)); s->mz = s->track(dev_upload(hs.mz));
        s->endC = s->track(dev_upload(hs.endC)); s->endR = s->track(dev_upload(hs.endR));
        s->adjJ = s->track(dev_upload(hs.adjJ)); s->adjL = s->track(dev_upload(hs.adjL));
        s->adjS = s->track(dev_upload(hs.adjS)); s->sprAB = s->track(dev_upload(hs.sprAB)); s->sprL = s->track(dev_upload(hs.sprL));
        s->slotTab = s->track(dev_upload(springTables.slot)); s->adjTab = s->track(dev_upload(springTables.adj));
        cell_pass_prepare(s->plan);   // function attributes are per device: set for THIS handle's device
        if (getenv("BCS_CP_CLOCK")) s->phaseClock = s->track(dev_alloc<unsigned long long>(8));
        s->counters = s->track(dev_alloc<Counters>(1));
        s->stagingLen = std::max(std::max(N, V), std::max(B, T));
        s->staging = s->track(dev_alloc<float>(3 * (size_t)s->stagingLen));
        s->sortP.allocate(N, std::max(s->pg.cells / 32 + 2, s->pg.ny * s->pg.nz + 2));
        s->sortT.allocate(T);

        // vein vertices -> device, triangle centres (calculateCentersKernel, run once), static triangle grid
        {
            float* sx = s->staging; float* sy = sx + s->stagingLen; float* sz = sy + s->stagingLen;
            BCS_CUDA(cudaMemcpy(sx, hs.vx.data(), V * sizeof(float), cudaMemcpyHostToDevice));
            BCS_CUDA(cudaMemcpy(sy, hs.vy.data(), V * sizeof(float), cudaMemcpyHostToDevice));
            BCS_CUDA(cudaMemcpy(sz, hs.vz.data(), V * sizeof(float), cudaMemcpyHostToDevice));
            // The copies come from pageable memory: cudaMemcpy returns when the source is staged, the last chunk may still be
            // on its way (legacy stream), and pack_kernel below runs on the handle's NON-BLOCKING stream.  Without this drain
            // the tail of the z array was occasionally packed as zeros (seen at 301 300 vertices: everything past the first
            // 1 MiB of floats) - a vein whose end lies flat in the z = 0 plane.
            BCS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
            pack_kernel<<<(V + 255) / 256, 256, 0, s->stream>>>(sx, sy, sz, s->vpos, V, s->types, s->collR, 0);
            launch_tri_centers(vein_args(s), s->tcent, s->stream);
            build_triangle_grid(s);
            BCS_CUDA(cudaStreamSynchronize(s->stream));
        }
        if (s->semantics == BCS_SEM_CLEAN && !getenv("BCS_NO_WALL_GRID")) setup_wall(s);
        s->overlap = !getenv("BCS_NO_OVERLAP");
        s->nearProbe = !getenv("BCS_NO_NEAR_PROBE");
        if (s->overlap) {
            for (cudaStream_t& q : s->side) BCS_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
            for (cudaEvent_t* e : {&s->evFork, &s->evSprings, &s->evWall, &s->evGather, &s->evMasked, &s->evVein, &s->evRebuilt, &s->evGrid})
                BCS_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        // fused run (bcs_step(n), n > 1): needs the row-directory grid (its count pass rides in the cell pass); single GPU
        s->fuseSteps = s->rows.enabled && !getenv("BCS_NO_FUSE") && !(slabOpts && getenv("BCS_SLAB_NO_FUSE"));
        if (slabOpts) {
            BCS_REQUIRE(slabOpts->struct_size == sizeof(bcs_slab_opts), BCS_ERR_INVALID, "bcs_slab_opts.struct_size mismatch");
            BCS_REQUIRE(s->semantics == BCS_SEM_CLEAN, BCS_ERR_UNSUPPORTED, "slab decomposition needs clean semantics");
            SlabInit in{};
            in.rank = slabOpts->rank; in.world = slabOpts->world; in.spawnRank = slabOpts->spawn_rank;
            in.yLo = slabOpts->y_lo; in.yHi = slabOpts->y_hi;
            in.haloWidth = slabOpts->halo_width > 0 ? slabOpts->halo_width : 32.0f;
            in.vertexHalo = slabOpts->vertex_halo > 0 ? slabOpts->vertex_halo : 2.0f * hs.triCellSize[1] + in.haloWidth + 20.0f;
            // defaults: 4x the expected number of particles in a halo layer / 2 % of the particles migrating at once
            const float extentY = hs.gsize[1] > 1.f ? hs.gsize[1] : 1.f;
            const int perHalo = (int)(4.0 * N * in.haloWidth / extentY) + 4096;
            in.capHalo = slabOpts->halo_capacity > 0 ? slabOpts->halo_capacity : std::min(N, perHalo);
            // migration: blood cells crossing one slab face (or teleporting to the spawn rank) in ONE step - a few cells;
            // default 2 % of a rank's share of the particles (+ slack).  Messages are sent at full capacity, so this is
            // also the per-step NVLink volume; overflow raises a sticky error
            in.capMig = slabOpts->migration_capacity > 0 ? slabOpts->migration_capacity : std::min(N, N / (50 * std::max(1, slabOpts->world)) + 2048);
            in.ncclId = slabOpts->nccl_unique_id;
            s->slab = slab_create(in, hs, s->tg, s->tids[1], s->tcellStart, s->tcellEnd, slab_ctx(s));
        }
        // set-up copies and fills ran on the legacy default stream (a pageable cudaMemcpy returns once the data is staged);
        // the handle's own stream is non-blocking and would not wait for them
        BCS_CUDA(cudaDeviceSynchronize());
        *out = s;
        return BCS_OK;
    } catch (const bcs::Error& e) {
        bcs::set_error(e.msg);
        destroy(s);
        return e.code;
    } catch (const std::exception& e) {
        bcs::set_error(e.what());
        destroy(s);
        return BCS_ERR_INVALID;
    }
}

int bcs_create(const bcs_scene* scene, const bcs_opts* opts, bcs_sim** out) { return create_impl(scene, opts, nullptr, out); }

int bcs_create_slab(const bcs_scene* scene, const bcs_opts* opts, const bcs_slab_opts* slab, bcs_sim** out)
{
    if (!slab) {
        bcs::set_error("null slab options");
        return BCS_ERR_INVALID;
    }
    return create_impl(scene, opts, slab, out);
}

int bcs_nccl_unique_id(char out[128])
{
    BCS_API_BEGIN
    BCS_REQUIRE(out, BCS_ERR_INVALID, "null argument");
    slab_unique_id(out);
    BCS_API_END
}

// ---- owned-only transfers (slab mode) ------------------------------------------------------------------------------------
static void check_device_flags(bcs_sim* s);

static void refresh_owned(bcs_sim* s)
{
    if (s->ownedHostValid) return;
    const int N = s->hs.N, B = s->hs.B;
    if (!s->slab->primed) slab_prime(s->slab, slab_ctx(s));
    if (!s->ownedIdxHost) {
        BCS_CUDA(cudaHostAlloc(&s->ownedIdxHost, (size_t)N * sizeof(int), cudaHostAllocDefault));
        BCS_CUDA(cudaHostAlloc(&s->hostStage, 3 * (size_t)N * sizeof(float), cudaHostAllocDefault));
        s->ownedIdxDev = s->track(dev_alloc<int>(N));
    }
    s->ownedHost.resize(B);
    BCS_CUDA(cudaMemcpyAsync(s->ownedHost.data(), s->slab->ownedCell, B, cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    s->ownedRuns.clear();
    int k = 0;
    for (int t = 0; t < s->types.n; ++t) {
        const TypeDev& ty = s->types.t[t];
        for (int c = 0; c < ty.count; ++c) {
            if (!s->ownedHost[ty.cStart + c]) continue;
            const int first = ty.pStart + c * ty.P;
            if (!s->ownedRuns.empty() && s->ownedRuns.back().x + s->ownedRuns.back().y == first) s->ownedRuns.back().y += ty.P;
            else s->ownedRuns.push_back(make_int2(first, ty.P));
            for (int j = 0; j < ty.P; ++j) s->ownedIdxHost[k++] = first + j;
        }
    }
    s->nOwnedParticles = k;
    BCS_CUDA(cudaMemcpyAsync(s->ownedIdxDev, s->ownedIdxHost, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    s->ownedHostValid = true;
}

// the ghosts of the neighbours mirror positions / velocities that an owned-only upload has just replaced: one exchange
// (no migration: ownership stays where it is until the next step decides) before anything reads them
static void refresh_halo_if_stale(bcs_sim* s)
{
    if (!s->slab || !s->haloStale) return;
    s->haloStale = false;
    if (!s->slab->primed) return;   // priming exchanges the halo anyway
    BCS_CUDA(cudaMemsetAsync(s->slab->moveTo, 0xFF, s->hs.B, s->stream));
    slab_end_of_step(s->slab, slab_ctx(s));
}

// device-visible aliases of three host arrays if all of them are pinned (cudaHostAlloc / cudaHostRegister) memory
static bool pinned_aliases(const float* x, const float* y, const float* z, const float* out[3])
{
    const float* in[3] = {x, y, z};
    for (int k = 0; k < 3; ++k) {
        cudaPointerAttributes at{};
        const bool ok = cudaPointerGetAttributes(&at, in[k]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer;
        cudaGetLastError();
        if (!ok) return false;
        out[k] = static_cast<const float*>(at.devicePointer);
    }
    return true;
}

static ActiveItems owned_items(bcs_sim* s)
{
    if (!s->slab->primed) slab_prime(s->slab, slab_ctx(s));
    ActiveItems it{};
    it.cells = s->slab->listCells; it.cellPrefix = s->slab->listCellPrefix;
    it.ghostList = s->slab->ghostList; it.ghostCount = s->slab->ghostCount;
    it.types = s->typesDev; it.maxP = s->maxP;
    return it;
}

int bcs_upload_owned(bcs_sim* s, int which, const float* x, const float* y, const float* z, int32_t n)
{
    if (s && !s->slab) return bcs_upload(s, which, x, y, z, n);
    BCS_API_BEGIN
    BCS_REQUIRE(s && x && y && z, BCS_ERR_INVALID, "null argument");
    BCS_REQUIRE(which == BCS_PARTICLE_POS || which == BCS_PARTICLE_VEL || which == BCS_PARTICLE_FRC, BCS_ERR_INVALID,
                "owned-only transfers move particle arrays (positions, velocities, forces)");
    BCS_REQUIRE(n == s->hs.N, BCS_ERR_INVALID, "array length mismatch");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    const float* dp[3];
    if (pinned_aliases(x, y, z, dp)) {
        // pinned arrays: as with bcs_upload the call is asynchronous - the arrays are read when the stream gets there
        Array a = array_of(s, which);
        const long long items = (long long)s->hs.B * s->maxP;
        upload_owned_direct_kernel<<<(int)std::max<long long>(1, std::min<long long>((items + 255) / 256, BOUNDED_BLOCKS)), 256, 0, s->stream>>>(
            dp[0], dp[1], dp[2], owned_items(s), a.ptr, s->collR, a.isParticlePos ? 1 : 0);
        BCS_CUDA(cudaGetLastError());
        if (which != BCS_PARTICLE_FRC) s->haloStale = true;
        return BCS_OK;
    }
    refresh_owned(s);
    const int m = s->nOwnedParticles;
    float* hx = s->hostStage; float* hy = hx + m; float* hz = hy + m;
    int k = 0;
    for (const int2& r : s->ownedRuns) {
        std::memcpy(hx + k, x + r.x, (size_t)r.y * sizeof(float));
        std::memcpy(hy + k, y + r.x, (size_t)r.y * sizeof(float));
        std::memcpy(hz + k, z + r.x, (size_t)r.y * sizeof(float));
        k += r.y;
    }
    if (m) {
        float* sx = s->staging;
        BCS_CUDA(cudaMemcpyAsync(sx, s->hostStage, 3 * (size_t)m * sizeof(float), cudaMemcpyHostToDevice, s->stream));   // one copy: x | y | z
        Array a = array_of(s, which);
        pack_owned_kernel<<<(m + 255) / 256, 256, 0, s->stream>>>(sx, sx + m, sx + 2 * (size_t)m, s->ownedIdxDev, m, a.ptr, s->types, s->collR,
                                                                  a.isParticlePos ? 1 : 0);
        BCS_CUDA(cudaGetLastError());
    }
    BCS_CUDA(cudaStreamSynchronize(s->stream));   // the pinned staging buffer is reused by the next call
    if (which != BCS_PARTICLE_FRC) s->haloStale = true;
    BCS_API_END
}

int bcs_download_owned(bcs_sim* s, int which, float* x, float* y, float* z, int32_t n)
{
    if (s && !s->slab) return bcs_download(s, which, x, y, z, n);
    BCS_API_BEGIN
    BCS_REQUIRE(s && x && y && z, BCS_ERR_INVALID, "null argument");
    BCS_REQUIRE(which == BCS_PARTICLE_POS || which == BCS_PARTICLE_VEL || which == BCS_PARTICLE_FRC, BCS_ERR_INVALID,
                "owned-only transfers move particle arrays (positions, velocities, forces)");
    BCS_REQUIRE(n == s->hs.N, BCS_ERR_INVALID, "array length mismatch");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    const float* dp[3];
    if (pinned_aliases(x, y, z, dp)) {
        Array a = array_of(s, which);
        const long long items = (long long)s->hs.B * s->maxP;
        download_owned_direct_kernel<<<(int)std::max<long long>(1, std::min<long long>((items + 255) / 256, BOUNDED_BLOCKS)), 256, 0, s->stream>>>(
            a.ptr, owned_items(s), const_cast<float*>(dp[0]), const_cast<float*>(dp[1]), const_cast<float*>(dp[2]));
        BCS_CUDA(cudaGetLastError());
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        check_device_flags(s);
        return BCS_OK;
    }
    refresh_owned(s);
    const int m = s->nOwnedParticles;
    if (m) {
        float* sx = s->staging;
        Array a = array_of(s, which);
        unpack_owned_kernel<<<(m + 255) / 256, 256, 0, s->stream>>>(a.ptr, s->ownedIdxDev, m, sx, sx + m, sx + 2 * (size_t)m);
        BCS_CUDA(cudaGetLastError());
        BCS_CUDA(cudaMemcpyAsync(s->hostStage, sx, 3 * (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    const float* hx = s->hostStage; const float* hy = hx + m; const float* hz = hy + m;
    int k = 0;
    for (const int2& r : s->ownedRuns) {
        std::memcpy(x + r.x, hx + k, (size_t)r.y * sizeof(float));
        std::memcpy(y + r.x, hy + k, (size_t)r.y * sizeof(float));
        std::memcpy(z + r.x, hz + k, (size_t)r.y * sizeof(float));
        k += r.y;
    }
    check_device_flags(s);
    BCS_API_END
}

int bcs_download_ownership(bcs_sim* s, uint8_t* owned, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && owned && n == s->hs.B, BCS_ERR_INVALID, "bad argument");
    BCS_CUDA(cudaSetDevice(s->device));
    if (!s->slab) {
        std::memset(owned, 1, n);
    } else {
        CtxScope scope(&s->ctx);
        if (!s->slab->primed) slab_prime(s->slab, slab_ctx(s));
        BCS_CUDA(cudaMemcpyAsync(owned, s->slab->ownedCell, n, cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
    }
    BCS_API_END
}

int bcs_slab_counts(bcs_sim* s, int32_t* active, int32_t* ghosts, int32_t* ownedCells)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && active && ghosts && ownedCells, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaSetDevice(s->device));
    if (!s->slab) { *active = s->hs.N; *ghosts = 0; *ownedCells = s->hs.B; return BCS_OK; }
    std::vector<unsigned char> o(s->hs.B);
    BCS_CUDA(cudaMemcpyAsync(active, s->slab->nActive, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(ghosts, s->slab->ghostCount, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(o.data(), s->slab->ownedCell, s->hs.B, cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    int k = 0;
    for (unsigned char v : o) k += v;
    *ownedCells = k;
    BCS_REQUIRE(!slab_check_error(s->slab, s->stream), BCS_ERR_STATE, "a halo / migration message overflowed its capacity (raise bcs_slab_opts capacities)");
    BCS_API_END
}

void bcs_destroy(bcs_sim* s) { destroy(s); }

int bcs_get_layout(const bcs_sim* s, bcs_layout* o)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && o, BCS_ERR_INVALID, "null argument");
    std::memset(o, 0, sizeof *o);
    const HostScene& hs = s->hs;
    o->n_types = (int)hs.types.size();
    o->n_particles = hs.N; o->n_cells = hs.B; o->n_model = hs.nModel; o->n_graph = hs.nGraph;
    o->n_vertices = hs.V; o->n_triangles = hs.T;
    for (int d = 0; d < 3; ++d) {
        o->grid_dims[d] = hs.gdims[d]; o->tri_grid_dims[d] = hs.tdims[d];
        o->grid_min[d] = hs.gmin[d]; o->grid_max[d] = hs.gmax[d]; o->grid_size[d] = hs.gsize[d];
    }
    o->grid_cells = s->pg.cells; o->tri_grid_cells = s->tg.cells;
    for (size_t i = 0; i < hs.types.size(); ++i) {
        const HostType& t = hs.types[i];
        o->types[i] = bcs_type_info{t.count, t.P, t.pStart, t.cStart, t.mStart, t.gStart, t.srcDef, t.warpSync, t.smallestRadius};
    }
    BCS_API_END
}

int bcs_get_table(bcs_sim* s, int table, void* dst, size_t bytes)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && dst, BCS_ERR_INVALID, "null argument");
    const HostScene& hs = s->hs;
    const void* src = nullptr;
    size_t n = 0;
    std::vector<float> tmp;
    switch (table) {
    case BCS_TABLE_SPRING_GRAPH: src = hs.graph.data(); n = hs.graph.size() * 4; break;
    case BCS_TABLE_MODEL_X: src = hs.mx.data(); n = hs.mx.size() * 4; break;
    case BCS_TABLE_MODEL_Y: src = hs.my.data(); n = hs.my.size() * 4; break;
    case BCS_TABLE_MODEL_Z: src = hs.mz.data(); n = hs.mz.size() * 4; break;
    case BCS_TABLE_COLLISION_RADII: src = hs.collR.data(); n = hs.collR.size() * 4; break;
    case BCS_TABLE_INITIAL_RADII: src = hs.initR.data(); n = hs.initR.size() * 4; break;
    case BCS_TABLE_VEIN_NBR_IDS: src = hs.nbrIds.data(); n = hs.nbrIds.size() * 4; break;
    case BCS_TABLE_VEIN_NBR_LEN: src = hs.nbrLen.data(); n = hs.nbrLen.size() * 4; break;
    case BCS_TABLE_TRI_CENTERS_X:
    case BCS_TABLE_TRI_CENTERS_Y:
    case BCS_TABLE_TRI_CENTERS_Z: {
        BCS_CUDA(cudaSetDevice(s->device));
        std::vector<float4> c(hs.T);
        BCS_CUDA(cudaMemcpyAsync(c.data(), s->tcent, hs.T * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        tmp.resize(hs.T);
        for (int i = 0; i < hs.T; ++i) tmp[i] = table == BCS_TABLE_TRI_CENTERS_X ? c[i].x : table == BCS_TABLE_TRI_CENTERS_Y ? c[i].y : c[i].z;
        src = tmp.data(); n = tmp.size() * 4;
        break;
    }
    default: throw Error{BCS_ERR_INVALID, "unknown table id"};
    }
    BCS_REQUIRE(bytes >= n, BCS_ERR_INVALID, "destination buffer too small");
    std::memcpy(dst, src, n);
    BCS_API_END
}

static void settle_grid_mode(bcs_sim* s, const float* y, const float* z, int n);

int bcs_upload(bcs_sim* s, int which, const float* x, const float* y, const float* z, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && x && y && z, BCS_ERR_INVALID, "null argument");
    BCS_REQUIRE(which != BCS_CELL_CENTERS, BCS_ERR_INVALID, "cell centres are derived, not uploadable");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    Array a = array_of(s, which);
    BCS_REQUIRE(n == a.n, BCS_ERR_INVALID, "array length mismatch");
    float* sx = s->staging; float* sy = sx + s->stagingLen; float* sz = sy + s->stagingLen;
    BCS_CUDA(cudaMemcpyAsync(sx, x, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    BCS_CUDA(cudaMemcpyAsync(sy, y, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    BCS_CUDA(cudaMemcpyAsync(sz, z, n * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    pack_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(sx, sy, sz, a.ptr, n, s->types, s->collR, a.isParticlePos ? 1 : 0);
    BCS_CUDA(cudaGetLastError());
    {
        // Pageable host memory: the caller may reuse its buffers as soon as this returns.  The runtime stages such copies
        // on most systems, but where the GPU can read pageable memory directly (HMM / ATS) the copy is genuinely
        // asynchronous - so the stream is drained unless the buffers are pinned (bcs_host_alloc / cudaHostAlloc).
        cudaPointerAttributes at{};
        const bool pinned = cudaPointerGetAttributes(&at, x) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!pinned) BCS_CUDA(cudaStreamSynchronize(s->stream));
    }
    if (s->slab && which == BCS_PARTICLE_POS) { s->slab->primed = false; s->ownedHostValid = false; }   // ownership is re-derived from the new positions
    if (which == BCS_PARTICLE_POS && !s->gridModeSettled) settle_grid_mode(s, y, z, n);
    if (s->wall.enabled && which == BCS_VEIN_POS) BCS_CUDA(cudaMemsetAsync(s->wall.dirty, 1, sizeof(int), s->stream));   // wall grid: rebuild
    if (which == BCS_VEIN_FRC) BCS_CUDA(cudaMemsetAsync(s->vsplat, 0, 3 * (size_t)s->hs.V * sizeof(long long), s->stream));   // the upload replaces parked splats too
    BCS_API_END
}

// The row-directory grid is chosen at creation from the grid's shape alone (particles per grid row).  What decides whether
// it beats the compact cell index is particles per OCCUPIED row - the order kernel ranks a particle against its row mates
// and the pair search scans runs of rows (measured, 100 k particles: 4.6 per occupied row 1.3x faster with rows, 9.6 per
// occupied row 1.2x slower) - and that is only known once positions are: the first position upload settles it.
static void settle_grid_mode(bcs_sim* s, const float* y, const float* z, int n)
{
    s->gridModeSettled = true;
    if (!s->rows.enabled || s->slab || getenv("BCS_GRID") || s->graphMid || s->graphLast || s->graphExec) return;
    const GridDev& g = s->pg;
    std::vector<unsigned long long> bits(((size_t)g.ny * g.nz + 63) / 64, 0ull);
    for (int i = 0; i < n; ++i) {
        const int cy = std::max(0, std::min(g.ny - 1, (int)std::floor((y[i] - g.miny) / (float)g.csy)));
        const int cz = std::max(0, std::min(g.nz - 1, (int)std::floor((z[i] - g.minz) / (float)g.csz)));
        const size_t row = (size_t)cz * g.ny + cy;
        bits[row >> 6] |= 1ull << (row & 63);
    }
    size_t occupied = 0;
    for (unsigned long long w : bits) occupied += (size_t)__builtin_popcountll(w);
    if ((double)n > 7.0 * (double)std::max<size_t>(1, occupied)) {
        s->rows.enabled = 0;     // dense rows: the compact cell index (always allocated) takes over
        s->fuseSteps = false;
    }
}

static void check_device_flags(bcs_sim* s);

int bcs_download(bcs_sim* s, int which, float* x, float* y, float* z, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && x && y && z, BCS_ERR_INVALID, "null argument");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    Array a = array_of(s, which);
    BCS_REQUIRE(n == a.n, BCS_ERR_INVALID, "array length mismatch");
    float* sx = s->staging; float* sy = sx + s->stagingLen; float* sz = sy + s->stagingLen;
    if (which == BCS_VEIN_FRC) launch_vein_fold_splats(vein_args(s), s->stream);   // wall splats of a half-finished step are parked in fixed point
    unpack_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(a.ptr, sx, sy, sz, n);
    BCS_CUDA(cudaGetLastError());
    BCS_CUDA(cudaMemcpyAsync(x, sx, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(y, sy, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(z, sz, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    if (s->slab) check_device_flags(s);   // a rank that dropped a halo / migration record must not hand out its state as good
    BCS_API_END
}

int bcs_device_ptrs(bcs_sim* s, bcs_device_view* o)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && o, BCS_ERR_INVALID, "null argument");
    o->particle_pos4 = s->pos; o->particle_vel4 = s->vel; o->particle_frc4 = s->frc;
    o->vein_pos4 = s->vpos; o->vein_vel4 = s->vvel; o->vein_frc4 = s->vfrc;
    o->stream = (void*)s->stream;
    BCS_API_END
}

int bcs_export_frame(bcs_sim* s, float* cellVertices6, float* offsets3, float* veinVertices6)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    BCS_CUDA(cudaSetDevice(s->device));
    const int N = s->hs.N, V = s->hs.V;
    if (cellVertices6 || offsets3) export_xyz_kernel<<<(N + 255) / 256, 256, 0, s->stream>>>(s->pos, N, cellVertices6, offsets3);
    if (veinVertices6) export_xyz_kernel<<<(V + 255) / 256, 256, 0, s->stream>>>(s->vpos, V, veinVertices6, nullptr);
    BCS_CUDA(cudaGetLastError());
    BCS_API_END
}

int bcs_host_alloc(void** out, size_t bytes)
{
    BCS_API_BEGIN
    BCS_REQUIRE(out, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    BCS_API_END
}
int bcs_host_free(void* p)
{
    BCS_API_BEGIN
    BCS_CUDA(cudaFreeHost(p));
    BCS_API_END
}

int bcs_run_stage(bcs_sim* s, int st)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    stage(s, st);
    BCS_API_END
}

int bcs_build_grid(bcs_sim* s)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    stage(s, BCS_STAGE_GRID_PARTICLES);
    stage(s, BCS_STAGE_GRID_TRIANGLES);
    BCS_API_END
}

int bcs_compute_forces(bcs_sim* s)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    for (int st = BCS_STAGE_VEIN_GATHER; st <= BCS_STAGE_VEIN_COLLISIONS; ++st) stage(s, st);
    BCS_API_END
}

int bcs_integrate(bcs_sim* s)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    for (int st = BCS_STAGE_INTEGRATE_PARTICLES; st <= BCS_STAGE_VEIN_END; ++st) stage(s, st);
    BCS_API_END
}

int bcs_step(bcs_sim* s, int32_t nsteps)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && nsteps >= 0, BCS_ERR_INVALID, "bad argument");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    if (s->slab && !s->slab->primed) slab_prime(s->slab, slab_ctx(s));   // ownership + first halo exchange, outside any capture
    refresh_halo_if_stale(s);
    if (nsteps > 0) s->ownedHostValid = false;   // blood cells may change owner
    if (s->fuseSteps) {
        run_fused(s, nsteps);
    } else if (!s->useGraph) {
        for (int i = 0; i < nsteps; ++i) enqueue_step(s);
    } else {
        if (!s->graphExec) {
            cudaGraph_t graph = nullptr;
            BCS_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            const unsigned long long before = s->ctx.launches;
            try {
                enqueue_step(s);
                s->kernelsPerGraph = s->ctx.launches - before;
                s->ctx.launches = before;   // captured, not executed
            } catch (...) {
                cudaStreamEndCapture(s->stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            BCS_CUDA(cudaStreamEndCapture(s->stream, &graph));
            cudaError_t e = cudaGraphInstantiate(&s->graphExec, graph, 0);
            cudaGraphDestroy(graph);
            BCS_CUDA(e);
        }
        for (int i = 0; i < nsteps; ++i) {
            BCS_CUDA(cudaGraphLaunch(s->graphExec, s->stream));
            s->ctx.launches += s->kernelsPerGraph;
        }
    }
    BCS_API_END
}

int bcs_get_launch_count(bcs_sim* s, uint64_t* kernels)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && kernels, BCS_ERR_INVALID, "null argument");
    *kernels = s->ctx.launches;
    BCS_API_END
}

int bcs_profile_steps(bcs_sim* s, int32_t nsteps, int32_t cap, char (*names)[BCS_KERNEL_NAME_LEN], float* ms_total, int32_t* launches,
                      int32_t* count)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && names && ms_total && launches && count && nsteps > 0 && cap > 0, BCS_ERR_INVALID, "bad argument");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    s->ctx.records.clear();
    if (s->slab && !s->slab->primed) slab_prime(s->slab, slab_ctx(s));
    refresh_halo_if_stale(s);
    s->ownedHostValid = false;
    s->ctx.timing = true;
    try {
        if (s->fuseSteps) run_fused(s, nsteps);
        else for (int i = 0; i < nsteps; ++i) enqueue_step(s);
    } catch (...) {
        s->ctx.timing = false;
        throw;
    }
    s->ctx.timing = false;
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    int k = 0;
    for (const LaunchRecord& r : s->ctx.records) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        int j = 0;
        for (; j < k; ++j)
            if (std::strncmp(names[j], r.name, BCS_KERNEL_NAME_LEN - 1) == 0) break;
        if (j == k) {
            if (k == cap) continue;
            std::strncpy(names[k], r.name, BCS_KERNEL_NAME_LEN - 1);
            names[k][BCS_KERNEL_NAME_LEN - 1] = 0;
            ms_total[k] = 0.f;
            launches[k] = 0;
            ++k;
        }
        ms_total[j] += ms;
        launches[j] += 1;
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    s->ctx.records.clear();
    *count = k;
    BCS_API_END
}

// sticky device-side error flags (halo / migration message overflow, a particle outside the rank's row window, wall-grid
// list overflow): a run that raised one has silently dropped or misplaced data, so every synchronising entry point reports it
static void check_device_flags(bcs_sim* s)
{
    if (s->slab) BCS_REQUIRE(!slab_check_error(s->slab, s->stream), BCS_ERR_STATE, "a halo / migration message overflowed its capacity (raise bcs_slab_opts capacities)");
    if (s->rows.enabled && s->rows.local) {
        int flag[4] = {0, 0, 0, 0};
        BCS_CUDA(cudaMemcpyAsync(flag, s->rows.error, sizeof flag, cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        if (flag[0]) {
            char msg[256];
            snprintf(msg, sizeof msg, "an active particle left the rank's window of grid rows (cell row %d of layer %d; the window is rows %d..%d): "
                     "raise bcs_slab_opts.halo_width", flag[1], flag[2], s->rows.y0, s->rows.y0 + s->rows.nyL - 1);
            throw Error{BCS_ERR_STATE, msg};
        }
    }
    if (s->wall.enabled) {
        int overflow = 0;
        BCS_CUDA(cudaMemcpyAsync(&overflow, s->wall.overflow, sizeof overflow, cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        BCS_REQUIRE(!overflow, BCS_ERR_STATE, "the wall grid outgrew its list capacity (vein deformed far beyond its rest shape)");
    }
}

int bcs_synchronize(bcs_sim* s)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s, BCS_ERR_INVALID, "null handle");
    BCS_CUDA(cudaSetDevice(s->device));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    check_device_flags(s);
    BCS_API_END
}

int bcs_get_step_count(const bcs_sim* s, int64_t* out)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && out, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaSetDevice(s->device));
    Counters c;
    BCS_CUDA(cudaMemcpyAsync(&c, s->counters, sizeof c, cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    *out = (int64_t)c.step;
    BCS_API_END
}

int bcs_set_step_count(bcs_sim* s, int64_t steps)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && steps >= 0, BCS_ERR_INVALID, "bad argument");
    BCS_CUDA(cudaSetDevice(s->device));
    const unsigned long long v = (unsigned long long)steps;
    BCS_CUDA(cudaMemcpyAsync(&s->counters->step, &v, sizeof v, cudaMemcpyHostToDevice, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    BCS_API_END
}

int bcs_get_stats(bcs_sim* s, bcs_stats* o)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && o, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaSetDevice(s->device));
    Counters c;
    BCS_CUDA(cudaMemcpyAsync(&c, s->counters, sizeof c, cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    o->pair_tests = c.pairTests; o->pair_hits = c.pairHits; o->triangle_tests = c.triTests;
    o->vein_hits = c.veinHits; o->teleported_cells = c.teleported; o->out_of_bounds = c.oob;
    o->wall_rebuilds = 0;
    if (s->wall.enabled) {
        unsigned long long builds = 0;
        BCS_CUDA(cudaMemcpy(&builds, s->wall.builds, sizeof builds, cudaMemcpyDeviceToHost));
        o->wall_rebuilds = builds;
    }
    check_device_flags(s);
    BCS_API_END
}

int bcs_download_grid(bcs_sim* s, int which, int32_t* keys, int32_t* ids, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && keys && ids, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaSetDevice(s->device));
    const int m = which ? s->hs.T : s->hs.N;
    BCS_REQUIRE(n == m, BCS_ERR_INVALID, "array length mismatch");
    BCS_REQUIRE(which || s->gridBuilt, BCS_ERR_STATE, "particle grid has not been built yet");
    BCS_CUDA(cudaMemcpyAsync(keys, which ? s->tkeys[1] : s->keys[1], n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(ids, which ? s->tids[1] : s->ids[1], n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    BCS_API_END
}

int bcs_download_cell_table(bcs_sim* s, int which, int32_t cap, int32_t* cells, int32_t* starts, int32_t* ends, int32_t* count)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && cells && starts && ends && count, BCS_ERR_INVALID, "null argument");
    BCS_CUDA(cudaSetDevice(s->device));
    if (!which && s->rows.enabled) {
        // row-directory mode keeps no per-cell table: the (debug) view is derived from the sorted keys
        BCS_REQUIRE(s->gridBuilt, BCS_ERR_STATE, "particle grid has not been built yet");
        int n = s->hs.N;
        if (s->slab) BCS_CUDA(cudaMemcpyAsync(&n, s->slab->nActive, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        std::vector<int> keys((size_t)std::max(n, 1));
        BCS_CUDA(cudaMemcpyAsync(keys.data(), s->keys[1], (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        int k = 0;
        for (int i = 0; i < n; ++i) {
            if (i == 0 || keys[i] != keys[i - 1]) {
                if (k < cap) { cells[k] = keys[i]; starts[k] = i; }
                ++k;
            }
            if (k - 1 < cap) ends[k - 1] = i;
        }
        *count = k;
        BCS_REQUIRE(k <= cap, BCS_ERR_INVALID, "capacity too small for the cell table");
        return BCS_OK;
    }
    if (!which && s->semantics == BCS_SEM_CLEAN) {
        // compact index: the occupied cells are stored explicitly
        BCS_REQUIRE(s->gridBuilt, BCS_ERR_STATE, "particle grid has not been built yet");
        int k = 0;
        BCS_CUDA(cudaMemcpyAsync(&k, s->numOcc, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        *count = k;
        BCS_REQUIRE(k <= cap, BCS_ERR_INVALID, "capacity too small for the cell table");
        std::vector<int> st_(k + 1);
        BCS_CUDA(cudaMemcpyAsync(cells, s->occKey, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaMemcpyAsync(st_.data(), s->occStart, (size_t)(k + 1) * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaStreamSynchronize(s->stream));
        for (int i = 0; i < k; ++i) { starts[i] = st_[i]; ends[i] = st_[i + 1] - 1; }
        return BCS_OK;
    }
    const int nc = which ? s->tg.cells : s->pg.cells;
    // debug path: copy the dense tables and compact on the host
    std::vector<int> hs_(nc), he_(nc);
    BCS_CUDA(cudaMemcpyAsync(hs_.data(), which ? s->tcellStart : s->cellStart, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaMemcpyAsync(he_.data(), which ? s->tcellEnd : s->cellEnd, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    BCS_CUDA(cudaStreamSynchronize(s->stream));
    const bool ref = s->semantics == BCS_SEM_REFERENCE;
    int k = 0;
    for (int c = 0; c < nc; ++c) {
        const bool keep = ref ? (hs_[c] != 0 || he_[c] != 0) : (he_[c] >= hs_[c]);
        if (!keep) continue;
        if (k < cap) { cells[k] = c; starts[k] = hs_[c]; ends[k] = he_[c]; }
        ++k;
    }
    *count = k;
    BCS_REQUIRE(k <= cap, BCS_ERR_INVALID, "capacity too small for the cell table");
    BCS_API_END
}

int bcs_debug_candidates(bcs_sim* s, int32_t* counts, uint64_t* sums, int32_t* hits, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && counts && sums && hits, BCS_ERR_INVALID, "null argument");
    BCS_REQUIRE(n == s->hs.N, BCS_ERR_INVALID, "array length mismatch");
    BCS_REQUIRE(s->gridBuilt, BCS_ERR_STATE, "particle grid has not been built yet");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    int* dc = dev_alloc<int>(n);
    int* dh = dev_alloc<int>(n);
    unsigned long long* ds = dev_alloc<unsigned long long>(n);
    CollideArgs a = collide_args(s);
    a.dbgCount = dc; a.dbgSum = ds; a.dbgHits = dh;
    cudaError_t e = cudaSuccess;
    try {
        if (s->rows.enabled) launch_particle_collisions_rows(a, s->stream);
        else launch_particle_collisions(a, s->stream);
        BCS_CUDA(cudaMemcpyAsync(counts, dc, n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaMemcpyAsync(hits, dh, n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaMemcpyAsync(sums, ds, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        e = cudaStreamSynchronize(s->stream);
    } catch (...) {
        cudaFree(dc); cudaFree(dh); cudaFree(ds);
        throw;
    }
    cudaFree(dc); cudaFree(dh); cudaFree(ds);
    BCS_CUDA(e);
    BCS_API_END
}

int bcs_debug_vein_hits(bcs_sim* s, int32_t* tri, float* t, int32_t n)
{
    BCS_API_BEGIN
    BCS_REQUIRE(s && tri && t, BCS_ERR_INVALID, "null argument");
    BCS_REQUIRE(n == s->hs.N, BCS_ERR_INVALID, "array length mismatch");
    CtxScope scope(&s->ctx);
    BCS_CUDA(cudaSetDevice(s->device));
    int* dt_ = dev_alloc<int>(n);
    float* df = dev_alloc<float>(n);
    VeinCollideArgs a = vein_collide_args(s);
    a.apply = false; a.dbgTri = dt_; a.dbgT = df;
    a.fast = false;   // the debug view is the reference's own exhaustive traversal
    cudaError_t e = cudaSuccess;
    try {
        launch_tri_refit(a, s->stream);
        launch_vein_collisions(a, s->stream);
        BCS_CUDA(cudaMemcpyAsync(tri, dt_, n * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        BCS_CUDA(cudaMemcpyAsync(t, df, n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        e = cudaStreamSynchronize(s->stream);
    } catch (...) {
        cudaFree(dt_); cudaFree(df);
        throw;
    }
    cudaFree(dt_); cudaFree(df);
    BCS_CUDA(e);
    BCS_API_END
}

}  // extern "C"

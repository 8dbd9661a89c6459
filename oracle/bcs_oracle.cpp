// ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (simulation-server_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and there only as the checker / the CPU baseline.
//
// A host-core (C++17 + OpenMP) restatement of the reference's per-step physics, stage by stage, written
// literally after the reference kernels so that it can serve as the correctness oracle for the CUDA path.
// Upstream has no CPU path; this is a port, not the reference ("kind": "port").
//
// PARITY PINNING: the reference's own tests hold no golden vector for this path.
// The oracle is pinned instead against dumps of the UNMODIFIED reference CUDA sources run headless on a
// B200 (oracle/ref_harness, oracle/build_ref.sh), committed under tests/golden/ together with the scripts
// that made them (tools/gpu_ref_goldens.sh, tools/make_goldens.py); tests/test_oracle_vs_reference.py
// checks every stage.  Independently of the dumps, tests/test_micro_scenes.py checks every stage against float64
// numpy restatements of the reference formulas on hand-built scenes (springs incl. the brake, collision force, grid
// boundary stencils, vein neighbour slots and springs, first-hit traversal, wall-hit effect, integrators, respawn).
//
// Reference lines followed (paths relative to the reference's src/):
//   layout / tables   meta_factory/blood_cell_factory.hpp:52-162,197-333; meta_factory/vein_factory.hpp:21-86,130-174
//   radii             simulation/simulation_controller.cu:93-153
//   math              utilities/math.cuh:11-104 (normalize: NaN -> 0)
//   cell id, grid     grids/uniform_grid.cu:20-80,129-155
//   vein springs      objects/vein_triangles.cu:14-27,88-154; simulation/physics.cuh:38-41
//   centres, springs  objects/blood_cells.cu:44-120; simulation/physics.cuh:24-27,53-78,102-120
//   collisions        simulation/particle_collisions.cuh:26-83,104-269; simulation/physics.cuh:133-145
//   vein collisions   simulation/vein_collisions.cuh:60-93; simulation/vein_collisions.cu:11-61,63-277
//   integration       objects/blood_cells.cu:155-179; objects/vein_triangles.cu:88-117
//   vein end          simulation/vein_end.cu:12-138
//   stage order       main.cu:175-208; simulation/simulation_controller.cu:246-331
//
// Floating point: plain IEEE single precision, one rounding per operation, evaluated in the reference's
// expression order (compile with -ffp-contract=off).  The device code of the reference (and of the
// product) may contract a*b+c into FMAs, so float results agree to ~1e-6 relative, integers exactly.  One
// expression is a decision threshold and therefore pinned to the FMA chain nvcc emits: d^2 of the touch test
// (stageParticleCollisions; the CUDA path spells out the same chain).
// Races of the reference are resolved as "snapshot" (springs read pre-stage forces) and
// "sequential sum in particle order" (vein force splats, Q9).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "bcs.h"

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------- math.cuh
struct f3 { float x, y, z; };
inline f3 mk(float x, float y, float z) { return {x, y, z}; }
inline f3 operator*(float a, f3 v) { return {a * v.x, a * v.y, a * v.z}; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator/(f3 v, float a) { return {v.x / a, v.y / a, v.z / a}; }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 u, f3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline float length_squared(f3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline float length(f3 v) { return std::sqrt(length_squared(v)); }
inline f3 normalize(f3 v)
{
    f3 vn = v / std::sqrt(dot(v, v));
    if (std::isnan(vn.x) || std::isnan(vn.y) || std::isnan(vn.z)) return {0, 0, 0};
    return vn;
}

struct V3 {   // SoA like cudaVec3
    std::vector<float> x, y, z;
    void resize(size_t n) { x.assign(n, 0.f); y.assign(n, 0.f); z.assign(n, 0.f); }
    f3 get(int i) const { return {x[i], y[i], z[i]}; }
    void set(int i, f3 v) { x[i] = v.x; y[i] = v.y; z[i] = v.z; }
    void add(int i, f3 v) { x[i] += v.x; y[i] += v.y; z[i] += v.z; }
    size_t size() const { return x.size(); }
};

// ---------------------------------------------------------------------------------------- Philox4x32-10
// Counter-based respawn RNG shared (by specification, not by code) with the product: key = seed,
// counter = (blood cell index, step, 0, 0); U = (x + 0.5) * 2^-32 like curand_uniform's (0,1].
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
inline float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

// ---------------------------------------------------------------------------------------- grid
struct Grid {
    int cs[3] = {1, 1, 1};
    int dims[3] = {0, 0, 0};
    int cells = 0;
    int n = 0;
    std::vector<int32_t> keys, ids, starts, ends;   // starts/ends: dense per-cell tables
    std::vector<int32_t> prevKeys;                  // clean mode: cells to un-write
};

struct Type {
    int count, P, pStart, cStart, mStart, gStart, srcDef, warpSync;
    float smallestRadius;
};

}  // namespace

struct orc_sim {
    int quirks = 0;
    bcs_physics ph{};
    int useBloodFlow = 1, reactionForce = 1, bigBrake = 1, bsCoeff = 3;
    uint64_t seed = 0;
    int64_t stepCount = 0;
    std::vector<Type> types;
    int N = 0, B = 0, nModel = 0, nGraph = 0, V = 0, T = 0;
    std::vector<float> graph, mx, my, mz, collR, initR;
    std::vector<int32_t> typeOfParticle;   // own type per particle id
    float gmin[3], gmax[3], gsize[3];
    V3 pos, vel, frc, centers;
    V3 vpos, vvel, vfrc, tcent;
    std::vector<uint32_t> vidx;
    std::vector<int32_t> nbrIds;   // [slot][vertex]
    std::vector<float> nbrLen;
    std::vector<float> endC, endR;
    Grid pg, tg;
    bool triGridBuilt = false;
    bcs_stats stats{};
    // debug outputs of the last collision stages
    std::vector<int32_t> dbgCount, dbgHits, dbgTri;
    std::vector<uint64_t> dbgSum;
    std::vector<float> dbgT;
};

namespace {

// ---------------------------------------------------------------------------------------- layout (meta_factory)
bool notPow2(int n) { return (n & (n - 1)) != 0; }            // the reference's `isPowerOfTwo` (:119-122)
bool orderBloodCells(int p1, int p2)                          // :130-147 incl. the dead `& 0` tests
{
    if (notPow2(p1) && !notPow2(p2)) return true;
    if (!notPow2(p1) && notPow2(p2)) return false;
    if ((p1 & 0) && (p2 & 1)) return true;
    if ((p1 & 1) && (p2 & 0)) return false;
    return true;
}
bool sameType(const bcs_cell_def& a, const bcs_cell_def& b)   // IsDuplicate :52-56
{
    if (a.particles_in_cell != b.particles_in_cell || a.n_springs != b.n_springs) return false;
    for (int i = 0; i < a.n_springs; ++i)
        if (a.springs[i].start != b.springs[i].start || a.springs[i].end != b.springs[i].end ||
            a.springs[i].length != b.springs[i].length)
            return false;
    return true;
}
std::vector<int> mpSort(const std::vector<int>& l, const bcs_cell_def* defs)   // boost::mp11::mp_sort (quicksort, first = pivot)
{
    if (l.size() <= 1) return l;
    int pivot = l[0];
    std::vector<int> a, b;
    for (size_t i = 1; i < l.size(); ++i)
        (orderBloodCells(defs[l[i]].particles_in_cell, defs[pivot].particles_in_cell) ? a : b).push_back(l[i]);
    std::vector<int> r = mpSort(a, defs);
    r.push_back(pivot);
    std::vector<int> s2 = mpSort(b, defs);
    r.insert(r.end(), s2.begin(), s2.end());
    return r;
}

void deriveLayout(orc_sim& s, const bcs_scene& sc)
{
    const bcs_cell_def* defs = sc.defs;
    std::vector<int> folded(sc.n_defs, 0), uniq;
    for (int i = 0; i < sc.n_defs; ++i)
        for (int j = 0; j < sc.n_defs; ++j)
            if (sameType(defs[i], defs[j])) folded[i] += defs[j].count;
    for (int i = 0; i < sc.n_defs; ++i) {
        bool dup = false;
        for (int j : uniq) dup = dup || sameType(defs[j], defs[i]);
        if (!dup) uniq.push_back(i);
    }
    std::vector<int> order = mpSort(uniq, defs);
    if ((int)order.size() > BCS_MAX_TYPES) throw std::runtime_error("too many blood cell types");
    int p = 0, c = 0, m = 0, g = 0;
    for (int i : order) {
        Type t{};
        t.count = folded[i]; t.P = defs[i].particles_in_cell;
        t.pStart = p; t.cStart = c; t.mStart = m; t.gStart = g; t.srcDef = i;
        // SelectSynchronizationType, vein_end.cu:23-30
        t.warpSync = (t.count * t.P <= 32 || (t.P > 0 && 32 % t.P == 0)) ? 1 : 0;
        p += t.count * t.P; c += t.count; m += t.P; g += t.P * t.P;
        s.types.push_back(t);
    }
    s.N = p; s.B = c; s.nModel = m; s.nGraph = g;
    s.graph.assign(g, 0.f);
    s.mx.resize(m); s.my.resize(m); s.mz.resize(m);
    for (auto& t : s.types) {
        const bcs_cell_def& d = defs[t.srcDef];
        for (int k = 0; k < d.n_springs; ++k) {          // springGraphGenerator :292-328
            const bcs_spring& sp = d.springs[k];
            if (sp.start < 0 || sp.end >= t.P || sp.end < 0 || sp.start >= t.P) throw std::runtime_error("ill-formed spring");
            s.graph[t.gStart + sp.start * t.P + sp.end] = sp.length * 1.0f;
            s.graph[t.gStart + sp.end * t.P + sp.start] = sp.length * 1.0f;
        }
        for (int j = 0; j < t.P; ++j) {
            s.mx[t.mStart + j] = d.vertices[3 * j];
            s.my[t.mStart + j] = d.vertices[3 * j + 1];
            s.mz[t.mStart + j] = d.vertices[3 * j + 2];
        }
    }
    s.typeOfParticle.resize(s.N);
    for (size_t ti = 0; ti < s.types.size(); ++ti)
        for (int i = 0; i < s.types[ti].count * s.types[ti].P; ++i) s.typeOfParticle[s.types[ti].pStart + i] = (int)ti;
}

// generateBoundingSpheres, simulation_controller.cu:93-153 (host code: double pow/sqrt, float result)
void deriveRadii(orc_sim& s)
{
    s.collR.assign(s.nModel, std::numeric_limits<float>::max());
    s.initR.assign(s.nModel, 0.f);
    for (auto& t : s.types) {
        t.smallestRadius = std::numeric_limits<float>::max();
        f3 center{0, 0, 0};
        for (int j = 0; j < t.P; ++j) {
            for (int k = 0; k < t.P; ++k) {
                if (j == k) continue;
                float dx = s.mx[t.mStart + j] - s.mx[t.mStart + k];
                float dy = s.my[t.mStart + j] - s.my[t.mStart + k];
                float dz = s.mz[t.mStart + j] - s.mz[t.mStart + k];
                float len = (float)(std::sqrt(std::pow((double)dx, 2) + std::pow((double)dy, 2) + std::pow((double)dz, 2)) /
                                    (2 * s.bsCoeff));
                if (len < s.collR[t.mStart + j]) s.collR[t.mStart + j] = len;
                if (len < t.smallestRadius) t.smallestRadius = len;
            }
            center = center + mk(s.mx[t.mStart + j], s.my[t.mStart + j], s.mz[t.mStart + j]);
        }
        center = center / (float)t.P;
        for (int j = 0; j < t.P; ++j)
            s.initR[t.mStart + j] = length(mk(s.mx[t.mStart + j], s.my[t.mStart + j], s.mz[t.mStart + j]) - center);
    }
}

// vein_factory.hpp:21-86 (bounds) and :130-174 (neighbour slots: sorted, duplicated, truncated to 9)
void deriveVein(orc_sim& s, const bcs_scene& sc)
{
    s.V = sc.n_vertices; s.T = sc.n_triangles;
    s.vpos.resize(s.V); s.vvel.resize(s.V); s.vfrc.resize(s.V);
    for (int i = 0; i < s.V; ++i) { s.vpos.x[i] = sc.vein_x[i]; s.vpos.y[i] = sc.vein_y[i]; s.vpos.z[i] = sc.vein_z[i]; }
    s.vidx.assign(sc.vein_indices, sc.vein_indices + 3 * (size_t)s.T);
    const float* c[3] = {sc.vein_x, sc.vein_y, sc.vein_z};
    for (int d = 0; d < 3; ++d) {
        float mn = *std::min_element(c[d], c[d] + s.V), mxv = *std::max_element(c[d], c[d] + s.V);
        float margin = d == 1 ? s.ph.grid_y_margin : s.ph.grid_xz_margin;
        s.gmin[d] = mn - margin;
        s.gmax[d] = mxv + margin;
        s.gsize[d] = s.gmax[d] - s.gmin[d];
    }
    std::vector<std::vector<uint32_t>> nb(s.V);
    for (int t = 0; t < s.T; ++t) {
        uint32_t i0 = s.vidx[3 * t], i1 = s.vidx[3 * t + 1], i2 = s.vidx[3 * t + 2];
        nb[i0].push_back(i1); nb[i0].push_back(i2);
        nb[i1].push_back(i0); nb[i1].push_back(i2);
        nb[i2].push_back(i0); nb[i2].push_back(i1);
    }
    s.nbrIds.assign((size_t)BCS_VEIN_MAX_NEIGHBORS * s.V, -1);
    s.nbrLen.assign((size_t)BCS_VEIN_MAX_NEIGHBORS * s.V, -1.0f);
    for (int i = 0; i < s.V; ++i) {
        std::sort(nb[i].begin(), nb[i].end());
        size_t m = std::min<size_t>(nb[i].size(), BCS_VEIN_MAX_NEIGHBORS);
        for (size_t j = 0; j < m; ++j) {
            uint32_t q = nb[i][j];
            s.nbrIds[j * s.V + i] = (int32_t)q;
            f3 d = s.vpos.get(i) - s.vpos.get((int)q);
            float l = length(d);
            s.nbrLen[j * s.V + i] = l < 0 ? -l : l;
        }
    }
    // calculateCentersKernel, vein_triangles.cu:14-27 (run once in the single-GPU constructor, :68)
    s.tcent.resize(s.T);
    for (int t = 0; t < s.T; ++t) {
        f3 a = s.vpos.get((int)s.vidx[3 * t]), b = s.vpos.get((int)s.vidx[3 * t + 1]), cc = s.vpos.get((int)s.vidx[3 * t + 2]);
        s.tcent.set(t, mk((a.x + b.x + cc.x) / 3, (a.y + b.y + cc.y) / 3, (a.z + b.z + cc.z) / 3));
    }
    s.endC.assign(sc.ending_centers, sc.ending_centers + 3 * (size_t)sc.n_endings);
    s.endR.assign(sc.ending_radii, sc.ending_radii + sc.n_endings);
}

void initGrid(const orc_sim& s, Grid& g, const int32_t cs[3], int n)
{
    for (int d = 0; d < 3; ++d) {
        g.cs[d] = cs[d];
        g.dims[d] = (int)(s.gsize[d] / (float)cs[d]);   // static_cast<int>(width / cellWidth), uniform_grid.cu:86-88
    }
    g.cells = g.dims[0] * g.dims[1] * g.dims[2];
    g.n = n;
    g.keys.assign(n, 0); g.ids.assign(n, 0);
    if (s.quirks) { g.starts.assign(g.cells, 0); g.ends.assign(g.cells, 0); }
    else { g.starts.assign(g.cells, 0); g.ends.assign(g.cells, -1); }
}

// calculateIdForCell, uniform_grid.cu:24-36, with the `max`/`min` macros of :20-21 written out.
inline int axisIndex(float p, float mn, float len, int cs)
{
    float q = (p - mn) / (float)cs;
    float m = (0 > q) ? 0 : q;          // max(0, q)   = ( a > b ? a : b )
    float r = (len > m) ? m : len;      // min(len, m) = ( a > b ? b : a )
    return (int)r;
}
inline int cellIdFor(const orc_sim& s, const Grid& g, float x, float y, float z, bool* oob)
{
    if (x < s.gmin[0] || x > s.gmax[0] || y < s.gmin[1] || y > s.gmax[1] || z < s.gmin[2] || z > s.gmax[2]) *oob = true;
    return axisIndex(z, s.gmin[2], s.gsize[2], g.cs[2]) * g.dims[0] * g.dims[1] +
           axisIndex(y, s.gmin[1], s.gsize[1], g.cs[1]) * g.dims[0] + axisIndex(x, s.gmin[0], s.gsize[0], g.cs[0]);
}

// UniformGrid::calculateGrid, uniform_grid.cu:129-155
void buildGrid(orc_sim& s, Grid& g, const V3& p)
{
    const int n = g.n;
    uint64_t oobCount = 0;
#pragma omp parallel for reduction(+ : oobCount) schedule(static)
    for (int i = 0; i < n; ++i) {
        bool oob = false;
        int id = cellIdFor(s, g, p.x[i], p.y[i], p.z[i], &oob);
        if (oob) ++oobCount;
        if (id < 0) id = 0;                     // memory-safety clamp for out-of-bounds positions only
        if (id >= g.cells) id = g.cells - 1;
        g.keys[i] = id;
        g.ids[i] = i;
    }
    s.stats.out_of_bounds += oobCount;
    // thrust::stable_sort_by_key(keys, ids): final order = lexicographic (cell id, object id)
    {
        std::vector<int32_t> k2(n), v2(n);
        int bits = 1;
        while ((1ll << bits) < g.cells) ++bits;
        for (int shift = 0; shift < bits; shift += 11) {
            std::vector<uint32_t> cnt(2049, 0);
            for (int i = 0; i < n; ++i) ++cnt[((g.keys[i] >> shift) & 2047) + 1];
            for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
            for (int i = 0; i < n; ++i) {
                uint32_t d = (g.keys[i] >> shift) & 2047;
                k2[cnt[d]] = g.keys[i]; v2[cnt[d]] = g.ids[i]; ++cnt[d];
            }
            g.keys.swap(k2); g.ids.swap(v2);
        }
    }
    // calculateStartAndEndOfCellKernel, uniform_grid.cu:51-80
    if (s.quirks) {
        // persistent, never cleared (Q1); thread N-1's stray `cellStarts[...] = N-1` lands last (Q2)
        for (int id = 0; id < n; ++id) {
            int c = g.keys[id];
            if (id > 0 && c != g.keys[id - 1]) g.starts[c] = id;
            if (id < n - 1 && c != g.keys[id + 1]) g.ends[c] = id;
        }
        if (n > 0) { g.starts[g.keys[0]] = 0; g.starts[g.keys[n - 1]] = n - 1; }
    } else {
        for (int32_t c : g.prevKeys) { g.starts[c] = 0; g.ends[c] = -1; }
        for (int id = 0; id < n; ++id) {
            int c = g.keys[id];
            if (id == 0 || c != g.keys[id - 1]) g.starts[c] = id;
            if (id == n - 1 || c != g.keys[id + 1]) g.ends[c] = id;
        }
        g.prevKeys = g.keys;
    }
}

// stencil trimming shared by particle_collisions.cuh:126-268 and vein_collisions.cu:86-230
inline void stencilRange(long long id, int count, int& lo, int& hi)
{
    if (id < 1) { lo = 0; hi = 1; }
    else if (id > (long long)count - 2) { lo = -1; hi = 0; }
    else { lo = -1; hi = 1; }
}

// ---------------------------------------------------------------------------------------- stages
void stageVeinGather(orc_sim& s)     // gatherForcesKernel, vein_triangles.cu:126-154
{
    const int V = s.V;
    std::vector<f3> add(V);
#pragma omp parallel for schedule(static)
    for (int id = 0; id < V; ++id) {
        f3 p = s.vpos.get(id), v = s.vvel.get(id), F{0, 0, 0};
        for (int slot = 0; slot < BCS_VEIN_MAX_NEIGHBORS; ++slot) {
            int nb = s.nbrIds[(size_t)slot * V + id];
            if (nb != -1) {
                float L = s.nbrLen[(size_t)slot * V + id];
                f3 q = s.vpos.get(nb);
                // springMassForceWithDampingForVein, physics.cuh:38-41
                float sf = (length(p - q) - L) * s.ph.vein_k_sniff + dot(normalize(p - q), (v - s.vvel.get(nb))) * s.ph.vein_d_fact;
                F = F + sf * normalize(q - p);
            }
        }
        add[id] = F;
    }
    for (int id = 0; id < V; ++id) s.vfrc.add(id, add[id]);
}

void stageSprings(orc_sim& s)        // calculateBloodCellsCenters + gatherForcesKernel, blood_cells.cu:44-120
{
    s.centers.resize(s.B);
    for (const Type& t : s.types) {
#pragma omp parallel for schedule(static)
        for (int c = 0; c < t.count; ++c) {
            int first = t.pStart + c * t.P;
            f3 center{0, 0, 0};
            for (int k = 0; k < t.P; ++k) center = center + s.pos.get(first + k);
            s.centers.set(t.cStart + c, center / (float)t.P);
        }
    }
    V3 out = s.frc;   // snapshot semantics: every thread reads pre-stage forces (Q7)
    const float dt = s.ph.dt;
    for (const Type& t : s.types) {
        const int nT = t.count * t.P;
#pragma omp parallel for schedule(static)
        for (int it = 0; it < nT; ++it) {
            int real = t.pStart + it, inCell = it % t.P, cell = t.cStart + it / t.P;
            f3 position = s.pos.get(real), velocity = s.vel.get(real), initialForce = s.frc.get(real), newForce{0, 0, 0};
            f3 radius = position - s.centers.get(cell);
            float initialRadius = s.initR[t.mStart + inCell];
            for (int j = 0; j < t.P; ++j) {
                float L = s.graph[t.gStart + j * t.P + inCell];
                if (L != 0.f) {
                    int nb = real - inCell + j;
                    f3 dP = position - s.pos.get(nb), dv = velocity - s.vel.get(nb);
                    f3 f2 = s.frc.get(nb);
                    // calculateParticlesSpringForceComponent (Heun branch), physics.cuh:53-78
                    f3 shift = normalize(-1.0f * dP);
                    f3 dv2 = dv + dt * (initialForce - f2);
                    float sm = (length(dP) - L) * s.ph.particle_k_sniff + dot(normalize(dP), dv2) * s.ph.particle_d_fact;
                    newForce = newForce + sm * shift;
                }
            }
            // accumulateEnvironmentForcesForParticles, physics.cuh:102-120
            float ratio = length(radius) / initialRadius;
            f3 G = mk(s.ph.gravity[0], s.ph.gravity[1], s.ph.gravity[2]);
            f3 env;
            if (s.bigBrake && ratio > s.ph.max_cell_size_factor_before_brake)
                env = G - (s.ph.viscous_damping * ratio * s.ph.big_particle_braking_intensity) * velocity;
            else
                env = G - s.ph.viscous_damping * velocity;
            newForce = newForce + env;
            out.set(real, (initialForce + newForce) / 2.0f);
        }
    }
    s.frc = std::move(out);
}

// physics::addResilientForceOnCollision, physics.cuh:133-145
inline f3 resilientForce(const orc_sim& s, f3 relPos, f3 relVel, float d2, float radius, float k)
{
    f3 dir = normalize(relPos);
    f3 tang = relVel - dot(relVel, dir) * dir;
    f3 spring = (-s.ph.collision_spring_coeff * (radius * 2 - std::sqrt(d2))) * dir;
    f3 damp = s.ph.collision_damping_coeff * relVel;
    f3 shear = s.ph.collision_shear_coeff * tang;
    return k * (spring + damp + shear);
}

inline int sliceTypeOfSlot(const orc_sim& s, int slot)
{
    for (size_t t = 0; t < s.types.size(); ++t)
        if (slot >= s.types[t].pStart && slot < s.types[t].pStart + s.types[t].count * s.types[t].P) return (int)t;
    return -1;
}

void stageParticleCollisions(orc_sim& s, bool apply)     // calculateParticleCollisions<UniformGrid>, particle_collisions.cuh:104-269
{
    const Grid& g = s.pg;
    const int N = s.N;
    s.dbgCount.assign(N, 0); s.dbgHits.assign(N, 0); s.dbgSum.assign(N, 0);
    uint64_t tests = 0, hits = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : tests, hits)
    for (int slot = 0; slot < N; ++slot) {
        const int pid = g.ids[slot];
        const f3 p1 = s.pos.get(pid), v1 = s.vel.get(pid);
        const int cell = g.keys[slot];
        const int xId = (int)((p1.x - s.gmin[0]) / (float)g.cs[0]);
        const int yId = (int)((p1.y - s.gmin[1]) / (float)g.cs[1]);
        const int zId = (int)((p1.z - s.gmin[2]) / (float)g.cs[2]);
        int x0, x1, y0, y1, z0, z1;
        stencilRange(xId, g.dims[0], x0, x1);
        stencilRange(yId, g.dims[1], y0, y1);
        stencilRange(zId, g.dims[2], z0, z1);
        // radius lookup: reference = the launching slice's (modelStart, particlesStart, P) for BOTH particles (Q4)
        const Type& own = s.types[s.typeOfParticle[pid]];
        const Type& sl = s.quirks ? s.types[sliceTypeOfSlot(s, slot)] : own;
        auto radiusOf = [&](int q) -> float {
            if (s.quirks) {
                int idx = sl.mStart + (q - sl.pStart) % sl.P;      // C++ remainder: sign of the dividend
                if (idx < 0) idx = 0;                              // the reference would read out of bounds here
                return s.collR[idx];
            }
            const Type& t = s.types[s.typeOfParticle[q]];
            return s.collR[t.mStart + (q - t.pStart) % t.P];
        };
        const float r1 = radiusOf(pid);
        f3 F = s.frc.get(pid);
        int cnt = 0, nh = 0;
        uint64_t sum = 0;
        for (int x = x0; x <= x1; ++x)
            for (int y = y0; y <= y1; ++y)
                for (int z = z0; z <= z1; ++z) {
                    long long nbr = (long long)cell + (long long)z * g.dims[0] * g.dims[1] + (long long)y * g.dims[0] + x;
                    if (nbr < 0 || nbr >= g.cells) continue;   // unreachable for in-bounds positions
                    for (int i = g.starts[nbr]; i <= g.ends[nbr]; ++i) {
                        int q = g.ids[i];
                        if (q == pid) continue;
                        ++cnt;
                        sum += (uint64_t)(q + 1) * 0x9E3779B97F4A7C15ull;
                        // detectCollision, particle_collisions.cuh:26-38
                        f3 rel = p1 - s.pos.get(q);
                        // length_squared as nvcc compiles it in the reference (-fmad=true): z*z + (y*y + (x*x)) as an FMA
                        // chain.  The touch decision is a threshold on this value, so the rounding sequence is part of
                        // the contract; the CUDA path pins the same chain.
                        float d2 = std::fmaf(rel.z, rel.z, std::fmaf(rel.y, rel.y, rel.x * rel.x));
                        float minD = r1 + radiusOf(q);
                        if (d2 <= minD * minD && d2 >= 0.0001f) {
                            ++nh;
                            f3 rv = v1 - s.vel.get(q);
                            f3 add = resilientForce(s, rel, rv, d2, r1, 0.5f);
                            F.x += add.x; F.y += add.y; F.z += add.z;
                        }
                    }
                }
        s.dbgCount[pid] = cnt; s.dbgHits[pid] = nh; s.dbgSum[pid] = sum;
        tests += cnt; hits += nh;
        if (apply) s.frc.set(pid, F);
    }
    if (apply) { s.stats.pair_tests += tests; s.stats.pair_hits += hits; }
}

struct Ray { f3 origin, direction, normal{0, 0, 0}; float t = 1e10f; int objectIndex = 0; };

// realCollisionDetection, vein_collisions.cu:11-45
inline bool rayTriangle(f3 v0, f3 v1, f3 v2, Ray& r, f3& refl)
{
    constexpr float EPS = 0.000001f;
    const f3 edge1 = v1 - v0, edge2 = v2 - v0;
    const f3 h = cross(r.direction, edge2);
    const float a = dot(edge1, h);
    if (a > -EPS && a < EPS) return false;
    const float f = 1 / a;
    const f3 sv = r.origin - v0;
    const float u = f * dot(sv, h);
    if (u < 0 || u > 1) return false;
    const f3 q = cross(sv, edge1);
    const float v = f * dot(r.direction, q);
    if (v < 0 || u + v > 1) return false;
    const float t = f * dot(edge2, q);
    if (t > EPS) {
        r.t = t;
        r.normal = normalize(cross(edge2, edge1));
        refl = r.direction - (2 * dot(r.direction, r.normal)) * r.normal;
        return true;
    }
    return false;
}

// calculateBaricentric, vein_collisions.cu:47-61
inline f3 barycentric(f3 point, f3 v0, f3 v1, f3 v2)
{
    f3 e0 = v1 - v0, e1 = v2 - v1, e2 = point - v0;
    float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(e2, e0), d21 = dot(e2, e1);
    float denom = d00 * d11 - d01 * d01;
    f3 b;
    b.x = (d11 * d20 - d01 * d21) / denom;
    b.y = (d00 * d21 - d01 * d20) / denom;
    b.z = 1.0f - b.x - b.y;
    return b;
}

inline long long toUnsignedIndex(float q)   // static_cast<unsigned int>(float): negatives saturate to 0 on the GPU (Q12)
{
    if (!(q > 0.f)) return 0;
    if (q >= 4294967296.f) return 4294967295ll;
    return (long long)(uint32_t)q;
}

void stageVeinCollisions(orc_sim& s, bool apply)     // detectVeinCollisions<UniformGrid>, vein_collisions.cu:63-277
{
    const Grid& g = s.tg;
    const int N = s.N;
    s.dbgTri.assign(N, -1); s.dbgT.assign(N, 1e10f);
    struct Hit { int pid, tri; f3 b, ds; };
    std::vector<std::vector<Hit>> perThread;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    perThread.resize(nthreads);
    uint64_t triTests = 0;
    const float impact2 = s.ph.vein_impact_distance * s.ph.vein_impact_distance;
    const float minForce2 = s.ph.vein_impact_minimal_force_distance * s.ph.vein_impact_minimal_force_distance;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : triTests)
    for (int pid = 0; pid < N; ++pid) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        const f3 velocity = s.vel.get(pid), pos = s.pos.get(pid);
        Ray r;
        r.origin = pos;
        r.direction = normalize(velocity);
        f3 refl{0, 0, 0};
        bool oob = false;
        int cell = cellIdFor(s, g, pos.x, pos.y, pos.z, &oob);
        long long xId = toUnsignedIndex((pos.x - s.gmin[0]) / (float)g.cs[0]);
        long long yId = toUnsignedIndex((pos.y - s.gmin[1]) / (float)g.cs[1]);
        long long zId = toUnsignedIndex((pos.z - s.gmin[2]) / (float)g.cs[2]);
        // `xId > cellCountX - 2` is an unsigned comparison in the reference
        auto range = [](long long id, int count, int& lo, int& hi) {
            long long lim = (long long)(uint32_t)(count - 2);
            if (id < 1) { lo = 0; hi = 1; }
            else if (id > lim) { lo = -1; hi = 0; }
            else { lo = -1; hi = 1; }
        };
        int x0, x1, y0, y1, z0, z1;
        range(xId, g.dims[0], x0, x1);
        range(yId, g.dims[1], y0, y1);
        range(zId, g.dims[2], z0, z1);
        bool hit = false;
        // calculateSideCollisions, vein_collisions.cuh:60-93: FIRST accepted triangle in traversal order (Q8)
        for (int x = x0; x <= x1 && !hit; ++x)
            for (int y = y0; y <= y1 && !hit; ++y)
                for (int z = z0; z <= z1 && !hit; ++z) {
                    long long nbr = (long long)cell + (long long)z * g.dims[0] * g.dims[1] + (long long)y * g.dims[0] + x;
                    if (nbr < 0 || nbr >= g.cells) continue;
                    for (int i = g.starts[nbr]; i <= g.ends[nbr]; ++i) {
                        int tri = g.ids[i];
                        ++triTests;
                        f3 v0 = s.vpos.get((int)s.vidx[3 * tri]), v1 = s.vpos.get((int)s.vidx[3 * tri + 1]),
                           v2 = s.vpos.get((int)s.vidx[3 * tri + 2]);
                        if (!rayTriangle(v0, v1, v2, r, refl)) continue;
                        r.objectIndex = tri;
                        hit = true;
                        break;
                    }
                }
        if (hit) { s.dbgTri[pid] = r.objectIndex; s.dbgT[pid] = r.t; }
        f3 relPos = pos - (pos + r.t * r.direction);
        float d2 = length_squared(relPos);
        if (apply && hit && d2 <= impact2) {
            if (d2 > minForce2) {
                if (s.reactionForce) {
                    f3 F = s.frc.get(pid);
                    f3 resp = ((-1.0f * dot(F, r.normal)) * r.normal) / dot(r.normal, r.normal);
                    s.frc.add(pid, resp);
                } else {
                    const Type& t = s.types[s.typeOfParticle[pid]];
                    f3 add = resilientForce(s, relPos, velocity, d2, s.collR[t.mStart + (pid - t.pStart) % t.P], 0.5f);
                    s.frc.add(pid, add);
                }
            }
            float speed = length(velocity);
            f3 dv = 1.0f * ((s.ph.velocity_collision_damping * speed) * refl - velocity);   // gpuCount = 1
            s.vel.add(pid, dv);
            f3 ds = s.ph.vein_collision_force_intensity * velocity;
            int tri = r.objectIndex;
            f3 v0 = s.vpos.get((int)s.vidx[3 * tri]), v1 = s.vpos.get((int)s.vidx[3 * tri + 1]), v2 = s.vpos.get((int)s.vidx[3 * tri + 2]);
            f3 b = barycentric(pos + r.t * r.direction, v0, v1, v2);
            perThread[tid].push_back({pid, tri, b, ds});
        }
    }
    if (apply) {
        // vein force splats (vein_collisions.cu:272-274), summed sequentially in particle order (Q9)
        std::vector<Hit> all;
        for (auto& v : perThread) all.insert(all.end(), v.begin(), v.end());
        std::sort(all.begin(), all.end(), [](const Hit& a, const Hit& b) { return a.pid < b.pid; });
        for (const Hit& h : all) {
            s.vfrc.add((int)s.vidx[3 * h.tri], h.b.x * h.ds);
            s.vfrc.add((int)s.vidx[3 * h.tri + 1], h.b.y * h.ds);
            s.vfrc.add((int)s.vidx[3 * h.tri + 2], h.b.z * h.ds);
        }
        s.stats.triangle_tests += triTests;
        s.stats.vein_hits += all.size();
    }
}

void stageIntegrateParticles(orc_sim& s)     // propagateParticleForcesKernel, blood_cells.cu:155-179 (Heun branch)
{
    const float dt = s.ph.dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; ++i) {
        f3 F = s.frc.get(i);
        f3 v0 = s.vel.get(i) / 1.0f;   // gpuCount = 1
        f3 v1 = v0 + dt * F;
        s.vel.set(i, v1);
        s.pos.add(i, (0.5f * dt) * (v1 + v0));
    }
}

void stageIntegrateVein(orc_sim& s)          // propagateForcesIntoPositionsKernel + memsets, vein_triangles.cu:88-117
{
    const float dt = s.ph.dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.V; ++i) {
        s.vvel.add(i, dt * s.vfrc.get(i));
        s.vpos.add(i, dt * s.vvel.get(i));
        s.vfrc.set(i, mk(0, 0, 0));
    }
}

void stageVeinEnd(orc_sim& s)                // HandleVeinEnd, vein_end.cu:12-173
{
    if (!s.useBloodFlow) return;
    const float upper = s.gmax[1] - 3 * s.ph.grid_y_margin / 4;
    const float lower = s.gmin[1] + s.ph.grid_y_margin / 2;
    const float right = s.gmax[0] - s.ph.grid_xz_margin / 2;
    const float left = s.gmin[0] + s.ph.grid_xz_margin / 2;
    const float front = s.gmax[2] - s.ph.grid_xz_margin / 2;
    const float back = s.gmin[2] + s.ph.grid_xz_margin / 2;
    uint64_t teleported = 0;
    for (const Type& t : s.types) {
#pragma omp parallel for schedule(static) reduction(+ : teleported)
        for (int c = 0; c < t.count; ++c) {
            bool tp = false;
            for (int k = 0; k < t.P; ++k) {
                int real = t.pStart + c * t.P + k;
                float px = s.pos.x[real], py = s.pos.y[real], pz = s.pos.z[real];
                bool one = false;
                if (!t.warpSync)   // handleVeinEndsWarpSync (:111-138) does not test the ending spheres
                    for (size_t e = 0; e < s.endR.size(); ++e)
                        one = one || length_squared(mk(px - s.endC[3 * e], py - s.endC[3 * e + 1], pz - s.endC[3 * e + 2])) <=
                                         s.endR[e] * s.endR[e];
                one = one || py <= lower || py >= upper || px <= left || px >= right || pz <= back || pz >= front;
                tp = tp || one;
            }
            if (!tp) continue;
            ++teleported;
            uint32_t ctr[4] = {(uint32_t)(t.cStart + c), (uint32_t)s.stepCount, (uint32_t)((uint64_t)s.stepCount >> 32), 0};
            philox4x32_10(ctr, (uint32_t)s.seed, (uint32_t)(s.seed >> 32));
            float u1 = u01(ctr[0]), u2 = u01(ctr[1]);
            for (int k = 0; k < t.P; ++k) {
                int real = t.pStart + c * t.P + k;
                s.pos.x[real] = (u1 - 0.5f) * 1.2f * s.ph.cylinder_radius + s.mx[t.mStart + k] - s.mx[t.mStart];
                s.pos.y[real] = s.ph.min_spawn_y + s.my[t.mStart + k] - s.my[t.mStart];
                s.pos.z[real] = (u2 - 0.5f) * 1.2f * s.ph.cylinder_radius + s.mz[t.mStart + k] - s.mz[t.mStart];
                s.vel.set(real, mk(s.ph.init_velocity[0], s.ph.init_velocity[1], s.ph.init_velocity[2]));
            }
        }
    }
    s.stats.teleported_cells += teleported;
}

void ensureTriGrid(orc_sim& s)
{
    // Single-GPU reference: centres are computed once, the triangle grid is re-sorted every frame from the
    // same centres (main.cu:176, Q14) - the result is identical every frame, so build it once.
    if (!s.triGridBuilt) { buildGrid(s, s.tg, s.tcent); s.triGridBuilt = true; }
}

int runStage(orc_sim& s, int stage)
{
    switch (stage) {
    case BCS_STAGE_GRID_PARTICLES: buildGrid(s, s.pg, s.pos); break;
    case BCS_STAGE_GRID_TRIANGLES: ensureTriGrid(s); break;
    case BCS_STAGE_VEIN_GATHER: stageVeinGather(s); break;
    case BCS_STAGE_SPRINGS: stageSprings(s); break;
    case BCS_STAGE_PARTICLE_COLLISIONS: stageParticleCollisions(s, true); break;
    case BCS_STAGE_VEIN_COLLISIONS: ensureTriGrid(s); stageVeinCollisions(s, true); break;
    case BCS_STAGE_INTEGRATE_PARTICLES: stageIntegrateParticles(s); break;
    case BCS_STAGE_INTEGRATE_VEIN: stageIntegrateVein(s); break;
    case BCS_STAGE_VEIN_END: stageVeinEnd(s); ++s.stepCount; break;
    default: g_err = "unknown stage"; return BCS_ERR_INVALID;
    }
    return BCS_OK;
}

V3* arrayOf(orc_sim& s, int which, int* n)
{
    switch (which) {
    case BCS_PARTICLE_POS: *n = s.N; return &s.pos;
    case BCS_PARTICLE_VEL: *n = s.N; return &s.vel;
    case BCS_PARTICLE_FRC: *n = s.N; return &s.frc;
    case BCS_VEIN_POS: *n = s.V; return &s.vpos;
    case BCS_VEIN_VEL: *n = s.V; return &s.vvel;
    case BCS_VEIN_FRC: *n = s.V; return &s.vfrc;
    case BCS_CELL_CENTERS: *n = s.B; return &s.centers;
    }
    return nullptr;
}

}  // namespace

extern "C" {

int orc_create(const bcs_scene* scene, const bcs_opts* opts, orc_sim** out)
{
    try {
        if (!scene || !out) { g_err = "null argument"; return BCS_ERR_INVALID; }
        auto* s = new orc_sim();
        s->quirks = opts ? (opts->semantics == BCS_SEM_REFERENCE) : 0;
        s->seed = opts ? opts->seed : 0;
        s->ph = scene->physics;
        s->useBloodFlow = scene->use_blood_flow; s->reactionForce = scene->enable_reaction_force;
        s->bigBrake = scene->enable_big_cells_brake; s->bsCoeff = scene->bounding_spheres_coeff;
        deriveLayout(*s, *scene);
        deriveRadii(*s);
        deriveVein(*s, *scene);
        s->pos.resize(s->N); s->vel.resize(s->N); s->frc.resize(s->N); s->centers.resize(s->B);
        initGrid(*s, s->pg, scene->cell_size, s->N);
        initGrid(*s, s->tg, scene->tri_cell_size, s->T);
        *out = s;
        return BCS_OK;
    } catch (const std::exception& e) {
        g_err = e.what();
        return BCS_ERR_INVALID;
    }
}
void orc_destroy(orc_sim* s) { delete s; }
const char* orc_last_error(void) { return g_err.c_str(); }
// launchers such as torchrun export OMP_NUM_THREADS=1 for every rank; the CPU baseline runs on ONE rank and takes the
// host cores it may use (bench.py passes the size of its affinity mask)
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_get_layout(const orc_sim* s, bcs_layout* o)
{
    std::memset(o, 0, sizeof *o);
    o->n_types = (int)s->types.size();
    o->n_particles = s->N; o->n_cells = s->B; o->n_model = s->nModel; o->n_graph = s->nGraph;
    o->n_vertices = s->V; o->n_triangles = s->T;
    for (int d = 0; d < 3; ++d) {
        o->grid_dims[d] = s->pg.dims[d]; o->tri_grid_dims[d] = s->tg.dims[d];
        o->grid_min[d] = s->gmin[d]; o->grid_max[d] = s->gmax[d]; o->grid_size[d] = s->gsize[d];
    }
    o->grid_cells = s->pg.cells; o->tri_grid_cells = s->tg.cells;
    for (size_t i = 0; i < s->types.size(); ++i) {
        const Type& t = s->types[i];
        o->types[i] = {t.count, t.P, t.pStart, t.cStart, t.mStart, t.gStart, t.srcDef, t.warpSync, t.smallestRadius};
    }
    return BCS_OK;
}

int orc_get_table(orc_sim* s, int table, void* dst, size_t bytes)
{
    const void* src = nullptr;
    size_t n = 0;
    switch (table) {
    case BCS_TABLE_SPRING_GRAPH: src = s->graph.data(); n = s->graph.size() * 4; break;
    case BCS_TABLE_MODEL_X: src = s->mx.data(); n = s->mx.size() * 4; break;
    case BCS_TABLE_MODEL_Y: src = s->my.data(); n = s->my.size() * 4; break;
    case BCS_TABLE_MODEL_Z: src = s->mz.data(); n = s->mz.size() * 4; break;
    case BCS_TABLE_COLLISION_RADII: src = s->collR.data(); n = s->collR.size() * 4; break;
    case BCS_TABLE_INITIAL_RADII: src = s->initR.data(); n = s->initR.size() * 4; break;
    case BCS_TABLE_VEIN_NBR_IDS: src = s->nbrIds.data(); n = s->nbrIds.size() * 4; break;
    case BCS_TABLE_VEIN_NBR_LEN: src = s->nbrLen.data(); n = s->nbrLen.size() * 4; break;
    case BCS_TABLE_TRI_CENTERS_X: src = s->tcent.x.data(); n = s->tcent.x.size() * 4; break;
    case BCS_TABLE_TRI_CENTERS_Y: src = s->tcent.y.data(); n = s->tcent.y.size() * 4; break;
    case BCS_TABLE_TRI_CENTERS_Z: src = s->tcent.z.data(); n = s->tcent.z.size() * 4; break;
    default: g_err = "unknown table"; return BCS_ERR_INVALID;
    }
    if (bytes < n) { g_err = "destination too small"; return BCS_ERR_INVALID; }
    std::memcpy(dst, src, n);
    return BCS_OK;
}

int orc_upload(orc_sim* s, int which, const float* x, const float* y, const float* z, int32_t n)
{
    int m = 0;
    V3* a = arrayOf(*s, which, &m);
    if (!a || which == BCS_CELL_CENTERS || n != m) { g_err = "bad array / length"; return BCS_ERR_INVALID; }
    std::copy(x, x + n, a->x.begin()); std::copy(y, y + n, a->y.begin()); std::copy(z, z + n, a->z.begin());
    return BCS_OK;
}
int orc_download(orc_sim* s, int which, float* x, float* y, float* z, int32_t n)
{
    int m = 0;
    V3* a = arrayOf(*s, which, &m);
    if (!a || n != m) { g_err = "bad array / length"; return BCS_ERR_INVALID; }
    std::copy(a->x.begin(), a->x.end(), x); std::copy(a->y.begin(), a->y.end(), y); std::copy(a->z.begin(), a->z.end(), z);
    return BCS_OK;
}

int orc_run_stage(orc_sim* s, int stage) { return runStage(*s, stage); }
int orc_build_grid(orc_sim* s) { buildGrid(*s, s->pg, s->pos); ensureTriGrid(*s); return BCS_OK; }
int orc_compute_forces(orc_sim* s)
{
    ensureTriGrid(*s);
    stageVeinGather(*s); stageSprings(*s); stageParticleCollisions(*s, true); stageVeinCollisions(*s, true);
    return BCS_OK;
}
int orc_integrate(orc_sim* s)
{
    stageIntegrateParticles(*s); stageIntegrateVein(*s); stageVeinEnd(*s); ++s->stepCount;
    return BCS_OK;
}
int orc_step(orc_sim* s, int32_t nsteps)
{
    for (int i = 0; i < nsteps; ++i) { orc_build_grid(s); orc_compute_forces(s); orc_integrate(s); }
    return BCS_OK;
}
int orc_synchronize(orc_sim*) { return BCS_OK; }
int orc_get_step_count(const orc_sim* s, int64_t* out) { *out = s->stepCount; return BCS_OK; }
int orc_set_step_count(orc_sim* s, int64_t steps) { s->stepCount = steps; return BCS_OK; }

int orc_download_grid(orc_sim* s, int which, int32_t* keys, int32_t* ids, int32_t n)
{
    if (which == 1) ensureTriGrid(*s);
    Grid& g = which ? s->tg : s->pg;
    if (n != g.n) { g_err = "bad length"; return BCS_ERR_INVALID; }
    std::copy(g.keys.begin(), g.keys.end(), keys); std::copy(g.ids.begin(), g.ids.end(), ids);
    return BCS_OK;
}
int orc_download_cell_table(orc_sim* s, int which, int32_t cap, int32_t* cells, int32_t* starts, int32_t* ends, int32_t* count)
{
    if (which == 1) ensureTriGrid(*s);
    Grid& g = which ? s->tg : s->pg;
    int k = 0;
    for (int c = 0; c < g.cells; ++c) {
        bool keep = s->quirks ? (g.starts[c] != 0 || g.ends[c] != 0) : (g.ends[c] >= g.starts[c]);
        if (!keep) continue;
        if (k < cap) { cells[k] = c; starts[k] = g.starts[c]; ends[k] = g.ends[c]; }
        ++k;
    }
    *count = k;
    if (k > cap) { g_err = "capacity too small"; return BCS_ERR_INVALID; }
    return BCS_OK;
}
int orc_debug_candidates(orc_sim* s, int32_t* counts, uint64_t* sums, int32_t* hits, int32_t n)
{
    if (n != s->N) { g_err = "bad length"; return BCS_ERR_INVALID; }
    stageParticleCollisions(*s, false);
    std::copy(s->dbgCount.begin(), s->dbgCount.end(), counts);
    std::copy(s->dbgSum.begin(), s->dbgSum.end(), sums);
    std::copy(s->dbgHits.begin(), s->dbgHits.end(), hits);
    return BCS_OK;
}
int orc_debug_vein_hits(orc_sim* s, int32_t* tri, float* t, int32_t n)
{
    if (n != s->N) { g_err = "bad length"; return BCS_ERR_INVALID; }
    ensureTriGrid(*s);
    stageVeinCollisions(*s, false);
    std::copy(s->dbgTri.begin(), s->dbgTri.end(), tri);
    std::copy(s->dbgT.begin(), s->dbgT.end(), t);
    return BCS_OK;
}
int orc_get_stats(orc_sim* s, bcs_stats* out) { *out = s->stats; return BCS_OK; }

}  // extern "C"

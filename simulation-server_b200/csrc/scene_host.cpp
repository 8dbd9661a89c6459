// Host-side scene derivation: the run-time equivalent of the reference's compile-time meta_factory
// (blood_cell_factory.hpp:52-333, vein_factory.hpp:21-174) and of the host setup in
// SimulationController::generateBoundingSpheres (simulation_controller.cu:93-153).
//
// Compiled WITHOUT floating-point contraction: the reference computes these tables on the host
// (constexpr / plain x86-64 code), one rounding per operation.
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

#include "bcs_internal.cuh"

namespace bcs {
namespace {

// ---- blood cell type list -----------------------------------------------------------------------------
// IsDuplicate (blood_cell_factory.hpp:52-56): equal particlesInCell and the same spring list *type*; at
// run time "same type" is "same sequence of (start, end, length)".
bool equivalent(const bcs_cell_def& a, const bcs_cell_def& b)
{
    if (a.particles_in_cell != b.particles_in_cell || a.n_springs != b.n_springs) return false;
    return std::equal(a.springs, a.springs + a.n_springs, b.springs, [](const bcs_spring& s, const bcs_spring& t) {
        return s.start == t.start && s.end == t.end && s.length == t.length;
    });
}

// orderBloodCells (blood_cell_factory.hpp:130-147).  The reference's helper named isPowerOfTwo returns
// n & (n-1), i.e. it is true for NON powers of two, and both "even" tests are `& 0` (always false);
// what remains is: comp(a, b) is false exactly when a is a power of two (or 0) and b is not.
bool comes_first(int pa, int pb)
{
    const bool aNonPow2 = (pa & (pa - 1)) != 0;
    const bool bNonPow2 = (pb & (pb - 1)) != 0;
    return !(!aNonPow2 && bNonPow2);
}

// boost::mp11::mp_sort is a quicksort whose pivot is the first element and whose partition keeps the
// input order (mp11/algorithm.hpp:609-619).  The comparator above is not a strict weak order, so the
// result depends on exactly this algorithm; reproduce it with an explicit work list.
std::vector<int> mp_sort_order(std::vector<int> items, const bcs_cell_def* defs)
{
    if (items.size() < 2) return items;
    const int pivot = items.front();
    std::vector<int> before, after;
    for (size_t k = 1; k < items.size(); ++k) {
        const int u = items[k];
        if (comes_first(defs[u].particles_in_cell, defs[pivot].particles_in_cell)) before.push_back(u);
        else after.push_back(u);
    }
    std::vector<int> out = mp_sort_order(std::move(before), defs);
    out.push_back(pivot);
    for (int u : mp_sort_order(std::move(after), defs)) out.push_back(u);
    return out;
}

void build_types(const bcs_scene& in, HostScene& hs)
{
    BCS_REQUIRE(in.n_defs > 0 && in.defs, BCS_ERR_INVALID, "scene has no blood cell definitions");
    const int n = in.n_defs;
    for (int i = 0; i < n; ++i) {
        const bcs_cell_def& d = in.defs[i];
        BCS_REQUIRE(d.count > 0 && d.particles_in_cell > 0, BCS_ERR_INVALID, "blood cell definition with non-positive count/size");
        BCS_REQUIRE(d.vertices != nullptr, BCS_ERR_INVALID, "blood cell definition without vertices");
        BCS_REQUIRE(d.n_springs == 0 || d.springs != nullptr, BCS_ERR_INVALID, "blood cell definition without springs");
        BCS_REQUIRE(d.particles_in_cell <= 256, BCS_ERR_UNSUPPORTED, "more than 256 particles per blood cell is not supported");
    }
    // class representative = first definition of each equivalence class; folded count = class total
    std::vector<int> rep(n);
    for (int i = 0; i < n; ++i) {
        rep[i] = i;
        for (int j = 0; j < i; ++j)
            if (equivalent(in.defs[j], in.defs[i])) { rep[i] = rep[j]; break; }
    }
    std::vector<int> total(n, 0), reps;
    for (int i = 0; i < n; ++i) total[rep[i]] += in.defs[i].count;
    for (int i = 0; i < n; ++i)
        if (rep[i] == i) reps.push_back(i);
    const std::vector<int> order = mp_sort_order(reps, in.defs);
    BCS_REQUIRE((int)order.size() <= BCS_MAX_TYPES, BCS_ERR_UNSUPPORTED, "more than 16 distinct blood cell types (maxCudaStreams)");

    int pAcc = 0, cAcc = 0, mAcc = 0, gAcc = 0;
    for (int src : order) {
        HostType t{};
        t.count = total[src];
        t.P = in.defs[src].particles_in_cell;
        t.pStart = pAcc; t.cStart = cAcc; t.mStart = mAcc; t.gStart = gAcc;
        t.srcDef = src;
        // SelectSynchronizationType (vein_end.cu:23-30)
        t.warpSync = (t.count * t.P <= 32 || 32 % t.P == 0) ? 1 : 0;
        pAcc += t.count * t.P; cAcc += t.count; mAcc += t.P; gAcc += t.P * t.P;
        hs.types.push_back(t);
    }
    hs.N = pAcc; hs.B = cAcc; hs.nModel = mAcc; hs.nGraph = gAcc;

    hs.graph.assign(gAcc, 0.0f);
    hs.mx.resize(mAcc); hs.my.resize(mAcc); hs.mz.resize(mAcc);
    for (const HostType& t : hs.types) {
        const bcs_cell_def& d = in.defs[t.srcDef];
        float* g = hs.graph.data() + t.gStart;
        for (int k = 0; k < d.n_springs; ++k) {     // springGraphGenerator (:292-328): symmetric fill, last entry wins
            const int a = d.springs[k].start, b = d.springs[k].end;
            BCS_REQUIRE(a >= 0 && a < t.P && b >= 0 && b < t.P, BCS_ERR_INVALID, "ill-formed spring definition");
            g[a * t.P + b] = d.springs[k].length;
            g[b * t.P + a] = d.springs[k].length;
        }
        for (int j = 0; j < t.P; ++j) {
            hs.mx[t.mStart + j] = d.vertices[3 * j + 0];
            hs.my[t.mStart + j] = d.vertices[3 * j + 1];
            hs.mz[t.mStart + j] = d.vertices[3 * j + 2];
        }
    }

    // ELL adjacency for the spring kernel: for particle i of a cell, its mates j in ascending order with
    // graph[j*P + i] != 0 (the order blood_cells.cu:88-110 visits them in)
    for (const HostType& t : hs.types) {
        const float* g = hs.graph.data() + t.gStart;
        int deg = 0;
        for (int i = 0; i < t.P; ++i) {
            int d = 0;
            for (int j = 0; j < t.P; ++j) d += g[j * t.P + i] != 0.0f;
            deg = std::max(deg, d);
        }
        hs.adjStart.push_back((int)hs.adjJ.size());
        hs.maxDeg.push_back(deg);
        const size_t base = hs.adjJ.size();
        hs.adjJ.resize(base + (size_t)deg * t.P, -1);
        hs.adjL.resize(base + (size_t)deg * t.P, 0.0f);
        hs.adjS.resize(base + (size_t)deg * t.P, -1);
        // undirected springs (a < b), numbered in (a, b) order
        hs.sprStart.push_back((int)hs.sprAB.size());
        std::vector<int> springOf((size_t)t.P * t.P, -1);
        int ns = 0;
        for (int a = 0; a < t.P; ++a)
            for (int b = a + 1; b < t.P; ++b)
                if (g[a * t.P + b] != 0.0f) {
                    springOf[a * t.P + b] = springOf[b * t.P + a] = ns++;
                    hs.sprAB.push_back(a | (b << 16));
                    hs.sprL.push_back(g[a * t.P + b]);
                }
        hs.nSpr.push_back(ns);
        for (int i = 0; i < t.P; ++i) {
            int d = 0;
            for (int j = 0; j < t.P; ++j) {
                const float L = g[j * t.P + i];
                if (L != 0.0f) {
                    hs.adjJ[base + (size_t)d * t.P + i] = j;
                    hs.adjL[base + (size_t)d * t.P + i] = L;
                    // a spring from a particle to itself (j == i) has no undirected entry; it contributes nothing
                    // (zero-length direction -> normalize gives 0), like in the reference
                    hs.adjS[base + (size_t)d * t.P + i] = (i == j) ? -1 : (springOf[i * t.P + j] | (i > j ? (int)0x80000000 : 0));
                    ++d;
                }
            }
        }
    }
}

// ---- radii ---------------------------------------------------------------------------------------------
void build_radii(HostScene& hs)
{
    hs.collR.assign(hs.nModel, 0.0f);
    hs.initR.assign(hs.nModel, 0.0f);
    const double denom = 2 * hs.bsCoeff;
    for (HostType& t : hs.types) {
        const float* x = hs.mx.data() + t.mStart;
        const float* y = hs.my.data() + t.mStart;
        const float* z = hs.mz.data() + t.mStart;
        float typeMin = std::numeric_limits<float>::max();
        float sx = 0.0f, sy = 0.0f, sz = 0.0f;
        for (int j = 0; j < t.P; ++j) {
            float best = std::numeric_limits<float>::max();
            for (int k = 0; k < t.P; ++k) {
                if (k == j) continue;
                // float differences, squared/summed/rooted in double, divided by 2*coeff, narrowed to float
                const double dx = (double)(x[j] - x[k]), dy = (double)(y[j] - y[k]), dz = (double)(z[j] - z[k]);
                const float half = (float)(std::sqrt(dx * dx + dy * dy + dz * dz) / denom);
                best = std::min(best, half);
            }
            hs.collR[t.mStart + j] = best;
            typeMin = std::min(typeMin, best);
            sx = sx + x[j]; sy = sy + y[j]; sz = sz + z[j];
        }
        t.smallestRadius = typeMin;
        const float cx = sx / (float)t.P, cy = sy / (float)t.P, cz = sz / (float)t.P;
        for (int j = 0; j < t.P; ++j) {
            const float ax = x[j] - cx, ay = y[j] - cy, az = z[j] - cz;
            hs.initR[t.mStart + j] = std::sqrt(ax * ax + ay * ay + az * az);
        }
    }
}

// ---- vein ----------------------------------------------------------------------------------------------
void build_vein(const bcs_scene& in, HostScene& hs)
{
    BCS_REQUIRE(in.n_vertices > 0 && in.vein_x && in.vein_y && in.vein_z, BCS_ERR_INVALID, "scene has no vein vertices");
    BCS_REQUIRE(in.n_triangles > 0 && in.vein_indices, BCS_ERR_INVALID, "scene has no vein triangles");
    hs.V = in.n_vertices;
    hs.T = in.n_triangles;
    hs.vx.assign(in.vein_x, in.vein_x + hs.V);
    hs.vy.assign(in.vein_y, in.vein_y + hs.V);
    hs.vz.assign(in.vein_z, in.vein_z + hs.V);
    hs.vidx.assign(in.vein_indices, in.vein_indices + 3 * (size_t)hs.T);
    for (uint32_t v : hs.vidx) BCS_REQUIRE(v < (uint32_t)hs.V, BCS_ERR_INVALID, "vein index out of range");

    // grid bounds = vein bounding box widened by the margins (vein_factory.hpp:21-86)
    const std::vector<float>* axis[3] = {&hs.vx, &hs.vy, &hs.vz};
    for (int d = 0; d < 3; ++d) {
        const auto [lo, hi] = std::minmax_element(axis[d]->begin(), axis[d]->end());
        const float margin = d == 1 ? hs.ph.grid_y_margin : hs.ph.grid_xz_margin;
        hs.gmin[d] = *lo - margin;
        hs.gmax[d] = *hi + margin;
        hs.gsize[d] = hs.gmax[d] - hs.gmin[d];
    }

    // neighbour slots (vein_factory.hpp:130-174): every triangle contributes both other corners to each
    // corner, lists are sorted and cut to 9 entries, duplicates included
    std::vector<int> degree(hs.V, 0);
    for (uint32_t v : hs.vidx) degree[v] += 2;
    std::vector<size_t> offset(hs.V + 1, 0);
    for (int i = 0; i < hs.V; ++i) offset[i + 1] = offset[i] + degree[i];
    std::vector<uint32_t> flat(offset[hs.V]);
    std::vector<size_t> cursor(offset.begin(), offset.end() - 1);
    for (int t = 0; t < hs.T; ++t) {
        const uint32_t c[3] = {hs.vidx[3 * t], hs.vidx[3 * t + 1], hs.vidx[3 * t + 2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                if (a != b) flat[cursor[c[a]]++] = c[b];
    }
    hs.nbrIds.assign((size_t)BCS_VEIN_MAX_NEIGHBORS * hs.V, -1);
    hs.nbrLen.assign((size_t)BCS_VEIN_MAX_NEIGHBORS * hs.V, -1.0f);
    for (int i = 0; i < hs.V; ++i) {
        auto first = flat.begin() + offset[i], last = flat.begin() + offset[i + 1];
        std::sort(first, last);
        const int keep = (int)std::min<size_t>(last - first, BCS_VEIN_MAX_NEIGHBORS);
        for (int s = 0; s < keep; ++s) {
            const uint32_t q = first[s];
            const float dx = hs.vx[i] - hs.vx[q], dy = hs.vy[i] - hs.vy[q], dz = hs.vz[i] - hs.vz[q];
            hs.nbrIds[(size_t)s * hs.V + i] = (int32_t)q;
            hs.nbrLen[(size_t)s * hs.V + i] = std::fabs(std::sqrt(dx * dx + dy * dy + dz * dz));
        }
    }

    hs.endC.assign(in.ending_centers, in.ending_centers + 3 * (size_t)std::max(0, in.n_endings));
    hs.endR.assign(in.ending_radii, in.ending_radii + std::max(0, in.n_endings));
}

}  // namespace

void derive_scene(const bcs_scene& in, HostScene& hs)
{
    BCS_REQUIRE(in.struct_size == sizeof(bcs_scene), BCS_ERR_INVALID, "bcs_scene.struct_size mismatch (ABI version?)");
    hs.ph = in.physics;
    hs.useBloodFlow = in.use_blood_flow;
    hs.reactionForce = in.enable_reaction_force;
    hs.bigBrake = in.enable_big_cells_brake;
    hs.bsCoeff = in.bounding_spheres_coeff;
    BCS_REQUIRE(hs.bsCoeff > 0, BCS_ERR_INVALID, "bounding_spheres_coeff must be positive");
    for (int d = 0; d < 3; ++d) {
        hs.cellSize[d] = in.cell_size[d];
        hs.triCellSize[d] = in.tri_cell_size[d];
        BCS_REQUIRE(hs.cellSize[d] > 0 && hs.triCellSize[d] > 0, BCS_ERR_INVALID, "grid cell sizes must be positive");
    }
    build_types(in, hs);
    build_radii(hs);
    build_vein(in, hs);
    long long cells = 1, tcells = 1;
    for (int d = 0; d < 3; ++d) {
        // static_cast<int>(width / cellWidth) (uniform_grid.cu:86-88)
        hs.gdims[d] = (int)(hs.gsize[d] / (float)hs.cellSize[d]);
        hs.tdims[d] = (int)(hs.gsize[d] / (float)hs.triCellSize[d]);
        BCS_REQUIRE(hs.gdims[d] > 0 && hs.tdims[d] > 0, BCS_ERR_INVALID, "grid has no cells along an axis");
        cells *= hs.gdims[d];
        tcells *= hs.tdims[d];
    }
    BCS_REQUIRE(cells < (1ll << 31) && tcells < (1ll << 31), BCS_ERR_UNSUPPORTED, "grid has more than 2^31 cells");
}

}  // namespace bcs

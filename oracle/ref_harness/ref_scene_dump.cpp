// TEST INFRASTRUCTURE (not product).  Compiles the REAL reference config +
// meta_factory headers where they lie (-I <reference>/src -I <reference>/Libraries/include)
// and serialises every table they derive into a BCSD scene file.  The scene
// file is (a) the input of all three implementations (product, oracle, reference
// harness use identical bytes) and (b) the golden for the product's and the
// oracle's own re-derivation of these tables (tests/test_scene_tables.py).
//
// Reference symbols read (all constexpr / inline const in the reference tree):
//   meta_factory/blood_cell_factory.hpp:162-334  BloodCellList, particleStarts, bloodCellTypesStarts,
//                                                bloodCellModelStarts, accumulatedGraphSizes, springGraph
//   meta_factory/vein_factory.hpp:12-185         veinPositionCount, triangleCount, min/max X/Y/Z, width/height/depth,
//                                                cylinderRadius, calculateSpringLengths(), veinVertexMaxNeighbors
//   config/simulation.hpp, config/physics.hpp    every constant
//   config/vein_definition.hpp                   veinPositions, veinIndices, VeinEndingCenters/Radii
//
// Build: see oracle/build_ref.sh.  Usage: ref_scene_dump <out.bcsd>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "meta_factory/blood_cell_factory.hpp"
#include "meta_factory/vein_factory.hpp"
#include "config/physics.hpp"
#include "config/simulation.hpp"

#include "bcsd_io.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <out.bcsd>\n", argv[0]);
        return 2;
    }
    bcsd::Writer w(argv[1]);

    // ---- blood cell types, in the order the reference's mp_sort produced -----------------------
    std::vector<int32_t> types, typeStarts, springCounts;
    std::vector<float> mx, my, mz;
    using TypeIdx = mp_iota_c<bloodCellTypeCount>;
    mp_for_each<TypeIdx>([&](auto i) {
        using Def = mp_at_c<BloodCellList, i>;
        types.push_back(Def::count);
        types.push_back(Def::particlesInCell);
        typeStarts.push_back(particleStarts[i]);
        typeStarts.push_back(bloodCellTypesStarts[i]);
        typeStarts.push_back(bloodCellModelStarts[i]);
        typeStarts.push_back(accumulatedGraphSizes[i]);
        springCounts.push_back((int32_t)mp_size<typename Def::List>::value);
        using Verts = typename Def::Vertices;
        mp_for_each<mp_iota_c<Def::particlesInCell>>([&](auto j) {
            mx.push_back(mp_at_c<Verts, j>::x);
            my.push_back(mp_at_c<Verts, j>::y);
            mz.push_back(mp_at_c<Verts, j>::z);
        });
    });
    w.put("types", types);
    w.put("type_starts", typeStarts);
    w.put("spring_counts", springCounts);
    w.put("model_x", mx);
    w.put("model_y", my);
    w.put("model_z", mz);
    w.put("spring_graph", springGraph.data(), springGraph.size());

    // the user's list order (before fold/unique/sort), for the ordering test
    std::vector<int32_t> userTypes;
    mp_for_each<mp_iota_c<mp_size<UserDefinedBloodCellList>::value>>([&](auto i) {
        using Def = mp_at_c<UserDefinedBloodCellList, i>;
        userTypes.push_back(Def::count);
        userTypes.push_back(Def::particlesInCell);
        userTypes.push_back((int32_t)mp_size<typename Def::List>::value);
    });
    w.put("user_types", userTypes);   // (count, P, nSprings) per user definition, in the user's order

    // user-level definitions (what a client writes into config/blood_cells_definition.hpp), flattened
    std::vector<int32_t> uSpringSE;   // start,end pairs
    std::vector<float> uSpringLen, uVerts;
    mp_for_each<mp_iota_c<mp_size<UserDefinedBloodCellList>::value>>([&](auto i) {
        using Def = mp_at_c<UserDefinedBloodCellList, i>;
        using SL = typename Def::List;
        mp_for_each<mp_iota_c<mp_size<SL>::value>>([&](auto j) {
            using S = mp_at_c<SL, j>;
            uSpringSE.push_back(S::start);
            uSpringSE.push_back(S::end);
            uSpringLen.push_back(S::length);
        });
        using Verts = typename Def::Vertices;
        mp_for_each<mp_iota_c<Def::particlesInCell>>([&](auto j) {
            uVerts.push_back(mp_at_c<Verts, j>::x);
            uVerts.push_back(mp_at_c<Verts, j>::y);
            uVerts.push_back(mp_at_c<Verts, j>::z);
        });
    });
    w.put("user_spring_se", uSpringSE);
    w.put("user_spring_len", uSpringLen);
    w.put("user_vertices", uVerts);

    std::vector<int32_t> totals = {particleCount, bloodCellCount, bloodCellTypeCount, particleDistinctCellsCount,
                                   totalGraphSize};
    w.put("totals", totals);

    // ---- grid ------------------------------------------------------------------------------
    std::vector<float> gmin = {minX, minY, minZ}, gmax = {maxX, maxY, maxZ}, whd = {width, height, depth};
    w.put("grid_min", gmin);
    w.put("grid_max", gmax);
    w.put("grid_whd", whd);
    std::vector<int32_t> cs = {cellWidth, cellHeight, cellDepth};
    std::vector<int32_t> tcs = {cellWidthTriangles, cellHeightTriangles, cellDepthTriangles};
    w.put("cell_size", cs);
    w.put("tri_cell_size", tcs);

    // ---- vein ------------------------------------------------------------------------------
    std::vector<float> vx(veinPositionCount), vy(veinPositionCount), vz(veinPositionCount);
    for (int i = 0; i < veinPositionCount; ++i) {
        vx[i] = veinPositions[i].x;
        vy[i] = veinPositions[i].y;
        vz[i] = veinPositions[i].z;
    }
    w.put("vein_x", vx);
    w.put("vein_y", vy);
    w.put("vein_z", vz);
    w.put("vein_indices", veinIndices.data(), veinIndices.size());

    {
        const auto pair = calculateSpringLengths();
        const auto& ids = std::get<0>(pair);
        const auto& lens = std::get<1>(pair);
        std::vector<int32_t> nbr;
        std::vector<float> len;
        for (int s = 0; s < veinVertexMaxNeighbors; ++s) {
            nbr.insert(nbr.end(), ids[s].begin(), ids[s].end());
            len.insert(len.end(), lens[s].begin(), lens[s].end());
        }
        w.put("vein_nbr_ids", nbr);   // [slot][vertex]
        w.put("vein_nbr_len", len);
    }

    std::vector<float> ec, er;
    mp_for_each<mp_iota_c<veinEndingCenterCount>>([&](auto i) {
        ec.push_back(mp_at_c<VeinEndingCenters, i>::x);
        ec.push_back(mp_at_c<VeinEndingCenters, i>::y);
        ec.push_back(mp_at_c<VeinEndingCenters, i>::z);
        er.push_back(mp_at_c<VeinEndingRadii, i>::value);
    });
    w.put("ending_centers", ec);
    w.put("ending_radii", er);

    // ---- constants (order fixed; twin: simulation-server_b200/scene.py PHYSICS_FIELDS) ------------
    std::vector<float> phys = {
        dt,
        velocity_collision_damping,
        particle_k_sniff,
        vein_k_sniff,
        particle_d_fact,
        vein_d_fact,
        vein_boundaries_velocity_damping,
        vein_collision_force_intensity,
        viscous_damping,
        collisionSpringCoeff,
        collisionDampingCoeff,
        collistionShearCoeff,
        maxCellSizeFactorBeforeBrake,
        bigParticleBrakingIntensity,
        initVelocityX,
        initVelocityY,
        initVelocityZ,
        randomVelocityModifier,
        veinImpactDistance,
        veinImpactMinimalForceDistance,
        Gx,
        Gy,
        Gz,
        gridYMargin,
        gridXZMargin,
        minSpawnY,
        cylinderRadius,
    };
    w.put("physics", phys);
    std::vector<int32_t> flags = {useBloodFlow ? 1 : 0, enableReactionForce ? 1 : 0, enableBigCellsBrake ? 1 : 0,
                                  boundingSpheresCoeff, maxFrames, gpuCount};
    w.put("flags", flags);
    std::printf("scene: N=%d cells=%d types=%d V=%d T=%d\n", particleCount, bloodCellCount, bloodCellTypeCount,
                veinPositionCount, triangleCount);
    return 0;
}

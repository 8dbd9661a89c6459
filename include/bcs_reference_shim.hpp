// bcs_reference_shim.hpp - fills a bcs_scene from the reference's OWN compile-time config.
//
// This is the piece a Simulation-Server maintainer compiles inside their tree: include it after the
// reference's headers (meta_factory/blood_cell_factory.hpp, meta_factory/vein_factory.hpp,
// config/physics.hpp, config/simulation.hpp - they bring UserDefinedBloodCellList, veinPositions,
// veinIndices, VeinEndingCenters/Radii and every constant) and call bcs_shim::fill(storage).
// The src/config headers stay the API surface: whatever the client app writes there
// (README.md:32 of the reference) is what libbcs simulates.  Nothing here depends on CUDA.
//
// Only USER-level data is forwarded (definitions in the user's order, vein mesh, constants); the
// fold/unique/sort of blood_cell_factory.hpp:60-162, the spring matrix, the grid bounds, the vein neighbour
// slots and the bounding-sphere radii are re-derived inside libbcs and can be cross-checked against the
// reference's own constexpr tables with bcs_get_layout()/bcs_get_table() (tests do exactly that).
#pragma once

#include <cstdint>
#include <vector>

#include "bcs.h"

namespace bcs_shim {

struct SceneStorage {
    std::vector<bcs_cell_def> defs;
    std::vector<std::vector<bcs_spring>> springs;
    std::vector<std::vector<float>> vertices;
    std::vector<float> vx, vy, vz, endingCenters, endingRadii;
    std::vector<uint32_t> indices;
    bcs_scene scene{};
};

inline const bcs_scene& fill(SceneStorage& st)
{
    using namespace boost::mp11;
    constexpr int nDefs = (int)mp_size<UserDefinedBloodCellList>::value;
    st.defs.resize(nDefs);
    st.springs.resize(nDefs);
    st.vertices.resize(nDefs);
    mp_for_each<mp_iota_c<nDefs>>([&](auto i) {
        using Def = mp_at_c<UserDefinedBloodCellList, i>;
        using Springs = typename Def::List;
        using Verts = typename Def::Vertices;
        mp_for_each<mp_iota_c<mp_size<Springs>::value>>([&](auto j) {
            using S = mp_at_c<Springs, j>;
            st.springs[i].push_back(bcs_spring{S::start, S::end, S::length});
        });
        mp_for_each<mp_iota_c<Def::particlesInCell>>([&](auto j) {
            using V = mp_at_c<Verts, j>;
            st.vertices[i].push_back(V::x);
            st.vertices[i].push_back(V::y);
            st.vertices[i].push_back(V::z);
        });
        st.defs[i] = bcs_cell_def{Def::count, Def::particlesInCell, (int32_t)st.springs[i].size(), st.springs[i].data(),
                                  st.vertices[i].data()};
    });

    const int V = (int)veinPositions.size();
    st.vx.resize(V); st.vy.resize(V); st.vz.resize(V);
    for (int k = 0; k < V; ++k) {
        st.vx[k] = veinPositions[k].x;
        st.vy[k] = veinPositions[k].y;
        st.vz[k] = veinPositions[k].z;
    }
    st.indices.assign(veinIndices.begin(), veinIndices.end());
    mp_for_each<mp_iota_c<mp_size<VeinEndingCenters>::value>>([&](auto e) {
        using C = mp_at_c<VeinEndingCenters, e>;
        st.endingCenters.push_back(C::x);
        st.endingCenters.push_back(C::y);
        st.endingCenters.push_back(C::z);
        st.endingRadii.push_back(mp_at_c<VeinEndingRadii, e>::value);
    });

    bcs_scene& s = st.scene;
    s.struct_size = sizeof(bcs_scene);
    s.n_defs = nDefs;
    s.defs = st.defs.data();
    s.n_vertices = V;
    s.vein_x = st.vx.data(); s.vein_y = st.vy.data(); s.vein_z = st.vz.data();
    s.n_triangles = (int32_t)(st.indices.size() / 3);
    s.vein_indices = st.indices.data();
    s.n_endings = (int32_t)st.endingRadii.size();
    s.ending_centers = st.endingCenters.data();
    s.ending_radii = st.endingRadii.data();
    s.cell_size[0] = cellWidth; s.cell_size[1] = cellHeight; s.cell_size[2] = cellDepth;
    s.tri_cell_size[0] = cellWidthTriangles; s.tri_cell_size[1] = cellHeightTriangles; s.tri_cell_size[2] = cellDepthTriangles;
    s.use_blood_flow = useBloodFlow ? 1 : 0;
    s.enable_reaction_force = enableReactionForce ? 1 : 0;
    s.enable_big_cells_brake = enableBigCellsBrake ? 1 : 0;
    s.bounding_spheres_coeff = boundingSpheresCoeff;
    bcs_physics& p = s.physics;
    p.dt = dt;
    p.velocity_collision_damping = velocity_collision_damping;
    p.particle_k_sniff = particle_k_sniff; p.vein_k_sniff = vein_k_sniff;
    p.particle_d_fact = particle_d_fact; p.vein_d_fact = vein_d_fact;
    p.vein_boundaries_velocity_damping = vein_boundaries_velocity_damping;
    p.vein_collision_force_intensity = vein_collision_force_intensity;
    p.viscous_damping = viscous_damping;
    p.collision_spring_coeff = collisionSpringCoeff; p.collision_damping_coeff = collisionDampingCoeff;
    p.collision_shear_coeff = collistionShearCoeff;
    p.max_cell_size_factor_before_brake = maxCellSizeFactorBeforeBrake;
    p.big_particle_braking_intensity = bigParticleBrakingIntensity;
    p.init_velocity[0] = initVelocityX; p.init_velocity[1] = initVelocityY; p.init_velocity[2] = initVelocityZ;
    p.random_velocity_modifier = randomVelocityModifier;
    p.vein_impact_distance = veinImpactDistance;
    p.vein_impact_minimal_force_distance = veinImpactMinimalForceDistance;
    p.gravity[0] = Gx; p.gravity[1] = Gy; p.gravity[2] = Gz;
    p.grid_y_margin = gridYMargin; p.grid_xz_margin = gridXZMargin;
    p.min_spawn_y = minSpawnY;
    p.cylinder_radius = cylinderRadius;
    return s;
}

}  // namespace bcs_shim

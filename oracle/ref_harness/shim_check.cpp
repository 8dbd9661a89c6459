// TEST INFRASTRUCTURE: compiles include/bcs_reference_shim.hpp INSIDE the reference's header tree (the way a
// maintainer would) and writes the bcs_scene it produces as the user-level arrays of a BCSD scene file, so
// that tests can compare it with what ref_scene_dump.cpp extracted independently.  usage: shim_check <out.bcsd>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "meta_factory/blood_cell_factory.hpp"
#include "meta_factory/vein_factory.hpp"
#include "config/physics.hpp"
#include "config/simulation.hpp"

#include "bcs_reference_shim.hpp"
#include "bcsd_io.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    bcs_shim::SceneStorage st;
    const bcs_scene& s = bcs_shim::fill(st);
    bcsd::Writer w(argv[1]);
    std::vector<int32_t> ut, se;
    std::vector<float> sl, uv;
    for (int i = 0; i < s.n_defs; ++i) {
        const bcs_cell_def& d = s.defs[i];
        ut.push_back(d.count); ut.push_back(d.particles_in_cell); ut.push_back(d.n_springs);
        for (int k = 0; k < d.n_springs; ++k) { se.push_back(d.springs[k].start); se.push_back(d.springs[k].end); sl.push_back(d.springs[k].length); }
        uv.insert(uv.end(), d.vertices, d.vertices + 3 * d.particles_in_cell);
    }
    w.put("user_types", ut); w.put("user_spring_se", se); w.put("user_spring_len", sl); w.put("user_vertices", uv);
    w.put("vein_x", s.vein_x, s.n_vertices); w.put("vein_y", s.vein_y, s.n_vertices); w.put("vein_z", s.vein_z, s.n_vertices);
    w.put("vein_indices", s.vein_indices, 3 * (uint64_t)s.n_triangles);
    w.put("ending_centers", s.ending_centers, 3 * (uint64_t)s.n_endings); w.put("ending_radii", s.ending_radii, s.n_endings);
    std::vector<int32_t> cs(s.cell_size, s.cell_size + 3), tcs(s.tri_cell_size, s.tri_cell_size + 3);
    w.put("cell_size", cs); w.put("tri_cell_size", tcs);
    const bcs_physics& p = s.physics;
    std::vector<float> ph = {p.dt, p.velocity_collision_damping, p.particle_k_sniff, p.vein_k_sniff, p.particle_d_fact, p.vein_d_fact,
                             p.vein_boundaries_velocity_damping, p.vein_collision_force_intensity, p.viscous_damping, p.collision_spring_coeff,
                             p.collision_damping_coeff, p.collision_shear_coeff, p.max_cell_size_factor_before_brake,
                             p.big_particle_braking_intensity, p.init_velocity[0], p.init_velocity[1], p.init_velocity[2],
                             p.random_velocity_modifier, p.vein_impact_distance, p.vein_impact_minimal_force_distance, p.gravity[0],
                             p.gravity[1], p.gravity[2], p.grid_y_margin, p.grid_xz_margin, p.min_spawn_y, p.cylinder_radius};
    w.put("physics", ph);
    std::vector<int32_t> fl = {s.use_blood_flow, s.enable_reaction_force, s.enable_big_cells_brake, s.bounding_spheres_coeff, maxFrames, gpuCount};
    w.put("flags", fl);
    return 0;
}

// Headless driver: the reference's programLoop (src/main.cu:102-266) minus OpenGL / streaming / client messages,
// running on libbcs through the host mirror (bcs_host.hpp).  Scene and initial state come from BCSD files
// (the reference generates its initial state on the device from a time(0) seed, simulation_controller.cu:60-61).
//
//   bcs_headless <scene.bcsd> <state.bcsd> <frames> [out_state.bcsd] [--fused]
//
// Prints what the reference prints at exit (main.cu:248-253): total seconds, average fps, average frame time.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "bcs_host.hpp"
#include "bcsd_io.hpp"

namespace {
struct SceneFile {
    std::vector<bcs_cell_def> defs;
    std::vector<std::vector<bcs_spring>> springs;
    std::vector<std::vector<float>> verts;
    std::vector<float> vx, vy, vz, ec, er;
    std::vector<uint32_t> idx;
    bcs_scene scene{};
};

void load_scene(const std::string& path, SceneFile& f)
{
    auto a = bcsd::read_all(path);
    auto ut = bcsd::get_vec<int32_t>(a, "user_types");
    auto se = bcsd::get_vec<int32_t>(a, "user_spring_se");
    auto sl = bcsd::get_vec<float>(a, "user_spring_len");
    auto uv = bcsd::get_vec<float>(a, "user_vertices");
    const int n = (int)ut.size() / 3;
    f.defs.resize(n); f.springs.resize(n); f.verts.resize(n);
    size_t so = 0, vo = 0;
    for (int i = 0; i < n; ++i) {
        const int cnt = ut[3 * i], p = ut[3 * i + 1], ns = ut[3 * i + 2];
        for (int k = 0; k < ns; ++k) f.springs[i].push_back(bcs_spring{se[2 * (so + k)], se[2 * (so + k) + 1], sl[so + k]});
        f.verts[i].assign(uv.begin() + 3 * vo, uv.begin() + 3 * (vo + p));
        f.defs[i] = bcs_cell_def{cnt, p, ns, f.springs[i].data(), f.verts[i].data()};
        so += ns; vo += p;
    }
    f.vx = bcsd::get_vec<float>(a, "vein_x"); f.vy = bcsd::get_vec<float>(a, "vein_y"); f.vz = bcsd::get_vec<float>(a, "vein_z");
    f.idx = bcsd::get_vec<uint32_t>(a, "vein_indices");
    f.ec = bcsd::get_vec<float>(a, "ending_centers"); f.er = bcsd::get_vec<float>(a, "ending_radii");
    auto cs = bcsd::get_vec<int32_t>(a, "cell_size"), tcs = bcsd::get_vec<int32_t>(a, "tri_cell_size"), fl = bcsd::get_vec<int32_t>(a, "flags");
    auto ph = bcsd::get_vec<float>(a, "physics");
    bcs_scene& s = f.scene;
    s.struct_size = sizeof(bcs_scene);
    s.n_defs = n; s.defs = f.defs.data();
    s.n_vertices = (int)f.vx.size(); s.vein_x = f.vx.data(); s.vein_y = f.vy.data(); s.vein_z = f.vz.data();
    s.n_triangles = (int)f.idx.size() / 3; s.vein_indices = f.idx.data();
    s.n_endings = (int)f.er.size(); s.ending_centers = f.ec.data(); s.ending_radii = f.er.data();
    for (int d = 0; d < 3; ++d) { s.cell_size[d] = cs[d]; s.tri_cell_size[d] = tcs[d]; }
    s.use_blood_flow = fl[0]; s.enable_reaction_force = fl[1]; s.enable_big_cells_brake = fl[2]; s.bounding_spheres_coeff = fl[3];
    // order of the "physics" vector: oracle/ref_harness/ref_scene_dump.cpp
    bcs_physics& p = s.physics;
    p.dt = ph[0]; p.velocity_collision_damping = ph[1]; p.particle_k_sniff = ph[2]; p.vein_k_sniff = ph[3];
    p.particle_d_fact = ph[4]; p.vein_d_fact = ph[5]; p.vein_boundaries_velocity_damping = ph[6];
    p.vein_collision_force_intensity = ph[7]; p.viscous_damping = ph[8]; p.collision_spring_coeff = ph[9];
    p.collision_damping_coeff = ph[10]; p.collision_shear_coeff = ph[11]; p.max_cell_size_factor_before_brake = ph[12];
    p.big_particle_braking_intensity = ph[13]; p.init_velocity[0] = ph[14]; p.init_velocity[1] = ph[15]; p.init_velocity[2] = ph[16];
    p.random_velocity_modifier = ph[17]; p.vein_impact_distance = ph[18]; p.vein_impact_minimal_force_distance = ph[19];
    p.gravity[0] = ph[20]; p.gravity[1] = ph[21]; p.gravity[2] = ph[22]; p.grid_y_margin = ph[23]; p.grid_xz_margin = ph[24];
    p.min_spawn_y = ph[25]; p.cylinder_radius = ph[26];
}

bcs_host::Vec3Host vec3_from(const std::map<std::string, bcsd::Array>& a, const std::string& prefix)
{
    bcs_host::Vec3Host v;
    v.x = bcsd::get_vec<float>(a, prefix + "_x"); v.y = bcsd::get_vec<float>(a, prefix + "_y"); v.z = bcsd::get_vec<float>(a, prefix + "_z");
    return v;
}
}  // namespace

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <scene.bcsd> <state.bcsd> <frames> [out_state.bcsd] [--fused]\n", argv[0]);
        return 2;
    }
    try {
        SceneFile sf;
        load_scene(argv[1], sf);
        const int maxFrames = std::atoi(argv[3]);
        const bool fused = argc > 4 && std::strcmp(argv[argc - 1], "--fused") == 0;
        const char* outPath = (argc > 4 && std::strcmp(argv[4], "--fused") != 0) ? argv[4] : nullptr;

        bcs_host::Simulation sim(sf.scene);
        auto st = bcsd::read_all(argv[2]);
        sim.upload(BCS_PARTICLE_POS, vec3_from(st, "pos"));
        sim.upload(BCS_PARTICLE_VEL, vec3_from(st, "vel"));
        if (bcsd::has(st, "frc_x")) sim.upload(BCS_PARTICLE_FRC, vec3_from(st, "frc"));

        std::cout << "started simulating...\n";
        auto begin = std::chrono::steady_clock::now();
        int frameCount = 0;
        bool shouldBeRunning = maxFrames > 0;
        if (fused) {
            sim.step(maxFrames);
            frameCount = maxFrames;
        } else {
            // MAIN LOOP - the four calls of main.cu:175-176,199,208, in that order
            while (shouldBeRunning) {
                // the reference's own call sites, arguments included (the library owns the device arrays: `positions` names them)
                sim.particleGrid.calculateGrid(bcs_host::Vec3Host{}, sim.particleCount());
                sim.triangleCentersGrid.calculateGrid(bcs_host::Vec3Host{}, sim.layout().n_triangles);
                sim.simulationController.calculateNextFrame();
                sim.simulationController.propagateAll();
                if (++frameCount >= maxFrames) shouldBeRunning = false;
            }
        }
        sim.synchronize();
        auto end = std::chrono::steady_clock::now();
        const double seconds = std::chrono::duration<double>(end - begin).count();
        std::cout << "finished simulating in " << seconds << " seconds\n";
        std::cout << "Average framerate: " << frameCount / seconds << " fps\n";
        std::cout << "Average single frame time: " << seconds / frameCount << " s\n";

        if (outPath) {
            const int n = sim.particleCount();
            auto pos = sim.download(BCS_PARTICLE_POS, n), vel = sim.download(BCS_PARTICLE_VEL, n), frc = sim.download(BCS_PARTICLE_FRC, n);
            bcsd::Writer w(outPath);
            w.put("pos_x", pos.x); w.put("pos_y", pos.y); w.put("pos_z", pos.z);
            w.put("vel_x", vel.x); w.put("vel_y", vel.y); w.put("vel_z", vel.z);
            w.put("frc_x", frc.x); w.put("frc_y", frc.y); w.put("frc_z", frc.z);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "bcs_headless: %s\n", e.what());
        return 1;
    }
}

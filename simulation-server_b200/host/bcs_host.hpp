// C++17 host-side mirror of the reference's simulation-loop interface, over the C ABI of libbcs.
//
// The reference's frame loop (src/main.cu:158-247) talks to four objects:
//     particleGrid.calculateGrid(positions, particleCount);            // grids/uniform_grid.cuh:53
//     triangleCentersGrid.calculateGrid(triangles.centers, T);
//     simulationController.calculateNextFrame();                       // simulation/simulation_controller.cuh:33
//     simulationController.propagateAll();                             // :38
// and reads bloodCells.particles.positions[0] / triangles.positions[0] for rendering (main.cu:183-184).
// The classes below keep those names and that call order, so a headless copy of the loop body needs no edits
// beyond the constructors; errors surface as exceptions instead of the reference's printf+exit
// (utilities/cuda_handle_error.cuh:18-25).
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "bcs.h"

namespace bcs_host {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc, const char* what)
{
    if (rc != BCS_OK) throw Error(rc, std::string(what) + ": " + bcs_last_error());
}

struct Vec3Host {   // host image of a cudaVec3 (utilities/cuda_vec3.cuh:11-73)
    std::vector<float> x, y, z;
    explicit Vec3Host(size_t n = 0) : x(n), y(n), z(n) {}
    size_t size() const { return x.size(); }
};

class Simulation;

// Stand-in for UniformGrid (grids/uniform_grid.cuh:15-65): calculateGrid() rebuilds the grid it names.
class UniformGrid {
    Simulation* sim_;
    int which_;   // 0 particles, 1 triangle centres
public:
    UniformGrid(Simulation* s, int which) : sim_(s), which_(which) {}
    // The reference's signature (grids/uniform_grid.cuh:53: calculateGrid(const cudaVec3& positions, int objectCount)), so
    // that main.cu:175-176 compiles unchanged.  The library owns the device arrays, so `positions` only names them (any
    // type is accepted and ignored); objectCount must be the number of objects of the grid.
    template <class Positions>
    void calculateGrid(const Positions& positions, int objectCount);
    void calculateGrid();
    void download(std::vector<int32_t>& cellIds, std::vector<int32_t>& objectIds);
};

// Stand-in for sim::SimulationController (simulation/simulation_controller.cuh:24-67)
class SimulationController {
    Simulation* sim_;
public:
    explicit SimulationController(Simulation* s) : sim_(s) {}
    void calculateNextFrame();
    void propagateAll();
};

class Simulation {
    bcs_sim* h_ = nullptr;
    bcs_layout layout_{};
public:
    UniformGrid particleGrid{this, 0};
    UniformGrid triangleCentersGrid{this, 1};
    SimulationController simulationController{this};

    Simulation(const bcs_scene& scene, const bcs_opts* opts = nullptr)
    {
        check(bcs_create(&scene, opts, &h_), "bcs_create");
        check(bcs_get_layout(h_, &layout_), "bcs_get_layout");
    }
    ~Simulation() { bcs_destroy(h_); }
    Simulation(const Simulation&) = delete;
    Simulation& operator=(const Simulation&) = delete;

    bcs_sim* handle() const { return h_; }
    const bcs_layout& layout() const { return layout_; }
    int particleCount() const { return layout_.n_particles; }

    void upload(int which, const Vec3Host& v) { check(bcs_upload(h_, which, v.x.data(), v.y.data(), v.z.data(), (int32_t)v.size()), "bcs_upload"); }
    Vec3Host download(int which, size_t n)
    {
        Vec3Host v(n);
        check(bcs_download(h_, which, v.x.data(), v.y.data(), v.z.data(), (int32_t)n), "bcs_download");
        return v;
    }
    // whole steps without host round trips (the loop body as one CUDA-graph replay)
    void step(int n) { check(bcs_step(h_, n), "bcs_step"); }
    void synchronize() { check(bcs_synchronize(h_), "bcs_synchronize"); }
};

template <class Positions>
inline void UniformGrid::calculateGrid(const Positions&, int objectCount)
{
    const int n = which_ == 0 ? sim_->layout().n_particles : sim_->layout().n_triangles;
    if (objectCount != n) throw Error(BCS_ERR_INVALID, "calculateGrid: objectCount does not match the grid's object count");
    calculateGrid();
}
inline void UniformGrid::calculateGrid()
{
    check(bcs_run_stage(sim_->handle(), which_ == 0 ? BCS_STAGE_GRID_PARTICLES : BCS_STAGE_GRID_TRIANGLES), "calculateGrid");
}
inline void UniformGrid::download(std::vector<int32_t>& cellIds, std::vector<int32_t>& objectIds)
{
    const int n = which_ == 0 ? sim_->layout().n_particles : sim_->layout().n_triangles;
    cellIds.resize(n);
    objectIds.resize(n);
    check(bcs_download_grid(sim_->handle(), which_, cellIds.data(), objectIds.data(), n), "bcs_download_grid");
}
inline void SimulationController::calculateNextFrame() { check(bcs_compute_forces(sim_->handle()), "calculateNextFrame"); }
inline void SimulationController::propagateAll() { check(bcs_integrate(sim_->handle()), "propagateAll"); }

}  // namespace bcs_host

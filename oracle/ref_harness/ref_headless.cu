// TEST INFRASTRUCTURE (not product): headless driver around the UNMODIFIED reference CUDA sources.
//
// This translation unit textually includes the reference's simulation_controller.cu from where it
// lies (-I <reference>/src); the other seven hot-path .cu files are compiled separately from the
// reference tree and linked in (oracle/build_ref.sh).  Nothing of the reference is copied into this
// repository.  What this file adds:
//   * state upload from a BCSD file (bypasses the time(0) cuRAND seeding, simulation_controller.cu:60-61)
//   * zero-initialisation of the never-initialised buffers (grid tables, vein velocities/forces)
//   * "plain" mode  : the loop body of main.cu:175-176,199,208 (calculateGrid x2, calculateNextFrame, propagateAll)
//   * "staged" mode : the same launches in the same order, issued stage by stage with a dump after each
//                     stage (stage order: simulation_controller.cu:246-331)
//   * "bench" mode  : plain loop, wall clock, for the "reference CUDA build on the same box" baseline
//
// usage: ref_headless <state.bcsd|-> <out_dir> <mode:plain|staged|bench> <nsteps> [dump_step ...]
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <variant>
#include <vector>

#include "bcsd_io.hpp"

#define private public
#include "simulation/simulation_controller.cu"
#undef private

namespace {

std::vector<float> d2h_f(const float* d, size_t n)
{
    std::vector<float> h(n);
    CUDACHECK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
    return h;
}
std::vector<int32_t> d2h_i(const int* d, size_t n)
{
    std::vector<int32_t> h(n);
    CUDACHECK(cudaMemcpy(h.data(), d, n * sizeof(int), cudaMemcpyDeviceToHost));
    return h;
}
void h2d_f(float* d, const std::vector<float>& h, size_t n)
{
    if (h.size() != n) throw std::runtime_error("state array has wrong length");
    CUDACHECK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
}

void put_vec3(bcsd::Writer& w, const std::string& name, const cudaVec3& v, size_t n)
{
    w.put(name + "_x", d2h_f(v.x, n));
    w.put(name + "_y", d2h_f(v.y, n));
    w.put(name + "_z", d2h_f(v.z, n));
}

void put_grid(bcsd::Writer& w, const std::string& name, UniformGrid& g)
{
    w.put(name + ".keys", d2h_i(g.gridCellIds[0], g.objectCount));
    w.put(name + ".ids", d2h_i(g.particleIds[0], g.objectCount));
    auto starts = d2h_i(g.gridCellStarts[0], g.cellCount);
    auto ends = d2h_i(g.gridCellEnds[0], g.cellCount);
    // sparse form: every cell whose (start,end) differs from the zero-initialised (0,0)
    std::vector<int32_t> cells, s, e;
    for (int c = 0; c < g.cellCount; ++c) {
        if (starts[c] != 0 || ends[c] != 0) {
            cells.push_back(c);
            s.push_back(starts[c]);
            e.push_back(ends[c]);
        }
    }
    w.put(name + ".table_cells", cells);
    w.put(name + ".table_starts", s);
    w.put(name + ".table_ends", e);
    std::vector<int32_t> dims = {g.cellCountX, g.cellCountY, g.cellCountZ, g.cellCount};
    w.put(name + ".dims", dims);
}

void put_particles(bcsd::Writer& w, const std::string& stage, BloodCells& bc, bool pos, bool vel, bool frc)
{
    if (pos) put_vec3(w, stage + ".pos", bc.particles.positions[0], particleCount);
    if (vel) put_vec3(w, stage + ".vel", bc.particles.velocities[0], particleCount);
    if (frc) put_vec3(w, stage + ".frc", bc.particles.forces[0], particleCount);
}

void put_vein(bcsd::Writer& w, const std::string& stage, VeinTriangles& t, bool pos, bool vel, bool frc)
{
    if (pos) put_vec3(w, stage + ".vein_pos", t.positions[0], veinPositionCount);
    if (vel) put_vec3(w, stage + ".vein_vel", t.velocities[0], veinPositionCount);
    if (frc) put_vec3(w, stage + ".vein_frc", t.forces[0], veinPositionCount);
}

void zero_vec3(cudaVec3& v, size_t n)
{
    CUDACHECK(cudaMemset(v.x, 0, n * sizeof(float)));
    CUDACHECK(cudaMemset(v.y, 0, n * sizeof(float)));
    CUDACHECK(cudaMemset(v.z, 0, n * sizeof(float)));
}

void zero_grid(UniformGrid& g)
{
    CUDACHECK(cudaMemset(g.gridCellIds[0], 0, g.objectCount * sizeof(int)));
    CUDACHECK(cudaMemset(g.particleIds[0], 0, g.objectCount * sizeof(int)));
    CUDACHECK(cudaMemset(g.gridCellStarts[0], 0, g.cellCount * sizeof(int)));
    CUDACHECK(cudaMemset(g.gridCellEnds[0], 0, g.cellCount * sizeof(int)));
}

}  // namespace

namespace sim {

// The launches of SimulationController::calculateNextFrame (simulation_controller.cu:246-313, single GPU:
// `calculate(0)`) and propagateAll (:315-331), issued one stage at a time so that a dump can be taken
// between them.  Launch shapes, streams and argument lists are the reference's.
struct StagedFrame {
    SimulationController& sc;
    BloodCells& bloodCells;
    VeinTriangles& triangles;
    UniformGrid& particleGrid;
    UniformGrid& triangleGrid;

    void veinGather()
    {
        triangles.gatherForcesFromNeighbors(0, verticesGpuStarts[0], verticesGpuEnds[0], sc.veinTrianglesThreads.blocks,
                                            sc.veinTrianglesThreads.threadsPerBlock);
        CUDACHECK(cudaDeviceSynchronize());
    }
    void springs()
    {
        bloodCells.gatherForcesFromNeighbors(0, bloodCellGpuStarts[0], bloodCellGpuEnds[0], particleGpuStarts[0],
                                             particleGpuEnds[0], sc.streams[0]);
        CUDACHECK(cudaDeviceSynchronize());
    }
    void particleCollisions()
    {
        using IndexList = mp_iota_c<bloodCellTypeCount>;
        mp_for_each<IndexList>([&](auto i) {
            using Def = mp_at_c<BloodCellList, i>;
            constexpr int pStart = particleStarts[i];
            constexpr int mStart = bloodCellModelStarts[i];
            CudaThreads threads(Def::count * Def::particlesInCell);
            calculateParticleCollisions<<<threads.blocks, threads.threadsPerBlock, 0, sc.streams[0][i]>>>(
                0, particleGpuStarts[0], particleGpuEnds[0], bloodCells, particleGrid, sc.cellModelsBoundingSpheres[0],
                Def::count, Def::particlesInCell, mStart, pStart);
        });
        CUDACHECK(cudaDeviceSynchronize());
    }
    void veinCollisions()
    {
        using IndexList = mp_iota_c<bloodCellTypeCount>;
        mp_for_each<IndexList>([&](auto i) {
            using Def = mp_at_c<BloodCellList, i>;
            constexpr int pStart = particleStarts[i];
            constexpr int mStart = bloodCellModelStarts[i];
            CudaThreads threads(Def::count * Def::particlesInCell);
            detectVeinCollisions<<<threads.blocks, threads.threadsPerBlock, 0, sc.streams[0][i]>>>(
                0, particleGpuStarts[0], particleGpuEnds[0], bloodCells, triangles, triangleGrid,
                sc.cellModelsBoundingSpheres[0], Def::count, Def::particlesInCell, mStart, pStart);
        });
        CUDACHECK(cudaDeviceSynchronize());
    }
    void integrateParticles()
    {
        bloodCells.propagateForcesIntoPositions(sc.bloodCellsThreads.blocks, sc.bloodCellsThreads.threadsPerBlock);
        CUDACHECK(cudaDeviceSynchronize());
    }
    void integrateVein()
    {
        triangles.propagateForcesIntoPositions(sc.veinVerticesThreads.blocks, sc.veinVerticesThreads.threadsPerBlock);
        CUDACHECK(cudaDeviceSynchronize());
    }
    void veinEnd()
    {
        if constexpr (useBloodFlow) {
            HandleVeinEnd(bloodCells, sc.devStates, sc.streams[0], sc.bloodCellModels);
            CUDACHECK(cudaDeviceSynchronize());
        }
    }
};

}  // namespace sim

int main(int argc, char** argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s <state.bcsd|-> <out_dir> <plain|staged|bench> <nsteps> [dump_step ...]\n", argv[0]);
        return 2;
    }
    const std::string statePath = argv[1], outDir = argv[2], mode = argv[3];
    const int nsteps = std::atoi(argv[4]);
    std::set<int> dumpSteps;
    for (int i = 5; i < argc; ++i) dumpSteps.insert(std::atoi(argv[i]));

    CUDACHECK(cudaSetDevice(0));
    {
        cudaDeviceProp prop;
        CUDACHECK(cudaGetDeviceProperties(&prop, 0));
        std::printf("ref_headless: device %s sm_%d%d  N=%d cells=%d types=%d V=%d T=%d\n", prop.name, prop.major, prop.minor,
                    particleCount, bloodCellCount, bloodCellTypeCount, veinPositionCount, triangleCount);
    }

    // Construction order of programLoop (main.cu:122-138)
    BloodCells bloodCells;
    VeinTriangles triangles;
    UniformGrid particleGrid(particleGridGpu, particleCount, cellWidth, cellHeight, cellDepth);
    UniformGrid triangleCentersGrid(veinGridGpu, triangleCount, cellWidthTriangles, cellHeightTriangles, cellDepthTriangles);
    sim::SimulationController sc(bloodCells, triangles, &particleGrid, &triangleCentersGrid);
    CUDACHECK(cudaDeviceSynchronize());

    // Deterministic start: zero what the reference leaves uninitialised.
    zero_grid(particleGrid);
    zero_grid(triangleCentersGrid);
    zero_vec3(triangles.velocities[0], veinPositionCount);
    zero_vec3(triangles.forces[0], veinPositionCount);

    if (statePath != "-") {
        auto st = bcsd::read_all(statePath);
        h2d_f(bloodCells.particles.positions[0].x, bcsd::get_vec<float>(st, "pos_x"), particleCount);
        h2d_f(bloodCells.particles.positions[0].y, bcsd::get_vec<float>(st, "pos_y"), particleCount);
        h2d_f(bloodCells.particles.positions[0].z, bcsd::get_vec<float>(st, "pos_z"), particleCount);
        h2d_f(bloodCells.particles.velocities[0].x, bcsd::get_vec<float>(st, "vel_x"), particleCount);
        h2d_f(bloodCells.particles.velocities[0].y, bcsd::get_vec<float>(st, "vel_y"), particleCount);
        h2d_f(bloodCells.particles.velocities[0].z, bcsd::get_vec<float>(st, "vel_z"), particleCount);
        if (bcsd::has(st, "frc_x")) {
            h2d_f(bloodCells.particles.forces[0].x, bcsd::get_vec<float>(st, "frc_x"), particleCount);
            h2d_f(bloodCells.particles.forces[0].y, bcsd::get_vec<float>(st, "frc_y"), particleCount);
            h2d_f(bloodCells.particles.forces[0].z, bcsd::get_vec<float>(st, "frc_z"), particleCount);
        } else {
            zero_vec3(bloodCells.particles.forces[0], particleCount);
        }
    }

    {   // setup tables the reference derives on the host at construction (simulation_controller.cu:93-153)
        bcsd::Writer w(outDir + "/setup.bcsd");
        w.put("bounding_spheres", d2h_f(sc.cellModelsBoundingSpheres[0], particleDistinctCellsCount));
        w.put("initial_radiuses", d2h_f(bloodCells.initialRadiuses[0], particleDistinctCellsCount));
        put_vec3(w, "models", sc.bloodCellModels, particleDistinctCellsCount);
        put_vec3(w, "tri_centers", triangles.centers, triangleCount);
        put_particles(w, "init", bloodCells, true, true, true);
        std::vector<float> srt(sc.smallestRadiusInType.begin(), sc.smallestRadiusInType.end());
        w.put("smallest_radius_in_type", srt);
    }

    sim::StagedFrame sf{sc, bloodCells, triangles, particleGrid, triangleCentersGrid};

    if (mode == "bench") {
        const int warm = dumpSteps.empty() ? 20 : *dumpSteps.begin();
        auto loop = [&](int n) {
            for (int s = 0; s < n; ++s) {
                particleGrid.calculateGrid(bloodCells.particles.positions[particleGridGpu], particleCount);
                triangleCentersGrid.calculateGrid(triangles.centers, triangleCount);
                sc.calculateNextFrame();
                sc.propagateAll();
            }
            CUDACHECK(cudaDeviceSynchronize());
        };
        loop(warm);
        auto t0 = std::chrono::steady_clock::now();
        loop(nsteps);
        auto t1 = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / nsteps;
        std::printf("{\"impl\": \"reference_cuda\", \"particles\": %d, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, "
                    "\"particle_steps_per_s\": %.6e}\n",
                    particleCount, nsteps, warm, ms, particleCount / (ms * 1e-3));
        return 0;
    }

    for (int step = 1; step <= nsteps; ++step) {
        const bool dump = dumpSteps.count(step) != 0;
        char fn[64];
        std::snprintf(fn, sizeof fn, "/step%05d.bcsd", step);
        if (mode == "plain") {
            particleGrid.calculateGrid(bloodCells.particles.positions[particleGridGpu], particleCount);
            triangleCentersGrid.calculateGrid(triangles.centers, triangleCount);
            sc.calculateNextFrame();
            if (dump) {
                bcsd::Writer w(outDir + fn);
                put_particles(w, "forces", bloodCells, false, true, true);
                put_vein(w, "forces", triangles, false, false, true);
                sc.propagateAll();
                put_particles(w, "end", bloodCells, true, true, true);
                put_vein(w, "end", triangles, true, true, true);
            } else {
                sc.propagateAll();
            }
            continue;
        }
        // staged
        if (!dump) {
            particleGrid.calculateGrid(bloodCells.particles.positions[particleGridGpu], particleCount);
            triangleCentersGrid.calculateGrid(triangles.centers, triangleCount);
            sf.veinGather(); sf.springs(); sf.particleCollisions(); sf.veinCollisions();
            sf.integrateParticles(); sf.integrateVein(); sf.veinEnd();
            continue;
        }
        bcsd::Writer w(outDir + fn);
        put_particles(w, "begin", bloodCells, true, true, true);
        put_vein(w, "begin", triangles, true, true, true);
        particleGrid.calculateGrid(bloodCells.particles.positions[particleGridGpu], particleCount);
        put_grid(w, "pgrid", particleGrid);
        triangleCentersGrid.calculateGrid(triangles.centers, triangleCount);
        put_grid(w, "tgrid", triangleCentersGrid);
        sf.veinGather();
        put_vein(w, "vein_gather", triangles, false, false, true);
        sf.springs();
        put_particles(w, "springs", bloodCells, false, false, true);
        put_vec3(w, "springs.centers", bloodCells.particleCenters[0], bloodCellCount);
        sf.particleCollisions();
        put_particles(w, "pcoll", bloodCells, false, false, true);
        sf.veinCollisions();
        put_particles(w, "vcoll", bloodCells, false, true, true);
        put_vein(w, "vcoll", triangles, false, false, true);
        sf.integrateParticles();
        put_particles(w, "integrate", bloodCells, true, true, false);
        sf.integrateVein();
        put_vein(w, "integrate", triangles, true, true, true);
        sf.veinEnd();
        put_particles(w, "end", bloodCells, true, true, true);
    }
    {
        bcsd::Writer w(outDir + "/final.bcsd");
        put_particles(w, "final", bloodCells, true, true, true);
        put_vein(w, "final", triangles, true, true, true);
        std::vector<int32_t> meta = {nsteps};
        w.put("nsteps", meta);
    }
    std::printf("ref_headless: done %d steps (%s)\n", nsteps, mode.c_str());
    return 0;
}

// sph_api.cpp -- host-side orchestration of one simulation step over the libsphk C-ABI.
// Each method states the reference code whose behaviour it reproduces (/root/reference/src/...).
// No device code here: this file is compiled by g++.
#include "sph_api.hpp"

#include <algorithm>
#include <cstdlib>
#include <limits>

using sphb200::check;

// ================================================================================================
// Engine
// ================================================================================================
namespace sphb200 {
Engine::Engine(int maxFluid, int maxBoundary, int3 cellSize, float cellLength, int3 origin) {
    // blocking stream: legacy-default-stream work of the caller (DArray memset, cudaMemcpy through the
    // raw accessors) orders with the engine's work exactly as it does in the single-stream reference
    CUDA_CALL(cudaStreamCreate(&stream_));
    sphk_grid g;
    g.cell_size[0] = cellSize.x; g.cell_size[1] = cellSize.y; g.cell_size[2] = cellSize.z;
    g.cell_length = cellLength;
    g.origin[0] = origin.x; g.origin[1] = origin.y; g.origin[2] = origin.z;
    const int rc = sphk_create(&ctx_, maxFluid, maxBoundary, &g, stream_);
    if (rc != 0) {
        // no CPU fallback: a missing device / failed allocation is reported and leaves the system inert
        printf("sphk_create failed: %s (%d)\n", sphk_error_string(rc), rc);
        ctx_ = nullptr;
    }
}
Engine::~Engine() {
    if (ctx_) { sphk_synchronize(ctx_); sphk_destroy(ctx_); }
    if (stream_) cudaStreamDestroy(stream_);
}
}  // namespace sphb200

// ================================================================================================
// Particles (Particles.h:22-25, Particles.cu:28-36)
// ================================================================================================
Particles::Particles(const std::vector<float3>& p) : pos(p.size()), vel(p.size()) {
    CUDA_CALL(cudaMemcpy(pos.addr(), p.data(), sizeof(float3) * p.size(), cudaMemcpyHostToDevice));
}

void Particles::advect(float dt) {
    // pos += dt * vel.  Not on the engine's hot path (BasicSPHSolver::advect uses the fused sphk_advect); kept for
    // API parity.  It moves particles behind the engine's back, so the packed records are re-packed before the
    // next sweep (shadowsStale).
    cudaStream_t st = nullptr;
    if (engine_ && engine_->ok()) { st = engine_->stream(); engine_->shadowsStale = true; }
    check(sphk_particles_advect(reinterpret_cast<float*>(pos.addr()), reinterpret_cast<const float*>(vel.addr()),
                                static_cast<int>(size()), dt, st), "sphk_particles_advect");
}

// ================================================================================================
// BasicSPHSolver (BasicSPHSolver.cu)
// ================================================================================================
bool BasicSPHSolver::beginStep(std::shared_ptr<SPHParticles>& fluids, const std::shared_ptr<SPHParticles>& boundaries,
                               const DArray<int>& cellStartFluid, const DArray<int>& cellStartBoundary, float radius,
                               bool neighborList, int listSkinPermille) {
    const auto& eng = fluids->engine();
    if (!eng || !eng->ok()) {
        printf("SPH solver: particles are not bound to a B200 engine (construct them through SPHSystem)\n");
        current_.ctx = nullptr;
        return false;
    }
    current_.ctx = eng->ctx();
    current_.abi.fluid = fluids->abi();
    current_.abi.boundary = boundaries->abi();
    current_.abi.cell_start_fluid = cellStartFluid.addr();
    current_.abi.cell_start_boundary = cellStartBoundary.addr();
    current_.abi.radius = radius;
    check(sphk_set_option(current_.ctx, SPHK_OPT_NEIGHBOR_LIST, neighborList ? 1 : 0), "sphk_set_option");
    check(sphk_set_option(current_.ctx, SPHK_OPT_LIST_SKIN, listSkinPermille), "sphk_set_option");
    if (eng->shadowsStale) { check(sphk_refresh(current_.ctx, &current_.abi), "sphk_refresh"); eng->shadowsStale = false; }
    return true;
}

// BasicSPHSolver::step, BasicSPHSolver.cu:237-260
void BasicSPHSolver::step(std::shared_ptr<SPHParticles>& fluids, const std::shared_ptr<SPHParticles>& boundaries,
                          const DArray<int>& cellStartFluid, const DArray<int>& cellStartBoundary, float3 spaceSize,
                          int3 cellSize, float cellLength, float radius, float dt, float rho0, float rhoB, float stiff,
                          float visc, float3 G, float surfaceTensionIntensity, float airPressure) {
    if (!beginStep(fluids, boundaries, cellStartFluid, cellStartBoundary, radius, true, 0)) return;
    const bool surface = surfaceTensionIntensity > EPSILON || airPressure > EPSILON;
    force(fluids, dt, G);
    if (fusedSweeps_) {
        // density and the colour gradient depend on positions and masses only: one sweep computes both (the
        // reference computes them in two sweeps, :277-330 and :32-83, with nothing in between that they read)
        densityAndColorGrad(nullptr, rho0, rhoB, surface);
        if (surface) produced(0, colorGradBuffer(), 3);
        diffuseAndSurface(rho0, rhoB, visc, dt, surfaceTensionIntensity, airPressure, surface, true);
        producedVel(fluids);
        check(sphk_pressure(current_.ctx, &current_.abi, rho0, stiff), "sphk_pressure");
        produced(0, fluids->getDensityPtr(), 1); produced(0, fluids->getPressurePtr(), 1);
        check(sphk_pressure_force(current_.ctx, &current_.abi, dt), "sphk_pressure_force");
    } else {
        diffuse(fluids, cellStartFluid, cellSize, cellLength, rho0, radius, visc, dt);
        producedVel(fluids);
        if (surface)
            handleSurface(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, rhoB, cellSize, cellLength, radius,
                          dt, surfaceTensionIntensity, airPressure);
        project(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, stiff, cellSize, cellLength, radius, dt);
    }
    advect(fluids, dt, spaceSize);
}

void BasicSPHSolver::densityAndColorGrad(float* alphaOrNull, float rho0, float rhoB, bool surface) {
    float* cg = reinterpret_cast<float*>(bufferColorGrad.addr());
    if (alphaOrNull) {
        if (surface) check(sphk_fused_dfsph_density_alpha_color_grad(current_.ctx, &current_.abi, alphaOrNull, cg, rho0, rhoB),
                           "sphk_fused_dfsph_density_alpha_color_grad");
        else check(sphk_dfsph_density_alpha(current_.ctx, &current_.abi, alphaOrNull), "sphk_dfsph_density_alpha");
    } else {
        if (surface) check(sphk_fused_density_color_grad(current_.ctx, &current_.abi, cg, rho0, rhoB), "sphk_fused_density_color_grad");
        else check(sphk_density(current_.ctx, &current_.abi), "sphk_density");
    }
}

void BasicSPHSolver::diffuseAndSurface(float rho0, float rhoB, float visc, float dt, float surfaceTensionIntensity,
                                       float airPressure, bool surface, bool colorGradReady) {
    float* dv = reinterpret_cast<float*>(bufferFloat3.addr());
    float* cg = reinterpret_cast<float*>(bufferColorGrad.addr());
    if (!surface) {
        check(sphk_viscosity(current_.ctx, &current_.abi, dv, rho0, visc, dt), "sphk_viscosity");
        return;
    }
    if (!colorGradReady) check(sphk_color_grad(current_.ctx, &current_.abi, cg, rho0, rhoB), "sphk_color_grad");
    check(sphk_fused_viscosity_surface(current_.ctx, &current_.abi, dv, cg, rho0, visc, dt, surfaceTensionIntensity, airPressure),
          "sphk_fused_viscosity_surface");
}

// :227-235
void BasicSPHSolver::force(std::shared_ptr<SPHParticles>&, float dt, float3 G) {
    if (!current_.ctx) return;
    const float g[3] = {G.x, G.y, G.z};
    check(sphk_gravity(current_.ctx, &current_.abi, dt, g), "sphk_gravity");
}

// :98-101
void BasicSPHSolver::advect(std::shared_ptr<SPHParticles>&, float dt, float3 spaceSize) {
    if (!current_.ctx) return;
    const float sp[3] = {spaceSize.x, spaceSize.y, spaceSize.z};
    check(sphk_advect(current_.ctx, &current_.abi, dt, sp), "sphk_advect");
}

// :167-181
void BasicSPHSolver::project(std::shared_ptr<SPHParticles>&, const std::shared_ptr<SPHParticles>&, const DArray<int>&,
                             const DArray<int>&, float rho0, float stiff, int3, float, float, float dt) {
    if (!current_.ctx) return;
    check(sphk_density(current_.ctx, &current_.abi), "sphk_density");
    check(sphk_pressure(current_.ctx, &current_.abi, rho0, stiff), "sphk_pressure");
    produced(0, current_.abi.fluid.density, 1); produced(0, current_.abi.fluid.pressure, 1);
    check(sphk_pressure_force(current_.ctx, &current_.abi, dt), "sphk_pressure_force");
}

// :211-225
void BasicSPHSolver::diffuse(std::shared_ptr<SPHParticles>&, const DArray<int>&, int3, float, float rho0, float,
                             float visc, float dt) {
    if (!current_.ctx) return;
    check(sphk_viscosity(current_.ctx, &current_.abi, reinterpret_cast<float*>(bufferFloat3.addr()), rho0, visc, dt),
          "sphk_viscosity");
}

// :262-275
void BasicSPHSolver::handleSurface(std::shared_ptr<SPHParticles>&, const std::shared_ptr<SPHParticles>&,
                                   const DArray<int>&, const DArray<int>&, float rho0, float rhoB, int3, float, float,
                                   float dt, float surfaceTensionIntensity, float airPressure) {
    if (!current_.ctx) return;
    float* cg = reinterpret_cast<float*>(bufferFloat3.addr());
    check(sphk_color_grad(current_.ctx, &current_.abi, cg, rho0, rhoB), "sphk_color_grad");
    produced(0, cg, 3);
    check(sphk_surface(current_.ctx, &current_.abi, cg, dt, rho0, surfaceTensionIntensity, airPressure), "sphk_surface");
    produced(1, current_.abi.fluid.vel, 3);
}

// ================================================================================================
// DFSPHSolver (DFSPHSolver.cu)
// ================================================================================================
// DFSPHSolver::step, DFSPHSolver.cu:33-72
void DFSPHSolver::step(std::shared_ptr<SPHParticles>& fluids, const std::shared_ptr<SPHParticles>& boundaries,
                       const DArray<int>& cellStartFluid, const DArray<int>& cellStartBoundary, float3 spaceSize,
                       int3 cellSize, float cellLength, float radius, float dt, float rho0, float rhoB, float,
                       float visc, float3 G, float surfaceTensionIntensity, float airPressure) {
    if (!beginStep(fluids, boundaries, cellStartFluid, cellStartBoundary, radius, true, 0)) return;
    const int num = static_cast<int>(fluids->size());
    const bool surface = surfaceTensionIntensity > EPSILON || airPressure > EPSILON;
    if (fusedSweeps_) {
        // density/alpha + colour gradient (positions only) + the first divergence error of :341 in one sweep
        check(sphk_fused_dfsph_density_alpha_div_error(current_.ctx, &current_.abi, alpha.addr(), surface ? colorGradBuffer() : nullptr,
                                                       rho0, rhoB, error.addr(), bufferFloat.addr(), dt),
              "sphk_fused_dfsph_density_alpha_div_error");
        produced(2, bufferFloat.addr(), 1);
        if (surface) produced(0, colorGradBuffer(), 3);
    } else {
        check(sphk_dfsph_density_alpha(current_.ctx, &current_.abi, alpha.addr()), "sphk_dfsph_density_alpha");
    }
    itDiv_ = correctDivergenceError(rho0, dt, divergenceErrorThreshold, maxIter, num, fusedSweeps_);
    force(fluids, dt, G);
    if (fusedSweeps_) {
        diffuseAndSurface(rho0, rhoB, visc, dt, surfaceTensionIntensity, airPressure, surface, true);
        producedVel(fluids);
    } else {
        BasicSPHSolver::diffuse(fluids, cellStartFluid, cellSize, cellLength, rho0, radius, visc, dt);
        producedVel(fluids);
        if (surface)
            handleSurface(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, rhoB, cellSize, cellLength, radius,
                          dt, surfaceTensionIntensity, airPressure);
    }
    itDen_ = project(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, cellSize, cellLength, radius, dt,
                     densityErrorThreshold, maxIter);
    advect(fluids, dt, spaceSize);
}

// DFSPHSolver::correctDivergenceError, DFSPHSolver.cu:331-363.  With a negative threshold the loop test
// `totalError > thr*num*rho0` is true for every possible error sum, so the (host-synchronising)
// reduction is skipped: same iteration count as the reference, no pipeline bubble (Q11).
int DFSPHSolver::correctDivergenceError(float rho0, float dt, float errorThreshold, int maxIter_, int num, bool firstErrorDone) {
    auto totalError = std::numeric_limits<float>::max();
    auto iter = 0;
    sphk_ctx* ctx = current_.ctx;
    if (!firstErrorDone) {
        check(sphk_dfsph_div_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0),
              "sphk_dfsph_div_error");
        produced(2, bufferFloat.addr(), 1);
    }
    const float thresholdCount = static_cast<float>(globalCount_ >= 0 ? globalCount_ : num);
    if (errorThreshold >= 0.0f && deviceLoops_ && !fieldHook_) {
        // the loop test runs on the device: maxIter bodies are enqueued, the ones after convergence return at once
        check(sphk_loop_begin(ctx, 0, 1, maxIter_, errorThreshold * num * rho0, 1), "sphk_loop_begin");
        for (int k = 0; k < maxIter_; ++k) {
            check(sphk_dfsph_div_correct(ctx, &current_.abi, bufferFloat.addr()), "sphk_dfsph_div_correct");
            check(sphk_dfsph_div_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0),
                  "sphk_dfsph_div_error");
            check(sphk_loop_next(ctx, 0, error.addr(), num), "sphk_loop_next");
        }
        check(sphk_loop_end(ctx, 0), "sphk_loop_end");
        return -1;      // known on the device; lastDivergenceIterations() reads it back
    }
    while ((iter < 1 || totalError > errorThreshold * thresholdCount * rho0) && iter < maxIter_) {
        check(sphk_dfsph_div_correct(ctx, &current_.abi, bufferFloat.addr()), "sphk_dfsph_div_correct");
        produced(1, current_.abi.fluid.vel, 3);
        check(sphk_dfsph_div_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0),
              "sphk_dfsph_div_error");
        produced(2, bufferFloat.addr(), 1);
        ++iter;
        if (errorThreshold >= 0.0f) totalError = reduceError(num);
    }
    return iter;
}

// DFSPHSolver::project, DFSPHSolver.cu:160-210
int DFSPHSolver::project(std::shared_ptr<SPHParticles>& fluids, const std::shared_ptr<SPHParticles>&, const DArray<int>&,
                         const DArray<int>&, float rho0, int3, float, float, float dt, float errorThreshold,
                         int maxIter_) {
    if (!current_.ctx) return 0;
    sphk_ctx* ctx = current_.ctx;
    const int num = static_cast<int>(fluids->size());
    auto totalError = std::numeric_limits<float>::max();
    auto iter = 0;
    // warm stiffness of the previous step follows its particle (:170-171)
    check(sphk_permute(ctx, denWarmStiff.addr(), 1, num), "sphk_permute");
    // (ghost particles carry their warm stiffness with them: it travels with the candidates and is permuted with the rest)
    check(sphk_dfsph_den_correct(ctx, &current_.abi, denWarmStiff.addr(), dt), "sphk_dfsph_den_correct");
    produced(1, current_.abi.fluid.vel, 3);
    check(sphk_dfsph_den_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0, nullptr),
          "sphk_dfsph_den_error");
    produced(2, bufferFloat.addr(), 1);
    check(sphk_copy(ctx, denWarmStiff.addr(), bufferFloat.addr(), num), "sphk_copy");     // :185
    const float thresholdCount = static_cast<float>(globalCount_ >= 0 ? globalCount_ : num);
    if (errorThreshold >= 0.0f && deviceLoops_ && !fieldHook_) {
        check(sphk_loop_begin(ctx, 1, 2, maxIter_, errorThreshold * num * rho0, 2), "sphk_loop_begin");
        for (int k = 0; k < maxIter_; ++k) {
            check(sphk_dfsph_den_correct(ctx, &current_.abi, bufferFloat.addr(), dt), "sphk_dfsph_den_correct");
            check(sphk_dfsph_den_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0,
                                       denWarmStiff.addr()),
                  "sphk_dfsph_den_error");                                                // + :199-203 fused
            check(sphk_loop_next(ctx, 1, error.addr(), num), "sphk_loop_next");
        }
        check(sphk_loop_end(ctx, 1), "sphk_loop_end");
        return -1;
    }
    while ((iter < 2 || totalError > errorThreshold * thresholdCount * rho0) && iter < maxIter_) {
        check(sphk_dfsph_den_correct(ctx, &current_.abi, bufferFloat.addr(), dt), "sphk_dfsph_den_correct");
        produced(1, current_.abi.fluid.vel, 3);
        check(sphk_dfsph_den_error(ctx, &current_.abi, alpha.addr(), error.addr(), bufferFloat.addr(), dt, rho0,
                                   denWarmStiff.addr()),
              "sphk_dfsph_den_error");                                                    // + :199-203 fused
        produced(2, bufferFloat.addr(), 1);
        ++iter;
        if (iter >= 2 && errorThreshold >= 0.0f) totalError = reduceError(num);
    }
    return iter;
}

// thrust::reduce(error, abs_plus), DFSPHSolver.cu:206,360 -- over the particles this rank owns, then over the ranks
float DFSPHSolver::reduceError(int num) {
    float local = 0.0f;
    if (reduceHook_) {
        int begin = 0, count = num;
        if (ownedRange_) ownedRange_(begin, count);
        check(sphk_reduce_abs_sum(current_.ctx, error.addr(begin), count, &local), "sphk_reduce_abs_sum");
        return static_cast<float>(reduceHook_(static_cast<double>(local)));
    }
    check(sphk_reduce_abs_sum(current_.ctx, error.addr(), num, &local), "sphk_reduce_abs_sum");
    return local;
}

int DFSPHSolver::loopIterations(int slot, int hostCount) const {
    if (hostCount >= 0 || !current_.ctx) return hostCount;
    int it = 0;
    check(sphk_loop_iterations(current_.ctx, slot, &it, nullptr), "sphk_loop_iterations");
    return it;
}

// ================================================================================================
// PBDSolver (PBDSolver.cu)
// ================================================================================================
// PBDSolver::step, PBDSolver.cu:34-73
void PBDSolver::step(std::shared_ptr<SPHParticles>& fluids, const std::shared_ptr<SPHParticles>& boundaries,
                     const DArray<int>& cellStartFluid, const DArray<int>& cellStartBoundary, float3 spaceSize,
                     int3 cellSize, float cellLength, float radius, float dt, float rho0, float rhoB, float, float,
                     float3 G, float surfaceTensionIntensity, float airPressure) {
    if (!posLastInitialized) {
        if (fluids->engine() && fluids->engine()->ok()) sphk_synchronize(fluids->engine()->ctx());
        initializePosLast(fluids->getPos());
        throw "PBD: The last position of fluids is initialized.";     // Q6, PBDSolver.cu:44-47
    }
    // positions move inside the step (Q7): the neighbour list carries a skin of 0.15 R; the engine tracks the
    // displacement on the device and falls back to the cell walk if the corrections ever exceed skin/2
    if (!beginStep(fluids, boundaries, cellStartFluid, cellStartBoundary, radius, true, 150)) return;
    updateNeighborhood(fluids);
    project(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, cellSize, spaceSize, cellLength, radius, maxIter);
    check(sphk_pbd_velocity_from_positions(current_.ctx, &current_.abi, reinterpret_cast<float*>(fluidPosLast.addr()), dt),
          "sphk_pbd_velocity_from_positions");
    const bool surface = surfaceTensionIntensity > EPSILON || airPressure > EPSILON;
    if (fusedSweeps_ && surface) {
        // XSPH + colour gradient in one sweep (positions are final after the projection), then the surface sweep
        float* cg = colorGradBuffer();
        check(sphk_fused_pbd_xsph_color_grad(current_.ctx, &current_.abi, xSPH_c, rho0, cg, rhoB), "sphk_fused_pbd_xsph_color_grad");
        producedVel(fluids); produced(0, cg, 3);
        check(sphk_surface(current_.ctx, &current_.abi, cg, dt, rho0, surfaceTensionIntensity, airPressure), "sphk_surface");
        producedVel(fluids);
    } else {
        diffuse(fluids, cellStartFluid, cellSize, cellLength, rho0, radius, xSPH_c);
        producedVel(fluids);
        if (surface)
            handleSurface(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, rhoB, cellSize, cellLength, radius,
                          dt, surfaceTensionIntensity, airPressure);
    }
    force(fluids, dt, G);
    predict(fluids, dt, spaceSize);
}

// :75-79
void PBDSolver::predict(std::shared_ptr<SPHParticles>& fluids, float dt, float3 spaceSize) {
    check(sphk_copy(current_.ctx, reinterpret_cast<float*>(fluidPosLast.addr()),
                    reinterpret_cast<const float*>(fluids->getPosPtr()), 3 * static_cast<int>(fluids->size())),
          "sphk_copy");
    advect(fluids, dt, spaceSize);
}

// :81-87
void PBDSolver::updateNeighborhood(const std::shared_ptr<SPHParticles>& particles) {
    check(sphk_permute(current_.ctx, reinterpret_cast<float*>(fluidPosLast.addr()), 3, static_cast<int>(particles->size())),
          "sphk_permute");
}

// :117-125
void PBDSolver::diffuse(std::shared_ptr<SPHParticles>&, const DArray<int>&, int3, float, float rho0, float, float visc) {
    check(sphk_pbd_xsph(current_.ctx, &current_.abi, visc, rho0), "sphk_pbd_xsph");
}

// :225-258
int PBDSolver::project(std::shared_ptr<SPHParticles>&, const std::shared_ptr<SPHParticles>&, const DArray<int>&,
                       const DArray<int>&, float rho0, int3, float3 spaceSize, float, float, int maxIter_) {
    const float sp[3] = {spaceSize.x, spaceSize.y, spaceSize.z};
    auto iter = 0;
    while (iter < maxIter_) {
        check(sphk_pbd_density_lambda(current_.ctx, &current_.abi, bufferFloat.addr(), rho0, relaxation),
              "sphk_pbd_density_lambda");
        produced(2, bufferFloat.addr(), 1);
        check(sphk_pbd_delta_pos_apply(current_.ctx, &current_.abi, bufferFloat.addr(),
                                       reinterpret_cast<float*>(bufferFloat3.addr()), rho0, sp),
              "sphk_pbd_delta_pos_apply");
        produced(4, current_.abi.fluid.pos, 3);
        ++iter;
    }
    return iter;
}

// ================================================================================================
// SPHSystem (SPHSystem.cu)
// ================================================================================================
SPHSystem::SPHSystem(std::shared_ptr<SPHParticles>& fluidParticles, std::shared_ptr<SPHParticles>& boundaryParticles,
                     std::shared_ptr<BaseSolver>& solver, const float3 spaceSize, const float sphCellLength,
                     const float sphSmoothingRadius, const float dt, const float sphM0, const float sphRho0,
                     const float sphRhoBoundary, const float sphStiff, const float sphVisc,
                     const float sphSurfaceTensionIntensity, const float sphAirPressure, const float3 sphG,
                     const int3 cellSize)
    // the caller's three shared_ptrs are moved from, as in the reference (SPHSystem.cu:50-51)
    : _fluids(std::move(fluidParticles)), _boundaries(std::move(boundaryParticles)), _solver(std::move(solver)),
      cellStartFluid(cellSize.x * cellSize.y * cellSize.z + 1), cellStartBoundary(cellSize.x * cellSize.y * cellSize.z + 1),
      _spaceSize(spaceSize), _sphSmoothingRadius(sphSmoothingRadius), _sphCellLength(sphCellLength), _dt(dt),
      _sphRho0(sphRho0), _sphRhoBoundary(sphRhoBoundary), _sphStiff(sphStiff), _sphG(sphG), _sphVisc(sphVisc),
      _sphSurfaceTensionIntensity(sphSurfaceTensionIntensity), _sphAirPressure(sphAirPressure), _cellSize(cellSize) {
    _engine = std::make_shared<sphb200::Engine>(fluidSize(), boundarySize(), cellSize, sphCellLength);
    _fluids->bindEngine(_engine);
    _boundaries->bindEngine(_engine);
    CUDA_CALL(cudaEventCreate(&_evStart));
    CUDA_CALL(cudaEventCreate(&_evStop));
    if (const char* g = std::getenv("SPHK_STEP_GRAPH")) _graphEnabled = g[0] != '0';
    if (!_engine->ok()) return;
    // SPHSystem.cu:68-76
    neighborSearch(_boundaries, cellStartBoundary);
    computeBoundaryMass();
    check(sphk_fill(_engine->ctx(), _fluids->getMassPtr(), fluidSize(), sphM0), "sphk_fill");
    neighborSearch(_fluids, cellStartFluid);
    step();
}

SPHSystem::~SPHSystem() noexcept {
    if (_engine && _engine->ok()) sphk_synchronize(_engine->ctx());
    if (_graphExec) cudaGraphExecDestroy(_graphExec);
    if (_evStart) cudaEventDestroy(_evStart);
    if (_evStop) cudaEventDestroy(_evStop);
}

// SPHSystem.cu:107-112
void SPHSystem::computeBoundaryMass() {
    const sphk_particles b = _boundaries->abi();
    check(sphk_boundary_mass(_engine->ctx(), &b, cellStartBoundary.addr(), _sphRhoBoundary, _sphSmoothingRadius),
          "sphk_boundary_mass");
}

// SPHSystem.cu:114-127
void SPHSystem::neighborSearch(const std::shared_ptr<SPHParticles>& particles, DArray<int>& cellStart) {
    const sphk_particles p = particles->abi();
    const int which = (particles.get() == _boundaries.get()) ? 1 : 0;
    check(sphk_neighbor_search(_engine->ctx(), which, &p, cellStart.addr()), "sphk_neighbor_search");
    // the fluid search re-packs the records from the API arrays: edits made behind the engine's back
    // (Particles::advect, raw pointer writes) are absorbed here, no separate sphk_refresh is needed
    if (which == 0) _engine->shadowsStale = false;
}

// SPHSystem.cu:129-158
bool SPHSystem::solverIsGraphSafe() const {
    const auto* known = dynamic_cast<const BasicSPHSolver*>(_solver.get());   // unknown BaseSolver subclasses: never
    return known && known->stepIsGraphSafe();
}

unsigned int SPHSystem::solverConfigEpoch() const {
    const auto* known = dynamic_cast<const BasicSPHSolver*>(_solver.get());
    return known ? known->configEpoch() : 0u;
}

void SPHSystem::dropStepGraph() {
    if (!_graphExec) return;
    sphk_synchronize(_engine->ctx());
    cudaGraphExecDestroy(_graphExec);
    _graphExec = nullptr;
}

void SPHSystem::setStepGraph(bool on) {
    _graphEnabled = on;
    if (!on) dropStepGraph();
}

// SPHSystem.cu:129-158.  Small scenes are launch-bound (~45 kernels of a few microseconds per DFSPH step): once
// two plain steps have run (all lazy allocations done), a fixed-iteration step is captured into a CUDA graph and
// replayed -- same kernels, same order, same arguments (all scene parameters are const members).
float SPHSystem::step() {
    if (!_engine->ok()) return 0.0f;
    cudaStream_t st = _engine->stream();
    sphk_ctx* ctx = _engine->ctx();
    CUDA_CALL(cudaEventRecord(_evStart, st));
    // a captured step replays a fixed kernel sequence: drop it when the solver's settings changed or it stopped being
    // replayable.  (Particles moved behind the engine's back need nothing special: the replayed step begins with the
    // fluid neighbour search, which re-packs the records from the API arrays.)
    if (_graphExec && (_graphConfigEpoch != solverConfigEpoch() || !solverIsGraphSafe())) dropStepGraph();
    if (_graphExec) {
        _engine->shadowsStale = false;
        CUDA_CALL(cudaGraphLaunch(_graphExec, st));
        sphk_add_launches(ctx, _graphLaunches);
        check(sphk_synchronize(ctx), "step");
    } else {
        const bool capture = _graphEnabled && _plainSteps >= 2 && solverIsGraphSafe();
        const long long launches0 = sphk_launch_count(ctx);
        bool captured = false;
        if (capture) captured = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        bool ok = false;
        try {
            neighborSearch(_fluids, cellStartFluid);
            _solver->step(_fluids, _boundaries, cellStartFluid, cellStartBoundary, _spaceSize, _cellSize, _sphCellLength,
                          _sphSmoothingRadius, _dt, _sphRho0, _sphRhoBoundary, _sphStiff, _sphVisc, _sphG,
                          _sphSurfaceTensionIntensity, _sphAirPressure);
            ok = true;
        } catch (const char* s) {
            std::cout << s << "\n";
        } catch (...) {
            std::cout << "Unknown Exception at " << __FILE__ << ": line " << __LINE__ << "\n";
        }
        if (captured) {
            cudaGraph_t graph = nullptr;
            if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph && ok &&
                cudaGraphInstantiate(&_graphExec, graph, 0) == cudaSuccess) {
                _graphLaunches = sphk_launch_count(ctx) - launches0;
                _graphConfigEpoch = solverConfigEpoch();
                CUDA_CALL(cudaGraphLaunch(_graphExec, st));     // the captured step has not run yet: run it now
            } else {
                cudaGetLastError();
                _graphExec = nullptr;
                _graphEnabled = false;                          // capture refused (e.g. a library call): stay on plain launches
                printf("SPHSystem: step graph capture failed, continuing with plain launches\n");
                neighborSearch(_fluids, cellStartFluid);
                try {
                    _solver->step(_fluids, _boundaries, cellStartFluid, cellStartBoundary, _spaceSize, _cellSize,
                                  _sphCellLength, _sphSmoothingRadius, _dt, _sphRho0, _sphRhoBoundary, _sphStiff, _sphVisc,
                                  _sphG, _sphSurfaceTensionIntensity, _sphAirPressure);
                } catch (const char* s) {
                    std::cout << s << "\n";
                } catch (...) {
                    std::cout << "Unknown Exception at " << __FILE__ << ": line " << __LINE__ << "\n";
                }
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (ok) ++_plainSteps;
        check(sphk_synchronize(ctx), "step");
    }
    float milliseconds = 0.0f;
    CUDA_CALL(cudaEventRecord(_evStop, st));
    CUDA_CALL(cudaEventSynchronize(_evStop));
    CUDA_CALL(cudaEventElapsedTime(&milliseconds, _evStart, _evStop));
    return milliseconds;
}

// vbo.cu:46-51
extern "C" void generate_dots(float3* dot, float3* color, const std::shared_ptr<SPHParticles> particles) {
    const auto& eng = particles->engine();
    if (!eng || !eng->ok()) { printf("generate_dots: particles are not bound to a B200 engine\n"); return; }
    const sphk_particles p = particles->abi();
    check(sphk_export_dots(eng->ctx(), &p, reinterpret_cast<float*>(dot), reinterpret_cast<float*>(color)), "sphk_export_dots");
}

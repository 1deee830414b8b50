// sph_app.cpp -- headless facade; see include/sph_app.h.
//
// Restates the reference's API call sites (main.cpp:86,117 particle construction; :119-130 solver
// construction; :131-134 SPHSystem construction; :302 step; vbo.cu:48 accessor use) and nothing
// else.  It is deliberately written ONLY against names the reference headers export, so that it
// builds unchanged on top of either engine (this repo's headers, or /root/reference/src).
#include <cstdio>
#include <iostream>
#include <vector>
#include <memory>
#include <cuda_runtime.h>
#ifdef SPH_APP_REFERENCE_ENGINE
#include <helper_math.h>
#include "global.h"
#endif
#include "DArray.h"
#include "Particles.h"
#include "SPHParticles.h"
#include "BaseSolver.h"
#include "BasicSPHSolver.h"
#include "DFSPHSolver.h"
#include "PBDSolver.h"
#include "SPHSystem.h"
#include "sph_app.h"

struct sph_app {
    std::shared_ptr<SPHSystem> system;
};

static std::vector<float3> to_float3(const float* xyz, int n) {
    std::vector<float3> v(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) v[i] = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return v;
}

extern "C" sph_app* sph_app_create(const float* fluid_xyz, int n_fluid,
                                   const float* boundary_xyz, int n_boundary,
                                   const sph_app_params* p) {
    if (!fluid_xyz || !boundary_xyz || !p || n_fluid <= 0 || n_boundary <= 0) return nullptr;
    auto fluidParticles = std::make_shared<SPHParticles>(to_float3(fluid_xyz, n_fluid));
    auto boundaryParticles = std::make_shared<SPHParticles>(to_float3(boundary_xyz, n_boundary));
    std::shared_ptr<BaseSolver> pSolver;
    const int num = static_cast<int>(fluidParticles->size());
    switch (p->solver) {
    case 2:
        if (p->max_iter > 0) pSolver = std::make_shared<PBDSolver>(num, p->max_iter);
        else pSolver = std::make_shared<PBDSolver>(num);
        break;
    case 1:
        if (p->max_iter > 0)
            pSolver = std::make_shared<DFSPHSolver>(num, p->density_error_threshold,
                                                    p->divergence_error_threshold, p->max_iter);
        else pSolver = std::make_shared<DFSPHSolver>(num);
        break;
    default:
        pSolver = std::make_shared<BasicSPHSolver>(num);
        break;
    }
    auto* app = new sph_app;
    app->system = std::make_shared<SPHSystem>(
        fluidParticles, boundaryParticles, pSolver,
        make_float3(p->space[0], p->space[1], p->space[2]), p->cell_length, p->radius, p->dt, p->m0,
        p->rho0, p->rho_boundary, p->stiff, p->visc, p->surface_tension, p->air_pressure,
        make_float3(p->gravity[0], p->gravity[1], p->gravity[2]),
        make_int3(p->cell_size[0], p->cell_size[1], p->cell_size[2]));
    if (cudaDeviceSynchronize() != cudaSuccess) {
        std::fprintf(stderr, "sph_app_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete app;
        return nullptr;
    }
    return app;
}

extern "C" void sph_app_destroy(sph_app* app) { delete app; }

extern "C" float sph_app_step(sph_app* app) { return app->system->step(); }

extern "C" int sph_app_fluid_size(const sph_app* app) { return app->system->fluidSize(); }
extern "C" int sph_app_boundary_size(const sph_app* app) { return app->system->boundarySize(); }

static int d2h(void* dst, const void* src, size_t bytes) {
    if (!dst) return 0;
    return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}

extern "C" int sph_app_download_fluid(sph_app* app, float* pos, float* vel, float* density,
                                      float* pressure, float* mass, int* p2c) {
    const auto f = app->system->getFluids();
    const size_t n = f->size();
    int bad = 0;
    bad |= d2h(pos, f->getPosPtr(), n * sizeof(float3));
    bad |= d2h(vel, f->getVelPtr(), n * sizeof(float3));
    bad |= d2h(density, f->getDensityPtr(), n * sizeof(float));
    bad |= d2h(pressure, f->getPressurePtr(), n * sizeof(float));
    bad |= d2h(mass, f->getMassPtr(), n * sizeof(float));
    bad |= d2h(p2c, f->getParticle2Cell(), n * sizeof(int));
    return bad;
}

extern "C" int sph_app_download_boundary(sph_app* app, float* pos, float* mass, int* p2c) {
    const auto b = app->system->getBoundaries();
    const size_t n = b->size();
    int bad = 0;
    bad |= d2h(pos, b->getPosPtr(), n * sizeof(float3));
    bad |= d2h(mass, b->getMassPtr(), n * sizeof(float));
    bad |= d2h(p2c, b->getParticle2Cell(), n * sizeof(int));
    return bad;
}

extern "C" int sph_app_upload_fluid(sph_app* app, const float* pos, const float* vel) {
    const auto f = app->system->getFluids();
    const size_t n = f->size();
    int bad = 0;
    if (pos) bad |= cudaMemcpy(f->getPosPtr(), pos, n * sizeof(float3), cudaMemcpyHostToDevice) != cudaSuccess;
    if (vel) bad |= cudaMemcpy(f->getVelPtr(), vel, n * sizeof(float3), cudaMemcpyHostToDevice) != cudaSuccess;
    return bad;
}

extern "C" const char* sph_app_engine(void) {
#ifdef SPH_APP_REFERENCE_ENGINE
    return "reference-cuda";
#else
    return "b200-native";
#endif
}

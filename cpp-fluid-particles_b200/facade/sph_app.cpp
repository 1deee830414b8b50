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

// One batch in flight through the pipelined host-buffer path (sph_app_submit / sph_app_wait)
struct sph_batch {
    float3* dIn[2] = {nullptr, nullptr};     // device staging: pos, vel of the batch
    float3* dOut[2] = {nullptr, nullptr};    // device staging of the result: pos, vel
    float* dOutDensity = nullptr;
    cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
    float* hPos = nullptr; float* hVel = nullptr; float* hDensity = nullptr;   // where the result goes (host)
    bool pending = false;                    // uploaded, not yet stepped
    float ms = 0.f;
};

struct sph_app {
    std::shared_ptr<SPHSystem> system;
    BaseSolver* solver = nullptr;            // owned by `system` (its constructor moves the caller's shared_ptr)
    // pipelined path (created on first sph_app_submit): uploads and downloads on their own streams (PCIe is full duplex)
    cudaStream_t copyStream = nullptr, downStream = nullptr;
    sph_batch batch[2];
    unsigned long long submitted = 0;
    bool pipeReady = false;
    ~sph_app() {
        if (pipeReady) {
            cudaStreamSynchronize(copyStream);
            cudaStreamSynchronize(downStream);
            for (auto& b : batch) {
                cudaFree(b.dIn[0]); cudaFree(b.dIn[1]); cudaFree(b.dOut[0]); cudaFree(b.dOut[1]); cudaFree(b.dOutDensity);
                cudaEventDestroy(b.uploaded); cudaEventDestroy(b.computed); cudaEventDestroy(b.downloaded);
            }
            cudaStreamDestroy(copyStream);
            cudaStreamDestroy(downStream);
        }
    }
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
    app->solver = pSolver.get();
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

// ---- pipelined host-buffer stepping -------------------------------------------------------------------------------
// The reference's own call sites move data with synchronous cudaMemcpy through the raw accessors (Particles.h:24,
// vbo.cu:48).  A caller that feeds a batch of particle states from host memory and reads every result back pays
// ~2 ms of PCIe time per 2M-particle step that way.  sph_app_submit overlaps those copies with the previous batch's
// step: uploads and downloads run on a copy stream into / out of device staging buffers, the step itself only adds
// two device-to-device copies.  Only public accessors and the CUDA runtime are used, so this builds on either engine.
static bool pipe_init(sph_app* app) {
    if (app->pipeReady) return true;
    const size_t n = app->system->getFluids()->size();
    if (cudaStreamCreateWithFlags(&app->copyStream, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamCreateWithFlags(&app->downStream, cudaStreamNonBlocking) != cudaSuccess) return false;
    for (auto& b : app->batch) {
        bool ok = true;
        for (int k = 0; k < 2; ++k) {
            ok = ok && cudaMalloc(reinterpret_cast<void**>(&b.dIn[k]), n * sizeof(float3)) == cudaSuccess;
            ok = ok && cudaMalloc(reinterpret_cast<void**>(&b.dOut[k]), n * sizeof(float3)) == cudaSuccess;
        }
        ok = ok && cudaMalloc(reinterpret_cast<void**>(&b.dOutDensity), n * sizeof(float)) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&b.computed, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&b.downloaded, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventRecord(b.downloaded, app->downStream) == cudaSuccess;      // "nothing to wait for" before the first use
        if (!ok) return false;
    }
    app->pipeReady = true;
    return true;
}

// steps the batch in `slot` (its upload has been enqueued) and enqueues the download of its result
static int pipe_run(sph_app* app, int slot) {
    sph_batch& b = app->batch[slot];
    if (!b.pending) return 0;
    const auto f = app->system->getFluids();
    const size_t n = f->size();
    int bad = 0;
    // legacy default stream: ordered with everything the engines enqueue (the reference uses it throughout; this
    // repository's engine stream is a blocking stream)
    bad |= cudaStreamWaitEvent(nullptr, b.uploaded, 0) != cudaSuccess;
    bad |= cudaMemcpyAsync(f->getPosPtr(), b.dIn[0], n * sizeof(float3), cudaMemcpyDeviceToDevice, nullptr) != cudaSuccess;
    bad |= cudaMemcpyAsync(f->getVelPtr(), b.dIn[1], n * sizeof(float3), cudaMemcpyDeviceToDevice, nullptr) != cudaSuccess;
    b.ms = app->system->step();
    bad |= cudaStreamWaitEvent(nullptr, b.downloaded, 0) != cudaSuccess;     // this slot's previous result has left the staging buffers
    bad |= cudaMemcpyAsync(b.dOut[0], f->getPosPtr(), n * sizeof(float3), cudaMemcpyDeviceToDevice, nullptr) != cudaSuccess;
    bad |= cudaMemcpyAsync(b.dOut[1], f->getVelPtr(), n * sizeof(float3), cudaMemcpyDeviceToDevice, nullptr) != cudaSuccess;
    bad |= cudaMemcpyAsync(b.dOutDensity, f->getDensityPtr(), n * sizeof(float), cudaMemcpyDeviceToDevice, nullptr) != cudaSuccess;
    bad |= cudaEventRecord(b.computed, nullptr) != cudaSuccess;
    bad |= cudaStreamWaitEvent(app->downStream, b.computed, 0) != cudaSuccess;
    if (b.hPos) bad |= cudaMemcpyAsync(b.hPos, b.dOut[0], n * sizeof(float3), cudaMemcpyDeviceToHost, app->downStream) != cudaSuccess;
    if (b.hVel) bad |= cudaMemcpyAsync(b.hVel, b.dOut[1], n * sizeof(float3), cudaMemcpyDeviceToHost, app->downStream) != cudaSuccess;
    if (b.hDensity) bad |= cudaMemcpyAsync(b.hDensity, b.dOutDensity, n * sizeof(float), cudaMemcpyDeviceToHost, app->downStream) != cudaSuccess;
    bad |= cudaEventRecord(b.downloaded, app->downStream) != cudaSuccess;
    b.pending = false;
    return bad;
}

extern "C" int sph_app_submit(sph_app* app, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                              float* density_out) {
    if (!app || !pos_in || !vel_in || !pipe_init(app)) return 1;
    const size_t n = app->system->getFluids()->size();
    const int slot = static_cast<int>(app->submitted & 1ull);
    sph_batch& b = app->batch[slot];
    int bad = 0;
    // (the slot's previous batch was stepped during the previous submit, so its input staging is free; its download reads
    // the output staging, which this batch overwrites only after ITS step and after waiting for the `downloaded` event)
    bad |= cudaMemcpyAsync(b.dIn[0], pos_in, n * sizeof(float3), cudaMemcpyHostToDevice, app->copyStream) != cudaSuccess;
    bad |= cudaMemcpyAsync(b.dIn[1], vel_in, n * sizeof(float3), cudaMemcpyHostToDevice, app->copyStream) != cudaSuccess;
    bad |= cudaEventRecord(b.uploaded, app->copyStream) != cudaSuccess;
    b.hPos = pos_out; b.hVel = vel_out; b.hDensity = density_out;
    b.pending = true;
    app->submitted++;
    // while this upload is in flight, step the batch submitted before it
    bad |= pipe_run(app, slot ^ 1);
    return bad;
}

extern "C" int sph_app_wait(sph_app* app) {
    if (!app || !app->pipeReady) return 0;
    int bad = 0;
    const int last = static_cast<int>((app->submitted + 1) & 1ull);     // slot of the most recent submit
    bad |= pipe_run(app, last ^ 1);
    bad |= pipe_run(app, last);
    bad |= cudaStreamSynchronize(app->copyStream) != cudaSuccess;
    bad |= cudaStreamSynchronize(app->downStream) != cudaSuccess;
    return bad;
}

// ---- introspection of this repository's engine (absent from the reference build of the facade) -------------------------
extern "C" int sph_app_dfsph_iterations(sph_app* app, int* divergence_iters, int* density_iters) {
#ifdef SPH_APP_REFERENCE_ENGINE
    (void)app; (void)divergence_iters; (void)density_iters;
    return 1;
#else
    auto* d = dynamic_cast<DFSPHSolver*>(app->solver);
    if (!d) return 1;
    if (divergence_iters) *divergence_iters = d->lastDivergenceIterations();
    if (density_iters) *density_iters = d->lastDensityIterations();
    return 0;
#endif
}

extern "C" int sph_app_set_option(sph_app* app, int option, int value) {
#ifdef SPH_APP_REFERENCE_ENGINE
    (void)app; (void)option; (void)value;
    return 1;
#else
    switch (option) {
    case 1: { auto* d = dynamic_cast<DFSPHSolver*>(app->solver); if (!d) return 1; d->setDeviceLoops(value != 0); return 0; }
    case 2: { auto* b = dynamic_cast<BasicSPHSolver*>(app->solver); if (!b) return 1; b->setFusedSweeps(value != 0); return 0; }
    case 3: app->system->setStepGraph(value != 0); return 0;
    default: return 1;
    }
#endif
}

extern "C" const char* sph_app_engine(void) {
#ifdef SPH_APP_REFERENCE_ENGINE
    return "reference-cuda";
#else
    return "b200-native";
#endif
}

/* sph_app.h -- headless C facade over the reference-shaped C++ class API
 * (SPHSystem / BaseSolver / SPHParticles / DArray).
 *
 * The facade source (cpp-fluid-particles_b200/facade/sph_app.cpp) restates the ONLY call sites of
 * that API in the reference -- initSPHSystem() /root/reference/src/main.cpp:117-134 (particle
 * upload, solver choice, 16-argument SPHSystem constructor) and oneStep() main.cpp:300-306
 * (`float ms = pSystem->step()`) -- without GLUT/OpenGL.  The very same source file is compiled
 *   (1) by g++ against this repository's headers  -> libsphhost.so   (product: B200-native engine)
 *   (2) by nvcc against /root/reference/src headers + the unmodified reference .cu files
 *                                                 -> oracle/_ref/libsphref.so (parity oracle, B1)
 * which is the drop-in claim made concrete: identical call sites, two engines underneath.
 */
#ifndef SPH_APP_H_
#define SPH_APP_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sph_app sph_app;

typedef struct sph_app_params {
    float space[3];        /* spaceSize                       main.cpp:54  */
    float cell_length;     /* sphCellLength                   main.cpp:57  */
    float radius;          /* sphSmoothingRadius              main.cpp:56  */
    float dt;              /*                                 main.cpp:58  */
    float m0;              /* sphM0                           main.cpp:61  */
    float rho0;            /* sphRho0                         main.cpp:59  */
    float rho_boundary;    /* sphRhoBoundary                  main.cpp:60  */
    float stiff;           /* sphStiff                        main.cpp:62  */
    float visc;            /* sphVisc                         main.cpp:64  */
    float surface_tension; /* sphSurfaceTensionIntensity      main.cpp:65  */
    float air_pressure;    /* sphAirPressure                  main.cpp:66  */
    float gravity[3];      /* sphG                            main.cpp:63  */
    int   cell_size[3];    /* cellSize                        main.cpp:67  */
    int   solver;          /* 0 = BasicSPHSolver, 1 = DFSPHSolver, 2 = PBDSolver (main.cpp:69-71) */
    int   max_iter;        /* <=0: solver constructor default (DFSPHSolver.h:30 / PBDSolver.h:28) */
    float density_error_threshold;    /* DFSPH only; used when max_iter > 0 */
    float divergence_error_threshold; /* DFSPH only; used when max_iter > 0 */
} sph_app_params;

/* Builds SPHParticles for fluid and boundary from host xyz triplets, the chosen solver and the
 * SPHSystem (whose constructor already performs one step, SPHSystem.cu:76).  NULL on failure. */
sph_app* sph_app_create(const float* fluid_xyz, int n_fluid,
                        const float* boundary_xyz, int n_boundary,
                        const sph_app_params* params);
void  sph_app_destroy(sph_app* app);
/* One SPHSystem::step(); returns the milliseconds the system itself reports (SPHSystem.cu:151-157). */
float sph_app_step(sph_app* app);
int   sph_app_fluid_size(const sph_app* app);
int   sph_app_boundary_size(const sph_app* app);
/* Device->host copies through the public accessors (getPosPtr, getVelPtr, getDensityPtr,
 * getPressurePtr, getMassPtr, getParticle2Cell).  Any pointer may be NULL.  Returns 0 on success. */
int   sph_app_download_fluid(sph_app* app, float* pos_xyz, float* vel_xyz, float* density,
                             float* pressure, float* mass, int* particle2cell);
int   sph_app_download_boundary(sph_app* app, float* pos_xyz, float* mass, int* particle2cell);
/* Host->device overwrite of fluid pos / vel through getPosPtr / getVelPtr (either may be NULL). */
int   sph_app_upload_fluid(sph_app* app, const float* pos_xyz, const float* vel_xyz);
/* Pipelined host-buffer stepping: one call = "advance the particle state held in HOST memory (pos_in, vel_in) by one
 * SPHSystem::step() and deliver pos / vel / density of the result to HOST memory".  The upload of this batch and the
 * download of the previous result overlap the previous batch's step (copy stream + device staging; the step itself
 * only adds device-to-device copies), so a stream of batches costs one step per batch instead of step + PCIe time.
 * Host buffers should be pinned (cudaHostAlloc / torch pin_memory) and must stay valid until sph_app_wait() returns;
 * results of batch k are complete after the call that submits batch k+2, or after sph_app_wait().  Output pointers
 * may be NULL.  Returns 0 on success. */
int   sph_app_submit(sph_app* app, const float* pos_in_xyz, const float* vel_in_xyz,
                     float* pos_out_xyz, float* vel_out_xyz, float* density_out);
/* Steps every submitted batch that has not run yet and waits for all downloads. */
int   sph_app_wait(sph_app* app);
/* Introspection / switches of this repository's engine (return 1 on the reference build of the facade):
 * iteration counts of the last DFSPH step; option 1 = device-side loop tests for adaptive DFSPH (default on; 0 = the
 * reference's host loop with one error sum read back per iteration), 2 = fused sweeps (default on), 3 = step graph. */
int   sph_app_dfsph_iterations(sph_app* app, int* divergence_iters, int* density_iters);
int   sph_app_set_option(sph_app* app, int option, int value);
/* Name of the engine underneath: "reference-cuda" or "b200-native". */
const char* sph_app_engine(void);

#ifdef __cplusplus
}
#endif
#endif /* SPH_APP_H_ */

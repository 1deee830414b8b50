/* sphk.h -- C-ABI of the B200-native SPH kernel library (libsphk.so, sm_100a).
 *
 * The reference (zhai-xiao/CPP-Fluid-Particles) has no FFI/plugin boundary: its boundary is the C++
 * class API SPHSystem / BaseSolver / SPHParticles / DArray compiled into one executable.  This
 * header is the boundary a maintainer binds to: ONE entry point per reference launch site on the
 * hot path SPHSystem::step() (each declaration cites the reference code it replaces, paths relative
 * to /root/reference/src).  The reference-shaped C++ classes in cpp-fluid-particles_b200/host/ are
 * written purely on top of these calls.
 *
 * Conventions
 *  - Plain C: raw DEVICE pointers, ints, floats, PODs.  No torch / thrust / C++ types.
 *  - "float3 arrays" are packed xyz triplets, 12-byte stride, exactly the reference's float3 DArrays.
 *  - Every function returns 0 on success or a negative SPHK_ERR_* / positive cudaError_t value, and
 *    never throws.  Nothing synchronises the host unless stated (the reference syncs on every Thrust
 *    call; here only sphk_reduce_abs_sum and sphk_synchronize do).
 *  - All work is enqueued on the stream given to sphk_create (a cudaStream_t passed as void*).
 *  - There is NO CPU fallback: without a CUDA device sphk_create fails with SPHK_ERR_NO_DEVICE.
 *  - The library never allocates or frees API-visible arrays; sphk_ctx owns only scratch (sort
 *    buffers, packed 32-byte particle records, neighbour lists).
 *
 * Record coherence rule: sphk_neighbor_search() packs the particle set into 32-byte records
 * {x,y,z,s | vx,vy,vz,m} that the sweep kernels gather; every sphk_* call that changes pos / vel writes
 * both the API array and the record.  If the caller edits pos / vel itself between calls it must call
 * sphk_neighbor_search (or sphk_refresh) before the next sweep.  A scalar array passed to a sweep that
 * gathers it from neighbours (stiffness, lambda) is mirrored into the records unless the previous sphk_*
 * call produced it; arrays modified behind the library's back need sphk_refresh.
 */
#ifndef SPHK_H_
#define SPHK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SPHK_VERSION 1

enum {
    SPHK_OK = 0,
    SPHK_ERR_INVALID = -1,    /* bad argument */
    SPHK_ERR_NO_DEVICE = -2,  /* no CUDA device / driver: there is no CPU path */
    SPHK_ERR_CAPACITY = -3,   /* n exceeds the capacity given to sphk_create */
    SPHK_ERR_STATE = -4,      /* call order violated (e.g. sweep before neighbour search) */
    SPHK_ERR_ALLOC = -5,
    SPHK_ERR_COMM = -6        /* multi-GPU: NCCL missing / failed, or a neighbour's message did not arrive */
};

typedef struct sphk_ctx sphk_ctx;

/* Uniform grid, CUDAFunctions.cuh:64-78: cell (x,y,z) -> (x*cs.y + y)*cs.z + z, z fastest. */
typedef struct sphk_grid {
    int   cell_size[3];
    float cell_length;
    int   origin[3];   /* cell coordinates of this context's cell (0,0,0) in the global grid: all zero for a
                          single-GPU system; a slab rank covers global x-planes [origin[0], origin[0]+cell_size[0]) */
} sphk_grid;

/* One SPHParticles object (SPHParticles.h:56-59 + Particles.h:47-48): device pointers. */
typedef struct sphk_particles {
    float* pos;            /* float3[n]   getPosPtr()          */
    float* vel;            /* float3[n]   getVelPtr()  (NULL allowed for boundaries) */
    float* mass;           /* float[n]    getMassPtr()         */
    float* density;        /* float[n]    getDensityPtr()      */
    float* pressure;       /* float[n]    getPressurePtr()     */
    int*   particle2cell;  /* int[n]      getParticle2Cell()   */
    int    n;
} sphk_particles;

/* What every solver step receives (BaseSolver.h:22-26): both particle sets + their cell ranges. */
typedef struct sphk_scene {
    sphk_particles fluid;
    sphk_particles boundary;
    const int* cell_start_fluid;     /* int[ncells+1] */
    const int* cell_start_boundary;  /* int[ncells+1] */
    float radius;                    /* sphSmoothingRadius */
} sphk_scene;

enum {
    SPHK_OPT_NEIGHBOR_LIST = 1,  /* 1: sweeps walk a per-step neighbour list while positions are
                                    unchanged since the search (WCSPH, DFSPH); 0: always walk the
                                    27 cells.  default 1 */
    SPHK_OPT_LIST_CAPACITY = 2,  /* max neighbours kept per particle; particles with more fall back
                                    to the cell walk individually.  default 96 */
    SPHK_OPT_STAGED_LIST_BUILD = 10, /* 1 (default): the list builder stages each tile's 9 + 9 candidate windows in shared memory
                                    (cp.async.bulk + mbarrier) and tests candidates from there; 0: candidates read from global memory */
    SPHK_OPT_PATCH = 11,         /* experiment (default 0 = off): S > 0 makes the 4 warps of a list-sweep block take their
                                    32-particle chunks from 4 runs that lie S chunks apart in the sorted order instead of 4
                                    consecutive chunks -- with S = chunks per cell column, a block covers 4 adjacent columns x 4 cells
                                    in z, a smaller neighbour footprint per block (L1 reuse).  Results are unchanged. */
    SPHK_OPT_SIMPLE_LIST_BUILD = 6, /* 1: build the list with the generic cell walk (reference for the tuned builder) */
    SPHK_OPT_LIST_SKIN = 5,      /* neighbour-list skin in 1/1000 of the radius (default 0).  With a skin the list
                                    stays valid while sphk_pbd_delta_pos_apply moves particles by less than skin/2
                                    (tracked on the device; beyond that every sweep falls back to the cell walk) */
    SPHK_OPT_TILE = 9            /* 1: tile lists.  A tile = 128 consecutive particles; the particles any of them can interact
                                    with lie in 9 + 9 contiguous windows of the sorted fluid / boundary records
                                    (CUDAFunctions.cuh:68: z is the fastest cell dimension), which every sweep stages into
                                    shared memory with cp.async.bulk behind an mbarrier; neighbour lists hold 16-bit slots of the
                                    staged windows (half the list traffic, no L1 misses on the gathers).
                                    0 (default): per-particle int32 lists gathered from global memory through L1 -- measured
                                    faster on B200 (profiles/r02: 123 vs 175 us for a 16-byte sweep, 194 vs 327 us for a 32-byte
                                    sweep at 2M particles): random 16-byte shared-memory reads cost what L1 hits cost, and the
                                    staged tiles cap the occupancy of the velocity sweeps at 4 blocks per SM */
};

/* ---- lifetime ------------------------------------------------------------------------------ */
int  sphk_create(sphk_ctx** ctx, int max_fluid, int max_boundary, const sphk_grid* grid, void* stream);
void sphk_destroy(sphk_ctx* ctx);
int  sphk_set_option(sphk_ctx* ctx, int option, int value);
/* re-target the context at another (sub-)grid with the same cell length (a slab rank whose cuts moved): both particle
 * sets must be searched again before the next sweep; capacities and buffers are kept */
int  sphk_set_grid(sphk_ctx* ctx, const sphk_grid* grid);
int  sphk_synchronize(sphk_ctx* ctx);
const char* sphk_error_string(int code);
/* number of kernels this library has launched on ctx since creation (bench.py's gpu_launches) */
long long sphk_launch_count(const sphk_ctx* ctx);
/* account for kernels replayed from a captured CUDA graph (the host-side counter does not see them) */
int  sphk_add_launches(sphk_ctx* ctx, long long n);
/* rcp.approx(cell_length) as the device computes it (bit pattern as float): lets a CPU checker
 * reproduce the GPU cell hash bit-for-bit */
int  sphk_device_rcp(sphk_ctx* ctx, float x, float* out_host);

/* ---- neighbour search: SPHSystem::neighborSearch, SPHSystem.cu:114-127 ------------------------
 * Replaces mapParticles2Cells_CUDA (CUDAFunctions.cuh:72-78), two thrust::sort_by_key with float3
 * payloads (SPHSystem.cu:119,121), thrust::fill + countingInCell_CUDA + thrust::exclusive_scan
 * (SPHSystem.cu:123-125).  Bit-exact contract: particle2cell (left in PRE-sort order, quirk Q2), the
 * stable-sort permutation applied to pos and vel, and cell_start[ncells+1].
 * which = 0: the fluid set (p = &scene->fluid, cell_start = cell_start_fluid);
 * which = 1: the boundary set (searched once at construction, SPHSystem.cu:69). */
int sphk_neighbor_search(sphk_ctx* ctx, int which, const sphk_particles* p, int* cell_start);

/* Applies the permutation of the last fluid neighbour search to a solver-private history array, in
 * place: replaces DFSPHSolver.cu:170-171 (denWarmStiff, width 1) and PBDSolver.cu:84-85
 * (fluidPosLast, width 3), which re-sort with particle2cell as keys and rely on sort stability. */
int sphk_permute(sphk_ctx* ctx, float* array, int width, int n);

/* Re-packs the shadows from the API arrays without re-sorting (after caller-side edits). */
int sphk_refresh(sphk_ctx* ctx, const sphk_scene* scene);

/* computeBoundaryMass_CUDA, SPHSystem.cu:79-112: mass_b = rhoB / max(eps, sum_k W(|x_b - x_k|)). */
int sphk_boundary_mass(sphk_ctx* ctx, const sphk_particles* boundary, const int* cell_start_boundary,
                       float rho_boundary, float radius);

/* Particles::advect, Particles.cu:28-36: pos += dt * vel on raw arrays (no context needed: the public method of the
 * class API; the solvers' own advection is the fused sphk_advect).  `stream` is a cudaStream_t or NULL. */
int sphk_particles_advect(float* pos, const float* vel, int n, float dt, void* stream);

/* thrust::fill, SPHSystem.cu:73 / BasicSPHSolver.cu:78 */
int sphk_fill(sphk_ctx* ctx, float* array, int n, float value);

/* ---- WCSPH: BasicSPHSolver.cu -------------------------------------------------------------- */
/* force(): vel += dt*G, BasicSPHSolver.cu:227-235 */
int sphk_gravity(sphk_ctx* ctx, const sphk_scene* s, float dt, const float G[3]);
/* viscosity_CUDA + thrust::transform(vel += deltaV), BasicSPHSolver.cu:183-225.  delta_v (float3[n])
 * is the reference's bufferFloat3 and receives visc*a*dt as in the reference. */
int sphk_viscosity(sphk_ctx* ctx, const sphk_scene* s, float* delta_v, float rho0, float visc, float dt);
/* computeColorGrad_CUDA, BasicSPHSolver.cu:277-330 */
int sphk_color_grad(sphk_ctx* ctx, const sphk_scene* s, float* color_grad, float rho0, float rho_boundary);
/* surfaceTensionAndAirPressure_CUDA, BasicSPHSolver.cu:332-381 */
int sphk_surface(sphk_ctx* ctx, const sphk_scene* s, const float* color_grad, float dt, float rho0,
                 float surface_tension, float air_pressure);
/* thrust::fill(density,0) + computeDensity_CUDA, BasicSPHSolver.cu:32-83 */
int sphk_density(sphk_ctx* ctx, const sphk_scene* s);
/* computePressure_CUDA, BasicSPHSolver.cu:103-111 */
int sphk_pressure(sphk_ctx* ctx, const sphk_scene* s, float rho0, float stiff);
/* pressureForce_CUDA, BasicSPHSolver.cu:113-165 */
int sphk_pressure_force(sphk_ctx* ctx, const sphk_scene* s, float dt);
/* Particles::advect (Particles.cu:28-36) + enforceBoundary_CUDA(pos,vel) (BasicSPHSolver.cu:85-101) */
int sphk_advect(sphk_ctx* ctx, const sphk_scene* s, float dt, const float space[3]);

/* ---- fused sweeps: the same quantities as the per-launch-site entries above, computed in fewer passes over the
 * neighbour lists (each neighbour record is gathered once for two operators).  Every quantity is formed by the
 * same operations in the same order as in its own sweep.  The class layer uses these by default. */
/* computeDensity_CUDA + computeColorGrad_CUDA (both depend on positions and masses only) */
int sphk_fused_density_color_grad(sphk_ctx* ctx, const sphk_scene* s, float* color_grad, float rho0, float rho_boundary);
/* computeDensityAlpha_CUDA + computeColorGrad_CUDA */
int sphk_fused_dfsph_density_alpha_color_grad(sphk_ctx* ctx, const sphk_scene* s, float* alpha, float* color_grad,
                                              float rho0, float rho_boundary);
/* computeDensityAlpha_CUDA (+ computeColorGrad_CUDA when color_grad != NULL) + the first computeDivergenceError_CUDA of
 * DFSPHSolver::correctDivergenceError (DFSPHSolver.cu:341), which reads the velocities the density sweep leaves
 * untouched and, per particle, the density and alpha that particle has just computed. */
int sphk_fused_dfsph_density_alpha_div_error(sphk_ctx* ctx, const sphk_scene* s, float* alpha, float* color_grad_or_null,
                                             float rho0, float rho_boundary, float* error, float* stiff, float dt);
/* viscosity_CUDA (+ vel += deltaV) followed by surfaceTensionAndAirPressure_CUDA.  delta_v and color_grad must be
 * distinct buffers (the reference reuses one buffer for both because it runs them one after the other). */
int sphk_fused_viscosity_surface(sphk_ctx* ctx, const sphk_scene* s, float* delta_v, const float* color_grad,
                                 float rho0, float visc, float dt, float surface_tension, float air_pressure);

/* ---- DFSPH: DFSPHSolver.cu ----------------------------------------------------------------- */
/* computeDensityAlpha_CUDA, DFSPHSolver.cu:212-259 */
int sphk_dfsph_density_alpha(sphk_ctx* ctx, const sphk_scene* s, float* alpha);
/* computeDivergenceError_CUDA, DFSPHSolver.cu:261-306 */
int sphk_dfsph_div_error(sphk_ctx* ctx, const sphk_scene* s, const float* alpha, float* error,
                         float* stiff, float dt, float rho0);
/* correctDivergenceError_CUDA, DFSPHSolver.cu:308-329 */
int sphk_dfsph_div_correct(sphk_ctx* ctx, const sphk_scene* s, const float* stiff);
/* computeDensityError_CUDA, DFSPHSolver.cu:74-116.  warm_accumulate (may be NULL) additionally does
 * warm[i] += stiff[i], i.e. the thrust::transform of DFSPHSolver.cu:199-203 fused in. */
int sphk_dfsph_den_error(sphk_ctx* ctx, const sphk_scene* s, const float* alpha, float* error,
                         float* stiff, float dt, float rho0, float* warm_accumulate);
/* correctDensityError_CUDA, DFSPHSolver.cu:118-158 */
int sphk_dfsph_den_correct(sphk_ctx* ctx, const sphk_scene* s, const float* stiff, float dt);
/* thrust::reduce(error, abs_plus), DFSPHSolver.cu:206,360.  Synchronises; result to *host_out. */
int sphk_reduce_abs_sum(sphk_ctx* ctx, const float* x, int n, float* host_out);
/* Device-side control of an adaptive solver loop: the loop tests of DFSPHSolver.cu:187 / :347
 *     while ((iter < min_iter || totalError > threshold) && iter < max_iter) { body; ++iter; [totalError = reduce] }
 * evaluated on the device.  The host enqueues max_iter bodies; between sphk_loop_begin and sphk_loop_end every sweep
 * returns at once when the loop has ended, so no error sum travels to the host and the launch sequence is fixed (a
 * CUDA graph can replay it).  sphk_loop_next ends one iteration: ++iter, totalError = sum |error| (the same fixed-shape
 * reduction as sphk_reduce_abs_sum) from iteration reduce_from_iter on (DFSPHSolver.cu:205: the density loop reduces
 * from its second iteration), then the loop test.  slot: 0 or 1.  threshold = errorThreshold * num * rho0. */
int sphk_loop_begin(sphk_ctx* ctx, int slot, int min_iter, int max_iter, float threshold, int reduce_from_iter);
int sphk_loop_next(sphk_ctx* ctx, int slot, const float* error, int n);
int sphk_loop_end(sphk_ctx* ctx, int slot);
/* iterations the loop executed and its last error sum (synchronises) */
int sphk_loop_iterations(sphk_ctx* ctx, int slot, int* iterations_host, float* last_total_host);
/* cudaMemcpy D2D of a float array, DFSPHSolver.cu:185 */
int sphk_copy(sphk_ctx* ctx, float* dst, const float* src, int n_floats);

/* ---- PBD: PBDSolver.cu --------------------------------------------------------------------- */
/* computeDensityLambda_CUDA, PBDSolver.cu:127-168 (incl. the `bool rho0` quirk Q4) */
int sphk_pbd_density_lambda(sphk_ctx* ctx, const sphk_scene* s, float* lambda, float rho0, float relaxation);
/* computeDeltaPos_CUDA + thrust::transform(pos += dpos) + enforceBoundary_CUDA(pos),
 * PBDSolver.cu:170-256.  delta_pos (float3[n]) receives dp/rho0 as in the reference. */
int sphk_pbd_delta_pos_apply(sphk_ctx* ctx, const sphk_scene* s, const float* lambda, float* delta_pos,
                             float rho0, const float space[3]);
/* thrust::transform(vel = (pos - posLast)/dt), PBDSolver.cu:55-60 */
int sphk_pbd_velocity_from_positions(sphk_ctx* ctx, const sphk_scene* s, const float* pos_last, float dt);
/* XSPHViscosity_CUDA, PBDSolver.cu:89-125; Jacobi (the reference updates in place and races, Q5) */
int sphk_pbd_xsph(sphk_ctx* ctx, const sphk_scene* s, float c, float rho0);

/* XSPHViscosity_CUDA + computeColorGrad_CUDA in one pass over the neighbour list (the colour gradient reads positions and
 * masses only, which PBD does not change after its projection) */
int sphk_fused_pbd_xsph_color_grad(sphk_ctx* ctx, const sphk_scene* s, float c, float rho0, float* color_grad, float rho_boundary);

/* Builds the per-step neighbour list now (sweeps otherwise build it lazily on first use). */
int sphk_build_neighbor_list(sphk_ctx* ctx, const sphk_scene* s);

/* ---- scene generation on the device: initSPHSystem(), main.cpp:73-116 ---------------------------------------
 * The dam-break scene without a host-side particle array: positions are bit-identical to the host code's (same float
 * operations in the same order, no contraction) and in the reference's push order.
 * fluid block nx * ny * nz at `origin` with `spacing` (main.cpp:76-85: y outermost, then x, z innermost), restricted to the
 * x-columns [j_begin, j_begin + j_count) -- a slab rank generates only its own columns; pos_out: float3[ny * j_count * nz]. */
int sphk_scene_fluid_block(float* pos_out, int nx, int ny, int nz, const float origin[3], float spacing, int j_begin, int j_count,
                           void* stream);
/* six-face boundary shell of the box `space` on the lattice 2 * cell_size (main.cpp:89-116); pos_out: float3[count] */
long long sphk_scene_boundary_count(const int cell_size[3]);
int sphk_scene_boundary_shell(float* pos_out, const int cell_size[3], const float space[3], void* stream);

/* ---- render export: generate_dots_CUDA, vbo.cu:26-44 -----------------------------------------------
 * dot[i] = pos[i]; colour[i] from density[i] (blue below 0.75, blend to white at 1.0, blend to pink above).  Plain
 * device buffers instead of a mapped GL vertex buffer; synchronises like the reference (vbo.cu:49). */
int sphk_export_dots(sphk_ctx* ctx, const sphk_particles* p, float* dot_xyz, float* color_rgb);

/* ---- multi-GPU slab support (no reference counterpart: the reference is single-GPU) ----------------
 * A slab rank keeps [ghost-left | owned | ghost-right] particles in one sorted set (cpp-fluid-particles_b200/
 * slabs.py).  Sweeps compute only the active (owned) range; ghost values arrive by halo exchange of the API
 * arrays (contiguous slices) and are mirrored into the packed records. */
/* restrict every subsequent sweep to particles [begin, begin+count) of the fluid set (count<0: all) */
int sphk_set_active_range(sphk_ctx* ctx, int begin, int count);
/* the same with {begin, count} in DEVICE memory, read by the kernels themselves (NULL: back to the host-side range) */
int sphk_set_active_range_device(sphk_ctx* ctx, const int* device_begin_count);
/* copy API data of particles [begin, begin+count) into the packed records: what is a bit mask -- 1: vel
 * (scene->fluid.vel), 2: neighbour scalar from `array` (float[n]), 4: pos (scene->fluid.pos; PBD ghosts moved by
 * their owner: counted against the neighbour-list skin like local moves) */
int sphk_push_range(sphk_ctx* ctx, const sphk_scene* s, int what, const float* array, int begin, int count);

/* The exchanges themselves (csrc/sphk_mg.cu).  One sphk_mg_comm per rank = one process per GPU; a rank talks to
 * ranks rank-1 ("left") and rank+1 ("right") only.  Everything is enqueued on the stream given to sphk_mg_init
 * (must be the sphk_ctx stream) and nothing synchronises the host unless stated.  Transports:
 *   NCCL point-to-point (resolved with dlopen at run time, so it is the process's one NCCL) for the once-per-step
 *   candidate exchange and as fallback; peer-memory mailboxes over NVLink (CUDA IPC) for the per-sweep halos:
 *   one kernel stores the boundary planes into the neighbours' memory, waits for theirs and unpacks them into the
 *   ghost slices and the packed records. */
typedef struct sphk_mg_comm sphk_mg_comm;
/* rank 0 creates the NCCL id; ship the 128 bytes to every rank by any side channel */
int  sphk_mg_unique_id(unsigned char id[128]);
/* collective over all ranks.  mailbox_floats > 0 also allocates this rank's halo mailboxes (payload capacity per
 * message, in floats; MUST be the same number on every rank -- sphk_mg_ipc_connect checks it) */
int  sphk_mg_init(sphk_mg_comm** comm, int rank, int world, const unsigned char id[128], void* stream, long long mailbox_floats);
void sphk_mg_destroy(sphk_mg_comm* comm);
/* CUDA IPC handle of this rank's mailboxes (64 bytes) / open the neighbours' (NULL where there is none) */
int  sphk_mg_ipc_handle(sphk_mg_comm* comm, unsigned char handle[64]);
int  sphk_mg_ipc_connect(sphk_mg_comm* comm, const unsigned char* left_handle64, const unsigned char* right_handle64);
/* halo transport: 0 = NCCL send/recv + sphk_push_range (default), 1 = peer-memory mailboxes (after ipc_connect) */
int  sphk_mg_set_transport(sphk_mg_comm* comm, int transport);
/* up to 8 ints to each neighbour and back (missing neighbour: zeros).  Synchronises the stream. */
int  sphk_mg_exchange_ints(sphk_mg_comm* comm, const int* to_left, const int* to_right, int* from_left, int* from_right, int count);
/* sum of one double over all ranks.  Synchronises the stream. */
int  sphk_mg_allreduce_sum(sphk_mg_comm* comm, double* inout_host);
/* NCCL: for every array a (widths[a] floats per particle) send particles send_left = {begin, count} of send_arrays[a]
 * to the left rank and send_right to the right rank; receive recv_left / recv_right = {begin, count} of
 * recv_arrays[a].  Counts must agree with the neighbours' (exchange them first).  One NCCL group. */
int  sphk_mg_exchange_slices(sphk_mg_comm* comm, int narrays, const float* const* send_arrays, float* const* recv_arrays,
                             const int* widths, const int send_left[2], const int send_right[2],
                             const int recv_left[2], const int recv_right[2]);
/* One halo of a per-particle API array (width 1 or 3 floats): ranges = {first_begin, first_count, last_begin,
 * last_count, ghostL_begin, ghostL_count, ghostR_begin, ghostR_count} in particles -- the first / last owned plane go
 * to the left / right rank, whose planes arrive in the ghost ranges and are mirrored into the packed records as
 * sphk_push_range(what) would (what = 0: an array no record mirrors; 1: array = scene->fluid.vel; 2: neighbour
 * scalar; 4: array = scene->fluid.pos). */
int  sphk_mg_halo(sphk_mg_comm* comm, sphk_ctx* ctx, const sphk_scene* s, int what, float* array, int width, const int ranges[8]);
/* Host-free variants (a slab rank's step without a single host synchronisation):
 * sphk_mg_plane_ranges: from the cell ranges of the freshly sorted local set (cell_start_fluid, plane_cells = cy*cz cells per
 * x-plane, w owned planes + 2 ghost planes) computes on the device, into device_out24: [0..7] the plane offsets, [8..15] the
 * halo ranges in sphk_mg_halo's layout, [16..17] the owned range {begin, count} (-> sphk_set_active_range_device),
 * [18..21] next step's candidate slices, [22..23] their counts (what the neighbours must learn); also copied to pinned host
 * memory for NEXT step's host-side sizing.  sphk_mg_halo_device: sphk_mg_halo with the ranges read from device memory
 * (mailbox transport).  sphk_mg_exchange_ints_async / sphk_mg_check_async: as the synchronising calls, device-resident input,
 * results into pinned host memory when the stream gets there. */
int  sphk_mg_plane_ranges(sphk_mg_comm* comm, const int* cell_start_fluid, int plane_cells, int w, int* device_out24, int* pinned_host_out24);
int  sphk_mg_halo_device(sphk_mg_comm* comm, sphk_ctx* ctx, const sphk_scene* s, int what, float* array, int width,
                         const int* device_ranges8, int max_particles);
int  sphk_mg_exchange_ints_async(sphk_mg_comm* comm, const int* device_to_left, const int* device_to_right, int count,
                                 int* pinned_from_left, int* pinned_from_right);
int  sphk_mg_check_async(sphk_mg_comm* comm, int* pinned_error_bits);
/* mailbox error word (0 = fine; bit 0/1 timed out waiting for left/right; bit 2/3 size mismatch from left/right; bit 4 more
 * strays than the routing capacity, see sphk_mg_strays_route).
 * Synchronises the stream. */
int  sphk_mg_check(sphk_mg_comm* comm, int* error_bits_host);
/* {bytes sent, messages sent} by this rank so far */
int  sphk_mg_stats(const sphk_mg_comm* comm, long long out_host[2]);
/* Strays: owned particles that crossed TWO OR MORE cell planes in x since the last search (the candidate exchange covers
 * one plane per step; the reference's own DFSPH benchmark setting shoots a few hundred particles across tens of planes
 * once the block hits the floor).  They are routed to every rank instead.  `arrays` are the carried per-particle arrays
 * (arrays[0] = positions, width 3; at most 4 arrays of width 1..3); a block is {int count, 3 pad, rows[capacity][sum of
 * widths]} = sphk_strays_block_floats floats.
 *   sphk_strays_collect: over the owned slots [own_begin, own_begin + own_count) of the set sorted by the last search
 *     (cell_start_fluid are ITS cell ranges; call before sphk_set_grid moves the window): rows of the strays into `block`,
 *     their positions out of the world (-1e6: dropped by the next search wherever a copy is held).
 *   sphk_mg_strays_route: all-gathers the blocks (NCCL) into `gathered` (world blocks) and appends world * capacity slots
 *     at dst_begin of `arrays` (the assembled set of the next search): the strays of rank 0, 1, ... -- the same order on
 *     every rank -- unused slots out of the world.  count > capacity (the surplus stayed in place) raises bit 4 of the
 *     error word of sphk_mg_check.  sphk_strays_append: the append alone, for blocks gathered through another channel.
 *   sphk_strays_counts: the per-rank counts of a gathered set (synchronises; introspection). */
long long sphk_strays_block_floats(int capacity, int narrays, const int* widths);
int  sphk_strays_collect(sphk_ctx* ctx, const int* cell_start_fluid, int own_begin, int own_count, int narrays,
                         float* const* arrays, const int* widths, float* block, int capacity);
int  sphk_strays_append(sphk_ctx* ctx, const float* gathered, int world, int capacity, int narrays, float* const* arrays,
                        const int* widths, int dst_begin);
int  sphk_mg_strays_route(sphk_mg_comm* comm, sphk_ctx* ctx, const float* block, float* gathered, int capacity, int narrays,
                          float* const* arrays, const int* widths, int dst_begin);
int  sphk_strays_counts(sphk_ctx* ctx, const float* gathered, int world, int capacity, int narrays, const int* widths, int* counts_host);

/* ---- introspection for parity tests --------------------------------------------------------- */
/* copies the stable-sort permutation of the last fluid search (perm[s] = pre-sort index) to device
 * memory `perm_out` (int[n]) */
int sphk_get_permutation(sphk_ctx* ctx, int* perm_out, int n);
/* raw copy of the current neighbour list (device to device): counts int[n]; entries int[capacity*max_fluid] in
 * the int4-packed layout nbr4[(k/4)*max_fluid + i].{x,y,z,w}.  Either pointer may be NULL. */
int sphk_get_neighbor_list(sphk_ctx* ctx, const sphk_scene* s, int* counts_out, int* entries_out);
/* largest displacement of any particle since the skin list was built (PBD; synchronises) */
int sphk_get_skin_displacement(sphk_ctx* ctx, float* host_out);
/* neighbour-list statistics of the last build: host ints {max_count, overflow_particles, total} */
int sphk_list_stats(sphk_ctx* ctx, const sphk_scene* s, long long out_host[3]);

#ifdef __cplusplus
}
#endif
#endif /* SPHK_H_ */

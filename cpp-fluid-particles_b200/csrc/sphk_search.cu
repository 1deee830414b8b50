// sphk_search.cu -- context lifetime + neighbour search (cell hash -> stable radix sort -> gather ->
// cell ranges) + permutation / element-wise utilities.  Replaces SPHSystem::neighborSearch,
// /root/reference/src/SPHSystem.cu:114-127.
//
// Design (B200-first, not a translation of the reference's two Thrust sorts with float3 payloads):
//   1. k_hash_snapshot : one pass over pos(+vel): bit-exact cell key (MUFU.RCP path), writes
//                        particle2cell (API, pre-sort order, quirk Q2), the sort key, the identity
//                        index, and a 16-byte-aligned float4 snapshot of pos / vel.
//   2. cub::DeviceRadixSort::SortPairs on (key, index), only ceil(log2(ncells+1)) key bits
//                        -> ONE stable sort of 8-byte pairs instead of two sorts of 16-byte pairs.
//   3. k_gather        : one gather pass writes the sorted API float3 arrays and the packed 32-byte
//                        records {x,y,z,s | vx,vy,vz,m} the sweep kernels gather (one sector per neighbour).
//   4. k_cell_start    : cell_start[c] = lower_bound(sortedKeys, c) -- no atomics, no scan, equals
//                        fill + countingInCell_CUDA + exclusive_scan (SPHSystem.cu:123-125) exactly.
// HBM-bound; algorithmic bytes per particle are listed in DESIGN.md.
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdlib>
#include <new>
#include "sphk_internal.cuh"

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPHK_BLOCK)
k_hash_snapshot(const float* __restrict__ pos, const float* __restrict__ vel, int n, float cellLength,
                int3 cs, int3 org, int* __restrict__ p2c, int* __restrict__ keys, int* __restrict__ idx,
                float4* __restrict__ snapPos, float4* __restrict__ snapVel) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float3 p = load3(pos, i);
    const int key = cell_index(cell_coord(p.x, cellLength) - org.x, cell_coord(p.y, cellLength) - org.y,
                               cell_coord(p.z, cellLength) - org.z, cs);
    p2c[i] = key;
    keys[i] = key;
    idx[i] = i;
    snapPos[i] = make_float4(p.x, p.y, p.z, 0.f);
    if (vel) { const float3 v = load3(vel, i); snapVel[i] = make_float4(v.x, v.y, v.z, 0.f); }
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_gather(const int* __restrict__ idxSorted, const float4* __restrict__ snapPos,
         const float4* __restrict__ snapVel, const float* __restrict__ mass, int n,
         float* __restrict__ pos, float* __restrict__ vel, Rec rec, int isFluid,
         unsigned int* __restrict__ massRange) {
    const int s = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    float m = 0.f;
    if (s < n) {
        const int src = idxSorted[s];
        float4 p = snapPos[src];
        store3(pos, s, xyz(p));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vel) {
            v = snapVel[src];
            store3(vel, s, xyz(v));
            if (!isFluid) v = make_float4(0.f, 0.f, 0.f, 0.f);   // boundary records carry zero velocity / scalar
        }
        m = mass[s];                        // mass is NOT permuted by the reference (Q2): slot s keeps mass[s]
        v.w = m;
        p.w = isFluid ? 0.f : m;            // A.w: neighbour scalar (fluid, none yet) / mass (boundary)
        rec_store(rec + s, p, v);
    }
    if (isFluid) {                          // min / max fluid mass (non-negative floats order like their bits)
        float lo = (s < n) ? m : 3.0e38f, hi = (s < n) ? m : 0.f;
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {      // almost never taken after the first warps: no atomic storm
            if (__float_as_uint(lo) < massRange[0]) atomicMin(massRange, __float_as_uint(lo));
            if (__float_as_uint(hi) > massRange[1]) atomicMax(massRange + 1, __float_as_uint(hi));
        }
    }
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_cell_start(const int* __restrict__ keysSorted, int n, int ncells, int* __restrict__ cellStart) {
    const int c = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (c > ncells) return;
    int lo = 0, hi = n;                 // first s with keysSorted[s] >= c
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keysSorted[mid] < c) lo = mid + 1; else hi = mid;
    }
    cellStart[c] = lo;
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_permute(const int* __restrict__ idxSorted, const float* __restrict__ src, float* __restrict__ dst,
          int width, int n) {
    const int s = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (s >= n) return;
    const int from = idxSorted[s];
    for (int k = 0; k < width; ++k) dst[s * width + k] = src[from * width + k];
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_repack(const float* __restrict__ pos, const float* __restrict__ vel, const float* __restrict__ mass,
         int n, Rec rec, int isFluid, unsigned int* __restrict__ massRange) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    float m = 0.f;
    if (i < n) {
        const float3 p = load3(pos, i);
        m = mass[i];
        float4 v = make_float4(0.f, 0.f, 0.f, m);
        if (vel) { const float3 w = load3(vel, i); v = make_float4(w.x, w.y, w.z, m); }
        rec_store(rec + i, make_float4(p.x, p.y, p.z, isFluid ? 0.f : m), v);
    }
    if (massRange) {
        float lo = (i < n) ? m : 3.0e38f, hi = (i < n) ? m : 0.f;
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {      // almost never taken after the first warps: no atomic storm
            if (__float_as_uint(lo) < massRange[0]) atomicMin(massRange, __float_as_uint(lo));
            if (__float_as_uint(hi) > massRange[1]) atomicMax(massRange + 1, __float_as_uint(hi));
        }
    }
}

__global__ void k_init_mass_range(unsigned int* r) { r[0] = 0x7f7fffffu; r[1] = 0u; }

__global__ void __launch_bounds__(SPHK_BLOCK) k_fill(float* __restrict__ a, int n, float v) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void k_rcp(float x, float* out) {
    float one = 1.0f, r;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(one), "f"(x));
    // the hash multiplies by MUFU.RCP(x); div.approx(1,x) = 1 * rcp(x) is that value
    *out = r;
}

// abs-sum reduction: fixed-shape two-stage tree (deterministic run to run)
__global__ void __launch_bounds__(256) k_abs_sum_partial(const float* __restrict__ x, int n, float* __restrict__ partial) {
    __shared__ float sm[8];
    float s = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s += fabsf(x[i]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = sm[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) k_abs_sum_final(const float* __restrict__ partial, int m, float* __restrict__ out) {
    __shared__ float sm[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < m; i += 256) s += partial[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = sm[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) *out = s;
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T> static cudaError_t dalloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(T) * (count ? count : 1));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, sizeof(T) * (count ? count : 1));
    return e;
}

extern "C" int sphk_create(sphk_ctx** out, int max_fluid, int max_boundary, const sphk_grid* grid, void* stream) {
    if (!out || !grid || max_fluid <= 0 || max_boundary < 0) return SPHK_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return SPHK_ERR_NO_DEVICE; }
    sphk_ctx* c = new (std::nothrow) sphk_ctx;
    if (!c) return SPHK_ERR_ALLOC;
    c->stream = static_cast<cudaStream_t>(stream);
    c->capF = max_fluid; c->capB = max_boundary;
    c->cs = make_int3(grid->cell_size[0], grid->cell_size[1], grid->cell_size[2]);
    c->org = make_int3(grid->origin[0], grid->origin[1], grid->origin[2]);
    c->ncells = c->cs.x * c->cs.y * c->cs.z;
    c->cellLength = grid->cell_length;
    c->endBit = 1;
    while (c->endBit < 32 && (1ll << c->endBit) <= static_cast<long long>(c->ncells)) ++c->endBit;
    const size_t cap = static_cast<size_t>(max_fluid > max_boundary ? max_fluid : max_boundary);
    const size_t tot = static_cast<size_t>(max_fluid) + static_cast<size_t>(max_boundary);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = dalloc(&c->keys, cap);
    if (e == cudaSuccess) e = dalloc(&c->keysSorted, cap);
    if (e == cudaSuccess) e = dalloc(&c->idx, cap);
    if (e == cudaSuccess) e = dalloc(&c->idxSorted, cap);
    if (e == cudaSuccess) e = dalloc(&c->snapA, cap);
    if (e == cudaSuccess) e = dalloc(&c->snapB, cap);
    if (e == cudaSuccess) e = dalloc(&c->rec.a, tot + 1);   // + the far-away zero-mass dummy record (list padding)
    if (e == cudaSuccess) e = dalloc(&c->rec.b, tot + 1);
    if (e == cudaSuccess) {
        const float4 dummy = make_float4(1.0e6f, 1.0e6f, 1.0e6f, 0.f);
        e = cudaMemcpy(c->rec.a + tot, &dummy, sizeof(float4), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = dalloc(&c->massRange, 2);
    if (e == cudaSuccess) e = dalloc(&c->dispMax, 1);
    if (e == cudaSuccess) { c->cellFlagCap = static_cast<size_t>(c->ncells) + 1; e = dalloc(&c->cellFlag, c->cellFlagCap); }
    if (e == cudaSuccess) e = dalloc(&c->tmpF, 3 * static_cast<size_t>(max_fluid));
    if (e == cudaSuccess) e = dalloc(&c->partial, 1024);
    if (e == cudaSuccess) e = dalloc(&c->loops, 2);
    if (e == cudaSuccess) e = dalloc(&c->cnt, static_cast<size_t>(max_fluid));
    if (e == cudaSuccess) e = dalloc(&c->tileWin, (static_cast<size_t>(max_fluid) / SPHK_BLOCK + 1) * SPHK_TILE_WINS);
    if (e == cudaSuccess) {
        int dev = 0;
        cudaGetDevice(&dev);
        e = cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, dev);
        if (c->numSMs > 256) c->numSMs = 256;
        if (const char* v = std::getenv("SPHK_TILE")) c->tile = v[0] == '1' ? 1 : 0;
    }
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&c->pinned), 64);
    if (e == cudaSuccess) {
        c->cubTempBytes = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, c->cubTempBytes, c->keys, c->keysSorted, c->idx, c->idxSorted,
                                            static_cast<int>(cap), 0, c->endBit, c->stream);
        if (e == cudaSuccess) e = cudaMalloc(&c->cubTemp, c->cubTempBytes ? c->cubTempBytes : 1);
    }
    if (e != cudaSuccess) { sphk_destroy(c); cudaGetLastError(); return e == cudaErrorMemoryAllocation ? SPHK_ERR_ALLOC : static_cast<int>(e); }
    *out = c;
    return SPHK_OK;
}

extern "C" void sphk_destroy(sphk_ctx* c) {
    if (!c) return;
    cudaFree(c->keys); cudaFree(c->keysSorted); cudaFree(c->idx); cudaFree(c->idxSorted);
    cudaFree(c->snapA); cudaFree(c->snapB); cudaFree(c->rec.a); cudaFree(c->rec.b); cudaFree(c->massRange); cudaFree(c->dispMax); cudaFree(c->cellFlag);
    cudaFree(c->tmpF); cudaFree(c->partial); cudaFree(c->loops); cudaFree(c->cnt); cudaFree(c->nbr); cudaFree(c->cubTemp); cudaFree(c->tileWin);
    if (c->pinned) cudaFreeHost(c->pinned);
    delete c;
}

extern "C" int sphk_set_option(sphk_ctx* c, int option, int value) {
    if (!c) return SPHK_ERR_INVALID;
    switch (option) {
    case SPHK_OPT_NEIGHBOR_LIST: c->useList = value != 0; return SPHK_OK;
    case SPHK_OPT_LIST_CAPACITY:
        if (value < 8 || value > 1024 || (value & 3)) return SPHK_ERR_INVALID;   // multiple of 4 (int4 batches)
        if (value != c->kmax) {
            if (c->nbr) { cudaStreamSynchronize(c->stream); cudaFree(c->nbr); c->nbr = nullptr; }
            c->kmax = value; c->listEpoch = ~0ull;
        }
        return SPHK_OK;
    case SPHK_OPT_SIMPLE_LIST_BUILD: c->simpleBuild = value != 0; c->listEpoch = ~0ull; return SPHK_OK;
    case SPHK_OPT_PATCH: if (value < 0) return SPHK_ERR_INVALID; c->patch = value; return SPHK_OK;
    case SPHK_OPT_STAGED_LIST_BUILD: c->stagedBuild = value != 0; c->listEpoch = ~0ull; return SPHK_OK;
    case SPHK_OPT_LIST_SKIN:
        if (value < 0 || value > 1000) return SPHK_ERR_INVALID;
        if (value * 0.001f != c->skin) { c->skin = value * 0.001f; c->listEpoch = ~0ull; }
        return SPHK_OK;
    case SPHK_OPT_TILE:
        if (value != 0 && value != 1) return SPHK_ERR_INVALID;
        if (value != c->tile) { c->tile = value; c->listEpoch = ~0ull; }
        return SPHK_OK;
    default: return SPHK_ERR_INVALID;
    }
}

// Re-targets the context at another (sub-)grid of the same cell length: a slab rank whose cuts moved (load re-balancing)
// keeps its context, buffers and lists' storage; both particle sets must be searched again before the next sweep.
extern "C" int sphk_set_grid(sphk_ctx* c, const sphk_grid* grid) {
    if (!c || !grid || grid->cell_size[0] <= 0 || grid->cell_size[1] <= 0 || grid->cell_size[2] <= 0) return SPHK_ERR_INVALID;
    if (grid->cell_length != c->cellLength) return SPHK_ERR_INVALID;
    c->cs = make_int3(grid->cell_size[0], grid->cell_size[1], grid->cell_size[2]);
    c->org = make_int3(grid->origin[0], grid->origin[1], grid->origin[2]);
    c->ncells = c->cs.x * c->cs.y * c->cs.z;
    c->endBit = 1;
    while (c->endBit < 32 && (1ll << c->endBit) <= static_cast<long long>(c->ncells)) ++c->endBit;
    if (static_cast<size_t>(c->ncells) + 1 > c->cellFlagCap) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->cellFlag);
        c->cellFlag = nullptr;
        c->cellFlagCap = static_cast<size_t>(c->ncells) + 1;
        if (dalloc(&c->cellFlag, c->cellFlagCap) != cudaSuccess) { cudaGetLastError(); c->cellFlagCap = 0; return SPHK_ERR_ALLOC; }
    }
    c->fluidSearched = false; c->boundarySearched = false; c->permValid = false;
    c->listEpoch = ~0ull; c->sTag = nullptr;
    c->actBegin = 0; c->actCount = -1; c->rangeDev = nullptr;
    return SPHK_OK;
}

extern "C" int sphk_synchronize(sphk_ctx* c) {
    if (!c) return SPHK_ERR_INVALID;
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" const char* sphk_error_string(int code) {
    switch (code) {
    case SPHK_OK: return "ok";
    case SPHK_ERR_INVALID: return "sphk: invalid argument";
    case SPHK_ERR_NO_DEVICE: return "sphk: no CUDA device (libsphk has no CPU path)";
    case SPHK_ERR_CAPACITY: return "sphk: particle count exceeds context capacity";
    case SPHK_ERR_STATE: return "sphk: call order violated (neighbour search missing or stale scene)";
    case SPHK_ERR_ALLOC: return "sphk: device allocation failed";
    case SPHK_ERR_COMM: return "sphk: multi-GPU exchange failed (NCCL unavailable / error, or mailbox wiring missing)";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "sphk: unknown error";
    }
}

extern "C" long long sphk_launch_count(const sphk_ctx* c) { return c ? c->launches : 0; }
extern "C" int sphk_add_launches(sphk_ctx* c, long long n) { if (!c) return SPHK_ERR_INVALID; c->launches += n; return SPHK_OK; }

extern "C" int sphk_device_rcp(sphk_ctx* c, float x, float* out_host) {
    if (!c || !out_host) return SPHK_ERR_INVALID;
    k_rcp<<<1, 1, 0, c->stream>>>(x, c->partial);
    c->launches++;
    SPHK_CUDA_TRY(cudaMemcpyAsync(c->pinned, c->partial, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_host = *c->pinned;
    return SPHK_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int sphk_neighbor_search(sphk_ctx* c, int which, const sphk_particles* p, int* cell_start) {
    if (!c || !p || !cell_start || !p->pos || !p->mass || !p->particle2cell || p->n <= 0) return SPHK_ERR_INVALID;
    const bool fluid = which == 0;
    if (fluid && !p->vel) return SPHK_ERR_INVALID;
    const int n = p->n;
    if (n > (fluid ? c->capF : c->capB)) return SPHK_ERR_CAPACITY;
    const int off = fluid ? 0 : c->capF;
    cudaStream_t st = c->stream;
    k_hash_snapshot<<<sphk_blocks(n), SPHK_BLOCK, 0, st>>>(p->pos, p->vel, n, c->cellLength, c->cs, c->org, p->particle2cell,
                                                          c->keys, c->idx, c->snapA, c->snapB);
    size_t tb = c->cubTempBytes;
    SPHK_CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->cubTemp, tb, c->keys, c->keysSorted, c->idx, c->idxSorted, n, 0,
                                                  c->endBit, st));
    if (fluid) k_init_mass_range<<<1, 1, 0, st>>>(reinterpret_cast<unsigned int*>(c->massRange));
    k_gather<<<sphk_blocks(n), SPHK_BLOCK, 0, st>>>(c->idxSorted, c->snapA, c->snapB, p->mass, n, p->pos, p->vel,
                                                   c->rec + off, fluid ? 1 : 0, reinterpret_cast<unsigned int*>(c->massRange));
    k_cell_start<<<sphk_blocks(c->ncells + 1), SPHK_BLOCK, 0, st>>>(c->keysSorted, n, c->ncells, cell_start);
    c->launches += 3 + 4;   // 3 own kernels + CUB onesweep (histogram, scan, <=3 passes): counted as 4
    if (fluid) { c->nF = n; c->fluidSearched = true; c->permValid = true; c->searchEpoch++; c->posDirty = false; c->advected = false; c->sTag = nullptr;
                 c->actBegin = 0; c->actCount = -1; }
    else { c->nB = n; c->boundarySearched = true; c->listEpoch = ~0ull; c->permValid = false; }
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_permute(sphk_ctx* c, float* array, int width, int n) {
    if (!c || !array || (width != 1 && width != 3)) return SPHK_ERR_INVALID;
    if (!c->permValid || n != c->nF) return SPHK_ERR_STATE;
    if (array == c->sTag) c->sTag = nullptr;
    k_permute<<<sphk_blocks(n), SPHK_BLOCK, 0, c->stream>>>(c->idxSorted, array, c->tmpF, width, n);
    c->launches++;
    SPHK_CUDA_TRY(cudaMemcpyAsync(array, c->tmpF, sizeof(float) * static_cast<size_t>(n) * width,
                                  cudaMemcpyDeviceToDevice, c->stream));
    return SPHK_OK;
}

extern "C" int sphk_refresh(sphk_ctx* c, const sphk_scene* s) {
    if (!c || !s) return SPHK_ERR_INVALID;
    if (!c->fluidSearched || s->fluid.n != c->nF) return SPHK_ERR_STATE;
    k_init_mass_range<<<1, 1, 0, c->stream>>>(reinterpret_cast<unsigned int*>(c->massRange));
    k_repack<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(s->fluid.pos, s->fluid.vel, s->fluid.mass, c->nF, c->rec, 1,
                                                              reinterpret_cast<unsigned int*>(c->massRange));
    c->launches++;
    c->posDirty = true;
    c->advected = true;
    c->sTag = nullptr;
    if (c->boundarySearched && s->boundary.pos && s->boundary.n == c->nB) {
        k_repack<<<sphk_blocks(c->nB), SPHK_BLOCK, 0, c->stream>>>(s->boundary.pos, nullptr, s->boundary.mass, c->nB,
                                                                  c->rec + c->capF, 0, nullptr);
        c->launches++;
    }
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

__global__ void __launch_bounds__(SPHK_BLOCK) k_axpy(float* __restrict__ pos, const float* __restrict__ vel, int n3, float dt) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i < n3) pos[i] = pos[i] + dt * vel[i];
}

extern "C" int sphk_particles_advect(float* pos, const float* vel, int n, float dt, void* stream) {
    if (!pos || !vel || n < 0) return SPHK_ERR_INVALID;
    if (n == 0) return SPHK_OK;
    k_axpy<<<sphk_blocks(3 * n), SPHK_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(pos, vel, 3 * n, dt);
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_fill(sphk_ctx* c, float* array, int n, float value) {
    if (!c || !array || n < 0) return SPHK_ERR_INVALID;
    if (n == 0) return SPHK_OK;
    if (array == c->sTag) c->sTag = nullptr;
    k_fill<<<sphk_blocks(n), SPHK_BLOCK, 0, c->stream>>>(array, n, value);
    c->launches++;
    // A fill of an API mass array AFTER the search that packed it (the reference kernels read mass[] live) is not
    // mirrored here: the next sphk_neighbor_search / sphk_refresh re-packs it.  SPHSystem's constructor fills the
    // fluid mass before the fluid search that precedes its first sweep (SPHSystem.cu:73-76), so the class layer never
    // needs the refresh.
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_copy(sphk_ctx* c, float* dst, const float* src, int n_floats) {
    if (!c || !dst || !src || n_floats < 0) return SPHK_ERR_INVALID;
    if (dst == c->sTag) c->sTag = nullptr;
    SPHK_CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(float) * static_cast<size_t>(n_floats), cudaMemcpyDeviceToDevice, c->stream));
    return SPHK_OK;
}

extern "C" int sphk_reduce_abs_sum(sphk_ctx* c, const float* x, int n, float* host_out) {
    if (!c || !x || !host_out || n < 0) return SPHK_ERR_INVALID;
    int blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    k_abs_sum_partial<<<blocks, 256, 0, c->stream>>>(x, n, c->partial);
    k_abs_sum_final<<<1, 256, 0, c->stream>>>(c->partial, blocks, c->partial);
    c->launches += 2;
    SPHK_CUDA_TRY(cudaMemcpyAsync(c->pinned, c->partial, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *host_out = *c->pinned;
    return SPHK_OK;
}

// ---- device-side scene generation: initSPHSystem(), main.cpp:73-116, without a host-side particle array ---------
// The fluid block in the reference's push order (i = y outermost, j = x, k = z innermost; main.cpp:76-85), optionally
// restricted to the x-columns [jBegin, jBegin + jCount) (a slab rank generates only its own columns; the restriction keeps
// the relative order, which is all the stable sort sees).  Arithmetic as the host code performs it: float multiply, then
// float add, no contraction -- the positions are bit-identical to scene.py / main.cpp.
__global__ void __launch_bounds__(SPHK_BLOCK)
k_scene_fluid_block(float* __restrict__ pos, int nz, int jBegin, int jCount, long long count, float ox, float oy, float oz, float spacing) {
    const long long p = static_cast<long long>(blockIdx.x) * SPHK_BLOCK + threadIdx.x;
    if (p >= count) return;
    const int k = static_cast<int>(p % nz);
    const long long q = p / nz;
    const int j = jBegin + static_cast<int>(q % jCount);
    const int i = static_cast<int>(q / jCount);
    pos[3 * p] = __fadd_rn(ox, __fmul_rn(spacing, static_cast<float>(j)));
    pos[3 * p + 1] = __fadd_rn(oy, __fmul_rn(spacing, static_cast<float>(i)));
    pos[3 * p + 2] = __fadd_rn(oz, __fmul_rn(spacing, static_cast<float>(k)));
}

extern "C" int sphk_scene_fluid_block(float* pos_out, int nx, int ny, int nz, const float origin[3], float spacing, int j_begin,
                                      int j_count, void* stream) {
    if (!pos_out || !origin || nx <= 0 || ny <= 0 || nz <= 0 || j_begin < 0 || j_count < 0 || j_begin + j_count > nx) return SPHK_ERR_INVALID;
    const long long count = static_cast<long long>(ny) * j_count * nz;
    if (count == 0) return SPHK_OK;
    const long long blocks = (count + SPHK_BLOCK - 1) / SPHK_BLOCK;
    if (blocks > 0x7fffffffLL) return SPHK_ERR_INVALID;
    k_scene_fluid_block<<<static_cast<unsigned int>(blocks), SPHK_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(
        pos_out, nz, j_begin, j_count, count, origin[0], origin[1], origin[2], spacing);
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

// The six-face boundary shell, main.cpp:89-116: lattice c = 2 * cellSize per axis, point (i,j,k) -> 0.99 * ((i,j,k) / (c - 1) *
// spaceSize) + 0.005 * spaceSize, pushed as front/back pairs (i in cx, j in cy), top/bottom pairs (i in cx, j in cz - 2),
// left/right pairs (i in cy - 2, j in cz - 2).  IEEE divide / multiply / add, no contraction.
__device__ __forceinline__ float shell_coord(int idx, int c, float space) {
    const float x = __fmul_rn(__fdiv_rn(static_cast<float>(idx), static_cast<float>(c - 1)), space);
    return __fadd_rn(__fmul_rn(0.99f, x), __fmul_rn(0.005f, space));
}
__global__ void __launch_bounds__(SPHK_BLOCK)
k_scene_boundary_shell(float* __restrict__ pos, int cx, int cy, int cz, float sx, float sy, float sz, long long n0, long long n1, long long n2) {
    const long long p = static_cast<long long>(blockIdx.x) * SPHK_BLOCK + threadIdx.x;
    if (p >= n0 + n1 + n2) return;
    int ix, iy, iz;
    if (p < n0) {                       // front and back
        const long long q = p >> 1; const int second = static_cast<int>(p & 1);
        ix = static_cast<int>(q / cy); iy = static_cast<int>(q % cy); iz = second ? cz - 1 : 0;
    } else if (p < n0 + n1) {           // top and bottom
        const long long r = p - n0, q = r >> 1; const int second = static_cast<int>(r & 1);
        ix = static_cast<int>(q / (cz - 2)); iz = static_cast<int>(q % (cz - 2)) + 1; iy = second ? cy - 1 : 0;
    } else {                            // left and right
        const long long r = p - n0 - n1, q = r >> 1; const int second = static_cast<int>(r & 1);
        iy = static_cast<int>(q / (cz - 2)) + 1; iz = static_cast<int>(q % (cz - 2)) + 1; ix = second ? cx - 1 : 0;
    }
    pos[3 * p] = shell_coord(ix, cx, sx);
    pos[3 * p + 1] = shell_coord(iy, cy, sy);
    pos[3 * p + 2] = shell_coord(iz, cz, sz);
}

extern "C" long long sphk_scene_boundary_count(const int cell_size[3]) {
    if (!cell_size) return -1;
    const long long cx = 2LL * cell_size[0], cy = 2LL * cell_size[1], cz = 2LL * cell_size[2];
    if (cx < 2 || cy < 2 || cz < 2) return -1;
    return 2 * cx * cy + 2 * cx * (cz - 2) + 2 * (cy - 2) * (cz - 2);
}

extern "C" int sphk_scene_boundary_shell(float* pos_out, const int cell_size[3], const float space[3], void* stream) {
    if (!pos_out || !cell_size || !space) return SPHK_ERR_INVALID;
    const long long total = sphk_scene_boundary_count(cell_size);
    if (total <= 0) return SPHK_ERR_INVALID;
    const int cx = 2 * cell_size[0], cy = 2 * cell_size[1], cz = 2 * cell_size[2];
    const long long n0 = 2LL * cx * cy, n1 = 2LL * cx * (cz - 2), n2 = 2LL * (cy - 2) * (cz - 2);
    const long long blocks = (total + SPHK_BLOCK - 1) / SPHK_BLOCK;
    k_scene_boundary_shell<<<static_cast<unsigned int>(blocks), SPHK_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(
        pos_out, cx, cy, cz, space[0], space[1], space[2], n0, n1, n2);
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

// ---- device-side loop control (adaptive DFSPH, DFSPHSolver.cu:187,205-207,347,360) ----------------------------------
__global__ void k_loop_begin(LoopState* st, int minIter, int maxIter, float threshold, int reduceFrom) {
    st->iters = 0; st->minIter = minIter; st->maxIter = maxIter; st->reduceFrom = reduceFrom;
    st->threshold = threshold; st->total = 3.402823466e38f;
    st->active = ((0 < minIter) || (st->total > threshold)) && (0 < maxIter) ? 1 : 0;
}
// first stage of the error sum of iteration iters + 1 (same fixed-shape tree as sphk_reduce_abs_sum: same sum, bit for bit)
__global__ void __launch_bounds__(256) k_loop_partial(const LoopState* st, const float* __restrict__ x, int n, float* __restrict__ partial) {
    if (!st->active || st->iters + 1 < st->reduceFrom) return;
    __shared__ float sm[8];
    float s = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s += fabsf(x[i]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = sm[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
// second stage + the loop test: ++iter; total = sum (from iteration reduceFrom on); active = (iter < min || total > thr) && iter < max
__global__ void __launch_bounds__(256) k_loop_finish(LoopState* st, const float* __restrict__ partial, int m) {
    if (!st->active) return;
    __shared__ float sm[8];
    const int iters = st->iters + 1;
    float total = st->total;
    if (iters >= st->reduceFrom) {
        float s = 0.f;
        for (int i = threadIdx.x; i < m; i += 256) s += partial[i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 8) {
            s = sm[threadIdx.x];
            for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
            if (threadIdx.x == 0) sm[0] = s;
        }
        __syncthreads();
        total = sm[0];
    }
    if (threadIdx.x == 0) {
        st->iters = iters;
        st->total = total;
        st->active = ((iters < st->minIter) || (total > st->threshold)) && (iters < st->maxIter) ? 1 : 0;
    }
}

extern "C" int sphk_loop_begin(sphk_ctx* c, int slot, int min_iter, int max_iter, float threshold, int reduce_from_iter) {
    if (!c || slot < 0 || slot > 1 || max_iter < 0) return SPHK_ERR_INVALID;
    k_loop_begin<<<1, 1, 0, c->stream>>>(c->loops + slot, min_iter, max_iter, threshold, reduce_from_iter);
    c->launches++;
    c->pred = &c->loops[slot].active;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_loop_next(sphk_ctx* c, int slot, const float* error, int n) {
    if (!c || slot < 0 || slot > 1 || !error || n < 0) return SPHK_ERR_INVALID;
    int blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    k_loop_partial<<<blocks, 256, 0, c->stream>>>(c->loops + slot, error, n, c->partial);
    k_loop_finish<<<1, 256, 0, c->stream>>>(c->loops + slot, c->partial, blocks);
    c->launches += 2;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_loop_end(sphk_ctx* c, int slot) {
    if (!c || slot < 0 || slot > 1) return SPHK_ERR_INVALID;
    c->pred = nullptr;
    return SPHK_OK;
}

extern "C" int sphk_loop_iterations(sphk_ctx* c, int slot, int* iterations_host, float* last_total_host) {
    if (!c || slot < 0 || slot > 1 || !iterations_host) return SPHK_ERR_INVALID;
    LoopState h;
    SPHK_CUDA_TRY(cudaMemcpyAsync(&h, c->loops + slot, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *iterations_host = h.iters;
    if (last_total_host) *last_total_host = h.total;
    return SPHK_OK;
}

// generate_dots_CUDA, vbo.cu:26-44
__global__ void __launch_bounds__(SPHK_BLOCK)
k_export_dots(const float* __restrict__ pos, const float* __restrict__ density, float* __restrict__ dot, float* __restrict__ color, int n) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    store3(dot, i, load3(pos, i));
    const float d = density[i];
    const float3 blue = f3(0.34f, 0.46f, 0.7f), white = f3(0.9f, 0.9f, 0.9f), pink = f3(1.0f, 0.4f, 0.7f);
    float3 c;
    if (d < 0.75f) c = blue;
    else if (d < 1.0f) { const float w = (d - 0.75f) * 4.0f; c = w * white + (1 - w) * blue; }
    else { float w = (powf(d, 2) - 1.0f) * 4.0f; w = fminf(w, 1.0f); c = (1 - w) * white + w * pink; }
    store3(color, i, c);
}

extern "C" int sphk_export_dots(sphk_ctx* c, const sphk_particles* p, float* dot, float* color) {
    if (!c || !p || !p->pos || !p->density || !dot || !color || p->n < 0) return SPHK_ERR_INVALID;
    if (p->n > 0) k_export_dots<<<sphk_blocks(p->n), SPHK_BLOCK, 0, c->stream>>>(p->pos, p->density, dot, color, p->n);
    c->launches++;
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_get_permutation(sphk_ctx* c, int* perm_out, int n) {
    if (!c || !perm_out) return SPHK_ERR_INVALID;
    if (!c->permValid || n != c->nF) return SPHK_ERR_STATE;
    SPHK_CUDA_TRY(cudaMemcpyAsync(perm_out, c->idxSorted, sizeof(int) * static_cast<size_t>(n), cudaMemcpyDeviceToDevice, c->stream));
    return SPHK_OK;
}

// sphk_internal.cuh -- shared device helpers and the context of libsphk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "sphk.h"

#define SPHK_EPS (1e-6f)                        // global.h:21
#define SPHK_PI (3.14159265358979323846f)       // global.h:22
#define SPHK_MAX_A (1000.0f)                    // global.h:26
#define SPHK_BLOCK 128

struct sphk_ctx {
    cudaStream_t stream = nullptr;
    int capF = 0, capB = 0;          // capacities; boundary lives at unified index capF + b
    int3 cs = {0, 0, 0};
    int ncells = 0;
    float cellLength = 0.f;
    int endBit = 32;                 // radix sort key width: ceil(log2(ncells + 1))
    // ---- scratch (device) ----
    int *keys = nullptr, *keysSorted = nullptr, *idx = nullptr, *idxSorted = nullptr;
    void* cubTemp = nullptr; size_t cubTempBytes = 0;
    float4 *snapA = nullptr, *snapB = nullptr;   // [max(capF,capB)] snapshot / Jacobi temp
    float4* posm = nullptr;                      // [capF + capB] xyz + mass, sorted order
    float4* vel4 = nullptr;                      // [capF] xyz (+w unused), sorted order
    float* aux = nullptr;                        // [capF + capB] per-sweep scalar (p/rho^2, |c|^2); boundary part 0
    float* tmpF = nullptr;                       // [3*capF] permute temp
    float* partial = nullptr;                    // [1024] reduction partials
    int* nbr = nullptr;                          // [kmax * capF] neighbour list, nbr[k*capF + i]
    int* cnt = nullptr;                          // [capF] true neighbour count (may exceed kmax)
    float* pinned = nullptr;                     // host pinned scalar
    // ---- state ----
    int nF = 0, nB = 0;
    int kmax = 96;
    bool useList = true, useTile = false;
    unsigned long long searchEpoch = 0, listEpoch = ~0ull;
    bool posDirty = false;
    bool fluidSearched = false, boundarySearched = false, permValid = false;
    const int* lastCsB = nullptr;
    long long launches = 0;
};

struct DevScene {
    const float4* __restrict__ posm;
    const int* __restrict__ csF;
    const int* __restrict__ csB;
    const int* __restrict__ nbr;
    const int* __restrict__ cnt;
    int nF, bOff, nbrStride, kmax;
    int3 cs;
    float cellLength, R, r2cut;
};

// ---- tiny float3 algebra (component-wise, left to right) ---------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__device__ __forceinline__ float3 load3(const float* __restrict__ p, int i) {
    return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
}
__device__ __forceinline__ void store3(float* __restrict__ p, int i, float3 v) {
    p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z;
}

// ---- the cell hash, bit-exact with the reference under -use_fast_math ----------------------------
// make_int3(pos / cellLength) compiles to MUFU.RCP + FMUL.FTZ + F2I.FTZ.TRUNC in the reference build
// (CUDAFunctions.cuh:76, src/CMakeLists.txt:43).  div.approx.ftz + cvt.rzi.ftz is that sequence.
__device__ __forceinline__ int cell_coord(float x, float cellLength) {
    float q; int c;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(x), "f"(cellLength));
    asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(c) : "f"(q));
    return c;
}
// particlePos2cellIdx, CUDAFunctions.cuh:64-70
__device__ __forceinline__ int cell_index(int x, int y, int z, int3 cs) {
    return (x >= 0 && x < cs.x && y >= 0 && y < cs.y && z >= 0 && z < cs.z) ? ((x * cs.y + y) * cs.z + z)
                                                                            : (cs.x * cs.y * cs.z);
}

// ---- smoothing kernels, CUDAFunctions.cuh:23-54,82-98 (r = |d| passed in where already known) ----
__device__ __forceinline__ float w_cubic(float r, float R) {
    const float q = 2.0f * fabsf(r) / R;
    if (q > 2.0f || q < SPHK_EPS) return 0.0f;
    const float a = 0.25f / (SPHK_PI * R * R * R);
    return a * ((q > 1.0f) ? (2.0f - q) * (2.0f - q) * (2.0f - q) : ((3.0f * q - 6.0f) * q * q + 4.0f));
}
__device__ __forceinline__ float3 grad_w_cubic(float3 d, float r, float R) {
    const float q = 2.0f * r / R;
    if (q > 2.0f) return f3(0.f, 0.f, 0.f);
    const float3 a = d / (SPHK_PI * (q + SPHK_EPS) * R * R * R * R * R);
    return a * ((q > 1.0f) ? ((12.0f - 3.0f * q) * q - 12.0f) : ((9.0f * q - 12.0f) * q));
}
__device__ __forceinline__ float lap_visc(float r, float R) {
    return (r <= R) ? (45.0f * (R - r) / (SPHK_PI * powf(R, 6))) : 0.0f;
}
__device__ __forceinline__ float3 grad_surface_tension(float3 d, float x, float R) {
    if (x > R || x < SPHK_EPS) return f3(0.f, 0.f, 0.f);
    const float R3 = R * R * R;
    const float3 a = 136.0241f * -d / (SPHK_PI * R3 * R3 * R3 * x);
    const float e = R - x;
    const float e3x3 = (e * e * e) * (x * x * x);
    return a * ((2.0f * x <= R) ? (2.0f * e3x3 - 0.0156f * R3 * R3) : e3x3);
}

#define SPHK_CUDA_TRY(expr)                                  \
    do {                                                     \
        cudaError_t e_ = (expr);                             \
        if (e_ != cudaSuccess) return static_cast<int>(e_);  \
    } while (0)

static inline int sphk_blocks(int n) { return (n + SPHK_BLOCK - 1) / SPHK_BLOCK; }

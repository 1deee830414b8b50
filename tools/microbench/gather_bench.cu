// Microbenchmark: random 16-byte / 32-byte gathers from an L1-resident global table vs from shared memory.
// Question it answers: would staging neighbour records in shared memory lift the
// L1-gather bound of the list sweeps?   nvcc -arch=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define TABLE 1296      // candidates of a 1x1x16-cell tile at 8 particles/cell
#define ITERS 32        // neighbours per particle
#define THREADS 128
template <int MODE>     // 0: global float4, 1: smem float4, 2: global 2xfloat4 (32 B), 3: smem 2xfloat4
__global__ void k(const float4* __restrict__ table, const unsigned short* __restrict__ idx, float* out, int ntiles) {
    extern __shared__ float4 sm[];
    const int tile = blockIdx.x;
    const float4* src = table + (size_t)tile * TABLE * 2;
    if (MODE == 1 || MODE == 3) {
        for (int t = threadIdx.x; t < TABLE * (MODE == 3 ? 2 : 1); t += THREADS) sm[t] = src[t];
        __syncthreads();
    }
    float acc = 0.f;
    const unsigned short* my = idx + ((size_t)tile * THREADS + threadIdx.x);
    const size_t stride = (size_t)ntiles * THREADS;
#pragma unroll 4
    for (int k2 = 0; k2 < ITERS; ++k2) {
        const int j = my[k2 * stride];
        if (MODE == 0) { float4 v = src[j]; acc += v.x * v.y + v.z * v.w; }
        if (MODE == 1) { float4 v = sm[j]; acc += v.x * v.y + v.z * v.w; }
        if (MODE == 2) { float4 v = src[2 * j], w = src[2 * j + 1]; acc += v.x * w.y + v.z * w.w; }
        if (MODE == 3) { float4 v = sm[2 * j], w = sm[2 * j + 1]; acc += v.x * w.y + v.z * w.w; }
    }
    out[(size_t)tile * THREADS + threadIdx.x] = acc;
}
int main() {
    const int ntiles = 16384;
    float4* table; unsigned short* idx; float* out;
    cudaMalloc(&table, sizeof(float4) * 2 * TABLE * ntiles);
    cudaMemset(table, 0, sizeof(float4) * 2 * TABLE * ntiles);
    cudaMalloc(&idx, sizeof(unsigned short) * (size_t)ntiles * THREADS * ITERS);
    cudaMalloc(&out, sizeof(float) * ntiles * THREADS);
    unsigned short* h = (unsigned short*)malloc(sizeof(unsigned short) * (size_t)ntiles * THREADS * ITERS);
    // neighbour pattern like the real lists: thread t (particle in cell t/8) picks indices clustered around 9 row windows
    for (int tile = 0; tile < ntiles; ++tile)
        for (int t = 0; t < THREADS; ++t)
            for (int k2 = 0; k2 < ITERS; ++k2) {
                int row = k2 * 9 / ITERS, cell = t / 8;
                int base = row * 144 + cell * 8;                 // window of 3 cells = 24 candidates
                h[((size_t)k2 * ntiles + tile) * THREADS + t] = (unsigned short)(base + rand() % 24);
            }
    cudaMemcpy(idx, h, sizeof(unsigned short) * (size_t)ntiles * THREADS * ITERS, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"global 16B", "smem   16B", "global 32B", "smem   32B"};
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            size_t sh = (mode == 1 ? TABLE : mode == 3 ? 2 * TABLE : 0) * sizeof(float4);
            if (mode == 0) k<0><<<ntiles, THREADS>>>(table, idx, out, ntiles);
            if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); k<1><<<ntiles, THREADS, sh>>>(table, idx, out, ntiles); }
            if (mode == 2) k<2><<<ntiles, THREADS>>>(table, idx, out, ntiles);
            if (mode == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); k<3><<<ntiles, THREADS, sh>>>(table, idx, out, ntiles); }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s gather: %.3f ms for %.1f M gathers (incl. table staging for smem)  err=%s\n", names[mode], best,
               (double)ntiles * THREADS * ITERS / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

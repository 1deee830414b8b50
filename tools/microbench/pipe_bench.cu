// pipe_bench.cu -- what a warp-wide scattered load costs on the B200 L1/shared-memory data pipe.
//
// Design input for the neighbour sweeps: the sweeps are bound by
// l1tex__data_pipe_lsu_wavefronts.  This measures, in SM clocks per warp-level load instruction at full
// occupancy, how that cost depends on
//   G16 / G32 : global loads of 16 / 32 bytes per lane whose 32 addresses fall into L distinct 128-byte lines
//               (random record inside a window of L lines; record stride 16 or 32 bytes), L1-resident tables;
//   S16       : shared-memory 16-byte loads, random slot inside a window of W 16-byte slots (bank conflicts);
//   S16x2     : two 16-byte shared loads per lane (a 32-byte record as two halves W slots apart);
//   active    : the same with only a subset of lanes active (row-synchronised list walks idle some lanes).
// Not a kernel of the product; its numbers are design input only (no claim in DESIGN.md rests on it alone).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define THREADS 256
#define ITERS 512

__device__ __forceinline__ unsigned int lcg(unsigned int& x) { x = x * 1664525u + 1013904223u; return x >> 4; }

// each warp owns a window of `lines` 128-byte lines inside a per-block region (L1-resident after the first touch)
template <int BYTES>   // 16: float4 records at 16-byte stride; 32: 32-byte records, one 256-bit load
__global__ void __launch_bounds__(THREADS) k_global(const float4* __restrict__ table, int lines, int regionLines,
                                                     unsigned int activeMask, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int recsPerLine = 128 / BYTES;
    const int W = lines * recsPerLine;
    unsigned int x = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const bool on = (activeMask >> lane) & 1u;
    // window start moves slowly (like consecutive tiles of a z-run), staying inside the block's region
    const char* region = reinterpret_cast<const char*>(table) + (size_t)(blockIdx.x % 64) * regionLines * 128;
    if (on) {
#pragma unroll 1
        for (int it = 0; it < ITERS; it += 4) {
            const int wbase = ((warp * 7 + (it >> 4)) % (regionLines - lines + 1)) * 128;
            float4 v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned int r = __umulhi(lcg(x) << 4, (unsigned int)W);
                const char* p = region + wbase + (size_t)r * BYTES;
                if (BYTES == 16) v[u] = *reinterpret_cast<const float4*>(p);
                else asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                  : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w), "=f"(w[u].x), "=f"(w[u].y), "=f"(w[u].z), "=f"(w[u].w)
                                  : "l"(p));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc += v[u].x * v[u].y + v[u].z * v[u].w; if (BYTES == 32) acc += w[u].x * w[u].w; }
        }
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

// shared memory: `slots` float4 staged per block; each warp reads random slots inside a window of W slots
template <int HALVES>
__global__ void __launch_bounds__(THREADS) k_shared(const float4* __restrict__ table, int slots, int W, int mode,
                                                     unsigned int activeMask, float* out) {
    extern __shared__ float4 sm[];
    for (int t = threadIdx.x; t < slots * HALVES; t += THREADS) sm[t] = table[t];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int x = (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const bool on = (activeMask >> lane) & 1u;
    if (on) {
#pragma unroll 1
        for (int it = 0; it < ITERS; it += 4) {
            const int wbase = (warp * 97 + (it >> 4) * 8) % (slots - W + 1);
            float4 v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                unsigned int r = __umulhi(lcg(x) << 4, (unsigned int)W);
                if (mode == 1) r = lane;                        // conflict-free reference
                if (mode == 2) r = (lane >> 3) * 3 + (r % 3);   // 8 lanes of a "cell" share 3 slots (broadcast-heavy)
                v[u] = sm[wbase + r];
                if (HALVES == 2) w[u] = sm[slots + wbase + r];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc += v[u].x * v[u].y + v[u].z * v[u].w; if (HALVES == 2) acc += w[u].x * w[u].w; }
        }
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

static float time_it(void (*launch)(void*), void* arg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); launch(arg); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

struct GArgs { const float4* table; int lines, regionLines, bytes; unsigned int mask; float* out; int blocks; };
static void launch_g(void* p) {
    GArgs* a = (GArgs*)p;
    if (a->bytes == 16) k_global<16><<<a->blocks, THREADS>>>(a->table, a->lines, a->regionLines, a->mask, a->out);
    else k_global<32><<<a->blocks, THREADS>>>(a->table, a->lines, a->regionLines, a->mask, a->out);
}
struct SArgs { const float4* table; int slots, W, mode, halves; unsigned int mask; float* out; int blocks; };
static void launch_s(void* p) {
    SArgs* a = (SArgs*)p;
    const size_t sh = sizeof(float4) * a->slots * a->halves;
    if (a->halves == 1) k_shared<1><<<a->blocks, THREADS, sh>>>(a->table, a->slots, a->W, a->mode, a->mask, a->out);
    else k_shared<2><<<a->blocks, THREADS, sh>>>(a->table, a->slots, a->W, a->mode, a->mask, a->out);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    const int blocks = sms * 8;
    const double ghz = clk * 1e-6;
    printf("device %s, %d SMs, %.3f GHz (attribute), %d blocks x %d threads, %d loads per lane\n", prop.name, sms, ghz, blocks, THREADS, ITERS);
    float4* table; float* out;
    const size_t tableBytes = 64ull * 512 * 128;
    cudaMalloc(&table, tableBytes); cudaMemset(table, 0, tableBytes);
    cudaMalloc(&out, sizeof(float) * blocks * THREADS);
    cudaFuncSetAttribute(k_shared<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_shared<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const double warpLoads = (double)blocks * (THREADS / 32) * ITERS;
    auto clkPer = [&](float ms) { return ms * 1e-3 * ghz * 1e9 * sms / warpLoads; };
    const int Ls[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64};
    for (int bytes : {16, 32})
        for (int L : Ls) {
            GArgs a{table, L, 160, bytes, 0xffffffffu, out, blocks};
            const float ms = time_it(launch_g, &a);
            printf("G%d  window %2d lines  all lanes : %7.3f ms  %6.2f clk per warp load\n", bytes, L, ms, clkPer(ms));
        }
    for (unsigned int mask : {0x0000ffffu, 0x55555555u, 0x00ff00ffu, 0x77777777u})
        for (int L : {3, 6, 12}) {
            GArgs a{table, L, 160, 16, mask, out, blocks};
            const float ms = time_it(launch_g, &a);
            printf("G16 window %2d lines  mask %08x : %7.3f ms  %6.2f clk per warp load\n", L, mask, ms, clkPer(ms));
        }
    for (int halves : {1, 2}) {
        for (int mode : {1, 2}) {
            SArgs a{table, 1296, 48, mode, halves, 0xffffffffu, out, blocks};
            const float ms = time_it(launch_s, &a);
            printf("S16x%d mode %s : %7.3f ms  %6.2f clk per warp load (pair)\n", halves, mode == 1 ? "conflict-free" : "cell-broadcast", ms, clkPer(ms));
        }
        for (int W : {8, 24, 48, 144, 432, 1296}) {
            SArgs a{table, 1296, W, 0, halves, 0xffffffffu, out, blocks};
            const float ms = time_it(launch_s, &a);
            printf("S16x%d random in window of %4d slots : %7.3f ms  %6.2f clk per warp load (pair)\n", halves, W, ms, clkPer(ms));
        }
    }
    for (unsigned int mask : {0x0000ffffu, 0x55555555u}) {
        SArgs a{table, 1296, 48, 0, 1, mask, out, blocks};
        const float ms = time_it(launch_s, &a);
        printf("S16x1 random window 48, mask %08x : %7.3f ms  %6.2f clk per warp load\n", mask, ms, clkPer(ms));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

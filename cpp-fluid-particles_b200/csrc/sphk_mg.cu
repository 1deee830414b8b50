// sphk_mg.cu -- multi-GPU slab support of libsphk: the neighbour exchanges of the x-slab decomposition
// (the reference is single-GPU, so nothing here replaces reference code).
//
// One process per GPU.  A rank talks to its two x-neighbours only.  Two transports, both enqueued on the
// context's stream with no host synchronisation:
//
//   * NCCL point-to-point (ncclSend/ncclRecv in one group), called natively.  NCCL is resolved with dlopen at
//     run time -- inside a torch process that is torch's own libnccl.so.2 (same soname), so there is exactly one
//     NCCL in the process; libsphk itself has no link-time NCCL dependency.  Used for the once-per-step
//     candidate exchange (variable size, several arrays) and as the fallback halo transport.
//
//   * peer-memory mailboxes (CUDA IPC over NVLink / NVSwitch) for the ~21 per-sweep halos of a step:
//     ONE kernel per halo stores this rank's first / last owned plane straight into the neighbours' mailboxes
//     (remote stores), publishes a sequence flag, waits for the neighbours' flags on its own (local) mailboxes and
//     unpacks them into the ghost slices of the API array AND the packed 32-byte records the sweeps gather
//     (what sphk_push_range does after an NCCL halo).  No packing kernels, no NCCL launch latency, no extra
//     record-mirror launches: 1 launch per halo instead of 3.
//     Flow control: message k goes to mailbox k & 1.  A rank sends k+2 only after it has received k+1, which
//     its neighbour sent after it had consumed k (stream order), so two mailboxes per direction suffice.
//     Each message carries its element count: a mismatch between the owner's plane and the neighbour's ghost
//     range raises a device-side error flag (read by sphk_mg_check) instead of hanging; waits are bounded.
#include "sphk_internal.cuh"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <new>

// ---- the NCCL entry points this file uses (nccl.h 2.27: ncclUniqueId is 128 opaque bytes, passed by value) -------
namespace {

struct NcclUid { char internal[128]; };
typedef void* NcclComm;
enum { kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclInt32 = 2, kNcclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi load_nccl() {
    NcclApi api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    bool all = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) all = false; return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = all;
    return api;
}

NcclApi& nccl() {
    static NcclApi api = load_nccl();       // resolved once (thread-safe initialisation)
    return api;
}

// mailbox layout (per rank, one cudaMalloc, exported over CUDA IPC):
//   info              : MailInfo (256 B): magic + payload capacity, so a neighbour can check that both sides agree on
//                       the layout before it computes addresses inside this block
//   box[side][parity] : Header (64 B) + capFloats floats        side 0 = written by the LEFT neighbour, 1 = by the RIGHT
//   tail              : flag[side] = last sequence number the neighbour on that side has published; done counter; error word
struct MailHeader { int count; int pad[15]; };
struct MailInfo { unsigned long long magic; unsigned long long capFloats; unsigned long long boxBytes; unsigned char pad[232]; };
static_assert(sizeof(MailInfo) == 256, "MailInfo is one 256-byte block");
constexpr unsigned long long kMailMagic = 0x5350484b4d41494cull;   // "SPHKMAIL"

}  // namespace

SkinTrack sphk_skin_track(const sphk_ctx* c, float radius, bool on);     // sphk_sweeps.cu

struct sphk_mg_comm {
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    NcclComm comm = nullptr;
    int* dInts = nullptr;         // [4 * SPHK_MG_MAX_INTS] device staging of the small host exchanges
    int* hInts = nullptr;         // pinned twin
    double* dDouble = nullptr;
    long long bytesSent = 0, messages = 0;
    // ---- peer-memory mailboxes ----
    size_t capFloats = 0;         // payload capacity of one mailbox
    size_t boxBytes = 0;          // sizeof(MailHeader) + capFloats * 4, rounded to 256
    unsigned char* mail = nullptr;        // this rank's mailboxes + flags (+ done counters, error word)
    unsigned char* peer[2] = {nullptr, nullptr};   // neighbours' mailbox blocks, opened through IPC (0 = left, 1 = right)
    bool connected = false;
    unsigned long long seq = 0;   // halos issued through the mailboxes so far
    int transport = 0;            // 0: NCCL halos, 1: mailbox halos
};

#define SPHK_MG_MAX_INTS 8

namespace {

int nccl_rc(int r) {
    if (r == 0) return SPHK_OK;
    std::fprintf(stderr, "sphk_mg: NCCL error %d (%s)\n", r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    return SPHK_ERR_COMM;
}
#define SPHK_NCCL_TRY(expr) do { const int r_ = nccl_rc(expr); if (r_ != SPHK_OK) return r_; } while (0)
// inside ncclGroupStart/ncclGroupEnd: an error must still close the group, or every later NCCL call of the process
// would be appended to it
#define SPHK_NCCL_TRY_IN_GROUP(expr) do { const int r_ = nccl_rc(expr); if (r_ != SPHK_OK) { nccl().GroupEnd(); return r_; } } while (0)

inline size_t flags_offset(const sphk_mg_comm* m) { return sizeof(MailInfo) + 4 * m->boxBytes; }
inline unsigned char* box_of(unsigned char* base, const sphk_mg_comm* m, int side, int parity) {
    return base + sizeof(MailInfo) + (static_cast<size_t>(side) * 2 + parity) * m->boxBytes;
}

// words after the four boxes: flag[2] (ull), done[2] (uint), error (uint)
struct MailTail { unsigned long long flag[2]; unsigned int done[2]; unsigned int error; unsigned int pad; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct HaloSide {
    const float* src;        // first float of the slice to send (nullptr: no neighbour on this side)
    int sendFloats;          // floats to send
    unsigned char* peerBox;  // neighbour's mailbox for this message (remote memory)
    unsigned long long* peerFlag;   // neighbour's flag for messages from this rank (remote memory)
    const unsigned char* myBox;     // this rank's mailbox the neighbour writes (local memory)
    int ghostBegin, ghostCount;     // particles [ghostBegin, +ghostCount) receive the message
};

struct HaloArgs {
    HaloSide side[2];
    MailTail* tail;          // local
    unsigned long long seq;
    float* array;            // API array the ghost slices belong to (width floats per particle)
    int width;               // 1 or 3
    int what;                // sphk_push_range bit mask (0: array only)
    Rec rec;
    SkinTrack track;         // skin tracking (PBD position halos); track.posBuild == nullptr: off
    unsigned long long timeoutNs;
    const int* rangesDev;    // non-null: the eight plane ranges live in device memory (sphk_mg_plane_ranges); the kernel
                             // derives the slices to send and the ghost ranges from them, side[].src/ghost* hold only the
                             // "has a neighbour" information
    unsigned int capFloats;  // mailbox payload capacity (device-side check in rangesDev mode)
};

constexpr int kHaloBlock = 256;

// One halo: send both plane slices into the neighbours' mailboxes, publish, wait, unpack both ghost slices.
// The grid is at most one block per SM (all blocks co-resident), so a block that waits never keeps a block that
// still has to send from running.
__global__ void __launch_bounds__(kHaloBlock) k_halo_mailbox(HaloArgs a) {
    const int nthreads = gridDim.x * kHaloBlock;
    const int tid = blockIdx.x * kHaloBlock + threadIdx.x;
    if (a.rangesDev) {       // ranges = {first_begin, first_count, last_begin, last_count, ghostL_begin, ghostL_count, ghostR_begin, ghostR_count}
        const int* r = a.rangesDev;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (a.side[s].peerBox) {
                const int n = r[2 * s + 1] * a.width;
                a.side[s].src = a.array + static_cast<size_t>(r[2 * s]) * a.width;
                a.side[s].sendFloats = static_cast<unsigned int>(n) > a.capFloats ? -1 : n;
            }
            if (a.side[s].myBox) { a.side[s].ghostBegin = r[4 + 2 * s]; a.side[s].ghostCount = r[5 + 2 * s]; }
        }
    }
    // ---- 1. remote stores -------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const HaloSide& h = a.side[s];
        if (!h.peerBox) continue;
        float* dst = reinterpret_cast<float*>(h.peerBox + sizeof(MailHeader));
        const float* src = h.src;
        const int n = h.sendFloats;
        if (n < 0) {                               // oversize slice (host-side capacity check failed): header only
            if (tid == 0) { int* hdr = reinterpret_cast<int*>(h.peerBox); hdr[0] = -1; hdr[1] = 0; }
            continue;
        }
        // the slice starts at an arbitrary particle: peel to 16-byte alignment of the SOURCE, the mailbox payload is
        // written with the same phase (payload offset = source misalignment), so both sides move float4s
        const int lead = static_cast<int>((reinterpret_cast<uintptr_t>(src) >> 2) & 3);   // floats past a 16-byte boundary
        const float* src0 = src - lead;
        float* dst0 = dst;                        // dst[k] holds src0[k]; payload proper starts at dst + lead
        const int total = n + lead;
        const int nvec = total >> 2;
        for (int v = tid; v < nvec; v += nthreads) {
            const int k = v << 2;
            if (k >= lead && k + 4 <= total) {
                *reinterpret_cast<float4*>(dst0 + k) = __ldg(reinterpret_cast<const float4*>(src0 + k));
            } else {
                for (int j = 0; j < 4; ++j)
                    if (k + j >= lead && k + j < total) dst0[k + j] = src0[k + j];
            }
        }
        if (tid == 0) {
            for (int k = nvec << 2; k < total; ++k)
                if (k >= lead) dst0[k] = src0[k];
            int* hdr = reinterpret_cast<int*>(h.peerBox);
            hdr[0] = n; hdr[1] = lead;
        }
    }
    // ---- 2. publish: the last block to finish its stores raises both neighbours' flags ----------------------------
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&a.tail->done[0], 1u);
        if (prev == gridDim.x - 1) {
            a.tail->done[0] = 0;
            __threadfence_system();
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (a.side[s].peerFlag) st_release_sys(a.side[s].peerFlag, a.seq);
        }
    }
    // ---- 3. wait for the neighbours' messages (flags live in LOCAL memory), 4. unpack -------------------------------
    __shared__ int sOk;
    float d2 = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const HaloSide& h = a.side[s];
        if (!h.myBox) continue;
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_timer_ns();
            int ok = (*reinterpret_cast<volatile unsigned int*>(&a.tail->error) & 3u) ? 0 : 1;   // an earlier wait timed out: do not stall again
            while (ok && ld_acquire_sys(&a.tail->flag[s]) < a.seq) {
                if (global_timer_ns() - t0 > a.timeoutNs) { ok = 0; atomicOr(&a.tail->error, 1u << s); break; }
                __nanosleep(64);
            }
            sOk = ok;
        }
        __syncthreads();
        const bool ok = sOk != 0;
        __syncthreads();
        if (!ok) continue;
        const int* hdr = reinterpret_cast<const int*>(h.myBox);
        const int n = __ldcg(hdr), lead = __ldcg(hdr + 1);
        if (n != h.ghostCount * a.width) {            // ordering contract violated (or -1: the sender's slice was oversize): report, do not touch memory
            if (tid == 0) atomicOr(&a.tail->error, 4u << s);
            continue;
        }
        const float* pay = reinterpret_cast<const float*>(h.myBox + sizeof(MailHeader)) + lead;
        for (int t = tid; t < h.ghostCount; t += nthreads) {
            const int i = h.ghostBegin + t;
            if (a.width == 1) {
                const float v = __ldcg(pay + t);
                a.array[i] = v;
                if (a.what & 2) rec_set_s(a.rec + i, v);
            } else {
                const float3 v = f3(__ldcg(pay + 3 * t), __ldcg(pay + 3 * t + 1), __ldcg(pay + 3 * t + 2));
                store3(a.array, i, v);
                if (a.what & 1) rec_set_vel(a.rec + i, v);
                if (a.what & 4) {
                    rec_set_pos(a.rec + i, v);
                    if (a.track.posBuild) d2 = fmaxf(d2, skin_track(a.track, i, v));
                }
            }
        }
    }
    if ((a.what & 4) && a.track.posBuild) skin_track_max(a.track, d2);   // ghosts moved by their owner count against the list skin (as k_push_range)
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 1;
    }
    return n;
}

}  // namespace

// ---- lifetime ------------------------------------------------------------------------------------------------------
extern "C" int sphk_mg_unique_id(unsigned char id[128]) {
    if (!id) return SPHK_ERR_INVALID;
    NcclApi& n = nccl();
    if (!n.ok) return SPHK_ERR_COMM;
    NcclUid u;
    SPHK_NCCL_TRY(n.GetUniqueId(&u));
    std::memcpy(id, u.internal, 128);
    return SPHK_OK;
}

extern "C" int sphk_mg_init(sphk_mg_comm** out, int rank, int world, const unsigned char id[128], void* stream,
                            long long mailbox_floats) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world || mailbox_floats < 0) return SPHK_ERR_INVALID;
    *out = nullptr;
    NcclApi& n = nccl();
    if (!n.ok) return SPHK_ERR_COMM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHK_ERR_NO_DEVICE;
    sphk_mg_comm* m = new (std::nothrow) sphk_mg_comm();
    if (!m) return SPHK_ERR_ALLOC;
    m->rank = rank; m->world = world; m->stream = static_cast<cudaStream_t>(stream);
    NcclUid u;
    std::memcpy(u.internal, id, 128);
    int rc = nccl_rc(n.CommInitRank(&m->comm, world, u, rank));
    if (rc != SPHK_OK) { delete m; return rc; }
    if (cudaMalloc(&m->dInts, 4 * SPHK_MG_MAX_INTS * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&m->hInts, 4 * SPHK_MG_MAX_INTS * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&m->dDouble, 2 * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        sphk_mg_destroy(m);
        return SPHK_ERR_ALLOC;
    }
    if (mailbox_floats > 0) {
        m->capFloats = static_cast<size_t>(mailbox_floats);
        m->boxBytes = (sizeof(MailHeader) + (m->capFloats + 4) * sizeof(float) + 255) / 256 * 256;
        const size_t total = sizeof(MailInfo) + 4 * m->boxBytes + 256;
        if (cudaMalloc(&m->mail, total) != cudaSuccess) { cudaGetLastError(); m->mail = nullptr; sphk_mg_destroy(m); return SPHK_ERR_ALLOC; }
        cudaMemset(m->mail, 0, total);
        MailInfo info;
        std::memset(&info, 0, sizeof(info));
        info.magic = kMailMagic; info.capFloats = m->capFloats; info.boxBytes = m->boxBytes;
        cudaMemcpy(m->mail, &info, sizeof(info), cudaMemcpyHostToDevice);
        cudaDeviceSynchronize();
    }
    *out = m;
    return SPHK_OK;
}

extern "C" void sphk_mg_destroy(sphk_mg_comm* m) {
    if (!m) return;
    cudaStreamSynchronize(m->stream);
    for (int s = 0; s < 2; ++s)
        if (m->peer[s]) cudaIpcCloseMemHandle(m->peer[s]);
    if (m->comm) nccl().CommDestroy(m->comm);
    cudaFree(m->mail);
    cudaFree(m->dInts);
    cudaFreeHost(m->hInts);
    cudaFree(m->dDouble);
    delete m;
}

// ---- mailbox wiring: handles travel by whatever side channel the caller has (torch.distributed all_gather) -------------
extern "C" int sphk_mg_ipc_handle(sphk_mg_comm* m, unsigned char handle[64]) {
    if (!m || !handle || !m->mail) return SPHK_ERR_INVALID;
    cudaIpcMemHandle_t h;
    SPHK_CUDA_TRY(cudaIpcGetMemHandle(&h, m->mail));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle, &h, 64);
    return SPHK_OK;
}

extern "C" int sphk_mg_ipc_connect(sphk_mg_comm* m, const unsigned char* left64, const unsigned char* right64) {
    if (!m || !m->mail) return SPHK_ERR_INVALID;
    const unsigned char* hs[2] = {m->rank > 0 ? left64 : nullptr, m->rank < m->world - 1 ? right64 : nullptr};
    for (int s = 0; s < 2; ++s) {
        if ((s == 0 && m->rank > 0 && !left64) || (s == 1 && m->rank < m->world - 1 && !right64)) return SPHK_ERR_INVALID;
        if (!hs[s]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hs[s], 64);
        void* p = nullptr;
        SPHK_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->peer[s] = static_cast<unsigned char*>(p);
        // every address inside the neighbour's block is computed from MY boxBytes: refuse a neighbour laid out differently
        MailInfo info;
        SPHK_CUDA_TRY(cudaMemcpy(&info, p, sizeof(info), cudaMemcpyDefault));
        if (info.magic != kMailMagic || info.capFloats != m->capFloats || info.boxBytes != m->boxBytes) {
            std::fprintf(stderr, "sphk_mg: rank %d: neighbour mailbox capacity %llu floats differs from mine (%zu): every rank must "
                                 "pass the same mailbox_floats to sphk_mg_init\n", m->rank, info.capFloats, m->capFloats);
            return SPHK_ERR_INVALID;
        }
    }
    m->connected = true;
    return SPHK_OK;
}

/* 0: NCCL halos, 1: mailbox halos (needs sphk_mg_ipc_connect on every rank) */
extern "C" int sphk_mg_set_transport(sphk_mg_comm* m, int transport) {
    if (!m || transport < 0 || transport > 1) return SPHK_ERR_INVALID;
    if (transport == 1 && !(m->connected || m->world == 1)) return SPHK_ERR_STATE;
    m->transport = transport;
    return SPHK_OK;
}

// ---- small host-visible exchanges (synchronise the stream) ---------------------------------------------------------
extern "C" int sphk_mg_exchange_ints(sphk_mg_comm* m, const int* to_left, const int* to_right, int* from_left, int* from_right,
                                     int count) {
    if (!m || count < 1 || count > SPHK_MG_MAX_INTS || !to_left || !to_right || !from_left || !from_right) return SPHK_ERR_INVALID;
    NcclApi& n = nccl();
    const int K = SPHK_MG_MAX_INTS;
    for (int k = 0; k < count; ++k) { m->hInts[k] = to_left[k]; m->hInts[K + k] = to_right[k]; m->hInts[2 * K + k] = 0; m->hInts[3 * K + k] = 0; }
    SPHK_CUDA_TRY(cudaMemcpyAsync(m->dInts, m->hInts, 4 * K * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    SPHK_NCCL_TRY(n.GroupStart());
    if (m->rank > 0) {
        SPHK_NCCL_TRY_IN_GROUP(n.Send(m->dInts, count, kNcclInt32, m->rank - 1, m->comm, m->stream));
        SPHK_NCCL_TRY_IN_GROUP(n.Recv(m->dInts + 2 * K, count, kNcclInt32, m->rank - 1, m->comm, m->stream));
    }
    if (m->rank < m->world - 1) {
        SPHK_NCCL_TRY_IN_GROUP(n.Send(m->dInts + K, count, kNcclInt32, m->rank + 1, m->comm, m->stream));
        SPHK_NCCL_TRY_IN_GROUP(n.Recv(m->dInts + 3 * K, count, kNcclInt32, m->rank + 1, m->comm, m->stream));
    }
    SPHK_NCCL_TRY(n.GroupEnd());
    SPHK_CUDA_TRY(cudaMemcpyAsync(m->hInts + 2 * K, m->dInts + 2 * K, 2 * K * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(m->stream));
    for (int k = 0; k < count; ++k) { from_left[k] = m->hInts[2 * K + k]; from_right[k] = m->hInts[3 * K + k]; }
    return SPHK_OK;
}

extern "C" int sphk_mg_allreduce_sum(sphk_mg_comm* m, double* inout_host) {
    if (!m || !inout_host) return SPHK_ERR_INVALID;
    NcclApi& n = nccl();
    SPHK_CUDA_TRY(cudaMemcpyAsync(m->dDouble, inout_host, sizeof(double), cudaMemcpyHostToDevice, m->stream));
    SPHK_NCCL_TRY(n.AllReduce(m->dDouble, m->dDouble + 1, 1, kNcclFloat64, kNcclSum, m->comm, m->stream));
    SPHK_CUDA_TRY(cudaMemcpyAsync(inout_host, m->dDouble + 1, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return SPHK_OK;
}

// ---- slice exchange over NCCL: for each array a, floats [begin*width, +count*width) ---------------------------------
// send_left / send_right / recv_left / recv_right = {begin, count} in particles; one NCCL group on the stream.
extern "C" int sphk_mg_exchange_slices(sphk_mg_comm* m, int narrays, const float* const* send_arrays, float* const* recv_arrays,
                                       const int* widths, const int send_left[2], const int send_right[2],
                                       const int recv_left[2], const int recv_right[2]) {
    if (!m || narrays < 1 || !send_arrays || !recv_arrays || !widths || !send_left || !send_right || !recv_left || !recv_right)
        return SPHK_ERR_INVALID;
    NcclApi& n = nccl();
    const bool L = m->rank > 0, R = m->rank < m->world - 1;
    bool any = false;
    for (int a = 0; a < narrays; ++a) {
        const size_t w = static_cast<size_t>(widths[a]);
        if (L && send_left[1] > 0) {
            if (!any) { SPHK_NCCL_TRY(n.GroupStart()); any = true; }
            SPHK_NCCL_TRY_IN_GROUP(n.Send(send_arrays[a] + send_left[0] * w, send_left[1] * w, kNcclFloat32, m->rank - 1, m->comm, m->stream));
            m->bytesSent += static_cast<long long>(send_left[1] * w * 4); m->messages++;
        }
        if (L && recv_left[1] > 0) {
            if (!any) { SPHK_NCCL_TRY(n.GroupStart()); any = true; }
            SPHK_NCCL_TRY_IN_GROUP(n.Recv(recv_arrays[a] + recv_left[0] * w, recv_left[1] * w, kNcclFloat32, m->rank - 1, m->comm, m->stream));
        }
        if (R && send_right[1] > 0) {
            if (!any) { SPHK_NCCL_TRY(n.GroupStart()); any = true; }
            SPHK_NCCL_TRY_IN_GROUP(n.Send(send_arrays[a] + send_right[0] * w, send_right[1] * w, kNcclFloat32, m->rank + 1, m->comm, m->stream));
            m->bytesSent += static_cast<long long>(send_right[1] * w * 4); m->messages++;
        }
        if (R && recv_right[1] > 0) {
            if (!any) { SPHK_NCCL_TRY(n.GroupStart()); any = true; }
            SPHK_NCCL_TRY_IN_GROUP(n.Recv(recv_arrays[a] + recv_right[0] * w, recv_right[1] * w, kNcclFloat32, m->rank + 1, m->comm, m->stream));
        }
    }
    if (any) SPHK_NCCL_TRY(n.GroupEnd());
    return SPHK_OK;
}

// ---- one halo: first / last owned plane -> the neighbours' ghost planes, then into the packed records ---------------------
// ranges = {first_begin, first_count, last_begin, last_count, ghostL_begin, ghostL_count, ghostR_begin, ghostR_count}
// what: sphk_push_range mask (1 vel: array = scene->fluid.vel, width 3; 2 scalar: width 1; 4 pos: array =
// scene->fluid.pos, width 3) or 0 for an array no record mirrors (colour gradient, density, pressure).
static int halo_impl(sphk_mg_comm* m, sphk_ctx* c, const sphk_scene* s, int what, float* array, int width, const int ranges[8],
                     const int* rangesDev, int maxParticles);

extern "C" int sphk_mg_halo(sphk_mg_comm* m, sphk_ctx* c, const sphk_scene* s, int what, float* array, int width, const int ranges[8]) {
    if (!ranges) return SPHK_ERR_INVALID;
    return halo_impl(m, c, s, what, array, width, ranges, nullptr, 0);
}

/* the same with the eight plane ranges in DEVICE memory (written by sphk_mg_plane_ranges; the host never reads them):
 * mailbox transport only.  max_particles bounds the size of a plane (grid sizing). */
extern "C" int sphk_mg_halo_device(sphk_mg_comm* m, sphk_ctx* c, const sphk_scene* s, int what, float* array, int width,
                                   const int* device_ranges8, int max_particles) {
    if (!device_ranges8 || max_particles < 0) return SPHK_ERR_INVALID;
    if (!m || m->transport != 1) return SPHK_ERR_STATE;
    static const int zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return halo_impl(m, c, s, what, array, width, zero, device_ranges8, max_particles);
}

static int halo_impl(sphk_mg_comm* m, sphk_ctx* c, const sphk_scene* s, int what, float* array, int width, const int ranges[8],
                     const int* rangesDev, int maxParticles) {
    if (!m || !c || !s || !array || (width != 1 && width != 3) || (what != 0 && what != 1 && what != 2 && what != 4))
        return SPHK_ERR_INVALID;
    if ((what == 2 && width != 1) || ((what == 1 || what == 4) && width != 3)) return SPHK_ERR_INVALID;
    for (int k = 0; k < 4; ++k)
        if (ranges[2 * k] < 0 || ranges[2 * k + 1] < 0 || ranges[2 * k] + ranges[2 * k + 1] > c->nF) return SPHK_ERR_INVALID;
    if (c->stream != m->stream) return SPHK_ERR_INVALID;
    const bool L = m->rank > 0, R = m->rank < m->world - 1;
    if (!L && !R) return SPHK_OK;
    if (m->transport == 0) {
        const float* sa[1] = {array};
        float* ra[1] = {array};
        const int w[1] = {width};
        int rc = sphk_mg_exchange_slices(m, 1, sa, ra, w, ranges + 0, ranges + 2, ranges + 4, ranges + 6);
        if (rc != SPHK_OK) return rc;
        if (what) {
            if (L && ranges[5] > 0) { rc = sphk_push_range(c, s, what, array, ranges[4], ranges[5]); if (rc != SPHK_OK) return rc; }
            if (R && ranges[7] > 0) { rc = sphk_push_range(c, s, what, array, ranges[6], ranges[7]); if (rc != SPHK_OK) return rc; }
        }
        return SPHK_OK;
    }
    // ---- mailbox transport ----
    if (!m->connected) return SPHK_ERR_STATE;
    // a slice that does not fit a mailbox: the message is still sent (header only, count -1) so that the neighbours'
    // sequence numbers stay in step and they see a size mismatch at once instead of a 20 s timeout
    const bool tooBig = static_cast<size_t>(ranges[1]) * width > m->capFloats || static_cast<size_t>(ranges[3]) * width > m->capFloats ||
                        static_cast<size_t>(ranges[5]) * width > m->capFloats || static_cast<size_t>(ranges[7]) * width > m->capFloats;
    const unsigned long long seq = ++m->seq;
    const int parity = static_cast<int>(seq & 1);
    MailTail* tail = reinterpret_cast<MailTail*>(m->mail + flags_offset(m));
    HaloArgs a;
    std::memset(&a, 0, sizeof(a));
    // to the LEFT neighbour I am its RIGHT side (side index 1 in its block), and vice versa
    if (L) {
        MailTail* ptail = reinterpret_cast<MailTail*>(m->peer[0] + flags_offset(m));
        a.side[0].src = array + static_cast<size_t>(ranges[0]) * width;
        a.side[0].sendFloats = tooBig ? -1 : ranges[1] * width;
        a.side[0].peerBox = box_of(m->peer[0], m, 1, parity);
        a.side[0].peerFlag = &ptail->flag[1];
        a.side[0].myBox = box_of(m->mail, m, 0, parity);
        a.side[0].ghostBegin = ranges[4]; a.side[0].ghostCount = ranges[5];
        m->bytesSent += static_cast<long long>(ranges[1]) * width * 4; m->messages++;
    }
    if (R) {
        MailTail* ptail = reinterpret_cast<MailTail*>(m->peer[1] + flags_offset(m));
        a.side[1].src = array + static_cast<size_t>(ranges[2]) * width;
        a.side[1].sendFloats = tooBig ? -1 : ranges[3] * width;
        a.side[1].peerBox = box_of(m->peer[1], m, 0, parity);
        a.side[1].peerFlag = &ptail->flag[0];
        a.side[1].myBox = box_of(m->mail, m, 1, parity);
        a.side[1].ghostBegin = ranges[6]; a.side[1].ghostCount = ranges[7];
        m->bytesSent += static_cast<long long>(ranges[3]) * width * 4; m->messages++;
    }
    a.tail = tail;
    a.seq = seq;
    a.array = array; a.width = width; a.what = what;
    a.rec = c->rec;
    const bool track = (what & 4) && c->listHasSkin && c->listEpoch == c->searchEpoch;
    a.track = sphk_skin_track(c, s->radius, track);
    a.timeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
    a.rangesDev = rangesDev;
    a.capFloats = static_cast<unsigned int>(m->capFloats);
    int work = ranges[1] > ranges[3] ? ranges[1] : ranges[3];
    if (ranges[5] > work) work = ranges[5];
    if (ranges[7] > work) work = ranges[7];
    if (rangesDev) work = maxParticles;
    int blocks = (work * width / 4 + kHaloBlock - 1) / kHaloBlock;
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count()) blocks = sm_count();
    k_halo_mailbox<<<blocks, kHaloBlock, 0, c->stream>>>(a);
    c->launches++;
    if (what & 4) c->posDirty = true;
    SPHK_CUDA_TRY(cudaGetLastError());
    return tooBig ? SPHK_ERR_CAPACITY : SPHK_OK;
}

// ---- host-free step bookkeeping -----------------------------------------------------------------------------------------
// Everything a slab rank derives from the cell ranges of its freshly sorted set, computed where the data is:
//   out[0..7]   plane offsets s0, s1, s2, s3, s_{w-1}, s_w, s_{w+1}, s_end (slabs.plane_ranges)
//   out[8..15]  halo ranges {first_begin, first_count, last_begin, last_count, ghostL_begin, ghostL_count, ghostR_begin, ghostR_count}
//   out[16..17] active (owned) range {begin, count}                                     -> sphk_set_active_range_device
//   out[18..21] next step's candidates {to_left_begin, to_left_count, to_right_begin, to_right_count}
//   out[22..23] what each neighbour has to know: {to_left_count, to_right_count}         -> sphk_mg_exchange_ints_async
__global__ void k_plane_ranges(const int* __restrict__ cs, int planeCells, int w, int* __restrict__ out) {
    const int s0 = cs[0], s1 = cs[planeCells], s2 = cs[2 * planeCells], s3 = cs[3 * planeCells];
    const int swm1 = cs[(w - 1 > 0 ? w - 1 : 0) * planeCells], sw = cs[w * planeCells], sw1 = cs[(w + 1) * planeCells],
              send = cs[(w + 2) * planeCells];
    out[0] = s0; out[1] = s1; out[2] = s2; out[3] = s3; out[4] = swm1; out[5] = sw; out[6] = sw1; out[7] = send;
    const int fb = s1, fe = (w >= 2) ? s2 : sw1;           // first owned plane
    const int lb = (w >= 2) ? sw : s1, le = sw1;           // last owned plane
    out[8] = fb; out[9] = fe - fb; out[10] = lb; out[11] = le - lb;
    out[12] = s0; out[13] = s1 - s0; out[14] = sw1; out[15] = send - sw1;
    out[16] = s1; out[17] = sw1 - s1;
    const int tlb = s1, tle = (w >= 2) ? (s3 < sw1 ? s3 : sw1) : sw1;
    const int trb = (w >= 2) ? (swm1 > s1 ? swm1 : s1) : s1, tre = sw1;
    out[18] = tlb; out[19] = tle - tlb; out[20] = trb; out[21] = tre - trb;
    out[22] = tle - tlb; out[23] = tre - trb;
}

extern "C" int sphk_mg_plane_ranges(sphk_mg_comm* m, const int* cell_start_fluid, int plane_cells, int w, int* device_out24,
                                    int* pinned_host_out24) {
    if (!m || !cell_start_fluid || plane_cells <= 0 || w < 1 || !device_out24) return SPHK_ERR_INVALID;
    k_plane_ranges<<<1, 1, 0, m->stream>>>(cell_start_fluid, plane_cells, w, device_out24);
    if (pinned_host_out24)
        SPHK_CUDA_TRY(cudaMemcpyAsync(pinned_host_out24, device_out24, 24 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

/* `count` ints to each neighbour and back, all on the stream and WITHOUT a host synchronisation: the values to send are
 * read from device memory, the received ones land in pinned host memory once the stream reaches the copy (record an event
 * after this call and wait for it before reading them).  A missing neighbour yields zeros. */
extern "C" int sphk_mg_exchange_ints_async(sphk_mg_comm* m, const int* device_to_left, const int* device_to_right, int count,
                                           int* pinned_from_left, int* pinned_from_right) {
    if (!m || count < 1 || count > SPHK_MG_MAX_INTS || !device_to_left || !device_to_right || !pinned_from_left || !pinned_from_right)
        return SPHK_ERR_INVALID;
    NcclApi& n = nccl();
    const int K = SPHK_MG_MAX_INTS;
    SPHK_CUDA_TRY(cudaMemsetAsync(m->dInts + 2 * K, 0, 2 * K * sizeof(int), m->stream));
    SPHK_NCCL_TRY(n.GroupStart());
    if (m->rank > 0) {
        SPHK_NCCL_TRY_IN_GROUP(n.Send(device_to_left, count, kNcclInt32, m->rank - 1, m->comm, m->stream));
        SPHK_NCCL_TRY_IN_GROUP(n.Recv(m->dInts + 2 * K, count, kNcclInt32, m->rank - 1, m->comm, m->stream));
    }
    if (m->rank < m->world - 1) {
        SPHK_NCCL_TRY_IN_GROUP(n.Send(device_to_right, count, kNcclInt32, m->rank + 1, m->comm, m->stream));
        SPHK_NCCL_TRY_IN_GROUP(n.Recv(m->dInts + 3 * K, count, kNcclInt32, m->rank + 1, m->comm, m->stream));
    }
    SPHK_NCCL_TRY(n.GroupEnd());
    SPHK_CUDA_TRY(cudaMemcpyAsync(pinned_from_left, m->dInts + 2 * K, count * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    SPHK_CUDA_TRY(cudaMemcpyAsync(pinned_from_right, m->dInts + 3 * K, count * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    return SPHK_OK;
}

/* the mailbox error word into pinned host memory, no synchronisation (see sphk_mg_check for the bits) */
extern "C" int sphk_mg_check_async(sphk_mg_comm* m, int* pinned_error_bits) {
    if (!m || !pinned_error_bits) return SPHK_ERR_INVALID;
    if (!m->mail) { *pinned_error_bits = 0; return SPHK_OK; }
    MailTail* tail = reinterpret_cast<MailTail*>(m->mail + flags_offset(m));
    SPHK_CUDA_TRY(cudaMemcpyAsync(pinned_error_bits, &tail->error, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    return SPHK_OK;
}

/* Reads the mailbox error word (synchronises the stream): 0 = fine; bit 0/1: timed out waiting for the left/right
 * neighbour; bit 2/3: the message from the left/right neighbour did not match the ghost range. */
extern "C" int sphk_mg_check(sphk_mg_comm* m, int* error_bits_host) {
    if (!m || !error_bits_host) return SPHK_ERR_INVALID;
    *error_bits_host = 0;
    if (!m->mail) return SPHK_OK;
    MailTail* tail = reinterpret_cast<MailTail*>(m->mail + flags_offset(m));
    unsigned int e = 0;
    SPHK_CUDA_TRY(cudaMemcpyAsync(&e, &tail->error, sizeof(e), cudaMemcpyDeviceToHost, m->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(m->stream));
    *error_bits_host = static_cast<int>(e);
    return SPHK_OK;
}

// ---- strays: owned particles that crossed two or more cell planes in x since the last search ----------------------------------
// The candidate exchange covers particles that move at most ONE plane per step (two candidate planes per side).  A violent
// impact produces a few hundred particles that do not (the reference's own DFSPH setting of BASELINE.json: 4 + 4 fixed
// iterations, dt = 0.004, shoots particles across tens of planes once the block hits the floor, around step 40 of the dam
// break).  They are taken out of the regular flow and routed to everybody instead:
//   collect : every owned particle compares the plane its sorted slot belongs to (binary search in the plane offsets of the
//             cell ranges) with the plane of its position now (the search's own hash); |difference| >= 2 -> its row of the carried
//             arrays goes into this rank's block and its position leaves the world, so whoever holds a copy of it (its owner,
//             a neighbour that receives it as a candidate) drops it at the next search
//   gather  : the fixed-size blocks of all ranks, one all-gather
//   append  : world * capacity slots behind the assembled set [left candidates | own | right candidates]: the strays of rank
//             0, 1, ... in block order -- the same sequence on every rank, so the ordering contract of the halos holds for
//             them as well -- unused slots out of the world.  The search keeps a stray on the rank(s) whose window it landed in.
namespace {
constexpr int kStrayHeader = 4;                 // floats: {int count, 3 x pad}
constexpr float kOutOfWorld = -1.0e6f;
struct StrayArrays { float* a[4]; int w[4]; int off[4]; int k; int stride; };

bool stray_arrays(StrayArrays& A, int narrays, float* const* arrays, const int* widths) {
    if (narrays < 1 || narrays > 4 || !arrays || !widths || widths[0] != 3) return false;
    A.k = narrays; A.stride = 0;
    for (int i = 0; i < narrays; ++i) {
        if (!arrays[i] || widths[i] < 1 || widths[i] > 3) return false;
        A.a[i] = arrays[i]; A.w[i] = widths[i]; A.off[i] = A.stride; A.stride += widths[i];
    }
    return true;
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_strays_collect(const int* __restrict__ cs, int planeCells, int nPlanes, int begin, int count, float cellLength, int orgX,
                 StrayArrays A, float* __restrict__ block, int capacity) {
    const int t = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (t >= count) return;
    const int i = begin + t;
    const int pNow = cell_coord(A.a[0][3 * static_cast<size_t>(i)], cellLength) - orgX;
    int lo = 0, hi = nPlanes;                   // the plane slot i was sorted into: largest p with cs[p * planeCells] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cs + static_cast<size_t>(mid) * planeCells) <= i) lo = mid; else hi = mid;
    }
    const int d = pNow - lo;
    if (d > -2 && d < 2) return;
    const int slot = atomicAdd(reinterpret_cast<int*>(block), 1);
    if (slot >= capacity) return;               // stays in place; count > capacity is reported by the append
    float* row = block + kStrayHeader + static_cast<size_t>(slot) * A.stride;
    for (int a = 0; a < A.k; ++a)
        for (int c = 0; c < A.w[a]; ++c) row[A.off[a] + c] = A.a[a][static_cast<size_t>(A.w[a]) * i + c];
    A.a[0][3 * static_cast<size_t>(i)] = kOutOfWorld; A.a[0][3 * static_cast<size_t>(i) + 1] = kOutOfWorld;
    A.a[0][3 * static_cast<size_t>(i) + 2] = kOutOfWorld;
}

__global__ void __launch_bounds__(SPHK_BLOCK)
k_strays_append(const float* __restrict__ gathered, int world, int capacity, StrayArrays A, int dstBegin, unsigned int* errorWord) {
    const int t = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (t >= world * capacity) return;
    const int r = t / capacity, j = t - r * capacity;
    const float* block = gathered + static_cast<size_t>(r) * (kStrayHeader + static_cast<size_t>(capacity) * A.stride);
    const int cnt = *reinterpret_cast<const int*>(block);
    const size_t dst = static_cast<size_t>(dstBegin) + t;
    if (j < cnt) {
        const float* row = block + kStrayHeader + static_cast<size_t>(j) * A.stride;
        for (int a = 0; a < A.k; ++a)
            for (int c = 0; c < A.w[a]; ++c) A.a[a][A.w[a] * dst + c] = row[A.off[a] + c];
    } else {
        for (int a = 0; a < A.k; ++a)
            for (int c = 0; c < A.w[a]; ++c) A.a[a][A.w[a] * dst + c] = a == 0 ? kOutOfWorld : 0.0f;
    }
    if (j == 0 && cnt > capacity && errorWord) atomicOr(errorWord, 16u);
}
}  // namespace

extern "C" long long sphk_strays_block_floats(int capacity, int narrays, const int* widths) {
    if (capacity < 0 || narrays < 1 || narrays > 4 || !widths) return -1;
    long long stride = 0;
    for (int i = 0; i < narrays; ++i) stride += widths[i];
    return kStrayHeader + static_cast<long long>(capacity) * stride;
}

extern "C" int sphk_strays_collect(sphk_ctx* c, const int* cell_start_fluid, int own_begin, int own_count, int narrays,
                                   float* const* arrays, const int* widths, float* block, int capacity) {
    StrayArrays A;
    if (!c || !cell_start_fluid || !block || capacity < 1 || own_begin < 0 || own_count < 0 || !stray_arrays(A, narrays, arrays, widths))
        return SPHK_ERR_INVALID;
    SPHK_CUDA_TRY(cudaMemsetAsync(block, 0, kStrayHeader * sizeof(float), c->stream));
    if (own_count > 0) {
        k_strays_collect<<<sphk_blocks(own_count), SPHK_BLOCK, 0, c->stream>>>(cell_start_fluid, c->cs.y * c->cs.z, c->cs.x, own_begin,
                                                                              own_count, c->cellLength, c->org.x, A, block, capacity);
        c->launches++;
    }
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

static int strays_append(sphk_ctx* c, const float* gathered, int world, int capacity, int narrays, float* const* arrays,
                         const int* widths, int dst_begin, unsigned int* errorWord) {
    StrayArrays A;
    if (!c || !gathered || world < 1 || capacity < 1 || dst_begin < 0 || !stray_arrays(A, narrays, arrays, widths)) return SPHK_ERR_INVALID;
    k_strays_append<<<sphk_blocks(world * capacity), SPHK_BLOCK, 0, c->stream>>>(gathered, world, capacity, A, dst_begin, errorWord);
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_strays_append(sphk_ctx* c, const float* gathered, int world, int capacity, int narrays, float* const* arrays,
                                  const int* widths, int dst_begin) {
    return strays_append(c, gathered, world, capacity, narrays, arrays, widths, dst_begin, nullptr);
}

/* all-gather of this rank's block + append (overflow raises bit 4 of the error word read by sphk_mg_check) */
extern "C" int sphk_mg_strays_route(sphk_mg_comm* m, sphk_ctx* c, const float* block, float* gathered, int capacity, int narrays,
                                    float* const* arrays, const int* widths, int dst_begin) {
    if (!m || !c || !block || !gathered || m->stream != c->stream) return SPHK_ERR_INVALID;
    const long long floats = sphk_strays_block_floats(capacity, narrays, widths);
    if (floats < 0) return SPHK_ERR_INVALID;
    SPHK_NCCL_TRY(nccl().AllGather(block, gathered, static_cast<size_t>(floats), kNcclFloat32, m->comm, m->stream));
    m->bytesSent += floats * 4; m->messages++;
    unsigned int* err = nullptr;
    if (m->mail) err = &reinterpret_cast<MailTail*>(m->mail + flags_offset(m))->error;
    return strays_append(c, gathered, m->world, capacity, narrays, arrays, widths, dst_begin, err);
}

/* {count of rank 0, count of rank 1, ...} of an all-gathered set of blocks (synchronises the stream): introspection / tests */
extern "C" int sphk_strays_counts(sphk_ctx* c, const float* gathered, int world, int capacity, int narrays, const int* widths, int* counts_host) {
    if (!c || !gathered || !counts_host || world < 1) return SPHK_ERR_INVALID;
    const long long floats = sphk_strays_block_floats(capacity, narrays, widths);
    if (floats < 0) return SPHK_ERR_INVALID;
    SPHK_CUDA_TRY(cudaMemcpy2DAsync(counts_host, sizeof(int), gathered, static_cast<size_t>(floats) * sizeof(float), sizeof(int),
                                    static_cast<size_t>(world), cudaMemcpyDeviceToHost, c->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SPHK_OK;
}

extern "C" int sphk_mg_stats(const sphk_mg_comm* m, long long out_host[2]) {
    if (!m || !out_host) return SPHK_ERR_INVALID;
    out_host[0] = m->bytesSent; out_host[1] = m->messages;
    return SPHK_OK;
}

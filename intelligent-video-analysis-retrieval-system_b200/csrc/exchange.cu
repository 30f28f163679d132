// exchange.cu -- K7: the cross-shard hit exchange over NVLink / NVSwitch peer memory.
//
// Every rank owns a MAILBOX in its own HBM: `slots` regions of [world][capacity] 64-bit candidate keys plus one
// 32-bit arrival flag per (slot, sender).  After its local search a rank PUSHES its [nq, k] keys straight into
// every peer's mailbox with 16-byte peer stores (one small kernel: each key is read once and written `world`
// times), fences, and the last CTA publishes the step number in the peers' flags.  The receiver does not spin on
// an SM: the arrival wait is a stream memory operation (cuStreamWaitValue32, >=) queued in front of the k-way merge
// kernel, which then reads local memory only.
//
// Why not the NCCL all-gather this replaces: a collective kernel waits for the slowest rank while holding SMs, and
// the persistent scoring kernels need all 148 of them (384 threads x 168 registers fill an SM's register file), so
// the exchange of search i could not overlap the scoring of search i+1.  With pushes + stream waits a fast GPU goes
// on scoring the next batch while a slow peer finishes the previous one; per-step straggler waits no longer add up.
//
// Slot reuse needs no acknowledgement with >= 2 slots when every rank issues push / wait / merge of consecutive
// searches in ONE stream: A's push(i+2) follows A's wait(i+1), which needs B's push(i+1), which B's stream orders
// after B's merge(i) -- the last reader of the slot A is about to overwrite.
//
// Semantic precedent in the reference: _search_with_remote_index concatenates the per-shard hit lists
// (system.py:1721-1746); here the "concatenate" is the mailbox and the sort + truncate is merge_kernel.
#include "common.cuh"

#include <cuda.h>
#include <string.h>

namespace ivr {

constexpr int kMaxRanks = 64;                       // merge_kernel folds at most 64 lists in one level

struct PeerTable {
    uint64_t* dst[kMaxRanks];                       // where MY keys go in every peer's slot
    uint32_t* flag[kMaxRanks];                      // my arrival flag in every peer's mailbox
};

__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint64_t* __restrict__ src, int64_t n, PeerTable t, int world, uint32_t epoch,
                     unsigned int* ticket, int vec16) {
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    if (vec16) {                                    // n even, source and every destination part 16-byte aligned
        const int64_t n2 = n >> 1;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        int64_t i = tid;
        for (; i + 3 * nthr < n2; i += 4 * nthr) {  // four independent loads in flight per thread
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = s4[i + u * nthr];
            for (int p = 0; p < world; ++p) {
                uint4* d = reinterpret_cast<uint4*>(t.dst[p]);
#pragma unroll
                for (int u = 0; u < 4; ++u) d[i + u * nthr] = v[u];
            }
        }
        for (; i < n2; i += nthr) {
            const uint4 v = s4[i];
            for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(t.dst[p])[i] = v;
        }
    } else {                                        // odd nq * k (tiny searches): 8-byte stores
        for (int64_t i = tid; i < n; i += nthr) {
            const uint64_t v = src[i];
            for (int p = 0; p < world; ++p) t.dst[p][i] = v;
        }
    }
    // all my stores are performed system-wide before my CTA counts itself; the last CTA publishes the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ticket, 1u);
        if (done == gridDim.x - 1) {
            *ticket = 0;                            // next push (stream-ordered behind this kernel)
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.flag[p]), "r"(epoch) : "memory");
        }
    }
}

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static WaitValue32Fn get_wait_fn() {
    static WaitValue32Fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WaitValue32Fn>(p);
        cudaGetLastError();
    }
    return fn;
}

}  // namespace ivr

using namespace ivr;

struct ivr_exchange {
    int device = 0, rank = 0, world = 1, slots = 2;
    int64_t capacity = 0;                           // keys per (slot, sender), even
    char* base = nullptr;                           // my mailbox (cudaMalloc: exportable through CUDA IPC)
    size_t bytes = 0;
    unsigned int* ticket = nullptr;                 // last-CTA-done counter of the push kernel
    char* peer[kMaxRanks] = {};                     // every rank's mailbox as mapped HERE (peer[rank] == base)
    bool  ipc_opened[kMaxRanks] = {};
    bool  connected = false;

    size_t flags_offset() const { return static_cast<size_t>(slots) * world * capacity * sizeof(uint64_t); }
    uint64_t* part(char* b, int slot, int sender, int64_t n) const {   // parts of one search are n entries apart
        return reinterpret_cast<uint64_t*>(b) + static_cast<size_t>(slot) * world * capacity + static_cast<size_t>(sender) * n;
    }
    uint32_t* flag(char* b, int slot, int sender) const {
        return reinterpret_cast<uint32_t*>(b + flags_offset()) + slot * world + sender;
    }
};

extern "C" {

int ivr_exchange_create(int device, int rank, int world, int slots, int64_t capacity, ivr_exchange** out) {
    if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || slots < 2 || slots > 8 || capacity < 1) {
        set_error("exchange_create: bad argument (world 1..%d, slots 2..8, capacity >= 1)", kMaxRanks);
        return IVR_EINVAL;
    }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available; this library has no CPU fallback");
        return IVR_ENODEVICE;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return IVR_EINVAL; }
    IVR_CUDA(cudaSetDevice(device));
    ivr_exchange* ex = new ivr_exchange();
    ex->device = device; ex->rank = rank; ex->world = world; ex->slots = slots;
    ex->capacity = (capacity + 1) & ~int64_t(1);
    ex->bytes = ex->flags_offset() + static_cast<size_t>(slots) * world * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->base), ex->bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ex->ticket), sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ex->base, 0, ex->bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->ticket, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("exchange_create: %s", cudaGetErrorString(e));
        cudaFree(ex->base); cudaFree(ex->ticket); cudaGetLastError();
        delete ex;
        return e == cudaErrorMemoryAllocation ? IVR_ENOMEM : IVR_ECUDA;
    }
    ex->peer[rank] = ex->base;
    ex->connected = world == 1;
    *out = ex;
    return IVR_OK;
}

int ivr_exchange_destroy(ivr_exchange* ex) {
    if (!ex) return IVR_OK;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < ex->world; ++p)
        if (ex->ipc_opened[p]) cudaIpcCloseMemHandle(ex->peer[p]);
    cudaFree(ex->base);
    cudaFree(ex->ticket);
    cudaGetLastError();
    delete ex;
    return IVR_OK;
}

void* ivr_exchange_base(ivr_exchange* ex) { return ex ? ex->base : nullptr; }

int ivr_exchange_ipc_handle(ivr_exchange* ex, uint8_t* handle) {
    if (!ex || !handle) { set_error("exchange_ipc_handle: bad argument"); return IVR_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == IVR_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    IVR_CUDA(cudaSetDevice(ex->device));
    cudaIpcMemHandle_t h;
    IVR_CUDA(cudaIpcGetMemHandle(&h, ex->base));
    memcpy(handle, &h, sizeof(h));
    return IVR_OK;
}

int ivr_exchange_connect_ipc(ivr_exchange* ex, const uint8_t* handles) {
    if (!ex || !handles) { set_error("exchange_connect_ipc: bad argument"); return IVR_EINVAL; }
    IVR_CUDA(cudaSetDevice(ex->device));
    for (int p = 0; p < ex->world; ++p) {
        if (p == ex->rank || ex->ipc_opened[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + static_cast<size_t>(p) * IVR_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("exchange_connect_ipc: cannot map the mailbox of rank %d (%s): no peer access between the GPUs, "
                      "or the ranks do not share an IPC namespace", p, cudaGetErrorString(e));
            return IVR_EUNSUPPORTED;
        }
        ex->peer[p] = static_cast<char*>(ptr);
        ex->ipc_opened[p] = true;
    }
    ex->connected = true;
    return IVR_OK;
}

int ivr_exchange_connect_ptrs(ivr_exchange* ex, void* const* bases) {
    if (!ex || !bases) { set_error("exchange_connect_ptrs: bad argument"); return IVR_EINVAL; }
    for (int p = 0; p < ex->world; ++p) {
        if (p == ex->rank) continue;
        if (!bases[p]) { set_error("exchange_connect_ptrs: NULL mailbox for rank %d", p); return IVR_EINVAL; }
        ex->peer[p] = static_cast<char*>(bases[p]);
    }
    ex->connected = true;
    return IVR_OK;
}

int ivr_exchange_push(ivr_exchange* ex, const uint64_t* keys_dev, int64_t n_entries, int slot, uint32_t epoch,
                      void* stream) {
    if (!ex || !keys_dev || n_entries < 1 || slot < 0 || slot >= ex->slots || epoch == 0) {
        set_error("exchange_push: bad argument");
        return IVR_EINVAL;
    }
    if (!ex->connected) { set_error("exchange_push: the peers' mailboxes are not connected yet"); return IVR_EINVAL; }
    if (n_entries > ex->capacity) {
        set_error("exchange_push: %lld keys exceed the mailbox capacity of %lld", static_cast<long long>(n_entries),
                  static_cast<long long>(ex->capacity));
        return IVR_EINVAL;
    }
    IVR_CUDA(cudaSetDevice(ex->device));
    // parts of one search lie n_entries apart (the layout ivr_topk_merge_keys_device reads); with an even count
    // every part is 16-byte aligned (capacity is even) and the copy uses 16-byte peer stores
    const int64_t stride = n_entries;
    const int vec16 = (n_entries & 1) == 0 && (reinterpret_cast<uintptr_t>(keys_dev) & 15) == 0;
    PeerTable t{};
    for (int p = 0; p < ex->world; ++p) {
        t.dst[p] = ex->part(ex->peer[p], slot, ex->rank, stride);
        t.flag[p] = ex->flag(ex->peer[p], slot, ex->rank);
    }
    const int64_t pairs = (n_entries + 1) / 2;
    int grid = static_cast<int>((pairs + 256 * 4 - 1) / (256 * 4));
    grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
    exchange_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(keys_dev, n_entries, t, ex->world, epoch,
                                                                              ex->ticket, vec16);
    IVR_CUDA(cudaGetLastError());
    return IVR_OK;
}

int ivr_exchange_wait(ivr_exchange* ex, int slot, uint32_t epoch, void* stream) {
    if (!ex || slot < 0 || slot >= ex->slots || epoch == 0) { set_error("exchange_wait: bad argument"); return IVR_EINVAL; }
    WaitValue32Fn wait = get_wait_fn();
    if (!wait) { set_error("cuStreamWaitValue32 is not available from the driver"); return IVR_EUNSUPPORTED; }
    IVR_CUDA(cudaSetDevice(ex->device));
    for (int p = 0; p < ex->world; ++p) {
        const CUresult r = wait(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(ex->flag(ex->base, slot, p)),
                                epoch, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) {
            set_error("cuStreamWaitValue32 failed with CUresult %d", static_cast<int>(r));
            return IVR_ECUDA;
        }
    }
    return IVR_OK;
}

const uint64_t* ivr_exchange_slot(ivr_exchange* ex, int slot) {
    if (!ex || slot < 0 || slot >= ex->slots) return nullptr;
    return ex->part(ex->base, slot, 0, 0);
}

}  // extern "C"

// common.cuh -- shared helpers for the sm_100a retrieval kernels.
//
// Candidate encoding used by every top-k stage: one 64-bit key per (score,row)
//     key = order_preserving(score) << 32 | ~row
// so that a plain unsigned DESCENDING order is "score descending, then row id
// ascending".  key 0 is the padding value (lower than any real candidate).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ivr_b200.h"

namespace ivr {

// ---------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);   // capi.cu (thread-local message)

#define IVR_CUDA(expr)                                                               \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            ivr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                      \
            return (_e == cudaErrorMemoryAllocation) ? IVR_ENOMEM : IVR_ECUDA;       \
        }                                                                            \
    } while (0)

#define IVR_TRY(expr)                  \
    do {                               \
        int _r = (expr);               \
        if (_r != IVR_OK) return _r;   \
    } while (0)

constexpr int kDimAlign = 64;          // rows are padded to a multiple of 64 fp16 (one 128B swizzle atom)
inline int pad_dim(int d) { return (d + kDimAlign - 1) / kDimAlign * kDimAlign; }

// list capacity (entries) used by the streaming selectors for a given k:
// C = 2*KCAP with KCAP = k rounded up to a power of two >= 128.
inline int kcap_for(int k) { int c = 128; while (c < k) c <<= 1; return c; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(f2ord(score)) << 32) | static_cast<uint64_t>(~row);
}
__device__ __forceinline__ float key_score(uint64_t key) { return ord2f(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~static_cast<uint32_t>(key); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_nc_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// two packed fp16 -> two fp32 (exact)
__device__ __forceinline__ float2 h2f2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Warp-wide bitonic sort, DESCENDING, of 32*E keys held E per lane in blocked
// order (global position g = lane*E + j).  Input order is irrelevant.
template <int E>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&v)[E], int lane) {
#pragma unroll
    for (int s = 2; s <= 32 * E; s <<= 1) {
#pragma unroll
        for (int t = s >> 1; t >= 1; t >>= 1) {
            if (t >= E) {                       // partner lives in another lane
                const int lm = t / E;
                const bool lower = (lane & lm) == 0;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const int g = lane * E + j;
                    const bool desc = (g & s) == 0;
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v[j], lm);
                    const uint64_t mx = umax64(v[j], o), mn = umin64(v[j], o);
                    v[j] = (lower == desc) ? mx : mn;
                }
            } else {                            // partner is in this lane
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const int p = j ^ t;
                    if (p > j) {
                        const int g = lane * E + j;
                        const bool desc = (g & s) == 0;
                        const uint64_t a = v[j], b = v[p];
                        const uint64_t mx = umax64(a, b), mn = umin64(a, b);
                        v[j] = desc ? mx : mn;
                        v[p] = desc ? mn : mx;
                    }
                }
            }
        }
    }
}

// Warp-cooperative compaction of a candidate list (capacity 32*E) to its k best.
// On return list[0..min(cnt,k)) holds the survivors sorted descending and the
// function returns the new admission threshold: the k-th best score if the list
// held at least k entries, else -inf.  All 32 lanes must call it.
template <int E>
__device__ __forceinline__ float warp_compact_topk(uint64_t* list, int cnt, int k, int lane) {
    uint64_t v[E];
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int g = j * 32 + lane;            // coalesced read; order is irrelevant
        v[j] = (g < cnt) ? list[g] : 0ull;
    }
    warp_sort_desc<E>(v, lane);
    __syncwarp();
    uint64_t kth = 0;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int g = lane * E + j;
        if (g < k && g < cnt) list[g] = v[j];
        if (g == k - 1) kth = v[j];
    }
    kth = __shfl_sync(0xffffffffu, kth, (k - 1) / E);
    __syncwarp();
    return (cnt >= k) ? key_score(kth) : __int_as_float(0xff800000);   // -inf
}

// Large-k variant (capacity C > 512): the same contract, but the sort runs in place in
// memory (warp-wide bitonic network over `cap` slots, cap a power of two, `stride` entries apart)
// instead of in registers.  Slow and rare: it exists so that k up to IVR_MAX_K stays exact.
static __device__ __noinline__ float warp_compact_topk_mem(uint64_t* list, int cnt, int k, int cap, int lane,
                                                           int stride = 1) {
    for (int g = cnt + lane; g < cap; g += 32) list[static_cast<int64_t>(g) * stride] = 0ull;
    __syncwarp();
    for (int s = 2; s <= cap; s <<= 1) {
        for (int t = s >> 1; t >= 1; t >>= 1) {
            for (int i = lane; i < (cap >> 1); i += 32) {
                const int lo = 2 * i - (i & (t - 1));
                const int hi = lo + t;
                const bool desc = (lo & s) == 0;
                const uint64_t a = list[static_cast<int64_t>(lo) * stride], b = list[static_cast<int64_t>(hi) * stride];
                if ((a < b) == desc) { list[static_cast<int64_t>(lo) * stride] = b; list[static_cast<int64_t>(hi) * stride] = a; }
            }
            __syncwarp();
        }
    }
    const uint64_t kth = list[static_cast<int64_t>(k - 1) * stride];
    __syncwarp();
    return (cnt >= k) ? key_score(kth) : __int_as_float(0xff800000);
}

// E > 0: register sort of 32*E slots; E == 0: in-memory sort of `cap` slots.
template <int E>
__device__ __forceinline__ float warp_compact(uint64_t* list, int cnt, int k, int cap, int lane) {
    if constexpr (E > 0) return warp_compact_topk<E>(list, cnt, k, lane);
    else return warp_compact_topk_mem(list, cnt, k, cap, lane);
}

}  // namespace ivr

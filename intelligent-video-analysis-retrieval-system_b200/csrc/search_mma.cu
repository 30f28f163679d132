// search_mma.cu -- K1+K2: batched exact search on the 5th-gen tensor cores.
//
// Replaces D, I = index.search(x, k) for query batches (unified_index.py:503 is
// called once per query by the reference; core.py:891 already passes [nq, d]).
//
//   S[q, r] = <Q[q,:], X[r,:]>      Q: [nq, dpad] bf16 (converted per call)
//                                   X: [ntotal, dpad] bf16 rows, HBM resident
//
// One persistent CTA per SM (CG = 1) or one CTA pair per TPC (CG = 2,
// tcgen05 cta_group::2).  A work unit is a (128*CG queries) x (256 rows) tile:
//   * the query tile stays RESIDENT in shared memory (dpad <= 512: 128 x dpad bf16
//     = 128 KB) -- only the row tiles stream through a TMA/mbarrier ring, which
//     halves the L2->SMEM traffic of a classic GEMM tile loop;
//   * warp 0 (one lane) issues TMA loads (cp.async.bulk.tensor, SWIZZLE_128B);
//   * warp 1 (one lane) issues tcgen05.mma kind::f16 (bf16 x bf16 -> fp32) into one
//     of two 256-column TMEM accumulators;
//   * warps 4-7 drain the other accumulator with tcgen05.ld (32 lanes x 32 columns
//     per instruction): each THREAD owns one query, compares its 256 scores against
//     that query's running admission threshold in registers and appends the few
//     survivors to the query's candidate list; full lists are compacted by a warp
//     bitonic sort.  The score matrix never reaches shared or global memory.
// The per-(CTA, query-tile) survivors are folded by topk_merge.cu.
//
// Algorithmic work: 2 * ntotal * dpad * nq FLOP per search.
#include "index.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdlib>

namespace ivr {

constexpr int kMmaThreads   = 256;
constexpr int kTileN        = 256;                 // DB rows per tile (UMMA N)
constexpr int kTileQ        = 128;                 // queries per CTA (UMMA M per CTA)
constexpr int kKBlock       = 64;                  // bf16 per 128-byte swizzle row
constexpr int kQBlockBytes  = kTileQ * 128;        // one k-block of the query tile: 16 KB
constexpr int kMaxKBlocks   = 8;                   // dpad <= 512 keeps the query tile resident
constexpr int kSmemBudget   = 227 * 1024;
constexpr int kBarrierBytes = 1024;

// ---------------------------------------------------------------- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    if constexpr (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
    } else {
        // both CTAs of the pair signal the LEADER's barrier (peer bit cleared)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(x), "r"(y) : "memory");
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// arrive on `bar` (in every CTA of the group) once all previously issued MMAs have completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B), [46,48) version = 1 (Blackwell),
//   [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1),
// A = B = BF16 (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------- the kernel ----
struct MmaParams {
    int64_t n_rows;          // rows in this shard
    int     nq;              // real queries
    int     k;
    int     C;               // candidate list capacity (2 * kcap)
    int     kblocks;         // dpad / 64
    int     stages;          // row-tile ring depth
    int     tq;              // query tiles
    int64_t nt;              // row tiles
    int     groups;          // CTA groups in the grid
    int     nq_pad;          // tq * 128 * CG
    uint64_t* lists;         // [grid CTAs][128][C] working candidate lists
    uint64_t* out_keys;      // [slots][nq_pad][k]
    int*      out_counts;    // [slots][nq_pad]
};

template <int CG, int E>
__global__ void __launch_bounds__(kMmaThreads, 1)
search_mma_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                  const MmaParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16B aligned: round up to the 1024 B the 128B swizzle needs
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int group = blockIdx.x / CG;

    constexpr int kRowsPerCta = kTileN / CG;                    // rows of each tile this CTA loads
    const uint32_t stage_bytes = kRowsPerCta * 128;
    const uint32_t q_bytes = p.kblocks * kQBlockBytes;
    const uint32_t smem_q = smem_base;
    const uint32_t smem_b = smem_q + q_bytes;
    const uint32_t bars = smem_b + p.stages * stage_bytes;      // barrier block
    auto full_bar  = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (16 + s); };
    const uint32_t q_full = bars + 8u * 32, q_empty = bars + 8u * 33;
    auto tfull_bar  = [&](int a) { return bars + 8u * (34 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (36 + a); };
    const uint32_t tmem_slot = bars + 8u * 40;

    // work units [u0, u1): unit u = (query tile u / nt, row tile u % nt)
    const int64_t U = static_cast<int64_t>(p.tq) * p.nt;
    const int64_t u0 = U * group / p.groups, u1 = U * (group + 1) / p.groups;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * CG); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_x); }
    if (warp == 2) tmem_alloc<CG>(tmem_slot, 512);
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0 && lane == 0) {
        // ============================ TMA producer ============================
        int stage = 0; uint32_t phase = 0; int seg = 0;
        for (int64_t u = u0; u < u1; ++u) {
            const int t = static_cast<int>(u / p.nt);
            const int64_t j = u % p.nt;
            if (u == u0 || j == 0) {                              // new query tile
                if (seg > 0) mbar_wait(q_empty, (seg - 1) & 1);   // MMA finished with the old tile
                if (cta_rank == 0) mbar_expect_tx(q_full, q_bytes * CG);
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_2d<CG>(smem_q + kb * kQBlockBytes, &tmap_q, q_full, kb * kKBlock,
                                    t * (kTileQ * CG) + static_cast<int>(cta_rank) * kTileQ);
                ++seg;
            }
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (cta_rank == 0) mbar_expect_tx(full_bar(stage), stage_bytes * CG);
                tma_load_2d<CG>(smem_b + stage * stage_bytes, &tmap_x, full_bar(stage), kb * kKBlock,
                                static_cast<int>(j * kTileN) + static_cast<int>(cta_rank) * kRowsPerCta);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && cta_rank == 0) {
        // ============================ MMA issuer ==============================
        constexpr uint32_t idesc = make_idesc(kTileQ * CG, kTileN);
        int stage = 0; uint32_t phase = 0; int seg = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int64_t u = u0; u < u1; ++u) {
            const int64_t j = u % p.nt;
            if (u == u0 || j == 0) { mbar_wait(q_full, seg & 1); ++seg; }
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);            // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kTileN);
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a0 = smem_q + kb * kQBlockBytes;
                const uint32_t b0 = smem_b + stage * stage_bytes;
#pragma unroll
                for (int k4 = 0; k4 < kKBlock / 16; ++k4)          // UMMA K = 16 bf16 = 32 bytes
                    umma_f16<CG>(tmem_d, make_smem_desc(a0 + k4 * 32), make_smem_desc(b0 + k4 * 32), idesc,
                                 (kb | k4) ? 1u : 0u);
                umma_commit<CG>(empty_bar(stage));                 // ring slot free when these MMAs retire
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit<CG>(tfull_bar(acc));                       // accumulator ready for the epilogue
            if (j == p.nt - 1 || u == u1 - 1) umma_commit<CG>(q_empty);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (seg > 0) mbar_wait(q_empty, (seg - 1) & 1);            // last async arrive has landed
    } else if (warp == 1 && lane == 0 && cta_rank != 0) {
        // non-leader CTA of a pair: its copy of q_empty receives the multicast arrive of every segment;
        // wait for the last one so the CTA cannot retire under an in-flight arrive
        if (u1 > u0) {
            const int nseg = static_cast<int>((u1 - 1) / p.nt - u0 / p.nt) + 1;
            mbar_wait(q_empty, (nseg - 1) & 1);
        }
    } else if (warp >= 4) {
        // ============================ epilogue: fused top-k ====================
        const int quarter = warp & 3;                              // TMEM lanes [32*quarter, +32)
        const int r = quarter * 32 + lane;                         // query row inside this CTA's tile
        uint64_t* warp_lists = p.lists + (static_cast<int64_t>(blockIdx.x) * kTileQ + quarter * 32) * p.C;
        uint64_t* my_list = warp_lists + static_cast<int64_t>(lane) * p.C;
        const float NEG_INF = __int_as_float(0xff800000), POS_INF = __int_as_float(0x7f800000);
        int cnt = 0; float tau = NEG_INF; int acc = 0; uint32_t acc_phase = 0;
        int cur_t = -1; int64_t q_global = 0;

        auto flush_segment = [&](int t) {
            // compact every query's list to its sorted top-k and publish it in the slot of (group, t)
            const int64_t U_ = static_cast<int64_t>(p.tq) * p.nt;
            const int64_t first_unit = static_cast<int64_t>(t) * p.nt;
            const int gmin = static_cast<int>(((first_unit + 1) * p.groups - 1) / U_);
            const int slot = group - gmin;
            for (int l = 0; l < 32; ++l) {
                const int c_l = __shfl_sync(0xffffffffu, cnt, l);
                const int64_t q_l = __shfl_sync(0xffffffffu, q_global, l);
                if (c_l == 0 || q_l >= p.nq) continue;             // warp-uniform
                uint64_t* lst = warp_lists + static_cast<int64_t>(l) * p.C;
                __syncwarp();
                warp_compact<E>(lst, c_l, p.k, p.C, lane);
                const int keep = min(c_l, p.k);
                uint64_t* dst = p.out_keys + (static_cast<int64_t>(slot) * p.nq_pad + q_l) * p.k;
                for (int i = lane; i < keep; i += 32) dst[i] = lst[i];
                if (lane == 0) p.out_counts[static_cast<int64_t>(slot) * p.nq_pad + q_l] = keep;
            }
        };

        for (int64_t u = u0; u < u1; ++u) {
            const int t = static_cast<int>(u / p.nt);
            const int64_t j = u % p.nt;
            if (t != cur_t) {
                if (cur_t >= 0) flush_segment(cur_t);
                cur_t = t;
                q_global = static_cast<int64_t>(t) * (kTileQ * CG) + cta_rank * kTileQ + r;
                cnt = 0;
                tau = (q_global < p.nq) ? NEG_INF : POS_INF;        // padded queries admit nothing
            }
            const int64_t row0 = j * kTileN;
            const int nvalid = static_cast<int>(min(static_cast<int64_t>(kTileN), p.n_rows - row0));
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * kTileN);
#pragma unroll 1
            for (int c = 0; c < kTileN / 32; ++c) {
                // make room: a chunk can add up to 32 entries to one list
                unsigned need = __ballot_sync(0xffffffffu, cnt > p.C - 32);
                while (need) {
                    const int l = __ffs(need) - 1;
                    need &= need - 1;
                    const int c_l = __shfl_sync(0xffffffffu, cnt, l);
                    __syncwarp();
                    const float t_l = warp_compact<E>(warp_lists + static_cast<int64_t>(l) * p.C, c_l, p.k, p.C, lane);
                    if (lane == l) { tau = t_l; cnt = min(cnt, p.k); }
                }
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                tmem_wait_ld();
                float m = __uint_as_float(v[0]);
#pragma unroll
                for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                if (__any_sync(0xffffffffu, m > tau)) {
                    const int col0 = c * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float s = __uint_as_float(v[i]);
                        if (s > tau && col0 + i < nvalid) {
                            my_list[cnt] = make_key(s, static_cast<uint32_t>(row0 + col0 + i));
                            ++cnt;
                        }
                    }
                }
            }
            // all tcgen05.ld of this accumulator have completed (wait::ld above): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive_local(tempty_bar(acc));
                else mbar_arrive_cluster(tempty_bar(acc), 0);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (cur_t >= 0) flush_segment(cur_t);
    }

    // ------------------------------------------------------------------ teardown ----
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

// fp32 [nq, dim] -> bf16 [nq_pad, dpad], zero padded
__global__ void queries_to_bf16_kernel(const float* __restrict__ q, __nv_bfloat16* __restrict__ out,
                                       int64_t nq, int64_t nq_pad, int dim, int dpad) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= nq_pad * dpad) return;
    const int64_t r = i / dpad;
    const int c = static_cast<int>(i % dpad);
    out[i] = __float2bfloat16_rn((r < nq && c < dim) ? q[r * dim + c] : 0.f);
}

// ------------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle
static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int cols, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return IVR_ECUDA; }
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kKBlock), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r)); return IVR_ECUDA; }
    return IVR_OK;
}

static int cta_group_mode() {
    // IVR_MMA_CTA_GROUP=2 selects the cta_group::2 (CTA pair) variant; read per call so tests can flip it
    const char* e = getenv("IVR_MMA_CTA_GROUP");
    return (e && atoi(e) == 2) ? 2 : 1;
}

bool mma_supported(const ivr_index* idx, int64_t nq, int k) {
    (void)nq;
    return idx->dpad <= kMaxKBlocks * kKBlock && k <= IVR_MAX_K;
}

template <int CG, int E>
static int launch_mma(const CUtensorMap& tq, const CUtensorMap& tx, const MmaParams& p, int grid, size_t smem,
                      cudaStream_t st) {
    auto kern = search_mma_kernel<CG, E>;
    IVR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    IVR_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tx, p));
    return IVR_OK;
}

int search_mma(ivr_index* idx, const float* q_dev, int64_t nq, int k, float* D_dev, int64_t* I_dev,
               int64_t id_offset, cudaStream_t st) {
    const int cg = cta_group_mode();
    const int kcap = kcap_for(k);
    const int C = 2 * kcap;
    const int mq = kTileQ * cg;
    MmaParams p{};
    p.n_rows = idx->ntotal; p.nq = static_cast<int>(nq); p.k = k; p.C = C;
    p.kblocks = idx->dpad / kKBlock;
    p.tq = static_cast<int>((nq + mq - 1) / mq);
    p.nt = (idx->ntotal + kTileN - 1) / kTileN;
    p.nq_pad = p.tq * mq;
    int grid = idx->sm_count / cg * cg;
    p.groups = grid / cg;
    const int64_t U = static_cast<int64_t>(p.tq) * p.nt;
    if (U < p.groups) { p.groups = static_cast<int>(U); grid = p.groups * cg; }
    // smem: resident query tile + as many ring stages as fit
    const int stage_bytes = (kTileN / cg) * 128;
    const int q_bytes = p.kblocks * kQBlockBytes;
    p.stages = std::min(12, (kSmemBudget - 1024 - kBarrierBytes - q_bytes) / stage_bytes);
    if (p.stages < 2) { set_error("search_mma: dim %d leaves no room for the row-tile ring", idx->dim); return IVR_EUNSUPPORTED; }
    const size_t smem = 1024 + q_bytes + static_cast<size_t>(p.stages) * stage_bytes + kBarrierBytes;
    // partial-result slots: the groups whose unit range touches one query tile
    auto g_of = [&](int64_t u) { return static_cast<int>(((u + 1) * p.groups - 1) / U); };
    int slots = 1;
    for (int t = 0; t < p.tq; ++t)
        slots = std::max(slots, g_of(static_cast<int64_t>(t) * p.nt + p.nt - 1) - g_of(static_cast<int64_t>(t) * p.nt) + 1);

    // workspace carve-up
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_q = carve(static_cast<size_t>(p.nq_pad) * idx->dpad * 2);
    const size_t o_l = carve(static_cast<size_t>(grid) * kTileQ * C * 8);
    const size_t o_k = carve(static_cast<size_t>(slots) * p.nq_pad * k * 8);
    const size_t o_c = carve(static_cast<size_t>(slots) * p.nq_pad * 4);
    const size_t tmp_keys = merge_tmp_entries(slots, nq, k);
    const size_t o_t = carve(tmp_keys * 8);
    const size_t o_tc = carve((static_cast<size_t>(slots) * nq + 64) * 4);
    IVR_TRY(ensure_ws(idx, off));
    char* ws = static_cast<char*>(idx->ws);
    __nv_bfloat16* q_bf = reinterpret_cast<__nv_bfloat16*>(ws + o_q);
    p.lists = reinterpret_cast<uint64_t*>(ws + o_l);
    p.out_keys = reinterpret_cast<uint64_t*>(ws + o_k);
    p.out_counts = reinterpret_cast<int*>(ws + o_c);

    if (idx->timing) cudaEventRecord(idx->ev[4], st);
    {
        const int64_t n = static_cast<int64_t>(p.nq_pad) * idx->dpad;
        queries_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(q_dev, q_bf, nq, p.nq_pad,
                                                                                      idx->dim, idx->dpad);
        IVR_CUDA(cudaGetLastError());
        idx->launches[2]++;
        IVR_CUDA(cudaMemsetAsync(p.out_counts, 0, static_cast<size_t>(slots) * p.nq_pad * 4, st));
    }
    if (idx->timing) cudaEventRecord(idx->ev[5], st);

    // TMA descriptors (the row descriptor is cached until the matrix moves or grows)
    CUtensorMap tmq;
    IVR_TRY(make_tmap(&tmq, q_bf, p.nq_pad, idx->dpad, kTileQ));
    if (idx->tmap_rows_base != idx->rows || idx->tmap_rows_n != idx->ntotal || idx->tmap_rows_box != kTileN / cg) {
        IVR_TRY(make_tmap(reinterpret_cast<CUtensorMap*>(idx->tmap_rows), idx->rows, idx->ntotal, idx->dpad, kTileN / cg));
        idx->tmap_rows_base = idx->rows; idx->tmap_rows_n = idx->ntotal; idx->tmap_rows_box = kTileN / cg;
    }
    const CUtensorMap& tmx = *reinterpret_cast<const CUtensorMap*>(idx->tmap_rows);

    if (idx->timing) cudaEventRecord(idx->ev[0], st);
    int rc;
    if (cg == 2) rc = (kcap == 128) ? launch_mma<2, 8>(tmq, tmx, p, grid, smem, st) : launch_mma<2, 0>(tmq, tmx, p, grid, smem, st);
    else         rc = (kcap == 128) ? launch_mma<1, 8>(tmq, tmx, p, grid, smem, st) : launch_mma<1, 0>(tmq, tmx, p, grid, smem, st);
    IVR_TRY(rc);
    idx->launches[0]++;
    if (idx->timing) { cudaEventRecord(idx->ev[1], st); cudaEventRecord(idx->ev[2], st); }

    MergeIn in{};
    in.entries = p.out_keys; in.counts = p.out_counts;
    in.list_stride = static_cast<int64_t>(p.nq_pad) * k; in.q_stride = k;
    in.cnt_list_stride = p.nq_pad; in.cnt_q_stride = 1;
    in.n_lists = slots; in.fixed_count = 0;
    IVR_TRY(merge_lists_final(in, nq, k, D_dev, I_dev, id_offset, reinterpret_cast<uint64_t*>(ws + o_t),
                              reinterpret_cast<int*>(ws + o_tc), st, &idx->launches[1]));
    if (idx->timing) {
        cudaEventRecord(idx->ev[3], st);
        idx->ev_valid[0] = idx->ev_valid[1] = idx->ev_valid[2] = true;
    }
    return IVR_OK;
}

}  // namespace ivr

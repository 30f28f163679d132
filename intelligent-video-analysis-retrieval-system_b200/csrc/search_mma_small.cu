// search_mma_small.cu -- K1s: the tcgen05 search for SMALL query batches (2 .. 128 queries).
//
// With a handful of queries the search is bound by streaming the rows from HBM once (SURVEY.md 8d:
// N * D * 2 bytes per batch).  The batched kernels of search_mma.cu put the QUERIES on the UMMA M
// axis, so 16 queries still cost a full M = 256 instruction per 256 rows -- profiling showed the
// tensor pipe 82 % busy (and the SM clock power-capped) multiplying padding, and 70-75 % of the HBM
// rate.  Here the operands are swapped:
//
//     D[rows (M = 128 TMEM lanes), queries (N = npad columns)] = X_tile[128, dpad] * Q[npad, dpad]^T
//
// so the tensor work is proportional to the number of queries (N = 16 .. 128) and the kernel is a
// pure row streamer:
//   * the query matrix (npad x dpad fp16, <= 128 KB) is loaded ONCE per CTA and stays in shared memory;
//   * warp 0 (one lane) streams 128-row tiles through a TMA/mbarrier ring of 16 KB stages
//     (one 64-column k-block of 128 rows each, SWIZZLE_128B), every CTA a contiguous range of tiles;
//   * warp 1 (one lane) issues tcgen05.mma cta_group::1 kind::f16 (M = 128, N = npad, K = 16) into one
//     of up to 8 TMEM accumulators of npad columns;
//   * warps 2-5 drain the accumulators: tcgen05.ld hands every THREAD one ROW's scores for 32 queries;
//     each is compared with its query's admission threshold (shared memory, refreshed from the
//     per-query threshold all CTAs share through global memory); survivors are appended with a warp
//     ballot to the (CTA, warp, query) candidate list; a list that could overflow on the next tile is
//     compacted to its k best by a warp bitonic sort (registers for k <= 128, in memory above), which
//     also raises the thresholds.
// Large shards are searched in up to three launches of growing size (one tile per CTA, then ~380x
// more, then the rest): the exact k-th best of the rows seen so far seeds the thresholds of the
// next launch, so that lists almost never need compaction (the cold start of 4 x 148 lists per
// query would otherwise cost more than the streaming itself).
// The per-(CTA, warp) survivors are folded by topk_merge.cu.
//
// Replaces D, I = index.search(x, k) for small batches (unified_index.py:503, core.py:891).
// Algorithmic bytes: ntotal * dpad * 2 per search.
#include "mma_common.cuh"

namespace ivr {

constexpr int kSmallThreads    = 192;                  // producer, MMA issuer, 4 epilogue warps
constexpr int kSmallTileRows   = 128;                  // rows per tile (UMMA M)
constexpr int kSmallStageBytes = kSmallTileRows * 128; // one k-block of a row tile: 16 KB
constexpr size_t kSmallMaxListBytes = 2ull << 30;      // candidate-list workspace bound (large k x many queries)
constexpr int kSmallMaxQ       = 128;                  // queries per launch (UMMA N)
constexpr int kSmallMaxBuf     = 8;                    // TMEM accumulators
constexpr int kSmallAuxBytes   = 4096;                 // barriers + thresholds + per-warp counts

struct SmallParams {
    int64_t n_rows;          // rows in this shard
    int     nq, npad;        // real queries; padded to a multiple of 16 (UMMA N)
    int     k;
    int     C;               // candidate list capacity: 2 * kcap(k) (256 for k <= 128: register sort)
    int     kblocks;         // dpad / 64
    int     stages;          // row ring depth (16 KB each)
    int     nbuf, buf_cols;  // TMEM accumulators and their column stride
    int64_t nt;              // row tiles of this launch
    int64_t tile0;           // first row tile of this launch (row id = (tile0 + j) * 128 + lane)
    uint64_t row_policy;     // L2 eviction priority of the row loads
    uint64_t* lists;         // [grid][4 warps][npad][C] raw candidate lists
    int*      counts;        // [grid][4 warps][npad]
    uint32_t* tau_g;         // [npad] shared per-query thresholds (order-preserving encoding)
    // DUMP mode (small shards: dump != nullptr): no candidate lists -- every score goes to dump[q * dump_stride + row]
    // (4 bytes per (query, row): a few % of the row bytes for the batches this kernel serves) and every 128-row tile
    // publishes its maximum per query, tile_max[q * n_tiles + tile]; merge_select_kernel then needs only the ~k tiles
    // whose maximum reaches the k-th largest tile maximum.  One launch, no threshold learning, no seeding launches.
    float*    dump;
    int64_t   dump_stride;   // floats per query (>= padded row count)
    uint32_t* tile_max;      // zero-initialised by the caller
    int64_t   n_tiles;       // tiles of the whole shard
    uint32_t* group_max;     // [npad][n_groups]: maxima of n_groups runs of consecutive tiles (zero-initialised)
    int       n_groups;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kSmallThreads, 1)
search_mma_small_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                        const SmallParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t q_block = static_cast<uint32_t>(p.npad) * 128u;   // one k-block of the query matrix
    const uint32_t q_bytes = p.kblocks * q_block;
    const uint32_t smem_q = smem_base;
    const uint32_t smem_a = smem_q + q_bytes;
    const uint32_t aux = smem_a + p.stages * kSmallStageBytes;
    auto full_bar   = [&](int s) { return aux + 8u * s; };
    auto empty_bar  = [&](int s) { return aux + 8u * (16 + s); };
    const uint32_t q_full = aux + 8u * 32;
    auto tfull_bar  = [&](int a) { return aux + 8u * (34 + a); };
    auto tempty_bar = [&](int a) { return aux + 8u * (42 + a); };
    const uint32_t tmem_slot = aux + 8u * 50;
    float* const tau_s = reinterpret_cast<float*>(gen_base + (aux - smem_base) + 512);   // [128]
    int*   const cnt_s = reinterpret_cast<int*>(gen_base + (aux - smem_base) + 1024);    // [4][128]

    const int64_t j0 = p.nt * blockIdx.x / gridDim.x, j1 = p.nt * (blockIdx.x + 1) / gridDim.x;   // owned row tiles

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(q_full, 1);
        for (int a = 0; a < p.nbuf; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_x); }
    if (warp == 1) tmem_alloc<1>(tmem_slot, 512);
    for (int i = threadIdx.x; i < 4 * kSmallMaxQ; i += kSmallThreads) cnt_s[i] = 0;
    for (int i = threadIdx.x; i < kSmallMaxQ; i += kSmallThreads)      // padded queries admit nothing
        tau_s[i] = (i < p.nq) ? ord2f(__ldcg(p.tau_g + i)) : __int_as_float(0x7f800000);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0 && lane == 0) {
        // ============================ TMA producer ============================
        mbar_expect_tx(q_full, q_bytes);
        for (int kb = 0; kb < p.kblocks; ++kb)
            tma_load_2d<1>(smem_q + kb * q_block, &tmap_q, q_full, kb * kKBlock, 0, kL2EvictLast);
        int stage = 0; uint32_t phase = 0;
        for (int64_t j = j0; j < j1; ++j) {
            const int row0 = static_cast<int>((p.tile0 + j) * kSmallTileRows);
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), kSmallStageBytes);
                tma_load_2d<1>(smem_a + stage * kSmallStageBytes, &tmap_x, full_bar(stage), kb * kKBlock, row0, p.row_policy);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ============================ MMA issuer ==============================
        const uint32_t idesc = make_idesc(kSmallTileRows, p.npad);
        mbar_wait(q_full, 0);
        tc_fence_after();
        int stage = 0; uint32_t phase = 0; int buf = 0; uint32_t bphase = 0;
        for (int64_t j = j0; j < j1; ++j) {
            mbar_wait(tempty_bar(buf), bphase ^ 1);                  // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * p.buf_cols);
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a0 = smem_a + stage * kSmallStageBytes;   // A: streamed rows (M)
                const uint32_t b0 = smem_q + kb * q_block;               // B: resident queries (N)
#pragma unroll
                for (int k4 = 0; k4 < kKBlock / 16; ++k4)
                    umma_f16<1>(tmem_d, make_smem_desc(a0 + k4 * 32), make_smem_desc(b0 + k4 * 32), idesc,
                                (kb | k4) ? 1u : 0u);
                umma_commit<1>(empty_bar(stage));                    // ring slot free when these MMAs retire
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit<1>(tfull_bar(buf));
            if (++buf == p.nbuf) { buf = 0; bphase ^= 1; }
        }
    } else if (warp >= 2) {
        // ============================ epilogue: fused top-k ====================
        const int quarter = warp & 3;                              // the TMEM lanes this warp may read
        int* const cnt_w = cnt_s + quarter * kSmallMaxQ;
        const int64_t wslot = static_cast<int64_t>(blockIdx.x) * 4 + quarter;
        uint64_t* const wl = p.lists + wslot * p.npad * p.C;
        const unsigned lt_mask = (1u << lane) - 1u;

        auto compact = [&](int q) {                                // all lanes; list of query q
            const int c = cnt_w[q];
            __syncwarp();
            uint64_t* const lst = wl + static_cast<int64_t>(q) * p.C;
            const float t = (p.C == 256) ? warp_compact_raw<8>(lst, c, p.k, p.C, lane)      // register bitonic sort
                                         : warp_compact_raw<0>(lst, c, p.k, p.C, lane);     // in-memory (k > 128)
            if (lane == 0) {
                cnt_w[q] = min(c, p.k);
                if (c >= p.k) {
                    if (t > tau_s[q]) tau_s[q] = t;                // benign race: any value is a valid lower bound
                    atomicMax(p.tau_g + q, f2ord(t));              // share with the other CTAs
                }
            }
            __syncwarp();
        };

        int buf = 0; uint32_t bphase = 0; int64_t n = 0;
        if (p.dump) {
            // ---------------- dump mode: scores + per-tile maxima, no lists ----------------
            for (int64_t j = j0; j < j1; ++j) {
                const int64_t tile = p.tile0 + j;
                const int64_t grp = tile * p.n_groups / p.n_tiles;
                const int64_t row = tile * kSmallTileRows + quarter * 32 + lane;
                const bool valid = row < p.n_rows;
                mbar_wait(tfull_bar(buf), bphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                       static_cast<uint32_t>(buf * p.buf_cols);
                for (int c0 = 0; c0 < p.npad; c0 += 32) {
                    uint32_t v[32];
                    const int ncol = min(32, p.npad - c0);
                    if (ncol == 32) tmem_ld_32x32(taddr + c0, v);
                    else tmem_ld_32x16(taddr + c0, v);                  // npad is a multiple of 16
                    tmem_wait_ld(v);
                    if (c0 + 32 >= p.npad) {                            // accumulator fully read: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(tempty_bar(buf));
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= ncol) break;                           // warp-uniform
                        const uint32_t bits = valid ? v[i] : 0xff800000u;   // rows past the end: -inf
                        float* dst = p.dump + static_cast<int64_t>(c0 + i) * p.dump_stride + row;
                        *dst = __uint_as_float(bits);                    // 32 consecutive rows: one 128-byte store per query
                        const uint32_t mx = __reduce_max_sync(0xffffffffu, f2ord(__uint_as_float(bits)));
                        if (lane == 0) {
                            atomicMax(p.tile_max + static_cast<int64_t>(c0 + i) * p.n_tiles + tile, mx);
                            atomicMax(p.group_max + static_cast<int64_t>(c0 + i) * p.n_groups + grp, mx);
                        }
                    }
                }
                if (++buf == p.nbuf) { buf = 0; bphase ^= 1; }
            }
        } else
        for (int64_t j = j0; j < j1; ++j, ++n) {
            // make room: one tile adds at most 32 entries (one per lane) to each of this warp's lists
            for (int qb = 0; qb < p.npad; qb += 32) {
                const int q = qb + lane;
                unsigned need = __ballot_sync(0xffffffffu, q < p.npad && cnt_w[q] > p.C - 32);
                while (need) { const int l = __ffs(need) - 1; need &= need - 1; compact(qb + l); }
            }
            // fold in the thresholds the other CTAs have published (one warp per tile, round robin)
            if (static_cast<int>(n & 3) == quarter)
                for (int q = lane; q < p.nq; q += 32) {
                    const float t = ord2f(__ldcg(p.tau_g + q));
                    if (t > tau_s[q]) tau_s[q] = t;
                }
            __syncwarp();
            const uint32_t row = static_cast<uint32_t>((p.tile0 + j) * kSmallTileRows + quarter * 32 + lane);
            const bool valid = static_cast<int64_t>(row) < p.n_rows;
            mbar_wait(tfull_bar(buf), bphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(buf * p.buf_cols);
            for (int c0 = 0; c0 < p.npad; c0 += 32) {
                uint32_t v[32];
                if (p.npad - c0 >= 32) {
                    tmem_ld_32x32(taddr + c0, v);
                } else {                                            // npad is a multiple of 16: a 16-column tail
#pragma unroll
                    for (int i = 16; i < 32; ++i) v[i] = 0xff800000u;
                    tmem_ld_32x16(taddr + c0, v);
                }
                tmem_wait_ld(v);
                if (c0 + 32 >= p.npad) {                            // accumulator fully read: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(tempty_bar(buf));
                }
                unsigned m = 0;
                const float4* t4 = reinterpret_cast<const float4*>(tau_s + c0);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 t = t4[i4];
                    m |= (__uint_as_float(v[4 * i4 + 0]) > t.x ? 1u : 0u) << (4 * i4 + 0);
                    m |= (__uint_as_float(v[4 * i4 + 1]) > t.y ? 1u : 0u) << (4 * i4 + 1);
                    m |= (__uint_as_float(v[4 * i4 + 2]) > t.z ? 1u : 0u) << (4 * i4 + 2);
                    m |= (__uint_as_float(v[4 * i4 + 3]) > t.w ? 1u : 0u) << (4 * i4 + 3);
                }
                if (!valid) m = 0;                                  // rows past the end of the shard (TMA zero fill)
                const unsigned wm = __reduce_or_sync(0xffffffffu, m);
                if (wm == 0) continue;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!(wm & (1u << i))) continue;                // warp-uniform
                    const bool hit = (m >> i) & 1u;
                    const unsigned hits = __ballot_sync(0xffffffffu, hit);
                    const int q = c0 + i;
                    const int base = cnt_w[q];
                    if (hit)
                        wl[static_cast<int64_t>(q) * p.C + base + __popc(hits & lt_mask)] =
                            static_cast<uint64_t>(v[i]) | (static_cast<uint64_t>(row) << 32);
                    __syncwarp();
                    if (lane == 0) cnt_w[q] = base + __popc(hits);
                }
                __syncwarp();
            }
            if (++buf == p.nbuf) { buf = 0; bphase ^= 1; }
        }
        __syncwarp();
        if (!p.dump)
            for (int q = lane; q < p.npad; q += 32) p.counts[wslot * p.npad + q] = cnt_w[q];
    }

    // ------------------------------------------------------------------ teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

// ------------------------------------------------------------------ host side ----
struct SmallShape { int npad, stages, nbuf, buf_cols, C; size_t smem; };

static bool small_shape(const ivr_index* idx, int64_t nq, int k, SmallShape* s) {
    if (k < 1 || k > IVR_MAX_K || nq < 1 || nq > kSmallMaxQ) return false;
    const int npad = static_cast<int>((nq + 15) / 16 * 16);
    const int C = 2 * kcap_for(k);
    if (static_cast<size_t>(4 * idx->sm_count) * npad * C * 8 > kSmallMaxListBytes) return false;   // huge k x many queries
    const int64_t q_bytes = static_cast<int64_t>(npad) * idx->dpad * 2;
    if (q_bytes > 128 * 1024) return false;
    const int stages = std::min<int64_t>(12, (kSmemBudget - 1024 - kSmallAuxBytes - q_bytes) / kSmallStageBytes);
    if (stages < 4) return false;
    s->npad = npad; s->stages = stages; s->C = C;
    s->buf_cols = std::max(npad, 32);
    s->nbuf = std::min(kSmallMaxBuf, 512 / s->buf_cols);
    s->smem = 1024 + static_cast<size_t>(q_bytes) + static_cast<size_t>(stages) * kSmallStageBytes + kSmallAuxBytes;
    return true;
}

bool mma_small_supported(const ivr_index* idx, int64_t nq, int k) {
    SmallShape s;
    return small_shape(idx, nq, k, &s);
}

// One launch in dump mode + the one-launch select (merge_select_kernel, dense input).
static int search_mma_small_dump(ivr_index* idx, const SmallShape& sh, const float* q_dev, int64_t nq, int k, float* D_dev,
                                 int64_t* I_dev, int64_t id_offset, cudaStream_t st) {
    const int64_t nt = (idx->ntotal + kSmallTileRows - 1) / kSmallTileRows;
    const int64_t stride = nt * kSmallTileRows;                      // padded rows per query
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_q  = carve(static_cast<size_t>(sh.npad) * idx->dpad * 2);
    const size_t o_sc = carve(static_cast<size_t>(sh.npad) * 4);
    const size_t o_tg = carve(static_cast<size_t>(sh.npad) * 4);
    const size_t o_s  = carve(static_cast<size_t>(sh.npad) * stride * 4);
    const int n_groups = static_cast<int>(std::min<int64_t>(nt, 1024));
    const size_t zero_bytes = (static_cast<size_t>(sh.npad) * (nt + n_groups) + 2 * nq) * 4;   // tile + group maxima, pool counters, tickets
    const size_t o_z  = carve(zero_bytes);
    const size_t o_p  = carve(static_cast<size_t>(nq) * kSelectPoolCap * 8);
    IVR_TRY(ensure_ws(idx, off));
    char* ws = static_cast<char*>(idx->ws);
    __half* q_h = reinterpret_cast<__half*>(ws + o_q);
    float* q_scale = reinterpret_cast<float*>(ws + o_sc);
    uint32_t* tau_g = reinterpret_cast<uint32_t*>(ws + o_tg);
    uint32_t* tile_max = reinterpret_cast<uint32_t*>(ws + o_z);
    uint32_t* group_max = tile_max + static_cast<size_t>(sh.npad) * nt;
    int* pool_cnt = reinterpret_cast<int*>(group_max + static_cast<size_t>(sh.npad) * n_groups);
    const bool timed = idx->timing;

    if (timed) cudaEventRecord(idx->ev[4], st);
    IVR_TRY(launch_queries_to_f16(q_dev, q_h, q_scale, tau_g, nq, sh.npad, idx->dim, idx->dpad, st));
    IVR_CUDA(cudaMemsetAsync(ws + o_z, 0, zero_bytes, st));
    idx->launches[2]++;
    if (timed) cudaEventRecord(idx->ev[5], st);

    CUtensorMap tmq;
    IVR_TRY(make_tmap(&tmq, q_h, sh.npad, idx->dpad, sh.npad));
    if (idx->tmap_rows_base != idx->rows || idx->tmap_rows_n != idx->ntotal || idx->tmap_rows_box != kSmallTileRows) {
        IVR_TRY(make_tmap(reinterpret_cast<CUtensorMap*>(idx->tmap_rows), idx->rows, idx->ntotal, idx->dpad, kSmallTileRows));
        idx->tmap_rows_base = idx->rows; idx->tmap_rows_n = idx->ntotal; idx->tmap_rows_box = kSmallTileRows;
    }
    const CUtensorMap& tmx = *reinterpret_cast<const CUtensorMap*>(idx->tmap_rows);
    IVR_CUDA(cudaFuncSetAttribute(search_mma_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sh.smem)));
    SmallParams p{};
    p.n_rows = idx->ntotal; p.nq = static_cast<int>(nq); p.npad = sh.npad; p.k = k; p.C = sh.C;
    p.kblocks = idx->dpad / kKBlock; p.stages = sh.stages; p.nbuf = sh.nbuf; p.buf_cols = sh.buf_cols;
    p.tile0 = 0; p.nt = nt;
    p.row_policy = kL2EvictFirst;
    p.tau_g = tau_g;
    p.dump = reinterpret_cast<float*>(ws + o_s); p.dump_stride = stride; p.tile_max = tile_max; p.n_tiles = nt;
    p.group_max = group_max; p.n_groups = n_groups;
    const int grid = static_cast<int>(std::min<int64_t>(idx->sm_count, std::max<int64_t>(nt, 1)));
    if (timed) cudaEventRecord(idx->ev[0], st);
    search_mma_small_kernel<<<grid, kSmallThreads, sh.smem, st>>>(tmq, tmx, p);
    IVR_CUDA(cudaGetLastError());
    idx->launches[0]++;
    if (timed) { cudaEventRecord(idx->ev[1], st); cudaEventRecord(idx->ev[2], st); }
    MergeIn in{};
    in.dense = p.dump; in.dense_q_stride = stride; in.dense_rows = idx->ntotal; in.dense_len = kSmallTileRows;
    in.cnt_list_stride = 1; in.cnt_q_stride = nt;                   // strides of the maxima array
    in.n_lists = static_cast<int>(nt);
    IVR_TRY(merge_select_final(in, tile_max, nq, k, D_dev, I_dev, id_offset, reinterpret_cast<uint64_t*>(ws + o_p),
                               pool_cnt, pool_cnt + nq, idx->sm_count, st, &idx->launches[1], q_scale, group_max, n_groups));
    if (timed) {
        cudaEventRecord(idx->ev[3], st);
        idx->ev_valid[0] = idx->ev_valid[1] = idx->ev_valid[2] = true;
    }
    idx->last_kernel = "search_mma_small_kernel";
    return IVR_OK;
}

int search_mma_small(ivr_index* idx, const float* q_dev, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                     int64_t id_offset, cudaStream_t st) {
    SmallShape sh;
    if (!small_shape(idx, nq, k, &sh)) { set_error("search_mma_small: unsupported shape (nq=%lld, k=%d, dim=%d)",
                                                   static_cast<long long>(nq), k, idx->dim); return IVR_EUNSUPPORTED; }
    const int64_t nt = (idx->ntotal + kSmallTileRows - 1) / kSmallTileRows;
    const int sms = idx->sm_count;
    // Small shards: one launch that materialises the scores (see SmallParams::dump).  The seeded candidate-list path
    // pays ~0.2 ms of fixed cost (2-3 launches, their merges, cold lists); writing the scores costs nq * 4 bytes per
    // row.  Measured (dump vs lists, ms): 1 M x 512: 16 q 0.26 vs 0.37, 32 q 0.30 vs 0.37, 64 q 0.44 vs 0.41;
    // 851 k x 768: 16 q 0.30 vs 0.41, 64 q 0.37 vs 0.43; 3 M x 512: 16 q 0.64 vs 0.68, 32 q 0.72 vs 0.68; 4 M: equal;
    // 6 M: lists win.  Rule: up to 3.5 M rows while the scores add at most 1/6 to the row bytes.
    // IVR_SMALL_DUMP_MAX_KROWS (0 = never) overrides the row limit: tests use it to reach both modes.
    const int64_t dump_rows = static_cast<int64_t>(env_int("IVR_SMALL_DUMP_MAX_KROWS", 3500)) * 1000;
    if (k <= kSelectMaxK && idx->ntotal <= dump_rows && sh.npad * 4 * 6 <= idx->dpad * 2)
        return search_mma_small_dump(idx, sh, q_dev, nq, k, D_dev, I_dev, id_offset, st);
    // launch boundaries (in row tiles): one tile per CTA, then as many rows as keep the expected admissions per
    // list around 64 (k * rows_now / rows_before spread over 4 * sms lists), then the rest
    int64_t bounds[4] = {0, 0, 0, 0};
    int n_phases = 1;
    if (env_int("IVR_MMA_TWO_PHASE", 1) && nt >= 4 * static_cast<int64_t>(sms)) {
        const int64_t b0 = sms;
        const int64_t ratio = std::max<int64_t>(2, env_int("IVR_MMA_SMALL_RATIO", 0) > 0 ? env_int("IVR_MMA_SMALL_RATIO", 0)
                                                                                     : static_cast<int64_t>(sh.C / 4) * 4 * sms / k);
        const int64_t b1 = b0 + b0 * ratio;
        bounds[n_phases++] = b0;
        if (nt > b1 + b1 / 2) bounds[n_phases++] = b1;
    }
    bounds[n_phases] = nt;

    const int max_grid = static_cast<int>(std::min<int64_t>(sms, std::max<int64_t>(nt, 1)));
    const int max_lists = std::max(4 * max_grid, 3);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_q  = carve(static_cast<size_t>(sh.npad) * idx->dpad * 2);
    const size_t o_sc = carve(static_cast<size_t>(sh.npad) * 4);
    const size_t o_tg = carve(static_cast<size_t>(sh.npad) * 4);
    const size_t o_l  = carve(static_cast<size_t>(max_lists) * sh.npad * sh.C * 8);
    const size_t o_c  = carve(static_cast<size_t>(max_lists) * sh.npad * 4);
    const size_t o_pk = carve(static_cast<size_t>(3) * nq * k * 8);      // per-launch merged keys [phase][nq][k]
    const size_t o_pc = carve(static_cast<size_t>(3) * nq * 4);
    const size_t o_t  = carve(merge_tmp_entries(max_lists, nq, k) * 8);
    const size_t o_tc = carve((static_cast<size_t>(max_lists) * nq + 64) * 4);
    IVR_TRY(ensure_ws(idx, off));
    char* ws = static_cast<char*>(idx->ws);
    __half* q_h = reinterpret_cast<__half*>(ws + o_q);
    float* q_scale = reinterpret_cast<float*>(ws + o_sc);
    uint32_t* tau_g = reinterpret_cast<uint32_t*>(ws + o_tg);
    uint64_t* ph_keys = reinterpret_cast<uint64_t*>(ws + o_pk);
    int* ph_counts = reinterpret_cast<int*>(ws + o_pc);
    uint64_t* tmp_e = reinterpret_cast<uint64_t*>(ws + o_t);
    int* tmp_c = reinterpret_cast<int*>(ws + o_tc);
    const bool timed = idx->timing;

    if (timed) cudaEventRecord(idx->ev[4], st);
    IVR_TRY(launch_queries_to_f16(q_dev, q_h, q_scale, tau_g, nq, sh.npad, idx->dim, idx->dpad, st));
    idx->launches[2]++;
    if (timed) cudaEventRecord(idx->ev[5], st);

    CUtensorMap tmq;
    IVR_TRY(make_tmap(&tmq, q_h, sh.npad, idx->dpad, sh.npad));
    if (idx->tmap_rows_base != idx->rows || idx->tmap_rows_n != idx->ntotal || idx->tmap_rows_box != kSmallTileRows) {
        IVR_TRY(make_tmap(reinterpret_cast<CUtensorMap*>(idx->tmap_rows), idx->rows, idx->ntotal, idx->dpad, kSmallTileRows));
        idx->tmap_rows_base = idx->rows; idx->tmap_rows_n = idx->ntotal; idx->tmap_rows_box = kSmallTileRows;
    }
    const CUtensorMap& tmx = *reinterpret_cast<const CUtensorMap*>(idx->tmap_rows);
    IVR_CUDA(cudaFuncSetAttribute(search_mma_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(sh.smem)));

    if (timed) cudaEventRecord(idx->ev[0], st);
    for (int ph = 0; ph < n_phases; ++ph) {
        SmallParams p{};
        p.n_rows = idx->ntotal; p.nq = static_cast<int>(nq); p.npad = sh.npad; p.k = k; p.C = sh.C;
        p.kblocks = idx->dpad / kKBlock; p.stages = sh.stages; p.nbuf = sh.nbuf; p.buf_cols = sh.buf_cols;
        p.tile0 = bounds[ph]; p.nt = bounds[ph + 1] - bounds[ph];
        p.row_policy = kL2EvictFirst;                                // every row is read exactly once
        p.lists = reinterpret_cast<uint64_t*>(ws + o_l);
        p.counts = reinterpret_cast<int*>(ws + o_c);
        p.tau_g = tau_g;
        const int grid = static_cast<int>(std::min<int64_t>(sms, std::max<int64_t>(p.nt, 1)));
        search_mma_small_kernel<<<grid, kSmallThreads, sh.smem, st>>>(tmq, tmx, p);
        IVR_CUDA(cudaGetLastError());
        idx->launches[0]++;
        MergeIn in{};
        in.entries = p.lists; in.counts = p.counts;
        in.list_stride = static_cast<int64_t>(sh.npad) * sh.C; in.q_stride = sh.C;
        in.cnt_list_stride = sh.npad; in.cnt_q_stride = 1;
        in.n_lists = 4 * grid; in.fixed_count = 0; in.raw = 1;
        if (n_phases == 1) {
            if (timed) { cudaEventRecord(idx->ev[1], st); cudaEventRecord(idx->ev[2], st); }
            IVR_TRY(merge_lists_final(in, nq, k, D_dev, I_dev, id_offset, tmp_e, tmp_c, st, &idx->launches[1], q_scale));
        } else {
            IVR_TRY(merge_lists_keys(in, nq, k, ph_keys + static_cast<size_t>(ph) * nq * k,
                                     ph_counts + static_cast<size_t>(ph) * nq, tmp_e, tmp_c, st, &idx->launches[1]));
            if (ph + 1 < n_phases) {
                // the k-th best so far: after launch 0 it is launch 0's; after launch 1 tau_g already holds
                // launch 0's, so seeding with launch 1's own k-th best can only raise it
                IVR_TRY(launch_seed_tau(ph_keys + static_cast<size_t>(ph) * nq * k, ph_counts + static_cast<size_t>(ph) * nq,
                                        tau_g, nq, k, st));
                idx->launches[1]++;
            }
        }
    }
    if (n_phases > 1) {
        if (timed) { cudaEventRecord(idx->ev[1], st); cudaEventRecord(idx->ev[2], st); }
        MergeIn in{};
        in.entries = ph_keys; in.counts = ph_counts;
        in.list_stride = nq * k; in.q_stride = k; in.cnt_list_stride = nq; in.cnt_q_stride = 1;
        in.n_lists = n_phases; in.fixed_count = 0; in.raw = 0;
        IVR_TRY(merge_lists_final(in, nq, k, D_dev, I_dev, id_offset, tmp_e, tmp_c, st, &idx->launches[1], q_scale));
    }
    if (timed) {
        cudaEventRecord(idx->ev[3], st);
        idx->ev_valid[0] = idx->ev_valid[1] = idx->ev_valid[2] = true;
    }
    idx->last_kernel = "search_mma_small_kernel";
    return IVR_OK;
}

}  // namespace ivr

// topk_merge.cu -- K5: exact k-way merge of candidate lists (one CTA per query).
//
// Used (a) after the streaming / MMA scoring kernels to fold the per-warp or
// per-CTA partial lists into the final [nq,k] result, and (b) after the NCCL
// all-gather of the row shards' local top-k (ivr_topk_merge_device).
//
// Method: MSB-first radix select on the 64-bit candidate keys (8 bits per pass,
// shared-memory histogram) finds the k-th largest key exactly, the k survivors
// are gathered into shared memory and bitonic-sorted.  Integer-exact; ties on
// score resolve to the lower row id because the id is part of the key.
#include "index.cuh"

namespace ivr {

constexpr int kMergeThreads = 256;
constexpr int kMergeFanIn   = 64;     // lists folded by one CTA in a non-final level

struct MergeOut {
    float*    D;            // final
    int64_t*  I;
    int64_t   id_offset;
    const float* q_scale;   // final: per-query score multiplier (nullptr = 1)
    uint64_t* entries;      // non-final: [groups, nq, k]
    int*      counts;       // non-final: [groups, nq]
    int64_t   nq;
};

template <typename F>
__device__ __forceinline__ void for_each_key(const MergeIn& in, int64_t q, int l0, int l1, F&& f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int l = l0 + warp; l < l1; l += nwarps) {
        const int cnt = in.counts ? in.counts[l * in.cnt_list_stride + q * in.cnt_q_stride]
                                  : in.fixed_count;
        const uint64_t* base = in.entries + l * in.list_stride +
                               (in.interleave ? (q >> 5) * in.q_stride * 32 + (q & 31) : q * in.q_stride);
        const int es = in.interleave ? 32 : 1;
        for (int i = lane; i < cnt; i += 32) {
            uint64_t key = base[static_cast<int64_t>(i) * es];
            if (in.raw) key = make_key(__uint_as_float(static_cast<uint32_t>(key)), static_cast<uint32_t>(key >> 32));
            if (key != 0ull) f(key);
        }
    }
}

template <bool FINAL>
__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(MergeIn in, MergeOut out, int lists_per_group, int k, int kpad) {
    extern __shared__ uint64_t s_keys[];               // kpad
    __shared__ int      s_hist[256];
    __shared__ uint64_t s_prefix, s_mask;
    __shared__ int      s_remaining, s_done, s_n, s_neq, s_total;

    const int64_t q = blockIdx.x;
    const int grp = blockIdx.y;
    const int l0 = grp * lists_per_group;
    const int l1 = min(l0 + lists_per_group, in.n_lists);
    const int tid = threadIdx.x;

    if (tid == 0) { s_total = 0; s_n = 0; s_neq = 0; s_prefix = 0; s_mask = 0; s_remaining = k; s_done = 0; }
    for (int i = tid; i < kpad; i += blockDim.x) s_keys[i] = 0ull;
    __syncthreads();

    {   // total number of real candidates
        int local = 0;
        for_each_key(in, q, l0, l1, [&](uint64_t) { ++local; });
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((tid & 31) == 0 && local) atomicAdd(&s_total, local);
    }
    __syncthreads();
    const int total = s_total;

    if (total <= k) {
        for_each_key(in, q, l0, l1, [&](uint64_t key) { s_keys[atomicAdd(&s_n, 1)] = key; });
    } else {
        uint64_t prefix = 0, mask = 0;
        int remaining = k;
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            for (int i = tid; i < 256; i += blockDim.x) s_hist[i] = 0;
            __syncthreads();
            for_each_key(in, q, l0, l1, [&](uint64_t key) {
                if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xff], 1);
            });
            __syncthreads();
            if (tid == 0) {
                int cum = 0, b = 255;
                for (; b > 0; --b) {
                    const int c = s_hist[b];
                    if (cum + c >= remaining) break;
                    cum += c;
                }
                s_remaining = remaining - cum;
                s_prefix = prefix | (static_cast<uint64_t>(b) << shift);
                s_mask = mask | (0xffull << shift);
                s_done = (s_hist[b] == remaining - cum);   // the whole bin is needed: stop early
            }
            __syncthreads();
            prefix = s_prefix; mask = s_mask; remaining = s_remaining;
            if (s_done) break;
        }
        for_each_key(in, q, l0, l1, [&](uint64_t key) {
            const uint64_t km = key & mask;
            if (km > prefix) {
                s_keys[atomicAdd(&s_n, 1)] = key;
            } else if (km == prefix) {
                const int e = atomicAdd(&s_neq, 1);
                if (e < remaining) s_keys[k - remaining + e] = key;   // tail slots, disjoint from the > slots
            }
        });
    }
    __syncthreads();

    // bitonic sort (descending) of s_keys[0..kpad)
    for (int s = 2; s <= kpad; s <<= 1) {
        for (int t = s >> 1; t > 0; t >>= 1) {
            for (int i = tid; i < (kpad >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (t - 1));
                const int hi = lo + t;
                const bool desc = (lo & s) == 0;
                const uint64_t a = s_keys[lo], b = s_keys[hi];
                if ((a < b) == desc) { s_keys[lo] = b; s_keys[hi] = a; }
            }
            __syncthreads();
        }
    }

    const int nvalid = min(total, k);
    if (FINAL) {
        const float scale = out.q_scale ? out.q_scale[q] : 1.0f;
        if (out.D == nullptr) {
            // packed-key output for the cross-shard exchange: final score, 32-bit GLOBAL row id, 0 = padding
            uint64_t* K = reinterpret_cast<uint64_t*>(out.I);
            const uint32_t off = static_cast<uint32_t>(out.id_offset);
            for (int i = tid; i < k; i += blockDim.x) {
                const uint64_t key = s_keys[i];
                K[q * k + i] = (i < nvalid) ? make_key(key_score(key) * scale, key_row(key) + off) : 0ull;
            }
            return;
        }
        for (int i = tid; i < k; i += blockDim.x) {
            const uint64_t key = s_keys[i];
            const bool ok = i < nvalid;
            out.D[q * k + i] = ok ? key_score(key) * scale : -3.402823466e+38f;
            out.I[q * k + i] = ok ? static_cast<int64_t>(key_row(key)) + out.id_offset : -1;
        }
    } else {
        uint64_t* dst = out.entries + (static_cast<int64_t>(grp) * out.nq + q) * k;
        for (int i = tid; i < k; i += blockDim.x) dst[i] = (i < nvalid) ? s_keys[i] : 0ull;
        if (tid == 0) out.counts[static_cast<int64_t>(grp) * out.nq + q] = nvalid;
    }
}

static int kpad_for(int k) { int p = 2; while (p < k) p <<= 1; return p; }

size_t merge_tmp_entries(int n_lists, int64_t nq, int k) {
    // levels shrink by kMergeFanIn; two ping-pong buffers sized for the first level
    if (n_lists <= kMergeFanIn) return 0;
    const int64_t g1 = (n_lists + kMergeFanIn - 1) / kMergeFanIn;
    const int64_t g2 = (g1 + kMergeFanIn - 1) / kMergeFanIn;
    return static_cast<size_t>((g1 + g2) * nq * k);
}

int merge_lists_final(const MergeIn& in0, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                      int64_t id_offset, uint64_t* tmp_entries, int* tmp_counts,
                      cudaStream_t st, int* n_launches, const float* q_scale) {
    if (nq <= 0) return IVR_OK;
    const int kpad = kpad_for(k);
    const size_t smem = static_cast<size_t>(kpad) * sizeof(uint64_t);
    MergeIn in = in0;
    int level = 0;
    uint64_t* ebuf = tmp_entries;
    int*      cbuf = tmp_counts;
    while (in.n_lists > kMergeFanIn) {
        const int groups = (in.n_lists + kMergeFanIn - 1) / kMergeFanIn;
        if (!ebuf || !cbuf) { set_error("merge: scratch missing for %d lists", in.n_lists); return IVR_EINVAL; }
        MergeOut out{};
        out.entries = ebuf; out.counts = cbuf; out.nq = nq;
        dim3 grid(static_cast<unsigned>(nq), static_cast<unsigned>(groups));
        merge_kernel<false><<<grid, kMergeThreads, smem, st>>>(in, out, kMergeFanIn, k, kpad);
        IVR_CUDA(cudaGetLastError());
        if (n_launches) ++*n_launches;
        MergeIn nx{};
        nx.entries = ebuf; nx.counts = cbuf;
        nx.list_stride = nq * k; nx.q_stride = k;
        nx.cnt_list_stride = nq; nx.cnt_q_stride = 1;
        nx.n_lists = groups; nx.fixed_count = 0; nx.raw = 0;
        ebuf += static_cast<size_t>(groups) * nq * k;
        cbuf += static_cast<size_t>(groups) * nq;
        in = nx;
        if (++level > 4) { set_error("merge: too many levels"); return IVR_EINVAL; }
    }
    MergeOut out{};
    out.D = D_dev; out.I = I_dev; out.id_offset = id_offset; out.nq = nq; out.q_scale = q_scale;
    dim3 grid(static_cast<unsigned>(nq), 1);
    merge_kernel<true><<<grid, kMergeThreads, smem, st>>>(in, out, in.n_lists, k, kpad);
    IVR_CUDA(cudaGetLastError());
    if (n_launches) ++*n_launches;
    return IVR_OK;
}

// Same reduction, but the result stays in key form: out_keys [nq, k] (descending, 0-padded) and
// out_counts [nq].  Used between the two phases of the batched search.
int merge_lists_keys(const MergeIn& in0, int64_t nq, int k, uint64_t* out_keys, int* out_counts,
                     uint64_t* tmp_entries, int* tmp_counts, cudaStream_t st, int* n_launches) {
    if (nq <= 0) return IVR_OK;
    const int kpad = kpad_for(k);
    const size_t smem = static_cast<size_t>(kpad) * sizeof(uint64_t);
    MergeIn in = in0;
    uint64_t* ebuf = tmp_entries;
    int*      cbuf = tmp_counts;
    for (int level = 0; ; ++level) {
        const bool last = in.n_lists <= kMergeFanIn;
        const int groups = last ? 1 : (in.n_lists + kMergeFanIn - 1) / kMergeFanIn;
        if (!last && (!ebuf || !cbuf)) { set_error("merge: scratch missing for %d lists", in.n_lists); return IVR_EINVAL; }
        MergeOut out{};
        out.entries = last ? out_keys : ebuf; out.counts = last ? out_counts : cbuf; out.nq = nq;
        dim3 grid(static_cast<unsigned>(nq), static_cast<unsigned>(groups));
        merge_kernel<false><<<grid, kMergeThreads, smem, st>>>(in, out, last ? in.n_lists : kMergeFanIn, k, kpad);
        IVR_CUDA(cudaGetLastError());
        if (n_launches) ++*n_launches;
        if (last) return IVR_OK;
        MergeIn nx{};
        nx.entries = ebuf; nx.counts = cbuf;
        nx.list_stride = nq * k; nx.q_stride = k;
        nx.cnt_list_stride = nq; nx.cnt_q_stride = 1;
        nx.n_lists = groups; nx.fixed_count = 0; nx.raw = 0;
        ebuf += static_cast<size_t>(groups) * nq * k;
        cbuf += static_cast<size_t>(groups) * nq;
        in = nx;
        if (level > 4) { set_error("merge: too many levels"); return IVR_EINVAL; }
    }
}

// ---------------------------------------------------------------------------
// (D, I) shard lists -> packed keys, for the post-all-gather merge.  The key holds a 32-bit row id:
// ids at or above 2^32 cannot be represented and raise the flag (the caller rejects the merge).
// ---------------------------------------------------------------------------
__global__ void pack_parts_kernel(const float* __restrict__ D, const int64_t* __restrict__ I,
                                  uint64_t* __restrict__ keys, int64_t n) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int64_t id = I[i];
    keys[i] = (id < 0 || id > 0xffffffffll) ? 0ull : make_key(D[i], static_cast<uint32_t>(id));
}

int pack_parts(const float* D, const int64_t* I, uint64_t* keys, int64_t n, cudaStream_t st) {
    if (n <= 0) return IVR_OK;
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    pack_parts_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(D, I, keys, n);
    IVR_CUDA(cudaGetLastError());
    return IVR_OK;
}

}  // namespace ivr

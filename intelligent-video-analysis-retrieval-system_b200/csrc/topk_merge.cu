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
#include "merge_select.cuh"

#include <algorithm>

namespace ivr {

constexpr int kMergeFanIn   = 64;     // lists folded by one CTA in a non-final level

static int kpad_for(int k) { int p = 2; while (p < k) p <<= 1; return p; }

template <bool FINAL>
__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(MergeIn in, MergeOut out, int lists_per_group, int k, int kpad) {
    extern __shared__ uint64_t s_keys[];               // kpad
    __shared__ MergeShared sh;
    const int64_t q = blockIdx.x;
    const int grp = blockIdx.y;
    const int l0 = grp * lists_per_group;
    const int l1 = min(l0 + lists_per_group, in.n_lists);
    const int tid = threadIdx.x;
    const int nvalid = block_select_sort(in, q, l0, l1, k, kpad, s_keys, sh);
    if (FINAL) {
        write_final(out, q, k, s_keys, nvalid);
    } else {
        uint64_t* dst = out.entries + (static_cast<int64_t>(grp) * out.nq + q) * k;
        for (int i = tid; i < k; i += blockDim.x) dst[i] = (i < nvalid) ? s_keys[i] : 0ull;
        if (tid == 0) out.counts[static_cast<int64_t>(grp) * out.nq + q] = nvalid;
    }
}

__global__ void __launch_bounds__(kMergeThreads)
merge_select_kernel(MergeIn in, MergeOut out, SelectArgs sa, int k, int kpad) {
    extern __shared__ uint64_t s_keys[];               // max(kpad, pool_cap)
    __shared__ SelectShared ss;
    select_merge_block(in, out, sa, k, kpad, blockIdx.y, blockIdx.x, gridDim.x, s_keys, ss);
}

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

// One-launch merge by the list-maxima bound (see merge_select_kernel).  `maxima` uses the strides of in.counts;
// `pool` holds nq * kSelectPoolCap keys, `pool_cnt` / `ticket` nq ints each, all zero on entry (the kernel re-zeroes them).
int merge_select_final(const MergeIn& in, const uint32_t* maxima, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                       int64_t id_offset, uint64_t* pool, int* pool_cnt, int* ticket, int sm_count, cudaStream_t st,
                       int* n_launches, const float* q_scale, const uint32_t* group_max, int n_groups) {
    if (nq <= 0) return IVR_OK;
    if (k > kSelectMaxK || nq > 65535) { set_error("merge_select: k=%d / nq=%lld outside its range", k, static_cast<long long>(nq)); return IVR_EINVAL; }
    const int kpad = kpad_for(k);
    const size_t smem = static_cast<size_t>(std::max(kpad, kSelectPoolCap)) * sizeof(uint64_t);
    MergeOut out{};
    out.D = D_dev; out.I = I_dev; out.id_offset = id_offset; out.nq = nq; out.q_scale = q_scale;
    SelectArgs sa{};
    sa.maxima = maxima; sa.pool = pool; sa.pool_cnt = pool_cnt; sa.ticket = ticket; sa.pool_cap = kSelectPoolCap;
    if (group_max) { sa.tmax = group_max; sa.t_stride_g = 1; sa.t_stride_q = n_groups; sa.n_t = n_groups; }
    else { sa.tmax = maxima; sa.t_stride_g = in.cnt_list_stride; sa.t_stride_q = in.cnt_q_stride; sa.n_t = in.n_lists; }
    // enough CTAs that the filter pass is spread over the machine, few enough that the redundant T search stays cheap
    const int ctas = std::max(1, std::min(in.n_lists / 16, std::max(1, 2 * sm_count / static_cast<int>(std::min<int64_t>(nq, 2 * sm_count)))));
    dim3 grid(static_cast<unsigned>(ctas), static_cast<unsigned>(nq));
    merge_select_kernel<<<grid, kMergeThreads, smem, st>>>(in, out, sa, k, kpad);
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

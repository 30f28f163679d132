// sequence.cu -- K6: similar-sequence search.
//
// Replaces TemporalAnalyzer.find_similar_sequences (core.py:3644-3702), which slides a window of
// `seq_len` frames over the target AND over the database and, for every (target_start, db_start)
// pair, averages `seq_len` sklearn cosine_similarity calls (core.py:3812-3832):
//
//     sim(t, j) = mean_i cos(T[t+i], D[j+i]),   hit iff sim >= threshold
//
// i.e. a diagonal band sum over the cosine matrix C = Tn * Dn^T.  Here:
//   1. seq_normalize_kernel   rows -> x / sqrt(sum x^2) (zero norm -> 1), fp32, one warp per row
//                             (sklearn normalises first, then dots);
//   2. seq_cosine_kernel      C = Tn * Dn^T in fp32 (64 x 64 x 16 shared-memory tiles, 4 x 4 per thread)
//                             for a chunk of database rows -- fp32 FMA on purpose: the threshold
//                             test must see the reference's precision, not fp16's;
//   3. seq_diag_kernel        one thread per (t, j): the diagonal mean, summed in the order NumPy's
//                             float32 np.mean uses (pairwise_sum: straight loop below 8 elements, 8
//                             interleaved accumulators up to 128), threshold, atomic append of the hit.
// The database is processed in chunks so that the cosine block stays below kMaxBlockBytes.
// Algorithmic work: 2 * nt * nd * d FLOP.
//
// The same normalise + cosine-matrix kernels give the eps-neighbourhoods of the DBSCAN phase of
// filter_research_update.py (cluster_similar_frames, 113-134): seq_adjacency_kernel packs
// `1 - cos <= eps` into one bit per frame pair; the labelling itself is sequential and runs on the host.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace ivr {

constexpr int    kSeqMaxLen     = 128;                    // NumPy's pairwise_sum block: one level, no recursion
constexpr size_t kMaxBlockBytes = 256u << 20;             // cosine block per chunk

__global__ void seq_normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (r >= n) return;
    const float* row = x + r * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s = fmaf(row[c], row[c], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float nrm = sqrtf(s);
    if (nrm == 0.f) nrm = 1.f;
    for (int c = lane; c < d; c += 32) out[r * d + c] = __fdiv_rn(row[c], nrm);
}

// C[a, b] = <A[a, :], B[b, :]>, A: [M, K], B: [N, K] row-major, C: [M, ldc]
__global__ void __launch_bounds__(256)
seq_cosine_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                  int M, int64_t N, int K, int64_t ldc) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * 64, n0 = static_cast<int64_t>(blockIdx.x) * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {       // transposed tile loads (k fastest in global)
            const int r = e >> 4, k = e & 15;
            As[k][r] = (m0 + r < M && k0 + k < K) ? A[(m0 + r) * K + k0 + k] : 0.f;
            Bs[k][r] = (n0 + r < N && k0 + k < K) ? B[(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[m * ldc + n] = acc[i][j];
        }
}

// float32 np.mean of v(0..n): NumPy's pairwise_sum for n <= 128, then one float32 division
template <typename F>
__device__ __forceinline__ float numpy_mean_f32(int n, F&& v) {
    float res;
    if (n < 8) {
        res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, v(i));
    } else {
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = v(k);
        int i = 8;
        for (; i < n - (n % 8); i += 8)
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], v(i + k));
        res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, v(i));
    }
    return __fdiv_rn(res, static_cast<float>(n));
}

// C block: rows = target frames [0, nt), columns = database frames [j0, j0 + ncols)
__global__ void seq_diag_kernel(const float* __restrict__ C, int nt, int64_t ncols, int64_t ldc, int64_t j0,
                                int seq_len, float thr, int64_t n_starts_db, int64_t max_hits,
                                int32_t* __restrict__ hit_t, int64_t* __restrict__ hit_j, float* __restrict__ hit_sim,
                                unsigned long long* __restrict__ n_hits) {
    const int64_t jl = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // window start inside the block
    const int t = blockIdx.y;
    if (jl + seq_len > ncols || j0 + jl >= n_starts_db || t + seq_len > nt) return;
    const float sim = numpy_mean_f32(seq_len, [&](int i) { return C[(t + i) * ldc + jl + i]; });
    if (sim >= thr) {
        const unsigned long long pos = atomicAdd(n_hits, 1ull);
        if (pos < static_cast<unsigned long long>(max_hits)) {
            hit_t[pos] = t; hit_j[pos] = j0 + jl; hit_sim[pos] = sim;
        }
    }
}

// All pointers are device pointers; *n_hits_dev must be zeroed by the caller.
int sequence_similarity_device(const float* target, int64_t nt, const float* db, int64_t nd, int dim, int seq_len,
                               float thr, int64_t max_hits, int32_t* hit_t, int64_t* hit_j, float* hit_sim,
                               unsigned long long* n_hits_dev, float* tn, float* dn, float* cblock,
                               int64_t block_cols, cudaStream_t st) {
    {
        const int64_t threads_t = nt * 32, threads_d = nd * 32;
        seq_normalize_kernel<<<static_cast<unsigned>((threads_t + 255) / 256), 256, 0, st>>>(target, tn, nt, dim);
        seq_normalize_kernel<<<static_cast<unsigned>((threads_d + 255) / 256), 256, 0, st>>>(db, dn, nd, dim);
        IVR_CUDA(cudaGetLastError());
    }
    const int64_t n_starts_db = nd - seq_len + 1;
    const int64_t step = block_cols - (seq_len - 1);              // window starts covered per chunk
    for (int64_t j0 = 0; j0 < n_starts_db; j0 += step) {
        const int64_t ncols = std::min<int64_t>(block_cols, nd - j0);
        dim3 g1(static_cast<unsigned>((ncols + 63) / 64), static_cast<unsigned>((nt + 63) / 64));
        seq_cosine_kernel<<<g1, 256, 0, st>>>(tn, dn + j0 * dim, cblock, static_cast<int>(nt), ncols, dim, block_cols);
        dim3 g2(static_cast<unsigned>((ncols + 255) / 256), static_cast<unsigned>(nt - seq_len + 1));
        seq_diag_kernel<<<g2, 256, 0, st>>>(cblock, static_cast<int>(nt), ncols, block_cols, j0, seq_len, thr, n_starts_db,
                                           max_hits, hit_t, hit_j, hit_sim, n_hits_dev);
        IVR_CUDA(cudaGetLastError());
    }
    return IVR_OK;
}

int64_t sequence_block_cols(int64_t nt, int64_t nd, int seq_len) {
    int64_t cols = static_cast<int64_t>(kMaxBlockBytes / sizeof(float)) / std::max<int64_t>(nt, 1);
    if (const char* e = getenv("IVR_SEQ_BLOCK_COLS")) { if (*e) cols = atoll(e); }   // tests: force several chunks
    cols = std::max<int64_t>(cols, 2 * static_cast<int64_t>(seq_len));
    return std::min(cols, nd);
}

int sequence_max_len() { return kSeqMaxLen; }

// ---- eps-neighbourhoods for the DBSCAN phase (filter_research_update.py:113-127) --------------------
// adj[i][w] bit b  <=>  1 - cos(e_i, e_{32 w + b}) <= eps   (float32, as NumPy evaluates `1 - sim <= eps`)
__global__ void seq_adjacency_kernel(const float* __restrict__ C, int n, float eps, uint32_t* __restrict__ adj, int words) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;   // one warp per (row, word)
    if (wid >= static_cast<int64_t>(n) * words) return;
    const int i = static_cast<int>(wid / words), w = static_cast<int>(wid % words);
    const int j = w * 32 + lane;
    const bool in = j < n && __fsub_rn(1.0f, C[static_cast<int64_t>(i) * n + j]) <= eps;
    const unsigned bits = __ballot_sync(0xffffffffu, in);
    if (lane == 0) adj[static_cast<int64_t>(i) * words + w] = bits;
}

// e: device float32 [n, d]; xn: [n, d] scratch; cmat: [n, n] scratch; adj: [n, words] out
int cosine_neighbors_device(const float* e, int n, int dim, float eps, float* xn, float* cmat, uint32_t* adj,
                            cudaStream_t st) {
    const int words = (n + 31) / 32;
    const int64_t threads = static_cast<int64_t>(n) * 32;
    seq_normalize_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(e, xn, n, dim);
    dim3 g1(static_cast<unsigned>((n + 63) / 64), static_cast<unsigned>((n + 63) / 64));
    seq_cosine_kernel<<<g1, 256, 0, st>>>(xn, xn, cmat, n, n, dim, n);
    const int64_t t2 = static_cast<int64_t>(n) * words * 32;
    seq_adjacency_kernel<<<static_cast<unsigned>((t2 + 255) / 256), 256, 0, st>>>(cmat, n, eps, adj, words);
    IVR_CUDA(cudaGetLastError());
    return IVR_OK;
}

}  // namespace ivr

// index.cuh -- the flat inner-product index handle (device-resident fp16 rows).
#pragma once

#include "common.cuh"

struct ivr_index {
    int      dim      = 0;        // logical dimension (as given by the caller)
    int      dpad     = 0;        // storage dimension: dim padded to a multiple of 64
    int      device   = 0;
    int      sm_count = 0;
    int64_t  ntotal   = 0;
    int64_t  capacity = 0;        // rows allocated
    __half* rows = nullptr;          // [capacity, dpad] row-major fp16, HBM

    // search window (ivr_index_set_window): searches score rows [win_first, win_first + win_count) only;
    // win_count < 0 = the whole index
    int64_t  win_first = 0, win_count = -1;

    cudaStream_t stream = nullptr;   // handle-owned stream for the host-pointer entry points

    // scratch, grown on demand (device)
    void*   ws      = nullptr;
    size_t  ws_bytes = 0;
    // device staging for the host-pointer search entry point (queries in, D/I out)
    void*   io       = nullptr;
    size_t  io_bytes = 0;
    // pinned host staging
    void*   pin      = nullptr;
    size_t  pin_bytes = 0;
    // counters / tickets of the one-launch select merge (topk_merge.cu): zeroed at creation, re-zeroed by the kernel itself
    int*    sel_state = nullptr;

    // TMA descriptor cache for the MMA path (opaque 128-byte CUtensorMap blobs)
    alignas(64) unsigned char tmap_rows[128];
    const void* tmap_rows_base = nullptr;
    int64_t     tmap_rows_n    = -1;
    int         tmap_rows_box  = 0;

    // recorded after every device-side add on its stream; searches on other streams wait on it
    cudaEvent_t rows_ready = nullptr;
    bool        rows_ready_set = false;

    // timing of the last device search
    bool        timing = false;
    cudaEvent_t ev[6]  = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool        ev_valid[3] = {false, false, false};
    int         launches[3] = {0, 0, 0};
    int         last_path = 0;
    const char* last_kernel = "";
};

namespace ivr {

int  ensure_ws(ivr_index* idx, size_t bytes);
int  ensure_pin(ivr_index* idx, size_t bytes);
int  ensure_io(ivr_index* idx, size_t bytes);
int  ensure_capacity(ivr_index* idx, int64_t rows, cudaStream_t st, bool exact = false);

// search back-ends (search_stream.cu / search_mma.cu / topk_merge.cu)
int search_stream(ivr_index* idx, const float* q_dev, int64_t nq, int k, float* D_dev,
                  int64_t* I_dev, int64_t id_offset, cudaStream_t st);
int search_mma(ivr_index* idx, const float* q_dev, int64_t nq, int k, float* D_dev,
               int64_t* I_dev, int64_t id_offset, cudaStream_t st);
bool mma_supported(const ivr_index* idx, int64_t nq, int k);
bool mma_small_supported(const ivr_index* idx, int64_t nq, int k);   // small-batch kernel (search_mma_small.cu)

// Generic list merge.  For query q, list l lives at entries + l*list_stride + q*q_stride
// (64-bit keys) and holds counts[l*cnt_list_stride + q*cnt_q_stride] entries (or
// `fixed_count` entries when counts == nullptr).
struct MergeIn {
    const uint64_t* entries;
    const int*      counts;
    int64_t list_stride, q_stride;
    int64_t cnt_list_stride, cnt_q_stride;
    int     n_lists;
    int     fixed_count;
    int     raw;             // entries are raw {score bits, row} pairs (converted to keys on load)
    int     interleave;      // lists of 32 consecutive queries are interleaved: entry i of query q lives at
                             // l*list_stride + (q / 32) * q_stride * 32 + i * 32 + q % 32
    // DENSE mode (dense != nullptr; `entries` unused): list l of query q is the `dense_len` consecutive SCORES
    // dense[q * dense_q_stride + l * dense_len + i], i.e. rows l * dense_len + i (< dense_rows) of a materialised
    // score matrix; keys are formed on load.
    const float* dense;
    int64_t      dense_q_stride, dense_rows;
    int          dense_len;
};
// final stage: writes D/I ([nq,k], padded with -FLT_MAX / -1); ids = row + id_offset;
// scores are multiplied by q_scale[q] when q_scale != nullptr (power-of-two query scaling)
int merge_lists_final(const MergeIn& in, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                      int64_t id_offset, uint64_t* tmp_entries, int* tmp_counts,
                      cudaStream_t st, int* n_launches, const float* q_scale = nullptr);
// one-launch merge for many short lists using the per-list maxima (topk_merge.cu); k <= kSelectMaxK
constexpr int kSelectMaxK    = 1024;
constexpr int kSelectPoolCap = 4096;
int merge_select_final(const MergeIn& in, const uint32_t* maxima, int64_t nq, int k, float* D_dev, int64_t* I_dev,
                       int64_t id_offset, uint64_t* pool, int* pool_cnt, int* ticket, int sm_count, cudaStream_t st,
                       int* n_launches, const float* q_scale = nullptr,
                       const uint32_t* group_max = nullptr, int n_groups = 0);   // [nq][n_groups]: bound searched in these
int merge_lists_keys(const MergeIn& in, int64_t nq, int k, uint64_t* out_keys, int* out_counts,
                     uint64_t* tmp_entries, int* tmp_counts, cudaStream_t st, int* n_launches);
size_t merge_tmp_entries(int n_lists, int64_t nq, int k);   // #keys of scratch the merge may need

}  // namespace ivr

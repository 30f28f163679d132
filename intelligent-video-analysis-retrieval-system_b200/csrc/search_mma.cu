// search_mma.cu -- placeholder until the tcgen05 kernel lands (replaced below).
#include "index.cuh"
namespace ivr {
bool mma_supported(const ivr_index*, int64_t, int) { return false; }
int search_mma(ivr_index*, const float*, int64_t, int, float*, int64_t*, int64_t, cudaStream_t) {
    set_error("tcgen05 path not built");
    return IVR_EUNSUPPORTED;
}
}  // namespace ivr

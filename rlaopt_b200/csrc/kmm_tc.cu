// tcgen05 tensor-core path — placeholder until the kernel lands (reports "unsupported").
#include "kmm_common.cuh"
#include "kmm_launch.h"

namespace kmm {
bool tc_supported_d(int64_t) { return false; }
bool tc_supported_k(int64_t) { return false; }
size_t tc_packed_bytes(int64_t, int64_t) { return 0; }
cudaError_t launch_tc_pack(const float*, int64_t, int64_t, int64_t, const int64_t*, float, const float*, void*,
                           cudaStream_t) {
    return cudaErrorNotSupported;
}
size_t tc_workspace_bytes(int64_t, int64_t, int64_t, int64_t, int) { return 0; }
cudaError_t launch_tc(const void*, int64_t, const void*, int64_t, int64_t, const float*, int64_t, int64_t, float*,
                      int64_t, int, float, int, void*, size_t, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace kmm

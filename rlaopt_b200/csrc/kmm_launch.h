// Internal launch interface between the C-ABI (kmm_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace kmm {

// ---- packing (kmm_pack.cu): row-major points -> feature-major, 1/lengthscale applied ----
// idx: optional gather list (negative entries wrap once, out-of-range entries pack as zero points)
template <typename T>
cudaError_t launch_pack(const T* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx, T inv_ls,
                        const T* inv_ls_vec, T* packed, cudaStream_t stream);

// column means of X[idx] (fp64 accumulation, fixed summation order): the common shift of the tensor-core packs
size_t column_mean_workspace_bytes(int64_t n, int64_t d);
cudaError_t launch_column_mean(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                               float* center, void* workspace, cudaStream_t stream);

// ---- CUDA-core fused matmat (kmm_simt.cu) ----
template <typename T>
struct SimtArgs {
    const T* Rt;  // packed row operand   [d_pad][n_pad]
    int64_t n, n_pad;
    const T* Ct;  // packed column operand [d_pad][m_pad]
    int64_t m, m_pad;
    int64_t d_pad;
    const T* V;  // [m][k], row stride ldv
    int64_t ldv, k;
    T* Y;  // [n][k], row stride ldy
    int64_t ldy;
    T scale;  // const_scaling
    int kid;
    cudaStream_t stream;
};

template <typename T>
size_t simt_workspace_bytes(int64_t n, int64_t m, int64_t k, int sm_count);

template <typename T>
cudaError_t launch_simt(const SimtArgs<T>& args, int sm_count, void* workspace, size_t workspace_bytes);

// sum of per-split partial results [splits][n][k] -> Y (deterministic, fixed order)
template <typename T>
cudaError_t launch_split_reduce(const T* part, int splits, int64_t n, int64_t k, T* Y, int64_t ldy, T scale,
                                cudaStream_t stream);

// ---- tcgen05 tensor-core fused matmat for the L2 kernels, fp32 in/out (kmm_tc.cu) ----
bool tc_supported_d(int64_t d);
bool tc_supported_k(int64_t k);
size_t tc_packed_bytes(int64_t n, int64_t d);
cudaError_t launch_tc_pack(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                           float inv_ls, const float* inv_ls_vec, const float* center, void* packed,
                           cudaStream_t stream);
size_t tc_stats_offset();  // byte offset of {max squared norm (float bits), bad-index count} in a packed operand
size_t tc_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count);
cudaError_t launch_tc(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                      const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kid, float scale,
                      int sm_count, void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace kmm

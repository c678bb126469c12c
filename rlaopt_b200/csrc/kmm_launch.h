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

// ---- fused output stage (kmm_fuse.cu): Y = alpha c K V + beta C[ci] + gamma B[bi], Gram L^T Y, column norms ----
template <typename T>
struct FuseArgs {
    T alpha;                       // multiplies const_scaling * (K V)
    T beta;
    const T* addend;               // C, row stride ld_addend, addend_rows rows; row of output i is addend_idx[i] (or i)
    int64_t ld_addend, addend_rows;
    const int64_t* addend_idx;
    T gamma;
    const T* rhs;                  // B, likewise
    int64_t ld_rhs, rhs_rows;
    const int64_t* rhs_idx;
    T* Y;                          // may be null: reductions only (no n x k result is written)
    int64_t ldy;
    const T* gram_lhs;             // L [n][gram_cols], row stride ld_gram_lhs; gram_out[gram_cols][k] = L^T Y
    int64_t ld_gram_lhs;
    int gram_cols;
    T* gram_out;
    int want_sqnorm;
    T* sqnorm_out;                 // [k]
};
template <typename T>
bool fuse_reductions_supported(int64_t k, int64_t gram_cols);
template <typename T>
size_t fuse_workspace_bytes(int64_t n, int64_t k, int64_t gram_cols, int want_sqnorm);
// part: un-scaled partial sums [splits][n][k]; workspace: fuse_workspace_bytes
template <typename T>
cudaError_t launch_fuse(const FuseArgs<T>& a, const T* part, int splits, int64_t n, int64_t k, T alpha, void* workspace,
                        cudaStream_t stream);

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

// keep_partials: leave the un-scaled column-split sums [splits][n][k] in the workspace (always >= 1 split then) and
// skip the split reduce -- the fused output stage consumes them
template <typename T>
size_t simt_workspace_bytes(int64_t n, int64_t m, int64_t k, int sm_count, bool keep_partials = false);

template <typename T>
cudaError_t launch_simt(const SimtArgs<T>& args, int sm_count, void* workspace, size_t workspace_bytes,
                        bool keep_partials = false, int* splits_out = nullptr);

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
size_t tc_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count, bool keep_partials = false);
// keep_partials: see launch_simt; *part_out receives the address of the partial sums inside the workspace
cudaError_t launch_tc(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                      const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kid, float scale,
                      int sm_count, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                      bool keep_partials = false, int* splits_out = nullptr, const float** part_out = nullptr);

}  // namespace kmm

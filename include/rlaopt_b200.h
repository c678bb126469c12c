/*
 * rlaopt_b200 — C ABI of the B200-native implicit kernel-matrix matmat
 *
 *     Y = const_scaling * K(A1[row_idx], A2[col_idx]) @ V        (or K^T @ V)
 *
 * for the RBF, Laplace and Matern-1/2, -3/2, -5/2 kernels.  This is the drop-in
 * boundary for the one hot path of udellgroup/rlaopt: every entry point below
 * replaces a PyKeOps LazyTensor reduction call site of the reference (the
 * reference has no FFI of its own for this path — it calls `K_lazy @ x` from
 * Python; file:line citations are relative to the reference repository).
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures (`stream` is a cudaStream_t
 *     passed as void*; NULL = legacy default stream);
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     name ends in `_host`; matrices are row-major with an explicit row stride
 *     (`ld*`, in elements);
 *   - functions never allocate, free or synchronise; they enqueue work on
 *     `stream` and return; all buffers are caller-owned;
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     RLAOPT_B200_E* code; `rlaopt_b200_last_error()` gives the message
 *     (thread-local).  The Python host raises RuntimeError on non-zero, matching
 *     the TORCH_CHECK -> RuntimeError behaviour of the reference's own ops
 *     (rlaopt/csrc/cpp/input_checks.cpp:9-69).
 */
#ifndef RLAOPT_B200_H_
#define RLAOPT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLAOPT_B200_ABI_VERSION 2

/* kernel ids — formulas of rlaopt/kernels/standard.py:46-85 */
#define RLAOPT_B200_KERNEL_RBF 0      /* exp(-|u|_2^2 / 2)                      standard.py:46-52 */
#define RLAOPT_B200_KERNEL_LAPLACE 1  /* exp(-|u|_1)                            standard.py:55-61 */
#define RLAOPT_B200_KERNEL_MATERN12 2 /* exp(-r)                                standard.py:64-69 */
#define RLAOPT_B200_KERNEL_MATERN32 3 /* (1 + sqrt3 r) exp(-sqrt3 r)            standard.py:72-77 */
#define RLAOPT_B200_KERNEL_MATERN52 4 /* (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)  standard.py:80-85 */
/*   u = (x_i - y_j) / lengthscale  (scalar or per-feature), r = |u|_2           standard.py:31-43 */

/* operand layouts produced by rlaopt_b200_pack_points_* */
#define RLAOPT_B200_LAYOUT_SIMT 0 /* feature-major fp32/fp64, CUDA-core path (all kernels)       */
#define RLAOPT_B200_LAYOUT_TC 1   /* split-precision tiles + norms, tcgen05 path (L2 kernels, fp32) */

#define RLAOPT_B200_EINVAL (-1)      /* bad argument                      */
#define RLAOPT_B200_EWORKSPACE (-2)  /* workspace too small               */
#define RLAOPT_B200_EUNSUPPORTED (-3)/* shape / kernel not supported by the requested layout */

int rlaopt_b200_abi_version(void);
const char* rlaopt_b200_last_error(void);

/* Number of SMs of the current device (cached); also proves a usable CUDA device. */
int rlaopt_b200_device_sm_count(void);

/* 1 if (kernel, element size, d, k) can run on `layout`, else 0.  LAYOUT_SIMT supports
 * everything; LAYOUT_TC supports the L2 kernels in fp32 for the d / k ranges its tiles cover.
 * Pure host logic (no CUDA call). */
int rlaopt_b200_layout_supported(int kernel_id, int elem_bytes, int64_t d, int64_t k, int layout);

/* ---------------------------------------------------------------------------
 * Packed operands.
 *
 * A set of points X[n][d] is packed once per operator (or once per oracle
 * block) into the kernel's streaming layout: optional row gather, division by
 * the lengthscale, transposition, padding.  Replaces
 *   LazyTensor(A1[:, None, :]) / LazyTensor(A2[None, :, :])   rlaopt/kernels/base.py:88-99
 *   A1[blk].to(device), A2[blk].to(device)                    rlaopt/kernels/utils.py:23,47-48
 *   (x - y) / lengthscale                                     rlaopt/kernels/standard.py:31-35
 * ------------------------------------------------------------------------- */

/* Bytes of the packed form of n points with d features (elem_bytes = 4 or 8). */
size_t rlaopt_b200_packed_bytes(int64_t n, int64_t d, int elem_bytes, int layout);

/* X has n_src rows; idx: optional int64 gather list of length n (rows X[idx[i]]; negative entries wrap once
 * like Python indices; an entry still outside [0, n_src) packs as a zero point and is counted, see
 * rlaopt_b200_packed_stats_host), NULL = identity (then n <= n_src).
 * inv_lengthscale_vec: optional per-feature 1/lengthscale (length d), NULL = use the scalar.
 * center: optional shift (length d, units of X) subtracted from every point before the division by the
 * lengthscale.  All five kernels are functions of x - y (standard.py:31-43), so the SAME center on both
 * operands leaves K unchanged; the tensor-core layout needs it for uncentred data, because its GEMM-form
 * distance |x|^2 + |y|^2 - 2 x.y carries an absolute error ~3e-7 (|x|^2 + |y|^2).  LAYOUT_SIMT forms direct
 * differences (exact under translation, like the reference's KeOps formula) and ignores it. */
int rlaopt_b200_pack_points_f32(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                float inv_lengthscale, const float* inv_lengthscale_vec, const float* center,
                                int layout, void* packed, void* stream);
int rlaopt_b200_pack_points_f64(const double* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                double inv_lengthscale, const double* inv_lengthscale_vec, const double* center,
                                int layout, void* packed, void* stream);

/* Column means of X[idx] (fp64 accumulation, fixed summation order -> bit-reproducible): the center to pass to
 * both pack calls of one operator (mean of the column operand A2).  workspace: *_workspace_bytes(n, d). */
size_t rlaopt_b200_column_mean_workspace_bytes(int64_t n, int64_t d);
int rlaopt_b200_column_mean_f32(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                float* center, void* workspace, size_t workspace_bytes, void* stream);

/* Statistics of a LAYOUT_TC pack, copied to HOST memory; synchronises `stream` (the one exception to the
 * "never synchronises" rule, hence the suffix).  max_sqnorm = max_i |(x_i - center) / lengthscale|^2: the
 * tensor-core path keeps the 1e-5 parity bar while
 *     3e-7 * c_kernel * (max_sqnorm(rows) + max_sqnorm(cols)) <= 1e-5,   c = 0.5 RBF, 1.5 Matern-3/2, 5/6 Matern-5/2
 * beyond that the host falls back to LAYOUT_SIMT (rlaopt_b200.ops.tc_accuracy_ok).  bad_index = number of gather
 * indices that were out of range.  Either output pointer may be NULL. */
int rlaopt_b200_packed_stats_host(const void* packed, int layout, float* max_sqnorm_host, int64_t* bad_index_host,
                                  void* stream);

/* ---------------------------------------------------------------------------
 * Fused matmat on packed operands:  Y[n][k] = const_scaling * K(rows, cols) @ V[m][k].
 *
 * Replaces the reductions
 *   K_lazy @ x            rlaopt/kernels/base.py:44, :112, :209 ; rlaopt/kernels/utils.py:29,56
 *   K_lazy.T @ x          rlaopt/kernels/base.py:47, :214   (call with rows/cols swapped)
 * and the post-scaling of rlaopt/linops/mixins.py:26-29.
 * k = 1 with ldv = ldy = 1 is the matvec.  Workspace: see *_workspace_bytes.
 * ------------------------------------------------------------------------- */
size_t rlaopt_b200_matmat_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int elem_bytes, int layout);

int rlaopt_b200_matmat_packed_f32(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                                  const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kernel_id,
                                  float const_scaling, int layout, void* workspace, size_t workspace_bytes,
                                  void* stream);
int rlaopt_b200_matmat_packed_f64(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                                  const double* V, int64_t k, int64_t ldv, double* Y, int64_t ldy, int kernel_id,
                                  double const_scaling, int layout, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* ---------------------------------------------------------------------------
 * Fused output stage: the same product with the callers' next passes folded in,
 *
 *     Y[i,:]     = alpha * const_scaling * (K V)[i,:] + beta * addend[ai(i),:] + gamma * rhs[ri(i),:]
 *     gram_out   = gram_lhs^T Y     (gram_cols x k, optional; k, gram_cols <= 64)
 *     sqnorm_out = column sums of Y^2 (k, optional; k <= 64)
 *
 * with ai(i) = addend_idx[i] (or i when NULL; negative entries wrap, out-of-range rows contribute 0), same for rhs.
 * One pass, no n x k temporary (Y may be NULL when only the reductions are wanted); the reductions are
 * deterministic (fixed summation order, fp64 across blocks).  Replaces
 *     A @ P + reg * P   and   P^T (A P)                 rlaopt/solvers/pcg.py:58-61   (beta = reg, addend = gram_lhs = P)
 *     B - (A @ W + reg * W), its column norms           rlaopt/models/linsys.py:96-99, rlaopt/solvers/pcg.py:33
 *                                                       (alpha = -1, beta = -reg, addend = W, gamma = 1, rhs = B)
 *     A_row_oracle(blk) @ Y + reg * Y[blk] - B[blk]     rlaopt/solvers/sap.py:113-127 (addend_idx = rhs_idx = blk)
 * Workspace: rlaopt_b200_matmat_fused_workspace_bytes.
 * ------------------------------------------------------------------------- */
typedef struct rlaopt_b200_epilogue_f32 {
    float alpha;
    float beta;
    const float* addend;       /* NULL = no term */
    int64_t ld_addend, addend_rows;
    const int64_t* addend_idx;
    float gamma;
    const float* rhs;          /* NULL = no term */
    int64_t ld_rhs, rhs_rows;
    const int64_t* rhs_idx;
    const float* gram_lhs;     /* [n][gram_cols], NULL = no Gram */
    int64_t ld_gram_lhs;
    int64_t gram_cols;
    float* gram_out;           /* [gram_cols][k], row-major */
    float* sqnorm_out;         /* [k], NULL = not wanted */
} rlaopt_b200_epilogue_f32;

typedef struct rlaopt_b200_epilogue_f64 {
    double alpha;
    double beta;
    const double* addend;
    int64_t ld_addend, addend_rows;
    const int64_t* addend_idx;
    double gamma;
    const double* rhs;
    int64_t ld_rhs, rhs_rows;
    const int64_t* rhs_idx;
    const double* gram_lhs;
    int64_t ld_gram_lhs;
    int64_t gram_cols;
    double* gram_out;
    double* sqnorm_out;
} rlaopt_b200_epilogue_f64;

size_t rlaopt_b200_matmat_fused_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int elem_bytes, int layout,
                                                int64_t gram_cols, int want_sqnorm);

int rlaopt_b200_matmat_packed_fused_f32(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m,
                                        int64_t d, const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy,
                                        int kernel_id, float const_scaling, int layout,
                                        const rlaopt_b200_epilogue_f32* epilogue, void* workspace,
                                        size_t workspace_bytes, void* stream);
int rlaopt_b200_matmat_packed_fused_f64(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m,
                                        int64_t d, const double* V, int64_t k, int64_t ldv, double* Y, int64_t ldy,
                                        int kernel_id, double const_scaling, int layout,
                                        const rlaopt_b200_epilogue_f64* epilogue, void* workspace,
                                        size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * One-shot entry: pack both operands into `workspace`, then the fused matmat.
 *   forward   (transpose = 0): Y[n'][k] = c * K(A1[row_idx], A2[col_idx])   @ V[m'][k]
 *   transpose (transpose = 1): Y[m'][k] = c * K(A1[row_idx], A2[col_idx])^T @ V[n'][k]
 * with n' = n_idx if row_idx else n, m' = m_idx if col_idx else m.
 * This is the whole of _KernelLinOp's matvec / rmatvec / row_oracle / blk_oracle
 * (rlaopt/kernels/base.py:43-47, 104-128) in one call.  With LAYOUT_TC the entry centres both operands on
 * the column means of A2[col_idx] (computed on the device, in the workspace).  It cannot evaluate the accuracy
 * guard (no synchronisation): callers with unnormalised data (|x / lengthscale|^2 >> 30 after centring) use
 * LAYOUT_SIMT, or the two-step API with rlaopt_b200_packed_stats_host.
 * ------------------------------------------------------------------------- */
size_t rlaopt_b200_kernel_matmat_workspace_bytes(int64_t n_rows, int64_t m_cols, int64_t d, int64_t k, int elem_bytes,
                                                 int layout);

int rlaopt_b200_kernel_matmat_f32(const float* A1, int64_t n, int64_t lda1, const float* A2, int64_t m, int64_t lda2,
                                  int64_t d, const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy,
                                  int kernel_id, float inv_lengthscale, const float* inv_lengthscale_vec,
                                  float const_scaling, int transpose, const int64_t* row_idx, int64_t n_idx,
                                  const int64_t* col_idx, int64_t m_idx, int layout, void* workspace,
                                  size_t workspace_bytes, void* stream);
int rlaopt_b200_kernel_matmat_f64(const double* A1, int64_t n, int64_t lda1, const double* A2, int64_t m,
                                  int64_t lda2, int64_t d, const double* V, int64_t k, int64_t ldv, double* Y,
                                  int64_t ldy, int kernel_id, double inv_lengthscale,
                                  const double* inv_lengthscale_vec, double const_scaling, int transpose,
                                  const int64_t* row_idx, int64_t n_idx, const int64_t* col_idx, int64_t m_idx,
                                  int layout, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Host-buffer convenience entry (fp32): all matrix pointers are HOST pointers;
 * the call allocates device scratch, copies in, runs the one-shot entry,
 * copies Y back and synchronises.  It exists for callers without a device
 * allocator of their own and for the end-to-end leg of bench.py.
 * ------------------------------------------------------------------------- */
int rlaopt_b200_kernel_matmat_host_f32(const float* A1_host, int64_t n, const float* A2_host, int64_t m, int64_t d,
                                       const float* V_host, int64_t k, float* Y_host, int kernel_id,
                                       float inv_lengthscale, float const_scaling, int transpose, int layout);

#ifdef __cplusplus
}
#endif
#endif /* RLAOPT_B200_H_ */

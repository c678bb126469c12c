// Operand packing: row-major points X[n][d] (optionally gathered through an index
// list) -> feature-major Xt[d_pad][n_pad] with 1/lengthscale folded in and zero
// padding, the layout every tile load of the fused kernels streams with 16-byte
// copies.  Replaces the per-call `A1[blk].to(device)` gathers of the reference
// (rlaopt/kernels/utils.py:23,47-48; rlaopt/kernels/base.py:92-99): the gather,
// the lengthscale division (rlaopt/kernels/standard.py:31-35) and the transpose
// happen in one pass, once per operator (or once per oracle block).
#include "kmm_common.cuh"
#include "kmm_launch.h"

namespace kmm {
namespace {

constexpr int TP = 32;

template <typename T>
__global__ void __launch_bounds__(TP * 8)
kmm_pack_kernel(const T* __restrict__ X, int64_t n, int64_t n_src, int64_t d, int64_t ldx,
                const int64_t* __restrict__ idx, T inv_ls, const T* __restrict__ inv_ls_vec, T* __restrict__ Xt,
                int64_t n_pad, int64_t d_pad) {
    __shared__ T tile[TP][TP + 1];
    const int64_t i0 = (int64_t)blockIdx.x * TP;
    const int64_t f0 = (int64_t)blockIdx.y * TP;
    // read: threadIdx.x walks features (contiguous in X), threadIdx.y walks points
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        const int64_t i = i0 + r, f = f0 + threadIdx.x;
        T v = T(0);
        if (i < n && f < d) {
            int64_t src = idx ? idx[i] : i;
            if (src < 0) src += n_src;  // Python-style wrap; anything still out of range packs as a zero point
            if (src >= 0 && src < n_src) {
                const T s = inv_ls_vec ? inv_ls_vec[f] : inv_ls;
                v = X[src * ldx + f] * s;
            }
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    // write: threadIdx.x walks points (contiguous in Xt)
    for (int r = threadIdx.y; r < TP; r += blockDim.y) {
        const int64_t f = f0 + r, i = i0 + threadIdx.x;
        if (f < d_pad && i < n_pad) Xt[f * n_pad + i] = tile[threadIdx.x][r];
    }
}

// ---- column means of a (gathered) point set, fp64, deterministic: per-block partial sums, then a fixed-order sum ----
constexpr int CM_ROWS_PER_BLOCK = 4096;

__global__ void __launch_bounds__(256)
column_partial_sum_kernel(const float* __restrict__ X, int64_t n, int64_t n_src, int64_t d, int64_t ldx,
                          const int64_t* __restrict__ idx, double* __restrict__ partial) {
    // thread (fx, ry): feature fx + 32 * (blockIdx.y), rows ry, ry + 8, ... of this block's row range
    __shared__ double red[8][33];
    const int fx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t f = (int64_t)blockIdx.y * 32 + fx;
    const int64_t i0 = (int64_t)blockIdx.x * CM_ROWS_PER_BLOCK;
    const int64_t i1 = min(n, i0 + CM_ROWS_PER_BLOCK);
    double s = 0.0;
    if (f < d) {
        for (int64_t i = i0 + ry; i < i1; i += 8) {
            int64_t src = idx ? idx[i] : i;
            if (src < 0) src += n_src;
            if (src >= 0 && src < n_src) {
                const float v = X[src * ldx + f];
                if (fabsf(v) < 3.0e38f) s += (double)v;  // inf / nan do not poison the shift
            }
        }
    }
    red[ry][fx] = s;
    __syncthreads();
    if (ry == 0 && f < d) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][fx];
        partial[(int64_t)blockIdx.x * d + f] = t;
    }
}

__global__ void column_mean_finish_kernel(const double* __restrict__ partial, int64_t blocks, int64_t n, int64_t d,
                                          float* __restrict__ center) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= d) return;
    double t = 0.0;
    for (int64_t b = 0; b < blocks; ++b) t += partial[b * d + f];
    center[f] = (float)(t / (double)n);
}

}  // namespace

size_t column_mean_workspace_bytes(int64_t n, int64_t d) {
    if (n <= 0 || d <= 0) return 0;
    return (size_t)((n + CM_ROWS_PER_BLOCK - 1) / CM_ROWS_PER_BLOCK) * (size_t)d * sizeof(double);
}

cudaError_t launch_column_mean(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                               float* center, void* workspace, cudaStream_t stream) {
    const int64_t blocks = (n + CM_ROWS_PER_BLOCK - 1) / CM_ROWS_PER_BLOCK;
    double* partial = static_cast<double*>(workspace);
    dim3 grid((unsigned)blocks, (unsigned)((d + 31) / 32));
    column_partial_sum_kernel<<<grid, 256, 0, stream>>>(X, n, n_src, d, ldx, idx, partial);
    column_mean_finish_kernel<<<(unsigned)((d + 127) / 128), 128, 0, stream>>>(partial, blocks, n, d, center);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_pack(const T* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx, T inv_ls,
                        const T* inv_ls_vec, T* packed, cudaStream_t stream) {
    const int64_t n_pad = round_up(n, PACK_ROWS), d_pad = round_up(d, PACK_FEATS);
    if (n_pad == 0 || d_pad == 0) return cudaSuccess;
    dim3 grid((unsigned)(n_pad / TP), (unsigned)((d_pad + TP - 1) / TP));
    dim3 block(TP, 8);
    kmm_pack_kernel<T><<<grid, block, 0, stream>>>(X, n, n_src, d, ldx, idx, inv_ls, inv_ls_vec, packed, n_pad, d_pad);
    return cudaGetLastError();
}

template cudaError_t launch_pack<float>(const float*, int64_t, int64_t, int64_t, int64_t, const int64_t*, float,
                                        const float*, float*, cudaStream_t);
template cudaError_t launch_pack<double>(const double*, int64_t, int64_t, int64_t, int64_t, const int64_t*, double,
                                         const double*, double*, cudaStream_t);

}  // namespace kmm

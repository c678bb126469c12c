// Fused output stage of the kernel matmat:
//
//   Y[i, :] = alpha * c * (K V)[i, :] + beta * C[ci(i), :] + gamma * B[bi(i), :]
//   gram[a, b] = sum_i L[i, a] * Y[i, b]          (optional, k and gram_cols <= 64)
//   sqn[b]     = sum_i Y[i, b]^2                  (optional, k <= 64)
//
// which is, in one pass over the n x k product and without an n x k temporary,
//   A P + reg P and P^T A P                         rlaopt/solvers/pcg.py:58-61
//   B - (A W + reg W) and its column norms          rlaopt/models/linsys.py:96-99, rlaopt/solvers/pcg.py:33
//   A[blk, :] Y + reg Y[blk] - B[blk]               rlaopt/solvers/sap.py:113-127
//
// The main kernels (kmm_tc.cu / kmm_simt.cu) leave the un-scaled sums of their column splits in the workspace
// ([splits][n][k]); this stage replaces their split-reduce kernel, so the fused forms cost no extra pass.  The two
// reductions are deterministic: every block owns a fixed set of 64-row tiles and writes one partial, a second
// kernel adds the partials in block order (fp64).
#include "kmm_common.cuh"
#include "kmm_launch.h"

namespace kmm {
namespace {

constexpr int FU_ROWS = 64;      // rows per tile
constexpr int FU_THREADS = 256;
constexpr int FU_MAX_RED = 64;   // the reductions cover k <= 64 and gram_cols <= 64
constexpr int FU_MAX_BLOCKS = 592;

template <typename T>
__global__ void __launch_bounds__(FU_THREADS)
kmm_fuse_kernel(const FuseArgs<T> a, const T* __restrict__ part, int splits, int64_t n, int k, T alpha,
                T* __restrict__ red_part) {
    // smem (reductions only): Ys[64][k] then Ls[64][gram_cols]
    extern __shared__ __align__(16) unsigned char fuse_smem[];
    T* Ys = reinterpret_cast<T*>(fuse_smem);
    T* Ls = Ys + FU_ROWS * k;
    const bool want_gram = a.gram_lhs != nullptr && a.gram_cols > 0;
    const bool want_sqn = a.want_sqnorm != 0;
    const bool reduce = want_gram || want_sqn;
    const int kg = want_gram ? a.gram_cols : 0;
    const int tid = threadIdx.x;
    const int64_t nk = n * (int64_t)k;
    const int64_t tiles = (n + FU_ROWS - 1) / FU_ROWS;

    constexpr int MAXP = FU_MAX_RED * FU_MAX_RED / FU_THREADS;  // Gram entries per thread
    T gacc[MAXP];
#pragma unroll
    for (int p = 0; p < MAXP; ++p) gacc[p] = T(0);
    T sacc = T(0);

    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t row0 = t * FU_ROWS;
        const int rows = (int)min((int64_t)FU_ROWS, n - row0);
        for (int e = tid; e < rows * k; e += FU_THREADS) {
            const int r = e / k, c = e - r * k;
            const int64_t i = row0 + r;
            T y = T(0);
            for (int z = 0; z < splits; ++z) y += part[(int64_t)z * nk + i * k + c];
            y *= alpha;
            if (a.addend) {
                int64_t s = a.addend_idx ? a.addend_idx[i] : i;
                if (s < 0) s += a.addend_rows;
                if (s >= 0 && s < a.addend_rows) y += a.beta * a.addend[s * a.ld_addend + c];
            }
            if (a.rhs) {
                int64_t s = a.rhs_idx ? a.rhs_idx[i] : i;
                if (s < 0) s += a.rhs_rows;
                if (s >= 0 && s < a.rhs_rows) y += a.gamma * a.rhs[s * a.ld_rhs + c];
            }
            if (a.Y) a.Y[i * a.ldy + c] = y;
            if (reduce) Ys[r * k + c] = y;
        }
        if (!reduce) continue;
        if (want_gram) {
            for (int e = tid; e < rows * kg; e += FU_THREADS) {
                const int r = e / kg, c = e - r * kg;
                Ls[r * kg + c] = a.gram_lhs[(row0 + r) * a.ld_gram_lhs + c];
            }
        }
        __syncthreads();
        if (want_gram) {
#pragma unroll
            for (int p = 0; p < MAXP; ++p) {
                const int q = tid + p * FU_THREADS;  // entry (ga, gb) = (q / k, q % k): gb fastest -> conflict-free Ys reads
                if (q < kg * k) {
                    const int ga = q / k, gb = q - ga * k;
                    T s = T(0);
                    for (int r = 0; r < rows; ++r) s += Ls[r * kg + ga] * Ys[r * k + gb];
                    gacc[p] += s;
                }
            }
        }
        if (want_sqn && tid < k) {
            T s = T(0);
            for (int r = 0; r < rows; ++r) s += Ys[r * k + tid] * Ys[r * k + tid];
            sacc += s;
        }
        __syncthreads();
    }
    if (!reduce) return;
    // block partial: [kg * k Gram entries | k squared column norms]
    T* mine = red_part + (int64_t)blockIdx.x * ((int64_t)kg * k + k);
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
        const int q = tid + p * FU_THREADS;
        if (q < kg * k) mine[q] = gacc[p];
    }
    if (tid < k) mine[(int64_t)kg * k + tid] = want_sqn ? sacc : T(0);
}

template <typename T>
__global__ void kmm_fuse_reduce_kernel(const T* __restrict__ red_part, int blocks, int kg, int k, T* __restrict__ gram_out,
                                       T* __restrict__ sqn_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = kg * k + k;
    if (e >= per) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += (double)red_part[(int64_t)b * per + e];
    if (e < kg * k) {
        if (gram_out) gram_out[e] = (T)s;
    } else if (sqn_out) {
        sqn_out[e - kg * k] = (T)s;
    }
}

int fuse_blocks(int64_t n) {
    const int64_t tiles = (n + FU_ROWS - 1) / FU_ROWS;
    return (int)(tiles < FU_MAX_BLOCKS ? (tiles < 1 ? 1 : tiles) : FU_MAX_BLOCKS);
}

}  // namespace

template <typename T>
bool fuse_reductions_supported(int64_t k, int64_t gram_cols) {
    return k >= 1 && k <= FU_MAX_RED && gram_cols >= 0 && gram_cols <= FU_MAX_RED;
}

template <typename T>
size_t fuse_workspace_bytes(int64_t n, int64_t k, int64_t gram_cols, int want_sqnorm) {
    if (gram_cols <= 0 && !want_sqnorm) return 0;
    return (size_t)fuse_blocks(n) * (size_t)(gram_cols * k + k) * sizeof(T);
}

template <typename T>
cudaError_t launch_fuse(const FuseArgs<T>& a, const T* part, int splits, int64_t n, int64_t k, T alpha, void* workspace,
                        cudaStream_t stream) {
    const bool want_gram = a.gram_lhs != nullptr && a.gram_cols > 0;
    const bool reduce = want_gram || a.want_sqnorm;
    if (reduce && !fuse_reductions_supported<T>(k, want_gram ? a.gram_cols : 0)) return cudaErrorInvalidValue;
    const int blocks = fuse_blocks(n);
    const int kg = want_gram ? a.gram_cols : 0;
    const size_t smem = reduce ? (size_t)FU_ROWS * (size_t)(k + kg) * sizeof(T) : 0;
    auto kern = kmm_fuse_kernel<T>;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    T* red_part = static_cast<T*>(workspace);
    kern<<<blocks, FU_THREADS, smem, stream>>>(a, part, splits, n, (int)k, alpha, red_part);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess || !reduce) return err;
    const int per = kg * (int)k + (int)k;
    kmm_fuse_reduce_kernel<T><<<(per + 255) / 256, 256, 0, stream>>>(red_part, blocks, kg, (int)k, a.gram_out,
                                                                     a.want_sqnorm ? a.sqnorm_out : nullptr);
    return cudaGetLastError();
}

template bool fuse_reductions_supported<float>(int64_t, int64_t);
template bool fuse_reductions_supported<double>(int64_t, int64_t);
template size_t fuse_workspace_bytes<float>(int64_t, int64_t, int64_t, int);
template size_t fuse_workspace_bytes<double>(int64_t, int64_t, int64_t, int);
template cudaError_t launch_fuse<float>(const FuseArgs<float>&, const float*, int, int64_t, int64_t, float, void*,
                                        cudaStream_t);
template cudaError_t launch_fuse<double>(const FuseArgs<double>&, const double*, int, int64_t, int64_t, double, void*,
                                         cudaStream_t);

}  // namespace kmm

// Shared definitions for the fused kernel-matrix matmat kernels (sm_100a).
//
//   Y = c * K(R, C) @ V,   K_ij = f(dist(r_i, c_j))
//
// Kernel formulas follow the reference's symbolic definitions
// (rlaopt/kernels/standard.py:31-85); the pointwise function is applied in
// registers, K is never written to HBM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kmm {

enum KernelId : int {
    KID_RBF = 0,       // exp(-|u|_2^2 / 2)                      standard.py:46-52
    KID_LAPLACE = 1,   // exp(-|u|_1)                            standard.py:55-61
    KID_MATERN12 = 2,  // exp(-r)                                standard.py:64-69
    KID_MATERN32 = 3,  // (1 + sqrt3 r) exp(-sqrt3 r)            standard.py:72-77
    KID_MATERN52 = 4,  // (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)  standard.py:80-85
    KID_COUNT = 5
};

// Packed operand tiling: rows padded to PACK_ROWS, features padded to PACK_FEATS.
constexpr int PACK_ROWS = 128;
constexpr int PACK_FEATS = 8;

__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <typename T>
__device__ __forceinline__ T pw_exp(T x);
// fp32: exp(x) = ex2(x log2 e) with the hardware approximation (2^-22 relative error, one FMUL + one MUFU instead of
// the ~8-instruction expf).  The product x log2 e is rounded once: |x| 2^-24 of relative error in the result, < 1e-6
// for every argument whose exponential is not negligible in a sum of kernel values.
template <>
__device__ __forceinline__ float pw_exp<float>(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.44269504088896340736f));
    return y;
}
template <>
__device__ __forceinline__ double pw_exp<double>(double x) { return exp(x); }

template <typename T>
__device__ __forceinline__ T pw_sqrt(T x);
template <>
__device__ __forceinline__ float pw_sqrt<float>(float x) { return sqrtf(x); }
template <>
__device__ __forceinline__ double pw_sqrt<double>(double x) { return sqrt(x); }

// Pointwise kernel function of the accumulated distance D
// (D = sum u^2 for the L2 kernels, sum |u| for Laplace).
template <int KID, typename T>
__device__ __forceinline__ T pointwise_k(T D) {
    if constexpr (KID == KID_RBF) {
        return pw_exp<T>(T(-0.5) * D);
    } else if constexpr (KID == KID_LAPLACE) {
        return pw_exp<T>(-D);
    } else if constexpr (KID == KID_MATERN12) {
        return pw_exp<T>(-pw_sqrt<T>(D));
    } else if constexpr (KID == KID_MATERN32) {
        const T s = T(1.7320508075688772935) * pw_sqrt<T>(D);
        return (T(1) + s) * pw_exp<T>(-s);
    } else {  // KID_MATERN52
        const T s = T(2.2360679774997896964) * pw_sqrt<T>(D);
        return (T(1) + s + T(5.0 / 3.0) * D) * pw_exp<T>(-s);
    }
}

// Apply f in place to a register tile; the switch on the (warp-uniform) kernel id
// is taken once per tile, not once per entry.
template <typename T, int R, int C>
__device__ __forceinline__ void pointwise_tile(int kid, T (&S)[R][C]) {
#define KMM_PW_CASE(KID)                                   \
    case KID:                                              \
        _Pragma("unroll") for (int r = 0; r < R; ++r)      \
        _Pragma("unroll") for (int c = 0; c < C; ++c)      \
            S[r][c] = pointwise_k<KID, T>(S[r][c]);        \
        break;
    switch (kid) {
        KMM_PW_CASE(KID_RBF)
        KMM_PW_CASE(KID_LAPLACE)
        KMM_PW_CASE(KID_MATERN12)
        KMM_PW_CASE(KID_MATERN32)
        default:
            KMM_PW_CASE(KID_MATERN52)
    }
#undef KMM_PW_CASE
}

// cp.async helpers (LDGSTS): 16-byte global->shared copies, src_bytes < 16 zero-fills.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes = 16) {
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

}  // namespace kmm

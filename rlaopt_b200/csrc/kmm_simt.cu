// CUDA-core fused kernel-matrix matmat (general path: all five kernels, fp32/fp64,
// arbitrary n, m, d, k).  Direct-difference distance, the numerical twin of the
// KeOps reduction the reference calls at rlaopt/kernels/base.py:43-47.
//
// One CTA owns a 128-row tile of the output and walks 64-column tiles of K:
//   phase 1  S[128x64]  = sum_d |r_i - c_j|^p      (FADD+FFMA, 8x4 register tile)
//            P          = f(S)                     (registers -> swizzled smem)
//   phase 2  Yacc[128xKC] += P @ V[64xKC]          (FFMA, RTx4 register tile)
// Operands arrive feature-major ("packed", see kmm_pack.cu) so every tile load is
// a run of 16-byte cp.async copies; V tiles are zero-filled past m / k so padded
// columns contribute exactly 0.  Long column sums are accumulated in two levels
// (registers for 1024 columns, then a thread-private smem slot) to keep fp32
// round-off growth ~sqrt(1024) instead of ~sqrt(m).
#include <stdlib.h>

#include "kmm_common.cuh"
#include "kmm_launch.h"

namespace kmm {
namespace {

constexpr int BM = 128;  // output rows per CTA
constexpr int BN = 64;   // K columns per step
constexpr int DC = PACK_FEATS;  // features staged per cp.async stage
// threads per CTA: every thread owns an 8 x TN block of the 128 x 64 distance tile; TN = 8 for fp32 with up to 32
// columns of V per chunk (128 threads: half the shared-memory loads and loop overhead per FADD), else 4 (256 threads)
template <typename T, int KC>
struct SimtShape {
    static constexpr int TN = (sizeof(T) == 4 && KC <= 32) ? 8 : 4;  // KC = 64 keeps 8 x 4 (64 phase-2 accumulators per thread otherwise)
    static constexpr int NT = BM * BN / (8 * TN);
};
constexpr int FLUSH_TILES = 16;

template <typename T, int KC>
struct alignas(16) SimtSmem {
    T As[2][DC][BM];
    T Bs[2][DC][BN];
    T Ps[BN][BM];  // P^T, 16-byte chunks XOR-swizzled by (j >> 2) & 7
    T Vs[BN][KC];
    T Yt[BM * KC / SimtShape<T, KC>::NT][SimtShape<T, KC>::NT];  // second-level accumulators, one private column per thread
};

template <typename T, int N>
__device__ __forceinline__ void lds_vec(T* dst, const T* src) {
    constexpr int BYTES = N * sizeof(T);
    if constexpr (BYTES % 16 == 0) {
#pragma unroll
        for (int q = 0; q < BYTES / 16; ++q)
            reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(src)[q];
    } else if constexpr (BYTES % 8 == 0) {
#pragma unroll
        for (int q = 0; q < BYTES / 8; ++q)
            reinterpret_cast<uint2*>(dst)[q] = reinterpret_cast<const uint2*>(src)[q];
    } else {
#pragma unroll
        for (int q = 0; q < N; ++q) dst[q] = src[q];
    }
}

// Packed fp32 pairs (Blackwell FADD2 / FFMA2): one issue slot for two lanes of the FP32 pipe.  The scalar operand of
// a pair built as {a, a} is encoded as a broadcast source (SASS `R.F32`), so no register copies are spent on it.
#ifndef KMM_SIMT_PACKED
#define KMM_SIMT_PACKED 1
#endif
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <typename T>
__device__ __forceinline__ T absval(T x);
template <>
__device__ __forceinline__ float absval<float>(float x) { return fabsf(x); }
template <>
__device__ __forceinline__ double absval<double>(double x) { return fabs(x); }

template <typename T>
__device__ __forceinline__ T fmadd(T a, T b, T c);
template <>
__device__ __forceinline__ float fmadd<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <>
__device__ __forceinline__ double fmadd<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T, bool L1, int KC>
__global__ void __launch_bounds__(SimtShape<T, KC>::NT, sizeof(T) == 4 ? (KC <= 32 ? 3 : 2) : 1)
kmm_simt_kernel(const T* __restrict__ Rt, int64_t n, int64_t n_pad,
                const T* __restrict__ Ct, int64_t m, int64_t m_pad, int d_pad,
                const T* __restrict__ V, int64_t ldv, int k, int v_vec_ok,
                T* __restrict__ out, int64_t ldo, int64_t split_stride,
                T scale, int kid, int tiles_per_split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SimtSmem<T, KC>& sm = *reinterpret_cast<SimtSmem<T, KC>*>(smem_raw);

    constexpr int VEC = 16 / sizeof(T);
    constexpr int TN = SimtShape<T, KC>::TN, NT = SimtShape<T, KC>::NT;
    constexpr int CT = 4;                 // phase-2 columns per thread
    constexpr int NACC = BM * KC / NT;    // phase-2 accumulators per thread
    constexpr int RT = NACC / CT;         // phase-2 rows per thread
    static_assert(RT >= 1 && RT * CT == NACC, "phase-2 tiling must cover the output tile");
    constexpr bool PACKED = KMM_SIMT_PACKED && sizeof(T) == 4;

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int kc0 = blockIdx.y * KC;
    const int64_t n_col_tiles = (m + BN - 1) / BN;
    const int64_t t_begin = (int64_t)blockIdx.z * tiles_per_split;
    const int64_t t_end = min(n_col_tiles, t_begin + (int64_t)tiles_per_split);
    const int nd = d_pad / DC;

    // phase-2 thread tile
    const int cg = tid % (KC / CT), rg = tid / (KC / CT);
    const int r0 = rg * RT, c0 = cg * CT;

    auto load_ab = [&](int64_t t, int c, int buf) {
        constexpr int A_CH = DC * BM / VEC;
#pragma unroll
        for (int q = tid; q < A_CH; q += NT) {
            const int dd = q / (BM / VEC), i = (q % (BM / VEC)) * VEC;
            cp_async16(&sm.As[buf][dd][i], Rt + (int64_t)(c * DC + dd) * n_pad + row0 + i);
        }
        constexpr int B_CH = DC * BN / VEC;
        const int64_t col0 = t * BN;
#pragma unroll
        for (int q = tid; q < B_CH; q += NT) {
            const int dd = q / (BN / VEC), j = (q % (BN / VEC)) * VEC;
            cp_async16(&sm.Bs[buf][dd][j], Ct + (int64_t)(c * DC + dd) * m_pad + col0 + j);
        }
    };

    auto load_v = [&](int64_t t) {
        const int64_t col0 = t * BN;
        if (v_vec_ok) {
            constexpr int CH_PER_ROW = KC / VEC;
#pragma unroll
            for (int q = tid; q < BN * CH_PER_ROW; q += NT) {
                const int j = q / CH_PER_ROW, c = (q % CH_PER_ROW) * VEC;
                const int64_t gj = col0 + j;
                const int gc = kc0 + c;
                int valid = 0;
                if (gj < m && gc < k) valid = min(VEC, k - gc) * (int)sizeof(T);
                const T* src = valid ? V + gj * ldv + gc : V;
                cp_async16(&sm.Vs[j][c], src, valid);
            }
        } else {
            for (int q = tid; q < BN * KC; q += NT) {
                const int j = q / KC, c = q % KC;
                const int64_t gj = col0 + j;
                const int gc = kc0 + c;
                sm.Vs[j][c] = (gj < m && gc < k) ? V[gj * ldv + gc] : T(0);
            }
        }
    };

    T acc[RT][CT];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] = T(0);
#pragma unroll
    for (int e = 0; e < NACC; ++e) sm.Yt[e][tid] = T(0);

    if (t_begin < t_end) {
        load_ab(t_begin, 0, 0);
        cp_async_commit();
    }

    int since_flush = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
        // ---------------- phase 1: distances ----------------
        T S[8][TN];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) S[r][c] = T(0);

        for (int c = 0; c < nd; ++c) {
            cp_async_wait_all();
            __syncthreads();  // chunk c landed; everyone is done with the other buffer, Vs and Ps
            if (c == 0) load_v(t);
            if (c + 1 < nd) load_ab(t, c + 1, (c + 1) & 1);
            cp_async_commit();
            const int buf = c & 1;
            // the 8 x 8 tile fully unrolled over a chunk is 16 KB of FADDs (ncu: 8 % instruction-fetch stalls): unroll by 2
            constexpr int UNR = TN == 8 ? 2 : DC;
#pragma unroll(UNR)
            for (int dd = 0; dd < DC; ++dd) {
                alignas(16) T a[8];
                alignas(16) T b[TN];
                lds_vec<T, 8>(a, &sm.As[buf][dd][ty * 8]);
                lds_vec<T, TN>(b, &sm.Bs[buf][dd][tx * TN]);
                if constexpr (PACKED) {
                    // a_r - {b_c, b_c+1} as one FADD2; |.| is a source modifier of the scalar FADD that accumulates
                    uint64_t bp[TN / 2];
#pragma unroll
                    for (int c2 = 0; c2 < TN / 2; ++c2) bp[c2] = pk2(b[2 * c2], b[2 * c2 + 1]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const uint64_t ar = pk2(a[r], a[r]);
#pragma unroll
                        for (int c2 = 0; c2 < TN / 2; ++c2) {
                            const uint64_t diff = sub2(ar, bp[c2]);
                            if constexpr (L1) {
                                float lo, hi;
                                up2(diff, lo, hi);
                                S[r][2 * c2] += fabsf(lo);
                                S[r][2 * c2 + 1] += fabsf(hi);
                            } else {
                                up2(fma2(diff, diff, pk2(S[r][2 * c2], S[r][2 * c2 + 1])), S[r][2 * c2], S[r][2 * c2 + 1]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int cc = 0; cc < TN; ++cc) {
                            const T diff = a[r] - b[cc];
                            if constexpr (L1) S[r][cc] += absval<T>(diff);
                            else S[r][cc] = fmadd<T>(diff, diff, S[r][cc]);
                        }
                }
            }
        }

        // ---------------- pointwise: P = f(S) -> swizzled smem ----------------
        pointwise_tile<T, 8, TN>(kid, S);
#pragma unroll
        for (int cc = 0; cc < TN; ++cc) {
            const int j = tx * TN + cc;
            const int swz = (j >> 2) & 7;
            alignas(16) T p[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) p[r] = S[r][cc];
            constexpr int NCH = 8 / VEC;
#pragma unroll
            for (int h = 0; h < NCH; ++h) {
                const int chunk = (ty * NCH + h) ^ swz;
                *reinterpret_cast<uint4*>(&sm.Ps[j][chunk * VEC]) = *reinterpret_cast<const uint4*>(&p[h * VEC]);
            }
        }
        cp_async_wait_all();
        __syncthreads();  // V landed, P visible, As/Bs free
        if (t + 1 < t_end) {
            load_ab(t + 1, 0, 0);
            cp_async_commit();
        }

        // ---------------- phase 2: Yacc += P @ V ----------------
#pragma unroll 32
        for (int j = 0; j < BN; ++j) {  // unrolled by the swizzle period: (j >> 2) & 7 is a compile-time constant
            const int swz = (j >> 2) & 7;
            alignas(16) T p[RT];
            alignas(16) T v[CT];
            if constexpr (RT >= VEC) {
#pragma unroll
                for (int h = 0; h < RT / VEC; ++h) {
                    const int chunk = (r0 / VEC + h) ^ swz;
                    lds_vec<T, VEC>(&p[h * VEC], &sm.Ps[j][chunk * VEC]);
                }
            } else {
                const int chunk = (r0 / VEC) ^ swz;
                lds_vec<T, RT>(p, &sm.Ps[j][chunk * VEC + (r0 % VEC)]);
            }
            lds_vec<T, CT>(v, &sm.Vs[j][c0]);
            if constexpr (PACKED) {
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const uint64_t pr = pk2(p[r], p[r]);
#pragma unroll
                    for (int c2 = 0; c2 < CT / 2; ++c2)
                        up2(fma2(pr, pk2(v[2 * c2], v[2 * c2 + 1]), pk2(acc[r][2 * c2], acc[r][2 * c2 + 1])),
                            acc[r][2 * c2], acc[r][2 * c2 + 1]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < CT; ++c) acc[r][c] = fmadd<T>(p[r], v[c], acc[r][c]);
            }
        }

        if (++since_flush == FLUSH_TILES) {
            since_flush = 0;
#pragma unroll
            for (int r = 0; r < RT; ++r)
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    sm.Yt[r * CT + c][tid] += acc[r][c];
                    acc[r][c] = T(0);
                }
        }
    }

    T* dst = out + (int64_t)blockIdx.z * split_stride;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const int64_t i = row0 + r0 + r;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int col = kc0 + c0 + c;
            if (i < n && col < k) dst[i * ldo + col] = (sm.Yt[r * CT + c][tid] + acc[r][c]) * scale;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fp32, narrow k-chunks (instantiated for KC = 8; KC = 16 works but measured slower, see launch_kc): P stays in registers.
//
// Same 128 x 64 tile and the same phase 1 (8 x 8 distance tile per thread, 128 threads), but the thread that owns
// P[8 rows][8 columns j] also contracts it: acc[8 rows][KC] += P[r][j] V[j][:] over ITS eight j.  The partial sums of
// the eight threads that share a row block (the lanes tx = 0..7 of a row group) are added once, at the end of the
// kernel, with three shuffles -- so the per-tile round trip of P through shared memory (64 stores + 64 loads per thread
// and one barrier) is gone; what phase 2 reads from shared memory is V only (KC / 4 broadcast loads per 64 FFMA2).
// V is staged as Vs[c / 4][perm(j)][4] with perm(j) = (j % 8) * 8 + j / 8: the eight lanes of a quarter-warp then read
// eight consecutive 16-byte rows (conflict-free) although their columns j = 8 tx + cc are 8 apart.
template <int KC>
struct alignas(16) SimtRegpSmem {
    float As[2][DC][BM];
    float Bs[2][DC][BN];
    float Vs[KC / 4][BN][4];
    float Yt[8 * KC][BM];  // second-level accumulators, one private column per thread (128 threads)
};

template <bool L1, int KC>
__global__ void __launch_bounds__(128, 2)
kmm_simt_regp_kernel(const float* __restrict__ Rt, int64_t n, int64_t n_pad, const float* __restrict__ Ct, int64_t m,
                     int64_t m_pad, int d_pad, const float* __restrict__ V, int64_t ldv, int k, int v_vec_ok,
                     float* __restrict__ out, int64_t ldo, int64_t split_stride, float scale, int kid,
                     int tiles_per_split) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SimtRegpSmem<KC>& sm = *reinterpret_cast<SimtRegpSmem<KC>*>(smem_raw);
    constexpr int NT = 128, TN = 8, VEC = 4, NACC = 8 * KC;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);  // 8 x 16 thread grid over the 64 x 128 tile
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int kc0 = blockIdx.y * KC;
    const int64_t n_col_tiles = (m + BN - 1) / BN;
    const int64_t t_begin = (int64_t)blockIdx.z * tiles_per_split;
    const int64_t t_end = min(n_col_tiles, t_begin + (int64_t)tiles_per_split);
    const int nd = d_pad / DC;

    auto load_ab = [&](int64_t t, int c, int buf) {
        constexpr int A_CH = DC * BM / VEC;
#pragma unroll
        for (int q = tid; q < A_CH; q += NT) {
            const int dd = q / (BM / VEC), i = (q % (BM / VEC)) * VEC;
            cp_async16(&sm.As[buf][dd][i], Rt + (int64_t)(c * DC + dd) * n_pad + row0 + i);
        }
        constexpr int B_CH = DC * BN / VEC;
        const int64_t col0 = t * BN;
#pragma unroll
        for (int q = tid; q < B_CH; q += NT) {
            const int dd = q / (BN / VEC), j = (q % (BN / VEC)) * VEC;
            cp_async16(&sm.Bs[buf][dd][j], Ct + (int64_t)(c * DC + dd) * m_pad + col0 + j);
        }
    };
    auto load_v = [&](int64_t t) {
        const int64_t col0 = t * BN;
        constexpr int CH_PER_ROW = KC / VEC;
        if (v_vec_ok) {
#pragma unroll
            for (int q = tid; q < BN * CH_PER_ROW; q += NT) {
                const int j = q / CH_PER_ROW, c4 = q % CH_PER_ROW;
                const int64_t gj = col0 + j;
                const int gc = kc0 + c4 * VEC;
                int valid = 0;
                if (gj < m && gc < k) valid = min(VEC, k - gc) * (int)sizeof(float);
                const float* src = valid ? V + gj * ldv + gc : V;
                cp_async16(&sm.Vs[c4][(j % 8) * 8 + j / 8][0], src, valid);
            }
        } else {
            for (int q = tid; q < BN * KC; q += NT) {
                const int j = q / KC, c = q % KC;
                const int64_t gj = col0 + j;
                const int gc = kc0 + c;
                sm.Vs[c / 4][(j % 8) * 8 + j / 8][c % 4] = (gj < m && gc < k) ? V[gj * ldv + gc] : 0.0f;
            }
        }
    };

    float acc[8][KC];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < KC; ++c) acc[r][c] = 0.0f;
#pragma unroll
    for (int e = 0; e < NACC; ++e) sm.Yt[e][tid] = 0.0f;

    if (t_begin < t_end) {
        load_ab(t_begin, 0, 0);
        cp_async_commit();
    }
    int since_flush = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
        // ---------------- phase 1: distances (as kmm_simt_kernel) ----------------
        float S[8][TN];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) S[r][c] = 0.0f;
        for (int c = 0; c < nd; ++c) {
            cp_async_wait_all();
            __syncthreads();  // chunk c landed; everyone is done with the other buffer and (c == 0) with Vs
            if (c == 0) load_v(t);
            if (c + 1 < nd) load_ab(t, c + 1, (c + 1) & 1);
            cp_async_commit();
            const int buf = c & 1;
#pragma unroll 1
            for (int dd = 0; dd < DC; ++dd) {
                alignas(16) float a[8];
                alignas(16) float b[TN];
                lds_vec<float, 8>(a, &sm.As[buf][dd][ty * 8]);
                lds_vec<float, TN>(b, &sm.Bs[buf][dd][tx * TN]);
                uint64_t bp[TN / 2];
#pragma unroll
                for (int c2 = 0; c2 < TN / 2; ++c2) bp[c2] = pk2(b[2 * c2], b[2 * c2 + 1]);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint64_t ar = pk2(a[r], a[r]);
#pragma unroll
                    for (int c2 = 0; c2 < TN / 2; ++c2) {
                        const uint64_t diff = sub2(ar, bp[c2]);
                        if constexpr (L1) {
                            float lo, hi;
                            up2(diff, lo, hi);
                            S[r][2 * c2] += fabsf(lo);
                            S[r][2 * c2 + 1] += fabsf(hi);
                        } else {
                            up2(fma2(diff, diff, pk2(S[r][2 * c2], S[r][2 * c2 + 1])), S[r][2 * c2], S[r][2 * c2 + 1]);
                        }
                    }
                }
            }
        }
        // ---------------- pointwise in place: P = f(S) ----------------
        pointwise_tile<float, 8, TN>(kid, S);
        cp_async_wait_all();
        __syncthreads();  // V landed; As / Bs are free
        if (t + 1 < t_end) {
            load_ab(t + 1, 0, 0);
            cp_async_commit();
        }
        // ---------------- phase 2: acc[r][:] += P[r][cc] V[8 tx + cc][:], P from registers ----------------
#pragma unroll
        for (int cc = 0; cc < TN; ++cc) {
#pragma unroll
            for (int c4 = 0; c4 < KC / 4; ++c4) {
                const float4 v = *reinterpret_cast<const float4*>(&sm.Vs[c4][cc * 8 + tx][0]);
                const uint64_t v01 = pk2(v.x, v.y), v23 = pk2(v.z, v.w);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint64_t pr = pk2(S[r][cc], S[r][cc]);
                    up2(fma2(pr, v01, pk2(acc[r][4 * c4], acc[r][4 * c4 + 1])), acc[r][4 * c4], acc[r][4 * c4 + 1]);
                    up2(fma2(pr, v23, pk2(acc[r][4 * c4 + 2], acc[r][4 * c4 + 3])), acc[r][4 * c4 + 2], acc[r][4 * c4 + 3]);
                }
            }
        }
        if (++since_flush == FLUSH_TILES) {
            since_flush = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    sm.Yt[r * KC + c][tid] += acc[r][c];
                    acc[r][c] = 0.0f;
                }
        }
    }
    // ---------------- row sums over the eight lanes that share a row group, then one writer per group ----------------
    float* dst = out + (int64_t)blockIdx.z * split_stride;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t i = row0 + ty * 8 + r;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            float y = sm.Yt[r * KC + c][tid] + acc[r][c];
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            y += __shfl_xor_sync(0xffffffffu, y, 4);
            const int col = kc0 + c;
            if (tx == 0 && i < n && col < k) dst[i * ldo + col] = y * scale;
        }
    }
}

template <typename T>
__global__ void kmm_split_reduce_kernel(const T* __restrict__ part, int splits, int64_t nk, int64_t k,
                                        T* __restrict__ Y, int64_t ldy, T scale) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nk) return;
    T s = T(0);
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * nk + e];
    Y[(e / k) * ldy + (e % k)] = s * scale;
}

template <typename T, bool L1, int KC>
cudaError_t launch_one(const SimtArgs<T>& a, int splits, int tiles_per_split, T* out, int64_t ldo,
                       int64_t split_stride, T scale) {
    using Smem = SimtSmem<T, KC>;
    auto kern = kmm_simt_kernel<T, L1, KC>;
    // the attribute lives in the device context (one per GPU of a multi-device process): set it on every launch
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (err != cudaSuccess) return err;
    const int64_t row_tiles = (a.n + BM - 1) / BM;
    const int k_chunks = (int)((a.k + KC - 1) / KC);
    dim3 grid((unsigned)row_tiles, (unsigned)k_chunks, (unsigned)splits);
    const int v_vec_ok = ((reinterpret_cast<uintptr_t>(a.V) % 16) == 0) && ((a.ldv * sizeof(T)) % 16 == 0);
    kern<<<grid, SimtShape<T, KC>::NT, sizeof(Smem), a.stream>>>(a.Rt, a.n, a.n_pad, a.Ct, a.m, a.m_pad, (int)a.d_pad, a.V, a.ldv,
                                               (int)a.k, v_vec_ok, out, ldo, split_stride, scale, a.kid,
                                               tiles_per_split);
    return cudaGetLastError();
}

// fp32 with at most 8 columns per chunk: the register-resident-P kernel (RLAOPT_B200_SIMT_REGP=0 keeps the
// shared-memory P kernel, for A/B runs)
template <bool L1, int KC>
cudaError_t launch_regp(const SimtArgs<float>& a, int splits, int tiles_per_split, float* out, int64_t ldo,
                        int64_t split_stride, float scale) {
    using Smem = SimtRegpSmem<KC>;
    auto kern = kmm_simt_regp_kernel<L1, KC>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (err != cudaSuccess) return err;
    const int64_t row_tiles = (a.n + BM - 1) / BM;
    const int k_chunks = (int)((a.k + KC - 1) / KC);
    dim3 grid((unsigned)row_tiles, (unsigned)k_chunks, (unsigned)splits);
    const int v_vec_ok = ((reinterpret_cast<uintptr_t>(a.V) % 16) == 0) && ((a.ldv * sizeof(float)) % 16 == 0);
    kern<<<grid, 128, sizeof(Smem), a.stream>>>(a.Rt, a.n, a.n_pad, a.Ct, a.m, a.m_pad, (int)a.d_pad, a.V, a.ldv, (int)a.k,
                                                v_vec_ok, out, ldo, split_stride, scale, a.kid, tiles_per_split);
    return cudaGetLastError();
}

bool simt_regp_enabled() {
    static const int on = [] {
        const char* v = getenv("RLAOPT_B200_SIMT_REGP");
        return (v && *v) ? atoi(v) : 1;
    }();
    return on != 0;
}

template <typename T, bool L1>
cudaError_t launch_kc(const SimtArgs<T>& a, int kc, int splits, int tiles_per_split, T* out, int64_t ldo,
                      int64_t split_stride, T scale) {
    if constexpr (sizeof(T) == 4) {
        // measured (profiles/r02_simt_regp_ab.log): +3 % for chunks of 8 columns (Laplace d=32 k=8 359 -> 372, d=8 k=10
        // 633 -> 650 Gentries/s); with 16 columns the 128 accumulators cost a third of the resident warps and the
        // shared-memory-P kernel stays ahead (332 vs 320), so only KC = 8 runs here
        if (simt_regp_enabled() && kc == 8)
            return launch_regp<L1, 8>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
    }
    switch (kc) {
        case 8: return launch_one<T, L1, 8>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
        case 16: return launch_one<T, L1, 16>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
        case 32: return launch_one<T, L1, 32>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
        default:
            if constexpr (sizeof(T) == 4)
                return launch_one<T, L1, 64>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
            else
                return launch_one<T, L1, 32>(a, splits, tiles_per_split, out, ldo, split_stride, scale);
    }
}

}  // namespace

template <typename T>
cudaError_t launch_split_reduce(const T* part, int splits, int64_t n, int64_t k, T* Y, int64_t ldy, T scale,
                                cudaStream_t stream) {
    const int64_t nk = n * k;
    const int threads = 256;
    kmm_split_reduce_kernel<T><<<(unsigned)((nk + threads - 1) / threads), threads, 0, stream>>>(part, splits, nk, k, Y,
                                                                                                ldy, scale);
    return cudaGetLastError();
}
template cudaError_t launch_split_reduce<float>(const float*, int, int64_t, int64_t, float*, int64_t, float, cudaStream_t);
template cudaError_t launch_split_reduce<double>(const double*, int, int64_t, int64_t, double*, int64_t, double,
                                                 cudaStream_t);

template <typename T>
int simt_pick_kc(int64_t k) {
    const int kc_max = sizeof(T) == 4 ? 64 : 32;
    int kc = 8;
    while (kc < kc_max && kc < k) kc *= 2;
    return kc;
}

template <typename T>
void simt_plan(int64_t n, int64_t m, int64_t k, int sm_count, int* kc_out, int* splits_out, int* tiles_per_split_out) {
    const int kc = simt_pick_kc<T>(k);
    const int64_t row_tiles = (n + BM - 1) / BM;
    const int64_t k_chunks = (k + kc - 1) / kc;
    const int64_t col_tiles = (m + BN - 1) / BN;
    const int64_t base = row_tiles * k_chunks;
    const int64_t target = (int64_t)sm_count * (sizeof(T) == 4 ? 2 : 1) * 2;  // two waves of resident CTAs
    int64_t splits = 1;
    if (base < target) {
        splits = (target + base - 1) / base;
        const int64_t max_splits = (col_tiles + 7) / 8;  // keep >= 8 column tiles (512 columns) per split
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        if (splits > 1024) splits = 1024;
    }
    int64_t tps = (col_tiles + splits - 1) / splits;
    if (tps < 1) tps = 1;
    splits = (col_tiles + tps - 1) / tps;
    if (splits < 1) splits = 1;
    *kc_out = kc;
    *splits_out = (int)splits;
    *tiles_per_split_out = (int)tps;
}

template <typename T>
size_t simt_workspace_bytes(int64_t n, int64_t m, int64_t k, int sm_count, bool keep_partials) {
    int kc, splits, tps;
    simt_plan<T>(n, m, k, sm_count, &kc, &splits, &tps);
    return (splits > 1 || keep_partials) ? (size_t)splits * (size_t)n * (size_t)k * sizeof(T) : 0;
}

template <typename T>
cudaError_t launch_simt(const SimtArgs<T>& a, int sm_count, void* workspace, size_t workspace_bytes, bool keep_partials,
                        int* splits_out) {
    int kc, splits, tps;
    simt_plan<T>(a.n, a.m, a.k, sm_count, &kc, &splits, &tps);
    const bool l1 = a.kid == KID_LAPLACE;
    if (splits_out) *splits_out = splits;
    if (splits > 1 || keep_partials) {
        const size_t need = (size_t)splits * (size_t)a.n * (size_t)a.k * sizeof(T);
        if (workspace == nullptr || workspace_bytes < need) return cudaErrorInvalidValue;
        T* part = static_cast<T*>(workspace);
        cudaError_t err = l1 ? launch_kc<T, true>(a, kc, splits, tps, part, a.k, a.n * a.k, T(1))
                             : launch_kc<T, false>(a, kc, splits, tps, part, a.k, a.n * a.k, T(1));
        if (err != cudaSuccess || keep_partials) return err;
        const int64_t nk = a.n * a.k;
        const int threads = 256;
        kmm_split_reduce_kernel<T><<<(unsigned)((nk + threads - 1) / threads), threads, 0, a.stream>>>(
            part, splits, nk, a.k, a.Y, a.ldy, a.scale);
        return cudaGetLastError();
    }
    return l1 ? launch_kc<T, true>(a, kc, 1, tps, a.Y, a.ldy, 0, a.scale)
              : launch_kc<T, false>(a, kc, 1, tps, a.Y, a.ldy, 0, a.scale);
}

template cudaError_t launch_simt<float>(const SimtArgs<float>&, int, void*, size_t, bool, int*);
template cudaError_t launch_simt<double>(const SimtArgs<double>&, int, void*, size_t, bool, int*);
template size_t simt_workspace_bytes<float>(int64_t, int64_t, int64_t, int, bool);
template size_t simt_workspace_bytes<double>(int64_t, int64_t, int64_t, int, bool);

}  // namespace kmm

// tcgen05 / TMEM fused kernel-matrix matmat for the L2-distance kernels (fp32 in/out).
//
//   Y[i,:] = c * sum_j f(|x_i|^2 + |y_j|^2 - 2 x_i.y_j) V[j,:]
//
// flash-attention-shaped, one CTA per 128 output rows, streaming 64-column sub-tiles:
//
//   producer warp   cp.async.bulk (TMA engine, UBLKCP) of pre-swizzled tile images
//                   {Y-tile fp16 hi/lo, V-tile tf32 hi/lo, |y|^2} into a smem ring
//   MMA warp        MMA1  S  = X.Y^T     kind::f16, A = X tile resident in TMEM,
//                                        3 products hi.lo + lo.hi + hi.hi, fp32 accumulate
//                   MMA2  O += P.V       kind::tf32, A = P from TMEM (in place over S),
//                                        3 products hi.lo + lo.hi + hi.hi
//   8 epilogue warps  tcgen05.ld S -> D = |x|^2+|y|^2-2S -> P = f(D) in registers
//                   -> split P into tf32 hi/lo -> tcgen05.st back into TMEM;
//                   -> drain the previous sub-tile's O from TMEM into fp32 registers: each
//                   sub-tile gets a fresh TMEM accumulator and the long column sum is done
//                   with round-to-nearest FADDs (the tensor core accumulates with
//                   truncation, measured bias -1e-5 over 1024 columns)
//
// Split-precision arithmetic (why fp32 parity holds, DESIGN.md "numerics"):
//   x*s = hi + lo (+2^-22), fp16 pair, s a power of two chosen per operand so |x*s| < 2^13
//   P   = hi + lo (+2^-21), tf32 pair;  V likewise
// K is never written to HBM; S and P never leave TMEM / registers.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kmm_common.cuh"
#include "kmm_launch.h"
#include "kmm_tmem_ldst.cuh"

namespace kmm {
namespace {

constexpr int TC_BM = 128;      // rows per CTA (TMEM lanes)
constexpr int TC_BN = 64;       // K columns per sub-tile
constexpr int TC_EPI_WARPS = 8; // warps 0..7: pointwise; warp 8: producer; warp 9: MMA issue
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_MIN_SPLIT_TILES = 16;  // a column split covers at least 16 sub-tiles (1024 columns)
constexpr int TC_KBLOCK_BYTES = 64 * 128;  // one K-block of a 64-row image: 64 rows x 128 B
constexpr int TC_HEADER_BYTES = 256;
constexpr int TC_MAX_D = 192;
constexpr int TC_SMEM_LIMIT = 220 * 1024;

struct TcHeader {
    unsigned int absmax_bits;  // max |x / lengthscale| as float bits (filled by the absmax kernel)
    float scale;               // power of two s with |x * s| in [2^12, 2^13)
    float inv_scale;
};

__host__ __device__ inline int tc_kblocks(int64_t d) { return (int)((d + 63) / 64); }
__host__ __device__ inline int64_t tc_npad(int64_t n) { return round_up(n, TC_BM); }
__host__ __device__ inline size_t tc_image_bytes(int kb) { return (size_t)2 * kb * TC_KBLOCK_BYTES; }
__host__ __device__ inline size_t tc_norm_offset() { return TC_HEADER_BYTES; }
__host__ __device__ inline size_t tc_image_offset(int64_t n) { return TC_HEADER_BYTES + (size_t)tc_npad(n) * sizeof(float); }

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: try_wait suspends the thread in hardware for a while before it returns false,
// so the counter only trips on a genuine protocol bug (trap instead of hanging the GPU).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 UMMA):
// rows at 128 B pitch, 8-row groups 1024 B apart (SBO), version 1, layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor: fp32 accumulate, K-major A and B, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int ab_format, int n) {
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(TC_BM >> 4) << 24);
}
constexpr int FMT_F16 = 0, FMT_TF32 = 2;

// round-to-nearest split of an fp32 value into a tf32-representable head and its exact remainder
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    const uint32_t h = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
    hi = h;
    lo = __float_as_uint(v - __uint_as_float(h));
}

// pointwise kernel function on the squared distance D >= 0 (fast-math forms; rel. error ~2e-7)
template <int KID>
__device__ __forceinline__ float tc_pointwise(float D) {
    if constexpr (KID == KID_RBF) {
        return ex2_approx(D * -0.72134752044448170368f);  // exp(-D/2)
    } else if constexpr (KID == KID_MATERN12) {
        return ex2_approx(sqrt_approx(D) * -1.44269504088896340736f);
    } else if constexpr (KID == KID_MATERN32) {
        const float s = 1.7320508075688772935f * sqrt_approx(D);
        return (1.0f + s) * ex2_approx(s * -1.44269504088896340736f);
    } else {  // KID_MATERN52
        const float s = 2.2360679774997896964f * sqrt_approx(D);
        return fmaf(D, 1.6666666666666667f, 1.0f + s) * ex2_approx(s * -1.44269504088896340736f);
    }
}

// ------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------
__global__ void tc_absmax_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx,
                                 const int64_t* __restrict__ idx, float inv_ls, const float* __restrict__ inv_ls_vec,
                                 TcHeader* hdr) {
    float mx = 0.0f;
    const int64_t total = n * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d, f = e % d;
        const int64_t src = idx ? idx[i] : i;
        const float s = inv_ls_vec ? inv_ls_vec[f] : inv_ls;
        const float v = fabsf(X[src * ldx + f] * s);
        if (v < 3.0e38f) mx = fmaxf(mx, v);  // ignore inf / nan here; they propagate through the values
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(&hdr->absmax_bits, __float_as_uint(mx));
}

// one thread per (padded) point: fp16 hi/lo split into the swizzled tile image + squared norm
__global__ void tc_pack_points_kernel(const float* __restrict__ X, int64_t n, int64_t d, int64_t ldx,
                                      const int64_t* __restrict__ idx, float inv_ls,
                                      const float* __restrict__ inv_ls_vec, unsigned char* __restrict__ packed, int kb_count) {
    TcHeader* hdr = reinterpret_cast<TcHeader*>(packed);
    const float absmax = __uint_as_float(hdr->absmax_bits);
    float s = 1.0f;
    if (absmax > 0.0f) s = ldexpf(1.0f, 12 - ilogbf(absmax));
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        hdr->scale = s;
        hdr->inv_scale = 1.0f / s;
    }
    const int64_t n_pad = tc_npad(n);
    if (i >= n_pad) return;
    float* norms = reinterpret_cast<float*>(packed + tc_norm_offset());
    unsigned char* img = packed + tc_image_offset(n) + (size_t)(i >> 6) * tc_image_bytes(kb_count);
    const int r = (int)(i & 63);
    const int64_t src = (i < n) ? (idx ? idx[i] : i) : 0;
    double nrm = 0.0;
    for (int kb = 0; kb < kb_count; ++kb) {
        for (int c = 0; c < 8; ++c) {
            alignas(16) __half hi[8];
            alignas(16) __half lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int64_t f = (int64_t)kb * 64 + c * 8 + e;
                float v = 0.0f;
                if (i < n && f < d) v = X[src * ldx + f] * (inv_ls_vec ? inv_ls_vec[f] : inv_ls) * s;
                const __half h = __float2half_rn(v);
                const __half l = __float2half_rn(v - __half2float(h));
                hi[e] = h;
                lo[e] = l;
                const double eff = (double)__half2float(h) + (double)__half2float(l);
                nrm += eff * eff;
            }
            const size_t off = (size_t)kb * TC_KBLOCK_BYTES + (size_t)r * 128 + (size_t)((c ^ (r & 7)) * 16);
            *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(img + (size_t)kb_count * TC_KBLOCK_BYTES + off) = *reinterpret_cast<const uint4*>(lo);
        }
    }
    norms[i] = (float)(nrm / ((double)s * (double)s));
}

// V[m][k] -> per (k-chunk, 64-column sub-tile) image of the MMA2 B operand:
//   [hi | lo] x [K-block of 32 j] x [row c of KP] x 128 B (32 floats, 16 B chunks XOR-swizzled by c & 7)
__global__ void tc_pack_v_kernel(const float* __restrict__ V, int64_t m, int64_t k, int64_t ldv,
                                 unsigned char* __restrict__ images, int kp, int64_t sub_tiles, int k_chunks) {
    const int64_t total = (int64_t)k_chunks * sub_tiles * 2 * kp * 8;  // one thread per 16-byte chunk
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int c = (int)(tid % kp);  // fastest: consecutive threads read consecutive columns of V
    int64_t rest = tid / kp;
    const int chunk = (int)(rest % 8);
    rest /= 8;
    const int kb = (int)(rest % 2);
    rest /= 2;
    const int64_t t = rest % sub_tiles;
    const int kc = (int)(rest / sub_tiles);
    const int64_t col = (int64_t)kc * kp + c;
    alignas(16) uint32_t hi[4];
    alignas(16) uint32_t lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int64_t j = t * TC_BN + kb * 32 + chunk * 4 + e;
        const float v = (j < m && col < k) ? V[j * ldv + col] : 0.0f;
        split_tf32(v, hi[e], lo[e]);
    }
    const size_t image_bytes = (size_t)kp * 512;
    unsigned char* img = images + ((size_t)kc * sub_tiles + t) * image_bytes;
    const size_t off = (size_t)kb * kp * 128 + (size_t)c * 128 + (size_t)((chunk ^ (c & 7)) * 16);
    *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + (size_t)kp * 256 + off) = *reinterpret_cast<const uint4*>(lo);
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
struct TcParams {
    const unsigned char* rows;  // packed row operand
    const unsigned char* cols;  // packed column operand
    const unsigned char* vimg;  // packed V images
    float* out;
    int64_t ldo, split_stride;
    int64_t n, m;
    int k, kb, nk1, stages, kid;
    float scale_out;
    int64_t sub_tiles;        // ceil(m / 64)
    int tiles_per_split;
};

// one elected lane of a converged warp (same lane every time for a full mask)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// tcgen05.mma with the shared-memory descriptor passed as (lo, hi) words: only the low word
// (start address) changes between K steps, so advancing an operand is one 32-bit add.
template <int FMT>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                        uint32_t idesc, uint32_t acc) {
    if constexpr (FMT == FMT_F16) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b64 bd;\n"
            "mov.b64 bd, {%2, %3};\n"
            "setp.ne.b32 p, %5, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b64 bd;\n"
            "mov.b64 bd, {%2, %3};\n"
            "setp.ne.b32 p, %5, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
            : "memory");
    }
}

// TMEM column map (512 columns x 128 lanes, fp32 words):
//   [0, 32 KB)              X tile, fp16 hi halves (2 per column)
//   [32 KB, 64 KB)          X tile, fp16 lo halves
//   [64 KB, +256)           two S/P buffers: S (64 cols, overwritten in place by P_hi) | P_lo (64 cols)
//   [64 KB + 256, +2 KP)    two O buffers (one fresh accumulator per sub-tile, alternating)
template <int KP>
__global__ void __launch_bounds__(TC_THREADS, 1) kmm_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = p.kb, STAGES = p.stages;
    const uint32_t a_img_bytes = (uint32_t)tc_image_bytes(KB);  // hi + lo
    constexpr uint32_t v_img_bytes = KP * 512;
    const uint32_t stage_bytes = a_img_bytes + v_img_bytes;
    float* ny_smem = reinterpret_cast<float*>(smem + (size_t)STAGES * stage_bytes);  // [STAGES][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ny_smem + STAGES * TC_BN);
    uint64_t* full = bars;              // [STAGES] producer -> MMA / epilogue
    uint64_t* empty = full + STAGES;    // [STAGES] MMA -> producer
    uint64_t* s_full = empty + STAGES;  // [2] MMA1 done
    uint64_t* p_full = s_full + 2;      // [2] P written (8 warps)
    uint64_t* p_free = p_full + 2;      // [2] MMA2 done reading P
    uint64_t* o_full = p_free + 2;      // [2] O buffer complete
    uint64_t* o_free = o_full + 2;      // [2] O buffer drained (8 warps)
    uint64_t* a_full = o_free + 2;      // [1] X tile resident in TMEM (8 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

    const int64_t row0 = (int64_t)blockIdx.x * TC_BM;
    const int kc = blockIdx.y;
    const int64_t t_begin = (int64_t)blockIdx.z * p.tiles_per_split;
    const int64_t t_end = min(p.sub_tiles, t_begin + (int64_t)p.tiles_per_split);
    const int T = (int)(t_end - t_begin);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&p_full[b], TC_EPI_WARPS);
            mbar_init(&p_free[b], 1);
            mbar_init(&o_full[b], 1);
            mbar_init(&o_free[b], TC_EPI_WARPS);
        }
        mbar_init(a_full, TC_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t col_a_hi = 0, col_a_lo = KB * 32, col_sp = KB * 64, col_o = KB * 64 + 256;

    const TcHeader* rh = reinterpret_cast<const TcHeader*>(p.rows);
    const TcHeader* ch = reinterpret_cast<const TcHeader*>(p.cols);
    const float* col_norms = reinterpret_cast<const float*>(p.cols + tc_norm_offset());
    const unsigned char* col_images = p.cols + tc_image_offset(p.m);

    if (warp == TC_EPI_WARPS) {
        // =============================== producer ===============================
        if (lane == 0) {
            const unsigned char* a_src = col_images + (size_t)t_begin * a_img_bytes;
            const unsigned char* v_src = p.vimg + ((size_t)kc * p.sub_tiles + t_begin) * v_img_bytes;
            const float* n_src = col_norms + t_begin * TC_BN;
            int s = 0;
            uint32_t ph = 1;  // a fresh barrier passes a wait on parity 1
            for (int u = 0; u < T; ++u) {
                mbar_wait(&empty[s], ph);
                unsigned char* dst = smem + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full[s], stage_bytes + TC_BN * 4);
                bulk_copy_g2s(dst, a_src, a_img_bytes, &full[s]);
                bulk_copy_g2s(dst + a_img_bytes, v_src, v_img_bytes, &full[s]);
                bulk_copy_g2s(ny_smem + s * TC_BN, n_src, TC_BN * 4, &full[s]);
                a_src += a_img_bytes;
                v_src += v_img_bytes;
                n_src += TC_BN;
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == TC_EPI_WARPS + 1) {
        // =============================== MMA issue ===============================
        // The whole warp runs the (warp-uniform) control flow so addresses live in uniform
        // registers; one elected lane issues the tcgen05 instructions.
        constexpr uint32_t idesc1 = umma_idesc(FMT_F16, TC_BN);
        constexpr uint32_t idesc2 = umma_idesc(FMT_TF32, KP);
        const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
        const uint32_t desc_lo0 = (uint32_t)(umma_desc_sw128(0) & 0xFFFFFFFFu);
        const int nk1 = p.nk1;
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t lo_off = (uint32_t)KB * TC_KBLOCK_BYTES;

        // S[b] = X . Y_tile^T : hi.lo + lo.hi + hi.hi, one K = 16 step per instruction
        auto issue_mma1 = [&](int b, int s) {
            const uint32_t d_t = tmem + col_sp + b * 128;
            const uint32_t img = smem_base + (uint32_t)s * stage_bytes;
            const uint32_t dlo_hi = desc_lo0 + (img >> 4);             // Y hi image
            const uint32_t dlo_lo = desc_lo0 + ((img + lo_off) >> 4);  // Y lo image
            if (elect_one()) {
                uint32_t acc = 0;
#pragma unroll 1
                for (int part = 0; part < 3; ++part) {
                    uint32_t a = tmem + (part == 1 ? col_a_lo : col_a_hi);
                    uint32_t bd = (part == 0 ? dlo_lo : dlo_hi);
                    int ks = 0;
#pragma unroll 1
                    for (; ks + 4 <= nk1; ks += 4) {  // one 64-wide K-block: 4 steps of 32 B inside the 128 B row
                        umma_ts<FMT_F16>(d_t, a, bd, desc_hi, idesc1, acc);
                        umma_ts<FMT_F16>(d_t, a + 8, bd + 2, desc_hi, idesc1, 1);
                        umma_ts<FMT_F16>(d_t, a + 16, bd + 4, desc_hi, idesc1, 1);
                        umma_ts<FMT_F16>(d_t, a + 24, bd + 6, desc_hi, idesc1, 1);
                        acc = 1;
                        a += 32;
                        bd += TC_KBLOCK_BYTES >> 4;
                    }
                    for (; ks < nk1; ++ks) {
                        umma_ts<FMT_F16>(d_t, a, bd, desc_hi, idesc1, acc);
                        acc = 1;
                        a += 8;
                        bd += 2;
                    }
                }
            }
            __syncwarp();
        };
        // O[b] = P[b] . V_tile : fresh accumulator per sub-tile, K = 8 (tf32) per instruction
        auto issue_mma2 = [&](int b, int s) {
            const uint32_t d_t = tmem + col_o + b * KP;
            const uint32_t p_hi = tmem + col_sp + b * 128, p_lo = p_hi + 64;
            const uint32_t img = smem_base + (uint32_t)s * stage_bytes + a_img_bytes;
            const uint32_t dlo_hi = desc_lo0 + (img >> 4);
            const uint32_t dlo_lo = desc_lo0 + ((img + KP * 256) >> 4);
            if (elect_one()) {
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const uint32_t a = (part == 1 ? p_lo : p_hi);
                    const uint32_t bd = (part == 0 ? dlo_lo : dlo_hi);
#pragma unroll
                    for (int ks = 0; ks < TC_BN / 8; ++ks) {
                        umma_ts<FMT_TF32>(d_t, a + ks * 8, bd + (ks >> 2) * ((KP * 128) >> 4) + (ks & 3) * 2, desc_hi,
                                          idesc2, (part | ks) != 0);
                    }
                }
            }
            __syncwarp();
        };
        auto commit = [&](uint64_t* bar) {
            if (elect_one()) umma_commit(bar);
            __syncwarp();
        };

        mbar_wait(a_full, 0);
        int s1 = 0, s2 = 0;        // ring slots of the tiles MMA1 / MMA2 work on
        uint32_t ph1 = 0;          // parity of full[s1]
        if (T > 0) {
            mbar_wait(&full[0], 0);
            tc_fence_after();
            issue_mma1(0, 0);
            commit(&s_full[0]);
            if (++s1 == STAGES) {
                s1 = 0;
                ph1 ^= 1;
            }
        }
        for (int u = 0; u < T; ++u) {
            const int b = u & 1;
            const uint32_t par = (uint32_t)((u >> 1) & 1);  // use count parity of buffer b
            if (u + 1 < T) {
                const int b1 = b ^ 1;
                const uint32_t par1 = (uint32_t)(((u + 1) >> 1) & 1);
                mbar_wait(&full[s1], ph1);
                mbar_wait(&p_free[b1], par1 ^ 1);  // MMA2 of tile u-1 has consumed P[b1]
                tc_fence_after();
                issue_mma1(b1, s1);
                commit(&s_full[b1]);
                if (++s1 == STAGES) {
                    s1 = 0;
                    ph1 ^= 1;
                }
            }
            mbar_wait(&p_full[b], par);
            mbar_wait(&o_free[b], par ^ 1);  // O[b] of tile u-2 has been drained
            tc_fence_after();
            issue_mma2(b, s2);
            if (elect_one()) {
                umma_commit(&empty[s2]);
                umma_commit(&p_free[b]);
                umma_commit(&o_full[b]);
            }
            __syncwarp();
            if (++s2 == STAGES) s2 = 0;
        }
    } else {
        // =============================== epilogue warps ===============================
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int h = warp >> 2;  // column half of the sub-tile / of O handled by this warp
        const int row = q * 32 + lane;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        const int64_t grow = row0 + row;
        const float nx = reinterpret_cast<const float*>(p.rows + tc_norm_offset())[grow];
        const float m2c = -2.0f * rh->inv_scale * ch->inv_scale;

        // ---- X tile -> TMEM (h = 0: hi halves, h = 1: lo halves) ----
        {
            const unsigned char* img = p.rows + tc_image_offset(p.n) + (size_t)(grow >> 6) * a_img_bytes +
                                       (size_t)h * KB * TC_KBLOCK_BYTES;
            const int r = (int)(grow & 63);
            for (int kb = 0; kb < KB; ++kb) {
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4*>(img + (size_t)kb * TC_KBLOCK_BYTES + (size_t)r * 128 +
                                                                    (size_t)((c ^ (r & 7)) * 16));
                    w[c * 4 + 0] = v.x;
                    w[c * 4 + 1] = v.y;
                    w[c * 4 + 2] = v.z;
                    w[c * 4 + 3] = v.w;
                }
                tmem_st32(tmem + lane_bits + (h ? col_a_lo : col_a_hi) + kb * 32, w);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }

        float acc[KP / 2];
#pragma unroll
        for (int c = 0; c < KP / 2; ++c) acc[c] = 0.0f;

        // fp32 round-to-nearest accumulation of one sub-tile's O buffer into registers
        auto drain = [&](int b, uint32_t par) {
            mbar_wait(&o_full[b], par);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < KP / 16; ++g) {
                uint32_t o[8];
                tmem_ld8(tmem + lane_bits + col_o + b * KP + h * (KP / 2) + g * 8, o);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[g * 8 + e] += __uint_as_float(o[e]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[b]);
        };

        int s = 0;
        uint32_t phs = 0;
        for (int u = 0; u < T; ++u) {
            const int b = u & 1;
            const uint32_t par = (uint32_t)((u >> 1) & 1);
            mbar_wait(&full[s], phs);  // |y|^2 of this sub-tile is in smem
            mbar_wait(&s_full[b], par);
            tc_fence_after();
            const uint32_t t_s = tmem + lane_bits + col_sp + b * 128 + h * 32;
            uint32_t sv[32];
            tmem_ld32(t_s, sv);
            tmem_wait_ld();
            const float4* nyv = reinterpret_cast<const float4*>(ny_smem + s * TC_BN + h * 32);
            uint32_t lo[32];
#define KMM_TC_PW(KID)                                                           \
    _Pragma("unroll") for (int g = 0; g < 8; ++g) {                              \
        const float4 ny4 = nyv[g];                                               \
        const float nyy[4] = {ny4.x, ny4.y, ny4.z, ny4.w};                       \
        _Pragma("unroll") for (int e = 0; e < 4; ++e) {                          \
            const int c = g * 4 + e;                                             \
            float D = fmaf(__uint_as_float(sv[c]), m2c, nx + nyy[e]);            \
            D = fmaxf(D, 0.0f);                                                  \
            split_tf32(tc_pointwise<KID>(D), sv[c], lo[c]);                      \
        }                                                                        \
    }
            switch (p.kid) {
                case KID_RBF: KMM_TC_PW(KID_RBF) break;
                case KID_MATERN12: KMM_TC_PW(KID_MATERN12) break;
                case KID_MATERN32: KMM_TC_PW(KID_MATERN32) break;
                default: KMM_TC_PW(KID_MATERN52) break;
            }
#undef KMM_TC_PW
            tmem_st32(t_s, sv);        // P_hi in place over S
            tmem_st32(t_s + 64, lo);   // P_lo
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            // drain the previous sub-tile's O while the tensor core works on this one
            if (u > 0) drain(b ^ 1, (uint32_t)(((u - 1) >> 1) & 1));
            if (++s == STAGES) {
                s = 0;
                phs ^= 1;
            }
        }
        if (T > 0) drain((T - 1) & 1, (uint32_t)(((T - 1) >> 1) & 1));

        // ---- write this thread's half row of Y ----
        if (grow < p.n) {
            float* dst = p.out + (int64_t)blockIdx.z * p.split_stride + grow * p.ldo;
#pragma unroll
            for (int c = 0; c < KP / 2; ++c) {
                const int col = kc * KP + h * (KP / 2) + c;
                if (col < p.k) dst[col] = acc[c] * p.scale_out;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, 512);
}

struct TcPlan {
    int kb, kp, k_chunks, stages, splits, tiles_per_split;
    int64_t sub_tiles;
    size_t smem_bytes, vimg_bytes, part_bytes;
};

bool tc_plan(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count, TcPlan* pl) {
    if (d < 1 || d > TC_MAX_D || n < 1 || m < 1 || k < 1) return false;
    const int kb = tc_kblocks(d);
    const int kp_max = kb <= 2 ? 64 : 32;  // TMEM columns: 64*KB (X) + 256 (S/P x2) + 2*KP (O x2) <= 512
    int kp = 16;
    while (kp < kp_max && kp < k) kp *= 2;
    const size_t stage = tc_image_bytes(kb) + (size_t)kp * 512;
    const size_t fixed = 16 * sizeof(uint64_t) + 64;  // barriers + tmem slot (upper bound incl. per-stage barriers below)
    int stages = 4;
    while (stages > 1 && stages * (stage + TC_BN * 4 + 2 * sizeof(uint64_t)) + fixed > (size_t)TC_SMEM_LIMIT) --stages;
    if (stages < 2) return false;
    pl->kb = kb;
    pl->kp = kp;
    pl->k_chunks = (int)((k + kp - 1) / kp);
    pl->stages = stages;
    pl->sub_tiles = (m + TC_BN - 1) / TC_BN;
    pl->smem_bytes = stages * (stage + TC_BN * 4 + 2 * sizeof(uint64_t)) + fixed;
    const int64_t base = ((n + TC_BM - 1) / TC_BM) * pl->k_chunks;
    const int64_t target = (int64_t)sm_count * 2;
    int64_t splits = 1;
    if (base < target) {
        splits = (target + base - 1) / base;
        const int64_t max_splits = (pl->sub_tiles + TC_MIN_SPLIT_TILES - 1) / TC_MIN_SPLIT_TILES;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int64_t tps = (pl->sub_tiles + splits - 1) / splits;
    splits = (pl->sub_tiles + tps - 1) / tps;
    pl->splits = (int)splits;
    pl->tiles_per_split = (int)tps;
    pl->vimg_bytes = (size_t)pl->k_chunks * pl->sub_tiles * kp * 512;
    pl->part_bytes = splits > 1 ? (size_t)splits * n * k * sizeof(float) : 0;
    return true;
}

}  // namespace

bool tc_supported_d(int64_t d) { return d >= 1 && d <= TC_MAX_D; }
bool tc_supported_k(int64_t k) { return k >= 1; }

size_t tc_packed_bytes(int64_t n, int64_t d) {
    if (n <= 0 || !tc_supported_d(d)) return 0;
    return tc_image_offset(n) + (size_t)(tc_npad(n) / 64) * tc_image_bytes(tc_kblocks(d));
}

cudaError_t launch_tc_pack(const float* X, int64_t n, int64_t d, int64_t ldx, const int64_t* idx, float inv_ls,
                           const float* inv_ls_vec, void* packed, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(packed, 0, TC_HEADER_BYTES, stream);
    if (err != cudaSuccess) return err;
    const int64_t total = n * d;
    int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    if (blocks < 1) blocks = 1;
    tc_absmax_kernel<<<blocks, 256, 0, stream>>>(X, n, d, ldx, idx, inv_ls, inv_ls_vec, reinterpret_cast<TcHeader*>(packed));
    const int64_t n_pad = tc_npad(n);
    tc_pack_points_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, stream>>>(
        X, n, d, ldx, idx, inv_ls, inv_ls_vec, static_cast<unsigned char*>(packed), tc_kblocks(d));
    return cudaGetLastError();
}

size_t tc_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count) {
    TcPlan pl;
    if (!tc_plan(n, m, d, k, sm_count, &pl)) return 0;
    return round_up((int64_t)pl.vimg_bytes, 256) + pl.part_bytes;
}

template <int KP>
static cudaError_t launch_tc_kp(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    auto kern = kmm_tc_kernel<KP>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)((n + TC_BM - 1) / TC_BM), (unsigned)pl.k_chunks, (unsigned)pl.splits);
    kern<<<grid, TC_THREADS, pl.smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_tc(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                      const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kid, float scale,
                      int sm_count, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    TcPlan pl;
    if (!tc_plan(n, m, d, k, sm_count, &pl)) return cudaErrorInvalidValue;
    const size_t v_bytes = (size_t)round_up((int64_t)pl.vimg_bytes, 256);
    if (workspace == nullptr || workspace_bytes < v_bytes + pl.part_bytes) return cudaErrorInvalidValue;
    unsigned char* vimg = static_cast<unsigned char*>(workspace);
    float* part = reinterpret_cast<float*>(vimg + v_bytes);
    {
        const int64_t total = (int64_t)pl.k_chunks * pl.sub_tiles * 2 * pl.kp * 8;
        tc_pack_v_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(V, m, k, ldv, vimg, pl.kp, pl.sub_tiles,
                                                                              pl.k_chunks);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    TcParams p;
    p.rows = static_cast<const unsigned char*>(rows_packed);
    p.cols = static_cast<const unsigned char*>(cols_packed);
    p.vimg = vimg;
    p.n = n;
    p.m = m;
    p.k = (int)k;
    p.kb = pl.kb;
    p.nk1 = (int)((d + 15) / 16);
    p.stages = pl.stages;
    p.kid = kid;
    p.sub_tiles = pl.sub_tiles;
    p.tiles_per_split = pl.tiles_per_split;
    if (pl.splits > 1) {
        p.out = part;
        p.ldo = k;
        p.split_stride = n * k;
        p.scale_out = 1.0f;
    } else {
        p.out = Y;
        p.ldo = ldy;
        p.split_stride = 0;
        p.scale_out = scale;
    }
    cudaError_t err;
    switch (pl.kp) {
        case 16: err = launch_tc_kp<16>(p, pl, n, stream); break;
        case 32: err = launch_tc_kp<32>(p, pl, n, stream); break;
        case 64: err = launch_tc_kp<64>(p, pl, n, stream); break;
        default: err = launch_tc_kp<128>(p, pl, n, stream); break;
    }
    if (err != cudaSuccess) return err;
    if (pl.splits > 1) return launch_split_reduce<float>(part, pl.splits, n, k, Y, ldy, scale, stream);
    return cudaSuccess;
}

}  // namespace kmm

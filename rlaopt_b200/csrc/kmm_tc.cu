// tcgen05 / TMEM fused kernel-matrix matmat for the L2-distance kernels (fp32 in/out).
//
//   Y[i,:] = c * sum_j f(|x_i|^2 + |y_j|^2 - 2 x_i.y_j) V[j,:]
//
// flash-attention-shaped, one CTA per 128 output rows, streaming 64-column sub-tiles.
// 12 warps:
//
//   warp 8 (producer)   cp.async.bulk (TMA engine, UBLKCP) of pre-swizzled tile images into two smem
//                       rings: A ring {Y-tile fp16 hi|lo}, V ring {V-tile fp16 hi|lo, 1/s_V, |y|^2}
//   warps 9, 11 (MMA1)  S[b] = X.Y^T   kind::f16, A = X tile resident in TMEM, 3 products
//                       hi.lo + lo.hi + hi.hi, fp32 accumulate; even / odd sub-tiles
//   warp 10 (MMA2)      O[u&1] = P'[b].V'   kind::f16, A = P' from TMEM (in place over S), 3 products
//   warps 0-3 / 4-7     two epilogue warpgroups ping-pong over the sub-tiles; a thread owns one full row:
//                       tcgen05.ld S -> z = c1*S + c2*|y|^2 + c3 (packed FFMA2) -> row extreme ->
//                       P' = f(D) * 2^E (E folded into the ex2 argument, max_j P' in [2^14, 2^15))
//                       -> fp16 hi/lo pair -> tcgen05.st over S -> drain the warpgroup's previous O
//                       buffer: acc += O * 2^-E / s_V with round-to-nearest FFMA2 (the tensor core
//                       accumulates with truncation, so every sub-tile gets a fresh accumulator)
//
// k <= 4 (single right-hand sides) runs in register-contraction mode (template parameter KV): no MMA2 -- the
// epilogue reads S in the 16x256b fragment pattern (4 rows x 16 columns per thread), evaluates f in fp32 and
// multiplies with the raw fp32 V tile; warp 10 feeds the V ring instead of issuing MMA2.
// The X-resident instantiations carry one kernel function each (template parameter KIDT).
//
// Split-precision arithmetic (why fp32 parity holds, DESIGN.md "numerics"):
//   x*s = hi + lo (+2^-22), fp16 pair, s a power of two chosen per operand so |x*s| < 2^13
//   P*2^E and V*s_V likewise (per row and sub-tile / per 64-row V tile power-of-two scales)
// K is never written to HBM; S and P never leave TMEM / registers.
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdlib.h>

#include <type_traits>

#include "kmm_common.cuh"
#include "kmm_launch.h"
#include "kmm_tmem_ldst.cuh"

namespace kmm {
namespace {

constexpr int TC_BM = 128;      // rows per CTA (TMEM lanes)
constexpr int TC_BN = 64;       // K columns per sub-tile
// NWG epilogue warpgroups (warps 0 .. 4 NWG - 1), then four control warps: producer, MMA1 issue (even tiles),
// MMA2 issue, MMA1 issue (odd tiles)
__host__ __device__ constexpr int tc_threads(int nwg) { return (nwg * 4 + 4) * 32; }
constexpr int TC_MIN_SPLIT_TILES = 16;  // a column split covers at least 16 sub-tiles (1024 columns)
constexpr int TC_KBLOCK_BYTES = 64 * 128;  // one K-block of a 64-row image: 64 rows x 128 B
constexpr int TC_HEADER_BYTES = 256;
constexpr int TC_MAX_D = 192;        // X tile resident in TMEM up to here
constexpr int TC_MAX_D_WIDE = 2048;  // beyond: X and Y stream through smem one 64-feature K-block at a time
constexpr uint32_t TC_WIDE_STAGE_BYTES = 8 * TC_KBLOCK_BYTES;  // {X hi, X lo, Y hi, Y lo} x 128 rows x 128 B
constexpr int TC_SMEM_LIMIT = 220 * 1024;

struct TcHeader {
    unsigned int absmax_bits;  // max |(x - center) / lengthscale| as float bits (filled by the absmax kernel)
    float scale;               // power of two s with |x * s| in [2^12, 2^13)
    float inv_scale;
    unsigned int max_sqnorm_bits;  // max_i |(x_i - center) / lengthscale|^2 as float bits: the accuracy guard of the host
    unsigned int bad_index;        // number of gather indices outside [-n_src, n_src) (packed as zero points)
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
// wait on two / three barriers at once: the polls are issued back to back, so the (long) shared-memory
// round trip of a successful poll is paid once instead of once per barrier
__device__ __forceinline__ void mbar_wait2(uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n"
            "and.pred p, p, q;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(b0)), "r"(p0), "r"(smem_u32(b1)), "r"(p1)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait3(uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1, uint64_t* b2, uint32_t p2) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p, q, r;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 r, [%5], %6;\n"
            "and.pred p, p, q;\n"
            "and.pred p, p, r;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(b0)), "r"(p0), "r"(smem_u32(b1)), "r"(p1), "r"(smem_u32(b2)), "r"(p2)
            : "memory");
        if (ok) break;
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
// CTA-pair variants: the copy lands at the same smem offset in every CTA of `mask` and signals the mbarrier at the
// same offset in each of them; the commit arrives on the barrier at that offset in every CTA of `mask`
__device__ __forceinline__ void bulk_copy_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                 uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- cta_group::2 (CTA pair issuing one MMA over both SMs; every tcgen05 instruction of such a kernel carries it) ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrives on the barrier at this smem offset in every CTA of `mask` once all MMAs issued so far by this thread are done
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// instruction descriptor: fp32 accumulate, K-major A and B, M = 128 (256 for a cta_group::2 MMA over a CTA pair)
__host__ __device__ constexpr uint32_t umma_idesc(int ab_format, int n, int m = TC_BM) {
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
constexpr int FMT_F16 = 0;

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, one issue slot for two lanes) ----
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// Pointwise stage, two passes over a thread's 32 entries of one row (DESIGN.md "epilogue"):
//   pass 1  z_j = phi * D_j (Matern; phi = 1, 3, 5) or -D_j/2 * log2(e) (RBF) from the raw MMA1 sums,
//           and the row extreme (min D / max exponent argument);
//   pass 2  P'_j = f(D_j) * 2^E with the power-of-two row scale E folded into the ex2 argument.
constexpr float TC_LOG2E = 1.44269504088896340736f;
template <int KID>
__device__ __forceinline__ float tc_phi() {  // factor folded into the distance before the square root
    return KID == KID_MATERN32 ? 3.0f : (KID == KID_MATERN52 ? 5.0f : 1.0f);
}
// value of the kernel function at z (used once per row for the scale)
template <int KID>
__device__ __forceinline__ float tc_value(float z) {
    const float s = sqrt_approx(fmaxf(z, 0.0f));
    const float e = ex2_approx(-TC_LOG2E * s);
    if constexpr (KID == KID_MATERN12) return e;
    else if constexpr (KID == KID_MATERN32) return (1.0f + s) * e;
    else return fmaf(s, fmaf(s, 1.0f / 3.0f, 1.0f), 1.0f) * e;
}

// Exact squared distance of two packed points (direct differences of the fp16 hi + lo reconstructions), used by the
// Matern-1/2 epilogue for (near-)coincident pairs, where the GEMM-form distance has no relative accuracy.
// `xi` / `yj`: address of the point's row in K-block 0 of the hi image; `sw`: its swizzle key (index & 7).
__device__ __noinline__ float tc_exact_dist2(const unsigned char* xi, int swx, const unsigned char* yj, int swy, int KB,
                                             float inv_sx, float inv_sy) {
    const size_t lo_off = (size_t)KB * (64 * 128);
    float D = 0.0f;
    for (int kb = 0; kb < KB; ++kb) {
        for (int c = 0; c < 8; ++c) {
            const size_t ox = (size_t)kb * (64 * 128) + (size_t)((c ^ swx) * 16);
            const size_t oy = (size_t)kb * (64 * 128) + (size_t)((c ^ swy) * 16);
            const uint4 xh = *reinterpret_cast<const uint4*>(xi + ox), xl = *reinterpret_cast<const uint4*>(xi + lo_off + ox);
            const uint4 yh = *reinterpret_cast<const uint4*>(yj + oy), yl = *reinterpret_cast<const uint4*>(yj + lo_off + oy);
            const __half2* xh2 = reinterpret_cast<const __half2*>(&xh);
            const __half2* xl2 = reinterpret_cast<const __half2*>(&xl);
            const __half2* yh2 = reinterpret_cast<const __half2*>(&yh);
            const __half2* yl2 = reinterpret_cast<const __half2*>(&yl);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(xh2[e]), al = __half22float2(xl2[e]);
                const float2 b = __half22float2(yh2[e]), bl = __half22float2(yl2[e]);
                const float d0 = (a.x + al.x) * inv_sx - (b.x + bl.x) * inv_sy;
                const float d1 = (a.y + al.y) * inv_sx - (b.y + bl.y) * inv_sy;
                D = fmaf(d0, d0, D);
                D = fmaf(d1, d1, D);
            }
        }
    }
    return D;
}

// ------------------------------------------------------------------------------------------
// packing kernels
// ------------------------------------------------------------------------------------------
// gather index with Python semantics: negative indices wrap once; anything still outside [0, n_src) is invalid (-1)
__device__ __forceinline__ int64_t tc_src_row(const int64_t* __restrict__ idx, int64_t i, int64_t n_src) {
    if (!idx) return i;
    int64_t s = idx[i];
    if (s < 0) s += n_src;
    return (s >= 0 && s < n_src) ? s : -1;
}

// abs-max of the centred, scaled points: a block owns TC_ABSMAX_ROWS rows; lanes walk the features of 8 rows at a time
// (128-byte row segments, no index arithmetic per element)
constexpr int TC_ABSMAX_ROWS = 512;
__global__ void __launch_bounds__(256)
tc_absmax_kernel(const float* __restrict__ X, int64_t n, int64_t n_src, int64_t d, int64_t ldx,
                 const int64_t* __restrict__ idx, float inv_ls, const float* __restrict__ inv_ls_vec,
                 const float* __restrict__ center, TcHeader* hdr) {
    const int fx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * TC_ABSMAX_ROWS;
    const int64_t r1 = min(n, r0 + TC_ABSMAX_ROWS);
    float mx = 0.0f;
    for (int64_t i = r0 + ry; i < r1; i += 8) {
        const int64_t src = tc_src_row(idx, i, n_src);
        if (src < 0) continue;
        const float* row = X + src * ldx;
        for (int64_t f = fx; f < d; f += 32) {
            const float s = inv_ls_vec ? inv_ls_vec[f] : inv_ls;
            const float c = center ? center[f] : 0.0f;
            const float v = fabsf((row[f] - c) * s);
            if (v < 3.0e38f) mx = fmaxf(mx, v);  // ignore inf / nan here; they propagate through the values
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ float red[8];
    if (fx == 0) red[ry] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        if (mx > 0.0f) atomicMax(&hdr->absmax_bits, __float_as_uint(mx));
    }
}

// One block per 64-point tile image: fp16 hi/lo split into the swizzled image + squared norms.
//   phase A  the tile's 64 points x 64 features of a K-block go through shared memory (row gather, center, 1/l and the
//            power-of-two scale applied on the way): global reads are 128-byte row segments;
//   phase B  thread (row r, physical 16-byte chunk j) converts logical chunk j ^ (r & 7): a warp stores four whole
//            128-byte image rows (512 contiguous bytes) per instruction, hi then lo.
// `center` (optional, one value per feature, in the units of X): every kernel here is a function of x - y, so the
// same vector subtracted from both operands leaves K unchanged while it shrinks |x|^2 + |y|^2 -- and with it the
// absolute error eps (|x|^2 + |y|^2) of the GEMM-form distance.
__global__ void __launch_bounds__(256)
tc_pack_points_kernel(const float* __restrict__ X, int64_t n, int64_t n_src, int64_t d, int64_t ldx,
                      const int64_t* __restrict__ idx, float inv_ls, const float* __restrict__ inv_ls_vec,
                      const float* __restrict__ center, unsigned char* __restrict__ packed, int kb_count) {
    __shared__ float tile[TC_BN][TC_BN + 1];
    __shared__ int64_t srcs[TC_BN];
    __shared__ float wred[8];
    __shared__ unsigned int bad_cnt;
    TcHeader* hdr = reinterpret_cast<TcHeader*>(packed);
    const float absmax = __uint_as_float(hdr->absmax_bits);
    float s = 1.0f;
    if (absmax > 0.0f) s = ldexpf(1.0f, 12 - ilogbf(absmax));
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid == 0) {
        hdr->scale = s;
        hdr->inv_scale = 1.0f / s;
    }
    const int64_t i0 = (int64_t)blockIdx.x * TC_BN;
    if (tid == 0) bad_cnt = 0;
    __syncthreads();
    if (tid < TC_BN) {
        const int64_t i = i0 + tid;
        const int64_t src = (i < n) ? tc_src_row(idx, i, n_src) : -1;
        srcs[tid] = src;
        if (i < n && src < 0) atomicAdd(&bad_cnt, 1u);
    }
    __syncthreads();
    float* norms = reinterpret_cast<float*>(packed + tc_norm_offset());
    unsigned char* img = packed + tc_image_offset(n) + (size_t)blockIdx.x * tc_image_bytes(kb_count);
    const int j = tid & 7;  // physical 16-byte chunk of the image row
    double nrm[2] = {0.0, 0.0};
    for (int kb = 0; kb < kb_count; ++kb) {
        {   // phase A: feature fa of rows (tid >> 6) + 4 it
            const int fa = tid & 63;
            const int64_t f = (int64_t)kb * 64 + fa;
            const float sc = (f < d) ? (inv_ls_vec ? inv_ls_vec[f] : inv_ls) : 0.0f;
            const float cf = (f < d && center) ? center[f] : 0.0f;
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
                const int r = (tid >> 6) + 4 * it;
                const int64_t src = srcs[r];
                float v = 0.0f;
                if (src >= 0 && f < d) v = (X[src * ldx + f] - cf) * sc * s;
                tile[r][fa] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // phase B
            const int r = (tid >> 3) + 32 * half;
            const int c = j ^ (r & 7);  // logical chunk stored at physical position j
            alignas(16) __half hi[8];
            alignas(16) __half lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float v = tile[r][c * 8 + e];
                const __half h = __float2half_rn(v);
                const __half l = __float2half_rn(v - __half2float(h));
                hi[e] = h;
                lo[e] = l;
                const double eff = (double)__half2float(h) + (double)__half2float(l);
                nrm[half] += eff * eff;
            }
            const size_t off = (size_t)kb * TC_KBLOCK_BYTES + (size_t)r * 128 + (size_t)(j * 16);
            *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(img + (size_t)kb_count * TC_KBLOCK_BYTES + off) = *reinterpret_cast<const uint4*>(lo);
        }
        __syncthreads();
    }
    // row norms: the eight lanes of a row hold the partial sums of their chunks (fp64, fixed order)
    float wmax = 0.0f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        double t = nrm[half];
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        const float nrm_f = (float)(t / ((double)s * (double)s));
        if (j == 0) norms[i0 + (tid >> 3) + 32 * half] = nrm_f;
        if (nrm_f < 3.0e38f) wmax = fmaxf(wmax, nrm_f);
    }
    // accuracy guard input: the largest squared norm of the set (non-negative floats order like their bit patterns)
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((tid & 31) == 0) wred[tid >> 5] = wmax;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) wmax = fmaxf(wmax, wred[w]);
        if (wmax > 0.0f) atomicMax(&hdr->max_sqnorm_bits, __float_as_uint(wmax));
        if (bad_cnt > 0) atomicAdd(&hdr->bad_index, bad_cnt);
    }
}

// bytes of one V record in the workspace: fp16 image (hi | lo), 16 B trailer {1 / s}, 64 column norms
__host__ __device__ constexpr size_t tc_v_record_bytes(int kp) { return (size_t)kp * 256 + 16 + TC_BN * 4; }

// V[m][k] -> per (k-chunk, 64-row sub-tile) image of the MMA2 B operand, one block per image:
//   [hi | lo] x [row c of KP] x 128 B (64 j as fp16, 16 B chunks XOR-swizzled by c & 7) | 16 B trailer {1/s}
// The tile is scaled by a power of two s (|v s| in [2^14, 2^15)) before the fp16 hi/lo split, so
// the pair keeps 22 significant bits of every element within 2^-29 of the tile maximum.
//   | 64 squared column norms (copied from the packed column operand): one record = one bulk copy per V-ring stage
__global__ void __launch_bounds__(256) tc_pack_v_kernel(const float* __restrict__ V, int64_t m, int64_t k, int64_t ldv,
                                                        const float* __restrict__ col_norms,
                                                        unsigned char* __restrict__ images, int kp, int64_t sub_tiles) {
    extern __shared__ float vt[];  // [64][kp + 1]
    __shared__ float red[8];
    const int64_t t = blockIdx.x;
    const int kc = blockIdx.y;
    const int tid = threadIdx.x;
    const int pitch = kp + 1;
    float mx = 0.0f;
    for (int e = tid; e < TC_BN * kp; e += 256) {
        const int j = e / kp, c = e % kp;
        const int64_t row = t * TC_BN + j, col = (int64_t)kc * kp + c;
        const float v = (row < m && col < k) ? V[row * ldv + col] : 0.0f;
        vt[j * pitch + c] = v;
        const float a = fabsf(v);
        if (a < 3.0e38f) mx = fmaxf(mx, a);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    int E = 0;
    if (mx > 0.0f) E = 14 - ilogbf(mx);
    E = max(-126, min(126, E));
    const float s = __uint_as_float((uint32_t)(127 + E) << 23);
    const size_t image_bytes = tc_v_record_bytes(kp);
    unsigned char* img = images + ((size_t)kc * sub_tiles + t) * image_bytes;
    if (tid < TC_BN)  // the packed operand holds norms for all padded points
        reinterpret_cast<float*>(img + (size_t)kp * 256 + 16)[tid] = col_norms[t * TC_BN + tid];
    for (int ch = tid; ch < kp * 8; ch += 256) {
        const int c = ch >> 3, q8 = ch & 7;
        alignas(16) __half hi[8];
        alignas(16) __half lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float v = vt[(q8 * 8 + e) * pitch + c] * s;
            const __half h = __float2half_rn(v);
            hi[e] = h;
            lo[e] = __float2half_rn(v - __half2float(h));
        }
        const size_t off = (size_t)c * 128 + (size_t)((q8 ^ (c & 7)) * 16);
        *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(img + (size_t)kp * 128 + off) = *reinterpret_cast<const uint4*>(lo);
    }
    if (tid == 0) {
        float4 tr;
        tr.x = __uint_as_float((uint32_t)(127 - E) << 23);  // 1 / s
        tr.y = tr.z = tr.w = 0.0f;
        *reinterpret_cast<float4*>(img + (size_t)kp * 256) = tr;
    }
}

// Register-contraction mode (k <= 4): V[m][k] and the column norms -> one record per sub-tile,
// [kv][64] zero-padded fp32 values of V then the 64 squared norms: one bulk copy per sub-tile fills a V-ring stage
__global__ void tc_pack_v_small_kernel(const float* __restrict__ V, int64_t m, int64_t k, int64_t ldv,
                                       const float* __restrict__ col_norms, float* __restrict__ tiles, int kv,
                                       int64_t sub_tiles) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int rec = (kv + 1) * TC_BN;
    if (e >= sub_tiles * rec) return;
    const int j = (int)(e % TC_BN);
    const int c = (int)((e / TC_BN) % (kv + 1));
    const int64_t row = (e / rec) * TC_BN + j;
    float v;
    if (c == kv) v = col_norms[row];  // the packed operand holds norms for all padded points
    else v = (row < m && c < k) ? V[row * ldv + c] : 0.0f;
    tiles[e] = v;
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
#ifdef KMM_TC_PROFILE
__device__ long long g_tc_prof[64];
#define TC_DIAG(bit) ((p.diag & (bit)) != 0)  // RLAOPT_B200_TC_DIAG knock-outs exist in the profile build only
#define TC_PROF_DECL long long prof_t = clock64(); long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define TC_PROF(i) { const long long now_ = clock64(); prof_acc[i] += now_ - prof_t; prof_t = now_; }
#define TC_PROF_FLUSH(base) if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) { for (int i_ = 0; i_ < 8; ++i_) g_tc_prof[(base) + i_] = prof_acc[i_]; }
#else
#define TC_DIAG(bit) false
#define TC_PROF_DECL
#define TC_PROF(i) {}
#define TC_PROF_FLUSH(base) {}
#endif

struct TcParams {
    const unsigned char* rows;  // packed row operand
    const unsigned char* cols;  // packed column operand
    const unsigned char* vimg;  // packed V images
    float* out;
    int64_t ldo, split_stride;
    int64_t n, m;
    int k, kb, nk1, kid;
    int k_chunks;            // chunks of KP columns of V (gridDim.y, or twice gridDim.y for the two-chunk kernels)
    int dual_mode;           // two-chunk kernels: 1 = drains after the pointwise stage, 2 = pointwise stage sliced between the drains
    int dual_overlap;        // sliced mode: 1 = the quarter's loads overlap the drain's (A/B knob)
    int a_stages, v_stages;  // smem ring depths (column-tile images / V images + norms)
    int nb, la;              // S/P buffers in TMEM; la = extra V-ring depth (V of tile t is consumed la tiles after its A image)
    int wide;                // 1: d > 192, feature-chunked MMA1 with X and Y K-blocks streamed through the A ring
    int pair;                // 1: launched as clusters of two CTAs that share every column-tile load (multicast halves)
    int diag;                // RLAOPT_B200_TC_DIAG knock-outs (-DKMM_TC_PROFILE build only): 1 no MMA, 2 no pointwise, 4 no drain, 8 no loads
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

// tcgen05.mma kind::f16, A from TMEM, B through a shared-memory descriptor passed as (lo, hi)
// words: only the low word (start address) changes between K steps, so advancing an operand is
// one 32-bit add.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                        uint32_t idesc, uint32_t acc) {
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
}

// the same MMA issued by the leader CTA of a pair for both SMs: D and A at the same TMEM address in each CTA, B split
// along N between the two CTAs' shared memory (each holds its N / 2 rows at the descriptor's offset)
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                         uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 bd;\n"
        "mov.b64 bd, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], bd, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]) {
    if constexpr (N == 32) tmem_ld32(taddr, r);
    else if constexpr (N == 16) tmem_ld16(taddr, r);
    else tmem_ld8(taddr, r);
}

// tcgen05.mma kind::f16 with both operands in shared memory (wide-d variant)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi,
                                        uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 ad, bd;\n"
        "mov.b64 ad, {%1, %3};\n"
        "mov.b64 bd, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}

// shared-memory footprint of one V-ring stage: image (hi | lo), 16 B trailer, 64 column norms; 1 KB aligned
__host__ __device__ constexpr uint32_t tc_v_stage_bytes(int kp) { return (uint32_t)kp * 256 + 1024; }

// TMEM column map (512 columns x 128 lanes, 32-bit words):
//   [0, 32 KB)                   X tile, fp16 hi halves (two features per column)
//   [32 KB, 64 KB)               X tile, fp16 lo halves
//   [64 KB, +64 NB)              NB S/P buffers: S (64 fp32 columns) is overwritten in place by
//                                P_hi (32 columns of fp16 pairs) | P_lo (32 columns)
//   [64 KB + 64 NB, +2 KP)       two O buffers (one fresh accumulator per sub-tile, alternating)
// M12: instantiation for Matern-1/2 (carries the near-point recompute; the other kernels stay free of its code)
// WIDE: d > 192 instantiation (K-block streaming MMA1); kept out of the common kernels, whose C2 throughput drops
// by 2 % when that code shares their instruction footprint
// KV > 0: k <= KV <= 4 columns of V are contracted on the CUDA cores (no MMA2, no P' split, no O buffers): the
// epilogue thread that owns a row multiplies its 64 kernel values with the raw fp32 V tile from the V ring.
// KIDT >= 0: kernel id fixed at compile time (register-contraction instantiations: the fully unrolled epilogue is
// sensitive to its instruction footprint, so it carries the code of one kernel function only)
// CG2: CTA pair with cta_group::2 MMAs (k > 64 family).  The leader CTA (cluster rank 0) issues MMA1 and MMA2 for both
// SMs (M = 256: each SM computes its own 128 rows); every CTA loads only HALF of each column-tile image and of each V
// image (its N / 2 rows of the B operand) into its own shared memory -- the per-SM ingress that bounds this family
// (48 KB per 1152 tensor cycles = 42 B/clk, the L2 -> SM limit) is halved.  The peer's epilogue warps arrive on the
// leader's barriers through the cluster window; two of its idle issue warps relay its "tile landed" barriers.
// DUAL (k > 128): one CTA contracts the SAME P' of a sub-tile with TWO 128-column chunks of V -- S = X.Y^T and the
// pointwise stage are computed once per 256 columns of V instead of once per 128.  The two O buffers are the two chunks
// (chunk A drains while MMA2 works on chunk B and vice versa); a thread carries 2 x 64 accumulators, which fits next
// to the pointwise stage only because that stage re-reads S from TMEM in quarters (the quarter-row epilogue).
template <int KP, int NWG, bool M12, bool WIDE, int KV = 0, int KIDT = -1, bool CG2 = false, bool DUAL = false>
__global__ void __launch_bounds__(tc_threads(NWG), 1) kmm_tc_kernel(const TcParams p) {
    static_assert(KV == 0 || (KV <= 4 && KP == 16 && !WIDE), "register contraction: k <= 4, X resident in TMEM");
    static_assert(!CG2 || (KP == 128 && !WIDE && KV == 0 && NWG == 2), "cta_group::2 is built for the k > 64 family");
    static_assert(!DUAL || (KP == 128 && NWG == 2 && !WIDE && KV == 0 && !M12 && !CG2), "two chunks per P': k > 128 family");
    constexpr int TC_EPI_WARPS = NWG * 4;
    constexpr int NOB = NWG;  // O buffers: one per epilogue warpgroup (tile u accumulates into O[u % NWG])
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = p.kb, SA = p.a_stages, SV = p.v_stages, NB = p.nb;
    const uint32_t a_img_bytes = (uint32_t)tc_image_bytes(KB);  // hi + lo image of one 64-point tile in HBM
    const uint32_t a_stage_bytes = WIDE ? TC_WIDE_STAGE_BYTES : a_img_bytes;  // one slot of the A ring
    // V tile as stored in HBM: fp16 image hi + lo + trailer, or (KV) the raw fp32 tile [KV][64]
    constexpr uint32_t v_img_bytes = KV ? (KV + 1) * 256 : KP * 256 + 16 + TC_BN * 4;  // the record ends with the 64 column norms
    constexpr uint32_t v_stage_bytes = tc_v_stage_bytes(KP);
    constexpr uint32_t v_norm_off = KV ? KV * 256 : KP * 256 + 16;
    unsigned char* a_ring = smem;
    unsigned char* v_ring = smem + (size_t)SA * a_stage_bytes;
    float* xchg = reinterpret_cast<float*>(v_ring + (size_t)SV * v_stage_bytes);  // [8][128] per-row tile scales (KP = 128 mode)
    // DUAL: one scale per row, tile and chunk, then [2 warpgroups][2][128] floats of column norms + V scales (dual_mode 2)
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + (DUAL ? 16 * TC_BM + 512 : 8 * TC_BM));
    uint64_t* a_full = bars;             // [SA] producer -> MMA1
    uint64_t* a_empty = a_full + SA;     // [SA] MMA1 done -> producer
    uint64_t* v_full = a_empty + SA;     // [SV] producer -> MMA2 / epilogue (norms, V scale)
    uint64_t* v_empty = v_full + SV;     // [SV] MMA2 done -> producer
    uint64_t* s_full = v_empty + SV;     // [NB] MMA1 done
    uint64_t* p_full = s_full + NB;      // [NB] P written (8 warps)
    uint64_t* p_free = p_full + NB;      // [NB] MMA2 done reading P
    uint64_t* o_full = p_free + NB;      // [NOB] O buffer complete
    uint64_t* o_free = o_full + NOB;     // [NOB] O buffer drained
    uint64_t* x_full = o_free + NOB;     // [1] X tile resident in TMEM (8 warps)
    uint64_t* a_peer = x_full + 1;       // [SA] CG2, leader: the peer's half of the column-tile image has landed
    uint64_t* v_peer = a_peer + SA;      // [SV] CG2, leader: the peer's half of the V image has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_peer + SV);

    const int64_t row0 = (int64_t)blockIdx.x * TC_BM;
    const int kc = DUAL ? 2 * blockIdx.y : blockIdx.y;  // DUAL: chunks kc and kc + 1 (the second may not exist)
    const int kc_b = (DUAL && kc + 1 < p.k_chunks) ? kc + 1 : kc;  // an absent second chunk re-reads the first (never stored)
    const int64_t t_begin = (int64_t)blockIdx.z * p.tiles_per_split;
    const int64_t t_end = min(p.sub_tiles, t_begin + (int64_t)p.tiles_per_split);
    const int T = (int)(t_end - t_begin);

    if (threadIdx.x == 0) {
        // multicast pairs: a ring slot is free once both CTAs of the pair have read it; CG2: one MMA reads both halves
        const uint32_t consumers = (p.pair && !CG2) ? 2 : 1;
        const uint32_t ctas = CG2 ? 2 : 1;  // CG2: the peer's epilogue warps arrive on the leader's barriers too
        for (int s = 0; s < SA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], consumers);
            mbar_init(&a_peer[s], 1);
        }
        for (int s = 0; s < SV; ++s) {
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], KV ? 4 : consumers);  // KV: released by the four warps that read the stage
            mbar_init(&v_peer[s], 1);
        }
        for (int b = 0; b < NB; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&p_full[b], 4 * ctas);  // the four warps of the owning warpgroup
            mbar_init(&p_free[b], KV ? 4 : 1);  // KV: S[b] is free once its four warps hold it in registers
        }
        for (int b = 0; b < NOB; ++b) {
            mbar_init(&o_full[b], 1);
            mbar_init(&o_free[b], (KP > 64 ? 8 : 4) * ctas);
        }
        mbar_init(x_full, 8 * ctas);  // warpgroups 0 and 1 load the X tile
        fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS) {
        if constexpr (CG2) tmem_alloc2(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (p.pair) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t x_cols = WIDE ? 0 : KB * 64;  // wide-d: X is not resident in TMEM
    const uint32_t col_a_hi = 0, col_a_lo = KB * 32, col_sp = x_cols, col_o = x_cols + NB * 64;

    // CG2: barriers the leader's issue warps wait on collect arrivals from both CTAs (cluster-scope arrive on rank 0)
    auto arrive_pair = [&](uint64_t* bar) {
        if constexpr (CG2) mbar_arrive_cluster(bar, 0);
        else mbar_arrive(bar);
    };
    const uint32_t crank2 = CG2 ? cluster_ctarank() : 0;

    const TcHeader* rh = reinterpret_cast<const TcHeader*>(p.rows);
    const TcHeader* ch = reinterpret_cast<const TcHeader*>(p.cols);
    const float* col_norms = reinterpret_cast<const float*>(p.cols + tc_norm_offset());
    const unsigned char* col_images = p.cols + tc_image_offset(p.m);

    if (warp >= TC_EPI_WARPS) {
    // registers move from this warpgroup (producer, MMA issue, two idle warps) to the epilogue warpgroups
    if constexpr (NWG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // NWG = 3: 152 for the epilogue, NWG = 4: 104
    if (warp == TC_EPI_WARPS) {
        // =============================== producer ===============================
        if (lane == 0) {
            const unsigned char* a_src = col_images + (size_t)t_begin * a_img_bytes;
            const unsigned char* v_src = p.vimg + ((size_t)kc * p.sub_tiles + t_begin) * v_img_bytes;
            const unsigned char* v_src_b = p.vimg + ((size_t)kc_b * p.sub_tiles + t_begin) * v_img_bytes;  // DUAL: second chunk
            int sa = 0, sv = 0;
            uint32_t pha = 1, phv = 1;  // a fresh barrier passes a wait on parity 1
            const uint32_t crank = p.pair ? cluster_ctarank() : 0;
            if constexpr (WIDE) {
                // Wide d: per segment of two column tiles one A-ring slot per 64-feature K-block, then their V images;
                // a slot holds {X hi, X lo} of this CTA's 128 rows and {Y hi, Y lo} of the segment's 128 columns.
                const bool rows_live = row0 < tc_npad(p.n);
                const unsigned char* x_img = p.rows + tc_image_offset(p.n) + (size_t)(row0 >> 6) * a_img_bytes;
                const size_t lo_img = (size_t)KB * TC_KBLOCK_BYTES;
                for (int u0 = 0; u0 < T; u0 += 2) {
                    const int cnt = (u0 + 1 < T) ? 2 : 1;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&a_empty[sa], pha);
                        unsigned char* dst = a_ring + (size_t)sa * a_stage_bytes;
                        mbar_arrive_expect_tx(&a_full[sa], (rows_live ? 4u : 0u) * TC_KBLOCK_BYTES + 2u * cnt * TC_KBLOCK_BYTES);
                        const size_t koff = (size_t)kb * TC_KBLOCK_BYTES;
                        if (rows_live) {
                            bulk_copy_g2s(dst, x_img + koff, TC_KBLOCK_BYTES, &a_full[sa]);
                            bulk_copy_g2s(dst + TC_KBLOCK_BYTES, x_img + a_img_bytes + koff, TC_KBLOCK_BYTES, &a_full[sa]);
                            bulk_copy_g2s(dst + 2 * TC_KBLOCK_BYTES, x_img + lo_img + koff, TC_KBLOCK_BYTES, &a_full[sa]);
                            bulk_copy_g2s(dst + 3 * TC_KBLOCK_BYTES, x_img + a_img_bytes + lo_img + koff, TC_KBLOCK_BYTES, &a_full[sa]);
                        }
                        for (int e = 0; e < cnt; ++e) {
                            const unsigned char* y_img = a_src + (size_t)e * a_img_bytes + koff;
                            if (!p.pair) {
                                bulk_copy_g2s(dst + (4 + e) * TC_KBLOCK_BYTES, y_img, TC_KBLOCK_BYTES, &a_full[sa]);
                                bulk_copy_g2s(dst + (6 + e) * TC_KBLOCK_BYTES, y_img + lo_img, TC_KBLOCK_BYTES, &a_full[sa]);
                            } else if ((uint32_t)e == crank || cnt == 1 && crank == 0) {
                                // CTA pair: the two CTAs have different rows (own X) but the same columns -- each fetches
                                // one of the segment's two Y tiles and multicasts it to both
                                bulk_copy_g2s_mc(dst + (4 + e) * TC_KBLOCK_BYTES, y_img, TC_KBLOCK_BYTES, &a_full[sa], 3);
                                bulk_copy_g2s_mc(dst + (6 + e) * TC_KBLOCK_BYTES, y_img + lo_img, TC_KBLOCK_BYTES, &a_full[sa], 3);
                            }
                        }
                        if (++sa == SA) {
                            sa = 0;
                            pha ^= 1;
                        }
                    }
                    // V images and column norms of the segment: needed only after its pointwise stage
                    for (int e = 0; e < cnt; ++e) {
                        mbar_wait(&v_empty[sv], phv);
                        unsigned char* vdst = v_ring + (size_t)sv * v_stage_bytes;
                        mbar_arrive_expect_tx(&v_full[sv], v_img_bytes);
                        if (!p.pair) {
                            bulk_copy_g2s(vdst, v_src, v_img_bytes, &v_full[sv]);
                        } else if (crank == 0) {
                            bulk_copy_g2s_mc(vdst, v_src, KP * 128, &v_full[sv], 3);
                        } else {
                            bulk_copy_g2s_mc(vdst + KP * 128, v_src + KP * 128, v_img_bytes - KP * 128, &v_full[sv], 3);
                        }
                        v_src += v_img_bytes;
                        if (++sv == SV) {
                            sv = 0;
                            phv ^= 1;
                        }
                    }
                    a_src += (size_t)cnt * a_img_bytes;
                }
            } else
            for (int u = 0; u < T; ++u) {
                mbar_wait(&a_empty[sa], pha);
                if (TC_DIAG(8)) {
                    mbar_arrive(&a_full[sa]);
                    mbar_wait(&v_empty[sv], phv);
                    mbar_arrive(&v_full[sv]);
                } else {
                if constexpr (KV > 0) {
                    // the V ring has its own producer (warp TC_EPI_WARPS + 2): at ~700 cycles per sub-tile one thread
                    // issuing both rings is the bottleneck
                    mbar_arrive_expect_tx(&a_full[sa], a_img_bytes);
                    bulk_copy_g2s(a_ring + (size_t)sa * a_img_bytes, a_src, a_img_bytes, &a_full[sa]);
                } else if constexpr (CG2) {
                    // this CTA's half of the B operands, into its own shared memory: points [32 r, 32 r + 32) of every
                    // K-block of the column tile (hi and lo), rows [64 r, 64 r + 64) of the V image (hi and lo), and
                    // the whole trailer + column norms (both CTAs' epilogues read them)
                    unsigned char* adst = a_ring + (size_t)sa * a_img_bytes;
                    mbar_arrive_expect_tx(&a_full[sa], a_img_bytes / 2);
                    for (int part = 0; part < 2; ++part)
                        for (int kb = 0; kb < KB; ++kb)
                            bulk_copy_g2s(adst + (size_t)(part * KB + kb) * (TC_KBLOCK_BYTES / 2),
                                          a_src + (size_t)(part * KB + kb) * TC_KBLOCK_BYTES + crank * (TC_KBLOCK_BYTES / 2),
                                          TC_KBLOCK_BYTES / 2, &a_full[sa]);
                    mbar_wait(&v_empty[sv], phv);
                    unsigned char* vdst = v_ring + (size_t)sv * v_stage_bytes;
                    constexpr uint32_t v_q = KP * 64;  // half of the hi (or lo) image: 64 rows x 128 B
                    mbar_arrive_expect_tx(&v_full[sv], 2 * v_q + 16 + TC_BN * 4);
                    bulk_copy_g2s(vdst, v_src + crank * v_q, v_q, &v_full[sv]);
                    bulk_copy_g2s(vdst + v_q, v_src + KP * 128 + crank * v_q, v_q, &v_full[sv]);
                    bulk_copy_g2s(vdst + KP * 256, v_src + KP * 256, 16 + TC_BN * 4, &v_full[sv]);
                } else if (!p.pair) {
                    mbar_arrive_expect_tx(&a_full[sa], a_img_bytes);
                    bulk_copy_g2s(a_ring + (size_t)sa * a_img_bytes, a_src, a_img_bytes, &a_full[sa]);
                    mbar_wait(&v_empty[sv], phv);
                    unsigned char* vdst = v_ring + (size_t)sv * v_stage_bytes;
                    mbar_arrive_expect_tx(&v_full[sv], v_img_bytes);
                    bulk_copy_g2s(vdst, v_src, v_img_bytes, &v_full[sv]);  // image, 1 / s_V and column norms: one record
                } else {
                    // each CTA of the pair fetches half of every image and multicasts it to both
                    const uint32_t a_half = a_img_bytes / 2;
                    mbar_arrive_expect_tx(&a_full[sa], a_img_bytes);
                    bulk_copy_g2s_mc(a_ring + (size_t)sa * a_img_bytes + crank * a_half, a_src + crank * a_half, a_half,
                                     &a_full[sa], 3);
                    mbar_wait(&v_empty[sv], phv);
                    unsigned char* vdst = v_ring + (size_t)sv * v_stage_bytes;
                    mbar_arrive_expect_tx(&v_full[sv], v_img_bytes);
                    constexpr uint32_t v_half = KP * 128;
                    if (crank == 0) {
                        bulk_copy_g2s_mc(vdst, v_src, v_half, &v_full[sv], 3);
                    } else {
                        bulk_copy_g2s_mc(vdst + v_half, v_src + v_half, v_img_bytes - v_half, &v_full[sv], 3);
                    }
                }
                if constexpr (DUAL) {
                    // the record of the second chunk of this sub-tile goes to the next V-ring stage
                    if (++sv == SV) {
                        sv = 0;
                        phv ^= 1;
                    }
                    mbar_wait(&v_empty[sv], phv);
                    unsigned char* vdst = v_ring + (size_t)sv * v_stage_bytes;
                    mbar_arrive_expect_tx(&v_full[sv], v_img_bytes);
                    constexpr uint32_t v_half = KP * 128;
                    if (!p.pair) bulk_copy_g2s(vdst, v_src_b, v_img_bytes, &v_full[sv]);
                    else if (crank == 0) bulk_copy_g2s_mc(vdst, v_src_b, v_half, &v_full[sv], 3);
                    else bulk_copy_g2s_mc(vdst + v_half, v_src_b + v_half, v_img_bytes - v_half, &v_full[sv], 3);
                    v_src_b += v_img_bytes;
                }
                }
                a_src += a_img_bytes;
                v_src += v_img_bytes;
                if (++sa == SA) {
                    sa = 0;
                    pha ^= 1;
                }
                if (++sv == SV) {
                    sv = 0;
                    phv ^= 1;
                }
            }
        }
    } else {
        // =============================== MMA issue (warps 9, 11: MMA1, warp 10: MMA2) ===============================
        // The whole warp runs the (warp-uniform) control flow so addresses live in uniform
        // registers; one elected lane issues the tcgen05 instructions.  The MMA1 issuers run ahead of
        // MMA2 as far as the NB S/P buffers allow, so neither MMA waits for the other's completion.
        constexpr uint32_t idesc1 = umma_idesc(FMT_F16, TC_BN, CG2 ? 2 * TC_BM : TC_BM);
        constexpr uint32_t idesc2 = umma_idesc(FMT_F16, KP, CG2 ? 2 * TC_BM : TC_BM);
        // CG2: a CTA's slot holds its half of each K-block (32 points x 128 B) and of each V image (64 rows x 128 B)
        constexpr uint32_t kblock_bytes = CG2 ? TC_KBLOCK_BYTES / 2 : TC_KBLOCK_BYTES;
        constexpr uint32_t v_lo_off = CG2 ? KP * 64 : KP * 128;
        auto mma_ts = [](uint32_t d_t, uint32_t a_t, uint32_t dlo, uint32_t dhi, uint32_t idesc, uint32_t acc) {
            if constexpr (CG2) umma_ts2(d_t, a_t, dlo, dhi, idesc, acc);
            else umma_ts(d_t, a_t, dlo, dhi, idesc, acc);
        };
        auto commit_bar = [&](uint64_t* bar, bool shared_slot) {  // shared_slot: a ring slot both CTAs of a pair read
            if constexpr (CG2) umma_commit2_mc(bar, 3);
            else if (shared_slot && p.pair) umma_commit_mc(bar, 3);
            else umma_commit(bar);
        };
        const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
        const uint32_t desc_lo0 = (uint32_t)(umma_desc_sw128(0) & 0xFFFFFFFFu);
        const int nk1 = p.nk1;
        const uint32_t a_ring_base = smem_u32(a_ring), v_ring_base = smem_u32(v_ring);
        const uint32_t lo_off = (uint32_t)KB * kblock_bytes;

        // S[b] = X . Y_tile^T : hi.lo + lo.hi + hi.hi, one K = 16 step per instruction
        auto issue_mma1 = [&](int b, int s) {
            const uint32_t d_t = tmem + col_sp + b * 64;
            const uint32_t img = a_ring_base + (uint32_t)s * a_img_bytes;
            const uint32_t dlo_hi = desc_lo0 + (img >> 4);             // Y hi image
            const uint32_t dlo_lo = desc_lo0 + ((img + lo_off) >> 4);  // Y lo image
            if (!TC_DIAG(1) && elect_one()) {
                uint32_t acc = 0;
#pragma unroll 1
                for (int part = 0; part < 3; ++part) {
                    uint32_t a = tmem + (part == 1 ? col_a_lo : col_a_hi);
                    uint32_t bd = (part == 0 ? dlo_lo : dlo_hi);
                    int ks = 0;
#pragma unroll 1
                    for (; ks + 4 <= nk1; ks += 4) {  // one 64-wide K-block: 4 steps of 32 B inside the 128 B row
                        mma_ts(d_t, a, bd, desc_hi, idesc1, acc);
                        mma_ts(d_t, a + 8, bd + 2, desc_hi, idesc1, 1);
                        mma_ts(d_t, a + 16, bd + 4, desc_hi, idesc1, 1);
                        mma_ts(d_t, a + 24, bd + 6, desc_hi, idesc1, 1);
                        acc = 1;
                        a += 32;
                        bd += kblock_bytes >> 4;
                    }
                    for (; ks < nk1; ++ks) {
                        mma_ts(d_t, a, bd, desc_hi, idesc1, acc);
                        acc = 1;
                        a += 8;
                        bd += 2;
                    }
                }
            }
            __syncwarp();
        };
        // O[ob] = P[b] . V_tile : fresh accumulator per sub-tile, fp16 pairs, K = 16 per instruction
        auto issue_mma2 = [&](int b, int ob, int s) {
            const uint32_t d_t = tmem + col_o + ob * KP;
            // P' layout in the S/P buffer: hi pairs in columns [0, 32), lo pairs in [32, 64); the quarter-row epilogue
            // (NWG = 4, DUAL) interleaves them per 16 entries: hi of K-step ks at 16 ks, lo at 16 ks + 8
            constexpr bool interleaved = NWG == 4 || DUAL;
            constexpr uint32_t p_step = interleaved ? 16 : 8;
            const uint32_t p_hi = tmem + col_sp + b * 64, p_lo = p_hi + (interleaved ? 8 : 32);
            const uint32_t img = v_ring_base + (uint32_t)s * v_stage_bytes;
            const uint32_t dlo_hi = desc_lo0 + (img >> 4);
            const uint32_t dlo_lo = desc_lo0 + ((img + v_lo_off) >> 4);
            if (!TC_DIAG(1) && elect_one()) {
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const uint32_t a = (part == 1 ? p_lo : p_hi);
                    const uint32_t bd = (part == 0 ? dlo_lo : dlo_hi);
#pragma unroll
                    for (int ks = 0; ks < TC_BN / 16; ++ks) {
                        mma_ts(d_t, a + ks * p_step, bd + ks * 2, desc_hi, idesc2, (part | ks) != 0);
                    }
                }
            }
            __syncwarp();
        };

        if (warp != TC_EPI_WARPS + 2) {
            // ---- MMA1 issuers: warp 9 takes the even sub-tiles, warp 11 the odd ones, each running ahead of
            // MMA2 as far as the NB S/P buffers and the A ring allow; while one of them polls its barriers the
            // other's instructions keep the tensor pipe fed ----
            const int par = (warp == TC_EPI_WARPS + 1) ? 0 : 1;
            if constexpr (WIDE) {
                // Wide d (warp 9 only): S[b0], S[b0 + 1] of a two-tile segment accumulate over the K-blocks with
                // M128 x N128 x K16 MMAs, both operands from the A-ring slot (SS mode runs at the full rate for N = 128)
                if (par == 0) {
                    constexpr uint32_t idesc_w = umma_idesc(FMT_F16, 2 * TC_BN);
                    int sa = 0;
                    uint32_t pha = 0;
                    for (int u0 = 0; u0 < T; u0 += 2) {
                        const int b0 = u0 % NB;                        // NB = 4: segments alternate between buffers {0,1} and {2,3}
                        const uint32_t use = (uint32_t)((u0 / NB) & 1);
                        mbar_wait2(&p_free[b0], use ^ 1, &p_free[b0 + 1], use ^ 1);  // MMA2 of the segment NB tiles back
                        const uint32_t d_t = tmem + col_sp + b0 * 64;
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(&a_full[sa], pha);
                            tc_fence_after();
                            const uint32_t st = a_ring_base + (uint32_t)sa * a_stage_bytes;
                            const uint32_t xh = desc_lo0 + (st >> 4), xl = desc_lo0 + ((st + 2 * TC_KBLOCK_BYTES) >> 4);
                            const uint32_t yh = desc_lo0 + ((st + 4 * TC_KBLOCK_BYTES) >> 4);
                            const uint32_t yl = desc_lo0 + ((st + 6 * TC_KBLOCK_BYTES) >> 4);
                            const int steps = min(4, nk1 - 4 * kb);
                            if (!TC_DIAG(1) && elect_one()) {
#pragma unroll 1
                                for (int part = 0; part < 3; ++part) {
                                    const uint32_t ad = (part == 1 ? xl : xh), bd = (part == 0 ? yl : yh);
#pragma unroll 1
                                    for (int ks = 0; ks < steps; ++ks)
                                        umma_ss(d_t, ad + 2 * ks, bd + 2 * ks, desc_hi, idesc_w, (uint32_t)((kb | part | ks) != 0));
                                }
                            }
                            __syncwarp();
                            if (elect_one()) {
                                if (p.pair) umma_commit_mc(&a_empty[sa], 3);
                                else umma_commit(&a_empty[sa]);
                            }
                            __syncwarp();
                            if (++sa == SA) {
                                sa = 0;
                                pha ^= 1;
                            }
                        }
                        if (elect_one()) {
                            umma_commit(&s_full[b0]);
                            umma_commit(&s_full[b0 + 1]);
                        }
                        __syncwarp();
                    }
                }
            } else if (CG2 && crank2 != 0) {
                // ---- peer CTA of a cta_group::2 pair: no MMA is issued here; warp 9 tells the leader when this CTA's
                // half of a column-tile image has landed ----
                if (par == 0) {
                    int sa = 0;
                    uint32_t pha = 0;
                    for (int t1 = 0; t1 < T; ++t1) {
                        mbar_wait(&a_full[sa], pha);
                        if (lane == 0) mbar_arrive_cluster(&a_peer[sa], 0);
                        __syncwarp();
                        if (++sa == SA) {
                            sa = 0;
                            pha ^= 1;
                        }
                    }
                }
            } else {
            mbar_wait(x_full, 0);
            int b1 = par % NB, sa = par % SA;
            uint32_t use1 = (uint32_t)((par / NB) & 1), pha = (uint32_t)((par / SA) & 1);
            TC_PROF_DECL
            for (int t1 = par; t1 < T; t1 += 2) {
                TC_PROF(7)
                // Three conditions, polled in this order:
                //   a_empty[sa]  MMA1 of the stage's previous tile t1 - SA has completed (the barrier the producer waits on
                //                before it loads tile t1 into the stage, so it costs nothing);
                //   p_free[b1]   MMA2 of tile t1 - NB has consumed P[b1];
                //   a_full[sa]   the image of tile t1 has landed (CG2: in both CTAs).
                // Why the first: every wait is a PARITY wait, ambiguous by two phases.  With an odd ring depth the two issue
                // warps share the A stages, so while the OTHER warp's tile t1 - SA is still in flight a_full[sa] reads
                // "complete" for tile t1.  p_free covers that only when NB <= SA; with NB = 4, SA = 3 (the C2 plan) a bulk copy
                // that lands ~5000 cycles after its successors would let MMA1 of tile t1 consume the half-landed image of
                // tile t1 - 3.  a_empty is exact (its previous phase belongs to this warp's own tile t1 - 2 SA, long complete).
                // Why in this order: try_wait suspends the thread up to a system time limit, so a barrier sampled EARLIER in
                // the same poll can be stale when a later one returns -- the guards have to be sampled before a_full.
                // Found with scripts/tc_protocol_model.py (tests/test_tc_protocol_model.py); never observed on the hardware.
                if constexpr (CG2) {
                    mbar_wait(&a_empty[sa], pha ^ 1);
                    mbar_wait3(&p_free[b1], use1 ^ 1, &a_full[sa], pha, &a_peer[sa], pha);
                } else {
                    mbar_wait3(&a_empty[sa], pha ^ 1, &p_free[b1], use1 ^ 1, &a_full[sa], pha);
                }
                TC_PROF(0)
                tc_fence_after();
                issue_mma1(b1, sa);
                if (elect_one()) {
                    commit_bar(&s_full[b1], false);
                    commit_bar(&a_empty[sa], true);
                }
                __syncwarp();
                TC_PROF(2)
                b1 += 2;
                if (b1 >= NB) {
                    b1 -= NB;
                    use1 ^= 1;
                }
                sa += 2;
                if (sa >= SA) {
                    sa -= SA;
                    pha ^= 1;
                }
            }
            if (par == 0) TC_PROF_FLUSH(0)
            }
        } else {
            // ---- MMA2 issuer (idle when the contraction runs on the CUDA cores) ----
            if constexpr (KV > 0) {
                // ---- V-ring producer: one record {V tile, column norms} per sub-tile ----
                if (lane == 0) {
                    const unsigned char* v_src = p.vimg + (size_t)t_begin * v_img_bytes;
                    int sv = 0;
                    uint32_t phv = 1;
                    for (int u = 0; u < T; ++u) {
                        mbar_wait(&v_empty[sv], phv);
                        mbar_arrive_expect_tx(&v_full[sv], v_img_bytes);
                        bulk_copy_g2s(v_ring + (size_t)sv * v_stage_bytes, v_src, v_img_bytes, &v_full[sv]);
                        v_src += v_img_bytes;
                        if (++sv == SV) {
                            sv = 0;
                            phv ^= 1;
                        }
                    }
                }
            } else if (CG2 && crank2 != 0) {
                // ---- peer CTA: relay "this CTA's half of the V image has landed" to the leader ----
                int sv = 0;
                uint32_t phv = 0;
                for (int u = 0; u < T; ++u) {
                    mbar_wait(&v_full[sv], phv);
                    if (lane == 0) mbar_arrive_cluster(&v_peer[sv], 0);
                    __syncwarp();
                    if (++sv == SV) {
                        sv = 0;
                        phv ^= 1;
                    }
                }
            } else if constexpr (DUAL) {
                // ---- two chunks per sub-tile: O[0] = P'[b] . V'_A, O[1] = P'[b] . V'_B; P'[b] is released after both ----
                int b2 = 0, sv = 0;
                uint32_t use2 = 0, phv = 0;
                for (int u = 0; u < T; ++u) {
                    const uint32_t opar = (uint32_t)(u & 1);  // each O buffer is used once per sub-tile
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        // P written (first chunk); V record landed; O[c] of the previous sub-tile has been drained
                        // bit 16 (sliced mode, early announcement): chunk A of sub-tile u is issued only after chunk B of
                        // sub-tile u - 1 has COMPLETED (the issuer itself waits for its completion barrier)
                        if (c == 0 && u > 0 && (p.dual_overlap & 16)) mbar_wait(&o_full[1], (uint32_t)((u - 1) & 1));
                        if (c == 0) mbar_wait3(&p_full[b2], use2, &v_full[sv], phv, &o_free[c], opar ^ 1);
                        else mbar_wait2(&v_full[sv], phv, &o_free[c], opar ^ 1);
                        tc_fence_after();
                        issue_mma2(b2, c, sv);
                        if (elect_one()) {
                            commit_bar(&v_empty[sv], true);
                            if (c == 1) commit_bar(&p_free[b2], false);
                            commit_bar(&o_full[c], false);
                        }
                        __syncwarp();
                        if (++sv == SV) {
                            sv = 0;
                            phv ^= 1;
                        }
                    }
                    if (++b2 == NB) {
                        b2 = 0;
                        use2 ^= 1;
                    }
                }
            } else {
            int b2 = 0, sv = 0;
            uint32_t use2 = 0, phv = 0;
            TC_PROF_DECL
            for (int u = 0; u < T; ++u) {
                const int ob = u % NOB;
                const uint32_t opar = (uint32_t)((u / NOB) & 1);  // use-count parity of O[ob]
                TC_PROF(7)
                // P written; V image landed; O[ob] of tile u - 2 has been drained (CG2: in both CTAs)
                if constexpr (CG2) mbar_wait(&v_peer[sv], phv);
                mbar_wait3(&p_full[b2], use2, &v_full[sv], phv, &o_free[ob], opar ^ 1);
                TC_PROF(3)
                tc_fence_after();
                issue_mma2(b2, ob, sv);
                if (elect_one()) {
                    commit_bar(&v_empty[sv], true);
                    commit_bar(&p_free[b2], false);
                    commit_bar(&o_full[ob], false);
                }
                __syncwarp();
                TC_PROF(6)
                if (++b2 == NB) {
                    b2 = 0;
                    use2 ^= 1;
                }
                if (++sv == SV) {
                    sv = 0;
                    phv ^= 1;
                }
            }
            TC_PROF_FLUSH(24)
            }
        }
    }
    } else {
        if constexpr (NWG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        else if constexpr (NWG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // =============================== epilogue warps ===============================
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int h = warp >> 2;  // column half of the sub-tile / of O handled by this warp
        const int row = q * 32 + lane;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        const int64_t grow = row0 + row;
        const bool live = grow < tc_npad(p.n);  // a pair's second CTA may lie past the last row block (odd count)
        const float nx = live ? reinterpret_cast<const float*>(p.rows + tc_norm_offset())[grow] : 0.0f;
        const float m2c = -2.0f * rh->inv_scale * ch->inv_scale;

        // ---- X tile -> TMEM (warpgroup 0: hi halves, warpgroup 1: lo halves) ----
        if (h < 2 && !WIDE) {
            const unsigned char* img = p.rows + tc_image_offset(p.n) + (size_t)(grow >> 6) * a_img_bytes +
                                       (size_t)h * KB * TC_KBLOCK_BYTES;
            const int r = (int)(grow & 63);
            for (int kb = 0; kb < KB; ++kb) {
                uint32_t w[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (live)
                        v = *reinterpret_cast<const uint4*>(img + (size_t)kb * TC_KBLOCK_BYTES + (size_t)r * 128 +
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
            if (lane == 0) arrive_pair(x_full);
        }

        // Two warpgroups ping-pong over the sub-tiles (warpgroup g owns tiles u = g, g+2, ...): each
        // thread owns one full row of S (64 entries), so the row scale needs no exchange and the two
        // warps that share an SM sub-partition work on different tiles, covering each other's
        // TMEM / MUFU latencies.
        const int g = h;
        // KP <= 64: a warpgroup drains the O buffers of its own tiles (all KP columns) and the two partial rows
        // are added at the end.  KP = 128 (k > 64): both warpgroups drain every tile, half of the columns each,
        // so a thread still carries 64 accumulators; the owner publishes the tile's scale through smem.
        constexpr bool SPLIT = KP > 64;
        constexpr bool FRAG = KV > 0 && !M12;
        // quarter-row epilogue (S re-read from TMEM in pieces): the four-warpgroup instantiations and the two-chunk kernels
        constexpr bool QUART = NWG == 4 || DUAL;
        static_assert(!QUART || (KV == 0 && !M12 && !WIDE), "quarter-row epilogue: X-resident MMA2 kernels, not Matern-1/2");
        constexpr int DW = SPLIT ? KP / 2 : KP;  // O columns one thread accumulates (per chunk)
        constexpr int NCH = DUAL ? 2 : 1;        // chunks of V contracted with the same P'
        uint64_t acc[NCH * DW / 2];              // fp32 pairs
#pragma unroll
        for (int c = 0; c < NCH * DW / 2; ++c) acc[c] = 0ull;

        // acc += O[ob] * dsc: one sub-tile's accumulator, un-scaled and added with round-to-nearest; CH (compile time):
        // which half of the accumulators (the chunk) it goes to
        auto drain = [&](auto CH, int ob, uint32_t par, float dsc, const float* dsc_ptr = nullptr) {
            constexpr int aoff = decltype(CH)::value * (DW / 2);
            mbar_wait(&o_full[ob], par);
            tc_fence_after();
            if (dsc_ptr) dsc = *dsc_ptr;
            const uint64_t d2 = pack2(dsc, dsc);
            constexpr int W = DW < 32 ? DW : 32;
            constexpr int NLD = DW / W;  // 1 or 2 loads in flight, one wait (a tcgen05.ld round trip is ~180 cycles)
            uint32_t o[NLD][W];
#pragma unroll
            for (int l = 0; l < NLD; ++l) tmem_ld_n<W>(tmem + lane_bits + col_o + ob * KP + (SPLIT ? g * DW : 0) + l * W, o[l]);
            tmem_wait_ld();
            if (!TC_DIAG(4))
#pragma unroll
            for (int l = 0; l < NLD; ++l)
#pragma unroll
                for (int e = 0; e < W / 2; ++e)
                    acc[aoff + l * (W / 2) + e] = fma2(pack2(__uint_as_float(o[l][2 * e]), __uint_as_float(o[l][2 * e + 1])), d2,
                                                       acc[aoff + l * (W / 2) + e]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_pair(&o_free[ob]);
        };
        constexpr std::integral_constant<int, 0> CH0{};
        constexpr std::integral_constant<int, NCH - 1> CH1{};
        float* dsc_sm = xchg;  // [8][128]([2] DUAL): per-row un-scale factors of the last 8 tiles (SPLIT mode)
        int next_drain = 0;    // SPLIT mode: next tile this warpgroup has to drain
        auto drain_through = [&](int last) {  // SPLIT mode: drain tiles next_drain .. last (inclusive)
            for (; next_drain <= last; ++next_drain) {
                const int t = next_drain;
                // the owner's scale is read inside drain(), after its wait on o_full (published before p_full ->
                // MMA2 -> o_full): one barrier poll per drain, not two (a successful poll costs ~100 cycles)
                if constexpr (DUAL) {  // O[0] / O[1] are the two chunks of tile t, each used once per tile
                    drain(CH0, 0, (uint32_t)(t & 1), -1.0f, &dsc_sm[((t & 7) * TC_BM + row) * 2]);
                    drain(CH1, 1, (uint32_t)(t & 1), -1.0f, &dsc_sm[((t & 7) * TC_BM + row) * 2 + 1]);
                } else {
                    drain(CH0, t % NOB, (uint32_t)((t / NOB) & 1), -1.0f, &dsc_sm[(t & 7) * TC_BM + row]);
                }
            }
        };

        // per-kernel constants of pass 1: z = S * za + (|y|^2 * zc + zx)
        const int kid = KIDT >= 0 ? KIDT : p.kid;
        float zc;
        if (kid == KID_RBF) zc = -0.5f * TC_LOG2E;
        else if (kid == KID_MATERN32) zc = 3.0f;
        else if (kid == KID_MATERN52) zc = 5.0f;
        else zc = 1.0f;
        const uint64_t za2 = pack2(m2c * zc, m2c * zc), zc2 = pack2(zc, zc), zx2 = pack2(nx * zc, nx * zc);
        const bool is_rbf = kid == KID_RBF;

        int b = g % NB, sv = g % SV;  // NWG <= NB, SV
        uint32_t use = (uint32_t)((g / NB) & 1), phv = (uint32_t)((g / SV) & 1);
        float dsc_prev = 0.0f;
        // register-contraction mode: level-1 / level-2 sums per column of V, as (even j, odd j) pairs
        uint64_t accv[KV ? KV : 1], accv2[KV ? KV : 1];
#pragma unroll
        for (int c = 0; c < (KV ? KV : 1); ++c) accv[c] = accv2[c] = 0ull;
        int lvl = 0;
        // FRAG (register contraction, all kernels but Matern-1/2): S is read in the 16x256b fragment pattern, so a
        // thread holds 4 rows x 16 columns of the sub-tile instead of 1 row x 64 columns and loads the norms and V
        // values of 16 columns only -- with one row per thread the broadcast loads (512 B per thread and sub-tile)
        // kept the shared-memory pipe 79 % busy and bounded the kernel.  Row r = 2 hf + s of a thread is TMEM lane
        // 32 q + lane / 4 + 8 s + 16 hf; its columns are 8 i + 2 (lane % 4) + {0, 1}, i = 0..7.
        const int tq = lane & 3;
        uint64_t zxr2[4];
        float accf[4][KV ? KV : 1], accf2[4][KV ? KV : 1];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t gr = row0 + q * 32 + (lane >> 2) + 8 * (r & 1) + 16 * (r >> 1);
            const float nxr = (FRAG && gr < tc_npad(p.n)) ? reinterpret_cast<const float*>(p.rows + tc_norm_offset())[gr] : 0.0f;
            zxr2[r] = pack2(nxr * zc, nxr * zc);
#pragma unroll
            for (int c = 0; c < (KV ? KV : 1); ++c) accf[r][c] = accf2[r][c] = 0.0f;
        }
        bool sliced = false;
        if constexpr (DUAL) {
            if (p.dual_mode == 2) {
                // ---- sliced two-chunk epilogue (three S/P buffers) ----
                // With the drains behind the pointwise stage (dual_mode 1) MMA2 of sub-tile t + 1 waits for the owner of
                // sub-tile t + 2, which sits in its ~1800-cycle pointwise stage and cannot drain O of sub-tile t.  Here both
                // warpgroups walk ALL sub-tiles, drain chunk A, then chunk B of each as soon as it completes, and fill the two
                // gaps with a quarter of the pointwise stage of the warpgroup's NEXT own sub-tile: tile u is prepared in the
                // slots of tiles u - 2 (row extreme + quarter 0, quarter 1) and u - 1 (quarters 2, 3, publish).  The column
                // norms and V scales of tile u come from global memory (the V ring only feeds MMA2): prefetched one own
                // tile ahead into a register, handed to the warpgroup through a double-buffered smem row.
                sliced = true;
                float* wny = xchg + 16 * TC_BM + g * 256;
                const unsigned char* rec_a = p.vimg + ((size_t)kc * p.sub_tiles + t_begin) * v_img_bytes;
                const unsigned char* rec_b = p.vimg + ((size_t)kc_b * p.sub_tiles + t_begin) * v_img_bytes;
                float pre = 0.0f;
                auto prefetch = [&](int u) {
                    if (u < T && row < 66) {
                        const unsigned char* ra = rec_a + (size_t)u * v_img_bytes;
                        if (row < 64) pre = __ldg(reinterpret_cast<const float*>(ra + v_norm_off) + row);
                        else if (row == 64) pre = __ldg(reinterpret_cast<const float*>(ra + KP * 256));
                        else pre = __ldg(reinterpret_cast<const float*>(rec_b + (size_t)u * v_img_bytes + KP * 256));
                    }
                };
                const uint64_t nl2 = pack2(-TC_LOG2E, -TC_LOG2E), one2 = pack2(1.0f, 1.0f),
                               third2 = pack2(1.0f / 3.0f, 1.0f / 3.0f);
                uint64_t E2 = 0ull;
                int Ecur = 0;
                int pending_pfull = -1;  // diagnostic (bit 8): p_full arrive deferred to the next slot
                // 16 entries of S -> P'_hi (8 words) | P'_lo (8 words), written in place
                auto quarter = [&](const uint32_t* s16, const float4* ny4p, uint32_t t_dst) {
                    uint32_t pq[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 ny4 = ny4p[i];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint64_t zz = fma2(pack2(__uint_as_float(s16[4 * i + 2 * e]), __uint_as_float(s16[4 * i + 2 * e + 1])),
                                                     za2, fma2(e == 0 ? pack2(ny4.x, ny4.y) : pack2(ny4.z, ny4.w), zc2, zx2));
                            float p0, p1;
                            if (is_rbf) {
                                float a0, a1;
                                unpack2(add2(zz, E2), a0, a1);
                                p0 = ex2_approx(a0);
                                p1 = ex2_approx(a1);
                            } else {
                                float z0, z1;
                                unpack2(zz, z0, z1);
                                const uint64_t r2 = pack2(sqrt_approx(fmaxf(z0, 0.0f)), sqrt_approx(fmaxf(z1, 0.0f)));
                                float a0, a1;
                                unpack2(fma2(r2, nl2, E2), a0, a1);
                                uint64_t pp = pack2(ex2_approx(a0), ex2_approx(a1));
                                if (kid == KID_MATERN32) pp = mul2(pp, add2(r2, one2));
                                else pp = mul2(pp, fma2(r2, fma2(r2, third2, one2), one2));
                                unpack2(pp, p0, p1);
                            }
                            const __half2 h2 = __floats2half2_rn(p0, p1);
                            const float2 f2 = __half22float2(h2);
                            float l0, l1;
                            unpack2(sub2(pack2(p0, p1), pack2(f2.x, f2.y)), l0, l1);
                            const __half2 l2 = __floats2half2_rn(l0, l1);
                            pq[2 * i + e] = *reinterpret_cast<const uint32_t*>(&h2);
                            pq[8 + 2 * i + e] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                    }
                    tmem_st16(t_dst, pq);
                };
                auto pw_slice = [&](int u, auto SL) {
                    constexpr int sl = decltype(SL)::value;
                    if (u < 0 || u >= T) return;
                    const int bu = u % NB;
                    const uint32_t t_s = tmem + lane_bits + col_sp + bu * 64;
                    float* nys = wny + ((u >> 1) & 1) * 128;
                    const bool bypass = (p.dual_overlap & 2) != 0;  // diagnostic: norms straight from global memory
                    const float* ny = bypass ? reinterpret_cast<const float*>(rec_a + (size_t)u * v_img_bytes + v_norm_off) : nys;
                    const float4* nyq = reinterpret_cast<const float4*>(ny);
                    if constexpr (sl == 0) {
                        if (row < 66) nys[row] = pre;
                        asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
                        prefetch(u + 2);
                        mbar_wait(&s_full[bu], (uint32_t)((u / NB) & 1));
                        tc_fence_after();
                        uint32_t s0[32], s1[32];
                        tmem_ld32(t_s, s0);
                        tmem_ld32(t_s + 32, s1);
                        tmem_wait_ld();
                        float ext = is_rbf ? -3.0e38f : 3.0e38f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 ny4 = nyq[i];
                            const uint32_t* sh = i < 8 ? &s0[4 * i] : &s1[4 * (i - 8)];
                            float z0, z1, z2, z3;
                            unpack2(fma2(pack2(__uint_as_float(sh[0]), __uint_as_float(sh[1])), za2,
                                         fma2(pack2(ny4.x, ny4.y), zc2, zx2)), z0, z1);
                            unpack2(fma2(pack2(__uint_as_float(sh[2]), __uint_as_float(sh[3])), za2,
                                         fma2(pack2(ny4.z, ny4.w), zc2, zx2)), z2, z3);
                            if (is_rbf) ext = fmaxf(max3(ext, z0, z1), fmaxf(z2, z3));
                            else ext = fminf(min3(ext, z0, z1), fminf(z2, z3));
                        }
                        int E;
                        if (is_rbf) {
                            E = 14 - __float2int_rd(fmaxf(ext, -200.0f));
                        } else {
                            float pmax;
                            if (kid == KID_MATERN32) pmax = tc_value<KID_MATERN32>(ext);
                            else pmax = tc_value<KID_MATERN52>(ext);
                            E = 14 + 127 - (int)((__float_as_uint(pmax) >> 23) & 0xFF);
                        }
                        E = max(0, min(E, 120));
                        Ecur = E;
                        const float Ef = (float)E;
                        E2 = pack2(Ef, Ef);
                        quarter(s0, nyq, t_s);
                    } else {
                        uint32_t sq[16];
                        tmem_ld16(t_s + sl * 16, sq);
                        tmem_wait_ld();
                        quarter(sq, nyq + sl * 4, t_s + sl * 16);
                    }
                    if constexpr (sl == 3) {
                        tmem_wait_st();
                        tc_fence_before();
                        const float e2 = __uint_as_float((uint32_t)(127 - Ecur) << 23);
                        dsc_sm[((u & 7) * TC_BM + row) * 2] = e2 * nys[64];  // published before p_full -> MMA2 -> o_full
                        dsc_sm[((u & 7) * TC_BM + row) * 2 + 1] = e2 * nys[65];
                        __syncwarp();
                        if (p.dual_overlap & 8) pending_pfull = bu;
                        else if (lane == 0) arrive_pair(&p_full[bu]);
                    }
                };
                constexpr std::integral_constant<int, 0> S0{};
                constexpr std::integral_constant<int, 1> S1{};
                constexpr std::integral_constant<int, 2> S2{};
                constexpr std::integral_constant<int, 3> S3{};
                prefetch(g);
                // t = -2, -1: the slots of the sub-tiles "before the first" prepare tiles 0 and 1 (nothing to drain yet)
                // one slot: drain chunk CH of sub-tile t, then work item WK of sub-tile u:
                //   0: row extreme + quarter 0    1: quarters 1 and 2    2: quarter 3 + publish (p_full)    3: nothing
                // P'(u) must be complete while MMA2 of chunk B of sub-tile u - 1 still runs, or the tensor pipe idles for a
                // whole slot: so the last item sits in slot A of sub-tile u - 1 and slot B of a foreign sub-tile only drains.
                // Items 1 and 2 overlap the drain: the quarter's S load and the first half of O are in flight together, the
                // second half of O loads while the quarter runs through the special-function pipe.
                auto slot = [&](auto CH, int t, int u, auto WK) {
                    constexpr int wk = decltype(WK)::value;
                    constexpr int ch = decltype(CH)::value;
                    const bool have_pw = wk != 3 && u >= 0 && u < T;
                    if (pending_pfull >= 0) {  // diagnostic: P'(u) is announced only once chunk B of sub-tile u - 1 has completed
                        if (t >= 0) mbar_wait(&o_full[ch], (uint32_t)(t & 1));
                        __syncwarp();
                        if (lane == 0) arrive_pair(&p_full[pending_pfull]);
                        pending_pfull = -1;
                    }
                    if (wk == 0 || wk >= 3 || t < 0 || !have_pw || (p.dual_overlap & 1) == 0) {
                        if (t >= 0) drain(CH, ch, (uint32_t)(t & 1), -1.0f, &dsc_sm[((t & 7) * TC_BM + row) * 2 + ch]);
                        if (have_pw) {
                            if constexpr (wk == 0) pw_slice(u, S0);
                            if constexpr (wk == 1) {
                                pw_slice(u, S1);
                                pw_slice(u, S2);
                            }
                            if constexpr (wk == 2) pw_slice(u, S3);
                            if constexpr (wk == 4) pw_slice(u, S1);
                            if constexpr (wk == 5) pw_slice(u, S2);
                        }
                        return;
                    }
                    constexpr int sl = wk == 1 ? 1 : 3;
                    const int bu = u % NB;
                    const uint32_t t_s = tmem + lane_bits + col_sp + bu * 64;
                    const float* nys = wny + ((u >> 1) & 1) * 128;
                    const float* ny = (p.dual_overlap & 2) ? reinterpret_cast<const float*>(rec_a + (size_t)u * v_img_bytes + v_norm_off) : nys;
                    uint32_t sq[16];
                    tmem_ld16(t_s + sl * 16, sq);
                    mbar_wait(&o_full[ch], (uint32_t)(t & 1));
                    tc_fence_after();
                    const float dsc = dsc_sm[((t & 7) * TC_BM + row) * 2 + ch];
                    const uint64_t d2 = pack2(dsc, dsc);
                    const uint32_t t_o = tmem + lane_bits + col_o + ch * KP + g * DW;
                    constexpr int aoff = ch * (DW / 2);
                    uint32_t o1[32], o2[32];
                    tmem_ld32(t_o, o1);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        acc[aoff + e] = fma2(pack2(__uint_as_float(o1[2 * e]), __uint_as_float(o1[2 * e + 1])), d2, acc[aoff + e]);
                    tmem_ld32(t_o + 32, o2);
                    quarter(sq, reinterpret_cast<const float4*>(ny) + sl * 4, t_s + sl * 16);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        acc[aoff + 16 + e] = fma2(pack2(__uint_as_float(o2[2 * e]), __uint_as_float(o2[2 * e + 1])), d2, acc[aoff + 16 + e]);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_pair(&o_free[ch]);
                    if constexpr (wk == 1) pw_slice(u, S2);
                    if constexpr (wk == 2) {
                        tmem_wait_st();
                        tc_fence_before();
                        const float e2 = __uint_as_float((uint32_t)(127 - Ecur) << 23);
                        dsc_sm[((u & 7) * TC_BM + row) * 2] = e2 * nys[64];
                        dsc_sm[((u & 7) * TC_BM + row) * 2 + 1] = e2 * nys[65];
                        __syncwarp();
                        if (p.dual_overlap & 8) pending_pfull = bu;
                        else if (lane == 0) arrive_pair(&p_full[bu]);
                    }
                };
                constexpr std::integral_constant<int, 4> S4{};
                constexpr std::integral_constant<int, 5> S5{};
                for (int t = -2; t < T; ++t) {
                    if (p.dual_overlap & 4) {  // diagnostic: one quarter per slot, publish in slot B of sub-tile u - 1
                        if ((t & 1) == g) {
                            slot(CH0, t, t + 2, S0);
                            slot(CH1, t, t + 2, S4);
                        } else {
                            slot(CH0, t, t + 1, S5);
                            slot(CH1, t, t + 1, S2);
                        }
                    } else if ((t & 1) == g) {
                        slot(CH0, t, t + 2, S0);
                        slot(CH1, t, t + 2, S1);
                    } else {
                        slot(CH0, t, t + 1, S2);
                        slot(CH1, t, t + 1, S3);
                    }
                }
                next_drain = T;
            }
        }
        TC_PROF_DECL
        for (int u = sliced ? T : g; u < T; u += NWG) {
            int sv_b = 0;  // DUAL: the V-ring stages of sub-tile u are 2 u (first chunk) and 2 u + 1 (second chunk)
            uint32_t phv_b = 0;
            if constexpr (DUAL) {
                sv = (2 * u) % SV;
                phv = (uint32_t)(((2 * u) / SV) & 1);
                sv_b = (2 * u + 1) % SV;
                phv_b = (uint32_t)(((2 * u + 1) / SV) & 1);
            }
            const unsigned char* vst = v_ring + (size_t)sv * v_stage_bytes;
            TC_PROF(7)
            // |y|^2 and the V scale of this sub-tile are in smem; MMA1 has written S[b]
            if constexpr (DUAL) mbar_wait3(&v_full[sv], phv, &v_full[sv_b], phv_b, &s_full[b], use);
            else mbar_wait2(&v_full[sv], phv, &s_full[b], use);
            TC_PROF(1)
            tc_fence_after();
            const uint32_t t_s = tmem + lane_bits + col_sp + b * 64;
            if constexpr (QUART) {
                // ---- four epilogue warpgroups, 104 registers per thread: the row is never held whole.  Pass 1 reads S in
                // two halves for the row extreme; pass 2 re-reads it in quarters (the load of quarter q + 1 is in flight
                // while quarter q runs through the special-function pipe), recomputes z and writes P' in place:
                // quarter q (S columns 16 q .. 16 q + 15) becomes P'_hi in columns 16 q .. 16 q + 7 and P'_lo in columns
                // 16 q + 8 .. 16 q + 15 -- nothing unread is overwritten (issue_mma2 addresses A with that stride) ----
                const float vinv = *reinterpret_cast<const float*>(vst + KP * 256);
                const float4* nyq = reinterpret_cast<const float4*>(vst + v_norm_off);
                float ext = is_rbf ? -3.0e38f : 3.0e38f;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t sh[32];
                    tmem_ld32(t_s + hh * 32, sh);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 ny4 = nyq[hh * 8 + i];
                        float z0, z1, z2, z3;
                        unpack2(fma2(pack2(__uint_as_float(sh[4 * i]), __uint_as_float(sh[4 * i + 1])), za2,
                                     fma2(pack2(ny4.x, ny4.y), zc2, zx2)), z0, z1);
                        unpack2(fma2(pack2(__uint_as_float(sh[4 * i + 2]), __uint_as_float(sh[4 * i + 3])), za2,
                                     fma2(pack2(ny4.z, ny4.w), zc2, zx2)), z2, z3);
                        if (is_rbf) ext = fmaxf(max3(ext, z0, z1), fmaxf(z2, z3));
                        else ext = fminf(min3(ext, z0, z1), fminf(z2, z3));
                    }
                }
                int E;
                if (is_rbf) {
                    E = 14 - __float2int_rd(fmaxf(ext, -200.0f));
                } else {
                    float pmax;
                    if (kid == KID_MATERN32) pmax = tc_value<KID_MATERN32>(ext);
                    else pmax = tc_value<KID_MATERN52>(ext);
                    E = 14 + 127 - (int)((__float_as_uint(pmax) >> 23) & 0xFF);
                }
                E = max(0, min(E, 120));
                const float Ef = (float)E;
                const float dsc = __uint_as_float((uint32_t)(127 - E) << 23) * vinv;
                const uint64_t E2 = pack2(Ef, Ef);
                const uint64_t nl2 = pack2(-TC_LOG2E, -TC_LOG2E), one2 = pack2(1.0f, 1.0f),
                               third2 = pack2(1.0f / 3.0f, 1.0f / 3.0f);
                uint32_t sq[2][16];
                tmem_ld16(t_s, sq[0]);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    tmem_wait_ld();  // quarter q4 is in registers
                    if (q4 < 3) tmem_ld16(t_s + (q4 + 1) * 16, sq[(q4 + 1) & 1]);
                    uint32_t pq[16];  // [0, 8): P'_hi pairs, [8, 16): P'_lo pairs
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 ny4 = nyq[q4 * 4 + i];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const uint64_t zz = fma2(pack2(__uint_as_float(sq[q4 & 1][4 * i + 2 * e]), __uint_as_float(sq[q4 & 1][4 * i + 2 * e + 1])),
                                                     za2, fma2(e == 0 ? pack2(ny4.x, ny4.y) : pack2(ny4.z, ny4.w), zc2, zx2));
                            float p0, p1;
                            if (is_rbf) {
                                float a0, a1;
                                unpack2(add2(zz, E2), a0, a1);
                                p0 = ex2_approx(a0);
                                p1 = ex2_approx(a1);
                            } else {
                                float z0, z1;
                                unpack2(zz, z0, z1);
                                const uint64_t r2 = pack2(sqrt_approx(fmaxf(z0, 0.0f)), sqrt_approx(fmaxf(z1, 0.0f)));
                                float a0, a1;
                                unpack2(fma2(r2, nl2, E2), a0, a1);
                                uint64_t pp = pack2(ex2_approx(a0), ex2_approx(a1));
                                if (kid == KID_MATERN32) pp = mul2(pp, add2(r2, one2));
                                else pp = mul2(pp, fma2(r2, fma2(r2, third2, one2), one2));
                                unpack2(pp, p0, p1);
                            }
                            const __half2 h2 = __floats2half2_rn(p0, p1);
                            const float2 f2 = __half22float2(h2);
                            float l0, l1;
                            unpack2(sub2(pack2(p0, p1), pack2(f2.x, f2.y)), l0, l1);
                            const __half2 l2 = __floats2half2_rn(l0, l1);
                            pq[2 * i + e] = *reinterpret_cast<const uint32_t*>(&h2);
                            pq[8 + 2 * i + e] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                    }
                    tmem_st16(t_s + q4 * 16, pq);
                }
                tmem_wait_st();
                tc_fence_before();
                if constexpr (DUAL) {  // per-chunk un-scale factors (the V images of the two chunks carry their own scales)
                    const float vinv_b = *reinterpret_cast<const float*>(v_ring + (size_t)sv_b * v_stage_bytes + KP * 256);
                    const float e2 = __uint_as_float((uint32_t)(127 - E) << 23);
                    dsc_sm[((u & 7) * TC_BM + row) * 2] = dsc;  // published before p_full -> MMA2 -> o_full
                    dsc_sm[((u & 7) * TC_BM + row) * 2 + 1] = e2 * vinv_b;
                }
                __syncwarp();
                if (lane == 0) arrive_pair(&p_full[b]);
                if constexpr (DUAL) drain_through(u - 1);
                else if (u >= NWG) drain(CH0, g, (uint32_t)(((u - NWG) / NWG) & 1), dsc_prev);
                dsc_prev = dsc;
            } else {
            uint32_t s0[32], s1[32];
            if constexpr (FRAG) {
                tmem_ld_16x256b_x8(t_s, s0);                // lanes 32q + [0, 16)
                tmem_ld_16x256b_x8(t_s + (16u << 16), s1);  // lanes 32q + [16, 32)
            } else {
                tmem_ld32(t_s, s0);
                tmem_ld32(t_s + 32, s1);
            }
            tmem_wait_ld();
            TC_PROF(2)
            const float4* nyv = reinterpret_cast<const float4*>(vst + v_norm_off);
            if constexpr (FRAG) {
                // ---- register contraction, fragment layout: 4 rows x 16 columns per thread ----
                tc_fence_before();  // S[b] is in registers: MMA1 may overwrite the buffer
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_free[b]);
                const float2* ny2 = reinterpret_cast<const float2*>(vst + v_norm_off);
                const float2* v2 = reinterpret_cast<const float2*>(vst);  // raw fp32 V tile, [KV][64]
                uint64_t ts[4][KV];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < KV; ++c) ts[r][c] = 0ull;
                const uint64_t nl2 = pack2(-TC_LOG2E, -TC_LOG2E), one2 = pack2(1.0f, 1.0f),
                               third2 = pack2(1.0f / 3.0f, 1.0f / 3.0f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 nyp = ny2[4 * i + tq];
                    const uint64_t nyc = mul2(pack2(nyp.x, nyp.y), zc2);
                    uint64_t vp[KV];
#pragma unroll
                    for (int c = 0; c < KV; ++c) {
                        const float2 vv = v2[c * 32 + 4 * i + tq];
                        vp[c] = pack2(vv.x, vv.y);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int o = 4 * i + 2 * (r & 1);
                        const uint64_t sp = (r < 2) ? pack2(__uint_as_float(s0[o]), __uint_as_float(s0[o + 1]))
                                                    : pack2(__uint_as_float(s1[o]), __uint_as_float(s1[o + 1]));
                        float z0, z1;
                        unpack2(fma2(sp, za2, add2(nyc, zxr2[r])), z0, z1);
                        uint64_t pp;
                        if (is_rbf) {
                            pp = pack2(ex2_approx(z0), ex2_approx(z1));
                        } else {
                            const uint64_t r2 = pack2(sqrt_approx(fmaxf(z0, 0.0f)), sqrt_approx(fmaxf(z1, 0.0f)));
                            float a0, a1;
                            unpack2(mul2(r2, nl2), a0, a1);
                            pp = pack2(ex2_approx(a0), ex2_approx(a1));
                            if (kid == KID_MATERN32) pp = mul2(pp, add2(r2, one2));
                            else if (kid == KID_MATERN52) pp = mul2(pp, fma2(r2, fma2(r2, third2, one2), one2));
                        }
#pragma unroll
                        for (int c = 0; c < KV; ++c) ts[r][c] = fma2(pp, vp[c], ts[r][c]);
                    }
                }
                __syncwarp();  // every lane has read the stage (norms and V)
                if (lane == 0) mbar_arrive(&v_empty[sv]);
                // three-level sum: 8 products per partial, one add per tile, one add per 32 tiles
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < KV; ++c) {
                        float lo, hi;
                        unpack2(ts[r][c], lo, hi);
                        accf[r][c] += lo + hi;
                    }
                if (++lvl == 32) {
                    lvl = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < KV; ++c) {
                            accf2[r][c] += accf[r][c];
                            accf[r][c] = 0.0f;
                        }
                }
            } else if constexpr (KV > 0) {
                // ---- register contraction: Y[row, c] += sum_j f(D_j) V[j, c], c < KV ----
                tc_fence_before();  // S[b] is in registers: MMA1 may overwrite the buffer
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_free[b]);
                uint64_t z[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 ny4 = nyv[i];
                    z[2 * i] = fma2(pack2(__uint_as_float(s0[4 * i]), __uint_as_float(s0[4 * i + 1])), za2,
                                    fma2(pack2(ny4.x, ny4.y), zc2, zx2));
                    z[2 * i + 1] = fma2(pack2(__uint_as_float(s0[4 * i + 2]), __uint_as_float(s0[4 * i + 3])), za2,
                                        fma2(pack2(ny4.z, ny4.w), zc2, zx2));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 ny4 = nyv[8 + i];
                    z[16 + 2 * i] = fma2(pack2(__uint_as_float(s1[4 * i]), __uint_as_float(s1[4 * i + 1])), za2,
                                         fma2(pack2(ny4.x, ny4.y), zc2, zx2));
                    z[16 + 2 * i + 1] = fma2(pack2(__uint_as_float(s1[4 * i + 2]), __uint_as_float(s1[4 * i + 3])), za2,
                                             fma2(pack2(ny4.z, ny4.w), zc2, zx2));
                }
                if constexpr (M12) {
                    // near-coincident pairs: same recompute as the tensor-core contraction path below
                    float ext = 3.0e38f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float z0, z1;
                        unpack2(z[i], z0, z1);
                        ext = min3(ext, z0, z1);
                    }
                    if (live && ext < 1.0e-3f * nx * nx) {
                        const int64_t j0 = (t_begin + u) * TC_BN;
                        const unsigned char* xi = p.rows + tc_image_offset(p.n) + (size_t)(grow >> 6) * a_img_bytes +
                                                  (size_t)(grow & 63) * 128;
                        const float* nyf = reinterpret_cast<const float*>(vst + v_norm_off);
                        float zl[TC_BN];
#pragma unroll
                        for (int i = 0; i < 32; ++i) unpack2(z[i], zl[2 * i], zl[2 * i + 1]);
#pragma unroll 1
                        for (int j = 0; j < TC_BN; ++j) {
                            const float s2 = nx + nyf[j];
                            if (zl[j] < 2.25e-4f * s2 * s2 && j0 + j < p.m) {
                                const int64_t jg = j0 + j;
                                const unsigned char* yj = col_images + (size_t)(jg >> 6) * a_img_bytes + (size_t)(jg & 63) * 128;
                                zl[j] = tc_exact_dist2(xi, (int)(grow & 7), yj, (int)(jg & 7), KB, rh->inv_scale, ch->inv_scale);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) z[i] = pack2(zl[2 * i], zl[2 * i + 1]);
                    }
                }
                const float4* v4 = reinterpret_cast<const float4*>(vst);  // raw fp32 V tile, [KV][64]
                uint64_t ts[KV];  // this tile's sums: (even j, odd j) partial pairs per column of V
#pragma unroll
                for (int c = 0; c < KV; ++c) ts[c] = 0ull;
                const uint64_t nl2 = pack2(-TC_LOG2E, -TC_LOG2E), one2 = pack2(1.0f, 1.0f),
                               third2 = pack2(1.0f / 3.0f, 1.0f / 3.0f);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    uint64_t pp[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float z0, z1;
                        unpack2(z[2 * i + e], z0, z1);
                        if (is_rbf) {
                            pp[e] = pack2(ex2_approx(z0), ex2_approx(z1));
                        } else {
                            const uint64_t r2 = pack2(sqrt_approx(fmaxf(z0, 0.0f)), sqrt_approx(fmaxf(z1, 0.0f)));
                            float a0, a1;
                            unpack2(mul2(r2, nl2), a0, a1);
                            pp[e] = pack2(ex2_approx(a0), ex2_approx(a1));
                            if (kid == KID_MATERN32) pp[e] = mul2(pp[e], add2(r2, one2));
                            else if (kid == KID_MATERN52) pp[e] = mul2(pp[e], fma2(r2, fma2(r2, third2, one2), one2));
                        }
                    }
#pragma unroll
                    for (int c = 0; c < KV; ++c) {
                        const float4 vv = v4[c * 16 + i];
                        ts[c] = fma2(pp[0], pack2(vv.x, vv.y), ts[c]);
                        ts[c] = fma2(pp[1], pack2(vv.z, vv.w), ts[c]);
                    }
                }
                __syncwarp();  // every lane has read the stage (norms and V)
                if (lane == 0) mbar_arrive(&v_empty[sv]);
                // three-level sum: 32 products per partial, one add per tile, one add per 32 tiles
#pragma unroll
                for (int c = 0; c < KV; ++c) accv[c] = add2(accv[c], ts[c]);
                if (++lvl == 32) {
                    lvl = 0;
#pragma unroll
                    for (int c = 0; c < KV; ++c) {
                        accv2[c] = add2(accv2[c], accv[c]);
                        accv[c] = 0ull;
                    }
                }
            } else {
            const float vinv = *reinterpret_cast<const float*>(vst + KP * 256);
            if (TC_DIAG(2)) {
                uint32_t phi[16], plo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) phi[i] = s0[i] + s1[i], plo[i] = s0[16 + i] + s1[16 + i];
                tmem_st16(t_s, phi);
                tmem_st16(t_s + 16, plo);
                tmem_st16(t_s + 32, plo);
                tmem_st16(t_s + 48, phi);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (SPLIT) dsc_sm[(u & 7) * TC_BM + row] = 1.0f;
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
                TC_PROF(4)
                if (SPLIT) drain_through(u - 1);
                else if (u >= NWG) drain(CH0, g, (uint32_t)(((u - NWG) / NWG) & 1), 1.0f);
                TC_PROF(5)
            } else {
            // ---- pass 1: z_j and the row extreme ----
            uint64_t z[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 ny4 = nyv[i];
                z[2 * i] = fma2(pack2(__uint_as_float(s0[4 * i]), __uint_as_float(s0[4 * i + 1])), za2,
                                fma2(pack2(ny4.x, ny4.y), zc2, zx2));
                z[2 * i + 1] = fma2(pack2(__uint_as_float(s0[4 * i + 2]), __uint_as_float(s0[4 * i + 3])), za2,
                                    fma2(pack2(ny4.z, ny4.w), zc2, zx2));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 ny4 = nyv[8 + i];
                z[16 + 2 * i] = fma2(pack2(__uint_as_float(s1[4 * i]), __uint_as_float(s1[4 * i + 1])), za2,
                                     fma2(pack2(ny4.x, ny4.y), zc2, zx2));
                z[16 + 2 * i + 1] = fma2(pack2(__uint_as_float(s1[4 * i + 2]), __uint_as_float(s1[4 * i + 3])), za2,
                                         fma2(pack2(ny4.z, ny4.w), zc2, zx2));
            }
            float ext = is_rbf ? -3.0e38f : 3.0e38f;
            if (is_rbf) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float z0, z1;
                    unpack2(z[i], z0, z1);
                    ext = max3(ext, z0, z1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float z0, z1;
                    unpack2(z[i], z0, z1);
                    ext = min3(ext, z0, z1);
                }
                // Matern-1/2 is not smooth at r = 0: exp(-sqrt(D)) turns the absolute error eps (|x|^2 + |y|^2) of the
                // GEMM-form D into eps (|x|^2 + |y|^2) / (2 r).  Pairs with D < 2.25e-4 (|x|^2 + |y|^2)^2 (error above
                // ~1e-5) are recomputed from direct differences of the packed points; rows without such a pair
                // (the usual case: only the diagonal of K(X, X) has them) pay one compare per tile.
                if constexpr (M12) if (live && ext < 1.0e-3f * nx * nx) {
                    const int64_t j0 = (t_begin + u) * TC_BN;
                    const unsigned char* xi = p.rows + tc_image_offset(p.n) + (size_t)(grow >> 6) * a_img_bytes +
                                              (size_t)(grow & 63) * 128;
                    const float* nyf = reinterpret_cast<const float*>(vst + v_norm_off);
                    // slow path: z goes through a thread-local array so that the scan can use a run-time index
                    // (the unrolled copies keep z itself in registers on the fast path)
                    float zl[TC_BN];
#pragma unroll
                    for (int i = 0; i < 32; ++i) unpack2(z[i], zl[2 * i], zl[2 * i + 1]);
#pragma unroll 1
                    for (int j = 0; j < TC_BN; ++j) {
                        const float s2 = nx + nyf[j];
                        if (zl[j] < 2.25e-4f * s2 * s2 && j0 + j < p.m) {
                            const int64_t jg = j0 + j;
                            const unsigned char* yj = col_images + (size_t)(jg >> 6) * a_img_bytes + (size_t)(jg & 63) * 128;
                            zl[j] = tc_exact_dist2(xi, (int)(grow & 7), yj, (int)(jg & 7), KB, rh->inv_scale, ch->inv_scale);
                        }
                    }
                    ext = 3.0e38f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        z[i] = pack2(zl[2 * i], zl[2 * i + 1]);
                        ext = min3(ext, zl[2 * i], zl[2 * i + 1]);
                    }
                }
            }
            // P' = P * 2^E with max_j P' in [2^14, 2^15): the fp16 hi/lo pair keeps 22 bits of the row's large entries
            int E;
            if (is_rbf) {
                E = 14 - __float2int_rd(fmaxf(ext, -200.0f));
            } else {
                float pmax;
                if (kid == KID_MATERN12) pmax = tc_value<KID_MATERN12>(ext);
                else if (kid == KID_MATERN32) pmax = tc_value<KID_MATERN32>(ext);
                else pmax = tc_value<KID_MATERN52>(ext);
                E = 14 + 127 - (int)((__float_as_uint(pmax) >> 23) & 0xFF);
            }
            E = max(0, min(E, 120));
            const float Ef = (float)E;
            const float dsc = __uint_as_float((uint32_t)(127 - E) << 23) * vinv;
            const uint64_t E2 = pack2(Ef, Ef);
            // ---- pass 2: P'_j, split into fp16 hi / lo pairs, 32 entries per TMEM store ----
#define KMM_TC_SPLIT(i, P0, P1)                                                   \
    {                                                                             \
        const __half2 h2 = __floats2half2_rn(P0, P1);                             \
        const float2 f2 = __half22float2(h2);                                     \
        float l0, l1;                                                             \
        unpack2(sub2(pack2(P0, P1), pack2(f2.x, f2.y)), l0, l1);                  \
        const __half2 l2 = __floats2half2_rn(l0, l1);                             \
        phi[i] = *reinterpret_cast<const uint32_t*>(&h2);                         \
        plo[i] = *reinterpret_cast<const uint32_t*>(&l2);                         \
    }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t phi[16], plo[16];
                if (is_rbf) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a0, a1;
                        unpack2(add2(z[half * 16 + i], E2), a0, a1);
                        const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
                        KMM_TC_SPLIT(i, p0, p1)
                    }
                } else {
                    const uint64_t nl2 = pack2(-TC_LOG2E, -TC_LOG2E), one2 = pack2(1.0f, 1.0f),
                                   third2 = pack2(1.0f / 3.0f, 1.0f / 3.0f);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float z0, z1;
                        unpack2(z[half * 16 + i], z0, z1);
                        const float r0 = sqrt_approx(fmaxf(z0, 0.0f)), r1 = sqrt_approx(fmaxf(z1, 0.0f));
                        const uint64_t r2 = pack2(r0, r1);
                        float a0, a1;
                        unpack2(fma2(r2, nl2, E2), a0, a1);
                        uint64_t pp = pack2(ex2_approx(a0), ex2_approx(a1));
                        if (kid == KID_MATERN32) pp = mul2(pp, add2(r2, one2));
                        else if (kid == KID_MATERN52) pp = mul2(pp, fma2(r2, fma2(r2, third2, one2), one2));
                        float p0, p1;
                        unpack2(pp, p0, p1);
                        KMM_TC_SPLIT(i, p0, p1)
                    }
                }
                tmem_st16(t_s + half * 16, phi);       // P_hi: columns [0, 32) of the buffer
                tmem_st16(t_s + 32 + half * 16, plo);  // P_lo: columns [32, 64)
            }
#undef KMM_TC_SPLIT
            TC_PROF(3)
            tmem_wait_st();
            tc_fence_before();
            if (SPLIT) dsc_sm[(u & 7) * TC_BM + row] = dsc;  // published before p_full -> MMA2 -> o_full
            __syncwarp();
            if (lane == 0) arrive_pair(&p_full[b]);
            TC_PROF(4)
            // drain finished sub-tiles while the tensor core works on this one
            if (SPLIT) drain_through(u - 1);
            else if (u >= NWG) drain(CH0, g, (uint32_t)(((u - NWG) / NWG) & 1), dsc_prev);
            TC_PROF(5)
            dsc_prev = dsc;
            }
            }
            }  // !QUART
            b += NWG;
            if (b >= NB) {
                b -= NB;
                use ^= 1;
            }
            sv += NWG;
            if (sv >= SV) {
                sv -= SV;
                phv ^= 1;
            }
        }
        if (warp == 0) TC_PROF_FLUSH(8)
        else if (warp == 4) TC_PROF_FLUSH(16)
        if constexpr (FRAG) {
            // row sums: the four lanes that share a row hold the partial sums of their column subsets
            float rs[4][KV];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < KV; ++c) {
                    float v = accf2[r][c] + accf[r][c];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    rs[r][c] = v;
                }
            float* ysm = reinterpret_cast<float*>(smem);  // [NWG][128][KV], reuses the A ring
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");  // every warpgroup is done with the rings
            if (tq == 0) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int rl = q * 32 + (lane >> 2) + 8 * (r & 1) + 16 * (r >> 1);
#pragma unroll
                    for (int c = 0; c < KV; ++c) ysm[(g * TC_BM + rl) * KV + c] = rs[r][c];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
            if (g == 0 && grow < p.n) {
                float* dst = p.out + (int64_t)blockIdx.z * p.split_stride + grow * p.ldo;
#pragma unroll
                for (int c = 0; c < KV; ++c) {
                    float y = 0.0f;
#pragma unroll
                    for (int w = 0; w < NWG; ++w) y += ysm[(w * TC_BM + row) * KV + c];
                    if (c < p.k) dst[c] = y * p.scale_out;
                }
            }
        } else {
        if constexpr (KV > 0) {
            float cs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int c = 0; c < KV; ++c) {
                float lo, hi;
                unpack2(add2(accv2[c], accv[c]), lo, hi);
                cs[c] = lo + hi;
            }
            acc[0] = pack2(cs[0], cs[1]);
            acc[1] = pack2(cs[2], cs[3]);
        } else if (SPLIT) {
            drain_through(T - 1);
        } else {
            const int last = ((T - 1 - g) / NWG) * NWG + g;  // this warpgroup's last tile (T > g)
            if (T > g) drain(CH0, g, (uint32_t)((last / NWG) & 1), TC_DIAG(2) ? 1.0f : dsc_prev);
        }

        if (SPLIT) {
            // ---- each warpgroup holds complete sums for its half of the columns ----
            if (grow < p.n) {
                float* dst = p.out + (int64_t)blockIdx.z * p.split_stride + grow * p.ldo;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    if (ch == 1 && kc_b == kc) break;  // DUAL: the second chunk does not exist (odd number of chunks)
#pragma unroll
                    for (int c = 0; c < DW / 2; ++c) {
                        const int col = (kc + ch) * KP + g * DW + 2 * c;
                        float y0, y1;
                        unpack2(acc[ch * (DW / 2) + c], y0, y1);
                        if (col < p.k) dst[col] = y0 * p.scale_out;
                        if (col + 1 < p.k) dst[col + 1] = y1 * p.scale_out;
                    }
                }
            }
        } else {
            // ---- Y rows = sum of the warpgroups' partial rows (through smem; all MMAs and loads are done) ----
            float* ysm = reinterpret_cast<float*>(smem);  // [NWG - 1][128][KP + 1], reuses the A ring
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");  // every warpgroup has drained its last tile
            if (g >= 1) {
                float* mine = ysm + (size_t)(g - 1) * TC_BM * (KP + 1);
#pragma unroll
                for (int c = 0; c < DW / 2; ++c) {
                    float y0, y1;
                    unpack2(acc[c], y0, y1);
                    mine[row * (KP + 1) + 2 * c] = y0;
                    mine[row * (KP + 1) + 2 * c + 1] = y1;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
            if (g == 0 && grow < p.n) {
                float* dst = p.out + (int64_t)blockIdx.z * p.split_stride + grow * p.ldo;
#pragma unroll
                for (int c = 0; c < DW / 2; ++c) {
                    const int col = kc * KP + 2 * c;
                    float y0, y1;
                    unpack2(acc[c], y0, y1);
#pragma unroll
                    for (int w = 0; w < NWG - 1; ++w) {
                        y0 += ysm[(size_t)w * TC_BM * (KP + 1) + row * (KP + 1) + 2 * c];
                        y1 += ysm[(size_t)w * TC_BM * (KP + 1) + row * (KP + 1) + 2 * c + 1];
                    }
                    if (col < p.k) dst[col] = y0 * p.scale_out;
                    if (col + 1 < p.k) dst[col + 1] = y1 * p.scale_out;
                }
            }
        }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.pair) cluster_sync_all();  // no CTA leaves while its peer can still multicast into it or signal its barriers
    if (warp == TC_EPI_WARPS) {
        if constexpr (CG2) tmem_dealloc2(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

struct TcPlan {
    int kb, kp, k_chunks, a_stages, v_stages, nb, la, nwg, wide, splits, tiles_per_split, pair, kv, cg2, dual, dual_mode;
    int64_t sub_tiles;
    size_t smem_bytes, vimg_bytes, part_bytes;
};

int tc_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// kid: kernel id when known (the launch), -1 for sizing queries -- the workspace does not depend on what it selects
bool tc_plan(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count, TcPlan* pl, int kid = -1) {
    if (d < 1 || d > TC_MAX_D_WIDE || n < 1 || m < 1 || k < 1) return false;
    const bool wide = d > TC_MAX_D;
    const int kb = tc_kblocks(d);
    int kp = 16;
    while (kp < 64 && kp < k) kp *= 2;
    // k > 64: 128-column chunks halve the number of times S and the pointwise stage are recomputed; they need
    // 256 TMEM columns for the two O buffers, which leaves two S/P buffers for d <= 128
    if (k > 64 && (wide ? 0 : 64 * kb) + (wide ? 4 : 2) * 64 + 2 * 128 <= 512 && tc_env_int("RLAOPT_B200_TC_KP128", 1)) kp = 128;
    // Small d and k: the tensor work per tile is small and the kernel is bound by the pointwise stage (MUFU, TMEM
    // and mbarrier latencies); a third epilogue warpgroup keeps three tiles in flight per SM sub-partition.
    const int nwg_env = tc_env_int("RLAOPT_B200_TC_NWG", 3);
    int nwg = (kb <= 2 && kp <= 32 && nwg_env >= 3) ? 3 : 2;
    // A fourth one (quarter-row epilogue, 104 registers per thread; k <= 16, every kernel but Matern-1/2, whose near-pair
    // recompute holds the whole row) was built to fill the special-function pipe (XU 75 % busy with three warps per
    // sub-partition, profiles/r02_ncu_tc_matern52_c3_summary.md).  Correct, but its extra TMEM round trips cost more than
    // the fourth warp brings: Matern-5/2 d=32 k=16 1603 -> 1576, RBF 2475 -> 2185 Gentries/s (profiles/r02_tc_nwg4_ab.log).
    // Only with RLAOPT_B200_TC_NWG=4.
    if (nwg == 3 && kp == 16 && nwg_env >= 4 && kid >= 0 && kid != KID_MATERN12 && k > 4) nwg = 4;
    // k <= 4 (single right-hand sides: PCG, SAP / ASkotch oracles): the contraction with V runs on the CUDA cores in
    // the epilogue -- one FFMA2 per two entries and column instead of the fp16 split of P, its TMEM store, MMA2 and
    // the accumulator drain.  Three epilogue warpgroups, no CTA pairs.  RLAOPT_B200_TC_KV=0 switches it off.
    // d <= 128 only: with three K-blocks the A ring holds three stages and the plan would be {NB 4, SA 3, SV 5} for three
    // epilogue warpgroups -- neither ring depth is a multiple of three, so both the s_full and the v_full parity wait of a
    // warpgroup are ambiguous while another warpgroup's tile is in flight, and nothing in the poll guards them (the V ring
    // has its own producer in this mode).  scripts/tc_protocol_model.py, run_own(nwg=3, kv=True, nb=4, sa=3, sv=5) finds the
    // stale reads at once; 128 < d <= 192 takes the MMA2 path with two warpgroups, whose waits are unambiguous.
    int kv = (!wide && kb <= 2 && k <= 4 && tc_env_int("RLAOPT_B200_TC_KV", 1)) ? (k == 1 ? 1 : (k == 2 ? 2 : 4)) : 0;
    if (kv) nwg = 3;  // a fourth epilogue warpgroup (640 threads, 104 registers) measured 3 % slower at k = 1
    const int x_cols = wide ? 0 : 64 * kb;  // wide d: X streams through smem, only S/P and O live in TMEM
    // TMEM columns: 64 KB (X hi/lo) + 64 NB (S/P) + NWG KP (O) <= 512
    int nb = (512 - x_cols - nwg * kp) / 64;
    if (nb > nwg + 2) nb = nwg + 2;
    if (nb < 2) return false;
    if (nb < nwg) nwg = 2, kv = 0;
    if (wide) nb = 4, nwg = 2;  // two-tile segments, double buffered
    if (!wide) nb = max(nwg, min(nb, tc_env_int("RLAOPT_B200_TC_NB", nb)));
    int la = nb >= 4 ? 2 : 1;
    la = max(1, min(nb >= 3 ? nb - 2 : 1, tc_env_int("RLAOPT_B200_TC_LA", la)));
    // smem: A ring (column-tile images) + V ring (la stages deeper: V of tile t is consumed la tiles after its A)
    const size_t a_stage = wide ? (size_t)TC_WIDE_STAGE_BYTES : tc_image_bytes(kb), v_stage = tc_v_stage_bytes(kp);
    // k > 128: two 128-column chunks of V per S / P' (the DUAL instantiations; every kernel but Matern-1/2, whose near-pair
    // recompute holds the whole row; the kernel id is known at launch -- sizing queries see the same workspace either way).
    // Measured (profiles/r02_tc_dual_ab.log): it pays where S is expensive -- 64 < d <= 128: RBF d=128 k=256 362 -> 413
    // Gentries/s -- and loses for d <= 64 (k=1000 133 -> 109, Matern-5/2 d=32 k=200 551 -> 396): the two O buffers are
    // then the two chunks of ONE sub-tile, so a buffer is reused after one sub-tile instead of two and MMA2 waits for a
    // warpgroup that is still in its (long) pointwise stage.  RLAOPT_B200_TC_DUAL: 1 = where it pays (default), 0 = never,
    // 2 = wherever it is possible, 3 = as 2, with the SLICED epilogue where three S/P buffers fit (d <= 64): both warpgroups
    // drain every chunk as it completes and run the pointwise stage of their next sub-tile in quarters between the drains.
    // That removes the loss for d <= 64 and gains 5-7 % over one chunk per CTA when the chunk count is even (k=1000
    // 133 -> 143, Matern-5/2 d=32 k=200 545 -> 567 Gentries/s, bit-identical results); it is opt-in because a faster
    // schedule of the same code (P' announced one slot earlier, 147 Gentries/s) mis-computed about one CTA in 500 for a
    // reason that is not understood yet (profiles/r02_tc_dual_sliced.md).
    const int dual_env = tc_env_int("RLAOPT_B200_TC_DUAL", 1);
    bool dual = kp == 128 && !wide && k > 128 && kid >= 0 && kid != KID_MATERN12 && (dual_env >= 2 || (dual_env == 1 && kb >= 2)) &&
                !tc_env_int("RLAOPT_B200_TC_CG2", 0);
    size_t fixed = 8 * TC_BM * sizeof(float) + 64 * sizeof(uint64_t) + 64;
    if (dual) fixed += 8 * TC_BM * sizeof(float) + 2048;  // one un-scale factor per row, tile and chunk; norms of the sliced mode
    // ring depth: 4 stages; deeper rings measured no gain even at one sub-tile per ~600 cycles (RLAOPT_B200_TC_SA)
    int sa = tc_env_int("RLAOPT_B200_TC_SA", 4), sv;
    if (sa < 2) sa = 2;
    if (sa > 8) sa = 8;
    if (wide) {  // V images of one or two segments in flight, the rest of smem for 64 KB K-block slots
        sv = kp > 64 ? 2 : 4;
        sa = 3;
        while (sa > 2 && sa * a_stage + sv * v_stage + fixed > (size_t)TC_SMEM_LIMIT) --sa;
    } else {
        while (sa > 2 && sa * a_stage + (sa + la) * v_stage + fixed > (size_t)TC_SMEM_LIMIT) --sa;
        sv = sa + la;
        // Two epilogue warpgroups own alternate sub-tiles: with an ODD V ring they share its stages, and a warpgroup's
        // v_full parity wait reads "complete" while the other warpgroup's record of tile u - SV is still in flight.  The
        // s_full wait of the same poll covers that, but it is sampled after v_full (try_wait may suspend in between).  For
        // the k > 64 plan {NB 2, SA 2, SV 3} one more stage fits and makes every wait of that kernel unambiguous (the ring
        // configuration the two-chunk kernels run with); scripts/tc_protocol_model.py, run_single(nb=2, sa=2, sv=3 / 4).
        if (!dual && nwg == 2 && kp == 128 && (sv & 1) && sa * a_stage + (sv + 1) * v_stage + fixed <= (size_t)TC_SMEM_LIMIT) ++sv;
        if (dual) {  // a sub-tile consumes two V records: at least two sub-tiles in flight
            if (sv < 4) sv = 4;
            while (sa > 2 && sa * a_stage + sv * v_stage + fixed > (size_t)TC_SMEM_LIMIT) --sa;
            if (sa * a_stage + sv * v_stage + fixed > (size_t)TC_SMEM_LIMIT) {
                dual = false;
                fixed -= 8 * TC_BM * sizeof(float) + 2048;
                sv = sa + la;
            }
        }
    }
    if (sa * a_stage + sv * v_stage + fixed > (size_t)TC_SMEM_LIMIT) return false;
    if (3 * sa + 3 * sv + 3 * nb + 2 * nwg + 1 > 64) return false;  // mbarriers (incl. the CG2 relay barriers) fit the array
    pl->kb = kb;
    pl->kp = kp;
    pl->k_chunks = (int)((k + kp - 1) / kp);
    pl->a_stages = sa;
    pl->v_stages = sv;
    pl->nb = nb;
    pl->nwg = nwg;
    pl->la = la;
    pl->wide = wide ? 1 : 0;
    pl->sub_tiles = (m + TC_BN - 1) / TC_BN;
    // CTA pairs (clusters of two row blocks sharing every column-tile load through multicast): +3 % under the
    // power cap at C2.  Default: on once the launch has at least two waves of row blocks; RLAOPT_B200_TC_PAIR=0/1 forces.
    {
        const int64_t row_blocks = (n + TC_BM - 1) / TC_BM;
        const int want = tc_env_int("RLAOPT_B200_TC_PAIR", -1);
        pl->pair = row_blocks >= 2 && (want < 0 ? row_blocks >= 2 * (int64_t)sm_count : want != 0);
        if (kv) pl->pair = 0;
        // k > 64 (128-column chunks): CTA pairs with cta_group::2 MMAs -- each SM receives half of every operand image.
        // Correct (same results to the last bit of the error budget) but measured 20 % SLOWER than single-CTA MMAs
        // (profiles/r02_tc_cg2_ab.log: k = 1000 131.7 -> 104.9 Gentries/s): this family is bound by its epilogue
        // (pointwise stage + two 64-column drains per tile and warpgroup, profiles/r02_tc_k1000_section_cycles.log), not
        // by the SM ingress, and coupling two SMs adds their barrier latencies.  Off unless RLAOPT_B200_TC_CG2=1.
        pl->cg2 = (kp == 128 && !wide && row_blocks >= 2 && tc_env_int("RLAOPT_B200_TC_CG2", 0)) ? 1 : 0;
        if (pl->cg2) pl->pair = 1;
    }
    pl->kv = kv;
    pl->dual = dual ? 1 : 0;
    // sliced mode (the pointwise stage of a sub-tile runs in four slices between the drains of the two sub-tiles before
    // it): needs three S/P buffers (d <= 64)
    pl->dual_mode = dual ? ((dual_env >= 3 && nb >= 3) ? 2 : 1) : 0;
    pl->smem_bytes = sa * a_stage + sv * v_stage + fixed;
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
    {
        // Multi-wave launches: cap the sub-tiles one CTA sweeps so that all row blocks visit an L2-sized
        // column chunk (about 48 MB of tile images) before the grid moves on -- blockIdx.z is the
        // slowest-varying index of the launch order.  Cuts DRAM re-reads of the column panel from ~1000x to
        // ~20x the algorithmic bytes at C2 (+3 % under the power cap) for splits x n x k floats of workspace.
        const size_t tile_bytes = tc_image_bytes(kb) + (size_t)kp * 256;
        int64_t cap = tc_env_int("RLAOPT_B200_TC_SPLIT_TILES", -1);
        const bool forced = cap > 0;  // tests force small chunks on small problems
        if (cap < 0) cap = (int64_t)((48u << 20) / tile_bytes);
        if (cap > 0 && (forced || (base >= target && (size_t)pl->sub_tiles * tile_bytes > ((size_t)64 << 20)))) {
            if (cap < TC_MIN_SPLIT_TILES) cap = TC_MIN_SPLIT_TILES;
            const int64_t floor_tps = (pl->sub_tiles + 31) / 32;  // at most 32 partial buffers
            if (cap < floor_tps) cap = floor_tps;
            if (cap < tps) tps = cap;
        }
    }
    splits = (pl->sub_tiles + tps - 1) / tps;
    pl->splits = (int)splits;
    pl->tiles_per_split = (int)tps;
    pl->vimg_bytes = kv ? (size_t)pl->sub_tiles * (kv + 1) * 256 : (size_t)pl->k_chunks * pl->sub_tiles * tc_v_record_bytes(kp);
    pl->part_bytes = splits > 1 ? (size_t)splits * n * k * sizeof(float) : 0;
    return true;
}

}  // namespace

#ifdef KMM_TC_PROFILE
extern "C" int kmm_tc_prof_read(long long* host64) {
    return (int)cudaMemcpyFromSymbol(host64, g_tc_prof, sizeof(long long) * 64);
}
#endif

bool tc_supported_d(int64_t d) { return d >= 1 && d <= TC_MAX_D_WIDE; }
bool tc_supported_k(int64_t k) { return k >= 1; }

size_t tc_packed_bytes(int64_t n, int64_t d) {
    if (n <= 0 || !tc_supported_d(d)) return 0;
    return tc_image_offset(n) + (size_t)(tc_npad(n) / 64) * tc_image_bytes(tc_kblocks(d));
}

cudaError_t launch_tc_pack(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                           float inv_ls, const float* inv_ls_vec, const float* center, void* packed,
                           cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(packed, 0, TC_HEADER_BYTES, stream);
    if (err != cudaSuccess) return err;
    const unsigned blocks = (unsigned)((n + TC_ABSMAX_ROWS - 1) / TC_ABSMAX_ROWS);
    tc_absmax_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(X, n, n_src, d, ldx, idx, inv_ls, inv_ls_vec, center,
                                                                 reinterpret_cast<TcHeader*>(packed));
    const int64_t n_pad = tc_npad(n);
    tc_pack_points_kernel<<<(unsigned)(n_pad / TC_BN), 256, 0, stream>>>(
        X, n, n_src, d, ldx, idx, inv_ls, inv_ls_vec, center, static_cast<unsigned char*>(packed), tc_kblocks(d));
    return cudaGetLastError();
}

// header fields the host reads back for its accuracy guard / index check (the copy is the caller's)
size_t tc_stats_offset() { return offsetof(TcHeader, max_sqnorm_bits); }

size_t tc_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int sm_count, bool keep_partials) {
    TcPlan pl;
    if (!tc_plan(n, m, d, k, sm_count, &pl)) return 0;
    const size_t part = keep_partials ? (size_t)pl.splits * n * k * sizeof(float) : pl.part_bytes;
    return round_up((int64_t)pl.vimg_bytes, 256) + part;
}

template <int KP, int NWG, bool M12, bool WIDE, int KV = 0, int KIDT = -1, bool CG2 = false, bool DUAL = false>
static cudaError_t launch_tc_inst(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    auto kern = kmm_tc_kernel<KP, NWG, M12, WIDE, KV, KIDT, CG2, DUAL>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (err != cudaSuccess) return err;
    unsigned row_blocks = (unsigned)((n + TC_BM - 1) / TC_BM);
    if (p.pair) row_blocks = (row_blocks + 1) & ~1u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(row_blocks, (unsigned)(pl.dual ? (pl.k_chunks + 1) / 2 : pl.k_chunks), (unsigned)pl.splits);
    cfg.blockDim = dim3(tc_threads(NWG));
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int KP, int NWG>
static cudaError_t launch_tc_kp(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    if constexpr (NWG == 2) {
        if (p.wide) {
            switch (p.kid) {
                case KID_RBF: return launch_tc_inst<KP, NWG, false, true, 0, KID_RBF>(p, pl, n, stream);
                case KID_MATERN32: return launch_tc_inst<KP, NWG, false, true, 0, KID_MATERN32>(p, pl, n, stream);
                case KID_MATERN52: return launch_tc_inst<KP, NWG, false, true, 0, KID_MATERN52>(p, pl, n, stream);
                default: return launch_tc_inst<KP, NWG, true, true, 0, KID_MATERN12>(p, pl, n, stream);
            }
        }
    }
    if constexpr (KP == 128) {
        if (pl.dual) {  // k > 128: two chunks of V per P'
            switch (p.kid) {
                case KID_RBF: return launch_tc_inst<KP, NWG, false, false, 0, KID_RBF, false, true>(p, pl, n, stream);
                case KID_MATERN32: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN32, false, true>(p, pl, n, stream);
                default: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN52, false, true>(p, pl, n, stream);
            }
        }
        if (pl.cg2) {  // CTA pairs with cta_group::2 MMAs
            switch (p.kid) {
                case KID_RBF: return launch_tc_inst<KP, NWG, false, false, 0, KID_RBF, true>(p, pl, n, stream);
                case KID_MATERN32: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN32, true>(p, pl, n, stream);
                case KID_MATERN52: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN52, true>(p, pl, n, stream);
                default: return launch_tc_inst<KP, NWG, true, false, 0, KID_MATERN12, true>(p, pl, n, stream);
            }
        }
    }
    {
        // X-resident families: one kernel function per instantiation, see KIDT (Matern-1/2, -3/2, -5/2 at d = 32,
        // k = 16: +21 %, +19 %, +10 %; RBF +3.5 %)
        switch (p.kid) {
            case KID_RBF: return launch_tc_inst<KP, NWG, false, false, 0, KID_RBF>(p, pl, n, stream);
            case KID_MATERN32: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN32>(p, pl, n, stream);
            case KID_MATERN52: return launch_tc_inst<KP, NWG, false, false, 0, KID_MATERN52>(p, pl, n, stream);
            default: return launch_tc_inst<KP, NWG, true, false, 0, KID_MATERN12>(p, pl, n, stream);
        }
    }
}

// four epilogue warpgroups (quarter-row epilogue): k <= 16, RBF / Matern-3/2 / Matern-5/2
static cudaError_t launch_tc_nwg4(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    switch (p.kid) {
        case KID_RBF: return launch_tc_inst<16, 4, false, false, 0, KID_RBF>(p, pl, n, stream);
        case KID_MATERN32: return launch_tc_inst<16, 4, false, false, 0, KID_MATERN32>(p, pl, n, stream);
        default: return launch_tc_inst<16, 4, false, false, 0, KID_MATERN52>(p, pl, n, stream);
    }
}

// register-contraction instantiations: (k <= 1, 2, 4) x (RBF, Matern-3/2, Matern-5/2 fixed at compile time; Matern-1/2
// with its near-pair recompute), three epilogue warpgroups
template <int KV, int KIDT>
static cudaError_t launch_tc_kv_kid(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    return launch_tc_inst<16, 3, false, false, KV, KIDT>(p, pl, n, stream);
}
template <int KV>
static cudaError_t launch_tc_kv_k(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    switch (p.kid) {
        case KID_RBF: return launch_tc_kv_kid<KV, KID_RBF>(p, pl, n, stream);
        case KID_MATERN32: return launch_tc_kv_kid<KV, KID_MATERN32>(p, pl, n, stream);
        case KID_MATERN52: return launch_tc_kv_kid<KV, KID_MATERN52>(p, pl, n, stream);
        default: return launch_tc_inst<16, 3, true, false, KV, KID_MATERN12>(p, pl, n, stream);
    }
}
static cudaError_t launch_tc_kv(const TcParams& p, const TcPlan& pl, int64_t n, cudaStream_t stream) {
    switch (pl.kv) {
        case 1: return launch_tc_kv_k<1>(p, pl, n, stream);
        case 2: return launch_tc_kv_k<2>(p, pl, n, stream);
        default: return launch_tc_kv_k<4>(p, pl, n, stream);
    }
}

cudaError_t launch_tc(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                      const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kid, float scale,
                      int sm_count, void* workspace, size_t workspace_bytes, cudaStream_t stream, bool keep_partials,
                      int* splits_out, const float** part_out) {
    TcPlan pl;
    if (!tc_plan(n, m, d, k, sm_count, &pl, kid)) return cudaErrorInvalidValue;
    const size_t v_bytes = (size_t)round_up((int64_t)pl.vimg_bytes, 256);
    const size_t part_bytes = keep_partials ? (size_t)pl.splits * n * k * sizeof(float) : pl.part_bytes;
    if (workspace == nullptr || workspace_bytes < v_bytes + part_bytes) return cudaErrorInvalidValue;
    unsigned char* vimg = static_cast<unsigned char*>(workspace);
    float* part = reinterpret_cast<float*>(vimg + v_bytes);
    if (pl.kv) {
        const int64_t total = pl.sub_tiles * (pl.kv + 1) * TC_BN;
        const float* col_norms = reinterpret_cast<const float*>(static_cast<const unsigned char*>(cols_packed) + tc_norm_offset());
        tc_pack_v_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(V, m, k, ldv, col_norms,
                                                                                    reinterpret_cast<float*>(vimg), pl.kv, pl.sub_tiles);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    } else {
        const size_t pack_smem = (size_t)TC_BN * (pl.kp + 1) * sizeof(float);
        dim3 grid((unsigned)pl.sub_tiles, (unsigned)pl.k_chunks);
        const float* col_norms = reinterpret_cast<const float*>(static_cast<const unsigned char*>(cols_packed) + tc_norm_offset());
        tc_pack_v_kernel<<<grid, 256, pack_smem, stream>>>(V, m, k, ldv, col_norms, vimg, pl.kp, pl.sub_tiles);
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
    p.k_chunks = pl.k_chunks;
    p.dual_mode = pl.dual_mode;
    // sliced mode, bits: 1 = quarter loads overlap the drain; 4 = one quarter per slot, P'(u) announced at the end of slot B of
    // sub-tile u - 1 (the schedule that reproduces the one-chunk kernel bit for bit); 8 = early schedule with the announcement
    // deferred to the start of that slot; 2 = column norms straight from global memory (diagnostic).  Without bits 4 / 8 the
    // announcement comes while MMA2 of chunk B of sub-tile u - 1 is still in flight: 10 % faster, and WRONG in about one CTA in
    // 500 (profiles/r02_tc_dual_sliced.md) -- kept for the investigation only.
    p.dual_overlap = tc_env_int("RLAOPT_B200_TC_DUAL_OVERLAP", 5);
    p.kb = pl.kb;
    p.nk1 = (int)((d + 15) / 16);
    p.a_stages = pl.a_stages;
    p.v_stages = pl.v_stages;
    p.nb = pl.nb;
    p.la = pl.la;
    p.diag = tc_env_int("RLAOPT_B200_TC_DIAG", 0);
    p.pair = pl.pair;
    p.wide = pl.wide;
    p.kid = kid;
    p.sub_tiles = pl.sub_tiles;
    p.tiles_per_split = pl.tiles_per_split;
    if (splits_out) *splits_out = pl.splits;
    if (part_out) *part_out = part;
    if (pl.splits > 1 || keep_partials) {
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
    if (pl.kv) {
        err = launch_tc_kv(p, pl, n, stream);
    } else
    switch (pl.kp) {
        case 16:
            err = pl.nwg == 4   ? launch_tc_nwg4(p, pl, n, stream)
                  : pl.nwg == 3 ? launch_tc_kp<16, 3>(p, pl, n, stream)
                                : launch_tc_kp<16, 2>(p, pl, n, stream);
            break;
        case 32: err = pl.nwg == 3 ? launch_tc_kp<32, 3>(p, pl, n, stream) : launch_tc_kp<32, 2>(p, pl, n, stream); break;
        case 64: err = launch_tc_kp<64, 2>(p, pl, n, stream); break;
        default: err = launch_tc_kp<128, 2>(p, pl, n, stream); break;
    }
    if (err != cudaSuccess) return err;
    if (pl.splits > 1 && !keep_partials) return launch_split_reduce<float>(part, pl.splits, n, k, Y, ldy, scale, stream);
    return cudaSuccess;
}

}  // namespace kmm

// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM, alone and while tcgen05.mma runs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int n, int m) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ts32(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .b64 bd;\nmov.b64 bd, {%2, %3};\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, 1;\n}\n"
                 ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr))
#define ST32(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
    :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

// warps 0..NW-1: TMEM ld/st loops (warp w touches lane quarter w & 3, columns 256 + (w >> 2) * 32 ...)
// warp NW: MMA issue (if do_mma), N = 64 TS, D at columns 64..127, A at 0..31
// op: 0 = ld only (batches of `depth` loads per wait), 1 = st only, 2 = none (MMA only)
__global__ void __launch_bounds__(288, 1) bench(int nw, int op, int depth, int do_mma, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp < nw && op != 2) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 128;
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = i + lane;
        uint32_t sink = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; i += depth) {
            if (op == 0) {
                for (int j = 0; j < depth; ++j) {
                    LD32(base + (j & 3) * 32, r);
                    if (depth > 1) {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        sink += r[0] + r[31];
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                sink += r[0] + r[31];
            } else {
                for (int j = 0; j < depth; ++j) ST32(base + (j & 3) * 32, r);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) {
            out[2 + warp] = t1 - t0;
            out[20] = sink;
        }
    }
    if (warp == 8 && do_mma) {
        constexpr uint32_t idesc = umma_idesc(0, 64, 128);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem));
        const uint32_t blo = (uint32_t)bdesc, bhi = (uint32_t)(bdesc >> 32);
        const uint32_t a_t = tmem, d0 = tmem + 64;
        long long t0 = clock64();
        for (int i = 0; i < do_mma; i += 8) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_ts32(d0 + (j & 1) * 64, a_t + (j & 3) * 8, blo + (j & 3) * 2, bhi, idesc);
            }
            __syncwarp();
        }
        if (elect_one()) commit(&bar);
        __syncwarp();
        while (!mbar_try_wait(&bar, 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 32 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    const int iters = 2048;
    for (int do_mma : {0, 4096}) {
        for (int op : {0, 1, 2}) {
            for (int nw : {4, 8}) {
                for (int depth : {1, 4}) {
                    if (op == 2 && (!do_mma || nw != 4 || depth != 1)) continue;
                    cudaMemset(out, 0, 32 * 8);
                    // scale MMA count so both loops last about as long
                    bench<<<148, 288, 16 * 1024>>>(nw, op, depth, do_mma, iters, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[32];
                    cudaMemcpy(h, out, 32 * 8, cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int w = 0; w < nw; ++w) mx = h[2 + w] > mx ? h[2 + w] : mx;
                    const double bytes = (double)nw * iters * 32 * 32 * 4;
                    printf("mma %d  op %s  warps %d  depth %d : ", do_mma ? 1 : 0, op == 0 ? "ld" : (op == 1 ? "st" : "--"), nw, depth);
                    if (op != 2) printf("%.1f clk per x32 op per warp, %.1f B/clk/SM  ", (double)mx / iters, bytes / mx);
                    if (do_mma) printf("| mma %.1f clk each (ideal 32)", (double)h[0] / do_mma);
                    printf("\n");
                }
            }
        }
    }
    return 0;
}

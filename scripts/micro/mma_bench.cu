// Micro-benchmark: tcgen05.mma issue/complete rate as a function of N, operand source (TMEM / smem A),
// and accumulator dependency pattern.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdint>
#include <cstdio>
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void mma_ss32(uint32_t d, uint32_t alo, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .b64 bd, ad;\nmov.b64 bd, {%2, %3};\nmov.b64 ad, {%1, %3};\ntcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, 1;\n}\n"
                 ::"r"(d), "r"(alo), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts32_tf32(uint32_t d, uint32_t a, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n.reg .b64 bd;\nmov.b64 bd, {%2, %3};\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, 1;\n}\n"
                 ::"r"(d), "r"(a), "r"(blo), "r"(bhi), "r"(idesc) : "memory");
}

// MODE: 0 TS same-D, 1 TS alternating 2 D, 2 SS same-D, 3 SS alternating 2 D, 4 TS tf32 same-D,
//       5 TS, same A columns every step (A reuse), 6 TS alternating 4 D
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc(MODE == 4 ? 2 : 0, N, 128);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem));
        const uint32_t blo = (uint32_t)bdesc, bhi = (uint32_t)(bdesc >> 32);
        const uint32_t alo = blo + (32768 >> 4);
        const uint32_t a_t = tmem;       // A operand: columns [0, 32)
        const uint32_t d0 = tmem + 64;   // D buffers from column 64
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t d = d0;
                    if (MODE == 1 || MODE == 3) d = d0 + (j & 1) * N;
                    if (MODE == 6) d = d0 + (j & 3) * N;
                    const uint32_t ks = j & 3;
                    if (MODE == 2 || MODE == 3) mma_ss32(d, alo + ks * 2, blo + ks * 2, bhi, idesc);
                    else if (MODE == 4) mma_ts32_tf32(d, a_t + ks * 8, blo + ks * 2, bhi, idesc);
                    else if (MODE == 5) mma_ts32(d, a_t, blo + ks * 2, bhi, idesc);
                    else mma_ts32(d, a_t + ks * 8, blo + ks * 2, bhi, idesc);
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) commit(&bar);
        __syncwarp();
        while (!mbar_try_wait(&bar, 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, int N>
void run(int grid, long long* out) {
    const char* names[] = {"TS same-D", "TS alt-2-D", "SS same-D", "SS alt-2-D", "TS tf32 same-D", "TS same-A same-D", "TS alt-4-D"};
    const int iters = 4096;
    cudaFuncSetAttribute(bench<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    bench<MODE, N><<<grid, 128, 64 * 1024>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[2];
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("grid %3d  %-18s N=%3d  issue %.1f clk/mma  complete %.1f clk/mma  (ideal %d)\n", grid, names[MODE], N,
           (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
}

template <int MODE>
void run_all(int grid, long long* out) {
    run<MODE, 16>(grid, out);
    run<MODE, 32>(grid, out);
    run<MODE, 64>(grid, out);
    if (MODE != 6) run<MODE, 128>(grid, out);
    if (MODE != 6 && MODE != 1 && MODE != 3) run<MODE, 256>(grid, out);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    for (int grid : {1, 148}) {
        run_all<0>(grid, out);
        run_all<1>(grid, out);
        run_all<2>(grid, out);
        run_all<3>(grid, out);
        run_all<4>(grid, out);
        run_all<5>(grid, out);
        run_all<6>(grid, out);
    }
    return 0;
}

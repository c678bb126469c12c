// Micro-benchmark 2 for the CUDA-core kernel: what runs "for free" next to the FADD2/FADD L1-distance loop?
//   mode 0: distance loop alone (8 x 8 tile from shared memory)
//   mode 1: + NX integer ops (VIADD + LOP3 pairs) per feature on independent registers
//   mode 2: + NM mma.sync.m16n8k8 tf32 per feature
//   mode 3: mma.sync alone (rate of the legacy tensor path on sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o simt_mix2_bench simt_mix2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int D = 32, BM = 128, BN = 64, NT = 128;

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool DIST, int NX, int NM>
__global__ void __launch_bounds__(NT, 3) bench(float* out, int iters) {
    __shared__ __align__(16) float As[D][BM];
    __shared__ __align__(16) float Bs[D][BN];
    const int tid = threadIdx.x, tx = tid % 8, ty = tid / 8;
    for (int q = tid; q < D * BM; q += NT) (&As[0][0])[q] = (q * 37 % 101) * 0.01f;
    for (int q = tid; q < D * BN; q += NT) (&Bs[0][0])[q] = (q * 53 % 103) * 0.01f;
    __syncthreads();
    float S[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) S[r][c] = 0.f;
    uint32_t xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) xi[e] = tid * 977 + e;
    float cf[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) cf[e][f] = 0.f;
    uint32_t af[4] = {(uint32_t)tid, (uint32_t)tid * 3, (uint32_t)tid * 5, (uint32_t)tid * 7};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int dd = 0; dd < D; ++dd) {
            if (DIST) {
                alignas(16) float a[8], b[8];
                *(float4*)&a[0] = *(const float4*)&As[dd][ty * 8]; *(float4*)&a[4] = *(const float4*)&As[dd][ty * 8 + 4];
                *(float4*)&b[0] = *(const float4*)&Bs[dd][tx * 8]; *(float4*)&b[4] = *(const float4*)&Bs[dd][tx * 8 + 4];
                uint64_t bp[4];
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) bp[c2] = pk2(b[2 * c2], b[2 * c2 + 1]);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint64_t ar = pk2(a[r], a[r]);
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) {
                        float lo, hi;
                        up2(sub2(ar, bp[c2]), lo, hi);
                        S[r][2 * c2] += fabsf(lo);
                        S[r][2 * c2 + 1] += fabsf(hi);
                    }
                    if (r < NX) {  // one VIADD + one LOP3 on a private register
                        asm volatile("add.u32 %0, %0, 4096;" : "+r"(xi[r % 8]));
                        asm volatile("and.b32 %0, %0, 0xffffe000;" : "+r"(xi[(r + 4) % 8]));
                    }
                    if (r < NM) mma_tf32(cf[r % 4], af, xi[0], xi[1]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < NM; ++r) mma_tf32(cf[r % 4], af, xi[0], xi[1]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += S[r][c];
#pragma unroll
    for (int e = 0; e < 8; ++e) s += (float)xi[e];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) s += cf[e][f];
    out[blockIdx.x * NT + tid] = s;
}

template <bool DIST, int NX, int NM>
void run(float* out, int sms, const char* what) {
    const int iters = 2000, grid = sms * 3;
    bench<DIST, NX, NM><<<grid, NT>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<DIST, NX, NM><<<grid, NT>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double cyc_per_feature_per_warp_slot = ms * 1e-3 * 1.965e9 / ((double)D * iters) / 3.0;  // 3 warps per SMSP
    printf("%-44s %8.3f ms  %7.1f SMSP-cycles per (warp, feature)  err=%s\n", what, ms, cyc_per_feature_per_warp_slot,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, 148 * 3 * NT * sizeof(float) * 2);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs; ideal distance loop = 128 FP32-pipe cycles per (warp, feature)\n", p.name, sms);
    run<true, 0, 0>(out, sms, "distance loop alone");
    run<true, 4, 0>(out, sms, "+ 4 VIADD + 4 LOP3 per feature");
    run<true, 8, 0>(out, sms, "+ 8 VIADD + 8 LOP3 per feature");
    run<true, 0, 2>(out, sms, "+ 2 mma.sync tf32 per feature");
    run<true, 0, 4>(out, sms, "+ 4 mma.sync tf32 per feature");
    run<true, 8, 4>(out, sms, "+ 8 VIADD + 8 LOP3 + 4 mma.sync per feature");
    run<false, 0, 8>(out, sms, "8 mma.sync tf32 per 'feature', nothing else");
    return 0;
}

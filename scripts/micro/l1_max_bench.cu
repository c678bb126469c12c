// Micro-benchmark 3: L1 distance with |a-b| = 2 max(a,b) - a - b on NMAX of every 3 features (FMNMX on the ALU pipe +
// FFMA2 accumulate) and the direct form (FADD2 + FADD |.|) on the rest.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_max_bench l1_max_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int D = 33, BM = 128, BN = 64, NT = 128;

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int NMAX>
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
    const uint64_t two = pk2(2.0f, 2.0f);
    for (int it = 0; it < iters; ++it) {
        for (int d3 = 0; d3 < D; d3 += 3) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int dd = d3 + u;
                alignas(16) float a[8], b[8];
                *(float4*)&a[0] = *(const float4*)&As[dd][ty * 8]; *(float4*)&a[4] = *(const float4*)&As[dd][ty * 8 + 4];
                *(float4*)&b[0] = *(const float4*)&Bs[dd][tx * 8]; *(float4*)&b[4] = *(const float4*)&Bs[dd][tx * 8 + 4];
                if (u < NMAX) {
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) {
                            const float m0 = fmaxf(a[r], b[2 * c2]), m1 = fmaxf(a[r], b[2 * c2 + 1]);
                            up2(fma2(pk2(m0, m1), two, pk2(S[r][2 * c2], S[r][2 * c2 + 1])), S[r][2 * c2], S[r][2 * c2 + 1]);
                        }
                } else {
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
                    }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += S[r][c];
    out[blockIdx.x * NT + tid] = s;
}

template <int NMAX>
void run(float* out, int sms) {
    const int iters = 2000, grid = sms * 3;
    bench<NMAX><<<grid, NT>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<NMAX><<<grid, NT>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("max-form on %d of 3 features: %8.3f ms  %7.1f SMSP-cycles per (warp, feature) at 1.965 GHz  err=%s\n", NMAX, ms,
           ms * 1e-3 * 1.965e9 / ((double)D * iters) / 3.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, 148 * 3 * NT * sizeof(float) * 2);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>(out, p.multiProcessorCount);
    run<1>(out, p.multiProcessorCount);
    run<2>(out, p.multiProcessorCount);
    run<3>(out, p.multiProcessorCount);
    return 0;
}

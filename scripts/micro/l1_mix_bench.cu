// Micro-benchmark: L1-distance register tile (8 x 8 per thread, operands from shared memory) with NI of the 8 rows
// accumulated by the integer pipe (VABSDIFF d = |a - b| + c on fixed-point operands) and the rest by the FP32 pipe
// (FADD2 difference + FADD |.| accumulate).  Prints feature-entries per second for NI = 0..8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_mix_bench l1_mix_bench.cu && ./l1_mix_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int D = 32, BM = 128, BN = 64, NT = 128;

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int NI>
__global__ void __launch_bounds__(NT, 3) bench(float* out, int iters) {
    __shared__ __align__(16) float As[D][BM];
    __shared__ __align__(16) float Bs[D][BN];
    __shared__ __align__(16) int Ai[D][BM];
    __shared__ __align__(16) int Bi[D][BN];
    const int tid = threadIdx.x, tx = tid % 8, ty = tid / 8;
    for (int q = tid; q < D * BM; q += NT) { (&As[0][0])[q] = (q * 37 % 101) * 0.01f; (&Ai[0][0])[q] = q * 37 % 101; }
    for (int q = tid; q < D * BN; q += NT) { (&Bs[0][0])[q] = (q * 53 % 103) * 0.01f; (&Bi[0][0])[q] = q * 53 % 103; }
    __syncthreads();
    float S[8][8];
    int Si[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) { S[r][c] = 0.f; Si[r][c] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int dd = 0; dd < D; ++dd) {
            alignas(16) float a[8], b[8];
            alignas(16) int ai[8], bi[8];
            if (NI < 8) {
                *(float4*)&a[0] = *(const float4*)&As[dd][ty * 8]; *(float4*)&a[4] = *(const float4*)&As[dd][ty * 8 + 4];
                *(float4*)&b[0] = *(const float4*)&Bs[dd][tx * 8]; *(float4*)&b[4] = *(const float4*)&Bs[dd][tx * 8 + 4];
            }
            if (NI > 0) {
                *(int4*)&ai[0] = *(const int4*)&Ai[dd][ty * 8]; *(int4*)&ai[4] = *(const int4*)&Ai[dd][ty * 8 + 4];
                *(int4*)&bi[0] = *(const int4*)&Bi[dd][tx * 8]; *(int4*)&bi[4] = *(const int4*)&Bi[dd][tx * 8 + 4];
            }
            uint64_t bp[4];
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) bp[c2] = pk2(b[2 * c2], b[2 * c2 + 1]);
            // interleave integer and float rows so both pipes always have work queued
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const bool is_int = NI == 8 ? true : NI == 0 ? false : ((r * NI) / 8 != ((r + 1) * NI) / 8);
                if (is_int) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) asm("sad.s32 %0, %1, %2, %0;" : "+r"(Si[r][c]) : "r"(ai[r]), "r"(bi[c]));
                } else {
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
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += S[r][c] + (float)Si[r][c];
    out[blockIdx.x * NT + tid] = s;
}

template <int NI>
void run(float* out, int sms) {
    const int iters = 2000, grid = sms * 3;
    bench<NI><<<grid, NT>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<NI><<<grid, NT>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double fe = (double)grid * NT * 64.0 * D * iters;
    printf("int rows %d/8: %8.3f ms  %8.1f G feature-entries/s  (%.2f per clk per SM at 1.965 GHz)  err=%s\n", NI, ms,
           fe / ms / 1e6, fe / ms / 1e6 / 1.965 / sms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out;
    cudaMalloc(&out, 148 * 3 * NT * sizeof(float) * 2);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>(out, p.multiProcessorCount);
    run<2>(out, p.multiProcessorCount);
    run<3>(out, p.multiProcessorCount);
    run<4>(out, p.multiProcessorCount);
    run<5>(out, p.multiProcessorCount);
    run<6>(out, p.multiProcessorCount);
    run<8>(out, p.multiProcessorCount);
    return 0;
}

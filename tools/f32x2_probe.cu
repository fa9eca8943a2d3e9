// Packed FP32 (FADD2/FMUL2/FFMA2, sm_100) issue-rate probe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/f32x2_probe tools/f32x2_probe.cu
// A: 8 independent FFMA chains            B: 8 independent FFMA2 chains (2 floats each)
// C: 8 FFMA + 8 integer LOP3/IADD3        D: 4 FFMA2 + 8 integer ops (same float work as C)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float m, float c, unsigned k, int iters) {
    float f[8];
    uint64_t p[8];
    unsigned u[8];
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3f + i; p[i] = pack(f[i], f[i] + 0.5f); u[i] = threadIdx.x * 7u + i; }
    const uint64_t m2 = pack(m, m), c2 = pack(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], m, c);
            }
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], m2, c2);
            }
            if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], m2, c2);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ k) + (u[i] >> 3);   // LOP3/SHF/IADD3 mix on the ALU pipe
            }
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; ++i) { float a, b; unpack(p[i], a, b); acc += f[i] + a + b + (float)u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
double run(float* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(out, 0.999f, 0.001f, 0x5bd1e995u, 16);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(out, 0.999f, 0.001f, 0x5bd1e995u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double threads = 148.0 * 8 * 256;
    double a = run<0>(out, iters), b = run<1>(out, iters), c = run<2>(out, iters), d = run<3>(out, iters);
    printf("{\"ffma_ms\": %.3f, \"ffma_tflops\": %.2f, \"ffma2_ms\": %.3f, \"ffma2_tflops\": %.2f, "
           "\"ffma8_alu_ms\": %.3f, \"ffma2x4_alu_ms\": %.3f, \"note\": \"C and D do the same float work (32 fma/iter/thread) next to the same integer work\"}\n",
           a, threads * iters * 32 * 2 / (a * 1e-3) / 1e12, b, threads * iters * 64 * 2 / (b * 1e-3) / 1e12, c, d);
    return 0;
}

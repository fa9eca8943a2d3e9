// Do the half-rate pipes of an sm_100 SM sub-partition (FP64, integer ALU, IMAD on the heavy FMA
// half) dispatch independently?  Each kernel runs long independent chains of one or two
// instruction kinds per thread; if a mix takes the SUM of the single-kind times the two kinds
// share a dispatch slot, if it takes the MAX they overlap.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIter = 4096;
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, unsigned* iout, double seed, unsigned useed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3;
    const double m = 1.0000001, c = 1e-9;
    unsigned i0 = useed + threadIdx.x, i1 = useed * 3, i2 = useed * 5, i3 = useed * 7;
    float f0 = (float)seed, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
#pragma unroll 8
    for (int it = 0; it < kIter; ++it) {
        if (MODE & 1) {   // 4 DFMA
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        }
        if (MODE & 2) {   // 4 IADD3 (ALU)
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i0) : "r"(i1));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i1) : "r"(i2));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i2) : "r"(i3));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(i3) : "r"(i0));
        }
        if (MODE & 4) {   // 4 LOP3 (ALU)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i0) : "r"(i1), "r"(i2));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i1) : "r"(i2), "r"(i3));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i2) : "r"(i3), "r"(i0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i3) : "r"(i0), "r"(i1));
        }
        if (MODE & 8) {   // 4 IMAD (heavy half of the FMA pipe)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(i0) : "r"(i1), "r"(i2));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(i1) : "r"(i2), "r"(i3));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(i2) : "r"(i3), "r"(i0));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(i3) : "r"(i0), "r"(i1));
        }
        if (MODE & 16) {  // 4 FFMA
            f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f);
            f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + (double)(f0 + f1 + f2 + f3);
    iout[blockIdx.x * blockDim.x + threadIdx.x] = i0 ^ i1 ^ i2 ^ i3;
}
template <int MODE>
float run(double* out, unsigned* iout) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(out, iout, 1.0, 3u);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) probe<MODE><<<148 * 8, 256>>>(out, iout, 1.0, 3u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}
int main() {
    double* out; unsigned* iout;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 256); cudaMalloc(&iout, sizeof(unsigned) * 148 * 8 * 256);
    const double inst = 148.0 * 8 * 8 * kIter * 4;   // warp instructions of one kind per launch
    struct { const char* name; float ms; } r[] = {
        {"dfma", run<1>(out, iout)}, {"iadd3", run<2>(out, iout)}, {"lop3", run<4>(out, iout)}, {"imad", run<8>(out, iout)},
        {"ffma", run<16>(out, iout)}, {"dfma+iadd3", run<3>(out, iout)}, {"dfma+lop3", run<5>(out, iout)},
        {"dfma+imad", run<9>(out, iout)}, {"dfma+ffma", run<17>(out, iout)}, {"iadd3+imad", run<10>(out, iout)},
        {"iadd3+lop3", run<6>(out, iout)}, {"dfma+iadd3+lop3+imad", run<15>(out, iout)}, {"iadd3+ffma", run<18>(out, iout)}};
    printf("{");
    for (auto& x : r) printf("\"%s_ms\": %.4f, ", x.name, x.ms);
    printf("\"warp_instr_per_kind\": %.0f, \"cycles_per_warp_instr_per_smsp_dfma\": %.3f}\n", inst,
           r[0].ms * 1e-3 * 1.965e9 / (inst / (148.0 * 4)));
    return cudaGetLastError() != cudaSuccess;
}

// dp_probe.cu -- FP64 pipe behaviour on B200: achieved DFMA warp-instructions per clock per SM
// as a function of resident warps per SM sub-partition and independent chains per thread.
#include <cuda_runtime.h>
#include <stdio.h>
template <int ILP>
__global__ void dfma_chain(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
static void run(int sms, int warps_per_smsp, double* buf, int clk_khz) {
    const int threads = 128;                     // 4 warps = 1 per SMSP
    const int blocks = sms * warps_per_smsp;     // resident together (<= 16 blocks/SM fine)
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    dfma_chain<ILP><<<blocks, threads>>>(buf, iters, 1.0000001, 0.5);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    dfma_chain<ILP><<<blocks, threads>>>(buf, iters, 1.0000001, 0.5);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double winstr = (double)blocks * 4 * iters * ILP;
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("warps/SMSP %d ILP %d : %.3f DFMA warp-instr/clk/SM, %.1f cycles per dependent step\n", warps_per_smsp, ILP,
           winstr / cycles / sms, cycles / iters);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* buf; cudaMalloc(&buf, (size_t)p.multiProcessorCount * 16 * 128 * 8);
    for (int w : {1, 2, 4, 8}) {
        run<1>(p.multiProcessorCount, w, buf, clk);
        run<2>(p.multiProcessorCount, w, buf, clk);
        run<4>(p.multiProcessorCount, w, buf, clk);
        run<8>(p.multiProcessorCount, w, buf, clk);
    }
    return 0;
}

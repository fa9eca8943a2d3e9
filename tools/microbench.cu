// microbench.cu -- CUDA-core peaks the rooflines of this repo are quoted against.
// MEASURED_PEAKS.json (driver written) has HBM and bf16 tensor numbers only; the
// RDF / S(q) kernels are bound by the FP32 / FP64 pipes, MUFU and shared-memory
// atomics, so those are measured here with register-resident loops.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench.cu -o tools/microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// non-fused multiply and add (what the bit-exact RDF arithmetic issues)
template <typename T, int ILP>
__global__ void muladd_kernel(T* out, int iters, T a, T b);
template <int ILP>
__global__ void muladd_kernel_f(float* out, int iters, float a, float b) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __fadd_rn(__fmul_rn(v[i], a), b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void muladd_kernel_d(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __dadd_rn(__dmul_rn(v[i], a), b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void mufu_kernel(float* out, int iters) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 1.0f + threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// f64 <-> f32 conversion throughput (the complex64-rounding step of strict sf2d)
template <int ILP>
__global__ void cvt_kernel(double* out, int iters, double a) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 1.0 + threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = (double)(float)(v[i] + a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory histogram atomics: `active` lanes of 32 hit pseudo-random bins
// (nbins must be a power of two).  inc == 1 compiles to ATOMS.POPC.INC, any other
// run-time value to ATOMS.ADD.
template <bool POPC>
__global__ void atoms_kernel(unsigned* out, int iters, int nbins, int active, unsigned inc) {
    extern __shared__ unsigned hist[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const bool on = (threadIdx.x & 31) < active;
    const unsigned mask = nbins - 1;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        if (on) {
            if (POPC) atomicAdd(&hist[(x >> 8) & mask], 1u);
            else atomicAdd(&hist[(x >> 8) & mask], inc);
        }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s += hist[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half of the lanes increment one common "dump" bin, the other half random bins: the
// access pattern of an unconditional histogram update where out-of-range pairs go to a
// dump slot.  POPC=true -> ATOMS.POPC.INC.
template <bool POPC>
__global__ void atoms_dump_kernel(unsigned* out, int iters, int nbins, unsigned inc) {
    extern __shared__ unsigned hist[];
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const unsigned mask = nbins - 1;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        const unsigned bin = (x & 0x10000u) ? ((x >> 8) & mask) : (unsigned)nbins;
        if (POPC) atomicAdd(&hist[bin], 1u);
        else atomicAdd(&hist[bin], inc);
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) s += hist[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk_khz);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    void* buf; CK(cudaMalloc(&buf, (size_t)blocks * threads * 8));
    const double nthreads = (double)blocks * threads;
    constexpr int ILP = 8;
    double ms;
    ms = time_ms([&] { fma_kernel<float, ILP><<<blocks, threads>>>((float*)buf, iters, 1.0001f, 0.5f); });
    printf(", \"fp32_ffma_tflops\": %.2f", nthreads * iters * ILP * 2 / ms / 1e9);
    ms = time_ms([&] { fma_kernel<double, ILP><<<blocks, threads>>>((double*)buf, iters, 1.0001, 0.5); });
    printf(", \"fp64_dfma_tflops\": %.2f", nthreads * iters * ILP * 2 / ms / 1e9);
    ms = time_ms([&] { muladd_kernel_f<ILP><<<blocks, threads>>>((float*)buf, iters, 1.0001f, 0.5f); });
    printf(", \"fp32_mul_add_unfused_tflops\": %.2f", nthreads * iters * ILP * 2 / ms / 1e9);
    ms = time_ms([&] { muladd_kernel_d<ILP><<<blocks, threads>>>((double*)buf, iters, 1.0001, 0.5); });
    printf(", \"fp64_mul_add_unfused_tflops\": %.2f", nthreads * iters * ILP * 2 / ms / 1e9);
    ms = time_ms([&] { mufu_kernel<ILP><<<blocks, threads>>>((float*)buf, iters); });
    printf(", \"mufu_sqrt_gops\": %.1f", nthreads * iters * ILP / ms / 1e6);
    ms = time_ms([&] { cvt_kernel<ILP><<<blocks, threads>>>((double*)buf, iters / 4, 0.25); });
    printf(", \"cvt_f64_f32_roundtrip_gops\": %.1f", nthreads * (iters / 4) * ILP / ms / 1e6);
    for (int active : {32, 16, 8}) {
        for (int nbins : {512, 64}) {
            ms = time_ms([&] { atoms_kernel<true><<<blocks, threads, nbins * 4>>>((unsigned*)buf, iters, nbins, active, 1u); });
            printf(", \"atoms_popc_inc_nbins%d_active%d_gops\": %.1f", nbins, active,
                   nthreads * iters * active / 32.0 / ms / 1e6);
            ms = time_ms([&] { atoms_kernel<false><<<blocks, threads, nbins * 4>>>((unsigned*)buf, iters, nbins, active, 2u); });
            printf(", \"atoms_add_nbins%d_active%d_gops\": %.1f", nbins, active,
                   nthreads * iters * active / 32.0 / ms / 1e6);
        }
    }
    ms = time_ms([&] { atoms_dump_kernel<true><<<blocks, threads, 513 * 4>>>((unsigned*)buf, iters, 512, 1u); });
    printf(", \"atoms_popc_inc_half_dump_gops\": %.1f", nthreads * iters / ms / 1e6);
    ms = time_ms([&] { atoms_dump_kernel<false><<<blocks, threads, 513 * 4>>>((unsigned*)buf, iters, 512, 2u); });
    printf(", \"atoms_add_half_dump_gops\": %.1f", nthreads * iters / ms / 1e6);
    printf("}\n");
    return 0;
}

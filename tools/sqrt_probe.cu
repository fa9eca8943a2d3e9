// Exhaustive error of MUFU.SQRT (sqrt.approx.ftz.f32) against the exact square root over all
// positive normal floats: the bound the table-free RDF binning would have to carry.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void probe(double* max_rel, unsigned long long* n_off1, unsigned long long* n_off2) {
    double worst = 0.0;
    unsigned long long off1 = 0, off2 = 0;
    for (unsigned long long b = 0x00800000ull + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b < 0x7f800000ull;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float s = __uint_as_float((unsigned)b);
        float d;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
        const double ex = sqrt((double)s);
        const double rel = fabs((double)d - ex) / ex;
        worst = fmax(worst, rel);
        const float rn = (float)ex;
        const int du = abs((int)__float_as_uint(d) - (int)__float_as_uint(rn));
        off1 += du >= 1;
        off2 += du >= 2;
    }
    for (int o = 16; o > 0; o >>= 1) {
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        off1 += __shfl_xor_sync(0xffffffffu, off1, o);
        off2 += __shfl_xor_sync(0xffffffffu, off2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax((unsigned long long*)max_rel, (unsigned long long)__double_as_longlong(worst));
        atomicAdd(n_off1, off1);
        atomicAdd(n_off2, off2);
    }
}
int main() {
    double* d_rel; unsigned long long *d1, *d2;
    cudaMalloc(&d_rel, 8); cudaMalloc(&d1, 8); cudaMalloc(&d2, 8);
    cudaMemset(d_rel, 0, 8); cudaMemset(d1, 0, 8); cudaMemset(d2, 0, 8);
    probe<<<148 * 8, 256>>>(d_rel, d1, d2);
    double rel; unsigned long long o1, o2;
    cudaMemcpy(&rel, d_rel, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&o1, d1, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&o2, d2, 8, cudaMemcpyDeviceToHost);
    printf("{\"mufu_sqrt_max_rel_err\": %.6e, \"as_power_of_two\": %.3f, \"values_off_by_ge1ulp_from_rn\": %llu, \"values_off_by_ge2ulp\": %llu, \"values\": %llu}\n",
           rel, log2(rel), o1, o2, 0x7f800000ull - 0x00800000ull);
    return cudaGetLastError() != cudaSuccess;
}

// place_probe.cu -- what bounds the cell build?  Isolated rates on one B200:
//   (a) random global atomics (with / without return value) into K counters
//   (b) random 16-byte stores into a region of R bytes
//   (c) streaming reads of 12-byte records: plain loads, 1D bulk copies into shared memory
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o place_probe tools/place_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <bool RET, int ILP>
__global__ void atomic_rand(unsigned* cnt, unsigned K, long long n, unsigned* out) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x * ILP + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x * ILP) {
        unsigned r[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const unsigned c = hash32((unsigned)(i + (long long)u * blockDim.x)) % K;
            if (RET) r[u] = atomicAdd(&cnt[c], 1u);
            else { atomicAdd(&cnt[c], 1u); r[u] = 0; }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) acc += r[u];
    }
    if (acc == 0xffffffffu) out[0] = acc;
}

template <int ILP>
__global__ void scatter16(float4* dst, unsigned R16, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x * ILP + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x * ILP) {
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const unsigned s = hash32((unsigned)(i + (long long)u * blockDim.x)) % R16;
            dst[s] = make_float4(1.f, 2.f, 3.f, 0.f);
        }
    }
}

// scattered 16-byte stores whose destinations are unique and cover the region exactly (a permutation)
__global__ void scatter16_perm(float4* dst, const unsigned* perm, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[perm[i]] = make_float4(1.f, 2.f, 3.f, 0.f);
}

__global__ void stream_plain(const float* p, long long n, float* out) {
    float acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += p[3 * i] + p[3 * i + 1] + p[3 * i + 2];
    if (acc == 123.456f) out[0] = acc;
}
__global__ void stream_vec4(const float4* p, long long n4, float* out) {
    float acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = p[i];
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int TILE_BYTES, int STAGES>
__global__ void stream_bulk(const unsigned char* p, long long ntiles, float* out) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) unsigned long long s_full[STAGES], s_empty[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_empty[s])), "r"(256));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= 256) {
        if (tid != 256) return;
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) {
                const unsigned par = (it / STAGES - 1) & 1;
                asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(&s_empty[s])), "r"(par) : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_full[s])), "r"(TILE_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s_dyn + s * TILE_BYTES)),
                         "l"(p + t * TILE_BYTES), "r"(TILE_BYTES), "r"(smem_addr(&s_full[s])) : "memory");
        }
        return;
    }
    float acc = 0;
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        const unsigned par = (it / STAGES) & 1;
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(&s_full[s])), "r"(par) : "memory");
        const float* sc = reinterpret_cast<const float*>(s_dyn + s * TILE_BYTES);
        for (int j = tid; j < TILE_BYTES / 12; j += 256) acc += sc[3 * j] + sc[3 * j + 1] + sc[3 * j + 2];
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&s_empty[s])) : "memory");
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

// a bijection of [0, 2^k): multiply by odd numbers and xor-shift inside k bits
__global__ void make_perm(unsigned* perm, unsigned n, unsigned k) {
    const unsigned mask = n - 1;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned x = i;
        x = (x * 0x9E3779B1u) & mask; x ^= x >> (k / 2);
        x = (x * 0x85EBCA6Bu) & mask; x ^= x >> (k / 2 + 1);
        x = (x * 0xC2B2AE35u) & mask; x ^= x >> (k / 2);
        perm[i] = x;
    }
}
__global__ void gather4(const unsigned* src, const unsigned* perm, long long n, unsigned* out) {
    unsigned acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += src[perm[i]];
    if (acc == 0x12345u) out[0] = acc;
}
template <int BYTES>
__global__ void scatter_perm(unsigned char* dst, const unsigned* perm, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (BYTES == 8) reinterpret_cast<float2*>(dst)[perm[i]] = make_float2(1.f, 2.f);
        if (BYTES == 16) reinterpret_cast<float4*>(dst)[perm[i]] = make_float4(1.f, 2.f, 3.f, 0.f);
        if (BYTES == 32) {
            reinterpret_cast<float4*>(dst)[2 * (size_t)perm[i]] = make_float4(1.f, 2.f, 3.f, 0.f);
            reinterpret_cast<float4*>(dst)[2 * (size_t)perm[i] + 1] = make_float4(1.f, 2.f, 3.f, 0.f);
        }
    }
}

int main() {
    const long long n = 64ll * 1000000;   // "64 frames of 1M particles"
    unsigned *cnt, *out, *perm;
    float4* dst;
    float* src;
    CK(cudaMalloc(&cnt, 64u << 20));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&dst, 1200ull << 20));
    CK(cudaMalloc(&src, 12ull * n));
    CK(cudaMalloc(&perm, 4ull * n));
    CK(cudaMemset(src, 0, 12ull * n));
    CK(cudaMemset(cnt, 0, 64u << 20));
    const int grid = 148 * 8;
    printf("{\n");
    for (unsigned K : {62500u, 1000000u, 4000000u, 16000000u}) {
        float a = time_ms([&] { atomic_rand<true, 4><<<grid, 256>>>(cnt, K, n, out); });
        float b = time_ms([&] { atomic_rand<false, 4><<<grid, 256>>>(cnt, K, n, out); });
        float c = time_ms([&] { atomic_rand<true, 1><<<grid, 256>>>(cnt, K, n, out); });
        printf(" \"atomic_rand_K%u\": {\"ret_ilp4_gops\": %.1f, \"noret_ilp4_gops\": %.1f, \"ret_ilp1_gops\": %.1f},\n", K, n / a * 1e-6, n / b * 1e-6, n / c * 1e-6);
    }
    for (unsigned long long R : {16ull << 20, 64ull << 20, 960ull << 20}) {
        const unsigned R16 = (unsigned)(R / 16);
        float a = time_ms([&] { scatter16<4><<<grid, 256>>>(dst, R16, n); });
        printf(" \"scatter16_region_%lluMB\": {\"gstores\": %.1f, \"GBps\": %.0f},\n", R >> 20, n / a * 1e-6, 16.0 * n / a * 1e-6);
    }
    {
        // permutation scatter inside windows of W = 2^k records (each window is written completely, once)
        for (unsigned k : {20u, 22u, 25u}) {
            const unsigned W = 1u << k;
            make_perm<<<grid, 256>>>(perm, W, k);
            CK(cudaDeviceSynchronize());
            const long long nn = 32ll << 20;
            float a8 = time_ms([&] { for (long long o = 0; o + W <= nn; o += W) scatter_perm<8><<<grid, 256>>>((unsigned char*)dst + o * 8, perm, W); });
            float a16 = time_ms([&] { for (long long o = 0; o + W <= nn; o += W) scatter_perm<16><<<grid, 256>>>((unsigned char*)dst + o * 16, perm, W); });
            float a32 = time_ms([&] { for (long long o = 0; o + W <= nn; o += W) scatter_perm<32><<<grid, 256>>>((unsigned char*)dst + o * 32, perm, W); });
            float g = time_ms([&] { for (long long o = 0; o + W <= nn; o += W) gather4<<<grid, 256>>>((const unsigned*)dst + o, perm, W, out); });
            printf(" \"perm_window_2^%u\": {\"scatter8_gops\": %.1f, \"scatter16_gops\": %.1f, \"scatter32_gops\": %.1f, \"gather4_gops\": %.1f},\n", k, nn / a8 * 1e-6, nn / a16 * 1e-6, nn / a32 * 1e-6, nn / g * 1e-6);
        }
    }
    {
        float a = time_ms([&] { stream_plain<<<grid, 256>>>(src, n, (float*)out); });
        float b = time_ms([&] { stream_vec4<<<grid, 256>>>((const float4*)src, n * 3 / 4, (float*)out); });
        printf(" \"stream_plain_GBps\": %.0f, \"stream_vec4_GBps\": %.0f,\n", 12.0 * n / a * 1e-6, 12.0 * n / b * 1e-6);
        {
            constexpr int TB = 12288, ST = 4;
            CK(cudaFuncSetAttribute(stream_bulk<TB, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB * ST));
            for (int per : {1, 2, 4}) {
                float c = time_ms([&] { stream_bulk<TB, ST><<<148 * per, 288, TB * ST>>>((const unsigned char*)src, 12ll * n / TB, (float*)out); });
                printf(" \"stream_bulk_12KBx4_blocks%d_GBps\": %.0f,\n", per, 12.0 * n / c * 1e-6);
            }
        }
        {
            constexpr int TB = 24576, ST = 3;
            CK(cudaFuncSetAttribute(stream_bulk<TB, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB * ST));
            for (int per : {1, 2, 3}) {
                float c = time_ms([&] { stream_bulk<TB, ST><<<148 * per, 288, TB * ST>>>((const unsigned char*)src, 12ll * n / TB, (float*)out); });
                printf(" \"stream_bulk_24KBx3_blocks%d_GBps\": %.0f,\n", per, 12.0 * n / c * 1e-6);
            }
        }
    }
    printf(" \"n\": %lld\n}\n", n);
    return 0;
}

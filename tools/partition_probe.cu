// partition_probe.cu -- prototype of a cell build without a per-particle global atomic and without
// globally scattered stores: two-level counting sort of F frames of N
// random 2D particles into square cells.
//   P1  strip histogram: strips = runs of cell rows; shared-memory counters per block tile, one
//       global atomic per (tile, strip)
//   S   exclusive scan of the strip counts of every frame (one block per frame)
//   P2  strip scatter: ranks inside the tile from shared-memory atomics, one returning global atomic
//       per (tile, strip) reserves the run, records written as (x, y)
//   P3  one block per (frame, strip): cell counters in shared memory (count, scan, cursor), cell
//       offsets written coalesced, records scattered inside the strip's own region
// Prints microseconds per frame of each pass, to be compared with the 27 us of the production build
// (statistics 3 + count 9 + scans 2 + fill 11) on the same frames.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o partition_probe tools/partition_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Grid {
    double g0x, g0y, inv_h;
    int nx, ny, rows_per_strip, nstrips;
};

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void fill_random(float* p, long long n, float L) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        p[3 * i] = (hash32((unsigned)(2 * i)) >> 8) * (1.0f / 16777216.0f) * L;
        p[3 * i + 1] = (hash32((unsigned)(2 * i + 1)) >> 8) * (1.0f / 16777216.0f) * L;
        p[3 * i + 2] = 0.f;
    }
}
__device__ __forceinline__ int cell_of(double x, double g0, double inv_h, int n) {
    int c = __double2int_rd((x - g0) * inv_h);
    return min(max(c, 0), n - 1);
}

constexpr int kThreads = 256;
template <int PER>
__global__ void __launch_bounds__(kThreads) p1_strip_hist(const float* __restrict__ pts, long long n, Grid G, unsigned* __restrict__ strip_count) {
    __shared__ unsigned s_h[512];
    const int f = blockIdx.y;
    for (int k = threadIdx.x; k < G.nstrips; k += kThreads) s_h[k] = 0;
    __syncthreads();
    const float* base = pts + (long long)f * n * 3;
    const long long i0 = (long long)blockIdx.x * (kThreads * PER) + threadIdx.x;
    float y[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        y[u] = i < n ? base[3 * i + 1] : -1e30f;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        if (i < n) atomicAdd(&s_h[cell_of((double)y[u], G.g0y, G.inv_h, G.ny) / G.rows_per_strip], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < G.nstrips; k += kThreads)
        if (s_h[k]) atomicAdd(&strip_count[f * G.nstrips + k], s_h[k]);
}

__global__ void strip_scan(const unsigned* __restrict__ cnt, unsigned* __restrict__ start, unsigned* __restrict__ cursor, int nstrips, long long n) {
    const int f = blockIdx.x;
    if (threadIdx.x == 0) {
        unsigned acc = (unsigned)(f * n);
        for (int k = 0; k < nstrips; ++k) {
            start[f * (nstrips + 1) + k] = acc;
            cursor[f * nstrips + k] = acc;
            acc += cnt[f * nstrips + k];
        }
        start[f * (nstrips + 1) + nstrips] = acc;
    }
}

template <int PER>
__global__ void __launch_bounds__(kThreads) p2_strip_scatter(const float* __restrict__ pts, long long n, Grid G, unsigned* __restrict__ cursor,
                                                            float2* __restrict__ mid) {
    __shared__ unsigned s_h[512], s_base[512];
    const int f = blockIdx.y;
    for (int k = threadIdx.x; k < G.nstrips; k += kThreads) s_h[k] = 0;
    __syncthreads();
    const float* base = pts + (long long)f * n * 3;
    const long long i0 = (long long)blockIdx.x * (kThreads * PER) + threadIdx.x;
    float x[PER], y[PER];
    unsigned st[PER], rk[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        x[u] = i < n ? base[3 * i] : 0.f;
        y[u] = i < n ? base[3 * i + 1] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        st[u] = cell_of((double)y[u], G.g0y, G.inv_h, G.ny) / G.rows_per_strip;
        if (i < n) rk[u] = atomicAdd(&s_h[st[u]], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < G.nstrips; k += kThreads)
        if (s_h[k]) s_base[k] = atomicAdd(&cursor[f * G.nstrips + k], s_h[k]);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        if (i < n) mid[s_base[st[u]] + rk[u]] = make_float2(x[u], y[u]);
    }
}

// variant: the tile's records are first ordered by strip in shared memory (with their destinations), then
// written by consecutive threads -- runs of ~8 consecutive records per strip instead of 32 scattered 8-byte stores
template <int PER>
__global__ void __launch_bounds__(kThreads) p2_strip_scatter_sorted(const float* __restrict__ pts, long long n, Grid G, unsigned* __restrict__ cursor,
                                                                   float2* __restrict__ mid) {
    __shared__ unsigned s_h[512], s_base[512], s_off[512];
    __shared__ unsigned s_warp[33];
    __shared__ float2 s_rec[kThreads * PER];
    __shared__ unsigned s_dst[kThreads * PER];
    const int f = blockIdx.y;
    for (int k = threadIdx.x; k < 512; k += kThreads) s_h[k] = 0;
    __syncthreads();
    const float* base = pts + (long long)f * n * 3;
    const long long i0 = (long long)blockIdx.x * (kThreads * PER) + threadIdx.x;
    float x[PER], y[PER];
    unsigned st[PER], rk[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        x[u] = i < n ? base[3 * i] : 0.f;
        y[u] = i < n ? base[3 * i + 1] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        st[u] = cell_of((double)y[u], G.g0y, G.inv_h, G.ny) / G.rows_per_strip;
        if (i < n) rk[u] = atomicAdd(&s_h[st[u]], 1u);
    }
    __syncthreads();
    // exclusive scan of the tile's strip counts (two strips per thread) + reservation
    const unsigned c0 = s_h[2 * threadIdx.x], c1 = s_h[2 * threadIdx.x + 1];
    unsigned inc = c0 + c1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    if (c0) s_base[2 * threadIdx.x] = atomicAdd(&cursor[f * G.nstrips + 2 * threadIdx.x], c0);
    if (c1) s_base[2 * threadIdx.x + 1] = atomicAdd(&cursor[f * G.nstrips + 2 * threadIdx.x + 1], c1);
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < kThreads / 32 ? s_warp[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;
    }
    __syncthreads();
    const unsigned e0 = s_warp[warp] + inc - (c0 + c1);
    s_off[2 * threadIdx.x] = e0;
    s_off[2 * threadIdx.x + 1] = e0 + c0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + (long long)u * kThreads;
        if (i < n) {
            const unsigned pos = s_off[st[u]] + rk[u];
            s_rec[pos] = make_float2(x[u], y[u]);
            s_dst[pos] = s_base[st[u]] + rk[u];
        }
    }
    __syncthreads();
    const int cnt = (int)min((long long)(kThreads * PER), n - (long long)blockIdx.x * (kThreads * PER));
    for (int j = threadIdx.x; j < cnt; j += kThreads) mid[s_dst[j]] = s_rec[j];
}

constexpr int kP3Threads = 256;
constexpr int kP3Cache = 8;
__global__ void __launch_bounds__(kP3Threads) p3_strip_sort(const float2* __restrict__ mid, const unsigned* __restrict__ strip_start, Grid G,
                                                           long long n, unsigned* __restrict__ cell_start, float2* __restrict__ sorted) {
    extern __shared__ unsigned s_cnt[];   // [cells of the strip] counts, then cursors
    __shared__ unsigned s_warp[33];
    const int f = blockIdx.y, s = blockIdx.x;
    const int row0 = s * G.rows_per_strip;
    const int nrows = min(G.rows_per_strip, G.ny - row0);
    const int ncell = nrows * G.nx;
    const unsigned a = strip_start[f * (G.nstrips + 1) + s], b = strip_start[f * (G.nstrips + 1) + s + 1];
    for (int k = threadIdx.x; k < ncell; k += kP3Threads) s_cnt[k] = 0;
    // the first kP3Cache * kP3Threads records of the strip stay in registers for the second sweep
    float2 cv[kP3Cache];
    int cc[kP3Cache];
#pragma unroll
    for (int u = 0; u < kP3Cache; ++u) {
        const unsigned i = a + threadIdx.x + u * kP3Threads;
        cv[u] = i < b ? mid[i] : make_float2(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kP3Cache; ++u) {
        const unsigned i = a + threadIdx.x + u * kP3Threads;
        const int cx = cell_of((double)cv[u].x, G.g0x, G.inv_h, G.nx), cy = cell_of((double)cv[u].y, G.g0y, G.inv_h, G.ny);
        cc[u] = (cy - row0) * G.nx + cx;
        if (i < b) atomicAdd(&s_cnt[cc[u]], 1u);
    }
    for (unsigned i = a + threadIdx.x + kP3Cache * kP3Threads; i < b; i += kP3Threads) {
        const float2 v = mid[i];
        const int cx = cell_of((double)v.x, G.g0x, G.inv_h, G.nx), cy = cell_of((double)v.y, G.g0y, G.inv_h, G.ny);
        atomicAdd(&s_cnt[(cy - row0) * G.nx + cx], 1u);
    }
    __syncthreads();
    // exclusive scan of s_cnt[0..ncell) with base a: contiguous chunk per thread
    const int per = (ncell + kP3Threads - 1) / kP3Threads;
    const int k0 = threadIdx.x * per, k1 = min(ncell, k0 + per);
    unsigned sum = 0;
    for (int k = k0; k < k1; ++k) sum += s_cnt[k];
    unsigned inc = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < kP3Threads / 32 ? s_warp[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;
    }
    __syncthreads();
    unsigned off = a + s_warp[warp] + inc - sum;
    unsigned* cs = cell_start + (long long)f * (G.nx * G.ny + 1) + row0 * G.nx;
    for (int k = k0; k < k1; ++k) {
        const unsigned c = s_cnt[k];
        cs[k] = off;
        s_cnt[k] = off;   // cursor
        off += c;
    }
    if (s == G.nstrips - 1 && threadIdx.x == kP3Threads - 1) cell_start[(long long)f * (G.nx * G.ny + 1) + G.nx * G.ny] = b;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kP3Cache; ++u) {
        const unsigned i = a + threadIdx.x + u * kP3Threads;
        if (i < b) sorted[atomicAdd(&s_cnt[cc[u]], 1u)] = cv[u];
    }
    for (unsigned i = a + threadIdx.x + kP3Cache * kP3Threads; i < b; i += kP3Threads) {
        const float2 v = mid[i];
        const int cx = cell_of((double)v.x, G.g0x, G.inv_h, G.nx), cy = cell_of((double)v.y, G.g0y, G.inv_h, G.ny);
        sorted[atomicAdd(&s_cnt[(cy - row0) * G.nx + cx], 1u)] = v;
    }
}

int main(int argc, char** argv) {
    const long long n = 1000000;
    const int F = 32;
    const float L = 1253.3f;
    Grid G;
    G.g0x = 0; G.g0y = 0; G.inv_h = 1.0 / 5.0;
    G.nx = G.ny = 251;
    float *pts;
    float2 *mid, *sorted;
    unsigned *scnt, *sstart, *scur, *cstart;
    CK(cudaMalloc(&pts, 12ull * n * F));
    CK(cudaMalloc(&mid, 8ull * n * F));
    CK(cudaMalloc(&sorted, 8ull * n * F));
    CK(cudaMalloc(&scnt, 4 * 512 * F)); CK(cudaMalloc(&sstart, 4 * 513 * F)); CK(cudaMalloc(&scur, 4 * 512 * F));
    CK(cudaMalloc(&cstart, 4ull * (G.nx * G.ny + 1) * F));
    fill_random<<<148 * 8, 256>>>(pts, n * F, L);
    CK(cudaDeviceSynchronize());
    cudaEvent_t ev[6];
    for (auto& e : ev) cudaEventCreate(&e);
    printf("{\n");
    for (int variant = 0; variant < 2; ++variant)
    for (int rps : {1, 2}) {
        const bool sorted_scatter = variant == 1;
        G.rows_per_strip = rps;
        G.nstrips = (G.ny + rps - 1) / rps;
        const size_t smem3 = sizeof(unsigned) * (size_t)rps * G.nx;
        CK(cudaFuncSetAttribute(p3_strip_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        float best[4] = {1e30f, 1e30f, 1e30f, 1e30f};
        constexpr int PER = 8;
        const dim3 tg((unsigned)((n + kThreads * PER - 1) / (kThreads * PER)), F);
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(scnt, 0, 4 * 512 * F));
            cudaEventRecord(ev[0]);
            p1_strip_hist<PER><<<tg, kThreads>>>(pts, n, G, scnt);
            cudaEventRecord(ev[1]);
            strip_scan<<<F, 32>>>(scnt, sstart, scur, G.nstrips, n);
            cudaEventRecord(ev[2]);
            if (sorted_scatter) p2_strip_scatter_sorted<PER><<<tg, kThreads>>>(pts, n, G, scur, mid);
            else p2_strip_scatter<PER><<<tg, kThreads>>>(pts, n, G, scur, mid);
            cudaEventRecord(ev[3]);
            p3_strip_sort<<<dim3(G.nstrips, F), kP3Threads, smem3>>>(mid, sstart, G, n, cstart, sorted);
            cudaEventRecord(ev[4]);
            CK(cudaDeviceSynchronize());
            for (int k = 0; k < 4; ++k) {
                float ms;
                cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
                if (ms < best[k]) best[k] = ms;
            }
        }
        // check: cell starts monotone, last = (f + 1) n, every record sits in its cell
        std::vector<unsigned> hs((size_t)(G.nx * G.ny + 1));
        CK(cudaMemcpy(hs.data(), cstart, 4 * hs.size(), cudaMemcpyDeviceToHost));
        std::vector<float2> hv(n);
        CK(cudaMemcpy(hv.data(), sorted, 8 * n, cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int c = 0; c < G.nx * G.ny; ++c) {
            if (hs[c] > hs[c + 1]) ++bad;
            for (unsigned i = hs[c]; i < hs[c + 1] && i < n; ++i) {
                const int cx = (int)(hv[i].x / 5.0), cy = (int)(hv[i].y / 5.0);
                if (std::min(cx, G.nx - 1) != c % G.nx || std::min(cy, G.ny - 1) != c / G.nx) ++bad;
            }
        }
        if (hs[G.nx * G.ny] != n) ++bad;
        printf(" \"%srows_per_strip_%d\": {\"nstrips\": %d, \"p1_us_per_frame\": %.2f, \"scan_us\": %.2f, \"p2_us_per_frame\": %.2f, \"p3_us_per_frame\": %.2f, \"total_us_per_frame\": %.2f, \"check_errors\": %lld},\n",
               sorted_scatter ? "sorted_scatter_" : "", rps, G.nstrips, best[0] * 1e3 / F, best[1] * 1e3 / F, best[2] * 1e3 / F, best[3] * 1e3 / F,
               (best[0] + best[1] + best[2] + best[3]) * 1e3 / F, bad);
    }
    printf(" \"frames\": %d, \"n\": %lld\n}\n", F, n);
    return 0;
}

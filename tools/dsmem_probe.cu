// dsmem_probe.cu -- rate of returning atomic adds on the distributed shared memory of a thread-block
// cluster (random counter in a random block of the cluster), the mechanism a cluster-per-frame count
// pass would use instead of L2 atomics (tools/place_probe.cu: 100 G/s).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_probe tools/dsmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

constexpr int kCounters = 8192;   // per block: 32 KB

template <int ILP, bool REMOTE>
__global__ void dsmem_atomics(long long per_block, unsigned* out) {
    __shared__ unsigned cnt[kCounters];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks();
    for (int k = threadIdx.x; k < kCounters; k += blockDim.x) cnt[k] = 0;
    cluster.sync();
    unsigned acc = 0;
    const unsigned seed = blockIdx.x * 7919u;
    for (long long i = threadIdx.x; i < per_block; i += (long long)blockDim.x * ILP) {
        unsigned r[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const unsigned h = hash32(seed + (unsigned)(i + (long long)u * blockDim.x));
            const unsigned c = (h >> 8) % kCounters;
            unsigned* base = REMOTE ? cluster.map_shared_rank(cnt, h % csize) : cnt;
            r[u] = atomicAdd(base + c, 1u);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) acc += r[u];
    }
    cluster.sync();
    if (acc == 0xffffffffu) out[0] = acc;
    if (threadIdx.x == 0 && cnt[0] == 0xffffffffu) out[1] = 1;
}

template <typename K>
static float run(K kern, int cluster, int blocks, int threads, long long per_block, unsigned* out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaLaunchKernelEx(&cfg, kern, per_block, out));
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        CK(cudaLaunchKernelEx(&cfg, kern, per_block, out));
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    unsigned* out;
    CK(cudaMalloc(&out, 64));
    const long long per_block = 4ll << 20;
    printf("{\n");
    for (int cluster : {1, 2, 4, 8}) {
        const int blocks = (144 / cluster) * cluster;
        for (int threads : {256, 1024}) {
            float t = run(dsmem_atomics<4, true>, cluster, blocks, threads, per_block, out);
            printf(" \"dsmem_atomic_ret_cluster%d_threads%d_gops\": %.1f,\n", cluster, threads, blocks * (double)per_block / t * 1e-6);
        }
    }
    float t = run(dsmem_atomics<4, false>, 1, 144, 1024, per_block, out);
    printf(" \"local_smem_atomic_ret_threads1024_gops\": %.1f\n}\n", 144 * (double)per_block / t * 1e-6);
    return 0;
}

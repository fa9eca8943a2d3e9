// amepb_internal.cuh -- context, workspace and error plumbing shared by the
// translation units of libamepb.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "amepb.h"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

enum {
    BUF_STATS = 0,     // stats partials
    BUF_PLANS,         // FramePlan[]
    BUF_THR,           // threshold tables
    BUF_TCOUNT,        // target cell counters
    BUF_TSTART,        // target cell offsets
    BUF_QCOUNT,        // query cell counters
    BUF_QSTART,        // query cell offsets
    BUF_TSORTED,       // images in cell order
    BUF_QSORTED,       // queries in cell order
    BUF_QRANK,         // rank of each query inside its cell (fused self placement)
    BUF_STRIPS,        // partitioned build: strip counts | cursors | starts
    BUF_MID,           // partitioned build: records in strip order
    BUF_SCAN,          // scan block sums
    BUF_MISC,          // task counter, candidate counter
    BUF_STAGE_A,       // host-space staging: coords
    BUF_STAGE_B,       // host-space staging: other
    BUF_STAGE_OUT,     // host-space staging: results
    BUF_SQ_A,          // S(q) workspaces
    BUF_SQ_B,
    BUF_ORD_A,         // order.cu: cell counters / block counters
    BUF_ORD_B,         // order.cu: cell offsets / block offsets
    BUF_ORD_C,         // order.cu: images in cell order
    BUF_COUNT
};

struct amepb_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t total_mem = 0;
    std::string err;
    int64_t launches = 0;
    DevBuf bufs[BUF_COUNT];
    void* pinned = nullptr;  // pinned host scratch (plans, stats read-back, small tables)
    size_t pinned_cap = 0;
    // recorded after the last asynchronous device read of `pinned` (upload kernel / H2D copy);
    // every host write to `pinned` waits for it first (pinned_host_acquire)
    cudaEvent_t ev_pinned = nullptr;
    bool pinned_in_flight = false;
    amepb_rdf_info rdf_info;
    // workspace budget of amepb_rdf_batch, queried once per problem shape (rdf.cu)
    size_t budget_bytes = 0, budget_per_frame = 0;
    // per-frame statistics launched ahead for the next sub-batch (rdf.cu: stats prefetch)
    cudaEvent_t ev_stats = nullptr;
    int64_t stats_ahead_f0 = -1;   // first frame of the sub-batch whose statistics are in flight, -1: none
    int stats_ahead_nf = 0;
    // last threshold table built on the host (rdf.cu): edges it belongs to, arithmetic type, values
    std::vector<double> thr_cache_edges;
    std::vector<char> thr_cache_values;
    int thr_cache_f64 = -1;
    int rdf_k_override = 0;
    bool rdf_symmetric = true;
    int rdf_shard_index = 0, rdf_shard_count = 1;   // amepb_rdf_set_shard
    int sf2d_shard_index = 0, sf2d_shard_count = 1; // amepb_sf2d_set_shard
    int pcf2d_self_code = 0;                         // amepb_pcf2d_set_query_offset: 0 = off, else offset + 2
    long long pcfa_query_offset = -1;                // amepb_pcf_angle_set_query_offset
    void* last_stream = nullptr;
    bool sq_attr_harm = false;
    cudaStream_t copy_stream = nullptr;   // host-space staging
    cudaStream_t stats_stream = nullptr;  // host-space staging: statistics of the next sub-batch (rdf.cu)
    cudaStream_t compute_stream = nullptr; // host-space calls made on the NULL stream
    static constexpr int kStageBuffers = 3;   // host-buffer path: staging buffers in flight
    cudaEvent_t ev_copied[kStageBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_computed[kStageBuffers] = {nullptr, nullptr, nullptr};
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;     // reusable events
    size_t ev_used = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pair, ev_build, ev_sq, ev_copy;
    bool sq_attr_sf2d = false;
    bool pcf_attr = false;   // dynamic shared-memory opt-in of the pcf2d cell kernels set
};

namespace amepb {

int fail(amepb_ctx* ctx, int code, const char* fmt, ...);
int fail_cuda(amepb_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);
int ensure_device(amepb_ctx* ctx, int which, size_t bytes);
int ensure_pinned(amepb_ctx* ctx, size_t bytes);
// The pinned scratch block is shared by asynchronous consumers.  Host code calls
// pinned_host_acquire() before it writes into ctx->pinned (blocks until the last recorded device
// read has run) and pinned_device_release() right after it has enqueued a device read of it.
int pinned_host_acquire(amepb_ctx* ctx);
int pinned_device_release(amepb_ctx* ctx, cudaStream_t stream);

cudaEvent_t timing_event(amepb_ctx* ctx, cudaStream_t stream);  // records and returns an event (or nullptr)
double timing_total_ms(amepb_ctx* ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& spans);
void timing_reset(amepb_ctx* ctx);

// shared by rdf.cu and order.cu
// exclusive prefix sum of `len` counters (device arrays, in != out), on `stream`
int exclusive_scan(amepb_ctx* ctx, const unsigned* in, unsigned* out, long long len, cudaStream_t stream);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace amepb

#define AMEPB_CUDA(ctx, call)                                                         \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) return amepb::fail_cuda(ctx, e__, #call, __FILE__, __LINE__); \
    } while (0)

#define AMEPB_LAUNCH_CHECK(ctx)                                                       \
    do {                                                                              \
        (ctx)->launches++;                                                            \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) return amepb::fail_cuda(ctx, e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

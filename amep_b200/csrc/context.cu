// context.cu -- context life cycle, error text, workspace growth.
#include <stdarg.h>

#include "amepb_internal.cuh"

namespace {
thread_local std::string g_create_error;
}

namespace amepb {

int fail(amepb_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

int fail_cuda(amepb_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    const char* base = file;
    for (const char* p = file; *p; ++p)
        if (*p == '/') base = p + 1;
    fail(ctx, (int)e, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base, line, what);
    return (int)e > 0 ? (int)e : 1;
}

int ensure_device(amepb_ctx* ctx, int which, size_t bytes) {
    DevBuf& b = ctx->bufs[which];
    if (b.cap >= bytes && b.p) return 0;
    if (b.p) {
        AMEPB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return fail(ctx, AMEPB_E_NOMEM, "device workspace %d of %zu bytes does not fit", which, bytes);
    }
    b.cap = want;
    return 0;
}

int pinned_host_acquire(amepb_ctx* ctx) {
    if (ctx->pinned_in_flight && ctx->ev_pinned) {
        AMEPB_CUDA(ctx, cudaEventSynchronize(ctx->ev_pinned));
        ctx->pinned_in_flight = false;
    }
    return 0;
}

int pinned_device_release(amepb_ctx* ctx, cudaStream_t stream) {
    if (!ctx->ev_pinned) AMEPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_pinned, cudaEventDisableTiming));
    AMEPB_CUDA(ctx, cudaEventRecord(ctx->ev_pinned, stream));
    ctx->pinned_in_flight = true;
    return 0;
}

int ensure_pinned(amepb_ctx* ctx, size_t bytes) {
    if (ctx->pinned_cap >= bytes && ctx->pinned) return 0;
    if (ctx->pinned) {
        // a queued device read of the old block must have run before it is freed
        int rc = pinned_host_acquire(ctx);
        if (rc) return rc;
        AMEPB_CUDA(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    AMEPB_CUDA(ctx, cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_cap = want;
    return 0;
}

cudaEvent_t timing_event(amepb_ctx* ctx, cudaStream_t stream) {
    if (!ctx->timing) return nullptr;
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->ev_pool.push_back(e);
    }
    cudaEvent_t e = ctx->ev_pool[ctx->ev_used++];
    cudaEventRecord(e, stream);
    return e;
}

double timing_total_ms(amepb_ctx* ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& spans) {
    double total = 0.0;
    for (auto& sp : spans) {
        if (!sp.first || !sp.second) continue;
        cudaEventSynchronize(sp.second);
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.first, sp.second) == cudaSuccess) total += ms;
    }
    return total;
}

void timing_reset(amepb_ctx* ctx) {
    ctx->ev_used = 0;
    ctx->ev_pair.clear();
    ctx->ev_build.clear();
    ctx->ev_sq.clear();
    ctx->ev_copy.clear();
}

}  // namespace amepb

extern "C" {

int amepb_abi_version(void) { return AMEPB_ABI_VERSION; }

int amepb_create(int device, amepb_ctx** out) {
    if (!out) return amepb::fail(nullptr, AMEPB_E_INVAL, "amepb_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return amepb::fail(nullptr, AMEPB_E_NODEVICE,
                           "amepb_create: no CUDA device (%s); libamepb has no CPU fallback",
                           e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= count)
        return amepb::fail(nullptr, AMEPB_E_INVAL, "amepb_create: device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return amepb::fail_cuda(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    if (prop.major < 10)
        return amepb::fail(nullptr, AMEPB_E_NODEVICE,
                           "amepb_create: device %d is sm_%d%d; libamepb is built for sm_100a (B200) only",
                           device, prop.major, prop.minor);
    amepb_ctx* ctx = new amepb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->total_mem = prop.totalGlobalMem;
    memset(&ctx->rdf_info, 0, sizeof(ctx->rdf_info));
    *out = ctx;
    return 0;
}

void amepb_destroy(amepb_ctx* ctx) {
    if (!ctx) return;
    amepb::DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < BUF_COUNT; ++i)
        if (ctx->bufs[i].p) cudaFree(ctx->bufs[i].p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_stats) cudaEventDestroy(ctx->ev_stats);
    if (ctx->ev_pinned) cudaEventDestroy(ctx->ev_pinned);
    for (int b = 0; b < amepb_ctx::kStageBuffers; ++b) {
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_computed[b]) cudaEventDestroy(ctx->ev_computed[b]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stats_stream) cudaStreamDestroy(ctx->stats_stream);
    if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
    delete ctx;
}

const char* amepb_last_error(const amepb_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int64_t amepb_kernel_launches(const amepb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int amepb_set_timing(amepb_ctx* ctx, int enable) {
    if (!ctx) return AMEPB_E_INVAL;
    ctx->timing = enable != 0;
    return 0;
}

double amepb_last_kernel_ms(amepb_ctx* ctx) {
    if (!ctx) return 0.0;
    amepb::DeviceGuard guard(ctx->device);
    return amepb::timing_total_ms(ctx, ctx->ev_sq);
}

int amepb_synchronize(amepb_ctx* ctx, void* stream) {
    if (!ctx) return AMEPB_E_INVAL;
    amepb::DeviceGuard guard(ctx->device);
    AMEPB_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

}  // extern "C"

// Particle counts on the grid of the histogram density estimate behind sf2d(mode='fft').
//
// Reference: amep/continuum.py:300-407 (hde: periodic images, np.arange edges,
// np.histogram2d) called from coords_to_density (amep/continuum.py:51-139) by
// amep/spatialcor.py:1911-1940.  The kernel reproduces the integer counts exactly:
//   * images as pbc_points(width=0.5, enforce_nd=2): float32 base + float32 shift along x and
//     y, kept iff strictly inside (centre - L, centre + L) in every dimension
//     (amep/pbc.py:270-338); without pbc the coordinates are binned as they are;
//   * bin i of an axis holds edges[i] <= v < edges[i+1], the last bin also its right edge,
//     values outside [edges[0], edges[-1]] and NaN are dropped (np.histogramdd with explicit
//     edges: searchsorted(side='right') and the on-edge correction), compared in float64.
// Everything after the counts (division by the cell area, FFT, normalisation, cropping) is
// array arithmetic on the grid and stays with the caller.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "amepb_internal.cuh"
#include "rdf_plan.h"

namespace amepb {

struct DensityFrame {
    double win_lo[3], win_hi[3];  // strict image window (+-inf without pbc)
    float shift[3];               // float32 box lengths
    int pbc;
};

// index of the bin holding v, or -1 (np.histogramdd semantics, see the header comment)
__device__ __forceinline__ int bin_of(double v, const double* __restrict__ e, int ne, double inv_d) {
    const double lo = e[0], hi = e[ne - 1];
    if (!(v >= lo) || !(v <= hi)) return -1;
    if (v == hi) return ne - 2;
    int g = __double2int_rd((v - lo) * inv_d);
    g = min(max(g, 0), ne - 2);
    while (v < e[g]) --g;
    while (v >= e[g + 1]) ++g;
    return g;
}

template <typename T>
__global__ void __launch_bounds__(256) density_counts_kernel(const T* __restrict__ pts, long long n,
                                                             const DensityFrame* __restrict__ frames,
                                                             const double* __restrict__ xe, int nxe,
                                                             const double* __restrict__ ye, int nye,
                                                             double inv_dx, double inv_dy,
                                                             unsigned* __restrict__ counts) {
    const int f = blockIdx.y;
    __shared__ DensityFrame F;
    if (threadIdx.x < sizeof(DensityFrame) / sizeof(int))
        reinterpret_cast<int*>(&F)[threadIdx.x] = reinterpret_cast<const int*>(frames + f)[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* p = pts + ((long long)f * n + i) * 3;
    unsigned* grid = counts + (size_t)f * (nye - 1) * (nxe - 1);
    if (!F.pbc) {
        const int bx = bin_of((double)p[0], xe, nxe, inv_dx), by = bin_of((double)p[1], ye, nye, inv_dy);
        if (bx >= 0 && by >= 0) atomicAdd(&grid[(size_t)by * (nxe - 1) + bx], 1u);
        return;
    }
    // z is not shifted (enforce_nd = 2) but takes part in the window test
    const double z = (double)(float)p[2];
    if (!(z < F.win_hi[2]) || !(z > F.win_lo[2])) return;
    int bxs[3], bys[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const float vf = (v == 0) ? 0.0f : (v == 1 ? -1.0f : 1.0f);
        const double x = (double)__fadd_rn((float)p[0], __fmul_rn(vf, F.shift[0]));
        const double y = (double)__fadd_rn((float)p[1], __fmul_rn(vf, F.shift[1]));
        bxs[v] = (x < F.win_hi[0] && x > F.win_lo[0]) ? bin_of(x, xe, nxe, inv_dx) : -1;
        bys[v] = (y < F.win_hi[1] && y > F.win_lo[1]) ? bin_of(y, ye, nye, inv_dy) : -1;
    }
#pragma unroll
    for (int vy = 0; vy < 3; ++vy) {
        if (bys[vy] < 0) continue;
#pragma unroll
        for (int vx = 0; vx < 3; ++vx)
            if (bxs[vx] >= 0) atomicAdd(&grid[(size_t)bys[vy] * (nxe - 1) + bxs[vx]], 1u);
    }
}

}  // namespace amepb

using namespace amepb;

extern "C" int amepb_density_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space,
                                    int64_t nframes, const double* box_boundary, int box_dtype, int pbc,
                                    const double* xedges, int32_t nxe, const double* yedges, int32_t nye,
                                    uint32_t* counts, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !box_boundary || !xedges || !yedges || !counts || n <= 0 || nframes < 0 || nxe < 2 || nye < 2 ||
        (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (space != AMEPB_DEVICE && space != AMEPB_HOST) ||
        nframes > 65535)
        return fail(ctx, AMEPB_E_INVAL, "amepb_density_counts: invalid argument");
    for (int i = 0; i + 1 < nxe; ++i)
        if (!(xedges[i] < xedges[i + 1])) return fail(ctx, AMEPB_E_INVAL, "amepb_density_counts: x edges must increase");
    for (int i = 0; i + 1 < nye; ++i)
        if (!(yedges[i] < yedges[i + 1])) return fail(ctx, AMEPB_E_INVAL, "amepb_density_counts: y edges must increase");
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const size_t esz = dtype == AMEPB_F32 ? 4 : 8;
    const size_t cbytes = (size_t)n * 3 * esz * nframes;
    const size_t gbytes = sizeof(unsigned) * (size_t)nframes * (nxe - 1) * (nye - 1);
    int rc;
    const void* d_coords = coords;
    unsigned* d_counts = counts;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, cbytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, cbytes, cudaMemcpyHostToDevice, stream));
        d_coords = ctx->bufs[BUF_STAGE_A].p;
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, gbytes))) return rc;
        d_counts = reinterpret_cast<unsigned*>(ctx->bufs[BUF_STAGE_OUT].p);
    }
    // small parameter block: per-frame windows, then the two edge arrays
    const size_t fbytes = (sizeof(DensityFrame) * (size_t)nframes + 15) & ~(size_t)15;
    const size_t pbytes = fbytes + sizeof(double) * ((size_t)nxe + nye);
    if ((rc = ensure_pinned(ctx, pbytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, pbytes))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;   // an earlier call may still be reading the block
    char* hp = reinterpret_cast<char*>(ctx->pinned);
    DensityFrame* hf = reinterpret_cast<DensityFrame*>(hp);
    for (int64_t f = 0; f < nframes; ++f) {
        DensityFrame& F = hf[f];
        F.pbc = pbc ? 1 : 0;
        if (pbc) {
            image_window(box_boundary + f * 6, box_dtype == AMEPB_F32, F.win_lo, F.win_hi, F.shift);
        } else {
            for (int d = 0; d < 3; ++d) {
                F.win_lo[d] = -INFINITY;
                F.win_hi[d] = INFINITY;
                F.shift[d] = 0.0f;
            }
        }
    }
    memcpy(hp + fbytes, xedges, sizeof(double) * nxe);
    memcpy(hp + fbytes + sizeof(double) * nxe, yedges, sizeof(double) * nye);
    char* dp = reinterpret_cast<char*>(ctx->bufs[BUF_SQ_A].p);
    AMEPB_CUDA(ctx, cudaMemcpyAsync(dp, hp, pbytes, cudaMemcpyHostToDevice, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, gbytes, stream));
    const double* d_xe = reinterpret_cast<const double*>(dp + fbytes);
    const double* d_ye = d_xe + nxe;
    const double inv_dx = (double)(nxe - 1) / (xedges[nxe - 1] - xedges[0]);
    const double inv_dy = (double)(nye - 1) / (yedges[nye - 1] - yedges[0]);
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)nframes);
    if (dtype == AMEPB_F32)
        density_counts_kernel<float><<<grid, 256, 0, stream>>>((const float*)d_coords, n, reinterpret_cast<const DensityFrame*>(dp),
                                                               d_xe, nxe, d_ye, nye, inv_dx, inv_dy, d_counts);
    else
        density_counts_kernel<double><<<grid, 256, 0, stream>>>((const double*)d_coords, n, reinterpret_cast<const DensityFrame*>(dp),
                                                                d_xe, nxe, d_ye, nye, inv_dx, inv_dy, d_counts);
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, gbytes, cudaMemcpyDeviceToHost, stream));
    }
    // the pinned parameter block is reused by the next call
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    return 0;
}

// Pair counts on the (dx, dy) grid behind spatialcor.pcf2d.
//
// Reference: amep/spatialcor.py:652-723 (__dhist2d) driven by amep/spatialcor.py:726-926 (pcf2d).
// For every particle n of other_coords the reference takes the plain difference vectors
// other_coords[n] - coords (no minimum image: the pbc_diff calls are commented out there, and
// the pbc argument is ignored), drops the entry with the same index when both sets are the
// same array, optionally rotates the difference vectors by -angle about the box centre
// (amep/utils.py:581-652 rotate_coords: subtract the centre, R.v, add the centre) and adds
// np.histogram2d(dx, dy, [xbins, ybins]).  The kernel reproduces the integer counts:
//   * the difference is taken in the promoted dtype of the two arrays (float32 - float32 stays
//     float32), "diff - centre" in the promoted dtype of that and the centre (float32 when the
//     box is float32), the rotation and "+ centre" in float64, products and sums rounded
//     separately;
//   * bin i of an axis holds edges[i] <= v < edges[i+1], the last bin also its right edge, NaN
//     and outside values are dropped, compared in float64 (np.histogramdd with explicit edges).
// Everything after the counts (bin areas, density, N_other, bin centres) is array arithmetic
// on the grid and stays with the caller.
//
// Layout: one thread per query particle (a row of other_coords), targets (rows of coords) staged
// through shared memory 1024 at a time and read by all threads at once (broadcast); the targets
// are split over blockIdx.y so that small systems still fill the SMs; counts are 64-bit and
// added with global atomics (a 500 x 500 grid does not fit shared memory).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "amepb_internal.cuh"
#include "rdf_plan.h"

namespace amepb {

struct Pcf2dFrame {
    double cx, cy;     // box centre (x, y)
    double r00, r01;   // first row of R(-angle):  cos, -sin
    double r10, r11;   // second row:              sin,  cos
    int rotate;        // angle != 0
    // float32 prefilter for float32 differences.  No rotation: the closed float32 interval that is
    // equivalent to the float64 test against the outer edges; rotation: squared radius about the
    // centre that holds the whole window (conservative: never rejects a pair that bins)
    float fxlo, fxhi, fylo, fyhi, fr2;
};

constexpr int kPcfThreads = 256;
constexpr int kPcfStage = 1024;

// index of the bin holding v, or -1 (np.histogramdd semantics)
__device__ __forceinline__ int pcf_bin_of(double v, const double* __restrict__ e, int ne, double inv_d) {
    const double lo = e[0], hi = e[ne - 1];
    if (!(v >= lo) || !(v <= hi)) return -1;
    if (v == hi) return ne - 2;
    int g = __double2int_rd((v - lo) * inv_d);
    g = min(max(g, 0), ne - 2);
    while (v < e[g]) --g;
    while (v >= e[g + 1]) ++g;
    return g;
}

// D: dtype of other_coords[n] - coords;  CF32: the centre is float32 (only matters if D is float)
template <typename TC, typename TO, typename D, bool CF32>
__global__ void __launch_bounds__(kPcfThreads) pcf2d_kernel(const TC* __restrict__ coords, long long n,
                                                            const TO* __restrict__ other, long long m, int same,
                                                            const Pcf2dFrame* __restrict__ frames,
                                                            const double* __restrict__ xedges, int nxe,
                                                            const double* __restrict__ yedges, int nye,
                                                            long long edge_stride, long long per_split,
                                                            const float* __restrict__ thr,
                                                            unsigned long long* __restrict__ counts) {
    extern __shared__ float s_thr[];   // float32 thresholds of the x edges, then of the y edges
    __shared__ D sx[kPcfStage], sy[kPcfStage];
    __shared__ Pcf2dFrame F;
    const int f = blockIdx.z;
    if (threadIdx.x < sizeof(Pcf2dFrame) / sizeof(int))
        reinterpret_cast<int*>(&F)[threadIdx.x] = reinterpret_cast<const int*>(frames + f)[threadIdx.x];
    const double* xe = xedges + (long long)f * edge_stride;
    const double* ye = yedges + (long long)f * edge_stride;
    const double xlo = xe[0], xhi = xe[nxe - 1], ylo = ye[0], yhi = ye[nye - 1];
    const double inv_dx = (double)(nxe - 1) / (xhi - xlo), inv_dy = (double)(nye - 1) / (yhi - ylo);
    const TC* cf = coords + (long long)f * n * 3;
    const TO* of = other + (long long)f * m * 3;
    unsigned long long* grid = counts + (size_t)f * (nxe - 1) * (nye - 1);
    const long long qi = (long long)blockIdx.x * kPcfThreads + threadIdx.x;
    const bool live = qi < m;
    const D qx = live ? (D)of[qi * 3 + 0] : (D)0, qy = live ? (D)of[qi * 3 + 1] : (D)0;
    const long long t_lo = (long long)blockIdx.y * per_split, t_hi = min(n, t_lo + per_split);
    // float32 differences without rotation: binning against float32 thresholds (thr[k] = smallest
    // float32 >= edges[k], so "v >= edges[k]" in float64 is "v >= thr[k]" in float32)
    const bool use_thr = sizeof(D) == 4 && thr != nullptr;
    if (use_thr) {
        const float* tf = thr + (long long)f * (nxe + nye);
        for (int i = threadIdx.x; i < nxe + nye; i += kPcfThreads) s_thr[i] = tf[i];
    }
    __syncthreads();
    const bool rotate = F.rotate != 0;
    const float* tx = s_thr;
    const float* ty = s_thr + nxe;
    const float finv_dx = (float)inv_dx, finv_dy = (float)inv_dy;
    for (long long t0 = t_lo; t0 < t_hi; t0 += kPcfStage) {
        const int cnt = (int)min((long long)kPcfStage, t_hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += kPcfThreads) {
            sx[i] = (D)cf[(t0 + i) * 3 + 0];
            sy[i] = (D)cf[(t0 + i) * 3 + 1];
        }
        __syncthreads();
        if (!live) continue;
        const int skip = (same && qi >= t0 && qi < t0 + cnt) ? (int)(qi - t0) : -1;
        for (int i = 0; i < cnt; ++i) {
            D dx, dy;
            if constexpr (sizeof(D) == 4) {
                dx = __fsub_rn(qx, sx[i]);
                dy = __fsub_rn(qy, sy[i]);
            } else {
                dx = __dsub_rn(qx, sx[i]);
                dy = __dsub_rn(qy, sy[i]);
            }
            double x, y;
            if constexpr (sizeof(D) == 4) {
                if (!rotate) {
                    if (!(dx >= F.fxlo) || !(dx <= F.fxhi) || !(dy >= F.fylo) || !(dy <= F.fyhi)) continue;
                    if (use_thr) {
                        if (i == skip) continue;
                        // guess from the uniform spacing, then settle against the thresholds
                        int bx = min(max(__float2int_rd((dx - F.fxlo) * finv_dx), 0), nxe - 2);
                        int by = min(max(__float2int_rd((dy - F.fylo) * finv_dy), 0), nye - 2);
                        while (dx < tx[bx]) --bx;
                        while (bx < nxe - 2 && dx >= tx[bx + 1]) ++bx;
                        while (dy < ty[by]) --by;
                        while (by < nye - 2 && dy >= ty[by + 1]) ++by;
                        atomicAdd(&grid[(size_t)bx * (nye - 1) + by], 1ULL);
                        continue;
                    }
                } else {
                    const float ux = dx - (float)F.cx, uy = dy - (float)F.cy;
                    if (!(ux * ux + uy * uy <= F.fr2)) continue;
                }
            }
            if (!rotate) {
                x = (double)dx;
                y = (double)dy;
                if (!(x >= xlo) || !(x <= xhi) || !(y >= ylo) || !(y <= yhi)) continue;
            } else {
                double vx, vy;
                if constexpr (sizeof(D) == 4 && CF32) {
                    vx = (double)__fsub_rn(dx, (float)F.cx);
                    vy = (double)__fsub_rn(dy, (float)F.cy);
                } else {
                    vx = __dsub_rn((double)dx, F.cx);
                    vy = __dsub_rn((double)dy, F.cy);
                }
                x = __dadd_rn(__dadd_rn(__dmul_rn(F.r00, vx), __dmul_rn(F.r01, vy)), F.cx);
                y = __dadd_rn(__dadd_rn(__dmul_rn(F.r10, vx), __dmul_rn(F.r11, vy)), F.cy);
                if (!(x >= xlo) || !(x <= xhi) || !(y >= ylo) || !(y <= yhi)) continue;
            }
            if (i == skip) continue;
            const int bx = pcf_bin_of(x, xe, nxe, inv_dx);
            const int by = pcf_bin_of(y, ye, nye, inv_dy);
            if (bx >= 0 && by >= 0) atomicAdd(&grid[(size_t)bx * (nye - 1) + by], 1ULL);
        }
    }
}

template <typename TC, typename TO>
static void launch_pcf2d(dim3 grid, cudaStream_t stream, bool center_f32, const void* coords, long long n,
                         const void* other, long long m, int same, const Pcf2dFrame* frames, const double* xe, int nxe,
                         const double* ye, int nye, long long edge_stride, long long per_split, const float* thr,
                         size_t smem, unsigned long long* counts) {
    constexpr bool both_f32 = sizeof(TC) == 4 && sizeof(TO) == 4;
    if constexpr (both_f32) {
        if (center_f32)
            pcf2d_kernel<TC, TO, float, true><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, counts);
        else
            pcf2d_kernel<TC, TO, float, false><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, counts);
    } else {
        pcf2d_kernel<TC, TO, double, false><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, counts);
    }
}

}  // namespace amepb

using namespace amepb;

extern "C" int amepb_pcf2d_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other,
                                  int other_dtype, int64_t m, int same, int space, int64_t nframes,
                                  const double* center, int center_f32, const double* rot, const double* xedges,
                                  int32_t nxe, const double* yedges, int32_t nye, int64_t edge_stride,
                                  uint64_t* counts, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !other || !center || !xedges || !yedges || !counts || n <= 0 || m <= 0 || nframes < 0 || nxe < 2 ||
        nye < 2 || (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (other_dtype != AMEPB_F32 && other_dtype != AMEPB_F64) ||
        (space != AMEPB_DEVICE && space != AMEPB_HOST) || nframes > 65535 ||
        (edge_stride != 0 && (edge_stride < nxe || edge_stride < nye)) || (same && (n != m || dtype != other_dtype)))
        return fail(ctx, AMEPB_E_INVAL, "amepb_pcf2d_counts: invalid argument");
    const int64_t nedge_sets = edge_stride ? nframes : 1;
    for (int64_t f = 0; f < nedge_sets; ++f) {
        const double* xe = xedges + f * edge_stride;
        const double* ye = yedges + f * edge_stride;
        for (int i = 0; i + 1 < nxe; ++i)
            if (!(xe[i] < xe[i + 1])) return fail(ctx, AMEPB_E_INVAL, "amepb_pcf2d_counts: x edges must increase");
        for (int i = 0; i + 1 < nye; ++i)
            if (!(ye[i] < ye[i + 1])) return fail(ctx, AMEPB_E_INVAL, "amepb_pcf2d_counts: y edges must increase");
    }
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const size_t cbytes = (size_t)n * 3 * (dtype == AMEPB_F32 ? 4 : 8) * nframes;
    const size_t obytes = (size_t)m * 3 * (other_dtype == AMEPB_F32 ? 4 : 8) * nframes;
    const size_t gbytes = sizeof(unsigned long long) * (size_t)nframes * (nxe - 1) * (nye - 1);
    int rc;
    const void* d_coords = coords;
    const void* d_other = other;
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(counts);
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, cbytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, cbytes, cudaMemcpyHostToDevice, stream));
        d_coords = ctx->bufs[BUF_STAGE_A].p;
        if (same || other == coords) {
            d_other = d_coords;
        } else {
            if ((rc = ensure_device(ctx, BUF_STAGE_B, obytes))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_B].p, other, obytes, cudaMemcpyHostToDevice, stream));
            d_other = ctx->bufs[BUF_STAGE_B].p;
        }
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, gbytes))) return rc;
        d_counts = reinterpret_cast<unsigned long long*>(ctx->bufs[BUF_STAGE_OUT].p);
    }
    // small parameter block: per-frame centre / rotation, then the x and y edge arrays
    const size_t fbytes = (sizeof(Pcf2dFrame) * (size_t)nframes + 15) & ~(size_t)15;
    const size_t estride = edge_stride ? (size_t)edge_stride : (size_t)std::max(nxe, nye);
    const size_t ebytes = sizeof(double) * estride * nedge_sets;
    // float32 thresholds of the edges, per frame (x then y), used by the float32 kernels when they
    // fit shared memory beside the target stage
    const bool all_f32 = dtype == AMEPB_F32 && other_dtype == AMEPB_F32;
    const bool with_thr = all_f32 && (size_t)nxe + nye <= 6144;
    const size_t tbytes = with_thr ? ((sizeof(float) * (size_t)nframes * (nxe + nye) + 15) & ~(size_t)15) : 0;
    const size_t pbytes = fbytes + 2 * ebytes + tbytes;
    if ((rc = ensure_pinned(ctx, pbytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, pbytes))) return rc;
    char* hp = reinterpret_cast<char*>(ctx->pinned);
    memset(hp, 0, pbytes);
    Pcf2dFrame* hf = reinterpret_cast<Pcf2dFrame*>(hp);
    for (int64_t f = 0; f < nframes; ++f) {
        Pcf2dFrame& F = hf[f];
        F.cx = center[f * 3 + 0];
        F.cy = center[f * 3 + 1];
        F.rotate = (rot && rot[f * 3 + 0] != 0.0) ? 1 : 0;
        const double c = rot ? rot[f * 3 + 1] : 1.0, s = rot ? rot[f * 3 + 2] : 0.0;
        F.r00 = c;
        F.r01 = -s;
        F.r10 = s;
        F.r11 = c;
        const double* xe = xedges + f * edge_stride;
        const double* ye = yedges + f * edge_stride;
        F.fxlo = float_above(xe[0], false);      // dx >= xlo in float64  <=>  dx >= fxlo in float32
        F.fxhi = float_below(xe[nxe - 1], false);
        F.fylo = float_above(ye[0], false);
        F.fyhi = float_below(ye[nye - 1], false);
        // |R v| = |v|: the rotated vector lies in the window only if |diff - centre| is at most
        // the distance from the centre to the farthest window corner (plus float32 slack: the
        // prefilter subtracts a float32-rounded centre and rounds its products)
        const double ax = std::max(fabs(xe[0] - F.cx), fabs(xe[nxe - 1] - F.cx));
        const double ay = std::max(fabs(ye[0] - F.cy), fabs(ye[nye - 1] - F.cy));
        const double rad = sqrt(ax * ax + ay * ay) * (1.0 + 1e-5) + 1e-5 * (fabs(F.cx) + fabs(F.cy)) + 1e-30;
        F.fr2 = float_above(rad * rad, false);
    }
    double* hxe = reinterpret_cast<double*>(hp + fbytes);
    double* hye = reinterpret_cast<double*>(hp + fbytes + ebytes);
    for (int64_t f = 0; f < nedge_sets; ++f) {
        memcpy(hxe + f * estride, xedges + f * edge_stride, sizeof(double) * nxe);
        memcpy(hye + f * estride, yedges + f * edge_stride, sizeof(double) * nye);
    }
    if (with_thr) {
        float* ht = reinterpret_cast<float*>(hp + fbytes + 2 * ebytes);
        for (int64_t f = 0; f < nframes; ++f) {
            const double* xe = xedges + f * edge_stride;
            const double* ye = yedges + f * edge_stride;
            float* t = ht + f * (nxe + nye);
            for (int i = 0; i < nxe; ++i) t[i] = float_above(xe[i], false);
            for (int i = 0; i < nye; ++i) t[nxe + i] = float_above(ye[i], false);
        }
    }
    char* dp = reinterpret_cast<char*>(ctx->bufs[BUF_SQ_A].p);
    AMEPB_CUDA(ctx, cudaMemcpyAsync(dp, hp, pbytes, cudaMemcpyHostToDevice, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, gbytes, stream));
    const Pcf2dFrame* d_frames = reinterpret_cast<const Pcf2dFrame*>(dp);
    const double* d_xe = reinterpret_cast<const double*>(dp + fbytes);
    const double* d_ye = reinterpret_cast<const double*>(dp + fbytes + ebytes);
    const long long dev_stride = edge_stride ? (long long)estride : 0;
    const float* d_thr = with_thr ? reinterpret_cast<const float*>(dp + fbytes + 2 * ebytes) : nullptr;
    const size_t smem = with_thr ? sizeof(float) * ((size_t)nxe + nye) : 0;

    // query tiles x target splits x frames: at least ~4 blocks per SM when the system allows it
    const long long qtiles = (m + kPcfThreads - 1) / kPcfThreads;
    const long long want = 4LL * ctx->sm_count;
    long long tsplit = (want + qtiles * nframes - 1) / (qtiles * nframes);
    const long long max_split = (n + kPcfStage - 1) / kPcfStage;
    tsplit = std::max(1LL, std::min({tsplit, max_split, 65535LL}));
    long long per_split = (n + tsplit - 1) / tsplit;
    per_split = (per_split + kPcfStage - 1) / kPcfStage * kPcfStage;
    tsplit = (n + per_split - 1) / per_split;
    const dim3 grid((unsigned)qtiles, (unsigned)tsplit, (unsigned)nframes);
    const bool cf32 = center_f32 != 0;
    if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32)
        launch_pcf2d<float, float>(grid, stream, cf32, d_coords, n, d_other, m, same, d_frames, d_xe, nxe, d_ye, nye, dev_stride, per_split, d_thr, smem, d_counts);
    else if (dtype == AMEPB_F32)
        launch_pcf2d<float, double>(grid, stream, cf32, d_coords, n, d_other, m, same, d_frames, d_xe, nxe, d_ye, nye, dev_stride, per_split, d_thr, smem, d_counts);
    else if (other_dtype == AMEPB_F32)
        launch_pcf2d<double, float>(grid, stream, cf32, d_coords, n, d_other, m, same, d_frames, d_xe, nxe, d_ye, nye, dev_stride, per_split, d_thr, smem, d_counts);
    else
        launch_pcf2d<double, double>(grid, stream, cf32, d_coords, n, d_other, m, same, d_frames, d_xe, nxe, d_ye, nye, dev_stride, per_split, d_thr, smem, d_counts);
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, gbytes, cudaMemcpyDeviceToHost, stream));
    }
    // the pinned parameter block is reused by the next call
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    return 0;
}

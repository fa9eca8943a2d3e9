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
// Two pair walks share one binning routine (PairBinner) and give identical counts:
//   * pcf2d_kernel (direct): one thread per query particle (a row of other_coords), targets (rows
//     of coords) staged through shared memory 1024 at a time and read by all threads at once
//     (broadcast); the targets are split over blockIdx.y so that small systems still fill the
//     SMs; counts are 64-bit and added with global atomics (a 500 x 500 grid does not fit shared
//     memory).  Best when few pairs bin (narrow window).
//   * pcf2d_cells_kernel (large systems, wide window): both sets counting-sorted into cells, a
//     block counts a (query cell, target cell) pair in a shared-memory window of the grid; see the
//     comment above that kernel.
// The rotation flag is a template parameter (all frames of a launch share it); float32
// differences are decided in float32 wherever that provably agrees with the float64 arithmetic
// of the reference (thresholds for unrotated pairs, an error margin for rotated ones).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <type_traits>
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
    // both edge arrays are uniform to within 0.2 bin widths and their float32 thresholds strictly
    // increase: the bin guessed from the spacing is off by at most one
    int uniform;
    // float32 copies of the rotation and the centre, and the margin within which the float32
    // evaluation of a rotated difference is guaranteed to lie of the float64 one (see bins())
    float fr00, fr01, fr10, fr11, fcx, fcy, fmargin;
};

constexpr int kPcfThreads = 256;
constexpr int kPcfStage = 1024;
constexpr size_t kPcfStageEdges = 2048;   // edges (x + y) up to this many are staged in shared memory

// index of the bin holding v, or -1 (np.histogramdd semantics)
__device__ __forceinline__ int pcf_bin_of(double v, const double* __restrict__ e, int ne, double inv_d, bool uniform) {
    const double lo = e[0], hi = e[ne - 1];
    if (!(v >= lo) || !(v <= hi)) return -1;
    if (v == hi) return ne - 2;
    int g = __double2int_rd((v - lo) * inv_d);
    g = min(max(g, 0), ne - 2);
    if (uniform) return g + (int)(g < ne - 2 && v >= e[g + 1]) - (int)(v < e[g]);   // off by one at most
    while (v < e[g]) --g;
    while (v >= e[g + 1]) ++g;
    return g;
}

// Bins one difference vector exactly as the reference does (see the header comment).
// D: dtype of other_coords[n] - coords;  CF32: the centre is float32 (only matters if D is float)
template <typename D, bool CF32, bool ROT>
struct PairBinner {
    const Pcf2dFrame* F;
    const double* xe;
    const double* ye;
    const float* tx;   // float32 thresholds of the x / y edges in shared memory (use_thr)
    const float* ty;
    double xlo, xhi, ylo, yhi, inv_dx, inv_dy;
    float finv_dx, finv_dy;
    // prefilter bounds and the float32 rotation in registers (the reject path is the hot one)
    float fxlo, fxhi, fylo, fyhi, fcx, fcy, fr2, fr00, fr01, fr10, fr11, fmargin;
    int nxe, nye;
    static constexpr bool rotate = ROT;   // the frame's `angle != 0`, resolved at compile time
    bool use_thr, uniform;

    __device__ __forceinline__ void init(const Pcf2dFrame* frame, const double* xedges, int nxe_, const double* yedges,
                                         int nye_, const float* s_thr, bool with_thr) {
        F = frame; xe = xedges; ye = yedges; nxe = nxe_; nye = nye_;
        xlo = xe[0]; xhi = xe[nxe - 1]; ylo = ye[0]; yhi = ye[nye - 1];
        inv_dx = (double)(nxe - 1) / (xhi - xlo);
        inv_dy = (double)(nye - 1) / (yhi - ylo);
        finv_dx = (float)inv_dx; finv_dy = (float)inv_dy;
        tx = s_thr; ty = s_thr + nxe;
        uniform = frame->uniform != 0;
        fxlo = frame->fxlo; fxhi = frame->fxhi; fylo = frame->fylo; fyhi = frame->fyhi;
        fcx = frame->fcx; fcy = frame->fcy; fr2 = frame->fr2;
        fr00 = frame->fr00; fr01 = frame->fr01; fr10 = frame->fr10; fr11 = frame->fr11; fmargin = frame->fmargin;
        use_thr = sizeof(D) == 4 && with_thr;
    }

    __device__ __forceinline__ bool bins(D dx, D dy, int& bx, int& by) const {
        double x, y;
        if constexpr (sizeof(D) == 4) {
            if constexpr (!ROT) {
                if (!(dx >= fxlo) || !(dx <= fxhi) || !(dy >= fylo) || !(dy <= fyhi)) return false;
                if (use_thr) {
                    // float32 differences without rotation: "v >= edges[k]" in float64 is
                    // "v >= thr[k]" in float32 (thr[k] = smallest float32 >= edges[k]); guess from
                    // the uniform spacing, then settle against the thresholds
                    bx = min(max(__float2int_rd((dx - fxlo) * finv_dx), 0), nxe - 2);
                    by = min(max(__float2int_rd((dy - fylo) * finv_dy), 0), nye - 2);
                    if (uniform) {
                        bx += (int)(bx < nxe - 2 && dx >= tx[bx + 1]) - (int)(dx < tx[bx]);
                        by += (int)(by < nye - 2 && dy >= ty[by + 1]) - (int)(dy < ty[by]);
                        return true;
                    }
                    while (dx < tx[bx]) --bx;
                    while (bx < nxe - 2 && dx >= tx[bx + 1]) ++bx;
                    while (dy < ty[by]) --by;
                    while (by < nye - 2 && dy >= ty[by + 1]) ++by;
                    return true;
                }
            } else {
                const float ux = dx - fcx, uy = dy - fcy;
                if (!(ux * ux + uy * uy <= fr2)) return false;
                if (use_thr) {
                    // Rotation in float32 first: the result lies within fmargin of the float64 value
                    // the reference bins (|error| <= 4e-7 (|ux| + |uy|) + 2.5e-7 (|cx| + |cy|), the
                    // margin is 1.01e-6 (|ux| + |uy| + |cx| + |cy|) plus two ulps of the largest
                    // edge).  A value farther than that from the thresholds of its bin, or from the
                    // edge range, is decided here; the rest (about 4 margin / bin width of the
                    // pairs) takes the float64 path below.
                    const float xa = fmaf(fr00, ux, fr01 * uy) + fcx;
                    const float ya = fmaf(fr10, ux, fr11 * uy) + fcy;
                    const float mg = fmaf(1.01e-6f, fabsf(ux) + fabsf(uy), fmargin);
                    if (xa < fxlo - mg || xa > fxhi + mg || ya < fylo - mg || ya > fyhi + mg) return false;
                    const int kx = min(max(__float2int_rd((xa - fxlo) * finv_dx), 0), nxe - 2);
                    const int ky = min(max(__float2int_rd((ya - fylo) * finv_dy), 0), nye - 2);
                    if (xa - tx[kx] > mg && tx[kx + 1] - xa > mg && ya - ty[ky] > mg && ty[ky + 1] - ya > mg) {
                        bx = kx;
                        by = ky;
                        return true;
                    }
                }
            }
        }
        if constexpr (!ROT) {
            x = (double)dx;
            y = (double)dy;
        } else {
            double vx, vy;
            if constexpr (sizeof(D) == 4 && CF32) {
                vx = (double)__fsub_rn(dx, (float)F->cx);
                vy = (double)__fsub_rn(dy, (float)F->cy);
            } else {
                vx = __dsub_rn((double)dx, F->cx);
                vy = __dsub_rn((double)dy, F->cy);
            }
            x = __dadd_rn(__dadd_rn(__dmul_rn(F->r00, vx), __dmul_rn(F->r01, vy)), F->cx);
            y = __dadd_rn(__dadd_rn(__dmul_rn(F->r10, vx), __dmul_rn(F->r11, vy)), F->cy);
        }
        if (!(x >= xlo) || !(x <= xhi) || !(y >= ylo) || !(y <= yhi)) return false;
        bx = pcf_bin_of(x, xe, nxe, inv_dx, uniform);
        by = pcf_bin_of(y, ye, nye, inv_dy, uniform);
        return bx >= 0 && by >= 0;
    }
};

template <typename D>
__device__ __forceinline__ D pcf_sub(D a, D b) {
    if constexpr (sizeof(D) == 4) return __fsub_rn(a, b);
    else return __dsub_rn(a, b);
}

template <typename TC, typename TO, typename D, bool CF32, bool ROT>
__global__ void __launch_bounds__(kPcfThreads) pcf2d_kernel(const TC* __restrict__ coords, long long n,
                                                            const TO* __restrict__ other, long long m, int same,
                                                            const Pcf2dFrame* __restrict__ frames,
                                                            const double* __restrict__ xedges, int nxe,
                                                            const double* __restrict__ yedges, int nye,
                                                            long long edge_stride, long long per_split,
                                                            const float* __restrict__ thr, int stage,
                                                            unsigned long long* __restrict__ counts) {
    // stage != 0: the x and y edges (float64), then their float32 thresholds (thr != nullptr)
    extern __shared__ double s_edges[];
    float* s_thr = reinterpret_cast<float*>(s_edges + (stage ? nxe + nye : 0));
    __shared__ D sx[kPcfStage], sy[kPcfStage];
    __shared__ Pcf2dFrame F;
    const int f = blockIdx.z;
    if (threadIdx.x < sizeof(Pcf2dFrame) / sizeof(int))
        reinterpret_cast<int*>(&F)[threadIdx.x] = reinterpret_cast<const int*>(frames + f)[threadIdx.x];
    const double* xe_f = xedges + (long long)f * edge_stride;
    const double* ye_f = yedges + (long long)f * edge_stride;
    if (stage) {
        for (int i = threadIdx.x; i < nxe; i += kPcfThreads) s_edges[i] = xe_f[i];
        for (int i = threadIdx.x; i < nye; i += kPcfThreads) s_edges[nxe + i] = ye_f[i];
    }
    if (sizeof(D) == 4 && thr != nullptr) {
        const float* tf = thr + (long long)f * (nxe + nye);
        for (int i = threadIdx.x; i < nxe + nye; i += kPcfThreads) s_thr[i] = tf[i];
    }
    const TC* cf = coords + (long long)f * n * 3;
    const TO* of = other + (long long)f * m * 3;
    unsigned long long* grid = counts + (size_t)f * (nxe - 1) * (nye - 1);
    const long long qi = (long long)blockIdx.x * kPcfThreads + threadIdx.x;
    const bool live = qi < m;
    const D qx = live ? (D)of[qi * 3 + 0] : (D)0, qy = live ? (D)of[qi * 3 + 1] : (D)0;
    const long long t_lo = (long long)blockIdx.y * per_split, t_hi = min(n, t_lo + per_split);
    __syncthreads();
    PairBinner<D, CF32, ROT> B;
    B.init(&F, stage ? s_edges : xe_f, nxe, stage ? s_edges + nxe : ye_f, nye, s_thr, thr != nullptr);
    for (long long t0 = t_lo; t0 < t_hi; t0 += kPcfStage) {
        const int cnt = (int)min((long long)kPcfStage, t_hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += kPcfThreads) {
            sx[i] = (D)cf[(t0 + i) * 3 + 0];
            sy[i] = (D)cf[(t0 + i) * 3 + 1];
        }
        __syncthreads();
        if (!live) continue;
        // same: 1 = identical arrays (index i pairs with index i), >= 2: the queries are the slice of the
        // targets that starts at index same - 2 (amepb_pcf2d_set_query_offset), 0 = unrelated arrays
        const long long self = same == 1 ? qi : (same >= 2 ? qi + (long long)(same - 2) : -1);
        const int skip = (self >= t0 && self < t0 + cnt) ? (int)(self - t0) : -1;
        for (int i = 0; i < cnt; ++i) {
            int bx, by;
            if (!B.bins(pcf_sub<D>(qx, sx[i]), pcf_sub<D>(qy, sy[i]), bx, by)) continue;
            if (i == skip) continue;
            atomicAdd(&grid[(size_t)bx * (nye - 1) + by], 1ULL);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Cell-sorted walk with a shared-memory window of the grid.
//
// With a wide window (the default rmax bins half of all pairs) the kernel above is bound by its
// global atomics (~1e11/s).  Here both particle sets are counting-sorted into square cells of side
// cs; the differences of a (query cell, target cell) pair span 2 cs per axis -- kPcfWin bins at
// most by the choice of cs, also after the rotation -- so a block counts a cell pair in a
// kPcfWin x kPcfWin window of 32-bit shared-memory counters and adds the non-zero ones to the
// grid afterwards: W^2 global atomics per cell pair instead of one per binned pair.  The bins
// themselves come from PairBinner on the original coordinate values (sorting only permutes), and
// a pair whose bin falls outside the window (the window origin is computed approximately) is
// added to the grid directly, so the counts do not depend on the cell geometry.
// ---------------------------------------------------------------------------------------
constexpr int kPcfWin = 96;
constexpr int kPcfCellStage = 512;   // targets of a cell per shared-memory stage

struct PcfCellPlan {
    float x0, y0, cs, inv_cs;   // cell (i, j) covers [x0 + i cs, x0 + (i+1) cs) x [y0 + j cs, ...)
    int gx, gy;
};

template <typename D> struct PcfVec2 { using type = float2; };
template <> struct PcfVec2<double> { using type = double2; };

__device__ __forceinline__ int pcf_float_order(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}

// bbox[0..3] = ordered-int min x, min y, max x, max y over the finite coordinates
template <typename T>
__global__ void pcf_bbox_kernel(const T* __restrict__ pts, long long n, int* __restrict__ bbox) {
    float lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = (float)pts[3 * i], y = (float)pts[3 * i + 1];
        if (isfinite(x)) { lox = fminf(lox, x); hix = fmaxf(hix, x); }
        if (isfinite(y)) { loy = fminf(loy, y); hiy = fmaxf(hiy, y); }
    }
    const int a = __reduce_min_sync(0xffffffffu, pcf_float_order(lox));
    const int b = __reduce_min_sync(0xffffffffu, pcf_float_order(loy));
    const int c = __reduce_max_sync(0xffffffffu, pcf_float_order(hix));
    const int d = __reduce_max_sync(0xffffffffu, pcf_float_order(hiy));
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], a); atomicMin(&bbox[1], b); atomicMax(&bbox[2], c); atomicMax(&bbox[3], d);
    }
}

template <typename T>
__device__ __forceinline__ int pcf_cell_of(const T* __restrict__ p, const PcfCellPlan& P) {
    // any deterministic assignment is correct (see above); non-finite values go to cell 0
    const float x = (float)p[0], y = (float)p[1];
    int cx = isfinite(x) ? (int)((x - P.x0) * P.inv_cs) : 0;
    int cy = isfinite(y) ? (int)((y - P.y0) * P.inv_cs) : 0;
    cx = min(max(cx, 0), P.gx - 1);
    cy = min(max(cy, 0), P.gy - 1);
    return cy * P.gx + cx;
}

template <typename T>
__global__ void pcf_cell_count_kernel(const T* __restrict__ pts, long long n, PcfCellPlan P, int* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[pcf_cell_of(pts + 3 * i, P)], 1);
}

// start[c] = exclusive prefix sum of count (start[ncell] = n), count zeroed for use as a cursor
__global__ void pcf_cell_scan_kernel(int* __restrict__ count, int* __restrict__ start, int ncell) {
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    const int per = (ncell + 1023) / 1024;
    const int lo = min(tid * per, ncell), hi = min(lo + per, ncell);
    int sum = 0;
    for (int c = lo; c < hi; ++c) sum += count[c];
    s_part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = tid >= off ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - sum;
    for (int c = lo; c < hi; ++c) {
        const int v = count[c];
        start[c] = run;
        run += v;
        count[c] = 0;
    }
    if (tid == 1023) start[ncell] = s_part[1023];
}

template <typename T, typename D>
__global__ void pcf_cell_scatter_kernel(const T* __restrict__ pts, long long n, PcfCellPlan P,
                                        const int* __restrict__ start, int* __restrict__ cursor,
                                        typename PcfVec2<D>::type* __restrict__ sorted, int* __restrict__ sidx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = pcf_cell_of(pts + 3 * i, P);
    const int pos = start[c] + atomicAdd(&cursor[c], 1);
    typename PcfVec2<D>::type v;
    v.x = (D)pts[3 * i];
    v.y = (D)pts[3 * i + 1];
    sorted[pos] = v;
    sidx[pos] = (int)i;
}

template <typename D, bool CF32, bool ROT>
__global__ void __launch_bounds__(kPcfThreads) pcf2d_cells_kernel(
    const typename PcfVec2<D>::type* __restrict__ qs, const int* __restrict__ qidx, const int* __restrict__ qstart,
    const typename PcfVec2<D>::type* __restrict__ ts, const int* __restrict__ tidx, const int* __restrict__ tstart,
    PcfCellPlan P, int same, const Pcf2dFrame* __restrict__ frame, const double* __restrict__ xe, int nxe,
    const double* __restrict__ ye, int nye, const float* __restrict__ thr, int stage,
    unsigned long long* __restrict__ grid) {
    // window, then (stage != 0) the float64 edges and (thr != nullptr) their float32 thresholds
    extern __shared__ __align__(8) unsigned s_dyn[];
    unsigned* win = s_dyn;
    double* s_edges = reinterpret_cast<double*>(s_dyn + kPcfWin * kPcfWin);
    float* s_thr = reinterpret_cast<float*>(s_edges + (stage ? nxe + nye : 0));
    if (stage) {
        for (int i = threadIdx.x; i < nxe; i += kPcfThreads) s_edges[i] = xe[i];
        for (int i = threadIdx.x; i < nye; i += kPcfThreads) s_edges[nxe + i] = ye[i];
    }
    __shared__ Pcf2dFrame F;
    __shared__ typename PcfVec2<D>::type s_tv[kPcfCellStage];
    __shared__ int s_ti[kPcfCellStage];
    if (threadIdx.x < sizeof(Pcf2dFrame) / sizeof(int))
        reinterpret_cast<int*>(&F)[threadIdx.x] = reinterpret_cast<const int*>(frame)[threadIdx.x];
    if (sizeof(D) == 4 && thr != nullptr)
        for (int i = threadIdx.x; i < nxe + nye; i += kPcfThreads) s_thr[i] = thr[i];
    const int qc = blockIdx.x, ncell = P.gx * P.gy;
    const int q0 = qstart[qc], nqc = qstart[qc + 1] - q0;
    if (nqc == 0) return;
    __syncthreads();
    PairBinner<D, CF32, ROT> B;
    B.init(&F, stage ? s_edges : xe, nxe, stage ? s_edges + nxe : ye, nye, s_thr, thr != nullptr);
    const int qcx = qc % P.gx, qcy = qc / P.gx;
    const double qxlo = (double)P.x0 + (double)qcx * P.cs, qylo = (double)P.y0 + (double)qcy * P.cs;
    for (int tc = blockIdx.y; tc < ncell; tc += gridDim.y) {
        const int t0 = tstart[tc], ntc = tstart[tc + 1] - t0;
        if (ntc == 0) continue;
        // differences of this cell pair (one cell side of slack: clamped and rounded members)
        const int tcx = tc % P.gx, tcy = tc / P.gx;
        const double txlo = (double)P.x0 + (double)tcx * P.cs, tylo = (double)P.y0 + (double)tcy * P.cs;
        const bool edge_cell = qcx == 0 || qcy == 0 || qcx == P.gx - 1 || qcy == P.gy - 1 || tcx == 0 || tcy == 0 ||
                               tcx == P.gx - 1 || tcy == P.gy - 1;
        double dlo[2] = {qxlo - (txlo + P.cs), qylo - (tylo + P.cs)};
        double dhi[2] = {qxlo + P.cs - txlo, qylo + P.cs - tylo};
        double wxlo, wxhi, wylo, wyhi;
        if (!ROT) {
            wxlo = dlo[0]; wxhi = dhi[0]; wylo = dlo[1]; wyhi = dhi[1];
        } else {
            wxlo = wylo = INFINITY; wxhi = wyhi = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double vx = ((k & 1) ? dhi[0] : dlo[0]) - F.cx, vy = ((k & 2) ? dhi[1] : dlo[1]) - F.cy;
                const double x = F.r00 * vx + F.r01 * vy + F.cx, y = F.r10 * vx + F.r11 * vy + F.cy;
                wxlo = fmin(wxlo, x); wxhi = fmax(wxhi, x); wylo = fmin(wylo, y); wyhi = fmax(wyhi, y);
            }
        }
        // cells on the rim also hold whatever lies outside the grid: no geometric statement there
        if (!edge_cell) {
            const double slack_x = 2.0 / B.inv_dx, slack_y = 2.0 / B.inv_dy;
            if (wxhi + slack_x < B.xlo || wxlo - slack_x > B.xhi || wyhi + slack_y < B.ylo || wylo - slack_y > B.yhi) continue;
        }
        const int bx0 = max(0, (int)floor((wxlo - B.xlo) * B.inv_dx) - 1);
        const int by0 = max(0, (int)floor((wylo - B.ylo) * B.inv_dy) - 1);
        const bool priv = (unsigned long long)nqc * (unsigned long long)ntc < 0xffffffffull;
        __syncthreads();
        for (int i = threadIdx.x; i < kPcfWin * kPcfWin; i += kPcfThreads) win[i] = 0u;
        __syncthreads();
        // targets of the cell staged through shared memory (read by all threads at once)
        for (int c0 = 0; c0 < ntc; c0 += kPcfCellStage) {
            const int cnt = min(kPcfCellStage, ntc - c0);
            if (c0 > 0) __syncthreads();
            for (int i = threadIdx.x; i < cnt; i += kPcfThreads) {
                s_tv[i] = ts[t0 + c0 + i];
                s_ti[i] = same ? tidx[t0 + c0 + i] : -1;
            }
            __syncthreads();
            for (int q = threadIdx.x; q < nqc; q += kPcfThreads) {
                const typename PcfVec2<D>::type qv = qs[q0 + q];
                const int skip = same == 1 ? qidx[q0 + q] : (same >= 2 ? qidx[q0 + q] + (same - 2) : -2);
                for (int i = 0; i < cnt; ++i) {
                    const typename PcfVec2<D>::type tv = s_tv[i];
                    int bx, by;
                    if (!B.bins(pcf_sub<D>(qv.x, tv.x), pcf_sub<D>(qv.y, tv.y), bx, by)) continue;
                    if (s_ti[i] == skip) continue;
                    const unsigned lx = (unsigned)(bx - bx0), ly = (unsigned)(by - by0);
                    if (priv && lx < (unsigned)kPcfWin && ly < (unsigned)kPcfWin) atomicAdd(&win[lx * kPcfWin + ly], 1u);
                    else atomicAdd(&grid[(size_t)bx * (nye - 1) + by], 1ULL);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kPcfWin * kPcfWin; i += kPcfThreads) {
            const unsigned v = win[i];
            if (v) atomicAdd(&grid[(size_t)(bx0 + i / kPcfWin) * (nye - 1) + (by0 + i % kPcfWin)], (unsigned long long)v);
        }
    }
}

template <typename TC, typename TO, bool ROT>
static void launch_pcf2d_rot(dim3 grid, cudaStream_t stream, bool center_f32, const void* coords, long long n,
                             const void* other, long long m, int same, const Pcf2dFrame* frames, const double* xe,
                             int nxe, const double* ye, int nye, long long edge_stride, long long per_split,
                             const float* thr, int stage, size_t smem, unsigned long long* counts) {
    constexpr bool both_f32 = sizeof(TC) == 4 && sizeof(TO) == 4;
    if constexpr (both_f32) {
        if (center_f32)
            pcf2d_kernel<TC, TO, float, true, ROT><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, stage, counts);
        else
            pcf2d_kernel<TC, TO, float, false, ROT><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, stage, counts);
    } else {
        pcf2d_kernel<TC, TO, double, false, ROT><<<grid, kPcfThreads, smem, stream>>>((const TC*)coords, n, (const TO*)other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, stage, counts);
    }
}

// all frames of one launch share `rot` (the kernels resolve it at compile time)
template <typename TC, typename TO>
static void launch_pcf2d(bool rot, dim3 grid, cudaStream_t stream, bool center_f32, const void* coords, long long n,
                         const void* other, long long m, int same, const Pcf2dFrame* frames, const double* xe, int nxe,
                         const double* ye, int nye, long long edge_stride, long long per_split, const float* thr,
                         int stage, size_t smem, unsigned long long* counts) {
    if (rot)
        launch_pcf2d_rot<TC, TO, true>(grid, stream, center_f32, coords, n, other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, stage, smem, counts);
    else
        launch_pcf2d_rot<TC, TO, false>(grid, stream, center_f32, coords, n, other, m, same, frames, xe, nxe, ye, nye, edge_stride, per_split, thr, stage, smem, counts);
}

__global__ void pcf_bbox_init_kernel(int* bbox) {
    bbox[0] = bbox[1] = pcf_float_order(INFINITY);
    bbox[2] = bbox[3] = pcf_float_order(-INFINITY);
}

inline float pcf_order_to_float(int i) {
    i ^= (i >> 31) & 0x7fffffff;
    float f;
    memcpy(&f, &i, sizeof f);
    return f;
}

// Sorts one particle set into the cells of P: sorted positions (as D), original indices, cell starts.
template <typename T, typename D>
static int pcf_sort_cells(amepb_ctx* ctx, cudaStream_t stream, const T* pts, long long n, const PcfCellPlan& P,
                          int buf_sorted, int buf_idx, int buf_count, int buf_start) {
    const int ncell = P.gx * P.gy;
    int rc;
    if ((rc = ensure_device(ctx, buf_sorted, sizeof(typename PcfVec2<D>::type) * (size_t)n))) return rc;
    if ((rc = ensure_device(ctx, buf_idx, sizeof(int) * (size_t)n))) return rc;
    if ((rc = ensure_device(ctx, buf_count, sizeof(int) * (size_t)ncell))) return rc;
    if ((rc = ensure_device(ctx, buf_start, sizeof(int) * ((size_t)ncell + 1)))) return rc;
    int* count = reinterpret_cast<int*>(ctx->bufs[buf_count].p);
    int* start = reinterpret_cast<int*>(ctx->bufs[buf_start].p);
    AMEPB_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (size_t)ncell, stream));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    pcf_cell_count_kernel<T><<<blocks, 256, 0, stream>>>(pts, n, P, count);
    pcf_cell_scan_kernel<<<1, 1024, 0, stream>>>(count, start, ncell);
    pcf_cell_scatter_kernel<T, D><<<blocks, 256, 0, stream>>>(
        pts, n, P, start, count, reinterpret_cast<typename PcfVec2<D>::type*>(ctx->bufs[buf_sorted].p),
        reinterpret_cast<int*>(ctx->bufs[buf_idx].p));
    AMEPB_LAUNCH_CHECK(ctx);
    return 0;
}

// One frame through the cell-sorted walk.  Returns 0 when done, 1 when the frame does not
// qualify (the caller then uses pcf2d_kernel), < 0 / cudaError on failure.
template <typename TC, typename TO>
static int pcf2d_cells_frame(amepb_ctx* ctx, cudaStream_t stream, const TC* coords, long long n, const TO* other,
                             long long m, int same, bool center_f32, const Pcf2dFrame& hF, const Pcf2dFrame* d_frame,
                             const double* h_xe, const double* d_xe, int nxe, const double* h_ye, const double* d_ye,
                             int nye, const float* d_thr, double min_occupancy, int* h_bbox,
                             unsigned long long* d_grid) {
    using D = typename std::conditional<sizeof(TC) == 4 && sizeof(TO) == 4, float, double>::type;
    if (n >= (1LL << 31) || m >= (1LL << 31)) return 1;
    int rc;
    if ((rc = ensure_device(ctx, BUF_MISC, 64))) return rc;
    int* d_bbox = reinterpret_cast<int*>(ctx->bufs[BUF_MISC].p);
    pcf_bbox_init_kernel<<<1, 1, 0, stream>>>(d_bbox);
    const unsigned rblocks = (unsigned)std::min<long long>((n + 255) / 256, 4LL * ctx->sm_count);
    pcf_bbox_kernel<TC><<<rblocks, 256, 0, stream>>>(coords, n, d_bbox);
    if (!same) {
        const unsigned oblocks = (unsigned)std::min<long long>((m + 255) / 256, 4LL * ctx->sm_count);
        pcf_bbox_kernel<TO><<<oblocks, 256, 0, stream>>>(other, m, d_bbox);
    }
    AMEPB_LAUNCH_CHECK(ctx);
    AMEPB_CUDA(ctx, cudaMemcpyAsync(h_bbox, d_bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    const float lox = pcf_order_to_float(h_bbox[0]), loy = pcf_order_to_float(h_bbox[1]);
    const float hix = pcf_order_to_float(h_bbox[2]), hiy = pcf_order_to_float(h_bbox[3]);
    if (!(lox <= hix) || !(loy <= hiy) || !std::isfinite(lox) || !std::isfinite(hix) || !std::isfinite(loy) ||
        !std::isfinite(hiy))
        return 1;
    // cell side: the differences of a cell pair span 2 cs per axis, (|cos| + |sin|) times that after
    // the rotation, and have to fit the window with a few bins to spare
    const double bwx = (h_xe[nxe - 1] - h_xe[0]) / (nxe - 1), bwy = (h_ye[nye - 1] - h_ye[0]) / (nye - 1);
    const double spread = hF.rotate ? fabs(hF.r00) + fabs(hF.r10) : 1.0;
    if (!(spread >= 1.0 - 1e-9) || !(spread <= 1.5)) return 1;   // NaN angle: nothing bins anyway
    const double cs = 0.5 * (kPcfWin - 6) * std::min(bwx, bwy) / spread;
    const double ex = (double)hix - lox, ey = (double)hiy - loy;
    if (!(cs > 0.0) || ex / cs > 4096.0 || ey / cs > 4096.0) return 1;
    PcfCellPlan P;
    P.x0 = lox; P.y0 = loy; P.cs = (float)cs; P.inv_cs = (float)(1.0 / cs);
    P.gx = (int)(ex / cs) + 1; P.gy = (int)(ey / cs) + 1;
    const long long ncell = (long long)P.gx * P.gy;
    if (ncell > 65536 || (double)n / ncell < min_occupancy || (double)m / ncell < min_occupancy) return 1;
    if ((rc = pcf_sort_cells<TC, D>(ctx, stream, coords, n, P, BUF_TSORTED, BUF_QRANK, BUF_TCOUNT, BUF_TSTART))) return rc;
    if (!same)
        if ((rc = pcf_sort_cells<TO, D>(ctx, stream, other, m, P, BUF_QSORTED, BUF_SCAN, BUF_QCOUNT, BUF_QSTART))) return rc;
    using V = typename PcfVec2<D>::type;
    const V* ts = reinterpret_cast<const V*>(ctx->bufs[BUF_TSORTED].p);
    const int* tidx = reinterpret_cast<const int*>(ctx->bufs[BUF_QRANK].p);
    const int* tstart = reinterpret_cast<const int*>(ctx->bufs[BUF_TSTART].p);
    const V* qs = same ? ts : reinterpret_cast<const V*>(ctx->bufs[BUF_QSORTED].p);
    const int* qidx = same ? tidx : reinterpret_cast<const int*>(ctx->bufs[BUF_SCAN].p);
    const int* qstart = same ? tstart : reinterpret_cast<const int*>(ctx->bufs[BUF_QSTART].p);
    // thresholds only while they fit the default 48 KB of dynamic shared memory beside the window
    const int stage = (size_t)nxe + nye <= kPcfStageEdges ? 1 : 0;
    const bool with_thr = d_thr != nullptr && stage;
    const size_t smem = sizeof(unsigned) * kPcfWin * kPcfWin +
                        (stage ? sizeof(double) * ((size_t)nxe + nye) : 0) + (with_thr ? sizeof(float) * ((size_t)nxe + nye) : 0);
    const long long ysplit = std::max(1LL, std::min<long long>({(8LL * ctx->sm_count + ncell - 1) / ncell, ncell, 64LL}));
    const dim3 grid((unsigned)ncell, (unsigned)ysplit);
    const float* thr = with_thr ? d_thr : nullptr;
    if (!ctx->pcf_attr) {
        const int cap = (int)(sizeof(unsigned) * kPcfWin * kPcfWin + (sizeof(double) + sizeof(float)) * kPcfStageEdges);
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<float, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<float, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<float, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<float, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<double, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf2d_cells_kernel<double, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        ctx->pcf_attr = true;
    }
    const bool rot = hF.rotate != 0;
#define AMEPB_PCF_CELLS(DT, CF, RT, THR) \
    pcf2d_cells_kernel<DT, CF, RT><<<grid, kPcfThreads, smem, stream>>>(qs, qidx, qstart, ts, tidx, tstart, P, (same ? 1 : ctx->pcf2d_self_code), d_frame, d_xe, nxe, d_ye, nye, THR, stage, d_grid)
    if constexpr (sizeof(D) == 4) {
        if (center_f32) { if (rot) AMEPB_PCF_CELLS(float, true, true, thr); else AMEPB_PCF_CELLS(float, true, false, thr); }
        else { if (rot) AMEPB_PCF_CELLS(float, false, true, thr); else AMEPB_PCF_CELLS(float, false, false, thr); }
    } else {
        if (rot) AMEPB_PCF_CELLS(double, false, true, nullptr); else AMEPB_PCF_CELLS(double, false, false, nullptr);
    }
#undef AMEPB_PCF_CELLS
    AMEPB_LAUNCH_CHECK(ctx);
    return 0;
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
    const int stage = (size_t)nxe + nye <= kPcfStageEdges ? 1 : 0;
    const bool with_thr = all_f32 && stage;
    const size_t tbytes = with_thr ? ((sizeof(float) * (size_t)nframes * (nxe + nye) + 15) & ~(size_t)15) : 0;
    const size_t pbytes = fbytes + 2 * ebytes + tbytes;
    const size_t bbox_off = (pbytes + 63) & ~(size_t)63;   // bounding-box read-back of the cell-sorted walk
    if ((rc = ensure_pinned(ctx, bbox_off + 64))) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, pbytes))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;   // an earlier call may still be reading the block
    char* hp = reinterpret_cast<char*>(ctx->pinned);
    memset(hp, 0, pbytes);
    Pcf2dFrame* hf = reinterpret_cast<Pcf2dFrame*>(hp);
    const bool no_f32rot = getenv("AMEPB_PCF2D_NO_F32ROT") != nullptr;   // diagnostics / tests
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
        F.fr00 = (float)F.r00; F.fr01 = (float)F.r01; F.fr10 = (float)F.r10; F.fr11 = (float)F.r11;
        F.fcx = (float)F.cx; F.fcy = (float)F.cy;
        const double emax = std::max(std::max(fabs(xe[0]), fabs(xe[nxe - 1])), std::max(fabs(ye[0]), fabs(ye[nye - 1])));
        // constant part of the margin; bins() adds 1.01e-6 (|ux| + |uy|) per pair
        F.fmargin = float_above(1.01e-6 * (fabs(F.cx) + fabs(F.cy)) + 2.5e-7 * emax + 1e-30, false);
        if (no_f32rot) F.fmargin = INFINITY;   // nothing is decided in float32: every pair takes the float64 path
        auto uniform_axis = [](const double* e, int ne) {
            const double w = (e[ne - 1] - e[0]) / (ne - 1);
            for (int k = 0; k < ne; ++k) {
                if (!(fabs(e[k] - (e[0] + k * w)) <= 0.2 * w)) return false;
                if (k > 0 && !(float_above(e[k - 1], false) < float_above(e[k], false))) return false;
            }
            return true;
        };
        F.uniform = uniform_axis(xe, nxe) && uniform_axis(ye, nye) ? 1 : 0;
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
    const size_t smem = (stage ? sizeof(double) * ((size_t)nxe + nye) : 0) + (with_thr ? sizeof(float) * ((size_t)nxe + nye) : 0);

    const bool cf32 = center_f32 != 0;
    // frames [f0, f0 + nf) through pcf2d_kernel: query tiles x target splits x frames, at least
    // ~4 blocks per SM when the system allows it
    auto launch_direct_run = [&](int64_t f0, int64_t nf, bool rot_run) {
        const long long qtiles = (m + kPcfThreads - 1) / kPcfThreads;
        const long long want = 4LL * ctx->sm_count;
        long long tsplit = (want + qtiles * nf - 1) / (qtiles * nf);
        const long long max_split = (n + kPcfStage - 1) / kPcfStage;
        tsplit = std::max(1LL, std::min({tsplit, max_split, 65535LL}));
        long long per_split = (n + tsplit - 1) / tsplit;
        per_split = (per_split + kPcfStage - 1) / kPcfStage * kPcfStage;
        tsplit = (n + per_split - 1) / per_split;
        const dim3 grid((unsigned)qtiles, (unsigned)tsplit, (unsigned)nf);
        const char* c0 = reinterpret_cast<const char*>(d_coords) + (size_t)f0 * n * 3 * (dtype == AMEPB_F32 ? 4 : 8);
        const char* o0 = reinterpret_cast<const char*>(d_other) + (size_t)f0 * m * 3 * (other_dtype == AMEPB_F32 ? 4 : 8);
        const Pcf2dFrame* fr = d_frames + f0;
        const double* xe0 = d_xe + f0 * dev_stride;
        const double* ye0 = d_ye + f0 * dev_stride;
        const float* th0 = d_thr ? d_thr + f0 * (nxe + nye) : nullptr;
        unsigned long long* g0 = d_counts + (size_t)f0 * (nxe - 1) * (nye - 1);
        if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32)
            launch_pcf2d<float, float>(rot_run, grid, stream, cf32, c0, n, o0, m, (same ? 1 : ctx->pcf2d_self_code), fr, xe0, nxe, ye0, nye, dev_stride, per_split, th0, stage, smem, g0);
        else if (dtype == AMEPB_F32)
            launch_pcf2d<float, double>(rot_run, grid, stream, cf32, c0, n, o0, m, (same ? 1 : ctx->pcf2d_self_code), fr, xe0, nxe, ye0, nye, dev_stride, per_split, th0, stage, smem, g0);
        else if (other_dtype == AMEPB_F32)
            launch_pcf2d<double, float>(rot_run, grid, stream, cf32, c0, n, o0, m, (same ? 1 : ctx->pcf2d_self_code), fr, xe0, nxe, ye0, nye, dev_stride, per_split, th0, stage, smem, g0);
        else
            launch_pcf2d<double, double>(rot_run, grid, stream, cf32, c0, n, o0, m, (same ? 1 : ctx->pcf2d_self_code), fr, xe0, nxe, ye0, nye, dev_stride, per_split, th0, stage, smem, g0);
    };
    // consecutive frames with the same `angle != 0` share a launch
    auto launch_direct = [&](int64_t f0, int64_t nf) {
        int64_t a = f0;
        while (a < f0 + nf) {
            int64_t b = a + 1;
            while (b < f0 + nf && hf[b].rotate == hf[a].rotate) ++b;
            launch_direct_run(a, b - a, hf[a].rotate != 0);
            a = b;
        }
    };
    // Large systems: cell-sorted walk with a shared-memory window, frame by frame (each frame
    // needs its bounding box on the host).  AMEPB_PCF2D_CELLS=0 disables it, =1 drops the
    // occupancy requirement (tests).
    const char* cells_env = getenv("AMEPB_PCF2D_CELLS");
    const int cells_mode = cells_env ? atoi(cells_env) : -1;
    const bool try_cells = cells_mode != 0 && (cells_mode == 1 || (double)n * (double)m >= 2.5e9);
    if (!try_cells) {
        launch_direct(0, nframes);
    } else {
        const double min_occ = cells_mode == 1 ? 0.0 : 128.0;
        for (int64_t f = 0; f < nframes; ++f) {
            const char* c0 = reinterpret_cast<const char*>(d_coords) + (size_t)f * n * 3 * (dtype == AMEPB_F32 ? 4 : 8);
            const char* o0 = reinterpret_cast<const char*>(d_other) + (size_t)f * m * 3 * (other_dtype == AMEPB_F32 ? 4 : 8);
            const double* hxf = hxe + f * (edge_stride ? estride : 0);
            const double* hyf = hye + f * (edge_stride ? estride : 0);
            const float* th0 = d_thr ? d_thr + f * (nxe + nye) : nullptr;
            unsigned long long* g0 = d_counts + (size_t)f * (nxe - 1) * (nye - 1);
            int* h_bbox = reinterpret_cast<int*>(hp + bbox_off);
            int r;
            if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32)
                r = pcf2d_cells_frame<float, float>(ctx, stream, (const float*)c0, n, (const float*)o0, m, same, cf32, hf[f], d_frames + f, hxf, d_xe + f * dev_stride, nxe, hyf, d_ye + f * dev_stride, nye, th0, min_occ, h_bbox, g0);
            else if (dtype == AMEPB_F32)
                r = pcf2d_cells_frame<float, double>(ctx, stream, (const float*)c0, n, (const double*)o0, m, same, cf32, hf[f], d_frames + f, hxf, d_xe + f * dev_stride, nxe, hyf, d_ye + f * dev_stride, nye, th0, min_occ, h_bbox, g0);
            else if (other_dtype == AMEPB_F32)
                r = pcf2d_cells_frame<double, float>(ctx, stream, (const double*)c0, n, (const float*)o0, m, same, cf32, hf[f], d_frames + f, hxf, d_xe + f * dev_stride, nxe, hyf, d_ye + f * dev_stride, nye, th0, min_occ, h_bbox, g0);
            else
                r = pcf2d_cells_frame<double, double>(ctx, stream, (const double*)c0, n, (const double*)o0, m, same, cf32, hf[f], d_frames + f, hxf, d_xe + f * dev_stride, nxe, hyf, d_ye + f * dev_stride, nye, th0, min_occ, h_bbox, g0);
            if (r == 1) launch_direct(f, 1);
            else if (r != 0) return r;
        }
    }
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, gbytes, cudaMemcpyDeviceToHost, stream));
    }
    // the pinned parameter block is reused by the next call
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int amepb_pcf2d_set_query_offset(amepb_ctx* ctx, int64_t offset) {
    if (!ctx || offset >= (1LL << 31) - 2) return AMEPB_E_INVAL;
    ctx->pcf2d_self_code = offset < 0 ? 0 : (int)offset + 2;
    return 0;
}

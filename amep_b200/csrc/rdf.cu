// rdf.cu -- periodic-boundary pair histograms (RDF) on sm_100a.
//
// Replaces, for a whole batch of frames at once:
//   amep/pbc.py:169-350         pbc_points(width=0.5)   -> images generated on the fly
//   amep/spatialcor.py:348-388  __rdf_diff              -> cell-list pair kernel
//   amep/spatialcor.py:566-600  chunk fan-out + sum     -> persistent blocks + one merge
//
// Pipeline per sub-batch of frames (all on the caller's stream):
//   stats      min/max/constant-column flags per frame          (HBM bound, 12 B/particle)
//   plan       host: grid, neighbour rows, tasks                (rdf_plan.h)
//   count      periodic images -> cell counters                  (HBM bound)
//   scan       exclusive scan of the cell counters
//   fill       images / queries written in cell order (float4)   (HBM bound)
//   pairs      persistent blocks: warp = (query cell, neighbour row); lanes hold
//              images in registers, queries are broadcast; exact float32 / float64
//              distance arithmetic without FMA; squared-distance thresholds for
//              binning; per-block shared-memory histograms, one merge per frame
#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#include <vector>

#include "amepb_internal.cuh"
#include "rdf_plan.h"

namespace amepb {

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
template <typename S>
struct Vec4;
template <>
struct __align__(16) Vec4<float> {
    float x, y, z, w;
};
template <>
struct __align__(32) Vec4<double> {
    double x, y, z, w;
};

__device__ __forceinline__ Vec4<float> load_vec4(const Vec4<float>* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    Vec4<float> r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
__device__ __forceinline__ Vec4<double> load_vec4(const Vec4<double>* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    Vec4<double> r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
}

// One record of a cell-sorted array.  compact (FramePlan::compact): float32 arrays of flat frames
// hold (x, y) pairs only -- half the bytes of the scattered stores of the fill pass, which are what
// bounds it (tools/place_probe.cu: 8-byte records 123 G/s, 16-byte records 75 G/s), and half the
// bytes the pair kernel gathers.
__device__ __forceinline__ void store_record(Vec4<float>* arr, unsigned slot, float x, float y, float z, bool compact) {
    if (compact) {
        reinterpret_cast<float2*>(arr)[slot] = make_float2(x, y);
    } else {
        Vec4<float> o;
        o.x = x; o.y = y; o.z = z; o.w = 0.0f;
        arr[slot] = o;
    }
}
__device__ __forceinline__ void store_record(Vec4<double>* arr, unsigned slot, double x, double y, double z, bool) {
    Vec4<double> o;
    o.x = x; o.y = y; o.z = z; o.w = 0.0;
    arr[slot] = o;
}

struct StatsOut {  // mirrors FrameStats, device side
    double lo[3];
    double hi[3];
    double first[3];
    int varies[3];
    int pad;
};

// ---------------------------------------------------------------------------
// stats kernel: per frame min / max / "column differs from particle 0"
// ---------------------------------------------------------------------------
// (minima and maxima are taken in the element type -- exact, and for float32 input free of the six
// float->double conversions per particle that kept this HBM-bound pass on the conversion pipe)
__device__ __forceinline__ float stat_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float stat_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double stat_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double stat_max(double a, double b) { return fmax(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) stats_partial_kernel(const T* __restrict__ pts, long long n,
                                                            StatsOut* __restrict__ partial) {
    const int f = blockIdx.y;
    const T* base = pts + (long long)f * n * 3;
    const T f0 = base[0], f1 = base[1], f2 = base[2];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const T tinf = (T)inf;
    T lo0 = tinf, lo1 = tinf, lo2 = tinf, hi0 = -tinf, hi1 = -tinf, hi2 = -tinf;
    int v0 = 0, v1 = 0, v2 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const T x = base[3 * i], y = base[3 * i + 1], z = base[3 * i + 2];
        v0 |= !(x == f0);
        v1 |= !(y == f1);
        v2 |= !(z == f2);
        lo0 = stat_min(lo0, x); hi0 = stat_max(hi0, x);
        lo1 = stat_min(lo1, y); hi1 = stat_max(hi1, y);
        lo2 = stat_min(lo2, z); hi2 = stat_max(hi2, z);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo0 = stat_min(lo0, __shfl_xor_sync(0xffffffffu, lo0, o));
        lo1 = stat_min(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
        lo2 = stat_min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
        hi0 = stat_max(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
        hi1 = stat_max(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
        hi2 = stat_max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
        v0 |= __shfl_xor_sync(0xffffffffu, v0, o);
        v1 |= __shfl_xor_sync(0xffffffffu, v1, o);
        v2 |= __shfl_xor_sync(0xffffffffu, v2, o);
    }
    __shared__ double s_red[8][6];
    __shared__ int s_var[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_red[warp][0] = (double)lo0; s_red[warp][1] = (double)lo1; s_red[warp][2] = (double)lo2;
        s_red[warp][3] = (double)hi0; s_red[warp][4] = (double)hi1; s_red[warp][5] = (double)hi2;
        s_var[warp][0] = v0; s_var[warp][1] = v1; s_var[warp][2] = v2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        StatsOut o;
        for (int d = 0; d < 3; ++d) { o.lo[d] = inf; o.hi[d] = -inf; o.varies[d] = 0; }
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w)
            for (int d = 0; d < 3; ++d) {
                o.lo[d] = fmin(o.lo[d], s_red[w][d]);
                o.hi[d] = fmax(o.hi[d], s_red[w][3 + d]);
                o.varies[d] |= s_var[w][d];
            }
        o.first[0] = (double)f0; o.first[1] = (double)f1; o.first[2] = (double)f2;
        o.pad = 0;
        partial[(long long)f * gridDim.x + blockIdx.x] = o;
    }
}

__global__ void __launch_bounds__(32) stats_final_kernel(const StatsOut* __restrict__ partial, int nparts,
                                                         StatsOut* __restrict__ out) {
    // one warp per frame; lanes stride over the partials
    const int f = blockIdx.x, lane = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    int var[3] = {0, 0, 0};
    for (int p = lane; p < nparts; p += 32) {
        const StatsOut& q = partial[(long long)f * nparts + p];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fmin(lo[d], q.lo[d]);
            hi[d] = fmax(hi[d], q.hi[d]);
            var[d] |= q.varies[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
            var[d] |= __shfl_xor_sync(0xffffffffu, var[d], o);
        }
    if (lane == 0) {
        StatsOut o = partial[(long long)f * nparts];
        for (int d = 0; d < 3; ++d) {
            o.lo[d] = lo[d];
            o.hi[d] = hi[d];
            o.varies[d] = var[d];
        }
        out[f] = o;
    }
}

// ---------------------------------------------------------------------------
// cell placement: count pass and fill pass share one body
// ---------------------------------------------------------------------------
__device__ __forceinline__ int cell_index(double x, double g0, double inv_h, int n) {
    int c = __double2int_rd((x - g0) * inv_h);  // NaN -> 0, saturating
    return min(max(c, 0), n - 1);
}

// The placement kernels read the head of their frame's plan (everything before the neighbour
// row tables) from shared memory.
constexpr int kPlanHeadWords = (int)(offsetof(FramePlan, row_oy) / sizeof(int));
__device__ __forceinline__ const FramePlan& stage_plan_head(const FramePlan* __restrict__ plans, int f, int* s_words) {
    const int* src = reinterpret_cast<const int*>(plans + f);
    for (int k = threadIdx.x; k < kPlanHeadWords; k += blockDim.x) s_words[k] = src[k];
    __syncthreads();
    return *reinterpret_cast<const FramePlan*>(s_words);
}

// Targets with periodic images.  T: input element type, image storage is
// float32 (amep/pbc.py:300-302).  FILL=false counts, FILL=true writes.
// An image is kept iff it passes the strict window of pbc_points and lies in the grid
// region; both tests are folded into one closed float32 interval per dimension
// (FramePlan::keep_lo/hi).  Most particles have no shifted image in the region: they take
// the short path (nothing at all in the symmetric mode, where the unshifted image is the query).
template <typename T, bool FILL>
__device__ __forceinline__ void place_images_body(const T* __restrict__ p, const FramePlan& P,
                                                  unsigned* __restrict__ count,
                                                  const unsigned* __restrict__ start,
                                                  Vec4<float>* __restrict__ sorted) {
    // Shortcut for the interior: almost every particle is farther than the cutoff from all box
    // faces and cannot have a kept shifted image (near_lo / near_hi, conservative bounds from the
    // host); it then needs no image arithmetic at all -- 70 of the 190 instructions per particle.
    {
        bool maybe = false;
        float b0[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            b0[d] = (float)p[d];
            if (P.nshift[d] == 3) maybe = maybe || !(b0[d] > P.near_lo[d] && b0[d] < P.near_hi[d]);   // (NaN: full test)
        }
        if (!maybe) {
            if (P.sym) return;   // the unshifted image is the query itself
            if (!(b0[0] >= P.keep_lo[0] && b0[0] <= P.keep_hi[0] && b0[1] >= P.keep_lo[1] && b0[1] <= P.keep_hi[1] &&
                  b0[2] >= P.keep_lo[2] && b0[2] <= P.keep_hi[2]))
                return;
            const int cx = cell_index((double)b0[0], P.g0[0], P.inv_h[0], P.n[0]);
            const int cy = cell_index((double)b0[1], P.g0[1], P.inv_h[1], P.n[1]);
            const int cz = cell_index((double)b0[2], P.g0[2], P.inv_h[2], P.n[2]);
            const unsigned c = P.cell_base + (unsigned)((cz * P.n[1] + cy) * P.n[0] + cx);
            if (!FILL) {
                atomicAdd(&count[c], 1u);
            } else {
                const unsigned slot = start[c] + atomicSub(&count[c], 1u) - 1u;
                store_record(sorted, slot, b0[0], b0[1], b0[2], P.compact != 0);
            }
            return;
        }
    }
    float img[3][3];
    bool ok[3][3];
    bool any_shift = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float b = (float)p[d];  // .astype(np.float32): round to nearest even
        const float len = P.shift[d];
        const float lo = P.keep_lo[d], hi = P.keep_hi[d];
        const bool shifted = P.nshift[d] == 3;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            // v = 0: no shift, 1: -L, 2: +L.  base + v*box in float32 (amep/pbc.py:326)
            const float vf = (v == 0) ? 0.0f : (v == 1 ? -1.0f : 1.0f);
            const float x = __fadd_rn(b, __fmul_rn(vf, len));
            img[d][v] = x;
            ok[d][v] = (v == 0 || shifted) && x >= lo && x <= hi;
            if (v) any_shift = any_shift || ok[d][v];
        }
    }
    auto place = [&](int vx, int vy, int vz) {
        const int cx = cell_index((double)img[0][vx], P.g0[0], P.inv_h[0], P.n[0]);
        const int cy = cell_index((double)img[1][vy], P.g0[1], P.inv_h[1], P.n[1]);
        const int cz = cell_index((double)img[2][vz], P.g0[2], P.inv_h[2], P.n[2]);
        const unsigned c = P.cell_base + (unsigned)((cz * P.n[1] + cy) * P.n[0] + cx);
        if (!FILL) {
            atomicAdd(&count[c], 1u);
        } else {
            const unsigned slot = start[c] + atomicSub(&count[c], 1u) - 1u;
            store_record(sorted, slot, img[0][vx], img[1][vy], img[2][vz], P.compact != 0);
        }
    };
    if (!any_shift) {
        if (!P.sym && ok[0][0] && ok[1][0] && ok[2][0]) place(0, 0, 0);
        return;
    }
#pragma unroll
    for (int vz = 0; vz < 3; ++vz) {
        if (!ok[2][vz]) continue;
#pragma unroll
        for (int vy = 0; vy < 3; ++vy) {
            if (!ok[1][vy]) continue;
#pragma unroll
            for (int vx = 0; vx < 3; ++vx) {
                if (!ok[0][vx]) continue;
                if (P.sym && (vx | vy | vz) == 0) continue;  // the unshifted image is the query itself
                place(vx, vy, vz);
            }
        }
    }
}

// Particles per thread of the placement kernels.  A thread loads all of them first and then issues
// its cell-counter atomics / offset gathers back to back, so that several L2 round trips are in
// flight per thread: with one particle per thread the kernels were latency bound (long-scoreboard
// stalls of 10-12 per issue, 21-27 % of the HBM rate; profiles/r01_place_v8_ncu_summary.txt).
#ifndef AMEPB_PLACE_ILP
#define AMEPB_PLACE_ILP 2
#endif
constexpr int kPlaceILP = AMEPB_PLACE_ILP;

template <typename T, bool FILL>
__global__ void __launch_bounds__(256) place_images_kernel(const T* __restrict__ pts, long long n,
                                                           const FramePlan* __restrict__ plans,
                                                           unsigned* __restrict__ count,
                                                           const unsigned* __restrict__ start,
                                                           Vec4<float>* __restrict__ sorted) {
    const int f = blockIdx.y;
    if (!plans[f].valid) return;
    __shared__ int s_words[kPlanHeadWords];
    const FramePlan& P = stage_plan_head(plans, f, s_words);
    const long long i0 = (long long)blockIdx.x * (blockDim.x * kPlaceILP) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        if (i < n) place_images_body<T, FILL>(pts + ((long long)f * n + i) * 3, P, count, start, sorted);
    }
}

// Self RDF with periodic images: queries and images of one particle in a single pass over the
// coordinates.  The count pass keeps each query's rank inside its cell, so the fill pass
// needs no second atomic for the queries.
template <typename T, typename AT, bool FILL>
__global__ void __launch_bounds__(256) place_self_kernel(const T* __restrict__ pts, long long n,
                                                         const FramePlan* __restrict__ plans,
                                                         unsigned* __restrict__ tcount,
                                                         const unsigned* __restrict__ tstart,
                                                         Vec4<float>* __restrict__ tsorted,
                                                         unsigned* __restrict__ qcount,
                                                         const unsigned* __restrict__ qstart,
                                                         Vec4<AT>* __restrict__ qsorted,
                                                         unsigned* __restrict__ qrank, int frame_fast) {
    // frame_fast (count pass): consecutive blocks belong to different frames, so the cell-counter
    // atomics in flight spread over the counters of the whole sub-batch (random L2 atomics run at
    // 126 G/s over >= 1M counters, at 100 G/s over the 62 500 of one configs[1] frame,
    // tools/place_probe.cu).  The fill pass keeps all blocks inside one frame: its scattered stores
    // must complete their sectors while the frame's 8-16 MB of sorted output are still in L2.
    const int f = frame_fast ? blockIdx.x : blockIdx.y;
    const unsigned tile = frame_fast ? blockIdx.y : blockIdx.x;
    if (!plans[f].valid) return;
    __shared__ int s_words[kPlanHeadWords];
    const long long i0 = (long long)tile * (blockDim.x * kPlaceILP) + threadIdx.x;
    const T* fbase = pts + (long long)f * n * 3;
    unsigned* frank = qrank + (long long)f * n;
    AT v[kPlaceILP][3];
    unsigned cc[kPlaceILP], aux[kPlaceILP];
    bool have[kPlaceILP];
    // (1) all loads of the thread's particles (consecutive threads read consecutive records),
    // issued before the plan is staged so that they are in flight across its barrier.
    // Measured alternatives (profiles/r02_place_variants.txt): 1 / 2 / 4 / 8 records per thread;
    // the block's span loaded element by element (one 128-byte line per warp load) into shared
    // memory and picked apart there -- slower (40 us per frame against 30): the extra barrier in
    // front of the dependent chain costs more than the three-lines-per-load pattern.
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        have[u] = i < n;
        const T* p = fbase + (have[u] ? i : 0) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) v[u][d] = (AT)p[d];
        if (FILL) aux[u] = have[u] ? frank[i] : 0u;
    }
    const FramePlan& P = stage_plan_head(plans, f, s_words);
    // (2) cells, then the dependent L2 accesses back to back
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        int c[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) c[d] = cell_index((double)v[u][d], P.g0[d], P.inv_h[d], P.n[d]);
        cc[u] = P.cell_base + (unsigned)((c[2] * P.n[1] + c[1]) * P.n[0] + c[0]);
    }
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        if (!have[u]) continue;
        if (!FILL) aux[u] = atomicAdd(&qcount[cc[u]], 1u);
        else aux[u] += __ldg(&qstart[cc[u]]);
    }
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        if (!have[u]) continue;
        const long long i = i0 + (long long)u * blockDim.x;
        if (!FILL) {
            frank[i] = aux[u];
        } else {
            store_record(qsorted, aux[u], v[u][0], v[u][1], v[u][2], P.compact != 0);
        }
    }
    // (3) shifted images: none for almost every particle when the cutoff is short.  In the symmetric
    // mode the test of place_images_body's shortcut is made here, on the values the thread already holds.
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        if (!have[u]) continue;
        if (P.sym) {
            bool maybe = false;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float b = (float)v[u][d];
                if (P.nshift[d] == 3) maybe = maybe || !(b > P.near_lo[d] && b < P.near_hi[d]);
            }
            if (!maybe) continue;
        }
        place_images_body<T, FILL>(fbase + i * 3, P, tcount, tstart, tsorted);
    }
}

// ---------------------------------------------------------------------------
// Partitioned build of the query array (self RDF with images, symmetric walk; round 2).
// place_self_kernel pays one returning L2 atomic per particle (count pass, 100-126 G/s) and one
// scattered store per particle that has to reach DRAM (fill pass, 75-123 G/s): 8-10 us + 8-13 us per
// 1M-particle frame whatever else the kernels do (tools/place_probe.cu).  Here no particle touches a
// global atomic and no store is scattered farther than its own strip:
//   strip_count_kernel    strips = runs of cell rows.  Counters of a tile of particles in shared memory,
//                         one global atomic per (tile, strip).  Shifted images of the few particles
//                         near a box face are counted as before (place_images_body).
//   strip_scan_kernel     exclusive scan of the strip counts of every frame
//   strip_scatter_kernel  ranks inside the tile from shared-memory atomics, one returning global atomic
//                         per (tile, strip) reserves a run of the strip's region, records go there;
//                         shifted images are filled
//   strip_sort_kernel     one block per (strip, frame): cell counters in shared memory -- count, scan,
//                         cursors --, the cell offsets of the strip written coalesced, the records
//                         scattered inside the strip's own few hundred kilobytes
// Prototype on 32 frames of 1M particles (tools/partition_probe.cu, profiles/r02_partition_probe.json):
// 2.5 + 0.5 + 7.4 + 6.1 = 16.5 us per frame against 24 us of count + scans + fill.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_warp, unsigned* total);

constexpr int kStripThreads = 256;
#ifndef AMEPB_STRIP_PER
#define AMEPB_STRIP_PER 8
#endif
#ifndef AMEPB_STRIP_BLOCKS
#define AMEPB_STRIP_BLOCKS 5      // resident blocks per SM the tile kernels are compiled for (48 registers; 4 and 6 measured slower)
#endif
constexpr int kStripPer = AMEPB_STRIP_PER;         // particles per thread of the two tile kernels
constexpr int kStripTile = kStripThreads * kStripPer;
constexpr int kMaxStrips = 512;
constexpr int kSortThreads = 256;
// records a thread keeps in registers between the two sweeps of strip_sort_kernel: 128 bytes' worth
// (16 (x, y) records cover a configs[1] strip of 4 k particles; measured 256 x 16: 20.6 us per frame for the
// whole build, 256 x 8: 21.1, 512 x 8: 21.0, 512 x 4: 20.8, 128 x 16: 21.7)
template <typename Rec>
struct SortCache {
    static constexpr int value = sizeof(Rec) <= 8 ? 16 : (sizeof(Rec) <= 16 ? 8 : 4);
};

// row / strip_rows by a multiply (strip_magic = 2^32 / strip_rows + 1; exact while rows * strip_rows < 2^31,
// which the host checks): an integer division per particle would cost as much as the rest of the tile kernel
__device__ __forceinline__ int strip_of(const FramePlan& P, int cy, int cz) {
    const unsigned row = (unsigned)(cz * P.n[1] + cy);
    const unsigned q = __umulhi(row, P.strip_magic);
    return (int)(P.strip_magic ? q : row);   // (a select, not a branch; magic 0 stands for strip_rows == 1)
}
// (not inlined: the image arithmetic would otherwise set the register count of the tile kernels)
template <typename T, bool FILL>
__device__ __noinline__ void place_images_listed(const T* p, const FramePlan* P, unsigned* count, const unsigned* start,
                                                 Vec4<float>* sorted) {
    place_images_body<T, FILL>(p, *P, count, start, sorted);
}

// record type of the intermediate (strip-ordered) and the sorted array: (x, y) pairs or four components
template <typename AT, bool XY>
struct StripRec;
template <>
struct StripRec<float, true> {
    typedef float2 type;
    static __device__ __forceinline__ float2 make(float x, float y, float) { return make_float2(x, y); }
    static __device__ __forceinline__ float z(const float2&) { return 0.f; }
};
template <>
struct StripRec<float, false> {
    typedef float4 type;
    static __device__ __forceinline__ float4 make(float x, float y, float z) { return make_float4(x, y, z, 0.f); }
    static __device__ __forceinline__ float z(const float4& v) { return v.z; }
};
template <>
struct StripRec<double, false> {
    typedef double4 type;
    static __device__ __forceinline__ double4 make(double x, double y, double z) { return make_double4(x, y, z, 0.0); }
    static __device__ __forceinline__ double z(const double4& v) { return v.z; }
};

// SCATTER = false: strip counts (+ count of the shifted images); true: records to their strips (+ fill of the images)
// FLAT: every frame of the launch has a single layer of cells (no cell arithmetic for z)
template <typename T, typename AT, bool XY, bool SCATTER, bool FLAT>
__global__ void __launch_bounds__(kStripThreads, AMEPB_STRIP_BLOCKS) strip_tile_kernel(const T* __restrict__ pts, long long n,
                                                                   const FramePlan* __restrict__ plans,
                                                                   unsigned* __restrict__ strip_count,
                                                                   unsigned* __restrict__ strip_cursor,
                                                                   typename StripRec<AT, XY>::type* __restrict__ mid,
                                                                   unsigned* __restrict__ tcount,
                                                                   const unsigned* __restrict__ tstart,
                                                                   Vec4<float>* __restrict__ tsorted) {
    const int f = blockIdx.y;
    __shared__ int s_words[kPlanHeadWords];
    __shared__ unsigned s_h[kMaxStrips], s_base[kMaxStrips];
    __shared__ unsigned short s_list[kStripTile];   // particles of the tile that may have a kept shifted image
    __shared__ unsigned s_nlist;
    if (threadIdx.x == 0) s_nlist = 0;
    const long long i0 = (long long)blockIdx.x * kStripTile + threadIdx.x;
    const T* fbase = pts + (long long)f * n * 3;
    AT v[kStripPer][3];
    unsigned st[kStripPer], rk[kStripPer];
#pragma unroll
    for (int u = 0; u < kStripPer; ++u) {
        const long long i = i0 + (long long)u * kStripThreads;
        const T* p = fbase + (i < n ? i : 0) * 3;
        // (x, y) records belong to flat frames without images along z: z is neither stored nor tested
#pragma unroll
        for (int d = 0; d < (XY ? 2 : 3); ++d) v[u][d] = (AT)p[d];
        if (XY) v[u][2] = (AT)0;
    }
    for (int k = threadIdx.x; k < kMaxStrips; k += kStripThreads) s_h[k] = 0;
    const FramePlan& P = stage_plan_head(plans, f, s_words);
    if (!P.valid) return;   // (block-uniform; asked only now so that the loads above are not behind a round trip)
    // (the plan values of the loop in registers: read through P, they were re-loaded from shared memory
    // for every particle)
    const double g0y = P.g0[1], ihy = P.inv_h[1], g0z = P.g0[2], ihz = P.inv_h[2];
    const int ny = P.n[1], nz = P.n[2];
    const unsigned magic = P.strip_magic;
    float nlo[3], nhi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // a dimension without shifted images never makes a particle a candidate
        const bool sh = P.nshift[d] == 3;
        nlo[d] = sh ? P.near_lo[d] : -__int_as_float(0x7f800000);
        nhi[d] = sh ? P.near_hi[d] : __int_as_float(0x7f800000);
    }
    unsigned near_mask = 0;   // bit u: particle u of this thread may have a kept shifted image
#pragma unroll
    for (int u = 0; u < kStripPer; ++u) {
        const long long i = i0 + (long long)u * kStripThreads;
        const int cy = cell_index((double)v[u][1], g0y, ihy, ny);
        const int cz = FLAT ? 0 : cell_index((double)v[u][2], g0z, ihz, nz);
        const unsigned row = (unsigned)(cz * ny + cy);
        st[u] = magic ? __umulhi(row, magic) : row;
        rk[u] = 0;
        if (i < n) {
            rk[u] = atomicAdd(&s_h[st[u]], 1u);
            bool maybe = false;
#pragma unroll
            for (int d = 0; d < (XY ? 2 : 3); ++d) {
                const float b = (float)v[u][d];
                maybe = maybe || !(b > nlo[d] && b < nhi[d]);   // (NaN: candidate)
            }
            if (maybe) near_mask |= 1u << u;
        }
    }
    // Shifted images (symmetric walk: the unshifted image is the query itself) exist for the few
    // particles near a box face only.  They are listed -- one shared atomic per thread that has any --
    // and handled by consecutive threads below: called in place, the image arithmetic ran for one or
    // two lanes of most warps, two thirds of all warp instructions of this kernel.
    if (near_mask) {
        unsigned at = atomicAdd(&s_nlist, (unsigned)__popc(near_mask));
#pragma unroll
        for (int u = 0; u < kStripPer; ++u)
            if (near_mask & (1u << u)) s_list[at++] = (unsigned short)(threadIdx.x + u * kStripThreads);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < P.nstrips; k += kStripThreads) {
        const unsigned c = s_h[k];
        if (!c) continue;
        if (!SCATTER) atomicAdd(&strip_count[f * kMaxStrips + k], c);
        else s_base[k] = atomicAdd(&strip_cursor[f * kMaxStrips + k], c);
    }
    if (SCATTER) {
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kStripPer; ++u) {
            const long long i = i0 + (long long)u * kStripThreads;
            if (i < n) mid[s_base[st[u]] + rk[u]] = StripRec<AT, XY>::make(v[u][0], v[u][1], v[u][2]);
        }
    }
    const unsigned nl = s_nlist;   // (complete since the barrier after the ranks)
    for (unsigned e = threadIdx.x; e < nl; e += kStripThreads)
        place_images_listed<T, SCATTER>(fbase + ((long long)blockIdx.x * kStripTile + s_list[e]) * 3, &P, tcount, tstart, tsorted);
}

// one warp per frame: exclusive scan of its strip counts from the frame's first query slot.  The strip
// arrays have a fixed stride per frame (kMaxStrips counters, kMaxStrips + 1 starts), so that the sort
// kernel can read its range before it knows the plan; unused entries hold the end of the frame's range
// (an invalid frame: zeros), i.e. empty strips.
__global__ void __launch_bounds__(32) strip_scan_kernel(const FramePlan* __restrict__ plans, const unsigned* __restrict__ strip_count,
                                                        unsigned* __restrict__ strip_cursor, unsigned* __restrict__ strip_start) {
    const int f = blockIdx.x;
    const FramePlan& P = plans[f];
    const int lane = threadIdx.x;
    const bool live = P.valid && P.strip_rows;
    const int ns = live ? P.nstrips : 0;
    unsigned carry = live ? P.qbase : 0u;
    for (int k0 = 0; k0 < kMaxStrips; k0 += 32) {
        const int k = k0 + lane;
        const unsigned c = k < ns ? strip_count[f * kMaxStrips + k] : 0u;
        unsigned inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        strip_start[f * (kMaxStrips + 1) + k] = carry + inc - c;
        strip_cursor[f * kMaxStrips + k] = carry + inc - c;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) strip_start[f * (kMaxStrips + 1) + kMaxStrips] = carry;
}

template <typename AT, bool XY, bool FLAT>
__global__ void __launch_bounds__(kSortThreads) strip_sort_kernel(const typename StripRec<AT, XY>::type* __restrict__ mid,
                                                                  const FramePlan* __restrict__ plans,
                                                                  const unsigned* __restrict__ strip_start, long long n,
                                                                  unsigned* __restrict__ qstart,
                                                                  typename StripRec<AT, XY>::type* __restrict__ sorted) {
    typedef typename StripRec<AT, XY>::type Rec;
    constexpr int kSortCache = SortCache<Rec>::value;
    extern __shared__ unsigned s_cnt[];   // [cells of the strip]: counts, then cursors
    __shared__ unsigned s_warp[33];
    __shared__ int s_words[kPlanHeadWords];
    const int f = blockIdx.y, s = blockIdx.x;
    // the strip's range and its first records are requested before the plan is staged (fixed-stride strip
    // arrays, see strip_scan_kernel): two dependent round trips per block instead of four
    const unsigned a = strip_start[f * (kMaxStrips + 1) + s], b = strip_start[f * (kMaxStrips + 1) + s + 1];
    Rec cv[kSortCache];
    int cc[kSortCache];
#pragma unroll
    for (int u = 0; u < kSortCache; ++u) {
        const unsigned i = a + threadIdx.x + u * kSortThreads;
        if (i < b) cv[u] = mid[i];
        else cv[u] = StripRec<AT, XY>::make((AT)0, (AT)0, (AT)0);
    }
    const FramePlan& P = stage_plan_head(plans, f, s_words);
    if (!P.valid || s >= P.nstrips) return;   // (block-uniform)
    const int nrows_total = P.n[1] * P.n[2];
    const int row0 = s * P.strip_rows;
    const int nrows = min(P.strip_rows, nrows_total - row0);
    const int nx = P.n[0];
    const int ncell = nrows * nx;
    for (int k = threadIdx.x; k < ncell; k += kSortThreads) s_cnt[k] = 0;
    auto local_cell = [&](const Rec& r) {
        const int cx = cell_index((double)r.x, P.g0[0], P.inv_h[0], P.n[0]);
        const int cy = cell_index((double)r.y, P.g0[1], P.inv_h[1], P.n[1]);
        const int cz = FLAT ? 0 : cell_index((double)StripRec<AT, XY>::z(r), P.g0[2], P.inv_h[2], P.n[2]);
        return (cz * P.n[1] + cy - row0) * nx + cx;
    };
    // (the first kSortCache * kSortThreads records of the strip stay in registers for the second sweep)
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kSortCache; ++u) {
        const unsigned i = a + threadIdx.x + u * kSortThreads;
        cc[u] = 0;
        if (i < b) {
            cc[u] = local_cell(cv[u]);
            atomicAdd(&s_cnt[cc[u]], 1u);
        }
    }
    for (unsigned i = a + threadIdx.x + kSortCache * kSortThreads; i < b; i += kSortThreads) atomicAdd(&s_cnt[local_cell(mid[i])], 1u);
    __syncthreads();
    // exclusive scan of the counters (a contiguous chunk per thread), offsets from the strip's first slot
    const int per = (ncell + kSortThreads - 1) / kSortThreads;
    const int k0 = min(ncell, (int)threadIdx.x * per), k1 = min(ncell, k0 + per);
    unsigned sum = 0;
    for (int k = k0; k < k1; ++k) sum += s_cnt[k];
    unsigned total;
    unsigned off = a + block_exclusive_scan(sum, s_warp, &total);
    unsigned* cs = qstart + P.cell_base + (unsigned)row0 * (unsigned)nx;
    for (int k = k0; k < k1; ++k) {
        const unsigned c = s_cnt[k];
        cs[k] = off;
        s_cnt[k] = off;   // cursor
        off += c;
    }
    // the entry behind the frame's last cell (the first one of the next frame, or the end of the array)
    if (s == P.nstrips - 1 && threadIdx.x == 0) qstart[P.cell_base + P.ncells] = b;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kSortCache; ++u) {
        const unsigned i = a + threadIdx.x + u * kSortThreads;
        if (i < b) sorted[atomicAdd(&s_cnt[cc[u]], 1u)] = cv[u];
    }
    for (unsigned i = a + threadIdx.x + kSortCache * kSortThreads; i < b; i += kSortThreads) {
        const Rec r = mid[i];
        sorted[atomicAdd(&s_cnt[local_cell(r)], 1u)] = r;
    }
}

// Plain points (queries, or targets without periodic images).  S: storage
// type.  DROP_OUTSIDE: skip points outside the grid region (targets); queries
// are clamped into the boundary cells instead.
template <typename T, typename S, bool FILL, bool DROP_OUTSIDE>
__global__ void __launch_bounds__(256) place_points_kernel(const T* __restrict__ pts, long long n,
                                                           const FramePlan* __restrict__ plans,
                                                           unsigned* __restrict__ count,
                                                           const unsigned* __restrict__ start,
                                                           Vec4<S>* __restrict__ sorted) {
    const int f = blockIdx.y;
    if (!plans[f].valid) return;
    __shared__ int s_words[kPlanHeadWords];
    const FramePlan& P = stage_plan_head(plans, f, s_words);
    const long long i0 = (long long)blockIdx.x * (blockDim.x * kPlaceILP) + threadIdx.x;
    const T* fbase = pts + (long long)f * n * 3;
    S v[kPlaceILP][3];
    unsigned cc[kPlaceILP], slot[kPlaceILP];
    bool keep[kPlaceILP];
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        keep[u] = i < n;
        const T* p = fbase + (keep[u] ? i : 0) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) v[u][d] = (S)p[d];
    }
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        int c[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double xd = (double)v[u][d];
            if (DROP_OUTSIDE) keep[u] = keep[u] && (xd >= P.reg_lo[d]) && (xd <= P.reg_hi[d]);
            c[d] = cell_index(xd, P.g0[d], P.inv_h[d], P.n[d]);
        }
        cc[u] = P.cell_base + (unsigned)((c[2] * P.n[1] + c[1]) * P.n[0] + c[0]);
    }
#pragma unroll
    for (int u = 0; u < kPlaceILP; ++u) {
        if (!keep[u]) continue;
        if (!FILL) atomicAdd(&count[cc[u]], 1u);
        else slot[u] = atomicSub(&count[cc[u]], 1u) - 1u;
    }
    if constexpr (FILL) {
#pragma unroll
        for (int u = 0; u < kPlaceILP; ++u)
            if (keep[u]) slot[u] += __ldg(&start[cc[u]]);
#pragma unroll
        for (int u = 0; u < kPlaceILP; ++u) {
            if (!keep[u]) continue;
            store_record(sorted, slot[u], v[u][0], v[u][1], v[u][2], P.compact != 0);
        }
    }
}

// ---------------------------------------------------------------------------
// exclusive scan (three small kernels; arrays hold at most a few million cells)
// ---------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanBlock * kScanPerThread;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_warp, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;                 // exclusive warp offsets
        if (lane == 31) s_warp[32] = winc;       // block total
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_local_kernel(const unsigned* __restrict__ in,
                                                                unsigned* __restrict__ out, long long len,
                                                                unsigned* __restrict__ block_sums) {
    __shared__ unsigned s_warp[33];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanPerThread;
    unsigned v[kScanPerThread];
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        v[k] = (base + k < len) ? in[base + k] : 0u;
        sum += v[k];
    }
    unsigned total;
    unsigned off = block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        if (base + k < len) out[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned* __restrict__ block_sums, int nblocks) {
    __shared__ unsigned s_warp[33];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nblocks ? block_sums[i] : 0u;
        unsigned total;
        const unsigned off = block_exclusive_scan(v, s_warp, &total);
        const unsigned carry = s_carry;
        if (i < nblocks) block_sums[i] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(unsigned* __restrict__ out, long long len,
                                                              const unsigned* __restrict__ block_sums) {
    const unsigned add = block_sums[blockIdx.x];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanPerThread;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k)
        if (base + k < len) out[base + k] += add;
}

// ---------------------------------------------------------------------------
// pair kernel
// ---------------------------------------------------------------------------
template <typename AT>
struct Arith;
template <>
struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    // float32 approximation of s (only feeds the bin guess)
    static __device__ __forceinline__ float approx(float s) { return s; }
    static __device__ __forceinline__ float load_thr(unsigned addr) {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
    }
    static constexpr int kThrShift = 2;
};
template <>
struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    // Truncating double->float on the integer pipe (F2F runs at 1/4 of the FP64 rate).
    // Exact enough for the bin guess when s is in the float32 normal range.
    static __device__ __forceinline__ float approx(double s) {
        const unsigned hi = (unsigned)__double2hiint(s), lo = (unsigned)__double2loint(s);
        // saturate: s >= 2^128, inf and NaN (and, through the unsigned wrap, s < 2^-126) map to
        // a huge finite float, whose guess clamps to the last slot; tiny s is below thr_lo anyway
        const unsigned e = min(hi - 0x38000000u, 0x0fefffffu);
        return __uint_as_float(__funnelshift_l(lo, e, 3));
    }
    static __device__ __forceinline__ double load_thr(unsigned addr) {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
        return v;
    }
    static constexpr int kThrShift = 3;
};

// Bin guess without an F2I: one MUFU.SQRT and one FFMA whose constant term carries the
// round-to-integer magic number 1.5*2^23.  With x = (sqrt(s) - e0) / width it returns
// k = round(x) clamped to [0, kclamp] (a negative k wraps and clamps too).  The error of x
// stays far below 1/2 for nbins <= 2^17 and uniform edges (checked on the host), so
// the pair lies in bin k-1 or k: one comparison with the start of bin k settles it.
__device__ __forceinline__ unsigned bin_guess(float s_approx, float inv_w, float c_magic, unsigned kclamp) {
    float d;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s_approx));
    const float t = fmaf(d, inv_w, c_magic);
    return min((unsigned)(__float_as_int(t) - 0x4b400000), kclamp);
}

// Packed float32 pairs: sm_100 executes add/mul/fma on two floats of a 64-bit register pair
// in one issue slot (FADD2 / FMUL2 / FFMA2, each lane IEEE round-to-nearest like the scalar
// instruction; tools/f32x2_probe.cu).  The pair kernel is issue-bound, so the float32
// regime evaluates two queries per lane and step with them.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// a + b with a single rounding, written as a*one + b where `one` is 1.0f passed at run time:
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (also with --fmad=false), which would
// skip the rounding of the product that the reference's (diff**2).sum() performs.  An fma
// cannot absorb the multiply that feeds it, and a*1 is exact.
__device__ __forceinline__ f32x2 add2_unfused(f32x2 a, f32x2 b, f32x2 one2) { return fma2(a, one2, b); }

struct PairArgs {
    const FramePlan* plans;
    const unsigned* tstart;
    const unsigned* qstart;
    const void* tsorted;
    const void* qsorted;
    const void* thr;                // [tables][nbins + 1] of the arithmetic type
    unsigned long long* hist;       // [frames][nbins]
    unsigned long long* counters;   // [0] task counter, [1] candidate pairs, [2] images
    long long total_tasks;
    long long total_cells;
    int nbins;
    int nframes;
    int uniform;                    // every edge table is uniform (linspace-like) and fit for the one-comparison binning
    float one;                      // 1.0f, opaque to the compiler (see add2_unfused)
    int shard_index, shard_count;   // this launch takes the block tasks t with t % shard_count == shard_index
};

// float32 regime: every warp gathers the neighbourhood of its query cell into a shared buffer
// of kStageCapF32 targets (x[] | y[] (| z[])) and walks it in full 32-lane passes
#ifndef AMEPB_STAGE_CAP
#define AMEPB_STAGE_CAP 256
#endif
constexpr unsigned kStageCapF32 = AMEPB_STAGE_CAP;   // targets per warp and chunk (float32 regime; half of it in float64)
constexpr unsigned kQueryCapF32 = 64;    // queries per warp and chunk
constexpr unsigned kUnitLimit = 1u << 21;                // flush after 2^31 counted pairs (units of 1024)
constexpr int kPairThreads = 256;
#ifndef AMEPB_PAIR_BLOCKS
#define AMEPB_PAIR_BLOCKS 4      // float32 kernels: 64 registers (four targets per lane need them)
#endif
constexpr int kPairBlocksPerSM = AMEPB_PAIR_BLOCKS;
#ifndef AMEPB_PAIR_BLOCKS_F64
#define AMEPB_PAIR_BLOCKS_F64 5
#endif
constexpr int kPairBlocksPerSMF64 = AMEPB_PAIR_BLOCKS_F64;   // float64 arithmetic: register pairs

// ZMODE 0: three coordinates; 1: every z equal, dz*dz is the per-frame constant dz2c;
// 2: like 1 with dz2c == 0 (s + 0 == s, the addition is dropped)
template <typename AT, int ZMODE>
__device__ __forceinline__ AT pair_s(const Vec4<AT>& qv, AT nx, AT ny, AT nz, AT dz2c) {
    // diff = c - coords; (diff**2).sum(axis=1) : ((dx*dx + dy*dy) + dz*dz), no FMA
    const AT dx = Arith<AT>::sub(qv.x, nx);
    const AT dy = Arith<AT>::sub(qv.y, ny);
    AT s = Arith<AT>::add(Arith<AT>::mul(dx, dx), Arith<AT>::mul(dy, dy));
    if (ZMODE == 2) return s;
    if (ZMODE == 1) return Arith<AT>::add(s, dz2c);
    const AT dz = Arith<AT>::sub(qv.z, nz);
    return Arith<AT>::add(s, Arith<AT>::mul(dz, dz));
}

template <typename AT, bool XY_ONLY>
__device__ __forceinline__ Vec4<AT> load_query(const Vec4<AT>* p);
template <>
__device__ __forceinline__ Vec4<float> load_query<float, true>(const Vec4<float>* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    Vec4<float> r;
    r.x = v.x; r.y = v.y; r.z = 0.0f; r.w = 0.0f;
    return r;
}
template <>
__device__ __forceinline__ Vec4<float> load_query<float, false>(const Vec4<float>* p) { return load_vec4(p); }
template <>
__device__ __forceinline__ Vec4<double> load_query<double, true>(const Vec4<double>* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    Vec4<double> r;
    r.x = v.x; r.y = v.y; r.z = 0.0; r.w = 0.0;
    return r;
}
template <>
__device__ __forceinline__ Vec4<double> load_query<double, false>(const Vec4<double>* p) { return load_vec4(p); }

// One warp: all queries of one cell against one neighbour row (contiguous in the
// sorted image array).  Lanes hold images in registers, queries are broadcast loads.
// General version: threshold table anywhere, search loops, shared or global histogram.
// record j of a sorted array (compact: (x, y) pairs, see store_record; float32 arrays only)
__device__ __forceinline__ Vec4<float> load_record(const Vec4<float>* arr, unsigned j, bool compact) {
    if (!compact) return load_vec4(arr + j);
    const float2 v = __ldg(reinterpret_cast<const float2*>(arr) + j);
    Vec4<float> r;
    r.x = v.x; r.y = v.y; r.z = 0.0f; r.w = 0.0f;
    return r;
}
__device__ __forceinline__ Vec4<double> load_record(const Vec4<double>* arr, unsigned j, bool) { return load_vec4(arr + j); }

template <typename AT, typename NT, int ZMODE, bool DIRECT>
__device__ __noinline__ void warp_pairs_general(const Vec4<AT>* __restrict__ qsorted, unsigned qa, unsigned qb,
                                                const Vec4<NT>* __restrict__ tsorted, unsigned ta, unsigned tb,
                                                const AT* __restrict__ thr, int nbins, AT dz2c, float inv_w,
                                                float c_magic, unsigned weight, unsigned* __restrict__ s_hist,
                                                unsigned long long* __restrict__ g_hist, bool compact = false) {
    const int lane = threadIdx.x & 31;
    const AT nan = (AT)__int_as_float(0x7fc00000);
    const AT thr_lo = thr[0], thr_hi = thr[nbins];
    for (unsigned base = ta; base < tb; base += 32) {
        const unsigned j = base + lane;
        AT nx = nan, ny = nan, nz = nan;
        if (j < tb) {
            const Vec4<NT> nb = load_record(tsorted, j, compact);
            nx = (AT)nb.x; ny = (AT)nb.y; nz = (AT)nb.z;
        }
        for (unsigned q = qa; q < qb; ++q) {
            const Vec4<AT> qv = load_record(qsorted, q, compact);
            const AT s = pair_s<AT, ZMODE>(qv, nx, ny, nz, dz2c);
            if (s >= thr_lo && s < thr_hi) {
                int g = max((int)bin_guess(Arith<AT>::approx(s), inv_w, c_magic, (unsigned)nbins) - 1, 0);
                while (s < thr[g]) --g;
                while (s >= thr[g + 1]) ++g;
                if (DIRECT) atomicAdd(&g_hist[g], (unsigned long long)weight);
                else atomicAdd(&s_hist[g], weight);
            }
        }
    }
}

// Hot version.  Shared memory holds T[k], k = 0..nbins+1: T[0] = thr[0] (the smallest counted
// s), T[k] = thr[k] -- the squared distance at which bin k starts, T[nbins] = end of the last
// bin, T[nbins+1] = NaN.  A histogram half has nbins + 2 slots: slot 0 and slot nbins + 1 are
// dump slots, bin b lives in slot b + 1.  With the guess k the pair goes to slot
// k + (s >= T[k]): pairs below thr[0] have k == 0 (host check: fast_binning_ok) and fall into
// slot 0, pairs beyond the last edge, NaN and inf into slot nbins + 1.
// No branches in the loop body; kQueryUnroll queries are in flight per lane.
constexpr int kQueryUnroll = 4;
// Unconditional +1 (ATOMS.POPC.INC: lanes hitting the same address are merged by the
// hardware, measured in tools/microbench.cu).  Out-of-range pairs are sent to a dump slot
// instead of being predicated off: ptxas wraps a predicated shared atomic in a branch
// region, which costs more issue slots than the extra lanes cost wavefronts.
__device__ __forceinline__ void smem_inc(unsigned addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

template <typename AT>
__device__ __forceinline__ void slot_update(AT s, unsigned k, unsigned thr_addr, unsigned hist_addr) {
    const AT t = Arith<AT>::load_thr(thr_addr + (k << Arith<AT>::kThrShift));
    unsigned addr = hist_addr + (k << 2);
    if (s >= t) addr += 4;
    smem_inc(addr);
}

template <typename AT, int ZMODE>
__device__ __forceinline__ void pair_update(const Vec4<AT>& qv, AT nx, AT ny, AT nz, AT dz2c, float inv_w,
                                            float c_magic, unsigned kclamp, unsigned thr_addr,
                                            unsigned hist_addr) {
    const AT s = pair_s<AT, ZMODE>(qv, nx, ny, nz, dz2c);
    slot_update<AT>(s, bin_guess(Arith<AT>::approx(s), inv_w, c_magic, kclamp), thr_addr, hist_addr);
}

// Two queries (packed) against the lane's target.
template <int ZMODE>
__device__ __forceinline__ void pair_update2(f32x2 qx, f32x2 qy, f32x2 qz, float nx, float ny, float nz,
                                             float one, float dz2c, float inv_w, float c_magic, unsigned kclamp,
                                             unsigned thr_addr, unsigned hist_addr) {
    const f32x2 one2 = pack2(one, one);
    const f32x2 dx = sub2(qx, pack2(nx, nx)), dy = sub2(qy, pack2(ny, ny));
    f32x2 s2 = add2_unfused(mul2(dx, dx), mul2(dy, dy), one2);
    if (ZMODE == 1) s2 = add2_unfused(s2, pack2(dz2c, dz2c), one2);
    if (ZMODE == 0) {
        const f32x2 dz = sub2(qz, pack2(nz, nz));
        s2 = add2_unfused(mul2(dz, dz), s2, one2);
    }
    float s0, s1, d0, d1, t0, t1;
    unpack2(s2, s0, s1);
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d0) : "f"(s0));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d1) : "f"(s1));
    unpack2(fma2(pack2(d0, d1), pack2(inv_w, inv_w), pack2(c_magic, c_magic)), t0, t1);
    slot_update<float>(s0, min((unsigned)(__float_as_int(t0) - 0x4b400000), kclamp), thr_addr, hist_addr);
    slot_update<float>(s1, min((unsigned)(__float_as_int(t1) - 0x4b400000), kclamp), thr_addr, hist_addr);
}

__host__ __device__ inline size_t pair_thr_bytes(int nbins, size_t at_size) {
    return ((size_t)(nbins + 2) * at_size + 15) & ~(size_t)15;
}

struct FastArgs {
    float inv_w, c_magic, one;
    unsigned kclamp, thr_addr;
};

// Explicit shared-space accesses (32-bit addresses): the staging buffers are addressed from one
// per-warp base register instead of generic pointers the compiler keeps re-deriving.
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_v4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ f32x2 lds_b64(unsigned addr) {
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// staging buffer element of the arithmetic type
__device__ __forceinline__ float lds_at(unsigned addr, float) { return lds_f32(addr); }
__device__ __forceinline__ double lds_at(unsigned addr, double) { return lds_f64(addr); }
__device__ __forceinline__ void sts_at(unsigned addr, float v) { sts_f32(addr, v); }
__device__ __forceinline__ void sts_at(unsigned addr, double v) { sts_f64(addr, v); }

// float <-> int with the same ordering (for REDUX min/max over a warp)
__device__ __forceinline__ int float_order(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// All staged queries of a warp against the lane's target, float32 regime.  The queries sit in
// shared memory as (x[2i], x[2i+1], y[2i], y[2i+1]) quads (+ z pairs), padded with NaN to an
// even count nq2: one broadcast LDS.128 feeds two packed pair evaluations.
template <int ZMODE>
__device__ __forceinline__ void query_loop_f32(const FastArgs& fa, unsigned sq_xy, unsigned sq_z, unsigned nq2,
                                               float nx, float ny, float nz, float dz2c, unsigned hist_addr) {
    auto packed = [&](const float4& xy, f32x2 z) {
        pair_update2<ZMODE>(pack2(xy.x, xy.y), pack2(xy.z, xy.w), z, nx, ny, nz, fa.one, dz2c, fa.inv_w, fa.c_magic,
                            fa.kclamp, fa.thr_addr, hist_addr);
    };
    unsigned u = 0;  // in pairs of queries: 16 bytes of sq_xy, 8 bytes of sq_z each
    const unsigned np = nq2 >> 1;
#pragma unroll 2
    for (; u + 2 <= np; u += 2) {
        const float4 xy0 = lds_v4(sq_xy + u * 16), xy1 = lds_v4(sq_xy + u * 16 + 16);
        f32x2 z0 = 0ull, z1 = 0ull;
        if (ZMODE == 0) {
            z0 = lds_b64(sq_z + u * 8);
            z1 = lds_b64(sq_z + u * 8 + 8);
        }
        packed(xy0, z0);
        packed(xy1, z1);
    }
    if (u < np) packed(lds_v4(sq_xy + u * 16), ZMODE == 0 ? lds_b64(sq_z + u * 8) : 0ull);
}

#ifndef AMEPB_TARGET_CASCADE
#define AMEPB_TARGET_CASCADE 1   // passes left over by the groups of four: two of them together (-1.5 % in 2D)
#endif
#ifndef AMEPB_QUERY_UNROLL
#define AMEPB_QUERY_UNROLL 2
#endif
constexpr int kMultiQueryUnroll = AMEPB_QUERY_UNROLL;
#ifndef AMEPB_TARGETS_PER_LANE
#define AMEPB_TARGETS_PER_LANE 4     // staged targets a lane holds while it walks the staged queries
#endif
constexpr int kTargetsPerLane = AMEPB_TARGETS_PER_LANE;
// Several targets per lane: every broadcast LDS.128 of a query pair feeds 2 * NT packed pair
// evaluations instead of two, and the loop overhead is shared.  Round 2: two targets per lane are
// 6-8 % faster than one on every configuration measured (2D uniform 6.07 -> 5.66 ms per 32 frames,
// clustered 0.332 -> 0.306 ms per frame, 3D 6.21 -> 5.82 ms per 2M-particle frame); four per lane
// at 64 registers and four blocks per SM another 3 % (5.53 / 0.295 / 5.60); 3, 6 or 8 per lane and
// 3 or 5 blocks per SM were slower (profiles/r02_targets_per_lane_sweep.txt).
template <int ZMODE, int NT>
__device__ __forceinline__ void query_loop_f32_multi(const FastArgs& fa, unsigned sq_xy, unsigned sq_z, unsigned nq2,
                                                     const float (&tx)[NT], const float (&ty)[NT], const float (&tz)[NT],
                                                     const unsigned (&th)[NT], float dz2c) {
    const unsigned np = nq2 >> 1;
#pragma unroll kMultiQueryUnroll
    for (unsigned u = 0; u < np; ++u) {
        const float4 xy = lds_v4(sq_xy + u * 16);
        const f32x2 z = ZMODE == 0 ? lds_b64(sq_z + u * 8) : 0ull;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            pair_update2<ZMODE>(pack2(xy.x, xy.y), pack2(xy.z, xy.w), z, tx[t], ty[t], tz[t], fa.one, dz2c, fa.inv_w,
                                fa.c_magic, fa.kclamp, fa.thr_addr, th[t]);
    }
}

template <typename AT, typename NT, int ZMODE>
__device__ __forceinline__ void warp_pairs_fast(const Vec4<AT>* __restrict__ qsorted, unsigned qa, unsigned qb,
                                                const Vec4<NT>* __restrict__ tsorted, unsigned ta, unsigned tb,
                                                AT dz2c, const FastArgs& fa, unsigned hist_lo, unsigned hist_hi,
                                                unsigned split) {
    // targets with index < split count into the histogram half at hist_lo, the others into
    // hist_hi (the symmetric pass mixes weight-1 own-cell pairs and weight-2 pairs in one range)
    const int lane = threadIdx.x & 31;
    const AT nan = (AT)__int_as_float(0x7fc00000);
    for (unsigned base = ta; base < tb; base += 32) {
        const unsigned j = base + lane;
        AT nx = nan, ny = nan, nz = nan;
        if (j < tb) {
            const Vec4<NT> nb = load_vec4(tsorted + j);
            nx = (AT)nb.x; ny = (AT)nb.y; nz = (AT)nb.z;
        }
        const unsigned hist_addr = j < split ? hist_lo : hist_hi;
        unsigned q = qa;
        {
            for (; q + kQueryUnroll <= qb; q += kQueryUnroll) {
                Vec4<AT> qv[kQueryUnroll];
#pragma unroll
                for (int u = 0; u < kQueryUnroll; ++u) qv[u] = load_query<AT, ZMODE != 0>(qsorted + q + u);
#pragma unroll
                for (int u = 0; u < kQueryUnroll; ++u)
                    pair_update<AT, ZMODE>(qv[u], nx, ny, nz, dz2c, fa.inv_w, fa.c_magic, fa.kclamp, fa.thr_addr,
                                           hist_addr);
            }
            for (; q < qb; ++q) {
                const Vec4<AT> qv = load_query<AT, ZMODE != 0>(qsorted + q);
                pair_update<AT, ZMODE>(qv, nx, ny, nz, dz2c, fa.inv_w, fa.c_magic, fa.kclamp, fa.thr_addr, hist_addr);
            }
        }
    }
}

// SYM: queries and unshifted targets are the same particles with identical stored values
// (self RDF in the float32 regime, or without periodic images).  Then s(i,j) == s(j,i) bit for
// bit, so real-real pairs are evaluated once over the upper half shell and counted twice; the
// shifted images ("ghosts", a separate cell-sorted array) are not symmetric in floating point
// and are evaluated for every ordered (query, ghost) pair.
template <typename AT, typename NT, int ZMODE, bool SMEM, bool SYM>
__global__ void __launch_bounds__(kPairThreads, (sizeof(AT) == 4 ? kPairBlocksPerSM : kPairBlocksPerSMF64)) pair_kernel(PairArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = a.nbins;
    AT* s_thr = reinterpret_cast<AT*>(smem_raw);  // T[0..nbins+1], see warp_pairs_fast
    unsigned* s_hist = reinterpret_cast<unsigned*>(smem_raw + pair_thr_bytes(nbins, sizeof(AT)));
    // SYM: a second histogram half collects the pairs that count twice, so that every
    // increment is +1 (ATOMS.POPC.INC); the halves are combined as h1 + 2*h2 at flush time.
    // A half has nbins + 2 slots: bin b in slot b + 1, slots 0 and nbins + 1 are dump slots.
    const int half_len = nbins + 2;
    const int hist_len = (SYM ? 2 : 1) * half_len;
    unsigned hist_addr = (unsigned)__cvta_generic_to_shared(s_hist);
    unsigned thr_base_addr = (unsigned)__cvta_generic_to_shared(s_thr);
    // (opaque values, see stage_addr below)
    asm volatile("mov.u32 %0, %0;" : "+r"(hist_addr));
    asm volatile("mov.u32 %0, %0;" : "+r"(thr_base_addr));
    const unsigned hist2_addr = hist_addr + 4u * (unsigned)half_len;
    constexpr unsigned kStageComp = ZMODE == 0 ? 3 : 2;
    constexpr unsigned kES = (unsigned)sizeof(AT);                             // staged element size
    constexpr unsigned kStageCap = kStageCapF32 * 4 / kES;                    // same bytes for both types
    constexpr unsigned kQueryCap = kQueryCapF32 * 4 / kES;
    constexpr unsigned kStageFloats = (kStageCapF32 + kQueryCapF32) * kStageComp;  // 4-byte words per warp
    // per-warp staging area of the float32 regime, as a shared-space address
    unsigned stage_addr = (unsigned)__cvta_generic_to_shared(smem_raw) +
                          (unsigned)(pair_thr_bytes(nbins, sizeof(AT)) + (((size_t)hist_len * sizeof(unsigned) + 15) & ~(size_t)15)) +
                          (threadIdx.x >> 5) * (unsigned)(kStageFloats * sizeof(float));
    // opaque to the compiler: under the 48-register cap it would otherwise re-derive this address
    // (a dozen instructions) inside the gather loops instead of keeping or spilling one value
    asm volatile("mov.u32 %0, %0;" : "+r"(stage_addr));
    __shared__ int s_next;
    __shared__ FramePlan s_plan;
    __shared__ long long s_task;
    __shared__ unsigned s_units;
    __shared__ int s_flush;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Vec4<AT>* qsorted = reinterpret_cast<const Vec4<AT>*>(a.qsorted);
    const Vec4<NT>* tsorted = reinterpret_cast<const Vec4<NT>*>(a.tsorted);
    int cur = -1;
    unsigned long long cand = 0;
    if (tid == 0) s_units = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) {
            s_task = (long long)atomicAdd(&a.counters[0], 1ull) * a.shard_count + a.shard_index;
            s_next = blockDim.x >> 5;
            const int fl = s_units >= kUnitLimit;
            s_flush = fl;
            if (fl) s_units = 0;
        }
        __syncthreads();
        const long long t = s_task;
        const bool done = t >= a.total_tasks;
        int f = cur < 0 ? 0 : cur;
        if (!done)
            while (t >= a.plans[f].task_begin + a.plans[f].ntasks) ++f;
        if (done || f != cur || s_flush) {
            if (SMEM && cur >= 0) {
                unsigned long long* row = a.hist + (long long)cur * nbins;
                for (int b = tid; b < nbins; b += blockDim.x) {
                    unsigned long long v = s_hist[b + 1];
                    if (SYM) v += 2ull * s_hist[half_len + b + 1];
                    if (v) atomicAdd(&row[b], v);
                }
                __syncthreads();  // the zeroing below maps bins to other threads than the loop above
            }
            if (done) break;
            if (f != cur) {
                const int* src = reinterpret_cast<const int*>(&a.plans[f]);
                int* dst = reinterpret_cast<int*>(&s_plan);
                for (int k = tid; k < (int)(sizeof(FramePlan) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
                if (SMEM) {
                    const AT* tsrc = reinterpret_cast<const AT*>(a.thr) + (long long)a.plans[f].thr_index * (nbins + 1);
                    for (int k = tid; k <= nbins + 1; k += blockDim.x)
                        s_thr[k] = k <= nbins ? tsrc[k] : (AT)__int_as_float(0x7fc00000);
                }
            }
            if (SMEM)
                for (int b = tid; b < hist_len; b += blockDim.x) s_hist[b] = 0;
            cur = f;
            __syncthreads();
        }
        const FramePlan& P = s_plan;
        const AT* thr = reinterpret_cast<const AT*>(a.thr) + (long long)P.thr_index * (nbins + 1);
        const AT dz2c = (AT)P.dz2c;
        // k = round(sqrt(s)*inv_w - e0*inv_w); the magic constant does the rounding
        const float inv_w = (float)P.bin_inv_w;
        const float c_magic = 12582912.0f - (float)(P.bin_e0 * P.bin_inv_w);
        FastArgs fa;
        fa.inv_w = inv_w;
        fa.c_magic = c_magic;
        fa.one = a.one;
        fa.kclamp = (unsigned)(nbins + 1);
        fa.thr_addr = thr_base_addr;
        unsigned long long* g_hist = a.hist + (long long)f * nbins;

        const long long lt = t - P.task_begin;
        const int run_i = (int)(lt % P.runs_per_row);
        const long long rowq = lt / P.runs_per_row;
        const int cy = P.cq_lo[1] + (int)(rowq % P.cq_n[1]);
        const int cz = P.cq_lo[2] + (int)(rowq / P.cq_n[1]);
        const int cx0 = P.cq_lo[0] + run_i * P.run;
        const int ncell = min(P.run, P.cq_lo[0] + P.cq_n[0] - cx0);
        const int nrows = P.nrows;
        const int row0 = P.row0;
        const unsigned long long direct_pairs = P.direct_pairs;
        const bool small_ok = direct_pairs >= (unsigned long long)kQueryCapF32 * kStageCapF32;
        // squared cutoff of the target culling (float32 regime), less the constant dz*dz of flat frames
        const float cull_r2 = P.cull_r2 - (ZMODE == 1 ? (float)P.dz2c : 0.0f);
        const int n0 = P.n[0], n1 = P.n[1], n2 = P.n[2];
        const unsigned qrow = P.cell_base + (unsigned)((cz * n1 + cy) * n0);

        // One warp per query cell of the run (rsplit == 1), or rsplit warps sharing a cell,
        // each walking every rsplit-th neighbour row.
        const int rsplit = P.rsplit;
        // the items of a block task are handed out dynamically, so that a warp with light cells
        // takes more of them instead of waiting at the barrier
        for (int wtk = warp; wtk < ncell * rsplit;
             wtk = __shfl_sync(0xffffffffu, lane == 0 ? atomicAdd(&s_next, 1) : 0, 0)) {
            const int part = rsplit == 1 ? 0 : wtk / ncell;
            const int cx = cx0 + (wtk - part * ncell);
            const unsigned qa = __ldg(&a.qstart[qrow + cx]), qb = __ldg(&a.qstart[qrow + cx + 1]);
            if (qa == qb) continue;
            if constexpr (SMEM && sizeof(AT) == 4) {
                // ---- float32 regime: stage queries and neighbourhood, then full passes ---------
                const unsigned bx = stage_addr;                     // targets x[] | y[] (| z[])
                const unsigned by = bx + 4 * kStageCap;
                const unsigned bz = by + 4 * kStageCap;             // ZMODE == 0 only
                const unsigned sq_xy = bx + 4 * kStageCap * kStageComp;  // queries, see query_loop_f32
                const unsigned sq_z = sq_xy + 8 * kQueryCap;        // ZMODE == 0 only
                const Vec4<float>* qarr = reinterpret_cast<const Vec4<float>*>(qsorted);
                const Vec4<float>* tarr = reinterpret_cast<const Vec4<float>*>(tsorted);
                const bool compact = P.compact != 0;            // flat frames: (x, y) records, see store_record
                const unsigned rec2 = compact ? 1u : 2u;        // record size in float2 units
                // entries: [0] own cell (SYM), [1, G] ghost rows (SYM, if any can be near),
                // then the real rows (SYM: upper half shell) -- weight-1 entries first
                int G = 0;
                if (SYM) {
                    const int kx0 = P.kd[0], ky = P.kd[1], kz = P.kd[2];
                    const bool gfree = cx - kx0 >= P.gfree_lo[0] && cx + kx0 <= P.gfree_hi[0] && cy - ky >= P.gfree_lo[1] &&
                                       cy + ky <= P.gfree_hi[1] && cz - kz >= P.gfree_lo[2] && cz + kz <= P.gfree_hi[2];
                    G = gfree ? 0 : nrows;
                }
                const int E = SYM ? 1 + G + (nrows - row0) : nrows;
                unsigned long long staged = 0, oversize = 0;
                for (unsigned qc = qa; qc < qb; qc += kQueryCap) {
                    const unsigned nqc = min(kQueryCap, qb - qc);
                    const unsigned nq2 = (nqc + 1u) & ~1u;
                    __syncwarp();
                    // bounding box of the staged queries (NaN coordinates never pair and are left out)
                    const float finf = __int_as_float(0x7f800000);
                    float lo[3] = {finf, finf, finf}, hi[3] = {-finf, -finf, -finf};
#pragma unroll 1
                    for (unsigned i = lane; i < nq2; i += 32) {
                        float x = __int_as_float(0x7fc00000), y = x, z = x;
                        if (i < nqc) {
                            if (ZMODE == 0) {
                                const Vec4<float> v = load_vec4(qarr + qc + i);
                                x = v.x; y = v.y; z = v.z;
                            } else {
                                const float2 v = __ldg(reinterpret_cast<const float2*>(qarr) + (size_t)(qc + i) * rec2);
                                x = v.x; y = v.y;
                            }
                            if (x == x) { lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x); }
                            if (y == y) { lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y); }
                            if (ZMODE == 0 && z == z) { lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z); }
                        }
                        const unsigned e = sq_xy + (i >> 1) * 16 + (i & 1u) * 4;
                        sts_f32(e, x);
                        sts_f32(e + 8, y);
                        if (ZMODE == 0) sts_f32(sq_z + 4 * i, z);
                    }
#pragma unroll
                    for (int d = 0; d < (ZMODE == 0 ? 3 : 2); ++d) {
                        lo[d] = order_float(__reduce_min_sync(0xffffffffu, float_order(lo[d])));
                        hi[d] = order_float(__reduce_max_sync(0xffffffffu, float_order(hi[d])));
                    }
                    unsigned fill = 0, lo_end = 0;
                    auto flush_stage = [&]() {
                        __syncwarp();
                        unsigned g0 = 0;
                        if constexpr (kTargetsPerLane > 1) {
                            // full groups of kTargetsPerLane passes (the last one may be ragged: NaN targets)
                            for (; g0 + 32 * (kTargetsPerLane - 1) < fill; g0 += 32 * kTargetsPerLane) {
                                float tx[kTargetsPerLane], ty[kTargetsPerLane], tz[kTargetsPerLane];
                                unsigned th[kTargetsPerLane];
#pragma unroll
                                for (int t = 0; t < kTargetsPerLane; ++t) {
                                    const unsigned g = g0 + 32 * t + lane;
                                    tx[t] = ty[t] = tz[t] = __int_as_float(0x7fc00000);
                                    if (g < fill) {
                                        tx[t] = lds_f32(bx + 4 * g);
                                        ty[t] = lds_f32(by + 4 * g);
                                        if (ZMODE == 0) tz[t] = lds_f32(bz + 4 * g);
                                    }
                                    th[t] = (!SYM || g < lo_end) ? hist_addr : hist2_addr;
                                }
                                query_loop_f32_multi<ZMODE, kTargetsPerLane>(fa, sq_xy, sq_z, nq2, tx, ty, tz, th, dz2c);
                            }
#if AMEPB_TARGET_CASCADE
                            if (kTargetsPerLane > 2 && g0 + 32 < fill) {   // two of the passes that are left, together
                                float tx[2], ty[2], tz[2];
                                unsigned th[2];
#pragma unroll
                                for (int t = 0; t < 2; ++t) {
                                    const unsigned g = g0 + 32 * t + lane;
                                    tx[t] = ty[t] = tz[t] = __int_as_float(0x7fc00000);
                                    if (g < fill) {
                                        tx[t] = lds_f32(bx + 4 * g);
                                        ty[t] = lds_f32(by + 4 * g);
                                        if (ZMODE == 0) tz[t] = lds_f32(bz + 4 * g);
                                    }
                                    th[t] = (!SYM || g < lo_end) ? hist_addr : hist2_addr;
                                }
                                query_loop_f32_multi<ZMODE, 2>(fa, sq_xy, sq_z, nq2, tx, ty, tz, th, dz2c);
                                g0 += 64;
                            }
#endif
                        }
                        for (; g0 < fill; g0 += 32) {
                            const unsigned g = g0 + lane;
                            float nx = __int_as_float(0x7fc00000), ny = nx, nz = nx;
                            if (g < fill) {
                                nx = lds_f32(bx + 4 * g);
                                ny = lds_f32(by + 4 * g);
                                if (ZMODE == 0) nz = lds_f32(bz + 4 * g);
                            }
                            // (idle lanes hold NaN targets and end up in a dump slot: without a second
                            // half that must be the first half's)
                            query_loop_f32<ZMODE>(fa, sq_xy, sq_z, nq2, nx, ny, nz, dz2c,
                                                  (!SYM || g < lo_end) ? hist_addr : hist2_addr);
                        }
                        __syncwarp();
                        fill = 0;
                        lo_end = 0;
                    };
                    // Gather n targets into the staging buffer behind position `at`, dropping those
                    // farther than the cutoff from the box around the staged queries: no query of this
                    // round can reach them (cull_r2 carries a 1e-4 relative margin, far above the
                    // float32 rounding of either side).  Returns the new fill level.
                    auto copy_targets = [&](const Vec4<float>* src0, unsigned first, unsigned n, unsigned at, bool cull) -> unsigned {
#pragma unroll 1
                        for (unsigned base = 0; base < n; base += 32) {
                            const unsigned i = base + lane;
                            float x = 0.f, y = 0.f, z = 0.f;
                            bool keep = i < n;
                            if (keep) {
                                if (ZMODE == 0) {
                                    const Vec4<float> v = load_vec4(src0 + first + i);
                                    x = v.x; y = v.y; z = v.z;
                                } else {
                                    const float2 v = __ldg(reinterpret_cast<const float2*>(src0) + (size_t)(first + i) * rec2);
                                    x = v.x; y = v.y;
                                }
                                if (cull) {
                                    const float ex = fmaxf(fmaxf(lo[0] - x, x - hi[0]), 0.f);
                                    const float ey = fmaxf(fmaxf(lo[1] - y, y - hi[1]), 0.f);
                                    float d2 = ex * ex + ey * ey;
                                    if (ZMODE == 0) {
                                        const float ez = fmaxf(fmaxf(lo[2] - z, z - hi[2]), 0.f);
                                        d2 += ez * ez;
                                    }
                                    keep = d2 <= cull_r2;   // false for NaN targets as well
                                }
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, keep);
                            if (keep) {
                                const unsigned o = 4 * (at + __popc(m & ((1u << lane) - 1u)));
                                sts_f32(bx + o, x);
                                sts_f32(by + o, y);
                                if (ZMODE == 0) sts_f32(bz + o, z);
                            }
                            at += __popc(m);
                        }
                        return at;
                    };
                    for (int eb = 0; eb < E; eb += 32) {
                        const int e = eb + lane;
                        unsigned start = 0, len = 0, flags = 0;  // flags bit 0: counts twice, bit 1: lives in the query array
                        if (e < E) {
                            int r = -1;
                            if (SYM) {
                                if (e == 0) {
                                    flags = 2;
                                    if (part == 0) {
                                        start = qa;
                                        len = qb - qa;
                                    }
                                } else if (e <= G) {
                                    r = e - 1;
                                    if (rsplit > 1 && r % rsplit != part) r = -1;
                                } else {
                                    r = row0 + (e - 1 - G);
                                    flags = 3;
                                    if (rsplit > 1 && (r - row0) % rsplit != part) r = -1;
                                }
                            } else {
                                r = e;
                                if (rsplit > 1 && r % rsplit != part) r = -1;
                            }
                            if (r >= 0) {
                                const int ny = cy + P.row_oy[r], nz = cz + P.row_oz[r];
                                if (ny >= 0 && ny < n1 && nz >= 0 && nz < n2) {
                                    const int kx = P.row_kx[r];
                                    int x_lo = max(cx - kx, 0);
                                    const int x_hi = min(cx + kx, n0 - 1);
                                    if (flags == 3 && r == row0) x_lo = cx + 1;  // own row: only the cells to the right
                                    if (x_lo <= x_hi) {
                                        const unsigned* starts = (flags & 2) ? a.qstart : a.tstart;
                                        const unsigned trow = P.cell_base + (unsigned)((nz * n1 + ny) * n0);
                                        start = __ldg(&starts[trow + x_lo]);
                                        len = __ldg(&starts[trow + x_hi + 1]) - start;
                                    }
                                }
                            }
                        }
                        unsigned todo = __ballot_sync(0xffffffffu, len != 0);
                        while (todo) {
                            const int src = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const unsigned st = __shfl_sync(0xffffffffu, start, src), ln = __shfl_sync(0xffffffffu, len, src);
                            const unsigned fl = __shfl_sync(0xffffffffu, flags, src);
                            const Vec4<float>* arr0 = (fl & 2) ? qarr : tarr;
                            const bool cull = fl != 2;   // the own cell holds the queries themselves
                            if (fill + ln <= kStageCap && small_ok) {  // the usual case: the row fits
                                const unsigned before = fill;
                                fill = copy_targets(arr0, st, ln, fill, cull);
                                if (!(fl & 1)) lo_end = fill;
                                staged += nqc * (fill - before);   // at most kQueryCap * kStageCap
                                continue;
                            }
                            const unsigned long long npairs = (unsigned long long)nqc * ln;
                            if (npairs > direct_pairs) {  // giant cells: straight to the global histogram
                                oversize += npairs;
                                warp_pairs_general<float, float, ZMODE == 0 ? 0 : 1, true>(
                                    qarr, qc, qc + nqc, arr0, st, st + ln, reinterpret_cast<const float*>(thr), nbins,
                                    (float)dz2c, inv_w, c_magic, (fl & 1) ? 2u : 1u, s_hist, g_hist, ZMODE != 0 && compact);
                                continue;
                            }
                            unsigned off = 0;
                            while (off < ln) {
                                const unsigned n = min(ln - off, kStageCap - fill);
                                const unsigned before = fill;
                                fill = copy_targets(arr0, st + off, n, fill, cull);
                                staged += (unsigned long long)nqc * (fill - before);
                                off += n;
                                if (!(fl & 1)) lo_end = fill;
                                if (fill == kStageCap) flush_stage();
                            }
                        }
                    }
                    if (fill) flush_stage();
                }
                if (lane == 0) {
                    cand += staged + oversize;
                    if (staged) atomicAdd(&s_units, (unsigned)(staged >> 10) + 1u);
                }
            } else if constexpr (SMEM) {
                // ---- float64 regimes: the same staged walk with scalar float64 arithmetic.  (Kept apart from
                // the float32 block above on purpose: a shared generic version cost the float32 kernels 2 %.)
                const unsigned bx = stage_addr;                        // targets x[] | y[] (| z[]) of type AT
                const unsigned by = bx + kES * kStageCap;
                const unsigned bz = by + kES * kStageCap;              // ZMODE == 0 only
                // queries: float32: packed pairs, see query_loop_f32; float64: x[] | y[] (| z[])
                const unsigned sq_xy = bx + kES * kStageCap * kStageComp;
                const unsigned sq_z = sq_xy + 2 * kES * kQueryCap;     // ZMODE == 0 only
                const Vec4<AT>* qarr = qsorted;
                const Vec4<NT>* tarr = tsorted;
                // entries: [0] own cell (SYM), [1, G] ghost rows (SYM, if any can be near),
                // then the real rows (SYM: upper half shell) -- weight-1 entries first
                int G = 0;
                if (SYM) {
                    const int kx0 = P.kd[0], ky = P.kd[1], kz = P.kd[2];
                    const bool gfree = cx - kx0 >= P.gfree_lo[0] && cx + kx0 <= P.gfree_hi[0] && cy - ky >= P.gfree_lo[1] &&
                                       cy + ky <= P.gfree_hi[1] && cz - kz >= P.gfree_lo[2] && cz + kz <= P.gfree_hi[2];
                    G = gfree ? 0 : nrows;
                }
                const int E = SYM ? 1 + G + (nrows - row0) : nrows;
                unsigned long long staged = 0, oversize = 0;
                for (unsigned qc = qa; qc < qb; qc += kQueryCap) {
                    const unsigned nqc = min(kQueryCap, qb - qc);
                    const unsigned nq2 = (nqc + 1u) & ~1u;
                    __syncwarp();
                    // bounding box of the staged queries in float32, rounded outwards (NaN coordinates
                    // never pair and are left out)
                    const float finf = __int_as_float(0x7f800000);
                    float lo[3] = {finf, finf, finf}, hi[3] = {-finf, -finf, -finf};
                    auto grow = [&](int d, AT v) {
                        if (v == v) {
                            if constexpr (sizeof(AT) == 4) {
                                lo[d] = fminf(lo[d], v);
                                hi[d] = fmaxf(hi[d], v);
                            } else {
                                lo[d] = fminf(lo[d], __double2float_rd(v));
                                hi[d] = fmaxf(hi[d], __double2float_ru(v));
                            }
                        }
                    };
#pragma unroll 1
                    for (unsigned i = lane; i < nq2; i += 32) {
                        AT x = (AT)__int_as_float(0x7fc00000), y = x, z = x;
                        if (i < nqc) {
                            const Vec4<AT> v = load_query<AT, ZMODE != 0>(qarr + qc + i);
                            x = v.x; y = v.y; z = v.z;
                            grow(0, x);
                            grow(1, y);
                            if (ZMODE == 0) grow(2, z);
                        }
                        if constexpr (sizeof(AT) == 4) {
                            const unsigned e = sq_xy + (i >> 1) * 16 + (i & 1u) * 4;
                            sts_f32(e, x);
                            sts_f32(e + 8, y);
                            if (ZMODE == 0) sts_f32(sq_z + 4 * i, z);
                        } else if (i < nqc) {
                            sts_f64(sq_xy + 8 * i, x);
                            sts_f64(sq_xy + 8 * kQueryCap + 8 * i, y);
                            if (ZMODE == 0) sts_f64(sq_z + 8 * i, z);
                        }
                    }
#pragma unroll
                    for (int d = 0; d < (ZMODE == 0 ? 3 : 2); ++d) {
                        lo[d] = order_float(__reduce_min_sync(0xffffffffu, float_order(lo[d])));
                        hi[d] = order_float(__reduce_max_sync(0xffffffffu, float_order(hi[d])));
                    }
                    unsigned fill = 0, lo_end = 0;
                    auto flush_stage = [&]() {
                        __syncwarp();
                        for (unsigned g0 = 0; g0 < fill; g0 += 32) {
                            const unsigned g = g0 + lane;
                            AT nx = (AT)__int_as_float(0x7fc00000), ny = nx, nz = nx;
                            if (g < fill) {
                                nx = lds_at(bx + kES * g, AT());
                                ny = lds_at(by + kES * g, AT());
                                if (ZMODE == 0) nz = lds_at(bz + kES * g, AT());
                            }
                            // (idle lanes hold NaN targets and end up in a dump slot: without a second
                            // half that must be the first half's)
                            const unsigned h_addr = (!SYM || g < lo_end) ? hist_addr : hist2_addr;
                            if constexpr (sizeof(AT) == 4) {
                                query_loop_f32<ZMODE>(fa, sq_xy, sq_z, nq2, nx, ny, nz, dz2c, h_addr);
                            } else {
#pragma unroll 2
                                for (unsigned u = 0; u < nqc; ++u) {
                                    Vec4<double> qv;
                                    qv.x = lds_f64(sq_xy + 8 * u);
                                    qv.y = lds_f64(sq_xy + 8 * kQueryCap + 8 * u);
                                    qv.z = ZMODE == 0 ? lds_f64(sq_z + 8 * u) : 0.0;
                                    qv.w = 0.0;
                                    pair_update<double, ZMODE>(qv, nx, ny, nz, dz2c, fa.inv_w, fa.c_magic, fa.kclamp,
                                                               fa.thr_addr, h_addr);
                                }
                            }
                        }
                        __syncwarp();
                        fill = 0;
                        lo_end = 0;
                    };
                    // Gather n targets into the staging buffer behind position `at`, dropping those
                    // farther than the cutoff from the box around the staged queries: no query of this
                    // round can reach them (cull_r2 carries a 1e-4 relative margin, far above the
                    // rounding of either side; the float64 regime tests against the outward-rounded
                    // float32 box in float64).  Returns the new fill level.
                    auto copy_targets = [&](auto* src, unsigned n, unsigned at, bool cull) -> unsigned {
                        typedef decltype(src->x) ST;   // element type of the source array
#pragma unroll 1
                        for (unsigned base = 0; base < n; base += 32) {
                            const unsigned i = base + lane;
                            AT x = (AT)0, y = (AT)0, z = (AT)0;
                            bool keep = i < n;
                            if (keep) {
                                const Vec4<ST> v = load_query<ST, ZMODE != 0>(src + i);
                                x = (AT)v.x; y = (AT)v.y; z = (AT)v.z;
                                if (cull) {
                                    const AT ex = max(max((AT)lo[0] - x, x - (AT)hi[0]), (AT)0);
                                    const AT ey = max(max((AT)lo[1] - y, y - (AT)hi[1]), (AT)0);
                                    AT d2 = ex * ex + ey * ey;
                                    if (ZMODE == 0) {
                                        const AT ez = max(max((AT)lo[2] - z, z - (AT)hi[2]), (AT)0);
                                        d2 += ez * ez;
                                    }
                                    keep = d2 <= (AT)cull_r2;   // false for NaN targets as well
                                }
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, keep);
                            if (keep) {
                                const unsigned o = kES * (at + __popc(m & ((1u << lane) - 1u)));
                                sts_at(bx + o, x);
                                sts_at(by + o, y);
                                if (ZMODE == 0) sts_at(bz + o, z);
                            }
                            at += __popc(m);
                        }
                        return at;
                    };
                    for (int eb = 0; eb < E; eb += 32) {
                        const int e = eb + lane;
                        unsigned start = 0, len = 0, flags = 0;  // flags bit 0: counts twice, bit 1: lives in the query array
                        if (e < E) {
                            int r = -1;
                            if (SYM) {
                                if (e == 0) {
                                    flags = 2;
                                    if (part == 0) {
                                        start = qa;
                                        len = qb - qa;
                                    }
                                } else if (e <= G) {
                                    r = e - 1;
                                    if (rsplit > 1 && r % rsplit != part) r = -1;
                                } else {
                                    r = row0 + (e - 1 - G);
                                    flags = 3;
                                    if (rsplit > 1 && (r - row0) % rsplit != part) r = -1;
                                }
                            } else {
                                r = e;
                                if (rsplit > 1 && r % rsplit != part) r = -1;
                            }
                            if (r >= 0) {
                                const int ny = cy + P.row_oy[r], nz = cz + P.row_oz[r];
                                if (ny >= 0 && ny < n1 && nz >= 0 && nz < n2) {
                                    const int kx = P.row_kx[r];
                                    int x_lo = max(cx - kx, 0);
                                    const int x_hi = min(cx + kx, n0 - 1);
                                    if (flags == 3 && r == row0) x_lo = cx + 1;  // own row: only the cells to the right
                                    if (x_lo <= x_hi) {
                                        const unsigned* starts = (flags & 2) ? a.qstart : a.tstart;
                                        const unsigned trow = P.cell_base + (unsigned)((nz * n1 + ny) * n0);
                                        start = __ldg(&starts[trow + x_lo]);
                                        len = __ldg(&starts[trow + x_hi + 1]) - start;
                                    }
                                }
                            }
                        }
                        unsigned todo = __ballot_sync(0xffffffffu, len != 0);
                        while (todo) {
                            const int src = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const unsigned st = __shfl_sync(0xffffffffu, start, src), ln = __shfl_sync(0xffffffffu, len, src);
                            const unsigned fl = __shfl_sync(0xffffffffu, flags, src);
                            const bool from_q = (fl & 2) != 0;   // entries of the symmetric walk that live in the query array
                            const bool cull = fl != 2;            // the own cell holds the queries themselves
                            auto gather = [&](unsigned off, unsigned n, unsigned at) -> unsigned {
                                if constexpr (sizeof(AT) == sizeof(NT)) {   // one array type: one copy of the loop
                                    const Vec4<AT>* arr = from_q ? qarr : reinterpret_cast<const Vec4<AT>*>(tarr);
                                    return copy_targets(arr + st + off, n, at, cull);
                                } else {
                                    return from_q ? copy_targets(qarr + st + off, n, at, cull)
                                                  : copy_targets(tarr + st + off, n, at, cull);
                                }
                            };
                            if (fill + ln <= kStageCap && small_ok) {  // the usual case: the row fits
                                const unsigned before = fill;
                                fill = gather(0u, ln, fill);
                                if (!(fl & 1)) lo_end = fill;
                                staged += nqc * (fill - before);   // at most kQueryCap * kStageCap
                                continue;
                            }
                            const unsigned long long npairs = (unsigned long long)nqc * ln;
                            if (npairs > direct_pairs) {  // giant cells: straight to the global histogram
                                oversize += npairs;
                                if (from_q)
                                    warp_pairs_general<AT, AT, ZMODE == 0 ? 0 : 1, true>(qarr, qc, qc + nqc, qarr + st, 0u, ln, thr,
                                                                                         nbins, dz2c, inv_w, c_magic,
                                                                                         (fl & 1) ? 2u : 1u, s_hist, g_hist);
                                else
                                    warp_pairs_general<AT, NT, ZMODE == 0 ? 0 : 1, true>(qarr, qc, qc + nqc, tarr + st, 0u, ln, thr,
                                                                                         nbins, dz2c, inv_w, c_magic,
                                                                                         (fl & 1) ? 2u : 1u, s_hist, g_hist);
                                continue;
                            }
                            unsigned off = 0;
                            while (off < ln) {
                                const unsigned n = min(ln - off, kStageCap - fill);
                                const unsigned before = fill;
                                fill = gather(off, n, fill);
                                staged += (unsigned long long)nqc * (fill - before);
                                off += n;
                                if (!(fl & 1)) lo_end = fill;
                                if (fill == kStageCap) flush_stage();
                            }
                        }
                    }
                    if (fill) flush_stage();
                }
                if (lane == 0) {
                    cand += staged + oversize;
                    if (staged) atomicAdd(&s_units, (unsigned)(staged >> 10) + 1u);
                }
            } else {
            // pass 0: real targets (SYM: upper half shell, own row to the right, own cell);
            // pass 1 (SYM only): shifted images, all rows, skipped where none can live
            const int npass = SYM ? 2 : 1;
            for (int pass = 0; pass < npass; ++pass) {
                const bool real = SYM && pass == 0;
                int r_begin = 0, r_end = nrows;
                if (real) r_begin = row0;
                if (SYM && pass == 1) {
                    const int kx0 = P.kd[0], ky = P.kd[1], kz = P.kd[2];
                    if (cx - kx0 >= P.gfree_lo[0] && cx + kx0 <= P.gfree_hi[0] && cy - ky >= P.gfree_lo[1] &&
                        cy + ky <= P.gfree_hi[1] && cz - kz >= P.gfree_lo[2] && cz + kz <= P.gfree_hi[2])
                        break;
                }
                const unsigned* starts = real ? a.qstart : a.tstart;
                for (int r = r_begin + part; r < r_end; r += rsplit) {
                    const int ny = cy + P.row_oy[r], nz = cz + P.row_oz[r];
                    if (ny < 0 || ny >= n1 || nz < 0 || nz >= n2) continue;
                    const int kx = P.row_kx[r];
                    int x_lo = max(cx - kx, 0);
                    const int x_hi = min(cx + kx, n0 - 1);
                    // in the query's own row only the own cell (both orders, weight 1) and the
                    // cells to its right (weight 2) are visited in the symmetric pass, as one range
                    const bool own_row = real && r == row0;
                    if (own_row) x_lo = cx;
                    const unsigned trow = P.cell_base + (unsigned)((nz * n1 + ny) * n0);
                    {
                        const unsigned ta = __ldg(&starts[trow + x_lo]), tb = __ldg(&starts[trow + x_hi + 1]);
                        if (ta == tb) continue;
                        const unsigned long long npairs = (unsigned long long)(qb - qa) * (tb - ta);
                        const bool direct = !SMEM || npairs > direct_pairs;
                        if (lane == 0) {
                            cand += npairs;
                            if (!direct) atomicAdd(&s_units, (unsigned)(npairs >> 10) + 1u);
                        }
                        if (real) {
                            // real targets are the cell-sorted queries themselves (NT == AT in this mode);
                            // the own cell is [qa, qb) of the same array
                            if (direct) {
                                if (own_row) {
                                    warp_pairs_general<AT, AT, ZMODE == 0 ? 0 : 1, true>(qsorted, qa, qb, qsorted, ta, qb, thr, nbins,
                                                                             dz2c, inv_w, c_magic, 1u, s_hist, g_hist);
                                    if (qb < tb)
                                        warp_pairs_general<AT, AT, ZMODE == 0 ? 0 : 1, true>(qsorted, qa, qb, qsorted, qb, tb, thr, nbins,
                                                                                 dz2c, inv_w, c_magic, 2u, s_hist, g_hist);
                                } else {
                                    warp_pairs_general<AT, AT, ZMODE == 0 ? 0 : 1, true>(qsorted, qa, qb, qsorted, ta, tb, thr, nbins,
                                                                             dz2c, inv_w, c_magic, 2u, s_hist, g_hist);
                                }
                            } else {
                                warp_pairs_fast<AT, AT, ZMODE>(qsorted, qa, qb, qsorted, ta, tb, dz2c, fa, hist_addr,
                                                               hist2_addr, own_row ? qb : 0u);
                            }
                        } else if (direct) {
                            warp_pairs_general<AT, NT, ZMODE == 0 ? 0 : 1, true>(qsorted, qa, qb, tsorted, ta, tb, thr, nbins, dz2c,
                                                                     inv_w, c_magic, 1u, s_hist, g_hist);
                        } else {
                            warp_pairs_fast<AT, NT, ZMODE>(qsorted, qa, qb, tsorted, ta, tb, dz2c, fa, hist_addr, hist_addr, 0u);
                        }
                    }
                }
            }
            }  // per-row walk (float64 regimes, tables outside shared memory)
        }
        __syncthreads();
    }
    if (lane == 0 && cand) atomicAdd(&a.counters[1], cand);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&a.counters[2], (unsigned long long)a.tstart[a.total_cells]);
}

// Small parameter blocks (plans, thresholds) are pulled from mapped pinned host memory by
// the SMs instead of a copy-engine transfer (see frame_stats).
__global__ void upload_kernel(int* __restrict__ dst, const int* __restrict__ src, long long nwords) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <typename T>
static int launch_stats(amepb_ctx* ctx, const T* pts, int64_t n, int nframes, StatsOut* d_partial,
                        StatsOut* d_out, int nparts, cudaStream_t stream) {
    dim3 grid(nparts, nframes);
    stats_partial_kernel<T><<<grid, 256, 0, stream>>>(pts, (long long)n, d_partial);
    AMEPB_LAUNCH_CHECK(ctx);
    stats_final_kernel<<<nframes, 32, 0, stream>>>(d_partial, nparts, d_out);
    AMEPB_LAUNCH_CHECK(ctx);
    return 0;
}

static int stats_parts(const amepb_ctx* ctx, int64_t n, int nframes) {
    int64_t per = (n + 256 * 8 - 1) / (256 * 8);
    int64_t want = std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / std::max(1, nframes));
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(per, want), 1024));
}

// stats of [nframes][n][3] device array -> host FrameStats[]
// Launch the statistics kernels of `nframes` frames (no synchronisation).  Results land in the
// context's pinned memory, slot 0: coords, slot 1: other.
static int frame_stats_launch(amepb_ctx* ctx, const void* pts, int dtype, int64_t n, int nframes,
                              cudaStream_t stream, int slot) {
    const int nparts = stats_parts(ctx, n, nframes);
    int rc = ensure_device(ctx, BUF_STATS, sizeof(StatsOut) * ((size_t)nframes * nparts * 2 + (size_t)nframes * 2));
    if (rc) return rc;
    rc = ensure_pinned(ctx, sizeof(StatsOut) * (size_t)nframes * 2);
    if (rc) return rc;
    // The per-frame result goes straight to pinned host memory (mapped into the device address
    // space under UVA): no copy-engine transfer, which would queue behind a large staging
    // copy of the host-buffer path and stall the whole pipeline.
    StatsOut* h = reinterpret_cast<StatsOut*>(ctx->pinned) + (size_t)slot * nframes;
    StatsOut* d_partial = reinterpret_cast<StatsOut*>(ctx->bufs[BUF_STATS].p) + (size_t)2 * nframes +
                          (size_t)slot * nframes * nparts;
    if (dtype == AMEPB_F32) return launch_stats<float>(ctx, (const float*)pts, n, nframes, d_partial, h, nparts, stream);
    return launch_stats<double>(ctx, (const double*)pts, n, nframes, d_partial, h, nparts, stream);
}

// Read the statistics back (the caller has waited for the kernels).
static void frame_stats_collect(amepb_ctx* ctx, int nframes, std::vector<FrameStats>& out, int slot) {
    const StatsOut* h = reinterpret_cast<const StatsOut*>(ctx->pinned) + (size_t)slot * nframes;
    out.resize(nframes);
    for (int f = 0; f < nframes; ++f) {
        for (int d = 0; d < 3; ++d) {
            out[f].lo[d] = h[f].lo[d];
            out[f].hi[d] = h[f].hi[d];
            out[f].first[d] = h[f].first[d];
            out[f].varies[d] = h[f].varies[d];
        }
    }
}

int frame_stats(amepb_ctx* ctx, const void* pts, int dtype, int64_t n, int nframes,
                std::vector<FrameStats>& out, cudaStream_t stream, int slot) {
    int rc = pinned_host_acquire(ctx);
    if (rc) return rc;
    rc = frame_stats_launch(ctx, pts, dtype, n, nframes, stream, slot);
    if (rc) return rc;
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    frame_stats_collect(ctx, nframes, out, slot);
    return 0;
}

int exclusive_scan(amepb_ctx* ctx, const unsigned* in, unsigned* out, long long len, cudaStream_t stream) {
    const int nblocks = (int)((len + kScanTile - 1) / kScanTile);
    int rc = ensure_device(ctx, BUF_SCAN, sizeof(unsigned) * (size_t)(nblocks + 1));
    if (rc) return rc;
    unsigned* sums = reinterpret_cast<unsigned*>(ctx->bufs[BUF_SCAN].p);
    scan_local_kernel<<<nblocks, kScanBlock, 0, stream>>>(in, out, len, sums);
    AMEPB_LAUNCH_CHECK(ctx);
    if (nblocks > 1) {
        scan_sums_kernel<<<1, 1024, 0, stream>>>(sums, nblocks);
        AMEPB_LAUNCH_CHECK(ctx);
        scan_add_kernel<<<nblocks, kScanBlock, 0, stream>>>(out, len, sums);
        AMEPB_LAUNCH_CHECK(ctx);
    }
    return 0;
}

template <typename AT, typename NT, bool SYM>
static int launch_pairs(amepb_ctx* ctx, const PairArgs& args, int zmode, cudaStream_t stream) {
    const int nbins = args.nbins;
    const size_t smem = pair_thr_bytes(nbins, sizeof(AT)) + (size_t)(nbins + 2) * sizeof(unsigned) * (SYM ? 2 : 1);
    // the one-comparison bin correction needs uniform edges (args.uniform) and nbins <= 2^17;
    // the tables must fit in shared memory
    const bool use_smem = smem <= 40 * 1024 && args.uniform && nbins <= (1 << 17);
    void (*kern)(PairArgs) = nullptr;
    if (use_smem)
        kern = zmode == 2 ? pair_kernel<AT, NT, 2, true, SYM>
                          : (zmode == 1 ? pair_kernel<AT, NT, 1, true, SYM> : pair_kernel<AT, NT, 0, true, SYM>);
    else kern = zmode != 0 ? pair_kernel<AT, NT, 1, false, SYM> : pair_kernel<AT, NT, 0, false, SYM>;
    size_t dyn = use_smem ? smem : 16;
    if (use_smem) {
        // per-warp staging buffers (see kStageCapF32; the float64 kernels stage half as many elements)
        dyn = ((dyn + 15) & ~(size_t)15) +
              (size_t)(kPairThreads / 32) * (kStageCapF32 + kQueryCapF32) * (zmode == 0 ? 3 : 2) * sizeof(float);
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    }
    int per_sm = 0;
    AMEPB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPairThreads, dyn));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)ctx->sm_count * per_sm;
    blocks = std::max<long long>(1, std::min<long long>(blocks, (args.total_tasks + args.shard_count - 1) / args.shard_count));
    kern<<<(unsigned)blocks, kPairThreads, dyn, stream>>>(args);
    AMEPB_LAUNCH_CHECK(ctx);
    return 0;
}

struct RdfCall {
    const void* coords;
    int coords_dtype;
    int64_t n;
    const void* other;  // may be NULL
    int other_dtype;
    int64_t m;
    int space;
    int64_t nframes;
    const double* box;
    int box_dtype;
    const double* edges;
    int64_t edges_stride;
    int nbins;
    int pbc;
    uint64_t* hist_out;
    int32_t* dim_out;
    cudaStream_t stream;
};

static size_t elem_size(int dtype) { return dtype == AMEPB_F32 ? 4 : 8; }

// The sub-batch after the current one: its statistics kernels are launched ahead of the current
// pair kernel, so that the host has them (and can plan and enqueue the next build) while the GPU
// is still busy -- the stream never runs dry waiting for the host.
struct NextBatch {
    int64_t f0;
    int nf;
    const void* d_coords;
    const void* d_other;
    cudaEvent_t ready;   // host-buffer path: the staging copy; prefetch only if it has completed
};

// One sub-batch: frames [f0, f0+nf) of the call; d_coords / d_other / d_hist
// are device pointers to this sub-batch's data.
static int rdf_sub_batch(amepb_ctx* ctx, const RdfCall& c, int64_t f0, int nf, const void* d_coords,
                         const void* d_other, unsigned long long* d_hist, bool* bad_dim,
                         const NextBatch* next = nullptr) {
    cudaStream_t stream = c.stream;
    const bool self = (c.other == nullptr);
    const int qdtype = self ? c.coords_dtype : c.other_dtype;
    const int64_t m = self ? c.n : c.m;
    const void* d_query = self ? d_coords : d_other;
    // arithmetic regime (amep/spatialcor.py:381): images are float32 with pbc,
    // otherwise the targets keep their dtype; the query keeps its own dtype.
    const bool targets_f64 = (!c.pbc && c.coords_dtype == AMEPB_F64);
    const bool arith_f64 = targets_f64 || qdtype == AMEPB_F64;
    const int nbins = c.nbins;

    AMEPB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (size_t)nf * nbins, stream));

    // ---- stats -> plans ------------------------------------------------------
    static const bool trace_host = getenv("AMEPB_TRACE") != nullptr;
    auto host_ms = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double th0 = trace_host ? host_ms() : 0.0;
    cudaEvent_t ev_b0 = timing_event(ctx, stream);
    std::vector<FrameStats> tstats, qstats;
    int rc = 0;
    if (ctx->stats_ahead_f0 == f0 && ctx->stats_ahead_nf == nf && ctx->ev_stats) {
        // launched ahead of the previous pair kernel
        AMEPB_CUDA(ctx, cudaEventSynchronize(ctx->ev_stats));
    } else {
        // the statistics land in the pinned block: no read queued by an earlier call may be pending
        if ((rc = pinned_host_acquire(ctx))) return rc;
        if ((rc = frame_stats_launch(ctx, d_coords, c.coords_dtype, c.n, nf, stream, 0))) return rc;
        if (!self && (rc = frame_stats_launch(ctx, d_other, c.other_dtype, c.m, nf, stream, 1))) return rc;
        AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    ctx->stats_ahead_f0 = -1;
    frame_stats_collect(ctx, nf, tstats, 0);
    if (!self) frame_stats_collect(ctx, nf, qstats, 1);
    const double th1 = trace_host ? host_ms() : 0.0;
    std::vector<FramePlan> plans(nf);
    long long total_cells = 0, total_tasks = 0, total_cap = 0;
    bool all_zconst = true;
    int k_used = 0;
    for (int f = 0; f < nf; ++f) {
        const FrameStats& ts = tstats[f];
        const FrameStats& qs = self ? tstats[f] : qstats[f];
        const int dim = dimension_from_stats(ts, c.n);
        if (c.dim_out) c.dim_out[f0 + f] = dim;
        FramePlan& P = plans[f];
        const double* edges = c.edges + (c.edges_stride ? (f0 + f) * c.edges_stride : 0);
        bool ok = true;
        if (c.pbc && dim != 2 && dim != 3) {
            *bad_dim = true;
            ok = false;
        }
        PlanInputs in;
        in.box_boundary = c.box + (f0 + f) * 6;
        in.box_is_f32 = (c.box_dtype == AMEPB_F32);
        in.pbc = c.pbc;
        in.rc = edges[nbins];
        in.n_targets = c.n;
        in.n_queries = m;
        in.cells_per_cutoff = ctx->rdf_k_override;
        in.max_cells = std::max<int64_t>(4096, c.n * (c.pbc ? 2 : 1));
        if (const char* kx = getenv("AMEPB_KX")) in.cells_per_cutoff_dim[0] = atoi(kx);   // experiments: anisotropic cells
        if (const char* ky = getenv("AMEPB_KY")) in.cells_per_cutoff_dim[1] = atoi(ky);
        if (const char* kz = getenv("AMEPB_KZ")) in.cells_per_cutoff_dim[2] = atoi(kz);
        if (const char* tp = getenv("AMEPB_TASK_PAIRS")) in.task_pairs = std::max(1024.0, atof(tp));
        if (const char* hm = getenv("AMEPB_HEAVY_RUN_MULT")) in.heavy_run_mult = std::max(0, atoi(hm));
        // edges must be usable: finite, non-decreasing
        for (int i = 0; ok && i < nbins; ++i)
            if (!(edges[i] <= edges[i + 1])) ok = false;
        if (ok) ok = plan_frame(in, dim, ts, qs, P);
        else memset(&P, 0, sizeof(P)), P.dim = dim;
        P.thr_index = c.edges_stride ? f : 0;
        P.cell_base = (unsigned)total_cells;
        P.task_begin = total_tasks;
        if (P.valid) {
            // (with z images -- dimension 3, e.g. a single target particle -- dz is not one constant)
            const bool zc = !ts.varies[2] && !qs.varies[2] && !(c.pbc && P.nshift[2] == 3);
            all_zconst = all_zconst && zc;
            if (zc) {
                if (arith_f64) {
                    const double qz = qs.first[2];
                    const double nz = targets_f64 ? ts.first[2] : (double)(float)ts.first[2];
                    const double dz = qz - nz;
                    P.dz2c = dz * dz;
                } else {
                    const float qz = (float)qs.first[2];
                    const float nz = (float)ts.first[2];
                    const volatile float dz = qz - nz;
                    const volatile float dz2 = dz * dz;
                    P.dz2c = (double)dz2;
                }
            }
            P.bin_e0 = edges[0];
            const double width = edges[nbins] - edges[0];
            P.bin_inv_w = width > 0 ? (double)nbins / width : 0.0;
            total_cells += P.ncells;
            total_tasks += P.ntasks;
            long long cap = c.n;
            if (c.pbc)
                for (int d = 0; d < 3; ++d) cap *= (P.nshift[d] == 3 ? 2 : 1);
            total_cap += cap;
            for (int r = 0; r < P.nrows; ++r) k_used = std::max<int>(k_used, std::max<int>(P.row_kx[r], std::max(abs((int)P.row_oy[r]), abs((int)P.row_oz[r]))));
        } else {
            P.ntasks = 0;
            P.ncells = 0;
        }
    }
    // symmetric mode: self RDF whose stored queries equal the unshifted targets bit for bit
    // (float32 regime, or no images at all) and whose particles all pass the image window
    bool sym = ctx->rdf_symmetric && self && (c.coords_dtype == AMEPB_F32 || !c.pbc);
    for (int f = 0; sym && f < nf; ++f) {
        const FramePlan& P = plans[f];
        if (!P.valid || !c.pbc) continue;
        for (int d = 0; d < 3; ++d) {
            const double lo = (double)(float)tstats[f].lo[d], hi = (double)(float)tstats[f].hi[d];
            if (!(lo > P.win_lo[d]) || !(hi < P.win_hi[d])) sym = false;
        }
    }
    for (int f = 0; f < nf; ++f) plans[f].sym = sym ? 1 : 0;
    ctx->rdf_info.frames += nf;
    ctx->rdf_info.cells += total_cells;
    ctx->rdf_info.block_tasks += total_tasks;
    ctx->rdf_info.sub_batches += 1;
    if (total_tasks == 0) return 0;
    if (total_cells >= (1ll << 31) || total_cap >= (1ll << 32) || (long long)m * nf >= (1ll << 32))
        return fail(ctx, AMEPB_E_NOMEM, "amepb_rdf_batch: sub-batch too large for 32-bit cell offsets");

    const double th2 = trace_host ? host_ms() : 0.0;
    // ---- thresholds ------------------------------------------------------------
    const int ntables = c.edges_stride ? nf : 1;
    const size_t at_size = arith_f64 ? 8 : 4;
    const size_t thr_bytes = (size_t)ntables * (nbins + 1) * at_size;
    const size_t plan_bytes = sizeof(FramePlan) * (size_t)nf;
    rc = ensure_pinned(ctx, plan_bytes + thr_bytes + 64);
    if (rc) return rc;
    rc = ensure_device(ctx, BUF_PLANS, plan_bytes);
    if (rc) return rc;
    rc = ensure_device(ctx, BUF_THR, thr_bytes);
    if (rc) return rc;
    // (the statistics of this sub-batch were read above, after every earlier read of the block on
    // this stream had run; a read queued by another call on another stream is waited for here)
    if ((rc = pinned_host_acquire(ctx))) return rc;
    char* hp = reinterpret_cast<char*>(ctx->pinned);
    char* hthr = hp + ((plan_bytes + 15) & ~(size_t)15);
    bool uniform = true;
    for (int tb = 0; tb < ntables; ++tb) {
        const double* edges = c.edges + (c.edges_stride ? (f0 + tb) * c.edges_stride : 0);
        const double w = (edges[nbins] - edges[0]) / nbins;
        for (int i = 0; i <= nbins && uniform; ++i)
            if (!(fabs(edges[i] - (edges[0] + i * w)) <= 0.01 * w)) uniform = false;
        // pairs below the first counted distance (d <= 1e-3 or d < edges[0]) must round to guess 0
        if (!(w > 0) || !((std::max(1e-3, edges[0]) - edges[0]) / w < 0.45)) uniform = false;
        // the exact thresholds cost ~0.1 ms of host time per table: reuse the last table when the
        // edges are the same (every sub-batch of a call, every frame with the same box)
        char* dst = hthr + (size_t)tb * (nbins + 1) * at_size;
        const size_t tbytes = (size_t)(nbins + 1) * at_size;
        const bool hit = ctx->thr_cache_f64 == (arith_f64 ? 1 : 0) && ctx->thr_cache_edges.size() == (size_t)nbins + 1 &&
                         memcmp(ctx->thr_cache_edges.data(), edges, sizeof(double) * (nbins + 1)) == 0;
        if (hit) {
            memcpy(dst, ctx->thr_cache_values.data(), tbytes);
        } else {
            if (arith_f64) make_thresholds<double>(edges, nbins, reinterpret_cast<double*>(dst));
            else make_thresholds<float>(edges, nbins, reinterpret_cast<float*>(dst));
            ctx->thr_cache_edges.assign(edges, edges + nbins + 1);
            ctx->thr_cache_values.assign(dst, dst + tbytes);
            ctx->thr_cache_f64 = arith_f64 ? 1 : 0;
        }
    }
    // z mode of the pair kernel (all frames flat: dz*dz is a per-frame constant, or zero)
    int zmode = all_zconst ? 2 : 0;
    for (int f = 0; f < nf && zmode == 2; ++f)
        if (plans[f].valid && plans[f].dz2c != 0.0) zmode = 1;
    // compact (x, y) records: flat frames in the float32 regime on the staged walk (the same test as launch_pairs)
    {
        const size_t tab = pair_thr_bytes(nbins, sizeof(float)) + (size_t)(nbins + 2) * sizeof(unsigned) * (sym ? 2 : 1);
        const bool no_compact = getenv("AMEPB_NO_COMPACT_RECORDS") != nullptr;   // A/B switch for measurements and tests
        const bool compact = !arith_f64 && !targets_f64 && zmode != 0 && tab <= 40 * 1024 && uniform && nbins <= (1 << 17) && !no_compact;
        for (int f = 0; f < nf; ++f) plans[f].compact = compact ? 1 : 0;
    }
    // partitioned build of the query array (strip_* kernels): self RDF with images on the symmetric walk,
    // enough (strip, frame) blocks to be worth it, the cell counters of a strip in shared memory
    long long total_strips = 0;
    int max_strips = 0;
    size_t sort_smem = 0;
    bool partitioned = false;
    {
        const bool no_part = getenv("AMEPB_NO_PARTITIONED_BUILD") != nullptr;   // A/B switch for measurements and tests
        bool ok = self && c.pbc && sym && !no_part;
        unsigned qbase = 0;
        for (int f = 0; f < nf; ++f) {
            FramePlan& P = plans[f];
            P.strip_rows = 0;
            P.nstrips = 0;
            P.qbase = qbase;
            if (!P.valid) continue;
            qbase += (unsigned)m;
            const long long rows = (long long)P.n[1] * P.n[2];
            const long long rps = std::max<long long>(1, (rows + kMaxStrips - 1) / kMaxStrips);
            const long long ns = (rows + rps - 1) / rps;
            const size_t smem = sizeof(unsigned) * (size_t)rps * (size_t)P.n[0];
            if (smem > 160 * 1024 || rows <= 0) ok = false;
            if (rows * rps >= (1ll << 31)) ok = false;
            P.strip_rows = (int)rps;
            P.strip_magic = rps == 1 ? 0u : (unsigned)((1ull << 32) / (unsigned long long)rps + 1ull);
            P.nstrips = (int)ns;
            total_strips += ns;
            max_strips = std::max<int>(max_strips, (int)ns);
            sort_smem = std::max(sort_smem, smem);
        }
        const bool force_part = getenv("AMEPB_FORCE_PARTITIONED_BUILD") != nullptr;   // tests: also for tiny batches
        // (measured down to one 1M-particle frame, 254 strips: never slower than the placement kernels)
        partitioned = ok && (total_strips >= 128 || force_part);
        if (!partitioned)
            for (int f = 0; f < nf; ++f) plans[f].strip_rows = 0;
    }
    memcpy(hp, plans.data(), plan_bytes);
    {
        const long long pw = (long long)(plan_bytes / sizeof(int)), tw = (long long)((thr_bytes + 3) / sizeof(int));
        upload_kernel<<<(unsigned)std::min<long long>(64, (pw + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<int*>(ctx->bufs[BUF_PLANS].p), reinterpret_cast<const int*>(hp), pw);
        AMEPB_LAUNCH_CHECK(ctx);
        upload_kernel<<<(unsigned)std::min<long long>(64, (tw + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<int*>(ctx->bufs[BUF_THR].p), reinterpret_cast<const int*>(hthr), tw);
        AMEPB_LAUNCH_CHECK(ctx);
        if ((rc = pinned_device_release(ctx, stream))) return rc;
    }
    const FramePlan* d_plans = reinterpret_cast<const FramePlan*>(ctx->bufs[BUF_PLANS].p);

    // ---- workspaces ------------------------------------------------------------
    const size_t cell_bytes = sizeof(unsigned) * (size_t)(total_cells + 1);
    const size_t tsz = targets_f64 ? sizeof(Vec4<double>) : sizeof(Vec4<float>);
    const size_t qsz = arith_f64 ? sizeof(Vec4<double>) : sizeof(Vec4<float>);
    if ((rc = ensure_device(ctx, BUF_TCOUNT, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_TSTART, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_QCOUNT, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_QSTART, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_TSORTED, tsz * (size_t)total_cap))) return rc;
    if ((rc = ensure_device(ctx, BUF_QSORTED, qsz * (size_t)m * nf))) return rc;
    if ((rc = ensure_device(ctx, BUF_MISC, 64))) return rc;
    unsigned* tcount = reinterpret_cast<unsigned*>(ctx->bufs[BUF_TCOUNT].p);
    unsigned* tstart = reinterpret_cast<unsigned*>(ctx->bufs[BUF_TSTART].p);
    unsigned* qcount = reinterpret_cast<unsigned*>(ctx->bufs[BUF_QCOUNT].p);
    unsigned* qstart = reinterpret_cast<unsigned*>(ctx->bufs[BUF_QSTART].p);
    AMEPB_CUDA(ctx, cudaMemsetAsync(tcount, 0, cell_bytes, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(qcount, 0, cell_bytes, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(ctx->bufs[BUF_MISC].p, 0, sizeof(unsigned long long), stream));

    // ---- cell build ------------------------------------------------------------
    constexpr int kPlaceTile = 256 * kPlaceILP;
    const dim3 tgrid((unsigned)((c.n + kPlaceTile - 1) / kPlaceTile), nf), qgrid((unsigned)((m + kPlaceTile - 1) / kPlaceTile), nf);
    auto place_targets = [&](bool fill) -> int {
        void* sorted = ctx->bufs[BUF_TSORTED].p;
        if (c.pbc) {
            if (c.coords_dtype == AMEPB_F32) {
                if (fill) place_images_kernel<float, true><<<tgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
                else place_images_kernel<float, false><<<tgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
            } else {
                if (fill) place_images_kernel<double, true><<<tgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
                else place_images_kernel<double, false><<<tgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
            }
        } else if (c.coords_dtype == AMEPB_F32) {
            if (fill) place_points_kernel<float, float, true, true><<<tgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
            else place_points_kernel<float, float, false, true><<<tgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<float>*)sorted);
        } else {
            if (fill) place_points_kernel<double, double, true, true><<<tgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<double>*)sorted);
            else place_points_kernel<double, double, false, true><<<tgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, (Vec4<double>*)sorted);
        }
        AMEPB_LAUNCH_CHECK(ctx);
        return 0;
    };
    auto place_queries = [&](bool fill) -> int {
        void* sorted = ctx->bufs[BUF_QSORTED].p;
        if (qdtype == AMEPB_F32 && !arith_f64) {
            if (fill) place_points_kernel<float, float, true, false><<<qgrid, 256, 0, stream>>>((const float*)d_query, m, d_plans, qcount, qstart, (Vec4<float>*)sorted);
            else place_points_kernel<float, float, false, false><<<qgrid, 256, 0, stream>>>((const float*)d_query, m, d_plans, qcount, qstart, (Vec4<float>*)sorted);
        } else if (qdtype == AMEPB_F32) {
            if (fill) place_points_kernel<float, double, true, false><<<qgrid, 256, 0, stream>>>((const float*)d_query, m, d_plans, qcount, qstart, (Vec4<double>*)sorted);
            else place_points_kernel<float, double, false, false><<<qgrid, 256, 0, stream>>>((const float*)d_query, m, d_plans, qcount, qstart, (Vec4<double>*)sorted);
        } else {
            if (fill) place_points_kernel<double, double, true, false><<<qgrid, 256, 0, stream>>>((const double*)d_query, m, d_plans, qcount, qstart, (Vec4<double>*)sorted);
            else place_points_kernel<double, double, false, false><<<qgrid, 256, 0, stream>>>((const double*)d_query, m, d_plans, qcount, qstart, (Vec4<double>*)sorted);
        }
        AMEPB_LAUNCH_CHECK(ctx);
        return 0;
    };
    const bool have_targets = !(sym && !c.pbc);  // symmetric mode without images: the queries are the targets
    const bool fused = self && c.pbc;              // one pass over the coordinates for queries and images
    unsigned* qrank = nullptr;
    if (fused) {
        if ((rc = ensure_device(ctx, BUF_QRANK, sizeof(unsigned) * (size_t)m * nf))) return rc;
        qrank = reinterpret_cast<unsigned*>(ctx->bufs[BUF_QRANK].p);
    }
    // count pass: frames as the fast block index (see place_self_kernel) when the tile count fits gridDim.y
    static const bool no_frame_fast = getenv("AMEPB_NO_FRAME_FAST") != nullptr;   // A/B switch for measurements
    const int frame_fast = (nf > 1 && tgrid.x <= 65535u && !no_frame_fast) ? 1 : 0;
    const dim3 cgrid = frame_fast ? dim3((unsigned)nf, tgrid.x) : tgrid;
    auto place_self = [&](bool fill) -> int {
        Vec4<float>* ts = reinterpret_cast<Vec4<float>*>(ctx->bufs[BUF_TSORTED].p);
        void* qs = ctx->bufs[BUF_QSORTED].p;
        if (c.coords_dtype == AMEPB_F32) {
            if (fill) place_self_kernel<float, float, true><<<tgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, ts, qcount, qstart, (Vec4<float>*)qs, qrank, 0);
            else place_self_kernel<float, float, false><<<cgrid, 256, 0, stream>>>((const float*)d_coords, c.n, d_plans, tcount, tstart, ts, qcount, qstart, (Vec4<float>*)qs, qrank, frame_fast);
        } else {
            if (fill) place_self_kernel<double, double, true><<<tgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, ts, qcount, qstart, (Vec4<double>*)qs, qrank, 0);
            else place_self_kernel<double, double, false><<<cgrid, 256, 0, stream>>>((const double*)d_coords, c.n, d_plans, tcount, tstart, ts, qcount, qstart, (Vec4<double>*)qs, qrank, frame_fast);
        }
        AMEPB_LAUNCH_CHECK(ctx);
        return 0;
    };
    // partitioned build: records in strip order, strip counters
    const bool xy_records = !arith_f64 && plans[0].compact != 0;
    bool flat_cells = true;   // every valid frame has one layer of cells
    for (int f = 0; f < nf; ++f)
        if (plans[f].valid && plans[f].n[2] != 1) flat_cells = false;
    // ((x, y) records: z is constant in every frame; a frame that still shifts images along z -- a single
    // particle counts as three-dimensional -- keeps the old build)
    for (int f = 0; f < nf; ++f)
        if (partitioned && xy_records && plans[f].valid && (plans[f].nshift[2] == 3 || plans[f].n[2] != 1)) partitioned = false;
    unsigned *strip_count = nullptr, *strip_cursor = nullptr, *strip_start = nullptr;
    const dim3 sgrid((unsigned)((c.n + kStripTile - 1) / kStripTile), nf);
    if (partitioned) {
        const size_t rec = xy_records ? sizeof(float2) : (arith_f64 ? sizeof(Vec4<double>) : sizeof(Vec4<float>));
        if ((rc = ensure_device(ctx, BUF_MID, rec * (size_t)m * nf))) return rc;
        const size_t per_frame = (size_t)kMaxStrips;   // fixed stride per frame, see strip_scan_kernel
        if ((rc = ensure_device(ctx, BUF_STRIPS, sizeof(unsigned) * ((size_t)nf * (3 * per_frame + 1) + 8)))) return rc;
        strip_count = reinterpret_cast<unsigned*>(ctx->bufs[BUF_STRIPS].p);
        strip_cursor = strip_count + (size_t)nf * per_frame;
        strip_start = strip_cursor + (size_t)nf * per_frame;
        AMEPB_CUDA(ctx, cudaMemsetAsync(strip_count, 0, sizeof(unsigned) * (size_t)nf * per_frame, stream));
    }
    auto strip_tiles = [&](bool scatter) -> int {
        Vec4<float>* ts = reinterpret_cast<Vec4<float>*>(ctx->bufs[BUF_TSORTED].p);
        void* mid = ctx->bufs[BUF_MID].p;
#define AMEPB_STRIP_TILE2(T, AT, XY, FL)                                                                                         \
    do {                                                                                                                         \
        if (scatter) strip_tile_kernel<T, AT, XY, true, FL><<<sgrid, kStripThreads, 0, stream>>>((const T*)d_coords, c.n, d_plans, strip_count, strip_cursor, (typename StripRec<AT, XY>::type*)mid, tcount, tstart, ts); \
        else strip_tile_kernel<T, AT, XY, false, FL><<<sgrid, kStripThreads, 0, stream>>>((const T*)d_coords, c.n, d_plans, strip_count, strip_cursor, (typename StripRec<AT, XY>::type*)mid, tcount, tstart, ts); \
    } while (0)
#define AMEPB_STRIP_TILE(T, AT, XY)                                                                                              \
    do {                                                                                                                         \
        if (flat_cells) AMEPB_STRIP_TILE2(T, AT, XY, true);                                                                      \
        else AMEPB_STRIP_TILE2(T, AT, XY, false);                                                                                \
    } while (0)
        if (c.coords_dtype == AMEPB_F32 && xy_records) AMEPB_STRIP_TILE(float, float, true);
        else if (c.coords_dtype == AMEPB_F32) AMEPB_STRIP_TILE(float, float, false);
        else AMEPB_STRIP_TILE(double, double, false);
#undef AMEPB_STRIP_TILE2
#undef AMEPB_STRIP_TILE
        AMEPB_LAUNCH_CHECK(ctx);
        return 0;
    };
    auto strip_sort = [&]() -> int {
        const dim3 grid((unsigned)max_strips, nf);
        void* mid = ctx->bufs[BUF_MID].p;
        void* qs = ctx->bufs[BUF_QSORTED].p;
#define AMEPB_STRIP_SORT2(AT, XY, FL)                                                                                            \
    do {                                                                                                                         \
        if (sort_smem > 48 * 1024)                                                                                               \
            AMEPB_CUDA(ctx, cudaFuncSetAttribute(strip_sort_kernel<AT, XY, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem)); \
        strip_sort_kernel<AT, XY, FL><<<grid, kSortThreads, sort_smem, stream>>>((const typename StripRec<AT, XY>::type*)mid, d_plans, strip_start, c.n, qstart, (typename StripRec<AT, XY>::type*)qs); \
    } while (0)
#define AMEPB_STRIP_SORT(AT, XY)                                                                                                 \
    do {                                                                                                                         \
        if (flat_cells) AMEPB_STRIP_SORT2(AT, XY, true);                                                                         \
        else AMEPB_STRIP_SORT2(AT, XY, false);                                                                                   \
    } while (0)
        if (c.coords_dtype == AMEPB_F32 && xy_records) AMEPB_STRIP_SORT(float, true);
        else if (c.coords_dtype == AMEPB_F32) AMEPB_STRIP_SORT(float, false);
        else AMEPB_STRIP_SORT(double, false);
#undef AMEPB_STRIP_SORT2
#undef AMEPB_STRIP_SORT
        AMEPB_LAUNCH_CHECK(ctx);
        return 0;
    };
    if (partitioned) {
        if ((rc = strip_tiles(false))) return rc;
    } else if (fused) {
        if ((rc = place_self(false))) return rc;
    } else {
        if (have_targets && (rc = place_targets(false))) return rc;
        if ((rc = place_queries(false))) return rc;
    }
    if ((rc = exclusive_scan(ctx, tcount, tstart, total_cells + 1, stream))) return rc;
    if (partitioned) {
        strip_scan_kernel<<<nf, 32, 0, stream>>>(d_plans, strip_count, strip_cursor, strip_start);
        AMEPB_LAUNCH_CHECK(ctx);
        if ((rc = strip_tiles(true))) return rc;
        if ((rc = strip_sort())) return rc;
    } else if ((rc = exclusive_scan(ctx, qcount, qstart, total_cells + 1, stream))) {
        return rc;
    }
    if (partitioned) {
        // (queries and shifted images are in place)
    } else if (fused) {
        if ((rc = place_self(true))) return rc;
    } else {
        if (have_targets && (rc = place_targets(true))) return rc;
        if ((rc = place_queries(true))) return rc;
    }
    // ---- statistics of the next sub-batch, ahead of this sub-batch's pair kernel -------------
    // Device-resident frames: on the caller's stream, in front of the pair kernel.  Staged frames
    // (host-buffer path): on a stream of their own that waits for the staging copy of the next
    // sub-batch -- enqueued on the caller's stream they would hold this sub-batch's pair kernel
    // back until that copy has landed, and launched only once the copy is known to be complete
    // (the round-1 rule) they were usually skipped, because copies and computation take about equally
    // long: every sub-batch then started with a statistics launch, a stream synchronisation and the
    // host-side planning while the GPU sat idle (about 0.4 ms per sub-batch of 32 frames).
    if (next && next->nf > 0) {
        if (!ctx->ev_stats) AMEPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stats, cudaEventDisableTiming));
        cudaStream_t ss = stream;
        if (next->ready) {
            if (!ctx->stats_stream) AMEPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stats_stream, cudaStreamNonBlocking));
            ss = ctx->stats_stream;
            AMEPB_CUDA(ctx, cudaStreamWaitEvent(ss, next->ready, 0));
            // the statistics land in the pinned block this sub-batch's plans were pulled from
            if (ctx->ev_pinned) AMEPB_CUDA(ctx, cudaStreamWaitEvent(ss, ctx->ev_pinned, 0));
        }
        if ((rc = frame_stats_launch(ctx, next->d_coords, c.coords_dtype, c.n, next->nf, ss, 0))) return rc;
        if (!self && (rc = frame_stats_launch(ctx, next->d_other, c.other_dtype, c.m, next->nf, ss, 1))) return rc;
        AMEPB_CUDA(ctx, cudaEventRecord(ctx->ev_stats, ss));
        ctx->stats_ahead_f0 = next->f0;
        ctx->stats_ahead_nf = next->nf;
    }
    cudaEvent_t ev_b1 = timing_event(ctx, stream);
    if (ev_b0 && ev_b1) ctx->ev_build.emplace_back(ev_b0, ev_b1);

    // ---- pairs -----------------------------------------------------------------
    PairArgs args;
    args.plans = d_plans;
    args.tstart = tstart;
    args.qstart = qstart;
    args.tsorted = ctx->bufs[BUF_TSORTED].p;
    args.qsorted = ctx->bufs[BUF_QSORTED].p;
    args.thr = ctx->bufs[BUF_THR].p;
    args.hist = d_hist;
    args.counters = reinterpret_cast<unsigned long long*>(ctx->bufs[BUF_MISC].p);
    args.total_tasks = total_tasks;
    args.total_cells = total_cells;
    args.nbins = nbins;
    args.nframes = nf;
    args.uniform = uniform ? 1 : 0;
    args.one = 1.0f;
    args.shard_index = ctx->rdf_shard_index;
    args.shard_count = ctx->rdf_shard_count;
    if (sym) {
        if (!arith_f64) rc = launch_pairs<float, float, true>(ctx, args, zmode, stream);
        else rc = launch_pairs<double, double, true>(ctx, args, zmode, stream);
    } else if (!arith_f64) rc = launch_pairs<float, float, false>(ctx, args, zmode, stream);
    else if (!targets_f64) rc = launch_pairs<double, float, false>(ctx, args, zmode, stream);
    else rc = launch_pairs<double, double, false>(ctx, args, zmode, stream);
    if (rc) return rc;
    cudaEvent_t ev_p1 = timing_event(ctx, stream);
    if (ev_b1 && ev_p1) ctx->ev_pair.emplace_back(ev_b1, ev_p1);

    ctx->rdf_info.cells_per_cutoff = k_used;
    if (trace_host)
        fprintf(stderr, "[amepb] host: %d frames: stats+sync %.3f ms, plan %.3f ms, thresholds+launches %.3f ms\n", nf,
                th1 - th0, th2 - th1, host_ms() - th2);
    return 0;
}

}  // namespace amepb

using namespace amepb;

extern "C" {

int amepb_rdf_set_symmetric(amepb_ctx* ctx, int enable) {
    if (!ctx) return AMEPB_E_INVAL;
    ctx->rdf_symmetric = enable != 0;
    return 0;
}

int amepb_rdf_set_shard(amepb_ctx* ctx, int index, int count) {
    if (!ctx || count < 1 || index < 0 || index >= count) return AMEPB_E_INVAL;
    ctx->rdf_shard_index = index;
    ctx->rdf_shard_count = count;
    return 0;
}

int amepb_rdf_set_cells_per_cutoff(amepb_ctx* ctx, int k) {
    if (!ctx || k < 0 || k > kMaxCellsPerCutoff) return AMEPB_E_INVAL;
    ctx->rdf_k_override = k;
    return 0;
}

int amepb_rdf_last_info(amepb_ctx* ctx, amepb_rdf_info* info) {
    if (!ctx || !info) return AMEPB_E_INVAL;
    DeviceGuard guard(ctx->device);
    if (ctx->bufs[BUF_MISC].p && ctx->rdf_info.block_tasks > 0) {
        unsigned long long h[3] = {0, 0, 0};
        AMEPB_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)ctx->last_stream));
        AMEPB_CUDA(ctx, cudaMemcpy(h, ctx->bufs[BUF_MISC].p, sizeof(h), cudaMemcpyDeviceToHost));
        ctx->rdf_info.candidate_pairs = (int64_t)h[1];
        ctx->rdf_info.images = (int64_t)h[2];
    }
    ctx->rdf_info.pair_kernel_ms = timing_total_ms(ctx, ctx->ev_pair);
    ctx->rdf_info.build_ms = timing_total_ms(ctx, ctx->ev_build);
    *info = ctx->rdf_info;
    return 0;
}

int amepb_frame_dimension(amepb_ctx* ctx, const void* coords, int dtype, int space, int64_t nframes,
                          int64_t n, int32_t* dim_out, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !dim_out || nframes < 0 || n <= 0 || (dtype != AMEPB_F32 && dtype != AMEPB_F64) ||
        (space != AMEPB_DEVICE && space != AMEPB_HOST))
        return fail(ctx, AMEPB_E_INVAL, "amepb_frame_dimension: invalid argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t frame_bytes = (size_t)n * 3 * elem_size(dtype);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nframes, std::min<int64_t>(32768, (int64_t)((2ull << 30) / frame_bytes) + 1)));
    std::vector<FrameStats> st;
    for (int64_t f0 = 0; f0 < nframes; f0 += chunk) {
        const int nf = (int)std::min<int64_t>(chunk, nframes - f0);
        const void* d = (const char*)coords + (size_t)f0 * frame_bytes;
        if (space == AMEPB_HOST) {
            int rc = ensure_device(ctx, BUF_STAGE_A, frame_bytes * nf);
            if (rc) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, d, frame_bytes * nf, cudaMemcpyHostToDevice, stream));
            d = ctx->bufs[BUF_STAGE_A].p;
        }
        int rc = frame_stats(ctx, d, dtype, n, nf, st, stream, 0);
        if (rc) return rc;
        for (int f = 0; f < nf; ++f) dim_out[f0 + f] = dimension_from_stats(st[f], n);
    }
    return 0;
}

static int rdf_batch_impl(amepb_ctx* ctx, const void* coords, int coords_dtype, int64_t n, const void* other,
                          int other_dtype, int64_t m, int space, int64_t nframes, const double* box_boundary,
                          int box_dtype, const double* edges, int64_t edges_stride, int32_t nbins, int pbc,
                          uint64_t* hist_out, int32_t* dim_out, void* stream_, bool* budget_was_cached);

int amepb_rdf_batch(amepb_ctx* ctx, const void* coords, int coords_dtype, int64_t n, const void* other,
                    int other_dtype, int64_t m, int space, int64_t nframes, const double* box_boundary,
                    int box_dtype, const double* edges, int64_t edges_stride, int32_t nbins, int pbc,
                    uint64_t* hist_out, int32_t* dim_out, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    bool cached = false;
    int rc = rdf_batch_impl(ctx, coords, coords_dtype, n, other, other_dtype, m, space, nframes, box_boundary, box_dtype,
                            edges, edges_stride, nbins, pbc, hist_out, dim_out, stream_, &cached);
    if (rc == AMEPB_E_NOMEM && cached) {
        // the cached workspace budget was stale (the caller allocated device memory in between)
        ctx->budget_bytes = 0;
        rc = rdf_batch_impl(ctx, coords, coords_dtype, n, other, other_dtype, m, space, nframes, box_boundary, box_dtype,
                            edges, edges_stride, nbins, pbc, hist_out, dim_out, stream_, &cached);
    }
    return rc;
}

static int rdf_batch_impl(amepb_ctx* ctx, const void* coords, int coords_dtype, int64_t n, const void* other,
                          int other_dtype, int64_t m, int space, int64_t nframes, const double* box_boundary,
                          int box_dtype, const double* edges, int64_t edges_stride, int32_t nbins, int pbc,
                          uint64_t* hist_out, int32_t* dim_out, void* stream_, bool* budget_was_cached) {
    auto dtype_ok = [](int d) { return d == AMEPB_F32 || d == AMEPB_F64; };
    if (!coords || !box_boundary || !edges || !hist_out)
        return fail(ctx, AMEPB_E_INVAL, "amepb_rdf_batch: NULL pointer argument");
    if (nframes < 0 || n <= 0 || nbins <= 0 || (other && m <= 0))
        return fail(ctx, AMEPB_E_INVAL, "amepb_rdf_batch: nframes=%lld n=%lld m=%lld nbins=%d out of range",
                    (long long)nframes, (long long)n, (long long)m, (int)nbins);
    if (!dtype_ok(coords_dtype) || (other && !dtype_ok(other_dtype)) || !dtype_ok(box_dtype))
        return fail(ctx, AMEPB_E_INVAL, "amepb_rdf_batch: unknown dtype code");
    if (space != AMEPB_DEVICE && space != AMEPB_HOST)
        return fail(ctx, AMEPB_E_INVAL, "amepb_rdf_batch: unknown memory space %d", space);
    if (edges_stride != 0 && edges_stride < nbins + 1)
        return fail(ctx, AMEPB_E_INVAL, "amepb_rdf_batch: edges_stride %lld < nbins+1", (long long)edges_stride);
    DeviceGuard guard(ctx->device);
    memset(&ctx->rdf_info, 0, sizeof(ctx->rdf_info));
    timing_reset(ctx);
    const int64_t launches0 = ctx->launches;
    if (nframes == 0) return 0;
    ctx->last_stream = stream_;
    {
        int rc0 = ensure_device(ctx, BUF_MISC, 64);
        if (rc0) return rc0;
        AMEPB_CUDA(ctx, cudaMemsetAsync(ctx->bufs[BUF_MISC].p, 0, 64, (cudaStream_t)stream_));
    }

    RdfCall c;
    c.coords = coords; c.coords_dtype = coords_dtype; c.n = n;
    c.other = other; c.other_dtype = other_dtype; c.m = other ? m : n;
    c.space = space; c.nframes = nframes; c.box = box_boundary; c.box_dtype = box_dtype;
    c.edges = edges; c.edges_stride = edges_stride; c.nbins = nbins; c.pbc = pbc ? 1 : 0;
    c.hist_out = hist_out; c.dim_out = dim_out; c.stream = (cudaStream_t)stream_;
    if (space == AMEPB_HOST && stream_ == nullptr) {
        // host-space calls on the NULL stream get a private compute stream: the legacy default
        // stream would serialise against the staging copies of other blocking streams
        if (!ctx->compute_stream) AMEPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking));
        if (!getenv("AMEPB_HOST_NULL_STREAM")) c.stream = ctx->compute_stream;
    }

    // sub-batch size from a workspace budget
    const size_t cbytes = (size_t)n * 3 * elem_size(coords_dtype);
    const size_t obytes = other ? (size_t)m * 3 * elem_size(other_dtype) : 0;
    const size_t per_frame = (size_t)n * (c.pbc ? 8 : 1) * 32 + (size_t)c.m * 32 +
                             (size_t)std::max<int64_t>(4096, n * 2) * 16 +
                             (space == AMEPB_HOST ? cbytes + obytes : 0);
    auto wall_ms = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double w0 = wall_ms();
    // The memory budget is queried once per problem shape: cudaMemGetInfo takes a driver-wide
    // lock and was seen to stall for tens of milliseconds on shared hosts.
    *budget_was_cached = true;
    if (ctx->budget_per_frame != per_frame || ctx->budget_bytes == 0) {
        *budget_was_cached = false;
        size_t free_b = 0, total_b = 0;
        AMEPB_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        ctx->budget_bytes = std::min<size_t>((size_t)24 << 30, (free_b + ctx->bufs[BUF_TSORTED].cap) / 3);
        ctx->budget_per_frame = per_frame;
    }
    const double w1 = wall_ms();
    size_t budget = std::max<size_t>(ctx->budget_bytes, per_frame);
    int64_t fb = (int64_t)std::max<size_t>(1, budget / per_frame);
    fb = std::min<int64_t>(fb, 16384);
    fb = std::min<int64_t>(fb, std::max<int64_t>(1, (int64_t)((1ll << 31) / std::max<int64_t>(1, n * 8))));
    if (const char* cap = getenv("AMEPB_MAX_FRAMES")) fb = std::min<int64_t>(fb, std::max(1, atoi(cap)));   // tests: force sub-batching
    fb = std::min<int64_t>(fb, nframes);

    bool bad_dim = false;
    ctx->stats_ahead_f0 = -1;   // nothing in flight from an earlier (possibly failed) call
    if (ctx->stats_stream) AMEPB_CUDA(ctx, cudaStreamSynchronize(ctx->stats_stream));
    if (space == AMEPB_DEVICE) {
        for (int64_t f0 = 0; f0 < nframes; f0 += fb) {
            const int nf = (int)std::min<int64_t>(fb, nframes - f0);
            const void* d_coords = (const char*)coords + (size_t)f0 * cbytes;
            const void* d_other = other ? (const char*)other + (size_t)f0 * obytes : nullptr;
            unsigned long long* d_hist = reinterpret_cast<unsigned long long*>(hist_out) + (size_t)f0 * nbins;
            NextBatch nb;
            nb.f0 = f0 + nf;
            nb.nf = (int)std::min<int64_t>(fb, nframes - nb.f0);
            nb.d_coords = (const char*)coords + (size_t)nb.f0 * cbytes;
            nb.d_other = other ? (const char*)other + (size_t)nb.f0 * obytes : nullptr;
            nb.ready = nullptr;
            int rc = rdf_sub_batch(ctx, c, f0, nf, d_coords, d_other, d_hist, &bad_dim, nb.nf > 0 ? &nb : nullptr);
            if (rc) return rc;
        }
        if (getenv("AMEPB_TRACE") && ctx->timing && !ctx->ev_build.empty()) {
            AMEPB_CUDA(ctx, cudaStreamSynchronize(c.stream));
            cudaEvent_t t0 = ctx->ev_build[0].first;
            auto rel = [&](cudaEvent_t e) {
                float ms = 0;
                cudaEventElapsedTime(&ms, t0, e);
                return ms;
            };
            for (size_t i = 0; i < ctx->ev_build.size() && i < ctx->ev_pair.size(); ++i)
                fprintf(stderr, "[amepb] sub %zu: build %.2f-%.2f  pair %.2f-%.2f ms\n", i, rel(ctx->ev_build[i].first),
                        rel(ctx->ev_build[i].second), rel(ctx->ev_pair[i].first), rel(ctx->ev_pair[i].second));
        }
    } else {
        // Host buffers: staging through kStageBuffers device buffers.  The copies of the next two
        // sub-batches run on the context's copy stream while sub-batch i is computed on the
        // caller's stream; the extra buffer absorbs the jitter of the host link on shared hosts.
        size_t stage_target = (size_t)384 << 20;
        if (const char* env = getenv("AMEPB_STAGE_MB")) stage_target = (size_t)std::max(1, atoi(env)) << 20;
        fb = std::min<int64_t>(fb, std::max<int64_t>(1, (int64_t)(stage_target / std::max<size_t>(1, cbytes + obytes))));
        // Sub-batches grow slowly (x1.15) from fb/8 to fb frames.  The first copy (not overlapped
        // with anything) stays short and later sub-batches amortise their fixed costs; the growth
        // factor follows from the rates: a copy takes about as long as the computation of the same
        // frames (0.217 vs 0.223 ms per 1M-particle frame), so a sub-batch much larger than its
        // predecessor would leave the GPU waiting for its copy (a doubling schedule idled 9 %)
        std::vector<std::pair<int64_t, int>> subs;
        {
            int64_t f0 = 0;
            double sz = (double)std::max<int64_t>(1, fb / 8);
            const int64_t min_sz = std::max<int64_t>(1, fb / 8);
            while (f0 < nframes) {
                int64_t nf = std::min<int64_t>((int64_t)(sz + 0.5), nframes - f0);
                // ... and they shrink again towards the end (never more than a third of what is left):
                // the pipeline is bound by the copies, so what remains after the last copy has landed
                // is the computation of the last sub-batch -- 7.5 ms for 32 frames, 1 ms for 4
                nf = std::min<int64_t>(nf, std::max<int64_t>(min_sz, (nframes - f0 + 2) / 3));
                if (nframes - f0 - nf < std::min<int64_t>(nf, min_sz) / 2) nf = nframes - f0;   // no tiny last sub-batch
                nf = std::min<int64_t>(nf, fb);
                subs.emplace_back(f0, (int)nf);
                f0 += nf;
                sz = std::min<double>((double)fb, std::max(sz + 1.0, sz * 1.15));
            }
        }
        const int64_t nsub = (int64_t)subs.size();
        constexpr int NB = amepb_ctx::kStageBuffers;
        int rc = ensure_device(ctx, BUF_STAGE_A, NB * cbytes * fb);
        if (rc) return rc;
        if (other && (rc = ensure_device(ctx, BUF_STAGE_B, NB * obytes * fb))) return rc;
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, sizeof(unsigned long long) * (size_t)nframes * nbins))) return rc;
        if (!ctx->copy_stream) {
            AMEPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int b = 0; b < NB; ++b) {
                AMEPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
                AMEPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_computed[b], cudaEventDisableTiming));
            }
        }
        char* stage_a = reinterpret_cast<char*>(ctx->bufs[BUF_STAGE_A].p);
        char* stage_b = other ? reinterpret_cast<char*>(ctx->bufs[BUF_STAGE_B].p) : nullptr;
        unsigned long long* d_hist = reinterpret_cast<unsigned long long*>(ctx->bufs[BUF_STAGE_OUT].p);
        auto enqueue_copy = [&](int64_t i) -> int {
            const int b = (int)(i % NB);
            const int64_t f0 = subs[i].first;
            const int nf = subs[i].second;
            if (i >= NB) AMEPB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_computed[b], 0));
            cudaEvent_t ec0 = timing_event(ctx, ctx->copy_stream);
            AMEPB_CUDA(ctx, cudaMemcpyAsync(stage_a + (size_t)b * cbytes * fb, (const char*)coords + (size_t)f0 * cbytes,
                                            cbytes * nf, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (other)
                AMEPB_CUDA(ctx, cudaMemcpyAsync(stage_b + (size_t)b * obytes * fb, (const char*)other + (size_t)f0 * obytes,
                                                obytes * nf, cudaMemcpyHostToDevice, ctx->copy_stream));
            AMEPB_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
            cudaEvent_t ec1 = timing_event(ctx, ctx->copy_stream);
            if (ec0 && ec1) ctx->ev_copy.emplace_back(ec0, ec1);
            return 0;
        };
        // the copy stream must not start before work already queued on the caller's stream
        // could still be reading the staging buffers of an earlier call
        AMEPB_CUDA(ctx, cudaStreamSynchronize(c.stream));
        const double w2 = wall_ms();
        for (int64_t i = 0; i < std::min<int64_t>(NB - 1, nsub); ++i)
            if ((rc = enqueue_copy(i))) return rc;
        const bool trace = getenv("AMEPB_TRACE") != nullptr;
        auto now_ms = []() {
            timespec ts;
            clock_gettime(CLOCK_MONOTONIC, &ts);
            return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
        };
        for (int64_t i = 0; i < nsub; ++i) {
            const int b = (int)(i % NB);
            const int64_t f0 = subs[i].first;
            const int nf = subs[i].second;
            const double t0 = now_ms();
            if (i + NB - 1 < nsub && (rc = enqueue_copy(i + NB - 1))) return rc;
            if (trace) fprintf(stderr, "[amepb] sub %lld: enqueue_copy took %.3f ms (host)\n", (long long)i, now_ms() - t0);
            AMEPB_CUDA(ctx, cudaStreamWaitEvent(c.stream, ctx->ev_copied[b], 0));
            NextBatch nb;
            nb.nf = 0;
            if (i + 1 < nsub) {
                const int b1 = (int)((i + 1) % NB);
                nb.f0 = subs[i + 1].first;
                nb.nf = subs[i + 1].second;
                nb.d_coords = stage_a + (size_t)b1 * cbytes * fb;
                nb.d_other = other ? stage_b + (size_t)b1 * obytes * fb : nullptr;
                nb.ready = ctx->ev_copied[b1];
            }
            rc = rdf_sub_batch(ctx, c, f0, nf, stage_a + (size_t)b * cbytes * fb,
                               other ? stage_b + (size_t)b * obytes * fb : nullptr, d_hist + (size_t)f0 * nbins, &bad_dim,
                               nb.nf > 0 ? &nb : nullptr);
            if (rc) return rc;
            AMEPB_CUDA(ctx, cudaEventRecord(ctx->ev_computed[b], c.stream));
        }
        // one device->host transfer of all histograms at the end (a per-sub-batch copy would
        // queue behind the next staging copy and stall the stream)
        AMEPB_CUDA(ctx, cudaMemcpyAsync(hist_out, d_hist, sizeof(unsigned long long) * (size_t)nframes * nbins,
                                        cudaMemcpyDeviceToHost, c.stream));
        const double w3 = wall_ms();
        AMEPB_CUDA(ctx, cudaStreamSynchronize(c.stream));
        const double w4 = wall_ms();
        if (const char* slow = getenv("AMEPB_TRACE_SLOW_MS"))
            if (w4 - w0 > atof(slow))
                fprintf(stderr, "[amepb] host-buffer call: %.1f ms wall = meminfo %.2f + setup %.2f + loop %.2f + final sync %.2f\n",
                        w4 - w0, w1 - w0, w2 - w1, w3 - w2, w4 - w3);
        bool trace_now = trace;
        if (const char* slow = getenv("AMEPB_TRACE_SLOW_MS")) {
            // print the timeline only for calls slower than the given time (ms)
            trace_now = false;
            if (ctx->timing && !ctx->ev_copy.empty() && !ctx->ev_pair.empty()) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ctx->ev_copy[0].first, ctx->ev_pair.back().second);
                trace_now = ms > atof(slow);
            }
        }
        if (trace_now && ctx->timing && !ctx->ev_build.empty()) {
            cudaEvent_t t0 = ctx->ev_copy.empty() ? ctx->ev_build[0].first : ctx->ev_copy[0].first;
            auto rel = [&](cudaEvent_t e) {
                float ms = 0;
                cudaEventElapsedTime(&ms, t0, e);
                return ms;
            };
            for (size_t i = 0; i < ctx->ev_build.size(); ++i) {
                fprintf(stderr, "[amepb] sub %zu: copy %.2f-%.2f  build %.2f-%.2f  pair %.2f-%.2f ms\n", i,
                        i < ctx->ev_copy.size() ? rel(ctx->ev_copy[i].first) : -1.f,
                        i < ctx->ev_copy.size() ? rel(ctx->ev_copy[i].second) : -1.f,
                        rel(ctx->ev_build[i].first), rel(ctx->ev_build[i].second),
                        i < ctx->ev_pair.size() ? rel(ctx->ev_pair[i].first) : -1.f,
                        i < ctx->ev_pair.size() ? rel(ctx->ev_pair[i].second) : -1.f);
            }
        }
    }
    ctx->rdf_info.kernel_launches = ctx->launches - launches0;
    if (bad_dim)
        return fail(ctx, AMEPB_E_DIMENSION,
                    "amep.pbc.pbc_points: Only 2d or 3d systems are supported. Please check coords and enforce_nd.");
    return 0;
}

}  // extern "C"

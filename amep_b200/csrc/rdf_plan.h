// rdf_plan.h -- host-side planning for the RDF pair kernel (pure C++, no CUDA).
//
// Two jobs, both exact by construction:
//  (1) squared-distance thresholds: for every bin edge e the smallest squared
//      distance s (in the arithmetic type of the regime) whose correctly
//      rounded square root reaches e.  Binning s against these thresholds is
//      identical to NumPy's explicit-edge histogram of d = sqrt(s)
//      (amep/spatialcor.py:382-387) because IEEE sqrt is monotone.
//  (2) the per-frame cell grid: region, cell size, neighbour rows and the
//      decomposition into block tasks.  The grid is only a candidate
//      generator; membership is decided by the exact arithmetic in the kernel.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

namespace amepb {

constexpr int kMaxRows = 169;  // (2*6+1)^2 neighbour rows at most
constexpr int kMaxCellsPerCutoff = 6;

// ---------------------------------------------------------------------------
// per-frame statistics produced by the stats kernel
// ---------------------------------------------------------------------------
struct FrameStats {
    double lo[3];      // min per component (NaNs ignored)
    double hi[3];      // max per component
    double first[3];   // component values of particle 0
    int varies[3];     // column differs from particle 0 somewhere
};

// dimension() of amep/utils.py:1617-1667 from the column flags.
inline int dimension_from_stats(const FrameStats& s, int64_t n) {
    int dim = s.varies[0] + s.varies[1] + s.varies[2];
    if (dim == 0 && n == 1) dim = 3;
    return dim;
}

// ---------------------------------------------------------------------------
// (1) thresholds
// ---------------------------------------------------------------------------
template <typename T>
struct SqrtOps;
template <>
struct SqrtOps<float> {
    static double root(float s) { return (double)sqrtf(s); }
    static float next_up(float s) { return nextafterf(s, std::numeric_limits<float>::infinity()); }
    static float next_down(float s) { return nextafterf(s, -std::numeric_limits<float>::infinity()); }
};
template <>
struct SqrtOps<double> {
    static double root(double s) { return sqrt(s); }
    static double next_up(double s) { return nextafter(s, std::numeric_limits<double>::infinity()); }
    static double next_down(double s) { return nextafter(s, -std::numeric_limits<double>::infinity()); }
};

// smallest s >= 0 of type T with  sqrt_T(s) >= e  (strict: > e), compared in
// float64 after widening, exactly as NumPy compares the float64 `dist` array
// with the edges.  Returns +inf if no finite s qualifies.
template <typename T>
inline T sq_threshold(double e, bool strict) {
    const T inf = std::numeric_limits<T>::infinity();
    if (e != e) return inf;  // NaN edge: nothing compares true
    auto ok = [&](T s) { double r = SqrtOps<T>::root(s); return strict ? (r > e) : (r >= e); };
    if (ok((T)0)) return (T)0;
    if (!(e < std::numeric_limits<double>::infinity())) return inf;
    double guess = e * e;
    T s = (guess >= (double)std::numeric_limits<T>::max()) ? std::numeric_limits<T>::max() : (T)guess;
    if (!(s > (T)0)) s = std::numeric_limits<T>::denorm_min();
    if (!ok(s)) {
        while (!ok(s)) {
            if (s == std::numeric_limits<T>::max()) return inf;
            s = SqrtOps<T>::next_up(s);
        }
        return s;
    }
    for (;;) {
        T d = SqrtOps<T>::next_down(s);
        if (d < (T)0 || !ok(d)) return s;
        s = d;
    }
}

// thr[0]      : smallest s that passes both d > 1e-3 and d >= edges[0]
// thr[i]      : smallest s with d >= edges[i]            (0 < i < nbins)
// thr[nbins]  : smallest s with d >  edges[nbins]        (exclusive upper end;
//               the last bin is right-inclusive)
// A pair is counted iff thr[0] <= s < thr[nbins]; its bin is the number of
// i in [1, nbins) with thr[i] <= s.
template <typename T>
inline void make_thresholds(const double* edges, int nbins, T* thr) {
    T a = sq_threshold<T>(1e-3, true);
    T b = sq_threshold<T>(edges[0], false);
    thr[0] = a > b ? a : b;
    for (int i = 1; i < nbins; ++i) thr[i] = sq_threshold<T>(edges[i], false);
    thr[nbins] = sq_threshold<T>(edges[nbins], true);
}

// ---------------------------------------------------------------------------
// (2) per-frame plan (this struct is copied to the device verbatim)
// ---------------------------------------------------------------------------
struct FramePlan {
    double g0[3];      // grid origin
    double inv_h[3];   // cells per unit length
    double win_lo[3];  // image window, exclusive (amep/pbc.py:332-338)
    double win_hi[3];
    double reg_lo[3];  // grid region, inclusive: only images inside are kept
    double reg_hi[3];
    float keep_lo[3];  // window and region as one closed float32 interval: a float32 image x is
    float keep_hi[3];  // kept iff keep_lo <= x <= keep_hi (same decisions as the four tests above)
    float near_lo[3];  // a particle with near_lo < x < near_hi in every shifted dimension cannot have a
    float near_hi[3];  // kept *shifted* image (conservative: only a shortcut, never a decision)
    double dz2c;       // constant dz*dz term of the z-constant kernels
    double bin_e0;     // first edge and 1/width for the bin guess
    double bin_inv_w;
    long long task_begin;  // first block task of this frame
    long long ntasks;
    float shift[3];        // float32 box lengths (amep/pbc.py:300)
    float cull_r2;         // (largest edge)^2 with a 1e-4 margin: targets farther than this from a query box are skipped
    int n[3];              // cells per dimension
    int nshift[3];         // 3 where images are shifted along the dimension, else 1
    int cq_lo[3];          // query cell range
    int cq_n[3];
    unsigned cell_base;    // offset of the frame's cells in the flat cell arrays
    unsigned ncells;
    int nrows;             // neighbour rows
    int run;               // query cells per block task (along x)
    int rsplit;            // warps sharing one query cell (strided over its neighbour rows)
    int runs_per_row;
    int thr_index;         // threshold table of this frame
    int valid;             // 0: nothing to do for this frame
    int dim;               // dimension(coords)
    unsigned direct_pairs; // warp tasks with more pairs than this bypass the 32-bit shared histogram
    int row0;              // index of the (oy, oz) = (0, 0) row; rows after it are the "upper" half shell
    int sym;               // symmetric mode: real-real pairs once (weight 2), ghosts separately
    int compact;           // the float32 sorted arrays hold (x, y) records (flat frames, staged float32 walk)
    // partitioned build (rdf.cu: strip_* kernels): strips = runs of strip_rows consecutive cell rows
    int strip_rows;        // cell rows (fixed y, z; all x) per strip; 0: this frame is not partitioned
    int nstrips;
    unsigned strip_magic;  // row / strip_rows == umulhi(row, strip_magic) (0: strip_rows == 1)
    unsigned qbase;        // first slot of the frame's queries in the sorted query array
    int kd[3];             // neighbourhood half-widths in cells
    int gfree_lo[3];       // cells [gfree_lo, gfree_hi] (all dims) cannot hold ghost images
    int gfree_hi[3];
    signed char row_oy[kMaxRows];
    signed char row_oz[kMaxRows];
    signed char row_kx[kMaxRows];
    char pad_[1];
};

// Smallest float32 f with f >= b (strict: f > b) / largest with f <= b (strict: f < b), as
// decided in float64 -- turns a float64 comparison of a float32 value into a float32 one.
inline float float_above(double b, bool strict) {
    if (b != b) return std::numeric_limits<float>::quiet_NaN();
    float f = (float)b;
    while (strict ? !((double)f > b) : !((double)f >= b)) {
        if (f == std::numeric_limits<float>::infinity()) return std::numeric_limits<float>::quiet_NaN();
        f = nextafterf(f, INFINITY);
    }
    return f;
}
inline float float_below(double b, bool strict) {
    if (b != b) return std::numeric_limits<float>::quiet_NaN();
    float f = (float)b;
    while (strict ? !((double)f < b) : !((double)f <= b)) {
        if (f == -std::numeric_limits<float>::infinity()) return std::numeric_limits<float>::quiet_NaN();
        f = nextafterf(f, -INFINITY);
    }
    return f;
}

struct PlanInputs {
    const double* box_boundary;  // [3][2] float64 values
    int box_is_f32;              // caller's box dtype was float32
    int pbc;
    double rc;                   // largest edge
    int64_t n_targets;           // particles in coords
    int64_t n_queries;
    int cells_per_cutoff;        // 0 = automatic
    int cells_per_cutoff_dim[3] = {0, 0, 0};   // per-dimension override (0 = follow cells_per_cutoff / automatic)
    int64_t max_cells;           // upper bound on cells for one frame
    double task_pairs = 262144.0;  // candidate pairs aimed at per block task
    int heavy_run_mult = 0;        // query cells per block task multiplier when warps share cells (0 = automatic)
};

// Image window of pbc_points(width=0.5): (centre - L32, centre + L32), with
// centre = mean(box_boundary, axis=1) in the box's own dtype and L32 the
// float32 box length (amep/pbc.py:270-273, 300, 332-338).
inline void image_window(const double* bb, int box_is_f32, double* win_lo, double* win_hi, float* l32) {
    for (int d = 0; d < 3; ++d) {
        double lo = bb[2 * d], hi = bb[2 * d + 1];
        if (box_is_f32) {
            float flo = (float)lo, fhi = (float)hi;
            float len = fhi - flo;               // float32 subtraction (amep/pbc.py:270)
            float centre = (flo + fhi) / 2.0f;   // np.mean of two float32 values
            float width = len * 1.0f;            // box*(.5+width), width = 0.5
            l32[d] = len;
            win_hi[d] = (double)(centre + width);
            win_lo[d] = (double)(centre - width);
        } else {
            double len = hi - lo;
            double centre = (lo + hi) / 2.0;
            float len32 = (float)len;            // box.astype(np.float32)
            float width = len32 * 1.0f;
            l32[d] = len32;
            win_hi[d] = centre + (double)width;
            win_lo[d] = centre - (double)width;
        }
    }
}

inline int cell_of(double x, double g0, double inv_h, int n) {
    double c = floor((x - g0) * inv_h);
    if (!(c > 0.0)) return 0;  // also NaN
    if (c > (double)(n - 1)) return n - 1;
    return (int)c;
}

// Plan one frame.  `dim` = dimension(coords); `tgt`/`qry` = stats of the
// target and query particles.  Returns false (plan.valid = 0) when no pair can
// be counted.
inline bool plan_frame(const PlanInputs& in, int dim, const FrameStats& tgt, const FrameStats& qry,
                       FramePlan& P) {
    memset(&P, 0, sizeof(P));
    P.dim = dim;
    const double inf = std::numeric_limits<double>::infinity();
    const double rc = in.rc;
    if (!(rc > 0.0) || !(rc < inf) || in.n_targets <= 0 || in.n_queries <= 0) return false;
    const double rcs = rc * (1.0 + 1e-5) + 1e-300;  // cutoff plus a rounding skin

    double ib_lo[3], ib_hi[3];  // bounds of all images that can exist
    if (in.pbc) {
        image_window(in.box_boundary, in.box_is_f32, P.win_lo, P.win_hi, P.shift);
        for (int d = 0; d < 3; ++d) {
            bool shifted = (d < 2) || (dim == 3);  // 2D shifts x,y only (amep/pbc.py:305-309)
            P.nshift[d] = shifted ? 3 : 1;
            double ext = shifted ? fabs((double)P.shift[d]) : 0.0;
            ib_lo[d] = std::max(P.win_lo[d], tgt.lo[d] - ext - 1e-3 * (fabs(tgt.lo[d]) + ext + 1));
            ib_hi[d] = std::min(P.win_hi[d], tgt.hi[d] + ext + 1e-3 * (fabs(tgt.hi[d]) + ext + 1));
        }
    } else {
        for (int d = 0; d < 3; ++d) {
            P.win_lo[d] = -inf;
            P.win_hi[d] = inf;
            P.shift[d] = 0.0f;
            P.nshift[d] = 1;
            ib_lo[d] = tgt.lo[d];
            ib_hi[d] = tgt.hi[d];
        }
    }
    double ext[3];
    for (int d = 0; d < 3; ++d) {
        double lo = std::max(ib_lo[d], qry.lo[d] - rcs);
        double hi = std::min(ib_hi[d], qry.hi[d] + rcs);
        if (!(lo <= hi)) return false;  // empty region (also NaN bounds)
        P.reg_lo[d] = lo;
        P.reg_hi[d] = hi;
        ext[d] = hi - lo;
        P.keep_lo[d] = std::max(float_above(P.win_lo[d], true), float_above(lo, false));
        P.keep_hi[d] = std::min(float_below(P.win_hi[d], true), float_below(hi, false));
        // x + L <= keep_hi needs x <= keep_hi - L (+ rounding), x - L >= keep_lo needs x >= keep_lo + L (- rounding)
        P.near_lo[d] = std::numeric_limits<float>::infinity();     // default: every particle takes the full test
        P.near_hi[d] = -std::numeric_limits<float>::infinity();
        if (in.pbc && P.nshift[d] == 3) {
            const double L = (double)P.shift[d];
            const double slack = 1e-5 * (fabs((double)P.keep_lo[d]) + fabs((double)P.keep_hi[d]) + fabs(L) + 1.0);
            if (L > 0 && P.keep_lo[d] <= P.keep_hi[d]) {
                P.near_lo[d] = float_above((double)P.keep_hi[d] - L + slack, false);
                P.near_hi[d] = float_below((double)P.keep_lo[d] + L - slack, false);
            }
        }
    }

    // expected number of images in the region (uniform-density estimate)
    double n_est = (double)in.n_targets;
    for (int d = 0; d < 3; ++d) {
        double span = tgt.hi[d] - tgt.lo[d];
        if (span > 0 && ext[d] > 0) n_est *= std::min(3.0, std::max(ext[d] / span, 1e-3));
    }

    // grid refinement: the largest k that still leaves a useful number of
    // particles per cell
    auto cells_for = [&](int k, int* n) {
        double total = 1;
        for (int d = 0; d < 3; ++d) {
            const int kdim = in.cells_per_cutoff_dim[d] > 0 ? in.cells_per_cutoff_dim[d] : k;
            double want = floor(ext[d] * kdim / rcs);
            n[d] = (int)std::min(std::max(want, 1.0), 1.0e6);
            total *= n[d];
        }
        return total;
    };
    int k = in.cells_per_cutoff;
    int n[3];
    if (k <= 0) {
        const int kmax = (dim == 3) ? 4 : kMaxCellsPerCutoff;
        const double want_ppc = (dim == 3) ? 20.0 : 12.0;
        k = 1;
        for (int kk = 2; kk <= kmax; ++kk) {
            double cells = cells_for(kk, n);
            if (n_est / cells >= want_ppc) k = kk;
        }
    }
    k = std::min(std::max(k, 1), kMaxCellsPerCutoff);
    double total = cells_for(k, n);
    // cap the number of cells (sparse systems): grow the cells uniformly
    double cap = (double)std::max<int64_t>(in.max_cells, 1);
    while (total > cap) {
        double f = pow(cap / total, 1.0 / 3.0) * 0.999;
        total = 1;
        for (int d = 0; d < 3; ++d) {
            if (n[d] > 1) n[d] = std::max(1, (int)floor(n[d] * f));
            total *= n[d];
        }
    }
    double h[3];
    int kd[3];
    for (int d = 0; d < 3; ++d) {
        P.n[d] = n[d];
        P.g0[d] = P.reg_lo[d];
        h[d] = ext[d] > 0 ? ext[d] / n[d] : inf;
        P.inv_h[d] = ext[d] > 0 ? (double)n[d] / ext[d] : 0.0;
        kd[d] = (n[d] == 1) ? 0 : (int)std::min<double>(n[d] - 1, ceil(rcs / h[d]));
    }
    if (kd[1] > kMaxCellsPerCutoff || kd[2] > kMaxCellsPerCutoff) return false;  // cannot happen: h >= rcs/k
    P.ncells = (unsigned)(n[0] * (long long)n[1] * n[2]);
    for (int d = 0; d < 3; ++d) P.kd[d] = kd[d];

    // neighbour rows: all (oy, oz) cell offsets that can hold an image within
    // the cutoff of a query in the centre cell, with the x half-width needed
    int nrows = 0;
    for (int oz = -kd[2]; oz <= kd[2]; ++oz) {
        for (int oy = -kd[1]; oy <= kd[1]; ++oy) {
            double gy = std::max(abs(oy) - 1, 0) * (n[1] > 1 ? h[1] : 0.0);
            double gz = std::max(abs(oz) - 1, 0) * (n[2] > 1 ? h[2] : 0.0);
            double rem = rcs * rcs - gy * gy - gz * gz;
            if (rem < 0) continue;
            int kx = (n[0] == 1) ? 0 : (int)std::min<double>(kd[0], floor(sqrt(rem) / h[0]) + 1);
            if (oy == 0 && oz == 0) P.row0 = nrows;
            P.row_oy[nrows] = (signed char)oy;
            P.row_oz[nrows] = (signed char)oz;
            P.row_kx[nrows] = (signed char)kx;
            ++nrows;
        }
    }
    P.nrows = nrows;

    // query cell range
    double cells_q = 1;
    for (int d = 0; d < 3; ++d) {
        int a = cell_of(qry.lo[d], P.g0[d], P.inv_h[d], n[d]);
        int b = cell_of(qry.hi[d], P.g0[d], P.inv_h[d], n[d]);
        P.cq_lo[d] = a;
        P.cq_n[d] = b - a + 1;
        cells_q *= P.cq_n[d];
    }

    // cells that cannot hold a shifted image (symmetric mode skips the ghost pass there):
    // a shifted image of a particle in [tgt.lo, tgt.hi] lies outside (tgt.hi - L, tgt.lo + L)
    // in at least one shifted dimension
    for (int d = 0; d < 3; ++d) {
        P.gfree_lo[d] = 0;
        P.gfree_hi[d] = n[d] - 1;
        if (in.pbc && P.nshift[d] == 3) {
            const double len = fabs((double)P.shift[d]);
            // slack: float32 rounding of the cast and of base + shift
            const double slack = 1e-5 * (fabs(tgt.lo[d]) + fabs(tgt.hi[d]) + len + 1.0);
            P.gfree_lo[d] = cell_of(tgt.hi[d] - len + slack, P.g0[d], P.inv_h[d], n[d]) + 1;
            P.gfree_hi[d] = cell_of(tgt.lo[d] + len - slack, P.g0[d], P.inv_h[d], n[d]) - 1;
        }
    }

    // block tasks: runs of query cells along x
    double ppc_t = n_est / std::max(1.0, (double)P.ncells);
    double ppc_q = (double)in.n_queries / std::max(1.0, cells_q);
    double pairs_per_cell = std::max(1.0, ppc_q) * std::max(1.0, ppc_t) * nrows * (2 * kd[0] + 1);
    // A block task is `run` consecutive query cells; its 8 warps each take one cell (run a
    // multiple of 8) or, for heavy cells, a strided 1/rsplit share of the cell's neighbour rows.
    // Aim at ~2.6e5 candidate pairs per block task.
    int run = (int)std::min<double>(256.0, std::max(1.0, floor(in.task_pairs / pairs_per_cell + 0.5)));
    int rsplit = 1;
    if (run >= 8) run = (run / 8) * 8;
    else if (run >= 4) { run = 4; rsplit = 2; }
    else if (run >= 2) { run = 2; rsplit = 4; }
    else { run = 1; rsplit = 8; }
    if (run > P.cq_n[0]) {
        run = P.cq_n[0];
        rsplit = run >= 8 ? 1 : (run >= 4 ? 2 : (run >= 2 ? 4 : 8));
    }
    // shared cells: hand out several times more items than warps, so that the dynamic hand-out
    // inside the task can even out rows of different length before the block barrier
    // (x4 when the frame still yields thousands of block tasks: 5 % on the 3D configs; fewer,
    // larger tasks would unbalance the persistent blocks of small frames instead)
    if (rsplit > 1) {
        int mult = in.heavy_run_mult;
        if (mult <= 0) {
            mult = 1;
            for (int m = 4; m > 1; m /= 2) {
                const int r = std::min<int>(P.cq_n[0], run * m);
                const long long tasks = (long long)((P.cq_n[0] + r - 1) / r) * P.cq_n[1] * P.cq_n[2];
                if (tasks >= 4096) {
                    mult = m;
                    break;
                }
            }
        }
        run = std::min<int>(P.cq_n[0], run * mult);
    }
    P.run = run;
    P.rsplit = rsplit;
    P.cull_r2 = (float)(rc * rc * (1.0 + 1e-4)) * (1.0f + 1e-6f);
    // 32-bit shared-memory counters: a block task has at most run * (2*nrows + 1) warp tasks
    // (rows, the split own row, the ghost pass); warp tasks with more pairs than direct_pairs
    // bypass the shared histogram, so a block task adds fewer than 2^31 increments and the
    // flush check between block tasks (at 2^31) keeps every counter below 2^32
    P.direct_pairs = (unsigned)std::min<double>(4194304.0, 2147483648.0 / ((double)run * (2 * nrows + 1)));
    P.runs_per_row = (P.cq_n[0] + run - 1) / run;
    P.ntasks = (long long)P.runs_per_row * P.cq_n[1] * P.cq_n[2];
    P.valid = 1;
    return true;
}

}  // namespace amepb

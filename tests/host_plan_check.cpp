// host_plan_check.cpp -- CPU check of amep_b200/csrc/rdf_plan.h (test only).
//
// (1) squared-distance thresholds bin exactly like np.histogram on d = sqrt(s);
// (2) a serial walk over the planned grid (same cell / row / task decomposition
//     the pair kernel uses) visits every in-range (query, image) pair exactly
//     once: its histogram equals the brute-force one.
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "rdf_plan.h"

using namespace amepb;

static int g_fail = 0;
#define CHECK(cond, ...)                       \
    do {                                       \
        if (!(cond)) {                         \
            printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            printf(__VA_ARGS__);               \
            printf("\n");                      \
            ++g_fail;                          \
        }                                      \
    } while (0)

// np.histogram with explicit edges on one value (after the d > 1e-3 filter)
static int numpy_bin(double d, const std::vector<double>& e) {
    int nb = (int)e.size() - 1;
    if (!(d > 1e-3)) return -1;
    if (!(d >= e[0]) || !(d <= e[nb])) return -1;
    int a = 0, b = nb;
    while (b - a > 1) {
        int m = (a + b) / 2;
        if (e[m] <= d) a = m; else b = m;
    }
    return a;
}

template <typename T>
static int thr_bin(T s, const std::vector<T>& thr, int nb) {
    if (!(s >= thr[0] && s < thr[nb])) return -1;
    int g = 0;
    for (int i = 1; i < nb; ++i)
        if (thr[i] <= s) g = i;
    return g;
}

template <typename T>
static void check_thresholds(std::mt19937_64& rng, bool edges_f32, double rmax, int nb) {
    std::vector<double> e(nb + 1);
    for (int i = 0; i <= nb; ++i) {
        double v = rmax * i / nb;
        e[i] = edges_f32 ? (double)(float)v : v;
    }
    e[nb] = edges_f32 ? (double)(float)rmax : rmax;
    std::vector<T> thr(nb + 1);
    make_thresholds<T>(e.data(), nb, thr.data());
    std::uniform_real_distribution<double> U(0.0, 1.0);
    // values at and around every threshold, plus random ones
    for (int i = 0; i <= nb; ++i) {
        T s = thr[i];
        for (int k = -3; k <= 3; ++k) {
            T v = s;
            for (int j = 0; j < abs(k); ++j) v = k < 0 ? SqrtOps<T>::next_down(v) : SqrtOps<T>::next_up(v);
            if (!(v >= 0)) continue;
            int want = numpy_bin(SqrtOps<T>::root(v), e);
            int got = thr_bin<T>(v, thr, nb);
            CHECK(want == got, "threshold edge %d k=%d want %d got %d", i, k, want, got);
        }
    }
    for (int t = 0; t < 200000; ++t) {
        double d = U(rng) * rmax * 1.1;
        T v = (T)(d * d);
        int want = numpy_bin(SqrtOps<T>::root(v), e);
        int got = thr_bin<T>(v, thr, nb);
        CHECK(want == got, "random s want %d got %d", want, got);
    }
}

struct Img {
    float x[3];
    int ghost;  // shifted image
};

// Walk the plan exactly like the kernels do and histogram the pairs.
template <typename QT>
static void check_plan(std::mt19937_64& rng, int dim, bool box_f32, const double bb[6], double rmax, int nb,
                       int n, int m, bool self, double spill, int k_override, const char* tag) {
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<QT> tgt(3 * (size_t)n), qry;
    auto fill = [&](std::vector<QT>& v, int cnt) {
        v.resize(3 * (size_t)cnt);
        for (int i = 0; i < cnt; ++i)
            for (int d = 0; d < 3; ++d) {
                double lo = bb[2 * d], hi = bb[2 * d + 1], span = hi - lo;
                double x = lo - spill * span + U(rng) * span * (1 + 2 * spill);
                if (dim == 2 && d == 2) x = 0.0;
                v[3 * (size_t)i + d] = (QT)x;
            }
    };
    fill(tgt, n);
    if (self) qry = tgt, m = n;
    else fill(qry, m);

    auto stats_of = [&](const std::vector<QT>& v, int cnt) {
        FrameStats s;
        for (int d = 0; d < 3; ++d) {
            s.lo[d] = 1e300; s.hi[d] = -1e300; s.varies[d] = 0; s.first[d] = (double)v[d];
        }
        for (int i = 0; i < cnt; ++i)
            for (int d = 0; d < 3; ++d) {
                double x = (double)v[3 * (size_t)i + d];
                s.lo[d] = std::min(s.lo[d], x); s.hi[d] = std::max(s.hi[d], x);
                if (!(v[3 * (size_t)i + d] == v[d])) s.varies[d] = 1;
            }
        return s;
    };
    FrameStats ts = stats_of(tgt, n), qs = stats_of(qry, m);
    CHECK(dimension_from_stats(ts, n) == dim, "%s: dimension", tag);

    std::vector<double> e(nb + 1);
    for (int i = 0; i <= nb; ++i) e[i] = rmax * i / nb;
    PlanInputs in;
    in.box_boundary = bb; in.box_is_f32 = box_f32; in.pbc = 1; in.rc = e[nb];
    in.n_targets = n; in.n_queries = m; in.cells_per_cutoff = k_override; in.max_cells = std::max(4096, 2 * n);
    FramePlan P;
    bool ok = plan_frame(in, dim, ts, qs, P);
    CHECK(ok, "%s: plan failed", tag);
    if (!ok) return;

    // all images, reference style
    std::vector<Img> all;
    for (int vz = -1; vz <= 1; ++vz)
        for (int vy = -1; vy <= 1; ++vy)
            for (int vx = -1; vx <= 1; ++vx) {
                if (dim == 2 && vz != 0) continue;
                int v[3] = {vx, vy, vz};
                for (int i = 0; i < n; ++i) {
                    Img im; bool keep = true;
                    im.ghost = (vx != 0 || vy != 0 || vz != 0);
                    for (int d = 0; d < 3; ++d) {
                        float b = (float)tgt[3 * (size_t)i + d];
                        im.x[d] = b + (float)v[d] * P.shift[d];
                        if (!((double)im.x[d] < P.win_hi[d]) || !((double)im.x[d] > P.win_lo[d])) keep = false;
                    }
                    // the placement kernels decide with one closed float32 interval per dimension
                    for (int d = 0; d < 3; ++d) {
                        const double x = (double)im.x[d];
                        const bool four = x < P.win_hi[d] && x > P.win_lo[d] && x >= P.reg_lo[d] && x <= P.reg_hi[d];
                        const bool one = im.x[d] >= P.keep_lo[d] && im.x[d] <= P.keep_hi[d];
                        CHECK(four == one, "%s: keep interval disagrees for image coordinate %.9g (dim %d)", tag, x, d);
                    }
                    if (keep) all.push_back(im);
                }
            }
    // brute force histogram (float64 compare of widened sqrt)
    std::vector<long long> want(nb, 0), got(nb, 0);
    auto dist = [&](const QT* c, const Img& im) -> double {
        if (sizeof(QT) == 4) {
            float dx = (float)c[0] - im.x[0], dy = (float)c[1] - im.x[1], dz = (float)c[2] - im.x[2];
            float sx = dx * dx, sy = dy * dy, sz = dz * dz;
            float s = sx + sy; s = s + sz;
            return (double)sqrtf(s);
        }
        double dx = (double)c[0] - (double)im.x[0], dy = (double)c[1] - (double)im.x[1], dz = (double)c[2] - (double)im.x[2];
        double sx = dx * dx, sy = dy * dy, sz = dz * dz;
        double s = sx + sy; s = s + sz;
        return sqrt(s);
    };
    for (int q = 0; q < m; ++q)
        for (const Img& im : all) {
            int b = numpy_bin(dist(&qry[3 * (size_t)q], im), e);
            if (b >= 0) want[b]++;
        }

    // grid walk
    std::vector<std::vector<int>> tcell(P.ncells), qcell(P.ncells);
    std::vector<Img> kept;
    for (const Img& im : all) {
        bool inreg = true; int c[3];
        for (int d = 0; d < 3; ++d) {
            double x = (double)im.x[d];
            if (!(x >= P.reg_lo[d] && x <= P.reg_hi[d])) inreg = false;
            c[d] = cell_of(x, P.g0[d], P.inv_h[d], P.n[d]);
        }
        if (!inreg) continue;
        tcell[(c[2] * P.n[1] + c[1]) * P.n[0] + c[0]].push_back((int)kept.size());
        kept.push_back(im);
    }
    for (int q = 0; q < m; ++q) {
        int c[3];
        for (int d = 0; d < 3; ++d) c[d] = cell_of((double)qry[3 * (size_t)q + d], P.g0[d], P.inv_h[d], P.n[d]);
        for (int d = 0; d < 3; ++d)
            CHECK(c[d] >= P.cq_lo[d] && c[d] < P.cq_lo[d] + P.cq_n[d], "%s: query cell outside range", tag);
        qcell[(c[2] * P.n[1] + c[1]) * P.n[0] + c[0]].push_back(q);
    }
    long long ntask_seen = 0;
    for (long long t = 0; t < P.ntasks; ++t) {
        int run_i = (int)(t % P.runs_per_row);
        long long rowq = t / P.runs_per_row;
        int cy = P.cq_lo[1] + (int)(rowq % P.cq_n[1]);
        int cz = P.cq_lo[2] + (int)(rowq / P.cq_n[1]);
        int cx0 = P.cq_lo[0] + run_i * P.run;
        int ncell = std::min(P.run, P.cq_lo[0] + P.cq_n[0] - cx0);
        ++ntask_seen;
        for (int wt = 0; wt < ncell * P.nrows; ++wt) {
            int c = wt % ncell, r = wt / ncell, cx = cx0 + c;
            int ny = cy + P.row_oy[r], nz = cz + P.row_oz[r];
            if (ny < 0 || ny >= P.n[1] || nz < 0 || nz >= P.n[2]) continue;
            int kx = P.row_kx[r];
            int x_lo = std::max(cx - kx, 0), x_hi = std::min(cx + kx, P.n[0] - 1);
            for (int q : qcell[(cz * P.n[1] + cy) * P.n[0] + cx])
                for (int x = x_lo; x <= x_hi; ++x)
                    for (int j : tcell[(nz * P.n[1] + ny) * P.n[0] + x]) {
                        int b = numpy_bin(dist(&qry[3 * (size_t)q], kept[j]), e);
                        if (b >= 0) got[b]++;
                    }
        }
    }
    long long sw = 0, sg = 0;
    for (int b = 0; b < nb; ++b) {
        sw += want[b]; sg += got[b];
        CHECK(want[b] == got[b], "%s: bin %d want %lld got %lld", tag, b, want[b], got[b]);
    }
    // symmetric mode (self RDF, float32): upper half shell x2, own row split, ghosts separately,
    // ghost-free neighbourhoods skipped -- the walk of pair_kernel<..., SYM=true>
    if (self && sizeof(QT) == 4) {
        bool inside = true;
        for (int d = 0; d < 3; ++d)
            if (!((double)(float)ts.lo[d] > P.win_lo[d]) || !((double)(float)ts.hi[d] < P.win_hi[d])) inside = false;
        if (inside) {
            std::vector<long long> sym(nb, 0);
            std::vector<std::vector<int>> gcell(P.ncells);
            for (size_t j = 0; j < kept.size(); ++j) {
                if (!kept[j].ghost) continue;
                int c[3];
                for (int d = 0; d < 3; ++d) c[d] = cell_of((double)kept[j].x[d], P.g0[d], P.inv_h[d], P.n[d]);
                gcell[(c[2] * P.n[1] + c[1]) * P.n[0] + c[0]].push_back((int)j);
            }
            auto qdist = [&](int q, int q2) {
                Img im;
                for (int d = 0; d < 3; ++d) im.x[d] = (float)qry[3 * (size_t)q2 + d];
                return dist(&qry[3 * (size_t)q], im);
            };
            for (long long t = 0; t < P.ntasks; ++t) {
                int run_i = (int)(t % P.runs_per_row);
                long long rowq = t / P.runs_per_row;
                int cy = P.cq_lo[1] + (int)(rowq % P.cq_n[1]);
                int cz = P.cq_lo[2] + (int)(rowq / P.cq_n[1]);
                int cx0 = P.cq_lo[0] + run_i * P.run;
                int ncell = std::min(P.run, P.cq_lo[0] + P.cq_n[0] - cx0);
                for (int wtk = 0; wtk < ncell * P.rsplit; ++wtk) {
                    int part = wtk / ncell, cx = cx0 + (wtk - part * ncell);
                    const std::vector<int>& qs = qcell[(cz * P.n[1] + cy) * P.n[0] + cx];
                    for (int pass = 0; pass < 2; ++pass) {
                        bool real = pass == 0;
                        int r_begin = real ? P.row0 : 0;
                        if (!real && cx - P.kd[0] >= P.gfree_lo[0] && cx + P.kd[0] <= P.gfree_hi[0] &&
                            cy - P.kd[1] >= P.gfree_lo[1] && cy + P.kd[1] <= P.gfree_hi[1] &&
                            cz - P.kd[2] >= P.gfree_lo[2] && cz + P.kd[2] <= P.gfree_hi[2])
                            break;
                        for (int r = r_begin + part; r < P.nrows; r += P.rsplit) {
                            int ny = cy + P.row_oy[r], nz = cz + P.row_oz[r];
                            if (ny < 0 || ny >= P.n[1] || nz < 0 || nz >= P.n[2]) continue;
                            int kx = P.row_kx[r];
                            int x_lo = std::max(cx - kx, 0), x_hi = std::min(cx + kx, P.n[0] - 1);
                            if (real && r == P.row0) x_lo = cx;
                            for (int x = x_lo; x <= x_hi; ++x) {
                                int cellid = (nz * P.n[1] + ny) * P.n[0] + x;
                                long long w = real ? ((r == P.row0 && x == cx) ? 1 : 2) : 1;
                                if (real) {
                                    for (int q : qs)
                                        for (int q2 : qcell[cellid]) {
                                            int b = numpy_bin(qdist(q, q2), e);
                                            if (b >= 0) sym[b] += w;
                                        }
                                } else {
                                    for (int q : qs)
                                        for (int j : gcell[cellid]) {
                                            int b = numpy_bin(dist(&qry[3 * (size_t)q], kept[j]), e);
                                            if (b >= 0) sym[b] += w;
                                        }
                                }
                            }
                        }
                    }
                }
            }
            for (int b = 0; b < nb; ++b)
                CHECK(want[b] == sym[b], "%s: symmetric walk bin %d want %lld got %lld", tag, b, want[b], sym[b]);
        }
    }
    printf("  %-28s cells=%dx%dx%d rows=%d run=%d tasks=%lld images=%zu/%zu pairs=%lld\n", tag, P.n[0], P.n[1],
           P.n[2], P.nrows, P.run, P.ntasks, kept.size(), all.size(), sw);
    CHECK(sw > 0, "%s: empty histogram", tag);
    (void)sg;
}

// float_above / float_below: the float32 neighbours of a float64 bound reproduce the float64
// comparison for every float32 value (checked on the floats around the bound)
static void check_float_bounds(std::mt19937_64& rng) {
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const double scales[] = {1e-30, 1e-3, 1.0, 37.5, 1253.3336, 1e6, 1e30};
    for (double sc : scales)
        for (int it = 0; it < 2000; ++it) {
            double b = U(rng) * sc;
            if (it % 7 == 0) b = (double)(float)b;          // exactly representable bounds too
            const float ge = float_above(b, false), gt = float_above(b, true);
            const float le = float_below(b, false), lt = float_below(b, true);
            float x = (float)b;
            for (int step = 0; step < 3; ++step) x = nextafterf(x, -INFINITY);
            for (int step = 0; step < 7; ++step, x = nextafterf(x, INFINITY)) {
                CHECK(((double)x >= b) == (x >= ge), "float_above(%.17g, false) wrong at %.9g", b, (double)x);
                CHECK(((double)x > b) == (x >= gt), "float_above(%.17g, true) wrong at %.9g", b, (double)x);
                CHECK(((double)x <= b) == (x <= le), "float_below(%.17g, false) wrong at %.9g", b, (double)x);
                CHECK(((double)x < b) == (x <= lt), "float_below(%.17g, true) wrong at %.9g", b, (double)x);
            }
        }
    const double inf = std::numeric_limits<double>::infinity();
    CHECK(float_above(-inf, true) == -std::numeric_limits<float>::max(), "float_above(-inf, strict)");
    CHECK(float_below(inf, true) == std::numeric_limits<float>::max(), "float_below(+inf, strict)");
    CHECK(float_above(-inf, false) == -std::numeric_limits<float>::infinity(), "float_above(-inf)");
}

int main() {
    std::mt19937_64 rng(12345);
    printf("thresholds\n");
    check_float_bounds(rng);
    check_thresholds<float>(rng, true, 32.0, 500);
    check_thresholds<float>(rng, false, 10.0, 500);
    check_thresholds<float>(rng, false, 0.05, 500);
    check_thresholds<double>(rng, false, 10.0, 500);
    check_thresholds<double>(rng, true, 5.5, 137);
    check_thresholds<double>(rng, false, 0.01, 64);
    // window of pbc_points for a float32 and a float64 box
    {
        double bb[6] = {0.0, 79.266, 0.0, 79.266, -0.5, 0.5};
        double lo[3], hi[3]; float l32[3];
        image_window(bb, 1, lo, hi, l32);
        float L = (float)79.266 - 0.0f;
        CHECK(l32[0] == L && hi[0] == (double)(((float)79.266 + 0.0f) / 2.0f + L), "window f32");
        image_window(bb, 0, lo, hi, l32);
        CHECK(hi[0] == 79.266 / 2.0 + (double)(float)79.266, "window f64");
    }
    printf("plans\n");
    double sq2[6] = {0, 64, 0, 64, -0.5, 0.5};
    check_plan<float>(rng, 2, true, sq2, 32.0, 50, 1200, 0, true, 0.0, 0, "2d f32 rmax=L/2");
    check_plan<float>(rng, 2, true, sq2, 6.0, 50, 3000, 0, true, 0.0, 0, "2d f32 rmax=6");
    check_plan<double>(rng, 2, false, sq2, 6.0, 50, 2000, 700, false, 0.15, 0, "2d f64 cross spill");
    double ns2[6] = {0, 30, -5, 15, -0.5, 0.5};
    check_plan<float>(rng, 2, true, ns2, 15.0, 40, 1000, 0, true, 0.1, 0, "2d f32 nonsquare spill");
    double cu3[6] = {0, 12, 0, 12, 0, 12};
    check_plan<float>(rng, 3, true, cu3, 2.0, 40, 2500, 0, true, 0.0, 0, "3d f32 rmax=2");
    check_plan<double>(rng, 3, false, cu3, 6.0, 40, 700, 0, true, 0.0, 0, "3d f64 rmax=L/2");
    double th3[6] = {-3, 7, 2, 10, 0, 2.5};
    check_plan<float>(rng, 3, true, th3, 4.0, 30, 1200, 500, false, 0.05, 0, "3d f32 thin cross");
    for (int k = 1; k <= 6; ++k) {
        char tag[64];
        snprintf(tag, sizeof(tag), "2d f32 k=%d", k);
        check_plan<float>(rng, 2, true, sq2, 5.0, 25, 2500, 0, true, 0.0, k, tag);
        snprintf(tag, sizeof(tag), "3d f32 k=%d", k);
        check_plan<float>(rng, 3, true, cu3, 2.5, 25, 1500, 0, true, 0.02, k, tag);
    }
    if (g_fail == 0) printf("ALL OK\n");
    else printf("%d FAILURES\n", g_fail);
    return g_fail ? 1 : 0;
}

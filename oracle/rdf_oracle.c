/*
 * rdf_oracle.c -- plain-C restatement of AMEP 1.3.0's rdf(mode='diff') pair
 * counting and of sfiso(mode='fast') / sf2d(mode='std') sums.
 *
 * TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs).  Nothing in amep_b200/ links or loads it.
 * Pinned against the reference's golden vectors in tests/test_oracle_golden.py
 * (through oracle/c_oracle.py).
 *
 * Follows, operation by operation (no FMA: build with -ffp-contract=off):
 *   amep/pbc.py:269-338          pbc_points(width=0.5, fold_coords=False)
 *   amep/spatialcor.py:374-388   __rdf_diff: d = sqrt((dx*dx+dy*dy)+dz*dz),
 *                                d > 1e-3, explicit-edge histogram
 *   amep/spatialcor.py:1440-1446 __sf_iso_chunk_fast (float64)
 *   amep/spatialcor.py:1715-1719 __sf2d_chunk_std (complex64 running sum)
 * Brute force over all query-image pairs, like the reference: O(M * N_img).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const int SHIFTS_2D[9][3] = {{0, 0, 0}, {-1, -1, 0}, {-1, 0, 0}, {-1, 1, 0}, {0, -1, 0},
                                    {0, 1, 0},  {1, -1, 0},  {1, 0, 0},  {1, 1, 0}};

/* images as float32 triples; returns the number kept.  out must hold 27*n*3 floats. */
long oracle_images(const void* coords, int is_f64, long n, const double* bb, int box_is_f32, int dim,
                   float* out) {
    float len32[3];
    double lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        if (box_is_f32) {
            float a = (float)bb[2 * d], b = (float)bb[2 * d + 1];
            float len = b - a, centre = (a + b) / 2.0f;
            len32[d] = len;
            hi[d] = (double)(centre + len * 1.0f);
            lo[d] = (double)(centre - len * 1.0f);
        } else {
            double a = bb[2 * d], b = bb[2 * d + 1];
            double centre = (a + b) / 2.0;
            len32[d] = (float)(b - a);
            hi[d] = centre + (double)(len32[d] * 1.0f);
            lo[d] = centre - (double)(len32[d] * 1.0f);
        }
    }
    long kept = 0;
    const int nshift = dim == 2 ? 9 : 27;
    for (int s = 0; s < nshift; ++s) {
        int v[3];
        if (dim == 2) {
            v[0] = SHIFTS_2D[s][0]; v[1] = SHIFTS_2D[s][1]; v[2] = 0;
        } else {
            v[0] = s / 9 - 1; v[1] = (s / 3) % 3 - 1; v[2] = s % 3 - 1;
        }
        for (long i = 0; i < n; ++i) {
            float img[3];
            int ok = 1;
            for (int d = 0; d < 3; ++d) {
                float base = is_f64 ? (float)((const double*)coords)[3 * i + d] : ((const float*)coords)[3 * i + d];
                float shift = (float)v[d] * len32[d];
                img[d] = base + shift;
                if (!((double)img[d] < hi[d]) || !((double)img[d] > lo[d])) ok = 0;
            }
            if (ok) {
                out[3 * kept] = img[0]; out[3 * kept + 1] = img[1]; out[3 * kept + 2] = img[2];
                ++kept;
            }
        }
    }
    return kept;
}

static inline void bin_distance(double d, const double* edges, int nbins, long long* hist) {
    if (!(d > 1e-3)) return;
    if (!(d >= edges[0]) || !(d <= edges[nbins])) return;
    /* largest i with edges[i] <= d; the last bin is closed on the right */
    int a = 0, b = nbins;
    while (b - a > 1) {
        int mid = (a + b) / 2;
        if (edges[mid] <= d) a = mid; else b = mid;
    }
    hist[a] += 1;
}

/*
 * Pair histogram.  targets: float32 images (targets_f64 == 0) or the float64
 * coordinates themselves (no pbc, float64 input).  queries in their own dtype.
 * q_begin/q_end select a slice of the queries (for bounded timing samples).
 */
void oracle_rdf_hist(const void* targets, int targets_f64, long nt, const void* queries, int queries_f64,
                     long q_begin, long q_end, const double* edges, int nbins, long long* hist,
                     int nthreads) {
    memset(hist, 0, sizeof(long long) * (size_t)nbins);
    const int f64 = targets_f64 || queries_f64;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        long long* local = (long long*)calloc((size_t)nbins, sizeof(long long));
#pragma omp for schedule(dynamic, 16)
        for (long q = q_begin; q < q_end; ++q) {
            if (!f64) {
                const float* c = (const float*)queries + 3 * q;
                const float* t = (const float*)targets;
                for (long j = 0; j < nt; ++j) {
                    float dx = c[0] - t[3 * j], dy = c[1] - t[3 * j + 1], dz = c[2] - t[3 * j + 2];
                    float sx = dx * dx, sy = dy * dy, sz = dz * dz;
                    float s = sx + sy;
                    s = s + sz;
                    bin_distance((double)sqrtf(s), edges, nbins, local);
                }
            } else {
                double c[3];
                for (int d = 0; d < 3; ++d)
                    c[d] = queries_f64 ? ((const double*)queries)[3 * q + d] : (double)((const float*)queries)[3 * q + d];
                for (long j = 0; j < nt; ++j) {
                    double t0, t1, t2;
                    if (targets_f64) {
                        const double* t = (const double*)targets + 3 * j;
                        t0 = t[0]; t1 = t[1]; t2 = t[2];
                    } else {
                        const float* t = (const float*)targets + 3 * j;
                        t0 = (double)t[0]; t1 = (double)t[1]; t2 = (double)t[2];
                    }
                    double dx = c[0] - t0, dy = c[1] - t1, dz = c[2] - t2;
                    double sx = dx * dx, sy = dy * dy, sz = dz * dz;
                    double s = sx + sy;
                    s = s + sz;
                    bin_distance(sqrt(s), edges, nbins, local);
                }
            }
        }
#pragma omp critical
        for (int b = 0; b < nbins; ++b) hist[b] += local[b];
        free(local);
    }
}

/* sums of sfiso(mode='fast') for one direction u: out[2*i] = sum cos(q_i p), out[2*i+1] = sum sin(q_i p),
 * p = (ux x + uy y) + uz z in float64.  (The reference uses NumPy's pairwise summation; plain
 * summation here differs from it at the 1e-13 level, far below the 1e-6 criterion.) */
void oracle_sfiso_sums(const void* coords, int is_f64, long n, const double* u, const double* q, int nq,
                       double* out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        double c = 0.0, s = 0.0;
        for (long j = 0; j < n; ++j) {
            double x, y, z;
            if (is_f64) {
                const double* p = (const double*)coords + 3 * j;
                x = p[0]; y = p[1]; z = p[2];
            } else {
                const float* p = (const float*)coords + 3 * j;
                x = p[0]; y = p[1]; z = p[2];
            }
            double proj = (u[0] * x + u[1] * y) + u[2] * z;
            double ph = q[i] * proj;
            c += cos(ph);
            s += sin(ph);
        }
        out[2 * i] = c;
        out[2 * i + 1] = s;
    }
}

/* rho(q) of sf2d(mode='std') for grid points [g_begin, g_end) of the flattened (ny, nx) grid:
 * complex64 accumulator per chunk, chunk sums added in order in complex64. out: float32 pairs. */
void oracle_sf2d_modes(const void* coords, int is_f64, long n, const double* qx, const double* qy,
                       long g_begin, long g_end, long chunksize, float* out, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long g = g_begin; g < g_end; ++g) {
        float tr = 0.0f, ti = 0.0f;
        for (long start = 0; start < n; start += chunksize) {
            long stop = start + chunksize < n ? start + chunksize : n;
            float ar = 0.0f, ai = 0.0f;
            for (long j = start; j < stop; ++j) {
                double x, y;
                if (is_f64) {
                    x = ((const double*)coords)[3 * j]; y = ((const double*)coords)[3 * j + 1];
                } else {
                    x = ((const float*)coords)[3 * j]; y = ((const float*)coords)[3 * j + 1];
                }
                double a = qx[g] * x, b = qy[g] * y;
                double ph = a + b;
                ar = (float)((double)ar + cos(ph));
                ai = (float)((double)ai - sin(ph));
            }
            tr = tr + ar;
            ti = ti + ai;
        }
        out[2 * (g - g_begin)] = tr;
        out[2 * (g - g_begin) + 1] = ti;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

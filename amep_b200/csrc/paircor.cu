// paircor.cu -- the remaining pair observables of amep.spatialcor on sm_100a.
//
//   amepb_pcf_angle_counts  <- amep/spatialcor.py:929-1041  __dhist_angle (+ the chunk sum of pcf_angle, :1268-1304)
//       For every ordered pair (query n of `other`, target i of `coords`): minimum-image difference
//       vector by amep.pbc.fold in NumPy's promoted dtype (amep/pbc.py:155-163, 597-602), its norm
//       in that dtype, the angle to the reference direction e through a float64 arccos, optional
//       rotation by the psi angle, np.histogram2d binning of (distance, angle).
//   amepb_spatialcor_sums   <- amep/spatialcor.py:47-135    __scor_chunk (+ the chunk sum of spatialcor, :322-330)
//       Periodic-KD-tree distances restated: tree points shifted and folded into [0, L)
//       (amep/pbc.py:632-655), query points shifted, folded about the box centre
//       (amep/spatialcor.py:315-319) and wrapped into [0, L) as scipy's periodic tree does, float64
//       minimum-image distances, pairs below max(box)/2 only; per distance bin the pair count and the
//       sum of conj(other_value) * value.
//
// Both are all-pairs kernels (the reference is O(N M) as well): one thread per query, targets
// staged through shared memory as broadcasts.  No fused multiply-add in any value that is
// compared against a bin edge.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "amepb_internal.cuh"
#include "rdf_plan.h"

namespace amepb {

template <typename R>
struct Rn;
template <>
struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float root(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float nearest(float a) { return rintf(a); }
};
template <>
struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double root(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double nearest(double a) { return rint(a); }
};

// np.histogramdd along one axis (numpy/lib/_histograms_impl.py): bin = searchsorted(edges, v,
// 'right') - 1, a value equal to the last edge belongs to the last bin; -1 = not counted.
__device__ __forceinline__ int edge_bin(double v, const double* __restrict__ e, int ne, double inv_w, bool uniform) {
    if (!(v >= e[0]) || !(v <= e[ne - 1])) return -1;   // also NaN
    if (v == e[ne - 1]) return ne - 2;
    int b;
    if (uniform) {
        b = min(max((int)((v - e[0]) * inv_w), 0), ne - 2);
        while (b > 0 && v < e[b]) --b;
        while (b < ne - 2 && v >= e[b + 1]) ++b;
    } else {
        int lo = 0, hi = ne - 1;   // e[lo] <= v < e[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (v >= e[mid]) lo = mid;
            else hi = mid;
        }
        b = lo;
    }
    return b;
}

// ------------------------------------------------------------------------------------------
// pcf_angle
// ------------------------------------------------------------------------------------------
struct PcfaArgs {
    double fbox[3], fcen[3];   // box lengths / centre of the centred box fold() works with, widened
    double angle;              // rotation angle (0: none)
    double inv_wd, inv_wa;     // 1 / bin width of the two edge arrays (uniform edges)
    double e[3], e_norm;       // one reference direction (e_stride == 0)
    int e_stride;              // 3: one direction per query particle
    int pbc, same;
    long long self_off;        // >= 0: query q is particle q + self_off of the targets and is not paired with it
    int nde, nae;              // number of distance / angle edges
    int uniform_d, uniform_a;
    float cert_margin;         // > 0: angle bins may be decided in float32 (in units of a bin), see the kernel
    float pre_r2;              // > 0: squared distance beyond which the float32 pre-test drops a pair
};

constexpr int kPcfaThreads = 128;
constexpr int kPcfaTile = 512;

// TC / TO: dtypes of coords / other; P: dtype the difference vector ends up in -- float only if
// both arrays and (with pbc) the box are float32.
// PRIV: the (r, theta) grid fits into shared memory as 32-bit counters (the default 500 x 100 grid:
// 200 KB): one block of 1024 queries per SM counts there and adds its non-zero bins to the global grid
// once.  With the default rmax = L/2 three pairs in four are binned, and one 64-bit global atomic per
// binned pair bounded the first version of this kernel (5.8e10 pairs/s at N = 100k).
constexpr int kPcfaThreadsPriv = 1024;
template <typename TC, typename TO, typename P, int THREADS, bool PRIV>
__global__ void __launch_bounds__(THREADS) pcf_angle_kernel(
    const TC* __restrict__ coords, long long n, const TO* __restrict__ other, long long m, PcfaArgs A,
    const double* __restrict__ edges /* distance edges, then angle edges */, const double* __restrict__ e_all,
    const double* __restrict__ e_norm_all, unsigned long long* __restrict__ counts) {
    // the difference coords - other is taken in NumPy's promotion of the two array dtypes
    using D = decltype(TC() - TO());
    extern __shared__ __align__(16) unsigned char smem_pcfa[];
    double* s_ed = reinterpret_cast<double*>(smem_pcfa);
    double* s_ea = s_ed + A.nde;
    TC* s_t = reinterpret_cast<TC*>(s_ea + A.nae);   // [tile][2]
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_t + 2 * kPcfaTile);   // PRIV: [nde - 1][nae - 1]
    const int nbins_total = (A.nde - 1) * (A.nae - 1);
    for (int k = threadIdx.x; k < A.nde + A.nae; k += blockDim.x) s_ed[k] = edges[k];
    if (PRIV)
        for (int k = threadIdx.x; k < nbins_total; k += blockDim.x) s_hist[k] = 0u;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have_q = q < m;
    D ox = 0, oy = 0;
    double e0 = A.e[0], e1 = A.e[1], e2 = A.e[2], en = A.e_norm;
    if (have_q) {
        ox = (D)other[3 * q];
        oy = (D)other[3 * q + 1];
        if (A.e_stride) {
            e0 = e_all[3 * q]; e1 = e_all[3 * q + 1]; e2 = e_all[3 * q + 2];
            en = e_norm_all[q];
        }
    }
    // the z component: both arrays are flat (z == 0, enforced by the caller like the reference), so
    // the folded z difference is one constant
    P vz = (P)0;
    if (A.pbc) {
        const P fb = (P)A.fbox[2], fc = (P)A.fcen[2];
        const P s = Rn<P>::div(Rn<P>::sub((P)0, fc), fb);
        vz = Rn<P>::add(Rn<P>::mul(fb, Rn<P>::sub(s, Rn<P>::nearest(s))), fc);
    }
    const P vz2 = Rn<P>::mul(vz, vz);
    const float fox = (float)ox, foy = (float)oy;
    const float plx = (float)A.fbox[0], ply = (float)A.fbox[1], pcx = (float)A.fcen[0], pcy = (float)A.fcen[1];
    const float pix = 1.0f / plx, piy = 1.0f / ply;
    const double two_pi = 6.283185307179586;   // 2*np.pi
    const long long tile0 = (long long)blockIdx.y * ((n + gridDim.y - 1) / gridDim.y);
    const long long tile1 = min(n, tile0 + (n + gridDim.y - 1) / gridDim.y);
    for (long long t0 = tile0; t0 < tile1; t0 += kPcfaTile) {
        const int cnt = (int)min((long long)kPcfaTile, tile1 - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            s_t[2 * k] = coords[3 * (t0 + k)];
            s_t[2 * k + 1] = coords[3 * (t0 + k) + 1];
        }
        __syncthreads();
        if (!have_q) continue;
        for (int k = 0; k < cnt; ++k) {
            if (A.self_off >= 0 && t0 + k == q + A.self_off) continue;   // coords[particlerange != n]
            if (A.pre_r2 > 0.f) {
                // Cheap float32 look at the pair first: minimum image by multiply / magic-number rounding /
                // fma (no division, no conversion -- the exact fold below keeps the XU pipe 71 % busy),
                // and a pair farther than the last distance edge by more than pre_r2's tolerance (8e-6 of
                // the box, far above the float32 rounding of either fold) cannot be counted.
                float ax = (float)s_t[2 * k] - fox, ay = (float)s_t[2 * k + 1] - foy;
                if (A.pbc) {
                    const float ux = ax - pcx, uy = ay - pcy;
                    const float rx = __fadd_rn(__fmaf_rn(ux, pix, 12582912.f), -12582912.f);
                    const float ry = __fadd_rn(__fmaf_rn(uy, piy, 12582912.f), -12582912.f);
                    ax = __fmaf_rn(-rx, plx, ux) + pcx;
                    ay = __fmaf_rn(-ry, ply, uy) + pcy;
                }
                if (ax * ax + ay * ay > A.pre_r2) continue;   // (NaN and huge values fall through to the exact test)
            }
            P vx = (P)Rn<D>::sub((D)s_t[2 * k], ox);
            P vy = (P)Rn<D>::sub((D)s_t[2 * k + 1], oy);
            if (A.pbc) {
                // fold(v, box - centre): s = (v - c) / L;  v = L * (s - round(s)) + c
                const P fbx = (P)A.fbox[0], fcx = (P)A.fcen[0], fby = (P)A.fbox[1], fcy = (P)A.fcen[1];
                const P sx = Rn<P>::div(Rn<P>::sub(vx, fcx), fbx);
                const P sy = Rn<P>::div(Rn<P>::sub(vy, fcy), fby);
                vx = Rn<P>::add(Rn<P>::mul(fbx, Rn<P>::sub(sx, Rn<P>::nearest(sx))), fcx);
                vy = Rn<P>::add(Rn<P>::mul(fby, Rn<P>::sub(sy, Rn<P>::nearest(sy))), fcy);
            }
            // np.linalg.norm(r_ij, axis=1): sqrt((x*x + y*y) + z*z) in the vector's dtype
            const P dist = Rn<P>::root(Rn<P>::add(Rn<P>::add(Rn<P>::mul(vx, vx), Rn<P>::mul(vy, vy)), vz2));
            if (!(dist > (P)0)) continue;
            const double dd = (double)dist;
            const int bd = edge_bin(dd, s_ed, A.nde, A.inv_wd, A.uniform_d != 0);
            if (bd < 0) continue;
            // Angle bin decided in float32 where that is certain (uniform angle edges from 0 to 2 pi).
            // theta~ = atan2(|e x r|, e.r), mirrored and rotated like the float64 value below, is within
            // 3e-6 rad of the *true* angle (products and sums 5e-7, atan2f 2 ulp = 5e-7, the float32
            // constants and the three additions 2.4e-7 each, the scaling 7.5e-7).  The reference's own
            // angle is arccos of a ratio that carries the float32 rounding of dist (2e-7 relative), so it
            // sits up to 2e-7 / sin(theta) from the true angle -- 3.5e-4 rad at the poles, where it
            // returns exactly 0 or pi for a whole cone of directions.  Hence: only pairs with
            // sin(theta~) >= 4e-3 (all but 0.25 %), and only a value farther than cert_margin = 6e-5 rad
            // (>= 3e-6 + 2e-7 / 4e-3) from both edges of its bin, are decided here; the rest -- one pair
            // in 250 for 100 angle bins -- takes the float64 arccos like the reference.
            if (A.cert_margin > 0.f) {
                const float fx = (float)vx, fy = (float)vy, fz = (float)vz;
                const float f0 = (float)e0, f1 = (float)e1, f2 = (float)e2;
                const float dotf = f0 * fx + f1 * fy + f2 * fz;
                const float cx = f1 * fz - f2 * fy, cy = f2 * fx - f0 * fz, cz = f0 * fy - f1 * fx;
                const float cr2 = cx * cx + cy * cy + cz * cz;
                float th = atan2f(sqrtf(cr2), dotf);
                if (vy < (P)0) th = 6.2831855f - th;
                if (A.angle != 0.0) {
                    th -= (float)A.angle;
                    if (th < 0.f) th += 6.2831855f;
                    if (th >= 6.2831855f) th -= 6.2831855f;
                }
                const float fb = th * (float)A.inv_wa;
                const float fl = floorf(fb);
                if (cr2 >= 1.6e-5f * (cr2 + dotf * dotf) && fb - fl > A.cert_margin && (fl + 1.0f) - fb > A.cert_margin &&
                    fl >= 0.f && fl < (float)(A.nae - 1)) {
                    const int bc = (int)fl;
                    if (PRIV) atomicAdd(&s_hist[bd * (A.nae - 1) + bc], 1u);
                    else atomicAdd(&counts[(long long)bd * (A.nae - 1) + bc], 1ull);
                    continue;
                }
            }
            // ratio = clip(dot(e, r) / (dist * |e|), -1, 1); theta = arccos(ratio) in float64
            const double dot = __dadd_rn(__dadd_rn(__dmul_rn(e0, (double)vx), __dmul_rn(e1, (double)vy)), __dmul_rn(e2, (double)vz));
            double ratio = __ddiv_rn(dot, __dmul_rn(dd, en));
            ratio = ratio < -1.0 ? -1.0 : (ratio > 1.0 ? 1.0 : ratio);   // NaN passes through, as in np.clip
            double theta = acos(ratio);
            if (vy < (P)0) theta = __dsub_rn(two_pi, theta);
            if (A.angle != 0.0) {
                theta = __dsub_rn(theta, A.angle);
                if (theta < 0.0) theta = __dadd_rn(theta, two_pi);
                if (theta >= two_pi) theta = __dsub_rn(theta, two_pi);
            }
            const int ba = edge_bin(theta, s_ea, A.nae, A.inv_wa, A.uniform_a != 0);
            if (ba < 0) continue;
            if (PRIV) atomicAdd(&s_hist[bd * (A.nae - 1) + ba], 1u);
            else atomicAdd(&counts[(long long)bd * (A.nae - 1) + ba], 1ull);
        }
    }
    if (PRIV) {
        __syncthreads();
        for (int k = threadIdx.x; k < nbins_total; k += blockDim.x) {
            const unsigned v = s_hist[k];
            if (v) atomicAdd(&counts[k], (unsigned long long)v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// spatialcor
// ------------------------------------------------------------------------------------------
struct ScorPrep {
    double box[3], center[3], half[3];   // L, centre, L/2 in the box dtype, widened
    int pbc;
};

// Positions as the periodic KD-tree sees them, float64.  TREE = true: kdtree() (amep/pbc.py:632-655);
// TREE = false: the query points of spatialcor (amep/spatialcor.py:315-319) + scipy's wrap into [0, L).
// T: array dtype, P: NumPy's promotion of the array and the box dtype.
template <typename T, typename P, bool TREE>
__global__ void scor_prep_kernel(const T* __restrict__ pts, long long n, ScorPrep S, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double r;
        if (!S.pbc) {
            r = (double)pts[3 * i + d];
        } else {
            const P L = (P)S.box[d], c = (P)S.center[d], h = (P)S.half[d];
            // coords + box/2. - center
            const P t = Rn<P>::sub(Rn<P>::add((P)pts[3 * i + d], h), c);
            // fold(t, box): the tree folds into [0, L] (box centre L/2), the queries about the box centre
            const P fc = TREE ? h : c;
            const P s = Rn<P>::div(Rn<P>::sub(t, fc), L);
            P f = Rn<P>::add(Rn<P>::mul(L, Rn<P>::sub(s, Rn<P>::nearest(s))), fc);
            if (TREE) {
                if (f == L) f = (P)0;   // coords[coords[:,d] == box[d], d] = 0
                r = (double)f;
            } else {
                // scipy's periodic tree wraps query points into [0, L) (float64)
                const double Ld = (double)L;
                double x = (double)f;
                x = x - floor(x / Ld) * Ld;
                while (x >= Ld) x -= Ld;
                while (x < 0.0) x += Ld;
                r = x;
            }
        }
        out[3 * i + d] = r;
    }
}

struct ScorArgs {
    double L[3], hb[3];    // box lengths and half lengths (float64 of the box dtype values)
    double ub;             // distance_upper_bound = max(box) / 2
    double inv_w, e0;
    double s_lo, s_hi;     // squared-distance window of a counted pair, see scor_pair_kernel
    int pbc, ne, uniform, ncomp, cplx, same;
    int copies;            // copies of the shared bins per block: one per warp, or 1
};

constexpr int kScorThreads = 128;
constexpr int kScorTile = 256;
constexpr int kScorMaxComp = 4;

// values as float64 (re, im) pairs: [particle][comp][2]
template <typename V>
__global__ void scor_values_kernel(const V* __restrict__ v, long long count, int cplx, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (cplx) {
        out[2 * i] = (double)v[2 * i];
        out[2 * i + 1] = (double)v[2 * i + 1];
    } else {
        out[2 * i] = (double)v[i];
        out[2 * i + 1] = 0.0;
    }
}

__global__ void __launch_bounds__(kScorThreads) scor_pair_kernel(
    const double* __restrict__ tpos, long long n, const double* __restrict__ qpos, long long m,
    const double* __restrict__ tval, const double* __restrict__ qval, ScorArgs A, const double* __restrict__ thr,
    unsigned long long* __restrict__ norm, double* __restrict__ cor) {
    // thr[k]: smallest squared distance whose float64 root reaches edges[k] (host, sq_threshold): the bin of
    // a pair is decided on s = dx^2 + dy^2 + dz^2 without the IEEE square root; A.s_hi: below it the pair
    // is inside the tree's distance_upper_bound and below the last edge.
    constexpr int kWarps = kScorThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_scor[];
    double* s_t = reinterpret_cast<double*>(smem_scor);        // thresholds [ne]
    const int nb = A.ne - 1;
    // [copies][nb]: one copy per warp when that fits (the float64 adds are compare-and-swap loops), else one
    const int ncopy = A.copies;
    double* s_re = s_t + A.ne;
    double* s_im = s_re + ncopy * nb;
    unsigned* s_n = reinterpret_cast<unsigned*>(s_im + ncopy * nb);
    double* s_pos = reinterpret_cast<double*>(s_n + ((ncopy * nb + 1) & ~1));  // [tile][3]   (even count: 8-byte alignment)
    double* s_val = s_pos + 3 * kScorTile;                                      // [tile][ncomp][2]
    for (int k = threadIdx.x; k < A.ne; k += blockDim.x) s_t[k] = thr[k];
    for (int k = threadIdx.x; k < ncopy * nb; k += blockDim.x) {
        s_re[k] = 0.0;
        s_im[k] = 0.0;
        s_n[k] = 0u;
    }
    const int wofs = (ncopy == kWarps ? (int)(threadIdx.x >> 5) : 0) * nb;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have_q = q < m;
    double qx = 0, qy = 0, qz = 0, qv[kScorMaxComp][2];
#pragma unroll
    for (int c = 0; c < kScorMaxComp; ++c) qv[c][0] = qv[c][1] = 0.0;
    if (have_q) {
        qx = qpos[3 * q]; qy = qpos[3 * q + 1]; qz = qpos[3 * q + 2];
#pragma unroll
        for (int c = 0; c < kScorMaxComp; ++c)
            if (c < A.ncomp) {
                qv[c][0] = qval[(q * A.ncomp + c) * 2];
                qv[c][1] = qval[(q * A.ncomp + c) * 2 + 1];
            }
    }
    // (box values in registers; without pbc the half lengths are infinite and no wrap happens)
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double Lx = A.L[0], Ly = A.L[1], Lz = A.L[2];
    const double hx = A.pbc ? A.hb[0] : inf, hy = A.pbc ? A.hb[1] : inf, hz = A.pbc ? A.hb[2] : inf;
    const double s_lo = A.s_lo, s_hi = A.s_hi;
    const float e0 = (float)A.e0, inv_w = (float)A.inv_w;
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    const long long tile0 = (long long)blockIdx.y * per, tile1 = min(n, tile0 + per);
    for (long long t0 = tile0; t0 < tile1; t0 += kScorTile) {
        const int cnt = (int)min((long long)kScorTile, tile1 - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += blockDim.x) s_pos[k] = tpos[3 * t0 + k];
        for (int k = threadIdx.x; k < cnt * A.ncomp * 2; k += blockDim.x) s_val[k] = tval[t0 * A.ncomp * 2 + k];
        __syncthreads();
        if (!have_q) continue;
        for (int k = 0; k < cnt; ++k) {
            double dx = __dsub_rn(qx, s_pos[3 * k]), dy = __dsub_rn(qy, s_pos[3 * k + 1]), dz = __dsub_rn(qz, s_pos[3 * k + 2]);
            // scipy BoxDist1D: one wrap by the box length where |d| exceeds half of it (selects, no branches)
            const double dxm = __dadd_rn(dx, Lx), dxp = __dsub_rn(dx, Lx);
            const double dym = __dadd_rn(dy, Ly), dyp = __dsub_rn(dy, Ly);
            const double dzm = __dadd_rn(dz, Lz), dzp = __dsub_rn(dz, Lz);
            dx = dx < -hx ? dxm : (dx > hx ? dxp : dx);
            dy = dy < -hy ? dym : (dy > hy ? dyp : dy);
            dz = dz < -hz ? dzm : (dz > hz ? dzp : dz);
            const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            // d < distance_upper_bound, d < edges[-1] (no right-inclusive last bin here), d >= edges[0]
            if (!(s < s_hi) || !(s >= s_lo)) continue;
            int b;
            if (A.uniform) {
                float r;
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)s));
                b = min(max((int)((r - e0) * inv_w), 0), nb - 1);
                while (b > 0 && s < s_t[b]) --b;
                while (b < nb - 1 && s >= s_t[b + 1]) ++b;
            } else {
                int lo = 0, hi = nb;   // s_t[lo] <= s < s_t[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s >= s_t[mid]) lo = mid;
                    else hi = mid;
                }
                b = lo;
            }
            // conj(other_value) * value, summed over the components
            double re = 0.0, im = 0.0;
            if (A.cplx) {
#pragma unroll
                for (int c = 0; c < kScorMaxComp; ++c)
                    if (c < A.ncomp) {
                        const double vr = s_val[(k * A.ncomp + c) * 2], vi = s_val[(k * A.ncomp + c) * 2 + 1];
                        re += qv[c][0] * vr + qv[c][1] * vi;
                        im += qv[c][0] * vi - qv[c][1] * vr;
                    }
            } else {   // real values: the imaginary parts are zero
#pragma unroll
                for (int c = 0; c < kScorMaxComp; ++c)
                    if (c < A.ncomp) re += qv[c][0] * s_val[(k * A.ncomp + c) * 2];
            }
            atomicAdd(&s_n[wofs + b], 1u);
            atomicAdd(&s_re[wofs + b], re);
            if (A.cplx) atomicAdd(&s_im[wofs + b], im);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
        unsigned cn = 0;
        double re = 0.0, im = 0.0;
        for (int w = 0; w < ncopy; ++w) {
            cn += s_n[w * nb + k];
            re += s_re[w * nb + k];
            im += s_im[w * nb + k];
        }
        if (cn) {
            atomicAdd(&norm[k], (unsigned long long)cn);
            atomicAdd(&cor[2 * k], re);
            if (A.cplx) atomicAdd(&cor[2 * k + 1], im);
        }
    }
}

static bool edges_uniform(const double* e, int ne, double* inv_w) {
    const double w = (e[ne - 1] - e[0]) / (ne - 1);
    *inv_w = w > 0 ? 1.0 / w : 0.0;
    if (!(w > 0)) return false;
    for (int i = 0; i < ne; ++i)
        if (!(fabs(e[i] - (e[0] + i * w)) <= 0.01 * w)) return false;
    return true;
}

// value of `x` after a round trip through the box dtype
static double in_box_dtype(double x, int box_dtype) { return box_dtype == AMEPB_F32 ? (double)(float)x : x; }

}  // namespace amepb

using namespace amepb;

extern "C" {

int amepb_pcf_angle_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other, int other_dtype,
                           int64_t m, int same, int space, const double* box_boundary, int box_dtype, int pbc,
                           const double* e, int64_t e_stride, const double* e_norm, double angle, const double* dedges,
                           int32_t nde, const double* aedges, int32_t nae, uint64_t* counts, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    auto ok_dtype = [](int d) { return d == AMEPB_F32 || d == AMEPB_F64; };
    if (!coords || !other || !box_boundary || !e || !e_norm || !dedges || !aedges || !counts || n <= 0 || m <= 0 ||
        nde < 2 || nae < 2 || (size_t)nde + nae > 5000 || !ok_dtype(dtype) || !ok_dtype(other_dtype) || !ok_dtype(box_dtype) ||
        (space != AMEPB_DEVICE && space != AMEPB_HOST) || (e_stride != 0 && e_stride != 3))
        return fail(ctx, AMEPB_E_INVAL, "amepb_pcf_angle_counts: invalid argument (at most 5000 bin edges in total)");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const size_t cb = (size_t)n * 3 * (dtype == AMEPB_F32 ? 4 : 8), ob = (size_t)m * 3 * (other_dtype == AMEPB_F32 ? 4 : 8);
    const size_t gbytes = sizeof(uint64_t) * (size_t)(nde - 1) * (nae - 1);
    int rc;
    const void* d_c = coords;
    const void* d_o = other;
    uint64_t* d_counts = counts;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, cb))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, cb, cudaMemcpyHostToDevice, stream));
        d_c = ctx->bufs[BUF_STAGE_A].p;
        if (other == coords) {
            d_o = d_c;
        } else {
            if ((rc = ensure_device(ctx, BUF_STAGE_B, ob))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_B].p, other, ob, cudaMemcpyHostToDevice, stream));
            d_o = ctx->bufs[BUF_STAGE_B].p;
        }
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, gbytes))) return rc;
        d_counts = reinterpret_cast<uint64_t*>(ctx->bufs[BUF_STAGE_OUT].p);
    }
    PcfaArgs A;
    for (int d = 0; d < 3; ++d) {
        // bbs = box_boundary - np.mean(box_boundary, axis=1)[:, None]; fold() then takes its lengths and centre
        const double lo = box_boundary[2 * d], hi = box_boundary[2 * d + 1];
        const double c = in_box_dtype(in_box_dtype(lo + hi, box_dtype) / 2.0, box_dtype);
        const double lo2 = in_box_dtype(lo - c, box_dtype), hi2 = in_box_dtype(hi - c, box_dtype);
        A.fbox[d] = in_box_dtype(hi2 - lo2, box_dtype);
        A.fcen[d] = in_box_dtype(in_box_dtype(lo2 + hi2, box_dtype) / 2.0, box_dtype);
    }
    A.angle = angle;
    A.uniform_d = edges_uniform(dedges, nde, &A.inv_wd) ? 1 : 0;
    A.uniform_a = edges_uniform(aedges, nae, &A.inv_wa) ? 1 : 0;
    // float32-certified angle bins: uniform edges from 0 to 2 pi (what pcf_angle always passes), margin 6e-5 rad
    const bool no_cert = getenv("AMEPB_PCFA_NO_F32_ANGLE") != nullptr;   // A/B switch for measurements and tests
    A.cert_margin = 0.f;
    if (A.uniform_a && !no_cert && aedges[0] == 0.0 && fabs(aedges[nae - 1] - 6.283185307179586) < 1e-9 &&
        6e-5 * A.inv_wa < 0.25)
        A.cert_margin = (float)(6e-5 * A.inv_wa);
    // float32 pre-test of the distance (see the kernel): last edge plus 8e-6 of the largest coordinate scale
    A.pre_r2 = 0.f;
    {
        const double scale = fabs(A.fbox[0]) + fabs(A.fbox[1]) + fabs(A.fcen[0]) + fabs(A.fcen[1]) + fabs(dedges[nde - 1]);
        const double r = dedges[nde - 1] + 8e-6 * scale;
        const bool no_pre = getenv("AMEPB_PCFA_NO_PRETEST") != nullptr;   // A/B switch for measurements and tests
        // (not worth its instructions when most pairs are in range, e.g. the default rmax = L/2 with periodic images)
        const bool mostly_in = pbc && 3.2 * r * r > 0.5 * fabs(A.fbox[0] * A.fbox[1]);
        if (!no_pre && !mostly_in && r > 0 && r < 1e15 && (!pbc || (A.fbox[0] > 0 && A.fbox[1] > 0 && scale < 1e15)))
            A.pre_r2 = (float)(r * r * (1.0 + 1e-6));
    }
    A.e[0] = e[0]; A.e[1] = e[1]; A.e[2] = e[2];
    A.e_norm = e_norm[0];
    A.e_stride = (int)e_stride;
    A.pbc = pbc ? 1 : 0;
    A.same = same ? 1 : 0;
    // same arrays: query q is target q; a slice of the targets passed as `other` (amepb_pcf_angle_set_query_offset)
    A.self_off = same ? 0 : (ctx->pcfa_query_offset >= 0 ? ctx->pcfa_query_offset : -1);
    A.nde = nde;
    A.nae = nae;
    // edge arrays (and per-particle directions) to the device
    const size_t ebytes = sizeof(double) * ((size_t)nde + nae);
    const size_t dirbytes = e_stride ? sizeof(double) * (size_t)m * 4 : 0;
    if ((rc = ensure_device(ctx, BUF_SQ_A, ebytes + dirbytes + 64))) return rc;
    if ((rc = ensure_pinned(ctx, ebytes))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;
    memcpy(ctx->pinned, dedges, sizeof(double) * nde);
    memcpy(reinterpret_cast<char*>(ctx->pinned) + sizeof(double) * nde, aedges, sizeof(double) * nae);
    char* dp = reinterpret_cast<char*>(ctx->bufs[BUF_SQ_A].p);
    AMEPB_CUDA(ctx, cudaMemcpyAsync(dp, ctx->pinned, ebytes, cudaMemcpyHostToDevice, stream));
    const double* d_edges = reinterpret_cast<const double*>(dp);
    const double* d_e = nullptr;
    const double* d_en = nullptr;
    if (e_stride) {
        char* de = dp + ((ebytes + 15) & ~(size_t)15);
        AMEPB_CUDA(ctx, cudaMemcpyAsync(de, e, sizeof(double) * (size_t)m * 3, cudaMemcpyHostToDevice, stream));
        AMEPB_CUDA(ctx, cudaMemcpyAsync(de + sizeof(double) * (size_t)m * 3, e_norm, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, stream));
        d_e = reinterpret_cast<const double*>(de);
        d_en = d_e + (size_t)m * 3;
    }
    AMEPB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, gbytes, stream));
    const bool all_f32 = dtype == AMEPB_F32 && other_dtype == AMEPB_F32;
    const bool p_f32 = all_f32 && (!pbc || box_dtype == AMEPB_F32);
    // shared-memory grid (see pcf_angle_kernel) when it fits next to the edges and the target tile
    const size_t base_smem = ebytes + (size_t)kPcfaTile * 2 * (dtype == AMEPB_F32 ? 4 : 8);
    const size_t priv_smem = base_smem + sizeof(unsigned) * (size_t)(nde - 1) * (nae - 1);
    const bool no_priv = getenv("AMEPB_PCFA_GLOBAL_ATOMICS") != nullptr;   // A/B switch for measurements and tests
    const bool priv = priv_smem <= 225 * 1024 && !no_priv;
    const int threads = priv ? kPcfaThreadsPriv : kPcfaThreads;
    const unsigned qblocks = (unsigned)((m + threads - 1) / threads);
    // split the targets so that a few thousand blocks are in flight (shared grid: two waves of one
    // block per SM, and at most 2^32 pairs per block for its 32-bit counters)
    const long long want_blocks = priv ? 2LL * ctx->sm_count : 8LL * ctx->sm_count;
    unsigned ysplit = (unsigned)std::max<long long>(1, std::min<long long>((n + kPcfaTile - 1) / kPcfaTile,
                                                                            (want_blocks + qblocks - 1) / qblocks));
    if (priv) ysplit = (unsigned)std::max<long long>(ysplit, ((long long)threads * n >> 31) + 1);
    ysplit = std::min(ysplit, 65535u);
    const dim3 grid(qblocks, ysplit);
    const size_t smem = priv ? priv_smem : base_smem;
#define AMEPB_PCFA(TC, TO, P)                                                                                                    \
    do {                                                                                                                         \
        if (priv) {                                                                                                              \
            AMEPB_CUDA(ctx, cudaFuncSetAttribute(pcf_angle_kernel<TC, TO, P, kPcfaThreadsPriv, true>,                            \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                       \
            pcf_angle_kernel<TC, TO, P, kPcfaThreadsPriv, true><<<grid, kPcfaThreadsPriv, smem, stream>>>(                       \
                (const TC*)d_c, n, (const TO*)d_o, m, A, d_edges, d_e, d_en, reinterpret_cast<unsigned long long*>(d_counts));   \
        } else {                                                                                                                 \
            pcf_angle_kernel<TC, TO, P, kPcfaThreads, false><<<grid, kPcfaThreads, smem, stream>>>(                              \
                (const TC*)d_c, n, (const TO*)d_o, m, A, d_edges, d_e, d_en, reinterpret_cast<unsigned long long*>(d_counts));   \
        }                                                                                                                        \
    } while (0)
    if (p_f32) AMEPB_PCFA(float, float, float);
    else if (all_f32) AMEPB_PCFA(float, float, double);
    else if (dtype == AMEPB_F32) AMEPB_PCFA(float, double, double);
    else if (other_dtype == AMEPB_F32) AMEPB_PCFA(double, float, double);
    else AMEPB_PCFA(double, double, double);
#undef AMEPB_PCFA
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) AMEPB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, gbytes, cudaMemcpyDeviceToHost, stream));
    // the pinned block and the direction arrays of the caller are reused after the call
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    return 0;
}

int amepb_pcf_angle_set_query_offset(amepb_ctx* ctx, int64_t offset) {
    if (!ctx) return AMEPB_E_INVAL;
    ctx->pcfa_query_offset = offset < 0 ? -1 : (long long)offset;
    return 0;
}

int amepb_spatialcor_sums(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other, int other_dtype,
                          int64_t m, int space, const double* box_boundary, int box_dtype, int pbc, const void* values,
                          const void* other_values, int value_kind, int32_t ncomp, const double* edges, int32_t nedges,
                          uint64_t* norm, double* cor, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    auto ok_dtype = [](int d) { return d == AMEPB_F32 || d == AMEPB_F64; };
    if (!coords || !box_boundary || !values || !edges || !norm || !cor || n <= 0 || nedges < 2 || nedges > 4096 ||
        ncomp < 1 || ncomp > kScorMaxComp || value_kind < 0 || value_kind > 3 || !ok_dtype(dtype) || !ok_dtype(box_dtype) ||
        (other && (!ok_dtype(other_dtype) || m <= 0 || !other_values)) || (space != AMEPB_DEVICE && space != AMEPB_HOST))
        return fail(ctx, AMEPB_E_INVAL, "amepb_spatialcor_sums: invalid argument (at most 4096 edges, %d value components)", kScorMaxComp);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const bool same = other == nullptr;
    if (same) {
        other = coords;
        other_dtype = dtype;
        m = n;
        other_values = values;
    }
    const bool cplx = value_kind >= 2;
    const size_t vsz = (value_kind == 0 || value_kind == 2) ? 4 : 8;
    const size_t cb = (size_t)n * 3 * (dtype == AMEPB_F32 ? 4 : 8), ob = (size_t)m * 3 * (other_dtype == AMEPB_F32 ? 4 : 8);
    const size_t vb = (size_t)n * ncomp * vsz * (cplx ? 2 : 1), wb = (size_t)m * ncomp * vsz * (cplx ? 2 : 1);
    int rc;
    const void *d_c = coords, *d_o = other, *d_v = values, *d_w = other_values;
    const int nb = nedges - 1;
    if (space == AMEPB_HOST) {
        // one staging block: coords | other | values | other_values
        const size_t a1 = (cb + 255) & ~(size_t)255, a2 = (ob + 255) & ~(size_t)255, a3 = (vb + 255) & ~(size_t)255;
        if ((rc = ensure_device(ctx, BUF_STAGE_A, a1 + a2 + a3 + wb + 256))) return rc;
        char* base = reinterpret_cast<char*>(ctx->bufs[BUF_STAGE_A].p);
        AMEPB_CUDA(ctx, cudaMemcpyAsync(base, coords, cb, cudaMemcpyHostToDevice, stream));
        AMEPB_CUDA(ctx, cudaMemcpyAsync(base + a1 + a2, values, vb, cudaMemcpyHostToDevice, stream));
        d_c = base;
        d_v = base + a1 + a2;
        if (same) {
            d_o = d_c;
            d_w = d_v;
        } else {
            AMEPB_CUDA(ctx, cudaMemcpyAsync(base + a1, other, ob, cudaMemcpyHostToDevice, stream));
            AMEPB_CUDA(ctx, cudaMemcpyAsync(base + a1 + a2 + a3, other_values, wb, cudaMemcpyHostToDevice, stream));
            d_o = base + a1;
            d_w = base + a1 + a2 + a3;
        }
    }
    // workspaces: float64 positions and values of both sets, edges, results
    const size_t pos_t = sizeof(double) * 3 * (size_t)n, pos_q = sizeof(double) * 3 * (size_t)m;
    const size_t val_t = sizeof(double) * 2 * (size_t)n * ncomp, val_q = sizeof(double) * 2 * (size_t)m * ncomp;
    if ((rc = ensure_device(ctx, BUF_SQ_B, pos_t + pos_q + val_t + val_q + 256))) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, sizeof(double) * nedges + (sizeof(uint64_t) + 2 * sizeof(double)) * nb + 64))) return rc;
    double* tpos = reinterpret_cast<double*>(ctx->bufs[BUF_SQ_B].p);
    double* qpos = tpos + 3 * (size_t)n;
    double* tval = qpos + 3 * (size_t)m;
    double* qval = tval + 2 * (size_t)n * ncomp;
    double* d_edges = reinterpret_cast<double*>(ctx->bufs[BUF_SQ_A].p);
    unsigned long long* d_norm = reinterpret_cast<unsigned long long*>(d_edges + nedges);
    double* d_cor = reinterpret_cast<double*>(d_norm + nb);
    if ((rc = ensure_pinned(ctx, sizeof(double) * nedges))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;
    // squared-distance thresholds of the edges: thr[k] = smallest s with sqrt(s) >= edges[k] (float64, IEEE root)
    {
        double* thr = reinterpret_cast<double*>(ctx->pinned);
        for (int k = 0; k < nedges; ++k) thr[k] = sq_threshold<double>(edges[k], false);
    }
    AMEPB_CUDA(ctx, cudaMemcpyAsync(d_edges, ctx->pinned, sizeof(double) * nedges, cudaMemcpyHostToDevice, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(d_norm, 0, (sizeof(uint64_t) + 2 * sizeof(double)) * nb, stream));
    ScorPrep S;
    ScorArgs A;
    double maxbox = -INFINITY;
    for (int d = 0; d < 3; ++d) {
        const double lo = box_boundary[2 * d], hi = box_boundary[2 * d + 1];
        S.box[d] = in_box_dtype(hi - lo, box_dtype);
        S.center[d] = in_box_dtype(in_box_dtype(lo + hi, box_dtype) / 2.0, box_dtype);
        S.half[d] = in_box_dtype(S.box[d] / 2.0, box_dtype);
        A.L[d] = S.box[d];
        A.hb[d] = 0.5 * S.box[d];
        maxbox = std::max(maxbox, S.box[d]);
    }
    S.pbc = pbc ? 1 : 0;
    A.ub = in_box_dtype(maxbox / 2.0, box_dtype);
    A.e0 = edges[0];
    A.s_lo = sq_threshold<double>(edges[0], false);
    A.s_hi = std::min(sq_threshold<double>(A.ub, false), sq_threshold<double>(edges[nedges - 1], false));
    A.pbc = S.pbc;
    A.ne = nedges;
    A.uniform = edges_uniform(edges, nedges, &A.inv_w) ? 1 : 0;
    A.ncomp = ncomp;
    A.cplx = cplx ? 1 : 0;
    A.same = same ? 1 : 0;
    const unsigned tb = (unsigned)((n + 255) / 256), qb = (unsigned)((m + 255) / 256);
    const bool box32 = box_dtype == AMEPB_F32;
    // tree points
    if (dtype == AMEPB_F32 && box32) scor_prep_kernel<float, float, true><<<tb, 256, 0, stream>>>((const float*)d_c, n, S, tpos);
    else if (dtype == AMEPB_F32) scor_prep_kernel<float, double, true><<<tb, 256, 0, stream>>>((const float*)d_c, n, S, tpos);
    else scor_prep_kernel<double, double, true><<<tb, 256, 0, stream>>>((const double*)d_c, n, S, tpos);
    AMEPB_LAUNCH_CHECK(ctx);
    // query points
    if (other_dtype == AMEPB_F32 && box32) scor_prep_kernel<float, float, false><<<qb, 256, 0, stream>>>((const float*)d_o, m, S, qpos);
    else if (other_dtype == AMEPB_F32) scor_prep_kernel<float, double, false><<<qb, 256, 0, stream>>>((const float*)d_o, m, S, qpos);
    else scor_prep_kernel<double, double, false><<<qb, 256, 0, stream>>>((const double*)d_o, m, S, qpos);
    AMEPB_LAUNCH_CHECK(ctx);
    const long long tcount = (long long)n * ncomp, qcount = (long long)m * ncomp;
    if (vsz == 4) {
        scor_values_kernel<float><<<(unsigned)((tcount + 255) / 256), 256, 0, stream>>>((const float*)d_v, tcount, A.cplx, tval);
        scor_values_kernel<float><<<(unsigned)((qcount + 255) / 256), 256, 0, stream>>>((const float*)d_w, qcount, A.cplx, qval);
    } else {
        scor_values_kernel<double><<<(unsigned)((tcount + 255) / 256), 256, 0, stream>>>((const double*)d_v, tcount, A.cplx, tval);
        scor_values_kernel<double><<<(unsigned)((qcount + 255) / 256), 256, 0, stream>>>((const double*)d_w, qcount, A.cplx, qval);
    }
    AMEPB_LAUNCH_CHECK(ctx);
    const unsigned qblocks = (unsigned)((m + kScorThreads - 1) / kScorThreads);
    unsigned ysplit = (unsigned)std::max<long long>(1, std::min<long long>((n + kScorTile - 1) / kScorTile,
                                                                            (8LL * ctx->sm_count + qblocks - 1) / qblocks));
    ysplit = std::min(ysplit, 65535u);
    int kw = kScorThreads / 32;
    if ((2 * sizeof(double) + sizeof(unsigned)) * (size_t)kw * nb > 96 * 1024) kw = 1;
    A.copies = kw;
    const size_t smem = sizeof(double) * nedges + 2 * sizeof(double) * kw * nb + sizeof(unsigned) * ((kw * nb + 1) & ~1) +
                        sizeof(double) * 3 * kScorTile + sizeof(double) * 2 * kScorTile * ncomp + 16;
    if (smem > 48 * 1024)
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(scor_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scor_pair_kernel<<<dim3(qblocks, ysplit), kScorThreads, smem, stream>>>(tpos, n, qpos, m, tval, qval, A, d_edges, d_norm, d_cor);
    AMEPB_LAUNCH_CHECK(ctx);
    const cudaMemcpyKind kind = space == AMEPB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    AMEPB_CUDA(ctx, cudaMemcpyAsync(norm, d_norm, sizeof(uint64_t) * nb, kind, stream));
    AMEPB_CUDA(ctx, cudaMemcpyAsync(cor, d_cor, 2 * sizeof(double) * nb, kind, stream));
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    return 0;
}

}  // extern "C"

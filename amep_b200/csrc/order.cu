// order.cu -- the psi_6 orientation prologue of evaluate.SF2d(rotate=True) / evaluate.PCF2d on sm_100a.
//
//   amepb_psi_k        <- amep/order.py:1183-1348  k_nearest_neighbors(k, pbc=True, enforce_nd=2)
//                         amep/order.py:1361-1501  psi_k
//       all nine float32 periodic images of every particle (pbc_points with its default
//       width=None: no window, amep/pbc.py:296-326) are counting-sorted into a cell grid; one
//       thread per particle searches rings of cells for its k+1 nearest images (float64 distances,
//       the unit of scipy's KDTree), drops the nearest (itself), folds the image numbers back with
//       `% N` and sums exp(i k theta) over the bonds to the *unshifted* neighbour positions, in the
//       dtype of the coordinates like NumPy (float32 -> complex64).
//   amepb_rotate_crop  <- amep/evaluate.py:1366-1377  in_box(rotate_coords(pbc_points(c, box), -theta, center), box)
//       (amep/utils.py:581-652, 655-742): every periodic image rotated about the box centre in
//       float64, the ones inside the closed box kept in the order of the image array
//       (ordered stream compaction: block counts, scan, write).
#include <math.h>
#include <float.h>

#include <algorithm>
#include <vector>

#include "amepb_internal.cuh"
#include "rdf_plan.h"

namespace amepb {

int frame_stats(amepb_ctx* ctx, const void* pts, int dtype, int64_t n, int nframes, std::vector<FrameStats>& out,
                cudaStream_t stream, int slot);

// shift vectors of pbc_points in the reference's order (amep/pbc.py:305-321)
__constant__ signed char c_shift2d[9][2] = {{0, 0}, {-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
__constant__ signed char c_shift3d[27][3] = {
    {0, 0, 0},   {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1}, {-1, 1, 0},
    {-1, 1, 1},  {0, -1, -1},  {0, -1, 0},  {0, -1, 1},  {0, 0, -1},  {0, 0, 1},  {0, 1, -1}, {0, 1, 0},   {0, 1, 1},
    {1, -1, -1}, {1, -1, 0},   {1, -1, 1},  {1, 0, -1},  {1, 0, 0},   {1, 0, 1},  {1, 1, -1}, {1, 1, 0},   {1, 1, 1}};

struct OrderGrid {
    double g0[2], inv_h[2];
    double hmin;       // smaller cell edge
    int n[2];
    float box32[3];    // float32 box lengths (amep/pbc.py:300)
};

__device__ __forceinline__ int grid_cell(double x, double g0, double inv_h, int n) {
    const int c = __double2int_rd((x - g0) * inv_h);  // NaN -> 0
    return min(max(c, 0), n - 1);
}

// image s of particle j: fl32(base + v * box) per component, float32 arithmetic (amep/pbc.py:300-326)
template <typename T>
__device__ __forceinline__ void image_2d(const T* __restrict__ p, int s, const float* box32, float& x, float& y, float& z) {
    x = __fadd_rn((float)p[0], __fmul_rn((float)c_shift2d[s][0], box32[0]));
    y = __fadd_rn((float)p[1], __fmul_rn((float)c_shift2d[s][1], box32[1]));
    z = __fadd_rn((float)p[2], __fmul_rn(0.0f, box32[2]));
}

template <typename T, bool FILL>
__global__ void __launch_bounds__(256) order_place_kernel(const T* __restrict__ coords, long long n, OrderGrid G,
                                                          unsigned* __restrict__ count, const unsigned* __restrict__ start,
                                                          float4* __restrict__ sorted) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 9 * n) return;
    const int s = (int)(e / n);
    const long long j = e - (long long)s * n;
    float x, y, z;
    image_2d(coords + 3 * j, s, G.box32, x, y, z);
    const int cx = grid_cell((double)x, G.g0[0], G.inv_h[0], G.n[0]);
    const int cy = grid_cell((double)y, G.g0[1], G.inv_h[1], G.n[1]);
    const unsigned c = (unsigned)(cy * G.n[0] + cx);
    if (!FILL) {
        atomicAdd(&count[c], 1u);
    } else {
        const unsigned slot = start[c] + atomicSub(&count[c], 1u) - 1u;
        sorted[slot] = make_float4(x, y, z, __int_as_float((int)e));
    }
}

constexpr int kOrderKMax = 8;   // k of psi_k this kernel supports (the evaluate classes use 6)

// NumPy's arithmetic for the bond sum in the coordinate dtype T (amep/order.py:1491-1499):
// theta = arctan2(dy, dx); term = exp(1j*k*theta) -> (cos(k theta), sin(k theta)); psi = 1/k * sum.
template <typename T>
struct BondSum;
template <>
struct BondSum<float> {
    float re = 0.f, im = 0.f;
    __device__ __forceinline__ void add(float dy, float dx, int k) {
        const float theta = (float)atan2((double)dy, (double)dx);   // float32 arctan2, correctly rounded
        const float arg = __fmul_rn((float)k, theta);                // complex64: (0 + k j) * theta
        double s, c;
        sincos((double)arg, &s, &c);
        re = __fadd_rn(re, (float)c);
        im = __fadd_rn(im, (float)s);
    }
    __device__ __forceinline__ void scale(int k) {
        const float w = (float)(1.0 / (double)k);   // the Python float 1./k takes the array's precision
        re = __fmul_rn(re, w);
        im = __fmul_rn(im, w);
    }
};
template <>
struct BondSum<double> {
    double re = 0.0, im = 0.0;
    __device__ __forceinline__ void add(double dy, double dx, int k) {
        const double theta = atan2(dy, dx);
        double s, c;
        sincos((double)k * theta, &s, &c);
        re += c;
        im += s;
    }
    __device__ __forceinline__ void scale(int k) {
        const double w = 1.0 / (double)k;
        re *= w;
        im *= w;
    }
};

template <typename T>
__global__ void __launch_bounds__(128) order_knn_psi_kernel(const T* __restrict__ coords, long long n, OrderGrid G,
                                                            const unsigned* __restrict__ start,
                                                            const float4* __restrict__ sorted, int k,
                                                            T* __restrict__ psi, int* __restrict__ neighbors,
                                                            unsigned long long* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qx = (double)coords[3 * i], qy = (double)coords[3 * i + 1], qz = (double)coords[3 * i + 2];
    const int cx = grid_cell(qx, G.g0[0], G.inv_h[0], G.n[0]);
    const int cy = grid_cell(qy, G.g0[1], G.inv_h[1], G.n[1]);
    // the k + 1 nearest images so far, ascending in (distance^2, image number)
    double bd[kOrderKMax + 1];
    int bi[kOrderKMax + 1];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int t = 0; t <= kOrderKMax; ++t) {
        bd[t] = inf;
        bi[t] = 0x7fffffff;
    }
    const int want = k + 1;
    auto visit = [&](int ax, int ay) {
        const unsigned c = (unsigned)(ay * G.n[0] + ax);
        const unsigned a = __ldg(&start[c]), b = __ldg(&start[c + 1]);
        for (unsigned t = a; t < b; ++t) {
            const float4 v = __ldg(&sorted[t]);
            const double dx = qx - (double)v.x, dy = qy - (double)v.y, dz = qz - (double)v.z;
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            const int id = __float_as_int(v.w);
            if (!(d2 < bd[kOrderKMax] || (d2 == bd[kOrderKMax] && id < bi[kOrderKMax]))) continue;   // (also NaN)
            // insertion into the sorted list (the list is short: unrolled bubble from the end)
            bd[kOrderKMax] = d2;
            bi[kOrderKMax] = id;
#pragma unroll
            for (int u = kOrderKMax; u > 0; --u) {
                const bool sw = bd[u] < bd[u - 1] || (bd[u] == bd[u - 1] && bi[u] < bi[u - 1]);
                if (sw) {
                    const double td = bd[u]; bd[u] = bd[u - 1]; bd[u - 1] = td;
                    const int ti = bi[u]; bi[u] = bi[u - 1]; bi[u - 1] = ti;
                }
            }
        }
    };
    const int rmax = max(G.n[0], G.n[1]);
    for (int r = 0; r <= rmax; ++r) {
        const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r;
        if (x0 < 0 && y0 < 0 && x1 >= G.n[0] && y1 >= G.n[1]) break;   // the ring lies outside the grid on every side
        if (r == 0) {
            visit(cx, cy);
        } else {
            for (int ax = max(x0, 0); ax <= min(x1, G.n[0] - 1); ++ax) {
                if (y0 >= 0) visit(ax, y0);
                if (y1 < G.n[1]) visit(ax, y1);
            }
            for (int ay = max(y0 + 1, 0); ay <= min(y1 - 1, G.n[1] - 1); ++ay) {
                if (x0 >= 0) visit(x0, ay);
                if (x1 < G.n[0]) visit(x1, ay);
            }
        }
        // every image not visited yet is at least r * hmin away (strict test: ties in distance are
        // broken by the image number, so an unvisited image at exactly this distance could still rank)
        const double reach = (double)r * G.hmin;
        double dsel = bd[kOrderKMax];     // distance of rank k + 1 (selected without a dynamic register index)
#pragma unroll
        for (int t = 0; t < kOrderKMax; ++t)
            if (t == want - 1) dsel = bd[t];
        if (dsel < reach * reach) break;
    }
    // ranks 2 .. k+1 (amep/order.py:1311: k = list(range(2, k+2)) drops the nearest, the particle itself)
    BondSum<T> sum;
    const T cix = coords[3 * i], ciy = coords[3 * i + 1];
    bool self_image = false;
#pragma unroll
    for (int t = 1; t <= kOrderKMax; ++t) {
        if (t > k) break;
        const int id = bi[t];
        long long j = (id == 0x7fffffff) ? -1 : (long long)(id % n);
        if (j < 0 || j == i) {   // fewer than k+1 images, or an image of the particle itself (amep/order.py:1340-1342)
            self_image = true;
            j = i;
        }
        if (neighbors) neighbors[i * k + (t - 1)] = (int)j;
        // bond to the unshifted position of particle j, in the dtype of the coordinates
        const T dx = coords[3 * j] - cix, dy = coords[3 * j + 1] - ciy;
        sum.add(dy, dx, k);
    }
    sum.scale(k);
    psi[2 * i] = sum.re;
    psi[2 * i + 1] = sum.im;
    if (self_image) atomicAdd(bad, 1ull);
}

// ------------------------------------------------------------------------------------------
// rotated periodic crop
// ------------------------------------------------------------------------------------------
struct CropArgs {
    double center[3];      // np.mean(box_boundary, axis=1) widened to float64
    double lo[3], hi[3];   // box boundary widened to float64
    double cs, sn;         // cosine and sine of the rotation angle (the caller's np.cos / np.sin values)
    float box32[3];
    float center32[3];
    int center_is_f32;     // the reference's `coords - center` is a float32 subtraction for a float32 box
    int dim;               // 2: 9 images, 3: 27 images
};

template <typename T>
__device__ __forceinline__ bool crop_point(const T* __restrict__ coords, long long n, long long e, const CropArgs& A,
                                           double* out3) {
    const int s = (int)(e / n);
    const long long j = e - (long long)s * n;
    const T* p = coords + 3 * j;
    float img[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float v = A.dim == 2 ? (d < 2 ? (float)c_shift2d[s][d] : 0.0f) : (float)c_shift3d[s][d];
        img[d] = __fadd_rn((float)p[d], __fmul_rn(v, A.box32[d]));
    }
    double u[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        u[d] = A.center_is_f32 ? (double)__fsub_rn(img[d], A.center32[d]) : __dsub_rn((double)img[d], A.center[d]);
    // np.dot(R, v) with R = rotation(-theta) about z (amep/utils.py:547-579, 643-647), float64
    const double rx = __dadd_rn(__dadd_rn(__dmul_rn(A.cs, u[0]), __dmul_rn(-A.sn, u[1])), __dmul_rn(0.0, u[2]));
    const double ry = __dadd_rn(__dadd_rn(__dmul_rn(A.sn, u[0]), __dmul_rn(A.cs, u[1])), __dmul_rn(0.0, u[2]));
    const double rz = __dadd_rn(__dadd_rn(__dmul_rn(0.0, u[0]), __dmul_rn(0.0, u[1])), u[2]);
    out3[0] = __dadd_rn(rx, A.center[0]);
    out3[1] = __dadd_rn(ry, A.center[1]);
    out3[2] = __dadd_rn(rz, A.center[2]);
    // in_box: closed interval in every dimension (amep/utils.py:727-733)
    return out3[0] >= A.lo[0] && out3[0] <= A.hi[0] && out3[1] >= A.lo[1] && out3[1] <= A.hi[1] &&
           out3[2] >= A.lo[2] && out3[2] <= A.hi[2];
}

// WRITE = false: kept images per block of 256 consecutive image numbers; WRITE = true: the kept
// images at block_start[block] + rank inside the block (ballot + popc: order preserved).
template <typename T, bool WRITE>
__global__ void __launch_bounds__(256) rotate_crop_kernel(const T* __restrict__ coords, long long n, long long total,
                                                          CropArgs A, unsigned* __restrict__ block_count,
                                                          const unsigned* __restrict__ block_start,
                                                          double* __restrict__ out, long long capacity) {
    __shared__ unsigned s_warp[8];
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double r[3] = {0.0, 0.0, 0.0};
    const bool keep = e < total && crop_point(coords, n, e, A, r);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    unsigned before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const unsigned c = s_warp[w];
        if (w < warp) before += c;
        all += c;
    }
    if (!WRITE) {
        if (threadIdx.x == 0) block_count[blockIdx.x] = all;
        return;
    }
    if (keep) {
        const long long slot = (long long)block_start[blockIdx.x] + before + __popc(m & ((1u << lane) - 1u));
        if (slot < capacity) {
            out[3 * slot] = r[0];
            out[3 * slot + 1] = r[1];
            out[3 * slot + 2] = r[2];
        }
    }
}

}  // namespace amepb

using namespace amepb;

extern "C" {

int amepb_psi_k(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space, const double* box_boundary,
                int box_dtype, int32_t k, void* psi_out, int32_t* neighbors_out, int64_t* bad_out, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !box_boundary || !psi_out || n <= 0 || 9 * n >= (1ll << 31) || k < 1 || k > kOrderKMax ||
        (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (box_dtype != AMEPB_F32 && box_dtype != AMEPB_F64) ||
        (space != AMEPB_DEVICE && space != AMEPB_HOST))
        return fail(ctx, AMEPB_E_INVAL, "amepb_psi_k: invalid argument (1 <= k <= %d, 9 n < 2^31)", kOrderKMax);
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const size_t esz = dtype == AMEPB_F32 ? 4 : 8;
    const size_t cbytes = (size_t)n * 3 * esz;
    int rc;
    const void* d_coords = coords;
    void* d_psi = psi_out;
    int32_t* d_nb = neighbors_out;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, cbytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, cbytes, cudaMemcpyHostToDevice, stream));
        d_coords = ctx->bufs[BUF_STAGE_A].p;
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, (size_t)n * 2 * esz + (size_t)n * k * sizeof(int32_t)))) return rc;
        d_psi = ctx->bufs[BUF_STAGE_OUT].p;
        d_nb = neighbors_out ? reinterpret_cast<int32_t*>((char*)d_psi + (size_t)n * 2 * esz) : nullptr;
    }
    // extent of the particles -> extent of all nine images
    std::vector<FrameStats> st;
    if ((rc = frame_stats(ctx, d_coords, dtype, n, 1, st, stream, 0))) return rc;
    OrderGrid G;
    double span[2], dom_lo[2], dom_hi[2];
    for (int d = 0; d < 3; ++d) {
        const double len = box_boundary[2 * d + 1] - box_boundary[2 * d];
        // box.astype(np.float32) of a float32 or float64 box (the difference of float32 values is float32)
        G.box32[d] = box_dtype == AMEPB_F32 ? (float)((float)box_boundary[2 * d + 1] - (float)box_boundary[2 * d]) : (float)len;
    }
    for (int d = 0; d < 2; ++d) {
        double lo = st[0].lo[d], hi = st[0].hi[d];
        if (!(lo <= hi)) lo = hi = 0.0;   // no finite coordinate
        const double L = fabs((double)G.box32[d]);
        span[d] = hi - lo;
        const double pad = 1e-3 * (fabs(lo) + fabs(hi) + L + 1.0);
        dom_lo[d] = lo - L - pad;
        dom_hi[d] = hi + L + pad;
    }
    // about three images per cell at the density of the particles' bounding box
    const double area = std::max(span[0], 1e-300) * std::max(span[1], 1e-300);
    double h = sqrt(3.0 * area / (double)n);
    if (!(h > 0) || !std::isfinite(h)) h = 1.0;
    const double max_cells = std::max(1024.0, 12.0 * (double)n);
    for (;;) {
        double total = 1;
        for (int d = 0; d < 2; ++d) {
            const double ext = dom_hi[d] - dom_lo[d];
            const double cells = std::min(1.0e6, std::max(1.0, ceil(ext / h)));
            G.n[d] = (int)cells;
            G.g0[d] = dom_lo[d];
            G.inv_h[d] = ext > 0 ? cells / ext : 0.0;
            total *= cells;
        }
        if (total <= max_cells) break;
        h *= 1.5;
    }
    {
        const double hx = (dom_hi[0] - dom_lo[0]) / G.n[0], hy = (dom_hi[1] - dom_lo[1]) / G.n[1];
        G.hmin = std::min(hx, hy);
    }
    const long long ncells = (long long)G.n[0] * G.n[1];
    const size_t cell_bytes = sizeof(unsigned) * (size_t)(ncells + 1);
    if ((rc = ensure_device(ctx, BUF_ORD_A, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_ORD_B, cell_bytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_ORD_C, sizeof(float4) * (size_t)9 * n))) return rc;
    if ((rc = ensure_device(ctx, BUF_MISC, 64))) return rc;
    unsigned* count = reinterpret_cast<unsigned*>(ctx->bufs[BUF_ORD_A].p);
    unsigned* start = reinterpret_cast<unsigned*>(ctx->bufs[BUF_ORD_B].p);
    float4* sorted = reinterpret_cast<float4*>(ctx->bufs[BUF_ORD_C].p);
    unsigned long long* bad = reinterpret_cast<unsigned long long*>(ctx->bufs[BUF_MISC].p) + 4;
    AMEPB_CUDA(ctx, cudaMemsetAsync(count, 0, cell_bytes, stream));
    AMEPB_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(unsigned long long), stream));
    const unsigned pgrid = (unsigned)((9 * n + 255) / 256);
    if (dtype == AMEPB_F32) order_place_kernel<float, false><<<pgrid, 256, 0, stream>>>((const float*)d_coords, n, G, count, start, sorted);
    else order_place_kernel<double, false><<<pgrid, 256, 0, stream>>>((const double*)d_coords, n, G, count, start, sorted);
    AMEPB_LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, count, start, ncells + 1, stream))) return rc;
    if (dtype == AMEPB_F32) order_place_kernel<float, true><<<pgrid, 256, 0, stream>>>((const float*)d_coords, n, G, count, start, sorted);
    else order_place_kernel<double, true><<<pgrid, 256, 0, stream>>>((const double*)d_coords, n, G, count, start, sorted);
    AMEPB_LAUNCH_CHECK(ctx);
    const unsigned kgrid = (unsigned)((n + 127) / 128);
    if (dtype == AMEPB_F32)
        order_knn_psi_kernel<float><<<kgrid, 128, 0, stream>>>((const float*)d_coords, n, G, start, sorted, k, (float*)d_psi, d_nb, bad);
    else
        order_knn_psi_kernel<double><<<kgrid, 128, 0, stream>>>((const double*)d_coords, n, G, start, sorted, k, (double*)d_psi, d_nb, bad);
    AMEPB_LAUNCH_CHECK(ctx);
    unsigned long long h_bad = 0;
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(psi_out, d_psi, (size_t)n * 2 * esz, cudaMemcpyDeviceToHost, stream));
        if (neighbors_out)
            AMEPB_CUDA(ctx, cudaMemcpyAsync(neighbors_out, d_nb, (size_t)n * k * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    }
    AMEPB_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad, sizeof(h_bad), cudaMemcpyDeviceToHost, stream));
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (bad_out) *bad_out = (int64_t)h_bad;
    return 0;
}

int amepb_rotate_crop(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space, const double* box_boundary,
                      int box_dtype, int dim, const double* rot, double* out, int64_t capacity, int64_t* count_out,
                      void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !box_boundary || !count_out || !rot || n <= 0 || (dim != 2 && dim != 3) || 27 * n >= (1ll << 40) ||
        (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (box_dtype != AMEPB_F32 && box_dtype != AMEPB_F64) ||
        (space != AMEPB_DEVICE && space != AMEPB_HOST) || (out && capacity < 0))
        return fail(ctx, AMEPB_E_INVAL, "amepb_rotate_crop: invalid argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    const size_t esz = dtype == AMEPB_F32 ? 4 : 8;
    const size_t cbytes = (size_t)n * 3 * esz;
    int rc;
    const void* d_coords = coords;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, cbytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, cbytes, cudaMemcpyHostToDevice, stream));
        d_coords = ctx->bufs[BUF_STAGE_A].p;
    }
    CropArgs A;
    A.dim = dim;
    A.center_is_f32 = box_dtype == AMEPB_F32;
    for (int d = 0; d < 3; ++d) {
        const double lo = box_boundary[2 * d], hi = box_boundary[2 * d + 1];
        A.lo[d] = lo;
        A.hi[d] = hi;
        if (box_dtype == AMEPB_F32) {
            const float flo = (float)lo, fhi = (float)hi;
            A.box32[d] = fhi - flo;
            A.center32[d] = (flo + fhi) / 2.0f;      // np.mean of two float32 values
            A.center[d] = (double)A.center32[d];
        } else {
            A.box32[d] = (float)(hi - lo);
            A.center[d] = (lo + hi) / 2.0;
            A.center32[d] = (float)A.center[d];
        }
    }
    A.cs = rot[0];
    A.sn = rot[1];
    const long long total = (long long)(dim == 2 ? 9 : 27) * n;
    const long long nblocks = (total + 255) / 256;
    if (nblocks >= (1ll << 31)) return fail(ctx, AMEPB_E_INVAL, "amepb_rotate_crop: too many particles");
    const size_t bbytes = sizeof(unsigned) * (size_t)(nblocks + 1);
    if ((rc = ensure_device(ctx, BUF_ORD_A, bbytes))) return rc;
    if ((rc = ensure_device(ctx, BUF_ORD_B, bbytes))) return rc;
    unsigned* bcount = reinterpret_cast<unsigned*>(ctx->bufs[BUF_ORD_A].p);
    unsigned* bstart = reinterpret_cast<unsigned*>(ctx->bufs[BUF_ORD_B].p);
    AMEPB_CUDA(ctx, cudaMemsetAsync(bcount, 0, bbytes, stream));
    if (dtype == AMEPB_F32)
        rotate_crop_kernel<float, false><<<(unsigned)nblocks, 256, 0, stream>>>((const float*)d_coords, n, total, A, bcount, bstart, nullptr, 0);
    else
        rotate_crop_kernel<double, false><<<(unsigned)nblocks, 256, 0, stream>>>((const double*)d_coords, n, total, A, bcount, bstart, nullptr, 0);
    AMEPB_LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, bcount, bstart, nblocks + 1, stream))) return rc;
    unsigned h_total = 0;
    AMEPB_CUDA(ctx, cudaMemcpyAsync(&h_total, bstart + nblocks, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    *count_out = (int64_t)h_total;
    if (!out) return 0;
    if (capacity < (int64_t)h_total)
        return fail(ctx, AMEPB_E_INVAL, "amepb_rotate_crop: %lld images are kept, the output holds %lld",
                    (long long)h_total, (long long)capacity);
    double* d_out = out;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, sizeof(double) * 3 * (size_t)std::max<unsigned>(h_total, 1)))) return rc;
        d_out = reinterpret_cast<double*>(ctx->bufs[BUF_STAGE_OUT].p);
    }
    if (dtype == AMEPB_F32)
        rotate_crop_kernel<float, true><<<(unsigned)nblocks, 256, 0, stream>>>((const float*)d_coords, n, total, A, bcount, bstart, d_out, (long long)h_total);
    else
        rotate_crop_kernel<double, true><<<(unsigned)nblocks, 256, 0, stream>>>((const double*)d_coords, n, total, A, bcount, bstart, d_out, (long long)h_total);
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * 3 * (size_t)h_total, cudaMemcpyDeviceToHost, stream));
        AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    return 0;
}

}  // extern "C"

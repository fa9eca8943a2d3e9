/*
 * amepb.h -- C ABI of libamepb.so: B200 (sm_100a) kernels for AMEP's
 * spatial-correlation hot path (periodic RDF and static structure factor).
 *
 * The reference (AMEP 1.3.0, pure Python) has no FFI for this path; its only
 * internal "operator" seam is the chunk-worker convention consumed by
 * amep/utils.py:1928-2021 (compute_parallel).  Each entry point below replaces
 * one of those workers (plus the image / grid set-up around it) and is what a
 * ctypes binding inside the reference would call:
 *
 *   amepb_frame_dimension  <- amep/utils.py:1617-1667       dimension()
 *   amepb_rdf_batch        <- amep/pbc.py:169-350           pbc_points(width=0.5)
 *                             amep/spatialcor.py:348-388    __rdf_diff
 *                             amep/spatialcor.py:566-600    chunk fan-out + sum
 *   amepb_sq_harmonics     <- amep/spatialcor.py:1402-1448  __sf_iso_chunk_fast
 *   amepb_sq_debye         <- amep/spatialcor.py:1324-1399  __sf_iso_chunk_std
 *   amepb_sq_debye_ref     <- the same, in the reference's float32 summation order
 *   amepb_sf2d_modes       <- amep/spatialcor.py:1690-1719  __sf2d_chunk_std
 *                             amep/spatialcor.py:1876-1907  chunk fan-out + sum
 *   amepb_density_counts   <- amep/continuum.py:300-407     hde (np.histogram2d of the images),
 *                             reached by sf2d/sfiso(mode='fft'), amep/spatialcor.py:1911-1940
 *   amepb_pcf2d_counts     <- amep/spatialcor.py:652-723    __dhist2d (difference vectors, rotation,
 *                             np.histogram2d) and the chunk fan-out + sum of pcf2d, :886-912
 *   amepb_pcf_angle_counts <- amep/spatialcor.py:929-1041   __dhist_angle and the chunk sum of pcf_angle, :1268-1304
 *   amepb_spatialcor_sums  <- amep/spatialcor.py:47-135     __scor_chunk (periodic KD-tree query + value
 *                             products per distance bin) and the chunk sum of spatialcor, :322-330
 *   amepb_psi_k            <- amep/order.py:1183-1348       k_nearest_neighbors (periodic KD-tree query)
 *                             amep/order.py:1361-1501       psi_k, the prologue of evaluate.SF2d(rotate=True)
 *                             and evaluate.PCF2d (amep/evaluate.py:1352-1357, 730-735)
 *   amepb_rotate_crop      <- amep/evaluate.py:1366-1377    in_box(rotate_coords(pbc_points(...)))
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary;
 *   - every call returns 0 on success, <0 for an invalid argument
 *     (AMEPB_E_*), >0 for a cudaError_t; amepb_last_error() gives the text;
 *   - `space` says where the *data* pointers (coords, other, out) live:
 *     AMEPB_DEVICE = device pointers (e.g. torch tensor.data_ptr()), work is
 *     enqueued on `stream`; AMEPB_HOST = host pointers, the library stages
 *     them through the device itself (pinned memory gives overlap) and the
 *     call returns when `out` is complete;
 *   - small per-frame parameter arrays (box, edges, directions, dim_out) are
 *     always HOST pointers;
 *   - the caller owns every buffer; inputs are never modified;
 *   - a context owns its device workspaces: use one per (device, host thread);
 *   - there is no CPU fallback: without a CUDA device amepb_create fails.
 */
#ifndef AMEPB_H
#define AMEPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMEPB_ABI_VERSION 1

/* element types of coordinate / box arrays */
#define AMEPB_F32 0
#define AMEPB_F64 1

/* memory space of data pointers */
#define AMEPB_DEVICE 0
#define AMEPB_HOST 1

/* error codes (negative) */
#define AMEPB_E_INVAL (-1)     /* bad argument                                          */
#define AMEPB_E_DIMENSION (-2) /* a frame is neither 2D nor 3D (amep/pbc.py:290-294)   */
#define AMEPB_E_NOMEM (-3)     /* workspace does not fit                                */
#define AMEPB_E_NODEVICE (-4)  /* no usable CUDA device                                 */

/* amepb_sf2d_modes accumulation */
#define AMEPB_ACCUM_C64_REFERENCE 0 /* complex64 running sum in the reference's order */
#define AMEPB_ACCUM_F64 1           /* float64 sums (more accurate than the reference) */

typedef struct amepb_ctx amepb_ctx;

/* counters of the last amepb_rdf_batch call on a context */
typedef struct amepb_rdf_info {
    int64_t frames;           /* frames processed                                  */
    int64_t images;           /* periodic images kept in the cell grids            */
    int64_t cells;            /* grid cells over all frames                        */
    int64_t block_tasks;      /* scheduling units of the pair kernel               */
    int64_t candidate_pairs;  /* query-image distances the pair kernel evaluated   */
    int64_t kernel_launches;  /* CUDA kernels launched by the call                 */
    int32_t cells_per_cutoff; /* grid refinement k (cell edge ~ rmax / k)          */
    int32_t sub_batches;      /* how many sub-batches the frames were split into   */
    double pair_kernel_ms;    /* CUDA-event time of the pair kernel launches (only  */
                              /* with amepb_set_timing(ctx, 1)), else 0            */
    double build_ms;          /* same for stats + cell build (count, scan, fill)   */
} amepb_rdf_info;

int amepb_abi_version(void);

/* Create / destroy a context bound to CUDA device `device`. */
int amepb_create(int device, amepb_ctx** out);
void amepb_destroy(amepb_ctx* ctx);

/* Text of the last error on `ctx` (ctx may be NULL: last amepb_create error). */
const char* amepb_last_error(const amepb_ctx* ctx);

/*
 * dimension() of every frame: number of non-constant columns of the (n,3)
 * coordinate array, 3 for a single particle (amep/utils.py:1617-1667).
 *   coords  [nframes][n][3] of `dtype`, in `space`
 *   dim_out [nframes] host
 */
int amepb_frame_dimension(amepb_ctx* ctx, const void* coords, int dtype, int space,
                          int64_t nframes, int64_t n, int32_t* dim_out, void* stream);

/*
 * Integer pair histograms of rdf(mode='diff') for a batch of frames.
 *
 *   coords        [nframes][n][3]   particles whose periodic images are the targets
 *   other         [nframes][m][3]   query particles, or NULL (queries = coords, m ignored)
 *   box_boundary  host [nframes][3][2] float64 (float32 boxes widened exactly);
 *                 box_dtype says which type the caller's box had -- it decides the
 *                 type `centre` is computed in (amep/pbc.py:273, 332-338)
 *   edges         host float64 [nbins+1] (edges_stride == 0, shared by all frames)
 *                 or [nframes][edges_stride]; the values NumPy compares against,
 *                 i.e. float32 edges widened exactly (amep/spatialcor.py:563, 387)
 *   pbc           non-zero: periodic images as pbc_points(width=0.5,
 *                 fold_coords=False); zero: targets are `coords` themselves
 *   hist_out      [nframes][nbins] uint64 in `space`, overwritten
 *   dim_out       host [nframes], may be NULL: dimension(coords) per frame (the
 *                 caller's epilogue needs it, amep/spatialcor.py:640-647)
 *
 * Arithmetic follows the reference exactly: images are
 * float32; float32 queries give float32 distances, float64 queries float64
 * ones; no fused multiply-add; d > 1e-3; explicit-edge binning with a
 * right-inclusive last bin.  Counts are bit-exact.
 *
 * Returns AMEPB_E_DIMENSION if pbc and some frame has dimension not in {2,3}
 * (dim_out is filled; hist_out rows of such frames are zero).
 */
int amepb_rdf_batch(amepb_ctx* ctx,
                    const void* coords, int coords_dtype, int64_t n,
                    const void* other, int other_dtype, int64_t m,
                    int space, int64_t nframes,
                    const double* box_boundary, int box_dtype,
                    const double* edges, int64_t edges_stride, int32_t nbins,
                    int pbc,
                    uint64_t* hist_out, int32_t* dim_out, void* stream);

/* Counters of the last amepb_rdf_batch call (valid after the call returned). */
int amepb_rdf_last_info(amepb_ctx* ctx, amepb_rdf_info* info);

/* Tuning knob (0 = automatic): grid refinement k used by amepb_rdf_batch. */
int amepb_rdf_set_cells_per_cutoff(amepb_ctx* ctx, int k);

/* Tuning knob (default 1): allow the symmetric evaluation of self RDFs (each unordered
 * pair of unshifted particles once, counted twice -- valid where s(i,j) == s(j,i) bit for
 * bit: float32 regime or no periodic images).  0 forces every ordered pair to be evaluated.
 * The counts are identical either way. */
int amepb_rdf_set_symmetric(amepb_ctx* ctx, int enable);

/* Intra-frame sharding (default 0 of 1): later amepb_rdf_batch calls on this context evaluate only
 * the scheduling units (runs of query cells) u with u % count == index.  Every in-range pair
 * belongs to exactly one unit, so the uint64 histograms of the `count` shards of the same call
 * -- one per GPU, each holding a copy of the frames -- add up to the full histogram exactly.
 * This is the split the reference makes over query chunks inside one frame
 * (amep/spatialcor.py:587-600), used when there are fewer frames than GPUs. */
int amepb_rdf_set_shard(amepb_ctx* ctx, int index, int count);

/*
 * Harmonic density modes along given wave vectors:
 *   out[f][d][h] = ( sum_i cos((h+1) k_d . r_i),  sum_i sin((h+1) k_d . r_i) )
 * for h = 0..nharm-1, in float64.  With k_d = dq * u_d these are the sums of
 * sfiso(mode='fast') (amep/spatialcor.py:1440-1446) for q = dq, 2dq, ...
 *
 *   coords   [nframes][n][3] of `dtype` in `space`
 *   kvec     host float64 [nframes][ndir][3]  (kvec_stride == ndir*3) or
 *            [ndir][3] shared by all frames (kvec_stride == 0)
 *   out      [nframes][ndir][nharm][2] float64 in `space`, overwritten
 */
int amepb_sq_harmonics(amepb_ctx* ctx, const void* coords, int dtype, int space,
                       int64_t nframes, int64_t n,
                       const double* kvec, int64_t kvec_stride, int32_t ndir, int32_t nharm,
                       double* out, void* stream);

/*
 * Debye scattering sum of sfiso(mode='std') (amep/spatialcor.py:1324-1399, distances from
 * amep/pbc.py:1178-1206: all ordered pairs, no periodic images):
 *   out[f][k] = sum_{i in coords, j in other} K(q_k d_ij),  K = J0 (twod) or sin(x)/x,
 * zero distances contributing 1.  Distances and arguments are rounded to float32 like the
 * reference; K and the sum are evaluated in float64 (the reference sums in float32).
 *   other  may be NULL (other = coords);  q  host float64 [nq] (q_stride 0) or [nframes][q_stride]
 *   out    [nframes][nq] float64 in `space`
 */
int amepb_sq_debye(amepb_ctx* ctx, const void* coords, int dtype, int64_t n,
                   const void* other, int other_dtype, int64_t m,
                   int space, int64_t nframes,
                   const double* q, int64_t q_stride, int32_t nq, int twod,
                   double* out, void* stream);

/*
 * The same Debye sum in the reference's own float32 arithmetic AND summation order
 * (amep/spatialcor.py:1363-1399 per chunk, :1617-1633 over the chunks): per chunk of `chunksize`
 * particles of `other` (the reference's compute_parallel tasks, amep/spatialcor.py:1583-1590),
 *   res = float32(#{d == 0}) + sum over the non-zero distances, pair order i * chunk_len + j,
 * the sum taken like NumPy takes it -- row after row (np.sum(axis=0) of the [pairs, nq] array) while
 * 4 * pairs * nq <= 1e9 bytes, else NumPy's pairwise summation per q -- and the chunk results added
 * in chunk order in float32.  J0 is evaluated in float64 and rounded (scipy's float32 loop), sin(x)/x
 * in float32 (np.sinc).  out[f][k] is the float32 sum S before the division by sqrt(N_a N_b).
 * Synchronous: the call returns when out is complete (it synchronises `stream` once per chunk).
 *   q  host float64 [nq] (q_stride 0) or [nframes][q_stride];  out [nframes][nq] float32 in `space`
 */
int amepb_sq_debye_ref(amepb_ctx* ctx, const void* coords, int dtype, int64_t n,
                       const void* other, int other_dtype, int64_t m,
                       int space, int64_t nframes,
                       const double* q, int64_t q_stride, int32_t nq, int twod,
                       int64_t chunksize, float* out, void* stream);

/* Intra-frame sharding of amepb_sf2d_modes (default 0 of 1): later calls on this context compute only the
 * rows of q tiles t (16 q_b values each in the reference-order mode, 64 in the float64 mode) with
 * t % count == index and return zeros elsewhere, so the rho arrays of the `count` shares of the same call
 * -- one per GPU, each holding a copy of the frame -- add up to the full array exactly (every element is
 * computed by one share, in the reference's order over all particles).  The split the reference makes
 * over particle chunks (amep/spatialcor.py:1876-1907) cannot be spread over devices without changing its
 * complex64 rounding sequence; a split over q can.  Used when there are fewer frames than GPUs. */
int amepb_sf2d_set_shard(amepb_ctx* ctx, int index, int count);

/*
 * Particle counts on the grid of the histogram density estimate used by sf2d(mode='fft')
 * (amep/continuum.py:300-407 hde, called through coords_to_density by amep/spatialcor.py:1918-1920):
 * periodic images as pbc_points(width=0.5, enforce_nd=2) when pbc != 0 (float32 images, strict
 * window, amep/pbc.py:270-338), binned like np.histogram2d with explicit edges (bin i holds
 * edges[i] <= v < edges[i+1], the last bin also its right edge; NaN and outside values dropped).
 *   xedges, yedges  host float64, strictly increasing, shared by all frames (np.arange values)
 *   counts          [nframes][nye-1][nxe-1] uint32 in `space` (y-major: the transposed histogram
 *                   the reference divides by delta**2)
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_density_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space,
                         int64_t nframes, const double* box_boundary, int box_dtype, int pbc,
                         const double* xedges, int32_t nxe, const double* yedges, int32_t nye,
                         uint32_t* counts, void* stream);

/*
 * Pair counts on the (dx, dy) grid of spatialcor.pcf2d (amep/spatialcor.py:652-723 __dhist2d,
 * summed over all chunks as amep/spatialcor.py:886-912 does): for every row o of `other` and every
 * row c of `coords` (skipping equal indices when same != 0) the difference o - c in the promoted
 * dtype of the two arrays -- no minimum image, exactly like the reference, whose pbc_diff calls
 * are commented out -- optionally rotated about the box centre (amep/utils.py:581-652: subtract the
 * centre, multiply by R(-angle) in float64, add the centre), binned like np.histogram2d with
 * explicit edges (bin i holds edges[i] <= v < edges[i+1], the last bin also its right edge).
 *   center      host float64 [nframes][3]: np.mean(box_boundary, axis=1) widened to float64;
 *               center_f32 != 0 says that it was float32 (then "diff - centre" of float32
 *               differences is a float32 subtraction, as NumPy's promotion makes it)
 *   rot         NULL (no rotation) or host float64 [nframes][3] = (flag, cos(-angle), sin(-angle)):
 *               flag != 0 where the reference's `angle != 0.0` holds
 *   xedges, yedges  host float64, strictly increasing; shared by all frames (edge_stride = 0) or
 *               one set per frame, edge_stride doubles apart
 *   counts      [nframes][nxe-1][nye-1] uint64 in `space` (x-major like np.histogram2d)
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_pcf2d_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other,
                       int other_dtype, int64_t m, int same, int space, int64_t nframes,
                       const double* center, int center_f32, const double* rot, const double* xedges,
                       int32_t nxe, const double* yedges, int32_t nye, int64_t edge_stride,
                       uint64_t* counts, void* stream);

/* Intra-frame sharding of amepb_pcf2d_counts (default: off, offset < 0): in later calls on this context
 * with `same == 0` the query array `other` is the slice of the target array `coords` that starts at particle
 * `offset`, and a query is not paired with the target of the same *global* index -- what the reference does
 * for other_coords=None (amep/spatialcor.py:700-703).  The uint64 grids of the slices of one frame -- one
 * per GPU -- add up to the grid of the whole frame exactly. */
int amepb_pcf2d_set_query_offset(amepb_ctx* ctx, int64_t offset);

/*
 * Pair counts on the (distance, angle) grid of spatialcor.pcf_angle (amep/spatialcor.py:929-1041): for
 * every row o of `other` and every row c of `coords` (skipping equal indices when same != 0) the
 * difference c - o in the promoted dtype of the two arrays, folded to the minimum image when pbc != 0
 * (amep.pbc.pbc_diff -> fold about the centred box, in the dtype NumPy promotes vector and box to,
 * amep/pbc.py:155-163, 597-602); dist = norm in that dtype, pairs with dist == 0 dropped;
 * theta = arccos(clip(dot(e, r) / (dist |e|), -1, 1)) in float64, 2 pi - theta where r_y < 0, minus
 * `angle` (wrapped into [0, 2 pi)) when angle != 0; (dist, theta) binned like np.histogram2d.
 * Both coordinate arrays must be flat (z == 0), as the reference enforces.
 *   e, e_norm   host float64: one direction [3] and its norm [1] (e_stride == 0) or one per row of
 *               `other`, [m][3] and [m] (e_stride == 3)
 *   dedges, aedges  host float64, increasing (float32 edges widened exactly)
 *   counts      [nde-1][nae-1] uint64 in `space`
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_pcf_angle_counts(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other,
                           int other_dtype, int64_t m, int same, int space, const double* box_boundary,
                           int box_dtype, int pbc, const double* e, int64_t e_stride, const double* e_norm,
                           double angle, const double* dedges, int32_t nde, const double* aedges, int32_t nae,
                           uint64_t* counts, void* stream);

/* Intra-frame sharding of amepb_pcf_angle_counts (default: off, offset < 0): in later calls on this context
 * with `same == 0` the query array `other` (and the per-particle directions `e`) is the slice of the target
 * array `coords` that starts at particle `offset`; a query is not paired with the target of the same global
 * index (amep/spatialcor.py:1003: coords[particlerange != n]).  The uint64 grids of the slices add up to the
 * grid of the whole frame exactly. */
int amepb_pcf_angle_set_query_offset(amepb_ctx* ctx, int64_t offset);

/*
 * Pair counts and value-product sums per distance bin of spatialcor.spatialcor
 * (amep/spatialcor.py:47-135): the periodic KD-tree of the reference restated -- `coords` shifted by
 * box/2 - centre and folded into [0, L) (amep/pbc.py:632-655), the rows of `other` shifted, folded
 * about the box centre (amep/spatialcor.py:315-319) and wrapped into [0, L) like scipy's periodic
 * tree wraps query points; float64 minimum-image distances d; pairs with d < max(box)/2 only;
 * bin i holds edges[i] <= d < edges[i+1] (no right-inclusive last bin, amep/spatialcor.py:112);
 *   norm[i] += 1,  cor[i] += sum over the components of conj(other_value) * value  (float64).
 * pbc == 0: plain float64 distances of the two arrays.
 *   other, other_values  NULL: the same particles as coords / values
 *   values      [n][ncomp], other_values [m][ncomp] in `space`; value_kind 0 float32, 1 float64,
 *               2 complex64, 3 complex128 (interleaved re, im); 1 <= ncomp <= 4
 *   norm        [nedges-1] uint64, cor [nedges-1][2] float64 (re, im), both in `space`
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_spatialcor_sums(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other,
                          int other_dtype, int64_t m, int space, const double* box_boundary, int box_dtype,
                          int pbc, const void* values, const void* other_values, int value_kind, int32_t ncomp,
                          const double* edges, int32_t nedges, uint64_t* norm, double* cor, void* stream);

/*
 * k-atic bond order parameter of every particle of one 2D frame, psi_k of amep/order.py:1361-1501
 * with its defaults (other_coords=None, pbc=True): the k nearest neighbours are ranks 2..k+1 of the
 * float64 distances (x, y and z) to all nine float32 periodic images of all particles (pbc_points
 * with enforce_nd=2 and no window, amep/pbc.py:296-326; rank 1 is the particle itself), image
 * numbers folded back with % n; psi = 1/k sum_j exp(i k theta_j) with theta_j the angle of the bond
 * to the UNSHIFTED position of neighbour j (amep/order.py:1488-1499), evaluated in the dtype of
 * `coords` like NumPy does (float32 coordinates give a complex64 result).
 *   coords        [n][3] of `dtype` in `space`;  1 <= k <= 8
 *   psi_out       [n][2] (re, im) of `dtype` in `space`
 *   neighbors_out [n][k] int32 in `space`, may be NULL: the neighbour ids in order of distance
 *   bad_out       host, may be NULL: number of particles with an image of themselves (or nothing)
 *                 among their neighbours -- the reference raises for such frames (ragged id array)
 * Equal distances are ranked by image number (scipy's KD-tree leaves that order unspecified).
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_psi_k(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space,
                const double* box_boundary, int box_dtype, int32_t k,
                void* psi_out, int32_t* neighbors_out, int64_t* bad_out, void* stream);

/*
 * Rotated periodic crop of amep/evaluate.py:1366-1377:
 *   in_box(rotate_coords(pbc_points(coords, box), angle, center), box)
 * every periodic image (dim = 2: 9, dim = 3: 27 float32 copies in the order of amep/pbc.py:305-321,
 * dim = dimension(coords)) minus the box centre (a float32 subtraction for a float32 box), rotated
 * about z in float64 (amep/utils.py:547-652), plus the centre; the images inside the closed box are
 * kept in the order of the image array (amep/utils.py:727-742).
 *   rot       host [2]: cos and sin of the rotation angle as the caller's NumPy computed them
 *   out       [capacity][3] float64 in `space`, or NULL to only count
 *   count_out host: number of images kept (an error if it exceeds `capacity`)
 * The call returns after the work is complete (it synchronises `stream`).
 */
int amepb_rotate_crop(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, int space,
                      const double* box_boundary, int box_dtype, int dim, const double* rot,
                      double* out, int64_t capacity, int64_t* count_out, void* stream);

/*
 * Density modes on the sf2d grid: rho[f][iy][ix] = sum_i exp(-i (qx[ix] x_i + qy[iy] y_i))
 * with qx = qy = (-q[nq-1], ..., -q[0], q[0], ..., q[nq-1])  (amep/spatialcor.py:1852-1857, 1874).
 *
 *   q         host float64, the positive half of the axis: [nq] (q_stride == 0,
 *             shared by all frames) or [nframes][q_stride]
 *   accum     AMEPB_ACCUM_C64_REFERENCE: per chunk of `chunksize` particles a
 *             complex64 accumulator receives exp(...) (evaluated in float64)
 *             particle by particle in index order, chunk sums are then added in
 *             order in complex64 (amep/spatialcor.py:1715-1719, 1890-1891);
 *             out is complex64 [nframes][2nq][2nq][2] float32.
 *             AMEPB_ACCUM_F64: float64 sums (chunksize ignored); out is
 *             [nframes][2nq][2nq][2] float64.
 */
int amepb_sf2d_modes(amepb_ctx* ctx, const void* coords, int dtype, int space,
                     int64_t nframes, int64_t n,
                     const double* q, int64_t q_stride, int32_t nq,
                     int accum, int64_t chunksize,
                     void* out, void* stream);

/* Record CUDA events around the kernels of amepb_rdf_batch / amepb_sq_* on the caller's
 * stream so that amepb_rdf_last_info / amepb_last_kernel_ms can report device times. */
int amepb_set_timing(amepb_ctx* ctx, int enable);

/* Device time (ms, CUDA events on the launching stream) of the main kernel of the last
 * amepb_sq_harmonics / amepb_sf2d_modes call; 0 unless timing is enabled. */
double amepb_last_kernel_ms(amepb_ctx* ctx);

/* CUDA kernels launched by this context since creation (for accounting). */
int64_t amepb_kernel_launches(const amepb_ctx* ctx);

/* Wait until all work the context enqueued on `stream` (and its own streams) is done. */
int amepb_synchronize(amepb_ctx* ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMEPB_H */

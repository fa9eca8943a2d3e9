// sq.cu -- static structure factor kernels on sm_100a (FP64 pipe).
//
//   amepb_sq_harmonics  <- amep/spatialcor.py:1402-1448 (__sf_iso_chunk_fast): for every
//       direction k_d the sums  sum_i cos(m k_d.r_i),  sum_i sin(m k_d.r_i),  m = 1..nharm.
//       One sincos per (particle, direction); the harmonics follow from the Chebyshev
//       recurrence c_{m+1} = 2 cos(phi) c_m - c_{m-1} (same for sin): 2 DFMA + 2 DADD per
//       (particle, harmonic) for both sums.  Lane partials are reduced 16 harmonics at a
//       time through a transposed shared-memory tile (no shuffles in the hot loop).
//   amepb_sf2d_modes    <- amep/spatialcor.py:1690-1719 (__sf2d_chunk_std) + :1876-1907.
//       exp(-i(qx x + qy y)) separates into per-axis tables built by a rotation recurrence.
//       F64 mode: four real product sums (cc, ss, cs, sc) per (|qx|, |qy|) give all four
//       quadrants: 4 DFMA per particle per quad point, 4x4 register tiles.
//       C64 mode: the reference's arithmetic -- per chunk a complex64 accumulator, every
//       term added in float64 and rounded to float32 (done on the FP64 pipe with the
//       1.5*2^(e+29) constant: F2F runs at a quarter of the FP64 rate), chunk sums added in
//       float32 in chunk order; particles strictly in index order.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "amepb_internal.cuh"

namespace amepb {

static size_t elem_bytes(int dtype) { return dtype == AMEPB_F32 ? 4 : 8; }

// ---------------------------------------------------------------------------------------
// harmonics (sfiso 'fast')
// ---------------------------------------------------------------------------------------
constexpr int kHarmThreads = 128;
constexpr int kHarmP = 8;        // particles per thread (the tree sum below is written for 8)
constexpr int kHarmBlockM = 16;  // harmonics reduced per round
constexpr int kHarmSeg = 512;    // harmonics per launch segment (shared accumulators)

template <typename T>
__global__ void __launch_bounds__(kHarmThreads) harmonics_kernel(
    const T* __restrict__ coords, long long n, const double* __restrict__ kvec, long long kvec_stride,
    int ndir, int h0, int hn, int nsplit, double* __restrict__ partial) {
    extern __shared__ double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x, dir = blockIdx.y, frame = blockIdx.z;
    double* my_part = smem_d + warp * (kHarmBlockM * 2 * 32);            // [16][2][32]
    double* acc_base = smem_d + 4 * (kHarmBlockM * 2 * 32);               // [4][hn*2]
    double* my_acc = acc_base + warp * (2 * hn);
    for (int i = lane; i < 2 * hn; i += 32) my_acc[i] = 0.0;
    __syncwarp();

    const double* kv = kvec + (long long)frame * kvec_stride + dir * 3;
    const double kx = kv[0], ky = kv[1], kz = kv[2];
    const T* base = coords + (long long)frame * n * 3;
    const long long tile_particles = (long long)kHarmThreads * kHarmP;
    const long long ntiles = (n + tile_particles - 1) / tile_particles;

    for (long long tile = split; tile < ntiles; tile += nsplit) {
        double cp[kHarmP], cc[kHarmP], sp[kHarmP], sc[kHarmP], tx[kHarmP];
#pragma unroll
        for (int p = 0; p < kHarmP; ++p) {
            const long long idx = tile * tile_particles + (long long)warp * 32 * kHarmP + p * 32 + lane;
            if (idx < n) {
                const double x = (double)base[3 * idx], y = (double)base[3 * idx + 1], z = (double)base[3 * idx + 2];
                const double phi = (kx * x + ky * y) + kz * z;
                double s1, c1;
                sincos(phi, &s1, &c1);
                tx[p] = 2.0 * c1;
                if (h0 == 0) {
                    cp[p] = 1.0; sp[p] = 0.0; cc[p] = c1; sc[p] = s1;
                } else {
                    sincos((double)h0 * phi, &sp[p], &cp[p]);
                    sincos((double)(h0 + 1) * phi, &sc[p], &cc[p]);
                }
            } else {
                cp[p] = cc[p] = sp[p] = sc[p] = tx[p] = 0.0;
            }
        }
        for (int hb = 0; hb < hn; hb += kHarmBlockM) {
#pragma unroll
            for (int j = 0; j < kHarmBlockM; ++j) {
                // pairwise tree (depth 3) instead of a serial chain of 7 dependent DADDs
                const double sumc = ((cc[0] + cc[1]) + (cc[2] + cc[3])) + ((cc[4] + cc[5]) + (cc[6] + cc[7]));
                const double sums = ((sc[0] + sc[1]) + (sc[2] + sc[3])) + ((sc[4] + sc[5]) + (sc[6] + sc[7]));
                my_part[(j * 2 + 0) * 32 + lane] = sumc;
                my_part[(j * 2 + 1) * 32 + lane] = sums;
#pragma unroll
                for (int p = 0; p < kHarmP; ++p) {
                    const double tc = fma(tx[p], cc[p], -cp[p]);
                    const double ts = fma(tx[p], sc[p], -sp[p]);
                    cp[p] = cc[p]; cc[p] = tc;
                    sp[p] = sc[p]; sc[p] = ts;
                }
            }
            __syncwarp();
            // lane = row (harmonic j = lane/2, component lane&1); rotated reads are conflict free
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                t0 += my_part[lane * 32 + ((i + lane) & 31)];
                t1 += my_part[lane * 32 + ((i + 1 + lane) & 31)];
                t2 += my_part[lane * 32 + ((i + 2 + lane) & 31)];
                t3 += my_part[lane * 32 + ((i + 3 + lane) & 31)];
            }
            const double tot = (t0 + t1) + (t2 + t3);
            const int h = hb + (lane >> 1);
            if (h < hn) my_acc[h * 2 + (lane & 1)] += tot;
            __syncwarp();
        }
    }
    __syncthreads();
    double* out = partial + ((((long long)frame * ndir + dir) * nsplit + split) * (long long)hn) * 2;
    for (int i = tid; i < 2 * hn; i += kHarmThreads)
        out[i] = (acc_base[i] + acc_base[2 * hn + i]) + (acc_base[4 * hn + i] + acc_base[6 * hn + i]);
}

// out[f][d][h0+h][c] = sum over splits, in split order (deterministic)
__global__ void harmonics_reduce_kernel(const double* __restrict__ partial, int nsplit, int hn, int h0,
                                        int nharm, long long nfd, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = nfd * hn * 2;
    if (i >= total) return;
    const long long fd = i / (2 * hn);
    const int r = (int)(i % (2 * hn));
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += partial[((fd * nsplit + k) * (long long)hn) * 2 + r];
    out[(fd * nharm + h0) * 2 + r] = s;
}

// ---------------------------------------------------------------------------------------
// sf2d tables: E[j][a] = (cos(q_a x_j), sin(q_a x_j)) for a run of 16 consecutive a
// ---------------------------------------------------------------------------------------
// Two exact seeds (sincos of fl(q*x), the reference's phase rounding) and a rotation
// recurrence for the rest; w = E[a0+1] * conj(E[a0]).
__device__ __forceinline__ void table_run16(double x, const double* __restrict__ q, int a0, int nq,
                                            double2* __restrict__ dst, int dst_stride) {
    const double2 zero = make_double2(0.0, 0.0);
    if (a0 >= nq) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[i * dst_stride] = zero;
        return;
    }
    double c0, s0, c1, s1;
    sincos(__dmul_rn(q[a0], x), &s0, &c0);
    dst[0] = make_double2(c0, s0);
    if (a0 + 1 >= nq) {
#pragma unroll
        for (int i = 1; i < 16; ++i) dst[i * dst_stride] = zero;
        return;
    }
    sincos(__dmul_rn(q[a0 + 1], x), &s1, &c1);
    dst[dst_stride] = make_double2(c1, s1);
    const double wr = c1 * c0 + s1 * s0, wi = s1 * c0 - c1 * s0;
    double cr = c1, ci = s1;
#pragma unroll
    for (int i = 2; i < 16; ++i) {
        const double nr = cr * wr - ci * wi;
        const double ni = cr * wi + ci * wr;
        cr = nr; ci = ni;
        dst[i * dst_stride] = (a0 + i < nq) ? make_double2(cr, ci) : zero;
    }
}

// ---------------------------------------------------------------------------------------
// sf2d, float64 product sums
// ---------------------------------------------------------------------------------------
constexpr int kF64Threads = 256;   // 8 warps x 32 lanes, 2 x 8 quad points per thread: 64 x 64 tile
constexpr int kF64Stage = 32;      // particles per shared-memory stage
constexpr int kF64Row = 65;        // 64 table entries per particle and axis + 1 of padding

template <typename T, int UNROLL>
__global__ void __launch_bounds__(kF64Threads, 1) sf2d_f64_kernel(
    const T* __restrict__ coords, long long n, const double* __restrict__ qall, long long q_stride, int nq,
    int ksplit, double* __restrict__ partial, int shard_index, int shard_count) {
    // (intra-frame sharding, amepb_sf2d_set_shard: a block whose row of tiles belongs to another share has nothing to do)
    if ((int)(blockIdx.x / ((nq + 63) / 64)) % shard_count != shard_index) return;
    // shared: X[stage][64 + 1], Y[stage][64 + 1] as double2 (65 KB, dynamic)
    extern __shared__ __align__(16) unsigned char smem_f64[];
    // (rows padded by one entry: the builder's lanes hold different particles, a row apart, and
    // would otherwise all store to the same four banks)
    double2 (*sX)[kF64Row] = reinterpret_cast<double2 (*)[kF64Row]>(smem_f64);
    double2 (*sY)[kF64Row] = reinterpret_cast<double2 (*)[kF64Row]>(smem_f64 + sizeof(double2) * kF64Stage * kF64Row);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;   // lanes: qx = lane, lane + 32; warp: 8 consecutive qy
    const int tiles = (nq + 63) / 64;
    const int tile_a = blockIdx.x % tiles, tile_b = blockIdx.x / tiles;
    const int split = blockIdx.y, frame = blockIdx.z;
    const double* q = qall + (long long)frame * q_stride;
    const T* base = coords + (long long)frame * n * 3;
    const long long per = (n + ksplit - 1) / ksplit;
    const long long j_begin = (long long)split * per;
    const long long j_end = j_begin + per < n ? j_begin + per : n;

    // 2 x 8 register tile: the two X entries are conflict-free per-lane loads, the eight Y
    // entries are warp-uniform broadcasts (16 shared-memory wavefronts per particle and warp
    // instead of 24 with a 4 x 4 tile over 16 x 16 threads)
    double acc[2][8][4];  // [ia][ib][cc, ss, cs, sc]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][k][c] = 0.0;

    for (long long j0 = j_begin; j0 < j_end; j0 += kF64Stage) {
        // table build: thread -> (particle tid % 32, run tid / 32): runs 0-3 are X, 4-7 are Y
        {
            const int jj = tid & 31, run = tid >> 5;
            const long long j = j0 + jj;
            const bool isx = run < 4;
            const int a0 = (isx ? tile_a : tile_b) * 64 + (run & 3) * 16;
            double2* dst = (isx ? &sX[jj][0] : &sY[jj][0]) + (run & 3) * 16;
            if (j < j_end) {
                const double v = (double)base[3 * j + (isx ? 0 : 1)];
                table_run16(v, q, a0, nq, dst, 1);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[i] = make_double2(0.0, 0.0);
            }
        }
        __syncthreads();
#pragma unroll UNROLL
        for (int jj = 0; jj < kF64Stage; ++jj) {
            double2 xa[2], yb[8];
#pragma unroll
            for (int i = 0; i < 2; ++i) xa[i] = sX[jj][lane + 32 * i];
#pragma unroll
            for (int k = 0; k < 8; ++k) yb[k] = sY[jj][warp * 8 + k];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[i][k][0] = fma(xa[i].x, yb[k].x, acc[i][k][0]);
                    acc[i][k][1] = fma(xa[i].y, yb[k].y, acc[i][k][1]);
                    acc[i][k][2] = fma(xa[i].x, yb[k].y, acc[i][k][2]);
                    acc[i][k][3] = fma(xa[i].y, yb[k].x, acc[i][k][3]);
                }
        }
        __syncthreads();
    }
    // partial[f][split][4][nq][nq]
    double* out = partial + (((long long)frame * ksplit + split) * 4) * (long long)nq * nq;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int a = tile_a * 64 + lane + 32 * i, b = tile_b * 64 + warp * 8 + k;
            if (a < nq && b < nq)
#pragma unroll
                for (int c = 0; c < 4; ++c) out[((long long)c * nq + b) * nq + a] = acc[i][k][c];
        }
}

// combine the splits and expand the four quadrants: rho[f][iy][ix] complex128
__global__ void sf2d_f64_finish_kernel(const double* __restrict__ partial, int ksplit, int nq, long long nframes,
                                       double2* __restrict__ rho, int shard_index, int shard_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)nq * nq;
    if (i >= nframes * per) return;
    const long long f = i / per;
    const int b = (int)((i % per) / nq), a = (int)(i % nq);
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if ((b / 64) % shard_count == shard_index)   // (rows of other shares stay zero: their partial sums were never written)
        for (int k = 0; k < ksplit; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[c] += partial[(((f * ksplit + k) * 4 + c) * (long long)nq + b) * nq + a];
    const double cc = s[0], ss = s[1], cs = s[2], sc = s[3];
    const double2 pp = make_double2(cc - ss, -(sc + cs));  // (+qa, +qb)
    const double2 mp = make_double2(cc + ss, sc - cs);     // (-qa, +qb)
    const int g = 2 * nq;
    double2* r = rho + f * (long long)g * g;
    const int ixp = nq + a, ixm = nq - 1 - a, iyp = nq + b, iym = nq - 1 - b;
    r[(long long)iyp * g + ixp] = pp;
    r[(long long)iyp * g + ixm] = mp;
    r[(long long)iym * g + ixm] = make_double2(pp.x, -pp.y);
    r[(long long)iym * g + ixp] = make_double2(mp.x, -mp.y);
}

// ---------------------------------------------------------------------------------------
// sf2d, complex64 accumulation in the reference's order
// ---------------------------------------------------------------------------------------
constexpr int kC64Threads = 256;  // 8 warps: lanes = 32 qx values, each warp 2 qy values
constexpr int kC64TileA = 32;
constexpr int kC64TileB = 16;
// rows of the shared tables are padded by one entry: the builder's lanes hold different particles
// (row stride apart) and would otherwise all store to the same four banks
constexpr int kC64RowA = kC64TileA + 1;
constexpr int kC64RowB = kC64TileB + 1;
constexpr int kC64Stage = 32;     // particles per shared-memory stage (two stages resident)
constexpr size_t kC64ChunkBufferBytes = (size_t)4 << 30;   // chunk sums of the split kernel, per launch

// fl32(v) for a double v without the FP64 pipe (which the two DFMA per accumulation already
// keep busy) and without F2F (a quarter of the FP64 rate): add half a float32 ulp to the bit
// pattern (64-bit integer add: IADD3 + IADD3.X) and clear the 29 low mantissa bits (LOP3).
// This is round-half-away instead of round-half-even: the two differ only on exact ties, one
// addition in 2^29 -- two orders of magnitude rarer than the rounding flips that the 1e-16
// differences between the tables and libm's exp already cause (section 4 of DESIGN.md).
__device__ __forceinline__ double round_to_f32(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v) + 0x10000000ull;
    return __longlong_as_double((long long)(b & 0xffffffffe0000000ull));
}

// One thread owns the quad points (qa, qb0) and (qa, qb1): the X entry is loaded once per
// particle and lane (conflict free, 4 wavefronts per warp), the two Y entries are
// warp-uniform broadcasts -- 3 shared-memory wavefronts per quad-point step instead of 8
// with a 16 x 16 thread tile, which had the shared pipe at 62 % (ncu) next to the FP64 pipe.
// Tables of one stage (2 X runs and 1 Y run of 16 entries per particle -> 3 * kC64Stage threads).
// Not inlined on purpose: the double-precision sincos would otherwise inflate the register
// allocation of the accumulation loop it is interleaved with.
template <typename T>
__device__ __noinline__ void c64_build_stage(const T* __restrict__ base, const double* __restrict__ q, long long j0,
                                             long long j_hi, int xa0, int ya0, int nq, double2* __restrict__ bX,
                                             double2* __restrict__ bY) {
    const int tid = threadIdx.x;
    if (tid >= 3 * kC64Stage) return;
    const int jj = tid % kC64Stage, run = tid / kC64Stage;   // run 0,1: X; run 2: Y
    const long long j = j0 + jj;
    if (j >= j_hi) return;
    if (run < 2) table_run16((double)base[3 * j], q, xa0 + run * 16, nq, bX + jj * kC64RowA + run * 16, 1);
    else table_run16((double)base[3 * j + 1], q, ya0, nq, bY + jj * kC64RowB, 1);
}

// SPLIT: gridDim.z blocks share the particles of a tile by whole chunks and write every chunk sum
// (complex64) to chunk_out instead of accumulating the total; sf2d_c64_combine_kernel then adds
// the chunk sums in chunk order in float32, exactly like the sequential kernel.  Used when a call
// has too few frames to fill the GPU (one frame of configs[3] is 128 blocks for 148 SMs).
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(kC64Threads, 4) sf2d_c64_kernel(
    const T* __restrict__ coords, long long n, const double* __restrict__ qall, long long q_stride, int nq,
    long long chunksize, float2* __restrict__ rho, float4* __restrict__ chunk_out, long long nchunks, int shard_index,
    int shard_count) {
    // (intra-frame sharding, amepb_sf2d_set_shard: rows of tiles of other shares are left to them; rho is zeroed beforehand)
    if ((int)(blockIdx.x / ((nq + kC64TileA - 1) / kC64TileA)) % shard_count != shard_index) return;
    extern __shared__ __align__(16) unsigned char smem_c64[];
    // (one spare row per table: the operand prefetch of particle jj + 1 then needs no index clamp,
    // which cost five uniform-datapath instructions per particle in the accumulation loop)
    constexpr size_t kBufBytes = sizeof(double2) * (kC64Stage + 1) * (kC64RowA + kC64RowB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_a = (nq + kC64TileA - 1) / kC64TileA;
    const int tile_a = blockIdx.x % tiles_a, tile_b = blockIdx.x / tiles_a;
    const int frame = blockIdx.y;
    const double* q = qall + (long long)frame * q_stride;
    const T* base = coords + (long long)frame * n * 3;

    // chunk accumulators (float32 values held in doubles) and running totals (float32):
    // [quad point 0 / 1][pp real, pp imag, mp real, mp imag]
    double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    float tot[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    long long j_lo = 0, j_hi = n, chunk_idx = 0;
    if (SPLIT) {
        chunk_idx = nchunks * blockIdx.z / gridDim.z;
        const long long c_hi = nchunks * (blockIdx.z + 1) / gridDim.z;
        j_lo = chunk_idx * chunksize;
        j_hi = c_hi * chunksize < n ? c_hi * chunksize : n;
    }
    long long chunk_end = j_lo + chunksize < n ? j_lo + chunksize : n;
    const int qa_idx = tile_a * kC64TileA + lane;

    auto buf_x = [&](int b) { return reinterpret_cast<double2*>(smem_c64 + b * kBufBytes); };
    auto buf_y = [&](int b) { return reinterpret_cast<double2*>(smem_c64 + b * kBufBytes + sizeof(double2) * (kC64Stage + 1) * kC64RowA); };
    c64_build_stage<T>(base, q, j_lo, j_hi, tile_a * kC64TileA, tile_b * kC64TileB, nq, buf_x(0), buf_y(0));
    __syncthreads();
    int cur = 0;
    for (long long j0 = j_lo; j0 < j_hi; j0 += kC64Stage, cur ^= 1) {
        // the tables of the next stage are built by three of the eight warps before they join the
        // accumulation of this stage; one barrier per stage
        if (j0 + kC64Stage < j_hi)
            c64_build_stage<T>(base, q, j0 + kC64Stage, j_hi, tile_a * kC64TileA, tile_b * kC64TileB, nq, buf_x(cur ^ 1), buf_y(cur ^ 1));
        const double2 (*sX)[kC64RowA] = reinterpret_cast<const double2 (*)[kC64RowA]>(buf_x(cur));
        const double2 (*sY)[kC64RowB] = reinterpret_cast<const double2 (*)[kC64RowB]>(buf_y(cur));
        const int lim = (int)((j_hi - j0) < kC64Stage ? (j_hi - j0) : kC64Stage);
        int jj = 0;
        while (jj < lim) {
            // run to the next chunk boundary (or the end of the stage) without per-particle checks
            const long long to_chunk = chunk_end - (j0 + jj);
            const int seg_end = jj + (int)(to_chunk < (long long)(lim - jj) ? to_chunk : (long long)(lim - jj));
            // operands of the next particle are fetched while the current one is accumulated
            double2 xa = sX[jj][lane], y0 = sY[jj][2 * warp], y1 = sY[jj][2 * warp + 1];
#pragma unroll 2
            for (; jj < seg_end; ++jj) {
                const double2 xa_n = sX[jj + 1][lane], y0_n = sY[jj + 1][2 * warp], y1_n = sY[jj + 1][2 * warp + 1];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const double2 yb = p == 0 ? y0 : y1;
                    // exp(-i(qa x + qb y)) = (cc - ss) - i (sc + cs);  exp(-i(-qa x + qb y)) = (cc + ss) + i (sc - cs)
                    // acc + term in float64 as two fused multiply-adds, then the float32 rounding of the reference
                    acc[p][0] = round_to_f32(fma(-xa.y, yb.y, fma(xa.x, yb.x, acc[p][0])));
                    acc[p][1] = round_to_f32(fma(-xa.x, yb.y, fma(-xa.y, yb.x, acc[p][1])));
                    acc[p][2] = round_to_f32(fma(xa.y, yb.y, fma(xa.x, yb.x, acc[p][2])));
                    acc[p][3] = round_to_f32(fma(-xa.x, yb.y, fma(xa.y, yb.x, acc[p][3])));
                }
                xa = xa_n; y0 = y0_n; y1 = y1_n;
            }
            if (j0 + jj == chunk_end) {
                // rhoq += res : complex64 + complex64 (amep/spatialcor.py:1890-1891)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (SPLIT) {
                        const int b = tile_b * kC64TileB + 2 * warp + p;
                        if (qa_idx < nq && b < nq)
                            chunk_out[(((long long)frame * nchunks + chunk_idx) * nq + b) * nq + qa_idx] =
                                make_float4((float)acc[p][0], (float)acc[p][1], (float)acc[p][2], (float)acc[p][3]);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (!SPLIT) tot[p][c] = __fadd_rn(tot[p][c], (float)acc[p][c]);
                        acc[p][c] = 0.0;
                    }
                }
                ++chunk_idx;
                chunk_end = chunk_end + chunksize < n ? chunk_end + chunksize : n;
            }
        }
        __syncthreads();
    }
    if (SPLIT) return;
    const int a = qa_idx;
    const int g = 2 * nq;
    float2* r = rho + (long long)frame * g * g;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int b = tile_b * kC64TileB + 2 * warp + p;
        if (a < nq && b < nq) {
            const int ixp = nq + a, ixm = nq - 1 - a, iyp = nq + b, iym = nq - 1 - b;
            r[(long long)iyp * g + ixp] = make_float2(tot[p][0], tot[p][1]);
            r[(long long)iyp * g + ixm] = make_float2(tot[p][2], tot[p][3]);
            r[(long long)iym * g + ixm] = make_float2(tot[p][0], -tot[p][1]);
            r[(long long)iym * g + ixp] = make_float2(tot[p][2], -tot[p][3]);
        }
    }
}

// Sequential float32 sum of the chunk sums of the SPLIT kernel, in chunk order.
__global__ void sf2d_c64_combine_kernel(const float4* __restrict__ chunk_out, long long nchunks, int nq, long long nframes,
                                        float2* __restrict__ rho, int shard_index, int shard_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nframes * nq * nq) return;
    const int a = (int)(i % nq), b = (int)((i / nq) % nq);
    const long long frame = i / ((long long)nq * nq);
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    const float4* src = chunk_out + (frame * nchunks * nq + b) * nq + a;
    const bool mine = (b / kC64TileB) % shard_count == shard_index;   // (other shares' rows: zeros)
    for (long long c = 0; mine && c < nchunks; ++c) {
        const float4 v = src[c * nq * nq];
        t0 = __fadd_rn(t0, v.x); t1 = __fadd_rn(t1, v.y); t2 = __fadd_rn(t2, v.z); t3 = __fadd_rn(t3, v.w);
    }
    const int g = 2 * nq;
    float2* r = rho + frame * g * g;
    const int ixp = nq + a, ixm = nq - 1 - a, iyp = nq + b, iym = nq - 1 - b;
    r[(long long)iyp * g + ixp] = make_float2(t0, t1);
    r[(long long)iyp * g + ixm] = make_float2(t2, t3);
    r[(long long)iym * g + ixm] = make_float2(t0, -t1);
    r[(long long)iym * g + ixp] = make_float2(t2, -t3);
}

// ---------------------------------------------------------------------------------------
// Debye scattering sum (sfiso 'std'): out[k] = sum over all ordered pairs (a_i, b_j) of
// J0(q_k d_ij) (2D) or sin(q_k d_ij)/(q_k d_ij) (3D), no periodic images; zero distances
// contribute 1 (amep/spatialcor.py:1363-1399).  Distances are rounded like the reference
// (float32 coordinates, float32 distance and argument); the kernel and the sum are float64.
// ---------------------------------------------------------------------------------------
constexpr int kDebyeThreads = 256;
constexpr int kDebyeTileB = 16;   // b particles held in registers per thread

template <typename TA, typename TB, bool TWOD>
__global__ void __launch_bounds__(kDebyeThreads) debye_kernel(
    const TA* __restrict__ a, long long na, const TB* __restrict__ b, long long nb,
    const double* __restrict__ qall, long long q_stride, int nq, double* __restrict__ partial) {
    extern __shared__ double s_acc[];  // [8 warps][nq]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    const TA* fa = a + (long long)frame * na * 3;
    const TB* fb = b + (long long)frame * nb * 3;
    const double* q = qall + (long long)frame * q_stride;
    double* my_acc = s_acc + (size_t)warp * nq;
    for (int k = lane; k < nq; k += 32) my_acc[k] = 0.0;
    __syncwarp();
    const long long tiles_a = (na + kDebyeThreads - 1) / kDebyeThreads;
    const long long tiles_b = (nb + kDebyeTileB - 1) / kDebyeTileB;
    for (long long tile = blockIdx.x; tile < tiles_a * tiles_b; tile += gridDim.x) {
        const long long ta = tile % tiles_a, tb = tile / tiles_a;
        const long long i = ta * kDebyeThreads + tid;
        float d32[kDebyeTileB];
        bool have[kDebyeTileB];
        float ax = 0.f, ay = 0.f, az = 0.f;
        const bool have_a = i < na;
        if (have_a) {
            ax = (float)fa[3 * i]; ay = (float)fa[3 * i + 1]; az = (float)fa[3 * i + 2];
        }
#pragma unroll
        for (int j = 0; j < kDebyeTileB; ++j) {
            const long long jj = tb * kDebyeTileB + j;
            have[j] = have_a && jj < nb;
            d32[j] = 0.f;
            if (have[j]) {
                // scipy cdist on float32 input: differences and the root in float64, then .astype(float32)
                const double dx = (double)ax - (double)(float)fb[3 * jj];
                const double dy = (double)ay - (double)(float)fb[3 * jj + 1];
                const double dz = (double)az - (double)(float)fb[3 * jj + 2];
                d32[j] = (float)sqrt(dx * dx + dy * dy + dz * dz);
            }
        }
        for (int k = 0; k < nq; ++k) {
            const float q32 = (float)q[k];
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < kDebyeTileB; ++j) {
                if (!have[j]) continue;
                double term = 1.0;
                if (d32[j] != 0.f) {
                    if (TWOD) {
                        term = j0((double)__fmul_rn(d32[j], q32));
                    } else {
                        // x = d * (q/pi) in float32, y = pi*x in float32, sin(y)/y (np.sinc)
                        const float x = __fmul_rn(d32[j], __fdiv_rn(q32, 3.14159274f));
                        const double y = (double)__fmul_rn(3.14159274f, x);
                        term = y != 0.0 ? sin(y) / y : 1.0;
                    }
                }
                sum += term;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) my_acc[k] += sum;
        }
    }
    __syncthreads();
    double* out = partial + ((long long)frame * gridDim.x + blockIdx.x) * nq;
    for (int k = tid; k < nq; k += kDebyeThreads) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kDebyeThreads / 32; ++w) v += s_acc[(size_t)w * nq + k];
        out[k] = v;
    }
}

__global__ void debye_reduce_kernel(const double* __restrict__ partial, int nblocks, int nq, long long nframes,
                                    double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nframes * nq) return;
    const long long f = i / nq;
    const int k = (int)(i % nq);
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(f * nblocks + b) * nq + k];
    out[i] = s;
}


// ---------------------------------------------------------------------------------------
// Debye sum in the reference's own float32 order (amep/spatialcor.py:1363-1399), one chunk of
// `other` at a time like compute_parallel hands them out:
//   dists = cdist(coords32, other32[chunk]).flatten().astype(float32)      pair o = i * clen + j
//   res   = float32(count(dists == 0));  dists = dists[dists != 0]
//   res  += np.sum(K(dists[:, None] * q32[None, :]), axis=0)    if 4 n nq <= 1e9: row after row
//   res[k] += np.sum(K(dists * q32[k]))                         else: NumPy's pairwise summation
// with K = float32(j0(float64(x))) (scipy's float loop evaluates cephes j0 in double) or np.sinc in
// float32.  The terms are evaluated in parallel; only the float32 additions follow the reference's
// order: sequential over the rows (one thread per q over a buffer of terms), or the pairwise tree --
// leaves of at most 128 consecutive elements summed with eight strided accumulators, one leaf and 32
// q values per warp, then a host-built postfix program of the tree (PUSH leaf / ADD) run by one
// thread per q.
// ---------------------------------------------------------------------------------------
constexpr int kRefLeaf = 128;   // NumPy's PW_BLOCKSIZE

template <typename TA, typename TB>
__device__ __forceinline__ float ref_pair_d32(const TA* __restrict__ a, const TB* __restrict__ b, long long i, long long j) {
    // scipy cdist on float32 input: differences, squares and the root in float64, then .astype(float32)
    const double dx = (double)(float)a[3 * i] - (double)(float)b[3 * j];
    const double dy = (double)(float)a[3 * i + 1] - (double)(float)b[3 * j + 1];
    const double dz = (double)(float)a[3 * i + 2] - (double)(float)b[3 * j + 2];
    return (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

template <bool TWOD>
__device__ __forceinline__ float ref_term(float d, float qk) {
    if (TWOD) return (float)j0((double)__fmul_rn(d, qk));     // qk = float32(q)
    // np.sinc: x = d * (q32 / pi) [qk], y = pi * where(x == 0, 1e-20, x), sin(y) / y -- all float32
    float x = __fmul_rn(d, qk);
    if (x == 0.f) x = 1.0e-20f;
    const float y = __fmul_rn(3.14159274f, x);
    return __fdiv_rn((float)sin((double)y), y);
}

// pairs of the chunk with distance zero (original pair indices, any order)
template <typename TA, typename TB>
__global__ void debye_ref_zero_kernel(const TA* __restrict__ a, long long na, const TB* __restrict__ b, long long clen,
                                      unsigned long long* __restrict__ count, long long* __restrict__ list, long long cap) {
    const long long total = na * clen;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const long long i = o / clen, j = o - i * clen;
        if (ref_pair_d32(a, b, i, j) == 0.f) {
            const unsigned long long at = atomicAdd(count, 1ull);
            if ((long long)at < cap) list[at] = o;
        }
    }
}

// original pair index of the e-th non-zero distance: e + #{k : zp[k] <= e}, zp[k] = zero[k] - k (sorted)
__device__ __forceinline__ long long ref_orig_index(long long e, const long long* __restrict__ zp, int nz) {
    int lo = 0, hi = nz;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zp[mid] <= e) lo = mid + 1;
        else hi = mid;
    }
    return e + lo;
}

// One warp per group of <= 128 consecutive non-zero distances [start, start + len).
// LEAF: out[g][k] = NumPy's block sum of the terms; else out[start + e][k] = term (row-major buffer).
template <typename TA, typename TB, bool TWOD, bool LEAF>
__global__ void __launch_bounds__(128) debye_ref_terms_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long clen,
                                                              const long long* __restrict__ zp, int nz,
                                                              const long long* __restrict__ gstart, const int* __restrict__ glen,
                                                              long long ngroups, const float* __restrict__ qk, int q0, int nqt,
                                                              float* __restrict__ out) {
    __shared__ float s_d[4][kRefLeaf];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long g = (long long)blockIdx.x * 4 + warp; g < ngroups; g += (long long)gridDim.x * 4) {
        const long long start = gstart[g];
        const int len = glen[g];
        __syncwarp();
        for (int e = lane; e < len; e += 32) {
            const long long o = ref_orig_index(start + e, zp, nz);
            const long long i = o / clen;
            s_d[warp][e] = ref_pair_d32(a, b, i, o - i * clen);
        }
        __syncwarp();
        for (int kt = lane; kt < nqt; kt += 32) {
            const float q = qk[q0 + kt];
            if (!LEAF) {
                for (int e = 0; e < len; ++e) out[(start + e) * nqt + kt] = ref_term<TWOD>(s_d[warp][e], q);
                continue;
            }
            float res;
            if (len < 8) {
                res = 0.f;
                for (int e = 0; e < len; ++e) res = __fadd_rn(res, ref_term<TWOD>(s_d[warp][e], q));
            } else {
                float r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = ref_term<TWOD>(s_d[warp][u], q);
                int e = 8;
                for (; e < len - (len % 8); e += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) r[u] = __fadd_rn(r[u], ref_term<TWOD>(s_d[warp][e + u], q));
                }
                res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
                for (; e < len; ++e) res = __fadd_rn(res, ref_term<TWOD>(s_d[warp][e], q));
            }
            out[g * nqt + kt] = res;
        }
    }
}

// np.sum(terms, axis=0): float32 rows added one after the other
__global__ void debye_ref_rows_kernel(const float* __restrict__ terms, long long nrows, int nqt, float* __restrict__ sums) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nqt) return;
    float acc = 0.f;
    if (nrows > 0) acc = terms[k];
    for (long long e = 1; e < nrows; ++e) acc = __fadd_rn(acc, terms[e * nqt + k]);
    sums[k] = acc;
}

// NumPy's pairwise tree over the leaf sums: postfix program, op >= 0: push leaf op, op < 0: add
__global__ void debye_ref_tree_kernel(const float* __restrict__ leaf, const int* __restrict__ prog, long long nops, int nqt,
                                      float* __restrict__ sums) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nqt) return;
    float st[64];
    int sp = 0;
    for (long long t = 0; t < nops; ++t) {
        const int op = prog[t];
        if (op >= 0) st[sp++] = leaf[(long long)op * nqt + k];
        else { --sp; st[sp - 1] = __fadd_rn(st[sp - 1], st[sp]); }
    }
    sums[k] = sp > 0 ? st[0] : 0.f;
}

// res = float32(count0) + sums;  S = first ? res : S + res   (amep/spatialcor.py:1374-1375, 1631-1633)
__global__ void debye_ref_accum_kernel(const float* __restrict__ sums, float count0, int q0, int nqt, int first, float* __restrict__ S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nqt) return;
    const float res = __fadd_rn(count0, sums[k]);
    S[q0 + k] = first ? res : __fadd_rn(S[q0 + k], res);
}

}  // namespace amepb

using namespace amepb;

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
namespace {

// frames per staging chunk for host-space inputs
int64_t stage_frames(int64_t nframes, size_t frame_bytes, size_t out_frame_bytes) {
    const size_t budget = (size_t)2 << 30;
    int64_t fb = (int64_t)std::max<size_t>(1, budget / std::max<size_t>(1, frame_bytes + out_frame_bytes));
    return std::max<int64_t>(1, std::min<int64_t>(fb, std::min<int64_t>(nframes, 32768)));
}

}  // namespace

extern "C" {

int amepb_sq_harmonics(amepb_ctx* ctx, const void* coords, int dtype, int space, int64_t nframes, int64_t n,
                       const double* kvec, int64_t kvec_stride, int32_t ndir, int32_t nharm, double* out,
                       void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !kvec || !out || nframes < 0 || n <= 0 || ndir <= 0 || nharm <= 0 ||
        (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (space != AMEPB_DEVICE && space != AMEPB_HOST) ||
        (kvec_stride != 0 && kvec_stride < (int64_t)ndir * 3) || ndir > 65535)
        return fail(ctx, AMEPB_E_INVAL, "amepb_sq_harmonics: invalid argument");
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    timing_reset(ctx);

    const size_t frame_bytes = (size_t)n * 3 * elem_bytes(dtype);
    const size_t out_frame = (size_t)ndir * nharm * 2 * sizeof(double);
    const int64_t fb = space == AMEPB_HOST ? stage_frames(nframes, frame_bytes, out_frame)
                                           : std::min<int64_t>(nframes, 32768);
    // direction table on the device
    const size_t kv_count = (size_t)(kvec_stride ? nframes * kvec_stride : (int64_t)ndir * 3);
    int rc = ensure_pinned(ctx, kv_count * sizeof(double));
    if (rc) return rc;
    rc = ensure_device(ctx, BUF_SQ_A, kv_count * sizeof(double));
    if (rc) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;
    memcpy(ctx->pinned, kvec, kv_count * sizeof(double));
    AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_SQ_A].p, ctx->pinned, kv_count * sizeof(double),
                                    cudaMemcpyHostToDevice, stream));
    if ((rc = pinned_device_release(ctx, stream))) return rc;
    const double* d_kvec = reinterpret_cast<const double*>(ctx->bufs[BUF_SQ_A].p);

    const long long tile_particles = (long long)kHarmThreads * kHarmP;
    const long long ntiles = (n + tile_particles - 1) / tile_particles;
    const size_t smem_max = (size_t)(4 * kHarmBlockM * 2 * 32 + 4 * 2 * kHarmSeg) * sizeof(double);
    if (!ctx->sq_attr_harm) {
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(harmonics_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(harmonics_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        ctx->sq_attr_harm = true;
    }

    for (int64_t f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = (int)std::min<int64_t>(fb, nframes - f0);
        const void* d_coords = (const char*)coords + (size_t)f0 * frame_bytes;
        double* d_out = out + (size_t)f0 * ndir * nharm * 2;
        if (space == AMEPB_HOST) {
            if ((rc = ensure_device(ctx, BUF_STAGE_A, frame_bytes * nf))) return rc;
            if ((rc = ensure_device(ctx, BUF_STAGE_OUT, out_frame * nf))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, d_coords, frame_bytes * nf, cudaMemcpyHostToDevice, stream));
            d_coords = ctx->bufs[BUF_STAGE_A].p;
            d_out = reinterpret_cast<double*>(ctx->bufs[BUF_STAGE_OUT].p);
        }
        // splits of the particle range per (frame, direction)
        const long long fd = (long long)nf * ndir;
        // split the particle range so that the blocks fill whole waves of resident blocks
        int occ = 1;
        {
            const size_t smem_seg = (size_t)(4 * kHarmBlockM * 2 * 32 + 4 * 2 * std::min<int>(kHarmSeg, nharm)) * sizeof(double);
            if (dtype == AMEPB_F32)
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, harmonics_kernel<float>, kHarmThreads, smem_seg);
            else
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, harmonics_kernel<double>, kHarmThreads, smem_seg);
            if (occ < 1) occ = 1;
        }
        const long long wave = (long long)ctx->sm_count * occ;
        long long nsplit = 1;
        {
            const long long max_split = std::max<long long>(1, std::min<long long>(ntiles, 4 * wave / std::max<long long>(1, fd) + 1));
            double best_eff = -1.0;
            for (long long k = 1; k <= max_split; ++k) {
                const long long blocks = fd * k;
                const long long waves = (blocks + wave - 1) / wave;
                const double eff = (double)blocks / (double)(waves * wave);
                if (eff > best_eff + 1e-9) {
                    best_eff = eff;
                    nsplit = k;
                }
                if (waves >= 2 && eff > 0.97) break;
            }
        }
        for (int h0 = 0; h0 < nharm; h0 += kHarmSeg) {
            const int hn = std::min<int>(kHarmSeg, nharm - h0);
            const size_t part_bytes = (size_t)fd * nsplit * hn * 2 * sizeof(double);
            if ((rc = ensure_device(ctx, BUF_SQ_B, part_bytes))) return rc;
            double* d_part = reinterpret_cast<double*>(ctx->bufs[BUF_SQ_B].p);
            const size_t smem = (size_t)(4 * kHarmBlockM * 2 * 32 + 4 * 2 * hn) * sizeof(double);
            dim3 grid((unsigned)nsplit, (unsigned)ndir, (unsigned)nf);
            const double* kv = d_kvec + (kvec_stride ? f0 * kvec_stride : 0);
            cudaEvent_t ev0 = timing_event(ctx, stream);
            if (dtype == AMEPB_F32)
                harmonics_kernel<float><<<grid, kHarmThreads, smem, stream>>>((const float*)d_coords, n, kv, kvec_stride, ndir, h0, hn, (int)nsplit, d_part);
            else
                harmonics_kernel<double><<<grid, kHarmThreads, smem, stream>>>((const double*)d_coords, n, kv, kvec_stride, ndir, h0, hn, (int)nsplit, d_part);
            AMEPB_LAUNCH_CHECK(ctx);
            cudaEvent_t ev1 = timing_event(ctx, stream);
            if (ev0 && ev1) ctx->ev_sq.emplace_back(ev0, ev1);
            const long long total = fd * hn * 2;
            harmonics_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_part, (int)nsplit, hn, h0, nharm, fd, d_out);
            AMEPB_LAUNCH_CHECK(ctx);
        }
        if (space == AMEPB_HOST) {
            AMEPB_CUDA(ctx, cudaMemcpyAsync(out + (size_t)f0 * ndir * nharm * 2, d_out, out_frame * nf, cudaMemcpyDeviceToHost, stream));
            AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
        }
    }
    return 0;
}

int amepb_sf2d_modes(amepb_ctx* ctx, const void* coords, int dtype, int space, int64_t nframes, int64_t n,
                     const double* q, int64_t q_stride, int32_t nq, int accum, int64_t chunksize, void* out,
                     void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    if (!coords || !q || !out || nframes < 0 || n <= 0 || nq <= 0 || nq > 16384 ||
        (dtype != AMEPB_F32 && dtype != AMEPB_F64) || (space != AMEPB_DEVICE && space != AMEPB_HOST) ||
        (q_stride != 0 && q_stride < nq) || (accum != AMEPB_ACCUM_C64_REFERENCE && accum != AMEPB_ACCUM_F64))
        return fail(ctx, AMEPB_E_INVAL, "amepb_sf2d_modes: invalid argument");
    if (accum == AMEPB_ACCUM_C64_REFERENCE && chunksize <= 0)
        return fail(ctx, AMEPB_E_INVAL, "amepb_sf2d_modes: chunksize must be positive in the complex64 mode");
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    timing_reset(ctx);

    const size_t frame_bytes = (size_t)n * 3 * elem_bytes(dtype);
    const size_t g = (size_t)2 * nq;
    const size_t out_elem = accum == AMEPB_ACCUM_F64 ? sizeof(double2) : sizeof(float2);
    const size_t out_frame = g * g * out_elem;
    int64_t fb = space == AMEPB_HOST ? stage_frames(nframes, frame_bytes, out_frame)
                                     : std::min<int64_t>(nframes, 32768);

    const size_t q_count = (size_t)(q_stride ? nframes * q_stride : nq);
    int rc = ensure_pinned(ctx, q_count * sizeof(double));
    if (rc) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, q_count * sizeof(double)))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;
    memcpy(ctx->pinned, q, q_count * sizeof(double));
    AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_SQ_A].p, ctx->pinned, q_count * sizeof(double), cudaMemcpyHostToDevice, stream));
    if ((rc = pinned_device_release(ctx, stream))) return rc;
    const double* d_q = reinterpret_cast<const double*>(ctx->bufs[BUF_SQ_A].p);

    const size_t smem_f64 = sizeof(double2) * kF64Stage * kF64Row * 2;
    const size_t smem_c64 = 2 * sizeof(double2) * (kC64Stage + 1) * (kC64RowA + kC64RowB);
    if (accum == AMEPB_ACCUM_C64_REFERENCE) {
        // frames per launch such that the chunk sums of the split kernel fit their buffer
        const size_t nchunks = (size_t)((n + chunksize - 1) / chunksize);
        const size_t per_frame = sizeof(float4) * nchunks * nq * nq;
        if (nchunks >= 4 && per_frame <= kC64ChunkBufferBytes)
            fb = std::min<int64_t>(fb, (int64_t)(kC64ChunkBufferBytes / per_frame));
    }
    if (!ctx->sq_attr_sf2d) {
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_f64_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_f64_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_f64_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_f64_kernel<double, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_c64_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_c64_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_c64_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c64));
        AMEPB_CUDA(ctx, cudaFuncSetAttribute(sf2d_c64_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c64));
        ctx->sq_attr_sf2d = true;
    }
    // float64 mode: bound the split-K scratch
    int ksplit = 1;
    if (accum == AMEPB_ACCUM_F64) {
        const int tiles = (nq + 63) / 64;
        const size_t per_split = (size_t)4 * nq * nq * sizeof(double);
        fb = std::min<int64_t>(fb, std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / per_split)));
        const long long blocks_fd = (long long)tiles * tiles * std::min<int64_t>(fb, nframes);
        // one block per SM (register bound): pick the particle split whose block count fills
        // whole waves of SMs best, preferring 2-6 waves
        const long long max_split = std::max<long long>(1, std::min<long long>(64, (n + 8 * kF64Stage - 1) / (8 * kF64Stage)));
        double best_eff = -1.0;
        for (long long k = 1; k <= max_split; ++k) {
            const long long blocks = blocks_fd * k;
            const long long waves = (blocks + ctx->sm_count - 1) / ctx->sm_count;
            double eff = (double)blocks / (double)(waves * ctx->sm_count);
            if (waves > 8) eff *= 0.98;  // more splits = more partial sums to combine
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                ksplit = (int)k;
            }
            if (waves >= 6 && eff > 0.95) break;
        }
        while (ksplit > 1 && (size_t)ksplit * per_split * std::min<int64_t>(fb, nframes) > ((size_t)4 << 30)) --ksplit;
    }

    for (int64_t f0 = 0; f0 < nframes; f0 += fb) {
        const int nf = (int)std::min<int64_t>(fb, nframes - f0);
        const void* d_coords = (const char*)coords + (size_t)f0 * frame_bytes;
        char* d_out = (char*)out + (size_t)f0 * out_frame;
        if (space == AMEPB_HOST) {
            if ((rc = ensure_device(ctx, BUF_STAGE_A, frame_bytes * nf))) return rc;
            if ((rc = ensure_device(ctx, BUF_STAGE_OUT, out_frame * nf))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, d_coords, frame_bytes * nf, cudaMemcpyHostToDevice, stream));
            d_coords = ctx->bufs[BUF_STAGE_A].p;
            d_out = reinterpret_cast<char*>(ctx->bufs[BUF_STAGE_OUT].p);
        }
        const double* qf = d_q + (q_stride ? f0 * q_stride : 0);
        cudaEvent_t ev0 = timing_event(ctx, stream);
        if (accum == AMEPB_ACCUM_F64) {
            const int tiles = (nq + 63) / 64;
            const size_t part_bytes = (size_t)nf * ksplit * 4 * nq * nq * sizeof(double);
            if ((rc = ensure_device(ctx, BUF_SQ_B, part_bytes))) return rc;
            double* d_part = reinterpret_cast<double*>(ctx->bufs[BUF_SQ_B].p);
            dim3 grid((unsigned)(tiles * tiles), (unsigned)ksplit, (unsigned)nf);
            const char* un = getenv("AMEPB_SF2D_UNROLL");
            const int unroll = un ? atoi(un) : 2;
            if (dtype == AMEPB_F32 && unroll == 1)
                sf2d_f64_kernel<float, 1><<<grid, kF64Threads, smem_f64, stream>>>((const float*)d_coords, n, qf, q_stride, nq, ksplit, d_part, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            else if (dtype == AMEPB_F32 && unroll == 4)
                sf2d_f64_kernel<float, 4><<<grid, kF64Threads, smem_f64, stream>>>((const float*)d_coords, n, qf, q_stride, nq, ksplit, d_part, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            else if (dtype == AMEPB_F32)
                sf2d_f64_kernel<float, 2><<<grid, kF64Threads, smem_f64, stream>>>((const float*)d_coords, n, qf, q_stride, nq, ksplit, d_part, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            else
                sf2d_f64_kernel<double, 2><<<grid, kF64Threads, smem_f64, stream>>>((const double*)d_coords, n, qf, q_stride, nq, ksplit, d_part, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            AMEPB_LAUNCH_CHECK(ctx);
            {
                cudaEvent_t ev1 = timing_event(ctx, stream);
                if (ev0 && ev1) ctx->ev_sq.emplace_back(ev0, ev1);
            }
            const long long total = (long long)nf * nq * nq;
            sf2d_f64_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_part, ksplit, nq, nf, reinterpret_cast<double2*>(d_out), ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            AMEPB_LAUNCH_CHECK(ctx);
        } else {
            const int tiles_a = (nq + kC64TileA - 1) / kC64TileA, tiles_b = (nq + kC64TileB - 1) / kC64TileB;
            // split the particles of every tile by whole chunks over up to 8 blocks (see SPLIT): measured
            // per-frame time for 1 ... 12 frames of configs[3] is lowest with 8 splits throughout
            // (28.7 -> 20.4 ms for a single frame, 20.5 -> 19.7 ms for 8), tools/sf2d_split_sweep.py
            const long long nchunks = (n + chunksize - 1) / chunksize;
            long long ksplit_c = 1;
            if (nchunks >= 4 && getenv("AMEPB_SF2D_NO_SPLIT") == nullptr &&
                sizeof(float4) * (size_t)nf * nchunks * nq * nq <= kC64ChunkBufferBytes)
                ksplit_c = std::min<long long>(8, nchunks / 2);
            if (const char* ks = getenv("AMEPB_SF2D_SPLIT")) ksplit_c = std::max<long long>(1, std::min<long long>(atoi(ks), nchunks / 2));
            if (ksplit_c > 1) {
                if ((rc = ensure_device(ctx, BUF_SQ_B, sizeof(float4) * (size_t)nf * nchunks * nq * nq))) return rc;
                float4* d_chunks = reinterpret_cast<float4*>(ctx->bufs[BUF_SQ_B].p);
                dim3 grid((unsigned)(tiles_a * tiles_b), (unsigned)nf, (unsigned)ksplit_c);
                if (dtype == AMEPB_F32)
                    sf2d_c64_kernel<float, true><<<grid, kC64Threads, smem_c64, stream>>>((const float*)d_coords, n, qf, q_stride, nq, chunksize, reinterpret_cast<float2*>(d_out), d_chunks, nchunks, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
                else
                    sf2d_c64_kernel<double, true><<<grid, kC64Threads, smem_c64, stream>>>((const double*)d_coords, n, qf, q_stride, nq, chunksize, reinterpret_cast<float2*>(d_out), d_chunks, nchunks, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
                AMEPB_LAUNCH_CHECK(ctx);
                const long long total = (long long)nf * nq * nq;
                sf2d_c64_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_chunks, nchunks, nq, nf, reinterpret_cast<float2*>(d_out), ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            } else {
                dim3 grid((unsigned)(tiles_a * tiles_b), (unsigned)nf);
                if (ctx->sf2d_shard_count > 1)   // rows of other shares stay zero
                    AMEPB_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(float2) * (size_t)nf * 4 * nq * nq, stream));
                if (dtype == AMEPB_F32)
                    sf2d_c64_kernel<float, false><<<grid, kC64Threads, smem_c64, stream>>>((const float*)d_coords, n, qf, q_stride, nq, chunksize, reinterpret_cast<float2*>(d_out), nullptr, nchunks, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
                else
                    sf2d_c64_kernel<double, false><<<grid, kC64Threads, smem_c64, stream>>>((const double*)d_coords, n, qf, q_stride, nq, chunksize, reinterpret_cast<float2*>(d_out), nullptr, nchunks, ctx->sf2d_shard_index, ctx->sf2d_shard_count);
            }
            AMEPB_LAUNCH_CHECK(ctx);
            cudaEvent_t ev1 = timing_event(ctx, stream);
            if (ev0 && ev1) ctx->ev_sq.emplace_back(ev0, ev1);
        }
        if (space == AMEPB_HOST) {
            AMEPB_CUDA(ctx, cudaMemcpyAsync((char*)out + (size_t)f0 * out_frame, d_out, out_frame * nf, cudaMemcpyDeviceToHost, stream));
            AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
        }
    }
    return 0;
}

int amepb_sq_debye(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other, int other_dtype,
                   int64_t m, int space, int64_t nframes, const double* q, int64_t q_stride, int32_t nq, int twod,
                   double* out, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    auto ok_dtype = [](int d) { return d == AMEPB_F32 || d == AMEPB_F64; };
    if (!coords || !q || !out || nframes < 0 || n <= 0 || nq <= 0 || nq > 4096 || !ok_dtype(dtype) ||
        (other && (!ok_dtype(other_dtype) || m <= 0)) || (space != AMEPB_DEVICE && space != AMEPB_HOST) ||
        (q_stride != 0 && q_stride < nq) || nframes > 65535)
        return fail(ctx, AMEPB_E_INVAL, "amepb_sq_debye: invalid argument");
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    timing_reset(ctx);
    if (!other) {
        other = coords;
        other_dtype = dtype;
        m = n;
    }
    const size_t abytes = (size_t)n * 3 * elem_bytes(dtype) * nframes;
    const size_t bbytes = (size_t)m * 3 * elem_bytes(other_dtype) * nframes;
    const void* d_a = coords;
    const void* d_b = other;
    double* d_out = out;
    int rc;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, abytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, abytes, cudaMemcpyHostToDevice, stream));
        d_a = ctx->bufs[BUF_STAGE_A].p;
        if (other == coords) {
            d_b = d_a;
        } else {
            if ((rc = ensure_device(ctx, BUF_STAGE_B, bbytes))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_B].p, other, bbytes, cudaMemcpyHostToDevice, stream));
            d_b = ctx->bufs[BUF_STAGE_B].p;
        }
        if ((rc = ensure_device(ctx, BUF_STAGE_OUT, sizeof(double) * (size_t)nframes * nq))) return rc;
        d_out = reinterpret_cast<double*>(ctx->bufs[BUF_STAGE_OUT].p);
    }
    const size_t q_count = (size_t)(q_stride ? nframes * q_stride : nq);
    if ((rc = ensure_pinned(ctx, q_count * sizeof(double)))) return rc;
    if ((rc = ensure_device(ctx, BUF_SQ_A, q_count * sizeof(double)))) return rc;
    if ((rc = pinned_host_acquire(ctx))) return rc;
    memcpy(ctx->pinned, q, q_count * sizeof(double));
    AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_SQ_A].p, ctx->pinned, q_count * sizeof(double), cudaMemcpyHostToDevice, stream));
    if ((rc = pinned_device_release(ctx, stream))) return rc;
    const double* d_q = reinterpret_cast<const double*>(ctx->bufs[BUF_SQ_A].p);

    const long long tiles = ((n + kDebyeThreads - 1) / kDebyeThreads) * ((m + kDebyeTileB - 1) / kDebyeTileB);
    const int nblocks = (int)std::max<long long>(1, std::min<long long>(tiles, std::max<long long>(1, (long long)ctx->sm_count * 8 / nframes)));
    if ((rc = ensure_device(ctx, BUF_SQ_B, sizeof(double) * (size_t)nframes * nblocks * nq))) return rc;
    double* d_part = reinterpret_cast<double*>(ctx->bufs[BUF_SQ_B].p);
    const size_t smem = sizeof(double) * (size_t)(kDebyeThreads / 32) * nq;
    dim3 grid((unsigned)nblocks, (unsigned)nframes);
    cudaEvent_t ev0 = timing_event(ctx, stream);
#define AMEPB_DEBYE(TA, TB)                                                                                       \
    do {                                                                                                          \
        if (twod)                                                                                                 \
            debye_kernel<TA, TB, true><<<grid, kDebyeThreads, smem, stream>>>((const TA*)d_a, n, (const TB*)d_b, m, d_q, q_stride, nq, d_part); \
        else                                                                                                      \
            debye_kernel<TA, TB, false><<<grid, kDebyeThreads, smem, stream>>>((const TA*)d_a, n, (const TB*)d_b, m, d_q, q_stride, nq, d_part); \
    } while (0)
    if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32) AMEPB_DEBYE(float, float);
    else if (dtype == AMEPB_F32) AMEPB_DEBYE(float, double);
    else if (other_dtype == AMEPB_F32) AMEPB_DEBYE(double, float);
    else AMEPB_DEBYE(double, double);
#undef AMEPB_DEBYE
    AMEPB_LAUNCH_CHECK(ctx);
    cudaEvent_t ev1 = timing_event(ctx, stream);
    if (ev0 && ev1) ctx->ev_sq.emplace_back(ev0, ev1);
    const long long total = (long long)nframes * nq;
    debye_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_part, nblocks, nq, nframes, d_out);
    AMEPB_LAUNCH_CHECK(ctx);
    if (space == AMEPB_HOST) {
        AMEPB_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)nframes * nq, cudaMemcpyDeviceToHost, stream));
        AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    return 0;
}

int amepb_sf2d_set_shard(amepb_ctx* ctx, int index, int count) {
    if (!ctx || count < 1 || index < 0 || index >= count) return AMEPB_E_INVAL;
    ctx->sf2d_shard_index = index;
    ctx->sf2d_shard_count = count;
    return 0;
}

namespace {
// NumPy's pairwise summation tree over n elements (numpy/_core/src/umath/loops_utils.h.src,
// @TYPE@_pairwise_sum): blocks of <= 128 elements are leaves, larger ranges split at
// n2 = n/2 - (n/2) % 8.  Emits the leaves in order and the postfix program that combines them.
void pairwise_tree(long long start, long long n, std::vector<long long>& gstart, std::vector<int>& glen, std::vector<int>& prog) {
    if (n <= kRefLeaf) {
        prog.push_back((int)gstart.size());
        gstart.push_back(start);
        glen.push_back((int)n);
        return;
    }
    long long n2 = n / 2;
    n2 -= n2 % 8;
    pairwise_tree(start, n2, gstart, glen, prog);
    pairwise_tree(start + n2, n - n2, gstart, glen, prog);
    prog.push_back(-1);
}
}  // namespace

int amepb_sq_debye_ref(amepb_ctx* ctx, const void* coords, int dtype, int64_t n, const void* other, int other_dtype,
                       int64_t m, int space, int64_t nframes, const double* q, int64_t q_stride, int32_t nq, int twod,
                       int64_t chunksize, float* out, void* stream_) {
    if (!ctx) return AMEPB_E_INVAL;
    auto ok_dtype = [](int d) { return d == AMEPB_F32 || d == AMEPB_F64; };
    if (!coords || !q || !out || nframes < 0 || n <= 0 || nq <= 0 || nq > 65536 || chunksize <= 0 || !ok_dtype(dtype) ||
        (other && (!ok_dtype(other_dtype) || m <= 0)) || (space != AMEPB_DEVICE && space != AMEPB_HOST) ||
        (q_stride != 0 && q_stride < nq))
        return fail(ctx, AMEPB_E_INVAL, "amepb_sq_debye_ref: invalid argument");
    if (nframes == 0) return 0;
    DeviceGuard guard(ctx->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->last_stream = stream_;
    timing_reset(ctx);
    if (!other) {
        other = coords;
        other_dtype = dtype;
        m = n;
    }
    if ((double)n * (double)std::min<int64_t>(chunksize, m) >= 9.0e18)
        return fail(ctx, AMEPB_E_INVAL, "amepb_sq_debye_ref: chunk too large");
    const size_t abytes = (size_t)n * 3 * elem_bytes(dtype) * nframes;
    const size_t bbytes = (size_t)m * 3 * elem_bytes(other_dtype) * nframes;
    const char* d_a = reinterpret_cast<const char*>(coords);
    const char* d_b = reinterpret_cast<const char*>(other);
    int rc;
    if (space == AMEPB_HOST) {
        if ((rc = ensure_device(ctx, BUF_STAGE_A, abytes))) return rc;
        AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_A].p, coords, abytes, cudaMemcpyHostToDevice, stream));
        d_a = reinterpret_cast<const char*>(ctx->bufs[BUF_STAGE_A].p);
        if (other == coords) {
            d_b = d_a;
        } else {
            if ((rc = ensure_device(ctx, BUF_STAGE_B, bbytes))) return rc;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(ctx->bufs[BUF_STAGE_B].p, other, bbytes, cudaMemcpyHostToDevice, stream));
            d_b = reinterpret_cast<const char*>(ctx->bufs[BUF_STAGE_B].p);
        }
    }
    // device blocks: [S: nq floats][sums: nq floats][qk: nq floats][zero count: 8 bytes][zero list / zp]
    constexpr long long kZeroCap = 1ll << 22;
    const size_t head_bytes = ((size_t)3 * nq * sizeof(float) + 15) & ~(size_t)15;
    if ((rc = ensure_device(ctx, BUF_SQ_A, head_bytes + 16 + sizeof(long long) * (size_t)kZeroCap))) return rc;
    char* blk = reinterpret_cast<char*>(ctx->bufs[BUF_SQ_A].p);
    float* d_S = reinterpret_cast<float*>(blk);
    float* d_sums = d_S + nq;
    float* d_qk = d_sums + nq;
    unsigned long long* d_zcount = reinterpret_cast<unsigned long long*>(blk + head_bytes);
    long long* d_zlist = reinterpret_cast<long long*>(blk + head_bytes + 16);
    std::vector<float> qk(nq);
    std::vector<long long> zhost, gstart;
    std::vector<int> glen, prog;
    std::vector<float> s_host(nq);
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t ev0 = timing_event(ctx, stream);
    for (int64_t f = 0; f < nframes; ++f) {
        const double* qf = q + (q_stride ? f * q_stride : 0);
        for (int k = 0; k < nq; ++k) {
            // float32(q), in 3D divided by pi in float32 (amep/spatialcor.py:1366, 1395)
            const volatile float q32 = (float)qf[k];
            const volatile float qpi = q32 / 3.14159274f;
            qk[k] = twod ? q32 : qpi;
        }
        // (small synchronous upload; this entry point synchronises once per chunk anyway)
        AMEPB_CUDA(ctx, cudaMemcpyAsync(d_qk, qk.data(), sizeof(float) * nq, cudaMemcpyHostToDevice, stream));
        AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
        const char* fa = d_a + (size_t)f * n * 3 * elem_bytes(dtype);
        const char* fbase = d_b + (size_t)f * m * 3 * elem_bytes(other_dtype);
        int first = 1;
        for (int64_t c0 = 0; c0 < m; c0 += chunksize) {
            const long long clen = std::min<int64_t>(chunksize, m - c0);
            const char* fb = fbase + (size_t)c0 * 3 * elem_bytes(other_dtype);
            const long long total = (long long)n * clen;
            // ---- zero distances -> compaction table ----
            AMEPB_CUDA(ctx, cudaMemsetAsync(d_zcount, 0, 8, stream));
#define AMEPB_REF_ZERO(TA, TB) debye_ref_zero_kernel<TA, TB><<<blocks, 256, 0, stream>>>((const TA*)fa, n, (const TB*)fb, clen, d_zcount, d_zlist, kZeroCap)
            if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32) AMEPB_REF_ZERO(float, float);
            else if (dtype == AMEPB_F32) AMEPB_REF_ZERO(float, double);
            else if (other_dtype == AMEPB_F32) AMEPB_REF_ZERO(double, float);
            else AMEPB_REF_ZERO(double, double);
#undef AMEPB_REF_ZERO
            AMEPB_LAUNCH_CHECK(ctx);
            unsigned long long nz64 = 0;
            AMEPB_CUDA(ctx, cudaMemcpyAsync(&nz64, d_zcount, 8, cudaMemcpyDeviceToHost, stream));
            AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
            if ((long long)nz64 > kZeroCap)
                return fail(ctx, AMEPB_E_INVAL, "amepb_sq_debye_ref: more than 2^22 coincident pairs in one chunk");
            const int nz = (int)nz64;
            zhost.resize(nz);
            if (nz) {
                AMEPB_CUDA(ctx, cudaMemcpyAsync(zhost.data(), d_zlist, sizeof(long long) * nz, cudaMemcpyDeviceToHost, stream));
                AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
                std::sort(zhost.begin(), zhost.end());
                for (int k = 0; k < nz; ++k) zhost[k] -= k;
                AMEPB_CUDA(ctx, cudaMemcpyAsync(d_zlist, zhost.data(), sizeof(long long) * nz, cudaMemcpyHostToDevice, stream));
                AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));   // zhost is reused by the next chunk
            }
            const long long nn = total - nz;   // non-zero distances
            // ---- the reference's branch: one 2D array of terms, or one pairwise sum per q ----
            const bool rows = !((double)(4 * nn) * (double)nq / 1e9 > 1.0);
            gstart.clear();
            glen.clear();
            prog.clear();
            if (rows) {
                for (long long st = 0; st < nn; st += kRefLeaf) {
                    gstart.push_back(st);
                    glen.push_back((int)std::min<long long>(kRefLeaf, nn - st));
                }
            } else {
                pairwise_tree(0, nn, gstart, glen, prog);
            }
            const long long ngroups = (long long)gstart.size();
            const size_t tab_bytes = (sizeof(long long) + sizeof(int)) * (size_t)ngroups + sizeof(int) * prog.size() + 64;
            if ((rc = ensure_device(ctx, BUF_SQ_B, tab_bytes))) return rc;
            char* tab = reinterpret_cast<char*>(ctx->bufs[BUF_SQ_B].p);
            long long* d_gstart = reinterpret_cast<long long*>(tab);
            int* d_glen = reinterpret_cast<int*>(tab + sizeof(long long) * (size_t)ngroups);
            int* d_prog = d_glen + ngroups;
            if (ngroups) {
                AMEPB_CUDA(ctx, cudaMemcpyAsync(d_gstart, gstart.data(), sizeof(long long) * ngroups, cudaMemcpyHostToDevice, stream));
                AMEPB_CUDA(ctx, cudaMemcpyAsync(d_glen, glen.data(), sizeof(int) * ngroups, cudaMemcpyHostToDevice, stream));
            }
            if (!prog.empty())
                AMEPB_CUDA(ctx, cudaMemcpyAsync(d_prog, prog.data(), sizeof(int) * prog.size(), cudaMemcpyHostToDevice, stream));
            // q tiles so that the buffer of terms / leaf sums stays below 2 GB
            const long long rows_out = rows ? nn : ngroups;
            int qt = nq;
            while (qt > 32 && (double)rows_out * qt * 4.0 > 2.0e9) qt = (qt + 1) / 2;
            if ((rc = ensure_device(ctx, BUF_ORD_C, std::max<size_t>(16, sizeof(float) * (size_t)rows_out * qt)))) return rc;
            float* d_buf = reinterpret_cast<float*>(ctx->bufs[BUF_ORD_C].p);
            for (int q0 = 0; q0 < nq; q0 += qt) {
                const int nqt = std::min(qt, nq - q0);
                if (ngroups) {
                    const unsigned tb = (unsigned)std::min<long long>((ngroups + 3) / 4, (long long)ctx->sm_count * 16);
#define AMEPB_REF_TERMS(TA, TB)                                                                                                   \
    do {                                                                                                                          \
        if (twod && rows) debye_ref_terms_kernel<TA, TB, true, false><<<tb, 128, 0, stream>>>((const TA*)fa, (const TB*)fb, clen, d_zlist, nz, d_gstart, d_glen, ngroups, d_qk, q0, nqt, d_buf); \
        else if (twod) debye_ref_terms_kernel<TA, TB, true, true><<<tb, 128, 0, stream>>>((const TA*)fa, (const TB*)fb, clen, d_zlist, nz, d_gstart, d_glen, ngroups, d_qk, q0, nqt, d_buf); \
        else if (rows) debye_ref_terms_kernel<TA, TB, false, false><<<tb, 128, 0, stream>>>((const TA*)fa, (const TB*)fb, clen, d_zlist, nz, d_gstart, d_glen, ngroups, d_qk, q0, nqt, d_buf); \
        else debye_ref_terms_kernel<TA, TB, false, true><<<tb, 128, 0, stream>>>((const TA*)fa, (const TB*)fb, clen, d_zlist, nz, d_gstart, d_glen, ngroups, d_qk, q0, nqt, d_buf); \
    } while (0)
                    if (dtype == AMEPB_F32 && other_dtype == AMEPB_F32) AMEPB_REF_TERMS(float, float);
                    else if (dtype == AMEPB_F32) AMEPB_REF_TERMS(float, double);
                    else if (other_dtype == AMEPB_F32) AMEPB_REF_TERMS(double, float);
                    else AMEPB_REF_TERMS(double, double);
#undef AMEPB_REF_TERMS
                    AMEPB_LAUNCH_CHECK(ctx);
                }
                const unsigned qb = (unsigned)((nqt + 63) / 64);
                if (rows) debye_ref_rows_kernel<<<qb, 64, 0, stream>>>(d_buf, nn, nqt, d_sums);
                else debye_ref_tree_kernel<<<qb, 64, 0, stream>>>(d_buf, d_prog, (long long)prog.size(), nqt, d_sums);
                AMEPB_LAUNCH_CHECK(ctx);
                debye_ref_accum_kernel<<<qb, 64, 0, stream>>>(d_sums, (float)nz, q0, nqt, first, d_S);
                AMEPB_LAUNCH_CHECK(ctx);
            }
            // (the host tables of this chunk are rebuilt for the next one: wait for their uploads)
            AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
            first = 0;
        }
        if (space == AMEPB_HOST) {
            AMEPB_CUDA(ctx, cudaMemcpyAsync(out + (size_t)f * nq, d_S, sizeof(float) * nq, cudaMemcpyDeviceToHost, stream));
        } else {
            AMEPB_CUDA(ctx, cudaMemcpyAsync(out + (size_t)f * nq, d_S, sizeof(float) * nq, cudaMemcpyDeviceToDevice, stream));
        }
        AMEPB_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    cudaEvent_t ev1 = timing_event(ctx, stream);
    if (ev0 && ev1) ctx->ev_sq.emplace_back(ev0, ev1);
    return 0;
}

}  // extern "C"

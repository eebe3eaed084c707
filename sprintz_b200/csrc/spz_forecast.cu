// spz_forecast.cu -- the stand-alone forecaster transforms:
//   encode/decode_{delta,doubledelta}_rowmajor_{8b,16b}   (delta.h:17-67, delta.cpp:34-744)
// Same-size in/out, with x[i] = 0 for i < 0 and all arithmetic mod 2^w:
//   delta       : y[i] = x[i] - x[i-D]                            x[i] = y[i] + x[i-D]
//   doubledelta : y[i] = (x[i] - x[i-D]) - (x[i-D] - x[i-2D])      d[i] = d[i-D] + y[i],  x[i] = x[i-D] + d[i]
// (the reference's AVX2 block path and its serial tail compute exactly this; pinned in tests/test_forecast.py).
//
// encode is element-wise: one thread per 16 output bytes, packed (SIMD-within-a-register) subtraction, the lagged
//        operands come from aligned word loads + a byte funnel shift.  HBM bound: N read + N written.
// decode is a per-dim prefix sum along rows -- here a single-pass chained scan (decoupled look-back): thread
//        (tile, dim) sums its tile of rows, publishes the aggregate, looks back over the earlier tiles of its dim
//        for the running state, then re-walks the tile (now L1/L2 resident) producing the output.  The state of
//        the double-delta recurrence is the pair (delta, value); a tile of n rows maps (d0, x0) to
//        (d0 + S1, x0 + n*d0 + S2) with S1 = sum y_k, S2 = sum (n-k)*y_k, which composes associatively.
#include "spz_host.h"
#include "spz_swar.cuh"
#include "spz_xff_forecast.cuh"
#include "../../include/sprintz_b200.h"

namespace spz {

namespace {

// word at byte offset `off` (any alignment) of `base` (4-byte aligned), zero before the buffer
__device__ __forceinline__ uint32_t word_at(const uint8_t* base, int64_t off)
{
    if (off <= -4) return 0u;
    const int64_t a = off & ~(int64_t)3;
    const uint32_t sh = (uint32_t)(off & 3) * 8u;
    const uint32_t w0 = a >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(base + a)) : 0u;
    if (sh == 0) return w0;
    const uint32_t w1 = a + 4 >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(base + a + 4)) : 0u;
    return __funnelshift_r(w0, w1, sh);
}

constexpr int kFcThreads = 256;

// y = transform(x): thread per 16 bytes. `bytes` = len * esz; src / dst 16-byte aligned (the tail word(s) are
// handled with byte stores).  Reads may touch up to 3 bytes past `bytes` inside the last aligned word (caller pads).
template <int W, int ORDER>
__global__ void __launch_bounds__(kFcThreads) spz_forecast_encode_kernel(const uint8_t* __restrict__ src,
                                                                         uint8_t* __restrict__ dst, uint64_t bytes,
                                                                         uint64_t lag_bytes)
{
    const uint64_t v = (uint64_t)blockIdx.x * kFcThreads + threadIdx.x;   // 16-byte unit
    const uint64_t o = v * 16;
    if (o >= bytes) return;
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src + o));
    const uint32_t xs[4] = {x0.x, x0.y, x0.z, x0.w};
    uint32_t y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t ob = (int64_t)o + 4 * k;
        const uint32_t x1 = word_at(src, ob - (int64_t)lag_bytes);
        uint32_t r = psub<W>(xs[k], x1);
        if (ORDER == 2) {
            const uint32_t x2 = word_at(src, ob - 2 * (int64_t)lag_bytes);
            r = psub<W>(r, psub<W>(x1, x2));
        }
        y[k] = r;
    }
    if (o + 16 <= bytes) {
        *reinterpret_cast<uint4*>(dst + o) = make_uint4(y[0], y[1], y[2], y[3]);
    } else {
        for (uint64_t b = o; b < bytes; b++) dst[b] = (uint8_t)(y[(b - o) >> 2] >> (8 * ((b - o) & 3)));
    }
}

constexpr unsigned long long kFcAgg = 1ull << 62;
constexpr unsigned long long kFcPrefix = 2ull << 62;

// x = inverse transform(y): thread (tile, dim); `state` has ntiles * D zeroed words.
template <int W, int ORDER>
__global__ void __launch_bounds__(kFcThreads) spz_forecast_decode_kernel(const uint8_t* __restrict__ src,
                                                                         uint8_t* __restrict__ dst, uint64_t len,
                                                                         uint32_t D, uint32_t tile_rows, uint64_t ntiles,
                                                                         unsigned long long* state, unsigned int* ticket)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr uint32_t WM = (1u << W) - 1u;
    // Tiles are handed out by ticket, not by blockIdx: a look-back only ever waits for tiles whose CTA has already
    // started, whatever order the hardware (or a serialising tool) schedules the blocks in.
    __shared__ unsigned int s_vb;
    if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint64_t g = (uint64_t)s_vb * kFcThreads + threadIdx.x;
    const uint64_t tile = g / D;
    const uint32_t d = (uint32_t)(g % D);
    if (tile >= ntiles) return;
    const uint_t* __restrict__ y = reinterpret_cast<const uint_t*>(src);
    uint_t* __restrict__ x = reinterpret_cast<uint_t*>(dst);
    const uint64_t row0 = tile * tile_rows;
    // rows of this tile that exist for dim d (the last row of the stream may be ragged)
    const uint64_t rows_d = (len > d) ? (len - d + D - 1) / D : 0;
    const uint32_t n = row0 >= rows_d ? 0u : (uint32_t)((rows_d - row0 < tile_rows) ? (rows_d - row0) : tile_rows);

    // ---- pass A: tile aggregate
    uint32_t s1 = 0, s2 = 0;
    {
        const uint_t* p = y + row0 * D + d;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t v = p[(uint64_t)k * D];
            s1 += v;
            if (ORDER == 2) s2 += (n - k) * v;
        }
    }
    // ---- chained scan over the tiles of this dim.  Every thread publishes BEFORE anyone looks back (tile 0's
    // aggregate is already its inclusive prefix), so a look-back never spins on a thread of its own warp.
    uint32_t dstart = 0, xstart = 0;                // state before the tile's first row (delta, value)
    unsigned long long* st = state + d;             // state[tile * D + d]
    const uint32_t lo0 = (ORDER == 2) ? s2 : s1;
    atomicExch(&st[tile * D], (tile == 0 ? kFcPrefix : kFcAgg) | ((unsigned long long)(s1 & WM) << 16) | (lo0 & WM));
    if (tile != 0) {
        uint32_t r1 = 0, r2 = 0, nr = 0;           // running aggregate over tiles (j, tile): S1, S2, rows
        for (long long j = (long long)tile - 1;; j--) {
            unsigned long long w;
            do { w = atomicAdd(&st[(uint64_t)j * D], 0ull); } while ((w >> 62) == 0);   // earlier tiles are already running
            const uint32_t hi = (uint32_t)(w >> 16) & WM, lo = (uint32_t)w & WM;
            if ((w >> 62) == 2) {                  // inclusive prefix of tile j: (delta_end, x_end)
                if (ORDER == 2) { dstart = hi + r1; xstart = lo + nr * hi + r2; }
                else            { xstart = lo + r1; }
                break;
            }
            // prepend tile j (always a full tile: only the last tile of a dim can be short)
            if (ORDER == 2) r2 = lo + nr * hi + r2;
            r1 += hi;
            nr += tile_rows;
        }
        const uint32_t dend = dstart + s1;
        const uint32_t xend = (ORDER == 2) ? (xstart + n * dstart + s2) : (xstart + s1);
        atomicExch(&st[tile * D], kFcPrefix | ((unsigned long long)(dend & WM) << 16) | (xend & WM));
    }
    // ---- pass B: produce the tile (second read of y: L1 / L2 resident)
    {
        const uint_t* p = y + row0 * D + d;
        uint_t* q = x + row0 * D + d;
        uint32_t dl = dstart, xv = xstart;
        for (uint32_t k = 0; k < n; k++) {
            const uint32_t v = p[(uint64_t)k * D];
            if (ORDER == 2) { dl += v; xv += dl; } else { xv += v; }
            q[(uint64_t)k * D] = (uint_t)xv;
        }
    }
}

}  // namespace

// Stand-alone FIRE transform (predict.h:15-30): the predictor state of a dim runs through the WHOLE stream and is not
// associative, so one stream offers ndims-way parallelism only -- one thread per dim, the per-dim code of
// spz_xff_forecast.cuh (pinned on the CPU by tests/test_xff_forecast_host.py).  A correctness path: long streams belong
// in the chunked frame codec, which is this transform plus bit packing at full parallelism.
template <int W, bool ENC>
__global__ void __launch_bounds__(128) spz_xff_forecast_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               uint64_t len, uint32_t D, int32_t nblocks)
{
    const uint32_t d = blockIdx.x * 128u + threadIdx.x;
    if (d < D) xffp_dim<W, ENC>(src, dst, len, D, d, nblocks);
}

cudaError_t launch_xff_forecast(int enc, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                cudaStream_t stream)
{
    if (len == 0 || D == 0) return cudaSuccess;
    const int32_t nb = xffp_nblocks(len, D, elem_sz);
    const unsigned grid = (D + 127u) / 128u;
    if (elem_sz == 1) {
        count_launch();
        if (enc) spz_xff_forecast_kernel<8, true><<<grid, 128, 0, stream>>>(src, dst, len, D, nb);
        else     spz_xff_forecast_kernel<8, false><<<grid, 128, 0, stream>>>(src, dst, len, D, nb);
    } else {
        count_launch();
        if (enc) spz_xff_forecast_kernel<16, true><<<grid, 128, 0, stream>>>(src, dst, len, D, nb);
        else     spz_xff_forecast_kernel<16, false><<<grid, 128, 0, stream>>>(src, dst, len, D, nb);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------ delta decode, vectorised (rows of 8, 16 or k*16 bytes)
// The scalar kernel above walks one element per step (a byte or a halfword per load): 0.13-0.25 of the copy bandwidth.
// When a row is 8 bytes, 16 bytes or a multiple of 16 bytes, a 16-byte unit holds the SAME dims in every row (or two whole
// rows), so the prefix sum along rows is a packed add on whole units: thread (tile, unit column) sums its tile of 16 or 64
// units, the CTA scans its tiles and chains to the earlier CTAs, then every thread re-walks its tile (L2 resident: 256 B
// per thread) and stores whole units.
constexpr uint32_t kVecSteps = 16;                    // units per tile, at least (the launcher picks 16 or 64)
struct VecState { uint4 agg; uint4 pre; };             // per (tile, column): tile aggregate, inclusive prefix

template <int W> __device__ __forceinline__ uint4 padd4(uint4 a, uint4 b)
{
    return make_uint4(padd<W>(a.x, b.x), padd<W>(a.y, b.y), padd<W>(a.z, b.z), padd<W>(a.w, b.w));
}

// ROW8: a row is 8 bytes, a unit holds two rows (state and aggregates are kept in both halves).
// A CTA owns TPC consecutive tiles of every unit column (thread = (tile, column)): tile aggregates are scanned inside the
// CTA in shared memory, ONE thread per column chains the CTA's total to the earlier CTAs (decoupled look-back over CTA
// records: a CTA moves 256 KiB, so the chain is short and its predecessors have usually published their prefix).
template <int W, bool ROW8>
__global__ void __launch_bounds__(kFcThreads) spz_delta_decode_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                                          uint64_t nsteps, uint32_t M, uint32_t TPC, uint32_t S, uint64_t ntiles,
                                                                          VecState* vs, unsigned int* flags, unsigned int* ticket)
{
    __shared__ unsigned int s_vb;
    __shared__ uint4 sh[kFcThreads];                    // inclusive scan of the tile aggregates, [tile in CTA][column]
    __shared__ uint4 sp[kFcThreads];                    // per column: values of the row before the CTA's first tile
    if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1u); // CTAs by ticket: a look-back only waits for running CTAs
    __syncthreads();
    const uint32_t cta = s_vb;
    const uint32_t tl = threadIdx.x / M, col = threadIdx.x % M;
    const uint64_t tile = (uint64_t)cta * TPC + tl;
    const bool active = tl < TPC && tile < ntiles;
    const uint64_t s0 = tile * S;
    const uint32_t n = active ? (uint32_t)(nsteps - s0 < S ? nsteps - s0 : S) : 0u;
    const uint4* p = src + s0 * M + col;
    // ---- pass A: tile aggregate
    uint4 A = make_uint4(0, 0, 0, 0);
    for (uint32_t k = 0; k < n; k++) A = padd4<W>(A, __ldcg(p + (uint64_t)k * M));
    if (ROW8) { const uint2 h = make_uint2(padd<W>(A.x, A.z), padd<W>(A.y, A.w)); A = make_uint4(h.x, h.y, h.x, h.y); }
    // ---- inclusive scan over the CTA's tiles of my column (neighbours are M threads apart)
    sh[threadIdx.x] = A;
    __syncthreads();
    for (uint32_t off = 1; off < TPC; off <<= 1) {
        uint4 v = make_uint4(0, 0, 0, 0);
        const bool take = tl < TPC && tl >= off;
        if (take) v = sh[threadIdx.x - off * M];
        __syncthreads();
        if (take) sh[threadIdx.x] = padd4<W>(sh[threadIdx.x], v);
        __syncthreads();
    }
    // ---- the last tile of a column holds the CTA's total: publish, look back, publish the inclusive prefix
    if (tl == TPC - 1) {
        const uint4 T = sh[threadIdx.x];
        const uint64_t e = (uint64_t)cta * M + col;
        uint4 X = make_uint4(0, 0, 0, 0);
        if (cta == 0) {
            vs[e].pre = T;
            __threadfence();
            atomicExch(&flags[e], 2u);
        } else {
            vs[e].agg = T;
            __threadfence();
            atomicExch(&flags[e], 1u);
            for (uint64_t j = cta - 1;; j--) {
                const uint64_t ej = j * M + col;
                unsigned int f;
                do { f = atomicAdd(&flags[ej], 0u); } while (f == 0u);
                __threadfence();
                const volatile uint4* q = (f == 2u) ? &vs[ej].pre : &vs[ej].agg;
                const uint4 t = make_uint4(q->x, q->y, q->z, q->w);
                X = padd4<W>(X, t);
                if (f == 2u) break;
            }
            vs[e].pre = padd4<W>(X, T);
            __threadfence();
            atomicExch(&flags[e], 2u);
        }
        sp[col] = X;
    }
    __syncthreads();
    // ---- pass B: the tile (second read: L2)
    if (n) {
        uint4 X = sp[col];
        if (tl > 0) X = padd4<W>(X, sh[threadIdx.x - M]);
        uint4* q = dst + s0 * M + col;
        for (uint32_t k = 0; k < n; k++) {
            uint4 v = __ldcg(p + (uint64_t)k * M);
            if (ROW8) { v.z = padd<W>(v.z, v.x); v.w = padd<W>(v.w, v.y); }     // the unit's second row on top of its first
            v = padd4<W>(v, X);
            q[(uint64_t)k * M] = v;
            X = ROW8 ? make_uint4(v.z, v.w, v.z, v.w) : v;
        }
    }
}

static uint32_t forecast_tile_rows(uint64_t len, uint32_t D)
{
    const uint64_t rows = (len + D - 1) / D;
    uint64_t want = rows * D / 262144;              // ~256 Ki (tile, dim) threads
    uint32_t tr = 64;
    while (tr < 4096 && tr < want) tr <<= 1;
    return tr;
}

uint64_t forecast_workspace_bytes(uint64_t len, uint32_t D)
{
    if (D == 0) return 0;
    const uint64_t rows = (len + D - 1) / D;
    const uint32_t tr = forecast_tile_rows(len, D);
    const uint64_t ntiles = (rows + tr - 1) / tr;
    const uint64_t scalar = 8 * (ntiles ? ntiles : 1) * D + 8;          // + the tile ticket
    // vector path (rows of 8 / 16 / k*16 bytes): a VecState + a flag per (tile of 64 units, column); sized for 2-byte elements
    const uint64_t units = (len * 2 + 15) / 16;
    const uint64_t entries = units / kVecSteps + 2ull * D + 64;
    const uint64_t vec = entries * (sizeof(VecState) + 4) + 64;
    return scalar > vec ? scalar : vec;
}

cudaError_t launch_forecast_encode(int order, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                   cudaStream_t stream)
{
    const uint64_t bytes = len * (uint64_t)elem_sz;
    if (bytes == 0) return cudaSuccess;
    const uint64_t units = (bytes + 15) / 16;
    const uint64_t grid = (units + kFcThreads - 1) / kFcThreads;
    if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
    const uint64_t lag = (uint64_t)D * elem_sz;
    count_launch();
#define SPZ_FE(Wv, Ov) spz_forecast_encode_kernel<Wv, Ov><<<(unsigned)grid, kFcThreads, 0, stream>>>(src, dst, bytes, lag)
    if (elem_sz == 1) { if (order == 1) SPZ_FE(8, 1); else SPZ_FE(8, 2); }
    else              { if (order == 1) SPZ_FE(16, 1); else SPZ_FE(16, 2); }
#undef SPZ_FE
    return cudaGetLastError();
}

cudaError_t launch_forecast_decode(int order, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                   unsigned long long* state, unsigned int* ticket, cudaStream_t stream)
{
    if (len == 0) return cudaSuccess;
    {   // vectorised delta decode: whole 16-byte units with fixed dims per byte position
        const uint64_t row_bytes = (uint64_t)D * elem_sz, bytes = len * (uint64_t)elem_sz;
        const bool row8 = row_bytes == 8;
        const bool rowk = row_bytes % 16 == 0;
        if (order == 1 && ticket == nullptr && (row8 || (rowk && row_bytes <= 2048)) && bytes % (row8 ? 16 : row_bytes) == 0 &&
            (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)state) & 15) == 0) {
            const uint32_t M = row8 ? 1u : (uint32_t)(row_bytes / 16);
            const uint32_t TPC = kFcThreads / M;                   // tiles of every column per CTA
            const uint64_t nsteps = bytes / 16 / M;
            // one unit column (rows of 8 / 16 bytes): short tiles keep the second pass in L2 (716 vs 656 GB/s); several columns:
            // longer tiles amortise the CTA's scan and look-back (1172 vs 876 GB/s at int16 x 16)
            const uint32_t S = M == 1 ? kVecSteps : 4 * kVecSteps;
            const uint64_t ntiles = (nsteps + S - 1) / S;
            const uint64_t nctas = (ntiles + TPC - 1) / TPC;
            const uint64_t entries = nctas * M;
            VecState* vs = reinterpret_cast<VecState*>(state);
            unsigned int* flags = reinterpret_cast<unsigned int*>(vs + entries);
            unsigned int* tk = flags + entries;
            cudaError_t e = cudaMemsetAsync(flags, 0, 4 * (entries + 1), stream);
            if (e != cudaSuccess) return e;
            const uint64_t grid = nctas;
            if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
            count_launch();
#define SPZ_FV(Wv, Rv) spz_delta_decode_vec_kernel<Wv, Rv><<<(unsigned)grid, kFcThreads, 0, stream>>>( \
                reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), nsteps, M, TPC, S, ntiles, vs, flags, tk)
            if (elem_sz == 1) { if (row8) SPZ_FV(8, true); else SPZ_FV(8, false); }
            else              { if (row8) SPZ_FV(16, true); else SPZ_FV(16, false); }
#undef SPZ_FV
            return cudaGetLastError();
        }
    }
    const uint64_t rows = (len + D - 1) / D;
    const uint32_t tr = forecast_tile_rows(len, D);
    const uint64_t ntiles = (rows + tr - 1) / tr;
    if (ticket == nullptr) ticket = reinterpret_cast<unsigned int*>(state + ntiles * D);     // last word of the workspace
    cudaError_t e = cudaMemsetAsync(state, 0, 8 * ntiles * D + 8, stream);
    if (e != cudaSuccess) return e;
    const uint64_t threads = ntiles * D;
    const uint64_t grid = (threads + kFcThreads - 1) / kFcThreads;
    if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
    count_launch();
#define SPZ_FD(Wv, Ov) spz_forecast_decode_kernel<Wv, Ov><<<(unsigned)grid, kFcThreads, 0, stream>>>(src, dst, len, D, tr, ntiles, state, ticket)
    if (elem_sz == 1) { if (order == 1) SPZ_FD(8, 1); else SPZ_FD(8, 2); }
    else              { if (order == 1) SPZ_FD(16, 1); else SPZ_FD(16, 2); }
#undef SPZ_FD
    return cudaGetLastError();
}

}  // namespace spz

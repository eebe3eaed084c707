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
    return 8 * (ntiles ? ntiles : 1) * D + 8;          // + the tile ticket
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

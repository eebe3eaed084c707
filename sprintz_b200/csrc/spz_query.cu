// spz_query.cu -- per-dim reductions over row-major elements, the "materialize" half of query-on-compressed
// (query.hpp:22-29: QueryParams{op = REDUCE_MAX | REDUCE_SUM, materialize}).
//
// The reference threads a query functor through its decoders (sprintz_delta_rle_query.cpp:149-192,
// sprintz_xff_rle_query.cpp:106-137) and is half-finished: the functor's state is never handed back to the caller
// (query.hpp:231).  Here the query has a result -- for every dim d, the unsigned maximum or the 64-bit sum of
// x[i], i % ndims == d, over ALL decoded elements -- and two ways of getting it:
//   * fused        spz_decode_lane_kernel<.., Q> (spz_decode_lane.cu): decoder lanes fold rows into registers instead
//                  of storing them; reads the compressed bytes only (frames of 8 / 16 dims, materialize == 0)
//   * materialized decode with the regular kernels, then the kernel below reduces the decoded rows (any shape)
//
// spz_reduce_rows_kernel: a thread owns one 16-byte unit per sweep of S units, S * (16 / esz) a multiple of ndims,
// so the dims of its 8 / 16 element positions never change and the accumulators are plain registers.  Threads fold
// into per-CTA shared arrays, CTAs into the result with one atomic per dim.  HBM bound: N bytes read.
#include "spz_host.h"

namespace spz {

namespace {

constexpr int kRedThreads = 256;

template <int W, int OP>
__global__ void __launch_bounds__(kRedThreads) spz_reduce_rows_kernel(const uint8_t* __restrict__ src, uint64_t len,
                                                                      uint32_t D, uint64_t S,
                                                                      unsigned long long* __restrict__ result)
{
    constexpr int ESZ = W / 8;
    constexpr int EPV = 16 / ESZ;
    extern __shared__ unsigned long long sh[];      // [D]
    for (uint32_t d = threadIdx.x; d < D; d += kRedThreads) sh[d] = 0ull;
    __syncthreads();

    const uint64_t g = (uint64_t)blockIdx.x * kRedThreads + threadIdx.x;
    const uint64_t units = (len + EPV - 1) / EPV;
    if (g < S) {
        // Packed accumulators: a 16-byte unit is 4 words; 8-bit elements are split into two words of 16-bit lanes
        // (even / odd bytes), so max is one VIMNMX.U16x2 and sum one IADD per two elements.  Lane sums are drained
        // into 32-bit (8-bit data, every 256 sweeps) / 64-bit totals before they can overflow.
        constexpr int NA = (W == 8) ? 8 : (OP == 1 ? 4 : 8);
        uint32_t acc[NA];
        uint32_t tot32[(W == 8 && OP == 2) ? 16 : 1];
        unsigned long long tot64[(W == 16 && OP == 2) ? 8 : 1];
#pragma unroll
        for (int k = 0; k < NA; k++) acc[k] = 0;
#pragma unroll
        for (int k = 0; k < (int)(sizeof(tot32) / 4); k++) tot32[k] = 0;
#pragma unroll
        for (int k = 0; k < (int)(sizeof(tot64) / 8); k++) tot64[k] = 0;
        auto drain = [&]() {
            if (W == 8 && OP == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) { tot32[2 * k] += acc[k] & 0xffffu; tot32[2 * k + 1] += acc[k] >> 16; acc[k] = 0; }
            }
            if (W == 16 && OP == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) { tot64[k] += acc[k]; acc[k] = 0; }
            }
        };
        uint32_t sweeps = 0;
        for (uint64_t u = g; u < units; u += S) {
            uint32_t w[4];
            if ((u + 1) * EPV <= len) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + u);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {                                // ragged last unit: missing elements read as 0 (identity of both ops)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    w[j] = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint64_t at = u * 16 + 4 * j + b;
                        if (at < len * ESZ) w[j] |= (uint32_t)src[at] << (8 * b);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (W == 8) {
                    const uint32_t ev = w[j] & 0x00ff00ffu, od = (w[j] >> 8) & 0x00ff00ffu;
                    if (OP == 1) { acc[2 * j] = __vmaxu2(acc[2 * j], ev); acc[2 * j + 1] = __vmaxu2(acc[2 * j + 1], od); }
                    else         { acc[2 * j] += ev; acc[2 * j + 1] += od; }
                } else {
                    if (OP == 1) acc[j] = __vmaxu2(acc[j], w[j]);
                    else         { acc[2 * j] += w[j] & 0xffffu; acc[2 * j + 1] += w[j] >> 16; }
                }
            }
            ++sweeps;
            if (OP == 2 && (sweeps & (W == 8 ? 0xffu : 0x7fffu)) == 0u) drain();     // 256 x 255 < 2^16; 32768 x 65535 < 2^32
        }
        drain();
        uint32_t d = (uint32_t)((g * EPV) % D);
#pragma unroll
        for (int k = 0; k < EPV; k++) {
            unsigned long long v;
            if (W == 8) {
                // element k: word k / 4, byte k % 4 -> even / odd split word, low / high lane
                const int a = 2 * (k / 4) + (k & 1), hi = (k >> 1) & 1;
                v = (OP == 1) ? ((acc[a] >> (16 * hi)) & 0xffffu) : tot32[2 * a + hi];
            } else {
                v = (OP == 1) ? ((acc[k / 2] >> (16 * (k & 1))) & 0xffffu) : tot64[k];
            }
            if (OP == 1) { if (v) atomicMax(&sh[d], v); } else { if (v) atomicAdd(&sh[d], v); }
            if (++d == D) d = 0;
        }
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < D; d += kRedThreads) {
        const unsigned long long v = sh[d];
        if (v) { if (OP == 1) atomicMax(result + d, v); else atomicAdd(result + d, v); }
    }
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; }

}  // namespace

// result[ndims] must be zeroed by the caller (0 is the identity of both ops). src 16-byte aligned.
cudaError_t launch_reduce_rows(int op, int elem_sz, const uint8_t* src, uint64_t len, uint32_t D,
                               unsigned long long* result, cudaStream_t stream)
{
    if (len == 0) return cudaSuccess;
    if (((uintptr_t)src & 15) != 0 || D == 0 || D > (uint32_t)kMaxDims) return cudaErrorInvalidValue;
    const uint32_t epv = 16u / (uint32_t)elem_sz;
    const uint64_t units = (len + epv - 1) / epv;
    const uint32_t m = D / gcd_u32(D, epv);         // S must be a multiple of m
    uint64_t grid = (units + kRedThreads - 1) / kRedThreads;
    const uint64_t max_grid = 148ull * 8ull;
    if (grid > max_grid) grid = max_grid;
    const uint64_t min_grid = (m + kRedThreads - 1) / kRedThreads;
    if (grid < min_grid) grid = min_grid;
    const uint64_t S = (grid * kRedThreads) / m * m;
    const size_t smem = 8ull * D;
    count_launch();
#define SPZ_RED(Wv, Ov) spz_reduce_rows_kernel<Wv, Ov><<<(unsigned)grid, kRedThreads, smem, stream>>>(src, len, D, S, result)
    if (elem_sz == 1) { if (op == 1) SPZ_RED(8, 1); else SPZ_RED(8, 2); }
    else              { if (op == 1) SPZ_RED(16, 1); else SPZ_RED(16, 2); }
#undef SPZ_RED
    return cudaGetLastError();
}

}  // namespace spz

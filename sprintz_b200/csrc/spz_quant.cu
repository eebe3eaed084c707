// spz_quant.cu -- float -> uint8 / uint16 quantiser.
//
// Restates python/datasets/compress_bench.py:45-60 (_quantize with axis = 0), the step that feeds the codecs:
//     mat -= min(mat, axis=0); mat = mat.astype(float32); mat /= max(1, max(mat, axis=0)); (mat * max_val).astype(dtype)
// Two HBM-bound passes: per-dim min / max (order-independent, so any reduction tree is exact), then the element-wise
// map with IEEE round-to-nearest float32 division and multiplication (no FMA contraction, no fast-math).
#include "spz_host.h"

namespace spz {

constexpr int kQThreads = 256;

// monotone float <-> uint encodings so min / max can use integer atomics
__device__ __forceinline__ unsigned long long enc_f64(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | (1ull << 63));
}
__device__ __forceinline__ double dec_f64(unsigned long long e)
{
    const unsigned long long b = (e >> 63) ? (e & ~(1ull << 63)) : ~e;
    return __longlong_as_double((long long)b);
}

// work: [D] encoded minima, [D] encoded maxima (u64 each)
__global__ void __launch_bounds__(kQThreads) spz_quant_init_kernel(unsigned long long* work, uint32_t D)
{
    const uint32_t i = blockIdx.x * kQThreads + threadIdx.x;
    if (i < D) { work[i] = ~0ull; work[D + i] = 0ull; }
}

// A thread walks elements i = g, g + S, ... with S a multiple of D, so its dim never changes.
template <typename T>
__global__ void __launch_bounds__(kQThreads) spz_quant_minmax_kernel(const T* __restrict__ src, uint64_t len, uint32_t D,
                                                                     uint64_t S, unsigned long long* work)
{
    const uint64_t g = (uint64_t)blockIdx.x * kQThreads + threadIdx.x;
    if (g >= S || g >= len) return;
    double lo = (double)src[g], hi = lo;
    for (uint64_t i = g + S; i < len; i += S) {
        const double v = (double)src[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    const uint32_t d = (uint32_t)(g % D);
    atomicMin(work + d, enc_f64(lo));
    atomicMax(work + D + d, enc_f64(hi));
}

template <typename T, typename U>
__global__ void __launch_bounds__(kQThreads) spz_quant_map_kernel(const T* __restrict__ src, uint64_t len, uint32_t D,
                                                                  const unsigned long long* __restrict__ work, U* __restrict__ dst)
{
    const uint64_t i = (uint64_t)blockIdx.x * kQThreads + threadIdx.x;
    if (i >= len) return;
    const uint32_t d = (uint32_t)(i % D);
    const T mn = (T)dec_f64(work[d]);
    const T mx = (T)dec_f64(work[D + d]);
    // mat -= min (input precision); astype(float32); the column maximum after the shift is float32(max - min)
    float y, top;
    if (sizeof(T) == 8) { y = __double2float_rn((double)src[i] - (double)mn); top = __double2float_rn((double)mx - (double)mn); }
    else                { y = __fsub_rn((float)src[i], (float)mn); top = __fsub_rn((float)mx, (float)mn); }
    const float den = top > 1.0f ? top : 1.0f;                   // max(1, column max)
    const float q = __fmul_rn(__fdiv_rn(y, den), sizeof(U) == 1 ? 255.0f : 65535.0f);
    dst[i] = (U)q;                                               // astype: truncation (values are >= 0)
}

uint64_t quantize_workspace_bytes(uint32_t D) { return 16ull * (D ? D : 1); }

cudaError_t launch_quantize(int elem_sz, int src_f64, const void* src, uint64_t len, uint32_t D, void* dst, void* work_,
                            cudaStream_t stream)
{
    if (len == 0) return cudaSuccess;
    unsigned long long* work = (unsigned long long*)work_;
    count_launch();
    spz_quant_init_kernel<<<(D + kQThreads - 1) / kQThreads, kQThreads, 0, stream>>>(work, D);
    // ~64 K threads in flight, stride a multiple of D
    uint64_t S = ((1ull << 16) + D - 1) / D * D;
    if (S > len) S = (len + D - 1) / D * D;
    const unsigned grid = (unsigned)((S + kQThreads - 1) / kQThreads);
    count_launch();
    if (src_f64) spz_quant_minmax_kernel<double><<<grid, kQThreads, 0, stream>>>((const double*)src, len, D, S, work);
    else         spz_quant_minmax_kernel<float><<<grid, kQThreads, 0, stream>>>((const float*)src, len, D, S, work);
    const uint64_t g2 = (len + kQThreads - 1) / kQThreads;
    if (g2 > 0x7fffffffull) return cudaErrorInvalidValue;
    count_launch();
    if (src_f64) {
        if (elem_sz == 1) spz_quant_map_kernel<double, uint8_t><<<(unsigned)g2, kQThreads, 0, stream>>>((const double*)src, len, D, work, (uint8_t*)dst);
        else              spz_quant_map_kernel<double, uint16_t><<<(unsigned)g2, kQThreads, 0, stream>>>((const double*)src, len, D, work, (uint16_t*)dst);
    } else {
        if (elem_sz == 1) spz_quant_map_kernel<float, uint8_t><<<(unsigned)g2, kQThreads, 0, stream>>>((const float*)src, len, D, work, (uint8_t*)dst);
        else              spz_quant_map_kernel<float, uint16_t><<<(unsigned)g2, kQThreads, 0, stream>>>((const float*)src, len, D, work, (uint16_t*)dst);
    }
    return cudaGetLastError();
}

}  // namespace spz

// spz_swar.cuh -- SIMD-within-a-register arithmetic on packed 8-bit / 16-bit lanes of a 32-bit word.
//
// Used by the delta / double-delta forecaster kernels (spz_forecast.cu: four int8 or two int16 elements per
// subtraction).  The per-byte zigzag and OR-reduction below are the pass-1 building blocks of the packed delta encoder
// planned for wide int8 rows (config 4 needs ~2 instead of ~10 instructions per sample; the DELTA
// predictor needs no multiply, so four dims fit one word end to end).  Everything here is __host__ __device__ and is
// checked exhaustively per lane on the CPU (tests/test_primitives_host.py).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SPZ_SWAR_HD __host__ __device__ __forceinline__
#else
#define SPZ_SWAR_HD inline
#endif

namespace spz {

// per-byte (W == 8) / per-halfword (W == 16) a - b, no borrow across lanes
template <int W> SPZ_SWAR_HD uint32_t psub(uint32_t a, uint32_t b)
{
    constexpr uint32_t H = (W == 8) ? 0x80808080u : 0x80008000u;
    return ((a | H) - (b & ~H)) ^ ((a ^ ~b) & H);
}

// per-lane a + b, no carry across lanes
template <int W> SPZ_SWAR_HD uint32_t padd(uint32_t a, uint32_t b)
{
    constexpr uint32_t H = (W == 8) ? 0x80808080u : 0x80008000u;
    return ((a & ~H) + (b & ~H)) ^ ((a ^ b) & H);
}

// per-lane zigzag (bitpack.h:302): (x << 1) ^ (x >> (W-1)), each lane a signed W-bit integer
template <int W> SPZ_SWAR_HD uint32_t pzigzag(uint32_t x)
{
    constexpr uint32_t H = (W == 8) ? 0x80808080u : 0x80008000u;
    const uint32_t sign = (x & H) >> (W - 1);                 // 1 in every negative lane
    const uint32_t fill = sign * ((1u << W) - 1u);            // all ones in every negative lane (no carries: lanes are disjoint)
    return ((x & ~H) << 1) ^ fill;                            // & ~H drops the sign bit before it could shift into the next lane
}

// per-lane inverse zigzag (bitpack.h:303): (u >> 1) ^ -(u & 1)
template <int W> SPZ_SWAR_HD uint32_t punzigzag(uint32_t u)
{
    constexpr uint32_t L = (W == 8) ? 0x01010101u : 0x00010001u;
    constexpr uint32_t H = (W == 8) ? 0x80808080u : 0x80008000u;
    const uint32_t fill = (u & L) * ((1u << W) - 1u);
    return ((u >> 1) & ~H) ^ fill;
}

}  // namespace spz

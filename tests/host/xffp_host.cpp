// Host build of sprintz_b200/csrc/spz_xff_forecast.cuh: the SAME per-dim code the CUDA kernel runs (one thread per dim),
// looped over the dims on the CPU so that the CPU test suite can pin it to the reference-checked restatement before any
// GPU is involved.  Built by tests/test_xff_forecast_host.py with g++; test infrastructure, not product.
#include "../../sprintz_b200/csrc/spz_xff_forecast.cuh"

extern "C" int64_t host_xffp(int enc, int elem_sz, const void* src, uint64_t len, void* dst, uint32_t ndims)
{
    if (ndims == 0) return enc ? -1 : 0;
    const int32_t nb = spz::xffp_nblocks(len, ndims, elem_sz);
    for (uint32_t d = 0; d < ndims; d++) {
        if (elem_sz == 1) { if (enc) spz::xffp_dim<8, true>(src, dst, len, ndims, d, nb); else spz::xffp_dim<8, false>(src, dst, len, ndims, d, nb); }
        else              { if (enc) spz::xffp_dim<16, true>(src, dst, len, ndims, d, nb); else spz::xffp_dim<16, false>(src, dst, len, ndims, d, nb); }
    }
    return (int64_t)len;
}

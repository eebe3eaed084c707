// spz_xff_forecast.cuh -- per-dim arithmetic of the stand-alone FIRE transform
//   encode/decode_xff_rowmajor_{8b,16b}   (predict.h:15-30, predict.cpp:56-527)
// shared by the CUDA kernel (spz_forecast.cu) and a host build of the same code that the CPU tests compile with g++
// and compare with the reference-pinned restatement (tests/test_xff_forecast_host.py): the GPU kernel is exactly this
// code with one thread per dim.
//
// The transform is an older prototype of FIRE whose arithmetic differs from the codec's (sprintz_xff_rle.cpp):
//   blocks    nblocks = (len / ndims) / 8, reduced so the reference's 32-byte vector stores stay inside the buffer
//             (predict.cpp:99-106); everything after the block part is plain delta coding with lag ndims (:268-279),
//             and when there is no block at all the first ndims elements are copied.
//   8-bit     counter int16, coef = (counter >> 5) << 4; an EVEN dim multiplies the previous delta as an UNSIGNED byte
//             (:163-164), an odd dim as a signed one; both keep bits 8..15 of the 16-bit product; the gradient sum
//             wraps at 8 bits (:189); counter += (int8)gsum >> 2.
//   16-bit    counter int32, coef = low16(counter >> 15) << 12 as int16; prediction = mulhi(prev_delta, coef) << 2
//             wrapped to 16 bits; gradient sum wraps at 16 bits; counter += (int16)gsum >> 2.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SPZ_HD __host__ __device__ __forceinline__
#else
#define SPZ_HD inline
#endif

namespace spz {

SPZ_HD int32_t xffp_nblocks(uint64_t len, uint32_t ndims, int elem_sz)
{
    const uint32_t vector_sz = 32u / (uint32_t)elem_sz;
    const uint64_t bse = 8ull * ndims;
    int64_t nblocks = (int64_t)((len / ndims) / 8u);
    const uint32_t overrun = vector_sz - (ndims % vector_sz);
    const uint64_t trailing = len % bse;
    if (overrun > trailing) {
        nblocks -= (int64_t)((overrun + bse - 1) / bse);
        if (nblocks < 0) nblocks = 0;
    }
    return (int32_t)nblocks;
}

template <int W> SPZ_HD int32_t xffp_wrap(int32_t v) { return (int32_t)((uint32_t)v << (32 - W)) >> (32 - W); }

template <int W> SPZ_HD int32_t xffp_predict(uint32_t d, int32_t pd, int32_t counter)
{
    if (W == 8) {
        const int32_t coef = xffp_wrap<16>((int32_t)((uint32_t)(xffp_wrap<16>(counter) >> 5) << 4));
        const uint32_t a = (d & 1u) ? (uint32_t)pd : ((uint32_t)pd & 0xffu);       // pd is sign-extended: odd dims signed
        const uint32_t prod = (a * (uint32_t)coef) & 0xffffu;
        return xffp_wrap<8>((int32_t)(prod >> 8));
    }
    const int32_t coef = xffp_wrap<16>((int32_t)(((uint32_t)(counter >> 15) << 12) & 0xffffu));
    const int32_t hi = (pd * coef) >> 16;                                          // mulhi_epi16
    return xffp_wrap<16>((int32_t)((uint32_t)hi << 2));
}

template <int W> SPZ_HD int32_t xffp_sign(int32_t pd, int32_t err)                 // psignb / psignw(prev_delta, err)
{
    return xffp_wrap<W>(err == 0 ? 0 : (err < 0 ? -pd : pd));
}

template <int W> SPZ_HD uint32_t xffp_load(const void* p, uint64_t i)
{
    return W == 8 ? (uint32_t)((const uint8_t*)p)[i] : (uint32_t)((const uint16_t*)p)[i];
}
template <int W> SPZ_HD void xffp_store(void* p, uint64_t i, uint32_t v)
{
    if (W == 8) ((uint8_t*)p)[i] = (uint8_t)v; else ((uint16_t*)p)[i] = (uint16_t)v;
}

// One dim of the whole transform: the FIRE block part, then the delta-coded remainder.  ENC: x -> errors; else back.
template <int W, bool ENC>
SPZ_HD void xffp_dim(const void* src, void* dst, uint64_t len, uint32_t D, uint32_t d, int32_t nblocks)
{
    constexpr uint32_t WM = (W == 8) ? 0xffu : 0xffffu;
    uint32_t pv = 0;
    int32_t pd = 0, counter = 0;
    for (int32_t b = 0; b < nblocks; b++) {
        int32_t gsum = 0;
        for (uint32_t i = 0; i < 8; i++) {
            const uint64_t at = ((uint64_t)b * 8 + i) * D + d;
            const uint32_t in = xffp_load<W>(src, at);
            const int32_t pred = xffp_predict<W>(d, pd, counter);
            int32_t delta, err;
            if (ENC) {
                delta = xffp_wrap<W>((int32_t)((in - pv) & WM));
                err = xffp_wrap<W>((int32_t)(((uint32_t)delta - (uint32_t)pred) & WM));
                xffp_store<W>(dst, at, (uint32_t)err & WM);
                pv = in;
            } else {
                err = xffp_wrap<W>((int32_t)in);
                delta = xffp_wrap<W>((int32_t)(((uint32_t)err + (uint32_t)pred) & WM));
                pv = (pv + (uint32_t)delta) & WM;
                xffp_store<W>(dst, at, pv);
            }
            if (i & 1u) gsum = xffp_wrap<W>((int32_t)(((uint32_t)gsum + (uint32_t)xffp_sign<W>(pd, err)) & WM));
            pd = delta;
        }
        if (W == 8) counter = xffp_wrap<16>(counter + (gsum >> 2));
        else counter = (int32_t)((uint32_t)counter + (uint32_t)(gsum >> 2));
    }
    // remainder: delta with lag D (predict.cpp:268-279); with no block at all the first D elements are copied
    uint64_t i = (uint64_t)nblocks * 8 * D + d;
    if (nblocks == 0) {
        if (d < len) xffp_store<W>(dst, d, xffp_load<W>(src, d));
        i = (uint64_t)D + d;
    }
    for (; i < len; i += D) {
        if (ENC) xffp_store<W>(dst, i, (xffp_load<W>(src, i) - xffp_load<W>(src, i - D)) & WM);
        else     xffp_store<W>(dst, i, (xffp_load<W>(src, i) + xffp_load<W>(dst, i - D)) & WM);
    }
}

}  // namespace spz

// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (not product code).
//
// extern "C" shim around the UNMODIFIED reference implementation
// (dblalock/sprintz, cpp/Compress/sprintz.h:16-32).  It is compiled together
// with the reference's own .cpp files *where they lie* under /root/reference
// (see oracle/Makefile); no reference source is copied into this repo.  The
// resulting oracle/_ref/libsprintz_ref.so is used by tests/ to pin the
// restated oracle (oracle/sprintz_oracle.c) and the CUDA path byte-for-byte,
// and by bench.py as the "reference" CPU baseline arm.
//
// The shim adds only plumbing: C linkage, a codec-id switch, and a
// std::thread pool that calls the reference on independent chunk slices (the
// way a caller of the single-threaded reference would parallelise it).

#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>

#include "sprintz.h"   // from -I$(REF)/cpp/Compress
#include "sprintz_delta.h"   // the non-RLE sibling codecs
#include "sprintz_xff.h"     // compress8b_rowmajor_xff
#include "delta.h"           // stand-alone forecaster transforms
#include "predict.h"         // stand-alone FIRE transform

extern "C" {

// codec ids shared with include/sprintz_b200.h: bit0 = predictor (0 delta,
// 1 xff/FIRE), elem size passed separately.
int64_t ref_compress_rowmajor(int kind, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                              int write_size);
int64_t ref_decompress_rowmajor(int kind, int elem_sz, const void* src, void* dest);

int64_t ref_compress(int predictor, int elem_sz, const void* src, uint32_t len,
                     void* dest, uint16_t ndims, int write_size)
{
    if (predictor >= 2) return ref_compress_rowmajor(predictor - 1, elem_sz, src, len, dest, ndims, write_size);
    if (elem_sz == 1) {
        return predictor
            ? sprintz_compress_xff_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0)
            : sprintz_compress_delta_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0);
    }
    return predictor
        ? sprintz_compress_xff_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0)
        : sprintz_compress_delta_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0);
}

int64_t ref_decompress(int predictor, int elem_sz, const void* src, void* dest)
{
    if (predictor >= 2) return ref_decompress_rowmajor(predictor - 1, elem_sz, src, dest);
    if (elem_sz == 1) {
        return predictor
            ? sprintz_decompress_xff_8b((const int8_t*)src, (uint8_t*)dest)
            : sprintz_decompress_delta_8b((const int8_t*)src, (uint8_t*)dest);
    }
    return predictor
        ? sprintz_decompress_xff_16b((const int16_t*)src, (uint16_t*)dest)
        : sprintz_decompress_delta_16b((const int16_t*)src, (uint16_t*)dest);
}

// non-RLE sibling codecs: kind 1 = compress_rowmajor_delta_*, kind 2 = compress_rowmajor_* (sprintz_delta.h:27-44)
int64_t ref_compress_rowmajor(int kind, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                              int write_size)
{
    if (kind == 3) return elem_sz == 1 ? compress8b_rowmajor_xff((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0) : -1;
    if (elem_sz == 1) {
        return kind == 1
            ? compress_rowmajor_delta_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0)
            : compress_rowmajor_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0);
    }
    return kind == 1
        ? compress_rowmajor_delta_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0)
        : compress_rowmajor_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0);
}

int64_t ref_decompress_rowmajor(int kind, int elem_sz, const void* src, void* dest)
{
    if (kind == 3) return elem_sz == 1 ? decompress8b_rowmajor_xff((const int8_t*)src, (uint8_t*)dest) : -1;
    if (elem_sz == 1) {
        return kind == 1 ? decompress_rowmajor_delta_8b((const int8_t*)src, (uint8_t*)dest)
                         : decompress_rowmajor_8b((const int8_t*)src, (uint8_t*)dest);
    }
    return kind == 1 ? decompress_rowmajor_delta_16b((const int16_t*)src, (uint16_t*)dest)
                     : decompress_rowmajor_16b((const int16_t*)src, (uint16_t*)dest);
}

// stand-alone forecaster transforms (delta.h:17-67): order 1 = delta, 2 = double delta
int64_t ref_encode_forecast(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims, int write_size)
{
    if (elem_sz == 1)
        return order == 1 ? encode_delta_rowmajor_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0)
                          : encode_doubledelta_rowmajor_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0);
    return order == 1 ? encode_delta_rowmajor_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0)
                      : encode_doubledelta_rowmajor_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0);
}
int64_t ref_decode_forecast(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims)
{
    if (elem_sz == 1)
        return order == 1 ? decode_delta_rowmajor_8b((const int8_t*)src, len, (uint8_t*)dest, ndims)
                          : decode_doubledelta_rowmajor_8b((const int8_t*)src, len, (uint8_t*)dest, ndims);
    return order == 1 ? decode_delta_rowmajor_16b((const int16_t*)src, len, (uint16_t*)dest, ndims)
                      : decode_doubledelta_rowmajor_16b((const int16_t*)src, len, (uint16_t*)dest, ndims);
}

// stand-alone FIRE transform (predict.h:15-30)
int64_t ref_encode_xff_forecast(int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims, int write_size)
{
    return elem_sz == 1 ? encode_xff_rowmajor_8b((const uint8_t*)src, len, (int8_t*)dest, ndims, write_size != 0)
                        : encode_xff_rowmajor_16b((const uint16_t*)src, len, (int16_t*)dest, ndims, write_size != 0);
}
int64_t ref_decode_xff_forecast(int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims)
{
    return elem_sz == 1 ? decode_xff_rowmajor_8b((const int8_t*)src, len, (uint8_t*)dest, ndims)
                        : decode_xff_rowmajor_16b((const int16_t*)src, len, (uint16_t*)dest, ndims);
}

// Compress `nchunks` consecutive slices of `src` (chunk c covers elements
// [c*chunk_len, min(len,(c+1)*chunk_len))) with `nthreads` threads.  Chunk c is
// written at dest + c*dest_stride bytes; out_elems[c] receives the reference's
// return value.  Returns wall seconds of the parallel region.
double ref_compress_chunks(int predictor, int elem_sz, const void* src, uint64_t len,
                           uint64_t chunk_len, uint16_t ndims, void* dest,
                           uint64_t dest_stride, int64_t* out_elems, int nthreads)
{
    uint64_t nchunks = (len + chunk_len - 1) / chunk_len;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            uint64_t beg = c * chunk_len;
            uint64_t n = (beg + chunk_len <= len) ? chunk_len : (len - beg);
            out_elems[c] = ref_compress(predictor, elem_sz,
                (const char*)src + beg * elem_sz, (uint32_t)n,
                (char*)dest + c * dest_stride, ndims, 1);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Decompress chunk streams laid out at src + c*src_stride into
// dest + c*chunk_len elements.  NOTE the reference decoder stores whole 32-byte
// vectors, so chunks must be decoded in increasing order per thread *or* into
// private slack; here each thread decodes into a private bounce buffer and
// memcpy's the exact chunk out, so concurrent neighbours cannot clobber each other.
double ref_decompress_chunks(int predictor, int elem_sz, const void* src,
                             uint64_t src_stride, uint64_t nchunks, uint64_t len,
                             uint64_t chunk_len, void* dest, int64_t* out_elems,
                             int nthreads)
{
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        std::vector<char> bounce((chunk_len + 4096) * elem_sz + 256);
        for (;;) {
            uint64_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            uint64_t beg = c * chunk_len;
            uint64_t n = (beg + chunk_len <= len) ? chunk_len : (len - beg);
            out_elems[c] = ref_decompress(predictor, elem_sz,
                (const char*)src + c * src_stride, bounce.data());
            memcpy((char*)dest + beg * elem_sz, bounce.data(), n * elem_sz);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

/* oracle/sprintz_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C scalar restatement of the reference Sprintz codec path
 * (dblalock/sprintz cpp/Compress, sprintz.h:16-32).  See sprintz_oracle.c for the
 * per-function reference citations and for how the restatement is pinned.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use this library; the product (sprintz_b200/) never links or loads it.
 */
#ifndef SPRINTZ_ORACLE_H
#define SPRINTZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* predictor: 0 = delta, 1 = xff (FIRE).  elem_sz: 1 or 2 bytes.
 * `len` and the return value are ELEMENT counts, exactly like the reference
 * (for elem_sz == 2 the return is floor(bytes/2), sprintz_delta_lowdim.cpp:380-383).
 * *true_nbytes (optional) receives the exact byte length of the stream.
 * Returns -1 for ndims == 0 (sprintz.cpp:69-76). */
int64_t oracle_compress(int predictor, int elem_sz, const void* src, uint32_t len,
                        void* dest, uint16_t ndims, int write_size,
                        uint64_t* true_nbytes);

/* Decodes a stream that starts with the 8-byte RLE metadata (format.h:35-62).
 * Returns the number of elements written. Never writes past that count. */
int64_t oracle_decompress(int predictor, int elem_sz, const void* src, void* dest);

/* 5-argument form for streams written with write_size == 0
 * (sprintz_delta.h:52-90, sprintz_xff.h:47-83). */
int64_t oracle_decompress_raw(int predictor, int elem_sz, const void* src, void* dest,
                              uint16_t ndims, uint32_t ngroups, uint16_t remaining_len);

/* Block-size axis (BASELINE config 5): the same codecs with block_rows = 8 | 16 | 32 | 64 rows per block.  8 is the
 * reference (the functions above call these with 8); the other sizes apply the reference's own derived constants to a
 * larger block_sz and have NO reference build to check against: parity unpinned beyond block 8. */
int64_t oracle_compress_bs(int predictor, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                           int write_size, uint64_t* true_nbytes, int block_rows);
int64_t oracle_decompress_bs(int predictor, int elem_sz, const void* src, void* dest, int block_rows);
int64_t oracle_decompress_raw_bs(int predictor, int elem_sz, const void* src, void* dest, uint16_t ndims,
                                 uint32_t ngroups, uint16_t remaining_len, int block_rows);

/* Non-RLE sibling codecs: kind 1 = compress_rowmajor_delta_{8b,16b}, kind 2 =
 * compress_rowmajor_{8b,16b} (sprintz_delta.h:27-44). 6-byte metadata [u32 len][u16 ndims]. */
int64_t oracle_compress_rowmajor(int kind, int elem_sz, const void* src, uint32_t len, void* dest,
                                 uint16_t ndims, int write_size, uint64_t* true_nbytes);
int64_t oracle_decompress_rowmajor(int kind, int elem_sz, const void* src, void* dest);

/* Stand-alone forecaster transforms (delta.h:17-67): order 1 = delta, 2 = double delta.
 * encode returns len + metadata elements; decode takes the transformed elements WITHOUT metadata. */
int64_t oracle_encode_forecast(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                               int write_size);
int64_t oracle_decode_forecast(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims);

/* Stand-alone FIRE transform (predict.h:15-30): encode returns len + metadata elements; decode takes the transformed
 * elements WITHOUT metadata.  Restatement only -- there is no product kernel for it yet. */
int64_t oracle_encode_xff_forecast(int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims, int write_size);
int64_t oracle_decode_xff_forecast(int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims);

/* Query-on-compressed: decode, then per-dim unsigned max (op 1) / 64-bit sum (op 2) into
 * result[ndims]; returns the decoded element count. */
int64_t oracle_query(int predictor, int elem_sz, const void* src, void* dest, int op, uint64_t* result);

/* Entropy stage, oracle/huff_oracle.c: OUR byte-wise canonical Huffman container -- PARITY UNPINNED (Huff0
 * is not in the reference tree).  encode returns the container bytes; decode returns n or -1 for a malformed container. */
void oracle_huff_lengths(const uint64_t* count256, uint8_t* len256);
uint64_t oracle_huff_max_bytes(uint64_t n);
int64_t oracle_huff_encode(const uint8_t* src, uint64_t n, uint8_t* dst);
int64_t oracle_huff_decode(const uint8_t* cont, uint64_t cont_bytes, uint8_t* dst, uint64_t n);

/* Worst-case stream size in bytes for `len` elements (our own bound, used by tests
 * to size buffers; the reference has no capacity arguments). */
uint64_t oracle_max_compressed_bytes(int elem_sz, uint64_t len, uint16_t ndims);

#ifdef __cplusplus
}
#endif
#endif

/* sprintz_b200.h -- C ABI of the B200-native Sprintz encode/decode path.
 *
 * This library replaces ONE path of dblalock/sprintz: the codec entry points of
 * cpp/Compress/sprintz.h:16-32 (sprintz_{compress,decompress}_{delta,xff}_{8b,16b}) and the kernels
 * behind them (sprintz_delta_rle.cpp, sprintz_delta_lowdim.cpp, sprintz_xff_rle.cpp,
 * sprintz_xff_lowdim.cpp).  Every stream it writes is byte-identical to the reference's output for the
 * same input, and every stream the reference writes decodes to the same integers.
 *
 * Three layers, all plain C (pointers + sizes; no torch / C++ types):
 *   1. reference drop-ins on HOST buffers        spz_{compress,decompress}_{delta,xff}_{8b,16b}
 *   2. single-stream calls on DEVICE buffers     spz_compress_dev / spz_decompress_dev
 *   3. the chunked FRAME (the throughput path)   spz_frame_*  -- a long stream cut into chunks, each chunk
 *      an independent reference stream, plus a chunk directory from a device-wide exclusive scan.
 *
 * There is no CPU fallback: every call runs CUDA kernels on the current device and fails with
 * SPZ_E_CUDA if it cannot.  `len` and element-count returns are in ELEMENTS, as in the reference.
 */
#ifndef SPRINTZ_B200_H
#define SPRINTZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPZ_DELTA 0   /* sprintz_*_delta_*  */
#define SPZ_XFF   1   /* sprintz_*_xff_*  (FIRE) */
#define SPZ_RM_DELTA 2 /* frames only: compress_rowmajor_delta_* (non-RLE, sprintz_delta.h:36-44) */
#define SPZ_RM_PLAIN 3 /* frames only: compress_rowmajor_*       (non-RLE, sprintz_delta.h:27-34) */
#define SPZ_RM_XFF   4 /* frames only: compress8b_rowmajor_xff   (non-RLE FIRE, 8-bit only, sprintz_xff.h:28-31) */

#define SPZ_OK            0
#define SPZ_E_ARG        -1   /* the reference's own error return (ndims == 0, sprintz.cpp:71) */
#define SPZ_E_CUDA       -2   /* CUDA runtime / launch failure (see spz_last_error) */
#define SPZ_E_CAPACITY   -3   /* destination or workspace too small */
#define SPZ_E_STREAM     -4   /* malformed stream (ndims mismatch, overrun) */
#define SPZ_E_UNSUPPORTED -5  /* ndims > SPZ_MAX_NDIMS */

#define SPZ_MAX_NDIMS 1024
#define SPZ_FRAME_MAGIC 0x315a5053u   /* "SPZ1" */

typedef void* spz_stream_t;           /* a cudaStream_t (0 = default stream) */

/* per-chunk record (device memory, 16 bytes) */
typedef struct spz_chunk_meta {
    uint32_t nbytes;      /* encoder: exact stream length in bytes | decoder: bytes consumed */
    uint32_t ngroups;     /* format.h:35-45 */
    uint32_t remaining;   /* encoder: trailing raw elements | decoder (streams with metadata): elements produced */
    uint32_t status;      /* 0 ok; 1 ndims mismatch; 2 capacity overflow */
} spz_chunk_meta;

/* 64-byte frame header (little endian), followed by u64 chunk_off[nchunks+1], u32 chunk_nbytes[nchunks],
 * padding to 16, then the chunk streams (each starting 16-byte aligned at payload_off + chunk_off[c]). */
typedef struct spz_frame_header {
    uint32_t magic;
    uint16_t version;
    uint8_t  elem_sz;
    uint8_t  predictor;
    uint32_t ndims;
    uint32_t chunk_rows;
    uint64_t len;            /* total elements */
    uint64_t nchunks;
    uint64_t payload_off;    /* byte offset of the first chunk stream inside the frame */
    uint64_t payload_bytes;  /* sum of 16-byte-rounded chunk stream sizes */
    uint64_t frame_bytes;    /* payload_off + payload_bytes + 16 bytes of read slack */
    uint64_t reserved;
} spz_frame_header;

const char* spz_version(void);
const char* spz_last_error(void);    /* thread-local description of the calling thread's last failure; reading clears it */

/* ---------------------------------------------------------------- 1. reference drop-ins (host buffers)
 * Same arguments, return values and stream bytes as sprintz.h:16-32 (the C++ default `write_size=true`
 * becomes an explicit int).  Each call copies the input to the GPU, runs the single-stream kernel (one
 * team of lanes: FIRE and the stream parse are serial by construction) and copies the
 * result back.  Like the reference there are no capacity arguments: `dest` must hold
 * spz_max_compressed_bytes() bytes (compress) / the decoded length (decompress; no overrun past it,
 * unlike the reference's 32-byte vector stores). */
int64_t spz_compress_delta_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
int64_t spz_decompress_delta_8b(const int8_t*  src, uint8_t*  dest);
int64_t spz_compress_xff_8b   (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
int64_t spz_decompress_xff_8b (const int8_t*  src, uint8_t*  dest);
int64_t spz_compress_delta_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
int64_t spz_decompress_delta_16b(const int16_t* src, uint16_t* dest);
int64_t spz_compress_xff_16b  (const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
int64_t spz_decompress_xff_16b(const int16_t* src, uint16_t* dest);

/* generic forms of the above; *true_nbytes (optional) = exact stream length (the 16-bit return value is
 * floor(bytes/2), sprintz_delta_lowdim.cpp:380-383) */
int64_t spz_compress_host(int predictor, int elem_sz, const void* src, uint32_t len, void* dest,
                          uint16_t ndims, int write_size, uint64_t* true_nbytes);
int64_t spz_decompress_host(int predictor, int elem_sz, const void* src, void* dest);
/* 5-argument decoders for streams written with write_size == 0 (sprintz_delta.h:52-90, sprintz_xff.h:47-83) */
int64_t spz_decompress_raw_host(int predictor, int elem_sz, const void* src, void* dest,
                                uint16_t ndims, uint32_t ngroups, uint16_t remaining_len);

/* worst-case stream size in bytes for `len` elements (multiple of 16) */
uint64_t spz_max_compressed_bytes(int elem_sz, uint64_t len, uint16_t ndims);
/* walks a host stream's group headers: exact stream length in bytes and decoded element count */
int spz_measure_stream(int elem_sz, const void* src, int has_meta, uint16_t ndims, uint32_t ngroups,
                       uint16_t remaining_len, uint64_t* stream_bytes, uint64_t* decoded_elems);

/* Non-RLE sibling codecs of the same file family (sprintz_delta.h:27-44): delta + bit packing
 * and plain bit packing with the 6-byte metadata [u32 len][u16 ndims] (format.h:65-85), fixed ngroups = len/(16*ndims),
 * no run slots.  Same arguments / returns / bytes as compress_rowmajor_delta_{8b,16b}, compress_rowmajor_{8b,16b}
 * and their decoders.  kind: 1 = delta, 2 = plain, 3 = FIRE (8-bit only, see below). */
int64_t spz_compress_rowmajor_delta_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
int64_t spz_decompress_rowmajor_delta_8b(const int8_t*  src, uint8_t*  dest);
int64_t spz_compress_rowmajor_delta_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
int64_t spz_decompress_rowmajor_delta_16b(const int16_t* src, uint16_t* dest);
int64_t spz_compress_rowmajor_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
int64_t spz_decompress_rowmajor_8b(const int8_t*  src, uint8_t*  dest);
int64_t spz_compress_rowmajor_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
int64_t spz_decompress_rowmajor_16b(const int16_t* src, uint16_t* dest);
/* compress8b_rowmajor_xff / decompress8b_rowmajor_xff (sprintz_xff.h:28-31): FIRE without run slots, 8-bit only,
 * 8-byte metadata [u48 len][u16 ndims]; kind 3 of the generic forms below. */
int64_t spz_compress8b_rowmajor_xff(const uint8_t* src, uint64_t len, int8_t* dest, uint16_t ndims, int write_size);
int64_t spz_decompress8b_rowmajor_xff(const int8_t* src, uint8_t* dest);
int64_t spz_compress_rowmajor_host(int kind, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                                   int write_size, uint64_t* true_nbytes);
int64_t spz_decompress_rowmajor_host(int kind, int elem_sz, const void* src, void* dest);

/* Stand-alone forecaster transforms (delta.h:17-67, delta.cpp:34-744): same-size in/out,
 * y[i] = x[i] - x[i-ndims] (delta) or its second difference (double delta), x[i<0] = 0, optional 6-byte metadata
 * [u32 len][u16 ndims] in front of the output.  Same arguments and returns (len + metadata elements / len) as the
 * reference functions of the same names; order: 1 = delta, 2 = double delta. */
uint32_t spz_encode_delta_rowmajor_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
uint32_t spz_decode_delta_rowmajor_8b (const int8_t*   src, uint32_t len, uint8_t*  dest, uint16_t ndims);
uint32_t spz_encode_delta_rowmajor_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
uint32_t spz_decode_delta_rowmajor_16b(const int16_t*  src, uint32_t len, uint16_t* dest, uint16_t ndims);
uint32_t spz_encode_doubledelta_rowmajor_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
uint32_t spz_decode_doubledelta_rowmajor_8b (const int8_t*   src, uint32_t len, uint8_t*  dest, uint16_t ndims);
uint32_t spz_encode_doubledelta_rowmajor_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
uint32_t spz_decode_doubledelta_rowmajor_16b(const int16_t*  src, uint32_t len, uint16_t* dest, uint16_t ndims);
int64_t spz_encode_forecast_host(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims, int write_size);
int64_t spz_decode_forecast_host(int order, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims);
/* Stand-alone FIRE transform, encode/decode_xff_rowmajor_{8b,16b} (predict.h:15-30, predict.cpp:56-527): same
 * arguments and returns.  The predictor state of a dim runs through the whole stream, so a single stream is decoded by
 * one thread per dim -- a correctness path (long streams belong in the frame codec). */
uint32_t spz_encode_xff_rowmajor_8b (const uint8_t*  src, uint32_t len, int8_t*  dest, uint16_t ndims, int write_size);
uint32_t spz_decode_xff_rowmajor_8b (const int8_t*   src, uint32_t len, uint8_t*  dest, uint16_t ndims);
uint32_t spz_encode_xff_rowmajor_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, int write_size);
uint32_t spz_decode_xff_rowmajor_16b(const int16_t*  src, uint32_t len, uint16_t* dest, uint16_t ndims);
int64_t spz_xff_forecast_host(int enc, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims, int write_size);
int spz_xff_forecast_dev(int enc, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest, spz_stream_t stream);
/* device buffers (16-byte aligned, readable up to the next 16-byte boundary past len elements); decode needs
 * spz_forecast_workspace_bytes() of scratch for its chained scan */
uint64_t spz_forecast_workspace_bytes(uint64_t len, uint16_t ndims);
int spz_forecast_encode_dev(int order, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest, spz_stream_t stream);
int spz_forecast_decode_dev(int order, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest,
                            void* d_work, uint64_t work_bytes, spz_stream_t stream);

/* Query-on-compressed (query.hpp:22-29 QueryParams{op, materialize}, sprintz_delta.h:92-98,
 * sprintz_xff.h:86-93).  The reference's functor state never reaches its caller (query.hpp:231); here `result`
 * receives ndims u64: for every dim d the unsigned maximum (SPZ_Q_MAX) or the sum (SPZ_Q_SUM) of x[i], i % ndims == d,
 * over all decoded elements.  materialize != 0 also writes the decoded elements to dest (exactly decompress);
 * materialize == 0 leaves dest untouched (may be NULL).  Returns the decoded element count like the reference. */
#define SPZ_Q_NOOP 0   /* QueryTypes::NOOP */
#define SPZ_Q_MAX  1   /* QueryTypes::REDUCE_MAX */
#define SPZ_Q_SUM  2   /* QueryTypes::REDUCE_SUM */
int64_t spz_query_rowmajor_delta_rle_8b (const int8_t*  src, uint8_t*  dest, int op, int materialize, uint64_t* result);
int64_t spz_query_rowmajor_delta_rle_16b(const int16_t* src, uint16_t* dest, int op, int materialize, uint64_t* result);
int64_t spz_query_rowmajor_xff_rle_8b   (const int8_t*  src, uint8_t*  dest, int op, int materialize, uint64_t* result);
int64_t spz_query_rowmajor_xff_rle_16b  (const int16_t* src, uint16_t* dest, int op, int materialize, uint64_t* result);
int64_t spz_query_host(int predictor, int elem_sz, const void* src, void* dest, int op, int materialize, uint64_t* result);
/* Frames on device buffers.  materialize == 0 on frames of many 8 / 16-dim chunks runs FUSED: the decoder lanes fold
 * rows into registers and the launch reads the compressed bytes only (spz_last_query_fused() == 1).  Every other
 * case decodes (into d_dest, or into d_work when materialize == 0: spz_frame_query_workspace_bytes) and reduces
 * the decoded rows with a second kernel.  d_result: ndims u64 (device). */
uint64_t spz_frame_query_workspace_bytes(int elem_sz, uint64_t len, int materialize);
int spz_frame_query_dev(int predictor, int elem_sz, int op, int materialize, const void* d_frame, uint64_t frame_bytes,
                        uint64_t len, uint16_t ndims, uint32_t chunk_rows, void* d_dest, uint64_t dest_cap_elems,
                        spz_chunk_meta* d_status, uint64_t* d_result, void* d_work, uint64_t work_bytes,
                        spz_stream_t stream);
int spz_last_query_fused(void);
/* per-dim max / sum of row-major device elements (d_src 16-byte aligned); d_result: ndims u64 */
int spz_reduce_rows_dev(int op, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims, uint64_t* d_result,
                        spz_stream_t stream);

/* ---------------------------------------------------------------- 2. single stream, device buffers
 * The same codecs as layer 1 (sprintz.cpp:92-251) without the host copies; no reference equivalent -- the reference is
 * host-only.  d_dest must hold spz_max_compressed_bytes(); d_meta (1 record) receives sizes. Asynchronous on `stream`. */
int spz_compress_dev(int predictor, int elem_sz, const void* d_src, uint32_t len, uint16_t ndims,
                     int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream);
/* has_meta == 0: d_meta->ngroups / ->remaining are inputs.  src_bytes = readable bytes at d_src; it must
 * cover the stream plus 8 bytes (the kernels read whole aligned words / 16-byte units, never past it). */
int spz_decompress_dev(int predictor, int elem_sz, const void* d_src, uint64_t src_bytes, int has_meta,
                       uint16_t ndims, void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_meta,
                       spz_stream_t stream);

/* ---------------------------------------------------------------- 3. chunked frame
 * No reference equivalent: one reference call is one serial stream with a uint32_t length (sprintz.cpp:92-109,
 * sprintz_delta_rle.cpp:54-72); the frame is how the path is made parallel -- every chunk IS one reference stream.
 * chunk c covers elements [c*chunk_rows*ndims, (c+1)*chunk_rows*ndims) and is encoded exactly as the
 * reference encodes that slice. chunk_rows should be a multiple of 16 (whole groups). */
uint64_t spz_frame_nchunks(uint64_t len, uint16_t ndims, uint32_t chunk_rows);
uint64_t spz_frame_max_bytes(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows);
uint64_t spz_frame_workspace_bytes(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows);

/* Device buffers, asynchronous on `stream`. *d_frame_bytes (device, optional) receives the frame size. */
int spz_frame_encode_dev(int predictor, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                         uint32_t chunk_rows, void* d_frame, uint64_t frame_cap,
                         void* d_work, uint64_t work_bytes, uint64_t* d_frame_bytes, spz_stream_t stream);
/* frame_bytes = readable bytes at d_frame (the frame size, or the buffer's capacity).
 * d_status (device, optional, nchunks records) receives per-chunk decode status. */
int spz_frame_decode_dev(int predictor, int elem_sz, const void* d_frame, uint64_t frame_bytes, uint64_t len,
                         uint16_t ndims, uint32_t chunk_rows, void* d_dest, uint64_t dest_cap_elems,
                         spz_chunk_meta* d_status, spz_stream_t stream);
/* synchronous copy of the 64-byte header from a device frame */
int spz_frame_read_header(const void* d_frame, spz_frame_header* out, spz_stream_t stream);

/* Host buffers: H2D -> kernels -> D2H inside the call (pinned host memory makes the copies asynchronous).
 * Returns frame bytes (compress) / elements (decompress), or a negative SPZ_E_* code. */
int64_t spz_frame_compress_host(int predictor, int elem_sz, const void* src, uint64_t len, uint16_t ndims,
                                uint32_t chunk_rows, void* frame, uint64_t frame_cap);
int64_t spz_frame_decompress_host(const void* frame, uint64_t frame_bytes, void* dest, uint64_t dest_cap_elems);

/* Optional timing of spz_frame_encode_dev's launches with CUDA events on the call's stream (up to 64 calls
 * between enable and collect). ms3 = summed milliseconds of {codec kernel, directory scan, compaction}; for a
 * pipelined encode: {all codec kernels with the overlapped scan + compaction, 0, the exposed remainder}.
 * Process-wide measurement switch: not meant to be toggled while other threads encode. */
int spz_profile_enable(int on);
int spz_profile_collect(float* ms3, int* ncalls);

/* spz_frame_encode_dev can cut a frame into SLABS of whole chunks and run slab k's directory scan + compaction on a
 * second stream while slab k+1 is being encoded.  The frame bytes do not depend on the slab count.  slabs: 0 or 1 =
 * one slab (codec kernel, scan, compaction back to back: the default), n > 1 = n slabs, every slab but the last
 * compacted by one-warp CTAs that fit beside the next slab's codec kernel.  Measured: three slabs
 * pay for the wide-row encoder (int8 x 80 delta: 3.08 -> 2.84 ms) and not for the lane encoders, hence the default.  The
 * same slab machinery carries the host-buffer path (PCIe overlap) and the multi-GPU gathers (NVLink overlap), where it
 * does pay.  Returns the previous setting;
 * spz_last_encode_slabs() = slabs of this thread's last frame encode. */
int spz_set_encode_pipeline(int slabs);
int spz_last_encode_slabs(void);

/* number of CUDA kernels this library has launched in this process so far */
uint64_t spz_kernel_launches(void);

/* Kernel family used by the codec launches: 0 = automatic (lane-per-octet kernels for frames with many chunks of
 * 8 / 16 dims, team-per-chunk kernels otherwise), 1 = team kernels only, 2 = lane kernels whenever the shape
 * allows (testing: exercises them on small inputs). All families are CUDA and produce identical bytes.
 * Returns the previous setting. */
int spz_set_kernel_path(int path);
/* Kernel family used by the most recent encode (decode = 0) or decode (decode = 1) launch of this process:
 * 0 = generic team kernel, 1 = fast team kernel (spz_*_hd_kernel), 2 = lane kernel (spz_*_lane_kernel),
 * 3 = wide-row octet kernel (spz_*_wide_kernel).  Per calling thread. */
int spz_last_kernel_family(int decode);

/* ---------------------------------------------------------------- 3b. block-size axis (BASELINE config 5)
 * The reference's block is a compile-time constant of 8 rows (sprintz_delta_rle.cpp:62-64, sprintz_xff_rle.cpp:78-82); its
 * benchmark sweep names 16 / 32 / 64-row blocks.  These entry points run the RLE codecs (SPZ_DELTA / SPZ_XFF) with
 * block_rows = 8 | 16 | 32 | 64: the same stream format with a block of block_rows rows, the reference's own derived
 * constants (min_data_size = 8 * block_sz * 2, FIRE counter += grad_sum >> (log2_block_sz - 1)) and -- for more than 8
 * rows -- the high-dim (row-major payload) layout at every ndims.  block_rows == 8 is exactly the calls above.  No
 * reference build exists for the larger blocks: PARITY UNPINNED BEYOND BLOCK 8 (the oracle's generalisation is pinned to
 * the reference at 8 and checked by round trips otherwise).  chunk_rows must be a multiple of 2 * block_rows.  A frame
 * records block_rows in spz_frame_header.reserved (0 = 8). */
int spz_compress_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint32_t len, uint16_t ndims,
                        int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream);
int spz_decompress_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint64_t src_bytes, int has_meta,
                          uint16_t ndims, void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_meta, spz_stream_t stream);
int spz_frame_encode_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint64_t len, uint16_t ndims,
                            uint32_t chunk_rows, void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                            uint64_t* d_frame_bytes, spz_stream_t stream);
int spz_frame_decode_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_frame, uint64_t frame_bytes, uint64_t len,
                            uint16_t ndims, uint32_t chunk_rows, void* d_dest, uint64_t dest_cap_elems,
                            spz_chunk_meta* d_status, spz_stream_t stream);

/* ---------------------------------------------------------------- 4. multi-GPU: one process per GPU
 * No reference equivalent (the reference is single-threaded host code).  The path shards at chunk granularity with no
 * data-path collective; the one exchange step is gathering the ranks' compressed frames into ONE decodable frame on
 * a destination rank, over NCCL (ncclSend / ncclRecv grouped, NVLink / NVSwitch).
 *
 * Partition of a global stream of total_len elements over nranks (spz_shard_*): the stream is cut into STRIPES of
 * stripe_chunks chunks dealt round-robin -- rank r owns global stripes r, r + nranks, ... and holds them back to back in
 * its local buffer; stripe_chunks == 0 means one contiguous, balanced range of whole chunks per rank.  Only the last
 * global stripe can be short. */
typedef void* spz_comm_t;
#define SPZ_COMM_ID_BYTES 256
int spz_comm_unique_id(void* id);                                   /* rank 0; broadcast the bytes to every rank (host harness) */
int spz_comm_init(spz_comm_t* comm, int nranks, int rank, const void* id);      /* on the current CUDA device */
int spz_comm_destroy(spz_comm_t comm);
/* elements of the global stream that rank `rank` owns */
uint64_t spz_shard_len(uint64_t total_len, uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks, int nranks, int rank);
/* gathers the ranks' local frames (each encoded from that rank's shard with identical codec parameters) into one frame
 * of the global stream on rank dst.  d_gathered: spz_frame_gather_bytes() on dst (ignored elsewhere).  Asynchronous on
 * `stream` after a short host-side size exchange. */
uint64_t spz_frame_gather_max_bytes(int elem_sz, uint64_t total_len, uint16_t ndims, uint32_t chunk_rows);
int spz_frame_gather_dev(spz_comm_t comm, int dst, uint64_t total_len, uint32_t stripe_chunks, const void* d_frame,
                         void* d_gathered, uint64_t gathered_cap, spz_stream_t stream);
/* encode + gather, streamed stripe by stripe: stripe k travels while stripe k+1 is being encoded (stripe_chunks > 0).
 * The local frame (d_frame) is complete and valid as after spz_frame_encode_dev. */
int spz_frame_encode_gather_dev(spz_comm_t comm, int dst, int predictor, int elem_sz, const void* d_src, uint64_t total_len,
                                uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks, void* d_frame, uint64_t frame_cap,
                                void* d_work, uint64_t work_bytes, void* d_gathered, uint64_t gathered_cap, spz_stream_t stream);

/* Peer-memory form of the same exchange (NVLink stores, no NCCL in the data path).  spz_comm_shared_alloc is collective:
 * every rank allocates a shared region (CUDA IPC, mapped by all peers) with room for a gathered frame of gathered_bytes
 * (0 on ranks that never are the destination) and gets the address of that area in *d_gathered.
 * spz_frame_encode_gather_p2p_dev encodes the rank's shard stripe by stripe; each stripe's compaction kernel writes the
 * chunk streams and their directory entries straight to their final place in the DESTINATION's shared area while the
 * next stripe is being encoded (sizes travel as tagged flag words in peer memory; no host synchronisation).  The
 * gathered frame lands in dst's shared area (the pointer spz_comm_shared_alloc returned there); d_frame only receives
 * the local directory.  Asynchronous on `stream`; a rank that fails leaves the gathered frame without magic. */
int spz_comm_shared_alloc(spz_comm_t comm, uint64_t gathered_bytes, void** d_gathered);
int spz_frame_encode_gather_p2p_dev(spz_comm_t comm, int dst, int predictor, int elem_sz, const void* d_src, uint64_t total_len,
                                    uint16_t ndims, uint32_t chunk_rows, uint32_t stripe_chunks, void* d_frame, uint64_t frame_cap,
                                    void* d_work, uint64_t work_bytes, spz_stream_t stream);

/* ---------------------------------------------------------------- 5. float -> integer quantiser
 * python/datasets/compress_bench.py:45-60 (_quantize, axis = 0): the step that turns a float dataset (rows x ndims,
 * row-major) into the integer stream the codecs eat.  Per dim d, exactly the reference's arithmetic:
 *     y = x - min_d            (in the input precision: float32 or float64)
 *     y = float32(y) / max(1, float32(max_d - min_d))          (float32 division)
 *     q = trunc(y * (2^w - 1))                                 (float32 product, w = 8 | 16)
 * src_f64 != 0: the input is float64.  d_work: spz_quantize_workspace_bytes(ndims) of device scratch (per-dim
 * min / max).  NaN inputs are not supported (numpy propagates them into an undefined integer cast). */
uint64_t spz_quantize_workspace_bytes(uint16_t ndims);
int spz_quantize_dev(int elem_sz, int src_f64, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest,
                     void* d_work, uint64_t work_bytes, spz_stream_t stream);
int64_t spz_quantize_host(int elem_sz, int src_f64, const void* src, uint64_t len, uint16_t ndims, void* dest);

/* ---------------------------------------------------------------- 6. entropy stage (BASELINE config 3)
 * The paper's "SprintzFIRE+Huf" entropy-codes the bit-packed bytes with Huff0, which is NOT in the reference tree
 * (cpp/Compress/entropy.cpp:46-58 is a stub; README.md:29,45 points to the un-vendored lzbench fork).  PARITY UNPINNED:
 * this is OUR byte-wise canonical Huffman stage -- order-0 Huffman over bytes, 11-bit length limit, one table per
 * container, independently decodable 4096-byte segments behind a directory (one lane per segment) -- specified byte for
 * byte by oracle/huff_oracle.c.  It runs over any device buffer (16-byte aligned), typically a frame's payload.
 * Container: [64 B header][256 x 4-bit code lengths][u64 seg_off[nsegs+1]][u32 seg_nbytes[nsegs]][pad 16][streams][16 B].
 * Asynchronous on `stream`.  *d_cont_bytes (device, optional) receives the container size; *d_status (device) receives 0
 * or a bit set: 1 bad table | 2 bad header | 4 bad directory | 8 corrupt / truncated segment. */
uint64_t spz_huff_max_bytes(uint64_t raw_bytes);
uint64_t spz_huff_workspace_bytes(uint64_t raw_bytes);
int spz_huff_encode_dev(const void* d_src, uint64_t raw_bytes, void* d_cont, uint64_t cont_cap, void* d_work,
                        uint64_t work_bytes, uint64_t* d_cont_bytes, spz_stream_t stream);
int spz_huff_decode_dev(const void* d_cont, uint64_t cont_bytes, uint64_t raw_bytes, void* d_dest, uint64_t dest_cap,
                        void* d_work, uint64_t work_bytes, unsigned int* d_status, spz_stream_t stream);

/* releases the cached device workspaces of the host-buffer entry points */
void spz_release_workspaces(void);

#ifdef __cplusplus
}
#endif
#endif /* SPRINTZ_B200_H */

// spz_host.h -- host-side declarations shared by the C-ABI translation units (spz_api.cu, spz_comm.cu).
// Frame / workspace layouts, the slab-pipelined frame encoder, error plumbing and the kernel launchers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "spz_common.cuh"
#include "../../include/sprintz_b200.h"

namespace spz {

// ------------------------------------------------------------------ kernel launchers (one per .cu file)
cudaError_t launch_encode(int elem_sz, int predictor, EncParams p, cudaStream_t stream);
cudaError_t launch_decode(int elem_sz, int predictor, DecParams p, cudaStream_t stream);
cudaError_t launch_scan_sizes(const ChunkMeta* meta, uint64_t nchunks, uint64_t* chunk_off, uint32_t* chunk_nbytes,
                              unsigned long long* tile_state, unsigned int* ticket,
                              const unsigned long long* base_in, unsigned long long* total_out, int is_last,
                              spz_frame_header* hdr, const spz_frame_header& hdr_in, uint64_t* d_frame_bytes,
                              unsigned int* d_err, cudaStream_t stream);
cudaError_t launch_gather(const uint8_t* slots, uint64_t slot_stride, uint64_t nchunks, const uint64_t* chunk_off,
                          const uint32_t* chunk_nbytes, uint8_t* payload, cudaStream_t stream);
cudaError_t launch_rebase_directory(uint64_t* chunk_off, uint64_t nchunks, const uint64_t* seg_lo, const int64_t* seg_delta,
                                    uint32_t nsegs, uint64_t total_payload, uint32_t* chunk_nbytes, spz_frame_header* hdr,
                                    const spz_frame_header& hdr_in, cudaStream_t stream);
cudaError_t launch_gather_thin(const uint8_t* slots, uint64_t slot_stride, uint64_t nchunks, const uint64_t* chunk_off,
                               const uint32_t* chunk_nbytes, uint8_t* payload, uint32_t ctas, cudaStream_t stream);
uint64_t scan_tiles(uint64_t nchunks);
// What launch_encode() would do for this call: kernel family (0 generic, 1 team, 2 lane, 3 wide) and how many chunks
// fill the machine once (resident CTAs x chunks per CTA); nothing is launched.
struct EncodeProbe { int family; uint64_t round_chunks; };
EncodeProbe probe_encode(int elem_sz, int predictor, EncParams p);
extern thread_local EncodeProbe* g_enc_probe;      // set by probe_encode(): the launchers describe the launch instead of making it
template <class K>
static inline cudaError_t probe_fill(EncodeProbe* pr, int family, K kernel, int threads, size_t smem, uint64_t chunks_per_cta)
{
    int nb = 0, dev = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    pr->family = family;
    pr->round_chunks = (uint64_t)nb * (uint64_t)sms * chunks_per_cta;
    return cudaSuccess;
}
void set_force_path(int v);
int get_force_path();
int last_encode_family();
int last_decode_family();
cudaError_t launch_query_fused(int elem_sz, int predictor, DecParams p, cudaStream_t stream);
cudaError_t launch_reduce_rows(int op, int elem_sz, const uint8_t* src, uint64_t len, uint32_t D,
                               unsigned long long* result, cudaStream_t stream);
cudaError_t launch_xff_forecast(int enc, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                cudaStream_t stream);
uint64_t forecast_workspace_bytes(uint64_t len, uint32_t D);
cudaError_t launch_forecast_encode(int order, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                   cudaStream_t stream);
cudaError_t launch_forecast_decode(int order, int elem_sz, const uint8_t* src, uint8_t* dst, uint64_t len, uint32_t D,
                                   unsigned long long* state, unsigned int* ticket, cudaStream_t stream);
cudaError_t launch_quantize(int elem_sz, int src_f64, const void* src, uint64_t len, uint32_t D, void* dst, void* work,
                            cudaStream_t stream);
uint64_t quantize_workspace_bytes(uint32_t D);
void count_launch();                 // every kernel launch of the library bumps spz_kernel_launches()

// ------------------------------------------------------------------ errors
int cuda_fail(cudaError_t e, const char* where);     // records spz_last_error(), returns SPZ_E_CUDA
int fail_msg(int code, const char* msg);             // records a message for a non-CUDA failure, returns code
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::spz::cuda_fail(e__, #call); } while (0)

static inline uint64_t align16(uint64_t x) { return (x + 15ull) & ~15ull; }
static inline uint64_t align32(uint64_t x) { return (x + 31ull) & ~31ull; }
static inline bool valid_codec(int predictor, int elem_sz)
{
    return (predictor == 0 || predictor == 1) && (elem_sz == 1 || elem_sz == 2);
}
// frames also carry the non-RLE sibling codecs: SPZ_RM_DELTA (2), SPZ_RM_PLAIN (3), SPZ_RM_XFF (4, 8-bit only)
static inline bool valid_frame_codec(int predictor, int elem_sz)
{
    return predictor >= 0 && predictor <= 4 && (elem_sz == 1 || elem_sz == 2) && !(predictor == 4 && elem_sz != 1);
}

// ------------------------------------------------------------------ layout of a frame and of the encode workspace
struct FrameLayout {
    uint64_t nchunks, chunk_len, off_dir, off_sizes, payload_off;
};
FrameLayout frame_layout(uint64_t len, uint16_t ndims, uint32_t chunk_rows);

constexpr uint64_t kMaxSlabs = 256;      // slabs of one pipelined frame encode (directory-scan states are per slab)

// workspace = [scan tile states (per slab) | tickets | running payload totals | error word | meta | slots]
struct WorkLayout {
    uint64_t off_state, off_ticket, off_tot, off_err, off_meta, off_slots, slot_stride, total, state_words;
};
WorkLayout work_layout(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows);

// ------------------------------------------------------------------ the slab-pipelined frame encoder
// A frame encode is cut into slabs of whole chunks.  Slab k's codec kernel runs on the caller's stream; its
// directory scan + compaction run on a side stream as soon as the kernel is done, i.e. WHILE slab k+1 is being
// encoded (the codec kernels are instruction-bound, the compaction is HBM-bound).  The scan carries the running
// payload offset from slab to slab, so the frame bytes do not depend on the slab size.
struct FrameJob {
    int predictor, elem_sz;
    uint16_t ndims;
    uint32_t chunk_rows;
    uint64_t len;
    FrameLayout f;
    WorkLayout w;
    uint64_t slab_chunks, nslabs, tiles_per_slab;
    uint32_t block_rows;                 // rows per block (8 = the reference's format)
    int family;                          // kernel family chosen for the WHOLE frame (slabs must not re-decide)
    const uint8_t* d_src;
    uint8_t* d_frame;
    uint8_t* work;
    uint64_t* d_frame_bytes;
    spz_frame_header h;
    // slab k covers chunks [c0, c0 + nc) = elements [e0, e0 + ne)
    void slab(uint64_t k, uint64_t& c0, uint64_t& nc, uint64_t& e0, uint64_t& ne) const
    {
        c0 = k * slab_chunks;
        nc = (c0 + slab_chunks <= f.nchunks) ? slab_chunks : (f.nchunks - c0);
        e0 = c0 * f.chunk_len;
        ne = (e0 + nc * f.chunk_len <= len) ? nc * f.chunk_len : (len - e0);
    }
    unsigned long long* d_tot() const { return (unsigned long long*)(work + w.off_tot); }   // [k] = payload offset before slab k
    unsigned int* d_err() const { return (unsigned int*)(work + w.off_err); }
};
// slab_chunks == 0: pick automatically (whole machine rounds of the codec kernel; one slab for small frames)
int frame_job_init(FrameJob& j, int predictor, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                   uint32_t chunk_rows, void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                   uint64_t* d_frame_bytes, uint64_t slab_chunks, uint32_t block_rows = 8);
int frame_job_reset(const FrameJob& j, cudaStream_t st);                       // zeroes scan states / tickets / totals / error word
int frame_job_encode_slab(const FrameJob& j, uint64_t k, cudaStream_t st);     // codec kernel of slab k
int frame_job_scan_slab(const FrameJob& j, uint64_t k, cudaStream_t st);       // directory scan of slab k only (slab order!)
int frame_job_compact_slab(const FrameJob& j, uint64_t k, cudaStream_t st, int thin_ctas = 0);   // directory scan + compaction of slab k (slab order!); thin_ctas > 0: one-warp CTAs that fit beside a running codec kernel

// side stream + events of the pipelined encode (one set per process, guarded by a mutex while a call enqueues)
struct SidePipe {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_enc[kMaxSlabs];
    cudaEvent_t ev_done = nullptr, ev_begin = nullptr;
    int nev = 0;
    int sms = 0;                         // SMs of the device the pipe was created on
    cudaError_t init(int nslabs);
};
SidePipe& side_pipe();

}  // namespace spz

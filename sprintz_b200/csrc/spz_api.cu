// spz_api.cu -- the C ABI (include/sprintz_b200.h) over the CUDA kernels.
//
// Host-side mirror of the reference's operator interface for this path: sprintz.h:16-32 (the eight
// compress/decompress entry points) and the ndims dispatch of sprintz.cpp:69-85 (which happens inside
// launch_encode / launch_decode).  No CPU codec lives here: the only host-side stream logic is
// spz_measure_stream(), which walks group headers to learn how many bytes to copy because the reference
// signatures carry no sizes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "spz_host.h"

using namespace spz;

static thread_local char g_err[256] = "";
static std::atomic<unsigned long long> g_launches{0};

namespace spz {
int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SPZ_E_CUDA;
}
int fail_msg(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace spz

// number of CUDA kernels this library has launched in this process (bench.py's gpu_launches)
extern "C" uint64_t spz_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------ optional per-kernel timing of the frame encoder
// bench.py needs the codec kernel's own duration (roofline), separate from scan + compaction. When enabled,
// spz_frame_encode_dev records CUDA events on its stream around each launch; nothing is synchronised
// until spz_profile_collect().
namespace {
constexpr int kProfSlots = 64;
struct Profile {
    bool on = false;
    int n = 0;
    cudaEvent_t ev[kProfSlots][4];
    bool made = false;
} g_prof;
}  // namespace

extern "C" int spz_profile_enable(int on)
{
    if (on && !g_prof.made) {
        for (int i = 0; i < kProfSlots; i++)
            for (int j = 0; j < 4; j++) CK(cudaEventCreate(&g_prof.ev[i][j]));
        g_prof.made = true;
    }
    g_prof.on = on != 0;
    g_prof.n = 0;
    return SPZ_OK;
}

// sums (ms) over the recorded frame-encode calls: [0] codec kernel, [1] directory scan, [2] compaction
extern "C" int spz_profile_collect(float* ms3, int* ncalls)
{
    ms3[0] = ms3[1] = ms3[2] = 0.f;
    const int n = g_prof.n < kProfSlots ? g_prof.n : kProfSlots;
    for (int i = 0; i < n; i++) {
        CK(cudaEventSynchronize(g_prof.ev[i][3]));
        for (int j = 0; j < 3; j++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, g_prof.ev[i][j], g_prof.ev[i][j + 1]));
            ms3[j] += ms;
        }
    }
    if (ncalls) *ncalls = n;
    g_prof.n = 0;
    return SPZ_OK;
}
#define PROF_MARK(j) do { if (g_prof.on && g_prof.n < kProfSlots) CK(cudaEventRecord(g_prof.ev[g_prof.n][j], stream)); } while (0)
#define PROF_NEXT() do { if (g_prof.on && g_prof.n < kProfSlots) g_prof.n++; } while (0)

extern "C" int spz_set_kernel_path(int path)
{
    const int old = get_force_path();
    if (path >= 0 && path <= 2) set_force_path(path);
    return old;
}

extern "C" int spz_last_kernel_family(int decode) { return decode ? last_decode_family() : last_encode_family(); }

extern "C" const char* spz_version(void) { return "sprintz_b200 0.1 (sm_100a)"; }
// description of this thread's most recent failure; reading it clears it
extern "C" const char* spz_last_error(void)
{
    static thread_local char out[256];
    memcpy(out, g_err, sizeof(out));
    g_err[0] = 0;
    return out;
}

extern "C" uint64_t spz_max_compressed_bytes(int elem_sz, uint64_t len, uint16_t ndims)
{
    const uint64_t D = ndims ? ndims : 1;
    const uint64_t hb = elem_sz == 1 ? 3 : 4;
    const uint64_t H = (2 * D * hb + 7) / 8;
    const uint64_t nblocks = len / (8 * D);
    return align16(8 + len * (uint64_t)elem_sz + (nblocks / 2 + 2) * H + 16) + 16;
}

// ------------------------------------------------------------------ layout of a frame

namespace spz {
FrameLayout frame_layout(uint64_t len, uint16_t ndims, uint32_t chunk_rows)
{
    FrameLayout f;
    f.chunk_len = (uint64_t)chunk_rows * ndims;
    f.nchunks = len == 0 ? 1 : (len + f.chunk_len - 1) / f.chunk_len;
    f.off_dir = sizeof(spz_frame_header);
    f.off_sizes = f.off_dir + 8 * (f.nchunks + 1);
    f.payload_off = align16(f.off_sizes + 4 * f.nchunks);
    return f;
}
}  // namespace spz

extern "C" uint64_t spz_frame_nchunks(uint64_t len, uint16_t ndims, uint32_t chunk_rows)
{
    if (ndims == 0 || chunk_rows == 0) return 0;
    return frame_layout(len, ndims, chunk_rows).nchunks;
}

static uint64_t slot_stride_of(int elem_sz, uint64_t chunk_len, uint16_t ndims)
{
    return align32(spz_max_compressed_bytes(elem_sz, chunk_len, ndims));      // whole sectors: 256-bit stores
}

extern "C" uint64_t spz_frame_max_bytes(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows)
{
    if (ndims == 0 || chunk_rows == 0) return 0;
    const FrameLayout f = frame_layout(len, ndims, chunk_rows);
    return f.payload_off + f.nchunks * slot_stride_of(elem_sz, f.chunk_len, ndims) + 16;
}

namespace spz {
// workspace = [scan tile states, per slab | tickets | running payload totals | error word | meta | slots]
WorkLayout work_layout(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows)
{
    const FrameLayout f = frame_layout(len, ndims, chunk_rows);
    WorkLayout w;
    w.off_state = 0;
    w.state_words = scan_tiles(f.nchunks) + kMaxSlabs;             // every slab rounds its tile count up
    w.off_ticket = align16(8 * w.state_words);
    w.off_tot = align16(w.off_ticket + 4 * kMaxSlabs);
    w.off_err = align16(w.off_tot + 8 * (kMaxSlabs + 1));
    w.off_meta = w.off_err + 16;
    w.off_slots = align32(w.off_meta + sizeof(ChunkMeta) * f.nchunks);
    w.slot_stride = slot_stride_of(elem_sz, f.chunk_len < len || len == 0 ? f.chunk_len : len, ndims);
    w.total = w.off_slots + f.nchunks * w.slot_stride + 16;
    return w;
}
}  // namespace spz

extern "C" uint64_t spz_frame_workspace_bytes(int elem_sz, uint64_t len, uint16_t ndims, uint32_t chunk_rows)
{
    if (ndims == 0 || chunk_rows == 0) return 0;
    return work_layout(elem_sz, len, ndims, chunk_rows).total;
}

// ------------------------------------------------------------------ device entry points

static int compress_dev_impl(uint32_t variant, int predictor, int elem_sz, const void* d_src, uint32_t len, uint16_t ndims,
                             int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream, uint32_t block_rows = 8);
static bool valid_block_rows(int r) { return r == 8 || r == 16 || r == 32 || r == 64; }

extern "C" int spz_compress_dev(int predictor, int elem_sz, const void* d_src, uint32_t len, uint16_t ndims,
                                int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream)
{
    return compress_dev_impl(0, predictor, elem_sz, d_src, len, ndims, write_size, d_dest, d_meta, stream);
}

extern "C" int spz_compress_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint32_t len, uint16_t ndims,
                                   int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream)
{
    if (!valid_block_rows(block_rows)) return SPZ_E_ARG;
    if (block_rows != 8 && 2ull * (uint64_t)block_rows * ndims + ndims > 0xffffull) return SPZ_E_UNSUPPORTED;     // u16 tail length
    return compress_dev_impl(0, predictor, elem_sz, d_src, len, ndims, write_size, d_dest, d_meta, stream, (uint32_t)block_rows);
}

static int compress_dev_impl(uint32_t variant, int predictor, int elem_sz, const void* d_src, uint32_t len, uint16_t ndims,
                             int write_size, void* d_dest, spz_chunk_meta* d_meta, spz_stream_t stream, uint32_t block_rows)
{
    if (!valid_codec(predictor, elem_sz) || ndims == 0) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    EncParams p;
    memset(&p, 0, sizeof(p));
    p.variant = variant; p.block_rows = block_rows;
    p.src = (const uint8_t*)d_src; p.dst = (uint8_t*)d_dest; p.meta = (ChunkMeta*)d_meta;
    p.len = len; p.chunk_len = len ? len : 1; p.nchunks = 1; p.slot_stride = 0;
    p.ndims = ndims; p.write_size = write_size ? 1 : 0;
    CK(launch_encode(elem_sz, predictor, p, (cudaStream_t)stream));
    return SPZ_OK;
}

static int decompress_dev_impl(uint32_t variant, int predictor, int elem_sz, const void* d_src, uint64_t src_bytes, int has_meta,
                               uint16_t ndims, void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_meta,
                               spz_stream_t stream, uint32_t block_rows = 8);

extern "C" int spz_decompress_dev(int predictor, int elem_sz, const void* d_src, uint64_t src_bytes, int has_meta,
                                  uint16_t ndims, void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_meta,
                                  spz_stream_t stream)
{
    return decompress_dev_impl(0, predictor, elem_sz, d_src, src_bytes, has_meta, ndims, d_dest, dest_cap_elems, d_meta, stream);
}

extern "C" int spz_decompress_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint64_t src_bytes,
                                     int has_meta, uint16_t ndims, void* d_dest, uint64_t dest_cap_elems,
                                     spz_chunk_meta* d_meta, spz_stream_t stream)
{
    if (!valid_block_rows(block_rows)) return SPZ_E_ARG;
    return decompress_dev_impl(0, predictor, elem_sz, d_src, src_bytes, has_meta, ndims, d_dest, dest_cap_elems, d_meta, stream,
                               (uint32_t)block_rows);
}

static int decompress_dev_impl(uint32_t variant, int predictor, int elem_sz, const void* d_src, uint64_t src_bytes, int has_meta,
                               uint16_t ndims, void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_meta,
                               spz_stream_t stream, uint32_t block_rows)
{
    if (!valid_codec(predictor, elem_sz) || ndims == 0) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    if (!has_meta && !d_meta) return SPZ_E_ARG;
    DecParams p;
    memset(&p, 0, sizeof(p));
    p.variant = variant; p.block_rows = block_rows;
    p.src = (const uint8_t*)d_src; p.offsets = nullptr; p.dst = (uint8_t*)d_dest; p.meta = (ChunkMeta*)d_meta;
    p.len = dest_cap_elems; p.chunk_len = dest_cap_elems ? dest_cap_elems : 1; p.nchunks = 1; p.slot_stride = 0;
    p.dst_cap = dest_cap_elems; p.ndims = ndims; p.has_meta = has_meta ? 1 : 0;
    p.src_end = (const uint8_t*)d_src + src_bytes;
    CK(launch_decode(elem_sz, predictor, p, (cudaStream_t)stream));
    return SPZ_OK;
}

// ------------------------------------------------------------------ the slab-pipelined frame encoder (spz_host.h)

namespace spz {

cudaError_t SidePipe::init(int nslabs)
{
    cudaError_t e;
    if (!side) {
        int lo = 0, hi = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&lo, &hi)) != cudaSuccess) return e;
        // highest priority: the few CTAs of the directory scan / compaction get SM slots as soon as encoder CTAs retire
        if ((e = cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming)) != cudaSuccess) return e;
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    for (; nev < nslabs && nev < (int)kMaxSlabs; nev++)
        if ((e = cudaEventCreateWithFlags(&ev_enc[nev], cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaSuccess;
}
SidePipe& side_pipe() { static SidePipe p; return p; }

}  // namespace spz

namespace {
std::mutex g_pipe_mu;                       // one call at a time enqueues on the shared side stream
std::atomic<int> g_pipeline_slabs{0};       // 0: automatic; 1: classic (scan + compaction after the encoder); n: n slabs
thread_local int g_last_encode_slabs = 1;
}  // namespace

extern "C" int spz_set_encode_pipeline(int slabs)
{
    const int old = g_pipeline_slabs.load();
    if (slabs >= 0 && slabs <= (int)kMaxSlabs) g_pipeline_slabs.store(slabs);
    return old;
}
extern "C" int spz_last_encode_slabs(void) { return g_last_encode_slabs; }

namespace spz {

int frame_job_init(FrameJob& j, int predictor, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                   uint32_t chunk_rows, void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                   uint64_t* d_frame_bytes, uint64_t slab_chunks, uint32_t block_rows)
{
    if (!valid_frame_codec(predictor, elem_sz) || ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    if (block_rows != 8 && (predictor > 1 || !valid_block_rows((int)block_rows))) return SPZ_E_ARG;
    if (block_rows != 8 && 2ull * block_rows * ndims + ndims > 0xffffull) return SPZ_E_UNSUPPORTED;   // the tail length is a u16 (format.h:35-45)
    j.block_rows = block_rows;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    j.predictor = predictor; j.elem_sz = elem_sz; j.ndims = ndims; j.chunk_rows = chunk_rows; j.len = len;
    j.f = frame_layout(len, ndims, chunk_rows);
    if (j.f.chunk_len > 0xffffffffull) return SPZ_E_ARG;           // a chunk is one reference stream: uint32_t len
    j.w = work_layout(elem_sz, len, ndims, chunk_rows);
    if (work_bytes < j.w.total) return SPZ_E_CAPACITY;
    if (frame_cap < spz_frame_max_bytes(elem_sz, len, ndims, chunk_rows)) return SPZ_E_CAPACITY;
    j.d_src = (const uint8_t*)d_src; j.d_frame = (uint8_t*)d_frame; j.work = (uint8_t*)d_work; j.d_frame_bytes = d_frame_bytes;
    memset(&j.h, 0, sizeof(j.h));
    j.h.magic = SPZ_FRAME_MAGIC; j.h.version = 1; j.h.elem_sz = (uint8_t)elem_sz; j.h.predictor = (uint8_t)predictor;
    j.h.ndims = ndims; j.h.chunk_rows = chunk_rows; j.h.len = len; j.h.nchunks = j.f.nchunks; j.h.payload_off = j.f.payload_off;
    j.h.reserved = block_rows == 8 ? 0 : block_rows;           // 0: the reference's 8-row blocks

    // kernel family + machine round, decided ONCE on the whole frame
    EncParams p;
    memset(&p, 0, sizeof(p));
    p.src = j.d_src; p.dst = j.work + j.w.off_slots; p.meta = (ChunkMeta*)(j.work + j.w.off_meta);
    p.len = len; p.chunk_len = j.f.chunk_len; p.nchunks = j.f.nchunks; p.slot_stride = j.w.slot_stride;
    p.ndims = ndims; p.write_size = 1; p.block_rows = block_rows;
    const EncodeProbe pr = probe_encode(elem_sz, predictor, p);
    if (pr.family < 0) return fail_msg(SPZ_E_CUDA, "probe_encode failed");
    j.family = pr.family;
    if (slab_chunks == 0) {
        const int want = g_pipeline_slabs.load();
        // Measured (profiles/r2_pipeline_sweep.jsonl): splitting the device-resident frame encode into slabs does NOT
        // pay -- the compaction CTAs displace encoder CTAs (the lane encoders run 4 CTAs per SM at 116 registers, nothing
        // else fits beside them) and every slab boundary drains the machine, so c3 goes 3.46 -> 3.66 ms with 7 slabs.
        // Automatic therefore means ONE slab; the slab machinery is what the host path (PCIe overlap) and the streamed
        // multi-GPU gather (NVLink overlap) are built on, and `want` > 1 forces it here for tests and tuning.  With the
        // thin (one-warp) compaction CTAs of this round three slabs DO pay for the wide-row encoder (c4: 3.08 -> 2.87 ms,
        // it runs 12 machine rounds of small CTAs) but still not for the lane encoders (c3: 3.50 -> 3.70 ms at best, also
        // with slabs cut at machine rounds): the bench reports the pipelined time beside the classic one.
        if (want > 1) slab_chunks = (j.f.nchunks + want - 1) / want;
        else slab_chunks = j.f.nchunks;
    }
    if (slab_chunks == 0) slab_chunks = 1;
    j.slab_chunks = slab_chunks;
    j.nslabs = (j.f.nchunks + slab_chunks - 1) / slab_chunks;
    if (j.nslabs > kMaxSlabs) { j.slab_chunks = (j.f.nchunks + kMaxSlabs - 1) / kMaxSlabs; j.nslabs = (j.f.nchunks + j.slab_chunks - 1) / j.slab_chunks; }
    j.tiles_per_slab = scan_tiles(j.slab_chunks);
    if (j.tiles_per_slab * j.nslabs > j.w.state_words) return SPZ_E_CAPACITY;     // cannot happen: state_words = tiles + kMaxSlabs
    return SPZ_OK;
}

int frame_job_reset(const FrameJob& j, cudaStream_t st)
{
    CK(cudaMemsetAsync(j.work, 0, j.w.off_meta, st));             // scan tile states, tickets, totals, error word
    return SPZ_OK;
}

int frame_job_encode_slab(const FrameJob& j, uint64_t k, cudaStream_t st)
{
    uint64_t c0, nc, e0, ne;
    j.slab(k, c0, nc, e0, ne);
    EncParams p;
    memset(&p, 0, sizeof(p));
    p.src = j.d_src + e0 * j.elem_sz; p.dst = j.work + j.w.off_slots + c0 * j.w.slot_stride;
    p.meta = (ChunkMeta*)(j.work + j.w.off_meta) + c0;
    p.len = ne; p.chunk_len = j.f.chunk_len; p.nchunks = nc; p.slot_stride = j.w.slot_stride;
    p.ndims = j.ndims; p.write_size = 1; p.block_rows = j.block_rows;
    p.family_pin = (uint32_t)j.family + 1u;
    CK(launch_encode(j.elem_sz, j.predictor, p, st));
    return SPZ_OK;
}

// directory scan of slab k alone: local chunk_off / chunk_nbytes and the running payload total; no header, no compaction
// (the peer-memory gather moves the streams itself, spz_comm.cu)
int frame_job_scan_slab(const FrameJob& j, uint64_t k, cudaStream_t st)
{
    uint64_t c0, nc, e0, ne;
    j.slab(k, c0, nc, e0, ne);
    ChunkMeta* meta = (ChunkMeta*)(j.work + j.w.off_meta);
    uint64_t* chunk_off = (uint64_t*)(j.d_frame + j.f.off_dir);
    uint32_t* chunk_nbytes = (uint32_t*)(j.d_frame + j.f.off_sizes);
    CK(launch_scan_sizes(meta + c0, nc, chunk_off + c0, chunk_nbytes + c0,
                         (unsigned long long*)(j.work + j.w.off_state) + k * j.tiles_per_slab,
                         (unsigned int*)(j.work + j.w.off_ticket) + k, j.d_tot() + k, j.d_tot() + k + 1,
                         0, (spz_frame_header*)j.d_frame, j.h, nullptr, j.d_err(), st));
    return SPZ_OK;
}

int frame_job_compact_slab(const FrameJob& j, uint64_t k, cudaStream_t st, int thin_ctas)
{
    uint64_t c0, nc, e0, ne;
    j.slab(k, c0, nc, e0, ne);
    ChunkMeta* meta = (ChunkMeta*)(j.work + j.w.off_meta);
    uint64_t* chunk_off = (uint64_t*)(j.d_frame + j.f.off_dir);
    uint32_t* chunk_nbytes = (uint32_t*)(j.d_frame + j.f.off_sizes);
    CK(launch_scan_sizes(meta + c0, nc, chunk_off + c0, chunk_nbytes + c0,
                         (unsigned long long*)(j.work + j.w.off_state) + k * j.tiles_per_slab,
                         (unsigned int*)(j.work + j.w.off_ticket) + k, j.d_tot() + k, j.d_tot() + k + 1,
                         k + 1 == j.nslabs, (spz_frame_header*)j.d_frame, j.h, k + 1 == j.nslabs ? j.d_frame_bytes : nullptr,
                         j.d_err(), st));
    if (thin_ctas > 0)
        CK(launch_gather_thin(j.work + j.w.off_slots + c0 * j.w.slot_stride, j.w.slot_stride, nc, chunk_off + c0, chunk_nbytes + c0,
                              j.d_frame + j.f.payload_off, (uint32_t)thin_ctas, st));
    else
        CK(launch_gather(j.work + j.w.off_slots + c0 * j.w.slot_stride, j.w.slot_stride, nc, chunk_off + c0, chunk_nbytes + c0,
                         j.d_frame + j.f.payload_off, st));
    return SPZ_OK;
}

}  // namespace spz

static int frame_encode_impl(int predictor, int elem_sz, uint32_t block_rows, const void* d_src, uint64_t len, uint16_t ndims,
                             uint32_t chunk_rows, void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                             uint64_t* d_frame_bytes, spz_stream_t stream_);

extern "C" int spz_frame_encode_dev(int predictor, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                                    uint32_t chunk_rows, void* d_frame, uint64_t frame_cap,
                                    void* d_work, uint64_t work_bytes, uint64_t* d_frame_bytes, spz_stream_t stream)
{
    return frame_encode_impl(predictor, elem_sz, 8, d_src, len, ndims, chunk_rows, d_frame, frame_cap, d_work, work_bytes,
                             d_frame_bytes, stream);
}

extern "C" int spz_frame_encode_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_src, uint64_t len, uint16_t ndims,
                                       uint32_t chunk_rows, void* d_frame, uint64_t frame_cap,
                                       void* d_work, uint64_t work_bytes, uint64_t* d_frame_bytes, spz_stream_t stream)
{
    if (!valid_block_rows(block_rows) || chunk_rows % (2u * (uint32_t)block_rows) != 0) return SPZ_E_ARG;     // whole groups per chunk
    return frame_encode_impl(predictor, elem_sz, (uint32_t)block_rows, d_src, len, ndims, chunk_rows, d_frame, frame_cap, d_work,
                             work_bytes, d_frame_bytes, stream);
}

static int frame_encode_impl(int predictor, int elem_sz, uint32_t block_rows, const void* d_src, uint64_t len, uint16_t ndims,
                             uint32_t chunk_rows, void* d_frame, uint64_t frame_cap, void* d_work, uint64_t work_bytes,
                             uint64_t* d_frame_bytes, spz_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    FrameJob j;
    const int rc = frame_job_init(j, predictor, elem_sz, d_src, len, ndims, chunk_rows, d_frame, frame_cap, d_work,
                                  work_bytes, d_frame_bytes, 0, block_rows);
    if (rc != SPZ_OK) return rc;
    g_last_encode_slabs = (int)j.nslabs;
    if (frame_job_reset(j, stream) != SPZ_OK) return SPZ_E_CUDA;

    if (j.nslabs == 1) {            // classic: codec kernel, directory scan, compaction back to back on the caller's stream
        PROF_MARK(0);
        if (int r = frame_job_encode_slab(j, 0, stream)) return r;
        PROF_MARK(1);
        uint64_t c0, nc, e0, ne;
        j.slab(0, c0, nc, e0, ne);
        ChunkMeta* meta = (ChunkMeta*)(j.work + j.w.off_meta);
        uint64_t* chunk_off = (uint64_t*)(j.d_frame + j.f.off_dir);
        uint32_t* chunk_nbytes = (uint32_t*)(j.d_frame + j.f.off_sizes);
        CK(launch_scan_sizes(meta, nc, chunk_off, chunk_nbytes, (unsigned long long*)(j.work + j.w.off_state),
                             (unsigned int*)(j.work + j.w.off_ticket), nullptr, j.d_tot() + 1, 1,
                             (spz_frame_header*)j.d_frame, j.h, d_frame_bytes, j.d_err(), stream));
        PROF_MARK(2);
        CK(launch_gather(j.work + j.w.off_slots, j.w.slot_stride, nc, chunk_off, chunk_nbytes, j.d_frame + j.f.payload_off, stream));
        PROF_MARK(3);
        PROF_NEXT();
        return SPZ_OK;
    }

    // pipelined: slab k's scan + compaction run on the side stream while slab k+1 is being encoded
    std::lock_guard<std::mutex> lk(g_pipe_mu);
    SidePipe& sp = side_pipe();
    CK(sp.init((int)j.nslabs));
    CK(cudaEventRecord(sp.ev_begin, stream));
    CK(cudaStreamWaitEvent(sp.side, sp.ev_begin, 0));              // the reset above, and whatever the caller queued before
    PROF_MARK(0);
    for (uint64_t k = 0; k < j.nslabs; k++) {
        if (int r = frame_job_encode_slab(j, k, stream)) return r;
        CK(cudaEventRecord(sp.ev_enc[k], stream));
        CK(cudaStreamWaitEvent(sp.side, sp.ev_enc[k], 0));
        // every slab but the last is compacted by one-warp CTAs that fit beside the next slab's codec kernel (four per
        // SM); the last slab's compaction runs alone at full width
        if (int r = frame_job_compact_slab(j, k, sp.side, (k + 1 < j.nslabs) ? 4 * sp.sms : 0)) return r;
    }
    PROF_MARK(1);
    PROF_MARK(2);
    CK(cudaEventRecord(sp.ev_done, sp.side));
    CK(cudaStreamWaitEvent(stream, sp.ev_done, 0));               // the exposed remainder: the last slab's compaction
    PROF_MARK(3);
    PROF_NEXT();
    return SPZ_OK;
}

// The directory of a frame is input from outside: it is validated on the device before any decoder runs, and the
// decoders bound every chunk by its directory span.  *d_err (optional) is OR-ed with 1 when the directory is bad.
static int frame_decode_params(DecParams& p, int predictor, int elem_sz, const void* d_frame, uint64_t frame_bytes,
                               uint64_t len, uint16_t ndims, uint32_t chunk_rows, void* d_dest, spz_chunk_meta* d_status)
{
    if (!valid_frame_codec(predictor, elem_sz) || ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    const FrameLayout f = frame_layout(len, ndims, chunk_rows);
    if (frame_bytes < f.payload_off + 16) return SPZ_E_STREAM;
    const uint8_t* frame = (const uint8_t*)d_frame;
    memset(&p, 0, sizeof(p));
    p.src = frame + f.payload_off; p.offsets = (const uint64_t*)(frame + f.off_dir);
    p.dst = (uint8_t*)d_dest; p.meta = (ChunkMeta*)d_status;
    p.len = len; p.chunk_len = f.chunk_len; p.nchunks = f.nchunks; p.dst_cap = len;
    p.ndims = ndims; p.has_meta = 1;
    p.src_end = frame + frame_bytes;
    return SPZ_OK;
}

extern "C" int spz_frame_decode_dev(int predictor, int elem_sz, const void* d_frame, uint64_t frame_bytes,
                                    uint64_t len, uint16_t ndims, uint32_t chunk_rows, void* d_dest,
                                    uint64_t dest_cap_elems, spz_chunk_meta* d_status, spz_stream_t stream)
{
    if (dest_cap_elems < len) return SPZ_E_CAPACITY;
    DecParams p;
    const int rc = frame_decode_params(p, predictor, elem_sz, d_frame, frame_bytes, len, ndims, chunk_rows, d_dest, d_status);
    if (rc != SPZ_OK) return rc;
    CK(launch_decode(elem_sz, predictor, p, (cudaStream_t)stream));
    return SPZ_OK;
}

extern "C" int spz_frame_decode_bs_dev(int predictor, int elem_sz, int block_rows, const void* d_frame, uint64_t frame_bytes,
                                       uint64_t len, uint16_t ndims, uint32_t chunk_rows, void* d_dest,
                                       uint64_t dest_cap_elems, spz_chunk_meta* d_status, spz_stream_t stream)
{
    if (dest_cap_elems < len) return SPZ_E_CAPACITY;
    if (!valid_block_rows(block_rows) || (block_rows != 8 && predictor > 1)) return SPZ_E_ARG;
    DecParams p;
    const int rc = frame_decode_params(p, predictor, elem_sz, d_frame, frame_bytes, len, ndims, chunk_rows, d_dest, d_status);
    if (rc != SPZ_OK) return rc;
    p.block_rows = (uint32_t)block_rows;
    CK(launch_decode(elem_sz, predictor, p, (cudaStream_t)stream));
    return SPZ_OK;
}

// ------------------------------------------------------------------ query-on-compressed
// query.hpp:22-29 QueryParams{op, materialize}; sprintz_delta.h:92-98, sprintz_xff.h:86-93.

static int g_last_query_fused = 0;
extern "C" int spz_last_query_fused(void) { return g_last_query_fused; }

extern "C" uint64_t spz_frame_query_workspace_bytes(int elem_sz, uint64_t len, int materialize)
{
    return materialize ? 0 : align16(len * (uint64_t)elem_sz) + 64;
}

extern "C" int spz_frame_query_dev(int predictor, int elem_sz, int op, int materialize, const void* d_frame,
                                   uint64_t frame_bytes, uint64_t len, uint16_t ndims, uint32_t chunk_rows,
                                   void* d_dest, uint64_t dest_cap_elems, spz_chunk_meta* d_status,
                                   uint64_t* d_result, void* d_work, uint64_t work_bytes, spz_stream_t stream)
{
    if (!valid_frame_codec(predictor, elem_sz) || ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    if (op < SPZ_Q_NOOP || op > SPZ_Q_SUM) return SPZ_E_ARG;
    if (op != SPZ_Q_NOOP && !d_result) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    g_last_query_fused = 0;
    if (op != SPZ_Q_NOOP) CK(cudaMemsetAsync(d_result, 0, 8ull * ndims, st));
    if (!materialize && op != SPZ_Q_NOOP) {
        const FrameLayout f = frame_layout(len, ndims, chunk_rows);
        if (frame_bytes < f.payload_off + 16) return SPZ_E_STREAM;
        const uint8_t* frame = (const uint8_t*)d_frame;
        DecParams p;
        memset(&p, 0, sizeof(p));
        p.src = frame + f.payload_off; p.offsets = (const uint64_t*)(frame + f.off_dir);
        p.dst = nullptr; p.meta = (ChunkMeta*)d_status;
        p.len = len; p.chunk_len = f.chunk_len; p.nchunks = f.nchunks; p.dst_cap = len;
        p.ndims = ndims; p.has_meta = 1;
        p.src_end = frame + frame_bytes;
        p.qop = (uint32_t)op; p.qresult = (unsigned long long*)d_result;
        const cudaError_t e = launch_query_fused(elem_sz, predictor, p, st);
        if (e == cudaSuccess) { g_last_query_fused = 1; return SPZ_OK; }
        if (e != cudaErrorNotSupported) CK(e);
        (void)cudaGetLastError();
    }
    // materialised: decode with the regular kernels (into d_dest, or the workspace when the caller does not want
    // the elements), then reduce the decoded rows
    void* out = d_dest;
    uint64_t cap = dest_cap_elems;
    if (!materialize) {
        if (op == SPZ_Q_NOOP) return SPZ_OK;           // nothing to produce
        if (!d_work || ((uintptr_t)d_work & 15) || work_bytes < align16(len * (uint64_t)elem_sz)) return SPZ_E_CAPACITY;
        out = d_work; cap = len;
    }
    if (!out) return SPZ_E_ARG;
    const int rc = spz_frame_decode_dev(predictor, elem_sz, d_frame, frame_bytes, len, ndims, chunk_rows, out, cap,
                                        d_status, stream);
    if (rc != SPZ_OK) return rc;
    if (op != SPZ_Q_NOOP)
        CK(launch_reduce_rows(op, elem_sz, (const uint8_t*)out, len, ndims, (unsigned long long*)d_result, st));
    return SPZ_OK;
}

extern "C" int spz_reduce_rows_dev(int op, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                                   uint64_t* d_result, spz_stream_t stream)
{
    if ((op != SPZ_Q_MAX && op != SPZ_Q_SUM) || (elem_sz != 1 && elem_sz != 2) || ndims == 0 || !d_result) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    if ((uintptr_t)d_src & 15) return SPZ_E_ARG;
    CK(cudaMemsetAsync(d_result, 0, 8ull * ndims, (cudaStream_t)stream));
    CK(launch_reduce_rows(op, elem_sz, (const uint8_t*)d_src, len, ndims, (unsigned long long*)d_result, (cudaStream_t)stream));
    return SPZ_OK;
}

extern "C" int spz_frame_read_header(const void* d_frame, spz_frame_header* out, spz_stream_t stream)
{
    CK(cudaMemcpyAsync(out, d_frame, sizeof(*out), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return out->magic == SPZ_FRAME_MAGIC ? SPZ_OK : SPZ_E_STREAM;
}

// ------------------------------------------------------------------ host-buffer entry points

namespace {
// Cached device buffers for the host-pointer calls (grown on demand, guarded by one mutex: the
// reference is re-entrant, so are we, but concurrent host calls serialise on the device anyway).
struct DevCache {
    std::mutex mu;
    void* buf[3] = {nullptr, nullptr, nullptr};
    uint64_t cap[3] = {0, 0, 0};
    cudaStream_t stream = nullptr;
    cudaError_t ensure(int i, uint64_t bytes)
    {
        if (cap[i] >= bytes) return cudaSuccess;
        if (buf[i]) { cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
        uint64_t want = bytes + (bytes >> 3) + 4096;
        cudaError_t e = cudaMalloc(&buf[i], want);
        if (e == cudaSuccess) cap[i] = want;
        return e;
    }
    cudaError_t get_stream(cudaStream_t* s)
    {
        if (!stream) { cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
        *s = stream; return cudaSuccess;
    }
    void release()
    {
        for (int i = 0; i < 3; i++) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
        if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    }
};
DevCache g_cache;
}  // namespace

static void release_pipe();
extern "C" void spz_release_workspaces(void)
{
    std::lock_guard<std::mutex> lk(g_cache.mu);
    g_cache.release();
    release_pipe();
}

extern "C" int spz_measure_stream(int elem_sz, const void* src_, int has_meta, uint16_t ndims, uint32_t ngroups,
                                  uint16_t remaining_len, uint64_t* stream_bytes, uint64_t* decoded_elems)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint64_t pos = 0;
    if (has_meta) {
        ngroups = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
        remaining_len = (uint16_t)(src[4] | (src[5] << 8));
        ndims = (uint16_t)(src[6] | (src[7] << 8));
        pos = 8;
    }
    uint64_t elems = 0;
    if (!(ngroups == 0 && remaining_len < kMinDataSize)) {
        if (ndims == 0) return SPZ_E_ARG;
        const uint32_t D = ndims, w = 8u * elem_sz, hb = elem_sz == 1 ? 3 : 4;
        const bool low = elem_sz == 1 ? (D <= 4) : (D <= 2);
        const uint32_t H = (2 * D * hb + 7) / 8;
        for (uint32_t g = 0; g < ngroups; g++) {
            const uint8_t* hdr = src + pos;
            pos += H;
            for (uint32_t s = 0; s < 2; s++) {
                uint32_t total = 0;
                for (uint32_t d = 0; d < D; d++) {
                    const uint32_t bit = (s * D + d) * hb;
                    uint32_t f = (((uint32_t)hdr[bit >> 3] | ((uint32_t)hdr[(bit >> 3) + 1] << 8)) >> (bit & 7)) & ((1u << hb) - 1);
                    total += (f == w - 1) ? w : f;
                }
                if (total == 0) {
                    uint32_t n = src[pos] & 0x7f;
                    if (src[pos] & 0x80) { n |= (uint32_t)src[pos + 1] << 7; pos += 2; } else pos += 1;
                    elems += (uint64_t)n * 8 * D;
                } else {
                    pos += low ? total : 8ull * ((total + 7) / 8);
                    elems += 8ull * D;
                }
            }
        }
    }
    pos += (uint64_t)remaining_len * elem_sz;
    elems += remaining_len;
    if (stream_bytes) *stream_bytes = pos;
    if (decoded_elems) *decoded_elems = elems;
    return SPZ_OK;
}

static int64_t compress_host_impl(uint32_t variant, int predictor, int elem_sz, const void* src, uint32_t len, void* dest,
                                  uint16_t ndims, int write_size, uint64_t* true_nbytes);

extern "C" int64_t spz_compress_host(int predictor, int elem_sz, const void* src, uint32_t len, void* dest,
                                     uint16_t ndims, int write_size, uint64_t* true_nbytes)
{
    return compress_host_impl(0, predictor, elem_sz, src, len, dest, ndims, write_size, true_nbytes);
}

static int64_t compress_host_impl(uint32_t variant, int predictor, int elem_sz, const void* src, uint32_t len, void* dest,
                                  uint16_t ndims, int write_size, uint64_t* true_nbytes)
{
    if (ndims == 0) { printf("Received invalid ndims %d\n", (int)ndims); return -1; }   // sprintz.cpp:71
    if (!valid_codec(predictor, elem_sz)) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    const uint64_t in_bytes = (uint64_t)len * elem_sz;
    const uint64_t out_cap = spz_max_compressed_bytes(elem_sz, len, ndims);
    CK(g_cache.ensure(0, in_bytes + 16));
    CK(g_cache.ensure(1, out_cap + 16));
    CK(g_cache.ensure(2, 64));
    if (in_bytes) CK(cudaMemcpyAsync(g_cache.buf[0], src, in_bytes, cudaMemcpyHostToDevice, st));
    int rc = compress_dev_impl(variant, predictor, elem_sz, g_cache.buf[0], len, ndims, write_size, g_cache.buf[1],
                               (spz_chunk_meta*)g_cache.buf[2], st);
    if (rc != SPZ_OK) return rc;
    spz_chunk_meta m;
    CK(cudaMemcpyAsync(&m, g_cache.buf[2], sizeof(m), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (m.status != 0) return SPZ_E_CAPACITY;
    CK(cudaMemcpyAsync(dest, g_cache.buf[1], m.nbytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (true_nbytes) *true_nbytes = m.nbytes;
    return (int64_t)(m.nbytes / (uint32_t)elem_sz);    // pointer difference in int_t units
}

static int64_t decompress_host_impl(int predictor, int elem_sz, const void* src, void* dest, int has_meta,
                                    uint16_t ndims, uint32_t ngroups, uint16_t remaining_len,
                                    int qop = 0, int materialize = 1, uint64_t* qresult = nullptr)
{
    if (!valid_codec(predictor, elem_sz)) return SPZ_E_ARG;
    if (qop < SPZ_Q_NOOP || qop > SPZ_Q_SUM || (qop != SPZ_Q_NOOP && !qresult)) return SPZ_E_ARG;
    uint64_t sbytes = 0, elems = 0;
    if (has_meta) {
        const uint8_t* s8 = (const uint8_t*)src;
        ngroups = (uint32_t)s8[0] | ((uint32_t)s8[1] << 8) | ((uint32_t)s8[2] << 16) | ((uint32_t)s8[3] << 24);
        remaining_len = (uint16_t)(s8[4] | (s8[5] << 8));
        ndims = (uint16_t)(s8[6] | (s8[7] << 8));
    }
    const bool raw = (ngroups == 0 && remaining_len < kMinDataSize);
    if (ndims == 0 && !raw) return 0;                              // sprintz_delta_rle.cpp:475-478
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    int rc = spz_measure_stream(elem_sz, src, has_meta, ndims, ngroups, remaining_len, &sbytes, &elems);
    if (rc != SPZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    CK(g_cache.ensure(0, sbytes + 32));
    CK(g_cache.ensure(1, elems * elem_sz + 16));
    CK(g_cache.ensure(2, 64 + 8ull * (ndims ? ndims : 1)));
    CK(cudaMemcpyAsync(g_cache.buf[0], src, sbytes, cudaMemcpyHostToDevice, st));
    spz_chunk_meta m;
    memset(&m, 0, sizeof(m));
    m.ngroups = ngroups; m.remaining = remaining_len;
    CK(cudaMemcpyAsync(g_cache.buf[2], &m, sizeof(m), cudaMemcpyHostToDevice, st));
    rc = spz_decompress_dev(predictor, elem_sz, g_cache.buf[0], sbytes + 16, has_meta, ndims ? ndims : 1,
                            g_cache.buf[1], elems, (spz_chunk_meta*)g_cache.buf[2], st);
    if (rc != SPZ_OK) return rc;
    CK(cudaMemcpyAsync(&m, g_cache.buf[2], sizeof(m), cudaMemcpyDeviceToHost, st));
    if (qop != SPZ_Q_NOOP) {
        // a single stream is one serial decode: the query is the reduction of its decoded rows
        const uint32_t D = ndims ? ndims : 1;
        unsigned long long* d_res = (unsigned long long*)((uint8_t*)g_cache.buf[2] + 64);
        CK(cudaMemsetAsync(d_res, 0, 8ull * D, st));
        CK(launch_reduce_rows(qop, elem_sz, (const uint8_t*)g_cache.buf[1], elems, D, d_res, st));
        CK(cudaMemcpyAsync(qresult, d_res, 8ull * D, cudaMemcpyDeviceToHost, st));
    }
    if (elems && materialize) CK(cudaMemcpyAsync(dest, g_cache.buf[1], elems * elem_sz, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (m.status != 0) return SPZ_E_STREAM;
    return (int64_t)elems;
}

// query_rowmajor_{delta,xff}_rle_{8b,16b}(src, dest, QueryParams) (sprintz_delta.h:95-98, sprintz_xff.h:90-93):
// QueryParams is flattened to (op, materialize); `result` receives ndims u64 (unsigned max / sum per dim).
extern "C" int64_t spz_query_host(int predictor, int elem_sz, const void* src, void* dest, int op, int materialize,
                                  uint64_t* result)
{
    if (materialize && !dest) return SPZ_E_ARG;
    return decompress_host_impl(predictor, elem_sz, src, dest, 1, 0, 0, 0, op, materialize, result);
}
extern "C" int64_t spz_query_rowmajor_delta_rle_8b(const int8_t* s, uint8_t* d, int op, int mat, uint64_t* r) { return spz_query_host(SPZ_DELTA, 1, s, d, op, mat, r); }
extern "C" int64_t spz_query_rowmajor_delta_rle_16b(const int16_t* s, uint16_t* d, int op, int mat, uint64_t* r) { return spz_query_host(SPZ_DELTA, 2, s, d, op, mat, r); }
extern "C" int64_t spz_query_rowmajor_xff_rle_8b(const int8_t* s, uint8_t* d, int op, int mat, uint64_t* r) { return spz_query_host(SPZ_XFF, 1, s, d, op, mat, r); }
extern "C" int64_t spz_query_rowmajor_xff_rle_16b(const int16_t* s, uint16_t* d, int op, int mat, uint64_t* r) { return spz_query_host(SPZ_XFF, 2, s, d, op, mat, r); }

extern "C" int64_t spz_decompress_host(int predictor, int elem_sz, const void* src, void* dest)
{
    return decompress_host_impl(predictor, elem_sz, src, dest, 1, 0, 0, 0);
}
extern "C" int64_t spz_decompress_raw_host(int predictor, int elem_sz, const void* src, void* dest,
                                           uint16_t ndims, uint32_t ngroups, uint16_t remaining_len)
{
    return decompress_host_impl(predictor, elem_sz, src, dest, 0, ndims, ngroups, remaining_len);
}

// the eight reference-named entry points (sprintz.h:16-32)
extern "C" int64_t spz_compress_delta_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return spz_compress_host(SPZ_DELTA, 1, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_delta_8b(const int8_t* s, uint8_t* d) { return spz_decompress_host(SPZ_DELTA, 1, s, d); }
extern "C" int64_t spz_compress_xff_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return spz_compress_host(SPZ_XFF, 1, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_xff_8b(const int8_t* s, uint8_t* d) { return spz_decompress_host(SPZ_XFF, 1, s, d); }
extern "C" int64_t spz_compress_delta_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return spz_compress_host(SPZ_DELTA, 2, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_delta_16b(const int16_t* s, uint16_t* d) { return spz_decompress_host(SPZ_DELTA, 2, s, d); }
extern "C" int64_t spz_compress_xff_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return spz_compress_host(SPZ_XFF, 2, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_xff_16b(const int16_t* s, uint16_t* d) { return spz_decompress_host(SPZ_XFF, 2, s, d); }


// ------------------------------------------------------------------ non-RLE sibling codecs
// compress_rowmajor_delta_{8b,16b} / compress_rowmajor_{8b,16b} and their decoders (sprintz_delta.h:27-44).
// kind 1 = delta, 2 = plain bit packing.  Same kernels as the RLE codecs (generic family, variant flag).

static int measure_rowmajor(int kind, int elem_sz, const uint8_t* src, uint64_t* stream_bytes, uint64_t* decoded_elems, uint16_t* ndims_out)
{
    const uint32_t len = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    const uint32_t at = kind == 3 ? 6u : 4u;                  // kind 3: [u48 len][u16 ndims]
    const uint32_t D = (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8);
    uint64_t pos = at + 2;
    *ndims_out = (uint16_t)D;
    uint64_t done = 0;
    if (len >= kMinDataSize) {
        if (D == 0) return SPZ_E_ARG;
        const uint32_t w = 8u * elem_sz, hb = elem_sz == 1 ? 3 : 4;
        const uint32_t H = (2 * D * hb + 7) / 8;
        const uint64_t ngroups = len / (16ull * D);
        for (uint64_t g = 0; g < ngroups; g++) {
            const uint8_t* hdr = src + pos;
            pos += H;
            for (uint32_t s = 0; s < 2; s++) {
                uint32_t total = 0;
                for (uint32_t d = 0; d < D; d++) {
                    const uint32_t bit = (s * D + d) * hb;
                    uint32_t f = (((uint32_t)hdr[bit >> 3] | ((uint32_t)hdr[(bit >> 3) + 1] << 8)) >> (bit & 7)) & ((1u << hb) - 1);
                    total += (f == w - 1) ? w : f;
                }
                pos += 8ull * ((total + 7) / 8);
            }
        }
        done = ngroups * 16ull * D;
    }
    pos += (uint64_t)(len - done) * elem_sz;
    *stream_bytes = pos;
    *decoded_elems = len;
    return SPZ_OK;
}

extern "C" int64_t spz_compress_rowmajor_host(int kind, int elem_sz, const void* src, uint32_t len, void* dest,
                                              uint16_t ndims, int write_size, uint64_t* true_nbytes)
{
    if (kind < 1 || kind > 3 || (kind == 3 && elem_sz != 1)) return SPZ_E_ARG;
    return compress_host_impl((uint32_t)kind, kind == 3 ? SPZ_XFF : SPZ_DELTA, elem_sz, src, len, dest, ndims, write_size, true_nbytes);
}

extern "C" int64_t spz_decompress_rowmajor_host(int kind, int elem_sz, const void* src, void* dest)
{
    if (kind < 1 || kind > 3 || (kind == 3 && elem_sz != 1) || !valid_codec(SPZ_DELTA, elem_sz)) return SPZ_E_ARG;
    uint64_t sbytes = 0, elems = 0;
    uint16_t ndims = 0;
    int rc = measure_rowmajor(kind, elem_sz, (const uint8_t*)src, &sbytes, &elems, &ndims);
    if (rc != SPZ_OK) return rc == SPZ_E_ARG ? 0 : rc;            // ndims == 0: the reference returns 0
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    CK(g_cache.ensure(0, sbytes + 32));
    CK(g_cache.ensure(1, elems * elem_sz + 16));
    CK(g_cache.ensure(2, 64));
    CK(cudaMemcpyAsync(g_cache.buf[0], src, sbytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(g_cache.buf[2], 0, sizeof(spz_chunk_meta), st));
    rc = decompress_dev_impl((uint32_t)kind, kind == 3 ? SPZ_XFF : SPZ_DELTA, elem_sz, g_cache.buf[0], sbytes + 16, 1, ndims ? ndims : 1,
                             g_cache.buf[1], elems, (spz_chunk_meta*)g_cache.buf[2], st);
    if (rc != SPZ_OK) return rc;
    spz_chunk_meta m;
    CK(cudaMemcpyAsync(&m, g_cache.buf[2], sizeof(m), cudaMemcpyDeviceToHost, st));
    if (elems) CK(cudaMemcpyAsync(dest, g_cache.buf[1], elems * elem_sz, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (m.status != 0) return SPZ_E_STREAM;
    return (int64_t)elems;
}

extern "C" int64_t spz_compress_rowmajor_delta_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return spz_compress_rowmajor_host(1, 1, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_rowmajor_delta_8b(const int8_t* s, uint8_t* d) { return spz_decompress_rowmajor_host(1, 1, s, d); }
extern "C" int64_t spz_compress_rowmajor_delta_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return spz_compress_rowmajor_host(1, 2, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_rowmajor_delta_16b(const int16_t* s, uint16_t* d) { return spz_decompress_rowmajor_host(1, 2, s, d); }
extern "C" int64_t spz_compress_rowmajor_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return spz_compress_rowmajor_host(2, 1, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress_rowmajor_8b(const int8_t* s, uint8_t* d) { return spz_decompress_rowmajor_host(2, 1, s, d); }
extern "C" int64_t spz_compress_rowmajor_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return spz_compress_rowmajor_host(2, 2, s, n, d, nd, ws, nullptr); }
extern "C" int64_t spz_compress8b_rowmajor_xff(const uint8_t* s, uint64_t n, int8_t* d, uint16_t nd, int ws) { return n > 0xffffffffull ? SPZ_E_ARG : spz_compress_rowmajor_host(3, 1, s, (uint32_t)n, d, nd, ws, nullptr); }
extern "C" int64_t spz_decompress8b_rowmajor_xff(const int8_t* s, uint8_t* d) { return spz_decompress_rowmajor_host(3, 1, s, d); }
extern "C" int64_t spz_decompress_rowmajor_16b(const int16_t* s, uint16_t* d) { return spz_decompress_rowmajor_host(2, 2, s, d); }

// ------------------------------------------------------------------ stand-alone forecaster transforms
// encode/decode_{delta,doubledelta}_rowmajor_{8b,16b} (delta.h:17-67).  order 1 = delta, 2 = double delta.

extern "C" uint64_t spz_forecast_workspace_bytes(uint64_t len, uint16_t ndims)
{
    return forecast_workspace_bytes(len, ndims) + 16;
}

extern "C" int spz_forecast_encode_dev(int order, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                                       void* d_dest, spz_stream_t stream)
{
    if ((order != 1 && order != 2) || (elem_sz != 1 && elem_sz != 2) || ndims == 0) return SPZ_E_ARG;
    if (((uintptr_t)d_src & 15) || ((uintptr_t)d_dest & 15)) return SPZ_E_ARG;
    CK(launch_forecast_encode(order, elem_sz, (const uint8_t*)d_src, (uint8_t*)d_dest, len, ndims, (cudaStream_t)stream));
    return SPZ_OK;
}

extern "C" int spz_forecast_decode_dev(int order, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims,
                                       void* d_dest, void* d_work, uint64_t work_bytes, spz_stream_t stream)
{
    if ((order != 1 && order != 2) || (elem_sz != 1 && elem_sz != 2) || ndims == 0) return SPZ_E_ARG;
    if (((uintptr_t)d_src & (elem_sz - 1)) || ((uintptr_t)d_dest & (elem_sz - 1)) || ((uintptr_t)d_work & 7)) return SPZ_E_ARG;
    if (work_bytes < forecast_workspace_bytes(len, ndims)) return SPZ_E_CAPACITY;
    CK(launch_forecast_decode(order, elem_sz, (const uint8_t*)d_src, (uint8_t*)d_dest, len, ndims,
                              (unsigned long long*)d_work, nullptr, (cudaStream_t)stream));
    return SPZ_OK;
}

// host buffers: returns len + metadata elements (encode) / len (decode), like the reference
extern "C" int64_t spz_encode_forecast_host(int order, int elem_sz, const void* src, uint32_t len, void* dest,
                                            uint16_t ndims, int write_size)
{
    if ((order != 1 && order != 2) || (elem_sz != 1 && elem_sz != 2) || ndims == 0) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    const uint64_t bytes = (uint64_t)len * elem_sz;
    CK(g_cache.ensure(0, bytes + 32));
    CK(g_cache.ensure(1, bytes + 32));
    uint8_t* d8 = (uint8_t*)dest;
    uint32_t meta = 0;
    if (write_size) {                                   // format.h:65-73
        d8[0] = (uint8_t)len; d8[1] = (uint8_t)(len >> 8); d8[2] = (uint8_t)(len >> 16); d8[3] = (uint8_t)(len >> 24);
        d8[4] = (uint8_t)ndims; d8[5] = (uint8_t)(ndims >> 8);
        d8 += 6; meta = 6u / (uint32_t)elem_sz;
    }
    if (bytes) {
        CK(cudaMemcpyAsync(g_cache.buf[0], src, bytes, cudaMemcpyHostToDevice, st));
        CK(launch_forecast_encode(order, elem_sz, (const uint8_t*)g_cache.buf[0], (uint8_t*)g_cache.buf[1], len, ndims, st));
        CK(cudaMemcpyAsync(d8, g_cache.buf[1], bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return (int64_t)len + meta;
}

extern "C" int64_t spz_decode_forecast_host(int order, int elem_sz, const void* src, uint32_t len, void* dest,
                                            uint16_t ndims)
{
    if ((order != 1 && order != 2) || (elem_sz != 1 && elem_sz != 2)) return SPZ_E_ARG;
    if (ndims == 0) return order == 2 ? 0 : SPZ_E_ARG;  // delta.cpp:631
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    const uint64_t bytes = (uint64_t)len * elem_sz;
    CK(g_cache.ensure(0, bytes + 32));
    CK(g_cache.ensure(1, bytes + 32));
    CK(g_cache.ensure(2, forecast_workspace_bytes(len, ndims) + 64));
    if (bytes) {
        CK(cudaMemcpyAsync(g_cache.buf[0], src, bytes, cudaMemcpyHostToDevice, st));
        CK(launch_forecast_decode(order, elem_sz, (const uint8_t*)g_cache.buf[0], (uint8_t*)g_cache.buf[1], len, ndims,
                                  (unsigned long long*)g_cache.buf[2], nullptr, st));
        CK(cudaMemcpyAsync(dest, g_cache.buf[1], bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return (int64_t)len;
}

// stand-alone FIRE transform (predict.h:15-30): one thread per dim (see spz_forecast.cu)
extern "C" int spz_xff_forecast_dev(int enc, int elem_sz, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest,
                                    spz_stream_t stream)
{
    if ((elem_sz != 1 && elem_sz != 2) || ndims == 0) return SPZ_E_ARG;
    if (((uintptr_t)d_src & (elem_sz - 1)) || ((uintptr_t)d_dest & (elem_sz - 1))) return SPZ_E_ARG;
    CK(launch_xff_forecast(enc, elem_sz, (const uint8_t*)d_src, (uint8_t*)d_dest, len, ndims, (cudaStream_t)stream));
    return SPZ_OK;
}

// host buffers: encode returns len + metadata elements, decode returns len, like the reference
extern "C" int64_t spz_xff_forecast_host(int enc, int elem_sz, const void* src, uint32_t len, void* dest, uint16_t ndims,
                                         int write_size)
{
    if (elem_sz != 1 && elem_sz != 2) return SPZ_E_ARG;
    if (ndims == 0) return enc ? SPZ_E_ARG : 0;         // predict.cpp:318
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    const uint64_t bytes = (uint64_t)len * elem_sz;
    CK(g_cache.ensure(0, bytes + 32));
    CK(g_cache.ensure(1, bytes + 32));
    uint8_t* d8 = (uint8_t*)dest;
    uint32_t meta = 0;
    if (enc && write_size) {                            // format.h:65-73
        d8[0] = (uint8_t)len; d8[1] = (uint8_t)(len >> 8); d8[2] = (uint8_t)(len >> 16); d8[3] = (uint8_t)(len >> 24);
        d8[4] = (uint8_t)ndims; d8[5] = (uint8_t)(ndims >> 8);
        d8 += 6; meta = 6u / (uint32_t)elem_sz;
    }
    if (bytes) {
        CK(cudaMemcpyAsync(g_cache.buf[0], src, bytes, cudaMemcpyHostToDevice, st));
        CK(launch_xff_forecast(enc, elem_sz, (const uint8_t*)g_cache.buf[0], (uint8_t*)g_cache.buf[1], len, ndims, st));
        CK(cudaMemcpyAsync(d8, g_cache.buf[1], bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return (int64_t)len + meta;
}
extern "C" uint32_t spz_encode_xff_rowmajor_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return (uint32_t)spz_xff_forecast_host(1, 1, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_xff_rowmajor_8b(const int8_t* s, uint32_t n, uint8_t* d, uint16_t nd) { return (uint32_t)spz_xff_forecast_host(0, 1, s, n, d, nd, 0); }
extern "C" uint32_t spz_encode_xff_rowmajor_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return (uint32_t)spz_xff_forecast_host(1, 2, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_xff_rowmajor_16b(const int16_t* s, uint32_t n, uint16_t* d, uint16_t nd) { return (uint32_t)spz_xff_forecast_host(0, 2, s, n, d, nd, 0); }

extern "C" uint32_t spz_encode_delta_rowmajor_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return (uint32_t)spz_encode_forecast_host(1, 1, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_delta_rowmajor_8b(const int8_t* s, uint32_t n, uint8_t* d, uint16_t nd) { return (uint32_t)spz_decode_forecast_host(1, 1, s, n, d, nd); }
extern "C" uint32_t spz_encode_delta_rowmajor_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return (uint32_t)spz_encode_forecast_host(1, 2, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_delta_rowmajor_16b(const int16_t* s, uint32_t n, uint16_t* d, uint16_t nd) { return (uint32_t)spz_decode_forecast_host(1, 2, s, n, d, nd); }
extern "C" uint32_t spz_encode_doubledelta_rowmajor_8b(const uint8_t* s, uint32_t n, int8_t* d, uint16_t nd, int ws) { return (uint32_t)spz_encode_forecast_host(2, 1, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_doubledelta_rowmajor_8b(const int8_t* s, uint32_t n, uint8_t* d, uint16_t nd) { return (uint32_t)spz_decode_forecast_host(2, 1, s, n, d, nd); }
extern "C" uint32_t spz_encode_doubledelta_rowmajor_16b(const uint16_t* s, uint32_t n, int16_t* d, uint16_t nd, int ws) { return (uint32_t)spz_encode_forecast_host(2, 2, s, n, d, nd, ws); }
extern "C" uint32_t spz_decode_doubledelta_rowmajor_16b(const int16_t* s, uint32_t n, uint16_t* d, uint16_t nd) { return (uint32_t)spz_decode_forecast_host(2, 2, s, n, d, nd); }

// ------------------------------------------------------------------ float -> integer quantiser
// python/datasets/compress_bench.py:45-60 (_quantize, axis = 0); kernels in spz_quant.cu.

extern "C" uint64_t spz_quantize_workspace_bytes(uint16_t ndims) { return quantize_workspace_bytes(ndims); }

extern "C" int spz_quantize_dev(int elem_sz, int src_f64, const void* d_src, uint64_t len, uint16_t ndims, void* d_dest,
                                void* d_work, uint64_t work_bytes, spz_stream_t stream)
{
    if ((elem_sz != 1 && elem_sz != 2) || ndims == 0 || (len && (!d_src || !d_dest))) return SPZ_E_ARG;
    if (!d_work || ((uintptr_t)d_work & 7) || work_bytes < quantize_workspace_bytes(ndims)) return SPZ_E_CAPACITY;
    CK(launch_quantize(elem_sz, src_f64 ? 1 : 0, d_src, len, ndims, d_dest, d_work, (cudaStream_t)stream));
    return SPZ_OK;
}

extern "C" int64_t spz_quantize_host(int elem_sz, int src_f64, const void* src, uint64_t len, uint16_t ndims, void* dest)
{
    if ((elem_sz != 1 && elem_sz != 2) || ndims == 0) return SPZ_E_ARG;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    cudaStream_t st;
    CK(g_cache.get_stream(&st));
    const uint64_t in_bytes = len * (src_f64 ? 8ull : 4ull);
    CK(g_cache.ensure(0, in_bytes + 16));
    CK(g_cache.ensure(1, len * elem_sz + 16));
    CK(g_cache.ensure(2, quantize_workspace_bytes(ndims) + 16));
    if (len) {
        CK(cudaMemcpyAsync(g_cache.buf[0], src, in_bytes, cudaMemcpyHostToDevice, st));
        CK(launch_quantize(elem_sz, src_f64 ? 1 : 0, g_cache.buf[0], len, ndims, g_cache.buf[1], g_cache.buf[2], st));
        CK(cudaMemcpyAsync(dest, g_cache.buf[1], len * elem_sz, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return (int64_t)len;
}

// ------------------------------------------------------------------ frame, host buffers
//
// Both calls are software pipelines over SLABS of whole chunks on three streams (upload / kernels /
// download), so the PCIe transfers of neighbouring slabs overlap in both directions and with the kernels:
//   compress : H2D raw slab k+1 | encode + directory scan + compaction of slab k | D2H payload of slab k-1
//   decompress: H2D payload slab k+1 | decode slab k | D2H decoded slab k-1
// The chunk streams and the directory are exactly those of spz_frame_encode_dev (the scan carries the running
// payload offset from slab to slab). With pageable host memory the copies serialise inside the driver but stay
// correct; pinned host buffers make them asynchronous.

namespace {
constexpr uint64_t kSlabBytes = 64ull << 20;     // raw bytes per pipeline slab

struct Pipe {
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    static constexpr int kMaxPipe = (int)kMaxSlabs;
    cudaEvent_t ev_up[kMaxPipe], ev_run[kMaxPipe];
    int nev = 0;
    unsigned long long* h_tot = nullptr;          // pinned: running payload offset after each slab
    unsigned int* h_err = nullptr;                // pinned: the encode job's error word
    cudaError_t init(int nslabs)
    {
        cudaError_t e;
        if (!up) {
            if ((e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&run, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaMallocHost(&h_tot, sizeof(unsigned long long) * kMaxPipe)) != cudaSuccess) return e;
            if ((e = cudaMallocHost(&h_err, 64)) != cudaSuccess) return e;
        }
        for (; nev < nslabs && nev < kMaxPipe; nev++) {
            if ((e = cudaEventCreateWithFlags(&ev_up[nev], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_run[nev], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // every exit of the host frame calls goes through here: nothing may still be reading the caller's buffers or
    // the cached device buffers once the call returns and the lock is released
    void drain()
    {
        if (up) { cudaStreamSynchronize(up); cudaStreamSynchronize(run); cudaStreamSynchronize(down); }
    }
    void release()
    {
        for (int i = 0; i < nev; i++) { cudaEventDestroy(ev_up[i]); cudaEventDestroy(ev_run[i]); }
        nev = 0;
        if (up) { cudaStreamDestroy(up); cudaStreamDestroy(run); cudaStreamDestroy(down); up = run = down = nullptr; }
        if (h_tot) { cudaFreeHost(h_tot); h_tot = nullptr; }
        if (h_err) { cudaFreeHost(h_err); h_err = nullptr; }
    }
};
Pipe g_pipe;
struct DrainOnExit { ~DrainOnExit() { g_pipe.drain(); } };
}  // namespace

extern "C" int64_t spz_frame_compress_host(int predictor, int elem_sz, const void* src_, uint64_t len, uint16_t ndims,
                                           uint32_t chunk_rows, void* frame_, uint64_t frame_cap)
{
    if (!valid_frame_codec(predictor, elem_sz) || ndims == 0 || chunk_rows == 0) return SPZ_E_ARG;
    if (ndims > SPZ_MAX_NDIMS) return SPZ_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* hframe = (uint8_t*)frame_;
    const FrameLayout f = frame_layout(len, ndims, chunk_rows);
    if (f.chunk_len > 0xffffffffull) return SPZ_E_ARG;
    const WorkLayout w = work_layout(elem_sz, len, ndims, chunk_rows);
    const uint64_t in_bytes = len * (uint64_t)elem_sz;
    const uint64_t fmax = spz_frame_max_bytes(elem_sz, len, ndims, chunk_rows);
    if (frame_cap < f.payload_off + 16) return SPZ_E_CAPACITY;
    const uint64_t chunk_bytes = f.chunk_len * elem_sz;
    uint64_t slab_chunks = kSlabBytes / (chunk_bytes ? chunk_bytes : 1);
    if (slab_chunks == 0) slab_chunks = 1;
    CK(g_cache.ensure(0, in_bytes + 16));
    CK(g_cache.ensure(1, fmax + 16));
    CK(g_cache.ensure(2, w.total + 16));
    uint8_t* d_in = (uint8_t*)g_cache.buf[0];
    uint8_t* d_frame = (uint8_t*)g_cache.buf[1];
    FrameJob j;
    int rc = frame_job_init(j, predictor, elem_sz, d_in, len, ndims, chunk_rows, d_frame, fmax + 16, g_cache.buf[2], w.total + 16,
                            nullptr, slab_chunks);
    if (rc != SPZ_OK) return rc;
    CK(g_pipe.init((int)j.nslabs));
    DrainOnExit drain;                                  // error exits included
    if ((rc = frame_job_reset(j, g_pipe.run)) != SPZ_OK) return rc;

    uint64_t copied = 0;        // payload bytes already on their way to the host
    for (uint64_t k = 0; k <= j.nslabs; k++) {
        if (k < j.nslabs) {
            uint64_t c0, nc, e0, ne;
            j.slab(k, c0, nc, e0, ne);
            if (ne) CK(cudaMemcpyAsync(d_in + e0 * elem_sz, src + e0 * elem_sz, ne * elem_sz, cudaMemcpyHostToDevice, g_pipe.up));
            CK(cudaEventRecord(g_pipe.ev_up[k], g_pipe.up));
            CK(cudaStreamWaitEvent(g_pipe.run, g_pipe.ev_up[k], 0));
            if ((rc = frame_job_encode_slab(j, k, g_pipe.run)) != SPZ_OK) return rc;
            if ((rc = frame_job_compact_slab(j, k, g_pipe.run)) != SPZ_OK) return rc;
            CK(cudaMemcpyAsync(&g_pipe.h_tot[k], j.d_tot() + k + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_pipe.run));
            if (k + 1 == j.nslabs)
                CK(cudaMemcpyAsync(g_pipe.h_err, j.d_err(), sizeof(unsigned int), cudaMemcpyDeviceToHost, g_pipe.run));
            CK(cudaEventRecord(g_pipe.ev_run[k], g_pipe.run));
        }
        if (k >= 1) {       // slab k-1 is compacted: its payload range can leave while slab k runs
            CK(cudaEventSynchronize(g_pipe.ev_run[k - 1]));
            const uint64_t upto = g_pipe.h_tot[k - 1];
            if (f.payload_off + upto + 16 > frame_cap) return SPZ_E_CAPACITY;
            CK(cudaStreamWaitEvent(g_pipe.down, g_pipe.ev_run[k - 1], 0));
            if (upto > copied)
                CK(cudaMemcpyAsync(hframe + f.payload_off + copied, d_frame + f.payload_off + copied, upto - copied,
                                   cudaMemcpyDeviceToHost, g_pipe.down));
            copied = upto;
        }
    }
    if (*g_pipe.h_err != 0) return fail_msg(SPZ_E_CAPACITY, "a chunk encoder overflowed its slot");
    // header + directory (written by the last slab's scan) and the 16 bytes of read slack
    CK(cudaMemcpyAsync(hframe, d_frame, f.payload_off, cudaMemcpyDeviceToHost, g_pipe.down));
    CK(cudaStreamSynchronize(g_pipe.down));
    CK(cudaStreamSynchronize(g_pipe.run));
    memset(hframe + f.payload_off + copied, 0, 16);
    return (int64_t)(f.payload_off + copied + 16);
}

extern "C" int64_t spz_frame_decompress_host(const void* frame_, uint64_t frame_bytes, void* dest_, uint64_t dest_cap_elems)
{
    if (frame_bytes < sizeof(spz_frame_header)) return SPZ_E_STREAM;
    const uint8_t* hframe = (const uint8_t*)frame_;
    uint8_t* dest = (uint8_t*)dest_;
    spz_frame_header h;
    memcpy(&h, hframe, sizeof(h));
    if (h.magic != SPZ_FRAME_MAGIC || h.frame_bytes > frame_bytes || !valid_frame_codec(h.predictor, h.elem_sz) ||
        h.ndims == 0 || h.ndims > SPZ_MAX_NDIMS || h.chunk_rows == 0) return SPZ_E_STREAM;
    if (dest_cap_elems < h.len) return SPZ_E_CAPACITY;
    if (h.len > (1ull << 48)) return SPZ_E_STREAM;                  // keeps the layout arithmetic far from overflow
    const FrameLayout f = frame_layout(h.len, (uint16_t)h.ndims, h.chunk_rows);
    if (f.nchunks != h.nchunks || f.payload_off != h.payload_off || h.frame_bytes < f.payload_off + 16 ||
        h.payload_bytes != h.frame_bytes - f.payload_off - 16) return SPZ_E_STREAM;
    // the directory is input from outside: every entry 16-byte aligned, non-decreasing, inside the payload
    const uint64_t* hoff = (const uint64_t*)(hframe + f.off_dir);
    if (hoff[0] != 0 || hoff[f.nchunks] != h.payload_bytes) return SPZ_E_STREAM;
    for (uint64_t c = 0; c < f.nchunks; c++)
        if ((hoff[c] & 15ull) || hoff[c + 1] < hoff[c] || hoff[c + 1] > h.payload_bytes) return SPZ_E_STREAM;
    std::lock_guard<std::mutex> lk(g_cache.mu);
    const int esz = h.elem_sz;
    const uint64_t chunk_bytes = f.chunk_len * esz;
    uint64_t slab_chunks = kSlabBytes / (chunk_bytes ? chunk_bytes : 1);
    if (slab_chunks == 0) slab_chunks = 1;
    uint64_t nslabs = (f.nchunks + slab_chunks - 1) / slab_chunks;
    if (nslabs > (uint64_t)Pipe::kMaxPipe) { slab_chunks = (f.nchunks + Pipe::kMaxPipe - 1) / Pipe::kMaxPipe; nslabs = (f.nchunks + slab_chunks - 1) / slab_chunks; }
    CK(g_pipe.init((int)nslabs));
    CK(g_cache.ensure(0, h.len * esz + 16));
    CK(g_cache.ensure(1, h.frame_bytes + 32));
    CK(g_cache.ensure(2, sizeof(spz_chunk_meta) * f.nchunks + 16));
    uint8_t* d_out = (uint8_t*)g_cache.buf[0];
    uint8_t* d_frame = (uint8_t*)g_cache.buf[1];
    spz_chunk_meta* d_status = (spz_chunk_meta*)g_cache.buf[2];
    DrainOnExit drain;                                  // error exits included

    // header + directory first (small), then payload slabs
    CK(cudaMemcpyAsync(d_frame, hframe, f.payload_off, cudaMemcpyHostToDevice, g_pipe.up));
    for (uint64_t k = 0; k < nslabs; k++) {
        const uint64_t c0 = k * slab_chunks;
        const uint64_t nc = (c0 + slab_chunks <= f.nchunks) ? slab_chunks : (f.nchunks - c0);
        const uint64_t b0 = hoff[c0], b1 = hoff[c0 + nc] + (k + 1 == nslabs ? 16 : 0);
        CK(cudaMemcpyAsync(d_frame + f.payload_off + b0, hframe + f.payload_off + b0, b1 - b0, cudaMemcpyHostToDevice, g_pipe.up));
        CK(cudaEventRecord(g_pipe.ev_up[k], g_pipe.up));
        CK(cudaStreamWaitEvent(g_pipe.run, g_pipe.ev_up[k], 0));
        const uint64_t e0 = c0 * f.chunk_len;
        const uint64_t ne = (e0 + nc * f.chunk_len <= h.len) ? nc * f.chunk_len : (h.len - e0);
        DecParams p;
        memset(&p, 0, sizeof(p));
        p.src = d_frame + f.payload_off; p.offsets = (const uint64_t*)(d_frame + f.off_dir) + c0;
        p.dst = d_out + e0 * esz; p.meta = (ChunkMeta*)(d_status + c0);
        p.len = ne; p.chunk_len = f.chunk_len; p.nchunks = nc; p.dst_cap = ne;
        p.ndims = h.ndims; p.has_meta = 1;
        p.src_end = d_frame + f.payload_off + hoff[c0 + nc] + 16;      // 16 bytes of read slack: the next slab's first bytes or the frame's own slack
        CK(launch_decode(esz, h.predictor, p, g_pipe.run));
        CK(cudaEventRecord(g_pipe.ev_run[k], g_pipe.run));
        CK(cudaStreamWaitEvent(g_pipe.down, g_pipe.ev_run[k], 0));
        if (ne) CK(cudaMemcpyAsync(dest + e0 * esz, d_out + e0 * esz, ne * esz, cudaMemcpyDeviceToHost, g_pipe.down));
    }
    // per-chunk status: any failure fails the call
    static thread_local std::vector<spz_chunk_meta> hst;
    hst.resize(f.nchunks);
    CK(cudaStreamWaitEvent(g_pipe.down, g_pipe.ev_run[nslabs - 1], 0));
    CK(cudaMemcpyAsync(hst.data(), d_status, sizeof(spz_chunk_meta) * f.nchunks, cudaMemcpyDeviceToHost, g_pipe.down));
    CK(cudaStreamSynchronize(g_pipe.down));
    for (uint64_t c = 0; c < f.nchunks; c++) if (hst[c].status != 0) return SPZ_E_STREAM;
    return (int64_t)h.len;
}

static void release_pipe() { g_pipe.release(); }

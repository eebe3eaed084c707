// spz_frame.cu -- chunk directory (device-wide exclusive scan) and compaction kernels of the frame path.
//
// The reference has no notion of a frame: one call == one serial stream (sprintz.cpp:92-109) with a
// uint32_t length.  The frame is how that path is made data-parallel: chunk c is an independent reference
// stream, and the directory built here says where each one starts.
#include "spz_host.h"

namespace spz {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

// Single-pass exclusive scan (decoupled look-back) of the 16-byte-rounded chunk sizes of one *segment* of
// `nchunks` consecutive chunks (the whole frame, or one slab of the pipelined host path). All pointers are
// already shifted to the segment's first chunk. *base_in (optional) is the payload offset at which the
// segment starts, *total_out receives the offset after it; the last segment also writes the frame header.
// tile_state[] and *ticket must be zero on entry.
__global__ void __launch_bounds__(kScanThreads)
spz_scan_sizes_kernel(const ChunkMeta* __restrict__ meta, uint64_t nchunks,
                      uint64_t* __restrict__ chunk_off, uint32_t* __restrict__ chunk_nbytes,
                      unsigned long long* tile_state, unsigned int* ticket,
                      const unsigned long long* base_in, unsigned long long* total_out, int is_last,
                      spz_frame_header* hdr, spz_frame_header hdr_in, uint64_t* d_frame_bytes, unsigned int* err)
{
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_excl;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;

    unsigned long long v[kScanItems];
    unsigned long long tsum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const uint64_t idx = base + j;
        unsigned long long a = 0;
        if (idx < nchunks) {
            const uint32_t nb = meta[idx].nbytes;
            if (meta[idx].status != kStatusOk) atomicOr(err, 2u);        // an encoder lane gave up (slot capacity): the frame is invalid
            chunk_nbytes[idx] = nb;
            a = ((unsigned long long)nb + 15ull) & ~15ull;
        }
        v[j] = tsum;          // exclusive within the thread
        tsum += a;
    }
    // block-wide exclusive scan of the thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = tsum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        if (w < warp) warp_excl += s_warp[w];
        block_total += s_warp[w];
    }
    const unsigned long long thread_excl = warp_excl + incl - tsum;

    if (threadIdx.x == 0) {
        unsigned long long excl = 0;
        __threadfence();                             // this tile's error flags are visible before its aggregate is
        if (tile == 0) {
            atomicExch(&tile_state[0], kFlagPrefix | block_total);
        } else {
            atomicExch(&tile_state[tile], kFlagAgg | block_total);
            for (long long j = (long long)tile - 1; j >= 0; j--) {
                unsigned long long st;
                do { st = atomicAdd(&tile_state[j], 0ull); } while ((st >> 62) == 0);   // predecessor tickets are already running
                excl += st & kValMask;
                if ((st >> 62) == 2) break;
            }
            atomicExch(&tile_state[tile], kFlagPrefix | (excl + block_total));
        }
        s_excl = excl;
    }
    __syncthreads();
    const unsigned long long tile_excl = s_excl + (base_in ? *base_in : 0ull);
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const uint64_t idx = base + j;
        if (idx < nchunks) chunk_off[idx] = tile_excl + thread_excl + v[j];
    }
    // the tile holding the last chunk also closes the directory and writes the frame header
    const uint64_t last_tile = (nchunks - 1) / kScanTile;
    if (tile == last_tile && threadIdx.x == 0) {
        const unsigned long long total = tile_excl + block_total;
        chunk_off[nchunks] = total;
        if (total_out) *total_out = total;
        if (is_last) {
            // zero the alignment gap between the size table and the payload so frames are deterministic
            uint8_t* gap = reinterpret_cast<uint8_t*>(chunk_nbytes + nchunks);
            uint8_t* pay = reinterpret_cast<uint8_t*>(hdr) + hdr_in.payload_off;
            while (gap < pay) *gap++ = 0;
            for (int i = 0; i < 16; i++) pay[total + i] = 0;             // the frame's 16 bytes of read slack
            spz_frame_header h = hdr_in;
            h.payload_bytes = total;
            h.frame_bytes = h.payload_off + total + 16;
            // every earlier tile published its aggregate (after a fence) before this tile's look-back ended, so all
            // error flags are in: a frame with a failed chunk gets no magic (spz_frame_read_header / decode reject it)
            __threadfence();
            if (atomicOr(err, 0u) != 0u) h.magic = 0;
            *hdr = h;
            if (d_frame_bytes) *d_frame_bytes = h.frame_bytes;
        }
    }
}

// Compaction: chunk c's stream moves from its worst-case slot to payload + chunk_off[c].
// One warp per chunk, 128-bit copies; bytes past the stream's true end are written as zero.
__global__ void __launch_bounds__(256)
spz_gather_kernel(const uint8_t* __restrict__ slots, uint64_t slot_stride, uint64_t nchunks,
                  const uint64_t* __restrict__ chunk_off, const uint32_t* __restrict__ chunk_nbytes,
                  uint8_t* __restrict__ payload)
{
    const uint64_t c = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    const int lane = threadIdx.x & 31;
    const uint32_t nb = chunk_nbytes[c];
    const uint32_t units = (nb + 15u) >> 4;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(slots + c * slot_stride);
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(payload + chunk_off[c]);
    for (uint32_t u = lane; u < units; u += 32) {
        uint4 x = s4[u];
        if (u == units - 1 && (nb & 15u)) {
            uint32_t w[4] = {x.x, x.y, x.z, x.w};
            const uint32_t keep = nb & 15u;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t lo = 4u * k;
                if (keep <= lo) w[k] = 0;
                else if (keep < lo + 4) w[k] &= (1u << (8 * (keep - lo))) - 1u;
            }
            x = make_uint4(w[0], w[1], w[2], w[3]);
        }
        d4[u] = x;
    }
}

// The same compaction as ONE-warp CTAs with eight 128-bit loads in flight per lane (~70 registers x 32 threads): thin
// enough to sit in the registers the codec kernels leave free (the lane encoders fill 59 of the 64 K registers of an SM
// with 4 CTAs), so a slab's compaction can run BESIDE the next slab's codec kernel instead of displacing its CTAs.
__global__ void __launch_bounds__(32)
spz_gather_thin_kernel(const uint8_t* __restrict__ slots, uint64_t slot_stride, uint64_t nchunks,
                       const uint64_t* __restrict__ chunk_off, const uint32_t* __restrict__ chunk_nbytes,
                       uint8_t* __restrict__ payload)
{
    const int lane = threadIdx.x;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint32_t nb = chunk_nbytes[c];
        const uint32_t units = (nb + 15u) >> 4;
        const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(slots + c * slot_stride);
        uint4* __restrict__ d4 = reinterpret_cast<uint4*>(payload + chunk_off[c]);
        for (uint32_t u0 = 0; u0 < units; u0 += 256u) {
            uint4 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t u = u0 + (uint32_t)lane + 32u * k;
                if (u < units) x[k] = s4[u];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t u = u0 + (uint32_t)lane + 32u * k;
                if (u < units) {
                    uint4 v = x[k];
                    if (u == units - 1 && (nb & 15u)) {
                        uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        const uint32_t keep = nb & 15u;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t lo = 4u * q;
                            if (keep <= lo) w[q] = 0;
                            else if (keep < lo + 4) w[q] &= (1u << (8 * (keep - lo))) - 1u;
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    d4[u] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ directory of a gathered frame (multi-GPU)
// After the ranks' directory slices have landed in the destination frame, entry c still holds the offset inside its
// OWNER's payload.  Segment s covers global chunks [seg_lo[s], seg_lo[s+1]) and moved by seg_delta[s] bytes.
__global__ void __launch_bounds__(256)
spz_rebase_directory_kernel(uint64_t* __restrict__ chunk_off, uint64_t nchunks, const uint64_t* __restrict__ seg_lo,
                            const int64_t* __restrict__ seg_delta, uint32_t nsegs, uint64_t total_payload,
                            uint32_t* chunk_nbytes, spz_frame_header* hdr, spz_frame_header hdr_in)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchunks) {
        uint32_t lo = 0, hi = nsegs;                 // last segment with seg_lo <= c
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_lo[mid] <= c) lo = mid; else hi = mid; }
        chunk_off[c] = (uint64_t)((int64_t)chunk_off[c] + seg_delta[lo]);
    }
    if (c == nchunks) {
        chunk_off[nchunks] = total_payload;
        uint8_t* gap = reinterpret_cast<uint8_t*>(chunk_nbytes + nchunks);
        uint8_t* pay = reinterpret_cast<uint8_t*>(hdr) + hdr_in.payload_off;
        while (gap < pay) *gap++ = 0;
        for (int i = 0; i < 16; i++) pay[total_payload + i] = 0;         // read slack
        spz_frame_header h = hdr_in;
        h.payload_bytes = total_payload;
        h.frame_bytes = h.payload_off + total_payload + 16;
        *hdr = h;
    }
}

cudaError_t launch_rebase_directory(uint64_t* chunk_off, uint64_t nchunks, const uint64_t* seg_lo, const int64_t* seg_delta,
                                    uint32_t nsegs, uint64_t total_payload, uint32_t* chunk_nbytes, spz_frame_header* hdr,
                                    const spz_frame_header& hdr_in, cudaStream_t stream)
{
    const uint64_t grid = (nchunks + 1 + 255) / 256;
    count_launch();
    spz_rebase_directory_kernel<<<(unsigned)grid, 256, 0, stream>>>(chunk_off, nchunks, seg_lo, seg_delta, nsegs, total_payload,
                                                                   chunk_nbytes, hdr, hdr_in);
    return cudaGetLastError();
}

cudaError_t launch_scan_sizes(const ChunkMeta* meta, uint64_t nchunks, uint64_t* chunk_off, uint32_t* chunk_nbytes,
                              unsigned long long* tile_state, unsigned int* ticket,
                              const unsigned long long* base_in, unsigned long long* total_out, int is_last,
                              spz_frame_header* hdr, const spz_frame_header& hdr_in, uint64_t* d_frame_bytes,
                              unsigned int* d_err, cudaStream_t stream)
{
    const uint64_t tiles = (nchunks + kScanTile - 1) / kScanTile;
    count_launch();
    spz_scan_sizes_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(meta, nchunks, chunk_off, chunk_nbytes,
                                                                      tile_state, ticket, base_in, total_out, is_last,
                                                                      hdr, hdr_in, d_frame_bytes, d_err);
    return cudaGetLastError();
}

uint64_t scan_tiles(uint64_t nchunks) { return (nchunks + kScanTile - 1) / kScanTile; }

cudaError_t launch_gather_thin(const uint8_t* slots, uint64_t slot_stride, uint64_t nchunks, const uint64_t* chunk_off,
                               const uint32_t* chunk_nbytes, uint8_t* payload, uint32_t ctas, cudaStream_t stream)
{
    if (ctas > nchunks) ctas = (uint32_t)nchunks;
    if (ctas == 0) return cudaSuccess;
    count_launch();
    spz_gather_thin_kernel<<<ctas, 32, 0, stream>>>(slots, slot_stride, nchunks, chunk_off, chunk_nbytes, payload);
    return cudaGetLastError();
}

cudaError_t launch_gather(const uint8_t* slots, uint64_t slot_stride, uint64_t nchunks, const uint64_t* chunk_off,
                          const uint32_t* chunk_nbytes, uint8_t* payload, cudaStream_t stream)
{
    const unsigned wpb = 8;
    const uint64_t grid = (nchunks + wpb - 1) / wpb;
    count_launch();
    spz_gather_kernel<<<(unsigned)grid, wpb * 32, 0, stream>>>(slots, slot_stride, nchunks, chunk_off, chunk_nbytes, payload);
    return cudaGetLastError();
}

}  // namespace spz

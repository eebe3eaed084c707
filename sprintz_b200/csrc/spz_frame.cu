// spz_frame.cu -- chunk directory (device-wide exclusive scan) and compaction kernels of the frame path.
//
// The reference has no notion of a frame: one call == one serial stream (sprintz.cpp:92-109) with a
// uint32_t length.  The frame is how that path is made data-parallel: chunk c is an independent reference
// stream, and the directory built here says where each one starts.
#include "spz_common.cuh"
#include "../../include/sprintz_b200.h"

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
                      spz_frame_header* hdr, spz_frame_header hdr_in, uint64_t* d_frame_bytes)
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
            spz_frame_header h = hdr_in;
            h.payload_bytes = total;
            h.frame_bytes = h.payload_off + total + 16;
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

// ------------------------------------------------------------------ overlapped directory + compaction ("chase" kernel)
//
// The compaction pass is HBM-bound (it moves N/r bytes twice) while the lane encoders are instruction-bound with
// DRAM 25-40 % busy, and run back to back they add up (17-26 % of frame-encode time).  This kernel does the work of
// spz_scan_sizes_kernel + spz_gather_kernel but is launched BEFORE the encoder on a second stream, with a grid small
// enough to stay resident next to it: CTA b owns tiles b, b + gridDim.x, ... of 128 consecutive chunks; a thread waits
// (acquire) for its chunk's record to be published by the encoder lane that owns it, the tile is scanned, chained to
// the earlier tiles by decoupled look-back, and copied from the worst-case slots (still in L2: they were just
// written) to its final place -- while later chunks are still being encoded.  Only the copy of the last wave of
// chunks is left exposed after the encoder finishes.  The encoder never waits on this kernel, so there is no cycle;
// every wait is bounded and raises *err instead of hanging.
constexpr int kChaseThreads = 128;
constexpr unsigned kChaseSpinLimit = 1u << 22;      // x >= 100 ns per poll: far beyond any encode time

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// FAST (uncapped probe mode only, unverified): the copy keeps four 128-bit loads in flight per lane.
template <bool FAST>
__global__ void __launch_bounds__(kChaseThreads)
spz_chase_kernel(const ChunkMeta* meta, uint64_t nchunks, const uint8_t* __restrict__ slots, uint64_t slot_stride,
                 uint64_t* __restrict__ chunk_off, uint32_t* __restrict__ chunk_nbytes, uint8_t* __restrict__ payload,
                 unsigned long long* tile_state, spz_frame_header* hdr, spz_frame_header hdr_in,
                 uint64_t* d_frame_bytes, unsigned int* err)
{
    __shared__ unsigned long long s_warp[kChaseThreads / 32];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned long long s_off[kChaseThreads];
    __shared__ uint32_t s_nb[kChaseThreads];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ntiles = (nchunks + kChaseThreads - 1) / kChaseThreads;

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t idx = tile * kChaseThreads + threadIdx.x;
        uint32_t nb = 0;
        if (idx < nchunks) {
            const unsigned int* flag = &meta[idx].nbytes;
            unsigned spins = 0;
            while ((nb = ld_acquire_u32(flag)) == 0u) {
                __nanosleep(200);
                if (++spins > kChaseSpinLimit) { atomicExch(err, 1u); break; }
            }
            chunk_nbytes[idx] = nb;
        }
        const unsigned long long a = ((unsigned long long)nb + 15ull) & ~15ull;
        // block-wide exclusive scan of the rounded sizes
        unsigned long long incl = a;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long warp_excl = 0, block_total = 0;
#pragma unroll
        for (int w = 0; w < kChaseThreads / 32; w++) {
            if (w < warp) warp_excl += s_warp[w];
            block_total += s_warp[w];
        }
        if (threadIdx.x == 0) {
            unsigned long long excl = 0;
            if (tile == 0) {
                atomicExch(&tile_state[0], kFlagPrefix | block_total);
            } else {
                atomicExch(&tile_state[tile], kFlagAgg | block_total);
                for (long long j = (long long)tile - 1; j >= 0; j--) {
                    unsigned long long st;
                    unsigned spins = 0;
                    while (((st = atomicAdd(&tile_state[j], 0ull)) >> 62) == 0) {
                        __nanosleep(100);
                        if (++spins > kChaseSpinLimit) { atomicExch(err, 2u); break; }
                    }
                    excl += st & kValMask;
                    if ((st >> 62) != 1) break;           // an inclusive prefix (or a timeout) ends the look-back
                }
                atomicExch(&tile_state[tile], kFlagPrefix | (excl + block_total));
            }
            s_excl = excl;
        }
        __syncthreads();
        const unsigned long long off = s_excl + warp_excl + incl - a;
        if (idx < nchunks) chunk_off[idx] = off;
        s_off[threadIdx.x] = off;
        s_nb[threadIdx.x] = nb;
        if (tile == ntiles - 1 && threadIdx.x == 0) {
            // closes the directory and writes the frame header (as spz_scan_sizes_kernel does for the last segment)
            const unsigned long long total = s_excl + block_total;
            chunk_off[nchunks] = total;
            uint8_t* gap = reinterpret_cast<uint8_t*>(chunk_nbytes + nchunks);
            uint8_t* pay = reinterpret_cast<uint8_t*>(hdr) + hdr_in.payload_off;
            while (gap < pay) *gap++ = 0;
            spz_frame_header h = hdr_in;
            h.payload_bytes = total;
            h.frame_bytes = h.payload_off + total + 16;
            if (atomicAdd(err, 0u) != 0u) h.magic = 0;      // a wait timed out: the frame is invalid (spz_frame_read_header fails)
            *hdr = h;
            if (d_frame_bytes) *d_frame_bytes = h.frame_bytes;
        }
        __syncthreads();
        // ---- copy the tile: one warp per chunk, 128-bit units, L2 loads (the slots were written by other SMs)
        for (int k = warp; k < kChaseThreads; k += kChaseThreads / 32) {
            const uint64_t c = tile * kChaseThreads + k;
            if (c >= nchunks) break;
            const uint32_t cnb = s_nb[k];
            const uint32_t units = (cnb + 15u) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(slots + c * slot_stride);
            uint4* d4 = reinterpret_cast<uint4*>(payload + s_off[k]);
            uint32_t u = lane;
            if (FAST && units > 128u) {
                // whole groups of 4 x 32 units that do not contain the chunk's last (possibly partial) unit
                for (; u + 96u + (31u - lane) < units - 1u; u += 128u) {
                    const uint4 x0 = __ldcg(s4 + u), x1 = __ldcg(s4 + u + 32), x2 = __ldcg(s4 + u + 64), x3 = __ldcg(s4 + u + 96);
                    d4[u] = x0; d4[u + 32] = x1; d4[u + 64] = x2; d4[u + 96] = x3;
                }
            }
            for (; u < units; u += 32) {
                uint4 x = __ldcg(s4 + u);
                if (u == units - 1 && (cnb & 15u)) {
                    uint32_t w[4] = {x.x, x.y, x.z, x.w};
                    const uint32_t keep = cnb & 15u;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t lo = 4u * q;
                        if (keep <= lo) w[q] = 0;
                        else if (keep < lo + 4) w[q] &= (1u << (8 * (keep - lo))) - 1u;
                    }
                    x = make_uint4(w[0], w[1], w[2], w[3]);
                }
                d4[u] = x;
            }
        }
        __syncthreads();                              // s_off / s_nb are reused by the next tile
    }
}

uint64_t chase_tiles(uint64_t nchunks) { return (nchunks + kChaseThreads - 1) / kChaseThreads; }

// tile_state: chase_tiles(nchunks) zeroed words; meta[].nbytes zeroed; *err zeroed.
cudaError_t launch_chase(const ChunkMeta* meta, uint64_t nchunks, const uint8_t* slots, uint64_t slot_stride,
                         uint64_t* chunk_off, uint32_t* chunk_nbytes, uint8_t* payload, unsigned long long* tile_state,
                         spz_frame_header* hdr, const spz_frame_header& hdr_in, uint64_t* d_frame_bytes, unsigned int* err,
                         bool fast, cudaStream_t stream)
{
    const uint64_t tiles = chase_tiles(nchunks);
    const unsigned grid = (unsigned)(tiles < 148 ? tiles : 148);       // one small CTA per SM: stays resident beside the encoder
    // Ask for the largest shared-memory carveout: this kernel needs 1.5 KB, but an SM keeps the carveout of the CTAs
    // resident on it, and the lane encoder needs 29 KB per CTA -- with a small carveout pinned on every SM by the
    // spinning chase CTAs the encoder could not be placed anywhere.  (The prime suspect for the stall seen when the
    // grid covered all 148 SMs; not yet re-measured, so spz_frame_encode_dev still limits this path to <= 148 tiles.)
    static bool carveout_set = false;
    if (!carveout_set) {
        cudaError_t ce = cudaFuncSetAttribute(spz_chase_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              (int)cudaSharedmemCarveoutMaxShared);
        if (ce != cudaSuccess) return ce;
        ce = cudaFuncSetAttribute(spz_chase_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)cudaSharedmemCarveoutMaxShared);
        if (ce != cudaSuccess) return ce;
        carveout_set = true;
    }
    if (fast)
        spz_chase_kernel<true><<<grid, kChaseThreads, 0, stream>>>(meta, nchunks, slots, slot_stride, chunk_off, chunk_nbytes,
                                                                  payload, tile_state, hdr, hdr_in, d_frame_bytes, err);
    else
        spz_chase_kernel<false><<<grid, kChaseThreads, 0, stream>>>(meta, nchunks, slots, slot_stride, chunk_off, chunk_nbytes,
                                                                   payload, tile_state, hdr, hdr_in, d_frame_bytes, err);
    return cudaGetLastError();
}

cudaError_t launch_scan_sizes(const ChunkMeta* meta, uint64_t nchunks, uint64_t* chunk_off, uint32_t* chunk_nbytes,
                              unsigned long long* tile_state, unsigned int* ticket,
                              const unsigned long long* base_in, unsigned long long* total_out, int is_last,
                              spz_frame_header* hdr, const spz_frame_header& hdr_in, uint64_t* d_frame_bytes,
                              cudaStream_t stream)
{
    const uint64_t tiles = (nchunks + kScanTile - 1) / kScanTile;
    spz_scan_sizes_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(meta, nchunks, chunk_off, chunk_nbytes,
                                                                      tile_state, ticket, base_in, total_out, is_last,
                                                                      hdr, hdr_in, d_frame_bytes);
    return cudaGetLastError();
}

uint64_t scan_tiles(uint64_t nchunks) { return (nchunks + kScanTile - 1) / kScanTile; }

cudaError_t launch_gather(const uint8_t* slots, uint64_t slot_stride, uint64_t nchunks, const uint64_t* chunk_off,
                          const uint32_t* chunk_nbytes, uint8_t* payload, cudaStream_t stream)
{
    const unsigned wpb = 8;
    const uint64_t grid = (nchunks + wpb - 1) / wpb;
    spz_gather_kernel<<<(unsigned)grid, wpb * 32, 0, stream>>>(slots, slot_stride, nchunks, chunk_off, chunk_nbytes, payload);
    return cudaGetLastError();
}

}  // namespace spz

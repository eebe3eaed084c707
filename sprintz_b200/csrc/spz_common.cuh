// spz_common.cuh -- shared device/host definitions for the B200 Sprintz path.
//
// Format facts restated from the reference (dblalock/sprintz cpp/Compress):
//   block = 8 rows x D dims, group = 2 slots (sprintz_delta_rle.cpp:50,63)
//   header field = 3 bits (8b) / 4 bits (16b), value w stored as w-1 (sprintz_delta_rle.cpp:61,199)
//   run slot = all-zero header fields + 1..2 length bytes, <= 0x7fff blocks (sprintz_delta_rle.cpp:68,238-245)
//   8-byte stream metadata [u32 ngroups][u16 remaining_len][u16 ndims] (format.h:35-62)
//   streams shorter than 128 elements are stored raw (sprintz_delta_rle.cpp:101-109)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace spz {

constexpr int      kBlockRows   = 8;
constexpr uint32_t kMaxRun      = 0x7fff;
constexpr uint32_t kMinDataSize = 128;
constexpr int      kThreads     = 128;     // CTA size of the codec kernels (4 warps)
constexpr int      kMaxDims     = 1024;    // supported ndims (see DESIGN.md "limits")

// per-chunk result record written by the encoder / status written by the decoder
struct ChunkMeta {
    uint32_t nbytes;      // exact byte length of the chunk's reference stream
    uint32_t ngroups;     // as stored in the stream metadata
    uint32_t remaining;   // trailing raw elements (full value; the stream stores it as u16)
    uint32_t status;      // 0 ok, else an SPZ_E* code
};

enum : uint32_t {
    kStatusOk = 0,
    kStatusBadDims = 1,       // stream's ndims disagrees with the call
    kStatusOverflow = 2,      // decoded output would exceed the destination capacity
    kStatusTruncated = 3,     // stream runs past its stated size
};

struct EncParams {
    const uint8_t* src;       // raw row-major elements (device)
    uint8_t*       dst;       // chunk c's stream is written at dst + c*slot_stride
    ChunkMeta*     meta;      // [nchunks]
    uint64_t       len;       // total elements
    uint64_t       chunk_len; // elements per chunk (last chunk ragged)
    uint64_t       nchunks;
    uint64_t       slot_stride;
    uint32_t       ndims;
    uint32_t       team;      // lanes per chunk, power of two <= 32
    uint32_t       dpl;       // dims per lane = ceil(ndims / team)
    uint32_t       write_size;
    uint32_t       team_smem; // bytes of shared memory per team
    uint32_t       nb_off, tile_off, tile_stride, out_off;   // layout inside team_smem
    uint32_t       out_cap;   // staging capacity (bytes)
    uint32_t       teams_per_cta;
    uint32_t       variant;   // 0: RLE codecs (sprintz.h) | 1: compress_rowmajor_delta_* | 2: compress_rowmajor_* (no RLE, 6-byte metadata)
    uint32_t       family_pin;// 0: the launcher picks the kernel family | f + 1: family f was chosen for the whole frame (slabs must not re-decide)
    uint32_t       block_rows;// rows per block: 0 / 8 = the reference's format; 16 / 32 / 64 = BASELINE config 5's block-size axis (generic kernels)
};

struct DecParams {
    const uint8_t*  src;       // stream(s)
    const uint64_t* offsets;   // chunk byte offsets into src (nullptr: c*slot_stride)
    uint8_t*        dst;       // decoded elements; chunk c goes to dst + c*chunk_len*esz
    ChunkMeta*      meta;      // optional [nchunks]: status out; for headerless streams ngroups/remaining in
    uint64_t        len;       // total elements expected
    uint64_t        chunk_len;
    uint64_t        nchunks;
    uint64_t        slot_stride;
    uint64_t        dst_cap;   // elements the destination can hold in total
    uint32_t        ndims;
    uint32_t        team;
    uint32_t        dpl;
    uint32_t        has_meta;  // 1: each stream starts with the 8-byte metadata
    uint32_t        team_smem;
    uint32_t        teams_per_cta;
    const uint8_t*  src_end;   // one past the last readable byte of the source buffer (fast path only)
    uint32_t        ring_cap;  // fast path: bytes of the per-team input ring (power of two)
    uint32_t        out_cap;   // fast path: bytes of the per-team output staging
    uint32_t        ring_mask4;// fast path: (ring_cap - 1) & ~3, as a kernel constant so wrap+base is one LOP3
    uint32_t        variant;   // as EncParams::variant
    uint32_t        qop;       // query-on-compressed (lane kernels only): 0 decode | 1 per-dim max | 2 per-dim sum; dst unused
    unsigned long long* qresult;   // [ndims] u64, zero-initialised by the caller
    uint32_t        block_rows;// as EncParams::block_rows
};

template <int W> struct Traits;
template <> struct Traits<8>  { typedef uint8_t  uint_t; static constexpr int HB = 3; static constexpr int CB = 16; };
template <> struct Traits<16> { typedef uint16_t uint_t; static constexpr int HB = 4; static constexpr int CB = 32; };

// ------------------------------------------------------------------ scalar arithmetic (bit-exact contract)

template <int BITS> __host__ __device__ __forceinline__ int32_t sx(int32_t v)
{
    return (int32_t)((uint32_t)v << (32 - BITS)) >> (32 - BITS);
}
template <> __host__ __device__ __forceinline__ int32_t sx<32>(int32_t v) { return v; }

// FIRE coefficient (sprintz_xff_rle.cpp:209-217 high-dim: 4 MSBs kept, int16_t; sprintz_xff_lowdim.cpp:170-173
// low-dim: counter >> 1, counter_t wide).
template <int W, bool LOW> __host__ __device__ __forceinline__ int32_t fire_coef(int32_t counter)
{
    if (LOW) return counter >> 1;
    constexpr int SH = W - 4;
    // 8-bit: the counter is an int16 (callers keep it sign-extended), so (counter >> 5) << 4 is already an int16
    if (W == 8) return (int32_t)((uint32_t)(counter >> (1 + SH)) << SH);
    return sx<16>((int32_t)((uint32_t)(counter >> (1 + SH)) << SH));
}

// prediction = (int_t)(((counter_t)prev_delta * coef) >> w), product wraps at 32 bits (sprintz_xff_rle.cpp:225)
template <int W> __host__ __device__ __forceinline__ int32_t fire_predict(int32_t prev_delta, int32_t coef)
{
    int32_t prod = (int32_t)((uint32_t)prev_delta * (uint32_t)coef);
    return sx<W>(prod >> W);
}

// icopysign(err, prev_delta) (util.h:63-68), W-bit
template <int W> __host__ __device__ __forceinline__ int32_t fire_grad(int32_t err, int32_t prev_delta)
{
    return err == 0 ? 0 : (err < 0 ? -prev_delta : prev_delta);   // caller wraps the sum to W bits
}

template <int W> __host__ __device__ __forceinline__ uint32_t zigzag_enc(int32_t e)   // bitpack.h:302
{
    return (((uint32_t)e << 1) ^ (uint32_t)(e >> 31)) & ((1u << W) - 1u);
}
template <int W> __host__ __device__ __forceinline__ int32_t zigzag_dec(uint32_t u)   // bitpack.h:303
{
    return sx<W>((int32_t)((u >> 1) ^ (0u - (u & 1u))));
}

// OR-mask -> field width with the variant's promotion rule
//   8b: 7->8 (bitpack.h:72-93); 16b high-dim: 7->8 and 15->16 (sprintz_delta_rle.cpp:177-187);
//   16b low-dim: only 15->16 (sprintz_delta_lowdim.cpp:172-173)
template <int W, bool LOW> __device__ __forceinline__ uint32_t nbits_rule(uint32_t mask)
{
    uint32_t n = 32u - (uint32_t)__clz((int)mask);
    if (n == (uint32_t)(W - 1)) n = W;
    if (W == 16 && !LOW && n == 7) n = 8;
    return n;
}

// A lane encoder hands its chunk to the chase kernel (spz_frame.cu) by writing the record's nbytes LAST, after a
// device-scope fence: a non-zero nbytes means every stream byte of the chunk is visible in L2.
__device__ __forceinline__ void publish_nbytes(uint32_t* p, uint32_t nbytes)
{
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(nbytes) : "memory");
}

// The output rings of the lane encoders are addressed with absolute 32-bit shared-memory addresses and explicit st/ld.shared:
// with C++ pointers the compiler rebuilt the shared window base (S2R + MOV + LEA) under every predicated store.
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.b8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// ------------------------------------------------------------------ team (sub-warp) primitives

__device__ __forceinline__ uint32_t team_mask_of(int lane, int team)
{
    return team == 32 ? 0xffffffffu : (((1u << team) - 1u) << (lane & ~(team - 1)));
}
__device__ __forceinline__ uint32_t team_incl_scan(uint32_t x, int t, int team, uint32_t tmask)
{
    for (int s = 1; s < team; s <<= 1) {
        uint32_t y = __shfl_up_sync(tmask, x, s, team);
        if (t >= s) x += y;
    }
    return x;
}
__device__ __forceinline__ uint32_t team_last(uint32_t x, int team, uint32_t tmask)
{
    return __shfl_sync(tmask, x, team - 1, team);
}

// 32-bit window of a little-endian bit stream starting at absolute bit `bit` of `base` (global or shared).
// Reads the two aligned words that cover it: callers guarantee 8 readable bytes past the last payload byte.
__device__ __forceinline__ uint32_t load_bits32(const uint8_t* base, uint32_t bit)
{
    const uint8_t* p = base + (bit >> 3);
    uintptr_t a = (uintptr_t)p;
    const uint32_t* wp = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)((a & 3) << 3) + (bit & 7);
    uint32_t w0 = wp[0], w1 = wp[1];
    return __funnelshift_r(w0, w1, sh);
}

// ------------------------------------------------------------------ shared-memory output staging
// One team's output bytes are staged in shared memory and leave for HBM as 128-bit stores.  buf[0]
// corresponds to the 16-byte aligned global address gdst; the first `skip` bytes (non-zero only before the
// first flush after (re)initialisation) lie before the region this team owns and are never written.
struct OutStage {
    uint8_t* buf;
    uint8_t* gdst;
    uint32_t cur;
    uint32_t skip;
    __device__ __forceinline__ void init(uint8_t* smem_buf, uint8_t* gptr)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)gptr & 15);
        buf = smem_buf; gdst = gptr - mis; skip = mis; cur = mis;
    }
};

__device__ __forceinline__ void stage_flush(OutStage& o, bool final_flush, int t, int team, uint32_t tmask)
{
    __syncwarp(tmask);
    const uint32_t hi = final_flush ? o.cur : (o.cur & ~15u);
    const uint32_t lo = o.skip;
    if (hi > lo) {
        uint32_t a = (lo + 15u) & ~15u;
        if (a > hi) a = hi;
        for (uint32_t b = lo + t; b < a; b += team) o.gdst[b] = o.buf[b];            // unaligned head
        const uint4* s4 = reinterpret_cast<const uint4*>(o.buf);
        uint4* g4 = reinterpret_cast<uint4*>(o.gdst);
        for (uint32_t u = (a >> 4) + t; u < (hi >> 4); u += team) g4[u] = s4[u];    // 128-bit body
        uint32_t tb = hi & ~15u;
        if (tb < a) tb = a;
        for (uint32_t b = tb + t; b < hi; b += team) o.gdst[b] = o.buf[b];           // tail (final only)
        __syncwarp(tmask);
        if (!final_flush) {
            const uint32_t rem = o.cur - hi;        // < 16, and hi >= 16: source and dest cannot overlap
            for (uint32_t b = t; b < rem; b += team) o.buf[b] = o.buf[hi + b];
            o.gdst += hi;
            o.cur = rem;
            o.skip = 0;
            __syncwarp(tmask);
        }
    }
}

}  // namespace spz

// spz_decode_wide.cu -- wide-row ("octet team") decode kernel: ndims a multiple of 8, 32..256; 8- and 16-bit elements,
// delta and FIRE.  The text below describes the 8-bit delta form (SIMD within a register); the FIRE and 16-bit forms share
// everything but the unpack / inverse predictor of a data slot (see the comment at the kernel).
//
// BASELINE config 4 (int8 x 80 dims, delta) ran on the team kernel of spz_decode_hd.cu at 0.38 of the roofline: ncu
// showed 26 thread instructions per sample, LSU pipe 65 % (a byte store per sample, two LDS + a funnel shift per
// field), 22 % of the shared wavefronts bank conflicts.  The delta codec needs no multiply, so here everything after the
// bit extraction is SIMD-within-a-register on FOUR samples per 32-bit word:
//
//   mapping  one lane owns 8 consecutive dims (an OCTET = two words per row) of one chunk; a chunk is decoded by a
//            team of ndims/8 lanes (10 for 80 dims), 32/(ndims/8) teams per warp, warp-synchronous like the team kernel
//   input    the compressed stream is pulled through a per-team shared-memory ring with coalesced 256-bit loads issued
//            one fill ahead
//   header   a lane reads its octet's 8 width fields as one 24-bit window, sums them with two multiplies and a popcount;
//            one packed shuffle scan over the team gives both slots' bit offsets
//   unpack   a quad of fields (<= 32 bits) is one window (2 LDS + funnel shift); field j is moved to byte j by a left
//            shift of 8j - (bits before it) >= 0 (as a multiply: the ALU pipe is the binding one), the four shifted copies are merged with three bit-selects: 6
//            instructions per 4 samples instead of 16
//   predict  zigzag^-1 on four bytes at once (mask, multiply by 0xff, shift, select-xor) and the running sum as a
//            carry-less byte-wise add: 9 instructions per 4 samples
//   output   8 rows x ndims bytes are staged in a shared tile (one 64-bit store per lane per row) and leave as
//            coalesced 128-bit stores that complete whole 32-byte sectors
//
// Reference behaviour restated: sprintz_delta_rle.cpp:417-772 (high-dim delta + RLE decode), sprintz_xff_rle.cpp:569-1179
// (high-dim FIRE + RLE decode).
#include "spz_host.h"

namespace spz {

constexpr int kWideThreads = 128;
constexpr unsigned kFullW = 0xffffffffu;

// per-quad constants of one slot: shifts that move field j to byte j, byte masks (split for the zigzag step)
struct QuadC { uint32_t m1, m2, m3, ones, half; };      // m_j = 2^(8j - bits before field j): the shifts run as multiplies (FMA pipe)

// f12: four 3-bit header fields.  Returns the quad's total width in bits.
__device__ __forceinline__ uint32_t quad_consts8(uint32_t f12, QuadC& q)
{
    const uint32_t f = (f12 & 7u) | ((f12 & 0x38u) << 5) | ((f12 & 0x1c0u) << 10) | ((f12 & 0xe00u) << 15);   // one field per byte
    const uint32_t n = f + (((f + 0x01010101u) >> 3) & 0x01010101u);      // width: 7 is stored for 8 (sprintz_delta_rle.cpp:199)
    const uint32_t pre = n * 0x01010100u;                                   // byte k = n_0 + ... + n_(k-1)  (<= 24)
    const uint32_t sl = 0x18100800u - pre;                                  // byte k = 8k - bits before field k  (>= 0)
    q.m1 = 1u << ((sl >> 8) & 0xffu); q.m2 = 1u << ((sl >> 16) & 0xffu); q.m3 = 1u << (sl >> 24);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) m |= ((1u << ((n >> (8 * k)) & 0xffu)) - 1u) << (8 * k);
    q.ones = m & 0x01010101u;
    q.half = (m >> 1) & 0x7f7f7f7fu;
    return (n * 0x01010101u) >> 24;
}

// sum of the 8 widths of an octet given its 24 header bits (fields + the number of fields equal to 7)
__device__ __forceinline__ uint32_t octet_width8(uint32_t fw)
{
    const uint32_t x = fw & 0x00ffffffu;
    const uint32_t e = (x & 0007070707u) + ((x >> 3) & 0007070707u);        // 4 six-bit lanes, each <= 14
    const uint32_t full = x & (x >> 1) & (x >> 2) & 0011111111u;            // bit 0 of a field: field == 7
    return (((e * 0x00041041u) >> 18) & 63u) + (uint32_t)__popc(full);
}

__device__ __forceinline__ uint32_t bsel(uint32_t a, uint32_t b, uint32_t c) { return (a & c) | (b & ~c); }   // one LOP3

// sum of the 8 widths of a 16-bit octet given its 32 header bits (fields + the number of fields equal to 15)
__device__ __forceinline__ uint32_t octet_width16(uint32_t fw)
{
    const uint32_t e = (fw & 0x0f0f0f0fu) + ((fw >> 4) & 0x0f0f0f0fu);      // 4 bytes, each <= 30
    const uint32_t full = fw & (fw >> 1) & (fw >> 2) & (fw >> 3) & 0x11111111u;
    return ((e * 0x01010101u) >> 24) + (uint32_t)__popc(full);
}

// p.team = lanes per chunk (ndims / 8), p.dpl = teams per warp, p.ring_cap / p.team_smem as in spz_decode_hd.cu
// FIRE: the inverse forecaster needs a multiply per sample, so a data slot is unpacked and predicted per element on
// left-aligned values exactly like the lane kernel's (spz_decode_lane.cu: one window per quad, one shift per field,
// zigzag^-1 and icopysign as multiply-adds) and the row is re-packed four samples to a word; stream ring, header scan,
// tiles and control are shared with the delta form.
// Registers: the 8-bit delta form is held to 80 (6 CTAs per SM, which is also what its shared memory allows): without a bound
// the compiler took 102 (5 CTAs, c4 decode 1.77 -> 1.82 ms), with 72 it schedules worse (1.88 ms); measured 1.79 ms.
template <int W, bool FIRE, bool HALF>
__global__ void __launch_bounds__(kWideThreads, (FIRE || W == 16) ? 4 : 6) spz_decode_wide_kernel(const DecParams p)
{
    static_assert(W == 8 || W == 16, "8- and 16-bit elements");
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr uint32_t ESZ = W / 8;
    constexpr uint32_t FMASK = (1u << HB) - 1u;
    constexpr bool ELEM = FIRE || W == 16;              // data slots are decoded per element (left-aligned values)
    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t NP = p.team, TPW = p.dpl;
    const uint32_t tw_raw = (uint32_t)lane / NP;
    const bool lane_on = tw_raw < TPW;                  // lanes past the last whole team of the warp idle along
    const uint32_t tw = lane_on ? tw_raw : 0u;
    const uint32_t t = lane_on ? (uint32_t)lane - tw * NP : 0u;
    const uint64_t c64 = ((uint64_t)blockIdx.x * (kWideThreads / 32) + warp) * TPW + tw;
    const bool valid = lane_on && c64 < p.nchunks;
    const uint64_t c = valid ? c64 : 0;
    const uint32_t D = NP * 8u;
    const uint32_t last_lane = tw * NP + NP - 1u;       // the lane holding my team's inclusive totals

    const uint32_t team_slot = (uint32_t)warp * TPW + tw;
    const uint32_t ring_off = team_slot * p.team_smem;  // team_smem is a multiple of ring_cap: wrap + base is one LOP3
    uint8_t* const ring = smem + ring_off;
    const uint32_t RCAP = p.ring_cap, RMASK = RCAP - 1u;
    const uint32_t TILE = 64u * NP * ESZ;               // 8 rows x D elements
    // two output tiles; team k's rows start where team k-1's end (mod the banks) so the 64-bit row stores of the
    // warp's teams fall into different banks
    uint8_t* const tile0 = ring + RCAP + 32 + ((tw * D * ESZ) & 112u);
#define WMASK4 (p.ring_mask4)

    // ---- input: byte offsets are relative to `abase`, the 32-byte aligned-down stream start
    const uint64_t soff = p.offsets ? p.offsets[c] : c * p.slot_stride;
    const uint64_t src_bytes = (uint64_t)(p.src_end - p.src);
    const bool bad_off = soff > src_bytes || src_bytes - soff < 8u;
    const uint8_t* const gstream = p.src + (bad_off ? 0 : soff);
    const uint32_t mis_in = (uint32_t)((uintptr_t)gstream & 31);
    const uint8_t* const abase = gstream - mis_in;
    const uint64_t lim64 = (uint64_t)(p.src_end - abase);
    const uint32_t limit = lim64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)lim64;   // readable bytes from abase
    const uint32_t FILL = NP * 32u;
    uint32_t wr = 0;                                    // stream bytes staged so far (team-uniform)
    uint32_t nx[8];
    auto prefetch = [&]() {
        const uint32_t o = wr + t * 32u;
        if (o + 32u <= limit) {
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(nx[0]), "=r"(nx[1]), "=r"(nx[2]), "=r"(nx[3]), "=r"(nx[4]), "=r"(nx[5]), "=r"(nx[6]), "=r"(nx[7])
                         : "l"(abase + o));
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) nx[j] = 0;
#pragma unroll 1
            for (uint32_t b = 0; b < 32u; b++)
                if (o + b < limit) nx[b >> 2] |= (uint32_t)abase[o + b] << (8 * (b & 3));
        }
    };
    // make stream bytes [.., need_end) available in the ring; warp-uniform call, teams that need nothing idle through
    auto ensure = [&](uint32_t need_end, bool on) {
        while (__any_sync(kFullW, on && wr < need_end)) {
            __syncwarp();                               // readers of the bytes about to be overwritten are done
            if (on && wr < need_end) {
                const uint32_t ro = (wr + t * 32u) & RMASK;
                *reinterpret_cast<uint4*>(ring + ro) = make_uint4(nx[0], nx[1], nx[2], nx[3]);
                *reinterpret_cast<uint4*>(ring + ro + 16) = make_uint4(nx[4], nx[5], nx[6], nx[7]);
                if (ro == 0) *reinterpret_cast<uint4*>(ring + RCAP) = make_uint4(nx[0], nx[1], nx[2], nx[3]);   // mirror: word reads run 4 bytes past the end
                wr += FILL;
                prefetch();
            }
            __syncwarp();
        }
    };
    // 32-bit window at ring-relative bit position `bit` (may exceed the ring: masked)
    auto ring_bits = [&](uint32_t bit) -> uint32_t {
        const uint32_t bo = ((bit >> 3) & WMASK4) | ring_off;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + bo);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + bo + 4);
        return __funnelshift_r(w0, w1, bit);
    };
    // 64-bit window (16-bit quads)
    auto ring_seg = [&](uint32_t bit, uint32_t& lo, uint32_t& hi) {
        const uint32_t bo = ((bit >> 3) & WMASK4) | ring_off;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + bo);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + bo + 4);
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(smem + bo + 8);
        lo = __funnelshift_r(w0, w1, bit);
        hi = __funnelshift_r(w1, w2, bit);
    };
    auto ring_byte = [&](uint32_t at) -> uint32_t { return ring[at & RMASK]; };

    // ---- output
    const uint64_t first = c * p.chunk_len;
    uint8_t* const out = p.dst + first * ESZ;
    const uint64_t cap64 = (!valid || first >= p.dst_cap) ? 0 :
        ((p.dst_cap - first < p.chunk_len) ? (p.dst_cap - first) : p.chunk_len);
    const uint32_t cap = cap64 > 0xffffffffull ? 0xffffffffu : (uint32_t)cap64;
    uint32_t tsel = 0;                                  // which tile the next data slot uses

    uint32_t xA = 0, xB = 0;                            // the previous row's values of my 8 dims (4 bytes per word)
    uint32_t pvh[ELEM ? 8 : 1];                         // per-element form: previous values / deltas left-aligned, FIRE counters (sprintz_xff_rle.cpp:683-688)
    int32_t pdh[FIRE ? 8 : 1], ct[FIRE ? 8 : 1];
#pragma unroll
    for (int j = 0; j < (ELEM ? 8 : 1); j++) pvh[j] = 0;
#pragma unroll
    for (int j = 0; j < (FIRE ? 8 : 1); j++) { pdh[j] = 0; ct[j] = 0; }
    constexpr int E1 = ELEM ? 1 : 0;
    auto store_row = [&](uint8_t* dst) {                // the top W bits of my 8 left-aligned values
        if (W == 8) {
            const uint32_t a0 = __byte_perm(pvh[0], pvh[1 * E1], 0x0073), a1 = __byte_perm(pvh[2 * E1], pvh[3 * E1], 0x0073);
            const uint32_t b0 = __byte_perm(pvh[4 * E1], pvh[5 * E1], 0x0073), b1 = __byte_perm(pvh[6 * E1], pvh[7 * E1], 0x0073);
            *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(a0, a1, 0x5410), __byte_perm(b0, b1, 0x5410));
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(__byte_perm(pvh[0], pvh[1 * E1], 0x7632), __byte_perm(pvh[2 * E1], pvh[3 * E1], 0x7632),
                                                        __byte_perm(pvh[4 * E1], pvh[5 * E1], 0x7632), __byte_perm(pvh[6 * E1], pvh[7 * E1], 0x7632));
        }
    };

    prefetch();
    uint32_t pos = mis_in;                              // consume offset from abase (team-uniform)
    uint32_t ngroups = 0, remaining = 0, status = kStatusOk;
    if (p.has_meta) {                                   // format.h:47-62
        ensure(pos + 8, lane_on);
        ngroups = ring_byte(pos) | (ring_byte(pos + 1) << 8) | (ring_byte(pos + 2) << 16) | (ring_byte(pos + 3) << 24);
        remaining = ring_byte(pos + 4) | (ring_byte(pos + 5) << 8);
        const uint32_t nd = ring_byte(pos + 6) | (ring_byte(pos + 7) << 8);
        if (nd != D) status = kStatusBadDims;
        pos += 8;
    } else {
        ngroups = p.meta[c].ngroups;
        remaining = p.meta[c].remaining & 0xffffu;
    }
    const uint32_t H = (2u * D * HB + 7u) >> 3;
    if (bad_off) { status = kStatusTruncated; remaining = 0; }
    {   // a group is at least H + 2 bytes: a group count the chunk's own span cannot hold is corrupt (and would spin)
        const uint64_t next = p.offsets ? p.offsets[c + 1] : src_bytes;
        const uint64_t cspan = (next > soff && next <= src_bytes) ? next - soff : (src_bytes > soff ? src_bytes - soff : 0);
        if ((uint64_t)ngroups * (H + 2u) > cspan) { status = kStatusTruncated; remaining = 0; }
    }
    if (!valid || status != kStatusOk) ngroups = 0;

    uint32_t o = 0;                                     // output element cursor (team-uniform)
    const uint32_t B = 8u * D;

    // One data slot: unpack + inverse predictor for my octet, rows staged in a tile, tile flushed with whole sectors.
    // Warp-uniform call; `doit` = my team really has this slot (state and global memory only change then).
    auto data_slot = [&](uint32_t fw, uint32_t offA, uint32_t total, bool doit) {
        if (ELEM) {
            // one tile; a 16-bit slot is pulled through the ring in two halves of 4 rows (the ring stays as small as the
            // 8-bit one: shared memory bounds the occupancy of this kernel)
            const uint32_t row_bytes = (total + 7u) >> 3;
            const uint32_t row_bits = row_bytes * 8u;
            uint8_t* const tile = tile0;
            uint32_t nb[8], fm[8];
            int32_t coef[FIRE ? 8 : 1], gsh[FIRE ? 8 : 1];
            uint32_t wA = 0;
            uint32_t bitA = (pos & RMASK) * 8u + offA;
            uint8_t* trow = tile + t * 8u * ESZ;
            __syncwarp();                               // the previous slot's tile has left
            auto rows4 = [&](uint32_t r0) {
#pragma unroll
                for (uint32_t r4 = 0; r4 < 4; r4++) {
                    const uint32_t r = r0 + r4;
                    // 8-bit: one 32-bit window per quad of fields; 16-bit: one 64-bit window
                    uint32_t lo, hi = 0;
                    if (W == 16) ring_seg(bitA, lo, hi); else lo = ring_bits(bitA);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (j == 4) { if (W == 16) ring_seg(bitA + wA, lo, hi); else lo = ring_bits(bitA + wA); }
                        // zigzag^-1 in doubled form: m = u ^ -(u & 1) == 2 * err; with b = u & 1 and tt = 1 - 2b, m = u * tt - b
                        const uint32_t u = lo & fm[j];
                        const uint32_t b = lo & fm[j] & 1u;
                        if (j != 7 && j != 3) {
                            if (W == 16) { lo = __funnelshift_r(lo, hi, nb[j]); hi >>= nb[j]; } else { lo >>= nb[j]; }
                        }
                        int32_t tt, mi;
                        asm("mad.lo.s32 %0, %1, -2, 1;" : "=r"(tt) : "r"(b));
                        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(mi) : "r"(u), "r"(tt), "r"(0u - b));
                        if (FIRE) {
                            const int f = FIRE ? j : 0;
                            if (r & 1u) { if (mi != 0) gsh[f] += pdh[f] * tt; }      // icopysign(err, prev_delta) (util.h:63-68)
                            // delta = ((prev_delta * coef) >> W) + err (sprintz_xff_rle.cpp:225), everything left-aligned
                            const int32_t th = __mulhi(pdh[f], coef[f]);
                            int32_t ths;
                            asm("mad.lo.s32 %0, %1, %2, 0;" : "=r"(ths) : "r"(th), "n"(1 << UP));
                            asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(pdh[f]) : "r"(mi), "n"(1 << (UP - 1)), "r"(ths));
                            pvh[j] += (uint32_t)pdh[f];
                        } else {
                            asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(pvh[j]) : "r"(mi), "n"(1 << (UP - 1)));
                        }
                    }
                    bitA += row_bits;
                    store_row(trow);
                    trow += D * ESZ;
                }
            };
            if (doit) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t f = (fw >> (j * HB)) & FMASK;
                    const uint32_t n = f + (f == FMASK ? 1u : 0u);   // W - 1 is stored for W (sprintz_xff_rle.cpp:296)
                    nb[j] = n;
                    if (j < 4) wA += n;
                    fm[j] = (1u << n) - 1u;
                    if (FIRE) { coef[FIRE ? j : 0] = fire_coef<W, false>(ct[FIRE ? j : 0]); gsh[FIRE ? j : 0] = 0; }
                }
                rows4(0);
            }
            if (HALF) ensure(pos + row_bytes * 8u + 16u, doit);     // second half of the slot (the caller staged the first)
            if (doit) {
                rows4(4);
                if (FIRE) {
#pragma unroll
                    for (int j = 0; j < (FIRE ? 8 : 1); j++) ct[j] = sx<CB>(ct[j] + ((gsh[j] >> UP) >> 2));   // sprintz_xff_rle.cpp:273-275
                }
            }
            __syncwarp();                               // the tile is complete
            if (doit) {
                const uint4* s4 = reinterpret_cast<const uint4*>(tile);
                uint4* g4 = reinterpret_cast<uint4*>(out + (size_t)o * ESZ);
#pragma unroll
                for (uint32_t i = 0; i < 4u * ESZ; i++) g4[t + i * NP] = s4[t + i * NP];
                pos += row_bytes * 8u;
                o += B;
            }
            return;
        }
        QuadC qa, qb;
        const uint32_t wA = quad_consts8(fw & 0xfffu, qa);
        (void)quad_consts8((fw >> 12) & 0xfffu, qb);
        const uint32_t row_bytes = (total + 7u) >> 3;
        const uint32_t row_bits = row_bytes * 8u;
        uint32_t bitA = (pos & RMASK) * 8u + offA;
        uint8_t* const tile = tile0 + tsel * (TILE + 128u);
        uint8_t* trow = tile + t * 8u;
        uint32_t a = xA, b = xB;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            // quad A: fields 0..3 of my octet; quad B: fields 4..7
            const uint32_t la = ring_bits(bitA);
            const uint32_t lb = ring_bits(bitA + wA);
            bitA += row_bits;
            const uint32_t wa = bsel(bsel(la, la * qa.m1, 0x000000ffu), bsel(la * qa.m2, la * qa.m3, 0x00ff0000u), 0x0000ffffu);
            const uint32_t wb = bsel(bsel(lb, lb * qb.m1, 0x000000ffu), bsel(lb * qb.m2, lb * qb.m3, 0x00ff0000u), 0x0000ffffu);
            // zigzag^-1 (bitpack.h:303) on four bytes: (u >> 1) ^ -(u & 1); bits above each field's width are masked off here
            const uint32_t da = ((wa >> 1) & qa.half) ^ ((wa & qa.ones) * 0xffu);
            const uint32_t db = ((wb >> 1) & qb.half) ^ ((wb & qb.ones) * 0xffu);
            // x += delta, byte-wise without carries between the bytes (sprintz_delta_rle.cpp:704-736)
            a = ((a & 0x7f7f7f7fu) + (da & 0x7f7f7f7fu)) ^ ((a ^ da) & 0x80808080u);
            b = ((b & 0x7f7f7f7fu) + (db & 0x7f7f7f7fu)) ^ ((b ^ db) & 0x80808080u);
            if (lane_on) *reinterpret_cast<uint2*>(trow) = make_uint2(a, b);   // idle lanes shadow team 0 / lane 0: they must not touch its tile
            trow += D;
        }
        __syncwarp();                                   // the tile is complete
        if (doit) {
            xA = a; xB = b;
            const uint4* s4 = reinterpret_cast<const uint4*>(tile);
            uint4* g4 = reinterpret_cast<uint4*>(out + o);
#pragma unroll
            for (int i = 0; i < 4; i++) g4[t + i * NP] = s4[t + i * NP];     // 4 x NP 16-byte units = 64 x NP bytes
            pos += row_bytes * 8u;
            o += B;
            tsel ^= 1u;
        }
    };

    // stream bytes a data slot of `tot` bits per row needs staged before data_slot(): the slot plus the window over-read;
    // 16-bit slots are staged in two halves (the second inside data_slot)
    auto slot_need = [&](uint32_t tot) -> uint32_t {
        return HALF ? ((tot + 7u) >> 3) * 4u + 16u : ((tot + 7u) & ~7u) + (W == 16 ? 16u : 8u);     // 32-bit windows over-read < 8 bytes, 64-bit ones < 12
    };
    const uint32_t gmax = __reduce_max_sync(kFullW, ngroups);
    for (uint32_t g = 0; g < gmax; g++) {
        if (g < ngroups && pos > limit) status = kStatusTruncated;      // ran past the readable bytes (the ring was fed zeros)
        const bool act = g < ngroups && status == kStatusOk;
        ensure(pos + H + 8, act);                       // the ring holds one header or one slot at a time (occupancy: half the ring)

        // ---- both slots' widths: my octet's fields, one packed scan over the team for the bit offsets
        const uint32_t hbit = (pos & RMASK) * 8u;
        const uint32_t fw0 = ring_bits(hbit + (8u * t) * HB);
        const uint32_t fw1 = ring_bits(hbit + (D + 8u * t) * HB);
        const uint32_t mine = W == 8 ? (octet_width8(fw0) | (octet_width8(fw1) << 16)) : (octet_width16(fw0) | (octet_width16(fw1) << 16));
        uint32_t incl = mine;
        for (uint32_t s = 1; s < NP; s <<= 1) {
            const uint32_t y = __shfl_up_sync(kFullW, incl, s);
            if (t >= s) incl += y;
        }
        const uint32_t tot = __shfl_sync(kFullW, incl, last_lane);
        const uint32_t excl = incl - mine;
        const uint32_t tot0 = tot & 0xffffu, tot1 = tot >> 16;
        if (act) pos += H;

        // Fast path: every live team of the warp has two data slots that fit.
        if (__all_sync(kFullW, !lane_on || (act && tot0 != 0u && tot1 != 0u && o + 2u * B <= cap))) {
            ensure(pos + slot_need(tot0), lane_on);
            data_slot(fw0, excl & 0xffffu, tot0, lane_on);
            ensure(pos + slot_need(tot1), lane_on);
            data_slot(fw1, excl >> 16, tot1, lane_on);
            continue;
        }

#pragma unroll 1
        for (uint32_t s = 0; s < 2; s++) {
            const uint32_t total = s ? tot1 : tot0;
            const bool is_run = act && status == kStatusOk && total == 0;
            ensure(pos + (is_run ? 10u : slot_need(total)), act && status == kStatusOk);
            if (is_run) {
                // ---- run slot (sprintz_delta_rle.cpp:640-667): the previous row repeats nrun * 8 times; rare, divergent
                uint32_t nrun = ring_byte(pos) & 0x7f;
                if (ring_byte(pos) & 0x80) { nrun |= ring_byte(pos + 1) << 7; pos += 2; } else { pos += 1; }
                const uint32_t nrows = nrun * 8u;
                if ((uint64_t)o + (uint64_t)nrows * D > cap) { status = kStatusOverflow; }
                else {
                    uint8_t* op = out + ((size_t)o + t * 8u) * ESZ;
                    if (ELEM) {     // errors of 0: every delta is its own prediction (sprintz_xff_rle.cpp:828-958), the counters rest;
                                    // delta coding repeats the previous row
                        int32_t coef[FIRE ? 8 : 1];
#pragma unroll
                        for (int j = 0; j < (FIRE ? 8 : 1); j++) coef[j] = fire_coef<W, false>(ct[j]);
#pragma unroll 1
                        for (uint32_t r = 0; r < nrows; r++) {
                            if (FIRE) {
#pragma unroll
                                for (int j = 0; j < (FIRE ? 8 : 1); j++) {
                                    const int32_t pred = __mulhi(pdh[j], coef[j]);
                                    pdh[j] = (int32_t)((uint32_t)pred << UP);
                                    pvh[j] += (uint32_t)pdh[j];
                                }
                            }
                            store_row(op);
                            op += D * ESZ;
                        }
                    } else {
                        for (uint32_t r = 0; r < nrows; r++) { *reinterpret_cast<uint2*>(op) = make_uint2(xA, xB); op += D; }
                    }
                    o += nrows * D;
                }
            }
            bool data = act && status == kStatusOk && total != 0;
            if (data && o + B > cap) { status = kStatusOverflow; data = false; }
            if (__any_sync(kFullW, data)) data_slot(s ? fw1 : fw0, s ? (excl >> 16) : (excl & 0xffffu), total, data);
        }
    }

    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (valid && status == kStatusOk) {
        if ((uint64_t)o + remaining > cap) {
            status = kStatusOverflow;
        } else if ((uint64_t)pos + (uint64_t)remaining * ESZ > (uint64_t)limit) {
            status = kStatusTruncated;                  // the raw tail is read straight from the source
        } else {
            const uint8_t* tsrc = abase + pos;
            uint8_t* tdst = out + (size_t)o * ESZ;
            for (uint32_t bI = t; bI < remaining * ESZ; bI += NP) tdst[bI] = tsrc[bI];
            o += remaining;
        }
    }
    if (valid && t == 0 && p.meta) {
        p.meta[c].status = status;
        if (p.has_meta) {
            p.meta[c].nbytes = pos - mis_in + remaining * ESZ;
            p.meta[c].ngroups = ngroups;
            p.meta[c].remaining = o;
        }
    }
#undef WMASK4
}

static inline uint32_t pow2_ceil_w(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }

// Returns cudaErrorNotSupported when the shape is not covered (the caller continues with the team kernels).
cudaError_t launch_decode_wide(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    if (get_force_path() == 1) return cudaErrorNotSupported;
    if ((elem_sz != 1 && elem_sz != 2) || (predictor != 0 && predictor != 1) || p.variant != 0 || p.qop != 0) return cudaErrorNotSupported;
    // 24 dims (three lanes per chunk, shared memory limits the occupancy): the team kernel measured faster (8-bit delta 1952
    // vs 1720 GB/s, FIRE 1328 vs 1287; 16-bit delta 1712 vs 840)
    if (p.ndims < 32u || p.ndims > 256 || (p.ndims & 7u)) return cudaErrorNotSupported;
    if (elem_sz == 2 && p.ndims == 32u) return cudaErrorNotSupported;   // one dim per lane of a full warp: team kernel 2739 vs 2060 GB/s
    if (p.src_end == nullptr || p.src_end <= p.src) return cudaErrorNotSupported;
    const uint32_t esz = (uint32_t)elem_sz;
    // 128-bit tile flushes: every block of every chunk must start 16-byte aligned in the output
    if (((uintptr_t)p.dst & 15) != 0) return cudaErrorNotSupported;
    if (p.nchunks > 1 && ((p.chunk_len * esz) & 15) != 0) return cudaErrorNotSupported;
    if (p.chunk_len % p.ndims != 0 && p.nchunks > 1) return cudaErrorNotSupported;
    if (p.chunk_len > 0x30000000ull) return cudaErrorNotSupported;      // 32-bit cursors inside one chunk
    const uint32_t NP = p.ndims / 8u, TPW = 32u / NP;
    const uint32_t H = (2u * p.ndims * (esz == 1 ? 3u : 4u) + 7u) / 8u;
    // 16-bit delta slots are staged in two halves (D = 80: 2165 -> 2456 GB/s, D = 256: 1986 -> 2865); the FIRE form would spill
    // the state it has to hold across the refill (1961 -> 1344 at D = 80)
    const bool half = esz == 2u && predictor == 0;
    const uint32_t slot = half ? 8u * p.ndims : 8u * p.ndims * esz; // 8 rows of 8-bit, 4 rows of 16-bit elements
    const uint32_t max_need = (H > slot ? H : slot) + 8u + 8u;      // one header or one (half) slot, plus the window over-read
    p.team = NP; p.dpl = TPW;
    p.ring_cap = pow2_ceil_w(max_need + NP * 32u + 32u);
    p.ring_mask4 = (p.ring_cap - 1u) & ~3u;
    const uint32_t tile = 64u * NP * esz;
    const uint32_t ntiles = (esz == 2u || predictor == 1) ? 1u : 2u;  // the per-element forms synchronise instead of alternating tiles
    p.team_smem = ((p.ring_cap + 32u + 128u + ntiles * (tile + 128u) + p.ring_cap - 1u) / p.ring_cap) * p.ring_cap;
    const uint32_t teams_per_cta = (kWideThreads / 32) * TPW;
    const size_t smem = (size_t)teams_per_cta * p.team_smem;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const uint64_t grid64 = (p.nchunks + teams_per_cta - 1) / teams_per_cta;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        count_launch();
        kernel<<<(uint32_t)grid64, kWideThreads, smem, stream>>>(p);
        return cudaGetLastError();
    };
    if (half) return go(spz_decode_wide_kernel<16, false, true>);
    if (esz == 2u) return predictor == 1 ? go(spz_decode_wide_kernel<16, true, false>) : go(spz_decode_wide_kernel<16, false, false>);
    return predictor == 1 ? go(spz_decode_wide_kernel<8, true, false>) : go(spz_decode_wide_kernel<8, false, false>);
}

}  // namespace spz

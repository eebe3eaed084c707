// spz_encode_wide.cu -- wide-row ("octet team") encode kernel: ndims 5..256 (whole octets in the aligned form, any ndims
// in the PART form); 8- and 16-bit elements, delta and FIRE.  The text below describes the 8-bit delta form (SIMD within a
// register); the FIRE and 16-bit forms share everything but the arithmetic of pass 1 and the field packing (see the
// comments at the kernel).
//
// BASELINE config 4 (int8 x 80 dims, delta) ran on the team kernel of spz_encode_hd.cu at 0.23 of the roofline: ncu
// showed 26 thread instructions per sample (lane-per-dim forecasting of single bytes, a cluster multiply-add bit
// emitter that re-reads every zigzagged byte and its multiplier from shared memory), LSU pipe 65 %, 22 % of the shared
// wavefronts bank conflicts.  The delta predictor needs no multiply, so everything up to the bit packing runs
// SIMD-within-a-register on FOUR samples per 32-bit word, and the per-dim packing constants live in registers:
//
//   mapping  one lane owns 8 consecutive dims (an OCTET = two words per row) of one chunk; a chunk is encoded by a team
//            of ndims/8 lanes (10 for 80 dims), 32/(ndims/8) teams per warp, warp-synchronous like the team kernels
//   input    a lane loads its 8 bytes of each of the block's 8 rows with 64-bit loads (a team's loads cover whole rows:
//            coalesced), one block ahead in registers
//   pass 1   packed subtract (delta), packed zigzag (one PRMT replicates the four sign bits), OR-mask: 9 instructions
//            per 4 samples; widths of 4 dims at once (7 -> 8 promotion, bit smear, per-byte popcount)
//   header   D % 8 == 0: an octet owns 3 whole header bytes per slot
//   pack     a quad of fields is squeezed to <= 32 contiguous bits with three multiply-adds (byte 1 and byte 3 move next
//            to bytes 0 and 2, then the high pair moves next to the low pair); two quads make the octet's <= 64-bit
//            piece of the row
//   rows     the pieces of a block go through shared memory ([row][octet], 64-bit); lane (team, row) then appends the
//            team's pieces of ITS row with a branch-free scalar bit writer straight into the output staging tile -- no
//            cross-lane bit merging, no atomics.  Whole-word stores only: a row's first word is stored with zero low
//            bytes, after a barrier the lane in front ORs its last partial word into it.  With one or two teams per warp
//            a row is cut into 4 / 2 segments of whole pieces handed over the same way, so all 32 lanes write rows.
//   control  the reference's run / group state machine (sprintz_delta_rle.cpp:134-312) as the per-team action generator
//            of spz_encode_hd.cu, executed in lock-step passes
//   output   shared staging tile, flushed with 128-bit stores
//
// Reference behaviour restated: sprintz_delta_rle.cpp:54-404 (high-dim delta + RLE encode), sprintz_xff_rle.cpp:61-555
// (high-dim FIRE + RLE encode).
#include "spz_host.h"
#include "spz_swar.cuh"

namespace spz {

constexpr int kWideEncThreads = 128;
constexpr unsigned kFullWE = 0xffffffffu;

enum : uint32_t { WACT_NONE = 0, WACT_RUN = 1, WACT_PAD = 2, WACT_DATA = 3 };

// per-byte zigzag (bitpack.h:302) of four int8 lanes: (d << 1) ^ (d >> 7)
__device__ __forceinline__ uint32_t zigzag4(uint32_t d)
{
    uint32_t sign;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(sign) : "r"(d));           // selector msb: replicate the byte's sign
    return ((d + d) & 0xfefefefeu) ^ sign;
}

// OR-masks of four dims -> their field widths, one per byte: bit length with 7 -> 8 (bitpack.h:72-93) => 0..6, 8
__device__ __forceinline__ uint32_t widths4(uint32_t m)
{
    m |= (m << 1) & 0x80808080u;                                        // bit 6 set => the width is 8
    uint32_t v = m | ((m >> 1) & 0x7f7f7f7fu);
    v |= (v >> 2) & 0x3f3f3f3fu;
    v |= (v >> 4) & 0x0f0f0f0fu;                                        // 2^n - 1 in every byte
    uint32_t n = v - ((v >> 1) & 0x55555555u);
    n = (n & 0x33333333u) + ((n >> 2) & 0x33333333u);
    return (n + (n >> 4)) & 0x0f0f0f0fu;
}

// four widths (one per byte) -> the 12 header bits of the quad; a width of 8 is stored as 7 (sprintz_delta_rle.cpp:199)
__device__ __forceinline__ uint32_t header12(uint32_t n4)
{
    const uint32_t f = n4 - ((n4 >> 3) & 0x01010101u);
    return (f & 7u) | ((f >> 5) & 0x38u) | ((f >> 10) & 0x1c0u) | ((f >> 15) & 0xe00u);
}

// four widths (one per byte) -> the 16 header bits of a 16-bit quad; a width of 16 is stored as 15
__device__ __forceinline__ uint32_t header16(uint32_t n4)
{
    const uint32_t f = n4 - ((n4 >> 4) & 0x01010101u);
    return (f & 0xfu) | ((f >> 4) & 0xf0u) | ((f >> 8) & 0xf00u) | ((f >> 12) & 0xf000u);
}

// multipliers that squeeze the four bytes of a zigzag word to contiguous fields
struct PackC { uint32_t c1, c3, c23; };
__device__ __forceinline__ void pack_consts(uint32_t n4, PackC& q)
{
    const uint32_t n0 = n4 & 0xffu, n1 = (n4 >> 8) & 0xffu, n2 = (n4 >> 16) & 0xffu;
    q.c1 = (1u << n0) - 256u;                    // byte 1 moves from bit 8 to bit n0
    q.c3 = ((1u << n2) - 256u) << 16;            // byte 3 moves from bit 24 to bit 16 + n2
    q.c23 = (1u << (n0 + n1)) - 65536u;          // the high pair moves from bit 16 to bit n0 + n1
}
__device__ __forceinline__ uint32_t pack_quad(uint32_t w, const PackC& q)
{
    const uint32_t w1 = w + __byte_perm(w, 0, 0x4441) * q.c1;
    const uint32_t w2 = w1 + (w >> 24) * q.c3;
    return w2 + (w2 >> 16) * q.c23;
}

// p.team = lanes per chunk (ndims / 8), p.dpl = teams per warp; team shared memory:
//   [piece buffers 2 x 64*NP][octet widths 2 x 32][row bit offsets of the octets 2 x 32 x u16][staging tile of out_cap bytes]
// LSEG: log2 of the segments a row is cut into for the row writers (0 with >= 3 teams per warp, 1 with two, 2 with one);
// RDX: the team total is one REDUX (three teams per warp) instead of a shuffle scan
// FIRE: the forecaster needs a multiply per sample, so pass 1 runs per element on left-aligned values exactly like the
// lane kernel's (spz_encode_lane.cu) and re-packs the zigzagged errors four to a word; everything behind pass 1 (widths,
// header, multiply-add packing, row writers, control) is shared with the delta form.
// PART: any ndims -- the last lane of a team owns a partial octet (its missing dims read as constant zeros: width 0, no
// bits), rows are not 8 / 16-byte aligned (two aligned loads and a byte funnel per row), and an octet's header fields are
// not whole bytes (they are OR-ed into the zeroed header with shared-memory atomics, two per lane and slot).
template <int W, bool FIRE, int MINB, int LSEG_, bool RDX, bool PART>
__global__ void __launch_bounds__(kWideEncThreads, MINB) spz_encode_wide_kernel(const EncParams p)
{
    static_assert(W == 8 || W == 16, "8- and 16-bit elements");
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr uint32_t ESZ = W / 8;
    constexpr uint32_t NQ = W / 8;                      // <= 64-bit pieces per octet and row: the octet (8-bit), its two quads (16-bit)
    constexpr int NXW = 2 * (W / 8);                    // words of an octet's row
    constexpr bool ELEM = FIRE || W == 16;              // pass 1 runs per element (left-aligned values)
    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t NP = p.team, TPW = p.dpl;
    const uint32_t tw_raw = (uint32_t)lane / NP;
    const bool lane_on = tw_raw < TPW;                  // lanes past the last whole team of the warp idle along
    const uint32_t tw = lane_on ? tw_raw : 0u;
    const uint32_t t = lane_on ? (uint32_t)lane - tw * NP : 0u;
    const uint64_t c64 = ((uint64_t)blockIdx.x * (kWideEncThreads / 32) + warp) * TPW + tw;
    const bool valid = lane_on && c64 < p.nchunks;
    const uint64_t c = valid ? c64 : 0;
    const uint32_t D = PART ? p.ndims : NP * 8u;
    const uint32_t dl = PART ? (D - 8u * t < 8u ? D - 8u * t : 8u) : 8u;     // dims of my octet
    const uint32_t last_lane = tw * NP + NP - 1u;
    // my team's lanes (an idle lane is a team of its own): a team-wide REDUX names exactly these
    const uint32_t tmask = lane_on ? ((NP == 32u ? 0xffffffffu : ((1u << NP) - 1u)) << (tw * NP)) : (1u << lane);

    // ---- shared memory of this team
    const uint32_t TS = p.team_smem;
    const uint32_t warp_base = (uint32_t)warp * TPW * TS;
    uint8_t* const sm = smem + warp_base + tw * TS;
    const uint32_t NPC = NP * NQ;                       // pieces of a row
    const uint32_t WCAP = PART ? ((NPC + 3u) & ~3u) : 32u * NQ;     // entries of a width / offset array (a constant in the aligned form)
    // row stride of the piece buffer, in pieces: a multiple of 16 pieces (128 bytes) would put the 8 row jobs of a team on
    // the same banks, so such rows are padded by one piece
    // (16-bit: by two, so that an octet's two pieces stay one aligned 128-bit store)
    const uint32_t NPS = NPC + ((NPC & 15u) ? 0u : NQ);
    const uint32_t PB = 64u * NPS;                      // one piece buffer: [8 rows][NPS] x 8 bytes
    // 8-bit: two piece buffers used in turn; 16-bit: one (emit_rows() ends with a barrier, and shared memory bounds the
    // occupancy of the 16-bit form: 80 dims 3 -> 4 CTAs per SM)
    constexpr uint32_t NBUF = W == 16 ? 1u : 2u;
    const uint32_t wid_off = NBUF * PB;                 // [NBUF][WCAP] piece widths (u8), then [NBUF][WCAP] bit offsets of the pieces in a row (u16)
    const uint32_t pre_off = wid_off + NBUF * WCAP;
    // rows are written by one lane each; with one or two teams per warp a row is cut into SEG segments of whole pieces
    // so that all 32 lanes work in that phase (ndims >= 88)
    constexpr uint32_t LSEG = LSEG_;                    // log2(SEG): every division by SEG is a shift
    constexpr uint32_t SEG = 1u << LSEG;
    uint8_t* const stage0 = sm + p.out_off;
    const uint32_t OCAP = p.out_cap;

    // ---- this chunk
    const uint64_t first = c * p.chunk_len;
    const int64_t L = valid ? (int64_t)((first + p.chunk_len <= p.len) ? p.chunk_len : (p.len - first)) : 0;
    const uint8_t* const gin = p.src + first * ESZ;     // octet rows are 8 / 16-byte aligned (checked on the host)
    uint8_t* const gstream = p.dst + c * p.slot_stride;
    const uint32_t B = (uint32_t)kBlockRows * D;
    const int64_t G = 2 * (int64_t)B;
    const uint32_t H = (2u * D * HB + 7u) >> 3;          // D % 8 == 0: an octet's fields are whole bytes
    const uint32_t SH = H >> 1;                          // header bytes of one slot
    const uint32_t meta_bytes = p.write_size ? 8u : 0u;
    const uint32_t max_group_bytes = H + 2u * B * ESZ + 4u;

    // ---- output staging: one tile; whole 16-byte units leave, the (< 16 byte) remainder moves to the front
    uint8_t* const ob = stage0;
    uint8_t* gdst;
    uint32_t ocur, oskip;
    {
        const uint32_t mis = (uint32_t)((uintptr_t)gstream & 15);
        gdst = gstream - mis; ocur = mis; oskip = mis;
    }
    auto flush = [&](bool on_, bool final_flush) {
        __syncwarp();
        uint32_t carry[4] = {0, 0, 0, 0};               // the < 16 remaining bytes move to the front as (up to) 4 words
        bool moved = false;
        if (on_) {
            const uint32_t hi = final_flush ? ocur : (ocur & ~15u);
            if (hi > oskip) {
                uint32_t a = (oskip + 15u) & ~15u;
                if (a > hi) a = hi;
                for (uint32_t b = oskip + t; b < a; b += NP) gdst[b] = ob[b];
                const uint4* s4 = reinterpret_cast<const uint4*>(ob);
                uint4* g4 = reinterpret_cast<uint4*>(gdst);
                for (uint32_t u = (a >> 4) + t; u < (hi >> 4); u += NP) g4[u] = s4[u];
                uint32_t tb = hi & ~15u;
                if (tb < a) tb = a;
                for (uint32_t b = tb + t; b < hi; b += NP) gdst[b] = ob[b];
                const uint32_t rem = ocur - hi;         // < 16 (0 on the final flush)
                {   // hi is a multiple of 16: whole words; the bytes of the last word past the cursor are don't-care
                    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(ob + hi);
#pragma unroll
                    for (uint32_t k = 0; k < 4u; k++)
                        if (4u * (t + k * NP) < rem) carry[k] = w32[t + k * NP];
                }
                gdst += hi; ocur = rem; oskip = 0;
                moved = rem != 0u;
            }
        }
        __syncwarp();                                   // the tile has been read
        if (moved) {
            uint32_t* w32 = reinterpret_cast<uint32_t*>(ob);
#pragma unroll
            for (uint32_t k = 0; k < 4u; k++)
                if (4u * (t + k * NP) < ocur) w32[t + k * NP] = carry[k];
        }
        __syncwarp();                                   // the carried words hold don't-care bytes past the cursor: later writers come after
    };
    if (lane_on && t == 0) for (uint32_t b = 0; b < meta_bytes; b++) ob[ocur + b] = 0;     // metadata placeholder
    ocur += meta_bytes;

    // ---- team-uniform encoder state (sprintz_delta_rle.cpp:132-158)
    uint32_t xA = 0, xB = 0;                            // the previous row's values of my 8 dims
    uint32_t xprev[ELEM ? 8 : 1];                       // per-element form: previous values / deltas left-aligned, FIRE counters (sprintz_xff_rle.cpp:148-152)
    int32_t dprev[FIRE ? 8 : 1], ct[FIRE ? 8 : 1];
#pragma unroll
    for (int e = 0; e < (ELEM ? 8 : 1); e++) xprev[e] = 0;
#pragma unroll
    for (int e = 0; e < (FIRE ? 8 : 1); e++) { dprev[e] = 0; ct[e] = 0; }
    uint32_t ngroups = 0, nslots = 0, run = 0, blk = 0;
    uint32_t hdr_off = 0;
    int64_t tail = 0;
    bool outer = true;
    bool on = valid && L >= (int64_t)kMinDataSize;
    // blk*B <= L-G  <=>  blk <= last_start ; blk*B < L-G  <=>  blk <= last_cont   (both -1 when L < G)
    const int32_t last_start = (L >= G) ? (int32_t)((L - G) / B) : -1;
    // delta: strict `<` (sprintz_delta_rle.cpp:226); FIRE: `<=` (sprintz_xff_rle.cpp:362)
    const int32_t last_cont = FIRE ? last_start : ((L > G) ? (int32_t)((L - G - 1) / B) : -1);

    // ---- input prefetch (registers), one block ahead
    uint32_t nx[kBlockRows][NXW];
    // PART: which bytes of my octet's row exist
    uint32_t vmask[NXW];
#pragma unroll
    for (int k = 0; k < NXW; k++) {
        const uint32_t per = 4u / ESZ;                  // dims per word
        const uint32_t have = dl > (uint32_t)k * per ? dl - (uint32_t)k * per : 0u;
        vmask[k] = have >= per ? 0xffffffffu : ((1u << (8u * ESZ * have)) - 1u);
    }
    auto prefetch = [&](uint32_t b) {
        const bool ok = on && (int64_t)(b + 1u) * B <= L;
        const uint8_t* src = gin + ((size_t)b * B + 8u * t) * ESZ;
#pragma unroll
        for (int r = 0; r < kBlockRows; r++) {
            if (PART) {
                // the 8 * ESZ bytes at an arbitrary address: the two aligned vectors around it, shifted by whole words
                // (selects) and by bytes (PRMT); the second vector is only read when my dims reach into it
                const uint8_t* a = src + (uint32_t)r * D * ESZ;
                const uint32_t o = (uint32_t)((uintptr_t)a & (8u * ESZ - 1u));
                const uint8_t* base = a - o;
                const bool two = ok && o + dl * ESZ > 8u * ESZ;
                uint32_t w[2 * NXW + 1];
                if (W == 8) {
                    uint2 q0 = make_uint2(0u, 0u), q1 = make_uint2(0u, 0u);
                    if (ok) q0 = __ldg(reinterpret_cast<const uint2*>(base));
                    if (two) q1 = __ldg(reinterpret_cast<const uint2*>(base + 8));
                    w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; w[4] = 0u;
                } else {
                    uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = make_uint4(0u, 0u, 0u, 0u);
                    if (ok) q0 = __ldg(reinterpret_cast<const uint4*>(base));
                    if (two) q1 = __ldg(reinterpret_cast<const uint4*>(base + 16));
                    constexpr int S = W == 16 ? 1 : 0;  // (this branch is dead code for 8-bit elements: keep its indices in range)
                    w[0] = q0.x; w[1] = q0.y; w[2 * S] = q0.z; w[3 * S] = q0.w;
                    w[4 * S] = q1.x; w[5 * S] = q1.y; w[6 * S] = q1.z; w[7 * S] = q1.w; w[8 * S] = 0u;
                }
                const uint32_t kw = o >> 2;             // whole words to skip: 0..NXW-1
                if (W == 16 && (kw & 2u)) {
#pragma unroll
                    for (int i = 0; i + 2 <= 2 * NXW; i++) w[i] = w[i + 2];
                }
                if (kw & 1u) {
#pragma unroll
                    for (int i = 0; i + 1 <= 2 * NXW; i++) w[i] = w[i + 1];
                }
                const uint32_t sel = 0x3210u + 0x1111u * (o & 3u);
#pragma unroll
                for (int k = 0; k < NXW; k++) nx[r][k] = __byte_perm(w[k], w[k + 1], sel) & vmask[k];
            } else if (W == 8) {
                uint2 v = make_uint2(0u, 0u);
                if (ok) v = __ldg(reinterpret_cast<const uint2*>(src + (uint32_t)r * D));
                nx[r][0] = v.x; nx[r][1] = v.y;
            } else {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (ok) v = __ldg(reinterpret_cast<const uint4*>(src + (uint32_t)r * D * ESZ));
                nx[r][0] = v.x; nx[r][1] = v.y; nx[r][NXW - 2] = v.z; nx[r][NXW - 1] = v.w;
            }
        }
    };

    // this block's zigzagged errors (four 8-bit / two 16-bit to a word) and layout; nA / nB: the widths of my dims 0..3 /
    // 4..7, one per byte
    uint32_t zz[kBlockRows][NXW];
    uint32_t nA = 0, nB = 0, wA = 0, wo = 0, row_bytes = 0, excl = 0;
    uint32_t info = 0;                                  // what the row lanes need from my team (lane 0 of the team is read)

    // One slot of the open group (team-predicated by the caller).  Data slots leave their pieces in the piece buffer;
    // emit_rows() turns them into payload bytes.
    auto emit_slot = [&](uint32_t act, uint32_t run_len, uint32_t pbsel) -> bool {
        const uint32_t sl = nslots & 1u;
        if (sl == 0u) {
            hdr_off = ocur; ocur += H; ngroups++;
            if (PART) {                                 // the fields are OR-ed in
                for (uint32_t b = t; b < H; b += NP) ob[hdr_off + b] = 0;
                __syncwarp(tmask);
            }
        }
        uint32_t h24 = 0;
        if (act == WACT_RUN) {       // zero header fields + 1..2 length bytes (sprintz_delta_rle.cpp:238-245)
            const uint32_t two = run_len > 0x7fu ? 1u : 0u;
            if (t == 0) {
                ob[ocur] = (uint8_t)((run_len & 0x7fu) | (two << 7));
                if (two) ob[ocur + 1] = (uint8_t)(run_len >> 7);
            }
            ocur += 1u + two;
        } else if (act == WACT_PAD) {
            if (t == 0) ob[ocur] = 0;
            ocur += 1u;
        } else {
            // SOLO (a one-lane team, ndims <= 8 in the PART form): the lane holds every piece of its rows, so it appends the 8
            // byte-padded rows itself with a scalar bit writer -- no piece buffer, no row jobs (int8 x 5 dims: 463 -> 606 GB/s, int16 x 5: 508 -> 809)
            const bool solo = PART && NP == 1u;
            uint32_t s_waddr = 0, s_fill = 0, s_a0 = 0;
            if (solo) {
                const uint32_t bit0 = (uint32_t)((ob + ocur) - smem) * 8u;
                s_waddr = (bit0 >> 5) << 2; s_fill = bit0 & 31u;
                if (s_fill) s_a0 = *reinterpret_cast<const uint32_t*>(smem + s_waddr) & ((1u << s_fill) - 1u);
            }
            auto append = [&](uint32_t x, uint32_t y, uint32_t w) {      // w <= 64 bits of (x, y), clean above w
                s_a0 |= x << s_fill;
                const uint32_t a1 = __funnelshift_l(x, y, s_fill), a2 = __funnelshift_l(y, 0u, s_fill);
                s_fill += w;
                const bool p1 = s_fill >= 32u, p2 = s_fill >= 64u;
                if (p1) *reinterpret_cast<uint32_t*>(smem + s_waddr) = s_a0;
                if (p2) *reinterpret_cast<uint32_t*>(smem + s_waddr + 4u) = a1;
                s_a0 = p2 ? a2 : (p1 ? a1 : s_a0);
                s_waddr += (s_fill >> 5) * 4u;
                s_fill &= 31u;
            };
            auto pad_bits = [&](uint32_t n) {                            // n < 8 zero bits: rows are byte-padded
                s_fill += n;
                if (s_fill >= 32u) { *reinterpret_cast<uint32_t*>(smem + s_waddr) = s_a0; s_a0 = 0; s_waddr += 4u; s_fill -= 32u; }
            };
            const uint32_t row_pad = row_bytes * 8u - wo;
            if (W == 8) {
                h24 = header12(nA) | (header12(nB) << 12);
                PackC qa, qb;
                pack_consts(nA, qa);
                pack_consts(nB, qb);
                uint2* pb = reinterpret_cast<uint2*>(sm + pbsel * PB) + t;
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) {
                    const uint32_t va = pack_quad(zz[r][0], qa), vb = pack_quad(zz[r][1], qb);
                    // octet piece = va | vb << wA (wA <= 32: the clamping funnel shifts give 0 / vb at 32)
                    const uint2 piece = make_uint2(va | __funnelshift_lc(0u, vb, wA), __funnelshift_lc(vb, 0u, wA));
                    if (solo) { append(piece.x, piece.y, wo); pad_bits(row_pad); }
                    else pb[(uint32_t)r * NPS] = piece;
                }
                if (!solo) {
                    sm[wid_off + pbsel * WCAP + t] = (uint8_t)wo;
                    if (SEG > 1u) reinterpret_cast<uint16_t*>(sm + pre_off)[pbsel * WCAP + t] = (uint16_t)excl;
                }
            } else {
                h24 = header16(nA) | (header16(nB) << 16);
                // a pair of fields p = lo | hi << 16 closes up as p + (p >> 16) * (2^n_lo - 2^16); two pairs make a quad's
                // <= 64-bit piece
                const uint32_t mA0 = (1u << (nA & 0xffu)) - 65536u, mA1 = (1u << ((nA >> 16) & 0xffu)) - 65536u;
                const uint32_t mB0 = (1u << (nB & 0xffu)) - 65536u, mB1 = (1u << ((nB >> 16) & 0xffu)) - 65536u;
                const uint32_t cA0 = (nA & 0xffu) + ((nA >> 8) & 0xffu), cB0 = (nB & 0xffu) + ((nB >> 8) & 0xffu);
                uint2* pb = reinterpret_cast<uint2*>(sm + pbsel * PB) + 2u * t;
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) {
                    const uint32_t a0 = zz[r][0] + (zz[r][0] >> 16) * mA0, a1 = zz[r][1] + (zz[r][1] >> 16) * mA1;
                    const uint32_t b0 = zz[r][NXW - 2] + (zz[r][NXW - 2] >> 16) * mB0, b1 = zz[r][NXW - 1] + (zz[r][NXW - 1] >> 16) * mB1;
                    const uint2 pa = make_uint2(a0 | __funnelshift_lc(0u, a1, cA0), __funnelshift_lc(a1, 0u, cA0));
                    const uint2 pq = make_uint2(b0 | __funnelshift_lc(0u, b1, cB0), __funnelshift_lc(b1, 0u, cB0));
                    if (solo) { append(pa.x, pa.y, wA); append(pq.x, pq.y, wo - wA); pad_bits(row_pad); }
                    else *reinterpret_cast<uint4*>(pb + (uint32_t)r * NPS) = make_uint4(pa.x, pa.y, pq.x, pq.y);
                }
                if (!solo) {
                    sm[wid_off + pbsel * WCAP + 2u * t] = (uint8_t)wA;
                    sm[wid_off + pbsel * WCAP + 2u * t + 1u] = (uint8_t)(wo - wA);
                }
                if (!solo && SEG > 1u) {
                    uint16_t* pre = reinterpret_cast<uint16_t*>(sm + pre_off) + pbsel * WCAP + 2u * t;
                    pre[0] = (uint16_t)excl; pre[1] = (uint16_t)(excl + wA);
                }
            }
            if (solo) { if (s_fill) *reinterpret_cast<uint32_t*>(smem + s_waddr) = s_a0; }     // zeros behind the payload: written later
            else info = 1u | (row_bytes << 1) | ((uint32_t)((ob + ocur) - smem) << 11);
            ocur += kBlockRows * row_bytes;
            blk++;
        }
        if (PART) {     // my octet's fields start at bit (sl * D + 8 * t) * HB of the header: one or two words
            if (h24) {
                const uint32_t hb = ((uint32_t)((ob + hdr_off) - smem)) * 8u + (sl * D + 8u * t) * (uint32_t)HB;
                uint32_t* hw = reinterpret_cast<uint32_t*>(smem) + (hb >> 5);
                const uint32_t sh = hb & 31u;
                atomicOr(hw, h24 << sh);
                const uint32_t hi = __funnelshift_l(h24, 0u, sh);       // h24 >> (32 - sh), 0 when sh == 0
                if (hi) atomicOr(hw + 1, hi);
            }
        } else {        // my octet's 3 / 4 header bytes of this slot
            uint8_t* hp = ob + hdr_off + sl * SH + (uint32_t)HB * t;
            hp[0] = (uint8_t)h24; hp[1] = (uint8_t)(h24 >> 8); hp[2] = (uint8_t)(h24 >> 16);
            if (W == 16) hp[HB - 1] = (uint8_t)(h24 >> 24);
        }
        nslots++;
        return (nslots & 1u) == 0u;
    };

    // Payload rows of the data slots just emitted: lane (team, row) appends the team's NP pieces of its row.
    // Warp-uniform call; `doit` = my team emitted a data slot in this pass.
    //
    // Rows are byte aligned, the writer stores whole words: a row's first word is stored with ZERO low bytes (they
    // belong to the previous row), so after a barrier the previous row's lane ORs its last partial word into it.  Row 0
    // saves the bytes in front of the payload before and restores them after; row 7 stores its tail with zeros behind
    // the payload (bytes that are only written later).  Rows shorter than 4 bytes (which may not complete a word) take
    // a byte-wise path.
    auto emit_rows = [&](bool doit, uint32_t pbsel) {
        if (!doit) info = 0;
        __syncwarp();                                   // pieces, widths, offsets and header bytes are in shared memory
        const uint32_t njobs = 8u * TPW * SEG;          // job = (team, row, segment): consecutive lanes are consecutive in the stream
        for (uint32_t j0 = 0; j0 < njobs; j0 += 32u) {
            const uint32_t j = j0 + (uint32_t)lane;
            const uint32_t jt_raw = j >> (3u + LSEG);
            const uint32_t jt = jt_raw < TPW ? jt_raw : TPW - 1u;
            const uint32_t inf = __shfl_sync(kFullWE, info, (int)(jt * NP));
            const bool act = j < njobs && (inf & 1u);
            const uint32_t r = (j >> LSEG) & 7u;
            uint32_t seg = j & (SEG - 1u);
            const uint32_t rb = (inf >> 1) & 0x3ffu;
            const uint32_t rowaddr = (inf >> 11) + r * rb;              // byte offset from smem[]
            const uint32_t end = rowaddr + rb;
            const uint8_t* tb = smem + warp_base + jt * TS;
            const uint8_t* wd = tb + wid_off + pbsel * WCAP;
            const uint16_t* pre = reinterpret_cast<const uint16_t*>(tb + pre_off) + pbsel * WCAP;
            // my pieces [u0, u1) and their bit range [s0, s1) inside the row; a segmented row needs >= 64 bits per segment
            // (every job must complete the word it starts in), else lane `seg == 0` writes the whole row
            uint32_t u0 = 0, u1 = NPC, s0 = 0, s1 = rb * 8u;
            bool last = true, mine = true;
            if (SEG > 1u) {
                bool segok = act;
                uint32_t lo = 0;
                for (uint32_t k = 1; k <= SEG && segok; k++) {
                    const uint32_t hi = k < SEG ? (uint32_t)pre[(k * NPC) >> LSEG] : rb * 8u;
                    segok = hi - lo >= 64u;
                    lo = hi;
                }
                if (segok) {
                    u0 = (seg * NPC) >> LSEG; u1 = ((seg + 1u) * NPC) >> LSEG;
                    s0 = pre[u0];
                    last = seg == SEG - 1u;
                    if (!last) s1 = pre[u1];
                } else {
                    mine = seg == 0u;
                    seg = 0u;
                }
            }
            // a row of >= 4 bytes completes the word it starts in, which is all the hand-over below needs (every word then has
            // exactly one job that stores it whole; a shorter row could share one word with BOTH neighbours)
            const bool fast = act && mine && rb >= 4u;
            const uint32_t bit0 = rowaddr * 8u + s0;
            uint32_t waddr = (bit0 >> 5) << 2;
            const uint32_t fill0 = bit0 & 31u;
            uint32_t fill = fill0;
            uint32_t a0 = 0, old0 = 0;
            const bool first_job = r == 0u && seg == 0u;   // the bytes in front of the payload are saved and restored
            const uint2* pc = reinterpret_cast<const uint2*>(tb + pbsel * PB) + r * NPS;
            if (fast) {
                if (first_job && fill0) old0 = *reinterpret_cast<const uint32_t*>(smem + waddr) & ((1u << fill0) - 1u);
                for (uint32_t u = u0; u < u1; u++) {
                    const uint2 pv = pc[u];
                    const uint32_t w = wd[u];
                    a0 |= pv.x << fill;
                    const uint32_t a1 = __funnelshift_l(pv.x, pv.y, fill);
                    const uint32_t a2 = __funnelshift_l(pv.y, 0u, fill);
                    fill += w;
                    // 0, 1 or 2 whole words are complete: predicated stores and selects, no branches
                    const bool p1 = fill >= 32u, p2 = fill >= 64u;
                    if (p1) *reinterpret_cast<uint32_t*>(smem + waddr) = a0;
                    if (p2) *reinterpret_cast<uint32_t*>(smem + waddr + 4u) = a1;
                    a0 = p2 ? a2 : (p1 ? a1 : a0);
                    waddr += (fill >> 5) * 4u;
                    fill &= 31u;
                }
            } else if (act && mine) {
                auto put = [&](uint32_t wa, uint32_t v) {
#pragma unroll
                    for (uint32_t i = 0; i < 4u; i++)
                        if (wa + i >= rowaddr && wa + i < end) smem[wa + i] = (uint8_t)(v >> (8u * i));
                };
                for (uint32_t u = 0; u < NPC; u++) {
                    const uint2 pv = pc[u];
                    const uint32_t w = wd[u];
                    a0 |= pv.x << fill;
                    uint32_t a1 = __funnelshift_l(pv.x, pv.y, fill);
                    const uint32_t a2 = __funnelshift_l(pv.y, 0u, fill);
                    fill += w;
                    if (fill >= 32u) {
                        put(waddr, a0); waddr += 4u; a0 = a1; a1 = a2; fill -= 32u;
                        if (fill >= 32u) { put(waddr, a0); waddr += 4u; a0 = a1; fill -= 32u; }
                    }
                }
                if (waddr < end) put(waddr, a0);
            }
            __syncwarp();                               // every whole word of every job is stored
            if (fast) {
                // my last, partial word: the job behind me (next segment / next row) stored it with ZERO low bits when
                // my range ends inside it.  The last job of the slot stores zeros behind the payload.
                const bool tail = last ? (waddr < end) : (fill != 0u);
                if (tail) {
                    const bool shared = last ? (r < 7u && (end & 3u)) : true;
                    if (shared) a0 |= *reinterpret_cast<const uint32_t*>(smem + waddr);
                    *reinterpret_cast<uint32_t*>(smem + waddr) = a0;
                }
                if (first_job && fill0) *reinterpret_cast<uint32_t*>(smem + ((rowaddr * 8u) >> 5 << 2)) |= old0;
            }
        }
        __syncwarp();                                   // rows are complete before anything is appended behind them
    };

    // Group finished: the staging tile leaves for HBM when the next group might not fit.  Warp-uniform call.
    auto close_group = [&](bool gc) {
        const bool need = gc && ocur + max_group_bytes + 16u > OCAP;
        if (__any_sync(kFullWE, need)) flush(need, false);
    };

    prefetch(0);
    for (uint32_t it = 0;; it++) {
        if (on && outer && (int32_t)blk > last_start) { tail = (int64_t)blk * B; on = false; }
        if (!__any_sync(kFullWE, on)) break;
        const uint32_t pbsel = NBUF == 2u ? (it & 1u) : 0u;

        // ================= pass 1: packed delta -> zigzag -> OR-mask of this block (sprintz_delta_rle.cpp:161-201)
        if (ELEM) {
            // per element on left-aligned values.  FIRE (sprintz_xff_rle.cpp:205-275): prediction = (prev_delta * coef) >> W,
            // error = delta - prediction; the counter moves by the block's sum of icopysign(err, prev_delta) over the odd rows
            int32_t coef[FIRE ? 8 : 1], gsh[FIRE ? 8 : 1];
#pragma unroll
            for (int e = 0; e < (FIRE ? 8 : 1); e++) { coef[e] = fire_coef<W, false>(ct[e]); gsh[e] = 0; }
            auto elem = [&](uint32_t vh, int e, int r) -> uint32_t {   // left-aligned value in, left-aligned zigzagged error out
                const int32_t dh = (int32_t)(vh - xprev[e]);
                xprev[e] = vh;
                int32_t eh = dh;
                if (FIRE) {
                    const int f = FIRE ? e : 0;
                    const int32_t pred = __mulhi(dprev[f], coef[f]);
                    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(eh) : "r"(pred), "n"(-(1 << UP)), "r"(dh));
                    const int32_t sg = eh >> 31;
                    if (r & 1) {
                        int32_t cs;
                        asm("mad.lo.s32 %0, %1, 2, 1;" : "=r"(cs) : "r"(sg));
                        if (eh != 0) gsh[f] += dprev[f] * cs;
                    }
                    dprev[f] = dh;
                }
                const int32_t sg = eh >> 31;
                int32_t e2;
                asm("mad.lo.s32 %0, %1, 2, 0;" : "=r"(e2) : "r"(eh));
                return (uint32_t)e2 ^ (uint32_t)sg;
            };
            uint32_t mk[NXW];
#pragma unroll
            for (int k = 0; k < NXW; k++) mk[k] = 0;
#pragma unroll
            for (int r = 0; r < kBlockRows; r++) {
#pragma unroll
                for (int k = 0; k < NXW; k++) {
                    const uint32_t x = nx[r][k];
                    uint32_t z;
                    if (W == 8) {
                        const uint32_t z0 = elem(x << 24, 4 * k + 0, r);
                        const uint32_t z1 = elem(__byte_perm(x, 0, 0x1444), 4 * k + 1, r);
                        const uint32_t z2 = elem(__byte_perm(x, 0, 0x2444), 4 * k + 2, r);
                        const uint32_t z3 = elem(x & 0xff000000u, 4 * k + 3, r);
                        const uint32_t a = __byte_perm(z0, z1, 0x0073), b = __byte_perm(z2, z3, 0x0073);
                        z = __byte_perm(a, b, 0x5410);
                    } else {
                        const uint32_t z0 = elem(x << 16, (2 * k + 0) & 7, r);
                        const uint32_t z1 = elem(x & 0xffff0000u, (2 * k + 1) & 7, r);
                        z = __byte_perm(z0, z1, 0x7632);
                    }
                    zz[r][k] = z; mk[k] |= z;
                }
            }
            if (FIRE) {
#pragma unroll
                for (int e = 0; e < (FIRE ? 8 : 1); e++) ct[e] = sx<CB>(ct[e] + ((gsh[e] >> UP) >> 2));   // sprintz_xff_rle.cpp:273-275
            }
            if (W == 8) {
                nA = widths4(mk[0]); nB = widths4(mk[1]);
            } else {    // bit length with 7 -> 8 and 15 -> 16 (sprintz_delta_rle.cpp:177-187), one width per byte
                nA = nbits_rule<16, false>(mk[0] & 0xffffu) | (nbits_rule<16, false>(mk[0] >> 16) << 8) |
                     (nbits_rule<16, false>(mk[1] & 0xffffu) << 16) | (nbits_rule<16, false>(mk[1] >> 16) << 24);
                nB = nbits_rule<16, false>(mk[NXW - 2] & 0xffffu) | (nbits_rule<16, false>(mk[NXW - 2] >> 16) << 8) |
                     (nbits_rule<16, false>(mk[NXW - 1] & 0xffffu) << 16) | (nbits_rule<16, false>(mk[NXW - 1] >> 16) << 24);
            }
        } else {
            uint32_t a = xA, b = xB, ma = 0, mb = 0;
#pragma unroll
            for (int r = 0; r < kBlockRows; r++) {
                const uint32_t va = nx[r][0], vb = nx[r][1];
                zz[r][0] = zigzag4(psub<8>(va, a)); ma |= zz[r][0]; a = va;
                zz[r][1] = zigzag4(psub<8>(vb, b)); mb |= zz[r][1]; b = vb;
            }
            if (on) { xA = a; xB = b; }
            nA = widths4(ma); nB = widths4(mb);
        }
        // 8-bit: the next block's loads are issued now and land during the rest of the iteration; 16-bit rows are twice the
        // registers, so they are issued once the zigzagged block has gone to shared memory
        if (W == 8) prefetch(it + 1u);
        wA = (nA * 0x01010101u) >> 24;
        wo = wA + ((nB * 0x01010101u) >> 24);
        // team total of the widths (and, for segmented rows, the bit offset of my octet inside a row).  Three teams per warp:
        // one REDUX per team with the team's own mask (the hardware reduces one mask at a time, so this only pays with few
        // teams and a full machine: c4 2.54 -> 2.39 ms; with 10 teams per warp it costs 20 %).  Otherwise a shuffle scan.
        uint32_t total;
        excl = 0;
        if (RDX) {
            total = __reduce_add_sync(tmask, wo);
        } else {
            uint32_t incl = wo;
            for (uint32_t sft = 1; sft < NP; sft <<= 1) {
                const uint32_t y = __shfl_up_sync(kFullWE, incl, sft);
                if (t >= sft) incl += y;
            }
            excl = incl - wo;
            total = __shfl_sync(kFullWE, incl, (int)last_lane);
        }
        const bool zero = (total == 0);
        row_bytes = (total + 7u) >> 3;

        // ================= control
        // Fast path: every live team has a plain data block (no run open, block not all-zero).
        if (__all_sync(kFullWE, !on || (!zero && run == 0u))) {
            bool gc = false;
            if (on) { gc = emit_slot(WACT_DATA, 0u, pbsel); outer = gc; }
            if (W == 16) prefetch(it + 1u);
            if (!(PART && NP == 1u)) emit_rows(on, pbsel);
            if (__any_sync(kFullWE, gc)) close_group(gc);
            continue;
        }
        // General case: per-team action generator (run slot / pad / data slot), executed in lock-step passes.
        // st: 0 = block just read; 1 = run flushed because the data ran out, maybe pad; 2 = finished; 3 = block consumed
        uint32_t st = on ? 0u : 3u;
        for (;;) {
            uint32_t act = WACT_NONE, run_len = 0;
            if (st == 0u) {
                if (zero && run < kMaxRun) {
                    run++; blk++;
                    if ((int32_t)blk <= last_cont) { outer = false; st = 3u; }
                    else { act = WACT_RUN; run_len = run; run = 0; st = 1u; }    // not enough data to finish the group
                } else if (run > 0) {
                    act = WACT_RUN; run_len = run; run = 0;                      // a run just ended; re-evaluate this block
                } else {
                    act = WACT_DATA; st = 3u;
                }
            } else if (st == 1u) {
                if (nslots & 1u) { act = WACT_PAD; }
                st = 2u;
            }
            if (st == 2u && act == WACT_NONE) { tail = (int64_t)blk * B; on = false; st = 3u; }
            if (!__any_sync(kFullWE, act != WACT_NONE)) break;

            bool gc = false;
            if (act != WACT_NONE) {
                gc = emit_slot(act, run_len, pbsel);
                if (act == WACT_DATA) outer = gc;
            }
            if (!(PART && NP == 1u) && __any_sync(kFullWE, act == WACT_DATA)) emit_rows(act == WACT_DATA, pbsel);
            if (__any_sync(kFullWE, gc)) close_group(gc);
        }
        if (W == 16) prefetch(it + 1u);
    }

    // ---- trailing raw elements (sprintz_delta_rle.cpp:398-403); also the whole input when L < 128
    {
        const int64_t rem_elems = valid ? (L - tail) : 0;
        const uint8_t* tsrc = gin + tail * ESZ;
        uint32_t left = (uint32_t)rem_elems * ESZ, done = 0;
        while (__any_sync(kFullWE, left > 0)) {
            const uint32_t space = OCAP - ocur;
            const uint32_t take = left < space ? left : space;
            for (uint32_t b = t; b < take; b += NP) ob[ocur + b] = tsrc[done + b];
            ocur += take; done += take; left -= take;
            flush(left > 0, false);
        }
        const uint32_t sbytes = (uint32_t)((gdst + ocur) - gstream);      // exact stream length
        flush(valid, true);
        __syncwarp();
        if (valid && t == 0) {
            if (p.write_size) {     // format.h:35-45, written last because ngroups is only known now
                const uint32_t rem16 = (uint32_t)rem_elems & 0xffffu;
                gstream[0] = (uint8_t)ngroups; gstream[1] = (uint8_t)(ngroups >> 8);
                gstream[2] = (uint8_t)(ngroups >> 16); gstream[3] = (uint8_t)(ngroups >> 24);
                gstream[4] = (uint8_t)rem16; gstream[5] = (uint8_t)(rem16 >> 8);
                gstream[6] = (uint8_t)D; gstream[7] = (uint8_t)(D >> 8);
            }
            ChunkMeta m;
            m.nbytes = sbytes;
            m.ngroups = ngroups;
            m.remaining = (uint32_t)rem_elems;
            m.status = kStatusOk;
            p.meta[c] = m;
        }
    }
}

static inline uint32_t up16w(uint32_t x) { return (x + 15u) & ~15u; }

// Returns cudaErrorNotSupported when the shape is not covered (the caller continues with the team kernels).
cudaError_t launch_encode_wide(int elem_sz, int predictor, EncParams p, cudaStream_t stream)
{
    if (get_force_path() == 1) return cudaErrorNotSupported;
    if ((elem_sz != 1 && elem_sz != 2) || (predictor != 0 && predictor != 1) || p.variant != 0) return cudaErrorNotSupported;
    // high-dim layouts only (sprintz.cpp:44-71: up to 4 / 2 dims use the low-dim codecs)
    // (int16 x 3 / 4 dims are high-dim too, but the team kernel measured faster: 663 / 782 vs 603 / 649 GB/s)
    if (p.ndims < 5u || p.ndims > 256) return cudaErrorNotSupported;
    if (p.ndims == 8u) return cudaErrorNotSupported;                    // lane kernels, else the team kernel
    if (p.chunk_len > 0x30000000ull) return cudaErrorNotSupported;      // 32-bit cursors inside one chunk
    // whole octets and 64 / 128-bit aligned octet rows in every chunk: the aligned form; anything else: the PART form
    const uint32_t esz = (uint32_t)elem_sz, amask = 8u * esz - 1u;
    const bool part = (p.ndims & 7u) || ((uintptr_t)p.src & amask) || (p.nchunks > 1 && ((p.chunk_len * esz) & amask));
    const uint32_t NP = (p.ndims + 7u) / 8u, TPW = 32u / NP, NPC = NP * esz;
    const uint32_t H = (2u * p.ndims * (esz == 1 ? 3u : 4u) + 7u) / 8u;
    p.team = NP; p.dpl = TPW;
    const uint32_t grp = H + 16u * p.ndims * esz + 4u;
    const uint32_t teams_per_cta = (kWideEncThreads / 32) * TPW;
    p.teams_per_cta = teams_per_cta;
    const uint32_t nbuf = esz == 2u ? 1u : 2u;
    const uint32_t NPS = NPC + ((NPC & 15u) ? 0u : esz);           // padded row stride of the piece buffer (see the kernel)
    const uint32_t wcap = part ? ((NPC + 3u) & ~3u) : 32u * esz;
    p.out_off = up16w(nbuf * (64u * NPS + wcap + 2u * wcap));       // piece buffers | piece widths | row bit offsets
    // the staging tile holds up to 4 groups, fewer when that would cost occupancy (6 CTAs per SM need <= 36 KB each; the
    // 16-bit and FIRE forms run 4 CTAs per SM: <= 56 KB)
    const uint32_t budget = (esz == 2 || predictor == 1) ? 56u * 1024u : 36u * 1024u;
    for (uint32_t ngrp = 4; ngrp >= 1; ngrp--) {
        p.out_cap = up16w(32u + ngrp * grp + 16u);
        p.team_smem = p.out_off + p.out_cap;
        // a team stride that is a multiple of 128 bytes puts the piece stores of ALL the warp's teams on the same banks (16
        // dims: stride 768, 74 % of the shared wavefronts were conflicts, 754 -> 1018 GB/s with the stagger); team k then starts
        // where team k-1's pieces of a row end.  Other strides are left alone: the row jobs' loads prefer different offsets
        // than the piece stores, and what the layout happens to give measured as well or better (int16 x 64: 1325 vs 1023)
        if (TPW > 1u && (p.team_smem & 127u) == 0u) p.team_smem += (8u * esz * NP + 15u) & ~15u & 127u;
        if ((size_t)teams_per_cta * p.team_smem <= budget) break;
    }
    const size_t smem = (size_t)teams_per_cta * p.team_smem;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const uint64_t grid64 = (p.nchunks + teams_per_cta - 1) / teams_per_cta;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    // 96 registers / 5 CTAs per SM: measured equal to 127 registers / 4 CTAs (the kernel is issue-bound, not latency-bound);
    // 80 registers / 6 CTAs spills and is 8 % slower
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (g_enc_probe) return probe_fill(g_enc_probe, 3, kernel, kWideEncThreads, smem, teams_per_cta);
        count_launch();
        kernel<<<(uint32_t)grid64, kWideEncThreads, smem, stream>>>(p);
        return cudaGetLastError();
    };
#define SPZ_WIDE_GO(Wv, Fv, Mv) \
    if (part) { \
        if (TPW == 1u) return go(spz_encode_wide_kernel<Wv, Fv, 4, 2, false, true>); \
        if (TPW == 2u) return go(spz_encode_wide_kernel<Wv, Fv, 4, 1, false, true>); \
        return go(spz_encode_wide_kernel<Wv, Fv, 4, 0, false, true>); \
    } \
    if (TPW == 1u) return go(spz_encode_wide_kernel<Wv, Fv, Mv, 2, false, false>); \
    if (TPW == 2u) return go(spz_encode_wide_kernel<Wv, Fv, Mv, 1, false, false>); \
    if (TPW == 3u) return go(spz_encode_wide_kernel<Wv, Fv, Mv, 0, true, false>); \
    return go(spz_encode_wide_kernel<Wv, Fv, Mv, 0, false, false>);
    // 16-bit rows are twice the registers, FIRE keeps 24 more registers of state per lane: 4 CTAs per SM
    if (esz == 2u) {
        if (predictor == 1) { SPZ_WIDE_GO(16, true, 4) }
        SPZ_WIDE_GO(16, false, 4)
    }
    if (predictor == 1) { SPZ_WIDE_GO(8, true, 4) }
    SPZ_WIDE_GO(8, false, 5)
#undef SPZ_WIDE_GO
}

}  // namespace spz

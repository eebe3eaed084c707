// spz_decode_lane.cu -- "lane-per-octet" decode kernels: the throughput path for frames of many chunks.
//
// The team kernels (spz_decode_hd.cu) spread one chunk over 8..32 lanes and pay ~180 warp instructions of
// cooperative bookkeeping per group (header scan shuffles, ring barriers, staged flush); ncu showed them
// instruction-issue bound at ~27 thread instructions per sample with 44 % of those being that overhead
// (profiles/r1_final_ncu_summary.csv).  When a frame has tens of thousands of chunks there is no need to
// cooperate at all: here ONE lane owns 8 consecutive dims of one chunk (ndims = 8: a lane per chunk,
// ndims = 16: two lanes per chunk) and is a plain scalar decoder, so there are no shuffles, votes or barriers
// and every per-group cost is paid once per 8 dims x 16 rows.
//
//   input    each lane pulls the chunk's stream through a private 128-byte ring in shared memory with
//            32-byte (one sector) loads issued one fill ahead.  The ring is word-interleaved across the warp
//            (word i of lane L lives at row i, column L), so a lane only ever touches its own bank: field
//            reads at lane-specific bit positions are conflict-free by construction.  A 33rd row mirrors row 0
//            so the two-word window never has to wrap.
//            Loads run one fill ahead in registers and four fills ahead as an L2 prefetch.
//   unpack   a lane's 8 fields of a row are CONTIGUOUS bits (<= 64 for 8-bit, 2 x <= 64 for 16-bit): three ring
//            words + two funnel shifts give the aligned 64-bit segment, then every field is "mask the low bits,
//            shift the segment right by its width" -- 2 shifts per field instead of an address computation,
//            two LDS and a funnel shift per field
//   predict  zigzag^-1 and delta | FIRE entirely in the left-aligned domain: m = u ^ -(u & 1) is 2*err, so
//            prev_delta<<(32-W) = (mulhi(prev_delta<<, coef) << (32-W)) + (m << (31-W)) and the running value
//            accumulates left-aligned too (natural W-bit wrap, PRMT picks the high halves on output); the 8
//            dims of a lane are independent -> instruction-level parallelism instead of occupancy
//   output   a lane owns 8 consecutive elements of every row: packed with PRMT and stored as one 64/128-bit
//            store per row (two rows per store when ndims * esz == 8); lanes of a chunk fill whole sectors
//   runs / tail / errors   plain divergent scalar code, as in the reference's decoder
//
// Reference behaviour restated: sprintz_delta_rle.cpp:417-772, sprintz_xff_rle.cpp:568-1179 (high-dim).
#include <atomic>

#include "spz_host.h"

namespace spz {

constexpr int kLaneThreads = 128;
#ifndef SPZ_LANE_MINB
#define SPZ_LANE_MINB 4
#endif
constexpr uint32_t kLaneWarpSmem = 8192;      // 33 rows x 128 B used; the stride keeps "row | lane" an OR

// Q = 0: decode.  Q = 1 / 2: query-on-compressed (query.hpp:22-29 REDUCE_MAX / REDUCE_SUM with
// materialize == false): the same decoder, but a row is folded into per-dim running maxima / sums instead of being
// stored, so the launch reads the compressed bytes only.  Lanes reduce across the warp, warps across the CTA, and
// one atomic per dim and CTA lands in p.qresult[ndims] (u64, zeroed by the caller).
template <int W, bool FIRE, int D, int Q>
__global__ void __launch_bounds__(kLaneThreads, SPZ_LANE_MINB) spz_decode_lane_kernel(const DecParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr int NP = D / 8;                       // lanes per chunk
    constexpr uint32_t FMASK = (1u << HB) - 1u;
    constexpr uint32_t H = (2u * D * HB + 7u) / 8u; // header bytes per group
    constexpr uint32_t B = 8u * D;                  // elements per block
    constexpr uint32_t LB = 8u * ESZ;               // bytes a lane owns in every row
    constexpr uint32_t ROWB = (uint32_t)D * ESZ;    // bytes per row
    static_assert(D % 8 == 0 && D <= 16, "lane kernels cover ndims 8 and 16");

    __shared__ __align__(16) uint8_t smem[(kLaneThreads / 32) * kLaneWarpSmem];

    const uint64_t gtid = (uint64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    uint64_t c = gtid / NP;
    const uint32_t part = (uint32_t)(gtid % NP);
    bool live = true;
    if (c >= p.nchunks) {
        if (!Q) return;                             // decoding lanes never cooperate: early exit is safe
        live = false; c = p.nchunks - 1;            // query lanes meet in the final reduction: shadow the last chunk
    }
    const uint32_t rboff = (threadIdx.x >> 5) * kLaneWarpSmem + (threadIdx.x & 31) * 4u;

    // ---- input ring
    // the directory is input from outside: an offset past the source makes the chunk "truncated", never a wild read
    const uint64_t soff = p.offsets ? p.offsets[c] : c * p.slot_stride;
    const bool bad_off = soff > (uint64_t)(p.src_end - p.src) || (uint64_t)(p.src_end - p.src) - soff < 8u;
    const uint8_t* const gstream = p.src + (bad_off ? 0 : soff);
    const uint32_t mis = (uint32_t)((uintptr_t)gstream & 31);
    const uint8_t* base32 = gstream - mis;
    asm volatile("" : "+l"(base32));                // keep the pointer in registers (no rematerialisation per fill)
    const uint64_t span = (uint64_t)(p.src_end - base32);               // readable bytes from base32
    const uint32_t limit = span > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)span;
    const uint32_t lim = (limit - 1u) & ~31u;       // last 32-byte block that still holds a readable byte
    uint32_t filled = 0;                            // stream bytes staged so far (multiple of 32)
    uint32_t nx[8];                                 // the next 32 stream bytes, loaded one fill ahead
    auto load_next = [&]() {
        const uint32_t o = filled < lim ? filled : lim;
        // one 256-bit load == one sector per lane (sm_100: LDG.E.256)
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(nx[0]), "=r"(nx[1]), "=r"(nx[2]), "=r"(nx[3]), "=r"(nx[4]), "=r"(nx[5]), "=r"(nx[6]), "=r"(nx[7])
                     : "l"(base32 + o));
    };
    auto fill = [&]() {
        const uint32_t so = ((filled << 5) & 0xC00u) | rboff;
#pragma unroll
        for (int j = 0; j < 8; j++) *reinterpret_cast<uint32_t*>(smem + so + j * 128) = nx[j];
        if ((filled & 96u) == 0u) {                  // once per trip around the ring
            // rows 32, 33 mirror rows 0, 1: a 3-word read never wraps
            *reinterpret_cast<uint32_t*>(smem + rboff + 32 * 128) = nx[0];
            *reinterpret_cast<uint32_t*>(smem + rboff + 33 * 128) = nx[1];
        }
        {   // the register prefetch below runs one fill ahead; an L1 prefetch two fills further ahead lets it hit L1
            const uint32_t o = filled + 96u < lim ? filled + 96u : lim;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base32 + o));
        }
        filled += 32u;
        load_next();
    };
    // 32-bit window of the stream at bit position `bit` (relative to base32)
    auto window = [&](uint32_t bit) -> uint32_t {
        const uint32_t a = ((bit << 2) & 0xF80u) | rboff;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + a);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + a + 128);
        return __funnelshift_r(w0, w1, bit);
    };
    // aligned 64-bit segment of the stream starting at bit position `bit`
    auto segment = [&](uint32_t bit, uint32_t& lo, uint32_t& hi) {
        const uint32_t a = ((bit << 2) & 0xF80u) | rboff;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + a);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + a + 128);
        const uint32_t w2 = *reinterpret_cast<const uint32_t*>(smem + a + 256);
        lo = __funnelshift_r(w0, w1, bit);
        hi = __funnelshift_r(w1, w2, bit);
    };
    // Every consumer takes at most 32 bytes between two calls, so one fill restores "48 bytes ahead"
    // (a unit of <= 32 bytes plus the 12-byte reach of a 3-word read from its last byte).
    // Fills are made warp-coherent: when any lane of the warp has to fill, every lane with room for 32 more
    // bytes fills too.  Lanes consume at similar rates, so this keeps their phases aligned: the warp executes the
    // fill sequence once per ~32 consumed bytes instead of once per lane that runs dry, and -- because the
    // scoreboard on the prefetch registers is per warp -- the one-fill-ahead load really has a whole fill
    // interval to land (with per-lane fills every STS waited on a load issued moments earlier for another lane).
    auto ensure = [&](uint32_t bit) {
        const uint32_t avail = filled - (bit >> 3);
        const bool need = avail < 48u;
        if (__any_sync(__activemask(), need)) {
            if (need || avail <= 88u) fill();
        }
    };

    load_next();
    fill(); fill(); fill();

    // ---- output
    const uint64_t first = c * p.chunk_len;
    uint_t* const out = reinterpret_cast<uint_t*>(p.dst) + first;
    const uint64_t cap64 = first >= p.dst_cap ? 0 : ((p.dst_cap - first < p.chunk_len) ? (p.dst_cap - first) : p.chunk_len);
    const uint32_t cap = cap64 > 0xffffffffull ? 0xffffffffu : (uint32_t)cap64;
    uint8_t* orow = reinterpret_cast<uint8_t*>(out) + part * LB;        // my bytes of the next row to write
    const bool al32 = ((uintptr_t)out & 31) == 0;                       // whole-sector stores may be 256-bit

    uint32_t pvh[8];                                // running values, left-aligned (value << (32-W))
    int32_t pdh[8], ct[8];                          // previous deltas (left-aligned), FIRE counters
#pragma unroll
    for (int j = 0; j < 8; j++) { pvh[j] = 0; pdh[j] = 0; ct[j] = 0; }
    // int8 rows are 8 bytes per lane: rows are held back so that a lane always completes a whole 32-byte sector with
    // back-to-back stores (partial-sector writes cost a read-modify-write in L2: measured on the low-dim decoder,
    // 16-byte stores -> whole sectors = +38 %).  NP == 1: rows r..r+3 leave as two 128-bit stores; NP == 2: the two
    // lanes of a chunk write rows r, r+1 (2 x 16 bytes, one sector) in consecutive instructions.
    uint32_t held[6] = {0, 0, 0, 0, 0, 0};

    // my 8 values of one row -> HBM
    auto store_row = [&](uint32_t r) {
        if (Q) {
            // rows are folded (qfold below): nothing is materialised
        } else if (W == 16) {
            uint4 v;
            v.x = __byte_perm(pvh[0], pvh[1], 0x7632); v.y = __byte_perm(pvh[2], pvh[3], 0x7632);
            v.z = __byte_perm(pvh[4], pvh[5], 0x7632); v.w = __byte_perm(pvh[6], pvh[7], 0x7632);
            *reinterpret_cast<uint4*>(orow) = v;
            orow += ROWB;
        } else {
            const uint32_t a = __byte_perm(pvh[0], pvh[1], 0x0073), b = __byte_perm(pvh[2], pvh[3], 0x0073);
            const uint32_t e = __byte_perm(pvh[4], pvh[5], 0x0073), f = __byte_perm(pvh[6], pvh[7], 0x0073);
            const uint32_t w0 = __byte_perm(a, b, 0x5410), w1 = __byte_perm(e, f, 0x5410);
            if (NP == 1) {                           // rows are 8 bytes: rows r..r+3 form one sector
                if ((r & 3u) != 3u) { held[2 * (r & 3u)] = w0; held[2 * (r & 3u) + 1] = w1; }
                else {
                    if (al32) {                      // one 256-bit store == one whole sector (sm_100: STG.E.256)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow), "r"(held[0]), "r"(held[1]),
                                     "r"(held[2]), "r"(held[3]), "r"(held[4]), "r"(held[5]), "r"(w0), "r"(w1) : "memory");
                    } else {
                        reinterpret_cast<uint4*>(orow)[0] = make_uint4(held[0], held[1], held[2], held[3]);
                        reinterpret_cast<uint4*>(orow)[1] = make_uint4(held[4], held[5], w0, w1);
                    }
                    orow += 4 * ROWB;
                }
            } else {                                 // 16-byte rows, 8 bytes per lane: rows r, r+1 form one sector
                if ((r & 1u) == 0u) { held[0] = w0; held[1] = w1; }
                else {
                    *reinterpret_cast<uint2*>(orow) = make_uint2(held[0], held[1]);
                    *reinterpret_cast<uint2*>(orow + ROWB) = make_uint2(w0, w1);
                    orow += 2 * ROWB;
                }
            }
        }
    };

    // ---- query accumulators (Q != 0): running maximum (left-aligned, order preserving) | 32-bit partial sums,
    // drained into lane-private 64-bit totals (rows 34..49 of the warp's shared region) every 4096 blocks
    uint32_t qa[8];
    uint32_t qblocks = 0;
    if (Q) {
#pragma unroll
        for (int j = 0; j < 8; j++) qa[j] = 0;
        if (Q == 2) {
#pragma unroll
            for (int j = 0; j < 16; j++) *reinterpret_cast<uint32_t*>(smem + rboff + (34 + j) * 128) = 0u;
        }
    }
    auto qdrain = [&]() {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t* lo = reinterpret_cast<uint32_t*>(smem + rboff + (34 + 2 * j) * 128);
            uint32_t* hi = reinterpret_cast<uint32_t*>(smem + rboff + (35 + 2 * j) * 128);
            const uint32_t t = *lo + qa[j];
            *hi += (t < qa[j]) ? 1u : 0u;
            *lo = t;
            qa[j] = 0;
        }
    };
    // fold `reps` identical rows (the current running values) into the accumulators
    auto qfold = [&](uint32_t reps) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (Q == 1) qa[j] = pvh[j] > qa[j] ? pvh[j] : qa[j];
            else if (reps == 1u) asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(qa[j]) : "r"(pvh[j]), "n"(1u << W));
            else {                                   // rare: a delta run; add the 64-bit product to the totals
                uint32_t* lo = reinterpret_cast<uint32_t*>(smem + rboff + (34 + 2 * j) * 128);
                uint32_t* hi = reinterpret_cast<uint32_t*>(smem + rboff + (35 + 2 * j) * 128);
                const unsigned long long t = (((unsigned long long)*hi << 32) | *lo) + qa[j]
                                             + (unsigned long long)(pvh[j] >> UP) * reps;
                *lo = (uint32_t)t; *hi = (uint32_t)(t >> 32);
                qa[j] = 0;
            }
        }
    };
    auto qblock_done = [&]() {
        if (Q == 2) { if ((++qblocks & 4095u) == 0u) qdrain(); }
    };

    uint32_t pos = mis * 8u;                         // bit cursor relative to base32
    uint32_t ngroups = 0, remaining = 0, status = kStatusOk;
    const uint32_t variant = p.variant;              // 1 / 2: the non-RLE sibling codecs (6-byte metadata, no run slots)
    if (variant) {                                   // format.h:75-85: [u32 len][u16 ndims]
        const uint32_t len32 = window(pos);
        const uint32_t w1 = window(pos + 32u);
        const uint32_t nd = variant == 3u ? (w1 >> 16) : (w1 & 0xffffu);       // variant 3: [u48 len][u16 ndims]
        pos += variant == 3u ? 64u : 48u;
        ensure(pos);
        if (len32 < kMinDataSize) { remaining = len32; }
        else {
            if (nd != (uint32_t)D) status = kStatusBadDims;
            ngroups = len32 / (16u * D);
            remaining = len32 - ngroups * 16u * D;
        }
    } else if (p.has_meta) {                         // format.h:47-62
        ngroups = window(pos);
        const uint32_t w1 = window(pos + 32u);
        remaining = w1 & 0xffffu;
        if ((w1 >> 16) != (uint32_t)D) status = kStatusBadDims;
        pos += 64u;
        ensure(pos);
    } else {
        ngroups = p.meta[c].ngroups;
        remaining = p.meta[c].remaining & 0xffffu;
    }
    if (bad_off) { status = kStatusTruncated; remaining = 0; }
    {   // a group is at least H + 2 bytes: a group count the chunk's own span cannot hold is corrupt (and would spin)
        const uint64_t src_bytes = (uint64_t)(p.src_end - p.src);
        const uint64_t next = p.offsets ? p.offsets[c + 1] : src_bytes;
        const uint64_t cspan = (next > soff && next <= src_bytes) ? next - soff : (src_bytes > soff ? src_bytes - soff : 0);
        // (the non-RLE siblings have no run slots: an all-zero block is a data slot with no payload at all)
        if ((uint64_t)ngroups * (H + (variant ? 0u : 2u)) > cspan) { status = kStatusTruncated; remaining = 0; }
    }
    if (status != kStatusOk) ngroups = 0;

    uint32_t o = 0;                                  // output element cursor of the chunk
#pragma unroll 1
    for (uint32_t g = 0; g < ngroups; g++) {
        if ((pos >> 3) > limit) { status = kStatusTruncated; break; }
        const uint32_t hbit = pos;
        pos += H * 8u;
        // both slots' header words are read before the ring moves on
        uint32_t mine[2], sum_before[2], total[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            mine[s] = window(hbit + (uint32_t)(s * D) * HB + part * (8u * HB));
            uint32_t before = 0, all = 0;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const uint32_t fw = (NP == 1) ? mine[s] : window(hbit + (uint32_t)(s * D + q * 8) * HB);
                // sum of the 8 widths of octet q: sum of the fields + the number of fields equal to W-1 (stored W)
                uint32_t sum;
                if (HB == 4) {
                    const uint32_t e = (fw & 0x0f0f0f0fu) + ((fw >> 4) & 0x0f0f0f0fu);      // 4 byte lanes, each <= 30
                    uint32_t full = fw & (fw >> 1); full &= full >> 2;                      // bit 0 of a nibble: nibble == 15
                    sum = ((e * 0x01010101u) >> 24) + (uint32_t)__popc(full & 0x11111111u);
                } else {
                    const uint32_t x = fw & 0x00ffffffu;
                    const uint32_t e = (x & 0007070707u) + ((x >> 3) & 0007070707u);        // 4 six-bit lanes, each <= 14
                    const uint32_t full = x & (x >> 1) & (x >> 2) & 0011111111u;            // bit 0 of a field: field == 7
                    sum = (((e * 0x00041041u) >> 18) & 63u) + (uint32_t)__popc(full);
                }
                all += sum;
                if ((uint32_t)q < part) before += sum;
            }
            sum_before[s] = before; total[s] = all;
        }
        ensure(pos);

#pragma unroll 1
        for (int s = 0; s < 2; s++) {
            const uint32_t tot = s ? total[1] : total[0];
            if (tot == 0u && variant == 0u) {
                // ---- run slot (sprintz_delta_rle.cpp:640-667, sprintz_xff_rle.cpp:828-958)
                const uint32_t rb = window(pos);
                uint32_t nrun = rb & 0x7fu;
                if (rb & 0x80u) { nrun |= ((rb >> 8) & 0xffu) << 7; pos += 16u; } else { pos += 8u; }
                ensure(pos);
                if ((uint64_t)o + (uint64_t)nrun * B > cap) { status = kStatusOverflow; break; }
                int32_t coef[8];
#pragma unroll
                for (int j = 0; j < 8; j++) coef[j] = FIRE ? fire_coef<W, false>(ct[j]) : 0;
                if (Q && !FIRE) {                    // a delta run repeats the previous row: fold it in one step
                    if (nrun) qfold(nrun * 8u);
                    o += nrun * B;
                    continue;
                }
#pragma unroll 1
                for (uint32_t b = 0; b < nrun; b++) {
                    if (Q) qblock_done();
#pragma unroll
                    for (uint32_t r = 0; r < 8; r++) {
                        if (FIRE) {
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                // err == 0 -> delta == prediction = (prev_delta * coef) >> W, wrapped to W bits
                                const int32_t pred = __mulhi(pdh[j], coef[j]);
                                pdh[j] = (int32_t)((uint32_t)pred << UP);
                                pvh[j] += (uint32_t)pdh[j];
                            }
                        }
                        if (Q) qfold(1u); else store_row(r);
                    }
                }
                o += nrun * B;
                continue;
            }

            // ---- data slot
            if (o + B > cap) { status = kStatusOverflow; break; }
            const uint32_t fw = s ? mine[1] : mine[0];
            uint32_t nb[8], fm[8];
            int32_t coef[8], gsh[8];
            const uint32_t offA = s ? sum_before[1] : sum_before[0];   // bit offset of my first dim inside a row
            uint32_t offB = offA;                                      // ... of my fifth dim (16-bit: second segment)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t f = (fw >> (j * HB)) & FMASK;
                const uint32_t n = f + (f == (uint32_t)(W - 1) ? 1u : 0u);
                nb[j] = n;
                if (j < 4) offB += n;
                fm[j] = (1u << n) - 1u;
                coef[j] = FIRE ? fire_coef<W, false>(ct[j]) : 0;
                gsh[j] = 0;
            }
            const uint32_t row_bits = ((tot + 7u) >> 3) * 8u;
            uint32_t rowbit = pos;
#pragma unroll
            for (uint32_t r = 0; r < 8; r++) {
                // int16: two 64-bit segments of 4 fields; int8: two 32-bit segments of 4 fields (one shift per field)
                uint32_t lo, hi = 0;
                if (W == 16) segment(rowbit + offA, lo, hi); else lo = window(rowbit + offA);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (j == 4) { if (W == 16) segment(rowbit + offB, lo, hi); else lo = window(rowbit + offB); }
                    // zigzag^-1 (bitpack.h:303) in doubled form: m = u ^ -(u & 1) == 2 * err (bit 0 is always clear).
                    // The kernel is bound by the half-rate INT32 ALU pipe (ncu: ALU 71 %, FMA 20 %), so the sign
                    // handling is phrased as multiply-adds: with b = u & 1 and t = 1 - 2b, m = u * t - b, and the
                    // same t is icopysign's factor in the gradient.
                    const uint32_t u = lo & fm[j];
                    const uint32_t b = lo & fm[j] & 1u;
                    if (j != 7 && j != 3) {
                        if (W == 16) { lo = __funnelshift_r(lo, hi, nb[j]); hi >>= nb[j]; } else { lo >>= nb[j]; }
                    }
                    int32_t t, mi;
                    asm("mad.lo.s32 %0, %1, -2, 1;" : "=r"(t) : "r"(b));
                    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(mi) : "r"(u), "r"(t), "r"(0u - b));
                    const uint32_t m = (uint32_t)mi;
                    if (FIRE) {
                        if (r & 1u) {
                            // icopysign(err, prev_delta) (util.h:63-68), accumulated left-aligned so the int_t wrap
                            // of grad_sum is the natural 32-bit wrap
                            if (m != 0u) gsh[j] += pdh[j] * t;
                        }
                        // prediction = (prev_delta * coef) >> W (sprintz_xff_rle.cpp:225): prev_delta is left-aligned
                        // by 32-W bits, so this is the high word of the product; delta = prediction + err, and
                        // (delta << UP) = (prediction << UP) + (m << (UP-1)) wraps to W bits by itself
                        const int32_t th = __mulhi(pdh[j], coef[j]);
                        int32_t ths;
                        asm("mad.lo.s32 %0, %1, %2, 0;" : "=r"(ths) : "r"(th), "n"(1 << UP));
                        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(pdh[j]) : "r"(mi), "n"(1 << (UP - 1)), "r"(ths));
                        pvh[j] += (uint32_t)pdh[j];
                    } else if (variant == 2u) {
                        pvh[j] = u << UP;                                    // plain bit packing: the field is the value
                    } else {
                        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(pvh[j]) : "r"(mi), "n"(1 << (UP - 1)));
                    }
                }
                if (Q) qfold(1u); else store_row(r);
                rowbit += row_bits;
                // a row is at most D*W/8 bytes: refill checks are only needed every 32 bytes of worst-case consumption
                constexpr uint32_t CHK = (D * W / 8 >= 32) ? 1u : (32u / (D * W / 8));
                if ((r + 1u) % CHK == 0u) ensure(rowbit);
            }
            pos = rowbit;
            if (FIRE) {
#pragma unroll
                for (int j = 0; j < 8; j++) ct[j] = sx<CB>(ct[j] + ((gsh[j] >> UP) >> 2));   // sprintz_xff_rle.cpp:273-275
            }
            o += B;
            if (Q) qblock_done();
        }
        if (status != kStatusOk) break;
    }

    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (status == kStatusOk) {
        if ((uint64_t)o + remaining > cap) {
            status = kStatusOverflow;
        } else if ((uint64_t)(pos >> 3) + (uint64_t)remaining * ESZ > (uint64_t)limit) {
            status = kStatusTruncated;              // the raw tail is read straight from the source
        } else {
            const uint8_t* tsrc = base32 + (pos >> 3);
            if (Q) {
                // raw tail elements go straight to the result (short, rare): element e belongs to dim e % D
                if (live && part == 0) {
                    for (uint32_t e = 0; e < remaining; e++) {
                        const unsigned long long v = (W == 16) ? (unsigned long long)(tsrc[2 * e] | (tsrc[2 * e + 1] << 8))
                                                               : (unsigned long long)tsrc[e];
                        if (Q == 1) atomicMax(p.qresult + (o + e) % D, v); else atomicAdd(p.qresult + (o + e) % D, v);
                    }
                }
            } else {
                uint8_t* tdst = reinterpret_cast<uint8_t*>(out + o);
                const uint32_t nb = remaining * ESZ;
                for (uint32_t b = part; b < nb; b += NP) tdst[b] = tsrc[b];
            }
            o += remaining;
        }
    }
    if (Q) {
        // ---- lanes -> warp -> CTA -> one atomic per dim
        unsigned long long v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (Q == 1) v[j] = qa[j] >> UP;
            else {
                const uint32_t lo = *reinterpret_cast<uint32_t*>(smem + rboff + (34 + 2 * j) * 128);
                const uint32_t hi = *reinterpret_cast<uint32_t*>(smem + rboff + (35 + 2 * j) * 128);
                v[j] = (((unsigned long long)hi << 32) | lo) + qa[j];
            }
            if (!live || status != kStatusOk) v[j] = 0;
#pragma unroll
            for (int off = NP; off < 32; off <<= 1) {
                const unsigned long long w = __shfl_xor_sync(0xffffffffu, v[j], off);
                v[j] = (Q == 1) ? (w > v[j] ? w : v[j]) : (v[j] + w);
            }
        }
        __syncthreads();                             // everyone is done with the rings: reuse the first bytes
        unsigned long long* cta = reinterpret_cast<unsigned long long*>(smem);
        if ((threadIdx.x & 31) < NP) {
#pragma unroll
            for (int j = 0; j < 8; j++) cta[(threadIdx.x >> 5) * D + part * 8 + j] = v[j];
        }
        __syncthreads();
        if (threadIdx.x < D) {
            unsigned long long r = cta[threadIdx.x];
#pragma unroll
            for (int w = 1; w < kLaneThreads / 32; w++) {
                const unsigned long long x = cta[w * D + threadIdx.x];
                r = (Q == 1) ? (x > r ? x : r) : (r + x);
            }
            if (Q == 1) { if (r) atomicMax(p.qresult + threadIdx.x, r); } else atomicAdd(p.qresult + threadIdx.x, r);
        }
        if (!live) return;
    }
    if (part == 0 && p.meta) {
        p.meta[c].status = status;
        if (p.has_meta || variant) {
            p.meta[c].nbytes = (pos >> 3) - mis + remaining * ESZ;
            p.meta[c].ngroups = ngroups;
            p.meta[c].remaining = o;
        }
    }
}

// ------------------------------------------------------------------ launch

static std::atomic<int> g_force_path{0};       // 0 auto, 1 team kernels only, 2 lane kernels whenever the shape allows (process-wide test switch)
void set_force_path(int v) { g_force_path.store(v); }
int get_force_path() { return g_force_path.load(); }

template <int W, bool FIRE, int D, int Q>
static cudaError_t launch_lane_q(const DecParams& p, cudaStream_t stream)
{
    const uint64_t threads = p.nchunks * (uint64_t)(D / 8);
    const uint64_t grid64 = (threads + kLaneThreads - 1) / kLaneThreads;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    count_launch();
    spz_decode_lane_kernel<W, FIRE, D, Q><<<(uint32_t)grid64, kLaneThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

template <int W, bool FIRE, int D>
static cudaError_t launch_lane(const DecParams& p, cudaStream_t stream)
{
    if (p.qop == 1) return launch_lane_q<W, FIRE, D, 1>(p, stream);
    if (p.qop == 2) return launch_lane_q<W, FIRE, D, 2>(p, stream);
    const uint64_t threads = p.nchunks * (uint64_t)(D / 8);
    const uint64_t grid64 = (threads + kLaneThreads - 1) / kLaneThreads;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    count_launch();
    spz_decode_lane_kernel<W, FIRE, D, 0><<<(uint32_t)grid64, kLaneThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape is not covered (caller continues with the team kernels).
cudaError_t launch_decode_lane(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    const int force = get_force_path();
    if (force == 1) return cudaErrorNotSupported;
    if (p.ndims != 8 && p.ndims != 16) return cudaErrorNotSupported;
    if (p.src_end == nullptr || p.src_end <= p.src) return cudaErrorNotSupported;
    // a lane is a whole scalar decoder: only worth it when there are enough lane tasks to fill the machine (one wave
    // is 148 SMs x 4 CTAs x 128 lanes = 75 776).  Measured crossover against the team kernels at 1024-row chunks
    // (gpurun s43, 64..512 MiB frames): ~20-28 Ki lane tasks for int8, ~50 Ki for int16.
    if (force != 2 && p.nchunks * (p.ndims / 8) < (elem_sz == 1 ? 24576u : 49152u)) return cudaErrorNotSupported;
    // vector stores: every chunk's output must start 16-byte aligned
    if (!p.qop && ((uintptr_t)p.dst & 15) != 0) return cudaErrorNotSupported;
    if (!p.qop && p.nchunks > 1 && ((p.chunk_len * (uint64_t)elem_sz) & 15) != 0) return cudaErrorNotSupported;
    if (p.chunk_len % p.ndims != 0 && p.nchunks > 1) return cudaErrorNotSupported;
    // 32-bit bit cursors inside one chunk
    if (p.chunk_len * (uint64_t)elem_sz > 0x0c000000ull) return cudaErrorNotSupported;
#define SPZ_LANE(Wv, Dv) return predictor ? launch_lane<Wv, true, Dv>(p, stream) : launch_lane<Wv, false, Dv>(p, stream)
    if (elem_sz == 1) { if (p.ndims == 8) SPZ_LANE(8, 8); else SPZ_LANE(8, 16); }
    else              { if (p.ndims == 8) SPZ_LANE(16, 8); else SPZ_LANE(16, 16); }
#undef SPZ_LANE
}

}  // namespace spz

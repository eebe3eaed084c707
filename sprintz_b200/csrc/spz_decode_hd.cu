// spz_decode_hd.cu -- fast high-dim decode kernels (compile-time team size / dims-per-lane).
//
// Same work decomposition as spz_decode.cu (one team of lanes per chunk, lane-per-dim serial recurrence),
// rebuilt after two ncu passes:
//   pass 1 (generic kernel): L1TEX-bound -- every field cost two global loads that touched one line per team.
//   pass 2 (first fast kernel): INT32-ALU-pipe bound (72 %) + ~6 instructions of MATCH/VOTE per sub-warp sync.
//
// What this version does about it:
//   input    the compressed stream is pulled through a per-team shared-memory ring with coalesced 128-bit
//            loads issued one fill ahead (register double buffering); a field is two broadcast LDS + one
//            funnel shift
//   sync     the kernel is WARP-SYNCHRONOUS: the 32/TEAM teams of a warp step through their chunks group by
//            group in lock-step, per-team differences (fewer groups, run slots, errors) are predicates, so
//            every shuffle / __syncwarp uses the full mask (one instruction, no convergence test)
//   pipes    INT32 ALU and FMA pipes are both half rate on sm_100; the recurrence is phrased so they share
//            the work: prev_delta is kept left-aligned (delta << (32-W)) so the FIRE prediction is one
//            IMAD.HI, the W-bit wrap is one IMAD.SHL, and the gradient sum wraps for free in the top bits
//   headers  both slots of a group are parsed at once; the two prefix sums share one packed shuffle scan
//   output   decoded rows are staged in a double-buffered shared tile and leave as 128-bit coalesced stores
//
// Reference behaviour restated: sprintz_delta_rle.cpp:417-772, sprintz_xff_rle.cpp:568-1179 (high-dim only;
// the low-dim column-major variants stay on the generic kernel).
#include "spz_host.h"

namespace spz {

constexpr unsigned kFull = 0xffffffffu;

template <int W, bool FIRE, int TEAM, int DPL, bool FULLD>
__global__ void __launch_bounds__(kThreads) spz_decode_hd_kernel(const DecParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;                      // left-alignment shift of W-bit signed values
    constexpr uint32_t FILL = TEAM * 16;            // bytes one ring fill brings in
    constexpr uint32_t FMASK = (1u << HB) - 1u;
    constexpr int TEAMS_PER_WARP = 32 / TEAM;
    // Several dims per lane: give a lane CONSECUTIVE dims (d = t*DPL + k) instead of strided ones whenever the
    // cover is exact and a lane's fields of one row fit a 64-bit segment.  Its header fields and its payload
    // fields are then contiguous bit strings: one scan per group instead of DPL, three ring words + two funnel
    // shifts per row instead of two LDS + a funnel shift per field (ncu on the 80-dim config: LSU pipe 67 % busy).
    constexpr bool CONTIG = FULLD && DPL > 1 && DPL * W <= 64 && DPL * Traits<W>::HB <= 24;

    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int t = lane & (TEAM - 1);
    const int team_in_cta = (int)(threadIdx.x / TEAM);
    const uint64_t c64 = (uint64_t)blockIdx.x * (kThreads / TEAM) + team_in_cta;
    const bool valid = c64 < p.nchunks;
    const uint64_t c = valid ? c64 : 0;            // invalid teams shadow chunk 0 but never store
    const int D = FULLD ? TEAM * DPL : (int)p.ndims;

    // team_smem is a multiple of ring_cap, so the ring's offset inside the CTA's dynamic shared memory is
    // ring_cap-aligned and "wrap + add base" is a single LOP3: (byte & mask) | ring_off
    const uint32_t ring_off = (uint32_t)team_in_cta * p.team_smem;
    uint8_t* const ring = smem + ring_off;                              // ring_cap bytes + 16 mirror bytes
    const uint32_t RCAP = p.ring_cap;
    const uint32_t RMASK = RCAP - 1u;
    // teams of a warp run in lock-step at similar stream positions and their rings are bank-aligned: store
    // stream byte x at ring position (x + skew) so their word reads fall into different banks
    const uint32_t skew = (uint32_t)(team_in_cta % TEAMS_PER_WARP) * 32u;
#define WMASK4 (p.ring_mask4)          /* constant-bank operand: (byte & mask) | ring_off is a single LOP3 */
    // two output tiles of out_cap bytes; skewed by 32 B per team of the warp so the lock-step teams' row stores
    // land in different banks (team regions are ring_cap-aligned, i.e. bank-aligned)
    uint8_t* const obuf0 = ring + RCAP + 16 + (uint32_t)(team_in_cta % TEAMS_PER_WARP) * 32u;
    const uint32_t OCAP = p.out_cap;

    // ---- input: byte offsets are relative to `abase`, the 16-byte aligned-down stream start
    // the directory is input from outside: an offset past the source makes the chunk "truncated", never a wild read
    const uint64_t soff = p.offsets ? p.offsets[c] : c * p.slot_stride;
    const uint64_t src_bytes = (uint64_t)(p.src_end - p.src);
    const bool bad_off = soff > src_bytes || src_bytes - soff < 8u;
    const uint8_t* const gstream = p.src + (bad_off ? 0 : soff);
    const uint32_t mis_in = (uint32_t)((uintptr_t)gstream & 15);
    const uint8_t* const abase = gstream - mis_in;
    const uint64_t lim64 = (uint64_t)(p.src_end - abase);
    const uint32_t limit = lim64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)lim64;   // readable bytes from abase
    uint32_t wr = 0;                                // stream bytes staged so far (team-uniform)
    uint4 nxt;

    auto prefetch = [&]() {
        const uint32_t o = wr + (uint32_t)t * 16u;
        if (o + 16u <= limit) {
            nxt = __ldg(reinterpret_cast<const uint4*>(abase + o));
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; b < 16; b++)
                if (o + b < limit) w[b >> 2] |= (uint32_t)abase[o + b] << (8 * (b & 3));
            nxt = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    // Make stream bytes [.., need_end) available in the ring. Called at warp-uniform points; teams that do
    // not need a fill idle through it.
    auto ensure = [&](uint32_t need_end, bool on) {
        while (__any_sync(kFull, on && wr < need_end)) {
            __syncwarp();                            // readers of the bytes about to be overwritten are done
            if (on && wr < need_end) {
                const uint32_t ro = (wr + skew + (uint32_t)t * 16u) & RMASK;
                *reinterpret_cast<uint4*>(ring + ro) = nxt;
                if (ro == 0) *reinterpret_cast<uint4*>(ring + RCAP) = nxt;  // mirror: word reads run 4 bytes past the end
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
    // same, with the byte position tracked separately by the caller (byte == bit >> 3): saves the shift
    auto ring_bits2 = [&](uint32_t bit, uint32_t byte) -> uint32_t {
        const uint32_t bo = (byte & WMASK4) | ring_off;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + bo);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + bo + 4);
        return __funnelshift_r(w0, w1, bit);
    };
    auto ring_byte = [&](uint32_t at) -> uint32_t { return ring[(at + skew) & RMASK]; };

    // ---- output
    const uint64_t first = c * p.chunk_len;
    uint_t* const out = reinterpret_cast<uint_t*>(p.dst) + first;
    const uint64_t cap64 = (!valid || first >= p.dst_cap) ? 0 :
        ((p.dst_cap - first < p.chunk_len) ? (p.dst_cap - first) : p.chunk_len);
    const uint32_t cap = cap64 > 0xffffffffull ? 0xffffffffu : (uint32_t)cap64;
    // Output staging: tile `ob` holds bytes for the 16-byte aligned global address gdst; the first `skip`
    // bytes precede this chunk's region. Two tiles alternate so a flush needs a single barrier.
    uint8_t* ob = obuf0;
    uint8_t* gdst;
    uint32_t ocur, oskip;
    {
        const uint32_t mis = (uint32_t)((uintptr_t)out & 15);
        gdst = reinterpret_cast<uint8_t*>(out) - mis; ocur = mis; oskip = mis;
    }
    // Writes staged bytes [oskip, hi) to HBM; hi = whole 16-byte units unless `final_flush`. The remainder
    // moves to the other tile, which becomes current.
    auto flush = [&](bool on, bool final_flush) {
        __syncwarp();
        if (on) {
            uint8_t* const nb = (ob == obuf0) ? obuf0 + OCAP : obuf0;
            if (FULLD && oskip == 0u && ocur == (uint32_t)(2 * kBlockRows * TEAM * DPL * ESZ)) {
                // common case: aligned output, exactly one group staged = DPL*ESZ 16-byte units per lane
                const uint4* s4 = reinterpret_cast<const uint4*>(ob);
                uint4* g4 = reinterpret_cast<uint4*>(gdst);
#pragma unroll
                for (int j = 0; j < DPL * ESZ; j++) g4[t + j * TEAM] = s4[t + j * TEAM];
                ob = nb; gdst += ocur; ocur = 0;
            } else if (oskip == 0u && (ocur & 15u) == 0u) {
                const uint4* s4 = reinterpret_cast<const uint4*>(ob);
                uint4* g4 = reinterpret_cast<uint4*>(gdst);
                for (uint32_t u = t; u < (ocur >> 4); u += TEAM) g4[u] = s4[u];
                ob = nb; gdst += ocur; ocur = 0;
            } else {
                const uint32_t hi = final_flush ? ocur : (ocur & ~15u);
                if (hi > oskip) {
                    uint32_t a = (oskip + 15u) & ~15u;
                    if (a > hi) a = hi;
                    for (uint32_t b = oskip + t; b < a; b += TEAM) gdst[b] = ob[b];
                    const uint4* s4 = reinterpret_cast<const uint4*>(ob);
                    uint4* g4 = reinterpret_cast<uint4*>(gdst);
                    for (uint32_t u = (a >> 4) + t; u < (hi >> 4); u += TEAM) g4[u] = s4[u];
                    uint32_t tb = hi & ~15u;
                    if (tb < a) tb = a;
                    for (uint32_t b = tb + t; b < hi; b += TEAM) gdst[b] = ob[b];
                    const uint32_t rem = ocur - hi;
                    for (uint32_t b = t; b < rem; b += TEAM) nb[b] = ob[hi + b];
                    ob = nb; gdst += hi; ocur = rem; oskip = 0;
                }
            }
        }
    };

    // ---- predictor state (registers); prev_delta is kept left-aligned: pdh = prev_delta << (32-W)
    uint32_t pv[DPL];
    int32_t pdh[DPL], ct[DPL];
#pragma unroll
    for (int k = 0; k < DPL; k++) { pv[k] = 0; pdh[k] = 0; ct[k] = 0; }

    prefetch();
    uint32_t pos = mis_in;              // consume offset from abase (team-uniform)
    uint32_t ngroups = 0, remaining = 0, status = kStatusOk;
    if (p.has_meta) {                   // format.h:47-62
        ensure(pos + 8, true);
        ngroups = ring_byte(pos) | (ring_byte(pos + 1) << 8) | (ring_byte(pos + 2) << 16) | (ring_byte(pos + 3) << 24);
        remaining = ring_byte(pos + 4) | (ring_byte(pos + 5) << 8);
        const uint32_t nd = ring_byte(pos + 6) | (ring_byte(pos + 7) << 8);
        if (nd != (uint32_t)D) status = kStatusBadDims;
        pos += 8;
    } else {
        ngroups = p.meta[c].ngroups;
        remaining = p.meta[c].remaining & 0xffffu;
    }
    if (bad_off) { status = kStatusTruncated; remaining = 0; }
    {   // a group is at least H + 2 bytes: a group count the chunk's own span cannot hold is corrupt (and would spin)
        const uint64_t next = p.offsets ? p.offsets[c + 1] : src_bytes;
        const uint64_t cspan = (next > soff && next <= src_bytes) ? next - soff : (src_bytes > soff ? src_bytes - soff : 0);
        const uint64_t hmin = ((2ull * (uint64_t)D * HB + 7ull) >> 3) + 2ull;
        if ((uint64_t)ngroups * hmin > cspan) { status = kStatusTruncated; remaining = 0; }
    }
    if (!valid || status != kStatusOk) ngroups = 0;

    uint32_t o = 0;                     // output element cursor (team-uniform)
    const uint32_t H = (2u * (uint32_t)D * HB + 7u) >> 3;
    const uint32_t B = (uint32_t)kBlockRows * D;
    const uint32_t max_group = H + 2u * B * ESZ;
    const uint32_t row_out = (uint32_t)D * ESZ;         // output bytes per row

    const uint32_t gmax = __reduce_max_sync(kFull, ngroups);
    for (uint32_t g = 0; g < gmax; g++) {
        if (g < ngroups && pos > limit) status = kStatusTruncated;     // ran past the readable bytes (the ring was fed zeros)
        const bool act = g < ngroups && status == kStatusOk;
        ensure(pos + max_group + 8, act);

        // ---- both slots' widths: fields (s*D + d); the prefix sums over dims share one packed scan
        const uint32_t hbit = ((pos + skew) & RMASK) * 8u;
        uint32_t nn[DPL], offs[DPL];
        uint32_t carry = 0;
        if (CONTIG) {
            const uint32_t fw0 = ring_bits(hbit + (uint32_t)(t * DPL) * HB);
            const uint32_t fw1 = ring_bits(hbit + (uint32_t)(D + t * DPL) * HB);
            uint32_t lanesum = 0;
#pragma unroll
            for (int k = 0; k < DPL; k++) {
                const uint32_t f0 = (fw0 >> (k * HB)) & FMASK, f1 = (fw1 >> (k * HB)) & FMASK;
                const uint32_t n0 = (f0 == (uint32_t)(W - 1)) ? (uint32_t)W : f0;
                const uint32_t n1 = (f1 == (uint32_t)(W - 1)) ? (uint32_t)W : f1;
                nn[k] = n0 | (n1 << 16);
                offs[k] = lanesum;                  // exclusive prefix inside the lane
                lanesum += nn[k];
            }
            uint32_t incl = lanesum;
#pragma unroll
            for (int s = 1; s < TEAM; s <<= 1) {
                asm volatile("{ .reg .pred p; .reg .b32 r; shfl.sync.up.b32 r|p, %0, %1, %2, 0xffffffff; @p add.u32 %0, %0, r; }"
                             : "+r"(incl) : "r"(s), "r"((32 - TEAM) << 8));
            }
            const uint32_t base = incl - lanesum;
#pragma unroll
            for (int k = 0; k < DPL; k++) offs[k] += base;
            carry = __shfl_sync(kFull, incl, TEAM - 1, TEAM);
        } else {
#pragma unroll
        for (int k = 0; k < DPL; k++) {
            const int d = k * TEAM + t;
            uint32_t packed = 0;
            if (FULLD || d < D) {
                const uint32_t f0 = ring_bits(hbit + (uint32_t)d * HB) & FMASK;
                const uint32_t f1 = ring_bits(hbit + (uint32_t)(D + d) * HB) & FMASK;
                const uint32_t n0 = (f0 == (uint32_t)(W - 1)) ? (uint32_t)W : f0;
                const uint32_t n1 = (f1 == (uint32_t)(W - 1)) ? (uint32_t)W : f1;
                packed = n0 | (n1 << 16);
            }
            uint32_t incl = packed;
#pragma unroll
            for (int s = 1; s < TEAM; s <<= 1) {
                // shfl.up's predicate says whether the source lane is inside my team segment: predicated add, no select
                asm volatile("{ .reg .pred p; .reg .b32 r; shfl.sync.up.b32 r|p, %0, %1, %2, 0xffffffff; @p add.u32 %0, %0, r; }"
                             : "+r"(incl) : "r"(s), "r"((32 - TEAM) << 8));
            }
            nn[k] = packed;
            offs[k] = carry + incl - packed;
            carry += __shfl_sync(kFull, incl, TEAM - 1, TEAM);
        }
        }
        if (act) pos += H;

        // One data slot: unpack + inverse predictor for my dims, rows staged at ob + ocur.
        auto data_slot = [&](uint32_t sh16, uint32_t total) {
            uint32_t row_bytes = (total + 7u) >> 3;
            asm volatile("" : "+r"(row_bytes));      // keep it a register: byte += row_bytes stays an add (FMA-pipe IMAD.IADD)
            const uint32_t row_bits = row_bytes * 8u;
            const uint32_t pbit = ((pos + skew) & RMASK) * 8u;
            if (CONTIG) {
                // row-major over my DPL consecutive dims: one aligned 64-bit segment per row
                uint8_t* op = ob + ocur + (uint32_t)(t * DPL) * ESZ;
                uint32_t nk[DPL], fmk[DPL];
                int32_t coef[DPL], gsh[DPL];
#pragma unroll
                for (int k = 0; k < DPL; k++) {
                    nk[k] = (nn[k] >> sh16) & 0xffffu;
                    fmk[k] = (1u << nk[k]) - 1u;
                    coef[k] = FIRE ? fire_coef<W, false>(ct[k]) : 0;
                    gsh[k] = 0;
                }
                uint32_t bit = pbit + ((offs[0] >> sh16) & 0xffffu);
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) {
                    const uint32_t bo = ((bit >> 3) & WMASK4) | ring_off;
                    const uint32_t w0 = *reinterpret_cast<const uint32_t*>(smem + bo);
                    const uint32_t w1 = *reinterpret_cast<const uint32_t*>(smem + bo + 4);
                    const uint32_t w2 = *reinterpret_cast<const uint32_t*>(smem + bo + 8);
                    uint32_t lo = __funnelshift_r(w0, w1, bit), hi = __funnelshift_r(w1, w2, bit);
                    bit += row_bits;
#pragma unroll
                    for (int k = 0; k < DPL; k++) {
                        const uint32_t xk = lo;
                        if (k != DPL - 1) {
                            if (DPL * W <= 32) { lo >>= nk[k]; } else { lo = __funnelshift_r(lo, hi, nk[k]); hi >>= nk[k]; }
                        }
                        const uint32_t b = xk & fmk[k] & 1u;
                        const int32_t err = (int32_t)((xk & fmk[k]) ^ (0u - b)) >> 1;
                        int32_t delta = err;
                        if (FIRE) {
                            if (r & 1) { if (err != 0) gsh[k] += pdh[k] * (int32_t)(1u - 2u * b); }
                            delta = __mulhi(pdh[k], coef[k]) + err;
                            pdh[k] = (int32_t)((uint32_t)delta << UP);
                        }
                        pv[k] += (uint32_t)delta;
                        *reinterpret_cast<uint_t*>(op + k * ESZ) = (uint_t)pv[k];
                    }
                    op += row_out;
                }
                if (FIRE) {
#pragma unroll
                    for (int k = 0; k < DPL; k++) ct[k] = sx<CB>(ct[k] + ((gsh[k] >> UP) >> 2));
                }
            } else {
            uint8_t* const obase = ob + ocur + (uint32_t)t * ESZ;
#pragma unroll
            for (int k = 0; k < DPL; k++) {
                const int d = k * TEAM + t;
                if (FULLD || d < D) {
                    const uint32_t n = (nn[k] >> sh16) & 0xffffu;
                    const uint32_t fmask = (1u << n) - 1u;
                    const uint32_t smask = fmask & 1u;         // the zigzag sign bit (0 when the field is empty)
                    uint32_t bit = pbit + ((offs[k] >> sh16) & 0xffffu);
                    uint32_t byte = bit >> 3;
                    uint32_t x[kBlockRows];
#pragma unroll
                    for (int r = 0; r < kBlockRows; r++) { x[r] = ring_bits2(bit, byte); bit += row_bits; byte += row_bytes; }

                    int32_t coef = 0, gsh = 0;
                    if (FIRE) coef = fire_coef<W, false>(ct[k]);
                    uint8_t* op = obase + (uint32_t)(k * TEAM) * ESZ;
#pragma unroll
                    for (int r = 0; r < kBlockRows; r++) {
                        // zigzag^-1 (bitpack.h:303) fused with the field mask: err = ((x & fmask) ^ -b) >> 1, b = sign bit
                        const uint32_t b = x[r] & smask;
                        const int32_t err = (int32_t)((x[r] & fmask) ^ (0u - b)) >> 1;
                        int32_t delta = err;
                        if (FIRE) {
                            if (r & 1) {
                                // icopysign(err, prev_delta) (util.h:63-68), accumulated left-aligned so the
                                // int_t wrap of grad_sum is the natural 32-bit wrap
                                if (err != 0) gsh += pdh[k] * (int32_t)(1u - 2u * b);
                            }
                            // prediction = (prev_delta * coef) >> W (sprintz_xff_rle.cpp:225): with prev_delta left-aligned
                            // by 32-W bits this is exactly the high word of the product -> one IMAD.HI with addend
                            delta = __mulhi(pdh[k], coef) + err;
                            pdh[k] = (int32_t)((uint32_t)delta << UP);       // wraps delta to W bits
                        }
                        pv[k] += (uint32_t)delta;
                        *reinterpret_cast<uint_t*>(op) = (uint_t)pv[k];
                        op += row_out;
                    }
                    if (FIRE) ct[k] = sx<CB>(ct[k] + ((gsh >> UP) >> 2));     // sprintz_xff_rle.cpp:273-275
                }
            }
            }
            pos += row_bytes * kBlockRows;
            ocur += B * ESZ;
            o += B;
        };

        const uint32_t tot0 = carry & 0xffffu, tot1 = carry >> 16;
        // Fast path: every team of the warp has two data slots that fit.
        if (__all_sync(kFull, act && tot0 != 0u && tot1 != 0u && o + 2u * B <= cap)) {
            data_slot(0u, tot0);
            data_slot(16u, tot1);
            flush(true, false);
            continue;
        }

#pragma unroll 1
        for (uint32_t s = 0; s < 2; s++) {
            const uint32_t sh16 = 16u * s;
            const uint32_t total = s ? tot1 : tot0;
            const bool is_run = act && status == kStatusOk && total == 0;
            if (__any_sync(kFull, is_run)) {
                // ---- run slot (sprintz_delta_rle.cpp:640-667): rare; the whole warp walks through it, teams
                // without a run idle. Staged bytes leave first, then rows are stored directly.
                uint32_t nrows = 0;
                if (is_run) {
                    uint32_t nrun = ring_byte(pos) & 0x7f;
                    if (ring_byte(pos) & 0x80) { nrun |= ring_byte(pos + 1) << 7; pos += 2; } else { pos += 1; }
                    nrows = nrun * kBlockRows;
                    if ((uint64_t)o + (uint64_t)nrows * D > cap) { status = kStatusOverflow; nrows = 0; }
                }
                const bool go = is_run && status == kStatusOk;
                flush(go, true);
                if (go) {
#pragma unroll
                    for (int k = 0; k < DPL; k++) {
                        const int d = CONTIG ? (t * DPL + k) : (k * TEAM + t);
                        if (FULLD || d < D) {
                            uint_t* op = out + o + d;
                            if (FIRE) {
                                const int32_t coef = fire_coef<W, false>(ct[k]);
                                for (uint32_t r = 0; r < nrows; r++) {
                                    // err == 0 -> delta == prediction = (prev_delta * coef) >> W, wrapped to W bits
                                    const int32_t pred = __mulhi(pdh[k], coef);
                                    pdh[k] = (int32_t)((uint32_t)pred << UP);
                                    pv[k] += (uint32_t)pred;
                                    op[(size_t)r * D] = (uint_t)pv[k];
                                }
                            } else {
                                for (uint32_t r = 0; r < nrows; r++) op[(size_t)r * D] = (uint_t)pv[k];
                            }
                        }
                    }
                    o += nrows * D;
                    const uint32_t mis = (uint32_t)((uintptr_t)(out + o) & 15);
                    gdst = reinterpret_cast<uint8_t*>(out + o) - mis; ocur = mis; oskip = mis;
                }
            }

            // ---- data slot
            bool data = act && status == kStatusOk && total != 0;
            if (data && o + B > cap) { status = kStatusOverflow; data = false; }
            if (data) data_slot(sh16, total);
        }
        flush(act, false);
    }
    flush(valid, true);

    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (valid && status == kStatusOk) {
        if ((uint64_t)o + remaining > cap) {
            status = kStatusOverflow;
        } else if ((uint64_t)pos + (uint64_t)remaining * ESZ > (uint64_t)limit) {
            status = kStatusTruncated;              // the raw tail is read straight from the source
        } else {
            const uint8_t* tsrc = abase + pos;
            uint8_t* tdst = reinterpret_cast<uint8_t*>(out + o);
            const uint32_t nb = remaining * ESZ;
            for (uint32_t b = t; b < nb; b += TEAM) tdst[b] = tsrc[b];
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
}

#undef WMASK4

// ------------------------------------------------------------------ launch

template <int W, bool FIRE, int TEAM, int DPL, bool FULLD>
static cudaError_t launch_hd(const DecParams& p, uint32_t grid, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(spz_decode_hd_kernel<W, FIRE, TEAM, DPL, FULLD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    count_launch();
    spz_decode_hd_kernel<W, FIRE, TEAM, DPL, FULLD><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline uint32_t pow2_ceil(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }

// Returns cudaErrorNotSupported when the shape is not covered by the fast path (caller falls back to the
// generic kernel, which is still CUDA -- there is no CPU path anywhere).
cudaError_t launch_decode_hd(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    const bool low = elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2);
    if (low || p.ndims > 256 || p.src_end == nullptr) return cudaErrorNotSupported;
    // 32-bit stream / output offsets inside one chunk
    if (p.chunk_len * (uint64_t)elem_sz > 0x70000000ull) return cudaErrorNotSupported;
    // team shape: a power-of-two slice of the warp with DPL dims per lane. Prefer an exact cover of ndims
    // (no padded lanes, compile-time row stride); D=80 -> 16x5, D=24 -> 8x3, D=48 -> 16x3.
    uint32_t team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
    uint32_t dpl = (p.ndims + team - 1) / team;
    if (team * dpl != p.ndims) {
        if (p.ndims == 80 || p.ndims == 48) { team = 16; dpl = p.ndims / 16; }
        else if (p.ndims == 24 || p.ndims == 40) { team = 8; dpl = p.ndims / 8; }
    }
    const bool fulld = (team * dpl == p.ndims);
    const uint32_t hb = elem_sz == 1 ? 3 : 4;
    const uint32_t H = (2 * p.ndims * hb + 7) / 8;
    const uint32_t max_group = H + 2 * 8 * p.ndims * elem_sz;
    p.team = team; p.dpl = dpl;
    p.ring_cap = pow2_ceil(max_group + 8 + team * 16);
    p.ring_mask4 = (p.ring_cap - 1u) & ~3u;
    p.out_cap = (16 + 2 * 8 * p.ndims * elem_sz + 16 + 15) & ~15u;
    // the two output tiles start at ring_cap + 16 + (team index within the warp) * 32 (bank skew, see the kernel):
    // up to 224 bytes for 4-lane teams
    const uint32_t skew_max = (32u / team - 1u) * 32u;
    p.team_smem = ((p.ring_cap + 16 + 2 * p.out_cap + (skew_max > 128u ? skew_max : 128u) + p.ring_cap - 1) / p.ring_cap) * p.ring_cap;
    const uint32_t teams_per_cta = kThreads / team;
    p.teams_per_cta = teams_per_cta;
    const size_t smem = (size_t)teams_per_cta * p.team_smem;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const uint64_t grid64 = (p.nchunks + teams_per_cta - 1) / teams_per_cta;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    const uint32_t grid = (uint32_t)grid64;

#define SPZ_HD2(Wv, Tv, Dv, Fv) \
    return predictor ? launch_hd<Wv, true, Tv, Dv, Fv>(p, grid, smem, stream) : launch_hd<Wv, false, Tv, Dv, Fv>(p, grid, smem, stream)
#define SPZ_HD(Wv, Tv, Dv) do { if (fulld) { SPZ_HD2(Wv, Tv, Dv, true); } else { SPZ_HD2(Wv, Tv, Dv, false); } } while (0)
    if (elem_sz == 1) {
        if (team == 8 && dpl == 1) SPZ_HD(8, 8, 1);
        if (team == 8 && dpl == 3) SPZ_HD2(8, 8, 3, true);
        if (team == 8 && dpl == 5) SPZ_HD2(8, 8, 5, true);
        if (team == 16 && dpl == 1) SPZ_HD(8, 16, 1);
        if (team == 16 && dpl == 3) SPZ_HD2(8, 16, 3, true);
        if (team == 16 && dpl == 5) SPZ_HD2(8, 16, 5, true);
        if (dpl == 1) SPZ_HD(8, 32, 1);
        if (dpl == 2) SPZ_HD(8, 32, 2);
        if (dpl == 3) SPZ_HD(8, 32, 3);
        if (dpl == 4) SPZ_HD(8, 32, 4);
        if (dpl == 5) SPZ_HD(8, 32, 5);
        if (dpl == 6) SPZ_HD(8, 32, 6);
        if (dpl == 7) SPZ_HD(8, 32, 7);
        if (dpl == 8) SPZ_HD(8, 32, 8);
    } else {
        if (team == 4) SPZ_HD(16, 4, 1);
        if (team == 8 && dpl == 1) SPZ_HD(16, 8, 1);
        if (team == 8 && dpl == 3) SPZ_HD2(16, 8, 3, true);
        if (team == 8 && dpl == 5) SPZ_HD2(16, 8, 5, true);
        if (team == 16 && dpl == 1) SPZ_HD(16, 16, 1);
        if (team == 16 && dpl == 3) SPZ_HD2(16, 16, 3, true);
        if (team == 16 && dpl == 5) SPZ_HD2(16, 16, 5, true);
        if (dpl == 1) SPZ_HD(16, 32, 1);
        if (dpl == 2) SPZ_HD(16, 32, 2);
        if (dpl == 3) SPZ_HD(16, 32, 3);
        if (dpl == 4) SPZ_HD(16, 32, 4);
        if (dpl == 5) SPZ_HD(16, 32, 5);
        if (dpl == 6) SPZ_HD(16, 32, 6);
        if (dpl == 7) SPZ_HD(16, 32, 7);
        if (dpl == 8) SPZ_HD(16, 32, 8);
    }
#undef SPZ_HD
#undef SPZ_HD2
    return cudaErrorNotSupported;
}

}  // namespace spz

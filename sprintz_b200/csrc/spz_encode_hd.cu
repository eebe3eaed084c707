// spz_encode_hd.cu -- fast high-dim encode kernels (compile-time team size / dims-per-lane).
//
// Same contract as spz_encode.cu (one team of lanes per chunk; every chunk is byte-identical to the
// reference's stream for that slice: sprintz_delta_rle.cpp:54-404, sprintz_xff_rle.cpp:60-555), rebuilt as a
// WARP-SYNCHRONOUS kernel after the first ncu pass (generic kernel: 113 issue slots per sample, 71 % lane
// utilisation, one 8-byte global load per team per row):
//
//   input    16 rows (one group) at a time through shared memory: 128-bit coalesced loads issued one group
//            ahead (register double buffering), lane-per-dim reads of the staged tile
//   phase 1  lane-per-dim forecast -> zigzag -> width, all 8 rows independent (the FIRE prediction only
//            needs the previous *input* delta and the block's coefficient); prev_delta is kept left-aligned
//            so the W-bit wraps are free and the work splits across the (half-rate) ALU and FMA pipes
//   control  the run-length / group state machine is a per-team *action generator* (run slot, pad, data
//            slot); the warp executes actions in lock-step passes, so every shuffle / barrier is full-mask
//   phase 2  at group end, lane-per-row bit-emit: fields are first combined into <=32-bit clusters by
//            multiply-add (multipliers 2^prefix precomputed per block), clusters are appended to the row with
//            one 64-bit multiply-add each; header nibbles are gathered into words by a few lanes
//   output   a double-buffered shared staging tile, flushed with 128-bit stores once per group
#include "spz_host.h"

namespace spz {


constexpr unsigned kFullE = 0xffffffffu;

enum : uint32_t { ACT_NONE = 0, ACT_RUN = 1, ACT_PAD = 2, ACT_DATA = 3 };

template <int W, bool FIRE, int TEAM, int DPL, bool FULLD>
__global__ void __launch_bounds__(kThreads) spz_encode_hd_kernel(const EncParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr uint32_t WMASK = (1u << W) - 1u;
    constexpr int CL = 32 / W;                      // fields per cluster (<= 32 bits together)
    constexpr int Dp = TEAM * DPL;                  // padded dims
    constexpr int NCL = Dp / CL;                    // clusters per row
    constexpr int NVEC = DPL * ESZ;                 // 16-byte vectors per lane per 16-row group (upper bound)
    constexpr uint32_t ZS = ((((uint32_t)Dp * ESZ + 3u) >> 2) | 1u) * 4u;   // zz tile row stride (odd #words)
    constexpr bool HFAST = FULLD && TEAM >= 8;      // D % 8 == 0: header bytes are owned by octets of dims

    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int t = lane & (TEAM - 1);
    const int team_in_cta = (int)(threadIdx.x / TEAM);
    const uint64_t c64 = (uint64_t)blockIdx.x * (kThreads / TEAM) + team_in_cta;
    const bool valid = c64 < p.nchunks;
    const uint64_t c = valid ? c64 : 0;
    const int D = FULLD ? Dp : (int)p.ndims;

    // ---- shared memory of this team
    uint8_t* const sm = smem + (size_t)team_in_cta * p.team_smem;
    uint8_t* const s_in = sm;                                           // [16][D*ESZ] staged input rows
    uint8_t* const s_zz = sm + p.nb_off;                                // [16][ZS]     zigzagged errors (2 halves of 8 rows)
    uint32_t* const s_mul = reinterpret_cast<uint32_t*>(sm + p.tile_off);   // [2][Dp] cluster-local multipliers 2^prefix
    uint8_t* const s_cw = reinterpret_cast<uint8_t*>(s_mul + 2 * Dp);   // [2][Dp] cluster widths (at the cluster's first dim)
    uint8_t* const s_hf = s_cw + 2 * Dp;                                // [2*Dp]  header nibbles of the open group
    uint8_t* const stage0 = sm + p.out_off;                             // two staging tiles of out_cap bytes
    const uint32_t OCAP = p.out_cap;
    const uint32_t max_group_bytes = ((2u * (FULLD ? (uint32_t)Dp : p.ndims) * HB + 7u) >> 3) + 2u * kBlockRows * (FULLD ? (uint32_t)Dp : p.ndims) * ESZ + 4u;

    // ---- this chunk
    const uint64_t first = c * p.chunk_len;
    const int64_t L = valid ? (int64_t)((first + p.chunk_len <= p.len) ? p.chunk_len : (p.len - first)) : 0;
    const uint8_t* const gin = p.src + first * ESZ;                     // 16-byte aligned (checked on the host)
    const uint32_t in_bytes = (uint32_t)(L * ESZ);
    uint8_t* const gstream = p.dst + c * p.slot_stride;
    const uint32_t B = (uint32_t)kBlockRows * D;
    const int64_t G = 2 * (int64_t)B;
    const uint32_t H = (2u * (uint32_t)D * HB + 7u) >> 3;
    const uint32_t meta_bytes = p.write_size ? 8u : 0u;
    const uint32_t grp_bytes = 2u * B * ESZ;                            // input bytes per 16-row group
    const uint32_t nvec_grp = (uint32_t)D * ESZ;                        // 16-byte vectors per group

    // ---- input prefetch (registers), one group ahead
    uint4 nxt[NVEC];
    auto prefetch = [&](uint32_t pair) {
        const uint32_t base = pair * grp_bytes;
#pragma unroll
        for (int j = 0; j < NVEC; j++) {
            const uint32_t v = (uint32_t)j * TEAM + t;
            const uint32_t o = base + v * 16u;
            nxt[j] = make_uint4(0, 0, 0, 0);
            if (FULLD || v < nvec_grp) {
                if (o + 16u <= in_bytes) {
                    nxt[j] = __ldg(reinterpret_cast<const uint4*>(gin + o));
                } else if (o < in_bytes) {          // the vector straddles the end of the chunk's input
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (uint32_t b = 0; b < 16u; b++)
                        if (o + b < in_bytes) w[b >> 2] |= (uint32_t)gin[o + b] << (8u * (b & 3u));
                    nxt[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    };

    // ---- output staging (double buffered)
    uint8_t* ob = stage0;
    uint32_t ob_sel = 0;                    // which of the two staging tiles is current
    uint8_t* gdst;
    uint32_t ocur, oskip;
    {
        const uint32_t mis = (uint32_t)((uintptr_t)gstream & 15);
        gdst = gstream - mis; ocur = mis; oskip = mis;
    }
    auto flush = [&](bool on, bool final_flush) {
        __syncwarp();
        if (on) {
            ob_sel ^= 1u;
            uint8_t* const nb = stage0 + ob_sel * OCAP;
            if (!final_flush && oskip == 0u) {
                // common case: whole 16-byte units leave, the (<16 byte) remainder moves to the other tile
                const uint32_t hi = ocur & ~15u;
                const uint4* s4 = reinterpret_cast<const uint4*>(ob);
                uint4* g4 = reinterpret_cast<uint4*>(gdst);
                for (uint32_t u = t; u < (hi >> 4); u += TEAM) g4[u] = s4[u];
                const uint32_t rem = ocur - hi;
                if ((uint32_t)t < rem) nb[t] = ob[hi + t];
                if (TEAM < 16 && (uint32_t)t + TEAM < rem) nb[t + TEAM] = ob[hi + t + TEAM];
                if (TEAM < 8 && (uint32_t)t + 2 * TEAM < rem) { nb[t + 2 * TEAM] = ob[hi + t + 2 * TEAM]; }
                if (TEAM < 8 && (uint32_t)t + 3 * TEAM < rem) { nb[t + 3 * TEAM] = ob[hi + t + 3 * TEAM]; }
                ob = nb; gdst += hi; ocur = rem;
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
                } else {
                    ob_sel ^= 1u;           // nothing left: stay on the current tile
                }
            }
        }
    };
    if (t == 0) for (uint32_t b = 0; b < meta_bytes; b++) ob[ocur + b] = 0;     // metadata placeholder
    ocur += meta_bytes;

    // ---- predictor state (registers), values left-aligned: xh = x << (32-W)
    uint32_t pvh[DPL];
    int32_t pdh[DPL], ct[DPL];
    uint32_t hf0[DPL], hf1[DPL];            // header nibbles of the open group (slot 0 / slot 1)
#pragma unroll
    for (int k = 0; k < DPL; k++) { pvh[k] = 0; pdh[k] = 0; ct[k] = 0; hf0[k] = 0; hf1[k] = 0; }

    // ---- team-uniform encoder state (sprintz_delta_rle.cpp:132-158)
    uint32_t ngroups = 0, nslots = 0, run = 0, blk = 0;     // blk = next block index
    uint32_t hdr_off = 0;
    uint32_t pend = 0, wh = 0;              // pending data blocks of the open group; zz half phase 1 writes next
    uint32_t pe_half0 = 0, pe_off0 = 0, pe_rb0 = 0, pe_half1 = 0, pe_off1 = 0, pe_rb1 = 0;
    int64_t tail = 0;
    bool outer = true;
    bool on = valid && L >= (int64_t)kMinDataSize;
    const bool strict = !FIRE;              // `<` vs `<=` in the continue-run test (high-dim: sprintz_xff_rle.cpp:362)
    // blk*B <= L-G  <=>  blk <= last_start ; blk*B < L-G  <=>  blk <= last_strict   (both -1 when L < G)
    const int32_t last_start = (L >= G) ? (int32_t)((L - G) / B) : -1;
    const int32_t last_strict = (L > G) ? (int32_t)((L - G - 1) / B) : -1;
    const int32_t last_cont = strict ? last_strict : last_start;

    // Group finished: bit-emit the pending data blocks, (generic shapes: gather the header), flush.
    // Warp-uniform call.
    auto close_group = [&](bool gc) {
        if (!HFAST && gc) {
#pragma unroll
            for (int k = 0; k < DPL; k++) {
                const int d = k * TEAM + t;
                if (FULLD || d < D) { s_hf[d] = (uint8_t)hf0[k]; s_hf[D + d] = (uint8_t)hf1[k]; }
                hf0[k] = 0; hf1[k] = 0;
            }
        }
        __syncwarp();
        if (gc) {
            // ---- row jobs: job j -> (pending entry j>>3, row j&7). Rows are emitted byte by byte (4 at a time):
            // no alignment cases, no branches besides the flush predicate.
            for (uint32_t j = t; j < pend * 8u; j += TEAM) {
                const uint32_t e = j >> 3, r = j & 7u;
                const uint32_t half = e ? pe_half1 : pe_half0;
                const uint32_t rb = e ? pe_rb1 : pe_rb0;
                uint8_t* dstp = ob + (e ? pe_off1 : pe_off0) + r * rb;
                const uint_t* zrow = reinterpret_cast<const uint_t*>(s_zz + (half * 8u + r) * ZS);
                const uint4* mul4 = reinterpret_cast<const uint4*>(s_mul + half * Dp);
                const uint8_t* cw = s_cw + half * Dp;
                uint64_t acc = 0;
                uint32_t pos = 0;
#pragma unroll 2
                for (int q = 0; q < NCL; q++) {
                    uint32_t qv;
                    // the cluster's zigzagged values come in as ONE aligned word (rows are 4-byte aligned, a cluster
                    // is 4 bytes): the kernel is L1 / LSU bound (ncu: l1tex 94 %), byte loads were 1 LSU op per sample
                    const uint32_t zw = reinterpret_cast<const uint32_t*>(zrow)[q];
                    if (CL == 4) {
                        const uint4 m = mul4[q];
                        qv = (zw & 0xffu) * m.x + __byte_perm(zw, 0, 0x4441) * m.y +
                             __byte_perm(zw, 0, 0x4442) * m.z + (zw >> 24) * m.w;
                    } else {
                        const uint2 m = reinterpret_cast<const uint2*>(mul4)[q];
                        qv = (zw & 0xffffu) * m.x + (zw >> 16) * m.y;
                    }
                    const uint32_t p2 = 1u << pos;
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(qv), "r"(p2));   // fields are disjoint: add == or
                    pos += cw[q * CL];
                    if (pos >= 32u) {
                        const uint32_t wv = (uint32_t)acc;
                        dstp[0] = (uint8_t)wv; dstp[1] = (uint8_t)(wv >> 8); dstp[2] = (uint8_t)(wv >> 16); dstp[3] = (uint8_t)(wv >> 24);
                        dstp += 4; acc >>= 32; pos -= 32u;
                    }
                }
                const uint32_t wv = (uint32_t)acc;      // < 32 bits left
                if (pos > 0u)  dstp[0] = (uint8_t)wv;
                if (pos > 8u)  dstp[1] = (uint8_t)(wv >> 8);
                if (pos > 16u) dstp[2] = (uint8_t)(wv >> 16);
                if (pos > 24u) dstp[3] = (uint8_t)(wv >> 24);
            }
            if (!HFAST) {
                // ---- header words: lane j gathers the nibbles overlapping bits [32j, 32j+32) of the group header
                const uint32_t nfields = 2u * (uint32_t)D;
                for (uint32_t j = t; j * 4u < H; j += TEAM) {
                    const uint32_t f_lo = (32u * j) / HB;
                    uint32_t f_hi = (32u * j + 31u) / HB;
                    if (f_hi > nfields - 1u) f_hi = nfields - 1u;
                    const int32_t sh0 = (int32_t)(f_lo * HB) - (int32_t)(32u * j);     // in (-HB, 0]
                    uint32_t w = (uint32_t)s_hf[f_lo] >> (uint32_t)(-sh0);
                    for (uint32_t f = f_lo + 1u; f <= f_hi; f++) w |= (uint32_t)s_hf[f] << (f * HB - 32u * j);
                    for (uint32_t b = 0; b < 4u; b++)
                        if (4u * j + b < H) ob[hdr_off + 4u * j + b] = (uint8_t)(w >> (8u * b));
                }
            }
            pend = 0;
        }
        __syncwarp();       // the row jobs have read the zigzag tile: the next block's phase 1 may overwrite it (racecheck)
        // several groups accumulate in the staging tile; it leaves for HBM when the next group might not fit
        const bool need = gc && ocur + max_group_bytes + 16u > OCAP;
        if (__any_sync(kFullE, need)) flush(need, false);
    };

    // One slot of the open group (team-predicated by the caller).
    auto emit_slot = [&](uint32_t act, uint32_t run_len, const uint32_t (&fh)[DPL], uint32_t row_bytes) -> bool {
        const uint32_t sl = nslots & 1u;
        if (sl == 0u) {
            hdr_off = ocur; ocur += H; ngroups++;
        }
        if (act == ACT_RUN) {       // zero header fields + 1..2 length bytes (sprintz_delta_rle.cpp:238-245)
            const uint32_t two = run_len > 0x7fu ? 1u : 0u;
            if (t == 0) {
                ob[ocur] = (uint8_t)((run_len & 0x7fu) | (two << 7));
                if (two) ob[ocur + 1] = (uint8_t)(run_len >> 7);
            }
            ocur += 1u + two;
        } else if (act == ACT_PAD) {
            if (t == 0) ob[ocur] = 0;
            ocur += 1u;
        } else {                    // data slot: payload bytes are produced at group end
            if (pend == 0u) { pe_half0 = wh; pe_off0 = ocur; pe_rb0 = row_bytes; }
            else            { pe_half1 = wh; pe_off1 = ocur; pe_rb1 = row_bytes; }
            pend++;
            if (!HFAST) {
#pragma unroll
                for (int k = 0; k < DPL; k++) { if (sl == 0u) hf0[k] = fh[k]; else hf1[k] = fh[k]; }
            }
            ocur += kBlockRows * row_bytes;
            wh ^= 1u; blk++;
        }
        nslots++;
        return (nslots & 1u) == 0u;
    };
    // HFAST shapes write a slot's header bytes as soon as the slot is emitted: D is a multiple of 8, so every 8
    // dims own 3 (8-bit) / 4 (16-bit) whole header bytes and an xor-butterfly over the octet assembles them; run
    // and pad slots write zeros. Called at a warp-uniform point: `doit` = my team emitted a slot, `isdata` = it
    // was a data slot, `sl` = its index in the group.
    auto emit_header_fast = [&](bool doit, bool isdata, uint32_t sl, uint32_t hoff, const uint32_t (&fh)[DPL]) {
#pragma unroll
        for (int k = 0; k < DPL; k++) {
            uint32_t v = isdata ? (fh[k] << (HB * (t & 7))) : 0u;
            v |= __shfl_xor_sync(kFullE, v, 1);
            v |= __shfl_xor_sync(kFullE, v, 2);
            v |= __shfl_xor_sync(kFullE, v, 4);
            const uint32_t bidx = (uint32_t)(t & 7);
            if (doit && bidx < (uint32_t)HB) {
                const uint32_t byte0 = ((sl * (uint32_t)D + (uint32_t)((k * TEAM + t) & ~7)) * HB) >> 3;
                ob[hoff + byte0 + bidx] = (uint8_t)(v >> (8u * bidx));
            }
        }
    };

    prefetch(0);
    for (uint32_t it = 0;; it++) {
        if (on && outer && (int32_t)blk > last_start) { tail = (int64_t)blk * B; on = false; }
        if (!__any_sync(kFullE, on)) break;

        // ---- stage the next 16 input rows (every other block)
        if ((it & 1u) == 0) {
            __syncwarp();                           // phase-1 readers of the previous 16 rows are done
            if (on) {
#pragma unroll
                for (int j = 0; j < NVEC; j++) {
                    const uint32_t v = (uint32_t)j * TEAM + t;
                    if (FULLD || v < nvec_grp) reinterpret_cast<uint4*>(s_in)[v] = nxt[j];
                }
                prefetch((it >> 1) + 1u);
            }
            __syncwarp();
        }

        // ================= phase 1: lane-per-dim forecasting of this block
        uint32_t nw[DPL], fh[DPL], offs[DPL], incls[DPL];
        uint32_t total = 0;
#pragma unroll
        for (int k = 0; k < DPL; k++) {
            const int d = k * TEAM + t;
            uint32_t n = 0;
            uint8_t* zcol = s_zz + (wh * 8u) * ZS + (uint32_t)d * ESZ;
            if (FULLD || d < D) {
                const uint_t* in = reinterpret_cast<const uint_t*>(s_in) + (it & 1u) * B + d;
                uint32_t xh[kBlockRows];
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) xh[r] = (uint32_t)in[r * D] << UP;
                int32_t coef = 0, gsh = 0;
                if (FIRE) coef = fire_coef<W, false>(ct[k]);
                uint32_t mask = 0;
                int32_t dprev = pdh[k];
                uint32_t xprev = pvh[k];
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) {
                    const int32_t dh = (int32_t)(xh[r] - xprev);            // delta, left-aligned (wraps to W bits)
                    int32_t eh = dh;
                    if (FIRE) {
                        // err = delta - ((prev_delta * coef) >> W)  (sprintz_xff_rle.cpp:225-237): prev_delta is left-aligned
                        // by 32-W bits, so the prediction is exactly the high word of the product
                        const int32_t pred = __mulhi(dprev, coef);
                        eh = dh - (int32_t)((uint32_t)pred << UP);
                    }
                    // zigzag (bitpack.h:302) of the left-aligned error
                    const uint32_t zz = (((uint32_t)eh >> (UP - 1)) ^ (uint32_t)(eh >> 31)) & WMASK;
                    if (FIRE && (r & 1)) {
                        // grad_sum += icopysign(err, prev_delta) (util.h:63-68), left-aligned so it wraps like int_t
                        if (eh > 0) gsh += dprev; else if (eh < 0) gsh -= dprev;
                    }
                    mask |= zz;
                    *reinterpret_cast<uint_t*>(zcol + (uint32_t)r * ZS) = (uint_t)zz;
                    xprev = xh[r]; dprev = dh;
                }
                if (on) {
                    pvh[k] = xprev; pdh[k] = dprev;
                    if (FIRE) ct[k] = sx<CB>(ct[k] + ((gsh >> UP) >> 2));      // sprintz_xff_rle.cpp:273-275
                }
                n = nbits_rule<W, false>(mask);
            } else {
#pragma unroll
                for (int r = 0; r < kBlockRows; r++) *reinterpret_cast<uint_t*>(zcol + (uint32_t)r * ZS) = 0;
            }
            uint32_t incl = n;
#pragma unroll
            for (int s = 1; s < TEAM; s <<= 1) {
                asm volatile("{ .reg .pred p; .reg .b32 r; shfl.sync.up.b32 r|p, %0, %1, %2, 0xffffffff; @p add.u32 %0, %0, r; }"
                             : "+r"(incl) : "r"(s), "r"((32 - TEAM) << 8));
            }
            nw[k] = n;
            fh[k] = n - (n == (uint32_t)W);                    // w stored as w-1 (sprintz_delta_rle.cpp:199)
            offs[k] = total + incl - n;
            incls[k] = total + incl;
            total += __shfl_sync(kFullE, incl, TEAM - 1, TEAM);
        }
        // cluster-local multipliers / widths for phase 2
#pragma unroll
        for (int k = 0; k < DPL; k++) {
            const int d = k * TEAM + t;
            const int cs = t & ~(CL - 1);
            const uint32_t off_cs = __shfl_sync(kFullE, offs[k], cs, TEAM);
            const uint32_t inc_ce = __shfl_sync(kFullE, incls[k], cs + CL - 1, TEAM);
            s_mul[wh * Dp + d] = 1u << (offs[k] - off_cs);
            if (t == cs) s_cw[wh * Dp + d] = (uint8_t)(inc_ce - off_cs);
        }
        const bool zero = (total == 0);
        const uint32_t row_bytes = (total + 7u) >> 3;

        // ================= control
        // Fast path: every live team has a plain data block (no run open, block not all-zero).
        if (__all_sync(kFullE, !on || (!zero && run == 0u))) {
            bool gc = false;
            const uint32_t sl = nslots & 1u;
            if (on) { gc = emit_slot(ACT_DATA, 0u, fh, row_bytes); outer = gc; }
            if (HFAST) emit_header_fast(on, on, sl, hdr_off, fh);
            if (__any_sync(kFullE, gc)) close_group(gc);
            continue;
        }
        // General case: per-team action generator (run slot / pad / data slot), executed in lock-step passes.
        // st: 0 = block just read; 1 = run flushed because the data ran out, maybe pad; 2 = finished; 3 = block consumed
        uint32_t st = on ? 0u : 3u;
        for (;;) {
            uint32_t act = ACT_NONE, run_len = 0;
            if (st == 0u) {
                if (zero && run < kMaxRun) {
                    run++; blk++;
                    if ((int32_t)blk <= last_cont) { outer = false; st = 3u; }
                    else { act = ACT_RUN; run_len = run; run = 0; st = 1u; }    // not enough data to finish the group
                } else if (run > 0) {
                    act = ACT_RUN; run_len = run; run = 0;                      // a run just ended; re-evaluate this block
                } else {
                    act = ACT_DATA; st = 3u;
                }
            } else if (st == 1u) {
                if (nslots & 1u) { act = ACT_PAD; }
                st = 2u;
            }
            if (st == 2u && act == ACT_NONE) { tail = (int64_t)blk * B; on = false; st = 3u; }
            if (!__any_sync(kFullE, act != ACT_NONE)) break;

            bool gc = false;
            const uint32_t sl = nslots & 1u;
            if (act != ACT_NONE) {
                gc = emit_slot(act, run_len, fh, row_bytes);
                if (act == ACT_DATA) outer = gc;
            }
            if (HFAST) emit_header_fast(act != ACT_NONE, act == ACT_DATA, sl, hdr_off, fh);
            if (__any_sync(kFullE, gc)) close_group(gc);
        }
    }

    // ---- trailing raw elements (sprintz_delta_rle.cpp:398-403); also the whole input when L < 128
    {
        const int64_t rem_elems = valid ? (L - tail) : 0;
        const uint8_t* tsrc = gin + tail * ESZ;
        uint32_t left = (uint32_t)rem_elems * ESZ, done = 0;
        while (__any_sync(kFullE, left > 0)) {
            const uint32_t space = OCAP - ocur;
            const uint32_t take = left < space ? left : space;
            for (uint32_t b = t; b < take; b += TEAM) ob[ocur + b] = tsrc[done + b];
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

// ------------------------------------------------------------------ launch

template <int W, bool FIRE, int TEAM, int DPL, bool FULLD>
static cudaError_t launch_hd(const EncParams& p, uint32_t grid, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(spz_encode_hd_kernel<W, FIRE, TEAM, DPL, FULLD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (g_enc_probe) return probe_fill(g_enc_probe, 1, spz_encode_hd_kernel<W, FIRE, TEAM, DPL, FULLD>, kThreads, smem, kThreads / TEAM);
    count_launch();
    spz_encode_hd_kernel<W, FIRE, TEAM, DPL, FULLD><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline uint32_t pow2_ceil(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }
static inline uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

// Returns cudaErrorNotSupported when the shape is not covered by the fast path (the caller then uses the
// generic CUDA kernel).
cudaError_t launch_encode_hd(int elem_sz, int predictor, EncParams p, cudaStream_t stream)
{
    const bool low = elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2);
    if (low || p.ndims > 256) return cudaErrorNotSupported;
    if (p.chunk_len * (uint64_t)elem_sz > 0x70000000ull) return cudaErrorNotSupported;     // 32-bit offsets
    // 128-bit input loads: every chunk must start 16-byte aligned
    if (((uintptr_t)p.src & 15) || (p.nchunks > 1 && ((p.chunk_len * elem_sz) & 15))) return cudaErrorNotSupported;
    // team shape: a power-of-two slice of the warp with DPL dims per lane. Prefer an exact cover of ndims
    // (no padded lanes, compile-time row stride); D=80 -> 16x5, D=24 -> 8x3, D=48 -> 16x3.
    uint32_t team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
    uint32_t dpl = (p.ndims + team - 1) / team;
    if (team * dpl != p.ndims) {
        if (p.ndims == 80 || p.ndims == 48) { team = 16; dpl = p.ndims / 16; }
        else if (p.ndims == 24 || p.ndims == 40) { team = 8; dpl = p.ndims / 8; }
    }
    const bool fulld = (team * dpl == p.ndims);
    const uint32_t Dp = team * dpl;
    const uint32_t hb = elem_sz == 1 ? 3 : 4;
    const uint32_t H = (2 * p.ndims * hb + 7) / 8;
    p.team = team; p.dpl = dpl;
    // layout: [s_in][s_zz][s_mul | s_cw | s_hf][stage x2]
    uint32_t off = up16(16 * p.ndims * elem_sz);
    p.nb_off = off;                                         // s_zz
    const uint32_t ZS = ((((Dp * (uint32_t)elem_sz) + 3u) >> 2) | 1u) * 4u;
    off += up16(16 * ZS);
    p.tile_off = off;                                       // s_mul, s_cw, s_hf
    off += up16(2 * Dp * 4 + 2 * Dp + 2 * Dp);
    p.out_off = off;
    {
        const uint32_t grp = H + 2 * 8 * p.ndims * elem_sz + 4;
        uint32_t ngrp = 4;
        while (ngrp > 1 && ngrp * grp > 1536) ngrp--;           // keep the two tiles small enough for 8 CTAs/SM
        p.out_cap = up16(32 + ngrp * grp + 16);
        if (p.out_cap < 256) p.out_cap = 256;
    }
    p.team_smem = off + 2 * p.out_cap;
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

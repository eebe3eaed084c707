// spz_encode_lane.cu -- "lane-per-chunk" encode kernels: the throughput path for frames of many chunks.
//
// The team kernels (spz_encode_hd.cu) spread one chunk over 8..32 lanes: lane-per-dim forecasting, a
// warp-synchronous action generator for the run/group state machine and a lane-per-row bit emit, ~45-50 thread
// instructions per sample, most of it cooperation overhead (ncu: profiles/r1_final_ncu_summary.csv).  With tens
// of thousands of chunks in a frame no cooperation is needed: here ONE lane encodes ONE chunk as a plain scalar
// encoder -- the reference's control flow restated literally (sprintz_delta_rle.cpp:134-312) -- and the only
// shared-memory traffic is to lane-private columns, conflict-free by construction.
//
//   input    a block (8 rows) lives in a lane-private staging column ([vec][lane] 16-byte interleave: 128-bit
//            accesses never conflict).  The NEXT block's sectors are copied asynchronously (cp.async, 2 x 16 bytes)
//            over the sectors of the current block as soon as those have been emitted, so HBM/L2 latency hides
//            behind the rest of the bit emit and nobody holds the data in registers; an L2 prefetch runs one block ahead.
//   pass 1   octet-major: 8 dims x 8 rows at a time; values, deltas and FIRE errors stay left-aligned
//            (prediction = one IMAD.HI, error = one IMAD, natural W-bit wrap), zigzag, OR-masks packed 2 or 4
//            dims per register; the zigzagged errors overwrite the raw block in place
//   widths   per-dim width rule -> header fields, per-row layout (cluster multipliers / bit counts in registers)
//   control  run-length / group pairing / tail rules exactly as the reference's state machine
//   pass 2   row-major: two (16-bit) or four (8-bit) fields are merged into a <= 32-bit cluster with
//            multiply-adds, a cluster is appended to a 64-bit accumulator with one IMAD.WIDE; full words go to a
//            lane-private output ring and leave for HBM as 256-bit stores.  The header of a group precedes both
//            of its slots, so slot 1's fields are patched in afterwards (ring or HBM, wherever the bytes are).
//
// Reference behaviour restated: sprintz_delta_rle.cpp:54-404, sprintz_xff_rle.cpp:60-555 (high-dim).
#include "spz_host.h"

namespace spz {

constexpr int kEncLaneThreads = 128;

template <int W, bool FIRE, int D>
// int8 x 8 fits 5 CTAs per SM without spills (96 registers); the int16 shapes spill there and stay at 4
__global__ void __launch_bounds__(kEncLaneThreads, (W == 8) ? 5 : 4) spz_encode_lane_kernel(const EncParams p)
{
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr int NO = D / 8;                              // octets of dims
    constexpr int ROWB = D * ESZ;                          // bytes per row: 8, 16 or 32
    constexpr int BLKB = 8 * ROWB;                         // bytes per block
    constexpr int NVEC = BLKB / 16;                        // 16-byte vectors per block
    constexpr int NSEC = BLKB / 32;                        // 32-byte sectors per block
    constexpr int RPS = 8 / NSEC;                          // rows per sector: 4, 2 or 1
    constexpr int NCL = (W == 16) ? D / 2 : D / 4;         // clusters per row (pairs / quads of fields)
    constexpr int32_t Bc = 8 * D, G = 16 * D;
    constexpr uint32_t HSLOT = (uint32_t)D * HB / 8;       // header bytes of one slot (3, 4 or 8)
    constexpr bool STRICT = !FIRE;                         // continue-run test: `<` for delta, `<=` for FIRE
    static_assert((W == 8 && D == 8) || (W == 16 && (D == 8 || D == 16)), "shapes covered by the lane encoder");
    constexpr int NW = kEncLaneThreads / 32;

    // [pad to 4 KB][NW rings of 32 words x 32 lanes][NW staging areas of NVEC vectors x 32 lanes]
    extern __shared__ __align__(16) uint8_t dsm[];

    const uint64_t c = (uint64_t)blockIdx.x * kEncLaneThreads + threadIdx.x;
    if (c >= p.nchunks) return;                            // lanes never cooperate: early exit is safe
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dsm);
    const uint32_t apad = (0u - sbase) & 4095u;            // the rings start 4 KB aligned: "row | lane" is an OR
    const uint32_t ringabs = sbase + apad + warp * 4096u + lane * 4u;   // word w of my output ring: ringabs | ((w & 31) << 7)
    uint8_t* const stg = dsm + apad + NW * 4096 + warp * (NVEC * 512) + lane * 16;   // vector v of my block: stg + v * 512

    // ---- this chunk
    const uint64_t first = c * p.chunk_len;
    const int32_t L = (int32_t)((first + p.chunk_len <= p.len) ? p.chunk_len : (p.len - first));
    const uint8_t* gin = p.src + first * ESZ;
    asm volatile("" : "+l"(gin));
    uint8_t* const gout = p.dst + c * p.slot_stride;
    // variant 1 / 2: the non-RLE sibling codecs (delta / plain bit packing, 6-byte metadata, no run slots)
    const uint32_t variant = p.variant;
    const uint32_t meta_bytes = p.write_size ? ((variant == 1u || variant == 2u) ? 6u : 8u) : 0u;   // variant 3: [u48 len][u16 ndims]

    if (L < (int32_t)kMinDataSize) {                       // stored raw (sprintz_delta_rle.cpp:101-109)
        if (meta_bytes == 8u) {
            *reinterpret_cast<uint2*>(gout) = variant == 3u ? make_uint2((uint32_t)L, (uint32_t)D << 16)
                                                            : make_uint2(0u, (uint32_t)L | ((uint32_t)D << 16));
        } else if (meta_bytes == 6u) {
            *reinterpret_cast<uint32_t*>(gout) = (uint32_t)L;
            *reinterpret_cast<uint16_t*>(gout + 4) = (uint16_t)D;
        }
        for (int32_t b = 0; b < L * ESZ; b++) gout[meta_bytes + b] = gin[b];
        p.meta[c].ngroups = 0; p.meta[c].remaining = (uint32_t)L; p.meta[c].status = kStatusOk;
        publish_nbytes(&p.meta[c].nbytes, meta_bytes + (uint32_t)L * ESZ);
        return;
    }

    // ---- bit writer: (lo, hi) accumulator -> lane-private ring of words -> 256-bit stores
    // Invariant between puts: fewer than 32 bits are pending, all of them in `lo`.
    uint32_t lo = 0, pos = 0;                              // pending bits, their count
    uint32_t wcount = 0;                                   // words written to the ring so far
    uint32_t flushed = 0;                                  // words already in HBM (multiple of 8)
    // The address of a ring word is rebuilt from the counter for every store (a fresh register): updating an
    // address register in place right after the STS that reads it stalls on the write-after-read scoreboard.
    auto ring_at = [&](uint32_t w) -> uint32_t { return ((w << 7) & 0xF80u) | ringabs; };
    auto put = [&](uint32_t v, uint32_t nb) {              // v < 2^nb, nb <= 32
        // (carry : t) = v << pos as ONE widening multiply on the FMA pipe (the two funnel shifts were ALU-pipe work)
        uint32_t t, carry;
        asm("{ .reg .b64 q; mul.wide.u32 q, %2, %3; mov.b64 {%0, %1}, q; }" : "=r"(t), "=r"(carry) : "r"(v), "r"(1u << pos));
        // the stored word gets its own register: the STS reads its operands late, and overwriting them right
        // away (the compiler would fold `lo = full ? carry : w` into a predicated move onto w) stalls every put
        uint32_t w;
        asm volatile("or.b32 %0, %1, %2;" : "=r"(w) : "r"(lo), "r"(t));
        pos += nb;
        const bool full = pos >= 32u;
        if (full) { sts32(ring_at(wcount), w); wcount++; }
        lo = full ? carry : (lo | t);
        pos &= 31u;
    };
    // last cluster of a row: append, then zero-pad to the next byte boundary (rows are byte padded).  Up to 7 pad
    // bits on top of a 32-bit cluster can complete a second word; that (rare) case is the extra branch.
    auto put_row_end = [&](uint32_t v, uint32_t nb) {
        uint32_t t, carry;
        asm("{ .reg .b64 q; mul.wide.u32 q, %2, %3; mov.b64 {%0, %1}, q; }" : "=r"(t), "=r"(carry) : "r"(v), "r"(1u << pos));
        uint32_t w;
        asm volatile("or.b32 %0, %1, %2;" : "=r"(w) : "r"(lo), "r"(t));
        pos = (pos + nb + 7u) & ~7u;
        const bool full = pos >= 32u;
        if (full) { sts32(ring_at(wcount), w); wcount++; }
        lo = full ? carry : (lo | t);
        if (pos >= 64u) { sts32(ring_at(wcount), lo); wcount++; lo = 0u; }
        pos &= 31u;
    };
    auto maybe_flush = [&]() {                             // at most 32 new bytes between two calls
        if (wcount - flushed >= 8u) {
            const uint32_t fa = ((flushed << 7) & 0xC00u) | ringabs;          // a multiple of 8 rows: no wrap
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = lds32(fa + j * 128);
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(gout + (size_t)flushed * 4),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            flushed += 8u;
        }
    };
    auto cur_bytes = [&]() -> uint32_t { return wcount * 4u + (pos >> 3); };
    // overwrite already-written bytes [off, off+nbytes) of the stream with the low bytes of (plo, phi)
    auto patch = [&](uint32_t off, uint32_t plo, uint32_t phi, uint32_t nbytes) {
        sts32(ring_at(wcount), lo);                                            // the partial word joins the ring
#pragma unroll
        for (uint32_t b = 0; b < 8; b++) {
            if (b < nbytes) {
                const uint32_t o = off + b;
                const uint32_t val = ((b < 4 ? plo : phi) >> (8u * (b & 3u))) & 0xffu;
                if ((o >> 2) >= flushed) sts8(ring_at(o >> 2) + (o & 3u), val);
                else gout[o] = (uint8_t)val;
            }
        }
        lo = lds32(ring_at(wcount));
    };

    // ---- block input
    auto load_sector = [&](const uint8_t* g, uint32_t (&v)[8]) {
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(g));
    };
    auto store_sector = [&](int s, const uint32_t (&v)[8]) {
        *reinterpret_cast<uint4*>(stg + (2 * s) * 512) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(stg + (2 * s + 1) * 512) = make_uint4(v[4], v[5], v[6], v[7]);
    };
    // The NEXT block's sector s goes straight from HBM / L2 into the staging column with two 16-byte asynchronous copies
    // (LDGSTS, no registers) issued as soon as the current block's sector s has been emitted; nobody waits for them before
    // the next block's pass 1.  (Through registers -- load before a sector's emission, store after it -- the store waited
    // for the load: 28 % of the kernel's stall samples at c3, profiles/r2_ncu_summary.csv.)
    auto async_sector = [&](int s, const uint8_t* g) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(stg + (2 * s) * 512);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a + 512u), "l"(g + 16) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_block = [&](int32_t blk) {                   // plain (latency-exposed) load: first block, zero blocks
        const uint8_t* g = gin + (size_t)blk * BLKB;
#pragma unroll
        for (int s = 0; s < NSEC; s++) { uint32_t v[8]; load_sector(g + s * 32, v); store_sector(s, v); }
    };
    const uint8_t* const gend = p.src + p.len * ESZ - 1;   // last readable input byte (prefetch clamp)

    // ---- predictor state, left-aligned (value << (32-W)); FIRE counters
    uint32_t xprev[D];
    int32_t dprev[D], ct[D];
#pragma unroll
    for (int d = 0; d < D; d++) { xprev[d] = 0; dprev[d] = 0; ct[d] = 0; }

    if (meta_bytes) { put(0u, 32u); put(0u, meta_bytes * 8u - 32u); }   // metadata placeholder (format.h:35-73, written last)

    int32_t i = 0;                                         // next block
    uint32_t run = 0, nslots = 0, ngroups = 0, hdr_off = 0;
    bool outer = true;                                     // next block is gated by the outer-loop test (:134-138)
    int32_t tail = 0;
    load_block(0);

#pragma unroll 1
    for (;;) {
        if (outer && i * Bc > L - G) { tail = i * Bc; break; }
        {   // L2 prefetch of the block pass 2 will fetch (lead time = pass 1); two blocks ahead over-fetched 19 % at c3
            const uint8_t* pf = gin + (size_t)(i + 1) * BLKB;
            if (pf > gend) pf = gend;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            if (BLKB > 128) { const uint8_t* pf2 = pf + 128 > gend ? gend : pf + 128; asm volatile("prefetch.global.L2 [%0];" ::"l"(pf2)); }
        }

        // ================= pass 1: forecast -> zigzag -> masks; zigzagged errors overwrite the block in place
        uint32_t mulp[NCL * (W == 16 ? 1 : 3)];            // cluster multipliers (see pass 2)
        uint32_t cbits[NCL];                               // bits per cluster
        uint32_t hlo = 0, hhi = 0, total = 0;              // this block's header fields (D*HB bits), sum of widths
        // the block's first half of sectors has landed (asynchronous copies complete in order); the second half is awaited
        // where pass 1 first touches it
        asm volatile("cp.async.wait_group %0;" ::"n"(NSEC / 2) : "memory");
#pragma unroll
        for (int oc = 0; oc < NO; oc++) {
            int32_t coef[8], gsh[8];
            uint32_t mk[W == 16 ? 4 : 2];
#pragma unroll
            for (int e = 0; e < 8; e++) { coef[e] = FIRE ? fire_coef<W, false>(ct[oc * 8 + e]) : 0; gsh[e] = 0; }
#pragma unroll
            for (int k = 0; k < (W == 16 ? 4 : 2); k++) mk[k] = 0;

            // one element: left-aligned value in, zigzagged error out LEFT-ALIGNED (top W bits valid, the rest is
            // sign fill that the PRMT packing drops).  The kernel is bound by the half-rate INT32 ALU pipe (ncu:
            // ALU 77 %, FMA 24 %), so shifts-by-constant and the conditional negate are phrased as multiply-adds.
            auto elem = [&](uint32_t vh, int e, int r) -> uint32_t {
                const int d = oc * 8 + e;
                if (!FIRE && variant == 2u) return vh;                         // plain bit packing: the value itself
                const int32_t dh = (int32_t)(vh - xprev[d]);                   // delta, wraps to W bits
                xprev[d] = vh;
                int32_t eh = dh;
                if (FIRE) {
                    // prediction = (prev_delta * coef) >> W (sprintz_xff_rle.cpp:225): prev_delta is left-aligned
                    const int32_t pred = __mulhi(dprev[d], coef[e]);
                    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(eh) : "r"(pred), "n"(-(1 << UP)), "r"(dh));   // dh - (pred << UP): err, wraps to W bits
                }
                const int32_t sg = eh >> 31;
                if (FIRE) {
                    if (r & 1) {                                               // icopysign(err, prev_delta), odd rows
                        int32_t cs;
                        asm("mad.lo.s32 %0, %1, 2, 1;" : "=r"(cs) : "r"(sg));  // +1 / -1
                        if (eh != 0) gsh[e] += dprev[d] * cs;
                    }
                    dprev[d] = dh;
                }
                int32_t e2;
                asm("mad.lo.s32 %0, %1, 2, 0;" : "=r"(e2) : "r"(eh));          // eh << 1 on the FMA pipe
                return (uint32_t)e2 ^ (uint32_t)sg;                            // zigzag (bitpack.h:302), left-aligned
            };

            if (W == 16) {
                // rows are rolled in pairs (the odd/even gradient rule stays static): code size, not speed,
                // limited the fully unrolled version (62 KB of SASS, instruction-cache stalls)
                uint4 q = *reinterpret_cast<const uint4*>(stg + oc * 512);    // row 0; later rows are fetched one ahead
#pragma unroll 1
                for (int rp = 0; rp < 8; rp += 2)
#pragma unroll
                for (int r1 = 0; r1 < 2; r1++) {
                    if (oc == 0 && r1 == 0 && rp == 2) asm volatile("cp.async.wait_all;" ::: "memory");   // row 3 fetches row 4 ahead
                    const int r = r1;                                          // only the parity matters to elem()
                    uint8_t* vp = stg + ((rp + r1) * NO + oc) * 512;
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    if (rp + r1 < 7) q = *reinterpret_cast<const uint4*>(vp + NO * 512);
                    uint32_t z[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t z0 = elem(w[k] << 16, 2 * k, r);
                        const uint32_t z1 = elem(w[k] & 0xffff0000u, 2 * k + 1, r);
                        z[k] = __byte_perm(z0, z1, 0x7632);
                        mk[k] |= z[k];
                    }
                    *reinterpret_cast<uint4*>(vp) = make_uint4(z[0], z[1], z[2], z[3]);
                }
            } else {
#pragma unroll
                for (int v = 0; v < 4; v++) {              // a vector holds rows 2v, 2v+1 (8 bytes each)
                    if (v == 2) asm volatile("cp.async.wait_all;" ::: "memory");     // the block's second sector
                    uint8_t* vp = stg + v * 512;
                    const uint4 q = *reinterpret_cast<const uint4*>(vp);
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    uint32_t z[4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const uint32_t x = w[2 * h + k];
                            const int r = 2 * v + h;
                            const uint32_t z0 = elem(x << 24, 4 * k + 0, r);
                            const uint32_t z1 = elem(__byte_perm(x, 0, 0x1444), 4 * k + 1, r);
                            const uint32_t z2 = elem(__byte_perm(x, 0, 0x2444), 4 * k + 2, r);
                            const uint32_t z3 = elem(x & 0xff000000u, 4 * k + 3, r);
                            const uint32_t a = __byte_perm(z0, z1, 0x0073), b = __byte_perm(z2, z3, 0x0073);
                            z[2 * h + k] = __byte_perm(a, b, 0x5410);
                            mk[k] |= z[2 * h + k];
                        }
                    }
                    *reinterpret_cast<uint4*>(vp) = make_uint4(z[0], z[1], z[2], z[3]);
                }
            }

            // ---- widths of this octet's dims (sprintz_delta_rle.cpp:177-199), counters, row layout
            uint32_t nb[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t m = (W == 16) ? ((e & 1) ? (mk[e >> 1] >> 16) : (mk[e >> 1] & 0xffffu))
                                             : ((mk[e >> 2] >> (8 * (e & 3))) & 0xffu);
                const uint32_t n = nbits_rule<W, false>(m);
                nb[e] = n;
                total += n;
                const uint32_t f = n - (n >> (W == 16 ? 4 : 3));               // W is stored as W-1
                const int fbit = (oc * 8 + e) * HB;
                if (fbit < 32) hlo |= f << fbit; else hhi |= f << (fbit - 32);
                if (FIRE) ct[oc * 8 + e] = sx<CB>(ct[oc * 8 + e] + ((gsh[e] >> UP) >> 2));   // sprintz_xff_rle.cpp:273-275
            }
            if (W == 16) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    // cluster = lo + hi * 2^n_lo, computed from the packed pair p as p + (p >> 16) * (2^n_lo - 2^16)
                    mulp[oc * 4 + k] = (1u << nb[2 * k]) - 65536u;
                    cbits[oc * 4 + k] = nb[2 * k] + nb[2 * k + 1];
                }
            } else {
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    // cluster = b0 + b1 * 2^n0 + b2 * 2^(n0+n1) + b3 * 2^(n0+n1+n2) from the packed quad
                    const uint32_t s1 = nb[4 * k], s2 = s1 + nb[4 * k + 1], s3 = s2 + nb[4 * k + 2];
                    mulp[k * 3 + 0] = (1u << s1) - (1u << 8);
                    mulp[k * 3 + 1] = (1u << s2) - (1u << 16);
                    mulp[k * 3 + 2] = (1u << s3) - (1u << 24);
                    cbits[k] = s3 + nb[4 * k + 3];
                }
            }
        }
        const bool zero = total == 0u && variant == 0u;   // the non-RLE codecs have no run slots

        // ================= control (sprintz_delta_rle.cpp:214-312 restated)
        auto begin_group = [&](uint32_t f_lo, uint32_t f_hi) {
            hdr_off = cur_bytes();
            ngroups++;
            if (HSLOT == 8) { put(f_lo, 32u); put(f_hi, 32u); put(0u, 32u); put(0u, 32u); }
            else { put(f_lo, HSLOT * 8u); put(0u, HSLOT * 8u); }
            maybe_flush();
        };
        auto emit_run = [&](uint32_t n) {                                      // :238-245
            if ((nslots & 1u) == 0u) begin_group(0u, 0u);
            put((n & 0x7fu) | (n > 0x7fu ? 0x80u : 0u), 8u);
            if (n > 0x7fu) put(n >> 7, 8u);
            nslots++;
            maybe_flush();
        };

        bool stop = false, emitted = false;
#pragma unroll 1
        for (;;) {
            if (zero && run < kMaxRun) {
                run++; i++;
                if (STRICT ? (i * Bc < L - G) : (i * Bc <= L - G)) { outer = false; break; }
                // not enough data to finish the group: flush the run, pad the pair (:249-252)
                emit_run(run); run = 0;
                if (nslots & 1u) { put(0u, 8u); nslots++; }
                tail = i * Bc; stop = true; break;
            }
            if (run > 0u) {                                                    // a run just ended (:263-312)
                emit_run(run); run = 0;
                if (zero) continue;                                            // only when the run hit kMaxRun
            }
            // ================= pass 2: data slot
            if ((nslots & 1u) == 0u) begin_group(hlo, hhi);
            else patch(hdr_off + HSLOT, hlo, hhi, HSLOT);
            const bool have_next = (i + 2) * Bc <= L;
            const uint8_t* gnext = gin + (size_t)(have_next ? i + 1 : i) * BLKB;
            // the zigzagged words of a sector are fetched one sector ahead (registers), so the emit never waits on LDS
            uint4 qa = *reinterpret_cast<const uint4*>(stg);
            uint4 qb = *reinterpret_cast<const uint4*>(stg + 512);
#pragma unroll(NSEC > 2 ? 2 : NSEC)
            for (int s = 0; s < NSEC; s++) {
                uint4 na = qa, nb4 = qb;
                if (s + 1 < NSEC) {
                    na = *reinterpret_cast<const uint4*>(stg + (2 * s + 2) * 512);
                    nb4 = *reinterpret_cast<const uint4*>(stg + (2 * s + 3) * 512);
                }
                const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                constexpr int WPR = ROWB / 4;                                  // words per row: 2, 4 or 8
#pragma unroll
                for (int rr = 0; rr < RPS; rr++) {
#pragma unroll
                    for (int k = 0; k < NCL; k++) {
                        const uint32_t pw = w[rr * WPR + k];
                        uint32_t cl;
                        if (W == 16) {
                            cl = pw + (pw >> 16) * mulp[k];
                        } else {
                            cl = pw + __byte_perm(pw, 0, 0x4441) * mulp[k * 3 + 0] + __byte_perm(pw, 0, 0x4442) * mulp[k * 3 + 1]
                                    + (pw >> 24) * mulp[k * 3 + 2];
                        }
                        if (k == NCL - 1) put_row_end(cl, cbits[k]); else put(cl, cbits[k]);
                    }
                    // flush checks every 32 bytes of worst-case output (a row is at most ROWB bytes)
                    if (ROWB == 32 || (ROWB == 16 && (rr & 1)) || rr == RPS - 1) maybe_flush();
                }
                async_sector(s, gnext + s * 32);                               // this sector's words are in registers: its slot is free
                qa = na; qb = nb4;
            }
            nslots++;
            i++;
            outer = (nslots & 1u) == 0u;
            emitted = true;
            break;
        }
        if (stop) break;
        // the next block must be in the staging column: pass 2 fetched it unless this block was part of a run
        if (!emitted && (i + 1) * Bc <= L) load_block(i);
    }

    // ---- finish: remaining ring words, the partial word, the raw tail, the metadata
    asm volatile("cp.async.wait_all;" ::: "memory");
    uint32_t nbytes = cur_bytes();
    {
        const uint32_t nw = wcount - flushed;
        for (uint32_t w = 0; w < nw; w++)
            *reinterpret_cast<uint32_t*>(gout + (size_t)(flushed + w) * 4) = lds32(ring_at(flushed + w));
        for (uint32_t b = 0; b < (pos >> 3); b++) gout[(size_t)(flushed + nw) * 4 + b] = (uint8_t)(lo >> (8u * b));
    }
    const uint32_t remaining = (uint32_t)(L - tail);
    {
        const uint8_t* tsrc = gin + (size_t)tail * ESZ;
        uint8_t* tdst = gout + nbytes;
        const uint32_t nb = remaining * ESZ;
        for (uint32_t b = 0; b < nb; b++) tdst[b] = tsrc[b];
        nbytes += nb;
    }
    if (meta_bytes == 8u) {                                                    // format.h:35-45 | sprintz_xff.cpp:64-69
        *reinterpret_cast<uint2*>(gout) = variant == 3u ? make_uint2((uint32_t)L, (uint32_t)D << 16)
                                                        : make_uint2(ngroups, (remaining & 0xffffu) | ((uint32_t)D << 16));
    } else if (meta_bytes == 6u) {                                             // format.h:65-73
        *reinterpret_cast<uint32_t*>(gout) = (uint32_t)L;
        *reinterpret_cast<uint16_t*>(gout + 4) = (uint16_t)D;
    }
    p.meta[c].ngroups = ngroups; p.meta[c].remaining = remaining; p.meta[c].status = kStatusOk;
    publish_nbytes(&p.meta[c].nbytes, nbytes);
}

// ------------------------------------------------------------------ launch

int get_force_path();

template <int W, bool FIRE, int D>
static cudaError_t launch_enc_lane(const EncParams& p, cudaStream_t stream)
{
    const uint64_t grid64 = (p.nchunks + kEncLaneThreads - 1) / kEncLaneThreads;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    constexpr int NW = kEncLaneThreads / 32, NVEC = 8 * D * (W / 8) / 16;
    const size_t smem = 4096 + NW * 4096 + NW * NVEC * 512;            // alignment slack + rings + staging
    cudaError_t e = cudaFuncSetAttribute(spz_encode_lane_kernel<W, FIRE, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (g_enc_probe) return probe_fill(g_enc_probe, 2, spz_encode_lane_kernel<W, FIRE, D>, kEncLaneThreads, smem, kEncLaneThreads);
    count_launch();
    spz_encode_lane_kernel<W, FIRE, D><<<(uint32_t)grid64, kEncLaneThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape is not covered (caller continues with the team kernels).
cudaError_t launch_encode_lane(int elem_sz, int predictor, EncParams p, cudaStream_t stream)
{
    const int force = get_force_path();
    if (force == 1) return cudaErrorNotSupported;
    if (!((elem_sz == 1 && p.ndims == 8) || (elem_sz == 2 && (p.ndims == 8 || p.ndims == 16)))) return cudaErrorNotSupported;
    if (p.meta == nullptr) return cudaErrorNotSupported;
    // a lane is a whole scalar encoder: only worth it when there are enough chunks to fill the machine
    // measured crossover against the team kernels (gpurun s43): 16 Ki chunks for 8 dims, ~24 Ki for int16 x 16
    if (force != 2 && p.family_pin != 3u && p.nchunks < (p.ndims == 16 ? 24576u : 16384u)) return cudaErrorNotSupported;
    // 256-bit loads / stores: every chunk's input and output slot must start 32-byte aligned
    if (((uintptr_t)p.src & 31) != 0 || ((uintptr_t)p.dst & 31) != 0) return cudaErrorNotSupported;
    if (p.nchunks > 1 && (((p.chunk_len * (uint64_t)elem_sz) & 31) != 0 || (p.slot_stride & 31) != 0)) return cudaErrorNotSupported;
    if (p.chunk_len > 0x04000000ull) return cudaErrorNotSupported;      // 32-bit element / byte cursors
#define SPZ_ELANE(Wv, Dv) return predictor ? launch_enc_lane<Wv, true, Dv>(p, stream) : launch_enc_lane<Wv, false, Dv>(p, stream)
    if (elem_sz == 1) SPZ_ELANE(8, 8);
    if (p.ndims == 8) SPZ_ELANE(16, 8);
    SPZ_ELANE(16, 16);
#undef SPZ_ELANE
}

}  // namespace spz

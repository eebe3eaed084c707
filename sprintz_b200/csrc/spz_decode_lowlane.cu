// spz_decode_lowlane.cu -- lane-per-chunk decode kernels for the LOW-DIM codecs (ndims <= 4 for 8-bit, <= 2 for 16-bit;
// sprintz_delta_lowdim.cpp:397-794, sprintz_xff_lowdim.cpp:413-1119), the univariate / few-variable case.
//
// The low-dim format differs from the high-dim one in three ways: the payload of a block is
// COLUMN-major (for every dim its 8 values' low nbits bits, LSB first = exactly nbits bytes), header fields are a
// flat [block][dim] array, and FIRE's coefficient is counter >> 1 at counter width (no 4-MSB truncation).
// Same decomposition as spz_decode_lane.cu: one lane is a complete scalar decoder of one chunk (all its dims), the
// stream is pulled through a lane-private, bank-private 128-byte ring in shared memory with 256-bit loads issued one
// fill ahead, predictor state lives in registers left-aligned, and a block (8 rows x ndims) leaves as 64/128-bit
// stores.  A column is one 64-bit segment (8-bit) / two 64-bit segments (16-bit) of the ring; every field is
// "mask the low bits, shift the segment by its width".
#include "spz_host.h"

namespace spz {

int get_force_path();

constexpr int kLowLaneThreads = 128;
constexpr uint32_t kLowLaneWarpSmem = 8192;   // 33 rows x 128 B used; the stride keeps "row | lane" an OR

// Q = 0: decode.  Q = 1 / 2: query-on-compressed (REDUCE_MAX / REDUCE_SUM without materialising, see
// spz_decode_lane.cu): rows are folded into per-dim registers instead of being packed and stored.
template <int W, bool FIRE, int D, int Q>
__global__ void __launch_bounds__(kLowLaneThreads, (D * W <= 16) ? 7 : 5) spz_decode_lowlane_kernel(const DecParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr uint32_t FMASK = (1u << HB) - 1u;
    constexpr uint32_t H = (2u * D * HB + 7u) / 8u; // header bytes per group (<= 3)
    constexpr uint32_t B = 8u * D;                  // elements per block
    constexpr int NWORD = 8 * D * ESZ / 4;          // 32-bit words of one decoded block
    static_assert((W == 8 && D >= 1 && D <= 4) || (W == 16 && D >= 1 && D <= 2), "low-dim shapes only");

    __shared__ __align__(16) uint8_t smem[(kLowLaneThreads / 32) * kLowLaneWarpSmem];

    uint64_t c = (uint64_t)blockIdx.x * kLowLaneThreads + threadIdx.x;
    bool live = true;
    if (c >= p.nchunks) {
        if (!Q) return;                             // decoding lanes never cooperate: early exit is safe
        live = false; c = p.nchunks - 1;            // query lanes meet in the final reduction: shadow the last chunk
    }
    const uint32_t rboff = (threadIdx.x >> 5) * kLowLaneWarpSmem + (threadIdx.x & 31) * 4u;

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
    uint8_t* oblk = reinterpret_cast<uint8_t*>(out);                    // where the next block goes
    const bool al32 = ((uintptr_t)out & 31) == 0;                       // whole-sector stores may be 256-bit

    uint32_t pvh[D];                                // running values, left-aligned (value << (32-W))
    int32_t pdh[D], ct[D];                          // previous deltas (left-aligned), FIRE counters
#pragma unroll
    for (int d = 0; d < D; d++) { pvh[d] = 0; pdh[d] = 0; ct[d] = 0; }

    uint32_t wd[NWORD];                             // one decoded block, row-major, packed
#pragma unroll
    for (int k = 0; k < NWORD; k++) wd[k] = 0;
    // element (row r, dim d) of the block <- the current running value of dim d
    // ---- query accumulators (Q != 0): running maximum (left-aligned) | 32-bit partial sums drained into lane-private
    // 64-bit totals (rows 34.. of the warp's shared region; the store staging is unused when nothing is stored)
    uint32_t qa[D];
    uint32_t qblocks = 0;
    if (Q) {
#pragma unroll
        for (int d = 0; d < D; d++) qa[d] = 0;
        if (Q == 2) {
#pragma unroll
            for (int j = 0; j < 2 * D; j++) *reinterpret_cast<uint32_t*>(smem + rboff + (34 + j) * 128) = 0u;
        }
    }
    auto qtot_add = [&](int d, unsigned long long v) {
        uint32_t* lo = reinterpret_cast<uint32_t*>(smem + rboff + (34 + 2 * d) * 128);
        uint32_t* hi = reinterpret_cast<uint32_t*>(smem + rboff + (35 + 2 * d) * 128);
        const unsigned long long t = (((unsigned long long)*hi << 32) | *lo) + v;
        *lo = (uint32_t)t; *hi = (uint32_t)(t >> 32);
    };
    auto qfold1 = [&](int d) {                      // the current running value of dim d, once
        if (Q == 1) qa[d] = pvh[d] > qa[d] ? pvh[d] : qa[d];
        else asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(qa[d]) : "r"(pvh[d]), "n"(1u << W));
    };
    auto qblock_done = [&]() {
        if (Q == 2 && (++qblocks & 4095u) == 0u) {
#pragma unroll
            for (int d = 0; d < D; d++) { qtot_add(d, qa[d]); qa[d] = 0; }
        }
    };
    auto put = [&](int r, int d) {
        if (Q) { qfold1(d); return; }
        const int at = (r * D + d) * ESZ;           // byte offset inside the block
        const int w = at >> 2;
        if (W == 16) wd[w] = __byte_perm(wd[w], pvh[d], (at & 2) ? 0x7610 : 0x3276);
        else wd[w] = __byte_perm(wd[w], pvh[d], (at & 3) == 0 ? 0x3217 : (at & 3) == 1 ? 0x3270 : (at & 3) == 2 ? 0x3710 : 0x7210);
    };
    // Blocks smaller than a 32-byte sector (8 / 16 / 24 bytes) are collected in a lane-private 16-word ring (rows
    // 34..49 of the warp's shared region, bank-private like the input ring) and leave as whole aligned sectors:
    // 8-byte stores per lane cost a partial-sector write each (measured: 1 dim decoded at half the speed of 4 dims).
    uint32_t swi = 0, sri = 0;                      // staged words written / flushed
    auto stage_at = [&](uint32_t w) -> uint32_t { return rboff + (34u + (w & 15u)) * 128u; };
    auto store_block = [&]() {
        if (Q) {
            qblock_done();
        } else if (NWORD == 8) {
            if (al32) {                              // one 256-bit store == one whole sector (sm_100: STG.E.256)
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(oblk), "r"(wd[0]), "r"(wd[1]), "r"(wd[2]),
                             "r"(wd[3]), "r"(wd[4]), "r"(wd[5]), "r"(wd[6]), "r"(wd[7]) : "memory");
            } else {
                reinterpret_cast<uint4*>(oblk)[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                reinterpret_cast<uint4*>(oblk)[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
            }
            oblk += 32;
        } else {
#pragma unroll
            for (int k = 0; k < NWORD; k++) *reinterpret_cast<uint32_t*>(smem + stage_at(swi + k)) = wd[k];
            swi += NWORD;
            if (swi - sri >= 8u) {
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = *reinterpret_cast<const uint32_t*>(smem + stage_at(sri + k));
                if (al32) {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(oblk), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
                } else {
                    reinterpret_cast<uint4*>(oblk)[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<uint4*>(oblk)[1] = make_uint4(v[4], v[5], v[6], v[7]);
                }
                oblk += 32;
                sri += 8u;
            }
        }
    };
    auto store_finish = [&]() {                      // the staged remainder (a multiple of 8 bytes)
        for (; sri != swi; sri += 2u) {
            *reinterpret_cast<uint2*>(oblk) = make_uint2(*reinterpret_cast<const uint32_t*>(smem + stage_at(sri)),
                                                         *reinterpret_cast<const uint32_t*>(smem + stage_at(sri + 1u)));
            oblk += 8;
        }
    };
    // FIRE prediction on the left-aligned previous delta (sprintz_xff_lowdim.cpp:170-173: coef = counter >> 1).
    // 8-bit: |coef| < 2^14, the product cannot wrap: the high word of (prev_delta << 24) * coef is (prev_delta * coef) >> 8.
    // 16-bit: the counter is 32 bits wide and the reference's product wraps at 32 bits: (int32)(pd * coef) >> 16, left-aligned
    // again, is the product with its low half cleared.
    auto predict = [&](int32_t pdh_, int32_t coef) -> int32_t {
        if (W == 8) return (int32_t)((uint32_t)__mulhi(pdh_, coef) << UP);
        return (int32_t)(((uint32_t)(pdh_ >> 16) * (uint32_t)coef) & 0xffff0000u);
    };

    uint32_t pos = mis * 8u;                         // bit cursor relative to base32
    uint32_t ngroups = 0, remaining = 0, status = kStatusOk;
    if (p.has_meta) {                                // format.h:47-62
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
        if ((uint64_t)ngroups * (H + 2u) > cspan) { status = kStatusTruncated; remaining = 0; }
    }
    if (status != kStatusOk) ngroups = 0;

    uint32_t o = 0;                                  // output element cursor of the chunk
#pragma unroll 1
    for (uint32_t g = 0; g < ngroups; g++) {
        if ((pos >> 3) > limit) { status = kStatusTruncated; break; }
        const uint32_t hw = window(pos);             // all 2 * D fields (<= 24 bits)
        pos += H * 8u;
        ensure(pos);                                 // a slot is up to 32 bytes: the header must be its own unit
#pragma unroll 1
        for (int s = 0; s < 2; s++) {
            const uint32_t fw = hw >> (uint32_t)(s * D * HB);
            uint32_t nb[D], tot = 0;
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t f = (fw >> (d * HB)) & FMASK;
                nb[d] = f + (f == (uint32_t)(W - 1) ? 1u : 0u);
                tot += nb[d];
            }
            if (tot == 0u) {
                // ---- run slot (sprintz_delta_lowdim.cpp:566-600, sprintz_xff_lowdim.cpp:640-760)
                const uint32_t rb = window(pos);
                uint32_t nrun = rb & 0x7fu;
                if (rb & 0x80u) { nrun |= ((rb >> 8) & 0xffu) << 7; pos += 16u; } else { pos += 8u; }
                ensure(pos);
                if ((uint64_t)o + (uint64_t)nrun * B > cap) { status = kStatusOverflow; break; }
                if (Q && !FIRE) {                    // a delta run repeats the previous row: fold it in one step
#pragma unroll
                    for (int d = 0; d < D; d++) {
                        if (Q == 1) qfold1(d);
                        else if (nrun) qtot_add(d, (unsigned long long)(pvh[d] >> UP) * (nrun * 8u));
                    }
                } else if (!FIRE) {                  // the previous row, repeated
#pragma unroll
                    for (int k = 0; k < NWORD; k++) wd[k] = 0;
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int d = 0; d < D; d++) put(r, d);
#pragma unroll 1
                    for (uint32_t b = 0; b < nrun; b++) store_block();
                } else {                             // err == 0: the predictor runs free, counters untouched
                    int32_t coef[D];
#pragma unroll
                    for (int d = 0; d < D; d++) coef[d] = fire_coef<W, true>(ct[d]);
#pragma unroll 1
                    for (uint32_t b = 0; b < nrun; b++) {
#pragma unroll
                        for (int r = 0; r < 8; r++) {
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                pdh[d] = predict(pdh[d], coef[d]);
                                pvh[d] += (uint32_t)pdh[d];
                                put(r, d);
                            }
                        }
                        store_block();
                    }
                }
                o += nrun * B;
                continue;
            }

            // ---- data slot: column d = nb[d] bytes holding 8 fields of nb[d] bits
            if (o + B > cap) { status = kStatusOverflow; break; }
            uint32_t colbit = pos;
            int32_t gsh[D];
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t n = nb[d];
                const uint32_t fm = (1u << n) - 1u;
                const int32_t coef = FIRE ? fire_coef<W, true>(ct[d]) : 0;
                gsh[d] = 0;
                uint32_t lo, hi;
                segment(colbit, lo, hi);
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (W == 16 && r == 4) segment(colbit + 4u * n, lo, hi);
                    // zigzag^-1 (bitpack.h:303) in doubled form: m = u ^ -(u & 1) == 2 * err
                    const uint32_t u = lo & fm;
                    const uint32_t b = u & 1u;
                    lo = __funnelshift_r(lo, hi, n); hi >>= n;
                    const int32_t t = 1 - 2 * (int32_t)b;
                    const int32_t mi = (int32_t)(u * (uint32_t)t) - (int32_t)b;
                    if (FIRE) {
                        if (r & 1) { if (mi != 0) gsh[d] += pdh[d] * t; }      // icopysign(err, prev_delta), util.h:63-68
                        pdh[d] = (int32_t)((uint32_t)predict(pdh[d], coef) + ((uint32_t)mi << (UP - 1)));
                        pvh[d] += (uint32_t)pdh[d];
                    } else {
                        pvh[d] += (uint32_t)mi << (UP - 1);
                    }
                    put(r, d);
                }
                if (FIRE) ct[d] = sx<CB>(ct[d] + ((gsh[d] >> UP) >> 2));        // sprintz_xff_lowdim.cpp:222-226
                colbit += n * 8u;
            }
            store_block();
            pos = colbit;
            o += B;
            ensure(pos);
        }
        if (status != kStatusOk) break;
    }

    if (!Q) store_finish();
    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (status == kStatusOk) {
        if ((uint64_t)o + remaining > cap) {
            status = kStatusOverflow;
        } else if ((uint64_t)(pos >> 3) + (uint64_t)remaining * ESZ > (uint64_t)limit) {
            status = kStatusTruncated;              // the raw tail is read straight from the source
        } else {
            const uint8_t* tsrc = base32 + (pos >> 3);
            if (Q) {                                 // raw tail elements go straight to the result (short, rare)
                if (live) {
                    for (uint32_t e = 0; e < remaining; e++) {
                        const unsigned long long v = (W == 16) ? (unsigned long long)(tsrc[2 * e] | (tsrc[2 * e + 1] << 8))
                                                               : (unsigned long long)tsrc[e];
                        if (Q == 1) atomicMax(p.qresult + (o + e) % D, v); else atomicAdd(p.qresult + (o + e) % D, v);
                    }
                }
            } else {
                uint8_t* tdst = reinterpret_cast<uint8_t*>(out + o);
                const uint32_t nbytes = remaining * ESZ;
                for (uint32_t b = 0; b < nbytes; b++) tdst[b] = tsrc[b];
            }
            o += remaining;
        }
    }
    if (Q) {
        // ---- lanes -> warp -> CTA -> one atomic per dim (every lane holds the same dims)
        unsigned long long v[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (Q == 1) v[d] = qa[d] >> UP;
            else {
                const uint32_t lo = *reinterpret_cast<uint32_t*>(smem + rboff + (34 + 2 * d) * 128);
                const uint32_t hi = *reinterpret_cast<uint32_t*>(smem + rboff + (35 + 2 * d) * 128);
                v[d] = (((unsigned long long)hi << 32) | lo) + qa[d];
            }
            if (!live || status != kStatusOk) v[d] = 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long w = __shfl_xor_sync(0xffffffffu, v[d], off);
                v[d] = (Q == 1) ? (w > v[d] ? w : v[d]) : (v[d] + w);
            }
        }
        __syncthreads();                             // everyone is done with the rings: reuse the first bytes
        unsigned long long* cta = reinterpret_cast<unsigned long long*>(smem);
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int d = 0; d < D; d++) cta[(threadIdx.x >> 5) * D + d] = v[d];
        }
        __syncthreads();
        if (threadIdx.x < D) {
            unsigned long long r = cta[threadIdx.x];
#pragma unroll
            for (int w = 1; w < kLowLaneThreads / 32; w++) {
                const unsigned long long x = cta[w * D + threadIdx.x];
                r = (Q == 1) ? (x > r ? x : r) : (r + x);
            }
            if (Q == 1) { if (r) atomicMax(p.qresult + threadIdx.x, r); } else atomicAdd(p.qresult + threadIdx.x, r);
        }
        if (!live) return;
    }
    if (p.meta) {
        p.meta[c].status = status;
        if (p.has_meta) {
            p.meta[c].nbytes = (pos >> 3) - mis + remaining * ESZ;
            p.meta[c].ngroups = ngroups;
            p.meta[c].remaining = o;
        }
    }
}


// ------------------------------------------------------------------ launch

template <int W, bool FIRE, int D>
static cudaError_t launch_lowlane(const DecParams& p, cudaStream_t stream)
{
    const uint64_t grid64 = (p.nchunks + kLowLaneThreads - 1) / kLowLaneThreads;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    count_launch();
    if (p.qop == 1) spz_decode_lowlane_kernel<W, FIRE, D, 1><<<(uint32_t)grid64, kLowLaneThreads, 0, stream>>>(p);
    else if (p.qop == 2) spz_decode_lowlane_kernel<W, FIRE, D, 2><<<(uint32_t)grid64, kLowLaneThreads, 0, stream>>>(p);
    else spz_decode_lowlane_kernel<W, FIRE, D, 0><<<(uint32_t)grid64, kLowLaneThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape is not covered (caller continues with the generic kernel).
cudaError_t launch_decode_lowlane(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    const int force = get_force_path();
    if (force == 1 || p.variant != 0) return cudaErrorNotSupported;
    if (!(elem_sz == 1 ? (p.ndims >= 1 && p.ndims <= 4) : (p.ndims >= 1 && p.ndims <= 2))) return cudaErrorNotSupported;
    if (p.src_end == nullptr || p.src_end <= p.src) return cudaErrorNotSupported;
    // a lane is a whole scalar decoder: only worth it when there are enough chunks to fill the machine
    if (force != 2 && p.nchunks < 16384) return cudaErrorNotSupported;
    // vector stores: every chunk's output must start 16-byte aligned
    if (!p.qop && ((uintptr_t)p.dst & 15) != 0) return cudaErrorNotSupported;
    if (!p.qop && p.nchunks > 1 && ((p.chunk_len * (uint64_t)elem_sz) & 15) != 0) return cudaErrorNotSupported;
    if (p.chunk_len % p.ndims != 0 && p.nchunks > 1) return cudaErrorNotSupported;
    if (p.chunk_len * (uint64_t)elem_sz > 0x0c000000ull) return cudaErrorNotSupported;   // 32-bit bit cursors
#define SPZ_LL(Wv, Dv) return predictor ? launch_lowlane<Wv, true, Dv>(p, stream) : launch_lowlane<Wv, false, Dv>(p, stream)
    if (elem_sz == 1) {
        if (p.ndims == 1) SPZ_LL(8, 1);
        if (p.ndims == 2) SPZ_LL(8, 2);
        if (p.ndims == 3) SPZ_LL(8, 3);
        SPZ_LL(8, 4);
    }
    if (p.ndims == 1) SPZ_LL(16, 1);
    SPZ_LL(16, 2);
#undef SPZ_LL
}

}  // namespace spz

// spz_encode_lowlane.cu -- lane-per-chunk encode kernels for the LOW-DIM codecs (ndims <= 4 for 8-bit, <= 2 for 16-bit;
// sprintz_delta_lowdim.cpp:38-384, sprintz_xff_lowdim.cpp:43-400), the univariate / few-variable case.
//
// One lane encodes one chunk as a plain scalar encoder, like spz_encode_lane.cu, with the low-dim format rules
// (SURVEY 8a row A4): widths = bit length of the OR-mask with only the w-1 -> w promotion, FIRE coefficient =
// counter >> 1 at counter width, header fields a flat [block][dim] array (H = ceil(2 * D * hb / 8) bytes per group, so
// slot 1's fields are OR-ed into bytes that already hold slot 0's), payload COLUMN-major: for every dim its 8
// zigzagged errors' low nbits bits, LSB first = exactly nbits bytes.  Both low-dim codecs use the strict `<`
// continue-run test (sprintz_delta_lowdim.cpp:190, sprintz_xff_lowdim.cpp:234).
// A block is 8..32 bytes: it lives in registers (fetched one block ahead), a column is built from its 8 fields with
// shifts and appended to the bit writer as two (8-bit) or four (16-bit) <= 32-bit pieces.
#include "spz_host.h"

namespace spz {

int get_force_path();

constexpr int kEncLowThreads = 128;

template <int W, bool FIRE, int D>
__global__ void __launch_bounds__(kEncLowThreads, (D * W == 8) ? 8 : (D * W == 16) ? 7 : 5) spz_encode_lowlane_kernel(const EncParams p)
{
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr int UP = 32 - W;
    constexpr int BLKB = 8 * D * ESZ;                      // bytes per block: 8, 16, 24 or 32
    constexpr int NW2 = BLKB / 8;                          // 8-byte words per block
    constexpr int32_t Bc = 8 * D, G = 16 * D;
    constexpr uint32_t H = (2u * D * HB + 7u) / 8u;        // header bytes per group (1..3)
    static_assert((W == 8 && D >= 1 && D <= 4) || (W == 16 && D >= 1 && D <= 2), "low-dim shapes only");
    constexpr int NW = kEncLowThreads / 32;

    // [pad to 4 KB][NW rings of 32 words x 32 lanes]
    extern __shared__ __align__(16) uint8_t dsm[];

    const uint64_t c = (uint64_t)blockIdx.x * kEncLowThreads + threadIdx.x;
    if (c >= p.nchunks) return;                            // lanes never cooperate: early exit is safe
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dsm);
    const uint32_t apad = (0u - sbase) & 4095u;            // the rings start 4 KB aligned: "row | lane" is an OR
    const uint32_t ringabs = sbase + apad + warp * 4096u + lane * 4u;   // word w of my output ring: ringabs | ((w & 31) << 7)

    // ---- this chunk
    const uint64_t first = c * p.chunk_len;
    const int32_t L = (int32_t)((first + p.chunk_len <= p.len) ? p.chunk_len : (p.len - first));
    const uint8_t* gin = p.src + first * ESZ;
    uint8_t* const gout = p.dst + c * p.slot_stride;
    const uint32_t meta_bytes = p.write_size ? 8u : 0u;
    constexpr uint32_t variant = 0u;                       // the non-RLE sibling codecs have no low-dim form

    if (L < (int32_t)kMinDataSize) {                       // stored raw (sprintz_delta_lowdim.cpp:72-80)
        if (meta_bytes) *reinterpret_cast<uint2*>(gout) = make_uint2(0u, (uint32_t)L | ((uint32_t)D << 16));
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


    // ---- block input: 8-byte words, the next block is fetched while the current one is encoded
    auto load_block = [&](int32_t blk, uint2 (&v)[NW2]) {
        const uint2* g = reinterpret_cast<const uint2*>(gin + (size_t)blk * BLKB);
#pragma unroll
        for (int k = 0; k < NW2; k++) v[k] = __ldg(g + k);
    };

    // ---- predictor state, left-aligned (value << (32-W)); FIRE counters
    uint32_t xprev[D];
    int32_t dprev[D], ct[D];
#pragma unroll
    for (int d = 0; d < D; d++) { xprev[d] = 0; dprev[d] = 0; ct[d] = 0; }
    // FIRE prediction on the left-aligned previous delta (sprintz_xff_lowdim.cpp:170-173: coef = counter >> 1); see
    // spz_decode_lowlane.cu for the 16-bit form (the reference's 32-bit product wraps)
    auto predict = [&](int32_t pdh_, int32_t coef) -> int32_t {
        if (W == 8) return (int32_t)((uint32_t)__mulhi(pdh_, coef) << UP);
        return (int32_t)(((uint32_t)(pdh_ >> 16) * (uint32_t)coef) & 0xffff0000u);
    };

    if (meta_bytes) { put(0u, 32u); put(0u, 32u); }        // metadata placeholder (format.h:35-45, written last)

    int32_t i = 0;                                         // next block
    uint32_t run = 0, nslots = 0, ngroups = 0, hdr_off = 0, hdr_val = 0;
    bool outer = true;                                     // next block is gated by the outer-loop test
    int32_t tail = 0;
    uint2 cur[NW2], nxt[NW2];
    load_block(0, cur);
    // Blocks smaller than a sector: a per-block 64-bit load per lane costs one L1 wavefront per lane for part of a
    // sector (ncu: L1 85 % busy, 32 sectors per request).  For int16 x 1 whole sectors are fetched with ONE 256-bit
    // load, parked in a lane-private shared column (bank-private, like the ring) and the blocks are read from there
    // (+13 %).  Measured slower for the int8 shapes (8-byte blocks: 790 -> 736 GB/s, the extra shared traffic and
    // branches cost more than the wavefronts they save), so those keep the plain loads.
    constexpr int BPS = (W == 16 && D == 1) ? 2 : 1;                          // blocks per sector when staging applies
    const bool stage = BPS > 1 && (((uintptr_t)gin & 31) == 0);
    const uint32_t stgabs = sbase + apad + NW * 4096u + warp * 1024u + lane * 4u;   // word k of my sector: stgabs + k * 128
    bool staged_ok = false;
    auto ldg256 = [&](const uint8_t* g, uint32_t (&v)[8]) {
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(g));
    };
    if (stage) {                                           // sector 0 (L >= 128 elements: always inside the chunk)
        uint32_t v[8];
        ldg256(gin, v);
#pragma unroll
        for (int k = 0; k < 8; k++) sts32(stgabs + k * 128, v[k]);
        staged_ok = true;
    }

#pragma unroll 1
    for (;;) {
        if (outer && i * Bc > L - G) { tail = i * Bc; break; }
        const bool have_next = (i + 2) * Bc <= L;
        uint32_t sv[8];
        bool newsec = false;
        if (stage) {
            if (have_next) {
                const int32_t nb_ = i + 1;
                if ((nb_ & (BPS - 1)) == 0) {              // first block of a sector: fetch it if it lies inside the chunk
                    if ((nb_ + BPS) * Bc <= L) { ldg256(gin + (size_t)nb_ * BLKB, sv); newsec = true; }
                    else staged_ok = false;                // ragged end: plain per-block loads from here on
                }
                if (!newsec && !staged_ok) load_block(nb_, nxt);
            }
        } else {
            load_block(have_next ? i + 1 : i, nxt);
        }

        // ================= pass 1: forecast -> zigzag -> masks (registers)
        uint32_t zz[D][8], nb[D];
        uint32_t fields = 0, total = 0;
        {
            uint32_t mk[D];
            int32_t coef[D], gsh[D];
#pragma unroll
            for (int d = 0; d < D; d++) { mk[d] = 0; gsh[d] = 0; coef[d] = FIRE ? fire_coef<W, true>(ct[d]) : 0; }
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const int at = (r * D + d) * ESZ;                          // byte offset inside the block
                    const uint32_t w32 = (at & 4) ? cur[at >> 3].y : cur[at >> 3].x;
                    const uint32_t vh = (w32 >> (8 * (at & 3))) << UP;         // left-aligned value
                    const int32_t dh = (int32_t)(vh - xprev[d]);               // delta, wraps to W bits
                    xprev[d] = vh;
                    int32_t eh = dh;
                    if (FIRE) {
                        eh = dh - predict(dprev[d], coef[d]);                  // err, wraps to W bits
                        if (r & 1) {                                           // icopysign(err, prev_delta), odd rows
                            if (eh != 0) gsh[d] += (eh < 0) ? -dprev[d] : dprev[d];
                        }
                        dprev[d] = dh;
                    }
                    const uint32_t z = ((uint32_t)eh << 1) ^ (uint32_t)(eh >> 31);   // zigzag (bitpack.h:302), left-aligned
                    mk[d] |= z;
                    zz[d][r] = z >> UP;
                }
            }
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t n = nbits_rule<W, true>(mk[d] >> UP);           // sprintz_delta_lowdim.cpp:172-173
                nb[d] = n;
                total += n;
                fields |= (n - (n >> (W == 16 ? 4 : 3))) << (d * HB);          // W is stored as W-1
                if (FIRE) ct[d] = sx<CB>(ct[d] + ((gsh[d] >> UP) >> 2));       // sprintz_xff_lowdim.cpp:222-226
            }
        }
        const bool zero = total == 0u;

        // ================= control (sprintz_delta_lowdim.cpp:176-283 restated, as in spz_encode_lane.cu)
        auto begin_group = [&](uint32_t f0) {
            hdr_off = cur_bytes();
            hdr_val = f0;
            ngroups++;
            put(f0, H * 8u);
            maybe_flush();
        };
        auto emit_run = [&](uint32_t n) {
            if ((nslots & 1u) == 0u) begin_group(0u);
            put((n & 0x7fu) | (n > 0x7fu ? 0x80u : 0u), 8u);
            if (n > 0x7fu) put(n >> 7, 8u);
            nslots++;
            maybe_flush();
        };

        bool stop = false;
#pragma unroll 1
        for (;;) {
            if (zero && run < kMaxRun) {
                run++; i++;
                if (i * Bc < L - G) { outer = false; break; }
                // not enough data to finish the group: flush the run, pad the pair
                emit_run(run); run = 0;
                if (nslots & 1u) { put(0u, 8u); nslots++; }
                tail = i * Bc; stop = true; break;
            }
            if (run > 0u) {                                                    // a run just ended
                emit_run(run); run = 0;
                if (zero) continue;                                            // only when the run hit kMaxRun
            }
            // ================= pass 2: data slot, column-major
            if ((nslots & 1u) == 0u) begin_group(fields);
            else patch(hdr_off, hdr_val | (fields << (D * HB)), 0u, H);
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t n = nb[d];
                if (n != 0u) {
                    if (W == 8) {
                        const uint32_t a = zz[d][0] | (zz[d][1] << n) | (zz[d][2] << (2u * n)) | (zz[d][3] << (3u * n));
                        const uint32_t b = zz[d][4] | (zz[d][5] << n) | (zz[d][6] << (2u * n)) | (zz[d][7] << (3u * n));
                        put(a, 4u * n); put(b, 4u * n);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) put(zz[d][2 * k] | (zz[d][2 * k + 1] << n), 2u * n);
                    }
                }
                if (W == 16 && D == 2 && d == 0) maybe_flush();
            }
            maybe_flush();
            nslots++;
            i++;
            outer = (nslots & 1u) == 0u;
            break;
        }
        if (stop) break;
#pragma unroll
        for (int k = 0; k < NW2; k++) cur[k] = nxt[k];
        if (stage && have_next) {
            if (newsec) {                                  // every block of the previous sector has been consumed
#pragma unroll
                for (int k = 0; k < 8; k++) sts32(stgabs + k * 128, sv[k]);
                staged_ok = true;
            }
            if (staged_ok) {                               // block i (already incremented) from the parked sector
                const uint32_t a = stgabs + (uint32_t)(i & (BPS - 1)) * (uint32_t)(BLKB / 4) * 128u;
#pragma unroll
                for (int k = 0; k < NW2; k++) { cur[k].x = lds32(a + (2 * k) * 128); cur[k].y = lds32(a + (2 * k + 1) * 128); }
            }
        }
    }

    // ---- finish: remaining ring words, the partial word, the raw tail, the metadata
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

template <int W, bool FIRE, int D>
static cudaError_t launch_enc_lowlane(const EncParams& p, cudaStream_t stream)
{
    const uint64_t grid64 = (p.nchunks + kEncLowThreads - 1) / kEncLowThreads;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    const size_t smem = 4096 + (kEncLowThreads / 32) * (4096 + 1024);          // alignment slack + rings + sector columns
    if (g_enc_probe) return probe_fill(g_enc_probe, 2, spz_encode_lowlane_kernel<W, FIRE, D>, kEncLowThreads, smem, kEncLowThreads);
    count_launch();
    spz_encode_lowlane_kernel<W, FIRE, D><<<(uint32_t)grid64, kEncLowThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// Returns cudaErrorNotSupported when the shape is not covered (caller continues with the generic kernel).
cudaError_t launch_encode_lowlane(int elem_sz, int predictor, EncParams p, cudaStream_t stream)
{
    const int force = get_force_path();
    if (force == 1 || p.variant != 0) return cudaErrorNotSupported;
    if (!(elem_sz == 1 ? (p.ndims >= 1 && p.ndims <= 4) : (p.ndims >= 1 && p.ndims <= 2))) return cudaErrorNotSupported;
    if (p.meta == nullptr) return cudaErrorNotSupported;
    // a lane is a whole scalar encoder: only worth it when there are enough chunks to fill the machine
    if (force != 2 && p.family_pin != 3u && p.nchunks < 16384) return cudaErrorNotSupported;
    // 64-bit loads / 256-bit stores: chunk inputs 8-byte aligned, output slots 32-byte aligned
    if (((uintptr_t)p.src & 7) != 0 || ((uintptr_t)p.dst & 31) != 0) return cudaErrorNotSupported;
    if (p.nchunks > 1 && (((p.chunk_len * (uint64_t)elem_sz) & 7) != 0 || (p.slot_stride & 31) != 0)) return cudaErrorNotSupported;
    if (p.chunk_len > 0x04000000ull) return cudaErrorNotSupported;            // 32-bit element / byte cursors
#define SPZ_ELL(Wv, Dv) return predictor ? launch_enc_lowlane<Wv, true, Dv>(p, stream) : launch_enc_lowlane<Wv, false, Dv>(p, stream)
    if (elem_sz == 1) {
        if (p.ndims == 1) SPZ_ELL(8, 1);
        if (p.ndims == 2) SPZ_ELL(8, 2);
        if (p.ndims == 3) SPZ_ELL(8, 3);
        SPZ_ELL(8, 4);
    }
    if (p.ndims == 1) SPZ_ELL(16, 1);
    SPZ_ELL(16, 2);
#undef SPZ_ELL
}

}  // namespace spz

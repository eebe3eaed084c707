// spz_encode.cu -- Sprintz encode kernels for sm_100a.
//
// One *team* of lanes (a power-of-two slice of a warp, >= ndims lanes when ndims <= 32, else the whole
// warp with several dims per lane) encodes one chunk == one complete reference stream, byte-identical
// to what sprintz_compress_{delta,xff}_{8b,16b} (sprintz.h:16-32) emits for that slice:
//
//   phase 1  lane-per-dim: forecast (delta | FIRE) -> zigzag -> OR-mask -> field width
//            (replaces the scalar per-dim loop sprintz_delta_rle.cpp:161-201 / sprintz_xff_rle.cpp:197-298)
//   control  team-uniform run-length / group state machine (sprintz_delta_rle.cpp:214-312)
//   phase 2  lane-per-row: bit-emit of the 8 rows + the header nibbles into a shared-memory staging
//            buffer (replaces the pext stripe loop sprintz_delta_rle.cpp:315-381)
//   flush    128-bit coalesced copies of finished groups to HBM
//
// No tensor cores: this is integer bit-twiddling, bounded by instruction issue and HBM.
#include "spz_host.h"

namespace spz {

namespace {

// Serial LSB-first bit packer used by one lane for one row (or the header). Writes to shared memory
// at an arbitrary byte address: the first and last words go out as bytes, the middle as aligned words.
struct BitPacker {
    uint8_t* wp;        // aligned word pointer
    uint64_t acc;
    uint32_t nacc;      // bits held in acc
    uint32_t a;         // bytes of the first word that are not ours (0..3); 0 once the first word is out
    __device__ __forceinline__ void init(uint8_t* p, uint32_t init_bits, uint32_t init_nbits)
    {
        uintptr_t ad = (uintptr_t)p;
        a = (uint32_t)(ad & 3);
        wp = p - a;
        acc = (uint64_t)init_bits << (8 * a);
        nacc = 8 * a + init_nbits;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n)
    {
        acc |= (uint64_t)v << nacc;
        nacc += n;
        if (nacc >= 32) {
            uint32_t wv = (uint32_t)acc;
            if (a) {
                for (uint32_t b = a; b < 4; b++) wp[b] = (uint8_t)(wv >> (8 * b));
                a = 0;
            } else {
                *reinterpret_cast<uint32_t*>(wp) = wv;
            }
            wp += 4;
            acc >>= 32;
            nacc -= 32;
        }
    }
    __device__ __forceinline__ void finish()
    {
        const uint32_t nb = (nacc + 7) >> 3;
        for (uint32_t b = a; b < nb; b++) wp[b] = (uint8_t)(acc >> (8 * b));
    }
};

}  // namespace

// R = rows per block: 8 is the reference's format (sprintz_delta_rle.cpp:62); 16 / 32 / 64 are BASELINE config 5's
// block-size axis with the reference's derived constants (min_data_size = 8 * block_sz * 2, the FIRE counter update
// shifts by log2_block_sz - 1: sprintz_xff_rle.cpp:78-82,271-275) and always the high-dim layout.
template <int W, bool FIRE, bool LOW, int R>
__global__ void __launch_bounds__(kThreads) spz_encode_kernel(const EncParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    static_assert(R == 8 || !LOW, "the low-dim (column-major) layout exists for 8-row blocks only");
    constexpr int GSH = (R == 8) ? 2 : (R == 16) ? 3 : (R == 32) ? 4 : 5;      // log2(R) - 1
    constexpr int64_t kMinData = 16 * R;

    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int TEAM = (int)p.team;
    const int DPL = (int)p.dpl;
    const int D = (int)p.ndims;
    const int t = lane & (TEAM - 1);
    const int team_in_cta = (int)(threadIdx.x / TEAM);
    const uint64_t c = (uint64_t)blockIdx.x * p.teams_per_cta + team_in_cta;
    if (c >= p.nchunks) return;      // blockDim.x == teams_per_cta * TEAM, so team_in_cta < teams_per_cta
    const uint32_t tmask = team_mask_of(lane, TEAM);
    const int Dp = TEAM * DPL;

    // ---- shared memory carve-up (per team); offsets computed on the host (enc_layout)
    uint8_t* sm = smem + (size_t)team_in_cta * p.team_smem;
    int32_t* s_pv = reinterpret_cast<int32_t*>(sm);                 // [Dp] previous value
    int32_t* s_pd = s_pv + Dp;                                      // [Dp] previous delta   (FIRE)
    int32_t* s_ct = s_pd + Dp;                                      // [Dp] FIRE counter     (FIRE)
    uint8_t* s_nb = sm + p.nb_off;                                  // [Dp] field widths of the current block
    uint8_t* s_tile = sm + p.tile_off;                              // high-dim: [8][tile_stride] zigzagged errors
                                                                    // low-dim : [Dp] packed columns (16 B each)
    const uint32_t tile_stride = p.tile_stride;
    uint8_t* s_out = sm + p.out_off;

    for (int k = 0; k < DPL; k++) {
        s_pv[k * TEAM + t] = 0;
        if (FIRE) { s_pd[k * TEAM + t] = 0; s_ct[k * TEAM + t] = 0; }
    }

    // ---- this chunk
    const uint64_t first = c * p.chunk_len;
    const int64_t L = (int64_t)((first + p.chunk_len <= p.len) ? p.chunk_len : (p.len - first));
    const uint_t* __restrict__ src = reinterpret_cast<const uint_t*>(p.src) + first;
    uint8_t* const gstream = p.dst + c * p.slot_stride;
    const int64_t B = (int64_t)R * D, G = 2 * B;
    const uint32_t H = (2u * (uint32_t)D * HB + 7u) >> 3;
    // variant 1 / 2: the non-RLE sibling codecs compress_rowmajor_delta_* / compress_rowmajor_*
    // (sprintz_delta.cpp:776-1017, :64-392): 6-byte metadata [u32 len][u16 ndims] (format.h:65-73), every block
    // is a data slot (possibly of zero bytes), no zigzag/delta for variant 2
    const uint32_t variant = p.variant;
    // variant 3: compress8b_rowmajor_xff (sprintz_xff.cpp:35-303): FIRE without run slots, 8-byte metadata [u48 len][u16 ndims]
    const uint32_t meta_bytes = p.write_size ? ((variant == 1u || variant == 2u) ? 6u : 8u) : 0u;

    OutStage o;
    o.init(s_out, gstream);
    if (t == 0) for (uint32_t b = 0; b < meta_bytes; b++) s_out[o.cur + b] = 0;      // metadata placeholder
    o.cur += meta_bytes;
    uint64_t stream_bytes = meta_bytes;     // bytes emitted so far (team-uniform)
    uint32_t ngroups = 0, nslots = 0, run = 0;
    uint32_t hdr_off = 0;                   // staging offset of the current group's header
    int64_t i = 0, tail = 0;
    bool outer = true;

    // run slot: zero header fields + 1..2 length bytes (sprintz_delta_rle.cpp:238-245)
    auto begin_slot = [&]() {
        if ((nslots & 1u) == 0) {
            hdr_off = o.cur;
            for (uint32_t b = t; b < H; b += TEAM) o.buf[hdr_off + b] = 0;
            o.cur += H; stream_bytes += H; ngroups++;
            __syncwarp(tmask);
        }
    };
    auto end_slot = [&]() {
        nslots++;
        __syncwarp(tmask);          // phase-2 readers of s_tile/s_nb are done before the next phase 1
        if ((nslots & 1u) == 0) stage_flush(o, false, t, TEAM, tmask);   // group complete -> HBM
    };
    auto emit_run = [&](uint32_t n) {
        begin_slot();
        const uint32_t two = n > 0x7f ? 1u : 0u;
        if (t == 0) {
            o.buf[o.cur] = (uint8_t)((n & 0x7f) | (two << 7));
            if (two) o.buf[o.cur + 1] = (uint8_t)(n >> 7);
        }
        o.cur += 1 + two; stream_bytes += 1 + two;
        end_slot();
    };

    if (L >= kMinData) {
        const bool strict = !(FIRE && !LOW);   // `<` vs `<=` in the continue-run test (sprintz_xff_rle.cpp:362)
        for (;;) {
            if (outer && i * B > L - G) { tail = i * B; break; }

            // ================= phase 1: lane-per-dim forecasting of block i
            const uint_t* __restrict__ blk = src + i * B;
            uint32_t total = 0;       // sum of field widths over all dims (team-uniform after the scan)
            uint32_t my_off[1];       // only used by LOW (DPL == 1): byte offset of my column
            my_off[0] = 0;
            for (int k = 0; k < DPL; k++) {
                const int d = k * TEAM + t;
                uint32_t n = 0;
                uint32_t zz[kBlockRows];
                if (d < D) {
                    uint32_t pv = (uint32_t)s_pv[d];
                    int32_t pd = 0, coef = 0, gs = 0;
                    if (FIRE) { pd = s_pd[d]; coef = fire_coef<W, LOW>(s_ct[d]); }
                    uint32_t mask = 0;
#pragma unroll 1
                    for (int sb = 0; sb < R / kBlockRows; sb++) {       // 8 rows at a time (R == 8: once)
                        uint32_t x[kBlockRows];
#pragma unroll
                        for (int r = 0; r < kBlockRows; r++) x[r] = blk[(sb * kBlockRows + r) * D + d];
#pragma unroll
                        for (int r = 0; r < kBlockRows; r++) {
                            const int32_t delta = sx<W>((int32_t)(x[r] - pv));
                            int32_t err = delta;
                            if (FIRE) {
                                err = sx<W>(delta - fire_predict<W>(pd, coef));
                                if (r & 1) gs += fire_grad<W>(err, pd);
                            }
                            zz[r] = (variant == 2u) ? x[r] : zigzag_enc<W>(err);
                            mask |= zz[r];
                            pv = x[r]; pd = delta;
                        }
                        if (!LOW) {
#pragma unroll
                            for (int r = 0; r < kBlockRows; r++)
                                reinterpret_cast<uint_t*>(s_tile + (sb * kBlockRows + r) * tile_stride)[d] = (uint_t)zz[r];
                        }
                    }
                    s_pv[d] = (int32_t)pv;
                    if (FIRE) {
                        s_pd[d] = pd;
                        s_ct[d] = sx<CB>(s_ct[d] + (sx<W>(gs) >> GSH));   // sprintz_xff_rle.cpp:271-275
                    }
                    n = nbits_rule<W, LOW>(mask);
                    s_nb[d] = (uint8_t)n;
                }
                const uint32_t incl = team_incl_scan(n, t, TEAM, tmask);
                if (LOW) {
                    // low-dim: column-major payload, dim d's 8 fields = n bytes at byte offset sum_{d'<d} n
                    // (sprintz_delta_lowdim.cpp:306-356). Packed right here by the owning lane.
                    my_off[0] = incl - n;
                    if (d < D && n > 0) {
                        // 8 fields of n bits -> 8n bits; build as two 4-field halves (<= 64 bits each)
                        uint64_t h0 = 0, h1 = 0;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            h0 |= (uint64_t)zz[r] << (r * n);
                            h1 |= (uint64_t)zz[r + 4] << (r * n);
                        }
                        // stash for emission after the control decision (n <= 16 -> 4n <= 64)
                        const uint32_t s4 = 4 * n;
                        uint64_t lo = h0 | (s4 < 64 ? (h1 << s4) : 0ull);
                        uint64_t hi = s4 < 64 ? (s4 ? (h1 >> (64 - s4)) : 0ull) : h1;
                        // keep in registers via the tile-less path: write into the s_pd-independent scratch
                        // area reserved for LOW right after s_nb (16 bytes per dim)
                        uint64_t* col = reinterpret_cast<uint64_t*>(s_tile) + 2 * d;
                        col[0] = lo; col[1] = hi;
                    }
                }
                total += team_last(incl, TEAM, tmask);
            }
            __syncwarp(tmask);
            const bool zero = (total == 0) && variant == 0u;      // the non-RLE codecs have no run slots

            // ================= control: run-length / group state machine (team-uniform)
            bool emit_block = true;
            for (;;) {   // the reference's `goto do_rle` retry when a run hits kMaxRun (sprintz_delta_rle.cpp:305-308)
                if (zero && run < kMaxRun) {
                    run++; i++;
                    emit_block = false;
                    if (strict ? (i * B < L - G) : (i * B <= L - G)) { outer = false; break; }
                    // not enough data left to finish the group: flush the run, pad the pair (:228-260)
                    emit_run(run); run = 0;
                    if (nslots & 1u) {
                        if (t == 0) o.buf[o.cur] = 0;
                        o.cur += 1; stream_bytes += 1;
                        end_slot();
                    }
                    tail = i * B;
                    goto stream_done;
                }
                if (run > 0) {
                    emit_run(run); run = 0;
                    if (zero) continue;
                }
                break;
            }
            if (!emit_block) continue;

            // ================= phase 2: emit a data slot
            begin_slot();
            {
                const uint32_t s = nslots & 1u;
                const uint32_t pay = o.cur;                         // staging offset of the payload
                const uint32_t row_bytes = (total + 7u) >> 3;
                const uint32_t pay_bytes = LOW ? total : row_bytes * R;
                if (LOW) {
                    // each lane copies its packed column (n bytes) to its byte offset
                    const int d = t;
                    if (d < D) {
                        const uint32_t n = s_nb[d];
                        const uint64_t* col = reinterpret_cast<const uint64_t*>(s_tile) + 2 * d;
                        const uint64_t lo = n ? col[0] : 0ull, hi = n ? col[1] : 0ull;
                        uint8_t* dstb = o.buf + pay + my_off[0];
                        for (uint32_t b = 0; b < n; b++)
                            dstb[b] = (uint8_t)(b < 8 ? (lo >> (8 * b)) : (hi >> (8 * (b - 8))));
                    }
                    if (t == 0) {   // header nibbles: few dims, one lane
                        const uint32_t hbit = s * (uint32_t)D * HB;
                        uint8_t* hp = o.buf + hdr_off + (hbit >> 3);
                        BitPacker bp; bp.init(hp, (uint32_t)hp[0] & ((1u << (hbit & 7)) - 1u), hbit & 7);
                        for (int dd = 0; dd < D; dd++) {
                            const uint32_t n = s_nb[dd];
                            bp.put(n - (n == (uint32_t)W), HB);
                        }
                        bp.finish();
                    }
                } else {
                    // jobs 0..R-1 = the payload rows, job R = this slot's header fields
                    for (int j = t; j < R + 1; j += TEAM) {
                        BitPacker bp;
                        if (j < R) {
                            bp.init(o.buf + pay + (uint32_t)j * row_bytes, 0, 0);
                            const uint_t* trow = reinterpret_cast<const uint_t*>(s_tile + j * tile_stride);
                            for (int dd = 0; dd < D; dd++) bp.put(trow[dd], s_nb[dd]);
                        } else {
                            const uint32_t hbit = s * (uint32_t)D * HB;
                            uint8_t* hp = o.buf + hdr_off + (hbit >> 3);
                            bp.init(hp, (uint32_t)hp[0] & ((1u << (hbit & 7)) - 1u), hbit & 7);
                            for (int dd = 0; dd < D; dd++) {
                                const uint32_t n = s_nb[dd];
                                bp.put(n - (n == (uint32_t)W), HB);     // w stored as w-1 (sprintz_delta_rle.cpp:199)
                            }
                        }
                        bp.finish();
                    }
                }
                o.cur += pay_bytes; stream_bytes += pay_bytes;
            }
            end_slot();
            i++;
            outer = ((nslots & 1u) == 0);
        }
    }
stream_done:
    __syncwarp(tmask);
    {
        // ---- trailing raw elements (sprintz_delta_rle.cpp:398-403); also the whole input when L < 128
        const int64_t rem_elems = L - tail;
        // low-dim truncates the tail length itself to u16 (sprintz_delta_lowdim.cpp:376-381)
        const uint64_t copy_elems = LOW ? (uint64_t)(rem_elems & 0xffff) : (uint64_t)rem_elems;
        const uint8_t* tsrc = reinterpret_cast<const uint8_t*>(src + tail);
        uint64_t left = copy_elems * ESZ, done = 0;
        while (left > 0) {
            uint32_t space = p.out_cap - o.cur;
            uint32_t take = left < space ? (uint32_t)left : space;
            for (uint32_t b = t; b < take; b += TEAM) o.buf[o.cur + b] = tsrc[done + b];
            o.cur += take; done += take; left -= take; stream_bytes += take;
            if (left > 0) stage_flush(o, false, t, TEAM, tmask);
        }
        stage_flush(o, true, t, TEAM, tmask);
        __syncwarp(tmask);
        if (t == 0) {
            if (p.write_size && variant) {      // format.h:65-73; sprintz_xff.cpp:64-69 for variant 3
                const uint32_t L32 = (uint32_t)L;
                gstream[0] = (uint8_t)L32; gstream[1] = (uint8_t)(L32 >> 8);
                gstream[2] = (uint8_t)(L32 >> 16); gstream[3] = (uint8_t)(L32 >> 24);
                const uint32_t at = variant == 3u ? 6u : 4u;
                if (variant == 3u) { gstream[4] = 0; gstream[5] = 0; }
                gstream[at] = (uint8_t)D; gstream[at + 1] = (uint8_t)(D >> 8);
            } else if (p.write_size) {     // format.h:35-45, written last because ngroups is only known now
                const uint32_t rem16 = (uint32_t)rem_elems & 0xffffu;
                gstream[0] = (uint8_t)ngroups; gstream[1] = (uint8_t)(ngroups >> 8);
                gstream[2] = (uint8_t)(ngroups >> 16); gstream[3] = (uint8_t)(ngroups >> 24);
                gstream[4] = (uint8_t)rem16; gstream[5] = (uint8_t)(rem16 >> 8);
                gstream[6] = (uint8_t)D; gstream[7] = (uint8_t)(D >> 8);
            }
            ChunkMeta m;
            m.nbytes = (uint32_t)stream_bytes;
            m.ngroups = ngroups;
            m.remaining = (uint32_t)rem_elems;
            m.status = (stream_bytes >> 32) ? kStatusOverflow : kStatusOk;
            p.meta[c] = m;
        }
    }
}

// ------------------------------------------------------------------ launch

template <int W, bool FIRE, bool LOW, int R = 8>
static cudaError_t launch_one(const EncParams& p, uint32_t grid, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(spz_encode_kernel<W, FIRE, LOW, R>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (g_enc_probe) return probe_fill(g_enc_probe, 0, spz_encode_kernel<W, FIRE, LOW, R>, (int)(p.teams_per_cta * p.team), smem, p.teams_per_cta);
    count_launch();
    spz_encode_kernel<W, FIRE, LOW, R><<<grid, p.teams_per_cta * p.team, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline uint32_t pow2_ceil(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

// Shared-memory layout of one team (all regions 16-byte aligned).
static void enc_layout(int elem_sz, bool low, EncParams& p, uint32_t R = 8)
{
    const uint32_t Dp = p.team * p.dpl;
    const uint32_t hb = elem_sz == 1 ? 3 : 4;
    const uint32_t H = (2 * p.ndims * hb + 7) / 8;
    uint32_t off = (3u * 4u * Dp + 15u) & ~15u;         // pv, pd, ct
    p.nb_off = off;
    off += (Dp + 15u) & ~15u;
    p.tile_off = off;
    p.tile_stride = low ? 0u : ((((Dp * (uint32_t)elem_sz + 3u) >> 2) | 1u) * 4u);   // odd #words: conflict-free rows
    off += low ? 16u * Dp : ((R * p.tile_stride + 15u) & ~15u);
    p.out_off = off;
    uint32_t cap = 32 + H + 2 * (R * p.ndims * (uint32_t)elem_sz) + 16;   // skip + one full group + slack
    cap = (cap + 15u) & ~15u;
    if (cap < 256) cap = 256;
    p.out_cap = cap;
    p.team_smem = off + cap;
}

cudaError_t launch_encode_hd(int elem_sz, int predictor, EncParams p, cudaStream_t stream);
cudaError_t launch_encode_lane(int elem_sz, int predictor, EncParams p, cudaStream_t stream);
cudaError_t launch_encode_lowlane(int elem_sz, int predictor, EncParams p, cudaStream_t stream);
cudaError_t launch_encode_wide(int elem_sz, int predictor, EncParams p, cudaStream_t stream);

thread_local EncodeProbe* g_enc_probe = nullptr;      // set by probe_encode(): launchers describe instead of launching
static thread_local int g_last_encode_family = 0;    // 0 generic team kernel, 1 fast team kernel (hd), 2 lane kernel, 3 wide-row kernel
int last_encode_family() { return g_last_encode_family; }

// What launch_encode() would do for this call (kernel family, chunks per machine round); launches nothing.
EncodeProbe probe_encode(int elem_sz, int predictor, EncParams p)
{
    EncodeProbe pr{0, 0};
    const int keep = g_last_encode_family;
    g_enc_probe = &pr;
    const cudaError_t e = launch_encode(elem_sz, predictor, p, nullptr);
    g_enc_probe = nullptr;
    g_last_encode_family = keep;
    if (e != cudaSuccess) { (void)cudaGetLastError(); pr.family = -1; pr.round_chunks = 0; }
    return pr;
}

cudaError_t launch_encode(int elem_sz, int predictor, EncParams p, cudaStream_t stream)
{
    if (predictor >= 2) { p.variant = (uint32_t)predictor - 1u; }                  // SPZ_RM_DELTA / SPZ_RM_PLAIN / SPZ_RM_XFF (frames)
    if (p.variant) predictor = (p.variant == 3u) ? 1 : 0;
    const uint32_t pin = p.family_pin;           // family + 1 chosen for the whole frame, or 0
    const uint32_t R = p.block_rows ? p.block_rows : 8u;
    if (R != 8u) {
        // block-size axis (16 / 32 / 64 rows per block): generic kernels, high-dim layout for every ndims, RLE codecs only
        if (p.variant != 0 || (R != 16u && R != 32u && R != 64u)) return cudaErrorInvalidValue;
        g_last_encode_family = 0;
        p.team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
        p.dpl = (p.ndims + p.team - 1) / p.team;
        enc_layout(elem_sz, false, p, R);
        uint32_t tpc = kThreads / p.team;
        while (tpc > 1 && (size_t)tpc * p.team_smem > 200 * 1024) tpc >>= 1;
        p.teams_per_cta = tpc;
        const size_t smem = (size_t)tpc * p.team_smem;
        if (smem > 227 * 1024) return cudaErrorInvalidValue;
        const uint64_t grid64 = (p.nchunks + tpc - 1) / tpc;
        if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
        const uint32_t grid = (uint32_t)grid64;
#define SPZ_BS(Rv) \
        if (R == Rv) { \
            if (elem_sz == 1) return predictor ? launch_one<8, true, false, Rv>(p, grid, smem, stream) : launch_one<8, false, false, Rv>(p, grid, smem, stream); \
            return predictor ? launch_one<16, true, false, Rv>(p, grid, smem, stream) : launch_one<16, false, false, Rv>(p, grid, smem, stream); \
        }
        SPZ_BS(16) SPZ_BS(32) SPZ_BS(64)
#undef SPZ_BS
        return cudaErrorInvalidValue;
    }
    if (pin == 0 || pin == 3)
    {   // throughput path: frames with many chunks of 8 / 16 dims (spz_encode_lane.cu)
        const cudaError_t e = launch_encode_lane(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_encode_family = 2; return e; }
    }
    if (p.variant == 0 && (pin == 0 || pin == 4))
    {   // wide rows (ndims 5..256), 8- and 16-bit, delta and FIRE: octet teams (spz_encode_wide.cu)
        const cudaError_t e = launch_encode_wide(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_encode_family = 3; return e; }
    }
    if (p.variant == 0 && (pin == 0 || pin == 2))
    {   // fast path for high-dim shapes; the generic kernel below covers everything else
        const cudaError_t e = launch_encode_hd(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_encode_family = 1; return e; }
    }
    const bool low = p.variant ? false : (elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2));
    if (low && (pin == 0 || pin == 3))
    {   // throughput path of the low-dim codecs: frames with many chunks (spz_encode_lowlane.cu)
        const cudaError_t e = launch_encode_lowlane(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_encode_family = 2; return e; }
    }
    g_last_encode_family = 0;
    p.team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
    p.dpl = (p.ndims + p.team - 1) / p.team;
    enc_layout(elem_sz, low, p);
    uint32_t teams_per_cta = kThreads / p.team;
    // keep a CTA under the 227 KB shared-memory limit
    while (teams_per_cta > 1 && (size_t)teams_per_cta * p.team_smem > 200 * 1024) teams_per_cta >>= 1;
    p.teams_per_cta = teams_per_cta;
    const size_t smem = (size_t)teams_per_cta * p.team_smem;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    const uint64_t grid64 = (p.nchunks + teams_per_cta - 1) / teams_per_cta;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    const uint32_t grid = (uint32_t)grid64;
#define SPZ_GO(Wv, Fv, Lv) return launch_one<Wv, Fv, Lv>(p, grid, smem, stream)
    if (elem_sz == 1) {
        if (predictor) { if (low) SPZ_GO(8, true, true); else SPZ_GO(8, true, false); }
        else           { if (low) SPZ_GO(8, false, true); else SPZ_GO(8, false, false); }
    } else {
        if (predictor) { if (low) SPZ_GO(16, true, true); else SPZ_GO(16, true, false); }
        else           { if (low) SPZ_GO(16, false, true); else SPZ_GO(16, false, false); }
    }
#undef SPZ_GO
}

}  // namespace spz

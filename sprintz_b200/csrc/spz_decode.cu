// spz_decode.cu -- Sprintz decode kernels for sm_100a.
//
// One team of lanes decodes one chunk == one complete reference stream (the inverse of spz_encode.cu;
// replaces sprintz_decompress_{delta,xff}_{8b,16b}, sprintz.cpp:110-129,232-251):
//
//   header   lane-per-dim extraction of the 3/4-bit width fields + team prefix sum -> bit offsets
//            (replaces the pdep/sad/pshufb header unpack, sprintz_delta_rle.cpp:537-617)
//   unpack   per lane: two aligned 32-bit loads + funnel shift + mask (replaces the pdep hot loop :672-701)
//   predict  per lane serial recurrence over the 8 rows: zigzag^-1, delta | FIRE (replaces :704-736 and
//            sprintz_xff_rle.cpp:1012-1128); FIRE's floor-multiply makes this serial in rows *and* blocks,
//            so parallelism comes from chunks x dims
//   runs     delta: repeat the previous row; FIRE: free-run the predictor with err = 0
//            (sprintz_delta_rle.cpp:640-667, sprintz_xff_rle.cpp:828-958 -- but with the *correct*
//            coefficients)
#include "spz_host.h"

namespace spz {

// R = rows per block (8 = the reference's format; 16 / 32 / 64 = BASELINE config 5's block-size axis, see spz_encode.cu)
template <int W, bool FIRE, bool LOW, int R>
__global__ void __launch_bounds__(kThreads) spz_decode_kernel(const DecParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr uint32_t WMASK = (1u << W) - 1u;
    static_assert(R == 8 || !LOW, "the low-dim (column-major) layout exists for 8-row blocks only");
    constexpr int GSH = (R == 8) ? 2 : (R == 16) ? 3 : (R == 32) ? 4 : 5;      // log2(R) - 1
    constexpr uint32_t kMinData = 16u * R;

    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int TEAM = (int)p.team;
    const int DPL = (int)p.dpl;
    const int D = (int)p.ndims;
    const int t = lane & (TEAM - 1);
    const int team_in_cta = (int)(threadIdx.x / TEAM);
    const uint64_t c = (uint64_t)blockIdx.x * p.teams_per_cta + team_in_cta;
    if (c >= p.nchunks) return;
    const uint32_t tmask = team_mask_of(lane, TEAM);
    const int Dp = TEAM * DPL;

    // ---- per-team shared memory: predictor state + the current slot's widths / bit offsets
    int32_t* s_pv = reinterpret_cast<int32_t*>(smem + (size_t)team_in_cta * p.team_smem);
    int32_t* s_pd = s_pv + Dp;
    int32_t* s_ct = s_pd + Dp;
    uint32_t* s_fo = reinterpret_cast<uint32_t*>(s_ct + Dp);   // [Dp] (bit offset << 8) | width
    for (int k = 0; k < DPL; k++) {
        s_pv[k * TEAM + t] = 0; s_pd[k * TEAM + t] = 0; s_ct[k * TEAM + t] = 0;
    }

    // Streams and directories are input from outside: every read below is bounded by `avail`, the readable bytes
    // from the chunk's first byte to src_end (the contract: the source covers the stream plus 8 bytes of slack).
    const uint64_t soff = p.offsets ? p.offsets[c] : c * p.slot_stride;
    const uint64_t src_bytes = p.src_end ? (uint64_t)(p.src_end - p.src) : ~0ull;
    const uint64_t avail = soff <= src_bytes ? src_bytes - soff : 0;
    const uint8_t* __restrict__ stream = p.src + (avail ? soff : 0);
    const uint64_t first = c * p.chunk_len;
    uint_t* __restrict__ out = reinterpret_cast<uint_t*>(p.dst) + first;
    // elements this chunk may write
    const uint64_t cap = (first >= p.dst_cap) ? 0 :
        ((p.dst_cap - first < p.chunk_len) ? (p.dst_cap - first) : p.chunk_len);

    uint32_t ngroups, remaining, status = kStatusOk;
    uint64_t pos = 0;                   // byte cursor in the stream (team-uniform)
    const uint32_t variant = p.variant; // 1 / 2: non-RLE sibling codecs (see spz_encode.cu)
    if (avail < (variant ? (variant == 3u ? 8u : 6u) : (p.has_meta ? 8u : 0u))) {
        ngroups = 0; remaining = 0; status = kStatusTruncated;
    } else if (variant) {               // format.h:75-85: [u32 len][u16 ndims]; ngroups is implied
        const uint32_t len32 = (uint32_t)stream[0] | ((uint32_t)stream[1] << 8) | ((uint32_t)stream[2] << 16) | ((uint32_t)stream[3] << 24);
        const uint32_t at = variant == 3u ? 6u : 4u;          // variant 3: [u48 len][u16 ndims] (sprintz_xff.cpp:64-69)
        const uint32_t nd = (uint32_t)stream[at] | ((uint32_t)stream[at + 1] << 8);
        pos = at + 2;
        if (len32 < kMinData) { ngroups = 0; remaining = len32; }
        else {
            if (nd != (uint32_t)D) status = kStatusBadDims;
            ngroups = len32 / (16u * (uint32_t)D);
            remaining = len32 - ngroups * 16u * (uint32_t)D;
        }
    } else if (p.has_meta) {            // format.h:47-62
        ngroups = (uint32_t)stream[0] | ((uint32_t)stream[1] << 8) | ((uint32_t)stream[2] << 16) | ((uint32_t)stream[3] << 24);
        remaining = (uint32_t)stream[4] | ((uint32_t)stream[5] << 8);
        const uint32_t nd = (uint32_t)stream[6] | ((uint32_t)stream[7] << 8);
        if (nd != (uint32_t)D) status = kStatusBadDims;
        pos = 8;
    } else {
        ngroups = p.meta[c].ngroups;
        remaining = p.meta[c].remaining & 0xffffu;
    }

    uint64_t o = 0;                     // output element cursor (team-uniform)
    const uint32_t H = (2u * (uint32_t)D * HB + 7u) >> 3;
    const uint64_t B = (uint64_t)R * D;

    if (status == kStatusOk && (variant ? ngroups != 0 : !(ngroups == 0 && remaining < kMinData))) {
        for (uint32_t g = 0; g < ngroups && status == kStatusOk; g++) {
            if (pos + H + 8 > avail) { status = kStatusTruncated; break; }      // header fields are read as aligned word pairs
            const uint8_t* hdr = stream + pos;
            pos += H;
            for (uint32_t s = 0; s < 2; s++) {
                // ---- widths of this slot: field (s*D + d) of the group header, prefix-summed over dims
                uint32_t total = 0;
                for (int k = 0; k < DPL; k++) {
                    const int d = k * TEAM + t;
                    uint32_t n = 0;
                    if (d < D) {
                        const uint32_t f = load_bits32(hdr, (s * (uint32_t)D + (uint32_t)d) * HB) & ((1u << HB) - 1u);
                        n = (f == (uint32_t)(W - 1)) ? (uint32_t)W : f;
                    }
                    const uint32_t incl = team_incl_scan(n, t, TEAM, tmask);
                    s_fo[d] = ((total + incl - n) << 8) | n;
                    total += team_last(incl, TEAM, tmask);
                }

                if (total == 0 && variant == 0u) {
                    // ---- run slot (sprintz_delta_rle.cpp:640-667)
                    if (pos + 2 > avail) { status = kStatusTruncated; break; }
                    uint32_t nrun = stream[pos] & 0x7f;
                    if (stream[pos] & 0x80) { nrun |= (uint32_t)stream[pos + 1] << 7; pos += 2; } else { pos += 1; }
                    const uint64_t nrows = (uint64_t)nrun * R;
                    if (o + nrows * D > cap) { status = kStatusOverflow; break; }
                    for (int k = 0; k < DPL; k++) {
                        const int d = k * TEAM + t;
                        if (d < D) {
                            uint32_t pv = (uint32_t)s_pv[d];
                            uint_t* op = out + o + d;
                            if (FIRE) {
                                int32_t pd = s_pd[d];
                                const int32_t coef = fire_coef<W, LOW>(s_ct[d]);
                                for (uint64_t r = 0; r < nrows; r++) {
                                    pd = fire_predict<W>(pd, coef);        // err == 0  ->  delta == prediction
                                    pv = (pv + (uint32_t)pd) & WMASK;
                                    op[r * D] = (uint_t)pv;
                                }
                                s_pd[d] = pd;
                                s_pv[d] = (int32_t)pv;
                            } else {
                                for (uint64_t r = 0; r < nrows; r++) op[r * D] = (uint_t)pv;
                            }
                        }
                    }
                    o += nrows * D;
                    continue;
                }

                // ---- data slot
                if (o + B > cap) { status = kStatusOverflow; break; }
                const uint32_t row_bytes = (total + 7u) >> 3;
                if (pos + (LOW ? total : row_bytes * R) + 8 > avail) { status = kStatusTruncated; break; }
                const uint8_t* pay = stream + pos;
                for (int k = 0; k < DPL; k++) {
                    const int d = k * TEAM + t;
                    if (d < D) {
                        const uint32_t fo = s_fo[d];
                        const uint32_t n = fo & 0xffu, off = fo >> 8;
                        const uint32_t fmask = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
                        // high-dim: row r's field at bit r*row_bytes*8 + off; low-dim: column at byte `off`, field r at r*n
                        const uint32_t bit0 = LOW ? off * 8u : off;
                        const uint32_t stride = LOW ? n : row_bytes * 8u;
                        uint32_t pv = (uint32_t)s_pv[d];
                        int32_t pd = 0, coef = 0, gs = 0;
                        if (FIRE) { pd = s_pd[d]; coef = fire_coef<W, LOW>(s_ct[d]); }
                        uint_t* op = out + o + d;
#pragma unroll 1
                        for (int sb = 0; sb < R / kBlockRows; sb++) {   // 8 rows at a time (R == 8: once)
                            uint32_t u[kBlockRows];
#pragma unroll
                            for (int r = 0; r < kBlockRows; r++)
                                u[r] = load_bits32(pay, bit0 + (uint32_t)(sb * kBlockRows + r) * stride) & fmask;
#pragma unroll
                            for (int r = 0; r < kBlockRows; r++) {
                                const int32_t err = zigzag_dec<W>(u[r]);
                                int32_t delta = err;
                                if (FIRE) {
                                    delta = sx<W>(err + fire_predict<W>(pd, coef));
                                    if (r & 1) gs += fire_grad<W>(err, pd);
                                }
                                pd = delta;
                                pv = (variant == 2u) ? u[r] : ((pv + (uint32_t)delta) & WMASK);
                                op[(size_t)(sb * kBlockRows + r) * D] = (uint_t)pv;
                            }
                        }
                        s_pv[d] = (int32_t)pv;
                        if (FIRE) {
                            s_pd[d] = pd;
                            s_ct[d] = sx<CB>(s_ct[d] + (sx<W>(gs) >> GSH));
                        }
                    }
                }
                pos += LOW ? total : row_bytes * R;
                o += B;
            }
        }
    }

    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (status == kStatusOk) {
        if (o + remaining > cap) {
            status = kStatusOverflow;
        } else if (pos + (uint64_t)remaining * ESZ > avail) {
            status = kStatusTruncated;
        } else {
            const uint8_t* tsrc = stream + pos;
            uint8_t* tdst = reinterpret_cast<uint8_t*>(out + o);
            const uint32_t nb = remaining * ESZ;
            for (uint32_t b = t; b < nb; b += TEAM) tdst[b] = tsrc[b];
            o += remaining;
        }
    }
    if (t == 0 && p.meta) {
        p.meta[c].status = status;
        if (p.has_meta || variant) { p.meta[c].nbytes = (uint32_t)(pos + (uint64_t)remaining * ESZ); p.meta[c].ngroups = ngroups; p.meta[c].remaining = (uint32_t)o; }
    }
}

// ------------------------------------------------------------------ launch

template <int W, bool FIRE, bool LOW, int R = 8>
static cudaError_t launch_one(const DecParams& p, uint32_t grid, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(spz_decode_kernel<W, FIRE, LOW, R>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    count_launch();
    spz_decode_kernel<W, FIRE, LOW, R><<<grid, p.teams_per_cta * p.team, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline uint32_t pow2_ceil(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

cudaError_t launch_decode_hd(int elem_sz, int predictor, DecParams p, cudaStream_t stream);
cudaError_t launch_decode_lane(int elem_sz, int predictor, DecParams p, cudaStream_t stream);
cudaError_t launch_decode_lowlane(int elem_sz, int predictor, DecParams p, cudaStream_t stream);
cudaError_t launch_decode_wide(int elem_sz, int predictor, DecParams p, cudaStream_t stream);

static thread_local int g_last_decode_family = 0;     // 0 generic team kernel, 1 fast team kernel (hd), 2 lane kernel, 3 wide-row kernel
int last_decode_family() { return g_last_decode_family; }

cudaError_t launch_decode(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    if (predictor >= 2) { p.variant = (uint32_t)predictor - 1u; }                  // SPZ_RM_DELTA / SPZ_RM_PLAIN / SPZ_RM_XFF (frames)
    if (p.variant) predictor = (p.variant == 3u) ? 1 : 0;
    const uint32_t R = p.block_rows ? p.block_rows : 8u;
    if (R != 8u) {
        // block-size axis (16 / 32 / 64 rows per block): generic kernels, high-dim layout for every ndims, RLE codecs only
        if (p.variant != 0 || (R != 16u && R != 32u && R != 64u)) return cudaErrorInvalidValue;
        g_last_decode_family = 0;
        p.team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
        p.dpl = (p.ndims + p.team - 1) / p.team;
        p.team_smem = (4u * 4u * p.team * p.dpl + 15u) & ~15u;
        uint32_t tpc = kThreads / p.team;
        while (tpc > 1 && (size_t)tpc * p.team_smem > 200 * 1024) tpc >>= 1;
        p.teams_per_cta = tpc;
        const size_t smem = (size_t)tpc * p.team_smem;
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
    {   // throughput path: frames with many chunks of 8 / 16 dims (spz_decode_lane.cu)
        const cudaError_t e = launch_decode_lane(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_decode_family = 2; return e; }
    }
    if (p.variant == 0)
    {   // wide rows (ndims a multiple of 8, 32..256), 8- and 16-bit, delta and FIRE: octet teams (spz_decode_wide.cu)
        const cudaError_t e = launch_decode_wide(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_decode_family = 3; return e; }
    }
    if (p.variant == 0)
    {   // fast path for high-dim shapes; the generic kernel below covers everything else
        const cudaError_t e = launch_decode_hd(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_decode_family = 1; return e; }
    }
    const bool low = p.variant ? false : (elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2));
    if (low)
    {   // throughput path of the low-dim codecs: frames with many chunks (spz_decode_lowlane.cu)
        const cudaError_t e = launch_decode_lowlane(elem_sz, predictor, p, stream);
        if (e != cudaErrorNotSupported) { g_last_decode_family = 2; return e; }
    }
    g_last_decode_family = 0;
    p.team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
    p.dpl = (p.ndims + p.team - 1) / p.team;
    p.team_smem = (4u * 4u * p.team * p.dpl + 15u) & ~15u;
    uint32_t teams_per_cta = kThreads / p.team;
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

// query-on-compressed without materialisation (p.qop != 0): only the lane kernels fold rows in registers.
// cudaErrorNotSupported: the caller decodes into a workspace and reduces that (spz_query.cu).
cudaError_t launch_query_fused(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    if (predictor >= 2) { p.variant = (uint32_t)predictor - 1u; }
    if (p.variant) predictor = (p.variant == 3u) ? 1 : 0;
    if (p.qop != 1u && p.qop != 2u) return cudaErrorInvalidValue;
    const bool low = p.variant ? false : (elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2));
    if (low) return launch_decode_lowlane(elem_sz, predictor, p, stream);
    return launch_decode_lane(elem_sz, predictor, p, stream);
}

}  // namespace spz

// spz_decode_hd.cu -- fast high-dim decode kernels (compile-time team size / dims-per-lane).
//
// Same work decomposition as spz_decode.cu (one team of lanes per chunk, lane-per-dim serial recurrence),
// rebuilt around the memory system after the first ncu pass showed the generic kernel L1TEX-bound:
//
//   input    the compressed stream is pulled through a per-team shared-memory ring with coalesced 128-bit
//            loads issued one fill ahead (register double buffering); field extraction is then two LDS of
//            words the whole team shares (broadcast) + one funnel shift
//   state    prev value / prev delta / FIRE counter live in registers (DPL is a template parameter)
//   headers  both slots of a group are parsed at once; the two prefix sums share one packed shuffle scan
//   output   decoded rows are staged in shared memory and leave as 128-bit coalesced stores
//
// Reference behaviour restated: sprintz_delta_rle.cpp:417-772, sprintz_xff_rle.cpp:568-1179 (high-dim only;
// the low-dim column-major variants stay on the generic kernel).
#include "spz_common.cuh"

namespace spz {

namespace {

__device__ __forceinline__ uint4 ldg_v4(const uint8_t* p)
{
    return __ldg(reinterpret_cast<const uint4*>(p));
}

}  // namespace

template <int W, bool FIRE, int TEAM, int DPL>
__global__ void __launch_bounds__(kThreads) spz_decode_hd_kernel(const DecParams p)
{
    typedef typename Traits<W>::uint_t uint_t;
    constexpr int ESZ = W / 8;
    constexpr int HB = Traits<W>::HB;
    constexpr int CB = Traits<W>::CB;
    constexpr uint32_t FILL = TEAM * 16;            // bytes one ring fill brings in
    constexpr uint32_t FMASK = (1u << HB) - 1u;

    extern __shared__ __align__(16) uint8_t smem[];

    const int lane = threadIdx.x & 31;
    const int t = lane & (TEAM - 1);
    const int team_in_cta = (int)(threadIdx.x / TEAM);
    const uint64_t c = (uint64_t)blockIdx.x * p.teams_per_cta + team_in_cta;
    if (c >= p.nchunks) return;
    const uint32_t tmask = team_mask_of(lane, TEAM);
    const int D = (int)p.ndims;

    uint8_t* const ring = smem + (size_t)team_in_cta * p.team_smem;     // ring_cap bytes + 16 mirror bytes
    uint8_t* const obuf = ring + p.ring_cap + 16;
    const uint32_t RCAP = p.ring_cap;
    const uint32_t RMASK = RCAP - 1u;

    // ---- input: absolute byte offsets are relative to `abase`, the 16-byte aligned-down stream start
    const uint8_t* const gstream = p.src + (p.offsets ? p.offsets[c] : c * p.slot_stride);
    const uint32_t mis_in = (uint32_t)((uintptr_t)gstream & 15);
    const uint8_t* const abase = gstream - mis_in;
    const uint64_t limit = (uint64_t)(p.src_end - abase);               // readable bytes from abase
    uint64_t wr = 0;                                                    // bytes of the stream staged so far
    uint4 nxt;

    auto prefetch = [&]() {
        const uint64_t o = wr + (uint64_t)t * 16;
        if (o + 16 <= limit) {
            nxt = ldg_v4(abase + o);
        } else {
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint32_t b = 0; b < 16; b++)
                if (o + b < limit) w[b >> 2] |= (uint32_t)abase[o + b] << (8 * (b & 3));
            nxt = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    // make stream bytes [.., need_end) available in the ring (team-uniform)
    auto ensure = [&](uint64_t need_end) {
        while (wr < need_end) {
            const uint32_t ro = ((uint32_t)wr + (uint32_t)t * 16u) & RMASK;
            *reinterpret_cast<uint4*>(ring + ro) = nxt;
            if (ro == 0) *reinterpret_cast<uint4*>(ring + RCAP) = nxt;      // mirror: word reads may run 4 bytes past the end
            wr += FILL;
            prefetch();
        }
        __syncwarp(tmask);
    };
    // 32-bit window at ring-relative bit position `bit` (may exceed the ring: masked)
    auto ring_bits = [&](uint32_t bit) -> uint32_t {
        const uint32_t bo = (bit >> 3) & (RMASK & ~3u);
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(ring + bo);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(ring + bo + 4);
        return __funnelshift_r(w0, w1, bit);
    };
    auto ring_byte = [&](uint64_t at) -> uint32_t { return ring[(uint32_t)at & RMASK]; };

    // ---- output
    const uint64_t first = c * p.chunk_len;
    uint_t* const out = reinterpret_cast<uint_t*>(p.dst) + first;
    const uint64_t cap = (first >= p.dst_cap) ? 0 :
        ((p.dst_cap - first < p.chunk_len) ? (p.dst_cap - first) : p.chunk_len);
    OutStage os;
    os.init(obuf, reinterpret_cast<uint8_t*>(out));

    // ---- predictor state (registers)
    uint32_t pv[DPL];
    int32_t pd[DPL], ct[DPL];
#pragma unroll
    for (int k = 0; k < DPL; k++) { pv[k] = 0; pd[k] = 0; ct[k] = 0; }

    prefetch();
    uint64_t pos = mis_in;              // absolute consume offset (team-uniform)
    uint32_t ngroups, remaining, status = kStatusOk;
    if (p.has_meta) {                   // format.h:47-62
        ensure(pos + 8);
        ngroups = ring_byte(pos) | (ring_byte(pos + 1) << 8) | (ring_byte(pos + 2) << 16) | (ring_byte(pos + 3) << 24);
        remaining = ring_byte(pos + 4) | (ring_byte(pos + 5) << 8);
        const uint32_t nd = ring_byte(pos + 6) | (ring_byte(pos + 7) << 8);
        if (nd != (uint32_t)D) status = kStatusBadDims;
        pos += 8;
    } else {
        ngroups = p.meta[c].ngroups;
        remaining = p.meta[c].remaining & 0xffffu;
    }

    uint64_t o = 0;                     // output element cursor (team-uniform)
    const uint32_t H = (2u * (uint32_t)D * HB + 7u) >> 3;
    const uint32_t B = (uint32_t)kBlockRows * D;
    const uint32_t max_group = H + 2u * B * ESZ;
    const uint32_t row_out = (uint32_t)D * ESZ;         // output bytes per row

    if (status == kStatusOk && !(ngroups == 0 && remaining < kMinDataSize)) {
        for (uint32_t g = 0; g < ngroups; g++) {
            __syncwarp(tmask);          // everyone is done reading the previous group before the ring advances
            ensure(pos + max_group + 8);

            // ---- both slots' widths: fields (s*D + d), prefix sums over dims share one packed scan
            const uint32_t hbit = ((uint32_t)pos & RMASK) * 8u;
            uint32_t nn[DPL], offs[DPL];
            uint32_t carry = 0;
#pragma unroll
            for (int k = 0; k < DPL; k++) {
                const int d = k * TEAM + t;
                uint32_t packed = 0;
                if (d < D) {
                    const uint32_t f0 = ring_bits(hbit + (uint32_t)d * HB) & FMASK;
                    const uint32_t f1 = ring_bits(hbit + (uint32_t)(D + d) * HB) & FMASK;
                    const uint32_t n0 = (f0 == (uint32_t)(W - 1)) ? (uint32_t)W : f0;
                    const uint32_t n1 = (f1 == (uint32_t)(W - 1)) ? (uint32_t)W : f1;
                    packed = n0 | (n1 << 16);
                }
                uint32_t incl = packed;
#pragma unroll
                for (int s = 1; s < TEAM; s <<= 1) {
                    const uint32_t y = __shfl_up_sync(tmask, incl, s, TEAM);
                    if (t >= s) incl += y;
                }
                nn[k] = packed;
                offs[k] = carry + incl - packed;
                carry += __shfl_sync(tmask, incl, TEAM - 1, TEAM);
            }
            pos += H;

#pragma unroll 1
            for (uint32_t s = 0; s < 2; s++) {
                const uint32_t sh16 = 16u * s;
                const uint32_t total = (carry >> sh16) & 0xffffu;
                if (total == 0) {
                    // ---- run slot (sprintz_delta_rle.cpp:640-667): rare, handled without the staging buffer
                    uint32_t nrun = ring_byte(pos) & 0x7f;
                    if (ring_byte(pos) & 0x80) { nrun |= ring_byte(pos + 1) << 7; pos += 2; } else { pos += 1; }
                    const uint64_t nrows = (uint64_t)nrun * kBlockRows;
                    if (o + nrows * D > cap) { status = kStatusOverflow; break; }
                    stage_flush(os, true, t, TEAM, tmask);
#pragma unroll
                    for (int k = 0; k < DPL; k++) {
                        const int d = k * TEAM + t;
                        if (d < D) {
                            uint_t* op = out + o + d;
                            if (FIRE) {
                                const int32_t coef = fire_coef<W, false>(ct[k]);
                                for (uint64_t r = 0; r < nrows; r++) {
                                    pd[k] = fire_predict<W>(pd[k], coef);      // err == 0  ->  delta == prediction
                                    pv[k] += (uint32_t)pd[k];
                                    op[r * D] = (uint_t)pv[k];
                                }
                            } else {
                                for (uint64_t r = 0; r < nrows; r++) op[r * D] = (uint_t)pv[k];
                            }
                        }
                    }
                    o += nrows * D;
                    __syncwarp(tmask);
                    os.init(obuf, reinterpret_cast<uint8_t*>(out + o));
                    continue;
                }

                // ---- data slot
                if (o + B > cap) { status = kStatusOverflow; break; }
                const uint32_t row_bits = ((total + 7u) >> 3) * 8u;
                const uint32_t pbit = ((uint32_t)pos & RMASK) * 8u;
#pragma unroll
                for (int k = 0; k < DPL; k++) {
                    const int d = k * TEAM + t;
                    if (d < D) {
                        const uint32_t n = (nn[k] >> sh16) & 0xffffu;
                        const uint32_t fmask = (1u << n) - 1u;
                        uint32_t bit = pbit + ((offs[k] >> sh16) & 0xffffu);
                        uint32_t u[kBlockRows];
#pragma unroll
                        for (int r = 0; r < kBlockRows; r++) { u[r] = ring_bits(bit) & fmask; bit += row_bits; }

                        int32_t coef = 0, gs = 0;
                        if (FIRE) coef = fire_coef<W, false>(ct[k]);
                        uint8_t* ob = os.buf + os.cur + (uint32_t)d * ESZ;
#pragma unroll
                        for (int r = 0; r < kBlockRows; r++) {
                            const int32_t err = (int32_t)((u[r] >> 1) ^ (0u - (u[r] & 1u)));    // zigzag^-1 (bitpack.h:303)
                            int32_t delta = err;
                            if (FIRE) {
                                delta = sx<W>(err + fire_predict<W>(pd[k], coef));
                                if (r & 1) gs += fire_grad<W>(err, pd[k]);
                            }
                            pd[k] = delta;
                            pv[k] += (uint32_t)delta;
                            *reinterpret_cast<uint_t*>(ob) = (uint_t)pv[k];
                            ob += row_out;
                        }
                        if (FIRE) ct[k] = sx<CB>(ct[k] + (sx<W>(gs) >> 2));
                    }
                }
                pos += (row_bits >> 3) * kBlockRows;
                os.cur += B * ESZ;
                o += B;
            }
            if (status != kStatusOk) break;
            stage_flush(os, false, t, TEAM, tmask);
        }
    }
    stage_flush(os, true, t, TEAM, tmask);

    // ---- trailing raw elements (also the whole stream when it was stored raw)
    if (status == kStatusOk) {
        if (o + remaining > cap) {
            status = kStatusOverflow;
        } else {
            const uint8_t* tsrc = abase + pos;
            uint8_t* tdst = reinterpret_cast<uint8_t*>(out + o);
            const uint32_t nb = remaining * ESZ;
            for (uint32_t b = t; b < nb; b += TEAM) tdst[b] = tsrc[b];
            o += remaining;
        }
    }
    if (t == 0 && p.meta) {
        p.meta[c].status = status;
        if (p.has_meta) {
            p.meta[c].nbytes = (uint32_t)(pos - mis_in + (uint64_t)remaining * ESZ);
            p.meta[c].ngroups = ngroups;
            p.meta[c].remaining = (uint32_t)o;
        }
    }
}

// ------------------------------------------------------------------ launch

template <int W, bool FIRE, int TEAM, int DPL>
static cudaError_t launch_hd(const DecParams& p, uint32_t grid, size_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(spz_decode_hd_kernel<W, FIRE, TEAM, DPL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    spz_decode_hd_kernel<W, FIRE, TEAM, DPL><<<grid, p.teams_per_cta * TEAM, smem, stream>>>(p);
    return cudaGetLastError();
}

static inline uint32_t pow2_ceil(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }

// Returns cudaErrorNotSupported when the shape is not covered by the fast path (caller falls back to the
// generic kernel, which is still CUDA -- there is no CPU path anywhere).
cudaError_t launch_decode_hd(int elem_sz, int predictor, DecParams p, cudaStream_t stream)
{
    const bool low = elem_sz == 1 ? (p.ndims <= 4) : (p.ndims <= 2);
    if (low || p.ndims > 128 || p.src_end == nullptr) return cudaErrorNotSupported;
    const uint32_t team = p.ndims >= 32 ? 32 : pow2_ceil(p.ndims);
    const uint32_t dpl = (p.ndims + team - 1) / team;
    const uint32_t hb = elem_sz == 1 ? 3 : 4;
    const uint32_t H = (2 * p.ndims * hb + 7) / 8;
    const uint32_t max_group = H + 2 * 8 * p.ndims * elem_sz;
    p.team = team; p.dpl = dpl;
    p.ring_cap = pow2_ceil(max_group + 8 + team * 16);
    p.out_cap = (16 + 2 * 8 * p.ndims * elem_sz + 16 + 15) & ~15u;
    p.team_smem = p.ring_cap + 16 + p.out_cap;
    uint32_t teams_per_cta = kThreads / team;
    p.teams_per_cta = teams_per_cta;
    const size_t smem = (size_t)teams_per_cta * p.team_smem;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const uint64_t grid64 = (p.nchunks + teams_per_cta - 1) / teams_per_cta;
    if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
    const uint32_t grid = (uint32_t)grid64;

#define SPZ_HD(Wv, Tv, Dv) \
    return predictor ? launch_hd<Wv, true, Tv, Dv>(p, grid, smem, stream) : launch_hd<Wv, false, Tv, Dv>(p, grid, smem, stream)
    if (elem_sz == 1) {
        if (team == 8) SPZ_HD(8, 8, 1);
        if (team == 16) SPZ_HD(8, 16, 1);
        if (dpl == 1) SPZ_HD(8, 32, 1);
        if (dpl == 2) SPZ_HD(8, 32, 2);
        if (dpl == 3) SPZ_HD(8, 32, 3);
        if (dpl == 4) SPZ_HD(8, 32, 4);
    } else {
        if (team == 4) SPZ_HD(16, 4, 1);
        if (team == 8) SPZ_HD(16, 8, 1);
        if (team == 16) SPZ_HD(16, 16, 1);
        if (dpl == 1) SPZ_HD(16, 32, 1);
        if (dpl == 2) SPZ_HD(16, 32, 2);
        if (dpl == 3) SPZ_HD(16, 32, 3);
        if (dpl == 4) SPZ_HD(16, 32, 4);
    }
#undef SPZ_HD
    return cudaErrorNotSupported;
}

}  // namespace spz

// spz_huff.cu -- entropy stage over compressed bytes ("SprintzFIRE+Huf", BASELINE config 3's "Huffman
// header").
//
// The reference tree holds no entropy coder: cpp/Compress/entropy.cpp:46-58 is a stub and README.md:29,45 points to Huff0
// inside the un-vendored dblalock/lzbench fork (commit 4e154a07).  PARITY UNPINNED: this stage is OUR byte-wise canonical
// Huffman coder, specified by a plain-C CPU statement in the test tree (the tests compare against it byte for byte) -- same
// idea as Huff0 (order-0 Huffman over bytes, 11-bit length limit, independently decodable streams), not its bit format.
//
// Container ("SPZH", little endian):
//   [64 B header][256 code lengths, 4 bits each = 128 B][u64 seg_off[nsegs+1]][u32 seg_nbytes[nsegs]][pad to 16]
//   [segment bit streams, each 16-byte aligned, LSB-first][16 B read slack]
// A segment is 4096 raw bytes (the last one may be short) coded with ONE table per container; segments are independent,
// so encode and decode run one lane per segment -- the same decomposition as the lane codecs (a lane is a complete
// scalar coder, no cooperation), the directory is the same device-wide exclusive scan as the frame's.
//
//   histogram   256 bins, per-CTA shared-memory counters, one 16-byte load per lane per step
//   table       one CTA: thread 0 builds the code lengths (two-queue Huffman over the sorted counts, deterministic ties,
//               lengths limited to 11 bits), all threads derive canonical codes (bit-reversed: the stream is LSB-first)
//   sizes       lane per segment: sum of code lengths -> bytes; exclusive scan (spz_scan_sizes_kernel) -> offsets
//   emit        lane per segment: 64-bit accumulator, 128-bit stores
//   decode      lane per segment: 11-bit lookup table in shared memory (symbol | length), 64-bit bit buffer
#include "spz_host.h"

namespace spz {

constexpr uint32_t kHuffMagic = 0x485a5053u;     // 'SPZH'
constexpr int kHuffMaxLen = 11;
constexpr uint32_t kHuffSeg = 4096;
constexpr uint32_t kHuffLut = 1u << kHuffMaxLen;
constexpr int kHuffThreads = 128;

struct HuffHeader {
    uint32_t magic;
    uint16_t version;
    uint8_t max_len;
    uint8_t pad;
    uint32_t seg_bytes;
    uint32_t reserved0;
    uint64_t raw_bytes;
    uint64_t nsegs;
    uint64_t payload_off;
    uint64_t payload_bytes;
    uint64_t container_bytes;
    uint64_t reserved1;
};
static_assert(sizeof(HuffHeader) == 64, "header is 64 bytes");

struct HuffLayout { uint64_t nsegs, off_len, off_dir, off_sizes, payload_off; };
static HuffLayout huff_layout(uint64_t raw_bytes)
{
    HuffLayout l;
    l.nsegs = raw_bytes ? (raw_bytes + kHuffSeg - 1) / kHuffSeg : 1;
    l.off_len = 64;
    l.off_dir = 64 + 128;
    l.off_sizes = l.off_dir + 8 * (l.nsegs + 1);
    l.payload_off = align16(l.off_sizes + 4 * l.nsegs);
    return l;
}
// workspace: [hist 256 x u64][enc table 256 x u32][ticket + pad 64 B][scan tile states][ChunkMeta x nsegs]
struct HuffWork { uint64_t off_hist, off_tab, off_ticket, off_state, off_meta, total; };
static HuffWork huff_work(uint64_t nsegs)
{
    HuffWork w;
    w.off_hist = 0;
    w.off_tab = 2048;
    w.off_ticket = 2048 + 1024;
    w.off_state = w.off_ticket + 64;
    w.off_meta = align16(w.off_state + 8 * (scan_tiles(nsegs) + 1));
    w.total = w.off_meta + 16 * nsegs;
    return w;
}

// ------------------------------------------------------------------ histogram
__global__ void __launch_bounds__(256) spz_huff_hist_kernel(const uint8_t* __restrict__ src, uint64_t nbytes, unsigned long long* hist)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t nvec = nbytes >> 4;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = s4[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&h[w[k] & 0xffu], 1u); atomicAdd(&h[(w[k] >> 8) & 0xffu], 1u);
            atomicAdd(&h[(w[k] >> 16) & 0xffu], 1u); atomicAdd(&h[w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (uint64_t b = (nvec << 4) + threadIdx.x; b < nbytes; b += blockDim.x) atomicAdd(&h[src[b]], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ------------------------------------------------------------------ code lengths (specified by the CPU statement in the test tree)
// Two-queue Huffman over the present symbols sorted by (count, symbol); ties take the leaf; depths become lengths; lengths
// above 11 are clipped and the Kraft sum is repaired by lengthening the deepest symbol still below 11 (rarest first).
__device__ void huff_build_lengths(const unsigned long long* cnt, uint8_t* len, uint16_t* idx, unsigned long long* w, uint16_t* parent)
{
    int n = 0;
    for (int s = 0; s < 256; s++) { len[s] = 0; if (cnt[s]) idx[n++] = (uint16_t)s; }
    if (n == 0) return;
    if (n == 1) { len[idx[0]] = 1; return; }
    for (int i = 1; i < n; i++) {                       // insertion sort by (count, symbol); idx starts symbol-sorted
        const uint16_t s = idx[i];
        int j = i - 1;
        while (j >= 0 && cnt[idx[j]] > cnt[s]) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = s;
    }
    for (int i = 0; i < n; i++) w[i] = cnt[idx[i]];
    int li = 0, ii = n, ni = n;
    for (int k = 0; k < n - 1; k++) {
        int a, b;
        if (li < n && (ii >= ni || w[li] <= w[ii])) a = li++; else a = ii++;
        if (li < n && (ii >= ni || w[li] <= w[ii])) b = li++; else b = ii++;
        w[ni] = w[a] + w[b];
        parent[a] = (uint16_t)ni; parent[b] = (uint16_t)ni;
        ni++;
    }
    // depths: internal nodes were created in order, so a parent's index is larger than its children's
    uint16_t* depth = reinterpret_cast<uint16_t*>(w);   // the weights are no longer needed
    const int root = 2 * n - 2;
    depth[root] = 0;
    int maxl = 0;
    for (int v = root - 1; v >= 0; v--) depth[v] = (uint16_t)(depth[parent[v]] + 1);
    for (int i = 0; i < n; i++) { len[idx[i]] = (uint8_t)(depth[i] > 255 ? 255 : depth[i]); if (depth[i] > maxl) maxl = depth[i]; }
    if (maxl <= kHuffMaxLen) return;
    unsigned long long K = 0;
    for (int i = 0; i < n; i++) { if (len[idx[i]] > kHuffMaxLen) len[idx[i]] = kHuffMaxLen; K += 1ull << (kHuffMaxLen - len[idx[i]]); }
    while (K > (1ull << kHuffMaxLen)) {
        int best = -1;                                  // idx is sorted by (count, symbol): the first hit at the deepest level is the rarest
        for (int i = 0; i < n; i++) {
            const int l = len[idx[i]];
            if (l < kHuffMaxLen && (best < 0 || l > len[idx[best]])) best = i;
        }
        const int s = idx[best];
        K -= 1ull << (kHuffMaxLen - len[s] - 1);
        len[s]++;
    }
}

// canonical codes (MSB-first, symbols of one length in increasing order), returned bit-reversed for the LSB-first stream
__device__ __forceinline__ uint32_t huff_code_of(const uint8_t* len, int s)
{
    const int l = len[s];
    uint32_t code = 0;
    for (int b = 1; b <= l; b++) {                      // next_code[b] = (next_code[b-1] + count[b-1]) << 1
        uint32_t c = 0;
        for (int t = 0; t < 256; t++) c += (len[t] == b - 1 && b > 1) ? 1u : 0u;
        code = (code + c) << 1;
    }
    for (int t = 0; t < s; t++) code += (len[t] == l) ? 1u : 0u;
    return __brev(code) >> (32 - l);
}

// encode side: lengths -> container nibbles + encode table (rev code | len << 16)
__global__ void __launch_bounds__(256) spz_huff_table_kernel(const unsigned long long* hist, uint8_t* nibbles, uint32_t* tab)
{
    __shared__ uint8_t len[256];
    __shared__ uint16_t idx[256];
    __shared__ unsigned long long w[512];
    __shared__ uint16_t parent[512];
    if (threadIdx.x == 0) huff_build_lengths(hist, len, idx, w, parent);
    __syncthreads();
    const int s = threadIdx.x;
    tab[s] = len[s] ? (huff_code_of(len, s) | ((uint32_t)len[s] << 16)) : 0u;
    if (s < 128) nibbles[s] = (uint8_t)(len[2 * s] | (len[2 * s + 1] << 4));
}

// decode side: container nibbles -> lookup table (symbol | len << 8; 0 = no code ends here: corrupt stream)
__global__ void __launch_bounds__(256) spz_huff_lut_kernel(const uint8_t* nibbles, uint16_t* lut, unsigned int* err)
{
    __shared__ uint8_t len[256];
    __shared__ unsigned int kraft;
    const int s = threadIdx.x;
    if (s == 0) kraft = 0;
    len[s] = (nibbles[s >> 1] >> ((s & 1) * 4)) & 15u;
    for (uint32_t e = s; e < kHuffLut; e += 256) lut[e] = 0;
    __syncthreads();
    const int l = len[s];
    if (l > kHuffMaxLen) atomicOr(err, 1u);
    if (l > 0 && l <= kHuffMaxLen) atomicAdd(&kraft, 1u << (kHuffMaxLen - l));
    __syncthreads();
    if (kraft > kHuffLut) { if (s == 0) atomicOr(err, 1u); return; }    // not a prefix code: the table is corrupt
    __threadfence();
    if (l > 0 && l <= kHuffMaxLen) {
        const uint32_t rev = huff_code_of(len, s);
        for (uint32_t f = 0; f < (1u << (kHuffMaxLen - l)); f++) lut[rev | (f << l)] = (uint16_t)(s | (l << 8));
    }
}

// ------------------------------------------------------------------ encode
__global__ void __launch_bounds__(kHuffThreads) spz_huff_size_kernel(const uint8_t* __restrict__ src, uint64_t nbytes, uint64_t nsegs,
                                                                     const uint32_t* __restrict__ tab, ChunkMeta* meta)
{
    __shared__ uint8_t slen[256];
    for (int s = threadIdx.x; s < 256; s += kHuffThreads) slen[s] = (uint8_t)(tab[s] >> 16);
    __syncthreads();
    const uint64_t seg = (uint64_t)blockIdx.x * kHuffThreads + threadIdx.x;
    if (seg >= nsegs) return;
    const uint64_t lo = seg * kHuffSeg;
    const uint32_t n = (uint32_t)(nbytes - lo < kHuffSeg ? nbytes - lo : kHuffSeg);
    uint32_t bits = 0;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + lo);
    for (uint32_t v = 0; v < (n >> 4); v++) {
        const uint4 q = s4[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
            bits += slen[w[k] & 0xffu] + slen[(w[k] >> 8) & 0xffu] + slen[(w[k] >> 16) & 0xffu] + slen[w[k] >> 24];
    }
    for (uint32_t b = n & ~15u; b < n; b++) bits += slen[src[lo + b]];
    ChunkMeta m;
    m.nbytes = (bits + 7u) >> 3; m.ngroups = 0; m.remaining = n; m.status = kStatusOk;
    meta[seg] = m;
}

__global__ void __launch_bounds__(kHuffThreads) spz_huff_emit_kernel(const uint8_t* __restrict__ src, uint64_t nbytes, uint64_t nsegs,
                                                                     const uint32_t* __restrict__ tab, const uint64_t* __restrict__ seg_off,
                                                                     uint8_t* __restrict__ payload)
{
    __shared__ uint32_t stab[256];
    for (int s = threadIdx.x; s < 256; s += kHuffThreads) stab[s] = tab[s];
    __syncthreads();
    const uint64_t seg = (uint64_t)blockIdx.x * kHuffThreads + threadIdx.x;
    if (seg >= nsegs) return;
    const uint64_t lo = seg * kHuffSeg;
    const uint32_t n = (uint32_t)(nbytes - lo < kHuffSeg ? nbytes - lo : kHuffSeg);
    uint4* out = reinterpret_cast<uint4*>(payload + seg_off[seg]);      // 16-byte aligned
    uint64_t acc = 0;                                   // pending bits (fewer than 32 between symbols)
    uint32_t nacc = 0;
    uint32_t ow[4];
    int nw = 0;
    auto put = [&](uint32_t sym) {
        const uint32_t e = stab[sym];
        acc |= (uint64_t)(e & 0xffffu) << nacc;
        nacc += e >> 16;
        if (nacc >= 32u) {
            ow[nw++] = (uint32_t)acc; acc >>= 32; nacc -= 32u;
            if (nw == 4) { *out++ = make_uint4(ow[0], ow[1], ow[2], ow[3]); nw = 0; }
        }
    };
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + lo);
    for (uint32_t v = 0; v < (n >> 4); v++) {
        const uint4 q = s4[v];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { put(w[k] & 0xffu); put((w[k] >> 8) & 0xffu); put((w[k] >> 16) & 0xffu); put(w[k] >> 24); }
    }
    for (uint32_t b = n & ~15u; b < n; b++) put(src[lo + b]);
    // the rest: whole 16-byte unit, zero padded (the directory rounds every segment to 16 bytes)
    if (nacc > 0u || nw > 0) {
        ow[nw++] = (uint32_t)acc;
        while (nw < 4) ow[nw++] = 0u;
        *out = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

__global__ void spz_huff_finish_kernel(HuffHeader* hdr, HuffHeader h, const uint64_t* seg_off, uint32_t* seg_nbytes, uint64_t nsegs,
                                       const unsigned int* err, uint64_t* d_out_bytes)
{
    const uint64_t total = seg_off[nsegs];
    uint8_t* gap = reinterpret_cast<uint8_t*>(seg_nbytes + nsegs);
    uint8_t* pay = reinterpret_cast<uint8_t*>(hdr) + h.payload_off;
    while (gap < pay) *gap++ = 0;
    for (int i = 0; i < 16; i++) pay[total + i] = 0;
    h.payload_bytes = total;
    h.container_bytes = h.payload_off + total + 16;
    if (*err) h.magic = 0;
    *hdr = h;
    if (d_out_bytes) *d_out_bytes = h.container_bytes;
}

// ------------------------------------------------------------------ decode
// status word: 1 bad table | 2 bad header | 4 bad directory | 8 a segment's stream is corrupt / truncated
__global__ void __launch_bounds__(kHuffThreads) spz_huff_decode_kernel(const uint8_t* __restrict__ cont, uint64_t cont_bytes, uint64_t raw_bytes,
                                                                       HuffLayout l, const uint16_t* __restrict__ lut_g,
                                                                       uint8_t* __restrict__ dst, unsigned int* err)
{
    __shared__ uint16_t lut[kHuffLut];
    for (uint32_t e = threadIdx.x; e < kHuffLut; e += kHuffThreads) lut[e] = lut_g[e];
    __syncthreads();
    const uint64_t seg = (uint64_t)blockIdx.x * kHuffThreads + threadIdx.x;
    if (seg >= l.nsegs) return;
    const HuffHeader* h = reinterpret_cast<const HuffHeader*>(cont);
    if (seg == 0 && (h->magic != kHuffMagic || h->seg_bytes != kHuffSeg || h->raw_bytes != raw_bytes || h->nsegs != l.nsegs ||
                     h->payload_off != l.payload_off || h->max_len != kHuffMaxLen))
        atomicOr(err, 2u);
    const uint64_t* seg_off = reinterpret_cast<const uint64_t*>(cont + l.off_dir);
    const uint32_t* seg_nb = reinterpret_cast<const uint32_t*>(cont + l.off_sizes);
    const uint64_t a = seg_off[seg];
    const uint32_t nb = seg_nb[seg];
    const uint64_t pay_room = cont_bytes - l.payload_off;              // the caller checked cont_bytes >= payload_off + 16
    if (a > pay_room || (uint64_t)nb + 16u > pay_room - a || (a & 15u)) { atomicOr(err, 4u); return; }
    const uint64_t lo = seg * kHuffSeg;
    const uint32_t n = (uint32_t)(raw_bytes - lo < kHuffSeg ? raw_bytes - lo : kHuffSeg);
    const uint32_t* in = reinterpret_cast<const uint32_t*>(cont + l.payload_off + a);
    const uint32_t nwords = (nb + 3u) >> 2;             // words that hold stream bits (the 16 bytes of slack cover the over-read)
    uint64_t acc = 0;
    uint32_t nacc = 0, wi = 0;
    uint4* out = reinterpret_cast<uint4*>(dst + lo);
    uint32_t ow[4] = {0, 0, 0, 0};
    bool bad = false;
    for (uint32_t i = 0; i < n; i++) {
        if (nacc < (uint32_t)kHuffMaxLen) {
            const uint32_t wv = wi < nwords ? in[wi] : 0u;
            wi++;
            acc |= (uint64_t)wv << nacc; nacc += 32u;
        }
        const uint32_t e = lut[(uint32_t)acc & (kHuffLut - 1u)];
        const uint32_t len = e >> 8;
        bad |= (len == 0u);
        acc >>= len; nacc -= len;
        ow[(i >> 2) & 3u] |= (e & 0xffu) << (8u * (i & 3u));
        if ((i & 15u) == 15u) { *out++ = make_uint4(ow[0], ow[1], ow[2], ow[3]); ow[0] = ow[1] = ow[2] = ow[3] = 0; }
    }
    if (n & 15u) {                                      // the stream's last, partial 16-byte unit
        uint8_t* tail = reinterpret_cast<uint8_t*>(out);
        for (uint32_t b = 0; b < (n & 15u); b++) tail[b] = (uint8_t)(ow[b >> 2] >> (8u * (b & 3u)));
    }
    // the bits consumed must be exactly the bits of the segment (up to the byte padding)
    const uint64_t used_bits = (uint64_t)wi * 32u - nacc;
    if (bad || used_bits > (uint64_t)nb * 8u || used_bits + 7u < (uint64_t)nb * 8u) atomicOr(err, 8u);
}

}  // namespace spz

using namespace spz;

// ------------------------------------------------------------------ C ABI

extern "C" uint64_t spz_huff_max_bytes(uint64_t raw_bytes)
{
    const HuffLayout l = huff_layout(raw_bytes);
    // worst case 11 bits per byte, every segment rounded up to 16 bytes
    return l.payload_off + l.nsegs * (uint64_t)(((kHuffSeg * kHuffMaxLen + 7) / 8 + 15) & ~15) + 16;
}

extern "C" uint64_t spz_huff_workspace_bytes(uint64_t raw_bytes)
{
    return huff_work(huff_layout(raw_bytes).nsegs).total + 2 * kHuffLut;      // + the decode side's lookup table
}

extern "C" int spz_huff_encode_dev(const void* d_src, uint64_t raw_bytes, void* d_cont, uint64_t cont_cap, void* d_work,
                                   uint64_t work_bytes, uint64_t* d_cont_bytes, spz_stream_t stream_)
{
    if ((!d_src && raw_bytes) || !d_cont || !d_work || ((uintptr_t)d_src & 15) || ((uintptr_t)d_cont & 15) || ((uintptr_t)d_work & 15)) return SPZ_E_ARG;
    if (cont_cap < spz_huff_max_bytes(raw_bytes)) return SPZ_E_CAPACITY;
    if (work_bytes < spz_huff_workspace_bytes(raw_bytes)) return SPZ_E_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream_;
    const HuffLayout l = huff_layout(raw_bytes);
    const HuffWork w = huff_work(l.nsegs);
    uint8_t* work = (uint8_t*)d_work;
    uint8_t* cont = (uint8_t*)d_cont;
    const uint8_t* src = (const uint8_t*)d_src;
    CK(cudaMemsetAsync(work, 0, w.off_meta, st));                       // histogram, table, ticket, scan states
    unsigned long long* hist = (unsigned long long*)(work + w.off_hist);
    uint32_t* tab = (uint32_t*)(work + w.off_tab);
    unsigned int* ticket = (unsigned int*)(work + w.off_ticket);
    unsigned int* err = ticket + 1;
    ChunkMeta* meta = (ChunkMeta*)(work + w.off_meta);
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint64_t hgrid = (raw_bytes / 16 + 255) / 256;
    if (hgrid > (uint64_t)sms * 8) hgrid = (uint64_t)sms * 8;
    if (hgrid == 0) hgrid = 1;
    count_launch();
    spz_huff_hist_kernel<<<(unsigned)hgrid, 256, 0, st>>>(src, raw_bytes, hist);
    CK(cudaGetLastError());
    count_launch();
    spz_huff_table_kernel<<<1, 256, 0, st>>>(hist, cont + l.off_len, tab);
    CK(cudaGetLastError());
    const uint64_t sgrid = (l.nsegs + kHuffThreads - 1) / kHuffThreads;
    if (sgrid > 0x7fffffffull) return SPZ_E_UNSUPPORTED;
    count_launch();
    spz_huff_size_kernel<<<(unsigned)sgrid, kHuffThreads, 0, st>>>(src, raw_bytes, l.nsegs, tab, meta);
    CK(cudaGetLastError());
    uint64_t* seg_off = (uint64_t*)(cont + l.off_dir);
    uint32_t* seg_nb = (uint32_t*)(cont + l.off_sizes);
    spz_frame_header dummy;
    memset(&dummy, 0, sizeof(dummy));
    CK(launch_scan_sizes(meta, l.nsegs, seg_off, seg_nb, (unsigned long long*)(work + w.off_state), ticket, nullptr, nullptr, 0,
                         nullptr, dummy, nullptr, err, st));
    count_launch();
    spz_huff_emit_kernel<<<(unsigned)sgrid, kHuffThreads, 0, st>>>(src, raw_bytes, l.nsegs, tab, seg_off, cont + l.payload_off);
    CK(cudaGetLastError());
    HuffHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kHuffMagic; h.version = 1; h.max_len = kHuffMaxLen; h.seg_bytes = kHuffSeg; h.raw_bytes = raw_bytes;
    h.nsegs = l.nsegs; h.payload_off = l.payload_off;
    count_launch();
    spz_huff_finish_kernel<<<1, 1, 0, st>>>((HuffHeader*)cont, h, seg_off, seg_nb, l.nsegs, err, d_cont_bytes);
    CK(cudaGetLastError());
    return SPZ_OK;
}

extern "C" int spz_huff_decode_dev(const void* d_cont, uint64_t cont_bytes, uint64_t raw_bytes, void* d_dest, uint64_t dest_cap,
                                   void* d_work, uint64_t work_bytes, unsigned int* d_status, spz_stream_t stream_)
{
    if (!d_cont || (!d_dest && raw_bytes) || !d_work || !d_status || ((uintptr_t)d_cont & 15) || ((uintptr_t)d_dest & 15) || ((uintptr_t)d_work & 15))
        return SPZ_E_ARG;
    if (dest_cap < raw_bytes) return SPZ_E_CAPACITY;
    if (work_bytes < 2 * kHuffLut) return SPZ_E_CAPACITY;
    const HuffLayout l = huff_layout(raw_bytes);
    if (cont_bytes < l.payload_off + 16) return SPZ_E_STREAM;
    cudaStream_t st = (cudaStream_t)stream_;
    CK(cudaMemsetAsync(d_status, 0, sizeof(unsigned int), st));
    uint16_t* lut = (uint16_t*)d_work;
    count_launch();
    spz_huff_lut_kernel<<<1, 256, 0, st>>>((const uint8_t*)d_cont + l.off_len, lut, d_status);
    CK(cudaGetLastError());
    const uint64_t grid = (l.nsegs + kHuffThreads - 1) / kHuffThreads;
    if (grid > 0x7fffffffull) return SPZ_E_UNSUPPORTED;
    count_launch();
    spz_huff_decode_kernel<<<(unsigned)grid, kHuffThreads, 0, st>>>((const uint8_t*)d_cont, cont_bytes, raw_bytes, l, lut, (uint8_t*)d_dest,
                                                                   d_status);
    CK(cudaGetLastError());
    return SPZ_OK;
}

/* oracle/huff_oracle.c -- TEST INFRASTRUCTURE ONLY: CPU specification of the entropy stage.
 *
 * PARITY UNPINNED.  The reference tree has no entropy coder: cpp/Compress/entropy.cpp:46-58 is a stub, the paper's
 * "SprintzFIRE+Huf" (communicate/method.tex:303) used Huff0 from the un-vendored dblalock/lzbench fork (README.md:29,45,
 * commit 4e154a07fdfe85b48af9b0f6ce174e337ed81122).  This file therefore does not restate reference code; it DEFINES our
 * byte-wise canonical Huffman stage (order-0 Huffman over bytes, 11-bit length limit, independently decodable 4 KB
 * segments -- the properties Huff0 has, not its bit format), and the CUDA stage (sprintz_b200/csrc/spz_huff.cu) must
 * produce exactly these container bytes.
 *
 * Container ("SPZH", little endian):
 *   [64 B header][256 code lengths, 4 bits each][u64 seg_off[nsegs+1]][u32 seg_nbytes[nsegs]][pad to 16]
 *   [segment bit streams, each starting 16-byte aligned, LSB-first, zero padded][16 B read slack]
 * header: u32 magic 'SPZH', u16 version 1, u8 max_len 11, u8 0, u32 seg_bytes 4096, u32 0, u64 raw_bytes, u64 nsegs,
 *         u64 payload_off, u64 payload_bytes, u64 container_bytes, u64 0
 *
 * Code lengths: the present symbols sorted by (count, symbol) feed a two-queue Huffman construction (a tie between the
 * next leaf and the next internal node takes the leaf); node depths are the lengths; if the deepest exceeds 11, lengths
 * are clipped to 11 and the Kraft sum is repaired by repeatedly lengthening the deepest symbol still below 11 (the rarest
 * of them, then the smallest symbol).  Codes are canonical (within a length in symbol order, MSB-first) and written
 * bit-reversed because the streams are LSB-first. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HUFF_MAGIC 0x485a5053u
#define HUFF_MAXLEN 11
#define HUFF_SEG 4096u

void oracle_huff_lengths(const uint64_t* cnt, uint8_t* len)
{
    int idx[256], n = 0;
    memset(len, 0, 256);
    for (int s = 0; s < 256; s++) if (cnt[s]) idx[n++] = s;
    if (n == 0) return;
    if (n == 1) { len[idx[0]] = 1; return; }
    for (int i = 1; i < n; i++) {                   /* stable insertion sort: (count, symbol) ascending */
        int s = idx[i], j = i - 1;
        while (j >= 0 && cnt[idx[j]] > cnt[s]) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = s;
    }
    uint64_t w[512];
    int parent[512], depth[512];
    for (int i = 0; i < n; i++) w[i] = cnt[idx[i]];
    int leaf = 0, inner = n, next = n;
    for (int k = 0; k < n - 1; k++) {
        int pick[2];
        for (int q = 0; q < 2; q++) {
            if (leaf < n && (inner >= next || w[leaf] <= w[inner])) pick[q] = leaf++;
            else pick[q] = inner++;
        }
        w[next] = w[pick[0]] + w[pick[1]];
        parent[pick[0]] = next; parent[pick[1]] = next;
        next++;
    }
    int root = 2 * n - 2, maxl = 0;
    depth[root] = 0;
    for (int v = root - 1; v >= 0; v--) depth[v] = depth[parent[v]] + 1;
    for (int i = 0; i < n; i++) { len[idx[i]] = (uint8_t)(depth[i] > 255 ? 255 : depth[i]); if (depth[i] > maxl) maxl = depth[i]; }
    if (maxl <= HUFF_MAXLEN) return;
    uint64_t K = 0;
    for (int i = 0; i < n; i++) {
        if (len[idx[i]] > HUFF_MAXLEN) len[idx[i]] = HUFF_MAXLEN;
        K += 1ull << (HUFF_MAXLEN - len[idx[i]]);
    }
    while (K > (1ull << HUFF_MAXLEN)) {
        int best = -1;
        for (int i = 0; i < n; i++) {
            int l = len[idx[i]];
            if (l < HUFF_MAXLEN && (best < 0 || l > len[idx[best]])) best = i;
        }
        int s = idx[best];
        K -= 1ull << (HUFF_MAXLEN - len[s] - 1);
        len[s]++;
    }
}

static uint32_t bitrev(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

/* canonical codes, bit-reversed; code[s] valid when len[s] > 0 */
static void canonical(const uint8_t* len, uint32_t* code)
{
    uint32_t count[HUFF_MAXLEN + 2] = {0}, next[HUFF_MAXLEN + 2] = {0};
    for (int s = 0; s < 256; s++) if (len[s] && len[s] <= HUFF_MAXLEN) count[len[s]]++;
    uint32_t c = 0;
    for (int b = 1; b <= HUFF_MAXLEN; b++) { c = (c + count[b - 1]) << 1; next[b] = c; }
    for (int s = 0; s < 256; s++) {
        code[s] = 0;
        if (len[s] && len[s] <= HUFF_MAXLEN) code[s] = bitrev(next[len[s]]++, len[s]);
    }
}

typedef struct { uint64_t nsegs, off_len, off_dir, off_sizes, payload_off; } layout_t;
static layout_t layout(uint64_t n)
{
    layout_t l;
    l.nsegs = n ? (n + HUFF_SEG - 1) / HUFF_SEG : 1;
    l.off_len = 64; l.off_dir = 64 + 128;
    l.off_sizes = l.off_dir + 8 * (l.nsegs + 1);
    l.payload_off = (l.off_sizes + 4 * l.nsegs + 15) & ~15ull;
    return l;
}

uint64_t oracle_huff_max_bytes(uint64_t n)
{
    layout_t l = layout(n);
    return l.payload_off + l.nsegs * (uint64_t)(((HUFF_SEG * HUFF_MAXLEN + 7) / 8 + 15) & ~15) + 16;
}

static void put64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static void put32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint64_t get64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
static uint32_t get32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); return v; }

/* -> container bytes */
int64_t oracle_huff_encode(const uint8_t* src, uint64_t n, uint8_t* dst)
{
    layout_t l = layout(n);
    uint64_t cnt[256] = {0};
    uint8_t len[256];
    uint32_t code[256];
    for (uint64_t i = 0; i < n; i++) cnt[src[i]]++;
    oracle_huff_lengths(cnt, len);
    canonical(len, code);
    memset(dst, 0, l.payload_off);
    for (int s = 0; s < 128; s++) dst[l.off_len + s] = (uint8_t)(len[2 * s] | (len[2 * s + 1] << 4));
    uint64_t off = 0;
    for (uint64_t g = 0; g < l.nsegs; g++) {
        uint64_t lo = g * HUFF_SEG, m = n - lo < HUFF_SEG ? n - lo : HUFF_SEG;
        uint64_t bits = 0;
        for (uint64_t i = 0; i < m; i++) bits += len[src[lo + i]];
        uint32_t nb = (uint32_t)((bits + 7) / 8), rounded = (nb + 15u) & ~15u;
        uint8_t* out = dst + l.payload_off + off;
        memset(out, 0, rounded);
        uint64_t bit = 0;
        for (uint64_t i = 0; i < m; i++) {
            uint32_t c = code[src[lo + i]];
            int ln = len[src[lo + i]];
            for (int b = 0; b < ln; b++, bit++) if ((c >> b) & 1u) out[bit >> 3] |= (uint8_t)(1u << (bit & 7));
        }
        put64(dst + l.off_dir + 8 * g, off);
        put32(dst + l.off_sizes + 4 * g, nb);
        off += rounded;
    }
    put64(dst + l.off_dir + 8 * l.nsegs, off);
    memset(dst + l.payload_off + off, 0, 16);
    put32(dst + 0, HUFF_MAGIC);
    dst[4] = 1; dst[5] = 0; dst[6] = HUFF_MAXLEN; dst[7] = 0;
    put32(dst + 8, HUFF_SEG); put32(dst + 12, 0);
    put64(dst + 16, n); put64(dst + 24, l.nsegs); put64(dst + 32, l.payload_off); put64(dst + 40, off);
    put64(dst + 48, l.payload_off + off + 16); put64(dst + 56, 0);
    return (int64_t)(l.payload_off + off + 16);
}

/* -> n, or -1 when the container is malformed */
int64_t oracle_huff_decode(const uint8_t* cont, uint64_t cont_bytes, uint8_t* dst, uint64_t n)
{
    layout_t l = layout(n);
    if (cont_bytes < l.payload_off + 16) return -1;
    if (get32(cont) != HUFF_MAGIC || get32(cont + 8) != HUFF_SEG || get64(cont + 16) != n || get64(cont + 24) != l.nsegs ||
        get64(cont + 32) != l.payload_off || cont[6] != HUFF_MAXLEN) return -1;
    uint8_t len[256];
    uint32_t code[256];
    uint64_t K = 0;
    for (int s = 0; s < 256; s++) {
        len[s] = (cont[l.off_len + (s >> 1)] >> ((s & 1) * 4)) & 15;
        if (len[s] > HUFF_MAXLEN) return -1;
        if (len[s]) K += 1ull << (HUFF_MAXLEN - len[s]);
    }
    if (K > (1ull << HUFF_MAXLEN)) return -1;
    canonical(len, code);
    static uint16_t lut[1 << HUFF_MAXLEN];
    memset(lut, 0, sizeof(lut));
    for (int s = 0; s < 256; s++)
        if (len[s]) for (uint32_t f = 0; f < (1u << (HUFF_MAXLEN - len[s])); f++) lut[code[s] | (f << len[s])] = (uint16_t)(s | (len[s] << 8));
    for (uint64_t g = 0; g < l.nsegs; g++) {
        uint64_t a = get64(cont + l.off_dir + 8 * g);
        uint32_t nb = get32(cont + l.off_sizes + 4 * g);
        uint64_t room = cont_bytes - l.payload_off;
        if (a > room || (uint64_t)nb + 16 > room - a || (a & 15)) return -1;
        const uint8_t* in = cont + l.payload_off + a;
        uint64_t lo = g * HUFF_SEG, m = n - lo < HUFF_SEG ? n - lo : HUFF_SEG, bit = 0;
        for (uint64_t i = 0; i < m; i++) {
            uint32_t win = 0;
            for (int b = 0; b < HUFF_MAXLEN; b++) {
                uint64_t at = bit + b;
                if ((at >> 3) < nb) win |= (uint32_t)((in[at >> 3] >> (at & 7)) & 1u) << b;
            }
            uint16_t e = lut[win];
            if ((e >> 8) == 0) return -1;
            dst[lo + i] = (uint8_t)e;
            bit += e >> 8;
        }
        if (bit > (uint64_t)nb * 8 || bit + 7 < (uint64_t)nb * 8) return -1;
    }
    return (int64_t)n;
}

/* oracle/sprintz_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, scalar, portable restatement of the reference Sprintz codec hot path:
 *   sprintz_{compress,decompress}_{delta,xff}_{8b,16b}      (sprintz.h:16-32, sprintz.cpp:92-251)
 * which dispatch on ndims (sprintz.cpp:69-85) to
 *   compress_rowmajor_delta_rle           sprintz_delta_rle.cpp:54-404     (high-dim)
 *   compress_rowmajor_delta_rle_lowdim    sprintz_delta_lowdim.cpp:38-384  (8b: D<=4, 16b: D<=2)
 *   compress_rowmajor_xff_rle             sprintz_xff_rle.cpp:60-555       (high-dim FIRE)
 *   compress_rowmajor_xff_rle_lowdim      sprintz_xff_lowdim.cpp:43-400    (low-dim FIRE)
 * and the matching decoders (sprintz_delta_rle.cpp:417-772, sprintz_delta_lowdim.cpp:397-794,
 * sprintz_xff_rle.cpp:568-1179, sprintz_xff_lowdim.cpp:413-1119).
 *
 * It is written from the *behaviour* of that code, not from its
 * structure: the reference encoder is a goto-driven state machine over pext'ed stripes and
 * the reference decoders are AVX2; this file is a slot-oriented bit-stream writer/reader.
 *
 * PINNING.  The reference ships no golden compressed bytes (its tests are round-trips,
 * test/compress_testing.hpp:23-51).  This restatement is therefore pinned against the
 * reference itself: oracle/_ref/libsprintz_ref.so is the unmodified reference compiled from
 * /root/reference by oracle/Makefile, and tests/test_oracle_vs_ref.py + the committed
 * fixtures under tests/golden/ (generated from that library by tests/golden/make_golden.py)
 * require byte-identical streams, identical return values and identical decoded integers.
 *
 * Known deliberate difference: the reference's 16-bit high-dim FIRE *decoder* mis-builds the
 * filter coefficients inside run slots (sprintz_xff_rle.cpp:892-929 vs the correct :1073-1081,
 * SURVEY.md section 8a row D4).  This decoder is the true inverse of the (bit-exact) encoder.
 */
#include "sprintz_oracle.h"

#include <stdlib.h>
#include <string.h>

#define BLOCK_SZ 8
#define MAX_RUN_NBLOCKS 0x7fff          /* sprintz_delta_rle.cpp:68 */
#define MIN_DATA_SIZE (8 * BLOCK_SZ * 2) /* sprintz_delta_rle.cpp:71 -> 128 elements */

typedef struct {
    int w;      /* element bits: 8 | 16 */
    int esz;    /* element bytes */
    int hb;     /* header field bits: 3 | 4   (sprintz_delta_rle.cpp:61) */
    int cb;     /* FIRE counter bits: 16 | 32 (util.h:39-47) */
    int fire;   /* 0 delta, 1 xff */
    int low;    /* low-dim variant? (sprintz.cpp:69-85) */
    int D;
    uint32_t H; /* header bytes per group */
    int R;      /* rows per block: 8 in the reference (sprintz_delta_rle.cpp:62, sprintz_xff_rle.cpp:78-82 block_sz) */
    int gshift; /* FIRE: counter += grad_sum >> (log2_block_sz - log2_learning_downsample) (sprintz_xff_rle.cpp:271-275) */
} cfg_t;

typedef struct {
    uint32_t* prev_val;   /* [D] */
    int32_t*  prev_delta; /* [D] */
    int32_t*  counter;    /* [D] */
} state_t;

static int64_t sx(int64_t v, int bits)   /* wrap to signed `bits`-bit */
{
    uint64_t m = (bits == 64) ? ~0ull : ((1ull << bits) - 1);
    uint64_t u = (uint64_t)v & m;
    if (u >> (bits - 1)) u |= ~m;
    return (int64_t)u;
}

static int bitlength(uint32_t x) { int n = 0; while (x) { n++; x >>= 1; } return n; }

static cfg_t make_cfg(int predictor, int elem_sz, int D)
{
    cfg_t c;
    c.esz = elem_sz; c.w = 8 * elem_sz;
    c.hb = (elem_sz == 1) ? 3 : 4;
    c.cb = (elem_sz == 1) ? 16 : 32;
    c.fire = predictor != 0;
    c.low = (elem_sz == 1) ? (D <= 4) : (D <= 2);
    c.D = D;
    c.H = (uint32_t)((2u * (uint32_t)D * (uint32_t)c.hb + 7u) / 8u);
    c.R = BLOCK_SZ; c.gshift = 2;
    return c;
}

static uint32_t load_elem(const uint8_t* p, uint64_t idx, int esz)
{
    return esz == 1 ? p[idx] : (uint32_t)p[2 * idx] | ((uint32_t)p[2 * idx + 1] << 8);
}
static void store_elem(uint8_t* p, uint64_t idx, int esz, uint32_t v)
{
    if (esz == 1) { p[idx] = (uint8_t)v; }
    else { p[2 * idx] = (uint8_t)v; p[2 * idx + 1] = (uint8_t)(v >> 8); }
}

/* FIRE coefficient from the counter.
 * high-dim: (int16_t)((counter >> (1 + w-4)) << (w-4))      sprintz_xff_rle.cpp:209-217
 * low-dim : counter >> 1, counter_t wide (no truncation)     sprintz_xff_lowdim.cpp:38-39,170-173 */
static int32_t fire_coef(const cfg_t* c, int32_t counter)
{
    if (c->low) return (int32_t)sx((int64_t)counter >> 1, c->cb);
    int sh = c->w - 4;
    return (int32_t)sx(((int64_t)counter >> (1 + sh)) * (1ll << sh), 16);
}

/* prediction = (int_t)(((counter_t)prev_delta * coef) >> w); the product is an `int`
 * in the reference, i.e. wraps at 32 bits (sprintz_xff_rle.cpp:225). */
static int32_t fire_predict(const cfg_t* c, int32_t prev_delta, int32_t coef)
{
    int64_t prod = sx((int64_t)prev_delta * (int64_t)coef, 32);
    return (int32_t)sx(prod >> c->w, c->w);
}

/* bits -> nbits with the variant's promotion rule.
 * 8b (all variants) : 7 -> 8                          bitpack.h:72-93, sprintz_delta_lowdim.cpp:172-173
 * 16b high-dim      : low byte rule 7->8, and if the high byte is used n = 8 + rule(high)
 *                     (so 15 -> 16)                  sprintz_delta_rle.cpp:177-187
 * 16b low-dim       : only 15 -> 16                   sprintz_delta_lowdim.cpp:172-173 */
static int nbits_rule(const cfg_t* c, uint32_t mask)
{
    int n = bitlength(mask);
    if (c->low) { if (n == c->w - 1) n = c->w; return n; }
    if (c->w == 8) { if (n == 7) n = 8; return n; }
    if (mask >> 8) { int u = bitlength(mask >> 8); if (u == 7) u = 8; return 8 + u; }
    if (n == 7) n = 8;
    return n;
}

/* One block (8 rows x D) of forecasting + zigzag + per-dim width.
 * delta: sprintz_delta_rle.cpp:161-201; FIRE: sprintz_xff_rle.cpp:197-298.
 * zz is [8][D] row-major. Returns sum of nbits. */
static uint32_t encode_block(const cfg_t* c, state_t* st, const uint8_t* src, uint64_t first_elem,
                             uint32_t* zz, uint8_t* nbits)
{
    const int D = c->D, w = c->w;
    const uint32_t wmask = (1u << w) - 1u;
    uint32_t total = 0;
    for (int d = 0; d < D; d++) {
        int32_t coef = c->fire ? fire_coef(c, st->counter[d]) : 0;
        uint32_t mask = 0, pv = st->prev_val[d];
        int32_t pd = st->prev_delta[d], gs = 0;
        for (int r = 0; r < c->R; r++) {
            uint32_t v = load_elem(src, first_elem + (uint64_t)r * D + d, c->esz);
            int32_t delta = (int32_t)sx((int64_t)v - (int64_t)pv, w);
            int32_t err = delta;
            if (c->fire) {
                int32_t pred = fire_predict(c, pd, coef);
                err = (int32_t)sx((int64_t)delta - pred, w);
                if (r & 1) {  /* learning_downsample = 2: odd rows only (sprintz_xff_rle.cpp:240) */
                    int32_t g = (err == 0) ? 0 : (err < 0 ? (int32_t)sx(-(int64_t)pd, w) : pd); /* util.h:63-68 */
                    gs = (int32_t)sx((int64_t)gs + g, w);
                }
            }
            uint32_t z = (((uint32_t)err << 1) ^ (uint32_t)(err >> (w - 1))) & wmask; /* bitpack.h:302 */
            zz[r * D + d] = z;
            mask |= z;
            pv = v; pd = delta;
        }
        st->prev_val[d] = pv; st->prev_delta[d] = pd;
        if (c->fire)  /* counter += grad_sum >> 2, counter_t wraps (sprintz_xff_rle.cpp:273-275) */
            st->counter[d] = (int32_t)sx((int64_t)st->counter[d] + (gs >> c->gshift), c->cb);
        nbits[d] = (uint8_t)nbits_rule(c, mask);
        total += nbits[d];
    }
    return total;
}

/* ---- little-endian LSB-first bit writer / reader over a byte buffer ---- */
static void put_bits(uint8_t* base, uint64_t bitpos, uint32_t v, int n)
{
    for (int i = 0; i < n; i++) {
        if ((v >> i) & 1u) base[(bitpos + i) >> 3] |= (uint8_t)(1u << ((bitpos + i) & 7));
    }
}
static uint32_t get_bits(const uint8_t* base, uint64_t bitpos, int n)
{
    uint32_t v = 0;
    for (int i = 0; i < n; i++)
        v |= (uint32_t)((base[(bitpos + i) >> 3] >> ((bitpos + i) & 7)) & 1u) << i;
    return v;
}

typedef struct {
    uint8_t* dest;      /* write cursor */
    uint8_t* hdr;       /* current group's header */
    uint32_t nslots;
    uint32_t ngroups;
} emitter_t;

static void begin_slot(const cfg_t* c, emitter_t* e)
{
    if ((e->nslots & 1u) == 0) {           /* a new pair of slots == a new group */
        e->hdr = e->dest;
        memset(e->hdr, 0, c->H);
        e->dest += c->H;
        e->ngroups++;
    }
}

/* run slot: header fields all zero, payload = 1 or 2 length bytes (sprintz_delta_rle.cpp:238-245) */
static void emit_run(const cfg_t* c, emitter_t* e, uint32_t run)
{
    begin_slot(c, e);
    *e->dest++ = (uint8_t)((run & 0x7f) | (run > 0x7f ? 0x80 : 0));
    if (run > 0x7f) *e->dest++ = (uint8_t)(run >> 7);
    e->nslots++;
}

/* data slot: header nibbles + packed payload.
 * high-dim payload: 8 rows, each the D fields LSB-first, byte padded (sprintz_delta_rle.cpp:204-212,345-381)
 * low-dim payload : per dim its 8 fields LSB-first == nbits[d] bytes (sprintz_delta_lowdim.cpp:306-356) */
static void emit_data(const cfg_t* c, emitter_t* e, const uint32_t* zz, const uint8_t* nbits, uint32_t total)
{
    const int D = c->D;
    begin_slot(c, e);
    uint32_t s = e->nslots & 1u;
    for (int d = 0; d < D; d++) {
        uint32_t field = nbits[d] - (nbits[d] == c->w);   /* w stored as w-1 (sprintz_delta_rle.cpp:199) */
        put_bits(e->hdr, (uint64_t)(s * (uint32_t)D + (uint32_t)d) * c->hb, field, c->hb);
    }
    if (c->low) {
        memset(e->dest, 0, (size_t)total * (c->R / 8));    /* a column is R fields of nbits bits = nbits * R / 8 bytes */
        uint64_t bit = 0;
        for (int d = 0; d < D; d++)
            for (int r = 0; r < c->R; r++) { put_bits(e->dest, bit, zz[r * D + d], nbits[d]); bit += nbits[d]; }
        e->dest += (size_t)total * (c->R / 8);
    } else {
        uint32_t row_bytes = (total + 7) / 8;
        memset(e->dest, 0, (size_t)row_bytes * c->R);
        for (int r = 0; r < c->R; r++) {
            uint64_t bit = (uint64_t)r * row_bytes * 8;
            for (int d = 0; d < D; d++) { put_bits(e->dest, bit, zz[r * D + d], nbits[d]); bit += nbits[d]; }
        }
        e->dest += (size_t)row_bytes * c->R;
    }
    e->nslots++;
}

static void write_meta(uint8_t* p, uint16_t ndims, uint32_t ngroups, uint16_t rem)  /* format.h:35-45 */
{
    p[0] = (uint8_t)ngroups; p[1] = (uint8_t)(ngroups >> 8); p[2] = (uint8_t)(ngroups >> 16); p[3] = (uint8_t)(ngroups >> 24);
    p[4] = (uint8_t)rem; p[5] = (uint8_t)(rem >> 8);
    p[6] = (uint8_t)ndims; p[7] = (uint8_t)(ndims >> 8);
}

/* block_rows: rows per block.  8 is the reference's constant; 16 / 32 / 64 apply the reference's own derived constants
 * (min_data_size = 8 * block_sz * group_sz_blocks, shift_to_get_mean = log2_block_sz - log2_learning_downsample) to a
 * larger block_sz -- BASELINE config 5's block-size axis, which no reference build can check ("parity unpinned"). */
static int set_block_rows(cfg_t* c, int block_rows)
{
    int lg = 0;
    while ((1 << lg) < block_rows) lg++;
    if ((1 << lg) != block_rows || block_rows < 8 || block_rows > 64) return -1;
    c->R = block_rows; c->gshift = lg - 1;
    if (block_rows > 8) c->low = 0;    /* the column-major low-dim layout is tied to 8-row blocks: larger blocks use the
                                        * high-dim layout, width rule, coefficient rule and continue-run test at every ndims */
    return 0;
}

int64_t oracle_compress_bs(int predictor, int elem_sz, const void* src_, uint32_t len, void* dest_,
                           uint16_t ndims, int write_size, uint64_t* true_nbytes, int block_rows)
{
    if (ndims == 0) return -1;                 /* sprintz.cpp:71 */
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* const orig = (uint8_t*)dest_;
    cfg_t c = make_cfg(predictor, elem_sz, ndims);
    if (set_block_rows(&c, block_rows)) return -1;
    const int D = c.D;
    const int BLOCK_ROWS = c.R;
    const int64_t B = (int64_t)BLOCK_ROWS * D, G = 2 * B, L = len;
    uint8_t* dest = orig;

    if (len < (uint32_t)(8 * BLOCK_ROWS * 2)) {    /* min_data_size, sprintz_delta_rle.cpp:71,101-109 */
        if (write_size) { write_meta(dest, ndims, 0, (uint16_t)len); dest += 8; }
        memcpy(dest, src, (size_t)len * c.esz);
        dest += (size_t)len * c.esz;
        if (true_nbytes) *true_nbytes = (uint64_t)(dest - orig);
        return (int64_t)(dest - orig) / c.esz;
    }
    if (write_size) dest += 8;

    state_t st;
    st.prev_val = (uint32_t*)calloc(D, sizeof(uint32_t));
    st.prev_delta = (int32_t*)calloc(D, sizeof(int32_t));
    st.counter = (int32_t*)calloc(D, sizeof(int32_t));
    uint32_t* zz = (uint32_t*)malloc(sizeof(uint32_t) * BLOCK_ROWS * D);
    uint8_t* nbits = (uint8_t*)malloc(D);

    /* continue-run test: `<` for delta and all low-dim codecs, `<=` for high-dim FIRE
     * (sprintz_delta_rle.cpp:226, sprintz_delta_lowdim.cpp:190, sprintz_xff_lowdim.cpp:234
     *  vs sprintz_xff_rle.cpp:362) */
    const int strict = !(c.fire && !c.low);

    emitter_t e; e.dest = dest; e.hdr = NULL; e.nslots = 0; e.ngroups = 0;
    int64_t i = 0;          /* next block index */
    uint32_t run = 0;
    int outer = 1;          /* next block read is gated by the outer-loop test (sprintz_delta_rle.cpp:138) */
    int64_t tail;
    for (;;) {
        if (outer && i * B > L - G) { tail = i * B; break; }
        uint32_t total = encode_block(&c, &st, src, (uint64_t)(i * B), zz, nbits);
        int zero = (total == 0);
    again:
        if (zero && run < MAX_RUN_NBLOCKS) {
            run++; i++;
            if (strict ? (i * B < L - G) : (i * B <= L - G)) { outer = 0; continue; }
            /* not enough data to finish the group: flush the run, pad the pair (:249-252) */
            emit_run(&c, &e, run); run = 0;
            if (e.nslots & 1u) { *e.dest++ = 0; e.nslots++; }
            tail = i * B; break;
        }
        if (run > 0) {      /* a run just ended (:263-312) */
            emit_run(&c, &e, run); run = 0;
            if (zero) goto again;              /* only when the run hit MAX_RUN_NBLOCKS (:305-308) */
        }
        emit_data(&c, &e, zz, nbits, total);
        i++;
        outer = ((e.nslots & 1u) == 0);
    }
    dest = e.dest;

    uint32_t remaining = (uint32_t)(L - tail);
    /* low-dim truncates the tail length itself to u16, high-dim only its stored copy
     * (sprintz_delta_lowdim.cpp:376-381 vs sprintz_delta_rle.cpp:398-402) */
    uint32_t copy_len = c.low ? (uint16_t)remaining : remaining;
    if (write_size) write_meta(orig, ndims, e.ngroups, (uint16_t)remaining);
    memcpy(dest, src + (size_t)tail * c.esz, (size_t)copy_len * c.esz);
    dest += (size_t)copy_len * c.esz;

    free(st.prev_val); free(st.prev_delta); free(st.counter); free(zz); free(nbits);
    if (true_nbytes) *true_nbytes = (uint64_t)(dest - orig);
    return (int64_t)(dest - orig) / c.esz;      /* pointer difference in int_t units */
}

int64_t oracle_compress(int predictor, int elem_sz, const void* src, uint32_t len, void* dest,
                        uint16_t ndims, int write_size, uint64_t* true_nbytes)
{
    return oracle_compress_bs(predictor, elem_sz, src, len, dest, ndims, write_size, true_nbytes, BLOCK_SZ);
}

/* ------------------------------------------------------------------ decode */

int64_t oracle_decompress_raw_bs(int predictor, int elem_sz, const void* src_, void* dest_,
                                 uint16_t ndims, uint32_t ngroups, uint16_t remaining_len, int block_rows)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* out = (uint8_t*)dest_;
    if (ngroups == 0 && remaining_len < (uint32_t)(8 * block_rows * 2)) {   /* sprintz_xff_rle.cpp:623-629 */
        memcpy(out, src, (size_t)remaining_len * elem_sz);
        return remaining_len;
    }
    if (ndims == 0) return 0;                              /* sprintz_delta_rle.cpp:475-478 */
    cfg_t c = make_cfg(predictor, elem_sz, ndims);
    if (set_block_rows(&c, block_rows)) return -1;
    const int BLOCK_ROWS = c.R;
    const int D = c.D, w = c.w;
    const uint32_t wmask = (1u << w) - 1u;

    state_t st;
    st.prev_val = (uint32_t*)calloc(D, sizeof(uint32_t));
    st.prev_delta = (int32_t*)calloc(D, sizeof(int32_t));
    st.counter = (int32_t*)calloc(D, sizeof(int32_t));
    uint8_t* nbits = (uint8_t*)malloc(D);
    int32_t* gs = (int32_t*)malloc(sizeof(int32_t) * D);
    int32_t* coef = (int32_t*)malloc(sizeof(int32_t) * D);

    uint64_t o = 0;     /* output element cursor */
    for (uint32_t g = 0; g < ngroups; g++) {
        const uint8_t* hdr = src; src += c.H;
        for (uint32_t s = 0; s < 2; s++) {
            uint32_t total = 0;
            for (int d = 0; d < D; d++) {
                uint32_t f = get_bits(hdr, (uint64_t)(s * (uint32_t)D + (uint32_t)d) * c.hb, c.hb);
                nbits[d] = (uint8_t)(f == (uint32_t)(w - 1) ? (uint32_t)w : f);
                total += nbits[d];
            }
            if (total == 0) {   /* run slot (sprintz_delta_rle.cpp:640-667, sprintz_xff_rle.cpp:828-958) */
                uint32_t n = src[0] & 0x7f;
                if (src[0] & 0x80) { n |= (uint32_t)src[1] << 7; src += 2; } else { src += 1; }
                for (int d = 0; d < D; d++) coef[d] = c.fire ? fire_coef(&c, st.counter[d]) : 0;
                for (uint64_t r = 0; r < (uint64_t)n * BLOCK_ROWS; r++) {
                    for (int d = 0; d < D; d++) {
                        if (c.fire) {   /* free-run the predictor with err = 0; counters untouched */
                            int32_t pred = fire_predict(&c, st.prev_delta[d], coef[d]);
                            st.prev_delta[d] = pred;
                            st.prev_val[d] = (st.prev_val[d] + (uint32_t)pred) & wmask;
                        }
                        store_elem(out, o + (uint64_t)d, c.esz, st.prev_val[d]);
                    }
                    o += (uint64_t)D;
                }
                continue;
            }
            uint32_t row_bytes = (total + 7) / 8;
            for (int d = 0; d < D; d++) { gs[d] = 0; coef[d] = c.fire ? fire_coef(&c, st.counter[d]) : 0; }
            uint64_t colbit = 0;    /* low-dim: start bit of dim d's column */
            uint32_t rowoff = 0;    /* high-dim: bit offset of dim d inside a row */
            for (int d = 0; d < D; d++) {
                for (int r = 0; r < BLOCK_ROWS; r++) {
                    uint64_t bit = c.low ? colbit + (uint64_t)r * nbits[d]
                                         : (uint64_t)r * row_bytes * 8 + rowoff;
                    uint32_t u = get_bits(src, bit, nbits[d]);
                    int32_t err = (int32_t)sx((int64_t)((u >> 1) ^ (0u - (u & 1u))), w);   /* bitpack.h:303 */
                    int32_t delta = err;
                    if (c.fire) {
                        int32_t pd = st.prev_delta[d];
                        int32_t pred = fire_predict(&c, pd, coef[d]);
                        delta = (int32_t)sx((int64_t)err + pred, w);
                        if (r & 1) {
                            int32_t gr = (err == 0) ? 0 : (err < 0 ? (int32_t)sx(-(int64_t)pd, w) : pd);
                            gs[d] = (int32_t)sx((int64_t)gs[d] + gr, w);
                        }
                    }
                    st.prev_delta[d] = delta;
                    st.prev_val[d] = (st.prev_val[d] + (uint32_t)delta) & wmask;
                    store_elem(out, o + (uint64_t)r * D + d, c.esz, st.prev_val[d]);
                }
                if (c.fire) st.counter[d] = (int32_t)sx((int64_t)st.counter[d] + (gs[d] >> c.gshift), c.cb);
                colbit += (uint64_t)BLOCK_ROWS * nbits[d];
                rowoff += nbits[d];
            }
            src += c.low ? total * (uint32_t)(BLOCK_ROWS / 8) : row_bytes * (uint32_t)BLOCK_ROWS;
            o += (uint64_t)BLOCK_ROWS * D;
        }
    }
    memcpy(out + o * c.esz, src, (size_t)remaining_len * c.esz);
    o += remaining_len;

    free(st.prev_val); free(st.prev_delta); free(st.counter); free(nbits); free(gs); free(coef);
    return (int64_t)o;
}

int64_t oracle_decompress_raw(int predictor, int elem_sz, const void* src, void* dest,
                              uint16_t ndims, uint32_t ngroups, uint16_t remaining_len)
{
    return oracle_decompress_raw_bs(predictor, elem_sz, src, dest, ndims, ngroups, remaining_len, BLOCK_SZ);
}

int64_t oracle_decompress_bs(int predictor, int elem_sz, const void* src_, void* dest, int block_rows)
{
    const uint8_t* p = (const uint8_t*)src_;
    uint32_t ngroups = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    uint16_t rem = (uint16_t)(p[4] | (p[5] << 8));
    uint16_t ndims = (uint16_t)(p[6] | (p[7] << 8));
    return oracle_decompress_raw_bs(predictor, elem_sz, p + 8, dest, ndims, ngroups, rem, block_rows);
}

int64_t oracle_decompress(int predictor, int elem_sz, const void* src_, void* dest)
{
    const uint8_t* p = (const uint8_t*)src_;      /* format.h:47-62 */
    uint32_t ngroups = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    uint16_t rem = (uint16_t)(p[4] | (p[5] << 8));
    uint16_t ndims = (uint16_t)(p[6] | (p[7] << 8));
    return oracle_decompress_raw(predictor, elem_sz, p + 8, dest, ndims, ngroups, rem);
}


/* ------------------------------------------------------------------ non-RLE sibling codecs
 *
 * compress_rowmajor_delta_{8b,16b} / compress_rowmajor_{8b,16b} and their decoders
 * (sprintz_delta.h:27-44; sprintz_delta.cpp:64-392 plain, :776-1017 delta, decoders :394-770, :1021-1391).
 * Same block / group / header / row-major payload layout as the RLE high-dim codecs for EVERY ndims (there is
 * no low-dim dispatch here), but:
 *   - 6-byte metadata [u32 len][u16 ndims] (format.h:65-85, kMetaDataLenBytesSimple)
 *   - ngroups = len / (16*ndims) is implied; every block is a data slot, a block whose widths are all zero
 *     simply has zero payload bytes; no run slots
 *   - kind 1: zigzag(delta) with the delta carried across blocks; kind 2: the raw values, no zigzag
 *   - the width rule is the high-dim one (7->8, 15->16) for all ndims
 */
#define KIND_DELTA 1
#define KIND_PLAIN 2
#define KIND_XFF8  3   /* compress8b_rowmajor_xff (sprintz_xff.cpp:35-303): FIRE exactly as the 8-bit RLE codec
                        * (coef = (counter >> 5) << 4, odd-row gradients, counter += grad_sum >> 2), no run slots,
                        * 8-byte metadata [u48 len][u16 ndims] (:64-69) */

static uint32_t rowmajor_block(const cfg_t* c, int kind, uint32_t* prev_val, int32_t* prev_delta, int32_t* counter,
                               const uint8_t* src, uint64_t first_elem, uint32_t* zz, uint8_t* nbits)
{
    const int D = c->D, w = c->w;
    const uint32_t wmask = (1u << w) - 1u;
    uint32_t total = 0;
    for (int d = 0; d < D; d++) {
        uint32_t mask = 0, pv = prev_val[d];
        int32_t pd = prev_delta[d], gs = 0;
        const int32_t coef = (kind == KIND_XFF8) ? fire_coef(c, counter[d]) : 0;
        for (int r = 0; r < BLOCK_SZ; r++) {
            uint32_t v = load_elem(src, first_elem + (uint64_t)r * D + d, c->esz);
            uint32_t z = v;
            if (kind != KIND_PLAIN) {
                int32_t delta = (int32_t)sx((int64_t)v - (int64_t)pv, w);
                int32_t err = delta;
                if (kind == KIND_XFF8) {
                    err = (int32_t)sx((int64_t)delta - fire_predict(c, pd, coef), w);
                    if (r & 1) {
                        int32_t g = (err == 0) ? 0 : (err < 0 ? (int32_t)sx(-(int64_t)pd, w) : pd);
                        gs = (int32_t)sx((int64_t)gs + g, w);
                    }
                }
                z = (((uint32_t)err << 1) ^ (uint32_t)(err >> (w - 1))) & wmask;
                pd = delta;
            }
            zz[r * D + d] = z;
            mask |= z;
            pv = v;
        }
        prev_val[d] = pv; prev_delta[d] = pd;
        if (kind == KIND_XFF8) counter[d] = (int32_t)sx((int64_t)counter[d] + (gs >> 2), c->cb);
        nbits[d] = (uint8_t)nbits_rule(c, mask);
        total += nbits[d];
    }
    return total;
}

int64_t oracle_compress_rowmajor(int kind, int elem_sz, const void* src_, uint32_t len, void* dest_,
                                 uint16_t ndims, int write_size, uint64_t* true_nbytes)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* const orig = (uint8_t*)dest_;
    uint8_t* dest = orig;
    cfg_t c = make_cfg(kind == KIND_XFF8, elem_sz, ndims ? ndims : 1);
    c.low = 0;                                            /* one layout / width rule for every ndims */
    const int D = c.D;
    if (write_size) {                                     /* format.h:65-73; sprintz_xff.cpp:64-69 for kind 3 */
        dest[0] = (uint8_t)len; dest[1] = (uint8_t)(len >> 8); dest[2] = (uint8_t)(len >> 16); dest[3] = (uint8_t)(len >> 24);
        if (kind == KIND_XFF8) { dest[4] = 0; dest[5] = 0; dest += 2; }
        dest[4] = (uint8_t)ndims; dest[5] = (uint8_t)(ndims >> 8);
        dest += 6;
    }
    uint64_t tail = 0;
    if (len >= MIN_DATA_SIZE && ndims != 0) {             /* sprintz_delta.cpp:817-821 */
        const uint64_t B = (uint64_t)BLOCK_SZ * D, G = 2 * B;
        const uint64_t ngroups = len / G;
        uint32_t* prev = (uint32_t*)calloc(D, sizeof(uint32_t));
        int32_t* pdel = (int32_t*)calloc(D, sizeof(int32_t));
        int32_t* cnt = (int32_t*)calloc(D, sizeof(int32_t));
        uint32_t* zz = (uint32_t*)malloc(sizeof(uint32_t) * BLOCK_SZ * D);
        uint8_t* nbits = (uint8_t*)malloc(D);
        for (uint64_t g = 0; g < ngroups; g++) {
            uint8_t* hdr = dest;
            memset(hdr, 0, c.H);
            dest += c.H;
            for (uint32_t b = 0; b < 2; b++) {
                uint32_t total = rowmajor_block(&c, kind, prev, pdel, cnt, src, (g * 2 + b) * B, zz, nbits);
                for (int d = 0; d < D; d++) {
                    uint32_t field = nbits[d] - (nbits[d] == c.w);
                    put_bits(hdr, (uint64_t)(b * (uint32_t)D + (uint32_t)d) * c.hb, field, c.hb);
                }
                uint32_t row_bytes = (total + 7) / 8;
                memset(dest, 0, (size_t)row_bytes * BLOCK_SZ);
                for (int r = 0; r < BLOCK_SZ; r++) {
                    uint64_t bit = (uint64_t)r * row_bytes * 8;
                    for (int d = 0; d < D; d++) { put_bits(dest, bit, zz[r * D + d], nbits[d]); bit += nbits[d]; }
                }
                dest += (size_t)row_bytes * BLOCK_SZ;
            }
        }
        tail = ngroups * G;
        free(prev); free(pdel); free(cnt); free(zz); free(nbits);
    }
    memcpy(dest, src + tail * c.esz, (size_t)(len - tail) * c.esz);
    dest += (size_t)(len - tail) * c.esz;
    if (true_nbytes) *true_nbytes = (uint64_t)(dest - orig);
    return (int64_t)(dest - orig) / c.esz;
}

int64_t oracle_decompress_rowmajor(int kind, int elem_sz, const void* src_, void* dest_)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* out = (uint8_t*)dest_;
    const uint32_t len = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    if (kind == KIND_XFF8) src += 2;                       /* u48 length (we only carry 32 bits of it) */
    const uint16_t ndims = (uint16_t)(src[4] | (src[5] << 8));
    src += 6;
    if (len < MIN_DATA_SIZE) { memcpy(out, src, (size_t)len * elem_sz); return len; }   /* sprintz_delta.cpp:1056-1060 */
    if (ndims == 0) return 0;
    cfg_t c = make_cfg(kind == KIND_XFF8, elem_sz, ndims);
    c.low = 0;
    const int D = c.D, w = c.w;
    const uint32_t wmask = (1u << w) - 1u;
    const uint64_t B = (uint64_t)BLOCK_SZ * D, G = 2 * B, ngroups = len / G;
    uint32_t* prev = (uint32_t*)calloc(D, sizeof(uint32_t));
    int32_t* pdel = (int32_t*)calloc(D, sizeof(int32_t));
    int32_t* cnt = (int32_t*)calloc(D, sizeof(int32_t));
    int32_t* gsum = (int32_t*)calloc(D, sizeof(int32_t));
    int32_t* coefs = (int32_t*)calloc(D, sizeof(int32_t));
    uint8_t* nbits = (uint8_t*)malloc(D);
    uint64_t o = 0;
    for (uint64_t g = 0; g < ngroups; g++) {
        const uint8_t* hdr = src; src += c.H;
        for (uint32_t b = 0; b < 2; b++) {
            uint32_t total = 0;
            for (int d = 0; d < D; d++) { gsum[d] = 0; coefs[d] = (kind == KIND_XFF8) ? fire_coef(&c, cnt[d]) : 0; }
            for (int d = 0; d < D; d++) {
                uint32_t f = get_bits(hdr, (uint64_t)(b * (uint32_t)D + (uint32_t)d) * c.hb, c.hb);
                nbits[d] = (uint8_t)(f == (uint32_t)(w - 1) ? (uint32_t)w : f);
                total += nbits[d];
            }
            uint32_t row_bytes = (total + 7) / 8;
            for (int r = 0; r < BLOCK_SZ; r++) {
                uint64_t bit = (uint64_t)r * row_bytes * 8;
                for (int d = 0; d < D; d++) {
                    uint32_t u = get_bits(src, bit, nbits[d]); bit += nbits[d];
                    uint32_t v = u;
                    if (kind != KIND_PLAIN) {
                        int32_t err = (int32_t)sx((int64_t)((u >> 1) ^ (0u - (u & 1u))), w);
                        int32_t delta = err;
                        if (kind == KIND_XFF8) {
                            delta = (int32_t)sx((int64_t)err + fire_predict(&c, pdel[d], coefs[d]), w);
                            if (r & 1) {
                                int32_t gr = (err == 0) ? 0 : (err < 0 ? (int32_t)sx(-(int64_t)pdel[d], w) : pdel[d]);
                                gsum[d] = (int32_t)sx((int64_t)gsum[d] + gr, w);
                            }
                            pdel[d] = delta;
                        }
                        v = (prev[d] + (uint32_t)delta) & wmask;
                        prev[d] = v;
                    }
                    store_elem(out, o + (uint64_t)r * D + d, c.esz, v);
                }
            }
            if (kind == KIND_XFF8)
                for (int d = 0; d < D; d++) cnt[d] = (int32_t)sx((int64_t)cnt[d] + (gsum[d] >> 2), c.cb);
            src += (size_t)row_bytes * BLOCK_SZ;
            o += B;
        }
    }
    memcpy(out + o * c.esz, src, (size_t)(len - o) * c.esz);
    free(prev); free(pdel); free(cnt); free(gsum); free(coefs); free(nbits);
    return (int64_t)len;
}

/* ------------------------------------------------------------------ stand-alone forecaster transforms
 *
 * encode/decode_{delta,doubledelta}_rowmajor_{8b,16b} (delta.h:17-67, delta.cpp:34-744): same-size in/out,
 * optional 6-byte metadata [u32 len][u16 ndims] (format.h:65-85).  Behaviourally (whatever vector / serial path the
 * reference takes) with x[i] = 0 for i < 0, all arithmetic mod 2^w:
 *   delta       : y[i] = x[i] - x[i-D]
 *   doubledelta : y[i] = (x[i] - x[i-D]) - (x[i-D] - x[i-2D])
 * Returns len + the metadata length in ELEMENTS (6 for 8-bit, 3 for 16-bit), like the reference.  order: 1 | 2. */
int64_t oracle_encode_forecast(int order, int elem_sz, const void* src_, uint32_t len, void* dest_, uint16_t ndims,
                               int write_size)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dest = (uint8_t*)dest_;
    const int w = 8 * elem_sz;
    const uint32_t wmask = (1u << w) - 1u;
    uint32_t meta = 0;
    if (write_size) {
        dest[0] = (uint8_t)len; dest[1] = (uint8_t)(len >> 8); dest[2] = (uint8_t)(len >> 16); dest[3] = (uint8_t)(len >> 24);
        dest[4] = (uint8_t)ndims; dest[5] = (uint8_t)(ndims >> 8);
        dest += 6; meta = 6 / (uint32_t)elem_sz;
    }
    const uint64_t D = ndims;
    for (uint64_t i = 0; i < len; i++) {
        uint32_t x0 = load_elem(src, i, elem_sz);
        uint32_t x1 = (D && i >= D) ? load_elem(src, i - D, elem_sz) : 0;
        uint32_t x2 = (D && i >= 2 * D) ? load_elem(src, i - 2 * D, elem_sz) : 0;
        uint32_t y = (order == 1) ? (x0 - x1) : (x0 - 2 * x1 + x2);
        store_elem(dest, i, elem_sz, y & wmask);
    }
    return (int64_t)len + meta;
}

/* src points at the transformed elements (no metadata) */
int64_t oracle_decode_forecast(int order, int elem_sz, const void* src_, uint32_t len, void* dest_, uint16_t ndims)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* out = (uint8_t*)dest_;
    const int w = 8 * elem_sz;
    const uint32_t wmask = (1u << w) - 1u;
    const uint64_t D = ndims;
    if (order == 2 && ndims == 0) return 0;                /* delta.cpp:631 */
    for (uint64_t i = 0; i < len; i++) {
        uint32_t y = load_elem(src, i, elem_sz);
        uint32_t x1 = (D && i >= D) ? load_elem(out, i - D, elem_sz) : 0;
        uint32_t x2 = (D && i >= 2 * D) ? load_elem(out, i - 2 * D, elem_sz) : 0;
        uint32_t x = (order == 1) ? (y + x1) : (y + 2 * x1 - x2);
        store_elem(out, i, elem_sz, x & wmask);
    }
    return (int64_t)len;
}

/* ------------------------------------------------------------------ stand-alone FIRE transform
 *
 * encode/decode_xff_rowmajor_{8b,16b} (predict.h:15-30, predict.cpp:56-527): same-size in/out like the delta
 * transforms, but the forecaster is an older prototype of FIRE whose arithmetic differs from the codec's
 * (sprintz_xff_rle.cpp) in several observable ways.  Restated per dim d (the AVX2 lanes beyond ndims only touch bytes
 * that are rewritten later):
 *   blocks    nblocks = (len / ndims) / 8, reduced so that the vector stores of the last block stay inside the
 *             buffer (predict.cpp:99-106); every element after the block part -- whole rows included -- is plain DELTA
 *             coded with lag ndims (:268-279), and when there is no block at all the first ndims elements are copied.
 *   8-bit     counter int16, coef = (counter >> 5) << 4 (learning_shift 1); the prediction of an EVEN dim multiplies the
 *             previous delta as an UNSIGNED byte (and_si256 with 0x00ff, :163-164), of an odd dim as a signed one;
 *             both keep bits 8..15 of the 16-bit product; the gradient sum wraps at 8 bits ("faster, but can
 *             overflow", :189); counter += (int8)gsum >> 2.
 *   16-bit    counter int32, coef = low16((counter >> 15)) << 12 as int16 (learning_shift 3); prediction =
 *             (mulhi(prev_delta, coef) << 2) wrapped to 16 bits; gradient sum wraps at 16 bits; counter += (int16)gsum >> 2.
 * TEST INFRASTRUCTURE ONLY.  No product kernel exists for this transform yet. */
static int32_t xffp_nblocks(uint32_t len, uint32_t ndims, int elem_sz)
{
    const uint32_t vector_sz = 32u / (uint32_t)elem_sz;
    const uint32_t bse = 8u * ndims;
    int32_t nblocks = (int32_t)((len / ndims) / 8u);
    const uint32_t overrun = vector_sz - (ndims % vector_sz);
    const uint32_t trailing = len % bse;
    if (overrun > trailing) {
        nblocks -= (int32_t)((overrun + bse - 1) / bse);
        if (nblocks < 0) nblocks = 0;
    }
    return nblocks;
}

static int32_t xffp_predict(int elem_sz, uint32_t d, int32_t pd, int32_t counter)
{
    if (elem_sz == 1) {
        const int32_t coef = (int32_t)(int16_t)(((int16_t)counter >> 5) << 4);
        const uint32_t a = (d & 1u) ? (uint32_t)(int32_t)(int8_t)pd : (uint32_t)(uint8_t)pd;
        const uint32_t prod = (a * (uint32_t)coef) & 0xffffu;
        return (int32_t)(int8_t)(prod >> 8);
    }
    const int32_t coef = (int32_t)(int16_t)((((uint32_t)(counter >> 15)) << 12) & 0xffffu);
    const int32_t hi = ((int32_t)(int16_t)pd * coef) >> 16;
    return (int32_t)(int16_t)((uint32_t)hi << 2);
}

static int32_t xffp_sign(int elem_sz, int32_t pd, int32_t err)   /* psignb / psignw(prev_delta, err) */
{
    int32_t g = err == 0 ? 0 : (err < 0 ? -pd : pd);
    return elem_sz == 1 ? (int32_t)(int8_t)g : (int32_t)(int16_t)g;
}

int64_t oracle_encode_xff_forecast(int elem_sz, const void* src_, uint32_t len, void* dest_, uint16_t ndims, int write_size)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dest = (uint8_t*)dest_;
    const int w = 8 * elem_sz;
    const uint32_t wmask = (1u << w) - 1u;
    uint32_t meta = 0;
    if (ndims == 0) return -1;
    if (write_size) {
        dest[0] = (uint8_t)len; dest[1] = (uint8_t)(len >> 8); dest[2] = (uint8_t)(len >> 16); dest[3] = (uint8_t)(len >> 24);
        dest[4] = (uint8_t)ndims; dest[5] = (uint8_t)(ndims >> 8);
        dest += 6; meta = 6 / (uint32_t)elem_sz;
    }
    const uint32_t D = ndims;
    const int32_t nblocks = xffp_nblocks(len, D, elem_sz);
    for (uint32_t d = 0; d < D; d++) {
        uint32_t pv = 0; int32_t pd = 0, counter = 0;
        for (int32_t b = 0; b < nblocks; b++) {
            int32_t gsum = 0;
            for (uint32_t i = 0; i < 8; i++) {
                const uint64_t at = ((uint64_t)b * 8 + i) * D + d;
                const uint32_t x = load_elem(src, at, elem_sz);
                const int32_t delta = (int32_t)sx((int32_t)((x - pv) & wmask), w);
                const int32_t pred = xffp_predict(elem_sz, d, pd, counter);
                const int32_t err = (int32_t)sx((int32_t)(((uint32_t)delta - (uint32_t)pred) & wmask), w);
                if (i & 1u) gsum = (int32_t)sx((int32_t)(((uint32_t)gsum + (uint32_t)xffp_sign(elem_sz, pd, err)) & wmask), w);
                store_elem(dest, at, elem_sz, (uint32_t)err & wmask);
                pd = delta; pv = x;
            }
            if (elem_sz == 1) counter = (int32_t)(int16_t)(counter + (gsum >> 2));
            else counter = (int32_t)((uint32_t)counter + (uint32_t)(gsum >> 2));
        }
    }
    /* everything after the block part: delta with lag ndims (predict.cpp:268-279) */
    uint64_t i0 = (uint64_t)nblocks * 8 * D;
    if (nblocks == 0) {
        const uint64_t cpy = D < len ? D : len;
        for (uint64_t i = 0; i < cpy; i++) store_elem(dest, i, elem_sz, load_elem(src, i, elem_sz));
        i0 = D;
    }
    for (uint64_t i = i0; i < len; i++)
        store_elem(dest, i, elem_sz, (load_elem(src, i, elem_sz) - load_elem(src, i - D, elem_sz)) & wmask);
    return (int64_t)len + meta;
}

/* src points at the transformed elements (no metadata) */
int64_t oracle_decode_xff_forecast(int elem_sz, const void* src_, uint32_t len, void* dest_, uint16_t ndims)
{
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* out = (uint8_t*)dest_;
    const int w = 8 * elem_sz;
    const uint32_t wmask = (1u << w) - 1u;
    if (ndims == 0) return 0;                              /* predict.cpp:318 */
    const uint32_t D = ndims;
    const int32_t nblocks = xffp_nblocks(len, D, elem_sz);
    for (uint32_t d = 0; d < D; d++) {
        uint32_t pv = 0; int32_t pd = 0, counter = 0;
        for (int32_t b = 0; b < nblocks; b++) {
            int32_t gsum = 0;
            for (uint32_t i = 0; i < 8; i++) {
                const uint64_t at = ((uint64_t)b * 8 + i) * D + d;
                const int32_t err = (int32_t)sx((int32_t)load_elem(src, at, elem_sz), w);
                const int32_t pred = xffp_predict(elem_sz, d, pd, counter);
                if (i & 1u) gsum = (int32_t)sx((int32_t)(((uint32_t)gsum + (uint32_t)xffp_sign(elem_sz, pd, err)) & wmask), w);
                const int32_t delta = (int32_t)sx((int32_t)(((uint32_t)err + (uint32_t)pred) & wmask), w);
                pv = (pv + (uint32_t)delta) & wmask;
                store_elem(out, at, elem_sz, pv);
                pd = delta;
            }
            if (elem_sz == 1) counter = (int32_t)(int16_t)(counter + (gsum >> 2));
            else counter = (int32_t)((uint32_t)counter + (uint32_t)(gsum >> 2));
        }
    }
    uint64_t i0 = (uint64_t)nblocks * 8 * D;
    if (nblocks == 0) {
        const uint64_t cpy = D < len ? D : len;
        for (uint64_t i = 0; i < cpy; i++) store_elem(out, i, elem_sz, load_elem(src, i, elem_sz));
        i0 = D;
    }
    for (uint64_t i = i0; i < len; i++)
        store_elem(out, i, elem_sz, (load_elem(src, i, elem_sz) + load_elem(out, i - D, elem_sz)) & wmask);
    return (int64_t)len;
}

/* ------------------------------------------------------------------ query-on-compressed
 *
 * query.hpp:22-29 QueryParams{op = REDUCE_MAX | REDUCE_SUM, materialize}; query_rowmajor_{delta,xff}_rle_{8b,16b}
 * (sprintz_delta_rle_query.cpp:149-192, sprintz_xff_rle_query.cpp:106-137).  The reference's functors never hand
 * their state back (query.hpp:231) and MaxQuery folds every stripe into state[0] (query.hpp:218, "TODO rm"), so
 * there is no reference output to pin against: per SURVEY 8f the oracle is "the reduction of the decoded oracle
 * output" -- decode with oracle_decompress, then for every dim d the unsigned max (op 1) or the 64-bit sum (op 2)
 * of x[i], i % ndims == d.  Parity of the decoded elements themselves stays pinned as before. */
int64_t oracle_query(int predictor, int elem_sz, const void* src, void* dest, int op, uint64_t* result)
{
    const uint8_t* s = (const uint8_t*)src;
    const uint32_t D = (uint32_t)s[6] | ((uint32_t)s[7] << 8);
    const int64_t n = oracle_decompress(predictor, elem_sz, src, dest);
    if (n < 0 || D == 0) return n;
    for (uint32_t d = 0; d < D; d++) result[d] = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint64_t v = load_elem((const uint8_t*)dest, (uint64_t)i, elem_sz);
        uint64_t* r = &result[(uint64_t)i % D];
        if (op == 1) { if (v > *r) *r = v; } else if (op == 2) *r += v;
    }
    return n;
}

uint64_t oracle_max_compressed_bytes(int elem_sz, uint64_t len, uint16_t ndims)
{
    uint64_t D = ndims ? ndims : 1;
    uint64_t hb = (elem_sz == 1) ? 3 : 4;
    uint64_t H = (2 * D * hb + 7) / 8;
    uint64_t nblocks = len / (8 * D);
    return 8 + len * (uint64_t)elem_sz + (nblocks / 2 + 2) * H + 16;
}

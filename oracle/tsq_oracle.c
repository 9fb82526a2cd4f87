/*
 * tsq_oracle.c -- CPU restatement of Turbosqueeze's per-block codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product (turbosqueeze_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (a) against the golden vectors under tests/golden/ that were produced by
 *       the compiled, unmodified reference (oracle/_ref, built by
 *       oracle/Makefile from /root/reference; generator script
 *       tests/golden/make_golden.py), and
 *   (b) byte-for-byte against oracle/_ref itself whenever that library is
 *       present (random / text / periodic inputs, 1 B .. 4 MiB).
 *
 * What is restated (reference file:line):
 *   oracle_encode()      tsq_encode.cpp:48-189 (no-ext) and :192-342 (ext)
 *   oracle_decode()      tsq_decode.cpp:42-126 (no-ext) and :129-315 (ext)
 *   probe_insert()       tsq_encode.cpp:74-79, :162-167
 *   finish_symbol()      tsq_encode.cpp:93-95, :157-159
 *   flush_literals()     tsq_encode.cpp:85-97, :105-117
 *   tail padding         tsq_encode.cpp:176-188
 *
 * Written from the behavioural spec in SURVEY.md section 8(a); it is organised
 * around an explicit emitter state instead of the reference's inline loops.
 *
 * Contract (identical to the reference's memory path, tsq_threads.cpp:109):
 *   - the encoder reads up to 19 bytes past `size` (in[size .. size+18]);
 *   - the encoder stores literals as fixed 16-byte copies, so up to 15 bytes
 *     past the returned length are written and "stale" trailing control/size
 *     bytes take whatever the output slot held before (zero-fill it first).
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define ORC_HASH_BITS 17u
#define ORC_HASH_SLOTS (1u << ORC_HASH_BITS)
#define ORC_MAX_BLOCK (1u << 22)

static inline uint32_t le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t le64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

typedef struct {
    uint8_t *out;
    uint32_t j;          /* next free output byte */
    uint32_t ctl_at;     /* where the open control byte lives */
    uint32_t nib_at;     /* where the open size byte lives */
    uint32_t n_sym;      /* symbols emitted so far */
    uint32_t pair_org;   /* input position at the start of the open pair ("rep_last_i") */
} emitter_t;

/* tsq_encode.cpp:93-95 / :157-159 -- shift the symbol's control bit and length
 * nibble into the open bytes (read-modify-write IN the output buffer, which is
 * why uninitialised bytes can leak), then open new bytes when full. */
static inline void finish_symbol(emitter_t *e, uint32_t is_literal, uint32_t nibble, uint32_t in_pos)
{
    e->n_sym++;
    e->out[e->ctl_at] = (uint8_t)((e->out[e->ctl_at] << 1) | is_literal);
    if ((e->n_sym & 7u) == 0) e->ctl_at = e->j++;
    e->out[e->nib_at] = (uint8_t)((e->out[e->nib_at] << 4) | nibble);
    if ((e->n_sym & 1u) == 0) { e->nib_at = e->j++; e->pair_org = in_pos; }
}

/* tsq_encode.cpp:85-97 -- pending literals [*from, upto) leave as chunks of at
 * most 16, each one a blind 16-byte store. */
static inline void flush_literals(emitter_t *e, const uint8_t *in, uint32_t *from, uint32_t upto)
{
    do {
        uint32_t n = upto - *from; if (n > 16) n = 16;
        uint64_t a = le64(in + *from), b = le64(in + *from + 8);
        memcpy(e->out + e->j, &a, 8); memcpy(e->out + e->j + 8, &b, 8);
        *from += n; e->j += n;
        finish_symbol(e, 1u, n - 1u, *from);
    } while (upto - *from > 0);
}

/* tsq_encode.cpp:74-79 -- hash the 4 bytes at i, fetch the remembered position
 * with that hash (16 bits, re-expanded into the 64 KiB behind i), remember i. */
static inline uint32_t probe_insert(uint16_t *table, const uint8_t *in, uint32_t i, uint32_t *word)
{
    uint32_t w = le32(in + i);
    uint32_t h = (w ^ (w >> 12)) & (ORC_HASH_SLOTS - 1u);
    uint32_t p = table[h];
    p += (p >= (i & 0xFFFFu)) ? (i & 0xFFFF0000u) - 65536u : (i & 0xFFFF0000u);
    table[h] = (uint16_t)i;
    *word = w;
    return p;
}

static inline uint32_t common_prefix(const uint8_t *a, const uint8_t *b, uint32_t cap)
{
    /* tsq_encode.cpp:126-137 (cap 16) / :276-290 (cap 64): whole 8-byte words
     * are compared; the count only keeps growing while a word matched fully. */
    uint32_t k = 0;
    for (;;) {
        uint64_t x = le64(a + k) ^ le64(b + k);
        uint32_t nb = x ? (uint32_t)(__builtin_ctzll(x) >> 3) : 8u;
        k += nb;
        if (nb != 8 || k >= cap) return k;
    }
}

/* length -> (nibble, advance).  tsq_encode.cpp:44-45 (mlen) and :154. */
static inline void match_code(uint32_t k, uint32_t *nibble, uint32_t *advance)
{
    if (k < 17)      { *nibble = k - 1; *advance = k; }          /* 4..16 */
    else if (k < 32) { *nibble = 15;    *advance = 16; }
    else if (k < 48) { *nibble = 0;     *advance = 32; }
    else if (k < 64) { *nibble = 1;     *advance = 48; }
    else             { *nibble = 2;     *advance = 64; }
}

/*
 * Encode one block.  `table` is the caller's 2^17 x u16 scratch (the reference
 * callers zero it before every block: tsq_context.cpp:77-80).
 * Returns the number of bytes produced.
 */
uint32_t oracle_encode(uint16_t *table, const uint8_t *in, uint32_t size, uint8_t *out, uint32_t with_ext)
{
    emitter_t e;
    const uint32_t cap = with_ext ? 64u : 16u;
    uint32_t i = 0, lit_from, word, pos, off;

    out[0] = (uint8_t)size; out[1] = (uint8_t)(size >> 8); out[2] = (uint8_t)(size >> 16);
    e.out = out; e.ctl_at = 3; e.nib_at = 4; e.j = 5; e.n_sym = 0; e.pair_org = 0;

    do {
        lit_from = i;
        /* literal scan, tsq_encode.cpp:70-100 */
        do {
            i++;
            pos = probe_insert(table, in, i, &word);
            off = e.pair_org - pos;                      /* NOT refreshed by the flush below */
            if (i - lit_from > 31) flush_literals(&e, in, &lit_from, i);
        } while (i < size && !(word == le32(in + pos) && (off - 4u) < 0xFFFBu));

        if (i - lit_from > 0) flush_literals(&e, in, &lit_from, i);   /* :103-118 */
        if (!(i < size)) break;

        /* match chain, tsq_encode.cpp:123-170 */
        do {
            uint32_t k = common_prefix(in + i, in + pos, cap);
            uint32_t room = e.pair_org - pos;            /* source must end before the pair */
            uint32_t nibble, adv;
            if (k > room) k = room - 1u;
            if (k < 4) break;
            off = e.pair_org - pos;
            if (!((off - 4u) < 0xFFFBu)) break;
            match_code(k, &nibble, &adv);
            out[e.j++] = (uint8_t)off; out[e.j++] = (uint8_t)(off >> 8);
            i += adv;
            finish_symbol(&e, 0u, nibble, i);
            pos = probe_insert(table, in, i, &word);
            off = e.pair_org - pos;
        } while (i < size - 5u && word == le32(in + pos) && (off - 4u) < 0xFFFBu);
    } while (i < size);

    /* tsq_encode.cpp:176-188 */
    {
        int shifted = 0;
        while (e.n_sym & 7u) {
            out[e.ctl_at] = (uint8_t)((out[e.ctl_at] << 1) | 1u);
            if (!shifted && (e.n_sym & 1u)) { out[e.nib_at] = (uint8_t)(out[e.nib_at] << 4); shifted = 1; }
            e.n_sym++;
        }
    }
    return e.j;
}

/*
 * Decode one block into out[0 .. size).  Unlike the reference (which copies a
 * blind 16/32/48/64 bytes per symbol and relies on slack after the buffer,
 * tsq_decode.cpp:60-90) this restatement clips every store at `size`; the
 * bytes in [0, size) are the same because a match source always lies before
 * the start of its pair (tsq_encode.cpp:139-141).
 * Returns the decoded size, or 0 when the header exceeds 4 MiB (tsq_decode.cpp:53).
 */
uint32_t oracle_decode(const uint8_t *in, uint8_t *out, uint32_t with_ext)
{
    uint32_t size = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16);
    uint32_t i = 3, j = 0;
    if (size > ORC_MAX_BLOCK) return 0;

    while (j < size) {
        uint32_t ctl = in[i++];
        for (int pair = 0; pair < 4; pair++) {
            uint32_t nib = in[i++];
            uint32_t org = j;                            /* offsets count back from here */
            for (int half = 0; half < 2; half++) {
                uint32_t len = (half ? (nib & 15u) : (nib >> 4)) + 1u;
                uint32_t lit = ctl & 0x80u; ctl <<= 1;
                const uint8_t *src;
                if (lit) { src = in + i; i += len; }
                else {
                    uint32_t off = (uint32_t)in[i] | ((uint32_t)in[i + 1] << 8);
                    src = out + (org - off); i += 2;
                    if (with_ext && len <= 3) len = 16u * (len + 1u);   /* tsq_decode.cpp:174-187 */
                }
                for (uint32_t b = 0; b < len && j + b < size; b++) out[j + b] = src[b];
                j += len;
            }
        }
    }
    return size;
}

/* ---- batch helpers under the parity contract of SURVEY.md 8(c) ------------- */

/* Worst case: header 3 + per 16-byte literal (16 + 1/8 ctl + 1/2 size) + slack. */
uint64_t oracle_bound(uint32_t size) { return 5u + (uint64_t)size + (size >> 4) + ((size + 15u) >> 4) + 32u; }

/*
 * Encode `total` bytes of `buf` as consecutive blocks of `block` bytes, in
 * place (block b over-reads into block b+1; `buf` must be followed by >= 32
 * readable bytes).  Block b's stream goes to out + b*stride (slot zero-filled
 * first), its length to sizes[b].
 */
void oracle_encode_blocks(const uint8_t *buf, uint64_t total, uint32_t block, uint8_t *out,
                          uint64_t stride, uint32_t *sizes, uint32_t with_ext)
{
    uint16_t *table = (uint16_t *)malloc(ORC_HASH_SLOTS * sizeof(uint16_t));
    uint64_t nb = (total + block - 1) / block;
    for (uint64_t b = 0; b < nb; b++) {
        uint64_t at = b * (uint64_t)block;
        uint32_t n = (uint32_t)((total - at < block) ? total - at : block);
        memset(table, 0, ORC_HASH_SLOTS * sizeof(uint16_t));
        memset(out + b * stride, 0, stride);
        sizes[b] = oracle_encode(table, buf + at, n, out + b * stride, with_ext);
    }
    free(table);
}

void oracle_decode_blocks(const uint8_t *comp, uint64_t stride, uint64_t nb, uint8_t *out,
                          uint32_t block, uint32_t *sizes, uint32_t with_ext)
{
    for (uint64_t b = 0; b < nb; b++)
        sizes[b] = oracle_decode(comp + b * stride, out + b * (uint64_t)block, with_ext);
}

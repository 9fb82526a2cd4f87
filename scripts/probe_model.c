/* Development helper (CPU): what would alternative encoder designs cost in table requests?
 * Runs the greedy parse of tsqEncodeNoext (tsq_encode.cpp:48-189; same statement as oracle/tsq_oracle.c, instrumented)
 * over 256 KiB blocks of a file and counts, per input byte:
 *   - real probes (the reference's own), hits among them, age of the candidates;
 *   - probes of a lane-per-position window of W positions that starts at the next real probe (the GPU kernel: W = 32);
 *   - how many window probes an exact 'ever written' bitmap of 1 bit per G slots answers;
 *   - how many window probes find their sector among the last K distinct sectors the block touched (an LRU model of
 *     the block's share of L2);
 *   - DRAM requests = window probes not answered by either, plus one per insert.
 * usage: probe_model <file> [block]      (build: gcc -O2 scripts/probe_model.c -o build/probe_model)  */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HB 17u
#define HS (1u << HB)
static inline uint32_t le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t le64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t hash17(uint32_t w) { return (w ^ (w >> 12)) & (HS - 1u); }

/* visited[i] = 1 when the reference probes (and inserts) position i; is_hit[i] = that probe started a match */
static void parse(const uint8_t* in, uint32_t size, uint8_t* visited, uint8_t* is_hit, uint32_t* cand_of)
{
    static uint16_t table[HS];
    memset(table, 0, sizeof table);
    uint32_t i = 0, n = 0, rep = 0, lit_from;
#define PROBE(pos_var, w_var) do { w_var = le32(in + i); uint32_t h_ = hash17(w_var); uint32_t p_ = table[h_]; \
        p_ += (p_ >= (i & 0xFFFFu)) ? (i & 0xFFFF0000u) - 65536u : (i & 0xFFFF0000u); table[h_] = (uint16_t)i; \
        pos_var = p_; visited[i] = 1; cand_of[i] = p_; } while (0)
#define SYM(in_pos) do { n++; if ((n & 1u) == 0) rep = (in_pos); } while (0)
#define FLUSH(upto) do { do { uint32_t c_ = (upto) - lit_from; if (c_ > 16) c_ = 16; lit_from += c_; SYM(lit_from); } while ((upto) - lit_from > 0); } while (0)
    do {
        uint32_t w, pos, off;
        lit_from = i;
        do {
            i++;
            PROBE(pos, w);
            off = rep - pos;
            if (i - lit_from > 31) FLUSH(i);
        } while (i < size && !(w == le32(in + pos) && (off - 4u) < 0xFFFBu));
        if (i - lit_from > 0) FLUSH(i);
        if (!(i < size)) break;
        do {
            uint32_t k = 0;
            for (;;) { uint64_t x = le64(in + i + k) ^ le64(in + pos + k); uint32_t nb = x ? (uint32_t)(__builtin_ctzll(x) >> 3) : 8u; k += nb; if (nb != 8 || k >= 16) break; }
            if (k > 16) k = 16;
            uint32_t room = rep - pos;
            if (k > room) k = room - 1u;
            if (k < 4) break;
            if (!((room - 4u) < 0xFFFBu)) break;
            is_hit[i] = 1;
            i += k;
            SYM(i);
            PROBE(pos, w);
            off = rep - pos;
        } while (i < size - 5u && w == le32(in + pos) && (off - 4u) < 0xFFFBu);
    } while (i < size);
}

/* LRU over the last K distinct keys: returns 1 on hit; O(1) amortised with a timestamp table */
typedef struct { uint32_t* last; uint32_t clock; } lru_t;

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: probe_model <file> [block]\n"); return 1; }
    const uint32_t block = argc > 2 ? (uint32_t)atoi(argv[2]) : 262144u;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    fseek(f, 0, SEEK_END); long total = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = calloc((size_t)total + 256, 1);
    if (fread(buf, 1, (size_t)total, f) != (size_t)total) { perror("read"); return 1; }
    fclose(f);
    uint8_t* visited = malloc(block + 64), *is_hit = malloc(block + 64);
    uint32_t* cand_of = malloc(4 * (block + 64));
    static uint8_t written[HS];
    static uint32_t touched_at[HS];                 /* request counter value at the slot's last touch (0 = never) */
    const uint32_t Ws[4] = {8, 16, 32, 64}, Gs[4] = {1, 2, 4, 8}, Ks[4] = {256, 512, 2048, 8192};
    double bytes = 0, real = 0, hits = 0, age_lt2k = 0, age_lt64k = 0, inserts = 0;
    double wprobes[4] = {0}, bm_answered[4] = {0}, lru_hit[4] = {0}, unwritten = 0, wp32_real = 0;
    for (long at = 0; at + (long)block <= total; at += block) {
        const uint8_t* in = buf + at;
        memset(visited, 0, block + 64); memset(is_hit, 0, block + 64);
        parse(in, block, visited, is_hit, cand_of);
        bytes += block;
        for (uint32_t i = 1; i < block; i++) if (visited[i]) {
            real++; inserts++;
            if (is_hit[i]) { hits++; uint32_t age = i - cand_of[i]; if (age < 2048) age_lt2k++; if (age < 65536) age_lt64k++; }
        }
        /* window models */
        for (int wi = 0; wi < 4; wi++) {
            const uint32_t W = Ws[wi];
            uint32_t base = 1;
            memset(written, 0, sizeof written); memset(touched_at, 0, sizeof touched_at);
            uint32_t req = 0;
            while (base < block) {
                /* the window probes positions base .. base + W - 1; it ends after the last of them the parse reaches, the next
                 * window starts at the first visited position behind it */
                uint32_t end = base + W; if (end > block) end = block;
                for (uint32_t x = base; x < end; x++) {
                    const uint32_t h = hash17(le32(in + x));
                    wprobes[wi]++;
                    if (W == 32) {
                        if (visited[x]) wp32_real++;
                        if (!written[h]) unwritten++;
                        for (int gi = 0; gi < 4; gi++) {
                            const uint32_t G = Gs[gi], g0 = h & ~(G - 1u);
                            int any = 0;
                            for (uint32_t t = 0; t < G; t++) any |= written[g0 + t];
                            if (!any) bm_answered[gi]++;
                        }
                        int any4 = written[h & ~3u] | written[(h & ~3u) + 1] | written[(h & ~3u) + 2] | written[(h & ~3u) + 3];
                        if (any4) {                                   /* goes to memory: L2 model */
                            req++;
                            for (int ki = 0; ki < 4; ki++) if (touched_at[h] && req - touched_at[h] <= Ks[ki]) lru_hit[ki]++;
                            touched_at[h] = req;
                        }
                    }
                }
                /* commits of the window */
                for (uint32_t x = base; x < end; x++) if (visited[x]) {
                    const uint32_t h = hash17(le32(in + x));
                    written[h] = 1;
                    if (W == 32) { req++; touched_at[h] = req; }
                }
                uint32_t nx = end;
                while (nx < block && !visited[nx]) nx++;
                base = nx;
            }
        }
    }
    printf("blocks of %u bytes: %.0f, bytes %.0f\n", block, bytes / block, bytes);
    printf("real probes per byte            %.3f   (hits %.3f per byte = %.1f %% of the real probes)\n", real / bytes, hits / bytes, 100 * hits / real);
    printf("match candidates younger than 2 KiB %.1f %%, than 64 KiB %.1f %%\n", 100 * age_lt2k / hits, 100 * age_lt64k / hits);
    for (int wi = 0; wi < 4; wi++) printf("window of %2u positions: %.3f probes per byte\n", Ws[wi], wprobes[wi] / bytes);
    printf("window of 32: real %.1f %% of its probes; slot never written %.1f %%\n", 100 * wp32_real / wprobes[2], 100 * unwritten / wprobes[2]);
    for (int gi = 0; gi < 4; gi++) printf("  'ever written' bitmap, 1 bit per %u slots (%u KiB): answers %.1f %% of the window probes\n", Gs[gi], (HS / Gs[gi]) / 8192, 100 * bm_answered[gi] / wprobes[2]);
    const double to_mem = wprobes[2] - bm_answered[2];
    for (int ki = 0; ki < 4; ki++) printf("  of the probes that go to memory (1 bit per 4 slots), sector touched within the block's last %4u requests: %.1f %%\n", Ks[ki], 100 * lru_hit[ki] / to_mem);
    printf("requests per byte (W = 32, bitmap 1/4): probes %.3f + commits %.3f\n", to_mem / bytes, inserts / bytes);
    return 0;
}

/*
 * tsq_workload.c -- deterministic synthetic inputs for BASELINE.json's configs.
 *
 * Not part of the codec: bench.py and tests/ use it to build the buffers the
 * encode/decode path is measured on.  Every generator is
 * a pure function of (seed, absolute byte offset), produced in independent
 * 64 KiB chunks, so any sub-range can be regenerated on any thread or rank.
 *
 *   kind 0  "enwik9-shape" text: Zipf-distributed vocabulary words, wiki/XML
 *           markup, numbers and fresh rare tokens, calibrated so the reference's
 *           no-ext ratio at 4 MiB blocks is ~0.6225 (README.md:93 of the reference)
 *   kind 1  uniform random bytes (incompressible)
 *   kind 2  one random 8-byte pattern repeated (in[i] = pat[i & 7])
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define CHUNK 65536u
#define VOCAB 32768u
#define MAXW 14u

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    uint8_t  word[VOCAB][MAXW];
    uint8_t  wlen[VOCAB];
    uint32_t cdf[VOCAB];        /* cumulative Zipf weights scaled to 2^32 */
    uint64_t seed;
    int      ready;
} vocab_t;

static vocab_t g_vocab;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static const char LETTERS[] = "etaoinshrdlcumwfgypbvkjxqz";
/* cumulative English letter frequencies (per 1000) in LETTERS order */
static const uint16_t LCUM[26] = {127,218,300,375,445,512,575,636,696,739,779,807,835,859,883,905,925,945,964,979,989,997,998,999,1000,1000};

static void build_vocab(uint64_t seed, double zipf_s)
{
    uint64_t s = seed ^ 0xC0FFEEull;
    double total = 0, acc = 0;
    for (uint32_t w = 0; w < VOCAB; w++) {
        /* frequent words are short */
        uint32_t base = w < 64 ? 2 : (w < 1024 ? 3 : 4);
        uint32_t len = base + (uint32_t)(sm64(&s) % (w < 64 ? 3 : (w < 1024 ? 5 : 8)));
        if (len > MAXW) len = MAXW;
        g_vocab.wlen[w] = (uint8_t)len;
        for (uint32_t c = 0; c < len; c++) {
            uint32_t r = (uint32_t)(sm64(&s) % 1000u), k = 0;
            while (LCUM[k] <= r) k++;
            g_vocab.word[w][c] = (uint8_t)LETTERS[k];
        }
        if ((sm64(&s) & 15u) == 0) g_vocab.word[w][0] = (uint8_t)(g_vocab.word[w][0] - 32); /* Capitalised */
    }
    for (uint32_t w = 0; w < VOCAB; w++) total += 1.0 / pow((double)w + 2.7, zipf_s);
    for (uint32_t w = 0; w < VOCAB; w++) {
        acc += 1.0 / pow((double)w + 2.7, zipf_s);
        double v = acc / total * 4294967296.0;
        g_vocab.cdf[w] = v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
    }
    g_vocab.cdf[VOCAB - 1] = 0xFFFFFFFFu;
    g_vocab.seed = seed;
    g_vocab.ready = 1;
}

static const char *MARKUP[] = { "[[", "]]", "''", "'''", "==", "{{", "}}", "|", "&quot;", "&amp;", "&lt;", "&gt;",
                                "<ref>", "</ref>", "<page>", "</page>", "<title>", "</title>", "<id>", "</id>",
                                "<text xml:space=\"preserve\">", "</text>", "<revision>", "</revision>",
                                "<timestamp>", "</timestamp>", "http://www.", ".com/", "Category:", "* ", "# ", ":" };
#define NMARKUP (sizeof(MARKUP) / sizeof(MARKUP[0]))

/* knobs (set by tsqw_text_params, defaults calibrated in DESIGN.md) */
static double g_zipf = 1.355;
static uint32_t g_rare_per_1024 = 30;     /* share of tokens that are fresh random strings */
static uint32_t g_markup_per_1024 = 80;
static uint32_t g_number_per_1024 = 40;

void tsqw_text_params(double zipf_s, uint32_t rare, uint32_t markup, uint32_t number)
{
    pthread_mutex_lock(&g_lock);
    g_zipf = zipf_s; g_rare_per_1024 = rare; g_markup_per_1024 = markup; g_number_per_1024 = number;
    g_vocab.ready = 0;
    pthread_mutex_unlock(&g_lock);
}

static void text_chunk(uint64_t seed, uint64_t chunk, uint8_t *dst /* CHUNK bytes */)
{
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + chunk * 0xD1B54A32D192ED03ull + 1;
    uint32_t n = 0, col = 0;
    uint8_t tmp[64];
    while (n < CHUNK) {
        uint64_t r = sm64(&s);
        uint32_t kind = (uint32_t)(r & 1023u), len = 0;
        r >>= 10;
        if (kind < g_rare_per_1024) {                       /* fresh token: names, typos, foreign words */
            len = 3 + (uint32_t)(r % 9u); r >>= 4;
            uint64_t q = sm64(&s);
            for (uint32_t c = 0; c < len; c++) { tmp[c] = (uint8_t)('a' + (q % 26u)); q /= 26u; if (c == 10) q = sm64(&s); }
            if (r & 1u) tmp[0] = (uint8_t)(tmp[0] - 32);
        } else if (kind < g_rare_per_1024 + g_markup_per_1024) {
            const char *m = MARKUP[r % NMARKUP];
            len = (uint32_t)strlen(m); memcpy(tmp, m, len);
        } else if (kind < g_rare_per_1024 + g_markup_per_1024 + g_number_per_1024) {
            len = 1 + (uint32_t)(r % 6u); r >>= 3;
            uint64_t q = sm64(&s);
            for (uint32_t c = 0; c < len; c++) { tmp[c] = (uint8_t)('0' + (q % 10u)); q /= 10u; }
        } else {
            uint32_t u = (uint32_t)sm64(&s), lo = 0, hi = VOCAB - 1;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (g_vocab.cdf[mid] < u) lo = mid + 1; else hi = mid; }
            len = g_vocab.wlen[lo]; memcpy(tmp, g_vocab.word[lo], len);
        }
        /* separator: space, sometimes punctuation, newline at paragraph ends */
        uint32_t sep = (uint32_t)(sm64(&s) & 255u);
        if (sep < 16) tmp[len++] = ',';
        else if (sep < 26) tmp[len++] = '.';
        col += len + 1;
        if (col > 60 + (sep & 63u) * 6u) { tmp[len++] = '\n'; col = 0; if (sep & 64u) tmp[len++] = '\n'; }
        else tmp[len++] = ' ';
        if (len > CHUNK - n) len = CHUNK - n;
        memcpy(dst + n, tmp, len);
        n += len;
    }
}

static void random_chunk(uint64_t seed, uint64_t chunk, uint8_t *dst)
{
    uint64_t s = seed * 0xA24BAED4963EE407ull + chunk * 0x9FB21C651E98DF25ull + 7;
    for (uint32_t k = 0; k < CHUNK; k += 8) { uint64_t v = sm64(&s); memcpy(dst + k, &v, 8); }
}

typedef struct { int kind; uint64_t seed, offset, n; uint8_t *dst; uint64_t c0, c1; } job_t;

static void fill_range(const job_t *jb, uint64_t c0, uint64_t c1)
{
    uint8_t tmp[CHUNK];
    for (uint64_t c = c0; c < c1; c++) {
        uint64_t lo = c * CHUNK, hi = lo + CHUNK;
        uint64_t a = lo < jb->offset ? jb->offset : lo;
        uint64_t b = hi > jb->offset + jb->n ? jb->offset + jb->n : hi;
        if (a >= b) continue;
        if (jb->kind == 0) text_chunk(jb->seed, c, tmp); else random_chunk(jb->seed, c, tmp);
        memcpy(jb->dst + (a - jb->offset), tmp + (a - lo), b - a);
    }
}

static void *worker(void *p) { job_t *jb = (job_t *)p; fill_range(jb, jb->c0, jb->c1); return 0; }

/*
 * Fill dst[0..n) with bytes [offset, offset+n) of the infinite stream (kind, seed).
 */
void tsqw_fill(int kind, uint64_t seed, uint64_t offset, uint64_t n, uint8_t *dst, int threads)
{
    if (n == 0) return;
    if (kind == 2) {
        uint64_t s = seed ^ 0x5EEDull; uint64_t pat = sm64(&s); uint8_t p[8]; memcpy(p, &pat, 8);
        for (uint64_t k = 0; k < n; k++) dst[k] = p[(offset + k) & 7u];
        return;
    }
    if (kind == 0) {
        pthread_mutex_lock(&g_lock);
        if (!g_vocab.ready || g_vocab.seed != seed) build_vocab(seed, g_zipf);
        pthread_mutex_unlock(&g_lock);
    }
    uint64_t c0 = offset / CHUNK, c1 = (offset + n + CHUNK - 1) / CHUNK;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > c1 - c0) threads = (int)(c1 - c0);
    if (threads == 1) { job_t jb = { kind, seed, offset, n, dst, c0, c1 }; fill_range(&jb, c0, c1); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
    uint64_t per = (c1 - c0 + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        uint64_t a = c0 + per * t, b = a + per; if (a > c1) a = c1; if (b > c1) b = c1;
        jobs[t] = (job_t){ kind, seed, offset, n, dst, a, b };
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    free(th); free(jobs);
}

/*
 * f2q_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the read -> feature -> count
 * path of 2FAST2Q v2.8.1 (reference: fast2q/fast2q.py).  It is the checker for the CUDA path and the
 * timed "port" baseline of bench.py; it is never linked, imported or called by the product
 * (2fast2q_b200/ and libf2q.so).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the unmodified reference from
 * /root/reference, runs reads_counter / border_finder / sequence_tinder on the committed inputs and
 * stores the reference's own outputs in tests/golden/*.json; tests/test_oracle_golden.py checks this file
 * against every one of them, plus the reference's own unit vectors (tests/test_mainfunctions.py:4-78).
 *
 * Each function cites the reference lines it follows.  The code is written from the behaviour of
 * those lines, not copied from them.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/f2q.h"

#if defined(__GNUC__)
#define F2QO_API __attribute__((visibility("default")))
#else
#define F2QO_API
#endif

/* ------------------------------------------------------------------------------------------
 * small byte-string hash map (stand-in for the Python dicts / sets on the path)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t *off;    /* key offset in arena per slot */
    uint32_t *len;    /* key length per slot */
    int64_t  *val;    /* value per slot; slot empty when used[slot]==0 */
    uint8_t  *used;
    uint64_t cap, n;
    uint8_t  *arena;
    uint64_t arena_len, arena_cap;
    /* insertion order (Python dicts preserve it) */
    uint64_t *order;  /* slot index per insertion */
} bmap;

static uint64_t fnv1a(const uint8_t *p, uint32_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (uint32_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return h;
}

static int bmap_init(bmap *m, uint64_t cap_pow2) {
    memset(m, 0, sizeof(*m));
    m->cap = cap_pow2;
    m->off = (uint64_t *)calloc(m->cap, sizeof(uint64_t));
    m->len = (uint32_t *)calloc(m->cap, sizeof(uint32_t));
    m->val = (int64_t *)calloc(m->cap, sizeof(int64_t));
    m->used = (uint8_t *)calloc(m->cap, 1);
    m->order = (uint64_t *)calloc(m->cap, sizeof(uint64_t));
    m->arena_cap = 1 << 16;
    m->arena = (uint8_t *)malloc(m->arena_cap);
    return (m->off && m->len && m->val && m->used && m->order && m->arena) ? 0 : -1;
}

static void bmap_free(bmap *m) {
    free(m->off); free(m->len); free(m->val); free(m->used); free(m->order); free(m->arena);
    memset(m, 0, sizeof(*m));
}

/* returns slot index of key, or the empty slot where it would go (check used[]) */
static uint64_t bmap_probe(const bmap *m, const uint8_t *k, uint32_t n) {
    uint64_t i = fnv1a(k, n) & (m->cap - 1);
    while (m->used[i]) {
        if (m->len[i] == n && memcmp(m->arena + m->off[i], k, n) == 0) return i;
        i = (i + 1) & (m->cap - 1);
    }
    return i;
}

static int bmap_grow(bmap *m) {
    bmap g;
    if (bmap_init(&g, m->cap * 2)) return -1;
    free(g.arena);
    g.arena = m->arena; g.arena_len = m->arena_len; g.arena_cap = m->arena_cap;
    for (uint64_t j = 0; j < m->n; j++) {
        uint64_t s = m->order[j];
        uint64_t i = bmap_probe(&g, g.arena + m->off[s], m->len[s]);
        g.used[i] = 1; g.off[i] = m->off[s]; g.len[i] = m->len[s]; g.val[i] = m->val[s];
        g.order[g.n++] = i;
    }
    free(m->off); free(m->len); free(m->val); free(m->used); free(m->order);
    *m = g;
    return 0;
}

/* find-or-insert; *inserted tells which.  returns slot or (uint64_t)-1 on OOM */
static uint64_t bmap_put(bmap *m, const uint8_t *k, uint32_t n, int64_t v_if_new, int *inserted) {
    if ((m->n + 1) * 2 > m->cap) { if (bmap_grow(m)) return (uint64_t)-1; }
    uint64_t i = bmap_probe(m, k, n);
    if (m->used[i]) { *inserted = 0; return i; }
    if (m->arena_len + n + 1 > m->arena_cap) {
        uint64_t nc = m->arena_cap * 2;
        while (nc < m->arena_len + n + 1) nc *= 2;
        uint8_t *na = (uint8_t *)realloc(m->arena, nc);
        if (!na) return (uint64_t)-1;
        m->arena = na; m->arena_cap = nc;
    }
    memcpy(m->arena + m->arena_len, k, n);
    m->used[i] = 1; m->off[i] = m->arena_len; m->len[i] = n; m->val[i] = v_if_new;
    m->arena_len += n;
    m->order[m->n++] = i;
    *inserted = 1;
    return i;
}

static int64_t bmap_get(const bmap *m, const uint8_t *k, uint32_t n, int *found) {
    uint64_t i = bmap_probe(m, k, n);
    *found = m->used[i];
    return m->used[i] ? m->val[i] : 0;
}

/* ------------------------------------------------------------------------------------------
 * Python semantics helpers
 * ---------------------------------------------------------------------------------------- */

/* bytes.rstrip(): strips b" \t\n\r\x0b\x0c" from the right (fast2q.py:326) */
static int64_t rstrip_len(const uint8_t *p, int64_t n) {
    while (n > 0) {
        uint8_t c = p[n - 1];
        if (c == ' ' || (c >= 9 && c <= 13)) n--; else break;
    }
    return n;
}

/* Python slice bounds seq[a:b] on a sequence of length n, for int a,b (fast2q.py:354-355, 252-253) */
static void py_slice(int64_t n, int64_t a, int64_t b, int64_t *lo, int64_t *hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    if (b < a) b = a;
    *lo = a; *hi = b;
}

/* bytes.upper(): ASCII a-z only (fast2q.py:354) */
static uint8_t up8(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* fail set of fast2q.py:1112-1129: quality_list = chr(33)..chr(126); set(quality_list[:ph-1]) with ph<=0 -> 1.
 * A set of bytes is a 256-bit mask (bit b of word b>>6). */
typedef struct { uint64_t w[4]; } byteset;

static byteset fail_set(int ph) {
    byteset s = { {0, 0, 0, 0} };
    if (ph <= 0) ph = 1;
    int n = ph - 1;            /* number of characters taken from the front of the 94-long list */
    if (n > 94) n = 94;
    for (int b = 33; b < 33 + n; b++) s.w[b >> 6] |= 1ULL << (b & 63);
    return s;
}

static int slice_fails(const uint8_t *q, int64_t lo, int64_t hi, const byteset *s) {
    for (int64_t i = lo; i < hi; i++) if ((s->w[q[i] >> 6] >> (q[i] & 63)) & 1) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * binary_subtract (fast2q.py:601-626): 1 if <= mismatch differing positions over zip(a,b)
 * ---------------------------------------------------------------------------------------- */
static int binary_subtract(const uint8_t *a, int64_t na, const uint8_t *b, int64_t nb, int mismatch) {
    int64_t n = na < nb ? na : nb;
    int miss = 0;
    for (int64_t i = 0; i < n; i++) {
        if (a[i] != b[i]) miss++;
        if (miss > mismatch) return 0;
    }
    return 1;
}

/* border_finder (fast2q.py:628-658): first i+start_place, or -1 for None */
static int64_t border_finder(const uint8_t *seq, int64_t s, const uint8_t *read, int64_t r, int mismatch,
                             int64_t start_place) {
    int64_t fall_over = r - s;
    int64_t lo, hi;
    py_slice(r, start_place, r, &lo, &hi);          /* read[start_place:] */
    int64_t iters = hi - lo;
    for (int64_t i = 0; i < iters; i++) {
        int64_t a, b;
        py_slice(r, start_place + i, s + start_place + i, &a, &b);
        int finder = binary_subtract(seq, s, read + a, b - a, mismatch);
        if (i + start_place > fall_over) return -1;
        if (finder) return i + start_place;
    }
    return -1;
}

F2QO_API int f2qo_border_finder(const uint8_t *seq, int32_t s, const uint8_t *read, int32_t r, int32_t mismatch,
                                int32_t start_place) {
    return (int)border_finder(seq, s, read, r, mismatch, start_place);
}

/* sequence_tinder (fast2q.py:215-285).  returns 1 and (start,end) or 0 for (None,None).
 * read = raw (not upper-cased) sequence line, as passed at fast2q.py:337. */
static int sequence_tinder(const f2q_config *c, int i, const uint8_t *read, int64_t r, const uint8_t *qual,
                           int64_t q, const byteset *set_up, const byteset *set_down, int64_t *start_out,
                           int64_t *end_out) {
    const byteset *fu = set_up, *fd = set_down;
    int64_t lo, hi;
    if (c->has_up && c->has_down) {
        int64_t ul = c->up_len[i], dl = c->down_len[i];
        int64_t start = border_finder(c->up[i], ul, read, r, c->miss_up, 0);
        if (start >= 0) {
            int64_t end = border_finder(c->down[i], dl, read, r, c->miss_down, start + ul);
            if (end >= 0) {
                int bad = 0;
                py_slice(q, start, start + ul, &lo, &hi); bad |= slice_fails(qual, lo, hi, fu);
                py_slice(q, end, end + dl, &lo, &hi);     bad |= slice_fails(qual, lo, hi, fd);
                if (!bad) { *start_out = start + ul; *end_out = end; return 1; }
            }
        }
    } else if (c->has_up) {
        int64_t ul = c->up_len[i];
        int64_t start = border_finder(c->up[i], ul, read, r, c->miss_up, 0);
        if (start >= 0) {
            py_slice(q, start, start + ul, &lo, &hi);
            if (!slice_fails(qual, lo, hi, fu)) {
                *start_out = start + ul; *end_out = start + ul + c->length; return 1;
            }
        }
    } else if (c->has_down) {
        int64_t dl = c->down_len[i];
        int64_t end = border_finder(c->down[i], dl, read, r, c->miss_down, 0);
        if (end >= 0) {
            py_slice(q, end, end + dl, &lo, &hi);
            if (!slice_fails(qual, lo, hi, fd)) { *start_out = end - c->length; *end_out = end; return 1; }
        }
    }
    return 0;
}

/* set_up/set_down: explicit 256-bit fail sets (uint64[4]) or NULL for the --qsu/--qsd derived ones */
F2QO_API int f2qo_sequence_tinder(const f2q_config *c, int32_t i, const uint8_t *read, int32_t r,
                                  const uint8_t *qual, int32_t q, const uint64_t *set_up, const uint64_t *set_down,
                                  int32_t *start, int32_t *end) {
    int64_t s = 0, e = 0;
    byteset su = fail_set(c->qual_up), sd = fail_set(c->qual_down);
    if (set_up) memcpy(su.w, set_up, 32);
    if (set_down) memcpy(sd.w, set_down, 32);
    int ok = sequence_tinder(c, i, read, r, qual, q, &su, &sd, &s, &e);
    *start = (int32_t)s; *end = (int32_t)e;
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * per-read key construction (fast2q.py:332-363).  returns 1 and the key, or 0 when every iteration
 * was flagged (=> quality_failed).  key buffer must hold n_iter*(line length + 1) bytes.
 * ---------------------------------------------------------------------------------------- */
static int build_key(const f2q_config *c, const uint8_t *R, int64_t r, const uint8_t *Q, int64_t q, uint8_t *key,
                     int64_t *key_len) {
    int fixed = !(c->has_up || c->has_down);
    byteset fqs = fail_set(c->phred), su = fail_set(c->qual_up), sd = fail_set(c->qual_down);
    const byteset *fq = &fqs;
    int any = 0;
    int64_t kl = 0;
    for (int i = 0; i < c->n_iter; i++) {
        int64_t start = 0, end = 0;
        int have = 1;
        if (!fixed) {
            have = sequence_tinder(c, i, R, r, Q, q, &su, &sd, &start, &end);   /* :337-340 */
            if (have && end < start) have = 0;                      /* :343-345 */
        } else {
            start = c->starts[i];                                    /* :349-351 */
            end = (int64_t)c->starts[i] + c->length;                 /* reads_counter :540 */
        }
        if (!have) continue;                                         /* flagged */
        int64_t slo, shi, qlo, qhi;
        py_slice(r, start, end, &slo, &shi);                         /* :354 */
        py_slice(q, start, end, &qlo, &qhi);                         /* :355 */
        if (slice_fails(Q, qlo, qhi, fq)) continue;                  /* :357-360, flagged */
        if (any) key[kl++] = ':';                                    /* ":" join, :358,363 */
        for (int64_t k = slo; k < shi; k++) key[kl++] = up8(R[k]);
        any = 1;
    }
    *key_len = kl;
    return any;
}

/* ------------------------------------------------------------------------------------------
 * record iteration (fast2q.py:324-328,392): lines split on '\n', every 4 lines = one read
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *data; uint64_t n, pos;
} line_iter;

/* next line (without its '\n'); returns 0 at end.  A final unterminated non-empty line counts. */
static int next_line(line_iter *it, const uint8_t **p, int64_t *len) {
    if (it->pos >= it->n) return 0;
    const uint8_t *s = it->data + it->pos;
    const uint8_t *e = (const uint8_t *)memchr(s, '\n', it->n - it->pos);
    if (e) { *p = s; *len = e - s; it->pos += (uint64_t)(e - s) + 1; }
    else   { *p = s; *len = (int64_t)(it->n - it->pos); it->pos = it->n; }
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Counter mode: fastq_parser + mismatch_search_handler + features_all_vs_all
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *bytes; const uint64_t *off; uint32_t n;
} library;

/* features_all_vs_all (fast2q.py:660-690): index of the only same-length entry within `mismatch`, else -1 */
static int64_t features_all_vs_all(const library *lib, const uint8_t *read, int64_t r, int mismatch) {
    int found = 0; int64_t found_guide = -1;
    for (uint32_t g = 0; g < lib->n; g++) {
        int64_t gl = (int64_t)(lib->off[g + 1] - lib->off[g]);
        if (gl == r) {
            if (binary_subtract(lib->bytes + lib->off[g], gl, read, r, mismatch)) {
                found++; found_guide = g;
                if (found >= 2) return -1;
            }
        }
    }
    return found == 1 ? found_guide : -1;
}

F2QO_API int f2qo_count(const f2q_config *c, const uint8_t *lib_bytes, const uint64_t *lib_off, uint32_t n_keys,
                        const uint8_t *data, uint64_t nbytes, uint64_t *counts, uint64_t *stats) {
    library lib = { lib_bytes, lib_off, n_keys };
    bmap dict, memo;            /* features dict; passed_reads (val>=0) + failed_reads (val=-1) memo (:724-731) */
    uint64_t cap = 16; while (cap < (uint64_t)n_keys * 2 + 2) cap *= 2;
    if (bmap_init(&dict, cap)) return F2Q_ENOMEM;
    if (bmap_init(&memo, 1 << 12)) { bmap_free(&dict); return F2Q_ENOMEM; }
    int ins;
    for (uint32_t g = 0; g < n_keys; g++) {
        /* first sequence wins (features_loader :160-165); callers pass unique keys anyway */
        bmap_put(&dict, lib_bytes + lib_off[g], (uint32_t)(lib_off[g + 1] - lib_off[g]), g, &ins);
    }
    memset(counts, 0, sizeof(uint64_t) * n_keys);
    memset(stats, 0, sizeof(uint64_t) * F2Q_N_STATS);

    line_iter it = { data, nbytes, 0 };
    const uint8_t *lp[4]; int64_t ll[4];
    int have = 0;
    uint8_t *key = NULL; int64_t key_cap = 0;
    const uint8_t *p; int64_t len;
    int rc = F2Q_OK;
    while (next_line(&it, &p, &len)) {
        lp[have] = p; ll[have] = rstrip_len(p, len); have++;                 /* :326 */
        if (have < 4) continue;                                              /* :328 */
        have = 0;
        int64_t need = (ll[1] + 2) * (int64_t)(c->n_iter > 0 ? c->n_iter : 1) + 8;
        if (need > key_cap) { free(key); key_cap = need * 2; key = (uint8_t *)malloc((size_t)key_cap); if (!key) { rc = F2Q_ENOMEM; break; } }
        int64_t kl;
        if (build_key(c, lp[1], ll[1], lp[3], ll[3], key, &kl)) {
            int found;
            int64_t g = bmap_get(&dict, key, (uint32_t)kl, &found);
            if (found) { counts[g]++; stats[F2Q_STAT_PERFECT]++; }           /* :365-367 */
            else if (c->miss > 0) {                                          /* :369-378 */
                int64_t m = bmap_get(&memo, key, (uint32_t)kl, &found);
                if (!found) {
                    m = -1;
                    for (int k = 1; k <= c->miss; k++) {                      /* :734-750 */
                        int64_t f = features_all_vs_all(&lib, key, kl, k);
                        if (f >= 0) { m = f; break; }
                    }
                    if (bmap_put(&memo, key, (uint32_t)kl, m, &ins) == (uint64_t)-1) { rc = F2Q_ENOMEM; break; }
                }
                if (m >= 0) { counts[m]++; stats[F2Q_STAT_IMPERFECT]++; }
                else stats[F2Q_STAT_NON_ALIGNED]++;
            } else stats[F2Q_STAT_NON_ALIGNED]++;                            /* :379-380 */
        } else stats[F2Q_STAT_QUALITY_FAILED]++;                             /* :389-390 */
        stats[F2Q_STAT_READS]++;                                             /* :393 */
    }
    free(key);
    bmap_free(&dict); bmap_free(&memo);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Extract + Count mode (fast2q.py:382-387): dict of novel keys, insertion-ordered
 * ---------------------------------------------------------------------------------------- */
typedef struct f2qo_ec { bmap m; } f2qo_ec;

F2QO_API f2qo_ec *f2qo_ec_run(const f2q_config *c, const uint8_t *data, uint64_t nbytes, uint64_t *stats) {
    f2qo_ec *h = (f2qo_ec *)calloc(1, sizeof(f2qo_ec));
    if (!h) return NULL;
    if (bmap_init(&h->m, 1 << 12)) { free(h); return NULL; }
    memset(stats, 0, sizeof(uint64_t) * F2Q_N_STATS);
    line_iter it = { data, nbytes, 0 };
    const uint8_t *lp[4]; int64_t ll[4];
    int have = 0, ins;
    uint8_t *key = NULL; int64_t key_cap = 0;
    const uint8_t *p; int64_t len;
    while (next_line(&it, &p, &len)) {
        lp[have] = p; ll[have] = rstrip_len(p, len); have++;
        if (have < 4) continue;
        have = 0;
        int64_t need = (ll[1] + 2) * (int64_t)(c->n_iter > 0 ? c->n_iter : 1) + 8;
        if (need > key_cap) { free(key); key_cap = need * 2; key = (uint8_t *)malloc((size_t)key_cap); if (!key) break; }
        int64_t kl;
        if (build_key(c, lp[1], ll[1], lp[3], ll[3], key, &kl)) {
            uint64_t s = bmap_put(&h->m, key, (uint32_t)kl, 0, &ins);
            if (s == (uint64_t)-1) break;
            h->m.val[s]++;                                                   /* :383-386 */
            stats[F2Q_STAT_PERFECT]++;                                       /* :387 */
        } else stats[F2Q_STAT_QUALITY_FAILED]++;
        stats[F2Q_STAT_READS]++;
    }
    free(key);
    return h;
}

F2QO_API void f2qo_ec_size(const f2qo_ec *h, uint64_t *n_keys, uint64_t *key_bytes) {
    *n_keys = h->m.n; *key_bytes = h->m.arena_len;
}

/* keys in insertion order (Python dict order) */
F2QO_API void f2qo_ec_get(const f2qo_ec *h, uint8_t *key_bytes, uint64_t *key_off, uint64_t *counts) {
    uint64_t o = 0;
    for (uint64_t j = 0; j < h->m.n; j++) {
        uint64_t s = h->m.order[j];
        key_off[j] = o;
        memcpy(key_bytes + o, h->m.arena + h->m.off[s], h->m.len[s]);
        o += h->m.len[s];
        counts[j] = (uint64_t)h->m.val[s];
    }
    key_off[h->m.n] = o;
}

F2QO_API void f2qo_ec_free(f2qo_ec *h) { if (h) { bmap_free(&h->m); free(h); } }

/* per-read key, for differential tests of the extraction rules alone.  returns 1/0 like build_key */
F2QO_API int f2qo_build_key(const f2q_config *c, const uint8_t *R, int32_t r, const uint8_t *Q, int32_t q,
                            uint8_t *key, int32_t *key_len) {
    int64_t kl = 0;
    int ok = build_key(c, R, r, Q, q, key, &kl);
    *key_len = (int32_t)kl;
    return ok;
}

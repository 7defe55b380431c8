/*
 * bm25_oracle.c — CPU restatement of NextSearch's query-time BM25 + top-k path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nextsearch-api_b200/ links, loads or calls this file;
 * it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg have an
 * independent statement of the reference algorithm to check the CUDA path against.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this oracle
 * is pinned against the reference ITSELF: oracle/_ref/ref_engine is the reference's unmodified
 * src/api_engine.cpp + src/api_segment.cpp (+ autocomplete/metadata/semantic objects) compiled
 * by oracle/Makefile, and tests/golden/ holds its outputs (tests/golden/make_golden.py is the
 * generating script).  tests/test_oracle_golden.py checks this file against those vectors, and
 * tests/test_oracle_golden.py (-m ref cases) re-runs the live reference binary when it is present.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 *
 * Differences from the reference, all deliberate:
 *   - postings are read into memory once instead of streamed from std::ifstream per query
 *     (same bytes, same order);
 *   - the per-segment std::unordered_map<uint32_t,float> (src/api_engine.cpp:445) is a dense
 *     float array + touched list: per-doc additions happen in the same (query-term) order, so
 *     every score is the same f32;
 *   - the top-K heap's tie behaviour (an artefact of hash-map iteration order, :485-504) is
 *     replaced by the stated total order: score desc, segment index asc, docId asc.  The strict
 *     '>' replacement rule at :488 means an earlier segment keeps its slot on equal score, which
 *     this order preserves.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -march: the reference's CMake sets
 * neither optimisation nor FMA flags, so every float op rounds once).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <dirent.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#define ORC_BARRELS_MAX 65536

typedef struct {
    char* term;        /* owned */
    uint32_t termId, df, count, barrel;
    uint64_t begin;    /* posting index into seg->post */
} orc_lex;

typedef struct {
    uint32_t N;        /* stats.bin */
    float avgdl;
    uint32_t ndocs;    /* docs.bin */
    uint32_t* doc_len;
    char** uid;
    uint64_t npost;
    uint32_t* post;    /* interleaved docId, tf */
    uint32_t nlex;
    orc_lex* lex;
    uint32_t hcap;     /* open-addressing table over lex[] (power of two) */
    int32_t* htab;
} orc_seg;

typedef struct orc_index {
    int nseg;
    char** names;
    orc_seg* segs;
    uint32_t max_docs;
    char err[512];
} orc_index;

static char g_err[512];
const char* orc_last_error(void) { return g_err; }

/* ---- include/indexio.hpp:13-29 : raw little-endian reads ---- */
typedef struct { const uint8_t* p; const uint8_t* end; int ok; } rd_t;
static uint32_t rd_u32(rd_t* r) { uint32_t v = 0; if (r->end - r->p < 4) { r->ok = 0; return 0; } memcpy(&v, r->p, 4); r->p += 4; return v; }
static uint64_t rd_u64(rd_t* r) { uint64_t v = 0; if (r->end - r->p < 8) { r->ok = 0; return 0; } memcpy(&v, r->p, 8); r->p += 8; return v; }
static float rd_f32(rd_t* r) { float v = 0; if (r->end - r->p < 4) { r->ok = 0; return 0; } memcpy(&v, r->p, 4); r->p += 4; return v; }
static char* rd_str(rd_t* r) {
    uint32_t n = rd_u32(r);
    if (!r->ok || (uint64_t)(r->end - r->p) < n) { r->ok = 0; return NULL; }
    char* s = (char*)malloc((size_t)n + 1);
    memcpy(s, r->p, n);
    s[n] = 0;
    r->p += n;
    return s;
}

static uint8_t* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* b = (uint8_t*)malloc(sz > 0 ? (size_t)sz : 1);
    size_t got = sz > 0 ? fread(b, 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) { free(b); return NULL; }
    *n = (size_t)sz;
    return b;
}

static int exists(const char* p) { struct stat st; return stat(p, &st) == 0; }

static uint32_t fnv1a(const char* s) {
    uint32_t h = 2166136261u;
    for (; *s; s++) { h ^= (uint8_t)*s; h *= 16777619u; }
    return h;
}

/* term -> lexicon entry; the reference's unordered_map::emplace keeps the FIRST entry of a term
 * (src/api_segment.cpp:60,99) */
static void lex_build(orc_seg* s) {
    uint32_t cap = 16;
    while (cap < s->nlex * 2u + 1u) cap <<= 1;
    s->hcap = cap;
    s->htab = (int32_t*)malloc(sizeof(int32_t) * cap);
    for (uint32_t i = 0; i < cap; i++) s->htab[i] = -1;
    for (uint32_t i = 0; i < s->nlex; i++) {
        uint32_t h = fnv1a(s->lex[i].term) & (cap - 1);
        int dup = 0;
        while (s->htab[h] >= 0) {
            if (strcmp(s->lex[s->htab[h]].term, s->lex[i].term) == 0) { dup = 1; break; }
            h = (h + 1) & (cap - 1);
        }
        if (!dup) s->htab[h] = (int32_t)i;
    }
}

static const orc_lex* lex_find(const orc_seg* s, const char* term) {
    if (!s->hcap) return NULL;
    uint32_t h = fnv1a(term) & (s->hcap - 1);
    while (s->htab[h] >= 0) {
        const orc_lex* e = &s->lex[s->htab[h]];
        if (strcmp(e->term, term) == 0) return e;
        h = (h + 1) & (s->hcap - 1);
    }
    return NULL;
}

/* one lexicon file: u32 count; count x {str term, u32 termId, u32 df, u64 offset, u32 count}
 * (src/api_segment.cpp:50-61, :88-99) */
static int load_lexicon(orc_seg* s, const char* path, uint32_t barrel, uint64_t base, uint64_t inv_bytes) {
    size_t n;
    uint8_t* b = slurp(path, &n);
    if (!b) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return 0; }
    rd_t r = {b, b + n, 1};
    uint32_t cnt = rd_u32(&r);
    s->lex = (orc_lex*)realloc(s->lex, sizeof(orc_lex) * ((size_t)s->nlex + cnt + 1));
    for (uint32_t i = 0; i < cnt && r.ok; i++) {
        orc_lex e;
        e.term = rd_str(&r);
        e.termId = rd_u32(&r);
        e.df = rd_u32(&r);
        uint64_t off = rd_u64(&r);
        e.count = rd_u32(&r);
        e.barrel = barrel;
        if (!r.ok) { free(e.term); break; }
        if (off % 8 || off + (uint64_t)e.count * 8 > inv_bytes) {
            snprintf(g_err, sizeof g_err, "lexicon entry outside inverted file in %s", path);
            free(e.term);
            free(b);
            return 0;
        }
        e.begin = base + off / 8;
        s->lex[s->nlex++] = e;
    }
    int ok = r.ok;
    free(b);
    if (!ok) snprintf(g_err, sizeof g_err, "truncated %s", path);
    return ok;
}

/* src/api_segment.cpp:105-136 (+ :45-102), include/barrels.hpp:33-71 */
static int load_segment(const char* segdir, orc_seg* s) {
    char path[4096];
    size_t n;
    memset(s, 0, sizeof *s);
    snprintf(path, sizeof path, "%s/stats.bin", segdir);
    uint8_t* b = slurp(path, &n);
    if (!b) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return 0; }
    { rd_t r = {b, b + n, 1}; s->N = rd_u32(&r); s->avgdl = rd_f32(&r); free(b); if (!r.ok) return 0; }

    snprintf(path, sizeof path, "%s/docs.bin", segdir);
    b = slurp(path, &n);
    if (!b) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return 0; }
    {
        rd_t r = {b, b + n, 1};
        s->ndocs = rd_u32(&r);
        s->doc_len = (uint32_t*)calloc((size_t)s->ndocs + 1, 4);
        s->uid = (char**)calloc((size_t)s->ndocs + 1, sizeof(char*));
        for (uint32_t i = 0; i < s->ndocs && r.ok; i++) {
            s->uid[i] = rd_str(&r);
            free(rd_str(&r)); /* title skipped */
            free(rd_str(&r)); /* json_relpath skipped */
            s->doc_len[i] = rd_u32(&r);
        }
        free(b);
        if (!r.ok) { snprintf(g_err, sizeof g_err, "truncated %s", path); return 0; }
    }

    char p0[4096], p1[4096], p2[4096];
    snprintf(p0, sizeof p0, "%s/barrels.bin", segdir);
    snprintf(p1, sizeof p1, "%s/inverted_b000.bin", segdir);
    snprintf(p2, sizeof p2, "%s/lexicon_b000.bin", segdir);
    if (exists(p0) && exists(p1) && exists(p2)) {
        b = slurp(p0, &n);
        if (!b) return 0;
        rd_t r = {b, b + n, 1};
        uint32_t B = rd_u32(&r);
        (void)rd_u32(&r); /* terms_per_barrel: not needed to read */
        free(b);
        if (!r.ok || B == 0 || B > ORC_BARRELS_MAX) { snprintf(g_err, sizeof g_err, "bad barrels.bin in %s", segdir); return 0; }
        uint64_t* base = (uint64_t*)calloc((size_t)B + 1, 8);
        uint64_t* bytes = (uint64_t*)calloc((size_t)B, 8);
        for (uint32_t i = 0; i < B; i++) {
            snprintf(path, sizeof path, "%s/inverted_b%03u.bin", segdir, i);
            struct stat st;
            if (stat(path, &st) != 0) { snprintf(g_err, sizeof g_err, "cannot open %s", path); free(base); free(bytes); return 0; }
            bytes[i] = (uint64_t)st.st_size;
            base[i + 1] = base[i] + bytes[i] / 8;
        }
        s->npost = base[B];
        s->post = (uint32_t*)malloc(s->npost * 8 + 8);
        for (uint32_t i = 0; i < B; i++) {
            snprintf(path, sizeof path, "%s/inverted_b%03u.bin", segdir, i);
            FILE* f = fopen(path, "rb");
            size_t got = f ? fread(s->post + base[i] * 2, 1, bytes[i], f) : 0;
            if (f) fclose(f);
            if (!f || got != bytes[i]) { snprintf(g_err, sizeof g_err, "short read %s", path); free(base); free(bytes); return 0; }
        }
        for (uint32_t i = 0; i < B; i++) {
            snprintf(path, sizeof path, "%s/lexicon_b%03u.bin", segdir, i);
            if (!load_lexicon(s, path, i, base[i], bytes[i])) { free(base); free(bytes); return 0; }
        }
        free(base);
        free(bytes);
    } else {
        snprintf(path, sizeof path, "%s/inverted.bin", segdir);
        size_t nb;
        uint8_t* inv = slurp(path, &nb);
        if (!inv) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return 0; }
        s->npost = nb / 8;
        s->post = (uint32_t*)inv;
        snprintf(path, sizeof path, "%s/lexicon.bin", segdir);
        if (!load_lexicon(s, path, 0, 0, nb)) return 0;
    }
    lex_build(s);
    return 1;
}

static int cmp_str(const void* a, const void* b) { return strcmp(*(char* const*)a, *(char* const*)b); }

/* Engine::reload's segment discovery: manifest.bin (src/api_segment.cpp:14-26) else sorted
 * scan of segments/seg_* (src/api_engine.cpp:57-70) */
orc_index* orc_open(const char* index_dir) {
    orc_index* ix = (orc_index*)calloc(1, sizeof *ix);
    char path[4096];
    snprintf(path, sizeof path, "%s/manifest.bin", index_dir);
    size_t n;
    uint8_t* b = slurp(path, &n);
    if (b) {
        rd_t r = {b, b + n, 1};
        uint32_t cnt = rd_u32(&r);
        ix->names = (char**)calloc((size_t)cnt + 1, sizeof(char*));
        for (uint32_t i = 0; i < cnt && r.ok; i++) {
            char* s = rd_str(&r);
            if (s) ix->names[ix->nseg++] = s;
        }
        free(b);
    }
    if (ix->nseg == 0) {
        snprintf(path, sizeof path, "%s/segments", index_dir);
        DIR* d = opendir(path);
        if (d) {
            struct dirent* e;
            int cap = 16;
            free(ix->names);
            ix->names = (char**)calloc((size_t)cap, sizeof(char*));
            while ((e = readdir(d))) {
                if (strncmp(e->d_name, "seg_", 4) != 0) continue;
                {   /* only directories (src/api_engine.cpp:64, e.is_directory() follows symlinks: stat, not lstat) */
                    char sub[4400];
                    struct stat sb;
                    snprintf(sub, sizeof sub, "%s/%s", path, e->d_name);
                    if (stat(sub, &sb) != 0 || !S_ISDIR(sb.st_mode)) continue;
                }
                if (ix->nseg == cap) { cap *= 2; ix->names = (char**)realloc(ix->names, sizeof(char*) * (size_t)cap); }
                ix->names[ix->nseg++] = strdup(e->d_name);
            }
            closedir(d);
            qsort(ix->names, (size_t)ix->nseg, sizeof(char*), cmp_str);
        }
    }
    if (ix->nseg == 0) { snprintf(g_err, sizeof g_err, "no segments under %s", index_dir); free(ix->names); free(ix); return NULL; }
    ix->segs = (orc_seg*)calloc((size_t)ix->nseg, sizeof(orc_seg));
    for (int i = 0; i < ix->nseg; i++) {
        snprintf(path, sizeof path, "%s/segments/%s", index_dir, ix->names[i]);
        if (!load_segment(path, &ix->segs[i])) { free(ix); return NULL; } /* leak on error: test code */
        if (ix->segs[i].ndocs > ix->max_docs) ix->max_docs = ix->segs[i].ndocs;
    }
    return ix;
}

void orc_close(orc_index* ix) {
    if (!ix) return;
    for (int i = 0; i < ix->nseg; i++) {
        orc_seg* s = &ix->segs[i];
        for (uint32_t d = 0; d < s->ndocs; d++) free(s->uid[d]);
        for (uint32_t t = 0; t < s->nlex; t++) free(s->lex[t].term);
        free(s->uid); free(s->doc_len); free(s->post); free(s->lex); free(s->htab);
        free(ix->names[i]);
    }
    free(ix->names); free(ix->segs); free(ix);
}

int orc_num_segments(const orc_index* ix) { return ix->nseg; }
const char* orc_segment_name(const orc_index* ix, int i) { return ix->names[i]; }
const char* orc_cord_uid(const orc_index* ix, uint32_t seg, uint32_t doc) {
    if ((int)seg >= ix->nseg || doc >= ix->segs[seg].ndocs) return "";
    return ix->segs[seg].uid[doc];
}
void orc_segment_stats(const orc_index* ix, int i, uint32_t* N, float* avgdl, uint32_t* T, uint64_t* P) {
    const orc_seg* s = &ix->segs[i];
    *N = s->N; *avgdl = s->avgdl; *T = s->nlex; *P = s->npost;
}
int orc_term_stats(const orc_index* ix, int i, const char* term, uint32_t* df, uint32_t* count) {
    const orc_lex* e = lex_find(&ix->segs[i], term);
    *df = e ? e->df : 0; *count = e ? e->count : 0;
    return e != NULL;
}

/* include/textutil.hpp:13-28 (tokenize) + :31-37 (is_stopword) + src/api_engine.cpp:391-397.
 * isalnum/tolower in the "C" locale: ASCII only. Returns number of kept terms (max `cap`). */
static const char* const STOP[] = {"the","a","an","and","or","of","to","in","for","on","with","by","as",
                                   "is","are","was","were","be","been","it","this","that","from","at"};
static int is_stop(const char* t) {
    for (size_t i = 0; i < sizeof STOP / sizeof STOP[0]; i++) if (strcmp(t, STOP[i]) == 0) return 1;
    return 0;
}
int orc_query_terms(const char* q, char** out, int cap) {
    int n = 0;
    size_t len = strlen(q), cur = 0;
    char* buf = (char*)malloc(len + 2);
    for (size_t i = 0; i <= len; i++) {
        unsigned char c = (unsigned char)q[i];
        int alnum = (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
        if (i < len && alnum) {
            buf[cur++] = (char)((c >= 'A' && c <= 'Z') ? c - 'A' + 'a' : c);
        } else if (cur) {
            buf[cur] = 0;
            if (cur >= 2 && !is_stop(buf) && n < cap) out[n++] = strdup(buf);
            cur = 0;
        }
    }
    free(buf);
    return n;
}

/* src/api_engine.cpp:45-47 */
static float bm25_idf(uint32_t N, uint32_t df) {
    return logf((((float)(uint32_t)(N - df) + 0.5f) / ((float)df + 0.5f)) + 1.0f);
}

typedef struct { float s; uint32_t seg, doc; } orc_hit;

static int hit_before(const orc_hit* a, const orc_hit* b) {
    if (a->s != b->s) return a->s > b->s;
    if (a->seg != b->seg) return a->seg < b->seg;
    return a->doc < b->doc;
}

typedef struct {
    float* score;       /* [max_docs] */
    uint8_t* touched;   /* [max_docs] */
    uint32_t* list;     /* touched docIds in first-touch order */
} orc_scratch;

static orc_scratch* scratch_new(const orc_index* ix) {
    orc_scratch* sc = (orc_scratch*)calloc(1, sizeof *sc);
    sc->score = (float*)calloc((size_t)ix->max_docs + 1, sizeof(float));
    sc->touched = (uint8_t*)calloc((size_t)ix->max_docs + 1, 1);
    sc->list = (uint32_t*)malloc(((size_t)ix->max_docs + 1) * 4);
    return sc;
}
static void scratch_free(orc_scratch* sc) { free(sc->score); free(sc->touched); free(sc->list); free(sc); }

#define ORC_MAX_TERMS 256

/* The scoring loop proper, src/api_engine.cpp:426-505, over an explicit (term, qweight) list — the
 * reference's qterms_w (:410-421): the query's own terms with weight 1.0f, or SemanticIndex::expand's
 * output when embeddings are loaded (src/semantic_embedding.cpp:148-229).
 * Returns the number of hits written (<= K), K = clamp(k,1,100) (:377). */
static int search_terms_impl(const orc_index* ix, orc_scratch* sc, int nt, char* const* terms, const float* weights,
                             int k, orc_hit* hits, uint64_t* found, int* has_found) {
    const float k1 = 1.2f, b = 0.75f;                                  /* :375-376 */
    const int K = k < 1 ? 1 : (k > 100 ? 100 : k);                     /* :377 */
    *found = 0;
    *has_found = 0;
    if (nt == 0 || ix->nseg == 0) return 0;                            /* :407, :424 */
    *has_found = 1;
    int nh = 0;
    uint64_t total_found = 0;
    for (int si = 0; si < ix->nseg; si++) {                            /* :441 */
        const orc_seg* seg = &ix->segs[si];
        uint32_t ntouched = 0;
        for (int t = 0; t < nt; t++) {                                 /* :449, query order, duplicates kept */
            const float qweight = weights ? weights[t] : 1.0f;         /* :451; 1.0f without expansion (:420) */
            const orc_lex* e = lex_find(seg, terms[t]);                /* :454-455 */
            if (!e) continue;
            if (e->df == 0) continue;                                  /* :458 */
            float idf = bm25_idf(seg->N, e->df);                       /* :461 */
            const uint32_t* p = seg->post + e->begin * 2;
            for (uint32_t i = 0; i < e->count; i++) {                  /* :473 */
                uint32_t docId = p[2 * i], tf = p[2 * i + 1];          /* :474-475 */
                float dl = (float)seg->doc_len[docId];                 /* :477 */
                float denom = (float)tf + k1 * (1.0f - b + b * (dl / seg->avgdl)); /* :478 */
                float s = idf * ((float)tf * (k1 + 1.0f)) / denom;     /* :479 */
                if (!sc->touched[docId]) { sc->touched[docId] = 1; sc->score[docId] = 0.0f; sc->list[ntouched++] = docId; }
                sc->score[docId] += qweight * s;                       /* :480 */
            }
        }
        /* :485-492 under the total order (score desc, seg asc, doc asc) */
        for (uint32_t i = 0; i < ntouched; i++) {
            orc_hit h = {sc->score[sc->list[i]], (uint32_t)si, sc->list[i]};
            if (nh == K && !hit_before(&h, &hits[K - 1])) continue;
            int pos = nh < K ? nh : K - 1;
            while (pos > 0 && hit_before(&h, &hits[pos - 1])) { hits[pos] = hits[pos - 1]; pos--; }
            hits[pos] = h;
            if (nh < K) nh++;
        }
        total_found += ntouched;                                       /* :495 */
        for (uint32_t i = 0; i < ntouched; i++) sc->touched[sc->list[i]] = 0;
    }
    *found = total_found;                                              /* :505 */
    return nh;
}

/* src/api_engine.cpp:369-505 with semantic expansion disabled (:418-421): tokenise, filter, weight 1.0f. */
static int search_impl(const orc_index* ix, orc_scratch* sc, const char* query, int k, orc_hit* hits,
                       uint64_t* found, int* has_found) {
    char* terms[ORC_MAX_TERMS];
    int nt = orc_query_terms(query, terms, ORC_MAX_TERMS);             /* :388-397 */
    int nh = search_terms_impl(ix, sc, nt, terms, NULL, k, hits, found, has_found);
    for (int t = 0; t < nt; t++) free(terms[t]);
    return nh;
}

/* Weighted entry: the caller supplies qterms_w in the reference's order (tests take it from the compiled
 * reference: oracle/ref_driver.cpp dumps SemanticIndex::expand's output next to every result). */
int orc_search_weighted(const orc_index* ix, int nterms, const char* const* terms, const float* weights, int k,
                        float* scores, uint32_t* segs, uint32_t* docs, uint64_t* found, int* has_found) {
    if (nterms < 0 || nterms > ORC_MAX_TERMS) return -1;
    orc_scratch* sc = scratch_new(ix);
    orc_hit hits[100];
    int n = search_terms_impl(ix, sc, nterms, (char* const*)terms, weights, k, hits, found, has_found);
    for (int i = 0; i < n; i++) { scores[i] = hits[i].s; segs[i] = hits[i].seg; docs[i] = hits[i].doc; }
    scratch_free(sc);
    return n;
}

int orc_search(const orc_index* ix, const char* query, int k, float* scores, uint32_t* segs, uint32_t* docs,
               uint64_t* found, int* has_found) {
    orc_scratch* sc = scratch_new(ix);
    orc_hit hits[100];
    int n = search_impl(ix, sc, query, k, hits, found, has_found);
    for (int i = 0; i < n; i++) { scores[i] = hits[i].s; segs[i] = hits[i].seg; docs[i] = hits[i].doc; }
    scratch_free(sc);
    return n;
}

/* Score of one (segment, doc) for a query, for tie-group membership checks against the as-is
 * reference.  *matched = 0 if no query term has a posting for the doc. */
void orc_score_doc(const orc_index* ix, const char* query, uint32_t si, uint32_t doc, float* score, int* matched) {
    const float k1 = 1.2f, b = 0.75f;
    char* terms[ORC_MAX_TERMS];
    int nt = orc_query_terms(query, terms, ORC_MAX_TERMS);
    *score = 0.0f;
    *matched = 0;
    if ((int)si >= ix->nseg) return;
    const orc_seg* seg = &ix->segs[si];
    for (int t = 0; t < nt; t++) {
        const orc_lex* e = lex_find(seg, terms[t]);
        if (!e || e->df == 0) continue;
        float idf = bm25_idf(seg->N, e->df);
        const uint32_t* p = seg->post + e->begin * 2;
        uint32_t lo = 0, hi = e->count;
        while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (p[2 * mid] < doc) lo = mid + 1; else hi = mid; }
        if (lo < e->count && p[2 * lo] == doc) {
            uint32_t tf = p[2 * lo + 1];
            float dl = (float)seg->doc_len[doc];
            float denom = (float)tf + k1 * (1.0f - b + b * (dl / seg->avgdl));
            float s = idf * ((float)tf * (k1 + 1.0f)) / denom;
            if (!*matched) { *matched = 1; *score = 0.0f; }
            *score += 1.0f * s;
        }
    }
    for (int t = 0; t < nt; t++) free(terms[t]);
}

/* ---- batch driver (tests at size, bench.py cpu_baseline "port") ---- */
typedef struct {
    const orc_index* ix;
    const char* const* queries;
    int nq, k, tid, nthreads;
    float* scores; uint32_t* segs; uint32_t* docs; uint32_t* nhits; uint64_t* found; uint8_t* has_found;
    uint64_t postings;
} orc_job;

static void* job_main(void* arg) {
    orc_job* j = (orc_job*)arg;
    orc_scratch* sc = scratch_new(j->ix);
    const int K = j->k < 1 ? 1 : (j->k > 100 ? 100 : j->k);
    orc_hit hits[100];
    for (int q = j->tid; q < j->nq; q += j->nthreads) {
        uint64_t found; int has;
        int n = search_impl(j->ix, sc, j->queries[q], j->k, hits, &found, &has);
        if (j->nhits) j->nhits[q] = (uint32_t)n;
        if (j->found) j->found[q] = found;
        if (j->has_found) j->has_found[q] = (uint8_t)has;
        for (int i = 0; i < n; i++) {
            if (j->scores) j->scores[(size_t)q * K + i] = hits[i].s;
            if (j->segs) j->segs[(size_t)q * K + i] = hits[i].seg;
            if (j->docs) j->docs[(size_t)q * K + i] = hits[i].doc;
        }
    }
    scratch_free(sc);
    return NULL;
}

/* Runs nq queries on nthreads threads; outputs are [nq][K] (K = clamp(k)); any output may be NULL.
 * Returns elapsed wall seconds. */
double orc_search_many(const orc_index* ix, const char* const* queries, int nq, int k, int nthreads,
                       float* scores, uint32_t* segs, uint32_t* docs, uint32_t* nhits, uint64_t* found,
                       uint8_t* has_found) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq && nq > 0) nthreads = nq;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    orc_job* jobs = (orc_job*)calloc((size_t)nthreads, sizeof(orc_job));
    for (int t = 0; t < nthreads; t++) {
        orc_job j = {ix, queries, nq, k, t, nthreads, scores, segs, docs, nhits, found, has_found, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Σ LexEntry.count over (query term occurrence, segment): the posting count behind the
 * "algorithmic bytes" of SURVEY.md §8d (8 B each). */
uint64_t orc_query_postings(const orc_index* ix, const char* query) {
    char* terms[ORC_MAX_TERMS];
    int nt = orc_query_terms(query, terms, ORC_MAX_TERMS);
    uint64_t total = 0;
    for (int si = 0; si < ix->nseg; si++)
        for (int t = 0; t < nt; t++) {
            const orc_lex* e = lex_find(&ix->segs[si], terms[t]);
            if (e && e->df) total += e->count;
        }
    for (int t = 0; t < nt; t++) free(terms[t]);
    return total;
}

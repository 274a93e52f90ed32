/*
 * oracle_c.c — plain-C restatement of the reference's hybrid query path.  TEST INFRASTRUCTURE.
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load
 * it; the product never does.  It restates, on integer-coded columns, the same algorithm as
 * oracle/oracle.py (which cites the sources line by line) and is checked against it in
 * tests/test_oracle_c.py:
 *   - dense: rows L2-normalised in fp32 when loaded (qdrant COSINE normalises at upsert,
 *     vector_store.py:93,:313), query normalised, score = dot (accumulated in double, rounded to
 *     fp32) — distances.py cosine_similarity;
 *   - sparse: query value * ln((N - df + 0.5)/(df + 0.5) + 1), two-pointer merge against EVERY
 *     row's index-sorted vector, double accumulation, fp32 result, no overlap => excluded —
 *     sparse_distances.py sparse_dot_product / local_collection.py _rescore_idf;
 *   - filter: alive && scope bit && inclusive range on the chosen timestamp column, a missing
 *     timestamp fails — vector_store.py:462-530 + payload_filters.py;
 *   - selection: top-k by (score desc, row asc);
 *   - fusion: voitta's min-max weighted sum (vector_store.py:659-697) or Qdrant RRF.
 * PARITY UNPINNED for the qdrant-internal parts (see oracle/oracle.py header).
 * Deviation that only helps the CPU baseline's speed: rows are normalised once at load instead of
 * on every query, and selection uses a heap instead of a full argsort.
 * Threads: OpenMP over rows (B == 1) or queries (B > 1).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TS_MISSING INT64_MIN

typedef struct {
    uint64_t n;
    int32_t dim;
    float* dense;          /* n x dim, unit rows */
    const int64_t* indptr; /* borrowed */
    const uint32_t* terms;
    const float* vals;
    const uint32_t* scope;
    const int64_t* created;
    const int64_t* modified;
    const uint8_t* alive;
    uint64_t n_live;
    /* df hash: open addressing */
    uint32_t* h_key;
    uint32_t* h_df;
    uint8_t* h_used;
    uint64_t h_cap;
} orc_corpus;

typedef struct {
    const uint32_t* scope_bits;
    uint32_t scope_words;
    int32_t ts_field; /* 0 none, 1 created, 2 modified */
    int64_t ts_lo, ts_hi;
} orc_filter;

typedef struct { float s; uint32_t r; } cand_t;

static uint64_t hash32(uint32_t x) { uint64_t h = x * 0x9E3779B97F4A7C15ull; return h ^ (h >> 29); }

static void df_add(orc_corpus* c, uint32_t t) {
    uint64_t i = hash32(t) & (c->h_cap - 1);
    while (c->h_used[i] && c->h_key[i] != t) i = (i + 1) & (c->h_cap - 1);
    if (!c->h_used[i]) { c->h_used[i] = 1; c->h_key[i] = t; c->h_df[i] = 0; }
    c->h_df[i] += 1;
}
static uint32_t df_get(const orc_corpus* c, uint32_t t) {
    if (!c->h_cap) return 0;
    uint64_t i = hash32(t) & (c->h_cap - 1);
    while (c->h_used[i]) { if (c->h_key[i] == t) return c->h_df[i]; i = (i + 1) & (c->h_cap - 1); }
    return 0;
}

orc_corpus* orc_build(uint64_t n, int32_t dim, const float* dense, const int64_t* indptr, const uint32_t* terms,
                      const float* vals, const uint32_t* scope, const int64_t* created, const int64_t* modified,
                      const uint8_t* alive) {
    orc_corpus* c = (orc_corpus*)calloc(1, sizeof *c);
    c->n = n; c->dim = dim; c->indptr = indptr; c->terms = terms; c->vals = vals; c->scope = scope;
    c->created = created; c->modified = modified; c->alive = alive;
    c->dense = (float*)malloc((size_t)n * dim * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)n; ++r) {
        const float* src = dense + (size_t)r * dim;
        float* dst = c->dense + (size_t)r * dim;
        double ss = 0.0;
        for (int j = 0; j < dim; ++j) ss += (double)src[j] * src[j];
        const float nrm = (float)sqrt(ss);
        for (int j = 0; j < dim; ++j) dst[j] = nrm > 0.f ? src[j] / nrm : src[j];
    }
    uint64_t live = 0, nnz_live = 0;
    for (uint64_t r = 0; r < n; ++r) if (!alive || alive[r]) { ++live; if (indptr) nnz_live += (uint64_t)(indptr[r + 1] - indptr[r]); }
    c->n_live = live;
    if (indptr && nnz_live) {
        uint64_t cap = 1024;
        while (cap < 2 * nnz_live) cap <<= 1;
        if (cap > (1ull << 28)) cap = 1ull << 28;
        c->h_cap = cap;
        c->h_key = (uint32_t*)malloc(cap * 4); c->h_df = (uint32_t*)malloc(cap * 4); c->h_used = (uint8_t*)calloc(cap, 1);
        for (uint64_t r = 0; r < n; ++r) {
            if (alive && !alive[r]) continue;
            for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) df_add(c, terms[p]);
        }
    }
    return c;
}

void orc_free(orc_corpus* c) {
    if (!c) return;
    free(c->dense); free(c->h_key); free(c->h_df); free(c->h_used); free(c);
}

uint64_t orc_df(const orc_corpus* c, uint32_t term) { return df_get(c, term); }

static int better(cand_t a, cand_t b) { return a.s > b.s || (a.s == b.s && a.r < b.r); }

/* min-heap on `better` order: heap[0] is the worst kept candidate */
static void heap_push(cand_t* h, int* n, int k, cand_t x) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = x;
        while (i > 0) { int p = (i - 1) / 2; if (better(h[p], h[i])) { cand_t t = h[p]; h[p] = h[i]; h[i] = t; i = p; } else break; }
    } else if (better(x, h[0])) {
        h[0] = x;
        int i = 0;
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < k && better(h[m], h[l])) m = l;
            if (r < k && better(h[m], h[r])) m = r;
            if (m == i) break;
            cand_t t = h[m]; h[m] = h[i]; h[i] = t; i = m;
        }
    }
}
static int cmp_desc(const void* a, const void* b) {
    cand_t x = *(const cand_t*)a, y = *(const cand_t*)b;
    return better(x, y) ? -1 : (better(y, x) ? 1 : 0);
}

static int passes(const orc_corpus* c, const orc_filter* f, uint64_t r) {
    if (c->alive && !c->alive[r]) return 0;
    if (!f) return 1;
    if (f->scope_bits) {
        const uint32_t s = c->scope ? c->scope[r] : 0;
        if ((s >> 5) >= f->scope_words || !((f->scope_bits[s >> 5] >> (s & 31)) & 1u)) return 0;
    }
    if (f->ts_field) {
        const int64_t* col = f->ts_field == 1 ? c->created : c->modified;
        const int64_t t = col ? col[r] : ORC_TS_MISSING;
        if (t == ORC_TS_MISSING || t < f->ts_lo || t > f->ts_hi) return 0;
    }
    return 1;
}

/* one query, rows [r0, r1): push dense and sparse candidates into the heaps */
static void score_rows(const orc_corpus* c, const orc_filter* f, const float* qn, int nq, const uint32_t* qt,
                       const double* qw, uint64_t r0, uint64_t r1, int k, cand_t* hd, int* nd, cand_t* hs, int* ns) {
    const int dim = c->dim;
    for (uint64_t r = r0; r < r1; ++r) {
        if (!passes(c, f, r)) continue;
        const float* v = c->dense + (size_t)r * dim;
        double acc = 0.0;
        for (int j = 0; j < dim; ++j) acc += (double)v[j] * (double)qn[j];
        cand_t x = {(float)acc, (uint32_t)r};
        x.s += 0.0f;
        heap_push(hd, nd, k, x);
        if (nq && c->indptr) {
            int64_t p = c->indptr[r], pe = c->indptr[r + 1];
            int i = 0, overlap = 0;
            double res = 0.0;
            while (i < nq && p < pe) {
                const uint32_t a = qt[i], b = c->terms[p];
                if (a == b) { overlap = 1; res += qw[i] * (double)c->vals[p]; ++i; ++p; }
                else if (a < b) ++i; else ++p;
            }
            if (overlap) { cand_t y = {(float)res, (uint32_t)r}; y.s += 0.0f; heap_push(hs, ns, k, y); }
        }
    }
}

static int fuse(int mode, double w, const cand_t* d, int nd, const cand_t* s, int ns, int limit,
                uint64_t* out_rows, double* out_scores) {
    /* candidates in first-seen order: dense list, then sparse-only */
    int m = 0;
    const int cap = nd + ns;
    double* fin = (double*)malloc(sizeof(double) * (cap > 0 ? cap : 1));
    uint32_t* row = (uint32_t*)malloc(sizeof(uint32_t) * (cap > 0 ? cap : 1));
    const double dw = 1.0 - w;
    double dmin = 0, dsp = 0, smin = 0, ssp = 0;
    if (nd) { dmin = d[nd - 1].s; dsp = (double)d[0].s - dmin; }
    if (ns) { smin = s[ns - 1].s; ssp = (double)s[0].s - smin; }
    for (int i = 0; i < nd; ++i) {
        int j = -1;
        for (int t = 0; t < ns; ++t) if (s[t].r == d[i].r) { j = t; break; }
        double f;
        if (mode == 1) {
            const double dn = dsp > 0 ? ((double)d[i].s - dmin) / dsp : 1.0;
            const double sn = j < 0 ? 0.0 : (ssp > 0 ? ((double)s[j].s - smin) / ssp : 1.0);
            const double a = dw * dn, b = w * sn;
            f = a + b;
        } else {
            f = 1.0 / (double)(2 + i);
            if (j >= 0) f += 1.0 / (double)(2 + j);
        }
        fin[m] = f; row[m] = d[i].r; ++m;
    }
    for (int j = 0; j < ns; ++j) {
        int in_d = 0;
        for (int t = 0; t < nd; ++t) if (d[t].r == s[j].r) { in_d = 1; break; }
        if (in_d) continue;
        double f;
        if (mode == 1) {
            const double sn = ssp > 0 ? ((double)s[j].s - smin) / ssp : 1.0;
            const double a = dw * 0.0, b = w * sn;
            f = a + b;
        } else f = 1.0 / (double)(2 + j);
        fin[m] = f; row[m] = s[j].r; ++m;
    }
    /* stable selection of the top `limit` (descending, first-seen order on ties) */
    int out = 0;
    char* used = (char*)calloc(m > 0 ? m : 1, 1);
    while (out < limit && out < m) {
        int best = -1;
        for (int i = 0; i < m; ++i) if (!used[i] && (best < 0 || fin[i] > fin[best])) best = i;
        used[best] = 1;
        out_rows[out] = row[best]; out_scores[out] = fin[best]; ++out;
    }
    free(used); free(fin); free(row);
    return out;
}

/* Batch search.  q_vals are the raw query values (IDF applied here iff apply_idf).  Outputs like vb_search. */
int orc_search(const orc_corpus* c, uint32_t B, const float* q, const int64_t* q_indptr, const uint32_t* q_terms,
               const double* q_vals, uint32_t n_filters, const orc_filter* filters, const int32_t* filter_of,
               uint32_t limit, uint32_t kprime, int32_t fusion, double w,
               uint64_t* out_rows, double* out_scores, int32_t* out_counts,
               uint64_t* d_rows, float* d_scores, int32_t* d_counts,
               uint64_t* s_rows, float* s_scores, int32_t* s_counts, int32_t n_threads, int32_t apply_idf) {
    (void)n_filters;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    const int T = omp_get_max_threads();
#else
    const int T = 1;
    (void)n_threads;
#endif
    const int dim = c->dim, k = (int)kprime;
    for (uint32_t b0 = 0; b0 < B; ++b0) { out_counts[b0] = 0; }
    const int par_queries = B >= (uint32_t)T;
#pragma omp parallel for schedule(dynamic, 1) if (par_queries)
    for (int64_t b = 0; b < (int64_t)B; ++b) {
        const orc_filter* f = (filter_of && filter_of[b] >= 0) ? &filters[filter_of[b]] : NULL;
        /* normalise query */
        float* qn = (float*)malloc(sizeof(float) * dim);
        double ss = 0.0;
        for (int j = 0; j < dim; ++j) ss += (double)q[(size_t)b * dim + j] * q[(size_t)b * dim + j];
        const float nrm = (float)sqrt(ss);
        for (int j = 0; j < dim; ++j) qn[j] = nrm > 0.f ? q[(size_t)b * dim + j] / nrm : 0.f;
        /* sparse query: sort by term, apply idf */
        int nq = 0;
        uint32_t* qt = NULL; double* qw = NULL;
        if (fusion != 0 && q_indptr && q_indptr[b + 1] > q_indptr[b]) {
            nq = (int)(q_indptr[b + 1] - q_indptr[b]);
            qt = (uint32_t*)malloc(4 * nq); qw = (double*)malloc(8 * nq);
            for (int i = 0; i < nq; ++i) { qt[i] = q_terms[q_indptr[b] + i]; qw[i] = q_vals[q_indptr[b] + i]; }
            for (int i = 1; i < nq; ++i) {      /* insertion sort by term id */
                uint32_t t = qt[i]; double v = qw[i]; int j = i - 1;
                while (j >= 0 && qt[j] > t) { qt[j + 1] = qt[j]; qw[j + 1] = qw[j]; --j; }
                qt[j + 1] = t; qw[j + 1] = v;
            }
            for (int i = 0; apply_idf && i < nq; ++i) {
                const double df = (double)df_get(c, qt[i]);
                qw[i] = qw[i] * log(((double)c->n_live - df + 0.5) / (df + 0.5) + 1.0);
            }
        }
        cand_t* hd = (cand_t*)malloc(sizeof(cand_t) * k * T);
        cand_t* hs = (cand_t*)malloc(sizeof(cand_t) * k * T);
        int* nd = (int*)calloc(T, sizeof(int));
        int* ns = (int*)calloc(T, sizeof(int));
        if (par_queries) {
            score_rows(c, f, qn, nq, qt, qw, 0, c->n, k, hd, &nd[0], hs, &ns[0]);
        } else {
#pragma omp parallel
            {
#ifdef _OPENMP
                const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
                const int t = 0, nt = 1;
#endif
                const uint64_t r0 = c->n * t / nt, r1 = c->n * (t + 1) / nt;
                score_rows(c, f, qn, nq, qt, qw, r0, r1, k, hd + (size_t)t * k, &nd[t], hs + (size_t)t * k, &ns[t]);
            }
            /* merge per-thread heaps into slot 0 */
            for (int t = 1; t < T; ++t) {
                for (int i = 0; i < nd[t]; ++i) heap_push(hd, &nd[0], k, hd[(size_t)t * k + i]);
                for (int i = 0; i < ns[t]; ++i) heap_push(hs, &ns[0], k, hs[(size_t)t * k + i]);
            }
        }
        qsort(hd, nd[0], sizeof(cand_t), cmp_desc);
        qsort(hs, ns[0], sizeof(cand_t), cmp_desc);
        if (d_counts) d_counts[b] = nd[0];
        if (s_counts) s_counts[b] = ns[0];
        for (int i = 0; i < nd[0]; ++i) { if (d_rows) d_rows[(size_t)b * k + i] = hd[i].r; if (d_scores) d_scores[(size_t)b * k + i] = hd[i].s; }
        for (int i = 0; i < ns[0]; ++i) { if (s_rows) s_rows[(size_t)b * k + i] = hs[i].r; if (s_scores) s_scores[(size_t)b * k + i] = hs[i].s; }
        if (nq == 0) {
            const int m = nd[0] < (int)limit ? nd[0] : (int)limit;
            for (int i = 0; i < m; ++i) { out_rows[(size_t)b * limit + i] = hd[i].r; out_scores[(size_t)b * limit + i] = (double)hd[i].s; }
            out_counts[b] = m;
        } else {
            out_counts[b] = fuse(fusion, w, hd, nd[0], hs, ns[0], (int)limit, out_rows + (size_t)b * limit, out_scores + (size_t)b * limit);
        }
        free(hd); free(hs); free(nd); free(ns); free(qn); free(qt); free(qw);
    }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

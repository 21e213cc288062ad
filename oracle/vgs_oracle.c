/*
 * vgs_oracle.c -- plain-C CPU restatement of the reference's VLMC distance path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the oracle's fast arm: the same algorithms as
 * oracle/vlmc_oracle.py (which is pinned against the compiled reference and the golden vectors),
 * on the flat model arrays, so that parity tests at thousands of models and bench.py's
 * `cpu_baseline` ("port") leg finish in seconds.  It is never linked into or loaded by the product
 * library (libvgs.so); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs load
 * it, and only as the checker.
 *
 * Parity status: PINNED through tests/test_oracle_vs_goldens.py (every entry point below is checked
 * against tests/golden/, i.e. against outputs of the reference itself).
 *
 * Reference citations are relative to /root/reference.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n_ctx;
    int order;              /* src/vlmc/vlmc.pyx:179-180 */
    const uint8_t *len;     /* [n_ctx] */
    const uint32_t *code;   /* [n_ctx] newest symbol in the two lowest bits */
    const double *prob;     /* [n_ctx][4] */
    double *logp;           /* [n_ctx][4]  ln p, or -1000 where p == 0 (vlmc.pyx:94-99) */
    uint32_t hmask;         /* hash capacity - 1 */
    uint32_t *hkey;         /* (1 << 2*len) | code, 0 = empty */
    int32_t *hval;
} vo_model;

typedef struct {
    int64_t n_models;
    vo_model *m;
} vo_set;

static inline uint32_t vo_key(int len, uint32_t code) { return (1u << (2 * len)) | code; }
static inline uint32_t vo_hash(uint32_t k) { k *= 0x9E3779B1u; return k ^ (k >> 15); }

static int32_t vo_find(const vo_model *m, int len, uint32_t code) {
    uint32_t k = vo_key(len, code), s = vo_hash(k) & m->hmask;
    while (m->hkey[s]) {
        if (m->hkey[s] == k) return m->hval[s];
        s = (s + 1) & m->hmask;
    }
    return -1;
}

void *vo_create(int64_t n_models, const int64_t *ctx_offsets, const uint8_t *ctx_len,
                const uint32_t *ctx_code, const double *prob) {
    vo_set *S = (vo_set *)calloc(1, sizeof(vo_set));
    S->n_models = n_models;
    S->m = (vo_model *)calloc((size_t)n_models, sizeof(vo_model));
    for (int64_t i = 0; i < n_models; ++i) {
        vo_model *m = &S->m[i];
        int64_t a = ctx_offsets[i], n = ctx_offsets[i + 1] - a;
        m->n_ctx = n;
        m->len = ctx_len + a;
        m->code = ctx_code + a;
        m->prob = prob + 4 * a;
        m->logp = (double *)malloc(sizeof(double) * 4 * (size_t)n);
        uint32_t cap = 8;
        while (cap < 2 * n) cap <<= 1;
        m->hmask = cap - 1;
        m->hkey = (uint32_t *)calloc(cap, sizeof(uint32_t));
        m->hval = (int32_t *)calloc(cap, sizeof(int32_t));
        m->order = 0;
        for (int64_t c = 0; c < n; ++c) {
            if (m->len[c] > m->order) m->order = m->len[c];
            for (int x = 0; x < 4; ++x) {
                double p = m->prob[4 * c + x];
                m->logp[4 * c + x] = (p == 0) ? -1000.0 : log(p);
            }
            uint32_t k = vo_key(m->len[c], m->code[c]), s = vo_hash(k) & m->hmask;
            while (m->hkey[s] && m->hkey[s] != k) s = (s + 1) & m->hmask;
            if (!m->hkey[s]) { m->hkey[s] = k; m->hval[s] = (int32_t)c; } /* dict: first position of a key */
        }
    }
    return S;
}

void vo_destroy(void *h) {
    vo_set *S = (vo_set *)h;
    if (!S) return;
    for (int64_t i = 0; i < S->n_models; ++i) { free(S->m[i].logp); free(S->m[i].hkey); free(S->m[i].hval); }
    free(S->m);
    free(S);
}

int vo_order(void *h, int64_t model) { return ((vo_set *)h)->m[model].order; }

/* src/vlmc/vlmc.pyx:131-141 -- longest suffix of the history (length <= order) that is a context. */
int32_t vo_lookup(void *h, int64_t model, uint32_t hist_code, int hist_len) {
    const vo_model *m = &((vo_set *)h)->m[model];
    int top = hist_len < m->order ? hist_len : m->order;
    for (int ln = top; ln >= 0; --ln) {
        uint32_t c = ln >= 16 ? hist_code : (hist_code & ((1u << (2 * ln)) - 1u));
        int32_t id = vo_find(m, ln, c);
        if (id >= 0) return id;
    }
    return -1; /* RuntimeError("get_context vlmc.pyx") */
}

/* src/vlmc/vlmc.pyx:143-154 -- bit l set <=> the length-l suffix is a context. */
uint32_t vo_lookup_all(void *h, int64_t model, uint32_t hist_code, int hist_len) {
    const vo_model *m = &((vo_set *)h)->m[model];
    int top = hist_len < m->order ? hist_len : m->order;
    uint32_t mask = 0;
    for (int ln = top; ln >= 0; --ln)
        if (vo_find(m, ln, hist_code & ((1u << (2 * ln)) - 1u)) >= 0) mask |= 1u << ln;
    return mask;
}

/* src/vlmc/vlmc.pyx:87-101.  skip < 0 means "order" (log_likelihood_ignore_initial_bias, :79-82).
 * Returns NaN if some history has no matching context (reference: RuntimeError). */
double vo_loglik(void *h, int64_t model, const uint8_t *codes, int64_t L, int64_t skip) {
    const vo_model *m = &((vo_set *)h)->m[model];
    if (skip < 0) skip = m->order;
    double acc = 0.0;
    uint32_t hist = 0;
    uint32_t omask = m->order >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m->order)) - 1u);
    for (int64_t t = 0; t < L; ++t) {
        if (t >= skip) {
            int hl = t < m->order ? (int)t : m->order;
            int32_t id = vo_lookup(h, model, hist, hl);
            if (id < 0) return NAN;
            acc += m->logp[4 * id + codes[t]];
        }
        hist = ((hist << 2) | codes[t]) & omask;
    }
    return acc;
}

/* M[s][m]; codes is [n_seq][L]; skip as above. */
void vo_score_matrix(void *h, const uint8_t *codes, int64_t n_seq, int64_t L, int64_t skip, double *M) {
    vo_set *S = (vo_set *)h;
    int64_t nm = S->n_models;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t k = 0; k < n_seq * nm; ++k) {
        int64_t s = k / nm, mi = k % nm;
        M[k] = vo_loglik(h, mi, codes + s * L, L, skip);
    }
}

/* src/distance/negloglikelihood.pyx:17-26, same operation order. */
void vo_nll_from_scores(const double *M, int64_t n, int64_t length, double *D) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) {
            double lr = (M[i * n + i] - M[i * n + j]) / (double)length;
            double rl = (M[j * n + j] - M[j * n + i]) / (double)length;
            D[i * n + j] = (lr + rl) / 2;
        }
}

/* src/distance/frobenius.pyx:21-53: float32 rows, float32 subtraction, then an fp64 sum of squares
 * (the reference's float32 BLAS dot has an implementation-defined order; see vlmc_oracle.frobenius).
 * mode 0 = intersection, 1 = union with longest-suffix fallback, 2 = Projection (absent -> zero row,
 * no normalisation, src/distance/projection.pyx:25-36). */
double vo_pair_distance(void *h, int64_t i, int64_t j, int mode) {
    vo_set *S = (vo_set *)h;
    const vo_model *a = &S->m[i], *b = &S->m[j];
    double ssq = 0.0;
    int64_t count = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const vo_model *x = pass ? b : a, *y = pass ? a : b;
        int64_t yi = pass ? i : j;
        for (int64_t c = 0; c < x->n_ctx; ++c) {
            int32_t id = vo_find(y, x->len[c], x->code[c]);
            if (vo_find(x, x->len[c], x->code[c]) != (int32_t)c) continue; /* duplicate key: dict keeps one */
            if (id >= 0) {
                if (pass) continue; /* shared contexts are counted once */
            } else {
                if (mode == 0) continue;
                if (mode == 1) {
                    id = vo_lookup(h, yi, x->code[c], x->len[c]);
                    if (id < 0) return NAN;
                }
            }
            ++count;
            for (int k = 0; k < 4; ++k) {
                float px = (float)x->prob[4 * c + k];
                float py = id >= 0 ? (float)y->prob[4 * id + k] : 0.0f;
                float d = px - py;
                ssq += (double)d * (double)d;
            }
        }
    }
    if (mode == 2) return sqrt(ssq);
    if (count == 0) return NAN;
    return sqrt(ssq) / sqrt((double)count);
}

void vo_pair_matrix(void *h, int mode, double *D) {
    vo_set *S = (vo_set *)h;
    int64_t n = S->n_models;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n * n; ++k) D[k] = vo_pair_distance(h, k / n, k % n, mode);
}

/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h, DOUBLE_pairwise_sum), which is
 * what scipy's power_divergence `terms.sum(axis=0)` runs (src/distance/estimate.pyx:99). */
static double vo_pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return vo_pairwise_sum(a, n2) + vo_pairwise_sum(a + n2, n - n2);
    }
}

double vo_np_sum(const double *a, int64_t n) { return vo_pairwise_sum(a, n); }

/* src/distance/estimate.pyx:19-108 -- left's contexts counted along `codes` (the right model's
 * sequence); pseudo-counts; Pearson statistic over visited contexts. */
double vo_estimate(void *h, int64_t left, const uint8_t *codes, int64_t L) {
    const vo_model *m = &((vo_set *)h)->m[left];
    int64_t *cnt = (int64_t *)calloc((size_t)(4 * m->n_ctx), sizeof(int64_t));
    uint32_t hist = 0;
    uint32_t omask = m->order >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m->order)) - 1u);
    for (int64_t t = 0; t < L; ++t) {
        int top = t < m->order ? (int)t : m->order;
        for (int ln = top; ln >= 0; --ln) {
            int32_t id = vo_find(m, ln, hist & ((1u << (2 * ln)) - 1u));
            if (id >= 0) cnt[4 * id + codes[t]]++;
        }
        hist = ((hist << 2) | codes[t]) & omask;
    }
    double *terms = (double *)malloc(sizeof(double) * 4 * (size_t)m->n_ctx);
    int64_t nt = 0;
    for (int64_t c = 0; c < m->n_ctx; ++c) {
        if (vo_find(m, m->len[c], m->code[c]) != (int32_t)c) continue;
        int64_t *k = cnt + 4 * c;
        int64_t tot = k[0] + k[1] + k[2] + k[3];
        if (tot != 0 && (k[0] == 0 || k[1] == 0 || k[2] == 0 || k[3] == 0)) { /* :73-78 */
            for (int x = 0; x < 4; ++x) k[x] += 1;
            tot += 4;
        }
        if (tot <= 0) continue;
        int last = -1;
        for (int x = 0; x < 4; ++x) if (m->prob[4 * c + x] > 0) last = x;
        for (int x = 0; x < 4; ++x) {
            double p = m->prob[4 * c + x];
            if (!(p > 0) || x == last) continue;
            double obs = (double)k[x] / (double)tot;
            double d = obs - p;
            terms[nt++] = d * d / p;
        }
    }
    double r = vo_pairwise_sum(terms, nt);
    free(terms);
    free(cnt);
    return r;
}

/* codes: [n_seq][L]; D[i][j] = estimate(left = i, right sequence = j) */
void vo_estimate_matrix(void *h, const uint8_t *codes, int64_t n_seq, int64_t L, double *D) {
    vo_set *S = (vo_set *)h;
    int64_t n = S->n_models;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < n * n_seq; ++k) D[k] = vo_estimate(h, k / n_seq, codes + (k % n_seq) * L, L);
}

/* ---- MT19937 exactly as CPython's _randommodule.c (random.random() == genrand_res53) ---------- */
typedef struct { uint32_t mt[624]; int idx; } vo_mt;

static uint32_t vo_mt_next(vo_mt *s) {
    if (s->idx >= 624) {
        uint32_t *mt = s->mt;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double vo_mt_random(vo_mt *s) {
    uint32_t a = vo_mt_next(s) >> 5, b = vo_mt_next(s) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

/* state: 625 words = random.getstate()[1] (624 words + position).  Fills out[n] and updates state. */
void vo_mt_uniforms(uint32_t *state, double *out, int64_t n) {
    vo_mt s;
    memcpy(s.mt, state, 624 * sizeof(uint32_t));
    s.idx = (int)state[624];
    for (int64_t i = 0; i < n; ++i) out[i] = vo_mt_random(&s);
    memcpy(state, s.mt, 624 * sizeof(uint32_t));
    state[624] = (uint32_t)s.idx;
}

/* src/vlmc/vlmc.pyx:156-177 with random.choices spelled out (see vlmc_oracle.generate_sequence):
 * total symbols = length + pre; keeps the last `length`.  Returns 0, or -1 on a failed lookup. */
int vo_sample(void *h, int64_t model, uint32_t *state, int64_t length, int64_t pre, uint8_t *out) {
    const vo_model *m = &((vo_set *)h)->m[model];
    vo_mt s;
    memcpy(s.mt, state, 624 * sizeof(uint32_t));
    s.idx = (int)state[624];
    uint32_t hist = 0;
    uint32_t omask = m->order >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m->order)) - 1u);
    for (int64_t t = 0; t < length + pre; ++t) {
        int hl = t < m->order ? (int)t : m->order;
        int32_t id = vo_lookup(h, model, hist, hl);
        if (id < 0) return -1;
        const double *p = m->prob + 4 * id;
        double c0 = p[0], c1 = c0 + p[1], c2 = c1 + p[2], c3 = c2 + p[3];
        double u = vo_mt_random(&s) * (c3 + 0.0);
        uint8_t x = (u < c0) ? 0 : (u < c1) ? 1 : (u < c2) ? 2 : 3; /* bisect_right(cum, u, 0, 3) */
        if (t >= pre) out[t - pre] = x;
        hist = ((hist << 2) | x) & omask;
    }
    memcpy(state, s.mt, 624 * sizeof(uint32_t));
    state[624] = (uint32_t)s.idx;
    return 0;
}

/* Average-link clustering with the reference's exact semantics (src/clustering/average_link_clustering.pyx:21-146),
 * in O(n^2) memory and ~O(n^2) time with a nearest-neighbour cache -- the scalable oracle for row f1.
 *
 *  - _find_min_edge (:62-80): scan the clusters in dict insertion order (singletons 0..n-1, every merged cluster
 *    appended at the end); a cluster's candidate is the top of its heap = min over the other clusters of the tuple
 *    (distance, key); keys are sorted member tuples of disjoint clusters, so the tuple order is the order of the
 *    smallest member; a candidate replaces the best one only if strictly smaller ('<', :75): first in dict order wins;
 *  - _new_dist (:119-122) with numpy's scalar promotion (NEP 50) as it runs in this container: entries between two
 *    never-merged singletons are numpy.float32, entries that came out of an earlier merge are Python floats
 *    (cdef double return); int * float32 -> float32, float32 + pyfloat -> float32 (the pyfloat is cast first),
 *    float32 / int -> float32, pyfloat arithmetic -> float64;
 *  - other clusters' rows are updated from THEIR distances to left/right (_merge_weights :101-117), the merged
 *    cluster's own row from left's and right's rows (_merge_heaps :124-146): a non-symmetric input stays consistent.
 * labels[i] = index of i's cluster when clusters are numbered by their smallest member; merges[n-k] = merge distances. */
static double vo_new_dist(int64_t nl, double ld, int l32, int64_t nr, double rd, int r32) {
    if (l32 && r32) {
        float a = (float)nl * (float)ld, b = (float)nr * (float)rd;
        float s = a + b;
        s = s / (float)(nl + nr);
        return (double)s;
    }
    if (l32 || r32) {
        float a = l32 ? (float)nl * (float)ld : (float)((double)nl * ld);
        float b = r32 ? (float)nr * (float)rd : (float)((double)nr * rd);
        float s = a + b;
        s = s / (float)(nl + nr);
        return (double)s;
    }
    double s = (double)nl * ld + (double)nr * rd;
    return s / (double)(nl + nr);
}

static void vo_rescan(const double *D, int64_t n, const int *alive, const int64_t *minmem, int64_t i, double *nn_d, int64_t *nn_j) {
    double bd = INFINITY;
    int64_t bj = -1, bm = -1;
    for (int64_t j = 0; j < n; ++j) {
        if (!alive[j] || j == i) continue;
        double d = D[i * n + j];
        if (bj < 0 || d < bd || (d == bd && minmem[j] < bm)) { bd = d; bj = j; bm = minmem[j]; }
    }
    *nn_d = bd; *nn_j = bj;
}

int64_t vo_average_link(const float *D32, int64_t n, int64_t k, int64_t *labels, double *merges) {
    double *D = (double *)malloc(sizeof(double) * (size_t)(n * n));
    int64_t *size = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *minmem = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *okey = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *nn_j = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *parent = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int *alive = (int *)malloc(sizeof(int) * (size_t)n), *single = (int *)malloc(sizeof(int) * (size_t)n);
    double *nn_d = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n * n; ++i) D[i] = D32[i];
    for (int64_t i = 0; i < n; ++i) { size[i] = 1; alive[i] = 1; single[i] = 1; minmem[i] = i; okey[i] = i; parent[i] = i; }
    for (int64_t i = 0; i < n; ++i) vo_rescan(D, n, alive, minmem, i, &nn_d[i], &nn_j[i]);
    int64_t n_merges = 0;
    for (int64_t step = 0; step < n - k; ++step) {
        double best = INFINITY;
        int64_t bl = -1, bkey = -1;
        for (int64_t i = 0; i < n; ++i) {
            if (!alive[i] || nn_j[i] < 0) continue;
            if (nn_d[i] < best || (nn_d[i] == best && bl >= 0 && okey[i] < bkey && 0)) { best = nn_d[i]; bl = i; bkey = okey[i]; }
        }
        /* strict '<' with a scan in dict order: among equal distances the smallest dict position wins */
        for (int64_t i = 0; i < n; ++i)
            if (alive[i] && nn_j[i] >= 0 && nn_d[i] == best && okey[i] < bkey) { bl = i; bkey = okey[i]; }
        if (bl < 0 || !(best < INFINITY)) break;
        const int64_t l = bl, r = nn_j[l];
        merges[n_merges++] = best;
        const int64_t nl = size[l], nr = size[r];
        /* other rows first (they read D[x][l], D[x][r]), then the merged row (reads D[l][x], D[r][x]) */
        double *newrow = (double *)malloc(sizeof(double) * (size_t)n);
        for (int64_t x = 0; x < n; ++x) {
            if (!alive[x] || x == l || x == r) continue;
            const double vx = vo_new_dist(nl, D[x * n + l], single[x] && single[l], nr, D[x * n + r], single[x] && single[r]);
            newrow[x] = vo_new_dist(nl, D[l * n + x], single[x] && single[l], nr, D[r * n + x], single[x] && single[r]);
            D[x * n + l] = vx;
        }
        for (int64_t x = 0; x < n; ++x)
            if (alive[x] && x != l && x != r) D[l * n + x] = newrow[x];
        free(newrow);
        alive[r] = 0;
        single[l] = 0;
        size[l] = nl + nr;
        if (minmem[r] < minmem[l]) minmem[l] = minmem[r];
        okey[l] = n + step;
        parent[r] = l;
        for (int64_t x = 0; x < n; ++x) {
            if (!alive[x]) continue;
            if (x == l || nn_j[x] == l || nn_j[x] == r) vo_rescan(D, n, alive, minmem, x, &nn_d[x], &nn_j[x]);
            else {
                const double d = D[x * n + l];
                if (d < nn_d[x] || (d == nn_d[x] && minmem[l] < minmem[nn_j[x]])) { nn_d[x] = d; nn_j[x] = l; }
            }
        }
    }
    /* labels: clusters numbered by smallest member */
    for (int64_t i = 0; i < n; ++i) { int64_t x = i; while (parent[x] != x) x = parent[x]; labels[i] = minmem[x]; }
    int64_t next = 0;
    int64_t *remap = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) remap[i] = -1;
    for (int64_t i = 0; i < n; ++i) if (remap[labels[i]] < 0 && labels[i] == i) remap[i] = next++;
    for (int64_t i = 0; i < n; ++i) labels[i] = remap[labels[i]];
    free(remap); free(D); free(size); free(minmem); free(okey); free(nn_j); free(parent); free(alive); free(single); free(nn_d);
    return n_merges;
}

/*
 * gnx_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See gnx_oracle.h.
 *
 * Restates, in plain C with the reference's int64 arithmetic, loop order and tie-break,
 * the DP fills and tracebacks of /root/reference/align (gonomics @ bd66b49b).  Each
 * function cites the Go lines it follows.  Storage is flattened (Go uses slices of
 * slices); arithmetic order and comparison order are kept so every tie resolves the
 * same way.
 */
#include "gnx_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- align/align.go:76-84 tripleMaxTrace: ties prefer a (ColM), then b (ColI), then c (ColD) */
static inline int64_t tmt(int64_t a, int64_t b, int64_t c, uint8_t *k)
{
    if (a >= b && a >= c) {
        *k = 0;
        return a;
    } else if (b >= c) {
        *k = 1;
        return b;
    }
    *k = 2;
    return c;
}

static int bases_ok(const uint8_t *s, int64_t len, int dim)
{
    for (int64_t i = 0; i < len; i++)
        if (s[i] >= (uint8_t)dim)
            return 0;
    return 1;
}

/* run-length route builder shared by all tracebacks.  Mirrors the Go idiom
 * `route := make([]Cigar, 1)` + "RunLength==0 -> start / same Op -> ++ / else append"
 * (affineGap_highMem.go:58,63-71; constGap_highMem.go:44-54; affineGap.go:311-319). */
typedef struct {
    orc_cigar *v;
    int64_t len, cap;
    int owned, oom;
} route_t;

static void route_init(route_t *r)
{
    r->cap = 16;
    r->v = (orc_cigar *)calloc((size_t)r->cap, sizeof(orc_cigar));
    r->len = 1; /* make([]Cigar, 1): one zero element */
    r->owned = 1;
    r->oom = (r->v == NULL);
}

static void route_append(route_t *r, int64_t run, uint8_t op)
{
    if (r->oom)
        return;
    if (r->len == r->cap) {
        int64_t nc = r->cap * 2;
        orc_cigar *nv = (orc_cigar *)realloc(r->v, (size_t)nc * sizeof(orc_cigar));
        if (!nv) {
            r->oom = 1;
            return;
        }
        memset(nv + r->cap, 0, (size_t)(nc - r->cap) * sizeof(orc_cigar));
        r->v = nv;
        r->cap = nc;
    }
    r->v[r->len].run_length = run;
    r->v[r->len].op = op;
    r->len++;
}

/* one traceback step's bookkeeping; idx is the Go `routeIdx` */
static void route_push(route_t *r, int64_t *idx, uint8_t k)
{
    if (r->oom)
        return;
    if (r->v[*idx].run_length == 0) {
        r->v[*idx].run_length = 1;
        r->v[*idx].op = k;
    } else if (r->v[*idx].op == k) {
        r->v[*idx].run_length += 1;
    } else {
        route_append(r, 1, k);
        (*idx)++;
    }
}

/* align/align.go:86-90 reverseCigar */
static void route_reverse(route_t *r)
{
    for (int64_t i = 0, j = r->len - 1; i < j; i++, j--) {
        orc_cigar t = r->v[i];
        r->v[i] = r->v[j];
        r->v[j] = t;
    }
}

static int route_emit(route_t *r, orc_cigar *out, int64_t cap, int64_t *n_out)
{
    int rc = ORC_OK;
    if (r->oom) {
        rc = ORC_ENOMEM;
    } else {
        if (n_out)
            *n_out = r->len;
        if (out) {
            if (r->len > cap)
                rc = ORC_ECAP;
            else
                memcpy(out, r->v, (size_t)r->len * sizeof(orc_cigar));
        }
    }
    free(r->v);
    r->v = NULL;
    return rc;
}

/* =====================================================================================
 * affineGap_highMem (align/affineGap_highMem.go:181-223) generalised over the per-cell
 * match score so AffineGapChunk (:227-272) and multipleAffineGap[Chunk] (:274-353) reuse
 * the identical fill; affineTrace (:57-89) follows.
 * ===================================================================================== */
typedef int64_t (*match_fn)(const void *ctx, int64_t i, int64_t j); /* 1-based cell (i,j) */

typedef struct {
    const uint8_t *alpha, *beta;
    const int64_t *scores;
    int dim;
} pair_ctx;

static int64_t match_pair(const void *c, int64_t i, int64_t j)
{
    const pair_ctx *p = (const pair_ctx *)c;
    return p->scores[(int64_t)p->alpha[i - 1] * p->dim + p->beta[j - 1]];
}

/* rows = n+1, cols = m+1.  step_ext is gapExtend (or gapExtend*chunkSize for the chunk forms). */
static int affine_fill_trace(int64_t n, int64_t m, match_fn sc, const void *ctx, int64_t gap_open,
                             int64_t step_ext, int free_end_gaps, int want_cigar, int64_t *score,
                             orc_cigar *out, int64_t cap, int64_t *n_out)
{
    const int64_t W = m + 1;
    /* initAffineScoringAndTrace (:13-27): 2x3 score rows and a 3 x (n+1) x (m+1) byte trace,
     * allocated (zeroed) per call and per row, as the Go `make` calls do. */
    int64_t *rows = (int64_t *)calloc((size_t)(6 * W), sizeof(int64_t));
    uint8_t **tr[3] = {NULL, NULL, NULL};
    int rc = ORC_OK;
    if (!rows)
        return ORC_ENOMEM;
    if (want_cigar) {
        for (int k = 0; k < 3; k++) {
            tr[k] = (uint8_t **)calloc((size_t)(n + 1), sizeof(uint8_t *));
            if (!tr[k]) {
                rc = ORC_ENOMEM;
                goto done;
            }
            for (int64_t i = 0; i <= n; i++) {
                tr[k][i] = (uint8_t *)calloc((size_t)W, 1);
                if (!tr[k][i]) {
                    rc = ORC_ENOMEM;
                    goto done;
                }
            }
        }
    }
    {
        int64_t *cur[3] = {rows, rows + W, rows + 2 * W};
        int64_t *prev[3] = {rows + 3 * W, rows + 4 * W, rows + 5 * W};
        uint8_t k0, k1, k2;
        for (int64_t i = 0; i <= n; i++) {
            for (int64_t j = 0; j <= m; j++) {
                if (i == 0 && j == 0) { /* :185-192 */
                    cur[0][j] = 0;
                    cur[1][j] = gap_open;
                    cur[2][j] = free_end_gaps ? 0 : gap_open;
                } else if (i == 0) { /* :193-197 */
                    cur[0][j] = ORC_VERY_NEG;
                    cur[1][j] = step_ext + cur[1][j - 1];
                    if (want_cigar)
                        tr[1][i][j] = 1;
                    cur[2][j] = ORC_VERY_NEG;
                } else if (j == 0) { /* :198-206 */
                    cur[0][j] = ORC_VERY_NEG;
                    cur[1][j] = ORC_VERY_NEG;
                    cur[2][j] = (free_end_gaps ? 0 : step_ext) + prev[2][j];
                    if (want_cigar)
                        tr[2][i][j] = 2;
                } else {
                    const int64_t s = sc(ctx, i, j);
                    const int64_t oe = gap_open + step_ext;
                    int64_t v0 = tmt(s + prev[0][j - 1], s + prev[1][j - 1], s + prev[2][j - 1], &k0);
                    int64_t v1 = tmt(oe + cur[0][j - 1], step_ext + cur[1][j - 1], oe + cur[2][j - 1], &k1);
                    int64_t v2;
                    if (free_end_gaps && j == m) /* :207-210 last column, free target overhang */
                        v2 = tmt(prev[0][j], prev[1][j], prev[2][j], &k2);
                    else /* :211-215 */
                        v2 = tmt(oe + prev[0][j], oe + prev[1][j], step_ext + prev[2][j], &k2);
                    cur[0][j] = v0;
                    cur[1][j] = v1;
                    cur[2][j] = v2;
                    if (want_cigar) {
                        tr[0][i][j] = k0;
                        tr[1][i][j] = k1;
                        tr[2][i][j] = k2;
                    }
                }
            }
            if (i < n) { /* :217-219 swap all but after the last row */
                for (int k = 0; k < 3; k++) {
                    int64_t *t = prev[k];
                    prev[k] = cur[k];
                    cur[k] = t;
                }
            }
        }
        /* affineTrace :57-89 */
        uint8_t k;
        *score = tmt(cur[0][m], cur[1][m], cur[2][m], &k);
        if (want_cigar) {
            route_t r;
            int64_t idx = 0;
            route_init(&r);
            for (int64_t i = n, j = m; i > 0 || j > 0;) {
                route_push(&r, &idx, k);
                uint8_t nk = tr[k][i][j];
                if (k == 0) {
                    i--;
                    j--;
                } else if (k == 1) {
                    j--;
                } else {
                    i--;
                }
                k = nk;
            }
            route_reverse(&r);
            rc = route_emit(&r, out, cap, n_out);
        } else if (n_out) {
            *n_out = 0;
        }
    }
done:
    for (int k = 0; k < 3; k++) {
        if (tr[k]) {
            for (int64_t i = 0; i <= n; i++)
                free(tr[k][i]);
            free(tr[k]);
        }
    }
    free(rows);
    return rc;
}

int orc_affine_highmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                       const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                       int free_end_gaps, int want_cigar, int64_t *score, orc_cigar *out,
                       int64_t cap, int64_t *n_out)
{
    /* Go indexes scores[alpha[i-1]][beta[j-1]] only for interior cells: no panic if n==0 or m==0 */
    if (n > 0 && m > 0 && (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim)))
        return ORC_EBASE;
    pair_ctx c = {alpha, beta, scores, dim};
    return affine_fill_trace(n, m, match_pair, &c, gap_open, gap_extend, free_end_gaps, want_cigar,
                             score, out, cap, n_out);
}

/* ---- AffineGapChunk (align/affineGap_highMem.go:227-272), ungappedRegionScore (ungapped.go:7-13) */
typedef struct {
    pair_ctx p;
    int64_t chunk;
} chunk_ctx;

static int64_t match_chunk(const void *c, int64_t i, int64_t j)
{
    const chunk_ctx *q = (const chunk_ctx *)c;
    int64_t a0 = (i - 1) * q->chunk, b0 = (j - 1) * q->chunk, ans = 0;
    for (int64_t t = 0; t < q->chunk; t++)
        ans += q->p.scores[(int64_t)q->p.alpha[a0 + t] * q->p.dim + q->p.beta[b0 + t]];
    return ans;
}

int orc_affine_chunk(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                     const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                     int64_t chunk, int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out)
{
    if (chunk <= 0 || n % chunk != 0 || m % chunk != 0) /* :229-234 log.Fatalf */
        return ORC_ECHUNK;
    if (n > 0 && m > 0 && (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim)))
        return ORC_EBASE;
    chunk_ctx c = {{alpha, beta, scores, dim}, chunk};
    int64_t n_ops = 0;
    int rc = affine_fill_trace(n / chunk, m / chunk, match_chunk, &c, gap_open, gap_extend * chunk, 0, 1,
                               score, out, cap, &n_ops);
    if (rc == ORC_OK && out) /* expandCigarRunLength :91-95 */
        for (int64_t i = 0; i < n_ops; i++)
            out[i].run_length *= chunk;
    if (n_out)
        *n_out = n_ops;
    return rc;
}

/* ---- multipleAffineGap[Chunk] (:274-353); scoreColumnMatch / ungappedRegionColumnScore
 *      (align/multiAlign.go:82-110): truncated-integer mean of pairwise scores, gaps ignored,
 *      lowercase folded to uppercase.  Go divides by zero (panics) when every pair has a gap;
 *      that is reported as ORC_EPANIC through the `bad` flag. */
typedef struct {
    const uint8_t *ga, *gb;
    int64_t na, nb, la, lb, chunk;
    const int64_t *scores;
    int dim;
    int bad;
} multi_ctx;

static int64_t column_match(multi_ctx *q, int64_t ac, int64_t bc)
{
    int64_t sum = 0, count = 0;
    for (int64_t x = 0; x < q->na; x++) {
        uint8_t a = q->ga[x * q->la + ac];
        if (a >= 5 && a <= 9)
            a -= 5;
        for (int64_t y = 0; y < q->nb; y++) {
            uint8_t b = q->gb[y * q->lb + bc];
            if (b >= 5 && b <= 9)
                b -= 5;
            if (a != 10 && b != 10) {
                if (a >= q->dim || b >= q->dim) {
                    q->bad = ORC_EBASE;
                    return 0;
                }
                sum += q->scores[(int64_t)a * q->dim + b];
                count++;
            }
        }
    }
    if (count == 0) {
        q->bad = ORC_EPANIC;
        return 0;
    }
    return sum / count; /* Go int64 division truncates toward zero, as C99 does */
}

static int64_t match_multi(const void *c, int64_t i, int64_t j)
{
    multi_ctx *q = (multi_ctx *)c;
    int64_t a0 = (i - 1) * q->chunk, b0 = (j - 1) * q->chunk, ans = 0;
    for (int64_t t = 0; t < q->chunk; t++)
        ans += column_match(q, a0 + t, b0 + t);
    return ans;
}

int orc_multi_affine_chunk(const uint8_t *ga, int64_t na_seq, int64_t n, const uint8_t *gb,
                           int64_t nb_seq, int64_t m, const int64_t *scores, int dim,
                           int64_t gap_open, int64_t gap_extend, int64_t chunk, int64_t *score,
                           orc_cigar *out, int64_t cap, int64_t *n_out)
{
    if (chunk <= 0 || n % chunk != 0 || m % chunk != 0)
        return ORC_ECHUNK;
    multi_ctx c = {ga, gb, na_seq, nb_seq, n, m, chunk, scores, dim, 0};
    int64_t n_ops = 0;
    int rc = affine_fill_trace(n / chunk, m / chunk, match_multi, &c, gap_open, gap_extend * chunk, 0, 1,
                               score, out, cap, &n_ops);
    if (c.bad)
        return c.bad;
    if (rc == ORC_OK && out)
        for (int64_t i = 0; i < n_ops; i++)
            out[i].run_length *= chunk;
    if (n_out)
        *n_out = n_ops;
    return rc;
}

/* =====================================================================================
 * ConstGap_highMem (align/constGap_highMem.go:11-67)
 * ===================================================================================== */
int orc_const_highmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                      const int64_t *scores, int dim, int64_t gap_pen, int want_cigar,
                      int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out)
{
    if (n > 0 && m > 0 && (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim)))
        return ORC_EBASE;
    const int64_t W = m + 1;
    int64_t *cur = (int64_t *)calloc((size_t)W, sizeof(int64_t));
    int64_t *prev = (int64_t *)calloc((size_t)W, sizeof(int64_t));
    uint8_t **tr = NULL;
    int rc = ORC_OK;
    if (!cur || !prev) {
        rc = ORC_ENOMEM;
        goto done;
    }
    if (want_cigar) {
        tr = (uint8_t **)calloc((size_t)(n + 1), sizeof(uint8_t *));
        if (!tr) {
            rc = ORC_ENOMEM;
            goto done;
        }
        for (int64_t i = 0; i <= n; i++) {
            tr[i] = (uint8_t *)calloc((size_t)W, 1);
            if (!tr[i]) {
                rc = ORC_ENOMEM;
                goto done;
            }
        }
    }
    for (int64_t i = 0; i <= n; i++) { /* :23-40 */
        for (int64_t j = 0; j <= m; j++) {
            if (i == 0 && j == 0) {
                cur[j] = 0;
            } else if (i == 0) {
                cur[j] = cur[j - 1] + gap_pen;
                if (tr)
                    tr[i][j] = 1;
            } else if (j == 0) {
                cur[j] = prev[j] + gap_pen;
                if (tr)
                    tr[i][j] = 2;
            } else {
                uint8_t k;
                cur[j] = tmt(prev[j - 1] + scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]],
                             cur[j - 1] + gap_pen, prev[j] + gap_pen, &k);
                if (tr)
                    tr[i][j] = k;
            }
        }
        if (i < n) {
            int64_t *t = prev;
            prev = cur;
            cur = t;
        }
    }
    *score = cur[m];
    if (want_cigar) { /* :43-65 */
        route_t r;
        int64_t idx = 0;
        route_init(&r);
        for (int64_t i = n, j = m; i > 0 || j > 0;) {
            uint8_t k = tr[i][j];
            route_push(&r, &idx, k);
            if (k == 0) {
                i--;
                j--;
            } else if (k == 1) {
                j--;
            } else {
                i--;
            }
        }
        route_reverse(&r);
        rc = route_emit(&r, out, cap, n_out);
    } else if (n_out) {
        *n_out = 0;
    }
done:
    if (tr) {
        for (int64_t i = 0; i <= n; i++)
            free(tr[i]);
        free(tr);
    }
    free(cur);
    free(prev);
    return rc;
}

/* =====================================================================================
 * Low-memory "checkerboard" drivers.  Faithful restatement, including the behaviours the
 * survey flags as probable reference defects (double-dagger note), because the reference's own
 * checker-size-3 tests (align/affineGap_test.go:57-81) exercise exactly this code.
 * Every slice index Go would bounds-check is checked here; a violation returns ORC_EPANIC.
 * ===================================================================================== */
#define LM_CHECK(cond)         \
    do {                       \
        if (!(cond)) {         \
            rc = ORC_EPANIC;   \
            goto done;         \
        }                      \
    } while (0)

static inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }

/* lastCigar (align/constGap.go:280-311) */
static void last_cigar(int64_t len_alpha, int64_t len_beta, route_t *r, int64_t *idx, uint8_t op_end)
{
    int64_t total = 0, last;
    if (op_end == 1) {
        for (int64_t q = 0; q < r->len; q++)
            if (r->v[q].op == 0 || r->v[q].op == 1)
                total += r->v[q].run_length;
        last = len_beta - total;
    } else {
        for (int64_t q = 0; q < r->len; q++)
            if (r->v[q].op == 0 || r->v[q].op == 2)
                total += r->v[q].run_length;
        last = len_alpha - total;
    }
    if (r->v[*idx].op == op_end) {
        r->v[*idx].run_length += last;
    } else {
        route_append(r, last, op_end);
        (*idx)++;
    }
}

int orc_affine_lowmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                      const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                      int64_t ci, int64_t cj, int64_t *score, orc_cigar *out, int64_t cap,
                      int64_t *n_out)
{
    if (n <= 0 || m <= 0 || ci <= 0 || cj <= 0)
        return ORC_EUNDEF; /* reference loops forever / indexes out of range on empty input */
    if (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim))
        return ORC_EBASE;
    int rc = ORC_OK;
    const int64_t W = m + 1, H = n + 1;
    const int64_t NI = n / ci + 1, NJ = m / cj + 1; /* initAffineScoring :20-41 */
    const int64_t TI = i64min(n, ci), TJ = i64min(m, cj); /* initAffineTrace :45-54 */
    const int64_t oe = gap_open + gap_extend;
    int64_t *rows = (int64_t *)calloc((size_t)(6 * W), sizeof(int64_t));
    int64_t *prep_i = (int64_t *)calloc((size_t)(3 * NI * W), sizeof(int64_t));
    int64_t *prep_j = (int64_t *)calloc((size_t)(3 * NJ * H), sizeof(int64_t));
    uint8_t *trace = (uint8_t *)calloc((size_t)(3 * TI * TJ), 1);
    int64_t *frow = (int64_t *)calloc((size_t)(6 * W), sizeof(int64_t));
    route_t r;
    route_init(&r);
    if (!rows || !prep_i || !prep_j || !trace || !frow) {
        rc = ORC_ENOMEM;
        goto done;
    }
#define PI(k, b, j) prep_i[((int64_t)(k) * NI + (b)) * W + (j)]
#define PJ(k, b, i) prep_j[((int64_t)(k) * NJ + (b)) * H + (i)]
#define TR(k, a, b) trace[((int64_t)(k) * TI + (a)) * TJ + (b)]
    uint8_t kk;
    int64_t sh;
    { /* Step 1: highestScore_affineGap (:151-207) */
        int64_t *cur[3] = {rows, rows + W, rows + 2 * W};
        int64_t *prev[3] = {rows + 3 * W, rows + 4 * W, rows + 5 * W};
        for (int64_t i = 0; i <= n; i++) {
            for (int64_t j = 0; j <= m; j++) {
                if (i == 0 && j == 0) {
                    cur[0][j] = 0;
                    cur[1][j] = gap_open;
                    cur[2][j] = gap_open;
                    for (int k = 0; k < 3; k++)
                        PJ(k, j / cj, i) = cur[k][j];
                } else if (i == 0) {
                    cur[0][j] = ORC_VERY_NEG;
                    cur[1][j] = gap_extend + cur[1][j - 1];
                    cur[2][j] = ORC_VERY_NEG;
                    if (j % cj == 0)
                        for (int k = 0; k < 3; k++)
                            PJ(k, j / cj, i) = cur[k][j];
                } else if (j == 0) {
                    cur[0][j] = ORC_VERY_NEG;
                    cur[1][j] = ORC_VERY_NEG;
                    cur[2][j] = gap_extend + prev[2][j];
                    for (int k = 0; k < 3; k++)
                        PJ(k, j / cj, i) = cur[k][j];
                } else {
                    const int64_t s = scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]];
                    int64_t v0 = tmt(s + prev[0][j - 1], s + prev[1][j - 1], s + prev[2][j - 1], &kk);
                    int64_t v1 = tmt(oe + cur[0][j - 1], gap_extend + cur[1][j - 1], oe + cur[2][j - 1], &kk);
                    int64_t v2 = tmt(oe + prev[0][j], oe + prev[1][j], gap_extend + prev[2][j], &kk);
                    cur[0][j] = v0;
                    cur[1][j] = v1;
                    cur[2][j] = v2;
                    if (j % cj == 0)
                        for (int k = 0; k < 3; k++)
                            PJ(k, j / cj, i) = cur[k][j];
                }
            }
            if (i < n) {
                if (i % ci == 0) /* :194-197 save row */
                    for (int k = 0; k < 3; k++)
                        memcpy(&PI(k, i / ci, 0), cur[k], (size_t)W * sizeof(int64_t));
                for (int k = 0; k < 3; k++) {
                    int64_t *t = prev[k];
                    prev[k] = cur[k];
                    cur[k] = t;
                }
            }
        }
        sh = tmt(cur[0][m], cur[1][m], cur[2][m], &kk);
    }
    *score = sh;
    {
        const int64_t sh_i = n, sh_j = m;
        int64_t i_min = -2, j_min = -2; /* :98-99 */
        int64_t ridx = 0;
        uint8_t k_max = 0, k_min = 0;
        int64_t guard = 0;
        for (int64_t k1 = (sh_i - 1) / ci, k2 = (sh_j - 1) / cj; k1 >= 0 && k2 >= 0;) {
            if (++guard > (NI + NJ + 4) * 4) { /* the Go loop would not terminate */
                rc = ORC_EPANIC;
                goto done;
            }
            /* ---- Step 2: fillTraceback_affineGap (:219-273) ---- */
            int64_t *cur[3] = {frow, frow + W, frow + 2 * W};
            int64_t *prev[3] = {frow + 3 * W, frow + 4 * W, frow + 5 * W};
            memset(frow, 0, (size_t)(6 * W) * sizeof(int64_t)); /* fresh make() per call :220-226 */
            LM_CHECK(k1 < NI);
            for (int k = 0; k < 3; k++)
                memcpy(prev[k], &PI(k, k1, 0), (size_t)W * sizeof(int64_t));
            int64_t i_max, j_max;
            if (i_min >= 0)
                i_max = ci * k1 + 1 + i_min;
            else
                i_max = i64min(ci * (k1 + 1), sh_i);
            if (j_min >= 0)
                j_max = cj * k2 + 1 + j_min;
            else
                j_max = i64min(cj * (k2 + 1), sh_j);
            const int64_t i_in_max = (i_max - 1) % ci, j_in_max = (j_max - 1) % cj;
            for (int64_t i = ci * k1 + 1; i <= i_max; i++) {
                const int64_t i_in = (i - 1) % ci;
                /* :252-254 -- note: the reference indexes the saved column with checkersize_j*k1 */
                const int64_t pj_idx = cj * k1 + 1 + i_in;
                LM_CHECK(k2 < NJ && pj_idx >= 0 && pj_idx < H && cj * k2 < W);
                for (int k = 0; k < 3; k++)
                    cur[k][cj * k2] = PJ(k, k2, pj_idx);
                for (int64_t j = cj * k2 + 1; j <= j_max; j++) {
                    const int64_t j_in = (j - 1) % cj;
                    LM_CHECK(i - 1 < n && j - 1 < m && j < W && i_in < TI && j_in < TJ);
                    const int64_t s = scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]];
                    uint8_t t0, t1, t2;
                    int64_t v0 = tmt(s + prev[0][j - 1], s + prev[1][j - 1], s + prev[2][j - 1], &t0);
                    int64_t v1 = tmt(oe + cur[0][j - 1], gap_extend + cur[1][j - 1], oe + cur[2][j - 1], &t1);
                    int64_t v2 = tmt(oe + prev[0][j], oe + prev[1][j], gap_extend + prev[2][j], &t2);
                    cur[0][j] = v0;
                    cur[1][j] = v1;
                    cur[2][j] = v2;
                    TR(0, i_in, j_in) = t0;
                    TR(1, i_in, j_in) = t1;
                    TR(2, i_in, j_in) = t2;
                }
                if (i <= ci * (k1 + 1) - 1 && i <= sh_i - 1) {
                    for (int k = 0; k < 3; k++) {
                        int64_t *t = prev[k];
                        prev[k] = cur[k];
                        cur[k] = t;
                    }
                }
            }
            LM_CHECK(j_max >= 0 && j_max < W);
            (void)tmt(cur[0][j_max], cur[1][j_max], cur[2][j_max], &k_max); /* :270 */

            /* ---- Step 3: writeCigar_affineGap (:287-344) ---- */
            int64_t wi_max = (i_min >= 0) ? i_min : i_in_max;
            int64_t wj_max = (j_min >= 0) ? j_min : j_in_max;
            uint8_t k_in = (i_min >= 0 && j_in_max >= 0) ? k_min : k_max; /* :305-309 */
            int64_t new_i_min = 0, new_j_min = 0; /* Go zero values if the loop never runs */
            uint8_t new_k_min = 0;
            for (int64_t a = wi_max, b = wj_max; a >= 0 && b >= 0;) {
                LM_CHECK(ridx < r.len);
                route_push(&r, &ridx, k_in);
                LM_CHECK(k_in <= 2 && a < TI && b < TJ);
                uint8_t nk = TR(k_in, a, b);
                if (k_in == 0) {
                    a--;
                    b--;
                } else if (k_in == 1) {
                    b--;
                } else {
                    a--;
                }
                k_in = nk;
                new_i_min = a;
                new_j_min = b;
                new_k_min = k_in;
            }
            i_min = new_i_min;
            j_min = new_j_min;
            k_min = new_k_min;
            if (i_min < 0 && j_min < 0) { /* :121-127 */
                k1--;
                k2--;
            } else if (i_min < 0) {
                k1--;
            } else if (j_min < 0) {
                k2--;
            }
        }
        /* Step 4 (:135-139) */
        if (i_min != -1 && j_min == -1)
            last_cigar(n, m, &r, &ridx, 2);
        else if (i_min == -1 && j_min != -1)
            last_cigar(n, m, &r, &ridx, 1);
        route_reverse(&r);
    }
done:
    free(rows);
    free(prep_i);
    free(prep_j);
    free(trace);
    free(frow);
    if (rc != ORC_OK) {
        free(r.v);
        return rc;
    }
    return route_emit(&r, out, cap, n_out);
#undef PI
#undef PJ
#undef TR
}

int orc_const_lowmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                     const int64_t *scores, int dim, int64_t gap_pen, int64_t ci, int64_t cj,
                     int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out)
{
    if (n <= 0 || m <= 0 || ci <= 0 || cj <= 0)
        return ORC_EUNDEF;
    if (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim))
        return ORC_EBASE;
    int rc = ORC_OK;
    const int64_t W = m + 1, H = n + 1;
    const int64_t NI = n / ci + 1, NJ = m / cj + 1;       /* constGap.go:132-141 */
    const int64_t TI = i64min(n, ci), TJ = i64min(m, cj); /* :90-95 */
    int64_t *rows = (int64_t *)calloc((size_t)(2 * W), sizeof(int64_t));
    int64_t *prep_i = (int64_t *)calloc((size_t)(NI * W), sizeof(int64_t));
    int64_t *prep_j = (int64_t *)calloc((size_t)(NJ * H), sizeof(int64_t));
    uint8_t *trace = (uint8_t *)calloc((size_t)(TI * TJ), 1);
    int64_t *frow = (int64_t *)calloc((size_t)(2 * W), sizeof(int64_t));
    route_t r;
    route_init(&r);
    if (!rows || !prep_i || !prep_j || !trace || !frow) {
        rc = ORC_ENOMEM;
        goto done;
    }
#define PI(b, j) prep_i[(int64_t)(b) * W + (j)]
#define PJ(b, i) prep_j[(int64_t)(b) * H + (i)]
#define TR(a, b) trace[(int64_t)(a) * TJ + (b)]
    uint8_t kk;
    { /* Step 1: highestScore (:129-176) */
        int64_t *cur = rows, *prev = rows + W;
        for (int64_t i = 0; i <= n; i++) {
            for (int64_t j = 0; j <= m; j++) {
                if (i == 0 && j == 0) {
                    cur[j] = 0;
                    PJ(j / cj, i) = cur[j];
                } else if (i == 0) {
                    cur[j] = cur[j - 1] + gap_pen;
                    if (j % cj == 0)
                        PJ(j / cj, i) = cur[j];
                } else if (j == 0) {
                    cur[j] = prev[j] + gap_pen;
                    PJ(j / cj, i) = cur[j];
                } else {
                    cur[j] = tmt(prev[j - 1] + scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]],
                                 cur[j - 1] + gap_pen, prev[j] + gap_pen, &kk);
                    if (j % cj == 0)
                        PJ(j / cj, i) = cur[j];
                }
            }
            if (i < n) {
                if (i % ci == 0)
                    memcpy(&PI(i / ci, 0), cur, (size_t)W * sizeof(int64_t));
                int64_t *t = prev;
                prev = cur;
                cur = t;
            }
        }
        *score = cur[m];
    }
    {
        const int64_t sh_i = n, sh_j = m;
        int64_t i_min = -2, j_min = -2, ridx = 0, guard = 0;
        for (int64_t k1 = (sh_i - 1) / ci, k2 = (sh_j - 1) / cj; k1 >= 0 && k2 >= 0;) {
            if (++guard > (NI + NJ + 4) * 4) {
                rc = ORC_EPANIC;
                goto done;
            }
            /* Step 2: fillTraceback (:185-222) */
            int64_t *cur = frow, *prev = frow + W;
            memset(frow, 0, (size_t)(2 * W) * sizeof(int64_t));
            LM_CHECK(k1 < NI);
            memcpy(prev, &PI(k1, 0), (size_t)W * sizeof(int64_t));
            int64_t i_max = (i_min >= 0) ? ci * k1 + 1 + i_min : i64min(ci * (k1 + 1), sh_i);
            int64_t j_max = (j_min >= 0) ? cj * k2 + 1 + j_min : i64min(cj * (k2 + 1), sh_j);
            const int64_t i_in_max = (i_max - 1) % ci, j_in_max = (j_max - 1) % cj;
            for (int64_t i = ci * k1 + 1; i <= i_max; i++) {
                const int64_t i_in = (i - 1) % ci;
                const int64_t pj_idx = cj * k1 + 1 + i_in; /* :209, same checkersize_j*k1 quirk */
                LM_CHECK(k2 < NJ && pj_idx >= 0 && pj_idx < H && cj * k2 < W);
                cur[cj * k2] = PJ(k2, pj_idx);
                for (int64_t j = cj * k2 + 1; j <= j_max; j++) {
                    const int64_t j_in = (j - 1) % cj;
                    LM_CHECK(i - 1 < n && j - 1 < m && j < W && i_in < TI && j_in < TJ);
                    uint8_t t;
                    cur[j] = tmt(prev[j - 1] + scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]],
                                 cur[j - 1] + gap_pen, prev[j] + gap_pen, &t);
                    TR(i_in, j_in) = t;
                }
                if (i <= ci * (k1 + 1) - 1 && i <= sh_i - 1) {
                    int64_t *t = prev;
                    prev = cur;
                    cur = t;
                }
            }
            /* Step 3: writeCigar (:230-275) */
            int64_t wi_max = (i_min >= 0) ? i_min : i_in_max;
            int64_t wj_max = (j_min >= 0) ? j_min : j_in_max;
            int64_t new_i_min = 0, new_j_min = 0;
            for (int64_t a = wi_max, b = wj_max; a >= 0 && b >= 0;) {
                LM_CHECK(ridx < r.len && a < TI && b < TJ);
                uint8_t k = TR(a, b);
                route_push(&r, &ridx, k);
                if (k == 0) {
                    a--;
                    b--;
                } else if (k == 1) {
                    b--;
                } else {
                    a--;
                }
                new_i_min = a;
                new_j_min = b;
            }
            i_min = new_i_min;
            j_min = new_j_min;
            if (i_min < 0 && j_min < 0) {
                k1--;
                k2--;
            } else if (i_min < 0) {
                k1--;
            } else if (j_min < 0) {
                k2--;
            }
        }
        if (i_min != -1 && j_min == -1)
            last_cigar(n, m, &r, &ridx, 2);
        else if (i_min == -1 && j_min != -1)
            last_cigar(n, m, &r, &ridx, 1);
        route_reverse(&r);
    }
done:
    free(rows);
    free(prep_i);
    free(prep_j);
    free(trace);
    free(frow);
    if (rc != ORC_OK) {
        free(r.v);
        return rc;
    }
    return route_emit(&r, out, cap, n_out);
#undef PI
#undef PJ
#undef TR
}

/* =====================================================================================
 * gsw extend step: LeftDynamicAln / RightDynamicAln, genomeGraph/search.go:234-321.
 * cigar.TripleMaxTrace (cigar/tools.go:58-66) has the same M >= I >= D tie order with ops 'M','I','D'.
 * ===================================================================================== */
static const uint8_t kOpChar[3] = {'M', 'I', 'D'};

/* the reference's route idiom here differs from align's: `len(route)==0 -> append` (search.go:250-257) */
static void ext_push(route_t *r, int64_t *idx, uint8_t op)
{
    if (r->oom)
        return;
    if (r->len == 0) {
        route_append(r, 1, op);
    } else if (r->v[*idx].op == op) {
        r->v[*idx].run_length += 1;
    } else {
        route_append(r, 1, op);
        (*idx)++;
    }
}

static int ext_fill(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m, const int64_t *scores, int dim,
                    int64_t g, int left, int64_t **m_out, uint8_t **t_out, int64_t *max_v, int64_t *max_i,
                    int64_t *max_j)
{
    const int64_t W = m + 1;
    int64_t *mm = (int64_t *)calloc((size_t)((n + 1) * W), sizeof(int64_t));
    uint8_t *tr = (uint8_t *)calloc((size_t)((n + 1) * W), 1);
    if (!mm || !tr) {
        free(mm);
        free(tr);
        return ORC_ENOMEM;
    }
    int64_t cur_max = 0, bi = 0, bj = 0;
    for (int64_t i = 0; i <= n; i++) {
        for (int64_t j = 0; j <= m; j++) {
            int64_t v;
            uint8_t k = 0;
            if (left) { /* search.go:236-251: zero boundaries, clip at 0 after recording the trace */
                if (i == 0 || j == 0) {
                    v = 0;
                } else {
                    v = tmt(mm[(i - 1) * W + j - 1] + scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]],
                            mm[i * W + j - 1] + g, mm[(i - 1) * W + j] + g, &k);
                    tr[i * W + j] = kOpChar[k];
                    if (v < 0)
                        v = 0;
                }
            } else { /* search.go:280-299 */
                if (i == 0 && j == 0) {
                    v = 0;
                } else if (i == 0) {
                    v = mm[j - 1] + g;
                    tr[j] = 'I';
                } else if (j == 0) {
                    v = mm[(i - 1) * W] + g;
                    tr[i * W] = 'D';
                } else {
                    v = tmt(mm[(i - 1) * W + j - 1] + scores[(int64_t)alpha[i - 1] * dim + beta[j - 1]],
                            mm[i * W + j - 1] + g, mm[(i - 1) * W + j] + g, &k);
                    tr[i * W + j] = kOpChar[k];
                }
                if (v > cur_max) { /* strict: the first arg-max in row-major order (:295-299) */
                    cur_max = v;
                    bi = i;
                    bj = j;
                }
            }
            mm[i * W + j] = v;
        }
    }
    *m_out = mm;
    *t_out = tr;
    *max_v = cur_max;
    *max_i = bi;
    *max_j = bj;
    return ORC_OK;
}

int orc_left_dynamic_aln(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                         const int64_t *scores, int dim, int64_t gap_pen, int64_t *score, orc_cigar *out,
                         int64_t cap, int64_t *n_out, int64_t *end_i, int64_t *end_j)
{
    if (n > 0 && m > 0 && (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim)))
        return ORC_EBASE;
    int64_t *mm;
    uint8_t *tr;
    int64_t mv, mi, mj;
    int rc = ext_fill(alpha, n, beta, m, scores, dim, gap_pen, 1, &mm, &tr, &mv, &mi, &mj);
    if (rc != ORC_OK)
        return rc;
    const int64_t W = m + 1;
    route_t r;
    route_init(&r);
    r.len = 0; /* the extend functions start from an empty route */
    int64_t idx = 0, i = n, j = m;
    while (mm[i * W + j] > 0) { /* search.go:252-272 */
        const uint8_t op = tr[i * W + j];
        ext_push(&r, &idx, op);
        if (op == 'M') {
            i--;
            j--;
        } else if (op == 'I') {
            j--;
        } else {
            i--;
        }
    }
    *score = mm[n * W + m];
    *end_i = i;
    *end_j = j;
    free(mm);
    free(tr);
    return route_emit(&r, out, cap, n_out);
}

int orc_right_dynamic_aln(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                          const int64_t *scores, int dim, int64_t gap_pen, int64_t *score, orc_cigar *out,
                          int64_t cap, int64_t *n_out, int64_t *max_i, int64_t *max_j)
{
    if (n > 0 && m > 0 && (!bases_ok(alpha, n, dim) || !bases_ok(beta, m, dim)))
        return ORC_EBASE;
    int64_t *mm;
    uint8_t *tr;
    int64_t mv, mi, mj;
    int rc = ext_fill(alpha, n, beta, m, scores, dim, gap_pen, 0, &mm, &tr, &mv, &mi, &mj);
    if (rc != ORC_OK)
        return rc;
    const int64_t W = m + 1;
    route_t r;
    route_init(&r);
    r.len = 0;
    int64_t idx = 0, i = mi, j = mj;
    while (i > 0 || j > 0) { /* search.go:301-319 */
        const uint8_t op = tr[i * W + j];
        ext_push(&r, &idx, op);
        if (op == 'M') {
            i--;
            j--;
        } else if (op == 'I') {
            j--;
        } else {
            i--;
        }
    }
    *score = mm[mi * W + mj];
    *max_i = mi;
    *max_j = mj;
    free(mm);
    free(tr);
    return route_emit(&r, out, cap, n_out);
}

/* =====================================================================================
 * Batched, threaded driver (CPU baseline shape: one worker per core over disjoint pair
 * ranges, cmd/gsw/pairedEndFastqs.go:33-35).
 * ===================================================================================== */
typedef struct {
    const uint8_t *alpha_cat, *beta_cat;
    const int64_t *alpha_off, *beta_off, *scores, *out_cigar_off;
    int64_t lo, hi, gap_open, gap_extend;
    int dim, mode, want_cigar, rc;
    int64_t *out_score, *out_cigar_n;
    orc_cigar *out_cigar;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *b = (batch_job *)arg;
    for (int64_t p = b->lo; p < b->hi; p++) {
        const uint8_t *a = b->alpha_cat + b->alpha_off[p];
        const uint8_t *q = b->beta_cat + b->beta_off[p];
        const int64_t n = b->alpha_off[p + 1] - b->alpha_off[p];
        const int64_t m = b->beta_off[p + 1] - b->beta_off[p];
        orc_cigar *oc = NULL;
        int64_t cap = 0, nops = 0;
        if (b->want_cigar) {
            oc = b->out_cigar + b->out_cigar_off[p];
            cap = b->out_cigar_off[p + 1] - b->out_cigar_off[p];
        }
        int rc;
        if (b->mode == 2)
            rc = orc_const_highmem(a, n, q, m, b->scores, b->dim, b->gap_open, b->want_cigar,
                                   &b->out_score[p], oc, cap, &nops);
        else
            rc = orc_affine_highmem(a, n, q, m, b->scores, b->dim, b->gap_open, b->gap_extend, b->mode == 1,
                                    b->want_cigar, &b->out_score[p], oc, cap, &nops);
        if (b->out_cigar_n)
            b->out_cigar_n[p] = nops;
        if (rc != ORC_OK && b->rc == ORC_OK)
            b->rc = rc;
    }
    return NULL;
}

int orc_batch(const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
              const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim,
              int64_t gap_open, int64_t gap_extend, int mode, int want_cigar, int n_threads,
              int64_t *out_score, orc_cigar *out_cigar, const int64_t *out_cigar_off,
              int64_t *out_cigar_n)
{
    if (n_threads < 1)
        n_threads = 1;
    if ((int64_t)n_threads > n_pairs)
        n_threads = n_pairs > 0 ? (int)n_pairs : 1;
    batch_job *jobs = (batch_job *)calloc((size_t)n_threads, sizeof(batch_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!jobs || !tid) {
        free(jobs);
        free(tid);
        return ORC_ENOMEM;
    }
    int rc = ORC_OK;
    for (int t = 0; t < n_threads; t++) {
        batch_job *b = &jobs[t];
        b->alpha_cat = alpha_cat;
        b->beta_cat = beta_cat;
        b->alpha_off = alpha_off;
        b->beta_off = beta_off;
        b->scores = scores;
        b->out_cigar_off = out_cigar_off;
        b->lo = n_pairs * t / n_threads;
        b->hi = n_pairs * (t + 1) / n_threads;
        b->gap_open = gap_open;
        b->gap_extend = gap_extend;
        b->dim = dim;
        b->mode = mode;
        b->want_cigar = want_cigar;
        b->rc = ORC_OK;
        b->out_score = out_score;
        b->out_cigar_n = out_cigar_n;
        b->out_cigar = out_cigar;
        if (t > 0 && pthread_create(&tid[t], NULL, batch_worker, b) != 0) {
            batch_worker(b); /* could not spawn: run inline */
            tid[t] = 0;
        }
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++)
        if (tid[t])
            pthread_join(tid[t], NULL);
    for (int t = 0; t < n_threads; t++)
        if (jobs[t].rc != ORC_OK && rc == ORC_OK)
            rc = jobs[t].rc;
    free(jobs);
    free(tid);
    return rc;
}

/*
 * gnx_synth.c -- deterministic synthetic read/reference pairs for benchmarks and parity tests
 * (SURVEY.md 8d): target = iid uniform ACGT; query = a window of the target with substitutions
 * (p=0.02) and indel events (p=0.005 per base, length Geometric(0.5) capped at 10, insert/delete
 * equiprobable), trimmed/padded to the nominal length; 10 % of the queries are unrelated iid.
 * Counter-based (splitmix64 keyed by seed and pair index), so the output does not depend on the
 * thread count.  Host-side workload generator: not part of the alignment path.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t splitmix(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t next(rng_t *r) { return splitmix(&r->s); }
static inline double unif(rng_t *r) { return (double)(next(r) >> 11) * (1.0 / 9007199254740992.0); }

typedef struct {
    uint64_t seed;
    int64_t lo, hi, n_len, m_len;
    uint8_t *alpha, *beta;
} job_t;

static void gen_pair(uint64_t seed, int64_t p, int64_t n, int64_t m, uint8_t *a, uint8_t *b)
{
    rng_t r;
    r.s = seed * 0xD1342543DE82EF95ULL + (uint64_t)p * 0x2545F4914F6CDD1DULL + 1;
    (void)next(&r);
    for (int64_t i = 0; i < n; i += 32) { /* 32 bases per 64-bit draw */
        uint64_t w = next(&r);
        for (int64_t k = 0; k < 32 && i + k < n; k++, w >>= 2)
            a[i + k] = (uint8_t)(w & 3);
    }
    if (m == 0)
        return;
    const int unrelated = unif(&r) < 0.10 || n == 0;
    if (unrelated) {
        for (int64_t i = 0; i < m; i += 32) {
            uint64_t w = next(&r);
            for (int64_t k = 0; k < 32 && i + k < m; k++, w >>= 2)
                b[i + k] = (uint8_t)(w & 3);
        }
        return;
    }
    int64_t span = n > m ? n - m : 0;
    int64_t src = span ? (int64_t)(next(&r) % (uint64_t)(span + 1)) : 0;
    int64_t out = 0;
    while (out < m && src < n) {
        uint64_t w = next(&r);
        double u = (double)(w >> 40) * (1.0 / 16777216.0); /* 24 bits decide the event */
        if (u < 0.005) {                                   /* indel event */
            int len = 1;
            uint64_t g = w;
            while (len < 10 && (g & 1)) {
                len++;
                g >>= 1;
            }
            if ((w >> 20) & 1) { /* insertion into the query */
                uint64_t x = next(&r);
                for (int k = 0; k < len && out < m; k++, x >>= 2)
                    b[out++] = (uint8_t)(x & 3);
            } else { /* deletion from the query */
                src += len;
            }
        } else if (u < 0.025) { /* substitution: one of the three other bases */
            b[out++] = (uint8_t)((a[src] + 1 + ((w >> 8) % 3)) & 3);
            src++;
        } else {
            b[out++] = a[src++];
        }
    }
    while (out < m) { /* pad with uniform bases */
        uint64_t x = next(&r);
        for (int k = 0; k < 32 && out < m; k++, x >>= 2)
            b[out++] = (uint8_t)(x & 3);
    }
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    for (int64_t p = j->lo; p < j->hi; p++)
        gen_pair(j->seed, p, j->n_len, j->m_len, j->alpha + p * j->n_len, j->beta + p * j->m_len);
    return NULL;
}

/* Fill alpha[n_pairs*n_len] and beta[n_pairs*m_len] (uniform lengths; offsets are p*len).
 * first_pair lets a rank generate only its shard of a global batch. */
int gnx_synth_pairs(uint64_t seed, int64_t first_pair, int64_t n_pairs, int64_t n_len, int64_t m_len,
                    uint8_t *alpha, uint8_t *beta, int n_threads)
{
    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    pthread_t tid[256];
    job_t jobs[256];
    for (int t = 0; t < n_threads; t++) {
        jobs[t].seed = seed;
        jobs[t].lo = first_pair + n_pairs * t / n_threads;
        jobs[t].hi = first_pair + n_pairs * (t + 1) / n_threads;
        jobs[t].n_len = n_len;
        jobs[t].m_len = m_len;
        jobs[t].alpha = alpha - first_pair * n_len;
        jobs[t].beta = beta - first_pair * m_len;
        if (pthread_create(&tid[t], NULL, worker, &jobs[t]) != 0) {
            worker(&jobs[t]);
            tid[t] = 0;
        }
    }
    for (int t = 0; t < n_threads; t++)
        if (tid[t])
            pthread_join(tid[t], NULL);
    return 0;
}

/* dnaTwoBit.NewTwoBit (dna/dnaTwoBit/dnaTwoBit.go:68-78) of n_seqs sequences of `len` bases each (bases 0..3), tightly
 * packed: the workload generator's way of handing the benchmark its inputs in the reference's packed form.  Host-side
 * test/benchmark utility, multi-threaded over sequences. */
typedef struct {
    const uint8_t *bases;
    uint64_t *words;
    int64_t lo, hi, len, wps;
} pack_job_t;

static void *pack_worker(void *arg)
{
    pack_job_t *j = (pack_job_t *)arg;
    for (int64_t s = j->lo; s < j->hi; s++) {
        const uint8_t *b = j->bases + s * j->len;
        uint64_t *w = j->words + s * j->wps;
        for (int64_t k = 0; k < j->wps; k++) {
            uint64_t v = 0;
            int64_t cnt = j->len - 32 * k < 32 ? j->len - 32 * k : 32;
            for (int64_t i = 0; i < cnt; i++)
                v = (v << 2) | (uint64_t)(b[32 * k + i] & 3);
            w[k] = v << (2 * (32 - cnt)); /* the last word is left-aligned (dnaTwoBit.go:39-41) */
        }
    }
    return 0;
}

int gnx_pack_uniform(const uint8_t *bases, int64_t n_seqs, int64_t len, uint64_t *words, int n_threads)
{
    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > 256)
        n_threads = 256;
    pthread_t th[256];
    pack_job_t jobs[256];
    const int64_t wps = (len + 31) / 32;
    for (int t = 0; t < n_threads; t++) {
        jobs[t].bases = bases;
        jobs[t].words = words;
        jobs[t].len = len;
        jobs[t].wps = wps;
        jobs[t].lo = n_seqs * t / n_threads;
        jobs[t].hi = n_seqs * (t + 1) / n_threads;
        pthread_create(&th[t], 0, pack_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], 0);
    return 0;
}

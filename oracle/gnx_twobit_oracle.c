/*
 * gnx_twobit_oracle.c -- CPU ORACLE (test infrastructure, NOT the product) for SURVEY.md row 8f-2:
 * gonomics' 2-bit DNA encoding and perfect-match seed machinery.
 *
 * Plain-C restatement of (paths relative to /root/reference, gonomics @ bd66b49b)
 *   dna/dnaTwoBit/dnaTwoBit.go:28-78     BasesToUint64LeftAln, GetBase, NewTwoBit
 *   dna/dnaTwoBit/rainbow.go:8-25        NewTwoBitRainbow (element k = NewTwoBit of k 'A's + seq)
 *   dna/dnaTwoBit/perfectAlign.go:10-85  CountRightMatches, CountLeftMatches
 *   genomeGraph/align.go:163-186         ChromAndPosToNumber, dnaToNumber, numberToChromAndPos
 *   genomeGraph/index.go:21-44           IndexGenomeIntoMap (nodes without edges)
 *   genomeGraph/search.go:425-452,567-602 extendToTheRightDev, seedMapMemPool (nodes without edges; the
 *                                        seeds in APPEND order, i.e. before SortSeedLen / heapSortSeeds)
 *
 * Parity status: CountRight/LeftMatches, NewTwoBit and GetBase are PINNED to the reference's known-answer
 * tests (dna/dnaTwoBit/perfectAlign_test.go:28-92, dnaTwoBit_test.go:9-42; tests/golden/twobit.json).
 * The seed enumeration has no asserting test in the reference: PARITY UNPINNED for orc_seeds_*.
 */
#include "gnx_oracle.h"

#include <stdlib.h>
#include <string.h>

/* BasesToUint64LeftAln (dnaTwoBit.go:28-42): `answer<<2 | uint64(seq[i])` for every base, then left-aligned.
 * The raw byte is OR-ed in, so a base > 3 (N, lowercase, gap) spills into the bits of the bases before it. */
static uint64_t bases_to_u64_left(const uint8_t *seq, int64_t start, int64_t end)
{
    uint64_t answer = 0;
    int64_t i = start;
    for (; i < end; i++) {
        answer = answer << 2;
        answer = answer | (uint64_t)seq[i];
    }
    for (; i < start + 32; i++)
        answer = answer << 2;
    return answer;
}

int64_t orc_twobit_words(int64_t len) { return (len + 31) / 32; }

/* NewTwoBitRainbow element `lead` (rainbow.go:8-25); lead = 0 is NewTwoBit (dnaTwoBit.go:68-78).
 * out has (len + lead + 31) / 32 words; TwoBit.Len = len + lead. */
int orc_new_twobit(const uint8_t *seq, int64_t len, int lead, uint64_t *out)
{
    const int64_t total = len + lead;
    uint8_t *clone = (uint8_t *)malloc((size_t)(total > 0 ? total : 1));
    if (!clone)
        return ORC_ENOMEM;
    memset(clone, 0, (size_t)lead); /* dna.A prepended `lead` times (rainbow.go:22) */
    if (len > 0)
        memcpy(clone + lead, seq, (size_t)len);
    const int64_t words = (total + 31) / 32;
    for (int64_t i = 0; i < words; i++) {
        const int64_t start = i * 32;
        const int64_t end = start + 32 < total ? start + 32 : total;
        out[i] = bases_to_u64_left(clone, start, end);
    }
    free(clone);
    return ORC_OK;
}

/* GetBase (dnaTwoBit.go:59-65) */
uint8_t orc_get_base(const uint64_t *words, uint64_t pos)
{
    const uint64_t idx = pos / 32, rem = pos % 32;
    const uint64_t shift = 64 - 2 * (rem + 1);
    return (uint8_t)((words[idx] >> shift) & 3);
}

static int lz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static int tz64(uint64_t x) { return x ? __builtin_ctzll(x) : 64; }
static int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

/* CountRightMatches (perfectAlign.go:10-47).  Returns the count, or
 *   -1  log.Fatalf "Different offsets"                       (:24-26)
 *   -2  Go runtime panic: index out of range on one.Seq[i] / two.Seq[j] (start beyond the last word, or < 0) */
int64_t orc_count_right(const uint64_t *one, int64_t one_len, const uint64_t *two, int64_t two_len,
                        int64_t start_one, int64_t start_two)
{
    if (start_one < 0 || start_two < 0)
        return -2;
    const int offset_one = (int)(start_one % 32) * 2, offset_two = (int)(start_two % 32) * 2;
    if (offset_one != offset_two)
        return -1;
    int64_t i = start_one / 32, j = start_two / 32;
    const int64_t i_end = (one_len + 31) / 32, j_end = (two_len + 31) / 32;
    if (i >= i_end || j >= j_end)
        return -2;
    uint64_t diff = one[i] ^ two[j];
    diff &= ~(uint64_t)0 >> offset_one;
    int bit_matches = lz64(diff);
    int64_t total = bit_matches - offset_one;
    for (i = i + 1, j = j + 1; i < i_end && j < j_end && bit_matches == 64; i++, j++) {
        diff = one[i] ^ two[j];
        bit_matches = lz64(diff);
        total += bit_matches;
    }
    return min64(min64(total / 2, one_len - start_one), two_len - start_two);
}

/* CountLeftMatches (perfectAlign.go:49-85); same error returns. */
int64_t orc_count_left(const uint64_t *one, int64_t one_len, const uint64_t *two, int64_t two_len,
                       int64_t start_one, int64_t start_two)
{
    if (start_one < 0 || start_two < 0)
        return -2;
    const int offset_one = (int)(start_one % 32) * 2, offset_two = (int)(start_two % 32) * 2;
    if (offset_one != offset_two)
        return -1;
    const int first_bits_no_look = 64 - offset_one - 2;
    int64_t i = start_one / 32, j = start_two / 32;
    if (i >= (one_len + 31) / 32 || j >= (two_len + 31) / 32)
        return -2;
    uint64_t diff = one[i] ^ two[j];
    diff &= ~(uint64_t)0 << first_bits_no_look;
    int bit_matches = tz64(diff);
    int64_t total = bit_matches - first_bits_no_look;
    for (i = i - 1, j = j - 1; i >= 0 && j >= 0 && bit_matches == 64; i--, j--) {
        diff = one[i] ^ two[j];
        bit_matches = tz64(diff);
        total += bit_matches;
    }
    return total / 2;
}

/* ------------------------------------------------------------------------------------------------
 * Seed enumeration over a genome of edge-less nodes (a linear reference: one node per chromosome).
 * ------------------------------------------------------------------------------------------------ */

/* dnaToNumber (genomeGraph/align.go:170-177) */
static uint64_t dna_to_number(const uint8_t *seq, int64_t start, int64_t end)
{
    uint64_t answer = (uint64_t)seq[start];
    for (int64_t i = start + 1; i < end; i++) {
        answer = answer << 2;
        answer = answer | (uint64_t)seq[i];
    }
    return answer;
}

typedef struct {
    uint64_t key, loc;
} orc_kmer;

static int kmer_cmp(const void *a, const void *b)
{
    const orc_kmer *x = (const orc_kmer *)a, *y = (const orc_kmer *)b;
    if (x->key != y->key)
        return x->key < y->key ? -1 : 1;
    return x->loc < y->loc ? -1 : (x->loc > y->loc ? 1 : 0); /* insertion order = (node, pos) ascending */
}

/* IndexGenomeIntoMap (genomeGraph/index.go:21-44) for nodes without edges: every pos = 0, step, 2*step ...
 * <= len - seedLen whose window holds no dna.N (4) is indexed under dnaToNumber(window) with the value
 * ChromAndPosToNumber(node, pos); a Go map of slices == entries grouped by key in insertion order, which
 * is what a (key, loc)-sorted array gives.  Returns the number of entries (<= cap) or -1 on overflow. */
int64_t orc_seed_index(const uint8_t *genome_cat, const int64_t *node_off, int64_t n_nodes, int seed_len,
                       int seed_step, uint64_t *out_key, uint64_t *out_loc, int64_t cap)
{
    int64_t k = 0;
    for (int64_t node = 0; node < n_nodes; node++) {
        const uint8_t *seq = genome_cat + node_off[node];
        const int64_t len = node_off[node + 1] - node_off[node];
        for (int64_t pos = 0; pos < len - seed_len + 1; pos += seed_step) {
            int has_n = 0;
            for (int64_t q = pos; q < pos + seed_len; q++)
                has_n |= seq[q] == 4;
            if (has_n)
                continue;
            if (k >= cap)
                return -1;
            out_key[k] = dna_to_number(seq, pos, pos + seed_len);
            out_loc[k] = ((uint64_t)node << 32) | (uint64_t)pos;
            k++;
        }
    }
    orc_kmer *tmp = (orc_kmer *)malloc(sizeof(orc_kmer) * (size_t)(k > 0 ? k : 1));
    if (!tmp)
        return -1;
    for (int64_t i = 0; i < k; i++) {
        tmp[i].key = out_key[i];
        tmp[i].loc = out_loc[i];
    }
    qsort(tmp, (size_t)k, sizeof(orc_kmer), kmer_cmp);
    for (int64_t i = 0; i < k; i++) {
        out_key[i] = tmp[i].key;
        out_loc[i] = tmp[i].loc;
    }
    free(tmp);
    return k;
}

/* seedMapMemPool (genomeGraph/search.go:567-602) for one read against edge-less nodes, seeds in APPEND
 * order (fwd hits then rev hits of readStart 0, 1, ...; the caller sorts them afterwards).  With no edges
 * extendToTheRightDev (:425-452) yields zero or one seed per hit and extendToTheLeftDev (:454-476) is the
 * identity.  read_rc is the reverse complement (fastq.FastqBig.SeqRc).  node_words/node_word_off/node_off:
 * NewTwoBit of every node.  Seeds are written as 6 x uint32: TargetId, TargetStart, QueryStart, Length,
 * PosStrand, TotalLength.  Returns the seed count or -1 if cap is too small, -2 on a reference panic /
 * Fatalf inside CountLeft/RightMatches. */
int64_t orc_seeds_for_read(const uint64_t *idx_key, const uint64_t *idx_loc, int64_t n_idx,
                           const uint64_t *node_words, const int64_t *node_word_off, const int64_t *node_off,
                           const uint8_t *read, const uint8_t *read_rc, int64_t read_len, int seed_len,
                           uint32_t *out, int64_t cap)
{
    /* the 32-element rainbows of both strands (fastq.FastqBig.Rainbow / RainbowRc) */
    const int64_t wmax = (read_len + 31 + 31) / 32 + 1;
    uint64_t *rb = (uint64_t *)calloc((size_t)(2 * 32 * wmax), sizeof(uint64_t));
    if (!rb)
        return -1;
    for (int strand = 0; strand < 2; strand++)
        for (int k = 0; k < 32; k++)
            orc_new_twobit(strand == 0 ? read : read_rc, read_len, k, rb + ((size_t)strand * 32 + k) * wmax);
    const unsigned key_shift = 64 - (unsigned)seed_len * 2;
    int64_t n_out = 0;
    for (int64_t read_start = 0; read_start < read_len - seed_len + 1; read_start++) {
        const int64_t key_idx = (read_start + 31) / 32;
        const int key_offset = 31 - (int)((read_start + 31) % 32);
        for (int strand = 0; strand < 2; strand++) {
            const uint64_t *rainbow = rb + (size_t)strand * 32 * wmax;
            const uint64_t seq_key = rainbow[(size_t)key_offset * wmax + key_idx] >> key_shift;
            /* seedHash[seqKey]: the run of equal keys in the sorted index */
            int64_t lo = 0, hi = n_idx;
            while (lo < hi) {
                const int64_t mid = (lo + hi) / 2;
                if (idx_key[mid] < seq_key)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            for (int64_t h = lo; h < n_idx && idx_key[h] == seq_key; h++) {
                const int64_t node = (int64_t)(idx_loc[h] >> 32), node_pos = (int64_t)(idx_loc[h] & 0xffffffffu);
                const uint64_t *nw = node_words + node_word_off[node];
                const int64_t node_len = node_off[node + 1] - node_off[node];
                const int node_offset = (int)(node_pos % 32);
                int read_offset = 31 - (int)((read_start - node_offset + 31) % 32);
                int64_t left = orc_count_left(nw, node_len, rainbow + (size_t)read_offset * wmax, read_len + read_offset,
                                              node_pos, read_start + read_offset);
                if (left < 0) {
                    free(rb);
                    return -2;
                }
                left = min64(read_start + 1, left);
                /* extendToTheRightDev(node, read, readStart-(left-1), nodePos-(left-1), strand) */
                const int64_t r_start = read_start - (left - 1), n_start = node_pos - (left - 1);
                const int n_off2 = (int)(n_start % 32);
                read_offset = 31 - (int)((r_start - n_off2 + 31) % 32);
                const int64_t right = orc_count_right(nw, node_len, rainbow + (size_t)read_offset * wmax,
                                                      read_len + read_offset, n_start, r_start + read_offset);
                if (right < 0) {
                    free(rb);
                    return -2;
                }
                if (right == 0)
                    continue; /* "nothing aligned here" -> nil */
                if (n_out >= cap) {
                    free(rb);
                    return -1;
                }
                uint32_t *s = out + 6 * n_out++;
                s[0] = (uint32_t)node;
                s[1] = (uint32_t)n_start;
                s[2] = (uint32_t)r_start;
                s[3] = (uint32_t)right;
                s[4] = strand == 0;
                s[5] = (uint32_t)right;
            }
        }
    }
    free(rb);
    return n_out;
}

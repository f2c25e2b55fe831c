/*
 * gnx_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the gonomics `align` package's dynamic-programming
 * hot path (reference: the .go files under /root/reference/align, gonomics @ bd66b49b).  It exists only to
 * check the CUDA path bit-for-bit and to serve as the timed CPU baseline in
 * bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (libgnxalign.so) never
 * links or calls anything in this directory.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
 * against the reference's own golden vectors: align/affineGap_test.go, align/view_test.go,
 * align/multiAlign_test.go, cmd/globalAlignmentAnchor/testdata/out_alignment.{1,2}.expected.tsv,
 * cmd/cigarToBed/testdata/{sethvsraven,firstTest}, cmd/globalAlignment/testdata.
 * The Go toolchain is absent from this image, so the reference itself cannot be
 * executed here (no oracle/_ref); the golden vectors are the pin.
 */
#ifndef GNX_ORACLE_H
#define GNX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* align.Cigar (align/align.go:21-24): {RunLength int64; Op ColType(uint8)} -> 16 bytes on amd64 */
typedef struct {
    int64_t run_length;
    uint8_t op; /* 0 = ColM, 1 = ColI, 2 = ColD  (align/align.go:14-18) */
} orc_cigar;

enum {
    ORC_OK = 0,
    ORC_EBASE = 1,  /* Go would panic: scores[a][b] index out of range (base >= dim)     */
    ORC_ECAP = 2,   /* caller's cigar buffer too small                                    */
    ORC_ECHUNK = 3, /* Go would log.Fatalf: length not a multiple of chunkSize            */
    ORC_EPANIC = 4, /* Go would panic with an index out of range inside the low-mem driver */
    ORC_EUNDEF = 5, /* reference behaviour undefined (empty input to a low-mem driver)    */
    ORC_ENOMEM = 6
};

/* align.veryNegNum (align/align.go:8) */
#define ORC_VERY_NEG (INT64_MIN / 2)

/* affineGap_highMem + affineTrace (align/affineGap_highMem.go:181-223, :57-89).
 * free_end_gaps = 0 -> AffineGap_highMem (:99), 1 -> AffineGapLocal(target=alpha, query=beta) (:105).
 * scores: dim*dim row-major [alpha][beta].  want_cigar=0 skips the trace matrix (score only). */
int orc_affine_highmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                       const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                       int free_end_gaps, int want_cigar, int64_t *score, orc_cigar *out,
                       int64_t cap, int64_t *n_out);

/* ConstGap_highMem (align/constGap_highMem.go:11-67) */
int orc_const_highmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                      const int64_t *scores, int dim, int64_t gap_pen, int want_cigar,
                      int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out);

/* AffineGap_customizeCheckersize (align/affineGap.go:73-144) incl. its multi-board quirks;
 * AffineGap (align/affineGap.go:59-68) is ci = cj = 10000. */
int orc_affine_lowmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                      const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                      int64_t ci, int64_t cj, int64_t *score, orc_cigar *out, int64_t cap,
                      int64_t *n_out);

/* ConstGap_customizeCheckersize (align/constGap.go:73-124); ConstGap (:13-68) is 10000x10000. */
int orc_const_lowmem(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                     const int64_t *scores, int dim, int64_t gap_pen, int64_t ci, int64_t cj,
                     int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out);

/* AffineGapChunk (align/affineGap_highMem.go:227-272) */
int orc_affine_chunk(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                     const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                     int64_t chunk, int64_t *score, orc_cigar *out, int64_t cap, int64_t *n_out);

/* multipleAffineGap / multipleAffineGapChunk (align/affineGap_highMem.go:274-353) over two
 * column-major-free "groups": group g is n_seq rows of equal length `len`, stored row-major
 * (seq s, column c) at base[s*len + c]; bases may be lowercase (5..9) or Gap (10)
 * (scoreColumnMatch, align/multiAlign.go:82-102).  chunk = 1 reproduces multipleAffineGap. */
int orc_multi_affine_chunk(const uint8_t *ga, int64_t na_seq, int64_t n, const uint8_t *gb,
                           int64_t nb_seq, int64_t m, const int64_t *scores, int dim,
                           int64_t gap_open, int64_t gap_extend, int64_t chunk, int64_t *score,
                           orc_cigar *out, int64_t cap, int64_t *n_out);

/* ---- "next" row 8f-1: the gsw extend step (genomeGraph/search.go:234-321), linear gap, cigar.Cigar ops
 * 'M','I','D' (cigar/cigar.go:15-18, tie-break cigar/tools.go:58-66), route in TRACEBACK order (the
 * reference does not reverse it here).  Restated for a clean dynamicScoreKeeper (empty route, currMax 0):
 * resetDynamicScore takes its argument by value (search.go:104-107), so in the reference a non-empty
 * route leaks in from the caller; that caller-side quirk is outside these functions.
 * PARITY UNPINNED: the reference has no asserting test for these two functions. */
int orc_left_dynamic_aln(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                         const int64_t *scores, int dim, int64_t gap_pen, int64_t *score, orc_cigar *out,
                         int64_t cap, int64_t *n_out, int64_t *end_i, int64_t *end_j);
int orc_right_dynamic_aln(const uint8_t *alpha, int64_t n, const uint8_t *beta, int64_t m,
                          const int64_t *scores, int dim, int64_t gap_pen, int64_t *score, orc_cigar *out,
                          int64_t cap, int64_t *n_out, int64_t *max_i, int64_t *max_j);

/* ---- "next" row 8f-2: dna/dnaTwoBit + the perfect-match seed step (gnx_twobit_oracle.c) ------------
 * NewTwoBit / NewTwoBitRainbow element `lead` (dna/dnaTwoBit/dnaTwoBit.go:68-78, rainbow.go:8-25);
 * GetBase (:59-65); CountRightMatches / CountLeftMatches (perfectAlign.go:10-85; -1 = log.Fatalf
 * "Different offsets", -2 = Go index-out-of-range panic).  PINNED: perfectAlign_test.go:28-92,
 * dnaTwoBit_test.go:9-42 (tests/golden/twobit.json). */
int64_t orc_twobit_words(int64_t len);
int orc_new_twobit(const uint8_t *seq, int64_t len, int lead, uint64_t *out);
uint8_t orc_get_base(const uint64_t *words, uint64_t pos);
int64_t orc_count_right(const uint64_t *one, int64_t one_len, const uint64_t *two, int64_t two_len,
                        int64_t start_one, int64_t start_two);
int64_t orc_count_left(const uint64_t *one, int64_t one_len, const uint64_t *two, int64_t two_len,
                       int64_t start_one, int64_t start_two);
/* IndexGenomeIntoMap (genomeGraph/index.go:21-44) and seedMapMemPool (genomeGraph/search.go:567-602) for
 * edge-less nodes, seeds in append order (before the reference's unstable sort).  PARITY UNPINNED. */
int64_t orc_seed_index(const uint8_t *genome_cat, const int64_t *node_off, int64_t n_nodes, int seed_len,
                       int seed_step, uint64_t *out_key, uint64_t *out_loc, int64_t cap);
int64_t orc_seeds_for_read(const uint64_t *idx_key, const uint64_t *idx_loc, int64_t n_idx,
                           const uint64_t *node_words, const int64_t *node_word_off, const int64_t *node_off,
                           const uint8_t *read, const uint8_t *read_rc, int64_t read_len, int seed_len,
                           uint32_t *out, int64_t cap);

/* Batched driver used as the CPU baseline: one affineGap_highMem (or ConstGap_highMem when
 * mode==2) per pair, pairs split into contiguous ranges over n_threads pthreads -- the
 * goroutine-per-worker shape of cmd/gsw/pairedEndFastqs.go:33-35.  Cigars are written to
 * out_cigar at out_cigar_off[p] with a per-pair capacity of (n_p + m_p + 1) implied by the
 * caller-provided offsets (out_cigar_off has n_pairs+1 entries, filled by the CALLER);
 * out_cigar_n[p] receives the op count.  mode: 0 global affine, 1 free-end affine, 2 const gap
 * (gap_open is the penalty). */
int orc_batch(const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
              const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim,
              int64_t gap_open, int64_t gap_extend, int mode, int want_cigar, int n_threads,
              int64_t *out_score, orc_cigar *out_cigar, const int64_t *out_cigar_off,
              int64_t *out_cigar_n);

#ifdef __cplusplus
}
#endif
#endif

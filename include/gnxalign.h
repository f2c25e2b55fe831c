/*
 * gnxalign.h -- C ABI of libgnxalign.so: B200 (sm_100a) implementation of the gonomics
 * `align` package's pairwise DP hot path (score-matrix fill, traceback, Cigar).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  It is what a
 * cgo shim in the reference's `align` package binds (see INTEGRATION.md for the Go side) and
 * what tests load through ctypes.  Every entry point cites the Go function it replaces
 * (paths relative to the gonomics tree @ bd66b49b).
 *
 * Conventions
 *   - Sequences are `[]dna.Base` byte arrays (dna/dna.go:5-21: A,C,G,T,N = 0..4); a batch is the
 *     concatenation of all alphas (resp. betas) plus an (n_pairs+1)-entry int64 offset array.
 *   - `scores` is the reference's `[][]int64` matrix flattened row-major [alpha][beta], dim x dim
 *     (align/align.go:28-64; dim = 5 for the stock matrices).
 *   - gnx_cigar is layout-identical to Go's align.Cigar{RunLength int64; Op ColType} (align/align.go:21-24),
 *     16 bytes, so results can be written straight into a Go-allocated []align.Cigar.
 *   - Inputs are borrowed and never modified or retained; outputs are caller-owned.
 *   - All functions return 0 (GNX_OK) or a GNX_E* code; gnx_last_error() gives the text.  The
 *     reference never returns an error on this path -- it panics (base >= dim indexes past the
 *     matrix) or log.Fatalf's (chunk misuse); the Go shim maps codes back to those behaviours.
 *   - A context is bound to one CUDA device and is NOT thread-safe: use one context per calling
 *     goroutine/thread (contexts are cheap after the first), exactly like the reference's
 *     per-worker scratch matrices in cmd/gsw.
 */
#ifndef GNXALIGN_H
#define GNXALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnx_ctx gnx_ctx;

/* align.Cigar (align/align.go:21-24); op: 0 = ColM, 1 = ColI, 2 = ColD (align/align.go:12-18) */
typedef struct {
    int64_t run_length;
    uint8_t op;
} gnx_cigar;

enum {
    GNX_OK = 0,
    GNX_EBASE = 1,  /* a base >= dim in a pair with n>0 and m>0: Go panics (index out of range)      */
    GNX_ECAP = 2,   /* cigar_cap too small; out_cigar_off is still filled, see gnx_copy_last_cigars  */
    GNX_ECHUNK = 3, /* AffineGapChunk: a length is not a multiple of chunk (Go: log.Fatalf)          */
    GNX_EEMPTY = 4, /* low-mem entry point called with an empty sequence (reference undefined)       */
    GNX_ECUDA = 5,  /* CUDA runtime failure, text in gnx_last_error                                  */
    GNX_EARG = 6,   /* bad argument (NULL pointer, dim out of range, ...)                            */
    GNX_ERANGE = 7, /* scores/penalties/lengths exceed the exact-arithmetic range of every kernel    */
    GNX_EDIVZERO = 8, /* profile DP: a column pair with no ungapped base pair (Go: integer divide by zero
                         in scoreColumnMatch, align/multiAlign.go:101)                                */
    GNX_EOFFSET = 9,  /* CountRight/LeftMatches: start offsets differ modulo 32 (Go: log.Fatalf "Different
                         offsets", dna/dnaTwoBit/perfectAlign.go:24-26,63-65)                         */
    GNX_EINDEX = 10   /* 2-bit sequences: position beyond the last word (Go: index out of range panic) */
};

/* mode for gnx_affine_batch */
enum {
    GNX_GLOBAL = 0,  /* AffineGap_highMem / AffineGap (align/affineGap_highMem.go:99, affineGap.go:59) */
    GNX_FREE_END = 1 /* AffineGapLocal(target=alpha, query=beta) (align/affineGap_highMem.go:105)      */
};

/* ---- lifetime ------------------------------------------------------------------------------ */
int gnx_device_count(void);
/* workspace_bytes: device memory the context may use for traceback matrices of the chunks in
 * flight (0 = default, 2/3 of the device's free memory at creation, capped at 128 GiB: long pairs need
 * ~0.8 byte per DP cell and enough pairs in flight to fill the SMs; buffers are grown on demand, so
 * read-sized batches only ever use a few GB of it). */
gnx_ctx *gnx_create(int device, size_t workspace_bytes);
void gnx_destroy(gnx_ctx *ctx);
const char *gnx_last_error(gnx_ctx *ctx); /* ctx may be NULL: error of the last failed gnx_create */
const char *gnx_version(void);

/* Page-locked host buffers: inputs/outputs placed here are DMA'd without a staging copy. */
void *gnx_host_alloc(size_t bytes);
void gnx_host_free(void *p);

/* ---- affine gap ---------------------------------------------------------------------------- *
 * Replaces, per pair p (alpha_p, beta_p):
 *   mode GNX_GLOBAL  : align.AffineGap_highMem(alpha, beta, scores, gapOpen, gapExtend)
 *                      (align/affineGap_highMem.go:99-101,181-223 + affineTrace :57-89), which for
 *                      1 <= len <= 10000 is also align.AffineGap / AffineGap_customizeCheckersize
 *                      (align/affineGap.go:59-144; single checkerboard).
 *   mode GNX_FREE_END: align.AffineGapLocal(target, query, ...) (align/affineGap_highMem.go:105-107)
 *                      and each element of the GoAffineGapLocalEngine stream (:120-179).
 * want_cigar = 0: scores only (out_cigar/out_cigar_off may be NULL).
 * out_cigar_off has n_pairs+1 entries; pair p's cigar is out_cigar[off[p] .. off[p+1]).  If the
 * total exceeds cigar_cap the call returns GNX_ECAP with scores and offsets filled; fetch the
 * cigars with gnx_copy_last_cigars after growing the buffer. */
int gnx_affine_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off,
                     const uint8_t *beta_cat, const int64_t *beta_off, int64_t n_pairs,
                     const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend, int mode,
                     int want_cigar, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off,
                     int64_t cigar_cap);

/* ---- affine gap on dnaTwoBit inputs ---------------------------------------------------------- *
 * Same computation as gnx_affine_batch with the sequences in the reference's own packed form,
 * dnaTwoBit.TwoBit{Seq []uint64; Len int} (dna/dnaTwoBit/dnaTwoBit.go:14-17,28-42: 32 bases per word,
 * first base in bits 63:62, the last word left-aligned) -- what cmd/gsw keeps its genome and reads in.
 * alpha_words is the concatenation of the TwoBit.Seq slices of all alphas, TIGHTLY packed: sequence p
 * occupies ceil(Len_p / 32) words right after sequence p-1; alpha_len[p] = its TwoBit.Len.  For a batch
 * whose alphas all have the same length pass alpha_len = NULL and that length in alpha_uniform_len (then
 * beta_len must be NULL too and the betas uniform): no per-pair metadata crosses PCIe at all.
 * A quarter of the bytes of gnx_affine_batch travel to the device; there the packed 16-bit kernels consume
 * the words directly (staged into shared memory by 1-D TMA, cp.async.bulk + mbarrier) and every other
 * kernel reads a device-side expansion.  Two bits cannot hold dna.N: callers with N / lowercase bases use
 * the byte entry point (NewTwoBit itself mis-packs them, dnaTwoBit.go:33-37).  Results and error codes
 * are those of gnx_affine_batch on the unpacked sequences. */
int gnx_affine_batch_twobit(gnx_ctx *ctx, const uint64_t *alpha_words, const int64_t *alpha_len,
                            int64_t alpha_uniform_len, const uint64_t *beta_words, const int64_t *beta_len,
                            int64_t beta_uniform_len, int64_t n_pairs, const int64_t *scores, int dim,
                            int64_t gap_open, int64_t gap_extend, int mode, int want_cigar, int64_t *out_score,
                            gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap);

/* ---- constant gap ---------------------------------------------------------------------------- *
 * Replaces align.ConstGap_highMem (align/constGap_highMem.go:11-67) and, for 1 <= len <= 10000,
 * align.ConstGap / ConstGap_customizeCheckersize (align/constGap.go:13-124). */
int gnx_const_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off,
                    const uint8_t *beta_cat, const int64_t *beta_off, int64_t n_pairs,
                    const int64_t *scores, int dim, int64_t gap_pen, int want_cigar,
                    int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap);

/* ---- chunked affine gap ---------------------------------------------------------------------- *
 * Replaces align.AffineGapChunk (align/affineGap_highMem.go:227-272): DP over chunk-sized blocks,
 * match = ungappedRegionScore (align/ungapped.go:7-13), gap step = gapExtend*chunk, run lengths
 * multiplied by chunk (expandCigarRunLength :91-95). */
int gnx_affine_chunk_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off,
                           const uint8_t *beta_cat, const int64_t *beta_off, int64_t n_pairs,
                           const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                           int64_t chunk, int64_t *out_score, gnx_cigar *out_cigar,
                           int64_t *out_cigar_off, int64_t cigar_cap);

/* ---- profile (group-vs-group) affine gap: the progressive-MSA inner loop ----------------------- *
 * Replaces multipleAffineGap / multipleAffineGapChunk (align/affineGap_highMem.go:274-353) with their
 * match score scoreColumnMatch / ungappedRegionColumnScore (align/multiAlign.go:82-110: truncated
 * integer mean of scores[a][b] over the ungapped base pairs of two alignment columns, lowercase folded),
 * batched over the (x, y) group pairs that nearestGroups / nearestGroupsChunk (multiAlign.go:27-57)
 * evaluate; chunk = 1 is multipleAffineGap.
 * Group g is a sub-alignment of group_nseq[g] (>= 1) sequences of equal length
 * L_g = (group_off[g+1] - group_off[g]) / group_nseq[g], stored row after row in group_cat (the
 * fasta.Fasta.Seq bytes: dna.Base incl. lowercase 5..9 and dna.Gap = 10).  DP p aligns
 * alpha = group pair_x[p] against beta = group pair_y[p]; run lengths are in bases (already x chunk).
 * Errors in the reference's order: GNX_ECHUNK (log.Fatalf, :310-315), then the first panic in pair /
 * row-major cell order: GNX_EBASE (a base >= dim opposite an ungapped base) or GNX_EDIVZERO. */
int gnx_multi_affine_chunk_batch(gnx_ctx *ctx, const uint8_t *group_cat, const int64_t *group_off,
                                 const int64_t *group_nseq, int64_t n_groups, const int64_t *pair_x,
                                 const int64_t *pair_y, int64_t n_pairs, const int64_t *scores, int dim,
                                 int64_t gap_open, int64_t gap_extend, int64_t chunk, int want_cigar,
                                 int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off,
                                 int64_t cigar_cap);

/* ---- gsw extend step ------------------------------------------------------------ *
 * Replaces genomeGraph.LeftDynamicAln (genomeGraph/search.go:234-274) and RightDynamicAln (:276-321),
 * the linear-gap DPs cmd/gsw's seed-and-extend runs on either side of a seed, called with a clean
 * dynamicScoreKeeper (empty route, currMax 0 -- what the callers pass, since resetDynamicScore takes its
 * argument by value, :104-107).  Tie-break cigar.TripleMaxTrace (cigar/tools.go:58-66), M >= I >= D.
 *   GNX_EXT_LEFT : zero boundaries, cells clipped at 0 after their trace is recorded; score = m(n,m);
 *                  the route is walked from (n,m) while the cell value is > 0; out_end_i/j = where it stopped.
 *   GNX_EXT_RIGHT: Needleman-Wunsch boundaries; score = the first strict maximum in row-major order
 *                  (0 at (0,0) if nothing is positive); route from there to (0,0); out_end_i/j = that cell.
 *                  gap_pen must be <= 0.
 * out_cigar holds cigar.Cigar{RunLength int; Op byte} records (same 16-byte layout as gnx_cigar) with
 * Op = 'M', 'I' or 'D' (cigar/cigar.go:15-18), in TRACEBACK order -- the reference does not reverse the
 * route inside these functions.  With want_cigar = 0 only scores (and, for the right side, the end cell;
 * -1 for the left side) are produced.  dim <= 5. */
/*   GNX_EXT_LEFT_LOCAL / GNX_EXT_RIGHT_LOCAL: genomeGraph.LeftLocal / RightLocal (genomeGraph/localAlignment.go:95-196),
 *                  the older forms of the same two DPs: cigar.TripleMaxTraceExtended (cigar/tools.go:69-81) writes a
 *                  diagonal step as '=' when the substitution score is positive and 'X' otherwise, and the route is
 *                  reversed into alignment order (cigar.ReverseCigar).  Scores and end cells are those of the
 *                  LEFT / RIGHT forms: LeftLocal returns (score, route, minI = out_end_i, len(alpha), minJ =
 *                  out_end_j, len(beta)), RightLocal (score, route, 0, maxI = out_end_i, 0, maxJ = out_end_j). */
enum { GNX_EXT_LEFT = 1, GNX_EXT_RIGHT = 2, GNX_EXT_LEFT_LOCAL = 3, GNX_EXT_RIGHT_LOCAL = 4 };
int gnx_extend_batch(gnx_ctx *ctx, int side, const uint8_t *alpha_cat, const int64_t *alpha_off,
                     const uint8_t *beta_cat, const int64_t *beta_off, int64_t n_pairs, const int64_t *scores,
                     int dim, int64_t gap_pen, int want_cigar, int64_t *out_score, int64_t *out_end_i,
                     int64_t *out_end_j, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap);

/* ---- dna/dnaTwoBit on the device ------------------------------------------------ *
 * A gnx_twobit is a SET of dnaTwoBit.TwoBit sequences resident in the context's device memory (a genome's
 * nodes, a batch of reads): upload and pack once, query many times.
 *
 * gnx_twobit_new      dnaTwoBit.NewTwoBit (dna/dnaTwoBit/dnaTwoBit.go:68-78) of every sequence
 *                     seq_cat[seq_off[s] .. seq_off[s+1]); lead = k (0..31) prepends k x dna.A first, i.e.
 *                     element k of NewTwoBitRainbow (rainbow.go:8-25) with TwoBit.Len = len + k.  Bit-exact
 *                     including bases > 3: BasesToUint64LeftAln ORs the raw dna.Base byte (:33-37), so N,
 *                     lowercase and gap codes spill into the bits of the bases before them.
 * gnx_twobit_download TwoBit.Seq of every sequence (concatenated; out_word_off has n_seqs+1 entries) and
 *                     TwoBit.Len.  Any output pointer may be NULL.
 * gnx_twobit_unpack   dnaTwoBit.GetBase (:59-65) for every position of every sequence, concatenated.
 * gnx_twobit_get_bases GetBase(set[q_seq[q]], q_pos[q]) for a list of queries; GNX_EINDEX where Go panics.
 * gnx_twobit_count_matches  dnaTwoBit.CountRightMatches (perfectAlign.go:10-47; dir GNX_MATCH_RIGHT) or
 *                     CountLeftMatches (:49-85; GNX_MATCH_LEFT) of one[q_one[q]] from q_start_one[q] against
 *                     two[q_two[q]] from q_start_two[q].  GNX_EOFFSET / GNX_EINDEX report the first query (in
 *                     order) on which the reference would log.Fatalf / panic.
 * gnx_twobit_pack_device  NewTwoBit of ONE sequence already in device memory, enqueued on cuda_stream
 *                     without synchronising (d_words: (n_bases + lead + 31) / 32 words). */
typedef struct gnx_twobit gnx_twobit;
enum { GNX_MATCH_RIGHT = 0, GNX_MATCH_LEFT = 1 };
int gnx_twobit_new(gnx_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int64_t n_seqs, int lead,
                   gnx_twobit **out);
void gnx_twobit_free(gnx_twobit *tb);
int gnx_twobit_info(const gnx_twobit *tb, int64_t *n_seqs, int64_t *total_words);
int gnx_twobit_download(gnx_ctx *ctx, const gnx_twobit *tb, uint64_t *out_words, int64_t *out_word_off,
                        int64_t *out_len);
int gnx_twobit_unpack(gnx_ctx *ctx, const gnx_twobit *tb, uint8_t *out_cat, int64_t out_cap);
int gnx_twobit_get_bases(gnx_ctx *ctx, const gnx_twobit *tb, const int64_t *q_seq, const int64_t *q_pos,
                         int64_t n_q, uint8_t *out);
int gnx_twobit_count_matches(gnx_ctx *ctx, int dir, const gnx_twobit *one, const gnx_twobit *two,
                             const int64_t *q_one, const int64_t *q_start_one, const int64_t *q_two,
                             const int64_t *q_start_two, int64_t n_q, int64_t *out_matches);
int gnx_twobit_pack_device(gnx_ctx *ctx, const uint8_t *d_seq, int64_t n_bases, int lead, uint64_t *d_words,
                           void *cuda_stream);
/* dnaTwoBit.NewTwoBit on the HOST (no GPU involved; this is the packer gnx_affine_batch runs while it stages pageable
 * bytes, exposed so that a caller can prepare gnx_affine_batch_twobit input with the same threads): `count`
 * sequences of `len` bases each, back to back in `bases`, into (len + 31) / 32 words per sequence, first base in
 * bits 63:62, tail left-aligned (dna/dnaTwoBit/dnaTwoBit.go:28-42).  GNX_EBASE if a base is >= 4 (the reference
 * would OR its high bits into the neighbouring bases; words are then unspecified). */
int gnx_pack_twobit_host(const uint8_t *bases, int64_t count, int64_t len, uint64_t *words);

/* ---- perfect-match seeds of cmd/gsw --------------------------------------------- *
 * gnx_seed_index_new  genomeGraph.IndexGenomeIntoMap(genome, seedLen, seedStep) (genomeGraph/index.go:21-44)
 *                     for a genome whose nodes have no edges (a linear reference: one node per chromosome):
 *                     node s = genome_cat[node_off[s] .. node_off[s+1]).  The map is kept on the device as
 *                     entries sorted by key (dnaToNumber, genomeGraph/align.go:170-177), each key's locations
 *                     (ChromAndPosToNumber, :163-168) in the reference's insertion order, together with the
 *                     nodes' TwoBit encoding.  seedLen outside 2..32 is the reference's log.Fatalf (GNX_EARG).
 * gnx_seed_batch      genomeGraph.seedMapMemPool (genomeGraph/search.go:567-602) for every read
 *                     reads_cat[read_off[r] .. read_off[r+1]): both strands (dna.ReverseComplement), every
 *                     readStart, every hit extended with CountLeftMatches / extendToTheRightDev (:425-452).
 *                     Read r's seeds are out_seeds[out_seed_off[r] .. out_seed_off[r+1]) in the reference's
 *                     APPEND order, i.e. before its final SortSeedLen / heapSortSeeds (:596-600), which is an
 *                     unstable sort the caller keeps applying on its side.  NextPart is always nil without
 *                     edges, so gnx_seed carries the remaining SeedDev fields (genomeGraph/index.go:11-19).
 *                     GNX_ECAP: seed_cap too small (out_seed_off is filled). GNX_EBASE: a read byte > 12. */
typedef struct gnx_seed_index gnx_seed_index;
typedef struct {
    uint32_t target_id, target_start, query_start, length, pos_strand, total_length;
} gnx_seed;
int gnx_seed_index_new(gnx_ctx *ctx, const uint8_t *genome_cat, const int64_t *node_off, int64_t n_nodes,
                       int seed_len, int seed_step, gnx_seed_index **out);
void gnx_seed_index_free(gnx_seed_index *ix);
int gnx_seed_index_info(const gnx_seed_index *ix, int64_t *n_entries);
int gnx_seed_index_download(gnx_ctx *ctx, const gnx_seed_index *ix, uint64_t *out_key, uint64_t *out_loc);
int gnx_seed_batch(gnx_ctx *ctx, const gnx_seed_index *ix, const uint8_t *reads_cat, const int64_t *read_off,
                   int64_t n_reads, gnx_seed *out_seeds, int64_t *out_seed_off, int64_t seed_cap);

/* ---- the per-read driver of cmd/gsw ---------------------------------------------- *
 * gnx_gsw_batch   genomeGraph.GraphSmithWatermanToGiraf (genomeGraph/toGiraf.go:17-72) for every read of a block
 *                 against a genome graph WITHOUT edges (the gnx_seed_index's nodes): seeds (seedMapMemPool), their
 *                 ordering (heapSortSeeds / SortSeedLen), the seedCouldBeBetter early exit (genomeGraph/index.go:
 *                 102-121), LeftAlignTraversal / RightAlignTraversal's linear-gap DPs (gap -600) on the windows of
 *                 getLeftTargetBases / getRightBases, score / position / path bookkeeping and the cigar assembly
 *                 (cigar.Append, Concat, AppendSoftClips; routes in the reference's traceback order).  The seed and
 *                 extend steps of the whole block run as batched GPU calls, the reference's sequential loop is
 *                 replayed per read over their results.  paired != 0: reads 2k / 2k+1 are the Fwd / Rev mates of
 *                 WrapPairGiraf (:117-137) and the flags get setGirafFlags' pair bits.
 * A gnx_giraf is the part of giraf.Giraf the aligner computes (QName / Seq / Qual / MapQ = 255 / Notes are the
 * caller's); cigar ops are the bytes 'M','I','D','S'; n_cigar = -1 is Go's nil Cigar (no seed scored above 0).
 * GNX_ECAP: cigar_cap too small, *out_n_cigar holds the number of elements the block needs (records are filled). */
typedef struct {
    int32_t q_start, q_end; /* giraf.QStart, QEnd */
    int32_t pos_strand;     /* giraf.PosStrand */
    int32_t t_start, t_end; /* giraf.Path.TStart, TEnd */
    int32_t node;           /* giraf.Path.Nodes = [node]; -1: empty path */
    int64_t aln_score;      /* giraf.AlnScore */
    int32_t flag;           /* giraf.Flag (getGirafFlags, + setGirafFlags when paired) */
    int32_t n_cigar;        /* len(giraf.Cigar); -1: nil */
    int64_t cigar_off;      /* first element in out_cigar */
} gnx_giraf;
int gnx_gsw_batch(gnx_ctx *ctx, const gnx_seed_index *ix, const uint8_t *reads_cat, const int64_t *read_off,
                  int64_t n_reads, const int64_t *scores, int dim, int paired, gnx_giraf *out, gnx_cigar *out_cigar,
                  int64_t cigar_cap, int64_t *out_n_cigar);

/* After a GNX_ECAP return: copy the retained cigars of the last batch call (total = the last
 * entry of that call's out_cigar_off). */
int gnx_copy_last_cigars(gnx_ctx *ctx, gnx_cigar *out_cigar, int64_t cigar_cap);

/* ---- device-resident form ------------------------------------------------------------------- *
 * Same computation with every array already in this context's device memory (pointers are device
 * pointers; *_off_host are optional host copies of the offset arrays used for planning -- pass
 * NULL to let the library read them back).  Work is enqueued on `cuda_stream` (a cudaStream_t;
 * NULL = the legacy default stream) and the call returns without synchronising unless it has to
 * read offsets back.  kind: 0 affine global, 1 affine free-end, 2 const gap (gap_open = penalty).
 * d_out_cigar_off (n_pairs+1) and d_out_cigar (cigar_cap entries) may be NULL when want_cigar=0.
 * d_status (one int32, may be NULL) receives a GNX_E* code discovered on the device
 * (GNX_EBASE, GNX_ECAP).
 * Ordering: successive calls on one context share its scratch, so the library orders each call after the
 * previous one (an event wait on `cuda_stream`), whatever streams they use; device buffers growing between
 * calls (cudaFree) synchronise the device.  Calls on DIFFERENT contexts are independent.  The host-buffer
 * entry points must not run concurrently with a device-resident call on the same context. */
int gnx_batch_device(gnx_ctx *ctx, int kind, const uint8_t *d_alpha_cat, const int64_t *d_alpha_off,
                     const uint8_t *d_beta_cat, const int64_t *d_beta_off,
                     const int64_t *alpha_off_host, const int64_t *beta_off_host, int64_t n_pairs,
                     const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
                     int want_cigar, int64_t *d_out_score, gnx_cigar *d_out_cigar,
                     int64_t *d_out_cigar_off, int64_t cigar_cap, int32_t *d_status, void *cuda_stream);

/* ---- several GPUs behind one call --------------------------------------------------- *
 * A gnx_multi owns one context per listed device.  A batch is cut into contiguous shards balanced by DP
 * cells (sum of n*m), shard r runs on device r from its own host thread, scores land directly in the
 * caller's array and the cigars are stitched into the caller's buffer in pair order: the result is
 * identical to the single-device call (same arguments, same error codes -- when several shards fail, the
 * one holding the lowest pair indices reports).  There is no data-path collective: pairs are independent.
 * This is what a Go process binds to use a whole box from ONE cgo call (align.AffineGap batches of
 * cmd/gsw-style callers); `devices` may list a device more than once (two contexts on it).
 *   gnx_multi_create      devices == NULL or n_devices <= 0: every visible device.  workspace: per device,
 *                         0 = the default of gnx_create.
 *   gnx_multi_shard_bounds  the n_devices + 1 pair indices a call with these offsets would cut at.
 *   gnx_multi_context     the context of shard `index` (options, statistics); owned by the gnx_multi. */
typedef struct gnx_multi gnx_multi;
gnx_multi *gnx_multi_create(const int *devices, int n_devices, size_t workspace_bytes_per_device);
void gnx_multi_destroy(gnx_multi *mg);
int gnx_multi_device_count(const gnx_multi *mg);
const char *gnx_multi_last_error(gnx_multi *mg); /* mg may be NULL: error of the last failed gnx_multi_create */
gnx_ctx *gnx_multi_context(gnx_multi *mg, int index);
int gnx_multi_shard_bounds(const gnx_multi *mg, const int64_t *alpha_off, const int64_t *beta_off, int64_t n_pairs,
                           int64_t *out_bounds);
int gnx_multi_affine_batch(gnx_multi *mg, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                           const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                           int64_t gap_extend, int mode, int want_cigar, int64_t *out_score, gnx_cigar *out_cigar,
                           int64_t *out_cigar_off, int64_t cigar_cap);
int gnx_multi_const_batch(gnx_multi *mg, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                          const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_pen,
                          int want_cigar, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off,
                          int64_t cigar_cap);
int gnx_multi_copy_last_cigars(gnx_multi *mg, gnx_cigar *out_cigar, int64_t cigar_cap);

/* Device-resident form of gnx_affine_batch_twobit for a UNIFORM batch: sequence p's words at
 * d_alpha_words + p * ceil(alpha_len / 32) (16-byte aligned arrays, readable 256 bytes past their end: the TMA of
 * a tail quad fetches a whole quad's words).  kind: 0 affine global, 1 affine free-end. */
int gnx_batch_device_twobit(gnx_ctx *ctx, int kind, const uint64_t *d_alpha_words, int64_t alpha_len,
                            const uint64_t *d_beta_words, int64_t beta_len, int64_t n_pairs, const int64_t *scores,
                            int dim, int64_t gap_open, int64_t gap_extend, int want_cigar, int64_t *d_out_score,
                            gnx_cigar *d_out_cigar, int64_t *d_out_cigar_off, int64_t cigar_cap, int32_t *d_status,
                            void *cuda_stream);

/* ---- introspection (used by bench.py / tests) ------------------------------------------------ */
/* Kernel launches issued by this context since creation (every launch of a libgnxalign kernel). */
int64_t gnx_launch_count(gnx_ctx *ctx);
/* Device time (ms, CUDA events on the launching stream) and launches of the DP fill kernels of
 * the last batch call; used for the roofline line of bench.py. */
int gnx_last_fill_stats(gnx_ctx *ctx, double *fill_ms, int64_t *fill_launches, int64_t *cells);
/* Kernel family the last batch call on this context planned: impl 1 first-generation kernels, 3 affine_fill3 /
 * const_fill3 (int32), 16 affine_fill16 (packed 16-bit, score only), 17 affine_fill16 with checkpoints +
 * affine_ckpt_trace (read-sized traceback), 18 affine_long (tile checkpoints, long pairs); flags: 1 ragged batch on
 * host-binned quads, 2 dnaTwoBit words staged by TMA, 4 int64 fallback, 8 multi-strip pairs. */
int gnx_last_kernel_path(gnx_ctx *ctx, int *impl, int *flags);
/* Tuning knobs (name/value), e.g. "cols_per_lane", "block_threads", "chunk_pairs"; "pack_stage" (default 1):
 * gnx_affine_batch packs large uniform batches to dnaTwoBit words on the host threads while it stages them
 * (a quarter of the PCIe bytes; falls back to bytes when a base >= 4 is met); "tb_tma" (default 1): 2-bit input is
 * read by the packed 16-bit kernels through TMA instead of being expanded first.  Returns GNX_EARG for an unknown
 * name. */
int gnx_set_option(gnx_ctx *ctx, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* GNXALIGN_H */

// gnx_twobit.cuh -- gonomics' dna/dnaTwoBit on the device: packing, GetBase,
// CountRightMatches / CountLeftMatches, and the perfect-match seed step of cmd/gsw for a linear reference.
//
// Reference (paths relative to the gonomics tree @ bd66b49b):
//   dna/dnaTwoBit/dnaTwoBit.go:28-78      BasesToUint64LeftAln / GetBase / NewTwoBit
//   dna/dnaTwoBit/rainbow.go:8-25         NewTwoBitRainbow: element k = NewTwoBit(k x 'A' + seq)
//   dna/dnaTwoBit/perfectAlign.go:10-85   CountRightMatches / CountLeftMatches
//   genomeGraph/index.go:21-44            IndexGenomeIntoMap
//   genomeGraph/search.go:425-452,567-602 extendToTheRightDev / seedMapMemPool
//
// These are HBM-bound byte/bit kernels: the pack kernel moves 1 byte in and 0.25 byte
// out per base with 16-byte loads and 8-byte coalesced stores; no shared-memory staging is needed because
// every input byte is consumed by exactly one thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnx {

// A set of TwoBit sequences resident on the device: sequence s owns words[word_off[s] .. word_off[s+1])
// and has TwoBit.Len = len[s] (which includes the rainbow lead).
struct TwoBitView {
    const uint64_t *words;
    const int64_t *word_off;
    const int64_t *len;
    int64_t n_seqs;
};

// Four bases of one little-endian 32-bit word (byte 0 = first base) -> 8 bits, first base in bits 7:6.
// Valid when every byte <= 3: the four 2-bit fields land on disjoint bits of the product's top byte.
__device__ __forceinline__ unsigned pack4(unsigned x) { return x * 0x40100401u; } // result in bits 31:24

// Exact OR-shift packing of 32 raw bytes held in x[0..7]: word = OR_i byte_i << (62 - 2i) (mod 2^64),
// i.e. BasesToUint64LeftAln's `answer<<2 | uint64(base)` including the spill of bases > 3 into the bits of
// the bases before them (dnaTwoBit.go:33-37).
__device__ __forceinline__ uint64_t pack32_exact(const unsigned (&x)[8])
{
    uint64_t w = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint64_t v = (x[j] >> (8 * b)) & 0xffu;
            const int sh = 62 - 2 * (4 * j + b);
            w |= sh >= 0 ? (v << sh) : (v >> -sh); // never negative; kept for clarity
        }
    }
    return w;
}

__device__ __forceinline__ uint64_t pack32(const unsigned (&x)[8])
{
    const unsigned any = x[0] | x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7];
    if ((any & 0xfcfcfcfcu) == 0) { // every base is A, C, G or T: multiply-gather, 2 instructions per 4 bases
        const unsigned hi = __byte_perm(__byte_perm(pack4(x[3]), pack4(x[2]), 0x0073), __byte_perm(pack4(x[1]), pack4(x[0]), 0x0073), 0x5410);
        const unsigned lo = __byte_perm(__byte_perm(pack4(x[7]), pack4(x[6]), 0x0073), __byte_perm(pack4(x[5]), pack4(x[4]), 0x0073), 0x5410);
        return ((uint64_t)hi << 32) | lo;
    }
    return pack32_exact(x);
}

struct PackParams {
    const uint8_t *seq;       // concatenated dna.Base bytes
    const int64_t *seq_off;   // n_seqs + 1 byte offsets
    const int64_t *word_off;  // n_seqs + 1 output word offsets
    uint64_t *words;
    int64_t n_seqs, total_words, total_bytes;
    int64_t uniform_words;    // > 0: every sequence has this many words (word -> sequence by division)
    int lead;                 // rainbow lead: `lead` x dna.A prepended to every sequence (0..31)
};

// One thread per output word.
__global__ void __launch_bounds__(256) twobit_pack_kernel(const PackParams P)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.total_words)
        return;
    int64_t s;
    if (P.n_seqs == 1) {
        s = 0;
    } else if (P.uniform_words > 0) {
        s = w / P.uniform_words;
    } else { // largest s with word_off[s] <= w
        int64_t lo = 0, hi = P.n_seqs;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (P.word_off[mid] <= w)
                lo = mid;
            else
                hi = mid;
        }
        s = lo;
    }
    const int64_t k = w - P.word_off[s];
    const int64_t B = P.seq_off[s], L = P.seq_off[s + 1] - B;
    const int64_t p0 = 32 * k - P.lead;   // sequence position of the word's first base (negative: lead 'A's)
    const int64_t A = B + p0;             // its byte address in seq
    const int lo_v = p0 < 0 ? (int)-p0 : 0;                  // window bytes [lo_v, hi_v) are real bases
    const int hi_v = (int)(L - p0 < 32 ? L - p0 : 32);
    unsigned x[8];
    if (lo_v == 0 && hi_v == 32 && (A & 15) == 0) { // interior word, 16-byte aligned: two vector loads
        const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(P.seq + A));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(P.seq + A + 16));
        x[0] = q0.x, x[1] = q0.y, x[2] = q0.z, x[3] = q0.w;
        x[4] = q1.x, x[5] = q1.y, x[6] = q1.z, x[7] = q1.w;
    } else {
        // nine aligned 4-byte loads cover any 32-byte window; funnel shifts realign it
        const int64_t A4 = A & ~(int64_t)3; // floor to a multiple of 4 (also for negative A)
        const unsigned sh = (unsigned)(A - A4) * 8;
        unsigned raw[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int64_t a = A4 + 4 * j;
            unsigned v = 0;
            if (a >= 0 && a + 4 <= P.total_bytes) {
                v = __ldg(reinterpret_cast<const unsigned *>(P.seq + a));
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (a + b >= 0 && a + b < P.total_bytes)
                        v |= (unsigned)P.seq[a + b] << (8 * b);
            }
            raw[j] = v;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            x[j] = __funnelshift_r(raw[j], raw[j + 1], sh);
        if (lo_v > 0 || hi_v < 32) { // clear the bytes that belong to the lead / a neighbouring sequence
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                unsigned m = 0xffffffffu;
                const int l = lo_v - 4 * j, h = hi_v - 4 * j;
                if (l > 0)
                    m &= l >= 4 ? 0u : (0xffffffffu << (8 * l));
                if (h < 4)
                    m &= h <= 0 ? 0u : (0xffffffffu >> (8 * (4 - h)));
                x[j] &= m;
            }
        }
    }
    P.words[w] = pack32(x);
}

// GetBase (dnaTwoBit.go:59-65) for every position of every sequence: one thread per word, 32 bytes out.
struct UnpackParams {
    TwoBitView tb;
    const int64_t *out_off; // n_seqs + 1 byte offsets of the output (prefix sums of len)
    uint8_t *out;
    int64_t total_words;
    int64_t uniform_words;
};

__global__ void __launch_bounds__(256) twobit_unpack_kernel(const UnpackParams P)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.total_words)
        return;
    int64_t s;
    if (P.tb.n_seqs == 1) {
        s = 0;
    } else if (P.uniform_words > 0) {
        s = w / P.uniform_words;
    } else {
        int64_t lo = 0, hi = P.tb.n_seqs;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (P.tb.word_off[mid] <= w)
                lo = mid;
            else
                hi = mid;
        }
        s = lo;
    }
    const int64_t k = w - P.tb.word_off[s];
    const int64_t L = P.tb.len[s];
    const uint64_t word = P.tb.words[w];
    uint8_t *dst = P.out + P.out_off[s] + 32 * k;
    const int cnt = (int)(L - 32 * k < 32 ? L - 32 * k : 32);
    unsigned x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const unsigned r = (unsigned)(word >> (56 - 8 * j)) & 0xffu; // four bases, first in bits 7:6
        x[j] = (r >> 6) | (((r >> 4) & 3u) << 8) | (((r >> 2) & 3u) << 16) | ((r & 3u) << 24);
    }
    if (cnt == 32 && ((uintptr_t)dst & 15) == 0) {
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(x[0], x[1], x[2], x[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(x[4], x[5], x[6], x[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (4 * j + b < cnt)
                    dst[4 * j + b] = (uint8_t)(x[j] >> (8 * b));
    }
}

// Uniform batches (every sequence `len` bases = `wps` words, tightly packed): GetBase of every position without
// any offset array -- sequence s = word / wps writes its bytes at out + s * len.  One thread per word.
__global__ void __launch_bounds__(256) twobit_unpack_uniform_kernel(const uint64_t *words, int64_t total_words, int64_t wps,
                                                                    int64_t len, uint8_t *out)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total_words)
        return;
    const int64_t s = w / wps, k = w - s * wps;
    const uint64_t word = words[w];
    uint8_t *dst = out + s * len + 32 * k;
    const int cnt = (int)(len - 32 * k < 32 ? len - 32 * k : 32);
    unsigned x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const unsigned r = (unsigned)(word >> (56 - 8 * j)) & 0xffu; // four bases, first in bits 7:6
        x[j] = (r >> 6) | (((r >> 4) & 3u) << 8) | (((r >> 2) & 3u) << 16) | ((r & 3u) << 24);
    }
    if (cnt == 32 && ((uintptr_t)dst & 15) == 0) {
        reinterpret_cast<uint4 *>(dst)[0] = make_uint4(x[0], x[1], x[2], x[3]);
        reinterpret_cast<uint4 *>(dst)[1] = make_uint4(x[4], x[5], x[6], x[7]);
    } else if (cnt == 32 && ((uintptr_t)dst & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            reinterpret_cast<unsigned *>(dst)[j] = x[j];
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (4 * j + b < cnt)
                    dst[4 * j + b] = (uint8_t)(x[j] >> (8 * b));
    }
}

// off[k] = (first + k) * stride for k = 0 .. count-1: the byte offsets of a uniform batch, made on the device
// instead of shipped over PCIe.
__global__ void __launch_bounds__(256) iota_offsets_kernel(int64_t *off, int64_t first, int64_t count, int64_t stride)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count)
        off[k] = (first + k) * stride;
}

// GetBase for a list of (sequence, position) queries.  A position beyond the sequence's last word is Go's
// index-out-of-range panic (status kEIndex, first offending query).
constexpr int kEOffset = 9, kEIndex = 10;

// first_bad holds min over failing elements of (element << 8 | code): one atomic, so the code always belongs
// to the smallest failing index (the error a sequential caller would hit first).
__device__ __forceinline__ void report(int *status, int64_t *first_bad, int code, int64_t q)
{
    (void)status;
    atomicMin((unsigned long long *)first_bad, ((unsigned long long)q << 8) | (unsigned)code);
}

__global__ void __launch_bounds__(256) twobit_get_bases_kernel(const TwoBitView tb, const int64_t *q_seq, const int64_t *q_pos,
                                                               int64_t n_q, uint8_t *out, int *status, int64_t *first_bad)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q)
        return;
    const int64_t s = q_seq[q], pos = q_pos[q];
    if (s < 0 || s >= tb.n_seqs || pos < 0 || pos / 32 >= tb.word_off[s + 1] - tb.word_off[s]) {
        report(status, first_bad, kEIndex, q);
        out[q] = 0;
        return;
    }
    const uint64_t word = tb.words[tb.word_off[s] + pos / 32];
    out[q] = (uint8_t)((word >> (64 - 2 * (pos % 32 + 1))) & 3);
}

// CountRightMatches (perfectAlign.go:10-47)
__device__ __forceinline__ int64_t count_right(const uint64_t *one, int64_t one_len, const uint64_t *two, int64_t two_len,
                                               int64_t start_one, int64_t start_two)
{
    const int offset = (int)(start_one & 31) * 2;
    int64_t i = start_one >> 5, j = start_two >> 5;
    const int64_t i_end = (one_len + 31) >> 5, j_end = (two_len + 31) >> 5;
    uint64_t diff = (one[i] ^ two[j]) & (~0ull >> offset);
    int bit_matches = __clzll((long long)diff); // 64 for diff == 0
    int64_t total = bit_matches - offset;
    for (++i, ++j; i < i_end && j < j_end && bit_matches == 64; ++i, ++j) {
        diff = one[i] ^ two[j];
        bit_matches = __clzll((long long)diff);
        total += bit_matches;
    }
    const int64_t r = total / 2;
    return min(min(r, one_len - start_one), two_len - start_two);
}

// CountLeftMatches (perfectAlign.go:49-85)
__device__ __forceinline__ int64_t count_left(const uint64_t *one, const uint64_t *two, int64_t start_one, int64_t start_two)
{
    const int offset = (int)(start_one & 31) * 2;
    const int no_look = 64 - offset - 2;
    int64_t i = start_one >> 5, j = start_two >> 5;
    uint64_t diff = (one[i] ^ two[j]) & (~0ull << no_look);
    int bit_matches = diff ? __ffsll((long long)diff) - 1 : 64;
    int64_t total = bit_matches - no_look;
    for (--i, --j; i >= 0 && j >= 0 && bit_matches == 64; --i, --j) {
        diff = one[i] ^ two[j];
        bit_matches = diff ? __ffsll((long long)diff) - 1 : 64;
        total += bit_matches;
    }
    return total / 2;
}

__global__ void __launch_bounds__(256) twobit_count_kernel(int dir, const TwoBitView one, const TwoBitView two, const int64_t *q_one,
                                                           const int64_t *q_start_one, const int64_t *q_two,
                                                           const int64_t *q_start_two, int64_t n_q, int64_t *out, int *status,
                                                           int64_t *first_bad)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q)
        return;
    const int64_t s1 = q_one[q], s2 = q_two[q], a = q_start_one[q], b = q_start_two[q];
    out[q] = 0;
    if (s1 < 0 || s1 >= one.n_seqs || s2 < 0 || s2 >= two.n_seqs || a < 0 || b < 0) {
        report(status, first_bad, kEIndex, q);
        return;
    }
    if ((a & 31) != (b & 31)) { // log.Fatalf "Different offsets" (:24-26, :63-65) comes before any indexing
        report(status, first_bad, kEOffset, q);
        return;
    }
    const int64_t w1 = one.word_off[s1], n1 = one.word_off[s1 + 1] - w1;
    const int64_t w2 = two.word_off[s2], n2 = two.word_off[s2 + 1] - w2;
    if ((a >> 5) >= n1 || (b >> 5) >= n2) { // Seq[i] / Seq[j] index out of range
        report(status, first_bad, kEIndex, q);
        return;
    }
    out[q] = dir == 0 ? count_right(one.words + w1, one.len[s1], two.words + w2, two.len[s2], a, b)
                      : count_left(one.words + w1, two.words + w2, a, b);
}


// ------------------------------------------------------------------------------------------------
// Seed index: genomeGraph.IndexGenomeIntoMap (genomeGraph/index.go:21-44) for edge-less nodes as a
// (key, location)-sorted array plus a bucket table over the key's top bits.  A Go map[uint64][]uint64
// filled in (node, pos) order is exactly "entries grouped by key, each group in insertion order", which
// a STABLE sort by key of entries emitted in (node, pos) order reproduces.
// ------------------------------------------------------------------------------------------------
struct SeedEmitParams {
    const uint8_t *genome;     // concatenated node bytes
    const int64_t *node_off;   // n_nodes + 1
    const int64_t *cand_off;   // n_nodes + 1: prefix sums of the candidate positions per node
    int64_t n_nodes, n_cand;
    int seed_len, seed_step;
    uint64_t *key, *loc;       // per candidate
    int *valid;                // per candidate: 1 = indexed (window holds no dna.N)
};

__global__ void __launch_bounds__(256) seed_emit_kernel(const SeedEmitParams P)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cand)
        return;
    int64_t lo = 0, hi = P.n_nodes;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (P.cand_off[mid] <= c)
            lo = mid;
        else
            hi = mid;
    }
    const int64_t node = lo, pos = (c - P.cand_off[node]) * P.seed_step;
    const uint8_t *seq = P.genome + P.node_off[node] + pos;
    // dna.CountBaseInterval(seq, dna.N, pos, pos+seedLen) == 0 (index.go:30) and dnaToNumber (align.go:170-177):
    // answer = seq[start]; then answer<<2 | seq[i] -- the raw byte, so bases > 4 spill upwards
    uint64_t key = 0;
    bool has_n = false;
    for (int i = 0; i < P.seed_len; ++i) {
        const uint8_t b = seq[i];
        has_n |= b == 4;
        key = (key << 2) | b;
    }
    P.key[c] = key;
    P.loc[c] = ((uint64_t)node << 32) | (uint64_t)pos; // ChromAndPosToNumber (align.go:163-168)
    P.valid[c] = has_n ? 0 : 1;
}

__global__ void __launch_bounds__(256) seed_compact_kernel(const uint64_t *key, const uint64_t *loc, const int *valid,
                                                           const int64_t *dst, int64_t n_cand, uint64_t *okey, uint64_t *oloc)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand || !valid[c])
        return;
    okey[dst[c]] = key[c];
    oloc[dst[c]] = loc[c];
}

struct SeedIndexView {
    const uint64_t *key, *loc; // sorted by key (stable)
    int64_t n;
    const int64_t *bucket;     // (1 << bucket_bits) + 1 entries: lower_bound(b << bucket_shift); last = n
    int bucket_bits, bucket_shift;
};

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *a, int64_t lo, int64_t hi, uint64_t v)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) seed_bucket_kernel(const uint64_t *key, int64_t n, int bucket_bits, int bucket_shift,
                                                          int64_t *bucket)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nb = (int64_t)1 << bucket_bits;
    if (b > nb)
        return;
    bucket[b] = b == nb ? n : lower_bound_u64(key, 0, n, (uint64_t)b << bucket_shift);
}

// [lo, hi) = the run of entries whose key equals `k` (seedHash[k])
__device__ __forceinline__ void seed_lookup(const SeedIndexView &ix, uint64_t k, int64_t &lo, int64_t &hi)
{
    const uint64_t nb1 = ((uint64_t)1 << ix.bucket_bits) - 1;
    uint64_t b = ix.bucket_shift >= 64 ? 0 : (k >> ix.bucket_shift);
    b = b < nb1 ? b : nb1; // keys with spilled high bits all live in the last bucket
    const int64_t s = ix.bucket[b], e = b == nb1 ? ix.n : ix.bucket[b + 1];
    lo = lower_bound_u64(ix.key, s, e, k);
    hi = lo;
    while (hi < e && ix.key[hi] == k)
        ++hi;
}

// ------------------------------------------------------------------------------------------------
// seedMapMemPool (genomeGraph/search.go:567-602) for a batch of reads against edge-less nodes: one warp
// per read.  For every readStart and strand the read's 2-bit k-mer is looked up; every hit is extended
// to the left (CountLeftMatches, clipped at the read start) and then to the right from the new start
// (extendToTheRightDev :425-452, which without edges yields one seed or nil).  Seeds come out in the
// reference's APPEND order (readStart ascending; forward hits, then reverse hits; hits in index order);
// the caller applies the reference's sort (SortSeedLen / heapSortSeeds) afterwards.
//
// The read's 32 rainbow encodings (fastq.FastqBig.Rainbow / RainbowRc) are never materialised: word j of
// rainbow[off] is a funnel shift of the offset-0 words kept in shared memory when every base is A/C/G/T,
// and the exact OR-shift of the staged bytes otherwise (bases > 3 spill into their neighbours' bits
// differently in each rainbow element).
// ------------------------------------------------------------------------------------------------
struct SeedRec {
    uint32_t target_id, target_start, query_start, length, pos_strand, total_length; // genomeGraph.SeedDev (index.go:11-19)
};

struct SeedParams {
    const uint8_t *reads;
    const int64_t *read_off; // n_reads + 1
    int64_t n_reads;
    SeedIndexView ix;
    TwoBitView genome;
    int seed_len;
    int max_len;            // longest read of the batch (shared-memory pitch)
    int pass;               // 0: count hits per read; 1: write seeds
    int *hit_count;         // pass 0 out
    const int64_t *tmp_off; // pass 1: where read r's seeds go (prefix sums of hit_count)
    SeedRec *tmp;           // pass 1 out
    int *seed_count;        // pass 1 out: seeds actually produced per read (<= hit_count)
    int *status;
    int64_t *first_bad;
};

struct ReadView { // one strand of one read in shared memory
    const uint8_t *bytes;
    const uint64_t *w0; // offset-0 words (valid when clean)
    int len, nw0;
    bool clean;
    // word j of NewTwoBitRainbow(read)[off]  (rainbow.go:8-25)
    __device__ __forceinline__ uint64_t word(int off, int j) const
    {
        if (clean) {
            const uint64_t cur = (j >= 0 && j < nw0) ? w0[j] : 0;
            if (off == 0)
                return cur;
            const uint64_t prev = (j - 1 >= 0 && j - 1 < nw0) ? w0[j - 1] : 0;
            return (prev << (64 - 2 * off)) | (cur >> (2 * off));
        }
        uint64_t w = 0;
        const int p0 = 32 * j - off;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int p = p0 + i;
            const uint64_t v = (p >= 0 && p < len) ? bytes[p] : 0;
            w |= v << (62 - 2 * i);
        }
        return w;
    }
};

__global__ void __launch_bounds__(128) seed_kernel(const SeedParams P)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = (P.max_len + 15) & ~15, nwmax = (P.max_len + 31) / 32;
    // per warp: fwd bytes, rc bytes, fwd words, rc words
    uint8_t *base = s_raw + (size_t)warp * (2 * pitch + 2 * nwmax * 8);
    uint8_t *s_fwd = base, *s_rc = base + pitch;
    uint64_t *s_wf = reinterpret_cast<uint64_t *>(base + 2 * pitch), *s_wr = s_wf + nwmax;
    const int warps_per_block = blockDim.x >> 5;
    const unsigned key_shift = 64 - 2 * (unsigned)P.seed_len;

    for (int64_t r = (int64_t)blockIdx.x * warps_per_block + warp; r < P.n_reads; r += (int64_t)gridDim.x * warps_per_block) {
        const int64_t b0 = P.read_off[r];
        const int len = (int)(P.read_off[r + 1] - b0);
        // stage the read and its reverse complement (dna.ReverseComplement, dna/modify.go:72,111-115)
        bool dirty = false, bad = false;
        for (int i = lane; i < len; i += 32) {
            const uint8_t f = P.reads[b0 + i], g = P.reads[b0 + len - 1 - i];
            // complementArray = {T,G,C,A,N,t,g,c,a,n,Gap,Dot,Nil}; a byte > 12 panics (index out of range)
            const uint8_t c = g <= 3 ? 3 - g : (g == 4 ? 4 : (g <= 8 ? 13 - g : g));
            s_fwd[i] = f;
            s_rc[i] = c;
            dirty |= f > 3;
            bad |= g > 12;
        }
        const bool clean = !__any_sync(FULL, dirty);
        if (__any_sync(FULL, bad)) {
            if (lane == 0)
                report(P.status, P.first_bad, 1 /* GNX_EBASE */, r);
            if (lane == 0) {
                if (P.pass == 0)
                    P.hit_count[r] = 0;
                else
                    P.seed_count[r] = 0;
            }
            continue;
        }
        __syncwarp();
        const int nw0 = (len + 31) / 32;
        if (clean) {
            for (int k = lane; k < 2 * nw0; k += 32) {
                const uint8_t *src = (k < nw0 ? s_fwd : s_rc) + 32 * (k % nw0);
                const int cnt = min(32, len - 32 * (k % nw0));
                uint64_t w = 0;
                for (int i = 0; i < cnt; ++i)
                    w |= (uint64_t)src[i] << (62 - 2 * i);
                (k < nw0 ? s_wf : s_wr)[k % nw0] = w;
            }
        }
        __syncwarp();
        ReadView rv[2];
        rv[0].bytes = s_fwd, rv[0].w0 = s_wf, rv[1].bytes = s_rc, rv[1].w0 = s_wr;
        rv[0].len = rv[1].len = len, rv[0].nw0 = rv[1].nw0 = nw0, rv[0].clean = rv[1].clean = clean;

        const int n_look = len - P.seed_len + 1 > 0 ? 2 * (len - P.seed_len + 1) : 0; // (readStart, strand) pairs
        int total_hits = 0; // pass 0
        int written = 0;    // pass 1: seeds written so far for this read
        SeedRec *out = P.pass == 1 ? P.tmp + P.tmp_off[r] : nullptr;
        for (int l0 = 0; l0 < n_look; l0 += 32) {
            // each lane looks one (readStart, strand) up; order = lookup index
            const int l = l0 + lane;
            int64_t lo = 0, hi = 0;
            if (l < n_look) {
                const int read_start = l >> 1, strand = l & 1;
                const int key_idx = (read_start + 31) / 32, key_off = 31 - ((read_start + 31) % 32);
                const uint64_t key = rv[strand].word(key_off, key_idx) >> key_shift;
                seed_lookup(P.ix, key, lo, hi);
            }
            const int cnt = (int)(hi - lo);
            // inclusive prefix sum of the hit counts over the lanes
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o)
                    incl += v;
            }
            const int batch_hits = __shfl_sync(FULL, incl, 31);
            if (P.pass == 0) {
                total_hits += batch_hits;
                continue;
            }
            // pass 1: hit k of this batch (k in lookup order, then index order) goes to lane k % 32
            for (int k0 = 0; k0 < batch_hits; k0 += 32) {
                const int k = k0 + lane;
                // owner = first lane whose inclusive sum exceeds k
                int owner = 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int probe = owner + o - 1;
                    const int v = __shfl_sync(FULL, incl, probe);
                    if (v <= k)
                        owner += o;
                }
                // (owner ends as the number of lanes with incl <= k, i.e. the owning lane, when k < batch_hits)
                const int own = min(owner, 31);
                const int o_incl = __shfl_sync(FULL, incl, own), o_cnt = __shfl_sync(FULL, cnt, own);
                const int64_t o_lo = __shfl_sync(FULL, lo, own);
                bool valid = false;
                SeedRec rec;
                if (k < batch_hits) {
                    const int ol = l0 + own;
                    const int read_start = ol >> 1, strand = ol & 1;
                    const int64_t h = o_lo + (k - (o_incl - o_cnt));
                    const uint64_t code = P.ix.loc[h];
                    const int64_t node = (int64_t)(code >> 32), node_pos = (int64_t)(code & 0xffffffffu); // numberToChromAndPos
                    const uint64_t *nw = P.genome.words + P.genome.word_off[node];
                    const int64_t node_len = P.genome.len[node];
                    const ReadView &q = rv[strand];
                    // CountLeftMatches(node, nodePos, rainbow[readOffset], readStart + readOffset)  (search.go:583)
                    int node_offset = (int)(node_pos & 31);
                    int read_offset = 31 - ((read_start - node_offset + 31) % 32);
                    int64_t left;
                    {
                        const int no_look = 64 - 2 * node_offset - 2;
                        int64_t i = node_pos >> 5, j = (read_start + read_offset) >> 5;
                        uint64_t diff = (nw[i] ^ q.word(read_offset, (int)j)) & (~0ull << no_look);
                        int bm = diff ? __ffsll((long long)diff) - 1 : 64;
                        int64_t total = bm - no_look;
                        for (--i, --j; i >= 0 && j >= 0 && bm == 64; --i, --j) {
                            diff = nw[i] ^ q.word(read_offset, (int)j);
                            bm = diff ? __ffsll((long long)diff) - 1 : 64;
                            total += bm;
                        }
                        left = min((int64_t)read_start + 1, total / 2);
                    }
                    // extendToTheRightDev(node, read, readStart-(left-1), nodePos-(left-1), strand)  (:425-452)
                    const int64_t r_start = read_start - (left - 1), n_start = node_pos - (left - 1);
                    node_offset = (int)(n_start & 31);
                    read_offset = 31 - (int)((r_start - node_offset + 31) % 32);
                    int64_t right;
                    {
                        const int offset = node_offset * 2;
                        const int64_t two_len = len + read_offset, start_two = r_start + read_offset;
                        int64_t i = n_start >> 5, j = start_two >> 5;
                        const int64_t i_end = (node_len + 31) >> 5, j_end = (two_len + 31) >> 5;
                        uint64_t diff = (nw[i] ^ q.word(read_offset, (int)j)) & (~0ull >> offset);
                        int bm = __clzll((long long)diff);
                        int64_t total = bm - offset;
                        for (++i, ++j; i < i_end && j < j_end && bm == 64; ++i, ++j) {
                            diff = nw[i] ^ q.word(read_offset, (int)j);
                            bm = __clzll((long long)diff);
                            total += bm;
                        }
                        right = min(min(total / 2, node_len - n_start), two_len - start_two);
                    }
                    valid = right != 0; // "nothing aligned here" -> nil (:440-442)
                    rec.target_id = (uint32_t)node;
                    rec.target_start = (uint32_t)n_start;
                    rec.query_start = (uint32_t)r_start;
                    rec.length = (uint32_t)right;
                    rec.pos_strand = strand == 0;
                    rec.total_length = (uint32_t)right;
                }
                const unsigned bal = __ballot_sync(FULL, valid);
                if (valid)
                    out[written + __popc(bal & ((1u << lane) - 1))] = rec;
                written += __popc(bal);
            }
        }
        if (lane == 0) {
            if (P.pass == 0)
                P.hit_count[r] = total_hits;
            else
                P.seed_count[r] = written;
        }
        __syncwarp();
    }
}

// tmp (offsets by hits) -> out (offsets by seeds actually produced)
__global__ void __launch_bounds__(256) seed_gather_kernel(const SeedRec *tmp, const int64_t *tmp_off, const int64_t *out_off,
                                                          int64_t n_reads, SeedRec *out, int64_t cap)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < n_reads; r += (int64_t)gridDim.x * 8) {
        const int64_t s = tmp_off[r], d = out_off[r], n = out_off[r + 1] - d;
        for (int64_t i = lane; i < n; i += 32)
            if (d + i < cap)
                out[d + i] = tmp[s + i];
    }
}

} // namespace gnx

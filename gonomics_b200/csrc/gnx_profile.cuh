// gnx_profile.cuh -- group-vs-group ("profile") match scores for the progressive-MSA inner loop.
//
// Reference: scoreColumnMatch / ungappedRegionColumnScore (align/multiAlign.go:82-110), called three
// times per DP cell by multipleAffineGap[Chunk] (align/affineGap_highMem.go:274-353) and O(|x|*|y|)
// each time.  Here a group's column is reduced ONCE to a profile (counts of the folded bases, number
// of ungapped and of invalid bases) and a column pair's score becomes
//     trunc( sum_a cntA[a] * (sum_b cntB[b] * scores[a][b])  /  (ungappedA * ungappedB) )
// -- the same integer (Go's int64 division truncates toward zero, as C++'s does), because the
// reference's double loop adds scores[a][b] once per ungapped (a,b) pair.  The dense matrix of cell
// scores S[i][j] (sum over the chunk's column pairs) feeds affine_fill_kernel<LOOKUP = 3>.
#pragma once
#include "gnx_kernels.cuh"

namespace gnx {

constexpr int kProfW = 10;  // per column: counts of bases 0..7, ungapped count, invalid count
constexpr int kGapBase = 10; // dna.Gap (dna/dna.go:7-21); 5..9 are the lowercase bases
constexpr int kEDivZero = 8; // mirrors GNX_EDIVZERO

// One thread per profile column (all groups' columns are numbered consecutively: col_off[g] is the first
// column of group g).  Adjacent threads read adjacent bytes of each sequence row.
__global__ void profile_count_kernel(const uint8_t *group_cat, const int64_t *group_off, const int64_t *group_nseq,
                                     const int64_t *col_off, int n_groups, int dim, const int64_t *scores64,
                                     int *prof, long long *vb)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= col_off[n_groups])
        return;
    int lo = 0, hi = n_groups - 1; // last group whose first column is <= col (empty groups are skipped)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (col_off[mid] <= col)
            lo = mid;
        else
            hi = mid - 1;
    }
    const int g = lo;
    const int64_t len = col_off[g + 1] - col_off[g], c = col - col_off[g];
    const uint8_t *base = group_cat + group_off[g] + c;
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int ungapped = 0, invalid = 0;
    for (int64_t s = 0; s < group_nseq[g]; ++s) {
        int b = base[s * len];
        if (b >= 5 && b <= 9) // lowercase -> uppercase (multiAlign.go:87-89,92-94)
            b -= 5;
        if (b == kGapBase)
            continue;
        ++ungapped;
        if (b < dim)
            ++cnt[b];
        else
            ++invalid;
    }
    int *p = prof + col * kProfW;
#pragma unroll
    for (int a = 0; a < 8; ++a)
        p[a] = cnt[a];
    p[8] = ungapped;
    p[9] = invalid;
    // the column seen from the beta side: vb[a] = sum_b cnt[b] * scores[a][b]
    long long *v = vb + col * 8;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        long long acc = 0;
        if (a < dim)
            for (int b = 0; b < dim; ++b)
                acc += (long long)cnt[b] * scores64[a * dim + b];
        v[a] = acc;
    }
}

struct ProfileScoreParams {
    const int *prof;
    const long long *vb;
    const int64_t *col_off;    // first profile column of each group
    const int64_t *pair_x, *pair_y;
    const int64_t *smat_off;   // per pair (+1): first cell of its S matrix
    int64_t pair_begin, pair_end;
    int64_t smat_base;         // smat_off value of pair_begin (S is allocated per sub-batch)
    int chunk;
    int *smat;
    unsigned long long *first_error; // min over (pair, cell, code): the panic Go would hit first
};

// S[i][j] for every pair of the sub-batch; one thread per cell, consecutive threads on consecutive j.
__global__ void profile_score_kernel(const ProfileScoreParams P)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = P.smat_off[P.pair_end] - P.smat_base;
    if (cell >= total)
        return;
    int64_t lo = P.pair_begin, hi = P.pair_end - 1; // last pair whose first cell is <= cell
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (P.smat_off[mid] - P.smat_base <= cell)
            lo = mid;
        else
            hi = mid - 1;
    }
    const int64_t pair = lo;
    const int64_t gx = P.pair_x[pair], gy = P.pair_y[pair];
    const int64_t m = (P.col_off[gy + 1] - P.col_off[gy]) / P.chunk;
    const int64_t local = cell - (P.smat_off[pair] - P.smat_base);
    const int64_t i = local / m, j = local - i * m;
    const int *pa = P.prof + (P.col_off[gx] + i * P.chunk) * kProfW;
    const int *pb = P.prof + (P.col_off[gy] + j * P.chunk) * kProfW;
    const long long *vb = P.vb + (P.col_off[gy] + j * P.chunk) * 8;
    long long s = 0;
    int err = 0;
    for (int u = 0; u < P.chunk && !err; ++u, pa += kProfW, pb += kProfW, vb += 8) {
        const long long ua = pa[8], ub = pb[8];
        if ((pa[9] > 0 && ub > 0) || (pb[9] > 0 && ua > 0)) {
            err = kEBase; // scores[a][b] with a base >= dim: index out of range
        } else if (ua * ub == 0) {
            err = kEDivZero; // sum / count with count == 0 (multiAlign.go:101)
        } else {
            long long dot = 0;
#pragma unroll
            for (int a = 0; a < 8; ++a)
                dot += (long long)pa[a] * vb[a];
            s += dot / (ua * ub);
        }
    }
    if (err) {
        const unsigned long long key = ((unsigned long long)pair << 40) | ((unsigned long long)local << 4) | (unsigned)err;
        atomicMin(P.first_error, key);
        s = 0;
    }
    P.smat[cell] = (int)s;
}

} // namespace gnx

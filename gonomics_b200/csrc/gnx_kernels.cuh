// gnx_kernels.cuh -- sm_100a kernels for the gonomics `align` DP hot path.
//
// Mapping (DESIGN.md "Kernels"): one (alpha, beta) pair per warp.  The beta (query) axis is cut
// into strips of 32*C columns; lane l owns C consecutive columns of the strip and sweeps the alpha
// (target) rows with a one-row skew per lane (lane l is on row t-l at step t), so every cell's
// three neighbours are either in the lane's own registers or arrive from lane l-1 through two
// __shfl_up_sync per step.  Integer max-plus only: VIMNMX3 / VIADDMNMX (DPX) -- no tensor cores.
//
// Bit-exactness (align/align.go:76-84 tripleMaxTrace, ties M >= I >= D) is obtained by carrying
// every plane value as 64*v and adding a 2-bit tag (M=2, I=1, D=0) to the three candidates of a
// max in a field private to that max (I-plane: bits 1:0, D-plane: bits 3:2, H=max(M,I,D): bits
// 5:4).  One 3-way integer max then yields value and winning plane with the reference's
// tie order; the three tags of a cell sum into its 6-bit traceback code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnx {

constexpr int kTagBits = 6;
constexpr int kScale = 1 << kTagBits; // 64
constexpr int kFI = 1;                // tag field of the I-plane max  (code bits 1:0)
constexpr int kFD = 4;                // tag field of the D-plane max  (code bits 3:2)
constexpr int kFH = 16;               // tag field of H = max(M,I,D)   (code bits 5:4)
constexpr int kNeg32 = -(1 << 30);    // "-inf" for scaled int32 planes (real -2^24)

// status codes mirrored from gnxalign.h (device side)
constexpr int kOk = 0, kEBase = 1, kECap = 2;

struct FillParams {
    const uint8_t *alpha;
    const int64_t *alpha_off; // absolute offsets into alpha, indexed by global pair id
    const uint8_t *beta;
    const int64_t *beta_off;
    int64_t pair_begin, pair_end; // chunk = [pair_begin, pair_end)
    const uint8_t *pair_class;    // per global pair: 0 = ACGT only, 1 = has base in [4,dim), 2 = invalid
    int want_class;               // warps skip pairs whose class differs
    int gap_open, gap_extend;     // unscaled
    int h00;                      // H(0,0) = T(0, O, D(0,0)) unscaled
    int dim;
    int scores[64];               // unscaled, [a*dim + b]
    uint32_t *trace;              // chunk trace buffer (TRACE kernels)
    const int64_t *trace_off;     // per pair in chunk (index pair - pair_begin): offset in 32-bit words
    int2 *edge;                   // per-warp strip hand-off buffers: 2 * edge_stride int2 per warp
    int64_t edge_stride;
    int64_t *out_score;           // indexed by global pair id
    int chunk;                    // AffineGapChunk: bases per DP cell (1 otherwise); LOOKUP == 2 kernels only
    int one;                      // always 1: an opaque multiplier that keeps adds on the FMA pipe (IMAD)
    int64_t *out_best;            // const_fill3_kernel<EXT 2>: (row << 32 | column) of the first maximal cell, per global pair
    const int *smat;              // LOOKUP == 3: dense per-pair cell scores S[i][j] (gnx_profile.cuh), unscaled
    const int64_t *smat_off;      // LOOKUP == 3: first cell of pair p's matrix inside smat, indexed by global pair id
    // 2-bit inputs (affine_fill16_kernel<TB>): dnaTwoBit words of a UNIFORM batch, sequence p at word p * wn / p * wm
    // (indexed by global pair id: the pointers are biased by the chunk's first pair), lengths n_uni x m_uni
    const uint64_t *alpha_words, *beta_words;
    int wn, wm, n_uni, m_uni;
    // ragged batches on the packed 16-bit kernels: quads binned by (last-column index, target length) on the host
    const int *quad_pairs;        // 4 chunk-local pair indices per quad (-1 = empty slot; slot 0 always filled)
    const int64_t *quad_ck_off;   // CKPT: first checkpoint word of every quad
    int64_t quad_first, n_quads;  // this launch covers quads quad_first .. quad_first + n_quads - 1
    // dynamic schedule of the packed 16-bit kernels: a cursor that is zero at launch and zero again when the grid is
    // done (the fetch that draws n_quads + gridDim.x - 1 is the last one and resets it); nullptr = static round-robin
    unsigned *quad_ctr;
};

__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); } // max(a+b, c)
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
__device__ __forceinline__ long long addmax(long long a, long long b, long long c) { return max(a + b, c); }
__device__ __forceinline__ long long max3(long long a, long long b, long long c) { return max(max(a, b), c); }
// PRMT in its default mode: selector nibble bit 3 replicates the sign of the selected byte.
// (__byte_perm masks that bit off, so the sign-extending 16-bit table lookup needs the PTX form.)
__device__ __forceinline__ int prmt(int a, int b, int sel)
{
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Number of 32-bit trace words per lane per step: 5 six-bit codes per word.
__host__ __device__ constexpr int trace_wpl(int C) { return (C + 4) / 5; }
// Trace words for one pair: strips * (n + 31) steps * wpl * 32 lanes.
__host__ __device__ inline int64_t trace_words(int64_t n, int64_t m, int C)
{
    if (n <= 0 || m <= 0)
        return 0;
    const int64_t strips = (m + 32 * C - 1) / (32 * C);
    return strips * (n + 31) * trace_wpl(C) * 32;
}

// ------------------------------------------------------------------------------------------------
// classify: per pair, the largest base value decides the kernel class (and the GNX_EBASE error).
// One warp per pair, 16-byte vector loads where alignment allows.
// ------------------------------------------------------------------------------------------------
// maxbase: the largest base value of a byte range (one chunk's alpha or beta bytes), 16-byte loads,
// HBM-bound.  When it is < dim (the normal case) no pair can be invalid and the per-pair classify pass
// below returns at once (`gate`).
__global__ void maxbase_kernel(const uint8_t *bytes, int64_t lo, int64_t hi, int *out_max)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    unsigned mx = 0;
    // 16-byte aligned interior of the ADDRESS range (the base pointer itself may be biased/unaligned)
    const uintptr_t base = reinterpret_cast<uintptr_t>(bytes);
    const int64_t a_lo = min(hi, (int64_t)(((base + lo + 15) & ~uintptr_t(15)) - base));
    const int64_t a_hi = max(a_lo, (int64_t)(((base + hi) & ~uintptr_t(15)) - base));
    if (a_lo < a_hi) {
        const uint4 *v = reinterpret_cast<const uint4 *>(bytes + a_lo);
        const int64_t nv = (a_hi - a_lo) >> 4;
        for (int64_t i = tid; i < nv; i += nth) {
            const uint4 q = v[i];
            mx = __vmaxu4(mx, __vmaxu4(__vmaxu4(q.x, q.y), __vmaxu4(q.z, q.w)));
        }
        for (int64_t i = lo + tid; i < a_lo; i += nth)
            mx = max(mx, (unsigned)bytes[i]);
        for (int64_t i = a_hi + tid; i < hi; i += nth)
            mx = max(mx, (unsigned)bytes[i]);
    } else {
        for (int64_t i = lo + tid; i < hi; i += nth)
            mx = max(mx, (unsigned)bytes[i]);
    }
    mx = max(max(mx & 0xffu, (mx >> 8) & 0xffu), max((mx >> 16) & 0xffu, mx >> 24));
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx > 0)
        atomicMax(out_max, (int)mx);
}

__global__ void classify_kernel(const uint8_t *alpha, const int64_t *alpha_off, const uint8_t *beta,
                                const int64_t *beta_off, int64_t pair_begin, int64_t pair_end, int dim,
                                uint8_t *pair_class, int *status, const int *gate)
{
    if (gate && *gate < dim)
        return; // every base of the chunk is valid: pair_class stays all-zero
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p = pair_begin + warp; p < pair_end; p += nwarps) {
        unsigned mx = 0;
        const int64_t a0 = alpha_off[p], a1 = alpha_off[p + 1], b0 = beta_off[p], b1 = beta_off[p + 1];
        for (int64_t i = a0 + lane; i < a1; i += 32)
            mx = max(mx, (unsigned)alpha[i]);
        for (int64_t i = b0 + lane; i < b1; i += 32)
            mx = max(mx, (unsigned)beta[i]);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) {
            uint8_t cls = mx < 4 ? 0 : (mx < (unsigned)dim ? 1 : 2);
            if (a1 == a0 || b1 == b0)
                cls = cls == 2 ? 1 : cls; // Go never indexes the matrix when a side is empty: no panic
            pair_class[p] = cls;
            if (cls == 2)
                atomicMax(status, kEBase);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Affine-gap fill.  Reference recurrence: align/affineGap_highMem.go:181-220 (global and
// freeEndGaps), identical arithmetic in align/affineGap.go:159-191.
//   C      columns per lane (strip width 32*C)
//   TRACE  write the 6-bit/cell traceback codes (values scaled by 64 + tags) or score only
//   FREE   freeEndGaps (AffineGapLocal): D(i,0) = 0 and an un-penalised D in the last column
//   LOOKUP 0: ACGT-only pairs, substitution scores from per-column 16-bit tables via PRMT
//          1: any base < dim, scores from a shared-memory copy of the matrix
//          2: AffineGapChunk (align/affineGap_highMem.go:227-272): a DP cell is a block of P.chunk bases,
//             its match score the sum of the chunk's substitution scores (ungappedRegionScore,
//             align/ungapped.go:7-13); the host passes gap_extend * chunk as the gap step
//          3: profile DP (multipleAffineGap[Chunk], align/affineGap_highMem.go:274-353): the cell's match
//             score is read from the dense matrix P.smat that profile_score_kernel produced; lengths are
//             in bases and divided by P.chunk like LOOKUP 2; alpha/beta are not read
// ------------------------------------------------------------------------------------------------
//   V      plane value type: int (exact while analyse() proves the range) or long long (any input)
template <int C, bool TRACE, bool FREE, int LOOKUP, typename V = int>
__global__ void __launch_bounds__(128) affine_fill_kernel(const FillParams P)
{
    struct EdgeT {
        V x, y;
    };
    constexpr int SC = TRACE ? kScale : 1;
    constexpr int FI = TRACE ? kFI : 0, FD = TRACE ? kFD : 0, FH = TRACE ? kFH : 0;
    constexpr int WPL = trace_wpl(C);
    const V NEG = sizeof(V) == 8 ? (V)(-(1LL << 61)) : (V)kNeg32;
    const V CLRV = ~(V)(kScale - 1);
    constexpr unsigned FULL = 0xffffffffu;

    __shared__ int s_scores[64];
    if (LOOKUP >= 1) {
        if (threadIdx.x < 64)
            s_scores[threadIdx.x] = P.scores[threadIdx.x] * SC;
        __syncthreads();
    }
    const int chunk = LOOKUP >= 2 ? P.chunk : 1;

    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);

    const int O = P.gap_open, E = P.gap_extend;
    const V oe_s = (V)(O + E) * SC, e_s = (V)E * SC;
    // addends of the I-plane max (candidates M, I, D of the cell to the left)
    const V iM = oe_s + 2 * FI, iI = e_s + FI, iD = oe_s;
    // addends of the D-plane max (candidates M, I, D of the cell above), regular columns
    const V dM = oe_s + 2 * FD, dI = oe_s + FD, dD = e_s;

    EdgeT *edge_a = P.edge ? reinterpret_cast<EdgeT *>(P.edge) + (size_t)warp * 2 * P.edge_stride : nullptr;
    EdgeT *edge_b = P.edge ? edge_a + P.edge_stride : nullptr;

    for (int64_t pair = P.pair_begin + warp; pair < P.pair_end; pair += nwarps) {
        if (P.pair_class && (P.want_class < 0 ? P.pair_class[pair] > 1 : P.pair_class[pair] != P.want_class))
            continue; // want_class -1: every valid pair (classes 0 and 1)
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0) / chunk; // DP rows / columns (chunks when LOOKUP == 2)
        const int m = (int)(P.beta_off[pair + 1] - b0) / chunk;
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        const int *__restrict__ smat = LOOKUP == 3 ? P.smat + P.smat_off[pair] : nullptr;

        if (n == 0 || m == 0) { // closed forms of the boundary rows (affineGap_highMem.go:185-206)
            if (lane == 0) {
                int64_t sc;
                if (n == 0 && m == 0)
                    sc = P.h00;
                else if (n == 0)
                    sc = (int64_t)O + (int64_t)m * E; // I(0,m); M and D are -inf
                else
                    sc = FREE ? 0 : (int64_t)O + (int64_t)n * E; // D(n,0)
                P.out_score[pair] = sc;
            }
            continue;
        }

        const int T = n + 31; // steps per strip
        const int strips = (m + 32 * C - 1) / (32 * C);
        uint32_t *tbase = nullptr;
        if (TRACE && P.trace)
            tbase = P.trace + P.trace_off[pair - P.pair_begin];

        for (int p = 0; p < strips; ++p) {
            const int jbase = p * 32 * C + lane * C; // columns before mine; my columns are jbase+1..jbase+C
            // ---- per-column constants -----------------------------------------------------------
            int q[C];
            int t01[C], t23[C];       // LOOKUP 0: packed int16 score tables for target base 0,1 | 2,3
            V aM[C], aI[C], aD[C];    // D-plane addends (FREE: zero in the last column)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                if (LOOKUP == 3)
                    q[c] = j - 1;
                else
                    q[c] = (j <= m) ? (LOOKUP == 2 ? (j - 1) * chunk : (int)beta[j - 1]) : 0; // LOOKUP 2: chunk start
                if (LOOKUP == 0) {
                    const int s0 = P.scores[0 * P.dim + q[c]] * SC, s1 = P.scores[1 * P.dim + q[c]] * SC;
                    const int s2 = P.scores[2 * P.dim + q[c]] * SC, s3 = P.scores[3 * P.dim + q[c]] * SC;
                    t01[c] = (s0 & 0xffff) | (s1 << 16);
                    t23[c] = (s2 & 0xffff) | (s3 << 16);
                }
                const bool last = FREE && (j == m);
                aM[c] = last ? (V)(2 * FD) : dM;
                aI[c] = last ? (V)FD : dI;
                aD[c] = last ? (V)0 : dD;
            }
            // ---- row 0 state (affineGap_highMem.go:193-197): M=-inf, I=O+jE, D=-inf --------------
            V Dt[C];     // D(r, col) for the row about to be processed (tagged when TRACE)
            V Hc[C];     // clean H(r-1, col) of the previous row
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const V i0 = ((V)O + (V)j * E) * SC;
                Hc[c] = i0; // T(-inf, I, -inf) = I
                Dt[c] = addmax(NEG, aM[c], addmax(i0, aI[c], NEG + aD[c]));
            }
            V hpL = (jbase == 0) ? (V)P.h00 * SC : ((V)O + (V)jbase * E) * SC; // H(0, jbase)
            V edgeI = 0, edgeH = 0;                                     // what lane+1 consumes
            const EdgeT *ein = (p & 1) ? edge_b : edge_a;               // written by strip p-1
            EdgeT *eout = (p & 1) ? edge_a : edge_b;
            uint32_t *tp = (TRACE && tbase) ? tbase + ((size_t)p * T * WPL) * 32 + lane : nullptr;

            // lane 0 boundary stream for its next row (column jbase): I'(r, jbase+1) and H(r, jbase)
            V bI = 0, bH = 0;
            auto boundary = [&](int r) {
                if (p == 0) {
                    const V d0 = FREE ? (V)0 : ((V)O + (V)r * E) * SC; // D(r,0); M(r,0)=I(r,0)=-inf
                    bI = d0 + iD;                               // T(-inf, -inf, D+oe): tag D (0)
                    bH = d0;
                } else {
                    const EdgeT v = ein[r];
                    bI = v.x;
                    bH = v.y;
                }
            };
            if (lane == 0)
                boundary(1);
            int a_next = 0;
            if (lane == 0 && LOOKUP < 2)
                a_next = alpha[0];

            for (int t = 0; t < T; ++t) {
                const int r = t - lane + 1; // my row this step
                V inI = __shfl_up_sync(FULL, edgeI, 1);
                V inH = __shfl_up_sync(FULL, edgeH, 1);
                if (lane == 0) {
                    inI = bI;
                    inH = bH;
                }
                const bool active = (r >= 1) && (r <= n);
                const int a = a_next;
                if (LOOKUP < 2 && r + 1 >= 1 && r + 1 <= n)
                    a_next = alpha[r]; // prefetch the next row's base
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    int sel = 0, rowoff = 0;
                    if (LOOKUP == 0)
                        sel = a * 0x2222 + 0x9910; // PRMT selector: sign-extended 16-bit entry a
                    else
                        rowoff = a * P.dim;
                    V It = inI;
                    V hp = hpL;
                    uint32_t w[WPL];
#pragma unroll
                    for (int k = 0; k < WPL; ++k)
                        w[k] = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        int s;
                        if (LOOKUP == 0) {
                            s = prmt(t01[c], t23[c], sel);
                        } else if (LOOKUP == 1) {
                            s = s_scores[rowoff + q[c]];
                        } else if (LOOKUP == 3) {
                            s = (jbase + c + 1 <= m) ? smat[(size_t)(r - 1) * m + q[c]] * SC : 0;
                        } else { // sum over the chunk (columns past m read nothing)
                            s = 0;
                            if (jbase + c + 1 <= m) {
                                const uint8_t *pa = alpha + (size_t)(r - 1) * chunk, *pb = beta + q[c];
                                for (int u = 0; u < chunk; ++u)
                                    s += s_scores[(int)pa[u] * P.dim + (int)pb[u]];
                            }
                        }
                        const V Mc = hp + s; // M(r,j) = s + H(r-1,j-1)
                        V cI, cD, Ht, cH;
                        if (TRACE) {
                            cI = It & CLRV;
                            cD = Dt[c] & CLRV;
                            Ht = max3(Mc + 2 * FH, cI + FH, cD);
                            cH = Ht & CLRV;
                            const unsigned code = (unsigned)((It + Dt[c] + Ht) & (V)(kScale - 1));
                            w[c / 5] = (w[c / 5] << kTagBits) | code;
                        } else {
                            cI = It;
                            cD = Dt[c];
                            Ht = max3(Mc, cI, cD);
                            cH = Ht;
                        }
                        // I(r, j+1) = T(M+O+E, I+E, D+O+E)   (affineGap_highMem.go:213)
                        It = addmax(Mc, iM, addmax(cD, iD, cI + iI));
                        // D(r+1, j) = T(M+O+E, I+O+E, D+E)   (:214; FREE last column :209 has no penalty)
                        Dt[c] = addmax(Mc, aM[c], addmax(cI, aI[c], cD + aD[c]));
                        hp = Hc[c];
                        Hc[c] = cH;
                    }
                    edgeI = It;
                    edgeH = Hc[C - 1];
                    hpL = inH;
                    if (TRACE && tp) {
#pragma unroll
                        for (int k = 0; k < WPL; ++k)
                            tp[(size_t)k * 32] = w[k];
                    }
                    if (lane == 31 && p + 1 < strips)
                        eout[r] = EdgeT{edgeI, edgeH};
                }
                if (TRACE && tp)
                    tp += WPL * 32;
            }
            // ---- score: H(n, m) sits in Hc[] of the lane that owns column m ----------------------
            const int pm = (m - 1) / (32 * C), lm = ((m - 1) % (32 * C)) / C, cm = (m - 1) % C;
            if (p == pm && lane == lm) {
                V h = Hc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (c == cm)
                        h = Hc[c];
                P.out_score[pair] = (int64_t)(h / SC); // clean values are exact multiples of SC
            }
            __syncwarp(); // order the strip's edge writes before the next strip's reads
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Constant-gap (Needleman-Wunsch) fill.  Reference: align/constGap_highMem.go:23-40.
//   m(i,j), tr = T(m(i-1,j-1)+s, m(i,j-1)+g, m(i-1,j)+g); values carried as 4*v + tag.
// Trace: 2 bits per cell, 16 codes per 32-bit word.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int const_wpl(int C) { return (C + 15) / 16; }
__host__ __device__ inline int64_t const_trace_words(int64_t n, int64_t m, int C)
{
    if (n <= 0 || m <= 0)
        return 0;
    const int64_t strips = (m + 32 * C - 1) / (32 * C);
    return strips * (n + 31) * const_wpl(C) * 32;
}

template <int C, bool TRACE, int LOOKUP>
__global__ void __launch_bounds__(128) const_fill_kernel(const FillParams P)
{
    constexpr int SC = TRACE ? 4 : 1;
    constexpr int WPL = const_wpl(C);
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ int s_scores[64];
    if (LOOKUP == 1) {
        if (threadIdx.x < 64)
            s_scores[threadIdx.x] = P.scores[threadIdx.x] * SC;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int g = P.gap_open; // the single gap penalty
    const int g_left = g * SC + (TRACE ? 1 : 0), g_up = g * SC;
    int2 *edge_a = P.edge ? P.edge + (size_t)warp * 2 * P.edge_stride : nullptr;
    int2 *edge_b = P.edge ? edge_a + P.edge_stride : nullptr;

    for (int64_t pair = P.pair_begin + warp; pair < P.pair_end; pair += nwarps) {
        if (P.pair_class && P.pair_class[pair] != P.want_class)
            continue;
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0);
        const int m = (int)(P.beta_off[pair + 1] - b0);
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        if (n == 0 || m == 0) {
            if (lane == 0)
                P.out_score[pair] = (int64_t)g * (n + m); // boundary row/column (constGap_highMem.go:27-32)
            continue;
        }
        const int T = n + 31;
        const int strips = (m + 32 * C - 1) / (32 * C);
        uint32_t *tbase = TRACE ? P.trace + P.trace_off[pair - P.pair_begin] : nullptr;
        for (int p = 0; p < strips; ++p) {
            const int jbase = p * 32 * C + lane * C;
            int q[C], t01[C], t23[C], Hc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                q[c] = (j <= m) ? (int)beta[j - 1] : 0;
                if (LOOKUP == 0) {
                    const int s0 = P.scores[0 * P.dim + q[c]] * SC + (TRACE ? 2 : 0);
                    const int s1 = P.scores[1 * P.dim + q[c]] * SC + (TRACE ? 2 : 0);
                    const int s2 = P.scores[2 * P.dim + q[c]] * SC + (TRACE ? 2 : 0);
                    const int s3 = P.scores[3 * P.dim + q[c]] * SC + (TRACE ? 2 : 0);
                    t01[c] = (s0 & 0xffff) | (s1 << 16);
                    t23[c] = (s2 & 0xffff) | (s3 << 16);
                }
                Hc[c] = j * g * SC; // row 0
            }
            int hpL = jbase * g * SC;
            int edgeH = 0;
            const int2 *ein = (p & 1) ? edge_b : edge_a;
            int2 *eout = (p & 1) ? edge_a : edge_b;
            uint32_t *tp = TRACE ? tbase + ((size_t)p * T * WPL) * 32 + lane : nullptr;
            int bH = 0;
            auto boundary = [&](int r) { bH = (p == 0) ? r * g * SC : ein[r].x; };
            if (lane == 0)
                boundary(1);
            int a_next = (lane == 0) ? (int)alpha[0] : 0;
            for (int t = 0; t < T; ++t) {
                const int r = t - lane + 1;
                int inH = __shfl_up_sync(FULL, edgeH, 1);
                if (lane == 0)
                    inH = bH;
                const bool active = (r >= 1) && (r <= n);
                const int a = a_next;
                if (r + 1 >= 1 && r + 1 <= n)
                    a_next = alpha[r];
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    int sel = 0, rowoff = 0;
                    if (LOOKUP == 0)
                        sel = a * 0x2222 + 0x9910;
                    else
                        rowoff = a * P.dim;
                    int left = inH; // clean m(r, j-1)
                    int hp = hpL;   // clean m(r-1, j-1)
                    uint32_t w[WPL];
#pragma unroll
                    for (int k = 0; k < WPL; ++k)
                        w[k] = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        int s;
                        if (LOOKUP == 0)
                            s = prmt(t01[c], t23[c], sel);
                        else
                            s = s_scores[rowoff + q[c]] + (TRACE ? 2 : 0);
                        // candidates: diag+s (tag 2 = ColM), left+g (tag 1 = ColI), up+g (tag 0 = ColD)
                        const int ht = max3(hp + s, left + g_left, Hc[c] + g_up);
                        int cH = ht;
                        if (TRACE) {
                            cH = ht & ~3;
                            w[c / 16] = (w[c / 16] << 2) | ((unsigned)ht & 3u);
                        }
                        hp = Hc[c];
                        Hc[c] = cH;
                        left = cH;
                    }
                    edgeH = left;
                    hpL = inH;
                    if (TRACE) {
#pragma unroll
                        for (int k = 0; k < WPL; ++k)
                            tp[(size_t)k * 32] = w[k];
                    }
                    if (lane == 31 && p + 1 < strips)
                        eout[r] = make_int2(edgeH, 0);
                }
                if (TRACE)
                    tp += WPL * 32;
            }
            const int pm = (m - 1) / (32 * C), lm = ((m - 1) % (32 * C)) / C, cm = (m - 1) % C;
            if (p == pm && lane == lm) {
                int h = Hc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (c == cm)
                        h = Hc[c];
                P.out_score[pair] = (int64_t)(h / SC);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback + run-length encoding, one thread per pair.
// Reference: affineTrace (align/affineGap_highMem.go:57-89) and the cigar loop of
// ConstGap_highMem (align/constGap_highMem.go:43-65).
// Ops are produced end-to-start; they are kept in a small per-pair slot (reversed on expansion).
// ------------------------------------------------------------------------------------------------
struct TraceParams {
    const int64_t *alpha_off, *beta_off;
    int64_t pair_begin, pair_end;
    const uint32_t *trace;
    const int64_t *trace_off;
    int C;          // columns per lane the fill kernel used
    int layout;     // 1: affine_fill_kernel word layout (codes shifted in from the bottom, one word per step),
                    // 3: affine_fill3_kernel (codes shifted in from the top, rows blocked four steps per 16-byte piece)
    int lpp;        // lanes per pair of the fill kernel (32, or 16 for fill3's two-pairs-per-warp form)
    int skew;       // rows between neighbouring lanes (1, or 2 for fill3's pipelined form)
    int chunk;      // AffineGapChunk: bases per DP cell (run lengths are multiplied by it), else 1
    int kind;       // 0 affine, 2 const gap
    int h00_plane;  // plane of T(0, O, D(0,0)) (affine)
    uint32_t *slots; // per pair in chunk: slot_cap entries, run<<2 | op, traceback order
    int slot_cap;
    int *counts;    // per pair in chunk: number of cigar elements
    // second pass (overflowing pairs only): write final gnx_cigar entries directly
    int64_t *cigar_off;  // per pair in chunk (+1): exclusive scan of counts, relative to chunk base
    void *out_cigar;     // gnx_cigar* base that cigar_off indexes
    int64_t out_cap;     // entries available behind out_cigar
    int pass;            // 0: fill slots + counts, 1: rewrite pairs whose count > slot_cap
    const uint8_t *pair_class; // per global pair (may be NULL): class 2 = invalid base, the fill skipped it
};

struct CigarOut {
    long long run_length;
    unsigned char op;
};

__device__ __forceinline__ unsigned affine_code(const uint32_t *tr, int T, int C, int i, int j, int layout = 1,
                                                int lpp = 32, int skew = 1)
{
    const int wpl = trace_wpl(C);
    const int jj = j - 1;
    const int strip = jj / (lpp * C);
    const int within = jj - strip * lpp * C;
    const int lane = within / C, c = within - lane * C;
    const int t = (i - 1) + skew * lane;
    const int nin = (c / 5 == wpl - 1) ? (C - 5 * (wpl - 1)) : 5; // codes held by this word
    if (layout == 3) { // T is the padded (multiple of 4) row count; [t/4][word][thread][t%4]
        const uint32_t w3 = tr[((((size_t)strip * (T >> 2) + (t >> 2)) * wpl + c / 5) * 32 + lane) * 4 + (t & 3)];
        return (w3 >> (32 - kTagBits * (nin - (c % 5)))) & (kScale - 1);
    }
    const uint32_t w = tr[(((size_t)strip * T + t) * wpl + c / 5) * 32 + lane];
    if (layout == 2) // codes funnel-shifted in from the top (gnx_fill2.cuh)
        return (w >> (32 - kTagBits * (nin - (c % 5)))) & (kScale - 1);
    return (w >> (kTagBits * (nin - 1 - (c % 5)))) & (kScale - 1);
}

__device__ __forceinline__ unsigned const_code(const uint32_t *tr, int T, int C, int i, int j)
{
    const int wpl = const_wpl(C);
    const int jj = j - 1;
    const int strip = jj / (32 * C);
    const int within = jj - strip * 32 * C;
    const int lane = within / C, c = within - lane * C;
    const int t = (i - 1) + lane;
    const int nin = (c / 16 == wpl - 1) ? (C - 16 * (wpl - 1)) : 16;
    const uint32_t w = tr[(((size_t)strip * T + t) * wpl + c / 16) * 32 + lane];
    return (w >> (2 * (nin - 1 - (c % 16)))) & 3u;
}

__global__ void traceback_kernel(const TraceParams P)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) { // no trace was written for this pair (the call returns GNX_EBASE)
        if (P.pass == 0)
            P.counts[idx] = 0;
        return;
    }
    const int chunk = P.chunk > 1 ? P.chunk : 1;
    const int n = (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]) / chunk;
    const int m = (int)(P.beta_off[pair + 1] - P.beta_off[pair]) / chunk;
    if (P.pass == 1 && P.counts[idx] <= P.slot_cap)
        return;
    uint32_t *slot = P.slots + (size_t)idx * P.slot_cap;
    CigarOut *dst = nullptr;
    int total = 0;
    if (P.pass == 1) {
        total = P.counts[idx];
        if (P.cigar_off[idx] + total > P.out_cap)
            return; // expand_kernel has already flagged GNX_ECAP
        dst = (CigarOut *)P.out_cigar + P.cigar_off[idx];
    }
    int cnt = 0;
    auto emit = [&](int op, int run) {
        run *= chunk; // expandCigarRunLength (align/affineGap_highMem.go:91-95)
        if (P.pass == 0) {
            if (cnt < P.slot_cap)
                slot[cnt] = ((uint32_t)run << 2) | (uint32_t)op;
        } else { // final order is reversed traceback order
            CigarOut o;
            o.run_length = run;
            o.op = (unsigned char)op;
            dst[total - 1 - cnt] = o;
        }
        ++cnt;
    };
    if (n == 0 && m == 0) { // route := make([]Cigar, 1): one zero element (affineGap_highMem.go:58)
        emit(0, 0);
        if (P.pass == 0)
            P.counts[idx] = cnt;
        return;
    }
    const uint32_t *tr = P.trace + P.trace_off[idx];
    int T = n + (P.kind == 0 ? P.skew * (P.lpp - 1) : 31);
    if (P.kind == 0 && P.layout == 3)
        T = (T + 3) & ~3;
    const int C = P.C;
    int i = n, j = m, cur = -1, run = 0;
    if (P.kind == 0) {
        // start plane: T(M,I,D)(n,m) = the H tag of cell (n,m); boundaries are closed-form
        int k;
        if (n == 0)
            k = 1;
        else if (m == 0)
            k = 2;
        else
            k = 2 - (int)((affine_code(tr, T, C, n, m, P.layout, P.lpp, P.skew) >> 4) & 3u);
        while (i > 0 || j > 0) {
            if (k == cur) {
                ++run;
            } else {
                if (cur >= 0)
                    emit(cur, run);
                cur = k;
                run = 1;
            }
            if (i == 0) { // row 0: trace[1][0][j] = ColI (:196)
                --j;
                k = 1;
            } else if (j == 0) { // column 0: trace[2][i][0] = ColD (:205)
                --i;
                k = 2;
            } else if (k == 0) { // M: the source plane is argmax(M,I,D) of the diagonal neighbour
                --i;
                --j;
                if (i == 0 && j == 0)
                    k = P.h00_plane;
                else if (i == 0)
                    k = 1;
                else if (j == 0)
                    k = 2;
                else
                    k = 2 - (int)((affine_code(tr, T, C, i, j, P.layout, P.lpp, P.skew) >> 4) & 3u);
            } else if (k == 1) {
                k = 2 - (int)(affine_code(tr, T, C, i, j, P.layout, P.lpp, P.skew) & 3u);
                --j;
            } else {
                k = 2 - (int)((affine_code(tr, T, C, i, j, P.layout, P.lpp, P.skew) >> 2) & 3u);
                --i;
            }
        }
    } else {
        while (i > 0 || j > 0) {
            int k;
            if (i == 0)
                k = 1; // trace[0][j] = 1 (constGap_highMem.go:28)
            else if (j == 0)
                k = 2; // trace[i][0] = 2 (:31)
            else
                k = 2 - (int)const_code(tr, T, C, i, j);
            if (k == cur) {
                ++run;
            } else {
                if (cur >= 0)
                    emit(cur, run);
                cur = k;
                run = 1;
            }
            if (k == 0) {
                --i;
                --j;
            } else if (k == 1) {
                --j;
            } else {
                --i;
            }
        }
    }
    if (cur >= 0)
        emit(cur, run);
    if (P.pass == 0)
        P.counts[idx] = cnt;
}

// Branch-light traceback for the affine kernels' packed traces (layouts 2 and 3, C = 5 or 10).
// One thread per pair, exactly one trace load per step, the same instruction sequence for every
// thread whatever plane it is in (the generic kernel above diverges three ways per step; ncu showed
// 12.7 of 32 lanes active).  The cell's (strip, lane, column) coordinates are walked incrementally
// instead of re-derived with integer divisions.  Boundary cells are given pseudo-codes:
//   (0,0): H tag = plane of T(0,O,D(0,0));  (0,j>0): every tag = I;  (i>0,0): every tag = D.
__global__ void traceback_affine_kernel(const TraceParams P)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) { // no trace was written for this pair (the call returns GNX_EBASE)
        if (P.pass == 0)
            P.counts[idx] = 0;
        return;
    }
    const int n = (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]);
    const int m = (int)(P.beta_off[pair + 1] - P.beta_off[pair]);
    if (P.pass == 1 && P.counts[idx] <= P.slot_cap)
        return;
    uint32_t *slot = P.slots + (size_t)idx * P.slot_cap;
    CigarOut *dst = nullptr;
    int total = 0;
    if (P.pass == 1) {
        total = P.counts[idx];
        if (P.cigar_off[idx] + total > P.out_cap)
            return;
        dst = (CigarOut *)P.out_cigar + P.cigar_off[idx];
    }
    int cnt = 0;
    auto emit = [&](int op, int run) {
        if (P.pass == 0) {
            if (cnt < P.slot_cap)
                slot[cnt] = ((uint32_t)run << 2) | (uint32_t)op;
        } else {
            CigarOut o;
            o.run_length = run;
            o.op = (unsigned char)op;
            dst[total - 1 - cnt] = o;
        }
        ++cnt;
    };
    if (n == 0 && m == 0) {
        emit(0, 0);
        if (P.pass == 0)
            P.counts[idx] = cnt;
        return;
    }
    const uint32_t *__restrict__ tr = P.trace + P.trace_off[idx];
    const int C = P.C, lpp = P.lpp, skew = P.skew, wpl = trace_wpl(C);
    const bool blocked = P.layout == 3;
    int T = n + skew * (lpp - 1);
    if (blocked)
        T = (T + 3) & ~3;
    const size_t strip_words = (size_t)T * wpl * 32;
    // coordinates of column j = m
    int strip = 0, lane = 0, c = 0;
    if (m > 0) {
        const int jj = m - 1;
        strip = jj / (lpp * C);
        const int within = jj - strip * lpp * C;
        lane = within / C;
        c = within - lane * C;
    }
    auto load = [&](int i) -> unsigned { // code of cell (i, j) whose column coordinates are (strip, lane, c)
        const int t = (i - 1) + skew * lane;
        const int wi = c >= 5 ? 1 : 0, cc = c - 5 * wi;
        const int nin = (wi == wpl - 1) ? (C - 5 * (wpl - 1)) : 5;
        size_t a;
        if (blocked)
            a = (size_t)strip * strip_words + ((((size_t)(t >> 2)) * wpl + wi) * 32 + lane) * 4 + (t & 3);
        else
            a = (size_t)strip * strip_words + (((size_t)t) * wpl + wi) * 32 + lane;
        return (__ldg(tr + a) >> (32 - kTagBits * (nin - cc))) & (kScale - 1);
    };
    const unsigned code00 = (unsigned)(2 - P.h00_plane) << 4;
    int i = n, j = m;
    unsigned cur = (i > 0 && j > 0) ? load(i) : (i == 0 ? 0x15u : 0x00u);
    int k = 2 - (int)((cur >> 4) & 3u);
    int run = 0, cur_op = k;
    while (i > 0 || j > 0) {
        // on the boundary the pseudo-codes point to themselves: column 0 stays in plane D up to (0,0), row 0 in
        // plane I -- the rest of the route is one run (freeEndGaps alignments end with hundreds of such cells)
        if ((j == 0 && k == 2) || (i == 0 && k == 1)) {
            const int len = j == 0 ? i : j;
            if (k == cur_op) {
                run += len;
            } else {
                emit(cur_op, run);
                cur_op = k;
                run = len;
            }
            break;
        }
        if (k == cur_op) {
            ++run;
        } else {
            emit(cur_op, run);
            cur_op = k;
            run = 1;
        }
        // next plane when the current one is I or D: its tag in the CURRENT cell's code
        const int kn = 2 - (int)((cur >> (k == 1 ? 0 : 2)) & 3u);
        const int di = k != 1, dj = k != 2;
        i -= di;
        if (dj) { // step one column to the left
            --j;
            if (--c < 0) {
                c = C - 1;
                if (--lane < 0) {
                    lane = lpp - 1;
                    --strip;
                }
            }
        }
        unsigned nw;
        if (i > 0 && j > 0)
            nw = load(i);
        else
            nw = (i == 0) ? (j == 0 ? code00 : 0x15u) : 0x00u;
        k = (k == 0) ? 2 - (int)((nw >> 4) & 3u) : kn; // M: argmax(M,I,D) of the diagonal neighbour
        cur = nw;
    }
    emit(cur_op, run);
    if (P.pass == 0)
        P.counts[idx] = cnt;
}

// Warp-per-pair form of traceback_affine_kernel for LONG pairs (layout 3).  A thread-per-pair walk is a chain
// of ~n+m dependent L2/HBM loads (20 000 for a 10 kb x 10 kb pair: ~18 ms per chunk whatever the pair count).
// Here every lane of the warp keeps the same walk state; while the path is in plane M, lane d-1 loads the code
// of the diagonal cell (i-d, j-d) -- 32 independent loads in flight -- and the run of leading cells whose H tag
// is M is consumed in one iteration (the code the walk lands on comes from the lane that loaded it), so a
// match run of L cells costs L/32 round trips.  I / D steps take one broadcast load each.  Same RLE, same
// boundary pseudo-codes, same two-pass protocol as traceback_affine_kernel.
__global__ void __launch_bounds__(128) traceback_affine_warp_kernel(const TraceParams P)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane_id = threadIdx.x & 31;
    const int64_t idx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) {
        if (P.pass == 0 && lane_id == 0)
            P.counts[idx] = 0;
        return;
    }
    const int n = (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]);
    const int m = (int)(P.beta_off[pair + 1] - P.beta_off[pair]);
    if (P.pass == 1 && P.counts[idx] <= P.slot_cap)
        return;
    uint32_t *slot = P.slots + (size_t)idx * P.slot_cap;
    CigarOut *dst = nullptr;
    int total = 0;
    if (P.pass == 1) {
        total = P.counts[idx];
        if (P.cigar_off[idx] + total > P.out_cap)
            return;
        dst = (CigarOut *)P.out_cigar + P.cigar_off[idx];
    }
    int cnt = 0;
    auto emit = [&](int op, int run) { // every lane counts, lane 0 stores
        if (lane_id == 0) {
            if (P.pass == 0) {
                if (cnt < P.slot_cap)
                    slot[cnt] = ((uint32_t)run << 2) | (uint32_t)op;
            } else {
                CigarOut o;
                o.run_length = run;
                o.op = (unsigned char)op;
                dst[total - 1 - cnt] = o;
            }
        }
        ++cnt;
    };
    if (n == 0 && m == 0) {
        emit(0, 0);
        if (P.pass == 0 && lane_id == 0)
            P.counts[idx] = cnt;
        return;
    }
    const uint32_t *__restrict__ tr = P.trace + P.trace_off[idx];
    const int C = P.C, lpp = P.lpp, skew = P.skew, wpl = trace_wpl(C);
    int T = (n + skew * (lpp - 1) + 3) & ~3; // layout 3: rows blocked four steps per 16-byte piece
    const size_t strip_words = (size_t)T * wpl * 32;
    const int strip_cols = lpp * C;
    auto load = [&](int i, int j) -> unsigned { // code of interior cell (i, j)
        const int jj = j - 1;
        const int strip = jj / strip_cols, within = jj - strip * strip_cols;
        const int lane = within / C, c = within - lane * C;
        const int t = (i - 1) + skew * lane;
        const int wi = c >= 5 ? 1 : 0, cc = c - 5 * wi;
        const int nin = (wi == wpl - 1) ? (C - 5 * (wpl - 1)) : 5;
        const size_t a = (size_t)strip * strip_words + ((((size_t)(t >> 2)) * wpl + wi) * 32 + lane) * 4 + (t & 3);
        return (__ldg(tr + a) >> (32 - kTagBits * (nin - cc))) & (kScale - 1);
    };
    const unsigned code00 = (unsigned)(2 - P.h00_plane) << 4;
    auto code_at = [&](int i, int j) -> unsigned {
        return (i > 0 && j > 0) ? load(i, j) : ((i == 0) ? (j == 0 ? code00 : 0x15u) : 0x00u);
    };
    int i = n, j = m;
    unsigned cur = code_at(i, j);
    int k = 2 - (int)((cur >> 4) & 3u);
    int run = 0, cur_op = k;
    while (i > 0 || j > 0) {
        if ((j == 0 && k == 2) || (i == 0 && k == 1)) { // boundary: the rest is one run
            const int len = j == 0 ? i : j;
            if (k == cur_op) {
                run += len;
            } else {
                emit(cur_op, run);
                cur_op = k;
                run = len;
            }
            break;
        }
        if (k == 0) {
            // look ahead along the diagonal: lane d-1 takes cell (i-d, j-d)
            const int d = lane_id + 1, ci = i - d, cj = j - d;
            const bool inside = ci > 0 && cj > 0;
            const unsigned cd = inside ? load(ci, cj) : 0u;
            const bool isM = inside && ((cd >> 4) & 3u) == 2u;
            const int skip = __ffs(~__ballot_sync(FULL, isM)) - 1; // 0..32 leading diagonal cells in plane M
            if (skip >= 1) {
                // the current cell and the next skip-1 cells are M; land on cell d = skip (plane M, code known)
                if (cur_op == 0) {
                    run += skip;
                } else {
                    emit(cur_op, run);
                    cur_op = 0;
                    run = skip;
                }
                i -= skip;
                j -= skip;
                cur = __shfl_sync(FULL, cd, skip - 1);
                continue;
            }
            // skip == 0: the diagonal neighbour is not an interior M cell -- one ordinary step, its code is in lane 0
            if (cur_op == 0) {
                ++run;
            } else {
                emit(cur_op, run);
                cur_op = 0;
                run = 1;
            }
            --i;
            --j;
            const unsigned c1 = __shfl_sync(FULL, cd, 0);
            const unsigned nw = (i > 0 && j > 0) ? c1 : ((i == 0) ? (j == 0 ? code00 : 0x15u) : 0x00u);
            k = 2 - (int)((nw >> 4) & 3u);
            cur = nw;
            continue;
        }
        // plane I or D: one step
        if (k == cur_op) {
            ++run;
        } else {
            emit(cur_op, run);
            cur_op = k;
            run = 1;
        }
        const int kn = 2 - (int)((cur >> (k == 1 ? 0 : 2)) & 3u);
        i -= (k != 1);
        j -= (k != 2);
        cur = code_at(i, j);
        k = kn;
    }
    emit(cur_op, run);
    if (P.pass == 0 && lane_id == 0)
        P.counts[idx] = cnt;
}

// Traceback of const_fill3_kernel traces (2-bit codes, 10 per word, rows blocked by four): one load per
// step, no plane state (the code IS the op).  Reference: constGap_highMem.go:43-65.
__global__ void traceback_const3_kernel(const TraceParams P)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) { // no trace was written for this pair (the call returns GNX_EBASE)
        if (P.pass == 0)
            P.counts[idx] = 0;
        return;
    }
    const int n = (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]);
    const int m = (int)(P.beta_off[pair + 1] - P.beta_off[pair]);
    if (P.pass == 1 && P.counts[idx] <= P.slot_cap)
        return;
    uint32_t *slot = P.slots + (size_t)idx * P.slot_cap;
    CigarOut *dst = nullptr;
    int total = 0;
    if (P.pass == 1) {
        total = P.counts[idx];
        if (P.cigar_off[idx] + total > P.out_cap)
            return;
        dst = (CigarOut *)P.out_cigar + P.cigar_off[idx];
    }
    int cnt = 0;
    auto emit = [&](int op, int run) {
        if (P.pass == 0) {
            if (cnt < P.slot_cap)
                slot[cnt] = ((uint32_t)run << 2) | (uint32_t)op;
        } else {
            CigarOut o;
            o.run_length = run;
            o.op = (unsigned char)op;
            dst[total - 1 - cnt] = o;
        }
        ++cnt;
    };
    if (n == 0 && m == 0) {
        emit(0, 0);
        if (P.pass == 0)
            P.counts[idx] = cnt;
        return;
    }
    const uint32_t *__restrict__ tr = P.trace + P.trace_off[idx];
    const int C = P.C, lpp = P.lpp;
    const int T = (n + lpp - 1 + 3) & ~3;
    const size_t strip_words = (size_t)T * 32;
    int strip = 0, lane = 0, c = 0;
    if (m > 0) {
        const int jj = m - 1;
        strip = jj / (lpp * C);
        const int within = jj - strip * lpp * C;
        lane = within / C;
        c = within - lane * C;
    }
    int i = n, j = m, cur_op = -1, run = 0;
    while (i > 0 || j > 0) {
        int k;
        if (i > 0 && j > 0) {
            const int t = (i - 1) + lane;
            const uint32_t w = __ldg(tr + (size_t)strip * strip_words + (((size_t)(t >> 2)) * 32 + lane) * 4 + (t & 3));
            k = 2 - (int)((w >> (32 - 2 * (C - c))) & 3u);
        } else {
            k = (i == 0) ? 1 : 2; // row 0: I, column 0: D (constGap_highMem.go:28,31)
        }
        if (k == cur_op) {
            ++run;
        } else {
            if (cur_op >= 0)
                emit(cur_op, run);
            cur_op = k;
            run = 1;
        }
        i -= (k != 1);
        if (k != 2) {
            --j;
            if (--c < 0) {
                c = C - 1;
                if (--lane < 0) {
                    lane = lpp - 1;
                    --strip;
                }
            }
        }
    }
    emit(cur_op, run);
    if (P.pass == 0)
        P.counts[idx] = cnt;
}

// Traceback of the gsw extend step (genomeGraph/search.go:252-272 left, :301-319 right) over
// const_fill3_kernel<EXT> traces.  The route stays in TRACEBACK order (the reference does not reverse it
// here) and ops are the cigar package's bytes 'M','I','D' (cigar/cigar.go:15-18); expand_kernel is told so.
//   left : from (n,m) while the cell's value is > 0.  Values are not stored: a cell with value > 0 was not
//          clipped, so m(prev) = m(cur) - (substitution score | gap penalty) exactly; the walk starts from
//          the pair's score m(n,m) and recomputes the value as it goes.  Returns where it stopped.
//   right: from the first maximal cell (out_best) to (0,0); row 0 is 'I', column 0 is 'D' (:284-289).
struct ExtTraceParams {
    TraceParams t;
    int side;                 // 1 left, 2 right
    const uint8_t *alpha, *beta;
    int dim, gap;
    int scores[64];
    const int64_t *score;     // per global pair (left: the starting value)
    const int64_t *best;      // per global pair (right: row << 32 | column)
    int64_t *end_i, *end_j;   // per pair in chunk
    int local;                // LeftLocal / RightLocal (genomeGraph/localAlignment.go:95-196): the same DPs with
                              // cigar.TripleMaxTraceExtended -- a diagonal step is '=' when the substitution score is
                              // positive (a > prev) and 'X' otherwise -- and the route reversed into alignment order
};

__global__ void traceback_ext_kernel(const ExtTraceParams E)
{
    const TraceParams &P = E.t;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) { // no trace was written for this pair (the call returns GNX_EBASE)
        if (P.pass == 0)
            P.counts[idx] = 0;
        return;
    }
    const int n = (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]);
    const int m = (int)(P.beta_off[pair + 1] - P.beta_off[pair]);
    if (P.pass == 1 && P.counts[idx] <= P.slot_cap)
        return;
    uint32_t *slot = P.slots + (size_t)idx * P.slot_cap;
    CigarOut *dst = nullptr;
    if (P.pass == 1) {
        if (P.cigar_off[idx] + P.counts[idx] > P.out_cap)
            return;
        dst = (CigarOut *)P.out_cigar + P.cigar_off[idx];
    }
    int cnt = 0;
    auto emit = [&](int op, int run) {
        if (P.pass == 0) {
            if (cnt < P.slot_cap)
                slot[cnt] = ((uint32_t)run << 2) | (uint32_t)op;
        } else {
            CigarOut o;
            o.run_length = run;
            if (E.local) {
                o.op = (unsigned char)(op == 0 ? '=' : (op == 1 ? 'I' : (op == 2 ? 'D' : 'X')));
                dst[P.counts[idx] - 1 - cnt] = o; // cigar.ReverseCigar
            } else {
                o.op = (unsigned char)(op == 0 ? 'M' : (op == 1 ? 'I' : 'D'));
                dst[cnt] = o; // traceback order is the final order
            }
        }
        ++cnt;
    };
    int i = n, j = m;
    long long v = 0;
    if (E.side == 1) {
        v = (n > 0 && m > 0) ? E.score[pair] : 0;
    } else {
        const long long b = (n > 0 && m > 0) ? E.best[pair] : 0;
        i = (int)(b >> 32);
        j = (int)(b & 0xffffffffll);
    }
    const int start_i = i, start_j = j;
    const uint32_t *__restrict__ tr = P.trace + P.trace_off[idx];
    const uint8_t *__restrict__ al = E.alpha + P.alpha_off[pair];
    const uint8_t *__restrict__ be = E.beta + P.beta_off[pair];
    const int C = P.C, lpp = P.lpp;
    const int T = (n + lpp - 1 + 3) & ~3;
    const size_t strip_words = (size_t)T * 32;
    int strip = 0, lane = 0, c = 0;
    if (j > 0) {
        const int jj = j - 1;
        strip = jj / (lpp * C);
        const int within = jj - strip * lpp * C;
        lane = within / C;
        c = within - lane * C;
    }
    int cur_op = -1, run = 0;
    int guard = n + m + 1; // a valid walk takes at most n + m steps: a corrupt trace must not spin forever
    while ((E.side == 1 ? v > 0 : (i > 0 || j > 0)) && guard-- > 0) {
        int k;
        if (i > 0 && j > 0) {
            const int t = (i - 1) + lane;
            const uint32_t w = __ldg(tr + (size_t)strip * strip_words + (((size_t)(t >> 2)) * 32 + lane) * 4 + (t & 3));
            k = 2 - (int)((w >> (32 - 2 * (C - c))) & 3u);
        } else {
            k = (i == 0) ? 1 : 2;
        }
        int sub = 0;
        if (k == 0 && (E.side == 1 || E.local))
            sub = E.scores[(int)al[i - 1] * E.dim + (int)be[j - 1]];
        const int kop = (E.local && k == 0 && sub <= 0) ? 3 : k; // run-length op: 3 = 'X' (TripleMaxTraceExtended: a <= prev)
        if (kop == cur_op) {
            ++run;
        } else {
            if (cur_op >= 0)
                emit(cur_op, run);
            cur_op = kop;
            run = 1;
        }
        if (E.side == 1)
            v -= (k == 0) ? (long long)sub : (long long)E.gap;
        i -= (k != 1);
        if (k != 2) {
            --j;
            if (--c < 0) {
                c = C - 1;
                if (--lane < 0) {
                    lane = lpp - 1;
                    --strip;
                }
            }
        }
    }
    if (cur_op >= 0)
        emit(cur_op, run);
    if (P.pass == 0) {
        P.counts[idx] = cnt;
        E.end_i[idx] = E.side == 1 ? i : start_i;
        E.end_j[idx] = E.side == 1 ? j : start_j;
    }
}

// Expand the per-pair slots into gnx_cigar records at the scanned offsets (reversing to start->end
// order, align/align.go:86-90 reverseCigar).  One thread per pair; cigars are short.
// ext != 0 (gsw extend step): keep the traceback order and write the ops as 'M','I','D'.
__global__ void expand_kernel(const uint32_t *slots, int slot_cap, const int *counts, const int64_t *cigar_off,
                              int64_t n_pairs, CigarOut *out, int64_t out_cap_remaining, int *status, int ext)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs)
        return;
    const int cnt = counts[idx];
    const int64_t off = cigar_off[idx];
    if (off + cnt > out_cap_remaining) {
        atomicMax(status, kECap);
        return;
    }
    if (cnt > slot_cap)
        return; // rewritten by traceback pass 1
    const uint32_t *slot = slots + (size_t)idx * slot_cap;
    for (int k = 0; k < cnt; ++k) {
        const uint32_t v = slot[k];
        CigarOut o;
        o.run_length = (long long)(v >> 2);
        o.op = (unsigned char)(v & 3u);
        if (ext == 2) { // LeftLocal / RightLocal: extended ops, alignment order
            o.op = (unsigned char)(o.op == 0 ? '=' : (o.op == 1 ? 'I' : (o.op == 2 ? 'D' : 'X')));
            out[off + cnt - 1 - k] = o;
        } else if (ext) {
            o.op = (unsigned char)(o.op == 0 ? 'M' : (o.op == 1 ? 'I' : 'D'));
            out[off + k] = o;
        } else {
            out[off + cnt - 1 - k] = o;
        }
    }
}

// Exclusive scan of int counts into int64 offsets (n+1 entries) in two launches:
//   scan_partial_kernel: each 256-thread block sums its kScanSeg-pair segment into partial[block];
//   scan_apply_kernel:   each block adds the partials before it (a few hundred values) to *running_total
//                        and scans its own segment; the last block publishes off[n] and the new total.
constexpr int kScanSeg = 2048;

__global__ void scan_partial_kernel(const int *counts, int64_t n, int64_t *partial)
{
    __shared__ int64_t s_sum[8];
    const int64_t lo = (int64_t)blockIdx.x * kScanSeg;
    int64_t sum = 0;
    for (int64_t i = lo + threadIdx.x; i < min(n, lo + kScanSeg); i += blockDim.x)
        sum += counts[i];
    for (int o = 16; o > 0; o >>= 1)
        sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0)
        s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            t += s_sum[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void scan_apply_kernel(const int *counts, int64_t n, const int64_t *partial, int64_t *off,
                                  const int64_t *running_total, int64_t *total_out)
{
    __shared__ int64_t s_base;
    __shared__ int64_t s_warp[8];
    // base of this block = running total + sum of the partials of the blocks before it
    int64_t b = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += blockDim.x)
        b += partial[k];
    for (int o = 16; o > 0; o >>= 1)
        b += __shfl_down_sync(0xffffffffu, b, o);
    if ((threadIdx.x & 31) == 0)
        s_warp[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = *running_total;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
            t += s_warp[w];
        s_base = t;
    }
    __syncthreads();
    // each thread scans kScanSeg / blockDim consecutive counts
    const int per = kScanSeg / 256;
    const int64_t lo = (int64_t)blockIdx.x * kScanSeg + (int64_t)threadIdx.x * per;
    int64_t mine = 0;
    for (int k = 0; k < per; ++k)
        if (lo + k < n)
            mine += counts[lo + k];
    // exclusive scan of `mine` across the block
    int64_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o)
            incl += v;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 31)
        s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int64_t wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w)
        wbase += s_warp[w];
    int64_t run = s_base + wbase + incl - mine;
    for (int k = 0; k < per; ++k) {
        if (lo + k < n) {
            off[lo + k] = run;
            run += counts[lo + k];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) { // holds the grand total
        off[n] = run;
        if (total_out)
            *total_out = run;
    }
}

// Legacy single-block scan (kept for reference / tiny inputs): n+1 entries, offsets continue from
// *running_total.
// Exclusive scan of int counts into int64 offsets (n+1 entries), single block, chunk-sized inputs.
// Chunks are at most a few hundred thousand pairs, so one 1024-thread block striding is enough.
__global__ void scan_counts_kernel(const int *counts, int64_t n, int64_t *off, int64_t *running_total)
{
    __shared__ int64_t s_part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (n + blockDim.x - 1) / blockDim.x;
    const int64_t lo = min(n, (int64_t)tid * per), hi = min(n, lo + per);
    int64_t sum = 0;
    for (int64_t i = lo; i < hi; ++i)
        sum += counts[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        int64_t run = *running_total; // offsets continue from the chunks scanned before
        for (int k = 0; k < (int)blockDim.x; ++k) {
            const int64_t v = s_part[k];
            s_part[k] = run;
            run += v;
        }
        off[n] = run;
        *running_total = run;
    }
    __syncthreads();
    int64_t run = s_part[tid];
    for (int64_t i = lo; i < hi; ++i) {
        off[i] = run;
        run += counts[i];
    }
}

} // namespace gnx

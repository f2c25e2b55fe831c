// gnx_fill2.cuh -- second-generation affine fill kernel (same mapping and bit-exact contract as
// affine_fill_kernel in gnx_kernels.cuh; see DESIGN.md "Kernels").
//
// What changed, and why (profiles/r01_microbench_pipes.txt, profiles/r01a_fill.md):
//   * every integer max / logic / permute instruction issues on the ALU pipe (64 lanes/clk/SM) while
//     IMAD issues on the FMA pipe (64 lanes/clk/SM) in parallel; v1 was ALU-bound (81 % vs 31 %).  The
//     cell update is rewritten as "plain adds + one VIMNMX3 per plane" so ptxas can place the adds on
//     the FMA pipe, the per-cell ALU work being: PRMT, 3 tag-clears, 3 maxes, XOR3 + funnel shift.
//   * one warp per CTA: pair index, lengths, strip counts and base pointers are CTA-uniform, so loop
//     control and address bases live in the uniform datapath instead of per-thread registers.
//   * the step loop is split into ramp-up / steady / ramp-down: the steady phase (all lanes on a valid
//     row) carries no activity predicate and no divergent branch.
//   * score-only uses H directly:  I' = max(I+E, H+O+E),  D' = max(D+E, H+O+E)  (valid for O <= 0,
//     otherwise the host dispatches the tagged kernel with stores disabled).
#pragma once
#include "gnx_kernels.cuh"
#include <type_traits>

namespace gnx {

__device__ __forceinline__ unsigned shf_r_wrap(unsigned lo, unsigned hi, unsigned n)
{
    unsigned d;
    asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(n));
    return d;
}
__device__ __forceinline__ int xor3(int a, int b, int c)
{
    int d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// a + b issued as IMAD a, one, b: the FMA pipe runs in parallel with the ALU pipe that all the integer
// max / logic ops share.  `one` is a kernel parameter so ptxas cannot fold the multiply and re-fuse the
// add into a VIADDMNMX (which would put it back on the saturated ALU pipe).
__device__ __forceinline__ int madd(int a, int one, int b)
{
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

// Trace word layout of fill2 (per lane, per step): codes are funnel-shifted in from the top, so with
// `nin` codes in a word, code k sits at bit 32 - 6*(nin - k).
__device__ __forceinline__ unsigned affine_code2(const uint32_t *tr, int T, int C, int i, int j)
{
    const int wpl = trace_wpl(C);
    const int jj = j - 1;
    const int strip = jj / (32 * C);
    const int within = jj - strip * 32 * C;
    const int lane = within / C, c = within - lane * C;
    const int t = (i - 1) + lane;
    const int nin = (c / 5 == wpl - 1) ? (C - 5 * (wpl - 1)) : 5;
    const uint32_t w = tr[(((size_t)strip * T + t) * wpl + c / 5) * 32 + lane];
    return (w >> (32 - kTagBits * (nin - (c % 5)))) & (kScale - 1);
}

//   STORE  write the traceback words (TRACE && !STORE: tagged arithmetic only -- score only with O > 0)
//   MULTI  pairs may span several strips (edge hand-off buffers); false = every pair fits one strip
template <int C, bool TRACE, bool STORE, bool FREE, int LOOKUP, bool MULTI>
__global__ void __launch_bounds__(32, 20) affine_fill2_kernel(const FillParams P)
{
    constexpr int SC = TRACE ? kScale : 1;
    constexpr int FI = TRACE ? kFI : 0, FD = TRACE ? kFD : 0, FH = TRACE ? kFH : 0;
    constexpr int WPL = trace_wpl(C);
    constexpr int NEG = kNeg32;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int CLR = ~(kScale - 1);

    __shared__ int s_scores[64];
    if (LOOKUP == 1) {
        s_scores[threadIdx.x] = P.scores[threadIdx.x] * SC + 2 * FH;
        s_scores[threadIdx.x + 32] = P.scores[threadIdx.x + 32] * SC + 2 * FH;
        __syncwarp();
    }
    const int lane = threadIdx.x;
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_s = (O + E) * SC, e_s = E * SC;
    // candidates relative to MH = M + 2*FH (the table already carries the H tag of the M candidate)
    const int kI = oe_s + 2 * FI - 2 * FH; // MH + kI = M + O + E + tag(M) in the I field
    const int iI = e_s + FI, iD = oe_s;
    // D-plane addends (regular column / freeEndGaps last column).  In the TRACE form the I candidate
    // is built from cIh = I + FH, so its addends carry -FH.
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;

    int2 *edge_a = MULTI ? P.edge + (size_t)blockIdx.x * 2 * P.edge_stride : nullptr;
    int2 *edge_b = MULTI ? edge_a + P.edge_stride : nullptr;
    const int fh_reg = FH * one; // register copy of the H tag of the I candidate (LOP3 operand)

    for (int64_t pair = P.pair_begin + blockIdx.x; pair < P.pair_end; pair += gridDim.x) {
        if (P.pair_class && P.pair_class[pair] != P.want_class)
            continue;
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0);
        const int m = (int)(P.beta_off[pair + 1] - b0);
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        if (n == 0 || m == 0) {
            if (lane == 0) {
                int64_t sc;
                if (n == 0 && m == 0)
                    sc = P.h00;
                else if (n == 0)
                    sc = (int64_t)O + (int64_t)m * E;
                else
                    sc = FREE ? 0 : (int64_t)O + (int64_t)n * E;
                P.out_score[pair] = sc;
            }
            continue;
        }
        const int T = n + 31;
        const int strips = MULTI ? (m + 32 * C - 1) / (32 * C) : 1;
        uint32_t *tbase = STORE ? P.trace + P.trace_off[pair - P.pair_begin] : nullptr;

        for (int p = 0; p < strips; ++p) {
            const int jbase = p * 32 * C + lane * C;
            int q[C], t01[C], t23[C];
            int aM[C], aI[C], aD[C]; // only materialised for FREE (constants otherwise)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                q[c] = (j <= m) ? (int)beta[j - 1] : 0;
                if (LOOKUP == 0) {
                    const int s0 = P.scores[0 * P.dim + q[c]] * SC + 2 * FH, s1 = P.scores[1 * P.dim + q[c]] * SC + 2 * FH;
                    const int s2 = P.scores[2 * P.dim + q[c]] * SC + 2 * FH, s3 = P.scores[3 * P.dim + q[c]] * SC + 2 * FH;
                    t01[c] = (s0 & 0xffff) | (s1 << 16);
                    t23[c] = (s2 & 0xffff) | (s3 << 16);
                }
                const bool last = FREE && (j == m);
                aM[c] = last ? dMl : dMn;
                aI[c] = last ? dIl : dIn;
                aD[c] = last ? dDl : dDn;
            }
            int Dt[C], Hc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const int i0 = (O + j * E) * SC;
                Hc[c] = i0;
                // D(1,j) = T(M(0,j)+.., I(0,j)+.., D(0,j)+..) with M = D = -inf on row 0
                Dt[c] = max3(NEG + 2 * FH + aM[c], i0 + FH + aI[c], NEG + aD[c]);
            }
            int hpL = (jbase == 0) ? P.h00 * SC : (O + jbase * E) * SC;
            int edgeI = 0, edgeH = 0;
            const int2 *ein = (p & 1) ? edge_b : edge_a;
            int2 *eout = (p & 1) ? edge_a : edge_b;
            uint32_t *tp = STORE ? tbase + ((size_t)p * T * WPL) * 32 + lane : nullptr;
            const bool store_edge = MULTI && (lane == 31) && (p + 1 < strips);

            int bI = 0, bH = 0; // lane 0: boundary stream for its next row
            auto boundary = [&](int r) {
                if (!MULTI || p == 0) {
                    const int d0 = FREE ? 0 : (O + r * E) * SC;
                    bI = d0 + iD;
                    bH = d0;
                } else {
                    const int2 v = ein[r];
                    bI = v.x;
                    bH = v.y;
                }
            };
            if (lane == 0)
                boundary(1);
            int a_next = (lane == 0) ? (int)alpha[0] : 0;

            // one wavefront step; CHECK = lanes may be off the matrix (ramp phases)
            auto step = [&](int t, auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int r = t - lane + 1;
                int inI = __shfl_up_sync(FULL, edgeI, 1);
                int inH = __shfl_up_sync(FULL, edgeH, 1);
                if (lane == 0) {
                    inI = bI;
                    inH = bH;
                }
                const int a = a_next;
                bool active = true;
                if (CHECK) {
                    active = (unsigned)(r - 1) < (unsigned)n;
                    if ((unsigned)r < (unsigned)n)
                        a_next = alpha[r];
                } else {
                    a_next = alpha[r]; // r <= n - 1 throughout the steady phase
                }
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    int sel = 0, rowoff = 0;
                    if (LOOKUP == 0)
                        sel = a * 0x2222 + 0x9910;
                    else
                        rowoff = a * P.dim;
                    int It = inI, hp = hpL;
                    unsigned w[WPL];
#pragma unroll
                    for (int k = 0; k < WPL; ++k)
                        w[k] = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        int s;
                        if (LOOKUP == 0)
                            s = prmt(t01[c], t23[c], sel);
                        else
                            s = s_scores[rowoff + q[c]];
                        const int MH = madd(hp, one, s); // M(r,j) + 2*FH
                        if (TRACE) {
                            // cIh = clean I with the H-stage tag of the I candidate OR-ed in: one LOP3
                            int cIh;
                            asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(cIh) : "r"(It), "r"(CLR), "r"(fh_reg));
                            const int cD = Dt[c] & CLR;
                            const int Ht = max3(MH, cIh, cD);
                            if (STORE)
                                w[c / 5] = shf_r_wrap(w[c / 5], (unsigned)xor3(It, Dt[c], Ht), kTagBits);
                            It = max3(madd(MH, one, kI), madd(cIh, one, iI - FH), madd(cD, one, iD));
                            Dt[c] = max3(madd(MH, one, aM[c]), madd(cIh, one, aI[c]), madd(cD, one, aD[c]));
                            hp = Hc[c];
                            Hc[c] = Ht & CLR;
                        } else {
                            const int H = max3(MH, It, Dt[c]);
                            const int Ho = madd(H, one, oe_s);
                            It = addmax(It, e_s, Ho);
                            // regular column: max(D+E, H+O+E); freeEndGaps last column: max(D, H) = H
                            Dt[c] = FREE ? addmax(Dt[c], aD[c], madd(H, one, aI[c])) : addmax(Dt[c], e_s, Ho);
                            hp = Hc[c];
                            Hc[c] = H;
                        }
                    }
                    edgeI = It;
                    edgeH = Hc[C - 1];
                    hpL = inH;
                    if (STORE) {
#pragma unroll
                        for (int k = 0; k < WPL; ++k)
                            tp[(size_t)k * 32] = w[k];
                    }
                    if (store_edge)
                        eout[r] = make_int2(edgeI, edgeH);
                }
                if (STORE)
                    tp += WPL * 32;
            };

            int t = 0;
            const int ramp = min(31, T);
            for (; t < ramp; ++t)
                step(t, std::true_type{});
            for (; t < n - 1; ++t) // steady: 1 <= t - lane + 1 <= n - 1 for every lane
                step(t, std::false_type{});
            for (; t < T; ++t)
                step(t, std::true_type{});

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

} // namespace gnx

// gnx_long.cuh -- LONG pairs (BASELINE config C4: 10 kb x 10 kb global affine + CIGAR): tile checkpoints and
// path-tile recompute instead of a traceback matrix.
//
// The reference's own answer to long sequences is a checkerboard (align/affineGap.go:73-144): a score-only pass
// that keeps boundary rows/columns (highestScore_affineGap, :151-207), then a re-fill with trace of only the
// boards the path crosses (fillTraceback_affineGap, :219-273).  The trace-matrix path (affine_fill3_kernel<MODE 2>
// + traceback_affine_warp_kernel) instead wrote 82 MB of 6-bit codes per 10 kb x 10 kb pair, ran the 15-slot
// tagged cell on every cell and could keep only ~1500 pairs in flight in 126 GB.  Here ONE kernel does, per warp
// and per pair (pairs are fetched from a device counter, so the tail of a batch spreads over all SMs):
//
//   pass 1  score-only int32 sweep (6 issue slots per cell) through the 320-column strips, lane l owning 10
//           columns with a one-row skew -- the wavefront of affine_fill3_kernel -- keeping
//             * every strip's right edge column (I', H per row; I' tagged with its source plane: the one tag a
//               recompute of the next strip cannot rebuild, because it belongs to the previous strip's last
//               column), and
//             * the wavefront's 23 state registers per lane every kLongR steps
//           in a scratch area private to the WARP (not to the pair): 2.6 + 3.7 MB at 10 kb x 10 kb, reused for
//           the warp's next pair, so the workspace is 6.4 MB x resident warps (15 GB at 16 warps/SM) whatever the
//           batch size, and every SM runs a full complement of warps.
//   pass 2  affineTrace (align/affineGap_highMem.go:57-89) from (n,m): for the tile (strip p, steps
//           kLongR*b+1 .. kLongR*(b+1)) holding the current cell, restore the checkpoint, re-run those steps with
//           the TAGGED arithmetic of affine_fill3_kernel (same instructions, hence the same M >= I >= D
//           tie-breaks), 6-bit codes into a 66 KB tile buffer (L2), and walk them warp-cooperatively (32 diagonal
//           cells looked up at once while the route is in plane M) until the route leaves the tile.  A 10 kb
//           route crosses ~77 of the pair's 1250 tiles: 6 % of the cells are recomputed.
//
// A checkpoint cannot carry the tags the tagged kernel computes one step ahead (source of D(i,j), source of a
// lane's incoming I), so a tile restarted at step s serves steps s+1 .. s+kLongR only (tile 0 starts from the true
// initial state and serves steps 0 .. kLongR) -- the same rule as gnx_ckpt.cuh.
//
// Requires gap_open <= 0 (score-only recurrences I' = max(I+E, H+O+E), D' likewise), dim <= 5 and the int32
// range proof of analyse() at scale 64.
#pragma once
#include "gnx_fill3.cuh"
#include <algorithm>

namespace gnx {

constexpr int kLongR = 256;      // steps between row checkpoints
constexpr int kLongRegs = 23;    // 32-bit words per lane per checkpoint: Hc[10], Dt[10], hpL, edgeI, edgeH
constexpr int kLongCols = 320;   // columns per strip (32 lanes x 10)

struct LongParams {
    uint8_t *scratch;          // per-CTA (= per-warp) scratch areas
    int64_t cta_stride;        // bytes per CTA
    int64_t edge_stride;       // int2 entries per strip edge column (>= n_max + 2)
    int64_t ckpt_off;          // byte offset of the checkpoint area inside a CTA's scratch
    int64_t ckpt_strip_words;  // 32-bit words per strip in the checkpoint area
    int64_t tile_off;          // byte offset of the trace tile ((kLongR + 1) * 2 * 32 words)
    int *next_pair;            // device work counter: chunk-local index of the next pair to take
    int64_t runs_off;          // byte offset of the warp's run buffer (n_max + m_max + 2 entries: a route never has more)
    // Cigars of long pairs run from a handful to thousands of elements, so a pair's runs (run << 2 | op, traceback
    // order) are appended to a chunk-wide pool at an offset taken from a device cursor; a pair that finds the pool
    // full is marked -1 and re-run by pass 1, which writes its final cigar directly.
    long long *slot64;         // per pair in chunk: offset of its runs in `pool`, or -1
    uint32_t *pool;
    long long pool_cap;        // entries
    unsigned long long *pool_cursor;
    int *counts;               // per pair in chunk
    int pass;                  // 0: score, runs, counts; 1: pairs marked -1 write their final cigar
    const int64_t *cigar_off;  // pass 1
    CigarOut *out_cigar;
    int64_t out_cap;
    int h00_plane;
};

// FORM 0: the cell of affine_fill3_kernel<MODE 0>: H = max3(M, I, D); I' = max(I + E, H + O + E); D' likewise --
//         3 ALU-pipe + 3 FMA-pipe instructions, but the I chain across a lane's 10 columns is three dependent
//         instructions per cell.
// FORM 1: X = max(M, D) + O + E first (independent of I), then I' = max(I + E, X): ONE dependent instruction per
//         cell on the chain; H + O + E = max(I + O + E, X).  4 ALU + 2 FMA; H is carried as H + O + E.
#ifndef GNX_LONG_MINB
#define GNX_LONG_MINB 20 // 96 registers, no spills: 2594 vs 2517 GCUPS at 16 (24: no further gain), profiles/r02d
#endif
template <bool FREE, int FORM>
__global__ void __launch_bounds__(32, FREE ? 12 : GNX_LONG_MINB) affine_long_kernel(const FillParams P, const LongParams Q)
{
    constexpr int C = 10, R = kLongR;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SC = kScale, FI = kFI, FD = kFD, FH = kFH;
    constexpr int NEG = kNeg32, CLR = ~(kScale - 1);
    __shared__ int s_tab[C * kDimP * 32]; // [c][a][lane]
    __shared__ int2 s_ein[32], s_eout[32]; // pass 1: 32-row rings of the strip's incoming / outgoing edge column
    const int lane = threadIdx.x;
    unsigned eout_s = (unsigned)__cvta_generic_to_shared(s_eout); // lane 31's row r = t - 30 goes to slot r & 31
    asm volatile("" : "+r"(eout_s)); // opaque: otherwise ptxas re-derives the shared window base at every store
    const int one = P.one, four = 4 * one;
    const int O = P.gap_open, E = P.gap_extend, oe = O + E;
    const int TO = FORM == 1 ? oe : 0; // pass 1 carries H as H + TO
    uint8_t *my = Q.scratch + (size_t)blockIdx.x * Q.cta_stride;
    int2 *edge = reinterpret_cast<int2 *>(my);
    uint32_t *ckpt = reinterpret_cast<uint32_t *>(my + Q.ckpt_off);
    uint32_t *tile = reinterpret_cast<uint32_t *>(my + Q.tile_off);
    uint32_t *runs = reinterpret_cast<uint32_t *>(my + Q.runs_off);
    const int np = (int)(P.pair_end - P.pair_begin);
    // tagged-domain constants (pass 2), as in affine_fill3_kernel
    const int oe_s = oe * SC, e_s = E * SC;
    const int kI = oe_s + 2 * FI - 2 * FH;
    const int iI = e_s + FI, iD = oe_s;
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;
    const int fh_reg = FH * one;

    // pass 0: move the warp's runs into the chunk's pool
    auto publish = [&](int idx, int cnt) {
        if (Q.pass == 0) {
            long long off = -1;
            if (lane == 0) {
                const unsigned long long o = atomicAdd(Q.pool_cursor, (unsigned long long)cnt);
                off = (o + (unsigned long long)cnt <= (unsigned long long)Q.pool_cap) ? (long long)o : -1;
                Q.slot64[idx] = off;
                Q.counts[idx] = cnt;
            }
            off = __shfl_sync(FULL, off, 0);
            __syncwarp();
            if (off >= 0)
                for (int k = lane; k < cnt; k += 32)
                    Q.pool[off + k] = __ldcg(runs + k);
        }
        __syncwarp();
    };
    while (true) {
        int idx = 0;
        if (lane == 0)
            idx = atomicAdd(Q.next_pair, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= np)
            break;
        const int64_t pair = P.pair_begin + idx;
        if (P.pair_class && P.pair_class[pair] > 1) { // invalid base: the call returns GNX_EBASE
            if (Q.pass == 0 && lane == 0) {
                Q.counts[idx] = 0;
                Q.slot64[idx] = 0;
            }
            continue;
        }
        int total = 0;
        CigarOut *dst = nullptr;
        if (Q.pass == 1) {
            total = Q.counts[idx];
            if (Q.slot64[idx] >= 0 || Q.cigar_off[idx] + total > Q.out_cap)
                continue;
            dst = Q.out_cigar + Q.cigar_off[idx];
        }
        int cnt = 0;
        auto emit = [&](int op, int len) { // every lane counts, lane 0 stores
            if (lane == 0) {
                if (Q.pass == 0) {
                    runs[cnt] = ((uint32_t)len << 2) | (uint32_t)op;
                } else {
                    CigarOut o;
                    o.run_length = len;
                    o.op = (unsigned char)op;
                    dst[total - 1 - cnt] = o;
                }
            }
            ++cnt;
        };
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0), m = (int)(P.beta_off[pair + 1] - b0);
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        if (n == 0 || m == 0) { // closed forms of the boundary row / column (affineGap_highMem.go:185-206, :58)
            if (lane == 0 && Q.pass == 0) {
                int64_t sc;
                if (n == 0 && m == 0)
                    sc = P.h00;
                else if (n == 0)
                    sc = (int64_t)O + (int64_t)m * E;
                else
                    sc = FREE ? 0 : (int64_t)O + (int64_t)n * E;
                P.out_score[pair] = sc;
            }
            if (n == 0 && m == 0)
                emit(0, 0); // route := make([]Cigar, 1)
            else
                emit(n == 0 ? 1 : 2, n == 0 ? m : n);
            publish(idx, cnt);
            continue;
        }
        const int T = n + 31;
        const int strips = (m + kLongCols - 1) / kLongCols;

        // =========================== pass 1: score-only sweep with checkpoints ===========================
        // Steps run in blocks of 32.  Per block the warp moves the strip's left boundary column (32 rows: loaded
        // from the previous strip's edge column one block ahead, or computed for strip 0) into a shared-memory
        // ring that lane 0 reads with one broadcast LDS per step, and flushes the 32 rows lane 31 left in the
        // outgoing ring with one coalesced 256-byte store -- the per-step edge traffic of affine_fill3_kernel
        // (two index shuffles, 64-bit address arithmetic and a lane-31 store) cost 40 of its 115 instructions
        // per step (ncu source page, profiles/r02b_long_first.md).
        for (int p = 0; p < strips; ++p) {
            const int jbase = p * kLongCols + lane * C;
            int aD[FREE ? C : 1], aH[FREE ? C : 1]; // D-plane addends: (E, O+E) regular, (0, 0) in the free-end column
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const int q = (j <= m) ? (int)beta[j - 1] : 0;
#pragma unroll
                for (int a = 0; a < kDimP; ++a) {
                    int v = 0;
                    if (a < P.dim && q < P.dim)
                        v = P.scores[a * P.dim + q];
                    s_tab[(c * kDimP + a) * 32 + lane] = v;
                }
                if (FREE) {
                    const bool last = j == m;
                    aD[c] = last ? 0 : E;
                    aH[c] = (last ? 0 : oe) - TO;
                }
            }
            int Dt[C], Hc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const int i0 = O + j * E;                    // H(0,j) = I(0,j)
                Hc[c] = i0 + TO;
                Dt[c] = i0 + ((FREE && j == m) ? 0 : oe);    // D(1,j) = T(-inf, I(0,j) + O + E, -inf)
            }
            int hpL = ((jbase == 0) ? P.h00 : (O + jbase * E)) + TO;
            int edgeI = 0, edgeH = 0;
            const int2 *ein = edge + (size_t)(p > 0 ? p - 1 : 0) * Q.edge_stride; // written by strip p - 1
            int2 *eout = edge + (size_t)p * Q.edge_stride;
            const bool has_next = p + 1 < strips;
            uint32_t *ck = ckpt + (size_t)p * Q.ckpt_strip_words + lane;
            // rows t0 + 1 .. t0 + 32 of the left boundary column: (4 * I'(r, jbase+1) + source plane tag, H(r, jbase))
            auto edge_rows = [&](int t0) {
                const int rho = t0 + 1 + lane;
                return (p > 0 && rho <= n) ? __ldcg(&ein[rho]) : make_int2(0, 0);
            };
            int2 eb_next = edge_rows(0);
            auto refill = [&](int t0) {
                int2 v = eb_next;
                if (p == 0) { // column 0: M = I = -inf, D(r,0) = O + rE (0 with free end gaps); I'(r,1) = D + O + E, tag D
                    const int d0 = FREE ? 0 : O + (t0 + 1 + lane) * E;
                    v = make_int2(4 * (d0 + oe), d0);
                } else {
                    eb_next = edge_rows(t0 + 32);
                }
                __syncwarp();
                s_ein[lane] = v;
                __syncwarp();
            };
            // lane 31 has finished rows <= t0 - 31 when block t0 starts: write rows t0 - 62 .. t0 - 31 out
            auto flush = [&](int first_row) {
                __syncwarp();
                const int rr = first_row + lane;
                if (rr >= 1 && rr <= n)
                    eout[rr] = s_eout[rr & 31];
                __syncwarp();
            };
            __syncwarp();
            int a_next = (lane == 0) ? (int)alpha[0] : 0;

            auto step = [&](int t, auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int r = t - lane + 1;
                const int2 eb = s_ein[t & 31]; // one broadcast LDS.64: lane 0's row t + 1
                int inI = __shfl_up_sync(FULL, edgeI, 1);
                int inH = __shfl_up_sync(FULL, edgeH, 1);
                if (lane == 0) {
                    inI = eb.x >> 2;
                    inH = eb.y + TO;
                }
                const int a = a_next;
                bool active = true;
                if (CHECK) {
                    active = (unsigned)(r - 1) < (unsigned)n;
                    if ((unsigned)r < (unsigned)n)
                        a_next = alpha[r];
                } else {
                    a_next = alpha[r];
                }
                if (active) {
                    const int *row = s_tab + a * 32 + lane;
                    int It = inI, hp = hpL, tg = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int s = row[c * kDimP * 32];
                        if (FORM == 0) {
                            const int MH = madd(hp, one, s);
                            const int H = max3(MH, It, Dt[c]);
                            const int Ho = madd(H, one, oe);
                            if (c == C - 1) // source plane of I'(r, j+1), M >= I >= D: the next strip's first I tag
                                tg = max3(madd(MH, four, 4 * oe + 2), madd(It, four, 4 * E + 1), madd(Dt[c], four, 4 * oe));
                            It = addmax(It, E, Ho);
                            if (FREE)
                                Dt[c] = addmax(Dt[c], aD[c], madd(H, one, aH[c]));
                            else
                                Dt[c] = addmax(Dt[c], E, Ho);
                            hp = Hc[c];
                            Hc[c] = H;
                        } else {
                            const int MHo = madd(hp, one, s);      // M + O + E (hp carries H + O + E)
                            const int Do = madd(Dt[c], one, oe);   // D + O + E
                            const int X = max(MHo, Do);
                            if (c == C - 1)
                                tg = max3(madd(MHo, four, 2), madd(It, four, 4 * E + 1), madd(Do, four, 0));
                            const int Ht = addmax(It, oe, X);      // H + O + E
                            It = addmax(It, E, X);                 // the only instruction on the I chain
                            if (FREE)
                                Dt[c] = addmax(Dt[c], aD[c], madd(Ht, one, aH[c]));
                            else
                                Dt[c] = addmax(Dt[c], E, Ht);
                            hp = Hc[c];
                            Hc[c] = Ht;
                        }
                    }
                    edgeI = It;
                    edgeH = Hc[C - 1];
                    hpL = inH;
                    if (lane == 31) // shared-space address formed by hand: the generic form costs six instructions per store
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(eout_s + (((unsigned)(t + 2) & 31u) << 3)), "r"(tg),
                                     "r"(Hc[C - 1] - TO)
                                     : "memory");
                }
            };
            auto save = [&](int s) { // state entering step s = R * k, k >= 1 (true values)
                uint32_t *d = ck + (size_t)(s / R - 1) * (kLongRegs * 32);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    d[c * 32] = (uint32_t)(Hc[c] - TO);
                    d[(C + c) * 32] = (uint32_t)Dt[c];
                }
                d[20 * 32] = (uint32_t)(hpL - TO);
                d[21 * 32] = (uint32_t)edgeI;
                d[22 * 32] = (uint32_t)(edgeH - TO);
            };
#pragma unroll 1
            for (int t0 = 0; t0 < T; t0 += 32) {
                if (t0 > 0 && (t0 & (R - 1)) == 0)
                    save(t0);
                refill(t0);
                if (has_next && t0 >= 32)
                    flush(t0 - 62);
                if (t0 >= 32 && t0 + 32 <= n - 1) { // steady: every lane is on a valid row < n for the whole block
#pragma unroll 2
                    for (int t = t0; t < t0 + 32; ++t)
                        step(t, std::false_type{});
                } else {
                    const int t1 = min(T, t0 + 32);
#pragma unroll 1
                    for (int t = t0; t < t1; ++t)
                        step(t, std::true_type{});
                }
            }
            if (has_next)
                flush(((T - 1) & ~31) - 30); // rows not yet written out (at most 32, see the flush rule above)
            if (p == strips - 1 && lane == ((m - 1) % kLongCols) / C && Q.pass == 0) {
                const int cm = (m - 1) % C;
                int h = Hc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (c == cm)
                        h = Hc[c];
                P.out_score[pair] = (int64_t)(h - TO);
            }
            __syncwarp(); // the strip's edge column and checkpoints are visible to the whole warp
        }

        // =========================== pass 2: walk + recompute of the route's tiles ===========================
        int i = n, j = m, k = 0, need_k = 1, cur_op = -1, run = 0;
        auto add_run = [&](int op, int len) {
            if (op == cur_op) {
                run += len;
            } else {
                if (cur_op >= 0)
                    emit(cur_op, run);
                cur_op = op;
                run = len;
            }
        };
        while (i > 0 || j > 0) {
            if (i == 0 || j == 0) {
                // row 0 is plane I, column 0 plane D up to (0,0) (M(0,j), D(0,j), M(i,0), I(i,0) are -inf, so no
                // finite route enters the boundary in another plane): the rest of the route is one run
                add_run(i == 0 ? 1 : 2, i == 0 ? j : i);
                break;
            }
            // ---- the tile serving cell (i, j): strip p, steps lo .. s0 + ulast ----
            const int p = (j - 1) / kLongCols;
            const int tcell = (i - 1) + ((j - 1) - p * kLongCols) / C;
            const int blk = max(tcell - 1, 0) / R;
            const int s0 = blk * R;
            const int lo = blk > 0 ? s0 + 1 : 0;
            const int ulast = tcell - s0;
            const int jbase = p * kLongCols + lane * C;
            {
                int aM[FREE ? C : 1], aI[FREE ? C : 1], aD[FREE ? C : 1];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int jj = jbase + c + 1;
                    const int q = (jj <= m) ? (int)beta[jj - 1] : 0;
#pragma unroll
                    for (int a = 0; a < kDimP; ++a) {
                        int v = 0;
                        if (a < P.dim && q < P.dim)
                            v = P.scores[a * P.dim + q] * SC + 2 * FH;
                        s_tab[(c * kDimP + a) * 32 + lane] = v;
                    }
                    if (FREE) {
                        const bool last = jj == m;
                        aM[c] = last ? dMl : dMn;
                        aI[c] = last ? dIl : dIn;
                        aD[c] = last ? dDl : dDn;
                    }
                }
                int Hc[C], Dt[C];
                int hpL, edgeI = 0, edgeH = 0;
                if (blk > 0) {
                    const uint32_t *src = ckpt + (size_t)p * Q.ckpt_strip_words + (size_t)(blk - 1) * (kLongRegs * 32) + lane;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        Hc[c] = (int)__ldcg(src + c * 32) * SC;
                        Dt[c] = (int)__ldcg(src + (C + c) * 32) * SC;
                    }
                    hpL = (int)__ldcg(src + 20 * 32) * SC;
                    edgeI = (int)__ldcg(src + 21 * 32) * SC;
                    edgeH = (int)__ldcg(src + 22 * 32) * SC;
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int jj = jbase + c + 1;
                        const int i0 = (O + jj * E) * SC;
                        const int am = FREE ? aM[c] : dMn, ai = FREE ? aI[c] : dIn, ad = FREE ? aD[c] : dDn;
                        Hc[c] = i0;
                        Dt[c] = max3(NEG + 2 * FH + am, i0 + FH + ai, NEG + ad);
                    }
                    hpL = (jbase == 0) ? P.h00 * SC : (O + jbase * E) * SC;
                }
                const int2 *ein = edge + (size_t)(p > 0 ? p - 1 : 0) * Q.edge_stride;
                const int r0 = s0 - lane + 1;
                int2 eb = make_int2(0, 0);
                int a_next = 0;
                if ((unsigned)(r0 - 1) < (unsigned)n) {
                    a_next = alpha[r0 - 1];
                    if (p > 0 && lane == 0)
                        eb = __ldcg(&ein[r0]);
                }
                __syncwarp();
#pragma unroll 1
                for (int u = 0; u <= ulast; ++u) {
                    const int t = s0 + u;
                    const int r = t - lane + 1;
                    int inI = __shfl_up_sync(FULL, edgeI, 1);
                    int inH = __shfl_up_sync(FULL, edgeH, 1);
                    if (lane == 0) {
                        if (p == 0) {
                            const int d0 = FREE ? 0 : (O + r * E) * SC;
                            inI = d0 + iD; // T(-inf, -inf, D + O + E): tag D (0)
                            inH = d0;
                        } else {
                            inI = (eb.x & ~3) * (SC / 4) + (eb.x & 3);
                            inH = eb.y * SC;
                        }
                    }
                    const bool active = (unsigned)(r - 1) < (unsigned)n;
                    const int a = a_next;
                    if ((unsigned)r < (unsigned)n) { // next step's row r + 1
                        a_next = alpha[r];
                        if (p > 0 && lane == 0)
                            eb = __ldcg(&ein[r + 1]);
                    }
                    unsigned w0 = 0, w1 = 0;
                    if (active) {
                        const int *row = s_tab + a * 32 + lane;
                        int It = inI, hp = hpL;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int s = row[c * kDimP * 32];
                            const int MH = madd(hp, one, s);
                            int cIh;
                            asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(cIh) : "r"(It), "r"(CLR), "r"(fh_reg));
                            const int cD = Dt[c] & CLR;
                            const int Ht = max3(MH, cIh, cD);
                            const unsigned code = (unsigned)xor3(It, Dt[c], Ht);
                            if (c < 5)
                                w0 = shf_r_wrap(w0, code, kTagBits);
                            else
                                w1 = shf_r_wrap(w1, code, kTagBits);
                            const int am = FREE ? aM[c] : dMn, ai = FREE ? aI[c] : dIn, ad = FREE ? aD[c] : dDn;
                            It = max3(madd(MH, one, kI), madd(cIh, one, iI - FH), madd(cD, one, iD));
                            Dt[c] = max3(madd(MH, one, am), madd(cIh, one, ai), madd(cD, one, ad));
                            hp = Hc[c];
                            Hc[c] = Ht & CLR;
                        }
                        edgeI = It;
                        edgeH = Hc[C - 1];
                        hpL = inH;
                    }
                    tile[(u * 2 + 0) * 32 + lane] = w0;
                    tile[(u * 2 + 1) * 32 + lane] = w1;
                }
            }
            __syncwarp();
            // ---- walk the tile: every lane keeps the same state; while the route is in plane M, lane d-1 looks at
            // the diagonal cell (i-d, j-d) and the run of leading "M" answers is consumed at once ----
            const int jlo = p * kLongCols; // columns of this strip are jlo + 1 .. jlo + 320
            auto in_tile = [&](int ci, int cj) { return cj > jlo && (ci - 1) + ((cj - 1) - jlo) / C >= lo; };
            auto load = [&](int ci, int cj) -> unsigned {
                const int within = (cj - 1) - jlo;
                const int l = within / C, c = within - l * C;
                const int u = (ci - 1) + l - s0;
                const int q = c >= 5 ? 1 : 0, cc = c - 5 * q;
                return (__ldcg(tile + (u * 2 + q) * 32 + l) >> (32 - kTagBits * (5 - cc))) & (kScale - 1);
            };
            unsigned cur = load(i, j);
            if (need_k) {
                k = 2 - (int)((cur >> 4) & 3u);
                need_k = 0;
            }
            while (true) {
                if (k == 0) {
                    const int d = lane + 1, ci = i - d, cj = j - d;
                    const bool inside = ci > 0 && cj > 0 && in_tile(ci, cj);
                    const unsigned cd = inside ? load(ci, cj) : 0u;
                    const bool isM = inside && ((cd >> 4) & 3u) == 2u;
                    const int skip = __ffs(~__ballot_sync(FULL, isM)) - 1; // 0..32 leading diagonal cells in plane M
                    if (skip >= 1) { // this cell and the next skip-1 are M; land on cell d = skip (plane M, code known)
                        add_run(0, skip);
                        i -= skip;
                        j -= skip;
                        cur = __shfl_sync(FULL, cd, skip - 1);
                        continue;
                    }
                    add_run(0, 1);
                    --i;
                    --j;
                    if (i == 0 || j == 0)
                        break;
                    if (!in_tile(i, j)) {
                        need_k = 1; // the plane is the H tag of a cell of another tile
                        break;
                    }
                    cur = __shfl_sync(FULL, cd, 0);
                    k = 2 - (int)((cur >> 4) & 3u);
                    continue;
                }
                add_run(k, 1);
                const int kn = 2 - (int)((cur >> (k == 1 ? 0 : 2)) & 3u);
                i -= (k != 1);
                j -= (k != 2);
                k = kn;
                if (i == 0 || j == 0)
                    break;
                if (!in_tile(i, j))
                    break;
                cur = load(i, j);
            }
            __syncwarp();
        }
        if (cur_op >= 0)
            emit(cur_op, run);
        publish(idx, cnt);
    }
}

// Scratch geometry for a batch whose longest target / query are max_n / max_m.
struct LongGeom {
    int64_t edge_stride, ckpt_strip_words, ckpt_off, tile_off, runs_off, cta_stride;
};
inline LongGeom long_geom(int64_t max_n, int64_t max_m)
{
    LongGeom g;
    const int64_t strips = (max_m + kLongCols - 1) / kLongCols;
    g.edge_stride = (max_n + 2 + 1) & ~int64_t(1);
    const int64_t nck = (max_n + 30) / kLongR; // checkpoints per strip: steps R, 2R, ... < n + 31
    g.ckpt_strip_words = nck * kLongRegs * 32;
    g.ckpt_off = ((strips > 1 ? strips - 1 : 1) * g.edge_stride * 8 + 255) & ~int64_t(255);
    g.tile_off = (g.ckpt_off + strips * g.ckpt_strip_words * 4 + 255) & ~int64_t(255);
    g.runs_off = (g.tile_off + (int64_t)(kLongR + 1) * 2 * 32 * 4 + 255) & ~int64_t(255);
    g.cta_stride = (g.runs_off + (max_n + max_m + 2) * 4 + 255) & ~int64_t(255);
    return g;
}

// Pool entries for a chunk of np pairs: the route of an n x m pair has at most n + m + 1 runs; related sequences
// produce tens to hundreds and unrelated 10 kb pairs ~1400, so 2048 per pair on average is generous.
inline int64_t long_pool_entries(int64_t np, int64_t max_n, int64_t max_m)
{
    return std::max<int64_t>(64, std::min<int64_t>(np * (max_n + max_m + 2), std::max<int64_t>(np * 2048, int64_t(1) << 22)));
}
// Layout of the chunk's run storage (the Slot's `slots` buffer): [cursor, 16 B][slot64 x np][pool]
inline size_t long_slot_bytes(int64_t np, int64_t pool_entries) { return 16 + (size_t)np * 8 + (size_t)pool_entries * 4; }

// Final cigars of the pairs whose runs are in the pool: a warp per pair, reversed into start -> end order
// (align/align.go:86-90 reverseCigar).
__global__ void __launch_bounds__(128) expand_pool_kernel(const long long *slot64, const uint32_t *pool, const int *counts,
                                                          const int64_t *cigar_off, int64_t np, CigarOut *out,
                                                          int64_t out_cap, int *status)
{
    const int lane = threadIdx.x & 31;
    const int64_t idx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= np)
        return;
    const int cnt = counts[idx];
    const int64_t off = cigar_off[idx];
    if (off + cnt > out_cap) {
        if (lane == 0)
            atomicMax(status, kECap);
        return;
    }
    const long long so = slot64[idx];
    if (so < 0)
        return; // rewritten by pass 1 of affine_long_kernel
    for (int k = lane; k < cnt; k += 32) {
        const uint32_t v = pool[so + k];
        CigarOut o;
        o.run_length = (long long)(v >> 2);
        o.op = (unsigned char)(v & 3u);
        out[off + cnt - 1 - k] = o;
    }
}

} // namespace gnx

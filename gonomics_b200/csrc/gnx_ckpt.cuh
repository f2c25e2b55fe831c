// gnx_ckpt.cuh -- second pass of the checkpoint-and-recompute traceback for uniform read-sized batches in
// freeEndGaps mode (AffineGapLocal, align/affineGap_highMem.go:105-107 -> affineGap_highMem :181-220, affineTrace
// :57-89): BASELINE config C3.
//
// The traced int32 kernel (affine_fill3_kernel<MODE 2>) is bound by its integer instruction count, not by HBM:
// ~15 issue slots per cell against 3.3 for the packed 16-bit score-only kernel.  A 500 x 150 semi-global
// alignment only needs trace codes where its path can go: the free-end column down to row r*, then ~m rows.
// So: pass 1 = affine_fill16_kernel<FREE, CM, CKPT> (score, r*, and the wavefront's register state every
// kCkK = 32 steps: 11.8 KB per pair instead of a 66 KB trace matrix); pass 2 (here) walks the path backwards
// block by block: restore the state entering step 32b, re-run 33 steps with the TAGGED arithmetic of
// affine_fill3_kernel (identical cell code, so identical tie-breaks), keep the 6-bit codes of those steps in
// shared memory, and walk them (state in lane 0 of the half-warp, the other lanes look 16 diagonal cells ahead so
// that match runs are consumed at once) until the path leaves the block.  On the C3 workload the path touches ~6
// of 16 blocks.  Between the passes ckpt_classify_kernel settles the pairs whose route is provably the ungapped
// diagonal (no trace needed at all) and queues the rest.
//
// Exactness: the checkpointed values are the clean plane values (H, D', edge I) the tagged kernel carries
// between steps (score-only I' = max(I+E, H+O+E) equals the tagged three-way max's value when O <= 0, which
// the 16-bit kernel requires).  What a checkpoint cannot carry are the TAGS of D(i,j) and of each lane's
// incoming I, which the tagged kernel computes one step before it stores them -- so the codes of a block's
// first re-run step are incomplete, and a block that restarts at step s only serves steps s+1 .. s+32
// (block 0 starts from the true initial state and serves steps 0 .. 32).
#pragma once
#include "gnx_fill16.cuh"

namespace gnx {

struct CkptParams {
    const uint32_t *ckpt;      // [quad][k][kCkRegs][32] packed 16x2 state words written by pass 1
    int64_t quad_words;        // words per quad
    const int64_t *rstar;      // per global pair: r* from pass 1
    uint32_t *slots;           // per pair in chunk: slot_cap entries, run << 2 | op, traceback order
    int slot_cap;
    int *counts;               // per pair in chunk
    int pass;                  // 0: slots + counts; 1: pairs whose count > slot_cap write their final cigar
    const int64_t *cigar_off;  // pass 1
    CigarOut *out_cigar;
    int64_t out_cap;
    int h00_plane;
    int *work;                 // pass 0: chunk-local indices of the pairs that need the recompute walk
    int *work_count;           // device counter of `work` (ckpt_classify_kernel)
};

// Screening pass between the two passes: one thread per pair.  If the pair's score equals the score of the
// UNGAPPED diagonal that ends at (r*, m), its route is known without any trace code:
//     D x (n - r*),  M x m,  D x (r* - m)          (traceback order; zero-length runs omitted)
// Proof.  Let c_k = (r* - m + k, k) and P_k the diagonal's prefix score.  M(c_k) >= P_k (the diagonal is a valid
// path: it starts from the free D(i,0) = 0 and M(c_{k+1}) >= s + M(c_k)), and M(r*,m) <= max(M,I)(r*,m) = S = P_m,
// so M(r*,m) = S >= I(r*,m): the walk enters in plane M.  If some X in {I, D} had X(c_{k-1}) > M(c_{k-1}), then
// M(c_k) >= s_k + X(c_{k-1}) > s_k + P_{k-1} = P_k and, propagating along the diagonal, M(c_m) > P_m = S:
// impossible.  Hence tripleMaxTrace (ties M >= I >= D, align/align.go:76-84) picks M at every diagonal cell and
// the walk stays on the diagonal down to column 0, where affineTrace's boundary is plane D up to (0,0).
// Everything else is queued for affine_ckpt_trace_kernel.  Reads with no indel against their window are the
// common case in practice; in the synthetic C3 workload they are ~45 % of the pairs.
__global__ void __launch_bounds__(128) ckpt_classify_kernel(const FillParams P, const CkptParams Q)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t pair = P.pair_begin + idx;
    if (pair >= P.pair_end)
        return;
    if (P.pair_class && P.pair_class[pair] > 1) {
        Q.counts[idx] = 0;
        return;
    }
    const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
    const int n = (int)(P.alpha_off[pair + 1] - a0), m = (int)(P.beta_off[pair + 1] - b0);
    const int rs = (int)Q.rstar[pair];
    bool shortcut = false;
    if (rs >= m && m >= 1) {
        const uint8_t *__restrict__ al = P.alpha + a0 + (rs - m);
        const uint8_t *__restrict__ be = P.beta + b0;
        long long sum = 0;
        for (int k = 0; k < m; ++k)
            sum += P.scores[(int)al[k] * P.dim + (int)be[k]];
        shortcut = sum == P.out_score[pair];
    }
    if (!shortcut) {
        Q.work[atomicAdd(Q.work_count, 1)] = (int)idx;
        return;
    }
    uint32_t *slot = Q.slots + (size_t)idx * Q.slot_cap; // slot_cap >= 3
    int cnt = 0;
    if (n > rs)
        slot[cnt++] = ((uint32_t)(n - rs) << 2) | 2u;
    slot[cnt++] = ((uint32_t)m << 2) | 0u;
    if (rs > m)
        slot[cnt++] = ((uint32_t)(rs - m) << 2) | 2u;
    Q.counts[idx] = cnt;
}

__global__ void __launch_bounds__(32, 12) affine_ckpt_trace_kernel(const FillParams P, const CkptParams Q)
{
    constexpr int C = 10, LPP = 16, WPL = 2;
    constexpr int SC = kScale, FI = kFI, FD = kFD, FH = kFH;
    constexpr int NEG = kNeg32, CLR = ~(kScale - 1);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kTgtPitch = kRing + 64;
    __shared__ int s_tab[C * kDimP * 32];
    __shared__ uint8_t s_tgt[2 * kTgtPitch];
    __shared__ uint32_t s_tr[(kCkK + 1) * WPL * 32];
    const int tid = threadIdx.x, lane = tid % LPP, half = tid / LPP;
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_s = (O + E) * SC, e_s = E * SC;
    const int kI = oe_s + 2 * FI - 2 * FH;
    const int iI = e_s + FI, iD = oe_s;
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;
    const int fh_reg = FH * one;
    const int64_t np = P.pair_end - P.pair_begin;
    // pass 0 takes two queued pairs per warp from the work list; pass 1 (the rare cigars longer than the slot) first
    // screens 32 candidate units per warp, one per lane, a unit being two pairs (pl, pl + 2) of a quad
    const int64_t n_work = Q.pass == 0 ? (int64_t)*Q.work_count : 0;
    const int64_t n_units = Q.pass == 0 ? (n_work + 1) / 2 : ((np + 3) / 4) * 2;
    const int G = Q.pass == 1 ? 32 : 1;
    for (int64_t ubase = (int64_t)blockIdx.x * G; ubase < n_units; ubase += (int64_t)gridDim.x * G) {
      unsigned todo = 1u;
      if (Q.pass == 1) {
          const int64_t u = ubase + tid;
          bool need = false;
          if (u < n_units) {
              const int64_t p0 = (u >> 1) * 4 + (u & 1); // chunk-local index of the unit's first pair; second is p0 + 2
              need = (p0 < np && Q.counts[p0] > Q.slot_cap) || (p0 + 2 < np && Q.counts[p0 + 2] > Q.slot_cap);
          }
          todo = __ballot_sync(FULL, need);
      }
      while (todo) {
        const int64_t unit = ubase + (__ffs(todo) - 1);
        todo &= todo - 1;
        // this half-warp's pair (chunk-local index pl) and where its checkpoint words live: quad pl / 4, lanes
        // 16 * ((pl % 4) / 2) .. + 15 of each record, 16-bit half pl % 2
        int64_t pl;
        if (Q.pass == 0)
            pl = (2 * unit + half < n_work) ? (int64_t)Q.work[2 * unit + half] : -1;
        else
            pl = (unit >> 1) * 4 + half * 2 + (unit & 1);
        const bool valid = pl >= 0 && pl < np;
        const int64_t idx = valid ? pl : 0;
        const int64_t pair = P.pair_begin + idx;
        const int64_t quad = idx >> 2;
        const int src_lane = (int)(((idx >> 1) & 1) * LPP) + lane; // lane of the pass-1 warp that held these columns
        const int sel = (int)(idx & 1);
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0), m = (int)(P.beta_off[pair + 1] - b0); // uniform batch
        bool want = valid && (!P.pair_class || P.pair_class[pair] <= 1);
        if (Q.pass == 1)
            want = want && Q.counts[idx] > Q.slot_cap && Q.cigar_off[idx] + Q.counts[idx] <= Q.out_cap;
        if (!__any_sync(FULL, want)) {
            if (Q.pass == 0 && valid && lane == 0 && !want)
                Q.counts[idx] = 0;
            continue;
        }
        if (Q.pass == 0 && valid && lane == 0 && !want)
            Q.counts[idx] = 0;
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        const int T = n + LPP - 1;
        const int jbase = lane * C;

        // ---- per-lane score tables, addends, staged target: exactly affine_fill3_kernel's set-up ----
        int aM[C], aI[C], aD[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c + 1;
            const int q = (j <= m) ? (int)beta[j - 1] : 0;
#pragma unroll
            for (int a = 0; a < kDimP; ++a) {
                int v = 0;
                if (a < P.dim && q < P.dim)
                    v = P.scores[a * P.dim + q] * SC + 2 * FH;
                s_tab[(c * kDimP + a) * 32 + tid] = v;
            }
            const bool last = j == m;
            aM[c] = last ? dMl : dMn;
            aI[c] = last ? dIl : dIn;
            aD[c] = last ? dDl : dDn;
        }
        for (int i = lane; i < n; i += LPP)
            s_tgt[half * kTgtPitch + i] = alpha[i];
        const uint8_t *tg = s_tgt + half * kTgtPitch;
        __syncwarp();

        // ---- walk state, kept by lane 0 of each half-warp ----
        const int rs = want ? (int)Q.rstar[pair] : 0;
        int wi = rs, wj = m;            // current cell
        int wk = 2;                     // current plane (0 M, 1 I, 2 D)
        int need_k = 1;                 // the plane of the current cell is the H tag of its code
        int cur_op = 2, run = n - rs;   // the free-end column's D run from (n,m) down to (r*,m)
        int cnt = 0;
        bool done = !want;
        uint32_t *slot = Q.slots + (size_t)idx * Q.slot_cap;
        int total = 0;
        CigarOut *dst = nullptr;
        if (Q.pass == 1 && want) {
            total = Q.counts[idx];
            dst = Q.out_cigar + Q.cigar_off[idx];
        }
        auto emit = [&](int op, int len) {
            if (Q.pass == 0) {
                if (cnt < Q.slot_cap)
                    slot[cnt] = ((uint32_t)len << 2) | (uint32_t)op;
            } else {
                CigarOut o;
                o.run_length = len;
                o.op = (unsigned char)op;
                dst[total - 1 - cnt] = o;
            }
            ++cnt;
        };
        const unsigned code00 = (unsigned)(2 - Q.h00_plane) << 4;

        while (true) {
            // block each half needs: the one serving the step of its current cell
            int blk = 0;
            {
                const int tcell = (wi - 1) + (wj - 1) / C;
                blk = (wi > 0 && wj > 0) ? max(tcell - 1, 0) / kCkK : -1; // -1: only boundary cells remain
            }
            blk = __shfl_sync(FULL, blk, 0, LPP);
            const bool hdone = __shfl_sync(FULL, (int)done, 0, LPP) != 0;
            if (__all_sync(FULL, hdone))
                break;
            const bool recompute = !hdone && blk >= 0;
            const int s0 = blk > 0 ? blk * kCkK : 0;
            // the walk enters the block at its current cell and only moves to earlier steps: no later step is needed
            int tin = (wi - 1) + (wj - 1) / C;
            tin = __shfl_sync(FULL, tin, 0, LPP);
            const int ulast_h = recompute ? min(tin - s0, kCkK) : 0;
#ifdef GNX_CK_NOTRIM
            const int ulast = kCkK;
            (void)ulast_h;
#else
            const int ulast = max(ulast_h, __shfl_xor_sync(FULL, ulast_h, 16));
#endif
            if (__any_sync(FULL, recompute)) {
                // ---- restore the state entering step s0 ----
                int Hc[C], Dt[C];
                int hpL, edgeI = 0, edgeH = 0;
                if (blk > 0 && recompute) {
                    const uint32_t *src = Q.ckpt + (size_t)quad * Q.quad_words + (size_t)(blk - 1) * (kCkRegs * 32) + src_lane;
                    auto val = [&](uint32_t x) { return ((int)((x >> (16 * sel)) & 0xffffu) - 32768) * SC; };
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        Hc[c] = val(__ldg(src + c * 32));
                        Dt[c] = val(__ldg(src + (C + c) * 32));
                    }
                    hpL = val(__ldg(src + 20 * 32));
                    edgeI = val(__ldg(src + 21 * 32));
                    edgeH = val(__ldg(src + 22 * 32));
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = jbase + c + 1;
                        const int i0 = (O + j * E) * SC;
                        Hc[c] = i0;
                        Dt[c] = max3(NEG + 2 * FH + aM[c], i0 + FH + aI[c], NEG + aD[c]);
                    }
                    hpL = (jbase == 0) ? P.h00 * SC : (O + jbase * E) * SC;
                }
                // ---- re-run steps s0 .. s0 + kCkK with the tagged arithmetic, codes into shared memory ----
#pragma unroll 1
                for (int u = 0; u <= ulast; ++u) {
                    const int t = s0 + u;
                    const int r = t - lane + 1;
                    int inI = __shfl_up_sync(FULL, edgeI, 1, LPP);
                    int inH = __shfl_up_sync(FULL, edgeH, 1, LPP);
                    if (lane == 0) { // freeEndGaps: D(i,0) = 0, so I's candidate from column 0 is O + E
                        inI = iD;
                        inH = 0;
                    }
                    const bool active = recompute && t < T && (unsigned)(r - 1) < (unsigned)n;
                    unsigned w[WPL] = {0, 0};
                    if (active) {
                        const int a = tg[r - 1];
                        const int *row = s_tab + a * 32 + tid;
                        int It = inI, hp = hpL;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int s = row[c * kDimP * 32];
                            const int MH = madd(hp, one, s);
                            int cIh;
                            asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(cIh) : "r"(It), "r"(CLR), "r"(fh_reg));
                            const int cD = Dt[c] & CLR;
                            const int Ht = max3(MH, cIh, cD);
                            w[c / 5] = shf_r_wrap(w[c / 5], (unsigned)xor3(It, Dt[c], Ht), kTagBits);
                            It = max3(madd(MH, one, kI), madd(cIh, one, iI - FH), madd(cD, one, iD));
                            Dt[c] = max3(madd(MH, one, aM[c]), madd(cIh, one, aI[c]), madd(cD, one, aD[c]));
                            hp = Hc[c];
                            Hc[c] = Ht & CLR;
                        }
                        edgeI = It;
                        edgeH = Hc[C - 1];
                        hpL = inH;
                    }
                    s_tr[(u * WPL + 0) * 32 + tid] = w[0];
                    s_tr[(u * WPL + 1) * 32 + tid] = w[1];
                }
            }
            __syncwarp();
            // ---- walk: affineTrace over the block's codes.  The state lives in lane 0 of each half-warp; the
            // other 15 lanes only help to skip through match runs: while the path is in plane M at (i,j), lane d
            // looks at the H tag of the diagonal cell (i-d, j-d), and the run of leading "M" answers (within the
            // block, off the boundaries) is consumed in one go -- the codes along the way are exactly what the
            // one-cell-at-a-time walk would have read.
            {
                const int lo = blk > 0 ? s0 + 1 : 0; // first step whose codes this block serves
                auto load = [&](int i, int j) -> unsigned {
                    const int l = (j - 1) / C, c = (j - 1) - l * C;
                    const int u = (i - 1) + l - s0;
                    const int q = c >= 5 ? 1 : 0, cc = c - 5 * q;
                    return (s_tr[(u * WPL + q) * 32 + half * LPP + l] >> (32 - kTagBits * (5 - cc))) & (kScale - 1);
                };
                unsigned cur = 0;
                bool walking = !done; // meaningful in lane 0; broadcast below
                if (lane == 0 && walking) {
                    if (wi > 0 && wj > 0)
                        cur = load(wi, wj);
                    else
                        cur = (wi == 0) ? (wj == 0 ? code00 : 0x15u) : 0x00u;
                    if (need_k)
                        wk = 2 - (int)((cur >> 4) & 3u);
                    need_k = 0;
                }
                while (true) {
                    // half-uniform view of lane 0's state
                    const int bi = __shfl_sync(FULL, wi, 0, LPP), bj = __shfl_sync(FULL, wj, 0, LPP);
                    const int bk = __shfl_sync(FULL, wk, 0, LPP);
                    const bool bw = __shfl_sync(FULL, (int)(walking && (wi > 0 || wj > 0)), 0, LPP) != 0;
                    if (!__any_sync(FULL, bw))
                        break;
                    // match-run look-ahead: cell (bi - d, bj - d), d = lane + 1
                    int skip = 0;
                    {
                        const int d = lane + 1, ci = bi - d, cj = bj - d;
                        bool isM = false;
                        if (bw && bk == 0 && ci > 0 && cj > 0 && (ci - 1) + (cj - 1) / C >= lo)
                            isM = ((load(ci, cj) >> 4) & 3u) == 2u; // H tag 2 = plane M
                        const unsigned ball = __ballot_sync(FULL, isM);
                        const unsigned mine = (ball >> (half * LPP)) & 0xffffu;
                        skip = __ffs(~mine) - 1; // leading diagonal cells that continue the match run (0..16)
#ifdef GNX_CK_NOSKIP
                        skip = 0;
#endif
                    }
                    if (lane == 0 && bw) {
                        if (skip >= 2) {
                            // consume skip - 1 cells of the run at once: the walk stands on (wi,wj) in plane M, the next
                            // skip cells are M as well; stop ON the last of them so that the normal step below reads
                            // its code
                            const int adv = skip - 1;
                            if (cur_op == 0) {
                                run += adv;
                            } else {
                                if (run > 0)
                                    emit(cur_op, run);
                                cur_op = 0;
                                run = adv;
                            }
                            wi -= adv;
                            wj -= adv;
                            cur = load(wi, wj);
                        }
                        // standing on the boundary: column 0 is plane D up to (0,0), row 0 plane I (the pseudo-codes of
                        // affineTrace's boundary cells point to themselves) -- the rest of the route is one run
                        const bool col0 = wj == 0 && wk == 2, row0 = wi == 0 && wk == 1;
                        if (col0 || row0) {
                            const int len = col0 ? wi : wj;
                            if (wk == cur_op) {
                                run += len;
                            } else {
                                if (run > 0)
                                    emit(cur_op, run);
                                cur_op = wk;
                                run = len;
                            }
                            emit(cur_op, run);
                            if (Q.pass == 0)
                                Q.counts[idx] = cnt;
                            wi = wj = 0;
                            done = true;
                            walking = false;
                            continue;
                        }
                        // one ordinary step of affineTrace
                        if (wk == cur_op) {
                            ++run;
                        } else {
                            if (run > 0)
                                emit(cur_op, run);
                            cur_op = wk;
                            run = 1;
                        }
                        const int kn = 2 - (int)((cur >> (wk == 1 ? 0 : 2)) & 3u);
                        wi -= (wk != 1);
                        wj -= (wk != 2);
                        unsigned nw = 0;
                        bool leave = false;
                        if (wi > 0 && wj > 0) {
                            if ((wi - 1) + (wj - 1) / C < lo) { // the path leaves this block: an earlier one is needed
                                need_k = (wk == 0);
                                if (!need_k)
                                    wk = kn;
                                leave = true;
                            } else {
                                nw = load(wi, wj);
                            }
                        } else {
                            nw = (wi == 0) ? (wj == 0 ? code00 : 0x15u) : 0x00u;
                        }
                        if (leave) {
                            walking = false;
                        } else {
                            wk = (wk == 0) ? 2 - (int)((nw >> 4) & 3u) : kn;
                            cur = nw;
                            if (wi == 0 && wj == 0) {
                                if (run > 0)
                                    emit(cur_op, run);
                                if (Q.pass == 0)
                                    Q.counts[idx] = cnt;
                                done = true;
                                walking = false;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
      }
    }
}

} // namespace gnx

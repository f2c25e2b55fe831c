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
#ifdef GNX_CK_DEBUG
#include <cstdio>
#endif

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
    int *work_count;           // device counter of `work` (ckpt_classify_kernel / ckpt_overflow_list_kernel)
    int *work_next;            // device cursor: the next entry of `work` to hand to a half-warp
    // ragged batches: where pass 1 put each pair (quad * 4 + slot) and each quad's checkpoint words
    const int *pair_slot;      // per pair in chunk, or nullptr (uniform batch: the pair's own index)
    const int64_t *quad_ck_off;
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
    // 2-bit batches (P.alpha_words set): the bases are read from the dnaTwoBit words, there is no byte copy on the device
    const bool tbm = P.alpha_words != nullptr;
    const int64_t a0 = tbm ? 0 : P.alpha_off[pair], b0 = tbm ? 0 : P.beta_off[pair];
    const int n = tbm ? P.n_uni : (int)(P.alpha_off[pair + 1] - a0), m = tbm ? P.m_uni : (int)(P.beta_off[pair + 1] - b0);
    const int rs = (int)Q.rstar[pair];
    bool shortcut = false;
    if (rs >= m && m >= 1) {
        long long sum = 0;
        if (tbm) {
            const uint32_t *wa = reinterpret_cast<const uint32_t *>(P.alpha_words + pair * P.wn);
            const uint32_t *wb = reinterpret_cast<const uint32_t *>(P.beta_words + pair * P.wm);
            for (int k = 0; k < m; ++k)
                sum += P.scores[tb_base(wa, rs - m + k) * P.dim + tb_base(wb, k)];
        } else {
            const uint8_t *__restrict__ al = P.alpha + a0 + (rs - m);
            const uint8_t *__restrict__ be = P.beta + b0;
            for (int k = 0; k < m; ++k)
                sum += P.scores[(int)al[k] * P.dim + (int)be[k]];
        }
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

// Work list of pass 1 (pairs whose cigar overflowed the slot and still fits the output): built on the device.
__global__ void __launch_bounds__(128) ckpt_overflow_list_kernel(const int *counts, const int64_t *cigar_off, int64_t np, int slot_cap,
                                                                 int64_t out_cap, int *work, int *work_count)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < np && counts[idx] > slot_cap && cigar_off[idx] + counts[idx] <= out_cap)
        work[atomicAdd(work_count, 1)] = (int)idx;
}

// Each HALF-warp works on one queued pair and takes the next one from the work list as soon as its pair is settled
// (Q.work_next is a device cursor): routes need anything from one block (an indel near the read's end, then the
// ungapped-tail shortcut) to all sixteen (unrelated sequences), and with static pairing a warp cost the maximum of its
// two pairs.  The batch is uniform (n x m), so everything but the query tables, the staged target, r* and the
// checkpoint address is kernel-invariant.
#ifndef GNX_CK_MINB
#define GNX_CK_MINB 12
#endif
__global__ void __launch_bounds__(32, GNX_CK_MINB) affine_ckpt_trace_kernel(const FillParams P, const CkptParams Q)
{
    constexpr int C = 10, LPP = 16, WPL = 2;
    // SK rows between neighbouring lanes, as in pass 1: cell (i,j) belongs to step (i - 1) + SK * ((j - 1) / C).  A
    // restarted block lacks the tags of D and of the incoming I for its first SK steps (pass 1 saved clean values; the
    // neighbour's edge a lane reads was produced SK steps earlier), so block b > 0 serves steps 32b + SK .. 32b + SK + 31.
    constexpr int SK = kCkSkew;
    constexpr int SC = kScale, FI = kFI, FD = kFD, FH = kFH;
    constexpr int NEG = kNeg32, CLR = ~(kScale - 1);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kTgtPitch = kRing + 64;
    __shared__ int s_tab[C * kDimP * 32];
    __shared__ uint8_t s_tgt[2 * kTgtPitch];
    __shared__ uint32_t s_tr[(kCkK + SK) * WPL * 32];
    // per half: prefix sums of the walker's diagonal (route shortcuts).  They live in the trace-code area: the shortcut
    // tests of an iteration are over before its recompute writes the codes, and 18.1 KB per CTA instead of 19.3 lets a
    // twelfth warp onto the SM.
    static_assert(sizeof(int) * 2 * (LPP * C + 1) <= sizeof(s_tr), "s_pre must fit in s_tr");
    int *const s_pre = reinterpret_cast<int *>(s_tr);
    const int tid = threadIdx.x, lane = tid % LPP, half = tid / LPP;
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_s = (O + E) * SC, e_s = E * SC;
    const int kI = oe_s + 2 * FI - 2 * FH;
    const int iI = e_s + FI, iD = oe_s;
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;
    const int fh_reg = FH * one;
    const int n_work = *Q.work_count;
    if (n_work == 0)
        return;
    // lengths of this half's pair and everything derived from them (set when the pair is taken)
    int n = 0, m = 0, T = 0;
    const int jbase = lane * C;
    int aM[C], aI[C], aD[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        aM[c] = dMn;
        aI[c] = dIn;
        aD[c] = dDn;
    }
    const uint8_t *tg = s_tgt + half * kTgtPitch;
    const unsigned code00 = (unsigned)(2 - Q.h00_plane) << 4;

    // ---- this half's pair ----
    bool have = false, exhausted = false, tested = false;
    int64_t idx = 0, pair = P.pair_begin, quad = 0;
    const uint32_t *ck_base = Q.ckpt; // the pair's quad's checkpoint words
    int src_lane = lane, sel = 0, rs = 0, S_pair = 0;
    // walk state, kept by lane 0 of each half-warp
    int wi = 0, wj = 0, wk = 2, need_k = 1, cur_op = 2, run = 0, cnt = 0, total = 0;
    bool done = true;
    uint32_t *slot = Q.slots;
    CigarOut *dst = nullptr;
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

    while (true) {
        bool all_idle = false;
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
            // ---- a half without a pair takes the next queued one ----
            int w = -1;
            if (lane == 0 && !have && !exhausted)
                w = atomicAdd(Q.work_next, 1);
            w = __shfl_sync(FULL, w, 0, LPP);
            bool fresh = false;
            if (!have && !exhausted) {
                if (w < n_work)
                    fresh = true;
                else
                    exhausted = true;
            }
            if (!__any_sync(FULL, have || fresh)) {
                all_idle = true;
                break;
            }
            if (__any_sync(FULL, fresh)) {
                if (fresh) {
                    // chunk-local index idx; its checkpoint words: quad idx / 4, lanes 16 * ((idx % 4) / 2) .. + 15 of each
                    // record, 16-bit half idx % 2
                    idx = Q.work[w];
                    pair = P.pair_begin + idx;
                    const int64_t slot_id = Q.pair_slot ? (int64_t)Q.pair_slot[idx] : idx; // where pass 1 held this pair
                    quad = slot_id >> 2;
                    src_lane = (int)(((slot_id >> 1) & 1) * LPP) + lane;
                    sel = (int)(slot_id & 1);
                    ck_base = Q.ckpt + (Q.quad_ck_off ? (size_t)Q.quad_ck_off[quad] : (size_t)quad * Q.quad_words);
                    const bool tbm = P.alpha_words != nullptr; // 2-bit batch: bases come from the dnaTwoBit words
                    const uint8_t *__restrict__ alpha = tbm ? nullptr : P.alpha + P.alpha_off[pair];
                    const uint8_t *__restrict__ beta = tbm ? nullptr : P.beta + P.beta_off[pair];
                    const uint32_t *wa = tbm ? reinterpret_cast<const uint32_t *>(P.alpha_words + pair * P.wn) : nullptr;
                    const uint32_t *wb = tbm ? reinterpret_cast<const uint32_t *>(P.beta_words + pair * P.wm) : nullptr;
                    n = tbm ? P.n_uni : (int)(P.alpha_off[pair + 1] - P.alpha_off[pair]);
                    m = tbm ? P.m_uni : (int)(P.beta_off[pair + 1] - P.beta_off[pair]);
                    T = n + SK * (LPP - 1);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const bool last = jbase + c + 1 == m;
                        aM[c] = last ? dMl : dMn;
                        aI[c] = last ? dIl : dIn;
                        aD[c] = last ? dDl : dDn;
                    }
                    // per-lane score tables and the staged target: exactly affine_fill3_kernel's set-up
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = jbase + c + 1;
                        const int q = (j <= m) ? (tbm ? tb_base(wb, j - 1) : (int)beta[j - 1]) : 0;
#pragma unroll
                        for (int a = 0; a < kDimP; ++a) {
                            int v = 0;
                            if (a < P.dim && q < P.dim)
                                v = P.scores[a * P.dim + q] * SC + 2 * FH;
                            s_tab[(c * kDimP + a) * 32 + tid] = v;
                        }
                    }
                    if (tbm) { // 16 bases (one 32-bit half word) per lane and iteration
                        for (int h = lane; 16 * h < n; h += LPP) {
                            const uint32_t v = wa[((h >> 1) << 1) | ((h & 1) ^ 1)];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (16 * h + i < n)
                                    s_tgt[half * kTgtPitch + 16 * h + i] = (uint8_t)((v >> (30 - 2 * i)) & 3u);
                        }
                    } else {
                        for (int i = lane; i < n; i += LPP)
                            s_tgt[half * kTgtPitch + i] = alpha[i];
                    }
                    rs = (int)Q.rstar[pair];
                    S_pair = (int)P.out_score[pair];
                    wi = rs;            // current cell
                    wj = m;
                    wk = 2;             // current plane (0 M, 1 I, 2 D)
                    need_k = 1;         // the plane of the current cell is the H tag of its code
                    cur_op = 2;         // the free-end column's D run from (n,m) down to (r*,m)
                    run = n - rs;
                    cnt = 0;
                    done = false;
                    slot = Q.slots + (size_t)idx * Q.slot_cap;
                    if (Q.pass == 1) {
                        total = Q.counts[idx];
                        dst = Q.out_cigar + Q.cigar_off[idx];
                    }
                    have = true;
                    tested = false;
                }
                __syncwarp();
            }
            // ---- route shortcuts that need no trace code (pass 0) ----------------------------------------------
            // The walker stands on cell (i,j) and has to learn its plane from the cell's H tag (need_k): either the
            // fresh pair at (r*, m), or a walk that has just left a block through a diagonal (M) step.  Its value
            // V = H(i,j) = max(M,I,D)(i,j) is known without any recompute: V = S - (cost of the route walked so far).
            //  (a) ungapped tail.  If V equals U(i,j), the score of the ungapped diagonal from column 0 (free
            //      D(i-j,0) = 0) to (i,j), then M(i,j) >= U = V forces M(i,j) = H(i,j), and by the argument of
            //      ckpt_classify_kernel no I or D value can beat M anywhere on that diagonal: tripleMaxTrace picks M
            //      at every cell down to column 0.  The rest of the route is M x j, D x (i - j).
            //  (b) checkpoint-verified diagonal jump.  Pass 1 saved H(32k - l, 10l + c + 1) for every lane l and
            //      column c at every checkpoint k (skew 2: H(32k - 2l, ...)); the walker's diagonal meets at most one such
            //      cell c_q per checkpoint ((10 + SK) l + c = j - i + 32k - 1).  If V = H(c_q) + (substitution scores of the d diagonal
            //      cells above c_q), then M(i,j) >= s + H(i-1,j-1) >= ... >= sum + H(c_q) = V >= M(i,j): every
            //      inequality is tight, so M = H at (i,j) and at each of the d - 1 cells in between, i.e. the route is
            //      M x d (ties prefer M) and lands on c_q with its plane still to be read (need_k).  The farthest
            //      verified cell is taken: the blocks in between are never recomputed.
            // Reads carry few indels: a typical route is settled by one or two recomputed blocks around each indel.
#ifndef GNX_CK_NOTAIL
            {
                const int ti = __shfl_sync(FULL, wi, 0, LPP), tj = __shfl_sync(FULL, wj, 0, LPP);
                const int tcnt = __shfl_sync(FULL, cnt, 0, LPP), trun = __shfl_sync(FULL, run, 0, LPP);
                const int tcur = __shfl_sync(FULL, cur_op, 0, LPP);
                const int tf = __shfl_sync(FULL, (int)(!done && need_k), 0, LPP);
                const bool initial = tcnt == 0 && tcur == 2;   // nothing walked yet: V = S at (r*, m)
                const bool tri = Q.pass == 0 && have && !tested && tf != 0 && tj > 0 && ti > 0 && tcnt < Q.slot_cap &&
                                 (initial || tcur == 0);
                if (__any_sync(FULL, tri)) {
                    auto sc = [&](int row, int col) { // substitution score of cell (row, col) from the per-lane tables
                        const int a = tg[row - 1], l = (col - 1) / C, c = (col - 1) - l * C;
                        return (s_tab[(c * kDimP + a) * 32 + half * LPP + l] - 2 * FH) >> kTagBits;
                    };
                    int acc = 0, gap = 0, ci = rs, cj = m;
                    const int e0 = (n - rs > 0) ? 1 : 0; // slot[0] is the free-end column's D run (cost 0)
                    const int tc = initial ? -1 : tcnt;  // index of the run in progress (an M run), -1: none
                    const int emax = max(tc, __shfl_xor_sync(FULL, tc, 16));
                    for (int e = 0; e <= emax; ++e) { // warp-uniform trip count: the shuffle below needs all 32 lanes
                        uint32_t v = 0;
                        if (lane == 0 && tri && e >= e0 && e < tc)
                            v = slot[e];
                        v = __shfl_sync(FULL, v, 0, LPP);
                        if (e == tc)
                            v = (uint32_t)trun << 2; // the M run in progress
                        if (!tri || e < e0 || e > tc)
                            continue;
                        const int op = (int)(v & 3u), len = (int)(v >> 2);
                        if (op == 0) {
                            for (int t = lane; t < len; t += LPP)
                                acc += sc(ci - t, cj - t);
                            ci -= len;
                            cj -= len;
                        } else {
                            gap += O + len * E; // a gap run of len cells: open once, extend len times
                            if (op == 1)
                                cj -= len;
                            else
                                ci -= len;
                        }
                    }
#pragma unroll
                    for (int o = LPP / 2; o > 0; o >>= 1)
                        acc += __shfl_xor_sync(FULL, acc, o);
                    const int V = S_pair - gap - acc;
                    const bool ok = tri && ci == ti && cj == tj;
                    // prefix sums of the walker's diagonal: pre[d] = sum of s(i - t, j - t), t < d; lane l owns t in [10l, 10l+10)
                    const int L = ok ? min(ti, tj) : 0;
                    int part[C], mine = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int t = lane * C + c;
                        part[c] = mine;
                        if (t < L)
                            mine += sc(ti - t, tj - t);
                    }
                    int incl = mine; // inclusive scan of the lanes' chunk sums
#pragma unroll
                    for (int o = 1; o < LPP; o <<= 1) {
                        const int up = __shfl_up_sync(FULL, incl, o, LPP);
                        if (lane >= o)
                            incl += up;
                    }
                    const int base = incl - mine;
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        s_pre[half * (LPP * C + 1) + lane * C + c] = base + part[c];
                    if (lane == LPP - 1)
                        s_pre[half * (LPP * C + 1) + LPP * C] = incl;
                    __syncwarp();
                    const int *pre = s_pre + half * (LPP * C + 1);
                    // (a) ungapped tail
                    const bool tail = ok && ti >= tj && V == pre[tj];
                    // (b) lane q looks at checkpoint k = q + 1
                    int dhit = 0;
                    if (ok && !tail) {
                        const int k = lane + 1;
                        const int X = tj - ti + kCkK * k - 1; // = (C + SK) l + c for the cell H(32k - SK l, C l + c + 1) on the walker's diagonal
                        if (kCkK * k < T && X >= 0) {
                            const int l2 = X / (C + SK), c2 = X - (C + SK) * l2;
                            const int col = l2 * C + c2 + 1, d = tj - col;
                            if (c2 < C && l2 < LPP && col <= m && d >= 1 && d <= L - 1) { // the cell itself is interior (row, col >= 1)
                                const uint32_t x = __ldg(ck_base + (size_t)(k - 1) * (kCkRegs * 32) + c2 * 32 + (src_lane - lane + l2));
                                const int hck = (int)((x >> (16 * sel)) & 0xffffu) - 32768;
                                if (V - pre[d] == hck)
                                    dhit = d;
                            }
                        }
                    }
#pragma unroll
                    for (int o = LPP / 2; o > 0; o >>= 1)
                        dhit = max(dhit, __shfl_xor_sync(FULL, dhit, o)); // the farthest verified cell of this half
                    __syncwarp(); // s_pre shares its storage with s_tr: its reads are over before this iteration's recompute writes codes
                    if (lane == 0 && (tail || dhit > 0)) {
                        const int adv = tail ? tj : dhit;
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
                        if (tail) {
                            emit(0, run);
                            if (wi > 0)
                                emit(2, wi);
                            Q.counts[idx] = cnt;
                            wi = wj = 0;
                            done = true;
                        }
                    }
                }
                if (tri)
                    tested = true; // one attempt per standing cell
            }
#endif
            // a pair settled by the tail shortcut frees its half at once: it is refilled in the second round
            if (__shfl_sync(FULL, (int)done, 0, LPP) != 0)
                have = false;
        }
        if (all_idle)
            break;
        {
            // block each half needs: the one serving the step of its current cell
            int blk = 0;
            {
                const int tcell = (wi - 1) + SK * ((wj - 1) / C);
                blk = (wi > 0 && wj > 0) ? max(tcell - SK, 0) / kCkK : -1; // -1: only boundary cells remain
            }
            blk = __shfl_sync(FULL, blk, 0, LPP);
            const int done0 = __shfl_sync(FULL, (int)done, 0, LPP); // executed by all 32 lanes (no short-circuit around it)
            const bool hdone = !have || done0 != 0;
            const bool recompute = !hdone && blk >= 0;
            const int s0 = blk > 0 ? blk * kCkK : 0;
            // the walk enters the block at its current cell and only moves to earlier steps: no later step is needed
            int tin = (wi - 1) + SK * ((wj - 1) / C);
            tin = __shfl_sync(FULL, tin, 0, LPP);
            const int ulast_h = recompute ? min(tin - s0, kCkK + SK - 1) : 0;
            const int ulast = max(ulast_h, __shfl_xor_sync(FULL, ulast_h, 16));
#ifdef GNX_CK_DEBUG
            if (lane == 0 && recompute)
                printf("pair %d block %d at (%d,%d) need_k %d cur_op %d\n", (int)idx, blk, wi, wj, need_k, cur_op);
#endif
            if (__any_sync(FULL, recompute)) {
                // ---- restore the state entering step s0 ----
                int Hc[C], Dt[C];
                int hpL, edgeI = 0, edgeH = 0, edgeIp = 0, edgeHp = 0;
                if (blk > 0 && recompute) {
                    const uint32_t *src = ck_base + (size_t)(blk - 1) * (kCkRegs * 32) + src_lane;
                    auto val = [&](uint32_t x) { return ((int)((x >> (16 * sel)) & 0xffffu) - 32768) * SC; };
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        Hc[c] = val(__ldg(src + c * 32));
                        Dt[c] = val(__ldg(src + (C + c) * 32));
                    }
                    hpL = val(__ldg(src + 20 * 32));
                    edgeI = val(__ldg(src + 21 * 32));
                    edgeH = val(__ldg(src + 22 * 32));
                    if (SK == 2) {
                        edgeIp = val(__ldg(src + 23 * 32));
                        edgeHp = val(__ldg(src + 24 * 32));
                    }
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
                    const int r = t - SK * lane + 1;
                    int inI = __shfl_up_sync(FULL, SK == 2 ? edgeIp : edgeI, 1, LPP);
                    int inH = __shfl_up_sync(FULL, SK == 2 ? edgeHp : edgeH, 1, LPP);
                    if (SK == 2) { // every step, active or not: the neighbour reads the edge of two steps ago
                        edgeIp = edgeI;
                        edgeHp = edgeH;
                    }
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
                const int lo = blk > 0 ? s0 + SK : 0; // first step whose codes this block serves
                auto load = [&](int i, int j) -> unsigned {
                    const int l = (j - 1) / C, c = (j - 1) - l * C;
                    const int u = (i - 1) + SK * l - s0;
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
                        if (bw && bk == 0 && ci > 0 && cj > 0 && (ci - 1) + SK * ((cj - 1) / C) >= lo)
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
                            if ((wi - 1) + SK * ((wj - 1) / C) < lo) { // the path leaves this block: an earlier one is needed
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
        // a settled pair frees its half for the next queued one
        if (__shfl_sync(FULL, (int)done, 0, LPP) != 0)
            have = false;
        tested = false; // the walker has moved
    }
}

} // namespace gnx

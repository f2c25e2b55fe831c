// gnx_fill16.cuh -- packed 16-bit score-only affine fill: FOUR pairs per warp.
//
// Same wavefront as affine_fill3_kernel<C=10, LPP=16> (two half-warps, 10 columns per lane), but every
// 32-bit register carries the same DP cell of TWO different pairs: pair "A" in bits 15:0 and pair "B"
// in bits 31:16.  The three maxes of a cell become one VIMNMX3.U16x2 and two VIADDMNMX.U16x2 (DPX, ALU
// pipe, same issue cost as their 32-bit forms -- profiles/r01_microbench_pipes.txt), i.e. half the ALU
// work per cell.  Halves are kept BIASED-UNSIGNED (u = v + 32768) so that
//   * maxes are the unsigned 16x2 forms, and
//   * the plain adds (M = H_diag + s, H + O + E) stay ordinary 32-bit integer adds on the FMA pipe:
//     X = uB*65536 + uA, so X + (dB*65536 + dA) = (uB+dB)*65536 + (uA+dA) exactly as long as each half
//     stays inside [0, 65535] -- which the host proves before dispatching this kernel (fill16_ok()).
// Score-only needs no -inf at all: row 0 is I(0,j), column 0 is D(i,0) and every other state is a max
// that contains a finite candidate (DESIGN.md "Arithmetic width"), so there is no sentinel to keep
// away from the range limits.  Padding columns (j > m) get zero substitution scores, which keeps their
// (unused) values within one gap-open of real cells.
// Requires: m <= 160, dim <= 5, O <= 0, the range proof of fill16_ok(); pairs in quads that share the target length
// (a uniform batch, or the host-binned quads of a ragged one); score only, or -- CKPT -- score + checkpoints + r*.
// Template parameters: FREE (freeEndGaps), CM (compile-time last-column index), CKPT, TB (dnaTwoBit words by TMA and
// the joint score table); SK / quad cursor: see below.
#pragma once
#include "gnx_fill3.cuh"

namespace gnx {

__device__ __forceinline__ unsigned pack16(int v) { return (unsigned)(v + 32768) * 65537u; } // same v in both halves
__device__ __forceinline__ unsigned wrap16x2(int d) { return ((unsigned)d & 0xffffu) * 65537u; } // per-half wrapping addend
__device__ __forceinline__ unsigned umax3_16x2(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned uaddmax_16x2(unsigned a, unsigned b, unsigned c) { return __viaddmax_u16x2(a, b, c); }

// CM >= 0 (FREE only): (m - 1) % C, the in-lane index of the freeEndGaps column, known at compile time -- only that
// column then pays the extra add for its zero D-plane addends; every other cell shares H + O + E between I' and D'.
//
// CKPT (needs FREE and CM >= 0): the first pass of the checkpoint-and-recompute traceback (affine_ckpt_trace_kernel
// below).  Every kCkK steps the warp saves its 23 state registers per lane (Hc[10], Dt[10], hpL, edgeI, edgeH --
// the complete wavefront state entering step kCkK*k) and the lane that owns the free-end column tracks
//     r* = the LAST row r whose max(M(r,m), I(r,m)) equals the running maximum of the column (>= I(0,m)),
// which is where the reference's traceback leaves the free-end column: walking up from (n,m) in plane D, the
// source of D(i,m) is T(M, I, D)(i-1,m) with ties M >= I >= D (align/align.go:76-84), D(i,m) being the running
// maximum itself, so the walk stays in D exactly until the last row that reached it.
#ifndef GNX_F16_SKEW
#define GNX_F16_SKEW 2      // rows between neighbouring lanes of the freeEndGaps kernels (see SK below)
#endif
constexpr int kCkK = 32;    // steps between checkpoints
// The checkpoint path keeps skew 1: measured with skew 2 (parity-clean, 88 tests) pass 1 gains 1 % (its steady phase
// is cut into 32-step runs by the saves) while affine_ckpt_trace_kernel loses 15 % (a route that crosses the 15 lanes
// spans 30 steps instead of 15, i.e. one more 32-step block to recompute per pair).
#ifndef GNX_CK_SKEW
#define GNX_CK_SKEW 1
#endif
constexpr int kCkSkew = GNX_CK_SKEW;                // geometry shared by pass 1 (here) and affine_ckpt_trace_kernel
constexpr int kCkRegs = kCkSkew == 2 ? 25 : 23;     // 32-bit words per lane per checkpoint (skew 2: + the edges of the step before)
// checkpoints of a pair with target length n: the states entering steps 32, 64, ... < n + kCkSkew * 15
__host__ __device__ inline int64_t ck_count(int64_t n) { return (n + kCkSkew * 15 - 1) / kCkK; }

// ---- 1-D TMA (cp.async.bulk) staging of 2-bit packed sequences: TB kernels ----------------------------------
// The async proxy writes the packed words of a quad's four targets and four queries into shared memory and
// signals an mbarrier with the byte count (SASS: UBLKCP + SYNCS); nothing but one elected lane touches the copy.
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

constexpr int kTbMaxN = 512;                 // TB kernels unpack the quad's targets into 4 x 512 bytes of shared memory
constexpr int kTbMaxWn = kTbMaxN / 32, kTbMaxWm = 5; // words per target / query (m <= 160)
// base `pos` of a dnaTwoBit sequence stored as uint64 words (dna/dnaTwoBit/dnaTwoBit.go:59-65 GetBase: first base
// in bits 63:62), read through a little-endian 32-bit view of the words
__device__ __forceinline__ int tb_base(const uint32_t *w32, int pos)
{
    const uint32_t v = w32[((pos >> 5) << 1) | (((pos >> 4) & 1) ^ 1)];
    return (int)((v >> (30 - 2 * (pos & 15))) & 3u);
}

// TB: the batch's bases arrive as dnaTwoBit words (P.alpha_words / P.beta_words, uniform lengths P.n_uni x P.m_uni,
// sequence p at word p * wn resp. p * wm): the quad's words are staged by TMA one quad ahead, the targets expanded to
// one byte per base in shared memory (the per-step base fetch is then an LDS), the queries read straight from the
// packed words while the score tables are built.  Bases are 0..3, so the tables have four rows.
//
// TB kernels keep ONE joint table per thread: entry [c][4 * aA + aB] = s(aA, qA_c) + 65536 * s(aB, qB_c), so a packed
// cell costs one LDS (the kernel is bound by shared-memory wavefronts: ncu, profiles/r02i_ckpt_path.md) and the row
// base of a step is one byte of s_ab = 4 * targetA[r] + targetB[r].  16 x 10 x 32 words = 20 KB per warp: 10 warps per
// SM instead of 16.
#ifndef GNX_TB_JOINT
#define GNX_TB_JOINT 1
#endif
//
// SK = rows between neighbouring lanes.  With SK = 1 the chain I(r,j) -> H -> H+O+E -> I(r,j+1) of a step must finish
// before the next step can start (its first cell needs the neighbour's edge of the step before), so a warp exposes one
// dependent chain and the kernel stalls on fixed-latency waits (ncu source page, profiles/r02l_fill16_joint.md).  With
// SK = 2 the neighbour's edge was produced two steps ago: consecutive steps are independent chains one cell apart and
// the unrolled loop interleaves them.  n + 2 * 15 steps instead of n + 15, so the ramps (lanes that have not reached
// row 1 yet / are past row n) must not cost more than a steady step: they run the SAME branch-free cell on every lane,
// made harmless by data --
//   * before row 1 a lane processes "virtual rows" whose M term is masked off: the row-0 boundary it was initialised
//     with (H = I = O + jE, D(next) = H + O + E, its edges likewise) is a fixed point of such a row when the first
//     column is free (H(r,0) = 0), so the lane still holds exactly that boundary when row 1 arrives;
//   * past row n a lane keeps computing on a clamped target index, after H(n, CM) has been copied aside.
// Used by the freeEndGaps kernels with a compile-time last column (score-only and CKPT; affine_ckpt_trace_kernel
// re-runs its blocks in the same geometry, kCkSkew).
// SHIFTED (GNX_F16_SHORT_CHAIN): the one-op chain.  With Hs = H + O + E kept instead of H (row boundary, diagonal and
// edge included; a uniform shift of every H), a cell is
//     Y   = max(D + O + E, Hs(r-1,j-1) + s)       VIADDMNMX      (no I in it)
//     I'  = max(I + E, Y)                         VIADDMNMX      <- the only op on the chain I(r,j) -> I(r,j+1)
//     Hs  = max(I + O + E, Y)                     VIADDMNMX      (= H + O + E: I + O + E <= I + E changes no maximum)
//     D'  = max(D + E, Hs)                        VIADDMNMX
// four DPX ops instead of VIMNMX3 + IMAD + 2 VIADDMNMX, but the next cell of the row waits for ONE of them instead
// of three in sequence: the kernel was stalled on fixed-latency waits, not on the ALU pipe (profiles/r02l_*.md).
// I + O + E and D + O + E need one more O + E of head room than the unshifted form: fill16_ok() proves it.
// Measured SLOWER (C2 5.03 against 5.35 TCUPS): the cell is paced by the DPX pipe once enough rows are in flight
// (tools/microbench2.cu: 7 SMSP cycles per three-op cell, 10 per four-op cell), so the default stays 0.
#ifndef GNX_F16_SHORT_CHAIN
#define GNX_F16_SHORT_CHAIN 0
#endif
template <bool FREE, int CM = -1, bool CKPT = false, bool TB = false>
__global__ void __launch_bounds__(32, (TB && GNX_TB_JOINT) ? 10 : 16) affine_fill16_kernel(const FillParams P)
{
    constexpr bool JT = TB && GNX_TB_JOINT;
    constexpr int SK = (FREE && CM >= 0) ? (CKPT ? kCkSkew : GNX_F16_SKEW) : 1;
    constexpr bool SH = GNX_F16_SHORT_CHAIN != 0;
    static_assert(CM < 0 || FREE, "CM selects the free-end column");
    static_assert(!CKPT || (FREE && CM >= 0), "checkpoints are taken on the freeEndGaps path only");
    constexpr int C = 10, LPP = 16;
    constexpr int ROWS = TB ? 4 : kDimP;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ int s_tabA[JT ? 1 : C * ROWS * 32]; // [c][a][thread], pair A: s sign-extended (a 16-bit LDS costs two
                                                   // shared-memory wavefronts: ncu counted 30 per step instead of 20)
    __shared__ int s_tabB[JT ? 1 : C * ROWS * 32]; // [c][a][thread], pair B: s * 65536
    __shared__ int s_tabJ[JT ? C * 16 * 32 : 1];   // [c][4 aA + aB][thread]: sA + 65536 sB
    __shared__ __align__(16) uint64_t s_pk[TB ? 4 * (kTbMaxWn + kTbMaxWm) : 1]; // TMA landing zone: targets, then queries
    // the quad's targets, one byte per base; JT: per half-warp, 4 * (base of target A) + (base of target B)
    __shared__ __align__(16) uint8_t s_tg[TB ? (JT ? 2 : 4) * kTbMaxN : 16];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int lane = tid % LPP, half = tid / LPP;
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_i = (O + E) * 65537;          // integer addend: +O+E in both halves
    const unsigned e_w = wrap16x2(E);          // per-half wrapping addend for VIADDMNMX.U16x2
    const unsigned oe_w = wrap16x2(O + E);
    const int hs = SH ? O + E : 0;             // the shift of every stored H
    // Quads: four consecutive pairs of a uniform batch, or -- RAGGED batches -- the rows quad_first .. quad_first +
    // n_quads - 1 of P.quad_pairs: four chunk-local pair indices (-1 = empty slot) that the host binned so that the
    // quad's pairs share the target length n and, with free end gaps, the in-lane index CM of the last query column;
    // query lengths may differ inside a quad (columns past a pair's m are padding that scores 0).
    const bool binned = !TB && P.quad_pairs != nullptr;
    const int64_t n_quads = binned ? P.n_quads : (P.pair_end - P.pair_begin + 3) / 4;
    // TB: one quad's packed words = 32 * wn bytes of targets + 32 * wm bytes of queries (always multiples of 16, and
    // 16-byte aligned because chunks start at a quad boundary of a 256-byte aligned buffer)
    const unsigned tb_bytes_t = TB ? 32u * (unsigned)P.wn : 0u, tb_bytes_q = TB ? 32u * (unsigned)P.wm : 0u;
    auto tb_issue = [&](int64_t quad) { // lane 0: stage quad's words (its first pair is pair_begin + 4 * quad)
        const int64_t p0 = P.pair_begin + quad * 4;
        mbar_expect_tx(&s_bar, tb_bytes_t + tb_bytes_q);
        bulk_g2s(s_pk, P.alpha_words + p0 * P.wn, tb_bytes_t, &s_bar);
        bulk_g2s(s_pk + 4 * kTbMaxWn, P.beta_words + p0 * P.wm, tb_bytes_q, &s_bar);
    };
    // Quads are drawn from a device cursor: a CTA is one warp and an SM holds 10 (TB) or 16 of them, i.e. 3+3+2+2 per
    // scheduler, so warps do not run at the same speed and equal static shares would leave two schedulers idle at the end.
    const bool dyn = P.quad_ctr != nullptr;
    auto fetch = [&]() -> int64_t {
        unsigned v = 0;
        if (tid == 0) {
            v = atomicAdd(P.quad_ctr, 1u);
            if (v == (unsigned)n_quads + gridDim.x - 1) // every warp draws exactly one index >= n_quads: this is the last
                atomicExch(P.quad_ctr, 0u);
        }
        return (int64_t)__shfl_sync(FULL, v, 0);
    };
    unsigned tb_phase = 0;
    if (TB && tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    int64_t quad = dyn ? fetch() : (int64_t)blockIdx.x;
    if (TB) {
        __syncwarp();
        if (tid == 0 && quad < n_quads)
            tb_issue(quad);
    }

    while (quad < n_quads) {
        int64_t pA0, pB0, pA, pB; // the half-warp's two pairs (global indices); pX0 < 0 / >= pair_end: empty slot
        if (binned) {
            const int *qp = P.quad_pairs + (P.quad_first + quad) * 4;
            const int iA = qp[half * 2], iB = qp[half * 2 + 1], i0 = qp[0]; // slot 0 always holds a pair
            pA0 = iA >= 0 ? P.pair_begin + iA : P.pair_end;
            pB0 = iB >= 0 ? P.pair_begin + iB : P.pair_end;
            pA = P.pair_begin + (iA >= 0 ? iA : i0);
            pB = P.pair_begin + (iB >= 0 ? iB : i0);
        } else {
            pA0 = P.pair_begin + quad * 4 + half * 2;
            pB0 = pA0 + 1;
            pA = min(pA0, P.pair_end - 1); // tail: recompute a valid pair
            pB = min(pB0, P.pair_end - 1);
        }
        int n, m, mB;
        const uint8_t *__restrict__ alA, *__restrict__ alB, *__restrict__ beA = nullptr, *__restrict__ beB = nullptr;
        const uint32_t *qwA = nullptr, *qwB = nullptr; // TB: 32-bit views of the two queries' packed words
        if (TB) {
            n = P.n_uni;
            m = mB = P.m_uni;
            mbar_wait(&s_bar, tb_phase); // the quad's words have landed
            tb_phase ^= 1;
            const int kA = (int)(pA - (P.pair_begin + quad * 4)), kB = (int)(pB - (P.pair_begin + quad * 4));
            // expand the four targets: 16 bases (one 32-bit half word) -> 16 bytes per iteration
            const uint32_t *pk32 = reinterpret_cast<const uint32_t *>(s_pk);
            const int hw_per = 2 * P.wn;
            if (JT) { // 2 x 16 bases of the half-warp's two targets -> 16 joint bytes per iteration
                const int last = (int)min((int64_t)3, P.pair_end - 1 - (P.pair_begin + quad * 4));
                for (int h = tid; h < 2 * hw_per; h += 32) {
                    const int hf = h / hw_per, hh = h - hf * hw_per;
                    const int k0 = min(2 * hf, last), k1 = min(2 * hf + 1, last);
                    const uint32_t v0 = pk32[(k0 * P.wn + (hh >> 1)) * 2 + ((hh & 1) ^ 1)];
                    const uint32_t v1 = pk32[(k1 * P.wn + (hh >> 1)) * 2 + ((hh & 1) ^ 1)];
                    uint32_t x[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t r0 = (v0 >> (24 - 8 * q)) & 0xffu, r1 = (v1 >> (24 - 8 * q)) & 0xffu;
                        const uint32_t x0 = (r0 >> 6) | (((r0 >> 4) & 3u) << 8) | (((r0 >> 2) & 3u) << 16) | ((r0 & 3u) << 24);
                        const uint32_t x1 = (r1 >> 6) | (((r1 >> 4) & 3u) << 8) | (((r1 >> 2) & 3u) << 16) | ((r1 & 3u) << 24);
                        x[q] = x0 * 4u + x1;
                    }
                    if (16 * hh < kTbMaxN)
                        *reinterpret_cast<uint4 *>(s_tg + hf * kTbMaxN + 16 * hh) = make_uint4(x[0], x[1], x[2], x[3]);
                }
            } else
            for (int h = tid; h < 4 * hw_per; h += 32) {
                const int k = h / hw_per, hh = h - k * hw_per;
                const uint32_t v = pk32[(k * P.wn + (hh >> 1)) * 2 + ((hh & 1) ^ 1)];
                uint32_t x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t r = (v >> (24 - 8 * q)) & 0xffu; // four bases, first in bits 7:6
                    x[q] = (r >> 6) | (((r >> 4) & 3u) << 8) | (((r >> 2) & 3u) << 16) | ((r & 3u) << 24);
                }
                if (16 * hh < kTbMaxN)
                    *reinterpret_cast<uint4 *>(s_tg + k * kTbMaxN + 16 * hh) = make_uint4(x[0], x[1], x[2], x[3]);
            }
            alA = s_tg + (JT ? half : kA) * kTbMaxN;
            alB = JT ? alA : s_tg + kB * kTbMaxN;
            qwA = pk32 + (4 * kTbMaxWn + kA * P.wm) * 2;
            qwB = pk32 + (4 * kTbMaxWn + kB * P.wm) * 2;
        } else {
            const int64_t a0A = P.alpha_off[pA], b0A = P.beta_off[pA];
            n = (int)(P.alpha_off[pA + 1] - a0A);
            m = (int)(P.beta_off[pA + 1] - b0A); // pair A's query length; pair B's may differ in a binned quad
            mB = (int)(P.beta_off[pB + 1] - P.beta_off[pB]);
            alA = P.alpha + a0A;
            alB = P.alpha + P.alpha_off[pB];
            beA = P.beta + b0A;
            beB = P.beta + P.beta_off[pB];
        }
        const int T = n + SK * (LPP - 1);
        const int jbase = lane * C;
        unsigned aD[C];
        int aH[C]; // D-plane addends: (E, O+E) regular, (0, 0) in the freeEndGaps last column
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c + 1;
            const bool realA = j <= m, realB = j <= mB;
            int qA = 0, qB = 0;
            if (realA)
                qA = TB ? tb_base(qwA, j - 1) : (int)beA[j - 1];
            if (realB)
                qB = TB ? tb_base(qwB, j - 1) : (int)beB[j - 1];
            int vAs[ROWS], vBs[ROWS];
#pragma unroll
            for (int a = 0; a < ROWS; ++a) {
                int vA = 0, vB = 0;
                if (a < P.dim) { // padding columns score 0 against everything
                    if (realA)
                        vA = P.scores[a * P.dim + qA];
                    if (realB)
                        vB = P.scores[a * P.dim + qB];
                }
                vAs[a] = vA;
                vBs[a] = vB * 65536;
                if (!JT) {
                    s_tabA[(c * ROWS + a) * 32 + tid] = vA;
                    s_tabB[(c * ROWS + a) * 32 + tid] = vB * 65536;
                }
            }
            if (JT) {
#pragma unroll
                for (int ab = 0; ab < 16; ++ab)
                    s_tabJ[(c * 16 + ab) * 32 + tid] = (int)((unsigned)vAs[(ab >> 2) % ROWS] + (unsigned)vBs[(ab & 3) % ROWS]);
            }
            const bool lastA = FREE && (j == m), lastB = FREE && (j == mB); // per pair: halves of the packed addends
            aD[c] = (lastA ? 0u : (e_w & 0xffffu)) | (lastB ? 0u : (e_w & 0xffff0000u));
            aH[c] = (lastA ? 0 : (O + E)) + (lastB ? 0 : (O + E) * 65536);
        }
        unsigned Dt[C], Hc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = jbase + c + 1;
            const unsigned h0 = pack16(O + j * E); // H(0,j) = I(0,j)
            Hc[c] = pack16(O + j * E + hs);
            Dt[c] = h0 + (unsigned)aH[c];          // D(1,j) = I(0,j) + O + E   (or I(0,m) in the free last column)
        }
        const bool lastlaneA = FREE && lane == (m - 1) / C, lastlaneB = FREE && lane == (mB - 1) / C;
        // addends of column CM in this lane (zero in a pair's free-end column), per 16-bit half
        const unsigned aDl = (lastlaneA ? 0u : (e_w & 0xffffu)) | (lastlaneB ? 0u : (e_w & 0xffff0000u));
        const int aHl = (lastlaneA ? 0 : (O + E)) + (lastlaneB ? 0 : (O + E) * 65536);
        unsigned hpL = (jbase == 0) ? pack16(P.h00 + hs) : pack16(O + jbase * E + hs);
        unsigned edgeI = 0, edgeH = 0, edgeIp = 0, edgeHp = 0; // SK = 2: ...p = the edge of the step before
        unsigned res = 0;                                      // SK = 2: H(n, CM) of the lane, copied aside at row n
        if (SK == 2) { // the row-0 fixed point of the virtual rows: I leaving the lane's last column, H of that column
            edgeI = edgeIp = pack16(O + (jbase + C + 1) * E);
            edgeH = edgeHp = pack16(O + (jbase + C) * E + hs);
        }
        unsigned bI = 0, bH = 0;
        // CKPT: (value << 16 | row) of the last row that reached the free-end column's running maximum, per pair;
        // row 0 stands for the boundary D(1,m) = I(0,m)
        // (SH: the column's values are compared in their shifted form)
        unsigned bestA = (unsigned)(O + m * E + hs + 32768) << 16, bestB = (unsigned)(O + mB * E + hs + 32768) << 16;
        uint32_t *ck = nullptr; // checkpoint words of this quad (edge_stride: words per quad of a uniform batch)
        if (CKPT)
            ck = P.trace + (binned ? (size_t)P.quad_ck_off[P.quad_first + quad] : (size_t)quad * P.edge_stride);
        auto boundary = [&](int r) {
            const int d0 = FREE ? 0 : (O + r * E);
            bI = pack16(d0 + O + E);
            bH = pack16(d0 + hs);
        };
        if (lane == 0)
            boundary(1);
        __syncwarp(); // tables and (TB) expanded targets are complete; the landing zone is free again
        const int64_t next = dyn ? fetch() : quad + gridDim.x;
        if (TB) {
            if (tid == 0 && next < n_quads) { // prefetch the next quad's words behind this quad's fill
                fence_proxy_async();          // our generic-proxy reads of s_pk precede the async-proxy writes
                tb_issue(next);
            }
        }
        int aA_next = 0, aB_next = 0;
        if (lane == 0 || SK == 2) { // (SK = 2: lanes on virtual rows need a valid table row, any one)
            aA_next = alA[0];
            if (!JT)
                aB_next = alB[0];
        }

        // mode 0: every lane is inside rows 1..n-1 (no checks); 1: lanes outside their rows skip the step (SK = 1);
        // 2: ramp step of SK = 2 (branch-free, see above)
        auto step = [&](int t, auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr bool CHECK = MODE == 1, RAMP = MODE == 2;
            const int r = t - SK * lane + 1;
            const unsigned vmask = (RAMP && r < 1) ? 0u : 0xffffffffu;
            unsigned inI = __shfl_up_sync(FULL, SK == 2 ? edgeIp : edgeI, 1, LPP);
            unsigned inH = __shfl_up_sync(FULL, SK == 2 ? edgeHp : edgeH, 1, LPP);
            if (SK == 2) {
                edgeIp = edgeI;
                edgeHp = edgeH;
            }
            if (lane == 0) {
                inI = bI;
                inH = bH;
            }
            const int aA = aA_next, aB = aB_next;
            bool active = true;
            if (CHECK) {
                active = (unsigned)(r - 1) < (unsigned)n;
                if ((unsigned)r < (unsigned)n) {
                    aA_next = alA[r];
                    if (!JT)
                        aB_next = alB[r];
                }
            } else {
                const int rr = RAMP ? min(max(r, 0), n - 1) : r;
                aA_next = alA[rr];
                if (!JT)
                    aB_next = alB[rr];
            }
            if (active) {
                if (!FREE && lane == 0 && r < n)
                    boundary(r + 1);
                const int *rowA = JT ? s_tabJ + aA * 32 + tid : s_tabA + aA * 32 + tid; // JT: aA is the joint index
                const int *rowB = s_tabB + aB * 32 + tid;
                unsigned It = inI, hp = hpL;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int sA = rowA[c * (JT ? 16 : ROWS) * 32];
                    const int sB = JT ? 0 : rowB[c * ROWS * 32]; // s * 65536
                    unsigned MH = hp + (unsigned)sA + (unsigned)sB; // one IADD3 (JT: one IADD)
                    if (RAMP)
                        MH &= vmask; // virtual row: no M term (0 = the smallest biased value)
                    if (SH) { // MH = M + O + E
                        if (CKPT && c == CM) {
                            const unsigned mx = uaddmax_16x2(It, oe_w, MH); // max(M, I) + O + E
                            bestA = max(bestA, __byte_perm((unsigned)r, mx, 0x5410));
                            bestB = max(bestB, __byte_perm((unsigned)r, mx, 0x7610));
                        }
                        const unsigned Y = uaddmax_16x2(Dt[c], oe_w, MH);
                        const unsigned Hs = uaddmax_16x2(It, oe_w, Y);
                        It = uaddmax_16x2(It, e_w, Y);
                        if (FREE && CM < 0)
                            Dt[c] = uaddmax_16x2(Dt[c], aD[c], (unsigned)madd((int)Hs, one, aH[c] - oe_i));
                        else if (FREE && c == CM)
                            Dt[c] = uaddmax_16x2(Dt[c], aDl, (unsigned)madd((int)Hs, one, aHl - oe_i));
                        else
                            Dt[c] = uaddmax_16x2(Dt[c], e_w, Hs);
                        hp = Hc[c];
                        Hc[c] = Hs;
                        continue;
                    }
                    if (CKPT && c == CM) { // max(M, I) of the free-end column (meaningful in its lane only)
                        const unsigned mx = __vmaxu2(MH, It);
                        // ramp steps: a virtual row holds I(0,m) and counts as row 0 (= the initial value); rows past n do not count
                        const unsigned rb = RAMP ? (unsigned)max(r, 0) : (unsigned)r;
                        const unsigned keep = (RAMP && r > n) ? 0u : 0xffffffffu;
                        bestA = max(bestA, __byte_perm(rb, mx, 0x5410) & keep); // mxA << 16 | r: later rows win ties
                        bestB = max(bestB, __byte_perm(rb, mx, 0x7610) & keep);
                    }
                    const unsigned H = umax3_16x2(MH, It, Dt[c]);
                    const unsigned Ho = (unsigned)madd((int)H, one, oe_i);
                    It = uaddmax_16x2(It, e_w, Ho);                      // I' = max(I + E, H + O + E)
                    if (FREE && CM < 0)
                        Dt[c] = uaddmax_16x2(Dt[c], aD[c], (unsigned)madd((int)H, one, aH[c]));
                    else if (FREE && c == CM)
                        Dt[c] = uaddmax_16x2(Dt[c], aDl, (unsigned)madd((int)H, one, aHl));
                    else
                        Dt[c] = uaddmax_16x2(Dt[c], e_w, Ho);            // D' = max(D + E, H + O + E)
                    hp = Hc[c];
                    Hc[c] = H;
                }
                edgeI = It;
                edgeH = Hc[C - 1];
                hpL = inH;
                if (RAMP && r == n)
                    res = Hc[CM < 0 ? 0 : CM];
            }
        };

        auto save = [&](int s) { // state entering step s = kCkK * k, k >= 1
            uint32_t *dst = ck + (size_t)(s / kCkK - 1) * (kCkRegs * 32) + tid;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                dst[c * 32] = Hc[c] - (unsigned)(hs * 65537); // checkpoints hold the unshifted values
                dst[(C + c) * 32] = Dt[c];
            }
            dst[20 * 32] = hpL - (unsigned)(hs * 65537);
            dst[21 * 32] = edgeI;
            dst[22 * 32] = edgeH - (unsigned)(hs * 65537);
            if (SK == 2) {
                dst[23 * 32] = edgeIp;
                dst[24 * 32] = edgeHp - (unsigned)(hs * 65537);
            }
        };
        using Steady = std::integral_constant<int, 0>;
        using Check = std::integral_constant<int, 1>;
        using Ramp = std::integral_constant<int, 2>;
        int t = 0;
        if (SK == 2) {
#pragma unroll 2
            for (; t < SK * (LPP - 1); ++t)
                step(t, Ramp{});
            if (CKPT) {
#pragma unroll 1
                while (t < n - 1) { // steady phase in runs that end at the next checkpoint
                    if ((t & (kCkK - 1)) == 0 && t > 0)
                        save(t);
                    const int tend = min(n - 1, (t | (kCkK - 1)) + 1);
#pragma unroll 4
                    for (; t < tend; ++t)
                        step(t, Steady{});
                }
#pragma unroll 2
                for (; t < T; ++t) {
                    if ((t & (kCkK - 1)) == 0 && t > 0)
                        save(t);
                    step(t, Ramp{});
                }
            } else {
#pragma unroll 4
                for (; t < n - 1; ++t)
                    step(t, Steady{});
#pragma unroll 2
                for (; t < T; ++t)
                    step(t, Ramp{});
            }
        } else {
#pragma unroll 1
        for (; t < SK * (LPP - 1); ++t)
            step(t, Check{});
        }
        if (SK == 2) {
        } else if (CKPT) {
#pragma unroll 1
            while (t < n - 1) { // steady phase in runs that end at the next checkpoint, each run unrolled by two
                if ((t & (kCkK - 1)) == 0 && t > 0)
                    save(t);
                const int tend = min(n - 1, (t | (kCkK - 1)) + 1);
#pragma unroll 2
                for (; t < tend; ++t)
                    step(t, Steady{});
            }
#pragma unroll 1
            for (; t < T; ++t) {
                if ((t & (kCkK - 1)) == 0 && t > 0)
                    save(t);
                step(t, Check{});
            }
        } else {
#pragma unroll 2
            for (; t < n - 1; ++t)
                step(t, Steady{});
#pragma unroll 1
            for (; t < T; ++t)
                step(t, Check{});
        }

        { // H(n, m) of each pair sits in the lane and column that own its last query column
            const int lmA = (m - 1) / C, cmA = (m - 1) % C, lmB = (mB - 1) / C, cmB = (mB - 1) % C;
            unsigned hA = Hc[0], hB = Hc[0];
#pragma unroll
            for (int c = 1; c < C; ++c) {
                if (c == cmA)
                    hA = Hc[c];
                if (c == cmB)
                    hB = Hc[c];
            }
            if (SK == 2)
                hA = hB = res; // (cmA == cmB == CM)
            if (lane == lmA && pA0 < P.pair_end) {
                P.out_score[pA0] = (int64_t)(int)(hA & 0xffffu) - 32768 - hs;
                if (CKPT)
                    P.out_best[pA0] = (int64_t)(bestA & 0xffffu);
            }
            if (lane == lmB && pB0 < P.pair_end) {
                P.out_score[pB0] = (int64_t)(int)(hB >> 16) - 32768 - hs;
                if (CKPT)
                    P.out_best[pB0] = (int64_t)(bestB & 0xffffu);
            }
        }
        __syncwarp();
        quad = next;
    }
}

} // namespace gnx

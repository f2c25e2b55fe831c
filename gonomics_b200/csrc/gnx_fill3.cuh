// gnx_fill3.cuh -- the production affine fill kernel (int32, tagged max).  Same recurrence, tie-break and
// 6-bit trace codes as the first-generation affine_fill_kernel (gnx_kernels.cuh), restructured after
// measuring the SM's pipes (profiles/r01_microbench_pipes.txt, profiles/r01a_fill.md):
//   * every integer max / logic / permute instruction issues on the ALU pipe (64 lanes/clk/SM) while IMAD
//     issues on the FMA pipe (64 lanes/clk/SM) in parallel; the first kernel was ALU-bound (81 % vs 31 %).
//     The cell update is "plain adds + one VIMNMX3 per plane", the adds forced onto the FMA pipe (madd()),
//     leaving per cell on the ALU pipe: 3 tag-clears, 3 maxes, XOR3 + funnel shift (trace code).
//   * one warp per CTA and a persistent, occupancy-sized grid: pair index, lengths and base pointers are
//     CTA-uniform; the step loop is split into ramp-up / steady / ramp-down so the steady phase carries no
//     activity predicate; score-only uses H directly (I' = max(I+E, H+O+E), D' likewise; needs O <= 0).
//   * LPP lanes per pair (32 or 16).  With LPP = 16 a warp carries two pairs side by side (lanes 0-15 and
//     16-31), each lane owning C = 10 columns: a 150-column read fills 15 of 16 lanes, the row skew is
//     15 steps instead of 31, and the per-step overhead (shuffles, base fetch, loop control, stores) is
//     paid once per 10 cells instead of once per 5.
//   * substitution scores come from a per-lane shared-memory table laid out [column][base][thread], read
//     with one LDS per cell at an immediate offset from a per-step base register (bank = thread, so
//     conflict-free).  The lookup therefore costs no ALU- or FMA-pipe slot (the PRMT of fill2 was an
//     ALU op) and handles N (dim 5) in the same kernel -- no ACGT/N class split.
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

#ifndef GNX_FILL3_MINB
#define GNX_FILL3_MINB 12
#endif
#ifndef GNX_F3_GROUP4
#define GNX_F3_GROUP4 1
#endif
#ifndef GNX_F3_STAGE
#define GNX_F3_STAGE 1
#endif
#ifndef GNX_F3_UNROLL_TRACE
#define GNX_F3_UNROLL_TRACE 1
#endif
#ifndef GNX_F3_UNROLL_SCORE
#define GNX_F3_UNROLL_SCORE 4
#endif

template <int C, int LPP>
__host__ __device__ inline int64_t trace_words3(int64_t n_max_of_group, int64_t m)
{
    if (n_max_of_group <= 0 || m <= 0)
        return 0;
    const int64_t strips = (m + LPP * C - 1) / (LPP * C);
    return strips * (n_max_of_group + LPP - 1) * trace_wpl(C) * 32;
}

constexpr int kDimP = 5; // rows of the per-lane score table (bases 0..4); matrices with dim > 5 use fill2
constexpr int kUnrollTrace = GNX_F3_UNROLL_TRACE, kUnrollScore = GNX_F3_UNROLL_SCORE;
constexpr bool kBlock4 = true; // trace rows are blocked four steps per 16-byte piece (layout 3)
constexpr bool kGroup4 = GNX_F3_GROUP4 != 0; // traced steady loop processes aligned groups of 4 steps (no window shift)
constexpr int kRing = 1024; // single-strip kernels stage the whole target (alpha) in shared memory: n <= kRing
// MODE 0: score only (untagged, needs O <= 0)   1: tagged arithmetic, no stores   2: tagged + trace stores
// SK   row skew between neighbouring lanes.  With SK = 2 lane l is two rows behind lane l-1, so the edge
//      values a step consumes were produced two steps earlier: the shuffles of consecutive steps no
//      longer serialise them and the (unrolled-by-2) steady loop overlaps the I-plane dependency chains
//      of two rows (ncu: the SK = 1 kernels sat in fixed-latency `wait` stalls with 1.3 eligible warps).
template <int C, int LPP, int MODE, bool FREE, bool MULTI, int SK>
__global__ void __launch_bounds__(32, (C == 5 ? 20 : GNX_FILL3_MINB)) affine_fill3_kernel(const FillParams P)
{
    constexpr bool TRACE = MODE >= 1, STORE = MODE == 2;
    constexpr int G = 32 / LPP;
    constexpr int SC = TRACE ? kScale : 1;
    constexpr int FI = TRACE ? kFI : 0, FD = TRACE ? kFD : 0, FH = TRACE ? kFH : 0;
    constexpr int WPL = trace_wpl(C);
    constexpr int NEG = kNeg32;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int CLR = ~(kScale - 1);
    static_assert(!(MULTI && LPP != 32), "multi-strip pairs use one pair per warp");

    __shared__ int s_tab[C * kDimP * 32]; // [c][a][thread]
    // !MULTI: the pair's target bases, staged once per pair (one coalesced pass) so that the per-step base
    // fetch is an LDS that never waits on L2/HBM; MULTI (long sequences) streams them with LDG instead.
    constexpr bool STAGE = !MULTI && GNX_F3_STAGE;
    constexpr int kTgtPitch = kRing + 64; // +64 B: the two pairs' rows land in different banks
    __shared__ uint8_t s_tgt[STAGE ? G * kTgtPitch : 4];
    const int tid = threadIdx.x;
    const int lane = tid % LPP, half = tid / LPP;
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_s = (O + E) * SC, e_s = E * SC;
    const int kI = oe_s + 2 * FI - 2 * FH;
    const int iI = e_s + FI, iD = oe_s;
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;
    const int fh_reg = FH * one;

    int2 *edge_a = MULTI ? P.edge + (size_t)blockIdx.x * 2 * P.edge_stride : nullptr;
    int2 *edge_b = MULTI ? edge_a + P.edge_stride : nullptr;
    const int64_t n_groups = (P.pair_end - P.pair_begin + G - 1) / G;

    for (int64_t group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int64_t pair = P.pair_begin + group * G + half;
        int n = 0, m = 0;
        const uint8_t *__restrict__ alpha = P.alpha;
        const uint8_t *__restrict__ beta = P.beta;
        bool mine = pair < P.pair_end && (!P.pair_class || P.pair_class[pair] <= 1);
        if (mine) {
            const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
            n = (int)(P.alpha_off[pair + 1] - a0);
            m = (int)(P.beta_off[pair + 1] - b0);
            alpha += a0;
            beta += b0;
            if (n == 0 || m == 0) { // closed forms of the boundary row / column
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
                mine = false;
                n = 0;
            }
        }
        if (!mine)
            n = 0, m = 0;
        int nmax = n, nmin = n, mmax = m;
        if (G > 1) {
            nmax = max(n, __shfl_xor_sync(FULL, n, 16));
            nmin = min(n, __shfl_xor_sync(FULL, n, 16));
            mmax = max(m, __shfl_xor_sync(FULL, m, 16));
        }
        if (nmax == 0)
            continue;
        // steps per strip for the whole warp; with STORE the trace rows are written four steps at a time
        // (one 16-byte store per lane), so the step count is padded to a multiple of 4
        const int T = (STORE && kBlock4) ? ((nmax + SK * (LPP - 1) + 3) & ~3) : nmax + SK * (LPP - 1);
        const int Tp = kBlock4 ? ((n + SK * (LPP - 1) + 3) & ~3) : n + SK * (LPP - 1); // strip pitch (rows), MULTI only
        const int strips = MULTI ? (mmax + LPP * C - 1) / (LPP * C) : 1;
        uint32_t *tbase = (STORE && mine) ? P.trace + P.trace_off[pair - P.pair_begin] : nullptr;

        for (int p = 0; p < strips; ++p) {
            const int jbase = p * LPP * C + lane * C;
            int aM[C], aI[C], aD[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const int q = (mine && j <= m) ? (int)beta[j - 1] : 0;
#pragma unroll
                for (int a = 0; a < kDimP; ++a) {
                    int v = 0;
                    if (a < P.dim && q < P.dim)
                        v = P.scores[a * P.dim + q] * SC + 2 * FH;
                    s_tab[(c * kDimP + a) * 32 + tid] = v;
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
                Dt[c] = max3(NEG + 2 * FH + aM[c], i0 + FH + aI[c], NEG + aD[c]);
            }
            int hpL = (jbase == 0) ? P.h00 * SC : (O + jbase * E) * SC;
            int edgeI = 0, edgeH = 0;   // what lane+1 consumes SK steps after it was produced
            int edgeI1 = 0, edgeH1 = 0; // SK == 2: the values produced one step ago
            const int2 *ein = (p & 1) ? edge_b : edge_a;
            int2 *eout = (p & 1) ? edge_a : edge_b;
            // blocked trace layout: uint4 index ((t/4)*WPL + k)*32 + thread, word t%4 inside it, so that the
            // four consecutive steps of one lane share a 16-byte piece of one sector (the traceback walks
            // mostly along a lane: 4x fewer sectors touched than with one word per (step, lane) row)
            uint4 *tp4 = STORE ? reinterpret_cast<uint4 *>(tbase + ((size_t)p * Tp * WPL) * 32) + lane : nullptr;
            uint4 wq[WPL];
#pragma unroll
            for (int k = 0; k < WPL; ++k)
                wq[k] = make_uint4(0, 0, 0, 0);
            const bool store_edge = MULTI && (lane == LPP - 1) && (p + 1 < strips);

            int bI = 0, bH = 0;
            // MULTI, p > 0: the left boundary column comes from the previous strip's edge buffer (L2).  A load per
            // step one step ahead sits on the critical path (L2 latency ~ 4 steps at low occupancy: long pairs
            // are occupancy-bound), so the warp prefetches it 32 rows at a time, one block ahead: lane l holds
            // row base + l and lane 0 picks its row up with two shuffles per step.
            int2 eb_cur = make_int2(0, 0), eb_next = make_int2(0, 0);
            int nbI = 0, nbH = 0;
            auto edge_block = [&](int first_row) {
                const int rho = first_row + lane;
                return (MULTI && p > 0 && rho <= n) ? __ldcg(&ein[rho]) : make_int2(0, 0);
            };
            if (MULTI && p > 0) {
                eb_cur = edge_block(1);
                eb_next = edge_block(33);
                nbI = eb_cur.x; // lane 0: row 1
                nbH = eb_cur.y;
            }
            auto boundary = [&](int r) {
                if (!MULTI || p == 0) {
                    const int d0 = FREE ? 0 : (O + r * E) * SC;
                    bI = d0 + iD;
                    bH = d0;
                } else {
                    bI = nbI;
                    bH = nbH;
                }
            };
            if (lane == 0)
                boundary(1);
            const uint8_t *tg = alpha; // where this lane reads its row's base
            if (STAGE) {
                for (int i = lane; i < n; i += LPP)
                    s_tgt[half * kTgtPitch + i] = alpha[i];
                tg = s_tgt + half * kTgtPitch;
            }
            __syncwarp();
            int a_next = (lane == 0 && mine) ? (int)tg[0] : 0;

            auto step = [&](int t, auto check_tag, auto slot_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int r = t - SK * lane + 1;
                if (MULTI && p > 0) { // lane 0 fetches row t + 2 for its next step (boundary(r + 1) below)
                    const int idx = (t + 1) & 31;
                    if (idx == 0) {
                        eb_cur = eb_next;
                        eb_next = edge_block(t + 34);
                    }
                    nbI = __shfl_sync(FULL, eb_cur.x, idx);
                    nbH = __shfl_sync(FULL, eb_cur.y, idx);
                }
                int inI = __shfl_up_sync(FULL, edgeI, 1, LPP);
                int inH = __shfl_up_sync(FULL, edgeH, 1, LPP);
                if (SK == 2) { // age the pipeline: the one-step-old edge becomes visible to lane+1 next step
                    edgeI = edgeI1;
                    edgeH = edgeH1;
                }
                if (lane == 0) {
                    inI = bI;
                    inH = bH;
                }
                const int a = a_next;
                bool active = true;
                if (CHECK) {
                    active = (unsigned)(r - 1) < (unsigned)n;
                    if ((unsigned)r < (unsigned)n)
                        a_next = tg[r];
                } else {
                    a_next = tg[r];
                }
                unsigned w[WPL];
#pragma unroll
                for (int k = 0; k < WPL; ++k)
                    w[k] = 0;
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    const int *row = s_tab + a * 32 + tid; // &s_tab[(0*kDimP + a)*32 + tid]
                    int It = inI, hp = hpL;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int s = row[c * kDimP * 32]; // LDS at an immediate offset, bank = thread
                        const int MH = madd(hp, one, s);
                        if (TRACE) {
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
                            Dt[c] = FREE ? addmax(Dt[c], aD[c], madd(H, one, aI[c])) : addmax(Dt[c], e_s, Ho);
                            hp = Hc[c];
                            Hc[c] = H;
                        }
                    }
                    if (SK == 2) {
                        edgeI1 = It;
                        edgeH1 = Hc[C - 1];
                    } else {
                        edgeI = It;
                        edgeH = Hc[C - 1];
                    }
                    hpL = inH;
                    if (store_edge)
                        eout[r] = make_int2(It, Hc[C - 1]);
                }
                if (STORE) {
                    constexpr int SLOT = decltype(slot_tag)::value; // 0..3: aligned group position, -1: runtime
                    if (SLOT < 0) { // age the 4-step window
#pragma unroll
                        for (int k = 0; k < WPL; ++k) {
                            wq[k].x = wq[k].y;
                            wq[k].y = wq[k].z;
                            wq[k].z = wq[k].w;
                            wq[k].w = w[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < WPL; ++k) {
                            if (SLOT == 0)
                                wq[k].x = w[k];
                            if (SLOT == 1)
                                wq[k].y = w[k];
                            if (SLOT == 2)
                                wq[k].z = w[k];
                            if (SLOT == 3)
                                wq[k].w = w[k];
                        }
                    }
                    if (SLOT == 3 || (SLOT < 0 && (t & 3) == 3)) { // every fourth step: 16 B per lane
                        if (tbase) {
#pragma unroll
                            for (int k = 0; k < WPL; ++k)
                                tp4[(size_t)k * 32] = wq[k];
                        }
                        tp4 += WPL * 32;
                    }
                }
            };

            using RT = std::integral_constant<int, -1>;
            int t = 0;
            if (STORE && kGroup4) {
                // ramp-up to a multiple of 4, then aligned groups of four steady steps (the 4-step trace
                // window is filled in place, one 16-byte store per group), then the checked remainder
#pragma unroll 1
                for (; t < ((SK * (LPP - 1) + 3) & ~3); ++t)
                    step(t, std::true_type{}, RT{});
#pragma unroll 1
                for (; t + 3 < nmin - 1; t += 4) {
                    step(t, std::false_type{}, std::integral_constant<int, 0>{});
                    step(t + 1, std::false_type{}, std::integral_constant<int, 1>{});
                    step(t + 2, std::false_type{}, std::integral_constant<int, 2>{});
                    step(t + 3, std::false_type{}, std::integral_constant<int, 3>{});
                }
            } else {
#pragma unroll 1
                for (; t < SK * (LPP - 1); ++t)
                    step(t, std::true_type{}, RT{});
                if (TRACE) {
#pragma unroll kUnrollTrace
                    for (; t < nmin - 1; ++t) // steady: every lane of every pair in the warp is on a valid row < n
                        step(t, std::false_type{}, RT{});
                } else {
#pragma unroll kUnrollScore
                    for (; t < nmin - 1; ++t)
                        step(t, std::false_type{}, RT{});
                }
            }
#pragma unroll 1
            for (; t < T; ++t)
                step(t, std::true_type{}, RT{});

            if (mine) {
                const int pm = (m - 1) / (LPP * C), lm = ((m - 1) % (LPP * C)) / C, cm = (m - 1) % C;
                if (p == pm && lane == lm) {
                    int h = Hc[0];
#pragma unroll
                    for (int c = 1; c < C; ++c)
                        if (c == cm)
                            h = Hc[c];
                    P.out_score[pair] = (int64_t)(h / SC);
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// affine_fill3w_kernel: one CTA of NW warps per LONG pair (BASELINE config C4: 10 kb x 10 kb).
//
// The one-warp-per-pair MULTI kernel above is occupancy-bound on long pairs: a 10 kb x 10 kb pair owns
// 82 MB of traceback matrix, so only ~1000 pairs (= warps) fit in the workspace and each SM sees 3-7 warps.
// Here the NW warps of a CTA sweep NW consecutive 320-column strips of the SAME pair concurrently: warp w
// runs strip k*NW + w in round k and receives its left boundary column (I, H per row) from warp w-1 through
// a shared-memory ring (8 stages of 16 rows, each stage guarded by a full/empty mbarrier pair -- no fence
// instruction at all: MEMBAR.CTA per batch measured 2x slower, the GPU-scope fence of the L2 ring 2.9x).  Warp 0 of round k+1 reads
// the column warp NW-1 wrote to the per-CTA global ping-pong buffer in round k; rounds are separated by
// __syncthreads() (the pipeline drains for ~(NW-1)*40 of ~10 000 steps).  Cell arithmetic, tags, trace
// layout and score output are identical to affine_fill3_kernel<C, 32, MODE, FREE, true, 1>.
// ------------------------------------------------------------------------------------------------
// mbarrier producer/consumer primitives (shared::cta): arrive has release.cta and try_wait acquire.cta
// semantics, which orders the plain ring stores/loads around them without any MEMBAR.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

constexpr int kWBatch = 16, kWStages = 8, kWRing = kWBatch * kWStages;

template <int C, int MODE, bool FREE, int NW>
__global__ void __launch_bounds__(32 * NW, GNX_FILL3_MINB / NW) affine_fill3w_kernel(const FillParams P)
{
    constexpr bool TRACE = MODE >= 1, STORE = MODE == 2;
    constexpr int LPP = 32;
    constexpr int SC = TRACE ? kScale : 1;
    constexpr int FI = TRACE ? kFI : 0, FD = TRACE ? kFD : 0, FH = TRACE ? kFH : 0;
    constexpr int WPL = trace_wpl(C);
    constexpr int NEG = kNeg32;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int CLR = ~(kScale - 1);

    __shared__ int s_tab_all[NW][C * kDimP * 32]; // per warp: [c][a][lane]
    __shared__ int2 s_ring[NW][kWRing];           // ring w: written by warp w (lane 31), read by warp w+1 (lane 0)
    __shared__ uint64_t s_full[NW][kWStages], s_empty[NW][kWStages]; // one mbarrier pair per 8-row stage of ring w
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_tab = s_tab_all[warp];
    const int one = P.one;
    const int O = P.gap_open, E = P.gap_extend;
    const int oe_s = (O + E) * SC, e_s = E * SC;
    const int kI = oe_s + 2 * FI - 2 * FH;
    const int iI = e_s + FI, iD = oe_s;
    const int dMn = oe_s + 2 * FD - 2 * FH, dIn = oe_s + FD - FH, dDn = e_s;
    const int dMl = 2 * FD - 2 * FH, dIl = FD - FH, dDl = 0;
    const int fh_reg = FH * one;
    int2 *edge_a = P.edge + (size_t)blockIdx.x * 2 * P.edge_stride;
    int2 *edge_b = edge_a + P.edge_stride;
    if (threadIdx.x < NW * kWStages) { // one arriving thread each: lane 31 of the producer / lane 0 of the consumer
        mbar_init(&s_full[0][0] + threadIdx.x, 1);
        mbar_init(&s_empty[0][0] + threadIdx.x, 1);
    }
    __syncthreads();
    // batches handed over so far on the ring this warp writes (g_out) / reads (g_in); both ends count the same
    // batches, so stage = g % 8 and phase parity = (g / 8) & 1 stay in step for the whole kernel
    unsigned g_out = 0, g_in = 0;

    for (int64_t pair = P.pair_begin + blockIdx.x; pair < P.pair_end; pair += gridDim.x) {
        if (P.pair_class && P.pair_class[pair] > 1)
            continue; // CTA-uniform
        const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
        const int n = (int)(P.alpha_off[pair + 1] - a0);
        const int m = (int)(P.beta_off[pair + 1] - b0);
        const uint8_t *__restrict__ alpha = P.alpha + a0;
        const uint8_t *__restrict__ beta = P.beta + b0;
        if (n == 0 || m == 0) { // closed forms of the boundary row / column
            if (threadIdx.x == 0) {
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
        const int T = STORE ? ((n + LPP - 1 + 3) & ~3) : n + LPP - 1;
        const int Tp = (n + LPP - 1 + 3) & ~3;
        const int strips = (m + LPP * C - 1) / (LPP * C);
        const int rounds = (strips + NW - 1) / NW;
        uint32_t *tbase = STORE ? P.trace + P.trace_off[pair - P.pair_begin] : nullptr;

        for (int k = 0; k < rounds; ++k) {
            __syncthreads(); // the previous round is complete: its global edge column is visible, the rings are idle
            const int p = k * NW + warp;
            if (p >= strips)
                continue;
            const int jbase = p * LPP * C + lane * C;
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
                    s_tab[(c * kDimP + a) * 32 + lane] = v;
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
                Dt[c] = max3(NEG + 2 * FH + aM[c], i0 + FH + aI[c], NEG + aD[c]);
            }
            int hpL = (jbase == 0) ? P.h00 * SC : (O + jbase * E) * SC;
            int edgeI = 0, edgeH = 0;
            const int2 *ein = (k & 1) ? edge_b : edge_a;   // written by warp NW-1 in round k-1
            int2 *eout = (k & 1) ? edge_a : edge_b;
            uint4 *tp4 = STORE ? reinterpret_cast<uint4 *>(tbase + ((size_t)p * Tp * WPL) * 32) + lane : nullptr;
            uint4 wq[WPL];
#pragma unroll
            for (int q = 0; q < WPL; ++q)
                wq[q] = make_uint4(0, 0, 0, 0);
            const bool has_next = p + 1 < strips;
            const bool to_ring = has_next && warp < NW - 1, to_global = has_next && warp == NW - 1;
            const int win = warp > 0 ? warp - 1 : 0;
            const int2 *ring_in = s_ring[win];
            const unsigned nb = (unsigned)(n + kWBatch - 1) / kWBatch; // batches per strip
            const unsigned gi0 = g_in, go0 = g_out;
            if (warp > 0)
                g_in += nb; // this strip consumes nb batches from warp-1 ...
            if (to_ring)
                g_out += nb; // ... and hands nb batches to warp+1

            int bI = 0, bH = 0;
            // warp 0, p > 0: block prefetch of the global edge column (see affine_fill3_kernel)
            const bool gin = warp == 0 && p > 0;
            int2 eb_cur = make_int2(0, 0), eb_next = make_int2(0, 0);
            int nbI = 0, nbH = 0;
            auto edge_block = [&](int first_row) {
                const int rho = first_row + lane;
                return (gin && rho <= n) ? __ldcg(&ein[rho]) : make_int2(0, 0);
            };
            if (gin) {
                eb_cur = edge_block(1);
                eb_next = edge_block(33);
                nbI = eb_cur.x;
                nbH = eb_cur.y;
            }
            auto boundary = [&](int r) { // lane 0 only: (I, H) of the column left of this strip, row r
                if (p == 0) {
                    const int d0 = FREE ? 0 : (O + r * E) * SC;
                    bI = d0 + iD;
                    bH = d0;
                } else if (warp == 0) {
                    bI = nbI;
                    bH = nbH;
                } else {
                    const unsigned g = gi0 + (unsigned)(r - 1) / kWBatch;
                    const int2 v = ring_in[(r - 1) & (kWRing - 1)];
                    bI = v.x;
                    bH = v.y;
                    if ((r & (kWBatch - 1)) == 0 || r == n) // last row of the batch: hand the stage back
                        mbar_arrive(&s_empty[win][g % kWStages]);
                }
            };
            // The ring waits are executed by the WHOLE warp (their conditions depend only on the step index):
            // a spin loop inside the lane-0 / lane-31 branches leaves the warp split for the rest of the step
            // (ncu: the cell loop issued twice, 16 active threads on average).
            //   consumer: lane 0 fetches row t + 2 during step t; at the first row of a batch wait until the
            //             producer has filled that batch AND the next one (batches complete in order), so that in
            //             steady state the consumer is never polling a barrier that is about to flip;
            //   producer: lane 31 writes row t - 30 during step t; at the first row of a batch the stage must
            //             have been handed back by the consumer.
            const bool rin = warp > 0;
            auto wait_full = [&](int row) {
                const unsigned g = gi0 + (unsigned)(row - 1) / kWBatch;
                const unsigned gw = min(g + 1, gi0 + nb - 1);
                mbar_wait(&s_full[win][gw % kWStages], (gw / kWStages) & 1);
            };
            if (rin)
                wait_full(1);
            if (lane == 0)
                boundary(1);
            int a_next = (lane == 0) ? (int)alpha[0] : 0;

            auto step = [&](int t, auto check_tag, auto slot_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int r = t - lane + 1;
                if (rin && ((t + 1) & (kWBatch - 1)) == 0 && t + 2 <= n)
                    wait_full(t + 2);
                if (to_ring && ((t - 31) & (kWBatch - 1)) == 0 && t >= 31 && t - 30 <= n) {
                    const unsigned g = go0 + (unsigned)(t - 31) / kWBatch;
                    mbar_wait(&s_empty[warp][g % kWStages], ((g / kWStages) & 1) ^ 1);
                }
                if (gin) {
                    const int idx = (t + 1) & 31;
                    if (idx == 0) {
                        eb_cur = eb_next;
                        eb_next = edge_block(t + 34);
                    }
                    nbI = __shfl_sync(FULL, eb_cur.x, idx);
                    nbH = __shfl_sync(FULL, eb_cur.y, idx);
                }
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
                    a_next = alpha[r];
                }
                unsigned w[WPL];
#pragma unroll
                for (int q = 0; q < WPL; ++q)
                    w[q] = 0;
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    const int *row = s_tab + a * 32 + lane;
                    int It = inI, hp = hpL;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int s = row[c * kDimP * 32];
                        const int MH = madd(hp, one, s);
                        if (TRACE) {
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
                            Dt[c] = FREE ? addmax(Dt[c], aD[c], madd(H, one, aI[c])) : addmax(Dt[c], e_s, Ho);
                            hp = Hc[c];
                            Hc[c] = H;
                        }
                    }
                    edgeI = It;
                    edgeH = Hc[C - 1];
                    hpL = inH;
                    if (lane == LPP - 1) {
                        if (to_global) {
                            eout[r] = make_int2(It, Hc[C - 1]);
                        } else if (to_ring) {
                            const unsigned g = go0 + (unsigned)(r - 1) / kWBatch;
                            s_ring[warp][(r - 1) & (kWRing - 1)] = make_int2(It, Hc[C - 1]);
                            if ((r & (kWBatch - 1)) == 0 || r == n)
                                mbar_arrive(&s_full[warp][g % kWStages]);
                        }
                    }
                }
                if (STORE) {
                    constexpr int SLOT = decltype(slot_tag)::value; // 0..3: aligned group position, -1: runtime
                    if (SLOT < 0) {
#pragma unroll
                        for (int q = 0; q < WPL; ++q) {
                            wq[q].x = wq[q].y;
                            wq[q].y = wq[q].z;
                            wq[q].z = wq[q].w;
                            wq[q].w = w[q];
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < WPL; ++q) {
                            if (SLOT == 0)
                                wq[q].x = w[q];
                            if (SLOT == 1)
                                wq[q].y = w[q];
                            if (SLOT == 2)
                                wq[q].z = w[q];
                            if (SLOT == 3)
                                wq[q].w = w[q];
                        }
                    }
                    if (SLOT == 3 || (SLOT < 0 && (t & 3) == 3)) {
#pragma unroll
                        for (int q = 0; q < WPL; ++q)
                            tp4[(size_t)q * 32] = wq[q];
                        tp4 += WPL * 32;
                    }
                }
            };

            using RT = std::integral_constant<int, -1>;
            int t = 0;
            if (STORE) {
#pragma unroll 1
                for (; t < ((LPP - 1 + 3) & ~3); ++t)
                    step(t, std::true_type{}, RT{});
#pragma unroll 1
                for (; t + 3 < n - 1; t += 4) {
                    step(t, std::false_type{}, std::integral_constant<int, 0>{});
                    step(t + 1, std::false_type{}, std::integral_constant<int, 1>{});
                    step(t + 2, std::false_type{}, std::integral_constant<int, 2>{});
                    step(t + 3, std::false_type{}, std::integral_constant<int, 3>{});
                }
            } else {
#pragma unroll 1
                for (; t < LPP - 1; ++t)
                    step(t, std::true_type{}, RT{});
#pragma unroll 2
                for (; t < n - 1; ++t)
                    step(t, std::false_type{}, RT{});
            }
#pragma unroll 1
            for (; t < T; ++t)
                step(t, std::true_type{}, RT{});

            const int pm = (m - 1) / (LPP * C), lm = ((m - 1) % (LPP * C)) / C, cm = (m - 1) % C;
            if (p == pm && lane == lm) {
                int h = Hc[0];
#pragma unroll
                for (int c = 1; c < C; ++c)
                    if (c == cm)
                        h = Hc[c];
                P.out_score[pair] = (int64_t)(h / SC);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Constant-gap (Needleman-Wunsch) fill on the same wavefront (align/constGap_highMem.go:23-40):
//   m(i,j), tr = T(m(i-1,j-1)+s, m(i,j-1)+g, m(i-1,j)+g),  boundaries m(0,j) = j*g, m(i,0) = i*g.
// One plane, values carried as 4*v + tag (diag 2 = ColM, left 1 = ColI, up 0 = ColD), so a cell is one
// VIMNMX3 + one tag-clear + one funnel shift on the ALU pipe, three IMAD adds on the FMA pipe and one LDS.
// Trace: 2 bits per cell, the lane's 10 codes per step in one word, rows blocked four steps per 16 bytes
// (layout [strip][step/4][thread][step%4]); code k of a word sits at bit 32 - 2*(10 - k).
// ------------------------------------------------------------------------------------------------
//
// EXT selects the gsw extend step's variants of the same one-plane DP (genomeGraph/search.go:234-321):
//   1  LeftDynamicAln  (:234-274): zero boundaries; a cell is clipped at 0 AFTER its trace is recorded; the score
//      is m(n,m).  The traceback (traceback_ext_kernel) walks from (n,m) while the cell value is > 0.
//   2  RightDynamicAln (:276-321): Needleman-Wunsch boundaries; the result is the first cell, in row-major
//      order, holding the strict maximum (currMax starts at 0, so (0,0) wins when nothing is positive).  Each
//      lane keeps key = 16*value + (15 - column) of its best cell and the row it was found in; rows are visited
//      in increasing order and a later row replaces the best only on a strictly larger VALUE, so per lane the
//      row-major-first maximum survives; lanes and strips are merged on (value, -row, -column).  Needs g <= 0.
template <int C, int LPP, bool STORE, bool MULTI, int EXT = 0>
__global__ void __launch_bounds__(32, 16) const_fill3_kernel(const FillParams P)
{
    constexpr int G = 32 / LPP;
    constexpr int SC = STORE ? 4 : 1;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(!(MULTI && LPP != 32), "multi-strip pairs use one pair per warp");
    static_assert(C <= 16, "one trace word per lane per step");
    __shared__ int s_tab[C * kDimP * 32];
    constexpr bool STAGE = !MULTI;
    constexpr int kTgtPitch = kRing + 64;
    __shared__ uint8_t s_tgt[STAGE ? G * kTgtPitch : 4];
    const int tid = threadIdx.x;
    const int lane = tid % LPP, half = tid / LPP;
    const int one = P.one;
    const int g = P.gap_open; // the single gap penalty
    const int g_left = g * SC + (STORE ? 1 : 0), g_up = g * SC;
    const int k16 = one * (16 / SC); // EXT 2: clean value -> key multiplier (opaque, so the key is one IMAD)
    int2 *edge_a = MULTI ? P.edge + (size_t)blockIdx.x * 2 * P.edge_stride : nullptr;
    int2 *edge_b = MULTI ? edge_a + P.edge_stride : nullptr;
    const int64_t n_groups = (P.pair_end - P.pair_begin + G - 1) / G;

    for (int64_t group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const int64_t pair = P.pair_begin + group * G + half;
        int n = 0, m = 0;
        const uint8_t *__restrict__ alpha = P.alpha;
        const uint8_t *__restrict__ beta = P.beta;
        bool mine = pair < P.pair_end && (!P.pair_class || P.pair_class[pair] <= 1);
        if (mine) {
            const int64_t a0 = P.alpha_off[pair], b0 = P.beta_off[pair];
            n = (int)(P.alpha_off[pair + 1] - a0);
            m = (int)(P.beta_off[pair + 1] - b0);
            alpha += a0;
            beta += b0;
            if (n == 0 || m == 0) { // boundary row / column (constGap_highMem.go:27-32)
                if (lane == 0) {
                    // EXT: the boundaries are 0 (left) or never positive (right, g <= 0): score 0 at (0,0)
                    P.out_score[pair] = EXT ? 0 : (int64_t)g * (n + m);
                    if (EXT == 2)
                        P.out_best[pair] = 0;
                }
                mine = false;
            }
        }
        if (!mine)
            n = 0, m = 0;
        int nmax = n, nmin = n, mmax = m;
        if (G > 1) {
            nmax = max(n, __shfl_xor_sync(FULL, n, 16));
            nmin = min(n, __shfl_xor_sync(FULL, n, 16));
            mmax = max(m, __shfl_xor_sync(FULL, m, 16));
        }
        if (nmax == 0)
            continue;
        const int T = STORE ? ((nmax + LPP - 1 + 3) & ~3) : nmax + LPP - 1;
        const int Tp = (n + LPP - 1 + 3) & ~3;
        const int strips = MULTI ? (mmax + LPP * C - 1) / (LPP * C) : 1;
        uint32_t *tbase = (STORE && mine) ? P.trace + P.trace_off[pair - P.pair_begin] : nullptr;
        long long bestK = (long long)0xFFFFF << 20 | 0xFFFFF; // EXT 2: value 0 at (0,0)

        for (int p = 0; p < strips; ++p) {
            const int jbase = p * LPP * C + lane * C;
            int Hc[C];
            int kadd[EXT == 2 ? C : 1];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int j = jbase + c + 1;
                const int q = (mine && j <= m) ? (int)beta[j - 1] : 0;
#pragma unroll
                for (int a = 0; a < kDimP; ++a) {
                    int v = 0;
                    if (a < P.dim && q < P.dim)
                        v = P.scores[a * P.dim + q] * SC + (STORE ? 2 : 0);
                    s_tab[(c * kDimP + a) * 32 + tid] = v;
                }
                Hc[c] = EXT == 1 ? 0 : j * g * SC; // row 0
                if (EXT == 2)
                    kadd[c] = (mine && j <= m) ? 15 - c : kNeg32; // padding columns never hold the maximum
            }
            int bestkey = 15, brow = 0; // EXT 2: this lane's best cell in this strip (value 0, row 0 = "none yet")
            int hpL = EXT == 1 ? 0 : jbase * g * SC;
            int edgeH = 0;
            const int2 *ein = (p & 1) ? edge_b : edge_a;
            int2 *eout = (p & 1) ? edge_a : edge_b;
            uint4 *tp4 = STORE ? reinterpret_cast<uint4 *>(tbase + ((size_t)p * Tp) * 32) + lane : nullptr;
            uint4 wq = make_uint4(0, 0, 0, 0);
            const bool store_edge = MULTI && (lane == LPP - 1) && (p + 1 < strips);
            int bH = 0;
            auto boundary = [&](int r) { bH = (!MULTI || p == 0) ? (EXT == 1 ? 0 : r * g * SC) : ein[r].x; };
            if (lane == 0)
                boundary(1);
            const uint8_t *tg = alpha;
            if (STAGE) {
                for (int i = lane; i < n; i += LPP)
                    s_tgt[half * kTgtPitch + i] = alpha[i];
                tg = s_tgt + half * kTgtPitch;
            }
            __syncwarp();
            int a_next = (lane == 0 && mine) ? (int)tg[0] : 0;

            auto step = [&](int t, auto check_tag) {
                constexpr bool CHECK = decltype(check_tag)::value;
                const int r = t - lane + 1;
                int inH = __shfl_up_sync(FULL, edgeH, 1, LPP);
                if (lane == 0)
                    inH = bH;
                const int a = a_next;
                bool active = true;
                if (CHECK) {
                    active = (unsigned)(r - 1) < (unsigned)n;
                    if ((unsigned)r < (unsigned)n)
                        a_next = tg[r];
                } else {
                    a_next = tg[r];
                }
                unsigned w = 0;
                if (active) {
                    if (lane == 0 && r < n)
                        boundary(r + 1);
                    const int *row = s_tab + a * 32 + tid;
                    int left = inH, hp = hpL;
                    int rowkey = kNeg32;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int s = row[c * kDimP * 32];
                        const int ht = max3(madd(hp, one, s), madd(left, one, g_left), madd(Hc[c], one, g_up));
                        int cH = ht;
                        if (STORE) {
                            w = shf_r_wrap(w, (unsigned)ht, 2);
                            cH = ht & ~3;
                        }
                        if (EXT == 1)
                            cH = max(cH, 0); // search.go:247-249: clipped after the trace was recorded
                        if (EXT == 2)
                            rowkey = max(rowkey, madd(cH, k16, kadd[c]));
                        hp = Hc[c];
                        Hc[c] = cH;
                        left = cH;
                    }
                    if (EXT == 2 && (rowkey >> 4) > (bestkey >> 4)) {
                        bestkey = rowkey;
                        brow = r;
                    }
                    edgeH = left;
                    hpL = inH;
                    if (store_edge)
                        eout[r] = make_int2(left, 0);
                }
                if (STORE) {
                    wq.x = wq.y;
                    wq.y = wq.z;
                    wq.z = wq.w;
                    wq.w = w;
                    if ((t & 3) == 3) {
                        if (tbase)
                            *tp4 = wq;
                        tp4 += 32;
                    }
                }
            };
            int t = 0;
#pragma unroll 1
            for (; t < LPP - 1; ++t)
                step(t, std::true_type{});
#pragma unroll 4
            for (; t < nmin - 1; ++t)
                step(t, std::false_type{});
#pragma unroll 1
            for (; t < T; ++t)
                step(t, std::true_type{});
            if (EXT == 2) {
                if (brow > 0) { // merge this strip's best on (value, smaller row, smaller column)
                    const int j = jbase + (15 - (bestkey & 15)) + 1;
                    const long long K = ((long long)(bestkey >> 4) << 40) | ((long long)(0xFFFFF - brow) << 20) |
                                        (long long)(0xFFFFF - j);
                    bestK = max(bestK, K);
                }
            } else if (mine) {
                const int pm = (m - 1) / (LPP * C), lm = ((m - 1) % (LPP * C)) / C, cm = (m - 1) % C;
                if (p == pm && lane == lm) {
                    int h = Hc[0];
#pragma unroll
                    for (int c = 1; c < C; ++c)
                        if (c == cm)
                            h = Hc[c];
                    P.out_score[pair] = (int64_t)(h / SC);
                }
            }
            __syncwarp();
        }
        if (EXT == 2) {
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1)
                bestK = max(bestK, __shfl_xor_sync(FULL, bestK, o));
            if (mine && lane == 0) {
                P.out_score[pair] = bestK >> 40;
                const long long bi = 0xFFFFF - ((bestK >> 20) & 0xFFFFF), bj = 0xFFFFF - (bestK & 0xFFFFF);
                P.out_best[pair] = (bi << 32) | bj;
            }
        }
    }
}

} // namespace gnx

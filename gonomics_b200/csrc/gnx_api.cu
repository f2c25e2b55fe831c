// gnx_api.cu -- libgnxalign.so: context, chunk pipeline and the exported C ABI (include/gnxalign.h).
//
// Host pipeline for the host-buffer entry points (DESIGN.md "Pipeline"): the batch is cut into
// chunks whose traceback matrices fit the per-slot share of the workspace; kSlots chunks are in
// flight on kSlots streams, each going  H2D -> classify -> fill -> traceback -> scan -> [total to
// host] -> expand -> D2H, so copies of one chunk overlap the DP fill of another.
#include "../../include/gnxalign.h"
#include "gnx_kernels.cuh"
#include "gnx_fill3.cuh"
#include "gnx_fill16.cuh"
#include "gnx_ckpt.cuh"
#include "gnx_long.cuh"
#include "gnx_profile.cuh"
#include "gnx_twobit.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

extern "C" bool gnx_pack_range_host(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen); // gnx_pack_host.cpp

namespace {

using namespace gnx;

constexpr int kSlots = 3;
constexpr int kSlotCap = 24; // cigar elements kept per pair before the overflow pass
constexpr int kSlotCapCkpt = 64;
constexpr int kSlotCapLong = 1024; // ... for multi-strip (long) pairs, whose cigars run to hundreds of elements and
                                   // whose second traceback pass is a second walk of a 20 000-step route

static_assert(sizeof(gnx_cigar) == 16, "gnx_cigar must match Go's align.Cigar layout");
static_assert(sizeof(CigarOut) == 16, "device cigar record must match gnx_cigar");

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + std::min<size_t>(bytes / 8, (size_t)256 << 20) + 256; // growth slack, bounded for the
                                                                                     // 100-GB trace buffers of long pairs
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return (T *)p; }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return (T *)p; }
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_total = nullptr, ev_done = nullptr;
    DevBuf alpha, beta, aoff, boff, cls, trace, trace_off, slots, counts, score, cig_off, cigars, edge, misc, partials;
    DevBuf best, endi, endj; // gsw extend step: first-maximum cell (right) and the traceback's end coordinates
    DevBuf work;             // checkpoint path: work list of the pairs that need the recompute walk (+ its counter)
    DevBuf qctr;             // packed 16-bit kernels: the quad cursor of their dynamic schedule (self-resetting, zeroed once)
    DevBuf rag;                 // ragged batches: quad_pairs | quad_ck_off | pair_slot (RagTables)
    PinBuf h_rag, h_rag2;       // page-locked staging of the tables (the device-resident path alternates between the two)
    cudaEvent_t ev_rag[2] = {nullptr, nullptr};
    bool ev_rag_set[2] = {false, false};
    DevBuf tb_a, tb_b, tb_meta; // 2-bit inputs: the chunk's packed words (+ word offsets / lengths of ragged chunks)
    PinBuf h_stage_a, h_stage_b, h_total, h_trace_off, h_score, h_off, h_cig, h_endi, h_endj, h_tbmeta;
    // chunk in flight
    int64_t begin = 0, end = 0;
    bool busy = false;
    bool ev_done_set = false; // ev_done has been recorded: h_trace_off may still be the source of a pending upload
};

struct FillEvent {
    cudaEvent_t a, b;
};

} // namespace

struct gnx_ctx {
    int device = 0;
    std::string err;
    size_t workspace = 0;
    Slot slot[kSlots];
    DevBuf status;       // int32 device status word
    DevBuf dr_misc;      // device-resident path scratch (running total, counters)
    // long-pair checkpoint path (gnx_long.cuh): per-warp scratch shared by the slots (its launches are chained on
    // ev_long, so at most one of them runs at a time) + a ring of work counters in its first 256 bytes
    // gnx_batch_device calls share slot[0]'s scratch and the status / running-total words: each call is ordered
    // after the previous one on this context (ev_dev, recorded at the end of every call), whatever stream it uses
    cudaEvent_t ev_dev = nullptr;
    bool ev_dev_set = false;
    DevBuf long_scratch;
    cudaEvent_t ev_long = nullptr;
    bool ev_long_set = false;
    int long_rr = 0;
    // profile (group-vs-group) batches: groups, profiles, pair lists, dense cell-score matrices
    DevBuf pf_cat, pf_goff, pf_nseq, pf_coloff, pf_scores, pf_px, pf_py, pf_aoff, pf_boff, pf_soff, pf_prof, pf_vb,
        pf_smat, pf_err;
    int64_t launches = 0;
    int gate_rr = 0;     // round-robin slot of the per-chunk max-base gate words (status[1..8])
    // options
    int opt_cols = 0;          // 0 = auto
    int64_t opt_chunk_pairs = 1 << 18;
    int opt_blocks_per_sm = 8;
    int opt_fill_impl = 3;     // 1: first-generation affine_fill_kernel (4 warps/CTA), 3: affine_fill3/fill16 kernels
    int opt_lpp = 0;           // fill3 lanes per pair: 0 auto, 16 or 32
    int opt_tb_impl = 2;       // 1: generic traceback_kernel, 2: traceback_affine_kernel for fill2/3 traces
    int opt_fill16 = 1;        // allow the packed 16-bit score-only kernel when its range proof holds
    int opt_ctas_per_sm = 32;  // fill2/3 persistent grid = SMs * min(this, occupancy)
    int opt_force_lookup = -1; // -1 auto; 0/1 force the PRMT / shared-memory score lookup for ACGT pairs
    int opt_ckpt = 1;          // allow the checkpoint-and-recompute traceback for uniform freeEndGaps batches
    int opt_wide_cta = -1;     // -1 auto; 0/1 never / always run multi-strip pairs on the 4-warp CTA-per-pair kernel
    int opt_rag = 1;           // ragged batches may use the packed 16-bit kernels (quads binned on the host)
    int opt_tb_tma = 1;        // 2-bit inputs: 1 = fill16 kernels read the packed words (TMA), 0 = always unpack to bytes first
    int opt_pack_stage = 1;    // pageable byte inputs of large uniform batches are packed to 2 bits per base while staged
    int pack_backoff = 0;      // calls left that skip the attempt (the last one met a base >= 4)
    int opt_long = -1;         // -1 auto; 0/1 never / always run multi-strip traceback batches on the tile-checkpoint kernel
    int opt_long_form = 0;     // cell formulation of its score-only pass (gnx_long.cuh FORM)
    int64_t opt_long_pool = 0; // its run-pool entries per chunk (0 = auto; tests shrink it to force the re-run pass)
    int sm_count = 148;
    // stats of the last batch call
    std::vector<FillEvent> fill_events;
    size_t fill_events_used = 0;
    double last_fill_ms = 0;
    int64_t last_fill_launches = 0, last_cells = 0;
    int last_impl = 0, last_flags = 0; // kernel family the last batch call planned (gnx_last_kernel_path)
    // 2-bit entry points: host-side offset arrays derived from the lengths (byte offsets a/b, word offsets a/b), kept
    // between calls with the same uniform shape (a streaming caller's batches) so that they are built once
    std::vector<int64_t> tb_off[4];
    int64_t tb_key[3] = {-1, -1, -1}; // n_pairs, n, m of the cached uniform arrays
    PinBuf gsw_pin[32]; // gnx_gsw_batch: page-locked scratch kept between calls (gnx_gsw.inl)
    // cigars retained after GNX_ECAP
    std::vector<gnx_cigar> retained;
    bool have_retained = false;
};

namespace {

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[512];                                                                            \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                                        \
            ctx->err = buf_;                                                                           \
            return GNX_ECUDA;                                                                          \
        }                                                                                              \
    } while (0)

int fail(gnx_ctx *ctx, int code, const char *msg)
{
    ctx->err = msg;
    return code;
}

struct FillCfg {
    int impl = 1; // 1 affine_fill_kernel / const_fill_kernel, 3 affine_fill3_kernel,
                  // 16 affine_fill16_kernel (packed 16-bit, score only, uniform batch)
    int C = 5;    // columns per lane
    int lpp = 32; // lanes per pair
    int skew = 1; // rows between neighbouring lanes (fill3: 2)
    bool multi = false; // some pair needs more than one strip
    int strips_max = 1; // fill3: strips of the widest pair
    int64_t m_uniform = 0; // fill16: the batch's (uniform) query length
    int64_t n_uniform = 0; // checkpoint path: the batch's (uniform) target length
    int64_t long_pool = 0; // impl 18: run-pool entries per chunk (0 = long_pool_entries())
    bool rag = false;      // impl 16 / 17 on a ragged batch: quads come from RagTables
    bool tb = false;       // impl 16 / 17 on 2-bit inputs: affine_fill16_kernel<TB> stages the packed words by TMA
};

struct Problem {
    FillCfg cfg;
    int kind; // 0 affine global, 1 affine free-end, 2 const gap
    int want_cigar;
    int dim;
    int64_t scores[64];
    int64_t gap_open, gap_extend;
    int h00, h00_plane;
    bool prmt_ok;  // 16-bit PRMT tables usable for ACGT-only pairs
    bool tagged;   // run the tagged (scaled) arithmetic: traceback wanted, or score only with O > 0
    int64_t chunk = 1; // AffineGapChunk: bases per DP cell
    bool wide = false; // int64 plane values (the int32 range proof failed)
    int ext = 0;       // gsw extend step on the const-gap machinery (kind 2): 1 LeftDynamicAln, 2 RightDynamicAln
    bool ext_local = false; // ... in their older LeftLocal / RightLocal form ('=' / 'X' ops, route in alignment order)
    bool twobit = false;                // the caller's sequences are dnaTwoBit words (gnx_*_twobit entry points)
    bool profile = false;               // match scores come from a dense per-pair matrix (gnx_profile.cuh)
    const int64_t *extra_words = nullptr; // profile: per-pair workspace words besides the trace (the S matrix)
};

inline int slot_cap_of(const Problem &pb);
inline size_t slots_bytes(const Problem &pb, int64_t np);
inline int slot_cap_of(const Problem &pb)
{
    if (pb.cfg.impl == 17)
        return kSlotCapCkpt; // the checkpoint path's second pass is a second recompute, not a second walk of a stored trace
    if (pb.cfg.impl == 18)
        return kSlotCapLong;
    return (pb.cfg.multi && !pb.profile && !pb.ext) ? kSlotCapLong : kSlotCap;
}

inline int64_t pool_entries_of(const Problem &pb, int64_t np)
{
    return pb.cfg.long_pool > 0 ? pb.cfg.long_pool : long_pool_entries(np, pb.cfg.n_uniform, pb.cfg.m_uniform);
}
inline size_t slots_bytes(const Problem &pb, int64_t np)
{
    if (pb.cfg.impl == 18) // cursor + per-pair pool offsets + run pool (gnx_long.cuh)
        return long_slot_bytes(np, pool_entries_of(pb, np));
    return (size_t)np * slot_cap_of(pb) * 4;
}

// Exact-arithmetic range analysis for the scaled int32 kernels (DESIGN.md "Arithmetic width").
// Every finite plane value v obeys |v| <= bound.  With values carried as scale*v and -inf = -2^30,
// 2*bound*scale < 2^30 guarantees that nothing wraps and that every "-inf + addend" stays strictly
// below every finite candidate, so the int32 kernels make exactly the comparisons the int64
// reference makes.
int analyse(gnx_ctx *ctx, Problem &pb, int64_t max_n, int64_t max_m)
{
    int64_t smin = 0, smax = 0;
    for (int i = 0; i < pb.dim * pb.dim; ++i) {
        smin = std::min(smin, pb.scores[i]);
        smax = std::max(smax, pb.scores[i]);
    }
    smin *= pb.chunk; // AffineGapChunk: a cell's match score is a sum over `chunk` bases
    smax *= pb.chunk; // (gap_extend already carries the factor)
    const int64_t O = pb.gap_open, E = pb.gap_extend;
    const int64_t absO = O < 0 ? -O : O, absE = E < 0 ? -E : E;
    const int64_t sabs = std::max(-smin, smax);
    const int64_t len = max_n + max_m + 2;
    int64_t core;
    if (O <= 0 && E <= 0) {
        // H(i,j) >= 2O + (i+j)E (all-insert-then-all-delete path); H <= min(n,m) * max score
        core = std::max(2 * absO + len * absE, std::min(max_n, max_m) * std::max<int64_t>(smax, 0));
    } else {
        core = 2 * absO + len * std::max<int64_t>({absE, sabs, absO, 1});
    }
    const int64_t bound = core + 2 * (absO + absE) + sabs + 64;
    pb.tagged = pb.want_cigar || (pb.kind != 2 && O > 0);
    const int64_t scale = pb.ext == 2 ? 16 : (pb.tagged ? (pb.kind == 2 ? 4 : kScale) : 1); // ext 2: 16*v + column keys
    if (max_n >= (1 << 24) || max_m >= (1 << 24))
        return fail(ctx, GNX_ERANGE, "sequence longer than 2^24 bases");
    if (2 * bound * scale >= (int64_t(1) << 30)) {
        // the int32 proof fails: the affine DP falls back to the int64 instantiation (exact for any input
        // whose finite values stay below 2^55, i.e. always in practice); other kernels have no wide form yet
        if (pb.kind == 2 || pb.chunk > 1 || pb.profile || bound >= (int64_t(1) << 54))
            return fail(ctx, GNX_ERANGE, "scores/penalties x lengths exceed the exact range of the DP kernels");
        pb.wide = true;
        pb.tagged = true; // the wide kernel is instantiated in its tagged form only
        pb.cfg.impl = 1;
        pb.cfg.C = 5;
        pb.cfg.lpp = 32;
        pb.cfg.skew = 1;
        pb.cfg.multi = max_m > 32 * 5;
    }
    pb.prmt_ok = sabs * scale + 4 <= 32767;
    // H(0,0) = tripleMaxTrace(0, O, D(0,0))  (affineGap_highMem.go:185-192 + affineTrace :62)
    const int64_t d00 = pb.kind == 1 ? 0 : O;
    if (0 >= O && 0 >= d00) {
        pb.h00 = 0;
        pb.h00_plane = 0;
    } else if (O >= d00) {
        pb.h00 = (int)O;
        pb.h00_plane = 1;
    } else {
        pb.h00 = (int)d00;
        pb.h00_plane = 2;
    }
    return GNX_OK;
}

void pick_cfg(const gnx_ctx *ctx, Problem &pb, int64_t max_m, int64_t max_n)
{
    FillCfg &c = pb.cfg;
    // fill3 (int32, per-lane smem score tables for bases 0..4) is the production affine kernel; the
    // first-generation kernels serve the constant-gap DP, AffineGapChunk and matrices with dim > 5.
    c.impl = (pb.profile || pb.chunk > 1 || pb.dim > kDimP || (ctx->opt_fill_impl == 1 && !pb.ext)) ? 1 : 3;
    c.lpp = 32;
    c.skew = 1;
    if (c.impl == 3) {
        c.C = 10;
        if (ctx->opt_lpp != 32 && max_m <= 16 * c.C)
            c.lpp = 16; // two pairs per warp
        c.multi = max_m > (int64_t)c.lpp * c.C || max_n > kRing; // single-strip kernels stage the target in smem
        if (c.multi)
            c.lpp = 32;
        c.strips_max = (int)((max_m + 32 * c.C - 1) / (32 * c.C));
    } else {
        c.C = (ctx->opt_cols == 5 || ctx->opt_cols == 10) ? ctx->opt_cols : (max_m <= 160 ? 5 : 10);
        c.multi = max_m > 32 * (int64_t)c.C;
    }
}

// Range proof for the packed 16-bit score-only kernel (gnx_fill16.cuh): every state it ever holds,
// including the padding columns up to 160, lies in [LB, UB]; with the +32768 bias both must fit 16 bits.
bool fill16_ok(const gnx_ctx *ctx, const Problem &pb, int64_t min_n, int64_t max_n, int64_t min_m, int64_t max_m,
               bool for_ckpt = false)
{
    if (!ctx->opt_fill16 || ctx->opt_fill_impl != 3 || pb.kind == 2 || pb.dim > kDimP)
        return false;
    if (!for_ckpt && pb.want_cigar)
        return false;
    // checkpoint-and-recompute traceback (gnx_ckpt.cuh): freeEndGaps only, target staged in shared memory, and
    // worth it only when the target is well longer than the query (the path then skips most of the rows)
    if (for_ckpt && (!pb.want_cigar || !ctx->opt_ckpt || pb.kind != 1 || pb.chunk > 1 || pb.profile || pb.ext || max_n > kRing ||
                     min_n < 2 * max_m))
        return false;
    if (min_n < 1 || min_m < 1 || max_m > 160) // every pair non-empty, queries within the 16 x 10 columns of a half-warp
        return false;
    // uniform batches address their quads arithmetically; ragged ones go through quads the host bins by
    // (last-column index, target length) -- see RagTables
    if ((min_n != max_n || min_m != max_m) && (!ctx->opt_rag || pb.chunk > 1 || pb.profile || pb.ext))
        return false;
    const int64_t O = pb.gap_open, E = pb.gap_extend;
    if (O > 0 || E > 0)
        return false;
    int64_t smin = 0, smax = 0;
    for (int i = 0; i < pb.dim * pb.dim; ++i) {
        smin = std::min(smin, pb.scores[i]);
        smax = std::max(smax, pb.scores[i]);
    }
    // free end gaps: H(i,j) >= O + jE (enter from D(i,0)=0);  global: H(i,j) >= 2O + (i+j)E
    const int64_t hlow = pb.kind == 1 ? O + 161 * E : 2 * O + (max_n + 161) * E;
    // two gap opens below the lowest H: the kernel forms I + O + E and D + O + E (GNX_F16_SHORT_CHAIN)
    const int64_t lb = hlow + 2 * (O + E) + smin - 64;
    const int64_t ub = smax * std::min(max_n, max_m) + smax + 64;
    return lb >= -32768 && ub <= 32767 && -smin <= 32000 && smax <= 32000;
}

// Trace words of one group of pairs that share 128-byte trace rows (1 pair, or 2 with lpp == 16);
// n_eff is the largest row count among the group's non-empty pairs.
inline int64_t group_trace_words(const Problem &pb, int64_t n_eff, int64_t m)
{
    if (n_eff <= 0 || m <= 0)
        return 0;
    const FillCfg &c = pb.cfg;
    if (c.impl == 18) // tile checkpoints live in per-warp scratch, not per pair
        return 0;
    if (c.impl == 17) // checkpoints: kCkRegs words per lane every kCkK steps; a group is half a quad
        return ck_count(n_eff) * kCkRegs * 16;
    if (pb.kind == 2 && c.impl != 3)
        return const_trace_words(n_eff, m, c.C);
    const int64_t strips = (m + (int64_t)c.lpp * c.C - 1) / ((int64_t)c.lpp * c.C);
    int64_t rows = n_eff + c.skew * (c.lpp - 1);
    if (c.impl == 3)
        rows = (rows + 3) & ~int64_t(3); // fill3 writes four steps per 16-byte piece
    return strips * rows * (pb.kind == 2 ? 1 : trace_wpl(c.C)) * 32; // const_fill3: one word per lane per step
}

unsigned host_threads();

// Per-pair trace offsets (32-bit words) of chunk [begin, begin+np); returns the chunk's total words.  Large chunks are
// cut into (even-aligned) pieces: every thread fills its piece with offsets relative to the piece, the pieces' totals
// are scanned, and a second sweep adds the bases -- the serial loop cost 2.6 ms per 262,144-pair chunk of a ragged batch.
int64_t compute_trace_offsets(const Problem &pb, const int64_t *aoff, const int64_t *boff, int64_t begin, int64_t np,
                              int64_t *to)
{
    auto len = [&](int64_t k, int64_t &n, int64_t &m) {
        n = aoff[begin + k + 1] - aoff[begin + k];
        m = boff[begin + k + 1] - boff[begin + k];
        if (n == 0 || m == 0)
            n = 0;
    };
    auto piece = [&](int64_t lo, int64_t hi) -> int64_t { // pairs [lo, hi), lo even; returns the piece's words
        int64_t acc = 0;
        if (pb.cfg.lpp == 16) {
            for (int64_t k = lo; k < hi; k += 2) {
                int64_t n0, m0, n1 = 0, m1 = 0;
                len(k, n0, m0);
                if (k + 1 < np)
                    len(k + 1, n1, m1);
                to[k] = acc;
                if (k + 1 < np)
                    to[k + 1] = acc + 16 * 4; // second half-warp: 16 threads x 4 words per 16-byte piece
                acc += group_trace_words(pb, std::max(n0, n1), std::max(m0, m1));
            }
        } else {
            for (int64_t k = lo; k < hi; ++k) {
                int64_t n, m;
                len(k, n, m);
                to[k] = acc;
                acc += group_trace_words(pb, n, m);
            }
        }
        return acc;
    };
    const int nt = np >= (1 << 16) ? (int)std::min(8u, host_threads()) : 1;
    if (nt <= 1) {
        const int64_t acc = piece(0, np);
        to[np] = acc;
        return acc;
    }
    std::vector<int64_t> cut((size_t)nt + 1), tot((size_t)nt + 1, 0);
    for (int t = 0; t <= nt; ++t)
        cut[(size_t)t] = t == nt ? np : (np * t / nt) & ~int64_t(1);
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t)
            th.emplace_back([&, t] { tot[(size_t)t + 1] = piece(cut[(size_t)t], cut[(size_t)t + 1]); });
        tot[1] = piece(cut[0], cut[1]);
        for (auto &x : th)
            x.join();
    }
    for (int t = 0; t < nt; ++t)
        tot[(size_t)t + 1] += tot[(size_t)t];
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t)
            th.emplace_back([&, t] {
                const int64_t base = tot[(size_t)t];
                for (int64_t k = cut[(size_t)t]; k < cut[(size_t)t + 1]; ++k)
                    to[k] += base;
            });
        for (auto &x : th)
            x.join();
    }
    to[np] = tot[(size_t)nt];
    return tot[(size_t)nt];
}

FillEvent &next_fill_event(gnx_ctx *ctx)
{
    if (ctx->fill_events_used == ctx->fill_events.size()) {
        FillEvent fe;
        cudaEventCreate(&fe.a);
        cudaEventCreate(&fe.b);
        ctx->fill_events.push_back(fe);
    }
    return ctx->fill_events[ctx->fill_events_used++];
}

template <int C, bool TRACE, bool FREE, int LOOKUP>
void launch_affine(const FillParams &fp, int grid, cudaStream_t st)
{
    affine_fill_kernel<C, TRACE, FREE, LOOKUP><<<grid, 128, 0, st>>>(fp);
}
template <int C, bool TRACE, int LOOKUP> void launch_const(const FillParams &fp, int grid, cudaStream_t st)
{
    const_fill_kernel<C, TRACE, LOOKUP><<<grid, 128, 0, st>>>(fp);
}

template <int C, int LOOKUP>
void dispatch_fill_cl(const Problem &pb, const FillParams &fp, int grid, cudaStream_t st)
{
    if (pb.kind == 2) {
        if (pb.want_cigar)
            launch_const<C, true, LOOKUP>(fp, grid, st);
        else
            launch_const<C, false, LOOKUP>(fp, grid, st);
    } else if (pb.kind == 1) {
        if (pb.tagged)
            launch_affine<C, true, true, LOOKUP>(fp, grid, st);
        else
            launch_affine<C, false, true, LOOKUP>(fp, grid, st);
    } else {
        if (pb.tagged)
            launch_affine<C, true, false, LOOKUP>(fp, grid, st);
        else
            launch_affine<C, false, false, LOOKUP>(fp, grid, st);
    }
}

// Persistent grid: every CTA must be resident at once (pairs are statically strided over CTAs), so the grid
// is SMs x min(requested CTAs/SM, what the kernel's registers/shared memory allow).
template <int C, int LPP, int MODE, bool FREE, bool MULTI, int SK>
void launch_fill3_sk(const FillParams &fp, int64_t groups, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    static std::atomic<int> occ{0}; // per instantiation; contexts on different threads may race to fill it
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_fill3_kernel<C, LPP, MODE, FREE, MULTI, SK>, 32,
                                                          0) != cudaSuccess || o < 1)
            o = 8;
        occ = o;
    }
    const int grid = (int)std::min<int64_t>(groups, (int64_t)sm_count * std::min(occ.load(), ctas_per_sm));
    affine_fill3_kernel<C, LPP, MODE, FREE, MULTI, SK><<<grid, 32, 0, st>>>(fp);
}

template <int C, int LPP, int MODE, bool FREE, bool MULTI>
void launch_fill3(const FillParams &fp, int64_t groups, int sm_count, int ctas_per_sm, cudaStream_t st, int skew)
{
    (void)skew; // the two-row skew (SK = 2) variant measured no faster and is not instantiated
    launch_fill3_sk<C, LPP, MODE, FREE, MULTI, 1>(fp, groups, sm_count, ctas_per_sm, st);
}

template <int C, int LPP, bool MULTI>
void dispatch_fill3_t(const Problem &pb, const FillParams &fp, int64_t groups, int sms, int cps, cudaStream_t st)
{
    const int mode = !pb.tagged ? 0 : (pb.want_cigar ? 2 : 1);
    if (pb.kind == 1) {
        if (mode == 0)
            launch_fill3<C, LPP, 0, true, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
        else if (mode == 1)
            launch_fill3<C, LPP, 1, true, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
        else
            launch_fill3<C, LPP, 2, true, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
    } else {
        if (mode == 0)
            launch_fill3<C, LPP, 0, false, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
        else if (mode == 1)
            launch_fill3<C, LPP, 1, false, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
        else
            launch_fill3<C, LPP, 2, false, MULTI>(fp, groups, sms, cps, st, pb.cfg.skew);
    }
}

template <bool FREE, int CM, bool TB = false>
void launch_fill16(const FillParams &fp, int64_t quads, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    static std::atomic<int> occ{0}; // per instantiation; contexts on different threads may race to fill it
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_fill16_kernel<FREE, CM, false, TB>, 32, 0) != cudaSuccess ||
            o < 1)
            o = 8;
        occ = o;
    }
    const int grid = (int)std::min<int64_t>(quads, (int64_t)sm_count * std::min(occ.load(), ctas_per_sm));
    affine_fill16_kernel<FREE, CM, false, TB><<<grid, 32, 0, st>>>(fp);
}

inline int64_t ckpt_quad_words(int64_t n) { return ck_count(n) * kCkRegs * 32; }

template <int CM, bool TB>
void launch_fill16_ckpt_t(const FillParams &fp, int64_t quads, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    static std::atomic<int> occ{0};
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_fill16_kernel<true, CM, true, TB>, 32, 0) != cudaSuccess ||
            o < 1)
            o = 8;
        occ = o;
    }
    const int grid = (int)std::min<int64_t>(quads, (int64_t)sm_count * std::min(occ.load(), ctas_per_sm));
    affine_fill16_kernel<true, CM, true, TB><<<grid, 32, 0, st>>>(fp);
}

template <bool TB>
void launch_fill16_ckpt_b(const FillParams &fp, int64_t quads, int cm, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    switch (cm) {
    case 0: launch_fill16_ckpt_t<0, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 1: launch_fill16_ckpt_t<1, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 2: launch_fill16_ckpt_t<2, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 3: launch_fill16_ckpt_t<3, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 4: launch_fill16_ckpt_t<4, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 5: launch_fill16_ckpt_t<5, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 6: launch_fill16_ckpt_t<6, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 7: launch_fill16_ckpt_t<7, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 8: launch_fill16_ckpt_t<8, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    default: launch_fill16_ckpt_t<9, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    }
}

void launch_fill16_ckpt(const FillParams &fp, int64_t quads, int cm, int sm_count, int ctas_per_sm, cudaStream_t st, bool tb)
{
    if (tb)
        launch_fill16_ckpt_b<true>(fp, quads, cm, sm_count, ctas_per_sm, st);
    else
        launch_fill16_ckpt_b<false>(fp, quads, cm, sm_count, ctas_per_sm, st);
}

// second pass of the checkpoint path: recompute + walk (pass 0: slots and counts; pass 1: overflowing pairs)
void launch_ckpt_trace(gnx_ctx *ctx, const Problem &pb, const FillParams &fp, const uint32_t *ckpt, const int64_t *rstar,
                       uint32_t *slots, int *counts, int pass, const int64_t *cig_off, gnx_cigar *cigars, int64_t cap,
                       int *work, int *work_count, cudaStream_t st, const int *pair_slot = nullptr, const int64_t *quad_ck_off = nullptr)
{
    static std::atomic<int> occ{0};
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_ckpt_trace_kernel, 32, 0) != cudaSuccess || o < 1)
            o = 8;
        occ = o;
    }
    CkptParams q;
    memset(&q, 0, sizeof q);
    q.ckpt = ckpt;
    q.quad_words = ckpt_quad_words(pb.cfg.n_uniform);
    q.rstar = rstar;
    q.slots = slots;
    q.slot_cap = slot_cap_of(pb);
    q.counts = counts;
    q.pass = pass;
    q.cigar_off = cig_off;
    q.out_cigar = (CigarOut *)cigars;
    q.out_cap = cap;
    q.h00_plane = pb.h00_plane;
    q.work = work;
    q.work_count = work_count;
    q.work_next = work_count + 1;
    q.pair_slot = pair_slot;
    q.quad_ck_off = quad_ck_off;
    const int64_t np = fp.pair_end - fp.pair_begin;
    cudaMemsetAsync(work_count, 0, 2 * sizeof(int), st); // list length and cursor
    if (pass == 0) // screening: indel-free routes are written directly, the rest is queued
        ckpt_classify_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(fp, q);
    else           // pairs whose cigar overflowed the slot (and fits the output)
        ckpt_overflow_list_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(counts, cig_off, np, q.slot_cap, cap, work, work_count);
    ctx->launches++;
    // persistent grid: every half-warp keeps taking queued pairs until the list is empty
    const int grid = (int)std::min<int64_t>((np + 1) / 2, (int64_t)ctx->sm_count * occ.load());
    affine_ckpt_trace_kernel<<<grid, 32, 0, st>>>(fp, q);
}

// freeEndGaps: the in-lane index of the last column is a template parameter (uniform batches: one value per call)
template <bool TB>
void launch_fill16_free_b(const FillParams &fp, int64_t quads, int cm, int sm_count, int ctas_per_sm, cudaStream_t st)
{
    switch (cm) {
    case 0: launch_fill16<true, 0, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 1: launch_fill16<true, 1, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 2: launch_fill16<true, 2, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 3: launch_fill16<true, 3, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 4: launch_fill16<true, 4, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 5: launch_fill16<true, 5, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 6: launch_fill16<true, 6, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 7: launch_fill16<true, 7, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    case 8: launch_fill16<true, 8, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    default: launch_fill16<true, 9, TB>(fp, quads, sm_count, ctas_per_sm, st); break;
    }
}

void launch_fill16_free(const FillParams &fp, int64_t quads, int cm, int sm_count, int ctas_per_sm, cudaStream_t st, bool tb)
{
    if (tb)
        launch_fill16_free_b<true>(fp, quads, cm, sm_count, ctas_per_sm, st);
    else
        launch_fill16_free_b<false>(fp, quads, cm, sm_count, ctas_per_sm, st);
}

// CTA-per-pair kernel for long pairs (affine_fill3w_kernel): persistent grid over the chunk's pairs.
template <int MODE, bool FREE>
void launch_fill3w(const FillParams &fp, int64_t np, int sm_count, cudaStream_t st)
{
    constexpr int NW = 4;
    static std::atomic<int> occ{0};
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_fill3w_kernel<10, MODE, FREE, NW>, 32 * NW, 0) != cudaSuccess ||
            o < 1)
            o = 2;
        occ = o;
    }
    const int grid = (int)std::min<int64_t>(np, (int64_t)sm_count * occ.load());
    affine_fill3w_kernel<10, MODE, FREE, NW><<<grid, 32 * NW, 0, st>>>(fp);
}

void dispatch_fill3w(const Problem &pb, const FillParams &fp, int64_t np, int sms, cudaStream_t st)
{
    const int mode = !pb.tagged ? 0 : (pb.want_cigar ? 2 : 1);
    if (pb.kind == 1) {
        if (mode == 0)
            launch_fill3w<0, true>(fp, np, sms, st);
        else if (mode == 1)
            launch_fill3w<1, true>(fp, np, sms, st);
        else
            launch_fill3w<2, true>(fp, np, sms, st);
    } else {
        if (mode == 0)
            launch_fill3w<0, false>(fp, np, sms, st);
        else if (mode == 1)
            launch_fill3w<1, false>(fp, np, sms, st);
        else
            launch_fill3w<2, false>(fp, np, sms, st);
    }
}

template <int LPP, bool STORE, bool MULTI, int EXT>
void launch_const3(const FillParams &fp, int64_t groups, int sms, int cps, cudaStream_t st)
{
    static std::atomic<int> occ{0};
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, const_fill3_kernel<10, LPP, STORE, MULTI, EXT>, 32, 0) !=
                cudaSuccess || o < 1)
            o = 8;
        occ = o;
    }
    const int grid = (int)std::min<int64_t>(groups, (int64_t)sms * std::min(occ.load(), cps));
    const_fill3_kernel<10, LPP, STORE, MULTI, EXT><<<grid, 32, 0, st>>>(fp);
}

template <int EXT>
void dispatch_const3_e(const Problem &pb, const FillParams &fp, int64_t groups, int sms, int cps, cudaStream_t st)
{
    const FillCfg &c = pb.cfg;
    if (pb.want_cigar) {
        if (c.lpp == 16)
            launch_const3<16, true, false, EXT>(fp, groups, sms, cps, st);
        else if (c.multi)
            launch_const3<32, true, true, EXT>(fp, groups, sms, cps, st);
        else
            launch_const3<32, true, false, EXT>(fp, groups, sms, cps, st);
    } else {
        if (c.lpp == 16)
            launch_const3<16, false, false, EXT>(fp, groups, sms, cps, st);
        else if (c.multi)
            launch_const3<32, false, true, EXT>(fp, groups, sms, cps, st);
        else
            launch_const3<32, false, false, EXT>(fp, groups, sms, cps, st);
    }
}

void dispatch_const3(const Problem &pb, const FillParams &fp, int64_t groups, int sms, int cps, cudaStream_t st)
{
    if (pb.ext == 1)
        dispatch_const3_e<1>(pb, fp, groups, sms, cps, st);
    else if (pb.ext == 2)
        dispatch_const3_e<2>(pb, fp, groups, sms, cps, st);
    else
        dispatch_const3_e<0>(pb, fp, groups, sms, cps, st);
}

void dispatch_fill3(const Problem &pb, const FillParams &fp, int64_t groups, int sms, int cps, cudaStream_t st)
{
    const FillCfg &c = pb.cfg;
    if (pb.kind == 2) {
        dispatch_const3(pb, fp, groups, sms, cps, st);
        return;
    }
    if (c.lpp == 16)
        dispatch_fill3_t<10, 16, false>(pb, fp, groups, sms, cps, st);
    else if (c.multi)
        dispatch_fill3_t<10, 32, true>(pb, fp, groups, sms, cps, st);
    else
        dispatch_fill3_t<10, 32, false>(pb, fp, groups, sms, cps, st);
}

void dispatch_fill(const Problem &pb, const FillParams &fp, int C, int lookup, int grid, cudaStream_t st)
{
    if (lookup == 2) { // AffineGapChunk: global affine with traceback only
        if (C == 5)
            launch_affine<5, true, false, 2>(fp, grid, st);
        else
            launch_affine<10, true, false, 2>(fp, grid, st);
        return;
    }
    if (C == 5) {
        if (lookup == 0)
            dispatch_fill_cl<5, 0>(pb, fp, grid, st);
        else
            dispatch_fill_cl<5, 1>(pb, fp, grid, st);
    } else {
        if (lookup == 0)
            dispatch_fill_cl<10, 0>(pb, fp, grid, st);
        else
            dispatch_fill_cl<10, 1>(pb, fp, grid, st);
    }
}

// Tile-checkpoint kernel for long pairs (gnx_long.cuh): one persistent launch per chunk, pairs taken from a device
// counter.  pass 0: scores, cigar slots and counts; pass 1: pairs whose cigar overflowed the slot, re-run entirely.
template <bool FREE, int FORM>
int launch_long_t(gnx_ctx *ctx, const Problem &pb, const FillParams &fp, LongParams q, cudaStream_t st)
{
    static std::atomic<int> occ{0};
    if (occ == 0) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, affine_long_kernel<FREE, FORM>, 32, 0) != cudaSuccess || o < 1)
            o = 8;
        occ = o;
    }
    const LongGeom g = long_geom(pb.cfg.n_uniform, pb.cfg.m_uniform);
    const int64_t np = fp.pair_end - fp.pair_begin;
    int64_t grid = std::min<int64_t>(np, (int64_t)ctx->sm_count * std::min(occ.load(), ctx->opt_ctas_per_sm));
    const int64_t fit = ((int64_t)ctx->workspace - 256) / g.cta_stride;
    if (fit < 1)
        return fail(ctx, GNX_ERANGE, "one long pair's checkpoints exceed the context workspace");
    grid = std::max<int64_t>(1, std::min(grid, fit));
    CU(ctx->long_scratch.ensure((size_t)(256 + grid * g.cta_stride)));
    int *counter = ctx->long_scratch.as<int>() + ctx->long_rr;
    ctx->long_rr = (ctx->long_rr + 1) % 64;
    q.scratch = ctx->long_scratch.as<uint8_t>() + 256;
    q.cta_stride = g.cta_stride;
    q.edge_stride = g.edge_stride;
    q.ckpt_off = g.ckpt_off;
    q.ckpt_strip_words = g.ckpt_strip_words;
    q.tile_off = g.tile_off;
    q.runs_off = g.runs_off;
    q.next_pair = counter;
    if (ctx->ev_long_set) // the scratch is shared by the slots' streams: chain the launches
        CU(cudaStreamWaitEvent(st, ctx->ev_long, 0));
    CU(cudaMemsetAsync(counter, 0, sizeof(int), st));
    affine_long_kernel<FREE, FORM><<<(int)grid, 32, 0, st>>>(fp, q);
    CU(cudaEventRecord(ctx->ev_long, st));
    ctx->ev_long_set = true;
    ctx->launches++;
    return GNX_OK;
}

int launch_long(gnx_ctx *ctx, const Problem &pb, const FillParams &fp, uint32_t *slots, int *counts, int pass,
                const int64_t *cig_off, gnx_cigar *cigars, int64_t cap, cudaStream_t st)
{
    LongParams q;
    memset(&q, 0, sizeof q);
    const int64_t np = fp.pair_end - fp.pair_begin;
    q.pool_cursor = reinterpret_cast<unsigned long long *>(slots); // layout: long_slot_bytes()
    q.slot64 = reinterpret_cast<long long *>(slots) + 2;
    q.pool = slots + 4 + 2 * np;
    q.pool_cap = pool_entries_of(pb, np);
    if (pass == 0)
        CU(cudaMemsetAsync(q.pool_cursor, 0, 8, st));
    q.counts = counts;
    q.pass = pass;
    q.cigar_off = cig_off;
    q.out_cigar = (CigarOut *)cigars;
    q.out_cap = cap;
    q.h00_plane = pb.h00_plane;
    const bool f1 = ctx->opt_long_form != 0;
    if (pb.kind == 1)
        return f1 ? launch_long_t<true, 1>(ctx, pb, fp, q, st) : launch_long_t<true, 0>(ctx, pb, fp, q, st);
    return f1 ? launch_long_t<false, 1>(ctx, pb, fp, q, st) : launch_long_t<false, 0>(ctx, pb, fp, q, st);
}

// 2-bit inputs of a host batch (gnx_affine_batch_twobit): dnaTwoBit words, tightly packed
struct TbIn {
    const uint64_t *a_words = nullptr, *b_words = nullptr;
    const int64_t *a_woff = nullptr, *b_woff = nullptr; // n_pairs + 1 word offsets (host; computed by the entry point)
    const int64_t *a_len = nullptr, *b_len = nullptr;   // the caller's length arrays
    bool uniform = false;           // every pair n x m
    int64_t n = 0, m = 0, wn = 0, wm = 0; // uniform: lengths and words per sequence
    // from_bytes: the caller passed one byte per base (gnx_affine_batch).  Pageable input has to be copied through a
    // page-locked stage anyway; that pass packs it to dnaTwoBit words instead (a quarter of the bytes to write, to DMA
    // and to read on the device) and the chunk then runs exactly like a gnx_affine_batch_twobit chunk.
    // Uniform batches only (word offsets are p * wn / p * wm, a_woff / b_woff stay NULL).
    bool from_bytes = false;
};
constexpr int GNX_RETRY_BYTES = -1000; // internal: a from_bytes chunk met a base >= 4 -- redo the call on the byte path

// Device buffers of one chunk, all addressed with GLOBAL pair indices (pointers are pre-biased).
struct ChunkDev {
    const uint8_t *alpha, *beta;      // biased so that absolute offsets index them
    const int64_t *aoff, *boff;       // biased: aoff[pair] valid for pair in [begin, end]
    uint8_t *cls;                     // biased by -begin
    uint32_t *trace;
    const int64_t *trace_off;         // chunk-local index
    int2 *edge;
    int64_t edge_stride;
    uint32_t *slots;
    int *counts;                      // chunk-local
    int64_t *score;                   // biased by -begin (global pair index)
    int64_t a_lo, a_hi, b_lo, b_hi;   // absolute byte ranges of the chunk inside alpha / beta
    int64_t *best;                    // ext 2: biased by -begin (global pair index)
    int64_t *end_i, *end_j;           // ext: chunk-local
    int *work, *work_count;           // checkpoint path: work list (chunk-local pair indices) and its device counter
    unsigned *quad_ctr;               // affine_fill16_kernel: quad cursor of this slot (nullptr: static round-robin)
    const uint64_t *alpha_words, *beta_words; // TB kernels: biased so that words + pair * wn / wm is the pair's sequence
    // ragged batches on the packed 16-bit kernels (RagTables): device copies + the quad range of every last-column group
    const int *quad_pairs, *pair_slot;
    const int64_t *quad_ck_off;
    int64_t cm_first[12];
    const int *smat;                  // profile batches: biased so that smat + smat_off[pair] is the pair's matrix
    const int64_t *smat_off;          // indexed by global pair id
};

// The slot's quad cursor: allocated and zeroed once; every affine_fill16_kernel launch leaves it at zero again.
cudaError_t slot_quad_ctr(Slot &s, ChunkDev &cd)
{
    if (!s.qctr.p) {
        cudaError_t e = s.qctr.ensure(64);
        if (e != cudaSuccess)
            return e;
        e = cudaMemset(s.qctr.p, 0, 64);
        if (e != cudaSuccess)
            return e;
    }
    cd.quad_ctr = s.qctr.as<unsigned>();
    return cudaSuccess;
}

// Ragged batches on the packed 16-bit kernels.  A quad's four pairs advance in lock step, so they must share the
// target length n (the step count) and -- with free end gaps -- the in-lane index (m - 1) % 10 of the last query column
// (a template parameter of the kernels); query lengths are otherwise free.  The host bins a chunk's pairs by
// (that index, n) with a counting sort and cuts every bin into quads; the last quad of a bin may have empty slots.
struct RagTables {
    std::vector<int> quad_pairs, pair_slot, count, keys, first_slot;
    std::vector<int64_t> quad_ck_off;
    int64_t cm_first[12];
    int64_t n_quads = 0, ck_words = 0;
};

void build_rag_tables(const Problem &pb, const int64_t *aoff, const int64_t *boff, int64_t begin, int64_t np, int64_t max_n,
                      RagTables &R)
{
    const bool free_end = pb.kind == 1;
    const int64_t stride = max_n + 1, groups = free_end ? 10 : 1, nkeys = groups * stride;
    // inside a group the quads run from the LONGEST targets to the shortest: the persistent grid takes quads
    // round-robin, so the launch ends on short quads instead of waiting for a straggler (key = ... + max_n - n)
    R.keys.resize((size_t)np);
    R.count.assign((size_t)nkeys + 1, 0);
    for (int64_t k = 0; k < np; ++k) {
        const int64_t n = aoff[begin + k + 1] - aoff[begin + k], m = boff[begin + k + 1] - boff[begin + k];
        const int key = (int)((free_end ? (m - 1) % 10 : 0) * stride + (max_n - n));
        R.keys[(size_t)k] = key;
        R.count[(size_t)key + 1]++;
    }
    // first slot (= 4 x first quad) and checkpoint words of every key
    R.first_slot.resize((size_t)nkeys);
    R.quad_ck_off.clear();
    int64_t nq = 0, words = 0;
    for (int64_t key = 0; key < nkeys; ++key) {
        if (key % stride == 0)
            R.cm_first[key / stride] = nq;
        R.first_slot[(size_t)key] = (int)(nq * 4);
        const int64_t q = (R.count[(size_t)key + 1] + 3) / 4;
        if (q > 0) {
            const int64_t n = max_n - key % stride;
            const int64_t w = pb.cfg.impl == 17 ? ckpt_quad_words(n) : 0;
            for (int64_t i = 0; i < q; ++i) {
                R.quad_ck_off.push_back(words);
                words += w;
            }
            nq += q;
        }
    }
    for (int64_t g = groups; g < 12; ++g)
        R.cm_first[g] = nq;
    R.quad_ck_off.push_back(words);
    R.n_quads = nq;
    R.ck_words = words;
    R.quad_pairs.assign((size_t)nq * 4, -1);
    R.pair_slot.resize((size_t)np);
    for (int64_t k = 0; k < np; ++k) {
        const int slot = R.first_slot[(size_t)R.keys[(size_t)k]]++;
        R.quad_pairs[(size_t)slot] = (int)k;
        R.pair_slot[(size_t)k] = slot;
    }
}

// Tables of every chunk of a plan, built up front by a few host threads (a chunk's tables take ~1 ms of scalar work
// per 100k pairs: built one chunk at a time in front of the launches they would throttle the pipeline).
void build_all_rag_tables(const Problem &pb, const int64_t *aoff, const int64_t *boff, const std::vector<int64_t> &bounds,
                          int64_t max_n, std::vector<RagTables> &all)
{
    const int64_t nc = (int64_t)bounds.size() - 1;
    all.resize((size_t)nc);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)8, (int64_t)std::thread::hardware_concurrency(), nc}));
    std::atomic<int64_t> next{0};
    auto work = [&] {
        for (int64_t c = next++; c < nc; c = next++)
            build_rag_tables(pb, aoff, boff, bounds[(size_t)c], bounds[(size_t)c + 1] - bounds[(size_t)c], max_n, all[(size_t)c]);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        th.emplace_back(work);
    work();
    for (auto &x : th)
        x.join();
}

// Upload a chunk's RagTables into the slot's device buffer and point the chunk at them.
int upload_rag_tables(gnx_ctx *ctx, Slot &s, PinBuf &stage, const RagTables &R, int64_t np, ChunkDev &cd, cudaStream_t st)
{
    const size_t b_qp = (size_t)R.n_quads * 4 * sizeof(int), b_off = ((size_t)R.n_quads + 1) * 8, b_ps = (size_t)np * sizeof(int);
    const size_t o_off = (b_qp + 15) & ~(size_t)15, o_ps = (o_off + b_off + 15) & ~(size_t)15, total = o_ps + b_ps;
    CU(s.rag.ensure(total + 16));
    CU(stage.ensure(total + 16));
    uint8_t *h = stage.as<uint8_t>();
    memcpy(h, R.quad_pairs.data(), b_qp);
    memcpy(h + o_off, R.quad_ck_off.data(), b_off);
    memcpy(h + o_ps, R.pair_slot.data(), b_ps);
    CU(cudaMemcpyAsync(s.rag.p, h, total, cudaMemcpyHostToDevice, st));
    cd.quad_pairs = s.rag.as<int>();
    cd.quad_ck_off = reinterpret_cast<const int64_t *>(s.rag.as<uint8_t>() + o_off);
    cd.pair_slot = reinterpret_cast<const int *>(s.rag.as<uint8_t>() + o_ps);
    memcpy(cd.cm_first, R.cm_first, sizeof cd.cm_first);
    return GNX_OK;
}

void launch_traceback_ext(const Problem &pb, const ChunkDev &cd, const TraceParams &tp, int64_t np, cudaStream_t st)
{
    ExtTraceParams ep;
    memset(&ep, 0, sizeof ep);
    ep.t = tp;
    ep.side = pb.ext;
    ep.local = pb.ext_local ? 1 : 0;
    ep.alpha = cd.alpha;
    ep.beta = cd.beta;
    ep.dim = pb.dim;
    ep.gap = (int)pb.gap_open;
    for (int i = 0; i < pb.dim * pb.dim; ++i)
        ep.scores[i] = (int)pb.scores[i];
    ep.score = cd.score;
    ep.best = cd.best;
    ep.end_i = cd.end_i;
    ep.end_j = cd.end_j;
    traceback_ext_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(ep);
}

// classify + fill (+ traceback pass 0) for chunk [begin, end) on stream st.
int enqueue_chunk_compute(gnx_ctx *ctx, const Problem &pb, const ChunkDev &cd, int64_t begin, int64_t end,
                          cudaStream_t st)
{
    const int C = pb.cfg.C;
    const bool any_long = pb.cfg.multi;
    const int64_t np = end - begin;
    if (np <= 0)
        return GNX_OK;
    int *status = ctx->status.as<int>();
    const int warps_per_block = 4;
    const int max_grid = ctx->sm_count * ctx->opt_blocks_per_sm;
    const int grid = (int)std::min<int64_t>((np + warps_per_block - 1) / warps_per_block, max_grid);
    if (pb.profile) {
        // no sequence bytes on this path: invalid bases were found while the profiles were built
    } else if (pb.twobit) {
        cudaMemsetAsync(cd.cls + begin, 0, (size_t)np, st); // two bits per base: every base is < 4 <= dim
    } else if (pb.cfg.impl == 3 || pb.cfg.impl == 16 || pb.cfg.impl == 17 || pb.cfg.impl == 18) {
        // these kernels take any base < dim, so the per-pair pass is only needed to find WHICH pair is
        // invalid; gate it on the chunk's largest base (vectorised, HBM-bound)
        int *gate = status + 1 + ctx->gate_rr;
        ctx->gate_rr = (ctx->gate_rr + 1) % 8;
        cudaMemsetAsync(gate, 0, sizeof(int), st);
        cudaMemsetAsync(cd.cls + begin, 0, (size_t)np, st);
        const int mb_grid = ctx->sm_count * 4;
        maxbase_kernel<<<mb_grid, 256, 0, st>>>(cd.alpha, cd.a_lo, cd.a_hi, gate);
        maxbase_kernel<<<mb_grid, 256, 0, st>>>(cd.beta, cd.b_lo, cd.b_hi, gate);
        classify_kernel<<<grid, 128, 0, st>>>(cd.alpha, cd.aoff, cd.beta, cd.boff, begin, end, pb.dim, cd.cls, status,
                                              gate);
        ctx->launches += 3;
    } else {
        classify_kernel<<<grid, 128, 0, st>>>(cd.alpha, cd.aoff, cd.beta, cd.boff, begin, end, pb.dim, cd.cls, status,
                                              nullptr);
        ctx->launches++;
    }

    FillParams fp;
    memset(&fp, 0, sizeof fp);
    fp.alpha = cd.alpha;
    fp.alpha_off = cd.aoff;
    fp.beta = cd.beta;
    fp.beta_off = cd.boff;
    fp.pair_begin = begin;
    fp.pair_end = end;
    fp.pair_class = pb.profile ? nullptr : cd.cls;
    fp.smat = cd.smat;
    fp.smat_off = cd.smat_off;
    fp.out_best = cd.best;
    fp.gap_open = (int)pb.gap_open;
    fp.gap_extend = (int)pb.gap_extend;
    fp.h00 = pb.h00;
    fp.dim = pb.dim;
    for (int i = 0; i < pb.dim * pb.dim; ++i)
        fp.scores[i] = (int)pb.scores[i];
    fp.trace = pb.want_cigar ? cd.trace : nullptr; // tagged score-only runs keep the arithmetic, skip the stores
    fp.trace_off = cd.trace_off;
    fp.edge = any_long ? cd.edge : nullptr;
    fp.edge_stride = cd.edge_stride;
    fp.out_score = cd.score;
    fp.one = 1;
    fp.chunk = (int)pb.chunk;
    fp.quad_ctr = cd.quad_ctr;
    if (pb.cfg.tb) {
        fp.alpha_words = cd.alpha_words;
        fp.beta_words = cd.beta_words;
        fp.n_uni = (int)pb.cfg.n_uniform;
        fp.m_uni = (int)pb.cfg.m_uniform;
        fp.wn = (int)((pb.cfg.n_uniform + 31) / 32);
        fp.wm = (int)((pb.cfg.m_uniform + 31) / 32);
    }

    // class 0 (ACGT only) with the PRMT tables when the matrix fits 16 bits, else the smem lookup
    FillEvent &fe = next_fill_event(ctx);
    cudaEventRecord(fe.a, st);
    const int lookup0 = pb.chunk > 1 ? 2 : ((ctx->opt_force_lookup == 1 || !pb.prmt_ok) ? 1 : 0);
    if (pb.cfg.rag) {
        fp.quad_pairs = cd.quad_pairs;
        fp.quad_ck_off = cd.quad_ck_off;
    }
    if (pb.cfg.rag && (pb.cfg.impl == 16 || pb.cfg.impl == 17)) {
        // one launch per last-column group (free end gaps: the index is a template parameter), each over its own quads
        fp.trace = cd.trace;
        const int groups = pb.kind == 1 ? 10 : 1;
        for (int g = 0; g < groups; ++g) {
            fp.quad_first = cd.cm_first[g];
            fp.n_quads = cd.cm_first[g + 1] - cd.cm_first[g];
            if (fp.n_quads <= 0)
                continue;
            if (pb.cfg.impl == 17)
                launch_fill16_ckpt(fp, fp.n_quads, g, ctx->sm_count, ctx->opt_ctas_per_sm, st, false);
            else if (pb.kind == 1)
                launch_fill16_free(fp, fp.n_quads, g, ctx->sm_count, ctx->opt_ctas_per_sm, st, false);
            else
                launch_fill16<false, -1>(fp, fp.n_quads, ctx->sm_count, ctx->opt_ctas_per_sm, st);
            ctx->launches++;
            ctx->last_fill_launches++;
        }
    } else if (pb.cfg.impl == 18) { // long pairs: score, checkpoints, recompute and walk in one persistent launch
        const int rc = launch_long(ctx, pb, fp, cd.slots, cd.counts, 0, nullptr, nullptr, 0, st);
        if (rc != GNX_OK)
            return rc;
        ctx->last_fill_launches++;
    } else if (pb.cfg.impl == 17) {
        const int64_t quads = (np + 3) / 4;
        fp.trace = cd.trace;                                   // checkpoint area
        fp.edge_stride = ckpt_quad_words(pb.cfg.n_uniform);    // words per quad
        launch_fill16_ckpt(fp, quads, (int)((pb.cfg.m_uniform - 1) % 10), ctx->sm_count, ctx->opt_ctas_per_sm, st, pb.cfg.tb);
        ctx->launches++;
        ctx->last_fill_launches++;
    } else if (pb.cfg.impl == 16) {
        const int64_t quads = (np + 3) / 4;
        if (pb.kind == 1)
            launch_fill16_free(fp, quads, (int)((pb.cfg.m_uniform - 1) % 10), ctx->sm_count, ctx->opt_ctas_per_sm, st, pb.cfg.tb);
        else if (pb.cfg.tb)
            launch_fill16<false, -1, true>(fp, quads, ctx->sm_count, ctx->opt_ctas_per_sm, st);
        else
            launch_fill16<false, -1>(fp, quads, ctx->sm_count, ctx->opt_ctas_per_sm, st);
        ctx->launches++;
        ctx->last_fill_launches++;
    } else if (pb.cfg.impl == 3) {
        // one launch: the per-lane score tables cover every base < dim, so there is no class split
        const int64_t groups = (np + (32 / pb.cfg.lpp) - 1) / (32 / pb.cfg.lpp);
        // long pairs with few pairs in flight (the traceback matrix of a 10 kb x 10 kb pair is 82 MB, so a chunk
        // holds ~1000 of them and the one-warp kernel leaves the SMs short of warps): four warps share one pair
        // (affine_fill3w_kernel).  Measured at 10 kb x 10 kb (profiles/r01g_kbench_c4.txt): a wave of SMs x 3 pairs
        // takes 0.75 x the time the one-warp kernel needs for its first pair, which then grows by 1/(2.6 waves)
        // per extra pair -- so the CTA-per-pair kernel wins for up to one wave and again near two full waves.
        const int64_t wave = (int64_t)ctx->sm_count * 3;
        const int64_t waves = (np + wave - 1) / wave;
        const bool w_faster = 0.75 * (double)waves < 1.0 + (double)np / (2.6 * (double)wave);
        const bool wide_cta = pb.kind != 2 && pb.cfg.multi &&
                              (ctx->opt_wide_cta == 1 || (ctx->opt_wide_cta < 0 && pb.cfg.strips_max >= 4 && w_faster));
        if (wide_cta)
            dispatch_fill3w(pb, fp, np, ctx->sm_count, st);
        else
            dispatch_fill3(pb, fp, groups, ctx->sm_count, ctx->opt_ctas_per_sm, st);
        ctx->launches++;
        ctx->last_fill_launches++;
    } else if (pb.profile) {
        fp.want_class = -1;
        if (pb.tagged) {
            if (C == 5)
                launch_affine<5, true, false, 3>(fp, grid, st);
            else
                launch_affine<10, true, false, 3>(fp, grid, st);
        } else {
            if (C == 5)
                launch_affine<5, false, false, 3>(fp, grid, st);
            else
                launch_affine<10, false, false, 3>(fp, grid, st);
        }
        ctx->launches++;
        ctx->last_fill_launches++;
    } else if (pb.wide) {
        // int64 plane values: one launch over every valid pair (shared-memory score lookup)
        fp.want_class = -1;
        if (pb.kind == 1)
            affine_fill_kernel<5, true, true, 1, long long><<<grid, 128, 0, st>>>(fp);
        else
            affine_fill_kernel<5, true, false, 1, long long><<<grid, 128, 0, st>>>(fp);
        ctx->launches++;
        ctx->last_fill_launches++;
    } else {
        fp.want_class = 0;
        dispatch_fill(pb, fp, C, lookup0, grid, st);
        ctx->launches++;
        ctx->last_fill_launches++;
        // class 1 (contains N or other bases < dim): generic lookup.  Warps skip pairs of the other class.
        fp.want_class = 1;
        dispatch_fill(pb, fp, C, pb.chunk > 1 ? 2 : 1, grid, st);
        ctx->launches++;
        ctx->last_fill_launches++;
    }
    if (!(pb.cfg.impl == 17 && pb.want_cigar)) // checkpoint path: the re-fill of the path's blocks is part of the DP fill
        cudaEventRecord(fe.b, st);

    if (pb.want_cigar && pb.cfg.impl != 18) {
        TraceParams tp;
        memset(&tp, 0, sizeof tp);
        tp.alpha_off = cd.aoff;
        tp.beta_off = cd.boff;
        tp.pair_begin = begin;
        tp.pair_end = end;
        tp.trace = cd.trace;
        tp.trace_off = cd.trace_off;
        tp.C = C;
        tp.layout = pb.cfg.impl == 3 ? 3 : 1;
        tp.lpp = pb.cfg.lpp;
        tp.skew = pb.cfg.skew;
        tp.chunk = (int)pb.chunk;
        tp.kind = pb.kind == 2 ? 2 : 0;
        tp.h00_plane = pb.h00_plane;
        tp.slots = cd.slots;
        tp.slot_cap = slot_cap_of(pb);
        tp.counts = cd.counts;
        tp.pass = 0;
        tp.pair_class = pb.profile ? nullptr : cd.cls;
        if (pb.cfg.impl == 17) {
            launch_ckpt_trace(ctx, pb, fp, cd.trace, cd.best, cd.slots, cd.counts, 0, nullptr, nullptr, 0, cd.work, cd.work_count, st,
                              pb.cfg.rag ? cd.pair_slot : nullptr, pb.cfg.rag ? cd.quad_ck_off : nullptr);
            cudaEventRecord(fe.b, st);
            ctx->last_fill_launches++;
        } else if (pb.ext)
            launch_traceback_ext(pb, cd, tp, np, st);
        else if (tp.kind == 2 && tp.layout == 3)
            traceback_const3_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
        else if (tp.kind == 0 && tp.layout == 3 && (ctx->opt_tb_impl == 3 || (ctx->opt_tb_impl == 2 && pb.cfg.multi)))
            traceback_affine_warp_kernel<<<(int)((np + 3) / 4), 128, 0, st>>>(tp); // long pairs: a warp per pair
        else if (tp.kind == 0 && tp.layout >= 2 && ctx->opt_tb_impl >= 2)
            traceback_affine_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
        else
            traceback_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
        ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("kernel launch failed: ") + cudaGetErrorString(e);
        return GNX_ECUDA;
    }
    return GNX_OK;
}

// Exclusive scan of the chunk's cigar counts into `off` (np+1 entries), continuing from *running_total.
// When `advance` is set the running total is moved forward on the device (device-resident batches).
int enqueue_scan(gnx_ctx *ctx, DevBuf &partials, const int *counts, int64_t np, int64_t *off,
                 const int64_t *running_total, int64_t *total_out, cudaStream_t st)
{
    const int blocks = (int)((np + kScanSeg - 1) / kScanSeg);
    if (partials.ensure((size_t)blocks * 8) != cudaSuccess) {
        ctx->err = "cudaMalloc failed (scan partials)";
        return GNX_ECUDA;
    }
    scan_partial_kernel<<<blocks, 256, 0, st>>>(counts, np, partials.as<int64_t>());
    // total_out must not alias running_total: other blocks may still be reading it
    scan_apply_kernel<<<blocks, 256, 0, st>>>(counts, np, partials.as<int64_t>(), off, running_total, total_out);
    ctx->launches += 2;
    return GNX_OK;
}

// expand slots (+ overflow traceback pass) into cigars[] at cig_off (chunk-local offsets) .
int enqueue_chunk_expand(gnx_ctx *ctx, const Problem &pb, const ChunkDev &cd, int64_t begin, int64_t end,
                         const int64_t *cig_off, gnx_cigar *cigars, int64_t cap, cudaStream_t st)
{
    const int C = pb.cfg.C;
    const int64_t np = end - begin;
    if (np <= 0)
        return GNX_OK;
    int *status = ctx->status.as<int>();
    if (pb.cfg.impl == 18)
        expand_pool_kernel<<<(int)((np + 3) / 4), 128, 0, st>>>(reinterpret_cast<const long long *>(cd.slots) + 2,
                                                               cd.slots + 4 + 2 * np, cd.counts, cig_off, np,
                                                               (CigarOut *)cigars, cap, status);
    else
        expand_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(cd.slots, slot_cap_of(pb), cd.counts, cig_off, np,
                                                              (CigarOut *)cigars, cap, status,
                                                              pb.ext ? (pb.ext_local ? 2 : 1) : 0);
    ctx->launches++;
    TraceParams tp;
    memset(&tp, 0, sizeof tp);
    tp.alpha_off = cd.aoff;
    tp.beta_off = cd.boff;
    tp.pair_begin = begin;
    tp.pair_end = end;
    tp.trace = cd.trace;
    tp.trace_off = cd.trace_off;
    tp.C = C;
    tp.layout = pb.cfg.impl == 3 ? 3 : 1;
    tp.lpp = pb.cfg.lpp;
    tp.skew = pb.cfg.skew;
    tp.chunk = (int)pb.chunk;
    tp.kind = pb.kind == 2 ? 2 : 0;
    tp.h00_plane = pb.h00_plane;
    tp.slots = cd.slots;
    tp.slot_cap = slot_cap_of(pb);
    tp.counts = cd.counts;
    tp.cigar_off = const_cast<int64_t *>(cig_off);
    tp.out_cigar = cigars;
    tp.out_cap = cap;
    tp.pass = 1;
    tp.pair_class = pb.profile ? nullptr : cd.cls;
    if (pb.cfg.impl == 17 || pb.cfg.impl == 18) {
        FillParams fp;
        memset(&fp, 0, sizeof fp);
        fp.alpha = cd.alpha;
        fp.alpha_off = cd.aoff;
        fp.beta = cd.beta;
        fp.beta_off = cd.boff;
        fp.pair_begin = begin;
        fp.pair_end = end;
        fp.pair_class = cd.cls;
        fp.gap_open = (int)pb.gap_open;
        fp.gap_extend = (int)pb.gap_extend;
        fp.h00 = pb.h00;
        fp.dim = pb.dim;
        for (int i = 0; i < pb.dim * pb.dim; ++i)
            fp.scores[i] = (int)pb.scores[i];
        fp.one = 1;
        fp.out_score = cd.score;
        if (pb.cfg.impl == 18) {
            const int rc = launch_long(ctx, pb, fp, cd.slots, cd.counts, 1, cig_off, cigars, cap, st);
            if (rc != GNX_OK)
                return rc;
            ctx->launches--; // counted below
        } else
            launch_ckpt_trace(ctx, pb, fp, cd.trace, cd.best, cd.slots, cd.counts, 1, cig_off, cigars, cap, cd.work, cd.work_count, st,
                              pb.cfg.rag ? cd.pair_slot : nullptr, pb.cfg.rag ? cd.quad_ck_off : nullptr);
    } else if (pb.ext)
        launch_traceback_ext(pb, cd, tp, np, st);
    else if (tp.kind == 2 && tp.layout == 3)
        traceback_const3_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
    else if (tp.kind == 0 && tp.layout == 3 && (ctx->opt_tb_impl == 3 || (ctx->opt_tb_impl == 2 && pb.cfg.multi)))
        traceback_affine_warp_kernel<<<(int)((np + 3) / 4), 128, 0, st>>>(tp);
    else if (tp.kind == 0 && tp.layout >= 2 && ctx->opt_tb_impl >= 2)
        traceback_affine_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
    else
        traceback_kernel<<<(int)((np + 127) / 128), 128, 0, st>>>(tp);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("kernel launch failed: ") + cudaGetErrorString(e);
        return GNX_ECUDA;
    }
    return GNX_OK;
}

// Greedy chunk planning: consecutive pairs until the traceback matrices fill `budget_words`
// or the chunk reaches opt_chunk_pairs.
struct Plan {
    std::vector<int64_t> bounds;      // chunk boundaries (pair indices), size = chunks+1
    int64_t max_n = 0, max_m = 0, min_n = INT64_MAX, min_m = INT64_MAX, cells = 0;
    bool any_long = false;            // some pair needs more than one strip
};

int make_plan(gnx_ctx *ctx, Problem &pb, const int64_t *aoff, const int64_t *boff, int64_t n_pairs,
              int64_t budget_words, Plan &plan)
{
    { // length statistics (min/max/cells), split over a few host threads for 10^7-pair batches
        struct Stat {
            int64_t max_n = 0, max_m = 0, min_n = INT64_MAX, min_m = INT64_MAX, cells = 0;
            bool bad = false;
        };
        const int nt = n_pairs >= (1 << 20) ? 8 : 1;
        std::vector<Stat> st((size_t)nt);
        auto work = [&](int t) {
            Stat q;
            const int64_t lo = n_pairs * t / nt, hi = n_pairs * (t + 1) / nt;
            for (int64_t p = lo; p < hi; ++p) {
                const int64_t n = aoff[p + 1] - aoff[p], m = boff[p + 1] - boff[p];
                q.bad |= (n < 0) | (m < 0);
                q.max_n = std::max(q.max_n, n);
                q.max_m = std::max(q.max_m, m);
                q.min_n = std::min(q.min_n, n);
                q.min_m = std::min(q.min_m, m);
                q.cells += n * m;
            }
            st[(size_t)t] = q;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t)
            th.emplace_back(work, t);
        work(0);
        for (auto &x : th)
            x.join();
        for (const Stat &q : st) {
            if (q.bad)
                return fail(ctx, GNX_EARG, "offset arrays must be non-decreasing");
            plan.max_n = std::max(plan.max_n, q.max_n);
            plan.max_m = std::max(plan.max_m, q.max_m);
            plan.min_n = std::min(plan.min_n, q.min_n);
            plan.min_m = std::min(plan.min_m, q.min_m);
            plan.cells += q.cells;
        }
    }
    pick_cfg(ctx, pb, plan.max_m, plan.max_n);
    {   // range proof first: it may switch the configuration to the int64 kernel, which changes the trace layout
        const int rc = analyse(ctx, pb, plan.max_n, plan.max_m);
        if (rc != GNX_OK)
            return rc;
    }
    if (!pb.wide && fill16_ok(ctx, pb, plan.min_n, plan.max_n, plan.min_m, plan.max_m)) {
        pb.cfg.impl = 16;
        pb.cfg.C = 10;
        pb.cfg.lpp = 16;
        pb.cfg.skew = 1;
        pb.cfg.multi = false;
        pb.cfg.m_uniform = plan.max_m;
    } else if (!pb.wide && fill16_ok(ctx, pb, plan.min_n, plan.max_n, plan.min_m, plan.max_m, true)) {
        pb.cfg.impl = 17; // fill16 with checkpoints + affine_ckpt_trace_kernel
        pb.cfg.n_uniform = plan.max_n;
        pb.cfg.C = 10;
        pb.cfg.lpp = 16;
        pb.cfg.skew = 1;
        pb.cfg.multi = false;
        pb.cfg.m_uniform = plan.max_m;
    }
    // 2-bit inputs on the packed 16-bit kernels: the quad's words are staged by TMA (targets expanded in 4 x 512 B)
    pb.cfg.rag = (pb.cfg.impl == 16 || pb.cfg.impl == 17) && (plan.min_n != plan.max_n || plan.min_m != plan.max_m);
    pb.cfg.tb = pb.twobit && ctx->opt_tb_tma && (pb.cfg.impl == 16 || pb.cfg.impl == 17) && plan.max_n <= kTbMaxN && !pb.cfg.rag;
    if (pb.cfg.impl == 16 && pb.cfg.tb)
        pb.cfg.n_uniform = plan.max_n;
    // long pairs with traceback: tile checkpoints + recompute of the route's tiles instead of a trace matrix.  The
    // recompute costs ~(351 + kLongR) x 320 cells per strip crossed, i.e. a fraction ~610 / min(n, m) of the pair: worth
    // it from a couple of million cells per pair; a batch of less than two pairs per SM is latency-bound and stays on
    // the CTA-per-pair kernel.
    if (pb.cfg.impl == 3 && pb.cfg.multi && pb.want_cigar && pb.kind != 2 && !pb.wide && !pb.ext && pb.gap_open <= 0 &&
        ctx->opt_long != 0 &&
        (ctx->opt_long == 1 || (plan.cells / n_pairs >= 2000000 && n_pairs >= 2 * (int64_t)ctx->sm_count))) {
        pb.cfg.impl = 18;
        pb.cfg.n_uniform = plan.max_n; // geometry of the per-warp scratch
        pb.cfg.m_uniform = plan.max_m;
        pb.cfg.long_pool = ctx->opt_long_pool;
    }
    plan.any_long = pb.cfg.multi && pb.cfg.impl != 18;
    ctx->last_impl = pb.cfg.impl;
    ctx->last_flags = (pb.cfg.rag ? 1 : 0) | (pb.cfg.tb ? 2 : 0) | (pb.wide ? 4 : 0) | (pb.cfg.multi ? 8 : 0);
    plan.bounds.push_back(0);
    if (!pb.extra_words && plan.min_n == plan.max_n && plan.min_m == plan.max_m) { // uniform batch: chunk bounds are arithmetic
        const int64_t gsz = 32 / pb.cfg.lpp; // pairs that share trace rows
        const int64_t wg = pb.want_cigar ? group_trace_words(pb, (plan.max_n && plan.max_m) ? plan.max_n : 0, plan.max_m) : 0;
        if (wg > budget_words)
            return fail(ctx, GNX_ERANGE, "one pair's traceback matrix exceeds the context workspace");
        int64_t per = ctx->opt_chunk_pairs;
        if (wg > 0)
            per = std::min(per, (budget_words / wg) * gsz);
        per = std::max<int64_t>(4, per & ~int64_t(3)); // whole groups (and whole fill16 quads)
        for (int64_t p = per; p < n_pairs; p += per)
            plan.bounds.push_back(p);
        plan.bounds.push_back(n_pairs);
        return GNX_OK;
    }
    if (!pb.extra_words) {
        // ragged batch whose chunks fit the workspace even if every group were the largest one (read-sized pairs): the
        // bounds are arithmetic as well -- the exact loop below cost 15 ms per 1.5 M pairs of a gsw extension round
        const int64_t gsz = 32 / pb.cfg.lpp;
        const int64_t wg = pb.want_cigar ? group_trace_words(pb, (plan.max_n && plan.max_m) ? plan.max_n : 0, plan.max_m) : 0;
        const int64_t per = std::max<int64_t>(4, ctx->opt_chunk_pairs & ~int64_t(3));
        if (wg <= budget_words && wg * ((per + gsz - 1) / gsz) <= budget_words) {
            for (int64_t p = per; p < n_pairs; p += per)
                plan.bounds.push_back(p);
            plan.bounds.push_back(n_pairs);
            return GNX_OK;
        }
    }
    // exact accounting of the trace words of the chunk being grown (groups of 32/lpp pairs share rows)
    int64_t words = 0, count = 0, prev_n = 0, prev_m = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        int64_t n = aoff[p + 1] - aoff[p], m = boff[p + 1] - boff[p];
        if (n == 0 || m == 0)
            n = 0;
        const int64_t extra = pb.extra_words ? pb.extra_words[p] : 0;
        int64_t w = extra;
        if (pb.want_cigar) {
            if (pb.cfg.lpp == 16 && (count & 1)) // second pair of a group: the group grows to the larger one
                w += group_trace_words(pb, std::max(n, prev_n), std::max(m, prev_m)) -
                     group_trace_words(pb, prev_n, prev_m);
            else
                w += group_trace_words(pb, n, m);
        }
        if (w > budget_words)
            return fail(ctx, GNX_ERANGE, "one pair's traceback matrix exceeds the context workspace");
        if (count > 0 && (words + w > budget_words || count >= ctx->opt_chunk_pairs)) {
            plan.bounds.push_back(p);
            words = 0;
            count = 0;
            w = extra + (pb.want_cigar ? group_trace_words(pb, n, m) : 0);
        }
        words += w;
        ++count;
        prev_n = n;
        prev_m = m;
    }
    plan.bounds.push_back(n_pairs);
    return GNX_OK;
}

int collect_fill_stats(gnx_ctx *ctx)
{
    double ms = 0;
    for (size_t i = 0; i < ctx->fill_events_used; ++i) {
        float f = 0;
        if (cudaEventElapsedTime(&f, ctx->fill_events[i].a, ctx->fill_events[i].b) == cudaSuccess)
            ms += f;
    }
    ctx->last_fill_ms = ms;
    return GNX_OK;
}

// Host threads this process may use for staging, packing and the gsw driver's host phases: the hardware concurrency,
// or GNX_HOST_THREADS when several processes share the box (one rank per GPU).
thread_local unsigned tl_host_thread_share = 0; // gnx_multi_*: a shard's thread takes 1 / shards of the process's threads

unsigned host_threads()
{
    static const unsigned n = [] {
        const char *e = getenv("GNX_HOST_THREADS");
        const int v = e ? atoi(e) : 0;
        return v >= 1 ? (unsigned)v : std::max(1u, std::thread::hardware_concurrency());
    }();
    return tl_host_thread_share ? std::max(1u, std::min(n, tl_host_thread_share)) : n;
}

// Pageable caller memory (a Go slice, a numpy array) cannot be DMA'd directly: it is copied through a page-locked
// stage.  One memcpy thread moves ~6 GB/s, far below PCIe 5 x16, so large copies are split over a few threads.
void par_memcpy(void *dst, const void *src, size_t bytes)
{
    constexpr size_t kMin = (size_t)4 << 20;
    const unsigned hw = host_threads();
    static const int cap = [] { // GNX_MEMCPY_THREADS overrides the default of 16 staging threads
        const char *e = getenv("GNX_MEMCPY_THREADS");
        const int v = e ? atoi(e) : 16;
        return v < 1 ? 1 : (v > 64 ? 64 : v);
    }();
    const int nt = (int)std::min<size_t>({(size_t)cap, (size_t)hw, bytes / kMin});
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (int t = 1; t < nt; ++t) {
        const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
        if (hi > lo)
            th.emplace_back([=] { memcpy((char *)dst + lo, (const char *)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, per));
    for (auto &x : th)
        x.join();
}

// dnaTwoBit.NewTwoBit of `count` sequences of `len` bases each (one byte per base, back to back) into `wlen` words per
// sequence: gnx_pack_host.cpp (scalar / AVX2).  Returns false if any base is >= 4 (such input cannot be packed: the
// caller falls back to bytes, where the kernels report the pair).
inline bool pack_range(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
    return gnx_pack_range_host(dst, src, count, len, wlen);
}

bool pack_stage(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
    const unsigned hw = host_threads();
    static const int cap = [] { // GNX_PACK_THREADS overrides the default of 16 packing threads
        const char *e = getenv("GNX_PACK_THREADS");
        const int v = e ? atoi(e) : 16;
        return v < 1 ? 1 : (v > 64 ? 64 : v);
    }();
    const int nt = (int)std::min<int64_t>({(int64_t)cap, (int64_t)hw, std::max<int64_t>(1, count * len >> 21)});
    if (nt <= 1)
        return pack_range(dst, src, count, len, wlen);
    std::vector<std::thread> th;
    std::vector<char> ok((size_t)nt, 1);
    for (int t = 1; t < nt; ++t) {
        const int64_t lo = count * t / nt, hi = count * (t + 1) / nt;
        th.emplace_back([=, &ok] { ok[(size_t)t] = pack_range(dst + lo * wlen, src + lo * len, hi - lo, len, wlen); });
    }
    ok[0] = pack_range(dst, src, count / nt, len, wlen);
    for (auto &x : th)
        x.join();
    for (char c : ok)
        if (!c)
            return false;
    return true;
}

// every offset equals p * len (p = 0 .. n_pairs): the batch is uniform and starts at byte 0
bool offsets_uniform(const int64_t *off, int64_t n_pairs, int64_t len)
{
    const int nt = n_pairs >= (1 << 20) ? (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::vector<char> ok((size_t)nt, 1);
    auto work = [&](int t) {
        const int64_t lo = (n_pairs + 1) * t / nt, hi = (n_pairs + 1) * (t + 1) / nt;
        bool good = true;
        for (int64_t p = lo; p < hi && good; ++p)
            good = off[p] == p * len;
        ok[(size_t)t] = good;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        th.emplace_back(work, t);
    work(0);
    for (auto &x : th)
        x.join();
    for (char c : ok)
        if (!c)
            return false;
    return true;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer batch: the pipelined path behind gnx_affine_batch / gnx_const_batch.
// ---------------------------------------------------------------------------------------------
// tb != nullptr: the bases arrive as dnaTwoBit words (alpha_cat / beta_cat are NULL; aoff / boff are the byte offsets
// the entry point derived from the lengths, used for planning only).  Each chunk's words are copied to the device
// (a quarter of the bytes), the packed 16-bit kernels read them directly (TMA), every other path unpacks them to
// one byte per base ON THE DEVICE first (twobit_unpack_*_kernel, HBM-bound: ~1 % of a fill).
int run_host_batch(gnx_ctx *ctx, Problem &pb, const uint8_t *alpha_cat, const int64_t *aoff, const uint8_t *beta_cat,
                   const int64_t *boff, int64_t n_pairs, int64_t *out_score, gnx_cigar *out_cigar,
                   int64_t *out_cigar_off, int64_t cigar_cap, int64_t *out_end_i = nullptr, int64_t *out_end_j = nullptr,
                   const TbIn *tb = nullptr)
{
    CU(cudaSetDevice(ctx->device));
    ctx->fill_events_used = 0;
    ctx->last_fill_launches = 0;
    ctx->last_fill_ms = 0;
    ctx->last_cells = 0;
    ctx->have_retained = false;
    ctx->retained.clear();
    if (n_pairs == 0) {
        if (out_cigar_off)
            out_cigar_off[0] = 0;
        return GNX_OK;
    }
    Plan plan;
    const int64_t budget_words = (int64_t)(ctx->workspace / kSlots / 4);
    int rc = make_plan(ctx, pb, aoff, boff, n_pairs, budget_words, plan);
    if (rc != GNX_OK)
        return rc;
    ctx->last_cells = plan.cells;
    CU(cudaMemsetAsync(ctx->status.p, 0, sizeof(int), ctx->slot[0].stream));
    CU(cudaStreamSynchronize(ctx->slot[0].stream));

    if (tb && tb->from_bytes && !pb.cfg.tb) { // the plan did not pick the kernels that read packed words: stay on bytes
        tb = nullptr;
        pb.twobit = false;
    }
    const bool pack = tb && tb->from_bytes;
    const bool pin_a = !pack && is_pinned(tb ? (const void *)tb->a_words : (const void *)alpha_cat);
    const bool pin_b = !pack && is_pinned(tb ? (const void *)tb->b_words : (const void *)beta_cat);
    const bool pin_ao = is_pinned(aoff), pin_bo = is_pinned(boff);
    const bool pin_score = is_pinned(out_score);
    const bool pin_off = out_cigar_off && is_pinned(out_cigar_off);
    const bool pin_cig = out_cigar && is_pinned(out_cigar);
    const int nwarps_total = ctx->sm_count * std::max(ctx->opt_blocks_per_sm * 4, ctx->opt_ctas_per_sm);
    const int64_t edge_stride = plan.max_n + 2;

    const int64_t n_chunks = (int64_t)plan.bounds.size() - 1;
    int64_t cig_total = 0;   // cigar elements produced by finished chunks
    bool overflow = false;   // user's cigar buffer too small -> retain in ctx

    struct Pending {
        int slot;
        int64_t begin, end;
        ChunkDev cd;
    };
    std::vector<RagTables> rag_all; // ragged batches on the packed 16-bit kernels: every chunk's quad tables
    if (pb.cfg.rag)
        build_all_rag_tables(pb, aoff, boff, plan.bounds, plan.max_n, rag_all);
    std::vector<Pending> pending;

    auto finish = [&](const Pending &pd) -> int {
        Slot &s = ctx->slot[pd.slot];
        const int64_t np = pd.end - pd.begin;
        int64_t total = 0;
        if (pb.want_cigar) {
            CU(cudaEventSynchronize(s.ev_total));
            total = *s.h_total.as<int64_t>();
            CU(s.cigars.ensure((size_t)std::max<int64_t>(total, 1) * sizeof(gnx_cigar)));
            rc = enqueue_chunk_expand(ctx, pb, pd.cd, pd.begin, pd.end, s.cig_off.as<int64_t>(),
                                      s.cigars.as<gnx_cigar>(), total, s.stream);
            if (rc != GNX_OK)
                return rc;
        }
        // scores
        if (pin_score) {
            CU(cudaMemcpyAsync(out_score + pd.begin, s.score.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s.stream));
        } else {
            CU(s.h_score.ensure((size_t)np * 8));
            CU(cudaMemcpyAsync(s.h_score.p, s.score.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s.stream));
        }
        if (pb.ext && out_end_i && out_end_j) {
            CU(s.h_endi.ensure((size_t)np * 8));
            CU(s.h_endj.ensure((size_t)np * 8));
            if (pb.want_cigar) {
                CU(cudaMemcpyAsync(s.h_endi.p, s.endi.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s.stream));
                CU(cudaMemcpyAsync(s.h_endj.p, s.endj.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s.stream));
            } else if (pb.ext == 2) {
                CU(cudaMemcpyAsync(s.h_endi.p, s.best.p, (size_t)np * 8, cudaMemcpyDeviceToHost, s.stream));
            }
        }
        if (pb.want_cigar) {
            CU(s.h_off.ensure((size_t)(np + 1) * 8));
            CU(cudaMemcpyAsync(s.h_off.p, s.cig_off.p, (size_t)(np + 1) * 8, cudaMemcpyDeviceToHost, s.stream));
            const bool fits = !overflow && out_cigar && cig_total + total <= cigar_cap;
            if (!fits && !overflow) { // first overflow: keep what the user already holds in the retained copy
                overflow = true;
                ctx->retained.resize((size_t)cig_total);
                if (out_cigar && cig_total > 0) // earlier chunks were synchronised when they finished
                    memcpy(ctx->retained.data(), out_cigar, (size_t)cig_total * sizeof(gnx_cigar));
            }
            const bool direct = fits && pin_cig;
            if (total > 0) {
                if (direct) {
                    CU(cudaMemcpyAsync(out_cigar + cig_total, s.cigars.p, (size_t)total * sizeof(gnx_cigar),
                                       cudaMemcpyDeviceToHost, s.stream));
                } else {
                    CU(s.h_cig.ensure((size_t)total * sizeof(gnx_cigar)));
                    CU(cudaMemcpyAsync(s.h_cig.p, s.cigars.p, (size_t)total * sizeof(gnx_cigar),
                                       cudaMemcpyDeviceToHost, s.stream));
                }
            }
            CU(cudaStreamSynchronize(s.stream));
            if (total > 0 && !direct) {
                if (fits) {
                    par_memcpy(out_cigar + cig_total, s.h_cig.p, (size_t)total * sizeof(gnx_cigar));
                } else {
                    ctx->retained.resize((size_t)(cig_total + total));
                    par_memcpy(ctx->retained.data() + cig_total, s.h_cig.p, (size_t)total * sizeof(gnx_cigar));
                }
            }
            const int64_t *ho = s.h_off.as<int64_t>();
            {   // chunk-relative offsets -> the caller's absolute ones (a few threads: 10^7 pairs per call add up)
                const int nt = np >= (1 << 16) ? (int)std::min(4u, host_threads()) : 1;
                const int64_t base_off = cig_total;
                int64_t *dst_off = out_cigar_off + pd.begin;
                auto work = [=](int t) {
                    for (int64_t k = np * t / nt, e = np * (t + 1) / nt; k < e; ++k)
                        dst_off[k] = base_off + ho[k];
                };
                std::vector<std::thread> th;
                for (int t = 1; t < nt; ++t)
                    th.emplace_back(work, t);
                work(0);
                for (auto &x : th)
                    x.join();
            }
            cig_total += total;
            out_cigar_off[pd.end] = cig_total;
        } else {
            CU(cudaStreamSynchronize(s.stream));
        }
        if (!pin_score)
            memcpy(out_score + pd.begin, s.h_score.p, (size_t)np * 8);
        if (pb.ext && out_end_i && out_end_j) { // the stream was synchronised above
            const int64_t *hi = s.h_endi.as<int64_t>(), *hj = s.h_endj.as<int64_t>();
            if (pb.want_cigar) {
                memcpy(out_end_i + pd.begin, hi, (size_t)np * 8);
                memcpy(out_end_j + pd.begin, hj, (size_t)np * 8);
            } else {
                for (int64_t k = 0; k < np; ++k) { // right: the packed first-maximum cell; left: unknown without a trace
                    out_end_i[pd.begin + k] = pb.ext == 2 ? (hi[k] >> 32) : -1;
                    out_end_j[pd.begin + k] = pb.ext == 2 ? (hi[k] & 0xffffffffll) : -1;
                }
            }
        }
        s.busy = false;
        return GNX_OK;
    };

    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        const int si = (int)(ci % kSlots);
        Slot &s = ctx->slot[si];
        // the slot's previous chunk must be fully retired before its buffers are reused
        if (s.busy) {
            auto it = std::find_if(pending.begin(), pending.end(), [&](const Pending &q) { return q.slot == si; });
            if (it != pending.end()) {
                Pending pd = *it;
                pending.erase(it);
                rc = finish(pd);
                if (rc != GNX_OK)
                    return rc;
            }
        }
        const int64_t begin = plan.bounds[ci], end = plan.bounds[ci + 1], np = end - begin;
        const int64_t a_lo = aoff[begin], a_hi = aoff[end], b_lo = boff[begin], b_hi = boff[end];
        CU(s.alpha.ensure((size_t)std::max<int64_t>(a_hi - a_lo, 1)));
        CU(s.beta.ensure((size_t)std::max<int64_t>(b_hi - b_lo, 1)));
        CU(s.aoff.ensure((size_t)(np + 1) * 8));
        CU(s.boff.ensure((size_t)(np + 1) * 8));
        CU(s.cls.ensure((size_t)np));
        CU(s.score.ensure((size_t)np * 8));
        // H2D (pinned user memory is DMA'd directly, pageable memory goes through a pinned stage)
        auto h2d = [&](void *dst, const void *src, size_t bytes, bool pinned, PinBuf &stage) -> int {
            if (bytes == 0)
                return GNX_OK;
            if (pinned) {
                CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.stream));
            } else {
                CU(stage.ensure(bytes));
                par_memcpy(stage.p, src, bytes);
                CU(cudaMemcpyAsync(dst, stage.p, bytes, cudaMemcpyHostToDevice, s.stream));
            }
            return GNX_OK;
        };
        const uint64_t *d_wa = nullptr, *d_wb = nullptr;
        if (!tb) {
            if ((rc = h2d(s.alpha.p, alpha_cat + a_lo, (size_t)(a_hi - a_lo), pin_a, s.h_stage_a)) != GNX_OK)
                return rc;
            if ((rc = h2d(s.beta.p, beta_cat + b_lo, (size_t)(b_hi - b_lo), pin_b, s.h_stage_b)) != GNX_OK)
                return rc;
            // offsets are small; always DMA from the caller's arrays when pinned, else pageable memcpy
            CU(cudaMemcpyAsync(s.aoff.p, aoff + begin, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            CU(cudaMemcpyAsync(s.boff.p, boff + begin, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s.stream));
        } else {
            const int64_t wa_lo = pack ? begin * tb->wn : tb->a_woff[begin], wa_n = (pack ? end * tb->wn : tb->a_woff[end]) - wa_lo;
            const int64_t wb_lo = pack ? begin * tb->wm : tb->b_woff[begin], wb_n = (pack ? end * tb->wm : tb->b_woff[end]) - wb_lo;
            CU(s.tb_a.ensure((size_t)wa_n * 8 + 256)); // slack: the TMA of a tail quad reads a whole quad's words
            CU(s.tb_b.ensure((size_t)wb_n * 8 + 256));
            if (pack) { // pageable bytes -> packed words in the page-locked stage -> device
                CU(s.h_stage_a.ensure((size_t)wa_n * 8 + 8));
                CU(s.h_stage_b.ensure((size_t)wb_n * 8 + 8));
                if (!pack_stage(s.h_stage_a.as<uint64_t>(), alpha_cat + a_lo, np, tb->n, tb->wn) ||
                    !pack_stage(s.h_stage_b.as<uint64_t>(), beta_cat + b_lo, np, tb->m, tb->wm))
                    return GNX_RETRY_BYTES;
                CU(cudaMemcpyAsync(s.tb_a.p, s.h_stage_a.p, (size_t)wa_n * 8, cudaMemcpyHostToDevice, s.stream));
                CU(cudaMemcpyAsync(s.tb_b.p, s.h_stage_b.p, (size_t)wb_n * 8, cudaMemcpyHostToDevice, s.stream));
            } else {
                if ((rc = h2d(s.tb_a.p, tb->a_words + wa_lo, (size_t)wa_n * 8, pin_a, s.h_stage_a)) != GNX_OK)
                    return rc;
                if ((rc = h2d(s.tb_b.p, tb->b_words + wb_lo, (size_t)wb_n * 8, pin_b, s.h_stage_b)) != GNX_OK)
                    return rc;
            }
            d_wa = s.tb_a.as<uint64_t>();
            d_wb = s.tb_b.as<uint64_t>();
            // the packed 16-bit kernels read the words; so do the screening and recompute kernels of the checkpoint path
            const bool need_bytes = !(pb.cfg.tb && (!pb.want_cigar || pb.cfg.impl == 17));
            if (tb->uniform) { // byte offsets made on the device, bases unpacked without any offset array
                const int g = (int)((np + 1 + 255) / 256);
                iota_offsets_kernel<<<g, 256, 0, s.stream>>>(s.aoff.as<int64_t>(), begin, np + 1, tb->n);
                iota_offsets_kernel<<<g, 256, 0, s.stream>>>(s.boff.as<int64_t>(), begin, np + 1, tb->m);
                ctx->launches += 2;
                if (need_bytes) {
                    if (wa_n > 0)
                        twobit_unpack_uniform_kernel<<<(int)((wa_n + 255) / 256), 256, 0, s.stream>>>(d_wa, wa_n, tb->wn, tb->n,
                                                                                                   s.alpha.as<uint8_t>());
                    if (wb_n > 0)
                        twobit_unpack_uniform_kernel<<<(int)((wb_n + 255) / 256), 256, 0, s.stream>>>(d_wb, wb_n, tb->wm, tb->m,
                                                                                                   s.beta.as<uint8_t>());
                    ctx->launches += 2;
                }
            } else { // ragged: byte offsets, chunk-relative word offsets and lengths travel with the chunk
                CU(cudaMemcpyAsync(s.aoff.p, aoff + begin, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s.stream));
                CU(cudaMemcpyAsync(s.boff.p, boff + begin, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s.stream));
                CU(s.tb_meta.ensure((size_t)(np + 1) * 8 * 4));
                CU(s.h_tbmeta.ensure((size_t)(np + 1) * 8 * 4)); // pinned staging (the slot's previous chunk has been retired)
                int64_t *hm = s.h_tbmeta.as<int64_t>();
                int64_t *h_woa = hm, *h_wob = hm + (np + 1), *h_la = hm + 2 * (np + 1), *h_lb = hm + 3 * (np + 1);
                for (int64_t k = 0; k <= np; ++k) {
                    h_woa[k] = tb->a_woff[begin + k] - wa_lo;
                    h_wob[k] = tb->b_woff[begin + k] - wb_lo;
                    if (k < np) {
                        h_la[k] = tb->a_len[begin + k];
                        h_lb[k] = tb->b_len[begin + k];
                    }
                }
                CU(cudaMemcpyAsync(s.tb_meta.p, hm, (size_t)(np + 1) * 8 * 4, cudaMemcpyHostToDevice, s.stream));
                const int64_t *dm = s.tb_meta.as<int64_t>();
                UnpackParams up;
                up.tb.words = d_wa;
                up.tb.word_off = dm;
                up.tb.len = dm + 2 * (np + 1);
                up.tb.n_seqs = np;
                up.out_off = s.aoff.as<int64_t>();          // absolute byte offsets ...
                up.out = s.alpha.as<uint8_t>() - a_lo;      // ... into a pointer biased by the chunk's first byte
                up.total_words = wa_n;
                up.uniform_words = 0;
                if (wa_n > 0)
                    twobit_unpack_kernel<<<(int)((wa_n + 255) / 256), 256, 0, s.stream>>>(up);
                up.tb.words = d_wb;
                up.tb.word_off = dm + (np + 1);
                up.tb.len = dm + 3 * (np + 1);
                up.out_off = s.boff.as<int64_t>();
                up.out = s.beta.as<uint8_t>() - b_lo;
                up.total_words = wb_n;
                if (wb_n > 0)
                    twobit_unpack_kernel<<<(int)((wb_n + 255) / 256), 256, 0, s.stream>>>(up);
                ctx->launches += 2;
            }
        }
        (void)pin_ao;
        (void)pin_bo;
        (void)pin_off;

        ChunkDev cd;
        memset(&cd, 0, sizeof cd);
        CU(slot_quad_ctr(s, cd));
        cd.alpha = s.alpha.as<uint8_t>() - a_lo;
        cd.beta = s.beta.as<uint8_t>() - b_lo;
        cd.aoff = s.aoff.as<int64_t>() - begin;
        cd.boff = s.boff.as<int64_t>() - begin;
        cd.cls = s.cls.as<uint8_t>() - begin;
        cd.score = s.score.as<int64_t>() - begin;
        cd.a_lo = a_lo;
        cd.a_hi = a_hi;
        cd.b_lo = b_lo;
        cd.b_hi = b_hi;
        if (tb && tb->uniform) { // TB kernels address pair p's words as base + p * wn (global pair index)
            cd.alpha_words = d_wa - begin * tb->wn;
            cd.beta_words = d_wb - begin * tb->wm;
        }
        if (pb.cfg.rag) { // the slot's previous chunk has been retired: its staging is free
            if ((rc = upload_rag_tables(ctx, s, s.h_rag, rag_all[(size_t)ci], np, cd, s.stream)) != GNX_OK)
                return rc;
        }
        if (pb.ext) {
            CU(s.best.ensure((size_t)np * 8));
            CU(s.endi.ensure((size_t)np * 8));
            CU(s.endj.ensure((size_t)np * 8));
            cd.best = s.best.as<int64_t>() - begin;
            cd.end_i = s.endi.as<int64_t>();
            cd.end_j = s.endj.as<int64_t>();
        }
        if (pb.want_cigar) {
            CU(s.h_trace_off.ensure((size_t)(np + 1) * 8));
            int64_t *to = s.h_trace_off.as<int64_t>();
            const bool need_to = pb.cfg.impl != 17; // the checkpoint path addresses its area by quad
            int64_t acc = need_to ? compute_trace_offsets(pb, aoff, boff, begin, np, to) : 0;
            if (pb.cfg.impl == 17) { // checkpoint area: whole quads; r* per pair
                acc = pb.cfg.rag ? std::max(acc, rag_all[(size_t)ci].ck_words)
                                 : std::max(acc, ((np + 3) / 4) * ckpt_quad_words(pb.cfg.n_uniform));
                CU(s.best.ensure((size_t)np * 8));
                cd.best = s.best.as<int64_t>() - begin;
                CU(s.work.ensure((size_t)np * 4 + 64));
                cd.work_count = s.work.as<int>();
                cd.work = s.work.as<int>() + 16;
            }
            CU(s.trace.ensure((size_t)std::max<int64_t>(acc, 1) * 4));
            CU(s.trace_off.ensure((size_t)(np + 1) * 8));
            if (need_to)
                CU(cudaMemcpyAsync(s.trace_off.p, to, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, s.stream));
            CU(s.slots.ensure(slots_bytes(pb, np)));
            CU(s.counts.ensure((size_t)np * 4));
            CU(s.cig_off.ensure((size_t)(np + 1) * 8));
            cd.trace = s.trace.as<uint32_t>();
            cd.trace_off = s.trace_off.as<int64_t>();
            cd.slots = s.slots.as<uint32_t>();
            cd.counts = s.counts.as<int>();
        }
        if (plan.any_long) {
            CU(s.edge.ensure((size_t)nwarps_total * 2 * edge_stride * sizeof(int2) * (pb.wide ? 2 : 1)));
            cd.edge = s.edge.as<int2>();
            cd.edge_stride = edge_stride;
        }
        rc = enqueue_chunk_compute(ctx, pb, cd, begin, end, s.stream);
        if (rc != GNX_OK)
            return rc;
        if (pb.want_cigar) {
            CU(s.misc.ensure(64));
            CU(cudaMemsetAsync(s.misc.p, 0, 8, s.stream));
            rc = enqueue_scan(ctx, s.partials, cd.counts, np, s.cig_off.as<int64_t>(), s.misc.as<int64_t>(), nullptr,
                              s.stream);
            if (rc != GNX_OK)
                return rc;
            CU(s.h_total.ensure(8));
            CU(cudaMemcpyAsync(s.h_total.p, s.cig_off.as<int64_t>() + np, 8, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaEventRecord(s.ev_total, s.stream));
        }
        s.busy = true;
        s.begin = begin;
        s.end = end;
        pending.push_back(Pending{si, begin, end, cd});
        // retire the oldest chunk once the pipeline is full so that its D2H overlaps this fill
        if ((int)pending.size() >= kSlots) {
            Pending pd = pending.front();
            pending.erase(pending.begin());
            rc = finish(pd);
            if (rc != GNX_OK)
                return rc;
        }
    }
    while (!pending.empty()) {
        Pending pd = pending.front();
        pending.erase(pending.begin());
        rc = finish(pd);
        if (rc != GNX_OK)
            return rc;
    }
    for (int k = 0; k < kSlots; ++k)
        CU(cudaStreamSynchronize(ctx->slot[k].stream));
    collect_fill_stats(ctx);
    int st = 0;
    CU(cudaMemcpy(&st, ctx->status.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (st == kEBase)
        return fail(ctx, GNX_EBASE, "a sequence holds a base >= dim (Go: index out of range in scores[a][b])");
    if (overflow) {
        ctx->have_retained = true;
        return fail(ctx, GNX_ECAP, "cigar_cap too small; call gnx_copy_last_cigars with a larger buffer");
    }
    return GNX_OK;
}

int fill_problem(gnx_ctx *ctx, Problem &pb, int kind, int want_cigar, const int64_t *scores, int dim, int64_t gap_open,
                 int64_t gap_extend)
{
    if (!scores || dim < 4 || dim > 8)
        return fail(ctx, GNX_EARG, "scores must be a dim x dim matrix with 4 <= dim <= 8");
    pb.kind = kind;
    pb.want_cigar = want_cigar ? 1 : 0;
    pb.dim = dim;
    memset(pb.scores, 0, sizeof pb.scores);
    for (int i = 0; i < dim * dim; ++i) {
        if (scores[i] > (1 << 20) || scores[i] < -(1 << 20))
            return fail(ctx, GNX_ERANGE, "score matrix entry out of range");
        pb.scores[i] = scores[i];
    }
    if (gap_open > (1 << 24) || gap_open < -(1 << 24) || gap_extend > (1 << 24) || gap_extend < -(1 << 24))
        return fail(ctx, GNX_ERANGE, "gap penalty out of range");
    pb.gap_open = gap_open;
    pb.gap_extend = gap_extend;
    pb.prmt_ok = false;
    pb.h00 = 0;
    pb.h00_plane = 0;
    return GNX_OK;
}

// ---------------------------------------------------------------------------------------------
// Profile (group-vs-group) batch: the inner loop of nearestGroups[Chunk] (align/multiAlign.go:27-57).
// Groups are uploaded and reduced to column profiles once; the pair list is cut into sub-batches whose
// cell-score matrices + traceback matrices fit the workspace; each sub-batch runs
//   profile_score_kernel -> affine_fill_kernel<LOOKUP 3> -> traceback -> scan -> expand
// on one stream (these batches are small next to the read-alignment ones; no copy/compute overlap).
// ---------------------------------------------------------------------------------------------
int run_profile_batch(gnx_ctx *ctx, Problem &pb, const uint8_t *group_cat, const int64_t *group_off,
                      const int64_t *group_nseq, int64_t n_groups, const int64_t *pair_x, const int64_t *pair_y,
                      int64_t n_pairs, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off,
                      int64_t cigar_cap)
{
    CU(cudaSetDevice(ctx->device));
    ctx->fill_events_used = 0;
    ctx->last_fill_launches = 0;
    ctx->last_fill_ms = 0;
    ctx->last_cells = 0;
    ctx->have_retained = false;
    ctx->retained.clear();
    if (out_cigar_off)
        out_cigar_off[0] = 0;
    if (n_pairs == 0)
        return GNX_OK;
    const int64_t chunk = pb.chunk;
    std::vector<int64_t> col_off((size_t)n_groups + 1, 0), len((size_t)n_groups, 0);
    for (int64_t g = 0; g < n_groups; ++g) {
        const int64_t bytes = group_off[g + 1] - group_off[g], ns = group_nseq[g];
        if (ns <= 0 || bytes < 0 || bytes % ns != 0) // Go: alpha[0] of an empty group is an index-out-of-range panic
            return fail(ctx, GNX_EARG, "every group needs >= 1 sequence and a byte count divisible by its sequence count");
        len[(size_t)g] = bytes / ns;
        col_off[(size_t)g + 1] = col_off[(size_t)g] + len[(size_t)g];
    }
    if (n_pairs >= (int64_t(1) << 23))
        return fail(ctx, GNX_EARG, "at most 2^23 - 1 group pairs per call");
    std::vector<int64_t> aoff((size_t)n_pairs + 1, 0), boff((size_t)n_pairs + 1, 0), soff((size_t)n_pairs + 1, 0),
        extra((size_t)n_pairs, 0);
    for (int64_t p = 0; p < n_pairs; ++p) {
        const int64_t x = pair_x[p], y = pair_y[p];
        if (x < 0 || x >= n_groups || y < 0 || y >= n_groups)
            return fail(ctx, GNX_EARG, "pair_x / pair_y index out of range");
        const int64_t n = len[(size_t)x], m = len[(size_t)y];
        if (n % chunk != 0 || m % chunk != 0) // affineGap_highMem.go:310-315: log.Fatalf
            return fail(ctx, GNX_ECHUNK, "a subalignment length is not a multiple of chunkSize");
        aoff[(size_t)p + 1] = aoff[(size_t)p] + n;
        boff[(size_t)p + 1] = boff[(size_t)p] + m;
        extra[(size_t)p] = (n / chunk) * (m / chunk);
        if (extra[(size_t)p] >= (int64_t(1) << 36))
            return fail(ctx, GNX_ERANGE, "a profile DP has 2^36 or more cells");
        soff[(size_t)p + 1] = soff[(size_t)p] + extra[(size_t)p];
    }
    pb.extra_words = extra.data();
    Plan plan;
    const int64_t budget_words = (int64_t)(ctx->workspace / 4);
    int rc = make_plan(ctx, pb, aoff.data(), boff.data(), n_pairs, budget_words, plan);
    pb.extra_words = nullptr;
    if (rc != GNX_OK)
        return rc;
    ctx->last_cells = soff[(size_t)n_pairs];

    Slot &s = ctx->slot[0];
    cudaStream_t st = s.stream;
    auto up = [&](DevBuf &d, const void *src, size_t bytes) -> int {
        CU(d.ensure(std::max<size_t>(bytes, 8)));
        if (bytes)
            CU(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, st));
        return GNX_OK;
    };
    const int64_t total_cols = col_off[(size_t)n_groups];
    if ((rc = up(ctx->pf_cat, group_cat, (size_t)(group_off[n_groups] - group_off[0]))) != GNX_OK ||
        (rc = up(ctx->pf_goff, group_off, (size_t)(n_groups + 1) * 8)) != GNX_OK ||
        (rc = up(ctx->pf_nseq, group_nseq, (size_t)n_groups * 8)) != GNX_OK ||
        (rc = up(ctx->pf_coloff, col_off.data(), (size_t)(n_groups + 1) * 8)) != GNX_OK ||
        (rc = up(ctx->pf_scores, pb.scores, sizeof pb.scores)) != GNX_OK ||
        (rc = up(ctx->pf_px, pair_x, (size_t)n_pairs * 8)) != GNX_OK ||
        (rc = up(ctx->pf_py, pair_y, (size_t)n_pairs * 8)) != GNX_OK ||
        (rc = up(ctx->pf_aoff, aoff.data(), (size_t)(n_pairs + 1) * 8)) != GNX_OK ||
        (rc = up(ctx->pf_boff, boff.data(), (size_t)(n_pairs + 1) * 8)) != GNX_OK ||
        (rc = up(ctx->pf_soff, soff.data(), (size_t)(n_pairs + 1) * 8)) != GNX_OK)
        return rc;
    CU(ctx->pf_prof.ensure((size_t)std::max<int64_t>(total_cols, 1) * kProfW * 4));
    CU(ctx->pf_vb.ensure((size_t)std::max<int64_t>(total_cols, 1) * 8 * 8));
    CU(ctx->pf_err.ensure(8));
    CU(cudaMemsetAsync(ctx->pf_err.p, 0xff, 8, st));
    CU(cudaMemsetAsync(ctx->status.p, 0, sizeof(int), st));
    if (total_cols > 0) {
        // group_off is absolute inside the caller's group_cat; pf_cat holds it from group_off[0] on
        profile_count_kernel<<<(int)((total_cols + 255) / 256), 256, 0, st>>>(
            ctx->pf_cat.as<uint8_t>() - group_off[0], ctx->pf_goff.as<int64_t>(), ctx->pf_nseq.as<int64_t>(),
            ctx->pf_coloff.as<int64_t>(), (int)n_groups, pb.dim, ctx->pf_scores.as<int64_t>(), ctx->pf_prof.as<int>(),
            ctx->pf_vb.as<long long>());
        ctx->launches++;
    }
    const int nwarps_total = ctx->sm_count * std::max(ctx->opt_blocks_per_sm * 4, ctx->opt_ctas_per_sm);
    const int64_t edge_stride = plan.max_n + 2;
    int64_t cig_total = 0;
    bool overflow = false;
    const int64_t n_chunks = (int64_t)plan.bounds.size() - 1;
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        const int64_t begin = plan.bounds[(size_t)ci], end = plan.bounds[(size_t)ci + 1], np = end - begin;
        const int64_t cells = soff[(size_t)end] - soff[(size_t)begin];
        CU(ctx->pf_smat.ensure((size_t)std::max<int64_t>(cells, 1) * 4));
        if (cells > 0) {
            ProfileScoreParams pp;
            pp.prof = ctx->pf_prof.as<int>();
            pp.vb = ctx->pf_vb.as<long long>();
            pp.col_off = ctx->pf_coloff.as<int64_t>();
            pp.pair_x = ctx->pf_px.as<int64_t>();
            pp.pair_y = ctx->pf_py.as<int64_t>();
            pp.smat_off = ctx->pf_soff.as<int64_t>();
            pp.pair_begin = begin;
            pp.pair_end = end;
            pp.smat_base = soff[(size_t)begin];
            pp.chunk = (int)chunk;
            pp.smat = ctx->pf_smat.as<int>();
            pp.first_error = ctx->pf_err.as<unsigned long long>();
            profile_score_kernel<<<(int)((cells + 255) / 256), 256, 0, st>>>(pp);
            ctx->launches++;
        }
        ChunkDev cd;
        memset(&cd, 0, sizeof cd);
        cd.aoff = ctx->pf_aoff.as<int64_t>();
        cd.boff = ctx->pf_boff.as<int64_t>();
        CU(s.score.ensure((size_t)np * 8));
        cd.score = s.score.as<int64_t>() - begin;
        cd.smat = ctx->pf_smat.as<int>() - soff[(size_t)begin];
        cd.smat_off = ctx->pf_soff.as<int64_t>();
        if (pb.want_cigar) {
            CU(s.h_trace_off.ensure((size_t)(np + 1) * 8));
            int64_t *to = s.h_trace_off.as<int64_t>();
            const int64_t acc = compute_trace_offsets(pb, aoff.data(), boff.data(), begin, np, to);
            CU(s.trace.ensure((size_t)std::max<int64_t>(acc, 1) * 4));
            CU(s.trace_off.ensure((size_t)(np + 1) * 8));
            CU(cudaMemcpyAsync(s.trace_off.p, to, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, st));
            CU(s.slots.ensure(slots_bytes(pb, np)));
            CU(s.counts.ensure((size_t)np * 4));
            CU(s.cig_off.ensure((size_t)(np + 1) * 8));
            cd.trace = s.trace.as<uint32_t>();
            cd.trace_off = s.trace_off.as<int64_t>();
            cd.slots = s.slots.as<uint32_t>();
            cd.counts = s.counts.as<int>();
        }
        if (plan.any_long) {
            CU(s.edge.ensure((size_t)nwarps_total * 2 * edge_stride * sizeof(int2)));
            cd.edge = s.edge.as<int2>();
            cd.edge_stride = edge_stride;
        }
        rc = enqueue_chunk_compute(ctx, pb, cd, begin, end, st);
        if (rc != GNX_OK)
            return rc;
        CU(s.h_score.ensure((size_t)np * 8));
        CU(cudaMemcpyAsync(s.h_score.p, s.score.p, (size_t)np * 8, cudaMemcpyDeviceToHost, st));
        if (pb.want_cigar) {
            CU(s.misc.ensure(64));
            CU(cudaMemsetAsync(s.misc.p, 0, 8, st));
            rc = enqueue_scan(ctx, s.partials, cd.counts, np, s.cig_off.as<int64_t>(), s.misc.as<int64_t>(), nullptr, st);
            if (rc != GNX_OK)
                return rc;
            CU(s.h_off.ensure((size_t)(np + 1) * 8));
            CU(cudaMemcpyAsync(s.h_off.p, s.cig_off.p, (size_t)(np + 1) * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st)); // also fences the pinned trace-offset staging for the next sub-batch
            const int64_t *ho = s.h_off.as<int64_t>();
            const int64_t total = ho[np];
            CU(s.cigars.ensure((size_t)std::max<int64_t>(total, 1) * sizeof(gnx_cigar)));
            rc = enqueue_chunk_expand(ctx, pb, cd, begin, end, s.cig_off.as<int64_t>(), s.cigars.as<gnx_cigar>(), total, st);
            if (rc != GNX_OK)
                return rc;
            CU(s.h_cig.ensure((size_t)std::max<int64_t>(total, 1) * sizeof(gnx_cigar)));
            CU(cudaMemcpyAsync(s.h_cig.p, s.cigars.p, (size_t)total * sizeof(gnx_cigar), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const bool fits = !overflow && out_cigar && cig_total + total <= cigar_cap;
            if (!fits && !overflow) {
                overflow = true;
                ctx->retained.resize((size_t)cig_total);
                if (out_cigar && cig_total > 0)
                    memcpy(ctx->retained.data(), out_cigar, (size_t)cig_total * sizeof(gnx_cigar));
            }
            if (fits) {
                memcpy(out_cigar + cig_total, s.h_cig.p, (size_t)total * sizeof(gnx_cigar));
            } else {
                ctx->retained.resize((size_t)(cig_total + total));
                memcpy(ctx->retained.data() + cig_total, s.h_cig.p, (size_t)total * sizeof(gnx_cigar));
            }
            for (int64_t k = 0; k < np; ++k)
                out_cigar_off[begin + k] = cig_total + ho[k];
            cig_total += total;
            out_cigar_off[end] = cig_total;
        } else {
            CU(cudaStreamSynchronize(st));
        }
        memcpy(out_score + begin, s.h_score.p, (size_t)np * 8);
    }
    collect_fill_stats(ctx);
    unsigned long long first = 0;
    CU(cudaMemcpy(&first, ctx->pf_err.p, 8, cudaMemcpyDeviceToHost));
    if (first != ~0ull) { // the panic the reference would hit first (pair order, then row-major cell order)
        if ((first & 15) == (unsigned)kEBase)
            return fail(ctx, GNX_EBASE, "a group holds a base >= dim opposite an ungapped base (Go: index out of range)");
        return fail(ctx, GNX_EDIVZERO, "a column pair has no ungapped base pair (Go: integer divide by zero in scoreColumnMatch)");
    }
    if (overflow) {
        ctx->have_retained = true;
        return fail(ctx, GNX_ECAP, "cigar_cap too small; call gnx_copy_last_cigars with a larger buffer");
    }
    return GNX_OK;
}

// 2-bit entry point shared by the host-buffer and (below) device-resident forms: derive the offset arrays
int twobit_offsets(gnx_ctx *ctx, const int64_t *alpha_len, int64_t alpha_ulen, const int64_t *beta_len, int64_t beta_ulen,
                   int64_t n_pairs, TbIn &tb)
{
    std::vector<int64_t> &ao = ctx->tb_off[0], &bo = ctx->tb_off[1], &aw = ctx->tb_off[2], &bw = ctx->tb_off[3];
    tb.a_len = alpha_len;
    tb.b_len = beta_len;
    if (!alpha_len && !beta_len) { // uniform batch: arithmetic offsets, cached between calls of the same shape
        if (alpha_ulen < 0 || beta_ulen < 0)
            return fail(ctx, GNX_EARG, "uniform lengths must be >= 0");
        tb.uniform = true;
        tb.n = alpha_ulen;
        tb.m = beta_ulen;
        tb.wn = (alpha_ulen + 31) / 32;
        tb.wm = (beta_ulen + 31) / 32;
        if (ctx->tb_key[0] != n_pairs || ctx->tb_key[1] != tb.n || ctx->tb_key[2] != tb.m) {
            for (auto &v : ctx->tb_off)
                v.resize((size_t)n_pairs + 1);
            const int nt = n_pairs >= (1 << 20) ? 8 : 1;
            auto work = [&](int t) {
                const int64_t lo = (n_pairs + 1) * t / nt, hi = (n_pairs + 1) * (t + 1) / nt;
                for (int64_t p = lo; p < hi; ++p) {
                    ao[(size_t)p] = p * tb.n;
                    bo[(size_t)p] = p * tb.m;
                    aw[(size_t)p] = p * tb.wn;
                    bw[(size_t)p] = p * tb.wm;
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t)
                th.emplace_back(work, t);
            work(0);
            for (auto &x : th)
                x.join();
            ctx->tb_key[0] = n_pairs;
            ctx->tb_key[1] = tb.n;
            ctx->tb_key[2] = tb.m;
        }
    } else {
        if (!alpha_len || !beta_len)
            return fail(ctx, GNX_EARG, "alpha_len and beta_len must both be given (or both NULL for a uniform batch)");
        ctx->tb_key[0] = -1;
        for (auto &v : ctx->tb_off)
            v.resize((size_t)n_pairs + 1);
        ao[0] = bo[0] = aw[0] = bw[0] = 0;
        for (int64_t p = 0; p < n_pairs; ++p) {
            if (alpha_len[p] < 0 || beta_len[p] < 0)
                return fail(ctx, GNX_EARG, "negative sequence length");
            ao[(size_t)p + 1] = ao[(size_t)p] + alpha_len[p];
            bo[(size_t)p + 1] = bo[(size_t)p] + beta_len[p];
            aw[(size_t)p + 1] = aw[(size_t)p] + (alpha_len[p] + 31) / 32;
            bw[(size_t)p + 1] = bw[(size_t)p] + (beta_len[p] + 31) / 32;
        }
        tb.uniform = false;
        tb.n = tb.m = tb.wn = tb.wm = 0;
    }
    tb.a_woff = aw.data();
    tb.b_woff = bw.data();
    return GNX_OK;
}

} // namespace

// =================================================================================================
// exported C ABI
// =================================================================================================
extern "C" {

int gnx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *gnx_version(void) { return "gnxalign 0.1 (sm_100a)"; }

gnx_ctx *gnx_create(int device, size_t workspace_bytes)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) {
        g_create_error = "device index out of range";
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return nullptr;
    }
    gnx_ctx *ctx = new gnx_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
        ctx->sm_count = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (workspace_bytes == 0)
        workspace_bytes = std::min<size_t>(free_b / 3 * 2, (size_t)128 << 30);
    ctx->workspace = workspace_bytes;
    for (int k = 0; k < kSlots; ++k) {
        cudaStreamCreateWithFlags(&ctx->slot[k].stream, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->slot[k].ev_total, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->slot[k].ev_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->slot[k].ev_rag[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->slot[k].ev_rag[1], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->ev_long, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_dev, cudaEventDisableTiming);
    if (ctx->status.ensure(64) != cudaSuccess || ctx->dr_misc.ensure(256) != cudaSuccess) {
        g_create_error = "cudaMalloc failed at context creation";
        delete ctx;
        return nullptr;
    }
    cudaMemset(ctx->status.p, 0, 64);
    return ctx;
}

void gnx_destroy(gnx_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < kSlots; ++k) {
        Slot &s = ctx->slot[k];
        DevBuf *d[] = {&s.alpha, &s.beta, &s.aoff, &s.boff, &s.cls, &s.trace, &s.trace_off, &s.slots,
                       &s.counts, &s.score, &s.cig_off, &s.cigars, &s.edge, &s.misc, &s.partials, &s.best, &s.endi, &s.endj, &s.work,
                       &s.tb_a, &s.tb_b, &s.tb_meta, &s.rag, &s.qctr};
        for (DevBuf *b : d)
            b->release();
        PinBuf *h[] = {&s.h_stage_a, &s.h_stage_b, &s.h_total, &s.h_trace_off, &s.h_score, &s.h_off, &s.h_cig, &s.h_endi, &s.h_endj,
                       &s.h_tbmeta, &s.h_rag, &s.h_rag2};
        for (PinBuf *b : h)
            b->release();
        if (s.stream)
            cudaStreamDestroy(s.stream);
        if (s.ev_total)
            cudaEventDestroy(s.ev_total);
        if (s.ev_done)
            cudaEventDestroy(s.ev_done);
        for (cudaEvent_t e : s.ev_rag)
            if (e)
                cudaEventDestroy(e);
    }
    for (auto &fe : ctx->fill_events) {
        cudaEventDestroy(fe.a);
        cudaEventDestroy(fe.b);
    }
    ctx->status.release();
    ctx->dr_misc.release();
    ctx->long_scratch.release();
    for (PinBuf &b : ctx->gsw_pin)
        b.release();
    if (ctx->ev_long)
        cudaEventDestroy(ctx->ev_long);
    if (ctx->ev_dev)
        cudaEventDestroy(ctx->ev_dev);
    DevBuf *pf[] = {&ctx->pf_cat, &ctx->pf_goff, &ctx->pf_nseq, &ctx->pf_coloff, &ctx->pf_scores, &ctx->pf_px, &ctx->pf_py,
                    &ctx->pf_aoff, &ctx->pf_boff, &ctx->pf_soff, &ctx->pf_prof, &ctx->pf_vb, &ctx->pf_smat, &ctx->pf_err};
    for (DevBuf *b : pf)
        b->release();
    delete ctx;
}

const char *gnx_last_error(gnx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void *gnx_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void gnx_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

int gnx_affine_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                     const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                     int64_t gap_extend, int mode, int want_cigar, int64_t *out_score, gnx_cigar *out_cigar,
                     int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score || (mode != GNX_GLOBAL && mode != GNX_FREE_END))
        return fail(ctx, GNX_EARG, "bad argument to gnx_affine_batch");
    if (want_cigar && !out_cigar_off)
        return fail(ctx, GNX_EARG, "out_cigar_off is required when want_cigar != 0");
    Problem pb;
    int rc = fill_problem(ctx, pb, mode == GNX_FREE_END ? 1 : 0, want_cigar, scores, dim, gap_open, gap_extend);
    if (rc != GNX_OK)
        return rc;
    // Large uniform batches in pageable memory: pack to 2 bits per base while staging (TbIn).  A base >= 4 (N under a
    // 5 x 5 matrix, or an invalid code) cannot be packed: the call is redone on bytes, and the next calls on this
    // context skip the attempt for a while.
    if (ctx->pack_backoff > 0)
        ctx->pack_backoff--;
    else if (ctx->opt_pack_stage && ctx->opt_tb_tma && n_pairs >= 4096 && dim >= 4 && alpha_cat && beta_cat) {
        const int64_t n = alpha_off[1] - alpha_off[0], m = beta_off[1] - beta_off[0];
        // (page-locked input is packed too: its DMA needs no staging pass, but a quarter of the PCIe bytes is worth the
        // host threads -- score-only batches are PCIe-bound otherwise)
        if (n >= 1 && n <= kTbMaxN && m >= 1 && m <= 160 && offsets_uniform(alpha_off, n_pairs, n) &&
            offsets_uniform(beta_off, n_pairs, m)) {
            Problem pb2 = pb;
            pb2.twobit = true;
            TbIn tb;
            tb.uniform = tb.from_bytes = true;
            tb.n = n;
            tb.m = m;
            tb.wn = (n + 31) / 32;
            tb.wm = (m + 31) / 32;
            rc = run_host_batch(ctx, pb2, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, out_score, out_cigar, out_cigar_off,
                                out_cigar ? cigar_cap : 0, nullptr, nullptr, &tb);
            if (rc != GNX_RETRY_BYTES)
                return rc;
            cudaDeviceSynchronize(); // chunks in flight are abandoned: the byte path redoes the whole call
            for (int k = 0; k < kSlots; ++k)
                ctx->slot[k].busy = false;
            ctx->pack_backoff = 16;
        }
    }
    return run_host_batch(ctx, pb, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, out_score, out_cigar,
                          out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_pack_twobit_host(const uint8_t *bases, int64_t count, int64_t len, uint64_t *words)
{
    if (count < 0 || len < 0 || (count > 0 && len > 0 && (!bases || !words)))
        return GNX_EARG;
    if (count == 0 || len == 0)
        return GNX_OK;
    return pack_stage(words, bases, count, len, (len + 31) / 32) ? GNX_OK : GNX_EBASE;
}

int gnx_affine_batch_twobit(gnx_ctx *ctx, const uint64_t *alpha_words, const int64_t *alpha_len, int64_t alpha_uniform_len,
                            const uint64_t *beta_words, const int64_t *beta_len, int64_t beta_uniform_len, int64_t n_pairs,
                            const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend, int mode, int want_cigar,
                            int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || !out_score || (mode != GNX_GLOBAL && mode != GNX_FREE_END) || (n_pairs > 0 && (!alpha_words || !beta_words)))
        return fail(ctx, GNX_EARG, "bad argument to gnx_affine_batch_twobit");
    if (want_cigar && !out_cigar_off)
        return fail(ctx, GNX_EARG, "out_cigar_off is required when want_cigar != 0");
    Problem pb;
    int rc = fill_problem(ctx, pb, mode == GNX_FREE_END ? 1 : 0, want_cigar, scores, dim, gap_open, gap_extend);
    if (rc != GNX_OK)
        return rc;
    pb.twobit = true;
    TbIn tb;
    tb.a_words = alpha_words;
    tb.b_words = beta_words;
    if ((rc = twobit_offsets(ctx, alpha_len, alpha_uniform_len, beta_len, beta_uniform_len, n_pairs, tb)) != GNX_OK)
        return rc;
    return run_host_batch(ctx, pb, nullptr, ctx->tb_off[0].data(), nullptr, ctx->tb_off[1].data(), n_pairs, out_score, out_cigar,
                          out_cigar_off, out_cigar ? cigar_cap : 0, nullptr, nullptr, &tb);
}

int gnx_const_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                    const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_pen,
                    int want_cigar, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off,
                    int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score)
        return fail(ctx, GNX_EARG, "bad argument to gnx_const_batch");
    if (want_cigar && !out_cigar_off)
        return fail(ctx, GNX_EARG, "out_cigar_off is required when want_cigar != 0");
    Problem pb;
    int rc = fill_problem(ctx, pb, 2, want_cigar, scores, dim, gap_pen, 0);
    if (rc != GNX_OK)
        return rc;
    return run_host_batch(ctx, pb, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, out_score, out_cigar,
                          out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_affine_chunk_batch(gnx_ctx *ctx, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                           const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                           int64_t gap_extend, int64_t chunk, int64_t *out_score, gnx_cigar *out_cigar,
                           int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score || !out_cigar_off)
        return fail(ctx, GNX_EARG, "bad argument to gnx_affine_chunk_batch");
    if (chunk <= 0 || chunk > 4096)
        return fail(ctx, GNX_ECHUNK, "chunkSize must be in 1..4096");
    for (int64_t p = 0; p < n_pairs; ++p) // align/affineGap_highMem.go:229-234: log.Fatalf on a ragged length
        if ((alpha_off[p + 1] - alpha_off[p]) % chunk != 0 || (beta_off[p + 1] - beta_off[p]) % chunk != 0)
            return fail(ctx, GNX_ECHUNK, "a sequence length is not a multiple of chunkSize");
    Problem pb;
    int rc = fill_problem(ctx, pb, 0, 1, scores, dim, gap_open, gap_extend * chunk);
    if (rc != GNX_OK)
        return rc;
    pb.chunk = chunk;
    return run_host_batch(ctx, pb, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, out_score, out_cigar,
                          out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_extend_batch(gnx_ctx *ctx, int side, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                     const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_pen,
                     int want_cigar, int64_t *out_score, int64_t *out_end_i, int64_t *out_end_j, gnx_cigar *out_cigar,
                     int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score || side < GNX_EXT_LEFT || side > GNX_EXT_RIGHT_LOCAL)
        return fail(ctx, GNX_EARG, "bad argument to gnx_extend_batch");
    const bool local = side >= GNX_EXT_LEFT_LOCAL;
    if (local)
        side -= 2;
    if (want_cigar && !out_cigar_off)
        return fail(ctx, GNX_EARG, "out_cigar_off is required when want_cigar != 0");
    if (dim > kDimP)
        return fail(ctx, GNX_EARG, "gnx_extend_batch supports score matrices up to 5 x 5");
    if (side == GNX_EXT_RIGHT && gap_pen > 0)
        return fail(ctx, GNX_EARG, "RightDynamicAln with a positive gap penalty is not supported");
    Problem pb;
    int rc = fill_problem(ctx, pb, 2, want_cigar, scores, dim, gap_pen, 0);
    if (rc != GNX_OK)
        return rc;
    pb.ext = side;
    pb.ext_local = local;
    return run_host_batch(ctx, pb, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, out_score, out_cigar,
                          out_cigar_off, out_cigar ? cigar_cap : 0, out_end_i, out_end_j);
}

int gnx_multi_affine_chunk_batch(gnx_ctx *ctx, const uint8_t *group_cat, const int64_t *group_off,
                                 const int64_t *group_nseq, int64_t n_groups, const int64_t *pair_x,
                                 const int64_t *pair_y, int64_t n_pairs, const int64_t *scores, int dim,
                                 int64_t gap_open, int64_t gap_extend, int64_t chunk, int want_cigar,
                                 int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (n_groups < 0 || n_pairs < 0 || !group_off || (n_groups > 0 && !group_nseq) || !out_score ||
        (n_pairs > 0 && (!pair_x || !pair_y)))
        return fail(ctx, GNX_EARG, "bad argument to gnx_multi_affine_chunk_batch");
    if (want_cigar && !out_cigar_off)
        return fail(ctx, GNX_EARG, "out_cigar_off is required when want_cigar != 0");
    if (chunk <= 0 || chunk > 4096)
        return fail(ctx, GNX_ECHUNK, "chunkSize must be in 1..4096");
    Problem pb;
    int rc = fill_problem(ctx, pb, 0, want_cigar, scores, dim, gap_open, gap_extend * chunk);
    if (rc != GNX_OK)
        return rc;
    pb.chunk = chunk;
    pb.profile = true;
    return run_profile_batch(ctx, pb, group_cat, group_off, group_nseq, n_groups, pair_x, pair_y, n_pairs, out_score,
                             out_cigar, out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_copy_last_cigars(gnx_ctx *ctx, gnx_cigar *out_cigar, int64_t cigar_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (!ctx->have_retained)
        return fail(ctx, GNX_EARG, "no retained cigars (the last batch call did not return GNX_ECAP)");
    if (!out_cigar || cigar_cap < (int64_t)ctx->retained.size())
        return fail(ctx, GNX_ECAP, "cigar_cap still too small");
    memcpy(out_cigar, ctx->retained.data(), ctx->retained.size() * sizeof(gnx_cigar));
    return GNX_OK;
}

// Device-resident batches: every array already in this context's device memory, work enqueued on `st`, no host
// synchronisation.  tbd != nullptr: a UNIFORM batch in dnaTwoBit form (d_alpha_cat / d_beta_cat / the offset arrays are
// NULL): byte offsets are made per chunk on the device and the chunk's bases expanded only when a kernel needs bytes.
struct TbDev {
    const uint64_t *a_words, *b_words;
    int64_t n, m, wn, wm;
};

static int run_device_batch(gnx_ctx *ctx, Problem &pb, const uint8_t *d_alpha_cat, const int64_t *d_alpha_off,
                            const uint8_t *d_beta_cat, const int64_t *d_beta_off, const int64_t *alpha_off_host,
                            const int64_t *beta_off_host, int64_t n_pairs, int64_t *d_out_score, gnx_cigar *d_out_cigar,
                            int64_t *d_out_cigar_off, int64_t cigar_cap, int32_t *d_status, cudaStream_t st, const TbDev *tbd)
{
    int rc;
    ctx->fill_events_used = 0;
    ctx->last_fill_launches = 0;
    ctx->last_fill_ms = 0;
    ctx->last_cells = 0;
    if (n_pairs == 0)
        return GNX_OK;
    Plan plan;
    const int64_t budget_words = (int64_t)(ctx->workspace / 4); // single slot: chunks run back to back
    rc = make_plan(ctx, pb, alpha_off_host, beta_off_host, n_pairs, budget_words, plan);
    if (rc != GNX_OK)
        return rc;
    ctx->last_cells = plan.cells;
    Slot &s = ctx->slot[0];
    const int nwarps_total = ctx->sm_count * std::max(ctx->opt_blocks_per_sm * 4, ctx->opt_ctas_per_sm);
    const int64_t edge_stride = plan.max_n + 2;
    if (ctx->ev_dev_set) // the previous call's kernels still own the slot scratch, status and running-total words
        CU(cudaStreamWaitEvent(st, ctx->ev_dev, 0));
    CU(cudaMemsetAsync(ctx->status.p, 0, sizeof(int), st));
    CU(cudaMemsetAsync(ctx->dr_misc.p, 0, 16, st)); // running cigar total (two ping-pong words)
    const int64_t n_chunks = (int64_t)plan.bounds.size() - 1;
    CU(s.cls.ensure((size_t)n_pairs));
    std::vector<RagTables> rag_all;
    if (pb.cfg.rag)
        build_all_rag_tables(pb, alpha_off_host, beta_off_host, plan.bounds, plan.max_n, rag_all);
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        const int64_t begin = plan.bounds[ci], end = plan.bounds[ci + 1], np = end - begin;
        ChunkDev cd;
        memset(&cd, 0, sizeof cd);
        CU(slot_quad_ctr(s, cd));
        cd.alpha = d_alpha_cat;
        cd.beta = d_beta_cat;
        cd.aoff = d_alpha_off;
        cd.boff = d_beta_off;
        cd.cls = s.cls.as<uint8_t>();
        cd.score = d_out_score;
        cd.a_lo = alpha_off_host[begin];
        cd.a_hi = alpha_off_host[end];
        cd.b_lo = beta_off_host[begin];
        cd.b_hi = beta_off_host[end];
        if (tbd) {
            CU(s.aoff.ensure((size_t)(np + 1) * 8));
            CU(s.boff.ensure((size_t)(np + 1) * 8));
            const int g = (int)((np + 1 + 255) / 256);
            iota_offsets_kernel<<<g, 256, 0, st>>>(s.aoff.as<int64_t>(), begin, np + 1, tbd->n);
            iota_offsets_kernel<<<g, 256, 0, st>>>(s.boff.as<int64_t>(), begin, np + 1, tbd->m);
            ctx->launches += 2;
            cd.aoff = s.aoff.as<int64_t>() - begin;
            cd.boff = s.boff.as<int64_t>() - begin;
            cd.alpha_words = tbd->a_words;
            cd.beta_words = tbd->b_words;
            if (!(pb.cfg.tb && (!pb.want_cigar || pb.cfg.impl == 17))) { // some kernel of this path reads bytes: expand the chunk
                CU(s.alpha.ensure((size_t)std::max<int64_t>(cd.a_hi - cd.a_lo, 1)));
                CU(s.beta.ensure((size_t)std::max<int64_t>(cd.b_hi - cd.b_lo, 1)));
                const int64_t wa = np * tbd->wn, wb = np * tbd->wm;
                if (wa > 0)
                    twobit_unpack_uniform_kernel<<<(int)((wa + 255) / 256), 256, 0, st>>>(tbd->a_words + begin * tbd->wn, wa, tbd->wn,
                                                                                        tbd->n, s.alpha.as<uint8_t>());
                if (wb > 0)
                    twobit_unpack_uniform_kernel<<<(int)((wb + 255) / 256), 256, 0, st>>>(tbd->b_words + begin * tbd->wm, wb, tbd->wm,
                                                                                        tbd->m, s.beta.as<uint8_t>());
                ctx->launches += 2;
                cd.alpha = s.alpha.as<uint8_t>() - cd.a_lo;
                cd.beta = s.beta.as<uint8_t>() - cd.b_lo;
            }
        }
        if (pb.cfg.rag) { // two staging buffers: the host stays two chunks ahead of the uploads
            const int par = (int)(ci & 1);
            if (s.ev_rag_set[par]) // the upload that last read this staging buffer (two chunks or one call ago)
                CU(cudaEventSynchronize(s.ev_rag[par]));
            if ((rc = upload_rag_tables(ctx, s, par ? s.h_rag2 : s.h_rag, rag_all[(size_t)ci], np, cd, st)) != GNX_OK)
                return rc;
            CU(cudaEventRecord(s.ev_rag[par], st));
            s.ev_rag_set[par] = true;
        }
        if (pb.want_cigar) {
            // the pinned staging is about to be rewritten by the host: fence on the upload that last read it -- the
            // previous chunk's, or the last chunk's of an earlier call that may still be queued behind other work
            if (s.ev_done_set)
                CU(cudaEventSynchronize(s.ev_done));
            CU(s.h_trace_off.ensure((size_t)(np + 1) * 8));
            int64_t *to = s.h_trace_off.as<int64_t>();
            // the checkpoint path addresses its area by quad: no per-pair trace offsets to compute or upload
            const bool need_to = pb.cfg.impl != 17;
            int64_t acc = need_to ? compute_trace_offsets(pb, alpha_off_host, beta_off_host, begin, np, to) : 0;
            if (pb.cfg.impl == 17) { // checkpoint area: whole quads; r* per pair
                acc = pb.cfg.rag ? std::max(acc, rag_all[(size_t)ci].ck_words)
                                 : std::max(acc, ((np + 3) / 4) * ckpt_quad_words(pb.cfg.n_uniform));
                CU(s.best.ensure((size_t)np * 8));
                cd.best = s.best.as<int64_t>() - begin;
                CU(s.work.ensure((size_t)np * 4 + 64));
                cd.work_count = s.work.as<int>();
                cd.work = s.work.as<int>() + 16;
            }
            CU(s.trace.ensure((size_t)std::max<int64_t>(acc, 1) * 4));
            CU(s.trace_off.ensure((size_t)(np + 1) * 8));
            if (need_to)
                CU(cudaMemcpyAsync(s.trace_off.p, to, (size_t)(np + 1) * 8, cudaMemcpyHostToDevice, st));
            CU(cudaEventRecord(s.ev_done, st));
            s.ev_done_set = true;
            CU(s.slots.ensure(slots_bytes(pb, np)));
            CU(s.counts.ensure((size_t)np * 4));
            cd.trace = s.trace.as<uint32_t>();
            cd.trace_off = s.trace_off.as<int64_t>();
            cd.slots = s.slots.as<uint32_t>();
            cd.counts = s.counts.as<int>();
        }
        if (plan.any_long) {
            CU(s.edge.ensure((size_t)nwarps_total * 2 * edge_stride * sizeof(int2) * (pb.wide ? 2 : 1)));
            cd.edge = s.edge.as<int2>();
            cd.edge_stride = edge_stride;
        }
        rc = enqueue_chunk_compute(ctx, pb, cd, begin, end, st);
        if (rc != GNX_OK)
            return rc;
        if (pb.want_cigar) {
            // the running cigar total ping-pongs between two device words from chunk to chunk
            rc = enqueue_scan(ctx, s.partials, cd.counts, np, d_out_cigar_off + begin,
                              ctx->dr_misc.as<int64_t>() + (ci & 1), ctx->dr_misc.as<int64_t>() + ((ci + 1) & 1), st);
            if (rc != GNX_OK)
                return rc;
            rc = enqueue_chunk_expand(ctx, pb, cd, begin, end, d_out_cigar_off + begin, d_out_cigar, cigar_cap, st);
            if (rc != GNX_OK)
                return rc;
        }
    }
    if (d_status)
        CU(cudaMemcpyAsync(d_status, ctx->status.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
    CU(cudaEventRecord(ctx->ev_dev, st));
    ctx->ev_dev_set = true;
    return GNX_OK;
}

int gnx_batch_device(gnx_ctx *ctx, int kind, const uint8_t *d_alpha_cat, const int64_t *d_alpha_off,
                     const uint8_t *d_beta_cat, const int64_t *d_beta_off, const int64_t *alpha_off_host,
                     const int64_t *beta_off_host, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                     int64_t gap_extend, int want_cigar, int64_t *d_out_score, gnx_cigar *d_out_cigar,
                     int64_t *d_out_cigar_off, int64_t cigar_cap, int32_t *d_status, void *cuda_stream)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || kind < 0 || kind > 2 || !d_alpha_off || !d_beta_off || !d_out_score)
        return fail(ctx, GNX_EARG, "bad argument to gnx_batch_device");
    if (want_cigar && (!d_out_cigar_off || !d_out_cigar))
        return fail(ctx, GNX_EARG, "d_out_cigar and d_out_cigar_off are required when want_cigar != 0");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Problem pb;
    int rc = fill_problem(ctx, pb, kind, want_cigar, scores, dim, gap_open, kind == 2 ? 0 : gap_extend);
    if (rc != GNX_OK)
        return rc;
    std::vector<int64_t> ha, hb;
    if (n_pairs > 0 && !alpha_off_host) {
        ha.resize((size_t)n_pairs + 1);
        CU(cudaMemcpyAsync(ha.data(), d_alpha_off, (size_t)(n_pairs + 1) * 8, cudaMemcpyDeviceToHost, st));
        alpha_off_host = ha.data();
    }
    if (n_pairs > 0 && !beta_off_host) {
        hb.resize((size_t)n_pairs + 1);
        CU(cudaMemcpyAsync(hb.data(), d_beta_off, (size_t)(n_pairs + 1) * 8, cudaMemcpyDeviceToHost, st));
        beta_off_host = hb.data();
    }
    if (!ha.empty() || !hb.empty())
        CU(cudaStreamSynchronize(st));
    return run_device_batch(ctx, pb, d_alpha_cat, d_alpha_off, d_beta_cat, d_beta_off, alpha_off_host, beta_off_host, n_pairs,
                            d_out_score, d_out_cigar, d_out_cigar_off, cigar_cap, d_status, st, nullptr);
}

int gnx_batch_device_twobit(gnx_ctx *ctx, int kind, const uint64_t *d_alpha_words, int64_t alpha_len, const uint64_t *d_beta_words,
                            int64_t beta_len, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                            int64_t gap_extend, int want_cigar, int64_t *d_out_score, gnx_cigar *d_out_cigar,
                            int64_t *d_out_cigar_off, int64_t cigar_cap, int32_t *d_status, void *cuda_stream)
{
    if (!ctx)
        return GNX_EARG;
    if (n_pairs < 0 || kind < 0 || kind > 1 || !d_out_score || (n_pairs > 0 && (!d_alpha_words || !d_beta_words)))
        return fail(ctx, GNX_EARG, "bad argument to gnx_batch_device_twobit");
    if (want_cigar && (!d_out_cigar_off || !d_out_cigar))
        return fail(ctx, GNX_EARG, "d_out_cigar and d_out_cigar_off are required when want_cigar != 0");
    if (((uintptr_t)d_alpha_words & 15) || ((uintptr_t)d_beta_words & 15))
        return fail(ctx, GNX_EARG, "device word arrays must be 16-byte aligned (cp.async.bulk source)");
    CU(cudaSetDevice(ctx->device));
    Problem pb;
    int rc = fill_problem(ctx, pb, kind, want_cigar, scores, dim, gap_open, gap_extend);
    if (rc != GNX_OK)
        return rc;
    pb.twobit = true;
    TbIn tb;
    if ((rc = twobit_offsets(ctx, nullptr, alpha_len, nullptr, beta_len, n_pairs, tb)) != GNX_OK)
        return rc;
    TbDev td;
    td.a_words = d_alpha_words;
    td.b_words = d_beta_words;
    td.n = tb.n;
    td.m = tb.m;
    td.wn = tb.wn;
    td.wm = tb.wm;
    return run_device_batch(ctx, pb, nullptr, nullptr, nullptr, nullptr, ctx->tb_off[0].data(), ctx->tb_off[1].data(), n_pairs,
                            d_out_score, d_out_cigar, d_out_cigar_off, cigar_cap, d_status, (cudaStream_t)cuda_stream, &td);
}

int64_t gnx_launch_count(gnx_ctx *ctx) { return ctx ? ctx->launches : 0; }

int gnx_last_fill_stats(gnx_ctx *ctx, double *fill_ms, int64_t *fill_launches, int64_t *cells)
{
    if (!ctx)
        return GNX_EARG;
    cudaSetDevice(ctx->device);
    collect_fill_stats(ctx);
    if (fill_ms)
        *fill_ms = ctx->last_fill_ms;
    if (fill_launches)
        *fill_launches = ctx->last_fill_launches;
    if (cells)
        *cells = ctx->last_cells;
    return GNX_OK;
}

int gnx_last_kernel_path(gnx_ctx *ctx, int *impl, int *flags)
{
    if (!ctx)
        return GNX_EARG;
    if (impl)
        *impl = ctx->last_impl;
    if (flags)
        *flags = ctx->last_flags;
    return GNX_OK;
}

int gnx_set_option(gnx_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name)
        return GNX_EARG;
    std::string k(name);
    if (k == "cols_per_lane") {
        if (value != 0 && value != 5 && value != 10)
            return fail(ctx, GNX_EARG, "cols_per_lane must be 0 (auto), 5 or 10");
        ctx->opt_cols = (int)value;
    } else if (k == "chunk_pairs") {
        if (value < 1)
            return fail(ctx, GNX_EARG, "chunk_pairs must be >= 1");
        ctx->opt_chunk_pairs = value;
    } else if (k == "blocks_per_sm") {
        if (value < 1 || value > 32)
            return fail(ctx, GNX_EARG, "blocks_per_sm must be in 1..32");
        ctx->opt_blocks_per_sm = (int)value;
    } else if (k == "fill_impl") {
        if (value != 1 && value != 3)
            return fail(ctx, GNX_EARG, "fill_impl must be 1 (first-generation kernels) or 3 (fill3/fill16)");
        ctx->opt_fill_impl = (int)value;
    } else if (k == "lanes_per_pair") {
        if (value != 0 && value != 16 && value != 32)
            return fail(ctx, GNX_EARG, "lanes_per_pair must be 0 (auto), 16 or 32");
        ctx->opt_lpp = (int)value;
    } else if (k == "tb_impl") {
        ctx->opt_tb_impl = (int)value;
    } else if (k == "fill16") {
        ctx->opt_fill16 = value ? 1 : 0;
    } else if (k == "force_lookup") {
        ctx->opt_force_lookup = (int)value;
    } else if (k == "ckpt") {
        ctx->opt_ckpt = value ? 1 : 0;
    } else if (k == "wide_cta") {
        ctx->opt_wide_cta = (int)value;
    } else if (k == "ragged16") {
        ctx->opt_rag = value ? 1 : 0;
    } else if (k == "pack_stage") {
        ctx->opt_pack_stage = value ? 1 : 0;
        ctx->pack_backoff = 0;
    } else if (k == "tb_tma") {
        ctx->opt_tb_tma = value ? 1 : 0;
    } else if (k == "long_ckpt") {
        ctx->opt_long = (int)value;
    } else if (k == "long_form") {
        ctx->opt_long_form = (int)value;
    } else if (k == "long_pool") {
        ctx->opt_long_pool = value;
    } else if (k == "ctas_per_sm") {
        if (value < 1 || value > 32)
            return fail(ctx, GNX_EARG, "ctas_per_sm must be in 1..32");
        ctx->opt_ctas_per_sm = (int)value;
    } else if (k == "workspace_bytes") {
        if (value < (1 << 20))
            return fail(ctx, GNX_EARG, "workspace_bytes must be >= 1 MiB");
        ctx->workspace = (size_t)value;
    } else {
        return fail(ctx, GNX_EARG, "unknown option");
    }
    return GNX_OK;
}

} // extern "C"

#include "gnx_twobit_api.inl"
#include "gnx_multi.inl"
#include "gnx_gsw.inl"

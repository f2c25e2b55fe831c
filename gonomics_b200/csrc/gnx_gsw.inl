// gnx_gsw.inl -- the per-read driver of cmd/gsw on top of the seed and extend kernels.
//
// Reference: genomeGraph.GraphSmithWatermanToGiraf (genomeGraph/toGiraf.go:17-72), WrapPairGiraf / setGirafFlags
// (:117-137), getGirafFlags / isProperPairAlign (:171-196), seedCouldBeBetter (genomeGraph/index.go:102-121), the
// base cases of Left/RightAlignTraversal (genomeGraph/search.go:169-181,206-216), cigar.Append / Concat /
// AppendSoftClips (cigar/tools.go:4-40) -- for a genome graph WITHOUT edges (one node per chromosome).
//
// The reference handles one read at a time: seeds, then seed after seed (longest first) a left and a right
// linear-gap DP, stopping as soon as seedCouldBeBetter says no remaining seed can beat the best score so far.  The
// DPs of a read depend on nothing but the read, the seed and the genome, so a block of reads runs in two phases:
//   1. gnx_seed_batch: the seeds of every read (GPU);
//   2. the seeds are ordered as the reference orders them and, per read, every leading seed that
//      seedCouldBeBetter admits at best score 0 -- a superset of what the sequential loop can reach, because the
//      predicate only gets stricter as the best score grows -- contributes a left and a right extension pair;
//      gnx_extend_batch aligns them all in two launches (GPU);
//   3. the reference's loop is replayed per read over the precomputed DP results (host threads): same predicate,
//      same strict ">" update, same cigar / soft-clip / path / flag assembly.
// Facts of the Go code this depends on are listed in oracle/gsw.py (by-value keepers, unreversed routes on
// edge-less nodes, empty traversal paths).
namespace {

inline bool gsw_could_be_better(int64_t seedLen, int64_t best, int64_t perfect, int64_t qlen)
{ // index.go:102-121 with the constants GraphSmithWatermanToGiraf passes (toGiraf.go:38)
    const int64_t maxMatch = 100, minMatch = 90, lsm = -196, lsmc = -296;
    const int64_t seeds = qlen / (seedLen + 1), rem = qlen % (seedLen + 1);
    if (seedLen * maxMatch >= best && perfect - ((qlen - seedLen) * minMatch) >= best)
        return true;
    if (seedLen * seeds * maxMatch + seeds * lsm >= best && perfect - rem * minMatch + seeds * lsmc >= best)
        return true;
    if (seedLen * seeds * maxMatch + rem * maxMatch + (seeds + 1) * lsm >= best && perfect + (seeds + 1) * lsmc >= best)
        return true;
    return false;
}

void gsw_heap_sort(gnx_seed *a, int64_t n)
{ // heapSortSeeds (search.go:339-373): min-heap on TotalLength => descending order, the reference's tie order
    auto heapify = [&](int64_t size, int64_t i) {
        for (;;) {
            const int64_t l = 2 * i + 1, r = 2 * i + 2;
            int64_t m = (l < size && a[l].total_length < a[i].total_length) ? l : i;
            if (r < size && a[r].total_length < a[m].total_length)
                m = r;
            if (m == i)
                return;
            std::swap(a[i], a[m]);
            i = m;
        }
    };
    if (n < 2)
        return;
    for (int64_t i = n / 2 - 1; i >= 0; --i)
        heapify(n, i);
    int64_t size = n;
    for (int64_t i = n - 1; i >= 1; --i) {
        std::swap(a[0], a[i]);
        --size;
        heapify(size, 0);
    }
}

typedef std::vector<std::pair<int64_t, uint8_t>> GswCigar; // (RunLength, Op byte)

inline void gsw_append(GswCigar &a, int64_t run, uint8_t op)
{ // cigar.Append
    if (!a.empty() && a.back().second == op)
        a.back().first += run;
    else
        a.emplace_back(run, op);
}

inline int gsw_threads(int64_t n)
{
    const unsigned hw = host_threads();
    return (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hw, (int64_t)32, n / 256 + 1}));
}

// fn(lo, hi, t) on thread t of gsw_threads(n): the split depends on n only, so two passes over the same n see the
// same ranges (run_round's counting pass leaves per-thread prefix sums that its gather pass completes)
template <typename F> void gsw_parallel(int64_t n, F &&fn)
{
    const int nt = gsw_threads(n);
    if (nt == 1) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        th.emplace_back([&, t] { fn(n * t / nt, n * (t + 1) / nt, t); });
    fn(0, n / nt, 0);
    for (auto &x : th)
        x.join();
}

} // namespace

namespace {

// page-locked scratch of the driver, kept in the context between calls (ctx->gsw_pin[...]).  The extension results
// (score, end cell, cigar offsets, cigars) stay where gnx_extend_batch wrote them: one set per (round, side).
enum { GP_SEEDS, GP_SOFF, GP_A, GP_B, GP_AO, GP_BO, GP_RES, GP_N = GP_RES + 2 * 2 * 5 };
enum { GR_SC, GR_EI, GR_EJ, GR_CO, GR_CG };
inline int gp_res(int round, int side, int what) { return GP_RES + (round * 2 + side) * 5 + what; }

struct GswSide { // one (round, side) result set
    const int64_t *score = nullptr, *end_i = nullptr, *end_j = nullptr, *coff = nullptr;
    const gnx_cigar *cig = nullptr;
};

} // namespace

extern "C" int gnx_gsw_batch(gnx_ctx *ctx, const gnx_seed_index *ix, const uint8_t *reads_cat, const int64_t *read_off, int64_t n_reads,
                             const int64_t *scores, int dim, int paired, gnx_giraf *out, gnx_cigar *out_cigar, int64_t cigar_cap,
                             int64_t *out_n_cigar)
{
    if (!ctx)
        return GNX_EARG;
    if (!ix || n_reads < 0 || !read_off || (n_reads > 0 && (!reads_cat || !out)) || !scores || dim < 5 || dim > 8 ||
        (paired && (n_reads & 1)))
        return fail(ctx, GNX_EARG, "bad argument to gnx_gsw_batch (dim must hold dna.N; paired batches hold an even number of reads)");
    if (out_n_cigar)
        *out_n_cigar = 0;
    if (n_reads == 0)
        return GNX_OK;
    const std::vector<uint8_t> &G = ix->h_genome;
    const std::vector<int64_t> &GO = ix->h_off;
    const int64_t gap_pen = -600; // the extension penalty LeftAlignTraversal / RightAlignTraversal pass (search.go:177,212)
    int rc;
    const bool timing = getenv("GNX_GSW_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = now();
    auto lap = [&](const char *what) {
        if (timing) {
            const auto t = now();
            fprintf(stderr, "[gnx_gsw_batch] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
            t_prev = t;
        }
    };
    PinBuf *pin = ctx->gsw_pin;
    CU(cudaSetDevice(ctx->device));

    // ---- phase 1: seeds of every read (GPU), straight into page-locked memory ----
    CU(pin[GP_SOFF].ensure((size_t)(n_reads + 1) * 8));
    int64_t *soff = pin[GP_SOFF].as<int64_t>();
    CU(pin[GP_SEEDS].ensure((size_t)std::max<int64_t>(4 * n_reads, 64) * sizeof(gnx_seed)));
    rc = gnx_seed_batch(ctx, ix, reads_cat, read_off, n_reads, pin[GP_SEEDS].as<gnx_seed>(), soff,
                        (int64_t)(pin[GP_SEEDS].cap / sizeof(gnx_seed)));
    if (rc == GNX_ECAP) {
        CU(pin[GP_SEEDS].ensure((size_t)soff[n_reads] * sizeof(gnx_seed)));
        rc = gnx_seed_batch(ctx, ix, reads_cat, read_off, n_reads, pin[GP_SEEDS].as<gnx_seed>(), soff,
                            (int64_t)(pin[GP_SEEDS].cap / sizeof(gnx_seed)));
    }
    if (rc != GNX_OK)
        return rc;
    gnx_seed *seeds = pin[GP_SEEDS].as<gnx_seed>();
    const int64_t n_seeds = soff[n_reads];
    lap("seeds (GPU)");

    // ---- phase 2: order every read's seeds in place, perfect scores, reverse complements ----
    const int64_t total_bases = read_off[n_reads] - read_off[0];
    std::unique_ptr<uint8_t[]> rc_cat(new uint8_t[(size_t)std::max<int64_t>(total_bases, 1)]); // fastq.FastqBig.SeqRc
    std::unique_ptr<int64_t[]> perfect(new int64_t[(size_t)n_reads]);
    static const uint8_t comp[13] = {3, 2, 1, 0, 4, 8, 7, 6, 5, 9, 10, 11, 12}; // dna/modify.go:72 complementArray
    std::atomic<bool> bad_base{false};
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            const uint8_t *rd = reads_cat + read_off[r];
            const int64_t L = read_off[r + 1] - read_off[r];
            uint8_t *rcp = rc_cat.get() + (read_off[r] - read_off[0]);
            int64_t pf = 0;
            for (int64_t i = 0; i < L; ++i) {
                const uint8_t b = rd[i];
                if (b >= dim || b > 12) {
                    bad_base = true;
                    continue;
                }
                pf += scores[b * dim + b]; // perfectMatchBig (align.go:73-79)
                rcp[L - 1 - i] = comp[b];
            }
            perfect[(size_t)r] = pf;
            gnx_seed *h = seeds + soff[r];
            const int64_t ns = soff[r + 1] - soff[r];
            if (ns > 100) // SortSeedLen: sort.Slice, order among equal lengths unspecified -- stable here
                std::stable_sort(h, h + ns, [](const gnx_seed &a, const gnx_seed &b) { return a.total_length > b.total_length; });
            else
                gsw_heap_sort(h, ns);
        }
    });
    if (bad_base)
        return fail(ctx, GNX_EBASE, "a read holds a base >= dim (Go: index out of range in scoreMatrix[b][b])");
    lap("order seeds");

    // ---- the extension DPs, in two rounds so that seeds the sequential loop can never reach are (mostly) not aligned:
    // round 1 extends every read's first (longest) seed; its score is the best score the loop holds when it looks at
    // the second seed, and the predicate only gets stricter from there, so round 2 extends seeds 1.. up to the first
    // one seedCouldBeBetter rejects at THAT score -- still a superset of what the replay below can ask for.
    std::unique_ptr<int32_t[]> ext_id(new int32_t[(size_t)std::max<int64_t>(n_seeds, 1)]); // per seed: extension id or -1
    GswSide X[2][2]; // [round][side]
    int64_t n_ext_round[2] = {0, 0};
    auto seed_score_of = [&](const gnx_seed &s, const uint8_t *cur) {
        int64_t v = 0; // scoreSeedSeq (align.go:81-87)
        for (uint32_t i = s.query_start; i < s.query_start + s.length; ++i)
            v += scores[cur[i] * dim + cur[i]];
        return v;
    };
    // per read: extension pairs and window / query bases of both sides, as prefix sums (thread-local after the counting
    // pass, completed with the thread's base in the gather pass)
    struct RoundCnt {
        int64_t n, la, lb, ra, rb;
    };
    std::unique_ptr<RoundCnt[]> pre(new RoundCnt[(size_t)n_reads + 1]);
    const int nthr = gsw_threads(n_reads);
    std::vector<RoundCnt> tbase((size_t)nthr + 1);
    // first[r] .. last[r]: the seed range (relative to the read) this round extends
    auto run_round = [&](int round, const int32_t *first, const int32_t *last) -> int {
        gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int t) {
            RoundCnt acc = {0, 0, 0, 0, 0};
            for (int64_t r = lo; r < hi; ++r) {
                pre[(size_t)r] = acc; // exclusive, relative to the thread's first read
                const int64_t L = read_off[r + 1] - read_off[r], ext = perfect[(size_t)r] / 600 + L; // sk.extension (toGiraf.go:32)
                for (int32_t k = first[(size_t)r]; k < last[(size_t)r]; ++k) {
                    const gnx_seed &s = seeds[soff[r] + k];
                    if ((int64_t)s.total_length == L)
                        continue; // the seed spans the read: no extension (toGiraf.go:45-49)
                    ++acc.n;
                    const int64_t e = ext - s.total_length, node_len = GO[s.target_id + 1] - GO[s.target_id];
                    const int64_t ref_end = s.target_start, start = (int64_t)s.target_start + s.length;
                    acc.la += std::max<int64_t>(0, std::min(ref_end, e));
                    acc.lb += s.query_start;
                    acc.ra += std::max<int64_t>(0, std::min(node_len - start, e));
                    acc.rb += L - (s.query_start + s.length);
                }
            }
            tbase[(size_t)t + 1] = acc;
        });
        tbase[0] = {0, 0, 0, 0, 0};
        for (int t = 0; t < nthr; ++t) {
            tbase[(size_t)t + 1].n += tbase[(size_t)t].n;
            tbase[(size_t)t + 1].la += tbase[(size_t)t].la;
            tbase[(size_t)t + 1].lb += tbase[(size_t)t].lb;
            tbase[(size_t)t + 1].ra += tbase[(size_t)t].ra;
            tbase[(size_t)t + 1].rb += tbase[(size_t)t].rb;
        }
        const RoundCnt tot = tbase[(size_t)nthr];
        const int64_t ne = tot.n;
        n_ext_round[round] = ne;
        lap("  count");
        if (ne == 0)
            return GNX_OK;
        const int64_t base = round ? n_ext_round[0] : 0;
        for (int side = 0; side < 2; ++side) { // 0 left (getLeftTargetBases), 1 right (getRightBases), search.go:133-145
            CU(pin[GP_A].ensure((size_t)(side ? tot.ra : tot.la) + 16));
            CU(pin[GP_B].ensure((size_t)(side ? tot.rb : tot.lb) + 16));
            CU(pin[GP_AO].ensure((size_t)(ne + 1) * 8));
            CU(pin[GP_BO].ensure((size_t)(ne + 1) * 8));
            uint8_t *A = pin[GP_A].as<uint8_t>(), *B = pin[GP_B].as<uint8_t>();
            int64_t *AO = pin[GP_AO].as<int64_t>(), *BO = pin[GP_BO].as<int64_t>();
            AO[0] = BO[0] = 0;
            gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int t) {
                const RoundCnt tb = tbase[(size_t)t];
                for (int64_t r = lo; r < hi; ++r) {
                    const int64_t L = read_off[r + 1] - read_off[r], ext = perfect[(size_t)r] / 600 + L;
                    const uint8_t *rd = reads_cat + read_off[r], *rcp = rc_cat.get() + (read_off[r] - read_off[0]);
                    const RoundCnt &pr = pre[(size_t)r];
                    int64_t x = tb.n + pr.n, ap = side ? tb.ra + pr.ra : tb.la + pr.la, bp = side ? tb.rb + pr.rb : tb.lb + pr.lb;
                    for (int32_t k = first[(size_t)r]; k < last[(size_t)r]; ++k) {
                        const gnx_seed &s = seeds[soff[r] + k];
                        if ((int64_t)s.total_length == L)
                            continue;
                        const uint8_t *cur = s.pos_strand ? rd : rcp;
                        const int64_t e = ext - s.total_length, node_len = GO[s.target_id + 1] - GO[s.target_id];
                        const uint8_t *node = G.data() + GO[s.target_id];
                        const int64_t ref_end = s.target_start, start = (int64_t)s.target_start + s.length;
                        if (side == 0) {
                            const int64_t w = std::max<int64_t>(0, std::min(ref_end, e)), q = s.query_start;
                            memcpy(A + ap, node + ref_end - w, (size_t)w);
                            memcpy(B + bp, cur, (size_t)q);
                            ap += w;
                            bp += q;
                            ext_id[(size_t)(soff[r] + k)] = (int32_t)(base + x);
                        } else {
                            const int64_t w = std::max<int64_t>(0, std::min(node_len - start, e)), q = L - (s.query_start + s.length);
                            memcpy(A + ap, node + start, (size_t)w);
                            memcpy(B + bp, cur + s.query_start + s.length, (size_t)q);
                            ap += w;
                            bp += q;
                        }
                        AO[x + 1] = ap;
                        BO[x + 1] = bp;
                        ++x;
                    }
                }
            });
            lap("  gather windows");
            PinBuf &p_sc = pin[gp_res(round, side, GR_SC)], &p_ei = pin[gp_res(round, side, GR_EI)], &p_ej = pin[gp_res(round, side, GR_EJ)],
                   &p_co = pin[gp_res(round, side, GR_CO)], &p_cg = pin[gp_res(round, side, GR_CG)];
            CU(p_sc.ensure((size_t)ne * 8));
            CU(p_ei.ensure((size_t)ne * 8));
            CU(p_ej.ensure((size_t)ne * 8));
            CU(p_co.ensure((size_t)(ne + 1) * 8));
            CU(p_cg.ensure((size_t)std::max<int64_t>(4 * ne, 64) * sizeof(gnx_cigar)));
            int64_t *co = p_co.as<int64_t>();
            int e = gnx_extend_batch(ctx, side ? GNX_EXT_RIGHT : GNX_EXT_LEFT, A, AO, B, BO, ne, scores, dim, gap_pen, 1,
                                     p_sc.as<int64_t>(), p_ei.as<int64_t>(), p_ej.as<int64_t>(), p_cg.as<gnx_cigar>(), co,
                                     (int64_t)(p_cg.cap / sizeof(gnx_cigar)));
            if (e == GNX_ECAP) {
                CU(p_cg.ensure((size_t)co[ne] * sizeof(gnx_cigar)));
                e = gnx_copy_last_cigars(ctx, p_cg.as<gnx_cigar>(), co[ne]);
            }
            if (e != GNX_OK)
                return e;
            X[round][side].score = p_sc.as<int64_t>();
            X[round][side].end_i = p_ei.as<int64_t>();
            X[round][side].end_j = p_ej.as<int64_t>();
            X[round][side].coff = co;
            X[round][side].cig = p_cg.as<gnx_cigar>();
            lap("  gnx_extend_batch");
        }
        return GNX_OK;
    };
    std::unique_ptr<int32_t[]> first(new int32_t[(size_t)n_reads]), last(new int32_t[(size_t)n_reads]);
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            first[(size_t)r] = 0;
            last[(size_t)r] = soff[r + 1] > soff[r] ? 1 : 0; // pred(seed 0, best 0) holds for every seed
        }
    });
    memset(ext_id.get(), 0xff, (size_t)std::max<int64_t>(n_seeds, 1) * sizeof(int32_t)); // -1
    if ((rc = run_round(0, first.get(), last.get())) != GNX_OK)
        return rc;
    lap("round 1 (first seeds)");
    const int64_t n_ext1 = n_ext_round[0];
    // extension x of either round: its result set and index there
    auto ext_of = [&](int32_t x, int side, int64_t &local) -> const GswSide & {
        const int round = x >= n_ext1 ? 1 : 0;
        local = x - (round ? n_ext1 : 0);
        return X[round][side];
    };
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            const int64_t ns = soff[r + 1] - soff[r], L = read_off[r + 1] - read_off[r];
            first[(size_t)r] = last[(size_t)r] = 1;
            if (ns < 2)
                continue;
            const gnx_seed &s0 = seeds[soff[r]];
            const uint8_t *cur = s0.pos_strand ? reads_cat + read_off[r] : rc_cat.get() + (read_off[r] - read_off[0]);
            int64_t sc0 = seed_score_of(s0, cur);
            const int32_t id = ext_id[(size_t)soff[r]];
            if (id >= 0)
                sc0 += X[0][0].score[(size_t)id] + X[0][1].score[(size_t)id];
            const int64_t best1 = std::max<int64_t>(sc0, 0); // currBest.AlnScore after the first seed
            int32_t k = 1;
            while (k < ns && gsw_could_be_better(seeds[soff[r] + k].total_length, best1, perfect[(size_t)r], L))
                ++k;
            last[(size_t)r] = k;
        }
    });
    if ((rc = run_round(1, first.get(), last.get())) != GNX_OK)
        return rc;
    lap("round 2 (remaining seeds)");
    const int64_t n_ext_total = n_ext_round[0] + n_ext_round[1];
    if (timing)
        fprintf(stderr, "[gnx_gsw_batch] reads %lld seeds %lld extension pairs %lld + %lld\n", (long long)n_reads, (long long)n_seeds,
                (long long)n_ext1, (long long)(n_ext_total - n_ext1));

    // ---- phase 3: replay of the reference's per-read loop over the precomputed DPs ----
    // every thread appends the cigars of its (contiguous) reads to its own arena; out[r].cigar_off is arena-relative
    // until the arenas are laid end to end in the caller's buffer below
    std::vector<std::vector<gnx_cigar>> arena((size_t)nthr);
    std::atomic<bool> missing{false};
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int t) {
        GswCigar left, right, cig, best_cig;
        std::vector<gnx_cigar> &mine = arena[(size_t)t];
        mine.reserve((size_t)(hi - lo) * 3);
        for (int64_t r = lo; r < hi; ++r) {
            bool has_cig = false;
            best_cig.clear();
            const int64_t L = read_off[r + 1] - read_off[r];
            const uint8_t *rd = reads_cat + read_off[r], *rcp = rc_cat.get() + (read_off[r] - read_off[0]);
            const int64_t pf = perfect[(size_t)r];
            gnx_giraf g;
            memset(&g, 0, sizeof g);
            g.pos_strand = 1;
            g.node = -1;
            left.clear();   // sk.leftAlignment / rightAlignment / queryEnd live across the seeds of ONE read
            right.clear();
            int64_t query_end = 0;
            const int64_t ns = soff[r + 1] - soff[r];
            for (int64_t k = 0; k < ns; ++k) {
                const gnx_seed &s = seeds[soff[r] + k];
                if (!gsw_could_be_better(s.total_length, g.aln_score, pf, L))
                    break;
                const uint8_t *cur = s.pos_strand ? rd : rcp;
                const int64_t seed_score = seed_score_of(s, cur);
                int64_t tstart, tend, qstart, score;
                if ((int64_t)s.total_length == L) {
                    tstart = s.target_start;
                    tend = (int64_t)s.target_start + s.length;
                    qstart = s.query_start;
                    score = seed_score;
                } else {
                    const int32_t x = ext_id[(size_t)(soff[r] + k)];
                    if (x < 0) { // cannot happen: the rounds extend a superset of what this loop reaches
                        missing = true;
                        break;
                    }
                    const int64_t e = pf / 600 + L - s.total_length;
                    const int64_t ref_end = s.target_start, start = (int64_t)s.target_start + s.length;
                    const int64_t wl = std::max<int64_t>(0, std::min(ref_end, e));
                    int64_t xl = 0;
                    const GswSide &Lx = ext_of(x, 0, xl), &Rx = ext_of(x, 1, xl);
                    left.clear();
                    for (int64_t c = Lx.coff[xl]; c < Lx.coff[xl + 1]; ++c)
                        left.emplace_back(Lx.cig[c].run_length, Lx.cig[c].op);
                    right.clear();
                    for (int64_t c = Rx.coff[xl]; c < Rx.coff[xl + 1]; ++c)
                        right.emplace_back(Rx.cig[c].run_length, Rx.cig[c].op);
                    tstart = ref_end - wl + Lx.end_i[xl]; // refEnd - len(s.Seq) - len(seq) + targetStart (search.go:178)
                    qstart = Lx.end_j[xl];
                    tend = Rx.end_i[xl] + start;          // targetEnd + start (:214)
                    query_end = Rx.end_j[xl];
                    score = Lx.score[xl] + seed_score + Rx.score[xl];
                }
                if (score > g.aln_score) { // toGiraf.go:56-64
                    g.q_start = (int32_t)qstart;
                    g.q_end = (int32_t)((int64_t)s.query_start + qstart + query_end + s.total_length - 1);
                    g.pos_strand = s.pos_strand ? 1 : 0;
                    g.t_start = (int32_t)tstart;
                    g.t_end = (int32_t)tend;
                    g.node = (int32_t)s.target_id;
                    g.aln_score = score;
                    // cigar.Concat(cigar.Append(left, {TotalLength, 'M'}), right)
                    cig = left;
                    gsw_append(cig, s.total_length, 'M');
                    if (!right.empty()) {
                        gsw_append(cig, right[0].first, right[0].second);
                        cig.insert(cig.end(), right.begin() + 1, right.end());
                    }
                    // cigar.AppendSoftClips(queryStart, len(currSeq), cig), its drop-the-body quirk included
                    int64_t qlen = 0;
                    for (const auto &c : cig)
                        if (c.second == 'M' || c.second == 'I' || c.second == 'S' || c.second == '=' || c.second == 'X')
                            qlen += c.first;
                    GswCigar &dst = best_cig;
                    if (qstart == 0 && qlen >= L) {
                        dst = cig;
                    } else {
                        dst.clear();
                        if (qstart > 0)
                            dst.emplace_back(qstart, 'S');
                        if (qstart + qlen < L) {
                            dst.insert(dst.end(), cig.begin(), cig.end());
                            dst.emplace_back(L - qstart - qlen, 'S');
                        }
                    }
                    has_cig = true;
                }
            }
            g.flag = (g.pos_strand ? 4 : 0) + (g.aln_score < 1200 ? 2 : 0); // getGirafFlags (:187-196)
            g.cigar_off = (int64_t)mine.size();
            g.n_cigar = has_cig ? (int32_t)best_cig.size() : -1; // -1: Cigar == nil (no seed beat score 0)
            for (const auto &c : best_cig) {
                gnx_cigar o = gnx_cigar();
                o.run_length = c.first;
                o.op = c.second;
                mine.push_back(o);
            }
            out[r] = g;
        }
    });
    if (missing)
        return fail(ctx, GNX_ECUDA, "gnx_gsw_batch: internal error (a replayed seed has no extension result)");
    lap("replay");
    if (paired) { // setGirafFlags (toGiraf.go:126-137): as written there, the forward mate gets 8 + 16 + 16
        gsw_parallel(n_reads / 2, [&](int64_t lo, int64_t hi, int) {
            for (int64_t q = lo; q < hi; ++q) {
                gnx_giraf &f = out[2 * q], &v = out[2 * q + 1];
                f.flag += 8 + 16 + 16;
                const int64_t d = (int64_t)f.t_start - v.t_start;
                bool proper = false;
                if ((d < 0 ? -d : d) < 10000) {
                    if (f.t_start < v.t_start && f.pos_strand && !v.pos_strand)
                        proper = true;
                    if (f.t_start > v.t_start && !f.pos_strand && v.pos_strand)
                        proper = true;
                }
                if (proper) {
                    f.flag += 1;
                    v.flag += 1;
                }
                f.flag &= 0xff; // Flag is a uint8
                v.flag &= 0xff;
            }
        });
    }
    std::vector<int64_t> abase((size_t)nthr + 1, 0);
    for (int t = 0; t < nthr; ++t)
        abase[(size_t)t + 1] = abase[(size_t)t] + (int64_t)arena[(size_t)t].size();
    const int64_t total = abase[(size_t)nthr];
    if (out_n_cigar)
        *out_n_cigar = total;
    const bool fits = total <= cigar_cap && (total == 0 || out_cigar);
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int t) { // same split as the replay: thread t owns arena t
        for (int64_t r = lo; r < hi; ++r)
            out[r].cigar_off += abase[(size_t)t];
        if (fits && !arena[(size_t)t].empty())
            memcpy(out_cigar + abase[(size_t)t], arena[(size_t)t].data(), arena[(size_t)t].size() * sizeof(gnx_cigar));
    });
    if (!fits)
        return fail(ctx, GNX_ECAP, "cigar_cap too small for the block's cigars (out_n_cigar holds the size needed)");
    lap("pack results");
    return GNX_OK;
}

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

struct GswSeed {
    uint32_t tid, tstart, qstart, len, pos, total;
};

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

void gsw_heap_sort(std::vector<GswSeed> &a)
{ // heapSortSeeds (search.go:339-373): min-heap on TotalLength => descending order, the reference's tie order
    auto heapify = [&](size_t size, size_t i) {
        for (;;) {
            const size_t l = 2 * i + 1, r = 2 * i + 2;
            size_t m = (l < size && a[l].total < a[i].total) ? l : i;
            if (r < size && a[r].total < a[m].total)
                m = r;
            if (m == i)
                return;
            std::swap(a[i], a[m]);
            i = m;
        }
    };
    if (a.size() < 2)
        return;
    for (size_t i = a.size() / 2; i-- > 0;)
        heapify(a.size(), i);
    size_t size = a.size();
    for (size_t i = a.size() - 1; i >= 1; --i) {
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

template <typename F> void gsw_parallel(int64_t n, F &&fn)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)hw, (int64_t)32, n / 256 + 1}));
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

    // ---- phase 1: seeds of every read ----
    std::vector<gnx_seed> seeds((size_t)std::max<int64_t>(8 * n_reads, 64));
    std::vector<int64_t> soff((size_t)n_reads + 1);
    rc = gnx_seed_batch(ctx, ix, reads_cat, read_off, n_reads, seeds.data(), soff.data(), (int64_t)seeds.size());
    if (rc == GNX_ECAP) {
        seeds.resize((size_t)soff[(size_t)n_reads]);
        rc = gnx_seed_batch(ctx, ix, reads_cat, read_off, n_reads, seeds.data(), soff.data(), (int64_t)seeds.size());
    }
    if (rc != GNX_OK)
        return rc;

    // ---- phase 2: order the seeds, count the extension pairs each read can need ----
    const int64_t total_bases = read_off[n_reads] - read_off[0];
    std::vector<uint8_t> rc_cat((size_t)total_bases); // dna.ReverseComplement of every read (fastq.FastqBig.SeqRc)
    std::vector<std::vector<GswSeed>> hits((size_t)n_reads);
    std::vector<int64_t> perfect((size_t)n_reads), n_cand((size_t)n_reads), ext_first((size_t)n_reads + 1, 0);
    std::vector<int64_t> la_len((size_t)n_reads + 1, 0), lb_len((size_t)n_reads + 1, 0), ra_len((size_t)n_reads + 1, 0),
        rb_len((size_t)n_reads + 1, 0);
    static const uint8_t comp[13] = {3, 2, 1, 0, 4, 8, 7, 6, 5, 9, 10, 11, 12}; // dna/modify.go:72 complementArray
    std::atomic<bool> bad_base{false};
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            const uint8_t *rd = reads_cat + read_off[r];
            const int64_t L = read_off[r + 1] - read_off[r];
            uint8_t *rcp = rc_cat.data() + (read_off[r] - read_off[0]);
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
            std::vector<GswSeed> &h = hits[(size_t)r];
            h.resize((size_t)(soff[(size_t)r + 1] - soff[(size_t)r]));
            for (size_t k = 0; k < h.size(); ++k) {
                const gnx_seed &s = seeds[(size_t)soff[(size_t)r] + k];
                h[k] = GswSeed{s.target_id, s.target_start, s.query_start, s.length, s.pos_strand, s.total_length};
            }
            if (h.size() > 100) // SortSeedLen: sort.Slice, order among equal lengths unspecified -- stable here
                std::stable_sort(h.begin(), h.end(), [](const GswSeed &a, const GswSeed &b) { return a.total > b.total; });
            else
                gsw_heap_sort(h);
            const int64_t ext = pf / 600 + L; // sk.extension (toGiraf.go:32)
            int64_t nc = 0, ne = 0, la = 0, lb = 0, ra = 0, rb = 0;
            for (const GswSeed &s : h) {
                if (!gsw_could_be_better(s.total, 0, pf, L))
                    break;
                ++nc;
                if ((int64_t)s.total == L)
                    continue; // the seed spans the read: no extension (toGiraf.go:45-49)
                ++ne;
                const int64_t e = ext - s.total, node_len = GO[s.tid + 1] - GO[s.tid];
                const int64_t ref_end = s.tstart, start = (int64_t)s.tstart + s.len;
                la += std::max<int64_t>(0, std::min(ref_end, e));
                lb += s.qstart;
                ra += std::max<int64_t>(0, std::min(node_len - start, e));
                rb += L - (s.qstart + s.len);
            }
            n_cand[(size_t)r] = nc;
            ext_first[(size_t)r + 1] = ne;
            la_len[(size_t)r + 1] = la;
            lb_len[(size_t)r + 1] = lb;
            ra_len[(size_t)r + 1] = ra;
            rb_len[(size_t)r + 1] = rb;
        }
    });
    if (bad_base)
        return fail(ctx, GNX_EBASE, "a read holds a base >= dim (Go: index out of range in scoreMatrix[b][b])");
    for (int64_t r = 0; r < n_reads; ++r) {
        ext_first[(size_t)r + 1] += ext_first[(size_t)r];
        la_len[(size_t)r + 1] += la_len[(size_t)r];
        lb_len[(size_t)r + 1] += lb_len[(size_t)r];
        ra_len[(size_t)r + 1] += ra_len[(size_t)r];
        rb_len[(size_t)r + 1] += rb_len[(size_t)r];
    }
    const int64_t n_ext = ext_first[(size_t)n_reads];

    // ---- the extension pairs: target windows (getLeftTargetBases / getRightBases, search.go:133-145) and read flanks ----
    std::vector<uint8_t> la_cat((size_t)la_len[(size_t)n_reads]), lb_cat((size_t)lb_len[(size_t)n_reads]),
        ra_cat((size_t)ra_len[(size_t)n_reads]), rb_cat((size_t)rb_len[(size_t)n_reads]);
    std::vector<int64_t> la_off((size_t)n_ext + 1, 0), lb_off((size_t)n_ext + 1, 0), ra_off((size_t)n_ext + 1, 0), rb_off((size_t)n_ext + 1, 0);
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            const int64_t L = read_off[r + 1] - read_off[r];
            const uint8_t *rd = reads_cat + read_off[r], *rcp = rc_cat.data() + (read_off[r] - read_off[0]);
            const int64_t ext = perfect[(size_t)r] / 600 + L;
            int64_t x = ext_first[(size_t)r], la = la_len[(size_t)r], lb = lb_len[(size_t)r], ra = ra_len[(size_t)r], rb = rb_len[(size_t)r];
            const std::vector<GswSeed> &h = hits[(size_t)r];
            for (int64_t k = 0; k < n_cand[(size_t)r]; ++k) {
                const GswSeed &s = h[(size_t)k];
                if ((int64_t)s.total == L)
                    continue;
                const uint8_t *cur = s.pos ? rd : rcp;
                const int64_t e = ext - s.total, node_len = GO[s.tid + 1] - GO[s.tid];
                const uint8_t *node = G.data() + GO[s.tid];
                const int64_t ref_end = s.tstart, start = (int64_t)s.tstart + s.len;
                const int64_t wl = std::max<int64_t>(0, std::min(ref_end, e)), wr = std::max<int64_t>(0, std::min(node_len - start, e));
                const int64_t ql = s.qstart, qr = L - (s.qstart + s.len);
                memcpy(la_cat.data() + la, node + ref_end - wl, (size_t)wl);
                memcpy(lb_cat.data() + lb, cur, (size_t)ql);
                memcpy(ra_cat.data() + ra, node + start, (size_t)wr);
                memcpy(rb_cat.data() + rb, cur + s.qstart + s.len, (size_t)qr);
                la += wl;
                lb += ql;
                ra += wr;
                rb += qr;
                la_off[(size_t)x + 1] = la;
                lb_off[(size_t)x + 1] = lb;
                ra_off[(size_t)x + 1] = ra;
                rb_off[(size_t)x + 1] = rb;
                ++x;
            }
        }
    });

    // ---- phase 2b: every left and right DP in two batched calls ----
    std::vector<int64_t> l_score((size_t)n_ext), l_i((size_t)n_ext), l_j((size_t)n_ext), l_coff((size_t)n_ext + 1, 0);
    std::vector<int64_t> r_score((size_t)n_ext), r_i((size_t)n_ext), r_j((size_t)n_ext), r_coff((size_t)n_ext + 1, 0);
    std::vector<gnx_cigar> l_cig, r_cig;
    auto extend = [&](int side, std::vector<uint8_t> &a, std::vector<int64_t> &ao, std::vector<uint8_t> &b, std::vector<int64_t> &bo,
                      std::vector<int64_t> &sc, std::vector<int64_t> &ei, std::vector<int64_t> &ej, std::vector<int64_t> &co,
                      std::vector<gnx_cigar> &cg) -> int {
        if (n_ext == 0)
            return GNX_OK;
        cg.resize((size_t)std::max<int64_t>(6 * n_ext, 64));
        uint8_t dummy = 0;
        int e = gnx_extend_batch(ctx, side, a.empty() ? &dummy : a.data(), ao.data(), b.empty() ? &dummy : b.data(), bo.data(), n_ext,
                                 scores, dim, gap_pen, 1, sc.data(), ei.data(), ej.data(), cg.data(), co.data(), (int64_t)cg.size());
        if (e == GNX_ECAP) {
            cg.resize((size_t)co[(size_t)n_ext]);
            e = gnx_copy_last_cigars(ctx, cg.data(), (int64_t)cg.size());
        }
        return e;
    };
    if ((rc = extend(GNX_EXT_LEFT, la_cat, la_off, lb_cat, lb_off, l_score, l_i, l_j, l_coff, l_cig)) != GNX_OK)
        return rc;
    if ((rc = extend(GNX_EXT_RIGHT, ra_cat, ra_off, rb_cat, rb_off, r_score, r_i, r_j, r_coff, r_cig)) != GNX_OK)
        return rc;

    // ---- phase 3: replay of the reference's per-read loop over the precomputed DPs ----
    std::vector<GswCigar> cig_out((size_t)n_reads);
    std::vector<uint8_t> has_cig((size_t)n_reads, 0);
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        GswCigar left, right, cig;
        for (int64_t r = lo; r < hi; ++r) {
            const int64_t L = read_off[r + 1] - read_off[r];
            const uint8_t *rd = reads_cat + read_off[r], *rcp = rc_cat.data() + (read_off[r] - read_off[0]);
            const int64_t pf = perfect[(size_t)r];
            gnx_giraf g;
            memset(&g, 0, sizeof g);
            g.pos_strand = 1;
            g.node = -1;
            left.clear();   // sk.leftAlignment / rightAlignment / queryEnd live across the seeds of ONE read
            right.clear();
            int64_t query_end = 0, x = ext_first[(size_t)r];
            const std::vector<GswSeed> &h = hits[(size_t)r];
            for (size_t k = 0; k < h.size(); ++k) {
                const GswSeed &s = h[k];
                if (!gsw_could_be_better(s.total, g.aln_score, pf, L))
                    break; // k < n_cand always holds here: the predicate at best >= 0 implies the predicate at 0
                const uint8_t *cur = s.pos ? rd : rcp;
                int64_t seed_score = 0; // scoreSeedSeq (align.go:81-87)
                for (uint32_t i = s.qstart; i < s.qstart + s.len; ++i)
                    seed_score += scores[cur[i] * dim + cur[i]];
                int64_t tstart, tend, qstart, score;
                if ((int64_t)s.total == L) {
                    tstart = s.tstart;
                    tend = (int64_t)s.tstart + s.len;
                    qstart = s.qstart;
                    score = seed_score;
                } else {
                    const int64_t e = pf / 600 + L - s.total, node_len = GO[s.tid + 1] - GO[s.tid];
                    const int64_t ref_end = s.tstart, start = (int64_t)s.tstart + s.len;
                    const int64_t wl = std::max<int64_t>(0, std::min(ref_end, e));
                    (void)node_len;
                    left.clear();
                    for (int64_t c = l_coff[(size_t)x]; c < l_coff[(size_t)x + 1]; ++c)
                        left.emplace_back(l_cig[(size_t)c].run_length, l_cig[(size_t)c].op);
                    right.clear();
                    for (int64_t c = r_coff[(size_t)x]; c < r_coff[(size_t)x + 1]; ++c)
                        right.emplace_back(r_cig[(size_t)c].run_length, r_cig[(size_t)c].op);
                    tstart = ref_end - wl + l_i[(size_t)x]; // refEnd - len(s.Seq) - len(seq) + targetStart (search.go:178)
                    qstart = l_j[(size_t)x];
                    tend = r_i[(size_t)x] + start;          // targetEnd + start (:214)
                    query_end = r_j[(size_t)x];
                    score = l_score[(size_t)x] + seed_score + r_score[(size_t)x];
                    ++x;
                }
                if (score > g.aln_score) { // toGiraf.go:56-64
                    g.q_start = (int32_t)qstart;
                    g.q_end = (int32_t)((int64_t)s.qstart + qstart + query_end + s.total - 1);
                    g.pos_strand = s.pos ? 1 : 0;
                    g.t_start = (int32_t)tstart;
                    g.t_end = (int32_t)tend;
                    g.node = (int32_t)s.tid;
                    g.aln_score = score;
                    // cigar.Concat(cigar.Append(left, {TotalLength, 'M'}), right)
                    cig = left;
                    gsw_append(cig, s.total, 'M');
                    if (!right.empty()) {
                        gsw_append(cig, right[0].first, right[0].second);
                        cig.insert(cig.end(), right.begin() + 1, right.end());
                    }
                    // cigar.AppendSoftClips(queryStart, len(currSeq), cig), its drop-the-body quirk included
                    int64_t qlen = 0;
                    for (const auto &c : cig)
                        if (c.second == 'M' || c.second == 'I' || c.second == 'S' || c.second == '=' || c.second == 'X')
                            qlen += c.first;
                    GswCigar &dst = cig_out[(size_t)r];
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
                    has_cig[(size_t)r] = 1;
                }
            }
            g.flag = (g.pos_strand ? 4 : 0) + (g.aln_score < 1200 ? 2 : 0); // getGirafFlags (:187-196)
            out[r] = g;
        }
    });
    if (paired) { // setGirafFlags (toGiraf.go:126-137): as written there, the forward mate gets 8 + 16 + 16
        for (int64_t p = 0; p + 1 < n_reads; p += 2) {
            gnx_giraf &f = out[p], &v = out[p + 1];
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
    }
    int64_t total = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        out[r].cigar_off = total;
        out[r].n_cigar = has_cig[(size_t)r] ? (int32_t)cig_out[(size_t)r].size() : -1; // -1: Cigar == nil (no seed beat score 0)
        total += (int64_t)cig_out[(size_t)r].size();
    }
    if (out_n_cigar)
        *out_n_cigar = total;
    if (total > cigar_cap || (total > 0 && !out_cigar))
        return fail(ctx, GNX_ECAP, "cigar_cap too small for the block's cigars (out_n_cigar holds the size needed)");
    gsw_parallel(n_reads, [&](int64_t lo, int64_t hi, int) {
        for (int64_t r = lo; r < hi; ++r) {
            gnx_cigar *dst = out_cigar + out[r].cigar_off;
            for (const auto &c : cig_out[(size_t)r]) {
                dst->run_length = c.first;
                dst->op = c.second;
                ++dst;
            }
        }
    });
    return GNX_OK;
}

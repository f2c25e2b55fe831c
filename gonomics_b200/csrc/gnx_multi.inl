// gnx_multi.inl -- several GPUs behind ONE C call (include/gnxalign.h "multi-GPU"): SURVEY.md 8e.
//
// Pairs are independent, so a batch is cut into contiguous, cell-balanced shards (the rule of
// gonomics_b200/shard.py:shard_bounds), one per device; each shard runs the ordinary single-device pipeline
// (gnx_affine_batch / gnx_const_batch) on its own context from its own host thread, scores land directly in the
// caller's array and the shards' cigars are stitched into the caller's buffer in pair order.  There is no
// data-path collective: the only "exchange" is the host-side concatenation of results.  A Go caller binds this
// through cgo exactly like the single-device entry points (integration/go/align/align_cuda.go).
//
// Written against the public ABI only (no access to gnx_ctx internals).

struct gnx_multi {
    std::vector<gnx_ctx *> ctx;
    std::vector<int> device;
    std::string err;
    // per-shard cigar staging (page-locked, grown on demand) and the cigars kept after GNX_ECAP
    std::vector<PinBuf> stage;
    std::vector<gnx_cigar> retained;
    bool have_retained = false;
    // stats of the last call
    std::vector<int64_t> last_bounds;
};

namespace {

// Contiguous pair ranges balanced by DP cells: shard r starts at the first pair whose cell prefix reaches
// total * r / world (shard.py:shard_bounds).
void multi_bounds(const int64_t *aoff, const int64_t *boff, int64_t n_pairs, int world, std::vector<int64_t> &cuts)
{
    cuts.assign((size_t)world + 1, 0);
    __int128 total = 0;
    for (int64_t p = 0; p < n_pairs; ++p)
        total += (__int128)(aoff[p + 1] - aoff[p]) * (boff[p + 1] - boff[p]);
    __int128 acc = 0;
    int r = 1;
    for (int64_t p = 0; p <= n_pairs && r < world; ++p) { // acc = cells of pairs [0, p)
        while (r < world && acc >= total * r / world)
            cuts[(size_t)r++] = p;
        if (p < n_pairs)
            acc += (__int128)(aoff[p + 1] - aoff[p]) * (boff[p + 1] - boff[p]);
    }
    for (; r < world; ++r)
        cuts[(size_t)r] = n_pairs;
    cuts[(size_t)world] = n_pairs;
    for (int k = 1; k <= world; ++k)
        cuts[(size_t)k] = std::min(std::max(cuts[(size_t)k], cuts[(size_t)k - 1]), n_pairs);
}

// kind 0 / 1: affine global / free end; 2: const gap (gap_open = penalty)
int multi_run(gnx_multi *mg, int kind, const uint8_t *alpha_cat, const int64_t *aoff, const uint8_t *beta_cat,
              const int64_t *boff, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open, int64_t gap_extend,
              int want_cigar, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap)
{
    const int world = (int)mg->ctx.size();
    mg->have_retained = false;
    mg->retained.clear();
    multi_bounds(aoff, boff, n_pairs, world, mg->last_bounds);
    const std::vector<int64_t> &cut = mg->last_bounds;
    std::vector<int> rc((size_t)world, GNX_OK);
    std::vector<int64_t> total((size_t)world, 0);
    std::vector<std::vector<int64_t>> soff((size_t)world), loff((size_t)world); // rebased offsets, shard cigar offsets
    const unsigned all_threads = host_threads();
    auto shard = [&](int r) {
        const int64_t lo = cut[(size_t)r], hi = cut[(size_t)r + 1], np = hi - lo;
        if (np == 0)
            return;
        // the shards stage and pack concurrently: each takes its share of the process's host threads
        tl_host_thread_share = std::max(2u, all_threads / (unsigned)world);
        // offsets rebased to the shard (the single-device entry points take absolute offsets into the arrays they
        // are given; rebasing lets the bases be passed as a sub-range without a copy)
        std::vector<int64_t> &so = soff[(size_t)r];
        so.resize((size_t)(np + 1) * 2);
        int64_t *ao = so.data(), *bo = so.data() + np + 1;
        for (int64_t k = 0; k <= np; ++k) {
            ao[k] = aoff[lo + k] - aoff[lo];
            bo[k] = boff[lo + k] - boff[lo];
        }
        gnx_ctx *c = mg->ctx[(size_t)r];
        std::vector<int64_t> &lo_off = loff[(size_t)r];
        gnx_cigar *cg = nullptr;
        int64_t cap = 0;
        if (want_cigar) {
            lo_off.resize((size_t)np + 1);
            // first guess: the caller's capacity split by pairs (+ slack); GNX_ECAP grows the stage and refetches
            cap = std::max<int64_t>(1024, (cigar_cap / std::max<int64_t>(n_pairs, 1) + 1) * np + 1024);
            if (mg->stage[(size_t)r].ensure((size_t)cap * sizeof(gnx_cigar)) != cudaSuccess) {
                rc[(size_t)r] = GNX_ECUDA;
                return;
            }
            cap = (int64_t)(mg->stage[(size_t)r].cap / sizeof(gnx_cigar));
            cg = mg->stage[(size_t)r].as<gnx_cigar>();
        }
        int e;
        if (kind == 2)
            e = gnx_const_batch(c, alpha_cat + aoff[lo], ao, beta_cat + boff[lo], bo, np, scores, dim, gap_open, want_cigar,
                                out_score + lo, cg, want_cigar ? lo_off.data() : nullptr, cap);
        else
            e = gnx_affine_batch(c, alpha_cat + aoff[lo], ao, beta_cat + boff[lo], bo, np, scores, dim, gap_open, gap_extend,
                                 kind == 1 ? GNX_FREE_END : GNX_GLOBAL, want_cigar, out_score + lo, cg,
                                 want_cigar ? lo_off.data() : nullptr, cap);
        if (e == GNX_ECAP && want_cigar) { // the shard produced more than its stage: grow and fetch the retained copy
            const int64_t need = lo_off[(size_t)np];
            // the pipeline may hold the old stage's pages: gnx_copy_last_cigars copies from the context's own retained vector
            if (mg->stage[(size_t)r].ensure((size_t)need * sizeof(gnx_cigar)) != cudaSuccess) {
                rc[(size_t)r] = GNX_ECUDA;
                return;
            }
            cg = mg->stage[(size_t)r].as<gnx_cigar>();
            e = gnx_copy_last_cigars(c, cg, need);
        }
        rc[(size_t)r] = e;
        if (want_cigar && e == GNX_OK)
            total[(size_t)r] = lo_off[(size_t)np];
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < world; ++r)
            th.emplace_back(shard, r);
        shard(0);
        tl_host_thread_share = 0; // shard 0 ran on the caller's thread: give it its full share back
        for (auto &t : th)
            t.join();
    }
    for (int r = 0; r < world; ++r)
        if (rc[(size_t)r] != GNX_OK) { // the reference would fail on the first offending pair: lowest shard wins
            mg->err = std::string("device ") + std::to_string(mg->device[(size_t)r]) + ": " + gnx_last_error(mg->ctx[(size_t)r]);
            return rc[(size_t)r];
        }
    if (!want_cigar)
        return GNX_OK;
    // stitch: shard r's cigars start at the sum of the totals before it
    std::vector<int64_t> base((size_t)world + 1, 0);
    for (int r = 0; r < world; ++r)
        base[(size_t)r + 1] = base[(size_t)r] + total[(size_t)r];
    const int64_t grand = base[(size_t)world];
    const bool fits = out_cigar && grand <= cigar_cap;
    gnx_cigar *dst = out_cigar;
    if (!fits) {
        mg->retained.resize((size_t)grand);
        dst = mg->retained.data();
    }
    auto stitch = [&](int r) {
        const int64_t lo = cut[(size_t)r], np = cut[(size_t)r + 1] - lo;
        if (np == 0)
            return;
        if (total[(size_t)r] > 0)
            memcpy(dst + base[(size_t)r], mg->stage[(size_t)r].p, (size_t)total[(size_t)r] * sizeof(gnx_cigar));
        const int64_t *lf = loff[(size_t)r].data();
        for (int64_t k = 0; k < np; ++k)
            out_cigar_off[lo + k] = base[(size_t)r] + lf[k];
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < world; ++r)
            th.emplace_back(stitch, r);
        stitch(0);
        for (auto &t : th)
            t.join();
    }
    out_cigar_off[n_pairs] = grand;
    if (!fits) {
        mg->have_retained = true;
        mg->err = "cigar_cap too small; call gnx_multi_copy_last_cigars with a larger buffer";
        return GNX_ECAP;
    }
    return GNX_OK;
}

} // namespace

extern "C" {

gnx_multi *gnx_multi_create(const int *devices, int n_devices, size_t workspace_bytes_per_device)
{
    const int avail = gnx_device_count();
    if (avail <= 0) {
        g_create_error = "no CUDA device";
        return nullptr;
    }
    if (n_devices <= 0) { // every visible device
        n_devices = avail;
        devices = nullptr;
    }
    gnx_multi *mg = new gnx_multi();
    for (int k = 0; k < n_devices; ++k) {
        const int d = devices ? devices[k] : k;
        gnx_ctx *c = gnx_create(d, workspace_bytes_per_device);
        if (!c) { // g_create_error holds the reason
            for (gnx_ctx *x : mg->ctx)
                gnx_destroy(x);
            delete mg;
            return nullptr;
        }
        mg->ctx.push_back(c);
        mg->device.push_back(d);
    }
    mg->stage.resize((size_t)n_devices);
    return mg;
}

void gnx_multi_destroy(gnx_multi *mg)
{
    if (!mg)
        return;
    for (gnx_ctx *c : mg->ctx)
        gnx_destroy(c);
    for (PinBuf &b : mg->stage)
        b.release();
    delete mg;
}

int gnx_multi_device_count(const gnx_multi *mg) { return mg ? (int)mg->ctx.size() : 0; }

const char *gnx_multi_last_error(gnx_multi *mg) { return mg ? mg->err.c_str() : g_create_error.c_str(); }

gnx_ctx *gnx_multi_context(gnx_multi *mg, int index)
{
    return (mg && index >= 0 && index < (int)mg->ctx.size()) ? mg->ctx[(size_t)index] : nullptr;
}

int gnx_multi_shard_bounds(const gnx_multi *mg, const int64_t *alpha_off, const int64_t *beta_off, int64_t n_pairs,
                           int64_t *out_bounds)
{
    if (!mg || !alpha_off || !beta_off || n_pairs < 0 || !out_bounds)
        return GNX_EARG;
    std::vector<int64_t> cuts;
    multi_bounds(alpha_off, beta_off, n_pairs, (int)mg->ctx.size(), cuts);
    memcpy(out_bounds, cuts.data(), cuts.size() * sizeof(int64_t));
    return GNX_OK;
}

int gnx_multi_affine_batch(gnx_multi *mg, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                           const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_open,
                           int64_t gap_extend, int mode, int want_cigar, int64_t *out_score, gnx_cigar *out_cigar,
                           int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!mg)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score || (mode != GNX_GLOBAL && mode != GNX_FREE_END) ||
        (want_cigar && !out_cigar_off)) {
        mg->err = "bad argument to gnx_multi_affine_batch";
        return GNX_EARG;
    }
    if (n_pairs == 0) {
        if (out_cigar_off)
            out_cigar_off[0] = 0;
        return GNX_OK;
    }
    return multi_run(mg, mode == GNX_FREE_END ? 1 : 0, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, scores, dim, gap_open,
                     gap_extend, want_cigar, out_score, out_cigar, out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_multi_const_batch(gnx_multi *mg, const uint8_t *alpha_cat, const int64_t *alpha_off, const uint8_t *beta_cat,
                          const int64_t *beta_off, int64_t n_pairs, const int64_t *scores, int dim, int64_t gap_pen,
                          int want_cigar, int64_t *out_score, gnx_cigar *out_cigar, int64_t *out_cigar_off, int64_t cigar_cap)
{
    if (!mg)
        return GNX_EARG;
    if (n_pairs < 0 || !alpha_off || !beta_off || !out_score || (want_cigar && !out_cigar_off)) {
        mg->err = "bad argument to gnx_multi_const_batch";
        return GNX_EARG;
    }
    if (n_pairs == 0) {
        if (out_cigar_off)
            out_cigar_off[0] = 0;
        return GNX_OK;
    }
    return multi_run(mg, 2, alpha_cat, alpha_off, beta_cat, beta_off, n_pairs, scores, dim, gap_pen, 0, want_cigar, out_score,
                     out_cigar, out_cigar_off, out_cigar ? cigar_cap : 0);
}

int gnx_multi_copy_last_cigars(gnx_multi *mg, gnx_cigar *out_cigar, int64_t cigar_cap)
{
    if (!mg)
        return GNX_EARG;
    if (!mg->have_retained) {
        mg->err = "no retained cigars (the last call did not return GNX_ECAP)";
        return GNX_EARG;
    }
    if (!out_cigar || cigar_cap < (int64_t)mg->retained.size()) {
        mg->err = "cigar_cap still too small";
        return GNX_ECAP;
    }
    memcpy(out_cigar, mg->retained.data(), mg->retained.size() * sizeof(gnx_cigar));
    return GNX_OK;
}

} // extern "C"

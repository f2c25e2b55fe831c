// gnx_twobit_api.inl -- host side of the dna/dnaTwoBit + seed entry points (included at the end of gnx_api.cu;
// kernels in gnx_twobit.cuh).  Everything here runs on the context's first stream; the 2-bit sets and seed
// indexes are device-resident handles so that a genome is uploaded and packed once.

struct gnx_twobit {
    gnx_ctx *ctx = nullptr;
    DevBuf words, word_off, len;
    std::vector<int64_t> h_word_off, h_len;
    int64_t n_seqs = 0, total_words = 0, uniform_words = 0;
    gnx::TwoBitView view() const
    {
        return gnx::TwoBitView{words.as<uint64_t>(), word_off.as<int64_t>(), len.as<int64_t>(), n_seqs};
    }
};

struct gnx_seed_index {
    gnx_ctx *ctx = nullptr;
    gnx_twobit *genome = nullptr;
    DevBuf key, loc, bucket;
    int64_t n = 0;
    int seed_len = 0, seed_step = 0, bucket_bits = 0, bucket_shift = 0;
    // host copy of the nodes' bases: gnx_gsw_batch cuts its extension windows out of it (inputs are never retained)
    std::vector<uint8_t> h_genome;
    std::vector<int64_t> h_off;
    gnx::SeedIndexView view() const
    {
        return gnx::SeedIndexView{key.as<uint64_t>(), loc.as<uint64_t>(), n, bucket.as<int64_t>(), bucket_bits, bucket_shift};
    }
};

namespace {

int tb_launch_check(gnx_ctx *ctx, int n = 1)
{
    ctx->launches += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string("kernel launch failed: ") + cudaGetErrorString(e);
        return GNX_ECUDA;
    }
    return GNX_OK;
}

// status word protocol of the 2-bit kernels: status[0] = code, first_bad (int64) = offending query / read
struct TbStatus {
    int *code;
    int64_t *first_bad;
};

int tb_status_reset(gnx_ctx *ctx, cudaStream_t st, TbStatus &s)
{
    CU(ctx->dr_misc.ensure(256));
    s.code = ctx->dr_misc.as<int>() + 32;
    s.first_bad = ctx->dr_misc.as<int64_t>() + 20;
    const int zero = 0;
    const int64_t big = INT64_MAX;
    CU(cudaMemcpyAsync(s.code, &zero, 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.first_bad, &big, 8, cudaMemcpyHostToDevice, st));
    return GNX_OK;
}

int tb_status_read(gnx_ctx *ctx, cudaStream_t st, const TbStatus &s, const char *what)
{
    int64_t packed = 0;
    CU(cudaMemcpyAsync(&packed, s.first_bad, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (packed == INT64_MAX)
        return GNX_OK;
    const int code = (int)(packed & 0xff);
    const int64_t bad = packed >> 8;
    char buf[256];
    const char *why = code == GNX_EOFFSET  ? "start offsets differ modulo 32 (Go: log.Fatalf \"Different offsets\")"
                      : code == GNX_EINDEX ? "position beyond the sequence's words (Go: index out of range)"
                      : code == GNX_EBASE  ? "base > 12 (Go: complementArray index out of range)"
                                           : "device-side error";
    snprintf(buf, sizeof buf, "%s: %s at element %lld", what, why, (long long)bad);
    ctx->err = buf;
    return code;
}

// Enqueue NewTwoBit of every sequence: d_seq / d_seq_off / d_word_off are device arrays.
int tb_enqueue_pack(gnx_ctx *ctx, const uint8_t *d_seq, int64_t total_bytes, const int64_t *d_seq_off, const int64_t *d_word_off,
                    int64_t n_seqs, int64_t total_words, int64_t uniform_words, int lead, uint64_t *d_words, cudaStream_t st)
{
    if (total_words == 0)
        return GNX_OK;
    gnx::PackParams P;
    P.seq = d_seq;
    P.seq_off = d_seq_off;
    P.word_off = d_word_off;
    P.words = d_words;
    P.n_seqs = n_seqs;
    P.total_words = total_words;
    P.total_bytes = total_bytes;
    P.uniform_words = uniform_words;
    P.lead = lead;
    const int64_t blocks = (total_words + 255) / 256;
    gnx::twobit_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(P);
    return tb_launch_check(ctx);
}

int tb_build(gnx_ctx *ctx, gnx_twobit *tb, const uint8_t *seq_cat, const int64_t *seq_off, int64_t n_seqs, int lead,
             DevBuf *keep_bytes, DevBuf *keep_off)
{
    cudaStream_t st = ctx->slot[0].stream;
    tb->ctx = ctx;
    tb->n_seqs = n_seqs;
    tb->h_word_off.assign((size_t)n_seqs + 1, 0);
    tb->h_len.assign((size_t)n_seqs, 0);
    bool uniform = n_seqs > 0;
    for (int64_t s = 0; s < n_seqs; ++s) {
        const int64_t L = seq_off[s + 1] - seq_off[s];
        if (L < 0)
            return fail(ctx, GNX_EARG, "seq_off must be non-decreasing");
        tb->h_len[s] = L + lead; // TwoBit.Len of rainbow element `lead` (rainbow.go:16)
        const int64_t w = (L + lead + 31) / 32;
        tb->h_word_off[s + 1] = tb->h_word_off[s] + w;
        uniform &= w == tb->h_word_off[1];
    }
    tb->total_words = tb->h_word_off[n_seqs];
    tb->uniform_words = (uniform && tb->total_words > 0) ? tb->h_word_off[1] : 0;
    const int64_t base = n_seqs ? seq_off[0] : 0, total_bytes = n_seqs ? seq_off[n_seqs] - base : 0;
    DevBuf tmp_bytes, tmp_off;
    DevBuf &d_bytes = keep_bytes ? *keep_bytes : tmp_bytes;
    DevBuf &d_off = keep_off ? *keep_off : tmp_off;
    int rc = GNX_OK;
    do {
        cudaError_t e;
        if ((e = d_bytes.ensure((size_t)total_bytes + 64)) != cudaSuccess || (e = d_off.ensure(((size_t)n_seqs + 1) * 8)) != cudaSuccess ||
            (e = tb->words.ensure((size_t)tb->total_words * 8 + 8)) != cudaSuccess ||
            (e = tb->word_off.ensure(((size_t)n_seqs + 1) * 8)) != cudaSuccess ||
            (e = tb->len.ensure((size_t)n_seqs * 8 + 8)) != cudaSuccess) {
            ctx->err = std::string("cudaMalloc failed (2-bit set): ") + cudaGetErrorString(e);
            rc = GNX_ECUDA;
            break;
        }
        std::vector<int64_t> rel((size_t)n_seqs + 1);
        for (int64_t s = 0; s <= n_seqs; ++s)
            rel[s] = seq_off[s] - base;
        if (total_bytes)
            cudaMemcpyAsync(d_bytes.p, seq_cat + base, (size_t)total_bytes, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_off.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(tb->word_off.p, tb->h_word_off.data(), tb->h_word_off.size() * 8, cudaMemcpyHostToDevice, st);
        if (n_seqs)
            cudaMemcpyAsync(tb->len.p, tb->h_len.data(), (size_t)n_seqs * 8, cudaMemcpyHostToDevice, st);
        rc = tb_enqueue_pack(ctx, d_bytes.as<uint8_t>(), total_bytes, d_off.as<int64_t>(), tb->word_off.as<int64_t>(), n_seqs,
                             tb->total_words, tb->uniform_words, lead, tb->words.as<uint64_t>(), st);
        cudaError_t es = cudaStreamSynchronize(st); // rel[] and the caller's buffers are released on return
        if (rc == GNX_OK && es != cudaSuccess) {
            ctx->err = std::string("2-bit packing failed: ") + cudaGetErrorString(es);
            rc = GNX_ECUDA;
        }
    } while (0);
    tmp_bytes.release();
    tmp_off.release();
    return rc;
}

} // namespace

extern "C" {

int gnx_twobit_new(gnx_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int64_t n_seqs, int lead, gnx_twobit **out)
{
    if (!ctx)
        return GNX_EARG;
    if (!out || n_seqs < 0 || !seq_off || lead < 0 || lead > 31 || (!seq_cat && n_seqs > 0 && seq_off[n_seqs] > seq_off[0]))
        return fail(ctx, GNX_EARG, "gnx_twobit_new: bad argument (lead must be 0..31)");
    CU(cudaSetDevice(ctx->device));
    gnx_twobit *tb = new gnx_twobit();
    const int rc = tb_build(ctx, tb, seq_cat, seq_off, n_seqs, lead, nullptr, nullptr);
    if (rc != GNX_OK) {
        gnx_twobit_free(tb);
        return rc;
    }
    *out = tb;
    return GNX_OK;
}

void gnx_twobit_free(gnx_twobit *tb)
{
    if (!tb)
        return;
    if (tb->ctx)
        cudaSetDevice(tb->ctx->device);
    tb->words.release();
    tb->word_off.release();
    tb->len.release();
    delete tb;
}

int gnx_twobit_info(const gnx_twobit *tb, int64_t *n_seqs, int64_t *total_words)
{
    if (!tb)
        return GNX_EARG;
    if (n_seqs)
        *n_seqs = tb->n_seqs;
    if (total_words)
        *total_words = tb->total_words;
    return GNX_OK;
}

int gnx_twobit_download(gnx_ctx *ctx, const gnx_twobit *tb, uint64_t *out_words, int64_t *out_word_off, int64_t *out_len)
{
    if (!ctx)
        return GNX_EARG;
    if (!tb || tb->ctx != ctx)
        return fail(ctx, GNX_EARG, "gnx_twobit_download: the set belongs to another context");
    CU(cudaSetDevice(ctx->device));
    if (out_words && tb->total_words)
        CU(cudaMemcpy(out_words, tb->words.p, (size_t)tb->total_words * 8, cudaMemcpyDeviceToHost));
    if (out_word_off)
        memcpy(out_word_off, tb->h_word_off.data(), tb->h_word_off.size() * 8);
    if (out_len && tb->n_seqs)
        memcpy(out_len, tb->h_len.data(), (size_t)tb->n_seqs * 8);
    return GNX_OK;
}

int gnx_twobit_unpack(gnx_ctx *ctx, const gnx_twobit *tb, uint8_t *out_cat, int64_t out_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (!tb || tb->ctx != ctx || (!out_cat && out_cap > 0))
        return fail(ctx, GNX_EARG, "gnx_twobit_unpack: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    std::vector<int64_t> off((size_t)tb->n_seqs + 1, 0);
    for (int64_t s = 0; s < tb->n_seqs; ++s)
        off[s + 1] = off[s] + tb->h_len[s];
    const int64_t total = off[tb->n_seqs];
    if (total > out_cap)
        return fail(ctx, GNX_ECAP, "gnx_twobit_unpack: output buffer smaller than the sum of the lengths");
    if (total == 0)
        return GNX_OK;
    Slot &sl = ctx->slot[0];
    CU(sl.alpha.ensure((size_t)total + 64));
    CU(sl.aoff.ensure(off.size() * 8));
    CU(cudaMemcpyAsync(sl.aoff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
    gnx::UnpackParams P;
    P.tb = tb->view();
    P.out_off = sl.aoff.as<int64_t>();
    P.out = sl.alpha.as<uint8_t>();
    P.total_words = tb->total_words;
    P.uniform_words = tb->uniform_words;
    gnx::twobit_unpack_kernel<<<(unsigned)((tb->total_words + 255) / 256), 256, 0, st>>>(P);
    int rc = tb_launch_check(ctx);
    if (rc != GNX_OK)
        return rc;
    CU(cudaMemcpyAsync(out_cat, sl.alpha.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GNX_OK;
}

int gnx_twobit_get_bases(gnx_ctx *ctx, const gnx_twobit *tb, const int64_t *q_seq, const int64_t *q_pos, int64_t n_q, uint8_t *out)
{
    if (!ctx)
        return GNX_EARG;
    if (!tb || tb->ctx != ctx || n_q < 0 || (n_q > 0 && (!q_seq || !q_pos || !out)))
        return fail(ctx, GNX_EARG, "gnx_twobit_get_bases: bad argument");
    if (n_q == 0)
        return GNX_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    Slot &sl = ctx->slot[0];
    CU(sl.aoff.ensure((size_t)n_q * 8));
    CU(sl.boff.ensure((size_t)n_q * 8));
    CU(sl.cls.ensure((size_t)n_q));
    CU(cudaMemcpyAsync(sl.aoff.p, q_seq, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.boff.p, q_pos, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    TbStatus ts;
    int rc = tb_status_reset(ctx, st, ts);
    if (rc != GNX_OK)
        return rc;
    gnx::twobit_get_bases_kernel<<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(tb->view(), sl.aoff.as<int64_t>(), sl.boff.as<int64_t>(), n_q,
                                                                                  sl.cls.as<uint8_t>(), ts.code, ts.first_bad);
    if ((rc = tb_launch_check(ctx)) != GNX_OK)
        return rc;
    CU(cudaMemcpyAsync(out, sl.cls.p, (size_t)n_q, cudaMemcpyDeviceToHost, st));
    return tb_status_read(ctx, st, ts, "GetBase");
}

int gnx_twobit_count_matches(gnx_ctx *ctx, int dir, const gnx_twobit *one, const gnx_twobit *two, const int64_t *q_one,
                             const int64_t *q_start_one, const int64_t *q_two, const int64_t *q_start_two, int64_t n_q,
                             int64_t *out_matches)
{
    if (!ctx)
        return GNX_EARG;
    if (!one || !two || one->ctx != ctx || two->ctx != ctx || (dir != GNX_MATCH_RIGHT && dir != GNX_MATCH_LEFT) || n_q < 0 ||
        (n_q > 0 && (!q_one || !q_start_one || !q_two || !q_start_two || !out_matches)))
        return fail(ctx, GNX_EARG, "gnx_twobit_count_matches: bad argument");
    if (n_q == 0)
        return GNX_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    Slot &sl = ctx->slot[0];
    CU(sl.misc.ensure((size_t)n_q * 8 * 5));
    int64_t *d = sl.misc.as<int64_t>();
    CU(cudaMemcpyAsync(d, q_one, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + n_q, q_start_one, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + 2 * n_q, q_two, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + 3 * n_q, q_start_two, (size_t)n_q * 8, cudaMemcpyHostToDevice, st));
    TbStatus ts;
    int rc = tb_status_reset(ctx, st, ts);
    if (rc != GNX_OK)
        return rc;
    gnx::twobit_count_kernel<<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(dir == GNX_MATCH_LEFT, one->view(), two->view(), d, d + n_q, d + 2 * n_q,
                                                                              d + 3 * n_q, n_q, d + 4 * n_q, ts.code, ts.first_bad);
    if ((rc = tb_launch_check(ctx)) != GNX_OK)
        return rc;
    CU(cudaMemcpyAsync(out_matches, d + 4 * n_q, (size_t)n_q * 8, cudaMemcpyDeviceToHost, st));
    return tb_status_read(ctx, st, ts, dir == GNX_MATCH_LEFT ? "CountLeftMatches" : "CountRightMatches");
}

int gnx_twobit_pack_device(gnx_ctx *ctx, const uint8_t *d_seq, int64_t n_bases, int lead, uint64_t *d_words, void *cuda_stream)
{
    if (!ctx)
        return GNX_EARG;
    if (n_bases < 0 || lead < 0 || lead > 31 || (n_bases > 0 && (!d_seq || !d_words)))
        return fail(ctx, GNX_EARG, "gnx_twobit_pack_device: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // single sequence: offsets {0, n_bases} and {0, words} live in the context's scratch
    CU(ctx->dr_misc.ensure(256));
    int64_t *d_off = ctx->dr_misc.as<int64_t>() + 24; // [24..25] seq_off, [26..27] word_off
    const int64_t words = (n_bases + lead + 31) / 32;
    const int64_t h[4] = {0, n_bases, 0, words};
    CU(cudaMemcpyAsync(d_off, h, sizeof h, cudaMemcpyHostToDevice, st));
    return tb_enqueue_pack(ctx, d_seq, n_bases, d_off, d_off + 2, 1, words, 0, lead, d_words, st);
}

// ---- seed index ----------------------------------------------------------------------------------
int gnx_seed_index_new(gnx_ctx *ctx, const uint8_t *genome_cat, const int64_t *node_off, int64_t n_nodes, int seed_len,
                       int seed_step, gnx_seed_index **out)
{
    if (!ctx)
        return GNX_EARG;
    if (!out || !node_off || n_nodes < 0 || (!genome_cat && n_nodes > 0))
        return fail(ctx, GNX_EARG, "gnx_seed_index_new: bad argument");
    if (seed_len < 2 || seed_len > 32) // index.go:22-24 log.Fatalf
        return fail(ctx, GNX_EARG, "seed length needs to be greater than 1 and less than 33");
    if (seed_step < 1)
        return fail(ctx, GNX_EARG, "seed step must be >= 1");
    if (n_nodes >= ((int64_t)1 << 31))
        return fail(ctx, GNX_ERANGE, "too many nodes");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    gnx_seed_index *ix = new gnx_seed_index();
    ix->ctx = ctx;
    ix->seed_len = seed_len;
    ix->seed_step = seed_step;
    ix->genome = new gnx_twobit();
    ix->h_off.assign(node_off, node_off + n_nodes + 1);
    for (auto &v : ix->h_off)
        v -= node_off[0];
    if (n_nodes > 0)
        ix->h_genome.assign(genome_cat + node_off[0], genome_cat + node_off[n_nodes]);
    DevBuf d_bytes, d_off, d_cand, d_key, d_loc, d_valid, d_dst, d_tmp, d_key2, d_loc2;
    int rc = GNX_OK;
    do {
        for (int64_t s = 0; s < n_nodes && rc == GNX_OK; ++s)
            if (node_off[s + 1] - node_off[s] >= ((int64_t)1 << 32))
                rc = fail(ctx, GNX_ERANGE, "node longer than 2^32 bases (ChromAndPosToNumber packs the position in 32 bits)");
        if (rc != GNX_OK)
            break;
        if ((rc = tb_build(ctx, ix->genome, genome_cat, node_off, n_nodes, 0, &d_bytes, &d_off)) != GNX_OK)
            break;
        std::vector<int64_t> cand((size_t)n_nodes + 1, 0);
        for (int64_t s = 0; s < n_nodes; ++s) {
            const int64_t L = node_off[s + 1] - node_off[s];
            cand[s + 1] = cand[s] + (L >= seed_len ? (L - seed_len) / seed_step + 1 : 0);
        }
        const int64_t n_cand = cand[n_nodes];
        int64_t n = 0;
        if (n_cand > 0) {
            cudaError_t e;
            if ((e = d_cand.ensure(cand.size() * 8)) != cudaSuccess || (e = d_key.ensure((size_t)n_cand * 8)) != cudaSuccess ||
                (e = d_loc.ensure((size_t)n_cand * 8)) != cudaSuccess || (e = d_valid.ensure(((size_t)n_cand + 1) * 4)) != cudaSuccess ||
                (e = d_dst.ensure(((size_t)n_cand + 1) * 8)) != cudaSuccess) {
                ctx->err = std::string("cudaMalloc failed (seed index): ") + cudaGetErrorString(e);
                rc = GNX_ECUDA;
                break;
            }
            cudaMemcpyAsync(d_cand.p, cand.data(), cand.size() * 8, cudaMemcpyHostToDevice, st);
            gnx::SeedEmitParams E;
            E.genome = d_bytes.as<uint8_t>();
            E.node_off = d_off.as<int64_t>();
            E.cand_off = d_cand.as<int64_t>();
            E.n_nodes = n_nodes;
            E.n_cand = n_cand;
            E.seed_len = seed_len;
            E.seed_step = seed_step;
            E.key = d_key.as<uint64_t>();
            E.loc = d_loc.as<uint64_t>();
            E.valid = d_valid.as<int>();
            gnx::seed_emit_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(E);
            if ((rc = tb_launch_check(ctx)) != GNX_OK)
                break;
            // compaction offsets (exclusive scan of the valid flags) and the stable sort by key: CUB (plumbing)
            size_t tb1 = 0, tb2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb1, d_valid.as<int>(), d_dst.as<int64_t>(), n_cand + 1, st);
            cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_key.as<uint64_t>(), d_key.as<uint64_t>(), d_loc.as<uint64_t>(),
                                            d_loc.as<uint64_t>(), n_cand, 0, 64, st);
            if (d_tmp.ensure(std::max(tb1, tb2) + 16) != cudaSuccess) {
                rc = fail(ctx, GNX_ECUDA, "cudaMalloc failed (seed index scratch)");
                break;
            }
            // the scan reads n_cand + 1 flags: the extra one must be defined
            cudaMemsetAsync(d_valid.as<int>() + n_cand, 0, 4, st);
            size_t tbs = d_tmp.cap;
            cub::DeviceScan::ExclusiveSum(d_tmp.p, tbs, d_valid.as<int>(), d_dst.as<int64_t>(), n_cand + 1, st);
            ctx->launches += 2;
            cudaMemcpyAsync(&n, d_dst.as<int64_t>() + n_cand, 8, cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) {
                rc = fail(ctx, GNX_ECUDA, "seed index: emit/scan failed");
                break;
            }
            if (n > 0) {
                if (d_key2.ensure((size_t)n * 8) != cudaSuccess || d_loc2.ensure((size_t)n * 8) != cudaSuccess ||
                    ix->key.ensure((size_t)n * 8) != cudaSuccess || ix->loc.ensure((size_t)n * 8) != cudaSuccess) {
                    rc = fail(ctx, GNX_ECUDA, "cudaMalloc failed (seed index entries)");
                    break;
                }
                gnx::seed_compact_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, st>>>(
                    d_key.as<uint64_t>(), d_loc.as<uint64_t>(), d_valid.as<int>(), d_dst.as<int64_t>(), n_cand, d_key2.as<uint64_t>(),
                    d_loc2.as<uint64_t>());
                if ((rc = tb_launch_check(ctx)) != GNX_OK)
                    break;
                tbs = d_tmp.cap;
                cub::DeviceRadixSort::SortPairs(d_tmp.p, tbs, d_key2.as<uint64_t>(), ix->key.as<uint64_t>(), d_loc2.as<uint64_t>(),
                                                ix->loc.as<uint64_t>(), n, 0, 64, st);
                ctx->launches += 8;
            }
        }
        ix->n = n;
        // bucket table over the key's top bits: ~1 entry per bucket, at most 2^24 buckets
        int bits = 1;
        while (bits < 24 && bits < 2 * seed_len && ((int64_t)1 << bits) < n)
            ++bits;
        ix->bucket_bits = bits;
        ix->bucket_shift = 2 * seed_len - bits;
        const int64_t nb = ((int64_t)1 << bits) + 1;
        if (ix->bucket.ensure((size_t)nb * 8) != cudaSuccess || ix->key.ensure(8) != cudaSuccess || ix->loc.ensure(8) != cudaSuccess) {
            rc = fail(ctx, GNX_ECUDA, "cudaMalloc failed (seed index buckets)");
            break;
        }
        gnx::seed_bucket_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(ix->key.as<uint64_t>(), n, bits, ix->bucket_shift,
                                                                               ix->bucket.as<int64_t>());
        if ((rc = tb_launch_check(ctx)) != GNX_OK)
            break;
        if (cudaStreamSynchronize(st) != cudaSuccess)
            rc = fail(ctx, GNX_ECUDA, "seed index: sort/bucket failed");
    } while (0);
    DevBuf *tmp[] = {&d_bytes, &d_off, &d_cand, &d_key, &d_loc, &d_valid, &d_dst, &d_tmp, &d_key2, &d_loc2};
    for (DevBuf *b : tmp)
        b->release();
    if (rc != GNX_OK) {
        gnx_seed_index_free(ix);
        return rc;
    }
    *out = ix;
    return GNX_OK;
}

void gnx_seed_index_free(gnx_seed_index *ix)
{
    if (!ix)
        return;
    if (ix->ctx)
        cudaSetDevice(ix->ctx->device);
    gnx_twobit_free(ix->genome);
    ix->key.release();
    ix->loc.release();
    ix->bucket.release();
    delete ix;
}

int gnx_seed_index_info(const gnx_seed_index *ix, int64_t *n_entries)
{
    if (!ix || !n_entries)
        return GNX_EARG;
    *n_entries = ix->n;
    return GNX_OK;
}

int gnx_seed_index_download(gnx_ctx *ctx, const gnx_seed_index *ix, uint64_t *out_key, uint64_t *out_loc)
{
    if (!ctx)
        return GNX_EARG;
    if (!ix || ix->ctx != ctx)
        return fail(ctx, GNX_EARG, "gnx_seed_index_download: the index belongs to another context");
    CU(cudaSetDevice(ctx->device));
    if (ix->n && out_key)
        CU(cudaMemcpy(out_key, ix->key.p, (size_t)ix->n * 8, cudaMemcpyDeviceToHost));
    if (ix->n && out_loc)
        CU(cudaMemcpy(out_loc, ix->loc.p, (size_t)ix->n * 8, cudaMemcpyDeviceToHost));
    return GNX_OK;
}

int gnx_seed_batch(gnx_ctx *ctx, const gnx_seed_index *ix, const uint8_t *reads_cat, const int64_t *read_off, int64_t n_reads,
                   gnx_seed *out_seeds, int64_t *out_seed_off, int64_t seed_cap)
{
    if (!ctx)
        return GNX_EARG;
    if (!ix || ix->ctx != ctx || n_reads < 0 || !read_off || !out_seed_off || (seed_cap > 0 && !out_seeds))
        return fail(ctx, GNX_EARG, "gnx_seed_batch: bad argument");
    out_seed_off[0] = 0;
    if (n_reads == 0)
        return GNX_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    Slot &sl = ctx->slot[0];
    const int64_t base = read_off[0], total_bytes = read_off[n_reads] - base;
    int64_t max_len = 0;
    std::vector<int64_t> rel((size_t)n_reads + 1);
    for (int64_t r = 0; r <= n_reads; ++r) {
        rel[r] = read_off[r] - base;
        if (r && rel[r] < rel[r - 1])
            return fail(ctx, GNX_EARG, "read_off must be non-decreasing");
        if (r)
            max_len = std::max(max_len, rel[r] - rel[r - 1]);
    }
    const int warps = 4;
    const size_t per_warp = 2 * (size_t)((max_len + 15) & ~(int64_t)15) + 16 * (size_t)((max_len + 31) / 32);
    const size_t smem = per_warp * warps;
    if (smem > 200 * 1024)
        return fail(ctx, GNX_ERANGE, "read too long for the seed kernel's shared-memory staging");
    if (smem > 48 * 1024)
        CU(cudaFuncSetAttribute(gnx::seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(sl.alpha.ensure((size_t)total_bytes + 64));
    CU(sl.aoff.ensure(rel.size() * 8));
    CU(sl.counts.ensure((size_t)n_reads * 4 * 2));
    CU(sl.cig_off.ensure(((size_t)n_reads + 1) * 8 * 2));
    CU(sl.misc.ensure(64));
    if (total_bytes)
        CU(cudaMemcpyAsync(sl.alpha.p, reads_cat + base, (size_t)total_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.aoff.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(sl.misc.p, 0, 64, st));
    TbStatus ts;
    int rc = tb_status_reset(ctx, st, ts);
    if (rc != GNX_OK)
        return rc;
    int *d_hits = sl.counts.as<int>(), *d_cnt = d_hits + n_reads;
    int64_t *d_tmp_off = sl.cig_off.as<int64_t>(), *d_out_off = d_tmp_off + n_reads + 1;
    int64_t *d_zero = sl.misc.as<int64_t>(), *d_tot_hits = d_zero + 1, *d_tot_seeds = d_zero + 2;
    gnx::SeedParams P;
    P.reads = sl.alpha.as<uint8_t>();
    P.read_off = sl.aoff.as<int64_t>();
    P.n_reads = n_reads;
    P.ix = ix->view();
    P.genome = ix->genome->view();
    P.seed_len = ix->seed_len;
    P.max_len = (int)max_len;
    P.pass = 0;
    P.hit_count = d_hits;
    P.tmp_off = nullptr;
    P.tmp = nullptr;
    P.seed_count = d_cnt;
    P.status = ts.code;
    P.first_bad = ts.first_bad;
    const int64_t want_blocks = (n_reads + warps - 1) / warps;
    const unsigned grid = (unsigned)std::min<int64_t>(want_blocks, (int64_t)ctx->sm_count * 16);
    gnx::seed_kernel<<<grid, warps * 32, smem, st>>>(P);
    if ((rc = tb_launch_check(ctx)) != GNX_OK)
        return rc;
    if ((rc = enqueue_scan(ctx, sl.partials, d_hits, n_reads, d_tmp_off, d_zero, d_tot_hits, st)) != GNX_OK)
        return rc;
    int64_t tot_hits = 0;
    CU(cudaMemcpyAsync(&tot_hits, d_tot_hits, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(sl.cigars.ensure((size_t)std::max<int64_t>(tot_hits, 1) * sizeof(gnx::SeedRec)));
    P.pass = 1;
    P.tmp_off = d_tmp_off;
    P.tmp = sl.cigars.as<gnx::SeedRec>();
    gnx::seed_kernel<<<grid, warps * 32, smem, st>>>(P);
    if ((rc = tb_launch_check(ctx)) != GNX_OK)
        return rc;
    if ((rc = enqueue_scan(ctx, sl.partials, d_cnt, n_reads, d_out_off, d_zero, d_tot_seeds, st)) != GNX_OK)
        return rc;
    CU(cudaMemcpyAsync(out_seed_off, d_out_off, ((size_t)n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    if ((rc = tb_status_read(ctx, st, ts, "seedMapMemPool")) != GNX_OK)
        return rc;
    const int64_t total = out_seed_off[n_reads];
    if (total > seed_cap)
        return fail(ctx, GNX_ECAP, "gnx_seed_batch: seed_cap too small (out_seed_off is filled; call again with a larger buffer)");
    if (total == 0)
        return GNX_OK;
    CU(sl.trace.ensure((size_t)total * sizeof(gnx::SeedRec)));
    gnx::seed_gather_kernel<<<(unsigned)std::min<int64_t>((n_reads + 7) / 8, (int64_t)ctx->sm_count * 32), 256, 0, st>>>(
        sl.cigars.as<gnx::SeedRec>(), d_tmp_off, d_out_off, n_reads, sl.trace.as<gnx::SeedRec>(), total);
    if ((rc = tb_launch_check(ctx)) != GNX_OK)
        return rc;
    CU(cudaMemcpyAsync(out_seeds, sl.trace.p, (size_t)total * sizeof(gnx::SeedRec), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GNX_OK;
}

} // extern "C"

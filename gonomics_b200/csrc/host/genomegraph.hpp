// genomegraph.hpp -- C++ host-side mirror of gonomics' `dna/dnaTwoBit` package and of the two GPU-backed steps of
// `genomeGraph`'s seed-and-extend over the libgnxalign C ABI.  Same names, argument order
// and meaning as the Go packages (file:line relative to the gonomics tree); Go's log.Fatalf -> std::runtime_error,
// Go's index-out-of-range panic -> std::out_of_range.  Every call goes to the GPU; there is no CPU path.
#pragma once
#include "align.hpp"

#include <algorithm>
#include <memory>

namespace gonomics {

namespace dnaTwoBit {

namespace detail {
inline void check(int rc)
{
    if (rc == GNX_OK)
        return;
    const std::string msg = gnx_last_error(align::detail::ctx());
    if (rc == GNX_EINDEX || rc == GNX_EBASE)
        throw std::out_of_range("runtime error: index out of range: " + msg);
    if (rc == GNX_EOFFSET) // perfectAlign.go:24-26, :63-65
        throw std::runtime_error("Error: Different offsets when comparing sequences");
    throw std::runtime_error("gnxalign: " + msg);
}
} // namespace detail

// A batch of TwoBit sequences resident on the GPU (gnx_twobit): packed once, queried many times.
class TwoBitSet {
  public:
    TwoBitSet(const std::vector<std::vector<dna::Base>> &seqs, int lead = 0)
    {
        std::vector<int64_t> off(seqs.size() + 1, 0);
        for (size_t i = 0; i < seqs.size(); ++i)
            off[i + 1] = off[i] + (int64_t)seqs[i].size();
        std::vector<dna::Base> cat((size_t)off.back() + 1);
        for (size_t i = 0; i < seqs.size(); ++i)
            std::copy(seqs[i].begin(), seqs[i].end(), cat.begin() + off[i]);
        detail::check(gnx_twobit_new(align::detail::ctx(), cat.data(), off.data(), (int64_t)seqs.size(), lead, &h_));
    }
    ~TwoBitSet() { gnx_twobit_free(h_); }
    TwoBitSet(const TwoBitSet &) = delete;
    TwoBitSet &operator=(const TwoBitSet &) = delete;
    gnx_twobit *handle() const { return h_; }

  private:
    gnx_twobit *h_ = nullptr;
};

// dnaTwoBit.TwoBit (dnaTwoBit.go:14-17): Seq / Len are the reference's fields; `set` keeps the device copy the
// package functions run on.
struct TwoBit {
    std::vector<uint64_t> Seq;
    int Len = 0;
    std::shared_ptr<TwoBitSet> set;
    int64_t idx = 0;
};

inline TwoBit fromSet(std::shared_ptr<TwoBitSet> set, int64_t idx = 0)
{
    int64_t n = 0, words = 0;
    gnx_twobit_info(set->handle(), &n, &words);
    std::vector<uint64_t> all((size_t)words + 1);
    std::vector<int64_t> woff((size_t)n + 1), len((size_t)n + 1);
    detail::check(gnx_twobit_download(align::detail::ctx(), set->handle(), all.data(), woff.data(), len.data()));
    TwoBit t;
    t.Seq.assign(all.begin() + woff[(size_t)idx], all.begin() + woff[(size_t)idx + 1]);
    t.Len = (int)len[(size_t)idx];
    t.set = std::move(set);
    t.idx = idx;
    return t;
}

// dnaTwoBit.NewTwoBit (dnaTwoBit.go:68)
inline TwoBit NewTwoBit(const std::vector<dna::Base> &inSeq)
{
    return fromSet(std::make_shared<TwoBitSet>(std::vector<std::vector<dna::Base>>{inSeq}, 0));
}
// dnaTwoBit.NewTwoBitRainbow (rainbow.go:8): element k holds k leading 'A's
inline std::vector<TwoBit> NewTwoBitRainbow(const std::vector<dna::Base> &inSeq)
{
    std::vector<TwoBit> out;
    for (int k = 0; k < 32; ++k)
        out.push_back(fromSet(std::make_shared<TwoBitSet>(std::vector<std::vector<dna::Base>>{inSeq}, k)));
    return out;
}
// dnaTwoBit.GetBase (dnaTwoBit.go:59)
inline dna::Base GetBase(const TwoBit &frag, unsigned pos)
{
    const int64_t s = frag.idx, p = (int64_t)pos;
    uint8_t b = 0;
    detail::check(gnx_twobit_get_bases(align::detail::ctx(), frag.set->handle(), &s, &p, 1, &b));
    return b;
}
namespace detail {
inline int count(int dir, const TwoBit &one, int startOne, const TwoBit &two, int startTwo)
{
    const int64_t q1 = one.idx, s1 = startOne, q2 = two.idx, s2 = startTwo;
    int64_t out = 0;
    check(gnx_twobit_count_matches(align::detail::ctx(), dir, one.set->handle(), two.set->handle(), &q1, &s1, &q2, &s2, 1,
                                   &out));
    return (int)out;
}
} // namespace detail
// dnaTwoBit.CountRightMatches (perfectAlign.go:10)
inline int CountRightMatches(const TwoBit &one, int startOne, const TwoBit &two, int startTwo)
{
    return detail::count(GNX_MATCH_RIGHT, one, startOne, two, startTwo);
}
// dnaTwoBit.CountLeftMatches (perfectAlign.go:49)
inline int CountLeftMatches(const TwoBit &one, int startOne, const TwoBit &two, int startTwo)
{
    return detail::count(GNX_MATCH_LEFT, one, startOne, two, startTwo);
}

} // namespace dnaTwoBit

namespace cigar {
// cigar.Cigar{RunLength int; Op byte} (cigar/cigar.go:21-24); Op is 'M', 'I' or 'D' here
struct Cigar {
    int64_t RunLength;
    char Op;
};
} // namespace cigar

namespace genomeGraph {

// genomeGraph.SeedDev (genomeGraph/index.go:11-19); NextPart is always nil for nodes without edges
struct SeedDev {
    uint32_t TargetId, TargetStart, QueryStart, Length;
    bool PosStrand;
    uint32_t TotalLength;
};

struct DynamicAln { // what LeftDynamicAln / RightDynamicAln return
    int64_t score;
    std::vector<cigar::Cigar> route; // traceback order, as in the reference
    int i, j;
};

namespace detail {
inline std::vector<DynamicAln> extend(int side, const std::vector<std::vector<dna::Base>> &alphas,
                                      const std::vector<std::vector<dna::Base>> &betas, const align::Matrix &scores,
                                      int64_t gapPen)
{
    const size_t n = alphas.size();
    std::vector<int64_t> aoff(n + 1, 0), boff(n + 1, 0);
    for (size_t k = 0; k < n; ++k) {
        aoff[k + 1] = aoff[k] + (int64_t)alphas[k].size();
        boff[k + 1] = boff[k] + (int64_t)betas[k].size();
    }
    std::vector<dna::Base> acat((size_t)aoff[n] + 1), bcat((size_t)boff[n] + 1);
    for (size_t k = 0; k < n; ++k) {
        std::copy(alphas[k].begin(), alphas[k].end(), acat.begin() + aoff[k]);
        std::copy(betas[k].begin(), betas[k].end(), bcat.begin() + boff[k]);
    }
    const std::vector<int64_t> flat = align::detail::flatten(scores);
    std::vector<int64_t> score(n + 1), ei(n + 1), ej(n + 1), coff(n + 1);
    std::vector<gnx_cigar> cig((size_t)(aoff[n] + boff[n]) + n + 1);
    align::detail::check(gnx_extend_batch(align::detail::ctx(), side, acat.data(), aoff.data(), bcat.data(), boff.data(),
                                          (int64_t)n, flat.data(), (int)scores.size(), gapPen, 1, score.data(), ei.data(),
                                          ej.data(), cig.data(), coff.data(), (int64_t)cig.size()));
    std::vector<DynamicAln> out(n);
    for (size_t k = 0; k < n; ++k) {
        out[k].score = score[k];
        out[k].i = (int)ei[k];
        out[k].j = (int)ej[k];
        for (int64_t c = coff[k]; c < coff[k + 1]; ++c)
            out[k].route.push_back({cig[(size_t)c].run_length, (char)cig[(size_t)c].op});
    }
    return out;
}
} // namespace detail

// genomeGraph.LeftDynamicAln (genomeGraph/search.go:234); the scratch matrix and score keeper of the reference
// carry no state into the call and are not parameters here
inline DynamicAln LeftDynamicAln(const std::vector<dna::Base> &alpha, const std::vector<dna::Base> &beta,
                                 const align::Matrix &scores, int64_t gapPen = -600)
{
    return detail::extend(GNX_EXT_LEFT, {alpha}, {beta}, scores, gapPen)[0];
}
// genomeGraph.RightDynamicAln (genomeGraph/search.go:276)
inline DynamicAln RightDynamicAln(const std::vector<dna::Base> &alpha, const std::vector<dna::Base> &beta,
                                  const align::Matrix &scores, int64_t gapPen = -600)
{
    return detail::extend(GNX_EXT_RIGHT, {alpha}, {beta}, scores, gapPen)[0];
}
// batched forms: one GPU call for a block of extension pairs
inline std::vector<DynamicAln> LeftDynamicAlnBatch(const std::vector<std::vector<dna::Base>> &alphas,
                                                   const std::vector<std::vector<dna::Base>> &betas,
                                                   const align::Matrix &scores, int64_t gapPen = -600)
{
    return detail::extend(GNX_EXT_LEFT, alphas, betas, scores, gapPen);
}
inline std::vector<DynamicAln> RightDynamicAlnBatch(const std::vector<std::vector<dna::Base>> &alphas,
                                                    const std::vector<std::vector<dna::Base>> &betas,
                                                    const align::Matrix &scores, int64_t gapPen = -600)
{
    return detail::extend(GNX_EXT_RIGHT, alphas, betas, scores, gapPen);
}

// genomeGraph.heapSortSeeds (genomeGraph/search.go:339-373): in-place min-heap sort = descending TotalLength with
// the reference's order among equal lengths
inline void heapSortSeeds(std::vector<SeedDev> &a)
{
    auto heapify = [&](size_t size, size_t i) {
        for (;;) {
            const size_t l = 2 * i + 1, r = 2 * i + 2;
            size_t m = (l < size && a[l].TotalLength < a[i].TotalLength) ? l : i;
            if (r < size && a[r].TotalLength < a[m].TotalLength)
                m = r;
            if (m == i)
                return;
            std::swap(a[i], a[m]);
            i = m;
        }
    };
    for (size_t i = a.size() / 2; i-- > 0;)
        heapify(a.size(), i);
    size_t size = a.size();
    for (size_t i = a.size(); i-- > 1;) {
        std::swap(a[0], a[i]);
        --size;
        heapify(size, 0);
    }
}

// genomeGraph.IndexGenomeIntoMap (genomeGraph/index.go:21) for nodes without edges, kept on the GPU together with
// the nodes' TwoBit encoding; SeedMapMemPool is genomeGraph.seedMapMemPool (search.go:567) for a block of reads.
class SeedIndex {
  public:
    SeedIndex(const std::vector<std::vector<dna::Base>> &nodes, int seedLen, int seedStep)
    {
        std::vector<int64_t> off(nodes.size() + 1, 0);
        for (size_t i = 0; i < nodes.size(); ++i)
            off[i + 1] = off[i] + (int64_t)nodes[i].size();
        std::vector<dna::Base> cat((size_t)off.back() + 1);
        for (size_t i = 0; i < nodes.size(); ++i)
            std::copy(nodes[i].begin(), nodes[i].end(), cat.begin() + off[i]);
        const int rc = gnx_seed_index_new(align::detail::ctx(), cat.data(), off.data(), (int64_t)nodes.size(), seedLen,
                                          seedStep, &h_);
        if (rc == GNX_EARG) // index.go:22-24 log.Fatalf
            throw std::runtime_error(std::string("Error: ") + gnx_last_error(align::detail::ctx()));
        dnaTwoBit::detail::check(rc);
    }
    ~SeedIndex() { gnx_seed_index_free(h_); }
    SeedIndex(const SeedIndex &) = delete;
    SeedIndex &operator=(const SeedIndex &) = delete;

    std::vector<std::vector<SeedDev>> SeedMapMemPool(const std::vector<std::vector<dna::Base>> &reads) const
    {
        const size_t n = reads.size();
        std::vector<int64_t> off(n + 1, 0), soff(n + 1, 0);
        for (size_t i = 0; i < n; ++i)
            off[i + 1] = off[i] + (int64_t)reads[i].size();
        std::vector<dna::Base> cat((size_t)off[n] + 1);
        for (size_t i = 0; i < n; ++i)
            std::copy(reads[i].begin(), reads[i].end(), cat.begin() + off[i]);
        std::vector<gnx_seed> seeds(8 * n + 64);
        int rc = gnx_seed_batch(align::detail::ctx(), h_, cat.data(), off.data(), (int64_t)n, seeds.data(), soff.data(),
                                (int64_t)seeds.size());
        if (rc == GNX_ECAP) { // soff is filled: retry with the exact size
            seeds.assign((size_t)soff[n] + 1, gnx_seed{});
            rc = gnx_seed_batch(align::detail::ctx(), h_, cat.data(), off.data(), (int64_t)n, seeds.data(), soff.data(),
                                (int64_t)seeds.size());
        }
        dnaTwoBit::detail::check(rc);
        std::vector<std::vector<SeedDev>> out(n);
        for (size_t r = 0; r < n; ++r) {
            for (int64_t k = soff[r]; k < soff[r + 1]; ++k) {
                const gnx_seed &s = seeds[(size_t)k];
                out[r].push_back({s.target_id, s.target_start, s.query_start, s.length, s.pos_strand != 0, s.total_length});
            }
            // the reference's final ordering (search.go:596-600); lists > 100 use Go's unstable sort.Slice there
            if (out[r].size() > 100)
                std::stable_sort(out[r].begin(), out[r].end(),
                                 [](const SeedDev &x, const SeedDev &y) { return x.TotalLength > y.TotalLength; });
            else
                heapSortSeeds(out[r]);
        }
        return out;
    }

  private:
    gnx_seed_index *h_ = nullptr;
};

} // namespace genomeGraph
} // namespace gonomics

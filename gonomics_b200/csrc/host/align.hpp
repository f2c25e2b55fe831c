// align.hpp -- C++ host-side mirror of gonomics' `align` package over the libgnxalign C ABI.
//
// The reference is Go and the build image has no Go toolchain, so this header is the compiled-language
// host layer above include/gnxalign.h: same function names, argument order and meaning as the Go package
// (file:line citations are relative to the gonomics tree), same results, and the reference's error
// behaviour mapped to C++ exceptions (Go panic -> std::out_of_range, log.Fatalf -> std::runtime_error).
// The cgo binding a gonomics maintainer would add is in integration/go/align/.
#pragma once
#include "../../../include/gnxalign.h"

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace gonomics {
namespace dna {
using Base = uint8_t; // dna/dna.go:5-21: A,C,G,T,N = 0..4, lower case 5..9, Gap 10
inline std::vector<Base> StringToBases(const std::string &s)
{ // dna/convert.go
    std::vector<Base> b;
    b.reserve(s.size());
    for (char ch : s) {
        switch (ch) {
        case 'A': b.push_back(0); break;
        case 'C': b.push_back(1); break;
        case 'G': b.push_back(2); break;
        case 'T': b.push_back(3); break;
        case 'N': b.push_back(4); break;
        case 'a': b.push_back(5); break;
        case 'c': b.push_back(6); break;
        case 'g': b.push_back(7); break;
        case 't': b.push_back(8); break;
        case 'n': b.push_back(9); break;
        case '-': b.push_back(10); break;
        default: throw std::runtime_error(std::string("invalid base '") + ch + "'");
        }
    }
    return b;
}
} // namespace dna

namespace align {

using ColType = uint8_t; // align/align.go:12-18
constexpr ColType ColM = 0, ColI = 1, ColD = 2;
using Cigar = gnx_cigar; // {int64 run_length (RunLength); uint8 op (Op)}: align/align.go:21-24
using Matrix = std::vector<std::vector<int64_t>>;

// align/align.go:28-64
inline const Matrix DefaultScoreMatrix = {{91, -114, -31, -123, -44}, {-114, 100, -125, -31, -43},
                                          {-31, -125, 100, -114, -43}, {-123, -31, -114, 91, -44},
                                          {-44, -43, -43, -44, -43}};
inline const Matrix HumanChimpTwoScoreMatrix = {{90, -330, -236, -356, -208}, {-330, 100, -318, -236, -196},
                                                {-236, -318, 100, -330, -196}, {-356, -236, -330, 90, -208},
                                                {-208, -196, -196, -208, -202}};

namespace detail {
struct Ctx { // one gnx_ctx per calling thread (contexts are not thread-safe)
    gnx_ctx *h;
    Ctx() : h(gnx_create(0, 0))
    {
        if (!h)
            throw std::runtime_error(std::string("gnxalign: ") + gnx_last_error(nullptr));
    }
    ~Ctx() { gnx_destroy(h); }
};
inline gnx_ctx *ctx()
{
    thread_local Ctx c;
    return c.h;
}
inline void check(int rc)
{
    if (rc == GNX_OK)
        return;
    const std::string msg = gnx_last_error(ctx());
    if (rc == GNX_EBASE) // Go: scores[alpha[i]][beta[j]] index out of range
        throw std::out_of_range("runtime error: index out of range: " + msg);
    throw std::runtime_error("gnxalign: " + msg);
}
inline std::vector<int64_t> flatten(const Matrix &s)
{
    std::vector<int64_t> f;
    for (const auto &row : s)
        f.insert(f.end(), row.begin(), row.end());
    return f;
}
// kind 0 AffineGap_highMem, 1 AffineGapLocal, 2 ConstGap_highMem, 3 AffineGapChunk
inline std::pair<int64_t, std::vector<Cigar>> one(int kind, const std::vector<dna::Base> &alpha,
                                                  const std::vector<dna::Base> &beta, const Matrix &scores, int64_t gapOpen,
                                                  int64_t gapExtend, int64_t chunk = 1)
{
    const std::vector<int64_t> flat = flatten(scores);
    const int64_t aoff[2] = {0, (int64_t)alpha.size()}, boff[2] = {0, (int64_t)beta.size()};
    int64_t score = 0, coff[2] = {0, 0};
    std::vector<Cigar> route(alpha.size() + beta.size() + 1);
    int rc;
    if (kind == 2)
        rc = gnx_const_batch(ctx(), alpha.data(), aoff, beta.data(), boff, 1, flat.data(), (int)scores.size(), gapOpen, 1,
                             &score, route.data(), coff, (int64_t)route.size());
    else if (kind == 3)
        rc = gnx_affine_chunk_batch(ctx(), alpha.data(), aoff, beta.data(), boff, 1, flat.data(), (int)scores.size(),
                                    gapOpen, gapExtend, chunk, &score, route.data(), coff, (int64_t)route.size());
    else
        rc = gnx_affine_batch(ctx(), alpha.data(), aoff, beta.data(), boff, 1, flat.data(), (int)scores.size(), gapOpen,
                              gapExtend, kind == 1 ? GNX_FREE_END : GNX_GLOBAL, 1, &score, route.data(), coff,
                              (int64_t)route.size());
    check(rc);
    route.resize((size_t)coff[1]);
    return {score, std::move(route)};
}
} // namespace detail

// align/affineGap_highMem.go:99
inline std::pair<int64_t, std::vector<Cigar>> AffineGap_highMem(const std::vector<dna::Base> &alpha,
                                                                const std::vector<dna::Base> &beta, const Matrix &scores,
                                                                int64_t gapOpen, int64_t gapExtend)
{
    return detail::one(0, alpha, beta, scores, gapOpen, gapExtend);
}
// align/affineGap_highMem.go:105
inline std::pair<int64_t, std::vector<Cigar>> AffineGapLocal(const std::vector<dna::Base> &target,
                                                             const std::vector<dna::Base> &query, const Matrix &scores,
                                                             int64_t gapOpen, int64_t gapExtend)
{
    return detail::one(1, target, query, scores, gapOpen, gapExtend);
}
namespace detail {
// One checkerboard: the low-memory drivers equal the high-memory result.  Past one board the reference's stitching
// has defects that the GPU path does not reproduce: fail instead of returning a different cigar.
inline void require_one_board(const std::vector<dna::Base> &alpha, const std::vector<dna::Base> &beta, int ci, int cj,
                              const char *what)
{
    if (alpha.empty() || beta.empty())
        throw std::out_of_range("runtime error: index out of range (empty sequence)");
    if ((int64_t)alpha.size() > ci || (int64_t)beta.size() > cj)
        throw std::invalid_argument(std::string(what) + ": input spans more than one checkerboard; the reference's "
                                    "multi-board stitching is not reproduced -- call the _highMem function");
}
} // namespace detail
// align/affineGap.go:73
inline std::pair<int64_t, std::vector<Cigar>> AffineGap_customizeCheckersize(const std::vector<dna::Base> &alpha,
                                                                             const std::vector<dna::Base> &beta,
                                                                             const Matrix &scores, int64_t gapOpen,
                                                                             int64_t gapExtend, int checkersize_i, int checkersize_j)
{
    detail::require_one_board(alpha, beta, checkersize_i, checkersize_j, "AffineGap_customizeCheckersize");
    return detail::one(0, alpha, beta, scores, gapOpen, gapExtend);
}
// align/affineGap.go:59
inline std::pair<int64_t, std::vector<Cigar>> AffineGap(const std::vector<dna::Base> &alpha, const std::vector<dna::Base> &beta,
                                                        const Matrix &scores, int64_t gapOpen, int64_t gapExtend)
{
    return AffineGap_customizeCheckersize(alpha, beta, scores, gapOpen, gapExtend, 10000, 10000);
}
// align/constGap_highMem.go:11
inline std::pair<int64_t, std::vector<Cigar>> ConstGap_highMem(const std::vector<dna::Base> &alpha,
                                                               const std::vector<dna::Base> &beta, const Matrix &scores,
                                                               int64_t gapPen)
{
    return detail::one(2, alpha, beta, scores, gapPen, 0);
}
// align/constGap.go:73
inline std::pair<int64_t, std::vector<Cigar>> ConstGap_customizeCheckersize(const std::vector<dna::Base> &alpha,
                                                                            const std::vector<dna::Base> &beta,
                                                                            const Matrix &scores, int64_t gapPen, int checkersize_i,
                                                                            int checkersize_j)
{
    detail::require_one_board(alpha, beta, checkersize_i, checkersize_j, "ConstGap_customizeCheckersize");
    return detail::one(2, alpha, beta, scores, gapPen, 0);
}
// align/constGap.go:13
inline std::pair<int64_t, std::vector<Cigar>> ConstGap(const std::vector<dna::Base> &alpha, const std::vector<dna::Base> &beta,
                                                       const Matrix &scores, int64_t gapPen)
{
    return ConstGap_customizeCheckersize(alpha, beta, scores, gapPen, 10000, 10000);
}
// align/affineGap_highMem.go:227
inline std::pair<int64_t, std::vector<Cigar>> AffineGapChunk(const std::vector<dna::Base> &alpha,
                                                             const std::vector<dna::Base> &beta, const Matrix &scores,
                                                             int64_t gapOpen, int64_t gapExtend, int64_t chunkSize)
{
    return detail::one(3, alpha, beta, scores, gapOpen, gapExtend, chunkSize);
}

// align/view.go:25-33
inline std::string PrintCigar(const std::vector<Cigar> &ops)
{
    std::string s;
    for (const Cigar &c : ops)
        s += std::to_string(c.run_length) + "MID"[c.op];
    return s;
}

// align/affineGap_highMem.go:110-115
struct TargetQueryPair {
    std::vector<dna::Base> Target, Query;
    int64_t Score = 0;
    std::vector<Cigar> Cigar_;
};

// A Go-style buffered channel (capacity 1000 in the reference, :121-122).
template <typename T> class Chan {
  public:
    explicit Chan(size_t cap) : cap_(cap) {}
    void send(T v)
    {
        std::unique_lock<std::mutex> l(m_);
        not_full_.wait(l, [&] { return q_.size() < cap_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    bool recv(T &out) // false once the channel is closed and drained
    {
        std::unique_lock<std::mutex> l(m_);
        not_empty_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty())
            return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    bool try_recv(T &out)
    {
        std::lock_guard<std::mutex> l(m_);
        if (q_.empty())
            return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close()
    {
        std::lock_guard<std::mutex> l(m_);
        closed_ = true;
        not_empty_.notify_all();
    }

  private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// align/affineGap_highMem.go:120-179 GoAffineGapLocalEngine: same channel interface and FIFO order, but the
// worker drains whatever is queued into ONE GPU batch per iteration (the performant boundary is batch-shaped).
struct AffineGapLocalEngine {
    std::shared_ptr<Chan<TargetQueryPair>> inputs, outputs;
    std::thread worker;
    ~AffineGapLocalEngine()
    {
        if (worker.joinable()) {
            inputs->close();
            worker.join();
        }
    }
};

inline std::unique_ptr<AffineGapLocalEngine> GoAffineGapLocalEngine(const Matrix &scores, int64_t gapOpen, int64_t gapExtend)
{
    auto e = std::make_unique<AffineGapLocalEngine>();
    e->inputs = std::make_shared<Chan<TargetQueryPair>>(1000);
    e->outputs = std::make_shared<Chan<TargetQueryPair>>(1000);
    auto in = e->inputs, out = e->outputs;
    e->worker = std::thread([in, out, scores, gapOpen, gapExtend] {
        const std::vector<int64_t> flat = detail::flatten(scores);
        std::vector<TargetQueryPair> batch;
        TargetQueryPair p;
        try { // an exception escaping a std::thread is std::terminate: a failing batch (e.g. a base >= dim, the reference
              // goroutine's panic) ends the stream instead -- outputs is closed, consumers never block forever
        while (in->recv(p)) {
            batch.clear();
            batch.push_back(std::move(p));
            while (batch.size() < (1u << 16) && in->try_recv(p))
                batch.push_back(std::move(p));
            const int64_t n = (int64_t)batch.size();
            std::vector<int64_t> aoff((size_t)n + 1, 0), boff((size_t)n + 1, 0), sc((size_t)n), coff((size_t)n + 1);
            for (int64_t k = 0; k < n; ++k) {
                aoff[(size_t)k + 1] = aoff[(size_t)k] + (int64_t)batch[(size_t)k].Target.size();
                boff[(size_t)k + 1] = boff[(size_t)k] + (int64_t)batch[(size_t)k].Query.size();
            }
            std::vector<dna::Base> acat((size_t)aoff[(size_t)n]), bcat((size_t)boff[(size_t)n]);
            for (int64_t k = 0; k < n; ++k) {
                std::copy(batch[(size_t)k].Target.begin(), batch[(size_t)k].Target.end(), acat.begin() + aoff[(size_t)k]);
                std::copy(batch[(size_t)k].Query.begin(), batch[(size_t)k].Query.end(), bcat.begin() + boff[(size_t)k]);
            }
            std::vector<Cigar> cig((size_t)(16 * n + 64));
            int rc = gnx_affine_batch(detail::ctx(), acat.data(), aoff.data(), bcat.data(), boff.data(), n, flat.data(),
                                      (int)scores.size(), gapOpen, gapExtend, GNX_FREE_END, 1, sc.data(), cig.data(),
                                      coff.data(), (int64_t)cig.size());
            if (rc == GNX_ECAP) {
                cig.resize((size_t)coff[(size_t)n]);
                rc = gnx_copy_last_cigars(detail::ctx(), cig.data(), (int64_t)cig.size());
            }
            detail::check(rc);
            for (int64_t k = 0; k < n; ++k) {
                batch[(size_t)k].Score = sc[(size_t)k];
                batch[(size_t)k].Cigar_.assign(cig.begin() + coff[(size_t)k], cig.begin() + coff[(size_t)k + 1]);
                out->send(std::move(batch[(size_t)k]));
            }
        }
        } catch (const std::exception &) {
        }
        out->close(); // close(outputs) when inputs closes (:178) -- and when a batch fails
    });
    return e;
}

} // namespace align
} // namespace gonomics

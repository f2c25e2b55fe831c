// Replays dna/dnaTwoBit's known-answer tests (perfectAlign_test.go:28-110 TestCounting, dnaTwoBit_test.go:17-42
// TestDnaToFromString) through the C++ host mirror, and checks the seed step and the extend step against plain
// base-level restatements (perfectAlign_test.go:81-95 currentMethodRight / currentMethodLeft style).  Needs a GPU.
#include "../../gonomics_b200/csrc/host/genomegraph.hpp"

#include <cstdio>
#include <cstdlib>
#include <random>

using namespace gonomics;

static int fails = 0;
#define EXPECT(cond, ...)                                                                                               \
    do {                                                                                                                \
        if (!(cond)) {                                                                                                  \
            ++fails;                                                                                                    \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__);                                                        \
            fprintf(stderr, __VA_ARGS__);                                                                               \
            fprintf(stderr, "\n");                                                                                      \
        }                                                                                                               \
    } while (0)

int main()
{
    const std::string rep = "TGCACTAGTCATACAGTA";
    auto longSeq = [&](const char *head, const char *tail) {
        std::string s = head;
        while (s.size() + rep.size() <= 143)
            s += rep;
        return s.substr(0, 143) + tail;
    };
    (void)longSeq;
    // the literals of perfectAlign_test.go:21-31
    const std::string one = "ATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGG";
    const std::string two = "ATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGT";
    const std::string three = "CTGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGT";
    const std::string sOne = "CCCCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGG";
    const std::string sTwo = "ACCTACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGTATGCACTAGTCATACAGT";
    struct Case { std::string a, b; int sa, sb, left, right; };
    const Case cases[] = {{sOne, sTwo, 1, 1, 1, 2}, {"T", "T", 0, 0, 1, 1}, {one, two, 100, 100, 101, 43}, {three, two, 90, 90, 90, 54}};
    for (const Case &c : cases) {
        const auto A = dnaTwoBit::NewTwoBit(dna::StringToBases(c.a)), B = dnaTwoBit::NewTwoBit(dna::StringToBases(c.b));
        EXPECT(dnaTwoBit::CountLeftMatches(A, c.sa, B, c.sb) == c.left, "left matches at %d", c.sa);
        EXPECT(dnaTwoBit::CountRightMatches(A, c.sa, B, c.sb) == c.right, "right matches at %d", c.sa);
    }
    for (const char *s : {"TCATACGTTTTTTTTTTTTTCTGTC", "TCAAAACCCCCGGGGTTTTTCTGTC", "TCATACGTACGTACGTCCCCCTGCCCC", "TCATGGGGGGGGCCAGTACGTTGGCT"}) {
        const auto frag = dnaTwoBit::NewTwoBit(dna::StringToBases(s));
        EXPECT(frag.Len == (int)std::string(s).size(), "Len");
        EXPECT(dnaTwoBit::GetBase(frag, 0) == 3 && dnaTwoBit::GetBase(frag, 1) == 1 && dnaTwoBit::GetBase(frag, 2) == 0 &&
                   dnaTwoBit::GetBase(frag, 21) == 3 && dnaTwoBit::GetBase(frag, 24) == 1, "GetBase on %s", s);
    }
    bool threw = false;
    try {
        const auto A = dnaTwoBit::NewTwoBit(dna::StringToBases(one));
        dnaTwoBit::CountRightMatches(A, 3, A, 4);
    } catch (const std::runtime_error &) {
        threw = true;
    }
    EXPECT(threw, "different offsets must be fatal");

    // seeds on a small two-node genome against base-level loops (clean A,C,G,T)
    std::mt19937 rng(17);
    auto randSeq = [&](size_t n) {
        std::vector<dna::Base> s(n);
        for (auto &b : s)
            b = (dna::Base)(rng() & 3);
        return s;
    };
    std::vector<std::vector<dna::Base>> nodes = {randSeq(3000), randSeq(1000)};
    const int seedLen = 20, seedStep = 8;
    genomeGraph::SeedIndex index(nodes, seedLen, seedStep);
    std::vector<std::vector<dna::Base>> reads;
    for (int r = 0; r < 40; ++r) {
        const auto &node = nodes[r & 1];
        const size_t s = rng() % (node.size() - 160);
        std::vector<dna::Base> read(node.begin() + s, node.begin() + s + 100 + rng() % 50);
        read[rng() % read.size()] ^= 1;
        if (r % 3 == 0) { // reverse complement
            std::reverse(read.begin(), read.end());
            for (auto &b : read)
                b = 3 - b;
        }
        reads.push_back(read);
    }
    const auto got = index.SeedMapMemPool(reads);
    for (size_t r = 0; r < reads.size(); ++r) {
        std::vector<genomeGraph::SeedDev> want;
        const auto &fw = reads[r];
        std::vector<dna::Base> rc(fw.rbegin(), fw.rend());
        for (auto &b : rc)
            b = 3 - b;
        for (int start = 0; start + seedLen <= (int)fw.size(); ++start)
            for (int strand = 1; strand >= 0; --strand) {
                const auto &q = strand ? fw : rc;
                for (size_t ni = 0; ni < nodes.size(); ++ni)
                    for (int pos = 0; pos + seedLen <= (int)nodes[ni].size(); pos += seedStep) {
                        if (!std::equal(q.begin() + start, q.begin() + start + seedLen, nodes[ni].begin() + pos))
                            continue;
                        int left = 0;
                        while (start - left >= 0 && pos - left >= 0 && q[start - left] == nodes[ni][pos - left])
                            ++left;
                        const int rs = start - (left - 1), ns = pos - (left - 1);
                        int right = 0;
                        while (rs + right < (int)q.size() && ns + right < (int)nodes[ni].size() && q[rs + right] == nodes[ni][ns + right])
                            ++right;
                        want.push_back({(uint32_t)ni, (uint32_t)ns, (uint32_t)rs, (uint32_t)right, strand == 1, (uint32_t)right});
                    }
            }
        genomeGraph::heapSortSeeds(want);
        bool same = want.size() == got[r].size();
        for (size_t k = 0; same && k < want.size(); ++k)
            same = want[k].TargetId == got[r][k].TargetId && want[k].TargetStart == got[r][k].TargetStart &&
                   want[k].QueryStart == got[r][k].QueryStart && want[k].Length == got[r][k].Length &&
                   want[k].PosStrand == got[r][k].PosStrand;
        EXPECT(same, "seeds of read %zu (%zu vs %zu)", r, got[r].size(), want.size());
    }

    // extend step: an exact copy extends with all matches; route in traceback order
    const auto tgt = randSeq(120);
    const std::vector<dna::Base> q(tgt.end() - 60, tgt.end());
    const auto L = genomeGraph::LeftDynamicAln(tgt, q, align::HumanChimpTwoScoreMatrix);
    int64_t perfect = 0;
    for (auto b : q)
        perfect += align::HumanChimpTwoScoreMatrix[b][b];
    EXPECT(L.score == perfect && L.route.size() == 1 && L.route[0].Op == 'M' && L.route[0].RunLength == 60 && L.i == 60 && L.j == 0,
           "LeftDynamicAln on an exact suffix");
    const std::vector<dna::Base> q2(tgt.begin(), tgt.begin() + 70);
    const auto R = genomeGraph::RightDynamicAln(tgt, q2, align::HumanChimpTwoScoreMatrix);
    int64_t perfect2 = 0;
    for (auto b : q2)
        perfect2 += align::HumanChimpTwoScoreMatrix[b][b];
    EXPECT(R.score == perfect2 && R.route.size() == 1 && R.route[0].Op == 'M' && R.route[0].RunLength == 70 && R.i == 70 && R.j == 70,
           "RightDynamicAln on an exact prefix");
    if (fails == 0)
        printf("genomegraph host mirror: all checks passed\n");
    return fails ? 1 : 0;
}

// Replays align/affineGap_test.go (TestAffineGapLocal, TestGoAffineGapLocalEngine), view_test.go samples and
// the chunk / const-gap known answers through the C++ host mirror.  Exit code 0 = all assertions hold.
#include "../../gonomics_b200/csrc/host/align.hpp"

#include <cstdio>
#include <cstdlib>

using namespace gonomics;
using align::PrintCigar;

static int fails = 0;
#define EXPECT(cond)                                                    \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            ++fails;                                                    \
        }                                                               \
    } while (0)

int main()
{
    struct L { const char *t, *q; int64_t o, e, score; const char *cig; };
    const L local[] = {{"TCACTTTCGCACGTT", "CACACG", -600, -150, 460, "7D6M2D"},
                       {"CACACACACACACACATTTGACATAGACATA", "CTTTTGA", -600, -150, 441, "14D7M10D"},
                       {"GACTTTT", "GAC", -600, -150, 291, "3M4D"},
                       {"TTTTGAC", "GAC", -600, -150, 291, "4D3M"},
                       {"TTTTATGCCCAAAAGGGATGTTTT", "ATGCCCGGGATG", -200, -50, 764, "4D6M4D6M4D"}};
    for (const L &c : local) { // align/affineGap_test.go:120-155
        auto r = align::AffineGapLocal(dna::StringToBases(c.t), dna::StringToBases(c.q), align::DefaultScoreMatrix, c.o, c.e);
        EXPECT(r.first == c.score && PrintCigar(r.second) == c.cig);
    }
    { // engine, FIFO (:157-192) + batched burst
        auto eng = align::GoAffineGapLocalEngine(align::DefaultScoreMatrix, -600, -150);
        for (int rep = 0; rep < 3; ++rep)
            for (int k = 0; k < 4; ++k) {
                align::TargetQueryPair p;
                p.Target = dna::StringToBases(local[k].t);
                p.Query = dna::StringToBases(local[k].q);
                eng->inputs->send(std::move(p));
            }
        for (int rep = 0; rep < 3; ++rep)
            for (int k = 0; k < 4; ++k) {
                align::TargetQueryPair p;
                EXPECT(eng->outputs->recv(p));
                EXPECT(p.Score == local[k].score && PrintCigar(p.Cigar_) == local[k].cig);
            }
        eng->inputs->close();
        align::TargetQueryPair p;
        EXPECT(!eng->outputs->recv(p));
    }
    { // SURVEY.md appendix A known answers
        auto r = align::AffineGap_highMem(dna::StringToBases("CGCGCGCGCG"), dna::StringToBases("CGAAAACGCGTTTTCGCG"),
                                          align::DefaultScoreMatrix, -400, -30);
        EXPECT(r.first == -40 && PrintCigar(r.second) == "2M4I4M4I4M");
        r = align::AffineGap(dna::StringToBases("TTGTTCGGG"), dna::StringToBases("TTGTTATTCAAAGGG"),
                             align::HumanChimpTwoScoreMatrix, -600, -150);
        EXPECT(r.first == -1070 && PrintCigar(r.second) == "5M6I4M");
        r = align::ConstGap(dna::StringToBases("TTGTTATTC"), dna::StringToBases("TTGTTC"), align::HumanChimpTwoScoreMatrix, -430);
        EXPECT(r.first == -730 && PrintCigar(r.second) == "3M3D3M");
        r = align::AffineGapChunk(dna::StringToBases("TTGTTCTTCTTCTTC"), dna::StringToBases("TTGTTCTTCTTATTATTATTCTTC"),
                                  align::DefaultScoreMatrix, -400, -30, 3);
        EXPECT(PrintCigar(r.second) == "9M9I6M");
    }
    { // error behaviour: lowercase base indexes past the matrix -> Go panics
        bool threw = false;
        try {
            align::AffineGap_highMem(dna::StringToBases("acg"), dna::StringToBases("ACG"), align::DefaultScoreMatrix, -400, -30);
        } catch (const std::out_of_range &) {
            threw = true;
        }
        EXPECT(threw);
    }
    std::printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
    return fails ? 1 : 0;
}

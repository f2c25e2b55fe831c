#!/usr/bin/env python
"""Extract the reference's golden vectors for the alignment hot path into JSON fixtures.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It only READS data and expectation tables from the reference's tests/testdata (no code is
copied, nothing is executed -- the reference is Go and there is no Go toolchain here) and writes
tests/golden/*.json, which are committed so the GPU box (no /root/reference) can replay them.
Sources (relative to /root/reference):
  align/affineGap_test.go:11-36,120-155   affine global / chunk tables, AffineGapLocal score+cigar
  align/view_test.go:9-24                 const-gap table
  align/testdata/multiAlignTest.*.fa      progressive MSA fixtures
  cmd/globalAlignmentAnchor/testdata      out_alignment.{1,2}.expected.tsv + toy genomes
  cmd/cigarToBed/testdata                 seth/raven, PanTro6/hg38 10 kb pair + ins/del BEDs
  cmd/globalAlignment/testdata            chelsea/eric + faOut_test.fa
  dna/dnaTwoBit/perfectAlign_test.go:21-92 CountLeft/RightMatches known answers (seedTestsShort/Long)
  dna/dnaTwoBit/dnaTwoBit_test.go:9-42    GetBase expectations on four packed strings
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def read_fasta(p):
    recs, name, seq = [], None, []
    for line in read(p).splitlines():
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = line[1:].strip(), []
        elif line.strip():
            seq.append(line.strip())
    if name is not None:
        recs.append((name, "".join(seq)))
    return recs


def go_table(src, var):
    """rows of string literals from `var <name> = []struct{...}{ {"a","b","c"}, ... }`"""
    m = re.search(r"var\s+" + var + r"\s*=\s*\[\]struct\s*\{.*?\}\s*\{(.*?)\n\}", src, re.S)
    rows = []
    for row in re.findall(r"\{((?:\s*\"(?:[^\"\\]|\\.)*\"\s*,?)+)\}", m.group(1)):
        rows.append([bytes(s, "utf-8").decode("unicode_escape") for s in re.findall(r"\"((?:[^\"\\]|\\.)*)\"", row)])
    return rows


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, indent=1)
        f.write("\n")
    print("wrote", name)


def main():
    aff = read("align/affineGap_test.go")
    dump("affine_global.json", {
        "source": "align/affineGap_test.go:11-25 (TestAffineGap, TestAffineGap_lowMem, TestAffineGapMulti)",
        "matrix": "Default", "gap_open": -400, "gap_extend": -30,
        "cases": [{"alpha": a, "beta": b, "view": v} for a, b, v in go_table(aff, "affineAlignTests")]})
    dump("affine_chunk.json", {
        "source": "align/affineGap_test.go:27-36 (TestAffineGapChunk)",
        "matrix": "Default", "gap_open": -400, "gap_extend": -30, "chunk": 3,
        "cases": [{"alpha": a, "beta": b, "view": v} for a, b, v in go_table(aff, "affineAlignChunkTests")]})

    # TestAffineGapLocal: tgt/qry literals followed by the call and the expected score/cigar
    body = aff[aff.index("func TestAffineGapLocal"):aff.index("func TestGoAffineGapLocalEngine")]
    local = []
    pat = re.compile(r"tgt\s*:?=\s*dna\.StringToBases\(\"(\w+)\"\)\s*qry\s*:?=\s*dna\.StringToBases\(\"(\w+)\"\)\s*"
                     r"score,\s*cig\s*:?=\s*AffineGapLocal\(tgt,\s*qry,\s*(\w+),\s*(-?\d+),\s*(-?\d+)\)\s*"
                     r"if\s+score\s*!=\s*(-?\d+)\s*\|\|\s*PrintCigar\(cig\)\s*!=\s*\"(\w+)\"", re.S)
    for t, q, mat, o, e, sc, cg in pat.findall(body):
        local.append({"target": t, "query": q, "matrix": mat.replace("ScoreMatrix", ""), "gap_open": int(o),
                      "gap_extend": int(e), "score": int(sc), "cigar": cg})
    assert len(local) == 5, local
    dump("affine_local.json", {"source": "align/affineGap_test.go:120-155 (TestAffineGapLocal; first 4 also "
                                         "TestGoAffineGapLocalEngine :157-192)", "cases": local})

    dump("const_gap.json", {
        "source": "align/view_test.go:9-38 (TestConstGap)", "matrix": "Default", "gap_pen": -430,
        "cases": [{"alpha": a, "beta": b, "view": v} for a, b, v in go_table(read("align/view_test.go"), "alignTests")]})

    dump("multi_align.json", {
        "source": "align/multiAlign_test.go:9-37 (TestMultiAlignGap): AllSeqAffine and AllSeqAffineChunk(chunk=2), "
                  "Default,-400,-30, equal ignoring order",
        "matrix": "Default", "gap_open": -400, "gap_extend": -30, "chunk": 2,
        "cases": [{"input": read_fasta("align/testdata/multiAlignTest.in.fa"),
                   "expected": read_fasta("align/testdata/multiAlignTest.expected.fa")},
                  {"input": read_fasta("align/testdata/multiAlignTest.in2.fa"),
                   "expected": read_fasta("align/testdata/multiAlignTest.expected2.fa")}]})

    # globalAlignmentAnchor: rows whose both names end in _gap were produced by
    # AffineGap_customizeCheckersize(seq1[start-1:end-1], seq2[...], HumanChimpTwo, -600, -150, 10000, 10000)
    # after dna.AllToUpper (cmd/globalAlignmentAnchor/globalAlignmentAnchor.go:378-384)
    d = "cmd/globalAlignmentAnchor/testdata/"
    g1 = dict(read_fasta(d + "hg38.toy.fa"))
    g2 = dict(read_fasta(d + "rheMac10.toy.fa"))
    rows, seen = [], set()
    for tsv in ("out_alignment.1.expected.tsv", "out_alignment.2.expected.tsv"):
        for line in read(d + tsv).splitlines():
            c = line.split("\t")
            if c[3] != "species1_gap" or c[7] != "species2_gap":
                continue
            key = tuple(c[:8])
            if key in seen:
                continue
            seen.add(key)
            s1 = g1[c[0]][int(c[1]) - 1:int(c[2]) - 1]
            s2 = g2[c[4]][int(c[5]) - 1:int(c[6]) - 1]
            cig = [[int(a), int(b)] for a, b in re.findall(r"\{(\d+) (\d+)\}", c[9])]
            rows.append({"file": tsv, "region1": c[:3], "region2": c[4:7], "alpha": s1, "beta": s2,
                         "score": int(c[8]), "cigar": cig})
    dump("anchor.json", {"source": d + "out_alignment.{1,2}.expected.tsv (TestGlobalAlignmentAnchorTests)",
                         "note": "sequences are soft-masked as in the toy genomes; upper-case before aligning",
                         "matrix": "HumanChimpTwo", "gap_open": -600, "gap_extend": -150, "cases": rows})

    # cigarToBed: AffineGap(upper(faOne), upper(faTwo), HumanChimpTwo, -600, -150), BEDs derived from the cigar
    d = "cmd/cigarToBed/testdata/"
    cases = []
    for one, two, fi, fd, chrom, ins, dele in (
            ("sethvsraven/seth.fa", "sethvsraven/raven.fa", 1, 1, "chr1",
             "sethvsraven/affineGap_sethvsraven_ins.bed", "sethvsraven/affineGap_sethvsraven_del.bed"),
            ("firstTest/testRegion10kb_PanTro6.fa", "firstTest/testRegion10kb_hg38.fa", 119320000, 116703287, "chr1",
             "firstTest/affineGap_PanTro6vshg38_ins.bed", "firstTest/affineGap_PanTro6vshg38_del.bed")):
        cases.append({"alpha": read_fasta(d + one)[0][1], "beta": read_fasta(d + two)[0][1],
                      "first_pos_ins": fi, "first_pos_del": fd, "chrom": chrom,
                      "ins_bed": read(d + ins), "del_bed": read(d + dele), "files": [one, two]})
    dump("cigar_to_bed.json", {"source": "cmd/cigarToBed/cigarToBed_test.go:10-21 + cigarToBed.go:86-129",
                               "matrix": "HumanChimpTwo", "gap_open": -600, "gap_extend": -150, "cases": cases})

    d = "cmd/globalAlignment/testdata/"
    out = read_fasta(d + "faOut_test.fa")
    dump("global_alignment.json", {
        "source": "cmd/globalAlignment/globalAlignment_test.go + globalAlignment.go:84 "
                  "(ConstGap(chelsea, eric, HumanChimpTwo, -430); faOut_test.fa is the View)",
        "matrix": "HumanChimpTwo", "gap_pen": -430,
        "alpha": read_fasta(d + "chelsea.fa")[0][1], "beta": read_fasta(d + "eric.fa")[0][1],
        "view": out[0][1] + "\n" + out[1][1] + "\n"})

    twobit()


def twobit():
    """dna/dnaTwoBit known answers: the seedTest tables and the GetBase checks."""
    src = read("dna/dnaTwoBit/perfectAlign_test.go")
    seqs = dict(re.findall(r"var\s+(\w+)\s+\[\]dna\.Base\s*=\s*dna\.StringToBases\(\"(\w+)\"\)", src))
    cases = []
    for a, b, sa, sb, left, right in re.findall(
            r"\{SeqA:\s*(\w+),.*?SeqB:\s*(\w+),.*?StartA:\s*(\d+),\s*StartB:\s*(\d+),\s*"
            r"TrueMatchesLeft:\s*(\d+),\s*TrueMatchesRight:\s*(\d+),", src, re.S):
        cases.append({"seq_a": seqs[a], "seq_b": seqs[b], "start_a": int(sa), "start_b": int(sb),
                      "left": int(left), "right": int(right), "names": [a, b]})
    assert len(cases) == 4, cases
    t = read("dna/dnaTwoBit/dnaTwoBit_test.go")
    strings = re.findall(r"\"([ACGT]+)\",", t[t.index("var dnaStrings"):t.index("func TestDnaToFromString")])
    checks = [[int(pos), base] for pos, base in
              re.findall(r"GetBase\(frag,\s*(\d+)\)\s*if\s+singleBase\s*!=\s*dna\.(\w)\b", t)]
    assert len(strings) == 4 and len(checks) == 5, (strings, checks)
    dump("twobit.json", {
        "source": "dna/dnaTwoBit/perfectAlign_test.go:21-92 (TestCounting) + dnaTwoBit_test.go:9-42 (TestDnaToFromString)",
        "count_cases": cases, "get_base_strings": strings, "get_base_checks": checks})


if __name__ == "__main__":
    main()

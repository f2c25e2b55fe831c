"""Debug driver: run extend-step cases one by one (each under its own timeout) and report mismatches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as orc
from golden_util import bases, random_pair
from gonomics_b200 import align, genomegraph as gg

which = sys.argv[1]
rng = np.random.default_rng(2341)
S = orc.HUMAN_CHIMP_TWO_SCORE_MATRIX
alphas, betas = [], []
if which == "rep":
    for unit in ("A", "AC", "ACG", "AAC"):
        for n, m in ((30, 30), (64, 17), (17, 64), (150, 150)):
            alphas.append(bases((unit * 200)[:n])); betas.append(bases((unit * 200)[1:m + 1]))
elif which == "n":
    for _ in range(40):
        a, b = random_pair(rng, int(rng.integers(1, 170)), int(rng.integers(1, 150)), alphabet=5)
        alphas.append(a); betas.append(b)
else:
    n, m = map(int, which.split("x"))
    a, b = random_pair(rng, n, m, identity=0.92)
    alphas.append(a); betas.append(b)
ctx = align.Context(0)
for side, ofn in ((1, orc.left_dynamic_aln), (2, orc.right_dynamic_aln)):
    got = gg.extend_pairs(side, alphas, betas, S, -600, ctx)
    bad = 0
    for p, (a, b) in enumerate(zip(alphas, betas)):
        want = ofn(a, b, S, -600)
        g = (got[p][0], [tuple(c) for c in got[p][1]], got[p][2], got[p][3])
        if g != want:
            bad += 1
            if bad <= 3:
                print("MISMATCH side", side, "pair", p, len(a), len(b), "\n  got ", g, "\n  want", want)
    print(which, "side", side, "bad", bad, "of", len(alphas))

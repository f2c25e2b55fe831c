"""TEST INFRASTRUCTURE (oracle): plain-Python restatements of the linear-gap extension DPs of gonomics' genomeGraph
package, written from the Go source statement by statement.  Slow (Python loops): for small cases only.

    LeftDynamicAln / RightDynamicAln   genomeGraph/search.go:234-321   (cigar.TripleMaxTrace, route in traceback order)
    LeftLocal / RightLocal             genomeGraph/localAlignment.go:95-196 (cigar.TripleMaxTraceExtended, route reversed)

left_dynamic_aln / right_dynamic_aln here are a SECOND, independent restatement of what oracle/gnx_oracle.c's
orc_left_dynamic_aln / orc_right_dynamic_aln implement in C (tests/test_gsw_oracle.py holds them against each other): the
reference has no asserting test for these functions, so two independent readings of the source are the pin.
"""
from __future__ import annotations

from typing import List, Tuple


def triple_max_trace(a: int, b: int, c: int) -> Tuple[int, str]:
    """cigar.TripleMaxTrace (cigar/tools.go:58-66)."""
    if a >= b and a >= c:
        return a, "M"
    if b >= c:
        return b, "I"
    return c, "D"


def triple_max_trace_extended(prev: int, a: int, b: int, c: int) -> Tuple[int, str]:
    """cigar.TripleMaxTraceExtended (cigar/tools.go:69-81)."""
    if a >= b and a >= c:
        return (a, "=") if a > prev else (a, "X")
    if b >= c:
        return b, "I"
    return c, "D"


def _walk(trace, i, j, cond, route_start_empty=True):
    route: List[List] = []
    idx = 0
    while cond(i, j):
        op = trace[i][j]
        if len(route) == 0:
            route.append([1, op])
        elif route[idx][1] == op:
            route[idx][0] += 1
        else:
            route.append([1, op])
            idx += 1
        if op in ("M", "=", "X"):
            i, j = i - 1, j - 1
        elif op == "I":
            j -= 1
        elif op == "D":
            i -= 1
        else:
            raise AssertionError("unexpected traceback")
    return [(r, o) for r, o in route], i, j


def left_dynamic_aln(alpha, beta, scores, gap_pen: int):
    """search.go:234-274: zero boundaries, cells clipped at 0 AFTER their trace is recorded, walk from (n,m) while > 0."""
    n, m = len(alpha), len(beta)
    M = [[0] * (m + 1) for _ in range(n + 1)]
    T = [[None] * (m + 1) for _ in range(n + 1)]
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            M[i][j], T[i][j] = triple_max_trace(M[i - 1][j - 1] + int(scores[alpha[i - 1]][beta[j - 1]]), M[i][j - 1] + gap_pen,
                                                M[i - 1][j] + gap_pen)
            if M[i][j] < 0:
                M[i][j] = 0
    route, i, j = _walk(T, n, m, lambda i, j: M[i][j] > 0)
    return M[n][m], route, i, j


def right_dynamic_aln(alpha, beta, scores, gap_pen: int):
    """search.go:276-321: Needleman-Wunsch boundaries, first strict maximum in row-major order, walk to (0,0)."""
    n, m = len(alpha), len(beta)
    M = [[0] * (m + 1) for _ in range(n + 1)]
    T = [[None] * (m + 1) for _ in range(n + 1)]
    curr_max, max_i, max_j = 0, 0, 0
    for i in range(n + 1):
        for j in range(m + 1):
            if i == 0 and j == 0:
                M[i][j] = 0
            elif i == 0:
                M[i][j] = M[i][j - 1] + gap_pen
                T[i][j] = "I"
            elif j == 0:
                M[i][j] = M[i - 1][j] + gap_pen
                T[i][j] = "D"
            else:
                M[i][j], T[i][j] = triple_max_trace(M[i - 1][j - 1] + int(scores[alpha[i - 1]][beta[j - 1]]), M[i][j - 1] + gap_pen,
                                                    M[i - 1][j] + gap_pen)
            if M[i][j] > curr_max:
                curr_max, max_i, max_j = M[i][j], i, j
    route, _, _ = _walk(T, max_i, max_j, lambda i, j: i > 0 or j > 0)
    return M[max_i][max_j], route, max_i, max_j


def left_local(alpha, beta, scores, gap_pen: int):
    """localAlignment.go:95-140: (score, route, minI, maxI, minJ, maxJ), route reversed into alignment order."""
    n, m = len(alpha), len(beta)
    M = [[0] * (m + 1) for _ in range(n + 1)]
    T = [[None] * (m + 1) for _ in range(n + 1)]
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            M[i][j], T[i][j] = triple_max_trace_extended(M[i - 1][j - 1], M[i - 1][j - 1] + int(scores[alpha[i - 1]][beta[j - 1]]),
                                                         M[i][j - 1] + gap_pen, M[i - 1][j] + gap_pen)
            if M[i][j] < 0:
                M[i][j] = 0
    route, min_i, min_j = _walk(T, n, m, lambda i, j: M[i][j] > 0)  # minI, minJ start at len(alpha), len(beta)
    return M[n][m], route[::-1], min_i, n, min_j, m


def right_local(alpha, beta, scores, gap_pen: int):
    """localAlignment.go:142-196."""
    n, m = len(alpha), len(beta)
    M = [[0] * (m + 1) for _ in range(n + 1)]
    T = [[None] * (m + 1) for _ in range(n + 1)]
    curr_max, max_i, max_j = 0, 0, 0
    for i in range(n + 1):
        for j in range(m + 1):
            if i == 0 and j == 0:
                M[i][j] = 0
            elif i == 0:
                M[i][j] = M[i][j - 1] + gap_pen
                T[i][j] = "I"
            elif j == 0:
                M[i][j] = M[i - 1][j] + gap_pen
                T[i][j] = "D"
            else:
                M[i][j], T[i][j] = triple_max_trace_extended(M[i - 1][j - 1], M[i - 1][j - 1] + int(scores[alpha[i - 1]][beta[j - 1]]),
                                                             M[i][j - 1] + gap_pen, M[i - 1][j] + gap_pen)
            if M[i][j] > curr_max:
                curr_max, max_i, max_j = M[i][j], i, j
    route, _, _ = _walk(T, max_i, max_j, lambda i, j: i > 0 or j > 0)
    return M[max_i][max_j], route[::-1], 0, max_i, 0, max_j

"""Sharding of an alignment batch over ranks (one process per GPU) and the final gather.

Pairs are independent, so the only data-path exchange is the gather of per-shard results
(north_star: "NCCL over NVLink used only for the final score/CIGAR gather").  Works with any
torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_bounds(alpha_off: np.ndarray, beta_off: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous pair ranges [lo, hi) per rank, balanced by DP cells (sum of n*m)."""
    n_pairs = len(alpha_off) - 1
    cells = np.diff(alpha_off).astype(np.int64) * np.diff(beta_off).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(cells)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cuts.append(int(np.searchsorted(csum, target, side="left")))
    cuts.append(n_pairs)
    for r in range(1, len(cuts)):  # monotone, never past the end
        cuts[r] = min(max(cuts[r], cuts[r - 1]), n_pairs)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slice_batch(alpha_cat, alpha_off, beta_cat, beta_off, lo: int, hi: int):
    """The sub-batch [lo, hi) with offsets rebased to zero (views, no copies of the bases)."""
    ao = alpha_off[lo:hi + 1] - alpha_off[lo]
    bo = beta_off[lo:hi + 1] - beta_off[lo]
    return (alpha_cat[alpha_off[lo]:alpha_off[hi]], ao.astype(np.int64),
            beta_cat[beta_off[lo]:beta_off[hi]], bo.astype(np.int64))


def gather_results(scores: np.ndarray, cigar_off: Optional[np.ndarray], cigars: Optional[np.ndarray], device=None):
    """all_gather the shards' (scores, cigar_off, cigars) into the global batch order on every rank.

    Two fixed-width collectives (counts, then payloads padded to the largest shard), as in SURVEY.md 8e."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = device or torch.device("cpu")

    def ag(t: "torch.Tensor") -> List["torch.Tensor"]:
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return out

    want_cigar = cigar_off is not None
    n_ops = int(cigar_off[-1]) if want_cigar else 0
    meta = torch.tensor([len(scores), n_ops], dtype=torch.int64, device=dev)
    metas = [m.cpu().numpy() for m in ag(meta)]
    max_pairs = max(int(m[0]) for m in metas)
    max_ops = max(int(m[1]) for m in metas)

    def padded(a: np.ndarray, n: int, dtype) -> "torch.Tensor":
        buf = np.zeros(n, dtype=dtype)
        buf[:len(a)] = a
        return torch.from_numpy(buf).to(dev)

    all_scores = ag(padded(scores, max_pairs, np.int64))
    out_scores = np.concatenate([t.cpu().numpy()[:int(m[0])] for t, m in zip(all_scores, metas)])
    if not want_cigar:
        return out_scores, None, None
    all_cnt = ag(padded(np.diff(cigar_off), max_pairs, np.int64))
    raw = np.ascontiguousarray(cigars).view(np.uint8)
    all_cig = ag(padded(raw, max_ops * cigars.dtype.itemsize, np.uint8))
    counts = np.concatenate([t.cpu().numpy()[:int(m[0])] for t, m in zip(all_cnt, metas)])
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    cig = np.concatenate([t.cpu().numpy()[:int(m[1]) * cigars.dtype.itemsize].view(cigars.dtype)
                          for t, m in zip(all_cig, metas)])
    return out_scores, off, cig

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


def gather_device(scores, cigar_off, cigars_u8, item_bytes: int = 16):
    """The path's one exchange, on device tensors (NCCL on the GPU box, gloo in the CPU tests): every rank ends up with
    every shard's scores, per-pair cigar counts and cigar records.  Shards may differ in pair count and cigar volume, so
    the payloads are padded to the largest shard (two small collectives size them, three move the data).

    scores: int64[P]; cigar_off: int64[P+1] (None: scores only); cigars_u8: uint8 view of the shard's gnx_cigar records.
    Returns (scores[world, Pmax], counts[world, Pmax] | None, cigars_u8[world, Tmax * item_bytes] | None,
             meta[world, 2] = (pairs, cigar elements) per rank, bytes moved per rank)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = scores.device
    P = scores.numel()
    n_ops = cigar_off[-1:].clone() if cigar_off is not None else torch.zeros(1, dtype=torch.int64, device=dev)
    meta = torch.cat([torch.tensor([P], dtype=torch.int64, device=dev), n_ops.to(torch.int64)])
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.view(world, 2)
    mx = metas.max(dim=0).values.cpu()  # the one host synchronisation: payload sizes
    pmax, tmax = int(mx[0]), int(mx[1])

    def padded(t, n):
        if t.numel() == n:
            return t.contiguous()
        out = torch.zeros(n, dtype=t.dtype, device=dev)
        out[:t.numel()] = t
        return out

    all_scores = torch.empty(world * pmax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_scores, padded(scores, pmax))
    moved = world * pmax * 8
    if cigar_off is None:
        return all_scores.view(world, pmax), None, None, metas, moved
    counts = (cigar_off[1:] - cigar_off[:-1]).to(torch.int32)
    all_counts = torch.empty(world * pmax, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_counts, padded(counts, pmax))
    nbytes = tmax * item_bytes
    all_cig = torch.empty(world * max(nbytes, 1), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_cig, padded(cigars_u8[:min(cigars_u8.numel(), nbytes)], max(nbytes, 1)))
    moved += world * pmax * 4 + world * nbytes
    return all_scores.view(world, pmax), all_counts.view(world, pmax), all_cig.view(world, max(nbytes, 1)), metas, moved


def compact_gathered(all_scores, all_counts, all_cig, metas, item_dtype):
    """Host-side view of gather_device's padded result in global pair order: (scores, cigar_off, cigars)."""
    metas = metas.cpu().numpy()
    sc = np.concatenate([all_scores[r, :int(metas[r, 0])].cpu().numpy() for r in range(len(metas))])
    if all_counts is None:
        return sc, None, None
    cnt = np.concatenate([all_counts[r, :int(metas[r, 0])].cpu().numpy() for r in range(len(metas))]).astype(np.int64)
    off = np.zeros(len(cnt) + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    isz = np.dtype(item_dtype).itemsize
    cig = np.concatenate([all_cig[r, :int(metas[r, 1]) * isz].cpu().numpy().view(item_dtype) for r in range(len(metas))])
    return sc, off, cig

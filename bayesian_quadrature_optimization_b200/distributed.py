"""Multi-GPU plumbing (one process per GPU, torch.distributed: NCCL on GPUs, gloo in CPU tests).

The hot path shards naturally: (candidate, hyper-sample) units are
independent given the training data, which every GPU holds in full.  Two partitions are used:

  * candidates across ranks (default; all S factors fit in HBM on every GPU): no data-path
    collective at all -- each rank scores its own candidates and ONE all_gather of
    (best value, global index, point) picks the winner;
  * hyper-samples across ranks (when S factors of size n^2 do not fit on one GPU): each rank
    produces per-candidate partial sums over its own hyper-samples and ONE all_gather followed by
    a fixed-order sum gives every rank the same bit-identical totals (an all_reduce would
    change the last ulp with the ring order and could flip an argmax tie).

The replaced reference mechanism is the fork/pickle process pool of sbo/lib/parallel.py:16-76.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allgather_best(value: torch.Tensor, global_index: torch.Tensor, point: torch.Tensor):
    """Every rank passes its local best (scalar value, scalar global index, point[d]); all ranks
    return the same global best.  Ties go to the smallest global index (np.argmax semantics over
    the concatenated candidate list, sbo/acquisition_functions/sbo.py:1896-1900)."""
    rank, ws = world()
    rec = torch.cat([value.reshape(1).to(torch.float64),
                     global_index.reshape(1).to(torch.float64),
                     point.reshape(-1).to(torch.float64)])
    if ws == 1:
        allr = rec.reshape(1, -1)
    else:
        bufs = [torch.empty_like(rec) for _ in range(ws)]
        dist.all_gather(bufs, rec)
        allr = torch.stack(bufs)
    vals = allr[:, 0].cpu().numpy()
    idxs = allr[:, 1].cpu().numpy()
    order = np.lexsort((idxs, -vals))  # value descending, then index ascending; NaN last
    finite = [r for r in order if not np.isnan(vals[r])]
    win = finite[0] if finite else int(order[0])
    return float(vals[win]), int(idxs[win]), allr[win, 2:].cpu().numpy()


def allgather_ordered_sum(partial: torch.Tensor):
    """Sum of per-rank partial sums in rank order (bit-reproducible on every rank)."""
    rank, ws = world()
    if ws == 1:
        return partial.clone()
    bufs = [torch.empty_like(partial) for _ in range(ws)]
    dist.all_gather(bufs, partial.contiguous())
    total = bufs[0].clone()
    for r in range(1, ws):
        total += bufs[r]
    return total


def evaluate_candidates_hyper_sharded(make_bq, thetas, cands, z, starts, max_mean=None, **kwargs):
    """Hyper-samples across ranks (the partition that bounds memory, one n x n factor
    per hyper-sample): rank r factorises and scores only its slice of `thetas`; the per-candidate
    sums over hyper-samples are combined with ONE all_gather and a rank-ordered sum, so every rank
    holds bit-identical totals and the argmax over candidates agrees everywhere.

    :param make_bq: callable(thetas_slice) -> BQEngine for those hyper-samples on this rank's GPU
    :return: {'evaluations': tensor[C], 'gradient': tensor[C][dim_m] or None, 'argmax': int}
    """
    rank, ws = world()
    thetas = np.asarray(thetas, dtype=float)
    S = thetas.shape[0]
    lo, hi = shard_range(S, rank, ws)
    if hi > lo:
        bq = make_bq(thetas[lo:hi])
        mm = None if max_mean is None else np.asarray(max_mean, dtype=float)[lo:hi]
        res = bq.evaluate_candidates(cands, z, starts, max_mean=mm, **kwargs)
        part_v = res['evaluations'] * float(hi - lo)
        part_g = None if res.get('gradient') is None else res['gradient'] * float(hi - lo)
        dev = part_v.device
    else:  # more ranks than hyper-samples: this rank contributes zeros
        raise ValueError("hyper-sample sharding needs at least one hyper-sample per rank")
    tot_v = allgather_ordered_sum(part_v) / float(S)
    tot_g = None if part_g is None else allgather_ordered_sum(part_g) / float(S)
    return {'evaluations': tot_v, 'gradient': tot_g, 'argmax': int(torch.argmax(tot_v).item())}

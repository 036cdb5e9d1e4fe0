"""Candidate sharding for acquisition scoring over the GPUs of one box.

The candidates of one acquisition pass (reference bayes_opt/helpers.py:51-68) are independent, so
rank r of R scores the contiguous range [lo_r, hi_r) of the M candidate rows and keeps a per-GPU
(best value, lowest global index).  The only exchange on the data path is an all-gather of those
R x 16 bytes (NCCL over NVLink on the GPUs; any torch.distributed backend works, the CPU tests use
gloo) followed by the deterministic merge "max value, then min index", which is numpy.argmax's
first-occurrence rule (helpers.py:67).  The fitted state (W = L^-1, V, scaled theta, kvec,
hyper-parameters) is broadcast from the rank that ran the fit.

This module holds host logic only (no arithmetic on candidates); it never imports the oracle.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

NEG_INF = float("-inf")
IDX_SENTINEL = np.iinfo(np.int64).max


def shard_range(M, rank, world):
    """Contiguous split of M candidate rows: rank r gets [lo, hi).  Sizes differ by at most one row
    block of 128 (the scoring tile), so that every shard except the last is tile aligned."""
    if world <= 1:
        return 0, int(M)
    tiles = (int(M) + 127) // 128
    base, rem = divmod(tiles, world)
    lo_t = rank * base + min(rank, rem)
    hi_t = lo_t + base + (1 if rank < rem else 0)
    return min(lo_t * 128, int(M)), min(hi_t * 128, int(M))


def merge_best(vals, idxs):
    """(max value, then min index) over per-shard results; empty shards carry (-inf, IDX_SENTINEL).
    NaN values never win (they fail every comparison), as in the device merge."""
    best_v, best_i = NEG_INF, IDX_SENTINEL
    for v, i in zip(np.asarray(vals, dtype=np.float64).tolist(), np.asarray(idxs, dtype=np.int64).tolist()):
        if v > best_v or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_v, int(best_i)


def pack_best(val, idx, device=None):
    """One 16-byte record [value bits, index] as int64[2] (bit cast: the collective only moves bytes)."""
    rec = torch.empty(2, dtype=torch.int64, device=device if device is not None else val.device)
    rec[0:1].view(torch.float64).copy_(val.reshape(1).to(torch.float64))
    rec[1:2].copy_(idx.reshape(1).to(torch.int64))
    return rec


def unpack_best(recs):
    recs = recs.reshape(-1, 2)
    vals = recs[:, 0].contiguous().view(torch.float64)
    return vals, recs[:, 1]


class ShardGroup:
    """The ranks that share one acquisition pass.  world == 1 (no process group) is the identity."""

    def __init__(self, group=None):
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    def range(self, M):
        return shard_range(M, self.rank, self.world)

    def broadcast_fit(self, tensors, src=0):
        """Broadcast the fitted state (a list of device tensors, same shapes on every rank) from
        `src`.  One collective per tensor: W is N^2*8 bytes (537 MB at N=8192), the rest is small."""
        if self.world == 1:
            return
        for t in tensors:
            dist.broadcast(t, src=src, group=self.group)

    def all_gather_best(self, rec):
        """rec: int64[2] from pack_best -> int64[world, 2] on every rank (asynchronous on the
        current stream for NCCL)."""
        if self.world == 1:
            return rec.reshape(1, 2)
        out = torch.empty(self.world * 2, dtype=torch.int64, device=rec.device)
        dist.all_gather_into_tensor(out, rec, group=self.group)
        return out.reshape(self.world, 2)

    def argmax(self, local_val, local_idx):
        """local (value, global index) tensors -> (best value, best global index) as python numbers,
        identical on every rank."""
        recs = self.all_gather_best(pack_best(local_val, local_idx))
        vals, idxs = unpack_best(recs)
        return merge_best(vals.cpu().numpy(), idxs.cpu().numpy())

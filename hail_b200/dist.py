"""Variant sharding across the GPUs of one box.

The reference scales this path exactly one way: independent row partitions, one Spark task each
(`mv.rvd.mapPartitionsWithContext`, LinearRegression.scala:95 / :274, `preservesPartitionCounts = true`), with the
driver's prologue shipped to the tasks by `sc.broadcast` (LR:74-78, LR:257).  Here: one process per GPU, contiguous
variant ranges per rank, the group bases broadcast from rank 0 and the fixed-width result rows all-gathered in
rank order (= row-key order).  There is no exchange inside the sweep.

`torch.distributed` is the transport (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

BASIS_FIELDS = ("complete_idx", "q_cols", "y_res", "qty", "yyp")


def variant_range(rank: int, world: int, n_variants: int):
    """Contiguous range [lo, hi) of rank `rank`: concatenating the ranks in order reproduces the row order."""
    lo = rank * n_variants // world
    hi = (rank + 1) * n_variants // world
    return lo, hi


class BasisTensors:
    """A GroupBasis flattened to tensors (what travels in the broadcast)."""

    def __init__(self, meta, tensors):
        self.n, self.K, self.P, self.has_intercept = (int(v) for v in meta)
        self.tensors = tensors  # complete_idx int32 [n]; q_cols f64 [Kd, n]; y_res [P, n]; qty [K, P]; yyp [P]

    @classmethod
    def from_group_basis(cls, b, device):
        t = [torch.from_numpy(np.ascontiguousarray(getattr(b, f))).to(device) for f in BASIS_FIELDS]
        return cls((b.n, b.K, b.P, int(b.has_intercept)), t)

    @property
    def d(self):
        return self.n - self.K - 1


def broadcast_bases(bases, device, src=0):
    """Rank `src` passes its list of GroupBasis (others pass None); every rank returns the list of BasisTensors."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [BasisTensors.from_group_basis(b, device) for b in bases]
    count = torch.tensor([len(bases) if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(count, src)
    out = []
    for g in range(int(count.item())):
        if rank == src:
            bt = BasisTensors.from_group_basis(bases[g], device)
            meta = torch.tensor([bt.n, bt.K, bt.P, bt.has_intercept], dtype=torch.int64, device=device)
        else:
            meta = torch.zeros(4, dtype=torch.int64, device=device)
        dist.broadcast(meta, src)
        n, K, P, hi = (int(v) for v in meta.tolist())
        if rank != src:
            bt = BasisTensors((n, K, P, hi), [
                torch.empty(n, dtype=torch.int32, device=device),
                torch.empty((K - hi, n), dtype=torch.float64, device=device),
                torch.empty((P, n), dtype=torch.float64, device=device),
                torch.empty((K, P), dtype=torch.float64, device=device),
                torch.empty(P, dtype=torch.float64, device=device),
            ])
        for x in bt.tensors:
            if x.numel():
                dist.broadcast(x, src)
        out.append(bt)
    return out


def gather_rows(local_rows: torch.Tensor, counts=None, out=None, group=None):
    """All-gather row blocks [m_r, W] (m_r may differ per rank) and concatenate them in rank order.

    Equal blocks (the sharded sweep's normal case) go straight into one [world * m, W] tensor (`out`, allocated when
    None) with all_gather_into_tensor: no per-rank temporaries, no concatenation pass."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world = dist.get_world_size()
    dev = local_rows.device
    if counts is not None and len(set(counts)) == 1 and counts[0] == local_rows.shape[0]:
        if out is None:
            out = torch.empty((world * local_rows.shape[0], local_rows.shape[1]), dtype=local_rows.dtype, device=dev)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    if counts is None:
        c = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=dev)
        cs = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(cs, c)
        counts = [int(x.item()) for x in cs]
    width = local_rows.shape[1]
    mmax = max(counts)
    padded = local_rows
    if local_rows.shape[0] < mmax:
        padded = torch.zeros((mmax, width), dtype=local_rows.dtype, device=dev)
        padded[: local_rows.shape[0]] = local_rows
    parts = [torch.empty((mmax, width), dtype=local_rows.dtype, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)

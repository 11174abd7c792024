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
        if getattr(b, "weighted", False):
            # a weighted group also carries sqrt(w) and must reach lrr_add_group_weighted: not part of this message
            raise NotImplementedError("BasisTensors: weighted groups (weights=) are not broadcast; run them on one device")
        t = [torch.from_numpy(np.ascontiguousarray(getattr(b, f))).to(device) for f in BASIS_FIELDS]
        return cls((b.n, b.K, b.P, int(b.has_intercept)), t)

    @property
    def d(self):
        return self.n - self.K - 1


MAX_GROUPS_MSG = 63   # groups one broadcast header can describe


def _basis_shapes(n, K, P, hi):
    """(dtype, shape) of BASIS_FIELDS for a group with these counts."""
    return [(torch.int32, (n,)), (torch.float64, (K - hi, n)), (torch.float64, (P, n)), (torch.float64, (K, P)),
            (torch.float64, (P,))]


def broadcast_bases(bases, device, src=0, group=None, message=None, payload_transport=None):
    """Rank `src` passes its list of GroupBasis (others pass None); every rank returns the list of BasisTensors.

    What `sc.broadcast` ships per call (LR:74-78, LR:257) travels as ONE flat message: a fixed-size header (group count and
    the four counts of every group) and one byte payload holding every array of every group; the receivers' tensors
    are views into that payload."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [BasisTensors.from_group_basis(b, device) for b in bases]
    header = torch.zeros(1 + 4 * MAX_GROUPS_MSG, dtype=torch.int64, device=device)
    mine = None
    if rank == src:
        if len(bases) > MAX_GROUPS_MSG:
            raise ValueError(f"broadcast_bases: at most {MAX_GROUPS_MSG} groups per call")
        mine = [BasisTensors.from_group_basis(b, device) for b in bases]
        header[0] = len(mine)
        for g, bt in enumerate(mine):
            header[1 + 4 * g: 5 + 4 * g] = torch.tensor([bt.n, bt.K, bt.P, bt.has_intercept], dtype=torch.int64)
    dist.broadcast(header, src, group=group)
    h = header.tolist()
    metas = [tuple(int(v) for v in h[1 + 4 * g: 5 + 4 * g]) for g in range(int(h[0]))]
    layout, total = [], 0
    for meta in metas:
        offs = []
        for dt, shape in _basis_shapes(*meta):
            nbytes = int(np.prod(shape)) * (4 if dt == torch.int32 else 8)
            offs.append((total, nbytes, dt, shape))
            total += (nbytes + 7) // 8 * 8
        layout.append(offs)
    payload = torch.zeros(max(total, 8), dtype=torch.uint8, device=device)
    if rank == src:
        for bt, offs in zip(mine, layout):
            for t, (off, nbytes, _, _) in zip(bt.tensors, offs):
                if nbytes:
                    payload[off:off + nbytes] = t.contiguous().view(torch.uint8).reshape(-1)
    if payload_transport is not None:
        payload = payload_transport(payload)     # e.g. PeerBroadcast: pulled over NVLink by the copy engines
    else:
        dist.broadcast(payload, src, group=group)
    out = []
    for meta, offs in zip(metas, layout):
        ts = [payload[off:off + nbytes].view(dt).reshape(shape) for off, nbytes, dt, shape in offs]
        out.append(BasisTensors(meta, ts))
    if message is not None:
        message.extend([header, payload])
    return out


class ShardedRegression:
    """`linear_regression_rows` over variant shards: one process per GPU, rank r holds a contiguous range of the
    variants (LinearRegression.scala:95 / :274: one task per partition).  Per call: ONE basis broadcast from rank 0
    (LR:74-78 / :257), the sweep of the local rows with no collective inside, and -- only on request -- an all-gather of
    the fixed-width result rows in rank order (= row order).  Results are bit-identical to the single-device run: every
    row's arithmetic is independent of where the row sits."""

    FIELDS = ("y_transpose_x", "beta", "standard_error", "t_stat", "p_value")

    def __init__(self, genotypes, group=None, transport="nccl"):
        """`transport`: how the per-call basis message travels -- "nccl" (dist.broadcast) or "peer" (PeerBroadcast: pulled
        over NVLink by the copy engines; needs torch symmetric memory, falls back to NCCL when it is unavailable)."""
        from . import _lib
        self.g = genotypes
        self.dev = genotypes.device
        self.ctx = _lib.context(self.dev.index)
        self.group = group
        self.bts = None
        self.transport = transport
        self._peer = None

    def set_bases(self, bases, src=0):
        """Collective.  Rank `src` passes the list of GroupBasis (the driver prologue ran there), the others None."""
        ctx, N = self.ctx, self.g.n_samples
        self._message, self._src = [], src
        self.bts = broadcast_bases(bases, self.dev, src, self.group, message=self._message,
                                   payload_transport=self._peer_payload if self.transport == "peer" else None)
        with torch.cuda.device(self.dev):
            # lrr_add_group copies on the library's default stream: the message must have landed (one host wait per call)
            torch.cuda.current_stream(self.dev).synchronize()
            ctx.check(ctx.lib.lrr_clear_groups(ctx.handle))
            for b in self.bts:
                t = b.tensors
                ctx.check(ctx.lib.lrr_add_group(ctx.handle, N, b.n, b.K, b.P, b.has_intercept, t[0].data_ptr(),
                                                t[1].data_ptr() if t[1].numel() else None, t[2].data_ptr(),
                                                t[3].data_ptr() if t[3].numel() else None, t[4].data_ptr()))
            ctx.check(ctx.lib.lrr_reserve(ctx.handle, self.g.n_variants))
        return self.bts

    def _peer_payload(self, payload):
        """The payload of the per-call message through PeerBroadcast (created, or re-created larger, collectively: every
        rank knows the size from the header)."""
        nbytes = payload.numel()
        if self._peer is None or self._peer.capacity < nbytes:
            try:
                self._peer = PeerBroadcast(nbytes, self.dev, self._src, self.group)
            except Exception as e:   # symmetric memory unavailable on this box / build
                self.transport = f"nccl (peer transport unavailable: {type(e).__name__}: {e})"
                dist.broadcast(payload, self._src, group=self.group)
                return payload
        return self._peer.send(payload, nbytes)

    def rebroadcast(self):
        """The per-call message once more (header + payload, asynchronous on the current stream): what a repeated call
        with the same phenotypes costs on the wire (bench.py)."""
        header, payload = self._message
        if self._peer is not None and self.transport == "peer":
            if getattr(self, "_staged", None) is None:
                self._staged = payload.clone()   # (the received payload IS the symmetric buffer: send from a copy)
            self._peer.send(self._staged, self._staged.numel())
            return
        dist.broadcast(header, self._src, group=self.group)
        dist.broadcast(payload, self._src, group=self.group)

    def alloc_outputs(self, want_log10_p=False):
        from . import _lib
        M, dev = self.g.n_variants, self.dev
        outs, arr = [], (_lib.GroupOut * len(self.bts))()
        for g, b in enumerate(self.bts):
            o = {"n": torch.empty(M, dtype=torch.int32, device=dev), "n_missing": torch.empty(M, dtype=torch.int32, device=dev),
                 "sum_x": torch.empty(M, dtype=torch.float64, device=dev)}
            for f in self.FIELDS + (("log10_p",) if want_log10_p else ()):
                o[f] = torch.empty((M, b.P), dtype=torch.float64, device=dev)
            for k, v in o.items():
                setattr(arr[g], k, v.data_ptr())
            if not want_log10_p:
                arr[g].log10_p = None
            outs.append(o)
        return outs, arr

    def run(self, kernel="auto", outs=None, arr=None, want_log10_p=False):
        """The local sweep + statistics (asynchronous on torch's current stream).  Returns the per-group output dicts."""
        from . import _lib
        if outs is None:
            outs, arr = self.alloc_outputs(want_log10_p)
        g, ctx = self.g, self.ctx
        with torch.cuda.device(self.dev):
            ctx.check(ctx.lib.lrr_run(ctx.handle, g.data.data_ptr(), g.flags_ptr(), g.n_variants, g.stride, g.n_samples, arr,
                                      len(self.bts), _lib.KERNELS[kernel], torch.cuda.current_stream(self.dev).cuda_stream))
        return outs

    @staticmethod
    def pack_rows(o, out=None):
        """One group's result fields as fixed-width float64 rows [M, 3 + 5 P] (n, n_missing, sum_x, then the P-wide fields)."""
        cols = [o["n"].to(torch.float64)[:, None], o["n_missing"].to(torch.float64)[:, None], o["sum_x"][:, None]]
        cols += [o[f] for f in ShardedRegression.FIELDS]
        return torch.cat(cols, dim=1, out=out)

    @staticmethod
    def unpack_rows(rows, P):
        o = {"n": rows[:, 0].to(torch.int32), "n_missing": rows[:, 1].to(torch.int32), "sum_x": rows[:, 2].contiguous()}
        for i, f in enumerate(ShardedRegression.FIELDS):
            o[f] = rows[:, 3 + i * P: 3 + (i + 1) * P].contiguous()
        return o

    def gather(self, outs):
        """All ranks' rows of every group, concatenated in rank order (collective)."""
        return [self.unpack_rows(gather_rows(self.pack_rows(o), group=self.group), b.P) for o, b in zip(outs, self.bts)]


def gather_rows(local_rows: torch.Tensor, counts=None, out=None, group=None):
    """All-gather row blocks [m_r, W] (m_r may differ per rank) and concatenate them in rank order.

    Equal blocks (the sharded sweep's normal case) go straight into one [world * m, W] tensor (`out`, allocated when
    None) with all_gather_into_tensor: no per-rank temporaries, no concatenation pass."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    dev = local_rows.device
    if counts is not None and len(set(counts)) == 1 and counts[0] == local_rows.shape[0]:
        if out is None:
            out = torch.empty((world * local_rows.shape[0], local_rows.shape[1]), dtype=local_rows.dtype, device=dev)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    if counts is None:
        c = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=dev)
        cs = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(cs, c, group=group)
        counts = [int(x.item()) for x in cs]
    width = local_rows.shape[1]
    mmax = max(counts)
    padded = local_rows
    if local_rows.shape[0] < mmax:
        padded = torch.zeros((mmax, width), dtype=local_rows.dtype, device=dev)
        padded[: local_rows.shape[0]] = local_rows
    parts = [torch.empty((mmax, width), dtype=local_rows.dtype, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class RowGather:
    """All-gather of fixed-width result rows [m, W] float64 (same m on every rank) into [world * m, W], rank order.

    `peer` mode: the output buffers live in symmetric memory (torch.distributed._symmetric_memory: every rank maps every
    peer's buffer over NVLink); a rank PUSHES its rows into its slot of every peer's buffer with plain device-to-device
    copies -- they run on the copy engines, so the gather takes no SM from a persistent sweep kernel running next to it
    (an NCCL all-gather's CTAs do: +0.8 ms on the 8-GPU step of round 1) -- and a symmetric-memory barrier publishes
    them.  `nccl` mode: all_gather_into_tensor on the caller's stream (own communicator).  `auto` tries peer first."""

    def __init__(self, m, width, device, n_buffers=2, mode="auto", group=None):
        self.m, self.width, self.dev = int(m), int(width), device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.mode = None
        self.why = None
        if mode in ("auto", "peer"):
            try:
                import torch.distributed._symmetric_memory as symm
                pg = group if group is not None else dist.group.WORLD
                self.bufs, self.hdls, self.peers = [], [], []
                for _ in range(n_buffers):
                    t = symm.empty((self.world * self.m, self.width), dtype=torch.float64, device=device)
                    h = symm.rendezvous(t, pg.group_name if hasattr(pg, "group_name") else pg)
                    self.bufs.append(t)
                    self.hdls.append(h)
                    self.peers.append([h.get_buffer(r, t.shape, t.dtype) for r in range(self.world)])
                self.mode = "peer (copy engines over symmetric memory)"
            except Exception as e:   # symmetric memory unavailable on this box / build
                self.why = f"{type(e).__name__}: {e}"
                if mode == "peer":
                    raise
        if self.mode is None:
            self.pg = dist.new_group(list(range(self.world))) if group is None else group
            self.bufs = [torch.empty((self.world * self.m, self.width), dtype=torch.float64, device=device) for _ in range(n_buffers)]
            self.mode = "nccl all_gather_into_tensor" + (f" (peer mode unavailable: {self.why})" if self.why else "")

    def gather(self, rows, b=0):
        """Asynchronous on the current stream; returns buffer b holding every rank's rows once the stream reaches here."""
        if self.mode.startswith("peer"):
            lo, hi = self.rank * self.m, (self.rank + 1) * self.m
            self.hdls[b].barrier(channel=0)          # every peer is done reading buffer b from its previous use
            for r in range(self.world):
                self.peers[b][(self.rank + r) % self.world][lo:hi].copy_(rows, non_blocking=True)
            self.hdls[b].barrier(channel=1)          # every push has landed everywhere
        else:
            dist.all_gather_into_tensor(self.bufs[b], rows.contiguous(), group=self.pg)
        return self.bufs[b]


class PeerBroadcast:
    """One-to-all copy of a byte message through symmetric memory: the root stages the message in its own symmetric buffer
    and signals every peer (point to point); a peer waits for that signal only, PULLS the bytes over NVLink with a plain
    device-to-device copy (copy engines) and signals the root back, which the root collects before it overwrites the
    buffer with the next message.  No rank ever waits for a rank other than the root, and no SM spins inside a collective
    kernel next to the sweep (the per-call NCCL broadcast cost the 8-GPU step of round 2 1.4 ms: it acts as a barrier over
    ranks whose sweep times differ by 15 %, and its polling CTAs draw from the same power budget)."""

    def __init__(self, nbytes, device, src=0, group=None):
        import torch.distributed._symmetric_memory as symm
        self.src, self.dev = src, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        pg = group if group is not None else dist.group.WORLD
        self.capacity = int(max(nbytes, 1 << 20))
        self.buf = symm.empty(self.capacity, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, pg.group_name if hasattr(pg, "group_name") else pg)
        self.root_view = self.hdl.get_buffer(src, (self.capacity,), torch.uint8)
        self.sent = 0

    def send(self, message, nbytes):
        """Collective, asynchronous on the current stream.  `message` (uint8, >= nbytes) is read on the root only; returns
        a uint8 view of the local copy (valid until the next send)."""
        if self.rank == self.src:
            if self.sent:
                for r in range(self.world):
                    if r != self.src:
                        self.hdl.wait_signal(r, channel=3)     # peer r has pulled the previous message
            self.buf[:nbytes].copy_(message[:nbytes], non_blocking=True)
            for r in range(self.world):
                if r != self.src:
                    self.hdl.put_signal(r, channel=2)
        else:
            self.hdl.wait_signal(self.src, channel=2)
            self.buf[:nbytes].copy_(self.root_view[:nbytes], non_blocking=True)
            self.hdl.put_signal(self.src, channel=3)
        self.sent += 1
        return self.buf[:nbytes]

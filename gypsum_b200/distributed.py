"""Multi-GPU sharding of the acquisition search (one process per GPU, torch.distributed for the plumbing).

The path partitions: (PRN, Doppler, block) cells share only read-only input (reference acquisition.py:59-67 loops
satellites independently).  Two shapes:

* many independent blocks (BASELINE config 5, bench.py): `ShardedBlockSearch` -- the IQ lives on rank 0's host; ONE
  scatter hands every rank its contiguous share of blocks, every rank searches the full PRN x Doppler grid on its share,
  ONE gather brings the per-cell records (or, with reduce="best", the per-(block, PRN) best-bin records of
  acquisition.py:179-189: 41..81 times fewer bytes) back to rank 0's host.  No collective between the two;
* ONE short block searched over all PRNs (configs 2/3, the real detector): `ShardedGridSearch` broadcasts the IQ block
  once, every rank searches its PRN rows, and the per-cell records are all-gathered -- the single broadcast + final
  gather of per-cell peaks BASELINE.json's north_star describes.  Two collectives of ~100 us each around ~30 us of work:
  it is SLOWER than one GPU for a single 1-ms block (bench.py reports the figure), which is why a receiver shards blocks.

Stream discipline: NCCL collectives are ordered against torch's CURRENT stream, so the engine is switched to launch on that
stream (`engine.set_stream`) when a search object is built on a CUDA device -- the kernels then run after the scatter /
broadcast has landed and the gather after the kernels, without host synchronisation in between.
"""
from __future__ import annotations

import numpy as np

RECORD_BYTES = 32


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _is_cuda(device) -> bool:
    return str(device) != "cpu"


def _adopt_current_stream(engine, device) -> None:
    """Make the engine launch on torch's current stream of `device` (see the module docstring)."""
    if _is_cuda(device) and hasattr(engine, "set_stream"):
        import torch

        engine.set_stream(torch.cuda.current_stream(device).cuda_stream)


class ShardedGridSearch:
    """`engine` needs the gypsum_b200._native.Engine methods bind_iq_device / acquire_grid_device (the tests pass
    a CPU stand-in to exercise the sharding and gather order under gloo)."""

    def __init__(self, engine, device, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.engine = engine
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        _adopt_current_stream(engine, device)

    def acquire_grid(self, iq, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int):
        """iq: complex64 ndarray on rank 0 (ignored elsewhere).  Returns the full record array
        [n_blocks, n_prn, n_doppler] on every rank."""
        import torch

        from gypsum_b200._native import RECORD_DTYPE

        prn_idx = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        n_samples = n_blocks * ms_per_block * self.engine.samples_per_ms
        buf = torch.empty(n_samples * 2, dtype=torch.float32, device=self.device)
        if self.rank == 0:
            buf.copy_(torch.from_numpy(np.ascontiguousarray(iq, dtype=np.complex64)[:n_samples].view(np.float32)))
        self.dist.broadcast(buf, src=0, group=self.group)  # the one broadcast of the IQ block

        mine = shard_range(prn_idx.size, self.rank, self.world)
        per_rank = -(-prn_idx.size // self.world)  # padded shard so the all-gather is one fixed-size call
        out = torch.zeros(n_blocks * per_rank * dop.size * RECORD_BYTES, dtype=torch.uint8, device=self.device)
        if len(mine):
            my_prn = np.ascontiguousarray(prn_idx[mine.start:mine.stop])
            self.engine.bind_iq_device(buf.data_ptr(), n_samples)
            rows = torch.empty(n_blocks * len(mine) * dop.size * RECORD_BYTES, dtype=torch.uint8, device=self.device)
            # same stream as the broadcast before and the copy / all-gather after: ordered on the device
            self.engine.acquire_grid_device(n_blocks, ms_per_block, my_prn, dop, kind, rows.data_ptr())
            out.view(n_blocks, per_rank, dop.size * RECORD_BYTES)[:, :len(mine)] = rows.view(n_blocks, len(mine), -1)
        gathered = torch.empty(self.world * out.numel(), dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(gathered, out, group=self.group)  # the one gather of per-cell peaks
        g = gathered.cpu().numpy().view(RECORD_DTYPE).reshape(self.world, n_blocks, per_rank, dop.size)
        full = np.empty((n_blocks, prn_idx.size, dop.size), dtype=RECORD_DTYPE)
        for r in range(self.world):
            rr = shard_range(prn_idx.size, r, self.world)
            full[:, rr.start:rr.stop] = g[r][:, :len(rr)]
        return full


class ShardedBlockSearch:
    """Many independent blocks (BASELINE config 5: 1000 x 1-ms blocks over 8 GPUs): rank 0 holds the IQ on its host, each
    rank receives only its contiguous share of blocks (one scatter), searches the full PRN x Doppler grid on them, and the
    records come back to rank 0's host with one gather.  Buffers (device shares, the pinned receive buffer on rank 0) are kept
    between calls of the same shape.  `last_bytes` reports what the last call moved: host->device and device->host on rank
    0, scatter / gather payload over the interconnect (bytes leaving / reaching rank 0, its own share excluded)."""

    def __init__(self, engine, device, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.engine = engine
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._shape = None
        self.last_bytes = {}
        _adopt_current_stream(engine, device)

    def _buffers(self, most: int, per_block: int, row_bytes: int):
        import torch

        shape = (most, per_block, row_bytes)
        if self._shape != shape:
            cuda = _is_cuda(self.device)
            self._mine = torch.empty(most * per_block, dtype=torch.float32, device=self.device)
            self._out = torch.zeros(most * row_bytes, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                self._all_iq = torch.empty(self.world * most * per_block, dtype=torch.float32, device=self.device)
                self._all_out = torch.empty(self.world * most * row_bytes, dtype=torch.uint8, device=self.device)
                self._h_out = torch.empty(self.world * most * row_bytes, dtype=torch.uint8, pin_memory=cuda)
            self._shape = shape
        return self._mine, self._out

    def acquire_blocks(self, iq, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int, reduce: str | None = None,
                       copy: bool = True):
        """iq: complex64[n_blocks * ms_per_block * N] on rank 0 (ignored elsewhere; DMA'd in place when it lives in pinned
        memory).  Returns on rank 0 the record array [n_blocks, n_prn, n_doppler] (RECORD_DTYPE), or with reduce="best" the
        array [n_blocks, n_prn] (BEST_DTYPE) of acquisition.py:179-189 per row; None on the other ranks.  copy=False hands
        out a view of the pinned receive buffer when the shares are equal (valid until the next call)."""
        import torch

        from gypsum_b200._native import BEST_DTYPE, RECORD_DTYPE

        if reduce not in (None, "best"):
            raise ValueError("reduce must be None or 'best'")
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        per_block = ms_per_block * self.engine.samples_per_ms * 2  # float32 words
        row_bytes = prn.size * (RECORD_BYTES if reduce else dop.size * RECORD_BYTES)  # per block
        shares = [shard_range(n_blocks, r, self.world) for r in range(self.world)]
        most = max(len(s) for s in shares)
        mine, out = self._buffers(most, per_block, row_bytes)
        cuda = _is_cuda(self.device)

        # ---- one scatter of equal-sized (padded) shares ----
        if self.rank == 0:
            words = np.ascontiguousarray(iq, dtype=np.complex64)[: n_blocks * per_block // 2].view(np.float32)
            src = torch.from_numpy(words)
            nb_ = cuda and src.is_pinned()  # pinned caller memory: asynchronous DMA straight from it, no staging copy
            rows = self._all_iq.view(self.world, most * per_block)
            if all(len(s) == most for s in shares):
                self._all_iq.copy_(src, non_blocking=nb_)
            else:
                for r, s in enumerate(shares):
                    rows[r, : len(s) * per_block].copy_(src[s.start * per_block: s.stop * per_block], non_blocking=nb_)
            self.dist.scatter(mine, list(rows.unbind(0)), src=0, group=self.group)
        else:
            self.dist.scatter(mine, None, src=0, group=self.group)

        # ---- the grid on this rank's share (same stream: after the scatter, before the gather) ----
        my = shares[self.rank]
        if len(my):
            self.engine.bind_iq_device(mine.data_ptr(), len(my) * per_block // 2)
            if reduce:
                self.engine.acquire_grid_best_device(len(my), ms_per_block, prn, dop, kind, out.data_ptr())
            else:
                self.engine.acquire_grid_device(len(my), ms_per_block, prn, dop, kind, out.data_ptr())

        # ---- one gather of the records ----
        if self.rank == 0:
            self.dist.gather(out, list(self._all_out.view(self.world, -1).unbind(0)), dst=0, group=self.group)
            self._h_out.copy_(self._all_out, non_blocking=cuda)
            if cuda:
                torch.cuda.current_stream(self.device).synchronize()
        else:
            self.dist.gather(out, None, dst=0, group=self.group)
        share_iq, share_out = most * per_block * 4, most * row_bytes
        self.last_bytes = {"h2d": self.world * share_iq if self.rank == 0 else 0,
                           "scatter": (self.world - 1) * share_iq, "gather": (self.world - 1) * share_out,
                           "d2h": self.world * share_out if self.rank == 0 else 0}
        if self.rank != 0:
            return None
        dtype = BEST_DTYPE if reduce else RECORD_DTYPE
        tail = (prn.size,) if reduce else (prn.size, dop.size)
        raw = self._h_out.numpy()  # bytes: numpy copies structured records field by field, raw bytes with memcpy
        if all(len(s) == most for s in shares):
            flat = raw[: n_blocks * row_bytes]
            return (flat.copy() if copy else flat).view(dtype).reshape((n_blocks,) + tail)
        full = np.empty(n_blocks * row_bytes, dtype=np.uint8)
        for r, s in enumerate(shares):
            full[s.start * row_bytes: s.stop * row_bytes] = raw[r * most * row_bytes: (r * most + len(s)) * row_bytes]
        return full.view(dtype).reshape((n_blocks,) + tail)

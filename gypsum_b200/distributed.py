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


class ShardedBlockStream:
    """ShardedBlockSearch for a STREAM of equally shaped jobs (the receiver's steady state, and bench.py's steps): `submit`
    enqueues one job -- rank 0's host->device copy on a copy stream, ONE scatter, the grid on every rank, ONE gather, rank 0's
    device->host copy on a second copy stream -- and returns; `collect` waits for the oldest job in flight and hands out its
    table.  Two jobs may be in flight: the copies of job k+1 / k-1 run on the copy engines under job k's kernels (they need
    no SM, unlike NCCL's kernels, which is why the two collectives stay between the kernel phases on the main stream).  Every
    rank calls submit / collect in the same order.  Equal shares only (n_blocks divisible by the world size)."""

    def __init__(self, engine, device, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int, reduce: str | None = None,
                 group=None):
        import torch
        import torch.distributed as dist

        from gypsum_b200._native import BEST_DTYPE, RECORD_DTYPE

        if reduce not in (None, "best"):
            raise ValueError("reduce must be None or 'best'")
        self.dist, self.engine, self.device, self.group = dist, engine, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if n_blocks % self.world:
            raise ValueError("ShardedBlockStream needs n_blocks divisible by the number of ranks")
        self.prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        self.dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        self.n_blocks, self.ms, self.kind, self.reduce = n_blocks, ms_per_block, kind, reduce
        self.share = n_blocks // self.world
        self.per_block = ms_per_block * engine.samples_per_ms * 2  # float32 words
        self.row_bytes = self.prn.size * (RECORD_BYTES if reduce else self.dop.size * RECORD_BYTES)
        self.dtype = BEST_DTYPE if reduce else RECORD_DTYPE
        self.tail = (self.prn.size,) if reduce else (self.prn.size, self.dop.size)
        self.cuda = _is_cuda(device)
        _adopt_current_stream(engine, device)
        self.main = torch.cuda.current_stream(device) if self.cuda else None
        self.s_in = torch.cuda.Stream(device) if self.cuda else None
        self.s_out = torch.cuda.Stream(device) if self.cuda else None
        self.slots = []
        for _ in range(2):
            sl = {"mine": torch.empty(self.share * self.per_block, dtype=torch.float32, device=device),
                  "out": torch.zeros(self.share * self.row_bytes, dtype=torch.uint8, device=device)}
            if self.rank == 0:
                sl["all_iq"] = torch.empty(n_blocks * self.per_block, dtype=torch.float32, device=device)
                sl["all_out"] = torch.empty(n_blocks * self.row_bytes, dtype=torch.uint8, device=device)
                sl["h_out"] = torch.empty(n_blocks * self.row_bytes, dtype=torch.uint8, pin_memory=self.cuda)
            if self.cuda:
                sl["e_in"], sl["e_scattered"], sl["e_done"], sl["e_host"] = (torch.cuda.Event() for _ in range(4))
            self.slots.append(sl)
        self.head = self.tail_ix = 0
        share_iq, share_out = self.share * self.per_block * 4, self.share * self.row_bytes
        self.bytes_per_job = {"h2d": self.world * share_iq, "scatter": (self.world - 1) * share_iq,
                              "gather": (self.world - 1) * share_out, "d2h": self.world * share_out}

    @property
    def in_flight(self) -> int:
        return self.head - self.tail_ix

    def submit(self, iq) -> None:
        """iq: complex64[n_blocks * ms_per_block * N] on rank 0 (pinned memory = asynchronous DMA), ignored elsewhere."""
        import torch

        if self.in_flight >= 2:
            raise RuntimeError("two jobs in flight: collect one first")
        sl = self.slots[self.head % 2]
        if self.rank == 0:
            src = torch.from_numpy(np.ascontiguousarray(iq, dtype=np.complex64)[: self.n_blocks * self.per_block // 2].view(np.float32))
            if self.cuda:
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_event(sl["e_scattered"])  # the scatter that last read this buffer
                    sl["all_iq"].copy_(src, non_blocking=src.is_pinned())
                    sl["e_in"].record(self.s_in)
                self.main.wait_event(sl["e_in"])
                self.main.wait_event(sl["e_host"])  # the device->host copy that last read all_out
            else:
                sl["all_iq"].copy_(src)
            self.dist.scatter(sl["mine"], list(sl["all_iq"].view(self.world, -1).unbind(0)), src=0, group=self.group)
        else:
            self.dist.scatter(sl["mine"], None, src=0, group=self.group)
        if self.cuda:
            sl["e_scattered"].record(self.main)
        self.engine.bind_iq_device(sl["mine"].data_ptr(), self.share * self.per_block // 2)
        fn = self.engine.acquire_grid_best_device if self.reduce else self.engine.acquire_grid_device
        fn(self.share, self.ms, self.prn, self.dop, self.kind, sl["out"].data_ptr())
        if self.rank == 0:
            self.dist.gather(sl["out"], list(sl["all_out"].view(self.world, -1).unbind(0)), dst=0, group=self.group)
            if self.cuda:
                sl["e_done"].record(self.main)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(sl["e_done"])
                    sl["h_out"].copy_(sl["all_out"], non_blocking=True)
                    sl["e_host"].record(self.s_out)
            else:
                sl["h_out"].copy_(sl["all_out"])
        else:
            self.dist.gather(sl["out"], None, dst=0, group=self.group)
            if self.cuda:
                sl["e_done"].record(self.main)
        self.head += 1

    def collect(self):
        """Rank 0: the oldest job's table as a view of its pinned receive buffer (valid until two more jobs were submitted);
        other ranks: None (after their share of that job has left the device)."""
        if not self.in_flight:
            raise RuntimeError("no job in flight")
        sl = self.slots[self.tail_ix % 2]
        self.tail_ix += 1
        if self.cuda:
            (sl["e_host"] if self.rank == 0 else sl["e_done"]).synchronize()
        if self.rank != 0:
            return None
        return sl["h_out"].numpy().view(self.dtype).reshape((self.n_blocks,) + self.tail)

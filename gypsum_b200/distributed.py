"""Multi-GPU sharding of the acquisition search (one process per GPU, torch.distributed for the plumbing).

The path partitions: (PRN, Doppler, block) cells share only read-only input (reference acquisition.py:59-67 loops
satellites independently).  Two shapes:

* many independent blocks (BASELINE config 5, bench.py): `shard_range` gives every rank its own blocks -- no
  collective in the data path;
* ONE short block searched over all PRNs (configs 2/3, the real detector): `ShardedGridSearch` broadcasts the IQ
  block once, every rank searches its PRN rows, and the per-cell records are all-gathered -- the single
  broadcast + final gather of per-cell peaks BASELINE.json's north_star describes.  Payloads are KB..MB, far
  below NVLink bandwidth; latency is what counts, so both collectives are single calls.
"""
from __future__ import annotations

import numpy as np

RECORD_BYTES = 32


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


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
            self.engine.acquire_grid_device(n_blocks, ms_per_block, my_prn, dop, kind, rows.data_ptr())
            if self.device != "cpu" and str(self.device) != "cpu":
                torch.cuda.current_stream().synchronize()
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
    """Many independent blocks (BASELINE config 5: 1000 x 1-ms blocks over 8 GPUs): rank 0 holds the IQ, each rank
    receives only its contiguous share of blocks (one scatter), searches the full PRN x Doppler grid on them, and the
    per-cell records come back with one gather.  No collective between the two."""

    def __init__(self, engine, device, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.engine = engine
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def acquire_blocks(self, iq, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int):
        """iq: complex64[n_blocks * ms_per_block * N] on rank 0 (ignored elsewhere).  Returns the record array
        [n_blocks, n_prn, n_doppler] on rank 0 and None on the other ranks."""
        import torch

        from gypsum_b200._native import RECORD_DTYPE

        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        per_block = ms_per_block * self.engine.samples_per_ms * 2  # float32 words
        shares = [shard_range(n_blocks, r, self.world) for r in range(self.world)]
        most = max(len(s) for s in shares)
        # one scatter of equal-sized (padded) shares
        mine = torch.empty(most * per_block, dtype=torch.float32, device=self.device)
        if self.rank == 0:
            words = torch.from_numpy(np.ascontiguousarray(iq, dtype=np.complex64)[: n_blocks * per_block // 2].view(np.float32))
            parts = []
            for s in shares:
                part = torch.zeros(most * per_block, dtype=torch.float32)
                part[: len(s) * per_block] = words[s.start * per_block: s.stop * per_block]
                parts.append(part.to(self.device))
            self.dist.scatter(mine, parts, src=0, group=self.group)
        else:
            self.dist.scatter(mine, None, src=0, group=self.group)

        my = shares[self.rank]
        rec_bytes = most * prn.size * dop.size * RECORD_BYTES
        out = torch.zeros(rec_bytes, dtype=torch.uint8, device=self.device)
        if len(my):
            self.engine.bind_iq_device(mine.data_ptr(), len(my) * per_block // 2)
            self.engine.acquire_grid_device(len(my), ms_per_block, prn, dop, kind, out.data_ptr())
            if str(self.device) != "cpu":
                torch.cuda.current_stream().synchronize()
        # one gather of the records
        gathered = [torch.empty_like(out) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(out, gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None
        full = np.empty((n_blocks, prn.size, dop.size), dtype=RECORD_DTYPE)
        for r, s in enumerate(shares):
            g = gathered[r].cpu().numpy().view(RECORD_DTYPE).reshape(most, prn.size, dop.size)
            full[s.start:s.stop] = g[: len(s)]
        return full

"""N > 1 host logic under gloo on the CPU (world_size 2 and 3): PRN sharding, the single broadcast and the gather
order of ShardedGridSearch, with a CPU stand-in for the engine (the oracle computes each rank's rows)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from gypsum_b200.distributed import shard_range

    for n in (0, 1, 5, 32, 1000):
        for world in (1, 2, 3, 8):
            parts = [list(shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


class _OracleEngine:
    """CPU stand-in with the engine methods ShardedGridSearch uses; 'device pointers' are torch CPU tensors."""
    samples_per_ms = 2046

    def __init__(self):
        self.iq = None

    def bind_iq_device(self, ptr, n):
        import ctypes
        self.iq = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(2 * n,)).view(np.complex64).copy()

    def acquire_grid_device(self, n_blocks, ms, prn, dop, kind, out_ptr):
        import ctypes
        from gypsum_b200._native import RECORD_DTYPE
        from oracle import gypsum_oracle as o

        n = self.samples_per_ms
        rec = np.zeros((n_blocks, prn.size, dop.size), dtype=RECORD_DTYPE)
        for b in range(n_blocks):
            peak, arg, total, count = o.grid_cells(self.iq[b * ms * n:(b + 1) * ms * n], n * 1000, n, [int(p) + 1 for p in prn], list(dop))
            rec[b]["peak"], rec[b]["argmax"], rec[b]["sum"], rec[b]["count"] = peak, arg, total, count
        ctypes.memmove(out_ptr, rec.ctypes.data, rec.nbytes)
        return rec

    def acquire_grid_best_device(self, n_blocks, ms, prn, dop, kind, out_ptr):
        """acquisition.py:179-189 per (block, prn) row of the same grid."""
        import ctypes
        from gypsum_b200._native import BEST_DTYPE

        scratch = np.zeros(n_blocks * prn.size * dop.size * 32, dtype=np.uint8)
        rec = self.acquire_grid_device(n_blocks, ms, prn, dop, kind, scratch.ctypes.data)
        best = np.zeros((n_blocks, prn.size), dtype=BEST_DTYPE)
        n = self.samples_per_ms
        for b in range(n_blocks):
            for a in range(prn.size):
                k = int(np.argmax(rec[b, a]["peak"]))
                r = rec[b, a, k]
                best[b, a] = (dop[k], r["peak"] / ((r["sum"] - r["count"] * r["peak"]) / (n - r["count"])), r["peak"], r["argmax"], k, 0)
        ctypes.memmove(out_ptr, best.ctypes.data, best.nbytes)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gypsum_b200.distributed import ShardedGridSearch
    from oracle import gypsum_oracle as o

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = o.synth_iq(4, 2046, 2, 2046000, [(5, 1000.0, 99, 0.0, 0.4)]) if rank == 0 else None
    search = ShardedGridSearch(_OracleEngine(), "cpu")
    full = search.acquire_grid(x, 2, 1, np.arange(5), [0.0, 1000.0], 2)
    q.put((rank, full["peak"].copy(), full["argmax"].copy()))
    dist.destroy_process_group()


def _block_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gypsum_b200.distributed import ShardedBlockSearch
    from oracle import gypsum_oracle as o

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = o.synth_iq(6, 2046, 5, 2046000, [(9, -500.0, 1234, 0.0, 0.4)]) if rank == 0 else None
    search = ShardedBlockSearch(_OracleEngine(), "cpu")
    full = search.acquire_blocks(x, 5, 1, np.array([8, 0]), [-500.0, 0.0], 2)
    moved = dict(search.last_bytes)
    best = search.acquire_blocks(x, 5, 1, np.array([8, 0]), [-500.0, 0.0], 2, reduce="best")
    moved_best = search.last_bytes
    view = search.acquire_blocks(x, 5, 1, np.array([8, 0]), [-500.0, 0.0], 2, copy=False)
    assert view is None or np.array_equal(view["peak"], full["peak"])
    # the pipelined stream form (equal shares): 6 blocks, two jobs in flight, tables identical to the one-call form
    from gypsum_b200.distributed import ShardedBlockStream
    x6 = o.synth_iq(7, 2046, 6, 2046000, [(9, -500.0, 1234, 0.0, 0.4)]) if rank == 0 else None
    if world in (2, 3):
        stream = ShardedBlockStream(_OracleEngine(), "cpu", 6, 1, np.array([8, 0]), [-500.0, 0.0], 2)
        stream.submit(x6)
        stream.submit(x6)
        a = stream.collect()
        b = stream.collect()
        one = search.acquire_blocks(x6, 6, 1, np.array([8, 0]), [-500.0, 0.0], 2)
        assert a is None or (np.array_equal(a["peak"], one["peak"]) and np.array_equal(b["argmax"], one["argmax"]))
        assert stream.bytes_per_job["gather"] == (world - 1) * (6 // world) * 2 * 2 * 32
    q.put((rank, None if full is None else (full["peak"].copy(), full["argmax"].copy(), best.copy(), moved, dict(moved_best))))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_block_search_scatter_and_gather(world):
    """5 blocks over 2 / 3 ranks (uneven shares): one scatter of blocks, one gather of records, rank 0 gets the table
    in block order."""
    from oracle import gypsum_oracle as o

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_block_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(got[r] is None for r in range(1, world))
    peak, arg, best, moved, moved_best = got[0]
    x = o.synth_iq(6, 2046, 5, 2046000, [(9, -500.0, 1234, 0.0, 0.4)])
    for b in range(5):
        pk, ag, _, _ = o.grid_cells(x[b * 2046:(b + 1) * 2046], 2046000, 2046, [9, 1], [-500.0, 0.0])
        assert np.allclose(peak[b], pk, rtol=1e-6) and np.array_equal(arg[b], ag)
        assert arg[b, 0, 0] == 1234
        # reduce="best": acquisition.py:179-189 of each (block, prn) row -- SV9 is found in bin 0 at code phase 1234
        assert (best[b, 0]["bin"], best[b, 0]["doppler"], best[b, 0]["code_phase"]) == (0, -500.0, 1234)
        for a in range(2):
            assert best[b, a]["bin"] == int(np.argmax(pk[a])) and np.isclose(best[b, a]["peak"], pk[a].max(), rtol=1e-6)
    most = -(-5 // world)
    assert moved["scatter"] == (world - 1) * most * 2046 * 8 and moved["gather"] == (world - 1) * most * 2 * 2 * 32
    assert moved_best["gather"] == (world - 1) * most * 2 * 32  # one 32-byte record per (block, prn) instead of per cell


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_grid_search_matches_single_process(world):
    from oracle import gypsum_oracle as o

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = o.synth_iq(4, 2046, 2, 2046000, [(5, 1000.0, 99, 0.0, 0.4)])
    for b in range(2):
        peak, arg, _, _ = o.grid_cells(x[b * 2046:(b + 1) * 2046], 2046000, 2046, [1, 2, 3, 4, 5], [0.0, 1000.0])
        for rank, pk, ag in got:  # every rank ends up with the full, correctly ordered table
            assert np.allclose(pk[b], peak, rtol=1e-6) and np.array_equal(ag[b], arg)
    assert got[0][2][0, 4, 1] == 99

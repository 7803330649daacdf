"""torchrun --nproc-per-node N tools/multi_gpu_search.py : ONE short block searched over all PRNs, sharded by PRN with
a single NCCL broadcast of the IQ block and a single all-gather of the per-cell records (gypsum_b200.distributed).
Checks the gathered table against a single-GPU run on rank 0 and prints the timing (max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gypsum_b200 import _native  # noqa: E402
from gypsum_b200.distributed import ShardedBlockSearch, ShardedGridSearch  # noqa: E402
from gypsum_b200.gps_ca_prn_codes import ca_code_chips  # noqa: E402
from gypsum_b200 import synth as o  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n, m in ((2046, 1), (4092, 10)):
    fs = n * 1000
    eng = _native.Engine(fs, n, device=local)
    eng.set_replicas(np.stack([ca_code_chips(sv) for sv in range(1, 33)]).astype(np.uint8))
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    eng.set_stream(st.cuda_stream)
    x = o.synth_iq(2, n, m, fs, [(25, 1500.0, 777, 0.3, 0.3), (3, -3000.0, 5, 1.0, 0.3)]) if rank == 0 else None
    dop = np.arange(-10000, 10001, 500.0)
    search = ShardedGridSearch(eng, torch.device("cuda", local))
    full = search.acquire_grid(x, 1, m, np.arange(32), dop, _native.NON_COHERENT)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record(st)
    for _ in range(20):
        full = search.acquire_grid(x, 1, m, np.arange(32), dop, _native.NON_COHERENT)
    e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        eng.upload_iq(x)
        one = eng.acquire_grid(1, m, np.arange(32), dop)
        ok = all(np.array_equal(full[k], one[k]) for k in ("peak", "argmax", "count"))
        b = int(np.argmax(full["peak"][0, 24]))
        print(json.dumps({"workload": f"PRN-sharded search, N={n}, {m} ms, 32x41 cells", "world": world, "ms_per_search": float(ms.item()),
                          "identical_to_single_gpu": bool(ok), "sv25": [float(dop[b]), int(full["argmax"][0, 24, b])]}), flush=True)
    eng.set_stream(0)
    eng.close()
# config-5 shape: independent 1-ms blocks @ 16.368 Msps, 32 x 81 cells each, blocks scattered over the ranks
n, m, nb = 16368, 1, 24
fs = n * 1000
eng = _native.Engine(fs, n, device=local)
eng.set_replicas(np.stack([ca_code_chips(sv) for sv in range(1, 33)]).astype(np.uint8))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng.set_stream(st.cuda_stream)
dop = np.arange(-10000, 10001, 250.0)
x = None
if rank == 0:
    rng = np.random.default_rng(3)
    x = ((rng.standard_normal(n * nb, dtype=np.float32) + 1j * rng.standard_normal(n * nb, dtype=np.float32)) * np.float32(0.7071)).astype(np.complex64)
    x[: n] += o.synth_iq(0, n, 1, fs, [(25, 1500.0, 7777, 0.3, 0.3)], sigma=0.0)
search = ShardedBlockSearch(eng, torch.device("cuda", local))
full = search.acquire_blocks(x, nb, m, np.arange(32), dop, _native.NON_COHERENT)
torch.cuda.synchronize(); dist.barrier()
import time
t0 = time.perf_counter()
full = search.acquire_blocks(x, nb, m, np.arange(32), dop, _native.NON_COHERENT)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    eng.upload_iq(x)
    one = eng.acquire_grid(nb, m, np.arange(32), dop)
    ok = all(np.array_equal(full[k], one[k]) for k in ("peak", "argmax", "count"))
    b = int(np.argmax(full["peak"][0, 24]))
    print(json.dumps({"workload": f"block-sharded search (scatter + gather), {nb} blocks @ 16.368 Msps, 32x81 cells", "world": world,
                      "seconds_host_to_host": dt, "identical_to_single_gpu": bool(ok), "sv25": [float(dop[b]), int(full["argmax"][0, 24, b])]}), flush=True)
eng.set_stream(0)
eng.close()
dist.destroy_process_group()

"""A few launch pairs of one bench shape, for ncu (DRAM bytes of the config-3 / config-5 correlate launches):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
        -k regex:k_correlate -s 3 -c 2 --csv --log-file gpurun_out/traffic_cfg3.csv python tools/profile_shapes.py config3"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gypsum_b200 import _native  # noqa: E402
from gypsum_b200.gps_ca_prn_codes import ca_code_chips  # noqa: E402

SHAPES = {"config3": (4092, 10, 41, 1), "config5": (16368, 1, 81, 24), "config2": (2046, 1, 41, 256)}
n, m, n_dop, n_blocks = SHAPES[sys.argv[1]]
eng = _native.Engine(n * 1000, n)
eng.set_replicas(np.stack([ca_code_chips(sv) for sv in range(1, 33)]).astype(np.uint8))
rng = np.random.default_rng(1)
x = (rng.standard_normal(n * m * n_blocks * 2).astype(np.float32)).view(np.complex64)
xd = torch.from_numpy(x).cuda()
out = torch.empty(n_blocks * 32 * n_dop * 32, dtype=torch.uint8, device="cuda")
eng.bind_iq_device(xd.data_ptr(), x.size)
prn = np.arange(32, dtype=np.int32)
dop = np.linspace(-10000, 10000, n_dop)
for _ in range(6):
    eng.acquire_grid_device(n_blocks, m, prn, dop, 2, out.data_ptr())
torch.cuda.synchronize()

"""Short tracking run for ncu: 32 channels x 300 ms.  usage (on the GPU box): ncu ... python tools/profile_tracker.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gypsum_b200 import _native  # noqa: E402
from gypsum_b200.gps_ca_prn_codes import ca_code_chips  # noqa: E402
from gypsum_b200 import synth as to  # noqa: E402

n, fs, n_ch, n_ms = 2046, 2046000, 32, 300
eng = _native.Engine(fs, n)
eng.set_replicas(np.stack([ca_code_chips(sv) for sv in range(1, 33)]).astype(np.uint8))
chans = [(sv, 1000.0 + 37.3 * sv, 0.0, (53 * sv) % n, 0.1 * sv, 0.004) for sv in range(1, n_ch + 1)]
x = to.synth_tracking_iq(5, n, n_ms, fs, chans)
trk = _native.Tracker(eng, [c[0] - 1 for c in chans], [c[1] for c in chans], [0.0] * n_ch, [c[3] for c in chans])
eng.upload_iq(x)
times = np.array([round(k * n / fs, 6) for k in range(n_ms)])
for _ in range(2):
    rec = trk.process(n_ms, times)
print("locked", rec["locked"].mean())

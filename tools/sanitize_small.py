"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck).
usage (GPU box): compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gypsum_b200 import _native  # noqa: E402
from gypsum_b200.gps_ca_prn_codes import ca_code_chips  # noqa: E402
from gypsum_b200 import synth as o  # noqa: E402
from gypsum_b200 import synth as to  # noqa: E402

chips = np.stack([ca_code_chips(sv) for sv in range(1, 33)]).astype(np.uint8)
for n, m in ((2046, 2), (4092, 1)):
    fs = n * 1000
    eng = _native.Engine(fs, n)
    eng.set_replicas(chips)
    x = o.synth_iq(1, n, m, fs, [(25, 1500.0, 777 % n, 0.3, 0.3)])
    eng.upload_iq(x)
    dop = np.arange(-2000, 2001, 500.0)
    g = eng.acquire_grid(1, m, [24, 0, 5], dop)                       # split path (rsplit > 1)
    c = eng.acquire_cells([24, 3, 24, 7], [1500.0, 0.0, 1000.0, -250.0], m, _native.COHERENT, probe_idx=[777 % n, 0, 1, 2])
    p = eng.correlation_profile(24, 1500.0, m, _native.NON_COHERENT)
    assert int(p.argmax()) == 777 % n == int(g["argmax"][0, 0, 7])
    if n == 2046:
        big = eng.acquire_grid(2, 1, np.arange(32), np.linspace(-10000, 10000, 161))  # whole-cell-per-pair path, 10-pair build
        r = eng.detect([24, 2], m)
        assert int(big["argmax"][0, 24, int(np.argmax(big["peak"][0, 24]))]) == 777 and int(r["code_phase"][0]) == 777
        xs = to.synth_tracking_iq(3, n, 12, fs, [(25, 1500.3, 0.0, 777, 0.3, 0.004)])
        eng.upload_iq(xs)
        t = _native.Tracker(eng, [24, 6], [1500.0, -100.0], [0.0, 0.0], [777, 5])
        rec, prof = t.process(12, [round(k * n / fs, 6) for k in range(12)], want_profiles=True)
        assert rec["symbol"].shape == (2, 12)
        ts = np.array([round(k * n / fs, 6) for k in range(12)])
        for _ in range(9):  # > 80 symbols: bit-phase search and bit emission of the navigation-bit kernel
            t.process(12, ts)
            bits = t.integrate_bits(12, ts, ts + 0.001)
        assert len(bits) == 2 and t.bit_state(0)["processed_pseudosymbol_count"] == 108
        t.close()
        # pipelined batch stream: three streams, pageable staging
        gs = _native.GridStream(eng, 2, 1, [24, 0, 5], dop, _native.NON_COHERENT, depth=2)
        xb = o.synth_iq(2, n, 2, fs, [(25, 1500.0, 777, 0.3, 0.3)])
        outs = []
        for k in range(4):
            if gs.in_flight == 2:
                outs.append(gs.collect().copy())
            gs.submit(xb)
        while gs.in_flight:
            outs.append(gs.collect().copy())
        assert all(int(u["argmax"][0, 0, 7]) == 777 for u in outs)
        gs.close()
    eng.close()
print("sanitize_small ok")

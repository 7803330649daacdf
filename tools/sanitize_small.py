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
        # round 2: multi-ms one-warp kernel, graph-replayed host grid, best-bin reduction, device ring, channel pool with
        # subset launches + undo, generic-replica kernel
        x10 = o.synth_iq(4, n, 3, fs, [(25, 1500.0, 777, 0.3, 0.3)])
        eng.upload_iq(x10)
        g3 = eng.acquire_grid(1, 3, np.arange(32), dop)
        assert int(g3["argmax"][0, 24, 7]) == 777
        for _ in range(3):
            gh = eng.acquire_grid_host(x10, 1, 3, np.arange(32), dop)
        assert np.array_equal(gh["argmax"], g3["argmax"])
        eng.upload_iq(x10)
        best = eng.acquire_grid_best(1, 3, np.arange(32), dop)
        assert int(best["code_phase"][0, 24]) == 777 and float(best["doppler"][0, 24]) == 1500.0
        ring = _native.Ring(eng, 4)
        for k in range(6):
            ring.append(xs[k * n:(k + 1) * n])
        ring.bind_newest(3)
        pool = _native.Tracker.pool(eng, 4)
        pool.reset_channel(2, 24, 1500.0, 0.0, 777)
        pool.reset_channel(0, 6, -100.0, 0.0, 5)
        r1 = pool.process_channels([2, 0], 1, [0.003], keep_undo=True)
        pool.undo_channel(0)
        r2 = pool.process_channels([0], 1, [0.003])
        assert r1["symbol"][1, 0] == r2["symbol"][0, 0]
        pool.close()
        ring.close()
        eng.upload_iq(x10)
        rng = np.random.default_rng(0)
        odd = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        pr = eng.correlation_profile_replica(odd, 250.0, 2, _native.NON_COHERENT)
        assert pr.shape == (n,)
    eng.close()
print("sanitize_small ok")

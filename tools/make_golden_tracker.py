"""Generate tests/golden/tracker_*.npz by running the LIVE reference tracker (/root/reference/gypsum/tracker.py).
Run:  python tools/make_golden_tracker.py     (takes ~10 s)"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore", category=DeprecationWarning)

from gypsum.antenna_sample_provider import AntennaSampleChunk, SampleProviderAttributes  # noqa: E402
from gypsum.gps_ca_prn_codes import GpsSatelliteId, generate_replica_prn_signals  # noqa: E402
from gypsum.satellite import GpsSatellite  # noqa: E402
from gypsum.tracker import GpsSatelliteTracker, GpsSatelliteTrackingParameters, LostSatelliteLockError  # noqa: E402

from oracle import tracker_oracle as t  # noqa: E402  (synthetic input + timestamps only)

OUT = os.path.join(ROOT, "tests", "golden")
N, FS = 2046, 2046000


def run(name, seed, n_ms, channel, init, sigma=0.02, N=N, FS=FS):
    codes = generate_replica_prn_signals()
    sv = channel[0]
    x = t.synth_tracking_iq(seed, N, n_ms, FS, [channel], sigma)
    sat = GpsSatellite(GpsSatelliteId(sv), codes[GpsSatelliteId(sv)], N // 1023)
    params = GpsSatelliteTrackingParameters(satellite=sat, current_doppler_shift=init[0],
                                            current_carrier_wave_phase_shift=init[1],
                                            current_prn_code_phase_shift=init[2], doppler_shifts=[])
    trk = GpsSatelliteTracker(params, SampleProviderAttributes(FS, N))
    rows, lost_at = [], -1
    for k in range(n_ms):
        t0, t1 = t.chunk_times(k, FS, N)
        try:
            ps = trk.process_samples(AntennaSampleChunk(t0, t1, x[k * N:(k + 1) * N]))
        except LostSatelliteLockError:
            lost_at = k
            break
        pk = params.correlation_peaks_rolling_buffer[-1]
        rows.append([pk.real, pk.imag, params.correlation_peak_strengths_rolling_buffer[-1], ps.pseudosymbol.as_val(),
                     params.carrier_wave_phase_errors[-1], params.discriminators[-2], params.current_doppler_shift,
                     params.current_carrier_wave_phase_shift, params.current_prn_code_phase_shift,
                     ps.start_of_pseudosymbol, ps.end_of_pseudosymbol, trk.phase,
                     # what the reference APPENDS to its histories (tracker.py:352-353): the values before the 6-second
                     # constellation adjustment of :370-387, which only the current_* fields above include
                     params.doppler_shifts[-1], params.carrier_wave_phases[-1]])
    np.savez_compressed(os.path.join(OUT, f"tracker_{name}.npz"), seed=np.int64(seed), n_ms=np.int64(n_ms),
                        channel=np.array(channel, dtype=np.float64), init=np.array(init, dtype=np.float64),
                        sigma=np.float64(sigma), rows=np.array(rows, dtype=np.float64), lost_at=np.int64(lost_at),
                        n=np.int64(N), fs=np.int64(FS))
    r = np.array(rows)
    print(name, "ms", len(rows), "lost_at", lost_at, "final doppler", r[-1, 6], "symbols +/-", (r[:, 3] > 0).sum(), (r[:, 3] < 0).sum())


if __name__ == "__main__":
    # (sv, doppler, doppler rate, code phase, carrier phase, amplitude); init = (doppler, carrier phase, code phase)
    run("short", 11, 700, (25, 1500.3, 0.0, 777, 0.3, 0.004), (1500.0, 0.0, 777))
    run("long", 12, 6300, (7, -2212.7, 0.5, 100, 1.0, 0.005), (-2210.0, 0.5, 100))   # crosses the 6 s circularity check
    run("noise", 13, 6100, (3, 800.0, 0.0, 5, 0.0, 0.0), (800.0, 0.0, 5))            # no signal: loses lock at the check
    # 4.092 Msps: the reference keeps its hard-wired 2046 (tracker.py:301-303, :319) -- SURVEY F12 -- so the code-phase
    # accumulator wraps at 2046 although a millisecond is 4092 samples; the planted phase stays below 2046
    # weak signal: circularity 0.89 at the 6-second check -> the -+5 Hz / +-pi/2 nudge of tracker.py:380-387 fires
    run("adjust", 21, 6100, (9, 432.1, 0.0, 300, 0.4, 0.0016), (430.0, 0.0, 300))
    run("fs4", 14, 500, (12, 640.4, 0.0, 1501, 0.7, 0.004), (640.0, 0.0, 1501), N=4092, FS=4092000)

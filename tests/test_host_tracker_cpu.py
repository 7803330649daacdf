"""Host-side tracking glue without a GPU: GpsSatelliteTracker.process_samples (tracker.py:331-389 as seen by the caller:
pseudosymbol with the code-phase delay, the histories the visualiser reads, host edits of the loop state pushed to the
device, LostSatelliteLockError) through a stand-in for the native channel whose milliseconds are computed by the tracker
oracle, against the trajectories recorded from the live reference.  The stand-in is test infrastructure."""
import os

import numpy as np
import pytest

from gypsum_b200 import _native
from oracle import tracker_oracle as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, FS = 2046, 2046000


class StandInEngine:
    def upload_iq(self, x):
        self.x = np.asarray(x)

    def set_replicas(self, chips):
        self.chips = np.asarray(chips)


class StandInChannel:
    """_native.Tracker for one channel: process(1, [t0]) = one TrackerOracle step on the engine's current chunk."""

    def __init__(self, engine, prn_idx, doppler, carrier_phase, code_phase):
        self.engine = engine
        sv = 1 + int(prn_idx[0])  # row r of the table holds SV r + 1 in these tests
        self.oracle = t.TrackerOracle(sv, doppler[0], carrier_phase[0], int(code_phase[0]), FS, N)
        self.set_calls = []

    def set_state(self, channel, doppler, carrier_phase, phase_acc, code_phase):
        self.set_calls.append((doppler, carrier_phase, phase_acc, code_phase))
        self.oracle.doppler, self.oracle.carrier_phase = doppler, carrier_phase
        self.oracle.phase, self.oracle.code_phase = phase_acc, code_phase

    def process(self, n_ms, start_times, want_profiles=False):
        assert n_ms == 1
        t0 = float(start_times[0])
        lost = 0
        try:
            r = self.oracle.step(self.engine.x, t0, round(t0 + N / FS, 6))
        except t.LostLock as exc:
            r, lost = exc.args[0], 1
            r.update(doppler=self.oracle.doppler, carrier_phase=self.oracle.carrier_phase)
        rec = np.zeros((1, 1), dtype=_native.TRACK_DTYPE)
        for k in ("doppler", "carrier_phase", "error", "disc", "strength", "code_phase", "symbol", "peak_offset"):
            rec[k] = r[k]
        rec["phase_acc"], rec["locked"], rec["lost"] = self.oracle.phase, int(r["locked"]), lost
        rec["peak_re"], rec["peak_im"] = r["peak"].real, r["peak"].imag
        prof = np.zeros((1, 1, N), dtype=np.float32)
        return (rec, prof) if want_profiles else rec


class Attrs:
    samples_per_second, samples_per_prn_transmission = FS, N


class Chunk:
    def __init__(self, k, x):
        self.start_time, self.end_time = t.chunk_times(k, FS, N)
        self.samples = x[k * N:(k + 1) * N]


def _tracker(monkeypatch, sv, init):
    from gypsum_b200 import tracker as trk_mod
    from gypsum_b200 import utils
    from gypsum_b200.gps_ca_prn_codes import GpsSatelliteId, generate_replica_prn_signals
    from gypsum_b200.satellite import GpsSatellite

    ent = {"engine": StandInEngine(), "codes": {}, "table": []}
    monkeypatch.setattr(utils.POOL, "get", lambda fs, n, device=0: ent)
    monkeypatch.setattr(trk_mod._native, "Tracker", StandInChannel)
    codes = generate_replica_prn_signals()
    # register SVs 1..sv so that table row r is SV r + 1 (what the stand-in assumes)
    for s in range(1, sv):
        utils.POOL.replica_index(ent, np.ascontiguousarray(np.asarray(codes[GpsSatelliteId(s)].inner) != 0, dtype=np.uint8))
    params = trk_mod.GpsSatelliteTrackingParameters(
        satellite=GpsSatellite(GpsSatelliteId(sv), codes[GpsSatelliteId(sv)], 2), current_doppler_shift=init[0],
        current_carrier_wave_phase_shift=init[1], current_prn_code_phase_shift=int(init[2]), doppler_shifts=[])
    return trk_mod.GpsSatelliteTracker(params, Attrs, keep_correlation_profiles=True), params, trk_mod


def _case(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"tracker_{name}.npz"))
    ch = z["channel"]
    ch = (int(ch[0]), ch[1], ch[2], int(ch[3]), ch[4], ch[5])
    return z, ch, t.synth_tracking_iq(int(z["seed"]), N, int(z["n_ms"]), FS, [ch], float(z["sigma"]))


def test_process_samples_fills_the_reference_histories(monkeypatch):
    z, ch, x = _case("short")
    trk, params, _ = _tracker(monkeypatch, ch[0], z["init"])
    rows = z["rows"]
    for k in range(300):
        ps = trk.process_samples(Chunk(k, x))
        g = rows[k]
        assert (ps.pseudosymbol.as_val(), ps.start_of_pseudosymbol, ps.end_of_pseudosymbol) == (int(g[3]), g[9], g[10])
        assert (params.current_doppler_shift, params.current_carrier_wave_phase_shift,
                params.current_prn_code_phase_shift, trk.phase) == (g[6], g[7], int(g[8]), g[11])
    # what tracker.py:299-353 appends per millisecond
    assert len(params.doppler_shifts) == len(params.carrier_wave_phases) == len(params.carrier_wave_phase_errors) == 300
    assert len(params.correlation_peaks_rolling_buffer) == len(params.correlation_peak_angles) == 300
    assert len(params.discriminators) == 600 and len(params.non_coherent_correlation_profiles) == 250
    assert params.doppler_shifts[-1] == rows[299, 6] and params.carrier_wave_phase_errors[-1] == rows[299, 4]
    assert abs(params.correlation_peaks_rolling_buffer[-1] - complex(rows[299, 0], rows[299, 1])) <= 1e-6 * abs(rows[299, 0])
    assert trk._native.set_calls == []  # no host edits: nothing pushed


def test_host_edits_of_the_loop_state_reach_the_channel(monkeypatch):
    z, ch, x = _case("short")
    trk, params, _ = _tracker(monkeypatch, ch[0], z["init"])
    trk.process_samples(Chunk(0, x))
    params.current_doppler_shift += 2.5  # e.g. the pipeline's re-acquisition path (pipeline.py:104-147)
    trk.process_samples(Chunk(1, x))
    assert len(trk._native.set_calls) == 1 and trk._native.set_calls[0][0] == pytest.approx(float(z["rows"][0, 6]) + 2.5)
    trk.process_samples(Chunk(2, x))
    assert len(trk._native.set_calls) == 1


def test_lost_lock_surfaces_as_the_reference_exception(monkeypatch):
    z, ch, x = _case("noise")
    trk, params, trk_mod = _tracker(monkeypatch, ch[0], z["init"])
    lost_at = int(z["lost_at"])
    with pytest.raises(trk_mod.LostSatelliteLockError):
        for k in range(lost_at + 1):
            trk.process_samples(Chunk(k, x))
    assert len(params.doppler_shifts) == lost_at + 1  # the failing millisecond's histories were appended first (tracker.py:346-353, :378)

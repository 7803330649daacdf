"""The call surface of receiver.py:85-146 / satellite_signal_processing_pipeline.py:56-77, end to end on the GPU:
file-backed provider -> rolling 10-ms window -> GpsSatelliteDetector -> one GpsSatelliteTracker per detection ->
pseudosymbols -> NavigationBitIntegrator -> navigation bits, compared step for step with the same flow run on the CPU
oracle."""
import numpy as np
import pytest

from oracle import gypsum_oracle as o
from oracle import tracker_oracle as t

pytestmark = pytest.mark.gpu
N, FS = 2046, 2046000


def test_acquire_then_track_like_the_receiver(tmp_path, native_lib):
    from gypsum_b200.acquisition import GpsSatelliteDetector
    from gypsum_b200.antenna_sample_provider import (AntennaSampleProviderBackedByFile, InputFileInfo, NoMoreSamplesError,
                                                     RollingSampleWindow)
    from gypsum_b200.gps_ca_prn_codes import GpsSatelliteId, generate_replica_prn_signals
    from gypsum_b200.navigation_bit_integrator import EmitNavigationBitEvent, NavigationBitIntegrator
    from gypsum_b200.satellite import GpsSatellite
    from gypsum_b200.tracker import (EmittedPseudosymbol, GpsSatelliteTracker, GpsSatelliteTrackingParameters,
                                     NavigationBitPseudosymbol)

    n_ms = 420
    planted = [(25, 1500.3, 0.0, 777, 0.3, 0.004), (7, -2212.7, 0.0, 100, 1.0, 0.005)]
    x = t.synth_tracking_iq(42, N, n_ms + 1, FS, planted)
    path = tmp_path / "recording"
    x.view(np.float32).tofile(path)  # interleaved float32 I,Q -- the reference's recording format

    provider = AntennaSampleProviderBackedByFile(InputFileInfo(path, FS))
    attrs = provider.get_attributes()
    codes = generate_replica_prn_signals()
    satellites = {sid: GpsSatellite(sid, code, attrs.samples_per_prn_transmission // 1023) for sid, code in codes.items()}
    detector = GpsSatelliteDetector(satellites)  # receiver.py:66
    window = RollingSampleWindow(attrs.samples_per_prn_transmission, 10)  # receiver.py:68
    search_for = [GpsSatelliteId(i) for i in (3, 7, 19, 25)]
    trackers, symbols, integrators, bits = {}, {}, {}, {}
    k = 0
    while True:
        try:
            chunk = provider.get_samples(attrs.samples_per_prn_transmission)  # receiver.py:93-95
        except NoMoreSamplesError:
            break
        window.append(chunk.samples)  # receiver.py:100
        if window.is_full() and not trackers:  # receiver.py:148-174 (first scan)
            found = detector.detect_satellites_in_antenna_data(search_for, window.window(), attrs)  # receiver.py:220-224
            for r in found:  # pipeline.py:56-63
                params = GpsSatelliteTrackingParameters(
                    satellite=satellites[r.satellite_id], current_doppler_shift=r.doppler_shift,
                    current_carrier_wave_phase_shift=r.carrier_wave_phase_shift,
                    current_prn_code_phase_shift=r.prn_phase_shift, doppler_shifts=[])
                trackers[r.satellite_id.id] = (GpsSatelliteTracker(params, attrs, keep_correlation_profiles=False), r)
                symbols[r.satellite_id.id] = []
                integrators[r.satellite_id.id] = NavigationBitIntegrator(r.satellite_id)  # pipeline.py:64
                bits[r.satellite_id.id] = []
        for sv, (trk, _) in trackers.items():  # receiver.py:237-257
            ps = trk.process_samples(chunk)  # pipeline.py:77
            symbols[sv].append(ps.pseudosymbol.as_val())
            for ev in integrators[sv].process_pseudosymbol(chunk.start_time, ps):  # pipeline.py:79
                assert isinstance(ev, EmitNavigationBitEvent)
                bits[sv].append((ev.receiver_timestamp, ev.trailing_edge_receiver_timestamp, ev.bit_value))
        k += 1
    assert k == n_ms and sorted(trackers) == [7, 25]

    # ---- the same flow on the CPU oracle ----
    first = x[: 10 * N]
    for sv, (trk, r) in trackers.items():
        ref = o.acquire_sv(sv, first, FS, N)
        assert (ref.doppler, ref.code_phase) == (r.doppler_shift, r.prn_phase_shift)
        assert abs(ref.strength - r.correlation_strength) <= 1e-4 * ref.strength
        d = abs(ref.carrier_phase - r.carrier_wave_phase_shift)
        assert min(d, 2 * np.pi - d) <= 1e-4
        tr = t.TrackerOracle(sv, ref.doppler, ref.carrier_phase, ref.code_phase, FS, N)
        want, want_bits = [], []
        integ = NavigationBitIntegrator(sv)
        for ms in range(9, n_ms):  # tracking starts with the chunk that completed the first window
            a, b = t.chunk_times(ms, FS, N)
            st = tr.step(x[ms * N:(ms + 1) * N], a, b)
            want.append(st["symbol"])
            ps = EmittedPseudosymbol(st["start"], st["end"], NavigationBitPseudosymbol.from_val(st["symbol"]), 0)
            want_bits += [(e.receiver_timestamp, e.trailing_edge_receiver_timestamp, e.bit_value)
                          for e in integ.process_pseudosymbol(a, ps)]
        assert symbols[sv] == want
        # bits: same values, same edges (the code phase is exact, so the timestamps are too); 20 ms each
        assert bits[sv] == want_bits and len(want_bits) >= 15
        assert all(abs((b1 - b0) - 0.020) < 2e-6 for b0, b1, _ in bits[sv])
        # the symbol stream carries 20-ms data bits: long runs of equal symbols, not noise
        tail = np.array(symbols[sv][60:])
        assert np.count_nonzero(np.diff(tail) != 0) <= len(tail) // 8
        assert abs(trk.tracking_params.current_doppler_shift - tr.doppler) <= 5e-3

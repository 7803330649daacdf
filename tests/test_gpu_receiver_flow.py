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


def _setup_flow(tmp_path, n_ms, planted, seed):
    from gypsum_b200.antenna_sample_provider import AntennaSampleProviderBackedByFile, InputFileInfo
    from gypsum_b200.gps_ca_prn_codes import generate_replica_prn_signals
    from gypsum_b200.satellite import GpsSatellite

    x = t.synth_tracking_iq(seed, N, n_ms + 1, FS, planted)
    path = tmp_path / "recording"
    x.view(np.float32).tofile(path)
    provider = AntennaSampleProviderBackedByFile(InputFileInfo(path, FS))
    attrs = provider.get_attributes()
    codes = generate_replica_prn_signals()
    satellites = {sid: GpsSatellite(sid, code, attrs.samples_per_prn_transmission // 1023) for sid, code in codes.items()}
    return x, provider, attrs, satellites


def test_device_ring_feeds_detector_and_trackers(tmp_path, native_lib):
    """every millisecond is uploaded ONCE into the device ring (receiver.py:100); the detector reads its
    10-ms window (receiver.py:219) and all trackers their chunk from the ring in place, the trackers of one chunk in one
    launch.  Same acquisitions and pseudosymbol streams as the oracle flow, and as the host-array flow."""
    from gypsum_b200.acquisition import GpsSatelliteDetector
    from gypsum_b200.antenna_sample_provider import DeviceSampleRing, NoMoreSamplesError
    from gypsum_b200.gps_ca_prn_codes import GpsSatelliteId
    from gypsum_b200.tracker import GpsSatelliteTracker, GpsSatelliteTrackingParameters
    from gypsum_b200.utils import POOL

    n_ms = 150
    planted = [(25, 1500.3, 0.0, 777, 0.3, 0.004), (7, -2212.7, 0.0, 100, 1.0, 0.005), (19, 3100.4, 0.0, 1999, 2.0, 0.005)]
    x, provider, attrs, satellites = _setup_flow(tmp_path, n_ms, planted, 43)
    detector = GpsSatelliteDetector(satellites)
    ring = DeviceSampleRing(attrs, 10)
    eng = POOL.get(FS, N)["engine"]
    launches0 = eng.launch_count
    search_for = [GpsSatelliteId(i) for i in (3, 7, 19, 25)]
    trackers, symbols, acq = {}, {}, {}
    k = 0
    while True:
        try:
            chunk = ring.append(provider.get_samples(attrs.samples_per_prn_transmission))
        except NoMoreSamplesError:
            break
        if ring.is_full() and not trackers:
            for r in detector.detect_satellites_in_antenna_data(search_for, ring.window(), attrs):
                params = GpsSatelliteTrackingParameters(
                    satellite=satellites[r.satellite_id], current_doppler_shift=r.doppler_shift,
                    current_carrier_wave_phase_shift=r.carrier_wave_phase_shift,
                    current_prn_code_phase_shift=r.prn_phase_shift, doppler_shifts=[])
                trackers[r.satellite_id.id] = GpsSatelliteTracker(params, attrs, keep_correlation_profiles=False)
                symbols[r.satellite_id.id], acq[r.satellite_id.id] = [], r
            launches_after_scan = eng.launch_count
        for sv, trk in trackers.items():
            symbols[sv].append(trk.process_samples(chunk).pseudosymbol.as_val())
        k += 1
    assert k == n_ms and sorted(trackers) == [7, 19, 25]
    tracked_ms = n_ms - 9
    assert eng.launch_count - launches_after_scan == tracked_ms  # ONE tracking launch per millisecond for the three channels
    first = x[: 10 * N]
    for sv, trk in trackers.items():
        ref = o.acquire_sv(sv, first, FS, N)
        assert (ref.doppler, ref.code_phase) == (acq[sv].doppler_shift, acq[sv].prn_phase_shift)
        tr = t.TrackerOracle(sv, ref.doppler, ref.carrier_phase, ref.code_phase, FS, N)
        want = [tr.step(x[ms * N:(ms + 1) * N], *t.chunk_times(ms, FS, N))["symbol"] for ms in range(9, n_ms)]
        assert symbols[sv] == want
        trk.close()
    ring.native.close()


def test_detector_call_between_two_trackers_does_not_leak_its_samples(native_lib):
    """The trackers, the detector and the utils helpers of one sample rate share one engine and its single IQ binding.  A
    detector scan (on OTHER samples) between two trackers' process_samples of the same chunk must not make the second
    tracker correlate against the detector's window."""
    from gypsum_b200.acquisition import GpsSatelliteDetector
    from gypsum_b200.antenna_sample_provider import AntennaSampleChunk, SampleProviderAttributes
    from gypsum_b200.gps_ca_prn_codes import GpsSatelliteId, generate_replica_prn_signals
    from gypsum_b200.satellite import GpsSatellite
    from gypsum_b200.tracker import GpsSatelliteTracker, GpsSatelliteTrackingParameters
    from gypsum_b200.utils import IntegrationType, integrate_correlation_with_doppler_shifted_prn

    attrs = SampleProviderAttributes(FS, N)
    chans = [(25, 1500.3, 0.0, 777, 0.3, 0.004), (7, -2212.7, 0.0, 100, 1.0, 0.005)]
    x = t.synth_tracking_iq(5, N, 30, FS, chans)
    other = o.synth_iq(8, N, 10, FS, [(3, 250.0, 9, 0.0, 0.3)])
    codes = generate_replica_prn_signals()
    sats = {sid: GpsSatellite(sid, code, 2) for sid, code in codes.items()}
    det = GpsSatelliteDetector(sats)

    def make(sv, f, p, c):
        params = GpsSatelliteTrackingParameters(satellite=sats[GpsSatelliteId(sv)], current_doppler_shift=f,
                                                current_carrier_wave_phase_shift=p, current_prn_code_phase_shift=c, doppler_shifts=[])
        return GpsSatelliteTracker(params, attrs, keep_correlation_profiles=False)

    ta, tb = make(25, 1500.0, 0.0, 777), make(7, -2210.0, 0.5, 100)
    oa, ob = t.TrackerOracle(25, 1500.0, 0.0, 777, FS, N), t.TrackerOracle(7, -2210.0, 0.5, 100, FS, N)
    for k in range(30):
        a, b = t.chunk_times(k, FS, N)
        chunk = AntennaSampleChunk(a, b, x[k * N:(k + 1) * N])
        sa = ta.process_samples(chunk).pseudosymbol.as_val()
        if k % 3 == 0:
            det.detect_satellites_in_antenna_data([GpsSatelliteId(3)], other, attrs)
        elif k % 3 == 1:
            integrate_correlation_with_doppler_shifted_prn(IntegrationType.NonCoherent, other, attrs, 250.0,
                                                           sats[GpsSatelliteId(3)].prn_as_complex)
        tb._pool.drop_ahead(tb._channel)  # force tracker B to really read the engine's samples for this chunk
        sb = tb.process_samples(chunk).pseudosymbol.as_val()
        assert sa == oa.step(chunk.samples, a, b)["symbol"], k
        assert sb == ob.step(chunk.samples, a, b)["symbol"], k
    ta.close()
    tb.close()


def test_many_drop_in_trackers_equal_the_channel_bank(native_lib):
    """32 GpsSatelliteTracker objects stepped one millisecond at a time (one launch per millisecond through the pool)
    produce exactly the records TrackerBank produces in one launch over the whole block."""
    from gypsum_b200.antenna_sample_provider import AntennaSampleChunk, SampleProviderAttributes
    from gypsum_b200.gps_ca_prn_codes import GpsSatelliteId, generate_replica_prn_signals
    from gypsum_b200.satellite import GpsSatellite
    from gypsum_b200.tracker import GpsSatelliteTracker, GpsSatelliteTrackingParameters, TrackerBank

    attrs = SampleProviderAttributes(FS, N)
    rng = np.random.default_rng(1)
    chans = [(sv, float(rng.integers(-5000, 5000)) + 0.3, 0.0, int(rng.integers(0, N)), float(rng.uniform(0, 6)), 0.004)
             for sv in range(1, 33)]
    n_ms = 60
    x = t.synth_tracking_iq(6, N, n_ms, FS, chans, sigma=0.01)
    codes = generate_replica_prn_signals()
    sats = {sv: GpsSatellite(GpsSatelliteId(sv), codes[GpsSatelliteId(sv)], 2) for sv in range(1, 33)}
    seeds = [(sats[c[0]], round(c[1]), 0.0, c[3]) for c in chans]
    bank = TrackerBank(seeds, attrs)
    tt = np.array([t.chunk_times(k, FS, N) for k in range(n_ms)])
    want = bank.process(x, tt[:, 0])
    trackers = []
    for sat, f, p, c in seeds:
        params = GpsSatelliteTrackingParameters(satellite=sat, current_doppler_shift=f, current_carrier_wave_phase_shift=p,
                                                current_prn_code_phase_shift=c, doppler_shifts=[])
        trackers.append(GpsSatelliteTracker(params, attrs, keep_correlation_profiles=False))
    for k in range(n_ms):
        chunk = AntennaSampleChunk(tt[k, 0], tt[k, 1], x[k * N:(k + 1) * N])
        for i, trk in enumerate(trackers):
            ps = trk.process_samples(chunk)
            assert ps.pseudosymbol.as_val() == want[i, k]["symbol"]
    for i, trk in enumerate(trackers):
        p = trk.tracking_params
        assert (p.current_doppler_shift, p.current_carrier_wave_phase_shift, p.current_prn_code_phase_shift) == (
            want[i, -1]["doppler"], want[i, -1]["carrier_phase"], want[i, -1]["code_phase"])
        assert p.doppler_shifts == list(want[i]["doppler_hist"])
        trk.close()

"""Drop-in for reference gypsum/tracker.py: GpsSatelliteTracker, GpsSatelliteTrackingParameters and the
pseudosymbol types, same names / fields / exception (tracker.py:33-110, :117-155, :206-389).

`GpsSatelliteTracker.process_samples(chunk)` keeps the reference's one-millisecond-per-call contract (one GPU
round trip per call).  `TrackerBank` is the throughput interface: many channels x many milliseconds in one
persistent-kernel launch.  The correlators, loop filters, lock heuristics and the 6-second constellation check
all run on the device (gypsum_b200/csrc/tracker.cu, tracker_core.cuh); this module only mirrors the host-visible
state and histories the rest of gypsum reads.
"""
from __future__ import annotations

import collections
from dataclasses import dataclass
from enum import Enum, auto

import numpy as np

from gypsum_b200 import _native
from gypsum_b200.constants import ONE_MILLISECOND
from gypsum_b200.utils import POOL, chips_of_replica


class LostSatelliteLockError(Exception):  # tracker.py:33
    pass


class BitValue(Enum):  # tracker.py:48-84
    UNKNOWN = auto()
    ZERO = auto()
    ONE = auto()

    @classmethod
    def from_val(cls, val: int) -> "BitValue":
        return {0: BitValue.ZERO, 1: BitValue.ONE}[val]

    def as_val(self) -> int:
        if self == BitValue.UNKNOWN:
            raise ValueError("Cannot convert an unknown bit value into an integer")
        return {BitValue.ZERO: 0, BitValue.ONE: 1}[self]

    def inverted(self) -> "BitValue":
        if self == BitValue.UNKNOWN:
            raise ValueError("Cannot invert an unknown bit value")
        return {BitValue.ZERO: BitValue.ONE, BitValue.ONE: BitValue.ZERO}[self]

    def __eq__(self, other) -> bool:
        return isinstance(other, BitValue) and self.value == other.value

    def __hash__(self) -> int:
        return hash(self.value)


class NavigationBitPseudosymbol(Enum):  # tracker.py:87-102
    MINUS_ONE = auto()
    ONE = auto()

    @classmethod
    def from_val(cls, val: int) -> "NavigationBitPseudosymbol":
        return {-1: NavigationBitPseudosymbol.MINUS_ONE, 1: NavigationBitPseudosymbol.ONE}[val]

    def as_val(self) -> int:
        return {NavigationBitPseudosymbol.MINUS_ONE: -1, NavigationBitPseudosymbol.ONE: 1}[self]


@dataclass
class EmittedPseudosymbol:  # tracker.py:105-110
    start_of_pseudosymbol: float
    end_of_pseudosymbol: float
    pseudosymbol: NavigationBitPseudosymbol
    cursor_at_emit_time: int


_TRACKER_ITERATIONS_PER_SECOND = 1000  # tracker.py:114


@dataclass
class GpsSatelliteTrackingParameters:
    """tracker.py:117-155: current loop state + the rolling histories the visualiser reads.  `is_locked` reports the
    decision the device made for the most recent millisecond (tracker.py:157-203 runs on the GPU)."""

    satellite: object
    current_doppler_shift: float
    current_carrier_wave_phase_shift: float
    current_prn_code_phase_shift: int
    doppler_shifts: list
    carrier_wave_phases: collections.deque = None
    carrier_wave_phase_errors: collections.deque = None
    correlation_peaks_rolling_buffer: collections.deque = None
    correlation_peak_angles: collections.deque = None
    non_coherent_correlation_profiles: collections.deque = None
    discriminators: collections.deque = None

    def __post_init__(self) -> None:
        for field in (self.correlation_peaks_rolling_buffer, self.correlation_peak_angles, self.carrier_wave_phases,
                      self.carrier_wave_phase_errors):
            if field is not None:
                raise RuntimeError("This field is not intended to be initialized at a call site.")  # tracker.py:145
        n = _TRACKER_ITERATIONS_PER_SECOND
        self.correlation_peaks_rolling_buffer = collections.deque(maxlen=n)
        self.correlation_peak_strengths_rolling_buffer = collections.deque(maxlen=n)
        self.correlation_peak_angles = collections.deque(maxlen=n)
        self.carrier_wave_phases = collections.deque(maxlen=n * 5)
        self.carrier_wave_phase_errors = collections.deque(maxlen=n * 5)
        self.non_coherent_correlation_profiles = collections.deque(maxlen=n // 4)
        self.discriminators = collections.deque(maxlen=n)
        self._last_is_locked = False

    def is_locked(self) -> bool:
        return self._last_is_locked


def _replica_index(ent, satellite, n: int) -> int:
    code = getattr(getattr(satellite, "prn_code", None), "inner", None)
    if code is not None and getattr(satellite, "scale_factor", n // 1023) == n // 1023:
        chips = np.ascontiguousarray(np.asarray(code) != 0, dtype=np.uint8)
    else:
        chips, roll = chips_of_replica(satellite.prn_as_complex, n)
        if roll:
            raise ValueError("satellite replica must not be rolled")
    return POOL.replica_index(ent, chips)


def _apply_record(params: GpsSatelliteTrackingParameters, rec, profile=None) -> None:
    """What tracker.py:299-353 appends / assigns during one process_samples call."""
    peak = complex(float(rec["peak_re"]), float(rec["peak_im"]))
    params.current_prn_code_phase_shift = int(rec["code_phase"])
    params.discriminators.append(float(rec["disc"]))
    params.discriminators.append(0)  # tracker.py:305 self.accumulator
    if profile is not None:
        params.non_coherent_correlation_profiles.append(profile)
    params.correlation_peaks_rolling_buffer.append(peak)
    params.correlation_peak_strengths_rolling_buffer.append(float(rec["strength"]))
    params.current_carrier_wave_phase_shift = float(rec["carrier_phase"])
    params.current_doppler_shift = float(rec["doppler"])
    params.carrier_wave_phase_errors.append(float(rec["error"]))
    params.correlation_peak_angles.append(float(np.angle(peak)))
    params.doppler_shifts.append(params.current_doppler_shift)
    params.carrier_wave_phases.append(params.current_carrier_wave_phase_shift)
    params._last_is_locked = bool(rec["locked"])


def _pseudosymbol(rec, start_time: float, end_time: float) -> EmittedPseudosymbol:
    delay = (int(rec["code_phase"]) / 2046) * ONE_MILLISECOND  # tracker.py:319
    return EmittedPseudosymbol(
        start_of_pseudosymbol=start_time + delay, end_of_pseudosymbol=end_time + delay,
        pseudosymbol=NavigationBitPseudosymbol.from_val(int(rec["symbol"])), cursor_at_emit_time=0)


class GpsSatelliteTracker:
    def __init__(self, tracking_params: GpsSatelliteTrackingParameters, stream_attributes,
                 keep_correlation_profiles: bool = True) -> None:
        self.tracking_params = tracking_params
        self.stream_attributes = stream_attributes
        self.keep_correlation_profiles = keep_correlation_profiles
        self.accumulator = 0
        self.phase = tracking_params.current_prn_code_phase_shift  # tracker.py:224
        fs, n = int(stream_attributes.samples_per_second), int(stream_attributes.samples_per_prn_transmission)
        self._ent = POOL.get(fs, n)
        self._eng = self._ent["engine"]
        idx = _replica_index(self._ent, tracking_params.satellite, n)
        self._native = _native.Tracker(self._eng, [idx], [tracking_params.current_doppler_shift],
                                       [tracking_params.current_carrier_wave_phase_shift],
                                       [tracking_params.current_prn_code_phase_shift])
        self._device_view = (float(tracking_params.current_doppler_shift),
                             float(tracking_params.current_carrier_wave_phase_shift),
                             int(tracking_params.current_prn_code_phase_shift), float(self.phase))

    def _push_host_edits(self) -> None:
        """If a caller changed tracking_params.current_* (or self.phase) since the last call, the device follows."""
        p = self.tracking_params
        now = (float(p.current_doppler_shift), float(p.current_carrier_wave_phase_shift),
               int(p.current_prn_code_phase_shift), float(self.phase))
        if now != self._device_view:
            self._native.set_state(0, now[0], now[1], now[3], now[2])
            self._device_view = now

    def process_samples(self, receiver_samples_chunk) -> EmittedPseudosymbol:
        """tracker.py:331-389."""
        self._push_host_edits()
        samples = receiver_samples_chunk.samples
        key = (id(samples), float(receiver_samples_chunk.start_time))
        if self._eng.iq_tag != key:  # several trackers usually share one chunk: upload it once.  The tag lives in the
            self._eng.upload_iq(samples, tag=key)  # engine wrapper and every other upload / bind clears it
        got = self._native.process(1, [receiver_samples_chunk.start_time], want_profiles=self.keep_correlation_profiles)
        rec, prof = (got[0][0, 0], got[1][0, 0]) if self.keep_correlation_profiles else (got[0, 0], None)
        if int(rec["symbol"]) == 0:
            raise KeyError(0)  # tracker.py:317: NavigationBitPseudosymbol.from_val has no entry for 0
        _apply_record(self.tracking_params, rec, None if prof is None else prof.astype(np.float64))
        self.phase = float(rec["phase_acc"])
        p = self.tracking_params
        self._device_view = (float(p.current_doppler_shift), float(p.current_carrier_wave_phase_shift),
                             int(p.current_prn_code_phase_shift), float(self.phase))
        if int(rec["lost"]):
            raise LostSatelliteLockError()  # tracker.py:378
        return _pseudosymbol(rec, receiver_samples_chunk.start_time, receiver_samples_chunk.end_time)


class TrackerBank:
    """Throughput interface: `n` channels advance through a block of milliseconds in one persistent-kernel launch
    (BASELINE config 4).  channels: iterable of (satellite, doppler_hz, carrier_phase_rad, code_phase_samples)."""

    def __init__(self, channels, stream_attributes, device: int = 0):
        fs, n = int(stream_attributes.samples_per_second), int(stream_attributes.samples_per_prn_transmission)
        self.samples_per_ms = n
        self._ent = POOL.get(fs, n, device)
        self.engine = self._ent["engine"]
        channels = list(channels)
        idx = [_replica_index(self._ent, c[0], n) for c in channels]
        self.native = _native.Tracker(self.engine, idx, [c[1] for c in channels], [c[2] for c in channels],
                                      [c[3] for c in channels])
        self.n_channels = len(channels)

    def process(self, samples: np.ndarray, start_times, want_profiles: bool = False):
        """samples: complex64[n_ms * N]; returns TRACK_DTYPE records [n_channels, n_ms] (and profiles)."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        n_ms = x.size // self.samples_per_ms
        self.engine.upload_iq(x[: n_ms * self.samples_per_ms])
        return self.native.process(n_ms, start_times, want_profiles)

    def integrate_bits(self, start_times, end_times) -> list:
        """Navigation bits of every channel from the records the last `process` call left on the device
        (navigation_bit_intergrator.py:278-288; one integrator per channel, persistent across calls).  Returns one
        _native.BIT_DTYPE array per channel: timestamps of the bit's edges, value 1 / 0 / -1 (unknown)."""
        return self.native.integrate_bits(len(start_times), start_times, end_times)

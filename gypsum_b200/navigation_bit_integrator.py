"""Drop-in for reference gypsum/navigation_bit_intergrator.py: 20 pseudosymbols -> one navigation
bit, with the reference's bit-phase search and resynchronisation rules.

This is the step AFTER the tracking path: 50 Hz integer logic on the +-1 stream the tracker emits.  It stays on the
host (there is nothing to accelerate: 20 integer adds per bit); what the GPU side contributes is that `TrackerBank`
delivers the symbols in bulk.  Same class / method / event names as the reference; the 20-phase confidence search
(:127-147) is vectorised with numpy, everything else follows the reference's bookkeeping exactly so that the emitted
bit stream is identical (tests/test_bit_integrator.py compares with events recorded from the live reference)."""
from __future__ import annotations

import collections

import numpy as np

from gypsum_b200.tracker import BitValue, EmittedPseudosymbol

PSEUDOSYMBOLS_PER_NAVIGATION_BIT = 20  # constants.py:24
BITS_PER_SECOND = 50  # constants.py:22
PSEUDOSYMBOLS_PER_SECOND = PSEUDOSYMBOLS_PER_NAVIGATION_BIT * BITS_PER_SECOND
RECALCULATE_PSEUDOSYMBOL_PHASE_PERIOD = 1  # config.py:38
RECALCULATE_PSEUDOSYMBOL_PHASE_BIT_HEALTH_MEMORY_SIZE = 10  # config.py:41
RECALCULATE_PSEUDOSYMBOL_PHASE_BIT_HEALTH_THRESHOLD = 50  # config.py:43


class Event:  # events.py
    pass


class EmitNavigationBitEvent(Event):  # navigation_bit_intergrator.py:29-39
    def __init__(self, receiver_timestamp, trailing_edge_receiver_timestamp, bit_value: BitValue) -> None:
        self.receiver_timestamp = receiver_timestamp
        self.trailing_edge_receiver_timestamp = trailing_edge_receiver_timestamp
        self.bit_value = bit_value


class CannotDetermineBitPhaseEvent(Event):  # :42-44
    def __init__(self, confidence: float) -> None:
        self.confidence = confidence


class LostBitCoherenceEvent(Event):  # :47-49
    def __init__(self, confidence: float) -> None:
        self.confidence = confidence


class LostBitPhaseCoherenceError(Exception):  # :52-53
    pass


class NavigationBitIntegratorHistory:
    """State of :56-103 (same attribute names: the reference's visualiser reads them)."""

    def __init__(self) -> None:
        self.last_seen_pseudosymbols = collections.deque(maxlen=1000)
        self.last_emitted_bits = collections.deque(maxlen=BITS_PER_SECOND)
        self.previous_bit_phase_decision = None
        self.determined_bit_phase = None
        self.failed_bit_count = 0
        self.emitted_bit_count = 0
        self.processed_pseudosymbol_count = 0
        self.sequential_unknown_bit_value_counter = 0
        self.queued_pseudosymbols: list[EmittedPseudosymbol] = []
        self.pseudosymbol_cursor_within_queue = 0
        self.rolling_average_window_size = PSEUDOSYMBOLS_PER_NAVIGATION_BIT // 2
        self.rolling_average_window = collections.deque(maxlen=self.rolling_average_window_size)


class NavigationBitIntegrator:
    def __init__(self, satellite_id) -> None:
        self.satellite_id = satellite_id
        self.history = NavigationBitIntegratorHistory()
        self.pseudosymbol_count_to_use_for_bit_phase_selection = PSEUDOSYMBOLS_PER_NAVIGATION_BIT * 4
        self.resynchronize_bit_phase_period = PSEUDOSYMBOLS_PER_SECOND * RECALCULATE_PSEUDOSYMBOL_PHASE_PERIOD
        self.resynchronize_bit_phase_memory_size = RECALCULATE_PSEUDOSYMBOL_PHASE_BIT_HEALTH_MEMORY_SIZE
        self.slide = 0

    # -- bit phase ------------------------------------------------------------------------------------------------
    def _redetermine_bit_phase(self):
        """:127-147: over the last 16 bits' worth of symbols, the phase whose 20-symbol blocks agree most; the first
        phase wins ties (dict insertion order under max())."""
        h = self.history
        if len(h.last_seen_pseudosymbols) < self.pseudosymbol_count_to_use_for_bit_phase_selection:
            return None
        recent = list(h.last_seen_pseudosymbols)[-PSEUDOSYMBOLS_PER_NAVIGATION_BIT * 16:]
        values = np.array([s.pseudosymbol.as_val() for s in recent], dtype=np.int64)
        whole = (values.size // PSEUDOSYMBOLS_PER_NAVIGATION_BIT) * PSEUDOSYMBOLS_PER_NAVIGATION_BIT
        best_phase, best_score = 0, None
        for phase in range(PSEUDOSYMBOLS_PER_NAVIGATION_BIT):
            rolled = np.roll(values, -phase)
            sums = rolled[:whole].reshape(-1, PSEUDOSYMBOLS_PER_NAVIGATION_BIT).sum(axis=1)
            # :105-125: mean |block sum| per bit, normalised by the block length
            score = (int(np.abs(sums).sum()) / (values.size / PSEUDOSYMBOLS_PER_NAVIGATION_BIT)) / PSEUDOSYMBOLS_PER_NAVIGATION_BIT
            if best_score is None or score > best_score:
                best_phase, best_score = phase, score
        return best_phase

    def _should_resynchronize_bit_phase(self) -> bool:
        """:217-246."""
        h = self.history
        if h.processed_pseudosymbol_count % self.resynchronize_bit_phase_period == 0:
            return True
        if h.processed_pseudosymbol_count % PSEUDOSYMBOLS_PER_NAVIGATION_BIT != 0:
            return False
        if h.previous_bit_phase_decision is None:
            return True
        recent = list(h.last_emitted_bits)[-self.resynchronize_bit_phase_memory_size:]
        if len(recent) == self.resynchronize_bit_phase_memory_size:
            unknown = sum(1 for b in recent if b == BitValue.UNKNOWN)
            if (unknown / len(recent)) * 100 >= RECALCULATE_PSEUDOSYMBOL_PHASE_BIT_HEALTH_THRESHOLD:
                return True
        return False

    def _resynchronize_bit_phase_if_necessary(self) -> list:
        """:248-276."""
        if not self._should_resynchronize_bit_phase():
            return []
        h = self.history
        before = h.previous_bit_phase_decision
        after = self._redetermine_bit_phase()
        h.previous_bit_phase_decision = after
        h.determined_bit_phase = after
        if before is None and after is not None:
            if after > 0:
                h.pseudosymbol_cursor_within_queue = after
                self.slide = after
        elif before is not None and after is not None and before != after:
            self.slide += after - before
            h.pseudosymbol_cursor_within_queue += after - before
        return []

    def _reset_selected_bit_phase(self) -> None:
        self.history.determined_bit_phase = None  # :112-114

    # -- bits -------------------------------------------------------------------------------------------------------
    @staticmethod
    def _get_bit_value_from_pseudosymbols(pseudosymbols) -> BitValue:
        """:149-161: sign of the sum; unresolved when |mean| <= 50 %."""
        total = sum(s.pseudosymbol.as_val() for s in pseudosymbols)
        value = BitValue.ONE if total > 0 else BitValue.ZERO
        if abs(int((total / len(pseudosymbols)) * 100)) <= 50:
            value = BitValue.UNKNOWN
        return value

    def _emit_bit_from_pseudosymbols(self, pseudosymbols) -> EmitNavigationBitEvent:
        """:163-192."""
        h = self.history
        value = self._get_bit_value_from_pseudosymbols(pseudosymbols)
        h.last_emitted_bits.append(value)
        if value == BitValue.UNKNOWN:
            h.sequential_unknown_bit_value_counter += 1
            h.failed_bit_count += 1
            if h.sequential_unknown_bit_value_counter >= 30:
                self._reset_selected_bit_phase()
        else:
            h.sequential_unknown_bit_value_counter = 0
        return EmitNavigationBitEvent(
            receiver_timestamp=pseudosymbols[0].start_of_pseudosymbol,
            trailing_edge_receiver_timestamp=pseudosymbols[-1].end_of_pseudosymbol,
            bit_value=value,
        )

    def _emit_bits_from_queued_pseudosymbols(self) -> list:
        """:194-215."""
        h = self.history
        if h.determined_bit_phase is None:
            return []
        events = []
        pending = h.queued_pseudosymbols[h.pseudosymbol_cursor_within_queue:]
        n = PSEUDOSYMBOLS_PER_NAVIGATION_BIT
        for i in range(0, len(pending) - n + 1, n):  # whole bits only (utils.py:28-38)
            events.append(self._emit_bit_from_pseudosymbols(pending[i:i + n]))
            h.pseudosymbol_cursor_within_queue += n
            h.emitted_bit_count += 1
        if len(h.queued_pseudosymbols) >= n:
            unread = len(h.queued_pseudosymbols) - h.pseudosymbol_cursor_within_queue
            h.queued_pseudosymbols = h.queued_pseudosymbols[-n:]
            h.pseudosymbol_cursor_within_queue = n - unread
        return events

    def process_pseudosymbol(self, receiver_timestamp, pseudosymbol: EmittedPseudosymbol) -> list:
        """:278-288."""
        h = self.history
        pseudosymbol.cursor_at_emit_time = self.slide
        h.queued_pseudosymbols.append(pseudosymbol)
        h.last_seen_pseudosymbols.append(pseudosymbol)
        if receiver_timestamp < 40:
            self._resynchronize_bit_phase_if_necessary()
        events = list(self._emit_bits_from_queued_pseudosymbols())
        h.processed_pseudosymbol_count += 1
        return events

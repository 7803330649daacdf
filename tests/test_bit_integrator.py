"""gypsum_b200.navigation_bit_integrator against events recorded from the live reference integrator
(tools/make_golden_bits.py): identical bit stream, timestamps and final bookkeeping."""
import os

import numpy as np
import pytest

from gypsum_b200.navigation_bit_integrator import EmitNavigationBitEvent, NavigationBitIntegrator
from gypsum_b200.tracker import BitValue, EmittedPseudosymbol, NavigationBitPseudosymbol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = {BitValue.ONE: 1, BitValue.ZERO: 0, BitValue.UNKNOWN: -1}


@pytest.mark.parametrize("case", ["long", "synth"])
def test_bit_stream_identical_to_reference(case):
    z = np.load(os.path.join(ROOT, "tests", "golden", "bit_integrator.npz"))
    integ = NavigationBitIntegrator(7)
    rows = []
    for k, (s, a, b) in enumerate(zip(z[f"{case}_symbols"], z[f"{case}_starts"], z[f"{case}_ends"])):
        ps = EmittedPseudosymbol(a, b, NavigationBitPseudosymbol.from_val(int(s)), 0)
        for ev in integ.process_pseudosymbol(a, ps):
            assert isinstance(ev, EmitNavigationBitEvent)
            rows.append([k, ev.receiver_timestamp, ev.trailing_edge_receiver_timestamp, CODE[ev.bit_value]])
    assert np.array_equal(np.array(rows, dtype=np.float64), z[f"{case}_events"])
    h = integ.history
    final = [h.emitted_bit_count, h.failed_bit_count, h.processed_pseudosymbol_count, integ.slide,
             -1 if h.determined_bit_phase is None else h.determined_bit_phase,
             -1 if h.previous_bit_phase_decision is None else h.previous_bit_phase_decision, h.pseudosymbol_cursor_within_queue]
    assert final == list(z[f"{case}_final"])


BIT_EVENT = np.dtype([("t0", "<f8"), ("t1", "<f8"), ("k", "<i4"), ("bit", "<i4"), ("slide", "<i4"), ("pad", "<i4")])


def _emu_run(emu, symbols, starts, ends, splits):
    """bits_core.cuh (the device state machine compiled for the host) over a stream cut at `splits`."""
    import ctypes as C

    emu.emu_bit_run.restype = C.c_int
    st = (C.c_char * emu.emu_bit_state_size())()
    emu.emu_bit_init(st)
    rows = []
    edges = [0, *splits, len(symbols)]
    for a, b in zip(edges[:-1], edges[1:]):
        n = b - a
        sym = np.ascontiguousarray(symbols[a:b], dtype=np.int32)
        t0 = np.ascontiguousarray(starts[a:b])
        t1 = np.ascontiguousarray(ends[a:b])
        ev = np.zeros(n // 20 + 8, dtype=BIT_EVENT)
        cnt = emu.emu_bit_run(st, n, sym.ctypes.data_as(C.c_void_p), t0.ctypes.data_as(C.c_void_p), t0.ctypes.data_as(C.c_void_p),
                              t1.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), ev.size)
        assert cnt <= ev.size
        rows += [[a + e["k"], e["t0"], e["t1"], e["bit"]] for e in ev[:cnt]]
    summary = np.zeros(8, dtype=np.int64)
    emu.emu_bit_summary(st, summary.ctypes.data_as(C.c_void_p))
    return np.array(rows, dtype=np.float64), summary


@pytest.mark.parametrize("case,splits", [("long", []), ("long", [1, 79, 80, 81, 1000, 1013]), ("synth", []),
                                         ("synth", list(range(777, 45000, 777)))])
def test_device_state_machine_on_host(emu_lib, case, splits):
    """The same streams through the state machine the GPU runs (one call, and cut into ragged calls: the state carries
    over exactly)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "bit_integrator.npz"))
    rows, summary = _emu_run(emu_lib, z[f"{case}_symbols"], z[f"{case}_starts"], z[f"{case}_ends"], splits)
    assert np.array_equal(rows, z[f"{case}_events"])
    assert list(summary[:7]) == list(z[f"{case}_final"])
    assert summary[7] == 0  # queue never overflowed


def test_state_machine_after_the_resync_horizon(emu_lib):
    """Symbols that start after 40 s of receiver time never get a bit phase (:283): no bits, bounded state."""
    n = 3000
    sym = np.ones(n)
    t0 = 41.0 + np.arange(n) * 0.001
    rows, summary = _emu_run(emu_lib, sym, t0, t0 + 0.001, [])
    assert rows.size == 0 and summary[0] == 0 and summary[2] == n and summary[7] == 1
    integ = NavigationBitIntegrator(1)
    assert all(not integ.process_pseudosymbol(a, EmittedPseudosymbol(a, a + 0.001, NavigationBitPseudosymbol.ONE, 0)) for a in t0)

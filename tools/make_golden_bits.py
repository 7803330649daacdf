"""Generate tests/golden/bit_integrator.npz by running the LIVE reference NavigationBitIntegrator
(/root/reference/gypsum/navigation_bit_intergrator.py) on recorded / synthetic pseudosymbol streams.
Run: python tools/make_golden_bits.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")

from gypsum.gps_ca_prn_codes import GpsSatelliteId  # noqa: E402
from gypsum.navigation_bit_intergrator import EmitNavigationBitEvent, NavigationBitIntegrator  # noqa: E402
from gypsum.tracker import BitValue, EmittedPseudosymbol, NavigationBitPseudosymbol  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CODE = {BitValue.ONE: 1, BitValue.ZERO: 0, BitValue.UNKNOWN: -1}


def run(symbols, starts, ends):
    integ = NavigationBitIntegrator(GpsSatelliteId(7))
    rows = []
    for k, (s, a, b) in enumerate(zip(symbols, starts, ends)):
        ps = EmittedPseudosymbol(start_of_pseudosymbol=a, end_of_pseudosymbol=b,
                                 pseudosymbol=NavigationBitPseudosymbol.from_val(int(s)), cursor_at_emit_time=0)
        for ev in integ.process_pseudosymbol(a, ps):
            assert isinstance(ev, EmitNavigationBitEvent)
            rows.append([k, ev.receiver_timestamp, ev.trailing_edge_receiver_timestamp, CODE[ev.bit_value]])
    h = integ.history
    final = [h.emitted_bit_count, h.failed_bit_count, h.processed_pseudosymbol_count, integ.slide,
             -1 if h.determined_bit_phase is None else h.determined_bit_phase,
             -1 if h.previous_bit_phase_decision is None else h.previous_bit_phase_decision, h.pseudosymbol_cursor_within_queue]
    return np.array(rows, dtype=np.float64), np.array(final, dtype=np.int64)


def main():
    out = {}
    # 1. the recorded tracker trajectory (tests/golden/tracker_long.npz): real pull-in, flips, data bits
    z = np.load(os.path.join(OUT, "tracker_long.npz"))
    rows = z["rows"]
    out["long_symbols"], out["long_starts"], out["long_ends"] = rows[:, 3], rows[:, 9], rows[:, 10]
    out["long_events"], out["long_final"] = run(rows[:, 3], rows[:, 9], rows[:, 10])
    # 2. synthetic streams: clean bits at phase 13; a noisy stretch that produces UNKNOWN bits and a phase change
    rng = np.random.default_rng(5)
    n = 45000  # crosses the receiver_timestamp < 40 s rule of :283
    bits = rng.integers(0, 2, n // 20 + 2) * 2 - 1
    sym = np.repeat(bits, 20)[13:13 + n].astype(np.float64)
    flip = rng.random(n) < 0.04
    flip[5000:9000] = rng.random(4000) < 0.45          # a bad stretch: unresolved bits, resynchronisation
    sym[9000:] = np.roll(sym, 7)[9000:]                 # the bit phase moves by 7 symbols afterwards
    sym[flip] *= -1
    starts = np.round(np.arange(n) * 0.001, 6) + 0.000379
    ends = starts + 0.001
    out["synth_symbols"], out["synth_starts"], out["synth_ends"] = sym, starts, ends
    out["synth_events"], out["synth_final"] = run(sym, starts, ends)
    np.savez_compressed(os.path.join(OUT, "bit_integrator.npz"), **out)
    for k in ("long", "synth"):
        ev = out[f"{k}_events"]
        print(k, "bits", len(ev), "unknown", int((ev[:, 3] < 0).sum()), "final", out[f"{k}_final"])


if __name__ == "__main__":
    main()

"""CPU oracle (numpy, float64/complex128) for the gypsum correlation hot path.

TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; gypsum_b200 never does.

Every function restates, in its own words, the arithmetic of one reference function and cites it
(paths are relative to /root/reference).  The FFT / exp / abs live in numpy (pocketfft), a third-party
dependency the reference pins as numpy==1.26.0 (requirements.txt:25); this container has numpy 2.3.5.
Both run the transforms in complex128 for this path (utils.py:97 promotes the chunk), so the precision
policy is the same.

PINNING.  The reference carries no tests or golden vectors for this path; the only
known-answer material is the IS-GPS-200 first-10-chip octal table (gps_ca_prn_codes.py:192-225), which
`ca_code()` is checked against in tests/test_oracle_golden.py.  Everything else is pinned against
outputs of the live reference itself, generated in the build container by tools/make_golden.py
(which imports /root/reference) and committed under tests/golden/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

PRN_CHIP_COUNT = 1023  # constants.py:7
ACQUISITION_INTEGRATION_PERIOD_MS = 10  # config.py:4
DETECTION_THRESHOLD = 3  # config.py:7
COHERENT = "coherent"
NON_COHERENT = "non_coherent"

# G2 output tap pairs for SV1..SV32, IS-GPS-200 table 3-Ia (gps_ca_prn_codes.py:145-178).
G2_TAPS = [
    (2, 6), (3, 7), (4, 8), (5, 9), (1, 9), (2, 10), (1, 8), (2, 9), (3, 10), (2, 3), (3, 4), (5, 6), (6, 7),
    (7, 8), (8, 9), (9, 10), (1, 4), (2, 5), (3, 6), (4, 7), (5, 8), (6, 9), (1, 3), (4, 6), (5, 7), (6, 8),
    (7, 9), (8, 10), (1, 6), (2, 7), (3, 8), (4, 9),
]
# First ten chips of each code in octal, IS-GPS-200 (gps_ca_prn_codes.py:192-225).
FIRST_10_CHIPS_OCTAL = [
    1440, 1620, 1710, 1744, 1133, 1455, 1131, 1454, 1626, 1504, 1642, 1750, 1764, 1772, 1775, 1776, 1156, 1467,
    1633, 1715, 1746, 1763, 1063, 1706, 1743, 1761, 1770, 1774, 1127, 1453, 1625, 1712,
]


def ca_code(sv: int) -> np.ndarray:
    """1023 chips in {0,1} for SV `sv` (1-based).

    gps_ca_prn_codes.py:100-131: two 10-stage LFSRs initialised to all ones; G1 feeds back stages 3^10 and
    outputs stage 10; G2 feeds back 2^3^6^8^9^10 and outputs the xor of the SV's two tap stages; chip = G1^G2.
    """
    t1, t2 = G2_TAPS[sv - 1]
    g1 = [1] * 10
    g2 = [1] * 10
    chips = np.empty(PRN_CHIP_COUNT, dtype=np.int64)
    for n in range(PRN_CHIP_COUNT):
        chips[n] = g1[9] ^ g2[t1 - 1] ^ g2[t2 - 1]
        f1 = g1[2] ^ g1[9]
        f2 = g2[1] ^ g2[2] ^ g2[5] ^ g2[7] ^ g2[8] ^ g2[9]
        g1 = [f1] + g1[:9]
        g2 = [f2] + g2[:9]
    return chips


def replica(sv: int, samples_per_ms: int) -> np.ndarray:
    """satellite.py:20-31: chips repeated N/1023 times, {0,1}->{-1,+1}, complex128."""
    scale = samples_per_ms // PRN_CHIP_COUNT
    return (2.0 * np.repeat(ca_code(sv), scale) - 1.0).astype(complex)


def correlate_1ms(samples: np.ndarray, prn: np.ndarray) -> np.ndarray:
    """utils.py:59-73: ifft(fft(x) * conj(fft(prn))) -- circular cross-correlation, raw (un-normalised) sums."""
    return np.fft.ifft(np.fft.fft(samples) * np.conj(np.fft.fft(prn)))


def integrate(kind: str, data: np.ndarray, fs: int, n: int, doppler: float, prn: np.ndarray) -> np.ndarray:
    """utils.py:77-108: per whole 1-ms chunk i, wipe off exp(-j*tau*f*t), t=(arange(N)/fs + i*N/fs), correlate,
    then sum the complex result (coherent) or its magnitude (non-coherent).  A trailing partial chunk is
    dropped (utils.py:34-38)."""
    out = np.zeros(n, dtype=complex if kind == COHERENT else np.float64)
    for i in range(len(data) // n):
        t = (np.arange(n) / fs) + ((i * n) / fs)
        carrier = np.exp(-1j * math.tau * doppler * t)
        c = correlate_1ms(data[i * n:(i + 1) * n] * carrier, prn)
        if kind == COHERENT:
            out += c
        elif kind == NON_COHERENT:
            out += np.abs(c)
        else:
            raise ValueError("Unexpected integration type")  # utils.py:106
    return out


def peak_strength(profile: np.ndarray) -> float:
    """utils.py:111-116: max / mean of every element that is not equal to the max."""
    m = np.max(profile)
    return m / np.mean(profile[profile != m])


def doppler_bins(center: float, spread: float) -> list[int]:
    """acquisition.py:163-167: range(int(c-s), int(c+s), int(s/10)) -- truncation toward zero, upper end open."""
    return list(range(int(center - spread), int(center + spread), int(spread / 10)))


@dataclass
class BestBin:
    doppler: int
    profile: np.ndarray
    peak_index: int
    strength: float


@dataclass
class Acquisition:
    sv: int
    doppler: int
    carrier_phase: float
    code_phase: int
    strength: float


def best_bin(center: float, spread: float, data: np.ndarray, fs: int, n: int, prn: np.ndarray, trace=None) -> BestBin:
    """acquisition.py:154-190: non-coherent profile per bin; the winner is the FIRST bin with the largest
    profile maximum (python max over dict insertion order); argmax = first index; strength per utils.py:111.
    trace (tests only): receives this pass's (bins, per-bin profile maxima, per-bin top-2 gap of the profile)."""
    best = None
    bins, peaks, gaps = doppler_bins(center, spread), [], []
    for f in bins:
        prof = integrate(NON_COHERENT, data, fs, n, f, prn)
        if trace is not None:
            top2 = np.partition(prof, -2)[-2:]
            peaks.append(float(top2[1]))
            gaps.append(float(top2[1] - top2[0]))
        if best is None or np.max(prof) > np.max(best[1]):
            best = (f, prof)
    f, prof = best
    out = BestBin(f, prof, int(np.argmax(prof)), float(peak_strength(prof)))
    if trace is not None:
        trace.append(dict(bins=bins, peaks=peaks, gaps=gaps, chosen=f, strength=out.strength))
    return out


def acquire_sv(sv: int, data: np.ndarray, fs: int, n: int, trace=None) -> Acquisition:
    """acquisition.py:70-152: spread 7000 halved while >= 10 (10 passes); each pass re-centres on that pass's
    best bin; the kept result is the pass with the strictly greatest strength; then one coherent integration
    at the kept Doppler gives the carrier phase at the kept (non-coherent) peak index."""
    prn = replica(sv, n)
    center, spread, kept = 0.0, 7000.0, None
    while spread >= 10:
        b = best_bin(center, spread, data, fs, n, prn, trace)
        spread /= 2
        center = b.doppler
        if kept is None or b.strength > kept.strength:
            kept = b
    coh = integrate(COHERENT, data, fs, n, kept.doppler, prn)
    return Acquisition(sv, kept.doppler, float(np.angle(coh[kept.peak_index])), kept.peak_index, kept.strength)


def detect(svs: list[int], data: np.ndarray, fs: int, n: int) -> list[Acquisition]:
    """acquisition.py:52-68: acquire each SV, keep those with strength > 3 (config.py:7)."""
    return [r for r in (acquire_sv(sv, data, fs, n) for sv in svs) if r.strength > DETECTION_THRESHOLD]


def grid_cells(data: np.ndarray, fs: int, n: int, svs: list[int], dopplers: list[float], kind: str = NON_COHERENT):
    """The benchmark grid: one utils.py:77 evaluation per (SV, Doppler) cell,
    reduced to (max, argmax, sum, count==max).  Returns arrays shaped [len(svs), len(dopplers)]."""
    shape = (len(svs), len(dopplers))
    peak = np.zeros(shape)
    arg = np.zeros(shape, dtype=np.int64)
    total = np.zeros(shape)
    count = np.zeros(shape, dtype=np.int64)
    for a, sv in enumerate(svs):
        prn = replica(sv, n)
        for b, f in enumerate(dopplers):
            prof = integrate(kind, data, fs, n, f, prn)
            mag = np.abs(prof) if kind == COHERENT else prof
            peak[a, b] = mag.max()
            arg[a, b] = int(np.argmax(mag))
            total[a, b] = mag.sum()
            count[a, b] = int(np.count_nonzero(mag == mag.max()))
    return peak, arg, total, count


def search_is_ambiguous(trace, rel: float) -> bool:
    """Tests only: True when a float32 implementation may legitimately take a different branch of acquire_sv than the
    float64 one -- in some pass two bins' profile maxima, or the winning bin's two largest profile values, or two passes'
    strengths around the kept one, lie within `rel` (relative) of each other."""
    strengths = [p["strength"] for p in trace]
    for p in trace:
        pk = np.sort(np.asarray(p["peaks"]))
        if len(pk) > 1 and pk[-1] - pk[-2] <= rel * pk[-1]:
            return True
        w = p["bins"].index(p["chosen"])
        if p["gaps"][w] <= rel * p["peaks"][w]:
            return True
    st = np.sort(np.asarray(strengths))
    return len(st) > 1 and st[-1] - st[-2] <= rel * st[-1]


def strength_from_record(peak: float, total: float, count: int, n: int) -> float:
    """utils.py:111-116 rewritten on the reduced record: max / ((sum - count*max) / (N - count))."""
    return peak / ((total - count * peak) / (n - count))


# ------------------------------------------------------------------------------------------------------------
# Synthetic IQ.  Shared by tests and bench so CPU and GPU legs see identical bytes.
# ------------------------------------------------------------------------------------------------------------
def synth_iq(seed: int, n: int, n_ms: int, fs: int, planted, sigma: float = 1.0, nav_bits: bool = False):
    """complex64[n_ms*n]: unit-variance-ish complex gaussian noise * sigma plus, for each planted
    (sv, doppler_hz, code_phase_samples, carrier_phase_rad, amplitude), amplitude * roll(replica, code_phase)
    tiled over n_ms and rotated by exp(+j(2 pi f t + phi)); optional +-1 data bits every 20 ms."""
    rng = np.random.default_rng(seed)
    total = n * n_ms
    x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)) * (sigma / math.sqrt(2.0))
    t = np.arange(total) / fs
    for sv, f, tau_s, phi, amp in planted:
        code = np.tile(np.roll(replica(sv, n).real, tau_s), n_ms)
        if nav_bits:
            bits = rng.integers(0, 2, size=n_ms // 20 + 1) * 2 - 1
            code = code * np.repeat(bits, 20 * n)[:total]
        x = x + amp * code * np.exp(1j * (math.tau * f * t + phi))
    return x.astype(np.complex64)

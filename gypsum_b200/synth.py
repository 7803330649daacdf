"""Synthetic IQ for benchmarks and demos: seeded complex gaussian noise plus planted satellites, in
the byte format of the reference's recordings (complex64 = interleaved float32 I,Q,
antenna_sample_provider.py:112-119).  Product-side twin of the generators the test oracle carries; a test checks
they produce identical bytes."""
from __future__ import annotations

import math

import numpy as np

from gypsum_b200.gps_ca_prn_codes import ca_code_chips


def replica_real(sv: int, samples_per_ms: int) -> np.ndarray:
    """+-1 chips of SV `sv`, each repeated samples_per_ms / 1023 times (satellite.py:20-31), float64."""
    return 2.0 * np.repeat(ca_code_chips(sv), samples_per_ms // 1023) - 1.0


def synth_iq(seed: int, n: int, n_ms: int, fs: int, planted, sigma: float = 1.0) -> np.ndarray:
    """complex64[n_ms * n]: noise (g1 + j g2) * sigma / sqrt 2 plus, per planted
    (sv, doppler_hz, code_phase_samples, carrier_phase_rad, amplitude),
    amplitude * roll(replica, code_phase) tiled over n_ms * exp(+j(2 pi f t + phi))."""
    rng = np.random.default_rng(seed)
    total = n * n_ms
    x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)) * (sigma / math.sqrt(2.0))
    t = np.arange(total) / fs
    for sv, f, tau_s, phi, amp in planted:
        code = np.tile(np.roll(replica_real(sv, n), tau_s), n_ms)
        x = x + amp * code * np.exp(1j * (math.tau * f * t + phi))
    return x.astype(np.complex64)


def synth_tracking_iq(seed: int, n: int, n_ms: int, fs: int, channels, sigma: float = 0.02) -> np.ndarray:
    """Tracking-regime stream: per channel (sv, doppler_hz, doppler_rate_hz_per_s, code_phase,
    carrier_phase, amplitude) with random +-1 data bits every 20 ms at a random bit phase."""
    rng = np.random.default_rng(seed)
    total = n * n_ms
    x = (rng.standard_normal(total) + 1j * rng.standard_normal(total)) * (sigma / math.sqrt(2.0))
    t = np.arange(total) / fs
    for sv, f, rate, tau_s, phi, amp in channels:
        code = np.tile(np.roll(replica_real(sv, n), tau_s), n_ms)
        bits = rng.integers(0, 2, size=n_ms // 20 + 2) * 2 - 1
        offset = int(rng.integers(0, 20))
        data = np.repeat(bits, 20 * n)[offset * n: offset * n + total]
        x = x + amp * code * data * np.exp(1j * (math.tau * (f * t + 0.5 * rate * t * t) + phi))
    return x.astype(np.complex64)

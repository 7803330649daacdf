"""Sample ingest: file-backed provider semantics of antenna_sample_provider.py:78-136 and the
rolling acquisition window of receiver.py:68,100,219."""
import numpy as np
import pytest

from gypsum_b200.antenna_sample_provider import (AntennaSampleProviderBackedByFile, InputFileInfo, NoMoreSamplesError,
                                                 RollingSampleWindow)


def test_file_provider_matches_reference_semantics(tmp_path):
    rng = np.random.default_rng(0)
    n, fs = 2046, 2046000
    words = rng.standard_normal(2 * n * 5).astype(np.float32)
    path = tmp_path / "rec"
    words.tofile(path)
    p = AntennaSampleProviderBackedByFile(InputFileInfo(path, fs))
    assert p.get_attributes().samples_per_second == fs and p.get_attributes().samples_per_prn_transmission == n
    expect = words[0::2] + 1j * words[1::2]  # antenna_sample_provider.py:119
    for k in range(4):
        peek = p.peek_samples(n)
        chunk = p.get_samples(n)
        assert chunk.samples.dtype == np.complex64 and np.array_equal(chunk.samples, expect[k * n:(k + 1) * n])
        assert np.array_equal(peek.samples, chunk.samples)
        assert chunk.start_time == round(k * n / fs, 6) and chunk.end_time == round((k + 1) * n / fs, 6)
    assert p.seconds_since_start() == round(4 * n / fs, 6)
    with pytest.raises(NoMoreSamplesError):  # `>=` at :107: a read ending exactly at EOF is refused
        p.get_samples(n)


def test_rolling_window_equals_concatenate_of_last_ten():
    n = 64
    w = RollingSampleWindow(n, 10, pinned=False)
    rng = np.random.default_rng(1)
    chunks = []
    for k in range(37):
        c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        chunks.append(c)
        w.append(c)
        assert w.is_full() == (k >= 9) and len(w) == min(k + 1, 10)
        assert np.array_equal(w.window(), np.concatenate(chunks[-10:]))
        assert w.window().flags["C_CONTIGUOUS"]
    with pytest.raises(ValueError):
        w.append(np.zeros(n + 1, np.complex64))

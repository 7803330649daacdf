"""Boundary types of the sample stream, mirroring reference gypsum/antenna_sample_provider.py:24-35.

Any object with the same attribute names (e.g. the reference's own dataclasses) is accepted wherever these are.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SampleProviderAttributes:  # antenna_sample_provider.py:24-28
    samples_per_second: int
    samples_per_prn_transmission: int


@dataclass
class AntennaSampleChunk:  # antenna_sample_provider.py:31-35
    start_time: float
    end_time: float
    samples: np.ndarray


# ---------------------------------------------------------------------------------------------------------------
# Sample ingest: the step before the correlation path.
# ---------------------------------------------------------------------------------------------------------------
class NoMoreSamplesError(Exception):  # antenna_sample_provider.py:20-21
    pass


@dataclass
class InputFileInfo:
    """The fields of reference gypsum/radio_input.py:21-27 that the file-backed provider reads."""
    path: object
    sdr_sample_rate: float
    utc_start_time: object = None
    sample_component_data_type: type = np.float32


class AntennaSampleProviderBackedByFile:
    """Same interface and semantics as reference antenna_sample_provider.py:78-136 (peek_samples, get_samples,
    seconds_since_start, get_attributes; timestamps round(cursor / fs, 6); NoMoreSamplesError when a read would
    reach the end of the file, `>=` as at :107).

    The file's interleaved float32 I,Q words ARE complex64 samples, so chunks are zero-copy views of a memory
    map (the reference rebuilds every chunk with words[0::2] + 1j*words[1::2], :112-119) and can be handed to
    the engine as they are."""

    def __init__(self, file_info) -> None:
        import os

        self.path = file_info.path
        self.cursor = 0
        self.sample_rate = file_info.sdr_sample_rate
        start = getattr(file_info, "utc_start_time", None)
        self.utc_start_time = start.timestamp() if start is not None else 0.0
        self.sample_component_data_type = getattr(file_info, "sample_component_data_type", np.float32)
        if np.dtype(self.sample_component_data_type) != np.dtype(np.float32):
            raise ValueError("only interleaved float32 recordings are supported (GNU Radio format)")
        self.file_size_in_bytes = os.path.getsize(str(self.path))
        self._map = np.memmap(str(self.path), dtype=np.complex64, mode="r", shape=(self.file_size_in_bytes // 8,))

    def _get_elapsed_seconds_at_cursor(self, cursor: int) -> float:
        return round(cursor / self.sample_rate, 6)

    def seconds_since_start(self) -> float:
        return self._get_elapsed_seconds_at_cursor(self.cursor)

    def peek_samples(self, sample_count: int) -> AntennaSampleChunk:
        start_timestamp = self.seconds_since_start()
        file_offset_end = (self.cursor + sample_count) * 8
        if file_offset_end >= self.file_size_in_bytes:
            raise NoMoreSamplesError(
                f"Ran out of samples at {self.file_size_in_bytes/1024/1024:.2f}MB ({self.seconds_since_start():.2f}s)")
        return AntennaSampleChunk(
            start_time=start_timestamp,
            end_time=self._get_elapsed_seconds_at_cursor(self.cursor + sample_count),
            samples=self._map[self.cursor:self.cursor + sample_count],
        )

    def get_samples(self, sample_count: int) -> AntennaSampleChunk:
        chunk = self.peek_samples(sample_count)
        self.cursor += sample_count
        return chunk

    def get_attributes(self) -> SampleProviderAttributes:
        from gypsum_b200.constants import PRN_REPETITIONS_PER_SECOND

        return SampleProviderAttributes(
            samples_per_second=int(self.sample_rate),
            samples_per_prn_transmission=int(self.sample_rate // PRN_REPETITIONS_PER_SECOND),
        )


class RollingSampleWindow:
    """The receiver's rolling acquisition window (receiver.py:68 deque(maxlen=10), :100 append, :219 np.concatenate)
    without the per-scan concatenate: the last `window_ms` chunks live contiguously in one pinned host buffer
    (each chunk is written twice, `window_ms` slots apart, so every window is a contiguous slice) that the engine
    can DMA from directly (gb200_upload_iq recognises pinned memory)."""

    def __init__(self, samples_per_ms: int, window_ms: int = 10, pinned: bool = True):
        self.n = int(samples_per_ms)
        self.window_ms = int(window_ms)
        self.count = 0
        total = 2 * self.window_ms * self.n
        self._torch_keepalive = None
        if pinned:
            try:
                import torch

                t = torch.empty(total * 2, dtype=torch.float32)
                if torch.cuda.is_available():
                    t = t.pin_memory()
                self._torch_keepalive = t
                self._buf = t.numpy().view(np.complex64)
            except Exception:
                pinned = False
        if not pinned:
            self._buf = np.empty(total, dtype=np.complex64)

    def append(self, samples: np.ndarray) -> None:
        x = np.asarray(samples)
        if x.shape != (self.n,):
            raise ValueError(f"expected one millisecond ({self.n} samples)")
        slot = self.count % self.window_ms
        self._buf[slot * self.n:(slot + 1) * self.n] = x
        self._buf[(slot + self.window_ms) * self.n:(slot + self.window_ms + 1) * self.n] = x
        self.count += 1

    def __len__(self) -> int:
        return min(self.count, self.window_ms)

    def is_full(self) -> bool:
        return self.count >= self.window_ms

    def window(self) -> np.ndarray:
        """complex64[len(self) * N], oldest chunk first -- what receiver.py:219 passes to the detector."""
        k = len(self)
        if self.count <= self.window_ms:
            return self._buf[: k * self.n]
        start = self.count % self.window_ms
        return self._buf[start * self.n:(start + self.window_ms) * self.n]


class DeviceSampleRing:
    """receiver.py:68,100,219 with the window on the GPU (gb200_ring_*): `append(chunk)` uploads the new millisecond once
    and returns the chunk tagged with its place in the ring; GpsSatelliteTracker.process_samples and
    GpsSatelliteDetector.detect_satellites_in_antenna_data recognise tagged chunks / `window()` and read the samples where
    they already are instead of uploading them again (one 8*N-byte copy per millisecond in total, whatever the number of
    tracked satellites; no 10-ms re-upload per acquisition scan)."""

    def __init__(self, stream_attributes, window_ms: int = 10, device: int = 0):
        from gypsum_b200 import _native
        from gypsum_b200.utils import POOL

        self.n = int(stream_attributes.samples_per_prn_transmission)
        self.window_ms = int(window_ms)
        self._ent = POOL.get(int(stream_attributes.samples_per_second), self.n, device)
        self.native = _native.Ring(self._ent["engine"], self.window_ms)
        self.count = 0

    def append(self, chunk: AntennaSampleChunk) -> AntennaSampleChunk:
        x = np.asarray(chunk.samples)
        if x.shape != (self.n,):
            raise ValueError(f"expected one millisecond ({self.n} samples)")
        self.native.append(x)
        chunk.device_ring = self
        chunk.ring_index = self.count
        self.count += 1
        return chunk

    def holds_newest(self, chunk) -> bool:
        return getattr(chunk, "device_ring", None) is self and chunk.ring_index == self.count - 1

    def __len__(self) -> int:
        return min(self.count, self.window_ms)

    def is_full(self) -> bool:
        return self.count >= self.window_ms

    def window(self) -> "DeviceWindow":
        """The newest len(self) milliseconds, oldest first -- what receiver.py:219 concatenates for the detector."""
        return DeviceWindow(self, len(self))


class DeviceWindow:
    """A view of the newest `n_ms` milliseconds of a DeviceSampleRing; accepted by the detector in place of the ndarray."""

    def __init__(self, ring: DeviceSampleRing, n_ms: int):
        self.ring = ring
        self.n_ms = int(n_ms)
        self.size = self.n_ms * ring.n

    def bind(self) -> None:
        self.ring.native.bind_newest(self.n_ms)

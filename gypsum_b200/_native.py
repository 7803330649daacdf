"""ctypes binding of include/gypsum_b200.h.  The product path: if the shared library is missing this raises --
there is no numpy/CPU route behind it."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GB200_LIB") or os.path.join(_HERE, "libgypsum_b200.so")  # GB200_LIB: experiment builds

OK, EINVAL, ECUDA, ESTATE = 0, 1, 2, 3
COHERENT, NON_COHERENT = 1, 2

RECORD_DTYPE = np.dtype(
    [("peak", "<f4"), ("argmax", "<i4"), ("sum", "<f8"), ("count", "<i4"), ("probe_re", "<f4"), ("probe_im", "<f4"),
     ("reserved", "<i4")]
)
assert RECORD_DTYPE.itemsize == 32
TRACK_DTYPE = np.dtype(
    [("doppler", "<f8"), ("carrier_phase", "<f8"), ("error", "<f8"), ("disc", "<f8"), ("phase_acc", "<f8"),
     ("doppler_hist", "<f8"), ("carrier_phase_hist", "<f8"), ("peak_re", "<f4"), ("peak_im", "<f4"), ("strength", "<f4"), ("early_re", "<f4"), ("early_im", "<f4"),
     ("late_re", "<f4"), ("late_im", "<f4"), ("code_phase", "<i4"), ("symbol", "<i4"), ("locked", "<i4"), ("lost", "<i4"),
     ("peak_offset", "<i4"), ("reserved0", "<i4"), ("reserved1", "<i4")]
)
assert TRACK_DTYPE.itemsize == 112
ACQ_DTYPE = np.dtype([("doppler", "<f8"), ("strength", "<f8"), ("probe_re", "<f4"), ("probe_im", "<f4"),
                      ("code_phase", "<i4"), ("reserved", "<i4")])
assert ACQ_DTYPE.itemsize == 32
BEST_DTYPE = np.dtype([("doppler", "<f8"), ("strength", "<f8"), ("peak", "<f4"), ("code_phase", "<i4"), ("bin", "<i4"),
                       ("reserved", "<i4")])  # gb200_best_record
assert BEST_DTYPE.itemsize == 32

# every symbol include/gypsum_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gb200_abi_version": (C.c_int, []),
    "gb200_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gb200_destroy": (C.c_int, [_P]),
    "gb200_last_error": (C.c_char_p, [_P]),
    "gb200_set_stream": (C.c_int, [_P, _P]),
    "gb200_set_replicas": (C.c_int, [_P, _P, C.c_int]),
    "gb200_upload_iq": (C.c_int, [_P, _P, C.c_int64]),
    "gb200_bind_iq_device": (C.c_int, [_P, _P, C.c_int64]),
    "gb200_acquire_grid": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "gb200_acquire_grid_device": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "gb200_acquire_grid_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "gb200_acquire_grid_best": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "gb200_acquire_grid_best_device": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P]),
    "gb200_ring_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "gb200_ring_destroy": (C.c_int, [_P]),
    "gb200_ring_append": (C.c_int, [_P, _P, C.c_int]),
    "gb200_ring_bind_newest": (C.c_int, [_P, C.c_int]),
    "gb200_ring_appended": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "gb200_acquire_cells": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gb200_detect": (C.c_int, [_P, C.c_int, _P, C.c_int, _P]),
    "gb200_correlation_profile": (C.c_int, [_P, C.c_int, C.c_double, C.c_int, C.c_int, _P]),
    "gb200_correlation_profile_replica": (C.c_int, [_P, _P, C.c_double, C.c_int, C.c_int, _P]),
    "gb200_tracker_create": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.POINTER(_P)]),
    "gb200_tracker_destroy": (C.c_int, [_P]),
    "gb200_tracker_process": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "gb200_tracker_process_device": (C.c_int, [_P, C.c_int, _P, _P]),
    "gb200_tracker_get_state": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gb200_tracker_set_state": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "gb200_tracker_create_pool": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "gb200_tracker_reset_channel": (C.c_int, [_P, C.c_int, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "gb200_tracker_process_channels": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P]),
    "gb200_tracker_undo_channel": (C.c_int, [_P, C.c_int]),
    "gb200_grid_stream_create": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gb200_grid_stream_submit": (C.c_int, [_P, _P, _P]),
    "gb200_grid_stream_collect": (C.c_int, [_P]),
    "gb200_grid_stream_destroy": (C.c_int, [_P]),
    "gb200_tracker_integrate_bits": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int32, _P]),
    "gb200_tracker_bit_state": (C.c_int, [_P, C.c_int, _P]),
    "gb200_set_fused": (C.c_int, [_P, C.c_int]),
    "gb200_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "gb200_enable_kernel_timing": (C.c_int, [_P, C.c_int]),
    "gb200_kernel_timing": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). gypsum_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.gb200_abi_version() != 2:
            raise ImportError("libgypsum_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Engine:
    """One engine per (device, sample rate).  Thin, exception-raising wrapper over the C ABI."""

    def __init__(self, samples_per_second: int, samples_per_ms: int, device: int = 0):
        self._lib = load()
        self._h = _P()
        self._host_call_cache = None
        rc = self._lib.gb200_create(int(device), int(samples_per_second), int(samples_per_ms), C.byref(self._h))
        if rc != OK:
            msg = (self._lib.gb200_last_error(None) or b"").decode()
            raise (ValueError if rc == EINVAL else RuntimeError)(f"gb200_create: {msg}")
        self.samples_per_second = int(samples_per_second)
        self.samples_per_ms = int(samples_per_ms)
        self.device = int(device)
        self.n_prn = 0
        # What the engine's single IQ binding currently holds: a caller-chosen tag (e.g. the chunk key the trackers use to
        # share one upload), cleared by EVERY call that rebinds the IQ -- so no caller can mistake another caller's samples
        # for its own.
        self.iq_tag = None
        self._children = weakref.WeakSet()  # trackers / grid streams: they hold device memory tied to this engine

    # -- plumbing ------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc == OK:
            return
        msg = (self._lib.gb200_last_error(self._h) or b"").decode()
        raise (ValueError if rc == EINVAL else RuntimeError)(f"{what}: {msg}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            for child in list(getattr(self, "_children", ())):
                child.close()  # before the engine goes: their buffers live behind its handle's device
            self._lib.gb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int) -> None:
        self._check(self._lib.gb200_set_stream(self._h, _P(cuda_stream_ptr or 0)), "gb200_set_stream")

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.gb200_launch_count(self._h, C.byref(n)), "gb200_launch_count")
        return n.value

    def set_fused(self, mode) -> None:
        """acquire_cells kernel choice: True / 1 = fused block-per-(PRN, Doppler) kernel, False / 0 = doppler_spectra +
        correlate_cells, None / -1 = automatic (default)."""
        m = -1 if mode is None else int(mode)
        self._check(self._lib.gb200_set_fused(self._h, m), "gb200_set_fused")

    def enable_kernel_timing(self, on: bool) -> None:
        self._check(self._lib.gb200_enable_kernel_timing(self._h, int(bool(on))), "gb200_enable_kernel_timing")

    def kernel_timing(self, which: int) -> tuple[float, int]:
        """(total device ms, launches) of doppler_spectra (0) / correlate_cells (1) since timing was enabled."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.gb200_kernel_timing(self._h, which, C.byref(ms), C.byref(n)), "gb200_kernel_timing")
        return ms.value, n.value

    # -- inputs --------------------------------------------------------------------------------------------
    def set_replicas(self, chips: np.ndarray) -> None:
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        if chips.ndim != 2 or chips.shape[1] != 1023:
            raise ValueError("chips must be [n_prn, 1023]")
        self._check(self._lib.gb200_set_replicas(self._h, _ptr(chips), chips.shape[0]), "gb200_set_replicas")
        self.n_prn = chips.shape[0]

    def upload_iq(self, samples: np.ndarray, tag=None) -> None:
        """tag: optional hashable naming these samples; `iq_tag` holds it until the next call that rebinds the IQ."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        self._iq_keepalive = x
        self.iq_tag = None
        self._check(self._lib.gb200_upload_iq(self._h, _ptr(x), x.size), "gb200_upload_iq")
        self.iq_tag = tag

    def upload_iq_ptr(self, host_ptr: int, n_samples: int) -> None:
        self.iq_tag = None
        self._check(self._lib.gb200_upload_iq(self._h, _P(host_ptr), n_samples), "gb200_upload_iq")

    def bind_iq_device(self, device_ptr: int, n_samples: int, tag=None) -> None:
        self.iq_tag = None
        self._check(self._lib.gb200_bind_iq_device(self._h, _P(device_ptr), n_samples), "gb200_bind_iq_device")
        self.iq_tag = tag

    # -- the hot path --------------------------------------------------------------------------------------
    def acquire_grid(self, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int = NON_COHERENT,
                     out: np.ndarray | None = None) -> np.ndarray:
        """out: optional preallocated RECORD_DTYPE array [n_blocks, n_prn, n_doppler]; if it lives in pinned memory
        (e.g. a view of a torch pinned tensor) the records are DMA'd straight into it."""
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        if out is None:
            out = np.empty((n_blocks, prn.size, dop.size), dtype=RECORD_DTYPE)
        elif out.dtype != RECORD_DTYPE or out.shape != (n_blocks, prn.size, dop.size) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous RECORD_DTYPE array of shape [n_blocks, n_prn, n_doppler]")
        self._check(
            self._lib.gb200_acquire_grid(self._h, n_blocks, ms_per_block, _ptr(prn), prn.size, _ptr(dop), dop.size, kind,
                                         _ptr(out)),
            "gb200_acquire_grid",
        )
        return out

    def acquire_grid_host(self, iq, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int = NON_COHERENT,
                          out: np.ndarray | None = None) -> np.ndarray:
        """upload + grid + records back in ONE call, replayed as a CUDA graph per shape (latency-bound callers).
        iq: complex64 array of n_blocks * ms_per_block * N samples, or an int host address of such a buffer."""
        # latency path: numpy's .ctypes costs ~1 us per array, so the addresses of the axes and of the record buffer are kept
        # for as long as the caller passes the very same array objects (their data cannot move while we hold them)
        c = self._host_call_cache
        if c is not None and c[0] is prn_idx and c[1] is doppler_hz and c[2] is out:
            prn, dop, p_prn, p_dop, p_out = c[3:]
        else:
            prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
            dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
            if out is None:
                out = np.empty((n_blocks, prn.size, dop.size), dtype=RECORD_DTYPE)
                cacheable = False
            else:
                cacheable = prn is prn_idx and dop is doppler_hz  # no converted copies that could go stale
            p_prn, p_dop, p_out = _ptr(prn), _ptr(dop), _ptr(out)
            self._host_call_cache = (prn_idx, doppler_hz, out, prn, dop, p_prn, p_dop, p_out) if cacheable else None
        if out.dtype != RECORD_DTYPE or out.shape != (n_blocks, prn.size, dop.size) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous RECORD_DTYPE array of shape [n_blocks, n_prn, n_doppler]")
        if isinstance(iq, int):
            keep, ptr = None, iq
        elif isinstance(iq, np.integer):
            keep, ptr = None, int(iq)
        else:
            keep = np.ascontiguousarray(iq, dtype=np.complex64)
            if keep.size < n_blocks * ms_per_block * self.samples_per_ms:
                raise ValueError("not enough samples for the grid")
            ptr = _ptr(keep)
        self.iq_tag = None
        rc = self._lib.gb200_acquire_grid_host(self._h, ptr, n_blocks, ms_per_block, p_prn, prn.size, p_dop, dop.size, kind, p_out)
        if rc:
            self._check(rc, "gb200_acquire_grid_host")
        return out

    def acquire_grid_best(self, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int = NON_COHERENT) -> np.ndarray:
        """acquisition.py:179-189 per (block, prn) row: BEST_DTYPE [n_blocks, n_prn]."""
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        out = np.empty((n_blocks, prn.size), dtype=BEST_DTYPE)
        self._check(
            self._lib.gb200_acquire_grid_best(self._h, n_blocks, ms_per_block, _ptr(prn), prn.size, _ptr(dop), dop.size, kind,
                                              _ptr(out)),
            "gb200_acquire_grid_best",
        )
        return out

    def acquire_grid_best_device(self, n_blocks, ms_per_block, prn: np.ndarray, dop: np.ndarray, kind: int, out_device_ptr: int):
        self._check(
            self._lib.gb200_acquire_grid_best_device(self._h, n_blocks, ms_per_block, _ptr(prn), prn.size, _ptr(dop), dop.size,
                                                     kind, _P(out_device_ptr)),
            "gb200_acquire_grid_best_device",
        )

    def acquire_grid_device(self, n_blocks, ms_per_block, prn: np.ndarray, dop: np.ndarray, kind: int, out_device_ptr: int):
        """prn (int32) / dop (float64) must be contiguous arrays kept alive by the caller; enqueue only."""
        self._check(
            self._lib.gb200_acquire_grid_device(self._h, n_blocks, ms_per_block, _ptr(prn), prn.size, _ptr(dop), dop.size,
                                                kind, _P(out_device_ptr)),
            "gb200_acquire_grid_device",
        )

    def acquire_cells(self, prn_idx, doppler_hz, n_ms: int, kind: int = NON_COHERENT, probe_idx=None) -> np.ndarray:
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        if prn.shape != dop.shape or prn.ndim != 1:
            raise ValueError("prn_idx and doppler_hz must be 1-D and the same length")
        probe = None if probe_idx is None else np.ascontiguousarray(probe_idx, dtype=np.int32)
        out = np.empty(prn.size, dtype=RECORD_DTYPE)
        self._check(
            self._lib.gb200_acquire_cells(self._h, prn.size, _ptr(prn), _ptr(dop), None if probe is None else _ptr(probe),
                                          n_ms, kind, _ptr(out)),
            "gb200_acquire_cells",
        )
        return out

    def detect(self, prn_idx, n_ms: int) -> np.ndarray:
        """acquisition.py:70-152 for every listed replica row, all ten passes + the coherent pass on the device."""
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        out = np.empty(prn.size, dtype=ACQ_DTYPE)
        self._check(self._lib.gb200_detect(self._h, prn.size, _ptr(prn), int(n_ms), _ptr(out)), "gb200_detect")
        return out

    def correlation_profile(self, prn_idx: int, doppler_hz: float, n_ms: int, kind: int) -> np.ndarray:
        n = self.samples_per_ms
        out = np.empty(n * (2 if kind == COHERENT else 1), dtype=np.float32)
        self._check(
            self._lib.gb200_correlation_profile(self._h, int(prn_idx), float(doppler_hz), int(n_ms), int(kind), _ptr(out)),
            "gb200_correlation_profile",
        )
        return out.view(np.complex64) if kind == COHERENT else out

    def correlation_profile_replica(self, replica: np.ndarray, doppler_hz: float, n_ms: int, kind: int) -> np.ndarray:
        """utils.py:77-108 against an ARBITRARY complex replica of samples_per_ms samples (direct evaluation)."""
        n = self.samples_per_ms
        rep = np.ascontiguousarray(replica, dtype=np.complex64)
        if rep.shape != (n,):
            raise ValueError(f"replica must have {n} samples")
        out = np.empty(n * (2 if kind == COHERENT else 1), dtype=np.float32)
        self._check(
            self._lib.gb200_correlation_profile_replica(self._h, _ptr(rep), float(doppler_hz), int(n_ms), int(kind), _ptr(out)),
            "gb200_correlation_profile_replica",
        )
        return out.view(np.complex64) if kind == COHERENT else out


class GridStream:
    """Pipelined stream of equally shaped grid batches (gb200_grid_stream_*): `submit` enqueues copy-in, the grid and
    copy-out of one batch, `collect` waits for the oldest batch in flight and returns its record array.  With depth >= 2
    the transfers of neighbouring batches run under the kernels of the current one."""

    def __init__(self, engine: Engine, n_blocks: int, ms_per_block: int, prn_idx, doppler_hz, kind: int = NON_COHERENT,
                 depth: int = 2):
        self._engine = engine
        self._lib = engine._lib
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        self.shape = (n_blocks, prn.size, dop.size)
        self.samples_per_batch = n_blocks * ms_per_block * engine.samples_per_ms
        self.depth = depth
        self._pending: list = []  # (iq keep-alive, out array) per batch in flight
        self._h = _P()
        engine._check(self._lib.gb200_grid_stream_create(engine._h, n_blocks, ms_per_block, _ptr(prn), prn.size, _ptr(dop),
                                                         dop.size, kind, depth, C.byref(self._h)), "gb200_grid_stream_create")
        engine._children.add(self)

    @property
    def in_flight(self) -> int:
        return len(self._pending)

    def submit(self, iq, out: np.ndarray | None = None) -> None:
        """iq: complex64 array of samples_per_batch samples, or an int host address of such a buffer.  out: optional
        RECORD_DTYPE array [n_blocks, n_prn, n_doppler] (pinned memory = direct DMA)."""
        if isinstance(iq, (int, np.integer)):
            keep, ptr = None, _P(int(iq))
        else:
            keep = np.ascontiguousarray(iq, dtype=np.complex64)
            if keep.size != self.samples_per_batch:
                raise ValueError(f"a batch is {self.samples_per_batch} samples, got {keep.size}")
            ptr = _ptr(keep)
        if out is None:
            out = np.empty(self.shape, dtype=RECORD_DTYPE)
        elif out.dtype != RECORD_DTYPE or out.shape != self.shape or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous RECORD_DTYPE array of shape [n_blocks, n_prn, n_doppler]")
        self._engine.iq_tag = None  # the stream rebinds the engine's IQ to its own slot buffer
        self._engine._check(self._lib.gb200_grid_stream_submit(self._h, ptr, _ptr(out)), "gb200_grid_stream_submit")
        self._pending.append((keep, out))

    def collect(self) -> np.ndarray:
        self._engine._check(self._lib.gb200_grid_stream_collect(self._h), "gb200_grid_stream_collect")
        return self._pending.pop(0)[1]

    def close(self) -> None:
        if getattr(self, "_h", None) and getattr(self._engine, "_h", None):
            self._lib.gb200_grid_stream_destroy(self._h)
        self._h = None
        self._pending = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ring:
    """Device-resident rolling window of the newest milliseconds (gb200_ring_*; receiver.py:68,100,219)."""

    def __init__(self, engine: Engine, capacity_ms: int):
        self._engine = engine
        self._lib = engine._lib
        self.capacity_ms = int(capacity_ms)
        self._h = _P()
        engine._check(self._lib.gb200_ring_create(engine._h, self.capacity_ms, C.byref(self._h)), "gb200_ring_create")
        engine._children.add(self)

    def append(self, samples) -> None:
        """samples: complex64[n_ms * N] (whole milliseconds)."""
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        n = self._engine.samples_per_ms
        if x.size == 0 or x.size % n:
            raise ValueError("append whole milliseconds")
        self._engine._check(self._lib.gb200_ring_append(self._h, _ptr(x), x.size // n), "gb200_ring_append")
        if isinstance(self._engine.iq_tag, tuple) and self._engine.iq_tag[:1] == ("ring",):
            self._engine.iq_tag = None  # a binding into this ring no longer names the newest samples

    @property
    def appended_ms(self) -> int:
        n = C.c_int64()
        self._engine._check(self._lib.gb200_ring_appended(self._h, C.byref(n)), "gb200_ring_appended")
        return n.value

    def bind_newest(self, n_ms: int) -> None:
        """The engine's IQ := the newest n_ms milliseconds, in place."""
        tag = ("ring", id(self), self.appended_ms, int(n_ms))
        if self._engine.iq_tag == tag:
            return
        self._engine.iq_tag = None
        self._engine._check(self._lib.gb200_ring_bind_newest(self._h, int(n_ms)), "gb200_ring_bind_newest")
        self._engine.iq_tag = tag

    def close(self) -> None:
        if getattr(self, "_h", None) and getattr(self._engine, "_h", None):
            self._lib.gb200_ring_destroy(self._h)
            self._engine.iq_tag = None
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tracker:
    """A bank of tracking channels on one engine (gb200_tracker_*).  Channels consume the engine's loaded IQ."""

    @classmethod
    def pool(cls, engine: Engine, capacity: int) -> "Tracker":
        """`capacity` idle channel slots (gb200_tracker_create_pool); seed them with reset_channel."""
        self = cls.__new__(cls)
        self._engine = engine
        self._lib = engine._lib
        self.n_channels = int(capacity)
        self._h = _P()
        engine._check(self._lib.gb200_tracker_create_pool(engine._h, int(capacity), C.byref(self._h)), "gb200_tracker_create_pool")
        engine._children.add(self)
        return self

    def reset_channel(self, channel: int, prn_idx: int, doppler: float, carrier_phase: float, code_phase: int) -> None:
        self._engine._check(self._lib.gb200_tracker_reset_channel(self._h, int(channel), int(prn_idx), float(doppler),
                                                                  float(carrier_phase), int(code_phase)),
                            "gb200_tracker_reset_channel")

    def process_channels(self, channels, n_ms: int, start_times, want_profiles: bool = False, keep_undo: bool = False):
        """gb200_tracker_process for a subset, one launch: records [len(channels), n_ms] (and profiles)."""
        sel = np.ascontiguousarray(channels, dtype=np.int32)
        ts = np.ascontiguousarray(start_times, dtype=np.float64)
        if ts.shape != (n_ms,):
            raise ValueError("start_times must hold one timestamp per millisecond")
        out = np.empty((sel.size, n_ms), dtype=TRACK_DTYPE)
        prof = np.empty((sel.size, n_ms, self._engine.samples_per_ms), dtype=np.float32) if want_profiles else None
        self._engine._check(
            self._lib.gb200_tracker_process_channels(self._h, sel.size, _ptr(sel), n_ms, _ptr(ts), int(bool(keep_undo)),
                                                     _ptr(out), None if prof is None else _ptr(prof)),
            "gb200_tracker_process_channels")
        return (out, prof) if want_profiles else out

    def undo_channel(self, channel: int) -> None:
        self._engine._check(self._lib.gb200_tracker_undo_channel(self._h, int(channel)), "gb200_tracker_undo_channel")

    def __init__(self, engine: Engine, prn_idx, doppler_hz, carrier_phase, code_phase):
        self._engine = engine
        self._lib = engine._lib
        prn = np.ascontiguousarray(prn_idx, dtype=np.int32)
        dop = np.ascontiguousarray(doppler_hz, dtype=np.float64)
        cph = np.ascontiguousarray(carrier_phase, dtype=np.float64)
        code = np.ascontiguousarray(code_phase, dtype=np.int32)
        if not (prn.shape == dop.shape == cph.shape == code.shape) or prn.ndim != 1:
            raise ValueError("per-channel arrays must be 1-D and the same length")
        self.n_channels = prn.size
        self._h = _P()
        engine._check(self._lib.gb200_tracker_create(engine._h, prn.size, _ptr(prn), _ptr(dop), _ptr(cph), _ptr(code),
                                                     C.byref(self._h)), "gb200_tracker_create")
        engine._children.add(self)

    def close(self) -> None:
        if getattr(self, "_h", None) and getattr(self._engine, "_h", None):
            self._lib.gb200_tracker_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, n_ms: int, start_times, want_profiles: bool = False):
        """Records [n_channels, n_ms] (TRACK_DTYPE) and, optionally, |prompt profile| [n_channels, n_ms, N]."""
        ts = np.ascontiguousarray(start_times, dtype=np.float64)
        if ts.shape != (n_ms,):
            raise ValueError("start_times must hold one timestamp per millisecond")
        out = np.empty((self.n_channels, n_ms), dtype=TRACK_DTYPE)
        prof = np.empty((self.n_channels, n_ms, self._engine.samples_per_ms), dtype=np.float32) if want_profiles else None
        self._engine._check(
            self._lib.gb200_tracker_process(self._h, n_ms, _ptr(ts), _ptr(out), None if prof is None else _ptr(prof)),
            "gb200_tracker_process")
        return (out, prof) if want_profiles else out

    def process_device(self, n_ms: int, start_times: np.ndarray, out_device_ptr: int) -> None:
        self._engine._check(self._lib.gb200_tracker_process_device(self._h, n_ms, _ptr(start_times), _P(out_device_ptr)),
                            "gb200_tracker_process_device")

    def integrate_bits(self, n_ms: int, start_times, end_times, records_device_ptr: int | None = None) -> list:
        """navigation_bit_intergrator.py:278-288 for every channel over records in device memory (default: the ones the
        last `process` call left there).  Returns one BIT_DTYPE array per channel."""
        t0 = np.ascontiguousarray(start_times, dtype=np.float64)
        t1 = np.ascontiguousarray(end_times, dtype=np.float64)
        if t0.shape != (n_ms,) or t1.shape != (n_ms,):
            raise ValueError("start_times / end_times must hold one timestamp per millisecond")
        cap = n_ms // 20 + 8  # whole bits in the call + the backlog a first phase decision releases (<= 81 symbols)
        ev = np.empty((self.n_channels, cap), dtype=BIT_DTYPE)
        cnt = np.empty(self.n_channels, dtype=np.int32)
        self._engine._check(
            self._lib.gb200_tracker_integrate_bits(self._h, n_ms, _ptr(t0), _ptr(t1),
                                                   None if records_device_ptr is None else _P(records_device_ptr),
                                                   _ptr(ev), cap, _ptr(cnt)), "gb200_tracker_integrate_bits")
        if (cnt > cap).any():
            raise RuntimeError("bit event buffer too small")  # cannot happen: see `cap`
        return [ev[c, : cnt[c]].copy() for c in range(self.n_channels)]

    def bit_state(self, channel: int) -> dict:
        out = np.zeros(8, dtype=np.int64)
        self._engine._check(self._lib.gb200_tracker_bit_state(self._h, channel, _ptr(out)), "gb200_tracker_bit_state")
        keys = ("emitted_bit_count", "failed_bit_count", "processed_pseudosymbol_count", "slide", "determined_bit_phase",
                "previous_bit_phase_decision", "pseudosymbol_cursor_within_queue", "stopped")
        d = dict(zip(keys, (int(v) for v in out)))
        for k in ("determined_bit_phase", "previous_bit_phase_decision"):
            d[k] = None if d[k] < 0 else d[k]
        return d

    def get_state(self, channel: int) -> dict:
        d, c, a = C.c_double(), C.c_double(), C.c_double()
        p, lost = C.c_int32(), C.c_int32()
        self._engine._check(self._lib.gb200_tracker_get_state(self._h, channel, C.byref(d), C.byref(c), C.byref(a), C.byref(p),
                                                              C.byref(lost)), "gb200_tracker_get_state")
        return {"doppler": d.value, "carrier_phase": c.value, "phase_acc": a.value, "code_phase": p.value, "lost": lost.value}

    def set_state(self, channel: int, doppler: float, carrier_phase: float, phase_acc: float, code_phase: int) -> None:
        self._engine._check(self._lib.gb200_tracker_set_state(self._h, channel, float(doppler), float(carrier_phase),
                                                              float(phase_acc), int(code_phase)), "gb200_tracker_set_state")


BIT_DTYPE = np.dtype([  # gb200_bit_event
    ("receiver_timestamp", "<f8"), ("trailing_edge_receiver_timestamp", "<f8"), ("ms_index", "<i4"), ("bit_value", "<i4"),
    ("slide", "<i4"), ("pad_", "<i4")])
assert BIT_DTYPE.itemsize == 32


def strength_from_records(rec: np.ndarray, n: int) -> np.ndarray:
    """utils.py:111-116 on the reduced record: peak / mean(profile[profile != peak])."""
    peak = rec["peak"].astype(np.float64)
    cnt = rec["count"].astype(np.float64)
    return peak / ((rec["sum"] - cnt * peak) / (n - cnt))

"""CPU-side checks of the boundary: the shared library builds, loads and exports every symbol the header
declares; host-side argument handling; replica recognition.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from oracle import gypsum_oracle as o
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    from gypsum_b200 import _native

    header = open(os.path.join(ROOT, "include", "gypsum_b200.h")).read()
    declared = set(re.findall(r"\b(gb200_[a-z_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    for name in declared:
        assert getattr(native_lib, name) is not None
    assert native_lib.gb200_abi_version() == 2
    assert _native.RECORD_DTYPE.itemsize == 32
    assert [_native.RECORD_DTYPE.fields[k][1] for k in ("peak", "argmax", "sum", "count", "probe_re", "probe_im")] == [
        0, 4, 8, 16, 20, 24]


def test_no_gpu_is_an_error_not_a_fallback(native_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gypsum_b200 import _native

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Engine(2046000, 2046)
    with pytest.raises(ValueError):
        _native.Engine(2046000, 2047)  # not a multiple of 1023


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gypsum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f"{f} imports the oracle"
                assert "/root/reference" not in text


def test_replica_recognition():
    from gypsum_b200.utils import chips_of_replica

    for n in (2046, 4092):
        prn = o.replica(5, n)
        chips, roll = chips_of_replica(prn, n)
        assert roll == 0 and np.array_equal(chips, o.ca_code(5).astype(np.uint8))
        s = n // 1023
        for shift in (1, s, 777, n - 1):
            c2, r2 = chips_of_replica(np.roll(prn, shift), n)
            assert np.array_equal(np.roll(np.repeat(2.0 * c2 - 1, s), r2), np.roll(prn, shift).real)
    with pytest.raises(ValueError):
        chips_of_replica(np.ones(2046) * 0.5, 2046)
    with pytest.raises(ValueError):
        chips_of_replica(o.replica(5, 2046)[:-1], 2046)


def test_integration_type_enum_matches_reference_values():
    from gypsum_b200.utils import IntegrationType, _kind

    assert IntegrationType.Coherent.value == 1 and IntegrationType.NonCoherent.value == 2
    assert _kind(IntegrationType.Coherent) == 1 and _kind(IntegrationType.NonCoherent) == 2
    with pytest.raises(ValueError, match="Unexpected integration type"):
        _kind("nope")


def test_product_synth_matches_oracle_generators():
    """gypsum_b200.synth (used by bench.py / tools) and the oracle's generators produce identical bytes."""
    from gypsum_b200 import synth
    from oracle import tracker_oracle as t

    planted = [(25, 1500.0, 777, 0.3, 0.3), (3, -3250.5, 5, 1.0, 0.2)]
    for n, m in ((2046, 3), (4092, 1)):
        assert np.array_equal(synth.synth_iq(9, n, m, n * 1000, planted), o.synth_iq(9, n, m, n * 1000, planted))
    ch = [(7, -2212.7, 0.5, 100, 1.0, 0.005)]
    assert np.array_equal(synth.synth_tracking_iq(4, 2046, 45, 2046000, ch), t.synth_tracking_iq(4, 2046, 45, 2046000, ch))


def test_header_is_plain_c_and_links_from_c(native_lib, tmp_path):
    """include/gypsum_b200.h must be consumable by a C compiler (the boundary is a C ABI, not C++), and a C program
    linked against the library must see the record layouts the header promises and get a clean error -- not a crash,
    not a fallback -- when no CUDA device is usable (GPU boxes: the engine comes up instead)."""
    import subprocess

    from gypsum_b200 import _native

    src = tmp_path / "probe.c"
    src.write_text(r'''
#include <stdio.h>
#include "gypsum_b200.h"
int main(void) {
    gb200_engine* e = NULL;
    int rc = gb200_create(0, 2046000, 2046, &e);
    printf("%d %d %d %d %d %d\n", gb200_abi_version(), (int)sizeof(gb200_cell_record), (int)sizeof(gb200_track_record),
           (int)sizeof(gb200_acquisition_result), (int)sizeof(gb200_bit_event), rc);
    if (rc != GB200_OK) { printf("%s\n", gb200_last_error(NULL)); return 0; }
    gb200_destroy(e);
    return 0;
}
''')
    exe = tmp_path / "probe"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe), _native.LIB_PATH,
                    f"-Wl,-rpath,{os.path.dirname(_native.LIB_PATH)}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    fields = [int(v) for v in out[0].split()]
    assert fields[0] == 2 and fields[1:5] == [32, 112, 32, 32]
    assert fields[5] == 0 or "no usable CUDA device" in out[1]


def test_acquire_grid_host_wrapper_address_cache():
    """Engine.acquire_grid_host keeps the addresses of the axis / record arrays only while the caller passes the very same
    array objects and no converted copy was made; everything else is looked up again, and the shape check runs every time."""
    from gypsum_b200 import _native

    calls = []

    class Lib:
        def gb200_acquire_grid_host(self, h, iq, nb, m, p_prn, n_prn, p_dop, n_dop, kind, p_out):
            calls.append((iq, nb, m, p_prn, n_prn, p_dop, n_dop, kind, p_out))
            return 0

    eng = object.__new__(_native.Engine)
    eng._lib, eng._h, eng._host_call_cache, eng.samples_per_ms, eng.iq_tag = Lib(), None, None, 2046, "chunk"
    prn = np.arange(4, dtype=np.int32)
    dop = np.linspace(-1000.0, 1000.0, 5)
    out = np.empty((1, 4, 5), dtype=_native.RECORD_DTYPE)
    iq = np.zeros(2046, dtype=np.complex64)

    assert eng.acquire_grid_host(iq, 1, 1, prn, dop, out=out) is out
    assert eng.iq_tag is None  # the engine's IQ binding changed hands
    assert calls[-1][3:7] == (prn.ctypes.data, 4, dop.ctypes.data, 5) and calls[-1][8] == out.ctypes.data
    cached = eng._host_call_cache
    assert cached is not None
    eng.acquire_grid_host(iq.ctypes.data, 1, 1, prn, dop, out=out)  # same objects, IQ by address: served from the cache
    assert eng._host_call_cache is cached and calls[-1][0] == iq.ctypes.data and calls[-1][8] == out.ctypes.data

    out2 = np.empty((1, 4, 5), dtype=_native.RECORD_DTYPE)
    eng.acquire_grid_host(iq, 1, 1, prn, dop, out=out2)  # another record buffer: looked up again
    assert calls[-1][8] == out2.ctypes.data and eng._host_call_cache is not cached

    eng.acquire_grid_host(iq, 1, 1, [0, 1, 2, 3], dop, out=out2)  # a list is converted to a temporary: nothing may be kept
    assert eng._host_call_cache is None and calls[-1][4] == 4
    made = eng.acquire_grid_host(iq, 1, 1, prn, dop)  # no buffer given: a fresh one each call, never cached
    assert made.shape == (1, 4, 5) and eng._host_call_cache is None

    eng.acquire_grid_host(iq, 1, 1, prn, dop, out=out)
    with pytest.raises(ValueError):
        eng.acquire_grid_host(iq, 2, 1, prn, dop, out=out)  # cached objects, wrong shape for this call
    with pytest.raises(ValueError):
        eng.acquire_grid_host(iq[:100], 1, 1, prn, dop, out=out)  # not enough samples

"""The reference's own MAPQ-cap unit tests (unittest/minimizer_mapper.cpp:112-252), run against the oracle's
faster_cap (minimizer_mapper.cpp:2946-3260): the cap is never infinite, whatever the agglomerations.
(The GPU faster_cap is compared with this oracle through every MAPQ of the map parity tests.)"""
import ctypes as C
import math

import numpy as np

import helpers as H
from vg_b200 import capi


def _cap(offset, length, agg_start, agg_len, read_len, quality):
    lib = H.oracle_lib()
    lib.oracle_faster_cap_all_g.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint8]
    lib.oracle_faster_cap_all_g.restype = C.c_double
    arrs = [np.ascontiguousarray(a, dtype=np.uint32) for a in (offset, length, agg_start, agg_len)]
    return lib.oracle_faster_cap_all_g(len(arrs[0]), *(capi.ptr(a) for a in arrs), read_len, quality)


def test_cap_not_confused_by_excessive_gs():
    """:154-176 with cover_in_minimizers(sequence, 25, 10, stride 1) (:112-151): 150 Gs at quality 0x1E."""
    L, core, flank = 150, 25, 10
    offset, length, start, alen = [], [], [], []
    for core_start in range(0, L - core, 1):
        if core_start <= flank:
            s, n = 0, core + flank + core_start
        elif L - core_start - core <= flank:
            s = core_start - flank; n = L - s - 1
        else:
            s, n = core_start - flank, core + 2 * flank
        offset.append(core_start); length.append(core); start.append(s); alen.append(n)
    cap = _cap(offset, length, start, alen, L, 0x1E)
    assert math.isfinite(cap) and cap > 0


def test_cap_not_confused_by_fuzzing_with_high_base_qualities():
    """:178-252: random unrealistic agglomerations on 100 Gs at quality 60 (numpy RNG instead of rand(); 3000 tries)."""
    rng = np.random.default_rng(1234)
    L = 100
    for _ in range(3000):
        count = int(rng.integers(0, 100)) + 5
        offset, length, start, alen = [], [], [], []
        for _ in range(count):
            core = int(rng.integers(0, min(L // 2 - 1, 31))) + 1
            run = int(rng.integers(0, min(L - core, 32 - core)))
            flank = int(rng.integers(0, 10))
            core_start = int(rng.integers(0, L - core - run))
            s, n = core_start, core + run + 2 * flank
            if flank > s:
                n -= flank - s; s = 0
            else:
                s -= flank
            if s + n > L:
                n = L - s
            offset.append(core_start); length.append(core); start.append(s); alen.append(n)
        cap = _cap(offset, length, start, alen, L, 60)
        assert not math.isinf(cap) and not math.isnan(cap)

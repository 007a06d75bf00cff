"""The metadata wire: values must equal what strconv.ParseFloat reads back from FFmpeg's printf
("%.3f" lavfi.r128.*, "%f" astats, "%g" aspectralstats, "%.2f" loudnorm JSON).  The library's
arithmetic fast path is checked against the literal snprintf/strtod round trip."""
import ctypes as C
import math
import numpy as np
from jivetalking_b200 import gpudsp


def test_fast_wire_equals_printf_round_trip():
    L = C.CDLL(gpudsp.LIB_PATH)
    L.jt_debug_wire.restype = C.c_double
    L.jt_debug_wire.argtypes = [C.c_char_p, C.c_double, C.c_int]
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(20000) * 30, rng.standard_normal(20000) * 1e-4, 10 ** rng.uniform(-14, 14, 20000),
        -(10 ** rng.uniform(-14, 14, 5000)), np.round(rng.standard_normal(5000) * 10, 3) + 0.0005,
        np.array([0.0, -0.0, 1.0, 999999.5, 99999.95, 0.0999999, 1e-5, 123456.5, 0.5, 1.0005, 2.0015, -23.0065,
                  float("inf"), -float("inf"), 1e300, 1e-300, 1.7976931348623157e308])])
    for fmt in (b"%.3f", b"%f", b"%g", b"%.2f"):
        for v in vals:
            a, b = L.jt_debug_wire(fmt, float(v), 0), L.jt_debug_wire(fmt, float(v), 1)
            assert a == b or (math.isnan(a) and math.isnan(b)), (fmt, v, a, b)
            if math.isfinite(b) and fmt != b"%g":
                assert b == float(fmt.decode() % v)
    assert math.isnan(L.jt_debug_wire(b"%.3f", float("nan"), 0))

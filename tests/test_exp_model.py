"""The constants of the device exponential `so_exps(p, q) = exp(p q ln2/256)` (scalaopt_b200/csrc/so_device.cuh), pinned on CPU:
the header is parsed, the algorithm is replayed in numpy (double arithmetic, the two FMAs of the argument reduction in
extended precision) and compared with 50-digit arithmetic.  Not a product path: the product runs the CUDA code; this test
only guards the numbers in the header (a mistyped coefficient would still pass the 1e-12 parity tests on smooth sums)."""
import os
import re

import numpy as np
import pytest

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scalaopt_b200", "csrc", "so_device.cuh")


def _consts():
    src = open(HDR).read()
    blk = src[src.index("static __constant__ ExpSConsts kExpS"):]
    blk = blk[:blk.index("};")]
    vals = [float.fromhex(v) if "x" in v else float(v) for v in re.findall(r"^\s*(0x[0-9a-fA-F.]+p[+-]?\d+|\d+\.\d+),", blk, re.M)]
    shift, c1, c2, c4, unscale = vals
    c3_bits = int(re.search(r"SO_EXPS_C3 __longlong_as_double\(0x([0-9A-Fa-f]+)LL\)", src).group(1), 16)
    c3 = np.array([c3_bits], dtype=np.uint64).view(np.float64)[0]
    scale = float.fromhex(re.search(r"kExpScale = (0x[0-9a-fA-F.]+p[+-]?\d+)", src).group(1))
    scale_sqrt = float.fromhex(re.search(r"kExpScaleSqrt = (0x[0-9a-fA-F.]+p[+-]?\d+)", src).group(1))
    return shift, c1, c2, float(c3), c4, unscale, scale, scale_sqrt


def test_constants_are_what_the_comments_say():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    shift, c1, c2, c3, c4, unscale, scale, scale_sqrt = _consts()
    c = mp.log(2) / 256
    assert shift == 1.5 * 2.0 ** 52
    assert c1 == float(c) and unscale == float(c)
    assert c2 == float(c ** 2 / 2)
    assert c4 == float(c ** 4 / 24)
    assert abs(c3 - float(c ** 3 / 6)) / float(c ** 3 / 6) < 2.0 ** -21          # rounded to a 32-bit immediate
    assert np.array([c3]).view(np.uint64)[0] & 0xFFFFFFFF == 0
    assert scale == float(1 / c) and abs(scale_sqrt - float(mp.sqrt(1 / c))) <= 2.0 ** -48


def test_replayed_algorithm_is_within_two_ulp_of_exp():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    shift, c1, c2, c3, c4, _, scale, _ = _consts()
    rng = np.random.default_rng(12345)
    n = 4000
    # arguments over the whole valid range |x| < 700, as products (rate * t) with the 256/ln2 folded into one factor
    x = np.concatenate([rng.uniform(-700, 700, n // 2), rng.uniform(-5, 5, n // 4), rng.normal(0, 1e-3, n // 4)])
    t = rng.uniform(0.1, 10.0, x.size)
    p, q = t, (x / t) * scale
    ld = np.longdouble
    kd = (ld(p) * ld(q) + ld(shift)).astype(np.float64)              # fma(p, q, shift), one rounding
    k = (kd - shift)                                                 # exact
    r = (ld(p) * ld(q) - ld(k)).astype(np.float64)                   # fma(p, q, -k), one rounding
    assert np.abs(r).max() <= 0.5 + 2.0 ** -10          # (the extended-precision replay rounds twice; a real FMA gives <= 0.5)
    s = r * c4 + c3
    s = s * r + c2
    s = s * r + c1
    pm1 = s * r
    ki = k.astype(np.int64)
    T = np.exp2(ld(ki & 255) / 256).astype(np.float64) * np.exp2((ki >> 8).astype(np.float64))
    got = T * pm1 + T
    worst = 0.0
    for pi, qi, g in zip(p, q, got):
        ref = mp.exp(mp.mpf(float(pi)) * mp.mpf(float(qi)) * mp.log(2) / 256)      # exp of the EXACT product
        worst = max(worst, abs(float((mp.mpf(float(g)) - ref) / ref)))
    assert worst < 2.0 * 2.0 ** -52, worst

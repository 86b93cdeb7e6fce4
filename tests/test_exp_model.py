"""The constants of the device exponential `so_exps(p, q) = exp(p q ln2/256)` and of the cross-entropy link (`so_log1p_unit`,
`so_rcp_fast`) in scalaopt_b200/csrc/so_device.cuh, pinned on CPU:
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


def _log_consts():
    src = open(HDR).read()
    m = re.search(r"static __constant__ LogConsts kLog = \{([^}]*)\}", src)
    vals = [eval(v, {"__builtins__": {}}) for v in m.group(1).split(",")]      # plain arithmetic literals: 6755399441055744.0, -1.0 / 128.0, ...
    size = int(re.search(r"kLogTableSize = (\d+)", src).group(1))
    return vals, size


def test_cross_entropy_link_replayed_against_50_digit_arithmetic():
    """so_log1p_unit(e) = log(1 + e) for e = exp(-|z|) in [0, 1] (129-entry table of (1/c_j, log c_j) + degree-6 series) and the
    Newton reciprocal 1/(1 + e), replayed in numpy from the header's constants: relative error of log1p below 1e-15 on the whole
    interval INCLUDING e -> 0 (where a table-free log(1 + e) would lose every digit), two Newton steps from a 23-bit seed exact
    to an ulp."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    (shift, n128, c6, c5, c4, c3), size = _log_consts()
    assert shift == 1.5 * 2.0 ** 52 and n128 == -1.0 / 128.0 and size == 129
    assert (c6, c5, c4, c3) == (-1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0)
    rng = np.random.default_rng(12345)
    e = np.concatenate([rng.uniform(0.0, 1.0, 3000), 10.0 ** rng.uniform(-300, -2, 1000), [0.0, 1.0, 2.0 ** -8, 1.0 / 256 + 1e-12,
                        0.5 / 128, 0.5 / 128 - 1e-17, 127.5 / 128, 1.0 - 2.0 ** -53]])
    ld = np.longdouble
    kd = (ld(e) * ld(128.0) + ld(shift)).astype(np.float64)          # fma(e, 128, shift): e * 128 is exact
    j = (kd - shift).astype(np.int64)
    assert j.min() >= 0 and j.max() <= 128
    c = 1.0 + j / 128.0
    tab_x, tab_y = 1.0 / c, np.array([float(mp.log(mp.mpf(float(v)))) for v in c])     # what log_table_init() stores (libdevice log: <= 1 ulp)
    r = ((ld(j.astype(np.float64)) * ld(n128) + ld(e)).astype(np.float64)) * tab_x      # fma(j, -1/128, e) is exact, then one product
    assert np.abs(r).max() <= 2.0 ** -8 * (1 + 2.0 ** -20)     # (the extended-precision replay of the first FMA rounds twice near ties)
    q = r * c6 + c5
    q = q * r + c4
    q = q * r + c3
    q = q * r - 0.5
    got = (r * r) * q + r + tab_y
    worst = 0.0
    for ei, g in zip(e, got):
        ref = mp.log1p(mp.mpf(float(ei)))
        if ref == 0:
            assert g == 0.0
            continue
        worst = max(worst, abs(float((mp.mpf(float(g)) - ref) / ref)))
    assert worst < 1.0e-15, worst
    # reciprocal of 1 + e: seed with 23 good bits (rcp.approx.ftz.f64 keeps the upper 32 bits), two Newton steps
    x = 1.0 + e[:3000]
    seed = (1.0 / x).view(np.uint64) & np.uint64(0xFFFFFFFF00000000)
    rr = seed.view(np.float64)
    for _ in range(2):
        rr = rr + rr * (1.0 - x * rr)
    assert np.max(np.abs(rr * x - 1.0)) <= 2.0 ** -52

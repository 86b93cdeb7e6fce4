"""GPU parity of the scalar-t families (exponential decay, Gaussian peak, Gaussian peak + exponential
background, polynomial) through the C ABI against the CPU oracle on identical inputs."""
import numpy as np
import pytest

from parity import assert_gram, assert_loss, assert_vec

pytestmark = pytest.mark.gpu

CASES = [
    ("EXPDECAY", [2.0, -1.3], 0.1),            # LevenbergMarquardtSpec.scala:62-63 shape
    ("GAUSS", [2.0, 0.5, 0.05], 0.05),
    ("GAUSS_EXPBKG", [2.0, 0.5, 0.05, 1.0, 3.0], 0.05),   # BASELINE config 3
    ("POLY", [1.0], 0.1),
    ("POLY", [1.0, -2.0], 0.1),
    ("POLY", [1.0, -2.0, 0.5, 3.0], 0.1),
    ("POLY", [1.0, -2.0, 0.5, 3.0, -1.0, 0.25, 2.0, -0.5], 0.1),
]


def _upload(native, ctx, X, y):
    ds = native.Dataset(ctx, y.size, X.shape[1])
    ds.append(X, y)
    return ds


@pytest.mark.parametrize("name,truth,sigma", CASES)
@pytest.mark.parametrize("m", [200_000, 200_001])
def test_all_four_evaluations_match_oracle(native, oracle, gpu_ctx, name, truth, sigma, m):
    fam = getattr(oracle, name)
    assert fam == getattr(native, name)
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=sigma)
    ds = _upload(native, gpu_ctx, X, y)
    assert ds.size() == m
    p = np.array(truth) * 1.1 + 0.01
    n = p.size
    # K1
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    loss_only, g_none = ds.loss_grad(fam, p, want_grad=False)
    assert g_none is None
    assert_loss(loss_only, oracle.apply(fam, X, y, p))
    # K2
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr, Jtr_o, what="Jtr")
    assert_loss(rtr, rtr_o)
    assert_loss(rtr / (2 * m), loss, rtol=1e-13)            # loss = rtr/(2m)
    assert_vec(Jtr / m, grad, rtol=1e-13, what="Jtr/m vs grad")
    # K3
    d = -np.asarray(grad) + 0.05 * np.cos(np.arange(n))
    alphas = [0.0, 1.0, 2.0, 4.0, 0.37][: (5 if n <= 5 else 4)]
    phi, dphi = ds.line(fam, p, d, alphas)
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
    assert_vec(phi, phi_o, what="phi")
    assert_vec(dphi, dphi_o, rtol=1e-11, what="dphi")
    assert_loss(phi[0], loss, rtol=1e-13)
    # K4
    Hd = ds.hess_vec(fam, p, d)
    assert_vec(Hd, oracle.hess_vec(fam, X, y, p, d), rtol=1e-11, what="Hd")
    ds.free()


@pytest.mark.parametrize("m", [1, 2, 3, 5, 511, 512, 513, 1025])
def test_ragged_sizes(native, oracle, gpu_ctx, m):
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=0.05, total_rows=max(m, 8))
    ds = _upload(native, gpu_ctx, X, y)
    p = truth * 0.9
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr, Jtr_o, what="Jtr")
    ds.free()


def test_slices_at_odd_offsets_and_additivity(native, oracle, gpu_ctx):
    """so_dataset_slice: sums over disjoint row ranges add up to the whole (the aggregate's combop)."""
    fam = oracle.GAUSS
    truth = np.array([2.0, 0.5, 0.05])
    m = 100_003
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=0.05)
    ds = _upload(native, gpu_ctx, X, y)
    p = truth * 1.05
    _, _, rtr = ds.normal_eq(fam, p)
    cuts = [0, 1, 4098, 50_001, m]
    parts = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        v = ds.slice(b, e)
        assert v.size() == e - b
        JtJ, Jtr, r = v.normal_eq(fam, p)
        JtJ_o, Jtr_o, r_o = oracle.normal_eq(fam, X[b:e], y[b:e], p)
        assert_gram(JtJ, JtJ_o)
        assert_loss(r, r_o)
        parts.append(r)
        v.free()
    assert sum(parts) == pytest.approx(rtr, rel=1e-13)
    ds.free()


def test_extreme_parameters_take_the_slow_exp_path(native, oracle, gpu_ctx):
    """|arg| >= 708 (underflow to 0 / overflow) goes through libdevice exp; results still match."""
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, 4096, 1, truth, noise_sigma=0.05)
    ds = _upload(native, gpu_ctx, X, y)
    p = np.array([2.0, 0.5, 0.001, 1.0, 900.0])    # z^2/2 up to 1.2e5, lambda t up to 900
    loss, grad = ds.loss_grad(fam, p)
    assert np.isfinite(loss)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p), rtol=1e-11)
    ds.free()


def test_normal_equations_pick_the_exp_range_check_from_the_input_range(native, oracle, gpu_ctx):
    """K2 of the two-exponential family drops its per-row exp range check when the shard's input range keeps every
    argument inside (-700, 700); whichever variant runs, the sums match the oracle, and the cached range follows the data:
    same parameters, new data with far-out inputs in the same reservation -> the checked kernel again."""
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    m = 6000
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=0.05)
    ds = native.Dataset(gpu_ctx, m, 1).append(X, y)

    def check(p, X, y):
        JtJ, Jtr, rtr = ds.normal_eq(fam, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
        assert_gram(JtJ, JtJ_o); assert_vec(Jtr, Jtr_o, what="Jtr"); assert_loss(rtr, rtr_o)

    check(truth * 1.1, X, y)                                     # inputs in [0, 1]: every argument small -> check-free kernel
    check(np.array([2.0, 0.5, 0.001, 1.0, 900.0]), X, y)         # same data, arguments up to 1.2e5 -> checked kernel, slow exp
    launches = gpu_ctx.stats().total_launches
    check(truth * 0.9, X, y)
    assert gpu_ctx.stats().total_launches == launches + 1        # the range pass runs once per data-set state, not per call
    # new data in the same reservation: inputs far from the peak (exp underflows to 0 for those rows)
    X2 = X.copy(); X2[::7] = 40.0; X2[3] = -2.0
    ds.clear(); ds.append(X2, y)
    launches = gpu_ctx.stats().total_launches
    check(truth * 1.1, X2, y)                                    # (40 - 0.55)^2 / (2 0.055^2) = 2.6e5 -> checked kernel again
    assert gpu_ctx.stats().total_launches == launches + 2        # range pass + K2
    # a view sees its own range
    v = ds.slice(4, 7)
    JtJ, Jtr, rtr = v.normal_eq(fam, truth * 1.1)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X2[4:7], y[4:7], truth * 1.1)
    assert_gram(JtJ, JtJ_o); assert_vec(Jtr, Jtr_o, what="Jtr"); assert_loss(rtr, rtr_o)
    v.free()
    # a non-finite input: no usable range, the checked kernel propagates it like the reference (NaN sums)
    X3 = X.copy(); X3[11] = np.nan
    ds.clear(); ds.append(X3, y)
    JtJ, Jtr, rtr = ds.normal_eq(fam, truth * 1.1)
    assert np.isnan(rtr) and np.isnan(JtJ).any()
    ds.free()


@pytest.mark.parametrize("seed", range(16))
def test_two_exponential_family_at_random_parameters(native, oracle, gpu_ctx, seed):
    """K1-K4 and the TSQR at parameter vectors drawn around (and far from) the truth, on inputs spread over [-3, 5]:
    some draws keep every exp argument in the fast path's range (check-free kernels), some do not (checked kernels,
    libdevice exp for the far rows) — the choice must never show in the sums."""
    fam = oracle.GAUSS_EXPBKG
    rng = np.random.default_rng(1000 + seed)
    m = 4001
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X = rng.uniform(-3.0, 5.0, size=(m, 1))
    t = X[:, 0]
    y = truth[0] * np.exp(-0.5 * ((t - truth[1]) / truth[2]) ** 2) + truth[3] * np.exp(-truth[4] * t) + 0.05 * rng.standard_normal(m)
    scale = [1.0, 1.0, 10.0 ** rng.uniform(-2.5, 1.0), 1.0, 10.0 ** rng.uniform(-1.0, 1.0)]     # sigma, lambda over decades
    p = truth * np.array(scale) * rng.uniform(0.8, 1.2, size=5)
    if seed % 3 == 0:
        p[4] = -abs(p[4]) * 0.3                                  # growing background: positive exp arguments
    ds = _upload(native, gpu_ctx, X, y)
    with np.errstate(over="ignore", invalid="ignore"):
        loss_o = oracle.apply(fam, X, y, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    if not (np.isfinite(loss_o) and np.all(np.isfinite(JtJ_o)) and np.all(np.isfinite(Jtr_o)) and np.abs(JtJ_o).max() < 1e140):
        ds.free(); pytest.skip("the reference overflows (or its sums cannot be squared by the comparison) at this draw")
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, loss_o)
    assert_vec(grad, oracle.gradient(fam, X, y, p), rtol=1e-11)
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    assert np.all(np.isfinite(JtJ)) and np.all(np.isfinite(Jtr)), (p, JtJ)
    assert_gram(JtJ, JtJ_o); assert_vec(Jtr, Jtr_o, rtol=1e-11, what="Jtr"); assert_loss(rtr, rtr_o)
    d = rng.standard_normal(5) * 0.01 * np.abs(p)
    phi, dphi = ds.line(fam, p, d, [0.0, 0.5, 1.0])
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, [0.0, 0.5, 1.0])
    assert_vec(phi, phi_o, what="phi"); assert_vec(dphi, dphi_o, rtol=1e-10, what="dphi")
    assert_vec(ds.hess_vec(fam, p, d), oracle.hess_vec(fam, X, y, p, d), rtol=1e-10, what="Hd")
    R = ds.qr(fam, p)
    G = np.block([[JtJ_o, Jtr_o[:, None]], [Jtr_o[None, :], np.array([[rtr_o]])]])
    sc = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    assert np.max(np.abs(R.T @ R - G) / sc) < 1e-10
    ds.free()


@pytest.mark.parametrize("name", ["GAUSS", "EXPDECAY", "POLY"])
@pytest.mark.parametrize("seed", range(6))
def test_other_single_input_families_at_random_parameters(native, oracle, gpu_ctx, name, seed):
    """same idea for the one-exponential families and the polynomial: inputs over [-3, 5], widths / rates over decades"""
    fam = getattr(oracle, name)
    rng = np.random.default_rng(2000 + seed)
    m = 3001
    t = rng.uniform(-3.0, 5.0, size=m)
    if name == "GAUSS":
        truth = np.array([2.0, 0.5, 0.3])
        y = truth[0] * np.exp(-0.5 * ((t - truth[1]) / truth[2]) ** 2)
        p = truth * np.array([1.0, 1.0, 10.0 ** rng.uniform(-2.5, 1.0)]) * rng.uniform(0.8, 1.2, size=3)
    elif name == "EXPDECAY":
        truth = np.array([2.0, -1.3])
        y = truth[0] * np.exp(truth[1] * t)
        p = truth * np.array([1.0, 10.0 ** rng.uniform(-1.0, 1.3)]) * rng.uniform(0.8, 1.2, size=2) * (1.0 if seed % 2 else -1.0)
    else:
        truth = np.array([1.0, -2.0, 0.5, 3.0, -0.25])
        y = sum(c * t ** k for k, c in enumerate(truth))
        p = truth * 10.0 ** rng.uniform(-2.0, 2.0, size=5) * rng.choice([-1.0, 1.0], size=5)
    y = y + 0.05 * rng.standard_normal(m)
    X = t.reshape(-1, 1)
    n = p.size
    with np.errstate(over="ignore", invalid="ignore"):
        loss_o = oracle.apply(fam, X, y, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    if not (np.isfinite(loss_o) and np.all(np.isfinite(JtJ_o)) and np.all(np.isfinite(Jtr_o)) and np.abs(JtJ_o).max() < 1e140):
        pytest.skip("the reference overflows at this draw")
    ds = _upload(native, gpu_ctx, X, y)
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, loss_o)
    assert_vec(grad, oracle.gradient(fam, X, y, p), rtol=1e-11)
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    assert_gram(JtJ, JtJ_o); assert_vec(Jtr, Jtr_o, rtol=1e-11, what="Jtr"); assert_loss(rtr, rtr_o)
    d = rng.standard_normal(n) * 0.01 * np.abs(p)
    phi, dphi = ds.line(fam, p, d, [0.0, 0.5, 1.0])
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, [0.0, 0.5, 1.0])
    assert_vec(phi, phi_o, what="phi"); assert_vec(dphi, dphi_o, rtol=1e-10, what="dphi")
    assert_vec(ds.hess_vec(fam, p, d), oracle.hess_vec(fam, X, y, p, d), rtol=1e-10, what="Hd")
    R = ds.qr(fam, p)
    G = np.block([[JtJ_o, Jtr_o[:, None]], [Jtr_o[None, :], np.array([[rtr_o]])]])
    sc = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    assert np.max(np.abs(R.T @ R - G) / sc) < 1e-10
    ds.free()


def test_errors_are_reported_not_swallowed(native, gpu_ctx):
    ds = native.Dataset(gpu_ctx, 16, 1)
    with pytest.raises(native.NativeError):       # empty data set
        ds.loss_grad(native.GAUSS, [1.0, 0.5, 0.1])
    ds.append(np.linspace(0, 1, 16).reshape(-1, 1), np.ones(16))
    with pytest.raises(native.NativeError):       # wrong parameter count
        ds.loss_grad(native.GAUSS, [1.0, 0.5])
    with pytest.raises(native.NativeError):       # does not fit
        ds.append(np.zeros((1, 1)), np.zeros(1))
    with pytest.raises(native.NativeError):       # polynomial degree outside the catalogue
        ds.loss_grad(native.POLY, np.ones(9))
    x, y = ds.head()
    assert x[0] == 0.0 and y == 1.0
    ds.free()


def test_device_generator_follows_the_oracle_specification(native, oracle, gpu_ctx):
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    m = 50_000
    ds = native.Dataset(gpu_ctx, m, 1).generate(fam, truth, seed=12345, noise_sigma=0.05)
    Xd, yd = ds.read(0, m)
    Xo, yo = oracle.generate(fam, m, 1, truth, seed=12345, noise_sigma=0.05)
    np.testing.assert_array_equal(Xd, Xo)                 # the t grid is exact
    np.testing.assert_allclose(yd, yo, rtol=0, atol=1e-14)  # libdevice vs libm log/cos/exp: a few ulp
    # and the evaluation on the device-generated bits matches the oracle fed the very same bits
    p = truth * 1.1
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, Xd, yd, p)
    assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr, Jtr_o, what="Jtr")
    ds.free()


def test_results_are_bitwise_reproducible(native, oracle, gpu_ctx):
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    ds = native.Dataset(gpu_ctx, 1_000_003, 1).generate(fam, truth, noise_sigma=0.05)
    a = ds.normal_eq(fam, truth * 1.1)
    for _ in range(3):
        b = ds.normal_eq(fam, truth * 1.1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    st = gpu_ctx.stats()
    assert st.launches == 1 and st.rows == 1_000_003 and st.bytes == 16 * 1_000_003 and st.kernel_ms > 0
    ds.free()

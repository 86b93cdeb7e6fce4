"""Negative log-likelihood objective: oracle restatement on the reference example's own sample (CPU)
and GPU parity through the C ABI."""
import numpy as np
import pytest

TRUTH = np.array([0.3, 125.0, 10.0, 0.05])      # (fSig, mu, sigma, lambda), ExSignalPlusBackgroundFit.scala:40-45
P0 = np.array([0.05, 125.0, 10.0, 1.0 / 15.0])  # :58
XMIN = [50.0]


def test_oracle_nll_on_the_reference_example(oracle):
    x = oracle.nll_example_data()                # Random(12345), 2000 variates
    assert x.size == 2000 and x.min() >= 50.0
    # the objective is sum -2 log pdf (not a mean), and the pdf integrates to ~1 over [xMin, inf)
    assert oracle.nll(oracle.PDF_GAUSS_EXPBKG, x, TRUTH, XMIN) < oracle.nll(oracle.PDF_GAUSS_EXPBKG, x, P0, XMIN)
    v = [oracle.pdf(0, TRUTH, XMIN, t) for t in np.linspace(50.0, 600.0, 55001)]
    assert np.trapezoid(v, dx=0.01) == pytest.approx(1.0, abs=2e-3)
    # fold order only changes rounding
    a = oracle.nll(0, x, TRUTH, XMIN)
    assert oracle.nll(0, x, TRUTH, XMIN, parts=4, threads=4) == pytest.approx(a, rel=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("m", [2000, 200_001])
def test_gpu_nll_matches_oracle(native, oracle, gpu_ctx, m):
    if m == 2000:
        x = oracle.nll_example_data()
    else:
        rng = np.random.default_rng(12345)
        sig = rng.random(m) < 0.3
        x = np.where(sig, rng.normal(125.0, 10.0, m), 50.0 - np.log(rng.random(m)) / 0.05)
    ds = native.Dataset(gpu_ctx, m, 1).append(x.reshape(-1, 1), np.zeros(m))
    for p in (TRUTH, P0, TRUTH * 1.07):
        got = ds.nll(native.PDF_GAUSS_EXPBKG, p, XMIN)
        ref = oracle.nll(oracle.PDF_GAUSS_EXPBKG, x, p, XMIN)
        assert abs(got - ref) <= 1e-12 * abs(ref)
    st = gpu_ctx.stats()
    assert st.bytes == 8 * m and st.launches == 1          # only the x column is read
    # a pdf that underflows to 0 somewhere gives +inf like the reference; a negative pdf gives NaN
    assert ds.nll(native.PDF_GAUSS_EXPBKG, [1.0, 125.0, 1e-3, 0.05], XMIN) == np.inf
    assert np.isnan(ds.nll(native.PDF_GAUSS_EXPBKG, [1.5, 125.0, 10.0, 0.05], XMIN))
    with pytest.raises(native.NativeError):
        ds.nll(native.PDF_GAUSS_EXPBKG, TRUTH[:3], XMIN)
    ds.free()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_nll_at_random_parameters(native, oracle, gpu_ctx, seed):
    """widths and rates over decades: some draws keep both exp arguments in the fast path's range (check-free kernel), some
    underflow the Gaussian for most events (checked kernel, libdevice exp); finite or infinite, the sums agree"""
    rng = np.random.default_rng(4000 + seed)
    m = 5003
    sig = rng.random(m) < 0.3
    x = np.where(sig, rng.normal(125.0, 10.0, m), 50.0 - np.log(rng.random(m)) / 0.05)
    p = [rng.uniform(0.05, 0.95), 125.0 * rng.uniform(0.9, 1.1), 10.0 * 10.0 ** rng.uniform(-1.5, 1.0), 0.05 * 10.0 ** rng.uniform(-1.5, 1.2)]
    ds = native.Dataset(gpu_ctx, m, 1).append(x.reshape(-1, 1), np.zeros(m))
    got = ds.nll(native.PDF_GAUSS_EXPBKG, p, XMIN)
    ref = oracle.nll(oracle.PDF_GAUSS_EXPBKG, x, p, XMIN)
    if np.isfinite(ref):
        assert abs(got - ref) <= 1e-12 * abs(ref), (p, got, ref)
    else:
        assert got == ref or (np.isnan(got) and np.isnan(ref)), (p, got, ref)
    ds.free()


def test_max_likelihood_fit_with_nelder_mead(oracle):
    """ExSignalPlusBackgroundFit.scala:56-66: maxLikelihoodFit(pdf, data, p0) with its default method NelderMead."""
    from scalaopt_b200 import host
    x = oracle.nll_example_data()
    f = host.Objective.from_callables(4, lambda p: oracle.nll(0, x, p, XMIN))
    xmin, res = host.minimize(host.NELDER_MEAD, f, P0)
    assert f(xmin) < f(TRUTH)                       # the fit beats the generating point on its own sample
    np.testing.assert_allclose(xmin, TRUTH, rtol=0.12)
    assert f.counters()["gradient_calls"] == 0      # derivative free: only f(x) is consumed


@pytest.mark.gpu
def test_max_likelihood_fit_on_the_gpu(native, oracle, gpu_ctx):
    """The same fit with every f(x) one so_eval_nll launch over the resident sample; the simplex path only depends on
    comparisons of f values, so GPU and oracle objectives give the same minimiser."""
    from scalaopt_b200 import host
    x = oracle.nll_example_data(nb_events=20000)
    ds = native.Dataset(gpu_ctx, x.size, 1).append(x.reshape(-1, 1), np.zeros(x.size))
    f_gpu = host.Objective.gpu_nll(gpu_ctx, ds, native.PDF_GAUSS_EXPBKG, 4, XMIN)
    f_ref = host.Objective.from_callables(4, lambda p: oracle.nll(0, x, p, XMIN))
    assert f_gpu(P0) == pytest.approx(f_ref(P0), rel=1e-12)
    xg, rg = host.minimize(host.NELDER_MEAD, f_gpu, P0)
    xr, rr = host.minimize(host.NELDER_MEAD, f_ref, P0)
    assert f_gpu.counters()["loss_calls"] > rg.iterations and gpu_ctx.stats().launches == 1
    np.testing.assert_allclose(xg, xr, rtol=1e-4)
    np.testing.assert_allclose(xg, TRUTH, rtol=0.05)
    f_gpu.free(); ds.free()

"""GPU end-to-end parity: the reference's optimizers (host mirror) driven by GpuMSEFunction versus the same
optimizers driven by the CPU oracle on identical inputs — same iteration count, minimisers within 1e-8
(BASELINE.json north_star)."""
import numpy as np
import pytest

from refdata import lm_spec_data, logistic_spec_data, oracle_objective

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from scalaopt_b200 import host as h
    h.load()
    return h


def _gpu(native, host, ctx, fam, X, y, n, tsqr=False, speculate=True):
    X = np.asarray(X, float).reshape(len(y), -1)
    ds = native.Dataset(ctx, len(y), X.shape[1]).append(X, y)
    return ds, host.Objective.gpu(ctx, ds, fam, n, tsqr=tsqr, speculate=speculate)


def test_gpu_objective_answers_the_objective_function_interface(native, oracle, host, gpu_ctx):
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, 50_000, 1, truth, noise_sigma=0.05)
    ds, f = _gpu(native, host, gpu_ctx, fam, X, y, 5)
    p = truth * 1.1
    d = np.array([0.1, -0.01, 0.002, 0.05, -0.2])
    assert f(p) == pytest.approx(oracle.apply(fam, X, y, p), rel=1e-12)            # apply = sum r^2/2 / m
    np.testing.assert_allclose(f.gradient(p), oracle.gradient(fam, X, y, p), rtol=1e-11)
    assert f.dirder(p, d) == pytest.approx(float(oracle.gradient(fam, X, y, p) @ d), rel=1e-10)
    np.testing.assert_allclose(f.dirHessian(p, d), oracle.hess_vec(fam, X, y, p, d), rtol=1e-10)
    c = f.counters()
    assert c["k1_passes"] == 1 and c["k4_passes"] == 1 and c["cache_hits"] == 2    # one pass per distinct point
    # jacobianAndResidualsMatrix: n+1 rows whose QR equals the QR of the m-row system up to row signs
    rows = f.jacobianAndResidualsMatrix(p)
    assert rows.shape == (6, 6)
    q_gpu = host.qr(rows, 5, pivoting=False)
    q_ref = oracle.qr(oracle.jacobian_matrix(fam, X, y, p), 5, pivoting=False)
    np.testing.assert_allclose(np.abs(q_gpu["R"]), np.abs(q_ref.Rmat()), rtol=1e-9, atol=1e-9 * np.abs(q_ref.Rmat()).max())
    np.testing.assert_allclose(np.abs(q_gpu["qtb"]), np.abs(q_ref.qtb), rtol=1e-8, atol=1e-10 * np.abs(q_ref.qtb).max())
    np.testing.assert_allclose(q_gpu["acNorms"], q_ref.acnorms, rtol=1e-12)
    assert q_gpu["bNorm"] == pytest.approx(q_ref.bnorm, rel=1e-12)
    f.free(); ds.free()


CASES = ["gauss_expbkg", "lm_spec_linear", "lm_spec_exponential", "expdecay_large", "poly", "linear_small"]


@pytest.mark.parametrize("mode", ["k2", "k2-speculative", "tsqr"])
@pytest.mark.parametrize("case", CASES)
def test_levenberg_marquardt_trajectory_matches_oracle(native, oracle, host, gpu_ctx, case, mode):
    """k2: K2 + Cholesky, trial-point losses from K1; k2-speculative (the default): trial-point losses from K2 passes whose
    sums serve the next Jacobian request; tsqr: the (n+1)-row system is the R factor from so_eval_qr."""
    tsqr, speculate = mode == "tsqr", mode != "k2"
    if case == "gauss_expbkg":         # BASELINE config 3 at oracle-sized m
        fam, n = oracle.GAUSS_EXPBKG, 5
        truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
        X, y = oracle.generate(fam, 200_000, 1, truth, noise_sigma=0.05)
        start = truth * 1.1
    elif case == "linear_small":      # linear regression with intercept, n = 4: the narrow-row TSQR (so_wide_qr.cu)
        fam, n = oracle.LINEAR, 4
        X, y = oracle.generate(fam, 30_000, 3, [80.0, 0.0, 1.0, 2.0])
        start = np.zeros(4)
    elif case == "lm_spec_linear":
        (X, y, start), _ = lm_spec_data(oracle)
        fam, n = oracle.LINEAR, 11
    elif case == "lm_spec_exponential":
        _, (X, y, _) = lm_spec_data(oracle)
        fam, n, start = oracle.EXPDECAY, 2, np.array([4.0, 0.5])
    elif case == "expdecay_large":
        fam, n = oracle.EXPDECAY, 2
        X, y = oracle.generate(fam, 100_001, 1, [2.0, -1.3], noise_sigma=0.1)
        start = np.array([2.5, -1.0])
    else:
        fam, n = oracle.POLY, 4
        X, y = oracle.generate(fam, 50_000, 1, [1.0, -2.0, 0.5, 3.0], noise_sigma=0.1)
        start = np.array([0.0, 0.0, 0.0, 0.0])
    ds, f_gpu = _gpu(native, host, gpu_ctx, fam, X, y, n, tsqr=tsqr, speculate=speculate)
    f_ref = oracle_objective(host, oracle, fam, X, y, n)
    x_gpu, r_gpu = host.minimize(host.LEVENBERG_MARQUARDT, f_gpu, start)
    x_ref, r_ref = host.minimize(host.LEVENBERG_MARQUARDT, f_ref, start)
    assert r_gpu.iterations == r_ref.iterations
    assert r_gpu.inner_iterations == r_ref.inner_iterations
    np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-8, atol=1e-8)
    c = f_gpu.counters()
    if speculate:
        # the first trial point after each Jacobian is a K2 pass (n <= 16); an accepted step's Jacobian request is served
        # from it; later trial points of the same outer iteration are plain K1 loss passes
        assert c["k1_passes"] + c["spec_passes"] == r_gpu.inner_iterations
        assert c["k2_passes"] + c["spec_hits"] == r_gpu.iterations + c["spec_passes"]
        assert c["spec_passes"] == r_gpu.iterations and c["spec_hits"] >= min(1, r_gpu.iterations - 1)
        assert tsqr or c["k1_passes"] == r_gpu.inner_iterations - r_gpu.iterations
    else:
        assert c["k2_passes"] == r_gpu.iterations            # one normal-equation pass per outer iteration
        assert c["k1_passes"] == r_gpu.inner_iterations      # one loss pass per trial point
    f_gpu.free(); ds.free()


@pytest.mark.parametrize("mode", ["k2", "k2-speculative", "tsqr"])
def test_levenberg_marquardt_logistic_regression_matches_oracle(native, oracle, host, gpu_ctx, mode):
    """FFNeuralNetworkTrainerSpec.scala:320-351 (logistic regression by LM) on the GPU.  The loss of this family is the
    mean cross-entropy, NOT r^T r / 2m: with TSQR (or speculation) on, every loss LM sees must still be the K1 loss."""
    X, y, truth = logistic_spec_data(oracle)
    ds, f_gpu = _gpu(native, host, gpu_ctx, oracle.LOGISTIC, X, y, 3, tsqr=mode == "tsqr", speculate=mode != "k2")
    f_ref = oracle_objective(host, oracle, oracle.LOGISTIC, X, y, 3)
    start = [-0.5, 0.6, 0.6]
    x_gpu, r_gpu = host.minimize(host.LEVENBERG_MARQUARDT, f_gpu, start)
    x_ref, r_ref = host.minimize(host.LEVENBERG_MARQUARDT, f_ref, start)
    assert r_gpu.iterations == r_ref.iterations
    assert r_gpu.inner_iterations == r_ref.inner_iterations
    np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-8, atol=1e-8)
    c = f_gpu.counters()
    assert c["spec_passes"] == 0                             # nothing to speculate on: the loss is not a function of r^T r
    assert c["k1_passes"] == r_gpu.inner_iterations          # one cross-entropy loss pass per trial point
    # every loss the optimizer saw is the cross-entropy at that point
    assert f_gpu(x_gpu) == pytest.approx(oracle.apply(oracle.LOGISTIC, X, y, x_gpu), rel=1e-12)
    f_gpu.free(); ds.free()


def test_bfgs_logistic_regression_matches_oracle(native, oracle, host, gpu_ctx):
    """FFNeuralNetworkTrainerSpec.scala:189-220 on the GPU."""
    X, y, truth = logistic_spec_data(oracle)
    ds, f_gpu = _gpu(native, host, gpu_ctx, oracle.LOGISTIC, X, y, 3)
    f_ref = oracle_objective(host, oracle, oracle.LOGISTIC, X, y, 3, exact_hessvec=False)
    start = [-0.5, 0.6, 0.6]
    x_gpu, r_gpu = host.minimize(host.BFGS, f_gpu, start)
    x_ref, r_ref = host.minimize(host.BFGS, f_ref, start)
    assert r_gpu.iterations == r_ref.iterations == 12
    np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-8, atol=1e-8)
    assert np.linalg.norm(x_gpu - truth) / np.linalg.norm(truth) <= 0.05
    c = f_gpu.counters()
    # the reference asks 24 losses + 48 gradients on 13 distinct points; memoisation makes that 13 fused passes
    assert c["loss_calls"] == 24 and c["k1_passes"] + c["k3_passes"] <= 14
    f_gpu.free(); ds.free()


@pytest.mark.parametrize("method", ["BFGS", "CONJUGATE_GRADIENT", "NEWTON_CG"])
def test_gradient_methods_on_wide_linear_least_squares(native, oracle, host, gpu_ctx, method):
    fam = oracle.LINEAR_NOINT
    f_, m = 32, 20_000
    truth = np.arange(1, f_ + 1) / f_
    X, y = oracle.generate(fam, m, f_, truth)
    ds, f_gpu = _gpu(native, host, gpu_ctx, fam, X, y, f_)
    f_ref = oracle_objective(host, oracle, fam, X, y, f_)
    meth = getattr(host, method)
    x_gpu, r_gpu = host.minimize(meth, f_gpu, np.zeros(f_))
    x_ref, r_ref = host.minimize(meth, f_ref, np.zeros(f_))
    assert r_gpu.iterations == r_ref.iterations
    np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-8, atol=1e-8)
    f_gpu.free(); ds.free()


def test_gpu_dataset_mirrors_the_dataset_trait(native, oracle, host, gpu_ctx):
    X, y = oracle.generate(oracle.LINEAR, 1000, 3, [1.0, 2.0, 3.0, 4.0])
    ds = native.Dataset(gpu_ctx, 1000, 3).append(X, y)
    assert ds.size() == 1000
    x0, y0 = ds.head()
    np.testing.assert_array_equal(x0, X[0]); assert y0 == y[0]
    Xr, yr = ds.read(10, 5)
    np.testing.assert_array_equal(Xr, X[10:15]); np.testing.assert_array_equal(yr, y[10:15])
    ds.free()


@pytest.mark.parametrize("tsqr", [False, True])
@pytest.mark.parametrize("case", ["linear_spec", "no_origin", "wide"])
def test_linear_lm_on_the_gpu(native, oracle, host, gpu_ctx, case, tsqr):
    """linear.lm(data, addOrigin) (linear/package.scala:54-63; LinearSpec.scala:32-54) on a GpuDataSet: one pass over the
    resident rows (K2 / TSQR) + the reference's QR.solution on n+1 rows, against the reference's 2n+1-pass QR of the
    materialised rows (oracle) and, for the spec case, the spec's own +-0.5 on every coefficient."""
    if case == "linear_spec":
        (X, y, truth), _ = lm_spec_data(oracle)
        origin = True
    elif case == "no_origin":
        truth = np.array([0.5, -1.0, 2.0, 0.25])
        X, y = oracle.generate(oracle.LINEAR_NOINT, 100_003, 4, truth)      # y takes both signs
        origin = False
    else:
        truth = np.concatenate([[3.0], np.linspace(-1.0, 1.0, 32)])            # cfg5's n = 33 shape (with origin)
        X, y = oracle.generate(oracle.LINEAR, 20_001, 32, truth)
        origin = True
    n = X.shape[1] + int(origin)
    ds = native.Dataset(gpu_ctx, len(y), X.shape[1]).append(X, y)
    launches0 = gpu_ctx.stats().total_launches
    beta = host.gpu_lm(gpu_ctx, ds, add_origin=origin, tsqr=tsqr)
    assert gpu_ctx.stats().total_launches - launches0 == 1                     # one pass over the data
    a = np.hstack([np.ones((len(y), 1)), X]) if origin else X
    beta_ref = oracle.qr(np.hstack([a, y[:, None]]), n).solution()              # lm(data) as the reference computes it
    assert beta.size == n
    np.testing.assert_allclose(beta, beta_ref, rtol=1e-9, atol=1e-9)
    if case == "linear_spec":
        np.testing.assert_allclose(beta, truth, rtol=0, atol=0.5)              # LinearSpec.scala:47-49
    ds.free()

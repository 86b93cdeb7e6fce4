"""The FFNeuralNetworkTrainer device family: FFNeuralNetwork(Vector(f, h, 1)) forward/backward on the
GPU against the oracle's restatement of FFNeuralNetwork.forward/backward (FFNeuralNetwork.scala:114-189)."""
import numpy as np
import pytest

from parity import assert_gram, assert_loss, assert_vec

pytestmark = pytest.mark.gpu

SHAPES = [(2, 2), (1, 1), (1, 4), (3, 2), (4, 4), (2, 1), (4, 1)]


def _data(fam_ce, m, f, h, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((m, f))
    truth = rng.uniform(-1.0, 1.0, h * (f + 1) + h + 1)
    hid = 1.0 / (1.0 + np.exp(-(truth[: h * (f + 1)].reshape(h, f + 1)[:, 0] + X @ truth[: h * (f + 1)].reshape(h, f + 1)[:, 1:].T)))
    out = truth[h * (f + 1)] + hid @ truth[h * (f + 1) + 1:]
    if fam_ce:
        y = (rng.random(m) < 1.0 / (1.0 + np.exp(-out))).astype(float)
    else:
        y = out + 0.1 * rng.standard_normal(m)
    return X, y, truth


@pytest.mark.parametrize("f,h", SHAPES)
@pytest.mark.parametrize("loss", ["MLP_MSE", "MLP_CE"])
@pytest.mark.parametrize("m", [50_001, 3])
def test_mlp_family_matches_oracle(native, oracle, gpu_ctx, f, h, loss, m):
    fam = getattr(oracle, loss)
    assert fam == getattr(native, loss)
    X, y, truth = _data(loss == "MLP_CE", m, f, h)
    if loss == "MLP_CE" and m > 3:
        y[:5] = [0.25, 0.5, 0.75, 0.0, 1.0]             # fractional targets take the y log y terms
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.8 + 0.05
    n = p.size
    loss_v, grad = ds.loss_grad(fam, p)
    assert_loss(loss_v, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    l0, g0 = ds.loss_grad(fam, p, want_grad=False)
    assert g0 is None
    assert_loss(l0, oracle.apply(fam, X, y, p))
    # K3 / K4 are composed from loss+gradient launches exactly as the reference composes them
    d = -np.asarray(grad) + 0.01 * np.cos(np.arange(n))
    alphas = [0.0, 1.0, 0.37]
    phi, dphi = ds.line(fam, p, d, alphas)
    assert gpu_ctx.stats().launches == len(alphas)
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
    assert_vec(phi, phi_o, what="phi")
    assert_vec(dphi, dphi_o, rtol=1e-11, what="dphi")
    Hd = ds.hess_vec(fam, p, d)
    Hd_o = oracle.hess_vec_fd(fam, X, y, p, d)
    np.testing.assert_allclose(Hd, Hd_o, rtol=0, atol=1e-6 * max(1.0, np.abs(Hd_o).max()))   # forward difference, eps = 1e-8
    # K2: Gram of the Neuron.jacobian rows and the residual column
    if n <= 9:
        JtJ, Jtr, rtr = ds.normal_eq(fam, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
        assert_gram(JtJ, JtJ_o)
        assert_vec(Jtr, Jtr_o, what="Jtr")
        assert_loss(rtr, rtr_o)
    else:
        with pytest.raises(native.NativeError) as e:
            ds.normal_eq(fam, p)
        assert e.value.code == -5
    ds.free()


def test_mlp_shapes_outside_the_catalogue_are_errors(native, gpu_ctx):
    ds = native.Dataset(gpu_ctx, 4, 5).append(np.ones((4, 5)), np.ones(4))
    with pytest.raises(native.NativeError) as e:          # f = 5 > 4
        ds.loss_grad(native.MLP_MSE, np.ones(2 * 7 + 1))
    assert e.value.code == -5
    with pytest.raises(native.NativeError):               # n is not h (f + 2) + 1
        ds.loss_grad(native.MLP_MSE, np.ones(9))
    ds.free()


def test_train_a_regression_network_with_levenberg_marquardt_and_bfgs(native, oracle, gpu_ctx):
    """FFNeuralNetworkTrainerSpec.scala:158-186,288-317 in spirit: Vector(2, 2, 1), MSE loss, data that mimic a
    network; the same optimizer trajectory whether the trainer is evaluated by the oracle or on the GPU."""
    from scalaopt_b200 import host
    from refdata import oracle_objective
    f, h, m = 2, 2, 20_000
    X, y, truth = _data(False, m, f, h, seed=5)
    start = truth + 0.1 * np.cos(np.arange(truth.size))
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    for method, kw in ((host.LEVENBERG_MARQUARDT, {}), (host.BFGS, {"tol": 1e-6})):
        f_gpu = host.Objective.gpu(gpu_ctx, ds, native.MLP_MSE, truth.size)
        f_ref = oracle_objective(host, oracle, oracle.MLP_MSE, X, y, truth.size, exact_hessvec=False)
        cfg = host.default_config(method, **kw)
        x_gpu, r_gpu = host.minimize(method, f_gpu, start, cfg)
        x_ref, r_ref = host.minimize(method, f_ref, start, cfg)
        assert r_gpu.iterations == r_ref.iterations
        np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-7, atol=1e-7)
        assert f_gpu(x_gpu) < 0.6 * f_gpu(start)        # (the weights themselves are not identifiable on noisy data)
        f_gpu.free()
    ds.free()


def test_train_a_classification_network_with_bfgs(native, oracle, gpu_ctx):
    """cross-entropy loss, logistic output (FFNeuralNetworkTrainerSpec.scala:74-117 shape Vector(2, 2, 1))"""
    from scalaopt_b200 import host
    from refdata import oracle_objective
    f, h, m = 2, 2, 20_000
    X, y, truth = _data(True, m, f, h, seed=7)
    truth = truth * 2.0
    start = truth + 0.2 * np.sin(np.arange(truth.size))
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    f_gpu = host.Objective.gpu(gpu_ctx, ds, native.MLP_CE, truth.size)
    f_ref = oracle_objective(host, oracle, oracle.MLP_CE, X, y, truth.size, exact_hessvec=False)
    cfg = host.default_config(host.BFGS, tol=1e-3)
    outcome = []
    for fn in (f_gpu, f_ref):
        try:
            outcome.append(host.minimize(host.BFGS, fn, start, cfg))
        except host.MaxIterException:
            outcome.append(None)
    assert (outcome[0] is None) == (outcome[1] is None)     # same Try outcome on both sides
    if outcome[0] is not None:
        (x_gpu, r_gpu), (x_ref, r_ref) = outcome
        assert r_gpu.iterations == r_ref.iterations
        np.testing.assert_allclose(x_gpu, x_ref, rtol=1e-6, atol=1e-6)
        assert f_gpu(x_gpu) <= f_gpu(start)
    f_gpu.free(); ds.free()

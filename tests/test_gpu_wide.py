"""GPU parity of the linear-predictor families (linear +- intercept, logistic) through the C ABI against the
CPU oracle on identical inputs: every lane mapping (f odd/even, 1..32 lanes per row) and every kernel."""
import numpy as np
import pytest

from parity import assert_gram, assert_loss, assert_vec

pytestmark = pytest.mark.gpu


def _truth(fam_name, f):
    if fam_name == "LINEAR":
        return np.concatenate([[80.0], np.arange(f, dtype=float)])          # LevenbergMarquardtSpec.scala:45
    if fam_name == "LINEAR_NOINT":
        return np.arange(1, f + 1, dtype=float) / f
    return np.array([(((j % 7) - 3) / 4.0) for j in range(f + 1)])          # logistic, SURVEY.md §8d cfg2


CASES = [
    ("LINEAR", 1), ("LINEAR", 3), ("LINEAR", 4), ("LINEAR", 10), ("LINEAR", 31), ("LINEAR", 32), ("LINEAR", 127),
    ("LINEAR_NOINT", 2), ("LINEAR_NOINT", 8), ("LINEAR_NOINT", 46), ("LINEAR_NOINT", 128), ("LINEAR_NOINT", 256),
    ("LINEAR_NOINT", 511), ("LINEAR_NOINT", 512),
    ("LOGISTIC", 2), ("LOGISTIC", 7), ("LOGISTIC", 32), ("LOGISTIC", 33), ("LOGISTIC", 64),
]


@pytest.mark.parametrize("name,f", CASES)
def test_loss_gradient_line_hessvec_match_oracle(native, oracle, gpu_ctx, name, f):
    fam = getattr(oracle, name)
    truth = _truth(name, f)
    m = 40_003 if f <= 64 else 6_001
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    n = p.size
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    loss2, _ = ds.loss_grad(fam, p, want_grad=False)
    assert_loss(loss2, loss, rtol=1e-15)
    d = -np.asarray(grad) / max(1.0, np.abs(grad).max()) + 0.01 * np.cos(np.arange(n))
    for alphas in ([1.0], [0.0, 1.0], [0.0, 1.0, 2.0, 4.0, 8.0], [0.0, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0, 0.37, 0.1, 3.0, 1.5]):
        phi, dphi = ds.line(fam, p, d, alphas)
        if name == "LOGISTIC" and max(alphas) > 4.0:
            continue   # the oracle's o = 1/(1+exp(-z)) saturates to 1.0 and its log((1-o)/..) overflows; skip
        phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
        assert_vec(phi, phi_o, what=f"phi k={len(alphas)}")
        assert_vec(dphi, dphi_o, rtol=1e-11, what=f"dphi k={len(alphas)}")
    Hd = ds.hess_vec(fam, p, d)
    assert_vec(Hd, oracle.hess_vec(fam, X, y, p, d), rtol=1e-11, what="Hd")
    ds.free()


GRAM_CASES = [("LINEAR", 1), ("LINEAR", 3), ("LINEAR", 7), ("LINEAR_NOINT", 8), ("LINEAR", 8), ("LINEAR", 10),
              ("LINEAR", 31), ("LINEAR", 32), ("LINEAR_NOINT", 33), ("LINEAR_NOINT", 47), ("LINEAR", 46),
              ("LOGISTIC", 2), ("LOGISTIC", 7), ("LOGISTIC", 32),
              # n + 1 > 48: warp-specialised shared-memory staged DMMA SYRK
              ("LINEAR_NOINT", 48), ("LINEAR_NOINT", 64), ("LINEAR", 127), ("LINEAR_NOINT", 128), ("LINEAR", 256),
              ("LINEAR_NOINT", 256), ("LOGISTIC", 64), ("LINEAR_NOINT", 511), ("LINEAR", 512),
              # ragged edge patches (T % 4 != 0), odd f (8-byte copies), one patch shared by several warps
              ("LINEAR", 100), ("LOGISTIC", 255), ("LINEAR_NOINT", 72), ("LINEAR", 50), ("LOGISTIC", 47)]


@pytest.mark.parametrize("name,f", GRAM_CASES)
@pytest.mark.parametrize("m", [30_001, 5])
def test_normal_equations_match_oracle(native, oracle, gpu_ctx, name, f, m):
    if f > 100 and m > 1000:
        m = 3_001 if f > 300 else 6_001       # keep the O(m n^2) oracle fold in seconds
    fam = getattr(oracle, name)
    truth = _truth(name, f)
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    if name == "LOGISTIC":
        # the reference forms (x * error) / residual with error == residual (Neuron.scala:67-73): 1-ulp noise per entry
        np.testing.assert_allclose(JtJ, JtJ_o, rtol=1e-11, atol=1e-11 * np.abs(JtJ_o).max())
        assert np.array_equal(JtJ, JtJ.T)
    else:
        assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr, Jtr_o, what="Jtr")
    assert_loss(rtr, rtr_o)
    ds.free()


def test_shapes_beyond_the_catalogue_are_an_error_not_a_fallback(native, gpu_ctx):
    ds = native.Dataset(gpu_ctx, 4, 513)
    ds.append(np.ones((4, 513)), np.ones(4))
    with pytest.raises(native.NativeError) as e:          # f > SO_MAX_FEATURES
        ds.normal_eq(native.LINEAR_NOINT, np.ones(513))
    assert e.value.code == -5
    ds.free()


def test_device_generated_logistic_data_is_usable(native, oracle, gpu_ctx):
    fam = oracle.LOGISTIC
    truth = _truth("LOGISTIC", 32)
    m = 100_000
    ds = native.Dataset(gpu_ctx, m, 32).generate(fam, truth, seed=12345)
    X, y = ds.read(0, m)
    assert set(np.unique(y)) <= {0.0, 1.0}
    assert abs(X.mean()) < 0.01 and abs(X.std() - 1.0) < 0.01
    p = np.zeros(33)
    loss, grad = ds.loss_grad(fam, p)
    assert loss == pytest.approx(np.log(2.0), rel=1e-12)       # CE at p = 0
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    ds.free()

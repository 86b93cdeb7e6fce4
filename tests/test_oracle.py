"""CPU tests of the oracle itself: pinned against the reference's golden vectors where they exist,
and internally cross-checked (analytic vs the reference's finite-difference path) elsewhere."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_java_util_random_first_draws(oracle):
    # new Random(12345).nextDouble() — every reference spec seeds with 12345
    r = oracle.JRandom(12345)
    assert r.next_double() == 0.3618031071604718
    g = [r.next_gaussian() for _ in range(4)]
    assert all(np.isfinite(g)) and len(set(g)) == 4


@pytest.mark.parametrize("mode", ["no_pivot", "pivot"])
def test_qr_matches_qrspec_goldens(oracle, mode):
    gold = json.load(open(os.path.join(HERE, "golden", "qrspec.json")))
    ab = np.array(gold["ab"])
    q = oracle.qr(ab, 3, pivoting=(mode == "pivot"))
    tol = gold["tol"]
    assert list(q.ipvt) == gold[mode]["ipvt"]
    np.testing.assert_allclose(q.Rmat(), np.array(gold[mode]["R"]), atol=tol)
    np.testing.assert_allclose(q.qtb, np.array(gold[mode]["qtb"]), atol=tol)
    np.testing.assert_allclose(q.solution(), np.array(gold["solution"]), atol=tol)


def test_qr_rejects_empty_and_short_systems(oracle):
    # QRSpec.scala:150-162: IllegalArgumentException
    with pytest.raises(ValueError):
        oracle.qr(np.empty((0, 4)), 3)
    with pytest.raises(ValueError):
        oracle.qr(np.array([[1.0, 2.0, 3.0, 0.0]]), 3)


CASES = [
    ("LINEAR", 4, [80.0, 0.5, 1.0, 2.0, 3.0], 1.0),
    ("LINEAR_NOINT", 3, [0.5, 1.0, 2.0], 1.0),
    ("EXPDECAY", 1, [2.0, -1.3], 0.1),
    ("GAUSS", 1, [2.0, 0.5, 0.05], 0.05),
    ("GAUSS_EXPBKG", 1, [2.0, 0.5, 0.05, 1.0, 3.0], 0.05),
    ("POLY", 1, [1.0, -2.0, 0.5, 3.0], 0.1),
]


@pytest.mark.parametrize("name,f,truth,sigma", CASES)
def test_analytic_matches_reference_finite_differences(oracle, name, f, truth, sigma):
    """FFNeuralNetworkTrainerSpec.scala:38-72 pattern: analytic gradient vs SimpleMSEFunction's FD, 1e-5."""
    fam = getattr(oracle, name)
    X, y = oracle.generate(fam, 2000, f, truth, noise_sigma=sigma)
    p = np.array(truth) * 1.1
    g = oracle.gradient(fam, X, y, p)
    g_fd = oracle.gradient_fd(fam, X, y, p)
    np.testing.assert_allclose(g, g_fd, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(g).max()))
    # Jacobian rows: FD (MSEFunction.scala:117-124) vs analytic, in absolute value like the spec (:120-156)
    Ja = oracle.jacobian_matrix(fam, X, y, p, fd=False)
    Jf = oracle.jacobian_matrix(fam, X, y, p, fd=True)
    np.testing.assert_allclose(Ja[:, -1], Jf[:, -1], rtol=0, atol=0)
    np.testing.assert_allclose(Ja[:, :-1], Jf[:, :-1], rtol=1e-4, atol=1e-5 * max(1.0, np.abs(Ja).max()))
    # exact Hessian-vector product vs the reference's finite difference of gradients
    d = np.cos(np.arange(p.size) + 1.0)
    hv = oracle.hess_vec(fam, X, y, p, d)
    hv_fd = oracle.hess_vec_fd(fam, X, y, p, d, eps=1e-6)
    np.testing.assert_allclose(hv, hv_fd, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(hv).max()))


def test_logistic_gradient_and_hessian(oracle):
    fam = oracle.LOGISTIC
    truth = [0.1, 1.5, 0.4]          # FFNeuralNetworkTrainerSpec.scala:189-220
    X, y = oracle.generate(fam, 5000, 2, truth)
    assert set(np.unique(y)) <= {0.0, 1.0}
    p = np.array([-0.5, 0.6, 0.6])
    g = oracle.gradient(fam, X, y, p)
    g_fd = oracle.gradient_fd(fam, X, y, p)
    np.testing.assert_allclose(g, g_fd, atol=1e-6)
    d = np.array([0.3, -0.2, 0.9])
    np.testing.assert_allclose(oracle.hess_vec(fam, X, y, p, d), oracle.hess_vec_fd(fam, X, y, p, d, eps=1e-6),
                               atol=1e-5)


def test_normal_equations_agree_with_the_reference_qr(oracle):
    """What LevenbergMarquardt needs from QR.apply (R, qtb, acNorms, bNorm) is a function of JtJ, Jtr, rtr."""
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, 3000, 1, truth, noise_sigma=0.05)
    p = truth * 1.1
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    rows = oracle.jacobian_matrix(fam, X, y, p)
    q = oracle.qr(rows, 5, pivoting=False)
    R = q.Rmat()
    np.testing.assert_allclose(R.T @ R, JtJ, rtol=1e-10, atol=1e-10 * np.abs(JtJ).max())
    np.testing.assert_allclose(R.T @ q.qtb, Jtr, rtol=1e-9, atol=1e-10 * np.abs(Jtr).max())
    np.testing.assert_allclose(q.acnorms, np.sqrt(np.diag(JtJ)), rtol=1e-12)
    assert q.bnorm == pytest.approx(np.sqrt(rtr), rel=1e-12)
    # loss = rtr / (2m), gradient = Jtr / m
    assert oracle.apply(fam, X, y, p) == pytest.approx(rtr / (2 * 3000), rel=1e-12)
    np.testing.assert_allclose(oracle.gradient(fam, X, y, p), Jtr / 3000, rtol=1e-11)


def test_partitioned_fold_matches_sequential_fold(oracle):
    """par[N] (Spark local[N] model) and seq (SeqDataSet) differ only by rounding."""
    fam = oracle.LINEAR
    truth = np.arange(9, dtype=float)
    X, y = oracle.generate(fam, 10007, 8, truth)
    p = truth + 0.25
    a = oracle.apply(fam, X, y, p, parts=1)
    for parts, threads in [(2, 1), (7, 3), (8, 8)]:
        b = oracle.apply(fam, X, y, p, parts=parts, threads=threads)
        assert b == pytest.approx(a, rel=1e-13)
    g1 = oracle.gradient(fam, X, y, p)
    g8 = oracle.gradient(fam, X, y, p, parts=8, threads=4)
    np.testing.assert_allclose(g1, g8, rtol=1e-12)


def test_line_scores_are_loss_and_directional_derivative(oracle):
    fam = oracle.EXPDECAY
    X, y = oracle.generate(fam, 500, 1, [2.0, 1.0], noise_sigma=0.1)
    p, d = np.array([4.0, 0.5]), np.array([-1.0, 0.3])
    phi, dphi = oracle.line(fam, X, y, p, d, [0.0, 0.5, 1.0])
    assert phi[0] == oracle.apply(fam, X, y, p)
    assert dphi[0] == pytest.approx(float(oracle.gradient(fam, X, y, p) @ d), rel=1e-14)
    h = 1e-6
    phi_h, _ = oracle.line(fam, X, y, p, d, [0.5 + h, 0.5 - h])
    assert (phi_h[0] - phi_h[1]) / (2 * h) == pytest.approx(dphi[1], rel=1e-5)


def test_faithful_lm_pass_model_runs(oracle):
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, 20000, 1, truth, noise_sigma=0.05)
    b1 = oracle.lm_reference_passes(fam, X, y, truth * 1.1, threads=1)
    b4 = oracle.lm_reference_passes(fam, X, y, truth * 1.1, threads=4)
    _, _, rtr = oracle.normal_eq(fam, X, y, truth * 1.1)
    assert b1 == pytest.approx(np.sqrt(rtr), rel=1e-12)
    assert b4 == pytest.approx(b1, rel=1e-12)


def test_mlp_restatement_reproduces_the_trainer_spec_properties(oracle):
    """FFNeuralNetworkTrainerSpec.scala:38-117: on Vector(2, 2, 1) the back-propagated gradient equals the finite
    difference gradient of a SimpleMSEFunction (loss r^2/2) for the MSE loss, and of the trainer's own apply for the
    cross-entropy loss, to 1e-5; the Jacobian rows times the residual give the gradient rows back (Neuron.scala:65-73)."""
    rng = np.random.default_rng(1)
    m, f, h = 400, 2, 2
    X = rng.standard_normal((m, f))
    w = rng.uniform(-1.0, 1.0, h * (f + 1) + h + 1)
    y = rng.standard_normal(m)
    g = oracle.gradient(oracle.MLP_MSE, X, y, w)
    fd = oracle.gradient_fd(oracle.MLP_MSE, X, y, w)          # FD of the trainer's apply = sum r^2 / m (not halved)
    np.testing.assert_allclose(g, 0.5 * fd, atol=1e-5)
    yc = (rng.random(m) < 0.5).astype(float)
    gc = oracle.gradient(oracle.MLP_CE, X, yc, w)
    np.testing.assert_allclose(gc, oracle.gradient_fd(oracle.MLP_CE, X, yc, w), atol=1e-5)
    for fam, yy in ((oracle.MLP_MSE, y), (oracle.MLP_CE, yc)):
        J = oracle.jacobian_matrix(fam, X, yy, w)
        np.testing.assert_allclose((J[:, :-1] * J[:, -1:]).sum(axis=0) / m, oracle.gradient(fam, X, yy, w), rtol=1e-12)
    # the network function itself: logistic inner neurons, linear output
    e = w[:6].reshape(2, 3)
    o = 1.0 / (1.0 + np.exp(-(e[:, 0] + X[0] @ e[:, 1:].T)))
    assert oracle.model(oracle.MLP_MSE, w, X[0]) == pytest.approx(w[6] + o @ w[7:], rel=1e-14)


def test_compensated_normal_equations_agree_with_the_fold_and_beat_it(oracle):
    """oracle.normal_eq_comp (exact products + Neumaier; the yardstick of the 1e9-row GPU test) against the restated left
    fold on a size where both run, and against an exactly representable case."""
    fam = oracle.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    X, y = oracle.generate(fam, 200_001, 1, truth, noise_sigma=0.05)
    p = truth * 1.1
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    for threads in (1, 3):
        hi, lo = oracle.normal_eq_comp(fam, X, y, p, threads=threads)
        ref = hi + lo
        np.testing.assert_allclose(ref[:25].reshape(5, 5), JtJ, rtol=1e-12, atol=1e-12 * np.abs(JtJ).max())
        np.testing.assert_allclose(ref[25:30], Jtr, rtol=1e-10, atol=1e-12 * np.abs(Jtr).max())
        assert ref[30] == pytest.approx(rtr, rel=1e-13)
    # linear family, integer data: every product and sum is exact, so hi is the exact answer and lo is zero
    Xi = np.arange(1.0, 2001.0).reshape(-1, 1)
    yi = 3.0 * Xi[:, 0] + 7.0
    hi, lo = oracle.normal_eq_comp(oracle.LINEAR, Xi, yi, np.array([1.0, 1.0]), threads=2)
    r = yi - (1.0 + Xi[:, 0])
    assert np.all(lo == 0.0)
    assert hi[-1] == float(np.sum(r.astype(object) ** 2))
    assert hi[0] == 2000.0 and hi[1] == float(np.sum(Xi)) and hi[3] == float(np.sum(Xi[:, 0].astype(object) ** 2))


def test_network_restatement_with_vector_targets(oracle):
    """The general FFNeuralNetwork restatement (oracle.net_*): what FFNeuralNetworkTrainerSpec.scala:38-117 asserts of the
    reference — back-propagated gradient = finite-difference gradient to 1e-5 — for the multi-class shapes of :222-260 and
    :353-376, with and without weight decay; equality with the one-output restatement; the penalty rows; and the refusal
    of mean squared error with several outputs (`???` in the reference)."""
    rng = np.random.default_rng(1)
    for sizes in ((4, 5, 3), (2, 3), (3, 2, 2, 2)):
        n = oracle.net_nweights(sizes)
        X = rng.standard_normal((60, sizes[0]))
        Y = np.eye(sizes[-1])[rng.integers(0, sizes[-1], 60)]
        w = rng.uniform(-0.5, 0.5, n)
        for decay in (0.0, 0.01):
            l, g = oracle.net_loss_grad(sizes, True, X, Y, w, decay=decay)
            eps = 1e-7
            gfd = np.array([(oracle.net_loss_grad(sizes, True, X, Y, w + eps * np.eye(n)[i], decay=decay, want_grad=False)[0] - l) / eps
                            for i in range(n)])
            np.testing.assert_allclose(g, gfd, rtol=0, atol=1e-5)
    w2 = rng.uniform(-0.5, 0.5, 9)
    X2 = rng.standard_normal((40, 2)); y2 = rng.integers(0, 2, 40).astype(float)
    for fam, ce in ((oracle.MLP_CE, True), (oracle.MLP_MSE, False)):
        l_new, g_new = oracle.net_loss_grad([2, 2, 1], ce, X2, y2.reshape(-1, 1), w2)
        assert l_new == oracle.apply(fam, X2, y2, w2)
        np.testing.assert_array_equal(g_new, oracle.gradient(fam, X2, y2, w2))
        np.testing.assert_array_equal(oracle.net_jacobian_matrix([2, 2, 1], ce, X2, y2.reshape(-1, 1), w2),
                                      oracle.jacobian_matrix(fam, X2, y2, w2))
    rows = oracle.net_jacobian_matrix([2, 2, 1], True, X2, y2.reshape(-1, 1), w2, decay=0.04)
    assert rows.shape == (49, 10)
    np.testing.assert_array_equal(rows[40:, :9], 0.2 * np.eye(9)); assert np.all(rows[40:, 9] == 0.0)
    with pytest.raises(NotImplementedError):
        oracle.net_loss_grad([2, 2, 3], False, X2, np.zeros((40, 3)), rng.uniform(-1, 1, 15))

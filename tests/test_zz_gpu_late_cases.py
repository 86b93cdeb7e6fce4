"""GPU parity cases written after the round's GPU minutes were spent — sorted last on purpose: their first run on a B200 is
the driver's; their logic was dry-run against an oracle-backed stand-in data set.  What they reach that no earlier case does:
  * K3 (`so_eval_line`) of the single-input families at k = 1, 6, 7, 8 (one kernel instantiation per k = 1..8, OpLine<Fam, KB> in
    so_scalar.cu; the other cases use k = 2..5) and the ceil(k/8)-pass split; k = 6, 7, 8 for the linear-predictor families;
  * K2 at the tile counts T = ceil(f/8) = 10, 11, 12 of k_gram_ph (f = 73..96) and the logistic family at T = 13, the largest
    shared-memory footprint of that kernel (229 184 of 232 448 bytes);
  * K1 / K3 / K4 on odd rows with 8 lanes x 4 units (f = 41..63 odd), which only smoke() reached;
  * the weight-decay terms of GpuMSEFunction on the K3 (batched line-search) route and in dirHessian — both were missing
    until these cases were written (host/gpu.hpp)."""
import numpy as np
import pytest

from parity import assert_gram, assert_loss, assert_vec

pytestmark = pytest.mark.gpu

CASES = [
    ("EXPDECAY", [2.0, -1.3], 0.1),
    ("GAUSS", [2.0, 0.5, 0.05], 0.05),
    ("GAUSS_EXPBKG", [2.0, 0.5, 0.05, 1.0, 3.0], 0.05),   # BASELINE config 3
    ("POLY", [1.0, -2.0, 0.5, 3.0], 0.1),
]
SCHEDULE = [0.0, 1.0, 2.0, 4.0, 0.37, 0.5, 3.0, 1.5, 0.1, 2.5, 0.75, 3.5, 0.25]


@pytest.mark.parametrize("name,truth,sigma", CASES)
@pytest.mark.parametrize("k", [1, 6, 7, 8, 9, 13])
def test_line_scores_at_every_batch_size(native, oracle, gpu_ctx, name, truth, sigma, k):
    fam = getattr(oracle, name)
    m = 100_003
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=sigma)
    ds = native.Dataset(gpu_ctx, m, 1).append(X, y)
    p = np.array(truth) * 1.05 + 0.01
    _, grad = ds.loss_grad(fam, p)
    d = -0.02 * np.abs(p) * np.sign(grad)                  # a descent direction that moves every parameter by <= 8 % at alpha = 4
    alphas = SCHEDULE[:k]
    phi, dphi = ds.line(fam, p, d, alphas)
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
    assert_vec(phi, phi_o, what=f"phi k={k}")
    assert_vec(dphi, dphi_o, rtol=1e-11, what=f"dphi k={k}")
    # the batch is the same function of each step size as a pass of its own
    for j in sorted({0, k - 1}):
        phi1, dphi1 = ds.line(fam, p, d, [alphas[j]])
        assert_vec([phi[j]], phi1, what=f"phi[{j}] vs single")
        assert_vec([dphi[j]], dphi1, rtol=1e-11, what=f"dphi[{j}] vs single")
    ds.free()


@pytest.mark.parametrize("name,f", [("LINEAR", 3), ("LOGISTIC", 7), ("LINEAR", 12), ("LINEAR_NOINT", 32), ("LOGISTIC", 33)])
@pytest.mark.parametrize("k", [6, 7, 8])
def test_line_scores_of_the_linear_predictor_families_at_6_to_8_step_sizes(native, oracle, gpu_ctx, name, f, k):
    fam = getattr(oracle, name)
    n = f + (0 if name == "LINEAR_NOINT" else 1)
    truth = np.array([(((j % 7) - 3) / 4.0) for j in range(n)])
    m = 20_001
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    d = 0.02 * np.cos(np.arange(n))
    alphas = SCHEDULE[:k]
    phi, dphi = ds.line(fam, p, d, alphas)
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
    assert_vec(phi, phi_o, what=f"phi k={k}")
    assert_vec(dphi, dphi_o, rtol=1e-11, what=f"dphi k={k}")
    ds.free()


@pytest.mark.parametrize("name,f", [("LINEAR_NOINT", 80), ("LINEAR", 88), ("LINEAR_NOINT", 96), ("LINEAR", 95), ("LOGISTIC", 96),
                                    ("LOGISTIC", 104), ("LOGISTIC", 79)])
@pytest.mark.parametrize("m", [6_001, 5])
def test_normal_equations_at_the_remaining_tile_counts(native, oracle, gpu_ctx, name, f, m):
    fam = getattr(oracle, name)
    n = f + (0 if name == "LINEAR_NOINT" else 1)
    truth = np.array([(((j % 7) - 3) / 4.0) for j in range(n)])
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    if name == "LOGISTIC":       # (x * error) / residual with error == residual in the reference: 1-ulp noise per entry
        np.testing.assert_allclose(JtJ, JtJ_o, rtol=1e-11, atol=1e-11 * np.abs(JtJ_o).max())
        assert np.array_equal(JtJ, JtJ.T)
    else:
        assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr, Jtr_o, what="Jtr")
    assert_loss(rtr, rtr_o)
    ds.free()


@pytest.mark.parametrize("name,f", [("LINEAR", 47), ("LOGISTIC", 63), ("LINEAR_NOINT", 41)])
def test_odd_rows_on_the_8_lane_mapping(native, oracle, gpu_ctx, name, f):
    fam = getattr(oracle, name)
    n = f + (0 if name == "LINEAR_NOINT" else 1)
    truth = np.array([(((j % 7) - 3) / 4.0) for j in range(n)])
    m = 20_001
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p))
    d = 0.02 * np.cos(np.arange(n))
    for alphas in ([0.5], [0.0, 1.0, 2.0], SCHEDULE[:8]):
        phi, dphi = ds.line(fam, p, d, alphas)
        phi_o, dphi_o = oracle.line(fam, X, y, p, d, alphas)
        assert_vec(phi, phi_o, what=f"phi k={len(alphas)}")
        assert_vec(dphi, dphi_o, rtol=1e-11, what=f"dphi k={len(alphas)}")
    assert_vec(ds.hess_vec(fam, p, d), oracle.hess_vec(fam, X, y, p, d), rtol=1e-11, what="Hd")
    ds.free()


@pytest.mark.parametrize("decay", [0.0, 0.05])
def test_batched_line_scores_carry_the_weight_decay(native, oracle, gpu_ctx, decay):
    """GpuMSEFunction scores the later trial points of a line search with K3 passes (score_ray, host/gpu.hpp).  The device
    returns the data term; the trainer's weight decay (Trainer.scala:59-69: loss |w|^2/2 decay / m, gradient w decay / m)
    is added on the host — on the K1 route since round 2, on the K3 route only since this test was written: phi and phi' of a
    point on the ray must be the trainer's loss and directional derivative there, decay included."""
    from scalaopt_b200 import host
    sizes = (4, 5, 3)
    n = oracle.net_nweights(sizes)
    rng = np.random.default_rng(7)
    lab = np.repeat(np.arange(3), 40)
    X = rng.standard_normal((120, 4)) + lab[:, None]
    Y = np.eye(3)[lab]
    w0 = rng.uniform(-0.5, 0.5, n)
    ds = native.Dataset(gpu_ctx, 120, 4, ny=3).append(X, Y)
    f = host.Objective.gpu(gpu_ctx, ds, native.MLP_CE, n, decay=decay)
    ref = lambda w: oracle.net_loss_grad(sizes, True, X, Y, w, decay=decay)
    l0, g0 = ref(w0)
    d = -g0
    assert f(w0) == pytest.approx(l0, rel=1e-12)
    assert f.dirder(w0, d) == pytest.approx(float(g0 @ d), rel=1e-11)            # K1 route; tells the objective the ray
    for alpha in (0.5, 0.125):
        w = w0 + alpha * d
        lr, gr = ref(w)
        assert f(w) == pytest.approx(lr, rel=1e-12)                              # K3 route: first request at this point
        assert f.dirder(w, d) == pytest.approx(float(gr @ d), rel=1e-10)
    assert f.counters()["k3_passes"] >= 1
    f.free(); ds.free()


def test_hessian_vector_product_carries_the_weight_decay(native, oracle, gpu_ctx):
    """FFNeuralNetworkTrainer.dirHessian = (gradient(w + eps d) - gradient(w)) / eps with gradients that contain w decay / m
    (Trainer.scala:64-69,76-80).  The device composes the data part; decay d / m is added by the host objective (it was
    missing until this test was written).  Against the same difference of the restated trainer's gradients; the decay is
    large here so that its term (1.3 % of max |H d| on these data) stands clear of the differencing noise (1e-8 .. 1e-6 relative)."""
    from scalaopt_b200 import host
    sizes, decay, eps = (4, 5, 3), 0.5, 1.0e-8
    n = oracle.net_nweights(sizes)
    rng = np.random.default_rng(11)
    lab = np.repeat(np.arange(3), 40)
    X = rng.standard_normal((120, 4)) + lab[:, None]
    Y = np.eye(3)[lab]
    w = rng.uniform(-0.5, 0.5, n)
    d = rng.standard_normal(n)
    ds = native.Dataset(gpu_ctx, 120, 4, ny=3).append(X, Y)
    grad = lambda v, dec: oracle.net_loss_grad(sizes, True, X, Y, v, decay=dec)[1]
    for dec in (0.0, decay):
        f = host.Objective.gpu(gpu_ctx, ds, native.MLP_CE, n, decay=dec)
        Hd = f.dirHessian(w, d)
        ref = (grad(w + d * eps, dec) - grad(w, dec)) / eps
        scale = np.abs(ref).max()
        assert np.abs(Hd - ref).max() <= 1e-4 * scale, (dec, np.abs(Hd - ref).max() / scale)
        f.free()
    # the decay term itself is well above that tolerance: without it the second comparison cannot pass
    assert decay * np.abs(d).max() / 120 > 1e-3 * np.abs((grad(w + d * eps, decay) - grad(w, decay)) / eps).max()
    ds.free()

"""Full-size checks at the BASELINE.json configurations, where the CPU oracle cannot finish: size-independent
properties of the evaluation (additivity over row ranges = the aggregate's combop, cross-kernel identities,
finite-difference consistency, optimizer recovery of the planted parameters) plus oracle parity on sampled
row windows of the very same device-resident bits."""
import numpy as np
import pytest

from parity import assert_gram, assert_loss, assert_vec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from scalaopt_b200 import host as h
    h.load()
    return h


def test_config3_gaussian_peak_1e9_rows(native, oracle, host, gpu_ctx):
    fam = native.GAUSS_EXPBKG
    truth = np.array([2.0, 0.5, 0.05, 1.0, 3.0])
    m = 1_000_000_000
    ds = native.Dataset(gpu_ctx, m, 1).generate(fam, truth, seed=12345, noise_sigma=0.05)
    assert ds.size() == m
    p = truth * 1.1
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    assert np.array_equal(JtJ, JtJ.T)
    np.linalg.cholesky(JtJ)                                           # positive definite
    # additivity over ragged row ranges (odd boundaries: unaligned slices)
    cuts = [0, 1, 333_333_333, 777_777_778, m]
    parts = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        v = ds.slice(b, e)
        parts.append(v.normal_eq(fam, p))
        v.free()
    assert_gram(sum(x[0] for x in parts), JtJ)
    assert_vec(sum(x[1] for x in parts), Jtr, what="Jtr additivity")
    assert_loss(sum(x[2] for x in parts), rtr)
    # K1 / K2 / K3 / K4 identities
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, rtr / (2.0 * m))                                 # MSEFunction.apply = sum r^2/2 / m
    assert_vec(grad, Jtr / m, what="grad = Jtr/m")
    d = -grad
    phi, dphi = ds.line(fam, p, d, [0.0, 1.0, 2.0])
    assert_loss(phi[0], loss)
    assert abs(dphi[0] - grad @ d) <= 1e-11 * abs(grad @ d)
    l1, g1 = ds.loss_grad(fam, p + d)
    assert_loss(phi[1], l1)
    assert abs(dphi[1] - g1 @ d) <= 1e-10 * max(abs(g1 @ d), abs(grad @ d))
    Hd = ds.hess_vec(fam, p, d)
    eps = 1e-6
    _, g_eps = ds.loss_grad(fam, p + eps * d)
    np.testing.assert_allclose(Hd, (g_eps - grad) / eps, rtol=1e-4, atol=1e-6 * np.abs(Hd).max())
    # oracle parity on sampled windows of the same bits (head, unaligned middle, tail)
    for b, rows in ((0, 1_000_000), (500_000_001, 1_000_001), (m - 999_999, 999_999)):
        X, y = ds.read(b, rows)
        v = ds.slice(b, b + rows)
        JtJ_w, Jtr_w, rtr_w = v.normal_eq(fam, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p, parts=8, threads=8)
        assert_gram(JtJ_w, JtJ_o); assert_vec(Jtr_w, Jtr_o, what="Jtr window"); assert_loss(rtr_w, rtr_o)
        v.free()
    # summation error at m = 1e9 ("parity at scale"): all 1e9 rows read back in windows and summed
    # on the host with exact products and Neumaier compensation (oracle.normal_eq_comp; the windows' hi/lo parts are
    # added exactly with math.fsum) — the GPU's block/warp tree is held to the north_star tolerances against THAT
    import math
    import os
    n, Q = 5, 5 * 5 + 5 + 1
    pieces = [[] for _ in range(Q)]
    W = 1 << 26
    threads = max(1, len(os.sched_getaffinity(0)))
    for b in range(0, m, W):
        X, y = ds.read(b, min(W, m - b))
        hi, lo = oracle.normal_eq_comp(fam, X, y, p, threads=threads)
        for q in range(Q):
            pieces[q] += [float(hi[q]), float(lo[q])]
    ref = np.array([math.fsum(v) for v in pieces])
    assert_gram(JtJ, ref[: n * n].reshape(n, n))
    assert_vec(Jtr, ref[n * n: n * n + n], rtol=1e-11, what="Jtr vs compensated sum of 1e9 rows")
    assert_loss(rtr, ref[-1])
    assert_loss(loss, ref[-1] / (2.0 * m))
    # run-to-run reproducibility at full size, every kernel that streams through the shared-memory ring (a refill that
    # races with the loads of the stage it overwrites shows up here and nowhere at small sizes: ~1600 chunks per block)
    def bits(*arrays):
        return np.concatenate([np.ravel(np.asarray(a, dtype=np.float64)) for a in arrays]).view(np.uint64)
    for what, call in (("K2", lambda: bits(*ds.normal_eq(fam, p))), ("K1", lambda: bits(*ds.loss_grad(fam, p))),
                       ("TSQR", lambda: bits(ds.qr(fam, p))), ("K3", lambda: bits(*ds.line(fam, p, d, [0.5, 1.0]))),
                       ("K4", lambda: bits(ds.hess_vec(fam, p, d)))):
        first = call()
        for _ in range(5):
            assert np.array_equal(call(), first), f"{what} is not reproducible at 1e9 rows"
    # Levenberg-Marquardt on all 1e9 observations recovers the planted parameters
    f = host.Objective.gpu(gpu_ctx, ds, fam, 5)
    x, res = host.minimize(host.LEVENBERG_MARQUARDT, f, truth * 1.1)
    np.testing.assert_allclose(x, truth, rtol=2e-4)
    c = f.counters()
    assert c["k1_passes"] + c["spec_passes"] == res.inner_iterations and c["k2_passes"] + c["spec_hits"] == res.iterations + c["spec_passes"]
    f.free(); ds.free()


def test_config2_logistic_1e8_x_32(native, oracle, host, gpu_ctx):
    fam = native.LOGISTIC
    f_ = 32
    truth = np.array([(((j % 7) - 3) / 4.0) for j in range(f_ + 1)])
    m = 100_000_000
    ds = native.Dataset(gpu_ctx, m, f_).generate(fam, truth, seed=12345)
    p = truth * 0.5
    loss, grad = ds.loss_grad(fam, p)
    halves = []
    for b, e in ((0, 33_333_333), (33_333_333, m)):
        v = ds.slice(b, e)
        l, g = v.loss_grad(fam, p)
        halves.append((l * (e - b), g * (e - b)))
        v.free()
    assert_loss(sum(h[0] for h in halves) / m, loss)
    assert_vec(sum(h[1] for h in halves) / m, grad, what="grad additivity")
    d = -grad
    alphas = [0.0, 1.0, 2.0, 4.0, 0.5]
    phi, dphi = ds.line(fam, p, d, alphas)
    assert_loss(phi[0], loss)
    l2, g2 = ds.loss_grad(fam, p + 2.0 * d)
    assert_loss(phi[2], l2)
    assert abs(dphi[2] - g2 @ d) <= 1e-10 * abs(grad @ d)
    Hd = ds.hess_vec(fam, p, d)
    _, g_eps = ds.loss_grad(fam, p + 1e-6 * d)
    np.testing.assert_allclose(Hd, (g_eps - grad) / 1e-6, rtol=1e-4, atol=1e-6 * np.abs(Hd).max())
    X, y = ds.read(50_000_000, 200_000)
    v = ds.slice(50_000_000, 50_200_000)
    lw, gw = v.loss_grad(fam, p)
    assert_loss(lw, oracle.apply(fam, X, y, p)); assert_vec(gw, oracle.gradient(fam, X, y, p))
    v.free()
    fobj = host.Objective.gpu(gpu_ctx, ds, fam, f_ + 1)
    x, res = host.minimize(host.BFGS, fobj, np.zeros(f_ + 1))          # BFGSConfig defaults, start 0 (cfg2)
    assert np.abs(x - truth).max() < 5e-3
    c = fobj.counters()
    assert c["k1_passes"] + c["k3_passes"] <= res.iterations + 2       # one pass per distinct point
    fobj.free(); ds.free()


def test_config4_linear_1e7_x_256(native, oracle, host, gpu_ctx):
    fam = native.LINEAR_NOINT
    f_ = 256
    truth = np.arange(1, f_ + 1) / f_
    m = 10_000_000
    ds = native.Dataset(gpu_ctx, m, f_).generate(fam, truth, seed=12345)
    p = truth * 0.9
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    assert np.array_equal(JtJ, JtJ.T)
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, rtr / (2.0 * m)); assert_vec(grad, Jtr / m, what="grad = Jtr/m")
    d = np.cos(np.arange(f_) + 1.0)
    Hd = ds.hess_vec(fam, p, d)
    assert_vec(Hd, JtJ @ d / m, rtol=1e-11, what="Hd = JtJ d / m (linear least squares)")
    X, y = ds.read(5_000_000, 20_000)
    v = ds.slice(5_000_000, 5_020_000)
    JtJ_w, Jtr_w, rtr_w = v.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p, parts=8, threads=8)
    assert_gram(JtJ_w, JtJ_o); assert_vec(Jtr_w, Jtr_o, what="Jtr window")
    v.free()
    fobj = host.Objective.gpu(gpu_ctx, ds, fam, f_)
    x, res = host.minimize(host.NEWTON_CG, fobj, np.zeros(f_))
    assert np.abs(x - truth).max() < 1e-2
    # the normal equations solved on the host agree with the Newton-CG minimiser
    x_ne = p + np.linalg.solve(JtJ, -Jtr)
    np.testing.assert_allclose(x, x_ne, atol=1e-4)
    fobj.free(); ds.free()

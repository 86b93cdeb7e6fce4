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
    # 5 units per lane (17-20, 33-40, 65-80, 129-160 units per row), even and odd f; thread-per-row up to n = 12
    ("LINEAR", 35), ("LINEAR_NOINT", 40), ("LINEAR_NOINT", 72), ("LOGISTIC", 79), ("LINEAR", 150), ("LINEAR_NOINT", 157),
    ("LINEAR_NOINT", 300), ("LINEAR", 319), ("LINEAR", 9), ("LINEAR_NOINT", 12), ("LOGISTIC", 11),
    # 2 lanes per row (5-8 units, register prefetch): what is left of it above the thread-per-row range, n = 13 .. 17
    ("LINEAR", 12), ("LINEAR_NOINT", 13), ("LOGISTIC", 14), ("LINEAR", 15), ("LINEAR_NOINT", 16),
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


@pytest.mark.parametrize("name,f", [("LINEAR", 9), ("LINEAR_NOINT", 13), ("LINEAR_NOINT", 31), ("LOGISTIC", 33), ("LINEAR", 127), ("LINEAR_NOINT", 255)])
def test_odd_row_lengths_on_views_at_any_offset(native, oracle, gpu_ctx, name, f):
    """f odd: the kernels stream a row as (f + 1)/2 aligned 16-byte units, i.e. one element of a NEIGHBOUR row travels with it
    (so_wide.cu, ODD).  Views that start on odd rows (misaligned base), end on either parity, and are followed by a row of
    NaNs (what lies behind the last row of a shard must never reach a sum) against the oracle on the same rows."""
    fam = getattr(oracle, name)
    truth = _truth(name, f)
    m = 2_004
    X, y = oracle.generate(fam, m, f, truth)
    Xp = X.copy()
    Xp[m - 1, :] = np.nan                                # the row behind ...
    Xp[0, :] = np.nan                                    # ... and the row before some of the views below
    ds = native.Dataset(gpu_ctx, m, f).append(Xp, y)
    p = truth * 0.9 + 0.05
    for lo, hi in [(1, m - 1), (1, m - 2), (2, m - 1), (2, m - 2), (7, 8), (8, 9), (3, 70)]:
        v = ds.slice(lo, hi)
        Xv, yv = X[lo:hi], y[lo:hi]
        loss, grad = v.loss_grad(fam, p)
        assert_loss(loss, oracle.apply(fam, Xv, yv, p))
        assert_vec(grad, oracle.gradient(fam, Xv, yv, p), what=f"grad [{lo},{hi})")
        d = np.cos(np.arange(p.size)) / p.size
        assert_vec(v.hess_vec(fam, p, d), oracle.hess_vec(fam, Xv, yv, p, d), rtol=1e-11, what=f"Hd [{lo},{hi})")
        phi, dphi = v.line(fam, p, d, [0.0, 0.5, 1.0])
        phi_o, dphi_o = oracle.line(fam, Xv, yv, p, d, [0.0, 0.5, 1.0])
        assert_vec(phi, phi_o, what=f"phi [{lo},{hi})")
        assert_vec(dphi, dphi_o, rtol=1e-11, what=f"dphi [{lo},{hi})")
        JtJ, Jtr, rtr = v.normal_eq(fam, p)
        JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, Xv, yv, p)
        np.testing.assert_allclose(JtJ, JtJ_o, rtol=1e-11, atol=1e-11 * np.abs(JtJ_o).max())
        assert_vec(Jtr, Jtr_o, what=f"Jtr [{lo},{hi})")
        v.free()
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


@pytest.mark.parametrize("name,f", [("LINEAR", 3), ("LINEAR", 20), ("LINEAR_NOINT", 7), ("LINEAR_NOINT", 33), ("LINEAR", 64),
                                    ("LOGISTIC", 4), ("LOGISTIC", 17), ("LOGISTIC", 40)])
@pytest.mark.parametrize("seed", range(3))
def test_linear_predictor_families_at_random_scales(native, oracle, gpu_ctx, name, f, seed):
    """columns scaled over six decades, random parameters: K1-K4, K2 and (n <= 8) the TSQR against the oracle"""
    fam = getattr(oracle, name)
    rng = np.random.default_rng(3000 + 17 * f + seed)
    m = 2503
    colscale = 10.0 ** rng.uniform(-3.0, 3.0, size=f)
    X = rng.standard_normal((m, f)) * colscale
    icpt = 0 if name == "LINEAR_NOINT" else 1
    w = rng.standard_normal(f) / colscale / np.sqrt(f)
    z = X @ w + (0.3 if icpt else 0.0)
    if name == "LOGISTIC":
        y = (rng.random(m) < 1.0 / (1.0 + np.exp(-z))).astype(float)
    else:
        y = z + 0.1 * rng.standard_normal(m)
    p = np.concatenate([[0.1] if icpt else [], w * rng.uniform(0.5, 1.5, size=f)])
    n = p.size
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    loss, grad = ds.loss_grad(fam, p)
    assert_loss(loss, oracle.apply(fam, X, y, p))
    assert_vec(grad, oracle.gradient(fam, X, y, p), rtol=1e-11)
    d = np.concatenate([[0.01] if icpt else [], 0.05 * rng.standard_normal(f) / colscale / np.sqrt(f)])   # |x . d| ~ 0.05 per row
    phi, dphi = ds.line(fam, p, d, [0.0, 0.5, 1.0, 2.0])
    phi_o, dphi_o = oracle.line(fam, X, y, p, d, [0.0, 0.5, 1.0, 2.0])
    assert_vec(phi, phi_o, what="phi"); assert_vec(dphi, dphi_o, rtol=1e-10, what="dphi")
    assert_vec(ds.hess_vec(fam, p, d), oracle.hess_vec(fam, X, y, p, d), rtol=1e-10, what="Hd")
    JtJ, Jtr, rtr = ds.normal_eq(fam, p)
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, X, y, p)
    if name == "LOGISTIC":
        dg = np.sqrt(np.diag(JtJ_o)); np.testing.assert_allclose(JtJ / np.outer(dg, dg), JtJ_o / np.outer(dg, dg), rtol=0, atol=1e-11)
    else:
        assert_gram(JtJ, JtJ_o)
    assert_vec(Jtr / np.sqrt(np.diag(JtJ_o)), Jtr_o / np.sqrt(np.diag(JtJ_o)), rtol=1e-10, what="Jtr (column-scaled)")
    assert_loss(rtr, rtr_o)
    if n <= 8:
        R = ds.qr(fam, p)
        G = np.block([[JtJ_o, Jtr_o[:, None]], [Jtr_o[None, :], np.array([[rtr_o]])]])
        sc = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
        assert np.max(np.abs(R.T @ R - G) / sc) < 1e-10
    ds.free()


@pytest.mark.parametrize("f", [2, 7, 32])
def test_logistic_saturated_predictors_match_a_50_digit_reference(native, oracle, gpu_ctx, f):
    """Where the oracle stops (it forms o = 1/(1+exp(-z)) like Neuron.scala and overflows its own log for |z| >~ 36,
    so the k > 4 line-search cases above are skipped for this family): loss, gradient, phi/dphi at every step size and
    the Hessian-vector product against 50-digit mpmath arithmetic on the same rows, predictors out to |z| ~ 300."""
    import mpmath as mp
    mp.mp.dps = 50
    fam = oracle.LOGISTIC
    truth = _truth("LOGISTIC", f)
    m = 1500
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 30.0 + 0.5                                 # |z| up to ~100 at alpha = 0, several hundred along the ray
    d = np.cos(np.arange(f + 1) + 0.3) * 2.0
    alphas = [0.0, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0, 0.37, 3.0]

    def reference(w):
        w = [mp.mpf(float(v)) for v in w]
        loss = mp.mpf(0)
        grad = [mp.mpf(0)] * (f + 1)
        zs = []
        for i in range(m):
            xt = [mp.mpf(1)] + [mp.mpf(float(v)) for v in X[i]]
            z = sum(a * b for a, b in zip(w, xt))
            zs.append(z)
            yi = mp.mpf(float(y[i]))
            # CE(o, y) with o = 1/(1+exp(-z)): -y log o - (1-y) log(1-o) = log(1+exp(-z)) + (1-y) z
            loss += mp.log1p(mp.exp(-z)) + (1 - yi) * z
            o = 1 / (1 + mp.exp(-z))
            grad = [g + (o - yi) * xv for g, xv in zip(grad, xt)]
        return loss / m, [g / m for g in grad], zs

    loss, grad = ds.loss_grad(fam, p)
    l_ref, g_ref, zs = reference(p)
    assert max(abs(float(z)) for z in zs) > 36.0           # the regime the oracle cannot do
    assert abs(loss - float(l_ref)) <= 1e-12 * abs(float(l_ref))
    g_ref_f = np.array([float(g) for g in g_ref])
    assert_vec(grad, g_ref_f, rtol=1e-12, what="gradient vs mpmath")
    phi, dphi = ds.line(fam, p, d, alphas)
    for j, a in enumerate(alphas):
        l_a, g_a, zs_a = reference(p + a * d)
        dphi_ref = float(sum(g * mp.mpf(float(dv)) for g, dv in zip(g_a, d)))
        assert abs(phi[j] - float(l_a)) <= 1e-12 * abs(float(l_a)), (a, phi[j], float(l_a))
        assert abs(dphi[j] - dphi_ref) <= 1e-11 * max(abs(dphi_ref), 1e-300), (a, dphi[j], dphi_ref)
    # H d = (1/m) sum o (1 - o) (x~ . d) x~
    Hd = ds.hess_vec(fam, p, d)
    H_ref = [mp.mpf(0)] * (f + 1)
    for i in range(m):
        xt = [mp.mpf(1)] + [mp.mpf(float(v)) for v in X[i]]
        o = 1 / (1 + mp.exp(-zs[i]))
        xd = sum(a * mp.mpf(float(b)) for a, b in zip(xt, d))
        H_ref = [h + o * (1 - o) * xd * xv for h, xv in zip(H_ref, xt)]
    assert_vec(Hd, np.array([float(h / m) for h in H_ref]), rtol=1e-11, what="Hd vs mpmath")
    ds.free()


@pytest.mark.parametrize("name,f", [("LINEAR", 31), ("LOGISTIC", 33), ("LINEAR_NOINT", 9), ("LINEAR", 127), ("LINEAR_NOINT", 511)])
@pytest.mark.parametrize("m", [20_000, 20_001, 1, 2])
def test_odd_row_lengths_row_pair_path(native, oracle, gpu_ctx, name, f, m):
    """odd f streams row PAIRS with 128-bit loads (k_wide_pair): even and odd row counts, a single row, and views that
    start on an odd row (8-byte aligned only: the 64-bit mapping takes those) must all give the oracle's sums"""
    fam = getattr(oracle, name)
    truth = _truth(name, f)
    if f > 200 and m > 1000:
        m = 3_000 + (m & 1)
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    d = np.cos(np.arange(p.size)) * 0.1
    views = [(0, m)] + ([(1, m), (2, m - 1)] if m > 4 else [])
    for b, e in views:
        v = ds.slice(b, e)
        loss, grad = v.loss_grad(fam, p)
        assert_loss(loss, oracle.apply(fam, X[b:e], y[b:e], p))
        assert_vec(grad, oracle.gradient(fam, X[b:e], y[b:e], p))
        assert_vec(v.hess_vec(fam, p, d), oracle.hess_vec(fam, X[b:e], y[b:e], p, d), rtol=1e-11, what="Hd")
        v.free()
    ds.free()

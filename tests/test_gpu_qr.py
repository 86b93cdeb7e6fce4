"""On-device Householder TSQR of [J | r] (so_eval_qr) against the oracle's restatement of QR.apply
(linalg/QR.scala:131-250) on the materialised Jacobian/residual rows, against LAPACK, and — on an ill-conditioned
Jacobian — against the normal-equations route it replaces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("EXPDECAY", [2.0, -1.3], 0.1),
    ("GAUSS", [2.0, 0.5, 0.05], 0.05),
    ("GAUSS_EXPBKG", [2.0, 0.5, 0.05, 1.0, 3.0], 0.05),   # BASELINE config 3
    ("POLY", [1.0], 0.1),
    ("POLY", [1.0, -2.0, 0.5, 3.0], 0.1),
    ("POLY", [1.0, -2.0, 0.5, 3.0, -1.0, 0.25, 2.0, -0.5], 0.1),
]


def _normalise(R):
    """row signs of an R factor are arbitrary: make the diagonal non-negative"""
    s = np.where(np.diag(R) < 0, -1.0, 1.0)
    return R * s[:, None]


@pytest.mark.parametrize("name,truth,sigma", CASES)
@pytest.mark.parametrize("m", [100_003, 7, 2])
def test_tsqr_matches_reference_qr_and_lapack(native, oracle, gpu_ctx, name, truth, sigma, m):
    fam = getattr(oracle, name)
    X, y = oracle.generate(fam, m, 1, truth, noise_sigma=sigma)
    ds = native.Dataset(gpu_ctx, m, 1).append(X, y)
    p = np.array(truth) * 1.1 + 0.01
    n = p.size
    R = ds.qr(fam, p)
    assert gpu_ctx.stats().launches == 1                      # one pass, against 2n+1 in QR.apply
    assert R.shape == (n + 1, n + 1) and np.all(np.tril(R, -1) == 0.0)
    # R^T R is the normal-equations system of K2 (same conventions and scaling)
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    G = np.block([[JtJ, Jtr[:, None]], [Jtr[None, :], np.array([[rtr]])]])
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    assert np.max(np.abs(R.T @ R - G) / scale) < 1e-11
    if m <= n:
        return                                               # fewer rows than columns: R is not unique
    # the reference's own algorithm on the materialised rows (analytic Jacobian): R, Q^T b up to row signs
    ab = oracle.jacobian_matrix(fam, X, y, p)
    q = oracle.qr(ab, n, pivoting=False)
    Rn = _normalise(R)
    Rref = q.Rmat()
    sref = np.where(np.diag(Rref) < 0, -1.0, 1.0)
    colnorm = np.sqrt(np.diag(JtJ))
    np.testing.assert_allclose(Rn[:n, :n] / colnorm, (Rref * sref[:, None]) / colnorm, rtol=0, atol=1e-9)
    np.testing.assert_allclose(Rn[:n, n], q.qtb[:n] * sref, rtol=0, atol=1e-9 * np.sqrt(rtr))
    # and LAPACK's Householder QR of the same rows
    Rl = _normalise(np.linalg.qr(ab, mode="r"))
    np.testing.assert_allclose(Rn / np.append(colnorm, np.sqrt(rtr)), Rl / np.append(colnorm, np.sqrt(rtr)), rtol=0, atol=1e-9)
    ds.free()


WIDE_CASES = [("LINEAR", 1), ("LINEAR", 3), ("LINEAR", 7), ("LINEAR_NOINT", 2), ("LINEAR_NOINT", 4), ("LINEAR_NOINT", 8),
              ("LOGISTIC", 2), ("LOGISTIC", 4), ("LOGISTIC", 7)]


@pytest.mark.parametrize("name,f", WIDE_CASES)
@pytest.mark.parametrize("m", [50_001, 11])
def test_tsqr_of_the_narrow_linear_predictor_families(native, oracle, gpu_ctx, name, f, m):
    """so_eval_qr for linear (+- intercept) and logistic rows with n <= 8: same three references as above."""
    fam = getattr(oracle, name)
    if name == "LINEAR":
        truth = np.concatenate([[80.0], np.arange(f, dtype=float)])
    elif name == "LINEAR_NOINT":
        truth = np.arange(1, f + 1, dtype=float) / f
    else:
        truth = np.array([(((j % 7) - 3) / 4.0) for j in range(f + 1)])
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    n = p.size
    R = ds.qr(fam, p)
    assert gpu_ctx.stats().launches == 1
    assert R.shape == (n + 1, n + 1) and np.all(np.tril(R, -1) == 0.0)
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    G = np.block([[JtJ, Jtr[:, None]], [Jtr[None, :], np.array([[rtr]])]])
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    assert np.max(np.abs(R.T @ R - G) / scale) < 1e-11
    ab = oracle.jacobian_matrix(fam, X, y, p)
    q = oracle.qr(ab, n, pivoting=False)
    Rn = _normalise(R)
    Rref = q.Rmat()
    sref = np.where(np.diag(Rref) < 0, -1.0, 1.0)
    colnorm = np.sqrt(np.diag(JtJ))
    np.testing.assert_allclose(Rn[:n, :n] / colnorm, (Rref * sref[:, None]) / colnorm, rtol=0, atol=1e-9)
    np.testing.assert_allclose(Rn[:n, n], q.qtb[:n] * sref, rtol=0, atol=1e-9 * np.sqrt(rtr))
    Rl = _normalise(np.linalg.qr(ab, mode="r"))
    np.testing.assert_allclose(Rn / np.append(colnorm, np.sqrt(rtr)), Rl / np.append(colnorm, np.sqrt(rtr)), rtol=0, atol=1e-9)
    ds.free()


COLS_CASES = [("LINEAR_NOINT", 9), ("LINEAR", 10), ("LINEAR", 16), ("LINEAR_NOINT", 31), ("LINEAR", 31), ("LINEAR_NOINT", 32),
              ("LOGISTIC", 32), ("LINEAR_NOINT", 47), ("LINEAR", 46), ("LOGISTIC", 12), ("LINEAR_NOINT", 63)]


@pytest.mark.parametrize("name,f", COLS_CASES)
@pytest.mark.parametrize("m", [20_003, 70, 5])
def test_tsqr_of_the_wide_linear_predictor_families(native, oracle, gpu_ctx, name, f, m):
    """so_eval_qr with 8 < n <= 63 (one triangle per warp, lanes are columns — so_wide_qrw.cu): LM spec case 1 is n = 11,
    BASELINE config 2 is n = 33, config 5 has n = 32.  Same three references as the narrow cases."""
    fam = getattr(oracle, name)
    if name == "LINEAR":
        truth = np.concatenate([[80.0], np.arange(f, dtype=float)])
    elif name == "LINEAR_NOINT":
        truth = np.arange(1, f + 1, dtype=float) / f
    else:
        truth = np.array([(((j % 7) - 3) / 4.0) for j in range(f + 1)])
    X, y = oracle.generate(fam, m, f, truth)
    ds = native.Dataset(gpu_ctx, m, f).append(X, y)
    p = truth * 0.9 + 0.05
    n = p.size
    R = ds.qr(fam, p)
    assert gpu_ctx.stats().launches == 1
    assert R.shape == (n + 1, n + 1) and np.all(np.tril(R, -1) == 0.0) and np.all(np.isfinite(R))
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    G = np.block([[JtJ, Jtr[:, None]], [Jtr[None, :], np.array([[rtr]])]])
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G))) + 1e-300
    assert np.max(np.abs(R.T @ R - G) / scale) < 1e-11
    if m <= n:
        ds.free()
        return                                               # fewer rows than columns: R is not unique
    ab = oracle.jacobian_matrix(fam, X, y, p)
    Rn = _normalise(R)
    colnorm = np.sqrt(np.diag(JtJ))
    Rl = _normalise(np.linalg.qr(ab, mode="r"))
    np.testing.assert_allclose(Rn / np.append(colnorm, np.sqrt(rtr)), Rl / np.append(colnorm, np.sqrt(rtr)), rtol=0, atol=1e-9)
    if n <= 16:                                              # the restated 2n+1-pass QR.apply on the same rows
        q = oracle.qr(ab, n, pivoting=False)
        Rref = q.Rmat()
        sref = np.where(np.diag(Rref) < 0, -1.0, 1.0)
        np.testing.assert_allclose(Rn[:n, :n] / colnorm, (Rref * sref[:, None]) / colnorm, rtol=0, atol=1e-9)
        np.testing.assert_allclose(Rn[:n, n], q.qtb[:n] * sref, rtol=0, atol=1e-9 * np.sqrt(rtr))
    # bit-reproducible
    assert np.array_equal(ds.qr(fam, p), R)
    ds.free()


def test_tsqr_beyond_the_catalogue_is_an_error(native, gpu_ctx):
    ds = native.Dataset(gpu_ctx, 32, 64).append(np.ones((32, 64)), np.ones(32))
    with pytest.raises(native.NativeError):
        ds.qr(native.LINEAR_NOINT, np.ones(64))           # n = 64 > 63: a row no longer fits two columns per lane
    ds.free()


def test_tsqr_does_not_square_the_condition_number(native, oracle, gpu_ctx):
    """A degree-5 polynomial on t in [10, 10.5]: cond(J) ~ 1e11, so J^T J is numerically singular (cond ~ 1e22) and the
    Cholesky route loses everything, while Householder on the rows keeps R to cond * eps."""
    m, n = 200_000, 6
    rng = np.random.default_rng(3)
    t = 10.0 + 0.5 * rng.random(m)
    truth = np.array([1.0, -2.0, 0.5, 3.0, -1.0, 0.25])
    y = np.polyval(truth[::-1], t) + 0.1 * rng.standard_normal(m)
    ds = native.Dataset(gpu_ctx, m, 1).append(t.reshape(-1, 1), y)
    p = truth * 1.05
    R = _normalise(ds.qr(native.POLY, p))
    ab = oracle.jacobian_matrix(oracle.POLY, t.reshape(-1, 1), y, p)
    # reference: LAPACK QR of the rows in extended column scaling (both are backward stable: agree to cond * eps)
    Rl = _normalise(np.linalg.qr(ab, mode="r"))
    cond = np.linalg.cond(ab[:, :n] / np.linalg.norm(ab[:, :n], axis=0))
    assert cond > 1e8
    rel = np.abs(np.diag(R) - np.diag(Rl)) / np.abs(np.diag(Rl))
    assert rel.max() < cond * 1e-14, (rel, cond)
    # the normal-equations route: Cholesky of J^T J either breaks down or is far off in the trailing diagonal
    JtJ, Jtr, rtr = ds.normal_eq(native.POLY, p)
    try:
        Rc = np.linalg.cholesky(JtJ).T
        relc = np.abs(np.diag(Rc) - np.diag(Rl)[:n]) / np.abs(np.diag(Rl)[:n])
        assert relc.max() > 100 * rel.max()
    except np.linalg.LinAlgError:
        pass                                                  # not positive definite in floating point
    ds.free()


def test_tsqr_extreme_scales(native, oracle, gpu_ctx):
    """column norms far outside the normal range of the fast reciprocal / square-root seeds take the IEEE path"""
    m = 10_001
    rng = np.random.default_rng(11)
    t = rng.random(m)
    for scale in (1e-160, 1e150):
        y = scale * (2.0 - 1.3 * t + 0.01 * rng.standard_normal(m))
        ds = native.Dataset(gpu_ctx, m, 1).append(t.reshape(-1, 1), y)
        p = np.array([scale * 2.2, scale * -1.0])
        R = ds.qr(native.POLY, p)
        assert np.all(np.isfinite(R))
        ab = oracle.jacobian_matrix(oracle.POLY, t.reshape(-1, 1), y, p)
        Rl = _normalise(np.linalg.qr(ab, mode="r"))
        np.testing.assert_allclose(_normalise(R), Rl, rtol=1e-9, atol=1e-9 * np.abs(Rl).max())
        ds.free()

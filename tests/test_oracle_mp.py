"""The oracle's ANALYTIC derivatives against an independent differentiation.

The reference holds no test that pins gradient / J^T J / H d / line-search scores beyond 1e-5 (its own derivatives are
forward differences with eps = 1e-8), so the 1e-12 / 1e-11 parity of the CUDA kernels is parity with the oracle's
hand-derived formulas.  This file takes the hand out of it: every family's model f(p, x) and objective
  MSE families   L(p) = (1/m) sum_i (y_i - f(p, x_i))^2 / 2            (function/MSEFunction.scala:33-36,45-48)
  logistic       L(p) = (1/m) sum_i CE(o_i, y_i)                        (nnet/FFNeuralNetwork.scala:80-88)
is written down once more in 50-digit mpmath arithmetic, straight from the definition, and differentiated NUMERICALLY at
that precision (mpmath.diff: step ~1e-17, error ~1e-30) — no derivative formula is shared with oracle.c.  Compared at
1e-13: the loss, the gradient, the Jacobian/residual sums (J^T J, J^T r, r^T r with the reference's conventions r = |y - f|,
J = -sign(y - f) df/dp), the exact Hessian-vector product and phi(alpha), phi'(alpha) along a direction."""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")

CASES = [
    # name, f, parameters at which everything is evaluated
    ("EXPDECAY", 1, [2.2, -1.1]),
    ("GAUSS", 1, [2.1, 0.47, 0.06]),
    ("GAUSS_EXPBKG", 1, [2.2, 0.55, 0.055, 1.1, 3.3]),         # BASELINE config 3
    ("POLY", 1, [1.0, -2.0, 0.5, 3.0]),
    ("LINEAR", 3, [80.0, 0.3, 1.2, -2.0]),
    ("LINEAR_NOINT", 2, [0.7, -1.3]),
    ("LOGISTIC", 2, [0.1, 1.5, 0.4]),                           # FFNeuralNetworkTrainerSpec.scala:189-220
]
TRUTH = {"EXPDECAY": [2.0, -1.3], "GAUSS": [2.0, 0.5, 0.05], "GAUSS_EXPBKG": [2.0, 0.5, 0.05, 1.0, 3.0],
         "POLY": [1.0, -2.0, 0.5, 3.0], "LINEAR": [80.0, 0.0, 1.0, 2.0], "LINEAR_NOINT": [0.5, -1.0],
         "LOGISTIC": [0.1, 1.5, 0.4]}
M = 24


def _model(name, p, x):
    if name == "EXPDECAY":
        return p[0] * mp.exp(p[1] * x[0])
    if name == "GAUSS":
        return p[0] * mp.exp(-(x[0] - p[1]) ** 2 / (2 * p[2] ** 2))
    if name == "GAUSS_EXPBKG":
        return p[0] * mp.exp(-(x[0] - p[1]) ** 2 / (2 * p[2] ** 2)) + p[3] * mp.exp(-p[4] * x[0])
    if name == "POLY":
        return sum(p[k] * x[0] ** k for k in range(len(p)))
    if name == "LINEAR":
        return p[0] + sum(p[1 + j] * x[j] for j in range(len(x)))
    if name == "LINEAR_NOINT":
        return sum(p[j] * x[j] for j in range(len(x)))
    if name == "LOGISTIC":
        return 1 / (1 + mp.exp(-(p[0] + sum(p[1 + j] * x[j] for j in range(len(x))))))
    raise KeyError(name)


def _objective(name, X, y):
    m = len(y)

    def L(*p):
        s = mp.mpf(0)
        for xi, yi in zip(X, y):
            o = _model(name, p, xi)
            if name == "LOGISTIC":
                s += -(yi * mp.log(o) + (1 - yi) * mp.log(1 - o))          # targets are 0 or 1
            else:
                s += (yi - o) ** 2 / 2
        return s / m
    return L


def _partial(fn, p, j):
    order = tuple(1 if k == j else 0 for k in range(len(p)))
    return mp.diff(fn, tuple(p), order)


def _rel(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name,f,p", CASES)
def test_oracle_derivatives_against_50_digit_differentiation(oracle, name, f, p):
    mp.mp.dps = 50
    fam = getattr(oracle, name)
    Xf, yf = oracle.generate(fam, M, f, TRUTH[name], noise_sigma=0.05 if f == 1 else 1.0)
    X = [[mp.mpf(float(v)) for v in row] for row in Xf]          # the doubles themselves, exactly
    y = [mp.mpf(float(v)) for v in yf]
    pm = [mp.mpf(v) for v in p]
    n = len(p)
    L = _objective(name, X, y)
    tol = 1e-13

    # loss and gradient
    assert abs(oracle.apply(fam, Xf, yf, p) - float(L(*pm))) <= tol * abs(float(L(*pm)))
    g_ref = [_partial(L, pm, j) for j in range(n)]
    assert _rel(oracle.gradient(fam, Xf, yf, p), [float(v) for v in g_ref]) < tol

    # Jacobian/residual sums.  MSE families: r = |y - f|, J = -sign(y - f) df/dp, so J^T J = sum df df^T, J^T r = -sum df (y - f);
    # logistic: r = o - y and J = d(excitation)/dp = (1, x)  (Neuron.scala:65-73: gradient / residual)
    JtJ = mp.zeros(n, n)
    Jtr = [mp.mpf(0)] * n
    rtr = mp.mpf(0)
    for xi, yi in zip(X, y):
        if name == "LOGISTIC":
            df = [mp.mpf(1)] + list(xi)
            r = _model(name, pm, xi) - yi
            sgn = 1
        else:
            df = [_partial(lambda *q: _model(name, q, xi), pm, j) for j in range(n)]
            r = yi - _model(name, pm, xi)
            sgn = -1
        for a in range(n):
            Jtr[a] += sgn * df[a] * r
            for b in range(n):
                JtJ[a, b] += df[a] * df[b]
        rtr += r * r
    JtJ_o, Jtr_o, rtr_o = oracle.normal_eq(fam, Xf, yf, p)
    G = np.array([[float(JtJ[a, b]) for b in range(n)] for a in range(n)])
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
    assert np.max(np.abs(JtJ_o - G) / scale) < tol
    assert _rel(Jtr_o, [float(v) for v in Jtr]) < tol
    assert abs(rtr_o - float(rtr)) <= tol * float(rtr)

    # exact Hessian-vector product: d/d(eps) grad L(p + eps d) at 0
    d = [(-1.0) ** j * (0.3 + 0.1 * j) * max(abs(p[j]), 0.05) for j in range(n)]
    dm = [mp.mpf(v) for v in d]
    Hd_ref = [mp.diff(lambda e, j=j: _partial(L, [pm[k] + e * dm[k] for k in range(n)], j), 0) for j in range(n)]
    assert _rel(oracle.hess_vec(fam, Xf, yf, p, d), [float(v) for v in Hd_ref]) < 1e-12

    # line-search scores phi(alpha) = L(p + alpha d), phi'(alpha) (variable/Point.scala:54-57) on a short step
    ds = [0.05 * v for v in d]
    alphas = [0.0, 0.5, 1.0, 2.0]
    phi_o, dphi_o = oracle.line(fam, Xf, yf, p, ds, alphas)
    along = lambda a: L(*[pm[k] + a * mp.mpf(ds[k]) for k in range(n)])
    assert _rel(phi_o, [float(along(mp.mpf(a))) for a in alphas]) < tol
    assert _rel(dphi_o, [float(mp.diff(along, mp.mpf(a))) for a in alphas]) < 1e-12


# ---- the network restatement (FFNeuralNetwork / FFNeuralNetworkTrainer): back-propagated gradient vs differentiation ----
def _net_forward(sizes, cross_entropy, w, x):
    """logistic inner neurons; outputs: linear (MSE), logistic (cross-entropy, one output), soft-max (several).
    Weights layer by layer, neuron by neuron, bias first (FFNeuralNetwork.scala:66-68,252-263)."""
    out, k = list(x), 0
    for l in range(1, len(sizes)):
        exc = []
        for _ in range(sizes[l]):
            exc.append(w[k] + sum(w[k + 1 + i] * out[i] for i in range(sizes[l - 1])))
            k += sizes[l - 1] + 1
        last = l == len(sizes) - 1
        if last and not cross_entropy:
            out = exc
        elif last and sizes[-1] > 1:
            e = [mp.exp(v) for v in exc]
            out = [v / sum(e) for v in e]
        else:
            out = [1 / (1 + mp.exp(-v)) for v in exc]
    return out


NET_CASES = [((4, 5, 3), True, 0.0), ((4, 5, 3), True, 0.05),        # FFNeuralNetworkTrainerSpec.scala:353-376 (Iris shape)
             ((2, 3), True, 0.0),                                      # :222-260, soft-max regression
             ((2, 3, 1), True, 0.01), ((3, 2, 1), False, 0.0), ((3, 2, 1), False, 0.1)]


@pytest.mark.parametrize("sizes,cross_entropy,decay", NET_CASES)
def test_network_gradient_against_50_digit_differentiation(oracle, sizes, cross_entropy, decay):
    """Loss as FFNeuralNetwork.loss defines it (:76-104; the mean-squared-error loss is residual^2, NOT halved) and the
    trainer's fold (Trainer.scala:59-69: starts from |w|^2/2 decay, divided by data.size).  The back-propagated gradient
    is the derivative of the cross-entropy objective, and of HALF the squared residual for the MSE one — restated as the
    reference has it, and checked here against differentiation of exactly those two functions."""
    mp.mp.dps = 50
    rng = np.random.default_rng(12345)
    m, f, K = 12, sizes[0], sizes[-1]
    Xf = rng.normal(size=(m, f))
    if cross_entropy:
        lab = rng.integers(0, max(K, 2), size=m)
        Yf = np.eye(K)[lab] if K > 1 else lab.astype(float).reshape(m, 1)
    else:
        Yf = rng.normal(size=(m, K))
    n = oracle.net_nweights(sizes)
    wf = rng.normal(scale=0.5, size=n)
    X = [[mp.mpf(float(v)) for v in r] for r in Xf]
    Y = [[mp.mpf(float(v)) for v in r] for r in Yf]

    def objective(half_mse):
        def L(*w):
            s = mp.mpf(decay) * sum(v * v for v in w) / 2
            for xi, yi in zip(X, Y):
                o = _net_forward(sizes, cross_entropy, w, xi)
                if not cross_entropy:
                    s += sum((a - b) ** 2 for a, b in zip(o, yi)) / (2 if half_mse else 1)
                elif K > 1:
                    s += -sum(b * mp.log(a) for a, b in zip(o, yi))
                else:
                    s += -(yi[0] * mp.log(o[0]) + (1 - yi[0]) * mp.log(1 - o[0]))
            return s / m
        return L

    loss_o, grad_o = oracle.net_loss_grad(sizes, cross_entropy, Xf, Yf, wf, decay=decay)
    wm = [mp.mpf(float(v)) for v in wf]
    loss_ref = float(objective(False)(*wm))
    assert abs(loss_o - loss_ref) <= 1e-13 * abs(loss_ref)
    Lg = objective(True)
    grad_ref = [float(_partial(Lg, wm, j)) for j in range(n)]
    assert _rel(grad_o, grad_ref) < 1e-13


def test_negative_log_likelihood_against_50_digit_arithmetic(oracle):
    """sum_i -2 log pdf(p, x_i) (stats/package.scala:97-103) for the example's pdf (ExSignalPlusBackgroundFit.scala:78-114:
    fSig N(mu, sigma) + (1 - fSig) lambda exp(-lambda (x - xMin))) on the example's own 2000 variates."""
    mp.mp.dps = 50
    x = oracle.nll_example_data()
    for p in ([0.3, 125.0, 10.0, 0.05], [0.05, 125.0, 10.0, 1.0 / 15.0]):
        fs, mu, sg, lam = [mp.mpf(v) for v in p]
        ref = mp.mpf(0)
        for v in x:
            t = mp.mpf(float(v))
            ref += -2 * mp.log(fs * mp.exp(-((t - mu) / sg) ** 2 / 2) / (mp.sqrt(2 * mp.pi) * sg) + (1 - fs) * lam * mp.exp(-lam * (t - 50)))
        got = oracle.nll(oracle.PDF_GAUSS_EXPBKG, x, p, [50.0])
        assert abs(got - float(ref)) <= 1e-13 * abs(float(ref))

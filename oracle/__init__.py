"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``scalaopt_b200/`` does.  See ``oracle/oracle.h`` for the
reference file:line each function follows and for the parity status (the Scala reference cannot run
here: no JVM).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

LINEAR, LINEAR_NOINT, LOGISTIC, EXPDECAY, GAUSS, GAUSS_EXPBKG, POLY, MLP_MSE, MLP_CE = range(9)
SOFTMAX = 9        # FFNeuralNetwork(Vector(f, K)), cross-entropy + soft-max: evaluated by the net_* functions below


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, n) for n in ("oracle.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Fold(C.Structure):
    _fields_ = [("parts", C.c_int), ("threads", C.c_int)]


class JRandomState(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("have_next", C.c_int), ("next_gaussian", C.c_double)]


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_jrandom_init.argtypes = [C.POINTER(JRandomState), C.c_int64]
        L.orc_jrandom_next_double.argtypes = [C.POINTER(JRandomState)]
        L.orc_jrandom_next_double.restype = C.c_double
        L.orc_jrandom_next_gaussian.argtypes = [C.POINTER(JRandomState)]
        L.orc_jrandom_next_gaussian.restype = C.c_double
        L.orc_generate.argtypes = [C.c_int, C.c_int64, C.c_int, _dp, C.c_int, C.c_uint64, C.c_double,
                                   C.c_int64, C.c_int64, _dp, _dp]
        L.orc_family_nparams.argtypes = [C.c_int, C.c_int]
        L.orc_model.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int]
        L.orc_model.restype = C.c_double
        L.orc_apply.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, _Fold]
        L.orc_apply.restype = C.c_double
        L.orc_gradient.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, _Fold, _dp]
        L.orc_gradient_fd.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, C.c_double, _Fold, _dp]
        L.orc_jacobian_row.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_double, _dp]
        L.orc_jacobian_row.restype = C.c_double
        L.orc_jacobian_row_fd.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, _dp]
        L.orc_jacobian_row_fd.restype = C.c_double
        L.orc_jacobian_matrix.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, C.c_int,
                                          C.c_double, _dp]
        L.orc_normal_eq.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, _Fold, _dp, _dp, _dp]
        L.orc_net_row.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_net_loss_grad.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp, C.c_int64, _dp, C.c_double, _dp, _dp]
        L.orc_net_jacobian_matrix.argtypes = [_ip, C.c_int, C.c_int, _dp, _dp, C.c_int64, _dp, C.c_double, _dp]
        L.orc_normal_eq_comp.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp]
        L.orc_qr.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp]
        L.orc_qr.restype = C.c_int
        L.orc_qr_solution.argtypes = [_dp, _dp, _dp, _ip, C.c_int, _dp]
        L.orc_line.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, _Fold,
                               _dp, _dp]
        L.orc_hess_vec_fd.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, _dp, C.c_int, C.c_double,
                                      _Fold, _dp]
        L.orc_hess_vec.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, _dp, C.c_int, _Fold, _dp]
        L.orc_pdf.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_double]
        L.orc_pdf.restype = C.c_double
        L.orc_nll.argtypes = [C.c_int, _dp, C.c_int64, _dp, C.c_int, _dp, _Fold]
        L.orc_nll.restype = C.c_double
        L.orc_nll_example_data.argtypes = [C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        L.orc_lm_reference_passes.argtypes = [C.c_int, _dp, _dp, C.c_int64, C.c_int, _dp, C.c_int, C.c_int]
        L.orc_lm_reference_passes.restype = C.c_double
        _lib = L
    return _lib


def _a(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _xy(X, y):
    X = _a(X)
    y = _a(y).reshape(-1)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    assert X.shape[0] == y.shape[0]
    return X, y, X.shape[0], X.shape[1]


class JRandom:
    """java.util.Random; scala.util.Random(seed) delegates to it."""

    def __init__(self, seed: int):
        self._s = JRandomState()
        lib().orc_jrandom_init(C.byref(self._s), seed)

    def next_double(self) -> float:
        return lib().orc_jrandom_next_double(C.byref(self._s))

    def next_gaussian(self) -> float:
        return lib().orc_jrandom_next_gaussian(C.byref(self._s))


def generate(family, m, f, p_true, seed=12345, noise_sigma=1.0, row_offset=0, total_rows=0):
    p = _a(p_true)
    X = np.empty((m, f), dtype=np.float64)
    y = np.empty(m, dtype=np.float64)
    lib().orc_generate(family, m, f, _p(p), p.size, seed, noise_sigma, row_offset, total_rows, _p(X), _p(y))
    return X, y


def model(family, p, x):
    p = _a(p)
    x = _a(x).reshape(-1)
    return lib().orc_model(family, _p(p), p.size, _p(x), x.size)


def apply(family, X, y, p, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    return lib().orc_apply(family, _p(X), _p(y), m, f, _p(p), p.size, _Fold(parts, threads))


def gradient(family, X, y, p, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    g = np.empty(p.size)
    lib().orc_gradient(family, _p(X), _p(y), m, f, _p(p), p.size, _Fold(parts, threads), _p(g))
    return g


def gradient_fd(family, X, y, p, eps=1.0e-8, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    g = np.empty(p.size)
    lib().orc_gradient_fd(family, _p(X), _p(y), m, f, _p(p), p.size, eps, _Fold(parts, threads), _p(g))
    return g


def jacobian_matrix(family, X, y, p, fd=False, eps=1.0e-8):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    rows = np.empty((m, p.size + 1))
    lib().orc_jacobian_matrix(family, _p(X), _p(y), m, f, _p(p), p.size, int(fd), eps, _p(rows))
    return rows


def normal_eq(family, X, y, p, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    n = p.size
    JtJ = np.empty((n, n))
    Jtr = np.empty(n)
    rtr = C.c_double()
    lib().orc_normal_eq(family, _p(X), _p(y), m, f, _p(p), n, _Fold(parts, threads), _p(JtJ), _p(Jtr),
                        C.byref(rtr))
    return JtJ, Jtr, rtr.value


def normal_eq_comp(family, X, y, p, threads=1):
    """compensated sums (hi, lo), each n*n + n + 1 doubles laid out [JtJ row-major | Jtr | rtr]: the yardstick for
    summation error at scale, not a restatement of the reference"""
    X, y, m, f = _xy(X, y)
    p = _a(p)
    n = p.size
    hi = np.empty(n * n + n + 1)
    lo = np.empty(n * n + n + 1)
    lib().orc_normal_eq_comp(family, _p(X), _p(y), m, f, _p(p), n, threads, _p(hi), _p(lo))
    return hi, lo


def _net_args(sizes, X, Y):
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, int(sizes[0]))
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(X.shape[0], int(sizes[-1]))
    return sizes, X, Y


def net_nweights(sizes):
    return int(sum(sizes[l] * (sizes[l - 1] + 1) for l in range(1, len(sizes))))


def net_loss_grad(sizes, cross_entropy, X, Y, w, decay=0.0, want_grad=True):
    """FFNeuralNetworkTrainer(FFNeuralNetwork(Vector(sizes...)), data, decay).apply(w) / .gradient(w) with vector targets"""
    sizes, X, Y = _net_args(sizes, X, Y)
    w = _a(w)
    loss = C.c_double()
    g = np.empty(w.size) if want_grad else None
    rc = lib().orc_net_loss_grad(sizes.ctypes.data_as(_ip), sizes.size, int(cross_entropy), _p(X), _p(Y), X.shape[0], _p(w),
                                 float(decay), C.byref(loss), _p(g) if want_grad else None)
    if rc != 0:
        raise NotImplementedError("the reference has `???` for this network (mean squared error with several outputs)")
    return loss.value, g


def net_jacobian_matrix(sizes, cross_entropy, X, Y, w, decay=0.0):
    """trainer.jacobianAndResidualsMatrix(w): (m [+ n]) x (n + 1) rows (jacobian | residual)"""
    sizes, X, Y = _net_args(sizes, X, Y)
    w = _a(w)
    n = w.size
    rows = np.empty((X.shape[0] + (n if decay > 0.0 else 0), n + 1))
    rc = lib().orc_net_jacobian_matrix(sizes.ctypes.data_as(_ip), sizes.size, int(cross_entropy), _p(X), _p(Y), X.shape[0],
                                       _p(w), float(decay), _p(rows))
    if rc != 0:
        raise NotImplementedError("the reference has `???` for this network")
    return rows


class QRResult:
    def __init__(self, r, rdiag, qtb, ipvt, acnorms, bnorm):
        self.r, self.rdiag, self.qtb, self.ipvt, self.acnorms, self.bnorm = r, rdiag, qtb, ipvt, acnorms, bnorm
        self.n = rdiag.size

    def R(self, i, j):
        """QR.R(i, j), linalg/QR.scala:55-67"""
        if i == j:
            return self.rdiag[i]
        return self.r[i, j] if i < j else 0.0

    def Rmat(self):
        return np.array([[self.R(i, j) for j in range(self.n)] for i in range(self.n)])

    def solution(self):
        x = np.empty(self.n)
        lib().orc_qr_solution(_p(self.r), _p(self.rdiag), _p(self.qtb), self.ipvt.ctypes.data_as(_ip), self.n,
                              _p(x))
        return x


def qr(ab, n, pivoting=True):
    """QR.apply over a materialised augmented matrix ab[m, n+1]; raises ValueError where the reference
    throws IllegalArgumentException."""
    ab = _a(ab).reshape(-1, n + 1) if np.size(ab) else np.empty((0, n + 1))
    m = ab.shape[0]
    r = np.zeros((n, n))
    rdiag = np.zeros(n)
    qtb = np.zeros(n)
    ipvt = np.zeros(n, dtype=np.int32)
    acn = np.zeros(n)
    bn = C.c_double()
    rc = lib().orc_qr(_p(ab), m, n, int(pivoting), _p(r), _p(rdiag), _p(qtb), ipvt.ctypes.data_as(_ip), _p(acn),
                      C.byref(bn))
    if rc != 0:
        raise ValueError("requirement failed: R is not a square matrix")
    return QRResult(r, rdiag, qtb, ipvt, acn, bn.value)


def line(family, X, y, p, d, alphas, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p, d, al = _a(p), _a(d), _a(alphas)
    phi = np.empty(al.size)
    dphi = np.empty(al.size)
    lib().orc_line(family, _p(X), _p(y), m, f, _p(p), _p(d), p.size, _p(al), al.size, _Fold(parts, threads),
                   _p(phi), _p(dphi))
    return phi, dphi


def hess_vec(family, X, y, p, d, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p, d = _a(p), _a(d)
    out = np.empty(p.size)
    lib().orc_hess_vec(family, _p(X), _p(y), m, f, _p(p), _p(d), p.size, _Fold(parts, threads), _p(out))
    return out


def hess_vec_fd(family, X, y, p, d, eps=1.0e-8, parts=1, threads=1):
    X, y, m, f = _xy(X, y)
    p, d = _a(p), _a(d)
    out = np.empty(p.size)
    lib().orc_hess_vec_fd(family, _p(X), _p(y), m, f, _p(p), _p(d), p.size, eps, _Fold(parts, threads), _p(out))
    return out


def lm_reference_passes(family, X, y, p, threads=1):
    X, y, m, f = _xy(X, y)
    p = _a(p)
    return lib().orc_lm_reference_passes(family, _p(X), _p(y), m, f, _p(p), p.size, threads)


PDF_GAUSS_EXPBKG = 0


def nll(pdf, x, p, consts, parts=1, threads=1):
    """negLogLikelihood(pdf, data)(p) = sum -2 log pdf(p, x)   (stats/package.scala:97-103)"""
    x, p, consts = _a(x).reshape(-1), _a(p), _a(consts)
    return lib().orc_nll(pdf, _p(x), x.size, _p(p), p.size, _p(consts), _Fold(parts, threads))


def pdf(pdf_id, p, consts, x):
    p, consts = _a(p), _a(consts)
    return lib().orc_pdf(pdf_id, _p(p), p.size, _p(consts), float(x))


def nll_example_data(seed=12345, nb_events=2000, f_signal=0.3, mu=125.0, sigma=10.0, lam=1.0 / 20.0, x_min=50.0):
    """the sample of ExSignalPlusBackgroundFit.scala:39-55"""
    x = np.empty(nb_events)
    lib().orc_nll_example_data(seed, nb_events, f_signal, mu, sigma, lam, x_min, _p(x))
    return x

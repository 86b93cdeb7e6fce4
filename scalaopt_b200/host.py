"""ctypes binding of ``libscalaopt_host.so`` (``include/scalaopt_host.h``): the reference's optimizer-facing
interface — ``LevenbergMarquardt / BFGS / ConjugateGradient / NewtonCG / SteihaugCG .minimize(f, x0)(pars)`` —
over either a GPU-backed objective (``GpuMSEFunction``) or an objective made of Python callables.

The optimizers are the reference's host-side control logic restated in C++ (``scalaopt_b200/host/``); the
data passes they trigger run in ``libscalaopt_cuda.so``.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import native

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscalaopt_host.so")

LEVENBERG_MARQUARDT, BFGS, CONJUGATE_GRADIENT, NEWTON_CG, STEIHAUG_CG, NELDER_MEAD = range(6)
OK, MAX_ITER, ILLEGAL_ARG, GPU_ERROR, ERROR = range(5)

EXPORTS = ["soh_default_config", "soh_objective_from_callbacks", "soh_objective_gpu", "soh_objective_use_tsqr", "soh_objective_set_decay", "soh_objective_speculate", "soh_objective_gpu_nll", "soh_objective_free",
           "soh_objective_counters", "soh_objective_apply", "soh_objective_gradient", "soh_objective_dirder",
           "soh_objective_dir_hessian", "soh_objective_jacobian_rows", "soh_minimize", "soh_zoom_step_length",
           "soh_qr", "soh_gpu_lm", "soh_last_error"]


class MaxIterException(RuntimeError):
    """core/MaxIterException.scala — thrown (not returned) by the gradient methods and the line search."""


class IllegalArgumentException(ValueError):
    pass


class Config(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int32), ("eps", C.c_double),
                ("max_iter_line", C.c_int32), ("max_iter_zoom", C.c_int32), ("max_iter_first_time", C.c_int32),
                ("c1", C.c_double), ("c2", C.c_double), ("c3", C.c_double), ("cg_method", C.c_int32),
                ("delta0", C.c_double), ("delta_max", C.c_double), ("eta", C.c_double),
                ("x_tol", C.c_double), ("g_tol", C.c_double), ("ratio_tol", C.c_double), ("step_bound", C.c_double),
                ("epsmch", C.c_double), ("max_inner_iter", C.c_int32), ("max_step_length_iter", C.c_int32),
                ("use_pivoting", C.c_int32),
                ("c_reflection", C.c_double), ("c_expansion", C.c_double), ("c_contraction", C.c_double),
                ("c_shrink", C.c_double), ("rel_delta", C.c_double), ("abs_delta", C.c_double)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("inner_iterations", C.c_int32), ("gauss_newton_steps", C.c_int32),
                ("damped_steps", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("loss_calls", "gradient_calls", "dirder_calls", "hessvec_calls",
                                         "jacobian_calls", "k1_passes", "k2_passes", "k3_passes", "k4_passes",
                                         "cache_hits", "line_hits", "spec_passes", "spec_hits")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
APPLY_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, _dp, C.c_int)
GRAD_FN = C.CFUNCTYPE(None, C.c_void_p, _dp, C.c_int, _dp)
HESS_FN = C.CFUNCTYPE(None, C.c_void_p, _dp, _dp, C.c_int, _dp)
JAC_FN = C.CFUNCTYPE(None, C.c_void_p, _dp, C.c_int, _dp, C.c_int64)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    native.load()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.soh_default_config.argtypes = [C.c_int, C.POINTER(Config)]
    L.soh_objective_from_callbacks.argtypes = [C.c_int, APPLY_FN, GRAD_FN, HESS_FN, JAC_FN, C.c_int64, vp, C.c_double]
    L.soh_objective_from_callbacks.restype = vp
    L.soh_objective_gpu.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.soh_objective_gpu.restype = vp
    L.soh_objective_use_tsqr.argtypes = [vp, C.c_int]
    L.soh_objective_speculate.argtypes = [vp, C.c_int]
    L.soh_objective_set_decay.argtypes = [vp, C.c_double]
    L.soh_objective_gpu_nll.argtypes = [vp, vp, C.c_int, C.c_int, _dp, C.c_int]
    L.soh_objective_gpu_nll.restype = vp
    L.soh_objective_free.argtypes = [vp]
    L.soh_objective_free.restype = None
    L.soh_objective_counters.argtypes = [vp, C.POINTER(Counters)]
    L.soh_objective_apply.argtypes = [vp, _dp, _dp]
    L.soh_objective_gradient.argtypes = [vp, _dp, _dp]
    L.soh_objective_dirder.argtypes = [vp, _dp, _dp, _dp]
    L.soh_objective_dir_hessian.argtypes = [vp, _dp, _dp, _dp]
    L.soh_objective_jacobian_rows.argtypes = [vp, _dp, _dp, C.c_int64, C.POINTER(C.c_int64)]
    L.soh_minimize.argtypes = [C.c_int, vp, _dp, C.c_int, C.POINTER(Config), _dp, C.POINTER(Result)]
    L.soh_zoom_step_length.argtypes = [vp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.POINTER(Config), _dp]
    L.soh_qr.argtypes = [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int32), _dp, _dp, _dp]
    L.soh_gpu_lm.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, _dp]
    L.soh_last_error.restype = C.c_char_p
    _lib = L
    return L


def _vec(x):
    return np.ascontiguousarray(x, dtype=np.float64).reshape(-1)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _raise(rc):
    msg = load().soh_last_error().decode()
    if rc == MAX_ITER:
        raise MaxIterException(msg)
    if rc == ILLEGAL_ARG:
        raise IllegalArgumentException(msg)
    raise RuntimeError(f"libscalaopt_host error {rc}: {msg}")


def default_config(method: int, **overrides) -> Config:
    cfg = Config()
    load().soh_default_config(method, C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class Objective:
    """A DifferentiableObjectiveFunction (and MSEFunction when it has a Jacobian) behind an soh_objective handle."""

    def __init__(self, handle, n, keepalive=()):
        if not handle:
            raise RuntimeError(load().soh_last_error().decode())
        self._h, self.n, self._keep = C.c_void_p(handle), n, keepalive

    @classmethod
    def from_callables(cls, n, f, df=None, hessvec=None, jacobian_rows=None, m=0, eps=1.0e-8):
        """(f, df) -> ObjectiveFunctionWithGradient; f alone -> ObjectiveFunctionFiniteDiffDerivatives;
        jacobian_rows(p) -> array [m, n+1] of (J_i | r_i) makes it an MSEFunction."""
        def _apply(_u, x, nn):
            return float(f(np.ctypeslib.as_array(x, shape=(nn,)).copy()))

        def _grad(_u, x, nn, out):
            np.ctypeslib.as_array(out, shape=(nn,))[:] = df(np.ctypeslib.as_array(x, shape=(nn,)).copy())

        def _hess(_u, x, d, nn, out):
            np.ctypeslib.as_array(out, shape=(nn,))[:] = hessvec(np.ctypeslib.as_array(x, shape=(nn,)).copy(),
                                                                 np.ctypeslib.as_array(d, shape=(nn,)).copy())

        def _jac(_u, p, nn, rows, mm):
            np.ctypeslib.as_array(rows, shape=(mm * (nn + 1),))[:] = np.asarray(
                jacobian_rows(np.ctypeslib.as_array(p, shape=(nn,)).copy()), dtype=np.float64).reshape(-1)

        cbs = (APPLY_FN(_apply), GRAD_FN(_grad) if df else GRAD_FN(), HESS_FN(_hess) if hessvec else HESS_FN(),
               JAC_FN(_jac) if jacobian_rows else JAC_FN())
        h = load().soh_objective_from_callbacks(n, cbs[0], cbs[1], cbs[2], cbs[3], m, None, eps)
        return cls(h, n, keepalive=cbs)

    @classmethod
    def gpu(cls, ctx: native.Context, ds: native.Dataset, family: int, n: int, tsqr: bool = False, speculate: bool = True,
            decay: float = 0.0):
        """GpuMSEFunction(family, GpuDataSet(ds)); tsqr=True: Jacobian system from so_eval_qr instead of K2 + Cholesky;
        speculate=False: LM trial-point losses from K1 loss passes instead of speculative K2 passes"""
        h = load().soh_objective_gpu(ctx._h, ds._h, ds.f, family, n)
        if not speculate:
            load().soh_objective_speculate(h, 0)
        if decay and load().soh_objective_set_decay(h, float(decay)) != 0:
            raise IllegalArgumentException(load().soh_last_error().decode())
        if tsqr and load().soh_objective_use_tsqr(h, 1) != 0:
            raise RuntimeError(load().soh_last_error().decode())
        return cls(h, n, keepalive=(ctx, ds))

    @classmethod
    def gpu_nll(cls, ctx: native.Context, ds: native.Dataset, pdf: int, n: int, consts):
        """negLogLikelihood(pdf, GpuDataSet(ds)) _ with finite-difference derivatives"""
        c = _vec(consts)
        h = load().soh_objective_gpu_nll(ctx._h, ds._h, pdf, n, _ptr(c), c.size)
        return cls(h, n, keepalive=(ctx, ds))

    def _check(self, rc):
        if rc != OK:
            _raise(rc)

    def __call__(self, x):
        x = _vec(x)
        out = C.c_double()
        self._check(load().soh_objective_apply(self._h, _ptr(x), C.byref(out)))
        return out.value

    def gradient(self, x):
        x = _vec(x)
        g = np.empty(self.n)
        self._check(load().soh_objective_gradient(self._h, _ptr(x), _ptr(g)))
        return g

    def dirder(self, x, d):
        x, d = _vec(x), _vec(d)
        out = C.c_double()
        self._check(load().soh_objective_dirder(self._h, _ptr(x), _ptr(d), C.byref(out)))
        return out.value

    def dirHessian(self, x, d):
        x, d = _vec(x), _vec(d)
        out = np.empty(self.n)
        self._check(load().soh_objective_dir_hessian(self._h, _ptr(x), _ptr(d), _ptr(out)))
        return out

    def jacobianAndResidualsMatrix(self, p, capacity_rows=None):
        p = _vec(p)
        cap = capacity_rows or (self.n + 1)
        rows = np.empty((cap, self.n + 1))
        m = C.c_int64()
        self._check(load().soh_objective_jacobian_rows(self._h, _ptr(p), _ptr(rows), cap, C.byref(m)))
        return rows[: m.value]

    def counters(self) -> dict:
        c = Counters()
        load().soh_objective_counters(self._h, C.byref(c))
        return c.as_dict()

    def free(self):
        if self._h:
            load().soh_objective_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def minimize(method: int, f: Objective, x0, config: Config | None = None):
    """method.minimize(f, x0)(config) -> (x, Result).  Raises MaxIterException like the reference throws it."""
    x0 = _vec(x0)
    out = np.empty(x0.size)
    res = Result()
    cfg = config if config is not None else default_config(method)
    rc = load().soh_minimize(method, f._h, _ptr(x0), x0.size, C.byref(cfg), _ptr(out), C.byref(res))
    if rc != OK:
        _raise(rc)
    return out, res


def zoom_step_length(f: Objective, x, d, a, b, config: Config | None = None):
    x, d = _vec(x), _vec(d)
    out = np.empty(x.size)
    cfg = config if config is not None else default_config(BFGS)
    rc = load().soh_zoom_step_length(f._h, _ptr(x), _ptr(d), x.size, a, b, C.byref(cfg), _ptr(out))
    if rc != OK:
        _raise(rc)
    return out


def qr(ab, n, pivoting=True):
    """QR(ab, n, pivoting) over an in-memory DataSet[AugmentedRow]; dict with R, qtb, ipvt, acNorms, bNorm, solution."""
    ab = np.ascontiguousarray(ab, dtype=np.float64).reshape(-1, n + 1) if np.size(ab) else np.empty((0, n + 1))
    R, qtb, acn, sol = np.zeros((n, n)), np.zeros(n), np.zeros(n), np.zeros(n)
    ipvt = np.zeros(n, dtype=np.int32)
    bn = C.c_double()
    rc = load().soh_qr(_ptr(ab), ab.shape[0], n, int(pivoting), _ptr(R), _ptr(qtb),
                       ipvt.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(acn), C.byref(bn), _ptr(sol))
    if rc != OK:
        _raise(rc)
    return {"R": R, "qtb": qtb, "ipvt": ipvt, "acNorms": acn, "bNorm": bn.value, "solution": sol}


def gpu_lm(ctx: native.Context, ds: native.Dataset, add_origin: bool = True, tsqr: bool = False):
    """linear.lm(GpuDataSet(ds), addOrigin) (linear/package.scala:54-63): beta of (1, x) beta = y in the least-squares sense from
    one pass over the resident rows (K2, or the on-device TSQR) + the reference's QR(...).solution on n+1 rows."""
    beta = np.zeros(ds.f + (1 if add_origin else 0))
    rc = load().soh_gpu_lm(ctx._h, ds._h, ds.f, int(add_origin), int(tsqr), _ptr(beta))
    if rc != OK:
        _raise(rc)
    return beta


def bench_lm(ctx, ds, family, x0, barrier=lambda: None, max_over_ranks=lambda v: v, repeats=3, speculate=True, runs=4):
    """LevenbergMarquardt.minimize over the resident data set: iterations/s (one iteration = one pass through
    `iterate`, LevenbergMarquardt.scala:118-150 = one normal-equation evaluation + >= 1 loss evaluation + lmPar).
    `runs` fits are timed back to back inside one barrier bracket, so that the bracket's own cost (a cross-rank barrier)
    stays small against fits that take a few milliseconds on 8 GPUs; best of `repeats` brackets."""
    x0 = _vec(x0)
    best = None
    for _ in range(repeats):
        fs = [Objective.gpu(ctx, ds, family, x0.size, speculate=speculate) for _ in range(runs)]
        iters = inner = 0
        barrier()
        t0 = time.perf_counter()
        for f in fs:
            x, res = minimize(LEVENBERG_MARQUARDT, f, x0)
            iters += int(res.iterations)
            inner += int(res.inner_iterations)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        cnt = fs[-1].counters()
        for f in fs:
            f.free()
        if best is None or dt / runs < best["seconds"]:
            best = {"seconds": dt / runs, "iterations": iters // runs, "inner_iterations": inner // runs,
                    "iterations_per_s": iters / dt, "minimiser": [float(v) for v in x],
                    "k2_passes": cnt["k2_passes"], "k1_passes": cnt["k1_passes"], "spec_hits": cnt["spec_hits"],
                    "speculate": bool(speculate), "fits_per_bracket": runs}
    return best

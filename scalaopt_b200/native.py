"""ctypes binding of ``libscalaopt_cuda.so`` — the stand-in for the JNI glue a Scala caller would use
(``INTEGRATION.md``).  One Python call = one C-ABI call; nothing is computed here.

There is no CPU fallback: if the shared library has not been built, or no B200 is visible, loading or
``so_init`` fails with an exception.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCALAOPT_CUDA_LIB: another build of the same library (same-box A/B runs of kernel variants, tools/)
LIB_PATH = os.environ.get("SCALAOPT_CUDA_LIB") or os.path.join(_HERE, "libscalaopt_cuda.so")

LINEAR, LINEAR_NOINT, LOGISTIC, EXPDECAY, GAUSS, GAUSS_EXPBKG, POLY, MLP_MSE, MLP_CE, SOFTMAX = range(10)
FAMILY_NAMES = {LINEAR: "linear", LINEAR_NOINT: "linear_noint", LOGISTIC: "logistic", EXPDECAY: "expdecay",
                GAUSS: "gauss", GAUSS_EXPBKG: "gauss_expbkg", POLY: "poly", MLP_MSE: "mlp_mse", MLP_CE: "mlp_ce", SOFTMAX: "softmax"}
PDF_GAUSS_EXPBKG = 0
MAX_ALPHAS = 32
NCCL_UID_BYTES = 128

# every symbol include/scalaopt_cuda.h declares
EXPORTS = [
    "so_init", "so_nccl_unique_id", "so_init_rank", "so_shutdown", "so_last_error", "so_abi_version",
    "so_host_alloc", "so_host_free",
    "so_shard_plan", "so_dataset_create", "so_dataset_create_multi", "so_dataset_append", "so_dataset_load_raw", "so_dataset_clear", "so_dataset_generate", "so_dataset_read",
    "so_dataset_size", "so_dataset_head", "so_dataset_slice", "so_dataset_free",
    "so_eval_loss_grad", "so_eval_normal_eq", "so_eval_qr", "so_eval_line", "so_eval_hess_vec", "so_eval_nll", "so_stats", "so_stats_reset",
]


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libscalaopt_cuda error {code}: {msg}")
        self.code = code


class CallStats(C.Structure):
    _fields_ = [("rows", C.c_int64), ("bytes", C.c_int64), ("kernel_ms", C.c_double), ("device_ms", C.c_double),
                ("launches", C.c_int64), ("total_launches", C.c_int64), ("total_evals", C.c_int64),
                ("finish_ms", C.c_double), ("sum_kernel_ms", C.c_double), ("sum_device_ms", C.c_double),
                ("sum_finish_ms", C.c_double), ("min_kernel_ms", C.c_double), ("finished_on_device", C.c_int64)]


_dp = C.POINTER(C.c_double)
_lib = None


def load():
    """dlopen the library (raises if it has not been built — there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). scalaopt_b200 has no CPU fallback.")
    if "SCALAOPT_NCCL_LIB" not in os.environ:
        # prefer the NCCL that ships with torch, so a process that also imports torch maps one libnccl.so.2
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia")
            for base in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["SCALAOPT_NCCL_LIB"] = cand
                    break
        except Exception:
            pass
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.so_init.argtypes = [i32, C.POINTER(i32), C.POINTER(vp)]
    L.so_nccl_unique_id.argtypes = [vp, C.c_size_t]
    L.so_init_rank.argtypes = [i32, i32, i32, vp, C.c_size_t, C.POINTER(vp)]
    L.so_shutdown.argtypes = [vp]
    L.so_shutdown.restype = None
    L.so_last_error.argtypes = [vp]
    L.so_last_error.restype = C.c_char_p
    L.so_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.so_host_free.argtypes = [vp]
    L.so_host_free.restype = None
    L.so_shard_plan.argtypes = [i64, i32, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    L.so_dataset_create.argtypes = [vp, i64, i32, C.POINTER(vp)]
    L.so_dataset_create_multi.argtypes = [vp, i64, i32, i32, C.POINTER(vp)]
    L.so_dataset_append.argtypes = [vp, vp, vp, i64]
    L.so_dataset_load_raw.argtypes = [vp, C.c_char_p, C.c_char_p, i64, i64]
    L.so_dataset_clear.argtypes = [vp]
    L.so_dataset_generate.argtypes = [vp, i32, _dp, i32, C.c_uint64, C.c_double, i64]
    L.so_dataset_read.argtypes = [vp, i64, i64, vp, vp]
    L.so_dataset_size.argtypes = [vp, C.POINTER(i64)]
    L.so_dataset_head.argtypes = [vp, _dp, _dp]
    L.so_dataset_slice.argtypes = [vp, i64, i64, C.POINTER(vp)]
    L.so_dataset_free.argtypes = [vp]
    L.so_dataset_free.restype = None
    L.so_eval_loss_grad.argtypes = [vp, i32, _dp, i32, _dp, _dp]
    L.so_eval_normal_eq.argtypes = [vp, i32, _dp, i32, _dp, _dp, _dp]
    L.so_eval_qr.argtypes = [vp, i32, _dp, i32, _dp]
    L.so_eval_line.argtypes = [vp, i32, _dp, _dp, i32, _dp, i32, _dp, _dp]
    L.so_eval_hess_vec.argtypes = [vp, i32, _dp, _dp, i32, _dp]
    L.so_eval_nll.argtypes = [vp, i32, _dp, i32, _dp, i32, _dp]
    L.so_stats.argtypes = [vp, C.POINTER(CallStats)]
    L.so_stats_reset.argtypes = [vp]
    _lib = L
    return L


def _vec(x):
    return np.ascontiguousarray(x, dtype=np.float64).reshape(-1)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def shard_plan(m: int, f: int, G: int, g: int):
    """(begin, rows) of shard g of G (so_shard_plan; needs no GPU)."""
    b, r = C.c_int64(), C.c_int64()
    rc = load().so_shard_plan(m, f, G, g, C.byref(b), C.byref(r))
    if rc != 0:
        raise NativeError(rc, "so_shard_plan: bad arguments")
    return b.value, r.value


class PinnedBuffer:
    """fp64 array in page-locked host memory (so_host_alloc)."""

    def __init__(self, count: int):
        self._p = C.c_void_p()
        rc = load().so_host_alloc(count * 8, C.byref(self._p))
        if rc != 0:
            raise NativeError(rc, load().so_last_error(None).decode())
        self.array = np.ctypeslib.as_array(C.cast(self._p, _dp), shape=(count,))

    def free(self):
        if self._p:
            load().so_host_free(self._p)
            self._p = C.c_void_p()
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """``so_ctx``: either this process drives ``ngpus`` devices, or it is rank ``rank`` of ``world``."""

    def __init__(self, ngpus: int = 1, device_ids=None, *, rank: int | None = None, world: int = 1,
                 device_id: int = 0, nccl_uid: bytes | None = None):
        L = load()
        self._h = C.c_void_p()
        if rank is None:
            ids = None
            if device_ids is not None:
                ids = (C.c_int * ngpus)(*device_ids)
            rc = L.so_init(ngpus, ids, C.byref(self._h))
        else:
            buf = C.create_string_buffer(nccl_uid, NCCL_UID_BYTES) if nccl_uid else None
            rc = L.so_init_rank(device_id, rank, world, buf, NCCL_UID_BYTES if buf else 0, C.byref(self._h))
        if rc != 0:
            raise NativeError(rc, L.so_last_error(None).decode())
        self._datasets = []      # weak references: data sets are released before the context is shut down

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(NCCL_UID_BYTES)
        rc = load().so_nccl_unique_id(buf, NCCL_UID_BYTES)
        if rc != 0:
            raise NativeError(rc, load().so_last_error(None).decode())
        return buf.raw

    def check(self, rc):
        if rc != 0:
            raise NativeError(rc, load().so_last_error(self._h).decode())

    def stats(self) -> CallStats:
        st = CallStats()
        self.check(load().so_stats(self._h, C.byref(st)))
        return st

    def stats_reset(self):
        self.check(load().so_stats_reset(self._h))

    def close(self):
        if self._h:
            for ref in self._datasets:
                ds = ref()
                if ds is not None:
                    ds.free()
            self._datasets = []
            load().so_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Dataset:
    """``so_dataset``: observations resident in HBM."""

    def __init__(self, ctx: Context, m: int, f: int, _handle=None, _parent=None, ny: int = 1):
        self.ctx, self.f, self.ny = ctx, f, ny
        self._parent = _parent
        if _handle is not None:
            self._h = _handle
        else:
            self._h = C.c_void_p()
            ctx.check(load().so_dataset_create_multi(ctx._h, m, f, ny, C.byref(self._h)))
        import weakref
        ctx._datasets.append(weakref.ref(self))

    def append(self, X, y):
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = _vec(y)
        rows = y.size // self.ny
        assert X.size == rows * self.f and y.size == rows * self.ny, "X must be rows x f, y rows x ny"
        self.ctx.check(load().so_dataset_append(self._h, X.ctypes.data, y.ctypes.data, rows))
        return self

    def append_raw(self, x_addr: int, y_addr: int, rows: int):
        self.ctx.check(load().so_dataset_append(self._h, x_addr, y_addr, rows))
        return self

    def load_raw(self, path_x: str, path_y: str, rows: int, file_row_offset: int = 0):
        """append rows from raw little-endian fp64 files (row-major X, y)"""
        self.ctx.check(load().so_dataset_load_raw(self._h, os.fsencode(path_x), os.fsencode(path_y), rows, file_row_offset))
        return self

    def clear(self):
        self.ctx.check(load().so_dataset_clear(self._h))
        return self

    def generate(self, family, p_true, seed=12345, noise_sigma=1.0, row_offset=0):
        p = _vec(p_true)
        self.ctx.check(load().so_dataset_generate(self._h, family, _ptr(p), p.size, seed, noise_sigma, row_offset))
        return self

    def read(self, row_begin, rows):
        X = np.empty((rows, self.f))
        y = np.empty(rows) if self.ny == 1 else np.empty((rows, self.ny))
        self.ctx.check(load().so_dataset_read(self._h, row_begin, rows, X.ctypes.data, y.ctypes.data))
        return X, y

    def size(self) -> int:
        m = C.c_int64()
        self.ctx.check(load().so_dataset_size(self._h, C.byref(m)))
        return m.value

    def head(self):
        x = np.empty(self.f)
        y = C.c_double()
        self.ctx.check(load().so_dataset_head(self._h, _ptr(x), C.byref(y)))
        return x, y.value

    def slice(self, begin, end) -> "Dataset":
        h = C.c_void_p()
        self.ctx.check(load().so_dataset_slice(self._h, begin, end, C.byref(h)))
        return Dataset(self.ctx, end - begin, self.f, _handle=h, _parent=self, ny=self.ny)

    # ---- the four evaluations ----
    def loss_grad(self, family, p, want_grad=True):
        p = _vec(p)
        loss = C.c_double()
        g = np.empty(p.size) if want_grad else None
        self.ctx.check(load().so_eval_loss_grad(self._h, family, _ptr(p), p.size, C.byref(loss),
                                                _ptr(g) if want_grad else None))
        return loss.value, g

    def normal_eq(self, family, p):
        p = _vec(p)
        n = p.size
        JtJ = np.empty((n, n))
        Jtr = np.empty(n)
        rtr = C.c_double()
        self.ctx.check(load().so_eval_normal_eq(self._h, family, _ptr(p), n, _ptr(JtJ), _ptr(Jtr), C.byref(rtr)))
        return JtJ, Jtr, rtr.value

    def qr(self, family, p):
        """R factor ((n+1) x (n+1), upper triangular) of [J | r] by on-device Householder TSQR."""
        p = _vec(p)
        R = np.empty((p.size + 1, p.size + 1))
        self.ctx.check(load().so_eval_qr(self._h, family, _ptr(p), p.size, _ptr(R)))
        return R

    def line(self, family, p, d, alphas):
        p, d, al = _vec(p), _vec(d), _vec(alphas)
        phi = np.empty(al.size)
        dphi = np.empty(al.size)
        self.ctx.check(load().so_eval_line(self._h, family, _ptr(p), _ptr(d), p.size, _ptr(al), al.size, _ptr(phi),
                                           _ptr(dphi)))
        return phi, dphi

    def hess_vec(self, family, p, d):
        p, d = _vec(p), _vec(d)
        out = np.empty(p.size)
        self.ctx.check(load().so_eval_hess_vec(self._h, family, _ptr(p), _ptr(d), p.size, _ptr(out)))
        return out

    def nll(self, pdf, p, consts):
        """sum_i -2 log pdf(p, x_i) (negLogLikelihood, stats/package.scala:97-103)"""
        p, c = _vec(p), _vec(consts)
        out = C.c_double()
        self.ctx.check(load().so_eval_nll(self._h, pdf, _ptr(p), p.size, _ptr(c), c.size, C.byref(out)))
        return out.value

    def free(self):
        if self._h:
            if self.ctx._h:          # after so_shutdown the library has already released everything it owned
                load().so_dataset_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

"""Shared parity helpers: tolerances from BASELINE.json's north_star, all against the CPU oracle."""
import numpy as np

LOSS_RTOL = 1e-12      # loss and gradient within 1e-12 relative
GRAD_RTOL = 1e-12
GRAM_RTOL = 1e-11      # J^T J within 1e-11 relative (Frobenius, and element-wise against sqrt(G_ii G_jj))


def rel_err_vec(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def assert_loss(a, b, rtol=LOSS_RTOL):
    assert abs(a - b) <= rtol * abs(b), f"loss {a!r} vs oracle {b!r}: rel {abs(a - b) / abs(b):.3e}"


def assert_vec(a, b, rtol=GRAD_RTOL, what="gradient"):
    e = rel_err_vec(a, b)
    assert e <= rtol, f"{what}: max-norm relative error {e:.3e} > {rtol:.1e}\n got {a}\n ref {b}"


def assert_gram(G, Gref, rtol=GRAM_RTOL):
    G, Gref = np.asarray(G), np.asarray(Gref)
    fro = np.linalg.norm(G - Gref) / np.linalg.norm(Gref)
    assert fro <= rtol, f"JtJ Frobenius relative error {fro:.3e}"
    dg = np.sqrt(np.abs(np.diag(Gref)))
    scale = np.outer(dg, dg)
    scale[scale == 0] = 1.0
    ew = (np.abs(G - Gref) / scale).max()
    assert ew <= rtol, f"JtJ element-wise relative error {ew:.3e}"
    assert np.array_equal(G, G.T), "JtJ must be exactly symmetric"

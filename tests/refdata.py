"""Inputs of the reference's own specs, regenerated with the java.util.Random clone (seed 12345 everywhere)."""
import numpy as np


def lm_spec_data(oracle):
    """LevenbergMarquardtSpec.scala:34-81: one Random(12345) shared by both cases."""
    rnd = oracle.JRandom(12345)
    n, m = 10, 1000
    p0 = np.concatenate([[80.0], np.arange(n, dtype=float)])
    X = np.empty((m, n))
    y = np.empty(m)
    for i in range(m):
        X[i] = [rnd.next_double() for _ in range(n)]           # randomVec(n, random), :83-86
        y[i] = oracle.model(oracle.LINEAR, p0, X[i]) + rnd.next_gaussian()
    t = np.array([i / 10.0 for i in range(10)]).reshape(-1, 1)
    x0 = np.array([2.0, 1.0])
    y2 = np.array([oracle.model(oracle.EXPDECAY, x0, t[i]) + 0.1 * rnd.next_gaussian() for i in range(10)])
    return (X, y, p0), (t, y2, x0)


def logistic_spec_data(oracle):
    """FFNeuralNetworkTrainerSpec.scala:189-220: 10 000 rows, 2 N(0,1) inputs, Bernoulli labels, truth (0.1, 1.5, 0.4)."""
    rnd = oracle.JRandom(12345)
    truth = np.array([0.1, 1.5, 0.4])
    m = 10000
    X = np.empty((m, 2))
    y = np.empty(m)
    for i in range(m):
        X[i] = [rnd.next_gaussian(), rnd.next_gaussian()]
        prob = oracle.model(oracle.LOGISTIC, truth, X[i])
        y[i] = 1.0 if rnd.next_double() < prob else 0.0
    return X, y, truth


def oracle_objective(host, oracle, fam, X, y, n, fd_jacobian=False, exact_hessvec=True, parts=1):
    """The reference-side MSEFunction / trainer on a Seq data set, evaluated by the CPU oracle."""
    return host.Objective.from_callables(
        n,
        lambda p: oracle.apply(fam, X, y, p, parts=parts),
        lambda p: oracle.gradient(fam, X, y, p, parts=parts),
        (lambda p, d: oracle.hess_vec(fam, X, y, p, d, parts=parts)) if exact_hessvec else None,
        lambda p: oracle.jacobian_matrix(fam, X, y, p, fd=fd_jacobian),
        m=len(y))


def compressed_rows(oracle, fam, X, y, p):
    """What GpuMSEFunction.jacobianAndResidualsMatrix builds, from the oracle's sums (CPU stand-in for K2)."""
    JtJ, Jtr, rtr = oracle.normal_eq(fam, X, y, p)
    n = len(p)
    R = np.linalg.cholesky(JtJ).T
    q = np.linalg.solve(R.T, Jtr)
    rho = np.sqrt(max(0.0, rtr - q @ q))
    rows = np.zeros((n + 1, n + 1))
    rows[:n, :n] = R
    rows[:n, n] = q
    rows[n, n] = rho
    return rows

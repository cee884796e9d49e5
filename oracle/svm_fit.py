"""TEST INFRASTRUCTURE (oracle side) -- the linear SVM fit of the late-fusion step, numpy restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Reference call site: `linearClassifier = svm.LinearSVC(); linearClassifier.fit(svmTrainData, svmTrainLabels)`
(Sheet03/combinedModel.py:34-35), i.e. scikit-learn's LinearSVC with its defaults (penalty l2, loss squared_hinge,
C=1, one-vs-rest, fit_intercept with intercept_scaling 1, tol 1e-4, max_iter 1000).  The arithmetic lives in a
third-party dependency that is not under /root/reference: scikit-learn's vendored LIBLINEAR (the reference pins no
version; installed here: scikit-learn 1.9.0, LIBLINEAR 2.x), solver L2R_L2LOSS_SVC_DUAL.  Its published algorithm
(Hsieh et al., "A Dual Coordinate Descent Method for Large-scale Linear SVM", ICML 2008, Algorithm 1; liblinear
`solve_l2r_l1l2_svc`) is restated here:

  per class c (one-vs-rest, y_i = +1 if label_i == classes[c] else -1), features x_i extended by one constant `bias`:
    minimise over alpha >= 0:  1/2 alpha' (Q + D) alpha - e' alpha,   Q_ij = y_i y_j x_i.x_j,  D_ii = 1 / (2 C)
    w = sum_i alpha_i y_i x_i ;  one pass ("epoch") visits every i once:
       G  = y_i w.x_i - 1 + D_ii alpha_i
       PG = G if alpha_i > 0 else min(G, 0)
       if |PG| > 1e-12:  alpha_i <- max(alpha_i - G / (x_i.x_i + D_ii), 0);  w += (alpha_i - old) y_i x_i
    stop after an epoch over ALL samples with max PG - min PG <= tol (or max_iter epochs); samples with alpha_i = 0
    whose gradient exceeds the previous epoch's largest PG sit out until then ("shrinking", see fit_linear_svc)
  coef_[c] = w[:F], intercept_[c] = bias * w[F].

The problem is strictly convex in w, so every converging visiting order reaches the same coef_/intercept_.  Two
deliberate differences from LIBLINEAR, neither of which changes the optimum: the visiting order is a deterministic
affine permutation per epoch (`epoch_order`; LIBLINEAR draws random swaps from an unseeded generator, so the reference
itself is not reproducible beyond the solver tolerance), and shrunk samples leave the active set at the end of the
epoch instead of on the spot.  The CUDA solver (`va_svm_fit`) follows exactly this statement, including `epoch_order`,
so GPU and oracle agree to rounding.

Pinning: `oracle/make_golden.py` fits scikit-learn's LinearSVC (this container) at tight tolerance on a seeded
problem and stores data + coefficients in tests/golden/svm_fit.npz; tests check this restatement against it.
"""
from __future__ import annotations

import math

import numpy as np


def _mix32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def epoch_order_params(epoch: int, n: int):
    """(a, b) of the epoch's visiting order i -> (a * i + b) mod n, gcd(a, n) == 1; pure uint32 arithmetic shared with
    the CUDA solver (csrc/va_svm_fit.cu::epoch_order_params)."""
    h = _mix32((epoch * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF)
    a = h % n
    while math.gcd(a, n) != 1:
        a = (a + 1) % n
    b = _mix32(h ^ 0x85EBCA6B) % n
    return a, b


def epoch_order(epoch: int, n: int) -> np.ndarray:
    a, b = epoch_order_params(epoch, n)
    return ((a * np.arange(n, dtype=np.int64) + b) % n).astype(np.int64)


def fit_linear_svc(X: np.ndarray, labels: np.ndarray, C: float = 1.0, bias: float = 1.0, tol: float = 1e-4,
                   max_iter: int = 1000, shrinking: bool = True, return_steps: bool = False):
    """Returns (coef [n_classes or 1, F], intercept, classes, epochs per class).  Two classes give ONE row (positive
    class = classes[1]), as scikit-learn stores it.  One class problem after the other, LIBLINEAR's loop structure:

      active set = all samples; PGmax_old = +inf
      epoch: visit the active samples in `epoch_order(epoch, n_active)`; for sample s
                 G = y_s w.x_s - 1 + D alpha_s
                 alpha_s == 0 and G > PGmax_old  -> s leaves the active set (no update, not counted in the PG range)
                 PG = G if alpha_s > 0 else min(G, 0);  track max / min PG;  update alpha_s, w if |PG| > 1e-12
             PG range <= tol:  all samples were visited -> done;  else re-activate all, PGmax_old = +inf, next epoch
             otherwise PGmax_old = PGmax if PGmax > 0 else +inf
      (LIBLINEAR's second bound, alpha == U with G < PGmin_old, never applies: U = inf for the squared hinge.)

    The samples that left are removed at the END of the epoch, keeping the order of the rest (LIBLINEAR swaps a leaving
    sample with the last active one on the spot; which heuristic trims the active set does not change the optimum)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    V, F = X.shape
    classes = np.unique(labels)
    if len(classes) < 2:
        raise ValueError("This solver needs samples of at least 2 classes in the data")
    pos = classes[1:] if len(classes) == 2 else classes
    use_bias = bias > 0
    Xe = np.concatenate([X, np.full((V, 1), bias)], axis=1) if use_bias else X
    D = 0.5 / C
    QD = np.einsum("ij,ij->i", Xe, Xe) + D
    W, epochs, steps = [], [], []
    for c in pos:
        y = np.where(labels == c, 1.0, -1.0)
        w = np.zeros(Xe.shape[1])
        alpha = np.zeros(V)
        idx = np.arange(V)
        n, pgmax_old, it, nsteps = V, np.inf, 0, 0
        while it < max_iter:
            pgmax, pgmin = -np.inf, np.inf
            keep = np.ones(n, dtype=bool)
            if n > 0:
                a, b = epoch_order_params(it, n)
                for i in range(n):
                    p_ = (a * i + b) % n
                    s = idx[p_]
                    nsteps += 1
                    ai = alpha[s]
                    G = y[s] * (w @ Xe[s]) - 1.0 + ai * D
                    if ai == 0.0 and shrinking and G > pgmax_old:
                        keep[p_] = False
                        continue
                    PG = G if ai > 0 else min(G, 0.0)
                    pgmax, pgmin = max(pgmax, PG), min(pgmin, PG)
                    if abs(PG) > 1e-12:
                        new = max(ai - G / QD[s], 0.0)
                        w += ((new - ai) * y[s]) * Xe[s]
                        alpha[s] = new
            it += 1
            if pgmax - pgmin <= tol:
                if n == V and keep.all():        # the range was taken over ALL samples
                    break
                idx, n, pgmax_old = np.arange(V), V, np.inf
                continue
            idx = idx[:n][keep]
            n = len(idx)
            pgmax_old = pgmax if pgmax > 0 else np.inf
        W.append(w)
        epochs.append(it)
        steps.append(nsteps)
    W = np.array(W)
    coef = W[:, :F].copy()
    intercept = W[:, F] * bias if use_bias else np.zeros(len(pos))
    if return_steps:
        return coef, intercept, classes, np.array(epochs), np.array(steps)
    return coef, intercept, classes, np.array(epochs)


def decision(X, coef, intercept, classes):
    """LinearSVC.decision_function / predict (combinedModel.py:38)."""
    s = np.asarray(X, dtype=np.float64) @ coef.T + intercept
    if coef.shape[0] == 1:
        return s, classes[(s[:, 0] > 0).astype(int)]
    return s, classes[np.argmax(s, axis=1)]


def primal_objective(X, labels, coef, intercept, classes, C=1.0, bias=1.0):
    """LIBLINEAR's primal per class: 1/2 (|w|^2 + (b/bias)^2) + C sum max(0, 1 - y (w.x + b))^2 -- the regularised
    intercept is the extended feature's weight b / bias."""
    pos = classes[1:] if len(classes) == 2 else classes
    Y = np.where(labels[None, :] == pos[:, None], 1.0, -1.0)
    m = 1.0 - Y * (np.asarray(X, dtype=np.float64) @ coef.T + intercept).T
    wb = intercept / bias if bias > 0 else np.zeros_like(intercept)
    return 0.5 * ((coef ** 2).sum(1) + wb ** 2) + C * (np.maximum(m, 0.0) ** 2).sum(1)


def fit_primal_newton(X: np.ndarray, labels: np.ndarray, C: float = 1.0, bias: float = 1.0, gtol: float = 1e-10,
                      max_newton: int = 200, max_cg: int = 400):
    """The SAME optimum by an independent method, for checking the coordinate-descent results at sizes where
    `fit_linear_svc` takes a minute: all one-vs-rest problems at once by a finite Newton method on the primal
    (Keerthi & DeCoste, "A Modified Finite Newton Method for Fast Solution of Large Scale Linear SVMs", JMLR 2005):
        f(w) = 1/2 |w|^2 + C sum_i max(0, 1 - y_i w.x_i)^2      (x extended by the constant `bias`)
    Newton direction from the generalised Hessian I + 2C X_A' X_A (A = samples with y w.x < 1) by conjugate gradients,
    then an exact line search on the piecewise-quadratic f along it.  Stops when |grad f| <= gtol * |grad f(0)| for
    every class.  Returns (coef, intercept, classes, newton iterations)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    V, F = X.shape
    classes = np.unique(labels)
    pos = classes[1:] if len(classes) == 2 else classes
    K = len(pos)
    Y = np.where(labels[:, None] == pos[None, :], 1.0, -1.0)                  # [V, K]
    use_bias = bias > 0
    Xe = np.concatenate([X, np.full((V, 1), bias)], axis=1) if use_bias else X
    W = np.zeros((K, Xe.shape[1]))
    M = np.zeros((V, K))
    g0 = None
    its = 0
    for its in range(1, max_newton + 1):
        act = (1.0 - Y * M) > 0
        g = W + (2.0 * C * act * (M - Y)).T @ Xe
        gn = np.sqrt((g * g).sum(1))
        if g0 is None:
            g0 = np.maximum(gn, 1e-300)
        live = gn > gtol * g0
        if not live.any():
            break
        s = np.zeros_like(W)
        r = np.where(live[:, None], -g, 0.0)
        p = r.copy()
        rr = (r * r).sum(1)
        for _ in range(max_cg):
            Hp = p + (2.0 * C * act * (Xe @ p.T)).T @ Xe
            pHp = (p * Hp).sum(1)
            a = np.where(pHp > 0, rr / np.where(pHp > 0, pHp, 1.0), 0.0)
            s += a[:, None] * p
            r -= a[:, None] * Hp
            rr_new = (r * r).sum(1)
            if (np.sqrt(rr_new) <= 1e-3 * gn).all():
                break
            beta = np.where(rr > 0, rr_new / np.where(rr > 0, rr, 1.0), 0.0)
            p = r + beta[:, None] * p
            rr = rr_new
        ms = Xe @ s.T                                                          # [V, K]
        t = np.ones(K)
        ss, ws = (s * s).sum(1), (W * s).sum(1)
        for _ in range(50):                                                    # Newton on the monotone phi'(t)
            Mt = M + t[None, :] * ms
            at = (1.0 - Y * Mt) > 0
            d1 = ws + t * ss + 2.0 * C * (at * (Mt - Y) * ms).sum(0)
            d2 = ss + 2.0 * C * (at * ms * ms).sum(0)
            step = np.where(d2 > 0, d1 / np.where(d2 > 0, d2, 1.0), 0.0)
            t = t - step
            if (np.abs(step) <= 1e-14 * np.maximum(np.abs(t), 1.0)).all():
                break
        W += t[:, None] * s
        M += t[None, :] * ms
    coef = W[:, :F].copy()
    intercept = W[:, F] * bias if use_bias else np.zeros(K)
    return coef, intercept, classes, its

"""Prototype (CPU, numpy) for the next Tucker eigen-solve: tridiagonalise the unit-trace Gram matrix,
find only the eigenvalues above the rank threshold by Sturm-sequence bisection (one eigenvalue index
per lane on the device), get their vectors by inverse iteration + Gram-Schmidt, back-transform.
A division-free Sturm count (sturm_count_nodiv) is checked alongside.
Checked here against numpy.linalg.eigh on Gram matrices of states produced by the dense Tucker step
(tests/tucker_dense_ref.py): same rank decisions, projector difference at rounding level.

    python scripts/prototypes/tridiag_bisect_eig.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def householder_tridiag(A):
    """Returns d, e (sub-diagonal e[1:]) and the list of reflectors (v, beta, k) with T = Q^T A Q."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = -np.copysign(np.linalg.norm(x), x[0]) if x[0] != 0 else -np.linalg.norm(x)
        if alpha == 0.0:
            continue
        v = x.copy()
        v[0] -= alpha
        vv = v @ v
        if vv == 0.0:
            continue
        beta = 2.0 / vv
        p = beta * (A[k + 1:, k + 1:] @ v)
        w = p - (beta * 0.5 * (v @ p)) * v
        A[k + 1:, k + 1:] -= np.outer(v, w) + np.outer(w, v)
        A[k + 1, k] = A[k, k + 1] = alpha
        A[k + 2:, k] = A[k, k + 2:] = 0.0
        refl.append((v, beta, k))
    d = np.diag(A).copy()
    e = np.zeros(n)
    e[1:] = np.diag(A, -1)
    return d, e, refl


def sturm_count(d, e2, x):
    """Number of eigenvalues of the tridiagonal (d, e) smaller than x."""
    cnt, q = 0, 1.0
    for i in range(len(d)):
        q = d[i] - x - (e2[i] / q if i > 0 else 0.0)
        if q == 0.0:
            q = 1e-300
        if q < 0:
            cnt += 1
    return cnt


def sturm_count_nodiv(d, e2, x):
    """The same count without the division (a DP division per step is what would make the device
    version slow): signs of the leading principal minors p_i = (d_i - x) p_{i-1} - e_i^2 p_{i-2},
    rescaled by a power of two when they drift.  Agrees with sturm_count for every x above the
    rounding level of the matrix (checked in main)."""
    cnt, pm2, pm1 = 0, 0.0, 1.0
    for i in range(len(d)):
        p = (d[i] - x) * pm1 - (e2[i] * pm2 if i > 0 else 0.0)
        if p == 0.0:
            p = -1e-300 if pm1 > 0 else 1e-300
        if (p < 0) != (pm1 < 0):
            cnt += 1
        m = max(abs(p), abs(pm1))
        if m < 1e-100 or m > 1e100:
            sc = 2.0 ** (-np.frexp(m)[1])
            p, pm1 = p * sc, pm1 * sc
        pm2, pm1 = pm1, p
    return cnt


def eig_above(d, e, thr, kmax):
    """Eigenvalues > thr (at most kmax, largest first) by bisection."""
    n = len(d)
    e2 = e * e
    lo = min(d[i] - abs(e[i]) - (abs(e[i + 1]) if i + 1 < n else 0) for i in range(n))
    hi = max(d[i] + abs(e[i]) + (abs(e[i + 1]) if i + 1 < n else 0) for i in range(n))
    nabove = n - sturm_count(d, e2, thr)
    k = max(1, min(kmax, nabove))
    vals = []
    for j in range(k):                      # j-th largest = index n-1-j in ascending order
        a, b = lo, hi
        for _ in range(200):
            m = 0.5 * (a + b)
            if m == a or m == b:
                break
            if sturm_count(d, e2, m) <= n - 1 - j:
                a = m
            else:
                b = m
        vals.append(0.5 * (a + b))
    return np.array(vals), nabove


def inverse_iteration(d, e, lam, prev, rng):
    n = len(d)
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    shift = lam + 1e-14 * max(1.0, abs(lam)) * (1 + len(prev))     # off the exact eigenvalue
    x = rng.standard_normal(n)
    for _ in range(3):
        x = np.linalg.solve(T - shift * np.eye(n), x)                # tridiagonal solve on the device
        for p in prev:
            x -= (p @ x) * p
        x /= np.linalg.norm(x)
    return x


def leading_eigs(G, eps, rmax):
    n = G.shape[0]
    tr = np.trace(G)
    A = G / tr if tr > 0 else G
    d, e, refl = householder_tridiag(A)
    thr = (eps / np.sqrt(3.0)) ** 2                                  # lambda threshold at unit trace
    vals, nabove = eig_above(d, e, thr, rmax)
    rng = np.random.default_rng(0)
    vecs = []
    for lam in vals:
        vecs.append(inverse_iteration(d, e, lam, vecs, rng))
    Z = np.array(vecs).T
    for v, beta, k in reversed(refl):                                # back-transform: U = Q Z
        Z[k + 1:, :] -= beta * np.outer(v, v @ Z[k + 1:, :])
    return vals * tr, Z


def reference(G, eps, rmax):
    lam, V = np.linalg.eigh(G)
    lam, V = lam[::-1], V[:, ::-1]
    sig = np.sqrt(np.maximum(lam, 0))
    thr = eps * np.sqrt(np.maximum(lam, 0).sum()) / np.sqrt(3.0)
    r = 1
    for j in range(1, len(sig)):
        if sig[j] > thr and r < rmax:
            r += 1
        else:
            break
    return lam[:r], V[:, :r]


def main():
    import tucker_dense_ref as tdr
    from np_ref import vgrid
    rng = np.random.default_rng(1)
    worst = 0.0
    for n in [(11, 11, 11), (16, 12, 10), (32, 24, 20)]:
        _, V = vgrid(n, [-3, -2.5, -2], [3, 2.5, 2])
        for trial in range(6):
            f = np.zeros(n[0] * n[1] * n[2])
            for _ in range(rng.integers(1, 5)):
                c, s = rng.uniform(-0.8, 0.8, 3), rng.uniform(0.4, 1.0, 3)
                f += rng.uniform(0.2, 1.5) * np.exp(-0.5 * sum(((V[k] - c[k]) / s[k]) ** 2 for k in range(3)))
            X = f.reshape(n, order="F") + 1e-9 * rng.standard_normal(n)
            for eps in (1e-4, 1e-6):
                for k in range(3):
                    M = tdr.unfold(X, k)
                    G = M @ M.T
                    dd, ee, _ = householder_tridiag(G / np.trace(G))
                    for x in 10.0 ** rng.uniform(-15, 0, 20):
                        assert sturm_count(dd, ee * ee, x) == sturm_count_nodiv(dd, ee * ee, x)
                    lr, Vr = reference(G, eps, max(n))
                    lp, Vp = leading_eigs(G, eps, max(n))
                    assert len(lr) == len(lp), (n, trial, eps, k, len(lr), len(lp))
                    P = Vr @ Vr.T - Vp @ Vp.T
                    err = np.abs(P).max()
                    orth = np.abs(Vp.T @ Vp - np.eye(Vp.shape[1])).max()
                    worst = max(worst, err, orth)
    print("ranks agree on all cases; worst projector / orthogonality deviation = %.2e" % worst)


if __name__ == "__main__":
    main()

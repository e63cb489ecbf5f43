"""CPU study for the next step of the device PCG (csrc/poisson.cu): the Chronopoulos-Gear form of preconditioned CG
needs ONE global reduction per iteration instead of two (the device kernel pays a grid-wide barrier plus a cross-rank
sum for each: 6.2 us per iteration on one GPU, 16.8 us on eight).  The question is whether it still reaches the
reference's stopping rule |r| <= eps_machine |b| (Eigen's default tolerance, src/poisson.cpp:39-53) on the matrix the
oracle assembles, and in how many iterations — the recurrences for p.Ap are known to cost attainable accuracy.

    python scripts/prototypes/cg_single_reduction.py [hexes per edge, default 12]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from vlasovtucker_b200 import synthetic  # noqa: E402


def classic(S, dinv, b, x0, tol, maxit):
    """Eigen's ConjugateGradient.h:28-91 (two reductions per iteration: p.Ap, then r.r and r.z together)."""
    x = x0.copy()
    r = b - S @ x
    thr = tol * tol * (b @ b)
    if r @ r < thr:
        return x, 0
    z = dinv * r
    p = z.copy()
    absnew = r @ z
    for it in range(1, maxit + 1):
        t = S @ p
        alpha = absnew / (p @ t)
        x += alpha * p
        r -= alpha * t
        if r @ r < thr:
            return x, it
        z = dinv * r
        absold, absnew = absnew, r @ z
        p = z + (absnew / absold) * p
    return x, maxit


def chronopoulos_gear(S, dinv, b, x0, tol, maxit, replace_every=0):
    """One reduction per iteration: gamma = r.z, delta = z.Az and |r|^2 together; p.Ap follows from the recurrence
    delta - beta gamma / alpha_old.  replace_every > 0 recomputes r = b - A x and q = A p every that many iterations."""
    x = x0.copy()
    r = b - S @ x
    thr = tol * tol * (b @ b)
    z = dinv * r
    s = S @ z
    gamma, delta, rr = r @ z, z @ s, r @ r
    if rr < thr:
        return x, 0
    p = np.zeros_like(b)
    q = np.zeros_like(b)
    alpha, gamma_old = 1.0, 1.0
    for it in range(1, maxit + 1):
        if it == 1:
            beta, alpha = 0.0, gamma / delta
        else:
            beta = gamma / gamma_old
            alpha = gamma / (delta - beta * gamma / alpha)
        p = z + beta * p
        q = s + beta * q
        x += alpha * p
        r -= alpha * q
        if replace_every and it % replace_every == 0:
            r = b - S @ x
            q = S @ p
        z = dinv * r
        s = S @ z
        gamma_old = gamma
        gamma, delta, rr = r @ z, z @ s, r @ r      # the single reduction
        if rr < thr:
            return x, it
    return x, maxit


def main(h=12):
    nodes, tets, tris, ents = synthetic.kuhn_box(h, h, h, (1.0, 1.0, 1.0))
    m = oracle.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    po = oracle.Poisson(m)
    po.initialize()
    ip, ix, va = po.csr()[:3]
    n = len(ip) - 1
    A = sp.csr_matrix((va, ix, ip), shape=(n, n))
    U = sp.triu(A).tocsr()
    S = (U + sp.triu(A, 1).T).tocsr()       # the Upper-triangle view the reference solves with (poisson.h:41-44)
    dinv = 1.0 / S.diagonal()
    rng = np.random.default_rng(0)
    c = m.tetCentroid
    xs = np.sin(2 * np.pi * c[:, 0]) * np.cos(2 * np.pi * c[:, 1]) + 0.1 * rng.standard_normal(n)
    b = S @ xs
    b[0] = 0.0 if abs(S[0, 0] - 1.0) < 1e-14 and S[0].nnz == 1 else b[0]
    tol = np.finfo(float).eps
    x0 = np.zeros(n)
    print(f"rows {n}, nnz {S.nnz}, tolerance {tol:.2e} (Eigen default)")
    ref, it0 = classic(S, dinv, b, x0, tol, 2 * n)
    res0 = np.linalg.norm(b - S @ ref) / np.linalg.norm(b)
    print(f"classic PCG                      iterations {it0:5d}  true residual {res0:.2e}")
    for rep in (0, 50, 20):
        x, it = chronopoulos_gear(S, dinv, b, x0, tol, 4 * it0 + 50, rep)
        res = np.linalg.norm(b - S @ x) / np.linalg.norm(b)
        dx = np.linalg.norm(x - ref) / np.linalg.norm(ref)
        tag = "no residual replacement" if rep == 0 else f"replacement every {rep}"
        print(f"one reduction, {tag:26s} iterations {it:5d}{' (not converged)' if it >= 4 * it0 + 50 else ''}  "
              f"true residual {res:.2e}  |x - x_classic| / |x| {dx:.2e}")
    # a looser rule, for comparison: where both stand at 1e-12
    _, it12 = classic(S, dinv, b, x0, 1e-12, 2 * n)
    _, it12b = chronopoulos_gear(S, dinv, b, x0, 1e-12, 2 * n)
    print(f"at tolerance 1e-12: classic {it12}, one reduction {it12b} iterations")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 12)

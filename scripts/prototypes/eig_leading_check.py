"""Builds eig_leading_check.cpp (g++) and compares the prototype eigen-solve of eig_leading.h with
numpy.linalg.eigh: ranks under the reference's rule (tucker.cpp:66-98: sigma_j > eps |sigma| / sqrt 3,
at least one, at most rmax) and the projector onto the kept subspace.

    python scripts/prototypes/eig_leading_check.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    so = os.path.join(tempfile.gettempdir(), "vt_eig_leading_check.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", HERE,
                           os.path.join(HERE, "eig_leading_check.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.leading_eig.restype = C.c_int
    return lib


def proto(lib, G, eps, kmax):
    n = G.shape[0]
    G = np.ascontiguousarray(G, np.float64)
    vals = np.zeros(kmax)
    U = np.zeros((kmax, n))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    r = lib.leading_eig(dp(G), n, C.c_double(eps), kmax, dp(vals), dp(U))
    return vals[:r], U[:r].T


def reference(G, eps, kmax):
    lam, V = np.linalg.eigh(G)
    lam, V = np.maximum(lam[::-1], 0), V[:, ::-1]
    thr2 = eps * eps * lam.sum() / 3.0
    r = int((lam > thr2).sum())
    r = max(1, min(kmax, r))
    return lam[:r], V[:, :r], lam, thr2


def compare(lib, G, eps, kmax, tag, stats):
    lr, Vr, lam, thr2 = reference(G, eps, kmax)
    lp, Vp = proto(lib, G, eps, kmax)
    # an eigenvalue within rounding of the threshold may legitimately fall on either side
    margin = np.abs(lam - thr2).min() if len(lam) else 1.0
    if len(lp) != len(lr):
        assert margin <= 1e-15 * max(lam.sum(), 1e-300), (tag, len(lr), len(lp), margin)
        stats["edge"] += 1
        return
    scale = max(lam.sum(), 1e-300)
    assert np.abs(lp - lr).max() <= 1e-14 * scale, (tag, lp, lr)
    orth = np.abs(Vp.T @ Vp - np.eye(Vp.shape[1])).max()
    assert orth <= 1e-12, (tag, orth)
    # projector difference weighted by what it does to the tensor: sqrt(lambda) carries the magnitude
    kept_gap = (lr[-1] - (lam[len(lr)] if len(lr) < len(lam) else 0.0)) / scale
    P = np.abs(Vr @ Vr.T - Vp @ Vp.T).max()
    stats["worst_proj"] = max(stats["worst_proj"], P if kept_gap > 1e-12 else 0.0)
    # residual of the kept pairs, the criterion that does not depend on gaps
    res = np.abs(G @ Vp - Vp * lp[None, :]).max() / scale
    stats["worst_res"] = max(stats["worst_res"], res)
    assert res <= 1e-13, (tag, res)
    stats["n"] += 1


def main():
    import tucker_dense_ref as tdr
    from np_ref import vgrid
    lib = build()
    rng = np.random.default_rng(3)
    stats = dict(n=0, edge=0, worst_proj=0.0, worst_res=0.0)
    # special matrices
    for n in (2, 3, 11, 32, 64):
        compare(lib, np.zeros((n, n)), 1e-6, n, "zero", stats)
        compare(lib, np.eye(n), 1e-6, n, "identity", stats)
        u = rng.standard_normal(n)
        compare(lib, np.outer(u, u), 1e-6, n, "rank1", stats)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.array([1.0, 1.0, 0.5, 0.5, 0.5][:n] + [1e-9] * max(0, n - 5))
        compare(lib, (Q * lam) @ Q.T, 1e-6, n, "clusters", stats)
        compare(lib, (Q * lam) @ Q.T, 1e-6, 3, "clusters capped", stats)
        lam = 10.0 ** (-np.arange(n) * 0.7)
        compare(lib, (Q * lam) @ Q.T, 1e-6, n, "geometric", stats)
    # Gram matrices of unfoldings of realistic states
    for n in [(11, 11, 11), (16, 12, 10), (32, 24, 20), (20, 64, 18)]:
        _, V = vgrid(n, [-3, -2.5, -2], [3, 2.5, 2])
        for trial in range(5):
            f = np.zeros(n[0] * n[1] * n[2])
            for _ in range(rng.integers(1, 5)):
                c, s = rng.uniform(-0.8, 0.8, 3), rng.uniform(0.4, 1.0, 3)
                f += rng.uniform(0.2, 1.5) * np.exp(-0.5 * sum(((V[k] - c[k]) / s[k]) ** 2 for k in range(3)))
            X = f.reshape(n, order="F") * (1 + 1e-7 * rng.standard_normal(n))
            for eps in (1e-4, 1e-6):
                for k in range(3):
                    M = tdr.unfold(X, k)
                    compare(lib, M @ M.T, eps, n[k], (n, trial, eps, k), stats)
                    compare(lib, M @ M.T, eps, 8, (n, trial, eps, k, "cap 8"), stats)
    print("cases %d (+%d at the threshold's rounding edge); worst residual %.1e, worst projector deviation %.1e"
          % (stats["n"], stats["edge"], stats["worst_res"], stats["worst_proj"]))


if __name__ == "__main__":
    main()

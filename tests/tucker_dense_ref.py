"""numpy statement of the formulation the device Tucker path uses (vlasovtucker_b200/csrc/tucker.cu):
every Tucker sum of Solver<Tucker>::_UpdatePDF (src/solver.cpp:141-212) evaluated densely and
rounded by the truncated HOSVD that Tucker::Compress (src/tucker.cpp:66-98, 442-465) amounts to.
Test plumbing: test_tucker_cpu.py checks it against the oracle's real Tucker algebra, so the
claim "Compress == truncated HOSVD of the represented tensor" is itself under test."""
import numpy as np

from np_ref import vgrid


def unfold(X, k):
    return np.moveaxis(X, k, 0).reshape(X.shape[k], -1, order="F")


def hosvd_factors(X, eps, rmax):
    Us = []
    for k in range(3):
        U, s, _ = np.linalg.svd(unfold(X, k), full_matrices=False)
        thr = eps * np.linalg.norm(s) / np.sqrt(3.0)
        r = 1
        for j in range(1, len(s)):
            if s[j] > thr and r < rmax:
                r += 1
            else:
                break
        Us.append(U[:, :r])
    return Us


def project(X, Us):
    core = np.einsum("abc,ai,bj,ck->ijk", X, Us[0], Us[1], Us[2])
    return core


def truncate(X, eps, rmax):
    Us = hosvd_factors(X, eps, rmax)
    core = project(X, Us)
    return np.einsum("ijk,ai,bj,ck->abc", core, Us[0], Us[1], Us[2]), core.shape


def step_dense(f, nbr, area, volume, normal, bc, n, vmin, vmax, qm, E, dt, eps, rmax, pair_codes=(0, 1), absorbing=3):
    """One Tucker-format step; f (nT, N) rows i0-fastest.  Returns (new f, ranks)."""
    nT = len(nbr)
    step, V = vgrid(n, vmin, vmax)
    shp = tuple(n)
    Vg = [v.reshape(shp, order="F") for v in V]
    out = np.empty_like(f)
    ranks = np.zeros((nT, 3), int)
    for t in range(nT):
        A = f[t].reshape(shp, order="F")
        rhs = np.zeros(shp)
        for j in range(4):
            vn = normal[t, j, 0] * Vg[0] + normal[t, j, 1] * Vg[1] + normal[t, j, 2] * Vg[2]
            va, _ = truncate(np.abs(vn), eps, 6)
            if bc[t, j] in pair_codes:
                B = f[nbr[t, j]].reshape(shp, order="F")
                flux = 0.5 * (vn * (B + A) - va * (B - A))
            elif bc[t, j] == absorbing:
                flux = 0.5 * (vn * A + va * A)
            else:
                flux = vn * A
            rhs = rhs - (area[t, j] / volume[t]) * flux
            rhs, _ = truncate(rhs, eps, rmax)
        for k in range(3):
            up = np.zeros(shp)
            dn = np.zeros(shp)
            sl_hi = [slice(None)] * 3
            sl_lo = [slice(None)] * 3
            sl_hi[k] = slice(1, None)
            sl_lo[k] = slice(None, -1)
            up[tuple(sl_lo)] = A[tuple(sl_hi)]
            dn[tuple(sl_hi)] = A[tuple(sl_lo)]
            rhs = rhs - (qm * E[t, k]) * (up - dn) / (2 * step[k])
        rhs, _ = truncate(rhs, eps, rmax)
        X, r = truncate(A + dt * rhs, eps, rmax)
        out[t] = X.ravel(order="F")
        ranks[t] = r
    return out, ranks

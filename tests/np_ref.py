"""Tiny numpy restatement of the interior/periodic part of Solver<Full>::_UpdatePDF
(src/solver.cpp:141-212, 314-328, 363-404) on flattened tables — test plumbing for the
partition tests, itself checked against the oracle in test_partition_cpu.py."""
import numpy as np


def vgrid(n, vmin, vmax):
    step = [(vmax[k] - vmin[k]) / (n[k] - 1) for k in range(3)]
    ax = [vmin[k] + np.arange(n[k]) * step[k] for k in range(3)]
    V = np.meshgrid(*ax, indexing="ij")
    return step, [v.ravel(order="F") for v in V]


def step_tables(f, nbr, area, volume, normal, n, vmin, vmax, qm, E, dt, rows=None):
    """One explicit step for the tets in `rows` (default: all rows of `nbr`); `f` holds owned rows
    followed by ghost rows, `nbr` indexes into it.  Returns the new owned rows."""
    nO = len(nbr)
    step, V = vgrid(n, vmin, vmax)
    N = n[0] * n[1] * n[2]
    out = np.empty((nO, N))
    idx = np.arange(N).reshape(n[2], n[1], n[0]).transpose(2, 1, 0)   # idx[i0,i1,i2] = flat index
    plus = [np.roll(idx, -1, axis=k).transpose(2, 1, 0).ravel() for k in range(3)]
    minus = [np.roll(idx, 1, axis=k).transpose(2, 1, 0).ravel() for k in range(3)]
    for t in range(nO):
        A = f[t]
        rhs = np.zeros(N)
        for j in range(4):
            B = f[nbr[t, j]]
            vn = normal[t, j, 0] * V[0] + normal[t, j, 1] * V[1] + normal[t, j, 2] * V[2]
            flux = 0.5 * (vn * (B + A) - np.abs(vn) * (B - A))
            rhs = rhs - (area[t, j] / volume[t]) * flux
        for k in range(3):
            der = (A[plus[k]] - A[minus[k]]) / (2 * step[k])
            rhs = rhs - (qm * E[t, k]) * der
        out[t] = A + dt * rhs
    return out

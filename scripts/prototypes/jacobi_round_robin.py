"""CPU statement of the parallel two-sided Jacobi method of csrc/tucker_slab.cu (jacobi3, jacobi_warp): circle-method
pairing, one rotation per pair and round, G <- J^T G J applied 2x2-block pair by block pair over the upper
triangle with a mirrored write, V <- V J; absolute + relative skip threshold (the device code adds the
tolerance-aware criterion, the active-index sets and the pivoted Cholesky reduction in front).  Checks eigenvalues and
eigenvectors against numpy.linalg.eigh on Gram matrices like the ones the Tucker rounding sees."""
import numpy as np


def jacobi(G, max_sweeps=30):
    n = G.shape[0]
    ne = n + (n & 1)
    A = np.zeros((ne, ne))
    A[:n, :n] = G
    V = np.eye(ne)
    thr = 1e-18 * np.trace(G)
    h = ne // 2
    sweeps = 0
    rounds_with_work = 0
    for sweep in range(max_sweeps):
        count = 0
        for s in range(ne - 1):
            rot = []
            for slot in range(h):
                if slot == 0:
                    p, q = ne - 1, s
                else:
                    p, q = (s + slot) % (ne - 1), (s - slot + ne - 1) % (ne - 1)
                if p > q:
                    p, q = q, p
                gpq, gpp, gqq = A[p, q], A[p, p], A[q, q]
                a = abs(gpq)
                c, sn, t, r = 1.0, 0.0, 0.0, 0
                if a > thr and a * a > 1e-32 * abs(gpp * gqq):
                    tau = (gqq - gpp) / (2.0 * gpq)
                    t = 1.0 / (abs(tau) + np.sqrt(tau * tau + 1.0))
                    if tau < 0:
                        t = -t
                    c = 1.0 / np.sqrt(t * t + 1.0)
                    sn = t * c
                    r = 1
                    count += 1
                rot.append((p, q, c, sn, t, r))
            assert sorted(x for pq in rot for x in pq[:2]) == list(range(ne))
            if any(r[5] for r in rot):
                rounds_with_work += 1
            B = A.copy()   # every block reads the state before the round (blocks are disjoint)
            for i in range(h):
                for j in range(i, h):
                    P, Q, ci, si, ti, ri = rot[i]
                    R, S, cj, sj, tj, rj = rot[j]
                    if not (ri or rj):
                        continue
                    if i == j:
                        gpq = A[P, Q]
                        B[P, P] = A[P, P] - ti * gpq
                        B[Q, Q] = A[Q, Q] + ti * gpq
                        B[P, Q] = B[Q, P] = 0.0
                        continue
                    mPR, mPS, mQR, mQS = A[P, R], A[P, S], A[Q, R], A[Q, S]
                    nPR, nPS = cj * mPR - sj * mPS, sj * mPR + cj * mPS
                    nQR, nQS = cj * mQR - sj * mQS, sj * mQR + cj * mQS
                    oPR, oQR = ci * nPR - si * nQR, si * nPR + ci * nQR
                    oPS, oQS = ci * nPS - si * nQS, si * nPS + ci * nQS
                    B[P, R] = B[R, P] = oPR
                    B[P, S] = B[S, P] = oPS
                    B[Q, R] = B[R, Q] = oQR
                    B[Q, S] = B[S, Q] = oQS
            A = B
            for (P, Q, c, sn, t, r) in rot:
                if not r:
                    continue
                vp, vq = V[:n, P].copy(), V[:n, Q].copy()
                V[:n, P] = c * vp - sn * vq
                V[:n, Q] = sn * vp + c * vq
        sweeps += 1
        if count == 0:
            break
    return np.diag(A)[:n].copy(), V[:n, :n].copy(), sweeps, rounds_with_work


def gram_cases(rng, sizes=(48, 33, 24, 17)):
    for n in sizes:
        # smooth, quickly decaying spectrum: unfolding of a sum of shifted Maxwellians
        x = np.linspace(-4, 4, n)
        X = sum(np.exp(-(x[:, None, None] - a) ** 2 - 1.3 * (x[None, :, None] - b) ** 2 - 0.7 * (x[None, None, :] - c) ** 2)
                for a, b, c in rng.uniform(-1, 1, (8, 3)))
        for mode in range(3):
            Xk = np.moveaxis(X, mode, 0).reshape(n, -1)
            yield f"maxwell n={n} mode={mode}", Xk @ Xk.T
        Z = rng.standard_normal((n, 3 * n))
        yield f"random n={n}", Z @ Z.T
        u = rng.standard_normal(n)
        yield f"rank one n={n}", np.outer(u, u)
        yield f"zero n={n}", np.zeros((n, n))


def main(sizes=(48, 33, 24, 17)):
    rng = np.random.default_rng(1)
    worst = 0.0
    for name, G in gram_cases(rng, sizes):
        lam, V, sweeps, work = jacobi(G)
        n = G.shape[0]
        tr = max(np.trace(G), 1e-300)
        ref = np.linalg.eigvalsh(G)[::-1]
        order = np.argsort(-lam, kind="stable")
        e_val = np.max(np.abs(lam[order] - ref)) / tr
        e_orth = np.max(np.abs(V.T @ V - np.eye(n)))
        e_res = np.max(np.abs(G @ V - V * lam)) / tr
        worst = max(worst, e_val, e_orth, e_res)
        print(f"{name:28s} sweeps {sweeps:2d}  rounds with work {work:4d}  eigenvalues {e_val:.1e}  orthogonality {e_orth:.1e}  residual {e_res:.1e}")
    assert worst < 1e-13, worst
    print("ok")


if __name__ == "__main__":
    main()

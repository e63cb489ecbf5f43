"""Tucker oracle (oracle/oracle_tucker.cpp) against the reference's own Tucker test
(test/tucker_test.cpp live block), the rank known-answers of SURVEY.md §8c, and the dense
truncated-HOSVD formulation the device path uses (tests/tucker_dense_ref.py)."""
import numpy as np
import pytest

from conftest import face_bc_arrays, mesh_path, rel_l2
import tucker_dense_ref as tdr
from np_ref import vgrid


def test_tucker_test_invariance(oracle_mod):
    """test/tucker_test.cpp:171-196: x += x; x -= 0.5x; x += x; x -= 0.5x; Compress(1e-6, 4), ten
    times: every printed Reconstructed() equals the first one."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (5, 3, 3))                    # Tensor::setRandom range
    t = oracle_mod.TuckerObj.from_full(x, 1e-6, 4)
    assert t.ranks() == (4, 3, 3)
    first = t.reconstructed()
    for _ in range(10):
        for expect in (8, 16, 32, 64):                   # ranks add under operator+ (tucker.cpp:190-228)
            t.axpy(1.0 if expect in (8, 32) else -0.5, t.clone())
            assert t.ranks()[0] == expect
        assert t.ranks() == (64, 48, 48)
        t.compress(1e-6, 4)
        assert t.ranks() == (4, 3, 3)
        assert np.abs(t.reconstructed() - first).max() <= 1e-13


def test_rank_known_answers(oracle_mod):
    n, vmin, vmax = (11, 11, 11), [-3, -0.1, -0.1], [3, 0.1, 0.1]
    _, V = vgrid(n, vmin, vmax)
    Vg = [v.reshape(n, order="F") for v in V]
    nrm = np.array([0.3, -0.5, 0.81])
    vn = nrm[0] * Vg[0] + nrm[1] * Vg[1] + nrm[2] * Vg[2]
    t = oracle_mod.TuckerObj.from_full(vn, 1e-6)
    assert t.ranks() == (2, 2, 2)                        # a sum of three rank-1 terms has multilinear rank 2
    assert rel_l2(t.reconstructed(), vn) <= 1e-13
    ta = oracle_mod.TuckerObj.from_full(np.abs(vn), 1e-6, 6)
    assert max(ta.ranks()) == 6                          # solver.cpp:282 caps |v.n| at rank 6
    _, Vc = vgrid(n, [-3, -3, -3], [3, 3, 3])
    vc = np.abs(nrm[0] * Vc[0] + nrm[1] * Vc[1] + nrm[2] * Vc[2]).reshape(n, order="F")
    tc = oracle_mod.TuckerObj.from_full(vc, 1e-6, 6)
    assert tc.ranks() == (6, 6, 6)
    assert 1e-4 < rel_l2(tc.reconstructed(), vc) < 5e-2  # the rank cap makes |v.n| approximate
    mx = np.exp(-0.5 * (Vg[0] ** 2 + (Vg[1] / 0.05) ** 2 + (Vg[2] / 0.05) ** 2))
    tm = oracle_mod.TuckerObj.from_full(mx, 1e-6)
    assert tm.ranks() == (1, 1, 1)                       # separable Maxwellian
    assert abs(tm.sum() - mx.sum()) <= 1e-12 * mx.sum()


def test_hadamard_and_sum(oracle_mod):
    rng = np.random.default_rng(2)
    a, b = rng.random((6, 5, 4)), rng.random((6, 5, 4))
    ta, tb = oracle_mod.TuckerObj.from_full(a), oracle_mod.TuckerObj.from_full(b)
    ta.hadamard(tb)
    assert ta.ranks() == (36, 25, 16)                    # Kronecker factors (tucker.cpp:259-300)
    assert rel_l2(ta.reconstructed(), a * b) <= 1e-13
    assert abs(ta.sum() - (a * b).sum()) <= 1e-12 * (a * b).sum()


def test_dense_truncation_equals_compress(oracle_mod):
    """Compress of a sum == truncated HOSVD of the tensor the sum represents."""
    rng = np.random.default_rng(3)
    n = (7, 6, 5)
    ax = [np.linspace(-1, 1, k) for k in n]
    smooth = lambda c: np.exp(-((ax[0][:, None, None] - c[0]) ** 2 + (ax[1][None, :, None] - c[1]) ** 2
                                + (ax[2][None, None, :] - c[2]) ** 2 + 0.3 * ax[0][:, None, None] * ax[1][None, :, None]))
    x, y = smooth((0.1, 0.2, -0.3)), smooth((-0.4, 0.0, 0.5))
    for eps in (1e-2, 1e-4, 1e-6):
        tx, ty = oracle_mod.TuckerObj.from_full(x), oracle_mod.TuckerObj.from_full(y)
        tx.axpy(0.7, ty)
        tx.compress(eps, 5)
        ref, r = tdr.truncate(x + 0.7 * y, eps, 5)
        assert tx.ranks() == tuple(r)
        assert rel_l2(tx.reconstructed(), ref) <= 1e-12


@pytest.mark.parametrize("mesh,pairs,spec", [
    ("fully_periodic_coarse.msh", [(1, 2), (3, 4), (5, 6)], {}),
    ("rectangle.msh", [(3, 4), (5, 6)], {1: ("Absorbing", False), 2: ("Free", False)}),
])
def test_update_pdf_dense_formulation(oracle_mod, mesh, pairs, spec):
    """Solver<Tucker>::_UpdatePDF in real Tucker algebra (oracle) == the dense + truncated-HOSVD
    statement of it, within the compression error."""
    m = oracle_mod.Mesh.load(mesh_path(mesh), pairs)
    n, vmin, vmax = (9, 7, 5), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    eps, dt, qm = 1e-6, 2e-3, 0.5
    _, V = vgrid(n, vmin, vmax)
    L = m.points[:, 0].max()
    dens = 1 + 0.3 * np.sin(2 * np.pi * m.tetCentroid[:, 0] / L)
    mx = np.exp(-0.5 * ((V[0] - 0.4) ** 2 + (V[1] / 0.5) ** 2 + (V[2] / 0.5) ** 2))
    f = dens[:, None] * mx[None, :]
    rng = np.random.default_rng(4)
    E = rng.standard_normal((m.nTets, 3))
    ts = oracle_mod.TuckerSim(m, n, vmin, vmax, 2.0, 1.0, eps)
    for e, (kind, _) in spec.items():
        ts.set_particle_bc(e, kind)
    ts.set_pdf(f)
    bc, _ = face_bc_arrays(m, spec)
    g = f.copy()
    for _ in range(2 if m.nTets > 500 else 3):   # the oracle's Tucker algebra costs ~20 ms per tet-step
        ts.update_pdf(dt, E)
        g, ranks = tdr.step_dense(g, m.adj, m.faceArea, m.tetVolume, m.faceNormal, bc, n, vmin, vmax, qm, E, dt, eps, max(n))
    fo = ts.get_pdf()
    assert rel_l2(g, fo) <= eps
    per_tet = np.linalg.norm(g - fo, axis=1) / np.linalg.norm(fo, axis=1)
    assert per_tet.max() <= 3 * eps
    assert rel_l2(g.sum(axis=1), fo.sum(axis=1)) <= eps
    assert np.abs(ranks - ts.ranks()).max() <= 1

"""CPU tests: the oracle against the known answers the reference's fixtures and tests give
(SURVEY.md §8c), and internal consistency of its two evaluation modes."""
import numpy as np
import pytest

from conftest import mesh_path, rel_l2

EPS0 = 8.85e-12   # src/constants.h:10
PI = 3.14159265358979323846


def test_mesh_rectangle_fine_golden(oracle_mod):
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 2), (3, 4), (5, 6)])
    assert (m.nPoints, m.nTets) == (353, 824)
    assert m.tets[0].tolist() == [309, 251, 231, 267]
    assert int((m.faceBoundary == 0).sum()) == 2604
    assert int((m.faceBoundary == 1).sum()) == 692
    assert [int((m.faceEntity == e).sum()) for e in range(1, 7)] == [14, 14, 166, 166, 166, 166]
    assert abs(m.tetVolume.sum() - 0.0025) < 1e-15
    assert abs(m.average_cell_size() - 0.0289170916933) < 1e-12
    assert (m.adj >= 0).all()                       # fully periodic: every face has a neighbour
    # outward normals: sum_f A_f n_f = 0 per tet
    assert np.abs((m.faceArea[..., None] * m.faceNormal).sum(1)).max() < 1e-17
    # adjacency is symmetric
    for t in range(m.nTets):
        for j in range(4):
            assert t in m.adj[m.adj[t, j]]
    labels = m.labels()
    assert labels[1] == ["Boundary: Periodic C", "Poisson: Periodic"]


@pytest.mark.parametrize("name,npts,ntets,tet0,vol,size", [
    ("rectangle.msh", 86, 196, [43, 6, 77, 56], 0.0025, 0.0471441161436),
    ("box_4955_tets.msh", 1211, 4955, [781, 873, 829, 885], 1.0, 0.118022197428),
])
def test_mesh_golden_counts(oracle_mod, name, npts, ntets, tet0, vol, size):
    m = oracle_mod.Mesh.load(mesh_path(name))
    assert (m.nPoints, m.nTets) == (npts, ntets)
    assert m.tets[0].tolist() == tet0
    assert abs(m.tetVolume.sum() - vol) < 1e-12
    assert abs(m.average_cell_size() - size) < 1e-11


def test_mesh_simple_unlabelled(oracle_mod):
    m = oracle_mod.Mesh.load(mesh_path("simple.msh"))
    assert (m.nPoints, m.nTets) == (8, 5)
    assert int((m.adj >= 0).sum()) == 8 and int((m.adj < 0).sum()) == 12
    assert int(m.faceBoundary.sum()) == 0           # no triangles in the file


def test_mesh_periodic_size_mismatch_raises(oracle_mod):
    with pytest.raises(RuntimeError, match="Mismatch between the sizes of the periodic planes"):
        oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 3)])


def test_poisson_known_answer_box(oracle_mod):
    """test/poisson_test.cpp:28-120 (Test 1) on box_4955_tets; values in SURVEY.md §8c."""
    m = oracle_mod.Mesh.load(mesh_path("box_4955_tets.msh"), [(5, 6)])
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Dirichlet")
    p.set_bc(2, "Dirichlet")
    p.set_bc(3, "Neumann")
    p.set_bc(4, "Neumann")
    p.initialize()
    c = m.tetCentroid
    rho = -EPS0 * np.cos(2 * PI * c[:, 1]) * (-(2 * PI * c[:, 0]) ** 2 + (2 * PI) ** 2 * c[:, 0] + 2)
    for _ in range(5):
        phi, E = p.solve(rho)
    assert p.last_error <= 2.3e-16
    phiA = c[:, 0] * (c[:, 0] - 1) * np.cos(2 * PI * c[:, 1])
    EA = np.stack([-(2 * c[:, 0] - 1) * np.cos(2 * PI * c[:, 1]),
                   2 * PI * (c[:, 0] ** 2 - c[:, 0]) * np.sin(2 * PI * c[:, 1]), 0 * c[:, 0]], 1)
    assert len(p.csr()[2]) == 23763
    assert abs(np.sqrt(((phi - phiA) ** 2).mean()) - 1.6960645569e-03) < 1e-11
    assert abs(np.sqrt(((E - EA) ** 2).sum(1).mean()) - 1.1027392059e-01) < 1e-9
    assert abs(phi[1] - 7.944633144110e-02) < 1e-11
    assert abs(phi[4954] + 1.013482891411e-01) < 1e-11


def test_poisson_known_answer_sphere(oracle_mod):
    """test/poisson_test.cpp:210-297 (Test 3) on sphere_2697_tets."""
    m = oracle_mod.Mesh.load(mesh_path("sphere_2697_tets.msh"))
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Dirichlet")
    p.initialize()
    c = m.tetCentroid
    r2 = (c ** 2).sum(1)
    for _ in range(5):
        phi, E = p.solve(-EPS0 * r2)
    assert len(p.csr()[2]) == 12623
    assert abs(np.sqrt(((phi - (r2 ** 2 - 1) / 20) ** 2).mean()) - 3.3977980013e-04) < 1e-11
    assert abs(np.sqrt(((E + c * r2[:, None] / 5) ** 2).sum(1).mean()) - 1.5673321758e-02) < 1e-10
    assert abs(phi[1] + 3.483064056422e-02) < 1e-11


def test_poisson_cg_matches_direct_solve(oracle_mod):
    """The oracle substitutes CG@2.2e-16 for the reference's default SparseLU: check against
    scipy's SuperLU on the same assembled matrix (first, uncorrected solve)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    m = oracle_mod.Mesh.load(mesh_path("sphere_2697_tets.msh"))
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Dirichlet", value=0.3)
    p.initialize()
    rp, ci, v = p.csr()
    A = sp.csr_matrix((v, ci, rp)).tocsc()
    assert abs(A - A.T).max() < 1e-12
    rho = EPS0 * np.sin(3 * m.tetCentroid[:, 0])
    # rebuild the uncorrected RHS (poisson.cpp:186, 246-274)
    rhs = -rho / EPS0 * m.tetVolume
    for t in range(m.nTets):
        for j in range(4):
            if m.faceEntity[t, j] == 1:
                d = m.faceCentroid[t, j] - m.tetCentroid[t]
                rhs[t] -= m.faceArea[t, j] / (d @ m.faceNormal[t, j]) * 0.3
    x = sla.splu(A).solve(rhs)
    y = p.solve_system(rhs)                      # the oracle's restated Eigen CG
    assert p.last_error <= 2.3e-16
    assert rel_l2(y, x) < 1e-12
    # warm start from a perturbed solution converges to the same answer
    y2 = p.solve_system(rhs, guess=x * (1 + 1e-3))
    assert rel_l2(y2, x) < 1e-12


def test_full_known_answers(oracle_mod):
    """test/test_tensors.cpp:10-18: Sum(ones 3^3) = 27 and (x+x)*(x+x) = 4."""
    L = oracle_mod
    m = L.Mesh.load(mesh_path("simple.msh"))
    s = L.Sim(m)
    sp = s.add_species([3, 3, 3], [-1, -1, -1], [1, 1, 1], 1.0, 1.0)
    s.set_pdf(sp, np.ones((m.nTets, 27)))
    # cellVolume = 1 here, so Density() = Sum()
    assert np.array_equal(s.density(sp), np.full(m.nTets, 27.0))


def _c1s(oracle_mod, fused, mesh="rectangle_fine.msh"):
    m = oracle_mod.Mesh.load(mesh_path(mesh), [(1, 2), (3, 4), (5, 6)])
    s = oracle_mod.Sim(m)
    q = 2.975e-5
    sp = s.add_species([11, 11, 11], [-3, -.1, -.1], [3, .1, .1], 1.0, q)
    s.set_maxwell(sp, 10 + 0.2 * np.sin(2 * PI * m.tetCentroid[:, 0]), 0.0)
    s.set_params(sp, 1e-4, background=-q * 10 * np.ones(m.nTets), fused=fused)
    s.begin()
    return m, s


def test_update_pdf_fused_equals_faithful(oracle_mod):
    """The single-pass CPU variant performs the same IEEE operations per element as the
    temporaries-per-operator restatement (solver.cpp:141-212): bit-identical results."""
    m, a = _c1s(oracle_mod, False)
    _, b = _c1s(oracle_mod, True)
    for it in range(3):
        a.step(it)
        b.step(it)
    assert np.array_equal(a.get_pdf(0), b.get_pdf(0))
    # delta initial condition: f(5,5,5) = n / cellVolume (particle_data.cpp:70-77)
    _, c = _c1s(oracle_mod, False)
    f0 = c.get_pdf(0)
    idx = 5 + 11 * (5 + 11 * 5)
    assert np.count_nonzero(f0) == m.nTets
    assert np.allclose(f0[:, idx] * 2.4e-4, 10 + 0.2 * np.sin(2 * PI * m.tetCentroid[:, 0]), rtol=1e-13)


def test_transport_conserves_particles(oracle_mod):
    """Periodic box, no field: sum_t V_t * Density_t is conserved by the face-flux form up to
    rounding (each face flux is computed twice with opposite normals)."""
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    s = oracle_mod.Sim(m)
    sp = s.add_species([6, 5, 4], [-1, -1, -1], [1, 1, 1], 1.0, 0.0)
    rng = np.random.default_rng(1)
    s.set_pdf(sp, rng.random((m.nTets, 120)))
    s.set_params(sp, 1e-3, fused=True)
    n0 = (s.density(sp) * m.tetVolume).sum()
    for _ in range(5):
        s.update_pdf(sp, np.zeros((m.nTets, 3)))
    n1 = (s.density(sp) * m.tetVolume).sum()
    assert abs(n1 - n0) / n0 < 1e-13


def test_sheath_constants(oracle_mod):
    """Plasma constants printed by examples/sheath.cpp:15-27."""
    kB, e, me = 1.38e-23, 1.6e-19, 9.1e-31
    T = 1 * 11604.518
    debye = np.sqrt(EPS0 * kB * T / 1e17) / e
    wp = e * np.sqrt(1e17 / (me * EPS0))
    assert abs(debye - 2.3529069315788665e-05) < 1e-19
    assert abs(wp - 1.782902734810009e10) < 1e-3


def test_poisson_known_answer_periodic_box(oracle_mod):
    """test/poisson_test.cpp:122-208 (Test 2): fully periodic box, pinned row 0 (no Dirichlet BC).
    SURVEY.md §8c: nnz 24,771, RMS phi error 1.04 (the pinned row offsets phi by -phi_A(tet 0) and the
    mesh has ~3.5 cells per wavelength), RMS E error 2.87 with |E|_rms ~ 10.9."""
    m = oracle_mod.Mesh.load(mesh_path("box_4955_tets.msh"), [(1, 2), (3, 4), (5, 6)])
    p = oracle_mod.Poisson(m)
    p.initialize()
    c = m.tetCentroid
    arg = 2 * PI * (c[:, 0] + c[:, 1] + 2 * c[:, 2])
    rho = EPS0 * 4 * PI * PI * 6 * np.sin(arg)
    for _ in range(5):
        phi, E = p.solve(rho)
    assert len(p.csr()[2]) == 24771
    assert phi[0] == 0.0                                   # pinned (poisson.cpp:128-134, 248-254)
    phiA = np.sin(arg)
    EA = np.stack([-2 * PI * np.cos(arg), -2 * PI * np.cos(arg), -4 * PI * np.cos(arg)], 1)
    assert abs(np.sqrt(((phi - phiA) ** 2).mean()) - 1.04) < 0.01
    assert abs(np.sqrt(((E - EA) ** 2).sum(1).mean()) - 2.87) < 0.01
    assert abs(np.sqrt((EA ** 2).sum(1).mean()) - 10.9) < 0.1


def test_poisson_known_answers_finer_box(oracle_mod):
    """The second mesh of the reference's convergence series (box_10021_tets): Test 1 RMS errors
    9.9208522907e-04 / 8.8083192941e-02 with nnz 48,313, Test 2 8.0e-02 / 2.27 with nnz 50,101
     — both smaller than on box_4955_tets, i.e. the discretisation converges."""
    m = oracle_mod.Mesh.load(mesh_path("box_10021_tets.msh"), [(5, 6)])
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Dirichlet")
    p.set_bc(2, "Dirichlet")
    p.set_bc(3, "Neumann")
    p.set_bc(4, "Neumann")
    p.initialize()
    c = m.tetCentroid
    rho = -EPS0 * np.cos(2 * PI * c[:, 1]) * (-(2 * PI * c[:, 0]) ** 2 + (2 * PI) ** 2 * c[:, 0] + 2)
    for _ in range(5):
        phi, E = p.solve(rho)
    phiA = c[:, 0] * (c[:, 0] - 1) * np.cos(2 * PI * c[:, 1])
    EA = np.stack([-(2 * c[:, 0] - 1) * np.cos(2 * PI * c[:, 1]),
                   2 * PI * (c[:, 0] ** 2 - c[:, 0]) * np.sin(2 * PI * c[:, 1]), 0 * c[:, 0]], 1)
    assert len(p.csr()[2]) == 48313
    assert abs(np.sqrt(((phi - phiA) ** 2).mean()) - 9.9208522907e-04) < 1e-11
    assert abs(np.sqrt(((E - EA) ** 2).sum(1).mean()) - 8.8083192941e-02) < 1e-9

    m2 = oracle_mod.Mesh.load(mesh_path("box_10021_tets.msh"), [(1, 2), (3, 4), (5, 6)])
    p2 = oracle_mod.Poisson(m2)
    p2.initialize()
    c = m2.tetCentroid
    arg = 2 * PI * (c[:, 0] + c[:, 1] + 2 * c[:, 2])
    for _ in range(5):
        phi, E = p2.solve(EPS0 * 4 * PI * PI * 6 * np.sin(arg))
    EA = np.stack([-2 * PI * np.cos(arg), -2 * PI * np.cos(arg), -4 * PI * np.cos(arg)], 1)
    assert len(p2.csr()[2]) == 50101
    assert abs(np.sqrt(((phi - np.sin(arg)) ** 2).mean()) - 8.0e-02) < 1e-3
    assert abs(np.sqrt(((E - EA) ** 2).sum(1).mean()) - 2.27) < 0.01


def test_synthetic_msh_file_through_the_oracle_reader(oracle_mod, tmp_path):
    """The C4-format MSH 2.2 file of the synthetic Kuhn box, read by the oracle's restatement of the
    reference's loader and Mesh::Reconstruct, equals the oracle mesh built from the arrays directly
    and the tables the bench uploads — the file format, the index contract and the periodic pairing
    agree across the three routes (writer -> reader, arrays, closed-form tables)."""
    from vlasovtucker_b200 import synthetic
    dims, L = (5, 4, 3), (1.0, 0.8, 0.6)
    nodes, tets, tris, ents = synthetic.kuhn_box(*dims, L)
    path = str(tmp_path / "kuhn.msh")
    synthetic.write_msh(path, nodes, tets, tris, ents)
    pairs = [(1, 2), (3, 4), (5, 6)]
    a = oracle_mod.Mesh.load(path, pairs)
    b = oracle_mod.Mesh.from_arrays(nodes, tets, tris, ents, pairs)
    mt = synthetic.periodic_kuhn_tables(*dims, L)
    assert a.nTets == b.nTets == mt.nTets == 6 * 5 * 4 * 3
    for name in ("adj", "faceEntity", "tetVolume", "faceArea", "faceNormal", "tetCentroid", "faceCentroid"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.adj, mt.nbr) and np.array_equal(a.faceEntity, mt.entity)
    assert np.array_equal(a.tetVolume, mt.volume) and np.array_equal(a.faceArea, mt.area)
    assert np.array_equal(a.faceNormal, mt.normal)

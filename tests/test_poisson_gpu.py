"""GPU parity tests of the Poisson path (PoissonSolver, src/poisson.cpp) and of the whole
time-step loop (Solver::Solve / MulticomponentSolver::Solve bodies) against the oracle.

The device CG and the oracle's restated Eigen CG both iterate to a 2.2e-16 relative residual;
their solutions agree to roughly cond(A)*eps, hence the 1e-9 bound on phi and E here, and the
north-star 1e-10 bound on the distribution function after N coupled steps.
"""
import numpy as np
import pytest

from conftest import face_bc_arrays, mesh_path, poisson_bc_arrays, rel_l2, tables_from_oracle

pytestmark = pytest.mark.gpu

EPS0 = 8.85e-12
PI = 3.14159265358979323846


@pytest.fixture(scope="module")
def vt():
    import vlasovtucker_b200 as vtb
    vtb.capi.load()
    return vtb


def _gpu_poisson(vt, m, spec, order=None):
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m, order=order))
    bc, val, ng = poisson_bc_arrays(m, spec)
    ctx.poisson_setup(bc, val, ng)
    return ctx


def test_poisson_box_dirichlet_neumann_periodic(vt, oracle_mod):
    """test/poisson_test.cpp:28-120 (Test 1)."""
    m = oracle_mod.Mesh.load(mesh_path("box_4955_tets.msh"), [(5, 6)])
    spec = {1: ("Dirichlet", 0, 0), 2: ("Dirichlet", 0, 0), 3: ("Neumann", 0, 0), 4: ("Neumann", 0, 0)}
    p = oracle_mod.Poisson(m)
    for e, (kind, v, g) in spec.items():
        p.set_bc(e, kind, v, g)
    p.initialize()
    ctx = _gpu_poisson(vt, m, spec, order=np.random.default_rng(0).permutation(m.nTets).astype(np.int32))
    c = m.tetCentroid
    rho = -EPS0 * np.cos(2 * PI * c[:, 1]) * (-(2 * PI * c[:, 0]) ** 2 + (2 * PI) ** 2 * c[:, 0] + 2)
    for r in range(5):
        phi_o, E_o = p.solve(rho)
        phi_g, E_g = ctx.poisson_solve(rho)
        it, res = ctx.poisson_stats()
        assert res <= 2.3e-16 and it > 0
        assert rel_l2(phi_g, phi_o) <= 1e-9, r
        assert rel_l2(E_g, E_o) <= 1e-9, r
    # and the known answer itself
    phiA = c[:, 0] * (c[:, 0] - 1) * np.cos(2 * PI * c[:, 1])
    assert abs(np.sqrt(((phi_g - phiA) ** 2).mean()) - 1.6960645569e-03) < 1e-10
    ctx.close()


def test_poisson_sphere_nonzero_dirichlet_and_neumann_values(vt, oracle_mod):
    m = oracle_mod.Mesh.load(mesh_path("sphere_2697_tets.msh"))
    spec = {1: ("Dirichlet", 0.25, 0)}
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Dirichlet", 0.25, 0)
    p.initialize()
    ctx = _gpu_poisson(vt, m, spec)
    r2 = (m.tetCentroid ** 2).sum(1)
    for r in range(3):
        phi_o, E_o = p.solve(-EPS0 * r2)
        phi_g, E_g = ctx.poisson_solve(-EPS0 * r2)
        assert rel_l2(phi_g, phi_o) <= 1e-9
        assert rel_l2(E_g, E_o) <= 1e-9
    ctx.close()


def test_poisson_neumann_wall_value_update(vt, oracle_mod):
    """Neumann data change between solves (solver.cpp:120-132): only the RHS moves."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    p = oracle_mod.Poisson(m)
    p.set_bc(1, "Neumann", 0, 0.0)
    p.set_bc(2, "Dirichlet", 0.0, 0)
    p.initialize()
    spec = {1: ("Neumann", 0, 0.0), 2: ("Dirichlet", 0.0, 0)}
    ctx = _gpu_poisson(vt, m, spec)
    rho = EPS0 * 1e3 * np.sin(2 * PI * m.tetCentroid[:, 0])
    for k, g in enumerate([0.0, 150.0, -40.0]):
        p.set_bc(1, "Neumann", 0, g)
        spec[1] = ("Neumann", 0, g)
        _, val, ng = poisson_bc_arrays(m, spec)
        ctx.poisson_update_bc_values(val, ng)
        phi_o, E_o = p.solve(rho)
        phi_g, E_g = ctx.poisson_solve(rho)
        assert rel_l2(phi_g, phi_o) <= 1e-9, k
        assert rel_l2(E_g, E_o) <= 1e-9, k
    ctx.close()


@pytest.mark.parametrize("mesh", ["fully_periodic_coarse.msh", "rectangle.msh"])
def test_poisson_fully_periodic_pinned_row(vt, oracle_mod, mesh):
    """test/poisson_test.cpp:122-208 (Test 2) setting: no Dirichlet BC -> row 0 pinned."""
    m = oracle_mod.Mesh.load(mesh_path(mesh), [(1, 2), (3, 4), (5, 6)])
    p = oracle_mod.Poisson(m)
    p.initialize()
    ctx = _gpu_poisson(vt, m, {}, order=np.random.default_rng(1).permutation(m.nTets).astype(np.int32))
    rho = EPS0 * np.sin(2 * PI * m.tetCentroid[:, 0] / m.tetCentroid[:, 0].max())
    for r in range(4):
        phi_o, E_o = p.solve(rho)
        phi_g, E_g = ctx.poisson_solve(rho)
        assert phi_g[0] == 0.0
        assert rel_l2(phi_g, phi_o) <= 1e-9
        assert rel_l2(E_g, E_o) <= 1e-9
    ctx.close()


def test_poisson_zero_rhs_gives_zero(vt, oracle_mod):
    """Eigen's CG returns x = 0 for a zero right-hand side (ConjugateGradient.h:48-54)."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(1, 2), (3, 4), (5, 6)])
    ctx = _gpu_poisson(vt, m, {})
    phi, E = ctx.poisson_solve(np.zeros(m.nTets))
    assert not phi.any() and not E.any()
    ctx.close()


def test_missing_field_bc_is_an_error(vt, oracle_mod):
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    bc, val, ng = poisson_bc_arrays(m, {})
    with pytest.raises(RuntimeError, match="without a field BC"):
        ctx.poisson_setup(bc, val, ng)
    ctx.close()


def _coupled_single(vt, oracle_mod, mesh, steps, q, n=(11, 11, 11)):
    """Solver<Full>::Solve loop body (solver.cpp:91-133), stable variant C1s of SURVEY.md §8d."""
    m = oracle_mod.Mesh.load(mesh_path(mesh), [(1, 2), (3, 4), (5, 6)])
    vmin, vmax = [-3, -.1, -.1], [3, .1, .1]
    L = m.tetCentroid[:, 0].max() + m.tetCentroid[:, 0].min()
    dens = 10 + 0.2 * np.sin(2 * PI * m.tetCentroid[:, 0] / L)
    bg = -q * 10 * np.ones(m.nTets)
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 1.0, q)
    s.set_maxwell(sp, dens, 0.0)
    s.set_params(sp, 1e-4, background=bg, fused=True)
    s.begin()

    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m, order=np.random.default_rng(2).permutation(m.nTets).astype(np.int32)))
    g = ctx.species_create(n, vmin, vmax, 1.0, q)
    bc, col = face_bc_arrays(m, {})
    ctx.set_face_bc(g, bc, col)
    ctx.set_maxwell(g, dens, 0.0)
    qb, val, ng = poisson_bc_arrays(m, {})
    ctx.poisson_setup(qb, val, ng)
    for it in range(steps):
        s.step(it)
        ctx.charge_density([g], bg)          # solver.cpp:98-105
        ctx.poisson_solve(download=False)    # solver.cpp:108-110
        ctx.step_full(g, 1e-4)               # solver.cpp:115
    return m, s, sp, ctx, g


# C1s of SURVEY.md §8d is rectangle_fine.msh for N = 200 iterations
@pytest.mark.parametrize("mesh,steps", [("fully_periodic_coarse.msh", 30), ("rectangle.msh", 30), ("rectangle_fine.msh", 200)])
def test_coupled_loop_parity(vt, oracle_mod, mesh, steps):
    m, s, sp, ctx, g = _coupled_single(vt, oracle_mod, mesh, steps, 2.975e-5)
    rho_o, phi_o, E_o = s.fields(sp)
    rho_g, phi_g, E_g = ctx.field_get()
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= 1e-10
    assert rel_l2(phi_g, phi_o) <= 1e-8
    assert rel_l2(E_g, E_o) <= 1e-8
    assert rel_l2(ctx.density(g), s.density(sp)) <= 1e-10
    ctx.close()


def test_c1_as_committed_first_step(vt, oracle_mod):
    """examples/oscillations.cpp exactly (q = 10, SI epsilon0, no background): one step."""
    m, s, sp, ctx, g = None, None, None, None, None
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (11, 11, 11), [-3, -.1, -.1], [3, .1, .1]
    dens = 10 + 0.2 * np.sin(1 * m.tetCentroid[:, 0] * (2 * PI))
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 1.0, 10.0)
    s.set_maxwell(sp, dens, 0.0)
    s.set_params(sp, 1e-4, fused=True)
    s.begin()
    s.step(0)
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create(n, vmin, vmax, 1.0, 10.0)
    bc, col = face_bc_arrays(m, {})
    ctx.set_face_bc(g, bc, col)
    ctx.set_maxwell(g, dens, 0.0)
    qb, val, ng = poisson_bc_arrays(m, {})
    ctx.poisson_setup(qb, val, ng)
    ctx.charge_density([g])
    ctx.poisson_solve(download=False)
    ctx.step_full(g, 1e-4)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= 1e-10
    ctx.close()


def test_sheath_two_species_loop(vt, oracle_mod):
    """examples/sheath.cpp through MulticomponentSolver::Solve's loop body
    (multicomponent_solver.cpp:55-126): electrons every step, ions every 10th with 10x dt,
    absorbing charged wall (entity 1), free Dirichlet wall (entity 2), shared Poisson solve."""
    kB, e, me, mi, eV = 1.38e-23, 1.6e-19, 9.1e-31, 1.66e-27, 11604.518
    Te, Ti, dens = 1 * eV, 400.0, 1e17
    debye = np.sqrt(EPS0 * kB * Te / dens) / e
    wp = e * np.sqrt(dens / (me * EPS0))
    dt = 1e-4 * 2 * PI / wp
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)], scale=22 * debye)
    maxVE = np.sqrt(-np.log(1e-6) * 2 * kB * Te / me)
    maxVI = np.sqrt(-np.log(1e-6) * 2 * kB * Ti / mi)
    grids = [((50, 5, 5), [-4 * maxVE, -maxVE, -maxVE], [4 * maxVE, maxVE, maxVE], me, -e, Te, 1),
             ((50, 5, 5), [-4 * maxVI, -maxVI, -maxVI], [4 * maxVI, maxVI, maxVI], mi, e, Ti, 10)]
    s = oracle_mod.Sim(m)
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    gs = []
    for (n, vmin, vmax, mass, q, T, mult) in grids:
        sp = s.add_species(n, vmin, vmax, mass, q, mult)
        s.set_maxwell(sp, np.full(m.nTets, dens), T)
        s.set_params(sp, dt * mult, fused=True)
        s.set_particle_bc(sp, 1, "Absorbing", True)
        s.set_particle_bc(sp, 2, "Free", False)
        g = ctx.species_create(n, vmin, vmax, mass, q)
        bc, col = face_bc_arrays(m, {1: ("Absorbing", True), 2: ("Free", False)})
        ctx.set_face_bc(g, bc, col)
        ctx.set_maxwell(g, np.full(m.nTets, dens), T)
        gs.append((g, mult))
    s.set_field_bc_charge(0, 1, 0.0)
    s.set_field_bc_potential(0, 2, 0.0)
    s.begin()
    spec = {1: ("Neumann", 0, 0.0), 2: ("Dirichlet", 0.0, 0)}
    qb, val, ng = poisson_bc_arrays(m, spec)
    ctx.poisson_setup(qb, val, ng)
    area = s.wall_area(0, 1)
    steps = 200   # C3 of SURVEY.md §8d: N = 200 iterations (the ions take 20 sub-cycled steps)
    for it in range(steps):
        s.step(it)
        ctx.charge_density([g for g, _ in gs])
        ctx.poisson_solve(download=False)
        for g, mult in gs:
            if it % mult == 0:
                ctx.step_full(g, dt * mult)
        # merged wall charge -> Neumann value sigma/(2 eps0) (multicomponent_solver.cpp:99-126, solver.cpp:56)
        Q = sum(ctx.wall_charge(g, 1) for g, _ in gs)
        spec[1] = ("Neumann", 0, (Q / area) / (2 * EPS0))
        _, val, ng = poisson_bc_arrays(m, spec)
        ctx.poisson_update_bc_values(val, ng)
    for k, (g, _) in enumerate(gs):
        assert rel_l2(ctx.get_pdf(g), s.get_pdf(k)) <= 1e-10, k
        qo, qg = s.wall_charge(k, 1), ctx.wall_charge(g, 1)
        assert qo != 0 and abs(qg - qo) <= 1e-9 * abs(qo)
    rho_o, phi_o, E_o = s.fields(0)
    rho_g, phi_g, E_g = ctx.field_get()
    assert rel_l2(phi_g, phi_o) <= 1e-7
    assert rel_l2(E_g, E_o) <= 1e-7
    ctx.close()

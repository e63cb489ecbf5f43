"""Halo path on ONE GPU: N "virtual ranks" = N contexts on the same device whose ghost rows are
wired straight to each other's buffers (vt_halo_attach_local) — the kernels, push lists and the
device-side barrier are exactly those of the one-process-per-GPU run (tests/test_multigpu_gpu.py
needs >= 2 GPUs and is skipped on a one-GPU box; this file is not).  SURVEY.md §4 (ii).

The partitioned update must reproduce the single-context update bit for bit: per-tet arithmetic is
identical, only the memory the neighbour values come from differs."""
import numpy as np
import pytest

from conftest import face_bc_arrays, mesh_path, rel_l2, tables_from_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vt():
    import vlasovtucker_b200 as vtb
    vtb.capi.load()
    return vtb


def _single(vt, mt, n, vmin, vmax, f0, E, dt, steps, chunk, variant, bc=None, col=None):
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 2.0)
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vt.PBC["Periodic"], np.uint8) if bc is None else bc, col)
    ctx.set_pdf(g, f0)
    ctx.field_set(E)
    ctx.step_config(chunk_planes=chunk, variant=variant)
    for _ in range(steps):
        ctx.step_full(g, dt)
    out = ctx.get_pdf(g), ctx.density(g)
    ctx.close()
    return out


@pytest.mark.parametrize("world,mode,n,chunk,variant", [
    (2, "block", (16, 8, 8), 2, 0),          # register-staged kernel, reference arithmetic
    (2, "rcb", (16, 8, 8), 2, 2),            # upwind select
    (3, "rcb", (16, 8, 8), None, 18),        # bulk-copy pipeline, 16 warps, three ranks
    (4, "block", (16, 8, 8), 2, 18),
    (2, "block", (32, 32, 4), None, 64),     # the bench kernel: boundary/halo launch + interior launch
    (4, "rcb", (32, 32, 4), 2, 64),          # ... chunked items
    (2, "rcb", (32, 32, 5), 2, 64 | 256),    # ... tet-major item order
    (2, "block", (32, 32, 4), None, 64 | 128),   # whole neighbour planes by bulk copy
    (2, "rcb", (64, 16, 3), None, 64),
])
def test_virtual_ranks_full_bit_identical(vt, world, mode, n, chunk, variant):
    from vlasovtucker_b200 import multigpu, partition as part, synthetic
    dims = (6, 4, 4)
    mt = synthetic.periodic_kuhn_tables(*dims, (1.5, 1.0, 1.0))
    vmin, vmax = [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    rng = np.random.default_rng(0)
    N = n[0] * n[1] * n[2]
    f0 = rng.random((mt.nTets, N))
    E = rng.standard_normal((mt.nTets, 3))
    steps, dt = 4, 1e-3
    if mode == "rcb":
        owner = part.rcb_owner(mt.tetCentroid, world)
    else:
        owner = part.block_owner(dims, part.rank_grid(world) if world != 4 else (2, 2, 1))
    vr = multigpu.VirtualRanks(mt, owner, world)
    ids = vr.species_create(n, vmin, vmax, 1.0, 2.0)
    for r, ctx in enumerate(vr.ctxs):
        ctx.set_pdf(ids[r], f0[vr.lps[r].owned])
        ctx.field_set(E[vr.lps[r].owned])
        ctx.step_config(chunk_planes=chunk, variant=variant)
    vr.fill_ghosts(ids)
    for _ in range(steps):
        vr.step_full(ids, dt)
    vr.sync()
    full = vr.gather([ctx.get_pdf(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    dens = vr.gather([ctx.density(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    vr.close()
    ref, dref = _single(vt, mt, n, vmin, vmax, f0, E, dt, steps, chunk, variant)
    assert np.array_equal(full, ref), f"max|diff| = {np.abs(full - ref).max():.3e}"
    assert np.array_equal(dens, dref)


def test_virtual_ranks_walls_unstructured(vt, oracle_mod):
    """Unstructured mesh with an absorbing, charge-collecting wall and a free wall, split over three
    virtual ranks: boundary-condition tets and halo tets share the generic instance of the kernel.
    State bit-identical to one context; wall charge (atomic accumulation) to rounding."""
    from vlasovtucker_b200 import multigpu, partition as part
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    mt = tables_from_oracle(m)
    n, vmin, vmax = (32, 32, 4), [-4.0, -1.0, -1.0], [4.0, 1.0, 1.0]
    rng = np.random.default_rng(3)
    f0 = rng.random((mt.nTets, n[0] * n[1] * n[2]))
    E = rng.standard_normal((mt.nTets, 3))
    bc, col = face_bc_arrays(m, {1: ("Absorbing", True), 2: ("Free", False)})
    world, steps, dt = 3, 4, 2e-4
    owner = part.rcb_owner(mt.tetCentroid, world)
    vr = multigpu.VirtualRanks(mt, owner, world)
    ids = vr.species_create(n, vmin, vmax, 1.0, 2.0, bc_type=bc, collect=col)
    for r, ctx in enumerate(vr.ctxs):
        ctx.set_pdf(ids[r], f0[vr.lps[r].owned])
        ctx.field_set(E[vr.lps[r].owned])
    vr.fill_ghosts(ids)
    for _ in range(steps):
        vr.step_full(ids, dt)
    vr.sync()
    full = vr.gather([ctx.get_pdf(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    q = sum(ctx.wall_charge(ids[r], 1) for r, ctx in enumerate(vr.ctxs))
    vr.close()
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 2.0)
    ctx.set_face_bc(g, bc, col)
    ctx.set_pdf(g, f0)
    ctx.field_set(E)
    for _ in range(steps):
        ctx.step_full(g, dt)
    ref = ctx.get_pdf(g)
    qref = ctx.wall_charge(g, 1)
    ctx.close()
    assert np.array_equal(full, ref)
    assert qref != 0 and abs(q - qref) <= 1e-12 * abs(qref)


@pytest.mark.parametrize("n,cap,kernel", [((12, 10, 8), 0, "general"), ((34, 12, 10), 8, "slab")])
def test_virtual_ranks_tucker_bit_identical(vt, n, cap, kernel):
    """Tucker species: the step kernel (the general one, and the slab-streaming one for grids above 32 nodes
    per axis) mirrors every boundary tet's new core, factors and ranks into the peers' ghost slots."""
    from vlasovtucker_b200 import multigpu, partition as part, synthetic
    dims = (4, 2, 2)
    mt = synthetic.periodic_kuhn_tables(*dims, (2.0, 1.0, 1.0))
    vmin, vmax = [-3.0, -2.0, -2.0], [3.0, 2.0, 2.0]
    ax = [np.linspace(vmin[k], vmax[k], n[k]) for k in range(3)]
    V0, V1, V2 = np.meshgrid(*ax, indexing="ij")
    x = mt.tetCentroid[:, 0]
    f0 = np.stack([((1 + 0.3 * np.sin(3.14159 * xx)) * np.exp(-((V0 - 0.5 * np.cos(3.14159 * xx)) ** 2 + V1 ** 2 + V2 ** 2) / 2)
                    ).ravel(order="F") for xx in x])
    E = np.random.default_rng(1).standard_normal((mt.nTets, 3)) * 0.3
    world, steps, dt, eps = 2, 3, 2e-3, 1e-6
    owner = part.rcb_owner(mt.tetCentroid, world)
    vr = multigpu.VirtualRanks(mt, owner, world)
    ids = vr.species_create(n, vmin, vmax, 1.0, 1.5, tucker=(eps, cap))
    for r, ctx in enumerate(vr.ctxs):
        ctx.tucker_set_pdf(ids[r], f0[vr.lps[r].owned])
        ctx.field_set(E[vr.lps[r].owned])
    vr.fill_ghosts(ids)
    for _ in range(steps):
        vr.step_tucker(ids, dt)
    vr.sync()
    assert all(ctx.tucker_last_kernel(ids[r]) == kernel for r, ctx in enumerate(vr.ctxs))
    full = vr.gather([ctx.tucker_get_pdf(ids[r], n[0] * n[1] * n[2]) for r, ctx in enumerate(vr.ctxs)])
    ranks = vr.gather([ctx.tucker_ranks(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    vr.close()
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 1.5)
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vt.PBC["Periodic"], np.uint8))
    ctx.tucker_enable(g, eps, cap)
    ctx.tucker_set_pdf(g, f0)
    ctx.field_set(E)
    for _ in range(steps):
        ctx.step_tucker(g, dt)
    ref = ctx.tucker_get_pdf(g, n[0] * n[1] * n[2])
    rref = ctx.tucker_ranks(g)
    ctx.close()
    assert np.array_equal(ranks, rref)
    assert np.array_equal(full, ref), f"rel diff {rel_l2(full, ref):.3e}"


# ---- partitioned Poisson solve (rows partitioned like the tets, ghost values by peer stores, the CG
# dot products summed over the ranks inside the solve kernel) against the one-context solve and the oracle

EPS0 = 8.85e-12
PI = 3.14159265358979323846


@pytest.mark.parametrize("mesh,pairs,spec,world", [
    ("box_4955_tets.msh", [(5, 6)], {1: ("Dirichlet", 0, 0), 2: ("Dirichlet", 0, 0), 3: ("Neumann", 0, 0), 4: ("Neumann", 0, 0)}, 3),
    ("fully_periodic_coarse.msh", [(1, 2), (3, 4), (5, 6)], {}, 2),          # pinned row 0 on one rank
    ("rectangle_fine.msh", [(3, 4), (5, 6)], {1: ("Neumann", 0, 75.0), 2: ("Dirichlet", 0.5, 0)}, 4),
])
@pytest.mark.timeout(300)
def test_virtual_ranks_poisson(vt, oracle_mod, monkeypatch, mesh, pairs, spec, world):
    from conftest import poisson_bc_arrays
    monkeypatch.setenv("VT_COMM_TIMEOUT_MS", "5000")   # a rank that never arrives fails the test in seconds
    from vlasovtucker_b200 import multigpu, partition as part
    m = oracle_mod.Mesh.load(mesh_path(mesh), pairs)
    mt = tables_from_oracle(m)
    bc, val, ng = poisson_bc_arrays(m, spec)
    c = m.tetCentroid
    L = c.max(0)
    rho = -EPS0 * np.cos(2 * PI * c[:, 1] / L[1]) * (1 + np.sin(2 * PI * c[:, 0] / L[0]))
    p = oracle_mod.Poisson(m)
    for e, (kind, v, g) in spec.items():
        p.set_bc(e, kind, v, g)
    p.initialize()
    one = vt.Context(0)
    one.mesh_upload(mt)
    one.poisson_setup(bc, val, ng)
    owner = part.rcb_owner(mt.tetCentroid, world)
    vr = multigpu.VirtualRanks(mt, owner, world)
    vr.poisson_setup(bc, val, ng)
    for r in range(3):       # first call: two solves (uncorrected, corrected); later calls: warm start
        phi_o, E_o = p.solve(rho * (1 + 0.1 * r))
        phi_1, E_1 = one.poisson_solve(rho * (1 + 0.1 * r))
        vr.poisson_solve(rho * (1 + 0.1 * r))
        vr.sync()
        got = [ctx.field_get() for ctx in vr.ctxs]
        phi_p = vr.gather([g[1] for g in got])
        E_p = vr.gather([g[2] for g in got])
        its = [ctx.poisson_stats() for ctx in vr.ctxs]
        assert len({i for i, _ in its}) == 1            # every rank took the same number of iterations
        assert all(res <= 2.3e-16 for _, res in its)
        # same algorithm, other summation order: cond(A) * eps apart
        assert rel_l2(phi_p, phi_1) <= 1e-10, (r, rel_l2(phi_p, phi_1))
        assert rel_l2(E_p, E_1) <= 1e-10, (r, rel_l2(E_p, E_1))
        assert rel_l2(phi_p, phi_o) <= 1e-9
        assert rel_l2(E_p, E_o) <= 1e-9
    vr.close()
    one.close()


@pytest.mark.timeout(300)
def test_virtual_ranks_coupled_loop(vt, oracle_mod, monkeypatch):
    """The whole loop body of Solver::Solve (solver.cpp:91-133) on three virtual ranks — charge density,
    partitioned Poisson solve, partitioned step with fused halo — against one context and the oracle
    (config C1s)."""
    from conftest import poisson_bc_arrays
    from vlasovtucker_b200 import multigpu, partition as part
    monkeypatch.setenv("VT_COMM_TIMEOUT_MS", "5000")
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(1, 2), (3, 4), (5, 6)])
    mt = tables_from_oracle(m)
    n, vmin, vmax, q, dt, steps, world = (11, 11, 11), [-3, -.1, -.1], [3, .1, .1], 2.975e-5, 1e-4, 20, 3
    Lx = m.tetCentroid[:, 0].max() + m.tetCentroid[:, 0].min()
    dens = 10 + 0.2 * np.sin(2 * PI * m.tetCentroid[:, 0] / Lx)
    bg = -q * 10 * np.ones(m.nTets)
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 1.0, q)
    s.set_maxwell(sp, dens, 0.0)
    s.set_params(sp, dt, background=bg, fused=True)
    s.begin()
    qb, val, ng = poisson_bc_arrays(m, {})
    owner = part.rcb_owner(mt.tetCentroid, world)
    vr = multigpu.VirtualRanks(mt, owner, world)
    ids = vr.species_create(n, vmin, vmax, 1.0, q)
    for r, ctx in enumerate(vr.ctxs):
        ctx.set_maxwell(ids[r], dens[vr.lps[r].owned], 0.0)
    vr.fill_ghosts(ids)
    vr.poisson_setup(qb, val, ng)
    for it in range(steps):
        s.step(it)
        for r, ctx in enumerate(vr.ctxs):
            ctx.charge_density([ids[r]], bg[vr.lps[r].owned])
        vr.poisson_solve()
        vr.step_full(ids, dt)
    vr.sync()
    f = vr.gather([ctx.get_pdf(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    phi = vr.gather([ctx.field_get()[1] for ctx in vr.ctxs])
    vr.close()
    assert rel_l2(f, s.get_pdf(sp)) <= 1e-10
    assert rel_l2(phi, s.fields(sp)[1]) <= 1e-8

"""GPU parity tests of the full-format hot path (K1) through the C ABI against the oracle.

Tolerance: relative L2 <= 1e-10 on the state (BASELINE.json north_star); the observed
differences are FMA-contraction level (~1e-16 per step).
"""
import numpy as np
import pytest

from conftest import face_bc_arrays, mesh_path, rel_l2, tables_from_oracle

pytestmark = pytest.mark.gpu

PI = 3.14159265358979323846
TOL = 1e-10


@pytest.fixture(scope="module")
def vt():
    import vlasovtucker_b200 as vtb
    vtb.capi.load()
    return vtb


def _smooth_state(m, n, vmin, vmax, seed=0):
    """Smooth positive f(x,v) with some structure in every direction."""
    rng = np.random.default_rng(seed)
    ax = [np.linspace(vmin[k], vmax[k], n[k]) for k in range(3)]
    V0, V1, V2 = np.meshgrid(*ax, indexing="ij")
    w = [0.35 * (vmax[k] - vmin[k]) for k in range(3)]
    base = np.exp(-(V0 / w[0]) ** 2 - (V1 / w[1]) ** 2 - (V2 / w[2]) ** 2)
    c = m.tetCentroid
    amp = 1 + 0.3 * np.sin(2 * PI * c[:, 0] / max(c[:, 0].max(), 1e-30)) + 0.05 * rng.random(m.nTets)
    f = amp[:, None] * base.ravel(order="F")[None, :]
    f *= 1 + 0.01 * rng.standard_normal(f.shape)
    return f


def _run_pair(vt, oracle_mod, m, n, vmin, vmax, mass, charge, f0, E, dt, steps, ext=(0, 0, 0), bc_spec=None,
              order=None, chunk=None, brick=None, variant=None, source=None):
    bc_spec = bc_spec or {}
    # oracle
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, mass, charge)
    s.set_pdf(sp, f0)
    s.set_params(sp, dt, ext=ext, fused=True)
    for e, (kind, collect) in bc_spec.items():
        s.set_particle_bc(sp, e, kind, collect, source_pdf=source if kind == "Source" else None)
    # GPU
    ctx = vt.Context(0)
    mt = tables_from_oracle(m, order=order)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, mass, charge)
    bc, col = face_bc_arrays(m, bc_spec)
    src_id = None
    if source is not None:
        ctx.set_source_pdfs(g, source[None, :])
        src_id = np.where(bc == vt.PBC["Source"], 0, -1).astype(np.int32)
    ctx.set_face_bc(g, bc, col, src_id)
    ctx.set_pdf(g, f0)
    if chunk is not None or brick is not None or variant is not None:
        ctx.step_config(chunk_planes=chunk, brick_tets=brick, variant=variant)
    ctx.field_set(E)
    return s, sp, ctx, g


@pytest.mark.parametrize("n,chunk,brick,variant", [
    # variant bits: 1 no shuffles, 2 upwind-select arithmetic, 16 persistent bulk-copy (TMA) pipeline, 32 eight
    # consumer warps, 128 whole neighbour planes, 256 tet-major items; 0 = register-staged kernel, reference shape
    ((11, 11, 11), None, None, None),     # C1 grid: odd n0 -> scalar path
    ((11, 11, 11), 3, 100, 2),            # ragged chunks and bricks, upwind arithmetic
    ((8, 6, 4), None, None, None),        # double2 path without shuffles
    ((8, 6, 4), 3, 50, 2),
    ((8, 6, 4), 3, 50, 18),               # TMA, ragged chunk
    ((16, 8, 8), None, None, None),       # shuffle path
    ((16, 8, 8), 2, 128, 1),              # shuffles disabled
    ((16, 8, 8), 2, 128, 2),              # upwind
    ((16, 8, 8), 2, 128, 16),             # TMA
    ((32, 4, 6), 3, 0, None),
    ((32, 32, 4), 2, 0, 2),               # the bench plane size
    ((48, 48, 3), None, 0, 18),           # TMA, 3 columns per thread
    ((50, 5, 5), None, None, None),       # sheath grid
    ((32, 32, 4), None, 0, None),         # library default at the bench plane size: bulk-copy pipeline, 8 warps x 2 columns
    ((32, 32, 4), 2, 0, 18),              # bulk-copy pipeline, 16 consumer warps, 2-plane items
    ((32, 32, 5), 1, 64, 50),             # single-plane items through the queue, bricks of 64 tets
    ((32, 32, 4), 2, 0, 0),               # register-staged kernel, reference expression shape
    ((64, 64, 3), None, 0, 18),           # 4 columns per consumer thread, 32 KiB planes (ring depth 1x... 4)
    ((24, 22, 4), 3, 0, 16),              # plane not a multiple of the consumer count, reference arithmetic
    ((32, 32, 4), None, 0, 64 | 128),     # default kernel with whole neighbour planes by bulk copy (round-1 path)
    ((32, 32, 7), 3, 64, 50),             # consumer-side neighbour loads, ragged chunks, bricks
    ((32, 32, 8), 2, 0, 50 | 256),        # ... work items in tet-major order, chunk fastest
    ((32, 32, 3), 1, 0, 50 | 256),        # ... single-plane items: the neighbour prefetch runs across 3 items
    ((64, 16, 5), None, 0, None),         # eight-warp layout on a non-square plane (no compile-time shape)
    ((16, 64, 6), 4, 0, 64 | 256),
    ((64, 16, 5), None, 0, 64 | 128),
])
def test_update_pdf_periodic_parity(vt, oracle_mod, n, chunk, brick, variant):
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 2), (3, 4), (5, 6)])
    vmin, vmax = [-3, -0.1, -0.2], [3, 0.1, 0.2]
    f0 = _smooth_state(m, n, vmin, vmax)
    rng = np.random.default_rng(5)
    E = rng.standard_normal((m.nTets, 3)) * 0.5
    ext = (0.1, -0.2, 0.05)
    order = rng.permutation(m.nTets).astype(np.int32)
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 1.0, 10.0, f0, E, 1e-4, 3, ext=ext, order=order,
                              chunk=chunk, brick=brick, variant=variant)
    for _ in range(3):
        s.update_pdf(sp, E)
        ctx.step_full(g, 1e-4, ext)
    fo, fg = s.get_pdf(sp), ctx.get_pdf(g)
    assert rel_l2(fg, fo) <= TOL
    assert rel_l2(ctx.density(g), s.density(sp)) <= TOL
    assert np.abs(ctx.velocity(g) - s.velocity(sp)).max() <= 1e-9 * np.abs(s.velocity(sp)).max()
    ctx.close()


@pytest.mark.parametrize("vmin,vmax", [([0.5, -0.1, -0.3], [3, 0.4, -0.05]), ([-3, -2, 1e-3], [-1, 2, 4]),
                                       ([-1e-9, -1e-9, -1e-9], [1e-9, 1e-9, 1e-9]), ([-2e6, -2e6, -2e6], [2e6, 2e6, 2e6])])
def test_inflow_only_neighbour_loads_on_lopsided_grids(vt, oracle_mod, vmin, vmax):
    """The default kernel at 32x32 planes loads a neighbour value only where v.n <= 0 (single-precision
    predicate with a guard band).  Grids that do not straddle v = 0, and very small / large velocity
    scales, move the inflow boundary to the edges of the planes or out of them."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(1, 2), (3, 4), (5, 6)])
    n = (32, 32, 6)
    f0 = _smooth_state(m, n, vmin, vmax, seed=2)
    scale = max(abs(x) for x in vmin + vmax)
    h = m.tetVolume.min() ** (1.0 / 3.0)
    dt = 0.05 * h / scale
    E = np.zeros((m.nTets, 3))
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 1.0, 1.0, f0, E, dt, 3,
                              order=np.random.default_rng(1).permutation(m.nTets).astype(np.int32))
    for _ in range(3):
        s.update_pdf(sp, E)
        ctx.step_full(g, dt)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    # and bit-identical to the same kernel fed with whole neighbour planes
    ctx2 = vt.Context(0)
    ctx2.mesh_upload(tables_from_oracle(m))
    g2 = ctx2.species_create(n, vmin, vmax, 1.0, 1.0)
    bc, col = face_bc_arrays(m, {})
    ctx2.set_face_bc(g2, bc, col)
    ctx2.set_pdf(g2, f0)
    ctx2.step_config(variant=64 | 128)
    ctx2.field_set(E)
    for _ in range(3):
        ctx2.step_full(g2, dt)
    assert np.array_equal(ctx2.get_pdf(g2), ctx.get_pdf(g))
    ctx.close()
    ctx2.close()


def test_c1_delta_one_step(vt, oracle_mod):
    """config C1 (examples/oscillations.cpp) as committed: parity is meaningful at N = 1 only
    (the case overflows after 7 steps); field from the oracle's Poisson solve."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (11, 11, 11), [-3, -.1, -.1], [3, .1, .1]
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 1.0, 10.0)
    dens = 10 + 0.2 * np.sin(1 * m.tetCentroid[:, 0] * (2 * PI))
    s.set_maxwell(sp, dens, 0.0)
    s.set_params(sp, 1e-4)
    f0 = s.get_pdf(sp)
    s.begin()
    s.step(0)
    _, _, E = s.fields(sp)

    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create(n, vmin, vmax, 1.0, 10.0)
    bc, col = face_bc_arrays(m, {})
    ctx.set_face_bc(g, bc, col)
    ctx.set_maxwell(g, dens, 0.0)
    assert np.array_equal(ctx.get_pdf(g), f0)          # same initial condition, bit for bit
    ctx.field_set(E)
    ctx.step_full(g, 1e-4)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    ctx.close()


def test_maxwellian_init_bit_exact(vt, oracle_mod):
    """SetMaxwellPDF (particle_data.cpp:23-66): same exp(), same summation order."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(3, 4), (5, 6)])
    n, vmin, vmax = (50, 5, 5), [-8e6, -2e6, -2e6], [8e6, 2e6, 2e6]
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 9.1e-31, -1.6e-19)
    dens = 1e17 * (1 + 0.01 * m.tetCentroid[:, 0])
    s.set_maxwell(sp, dens, 11604.518, (1e5, 0, 0))
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create(n, vmin, vmax, 9.1e-31, -1.6e-19)
    ctx.set_maxwell(g, dens, 11604.518, (1e5, 0, 0))
    assert np.array_equal(ctx.get_pdf(g), s.get_pdf(sp))
    assert rel_l2(ctx.density(g), s.density(sp)) <= 1e-14
    ctx.close()


@pytest.mark.parametrize("n,variant", [((50, 5, 5), None), ((32, 16, 5), 18), ((32, 32, 4), 50), ((32, 32, 4), 50 | 128),
                                       ((64, 16, 3), None)])
def test_wall_bcs_and_charge(vt, oracle_mod, n, variant):
    """Sheath-style boundaries (examples/sheath.cpp:100-110): entity 1 Absorbing+collectCharge,
    entity 2 Free, {3,4},{5,6} periodic; wall charge as solver.cpp:171-178."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    vmin, vmax = [-4, -1, -1], [4, 1, 1]
    f0 = _smooth_state(m, n, vmin, vmax, seed=3)
    rng = np.random.default_rng(7)
    E = rng.standard_normal((m.nTets, 3))
    spec = {1: ("Absorbing", True), 2: ("Free", False)}
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 2.0, -3.0, f0, E, 2e-4, 4, bc_spec=spec, variant=variant)
    s.init_wall()
    for _ in range(4):
        s.update_pdf(sp, E)
        ctx.step_full(g, 2e-4)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    qo, qg = s.wall_charge(sp, 1), ctx.wall_charge(g, 1)
    assert qo != 0.0
    assert abs(qg - qo) <= 1e-10 * abs(qo)
    assert ctx.wall_charge(g, 2) == 0.0
    ctx.close()


@pytest.mark.parametrize("n,variant", [((8, 6, 4), None), ((32, 16, 4), 18), ((32, 32, 4), None), ((32, 32, 4), 50 | 128)])
def test_source_bc(vt, oracle_mod, n, variant):
    """Source faces use ParticleBC::sourcePDF in place of the neighbour (solver.cpp:334-339)."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    vmin, vmax = [-4, -1, -1], [4, 1, 1]
    f0 = _smooth_state(m, n, vmin, vmax, seed=4)
    src = 2.0 * f0[17].copy()
    E = np.zeros((m.nTets, 3))
    spec = {1: ("Source", False), 2: ("Absorbing", False)}
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 1.0, 1.0, f0, E, 1e-4, 3, bc_spec=spec, source=src, variant=variant)
    for _ in range(3):
        s.update_pdf(sp, E)
        ctx.step_full(g, 1e-4)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    ctx.close()


def test_dangling_boundary_face_is_an_error(vt, oracle_mod):
    """A boundary face with neither neighbour nor particle BC is a null dereference in the
    reference (solver.cpp:319); the library refuses to step instead."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    ctx = vt.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create((8, 6, 4), [-1, -1, -1], [1, 1, 1], 1.0, 1.0)
    bc, col = face_bc_arrays(m, {})
    ctx.set_face_bc(g, bc, col)
    with pytest.raises(RuntimeError, match="no neighbour and no particle BC"):
        ctx.step_full(g, 1e-4)
    ctx.close()


def test_order_independence_bitwise(vt, oracle_mod):
    """The locality permutation is a pure layout choice: bit-identical states."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (16, 8, 8), [-3, -1, -1], [3, 1, 1]
    f0 = _smooth_state(m, n, vmin, vmax, seed=9)
    E = np.random.default_rng(2).standard_normal((m.nTets, 3))
    outs = []
    for order in (None, np.random.default_rng(3).permutation(m.nTets).astype(np.int32)):
        ctx = vt.Context(0)
        ctx.mesh_upload(tables_from_oracle(m, order=order))
        g = ctx.species_create(n, vmin, vmax, 1.0, 2.0)
        bc, col = face_bc_arrays(m, {})
        ctx.set_face_bc(g, bc, col)
        ctx.set_pdf(g, f0)
        ctx.field_set(E)
        for _ in range(2):
            ctx.step_full(g, 1e-4)
        outs.append((ctx.get_pdf(g), ctx.density(g)))
        ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


def test_step_host_roundtrip(vt, oracle_mod):
    """vt_step_full_host: E in through host memory, Density() out."""
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (16, 8, 8), [-3, -1, -1], [3, 1, 1]
    f0 = _smooth_state(m, n, vmin, vmax, seed=11)
    E = np.random.default_rng(4).standard_normal((m.nTets, 3))
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 1.0, 2.0, f0, np.zeros_like(E), 1e-4, 1,
                              order=np.random.default_rng(8).permutation(m.nTets).astype(np.int32))
    dens = np.empty(m.nTets)
    ctx.step_full_host(g, 1e-4, E, dens)
    s.update_pdf(sp, E)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    assert rel_l2(dens, s.density(sp)) <= TOL
    ctx.close()


def test_bench_configuration_against_oracle(vt, oracle_mod):
    """The exact bench configuration — 32^3 velocity grid, whole-tensor work items (chunk 0), the
    library's default kernel, C4 physics (bench.c4_setup) — on the 8x8x8-hex Kuhn box, against the
    oracle after 3 steps: state and density to relative L2 <= 1e-10."""
    import bench
    from vlasovtucker_b200 import synthetic
    hexes, nv = (8, 8, 8), 32
    cfg = bench.c4_setup(hexes, nv)
    nodes, tets, tris, ents = synthetic.kuhn_box(*hexes, cfg["lengths"])
    m = oracle_mod.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    s = oracle_mod.Sim(m)
    sp = s.add_species(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
    x = m.tetCentroid[:, 0] / cfg["lengths"][0]
    dens0 = cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x))
    s.set_maxwell(sp, dens0, cfg["T"])
    s.set_params(sp, cfg["dt"], fused=True)
    E = np.zeros((m.nTets, 3))
    E[:, 0] = 1e3 * np.cos(2 * PI * x)
    mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"], brick=(4, 4, 4))
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vt.PBC["Periodic"], np.uint8))
    ctx.set_maxwell(g, dens0, cfg["T"])
    ctx.field_set(E)
    ctx.step_config(chunk_planes=0, variant=64)
    for _ in range(3):
        s.update_pdf(sp, E)
        ctx.step_full(g, cfg["dt"])
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    assert rel_l2(ctx.density(g), s.density(sp)) <= TOL
    ctx.close()


def test_kuhn_box_full_size_properties(vt):
    """At the bench grid (32^3), on top of the comparison above: size-independent properties — a uniform
    state is a fixed point of transport, particles are conserved, and the step commutes with
    scaling (linearity)."""
    from vlasovtucker_b200 import synthetic
    mt = synthetic.periodic_kuhn_tables(6, 6, 6, brick=(3, 3, 3))
    n, vmin, vmax = (32, 32, 32), [-6, -6, -6], [6, 6, 6]
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 1.0)
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vt.PBC["Periodic"], np.uint8))
    ax = np.linspace(-6, 6, 32)
    V0, V1, V2 = np.meshgrid(ax, ax, ax, indexing="ij")
    maxw = np.exp(-(V0 ** 2 + V1 ** 2 + V2 ** 2) / 2).ravel(order="F")
    # (1) uniform in space, no field: every face flux cancels -> state unchanged to rounding
    f0 = np.tile(maxw, (mt.nTets, 1))
    ctx.set_pdf(g, f0)
    ctx.field_set(np.zeros((mt.nTets, 3)))
    ctx.step_full(g, 1e-3)
    f1 = ctx.get_pdf(g)
    assert np.abs(f1 - f0).max() <= 1e-13
    # (2) conservation + (3) linearity with a spatially varying state
    amp = 1 + 0.2 * np.sin(2 * PI * mt.tetCentroid[:, 0])
    fa = amp[:, None] * maxw[None, :]
    ctx.set_pdf(g, fa)
    n_before = (ctx.density(g) * mt.volume).sum()
    ctx.step_full(g, 1e-3)
    fa1 = ctx.get_pdf(g)
    n_after = (ctx.density(g) * mt.volume).sum()
    assert abs(n_after - n_before) <= 1e-12 * n_before
    ctx.set_pdf(g, 4.0 * fa)          # power of two: exact scaling
    ctx.step_full(g, 1e-3)
    assert np.array_equal(ctx.get_pdf(g), 4.0 * fa1)
    ctx.close()


@pytest.mark.parametrize("n,chunk", [((48, 48, 4), None), ((40, 40, 6), 3), ((36, 44, 5), None), ((50, 50, 3), None)])
def test_eight_warp_instances_for_planes_above_32x32(vt, oracle_mod, monkeypatch, n, chunk):
    """VT_STEP_WIDE8=1 (opt-in): planes above 32x32 on eight consumer warps with three to five columns per thread and
    consumer-side inflow-half neighbour loads — instead of the sixteen-warp instances, which spill from three columns
    on.  Same arithmetic with explicit roundings: parity with the oracle as every other instance, on a mesh with
    wall faces (the boundary instance) and with planes that do not fill the last column of every thread."""
    monkeypatch.setenv("VT_STEP_WIDE8", "1")
    m = oracle_mod.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)])
    vmin, vmax = [-3, -0.1, -0.2], [3, 0.1, 0.2]
    f0 = _smooth_state(m, n, vmin, vmax)
    rng = np.random.default_rng(7)
    E = rng.standard_normal((m.nTets, 3)) * 0.5
    spec = {1: ("Absorbing", True), 2: ("Free", False)}
    s, sp, ctx, g = _run_pair(vt, oracle_mod, m, n, vmin, vmax, 1.0, 10.0, f0, E, 1e-4, 3, bc_spec=spec, chunk=chunk, variant=64)
    for _ in range(3):
        s.update_pdf(sp, E)
        ctx.step_full(g, 1e-4)
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= TOL
    assert rel_l2(ctx.density(g), s.density(sp)) <= TOL
    ctx.close()

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


def test_virtual_ranks_tucker_bit_identical(vt):
    """Tucker species: the step kernel mirrors every boundary tet's new core, factors and ranks into
    the peers' ghost slots."""
    from vlasovtucker_b200 import multigpu, partition as part, synthetic
    dims = (4, 2, 2)
    mt = synthetic.periodic_kuhn_tables(*dims, (2.0, 1.0, 1.0))
    n, vmin, vmax = (12, 10, 8), [-3.0, -2.0, -2.0], [3.0, 2.0, 2.0]
    ax = [np.linspace(vmin[k], vmax[k], n[k]) for k in range(3)]
    V0, V1, V2 = np.meshgrid(*ax, indexing="ij")
    x = mt.tetCentroid[:, 0]
    f0 = np.stack([((1 + 0.3 * np.sin(3.14159 * xx)) * np.exp(-((V0 - 0.5 * np.cos(3.14159 * xx)) ** 2 + V1 ** 2 + V2 ** 2) / 2)
                    ).ravel(order="F") for xx in x])
    E = np.random.default_rng(1).standard_normal((mt.nTets, 3)) * 0.3
    world, steps, dt, eps = 2, 3, 2e-3, 1e-6
    owner = part.rcb_owner(mt.tetCentroid, world)
    vr = multigpu.VirtualRanks(mt, owner, world)
    ids = vr.species_create(n, vmin, vmax, 1.0, 1.5, tucker=(eps, 0))
    for r, ctx in enumerate(vr.ctxs):
        ctx.tucker_set_pdf(ids[r], f0[vr.lps[r].owned])
        ctx.field_set(E[vr.lps[r].owned])
    vr.fill_ghosts(ids)
    for _ in range(steps):
        vr.step_tucker(ids, dt)
    vr.sync()
    full = vr.gather([ctx.tucker_get_pdf(ids[r], n[0] * n[1] * n[2]) for r, ctx in enumerate(vr.ctxs)])
    ranks = vr.gather([ctx.tucker_ranks(ids[r]) for r, ctx in enumerate(vr.ctxs)])
    vr.close()
    ctx = vt.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 1.5)
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vt.PBC["Periodic"], np.uint8))
    ctx.tucker_enable(g, eps, 0)
    ctx.tucker_set_pdf(g, f0)
    ctx.field_set(E)
    for _ in range(steps):
        ctx.step_tucker(g, dt)
    ref = ctx.tucker_get_pdf(g, n[0] * n[1] * n[2])
    rref = ctx.tucker_ranks(g)
    ctx.close()
    assert np.array_equal(ranks, rref)
    assert np.array_equal(full, ref), f"rel diff {rel_l2(full, ref):.3e}"

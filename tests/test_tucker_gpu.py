"""Device Tucker path (vt_tucker_*, vt_step_tucker) against the oracle's Tucker algebra
(oracle/oracle_tucker.cpp) through the C ABI.  Tolerance: the reconstructed tensors and the
densities agree within comprErr + 1e-10 (relative L2), the bar BASELINE.json states."""
import numpy as np
import pytest

from conftest import face_bc_arrays, mesh_path, rel_l2, tables_from_oracle
from np_ref import vgrid
import tucker_dense_ref as tdr

pytestmark = pytest.mark.gpu


def _initial(m, n, vmin, vmax, drift=0.4):
    _, V = vgrid(n, vmin, vmax)
    L = m.points[:, 0].max()
    dens = 1 + 0.3 * np.sin(2 * np.pi * m.tetCentroid[:, 0] / L)
    mx = np.exp(-0.5 * ((V[0] - drift) ** 2 + (V[1] / 0.5) ** 2 + (V[2] / 0.5) ** 2))
    return dens[:, None] * mx[None, :]


def _ctx(m, n, vmin, vmax, mass, charge, spec, eps, max_rank=0):
    import vlasovtucker_b200 as vtb
    ctx = vtb.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create(n, vmin, vmax, mass, charge)
    bc, col = face_bc_arrays(m, spec)
    ctx.set_face_bc(g, bc, col)
    ctx.tucker_enable(g, eps, max_rank)
    return ctx, g, bc


def test_set_get_roundtrip(oracle_mod):
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    n = (9, 7, 5)
    ctx, g, _ = _ctx(m, n, [-1, -1, -1], [1, 1, 1], 1.0, 1.0, {}, 1e-6)
    rng = np.random.default_rng(0)
    f = rng.random((m.nTets, 9 * 7 * 5))
    ctx.tucker_set_pdf(g, f)
    assert rel_l2(ctx.tucker_get_pdf(g, f.shape[1]), f) <= 1e-12       # precision 0: exact (particle_data.cpp:64-69)
    assert rel_l2(ctx.tucker_density(g), f.sum(axis=1) * (2 / 8) * (2 / 6) * (2 / 4)) <= 1e-12
    core, U = ctx.tucker_factors(g, 3, n)
    for k in range(3):
        assert np.abs(U[k].T @ U[k] - np.eye(U[k].shape[1])).max() <= 1e-12   # orthonormal factors
    rec = np.einsum("ijk,ai,bj,ck->abc", core, U[0], U[1], U[2]).ravel(order="F")
    assert rel_l2(rec, f[3]) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("mesh,pairs,spec", [
    ("fully_periodic_coarse.msh", [(1, 2), (3, 4), (5, 6)], {}),
    ("rectangle.msh", [(3, 4), (5, 6)], {1: ("Absorbing", False), 2: ("Free", False)}),
])
@pytest.mark.parametrize("eps", [1e-6, 1e-4])
def test_step_parity(oracle_mod, mesh, pairs, spec, eps):
    m = oracle_mod.Mesh.load(mesh_path(mesh), pairs)
    n, vmin, vmax = (9, 7, 5), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    dt, mass, charge = 2e-3, 2.0, 1.0
    f = _initial(m, n, vmin, vmax)
    E = np.random.default_rng(4).standard_normal((m.nTets, 3))
    ts = oracle_mod.TuckerSim(m, n, vmin, vmax, mass, charge, eps)
    for e, (kind, _) in spec.items():
        ts.set_particle_bc(e, kind)
    ts.set_pdf(f)
    ctx, g, bc = _ctx(m, n, vmin, vmax, mass, charge, spec, eps)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(E)
    for _ in range(3):
        ts.update_pdf(dt, E)
        ctx.step_tucker(g, dt)
    fo, fg = ts.get_pdf(), ctx.tucker_get_pdf(g, f.shape[1])
    tol = eps + 1e-10
    assert rel_l2(fg, fo) <= tol
    assert rel_l2(ctx.tucker_density(g), ts.density()) <= tol
    assert np.abs(ctx.tucker_ranks(g) - ts.ranks()).max() <= 1
    # the same formulation in numpy (SVD instead of Gram/Jacobi): far tighter than eps
    gd = f.copy()
    for _ in range(3):
        gd, _ = tdr.step_dense(gd, m.adj, m.faceArea, m.tetVolume, m.faceNormal, bc, n, vmin, vmax, charge / mass, E, dt, eps, max(n))
    assert rel_l2(fg, gd) <= tol
    ctx.close()


def test_source_and_wall_charge(oracle_mod):
    """Source faces use sourcePDF in the neighbour's place (solver.cpp:335-338); Absorbing faces
    with collectCharge accumulate charge*dt*area*flux.Sum()*cellVolume (solver.cpp:171-178)."""
    import vlasovtucker_b200 as vtb
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(3, 4), (5, 6)])
    n, vmin, vmax = (9, 7, 5), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    eps, dt, mass, charge = 1e-6, 2e-3, 2.0, -1.5
    f = _initial(m, n, vmin, vmax, drift=-0.5)
    src = 1.3 * _initial(m, n, vmin, vmax, drift=0.8)[0]
    E = np.random.default_rng(6).standard_normal((m.nTets, 3))
    spec = {1: ("Absorbing", True), 2: ("Source", False)}
    ts = oracle_mod.TuckerSim(m, n, vmin, vmax, mass, charge, eps)
    ts.set_particle_bc(1, "Absorbing", True)
    ts.set_particle_bc(2, "Source", False, src)
    ts.set_pdf(f)
    ctx = vtb.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create(n, vmin, vmax, mass, charge)
    bc, col = face_bc_arrays(m, spec)
    ctx.set_source_pdfs(g, src[None, :])
    ctx.set_face_bc(g, bc, col, np.where(bc == vtb.PBC["Source"], 0, -1).astype(np.int32))
    ctx.tucker_enable(g, eps)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(E)
    for _ in range(3):
        ts.update_pdf(dt, E)
        ctx.step_tucker(g, dt)
    tol = eps + 1e-10
    assert rel_l2(ctx.tucker_get_pdf(g, f.shape[1]), ts.get_pdf()) <= tol
    qo, qg = ts.wall_charge(1), ctx.wall_charge(g, 1)
    assert qo != 0.0 and abs(qg - qo) <= tol * abs(qo)
    ctx.close()


def test_generic_entry_points_on_tucker_species(oracle_mod):
    """vt_species_set_maxwell / density / velocity / get_pdf serve a Tucker species too
    (ParticleData<Tucker>::SetMaxwellPDF, Density, Velocity — particle_data.cpp:23-128)."""
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (10, 8, 6), [-3.0, -2.0, -2.0], [3.0, 2.0, 2.0]
    dens = 1e3 * (1 + 0.2 * np.cos(m.tetCentroid[:, 1]))
    s = oracle_mod.Sim(m)
    sp = s.add_species(n, vmin, vmax, 9.1e-31, 1.0)
    T = 0.4 * 9.1e-31 / 1.38e-23
    s.set_maxwell(sp, dens, T, [0.3, 0.0, -0.2])
    ctx, g, _ = _ctx(m, n, vmin, vmax, 9.1e-31, 1.0, {}, 1e-6)
    ctx.set_maxwell(g, dens, T, [0.3, 0.0, -0.2])
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= 1e-12
    assert rel_l2(ctx.density(g), s.density(sp)) <= 1e-12
    assert np.abs(ctx.velocity(g) - s.velocity(sp)).max() <= 1e-10
    ctx.tucker_enable(g, 1e-6, 3)                        # re-configuration keeps the state
    assert ctx.tucker_ranks(g).max() <= 3
    assert rel_l2(ctx.get_pdf(g), s.get_pdf(sp)) <= 1e-9  # a Maxwellian has rank (1,1,1)
    ctx.close()


@pytest.mark.parametrize("n,cap", [((32, 24, 20), 0), ((48, 40, 36), 8), ((20, 64, 18), 8),
                                   ((32, 32, 32), 8), ((33, 25, 17), 8), ((24, 36, 18), 16), ((48, 48, 40), 12)])
def test_large_grids_against_dense_statement(n, cap):
    """Velocity grids above 16 nodes per axis take the 256-thread kernel (tiled Gram matrices with
    and without register prefetch, C5 shape 48^3 at rank 8).  The oracle's Tucker algebra is too
    slow there, so the check is against the numpy dense + SVD statement of the same update
    (tests/tucker_dense_ref.py, itself pinned to the oracle in test_tucker_cpu.py)."""
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    mt = synthetic.periodic_kuhn_tables(2, 1, 1, (1.0, 0.5, 0.5), brick=(1, 1, 1))
    vmin, vmax = [-3.0, -2.5, -2.0], [3.0, 2.5, 2.0]
    eps, dt = 1e-6, 2e-3
    _, V = vgrid(n, vmin, vmax)
    rng = np.random.default_rng(8)
    f = np.zeros((mt.nTets, n[0] * n[1] * n[2]))
    for t in range(mt.nTets):      # a few drifting anisotropic Maxwellians per tet: rank > 1
        for _ in range(3):
            c, s = rng.uniform(-0.8, 0.8, 3), rng.uniform(0.5, 1.0, 3)
            f[t] += rng.uniform(0.5, 1.5) * np.exp(-0.5 * sum(((V[k] - c[k]) / s[k]) ** 2 for k in range(3)))
    E = rng.standard_normal((mt.nTets, 3))
    ctx = vtb.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, vmin, vmax, 1.0, 1.0)
    bc = np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8)
    ctx.set_face_bc(g, bc)
    ctx.tucker_enable(g, eps, cap)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(E)
    ctx.step_tucker(g, dt)
    rmax = cap if cap else max(n)
    if cap:   # the initial tensors are stored with rank <= cap
        f = np.stack([tdr.truncate(row.reshape(n, order="F"), 0.0, cap)[0].ravel(order="F") for row in f])
    gd, ranks = tdr.step_dense(f, mt.nbr, mt.area, mt.volume, mt.normal, bc, n, vmin, vmax, 1.0, E, dt, eps, rmax)
    fg = ctx.tucker_get_pdf(g, f.shape[1])
    assert rel_l2(fg, gd) <= eps + 1e-10
    assert np.abs(ctx.tucker_ranks(g) - ranks).max() <= 1
    assert ctx.tucker_ranks(g).max() > 1
    ctx.close()


def test_max_rank_cap(oracle_mod):
    """ParticleData::SetMaxRank (C5: 48^3 at r = 8): ranks never exceed the cap and the capped
    rounding matches the oracle's."""
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (12, 10, 8), [-3.0, -2.0, -2.0], [3.0, 2.0, 2.0]
    eps, cap, dt = 1e-6, 3, 2e-3
    f = _initial(m, n, vmin, vmax)
    E = np.random.default_rng(5).standard_normal((m.nTets, 3))
    ts = oracle_mod.TuckerSim(m, n, vmin, vmax, 1.0, 1.0, eps, cap)
    ts.set_pdf(f)
    ctx, g, _ = _ctx(m, n, vmin, vmax, 1.0, 1.0, {}, eps, cap)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(E)
    for _ in range(2):
        ts.update_pdf(dt, E)
        ctx.step_tucker(g, dt)
    assert ctx.tucker_ranks(g).max() <= cap
    assert np.array_equal(ctx.tucker_ranks(g), ts.ranks())
    # with a binding cap the truncation error is large, and both sides make the same one
    assert rel_l2(ctx.tucker_get_pdf(g, f.shape[1]), ts.get_pdf()) <= 1e-6
    ctx.close()


def test_uniform_state_is_steady(oracle_mod):
    """Property: a spatially uniform separable state with E = 0 on a periodic mesh stays put up to
    the |v.n| rank-6 error times the (vanishing) jump — i.e. exactly, to rounding."""
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (16, 8, 8), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    _, V = vgrid(n, vmin, vmax)
    mx = np.exp(-0.5 * (V[0] ** 2 + (V[1] / 0.4) ** 2 + (V[2] / 0.4) ** 2))
    f = np.tile(mx, (m.nTets, 1))
    ctx, g, _ = _ctx(m, n, vmin, vmax, 1.0, 1.0, {}, 1e-6)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(np.zeros((m.nTets, 3)))
    for _ in range(5):
        ctx.step_tucker(g, 1e-3)
    assert rel_l2(ctx.tucker_get_pdf(g, f.shape[1]), f) <= 1e-9
    assert (ctx.tucker_ranks(g) == 1).all()
    ctx.close()


def test_errors(oracle_mod):
    import vlasovtucker_b200 as vtb
    m = oracle_mod.Mesh.load(mesh_path("rectangle.msh"), [(3, 4), (5, 6)])
    ctx = vtb.Context(0)
    ctx.mesh_upload(tables_from_oracle(m))
    g = ctx.species_create((8, 4, 4), [-1, -1, -1], [1, 1, 1], 1.0, 1.0)
    with pytest.raises(RuntimeError, match="Tucker"):
        ctx.step_tucker(g, 1e-3)                         # not enabled
    ctx.tucker_enable(g, 1e-6)
    ctx.tucker_set_pdf(g, np.ones((m.nTets, 128)))
    ctx.field_set(np.zeros((m.nTets, 3)))
    with pytest.raises(RuntimeError, match="boundary faces"):
        ctx.step_tucker(g, 1e-3)                         # x faces have no particle BC yet
    ctx.close()


@pytest.mark.parametrize("eps", [1e-8, 1e-10])
def test_step_parity_small_compression_error(oracle_mod, eps):
    """comprErr below ~1e-7 (the class default is 1e-10, particle_data.h:49): singular values taken from
    Gram-matrix eigenvalues alone are noise below 1e-8 |sigma|; the device refines the trailing
    eigen-directions inside their own subspace (csrc/tucker_kernel.inl, hosvd_truncate).  Against the oracle,
    whose rounding uses a one-sided Jacobi SVD of the unfoldings as the reference uses Eigen's SVD."""
    m = oracle_mod.Mesh.load(mesh_path("fully_periodic_coarse.msh"), [(1, 2), (3, 4), (5, 6)])
    n, vmin, vmax = (9, 7, 5), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    dt, mass, charge = 2e-3, 2.0, 1.0
    f = _initial(m, n, vmin, vmax)
    E = np.random.default_rng(4).standard_normal((m.nTets, 3))
    ts = oracle_mod.TuckerSim(m, n, vmin, vmax, mass, charge, eps)
    ts.set_pdf(f)
    ctx, g, bc = _ctx(m, n, vmin, vmax, mass, charge, {}, eps)
    ctx.tucker_set_pdf(g, f)
    ctx.field_set(E)
    for _ in range(3):
        ts.update_pdf(dt, E)
        ctx.step_tucker(g, dt)
    fo, fg = ts.get_pdf(), ctx.tucker_get_pdf(g, f.shape[1])
    tol = eps + 1e-10
    assert rel_l2(fg, fo) <= tol, rel_l2(fg, fo)
    assert rel_l2(ctx.tucker_density(g), ts.density()) <= tol
    assert np.abs(ctx.tucker_ranks(g) - ts.ranks()).max() <= 1
    ctx.close()


def test_truncation_rule_resolves_tiny_singular_values():
    """A tensor with prescribed multilinear singular values 1, 1e-3, 1e-6, 1e-9, 3e-11, 1e-13: compressing
    at eps = 1e-10 must keep the first four in every mode (sigma_j > eps |sigma| / sqrt(3), tucker.cpp:
    450-461) and reconstruct to ~eps, which needs singular values resolved far below the 1e-8 a Gram
    matrix gives."""
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    rng = np.random.default_rng(7)
    n = (12, 10, 8)
    sig = np.array([1.0, 1e-3, 1e-6, 1e-9, 3e-11, 1e-13])
    Q = [np.linalg.qr(rng.standard_normal((n[k], 6)))[0] for k in range(3)]
    core = np.zeros((6, 6, 6))
    for j in range(6):
        core[j, j, j] = sig[j]          # superdiagonal core: mode-k singular values are exactly sig
    X = np.einsum("ijk,ai,bj,ck->abc", core, Q[0], Q[1], Q[2])
    mt = synthetic.periodic_kuhn_tables(1, 1, 1)
    ctx = vtb.Context(0)
    ctx.mesh_upload(mt)
    g = ctx.species_create(n, [-1, -1, -1], [1, 1, 1], 1.0, 1.0)
    ctx.set_face_bc(g, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
    ctx.tucker_enable(g, 1e-10, 0)
    f = np.tile(X.ravel(order="F"), (mt.nTets, 1))
    ctx.tucker_set_pdf(g, f)                      # exact (precision 0)
    ctx.field_set(np.zeros((mt.nTets, 3)))
    ctx.step_tucker(g, 0.0)                       # dt = 0: the state is only re-rounded at eps = 1e-10
    r = ctx.tucker_ranks(g)
    assert (r == 4).all(), r
    out = ctx.tucker_get_pdf(g, f.shape[1])
    assert rel_l2(out, f) <= 1e-10
    ctx.close()


def test_gram_dfma_and_dmma_agree(oracle_mod):
    """The Gram matrices of the rounding come from the FP64 tensor cores (mma.sync m8n8k4, the default) or
    from DFMA (VT_TUCKER_GRAM=dfma); both must give the same step within rounding.  Run as two child
    processes because the switch is read once per process."""
    import os
    import subprocess
    import sys
    import tempfile
    from conftest import ROOT
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import vlasovtucker_b200 as vtb
from vlasovtucker_b200 import synthetic
mt = synthetic.periodic_kuhn_tables(2, 1, 1)
n = (20, 13, 9)
ax = [np.linspace(-3, 3, k) for k in n]
V0, V1, V2 = np.meshgrid(*ax, indexing="ij")
x = mt.tetCentroid[:, 0]
f = np.stack([((1 + 0.3 * np.sin(3 * xx)) * np.exp(-((V0 - 0.5 * np.cos(2 * xx)) ** 2 + V1 ** 2 + (V2 + 0.2) ** 2) / 2)).ravel(order="F") for xx in x])
ctx = vtb.Context(0)
ctx.mesh_upload(mt)
g = ctx.species_create(n, [-3, -3, -3], [3, 3, 3], 1.0, 1.5)
ctx.set_face_bc(g, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
ctx.tucker_enable(g, 1e-6, 0)
ctx.tucker_set_pdf(g, f)
ctx.field_set(0.3 * np.random.default_rng(1).standard_normal((mt.nTets, 3)))
for _ in range(3):
    ctx.step_tucker(g, 2e-3)
np.save(sys.argv[1], ctx.tucker_get_pdf(g, f.shape[1]))
''' % ROOT
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for mode in ("dmma", "dfma"):
            path = os.path.join(td, mode + ".npy")
            env = dict(os.environ, VT_TUCKER_GRAM=mode)
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(path))
    assert rel_l2(outs[0], outs[1]) <= 1e-9


@pytest.mark.parametrize("n,cap", [((34, 20, 18), 8), ((40, 33, 26), 16)])
def test_slab_kernel_agrees_with_general_kernel(n, cap):
    """csrc/tucker_slab.cu (slab-streaming, tensor-core contractions, parallel Jacobi) serves grids of
    33..48 nodes per axis with rank caps <= 16 and eps >= 5.5e-7; VT_TUCKER_KERNEL=general keeps k_tucker.
    Both evaluate the same six truncated HOSVDs: same ranks, state and moments within the compression
    error, on a mesh with Absorbing (collecting), Free and Source faces next to the periodic ones."""
    import os
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    mt = synthetic.periodic_kuhn_tables(2, 2, 1, (1.0, 1.0, 0.5), brick=(1, 1, 1))
    vmin, vmax = [-3.0, -2.5, -2.0], [3.0, 2.5, 2.0]
    eps, dt = 1e-6, 2e-3
    _, V = vgrid(n, vmin, vmax)
    rng = np.random.default_rng(11)
    f = np.zeros((mt.nTets, n[0] * n[1] * n[2]))
    for t in range(mt.nTets):
        for _ in range(3):
            c, s = rng.uniform(-0.8, 0.8, 3), rng.uniform(0.5, 1.0, 3)
            f[t] += rng.uniform(0.5, 1.5) * np.exp(-0.5 * sum(((V[k] - c[k]) / s[k]) ** 2 for k in range(3)))
    src = np.exp(-0.5 * sum((V[k] / 0.8) ** 2 for k in range(3)))[None, :]
    E = rng.standard_normal((mt.nTets, 3))
    bc = np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8)
    collect = np.zeros((mt.nTets, 4), np.uint8)
    source = np.full((mt.nTets, 4), -1, np.int32)
    kinds = [vtb.PBC["Absorbing"], vtb.PBC["Free"], vtb.PBC["Source"], vtb.PBC["Absorbing"]]
    ent = np.asarray(mt.entity).reshape(mt.nTets, 4)
    for i, (t, j) in enumerate(zip(*np.nonzero(ent > 0))):   # every second face on the box surface gets a wall BC
        if i % 2:
            continue
        bc[t, j] = kinds[(i // 2) % 4]
        if bc[t, j] == vtb.PBC["Absorbing"]:
            collect[t, j] = 1
        if bc[t, j] == vtb.PBC["Source"]:
            source[t, j] = 0
    out = {}
    for kernel in ("general", "slab"):
        os.environ["VT_TUCKER_KERNEL"] = kernel
        try:
            ctx = vtb.Context(0)
            ctx.mesh_upload(mt)
            g = ctx.species_create(n, vmin, vmax, 1.0, -1.0)
            ctx.set_source_pdfs(g, src)
            ctx.set_face_bc(g, bc, collect, source)
            ctx.tucker_enable(g, eps, cap)
            ctx.tucker_set_pdf(g, f)
            ctx.field_set(E)
            for _ in range(2):
                ctx.step_tucker(g, dt)
            ents = sorted(set(int(e) for e in np.asarray(mt.entity).ravel() if e > 0))
            out[kernel] = (ctx.tucker_get_pdf(g, f.shape[1]), ctx.tucker_ranks(g).copy(), ctx.tucker_density(g).copy(),
                           np.array([ctx.wall_charge(g, e) for e in ents]))
            ctx.close()
        finally:
            os.environ.pop("VT_TUCKER_KERNEL", None)
    fa, ra, da, wa = out["general"]
    fb, rb, db, wb = out["slab"]
    assert rel_l2(fb, fa) <= 2 * eps
    assert np.abs(ra - rb).max() <= 1
    assert rel_l2(db, da) <= 2 * eps
    assert np.allclose(wb, wa, rtol=1e-5, atol=1e-12 * max(1.0, np.abs(wa).max()))
    assert np.abs(wa).max() > 0

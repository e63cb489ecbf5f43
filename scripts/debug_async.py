import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle, vlasovtucker_b200 as vt
from conftest import tables_from_oracle, face_bc_arrays, mesh_path
m = oracle.Mesh.load(mesh_path("rectangle_fine.msh"), [(1, 2), (3, 4), (5, 6)])
for n, chunk, brick, variant in [((8, 6, 8), 1, 50, 2)]:
    vmin, vmax = [-3, -0.1, -0.2], [3, 0.1, 0.2]
    N = n[0] * n[1] * n[2]
    rng = np.random.default_rng(5)
    f0 = rng.random((m.nTets, N)); E = rng.standard_normal((m.nTets, 3)) * 0.5
    s = oracle.Sim(m); sp = s.add_species(n, vmin, vmax, 1.0, 10.0); s.set_pdf(sp, f0); s.set_params(sp, 1e-4, fused=True)
    ctx = vt.Context(0); ctx.mesh_upload(tables_from_oracle(m)); g = ctx.species_create(n, vmin, vmax, 1.0, 10.0)
    bc, col = face_bc_arrays(m, {}); ctx.set_face_bc(g, bc, col); ctx.set_pdf(g, f0); ctx.field_set(E)
    ctx.step_config(chunk_planes=chunk, brick_tets=brick, variant=variant)
    s.update_pdf(sp, E); ctx.step_full(g, 1e-4)
    fo, fg = s.get_pdf(sp), ctx.get_pdf(g)
    os.makedirs('gpurun_out', exist_ok=True); np.save('gpurun_out/dbg_fg.npy', fg)
    d = np.abs(fg - fo).reshape(m.nTets, n[2], n[0] * n[1])
    bad = np.argwhere(d.max(2) > 1e-9)
    byplane = np.bincount(bad[:, 1], minlength=n[2]) if len(bad) else []
    # within a bad plane: how many elements are off, and is the error confined to some lines?
    info = ""
    if len(bad):
        t, pl = bad[0]
        e = d[t, pl].reshape(n[1], n[0])
        info = f"first bad tet {t} plane {pl}: bad elems {(e > 1e-9).sum()} of {e.size}, bad lines {np.flatnonzero((e > 1e-9).any(1)).tolist()[:12]}"
    for (t, pl) in bad[:12].tolist() + bad[-6:].tolist():
        e = d[t, pl]
        print('BADITEM', t, pl, 'bad elems', np.flatnonzero(e > 1e-9).tolist())
    print(n, chunk, brick, variant, "rel %.2e" % (np.linalg.norm(fg - fo) / np.linalg.norm(fo)), "bad items", len(bad), "by plane", list(byplane), info, flush=True)
    ctx.close()

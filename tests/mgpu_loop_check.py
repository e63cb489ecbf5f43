#!/usr/bin/env python
"""Multi-GPU coupled loop (run under torchrun by tests/test_multigpu_gpu.py): sheath-style electrons on
rectangle_fine.msh partitioned by recursive coordinate bisection — absorbing charged wall, free
Dirichlet wall, Poisson solve partitioned like the tets, halo exchange fused into the step kernel,
wall charge summed over ranks — against the same loop on one GPU.  Not bit-identical by
construction (wall charge is accumulated with atomics, its value feeds the field): tolerance 1e-11."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402  (mesh tables only: test plumbing)
import vlasovtucker_b200 as vtb  # noqa: E402
from conftest import face_bc_arrays, mesh_path, poisson_bc_arrays, rel_l2, tables_from_oracle  # noqa: E402
from vlasovtucker_b200 import multigpu, partition as part  # noqa: E402

EPS0 = 8.85e-12
PI = 3.14159265358979323846


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kB, e, me, eV = 1.38e-23, 1.6e-19, 9.1e-31, 11604.518
    Te, dens = 1 * eV, 1e17
    debye = np.sqrt(EPS0 * kB * Te / dens) / e
    wp = e * np.sqrt(dens / (me * EPS0))
    dt = 1e-4 * 2 * PI / wp
    m = oracle.Mesh.load(mesh_path("rectangle_fine.msh"), [(3, 4), (5, 6)], scale=22 * debye)
    mt = tables_from_oracle(m)
    maxV = np.sqrt(-np.log(1e-6) * 2 * kB * Te / me)
    n, vmin, vmax = (32, 16, 4), [-4 * maxV, -maxV, -maxV], [4 * maxV, maxV, maxV]
    variant = int(os.environ.get("VT_VARIANT", "64"))
    bc, col = face_bc_arrays(m, {1: ("Absorbing", True), 2: ("Free", False)})
    spec = {1: ("Neumann", 0, 0.0), 2: ("Dirichlet", 0.0, 0)}
    qb, val, ng = poisson_bc_arrays(m, spec)
    area = float(m.faceArea[m.faceEntity == 1].sum())
    background = np.full(m.nTets, e * dens)          # immobile ions
    steps = 12

    def neumann(Q):
        spec[1] = ("Neumann", 0, (Q / area) / (2 * EPS0))
        _, v, g = poisson_bc_arrays(m, spec)
        return v, g

    owner = part.rcb_owner(mt.tetCentroid, world)
    lp = part.partition(mt, owner, rank)
    ctx = vtb.Context(local)
    ctx.mesh_upload(lp.tables)
    ps = multigpu.PartitionedSpecies(ctx, lp, dist, n, vmin, vmax, me, -e, bc_type=bc[lp.owned])
    ctx.set_face_bc(ps.sp, bc[lp.owned], col[lp.owned])
    ctx.set_maxwell(ps.sp, np.full(len(lp.owned), dens), Te)
    ctx.step_config(variant=variant)
    ps.fill_ghosts()
    # the field solve is partitioned like the tets (rows = owned tets, ghost values by peer stores, the CG
    # dot products summed over the ranks inside the solve kernel); Neumann data of the charged wall are
    # the only thing that goes through the host each step, as in the reference's loop (solver.cpp:120-132)
    multigpu.PartitionedPoisson(ctx, lp, dist, qb[lp.owned], val[lp.owned], ng[lp.owned])
    for _ in range(steps):
        ctx.charge_density([ps.sp], background[lp.owned])
        ctx.poisson_solve(download=False)
        ctx.step_full(ps.sp, dt)
        ctx.halo_barrier()
        Q = multigpu.wall_charge_total(ctx, ps.sp, 1, dist)
        v, g_ = neumann(Q)
        ctx.poisson_update_bc_values(v[lp.owned], g_[lp.owned])
    ctx.sync()
    out = [None] * world
    dist.all_gather_object(out, (lp.owned, ctx.get_pdf(ps.sp), ctx.field_get()[1]))
    if rank == 0:
        full = np.zeros((m.nTets, n[0] * n[1] * n[2]))
        phi = np.zeros(m.nTets)
        for ids, rows, ph in out:
            full[ids] = rows
            phi[ids] = ph
        one = vtb.Context(local)
        one.mesh_upload(mt)
        g = one.species_create(n, vmin, vmax, me, -e)
        one.set_face_bc(g, bc, col)
        one.set_maxwell(g, np.full(m.nTets, dens), Te)
        one.step_config(variant=variant)
        one.poisson_setup(qb, val, ng)
        for _ in range(steps):
            one.charge_density([g], background)
            phi1, E1 = one.poisson_solve()
            one.step_full(g, dt)
            one.poisson_update_bc_values(*neumann(one.wall_charge(g, 1)))
        ef = rel_l2(full, one.get_pdf(g))
        eq = abs(Q - one.wall_charge(g, 1)) / abs(one.wall_charge(g, 1))
        ephi = rel_l2(phi, phi1)
        print(f"mgpu_loop_check world={world}: f rel L2 {ef:.3e}  wall charge rel {eq:.3e}  phi rel L2 {ephi:.3e}  Q={Q:.6e}", flush=True)
        assert Q != 0.0 and ef <= 1e-11 and eq <= 1e-11 and ephi <= 1e-9
        one.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_LOOP_OK", flush=True)


if __name__ == "__main__":
    main()

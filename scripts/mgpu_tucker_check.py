#!/usr/bin/env python
"""Multi-GPU parity of the Tucker step (run under torchrun): the partitioned update, with every
boundary tet's new core/factors/ranks stored into the peers' ghost slots by the step kernel, must
reproduce the single-GPU Tucker update of the same global mesh bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlasovtucker_b200 as vtb  # noqa: E402
from vlasovtucker_b200 import multigpu, partition as part, synthetic  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (4, 4, 2)
    mt = synthetic.periodic_kuhn_tables(*dims, (1.0, 1.0, 0.5))
    n, vmin, vmax = (12, 10, 8), [-3.0, -2.5, -2.0], [3.0, 2.5, 2.0]
    N = n[0] * n[1] * n[2]
    ax = [np.linspace(vmin[k], vmax[k], n[k]) for k in range(3)]
    V = [a.ravel(order="F") for a in np.meshgrid(*ax, indexing="ij")]
    rng = np.random.default_rng(1)
    f0 = np.zeros((mt.nTets, N))
    for t in range(mt.nTets):
        for _ in range(2):
            c, s = rng.uniform(-0.8, 0.8, 3), rng.uniform(0.5, 1.0, 3)
            f0[t] += rng.uniform(0.5, 1.5) * np.exp(-0.5 * sum(((V[k] - c[k]) / s[k]) ** 2 for k in range(3)))
    E = rng.standard_normal((mt.nTets, 3))
    steps, dt, eps = 3, 2e-3, 1e-6

    owner = part.rcb_owner(mt.tetCentroid, world)
    lp = part.partition(mt, owner, rank)
    ctx = vtb.Context(local)
    ctx.mesh_upload(lp.tables)
    ps = multigpu.PartitionedSpecies(ctx, lp, dist, n, vmin, vmax, 1.0, 2.0, tucker=(eps, 0))
    ctx.tucker_set_pdf(ps.sp, f0[lp.owned])
    ctx.field_set(E[lp.owned])
    ps.fill_ghosts()
    dist.barrier()
    for _ in range(steps):
        ctx.step_tucker(ps.sp, dt)
        ctx.halo_barrier()
    ctx.sync()
    mine = (lp.owned, ctx.tucker_get_pdf(ps.sp, N), ctx.tucker_density(ps.sp), ctx.tucker_ranks(ps.sp))
    out = [None] * world
    dist.all_gather_object(out, mine)
    if rank == 0:
        full, dens, ranks = np.zeros_like(f0), np.zeros(mt.nTets), np.zeros((mt.nTets, 3), np.int32)
        for ids, rows, d, r in out:
            full[ids], dens[ids], ranks[ids] = rows, d, r
        one = vtb.Context(local)
        one.mesh_upload(mt)
        g = one.species_create(n, vmin, vmax, 1.0, 2.0)
        one.set_face_bc(g, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
        one.tucker_enable(g, eps, 0)
        one.tucker_set_pdf(g, f0)
        one.field_set(E)
        for _ in range(steps):
            one.step_tucker(g, dt)
        ref = one.tucker_get_pdf(g, N)
        same = np.array_equal(full, ref)
        rsame = np.array_equal(ranks, one.tucker_ranks(g))
        dsame = np.array_equal(dens, one.tucker_density(g))
        print(f"mgpu_tucker_check world={world}: state bit-identical={same} ranks equal={rsame} density bit-identical={dsame} "
              f"max|diff|={np.abs(full - ref).max():.3e} mean rank={ranks.mean():.2f}", flush=True)
        assert same and rsame and dsame
        one.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_TUCKER_OK", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun): the partitioned, halo-fused update over N GPUs
must reproduce the single-GPU update of the same global mesh bit for bit.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py
"""
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
    dims = (6, 4, 4)
    mt = synthetic.periodic_kuhn_tables(*dims, (1.5, 1.0, 1.0))
    n, vmin, vmax = (16, 8, 8), [-3.0, -1.0, -1.0], [3.0, 1.0, 1.0]
    rng = np.random.default_rng(0)
    f0 = rng.random((mt.nTets, 16 * 8 * 8))
    E = rng.standard_normal((mt.nTets, 3))
    steps, dt = 5, 1e-3

    for mode in ("rcb", "block"):
        owner = part.rcb_owner(mt.tetCentroid, world) if mode == "rcb" else part.block_owner(dims, part.rank_grid(world))
        lp = part.partition(mt, owner, rank)
        ctx = vtb.Context(local)
        ctx.mesh_upload(lp.tables)
        ps = multigpu.PartitionedSpecies(ctx, lp, dist, n, vmin, vmax, 1.0, 2.0)
        ctx.set_pdf(ps.sp, f0[lp.owned])
        ctx.field_set(E[lp.owned])
        ctx.step_config(chunk_planes=2, variant=int(os.environ.get("VT_VARIANT", "0")))
        ps.fill_ghosts()
        dist.barrier()
        for _ in range(steps):
            ctx.step_full(ps.sp, dt)
            ctx.halo_barrier()
        ctx.sync()
        mine = (lp.owned, ctx.get_pdf(ps.sp), ctx.density(ps.sp))
        out = [None] * world
        dist.all_gather_object(out, mine)
        if rank == 0:
            full = np.zeros_like(f0)
            dens = np.zeros(mt.nTets)
            for ids, rows, d in out:
                full[ids] = rows
                dens[ids] = d
            single = vtb.Context(local)
            single.mesh_upload(mt)
            g = single.species_create(n, vmin, vmax, 1.0, 2.0)
            single.set_face_bc(g, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
            single.set_pdf(g, f0)
            single.field_set(E)
            single.step_config(chunk_planes=2, variant=int(os.environ.get("VT_VARIANT", "0")))
            for _ in range(steps):
                single.step_full(g, dt)
            ref = single.get_pdf(g)
            same = np.array_equal(full, ref)
            dsame = np.array_equal(dens, single.density(g))
            print(f"mgpu_check[{mode}] world={world}: state bit-identical={same} density bit-identical={dsame} "
                  f"max|diff|={np.abs(full - ref).max():.3e}", flush=True)
            assert same and dsame
            single.close()
        dist.barrier()
        ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()

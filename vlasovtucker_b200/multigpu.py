"""One-process-per-GPU driver of the partitioned full-format update (bench.py --gpus N, tests).

Plumbing only: torch.distributed carries the 192-byte CUDA-IPC handles once at set-up; the
per-step boundary exchange is done by the step kernel itself (peer stores over NVLink into the
neighbours' ghost rows) followed by a device-side flag barrier — no NCCL call in the data path.
"""
import numpy as np

from . import partition as part
from . import synthetic
from .context import PBC, Context

PI = 3.14159265358979323846


class PartitionedSpecies:
    """A species on one rank of a partitioned mesh: owned + ghost rows, halo wired up."""

    def __init__(self, ctx, lp, dist, n, vmin, vmax, mass, charge, bc_type=None):
        self.ctx, self.lp, self.dist = ctx, lp, dist
        self.sp = ctx.species_create(n, vmin, vmax, mass, charge)
        nO = len(lp.owned)
        ctx.set_face_bc(self.sp, np.full((nO, 4), PBC["Periodic"], np.uint8) if bc_type is None else bc_type)
        mine = ctx.halo_export(self.sp)
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, mine.tobytes())
        if lp.peers:
            handles = np.stack([np.frombuffer(gathered[q], np.uint8) for q in lp.peers])
            ctx.halo_attach(self.sp, lp.rank, lp.peers, handles)
            ctx.halo_set_push(self.sp, lp.push_peer, lp.push_row)

    def fill_ghosts(self):
        """Initial ghost fill: push the current owned boundary rows to the peers."""
        self.ctx.halo_push_current(self.sp)
        self.ctx.halo_barrier()
        self.ctx.sync()


class WeakScaledBox:
    """Config C4 weak scaling: every rank owns one block of `hexes` Kuhn hexes of a periodic box
    that is `rank_grid(world)` blocks large."""

    def __init__(self, rank, world, local_device, hexes, cfg, brick, dist):
        self.rank, self.world, self.dist = rank, world, dist
        grid = part.rank_grid(world)
        ghex = tuple(h * g for h, g in zip(hexes, grid))
        glen = tuple(l * g for l, g in zip(cfg["lengths"], grid))
        mtg = synthetic.periodic_kuhn_tables(*ghex, glen)
        owner = part.block_owner(ghex, grid)
        order, brick_tets = synthetic.brick_order(*ghex, brick)
        local_order = order[owner[order] == rank]
        self.lp = lp = part.partition(mtg, owner, rank, local_order)
        lp.tables.brickTets = brick_tets
        self.mt = lp.tables
        self.ctx = ctx = Context(local_device)
        ctx.mesh_upload(lp.tables)
        self.ps = PartitionedSpecies(ctx, lp, dist, cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
        self.sp = self.ps.sp
        x = lp.tables.tetCentroid[:, 0] / glen[0]
        ctx.set_maxwell(self.sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x)), cfg["T"])
        self.E = np.zeros((lp.tables.nTets, 3))
        self.E[:, 0] = 1e3 * np.cos(2 * PI * x)
        ctx.field_set(self.E)
        self.ps.fill_ghosts()
        dist.barrier()

    def step(self, dt):
        self.ctx.step_full(self.sp, dt)
        self.ctx.halo_barrier()

    def e2e(self, dt, steps, barrier):
        import time
        dens = np.empty(self.mt.nTets)
        self.ctx.step_full_host(self.sp, dt, self.E, dens)
        self.ctx.halo_barrier()
        barrier()
        t0 = time.perf_counter()
        self.ctx.profile_begin()
        for _ in range(steps):
            self.ctx.step_full_host(self.sp, dt, self.E, dens)
            self.ctx.halo_barrier()
        region_ms, _, _ = self.ctx.profile_end()
        return max(region_ms, (time.perf_counter() - t0) * 1e3)

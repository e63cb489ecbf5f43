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

    def __init__(self, ctx, lp, dist, n, vmin, vmax, mass, charge, bc_type=None, tucker=None):
        """tucker = (comprErr, maxRank) keeps the species in Tucker format (ParticleData<Tucker>)."""
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
        if tucker is not None:
            ctx.tucker_enable(self.sp, tucker[0], tucker[1])
            mine = ctx.tucker_halo_export(self.sp)
            dist.all_gather_object(gathered, mine.tobytes())
            if lp.peers:
                ctx.tucker_halo_attach(self.sp, np.stack([np.frombuffer(gathered[q], np.uint8) for q in lp.peers]))

    def fill_ghosts(self):
        """Initial ghost fill: push the current owned boundary rows to the peers."""
        self.ctx.halo_push_current(self.sp)
        self.ctx.halo_barrier()
        self.ctx.sync()


class VirtualRanks:
    """All ranks of a partitioned mesh inside ONE process: one context per rank, on the devices given
    (the same device for every rank = "virtual ranks"), ghost rows wired directly
    to the peers' buffers (vt_halo_attach_local).  Exactly the kernels, push lists and device-side
    barrier of the one-process-per-GPU run; only the handle exchange differs.  A box with a single
    GPU can therefore check the halo path."""

    def __init__(self, tables, owner, world, devices=None, order=None):
        devices = [0] * world if devices is None else list(devices)
        self.world = world
        owner = np.asarray(owner)
        self.owner = owner
        self.lps = [part.partition(tables, owner, r, None if order is None else np.asarray(order)[owner[np.asarray(order)] == r])
                    for r in range(world)]
        self.ctxs = [Context(devices[r]) for r in range(world)]
        for ctx, lp in zip(self.ctxs, self.lps):
            ctx.mesh_upload(lp.tables)
        self.nGlobal = tables.nTets

    def species_create(self, n, vmin, vmax, mass, charge, bc_type=None, collect=None, tucker=None):
        """One species on every rank; bc_type/collect are global (nTets, 4) arrays.  Returns the
        per-rank species ids (equal on all ranks when the calls are made in the same order)."""
        ids = []
        for ctx, lp in zip(self.ctxs, self.lps):
            sp = ctx.species_create(n, vmin, vmax, mass, charge)
            nO = len(lp.owned)
            bc = np.full((nO, 4), PBC["Periodic"], np.uint8) if bc_type is None else bc_type[lp.owned]
            ctx.set_face_bc(sp, bc, None if collect is None else collect[lp.owned])
            ids.append(sp)
        for r, (ctx, lp) in enumerate(zip(self.ctxs, self.lps)):
            if lp.peers:
                ctx.halo_attach_local(ids[r], r, lp.peers, [self.ctxs[q] for q in lp.peers], [ids[q] for q in lp.peers])
                ctx.halo_set_push(ids[r], lp.push_peer, lp.push_row)
        if tucker is not None:
            for r, ctx in enumerate(self.ctxs):
                ctx.tucker_enable(ids[r], tucker[0], tucker[1])
            for r, (ctx, lp) in enumerate(zip(self.ctxs, self.lps)):
                if lp.peers:
                    ctx.tucker_halo_attach_local(ids[r], [self.ctxs[q] for q in lp.peers], [ids[q] for q in lp.peers])
        return ids

    def poisson_setup(self, bc_type, bc_value=None, bc_normal_grad=None):
        """Partitioned PoissonSolver: every rank assembles its rows (global (nTets, 4) BC arrays in),
        then the ranks are wired to each other (ghost values by peer stores, in-kernel reductions)."""
        any_dirichlet = bool(np.any(np.asarray(bc_type) == 2))
        for ctx, lp in zip(self.ctxs, self.lps):
            ctx.poisson_set_global_dirichlet(any_dirichlet)
            ctx.poisson_setup(np.asarray(bc_type)[lp.owned],
                              None if bc_value is None else np.asarray(bc_value)[lp.owned],
                              None if bc_normal_grad is None else np.asarray(bc_normal_grad)[lp.owned])
        for r, (ctx, lp) in enumerate(zip(self.ctxs, self.lps)):
            ctx.poisson_comm_attach_local(r, self.ctxs)
            ctx.poisson_set_push(lp.push_rank, lp.push_row)

    def poisson_solve(self, rho=None):
        """Launch every rank's solve (nothing waits on the host in between); rho = global array or None
        (device-resident charge density)."""
        for ctx, lp in zip(self.ctxs, self.lps):
            ctx.poisson_solve(None if rho is None else np.asarray(rho)[lp.owned], download=False)

    def scatter(self, a):
        """Global per-tet array -> list of per-rank owned slices."""
        return [np.ascontiguousarray(a[lp.owned]) for lp in self.lps]

    def gather(self, parts):
        out = np.zeros((self.nGlobal,) + parts[0].shape[1:], parts[0].dtype)
        for lp, p in zip(self.lps, parts):
            out[lp.owned] = p
        return out

    def fill_ghosts(self, ids):
        for r, ctx in enumerate(self.ctxs):
            ctx.halo_push_current(ids[r])
        self.barrier()
        self.sync()

    def barrier(self):
        """Every rank's device-side barrier, queued on its own stream (they wait for each other on
        the device, not on the host)."""
        for ctx, lp in zip(self.ctxs, self.lps):
            ctx.halo_barrier()

    def sync(self):
        for ctx in self.ctxs:
            ctx.sync()

    def step_full(self, ids, dt, ext=(0.0, 0.0, 0.0)):
        for r, ctx in enumerate(self.ctxs):
            ctx.step_full(ids[r], dt, ext)
        self.barrier()

    def step_tucker(self, ids, dt, ext=(0.0, 0.0, 0.0)):
        for r, ctx in enumerate(self.ctxs):
            ctx.step_tucker(ids[r], dt, ext)
        self.barrier()

    def close(self):
        self.sync()
        for ctx in self.ctxs:
            ctx.close()


class WeakScaledBox:
    """Config C4 weak scaling: every rank owns one block of `hexes` Kuhn hexes of a periodic box
    that is `rank_grid(world)` blocks large."""

    def __init__(self, rank, world, local_device, hexes, cfg, brick, dist, tucker=None, init=None):
        """tucker = (comprErr, maxRank): the species is kept in Tucker format; init(ctx, sp, x) sets the
        initial distribution functions from the normalised x coordinate of the owned tets (default:
        the C4 Maxwellian with a 1 % density wave)."""
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
        self.ps = PartitionedSpecies(ctx, lp, dist, cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"],
                                     tucker=tucker)
        self.sp = self.ps.sp
        self.tucker = tucker is not None
        x = lp.tables.tetCentroid[:, 0] / glen[0]
        if init is not None:
            init(ctx, self.sp, x)
        else:
            ctx.set_maxwell(self.sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x)), cfg["T"])
        self.E = np.zeros((lp.tables.nTets, 3))
        self.E[:, 0] = 1e3 * np.cos(2 * PI * x)
        ctx.field_set(self.E)
        self.ps.fill_ghosts()
        dist.barrier()

    def step(self, dt):
        if self.tucker:
            self.ctx.step_tucker(self.sp, dt)
        else:
            self.ctx.step_full(self.sp, dt)
        self.ctx.halo_barrier()

    def step_host(self, dt, dens):
        """One step through host buffers: E in, Density() out."""
        if self.tucker:
            self.ctx.field_set(self.E)
            self.ctx.step_tucker(self.sp, dt)
            dens[:] = self.ctx.tucker_density(self.sp)
        else:
            self.ctx.step_full_host(self.sp, dt, self.E, dens)
        self.ctx.halo_barrier()

    def e2e(self, dt, steps, barrier):
        import time
        dens = np.empty(self.mt.nTets)
        self.step_host(dt, dens)
        barrier()
        t0 = time.perf_counter()
        self.ctx.profile_begin()
        for _ in range(steps):
            self.step_host(dt, dens)
        region_ms, _, _ = self.ctx.profile_end()
        return max(region_ms, (time.perf_counter() - t0) * 1e3)


class PartitionedPoisson:
    """PoissonSolver of a partitioned run, one process per GPU: rows = owned tets, phi / z / grad(phi) of
    the ghost rows by peer stores (CUDA-IPC), the dot products of every CG iteration summed over the
    ranks inside the solve kernel.  torch.distributed only carries the 128-byte handles at set-up."""

    def __init__(self, ctx, lp, dist, bc_type, bc_value=None, bc_normal_grad=None, any_dirichlet=None):
        world = dist.get_world_size()
        if any_dirichlet is None:
            flags = [None] * world
            dist.all_gather_object(flags, bool(np.any(np.asarray(bc_type) == 2)))
            any_dirichlet = any(flags)
        ctx.poisson_set_global_dirichlet(any_dirichlet)
        ctx.poisson_setup(bc_type, bc_value, bc_normal_grad)
        mine = ctx.poisson_comm_export()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tobytes())
        ctx.poisson_comm_attach(lp.rank, np.stack([np.frombuffer(g, np.uint8) for g in gathered]))
        ctx.poisson_set_push(lp.push_rank, lp.push_row)
        dist.barrier()


def wall_charge_total(ctx, sp, entity, dist):
    """Absorbed charge of one boundary entity summed over the ranks (solver.cpp:171-178 accumulates
    it per face; the faces of an entity are spread over the partitions)."""
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, ctx.wall_charge(sp, entity))
    return float(np.sum(parts))   # fixed rank order: identical on every rank

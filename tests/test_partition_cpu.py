"""CPU tests of the multi-GPU host logic: partition tables, ghost/push lists, and a
world_size-2 gloo run of the halo pattern (push owned boundary rows into the peers' ghost rows,
barrier, step) that must reproduce the single-rank result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, mesh_path, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tests"))
import np_ref  # noqa: E402


def _case(dims=(4, 3, 3)):
    from vlasovtucker_b200 import synthetic
    mt = synthetic.periodic_kuhn_tables(*dims, (1.0, 0.8, 0.9))
    n, vmin, vmax = (4, 3, 3), [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    rng = np.random.default_rng(0)
    f0 = rng.random((mt.nTets, 36))
    E = rng.standard_normal((mt.nTets, 3))
    return mt, n, vmin, vmax, f0, E


def test_numpy_step_matches_oracle(oracle_mod):
    """Anchor the table-level numpy step used below to the oracle (bit for bit: same IEEE ops)."""
    from vlasovtucker_b200 import synthetic
    dims, lengths = (3, 3, 3), (1.0, 0.8, 0.9)
    nodes, tets, tris, ents = synthetic.kuhn_box(*dims, lengths)
    om = oracle_mod.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    mt = synthetic.periodic_kuhn_tables(*dims, lengths)
    n, vmin, vmax = (4, 3, 3), [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    rng = np.random.default_rng(0)
    f0 = rng.random((mt.nTets, 36))
    E = rng.standard_normal((mt.nTets, 3))
    s = oracle_mod.Sim(om)
    sp = s.add_species(n, vmin, vmax, 2.0, 3.0)
    s.set_pdf(sp, f0)
    s.set_params(sp, 1e-3, fused=True)
    s.update_pdf(sp, E)
    f1 = np_ref.step_tables(f0, mt.nbr, mt.area, mt.volume, mt.normal, n, vmin, vmax, 3.0 / 2.0, E, 1e-3)
    assert rel_l2(f1, s.get_pdf(sp)) < 1e-15


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_partition_tables_consistent(world):
    from vlasovtucker_b200 import partition as part
    mt, *_ = _case((4, 4, 4))
    owner = part.rcb_owner(mt.tetCentroid, world)
    assert sorted(np.unique(owner).tolist()) == list(range(world))
    counts = np.bincount(owner)
    assert counts.max() - counts.min() <= 1
    parts = [part.partition(mt, owner, r) for r in range(world)]
    for lp in parts:
        nO = len(lp.owned)
        # local neighbour indices resolve to the right global tets
        loc2glob = np.concatenate([lp.owned, lp.ghost])
        assert np.array_equal(loc2glob[lp.tables.nbr], mt.nbr[lp.owned])
        assert lp.tables.nGhost == len(lp.ghost)
        # ghosts grouped by owner rank, sorted inside a group
        key = lp.ghost_owner.astype(np.int64) * mt.nTets + lp.ghost
        assert np.all(np.diff(key) > 0)
        # every push entry lands on the row where the peer expects this tet
        for t in range(nO):
            for q in range(4):
                if lp.push_peer[t, q] < 0:
                    continue
                peer = parts[lp.peers[lp.push_peer[t, q]]]
                row = lp.push_row[t, q]
                assert row >= len(peer.owned)
                assert peer.ghost[row - len(peer.owned)] == lp.owned[t]
    # every ghost row of every rank is pushed exactly once
    for lp in parts:
        hits = np.zeros(len(lp.ghost), int)
        for other in parts:
            if other.rank == lp.rank or lp.rank not in other.peers:
                continue
            pi = other.peers.index(lp.rank)
            rows = other.push_row[other.push_peer == pi]
            np.add.at(hits, rows - len(lp.owned), 1)
        assert np.all(hits == 1)


def test_block_owner_and_rank_grid():
    from vlasovtucker_b200 import partition as part
    assert part.rank_grid(1) == (1, 1, 1)
    assert part.rank_grid(2) == (2, 1, 1)
    assert part.rank_grid(4) == (2, 2, 1)
    assert part.rank_grid(8) == (2, 2, 2)
    owner = part.block_owner((4, 4, 2), (2, 2, 1))
    assert np.array_equal(np.bincount(owner), np.full(4, 6 * 8))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from vlasovtucker_b200 import partition as part
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    import torch
    mt, n, vmin, vmax, f0, E = _case()
    owner = part.rcb_owner(mt.tetCentroid, world)
    lp = part.partition(mt, owner, rank)
    nO, nG = len(lp.owned), len(lp.ghost)
    N = f0.shape[1]
    state = np.zeros((nO + nG, N))
    state[:nO] = f0[lp.owned]

    def exchange(rows_owned):
        """The halo pattern of the CUDA path with gloo standing in for the NVLink peer stores:
        each rank delivers its pushed rows into the peers' ghost rows, then everybody waits."""
        sends = {}
        for pi, q in enumerate(lp.peers):
            sel = np.argwhere(lp.push_peer == pi)
            rows = lp.push_row[lp.push_peer == pi]
            sends[q] = (rows, rows_owned[sel[:, 0]])
        gathered = [None] * world
        dist.all_gather_object(gathered, sends)
        for src in range(world):
            if src == rank or rank not in gathered[src]:
                continue
            rows, data = gathered[src][rank]
            state[rows] = data
        dist.barrier()

    exchange(state[:nO])
    for _ in range(3):
        new = np_ref.step_tables(state, lp.tables.nbr, lp.tables.area, lp.tables.volume, lp.tables.normal, n, vmin,
                                 vmax, 1.5, E[lp.owned], 1e-3)
        state[:nO] = new
        exchange(new)
    out = [None] * world
    dist.all_gather_object(out, (lp.owned, state[:nO]))
    if rank == 0:
        full = np.zeros_like(f0)
        for ids, rows in out:
            full[ids] = rows
        q.put(full)
    dist.destroy_process_group()


def test_gloo_two_rank_halo_reproduces_single_rank():
    import torch.multiprocessing as mp
    mt, n, vmin, vmax, f0, E = _case()
    ref = f0.copy()
    for _ in range(3):
        ref = np_ref.step_tables(ref, mt.nbr, mt.area, mt.volume, mt.normal, n, vmin, vmax, 1.5, E, 1e-3)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full, ref)       # partition independence: bit-identical

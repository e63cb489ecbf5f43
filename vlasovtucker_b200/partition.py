"""Mesh partitioning for the multi-GPU path (host logic, numpy only).

The reference is single-process; partitioning is new.  Every rank derives the
same partition from the same global tables, so no set-up communication is needed beyond the
CUDA-IPC handle exchange:

* ``owner[t]``      rank that updates tet t (any deterministic function of the global mesh);
* local rows        owned tets first (in the rank's locality order), then ghost tets — the
                    neighbours of owned tets that another rank owns — grouped by owner rank
                    and sorted by global tet index inside a group;
* push lists        for every owned tet, the (peer, ghost row on that peer) pairs where its
                    new state must also be stored each step (the fused halo exchange of K1).
"""
from dataclasses import dataclass

import numpy as np

from .context import MeshTables


@dataclass
class LocalPart:
    rank: int
    owned: np.ndarray          # global tet ids of the owned rows, local order
    ghost: np.ndarray          # global tet ids of the ghost rows, local order
    ghost_owner: np.ndarray    # owner rank of each ghost row
    tables: MeshTables         # local tables: nbr indexes local rows, nGhost set
    peers: list                # ranks this rank exchanges with (sorted)
    push_peer: np.ndarray      # (nOwned,4) index into `peers`, -1 unused
    push_row: np.ndarray       # (nOwned,4) ghost row index on that peer
    recv_rows: dict            # peer rank -> slice of local ghost rows filled by that peer
    push_rank: np.ndarray = None   # (nOwned,4) push_peer as ranks instead of indices into `peers` (Poisson halo)


def ghost_list(nbr, owner, rank):
    """Global ids of rank's ghost tets grouped by owner rank, sorted by global id in a group."""
    mine = owner == rank
    nb = nbr[mine].ravel()
    nb = nb[nb >= 0]
    g = np.unique(nb[owner[nb] != rank])
    order = np.lexsort((g, owner[g]))
    return g[order]


def partition(mt: MeshTables, owner, rank, local_order=None):
    """Extract rank's part of the global tables ``mt`` (reference tet order)."""
    owner = np.asarray(owner)
    nT = mt.nTets
    owned = np.flatnonzero(owner == rank) if local_order is None else np.asarray(local_order)
    assert np.all(owner[owned] == rank) and len(np.unique(owned)) == int((owner == rank).sum())
    ghost = ghost_list(mt.nbr, owner, rank)
    nO, nG = len(owned), len(ghost)
    g2l = np.full(nT, -1, np.int64)
    g2l[owned] = np.arange(nO)
    g2l[ghost] = nO + np.arange(nG)
    nbr = mt.nbr[owned]
    lnbr = np.where(nbr >= 0, g2l[np.maximum(nbr, 0)], -1).astype(np.int32)
    assert not np.any((nbr >= 0) & (lnbr < 0))
    gnb = mt.nbr[ghost] if nG else np.zeros((0, 4), np.int64)
    ghost_nbr = np.where(gnb >= 0, g2l[np.maximum(gnb, 0)], -1).astype(np.int32)
    local = MeshTables(nbr=lnbr, area=mt.area[owned], volume=mt.volume[owned], normal=mt.normal[owned],
                       entity=mt.entity[owned], tetCentroid=mt.tetCentroid[owned],
                       faceCentroid=mt.faceCentroid[owned], nGhost=nG, periodic=list(mt.periodic),
                       globalTets=nT, globalId=np.concatenate([owned, ghost]).astype(np.int32), ghostNbr=ghost_nbr,
                       ghostArea=mt.area[ghost], ghostNormal=mt.normal[ghost], ghostTetCentroid=mt.tetCentroid[ghost],
                       ghostFaceCentroid=mt.faceCentroid[ghost])
    gowner = owner[ghost]
    # peers: ranks that own my ghosts or hold my tets as ghosts (symmetric for face adjacency)
    peers = sorted(set(int(r) for r in np.unique(gowner)))
    recv_rows = {}
    for q in peers:
        idx = np.flatnonzero(gowner == q)
        recv_rows[q] = slice(nO + int(idx[0]), nO + int(idx[-1]) + 1)
    push_peer = np.full((nO, 4), -1, np.int32)
    push_row = np.full((nO, 4), -1, np.int32)
    fill = np.zeros(nO, np.int64)
    for pi, q in enumerate(peers):
        gq = ghost_list(mt.nbr, owner, q)                # q's ghost rows, q's local order
        nOq = int((owner == q).sum())
        sel = np.flatnonzero(owner[gq] == rank)          # those owned by me
        mine_local = g2l[gq[sel]]
        rows_on_q = nOq + sel
        slot = fill[mine_local]
        if np.any(slot >= 4):
            raise RuntimeError("a tet is a ghost on more than 4 peers")
        push_peer[mine_local, slot] = pi
        push_row[mine_local, slot] = rows_on_q
        fill[mine_local] += 1
    push_rank = np.where(push_peer >= 0, np.asarray(peers + [0], np.int32)[np.maximum(push_peer, 0)], -1).astype(np.int32)
    return LocalPart(rank=rank, owned=owned, ghost=ghost, ghost_owner=gowner, tables=local, peers=peers,
                     push_peer=push_peer, push_row=push_row, recv_rows=recv_rows, push_rank=push_rank)


def rcb_owner(centroids, nparts):
    """Recursive coordinate bisection on tet centroids, ties broken by tet index: a deterministic
    partition of an unstructured mesh into ``nparts`` (any positive integer) parts."""
    c = np.asarray(centroids)
    owner = np.zeros(len(c), np.int32)

    def split(ids, lo, n):
        if n == 1:
            owner[ids] = lo
            return
        ext = c[ids].max(0) - c[ids].min(0)
        ax = int(np.argmax(ext))
        o = np.lexsort((ids, c[ids, ax]))
        nl = n // 2
        cut = (len(ids) * nl) // n
        split(ids[o[:cut]], lo, nl)
        split(ids[o[cut:]], lo + nl, n - nl)

    split(np.arange(len(c)), 0, int(nparts))
    return owner


def rank_grid(world):
    """Near-cubic process grid (gx >= gy >= gz) with gx*gy*gz == world."""
    best = None
    for gx in range(1, world + 1):
        if world % gx:
            continue
        for gy in range(1, world // gx + 1):
            if (world // gx) % gy:
                continue
            gz = world // gx // gy
            dims = tuple(sorted((gx, gy, gz), reverse=True))
            score = dims[0] - dims[2]
            if best is None or score < best[0]:
                best = (score, dims)
    return best[1]


def block_owner(global_hexes, grid):
    """Owner rank of every tet of the Kuhn box (6 tets per hex, hexes x-fastest) for a block
    decomposition of the hexes over a process grid."""
    nx, ny, nz = global_hexes
    gx, gy, gz = grid
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    r = ((hi.ravel() // (nx // gx)) * gy + (hj.ravel() // (ny // gy))) * gz + (hk.ravel() // (nz // gz))
    return np.repeat(r.astype(np.int32), 6)

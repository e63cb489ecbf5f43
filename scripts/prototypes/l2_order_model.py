"""CPU model of the DRAM traffic of the full-format step as a function of the ORDER in which tets are
processed: an LRU cache of C whole rows (one row = one tet's n0*n1*n2 doubles), accessed as the
kernel does (own row + the four neighbour rows per tet, `width` tets in flight at once).  Calibrated
on the measured traffic of the default order (4x4x4-hex bricks: 101.3 GB per launch on the
28x28x28-hex share = 1.93 row reads per tet), then used to rank other orders.  A guide for what to measure, not a prediction: a plane-granular
version of the same model that fits this point (capacity ~230 rows, i.e. about half of the 126 MB L2 —
each die keeping its own copy of rows read from both) over-estimates the measured traffic of
velocity-chunked items (7x7x4-hex bricks, 4 planes: 1.35 row reads per tet measured, 1.7 modelled).

    python scripts/prototypes/l2_order_model.py
"""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vlasovtucker_b200 import synthetic  # noqa: E402


def misses_per_tet(nbr, order, cap, width=148):
    """Row reads that miss an LRU of `cap` rows; `width` consecutive tets touch their rows together
    (own rows first, as the producers do, then neighbours)."""
    lru = OrderedDict()
    miss = 0

    def touch(r):
        nonlocal miss
        if r in lru:
            lru.move_to_end(r)
        else:
            miss += 1
            lru[r] = None
            if len(lru) > cap:
                lru.popitem(last=False)

    n = len(order)
    for s in range(0, n, width):
        grp = order[s:s + width]
        for t in grp:
            touch(int(t))
        for f in range(4):
            for t in grp:
                touch(int(nbr[t, f]))
    return miss / n


def hex_orders(nx, ny, nz):
    hk, hj, hi = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    hi, hj, hk = hi.ravel(), hj.ravel(), hk.ravel()
    hexid = np.arange(nx * ny * nz)            # (k*ny + j)*nx + i, as synthetic.brick_order numbers hexes

    def tets(ho):
        return (ho[:, None] * 6 + np.arange(6)[None, :]).reshape(-1)

    out = {}
    for b in [(2, 2, 2), (4, 4, 4), (7, 7, 4), (7, 7, 7)]:
        keys = np.stack([hi % b[0], hj % b[1], hk % b[2], hi // b[0], hj // b[1], hk // b[2]], 0)
        out["brick %dx%dx%d" % b] = tets(hexid[np.lexsort(keys)])
    # pencils: a x a hexes in (i, j), swept along k; pencils themselves in serpentine order
    for a in (2, 4, 7):
        pi, pj = hi // a, hj // a
        npj = (ny + a - 1) // a
        pj_s = np.where(pi % 2 == 0, pj, npj - 1 - pj)
        pen = pi * npj + pj_s
        kk = np.where(pen % 2 == 0, hk, nz - 1 - hk)               # sweep back and forth
        keys = np.stack([hi % a, hj % a, kk, pen], 0)
        out["pencil %dx%d along k" % (a, a)] = tets(hexid[np.lexsort(keys)])
    # slabs: a hexes thick in i, swept over (j, k) row by row
    for a in (2, 4):
        sl = hi // a
        jj = np.where(sl % 2 == 0, hj, ny - 1 - hj)
        keys = np.stack([hi % a, hk, jj, sl], 0)
        out["slab %d thick" % a] = tets(hexid[np.lexsort(keys)])

    def morton(i, j, k):
        m = np.zeros_like(i, dtype=np.int64)
        for b in range(6):
            m |= ((i >> b) & 1) << (3 * b) | ((j >> b) & 1) << (3 * b + 1) | ((k >> b) & 1) << (3 * b + 2)
        return m
    out["morton"] = tets(hexid[np.argsort(morton(hi, hj, hk), kind="stable")])
    out["natural (i fastest)"] = tets(hexid)
    return out


def main():
    nx = ny = nz = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    mt = synthetic.periodic_kuhn_tables(nx, ny, nz)
    nbr = np.asarray(mt.nbr)
    orders = hex_orders(nx, ny, nz)
    caps = [160, 200, 230, 260, 300, 360, 420, 480]
    print("row reads per tet (1.0 = every row read once); measured default: 1.93")
    print("%-24s" % "order" + "".join("%8d" % c for c in caps))
    for name, order in orders.items():
        print("%-24s" % name + "".join("%8.2f" % misses_per_tet(nbr, order, c) for c in caps), flush=True)


if __name__ == "__main__":
    main()

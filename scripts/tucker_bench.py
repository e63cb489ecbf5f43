"""Timing of the device Tucker step (vt_step_tucker) on a periodic Kuhn box: C2-like
(11^3 velocity grid, comprErr 1e-6) and C5-like (48^3, max rank 8).  Prints one JSON line per case."""
import argparse
import json
import time

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import vlasovtucker_b200 as vtb
from vlasovtucker_b200 import synthetic


def run(hexes, nv, eps, max_rank, steps):
    mt = synthetic.periodic_kuhn_tables(*hexes, (1.0, 1.0, 1.0), brick=(4, 4, 4))
    n = (nv, nv, nv)
    vmin, vmax = [-3.0] * 3, [3.0] * 3
    ax = np.linspace(-3, 3, nv)
    g = np.exp(-0.5 * ax ** 2)
    mx = (g[:, None, None] * g[None, :, None] * g[None, None, :]).ravel(order="F")
    x = (np.arange(mt.nTets) % 97) / 97.0
    f = (1 + 0.2 * np.sin(2 * np.pi * x))[:, None] * mx[None, :]
    ctx = vtb.Context(0)
    ctx.mesh_upload(mt)
    sp = ctx.species_create(n, vmin, vmax, 1.0, 1.0)
    ctx.set_face_bc(sp, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
    ctx.tucker_enable(sp, eps, max_rank)
    ctx.tucker_set_pdf(sp, f)
    ctx.field_set(0.1 * np.random.default_rng(0).standard_normal((mt.nTets, 3)))
    ctx.step_tucker(sp, 1e-3)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.step_tucker(sp, 1e-3)
    ctx.sync()
    ms = (time.perf_counter() - t0) / steps * 1e3
    r = ctx.tucker_ranks(sp)
    print(json.dumps(dict(case=f"{mt.nTets} tets x {nv}^3", eps=eps, max_rank=max_rank, ms_per_step=ms,
                          tet_updates_per_s=mt.nTets / ms * 1e3, mean_rank=float(r.mean()), max_rank_seen=int(r.max()))), flush=True)
    ctx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--case", type=int, default=-1, help="run only this case (0, 1 or 2)")
    a = ap.parse_args()
    cases = [((8, 8, 8), 11, 1e-6, 0, a.steps), ((8, 8, 8), 32, 1e-6, 8, max(1, a.steps // 2)),
             ((4, 4, 4), 48, 1e-6, 8, max(1, a.steps // 2))]
    for i, c in enumerate(cases):
        if a.case < 0 or a.case == i:
            run(*c)

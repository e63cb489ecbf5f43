#!/usr/bin/env python
"""Tuning sweep of the full-format step kernel on one GPU: L2 brick size x planes per CTA."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vlasovtucker_b200 as vtb  # noqa: E402
from vlasovtucker_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hexes", type=int, nargs=3, default=[28, 28, 28])
    ap.add_argument("--nv", type=int, default=32)
    ap.add_argument("--bricks", default="4,4,4;7,7,7;14,14,14;28,28,28;7,7,4;14,7,7")
    ap.add_argument("--chunks", default="1,2,4,8,32")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/sweep_full.jsonl")
    args = ap.parse_args()
    hexes = tuple(args.hexes)
    cfg = bench.c4_setup(hexes, args.nv)
    peak, _ = bench.measured_peak()
    out = open(args.out, "a")
    for bs in args.bricks.split(";"):
        if bs.startswith("p"):             # "pA,D": A x A-hex pencils swept along z, bricks of A x A x D
            a, depth = (int(x) for x in bs[1:].split(","))
            if hexes[0] % a or hexes[1] % a or hexes[2] % depth:
                continue
            brick = ("pencil", a, depth)
            mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"])
            mt.order, mt.brickTets = synthetic.pencil_order(*hexes, a, depth)
        else:
            brick = tuple(int(x) for x in bs.split(","))
            if any(h % b for h, b in zip(hexes, brick)):
                continue
            mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"], brick=brick)
        ctx = vtb.Context(0)
        ctx.mesh_upload(mt)
        sp = ctx.species_create(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
        ctx.set_face_bc(sp, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
        x = mt.tetCentroid[:, 0] / cfg["lengths"][0]
        ctx.set_maxwell(sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * bench.PI * x)), cfg["T"])
        E = np.zeros((mt.nTets, 3))
        E[:, 0] = 1e3 * np.cos(2 * bench.PI * x)
        ctx.field_set(E)
        for cp, var in itertools.product([int(c) for c in args.chunks.split(",")], [int(v) for v in args.variants.split(",")]):
            ctx.step_config(chunk_planes=cp, variant=var)
            for _ in range(2):
                ctx.step_full(sp, cfg["dt"])
            ctx.profile_begin()
            for _ in range(args.steps):
                ctx.step_full(sp, cfg["dt"])
            region, kern, n = ctx.profile_end()
            ms = kern / n
            gbs = 16.0 * mt.nTets * args.nv ** 3 / (ms * 1e-3) / 1e9
            rec = dict(hexes=hexes, brick=brick, chunk_planes=cp, variant=var, kernel_ms=ms, step_ms=region / args.steps,
                       algo_gbs=gbs, frac=gbs / peak)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            out.flush()
        ctx.close()


if __name__ == "__main__":
    main()

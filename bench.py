#!/usr/bin/env python
"""bench.py — headline benchmark of the full-format kinetic update (config C4 of SURVEY.md §8d).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm

Metric (BASELINE.json): cell x v-node updates per second per step.  A "step" is one
Solver<Full>::_UpdatePDF pass (src/solver.cpp:141-212: upwind face fluxes + E-field
acceleration + explicit Euler) plus the Density() moment the next step needs, over every tet
of the rank's mesh block.  Workload at N GPUs: weak scaling, one 28x28x28-hex Kuhn block
(131,712 tets) x 32^3 velocity nodes per GPU — the per-GPU share of the 1,053,696-tet C4 box.

One JSON line is printed by rank 0.  `value` is timed with the state resident in HBM; `e2e`
is the same step through the host-buffer C-ABI call (E field in from pinned host memory,
Density() out) — what one call of the reference-facing Solver exchanges per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# physical set-up of C4 with the reference's constants (src/constants.h:5-12)
EL_MASS, EL_CHARGE, EPS0, KB, EV = 9.1e-31, 1.6e-19, 8.85e-12, 1.38e-23, 11604.518
PI = 3.14159265358979323846
ALGO_BYTES_PER_UPDATE = 16.0     # read f^n once + write f^{n+1} once


def c4_setup(hexes, nv):
    T = 1 * EV
    dens = 1e17
    debye = np.sqrt(EPS0 * KB * T / dens) / EL_CHARGE          # particle_data.cpp:155-157
    wp = EL_CHARGE * np.sqrt(dens / (EL_MASS * EPS0))          # particle_data.cpp:159-162
    vth = np.sqrt(KB * T / EL_MASS)
    cell = 50 * debye / 28.0                                   # 28 hexes per 50 Debye lengths
    lengths = tuple(cell * h for h in hexes)
    return dict(T=T, dens=dens, lengths=lengths, vmin=[-6 * vth] * 3, vmax=[6 * vth] * 3, n=[nv] * 3,
                dt=1e-3 * 2 * PI / wp, mass=EL_MASS, charge=-EL_CHARGE)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    smax.append(float(c[2]))
                except ValueError:
                    continue
                for name, val in zip(names, c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of the step kernel from the committed ncu capture of the default
    bench command (profiles/k_full_step_traffic.json, written from an `ncu --set full` report by
    scripts/ncu_summary.py), if it is for the kernel this run uses."""
    p = os.path.join(ROOT, "profiles", "k_full_step_traffic.json")
    try:
        d = json.load(open(p))
        return d if d.get("kernel") == kernel else None
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def cpu_baseline(steps=None, warmup=1, hexes=(8, 8, 8), nv=32, threads=None, fused=False, keep=None, budget_s=12.0):
    """The oracle's reference-faithful restatement of _UpdatePDF (one temporary per tensor
    operator, OpenMP over tets as src/solver.cpp:159,187,204) on a bounded sample of C4, timed with
    whichever build of the oracle this process loaded (see cpu_baseline_native)."""
    threads = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import oracle
    if not os.environ.get("VT_ORACLE_SO"):
        oracle.build()
    from vlasovtucker_b200 import synthetic
    cfg = c4_setup(hexes, nv)
    nodes, tets, tris, ents = synthetic.kuhn_box(*hexes, cfg["lengths"])
    m = oracle.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    s = oracle.Sim(m)
    sp = s.add_species(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
    x = m.tetCentroid[:, 0] / cfg["lengths"][0]
    s.set_maxwell(sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x)), cfg["T"])
    s.set_params(sp, cfg["dt"], fused=fused)
    E = np.zeros((m.nTets, 3))
    E[:, 0] = 1e3 * np.cos(2 * PI * x)
    tw = time.perf_counter()
    for _ in range(warmup):
        s.update_pdf(sp, E)
    tw = (time.perf_counter() - tw) / max(1, warmup)
    if steps is None:   # bounded sample
        steps = int(min(40, max(2, round(budget_s / max(tw, 1e-3)))))
    t0 = time.perf_counter()
    for _ in range(steps):
        s.update_pdf(sp, E)
    dt = (time.perf_counter() - t0) / steps
    updates = m.nTets * nv ** 3
    if keep is not None:    # hand the oracle's end state to the parity check (bench line key "parity")
        keep.update(sim=s, species=sp, mesh=m, steps_done=warmup + steps, cfg=cfg, E=E, hexes=hexes, nv=nv)
    return dict(value=updates / dt, unit="updates/s", cores=threads, kind="port",
                sample=f"Kuhn box {hexes[0]}x{hexes[1]}x{hexes[2]}x6={m.nTets} tets x {nv}^3, "
                       f"{steps} timed _UpdatePDF steps ({'fused single-pass' if fused else 'reference-faithful temporaries'}, "
                       f"OpenMP {threads} threads)", ms_per_step=dt * 1e3, steps=steps, warmup=warmup)


def cpu_baseline_native(steps=None, warmup=1, budget_s=10.0):
    """CPU numbers with the reference's own compiler flags (CMakeLists.txt:6: -O3 -fno-math-errno
    -march=native, default contraction) — the oracle rebuilt ON THIS MACHINE (oracle/_native/) and timed
    in a child process: the reference-faithful mode (the reference's allocation pattern) and the fused
    single-pass mode (what a careful CPU implementation would do).  Returns None when that build is not
    possible here (no compiler on the box); the caller then falls back to the portable checker build."""
    try:
        import oracle
        so = oracle.build_native()
    except Exception:
        return None
    out = {}
    for mode in ("faithful", "fused"):
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", mode, "--warmup", str(warmup)]
        if steps is not None:
            cmd += ["--steps", str(steps)]
        env = dict(os.environ, VT_ORACLE_SO=so, VT_CPU_BUDGET_S=str(budget_s if mode == "faithful" else budget_s / 2))
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            return None
        out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    return out


def cpu_legs(steps=None, warmup=1, keep=None):
    """cpu_baseline object of the bench line: the reference-flag build where it can be made, the
    portable checker build (x86-64-v3, -ffp-contract=off: the one the parity tests use) beside it."""
    native = cpu_baseline_native(steps=steps, warmup=warmup)
    # the checker build also advances the parity sample (few steps: its job here is the end state)
    port = cpu_baseline(steps=2 if native else steps, warmup=warmup, keep=keep)
    if native is None:
        cb = {k: port[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cb["flags"] = "-O3 -fno-math-errno -march=x86-64-v3 -ffp-contract=off -fopenmp (portable checker build; native build unavailable)"
        return cb, port
    f, u = native["faithful"], native["fused"]
    cb = {"value": f["value"], "unit": f["unit"], "cores": f["cores"], "kind": "port", "sample": f["sample"],
          "flags": "-O3 -fno-math-errno -march=native -fopenmp (the reference's CMakeLists.txt:6 + OpenMP), built on this host",
          "fused_cpu": {"value": u["value"], "sample": u["sample"]},
          "portable_checker_build": {"value": port["value"], "flags": "-march=x86-64-v3 -ffp-contract=off", "sample": port["sample"]}}
    return cb, f


def parity_vs_oracle(keep, device, variant, chunk_planes):
    """The bench kernel (same variant, same items) against the oracle on the CPU-baseline sample:
    the oracle has just advanced that mesh by `steps_done` steps; do the same on the GPU from the
    same initial condition and compare the end states (tolerance of north_star: rel. L2 <= 1e-10)."""
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    cfg, hexes, m = keep["cfg"], keep["hexes"], keep["mesh"]
    mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"], brick=(4, 4, 4))
    ctx = vtb.Context(device)
    ctx.mesh_upload(mt)
    sp = ctx.species_create(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
    ctx.set_face_bc(sp, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
    x = m.tetCentroid[:, 0] / cfg["lengths"][0]
    ctx.set_maxwell(sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x)), cfg["T"])
    ctx.field_set(keep["E"])
    ctx.step_config(chunk_planes=chunk_planes, variant=variant)
    for _ in range(keep["steps_done"]):
        ctx.step_full(sp, cfg["dt"])
    s, osp = keep["sim"], keep["species"]
    fo = s.get_pdf(osp)
    fg = ctx.get_pdf(sp)
    ef = float(np.linalg.norm((fg - fo).ravel()) / np.linalg.norm(fo.ravel()))
    do, dg = s.density(osp), ctx.density(sp)
    ed = float(np.linalg.norm(dg - do) / np.linalg.norm(do))
    ctx.close()
    return {"rel_l2_f": ef, "rel_l2_density": ed, "steps": int(keep["steps_done"]), "tolerance": 1e-10,
            "ok": bool(ef <= 1e-10 and ed <= 1e-10),
            "what": f"bench kernel (variant {variant}, chunk_planes {chunk_planes}) vs the CPU oracle on the cpu_baseline sample "
                    f"({hexes[0]}x{hexes[1]}x{hexes[2]}x6 tets x {keep['nv']}^3), same initial condition and field"}


# ---------------------------------------------------------------------------------------------------
# Tucker format (config C5 of SURVEY.md §8d): per-cell Tucker update with rounding after every face,
# after the acceleration term and after the Euler update (src/solver.cpp:141-212 with T = Tucker,
# src/tucker.cpp:66-98).  value = cell x v-node updates/s as for the full format; the bounding roofline
# is FP64: flops counted with the formula of SURVEY.md §8d for the reference's algorithm, divided by
# the DFMA peak measured in the same run (vt_measure_dfma_peak).

def c5_terms(nv, rank, vth):
    """Velocity-space factors of `rank` shifted, anisotropic Maxwellians (multilinear rank = rank)."""
    ax = np.linspace(-6 * vth, 6 * vth, nv)
    k = np.arange(rank)
    a = []
    for j in range(3):
        shift = vth * 1.5 * np.cos(2 * PI * (k + 0.37 * j) / max(rank, 1) + j)
        width = vth * (0.8 + 0.35 * ((k + j) % 3))
        a.append(np.exp(-0.5 * ((ax[None, :] - shift[:, None]) / width[:, None]) ** 2))
    return a


def c5_init(cfg, rank):
    vth = np.sqrt(KB * cfg["T"] / EL_MASS)
    a0, a1, a2 = c5_terms(cfg["n"][0], rank, vth)
    norm = (a0.sum(1) * a1.sum(1) * a2.sum(1)) * np.prod([(12 * vth) / (cfg["n"][j] - 1) for j in range(3)])

    def init(ctx, sp, x):
        k = np.arange(rank)
        amp = cfg["dens"] / rank * (1 + 0.01 * np.sin(2 * PI * x[:, None] + k[None, :])) / norm[None, :]
        ctx.set_separable(sp, amp, a0, a1, a2)
    return init


def tucker_flops_survey(n, r, rmax):
    """Flops of one tet-step of the REFERENCE's algorithm (six Compress calls, SURVEY.md §8d formula):
    stacked ranks R = r_rhs + 16 r for the four faces, r_rhs + 3 r for the acceleration term, r + r_rhs
    for the Euler update, R' = min(n, R), result rank r' = min(rmax, n)."""
    rp = min(rmax, n)

    def compress(R):
        Rp = min(n, R)
        qr = 3 * 2 * n * R * Rp
        core = 3 * 2 * Rp ** 4
        svd = 3 * (2 * Rp ** 4 + 12 * Rp ** 3)
        proj = 2 * (Rp ** 3 * rp + Rp ** 2 * rp ** 2 + Rp * rp ** 3)
        upd = 3 * 2 * n * Rp * rp
        return qr + core + svd + proj + upd
    total, rrhs = 0, 1
    for _ in range(4):
        total += compress(rrhs + 16 * r)
        rrhs = rp
    total += compress(rrhs + 3 * r)
    total += compress(r + rrhs)
    return float(total)


def tucker_flops_executed(n):
    """Flops of the dense formulation the device kernel executes per tet-step (csrc/tucker.cu): six
    truncated HOSVDs of an n^3 tensor — three Gram matrices (n^4 FMA each) and three projections —
    plus the reconstructions of the own, neighbour and |v.n| tensors (bounded by n^4 FMA each)."""
    gram = 3 * 2 * n ** 4
    proj = 3 * 2 * n ** 4          # upper bound: full-rank projection
    return float(6 * (gram + proj) + 9 * 3 * 2 * n ** 4)


def tucker_flops_slab(n, r):
    """Flops the slab-streaming kernel (csrc/tucker_slab.cu) executes per tet-step on an n^3 grid at rank r, all faces
    paired: its tensor-core contractions counted DMMA by DMMA (512 flops each) — slab expansions, Gram matrices
    (upper triangle of 8x8 blocks), projections — plus ~20 flops per node and rounding for the flux / update arithmetic.
    The eigen-solves (pivoted Cholesky + Jacobi on ~15 x 15 matrices) are latency, not flops, and are left out."""
    p = (n + 7) // 8 * 8
    nb = p // 8
    pairs = nb * (nb + 1) // 2
    k4 = lambda x: (x + 3) // 4
    b8 = lambda x: (x + 7) // 8
    dmma = 0
    for ops in (3, 3, 3, 3, 1, 1):                       # four faces (neighbour, |v.n|, previous rhs / own f), acceleration, Euler update
        ranks = [r, 6, r][:ops] if ops == 3 else [r]
        expand = sum(nb * b8(q) * k4(q) + nb * nb * k4(q) for q in ranks)     # T = U0 M2, slab = T U1^T
        gram01 = 2 * pairs * (p // 4)
        gram2 = pairs * (p // 4)
        proj = b8(r) * nb * (p // 4) + b8(r) * b8(r) * (p // 4)
        dmma += n * (expand + gram01 + proj) + n * gram2
    return float(512 * dmma + 6 * 20 * n ** 3)


def tucker_cpu_baseline(nv, rank, eps, hexes=(1, 1, 1), budget_s=20.0):
    """The oracle's Tucker algebra (oracle/oracle_tucker.cpp: operator+, Hadamard product, QR + HOSVD
    rounding as src/tucker.cpp) on a small Kuhn box with the C5 inputs, all host cores."""
    threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    import oracle
    oracle.build()
    from vlasovtucker_b200 import synthetic
    cfg = c4_setup(hexes, nv)
    nodes, tets, tris, ents = synthetic.kuhn_box(*hexes, cfg["lengths"])
    m = oracle.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    vth = np.sqrt(KB * cfg["T"] / EL_MASS)
    a0, a1, a2 = c5_terms(nv, rank, vth)
    x = m.tetCentroid[:, 0] / cfg["lengths"][0]
    k = np.arange(rank)
    amp = cfg["dens"] / rank * (1 + 0.01 * np.sin(2 * PI * x[:, None] + k[None, :]))
    f = np.einsum("tk,ka,kb,kc->tcba", amp, a0, a1, a2).reshape(m.nTets, -1)
    ts = oracle.TuckerSim(m, cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"], eps, rank)
    ts.set_pdf(f)
    E = np.zeros((m.nTets, 3))
    E[:, 0] = 1e3 * np.cos(2 * PI * x)
    ts.update_pdf(cfg["dt"], E)      # untimed: the first call also tabulates v.n / |v.n| (solver.cpp:258-293, once per run)
    t0 = time.perf_counter()
    ts.update_pdf(cfg["dt"], E)
    one = time.perf_counter() - t0
    steps = int(min(10, max(1, round(budget_s / max(one, 1e-3)))))
    t1 = time.perf_counter()
    for _ in range(steps):
        ts.update_pdf(cfg["dt"], E)
    per = (time.perf_counter() - t1) / steps
    return dict(value=m.nTets * nv ** 3 / per, unit="updates/s", cores=threads, kind="port",
                sample=f"Kuhn box {hexes[0]}x{hexes[1]}x{hexes[2]}x6={m.nTets} tets x {nv}^3, rank {rank}, eps {eps:g}: "
                       f"{steps} timed Tucker _UpdatePDF steps of the oracle (real Tucker algebra, OpenMP {threads} threads)",
                ms_per_step=per * 1e3)


def tucker_cpu_worker_start(nv, rank_t, eps):
    """The oracle's Tucker step costs ~40 core-seconds per tet at 48^3 — start it as a child process at
    the beginning of the run (it uses 6 host cores while the GPU legs, which keep one core busy, run)
    and collect the number at the end."""
    cmd = [sys.executable, os.path.abspath(__file__), "--tucker-cpu-worker", "--tucker-nv", str(nv),
           "--tucker-rank", str(rank_t), "--tucker-eps", repr(eps)]
    try:
        return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def tucker_cpu_worker_collect(proc, timeout=600):
    if proc is None:
        return {"error": "could not start the CPU worker"}
    try:
        out, _ = proc.communicate(timeout=timeout)
        cb = json.loads(out.strip().splitlines()[-1])
        return {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    except Exception as exc:
        try:
            proc.kill()
        except Exception:
            pass
        return {"error": str(exc)[:200]}


def run_tucker(args, rank_world_local, dist, torch, hexes, nv, rank_t, eps, steps, warmup, cpu=True, cpu_proc=None):
    """One Tucker measurement; returns the dict (rank 0) or None."""
    rank, world, local = rank_world_local
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    cfg = c4_setup(hexes, nv)
    init = c5_init(cfg, rank_t)
    brick = tuple(b if h % b == 0 else 1 for h, b in zip(hexes, (4, 4, 4)))
    if world > 1:
        from vlasovtucker_b200 import multigpu
        runner = multigpu.WeakScaledBox(rank, world, local, hexes, cfg, brick=brick, dist=dist, tucker=(eps, rank_t), init=init)
        ctx, sp, mt = runner.ctx, runner.sp, runner.mt
        E = runner.E
    else:
        runner = None
        ctx = vtb.Context(local)
        mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"], brick=brick)
        ctx.mesh_upload(mt)
        sp = ctx.species_create(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
        ctx.set_face_bc(sp, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
        ctx.tucker_enable(sp, eps, rank_t)
        x = mt.tetCentroid[:, 0] / cfg["lengths"][0]
        init(ctx, sp, x)
        E = np.zeros((mt.nTets, 3))
        E[:, 0] = 1e3 * np.cos(2 * PI * x)
        ctx.field_set(E)
    nT, N, dt = mt.nTets, nv ** 3, cfg["dt"]

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        if runner is not None:
            runner.step(dt)
        else:
            ctx.step_tucker(sp, dt)

    for _ in range(warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    ctx.profile_begin()
    for _ in range(steps):
        step()
    region_ms, kern_ms, kern_n = ctx.profile_end()
    launches = ctx.launch_count() - l0
    barrier()
    t_ms = region_ms
    if dist is not None:
        t = torch.tensor([region_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    ms_per_step = t_ms / steps
    ranks = ctx.tucker_ranks(sp)
    which = ctx.tucker_last_kernel(sp)
    # end to end: E from host memory in, Density() out to host memory, every step
    if runner is not None:
        e2e_ms = runner.e2e(dt, max(1, steps // 2), barrier) / max(1, steps // 2)
    else:
        ctx.field_set(E)
        ctx.step_tucker(sp, dt)
        ctx.tucker_density(sp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, steps // 2)):
            ctx.field_set(E)
            ctx.step_tucker(sp, dt)
            ctx.tucker_density(sp)
        ctx.sync()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, steps // 2)
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    peak = ctx.dfma_peak_tflops()
    if runner is not None:
        runner.ctx.close()
    else:
        ctx.close()
    if rank != 0:
        return None
    kern_avg = kern_ms / max(1, kern_n)
    fl_ref = tucker_flops_survey(nv, rank_t, rank_t) * nT
    fl_exe = (tucker_flops_slab(nv, rank_t) if which == "slab" else tucker_flops_executed(nv)) * nT
    out = {
        "metric": "cell x v-node updates/s per step (Tucker format)", "value": nT * N * world / (ms_per_step * 1e-3),
        "unit": "updates/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5: periodic Kuhn box {hexes[0]}x{hexes[1]}x{hexes[2]} hexes x6 = {nT} tets/GPU x {nv}^3 velocity nodes, "
                               f"Tucker format, f = sum of {rank_t} shifted anisotropic Maxwellians, SetMaxRank({rank_t}), comprErr {eps:g}",
                   "tets_per_gpu": nT, "v_nodes": N, "max_rank": rank_t, "compr_err": eps,
                   "mean_rank_after": float(ranks.mean()), "max_rank_after": int(ranks.max()),
                   "tet_updates_per_s": nT * world / (ms_per_step * 1e-3)},
        "roofline": {"bound": "fp64", "achieved": fl_ref / (kern_avg * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": fl_ref / (kern_avg * 1e-3) / 1e12 / peak, "traffic": None,
                     "peak_source": "DFMA probe in this run (vt_measure_dfma_peak)",
                     "kernel": "k_tucker_slab" if which == "slab" else "k_tucker", "kernel_ms": kern_avg,
                     "flops_per_tet_step_counted": fl_ref / nT,
                     "counted": "SURVEY.md §8d formula for the reference's six Compress calls (QR + core transform + 3 SVDs + projection)",
                     "executed_flops_per_tet_step": fl_exe / nT, "executed_tflops": fl_exe / (kern_avg * 1e-3) / 1e12,
                     "executed": ("slab-streaming kernel csrc/tucker_slab.cu: its DMMA contractions counted one by one (slab expansions, "
                                  "upper-triangle Gram blocks, projections) + ~20 flops per node and rounding" if which == "slab" else
                                  "dense formulation of csrc/tucker.cu: 6 x (3 Gram matrices + 3 projections) + reconstructions"),
                     "kernel_share_of_step": kern_ms / region_ms},
        "e2e": {"value": nT * N * world / (e2e_ms * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": 3 * nT * 8 * world,
                "d2h_bytes_per_step": nT * 8 * world, "ms_per_step": e2e_ms,
                "what": "vt_field_set (E from host) + vt_step_tucker + vt_tucker_density (to host); compressed state stays resident"},
        "gpu_launches": int(launches),
    }
    if cpu and world == 1:
        out["cpu_baseline"] = tucker_cpu_worker_collect(cpu_proc if cpu_proc is not None else tucker_cpu_worker_start(nv, rank_t, eps))
    return out


def workload_name(hexes, nv):
    """The workload both arms report (config.workload)."""
    nT = 6 * hexes[0] * hexes[1] * hexes[2]
    return (f"C4 per-GPU share: periodic Kuhn box {hexes[0]}x{hexes[1]}x{hexes[2]} hexes x6 = {nT} tets/GPU "
            f"x {nv}^3 velocity nodes, full format, electrons (Maxwellian 1 eV, 1% density wave), dt=1e-3 T_p")


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cb, timed = cpu_legs(steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": "cell x v-node updates/s per step (full format)", "value": cb["value"],
        "unit": "updates/s", "n_gpus": args.gpus, "steps": timed["steps"], "warmup": timed["warmup"],
        "ms_per_step": timed["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(tuple(args.hexes), args.nv),
                   "cpu_arm": "timed on a bounded sample of the workload (see cpu_baseline.sample)",
                   "note": "reference itself cannot be compiled here (vendored Eigen lacks Eigen/Core); this is the oracle's line-by-line restatement"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import torch
    import vlasovtucker_b200 as vtb
    from vlasovtucker_b200 import synthetic
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    hexes = tuple(args.hexes)
    if args.scaling == "strong":
        # the whole C4 box (56^3 hexes = 1,053,696 tets) divided over the ranks; fits from 4 GPUs up
        from vlasovtucker_b200 import partition as part_
        grid = part_.rank_grid(world)
        if any(g % r for g, r in zip(args.global_hexes, grid)):
            raise SystemExit("--global-hexes must be divisible by the rank grid")
        hexes = tuple(g // r for g, r in zip(args.global_hexes, grid))
        need = 2 * 6 * hexes[0] * hexes[1] * hexes[2] * args.nv ** 3 * 8
        if need > 170e9:
            raise SystemExit(f"strong scaling: {need / 1e9:.0f} GB of state per GPU does not fit (use >= 4 GPUs for the C4 box)")
    nv = args.nv
    cfg = c4_setup(hexes, nv)
    tucker_cpu = None
    if rank == 0 and world == 1 and not args.no_tucker and not args.no_cpu_baseline:
        tucker_cpu = tucker_cpu_worker_start(args.tucker_nv, args.tucker_rank, args.tucker_eps)
    if world > 1:
        from vlasovtucker_b200 import multigpu
        runner = multigpu.WeakScaledBox(rank, world, local, hexes, cfg, brick=tuple(args.brick), dist=dist)
        ctx, sp, mt = runner.ctx, runner.sp, runner.mt
    else:
        runner = None
        ctx = vtb.Context(local)
        mt = synthetic.periodic_kuhn_tables(*hexes, cfg["lengths"], brick=tuple(args.brick))
        if args.pencil:                    # experiment: pencil sweep instead of bricks (N=1 only)
            mt.order, mt.brickTets = synthetic.pencil_order(*hexes, args.pencil[0], args.pencil[1])
        ctx.mesh_upload(mt)
        sp = ctx.species_create(cfg["n"], cfg["vmin"], cfg["vmax"], cfg["mass"], cfg["charge"])
        ctx.set_face_bc(sp, np.full((mt.nTets, 4), vtb.PBC["Periodic"], np.uint8))
        x = mt.tetCentroid[:, 0] / cfg["lengths"][0]
        ctx.set_maxwell(sp, cfg["dens"] * (1 + 0.01 * np.sin(2 * PI * x)), cfg["T"])
        E = np.zeros((mt.nTets, 3))
        E[:, 0] = 1e3 * np.cos(2 * PI * x)
        ctx.field_set(E)
    ctx.step_config(chunk_planes=args.chunk_planes, variant=args.variant)
    nT = mt.nTets
    N = nv ** 3
    updates_rank = nT * N
    dt = cfg["dt"]

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        if runner is not None:
            runner.step(dt)
        else:
            ctx.step_full(sp, dt)

    # ---- device-resident timing
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.profile_begin()
    for _ in range(args.steps):
        step()
    region_ms, kern_ms, kern_n = ctx.profile_end()
    launches = ctx.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = region_ms
    if dist is not None:
        t = torch.tensor([region_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    ms_per_step = t_ms / args.steps
    value = updates_rank * world / (ms_per_step * 1e-3)

    # ---- end to end through host buffers (E in from host, Density() out to host)
    if runner is not None:
        e2e_ms = runner.e2e(dt, args.steps, barrier)
    else:
        Eh = E.copy()
        dens = np.empty(nT)
        ctx.step_full_host(sp, dt, Eh, dens)
        barrier()
        t0 = time.perf_counter()
        ctx.profile_begin()
        for _ in range(args.steps):
            ctx.step_full_host(sp, dt, Eh, dens)
        e2e_region_ms, _, _ = ctx.profile_end()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(e2e_region_ms, wall_ms)     # host packing is part of the call: take the slower clock
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = updates_rank * world / (e2e_ms / args.steps * 1e-3)

    # ---- the whole loop body of Solver::Solve (solver.cpp:91-133) on the same mesh: charge density,
    # Poisson solve (periodic, pinned row, Jacobi-PCG to 2.2e-16), step — reported beside the metric
    coupled = None
    if not args.no_coupled:
        try:
            if runner is not None:
                from vlasovtucker_b200 import multigpu
                multigpu.PartitionedPoisson(ctx, runner.lp, dist, np.full((nT, 4), vtb.QBC["Periodic"], np.uint8),
                                            any_dirichlet=False)
            else:
                ctx.poisson_setup(np.full((nT, 4), vtb.QBC["Periodic"], np.uint8))
            bg = np.full(nT, -cfg["charge"] * cfg["dens"])

            def loop_body():
                ctx.charge_density([sp], bg)
                ctx.poisson_solve(download=False)
                step()

            for _ in range(2):
                loop_body()
            barrier()
            nc = max(3, args.steps // 4)
            t0 = time.perf_counter()
            for _ in range(nc):
                loop_body()
            barrier()
            loop_ms = (time.perf_counter() - t0) * 1e3 / nc
            if dist is not None:
                t = torch.tensor([loop_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                loop_ms = float(t.item())
            its, res = ctx.poisson_stats()
            coupled = {"ms_per_iteration": loop_ms, "poisson_ms": loop_ms - ms_per_step, "pcg_iterations": int(its),
                       "pcg_rel_residual": float(res), "iterations_timed": nc, "rows": nT * world,
                       "ratio_to_step_alone": loop_ms / ms_per_step,
                       "what": "vt_charge_density + vt_poisson_solve (fields stay on the device"
                               + (", rows partitioned like the tets, in-kernel reductions over the ranks" if world > 1 else "")
                               + ") + vt_step_full" + (" + halo barrier" if world > 1 else "")}
        except Exception as exc:   # the headline metric must not depend on this extra
            coupled = {"error": str(exc)[:300]}

    # ---- Tucker format beside it (BASELINE.json's metric names both formats): C5 per-GPU share, same
    # rank grid, after the full-format state has been released
    tucker = None
    if not args.no_tucker:
        if runner is not None:
            runner.ctx.close()
        else:
            ctx.close()
        try:
            tucker = run_tucker(args, (rank, world, local), dist, torch, tuple(args.tucker_hexes), args.tucker_nv,
                                args.tucker_rank, args.tucker_eps, max(3, args.steps // 4), 2,
                                cpu=not args.no_cpu_baseline, cpu_proc=tucker_cpu)
        except Exception as exc:
            tucker = {"error": str(exc)[:300]} if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        kern_avg_ms = kern_ms / max(1, kern_n)
        achieved = ALGO_BYTES_PER_UPDATE * updates_rank / (kern_avg_ms * 1e-3) / 1e9
        bulk = (args.variant & 64 and nv % 2 == 0 and nv * nv * 8 >= 4096) or (args.variant & 16)
        kernel_name = "k_full_step_bulk" if bulk else "k_full_step"
        tr = ncu_traffic(kernel_name)
        line = {
            "metric": "cell x v-node updates/s per step (full format)", "value": value, "unit": "updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(hexes, nv) + (f" (strong scaling: the {args.global_hexes[0]}x{args.global_hexes[1]}x{args.global_hexes[2]}-hex box over {world} GPUs)" if args.scaling == "strong" else ""),
                "tets_per_gpu": nT, "v_nodes": N, "state_bytes_per_gpu": 2 * nT * N * 8,
                "l2_policy": "inputs larger than L2 (state is %.1f GB per copy); no flush needed" % (nT * N * 8 / 1e9),
                "brick_hexes": list(args.brick) if not (args.pencil and world == 1) else None,
                "pencil_hexes": args.pencil if world == 1 else None, "chunk_planes": args.chunk_planes, "kernel_variant": args.variant,
                "kernel": kernel_name,
                "step": "K1 full_step (flux+accel+Euler+density partials) + density reduce" + ("; halo push over NVLink" if world > 1 else ""),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if tr is None else tr.get("dram_bytes_per_launch"),
                         "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": kern_avg_ms,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * updates_rank,
                         "kernel_share_of_step": kern_ms / region_ms},
            "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": 3 * nT * 8 * world,
                    "d2h_bytes_per_step": nT * 8 * world, "ms_per_step": e2e_ms / args.steps,
                    "what": "vt_step_full_host: E (3 doubles/tet) host->device, step, Density() device->host; state stays resident"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if coupled is not None:
            line["coupled_loop"] = coupled
        if tucker is not None:
            line["tucker"] = tucker
        if not args.no_cpu_baseline and world == 1:
            keep = {}
            cb, _ = cpu_legs(keep=keep)
            line["cpu_baseline"] = cb
            try:
                line["parity"] = parity_vs_oracle(keep, local, args.variant, args.chunk_planes)
            except Exception as exc:
                line["parity"] = {"error": str(exc)[:300]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_gpu_tucker(args):
    """`--format tucker`: the C5 line as the main JSON line."""
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    line = run_tucker(args, (rank, world, local), dist, torch, tuple(args.tucker_hexes), args.tucker_nv, args.tucker_rank,
                      args.tucker_eps, args.steps, args.warmup, cpu=not args.no_cpu_baseline)
    clocks = sampler.stop() if rank == 0 else None
    sweep = []
    if args.tucker_sweep:
        for r in (4, 8, 12, 16):
            for eps in (1e-4, 1e-6, 1e-8):
                pt = run_tucker(args, (rank, world, local), dist, torch, (8, 8, 8), args.tucker_nv, r, eps, 2, 1, cpu=False)
                if rank == 0:
                    sweep.append({"max_rank": r, "compr_err": eps, "tets_per_gpu": pt["config"]["tets_per_gpu"],
                                  "ms_per_step": pt["ms_per_step"], "value": pt["value"], "mean_rank_after": pt["config"]["mean_rank_after"],
                                  "roofline_frac": pt["roofline"]["frac"], "executed_tflops": pt["roofline"]["executed_tflops"],
                                  "kernel": pt["roofline"]["kernel"]})
    if rank == 0:
        line["vs_baseline"] = None
        line["clocks"] = clocks
        if sweep:
            line["sweep"] = sweep
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hexes", type=int, nargs=3, default=[28, 28, 28], help="hexes per GPU block")
    ap.add_argument("--nv", type=int, default=32)
    ap.add_argument("--brick", type=int, nargs=3, default=[4, 4, 4], help="L2 brick in hexes")
    ap.add_argument("--pencil", type=int, nargs=2, default=None, metavar=("A", "DEPTH"),
                    help="experiment: order tets in AxA-hex pencils swept along z (bricks of AxAxDEPTH)")
    ap.add_argument("--chunk-planes", type=int, default=0, help="i2-planes per work item (0 = whole tensor)")
    ap.add_argument("--variant", type=int, default=64,
                    help="vt_step_config variant bits; 64 = the library's own choice (bulk-copy pipeline, upwind-select "
                         "arithmetic for 32^3), 2 = register-staged kernel")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --hexes per GPU (default); strong: --global-hexes divided over the GPUs (>= 4 GPUs for C4)")
    ap.add_argument("--global-hexes", type=int, nargs=3, default=[56, 56, 56], help="--scaling strong: hexes of the whole box")
    ap.add_argument("--format", default="full", choices=["full", "tucker"],
                    help="full: the headline line (with a `tucker` object beside it); tucker: the C5 Tucker line alone")
    ap.add_argument("--tucker-hexes", type=int, nargs=3, default=[16, 16, 16], help="C5: hexes per GPU block")
    ap.add_argument("--tucker-nv", type=int, default=48)
    ap.add_argument("--tucker-rank", type=int, default=8, help="C5: rank of the initial condition = SetMaxRank")
    ap.add_argument("--tucker-eps", type=float, default=1e-6, help="C5: comprErr")
    ap.add_argument("--tucker-sweep", action="store_true",
                    help="--format tucker: r in {4,8,12,16} x eps in {1e-4,1e-6,1e-8} on a reduced box, one object per point in `sweep`")
    ap.add_argument("--no-tucker", action="store_true", help="skip the Tucker leg of the default line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-coupled", action="store_true", help="skip the coupled-loop (Poisson + step) measurement")
    ap.add_argument("--cpu-worker", default=None, choices=["faithful", "fused"], help=argparse.SUPPRESS)
    ap.add_argument("--tucker-cpu-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.tucker_cpu_worker:
        print(json.dumps(tucker_cpu_baseline(args.tucker_nv, args.tucker_rank, args.tucker_eps)), flush=True)
        return
    if args.cpu_worker:     # child of cpu_baseline_native: time one mode with the library VT_ORACLE_SO names
        explicit = any(a == "--steps" for a in sys.argv)
        cb = cpu_baseline(steps=args.steps if explicit else None, warmup=max(1, args.warmup), fused=args.cpu_worker == "fused",
                          budget_s=float(os.environ.get("VT_CPU_BUDGET_S", "10")))
        print(json.dumps(cb), flush=True)
        return
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.format == "tucker":
        run_gpu_tucker(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

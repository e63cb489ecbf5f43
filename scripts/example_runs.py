"""Wall time per iteration of the reference's example configurations (C1s oscillations Full and
Tucker, C3 sheath with two species) through the C++ host classes on the GPU: difference of two run
lengths of tests/cpp/host_parity.cpp (the reference's drivers with a dump at the end), so that mesh
loading and set-up cancel."""
import json
import os
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "vlasovtucker_b200", "build", "host_parity")


def run(case, mesh, iters):
    t0 = time.perf_counter()
    subprocess.run([BIN, case, os.path.join(ROOT, "tests", "data", mesh), str(iters), f"/tmp/ex_{case}.bin"],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0


for case, mesh, a, b in [("oscillations", "rectangle_fine.msh", 20, 220),
                         ("oscillations_tucker", "fully_periodic_coarse.msh", 5, 45),
                         ("sheath", "rectangle_fine.msh", 20, 220)]:
    run(case, mesh, 2)   # warm the driver / page cache
    ta, tb = run(case, mesh, a), run(case, mesh, b)
    print(json.dumps(dict(case=f"{case} ({mesh})", iterations=b - a, ms_per_iteration=(tb - ta) / (b - a) * 1e3,
                          setup_s=ta - a * (tb - ta) / (b - a))), flush=True)

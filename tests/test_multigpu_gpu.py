"""Multi-GPU parity, one process per GPU under torchrun (CUDA-IPC peer mappings): the partitioned update
with the halo exchange fused into the step kernel reproduces the single-GPU update bit for bit, the
coupled loop with the partitioned Poisson solve matches one GPU.  On a box with a single GPU the same
partitions, kernels, push lists and device-side barriers run as virtual ranks of one process
(tests/test_virtual_ranks_gpu.py) — the tests below then run that form instead of being skipped."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _virtual(test_name, *args):
    """One GPU: the virtual-rank form of the same check."""
    import vlasovtucker_b200 as vtb
    import test_virtual_ranks_gpu as tv
    vtb.capi.load()
    getattr(tv, test_name)(vtb, *args)


@pytest.mark.parametrize("variant", ["0", "2", "18"])
def test_partitioned_update_bit_identical(variant):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        return _virtual("test_virtual_ranks_full_bit_identical", 2, "rcb", (16, 8, 8), 2, int(variant))
    world = 4 if n >= 4 else 2
    env = dict(os.environ, VT_VARIANT=variant)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_CHECK_OK" in r.stdout


def test_partitioned_coupled_loop(oracle_mod, monkeypatch):
    """Density -> partitioned Poisson -> partitioned step -> wall charge over ranks, against one GPU."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        return _virtual("test_virtual_ranks_coupled_loop", oracle_mod, monkeypatch)
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "mgpu_loop_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_LOOP_OK" in r.stdout


def test_partitioned_tucker_bit_identical():
    """Tucker species: slots of boundary tets mirrored into the peers' ghost rows by the step kernel."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        return _virtual("test_virtual_ranks_tucker_bit_identical", (12, 10, 8), 0, "general")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29521", os.path.join(ROOT, "scripts", "mgpu_tucker_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_TUCKER_OK" in r.stdout

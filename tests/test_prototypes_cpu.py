"""Keeps the CPU-checked prototypes under scripts/prototypes/ honest: the device-style leading-eigenpair
solver planned for the Tucker rounding (eig_leading.h) against numpy on its whole case list."""
import importlib.util
import os

from conftest import ROOT


def _load(name):
    path = os.path.join(ROOT, "scripts", "prototypes", name + ".py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_leading_eigenpairs_prototype_matches_numpy(capsys):
    mod = _load("eig_leading_check")
    mod.main()
    out = capsys.readouterr().out
    assert out.startswith("cases 270")


def test_round_robin_jacobi_statement_matches_numpy(capsys):
    """The parallel two-sided Jacobi method of csrc/tucker_slab.cu, stated on the CPU (round-robin pairing, 2x2 block
    updates with mirrored writes): eigenvalues, orthogonality and residuals against numpy on Gram matrices."""
    mod = _load("jacobi_round_robin")
    mod.main(sizes=(17, 12))
    assert capsys.readouterr().out.rstrip().endswith("ok")


def test_single_reduction_cg_reaches_the_reference_stopping_rule(capsys):
    """scripts/prototypes/cg_single_reduction.py: the Chronopoulos-Gear form of PCG (one global reduction per iteration,
    the planned next step of csrc/poisson.cu) meets Eigen's |r| <= eps |b| on the oracle's Poisson matrix in as many
    iterations as the classic form and gives the same solution."""
    mod = _load("cg_single_reduction")
    mod.main(6)
    out = capsys.readouterr().out.splitlines()
    it_classic = int(out[1].split("iterations")[1].split()[0])
    one = out[2]
    it_one = int(one.split("iterations")[1].split()[0])
    dx = float(one.split("|x - x_classic| / |x|")[1])
    assert "not converged" not in one
    assert abs(it_one - it_classic) <= 3 and dx <= 1e-11

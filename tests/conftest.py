import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; ignored when the plugin is absent)")


def mesh_path(name):
    return os.path.join(DATA, name)


def tables_from_oracle(m, order=None, brick_tets=0):
    """MeshTables (the C-ABI input) from the oracle's restated Mesh — test plumbing only."""
    from vlasovtucker_b200 import MeshTables
    return MeshTables(nbr=m.adj.copy(), area=m.faceArea.copy(), volume=m.tetVolume.copy(),
                      normal=m.faceNormal.copy(), entity=m.faceEntity.copy(),
                      tetCentroid=m.tetCentroid.copy(), faceCentroid=m.faceCentroid.copy(),
                      order=order, brickTets=brick_tets, periodic=list(getattr(m, "periodic", [])))


def rel_l2(a, b):
    a = np.asarray(a, float).ravel()
    b = np.asarray(b, float).ravel()
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def face_bc_arrays(m, spec):
    """Per-face particle-BC arrays from {entity: (kind, collect)} as Solver::SetParticleBC
    (src/solver.cpp:62-71) would assign them; periodic pairs of the mesh get Periodic."""
    from vlasovtucker_b200 import PBC
    bc = np.zeros((m.nTets, 4), np.uint8)
    col = np.zeros((m.nTets, 4), np.uint8)
    for pair in getattr(m, "periodic", []):
        for e in pair:
            bc[m.faceEntity == e] = PBC["Periodic"]
    for e, (kind, collect) in spec.items():
        bc[m.faceEntity == e] = PBC[kind]
        col[m.faceEntity == e] = 1 if collect else 0
    return bc, col


def poisson_bc_arrays(m, spec):
    """Per-face field-BC arrays from {entity: (kind, value, normalGrad)} as PoissonSolver::SetBC
    (src/poisson.cpp:85-92) would assign them; periodic pairs of the mesh get Periodic."""
    from vlasovtucker_b200 import QBC
    bc = np.zeros((m.nTets, 4), np.uint8)
    val = np.zeros((m.nTets, 4))
    ng = np.zeros((m.nTets, 4))
    for pair in getattr(m, "periodic", []):
        for e in pair:
            bc[m.faceEntity == e] = QBC["Periodic"]
    for e, (kind, value, grad) in spec.items():
        sel = m.faceEntity == e
        bc[sel] = QBC[kind]
        val[sel] = value
        ng[sel] = grad
    return bc, val, ng


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle

"""vlasovtucker_b200 — B200-native (sm_100a) implementation of VlasovTucker's per-time-step
kinetic update behind a C ABI (include/vt_b200.h).

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI, ``host/`` the C++ host classes with
the reference's names (Mesh, VelocityGrid, ParticleData, Full, Tucker, PoissonSolver, Solver,
MulticomponentSolver), and this thin ctypes layer used by the tests and ``bench.py``.
There is no CPU fallback: every compute call goes through ``libvt_b200.so``.
"""
from . import build, capi  # noqa: F401
from .context import Context, MeshTables, PBC, QBC  # noqa: F401

"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/vt_b200.h declares, and refuses to run without a GPU (no CPU fallback); plus the
host-side table builders against the oracle's restated Mesh::Reconstruct."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from vlasovtucker_b200 import capi
    decl = _declared_symbols()
    assert len(decl) >= 25
    assert sorted(capi.SIGNATURES) == decl


def test_library_exports_every_declared_symbol():
    from vlasovtucker_b200 import build, capi
    build.build_lib()
    lib = capi.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vt_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vlasovtucker_b200 import capi
    lib = capi.load()
    h = C.c_void_p()
    assert lib.vt_ctx_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.vt_last_error()
    import vlasovtucker_b200 as vtb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vtb.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vlasovtucker_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


@pytest.mark.parametrize("dims,lengths", [((4, 3, 5), (1.0, 0.7, 1.3)), ((3, 3, 3), (1.0, 1.0, 1.0))])
def test_synthetic_kuhn_tables_match_reconstruct(oracle_mod, dims, lengths):
    """The numpy table builder reproduces the restated Mesh::Reconstruct bit for bit: indices,
    adjacency (incl. the sort-based periodic pairing, mesh.cpp:224-303) and geometry."""
    from vlasovtucker_b200 import synthetic
    nodes, tets, tris, ents = synthetic.kuhn_box(*dims, lengths)
    mt = synthetic.periodic_kuhn_tables(*dims, lengths)
    om = oracle_mod.Mesh.from_arrays(nodes, tets, tris, ents, [(1, 2), (3, 4), (5, 6)])
    assert mt.nTets == om.nTets == 6 * dims[0] * dims[1] * dims[2]
    assert np.array_equal(om.adj, mt.nbr)
    assert np.array_equal(om.faceEntity, mt.entity)
    for a, b in [(om.faceArea, mt.area), (om.tetVolume, mt.volume), (om.faceNormal, mt.normal),
                 (om.tetCentroid, mt.tetCentroid), (om.faceCentroid, mt.faceCentroid)]:
        assert np.array_equal(a, b)
    assert abs(mt.volume.sum() - np.prod(lengths)) < 1e-14


def test_brick_order_is_a_permutation_of_compact_bricks():
    from vlasovtucker_b200 import synthetic
    mt = synthetic.periodic_kuhn_tables(6, 6, 6, brick=(3, 3, 3))
    assert sorted(mt.order.tolist()) == list(range(mt.nTets))
    assert mt.brickTets == 6 * 27
    c = mt.tetCentroid[mt.order[:mt.brickTets]]
    assert (c.max(0) - c.min(0)).max() < 0.5 + 1e-12     # first brick spans 3 of 6 cells per axis


def test_pencil_order_sweeps_contiguously():
    """pencil_order: a permutation; every brick is one a x a x depth block of hexes; consecutive bricks
    touch (the sweep never jumps), which is what keeps the rows in flight next to the cached ones."""
    from vlasovtucker_b200 import synthetic
    nx = ny = nz = 6
    mt = synthetic.periodic_kuhn_tables(nx, ny, nz)
    order, bt = synthetic.pencil_order(nx, ny, nz, 2, 3)
    assert sorted(order.tolist()) == list(range(mt.nTets)) and bt == 6 * 2 * 2 * 3
    cen = mt.tetCentroid[order].reshape(-1, bt, 3)
    ext = cen.max(1) - cen.min(1)
    assert (ext[:, :2] < 2 / 6).all() and (ext[:, 2] < 3 / 6).all()
    mid = cen.mean(1)
    step = np.abs(np.diff(mid, axis=0))
    assert (step.max(1) <= 3 / 6 + 1e-12).all()            # next brick is an x, y or z neighbour
    assert ((step > 1e-12).sum(1) == 1).all()
    with pytest.raises(ValueError):
        synthetic.pencil_order(6, 6, 6, 4)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header compiles as C99 and as C++ without any CUDA or
    torch type in a signature."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "vt_b200.h")
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        r = subprocess.run(["/usr/bin/gcc", "-x", lang, std, "-Wall", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)      # declarations only, comments stripped
    for banned in ("cuda", "torch", "at::", "std::", "Tensor"):
        assert banned not in code, banned

"""CPU tests of the C++ host layer: it builds, its Mesh reproduces the oracle's restated
Mesh::Reconstruct index for index, and the reference's own print-only tensor tests give their
known answers when compiled, unchanged, against it (only where /root/reference is present)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, mesh_path

HOST = os.path.join(ROOT, "vlasovtucker_b200", "host")
BUILD = os.path.join(ROOT, "vlasovtucker_b200", "build")
REF = "/root/reference"


@pytest.fixture(scope="module")
def host_lib():
    from vlasovtucker_b200 import build
    build.build_lib()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    subprocess.check_call(["make", "-C", HOST, "-s", "parity"])
    return os.path.join(ROOT, "vlasovtucker_b200", "lib", "libvlasov_tucker.so")


def _mesh_dump(host_lib, tmp_path):
    exe = os.path.join(BUILD, "mesh_dump")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", f"-I{HOST}", f"-I{HOST}/standin", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "mesh_dump.cpp"), f"-L{ROOT}/vlasovtucker_b200/lib",
                           "-lvlasov_tucker", "-lvt_b200", "-Wl,-rpath," + os.path.join(ROOT, "vlasovtucker_b200", "lib")])
    return exe


@pytest.mark.parametrize("name,pairs", [
    ("rectangle_fine.msh", [(1, 2), (3, 4), (5, 6)]),
    ("rectangle_fine.msh", [(3, 4), (5, 6)]),
    ("box_4955_tets.msh", [(5, 6)]),
    ("sphere_2697_tets.msh", []),
    ("simple.msh", []),
])
def test_host_mesh_matches_oracle(oracle_mod, host_lib, tmp_path, name, pairs):
    exe = _mesh_dump(host_lib, tmp_path)
    out = str(tmp_path / "mesh.bin")
    args = [exe, mesh_path(name), out] + [str(x) for p in pairs for x in p]
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    m = oracle_mod.Mesh.load(mesh_path(name), pairs)
    buf = open(out, "rb").read()
    nT, nP = np.frombuffer(buf, np.int32, 2)
    assert (nT, nP) == (m.nTets, m.nPoints)
    off = 8

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off)
        off += a.nbytes
        return a
    assert np.array_equal(take(np.int32, 4 * nT).reshape(nT, 4), m.tets)             # bit-exact indexing
    assert np.array_equal(take(np.int32, 4 * nT).reshape(nT, 4), m.adj)
    assert np.array_equal(take(np.int32, 4 * nT).reshape(nT, 4), m.faceEntity)
    assert np.allclose(take(np.float64, nT), m.tetVolume, rtol=1e-14, atol=0)
    assert np.array_equal(take(np.float64, 4 * nT).reshape(nT, 4), m.faceArea)
    assert np.array_equal(take(np.float64, 12 * nT).reshape(nT, 4, 3), m.faceNormal)
    assert np.array_equal(take(np.float64, 3 * nT).reshape(nT, 3), m.tetCentroid)
    assert np.array_equal(take(np.float64, 12 * nT).reshape(nT, 4, 3), m.faceCentroid)
    order = take(np.int32, nT)
    assert sorted(order.tolist()) == list(range(nT))
    assert abs(take(np.float64, 1)[0] - m.average_cell_size()) < 1e-14


def test_host_mesh_periodic_mismatch_message(host_lib, tmp_path):
    exe = _mesh_dump(host_lib, tmp_path)
    r = subprocess.run([exe, mesh_path("rectangle_fine.msh"), str(tmp_path / "x.bin"), "1", "3"], capture_output=True, text=True)
    assert r.returncode == 3
    assert "Mismatch between the sizes of the periodic planes 1 and  3" in r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_reference_drivers_compile_unchanged_and_known_answers(host_lib):
    subprocess.check_call(["make", "-C", HOST, "-s", "drivers"])
    for exe in ("oscillations", "sheath", "test_tensors", "tucker_test", "poisson_test", "mesh_test"):
        assert os.path.exists(os.path.join(BUILD, exe)), exe
    # test/test_tensors.cpp:10-18: Sum(ones 3^3) = 27, (x+x)*(x+x) = 4 everywhere
    out = subprocess.run([os.path.join(BUILD, "test_tensors")], capture_output=True, text=True, check=True).stdout.split()
    vals = [float(x) for x in out]
    assert vals[:27] == [1.0] * 27 and vals[27] == 27.0 and vals[28:] == [4.0] * 27
    # test/tucker_test.cpp:173-203: ten add/scale/round cycles print the same tensor
    out = subprocess.run([os.path.join(BUILD, "tucker_test")], capture_output=True, text=True, check=True).stdout
    blocks = [np.array([float(x) for x in b.split()]) for b in out.strip().split("\n\n")]
    assert len(blocks) == 10 and all(b.size == 45 for b in blocks)
    for b in blocks[1:]:
        assert np.abs(b - blocks[0]).max() < 1e-5      # printed with 6 significant digits


def test_host_tucker_class_matches_oracle(oracle_mod, host_lib, tmp_path):
    """The user-facing Tucker value class of the host API against the oracle's restated tucker.cpp:
    construction, operator+, scalar and Hadamard products, Compress, Sum — ranks equal, tensors to
    1e-12 (singular-vector signs are implementation defined, so only reconstructions compare)."""
    exe = os.path.join(BUILD, "tucker_dump")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", f"-I{HOST}", f"-I{HOST}/standin", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "tucker_dump.cpp"), f"-L{ROOT}/vlasovtucker_b200/lib",
                           "-lvlasov_tucker", "-lvt_b200", "-Wl,-rpath," + os.path.join(ROOT, "vlasovtucker_b200", "lib")])
    out = subprocess.run([exe, "errors"], capture_output=True, text=True, check=True).stdout
    assert "sum: Different shapes in sum" in out and "mult: Different shapes in mult" in out   # tucker.cpp:193, 262
    n, eps, rmax = (9, 7, 6), 1e-6, 5
    ax = [np.linspace(-1, 1, k) for k in n]

    def smooth(c, w):
        return np.exp(-(((ax[0][:, None, None] - c[0]) / w) ** 2 + ((ax[1][None, :, None] - c[1]) / w) ** 2
                        + ((ax[2][None, None, :] - c[2]) / w) ** 2 + 0.4 * ax[0][:, None, None] * ax[2][None, None, :]))
    a = smooth((0.1, -0.2, 0.3), 0.6) + 0.3 * smooth((-0.5, 0.4, 0.0), 0.4)
    b = smooth((-0.2, 0.1, -0.1), 0.8)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([a.ravel(order="F"), b.ravel(order="F")]).tofile(fin)
    subprocess.check_call([exe, fin, fout] + [str(k) for k in n] + [str(eps), str(rmax)])
    buf = np.fromfile(fout)
    N = a.size
    recs = [(tuple(int(x) for x in buf[i:i + 3]), buf[i + 3:i + 3 + N].reshape(n, order="F"), buf[i + 3 + N])
            for i in range(0, buf.size, N + 4)]
    assert len(recs) == 5
    T = oracle_mod.TuckerObj
    ta, tb = T.from_full(a, eps, rmax), T.from_full(b)
    want = [ta.clone()]
    s = ta.clone().axpy(0.7, tb)
    want.append(s.clone())
    want.append(s.clone().compress(eps, rmax))
    h = ta.clone().hadamard(tb)
    want.append(h.clone())
    want.append(h.clone().axpy(-0.25, ta).compress(eps, rmax))
    for k, ((ranks, rec, total), w) in enumerate(zip(recs, want)):
        assert ranks == w.ranks(), k
        ref = w.reconstructed()
        assert np.abs(rec - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), k
        assert abs(total - w.sum()) <= 1e-11 * max(1.0, abs(w.sum())), k


def _read_mesh_dump(path):
    buf = open(path, "rb").read()
    nT, nP = np.frombuffer(buf, np.int32, 2)
    off = 8

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off)
        off += a.nbytes
        return a
    d = dict(nT=int(nT), nP=int(nP))
    d["tets"] = take(np.int32, 4 * nT).reshape(nT, 4)
    d["nbr"] = take(np.int32, 4 * nT).reshape(nT, 4)
    d["entity"] = take(np.int32, 4 * nT).reshape(nT, 4)
    d["volume"] = take(np.float64, nT)
    d["area"] = take(np.float64, 4 * nT).reshape(nT, 4)
    d["normal"] = take(np.float64, 12 * nT).reshape(nT, 4, 3)
    return d


def test_host_mesh_reads_the_synthetic_c4_file(host_lib, tmp_path):
    """The C4 input path: the Kuhn box written as MSH 2.2 ASCII, read by the host
    `Mesh(std::string)` + `SetPeriodicBounaries` + `Reconstruct`, gives exactly the tables the bench
    builds directly (synthetic.periodic_kuhn_tables) — neighbours, entities, volumes, areas, normals —
    and the load scales linearly (a 24^3-hex, 82,944-tet file in seconds, not minutes)."""
    import time
    from vlasovtucker_b200 import synthetic
    exe = _mesh_dump(host_lib, tmp_path)
    times = {}
    for nh in (6, 24):
        nodes, tets, tris, ents = synthetic.kuhn_box(nh, nh, nh)
        msh, out = str(tmp_path / f"kuhn{nh}.msh"), str(tmp_path / f"kuhn{nh}.bin")
        synthetic.write_msh(msh, nodes, tets, tris, ents)
        t0 = time.perf_counter()
        subprocess.check_call([exe, msh, out, "1", "2", "3", "4", "5", "6"])
        times[nh] = time.perf_counter() - t0
        d = _read_mesh_dump(out)
        mt = synthetic.periodic_kuhn_tables(nh, nh, nh)
        assert d["nT"] == 6 * nh ** 3 and np.array_equal(d["tets"], tets)
        assert np.array_equal(d["nbr"], mt.nbr)
        assert np.array_equal(d["entity"], mt.entity)
        assert np.array_equal(d["volume"], mt.volume)
        assert np.array_equal(d["area"], mt.area)
        assert np.array_equal(d["normal"], mt.normal)
    assert times[24] < 60.0
    assert times[24] < 64 * 4 * max(times[6], 0.05)     # 64x the tets: no worse than ~linear with slack

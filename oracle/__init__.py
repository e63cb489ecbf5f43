"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so``, the Eigen-free CPU restatement of the
reference's per-time-step kinetic update (see ``oracle/oracle.h`` for the parity status
and the file:line citations).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product package ``vlasovtucker_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PBC = dict(NonBoundary=0, Periodic=1, Source=2, Absorbing=3, Free=4)   # solver.h:25
QBC = dict(NonBoundary=0, Neumann=1, Dirichlet=2, Periodic=3)           # poisson.h:47

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_native():
    """The timing build with the reference's flags (-march=native), made on the machine that runs it."""
    subprocess.check_call(["make", "-C", _HERE, "_native/liboracle_native.so"], stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "_native", "liboracle_native.so")


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("VT_ORACLE_SO") or os.path.join(_HERE, "liboracle.so")   # bench.py's timing build
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_mesh_load", "orc_mesh_from_arrays", "orc_poisson_create", "orc_sim_create",
                     "orc_tucker_from_full", "orc_tucker_clone"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_void_p
        for name in ("orc_mesh_average_cell_size", "orc_poisson_last_error", "orc_sim_wall_charge",
                     "orc_sim_wall_area", "orc_tucker_sum", "orc_tucker_norm"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_double
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _check(rc):
    if rc:
        raise RuntimeError(lib().orc_last_error().decode())


class Mesh:
    """Restated ``Mesh`` after ``Reconstruct`` (mesh.cpp:94-303), flattened to arrays."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(lib().orc_last_error().decode())
        self.h = C.c_void_p(handle)
        L = lib()
        nt = self.nTets = L.orc_mesh_ntets(self.h)
        npnt = self.nPoints = L.orc_mesh_npoints(self.h)
        self.points = np.zeros((npnt, 3))
        self.tets = np.zeros((nt, 4), np.int32)
        self.tetCentroid = np.zeros((nt, 3))
        self.tetVolume = np.zeros(nt)
        self.facePoints = np.zeros((nt, 4, 3), np.int32)
        self.faceNormal = np.zeros((nt, 4, 3))
        self.faceCentroid = np.zeros((nt, 4, 3))
        self.faceArea = np.zeros((nt, 4))
        self.faceEntity = np.zeros((nt, 4), np.int32)
        self.faceBoundary = np.zeros((nt, 4), np.uint8)
        self.adj = np.zeros((nt, 4), np.int32)
        L.orc_mesh_get(self.h, _d(self.points), _i(self.tets), _d(self.tetCentroid), _d(self.tetVolume),
                       _i(self.facePoints), _d(self.faceNormal), _d(self.faceCentroid), _d(self.faceArea),
                       _i(self.faceEntity), self.faceBoundary.ctypes.data_as(C.POINTER(C.c_ubyte)),
                       _i(self.adj))

    @classmethod
    def load(cls, path, periodic=(), scale=1.0):
        pairs = np.ascontiguousarray(np.array(list(periodic), np.int32).reshape(-1, 2))
        m = cls(lib().orc_mesh_load(path.encode(), _i(pairs), len(pairs), C.c_double(scale)))
        m.periodic = [tuple(p) for p in pairs.tolist()]
        return m

    @classmethod
    def from_arrays(cls, nodes, tets, tris, tri_entity, periodic=(), scale=1.0):
        nodes = np.ascontiguousarray(nodes, np.float64)
        tets = np.ascontiguousarray(tets, np.int32)
        tris = np.ascontiguousarray(tris, np.int32)
        tri_entity = np.ascontiguousarray(tri_entity, np.int32)
        pairs = np.ascontiguousarray(np.array(list(periodic), np.int32).reshape(-1, 2))
        m = cls(lib().orc_mesh_from_arrays(_d(nodes), len(nodes), _i(tets), len(tets), _i(tris),
                                           _i(tri_entity), len(tris), _i(pairs), len(pairs),
                                           C.c_double(scale)))
        m.periodic = [tuple(p) for p in pairs.tolist()]
        return m

    def average_cell_size(self):
        return lib().orc_mesh_average_cell_size(self.h)

    def entity_faces(self, entity):
        n = lib().orc_mesh_entity_faces(self.h, entity, None, 0)
        if n < 0:
            raise KeyError(entity)
        out = np.zeros(n, np.int32)
        lib().orc_mesh_entity_faces(self.h, entity, _i(out), n)
        return out

    def labels(self):
        buf = C.create_string_buffer(1 << 16)
        lib().orc_mesh_labels(self.h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            k, v = line.split(":", 1)
            out[int(k)] = v.split("|") if v else []
        return out

    def __del__(self):
        try:
            lib().orc_mesh_free(self.h)
        except Exception:
            pass


class Poisson:
    """Restated stand-alone ``PoissonSolver`` (poisson.cpp), as test/poisson_test.cpp uses it."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = C.c_void_p(lib().orc_poisson_create(mesh.h))

    def set_bc(self, entity, kind, value=0.0, normal_grad=0.0):
        _check(lib().orc_poisson_set_bc(self.h, entity, QBC[kind], C.c_double(value), C.c_double(normal_grad)))

    def initialize(self):
        _check(lib().orc_poisson_initialize(self.h))

    def csr(self):
        n = self.mesh.nTets
        nnz = lib().orc_poisson_nnz(self.h)
        rp = np.zeros(n + 1, np.int32)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz)
        lib().orc_poisson_csr(self.h, _i(rp), _i(ci), _d(v))
        return rp, ci, v

    def solve(self, rho):
        rho = np.ascontiguousarray(rho, np.float64)
        n = self.mesh.nTets
        phi = np.zeros(n)
        E = np.zeros((n, 3))
        _check(lib().orc_poisson_solve(self.h, _d(rho), _d(phi), _d(E)))
        return phi, E

    def solve_system(self, rhs, guess=None):
        rhs = np.ascontiguousarray(rhs, np.float64)
        guess = None if guess is None else np.ascontiguousarray(guess, np.float64)
        x = np.zeros(self.mesh.nTets)
        _check(lib().orc_poisson_solve_system(self.h, _d(rhs), _d(guess), _d(x)))
        return x

    @property
    def last_iterations(self):
        return lib().orc_poisson_last_iterations(self.h)

    @property
    def last_error(self):
        return lib().orc_poisson_last_error(self.h)

    def __del__(self):
        try:
            lib().orc_poisson_free(self.h)
        except Exception:
            pass


class Sim:
    """Restated ``Solver<Full>`` (one species) / ``MulticomponentSolver<Full>`` (several)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.h = C.c_void_p(lib().orc_sim_create(mesh.h))
        self.grids = []

    def add_species(self, n, vmin, vmax, mass, charge, multiplier=1):
        n = np.asarray(n, np.int32)
        vmin = np.asarray(vmin, np.float64)
        vmax = np.asarray(vmax, np.float64)
        sp = lib().orc_sim_add_species(self.h, _i(n), _d(vmin), _d(vmax), C.c_double(mass), C.c_double(charge),
                                       int(multiplier))
        self.grids.append((tuple(int(x) for x in n), vmin.copy(), vmax.copy()))
        return sp

    def N(self, sp):
        n = self.grids[sp][0]
        return n[0] * n[1] * n[2]

    def set_maxwell(self, sp, density, temperature, mpv=(0, 0, 0)):
        density = np.ascontiguousarray(density, np.float64)
        mpv = np.asarray(mpv, np.float64)
        _check(lib().orc_sim_set_maxwell(self.h, sp, _d(density), C.c_double(temperature), _d(mpv)))

    def set_pdf(self, sp, f):
        f = np.ascontiguousarray(f, np.float64)
        assert f.size == self.mesh.nTets * self.N(sp)
        lib().orc_sim_set_pdf(self.h, sp, _d(f))

    def get_pdf(self, sp):
        f = np.zeros((self.mesh.nTets, self.N(sp)))
        lib().orc_sim_get_pdf(self.h, sp, _d(f))
        return f

    def set_params(self, sp, dt, ext=None, background=None, fused=False):
        ext = None if ext is None else np.asarray(ext, np.float64)
        background = None if background is None else np.ascontiguousarray(background, np.float64)
        lib().orc_sim_set_params(self.h, sp, C.c_double(dt), _d(ext), _d(background), int(fused))

    def set_particle_bc(self, sp, entity, kind, collect=False, source_pdf=None):
        src = None if source_pdf is None else np.ascontiguousarray(source_pdf, np.float64)
        _check(lib().orc_sim_set_particle_bc(self.h, sp, entity, PBC[kind], int(collect), _d(src)))

    def set_field_bc_potential(self, sp, entity, potential):
        _check(lib().orc_sim_set_field_bc(self.h, sp, entity, 1, C.c_double(potential)))

    def set_field_bc_charge(self, sp, entity, sigma):
        _check(lib().orc_sim_set_field_bc(self.h, sp, entity, 0, C.c_double(sigma)))

    def begin(self):
        _check(lib().orc_sim_begin(self.h))

    def init_wall(self):
        lib().orc_sim_init_wall(self.h)

    def step(self, iteration=0):
        _check(lib().orc_sim_step(self.h, iteration))

    def update_pdf(self, sp, E):
        E = np.ascontiguousarray(E, np.float64)
        _check(lib().orc_sim_update_pdf(self.h, sp, _d(E)))

    def fields(self, sp=0):
        n = self.mesh.nTets
        rho, phi, E = np.zeros(n), np.zeros(n), np.zeros((n, 3))
        lib().orc_sim_get_fields(self.h, sp, _d(rho), _d(phi), _d(E))
        return rho, phi, E

    def density(self, sp):
        out = np.zeros(self.mesh.nTets)
        lib().orc_sim_density(self.h, sp, _d(out))
        return out

    def velocity(self, sp):
        out = np.zeros((self.mesh.nTets, 3))
        lib().orc_sim_velocity(self.h, sp, _d(out))
        return out

    def wall_charge(self, sp, entity):
        return lib().orc_sim_wall_charge(self.h, sp, entity)

    def wall_area(self, sp, entity):
        return lib().orc_sim_wall_area(self.h, sp, entity)

    def poisson_iterations(self):
        return lib().orc_sim_poisson_iterations(self.h)

    def __del__(self):
        try:
            lib().orc_sim_free(self.h)
        except Exception:
            pass


class TuckerObj:
    """Restated stand-alone ``Tucker`` tensor (tucker.cpp), as test/tucker_test.cpp uses it."""

    def __init__(self, handle, n):
        self.h = C.c_void_p(handle)
        self.n = tuple(int(x) for x in n)

    @classmethod
    def from_full(cls, x, precision=0.0, rmax=1000000):
        """x: array of shape (n0, n1, n2); stored i0-fastest like Eigen::Tensor."""
        x = np.asarray(x, np.float64)
        n = np.asarray(x.shape, np.int32)
        flat = np.ascontiguousarray(x.ravel(order="F"))
        return cls(lib().orc_tucker_from_full(_d(flat), _i(n), C.c_double(precision), int(rmax)), x.shape)

    def clone(self):
        return TuckerObj(lib().orc_tucker_clone(self.h), self.n)

    def ranks(self):
        r = np.zeros(3, np.int32)
        lib().orc_tucker_ranks(self.h, _i(r))
        return tuple(int(x) for x in r)

    def reconstructed(self):
        out = np.zeros(self.n[0] * self.n[1] * self.n[2])
        lib().orc_tucker_reconstruct(self.h, _d(out))
        return out.reshape(self.n, order="F")

    def compress(self, precision=0.0, rmax=1000000):
        lib().orc_tucker_compress(self.h, C.c_double(precision), int(rmax))
        return self

    def sum(self):
        return lib().orc_tucker_sum(self.h)

    def axpy(self, s, other):
        """self <- self + s*other (operator+= / operator-= with the scalar product)."""
        lib().orc_tucker_axpy(self.h, C.c_double(s), other.h)
        return self

    def hadamard(self, other):
        lib().orc_tucker_hadamard(self.h, other.h)
        return self

    def __del__(self):
        try:
            lib().orc_tucker_free(self.h)
        except Exception:
            pass


class TuckerSim:
    """Restated ``Solver<Tucker>::_UpdatePDF`` + ``ParticleData<Tucker>`` for one species."""

    def __init__(self, mesh, n, vmin, vmax, mass, charge, compr_err, max_rank=0):
        self.mesh = mesh
        self.n = tuple(int(x) for x in n)
        n_ = np.asarray(n, np.int32)
        vmin_ = np.asarray(vmin, np.float64)
        vmax_ = np.asarray(vmax, np.float64)
        L = lib()
        L.orc_tsim_create.restype = C.c_void_p
        self.h = C.c_void_p(L.orc_tsim_create(mesh.h, _i(n_), _d(vmin_), _d(vmax_), C.c_double(mass),
                                              C.c_double(charge), C.c_double(compr_err), int(max_rank)))

    @property
    def N(self):
        return self.n[0] * self.n[1] * self.n[2]

    def set_particle_bc(self, entity, kind, collect=False, source=None):
        src = None if source is None else np.ascontiguousarray(source, np.float64)
        lib().orc_tsim_set_particle_bc_ex(self.h, int(entity), PBC[kind], int(bool(collect)), _d(src))

    def wall_charge(self, entity):
        L = lib()
        L.orc_tsim_wall_charge.restype = C.c_double
        return L.orc_tsim_wall_charge(self.h, int(entity))

    def set_pdf(self, f):
        f = np.ascontiguousarray(f, np.float64)
        assert f.size == self.mesh.nTets * self.N
        lib().orc_tsim_set_pdf(self.h, _d(f))

    def get_pdf(self):
        f = np.zeros((self.mesh.nTets, self.N))
        lib().orc_tsim_get_pdf(self.h, _d(f))
        return f

    def ranks(self):
        r = np.zeros((self.mesh.nTets, 3), np.int32)
        lib().orc_tsim_ranks(self.h, _i(r))
        return r

    def vnabs(self, face):
        out = np.zeros(self.N)
        lib().orc_tsim_vnabs(self.h, int(face), _d(out))
        return out

    def update_pdf(self, dt, E, ext=None):
        E = np.ascontiguousarray(E, np.float64)
        ext = None if ext is None else np.asarray(ext, np.float64)
        lib().orc_tsim_update_pdf(self.h, C.c_double(dt), _d(E), _d(ext))

    def density(self):
        out = np.zeros(self.mesh.nTets)
        lib().orc_tsim_density(self.h, _d(out))
        return out

    def __del__(self):
        try:
            lib().orc_tsim_free(self.h)
        except Exception:
            pass

"""Thin Python mirror of the C ABI (include/vt_b200.h) used by tests and bench.py.

The names follow the reference's domain: a context owns the flattened ``Mesh`` tables, each
species owns a ``VelocityGrid`` + ``ParticleData<Full>`` state on the device, ``step_full`` is
``Solver<Full>::_UpdatePDF`` (src/solver.cpp:141-212).
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi

PBC = dict(NonBoundary=0, Periodic=1, Source=2, Absorbing=3, Free=4)   # src/solver.h:25
QBC = dict(NonBoundary=0, Neumann=1, Dirichlet=2, Periodic=3)           # src/poisson.h:47


@dataclass
class MeshTables:
    """Flattened ``Mesh`` after ``Reconstruct`` (src/mesh.cpp:94-111), reference tet order."""
    nbr: np.ndarray            # (nT,4) int32, -1 where the reference holds nullptr
    area: np.ndarray           # (nT,4)
    volume: np.ndarray         # (nT,)
    normal: np.ndarray         # (nT,4,3)
    entity: np.ndarray         # (nT,4) int32, -1 internal
    tetCentroid: np.ndarray    # (nT,3)
    faceCentroid: np.ndarray   # (nT,4,3)
    order: np.ndarray = None   # locality permutation for the device layout (optional)
    brickTets: int = 0         # tets per L2 brick matching `order` (0 = library default)
    nGhost: int = 0
    periodic: list = field(default_factory=list)
    # partition only (partition.partition fills them): the reference's tet index of every local row and
    # the geometry of the ghost tets, which the partitioned Poisson assembly needs
    globalTets: int = 0
    globalId: np.ndarray = None          # (nT + nGhost,)
    ghostNbr: np.ndarray = None          # (nGhost, 4) local index or -1
    ghostArea: np.ndarray = None         # (nGhost, 4)
    ghostNormal: np.ndarray = None       # (nGhost, 4, 3)
    ghostTetCentroid: np.ndarray = None  # (nGhost, 3)
    ghostFaceCentroid: np.ndarray = None  # (nGhost, 4, 3)

    @property
    def nTets(self):
        return len(self.volume)


class Context:
    def __init__(self, device=0):
        """device: an index, or a list of indices for a device group (vt_ctx_create_group; entries may
        repeat = virtual ranks on one GPU)."""
        self.lib = capi.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            capi.check(self.lib.vt_ctx_create_group(arr, len(device), C.byref(h)))
        else:
            capi.check(self.lib.vt_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.nOwned = 0
        self.grids = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.vt_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm = C.c_int()
        l2 = C.c_size_t()
        hbm = C.c_size_t()
        capi.check(self.lib.vt_device_info(self.h, C.byref(sm), C.byref(l2), C.byref(hbm)))
        return dict(sm_count=sm.value, l2_bytes=l2.value, hbm_bytes=hbm.value)

    def sync(self):
        capi.check(self.lib.vt_sync(self.h))

    # ---- mesh
    def mesh_upload(self, mt: MeshTables):
        nbr = capi.i32(mt.nbr)
        area = capi.f64(mt.area)
        vol = capi.f64(mt.volume)
        nrm = capi.f64(mt.normal)
        ent = capi.i32(mt.entity)
        order = capi.i32(mt.order)
        self.nOwned = len(vol)
        self.mesh = mt
        capi.check(self.lib.vt_mesh_upload(self.h, self.nOwned, int(mt.nGhost), capi.ip(nbr), capi.dp(area),
                                           capi.dp(vol), capi.dp(nrm), capi.ip(ent), capi.ip(order)))
        if mt.brickTets:
            self.step_config(brick_tets=mt.brickTets)
        if mt.globalId is not None:
            gid = capi.i32(mt.globalId)
            gn = capi.i32(mt.ghostNbr if mt.nGhost else np.zeros((1, 4), np.int32))
            ga = capi.f64(mt.ghostArea if mt.nGhost else np.zeros((1, 4)))
            gnr = capi.f64(mt.ghostNormal if mt.nGhost else np.zeros((1, 4, 3)))
            gc = capi.f64(mt.ghostTetCentroid if mt.nGhost else np.zeros((1, 3)))
            gfc = capi.f64(mt.ghostFaceCentroid if mt.nGhost else np.zeros((1, 4, 3)))
            capi.check(self.lib.vt_mesh_set_ghost_geometry(self.h, int(mt.globalTets), capi.ip(gid), capi.ip(gn), capi.dp(ga),
                                                           capi.dp(gnr), capi.dp(gc), capi.dp(gfc)))

    # ---- species
    def species_create(self, n, vmin, vmax, mass, charge):
        n = capi.i32(n)
        vmin = capi.f64(vmin)
        vmax = capi.f64(vmax)
        sp = C.c_int()
        capi.check(self.lib.vt_species_create(self.h, capi.ip(n), capi.dp(vmin), capi.dp(vmax), float(mass),
                                              float(charge), C.byref(sp)))
        self.grids.append(tuple(int(x) for x in n))
        return sp.value

    def N(self, sp):
        n = self.grids[sp]
        return n[0] * n[1] * n[2]

    def set_face_bc(self, sp, bc_type, collect=None, source_id=None):
        bc = capi.u8(bc_type)
        col = capi.u8(collect)
        src = capi.i32(source_id)
        capi.check(self.lib.vt_species_set_face_bc(self.h, sp, capi.u8p(bc), capi.u8p(col), capi.ip(src)))

    def set_source_pdfs(self, sp, pdfs):
        pdfs = capi.f64(pdfs).reshape(-1, self.N(sp))
        capi.check(self.lib.vt_species_set_source_pdfs(self.h, sp, len(pdfs), capi.dp(pdfs)))

    def set_pdf(self, sp, f, first=0):
        f = capi.f64(f).reshape(-1, self.N(sp))
        capi.check(self.lib.vt_species_set_pdf(self.h, sp, int(first), len(f), capi.dp(f)))

    def get_pdf(self, sp, first=0, count=None):
        count = self.nOwned - first if count is None else count
        out = np.empty((count, self.N(sp)))
        capi.check(self.lib.vt_species_get_pdf(self.h, sp, int(first), int(count), capi.dp(out)))
        return out

    def set_maxwell(self, sp, density, temperature, mpv=(0.0, 0.0, 0.0)):
        d = capi.f64(density)
        v = capi.f64(mpv)
        capi.check(self.lib.vt_species_set_maxwell(self.h, sp, capi.dp(d), float(temperature), capi.dp(v)))

    def density(self, sp, download=True):
        out = np.empty(self.nOwned) if download else None
        capi.check(self.lib.vt_species_density(self.h, sp, capi.dp(out)))
        return out

    def velocity(self, sp):
        out = np.empty((self.nOwned, 3))
        capi.check(self.lib.vt_species_velocity(self.h, sp, capi.dp(out)))
        return out

    # ---- field
    def field_set(self, E):
        E = capi.f64(E)
        capi.check(self.lib.vt_field_set(self.h, capi.dp(E)))

    def field_get(self):
        rho, phi, E = np.empty(self.nOwned), np.empty(self.nOwned), np.empty((self.nOwned, 3))
        capi.check(self.lib.vt_field_get(self.h, capi.dp(rho), capi.dp(phi), capi.dp(E)))
        return rho, phi, E

    # ---- hot path
    def step_full(self, sp, dt, ext=(0.0, 0.0, 0.0)):
        ext = capi.f64(ext)
        capi.check(self.lib.vt_step_full(self.h, sp, float(dt), capi.dp(ext)))

    def step_full_host(self, sp, dt, E, density_out, ext=(0.0, 0.0, 0.0)):
        ext = capi.f64(ext)
        assert E.dtype == np.float64 and E.flags.c_contiguous and density_out.flags.c_contiguous
        capi.check(self.lib.vt_step_full_host(self.h, sp, float(dt), capi.dp(ext), capi.dp(E), capi.dp(density_out)))

    def step_config(self, chunk_planes=None, brick_tets=None, variant=None):
        self._cfg = getattr(self, "_cfg", [0, 0, 64])
        if chunk_planes is not None:
            self._cfg[0] = int(chunk_planes)
        if brick_tets is not None:
            self._cfg[1] = int(brick_tets)
        if variant is not None:
            self._cfg[2] = int(variant)
        capi.check(self.lib.vt_step_config(self.h, *self._cfg))

    def step_last_ms(self):
        ms = C.c_float()
        capi.check(self.lib.vt_step_last_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.vt_launch_count(self.h))

    def profile_begin(self):
        capi.check(self.lib.vt_profile_begin(self.h))

    def profile_end(self):
        """(region_ms, step_kernel_ms, step_kernels) measured with CUDA events on the context stream."""
        r, k, n = C.c_float(), C.c_float(), C.c_int()
        capi.check(self.lib.vt_profile_end(self.h, C.byref(r), C.byref(k), C.byref(n)))
        return r.value, k.value, n.value

    def wall_charge(self, sp, entity):
        q = C.c_double()
        capi.check(self.lib.vt_wall_charge_get(self.h, sp, int(entity), C.byref(q)))
        return q.value

    def wall_charge_reset(self, sp):
        capi.check(self.lib.vt_wall_charge_reset(self.h, sp))

    # ---- multi-GPU halo
    HALO_HANDLE_BYTES = 192

    def set_separable(self, sp, amp, a0, a1, a2):
        """pdf[t] = sum_k amp[t,k] a0[k] (x) a1[k] (x) a2[k], filled on the device."""
        amp = capi.f64(amp).reshape(self.nOwned, -1)
        k = amp.shape[1]
        a0, a1, a2 = (capi.f64(a).reshape(k, -1) for a in (a0, a1, a2))
        capi.check(self.lib.vt_species_set_separable(self.h, sp, k, capi.dp(amp), capi.dp(a0), capi.dp(a1), capi.dp(a2)))

    def dfma_peak_tflops(self):
        v = C.c_double()
        capi.check(self.lib.vt_measure_dfma_peak(self.h, C.byref(v)))
        return v.value

    def halo_export(self, sp):
        buf = np.zeros(self.HALO_HANDLE_BYTES, np.uint8)
        capi.check(self.lib.vt_halo_export(self.h, sp, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def halo_attach(self, sp, my_rank, peer_ranks, peer_handles):
        pr = capi.i32(peer_ranks)
        ph = np.ascontiguousarray(peer_handles, np.uint8).reshape(len(pr), self.HALO_HANDLE_BYTES)
        capi.check(self.lib.vt_halo_attach(self.h, sp, int(my_rank), len(pr), capi.ip(pr), ph.ctypes.data_as(C.c_void_p)))

    def halo_attach_local(self, sp, my_rank, peer_ranks, peer_ctxs, peer_species):
        """In-process peers (virtual ranks on one device, or peer-accessible devices): no CUDA IPC."""
        pr = capi.i32(peer_ranks)
        arr = (C.c_void_p * len(pr))(*[c.h for c in peer_ctxs])
        ps = capi.i32(peer_species)
        capi.check(self.lib.vt_halo_attach_local(self.h, sp, int(my_rank), len(pr), capi.ip(pr), arr, capi.ip(ps)))

    def halo_set_push(self, sp, push_peer, push_row):
        pp = capi.i32(push_peer).reshape(self.nOwned, 4)
        pr = capi.i32(push_row).reshape(self.nOwned, 4)
        capi.check(self.lib.vt_halo_set_push(self.h, sp, capi.ip(pp), capi.ip(pr)))

    def halo_push_current(self, sp):
        capi.check(self.lib.vt_halo_push_current(self.h, sp))

    def halo_barrier(self):
        capi.check(self.lib.vt_halo_barrier(self.h))

    # ---- Poisson
    def poisson_setup(self, bc_type, bc_value=None, bc_normal_grad=None):
        mt = self.mesh
        nT = self.nOwned
        bc = capi.u8(bc_type).reshape(nT, 4)
        val = capi.f64(np.zeros((nT, 4)) if bc_value is None else bc_value)
        ng = capi.f64(np.zeros((nT, 4)) if bc_normal_grad is None else bc_normal_grad)
        tc = capi.f64(mt.tetCentroid)
        fc = capi.f64(mt.faceCentroid)
        capi.check(self.lib.vt_poisson_setup(self.h, capi.dp(tc), capi.dp(fc), capi.u8p(bc), capi.dp(val), capi.dp(ng)))

    # partitioned solve (see include/vt_b200.h)
    POISSON_HANDLE_BYTES = 128

    def poisson_set_global_dirichlet(self, any_dirichlet):
        capi.check(self.lib.vt_poisson_set_global_dirichlet(self.h, int(bool(any_dirichlet))))

    def poisson_comm_export(self):
        buf = np.zeros(self.POISSON_HANDLE_BYTES, np.uint8)
        capi.check(self.lib.vt_poisson_comm_export(self.h, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def poisson_comm_attach(self, my_rank, handles):
        h = np.ascontiguousarray(handles, np.uint8).reshape(-1, self.POISSON_HANDLE_BYTES)
        capi.check(self.lib.vt_poisson_comm_attach(self.h, int(my_rank), len(h), h.ctypes.data_as(C.c_void_p)))

    def poisson_comm_attach_local(self, my_rank, ctxs):
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        capi.check(self.lib.vt_poisson_comm_attach_local(self.h, int(my_rank), len(ctxs), arr))

    def poisson_set_push(self, push_rank, push_row):
        pr = capi.i32(push_rank).reshape(self.nOwned, 4)
        prow = capi.i32(push_row).reshape(self.nOwned, 4)
        capi.check(self.lib.vt_poisson_set_push(self.h, capi.ip(pr), capi.ip(prow)))

    def poisson_update_bc_values(self, bc_value, bc_normal_grad):
        val = capi.f64(bc_value)
        ng = capi.f64(bc_normal_grad)
        capi.check(self.lib.vt_poisson_update_bc_values(self.h, capi.dp(val), capi.dp(ng)))

    def poisson_solve(self, rho=None, download=True):
        rho = capi.f64(rho)
        phi = np.empty(self.nOwned) if download else None
        E = np.empty((self.nOwned, 3)) if download else None
        capi.check(self.lib.vt_poisson_solve(self.h, capi.dp(rho), capi.dp(phi), capi.dp(E)))
        return phi, E

    def poisson_stats(self):
        it = C.c_int()
        res = C.c_double()
        capi.check(self.lib.vt_poisson_stats(self.h, C.byref(it), C.byref(res)))
        return it.value, res.value

    def charge_density(self, species, background=None):
        sp = capi.i32(species)
        bg = capi.f64(background)
        capi.check(self.lib.vt_charge_density(self.h, capi.ip(sp), len(sp), capi.dp(bg)))

    # ---- Tucker format (ParticleData<Tucker> / Solver<Tucker>) ----
    def tucker_enable(self, sp, compr_err=1e-10, max_rank=0):
        capi.check(self.lib.vt_tucker_enable(self.h, sp, float(compr_err), int(max_rank)))
        self._tucker_cap = getattr(self, "_tucker_cap", {})
        self._tucker_cap[sp] = int(max_rank)

    def tucker_set_pdf(self, sp, dense):
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        assert dense.shape[0] == self.nOwned
        capi.check(self.lib.vt_tucker_set_pdf(self.h, sp, capi.dp(dense)))

    def tucker_get_pdf(self, sp, N):
        out = np.empty((self.nOwned, N))
        capi.check(self.lib.vt_tucker_get_pdf(self.h, sp, capi.dp(out)))
        return out

    def tucker_ranks(self, sp):
        r = np.empty((self.nOwned, 3), np.int32)
        capi.check(self.lib.vt_tucker_get_ranks(self.h, sp, capi.ip(r)))
        return r

    def tucker_factors(self, sp, tet, n):
        """(core[r0,r1,r2] i0-fastest, [U0,U1,U2] column-major n_k x r_k) of one tet."""
        r = np.zeros(3, np.int32)
        core = np.empty(n[0] * n[1] * n[2])
        us = [np.empty(n[k] * n[k]) for k in range(3)]
        capi.check(self.lib.vt_tucker_get_factors(self.h, sp, int(tet), capi.ip(r), capi.dp(core), capi.dp(us[0]),
                                                  capi.dp(us[1]), capi.dp(us[2])))
        c = core[:r[0] * r[1] * r[2]].reshape(r[2], r[1], r[0]).transpose(2, 1, 0)
        U = [us[k][:n[k] * r[k]].reshape(r[k], n[k]).T for k in range(3)]
        return c, U

    def tucker_density(self, sp):
        d = np.empty(self.nOwned)
        capi.check(self.lib.vt_tucker_density(self.h, sp, capi.dp(d)))
        return d

    def tucker_last_kernel(self, sp):
        """'general' (csrc/tucker.cu) or 'slab' (csrc/tucker_slab.cu): the kernel the last step_tucker ran."""
        k = C.c_int()
        capi.check(self.lib.vt_tucker_last_kernel(self.h, sp, C.byref(k)))
        return {0: None, 1: "general", 2: "slab"}[k.value]

    TUCKER_HALO_HANDLE_BYTES = 128

    def tucker_halo_export(self, sp):
        buf = np.zeros(self.TUCKER_HALO_HANDLE_BYTES, np.uint8)
        capi.check(self.lib.vt_tucker_halo_export(self.h, sp, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def tucker_halo_attach(self, sp, peer_handles):
        h = np.ascontiguousarray(peer_handles, np.uint8).reshape(-1, self.TUCKER_HALO_HANDLE_BYTES)
        capi.check(self.lib.vt_tucker_halo_attach(self.h, sp, len(h), h.ctypes.data_as(C.c_void_p)))

    def tucker_halo_attach_local(self, sp, peer_ctxs, peer_species):
        arr = (C.c_void_p * len(peer_ctxs))(*[c.h for c in peer_ctxs])
        ps = capi.i32(peer_species)
        capi.check(self.lib.vt_tucker_halo_attach_local(self.h, sp, len(peer_ctxs), arr, capi.ip(ps)))

    def step_tucker(self, sp, dt, ext=(0.0, 0.0, 0.0)):
        ext = capi.f64(ext)
        capi.check(self.lib.vt_step_tucker(self.h, sp, float(dt), capi.dp(ext)))

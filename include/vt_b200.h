/* vt_b200.h — C ABI of libvt_b200.so, the B200 (sm_100a) implementation of VlasovTucker's
 * per-time-step kinetic update.
 *
 * The reference (DmitriiGurev/VlasovTucker, C++) has no FFI of its own: its boundary for this
 * path is the header-level C++ API of src/header.h.  The entry points below are what the host
 * classes of that API (vlasovtucker_b200/host/, same names as the reference's) bind, one per
 * reference function on the hot path; each comment cites the reference code it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; vt_last_error() gives the text
 *    (the host classes rethrow it as the std::runtime_error / std::invalid_argument the
 *    reference would have thrown);
 *  - all pointers are HOST pointers owned by the caller unless the name ends in _dev;
 *  - tet-indexed arrays are in the caller's (= reference's) tet order; the library keeps its
 *    own locality order internally (vt_mesh_upload's `order`) and translates on every call;
 *  - FP64 throughout; indices are 32-bit; tensors are column-major, i0 fastest, exactly as
 *    Eigen::Tensor<double,3> in the reference (src/typedefs.h:9);
 *  - one host thread per context; calls are stream-ordered inside the context;
 *  - there is no CPU fallback: without a CUDA device vt_ctx_create fails.
 */
#ifndef VT_B200_H
#define VT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vt_ctx vt_ctx;

/* ParticleBCType, src/solver.h:25 */
enum { VT_PBC_NONBOUNDARY = 0, VT_PBC_PERIODIC = 1, VT_PBC_SOURCE = 2, VT_PBC_ABSORBING = 3, VT_PBC_FREE = 4 };
/* PoissonBCType, src/poisson.h:47 */
enum { VT_QBC_NONBOUNDARY = 0, VT_QBC_NEUMANN = 1, VT_QBC_DIRICHLET = 2, VT_QBC_PERIODIC = 3 };

const char* vt_last_error(void);
int vt_version(void);

/* ---- context ------------------------------------------------------------------------------ */
int vt_ctx_create(int device, vt_ctx** out);
/* One context that drives several GPUs from one host thread (devices may repeat: "virtual ranks" on
 * one GPU).  The group partitions the mesh of vt_mesh_upload over its members (contiguous chunks of
 * the locality order), wires their ghost rows directly to each other and translates every per-tet
 * call below into calls on the members; per step it runs exactly what the one-process-per-GPU path
 * runs (fused halo push, device-side barriers, partitioned Poisson solve).  This is how the C++ host
 * classes use all GPUs of a box (VT_DEVICES=0,1,...; the reference's Solver<T>::Solve,
 * src/solver.cpp:80-138, is single-process).  The vt_halo_*, vt_poisson_comm_*, vt_profile_* and
 * vt_step_full_host entry points apply to single-device contexts only. */
int vt_ctx_create_group(const int* devices, int nDevices, vt_ctx** out);
int vt_group_size(vt_ctx* ctx);
void vt_ctx_destroy(vt_ctx* ctx);
int vt_sync(vt_ctx* ctx);
/* device properties the bench reports (SM count, L2 bytes, HBM bytes) */
int vt_device_info(vt_ctx* ctx, int* sm_count, size_t* l2_bytes, size_t* hbm_bytes);

/* ---- mesh tables: the flattened Mesh the face loop of Solver::_UpdatePDF walks --------------
 * replaces the Tet / Face pointer-graph reads of src/solver.cpp:163-168, 319 and
 * src/primitives.h:62-97.
 *   nOwned        tets this context updates
 *   nGhost        extra read-only rows appended after the owned ones (multi-GPU halo); a
 *                 neighbour index in [nOwned, nOwned+nGhost) names a ghost row
 *   nbr[4*t+j]    tet across face j (tet->adjTets[j]->index), -1 where the reference holds
 *                 nullptr
 *   area, volume  face->area (4 per tet), tet->volume
 *   normal[12*t+3*j+k]   face->normal
 *   entity[4*t+j] face->entity (-1 internal)
 *   order         optional permutation: order[p] = caller index of the tet stored at device
 *                 row p (NULL = identity).  Pure layout hint; results do not depend on it.
 */
int vt_mesh_upload(vt_ctx* ctx, int nOwned, int nGhost, const int32_t* nbr, const double* area,
                   const double* volume, const double* normal, const int32_t* entity,
                   const int32_t* order);

/* ---- species: VelocityGrid + ParticleData<Full> + the per-face particle BCs of Solver -------
 * replaces VelocityGrid (src/velocity_grid.cpp:9-53) and ParticleData (src/particle_data.h:21-52).
 */
int vt_species_create(vt_ctx* ctx, const int32_t n[3], const double vmin[3], const double vmax[3],
                      double mass, double charge, int* species);
/* ParticleData::mass / ::charge are public members the drivers assign after construction
 * (examples/oscillations.cpp:32-33); push the current values before they are used */
int vt_species_set_params(vt_ctx* ctx, int species, double mass, double charge);
/* Solver::SetParticleBC (src/solver.cpp:62-71): per-face BC type (uint8, 4 per tet), the
 * collectCharge flag, and for Source faces an index into the source-PDF table (-1 otherwise). */
int vt_species_set_face_bc(vt_ctx* ctx, int species, const uint8_t* bcType, const uint8_t* collect,
                           const int32_t* sourceId);
/* ParticleBC::sourcePDF tensors (src/solver.h:31), nSource x N doubles */
int vt_species_set_source_pdfs(vt_ctx* ctx, int species, int nSource, const double* pdf);
/* pdf[t] = Full(tensor) for t in [first, first+count): caller order, N doubles per tet */
int vt_species_set_pdf(vt_ctx* ctx, int species, int first, int count, const double* pdf);
/* pdf[t].Reconstructed() (src/full.cpp:24-27) */
int vt_species_get_pdf(vt_ctx* ctx, int species, int first, int count, double* pdf);
/* SetMaxwellPDF on the device (src/particle_data.cpp:23-90) */
int vt_species_set_maxwell(vt_ctx* ctx, int species, const double* physDensity, double temperature,
                           const double mostProbableV[3]);
/* pdf[t] = sum_k amp[t*nTerms+k] * a0_k (x) a1_k (x) a2_k, built on the device (a0: nTerms x n0, ...): the
 * C5 inputs of the Tucker sweep — sums of shifted / anisotropic Maxwellians, multilinear rank nTerms —
 * without a dense host array (the loop over tets of src/particle_data.cpp:33-66, generalised) */
int vt_species_set_separable(vt_ctx* ctx, int species, int nTerms, const double* amp, const double* a0,
                             const double* a1, const double* a2);
/* ParticleData::Density (src/particle_data.cpp:93-102) and Velocity (:105-125);
 * out may be NULL (result stays on the device for the Poisson step) */
int vt_species_density(vt_ctx* ctx, int species, double* density);
int vt_species_velocity(vt_ctx* ctx, int species, double* velocity /* 3 per tet */);

/* ---- field ---------------------------------------------------------------------------------- */
/* copy a caller-computed field (3 doubles per tet) to the device; replaces Solver::_field
 * (src/solver.cpp:110) when the Poisson solve ran elsewhere */
int vt_field_set(vt_ctx* ctx, const double* E);
int vt_field_get(vt_ctx* ctx, double* rho, double* phi, double* E);

/* ---- the hot path: Solver<Full>::_UpdatePDF (src/solver.cpp:141-212) ------------------------
 * rhs = -sum_f (A_f/V) flux_f - sum_k (q/m)(E_k+ext_k) d f/d v_k ;  f += dt*rhs, for every owned
 * tet, using the field currently on the device.  Also leaves Density() of the new state on the
 * device and accumulates the absorbed wall charge (src/solver.cpp:171-178). */
int vt_step_full(vt_ctx* ctx, int species, double dt, const double ext[3]);
/* the same through host buffers, as one call of the reference-facing Solver would do it: copies
 * E (3*nOwned doubles) host->device, steps, copies Density() (nOwned doubles) device->host */
int vt_step_full_host(vt_ctx* ctx, int species, double dt, const double ext[3], const double* E,
                      double* density);
/* tuning knobs of the step kernel: planes of the velocity grid per work item (0 = whole tensor),
 * tets per L2 brick (0 = no bricks) and the kernel variant as a bit set:
 *   64 the library's choice from the velocity grid (the default: bulk-copy pipeline where planes of
 *      >= 4 KiB can be staged, register-staged kernel otherwise, upwind-select arithmetic);
 *    1 no warp shuffles, 2 upwind-select arithmetic (vn>0 ? vn*f : vn*fa instead of the reference's
 *      0.5*(vn*(fa+f) - |vn|*(fa-f)); equal up to rounding), 4 three resident CTAs per SM,
 *    (8 unused,) 16 persistent bulk-copy (cp.async.bulk + mbarrier)
 *      producer/consumer pipeline, 32 with 16: eight consumer warps x two columns instead of sixteen;
 *  128 with 16|32|2: neighbour planes as whole-plane bulk copies instead of the consumer-side loads of
 *      only the inflow half of every neighbour row (the default where the eight-warp layout applies),
 *  256 with 16: work items in tet-major order (the chunks of one tet adjacent in the queue) instead of
 *      brick, chunk, tet.
 * Every variant computes the same update; they differ in scheduling only (tests/test_full_gpu.py). */
int vt_step_config(vt_ctx* ctx, int chunkPlanes, int brickTets, int variant);
/* device time of the last vt_step_full kernel in milliseconds (CUDA events) and launches so far */
int vt_step_last_ms(vt_ctx* ctx, float* ms);
long vt_launch_count(vt_ctx* ctx);
/* FP64 FMA peak of the device in TFLOP/s from a short register-resident probe (the denominator of
 * the Tucker roofline; MEASURED_PEAKS.json carries no FP64 figure) */
int vt_measure_dfma_peak(vt_ctx* ctx, double* tflops);
/* measurement on the context's own stream (CUDA events): vt_profile_begin marks the start of a
 * timed region; vt_profile_end synchronises and returns the region's device time, the summed
 * device time of the step kernels launched inside it and how many there were */
int vt_profile_begin(vt_ctx* ctx);
int vt_profile_end(vt_ctx* ctx, float* region_ms, float* step_kernel_ms, int* step_kernels);

/* ---- multi-GPU halo (one process per GPU; the reference is single-process, so this has no
 * reference counterpart: it is what keeps Solver::_UpdatePDF's neighbour reads, src/solver.cpp:
 * 319-327, valid when the mesh is partitioned).  Ghost rows live after the owned rows of each
 * species' state buffers.  The step kernel itself writes every boundary tet's new state into
 * the ghost rows of the peer GPUs (NVLink peer stores through CUDA-IPC mappings), so the
 * exchange is fused with the sweep; vt_halo_barrier is the only synchronisation between steps.
 *   vt_halo_export   3 x 64-byte CUDA IPC handles: state buffer 0, state buffer 1, barrier flags
 *   vt_halo_attach   open the peers' handles (peerHandles = nPeers x 192 bytes, same layout)
 *   vt_halo_set_push for each owned tet (caller order) up to 4 (peer index, ghost row on that
 *                    peer) pairs, -1 = unused
 *   vt_halo_push_current  copy the current state of the pushed tets to the peers (initial fill)
 *   vt_halo_barrier  device-side barrier with all attached peers on the context stream */
int vt_halo_export(vt_ctx* ctx, int species, void* handles);
int vt_halo_attach(vt_ctx* ctx, int species, int myRank, int nPeers, const int32_t* peerRanks,
                   const void* peerHandles);
/* the same wiring for contexts of ONE process — several devices with peer access driven by one host
 * thread (the C++ Solver under VT_DEVICES), or several "virtual ranks" on one device (tests on a
 * one-GPU box): the ghost rows point straight at the peers' buffers, no CUDA IPC involved.
 * peerSpecies[i] is the species id of the same species in peerCtx[i]. */
int vt_halo_attach_local(vt_ctx* ctx, int species, int myRank, int nPeers, const int32_t* peerRanks,
                         vt_ctx* const* peerCtx, const int32_t* peerSpecies);
int vt_halo_set_push(vt_ctx* ctx, int species, const int32_t* pushPeer, const int32_t* pushRow);
int vt_halo_push_current(vt_ctx* ctx, int species);
int vt_halo_barrier(vt_ctx* ctx);

/* wall charge, src/solver.cpp:171-178, 296-311: accumulated charge of entity */
int vt_wall_charge_get(vt_ctx* ctx, int species, int entity, double* charge);
int vt_wall_charge_reset(vt_ctx* ctx, int species);

/* ---- Poisson: PoissonSolver (src/poisson.cpp) ----------------------------------------------- */
/* geometry the field kernels need beyond vt_mesh_upload: centroids (3 per tet) and face
 * centroids (12 per tet), plus per-face BC (type uint8, value, normalGrad)
 * — PoissonSolver::SetBC / Initialize (src/poisson.cpp:85-124) */
int vt_poisson_setup(vt_ctx* ctx, const double* tetCentroid, const double* faceCentroid,
                     const uint8_t* bcType, const double* bcValue, const double* bcNormalGrad);
/* ---- partitioned Poisson solve (one rank per GPU, or several contexts of one process): the rows are
 * the rank's owned tets, ghost values of phi / z / grad(phi) arrive by peer stores, and the two dot
 * products of every CG iteration are summed over the ranks inside the persistent solve kernel
 * (rank-ordered, so every rank holds the same bits) — the role north_star gives an NCCL allreduce.
 * No reference counterpart (the reference is one process); the matrix, right-hand side, correction and
 * gradient are those of src/poisson.cpp, assembled per rank from its rows plus the ghost geometry.
 *   vt_mesh_set_ghost_geometry  after vt_mesh_upload: globalId[nOwned+nGhost] = the reference's tet index of
 *                               every local row (decides the Upper-triangle entry and the pinned row 0);
 *                               per ghost tet its 4 neighbours as local indices (-1 = not on this rank),
 *                               face areas, normals, centroid and face centroids
 *   vt_poisson_set_global_dirichlet  before vt_poisson_setup: whether ANY rank holds a Dirichlet BC
 *                               (poisson.cpp:87-88 decides the pinned row from it)
 *   vt_poisson_comm_export      128-byte handle of this rank's exchange block (after vt_poisson_setup)
 *   vt_poisson_comm_attach      all ranks' handles (world x 128 bytes, own entry ignored)
 *   vt_poisson_comm_attach_local  the same for contexts of one process
 *   vt_poisson_set_push         per owned tet (caller order) up to 4 (rank, ghost row on that rank) pairs */
int vt_mesh_set_ghost_geometry(vt_ctx* ctx, int globalTets, const int32_t* globalId, const int32_t* ghostNbr,
                               const double* ghostArea, const double* ghostNormal, const double* ghostTetCentroid,
                               const double* ghostFaceCentroid);
int vt_poisson_set_global_dirichlet(vt_ctx* ctx, int anyDirichlet);
int vt_poisson_comm_export(vt_ctx* ctx, void* handle);
int vt_poisson_comm_attach(vt_ctx* ctx, int myRank, int world, const void* handles);
int vt_poisson_comm_attach_local(vt_ctx* ctx, int myRank, int world, vt_ctx* const* ranks);
int vt_poisson_set_push(vt_ctx* ctx, const int32_t* pushRank, const int32_t* pushRow);
/* Neumann/Dirichlet data may change between solves (src/solver.cpp:120-132) */
int vt_poisson_update_bc_values(vt_ctx* ctx, const double* bcValue, const double* bcNormalGrad);
/* PoissonSolver::Solve (src/poisson.cpp:179-213); rho NULL = use the device-resident charge
 * density assembled by vt_charge_density.  phi/E may be NULL. */
int vt_poisson_solve(vt_ctx* ctx, const double* rho, double* phi, double* E);
/* iterations and relative residual of the last solve (synchronises the context's stream: the solve
 * itself is launched asynchronously) */
int vt_poisson_stats(vt_ctx* ctx, int* lastIterations, double* lastRelResidual);
/* rho = sum_s charge_s * Density_s + background (src/solver.cpp:98-105,
 * src/multicomponent_solver.cpp:61-74); background may be NULL */
int vt_charge_density(vt_ctx* ctx, const int* species, int nSpecies, const double* background);

/* ---- Tucker format: ParticleData<Tucker> + Solver<Tucker> (src/tucker.cpp, src/solver.cpp) ---- */
/* switch a species to Tucker storage: ParticleData::SetCompressionError / SetMaxRank
 * (src/particle_data.cpp:18, 174-185); maxRank <= 0 = max(n) as the reference defaults */
int vt_tucker_enable(vt_ctx* ctx, int species, double comprErr, int maxRank);
/* dense rows (nOwned x N, caller order) -> exact Tucker tensors, as ParticleData<Tucker>::
 * SetMaxwellPDF builds them with precision 0 (src/particle_data.cpp:64-69) */
int vt_tucker_set_pdf(vt_ctx* ctx, int species, const double* dense);
/* Tucker::Reconstructed() of every tet (src/tucker.cpp:100-104) */
int vt_tucker_get_pdf(vt_ctx* ctx, int species, double* dense);
/* Tucker::Ranks() of every tet, 3 per tet (src/tucker.cpp:126-129) */
int vt_tucker_get_ranks(vt_ctx* ctx, int species, int32_t* ranks);
/* Tucker::Core()/U() of one tet; buffers sized for the rank capacity min(maxRank, n_k) */
int vt_tucker_get_factors(vt_ctx* ctx, int species, int tet, int32_t ranks[3], double* core, double* u0,
                          double* u1, double* u2);
/* ParticleData<Tucker>::Density (src/particle_data.cpp:93-102); density may be NULL */
int vt_tucker_density(vt_ctx* ctx, int species, double* density);
/* which kernel the last vt_step_tucker of this species ran: 0 none yet, 1 general (csrc/tucker_kernel.inl, any grid up to
 * 64 nodes per axis, any compression error), 2 slab-streaming (csrc/tucker_slab.cu, 33..48 nodes per axis, rank cap
 * <= 16, comprErr >= 5.5e-7).  Diagnostics for bench.py and the tests; no reference counterpart. */
int vt_tucker_last_kernel(vt_ctx* ctx, int species, int* kernel);
/* Solver<Tucker>::_UpdatePDF for one species (src/solver.cpp:141-212): flux per face, rounding
 * after each face, acceleration term, rounding, Euler update, rounding */
int vt_step_tucker(vt_ctx* ctx, int species, double dt, const double ext[3]);
/* multi-GPU Tucker species (after vt_halo_export/attach/set_push of the same species and after the
 * last vt_tucker_enable): 128-byte handle of the compressed state to hand to the peers, and the
 * peers' handles in the order of vt_halo_attach.  vt_step_tucker then also stores every boundary
 * tet's new core, factors and ranks into the peers' ghost rows (fixed-capacity slots), and
 * vt_halo_push_current / vt_halo_barrier serve the species as they do for the full format. */
int vt_tucker_halo_export(vt_ctx* ctx, int species, void* handle);
int vt_tucker_halo_attach(vt_ctx* ctx, int species, int nPeers, const void* peerHandles);
/* in-process counterpart (after vt_halo_attach_local with the same peers) */
int vt_tucker_halo_attach_local(vt_ctx* ctx, int species, int nPeers, vt_ctx* const* peerCtx,
                                const int32_t* peerSpecies);

#ifdef __cplusplus
}
#endif
#endif /* VT_B200_H */

/* include/mdpy_b200.h — C ABI of libmdpyb200.so, the B200-native (sm_100a) nonbonded
 * hot path behind mdpy's Constraint / Integrator API.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / numpy /
 * C++ types), returns an int status (MDK_OK == 0, negative == error, text from
 * mdk_last_error) and never throws or aborts across the boundary.  Host pointers are
 * borrowed for the duration of the call only.  One mdk_ctx == one CUDA device == one host
 * thread at a time (not re-entrant).  File:line citations are relative to the reference
 * tree (mdpy v0.2.x); INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Units are mdpy's internal ones: angstrom, femtosecond, dalton, e; energy Da*A^2/fs^2
 * (mdpy/unit/__init__.py:31-41).
 */
#ifndef MDPY_B200_H
#define MDPY_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MDK_API __attribute__((visibility("default")))
#else
#define MDK_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdk_ctx mdk_ctx;

/* ---- status codes (mapped by the Python wrapper onto mdpy/error.py classes) ---- */
enum {
    MDK_OK = 0,
    MDK_ERR_BAD_ARG = -1,          /* ValueError / ArrayDimError (mdpy/error.py:37) */
    MDK_ERR_CUDA = -2,             /* RuntimeError */
    MDK_ERR_NOT_BOUND = -3,        /* NonBoundedError (mdpy/error.py:87, constraint.py:39-43) */
    MDK_ERR_CUTOFF_TOO_LARGE = -4, /* CellListPoorDefinedError (mdpy/error.py:123, cell_list.py:58-69) */
    MDK_ERR_PARTICLE_LOST = -5,    /* ParticleLossError (mdpy/error.py:203, utils/pbc.py:30-34) */
    MDK_ERR_OOM = -6,
    MDK_ERR_NCCL = -7,
    MDK_ERR_NLIST_STALE = -8       /* an atom moved more than skin/2 between rebuilds */
};

/* ---- force / energy terms (bit flags for mdk_compute, indices into energies[]) ---- */
enum {
    MDK_TERM_LJ = 1u << 0,          /* CharmmNonbondedConstraint.update, charmm_nonbonded_constraint.py:183-226 */
    MDK_TERM_COUL_DIRECT = 1u << 1, /* erfc(alpha r)/r inside rc over the tile list [not in the reference tree] */
    MDK_TERM_PME_RECIP = 1u << 2,   /* spread -> FFT -> influence function -> IFFT -> gather (+self, background, excluded-pair terms) */
    MDK_TERM_COUL_BARE = 1u << 3,   /* ElectrostaticConstraint.update, electrostatic_constraint.py:137-174 (all pairs, minimum image) */
    MDK_TERM_BOND = 1u << 4,        /* charmm_bond_constraint.py:53-73 */
    MDK_TERM_ANGLE = 1u << 5,       /* charmm_angle_constraint.py:55-96 (incl. Urey-Bradley) */
    MDK_TERM_DIHEDRAL = 1u << 6,    /* charmm_dihedral_constraint.py:59-95 */
    MDK_TERM_IMPROPER = 1u << 7     /* charmm_improper_constraint.py:57-94 */
};
enum {
    MDK_E_LJ = 0,
    MDK_E_COUL_DIRECT = 1,
    MDK_E_PME_RECIP = 2,
    MDK_E_PME_SELF = 3,  /* self + neutralising background */
    MDK_E_PME_EXCL = 4,  /* excluded-pair erf correction */
    MDK_E_COUL_BARE = 5,
    MDK_E_BOND = 6,
    MDK_E_ANGLE = 7,
    MDK_E_DIHEDRAL = 8,
    MDK_E_IMPROPER = 9,
    MDK_E_KINETIC = 10,
    MDK_NUM_ENERGIES = 16
};

/* ---- lifetime ---- */
/* Replaces the per-call cuda.to_device allocations of the reference
 * (charmm_nonbonded_constraint.py:197-209): the context owns every device buffer,
 * stream, cuFFT plan.  device = CUDA ordinal.  Fails (MDK_ERR_CUDA) when no sm_100 GPU
 * is present — there is no CPU fallback. */
MDK_API int mdk_create(int device, mdk_ctx **out);
MDK_API void mdk_destroy(mdk_ctx *ctx);
/* Last error text of this ctx (or of a failed mdk_create when ctx == NULL). */
MDK_API const char *mdk_last_error(const mdk_ctx *ctx);
/* Run all work of this ctx on an existing CUDA stream (cudaStream_t handle, e.g.
 * torch.cuda.current_stream().cuda_stream) instead of the private one. */
MDK_API int mdk_set_stream(mdk_ctx *ctx, void *cuda_stream);

/* ---- system definition (bind time; reference: Constraint.bind_ensemble) ---- */
/* Orthorhombic box edge lengths = diag(State.pbc_matrix) (state.py:47-54; SURVEY Q3: the
 * reference's CUDA kernels and cell list read only the diagonal). */
MDK_API int mdk_set_box(mdk_ctx *ctx, const double box[3]);
/* topology.charges / topology.masses, float32 [n] (topology.py:59-68). */
MDK_API int mdk_set_atoms(mdk_ctx *ctx, int n, const float *charges, const float *masses);
/* Per-atom [eps, sigma, eps14, sigma14] table float32 [n,4] exactly as
 * CharmmNonbondedConstraint.bind_ensemble builds it (charmm_nonbonded_constraint.py:48-62).
 * rc = cutoff (inclusive, :90).  r_switch >= rc: the reference's plain truncation;
 * r_switch < rc: CHARMM energy switch on (r_switch, rc] [not in the reference tree]. */
MDK_API int mdk_set_lj(mdk_ctx *ctx, const float *eps_sigma, float rc, float r_switch);
/* env.set_precision('DOUBLE') (mdpy/environment.py:23-42 switches the reference's arithmetic type).  With
 * double_precision != 0 and the float64 parameter copies of mdk_set_params_f64 (charges [n], eps/sigma table [n,4]; either
 * may be NULL), LJ, erfc direct space, the bonded terms, the excluded-pair correction and the all-pairs Coulomb sum are
 * evaluated in float64 on the float64 positions (mdk_upload_positions_f64); sums stay int64 fixed point (2^-40).
 * The PME mesh (spreading, FFT, gather) stays float32. */
MDK_API int mdk_set_precision(mdk_ctx *ctx, int double_precision);
MDK_API int mdk_set_params_f64(mdk_ctx *ctx, const double *charges, const double *eps_sigma);
/* topology.bonded_particles (1-2 and 1-3 partners: excluded, charmm_nonbonded_constraint.py:83)
 * and topology.scaling_particles (1-4 partners: eps14/sigma14, :92-97), int32, -1 padded,
 * row widths wb / ws (topology.py:69-79).  Either pointer may be NULL with width 0. */
MDK_API int mdk_set_exclusions(mdk_ctx *ctx, const int32_t *bonded, int wb, const int32_t *scaling, int ws);
/* k_e = 1/(4 pi eps0) in internal units, taken from mdpy's own EPSILON0 at run time
 * (electrostatic_constraint.py:21,60).  alpha / rc only matter for
 * MDK_TERM_COUL_DIRECT and MDK_TERM_PME_RECIP. */
MDK_API int mdk_set_coulomb(mdk_ctx *ctx, double k_e, double alpha, float rc);
/* PME mesh and B-spline order (4, 5, 6 or 8). */
MDK_API int mdk_set_pme(mdk_ctx *ctx, int nx, int ny, int nz, int order);
/* Verlet buffer of the tile list (the reference rebuilds its cell list on every
 * set_positions, state.py:61; here the list is reused until an atom has moved skin/2). */
MDK_API int mdk_set_nlist(mdk_ctx *ctx, float skin);
/* Bonded terms, parameters in internal units, one row per term:
 *   bonds     idx [n,2], par [n,2] = (k, r0)                E = k (r - r0)^2
 *   angles    idx [n,3], par [n,4] = (k, theta0, k_ub, r_ub) E = k (th - th0)^2 + k_ub (r13 - r_ub)^2
 *   dihedrals idx [n,4], par [n,3] = (k, n, delta)          E = k (1 + cos(n phi - delta))
 *   impropers idx [n,4], par [n,2] = (k, psi0)              E = k (psi - psi0)^2 */
MDK_API int mdk_set_bonded(mdk_ctx *ctx, int kind /* 0 bond 1 angle 2 dihedral 3 improper */, int n,
                   const int32_t *idx, const float *par);

/* Rigid three-site waters — what the reference's is_SHAKE flag asks for (forcefield/charmm_forcefield.py:24,32; the
 * reference itself has no constraint code).  triplets int32 [n,3] = (O, H, H) matrix ids, all molecules with the same
 * masses; d_oh / d_hh = the constrained O-H and H-H distances.  The Langevin step then integrates these molecules with
 * SETTLE (positions) + a RATTLE velocity projection fused into the position / velocity update; their bond / angle terms
 * must not be passed to mdk_set_bonded.  Positions that come from outside are projected onto the rigid geometry before
 * the next step call.  n_waters = 0 switches the constraints off.  Single domain only. */
MDK_API int mdk_set_rigid_waters(mdk_ctx *ctx, int n_waters, const int32_t *triplets, double d_oh, double d_hh);

/* ---- state ---- */
/* State.set_positions (state.py:56-61): wraps into [-L/2, L/2] (utils/pbc.py:28-36; an
 * atom >= 2 images away -> MDK_ERR_PARTICLE_LOST) and marks the tile list for a
 * displacement check.  xyz float32 [n,3] in topology (matrix_id) order. */
MDK_API int mdk_upload_positions(mdk_ctx *ctx, const float *xyz);
MDK_API int mdk_upload_positions_f64(mdk_ctx *ctx, const double *xyz);
MDK_API int mdk_upload_velocities(mdk_ctx *ctx, const float *v);
MDK_API int mdk_download_positions(mdk_ctx *ctx, float *xyz_wrapped);
MDK_API int mdk_download_positions_f64(mdk_ctx *ctx, double *xyz_unwrapped);
MDK_API int mdk_download_velocities(mdk_ctx *ctx, float *v);

/* ---- hot path ---- */
/* Force a neighbour (tile) list rebuild now.  Writes list statistics if non-NULL:
 * stats[0]=i-blocks, [1]=work units, [2]=j-chunks, [3]=masked chunks, [4]=pair slots. */
MDK_API int mdk_build_nlist(mdk_ctx *ctx, int64_t *stats);
/* Constraint.update for the selected terms (bit-or of MDK_TERM_*): zeroes the force
 * accumulator, rebuilds the tile list if an atom has moved more than skin/2, evaluates
 * the terms and writes their energies (float64, internal units) into
 * energies[MDK_NUM_ENERGIES] (may be NULL).  Synchronous. */
MDK_API int mdk_compute(mdk_ctx *ctx, unsigned terms, double *energies);
/* Constraint.forces: sum of the terms of the last mdk_compute, float32 [n,3] in
 * matrix_id order. */
MDK_API int mdk_download_forces(mdk_ctx *ctx, float *out);
MDK_API int mdk_download_forces_f64(mdk_ctx *ctx, double *out);

/* ---- integrators (the per-step position/velocity update) ---- */
/* VerletIntegrator.integrate (verlet_integrator.py:20-50), device resident.
 * reference_quirks != 0 reproduces the reference bit of behaviour the survey flags
 * (first step a dt^2, reported velocity (x_n+1 - x_n)/(2 dt)); 0 gives the
 * textbook initialisation (a dt^2/2) and central-difference velocities. */
MDK_API int mdk_step_verlet(mdk_ctx *ctx, double dt, int nsteps, unsigned terms, int reference_quirks);
MDK_API void mdk_verlet_reset(mdk_ctx *ctx); /* Integrator.erase_cache, integrator.py:20-22 */
/* G-JF Langevin with the reference's coefficients a, b, sigma
 * (langevin_integrator.py:23-31; textbook update), Philox4x32-10 noise. */
MDK_API int mdk_step_langevin(mdk_ctx *ctx, double dt, double kT, double gamma, uint64_t seed,
                      int nsteps, unsigned terms);
/* LangevinIntegrator.integrate(ensemble, nsteps) with the host State as input and output
 * (langevin_integrator.py:38-72 reads ensemble.state.positions / velocities and writes them back;
 * integrator.py:14-50).  x_in / v_in: float32 [n,3] host arrays (NULL = keep the device state);
 * an atom whose host value equals the float32 image of the device state keeps its float64 device
 * coordinate, so feeding back what the previous call returned continues the trajectory without a
 * restart; any other atom takes the host value and the step caches are dropped.  x_out / v_out:
 * float32 [n,3] (wrapped positions / velocities after the last step; NULL = skip), energies
 * [MDK_NUM_ENERGIES] or NULL.  Buffers from mdk_host_alloc are copied without staging.  One stream
 * synchronisation after the upload, one at the end. */
MDK_API int mdk_step_langevin_host(mdk_ctx *ctx, const float *x_in, const float *v_in, float *x_out, float *v_out,
                           double dt, double kT, double gamma, uint64_t seed, int nsteps, unsigned terms,
                           double *energies);
/* SteepestDescentMinimizer.minimize (mdpy/minimizer/steepest_descent_minimizer.py:30-53), device resident: up to
 * max_iterations of x_i += alpha F_i / |F_i| (per-atom unit force vectors, as the reference normalises them), one force
 * evaluation each, until |E_k - E_k-1| / |E_k-1| < energy_tolerance.  *iterations = iterations done;
 * energy_first_prev_last[3] = potential energy before the first, before the last and after the last iteration. */
MDK_API int mdk_minimize_sd(mdk_ctx *ctx, double alpha, double energy_tolerance, int max_iterations, unsigned terms,
                    int *iterations, double *energy_first_prev_last, double *energies);
/* Trajectory frames for the dumpers (reference: mdpy/dumper/*.py read ensemble.state.positions once per dump period).
 * stride > 0: every stride-th step of a following step call leaves the wrapped float32 positions [n,3] in a page-locked
 * ring of max_frames frames, copied out by a separate copy stream while the next steps run.  mdk_get_frames waits for
 * the copies, hands the frames over (at most max_frames into out; *n_frames = frames captured since the last hand-over)
 * and rewinds the ring.  stride = 0 switches the capture off.  Single domain. */
MDK_API int mdk_set_frame_capture(mdk_ctx *ctx, int stride, int max_frames);
MDK_API int mdk_get_frames(mdk_ctx *ctx, float *out, int max_frames, int *n_frames);
/* Page-locked host memory owned by the ctx (freed by mdk_destroy at the latest): State arrays that
 * live in it move to / from the device by DMA without a staging copy. */
MDK_API int mdk_host_alloc(mdk_ctx *ctx, size_t bytes, void **out);
MDK_API int mdk_host_free(mdk_ctx *ctx, void *ptr);
/* Energies of the most recent force evaluation inside a step call (no extra work). */
MDK_API int mdk_last_energies(mdk_ctx *ctx, double *energies);

/* ---- test / measurement hooks ---- */
/* The in-cutoff (rc of mdk_set_lj), non-excluded pair set the tile list yields, as
 * matrix_id pairs i<j (unsorted).  *n_out = count; at most cap are written. */
MDK_API int mdk_get_pairs(mdk_ctx *ctx, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out);
/* The same set as the PRODUCTION pair kernel decides it: a debug instantiation of k_pair (same template
 * arguments as the launch mdk_compute / the step graphs make for the current system: hoisted minimum
 * image where the box allows it, CHARMM switch, shared cutoff) writes out every slot that passes its own
 * cutoff / exclusion test, on the tile list and tile-order positions exactly as the last force evaluation
 * left them (no refresh, no rebuild: after a step call this is the list the last in-graph rebuild made). */
MDK_API int mdk_get_pairs_production(mdk_ctx *ctx, int32_t *out_i, int32_t *out_j, int64_t cap, int64_t *n_out);
/* Device time (ms, CUDA events on the ctx stream) of the last mdk_compute / step call,
 * per phase: [0]=nlist rebuild [1]=pair kernel [2]=pme spread [3]=fft+convolve
 * [4]=pme gather [5]=bonded+special pairs [6]=integrate [7]=bare coulomb
 * [8]=total [9]=NCCL all-reduce; plus counters [10]=kernel launches [11]=nlist rebuilds
 * [12]=pair-kernel launches [13]=work units [14]=j-chunks [15]=masked chunks [16]=chunks per unit
 * [17]=i-blocks. */
MDK_API int mdk_get_timing(mdk_ctx *ctx, double *out24);
/* Event timing level: 0 off (default), 1 whole-call CUDA events only, 2 per-phase events (adds stream syncs). */
MDK_API int mdk_set_profiling(mdk_ctx *ctx, int level);
/* Raw device pointer + element count of the int64 fixed-point force accumulator in
 * tile order (multi-GPU reduction by the host layer; scale = 2^40). */
MDK_API int mdk_force_accumulator(mdk_ctx *ctx, void **dev_ptr, int64_t *n_int64);
/* Execution options: key 0 = CUDA-graph integrator steps (default 1), 1 = PME / bonded kernels on side
 * streams beside the pair kernel (default 1), 2 = always apply the canonical minimum image per pair
 * (default 0: hoisted out of the pair loop when the box allows it), 3 = energy sums in every graph
 * step, 4 = NCCL inside the captured step (N > 1; default 0), 5 = persistent pair-kernel blocks per SM
 * (default 4), 6 = cuFFT also for small power-of-two meshes (default 0: fused mesh kernels),
 * 7 = shared-memory staged charge spreading (default 1; 0 = one global atomic per spline point), 8 = filter-then-compute
 * pair kernel (default 0 = the rotation-ring kernel; the variant measured slower), 9 = work units per
 * resident warp the list planner aims for (default 8), 10 = list order: the j-atoms of a block's list that have no i-atom
 * within the cutoff itself (skin shell) go last, into chunks of their own (default 1), 11 = work units a pair-kernel warp takes
 * before it retires (default 0 = persistent blocks fed by an atomic cursor; > 0 = short-lived blocks, which lets the side
 * streams' kernels in between), 12 = staging threshold of the list builder's far class (default 992 atoms per block part;
 * tests lower it to drive the early-flush path), 13 = decomposed step: spread after the halo exchange and send the sub-meshes in
 * an exchange of their own (default 0: spread first, sub-meshes and halo positions in one grouped exchange), 14 = decomposed
 * step, NCCL backend: the potential boxes return on the PME side stream and every rank posts its receive before it launches the
 * pair kernel (default 0: measured no gain at 8 GPUs, 4 % slower at 2). */
MDK_API int mdk_set_option(mdk_ctx *ctx, int key, double value);
/* Benchmark hygiene: overwrite a 256 MB scratch buffer on the ctx stream (evicts the 126 MB L2). */
MDK_API int mdk_flush_l2(mdk_ctx *ctx);
/* Multi-GPU: one process per GPU.  mdk_comm_unique_id wraps ncclGetUniqueId (rank 0 calls it and
 * ships the 128 bytes to the other ranks by any means, e.g. torch.distributed.broadcast);
 * mdk_comm_init joins the communicator. */
MDK_API int mdk_comm_unique_id(void *out128);
MDK_API int mdk_comm_init(mdk_ctx *ctx, int rank, int nranks, const void *unique_id128);
/* Spatial domain decomposition with halo exchange (mdk_dd.cu; the reference has one State on one device,
 * core/state.py:18-28).  The box is cut into px x py x pz domains, rank r = (rz * py + ry) * px + rx owns one:
 * it lists and evaluates the pair work of its own i-blocks, owns the bonded / excluded-pair terms of its atoms,
 * spreads / gathers them on the PME mesh and integrates only them.  Afterwards mdk_compute, mdk_step_langevin
 * and mdk_step_langevin_host of this ctx are collective calls (every rank makes the same call with the same
 * state); per step: grouped ncclSend / ncclRecv of halo positions and of halo forces, PME sub-meshes to and from
 * the mesh rank (the last one), an all-gather of the state at every list rebuild and at the end of a call.
 * local_group < 0: NCCL backend (mdk_comm_init first).  local_group >= 0: the ranks are contexts of THIS process on
 * ONE device (tests on a single-GPU box): transfers are device-to-device copies and the group is driven through
 * the *_group calls below. */
MDK_API int mdk_dd_init(mdk_ctx *ctx, int rank, int nranks, int px, int py, int pz, int local_group);
/* Local groups: ctxs = every context of the group in rank order; same semantics as mdk_compute / mdk_step_langevin. */
MDK_API int mdk_dd_compute_group(mdk_ctx *const *ctxs, int n, unsigned terms, double *energies);
MDK_API int mdk_dd_step_langevin_group(mdk_ctx *const *ctxs, int n, double dt, double kT, double gamma, uint64_t seed,
                               int nsteps, unsigned terms, double *energies);
/* out8: own tile slots [lo, hi), halo atoms received per step, own atoms sent per step, exchanges so far, list
 * rebuilds, PME sub-mesh points of this rank, ranks. */
/* Relative pair-work share of every rank's domain (nranks doubles, the same on all ranks; NULL = equal).  The domains are a
 * recursive bisection of the cell grid (x, then y inside each x slab, then z inside each column) whose volumes follow the
 * weights: the rank that also runs the PME mesh chain gets a smaller domain. */
MDK_API int mdk_dd_set_weights(mdk_ctx *ctx, const double *weights);
MDK_API int mdk_dd_stats(mdk_ctx *ctx, int64_t *out8);
/* Phase trace of the decomposed step (measurement hook): on != 0 makes every phase of the following calls end with a stream
 * synchronisation and accumulates its wall time; out16 (may be NULL) receives the sums so far in ms: halo positions,
 * rebuild (state gather / sort + lists / halo lists), O(N) terms + spreading, sub-meshes in, pair kernel, potential boxes out,
 * gather, halo forces, update, end of call, [12] = steps counted. */
MDK_API int mdk_dd_trace(mdk_ctx *ctx, int on, double *out16);

#ifdef __cplusplus
}
#endif
#endif /* MDPY_B200_H */

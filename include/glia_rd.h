/* glia_rd.h -- C ABI of libglia_rd.so: the B200-native (sm_100a) replacement for the
 * CUDA seam of GLIA's reaction-diffusion forward/adjoint hot path.
 *
 * The reference has no plugin / FFI layer: its seam is the `#ifdef CUDA` blocks inside
 * SpectralOperators, DiffCoef, DiffusionSolver and PdeOperatorsRD, which call free
 * functions on raw device pointers borrowed from PETSc Vecs (vecGetArray,
 * src/utils/Utils.cpp:69-93).  Each entry point below names the reference method it
 * replaces (paths relative to the GLIA repository root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (never exits the process,
 *     mirroring PetscErrorCode); glia_rd_last_error() returns the message;
 *   - `void*` field arguments are DEVICE pointers to `float` (precision 4) or `double`
 *     (precision 8) in the PETSc-Vec layout of the reference: C order [n0][n1][n2], z
 *     fastest (src/mat/DiffCoef.cpp:150); they stay owned by the caller;
 *   - calls are enqueued on the handle's stream and are synchronous at return;
 *   - the handle's stream is NON-BLOCKING: it has no implicit ordering with the legacy default
 *     stream.  Input fields must be complete when a call is made -- synchronise their producer, or
 *     order it in front of the library's work with glia_rd_wait_stream();
 *   - a handle is not thread-safe (one call at a time), but it may be used from any host thread:
 *     every entry point makes the handle's CUDA device current for the calling thread;
 *   - slab handles: create / resize_history / destroy are collective and must follow the
 *     disconnect -> barrier -> free order described at glia_rd_ipc_disconnect().
 */
#ifndef GLIA_RD_H
#define GLIA_RD_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glia_rd glia_rd_t;

#define GLIA_RD_F32 4
#define GLIA_RD_F64 8

/* history selectors for glia_rd_history() */
#define GLIA_HIST_C 0      /* c_[i],      i = 0..nt   (include/pde/PdeOperators.h) */
#define GLIA_HIST_P 1      /* p_[i],      i = 0..nt   */
#define GLIA_HIST_C_HALF 2 /* c_half_[i], i = 0..nt-1 */

/* Library / ABI version and build flavour ("cuda-sm_100a"). */
int glia_rd_abi_version(void);
const char* glia_rd_build_info(void);

/* Grid + plan set-up.  Replaces SpectralOperators::setup / initializeGrid
 * (src/grad/SpectralOperators.cpp:19-66, 424-446).  n[i] must be a power of two in
 * [32, 512].  `device` is the CUDA ordinal.  dt_ctx initialises the diffusion solver's
 * context time step exactly like params->tu_->dt_ does in the DiffusionSolver ctor
 * (src/pde/DiffusionSolver.cpp:11); pass 0.5 for the reference default. */
int glia_rd_create(glia_rd_t** h, const int n[3], int precision, int device, double dt_ctx);
/* Slab-decomposed handle: one process per GPU, the grid cut along x into `nranks` (1, 2, 4 or
 * 8) slabs, rank r owning x-planes [r*n0/nranks, (r+1)*n0/nranks) -- the local block AccFFT
 * hands each rank for c_dims = {nranks, 1} (accfftCreateComm / accfft_local_size_dft_r2c,
 * src/grad/SpectralOperators.cpp:24, 398-421; Grid::isize/istart, include/Parameters.h:374-451).
 * `n` stays the GLOBAL grid; every field pointer passed to a slab handle is that rank's local
 * block [n0/nranks][n1][n2].  Where the reference exchanges data with MPI all-to-all inside
 * every FFT, the ranks here map each other's field arenas (CUDA IPC over NVLink) and the
 * x-axis sweeps read / write the owners' memory directly.  Set-up is collective:
 *     create_slab  ->  ipc_export(0)  ->  [all-gather the 64-byte handles]  ->  ipc_connect(0)
 * and again with which = 1 after every glia_rd_resize_history.  Afterwards every call below
 * is collective over the ranks (same calls, same order), like the MPI reference. */
int glia_rd_create_slab(glia_rd_t** h, const int n[3], int precision, int device, double dt_ctx, int rank,
                        int nranks);
/* Ensemble handle (BASELINE config 5; the reference runs such ensembles as one process per member,
 * scripts/gridcont/run_sparsetil_multilevel_multigpu.py:25-135): `nbatch` INDEPENDENT members share one handle and one
 * set of kernel launches.  Every field argument is then [nbatch][n0][n1][n2] (member-major); each member has its own
 * k(x), rho(x), k-bar, PCG state and iteration count (a converged member stops while the others continue, exactly
 * as nbatch separate handles would).  Available on such a handle: the coefficient setters (glia_rd_set_coefficients_batch
 * for a (kappa, rho) pair per member over shared tissue maps; the plain setters take batch-sized fields), prec_factor,
 * diffusion_solve, reaction, apply_D, resize_history / history, solve_state(0 | 1), solve_adjoint, forward_adjoint.
 * The int iteration counts these return are sums over the members; glia_rd_batch_iterations gives them per member
 * (accumulated = 0: the last diffusion solve; 1: the totals of the last solve_state / solve_adjoint). */
int glia_rd_create_batch(glia_rd_t** h, const int n[3], int precision, int device, double dt_ctx, int nbatch);
int glia_rd_batch_size(glia_rd_t* h, int* nbatch);
int glia_rd_batch_iterations(glia_rd_t* h, int* its_per_member, int accumulated);
/* wm, gm, csf: ONE member's tissue maps [n0][n1][n2]; k_scale[nbatch], rho_scale[nbatch] (HOST arrays) */
int glia_rd_set_coefficients_batch(glia_rd_t* h, const void* wm, const void* gm, const void* csf, const double* k_scale,
                                   double k_gm_wm, double k_glm_wm, double filter_sum, const double* rho_scale,
                                   double r_gm_wm, double r_glm_wm);
#define GLIA_IPC_HANDLE_BYTES 64
/* which: 0 = work arena (after create), 1 = time-history arena (after resize_history) */
int glia_rd_ipc_export(glia_rd_t* h, int which, void* handle64);
/* handles: nranks * 64 bytes in rank order (the all-gathered exports) */
int glia_rd_ipc_connect(glia_rd_t* h, int which, const void* handles);
/* Closes this rank's mappings of the peers' arenas (which: 0 = work arena, 1 = time histories, -1 = both).
 * CUDA leaves freeing an exported allocation that another process still has open undefined, so teardown
 * and every re-allocation of the histories are
 *     every rank: glia_rd_ipc_disconnect  ->  caller-side barrier  ->  glia_rd_destroy / glia_rd_resize_history
 * (the reference's AccFFT / PETSc objects are collective in the same way: accfft_destroy_plan, VecDestroy). */
int glia_rd_ipc_disconnect(glia_rd_t* h, int which);
int glia_rd_destroy(glia_rd_t* h);
const char* glia_rd_last_error(const glia_rd_t* h);
/* Orders everything enqueued so far on `producer_stream` (a cudaStream_t; NULL = the legacy default stream)
 * in front of the handle's later work, without a host or device-wide synchronisation -- what the reference got
 * for free from default-stream semantics. */
int glia_rd_wait_stream(glia_rd_t* h, void* producer_stream);
/* the CUDA stream (cudaStream_t) all work of this handle is enqueued on */
void* glia_rd_stream(glia_rd_t* h);
/* number of kernel launches issued by this handle since creation */
long long glia_rd_launch_count(const glia_rd_t* h);

/* ---- L0: SpectralOperators ------------------------------------------------------- */
/* executeFFTR2C / executeFFTC2R (src/grad/SpectralOperators.cpp:68-98): unnormalised,
 * fhat is [n0][n1][n2/2+1] interleaved complex. */
int glia_rd_fft_r2c(glia_rd_t* h, const void* f, void* fhat);
int glia_rd_fft_c2r(glia_rd_t* h, const void* fhat, void* f);
/* computeGradient (SpectralOperators.cpp:100-177); xyz_mask bit0=x, bit1=y, bit2=z;
 * unrequested components may be NULL. */
int glia_rd_gradient(glia_rd_t* h, void* gx, void* gy, void* gz, const void* x, int xyz_mask);
/* computeDivergence (SpectralOperators.cpp:179-261). */
int glia_rd_divergence(glia_rd_t* h, void* div, const void* dx, const void* dy, const void* dz);

/* ---- L1: DiffCoef / ReacCoef ------------------------------------------------------ */
/* Isotropic diffusion coefficient field kxx=kyy=kzz=k and its three preconditioner
 * averages (kxx_avg_, kyy_avg_, kzz_avg_; DiffCoef.cpp:103-116).  The field is copied.
 * k_scale is DiffCoef::k_scale_ (solve() returns immediately when it is 0,
 * DiffusionSolver.cpp:227). */
int glia_rd_set_diffusion(glia_rd_t* h, const void* k, const double kavg[3], double k_scale);
/* DiffCoef::setValues (DiffCoef.cpp:77-131): k = k_scale*(wm + k_gm_wm*gm + k_glm_wm*csf),
 * negative ratios clamp to 0; averages = sum(k)/filter_sum (filter_sum = sum of the
 * brain mask, MatProp.cpp:180-185; pass n0*n1*n2 for setValuesSinusoidal semantics). */
int glia_rd_set_diffusion_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf,
                                 double k_scale, double k_gm_wm, double k_glm_wm, double filter_sum);
/* DiffCoef::setSecondaryCoefficients (DiffCoef.cpp:44-59): ktilde field (copied). */
int glia_rd_set_secondary_k(glia_rd_t* h, const void* ktilde);
/* ReacCoef::rho_vec_ (src/mat/ReacCoef.cpp:13-38); the field is copied. */
int glia_rd_set_reaction(glia_rd_t* h, const void* rho);
int glia_rd_set_reaction_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf,
                                double rho_scale, double r_gm_wm, double r_glm_wm);
/* PdeOperatorsMassEffect::updateReacAndDiffCoefficients (src/pde/PdeOperatorsMassEffect.cpp:98-138),
 * the coefficient refresh the mass-effect models do before every time step:
 *   rho = rho_scale * max(0, 1 - (bg + gm_r_scale*gm + vt + csf)),
 *   k   = k_scale   * max(0, 1 - (bg + gm_k_scale*gm + vt + csf)),
 * gm_r_scale = 1 - r_gm_wm_ratio, gm_k_scale = 1 - k_gm_wm_ratio.  One pass, both fields written
 * inside the handle.  As in the reference the averages k-bar used by glia_rd_prec_factor are NOT
 * recomputed (they keep the value of the last glia_rd_set_diffusion[_tissue]); the time step is then
 * glia_rd_prec_factor, [advection, not in this library], glia_rd_diffusion_solve(c, dt),
 * glia_rd_reaction(c, NULL, dt) -- PdeOperatorsMassEffect.cpp:578-631. */
int glia_rd_update_reac_diff(glia_rd_t* h, const void* bg, const void* gm, const void* vt, const void* csf,
                             double rho_scale, double k_scale, double gm_r_scale, double gm_k_scale);
/* DiffCoef::applyD / applyDWithSecondaryCoeffs (DiffCoef.cpp:249-300); dc may alias c. */
int glia_rd_apply_D(glia_rd_t* h, void* dc, const void* c, int secondary);

/* ---- L2: DiffusionSolver ---------------------------------------------------------- */
/* DiffusionSolver::precFactor (DiffusionSolver.cpp:119-180): freezes the preconditioner
 * symbol from the CURRENT context dt and averages (the stale-dt behaviour of the
 * reference is state and is reproduced on purpose). */
int glia_rd_prec_factor(glia_rd_t* h);
/* DiffusionSolver::solve(c, dt) (DiffusionSolver.cpp:217-250): Crank-Nicolson step by
 * PETSc-semantics preconditioned CG (rtol 1e-6, abstol 1e-50, maxit 5000, non-zero
 * initial guess, preconditioned norm).  c is updated in place; *ksp_its = ksp_itr_. */
int glia_rd_diffusion_solve(glia_rd_t* h, void* c, double dt, int* ksp_its);
int glia_rd_set_ksp_tolerances(glia_rd_t* h, double rtol, double abstol, double dtol, int maxit);

/* ---- L2a: PdeOperatorsRD ----------------------------------------------------------- */
/* PdeOperatorsRD ctor / resizeTimeHistory (src/pde/PdeOperators.cpp:17-103):
 * (nt+1) state + (nt+1) adjoint + nt half-step fields, zero-initialised. */
int glia_rd_resize_history(glia_rd_t* h, int nt, double dt);
int glia_rd_history(glia_rd_t* h, int which, int i, void** dev_ptr);
/* params->tu_->order_ (PdeOperators.cpp:271-290, 389-398): 2 = Strang splitting (default; half diffusion,
 * reaction, half diffusion), 1 = full-dt diffusion then reaction, in solve_state and solve_adjoint. */
int glia_rd_set_splitting_order(glia_rd_t* h, int order);
/* PdeOperatorsRD::reaction (PdeOperators.cpp:140-190): c_lin == NULL -> nonlinear
 * logistic step, else the linearised step about c_lin. */
int glia_rd_reaction(glia_rd_t* h, void* c_t, const void* c_lin, double dt);
/* PdeOperatorsRD::solveState(linearized) (PdeOperators.cpp:235-316).  c0 -> cT (may be
 * NULL); linearized 0/1/2; *ksp_its_total = diff_ksp_itr_state_. */
int glia_rd_solve_state(glia_rd_t* h, const void* c0, void* cT, int linearized, int* ksp_its_total);
/* PdeOperatorsRD::solveAdjoint(linearized) (PdeOperators.cpp:372-420).  pT -> p0 (may be
 * NULL); adjoint_store selects c_half_ (1) or the re-diffusion of c_[k] (0). */
int glia_rd_solve_adjoint(glia_rd_t* h, const void* pT, void* p0, int linearized, int adjoint_store,
                          int* ksp_its_total);

/* ---- L2b: gradient time integrals -------------------------------------------------- */
/* DerivativeOperators::gradDiffusion + gradReaction (src/grad/DerivativeOperators.cpp:
 * 189-321): out = h^3 * { <wm,Tk>, <gm,Tk>, <csf,Tk>, <wm,Tr>, <gm,Tr>, <csf,Tr> } with
 * Tk = sum_i w_i dt grad c_i . grad p_i, Tr = sum_i w_i dt p_i (c_i^2 - c_i) over the
 * stored histories (trapezoid weights).  The same call serves the Hkp / Hkk integrals of
 * evaluateHessian (DerivativeOperatorsRD.cpp:270-320, 355-407). */
int glia_rd_grad_kappa_rho(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double out[6]);

/* ---- L2b: DerivativeOperatorsRD drivers, in field space ------------------------------- */
/* The Phi basis (src/mat/Phi.cpp) is outside this library: the initial condition c(0) = Phi p is
 * an input field and the p-block of a gradient / Hessian product is returned as a field g with
 * g_p = Phi^T g.  `obs` is the observation mask of Obs::apply / applyT (src/mat/Obs.cpp:75-140;
 * NULL means O = I); regularisation is the L2 form (DerivativeOperatorsRD.cpp:36-39). */
/* DiffCoef::setSecondaryCoefficients (src/mat/DiffCoef.cpp:44-59): k~ = k1 wm + k2 gm + k3 csf
 * (the caller resolves the nk == 1 ratios). */
int glia_rd_set_secondary_tissue(glia_rd_t* h, const void* wm, const void* gm, const void* csf, double k1, double k2,
                                 double k3);
/* params->tu_->two_time_points_ (DerivativeOperatorsRD.cpp:30-34, 149-153, 216-222): data d0 at t = 0 with its
 * own observation mask (Obs::filter_0_, src/mat/Obs.cpp:60-100; NULL = identity).  Both fields are copied;
 * d0 == NULL switches the two-snapshot terms off again.  While on, glia_rd_hessian_matvec fails like the
 * reference's evaluateHessian does ("not implemented for two-snapshot scenario", :234). */
int glia_rd_set_two_snapshot(glia_rd_t* h, const void* d0, const void* obs0);
/* evaluateObjectiveAndGradient (DerivativeOperatorsRD.cpp:130-226): solveState(0), mismatch,
 * p_T = -O^T(O c(1) - d1), solveAdjoint(1), gradDiffusion + gradReaction.
 *   J[4]   = { J, h^3/2 ||O c(1) - d1||^2, beta/2 h^3 ||c0||^2, h^3/2 ||O0 c(0) - d0||^2 (0 unless two-snapshot) }
 *   g_c0   = -h^3 (alpha(0) - beta c0) [+ h^3 O0^T (O0 c0 - d0)]     (device field, may be NULL)
 *   g[6]   = as glia_rd_grad_kappa_rho;  ksp_its[2] = { state, adjoint } iteration totals. */
int glia_rd_objective_gradient(glia_rd_t* h, const void* c0, const void* d1, const void* obs, double beta,
                               const void* wm, const void* gm, const void* csf, double J[4], void* g_c0, double g[6],
                               int ksp_its[2]);
/* evaluateHessian (DerivativeOperatorsRD.cpp:229-438), Gauss-Newton product about the state of
 * the last objective_gradient call.  c0_tilde = Phi p~.  diffusivity_inversion = 0: y_c0 =
 * h^3 (beta c0~ - alpha~(0)) only.  Otherwise also the Hkp / Hpk / Hkk blocks with k~ from
 * glia_rd_set_secondary[_tissue]:  y_c0 -= h^3 alpha~_k(0),
 *   hk[0..2] = h^3 <wm|gm|csf, int grad c . grad alpha~ dt>      (Hkp p~)
 *   hk[3..5] = the same integral after the k~ solves              (Hkk k~)
 * including the reference's stale p_[nt] term.  ksp_its[4] = state(1), adjoint(2), state(2),
 * adjoint(2) iteration totals. */
int glia_rd_hessian_matvec(glia_rd_t* h, const void* c0_tilde, const void* obs, double beta, int diffusivity_inversion,
                           const void* wm, const void* gm, const void* csf, void* y_c0, double hk[6], int ksp_its[4]);

/* ---- callers either side of the path: smoother, MatProp, Phi (single-GPU handles) ---------- */
/* SpectralOperators::weierstrassSmoother (src/grad/SpectralOperators.cpp:263-381): periodic
 * Gaussian smoothing with width sigma; out may alias in; sigma == 0 copies. */
int glia_rd_smooth(glia_rd_t* h, void* out, const void* in, double sigma);
/* MatProp::setValuesCustom (src/mat/MatProp.cpp:135-201): clips gm / wm / vt / csf at 0 IN PLACE
 * (null = absent map), bg = 1 - sum, filter = (wm > 0.1 || gm > 0.1) && vt < 0.8 (bg, filter may
 * be null); *filter_sum = sum(filter), the divisor of DiffCoef's average coefficients. */
int glia_rd_mat_prop(glia_rd_t* h, void* gm, void* wm, void* vt, void* csf, void* bg, void* filter,
                     double* filter_sum);
/* Phi::setGaussians / setValues (src/mat/Phi.cpp:24-120): np Gaussians with HOST centres
 * centers[3*np] (radians) and width sigma_phi, the MatProp filter (device field, copied; null =
 * none) and sigma_smooth = smoothing_factor * 2 pi / n0 (Phi.cpp:338). */
int glia_rd_phi_set(glia_rd_t* h, int np, const double* centers, double sigma_phi, const void* filter,
                    double sigma_smooth);
/* Phi::apply in on-the-fly mode (Phi.cpp:324-383): out = sum_i p_i phi_i / max_i max(phi_i),
 * phi_i = truncate_{5 sigma}(W(Gaussian_i . filter)); p is a HOST array of np values. */
int glia_rd_phi_apply(glia_rd_t* h, void* out, const double* p);
/* Phi::applyTranspose (Phi.cpp:385-434): pout_i = <phi_i, in> / max_i max(phi_i) (HOST, np). */
int glia_rd_phi_apply_transpose(glia_rd_t* h, double* pout, const void* in);

/* ---- data formats either side of the path ---------------------------------- */
/* dataIn (src/utils/IO.cpp:511-538, PnetCDF there): reads variable "data" (dims x, y, z; any numeric
 * nc_type; NetCDF classic CDF-1 / CDF-2 / CDF-5 headers) into the device field -- a slab handle reads
 * its own x rows, like ncmpi_get_vara_all with the rank's istart / isize. */
int glia_rd_data_in(glia_rd_t* h, const char* path, void* field);
/* dataOut (IO.cpp:540-612): writes the device field as a CDF-2 file laid out like the reference's
 * (dims x y z, "data" NC_FLOAT / NC_DOUBLE, global attribute "CDF-5 mode" = 0); collective on slab
 * handles (every rank writes its rows). */
int glia_rd_data_out(glia_rd_t* h, const char* path, const void* field);
/* splitSegmentation, atlas form (src/utils/Utils.cpp:592-657 with tu == ed == nullptr): one-hot maps
 * for labels = {wm, gm, vt, csf} (a label <= 0 gives a zero map; null outputs are skipped). */
int glia_rd_split_segmentation(glia_rd_t* h, const void* seg, const int labels[4], void* wm, void* gm, void* vt,
                               void* csf);

/* ---- per-kernel profile (CUDA events around every launch on the handle's stream) ---- */
/* begin: start recording; end: stop, and write one "tag launches total_ms" line per kernel
 * family into buf (NUL-terminated, truncated to buflen).  Replaces the reference's
 * EventRegistry timers (3rdparty/timings/EventTimings.hpp:140-176) for this path. */
int glia_rd_profile_begin(glia_rd_t* h);
int glia_rd_profile_end(glia_rd_t* h, char* buf, int buflen);

/* measurement probe of the slab x sweeps (collective; the work buffers are overwritten):
 * what 0 = preconditioner x sweep, 1 = D-apply x sweep; local_mask redirects the peer reads (1)
 * and / or writes (2) to local memory, to separate NVLink pull, push and compute time. */
int glia_rd_probe_xsweep(glia_rd_t* h, int what, int local_mask, int reps, double* ms_per_sweep);

/* ---- timing helper (CUDA events on the handle's stream) --------------------------- */
int glia_rd_timer_start(glia_rd_t* h);
int glia_rd_timer_stop_ms(glia_rd_t* h, double* ms);

/* ---- one objective-gradient evaluation's PDE work ------------------------------------ */
/* solveState(0), terminal condition p_T = -(c(T) - d1) (O = I; DerivativeOperatorsRD.cpp:
 * 155-162), solveAdjoint(1) -- the state/adjoint pair every evaluateObjectiveAndGradient
 * performs.  DEVICE buffers; cT and p0 may be NULL; histories are left in the handle. */
int glia_rd_forward_adjoint(glia_rd_t* h, const void* c0, const void* d1, void* cT, void* p0,
                            int* ksp_state, int* ksp_adj);
/* The same with HOST buffers (plugin-style end-to-end call): H2D of c0 and d1 and D2H of
 * c(T) and p(0) happen inside the call; page-locked buffers are DMA'd directly. */
int glia_rd_forward_adjoint_host(glia_rd_t* h, const void* c0_host, const void* d1_host, void* cT_host,
                                 void* p0_host, int* ksp_state, int* ksp_adj);

#ifdef __cplusplus
}
#endif
#endif /* GLIA_RD_H */

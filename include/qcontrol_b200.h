/*
 * qcontrol_b200 -- C ABI of the B200-native propagation hot path.
 *
 * The reference (irenemizus/qcontrol) is pure Python; it has no FFI.  The seam
 * this library replaces is the Python class protocol around
 *   phys_base.prop_cpu            (phys_base.py:129-178)   one polynomial step
 *   PropagationSolver.step        (propagation.py:366-459) one time step
 *   FittingSolver sweeps          (fitter.py:213-422, 650-811) Krotov / UT sweeps
 * so every entry point below names the reference code it stands in for.
 *
 * Conventions
 *   - every function returns 0 on success, a negative qc_status otherwise;
 *     qc_last_error() gives the message of the calling thread's last failure.
 *   - all host buffers are caller-owned, C-contiguous, little-endian float64;
 *     complex values are interleaved (re, im) pairs  == numpy complex128.
 *   - the library owns all device buffers (inside a plan); nothing here takes
 *     or returns a torch type.
 *   - a context binds one CUDA device and one stream.  Calls on one context are
 *     ordered on that stream; set/get calls are synchronous with the host,
 *     sweep calls are asynchronous unless stated.  Different contexts may be used
 *     from different host threads (the reference farms runs over a thread pool,
 *     newcheb.py:656-659); the library keeps no unguarded global state.
 */
#ifndef QCONTROL_B200_H
#define QCONTROL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QC_ABI_VERSION 1

typedef enum {
    QC_OK = 0,
    QC_ERR_ARG = -1,      /* bad argument / unsupported size          */
    QC_ERR_CUDA = -2,     /* CUDA runtime failure                      */
    QC_ERR_STATE = -3,    /* call order (e.g. sweep before set_static) */
    QC_ERR_NOMEM = -4
} qc_status;

/* hamil_2d.py functor classes (hamil_2d.py:190-223) */
typedef enum {
    QC_HAMIL_NONTRIV = 0,   /* Hamil2DNonTrivial : FFT grid, 2 levels   (hamil_2d.py:34-73)  */
    QC_HAMIL_TWO_LEVELS = 1,/* Hamil2DTwoLevels                        (hamil_2d.py:76-97)   */
    QC_HAMIL_BH_X = 2,      /* Hamil2DBHX                              (hamil_2d.py:100-127) */
    QC_HAMIL_BH_Z = 3,      /* Hamil2DBHZ                              (hamil_2d.py:130-157) */
    QC_HAMIL_QUBIT = 4      /* Hamil2DQubit                            (hamil_2d.py:160-187) */
} qc_hamil;

/* how the field of a step is obtained (fitter.py:650-811) */
typedef enum {
    QC_FIELD_PRESCRIBED = 0,   /* E[l] supplied by the caller (single_pot, filtering, trans_wo_control,
                                  intuitive_control, Krotov iteration 0 forward, UT backward)            */
    QC_FIELD_KROTOV_FWD = 1,   /* E = -2 Im<chi_old[nt-l].f[1] | psi_w.f[0]> / h_lambda   (fitter.py:700-706) */
    QC_FIELD_KROTOV_BWD = 2,   /* E = -2 Im<chi_w.f[1] | psi_old[nt-l].f[0]> / h_lambda   (fitter.py:791-796) */
    QC_FIELD_UT_FWD = 3,       /* E = E_base[l] - s[l] Im(sum_v a0[v] <chi[-l]|psi_cur>) / h_lambda
                                  (fitter.py:739-778)                                                     */
    QC_FIELD_LOCAL_PROJ = 4    /* local_control_projection: E[l] supplied; the carrier-frequency multiplier
                                  follows its ODE towards the value fed back from the previous step, so eL,
                                  exp_L, the energy window and the Newton table are rebuilt on the device
                                  every step (fitter.py:598-623, 814-836; propagation.py:379-407)         */
} qc_field_mode;

typedef struct qc_context qc_context;
typedef struct qc_plan qc_plan;

typedef struct {
    int32_t hamil;      /* qc_hamil                                                  */
    int32_t np;         /* grid points: 256..4096 (power of two) for NONTRIV (4096: two-CTA cluster per run), 1 else */
    int32_t nlevs;      /* 2 for NONTRIV/TWO_LEVELS; 2..8 for BH_X/BH_Z/QUBIT        */
    int32_t nb;         /* basis vectors per run (1 for NONTRIV)                     */
    int32_t nch;        /* interpolation points (power of two, <= 128)               */
    int32_t batch;      /* independent runs resident in this plan                    */
    int32_t nt_max;     /* longest time grid of any run (trajectory capacity)        */
    int32_t hf_hide;    /* rotating-frame transform per step (propagation.py:383-402)*/
    int32_t store_traj; /* 1: allocate the two trajectory buffers (Krotov / UT)      */
    int32_t reserved;
} qc_plan_desc;

/* ---- library / context --------------------------------------------------- */
int32_t qc_abi_version(void);
const char *qc_last_error(void);
/* stream: a cudaStream_t cast to void* (e.g. torch.cuda.current_stream().cuda_stream) or
 * NULL to let the library create its own non-blocking stream. */
int32_t qc_create(int32_t device, void *stream, qc_context **out);
int32_t qc_destroy(qc_context *ctx);
int32_t qc_synchronize(qc_context *ctx);
/* CUDA-event timing on the context's stream (used by bench.py for the roofline figure) */
int32_t qc_timer_start(qc_context *ctx);
int32_t qc_timer_stop(qc_context *ctx, double *elapsed_ms); /* synchronises */
/* kernels launched by this context since creation (bench.py "gpu_launches") */
int64_t qc_launch_count(qc_context *ctx);
/* FP64 FMA throughput micro-benchmark on this device, TFLOP/s (roofline denominator) */
int32_t qc_measure_fp64_peak(qc_context *ctx, double *tflops);
/* free / total device memory of the context's GPU in bytes: the reference keeps every trajectory in host RAM
 * (psi_tlist / chi_tlist, fitter.py:322-329); here they live in HBM, so batch entry points size their plans with
 * this and qc_plan_footprint instead of failing with QC_ERR_NOMEM */
int32_t qc_mem_info(qc_context *ctx, int64_t *free_bytes, int64_t *total_bytes);
/* number of SMs of the context's GPU: batch entry points cut big batches into pieces of whole waves of CTAs so that the
 * host set-up of one piece overlaps the propagation of the previous one (the reference farms its jobs over a thread
 * pool instead, newcheb.py:656-659) */
int32_t qc_sm_count(qc_context *ctx, int32_t *sm_count);
/* Reserve `bytes` of device memory as an arena from which the plans of this device are cut until it is given back
 * with bytes = 0 (or the last context of the device is destroyed).  A batch entry point that sets the next sub-batch up
 * beside the propagation of the current one reserves its working set while the GPU is idle: an allocation by the
 * driver beside a running sweep returns only when the sweep ends.  Requests the arena cannot serve fall back to the
 * driver. */
int32_t qc_pool_reserve(qc_context *ctx, int64_t bytes);

/* ---- plan ---------------------------------------------------------------- */
int32_t qc_plan_create(qc_context *ctx, const qc_plan_desc *desc, qc_plan **out);
int32_t qc_plan_destroy(qc_plan *plan);
/* upper bound of the device bytes a plan of this shape allocates over its life (needs no GPU) */
int32_t qc_plan_footprint(const qc_plan_desc *desc, int64_t *bytes);

/* Potentials and kinetic multiplier: what Hamil2D*.__init__(v, akx2, np, U, W, delta, ntriv)
 * receives (hamil_2d.py:190-199).
 *   v        [v_batched ? batch : 1][nlevs][np]   the arrays v[n][1]
 *   akx2_re  [np]  real part of akx2 (its imaginary part is -0, task_manager.py:501-504); may be NULL if np == 1
 *   uwd      [uwd_batched ? batch : 1][3]  (U, W, delta) for BH_X / BH_Z / QUBIT; may be NULL otherwise
 *   dx       [dx_batched ? batch : 1]      grid step used by every inner product (math_base.py:7-21) */
int32_t qc_set_static(qc_plan *plan, const double *v, int32_t v_batched, const double *akx2_re,
                      const double *uwd, int32_t uwd_batched, const double *dx, int32_t dx_batched);

/* Interpolation data of one sweep direction: the outputs of math_base.points (math_base.py:153-188)
 * computed on the host, plus the scalars residum_cpu / prop_cpu derive from emin, emax, t_sc
 * (phys_base.py:100-101, 174).
 *   slot     0 or 1 (e.g. forward / backward coefficients kept side by side)
 *   xp       [nch]                      nodes (shared: they depend on nch only)
 *   dv       [batched ? batch : 1][nch] complex divided differences
 *   sc       [batched ? batch : 1][4]   (emin, emax, t_sc, eL)  eL = nu_L*freq_mult*Hz_to_cm/2 (propagation.py:379) */
int32_t qc_set_poly(qc_plan *plan, int32_t slot, const double *xp, const double *dv, const double *sc,
                    int32_t batched);

/* OPT-IN truncation of the interpolation series: orders[batched ? batch : 1] = number of terms dv_0 .. dv_{J-1}
 * actually summed by sweeps with this slot (2 <= J <= nch); NULL restores the full series.  The reference always sums
 * all nch terms (phys_base.py:165), although math_base.points leaves divided differences of round-off size (~1e-17)
 * behind the converged part; the host chooses J from a rigorous bound on the dropped tail
 * (qcontrol_b200.math_base.truncation_order).  Off unless called: results then differ from the reference's by that
 * bound per step instead of by round-off. */
int32_t qc_set_poly_orders(qc_plan *plan, int32_t slot, const int32_t *orders, int32_t batched);

/* State psi[batch][nb][nlevs][np] complex: DynamicState.psi (propagation.py:55-71). */
int32_t qc_set_state(qc_plan *plan, const double *psi);
int32_t qc_get_state(qc_plan *plan, double *psi);

/* Per-step scalars known before the sweep (they depend on t only):
 *   which = 0  E_base [b][nt_max+1] complex : prescribed field / UT base field E_tlist   (fitter.py:653-658, 775-778)
 *   which = 1  exp_L  [b][nt_max+1] complex : csqrt(laser_field_hf(t_l)) used with poly slot 0 (propagation.py:385)
 *   which = 3  exp_L  for poly slot 1 (the opposite time direction has different t_l)
 *   which = 2  s_env  [b][nt_max+1] complex : (envelope(t_l - |dt|/2)/E0, 0)              (fitter.py:753-754)
 * b = batched ? batch : 1. */
int32_t qc_set_step_scalars(qc_plan *plan, int32_t which, const double *vals, int32_t batched);

/* Per-run control scalars: ctl[batched ? batch : 1][4] = (h_lambda, nt, retired, reserved);
 * retired != 0 (N-level plans): the run takes no further part in sweeps -- the control loop of a batch keeps
 * iterating the runs that have not converged yet (fitter.py:545-546) while finished ones keep their slot;
 * a0[batch][nb] complex = aF_type(...) (task_manager.py:209-268), may be NULL. */
int32_t qc_set_control(qc_plan *plan, const double *ctl, int32_t batched, const double *a0);

typedef struct {
    int32_t poly_slot;    /* which qc_set_poly slot                                        */
    int32_t field_mode;   /* qc_field_mode                                                 */
    int32_t l_first;      /* first step index l (1-based as DynamicState.l, propagation.py:364) */
    int32_t nsteps;       /* steps to run in this launch (runs with nt < l stop early)     */
    int32_t renorm;       /* 1: renormalise every step (propagation.py:424-433, 450-459)   */
    int32_t traj_read;    /* buffer (0/1) holding the previous opposite sweep, -1: none    */
    int32_t traj_write;   /* buffer (0/1) to record this sweep into, -1: none              */
    int32_t reversed_E;   /* 1: prescribed field is read as E_base[nt - l] ... (fitter.py:804-807) */
    int32_t write_E_base; /* 1: copy the field actually used back into E_base[l] (fitter.py:336) */
    int32_t reserved[3];
} qc_sweep_args;

/* Run nsteps time steps of every run in the plan: PropagationSolver.step (propagation.py:366-459)
 * including the laser_field_envelope callback of the given mode.  Asynchronous. */
int32_t qc_sweep(qc_plan *plan, const qc_sweep_args *args);

/* One polynomial step on the resident state with an explicit field, no transform, no
 * renormalisation: the drop-in for prop_cpu(psi, hamil2D, t_sc, nch, np, emin, emax, E, eL, E_full,
 * orig=False) (phys_base.py:129-178).  E[batch] complex (real part used on the grid). Synchronous. */
int32_t qc_prop_step(qc_plan *plan, int32_t poly_slot, const double *E);

/* One application of an operator to the resident state, result to the host, state unchanged:
 *   which = 0  phi = H psi : the functor call hamil2D(orig, psi, E, eL, E_full) (hamil_2d.py:190-223);
 *              E[batch] complex is E (orig = 0, real part used on the grid) or E_full (orig = 1)
 *   which = 1  phi = T psi = ifft(akx2 * fft(psi))             (phys_base.diff_cpu, phys_base.py:18-36)
 *   which = 2  phi = P psi with the momentum multiplier of qc_set_instr_grid   (phys_base.py:238-244)
 * out[batch][nb][nlevs][np] complex. */
int32_t qc_apply_hamil(qc_plan *plan, int32_t which, int32_t orig, const double *E, double eL, double *out);

/* Grid data only the instrumentation needs: x[np] and the real momentum multiplier
 * akx[np] = initak(np, dx, 1, ntriv) * hart_to_cm / (-i) / dalt_to_au  (phys_base.py:238-240); akx may be NULL. */
int32_t qc_set_instr_grid(qc_plan *plan, const double *x, const double *akx_re);

/* The instrumentation block of PropagationSolver.step (propagation.py:462-470) on the device:
 * out[batch][nb][nlevs][20] = cnorm(2) cener(2) overlp0(2) overlpf(2) <x>(2) <x^2>(2) <p>(2) <p^2>(2)
 * max|psi|, Re psi at argmax|Re psi|, cprod(psi_0, psi_1)(2).  E_full[batch] complex (may be NULL = 0);
 * two_m = 2*m; psi0_slot / psif_slot: snapshot slots of psi0 / psif (-1 to skip). */
int32_t qc_instrument(qc_plan *plan, const double *E_full, double two_m, int32_t psi0_slot, int32_t psif_slot,
                      double *out);

/* Local control with a fed-back frequency multiplier (grid plans with hf_hide, QC_FIELD_LOCAL_PROJ sweeps).
 * params[batch][12] = the four energy extremes of propagation.py:371-375 (v0[0]+|akx2[np/2-1]|, vmin0,
 * v1[0]+|akx2[np/2-1]|, vmin1), E0, nu_L, carrier frequency nu, k_E, lamb, pow, epsilon, signed dt;
 * t_tab[batch][nt_max+1] = the step times t_l; xp[nch] = the interpolation nodes (math_base.py:166-170).
 * Resets the state of FitterDynamicState (fitter.py:22-31): freq_mult = freq_mult_patched = 1, velocities 0.
 * The multiplier used at step l is returned in the imaginary slot of E_out[run][l] (qc_get_fields). */
int32_t qc_set_local_control(qc_plan *plan, const double *params, const double *t_tab, const double *xp);
/* out[batch][11] = freq_mult, freq_mult_vel, freq_mult_patched, dAdt_happy, error (0 ok, 1 "freq_mult_patched
 * has to be positive or zero", 2 "Frequency multiplier has to be positive or zero": the run stopped there, like
 * the RuntimeError of fitter.py:817-821), psigc_psie (2), psigc_dv_psie (2), dA/dt of the last step, error step. */
int32_t qc_get_local_state(qc_plan *plan, double *out);
/* Overwrite the feedback state: state[batch][4] = freq_mult, freq_mult_vel, freq_mult_patched, dAdt_happy
 * (restart of a run from a stored time step; FitterDynamicState.from_json, fitter.py:33-50). */
int32_t qc_set_local_state(qc_plan *plan, const double *state);

/* Field used at each step of the last sweep(s): E_out[batch][nt_max+1] complex (E_tlist, fitter.py:281,326). */
int32_t qc_get_fields(qc_plan *plan, double *E_out);
/* sum_{l=1..nt} |E_l|^2 of the fields of the last sweep, per run: out[batch] (E_int / dt, fitter.py:338-341),
 * so that a batch can iterate without copying whole field lists to the host. */
int32_t qc_field_integral(qc_plan *plan, double *out);
/* Per-level norms before renormalisation, norms[batch][nb][nlevs] of the LAST step run (cnorm, propagation.py:424-429). */
int32_t qc_get_norms(qc_plan *plan, double *norms);

/* Device-resident snapshots of a full state (slots 0..3): the initial basis psi_init and the goal basis
 * psi_goal stay in HBM across iterations (FittingSolver keeps them as psi_init_basis / psi_goal_basis,
 * fitter.py:198-199).  psi_host == NULL snapshots the current device state. */
int32_t qc_snapshot_set(qc_plan *plan, int32_t slot, const double *psi_host);
/* state <- snapshot (the deep copies at the start of every sweep, fitter.py:227, 242) */
int32_t qc_snapshot_load(qc_plan *plan, int32_t slot);

/* <goal|psi> per run and basis vector: gc[batch][nb] complex = sum_n cprod(goal.f[n], psi.f[n])
 * (fitter.py:373-374) with goal = snapshot goal_slot.  gc may be NULL (result stays on the device for
 * qc_krotov_terminal). */
int32_t qc_goal_overlap(qc_plan *plan, int32_t goal_slot, double *gc);

/* Krotov terminal condition chi(T) (fitter.py:376-392): chi_l = 0, chi_u = gc * goal_u renormalised;
 * the state becomes chi(T) and trajectory buffer traj_write[0] receives chi_u * csqrt(hf(T)).
 * gc[batch] complex or NULL (= result of the last qc_goal_overlap); sqrt_hf_T[batch or 1] complex. */
int32_t qc_krotov_terminal(qc_plan *plan, int32_t goal_slot, const double *gc, const double *sqrt_hf_T,
                           int32_t hf_batched, int32_t traj_write);

/* Copy one recorded trajectory slice to the host: out[batch][nrec][np] complex where nrec is nb*nlevs
 * for N-level plans and 1 (the recorded level) for grid plans. */
int32_t qc_get_traj(qc_plan *plan, int32_t buffer, int32_t index, double *out);
/* Store the current state (lab frame) as slice `index` of a trajectory buffer (psi_tlist[0] = psi_init,
 * fitter.py:227-229; chi_tlist[0] = psi_goal for UT, fitter.py:241-244). level: recorded level for grid plans. */
int32_t qc_record_state(qc_plan *plan, int32_t buffer, int32_t index, int32_t level);

/* End-to-end convenience used by the reference-facing plugin layer and bench.py's e2e figure:
 * H2D of psi, nsteps prescribed-field steps, D2H of psi, all on the context's stream. Synchronous. */
int32_t qc_propagate_host(qc_plan *plan, const qc_sweep_args *args, const double *psi_in, double *psi_out);

/* ---- two-dimensional grids (synthetic extension: "hamil_2d 512x512" of BASELINE config 5; the reference has no
 * 2-D operator) ----------------------------------------------------------------------------------
 * One-level wavefunctions psi[batch][n][n] (row-major, y contiguous) under H = T_x + T_y + V(x,y), propagated by the
 * same Newton polynomial as phys_base.prop_cpu (phys_base.py:129-178) with the 2-D analogue of hamil_cpu
 * (phys_base.py:39-70): T_x = ifft_x(akx2 * fft_x .), T_y likewise.  n = 512. */
typedef struct qc_plan2d qc_plan2d;
int32_t qc_plan2d_create(qc_context *ctx, int32_t n, int32_t nch, int32_t batch, qc_plan2d **out);
int32_t qc_plan2d_destroy(qc_plan2d *plan);
/* v[batch|1][n][n] potential, akx2_re[n] / aky2_re[n] kinetic multipliers (math_base.initak scaled as in
 * task_manager.py:501-504), dxdy = area element of the norm */
int32_t qc_plan2d_set_static(qc_plan2d *plan, const double *v, int32_t v_batched, const double *akx2_re,
                             const double *aky2_re, double dxdy);
/* xp[nch], dv[nch] complex (math_base.points), sc = {emin, emax, t_sc} */
int32_t qc_plan2d_set_poly(qc_plan2d *plan, const double *xp, const double *dv, const double *sc);
int32_t qc_plan2d_set_state(qc_plan2d *plan, const double *psi);
int32_t qc_plan2d_get_state(qc_plan2d *plan, double *psi);
/* nsteps time steps of every wavefunction; renorm != 0: divide by sqrt(|<psi|psi>|) after each (propagation.py:450-459) */
int32_t qc_plan2d_sweep(qc_plan2d *plan, int32_t nsteps, int32_t renorm);
/* norms[batch]: <psi|psi> of the last step before renormalisation */
int32_t qc_plan2d_get_norms(qc_plan2d *plan, double *norms);

#ifdef __cplusplus
}
#endif
#endif /* QCONTROL_B200_H */

/*
 * mpcb200.h -- C ABI of libmpcb200.so: the B200-native replacement for the MPC solve loop of
 * AutomationLabsModelPredictiveControl.jl (reference paths below are relative to /root/reference).
 *
 * The reference has no FFI of its own: its hot path is `JuMP.optimize!(C.tuning.modeler)`
 * (src/main/computation_mpc.jl:41) into OSQP / Ipopt.  The seam this library plugs into is the solver table
 * `_IMPLEMENTATION_SOLVER_LIST` + `_selection_solver_JuMP_model` (src/sub/solver_selection.jl:9-14, 92-114) and the
 * untyped `ModelPredictiveControlTuning.modeler::Any` (src/types/types.jl:115): with `mpc_solver = "b200"` the
 * modeler holds an opaque `mpcb_handle*` instead of a JuMP model, and `update_initialization!` / `calculate!`
 * (src/main/computation_mpc.jl:17-55) `ccall` the entry points declared here.
 *
 * Conventions: plain C, no C++ types, no exceptions cross the boundary.  All matrices are COLUMN-MAJOR Float64
 * exactly as Julia stores them.  Every function returning `int` returns 0 on success and a negative code on failure;
 * `mpcb_last_error()` then holds a thread-local message.  Per-problem solver outcomes are DATA (status[] uses OSQP's
 * codes), not errors -- the reference itself never inspects termination_status (computation_mpc.jl:41-53).
 * A handle is not re-entrant (one call in flight per handle, like a JuMP model); distinct handles may be used from
 * distinct threads.  The caller owns every array it passes for the duration of the call only.
 */
#ifndef MPCB200_H
#define MPCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_VERSION 100 /* 0.1.0 */

/* error codes */
#define MPCB_OK 0
#define MPCB_ERR_INVALID -1   /* bad argument / unsupported configuration */
#define MPCB_ERR_CUDA -2      /* CUDA runtime failure (message has the cudaError string) */
#define MPCB_ERR_NUMERIC -3   /* host-side design failed (DARE diverged, K not positive definite, ...) */
#define MPCB_ERR_NO_DEVICE -4 /* no usable sm_100 device: there is NO CPU fallback */

/* per-problem status, OSQP's codes (what JuMP.termination_status would be derived from) */
#define MPCB_STATUS_SOLVED 1
#define MPCB_STATUS_SOLVED_INACCURATE 2 /* NMPC only: the SQP line search found no further descent (kink of a relu network) */
#define MPCB_STATUS_MAX_ITER -2
#define MPCB_STATUS_PRIMAL_INFEASIBLE -3
#define MPCB_STATUS_DESIGN_FAILED -20 /* re-linearised solve: this problem's Riccati equation has no stabilising solution (outside OSQP's code range: -4 is its dual-infeasible) */
#define MPCB_STATUS_UNSOLVED -10

/* terminal ingredient (src/sub/design_mpc.jl:298-394).  "contractive" (e_H' e_H <= 0.9 e_0' e_0, :333-340) is a quadratic
 * constraint OSQP cannot take; here it is a ball projection inside the ADMM (linear path, on-chip kernel: nz + nx <= 64).
 * "neighborhood" is unimplemented in the reference (:342-345) -> rejected with MPCB_ERR_INVALID. */
#define MPCB_TERMINAL_NONE 0
#define MPCB_TERMINAL_EQUALITY 1
#define MPCB_TERMINAL_CONTRACTIVE 2

/* kernel selection */
#define MPCB_KERNEL_AUTO 0
#define MPCB_KERNEL_ONCHIP 1   /* register/shared-memory resident DMMA ADMM (nz + m_g <= 64) */
#define MPCB_KERNEL_STREAMED 2 /* per-iteration FP64 tensor GEMM over HBM/L2-resident state */
#define MPCB_KERNEL_ONCHIP_SMEM 3 /* shared-memory resident DMMA ADMM: box-only problems with 64 < nz <= 120, and (round 2) problems WITH general
                                     rows -- terminal equality, contractive ball, state box -- with 48 < nz + m_g <= 120 by default (from 32 when
                                     asked for): the operator stays in shared memory, the state in per-warp slices, no host synchronisation */
#define MPCB_KERNEL_RICCATI 4     /* stage-wise ("sparse") ADMM: the x-update by a cached Riccati sweep over the horizon, O(H) per iteration;
                                     box-only problems without the S term, small (nx, nu); the long-horizon kernel (linear.jl:48-60 is the
                                     reference's own stage-wise formulation) */

typedef struct mpcb_handle mpcb_handle;

/* Solver settings.  Defaults (mpcb_default_settings) are OSQP 0.6's -- what the reference gets because
 * solver_selection.jl:94-95 sets no attribute -- except rho, which is chosen per system because the KKT factor is
 * cached once for the whole batch (rho <= 0: sqrt(lambda_min * lambda_max) of the condensed Hessian). */
typedef struct {
  double eps_abs;       /* 1e-3 */
  double eps_rel;       /* 1e-3 */
  double eps_prim_inf;  /* 1e-4 */
  double rho;           /* <= 0 : automatic */
  double rho_eq_scale;  /* 1e3  (OSQP RHO_EQ_OVER_RHO_INEQ) */
  double sigma;         /* 1e-6 */
  double alpha;         /* 1.6 */
  int32_t max_iter;     /* 4000 (rounded up to a multiple of check_every) */
  int32_t check_every;  /* 25 */
  int32_t device;       /* CUDA device ordinal this handle lives on */
  int32_t kernel;       /* MPCB_KERNEL_* */
  int32_t ladder_iter;  /* 0 (off).  > 0: rho ladder for controllers with state-box rows (register- and shared-memory resident kernels, streamed kernel) -- problems still unsolved
                           after ladder_iter iterations continue from their iterate with the step size of the state-box rows
                           multiplied by ladder_kappa (a second cached operator), for the remaining max_iter - ladder_iter iterations.
                           A batch-wide fixed rho leaves a few problems per 10^4 with thousands of iterations when many state
                           bounds are active (OSQP would adapt rho per problem); the second rung bounds that tail.  On the register- and shared-memory
                           resident kernels a small second rung (the usual case) runs on a CTA-cooperative kernel -- one CTA per eight stragglers --
                           without a host round trip: it is latency, not throughput. */
  int32_t ladder_kappa; /* 10 when ladder_iter > 0 and this is <= 0 */
  int32_t n_devices;    /* <= 1: the handle lives on `device`.  2..8: ONE handle drives device_ids[0..n_devices-1] from one process:
                           the per-system constants are replicated on every device at create; a batch is cut into n_devices contiguous shards
                           (sizes differ by at most one problem) that are solved concurrently with no exchange during the solve.  Host entry
                           (mpcb_solve_linear_batch, mpcb_closed_loop_linear_batch, and the NMPC host entries): one host thread and one stream pair
                           per device, results land directly in the caller's arrays.  Device entry: buffers live on device_ids[0]; the shards of the other
                           devices travel over NVLink as peer copies ordered by events (inputs out, results back into the caller's arrays), still
                           asynchronous with respect to the host. */
  int32_t device_ids[8];
  int32_t cold_init;    /* 0 (default): OSQP's cold start x = z = y = 0.  1: a cold start (no warm_u / warm_y) begins at the clipped unconstrained optimum
                           x = clip(-Pc^-1 q(p), umin, umax), z = [x; G x], with the dual guess y = -MPCB_INIT_KAPPA rho (x - v_unc) on the input-box rows
                           (zero where the box is inactive, the sign of the multiplier where it clips) and y = 0 on the general rows -- v_unc(p) is one
                           more per-system linear map of the parameters, like q(p).  Same iteration, fixed point and termination test.  Measured
                           (profiles/r02/cold_init_twin.txt, coldinit_ab_v1.jsonl): 8-20 % fewer iterations on AVERAGE, but the problems with many
                           active bounds -- the ones a batch waits for -- gain nothing: -3 % time on the on-chip kernel (H = 20, 65 536 problems),
                           +2..9 % on the shared-memory and stage-wise kernels, whose time follows the maximum over a tile.  Hence opt-in. */
} mpcb_settings;
#define MPCB_INIT_KAPPA 2.0

/* Linear (or linearised) MPC description = the data `_model_predictive_control_design` assembles for a
 * ConstrainedLinearControlDiscreteSystem (src/sub/design_mpc.jl:54-129):
 *   A, B        system.A, system.B                         (linear.jl:41-42)
 *   Q, R, S     WeightsCoefficient                          (design_mpc.jl:264-283)
 *   P           TerminalIngredient.P = are(Discrete,A,B,Q,R)  (design_mpc.jl:327); NULL -> computed here (mpcb_dare)
 *   umin/umax   first/last vertex of system.U               (linear.jl:35-38, 73-78)
 *   xmin/xmax   first/last vertex of system.X, used only when state_constraint != 0 (kw `mpc_state_constraint`
 *               present, linear.jl:62-70)
 */
typedef struct {
  int32_t nx, nu, horizon;
  const double* A; /* nx x nx */
  const double* B; /* nx x nu */
  const double* Q; /* nx x nx */
  const double* R; /* nu x nu */
  const double* S; /* nu x nu or NULL (= 0) */
  const double* P; /* nx x nx or NULL */
  const double* umin;
  const double* umax; /* nu */
  const double* xmin;
  const double* xmax; /* nx or NULL */
  int32_t state_constraint;
  int32_t terminal_mode; /* MPCB_TERMINAL_* */
} mpcb_linear_desc;

/* What the design produced (for logging, tests and the roofline arithmetic). */
typedef struct {
  int32_t nx, nu, horizon;
  int32_t nz;      /* nu*horizon decision variables (absolute inputs, stage-major) */
  int32_t mg;      /* general (non-box) constraint rows */
  int32_t nt;      /* nz + mg rows of the stacked ADMM operator */
  int32_t nt_pad;  /* padded to the MMA tile */
  int32_t kernel;  /* MPCB_KERNEL_* actually selected */
  int32_t device;
  int32_t sm_count;
  double rho;      /* rho in use */
  double lambda_min, lambda_max; /* of the condensed Hessian */
} mpcb_info;

/* One batch of independent problems sharing the designed controller.  Pointers are HOST pointers for
 * mpcb_solve_linear_batch and DEVICE pointers for mpcb_solve_linear_batch_device.  Layouts (column-major):
 *   x0    nx x batch                      the `initialization` of update_initialization! (computation_mpc.jl:17-29)
 *   xref  nx x batch, or nx x 1 if xref_broadcast   constant-over-horizon references (main_mpc.jl:105-117)
 *   uref  nu x batch, or nu x 1 if uref_broadcast
 *   u,e_u nu x horizon x batch ; x,e_x  nx x (horizon+1) x batch   == computation_results of calculate!
 *         (computation_mpc.jl:50-53), one reference-shaped matrix per problem, back to back
 *   u0    nu x batch                      first input column (what a closed loop applies)
 *   warm_u  nu x horizon x batch absolute inputs, warm_y  nt x batch duals (both or neither; NULL = cold start)
 *   y     nt x batch duals of [box rows; general rows] (for warm starting the next call)
 * Any output pointer may be NULL (skipped). */
typedef struct {
  int64_t batch;
  const double* x0;
  const double* xref;
  const double* uref;
  int32_t xref_broadcast;
  int32_t uref_broadcast;
  const double* warm_u;
  const double* warm_y;
  double* u;
  double* e_u;
  double* x;
  double* e_x;
  double* u0;
  int32_t* status;
  int32_t* iters;
  double* prim_res;
  double* dual_res;
  double* objective; /* the reference's J (design_mpc.jl:449-456), constants included */
  double* y;
  int32_t* inner_iters; /* NMPC only: total ADMM iterations over all SQP iterations (NULL = skip) */
} mpcb_batch_io;

/* CUDA-event timings of the last solve call on this handle, milliseconds (all 0 on the zero-copy small-batch path). */
typedef struct {
  float h2d_ms, solve_ms, recover_ms, d2h_ms, total_ms;
  int64_t batch;
  int64_t total_iterations; /* sum over problems of ADMM iterations */
  int32_t kernel_launches;  /* kernels of this library launched by the call */
  int32_t chunks;           /* > 1: the host entry pipelined the batch in this many chunks (download of chunk c overlaps the
                               solve of chunk c+1); the per-phase times are then 0 and only total_ms is meaningful */
} mpcb_timing;

int mpcb_version(void);
int mpcb_device_count(void);
const char* mpcb_last_error(void);
void mpcb_default_settings(mpcb_settings* s);

/* Discrete algebraic Riccati equation (replaces ControlSystems.are, design_mpc.jl:327), structure-preserving
 * doubling on the host.  All matrices column-major. */
int mpcb_dare(int32_t nx, int32_t nu, const double* A, const double* B, const double* Q, const double* R, double* P_out);
/* The same equation for MANY systems on the GPU, one warp per system (device-side design): A
 * [batch][nx*nx] and B [batch][nx*nu] column-major per system, shared weights Q (nx x nx), R (nu x nu, non-singular; HOST
 * pointers in both variants), P_out [batch][nx*nx]; status (may be NULL) = doubling steps taken (> 0) or -1 when the
 * recurrence did not settle ((A_i, B_i) not stabilisable; P_out of that system is NaN).  nu <= 3 nx. */
int mpcb_dare_batch(int32_t device, int64_t batch, int32_t nx, int32_t nu, const double* A, const double* B, const double* Q, const double* R,
                    double* P_out, int32_t* status);
int mpcb_dare_batch_device(int32_t device, int64_t batch, int32_t nx, int32_t nu, const double* dA, const double* dB, const double* Q, const double* R,
                           double* dP_out, int32_t* dstatus, void* cuda_stream);

/* Step-size selection on a sample of the workload.  OSQP adapts rho per problem and refactors (its default, which the reference gets:
 * solver_selection.jl:94-95 sets no attribute); here the KKT factor is cached once for the whole batch, so rho is chosen per CONTROLLER --
 * by default from the spectrum of the condensed Hessian (sqrt(lambda_min lambda_max)), which is right when many bounds are active and can be
 * several times too large when few are.  This call designs the controller for n_candidates step sizes rho0 * factor^j (j centred on 0,
 * rho0 = settings->rho if > 0, else the automatic value), solves the caller's SAMPLE batch (host pointers; only x0 / xref / uref are read)
 * with each on settings->device, and returns the candidate whose solve of the sample takes the least GPU time (CUDA events; cand_mean_iters then holds
 * milliseconds).  Samples of less than ~100 KB are not timed: they are scored by the mean over consecutive groups of 32 problems of the group's MAXIMUM
 * iteration count (what a tile / slot group of the kernels waits for).  Take the sample from the workload itself, a few thousand problems.  The
 * caller then passes *best_rho as settings->rho to mpcb_create_linear.  A design-time step, like OSQP's setup; solutions do not depend on it
 * beyond the termination tolerance.  cand_rho / cand_mean_iters (n_candidates each) may be NULL. */
int mpcb_tune_rho(const mpcb_linear_desc* desc, const mpcb_settings* settings, const mpcb_batch_io* host_sample, int32_t n_candidates, double factor,
                  double* best_rho, double* cand_rho, double* cand_mean_iters);

/* Design: condense, choose rho, factor K, build the stacked operators, upload to `settings->device`.
 * Replaces modeler construction + OSQP setup (linear.jl:20-103, solver_selection.jl:92-98). */
int mpcb_create_linear(const mpcb_linear_desc* desc, const mpcb_settings* settings, mpcb_handle** out);
void mpcb_destroy(mpcb_handle* h);
int mpcb_get_info(const mpcb_handle* h, mpcb_info* info);
int mpcb_get_timing(const mpcb_handle* h, mpcb_timing* t);

/* Host copies of the design matrices, for tests / inspection.  Any pointer may be NULL.
 *   Pc nz x nz, Lq nz x (2nx+nu), G mg x nz, Lb mg x (2nx+nu), T nt x nt  (all column-major) */
int mpcb_get_design(const mpcb_handle* h, double* Pc, double* Lq, double* G, double* Lb, double* T);

/* The hot path: replaces update_initialization! + calculate! (computation_mpc.jl:17-55) for `batch` problems.
 * A problem's results do not depend on the batch it is solved in: batches of up to eight problems per SM (the reference's own closed-loop use is one
 * problem per call) run on CTA-cooperative kernels and a direct recover kernel that are bit-identical to the throughput kernels of larger batches; with
 * up to 96 KB of inputs + outputs the kernels work directly on one page-locked block (no DMA operation): two launches and one synchronisation. */
int mpcb_solve_linear_batch(mpcb_handle* h, const mpcb_batch_io* host_io);
/* Same with device-resident buffers on the handle's device (device_ids[0] of a multi-device handle); `cuda_stream` is a cudaStream_t
 * (NULL = default stream).  Asynchronous: returns after enqueueing -- EXCEPT for controllers on MPCB_KERNEL_STREAMED, whose check
 * periods are driven from the host (one stream synchronisation per check; not capturable into a CUDA graph).
 * One call in flight per handle: the handle owns the work-queue counters and scratch buffers, so a call enqueued while the
 * previous one still runs -- on any stream -- is ordered behind it by an event (it never races, it may serialise). */
int mpcb_solve_linear_batch_device(mpcb_handle* h, const mpcb_batch_io* dev_io, void* cuda_stream);

/* Closed-loop batched simulation, resident on the GPU: repeats { solve from the current state (warm-started from the
 * previous step's solution and duals when warm_start != 0, as OSQP does inside one JuMP model); apply the first input;
 * advance the controller's own deviation model  x+ = x_ref + A (x - x_ref) + B (u0 - u_ref)  (linear.jl:59) } `steps`
 * times without leaving the device -- the update_initialization! / calculate! loop a user of the reference writes
 * (computation_mpc.jl:17-55; pattern of test/computation_mpc_test.jl:94-103) for a whole batch of plants.
 * HOST pointers.  x0 nx x batch; x_traj nx x (steps+1) x batch (column 1 = x0); u_traj nu x steps x batch; any output
 * may be NULL. */
typedef struct {
  int64_t batch;
  int32_t steps;
  int32_t warm_start;
  const double* x0;
  const double* xref;
  const double* uref;
  int32_t xref_broadcast;
  int32_t uref_broadcast;
  double* x_traj;
  double* u_traj;
  int32_t* iters_total;    /* per plant: ADMM iterations summed over the steps */
  int32_t* unsolved_steps; /* per plant: steps whose solve did not end with MPCB_STATUS_SOLVED */
} mpcb_closed_loop_io;
int mpcb_closed_loop_linear_batch(mpcb_handle* h, const mpcb_closed_loop_io* host_io);

/* Page-locked host memory helpers: arrays allocated here are copied to/from the device without the extra staging
 * copy that pageable memory needs (Julia: unsafe_wrap the pointer; Python: numpy.frombuffer). */
void* mpcb_alloc_pinned(size_t bytes);
void mpcb_free_pinned(void* p);

/* ------------------------------------------------------------------------------------------------------------------
 * Nonlinear path: Flux neural dynamics (fnn / resnet) + NMPC.
 *
 * Replaces, for a batch, the reference's NonLinearProgramming modelers + Ipopt
 *   src/sub/model_modeler_implementation/fnn/mpc_modeler_implementation_fnn.jl:63-189
 *   src/sub/model_modeler_implementation/resnet/mpc_modeler_implementation_resnet.jl:62-188
 *   src/sub/solver_selection.jl:100-106 (Ipopt, no attributes)
 * and the linearisation `AutomationLabsSystems.proceed_system_linearization` the LinearProgramming method applies to
 * black-box models (fnn.jl:37-46, resnet.jl:37-46; terminal cost: design_mpc.jl:312-327).
 * The network is described exactly as those modelers parse `Flux.params(system.f)` (fnn.jl:88-107):
 *   params[1] = W_in (n_neurons x (nx+nu), NO bias); then n_hidden pairs (W_j, b_j); params[end] = W_out (nx x n_neurons,
 *   NO bias).  fnn:  y_j = act(W_j y_{j-1} + b_j);  resnet:  y_j = y_{j-1} + act(W_j y_{j-1} + b_j);  polynet: see below;
 *   x+ = W_out y_end.
 * ------------------------------------------------------------------------------------------------------------------ */
#define MPCB_NN_FNN 0
#define MPCB_NN_RESNET 1
#define MPCB_NN_DENSENET 3 /* y_j = [act(W_j y_{j-1} + b_{j-1}); y_{j-1}]: W_j is n_neurons x ((j-1) n_neurons), W_out nx x ((n_hidden+1) n_neurons)
                               (densenet/mpc_modeler_implementation_densenet.jl:128-162) */
#define MPCB_NN_POLYNET 2 /* br = act(W y + b); y+ = y + br + act(W br + b)  (polynet/mpc_modeler_implementation_polynet.jl:132-149) */

#define MPCB_ACT_RELU 0
#define MPCB_ACT_TANH 1
#define MPCB_ACT_SIGMOID 2
#define MPCB_ACT_SWISH 3
#define MPCB_ACT_IDENTITY 4

typedef struct mpcb_nn mpcb_nn;
typedef struct mpcb_nmpc mpcb_nmpc;

typedef struct {
  int32_t arch;       /* MPCB_NN_* */
  int32_t activation; /* MPCB_ACT_*  (design_mpc.jl:472-496 reads it off the first hidden layer) */
  int32_t nx, nu, n_neurons, n_hidden;
  const double* W_in;     /* n_neurons x (nx+nu) column-major */
  const double* W_hidden; /* n_hidden matrices n_neurons x n_neurons (densenet: n_neurons x (j n_neurons), j = 1..n_hidden), back to back */
  const double* b_hidden; /* n_hidden vectors of n_neurons, back to back */
  const double* W_out;    /* nx x n_neurons (densenet: nx x ((n_hidden+1) n_neurons)) */
} mpcb_nn_desc;

/* A network resident on `device`. */
int mpcb_create_nn(const mpcb_nn_desc* desc, int32_t device, mpcb_nn** out);
void mpcb_destroy_nn(mpcb_nn* n);

/* Batched rollout x_{k+1} = f(x_k, u_k): x0 nx x batch, u nu x horizon x batch -> x nx x (horizon+1) x batch.
 * HOST pointers; the *_device variants take device pointers and enqueue on `cuda_stream`. */
int mpcb_nn_rollout_batch(mpcb_nn* n, int64_t batch, int32_t horizon, const double* x0, const double* u, double* x);
int mpcb_nn_rollout_batch_device(mpcb_nn* n, int64_t batch, int32_t horizon, const double* x0, const double* u, double* x, void* cuda_stream);
/* Batched forward-mode Jacobians at (x, u): x nx x batch, u nu x batch -> f nx x batch (may be NULL), A nx x nx x batch,
 * B nx x nu x batch.  Replaces proceed_system_linearization (fnn.jl:42). */
int mpcb_nn_jacobian_batch(mpcb_nn* n, int64_t batch, const double* x, const double* u, double* f, double* A, double* B);
int mpcb_nn_jacobian_batch_device(mpcb_nn* n, int64_t batch, const double* x, const double* u, double* f, double* A, double* B, void* cuda_stream);

/* NMPC controller = NL modeler + terminal ingredient + cost (design_mpc.jl:143-225).  xref/uref are the design references
 * (constant over the horizon, main_mpc.jl:105-117): the network is linearised there for P = are(...) (design_mpc.jl:312-327,
 * computed here when P is NULL) and for the ADMM step size rho. */
typedef struct {
  const mpcb_nn_desc* nn;
  int32_t horizon;
  const double* Q; /* nx x nx */
  const double* R; /* nu x nu */
  const double* S; /* nu x nu or NULL */
  const double* P; /* nx x nx or NULL */
  const double* umin;
  const double* umax;
  const double* xref; /* nx */
  const double* uref; /* nu */
  int32_t terminal_mode; /* MPCB_TERMINAL_NONE | _EQUALITY (e_x[:,end] == 0, design_mpc.jl:330-331) | _CONTRACTIVE (e_H'e_H <= 0.9 e_0'e_0, :333-340) */
  int32_t state_constraint; /* != 0: kw `mpc_state_constraint` present -> xmin <= x[:,k] <= xmax for k = 2..H+1 (fnn.jl:146-154) */
  const double* xmin; /* nx, first / last vertex of system.X; read only when state_constraint != 0 */
  const double* xmax;
} mpcb_nmpc_desc;

typedef struct {
  mpcb_settings qp;        /* inner ADMM; defaults: eps_abs = 1e-9, eps_rel = 0 (|q| is the cost gradient, so a relative dual
                              tolerance would be far looser than the SQP step tolerance), check_every = 5, sigma = 0, max_iter = 1000 per QP, rest as OSQP */
  double sqp_tol;          /* stop when the SQP step ||d||_inf <= sqp_tol          (1e-6) */
  double ls_armijo;        /* sufficient-decrease constant                           (1e-4) */
  double ls_noise;         /* round-off floor of a cost evaluation, relative to max(1,|J|): added to the Armijo bound (1e-10) */
  int32_t sqp_max_iter;    /* (20) */
  int32_t ls_max_halvings; /* (12) */
} mpcb_nmpc_settings;

void mpcb_default_nmpc_settings(mpcb_nmpc_settings* s);
int mpcb_create_nmpc(const mpcb_nmpc_desc* desc, const mpcb_nmpc_settings* settings, mpcb_nmpc** out);
void mpcb_destroy_nmpc(mpcb_nmpc* h);
/* rho in use, and host copies of the design linearisation (any pointer may be NULL): A nx x nx, B nx x nu, P nx x nx */
int mpcb_nmpc_get_design(const mpcb_nmpc* h, double* rho, double* A, double* B, double* P);
int mpcb_nmpc_get_timing(const mpcb_nmpc* h, mpcb_timing* t);

/* The hot path for the nonlinear method: update_initialization! + calculate! (computation_mpc.jl:17-55) for a batch.
 * Uses mpcb_batch_io with these meanings: warm_u = initial guess of u (NULL: the reference input clipped to the box),
 * warm_y = duals of [input box (nu*horizon) | state-box rows (nx*horizon, if state_constraint) | terminal rows (nx, with
 * the terminal equality)] per problem (may be NULL independently), y = those duals on exit,
 * iters = SQP iterations, inner_iters = total ADMM iterations, prim_res = last SQP step ||d||_inf,
 * dual_res = dual residual of the last QP, status = MPCB_STATUS_SOLVED / _SOLVED_INACCURATE / _MAX_ITER /
 * _PRIMAL_INFEASIBLE (a linearised terminal constraint that the input box does not admit within the inner iteration cap). */
int mpcb_solve_nmpc_batch(mpcb_nmpc* h, const mpcb_batch_io* host_io);
int mpcb_solve_nmpc_batch_device(mpcb_nmpc* h, const mpcb_batch_io* dev_io, void* cuda_stream);

/* The reference's LINEAR method on a black-box model (design_mpc.jl:319-327: proceed_system_linearization at the
 * reference, P = are(A, B, Q, R), then the linear modeler of linear.jl:45-93), re-designed PER PROBLEM on the device
 * (SURVEY section 8f rank 2): every problem of the batch carries its own reference (xref_i, uref_i), so the kernel
 * linearises the network there, solves that problem's Riccati equation for the terminal weight (always: the P of the
 * handle belongs to the design reference only), condenses and factors that problem's QP and solves it once -- a batch of B different
 * LTI controllers designed and evaluated in one launch.  Same handle, io meanings and constraints (input box, optional
 * state box, terminal "none"/"equality") as mpcb_solve_nmpc_batch; x / e_x are the predictions of the linearised model,
 * iters = 1, prim_res / dual_res = residuals of the QP, status = _SOLVED / _MAX_ITER / _DESIGN_FAILED. */
/* Closed-loop batched simulation of the nonlinear controller, resident on the GPU (NN plant):
 * repeats { SQP solve from the current state; apply the first input to the network itself, x+ = f(x, u0) } `steps` times
 * for a whole batch of plants.  Same io struct and meanings as mpcb_closed_loop_linear_batch; iters_total counts inner
 * ADMM iterations; with warm_start != 0 each solve starts from the previous solution and input-box duals shifted by one
 * stage (last stage repeated), the duals of state / terminal rows carried as they are. */
int mpcb_closed_loop_nmpc_batch(mpcb_nmpc* h, const mpcb_closed_loop_io* host_io);

int mpcb_solve_relinearized_batch(mpcb_nmpc* h, const mpcb_batch_io* host_io);
int mpcb_solve_relinearized_batch_device(mpcb_nmpc* h, const mpcb_batch_io* dev_io, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MPCB200_H */

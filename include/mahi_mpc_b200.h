/* mahi-mpc-b200 -- C ABI of the B200-native multiple-shooting MPC hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, caller-owned buffers, no exceptions and no
 * C++/torch types.  Everything the reference does between "ModelControl::calc_u hands vectors to
 * casadi::nlpsol" and "gets res["x"] back" (src/Mahi/Mpc/ModelControl.cpp:116-172) happens behind
 * mmpc_solve(); the CasADi-generated nlp_* callbacks (emitted by ModelGenerator.cpp:238) are behind
 * mmpc_nlp_eval() / mmpc_eval_nodes().  The C++ facade in include/Mahi/Mpc/ *.hpp (same class and
 * method names as the reference) is a thin layer over these calls; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns MMPC_OK (0) or a negative mmpc_status; mmpc_last_error() gives the
 *     thread-local message of the last failure on the calling thread;
 *   - all arithmetic is fp64; vectors are dense `double`, laid out exactly as the reference lays
 *     them out (V, p, g: ModelGenerator.cpp:61-112,129-187,206-225), instance-major for batches;
 *   - a pointer argument is a HOST pointer if `mem == MMPC_MEM_HOST`, a DEVICE pointer (on the
 *     solver's device) if `mem == MMPC_MEM_DEVICE`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream); calls are
 *     synchronous w.r.t. the host for MMPC_MEM_HOST and stream-ordered for MMPC_MEM_DEVICE;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     MMPC_ERR_NO_DEVICE.
 */
#ifndef MAHI_MPC_B200_H
#define MAHI_MPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MMPC_ABI_VERSION 1

typedef enum mmpc_status {
    MMPC_OK = 0,
    MMPC_ERR_INVALID = -1,    /* bad argument / size mismatch */
    MMPC_ERR_NO_DEVICE = -2,  /* no usable CUDA device (there is no CPU path) */
    MMPC_ERR_CUDA = -3,       /* CUDA runtime error, see mmpc_last_error() */
    MMPC_ERR_IO = -4,         /* model artefact missing / unreadable */
    MMPC_ERR_UNKNOWN_MODEL = -5,
    MMPC_ERR_ALLOC = -6,
    MMPC_ERR_UNSUPPORTED = -7
} mmpc_status;

typedef enum mmpc_mem { MMPC_MEM_HOST = 0, MMPC_MEM_DEVICE = 1 } mmpc_mem;

/* integrator of the one-step map F: EULER is what the library's NLP uses
 * (ModelGenerator.cpp:33-34); RK4 (x `substeps`) is the MATLAB prototypes' (workingMPC_8_30_21.m:46-59). */
typedef enum mmpc_integrator { MMPC_EULER = 0, MMPC_RK4 = 1 } mmpc_integrator;
/* Lagrangian Hessian of the NLP under RK4 shooting: EXACT differentiates twice through the integrator (what CasADi's AD hands
 * IPOPT in the prototypes, workingMPC_8_30_21.m:46-59,109-112; up to 4 sub-steps); GAUSS_NEWTON keeps the objective's
 * first-order part only (same KKT point, linear convergence).  Euler / linearised shooting is always exact. */
typedef enum mmpc_hessian { MMPC_HESSIAN_EXACT = 0, MMPC_HESSIAN_GAUSS_NEWTON = 1 } mmpc_hessian;
/* Where the tracking error of the objective sits.  SHOOTING: on the one-step map, e_k = F(x_k,u_k) - r_k -- the library's NLP
 * (ModelGenerator.cpp:207-214).  NODE: on the decision variable, e_k = x_{k+1} - r_k -- the MATLAB prototypes' NLP
 * (workingMPC_8_30_21.m:90-106: the error of x(:,k) against the reference at its own time; there Q <- 1000 Q, no du term and
 * Rm <- R; x_0 is pinned, so its term is a constant and is left out).  Same minimisers on the feasible set, a different
 * Lagrangian (constant 2Q blocks, no sum 2Q e_i d2F_i term).  NODE runs on T4 / T1. */
typedef enum mmpc_cost { MMPC_COST_SHOOTING = 0, MMPC_COST_NODE = 1 } mmpc_cost;

/* per-instance solve status (the reference never inspects IPOPT's, ModelControl.cpp:159-160) */
typedef enum mmpc_solve_status {
    MMPC_SOLVE_CONVERGED = 0,
    MMPC_SOLVE_MAX_ITER = 1,
    MMPC_SOLVE_NUMERICAL = 2 /* NaN/Inf or regularisation exhausted */
} mmpc_solve_status;

typedef struct mmpc_model* mmpc_model_t;
typedef struct mmpc_solver* mmpc_solver_t;

/* Field-for-field mirror of mahi::mpc::ModelParameters (include/Mahi/Mpc/ModelParameters.hpp:11-28);
 * times in microseconds as in the JSON artefact (ModelParameters.cpp:37-50). Bound pointers may be
 * NULL (= the reference's "no bound" sentinel +-10e30, ModelParameters.cpp:21-24). */
typedef struct mmpc_model_params {
    const char* name;
    long long timespan_us;
    long long step_size_us;
    int num_x;
    int num_u;
    int num_shooting_nodes;
    const double* x_min;
    const double* u_min;
    const double* x_max;
    const double* u_max;
    const char* dll_filepath; /* "<name>.so" (generated CUDA model library) or "builtin:<kernel set>" */
    int is_linear;
} mmpc_model_params;

/* Solver options. Defaults = the reference's hard-coded IPOPT options (ModelControl.cpp:54-59). */
typedef struct mmpc_options {
    double tol;            /* 1e-5  scaled KKT error (IPOPT "tol") */
    int max_iter;          /* 200 */
    int integrator;        /* MMPC_EULER */
    int substeps;          /* 1 */
    double mu_init;        /* 0.1   barrier start when finite bounds exist (IPOPT mu_init) */
    int warm_start;        /* 1: start from the previous solution of the same instance slot
                                 (ModelControl.cpp:161: args["x0"] = result), 0: always cold (zeros) */
    int warm_start_duals;  /* 0: reference behaviour (primal only); 1: also reuse lam_g */
    int kernel_variant;    /* 0 = auto: T4 for nx+2nu <= 8 without finite bounds, T3 with them, else T2; T1 for what
                              those do not implement.  1 / 2 / 3 / 4 request T1 / T2 / T3 / T4; 3 + 256*k and 4 + 256*k force a
                              tile configuration (mmpc_solve_t3.cuh, mmpc_solve_t4.cuh) -- experiments and tests only.
                              Changing it forgets the warm start. */
    int hessian;           /* MMPC_HESSIAN_EXACT (default) / MMPC_HESSIAN_GAUSS_NEWTON: RK4 shooting only */
    int cost;              /* MMPC_COST_SHOOTING (default, the library's objective) / MMPC_COST_NODE (the prototypes') */
    int reserved[6];
} mmpc_options;

/* Per-instance statistics returned by mmpc_solve (40 bytes). */
typedef struct mmpc_stats {
    double objective;  /* f at the returned point */
    double kkt_error;  /* scaled optimality error at the returned point */
    double mu;         /* final barrier parameter (0 if no finite bounds) */
    int iterations;
    int status;        /* mmpc_solve_status */
    int ls_steps;      /* total backtracking steps */
    int reg_steps;     /* total Hessian regularisation retries */
} mmpc_stats;

const char* mmpc_last_error(void);
int mmpc_abi_version(void);
int mmpc_device_count(int* count);
int mmpc_options_default(mmpc_options* opts);

/* ---- model artefact -------------------------------------------------------------------------
 * replaces ModelControl::load_model's json + nlpsol(dll) + casadi::external loading
 * (src/Mahi/Mpc/ModelControl.cpp:21-26,61-62,70-72). */
int mmpc_model_create(const mmpc_model_params* params, mmpc_model_t* out);
/* reads "<model_name>.json" in the reference's schema (ModelParameters.cpp:37-72) */
int mmpc_model_load(const char* model_name, mmpc_model_t* out);
int mmpc_model_free(mmpc_model_t model);
/* nx, nu, N, NV = nx(N+1)+nu N, M = N nx, np = length of p (ModelGenerator.cpp:61,129-143) */
int mmpc_model_dims(mmpc_model_t model, int* nx, int* nu, int* N, int* NV, int* M, int* np);
int mmpc_model_step_size(mmpc_model_t model, double* seconds);
int mmpc_model_is_linear(mmpc_model_t model, int* is_linear);
/* straight-line op counts of the generated device code (arith, transcendental) for f, f+A+B, f+A+B+H */
int mmpc_model_op_counts(mmpc_model_t model, int counts[6]);

/* ---- <name>_get_A / _get_B / _get_x_dot_init (ModelGenerator.cpp:51-53; ModelControl.cpp:127-133)
 * n points; x[n*nx], u[n*nu] -> xdot[n*nx], A[n*nx*nx], B[n*nx*nu], A and B COLUMN-major per point
 * (the order ModelControl flattens them into p). Any output may be NULL. */
int mmpc_linearise(mmpc_model_t model, int device, int n, const double* x, const double* u, double* xdot, double* A,
                   double* B, mmpc_mem mem, void* stream);

/* ---- shooting-node evaluation (BASELINE config 5; SURVEY 8 a-2,a-5,a-7,a-8) --------------------
 * n independent nodes, STRUCTURE-OF-ARRAYS: component i of node j at [i*n + j].
 *   in : x[nx*n], u[nu*n], x_next[nx*n], w[nx*n] (may be NULL)
 *   out: g[nx*n]      = F(x,u) - x_next
 *        Fx[nx*nx*n]  = dF/dx  (entry (r,c) at [(r*nx+c)*n + j])
 *        Fu[nx*nu*n]  = dF/du
 *        H[s(s+1)/2*n] = sum_i w_i d2F_i/dz2, packed upper triangle row-major, only if w != NULL
 * All pointers in `mem` space. */
int mmpc_eval_nodes(mmpc_model_t model, int device, const mmpc_options* opts, long long n, const double* x, const double* u,
                    const double* x_next, const double* w, double* g, double* Fx, double* Fu, double* H, mmpc_mem mem,
                    void* stream);

/* ---- whole-NLP callbacks, batched: nlp_f, nlp_g, nlp_grad_f, nlp_jac_g, nlp_hess_l -------------
 * (the functions CasADi generates from ModelGenerator.cpp:225-232,238).  Instance-major AoS inputs in the
 * reference's layouts: V[B*NV], p[B*np], lam_g[B*M]; outputs (any may be NULL):
 *   f[B], g[B*M], grad_f[B*NV],
 *   jac blocks  Fx[B*N*nx*nx], Fu[B*N*nx*nu] (row-major per node; d g_k/d x_{k+1} = -I is implicit),
 *   hess blocks Hzz[B*N*s*s] (row-major full stage blocks of lam_f*f + lam_g'g w.r.t. z_k=(x_k,u_k);
 *   the (u_k,u_{k+1}) coupling block is -2*lam_f*diag(R), the x_N block is zero). */
int mmpc_nlp_eval(mmpc_model_t model, int device, const mmpc_options* opts, int batch, const double* V, const double* p,
                  const double* lam_g, double lam_f, double* f, double* g, double* grad_f, double* Fx, double* Fu,
                  double* Hzz, mmpc_mem mem, void* stream);

/* ---- solver ---------------------------------------------------------------------------------- */
int mmpc_solver_create(mmpc_model_t model, int device, int batch, const mmpc_options* opts, mmpc_solver_t* out);
int mmpc_solver_free(mmpc_solver_t solver);
int mmpc_solver_batch(mmpc_solver_t solver, int* batch);
int mmpc_set_options(mmpc_solver_t solver, const mmpc_options* opts);
/* Q[nx], R[nu], Rm[nu] shared by all instances (per_instance = 0) or Q[B*nx] ... (per_instance = 1);
 * HOST pointers. Mirrors ModelControl ctor / update_weights (ModelControl.cpp:7-14,199-203): a NULL
 * pointer leaves that weight unchanged. */
int mmpc_set_weights(mmpc_solver_t solver, const double* Q, const double* R, const double* Rm, int per_instance);
/* x_min/x_max[nx], u_min/u_max[nu], HOST pointers, NULL = keep. |b| >= 1e19 means "no bound"
 * (ModelControl::update_control_limits, ModelControl.cpp:205-209; IPOPT nlp_*_bound_inf). */
int mmpc_set_bounds(mmpc_solver_t solver, const double* x_min, const double* x_max, const double* u_min, const double* u_max);
/* forget the stored warm start (next solve starts from zeros, ModelControl.cpp:28-50,68) */
int mmpc_reset_warm_start(mmpc_solver_t solver);

/* One batched solve = `batch` independent ModelControl::calc_u calls (ModelControl.cpp:116-163).
 *   state[B*nx]   current state, pinned as x_0           (ModelControl.cpp:144-145)
 *   control[B*nu] previous control = u_init in p         (ModelControl.cpp:136)
 *   traj[B*N*nx]  reference r_0..r_{N-1}                 (ModelControl.cpp:116,120)
 *   V_out[B*NV]   res["x"] in the reference layout       (ModelControl.cpp:160)
 *   lam_out[B*M]  constraint multipliers (may be NULL)
 *   stats[B]      (may be NULL)
 * For is_linear models A, B, xdot_init are computed on the device at (state, control)
 * (ModelControl.cpp:125-135).
 * MMPC_MEM_HOST transfers: (1) small batches (staging <= 256 KB, e.g. one calc_u) are copied into the handle's page-locked
 * staging buffers, which the tile kernels read and write IN PLACE over PCIe (zero-copy: one launch + one synchronisation, no
 * copy-engine transfers; MMPC_ZERO_COPY=0 restores the packed H2D / D2H copies); (2) page-locked caller buffers that the device
 * can address (cudaHostAlloc / cudaHostRegister, torch pin_memory()) are used in place the same way, in up to 8 sub-batches on
 * their own streams (MMPC_ZC_BIG=0: copy engines); (3) pageable buffers go through cudaMemcpyAsync.  Same bits on every path. */
int mmpc_solve(mmpc_solver_t solver, const double* state, const double* control, const double* traj, double* V_out,
               double* lam_out, mmpc_stats* stats, mmpc_mem mem, void* stream);

/* Closed-loop simulation of `batch` independent control loops, device resident (the loops of
 * examples/model_generate_example.cpp:156-230 and examples/model_control_example.cpp:54-93, batched): every cycle
 * re-plans with mmpc_solve against the window [c, c+N) of the long reference, applies the first control of the plan
 * and advances the plant one step_size with plant_integrator = MMPC_RK4 (model_generate_example.cpp:207-213) or
 * MMPC_EULER (model_control_example.cpp:81-86).  No host synchronisation inside the loop.
 *   state[B*nx], control[B*nu]      in: initial state / previous control; out: final ones
 *   traj_long[B*(cycles+N-1)*nx]    reference samples r(t0 + m*step_size), m = 0..cycles+N-2
 *   X_log[B*cycles*nx], U_log[B*cycles*nu]   state before each step and applied control (either may be NULL)
 *   iters_log[B*cycles]             Newton iterations of every solve (may be NULL)
 * The handle's warm start carries from cycle to cycle exactly as in ModelControl (ModelControl.cpp:161). */
int mmpc_closed_loop(mmpc_solver_t solver, int cycles, int plant_integrator, double* state, double* control, const double* traj_long,
                     double* X_log, double* U_log, int* iters_log, mmpc_mem mem, void* stream);

/* Time of the device work of the last mmpc_solve on this handle in milliseconds (CUDA events on the
 * launching stream): [0] H2D staging, [1] kernels, [2] D2H staging.  Host-buffer calls that are pipelined over sub-batches
 * (batches >= 4096) report the whole H2D + solve + D2H span in [1] and zeros in [0], [2]; zero-copy small-batch calls report
 * launch .. synchronised on the HOST clock in [1] (no events are recorded on that path) and zeros in [0], [2].
 * Tracing: every entry point opens an NVTX range ("mmpc_solve", ...); with MMPC_STATS_JSONL=<path> in the environment every
 * host-buffer mmpc_solve that returns statistics appends one JSON line (batch, N, kernel variant, times, iteration and
 * status counts) to <path>. */
int mmpc_last_timing(mmpc_solver_t solver, double ms[3]);
/* number of kernels the last mmpc_solve launched */
int mmpc_last_launch_count(mmpc_solver_t solver, int* launches);

/* ---- several GPUs in one process -------------------------------------------------------------
 * Independent MPC instances shard trivially (BASELINE north_star): a group owns one solver handle per
 * device, each holding a contiguous shard of `batch_total` instances (the first batch_total % n shards are one longer).
 * There is no collective on the solve path; the only cross-device traffic is the gather of plans and statistics.
 * `devices` may name a device more than once (several shards on one GPU).  The reference has no counterpart: its
 * ModelControl solves one instance on one CPU thread (ModelControl.cpp:116-172). */
typedef struct mmpc_group* mmpc_group_t;
int mmpc_group_create(mmpc_model_t model, const int* devices, int n_devices, int batch_total, const mmpc_options* opts, mmpc_group_t* out);
int mmpc_group_free(mmpc_group_t group);
/* shard i: its device, first instance and instance count; its handle (for per-shard calls such as mmpc_set_bounds) */
int mmpc_group_shard(mmpc_group_t group, int i, int* device, int* offset, int* count, mmpc_solver_t* solver);
int mmpc_group_size(mmpc_group_t group, int* n_shards, int* batch_total);
/* weights shared by all instances (as mmpc_set_weights with per_instance = 0) / control and state limits, on every shard */
int mmpc_group_set_weights(mmpc_group_t group, const double* Q, const double* R, const double* Rm);
int mmpc_group_set_bounds(mmpc_group_t group, const double* x_min, const double* x_max, const double* u_min, const double* u_max);
/* mmpc_solve over all shards concurrently, HOST buffers of the whole batch (layouts of mmpc_solve): every device copies its
 * shard in, solves, and copies its plans / multipliers / statistics straight into the caller's arrays -- the gather to the host. */
int mmpc_group_solve(mmpc_group_t group, const double* state, const double* control, const double* traj, double* V_out,
                     double* lam_out, mmpc_stats* stats);
/* Gather of the last mmpc_group_solve's plans and statistics to ONE device (peer-to-peer copies over NVLink where the
 * devices allow it): V_root[batch_total*NV] and stats_root[batch_total] (may be NULL) are DEVICE pointers on `root_device`.
 * ms_out (may be NULL): duration of the copies on the root device, CUDA events. */
int mmpc_group_gather(mmpc_group_t group, int root_device, double* V_root, mmpc_stats* stats_root, double* ms_out);

/* ---- measurement utility ----------------------------------------------------------------------
 * Measured fp64 FMA throughput of `device` in TFLOP/s (independent DFMA chains, every SM busy,
 * CUDA-event timed) -- the denominator of the solve kernels' FP64 roofline fraction
 * (MEASURED_PEAKS.json has no fp64 entry). */
int mmpc_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MAHI_MPC_B200_H */

// mahi-mpc-b200 :: internal interface between the core library (csrc/abi) and a model's kernel set
// (one translation unit per model, produced by the ModelGenerator: generated model functor +
// mmpc_kernels.cuh + MMPC_DEFINE_MODEL).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

#define MMPC_MAXNX 16
#define MMPC_MAXNU 8
#define MMPC_VTABLE_VERSION 9

struct mmpc_stats;  // public struct (include/mahi_mpc_b200.h)

// Problem constants shared by every instance of a batch (passed to kernels by value).
struct mmpc_problem {
    int N;           // shooting nodes
    int is_linear;   // F_lin instead of F (ModelGenerator.cpp:47-48,58)
    int integrator;  // MMPC_EULER / MMPC_RK4
    int substeps;
    double h;        // step_size in seconds
    int bounded;     // any finite bound present
    int gauss_newton;  // RK4 shooting only: drop the second derivatives of the integrator from the Lagrangian Hessian
    int node_cost;     // tracking error on the decision variable x_{k+1} instead of F(x_k, u_k) (MMPC_COST_NODE)
    double x_min[MMPC_MAXNX], x_max[MMPC_MAXNX], u_min[MMPC_MAXNU], u_max[MMPC_MAXNU];
};

// Device workspace of one solver handle. Iterates are structure-of-arrays with the instance index
// fastest: element (row r, instance b) of an array with `rows` rows lives at [r * Bp + b].
struct mmpc_workspace {
    int B, Bp;
    // iterate (persists between solves = warm start)
    double *X;     // (N+1)*nx rows : x_k[i] at row k*nx+i
    double *U;     // N*nu rows
    double *LAM;   // N*nx rows     : multiplier of g_k
    double *ZLU, *ZUU, *ZLX, *ZUX;  // bound multipliers for u (N*nu rows) and x_1..x_N (N*nx rows)
    // Newton step
    double *DX, *DU, *LAMN, *DZLU, *DZUU, *DZLX, *DZUX;
    // Riccati products per stage
    double *KX;    // N*nu*nx rows : feedback on dx
    double *KV;    // N*nu*nu rows : feedback on du_{k-1}
    double *K0;    // N*nu rows    : feed-forward
    double *PXX;   // N*nx*nx rows : value-function Hessian blocks of stage k+1 (stored at k)
    double *PXV;   // N*nx*nu rows
    double *PX;    // N*nx rows
    // linearisation data for is_linear models (nx*nx + nx*nu + nx rows)
    double *LIN;
    // weights, SoA: Q nx rows, R nu rows, Rm nu rows
    double *Qw, *Rw, *Rmw;
    int have_warm;  // X/U (and LAM) hold a previous solution
    // ---- variant T2 (sub-warp cooperative kernel): instance-major (AoS) arrays ----
    double *V2;     // B * NV   iterate in the reference's V layout (persists = warm start)
    double *LAM2;   // B * M
    double *ZL2, *ZU2;  // B * NV  bound multipliers in V layout
    double *DV2;    // B * NV   Newton step in V layout
    double *LAMN2;  // B * M    multipliers of the QP
    double *DZL2, *DZU2;  // B * NV
    double *FG2;    // B * N * (nx*nx + nx*nu + nx + nx+2nu)   Fx, Fu, c, objective-gradient pieces of every node
    double *KG2;    // B * N * (nu*(nx+nu) + nu)               feedback gains
    double *PG2;    // B * N * (nx*(nx+nu+1))                  rows of the value function needed for lam
    double *LIN2;   // B * (nx*nx + nx*nu + nx)
    int variant;    // kernel variant the persistent state belongs to (1 = T1, 2 = T2, 3 = T3, 4 = T4)
};

// Tile width (instances per CTA) of kernel variant T4 for a handle of `batch` instances (mmpc_solve_t4.cuh): 4 (8 node slots per
// instance, 64 threads, two CTAs per SM at most) while the batch is at most 8 * 148 instances, 32 beyond.  `forced` =
// kernel_variant >> 8 (1: 8, 2: 28, 3: 32, 4: 4).  The ABI layer cuts host-buffer sub-batches on multiples of it.
static inline int mmpc_t4_tile_width(int batch, int forced) {
    if (forced == 1) return 8;
    if (forced == 2) return 28;
    if (forced == 4) return 4;
    if (forced >= 3) return 32;
    return batch <= 148 * 8 ? 4 : 32;
}

struct mmpc_solve_io {
    const double *state, *control, *traj;  // AoS device pointers (B*nx, B*nu, B*N*nx)
    double *V_out, *lam_out;               // AoS device pointers (B*NV, B*M); lam_out may be null
    mmpc_stats* stats;                     // device pointer (B) or null
    double tol, mu_init;
    int max_iter, warm_start, warm_start_duals, variant;
    int b0, b1;  // instance range [b0, b1) of this launch (b0 a multiple of 32): lets the ABI layer pipeline sub-batches
    // Zero-copy inputs (T3 / T4): when traj_stage != NULL, state / control / traj point into page-locked HOST memory that the
    // device reads in place over PCIe (no copy-engine transfers before the kernel): each CTA first copies the rows of its tile's
    // instances into these device buffers (same indexing as the inputs) and reads them from there -- the reference trajectories
    // are read in every sweep.  NULL: the inputs are device memory and are read in place.
    double *state_stage, *control_stage, *traj_stage;
    // Zero-copy outputs (T3 / T4): V_out / lam_out / stats point into page-locked HOST memory; the kernel writes them in place
    // with a transposing epilogue (whole 256-byte runs per warp instead of 8-byte PCIe writes).  0: device memory, plain epilogue.
    int host_out;
};

struct mmpc_model_vtable {
    int version;
    const char* name;
    int nx, nu;
    int op_counts[6];
    // launches return cudaError_t of the launch; *launches += number of kernels launched
    cudaError_t (*eval_nodes)(const mmpc_problem& P, long long n, const double* x, const double* u, const double* xn,
                              const double* w, double* g, double* Fx, double* Fu, double* H, cudaStream_t st, int* launches);
    cudaError_t (*linearise)(int n, const double* x, const double* u, double* xdot, double* A, double* B, cudaStream_t st,
                             int* launches);
    cudaError_t (*nlp_eval)(const mmpc_problem& P, int B, const double* V, const double* p, const double* lam, double lam_f,
                            double* f, double* g, double* grad_f, double* Fx, double* Fu, double* Hzz, cudaStream_t st,
                            int* launches);
    cudaError_t (*solve)(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st,
                         int* launches);
    // plant of the closed-loop drivers: apply the first control of each plan V for one step of length h
    // (plant_integrator MMPC_EULER: examples/model_control_example.cpp:81-86, MMPC_RK4: model_generate_example.cpp:207-213);
    // logs the state before the step and the applied control at log slot `cycle` of `cycles`.
    cudaError_t (*plant_step)(int B, int N, double h, int plant_integrator, const double* V, double* state, double* control,
                              double* X_log, double* U_log, int cycle, int cycles, cudaStream_t st, int* launches);
};

extern "C" typedef const mmpc_model_vtable* (*mmpc_vtable_getter)(void);

// mahi-mpc-b200 :: sm_100a kernels of the multiple-shooting MPC hot path, templated on a generated
// model functor `M` (struct mmpc_model_<name>: NX, NU, f, fAB, fABH -- see csrc/modelgen/codegen.cpp).
//
// What replaces what (reference file:line):
//   eval_nodes_kernel   the per-node body of the CasADi-generated nlp_g / nlp_jac_g / nlp_hess_l
//                       (node loop ModelGenerator.cpp:191-222; F = x + h f, :33-34)           SURVEY 8 a-2,a-5,a-7,a-8
//   linearise_kernel    <name>_get_A/_get_B/_get_x_dot_init (ModelGenerator.cpp:45-46,51-53)      SURVEY 8 a-9
//   nlp_eval_kernel     nlp_f, nlp_g, nlp_grad_f, nlp_jac_g, nlp_hess_l in the reference layouts    SURVEY 8 a-5..a-8
//   solve_kernel        m_solver(m_solver_args) = IPOPT + MUMPS (ModelControl.cpp:159): here an exact-Hessian
//                       Newton / primal-dual interior-point iteration whose KKT system is solved by a
//                       stage-wise Riccati recursion on the augmented state (x_k, u_{k-1})           SURVEY 8 a-10
//
// All arithmetic fp64 on the DFMA pipe; no tensor cores (tcgen05 has no fp64 type).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <math.h>

#include "../../../include/mahi_mpc_b200.h"
#include "mmpc_vtable.h"

namespace mmpc {

#define MMPC_UNROLL _Pragma("unroll")

template <class M>
struct Dim {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, NH = S * (S + 1) / 2;
};

// packed upper-triangular index (row-major), requires i <= j
__host__ __device__ constexpr int tri(int i, int j, int S) { return i * S - (i * (i - 1)) / 2 + (j - i); }
__host__ __device__ constexpr int trisym(int i, int j, int S) { return i <= j ? tri(i, j, S) : tri(j, i, S); }

// ---------------------------------------------------------------------------------------------
// one-step map F and its derivatives (explicit Euler: F = x + h f, ModelGenerator.cpp:33-34)
//   WANT 0: F          1: F, Fx, Fu        2: F, Fx, Fu, Hs
// Hs = sum_i w_i d2F_i/dz2 (packed upper triangle) with the NLP's weights
//   w_i = wq_i * e_i + lam_i,  e_i = F_i - r_i   (wq = 2*lam_f*Q for nlp_hess_l; 0 for a plain contraction)
// Because e depends on f itself, the generated functor takes the weights in affine form
//   w_i = wa_i f_i + wb_i,  wa_i = wq_i h^2,  wb_i = h (wq_i (x_i - r_i) + lam_i).
// ---------------------------------------------------------------------------------------------
template <class M, int WANT>
__device__ __forceinline__ void step_euler(double h, const double* x, const double* u, const double* wq, const double* r,
                                           const double* lam, double* F, double* Fx, double* Fu, double* Hs) {
    constexpr int NX = M::NX, NU = M::NU;
    double xd[NX];
    if (WANT == 0) {
        M::f(x, u, xd);
    } else if (WANT == 1) {
        M::fAB(x, u, xd, Fx, Fu);
    } else {
        double wa[NX], wb[NX];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            wa[i] = wq[i] * h * h;
            wb[i] = h * (wq[i] * (x[i] - r[i]) + lam[i]);
        }
        M::fABH(x, u, wa, wb, xd, Fx, Fu, Hs);
    }
    MMPC_UNROLL for (int i = 0; i < NX; ++i) F[i] = fma(h, xd[i], x[i]);
    if (WANT >= 1) {
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            MMPC_UNROLL for (int j = 0; j < NX; ++j) Fx[i * NX + j] = (i == j ? 1.0 : 0.0) + h * Fx[i * NX + j];
            MMPC_UNROLL for (int j = 0; j < NU; ++j) Fu[i * NU + j] = h * Fu[i * NU + j];
        }
    }
}

// linearised one-step map F_lin = x + h (A (x - x_init) + B (u - u_init) + xdot_init) (ModelGenerator.cpp:47-48)
// A, B row-major here.
template <class M, int WANT>
__device__ __forceinline__ void step_linear(double h, const double* x, const double* u, const double* A, const double* B,
                                            const double* xdot0, const double* x_init, const double* u_init, double* F,
                                            double* Fx, double* Fu, double* Hs) {
    constexpr int NX = M::NX, NU = M::NU;
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        double acc = xdot0[i];
        MMPC_UNROLL for (int j = 0; j < NX; ++j) acc = fma(A[i * NX + j], x[j] - x_init[j], acc);
        MMPC_UNROLL for (int j = 0; j < NU; ++j) acc = fma(B[i * NU + j], u[j] - u_init[j], acc);
        F[i] = fma(h, acc, x[i]);
    }
    if (WANT >= 1) {
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            MMPC_UNROLL for (int j = 0; j < NX; ++j) Fx[i * NX + j] = (i == j ? 1.0 : 0.0) + h * A[i * NX + j];
            MMPC_UNROLL for (int j = 0; j < NU; ++j) Fu[i * NU + j] = h * B[i * NU + j];
        }
    }
    if (WANT >= 2) {
        MMPC_UNROLL for (int i = 0; i < Dim<M>::NH; ++i) Hs[i] = 0.0;
    }
}

// RK4 x `substeps` shooting interval with exact forward sensitivities (the MATLAB prototypes' integrator:
// workingMPC_8_30_21.m:46-59, 4 RK4 sub-steps per interval; plant of examples/model_generate_example.cpp:207-213).
//   WANT 0: F           1: F, Fx = dF/dx, Fu = dF/du
// Second derivatives through the integrator are not formed: under RK4 the solvers use the Gauss-Newton part of the
// Lagrangian Hessian, which reaches the same KKT point.
template <class M, int WANT>
__device__ void step_rk4(double h, int substeps, const double* x, const double* u, double* F, double* Fx, double* Fu) {
    constexpr int NX = M::NX, NU = M::NU;
    const double hs = h / (double)substeps;
    double xs[NX], Sx[NX * NX], Su[NX * NU];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) xs[i] = x[i];
    if (WANT >= 1) {
        MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) Sx[i] = (i / NX == i % NX) ? 1.0 : 0.0;
        MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Su[i] = 0.0;
    }
    for (int m = 0; m < substeps; ++m) {
        double xa[NX], acc[NX], accx[NX * NX], accu[NX * NU], Tx[NX * NX], Tu[NX * NU];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) { xa[i] = xs[i]; acc[i] = 0.0; }
        if (WANT >= 1) {
            MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) { Tx[i] = Sx[i]; accx[i] = 0.0; }
            MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) { Tu[i] = Su[i]; accu[i] = 0.0; }
        }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            const double wgt = (st == 0 || st == 3) ? 1.0 / 6.0 : 1.0 / 3.0;   // RK4 weights
            const double adv = (st == 2) ? 1.0 : 0.5;                          // next stage point = xs + adv*hs*k
            double k[NX], kx[NX * NX], ku[NX * NU];
            if (WANT == 0) {
                M::f(xa, u, k);
            } else {
                double A[NX * NX], B[NX * NU];
                M::fAB(xa, u, k, A, B);
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    MMPC_UNROLL for (int j = 0; j < NX; ++j) {
                        double t = 0.0;
                        MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(A[i * NX + l], Tx[l * NX + j], t);
                        kx[i * NX + j] = t;
                    }
                    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                        double t = B[i * NU + j];
                        MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(A[i * NX + l], Tu[l * NU + j], t);
                        ku[i * NU + j] = t;
                    }
                }
            }
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                acc[i] = fma(wgt, k[i], acc[i]);
                xa[i] = fma(adv * hs, k[i], xs[i]);
            }
            if (WANT >= 1) {
                MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) { accx[i] = fma(wgt, kx[i], accx[i]); Tx[i] = fma(adv * hs, kx[i], Sx[i]); }
                MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) { accu[i] = fma(wgt, ku[i], accu[i]); Tu[i] = fma(adv * hs, ku[i], Su[i]); }
            }
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) xs[i] = fma(hs, acc[i], xs[i]);
        if (WANT >= 1) {
            MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) Sx[i] = fma(hs, accx[i], Sx[i]);
            MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Su[i] = fma(hs, accu[i], Su[i]);
        }
    }
    MMPC_UNROLL for (int i = 0; i < NX; ++i) F[i] = xs[i];
    if (WANT >= 1) {
        MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) Fx[i] = Sx[i];
        MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Fu[i] = Su[i];
    }
}

// RK4 x `substeps` with EXACT second derivatives: Hs = sum_i wbar_i d2F_i/dz2 (packed upper triangle over z = (x, u)), where
// wbar_i = wq_i (F_i - r_i) + lam_i are the NLP's weights of nlp_hess_l (wq == nullptr: wbar = lam, a plain contraction).
// The interval map is a composition of 4 * substeps evaluations of f joined by linear operations, so
//     d2(wbar' F)/dz2 = sum over evaluations e of  W_e' [ sum_i kbar_{e,i} d2f_i/d(x,u)2 at the stage point ] W_e,
// with W_e = d(stage point, u)/dz (forward sensitivities) and kbar_e the adjoint of the evaluation's output (reverse sweep
// through the Butcher tableau: kbar_4 = hs b_4 v, xbar_4 = A_4' kbar_4, kbar_3 = hs b_3 v + hs xbar_4, ...,
// v_prev = v + sum_j xbar_j).  Pass 1 (forward) stores the stage points and their sensitivities in local memory; pass 2
// (reverse) calls the model's fABH at every stage point with weights kbar (it returns A for the adjoint and the contraction).
// This is the Hessian CasADi's AD gives IPOPT for the prototypes' NLP (workingMPC_8_30_21.m:46-59,109-112: RK4 x 4 inside
// opti.subject_to(x(:,k+1)==F(x(:,k),u(:,k)))).  substeps <= MMPC_RK4_HESS_MAX_SUBSTEPS.
#define MMPC_RK4_HESS_MAX_SUBSTEPS 4
template <class M>
__device__ void step_rk4_hess(double h, int substeps, const double* x, const double* u, const double* wq, const double* r, const double* lam,
                              double* F, double* Fx, double* Fu, double* Hs) {
    constexpr int NX = M::NX, NU = M::NU, S = NX + NU, NH = S * (S + 1) / 2, MAXE = 4 * MMPC_RK4_HESS_MAX_SUBSTEPS;
    const double hs = h / (double)substeps;
    double xst[MAXE][NX], Tst[MAXE][NX * S];  // stage points and d(stage point)/dz, z = (x, u)
    double xs[NX], Sz[NX * S];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) xs[i] = x[i];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) MMPC_UNROLL for (int j = 0; j < S; ++j) Sz[i * S + j] = (i == j) ? 1.0 : 0.0;
    for (int m = 0; m < substeps; ++m) {
        double xa[NX], acc[NX], accz[NX * S], Tz[NX * S];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) { xa[i] = xs[i]; acc[i] = 0.0; }
        MMPC_UNROLL for (int i = 0; i < NX * S; ++i) { Tz[i] = Sz[i]; accz[i] = 0.0; }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            const double wgt = (st == 0 || st == 3) ? 1.0 / 6.0 : 1.0 / 3.0;
            const double adv = (st == 2) ? 1.0 : 0.5;
            const int e = 4 * m + st;
            MMPC_UNROLL for (int i = 0; i < NX; ++i) xst[e][i] = xa[i];
            MMPC_UNROLL for (int i = 0; i < NX * S; ++i) Tst[e][i] = Tz[i];
            double k[NX], A[NX * NX], B[NX * NU], kz[NX * S];
            M::fAB(xa, u, k, A, B);
            MMPC_UNROLL for (int i = 0; i < NX; ++i) MMPC_UNROLL for (int j = 0; j < S; ++j) {
                double t = j >= NX ? B[i * NU + j - NX] : 0.0;
                MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(A[i * NX + l], Tz[l * S + j], t);
                kz[i * S + j] = t;
            }
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                acc[i] = fma(wgt, k[i], acc[i]);
                xa[i] = fma(adv * hs, k[i], xs[i]);
            }
            MMPC_UNROLL for (int i = 0; i < NX * S; ++i) { accz[i] = fma(wgt, kz[i], accz[i]); Tz[i] = fma(adv * hs, kz[i], Sz[i]); }
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) xs[i] = fma(hs, acc[i], xs[i]);
        MMPC_UNROLL for (int i = 0; i < NX * S; ++i) Sz[i] = fma(hs, accz[i], Sz[i]);
    }
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        F[i] = xs[i];
        MMPC_UNROLL for (int j = 0; j < NX; ++j) Fx[i * NX + j] = Sz[i * S + j];
        MMPC_UNROLL for (int j = 0; j < NU; ++j) Fu[i * NU + j] = Sz[i * S + NX + j];
    }
    // ---- reverse sweep ----
    double v[NX], zero[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) { v[i] = (wq ? wq[i] * (xs[i] - r[i]) : 0.0) + lam[i]; zero[i] = 0.0; }
    MMPC_UNROLL for (int i = 0; i < NH; ++i) Hs[i] = 0.0;
    for (int m = substeps - 1; m >= 0; --m) {
        double xsum[NX], xbar_next[NX];  // sum of the stage-point adjoints; adjoint of the NEXT stage point (feeds this stage's k)
        MMPC_UNROLL for (int i = 0; i < NX; ++i) { xsum[i] = 0.0; xbar_next[i] = 0.0; }
#pragma unroll 1
        for (int st = 3; st >= 0; --st) {
            const double wgt = (st == 0 || st == 3) ? 1.0 / 6.0 : 1.0 / 3.0;
            const double adv_next = (st == 3) ? 0.0 : ((st == 2) ? 1.0 : 0.5);  // stage st+1's point = xs + adv_next hs k_st
            const int e = 4 * m + st;
            double kbar[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) kbar[i] = hs * fma(wgt, v[i], adv_next * xbar_next[i]);
            double kk[NX], A[NX * NX], B[NX * NU], Hl[NH], xa[NX], T[NX * S];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) xa[i] = xst[e][i];
            MMPC_UNROLL for (int i = 0; i < NX * S; ++i) T[i] = Tst[e][i];
            M::fABH(xa, u, zero, kbar, kk, A, B, Hl);
            // Hs += W' Hl W,  W = [T; (0 I)]  (S x S): first Y = Hl W (rows p of the symmetric Hl), then W' Y, upper triangle
            double Y[S * S];
            MMPC_UNROLL for (int p = 0; p < S; ++p) MMPC_UNROLL for (int j = 0; j < S; ++j) {
                double t = (j >= NX) ? Hl[trisym(p, j, S)] : 0.0;  // u rows of W are (0 I)
                MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(Hl[trisym(p, l, S)], T[l * S + j], t);
                Y[p * S + j] = t;
            }
            MMPC_UNROLL for (int a = 0; a < S; ++a) MMPC_UNROLL for (int j = a; j < S; ++j) {
                double t = (a >= NX) ? Y[a * S + j] : 0.0;
                MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(T[l * S + a], Y[l * S + j], t);
                Hs[tri(a, j, S)] += t;
            }
            // adjoint of this stage's point: A' kbar
            MMPC_UNROLL for (int j = 0; j < NX; ++j) {
                double t = 0.0;
                MMPC_UNROLL for (int i = 0; i < NX; ++i) t = fma(A[i * NX + j], kbar[i], t);
                xbar_next[j] = t;
                xsum[j] += t;
            }
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) v[i] += xsum[i];
    }
}

// ---------------------------------------------------------------------------------------------
// eval_nodes: n independent shooting nodes, SoA in / SoA out, one thread per node (HBM-bound stream)
// ---------------------------------------------------------------------------------------------
template <class M, bool WITH_H, bool RK4>
__global__ void __launch_bounds__(128) eval_nodes_kernel(double h, int rk4_substeps, long long n, const double* __restrict__ x,
                                                         const double* __restrict__ u, const double* __restrict__ xn,
                                                         const double* __restrict__ w, double* __restrict__ g,
                                                         double* __restrict__ Fx, double* __restrict__ Fu,
                                                         double* __restrict__ H) {
    constexpr int NX = M::NX, NU = M::NU, NH = Dim<M>::NH;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        double xr[NX], ur[NU], F[NX], fx[NX * NX], fu[NX * NU];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) xr[i] = x[i * n + j];
        MMPC_UNROLL for (int i = 0; i < NU; ++i) ur[i] = u[i * n + j];
        if (WITH_H) {
            double wr[NX], zero[NX], hs[NH];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                wr[i] = w[i * n + j];
                zero[i] = 0.0;
            }
            if (RK4) step_rk4_hess<M>(h, rk4_substeps, xr, ur, nullptr, nullptr, wr, F, fx, fu, hs);  // exact, through the integrator
            else step_euler<M, 2>(h, xr, ur, zero, zero, wr, F, fx, fu, hs);
            MMPC_UNROLL for (int i = 0; i < NH; ++i) __stcs(&H[i * n + j], hs[i]);
        } else if (RK4) {  // separate instantiation: keeps the Euler streaming kernel at its low register count
            step_rk4<M, 1>(h, rk4_substeps, xr, ur, F, fx, fu);
        } else {
            step_euler<M, 1>(h, xr, ur, nullptr, nullptr, nullptr, F, fx, fu, nullptr);
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) __stcs(&g[i * n + j], F[i] - xn[i * n + j]);
        MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) __stcs(&Fx[i * n + j], fx[i]);
        MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) __stcs(&Fu[i * n + j], fu[i]);
    }
}

// <name>_get_x_dot_init/_get_A/_get_B: AoS, A and B column-major per point (ModelControl.cpp:127-133)
template <class M>
__global__ void linearise_kernel(int n, const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ xdot,
                                 double* __restrict__ A, double* __restrict__ B) {
    constexpr int NX = M::NX, NU = M::NU;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double xr[NX], ur[NU], xd[NX], a[NX * NX], b[NX * NU];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) xr[i] = x[(size_t)j * NX + i];
    MMPC_UNROLL for (int i = 0; i < NU; ++i) ur[i] = u[(size_t)j * NU + i];
    M::fAB(xr, ur, xd, a, b);
    if (xdot) MMPC_UNROLL for (int i = 0; i < NX; ++i) xdot[(size_t)j * NX + i] = xd[i];
    if (A) MMPC_UNROLL for (int r = 0; r < NX; ++r) MMPC_UNROLL for (int c = 0; c < NX; ++c) A[(size_t)j * NX * NX + c * NX + r] = a[r * NX + c];
    if (B) MMPC_UNROLL for (int r = 0; r < NX; ++r) MMPC_UNROLL for (int c = 0; c < NU; ++c) B[(size_t)j * NX * NU + c * NX + r] = b[r * NU + c];
}

// plant of the closed-loop drivers: one thread per instance applies u_0 of its plan for one step
template <class M>
__global__ void plant_step_kernel(int B, int N, double h, int rk4, const double* __restrict__ V, double* __restrict__ state,
                                  double* __restrict__ control, double* __restrict__ X_log, double* __restrict__ U_log, int cycle, int cycles) {
    constexpr int NX = M::NX, NU = M::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t NV = (size_t)NX * (N + 1) + (size_t)NU * N;
    double x[NX], u[NU], F[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) x[i] = state[(size_t)b * NX + i];
    MMPC_UNROLL for (int j = 0; j < NU; ++j) u[j] = V[(size_t)b * NV + NX + j];  // first move of the plan (ModelControl.cpp:179-183)
    if (X_log) MMPC_UNROLL for (int i = 0; i < NX; ++i) X_log[((size_t)b * cycles + cycle) * NX + i] = x[i];
    if (U_log) MMPC_UNROLL for (int j = 0; j < NU; ++j) U_log[((size_t)b * cycles + cycle) * NU + j] = u[j];
    if (rk4) step_rk4<M, 0>(h, 1, x, u, F, nullptr, nullptr);
    else step_euler<M, 0>(h, x, u, nullptr, nullptr, nullptr, F, nullptr, nullptr, nullptr);
    MMPC_UNROLL for (int i = 0; i < NX; ++i) state[(size_t)b * NX + i] = F[i];
    MMPC_UNROLL for (int j = 0; j < NU; ++j) control[(size_t)b * NU + j] = u[j];
}

// ---------------------------------------------------------------------------------------------
// views of the reference's parameter vector p (ModelGenerator.cpp:129-187)
// ---------------------------------------------------------------------------------------------
template <class M>
struct PView {
    const double *r, *Q, *R, *Rm, *A, *B, *xdot_init, *x_init, *u_init;
    __host__ __device__ static int np(int N, int is_linear) {
        return N * M::NX + M::NX + 3 * M::NU + (is_linear ? M::NX * M::NX + M::NX * M::NU + 2 * M::NX : 0);
    }
    __device__ PView(const double* p, int N, int is_linear) {
        r = p; p += N * M::NX;
        Q = p; p += M::NX;
        R = p; p += M::NU;
        Rm = p; p += M::NU;
        A = B = xdot_init = x_init = nullptr;
        if (is_linear) {
            A = p; p += M::NX * M::NX;
            B = p; p += M::NX * M::NU;
            xdot_init = p; p += M::NX;
            x_init = p; p += M::NX;
        }
        u_init = p;
    }
};

// One (instance, node) of the whole-NLP callbacks. grid: (B, ceil(N/128)) -- the batch on grid.x (no 65535 limit), block 128.
// f and grad_f are accumulated with atomics (f[b] and the u_{k-1} coupling entries), so the caller
// zero-fills them first.
template <class M>
__global__ void __launch_bounds__(128) nlp_eval_kernel(mmpc_problem P, int B, const double* __restrict__ V,
                                                       const double* __restrict__ p, const double* __restrict__ lamg,
                                                       double lam_f, double* __restrict__ f, double* __restrict__ g,
                                                       double* __restrict__ grad_f, double* __restrict__ Fx,
                                                       double* __restrict__ Fu, double* __restrict__ Hzz) {
    constexpr int NX = M::NX, NU = M::NU, S = Dim<M>::S, NH = Dim<M>::NH;
    const int k = blockIdx.y * blockDim.x + threadIdx.x, b = blockIdx.x, N = P.N;
    if (k >= N || b >= B) return;
    const int NV = NX * (N + 1) + NU * N, Mc = N * NX;
    const double* Vb = V + (size_t)b * NV;
    const PView<M> pv(p + (size_t)b * PView<M>::np(N, P.is_linear), N, P.is_linear);
    double x[NX], u[NU], up[NU], xn[NX], r[NX], lam[NX], wq[NX], wqh[NX], F[NX], fx[NX * NX], fu[NX * NU], hs[NH], e[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        x[i] = Vb[k * S + i];
        xn[i] = Vb[(k + 1) * S + i];
        r[i] = pv.r[k * NX + i];
        lam[i] = lamg ? lamg[(size_t)b * Mc + k * NX + i] : 0.0;
        wq[i] = 2.0 * lam_f * pv.Q[i];
        wqh[i] = P.node_cost ? 0.0 : wq[i];  // MMPC_COST_NODE: the tracking term does not see F, the Hessian weights are lam alone
    }
    const bool node = P.node_cost != 0;
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        u[j] = Vb[k * S + NX + j];
        up[j] = k ? Vb[(k - 1) * S + NX + j] : pv.u_init[j];
    }
    if (P.is_linear) {
        double A[NX * NX], Bm[NX * NU];  // p holds them column-major
        MMPC_UNROLL for (int rr = 0; rr < NX; ++rr) {
            MMPC_UNROLL for (int c = 0; c < NX; ++c) A[rr * NX + c] = pv.A[c * NX + rr];
            MMPC_UNROLL for (int c = 0; c < NU; ++c) Bm[rr * NU + c] = pv.B[c * NX + rr];
        }
        step_linear<M, 2>(P.h, x, u, A, Bm, pv.xdot_init, pv.x_init, pv.u_init, F, fx, fu, hs);
    } else if (P.integrator == MMPC_RK4) {
        if (Hzz && !P.gauss_newton) {
            step_rk4_hess<M>(P.h, P.substeps, x, u, wqh, r, lam, F, fx, fu, hs);
        } else {
            step_rk4<M, 1>(P.h, P.substeps, x, u, F, fx, fu);
            MMPC_UNROLL for (int i = 0; i < NH; ++i) hs[i] = 0.0;  // Gauss-Newton part only
        }
    } else {
        step_euler<M, 2>(P.h, x, u, wqh, r, lam, F, fx, fu, hs);
    }
    double fk = 0.0;
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        e[i] = (node ? xn[i] : F[i]) - r[i];
        fk = fma(pv.Q[i] * e[i], e[i], fk);
        if (g) g[(size_t)b * Mc + k * NX + i] = F[i] - xn[i];
    }
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        const double du = u[j] - up[j];
        fk = fma(pv.R[j] * du, du, fk);
        fk = fma(pv.Rm[j] * u[j], u[j], fk);
    }
    if (f) atomicAdd(&f[b], fk);
    if (Fx) MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) Fx[((size_t)b * N + k) * NX * NX + i] = fx[i];
    if (Fu) MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Fu[((size_t)b * N + k) * NX * NU + i] = fu[i];
    if (grad_f) {
        double* gb = grad_f + (size_t)b * NV;
        MMPC_UNROLL for (int j = 0; j < NX; ++j) {
            double a = 0.0;
            if (node) a = k ? 2.0 * pv.Q[j] * (x[j] - pv.r[(k - 1) * NX + j]) : 0.0;  // d/dx_k of (x_k - r_{k-1})' Q (x_k - r_{k-1})
            else MMPC_UNROLL for (int i = 0; i < NX; ++i) a = fma(fx[i * NX + j], 2.0 * pv.Q[i] * e[i], a);
            gb[k * S + j] = a;  // x_k gets a contribution from node k only
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            double a = 0.0;
            if (!node) MMPC_UNROLL for (int i = 0; i < NX; ++i) a = fma(fu[i * NU + j], 2.0 * pv.Q[i] * e[i], a);
            const double du = u[j] - up[j];
            atomicAdd(&gb[k * S + NX + j], a + 2.0 * pv.R[j] * du + 2.0 * pv.Rm[j] * u[j]);
            if (k) atomicAdd(&gb[(k - 1) * S + NX + j], -2.0 * pv.R[j] * du);
        }
        if (k == N - 1) MMPC_UNROLL for (int j = 0; j < NX; ++j) gb[N * S + j] = node ? 2.0 * pv.Q[j] * e[j] : 0.0;
    }
    if (Hzz) {
        double* Hb = Hzz + ((size_t)b * N + k) * S * S;
        MMPC_UNROLL for (int a = 0; a < S; ++a) MMPC_UNROLL for (int c = 0; c < S; ++c) {
            double acc = hs[trisym(a, c, S)];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                const double fa = a < NX ? fx[i * NX + a] : fu[i * NU + a - NX];
                const double fc = c < NX ? fx[i * NX + c] : fu[i * NU + c - NX];
                acc = fma(wqh[i] * fa, fc, acc);
            }
            if (node && a == c && a < NX && k) acc += wq[a];  // constant block of the tracking term on x_k (x_N's block, 2 lam_f Q, is implicit)
            if (a == c && a >= NX) acc += 2.0 * lam_f * (pv.R[a - NX] + pv.Rm[a - NX]) + (k < N - 1 ? 2.0 * lam_f * pv.R[a - NX] : 0.0);
            Hb[a * S + c] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// solve_kernel, variant T1: one thread per MPC instance, everything fused.
//
// Per Newton iteration (unbounded case; the problem is equality-constrained because x_0 is pinned,
// ModelControl.cpp:144-145, and all other bounds default to +-10e30):
//   backward sweep k = N-1..0 : evaluate F, Fx, Fu, Lagrangian stage Hessian at (z_k, lam_k); accumulate the
//                               KKT residual norms of the CURRENT iterate; Riccati update of the value
//                               function of the augmented state (x_k, u_{k-1}); store gains and P.
//   -> converged? (IPOPT's scaled optimality error <= tol)  -> write out
//   forward sweep k = 0..N-1  : dz from the gains, new multipliers from P, directional derivative of the l1 merit
//                               function, and -- fused -- the merit value of the full step (alpha = 1)
//   backtracking (rare)       : re-evaluate the merit at alpha/2, ...
// ---------------------------------------------------------------------------------------------
template <class M>
struct RiccatiState {
    double Pxx[M::NX * M::NX], Pxv[M::NX * M::NU], Pvv[M::NU * M::NU], px[M::NX], pv[M::NU];
};

// in-place Cholesky of the NU x NU matrix G (full storage, lower triangle used); the diagonal holds 1/l_jj so
// the solves multiply instead of divide. returns false if not PD.
template <int NU>
__device__ __forceinline__ bool chol(double* G) {
    bool ok = true;
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        double d = G[j * NU + j];
        MMPC_UNROLL for (int k = 0; k < NU; ++k) if (k < j) d = fma(-G[j * NU + k], G[j * NU + k], d);
        ok = ok && (d > 0.0) && isfinite(d);
        const double il = rsqrt(d > 0.0 ? d : 1.0);
        G[j * NU + j] = il;
        MMPC_UNROLL for (int i = 0; i < NU; ++i) if (i > j) {
            double s = G[i * NU + j];
            MMPC_UNROLL for (int k = 0; k < NU; ++k) if (k < j) s = fma(-G[i * NU + k], G[j * NU + k], s);
            G[i * NU + j] = s * il;
        }
    }
    return ok;
}
// solve L L' y = b in place (L from chol)
template <int NU>
__device__ __forceinline__ void chol_solve(const double* L, double* b) {
    MMPC_UNROLL for (int i = 0; i < NU; ++i) {
        double s = b[i];
        MMPC_UNROLL for (int k = 0; k < NU; ++k) if (k < i) s = fma(-L[i * NU + k], b[k], s);
        b[i] = s * L[i * NU + i];
    }
    MMPC_UNROLL for (int i = NU - 1; i >= 0; --i) {
        double s = b[i];
        MMPC_UNROLL for (int k = 0; k < NU; ++k) if (k > i) s = fma(-L[k * NU + i], b[k], s);
        b[i] = s * L[i * NU + i];
    }
}

}  // namespace mmpc
#include "mmpc_solve_t1.cuh"
#include "mmpc_solve_t2.cuh"
#include "mmpc_solve_t3.cuh"
#include "mmpc_solve_t4.cuh"
namespace mmpc {

// ---------------------------------------------------------------------------------------------
// host-side launchers + vtable
// ---------------------------------------------------------------------------------------------
template <class M>
struct Launch {
    static cudaError_t eval_nodes(const mmpc_problem& P, long long n, const double* x, const double* u, const double* xn,
                                  const double* w, double* g, double* Fx, double* Fu, double* H, cudaStream_t st, int* launches) {
        if (P.is_linear) return cudaErrorNotSupported;
        if (P.integrator == MMPC_RK4 && w && P.substeps > MMPC_RK4_HESS_MAX_SUBSTEPS) return cudaErrorNotSupported;
        if (n <= 0) return cudaSuccess;
        const int rk = P.integrator == MMPC_RK4 ? P.substeps : 0;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long want = (n + 127) / 128;
        const int grid = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);  // grid-stride, multiple of the SM count
        if (w && rk) eval_nodes_kernel<M, true, true><<<grid, 128, 0, st>>>(P.h, rk, n, x, u, xn, w, g, Fx, Fu, H);
        else if (w) eval_nodes_kernel<M, true, false><<<grid, 128, 0, st>>>(P.h, 0, n, x, u, xn, w, g, Fx, Fu, H);
        else if (rk) eval_nodes_kernel<M, false, true><<<grid, 128, 0, st>>>(P.h, rk, n, x, u, xn, w, g, Fx, Fu, H);
        else eval_nodes_kernel<M, false, false><<<grid, 128, 0, st>>>(P.h, 0, n, x, u, xn, w, g, Fx, Fu, H);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    static cudaError_t linearise(int n, const double* x, const double* u, double* xdot, double* A, double* B, cudaStream_t st,
                                 int* launches) {
        if (n <= 0) return cudaSuccess;
        linearise_kernel<M><<<(n + 127) / 128, 128, 0, st>>>(n, x, u, xdot, A, B);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    static cudaError_t nlp_eval(const mmpc_problem& P, int B, const double* V, const double* p, const double* lam, double lam_f,
                                double* f, double* g, double* grad_f, double* Fx, double* Fu, double* Hzz, cudaStream_t st,
                                int* launches) {
        if (P.integrator == MMPC_RK4 && P.is_linear) return cudaErrorNotSupported;
        if (B <= 0) return cudaSuccess;
        const int NV = M::NX * (P.N + 1) + M::NU * P.N;
        cudaError_t e;
        if (f && (e = cudaMemsetAsync(f, 0, sizeof(double) * B, st)) != cudaSuccess) return e;
        if (grad_f && (e = cudaMemsetAsync(grad_f, 0, sizeof(double) * (size_t)B * NV, st)) != cudaSuccess) return e;
        dim3 grid(B, (P.N + 127) / 128);
        nlp_eval_kernel<M><<<grid, 128, 0, st>>>(P, B, V, p, lam, lam_f, f, g, grad_f, Fx, Fu, Hzz);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    static cudaError_t plant_step(int B, int N, double h, int plant_integrator, const double* V, double* state, double* control,
                                  double* X_log, double* U_log, int cycle, int cycles, cudaStream_t st, int* launches) {
        if (B <= 0) return cudaSuccess;
        plant_step_kernel<M><<<(B + 127) / 128, 128, 0, st>>>(B, N, h, plant_integrator == MMPC_RK4, V, state, control, X_log, U_log, cycle, cycles);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    // per-variant launchers; each may live in its own translation unit (MMPC_DEFINE_SOLVE_T*)
    // T1's two instantiations (free / interior point) are the longest single kernels to compile: one translation unit each
    template <bool BOUNDED>
    static cudaError_t solve_t1(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        const int grid = (io.b1 - io.b0 + 31) / 32;
        solve_kernel_t1<M, BOUNDED><<<grid, 32, 0, st>>>(P, W, io);
        return cudaGetLastError();
    }
    static cudaError_t solve_t2(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        return LaunchT2<M>::run(P, W, io, st);
    }
    template <int PART>
    static cudaError_t solve_t3(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        return LaunchT3<M>::template run_part<PART>(P, W, io, st);
    }
    template <int PART>
    static cudaError_t solve_t4(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        return LaunchT4<M>::template run_part<PART>(P, W, io, st);
    }
};

}  // namespace mmpc

// ---- translation-unit macros --------------------------------------------------------------------------------
// The solve variants of a model can be compiled in separate units (parallel builds of large generated models):
#define MMPC_DEFINE_SOLVE_T1(MODEL)                                                                                    \
    extern "C" cudaError_t mmpc_solve_t1b_##MODEL(const mmpc_problem*, const mmpc_workspace*, const mmpc_solve_io*, cudaStream_t); \
    extern "C" cudaError_t mmpc_solve_t1_##MODEL(const mmpc_problem* P, const mmpc_workspace* W, const mmpc_solve_io* io, cudaStream_t st) { \
        if (P->bounded) return mmpc_solve_t1b_##MODEL(P, W, io, st);                                                    \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t1<false>(*P, *W, *io, st);                                      \
    }
#define MMPC_DEFINE_SOLVE_T1B(MODEL)                                                                                   \
    extern "C" cudaError_t mmpc_solve_t1b_##MODEL(const mmpc_problem* P, const mmpc_workspace* W, const mmpc_solve_io* io, cudaStream_t st) { \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t1<true>(*P, *W, *io, st);                                       \
    }
#define MMPC_DEFINE_SOLVE_T2(MODEL)                                                                                    \
    extern "C" cudaError_t mmpc_solve_t2_##MODEL(const mmpc_problem* P, const mmpc_workspace* W, const mmpc_solve_io* io, cudaStream_t st) { \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t2(*P, *W, *io, st);                                             \
    }
// T3 / T4: part 0 (with the dispatcher over the parts) and one unit per further part (LaunchT3 / LaunchT4::run_part)
#define MMPC_SOLVE_FN_ARGS const mmpc_problem* P, const mmpc_workspace* W, const mmpc_solve_io* io, cudaStream_t st
#define MMPC_DEFINE_SOLVE_T3(MODEL)                                                                                    \
    extern "C" cudaError_t mmpc_solve_t3p1_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t3p2_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t3_##MODEL(MMPC_SOLVE_FN_ARGS) {                                                  \
        const int part = mmpc::LaunchT3<mmpc_model_##MODEL>::part_of(*P);                                               \
        if (part == 1) return mmpc_solve_t3p1_##MODEL(P, W, io, st);                                                    \
        if (part == 2) return mmpc_solve_t3p2_##MODEL(P, W, io, st);                                                    \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t3<0>(*P, *W, *io, st);                                          \
    }
#define MMPC_DEFINE_SOLVE_T3_PART(MODEL, K)                                                                            \
    extern "C" cudaError_t mmpc_solve_t3p##K##_##MODEL(MMPC_SOLVE_FN_ARGS) {                                            \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t3<K>(*P, *W, *io, st);                                          \
    }
// MMPC_DEFAULT_TILES_ONLY (set by ModelGenerator::compile_model): parts 1 and 3 -- tile widths that are only ever selected on
// request (tests of other tile shapes, kernel_variant = 4 + 256 k) -- are not built for a user model; asking for them is
// MMPC_ERR_UNSUPPORTED.
#ifdef MMPC_DEFAULT_TILES_ONLY
#define MMPC_T4_EXTRA_PART(MODEL, K) cudaErrorNotSupported
#else
#define MMPC_T4_EXTRA_PART(MODEL, K) mmpc_solve_t4p##K##_##MODEL(P, W, io, st)
#endif
#define MMPC_DEFINE_SOLVE_T4(MODEL)                                                                                    \
    extern "C" cudaError_t mmpc_solve_t4p1_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t4p2_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t4p3_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t4p4_##MODEL(MMPC_SOLVE_FN_ARGS);                                                 \
    extern "C" cudaError_t mmpc_solve_t4_##MODEL(MMPC_SOLVE_FN_ARGS) {                                                  \
        const int part = mmpc::LaunchT4<mmpc_model_##MODEL>::part_of(*P, *W, *io);                                      \
        if (part == 1) return MMPC_T4_EXTRA_PART(MODEL, 1);                                                             \
        if (part == 2) return mmpc_solve_t4p2_##MODEL(P, W, io, st);                                                    \
        if (part == 3) return MMPC_T4_EXTRA_PART(MODEL, 3);                                                             \
        if (part == 4) return mmpc_solve_t4p4_##MODEL(P, W, io, st);                                                    \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t4<0>(*P, *W, *io, st);                                          \
    }
#define MMPC_DEFINE_SOLVE_T4_PART(MODEL, K)                                                                            \
    extern "C" cudaError_t mmpc_solve_t4p##K##_##MODEL(MMPC_SOLVE_FN_ARGS) {                                            \
        return mmpc::Launch<mmpc_model_##MODEL>::solve_t4<K>(*P, *W, *io, st);                                          \
    }
// ... and this one defines the evaluators, the solve dispatcher and `mmpc_vtable_<MODEL>()`:
#define MMPC_DEFINE_MODEL_CORE(MODEL)                                                                                   \
    extern "C" cudaError_t mmpc_solve_t1_##MODEL(const mmpc_problem*, const mmpc_workspace*, const mmpc_solve_io*, cudaStream_t); \
    extern "C" cudaError_t mmpc_solve_t2_##MODEL(const mmpc_problem*, const mmpc_workspace*, const mmpc_solve_io*, cudaStream_t); \
    extern "C" cudaError_t mmpc_solve_t3_##MODEL(const mmpc_problem*, const mmpc_workspace*, const mmpc_solve_io*, cudaStream_t); \
    extern "C" cudaError_t mmpc_solve_t4_##MODEL(const mmpc_problem*, const mmpc_workspace*, const mmpc_solve_io*, cudaStream_t); \
    static cudaError_t mmpc_solve_dispatch_##MODEL(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io,  \
                                                   cudaStream_t st, int* launches) {                                    \
        if (P.integrator == MMPC_RK4 && P.is_linear) return cudaErrorNotSupported;                                      \
        /* bounds and RK4: T1, and T3 for small models (the ABI layer switches variants) */                              \
        if (W.variant == 2 && (P.bounded || P.node_cost)) return cudaErrorNotSupported;                                 \
        if (W.variant == 2 && P.integrator != MMPC_EULER) return cudaErrorNotSupported;                                 \
        if (launches) ++*launches;                                                                                      \
        if (W.variant == 2) return mmpc_solve_t2_##MODEL(&P, &W, &io, st);                                              \
        if (W.variant == 3) return mmpc_solve_t3_##MODEL(&P, &W, &io, st);                                              \
        if (W.variant == 4) return mmpc_solve_t4_##MODEL(&P, &W, &io, st);                                              \
        return mmpc_solve_t1_##MODEL(&P, &W, &io, st);                                                                  \
    }                                                                                                                   \
    extern "C" const mmpc_model_vtable* mmpc_vtable_##MODEL(void) {                                                      \
        using M = mmpc_model_##MODEL;                                                                                    \
        static const mmpc_model_vtable vt = {MMPC_VTABLE_VERSION, M::MODEL_NAME, M::NX, M::NU,                           \
                                             {M::OPS_F, M::TRANS_F, M::OPS_FAB, M::TRANS_FAB, M::OPS_FABH, M::TRANS_FABH}, \
                                             &mmpc::Launch<M>::eval_nodes, &mmpc::Launch<M>::linearise,                  \
                                             &mmpc::Launch<M>::nlp_eval, &mmpc_solve_dispatch_##MODEL,                   \
                                             &mmpc::Launch<M>::plant_step};                                              \
        return &vt;                                                                                                      \
    }
// everything in one unit (what ModelGenerator::generate_c_code emits for a user model)
#define MMPC_DEFINE_MODEL(MODEL) \
    MMPC_DEFINE_SOLVE_T1(MODEL) MMPC_DEFINE_SOLVE_T1B(MODEL) MMPC_DEFINE_SOLVE_T2(MODEL)                                  \
    MMPC_DEFINE_SOLVE_T3(MODEL) MMPC_DEFINE_SOLVE_T3_PART(MODEL, 1) MMPC_DEFINE_SOLVE_T3_PART(MODEL, 2)                    \
    MMPC_DEFINE_SOLVE_T4(MODEL) MMPC_DEFINE_SOLVE_T4_PART(MODEL, 1) MMPC_DEFINE_SOLVE_T4_PART(MODEL, 2) MMPC_DEFINE_SOLVE_T4_PART(MODEL, 3) \
    MMPC_DEFINE_SOLVE_T4_PART(MODEL, 4)                                                                                   \
    MMPC_DEFINE_MODEL_CORE(MODEL)
// Additionally exports the generic entry point a stand-alone generated model library is dlopen'ed by
// (the role `nlpsol("nlpsol","ipopt",dll,...)` plays at ModelControl.cpp:61-62).
// Generated libraries are compiled with -fvisibility=hidden: only this entry is exported, so a user model whose
// identifier equals a built-in's cannot be interposed by the copies inside libmahi_mpc_b200.so (global scope).
#define MMPC_EXPORT_MODEL_ENTRY(MODEL) \
    extern "C" __attribute__((visibility("default"))) const mmpc_model_vtable* mmpc_model_vtable_get(void) { return mmpc_vtable_##MODEL(); }

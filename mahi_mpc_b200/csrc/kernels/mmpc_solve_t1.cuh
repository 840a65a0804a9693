// mahi-mpc-b200 :: solve kernel, variant T1 -- one thread per MPC instance, everything fused.
//
// The reference algorithm this replaces is m_solver(m_solver_args) = IPOPT + MUMPS
// (src/Mahi/Mpc/ModelControl.cpp:159) on the NLP of ModelGenerator.cpp:23-233.  Here:
//   * exact-Hessian Newton on the KKT conditions; the KKT system of every iteration is solved by a stage-wise
//     Riccati recursion on the augmented state (x_k, u_{k-1}) (the cost couples consecutive controls through
//     du'R du, ModelGenerator.cpp:217-218);
//   * finite bounds (ModelParameters x_min..u_max, ModelControl::update_control_limits) turn it into a
//     primal-dual interior-point method: log barrier with monotone mu (Fiacco-McCormick), Sigma = z/s on the
//     stage Hessian diagonal, fraction-to-the-boundary rule, IPOPT's kappa_sigma safeguard [upstream:
//     Waechter & Biegler 2006];
//   * l1-merit backtracking line search; inertia correction by a diagonal shift when a G_uu is not PD;
//   * termination on IPOPT's scaled optimality error (eq. 5-6 of the same paper), tol/max_iter from
//     ModelControl.cpp:54-55.
// T1 is the latency-bound baseline variant (one warp per 32 instances); T2 (mmpc_solve_t2.cuh) spreads an
// instance over several lanes.  T1 is the variant that implements bounds.
#pragma once

namespace mmpc {

#define MMPC_INF_THRESH 1e19  // IPOPT nlp_{lower,upper}_bound_inf [upstream]

template <class M>
struct SolveCtx {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, NH = S * (S + 1) / 2;
    const mmpc_problem& P;
    const mmpc_workspace& W;
    const mmpc_solve_io& io;
    int b;  // instance
    size_t Bp;
    double Q[NX], R[NU], Rm[NU], uinit[NU];
    __device__ SolveCtx(const mmpc_problem& P_, const mmpc_workspace& W_, const mmpc_solve_io& io_, int b_)
        : P(P_), W(W_), io(io_), b(b_), Bp(W_.Bp) {}
    __device__ __forceinline__ double ld(const double* a, int row) const { return a[(size_t)row * Bp + b]; }
    __device__ __forceinline__ void st(double* a, int row, double v) const { a[(size_t)row * Bp + b] = v; }
    __device__ __forceinline__ double traj(int k, int i) const { return io.traj[((size_t)b * P.N + k) * NX + i]; }

    template <int WANT>
    __device__ __forceinline__ void step(const double* x, const double* u, const double* wq, const double* r, const double* lam,
                                         double* F, double* Fx, double* Fu, double* Hs) const {
        if (P.is_linear) {
            double A[NX * NX], Bm[NX * NU], xd0[NX], x0[NX];
            MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) A[i] = ld(W.LIN, i);
            MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Bm[i] = ld(W.LIN, NX * NX + i);
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                xd0[i] = ld(W.LIN, NX * NX + NX * NU + i);
                x0[i] = ld(W.X, i);  // x_init = pinned state (ModelControl.cpp:134)
            }
            step_linear<M, WANT>(P.h, x, u, A, Bm, xd0, x0, uinit, F, Fx, Fu, Hs);
        } else if (P.integrator == MMPC_RK4) {
            if (WANT >= 2 && !P.gauss_newton) {
                step_rk4_hess<M>(P.h, P.substeps, x, u, wq, r, lam, F, Fx, Fu, Hs);  // exact Newton through the integrator
            } else {
                step_rk4<M, (WANT >= 1 ? 1 : 0)>(P.h, P.substeps, x, u, F, Fx, Fu);
                if (WANT >= 2) MMPC_UNROLL for (int i = 0; i < NH; ++i) Hs[i] = 0.0;  // Gauss-Newton Hessian on request
            }
        } else {
            step_euler<M, WANT>(P.h, x, u, wq, r, lam, F, Fx, Fu, Hs);
        }
    }
};

struct SweepNorms {
    double dual_inf, prim_inf, theta, fval, lam_l1;
    double z_l1, cmax, cmin, barrier;  // interior-point extras: |z|_1, max/min of s*z, -sum log s
};

// barrier bookkeeping of one bounded scalar v in [lb, ub] with multipliers zl, zu
struct BoundTerm {
    double sigma, gphi, dres;  // z/s sum, barrier gradient -mu/sl + mu/su, -zl + zu
};
__device__ __forceinline__ BoundTerm bound_term(double v, double lb, double ub, double zl, double zu, double mu, SweepNorms& n) {
    BoundTerm t{0.0, 0.0, 0.0};
    double sprod = 1.0;  // product of the slacks: one logarithm per variable (see barrier_of)
    if (lb > -MMPC_INF_THRESH) {
        const double s = v - lb;
        t.sigma += zl / s; t.gphi -= mu / s; t.dres -= zl;
        n.z_l1 += fabs(zl); n.cmax = fmax(n.cmax, s * zl); n.cmin = fmin(n.cmin, s * zl); sprod *= s;
    }
    if (ub < MMPC_INF_THRESH) {
        const double s = ub - v;
        t.sigma += zu / s; t.gphi += mu / s; t.dres += zu;
        n.z_l1 += fabs(zu); n.cmax = fmax(n.cmax, s * zu); n.cmin = fmin(n.cmin, s * zu); sprod *= s;
    }
    if (lb > -MMPC_INF_THRESH || ub < MMPC_INF_THRESH) n.barrier -= log(sprod);
    return t;
}

// backward sweep over the horizon at the current iterate.
//   RICCATI = false : residual norms only (needs F, Fx, Fu)
//   RICCATI = true  : norms + Riccati recursion; gains and value-function blocks are stored for the forward sweep.
// returns false if some G_uu was not positive definite (caller regularises and retries)
template <class M, bool BOUNDED, bool RICCATI>
__device__ bool backward_sweep(const SolveCtx<M>& c, double delta, double mu, SweepNorms& nrm) {
    constexpr int NX = M::NX, NU = M::NU, S = NX + NU, NH = S * (S + 1) / 2;
    const int N = c.P.N;
    double Pxx[NX * NX], Pxv[NX * NU], Pvv[NU * NU], px[NX], pv[NU];
    MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) Pxx[i] = 0.0;
    MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) Pxv[i] = 0.0;
    MMPC_UNROLL for (int i = 0; i < NU * NU; ++i) Pvv[i] = 0.0;
    MMPC_UNROLL for (int i = 0; i < NX; ++i) px[i] = 0.0;
    MMPC_UNROLL for (int i = 0; i < NU; ++i) pv[i] = 0.0;
    double wq[NX], wqh[NX];
    const bool node = c.P.node_cost != 0;  // MMPC_COST_NODE: tracking error on the decision variable x_{k+1}
    MMPC_UNROLL for (int i = 0; i < NX; ++i) { wq[i] = 2.0 * c.Q[i]; wqh[i] = node ? 0.0 : wq[i]; }
    double du_next[NU];  // 2 R (u_{k+1} - u_k): enters the u_k stationarity residual
    MMPC_UNROLL for (int j = 0; j < NU; ++j) du_next[j] = 0.0;
    double xn[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) xn[i] = c.ld(c.W.X, N * NX + i);
    nrm.dual_inf = nrm.prim_inf = nrm.theta = nrm.fval = nrm.lam_l1 = 0.0;
    nrm.z_l1 = nrm.cmax = nrm.barrier = 0.0;
    nrm.cmin = INFINITY;
    double xn_dres[NX];  // bound-multiplier part of the stationarity residual of x_{k+1}
    MMPC_UNROLL for (int i = 0; i < NX; ++i) xn_dres[i] = 0.0;
    if (BOUNDED) {  // terminal node x_N carries only its barrier
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            const BoundTerm t = bound_term(xn[i], c.P.x_min[i], c.P.x_max[i], c.ld(c.W.ZLX, (N - 1) * NX + i),
                                           c.ld(c.W.ZUX, (N - 1) * NX + i), mu, nrm);
            Pxx[i * NX + i] = t.sigma;
            px[i] = t.gphi;
            xn_dres[i] = t.dres;
        }
    }
    if (node) {  // the tracking term of x_N is the terminal value function
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            Pxx[i * NX + i] += wq[i];
            px[i] += wq[i] * (xn[i] - c.traj(N - 1, i));
        }
    }
    bool ok = true;
    for (int k = N - 1; k >= 0; --k) {
        double x[NX], u[NU], v[NU], r[NX], lam[NX], lam_prev[NX];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            x[i] = c.ld(c.W.X, k * NX + i);
            r[i] = c.traj(k, i);
            lam[i] = c.ld(c.W.LAM, k * NX + i);
            lam_prev[i] = k ? c.ld(c.W.LAM, (k - 1) * NX + i) : 0.0;
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            u[j] = c.ld(c.W.U, k * NU + j);
            v[j] = k ? c.ld(c.W.U, (k - 1) * NU + j) : c.uinit[j];
        }
        double F[NX], Fx[NX * NX], Fu[NX * NU], H[NH];
        c.template step<RICCATI ? 2 : 1>(x, u, wqh, r, lam, F, Fx, Fu, H);
        double cres[NX], qe[NX];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            const double e = (node ? xn[i] : F[i]) - r[i];
            cres[i] = F[i] - xn[i];
            qe[i] = wq[i] * e;
            nrm.fval = fma(c.Q[i] * e, e, nrm.fval);
            nrm.prim_inf = fmax(nrm.prim_inf, fabs(cres[i]));
            nrm.theta += fabs(cres[i]);
            nrm.lam_l1 += fabs(lam[i]);
            // stationarity w.r.t. x_{k+1}: (its own stage terms were added when it was x of stage k+1) - lam_k - zl + zu
            if (k == N - 1) nrm.dual_inf = fmax(nrm.dual_inf, fabs((node ? qe[i] : 0.0) - lam[i] + xn_dres[i]));
        }
        // barrier terms of this stage's variables
        double sgx[NX], gpx[NX], drx[NX], sgu[NU], gpu_[NU], dru[NU];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) { sgx[i] = 0.0; gpx[i] = 0.0; drx[i] = 0.0; }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) { sgu[j] = 0.0; gpu_[j] = 0.0; dru[j] = 0.0; }
        if (BOUNDED) {
            if (k) MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                const BoundTerm t = bound_term(x[i], c.P.x_min[i], c.P.x_max[i], c.ld(c.W.ZLX, (k - 1) * NX + i),
                                               c.ld(c.W.ZUX, (k - 1) * NX + i), mu, nrm);
                sgx[i] = t.sigma; gpx[i] = t.gphi; drx[i] = t.dres;
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                const BoundTerm t = bound_term(u[j], c.P.u_min[j], c.P.u_max[j], c.ld(c.W.ZLU, k * NU + j), c.ld(c.W.ZUU, k * NU + j), mu, nrm);
                sgu[j] = t.sigma; gpu_[j] = t.gphi; dru[j] = t.dres;
            }
        }
        // objective gradient pieces (+ barrier gradient) and the stationarity residual of the current iterate
        double qx[NX], qu[NU], qv[NU];
        MMPC_UNROLL for (int j = 0; j < NX; ++j) {
            double a = 0.0, l = 0.0;
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                if (!node) a = fma(Fx[i * NX + j], qe[i], a);
                l = fma(Fx[i * NX + j], lam[i], l);
            }
            if (node && k) a = wq[j] * (x[j] - c.traj(k - 1, j));  // x_k's own tracking term (x_0 is pinned)
            qx[j] = a + gpx[j];
            if (k) nrm.dual_inf = fmax(nrm.dual_inf, fabs(a + l - lam_prev[j] + drx[j]));
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            double a = 0.0, l = 0.0;
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                if (!node) a = fma(Fu[i * NU + j], qe[i], a);
                l = fma(Fu[i * NU + j], lam[i], l);
            }
            const double du = u[j] - v[j], dRu = 2.0 * c.R[j] * du;
            const double g = a + dRu + 2.0 * c.Rm[j] * u[j];
            qu[j] = g + gpu_[j];
            qv[j] = -dRu;
            nrm.fval = fma(c.R[j] * du, du, nrm.fval);
            nrm.fval = fma(c.Rm[j] * u[j], u[j], nrm.fval);
            nrm.dual_inf = fmax(nrm.dual_inf, fabs(g + l - du_next[j] + dru[j]));
            du_next[j] = dRu;
        }
        if (RICCATI) {
            // full stage Hessian: H += Fz' (2Q) Fz ; uu diagonal += 2(R + Rm) ; + Sigma + delta on the diagonal
            MMPC_UNROLL for (int a = 0; a < S; ++a) MMPC_UNROLL for (int d = a; d < S; ++d) {
                double acc = H[tri(a, d, S)];
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    const double fa = a < NX ? Fx[i * NX + a] : Fu[i * NU + a - NX];
                    const double fd = d < NX ? Fx[i * NX + d] : Fu[i * NU + d - NX];
                    acc = fma(wqh[i] * fa, fd, acc);
                }
                if (a == d) acc += delta + (a >= NX ? 2.0 * (c.R[a - NX] + c.Rm[a - NX]) + sgu[a - NX] : sgx[a] + ((node && k) ? wq[a] : 0.0));
                H[tri(a, d, S)] = acc;
            }
            double y[NX], yv[NU];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                double a = px[i];
                MMPC_UNROLL for (int j = 0; j < NX; ++j) a = fma(Pxx[i * NX + j], cres[j], a);
                y[i] = a;
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                double a = pv[j];
                MMPC_UNROLL for (int i = 0; i < NX; ++i) a = fma(Pxv[i * NU + j], cres[i], a);
                yv[j] = a;
            }
            double PA[NX * NX], PB[NX * NU];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                MMPC_UNROLL for (int j = 0; j < NX; ++j) {
                    double a = 0.0;
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Pxx[i * NX + l], Fx[l * NX + j], a);
                    PA[i * NX + j] = a;
                }
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    double a = Pxv[i * NU + j];
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Pxx[i * NX + l], Fu[l * NU + j], a);
                    PB[i * NU + j] = a;
                }
            }
            double Gxx[NX * NX], Gxu[NX * NU], Guu[NU * NU], gx[NX], gu[NU];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                MMPC_UNROLL for (int j = i; j < NX; ++j) {
                    double a = H[tri(i, j, S)];
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Fx[l * NX + i], PA[l * NX + j], a);
                    Gxx[i * NX + j] = a;
                    Gxx[j * NX + i] = a;
                }
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    double a = H[tri(i, NX + j, S)];
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Fx[l * NX + i], PB[l * NU + j], a);
                    Gxu[i * NU + j] = a;
                }
                double a = qx[i];
                MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Fx[l * NX + i], y[l], a);
                gx[i] = a;
            }
            MMPC_UNROLL for (int i = 0; i < NU; ++i) {
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    double a = (i <= j ? H[tri(NX + i, NX + j, S)] : H[tri(NX + j, NX + i, S)]) + Pvv[i * NU + j];
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) {
                        a = fma(Fu[l * NU + i], PB[l * NU + j], a);
                        a = fma(Pxv[l * NU + i], Fu[l * NU + j], a);
                    }
                    Guu[i * NU + j] = a;
                }
                double a = qu[i] + yv[i];
                MMPC_UNROLL for (int l = 0; l < NX; ++l) a = fma(Fu[l * NU + i], y[l], a);
                gu[i] = a;
            }
            MMPC_UNROLL for (int i = 0; i < NU; ++i) MMPC_UNROLL for (int j = i + 1; j < NU; ++j) {  // symmetrise
                const double a = 0.5 * (Guu[i * NU + j] + Guu[j * NU + i]);
                Guu[i * NU + j] = Guu[j * NU + i] = a;
            }
            ok = chol<NU>(Guu) && ok;
            // gains: Kx = -Guu^-1 Gxu', Kv = Guu^-1 (2R), k0 = -Guu^-1 gu
            double Kx[NU * NX], Kv[NU * NU], k0[NU];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                double col[NU];
                MMPC_UNROLL for (int j = 0; j < NU; ++j) col[j] = -Gxu[i * NU + j];
                chol_solve<NU>(Guu, col);
                MMPC_UNROLL for (int j = 0; j < NU; ++j) Kx[j * NX + i] = col[j];
            }
            MMPC_UNROLL for (int l = 0; l < NU; ++l) {
                double col[NU];
                MMPC_UNROLL for (int j = 0; j < NU; ++j) col[j] = (j == l) ? 2.0 * c.R[l] : 0.0;
                chol_solve<NU>(Guu, col);
                MMPC_UNROLL for (int j = 0; j < NU; ++j) Kv[j * NU + l] = col[j];
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) k0[j] = -gu[j];
            chol_solve<NU>(Guu, k0);
            // store gains and the value function of stage k+1 (needed by the forward sweep for lam_k)
            MMPC_UNROLL for (int i = 0; i < NU * NX; ++i) c.st(c.W.KX, k * NU * NX + i, Kx[i]);
            MMPC_UNROLL for (int i = 0; i < NU * NU; ++i) c.st(c.W.KV, k * NU * NU + i, Kv[i]);
            MMPC_UNROLL for (int i = 0; i < NU; ++i) c.st(c.W.K0, k * NU + i, k0[i]);
            MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) c.st(c.W.PXX, k * NX * NX + i, Pxx[i]);
            MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) c.st(c.W.PXV, k * NX * NU + i, Pxv[i]);
            MMPC_UNROLL for (int i = 0; i < NX; ++i) c.st(c.W.PX, k * NX + i, px[i]);
            // new value function
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                MMPC_UNROLL for (int j = i; j < NX; ++j) {
                    double a = Gxx[i * NX + j];
                    MMPC_UNROLL for (int l = 0; l < NU; ++l) a = fma(Gxu[i * NU + l], Kx[l * NX + j], a);
                    Pxx[i * NX + j] = a;
                    Pxx[j * NX + i] = a;
                }
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    double a = 0.0;
                    MMPC_UNROLL for (int l = 0; l < NU; ++l) a = fma(Gxu[i * NU + l], Kv[l * NU + j], a);
                    Pxv[i * NU + j] = a;
                }
                double a = gx[i];
                MMPC_UNROLL for (int l = 0; l < NU; ++l) a = fma(Gxu[i * NU + l], k0[l], a);
                px[i] = a;
            }
            MMPC_UNROLL for (int i = 0; i < NU; ++i) {
                MMPC_UNROLL for (int j = 0; j < NU; ++j) Pvv[i * NU + j] = (i == j ? 2.0 * c.R[i] : 0.0) - 2.0 * c.R[i] * Kv[i * NU + j];
                pv[i] = qv[i] - 2.0 * c.R[i] * k0[i];
            }
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) xn[i] = x[i];
    }
    return ok;
}

struct FwdOut {
    double gradf_d, lam_inf, f1, theta1, bar1, amax, az;
    bool finite;
};

// bound bookkeeping of one scalar in the forward sweep: multiplier steps and fraction-to-the-boundary limits
__device__ __forceinline__ void bound_step(double v, double d, double lb, double ub, double zl, double zu, double mu, double tau,
                                           double& dzl, double& dzu, FwdOut& o) {
    dzl = 0.0; dzu = 0.0;
    if (lb > -MMPC_INF_THRESH) {
        const double s = v - lb;
        dzl = (mu - zl * d) / s - zl;
        if (d < 0.0) o.amax = fmin(o.amax, -tau * s / d);
        if (dzl < 0.0) o.az = fmin(o.az, -tau * zl / dzl);
        o.bar1 -= log(fmax(s + d, 1e-300));
    }
    if (ub < MMPC_INF_THRESH) {
        const double s = ub - v;
        dzu = (mu + zu * d) / s - zu;
        if (d > 0.0) o.amax = fmin(o.amax, tau * s / d);
        if (dzu < 0.0) o.az = fmin(o.az, -tau * zu / dzu);
        o.bar1 -= log(fmax(s - d, 1e-300));
    }
}
__device__ __forceinline__ double barrier_of(double v, double lb, double ub) {
    const bool hl = lb > -MMPC_INF_THRESH, hu = ub < MMPC_INF_THRESH;
    // a two-sided bound takes one logarithm: -log(sl) - log(su) = -log(sl su) (the slacks are far from the under/overflow range)
    if (hl && hu) return -log(fmax(v - lb, 1e-300) * fmax(ub - v, 1e-300));
    if (hl) return -log(fmax(v - lb, 1e-300));
    if (hu) return -log(fmax(ub - v, 1e-300));
    return 0.0;
}

// forward sweep: step, new multipliers, directional derivative of the (barrier) objective, and -- fused -- the
// merit pieces of the full step.
template <class M, bool BOUNDED>
__device__ void forward_sweep(const SolveCtx<M>& c, double mu, FwdOut& o) {
    constexpr int NX = M::NX, NU = M::NU;
    const int N = c.P.N;
    const double tau = fmax(0.99, 1.0 - mu);  // IPOPT tau_min = 0.99 [upstream]
    double dx[NX], dv[NU], wq[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        dx[i] = 0.0;
        wq[i] = 2.0 * c.Q[i];
    }
    MMPC_UNROLL for (int j = 0; j < NU; ++j) dv[j] = 0.0;
    o.gradf_d = o.lam_inf = o.f1 = o.theta1 = o.bar1 = 0.0;
    o.amax = o.az = 1.0;
    o.finite = true;
    double x[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) x[i] = c.ld(c.W.X, i);
    for (int k = 0; k < N; ++k) {
        double u[NU], v[NU], r[NX], xn[NX];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            r[i] = c.traj(k, i);
            xn[i] = c.ld(c.W.X, (k + 1) * NX + i);
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            u[j] = c.ld(c.W.U, k * NU + j);
            v[j] = k ? c.ld(c.W.U, (k - 1) * NU + j) : c.uinit[j];
        }
        double F[NX], Fx[NX * NX], Fu[NX * NU];
        c.template step<1>(x, u, nullptr, nullptr, nullptr, F, Fx, Fu, nullptr);
        double du[NU];
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            double a = c.ld(c.W.K0, k * NU + j);
            MMPC_UNROLL for (int i = 0; i < NX; ++i) a = fma(c.ld(c.W.KX, (k * NU + j) * NX + i), dx[i], a);
            MMPC_UNROLL for (int l = 0; l < NU; ++l) a = fma(c.ld(c.W.KV, (k * NU + j) * NU + l), dv[l], a);
            du[j] = a;
            c.st(c.W.DU, k * NU + j, a);
            o.finite = o.finite && isfinite(a);
        }
        double dxn[NX];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            double a = F[i] - xn[i];
            MMPC_UNROLL for (int j = 0; j < NX; ++j) a = fma(Fx[i * NX + j], dx[j], a);
            MMPC_UNROLL for (int j = 0; j < NU; ++j) a = fma(Fu[i * NU + j], du[j], a);
            dxn[i] = a;
            c.st(c.W.DX, (k + 1) * NX + i, a);
            // directional derivative of e'Qe: 2 Q e . (Fx dx + Fu du) = 2 Q e . (dxn - c)
            if (c.P.node_cost) o.gradf_d = fma(wq[i] * (xn[i] - r[i]), a, o.gradf_d);  // 2 Q (x_{k+1} - r_k) . dx_{k+1}
            else o.gradf_d = fma(wq[i] * (F[i] - r[i]), a - (F[i] - xn[i]), o.gradf_d);
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            o.gradf_d = fma(2.0 * c.R[j] * (u[j] - v[j]), du[j] - dv[j], o.gradf_d);
            o.gradf_d = fma(2.0 * c.Rm[j] * u[j], du[j], o.gradf_d);
        }
        if (BOUNDED) {  // multiplier steps, step-length limits, barrier gradient . d
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                double dzl, dzu;
                const double zl = c.ld(c.W.ZLU, k * NU + j), zu = c.ld(c.W.ZUU, k * NU + j);
                bound_step(u[j], du[j], c.P.u_min[j], c.P.u_max[j], zl, zu, mu, tau, dzl, dzu, o);
                c.st(c.W.DZLU, k * NU + j, dzl);
                c.st(c.W.DZUU, k * NU + j, dzu);
                if (c.P.u_min[j] > -MMPC_INF_THRESH) o.gradf_d -= mu / (u[j] - c.P.u_min[j]) * du[j];
                if (c.P.u_max[j] < MMPC_INF_THRESH) o.gradf_d += mu / (c.P.u_max[j] - u[j]) * du[j];
            }
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                double dzl, dzu;
                const double zl = c.ld(c.W.ZLX, k * NX + i), zu = c.ld(c.W.ZUX, k * NX + i);
                bound_step(xn[i], dxn[i], c.P.x_min[i], c.P.x_max[i], zl, zu, mu, tau, dzl, dzu, o);
                c.st(c.W.DZLX, k * NX + i, dzl);
                c.st(c.W.DZUX, k * NX + i, dzu);
                if (c.P.x_min[i] > -MMPC_INF_THRESH) o.gradf_d -= mu / (xn[i] - c.P.x_min[i]) * dxn[i];
                if (c.P.x_max[i] < MMPC_INF_THRESH) o.gradf_d += mu / (c.P.x_max[i] - xn[i]) * dxn[i];
            }
        }
        // multiplier of g_k from the value function of stage k+1
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            double a = c.ld(c.W.PX, k * NX + i);
            MMPC_UNROLL for (int j = 0; j < NX; ++j) a = fma(c.ld(c.W.PXX, (k * NX + i) * NX + j), dxn[j], a);
            MMPC_UNROLL for (int j = 0; j < NU; ++j) a = fma(c.ld(c.W.PXV, (k * NX + i) * NU + j), du[j], a);
            c.st(c.W.LAMN, k * NX + i, a);
            o.lam_inf = fmax(o.lam_inf, fabs(a));
        }
        // fused trial evaluation of the full step at node k
        {
            double xt[NX], ut[NU], Ft[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) xt[i] = x[i] + dx[i];
            MMPC_UNROLL for (int j = 0; j < NU; ++j) ut[j] = u[j] + du[j];
            c.template step<0>(xt, ut, nullptr, nullptr, nullptr, Ft, nullptr, nullptr, nullptr);
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                const double e = (c.P.node_cost ? xn[i] + dxn[i] : Ft[i]) - r[i];
                o.f1 = fma(c.Q[i] * e, e, o.f1);
                o.theta1 += fabs(Ft[i] - (xn[i] + dxn[i]));
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                const double d = ut[j] - (v[j] + dv[j]);
                o.f1 = fma(c.R[j] * d, d, o.f1);
                o.f1 = fma(c.Rm[j] * ut[j], ut[j], o.f1);
            }
        }
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            x[i] = xn[i];
            dx[i] = dxn[i];
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) dv[j] = du[j];
    }
}

// merit pieces at z + alpha dz
template <class M, bool BOUNDED>
__device__ void trial_eval(const SolveCtx<M>& c, double alpha, double& f1, double& theta1, double& bar1) {
    constexpr int NX = M::NX, NU = M::NU;
    const int N = c.P.N;
    f1 = 0.0; theta1 = 0.0; bar1 = 0.0;
    double x[NX], v[NU];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) x[i] = c.ld(c.W.X, i);
    MMPC_UNROLL for (int j = 0; j < NU; ++j) v[j] = c.uinit[j];
    for (int k = 0; k < N; ++k) {
        double u[NU], xn[NX], F[NX];
        MMPC_UNROLL for (int j = 0; j < NU; ++j) u[j] = fma(alpha, c.ld(c.W.DU, k * NU + j), c.ld(c.W.U, k * NU + j));
        MMPC_UNROLL for (int i = 0; i < NX; ++i) xn[i] = fma(alpha, c.ld(c.W.DX, (k + 1) * NX + i), c.ld(c.W.X, (k + 1) * NX + i));
        c.template step<0>(x, u, nullptr, nullptr, nullptr, F, nullptr, nullptr, nullptr);
        MMPC_UNROLL for (int i = 0; i < NX; ++i) {
            const double e = (c.P.node_cost ? xn[i] : F[i]) - c.traj(k, i);
            f1 = fma(c.Q[i] * e, e, f1);
            theta1 += fabs(F[i] - xn[i]);
            x[i] = xn[i];
            if (BOUNDED) bar1 += barrier_of(xn[i], c.P.x_min[i], c.P.x_max[i]);
        }
        MMPC_UNROLL for (int j = 0; j < NU; ++j) {
            const double d = u[j] - v[j];
            f1 = fma(c.R[j] * d, d, f1);
            f1 = fma(c.Rm[j] * u[j], u[j], f1);
            v[j] = u[j];
            if (BOUNDED) bar1 += barrier_of(u[j], c.P.u_min[j], c.P.u_max[j]);
        }
    }
}

// IPOPT's kappa_sigma safeguard on a bound multiplier after the step (eq. 16, kappa_sigma = 1e10) [upstream]
__device__ __forceinline__ double z_safeguard(double z, double s, double mu) { return fmin(fmax(z, mu / (1e10 * s)), 1e10 * mu / s); }

template <class M, bool BOUNDED>
__global__ void __launch_bounds__(32) solve_kernel_t1(mmpc_problem P, mmpc_workspace W, mmpc_solve_io io) {
    constexpr int NX = M::NX, NU = M::NU, S = NX + NU;
    const int b = io.b0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.b1) return;
    const int N = P.N;
    SolveCtx<M> c(P, W, io, b);
    MMPC_UNROLL for (int i = 0; i < NX; ++i) c.Q[i] = c.ld(W.Qw, i);
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        c.R[j] = c.ld(W.Rw, j);
        c.Rm[j] = c.ld(W.Rmw, j);
        c.uinit[j] = io.control[(size_t)b * NU + j];
    }
    // ---- start point: warm = previous solution of this slot (ModelControl.cpp:161), else zeros (:28-50) ----
    const bool warm = io.warm_start && W.have_warm;
    if (!warm) {
        for (int r = NX; r < (N + 1) * NX; ++r) c.st(W.X, r, 0.0);
        for (int r = 0; r < N * NU; ++r) c.st(W.U, r, 0.0);
    }
    if (!(warm && io.warm_start_duals))
        for (int r = 0; r < N * NX; ++r) c.st(W.LAM, r, 0.0);
    double x0[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        x0[i] = io.state[(size_t)b * NX + i];
        c.st(W.X, i, x0[i]);  // x_0 pinned (ModelControl.cpp:144-145)
    }
    if (P.is_linear) {  // A, B, xdot at (state, control) (ModelControl.cpp:125-135), kept row-major
        double xd[NX], A[NX * NX], Bm[NX * NU];
        M::fAB(x0, c.uinit, xd, A, Bm);
        MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) c.st(W.LIN, i, A[i]);
        MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) c.st(W.LIN, NX * NX + i, Bm[i]);
        MMPC_UNROLL for (int i = 0; i < NX; ++i) c.st(W.LIN, NX * NX + NX * NU + i, xd[i]);
    }
    double mu = 0.0;
    int nbounded = 0;
    if (BOUNDED) {
        // push the start into the interior (IPOPT bound_push = bound_frac = 1e-2) and set z = 1 [upstream]
        mu = io.mu_init;
        for (int k = 0; k < N; ++k) {
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                const double lb = P.u_min[j], ub = P.u_max[j];
                const bool hl = lb > -MMPC_INF_THRESH, hu = ub < MMPC_INF_THRESH;
                double v = c.ld(W.U, k * NU + j);
                if (hl) { double p = 1e-2 * fmax(1.0, fabs(lb)); if (hu) p = fmin(p, 1e-2 * (ub - lb)); v = fmax(v, lb + p); }
                if (hu) { double p = 1e-2 * fmax(1.0, fabs(ub)); if (hl) p = fmin(p, 1e-2 * (ub - lb)); v = fmin(v, ub - p); }
                c.st(W.U, k * NU + j, v);
                c.st(W.ZLU, k * NU + j, hl ? 1.0 : 0.0);
                c.st(W.ZUU, k * NU + j, hu ? 1.0 : 0.0);
                nbounded += (hl || hu);
            }
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {  // x_{k+1}
                const double lb = P.x_min[i], ub = P.x_max[i];
                const bool hl = lb > -MMPC_INF_THRESH, hu = ub < MMPC_INF_THRESH;
                double v = c.ld(W.X, (k + 1) * NX + i);
                if (hl) { double p = 1e-2 * fmax(1.0, fabs(lb)); if (hu) p = fmin(p, 1e-2 * (ub - lb)); v = fmax(v, lb + p); }
                if (hu) { double p = 1e-2 * fmax(1.0, fabs(ub)); if (hl) p = fmin(p, 1e-2 * (ub - lb)); v = fmin(v, ub - p); }
                c.st(W.X, (k + 1) * NX + i, v);
                c.st(W.ZLX, k * NX + i, hl ? 1.0 : 0.0);
                c.st(W.ZUX, k * NX + i, hu ? 1.0 : 0.0);
                nbounded += (hl || hu);
            }
        }
    }
    double nu_merit = 1.0, delta_last = 0.0, err = INFINITY, fval = 0.0;
    int it = 0, status = MMPC_SOLVE_MAX_ITER, ls_total = 0, reg_total = 0;
    const int nfree = N * S, ncon = N * NX;  // free variables: u_0, (x_k,u_k) k=1..N-1, x_N
    for (it = 0;; ++it) {
        SweepNorms nrm;
        double delta = 0.0;
        bool ok = true;
        if (BOUNDED) backward_sweep<M, BOUNDED, false>(c, 0.0, mu, nrm);  // residuals first: mu is chosen from them
        else ok = backward_sweep<M, BOUNDED, true>(c, delta, mu, nrm);
        // IPOPT scaled optimality error (Waechter & Biegler 2006 eq. 5-6) [upstream]: s_max = 100
        const double sd = fmax(100.0, (nrm.lam_l1 + nrm.z_l1) / (double)(nfree + ncon)) / 100.0;
        const double sc = fmax(100.0, nrm.z_l1 / (double)nfree) / 100.0;
        const double e_dp = fmax(nrm.dual_inf / sd, nrm.prim_inf);
        err = BOUNDED ? fmax(e_dp, nrm.cmax / sc) : e_dp;
        fval = nrm.fval;
        if (!isfinite(err)) { status = MMPC_SOLVE_NUMERICAL; break; }
        if (err <= io.tol) { status = MMPC_SOLVE_CONVERGED; break; }
        if (it >= io.max_iter) { status = MMPC_SOLVE_MAX_ITER; break; }
        if (BOUNDED) {
            // monotone barrier update (kappa_eps = 10, kappa_mu = 0.2, theta_mu = 1.5) [upstream]
            while (mu > io.tol / 10.0 && fmax(e_dp, fmax(nrm.cmax - mu, mu - nrm.cmin) / sc) <= 10.0 * mu)
                mu = fmax(io.tol / 10.0, fmin(0.2 * mu, mu * sqrt(mu)));
            SweepNorms nb;
            ok = backward_sweep<M, BOUNDED, true>(c, delta, mu, nb);
            nrm.barrier = nb.barrier;
        }
        // inertia correction: retry the sweep with delta on the diagonal until every G_uu is PD
        int tries = 0;
        while (!ok && tries < 24) {
            delta = (delta == 0.0) ? (delta_last > 0.0 ? fmax(1e-20, delta_last / 3.0) : 1e-4) : delta * 8.0;
            SweepNorms dummy;
            ok = backward_sweep<M, BOUNDED, true>(c, delta, mu, dummy);
            ++tries; ++reg_total;
        }
        if (!ok) { status = MMPC_SOLVE_NUMERICAL; break; }
        if (delta > 0.0) delta_last = delta;
        FwdOut fo;
        forward_sweep<M, BOUNDED>(c, mu, fo);
        if (!fo.finite) { status = MMPC_SOLVE_NUMERICAL; break; }
        nu_merit = fmax(nu_merit, 1.1 * fo.lam_inf + 1e-8);
        const double m0 = nrm.fval + mu * nrm.barrier + nu_merit * nrm.theta;
        const double D = fo.gradf_d - nu_merit * nrm.theta;
        double alpha = BOUNDED ? fo.amax : 1.0;
        double f1 = fo.f1, th1 = fo.theta1, bar1 = fo.bar1;
        if (BOUNDED && alpha < 1.0) trial_eval<M, BOUNDED>(c, alpha, f1, th1, bar1);
        for (int ls = 0; ls < 40; ++ls) {
            if (f1 + mu * bar1 + nu_merit * th1 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0)) break;
            alpha *= 0.5;
            ++ls_total;
            trial_eval<M, BOUNDED>(c, alpha, f1, th1, bar1);
        }
        for (int r = NX; r < (N + 1) * NX; ++r) c.st(W.X, r, fma(alpha, c.ld(W.DX, r), c.ld(W.X, r)));
        for (int r = 0; r < N * NU; ++r) c.st(W.U, r, fma(alpha, c.ld(W.DU, r), c.ld(W.U, r)));
        for (int r = 0; r < N * NX; ++r) {
            const double l = c.ld(W.LAM, r);
            c.st(W.LAM, r, fma(alpha, c.ld(W.LAMN, r) - l, l));
        }
        if (BOUNDED) {
            for (int k = 0; k < N; ++k) {
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    const int r = k * NU + j;
                    const double v = c.ld(W.U, r);
                    if (P.u_min[j] > -MMPC_INF_THRESH) c.st(W.ZLU, r, z_safeguard(fma(fo.az, c.ld(W.DZLU, r), c.ld(W.ZLU, r)), v - P.u_min[j], mu));
                    if (P.u_max[j] < MMPC_INF_THRESH) c.st(W.ZUU, r, z_safeguard(fma(fo.az, c.ld(W.DZUU, r), c.ld(W.ZUU, r)), P.u_max[j] - v, mu));
                }
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    const int r = k * NX + i;
                    const double v = c.ld(W.X, (k + 1) * NX + i);
                    if (P.x_min[i] > -MMPC_INF_THRESH) c.st(W.ZLX, r, z_safeguard(fma(fo.az, c.ld(W.DZLX, r), c.ld(W.ZLX, r)), v - P.x_min[i], mu));
                    if (P.x_max[i] < MMPC_INF_THRESH) c.st(W.ZUX, r, z_safeguard(fma(fo.az, c.ld(W.DZUX, r), c.ld(W.ZUX, r)), P.x_max[i] - v, mu));
                }
            }
        }
    }
    // ---- outputs in the reference's V layout (ModelGenerator.cpp:76-112; unpacked ModelControl.cpp:174-190) ----
    const int NV = NX * (N + 1) + NU * N;
    double* Vb = io.V_out + (size_t)b * NV;
    for (int k = 0; k <= N; ++k) {
        MMPC_UNROLL for (int i = 0; i < NX; ++i) Vb[k * S + i] = c.ld(W.X, k * NX + i);
        if (k < N) MMPC_UNROLL for (int j = 0; j < NU; ++j) Vb[k * S + NX + j] = c.ld(W.U, k * NU + j);
    }
    if (io.lam_out)
        for (int r = 0; r < N * NX; ++r) io.lam_out[(size_t)b * N * NX + r] = c.ld(W.LAM, r);
    if (io.stats) {
        mmpc_stats s;
        s.objective = fval; s.kkt_error = err; s.mu = mu; s.iterations = it; s.status = status;
        s.ls_steps = ls_total; s.reg_steps = reg_total;
        io.stats[b] = s;
    }
    (void)nbounded;
}

}  // namespace mmpc

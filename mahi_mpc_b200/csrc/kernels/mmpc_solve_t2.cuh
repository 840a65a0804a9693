// mahi-mpc-b200 :: solve kernel, variant T2 -- G lanes of a warp cooperate on one MPC instance.
//
// Why: at the benchmark batch (4096 instances) a thread-per-instance kernel (T1) keeps one warp per
// SM busy and is bound by instruction latency (ncu: fp64 pipe 3.5 % active, profiles/r1_solve_t1_*).
// T2 spreads every instance over G = 8 (nx+2nu <= 8) or 16 lanes, so the same batch runs on G times
// more warps AND the serial chain per stage gets shorter:
//
//   parallel phases   lane l of a group evaluates shooting node k0+l of a chunk of G consecutive nodes
//                     (F, dF/dx, dF/du, Lagrangian stage Hessian, residual norms; later the merit terms
//                     of the trial step) and leaves a record in shared memory;
//   serial phases     the G lanes run the Riccati recursion over the chunk's stages together: lane a owns
//                     row a of the (nx+2nu)^2 stage matrix M = H~ + Z' P~ Z over w = (x_k, u_{k-1}, u_k),
//                     rows of P~ live in registers, the few shared operands (T = P~ Z, the u-rows of M,
//                     the gains) go through shared memory with __syncwarp() in between; the forward
//                     recursion shares dx/du with width-G shuffles.
//
// Everything is warp-local: no __syncthreads, no grid sync, one launch per batched solve; groups of a
// warp that converge early keep executing in lock-step but stop updating their iterate.
// Replaces m_solver(m_solver_args) = IPOPT + MUMPS (src/Mahi/Mpc/ModelControl.cpp:159); the NLP is the
// one of ModelGenerator.cpp:23-233.
#pragma once

namespace mmpc {

template <class M>
struct T2 {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, N2 = NX + 2 * NU, NH = S * (S + 1) / 2;
    static constexpr int G = N2 <= 8 ? 8 : (N2 <= 16 ? 16 : 32);
    static_assert(N2 <= 32, "nx + 2 nu must fit one warp");
    static constexpr int GPW = 32 / G;  // instances per warp
    // backward record (shared memory, per node): Fx | Fu | H (packed, full stage Hessian) | c | q~ (x,v,u)
    static constexpr int RB_FX = 0, RB_FU = RB_FX + NX * NX, RB_H = RB_FU + NX * NU, RB_C = RB_H + NH, RB_Q = RB_C + NX,
                         RB_RAW = RB_Q + N2, RB = RB_RAW | 1;  // odd stride: lanes hit distinct banks
    // global per-node records
    static constexpr int FG_FX = 0, FG_FU = NX * NX, FG_C = FG_FU + NX * NU, FG_Q = FG_C + NX, FG = FG_Q + N2;
    static constexpr int KG = NU * S + NU;      // K stored [a][l] (a < S, l < NU), then k0[l]
    static constexpr int PG = NX * (S + 1);     // rows 0..NX-1 of P~ (S entries) + p~
    // forward record (shared memory, per node): FG | KG | PG
    static constexpr int RF_K = FG, RF_P = FG + KG, RF_RAW = FG + KG + PG, RF = RF_RAW | 1;
    static constexpr int REC = RB > RF ? RB : RF;
    // exchange area per group
    static constexpr int X_T = 0, X_Y = X_T + S * S, X_MU = X_Y + S, X_K = X_MU + NU * (S + 1), X_END = X_K + NU * (S + 1);
    static constexpr int GS = (G * REC + X_END) | 1;  // doubles per group, odd
    static constexpr int LIN = NX * NX + NX * NU + NX;
};

// 8-byte asynchronous global->shared copy (LDGSTS): many can be in flight per lane, no register staging
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double shfl_g(double v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
template <int G>
__device__ __forceinline__ double group_sum(double v) {
    MMPC_UNROLL for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v) {
    MMPC_UNROLL for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o, G));
    return v;
}

template <class M>
struct T2Ctx {
    using C = T2<M>;
    static constexpr int NX = M::NX, NU = M::NU, S = C::S, N2 = C::N2, G = C::G;
    const mmpc_problem& P;
    int N, gl;             // nodes; lane in group
    double* gs;            // group shared memory
    // instance-major global arrays of this instance
    double *V, *LAM, *DV, *LAMN, *FGp, *KGp, *PGp, *LINp;
    const double* traj;
    double Q[NX], R[NU], Rm[NU], uinit[NU];

    __device__ T2Ctx(const mmpc_problem& P_) : P(P_) {}

    template <int WANT>
    __device__ __forceinline__ void step(const double* x, const double* u, const double* wq, const double* r, const double* lam,
                                         double* F, double* Fx, double* Fu, double* Hs) const {
        if (P.is_linear) {
            double x0[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) x0[i] = V[i];  // x_init = pinned state (ModelControl.cpp:134)
            step_linear<M, WANT>(P.h, x, u, LINp, LINp + NX * NX, LINp + NX * NX + NX * NU, x0, uinit, F, Fx, Fu, Hs);
        } else {
            step_euler<M, WANT>(P.h, x, u, wq, r, lam, F, Fx, Fu, Hs);
        }
    }
};

struct T2Norms {
    double dual_inf, prim_inf, theta, fval, lam_l1;
};

// ---- backward sweep: parallel node evaluation + cooperative Riccati -------------------------------
// returns (group-uniform) whether every G_uu was positive definite
template <class M>
__device__ __noinline__ bool t2_backward(const T2Ctx<M>& cin, double delta, T2Norms& nrm) {
    const T2Ctx<M> c = cin;  // callee-local copy: keeps the context in registers instead of the caller's stack
    using C = T2<M>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, N2 = C::N2, NH = C::NH, G = C::G;
    const int N = c.N, a = c.gl;
    double* X = c.gs + G * C::REC;  // exchange area
    double Prow[S], prow = 0.0;
    MMPC_UNROLL for (int j = 0; j < S; ++j) Prow[j] = 0.0;
    double wq[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) wq[i] = 2.0 * c.Q[i];
    nrm.dual_inf = nrm.prim_inf = nrm.theta = nrm.fval = nrm.lam_l1 = 0.0;
    bool ok = true;
    const int nch = (N + G - 1) / G;
    for (int ch = nch - 1; ch >= 0; --ch) {
        const int k0 = ch * G, kn = min(G, N - k0);
        // ================= parallel phase: lane a evaluates node k0 + a =================
        if (a < kn) {
            const int k = k0 + a;
            const double* z = c.V + (size_t)k * S;
            double x[NX], u[NU], v[NU], r[NX], lam[NX], xn[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                x[i] = z[i];
                xn[i] = z[S + i];
                r[i] = c.traj[(size_t)k * NX + i];
                lam[i] = c.LAM[(size_t)k * NX + i];
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                u[j] = z[NX + j];
                v[j] = k ? z[NX + j - S] : c.uinit[j];
            }
            double F[NX], Fx[NX * NX], Fu[NX * NU], H[NH];
            c.template step<2>(x, u, wq, r, lam, F, Fx, Fu, H);
            double* rec = c.gs + a * C::REC;
            double* fg = c.FGp + (size_t)k * C::FG;
            double qe[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                const double e = F[i] - r[i], cr = F[i] - xn[i];
                qe[i] = wq[i] * e;
                nrm.fval = fma(c.Q[i] * e, e, nrm.fval);
                nrm.prim_inf = fmax(nrm.prim_inf, fabs(cr));
                nrm.theta += fabs(cr);
                nrm.lam_l1 += fabs(lam[i]);
                rec[C::RB_C + i] = cr;
                fg[C::FG_C + i] = cr;
                if (k == N - 1) nrm.dual_inf = fmax(nrm.dual_inf, fabs(lam[i]));  // stationarity w.r.t. x_N is -lam_{N-1}
            }
            MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) { rec[C::RB_FX + i] = Fx[i]; fg[C::FG_FX + i] = Fx[i]; }
            MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) { rec[C::RB_FU + i] = Fu[i]; fg[C::FG_FU + i] = Fu[i]; }
            // objective gradient pieces q~ = (qx, qv, qu) and the stationarity residual of the current iterate
            MMPC_UNROLL for (int j = 0; j < NX; ++j) {
                double g = 0.0, l = 0.0;
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    g = fma(Fx[i * NX + j], qe[i], g);
                    l = fma(Fx[i * NX + j], lam[i], l);
                }
                rec[C::RB_Q + j] = g;
                fg[C::FG_Q + j] = g;
                if (k) nrm.dual_inf = fmax(nrm.dual_inf, fabs(g + l - c.LAM[(size_t)(k - 1) * NX + j]));
            }
            MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                double g = 0.0, l = 0.0;
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    g = fma(Fu[i * NU + j], qe[i], g);
                    l = fma(Fu[i * NU + j], lam[i], l);
                }
                const double du = u[j] - v[j], dRu = 2.0 * c.R[j] * du;
                const double qu = g + dRu + 2.0 * c.Rm[j] * u[j];
                const double dRn = (k < N - 1) ? 2.0 * c.R[j] * (z[S + NX + j] - u[j]) : 0.0;  // 2R (u_{k+1} - u_k)
                nrm.fval = fma(c.R[j] * du, du, nrm.fval);
                nrm.fval = fma(c.Rm[j] * u[j], u[j], nrm.fval);
                nrm.dual_inf = fmax(nrm.dual_inf, fabs(qu + l - dRn));
                rec[C::RB_Q + NX + j] = -dRu;  fg[C::FG_Q + NX + j] = -dRu;   // qv
                rec[C::RB_Q + S + j] = qu;     fg[C::FG_Q + S + j] = qu;      // qu
            }
            // full stage Hessian over (x,u): H += Fz' (2Q) Fz ; uu diagonal += 2(R + Rm); + delta I
            MMPC_UNROLL for (int p = 0; p < S; ++p) MMPC_UNROLL for (int d = p; d < S; ++d) {
                double acc = H[tri(p, d, S)];
                MMPC_UNROLL for (int i = 0; i < NX; ++i) {
                    const double fa = p < NX ? Fx[i * NX + p] : Fu[i * NU + p - NX];
                    const double fd = d < NX ? Fx[i * NX + d] : Fu[i * NU + d - NX];
                    acc = fma(wq[i] * fa, fd, acc);
                }
                if (p == d) acc += delta + (p >= NX ? 2.0 * (c.R[p - NX] + c.Rm[p - NX]) : 0.0);
                rec[C::RB_H + tri(p, d, S)] = acc;
            }
        }
        __syncwarp();
        // ================= serial phase: Riccati over the chunk's stages, G lanes per instance =================
        for (int s = kn - 1; s >= 0; --s) {
            const int k = k0 + s;
            const double* rec = c.gs + s * C::REC;
            const double *Fx = rec + C::RB_FX, *Fu = rec + C::RB_FU, *Hp = rec + C::RB_H, *cc = rec + C::RB_C, *qq = rec + C::RB_Q;
            // (1) lane a < S: row a of T = P~ Z (columns x,u) and y_a = (P~ c~ + p~)_a
            if (a < S) {
                double y = prow;
                MMPC_UNROLL for (int l = 0; l < NX; ++l) y = fma(Prow[l], cc[l], y);
                X[C::X_Y + a] = y;
                MMPC_UNROLL for (int j = 0; j < NX; ++j) {
                    double t = 0.0;
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(Prow[l], Fx[l * NX + j], t);
                    X[C::X_T + a * S + j] = t;
                }
                MMPC_UNROLL for (int j = 0; j < NU; ++j) {
                    double t = Prow[NX + j];
                    MMPC_UNROLL for (int l = 0; l < NX; ++l) t = fma(Prow[l], Fu[l * NU + j], t);
                    X[C::X_T + a * S + NX + j] = t;
                }
                // the value function of stage k+1 is what the forward sweep needs for lam_k
                if (a < NX) {
                    double* pg = c.PGp + (size_t)k * C::PG + a * (S + 1);
                    MMPC_UNROLL for (int j = 0; j < S; ++j) pg[j] = Prow[j];
                    pg[S] = prow;
                }
            }
            __syncwarp();
            // (2) lane a < N2: row a of M over the (x,u) columns, its v-columns, and m_a
            double Mxu[S], Mv[NU], ma = 0.0;
            {
                double zc[S];  // column a of Z = [[Fx 0 Fu],[0 0 I]]
                MMPC_UNROLL for (int j = 0; j < S; ++j) { Mxu[j] = 0.0; zc[j] = 0.0; }
                MMPC_UNROLL for (int j = 0; j < NU; ++j) Mv[j] = 0.0;
                if (a < NX) {
                    MMPC_UNROLL for (int i = 0; i < NX; ++i) zc[i] = Fx[i * NX + a];
                    MMPC_UNROLL for (int j = 0; j < S; ++j) Mxu[j] = Hp[a <= j ? tri(a, j, S) : tri(j, a, S)];
                    ma = qq[a];
                } else if (a < S) {  // v row: only the control-increment weight
                    ma = qq[a];
                    MMPC_UNROLL for (int j = 0; j < NU; ++j) if (a - NX == j) {
                        Mv[j] = 2.0 * c.R[j];        // M[v_j][v_j]
                        Mxu[NX + j] = -2.0 * c.R[j];  // M[v_j][u_j]
                    }
                } else if (a < N2) {  // u row
                    const int ju = a - S;
                    MMPC_UNROLL for (int i = 0; i < NX; ++i) zc[i] = Fu[i * NU + ju];
                    MMPC_UNROLL for (int j = 0; j < NU; ++j) if (j == ju) { zc[NX + j] = 1.0; Mv[j] = -2.0 * c.R[j]; }
                    const int au = NX + ju;
                    MMPC_UNROLL for (int j = 0; j < S; ++j) Mxu[j] = Hp[au <= j ? tri(au, j, S) : tri(j, au, S)];
                    ma = qq[a];
                }
                MMPC_UNROLL for (int i = 0; i < S; ++i) {
                    ma = fma(zc[i], X[C::X_Y + i], ma);
                    MMPC_UNROLL for (int j = 0; j < S; ++j) Mxu[j] = fma(zc[i], X[C::X_T + i * S + j], Mxu[j]);
                }
            }
            // (3) the u lanes publish their rows
            if (a >= S && a < N2) {
                double* mu = X + C::X_MU + (a - S) * (S + 1);
                MMPC_UNROLL for (int j = 0; j < S; ++j) mu[j] = Mxu[j];
                mu[S] = ma;
            }
            __syncwarp();
            // (4) every lane factors M_uu; lane a < S solves for column a of K, lane S for k0
            double L[NU * NU];
            MMPC_UNROLL for (int i = 0; i < NU; ++i) MMPC_UNROLL for (int j = 0; j < NU; ++j)
                L[i * NU + j] = 0.5 * (X[C::X_MU + i * (S + 1) + NX + j] + X[C::X_MU + j * (S + 1) + NX + i]);
            ok = chol<NU>(L) && ok;
            {
                double col[NU];
                MMPC_UNROLL for (int l = 0; l < NU; ++l) {
                    double vv = 0.0;
                    if (a < NX) vv = X[C::X_MU + l * (S + 1) + a];          // M[u_l][x_a]
                    else if (a < S) vv = (a - NX == l) ? -2.0 * c.R[l] : 0.0;  // M[u_l][v_a]
                    else vv = X[C::X_MU + l * (S + 1) + S];                  // m_u (used by lane S)
                    col[l] = -vv;
                }
                chol_solve<NU>(L, col);
                if (a <= S) MMPC_UNROLL for (int l = 0; l < NU; ++l) X[C::X_K + l * (S + 1) + a] = col[l];
                // gains for the forward sweep: K[a][l] (a < S), k0[l]
                if (a <= S) {
                    double* kg = c.KGp + (size_t)k * C::KG + a * NU;
                    MMPC_UNROLL for (int l = 0; l < NU; ++l) kg[l] = col[l];
                }
            }
            __syncwarp();
            // (5) lane a < S: new row of P~ and p~_a
            if (a < S) {
                double Mau[NU];
                MMPC_UNROLL for (int l = 0; l < NU; ++l) Mau[l] = Mxu[NX + l];
                MMPC_UNROLL for (int j = 0; j < S; ++j) {
                    double t = j < NX ? Mxu[j] : Mv[j - NX];
                    MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Mau[l], X[C::X_K + l * (S + 1) + j], t);
                    Prow[j] = t;
                }
                double t = ma;
                MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Mau[l], X[C::X_K + l * (S + 1) + S], t);
                prow = t;
            }
            __syncwarp();
        }
    }
    // group reductions of the residual norms
    nrm.dual_inf = group_max<G>(nrm.dual_inf);
    nrm.prim_inf = group_max<G>(nrm.prim_inf);
    nrm.theta = group_sum<G>(nrm.theta);
    nrm.fval = group_sum<G>(nrm.fval);
    nrm.lam_l1 = group_sum<G>(nrm.lam_l1);
    return ok;
}

// merit pieces of node k at z + alpha dz (lane-local contribution)
template <class M>
__device__ __forceinline__ void t2_trial_node(const T2Ctx<M>& c, int k, double alpha, double& f1, double& th1) {
    using C = T2<M>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S;
    const double* z = c.V + (size_t)k * S;
    const double* d = c.DV + (size_t)k * S;
    double x[NX], u[NU], v[NU], xn[NX], F[NX];
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        x[i] = fma(alpha, d[i], z[i]);
        xn[i] = fma(alpha, d[S + i], z[S + i]);
    }
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        u[j] = fma(alpha, d[NX + j], z[NX + j]);
        v[j] = k ? fma(alpha, d[NX + j - S], z[NX + j - S]) : c.uinit[j];
    }
    c.template step<0>(x, u, nullptr, nullptr, nullptr, F, nullptr, nullptr, nullptr);
    MMPC_UNROLL for (int i = 0; i < NX; ++i) {
        const double e = F[i] - c.traj[(size_t)k * NX + i];
        f1 = fma(c.Q[i] * e, e, f1);
        th1 += fabs(F[i] - xn[i]);
    }
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        const double dd = u[j] - v[j];
        f1 = fma(c.R[j] * dd, dd, f1);
        f1 = fma(c.Rm[j] * u[j], u[j], f1);
    }
}

// ---- forward sweep: step, new multipliers, directional derivative, merit of the full step ----------
template <class M>
__device__ __noinline__ void t2_forward(const T2Ctx<M>& cin, double& gradf_d, double& lam_inf, double& f1, double& th1, bool& finite) {
    const T2Ctx<M> c = cin;
    using C = T2<M>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, N2 = C::N2, G = C::G;
    const int N = c.N, a = c.gl;
    double dxt[S];  // (dx_k, du_{k-1}), replicated in every lane of the group
    MMPC_UNROLL for (int j = 0; j < S; ++j) dxt[j] = 0.0;
    gradf_d = 0.0; lam_inf = 0.0; f1 = 0.0; th1 = 0.0;
    double bad = 0.0;
    const int nch = (N + G - 1) / G;
    for (int ch = 0; ch < nch; ++ch) {
        const int k0 = ch * G, kn = min(G, N - k0);
        // staging: the chunk's records (Fx,Fu,c,q~ | gains | value-function rows) are contiguous per instance in
        // global memory; the G lanes copy them into shared memory together with asynchronous 8-byte copies
        {
            const double* fg0 = c.FGp + (size_t)k0 * C::FG;
            const double* kg0 = c.KGp + (size_t)k0 * C::KG;
            const double* pg0 = c.PGp + (size_t)k0 * C::PG;
            for (int s = 0; s < kn; ++s) {
                double* rec = c.gs + s * C::REC;
                MMPC_UNROLL for (int i = 0; i < (C::FG + G - 1) / G; ++i) if (i * G + a < C::FG) cp_async8(rec + i * G + a, fg0 + s * C::FG + i * G + a);
                MMPC_UNROLL for (int i = 0; i < (C::KG + G - 1) / G; ++i) if (i * G + a < C::KG) cp_async8(rec + C::RF_K + i * G + a, kg0 + s * C::KG + i * G + a);
                MMPC_UNROLL for (int i = 0; i < (C::PG + G - 1) / G; ++i) if (i * G + a < C::PG) cp_async8(rec + C::RF_P + i * G + a, pg0 + s * C::PG + i * G + a);
            }
            cp_async_wait_all();
        }
        __syncwarp();
        // serial phase
        for (int s = 0; s < kn; ++s) {
            const int k = k0 + s;
            const double* rec = c.gs + s * C::REC;
            const double *Fx = rec + C::FG_FX, *Fu = rec + C::FG_FU, *cc = rec + C::FG_C, *qq = rec + C::FG_Q;
            const double *K = rec + C::RF_K, *Pr = rec + C::RF_P;
            // du_l on lane l < NU
            double mine = 0.0;
            if (a < NU) {
                mine = K[S * NU + a];  // k0
                MMPC_UNROLL for (int j = 0; j < S; ++j) mine = fma(K[j * NU + a], dxt[j], mine);
            }
            double du[NU];
            MMPC_UNROLL for (int l = 0; l < NU; ++l) du[l] = shfl_g(mine, l, G);
            // directional derivative of the objective: q~ . (dx, dv, du)
            if (a == 0) {
                MMPC_UNROLL for (int j = 0; j < S; ++j) gradf_d = fma(qq[j], dxt[j], gradf_d);
                MMPC_UNROLL for (int l = 0; l < NU; ++l) gradf_d = fma(qq[S + l], du[l], gradf_d);
            }
            // dx_{k+1} on lane i < NX
            mine = 0.0;
            if (a < NX) {
                mine = cc[a];
                MMPC_UNROLL for (int j = 0; j < NX; ++j) mine = fma(Fx[a * NX + j], dxt[j], mine);
                MMPC_UNROLL for (int l = 0; l < NU; ++l) mine = fma(Fu[a * NU + l], du[l], mine);
            }
            MMPC_UNROLL for (int i = 0; i < NX; ++i) dxt[i] = shfl_g(mine, i, G);
            MMPC_UNROLL for (int l = 0; l < NU; ++l) dxt[NX + l] = du[l];
            // lam_k = (P~_{k+1} (dx_{k+1}, du_k) + p~_{k+1})_x on lane i < NX ; write the step
            if (a < NX) {
                double lm = Pr[a * (S + 1) + S];
                MMPC_UNROLL for (int j = 0; j < S; ++j) lm = fma(Pr[a * (S + 1) + j], dxt[j], lm);
                c.LAMN[(size_t)k * NX + a] = lm;
                lam_inf = fmax(lam_inf, fabs(lm));
                c.DV[(size_t)(k + 1) * S + a] = mine;
                bad += (isfinite(mine) && isfinite(lm)) ? 0.0 : 1.0;
            }
            if (a < NU) c.DV[(size_t)k * S + NX + a] = du[a];
        }
        __syncwarp();
        // parallel phase: merit terms of the full step at the chunk's nodes
        if (a < kn) t2_trial_node<M>(c, k0 + a, 1.0, f1, th1);
    }
    gradf_d = shfl_g(gradf_d, 0, G);
    lam_inf = group_max<G>(lam_inf);
    f1 = group_sum<G>(f1);
    th1 = group_sum<G>(th1);
    finite = group_sum<G>(bad) == 0.0 && isfinite(f1);
}

template <class M>
__device__ __noinline__ void t2_trial(const T2Ctx<M>& cin, double alpha, double& f1, double& th1) {
    const T2Ctx<M> c = cin;
    constexpr int G = T2<M>::G;
    f1 = 0.0; th1 = 0.0;
    for (int k = c.gl; k < c.N; k += G) t2_trial_node<M>(c, k, alpha, f1, th1);
    f1 = group_sum<G>(f1);
    th1 = group_sum<G>(th1);
}

template <class M, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) solve_kernel_t2(mmpc_problem P, mmpc_workspace W, mmpc_solve_io io) {
    using C = T2<M>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, G = C::G, GPW = C::GPW;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int gl = lane % G, gw = lane / G;
    const int N = P.N, NV = NX * (N + 1) + NU * N, Mc = N * NX;
    const int inst_raw = io.b0 + (blockIdx.x * WARPS + wic) * GPW + gw;
    const bool valid = inst_raw < io.b1;
    const int b = inst_raw;                          // workspace slot (arrays are sized for the padded batch)
    const int bi = valid ? inst_raw : io.b1 - 1;     // surplus groups read the last instance's inputs, never write outputs

    T2Ctx<M> c(P);
    c.N = N; c.gl = gl;
    c.gs = smem + (size_t)(wic * GPW + gw) * C::GS;
    c.V = W.V2 + (size_t)b * NV; c.LAM = W.LAM2 + (size_t)b * Mc;
    c.DV = W.DV2 + (size_t)b * NV; c.LAMN = W.LAMN2 + (size_t)b * Mc;
    c.FGp = W.FG2 + (size_t)b * N * C::FG; c.KGp = W.KG2 + (size_t)b * N * C::KG; c.PGp = W.PG2 + (size_t)b * N * C::PG;
    c.LINp = W.LIN2 + (size_t)b * C::LIN;
    c.traj = io.traj + (size_t)bi * N * NX;
    MMPC_UNROLL for (int i = 0; i < NX; ++i) c.Q[i] = W.Qw[(size_t)i * W.Bp + bi];
    MMPC_UNROLL for (int j = 0; j < NU; ++j) {
        c.R[j] = W.Rw[(size_t)j * W.Bp + bi];
        c.Rm[j] = W.Rmw[(size_t)j * W.Bp + bi];
        c.uinit[j] = io.control[(size_t)bi * NU + j];
    }
    // ---- start point (ModelControl.cpp:28-50,68 cold; :161 warm), x_0 pinned (:144-145) ----
    const bool warm = io.warm_start && W.have_warm;
    {
        if (!warm) for (int e = NX + gl; e < NV; e += G) c.V[e] = 0.0;
        if (!(warm && io.warm_start_duals)) for (int e = gl; e < Mc; e += G) c.LAM[e] = 0.0;
        if (gl < NX) {
            c.V[gl] = io.state[(size_t)bi * NX + gl];
            c.DV[gl] = 0.0;
        }
    }
    __syncwarp();
    if (P.is_linear && gl == 0) {  // A, B, xdot at (state, control) (ModelControl.cpp:125-135), row-major
        double x0[NX], xd[NX], A[NX * NX], Bm[NX * NU];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) x0[i] = c.V[i];
        M::fAB(x0, c.uinit, xd, A, Bm);
        MMPC_UNROLL for (int i = 0; i < NX * NX; ++i) c.LINp[i] = A[i];
        MMPC_UNROLL for (int i = 0; i < NX * NU; ++i) c.LINp[NX * NX + i] = Bm[i];
        MMPC_UNROLL for (int i = 0; i < NX; ++i) c.LINp[NX * NX + NX * NU + i] = xd[i];
    }
    __syncwarp();

    double nu_merit = 1.0, delta_last = 0.0, err = INFINITY, fval = 0.0;
    int it = 0, it_done = 0, status = MMPC_SOLVE_MAX_ITER, ls_total = 0, reg_total = 0;
    bool active = valid;
    const int nfree = N * S, ncon = N * NX;
    for (it = 0;; ++it) {
        T2Norms nrm;
        double delta = 0.0;
        bool ok = t2_backward<M>(c, delta, nrm);
        if (active) {
            const double sd = fmax(100.0, nrm.lam_l1 / (double)(nfree + ncon)) / 100.0;  // IPOPT s_d, s_max = 100 [upstream]
            err = fmax(nrm.dual_inf / sd, nrm.prim_inf);
            fval = nrm.fval;
            it_done = it;
            if (!isfinite(err)) { status = MMPC_SOLVE_NUMERICAL; active = false; }
            else if (err <= io.tol) { status = MMPC_SOLVE_CONVERGED; active = false; }
            else if (it >= io.max_iter) { status = MMPC_SOLVE_MAX_ITER; active = false; }
        }
        if (!__any_sync(0xffffffffu, active)) break;
        // inertia correction: groups whose sweep met a non-PD G_uu retry with delta on the diagonal
        int tries = 0;
        while (__any_sync(0xffffffffu, active && !ok)) {
            if (active && !ok) {
                delta = (delta == 0.0) ? (delta_last > 0.0 ? fmax(1e-20, delta_last / 3.0) : 1e-4) : delta * 8.0;
                ++reg_total;
            }
            T2Norms dummy;
            const bool ok2 = t2_backward<M>(c, delta, dummy);
            if (active && !ok) ok = ok2;
            if (++tries >= 24) {
                if (active && !ok) { status = MMPC_SOLVE_NUMERICAL; active = false; }
                break;
            }
        }
        if (delta > 0.0) delta_last = delta;
        if (!__any_sync(0xffffffffu, active)) break;
        double gd, lam_inf, f1, th1;
        bool finite;
        t2_forward<M>(c, gd, lam_inf, f1, th1, finite);
        if (active && !finite) { status = MMPC_SOLVE_NUMERICAL; active = false; }
        nu_merit = fmax(nu_merit, 1.1 * lam_inf + 1e-8);
        const double m0 = nrm.fval + nu_merit * nrm.theta;
        const double D = gd - nu_merit * nrm.theta;
        double alpha = 1.0;
        int ls = 0;
        bool need = active && !(f1 + nu_merit * th1 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0));
        while (__any_sync(0xffffffffu, need)) {
            if (need) { alpha *= 0.5; ++ls_total; }
            double f2, t2;
            t2_trial<M>(c, alpha, f2, t2);
            if (need) need = !(f2 + nu_merit * t2 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0)) && (++ls < 40);
        }
        if (active) {  // batched so that the loads of one batch are all in flight together
            constexpr int UB = 8;
            for (int e0 = NX + gl; e0 < NV; e0 += G * UB) {
                double vz[UB], vd[UB];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * G < NV) { vz[q] = c.V[e0 + q * G]; vd[q] = c.DV[e0 + q * G]; }
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * G < NV) c.V[e0 + q * G] = fma(alpha, vd[q], vz[q]);
            }
            for (int e0 = gl; e0 < Mc; e0 += G * UB) {
                double vl[UB], vn[UB];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * G < Mc) { vl[q] = c.LAM[e0 + q * G]; vn[q] = c.LAMN[e0 + q * G]; }
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * G < Mc) c.LAM[e0 + q * G] = fma(alpha, vn[q] - vl[q], vl[q]);
            }
        }
        __syncwarp();
    }
    // ---- outputs: the iterate already is the reference's V layout ----
    if (valid) {
        double* Vo = io.V_out + (size_t)b * NV;
        for (int e = gl; e < NV; e += G) Vo[e] = c.V[e];
        if (io.lam_out) for (int e = gl; e < Mc; e += G) io.lam_out[(size_t)b * Mc + e] = c.LAM[e];
        if (io.stats && gl == 0) {
            mmpc_stats s;
            s.objective = fval; s.kkt_error = err; s.mu = 0.0; s.iterations = it_done; s.status = status;
            s.ls_steps = ls_total; s.reg_steps = reg_total;
            io.stats[b] = s;
        }
    }
}

template <class M>
struct LaunchT2 {
    static constexpr int WARPS = 1;
    static cudaError_t run(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        using C = T2<M>;
        const size_t smem = (size_t)WARPS * C::GPW * C::GS * sizeof(double);
        // the opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute (one process may drive
        // several GPUs through several handles); set once per device, atomically because handles may solve from several threads
        static std::atomic<unsigned long long> attr_done{0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem > 48 * 1024 && !((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ULL)) {
            cudaError_t e = cudaFuncSetAttribute(solve_kernel_t2<M, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attr_done.fetch_or(1ULL << (dev & 63), std::memory_order_release);
        }
        const int per_cta = WARPS * C::GPW;
        const int grid = (io.b1 - io.b0 + per_cta - 1) / per_cta;
        solve_kernel_t2<M, WARPS><<<grid, 32 * WARPS, smem, st>>>(P, W, io);
        return cudaGetLastError();
    }
};

}  // namespace mmpc

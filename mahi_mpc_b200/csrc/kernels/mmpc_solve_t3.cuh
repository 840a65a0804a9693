// mahi-mpc-b200 :: solve kernel, variant T3 -- a CTA per tile of I instances; worker warps evaluate the model for the
// next chunk of shooting nodes while warp 0 runs the Riccati recursion, one lane per instance, on the current chunk.
//
// Why (measured, profiles/r1_solve_t1_*, r1_solve_t2_*): T1 (thread per instance) executes the fewest instructions
// but evaluates the model inside the serial chain and keeps one warp per SM busy; T2 (8 lanes per instance)
// parallelises the evaluation but its cooperative Riccati spends as many issue slots per stage as one lane would.
// T3 keeps the best half of each and overlaps them:
//   * worker warps (threads 32..): thread (i, j) evaluates node k0+j of instance i for a chunk of C nodes (F, dF/dx,
//     dF/du, Lagrangian stage Hessian, residual norms; in the forward sweep the merit terms of the trial step and the
//     cp.async staging of the next chunk) and leaves a record in shared memory, field-major with the instance index
//     fastest so the lanes of a warp read / write consecutive 8-byte words (two wavefronts per access);
//   * warp 0: lane i runs T1's register-resident Riccati recursion (backward) and the linear rollout (forward) of
//     instance i from those records -- no exchange, no redundant work; gains and value-function blocks go to global
//     memory tile-wise ([node][field][instance]) so the lanes write full 256-byte lines and the forward sweep stages
//     a whole chunk with 16-byte cp.async copies;
//   * the two run concurrently on double-buffered records: one __syncthreads per chunk, no other synchronisation.
// One launch per batched solve; no grid-wide synchronisation.  Replaces IPOPT+MUMPS at ModelControl.cpp:159.
#pragma once
#if defined(MMPC_T3_PROFILE) || defined(MMPC_T3_PROFILE_SWEEPS) || defined(MMPC_T3_PROFILE_CTAS)
#include <cstdio>
#endif

namespace mmpc {

template <class M, int I_, int CB_, int CF_, int SK_ = 0, int RK_ = 0>
struct T3 {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, N2 = NX + 2 * NU, NH = S * (S + 1) / 2;
    static constexpr int I = I_;    // instances per CTA (<= 32: one serial lane each)
    static constexpr int CB = CB_;  // nodes per chunk in the backward sweep (one node per worker thread)
    static constexpr int CF = CF_;  // nodes per chunk in the forward sweep (CF <= CB)
    static_assert(CF_ <= CB_, "the forward chunk's merit evaluations use the backward chunk's worker slots");
    static constexpr int WORKERS = I * CB;
    // SK: warp 4 idles, so that the serial warp (warp 0) has its SM sub-partition -- and its fp64 pipe -- to itself
    static constexpr int SK = SK_;
    static constexpr int THREADS = 32 + WORKERS + (SK ? 32 : 0);
    __device__ static __forceinline__ bool is_worker(int tid) { return tid >= 32 && !(SK && tid >= 128 && tid < 160); }
    __device__ static __forceinline__ int widx(int tid) { return tid - 32 - ((SK && tid >= 160) ? 32 : 0); }
    // Node slot of worker group g = widx / I.  With one warp per slot (I = 32) warp 4 shares its SM sub-partition -- and its
    // fp64 pipe -- with the serial warp (warp 0): it gets the LAST slot, which the balanced chunks leave empty most of the
    // time (N = 50, CB = 7: 6 of 8 chunks have 6 nodes) and which takes no part in the forward sweep's merit evaluation.
    __device__ static __forceinline__ int slot_of(int g) { return (I == 32 && CB > 4) ? (g == 3 ? CB - 1 : (g == CB - 1 ? 3 : g)) : g; }
    // Fx and Fu are stored compacted to their structural non-zeros (generator masks; dense if a mask does not fit 64 bits);
    // the value-function block Pxx is stored as its upper triangle
    // RK_ is a mode mask.  bit 0: RK4 shooting intervals -- dense sensitivities (the one-step masks below hold for Euler /
    // linearised steps only); bit 1: finite bounds -- primal-dual interior point (T1's algorithm, mmpc_solve_t1.cuh)
    static constexpr int MODE = RK_, RK = RK_ & 1, BD = (RK_ >> 1) & 1;
    static constexpr bool SPARSE = !RK && NX * NX <= 64 && NX * NU <= 64;
    static constexpr unsigned long long fx_mask() {
        unsigned long long m = SPARSE ? M::A_MASK : ~0ULL;
        if (SPARSE) for (int i = 0; i < NX; ++i) m |= 1ULL << (i * NX + i);
        return m;
    }
    static constexpr unsigned long long fu_mask() { return SPARSE ? M::B_MASK : ~0ULL; }
    static constexpr int popc(unsigned long long m, int nbits) {
        int c = 0;
        for (int i = 0; i < nbits && i < 64; ++i) c += (int)((m >> i) & 1ULL);
        return c;
    }
    static constexpr unsigned long long FXM = fx_mask(), FUM = fu_mask();
    static constexpr int NFX = SPARSE ? popc(FXM, NX * NX) : NX * NX, NFU = SPARSE ? popc(FUM, NX * NU) : NX * NU;
    static constexpr int NPX = NX * (NX + 1) / 2;
    // backward record fields (shared): Fx | Fu | H (packed full stage Hessian) | c | q~ = (qx, qv, qu)
    // bounded: + Sigma and barrier gradient of x_{k+1} (they enter the value function of stage k+1)
    static constexpr int RB_FX = 0, RB_FU = RB_FX + NFX, RB_H = RB_FU + NFU, RB_C = RB_H + NH, RB_Q = RB_C + NX, RB_SN = RB_Q + N2, RB_GN = RB_SN + NX,
                         RB = RB_Q + N2 + (BD ? 2 * NX : 0);
    // global tile records per node
    static constexpr int FG_FX = 0, FG_FU = NFX, FG_C = FG_FU + NFU, FG_Q = FG_C + NX, FG_GN = FG_Q + N2, FG = FG_Q + N2 + (BD ? NX : 0);
    static constexpr int KG_KX = 0, KG_KV = NU * NX, KG_K0 = KG_KV + NU * NU, KG = KG_K0 + NU;
    static constexpr int PG_XX = 0, PG_XV = NPX, PG_X = PG_XV + NX * NU, PG = PG_X + NX;
    // node strides of the three global record arrays (T4 interleaves them into one record per node and overrides these)
    static constexpr int FG_NS = FG, KG_NS = KG, PG_NS = PG;
    // forward record (shared) = FG | KG | PG
    static constexpr int RF_K = FG, RF_P = FG + KG, RF = FG + KG + PG;
    static constexpr int BUF = (CB * RB > CF * RF ? CB * RB : CF * RF) * I;  // one record buffer (two are used)
    static constexpr int NRED = BD ? 9 : 5;  // per-(node,instance) reduction slots
    static constexpr int REDBUF = NRED * CB * I;
    static constexpr int SM_REC = 0, SM_RED = 2 * BUF, SM_INST = SM_RED + 2 * REDBUF, SM_END = SM_INST + 6 * I;  // per instance: delta, alpha, flag, mu, alpha_z, spare
    static constexpr size_t SMEM_BYTES = sizeof(double) * SM_END;
    static constexpr int LIN = NX * NX + NX * NU + NX;
};

// CTA-wide barrier usable from role-specific code paths (every thread of the CTA executes the same NUMBER of these;
// bar.sync counts arrivals per barrier, not per call site)
// __syncwarp first: with tile widths that are not multiples of 32 a warp holds lanes of several node groups; bar.sync is an
// aligned barrier, so the lanes of a warp arrive together
__device__ __forceinline__ void t3_cta_barrier() {
    __syncwarp();
    asm volatile("bar.sync 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}

// Structural sparsity of the one-step Jacobians under Euler / linearised steps: Fx = I + h A and Fu = h B inherit the
// pattern of A and B (e.g. the kinematic rows q' = qdot), which the generator reports as bit masks.  In the fully unrolled
// loops below the guards fold at compile time, so structurally zero products are never issued (CasADi's generated code
// exploits the same sparsity).  Not valid for RK4 (fill-in); T3 is Euler/linearised only.
#ifndef T3_EXPECT_CONVERGED_BELOW
#define T3_EXPECT_CONVERGED_BELOW 0.1
#endif
// record stores.  MMPC_T3_ST: 0 plain, 1 evict-first (st.global.cs), 2 L2 evict-last for a fraction of the lines.  The records
// of a whole batch (100 MB at B = 4096, N = 50) plus the iterates exceed the 126 MB L2 and are written and read in opposite
// orders: giving half of the lines priority keeps that half resident instead of letting LRU thrash all of them.  Measured at
// B = 4096 (T4, 32 x 7 x 7): plain 0.323, evict-first 0.325, evict-last 1.0 / 0.75 / 0.5 / 0.25: 0.323 / 0.318 / 0.318 / 0.317 ms.
#ifndef MMPC_T3_ST
#define MMPC_T3_ST 2
#endif
__device__ __forceinline__ void t3_st_evict_last(double* p, double v) {
    unsigned long long pol;
#ifndef MMPC_T3_ST_FRACTION
#define MMPC_T3_ST_FRACTION 0.5
#endif
#define MMPC_STR2(x) #x
#define MMPC_STR(x) MMPC_STR2(x)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, " MMPC_STR(MMPC_T3_ST_FRACTION) ";" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
#if MMPC_T3_ST == 0
#define T3_ST(p, v) (*(p) = (v))
#elif MMPC_T3_ST == 1
#define T3_ST(p, v) __stcs((p), (v))
#else
#define T3_ST(p, v) t3_st_evict_last((p), (v))
#endif
#define T3_FXNZ(r, c) (!C::SPARSE || ((C::FXM >> ((r) * NX + (c))) & 1ULL))
#define T3_FUNZ(r, c) (!C::SPARSE || ((C::FUM >> ((r) * NU + (c))) & 1ULL))
// slot of flat entry a in the compacted storage (folds to a constant in the unrolled loops)
#define T3_FXSLOT(a) (C::SPARSE ? __popcll(C::FXM & ((1ULL << (a)) - 1ULL)) : (a))
#define T3_FUSLOT(a) (C::SPARSE ? __popcll(C::FUM & ((1ULL << (a)) - 1ULL)) : (a))

// bounded problems only: multipliers of the lower / upper bounds and their steps, in V layout like the iterate
template <class M, bool BD>
struct T3CtxBounds {};
template <class M>
struct T3CtxBounds<M, true> {
    double *ZL, *ZU, *DZL, *DZU;
    double x_min[M::NX], x_max[M::NX], u_min[M::NU], u_max[M::NU];  // copies: reached through a reference, the kernel's parameter struct is read from local memory
};

template <class M, int I_, int MODE_ = 0>
struct T3Ctx : T3CtxBounds<M, ((MODE_ >> 1) & 1) != 0> {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, I = I_;
    double h;                     // step size; is_linear / substeps: copies of the problem constants the sweeps need
    int is_linear, substeps, gauss_newton, node_cost;
    int N;
    double *V, *LAM, *DV, *LAMN;  // this thread's instance inside the tile arrays: element e lives at [e * I]
    double* LINp;                 // linearisation data of the instance (AoS)
    const double* traj;
    double Q[NX], R[NU], Rm[NU], uinit[NU];
    template <int WANT>
    __device__ __forceinline__ void step(const double* x, const double* u, const double* wq, const double* r, const double* lam,
                                         double* F, double* Fx, double* Fu, double* Hs) const {
        if (is_linear) {
            double x0[NX];
            MMPC_UNROLL for (int i = 0; i < NX; ++i) x0[i] = V[(size_t)i * I];
            step_linear<M, WANT>(h, x, u, LINp, LINp + NX * NX, LINp + NX * NX + NX * NU, x0, uinit, F, Fx, Fu, Hs);
        } else if (MODE_ & 1) {
            if (WANT >= 2 && !gauss_newton) {
                step_rk4_hess<M>(h, substeps, x, u, wq, r, lam, F, Fx, Fu, Hs);  // exact Newton through the integrator
            } else {
                step_rk4<M, (WANT >= 1 ? 1 : 0)>(h, substeps, x, u, F, Fx, Fu);
                if (WANT >= 2) MMPC_UNROLL for (int i = 0; i < S * (S + 1) / 2; ++i) Hs[i] = 0.0;  // Gauss-Newton Hessian on request
            }
        } else {
            step_euler<M, WANT>(h, x, u, wq, r, lam, F, Fx, Fu, Hs);
        }
    }
};

// L2 -> L1 prefetch of the iterate rows node k's evaluation reads (issued one chunk ahead: the evaluation otherwise
// starts with a dependent chain of L2 round trips)
__device__ __forceinline__ void t3_prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Zero-copy inputs (mmpc_solve_io::traj_stage != NULL): state / control / traj point into page-locked HOST memory.  The CTA brings
// the rows of its tile's instances [bA, bB) on the device once, with coalesced loads (consecutive threads, consecutive words: whole
// PCIe read requests) issued eight at a time per thread so that few round trips cover them, and reads the device copies from then
// on (the reference trajectories are read in every sweep).  Ends with a CTA barrier that publishes the copies to the tile.
// One non-inlined copy for all kernels, every argument by value: the solve kernels' own code (register allocation of the sweeps)
// is the same with and without this path (inlined, it cost the device-resident solve 2-3 %: profiles/r2_latency.txt).
__device__ __forceinline__ void t3_stage_rows(const double* __restrict__ src, double* __restrict__ dst, int n, int tid, int nthreads) {
    constexpr int UB = 8;
    for (int e0 = tid; e0 < n; e0 += UB * nthreads) {
        double t[UB];
        MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * nthreads < n) t[q] = __ldcv(src + e0 + q * nthreads);
        MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * nthreads < n) dst[e0 + q * nthreads] = t[q];
    }
}
static __device__ __noinline__ void t3_stage_inputs(const double* state, const double* control, const double* traj, double* state_stage,
                                             double* control_stage, double* traj_stage, int bA, int bB, int nx, int nu, int N) {
    const int tid = threadIdx.x, nthreads = blockDim.x, nb = bB - bA;
    // the two short rows first, their loads in flight together with the first batch of the trajectories' loads
    const int ns = nb * nx, nc = nb * nu;
    double ts = 0.0, tc = 0.0;
    const bool small = ns <= nthreads && nc <= nthreads;
    if (small) {
        if (tid < ns) ts = __ldcv(state + (size_t)bA * nx + tid);
        if (tid < nc) tc = __ldcv(control + (size_t)bA * nu + tid);
    } else {
        t3_stage_rows(state + (size_t)bA * nx, state_stage + (size_t)bA * nx, ns, tid, nthreads);
        t3_stage_rows(control + (size_t)bA * nu, control_stage + (size_t)bA * nu, nc, tid, nthreads);
    }
    if (traj != traj_stage)  // (the trajectories may have been brought on the device by a copy engine already: MMPC_ZC_BIG=3)
        t3_stage_rows(traj + (size_t)bA * N * nx, traj_stage + (size_t)bA * N * nx, nb * N * nx, tid, nthreads);
    if (small) {
        if (tid < ns) state_stage[(size_t)bA * nx + tid] = ts;
        if (tid < nc) control_stage[(size_t)bA * nu + tid] = tc;
    }
    __syncthreads();
}

// Zero-copy outputs (mmpc_solve_io::host_out): V_out / lam_out / stats point into page-locked HOST memory and are written by the
// kernel in place (posted PCIe writes).  The iterate lives tile-major ([element][instance]), the outputs are instance-major: the
// plain epilogue's stores (lane = instance, 8 bytes every NV * 8 bytes) would leave the SM as 8-byte PCIe writes.  Here 64 elements of
// every instance pass through shared memory at a time, so each warp stores 256 contiguous bytes of one instance's vector.
// Non-inlined, arguments by value, tile width at run time (see t3_stage_inputs).
__device__ __forceinline__ void t3_rows_to_host(double* smem, const double* __restrict__ src_tile, double* __restrict__ dst, int rowlen,
                                                int nvalid, int tid, int nthreads, int I) {
    constexpr int E = 64, LD = E + 1, UB = 8;
    for (int e0 = 0; e0 < rowlen; e0 += E) {
        const int ne = min(E, rowlen - e0), n = ne * I;
        const double* src = src_tile + (size_t)e0 * I;
        __syncthreads();  // the previous chunk (or whatever used this shared memory before) has been consumed
        for (int t0 = tid; t0 < n; t0 += UB * nthreads) {
            double v[UB];
            MMPC_UNROLL for (int q = 0; q < UB; ++q) if (t0 + q * nthreads < n) v[q] = src[t0 + q * nthreads];
            MMPC_UNROLL for (int q = 0; q < UB; ++q) if (t0 + q * nthreads < n) {
                const int t = t0 + q * nthreads;
                smem[(t % I) * LD + t / I] = v[q];
            }
        }
        __syncthreads();
        for (int t = tid; t < nvalid * E; t += nthreads) {
            const int ii = t / E, r = t % E;
            if (r < ne) __stcs(&dst[(size_t)ii * rowlen + e0 + r], smem[ii * LD + r]);
        }
    }
}
// The statistics are handed over through shared memory (the kernel's serial lanes store their mmpc_stats at smem + I * 65 doubles,
// behind the transpose buffer): no solver state among the arguments.
static __device__ __noinline__ void t3_outputs_to_host_s(double* smem, double* V_out, double* lam_out, mmpc_stats* stats, int b1, int tile, int I,
                                                         int NV, int Mc, const double* V_tile, const double* LAM_tile, double* Lcan_tile) {
    static_assert(sizeof(mmpc_stats) == 5 * sizeof(double), "stats rows are copied as five words");
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int nvalid = min(I, b1 - tile * I);
    t3_rows_to_host(smem, V_tile, V_out + (size_t)tile * I * NV, NV, nvalid, tid, nthreads, I);
    if (lam_out) t3_rows_to_host(smem, LAM_tile, lam_out + (size_t)tile * I * Mc, Mc, nvalid, tid, nthreads, I);
    if (Lcan_tile) for (int t = tid; t < Mc * I; t += nthreads) Lcan_tile[t] = LAM_tile[t];  // warm start: same tile layout
    if (stats) {
        const double* src = smem + I * 65;
        double* dst = reinterpret_cast<double*>(stats + (size_t)tile * I);
        for (int t = tid; t < nvalid * 5; t += nthreads) __stcs(&dst[t], src[t]);
    }
}

template <class M, class C>
__device__ __forceinline__ void t3_prefetch_node(const T3Ctx<M, C::I, C::MODE>& c, int k) {
    constexpr int NX = M::NX, S = C::S, I = C::I;
    const double* z = c.V + (size_t)k * S * I;
    MMPC_UNROLL for (int a = 0; a < S + NX; ++a) t3_prefetch(z + a * I);  // x_k, u_k, x_{k+1}; u_{k-1} and u_{k+1} belong to the neighbours
    MMPC_UNROLL for (int a = 0; a < NX; ++a) t3_prefetch(c.LAM + ((size_t)k * NX + a) * I);
    t3_prefetch(c.traj + (size_t)k * NX);
}

// ---- parallel phase of the backward sweep: thread (i, j) evaluates node k = k0 + j of instance i -------------
// red slots: 0 dual_inf(max) 1 prim_inf(max) 2 theta 3 fval 4 lam_l1
template <class M, class C>
__device__ __forceinline__ void t3_eval_node(const T3Ctx<M, C::I, C::MODE>& c, int k, int i, int j, double delta, double mu, double* rec, double* red,
                                             double* fg_tile) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, NH = C::NH, I = C::I;
    const int N = c.N;
    const double* z = c.V + (size_t)k * S * I;
    double x[NX], u[NU], v[NU], r[NX], lam[NX], xn[NX], wq[NX];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        x[a] = z[a * I];
        xn[a] = z[(S + a) * I];
        r[a] = c.traj[(size_t)k * NX + a];
        lam[a] = c.LAM[((size_t)k * NX + a) * I];
        wq[a] = 2.0 * c.Q[a];
    }
    double lam_prev[NX], u_next[NU];  // neighbours, loaded up front so that every global load of the node is in flight together
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        u[a] = z[(NX + a) * I];
        v[a] = k ? z[(NX + a - S) * I] : c.uinit[a];
        u_next[a] = (k < N - 1) ? z[(S + NX + a) * I] : 0.0;
    }
    MMPC_UNROLL for (int a = 0; a < NX; ++a) lam_prev[a] = k ? c.LAM[((size_t)(k - 1) * NX + a) * I] : 0.0;
    // bounded: Sigma = z/s and the barrier gradient of u_k (stage Hessian / linear term) and of x_{k+1} (value function of
    // stage k+1); the residual norms of the bounded problem come from t3_residuals, not from here
    double sgu[NU], gpu_[NU], sgn[NX], gpn[NX];
    if constexpr (C::BD) {
        SweepNorms unused{};
        const double* zl = c.ZL + (size_t)k * S * I;
        const double* zu = c.ZU + (size_t)k * S * I;
        MMPC_UNROLL for (int a = 0; a < NU; ++a) {
            const BoundTerm t = bound_term(u[a], c.u_min[a], c.u_max[a], zl[(NX + a) * I], zu[(NX + a) * I], mu, unused);
            sgu[a] = t.sigma; gpu_[a] = t.gphi;
        }
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            const BoundTerm t = bound_term(xn[a], c.x_min[a], c.x_max[a], zl[(S + a) * I], zu[(S + a) * I], mu, unused);
            sgn[a] = t.sigma; gpn[a] = t.gphi;
        }
    }
    double F[NX], Fx[NX * NX], Fu[NX * NU], H[NH];
    // MMPC_COST_NODE: the tracking error sits on the decision variable x_{k+1}; node k then carries x_k's own term
    // 2 Q (x_k - r_{k-1}) (gradient) and 2 Q (Hessian diagonal) for k >= 1, x_N's term is the terminal value function
    // (t3_backward), the Hessian weights are lam alone and there is no Gauss-Newton block
    const bool nc = c.node_cost != 0;
    double wqh[NX], gxn[NX];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        wqh[a] = nc ? 0.0 : wq[a];
        gxn[a] = (nc && k) ? wq[a] * (x[a] - c.traj[(size_t)(k - 1) * NX + a]) : 0.0;
    }
    c.template step<2>(x, u, wqh, r, lam, F, Fx, Fu, H);
    double* const recp = rec + (size_t)j * C::RB * I + i;
    double* const fgp = fg_tile + (size_t)k * C::FG_NS * I + i;
#define T3_REC(f) recp[(f) * I]
#define T3_FG(f, v) T3_ST(&fgp[(f) * I], (v))
    double qe[NX], dual = 0.0, prim = 0.0, theta = 0.0, fval = 0.0, lam1 = 0.0;
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        const double e = (nc ? xn[a] : F[a]) - r[a], cr = F[a] - xn[a];
        qe[a] = wqh[a] * e;  // pulled back through Fx, Fu below (shooting cost only)
        fval = fma(c.Q[a] * e, e, fval);
        prim = fmax(prim, fabs(cr));
        theta += fabs(cr);
        lam1 += fabs(lam[a]);
        T3_REC(C::RB_C + a) = cr;
        T3_FG(C::FG_C + a, cr);
        if (k == N - 1) dual = fmax(dual, fabs((nc ? wq[a] * e : 0.0) - lam[a]));  // stationarity w.r.t. x_N: (2 Q e_N) - lam_{N-1}
    }
    MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) if (T3_FXNZ(a / NX, a % NX)) { T3_REC(C::RB_FX + T3_FXSLOT(a)) = Fx[a]; T3_FG(C::FG_FX + T3_FXSLOT(a), Fx[a]); }
    MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) if (T3_FUNZ(a / NU, a % NU)) { T3_REC(C::RB_FU + T3_FUSLOT(a)) = Fu[a]; T3_FG(C::FG_FU + T3_FUSLOT(a), Fu[a]); }
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        double g = 0.0, l = 0.0;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(q, a)) {
            g = fma(Fx[q * NX + a], qe[q], g);
            l = fma(Fx[q * NX + a], lam[q], l);
        }
        g += gxn[a];
        T3_REC(C::RB_Q + a) = g;
        T3_FG(C::FG_Q + a, g);
        if (k) dual = fmax(dual, fabs(g + l - lam_prev[a]));
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        double g = 0.0, l = 0.0;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FUNZ(q, a)) {
            g = fma(Fu[q * NU + a], qe[q], g);
            l = fma(Fu[q * NU + a], lam[q], l);
        }
        const double du = u[a] - v[a], dRu = 2.0 * c.R[a] * du;
        double qu = g + dRu + 2.0 * c.Rm[a] * u[a];
        const double dRn = (k < N - 1) ? 2.0 * c.R[a] * (u_next[a] - u[a]) : 0.0;
        fval = fma(c.R[a] * du, du, fval);
        fval = fma(c.Rm[a] * u[a], u[a], fval);
        dual = fmax(dual, fabs(qu + l - dRn));
        if constexpr (C::BD != 0) { qu += gpu_[a]; }
        T3_REC(C::RB_Q + NX + a) = -dRu;  T3_FG(C::FG_Q + NX + a, -dRu);
        T3_REC(C::RB_Q + S + a) = qu;     T3_FG(C::FG_Q + S + a, qu);
    }
    MMPC_UNROLL for (int p = 0; p < S; ++p) MMPC_UNROLL for (int d = p; d < S; ++d) {
        double acc = H[tri(p, d, S)];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) {
            const bool nzp = p < NX ? T3_FXNZ(q, p) : T3_FUNZ(q, p - NX), nzd = d < NX ? T3_FXNZ(q, d) : T3_FUNZ(q, d - NX);
            if (nzp && nzd) {
                const double fa = p < NX ? Fx[q * NX + p] : Fu[q * NU + p - NX];
                const double fd = d < NX ? Fx[q * NX + d] : Fu[q * NU + d - NX];
                acc = fma(wqh[q] * fa, fd, acc);
            }
        }
        if (p == d) acc += delta + (p >= NX ? 2.0 * (c.R[p - NX] + c.Rm[p - NX]) : ((nc && k) ? wq[p] : 0.0));
        if constexpr (C::BD != 0) {
            if (p == d && p >= NX) acc += sgu[p - NX];
        }
        T3_REC(C::RB_H + tri(p, d, S)) = acc;
    }
    if constexpr (C::BD != 0) {
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            T3_REC(C::RB_SN + a) = sgn[a];
            T3_REC(C::RB_GN + a) = gpn[a];
            T3_FG(C::FG_GN + a, gpn[a]);
        }
    }
#undef T3_REC
#undef T3_FG
    red[(0 * C::CB + j) * I + i] = dual;
    red[(1 * C::CB + j) * I + i] = prim;
    red[(2 * C::CB + j) * I + i] = theta;
    red[(3 * C::CB + j) * I + i] = fval;
    red[(4 * C::CB + j) * I + i] = lam1;
}

// ---- serial phase of the backward sweep: one lane per instance, T1's Riccati update on the augmented state -----
template <class M>
struct T3Riccati {
    static constexpr int NX = M::NX, NU = M::NU;
    double Pxx[NX * NX], Pxv[NX * NU], Pvv[NU * NU], px[NX], pv[NU];
};

template <class M, class C>
__device__ __forceinline__ bool t3_riccati_stage(T3Riccati<M>& V, const double* rec, int s, int i, const double* R, double* kg, double* pg) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    const double* const recp = rec + (size_t)s * C::RB * I + i;
#define T3_R(f) recp[(f) * I]
    double Fx[NX * NX], Fu[NX * NU], cres[NX];
    MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) Fx[a] = T3_FXNZ(a / NX, a % NX) ? T3_R(C::RB_FX + T3_FXSLOT(a)) : 0.0;
    MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) Fu[a] = T3_FUNZ(a / NU, a % NU) ? T3_R(C::RB_FU + T3_FUSLOT(a)) : 0.0;
    MMPC_UNROLL for (int a = 0; a < NX; ++a) cres[a] = T3_R(C::RB_C + a);
    // NOTE: keep `#pragma unroll` loops of an `if constexpr` branch inside braces.  With nvcc 12.9 the form
    // `if constexpr (C::BD) _Pragma("unroll") for (...) {...}` changed the multipliers this function's callers computed in the
    // instantiations where the branch is discarded (one more Newton iteration per solve; tools/it_cmp.py shows it).  The cause is
    // not established: a minimal reproducer of the construct (an `if constexpr` followed by a pragma line and a loop) discards
    // the loop correctly in PTX, and compute-sanitizer is not available on this GPU pool.  The braced form is what the
    // iteration-count parity tests against the oracle (every variant, all 4096 instances of the bench batch) pin.
    if constexpr (C::BD != 0) {  // barrier of x_{k+1}: part of the value function of stage k+1
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            V.Pxx[a * NX + a] += T3_R(C::RB_SN + a);
            V.px[a] += T3_R(C::RB_GN + a);
        }
    }
    // value function of stage k+1 -> global (the forward sweep needs it for lam_k)
    MMPC_UNROLL for (int a = 0; a < NX; ++a) MMPC_UNROLL for (int b = a; b < NX; ++b) T3_ST(&pg[(C::PG_XX + tri(a, b, NX)) * I], V.Pxx[a * NX + b]);
    MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) T3_ST(&pg[(C::PG_XV + a) * I], V.Pxv[a]);
    MMPC_UNROLL for (int a = 0; a < NX; ++a) T3_ST(&pg[(C::PG_X + a) * I], V.px[a]);
    double y[NX], yv[NU];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        double t = V.px[a];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(V.Pxx[a * NX + q], cres[q], t);
        y[a] = t;
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        double t = V.pv[a];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(V.Pxv[q * NU + a], cres[q], t);
        yv[a] = t;
    }
    double PA[NX * NX], PB[NX * NU];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        MMPC_UNROLL for (int b = 0; b < NX; ++b) {
            double t = 0.0;
            bool first = true;  // folds at compile time: the first product initialises the accumulator (no zeroing, no fma with 0)
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(q, b)) {
                t = first ? V.Pxx[a * NX + q] * Fx[q * NX + b] : fma(V.Pxx[a * NX + q], Fx[q * NX + b], t);
                first = false;
            }
            PA[a * NX + b] = t;
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) {
            double t = V.Pxv[a * NU + b];
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FUNZ(q, b)) t = fma(V.Pxx[a * NX + q], Fu[q * NU + b], t);
            PB[a * NU + b] = t;
        }
    }
    double Gxx[NX * NX], Gxu[NX * NU], Guu[NU * NU], gx[NX], gu[NU];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        MMPC_UNROLL for (int b = a; b < NX; ++b) {
            double t = T3_R(C::RB_H + tri(a, b, S));
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(q, a)) t = fma(Fx[q * NX + a], PA[q * NX + b], t);
            Gxx[a * NX + b] = t;
            Gxx[b * NX + a] = t;
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) {
            double t = T3_R(C::RB_H + tri(a, NX + b, S));
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(q, a)) t = fma(Fx[q * NX + a], PB[q * NU + b], t);
            Gxu[a * NU + b] = t;
        }
        double t = T3_R(C::RB_Q + a);
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(q, a)) t = fma(Fx[q * NX + a], y[q], t);
        gx[a] = t;
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        MMPC_UNROLL for (int b = a; b < NU; ++b) {
            double t = T3_R(C::RB_H + tri(NX + a, NX + b, S)) + 0.5 * (V.Pvv[a * NU + b] + V.Pvv[b * NU + a]);
            MMPC_UNROLL for (int q = 0; q < NX; ++q) {
                if (T3_FUNZ(q, a)) t = fma(Fu[q * NU + a], PB[q * NU + b], t);
                if (T3_FUNZ(q, b)) t = fma(V.Pxv[q * NU + a], Fu[q * NU + b], t);
            }
            Guu[a * NU + b] = t;
            Guu[b * NU + a] = t;
        }
        double t = T3_R(C::RB_Q + S + a) + yv[a];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FUNZ(q, a)) t = fma(Fu[q * NU + a], y[q], t);
        gu[a] = t;
    }
    bool ok;
    double Kx[NU * NX], Kv[NU * NU], k0[NU];
    if constexpr (NU == 2) {
        // two controls: explicit inverse of the 2 x 2 block (one reciprocal on the stage's dependent chain instead of the two
        // dependent rsqrt of the Cholesky factor and its substitutions); positive definite <=> G00 > 0 and det > 0
        const double det = fma(Guu[0], Guu[3], -Guu[1] * Guu[1]);
        ok = (Guu[0] > 0.0) && (det > 0.0) && isfinite(det);
        const double r = 1.0 / (ok ? det : 1.0);
        const double i00 = Guu[3] * r, i01 = -Guu[1] * r, i11 = Guu[0] * r;
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            Kx[0 * NX + a] = -fma(i00, Gxu[a * NU + 0], i01 * Gxu[a * NU + 1]);
            Kx[1 * NX + a] = -fma(i01, Gxu[a * NU + 0], i11 * Gxu[a * NU + 1]);
        }
        Kv[0 * NU + 0] = i00 * (2.0 * R[0]); Kv[1 * NU + 0] = i01 * (2.0 * R[0]);
        Kv[0 * NU + 1] = i01 * (2.0 * R[1]); Kv[1 * NU + 1] = i11 * (2.0 * R[1]);
        k0[0] = -fma(i00, gu[0], i01 * gu[1]);
        k0[1] = -fma(i01, gu[0], i11 * gu[1]);
    } else {
        ok = chol<NU>(Guu);
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            double col[NU];
            MMPC_UNROLL for (int b = 0; b < NU; ++b) col[b] = -Gxu[a * NU + b];
            chol_solve<NU>(Guu, col);
            MMPC_UNROLL for (int b = 0; b < NU; ++b) Kx[b * NX + a] = col[b];
        }
        MMPC_UNROLL for (int l = 0; l < NU; ++l) {
            double col[NU];
            MMPC_UNROLL for (int b = 0; b < NU; ++b) col[b] = (b == l) ? 2.0 * R[l] : 0.0;
            chol_solve<NU>(Guu, col);
            MMPC_UNROLL for (int b = 0; b < NU; ++b) Kv[b * NU + l] = col[b];
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) k0[b] = -gu[b];
        chol_solve<NU>(Guu, k0);
    }
    MMPC_UNROLL for (int a = 0; a < NU * NX; ++a) T3_ST(&kg[(C::KG_KX + a) * I], Kx[a]);
    MMPC_UNROLL for (int a = 0; a < NU * NU; ++a) T3_ST(&kg[(C::KG_KV + a) * I], Kv[a]);
    MMPC_UNROLL for (int a = 0; a < NU; ++a) T3_ST(&kg[(C::KG_K0 + a) * I], k0[a]);
    const double qv0 = 0.0;
    (void)qv0;
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        MMPC_UNROLL for (int b = a; b < NX; ++b) {
            double t = Gxx[a * NX + b];
            MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Gxu[a * NU + l], Kx[l * NX + b], t);
            V.Pxx[a * NX + b] = t;
            V.Pxx[b * NX + a] = t;
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) {
            double t = Gxu[a * NU + 0] * Kv[0 * NU + b];
            MMPC_UNROLL for (int l = 1; l < NU; ++l) t = fma(Gxu[a * NU + l], Kv[l * NU + b], t);
            V.Pxv[a * NU + b] = t;
        }
        double t = gx[a];
        MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Gxu[a * NU + l], k0[l], t);
        V.px[a] = t;
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        MMPC_UNROLL for (int b = 0; b < NU; ++b) V.Pvv[a * NU + b] = (a == b ? 2.0 * R[a] : 0.0) - 2.0 * R[a] * Kv[a * NU + b];
        V.pv[a] = T3_R(C::RB_Q + NX + a) - 2.0 * R[a] * k0[a];
    }
#undef T3_R
    return ok;
}

struct T3Norms {
    double dual_inf, prim_inf, theta, fval, lam_l1;
};

// whole backward sweep for the CTA tile, software pipelined: while warp 0 runs the Riccati recursion over chunk ch (one
// lane per instance) the worker warps evaluate chunk ch-1 into the other record buffer.
// returns (valid on the serial lanes) whether every G_uu was positive definite
template <class M, class C>
__device__ __noinline__ bool t3_backward(const T3Ctx<M, C::I, C::MODE>& cin, double* smem, int tid, double delta_self, T3Norms& nrm_out,
                                         double* fg_tile, double* kg_tile, double* pg_tile) {
    constexpr int NX = M::NX, NU = M::NU, I = C::I;
    const T3Ctx<M, C::I, C::MODE> c = cin;  // callee-local copy: the context and all accumulators stay in registers (the caller's copy has
                             // its address taken and therefore lives in local memory)
    const int N = c.N;
    const bool worker = C::is_worker(tid);
    const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
    double* dl = smem + C::SM_INST;  // per-instance delta, published by the serial lanes
    double n_dual = 0.0, n_prim = 0.0, n_theta = 0.0, n_fval = 0.0, n_lam = 0.0;
    bool ok = true;
    if (tid < I) dl[tid] = delta_self;
    __syncthreads();
    const int nch = (N + C::CB - 1) / C::CB;
#ifdef MMPC_T3_PROFILE_SWEEPS
    long long t_work = 0, t_all0 = clock64();
#endif
    // step = nch: workers evaluate the last chunk (prologue); step = ch < nch: warp 0 consumes chunk ch while the workers
    // evaluate chunk ch-1.  The two roles run their own copy of the step loop (same trip count, one CTA barrier per step):
    // the Riccati state then is not live across the evaluator and the evaluator's temporaries are not live across the
    // recursion, which is worth ~70 registers to each side.
    if (worker) {
        const double delta = dl[wi];
        const double mu = C::BD ? smem[C::SM_INST + 3 * I + wi] : 0.0;  // barrier parameter of the instance (bounded problems)
        for (int step = nch; step >= 0; --step) {
#ifdef MMPC_T3_PROFILE_SWEEPS
            const long long t0 = clock64();
#endif
            const int ch = step - 1;
            if (ch >= 0) {
                const int k0 = (int)((long long)ch * N / nch), kn = (int)((long long)(ch + 1) * N / nch) - k0;  // balanced chunks (<= CB nodes)
                if (ch > 0 && wj < k0 - (int)((long long)(ch - 1) * N / nch)) t3_prefetch_node<M, C>(c, (int)((long long)(ch - 1) * N / nch) + wj);
                if (wj < kn) t3_eval_node<M, C>(c, k0 + wj, wi, wj, delta, mu, smem + C::SM_REC + (ch & 1) * C::BUF, smem + C::SM_RED + (ch & 1) * C::REDBUF, fg_tile);
            }
#ifdef MMPC_T3_PROFILE_SWEEPS
            t_work += clock64() - t0;
#endif
            t3_cta_barrier();
        }
    } else {
        T3Riccati<M> V;
        MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) V.Pxx[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) V.Pxv[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NU * NU; ++a) V.Pvv[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NX; ++a) V.px[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NU; ++a) V.pv[a] = 0.0;
        if (c.node_cost && tid < I) {  // MMPC_COST_NODE: x_N's tracking term is the terminal value function
            MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                V.Pxx[a * NX + a] = 2.0 * c.Q[a];
                V.px[a] = 2.0 * c.Q[a] * (c.V[((size_t)N * C::S + a) * I] - c.traj[(size_t)(N - 1) * NX + a]);
            }
        }
        double R[NU];
        MMPC_UNROLL for (int a = 0; a < NU; ++a) R[a] = c.R[a];
        for (int step = nch; step >= 0; --step) {
#ifdef MMPC_T3_PROFILE_SWEEPS
            const long long t0 = clock64();
#endif
            if (tid < I && step < nch) {
                const int ch = step, k0 = (int)((long long)ch * N / nch), kn = (int)((long long)(ch + 1) * N / nch) - k0;
                const double* rec = smem + C::SM_REC + (ch & 1) * C::BUF;
                const double* red = smem + C::SM_RED + (ch & 1) * C::REDBUF;
                for (int s = kn - 1; s >= 0; --s) {
                    const int k = k0 + s;
                    ok = t3_riccati_stage<M, C>(V, rec, s, tid, R, kg_tile + (size_t)k * C::KG_NS * I + tid, pg_tile + (size_t)k * C::PG_NS * I + tid) && ok;
                    n_dual = fmax(n_dual, red[(0 * C::CB + s) * I + tid]);
                    n_prim = fmax(n_prim, red[(1 * C::CB + s) * I + tid]);
                    n_theta += red[(2 * C::CB + s) * I + tid];
                    n_fval += red[(3 * C::CB + s) * I + tid];
                    n_lam += red[(4 * C::CB + s) * I + tid];
                }
            }
#ifdef MMPC_T3_PROFILE_SWEEPS
            t_work += clock64() - t0;
#endif
            t3_cta_barrier();
        }
    }
#ifdef MMPC_T3_PROFILE_SWEEPS
    if (blockIdx.x == 0 && (tid == 0 || tid == 32))
        printf("[t3 backward] tid %d busy %lld of %lld cycles (%d chunks of %d)\n", tid, t_work, clock64() - t_all0, nch, C::CB);
#endif
    nrm_out.dual_inf = n_dual; nrm_out.prim_inf = n_prim; nrm_out.theta = n_theta; nrm_out.fval = n_fval; nrm_out.lam_l1 = n_lam;
    return ok;
}

// ---- KKT residuals of the current iterate without a Riccati sweep (T1's backward_sweep<RICCATI = false>).  Bounded problems:
// the barrier parameter is chosen from them BEFORE the Riccati sweep, so they get their own pass.  Unbounded problems: the
// convergence test of an iterate that is expected to pass it (see the kernel).  Workers only, node k of instance i per
// thread; first derivatives only.  Slots: 0 dual_inf 1 prim_inf 2 theta 3 fval 4 lam_l1 5 z_l1 6 max s*z 7 min s*z 8 -sum log s
struct T3NormsBD {
    double dual_inf, prim_inf, theta, fval, lam_l1, z_l1, cmax, cmin, barrier;
};
template <class M, class C>
__device__ __noinline__ void t3_residuals(const T3Ctx<M, C::I, C::MODE>& cin, double* smem, int tid, T3NormsBD& out) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    const T3Ctx<M, C::I, C::MODE> c = cin;
    const int N = c.N;
    const bool worker = C::is_worker(tid);
    const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
    double* red = smem + C::SM_RED;
    if (worker) {
        SweepNorms n;
        n.dual_inf = n.prim_inf = n.theta = n.fval = n.lam_l1 = n.z_l1 = n.cmax = n.barrier = 0.0;
        n.cmin = INFINITY;
        for (int k = wj; k < N; k += C::CB) {
            const double* z = c.V + (size_t)k * S * I;
            double x[NX], u[NU], v[NU], r[NX], lam[NX], xn[NX], lam_prev[NX], u_next[NU];
            double zlx[NX], zux[NX], zlu[NU], zuu[NU], zln[NX], zun[NX];
            MMPC_UNROLL for (int a = 0; a < NX; ++a) zlx[a] = zux[a] = zln[a] = zun[a] = 0.0;
            MMPC_UNROLL for (int a = 0; a < NU; ++a) zlu[a] = zuu[a] = 0.0;
            if constexpr (C::BD != 0) {
                const double* zl = c.ZL + (size_t)k * S * I;
                const double* zu = c.ZU + (size_t)k * S * I;
                MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                    zlx[a] = k ? zl[a * I] : 0.0; zux[a] = k ? zu[a * I] : 0.0;
                    zln[a] = zl[(S + a) * I]; zun[a] = zu[(S + a) * I];
                }
                MMPC_UNROLL for (int a = 0; a < NU; ++a) { zlu[a] = zl[(NX + a) * I]; zuu[a] = zu[(NX + a) * I]; }
            }
            MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                x[a] = z[a * I];
                xn[a] = z[(S + a) * I];
                r[a] = c.traj[(size_t)k * NX + a];
                lam[a] = c.LAM[((size_t)k * NX + a) * I];
                lam_prev[a] = k ? c.LAM[((size_t)(k - 1) * NX + a) * I] : 0.0;
            }
            MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                u[a] = z[(NX + a) * I];
                v[a] = k ? z[(NX + a - S) * I] : c.uinit[a];
                u_next[a] = (k < N - 1) ? z[(S + NX + a) * I] : 0.0;
            }
            double F[NX], Fx[NX * NX], Fu[NX * NU], qe[NX];
            c.template step<1>(x, u, nullptr, nullptr, nullptr, F, Fx, Fu, nullptr);
            SweepNorms own{};  // x_k's terms are accumulated by node k-1 (as its x_{k+1}); here only its -zl + zu is needed
            const bool nc = c.node_cost != 0;  // MMPC_COST_NODE: see t3_eval_node
            double qxn[NX];
            MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                const double e = (nc ? xn[a] : F[a]) - r[a], cr = F[a] - xn[a];
                qe[a] = nc ? 0.0 : 2.0 * c.Q[a] * e;
                qxn[a] = nc ? 2.0 * c.Q[a] * e : 0.0;
                n.fval = fma(c.Q[a] * e, e, n.fval);
                n.prim_inf = fmax(n.prim_inf, fabs(cr));
                n.theta += fabs(cr);
                n.lam_l1 += fabs(lam[a]);
                double dres_n = 0.0;
                if constexpr (C::BD != 0) { dres_n = bound_term(xn[a], c.x_min[a], c.x_max[a], zln[a], zun[a], 0.0, n).dres; }
                if (k == N - 1) n.dual_inf = fmax(n.dual_inf, C::BD ? fabs(qxn[a] - lam[a] + dres_n) : fabs(qxn[a] - lam[a]));  // stationarity w.r.t. x_N
            }
            MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                double g = 0.0, l = 0.0;
                MMPC_UNROLL for (int q = 0; q < NX; ++q) {
                    g = fma(Fx[q * NX + a], qe[q], g);
                    l = fma(Fx[q * NX + a], lam[q], l);
                }
                if (nc && k) g += 2.0 * c.Q[a] * (x[a] - c.traj[(size_t)(k - 1) * NX + a]);
                if (k) {
                    if constexpr (C::BD != 0) {
                        const BoundTerm tx = bound_term(x[a], c.x_min[a], c.x_max[a], zlx[a], zux[a], 0.0, own);
                        n.dual_inf = fmax(n.dual_inf, fabs(g + l - lam_prev[a] + tx.dres));
                    } else {
                        n.dual_inf = fmax(n.dual_inf, fabs(g + l - lam_prev[a]));
                    }
                }
            }
            MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                double g = 0.0, l = 0.0;
                MMPC_UNROLL for (int q = 0; q < NX; ++q) {
                    g = fma(Fu[q * NU + a], qe[q], g);
                    l = fma(Fu[q * NU + a], lam[q], l);
                }
                const double du = u[a] - v[a], dRu = 2.0 * c.R[a] * du;
                const double dRn = (k < N - 1) ? 2.0 * c.R[a] * (u_next[a] - u[a]) : 0.0;
                n.fval = fma(c.R[a] * du, du, n.fval);
                n.fval = fma(c.Rm[a] * u[a], u[a], n.fval);
                const double qu = g + dRu + 2.0 * c.Rm[a] * u[a];
                if constexpr (C::BD != 0) {
                    const BoundTerm tu = bound_term(u[a], c.u_min[a], c.u_max[a], zlu[a], zuu[a], 0.0, n);
                    n.dual_inf = fmax(n.dual_inf, fabs(qu + l - dRn + tu.dres));
                } else {
                    n.dual_inf = fmax(n.dual_inf, fabs(qu + l - dRn));
                }
            }
        }
        red[(0 * C::CB + wj) * I + wi] = n.dual_inf; red[(1 * C::CB + wj) * I + wi] = n.prim_inf;
        red[(2 * C::CB + wj) * I + wi] = n.theta;    red[(3 * C::CB + wj) * I + wi] = n.fval;
        red[(4 * C::CB + wj) * I + wi] = n.lam_l1;
        if constexpr (C::BD != 0) {
            red[(5 * C::CB + wj) * I + wi] = n.z_l1;  red[(6 * C::CB + wj) * I + wi] = n.cmax;
            red[(7 * C::CB + wj) * I + wi] = n.cmin;  red[(8 * C::CB + wj) * I + wi] = n.barrier;
        }
    }
    __syncwarp();  // node groups of one warp may differ in trip count (tile widths below 32): arrive at the aligned CTA barrier together
    __syncthreads();
    T3NormsBD o;
    o.dual_inf = o.prim_inf = o.theta = o.fval = o.lam_l1 = o.z_l1 = o.cmax = o.barrier = 0.0;
    o.cmin = INFINITY;
    if (tid < I) MMPC_UNROLL for (int s = 0; s < C::CB; ++s) {
        o.dual_inf = fmax(o.dual_inf, red[(0 * C::CB + s) * I + tid]); o.prim_inf = fmax(o.prim_inf, red[(1 * C::CB + s) * I + tid]);
        o.theta += red[(2 * C::CB + s) * I + tid];                      o.fval += red[(3 * C::CB + s) * I + tid];
        o.lam_l1 += red[(4 * C::CB + s) * I + tid];
        if constexpr (C::BD != 0) {
            o.z_l1 += red[(5 * C::CB + s) * I + tid];              o.cmax = fmax(o.cmax, red[(6 * C::CB + s) * I + tid]);
            o.cmin = fmin(o.cmin, red[(7 * C::CB + s) * I + tid]); o.barrier += red[(8 * C::CB + s) * I + tid];
        }
    }
    __syncthreads();
    out = o;
}

// merit pieces of node k at z + alpha dz
// bounded problems: barrier of the trial point (full = false: merit passes, t3_trial) or, in the forward sweep (full = true),
// the multiplier steps and the fraction-to-the-boundary limits only -- the merit of a bounded problem is evaluated at
// alpha = alpha_max by one t3_trial pass after the sweep, so the sweep itself needs neither the model nor the logarithms
struct T3TrialBD {
    double bar1, amax, az, mu, tau;
    bool full;
};
// T1's bound_step without the barrier of the full step
__device__ __forceinline__ void t3_bound_step(double v, double d, double lb, double ub, double zl, double zu, double mu, double tau, double& dzl,
                                              double& dzu, double& amax, double& az) {
    dzl = 0.0; dzu = 0.0;
    if (lb > -MMPC_INF_THRESH) {
        const double s = v - lb;
        dzl = (mu - zl * d) / s - zl;
        if (d < 0.0) amax = fmin(amax, -tau * s / d);
        if (dzl < 0.0) az = fmin(az, -tau * zl / dzl);
    }
    if (ub < MMPC_INF_THRESH) {
        const double s = ub - v;
        dzu = (mu + zu * d) / s - zu;
        if (d > 0.0) amax = fmin(amax, tau * s / d);
        if (dzu < 0.0) az = fmin(az, -tau * zu / dzu);
    }
}
template <class M, class C>
__device__ __forceinline__ void t3_trial_node(const T3Ctx<M, C::I, C::MODE>& c, int k, double alpha, double& f1, double& th1, T3TrialBD& bd,
                                              bool keep_lam = false) {
    constexpr int NX = M::NX, NU = M::NU, S = NX + NU, I = C::I;
    const double* z = c.V + (size_t)k * S * I;
    const double* d = c.DV + (size_t)k * S * I;
    double x[NX], u[NU], v[NU], xn[NX], F[NX], r[NX];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) r[a] = c.traj[(size_t)k * NX + a];  // with the other loads: one L2 round trip, not two
    // bounded, full step: the bound multipliers of u_k and x_{k+1}, loaded with everything else (one L2 round trip)
    double zlq[NX + NU], zuq[NX + NU];
    if constexpr (C::BD != 0) {
        if (bd.full) {
            const double* zl = c.ZL + ((size_t)k * S + NX) * I;
            const double* zu = c.ZU + ((size_t)k * S + NX) * I;
            MMPC_UNROLL for (int q = 0; q < S; ++q) {  // q < NU: u_k; q >= NU: x_{k+1}
                const double lb = q < NU ? c.u_min[q] : c.x_min[q - NU], ub = q < NU ? c.u_max[q] : c.x_max[q - NU];
                zlq[q] = lb > -MMPC_INF_THRESH ? zl[q * I] : 0.0;
                zuq[q] = ub < MMPC_INF_THRESH ? zu[q * I] : 0.0;
            }
        }
    }
    if (keep_lam) {  // forward sweep, instance that no longer iterates: LAMN must hold its multipliers when LAM and LAMN swap
        double lk[NX];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) lk[a] = c.LAM[((size_t)k * NX + a) * I];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) c.LAMN[((size_t)k * NX + a) * I] = lk[a];
    }
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        x[a] = fma(alpha, d[a * I], z[a * I]);
        xn[a] = fma(alpha, d[(S + a) * I], z[(S + a) * I]);
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        u[a] = fma(alpha, d[(NX + a) * I], z[(NX + a) * I]);
        v[a] = k ? fma(alpha, d[(NX + a - S) * I], z[(NX + a - S) * I]) : c.uinit[a];
    }
    if (!(C::BD && bd.full)) {
        c.template step<0>(x, u, nullptr, nullptr, nullptr, F, nullptr, nullptr, nullptr);
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            const double e = (c.node_cost ? xn[a] : F[a]) - r[a];
            f1 = fma(c.Q[a] * e, e, f1);
            th1 += fabs(F[a] - xn[a]);
        }
        MMPC_UNROLL for (int a = 0; a < NU; ++a) {
            const double dd = u[a] - v[a];
            f1 = fma(c.R[a] * dd, dd, f1);
            f1 = fma(c.Rm[a] * u[a], u[a], f1);
        }
    }
    if constexpr (C::BD) {
        if (bd.full) {  // multiplier steps and step-length limits from the iterate and the step
            double* dzl = c.DZL + (size_t)k * S * I;
            double* dzu = c.DZU + (size_t)k * S * I;
            MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                double dl, du_;
                t3_bound_step(z[(NX + a) * I], d[(NX + a) * I], c.u_min[a], c.u_max[a], zlq[a], zuq[a], bd.mu, bd.tau, dl, du_, bd.amax, bd.az);
                if (c.u_min[a] > -MMPC_INF_THRESH) dzl[(NX + a) * I] = dl;
                if (c.u_max[a] < MMPC_INF_THRESH) dzu[(NX + a) * I] = du_;
            }
            MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                double dl, du_;
                t3_bound_step(z[(S + a) * I], d[(S + a) * I], c.x_min[a], c.x_max[a], zlq[NU + a], zuq[NU + a], bd.mu, bd.tau, dl, du_, bd.amax, bd.az);
                if (c.x_min[a] > -MMPC_INF_THRESH) dzl[(S + a) * I] = dl;
                if (c.x_max[a] < MMPC_INF_THRESH) dzu[(S + a) * I] = du_;
            }
        } else {
            MMPC_UNROLL for (int a = 0; a < NU; ++a) bd.bar1 += barrier_of(u[a], c.u_min[a], c.u_max[a]);
            MMPC_UNROLL for (int a = 0; a < NX; ++a) bd.bar1 += barrier_of(xn[a], c.x_min[a], c.x_max[a]);
        }
    }
}

struct T3Fwd {
    double gradf_d, lam_inf, f1, th1;
    double bar1, amax, az;  // bounded problems: barrier at the full step, step-length limits of the primal and the bound multipliers
    bool finite;
};

// forward sweep for the tile, software pipelined: while warp 0 rolls the step through chunk ch, the worker warps stage
// chunk ch+1's records (cp.async) and evaluate the merit terms of the full step on chunk ch-1.
template <class M, class C>
__device__ __noinline__ void t3_forward(const T3Ctx<M, C::I, C::MODE>& cin, double* smem, int tid, T3Fwd& o_out, const double* fg_tile, const double* kg_tile,
                                        const double* pg_tile) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    const T3Ctx<M, C::I, C::MODE> c = cin;  // register-resident copies (see t3_backward)
    T3Fwd o;
    const int N = c.N;
    const bool worker = C::is_worker(tid);
    const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
    double* red = smem + C::SM_RED;
    double dx[NX], dv[NU];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) dx[a] = 0.0;
    MMPC_UNROLL for (int a = 0; a < NU; ++a) dv[a] = 0.0;
    o.gradf_d = o.lam_inf = o.f1 = o.th1 = o.bar1 = 0.0;
    o.amax = o.az = 1.0;
    o.finite = true;
    const int nch = (N + C::CF - 1) / C::CF;
    auto stage_chunk = [&](int ch) {  // [node][field][instance] blocks are contiguous in the tile arrays
        const int k0 = ch * C::CF, kn = min(C::CF, N - k0);
        double* buf = smem + C::SM_REC + (ch & 1) * C::BUF;
        for (int s = 0; s < kn; ++s) {
            const double* fg = fg_tile + (size_t)(k0 + s) * C::FG * I;
            const double* kg = kg_tile + (size_t)(k0 + s) * C::KG * I;
            const double* pg = pg_tile + (size_t)(k0 + s) * C::PG * I;
            double* dst = buf + (size_t)s * C::RF * I;
            for (int e = w * 2; e < C::FG * I; e += C::WORKERS * 2) cp_async16(dst + e, fg + e);
            for (int e = w * 2; e < C::KG * I; e += C::WORKERS * 2) cp_async16(dst + C::RF_K * I + e, kg + e);
            for (int e = w * 2; e < C::PG * I; e += C::WORKERS * 2) cp_async16(dst + C::RF_P * I + e, pg + e);
        }
    };
    double f1 = 0.0, th1 = 0.0;  // worker-local merit sums
    // does this thread's instance still iterate?  lam+ goes to LAMN, which swaps with LAM when the step is accepted: the serial
    // lane writes it for the instances that iterate, the merit worker of a node copies the old multipliers for the others
    const bool act = smem[C::SM_INST + 2 * I + (worker ? wi : (tid < I ? tid : 0))] != 0.0;
    T3TrialBD bd;
    bd.bar1 = 0.0; bd.amax = bd.az = 1.0; bd.full = true;
    bd.mu = C::BD ? smem[C::SM_INST + 3 * I + (worker ? wi : 0)] : 0.0;
    bd.tau = fmax(0.99, 1.0 - bd.mu);  // IPOPT tau_min = 0.99 [upstream]
    auto trial_chunk = [&](int ch) {  // workers with node slot < CF take part
        const int k0 = ch * C::CF, kn = min(C::CF, N - k0);
        if (wj < kn) t3_trial_node<M, C>(c, k0 + wj, 1.0, f1, th1, bd, !act);
    };
#ifdef MMPC_T3_PROFILE_SWEEPS
    long long t_work = 0, t_all0 = clock64();
#endif
    if (worker) {
        stage_chunk(0);
        cp_async_wait_all();
    }
    __syncthreads();
    for (int ch = 0; ch < nch; ++ch) {
#ifdef MMPC_T3_PROFILE_SWEEPS
        const long long t0 = clock64();
#endif
        if (worker) {
            if (ch + 1 < nch) stage_chunk(ch + 1);
            if (wj < C::CF && ch * C::CF + wj < N) {  // next step's merit evaluation: iterate rows and reference -> L1
                const int kp = ch * C::CF + wj;
                MMPC_UNROLL for (int a = 0; a < S + NX; ++a) t3_prefetch(c.V + ((size_t)kp * S + a) * I);
                t3_prefetch(c.traj + (size_t)kp * NX);
            }
            if (ch > 0) trial_chunk(ch - 1);
            cp_async_wait_all();
        } else if (tid < I) {
            const int k0 = ch * C::CF, kn = min(C::CF, N - k0);
            const double* rec = smem + C::SM_REC + (ch & 1) * C::BUF;
            for (int s = 0; s < kn; ++s) {
                const int k = k0 + s;
                const double* const recp = rec + (size_t)s * C::RF * I + tid;
#define T3_F(f) recp[(f) * I]
                double du[NU], dxn[NX];
                MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                    double t = T3_F(C::RF_K + C::KG_K0 + a);
                    MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(T3_F(C::RF_K + C::KG_KX + a * NX + q), dx[q], t);
                    MMPC_UNROLL for (int q = 0; q < NU; ++q) t = fma(T3_F(C::RF_K + C::KG_KV + a * NU + q), dv[q], t);
                    du[a] = t;
                }
                MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                    double t = T3_F(C::FG_C + a);
                    MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(a, q)) t = fma(T3_F(C::FG_FX + T3_FXSLOT(a * NX + q)), dx[q], t);
                    MMPC_UNROLL for (int q = 0; q < NU; ++q) if (T3_FUNZ(a, q)) t = fma(T3_F(C::FG_FU + T3_FUSLOT(a * NU + q)), du[q], t);
                    dxn[a] = t;
                }
                // directional derivative of the objective: q~ . (dx, dv, du)
                MMPC_UNROLL for (int a = 0; a < NX; ++a) o.gradf_d = fma(T3_F(C::FG_Q + a), dx[a], o.gradf_d);
                MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                    o.gradf_d = fma(T3_F(C::FG_Q + NX + a), dv[a], o.gradf_d);
                    o.gradf_d = fma(T3_F(C::FG_Q + S + a), du[a], o.gradf_d);
                }
                if constexpr (C::BD != 0) {  // barrier gradient of x_{k+1} . dx_{k+1}
                    MMPC_UNROLL for (int a = 0; a < NX; ++a) o.gradf_d = fma(T3_F(C::FG_GN + a), dxn[a], o.gradf_d);
                }
                MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                    double t = T3_F(C::RF_P + C::PG_X + a);
                    MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(T3_F(C::RF_P + C::PG_XX + trisym(a, q, NX)), dxn[q], t);
                    MMPC_UNROLL for (int q = 0; q < NU; ++q) t = fma(T3_F(C::RF_P + C::PG_XV + a * NU + q), du[q], t);
                    if (act) c.LAMN[((size_t)k * NX + a) * I] = t;
                    o.lam_inf = fmax(o.lam_inf, fabs(t));
                    o.finite = o.finite && isfinite(t) && isfinite(dxn[a]);
                }
#undef T3_F
                MMPC_UNROLL for (int a = 0; a < NU; ++a) c.DV[((size_t)k * S + NX + a) * I] = du[a];
                MMPC_UNROLL for (int a = 0; a < NX; ++a) c.DV[((size_t)(k + 1) * S + a) * I] = dxn[a];
                MMPC_UNROLL for (int a = 0; a < NX; ++a) dx[a] = dxn[a];
                MMPC_UNROLL for (int a = 0; a < NU; ++a) dv[a] = du[a];
            }
        }
#ifdef MMPC_T3_PROFILE_SWEEPS
        t_work += clock64() - t0;
#endif
        __syncthreads();  // the step of chunk ch is visible to the workers; buffer (ch+1)&1 is staged
    }
#ifdef MMPC_T3_PROFILE_SWEEPS
    if (blockIdx.x == 0 && (tid == 0 || tid == 32))
        printf("[t3 forward ] tid %d busy %lld of %lld cycles (%d chunks of %d)\n", tid, t_work, clock64() - t_all0, nch, C::CF);
#endif
    if (worker) {
        trial_chunk(nch - 1);
        red[(0 * C::CB + wj) * I + wi] = f1;
        red[(1 * C::CB + wj) * I + wi] = th1;
        if constexpr (C::BD) {
            red[(2 * C::CB + wj) * I + wi] = bd.bar1;
            red[(3 * C::CB + wj) * I + wi] = bd.amax;
            red[(4 * C::CB + wj) * I + wi] = bd.az;
        }
    }
    __syncthreads();
    if (tid < I) MMPC_UNROLL for (int s = 0; s < C::CF; ++s) {
        o.f1 += red[(0 * C::CB + s) * I + tid];
        o.th1 += red[(1 * C::CB + s) * I + tid];
        if constexpr (C::BD) {
            o.bar1 += red[(2 * C::CB + s) * I + tid];
            o.amax = fmin(o.amax, red[(3 * C::CB + s) * I + tid]);
            o.az = fmin(o.az, red[(4 * C::CB + s) * I + tid]);
        }
    }
    __syncthreads();
    o_out = o;
}

// merit of z + alpha_i dz for every instance of the tile (alpha per instance in shared memory)
template <class M, class C>
__device__ __noinline__ void t3_trial(const T3Ctx<M, C::I, C::MODE>& cin, double* smem, int tid, double& f1, double& th1, double& bar1) {
    constexpr int I = C::I;
    const T3Ctx<M, C::I, C::MODE> c = cin;
    const bool worker = C::is_worker(tid);
    const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
    double* red = smem + C::SM_RED;
    if (worker) {
        const double alpha = smem[C::SM_INST + I + wi];
        double a1 = 0.0, a2 = 0.0;
        T3TrialBD bd;
        bd.bar1 = 0.0; bd.amax = bd.az = 1.0; bd.mu = bd.tau = 0.0; bd.full = false;
        for (int k = wj; k < c.N; k += C::CB) t3_trial_node<M, C>(c, k, alpha, a1, a2, bd);
        red[(0 * C::CB + wj) * I + wi] = a1;
        red[(1 * C::CB + wj) * I + wi] = a2;
        if constexpr (C::BD) red[(2 * C::CB + wj) * I + wi] = bd.bar1;
    }
    __syncwarp();  // as in t3_residuals
    __syncthreads();
    f1 = 0.0; th1 = 0.0; bar1 = 0.0;
    if (tid < I) MMPC_UNROLL for (int s = 0; s < C::CB; ++s) {
        f1 += red[(0 * C::CB + s) * I + tid];
        th1 += red[(1 * C::CB + s) * I + tid];
        if constexpr (C::BD) bar1 += red[(2 * C::CB + s) * I + tid];
    }
    __syncthreads();
}

template <class M, int I_, int CB_, int CF_, int SK_, int RK_>
__global__ void __launch_bounds__(T3<M, I_, CB_, CF_, SK_, RK_>::THREADS) solve_kernel_t3(mmpc_problem P, mmpc_workspace W, mmpc_solve_io io) {
    using C = T3<M, I_, CB_, CF_, SK_, RK_>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    extern __shared__ __align__(16) double smem[];
#ifdef MMPC_T3_PROFILE_CTAS
    unsigned long long gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
#ifdef MMPC_T3_PROFILE
    const long long t_entry = clock64();
    long long t_bwd[8] = {0}, t_fwd[8] = {0}, t_post[8] = {0}, t_res[8] = {0}, t_glue[4] = {0}, t_iter[8] = {0}, t_it0 = 0, ts[10] = {0};
#endif
    const int tid = threadIdx.x, tile = io.b0 / C::I + blockIdx.x;  // b0 is a multiple of 32*I-compatible tile starts
    const bool worker = C::is_worker(tid);
    // a thread's instance: serial lane tid -> instance tid; worker w -> instance w % I (node slot w / I)
    const int i = worker ? C::widx(tid) % I : (tid < I ? tid : 0);
    const int j = worker ? C::slot_of(C::widx(tid) / I) : 0;
    const int N = P.N, NV = NX * (N + 1) + NU * N, Mc = N * NX;
    const int b = tile * I + i;                 // workspace slot (arrays are sized for the batch padded to 32)
    const bool valid = b >= io.b0 && b < io.b1;
    const int bi = valid ? b : io.b1 - 1;       // surplus slots read the last instance's inputs and never write outputs

    T3Ctx<M, I, C::MODE> c;
    c.N = N;
    c.h = P.h; c.is_linear = P.is_linear; c.substeps = P.substeps; c.gauss_newton = P.gauss_newton; c.node_cost = P.node_cost;
    // the iterate lives tile-wise, [element][instance]: every access of a warp is a run of full sectors whatever I is
    c.V = W.V2 + (size_t)tile * I * NV + i; c.LAM = W.LAM2 + (size_t)tile * I * Mc + i;
    c.DV = W.DV2 + (size_t)tile * I * NV + i; c.LAMN = W.LAMN2 + (size_t)tile * I * Mc + i;
    c.LINp = W.LIN2 + (size_t)b * C::LIN;
    if constexpr (C::BD != 0) {
        MMPC_UNROLL for (int a = 0; a < NX; ++a) { c.x_min[a] = P.x_min[a]; c.x_max[a] = P.x_max[a]; }
        MMPC_UNROLL for (int a = 0; a < NU; ++a) { c.u_min[a] = P.u_min[a]; c.u_max[a] = P.u_max[a]; }
        c.ZL = W.ZL2 + (size_t)tile * I * NV + i; c.ZU = W.ZU2 + (size_t)tile * I * NV + i;
        c.DZL = W.DZL2 + (size_t)tile * I * NV + i; c.DZU = W.DZU2 + (size_t)tile * I * NV + i;
    }
    const double *in_state = io.state, *in_control = io.control;
    c.traj = io.traj + (size_t)bi * N * NX;
    if (io.traj_stage) {  // zero-copy inputs (uniform over the grid): the tile's rows are brought on the device first
        t3_stage_inputs(io.state, io.control, io.traj, io.state_stage, io.control_stage, io.traj_stage, max(tile * I, io.b0),
                        min(tile * I + I, io.b1), NX, NU, N);
        in_state = io.state_stage; in_control = io.control_stage;
        c.traj = io.traj_stage + (size_t)bi * N * NX;
    }
    double* fg_tile = W.FG2 + (size_t)tile * I * N * C::FG;
    double* kg_tile = W.KG2 + (size_t)tile * I * N * C::KG;
    double* pg_tile = W.PG2 + (size_t)tile * I * N * C::PG;
    MMPC_UNROLL for (int a = 0; a < NX; ++a) c.Q[a] = W.Qw[(size_t)a * W.Bp + bi];
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        c.R[a] = W.Rw[(size_t)a * W.Bp + bi];
        c.Rm[a] = W.Rmw[(size_t)a * W.Bp + bi];
        c.uinit[a] = in_control[(size_t)bi * NU + a];
    }
    // ---- start point (ModelControl.cpp:28-50,68 cold; :161 warm), x_0 pinned (:144-145); the C workers of an instance help ----
    const bool warm = io.warm_start && W.have_warm;
    if (worker) {
        if (!warm) for (int e = NX + j; e < NV; e += C::CB) c.V[(size_t)e * I] = 0.0;
        if (!(warm && io.warm_start_duals)) for (int e = j; e < Mc; e += C::CB) c.LAM[(size_t)e * I] = 0.0;
        for (int e = j; e < NX; e += C::CB) {
            c.V[(size_t)e * I] = in_state[(size_t)bi * NX + e];
            c.DV[(size_t)e * I] = 0.0;
        }
    }
    if constexpr (C::BD) {
        // push the start into the interior (IPOPT bound_push = bound_frac = 1e-2) and set z = 1 [upstream]; x_0 is pinned
        if (worker) for (int e = NX + j; e < NV; e += C::CB) {
            const int r = e % S;
            const double lb = r < NX ? P.x_min[r] : P.u_min[r - NX], ub = r < NX ? P.x_max[r] : P.u_max[r - NX];
            const bool hl = lb > -MMPC_INF_THRESH, hu = ub < MMPC_INF_THRESH;
            double v = c.V[(size_t)e * I];
            if (hl) { double p = 1e-2 * fmax(1.0, fabs(lb)); if (hu) p = fmin(p, 1e-2 * (ub - lb)); v = fmax(v, lb + p); }
            if (hu) { double p = 1e-2 * fmax(1.0, fabs(ub)); if (hl) p = fmin(p, 1e-2 * (ub - lb)); v = fmin(v, ub - p); }
            c.V[(size_t)e * I] = v;
            c.ZL[(size_t)e * I] = hl ? 1.0 : 0.0;
            c.ZU[(size_t)e * I] = hu ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    if (P.is_linear && tid < I) {
        double x0[NX], xd[NX], A[NX * NX], Bm[NX * NU];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) x0[a] = c.V[(size_t)a * I];
        M::fAB(x0, c.uinit, xd, A, Bm);
        MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) c.LINp[a] = A[a];
        MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) c.LINp[NX * NX + a] = Bm[a];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) c.LINp[NX * NX + NX * NU + a] = xd[a];
    }
    __syncthreads();

    // plain local copies for the element-wise loops below (the struct `c` is passed by reference to the sweeps, so it lives
    // in local memory; these do not)
    double* const Vp = c.V; const double* const DVp = c.DV;
    // per-instance solver state lives in the serial lane (tid < I) of the instance
    double nu_merit = 1.0, delta_last = 0.0, err = INFINITY, fval = 0.0;
    int it = 0, it_done = 0, status = MMPC_SOLVE_MAX_ITER, ls_total = 0, reg_total = 0;
    bool active = valid && tid < I;
    double* sh_alpha = smem + C::SM_INST + I;      // per-instance step length
    double* sh_flag = smem + C::SM_INST + 2 * I;   // per-instance "apply the step" flag
    const int nfree = N * S, ncon = N * NX;
#ifdef MMPC_T3_PROFILE
    const long long t_kernel0 = clock64();
#endif
    double mu = C::BD ? io.mu_init : 0.0, barrier0 = 0.0;
    bool expect_converged = false;  // tile-uniform, unbounded problems (see the loop)
    double* sh_mu = smem + C::SM_INST + 3 * I;  // per-instance barrier parameter
    double* sh_az = smem + C::SM_INST + 4 * I;  // per-instance step length of the bound multipliers
    for (it = 0;; ++it) {
#ifdef MMPC_T3_PROFILE
        { const long long tn = clock64(); if (it > 0 && it <= 8) t_iter[it - 1] = tn - t_it0; t_it0 = tn; if (it == 1) ts[0] = tn; if (it == 2) ts[9] = tn; }
#endif
        T3Norms nrm;
        double delta = 0.0;
        bool ok = true;
        if constexpr (C::BD) {
            // residuals first: the barrier parameter is chosen from them (T1's bounded iteration, mmpc_solve_t1.cuh)
            T3NormsBD nb;
#ifdef MMPC_T3_PROFILE
            const long long tr0 = clock64();
#endif
            t3_residuals<M, C>(c, smem, tid, nb);
#ifdef MMPC_T3_PROFILE
            if (it < 8) t_res[it] = clock64() - tr0;
#endif
            nrm.dual_inf = nb.dual_inf; nrm.prim_inf = nb.prim_inf; nrm.theta = nb.theta; nrm.fval = nb.fval; nrm.lam_l1 = nb.lam_l1;
            barrier0 = nb.barrier;
            if (active) {
                // IPOPT scaled optimality error (Waechter & Biegler 2006 eq. 5-6) [upstream]: s_max = 100
                const double sd = fmax(100.0, (nb.lam_l1 + nb.z_l1) / (double)(nfree + ncon)) / 100.0;
                const double sc = fmax(100.0, nb.z_l1 / (double)nfree) / 100.0;
                const double e_dp = fmax(nb.dual_inf / sd, nb.prim_inf);
                err = fmax(e_dp, nb.cmax / sc);
                fval = nb.fval;
                it_done = it;
                if (!isfinite(err)) { status = MMPC_SOLVE_NUMERICAL; active = false; }
                else if (err <= io.tol) { status = MMPC_SOLVE_CONVERGED; active = false; }
                else if (it >= io.max_iter) { status = MMPC_SOLVE_MAX_ITER; active = false; }
                else {
                    // monotone barrier update (kappa_eps = 10, kappa_mu = 0.2, theta_mu = 1.5) [upstream]
                    while (mu > io.tol / 10.0 && fmax(e_dp, fmax(nb.cmax - mu, mu - nb.cmin) / sc) <= 10.0 * mu)
                        mu = fmax(io.tol / 10.0, fmin(0.2 * mu, mu * sqrt(mu)));
                }
            }
            if (tid < I) sh_mu[tid] = mu;
            if (!__syncthreads_or(active)) break;
            T3Norms unused;
#ifdef MMPC_T3_PROFILE
            const long long tp0 = clock64();
#endif
            ok = t3_backward<M, C>(c, smem, tid, delta, unused, fg_tile, kg_tile, pg_tile);
#ifdef MMPC_T3_PROFILE
            if (it < 8) t_bwd[it] = clock64() - tp0;
#endif
        } else {
        if (expect_converged) {
            // every instance of the tile that still iterates took a full Newton step from a small KKT error: the new iterate
            // is expected to pass the convergence test, which needs first derivatives only.  Test it with the cheap residual
            // sweep; the full sweep below (second derivatives, Riccati recursion, records) runs only if some instance fails.
            T3NormsBD nb;
            t3_residuals<M, C>(c, smem, tid, nb);
            if (active) {
                const double sd = fmax(100.0, nb.lam_l1 / (double)(nfree + ncon)) / 100.0;
                const double e1 = fmax(nb.dual_inf / sd, nb.prim_inf);
                if (isfinite(e1) && e1 <= io.tol) {
                    err = e1; fval = nb.fval; it_done = it;
                    status = MMPC_SOLVE_CONVERGED;
                    active = false;
                }
            }
            if (!__syncthreads_or(active)) break;
        }
#ifdef MMPC_T3_PROFILE
        const long long tp0 = clock64();
        if (it == 1) ts[1] = tp0;
#endif
        ok = t3_backward<M, C>(c, smem, tid, delta, nrm, fg_tile, kg_tile, pg_tile);
#ifdef MMPC_T3_PROFILE
        if (it == 1) ts[2] = clock64();
        if (it < 8) t_bwd[it] = clock64() - tp0;
#endif
        if (active) {
            const double sd = fmax(100.0, nrm.lam_l1 / (double)(nfree + ncon)) / 100.0;  // IPOPT s_d, s_max = 100 [upstream]
            err = fmax(nrm.dual_inf / sd, nrm.prim_inf);
            fval = nrm.fval;
            it_done = it;
            if (!isfinite(err)) { status = MMPC_SOLVE_NUMERICAL; active = false; }
            else if (err <= io.tol) { status = MMPC_SOLVE_CONVERGED; active = false; }
            else if (it >= io.max_iter) { status = MMPC_SOLVE_MAX_ITER; active = false; }
        }
        }
#ifdef MMPC_T3_PROFILE
        const long long tg0 = clock64();
        if (it == 1) ts[3] = tg0;
#endif
        if (!__syncthreads_or(active)) break;
#ifdef MMPC_T3_PROFILE
        const long long tg1 = clock64();
#endif
        int tries = 0;
        while (__syncthreads_or(active && !ok)) {  // inertia correction for the instances that need it
            if (active && !ok) {
                delta = (delta == 0.0) ? (delta_last > 0.0 ? fmax(1e-20, delta_last / 3.0) : 1e-4) : delta * 8.0;
                ++reg_total;
            }
            T3Norms dummy;
            const bool ok2 = t3_backward<M, C>(c, smem, tid, delta, dummy, fg_tile, kg_tile, pg_tile);
            if (active && !ok) ok = ok2;
            if (++tries >= 24) {
                if (active && !ok) { status = MMPC_SOLVE_NUMERICAL; active = false; }
                break;
            }
        }
        if (delta > 0.0) delta_last = delta;
        if (tid < I) sh_flag[tid] = active ? 1.0 : 0.0;  // which instances still iterate: read by the forward sweep
#ifdef MMPC_T3_PROFILE
        const long long tg2 = clock64();
#endif
        if (!__syncthreads_or(active)) break;
#ifdef MMPC_T3_PROFILE
        if (it == 1 && blockIdx.x == 0 && tid == 0) { t_glue[0] = tg1 - tg0; t_glue[1] = tg2 - tg1; t_glue[2] = clock64() - tg2; }
#endif
        T3Fwd fo;
#ifdef MMPC_T3_PROFILE
        const long long tp1 = clock64();
        if (it == 1) { ts[4] = tg2; ts[5] = tp1; }
#endif
        t3_forward<M, C>(c, smem, tid, fo, fg_tile, kg_tile, pg_tile);
#ifdef MMPC_T3_PROFILE
        const long long tp2 = clock64();
        if (it == 1) ts[6] = tp2;
        if (it < 8) t_fwd[it] = tp2 - tp1;
#endif
        if (active && !(fo.finite && isfinite(fo.f1))) { status = MMPC_SOLVE_NUMERICAL; active = false; }
        nu_merit = fmax(nu_merit, 1.1 * fo.lam_inf + 1e-8);
        const double m0 = nrm.fval + mu * barrier0 + nu_merit * nrm.theta;  // mu = 0 without bounds
        const double D = fo.gradf_d - nu_merit * nrm.theta;
        double alpha = C::BD ? fo.amax : 1.0;  // fraction to the boundary
        double f1 = fo.f1, th1 = fo.th1, bar1 = fo.bar1;
        int ls = 0;
        if constexpr (C::BD) {
            // the forward sweep of a bounded problem computes no merit terms (alpha_max is known only when it ends)
            if (tid < I) sh_alpha[tid] = alpha;
            __syncthreads();
            t3_trial<M, C>(c, smem, tid, f1, th1, bar1);
        }
        bool need = active && !(f1 + mu * bar1 + nu_merit * th1 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0));
        while (__syncthreads_or(need)) {
            if (need) { alpha *= 0.5; ++ls_total; }
            if (tid < I) sh_alpha[tid] = alpha;
            __syncthreads();
            double f2, t2, b2;
            t3_trial<M, C>(c, smem, tid, f2, t2, b2);
            if (need) need = !(f2 + mu * b2 + nu_merit * t2 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0)) && (++ls < 40);
        }
        // accept the step.  Multipliers: LAMN already holds lam+ (instances that iterate) or the old values (the others), so
        // a full step only swaps the two arrays; a shortened step (some instance of the tile) blends in place first.
        if (tid < I) {
            sh_alpha[tid] = active ? alpha : 0.0;
            if constexpr (C::BD) sh_az[tid] = fo.az;
        }
        const bool shortened = __syncthreads_or(active && alpha != 1.0);
        if (worker) {
            const double al = sh_alpha[i];
            constexpr int UB = 8;
            if (al != 0.0) for (int e0 = NX + j; e0 < NV; e0 += C::CB * UB) {  // the C workers of an instance update its iterate together, loads batched
                double vz[UB], vd[UB];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < NV) { vz[q] = Vp[(size_t)(e0 + q * C::CB) * I]; vd[q] = DVp[(size_t)(e0 + q * C::CB) * I]; }
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < NV) Vp[(size_t)(e0 + q * C::CB) * I] = fma(al, vd[q], vz[q]);
            }
            if (shortened) for (int e0 = j; e0 < Mc; e0 += C::CB * UB) {
                double vl[UB], vn[UB];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < Mc) { vl[q] = c.LAM[(size_t)(e0 + q * C::CB) * I]; vn[q] = c.LAMN[(size_t)(e0 + q * C::CB) * I]; }
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < Mc) c.LAMN[(size_t)(e0 + q * C::CB) * I] = (al != 0.0) ? fma(al, vn[q] - vl[q], vl[q]) : vl[q];
            }
        }
        if constexpr (C::BD != 0) {
            // bound multipliers: z += alpha_z dz, then IPOPT's kappa_sigma safeguard at the new point (T1).  Node by node, so
            // that the loads of one node are in flight together; only variables with a finite bound are touched.
            __syncthreads();  // the iterate above was updated element-wise by other threads of the instance
            if (worker && sh_alpha[i] != 0.0) {
                const double az = sh_az[i], mu_i = sh_mu[i];
                for (int k = j; k < N; k += C::CB) {
                    // q < NU: u_k, element k S + NX + q;  q >= NU: x_{k+1}, element (k+1) S + (q - NU) = k S + NX + q as well
                    double v[S], zl[S], zu[S], dl[S], du[S];
                    MMPC_UNROLL for (int q = 0; q < S; ++q) {
                        const double lb = q < NU ? P.u_min[q] : P.x_min[q - NU], ub = q < NU ? P.u_max[q] : P.x_max[q - NU];
                        const size_t e = ((size_t)k * S + NX + q) * I;
                        if (lb > -MMPC_INF_THRESH || ub < MMPC_INF_THRESH) v[q] = c.V[e];
                        if (lb > -MMPC_INF_THRESH) { zl[q] = c.ZL[e]; dl[q] = c.DZL[e]; }
                        if (ub < MMPC_INF_THRESH) { zu[q] = c.ZU[e]; du[q] = c.DZU[e]; }
                    }
                    MMPC_UNROLL for (int q = 0; q < S; ++q) {
                        const double lb = q < NU ? P.u_min[q] : P.x_min[q - NU], ub = q < NU ? P.u_max[q] : P.x_max[q - NU];
                        const size_t e = ((size_t)k * S + NX + q) * I;
                        if (lb > -MMPC_INF_THRESH) c.ZL[e] = z_safeguard(fma(az, dl[q], zl[q]), v[q] - lb, mu_i);
                        if (ub < MMPC_INF_THRESH) c.ZU[e] = z_safeguard(fma(az, du[q], zu[q]), ub - v[q], mu_i);
                    }
                }
            }
        }
        __syncthreads();
        { double* t = c.LAM; c.LAM = c.LAMN; c.LAMN = t; }
        // Is the next iterate expected to meet tol?  Then its test is done by the cheap residual sweep first.
        //   exact Hessian: Newton converges quadratically (KKT error 0.3 -> 3e-6, 0.05 -> 6e-9 on the bench workload): error < 0.1;
        //   linearised model: the problem is a QP, one full step solves it;
        //   RK4 (Gauss-Newton, linear rate): not predicted -- extrapolating the observed contraction was measured a loss at
        //   one sub-step (1.61 -> 1.71 ms) and a small gain at four (3.58 -> 3.44 ms).
        if constexpr (C::BD == 0) {
            const bool newton = C::RK == 0 || !P.gauss_newton;
            const bool expect = newton && alpha == 1.0 && (P.is_linear || err <= T3_EXPECT_CONVERGED_BELOW);
            expect_converged = !__syncthreads_or(active && !expect);
        }
#ifdef MMPC_T3_PROFILE
        if (it < 8) t_post[it] = clock64() - tp2;
        if (it == 1) ts[7] = clock64();
#endif
    }
#ifdef MMPC_T3_PROFILE
    const long long t_loop_end = clock64();
#endif
    // ---- outputs: the reference's V layout, instance-major.  The warm start lives in LAM2: after an odd number of accepted
    // steps the multipliers are copied back there ----
    double* const Lcan = W.LAM2 + (size_t)tile * I * Mc + i;
    const bool back = c.LAM != Lcan;  // tile-uniform
#ifdef MMPC_WIDE_TILES_NO_HOSTOUT
    constexpr bool HOST_OUT = I <= 8;  // wide tiles: the ABI layer fetches the outputs with a copy kernel behind this one
#else
    constexpr bool HOST_OUT = true;
#endif
    const bool host_out = HOST_OUT && io.host_out != 0;  // zero-copy outputs: the transposing epilogue runs instead of the plain one (see T4's note)
    if (worker && !host_out) {
        const double* LAMp = c.LAM;
        constexpr int UB = 8;  // batched so that the loads of one batch are in flight together
        if (valid) {
            double* Vo = io.V_out + (size_t)b * NV;
            for (int e0 = j; e0 < NV; e0 += C::CB * UB) {
                double t[UB];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < NV) t[q] = Vp[(size_t)(e0 + q * C::CB) * I];
                MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < NV) __stcs(&Vo[e0 + q * C::CB], t[q]);
            }
        }
        if ((valid && io.lam_out) || back) for (int e0 = j; e0 < Mc; e0 += C::CB * UB) {
            double t[UB];
            MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < Mc) t[q] = LAMp[(size_t)(e0 + q * C::CB) * I];
            MMPC_UNROLL for (int q = 0; q < UB; ++q) if (e0 + q * C::CB < Mc) {
                if (valid && io.lam_out) __stcs(&io.lam_out[(size_t)b * Mc + e0 + q * C::CB], t[q]);
                if (back) Lcan[(size_t)(e0 + q * C::CB) * I] = t[q];
            }
        }
    }
    // statistics: to the caller's array, or -- zero-copy outputs -- to shared memory behind the transpose buffer for the epilogue
    if (tid < I && io.stats && (valid || host_out)) {
        mmpc_stats s;
        s.objective = fval; s.kkt_error = err; s.mu = mu; s.iterations = it_done; s.status = status;
        s.ls_steps = ls_total; s.reg_steps = reg_total;
        mmpc_stats* dst = host_out ? reinterpret_cast<mmpc_stats*>(smem + I * 65) + tid : io.stats + b;
        *dst = s;
    }
    if constexpr (HOST_OUT) if (host_out)
        t3_outputs_to_host_s(smem, io.V_out, io.lam_out, io.stats, io.b1, tile, I, NV, Mc, Vp - i, c.LAM - i, back ? Lcan - i : nullptr);
#ifdef MMPC_T3_PROFILE_CTAS
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        printf("[t3 cta] %d sm %u start %llu end %llu\n", (int)blockIdx.x, smid, gt0, gt1);
    }
#endif
#ifdef MMPC_T3_PROFILE
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        const long long t_end = clock64();
        printf("[t3 kernel] prologue %lld  loop %lld  epilogue %lld cycles, %d iterations\n", t_kernel0 - t_entry, t_loop_end - t_kernel0, t_end - t_loop_end, it);
        for (int q = 0; q <= it && q < 8; ++q) printf("[t3 it %d] residuals %lld  backward %lld  forward %lld  line search + step %lld  whole iteration %lld\n", q, t_res[q], t_bwd[q], t_fwd[q], t_post[q], t_iter[q]);
        printf("[t3 stamps it 1] top->tp0 %lld  bwd %lld  ->tg0 %lld  votes %lld  ->tp1 %lld  fwd %lld  post %lld  ->next top %lld\n", ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], ts[6] - ts[5], ts[7] - ts[6], ts[9] - ts[7]);
        printf("[t3 glue it 1] first vote %lld  inertia loop %lld  second vote %lld\n", t_glue[0], t_glue[1], t_glue[2]);
    }
#endif
}

template <class M>
struct LaunchT3 {
    template <int I_, int CB_, int CF_, int SK_ = 0, int RK_ = 0>
    static cudaError_t go(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        using C = T3<M, I_, CB_, CF_, SK_, RK_>;
        // the opt-in to more than 48 KB of dynamic shared memory is a per-device function attribute (one process may drive
        // several GPUs through several handles); set once per device, atomically because handles may solve from several threads
        static std::atomic<unsigned long long> attr_done{0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (C::SMEM_BYTES > 48 * 1024 && !((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ULL)) {
            cudaError_t e = cudaFuncSetAttribute(solve_kernel_t3<M, I_, CB_, CF_, SK_, RK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            attr_done.fetch_or(1ULL << (dev & 63), std::memory_order_release);
        }
        const int t0 = io.b0 / C::I, t1 = (io.b1 + C::I - 1) / C::I;  // tiles touched by [b0, b1)
        const int grid = t1 - t0;
        solve_kernel_t3<M, I_, CB_, CF_, SK_, RK_><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(P, W, io);
        return cudaGetLastError();
    }
    // Configurations (tile width I, backward chunk CB, forward chunk CF), measured at N = 50 (profiles/r1_t3_configs.txt):
    //   1   8 x 4 x 4   64 threads,  41 KB, 4 CTAs/SM: lowest latency; best while the batch leaves at most one tile per SM (B <= 1184)
    //   2  32 x 7 x 5  256 threads, 182 KB, 1 CTA/SM: the serial warp is full (an fp64 instruction costs the same issue
    //                  slots for 8 active lanes as for 32) and 7 worker warps evaluate 7 nodes at a time; best beyond that
    // Tried and slower: 14 x 8 x 4, 16 x 4 x 4, 32 x 4 x 4, 32 x 7 x {4, 6}, 32 x 6 x {4, 5} with warp 4 idle (SK).
    // The tile width fixes the layout of the persistent iterate, so the choice depends only on the batch size of the
    // handle.  kernel_variant = 3 + 256*k forces configuration k.
    // The instantiations are compiled as three parts, one translation unit each (MMPC_DEFINE_SOLVE_T3 / _T3_PART in mmpc_kernels.cuh):
    //   part 0  Euler / linearised      part 1  finite bounds (interior point)      part 2  RK4 shooting
    static int part_of(const mmpc_problem& P) { return P.bounded ? 1 : ((P.integrator == MMPC_RK4 && !P.is_linear) ? 2 : 0); }
    template <int PART>
    static cudaError_t run_part(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        if (P.node_cost) return cudaErrorNotSupported;  // MMPC_COST_NODE: T4 / T1 (T3's forward sweep has no terminal objective term)
        const int forced = io.variant >> 8;
        const int cfg = forced ? forced : (W.B > 8 * 148 ? 2 : 1);
        if constexpr (PART == 0) return cfg == 2 ? go<32, 7, 5>(P, W, io, st) : go<8, 4, 4>(P, W, io, st);
        // finite bounds: interior point (Euler / linearised steps; bounded RK4 solves stay on T1); RK4 shooting: dense sensitivities.
        // Both are instantiated for the small models the tile kernels are the default for.
        if constexpr (M::NX + 2 * M::NU <= 8) {
            if constexpr (PART == 1) {
                if (P.integrator == MMPC_RK4 && !P.is_linear) return cudaErrorNotSupported;
                return cfg == 2 ? go<32, 6, 5, 0, 2>(P, W, io, st) : go<8, 4, 4, 0, 2>(P, W, io, st);
            }
            if constexpr (PART == 2) return cfg == 2 ? go<32, 7, 5, 0, 1>(P, W, io, st) : go<8, 4, 4, 0, 1>(P, W, io, st);
        }
        return cudaErrorNotSupported;
    }
};

}  // namespace mmpc

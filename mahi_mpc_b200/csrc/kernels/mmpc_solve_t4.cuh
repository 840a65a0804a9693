// mahi-mpc-b200 :: solve kernel, variant T4 -- one 512-thread CTA per tile of I instances, sized so that a batch of 4096
// fills all 148 SMs with one CTA each (I = 28 -> 147 tiles).  Same Newton / Riccati iteration as T1-T3 (same iterates),
// reorganised around what ncu showed for T3 (profiles/r1_bench_solve_t3_ncu_full.txt: 12.5 % warps active, the solve time is
// one tile's dependent chain: a 1.3 k-cycle Riccati stage per node in ONE lane, 8.5 k-cycle node evaluations in 7 warps):
//
//   * serial warps 0..3 (one warpgroup): the Riccati recursion and the linear roll-out of an instance are spread over
//     LPI = NX lanes -- lane a owns row a of the value function (Pxx, Pxv, px) and of the stage matrices; the rows meet
//     twice per stage through a per-warp shared-memory scratch (P Fx | P Fu | y, then the gain columns), __syncwarp only.
//     8 instances per warp, each warp on its own SM sub-partition (its own fp64 pipe): a stage costs ~1/3 of T3's chain.
//   * worker warps 4..15 (three warpgroups, 364 of 384 threads busy for I = 28): thread (i, j) evaluates node u = j, j+CB,
//     ... of instance i in sweep order (F, dF/dx, dF/du, exact Lagrangian stage Hessian, KKT residual pieces) -- 13 nodes
//     of every instance in flight instead of 7.
//   * the two sides are decoupled by a RING of node records in shared memory guarded by mbarriers (full / empty per slot):
//     no CTA-wide barrier inside a sweep, the recursion trails the evaluation by a few nodes.
//   * forward sweep: the per-node record (Fx, Fu, c, q~ | gains | value-function rows: one contiguous block per node in the
//     tile-major global layout) is brought into the ring by ONE cp.async.bulk (TMA) per node, completion on the slot's
//     mbarrier (complete_tx); the worker threads evaluate the merit terms of the trial point node by node as soon as the
//     roll-out has passed it.
//   * register budgets by role (setmaxnreg): the serial warpgroup gives registers to the evaluators.
// Euler / linearised shooting, no finite bounds (those stay on T3 / T1).  Replaces IPOPT+MUMPS at ModelControl.cpp:159.
#pragma once
#include <cstdint>
#if defined(MMPC_T4_DEBUG) || defined(MMPC_T4_PROFILE)
#include <cstdio>
#endif
#ifdef MMPC_T4_PROFILE
#define T4_CLK() clock64()
#else
#define T4_CLK() 0LL
#endif

namespace mmpc {

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ unsigned t4_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t4_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t4_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void t4_mbar_inval(unsigned long long* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(t4_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t4_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t4_smem_u32(bar)) : "memory");
}
// arrival that orders nothing: for "I have finished READING this slot" (the reads fed arithmetic that has completed) -- a
// releasing arrival would first wait for the thread's outstanding global stores (gains, value-function rows)
__device__ __forceinline__ void t4_mbar_arrive_relaxed(unsigned long long* bar) {
#ifdef MMPC_T4_RELEASE_ARRIVE
    t4_mbar_arrive(bar);
#else
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(t4_smem_u32(bar)) : "memory");
#endif
}
__device__ __forceinline__ void t4_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t4_smem_u32(bar)), "r"(bytes) : "memory");
}
#ifdef MMPC_T4_DEBUG
// developer build: a wait that gives up after ~1e6 polls, says where, and lets the kernel run to its end (garbage results)
__device__ __forceinline__ void t4_mbar_wait(unsigned long long* bar, unsigned parity, int site = 0, int aux = 0) {
    for (int n = 0; n < (1 << 20); ++n) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(t4_smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    printf("[t4 wait timeout] block %d tid %d site %d aux %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, site, aux, parity);
}
#else
__device__ __forceinline__ void t4_mbar_wait(unsigned long long* bar, unsigned parity, int = 0, int = 0) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "T4_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra T4_DONE;\n"
        "bra T4_WAIT;\n"
        "T4_DONE:\n"
        "}\n" ::"r"(t4_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif
// one contiguous block global -> shared through the TMA unit; completion (bytes) is signalled on `bar`
__device__ __forceinline__ void t4_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(t4_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(t4_smem_u32(bar))
                 : "memory");
}
// generic-proxy writes (st.shared into the ring, st.global into the node records) must be ordered before the async proxy (the
// bulk copies) touches the same bytes: all state spaces
__device__ __forceinline__ void t4_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void t4_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// CTA barrier after warp-divergent code.  bar.sync is an ALIGNED barrier: a warp counts as arrived -- and is suspended -- as soon
// as some of its lanes execute it.  With tile widths that are not multiples of 32 a warp holds lanes of several node groups
// (different trip counts) and idle lanes; a lane that ran ahead into the barrier stalled its warp-mates still polling an
// mbarrier (observed: the ring's "full" barrier of their node never completed).  __syncwarp first: lanes wait for each other
// per thread, then the whole warp arrives together.
__device__ __forceinline__ void t4_cta_sync() {
    __syncwarp();
    __syncthreads();
}
template <int R>
__device__ __forceinline__ void t4_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void t4_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

#ifndef T4_SWEEP
#define T4_SWEEP __noinline__
#endif
// setmaxnreg only pays when the two roles never share code after the split (ptxas allocates for the smaller budget on every
// path both can reach, and crashes on it across non-inlined calls): off until the kernel body is split by role
#ifndef MMPC_T4_SETMAXNREG
#define MMPC_T4_SETMAXNREG 0
#endif
#ifndef MMPC_T4_SERIAL_REGS
#define MMPC_T4_SERIAL_REGS 96
#endif
#ifndef MMPC_T4_WORKER_REGS
#define MMPC_T4_WORKER_REGS 136   // 128 x 96 + 384 x 136 = 64512 <= 65536
#endif

template <class M, int I_>
struct T4 {
    static constexpr int NX = M::NX, NU = M::NU, S = NX + NU, N2 = NX + 2 * NU, NH = S * (S + 1) / 2;
    static constexpr int I = I_;          // instances per CTA
    static constexpr int LPI = NX;        // lanes per instance in the serial warps
    static_assert(LPI == 2 || LPI == 4 || LPI == 8, "the serial warps spread an instance over NX lanes");
    static constexpr int IPW = 32 / LPI;  // instances per serial warp
    static constexpr int SWARPS = 4, STHREADS = 32 * SWARPS;
    static_assert(I <= SWARPS * IPW, "one warpgroup of serial warps per CTA");
    static constexpr int NSER = (I + IPW - 1) / IPW;  // serial warps that own instances
    static constexpr int WTHREADS = 384, THREADS = STHREADS + WTHREADS;
    static constexpr int CB = (WTHREADS / I) > 128 ? 128 : (WTHREADS / I);  // node slots of the worker threads
    static constexpr int CF = CB, WORKERS = I * CB, SK = 0;
    // Ring slots (node records in shared memory) = node slots of the workers: group j of worker threads always produces into
    // slot j (its nodes are u = j, j + CB, ...), so the "empty" phase it waits for always follows its OWN previous "full"
    // phase -- with several independent producers per slot a fast group could otherwise see the parity of a phase two uses
    // back as "complete" (mbarrier phase parity is one bit)
    static constexpr int NSLOT = CB;
    __device__ static __forceinline__ bool is_worker(int tid) { return tid >= STHREADS && tid - STHREADS < WORKERS; }
    __device__ static __forceinline__ int widx(int tid) { return tid - STHREADS; }
    __device__ static __forceinline__ int slot_of(int g) { return g; }
    static constexpr int MODE = 0, RK = 0, BD = 0;
    static constexpr bool SPARSE = NX * NX <= 64 && NX * NU <= 64;
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
    // backward ring record (shared memory only): Fx dense | Fu dense | H packed | c | q~ = (qx, qv, qu).  Dense Jacobians:
    // lane a of the recursion reads COLUMN a of Fx (a is a run-time lane index, so no compile-time compaction there)
    static constexpr int RB_FX = 0, RB_FU = NX * NX, RB_H = RB_FU + NX * NU, RB_C = RB_H + NH, RB_Q = RB_C + NX, RB = RB_Q + N2;
    // node record in global memory = ring record of the forward sweep: FG (Fx, Fu compacted | c | q~) | KG (gains) | PG (value function)
    static constexpr int FG_FX = 0, FG_FU = NFX, FG_C = FG_FU + NFU, FG_Q = FG_C + NX, FG = FG_Q + N2;
    static constexpr int KG_KX = 0, KG_KV = NU * NX, KG_K0 = KG_KV + NU * NU, KG = KG_K0 + NU;
    static constexpr int PG_XX = 0, PG_XV = NPX, PG_X = PG_XV + NX * NU, PG = PG_X + NX;
    static constexpr int RF_K = FG, RF_P = FG + KG, RF = FG + KG + PG;
    static constexpr int SLOT = (RB > RF ? RB : RF) * I;  // doubles per ring slot
    static_assert((RF * I) % 2 == 0, "a node record is moved by one bulk copy: a multiple of 16 bytes");
    static constexpr unsigned REC_BYTES = (unsigned)(RF * I * sizeof(double));
    static constexpr int NRED = 5, REDBUF = NRED * CB * I;
    // exchange scratch of the serial warps, per instance: P Fx (NX x NX) | P Fu + Pxv (NX x NU) | y (NX) | Pxv (NX x NU) | Kx columns (NX x NU)
    static constexpr int XS_PA = 0, XS_PB = NX * NX, XS_Y = XS_PB + NX * NU, XS_PV = XS_Y + NX, XS_KX = XS_PV + NX * NU,
                         XS_RAW = XS_KX + NX * NU;
    // stride between instances: an odd number of 16-byte units, so that the instances of a warp hit different bank groups
    static constexpr int XSI = ((XS_RAW + 1) / 2 * 2) + ((((XS_RAW + 1) / 2) % 2 == 0) ? 2 : 0);
    // per-instance constants the evaluators read from shared memory: Q | R | Rm | u_init | regularisation delta
    static constexpr int CS_Q = 0, CS_R = NX, CS_RM = NX + NU, CS_U0 = NX + 2 * NU, CS_DELTA = NX + 3 * NU, NCST = CS_DELTA + 1;
    static constexpr int SM_REC = 0, SM_RED = NSLOT * SLOT, SM_INST = SM_RED + REDBUF, SM_XS = SM_INST + 6 * I, SM_CST = SM_XS + SWARPS * IPW * XSI,
                         SM_BAR = SM_CST + NCST * I, SM_END = SM_BAR + 2 * NSLOT;
    static constexpr size_t SMEM_BYTES = sizeof(double) * SM_END;
    static constexpr int LIN = NX * NX + NX * NU + NX;
};

#define T4_FXNZ(r, c) (!C::SPARSE || ((C::FXM >> ((r) * NX + (c))) & 1ULL))
#define T4_FUNZ(r, c) (!C::SPARSE || ((C::FUM >> ((r) * NU + (c))) & 1ULL))
#define T4_FXSLOT(a) (C::SPARSE ? __popcll(C::FXM & ((1ULL << (a)) - 1ULL)) : (a))
#define T4_FUSLOT(a) (C::SPARSE ? __popcll(C::FUM & ((1ULL << (a)) - 1ULL)) : (a))

// element `a` (run-time) of a small register array without dynamic indexing (which would put the array in local memory)
template <int n>
__device__ __forceinline__ double t4_pick(const double* v, int a) {
    double r = v[0];
    MMPC_UNROLL for (int q = 1; q < n; ++q) r = (a == q) ? v[q] : r;
    return r;
}

// ---- worker: node k of instance i -> ring slot (recursion inputs) + global record (roll-out inputs) + residual pieces ----
// Per-instance constants (weights, u_init, regularisation) come from shared memory and the residual pieces are accumulated in
// shared memory: the evaluator's registers hold the model evaluation and little else (128 per thread at 512 threads per CTA).
template <class M, class C>
__device__ __forceinline__ void t4_eval_node(const T3Ctx<M, C::I, 0>& c, int k, const double* cst_i, double* slot_i, double* recg_i, double* red_i,
                                             unsigned long long* empty_bar, int empty_parity) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, NH = C::NH, I = C::I;
    const int N = c.N;
    const double* z = c.V + (size_t)k * S * I;
    double x[NX], u[NU], r[NX], lam[NX], wq[NX];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        x[a] = z[a * I];
        r[a] = c.traj[(size_t)k * NX + a];
        lam[a] = c.LAM[((size_t)k * NX + a) * I];
        wq[a] = 2.0 * cst_i[(C::CS_Q + a) * I];
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) u[a] = z[(NX + a) * I];
    double F[NX], Fx[NX * NX], Fu[NX * NU], H[NH];
    c.template step<2>(x, u, wq, r, lam, F, Fx, Fu, H);
    // the neighbours' values are needed only from here on (they were prefetched into L1 a round ago): loading them late keeps
    // them out of the registers during the model evaluation
    double xn[NX], v[NU], u_next[NU], lam_prev[NX];
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        xn[a] = z[(S + a) * I];
        lam_prev[a] = k ? c.LAM[((size_t)(k - 1) * NX + a) * I] : 0.0;
    }
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        v[a] = k ? z[(NX + a - S) * I] : cst_i[(C::CS_U0 + a) * I];
        u_next[a] = (k < N - 1) ? z[(S + NX + a) * I] : 0.0;
    }
    if (empty_parity >= 0) t4_mbar_wait(empty_bar, (unsigned)empty_parity, 1, k);  // the recursion has released the slot's previous node
#define T4_REC(f) slot_i[(f) * I]
#define T4_FG(f, val) recg_i[(f) * I] = (val)
    double qe[NX], dual = 0.0, prim = 0.0, theta = 0.0, fval = 0.0, lam1 = 0.0;
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        const double e = F[a] - r[a], cr = F[a] - xn[a];
        qe[a] = wq[a] * e;
        fval = fma(0.5 * wq[a] * e, e, fval);
        prim = fmax(prim, fabs(cr));
        theta += fabs(cr);
        lam1 += fabs(lam[a]);
        T4_REC(C::RB_C + a) = cr;
        T4_FG(C::FG_C + a, cr);
        if (k == N - 1) dual = fmax(dual, fabs(lam[a]));  // stationarity w.r.t. x_N is -lam_{N-1}
    }
    MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) {
        T4_REC(C::RB_FX + a) = T4_FXNZ(a / NX, a % NX) ? Fx[a] : 0.0;
        if (T4_FXNZ(a / NX, a % NX)) T4_FG(C::FG_FX + T4_FXSLOT(a), Fx[a]);
    }
    MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) {
        T4_REC(C::RB_FU + a) = T4_FUNZ(a / NU, a % NU) ? Fu[a] : 0.0;
        if (T4_FUNZ(a / NU, a % NU)) T4_FG(C::FG_FU + T4_FUSLOT(a), Fu[a]);
    }
    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
        double g = 0.0, l = 0.0;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T4_FXNZ(q, a)) {
            g = fma(Fx[q * NX + a], qe[q], g);
            l = fma(Fx[q * NX + a], lam[q], l);
        }
        T4_REC(C::RB_Q + a) = g;
        T4_FG(C::FG_Q + a, g);
        if (k) dual = fmax(dual, fabs(g + l - lam_prev[a]));
    }
    double RRm2[NU];  // 2 (R + Rm): diagonal of the control block
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        double g = 0.0, l = 0.0;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T4_FUNZ(q, a)) {
            g = fma(Fu[q * NU + a], qe[q], g);
            l = fma(Fu[q * NU + a], lam[q], l);
        }
        const double Ra = cst_i[(C::CS_R + a) * I], Rma = cst_i[(C::CS_RM + a) * I];
        RRm2[a] = 2.0 * (Ra + Rma);
        const double du = u[a] - v[a], dRu = 2.0 * Ra * du;
        const double qu = g + dRu + 2.0 * Rma * u[a];
        const double dRn = (k < N - 1) ? 2.0 * Ra * (u_next[a] - u[a]) : 0.0;
        fval = fma(Ra * du, du, fval);
        fval = fma(Rma * u[a], u[a], fval);
        dual = fmax(dual, fabs(qu + l - dRn));
        T4_REC(C::RB_Q + NX + a) = -dRu;  T4_FG(C::FG_Q + NX + a, -dRu);
        T4_REC(C::RB_Q + S + a) = qu;     T4_FG(C::FG_Q + S + a, qu);
    }
    // residual pieces of the iterate: running max / sums of this worker thread over its nodes, kept in shared memory
    red_i[0 * C::CB * I] = fmax(red_i[0 * C::CB * I], dual);
    red_i[1 * C::CB * I] = fmax(red_i[1 * C::CB * I], prim);
    red_i[2 * C::CB * I] += theta;
    red_i[3 * C::CB * I] += fval;
    red_i[4 * C::CB * I] += lam1;
    const double delta = cst_i[C::CS_DELTA * I];
    MMPC_UNROLL for (int p = 0; p < S; ++p) MMPC_UNROLL for (int d = p; d < S; ++d) {
        double t = H[tri(p, d, S)];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) {
            const bool nzp = p < NX ? T4_FXNZ(q, p) : T4_FUNZ(q, p - NX), nzd = d < NX ? T4_FXNZ(q, d) : T4_FUNZ(q, d - NX);
            if (nzp && nzd) {
                const double fa = p < NX ? Fx[q * NX + p] : Fu[q * NU + p - NX];
                const double fd = d < NX ? Fx[q * NX + d] : Fu[q * NU + d - NX];
                t = fma(wq[q] * fa, fd, t);
            }
        }
        if (p == d) t += delta + (p >= NX ? RRm2[p - NX] : 0.0);
        T4_REC(C::RB_H + tri(p, d, S)) = t;
    }
#undef T4_REC
#undef T4_FG
}

// ---- serial warps: the value function of one instance, row a in lane a ------------------------------------------------
template <class M>
struct T4Riccati {
    static constexpr int NX = M::NX, NU = M::NU;
    double Pxx[NX], Pxv[NU], px;     // row a of Pxx and Pxv, entry a of px
    double Pvv[NU * NU], pv[NU];     // replicated in the lanes of the instance
};

// per-lane constants of the row-wise recursion (run-time row index a)
template <class C>
struct T4Lane {
    int a;             // row owned by this lane
    int hoff[C::S];    // (RB_H + packed index of H[a][b]) * I
    int pxoff[C::NX];  // (PG_XX + packed index of Pxx[a][q]) * I
};

// One stage.  Written as "load everything the phase needs, then compute": with the loads of a phase issued back to back their
// ~30-cycle shared-memory latencies overlap (the first version interleaved every LDS with the DFMA that needed it: 2.2 k cycles
// per stage on the GPU).
template <class M, class C>
__device__ __forceinline__ bool t4_riccati_stage(T4Riccati<M>& V, const T4Lane<C>& ln, const double* rp, double* xs, const double* R, double* recg, bool wr) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    const int a = ln.a;
    const double* rpa = rp + a * I;
    // ---- loads of the whole stage from the ring record: uniform fields (every lane of the instance reads the same ones) and
    // the lane's own row / column ----
    double Fx[NX * NX], Fu[NX * NU], cres[NX], Fxc[NX], Hx[NX], Hu[NU], Huu[NU * (NU + 1) / 2], qu[NU], qvv[NU];
    MMPC_UNROLL for (int q = 0; q < NX * NX; ++q) Fx[q] = T4_FXNZ(q / NX, q % NX) ? rp[(C::RB_FX + q) * I] : 0.0;
    MMPC_UNROLL for (int q = 0; q < NX * NU; ++q) Fu[q] = T4_FUNZ(q / NU, q % NU) ? rp[(C::RB_FU + q) * I] : 0.0;
    MMPC_UNROLL for (int q = 0; q < NX; ++q) cres[q] = rp[(C::RB_C + q) * I];
    MMPC_UNROLL for (int q = 0; q < NX; ++q) Fxc[q] = rpa[(C::RB_FX + q * NX) * I];  // column a of Fx
    MMPC_UNROLL for (int b = 0; b < NX; ++b) Hx[b] = rp[ln.hoff[b]];
    MMPC_UNROLL for (int b = 0; b < NU; ++b) Hu[b] = rp[ln.hoff[NX + b]];
    MMPC_UNROLL for (int p = 0; p < NU; ++p) MMPC_UNROLL for (int d = p; d < NU; ++d) Huu[tri(p, d, NU)] = rp[(C::RB_H + tri(NX + p, NX + d, S)) * I];
    const double qxa = rpa[C::RB_Q * I];
    MMPC_UNROLL for (int p = 0; p < NU; ++p) { qvv[p] = rp[(C::RB_Q + NX + p) * I]; qu[p] = rp[(C::RB_Q + S + p) * I]; }
    // ---- value function of stage k+1 -> global record (the roll-out needs it for lam_k); lane a writes its row ----
    if (wr) {
        double* pg = recg + (size_t)C::RF_P * I;
        MMPC_UNROLL for (int b = 0; b < NX; ++b) if (b >= a) pg[ln.pxoff[b]] = V.Pxx[b];
        MMPC_UNROLL for (int b = 0; b < NU; ++b) pg[(C::PG_XV + a * NU + b) * I] = V.Pxv[b];
        pg[(C::PG_X + a) * I] = V.px;
    }
    // ---- step 1: row a of P Fx, P Fu + Pxv and y = p + P c ----
    {
        double y = V.px;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) y = fma(V.Pxx[q], cres[q], y);
        xs[C::XS_Y + a] = y;
        MMPC_UNROLL for (int b = 0; b < NX; ++b) {
            double t = 0.0;
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T4_FXNZ(q, b)) t = fma(V.Pxx[q], Fx[q * NX + b], t);
            xs[C::XS_PA + a * NX + b] = t;
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) {
            double t = V.Pxv[b];
            MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T4_FUNZ(q, b)) t = fma(V.Pxx[q], Fu[q * NU + b], t);
            xs[C::XS_PB + a * NU + b] = t;
            xs[C::XS_PV + a * NU + b] = V.Pxv[b];
        }
    }
    __syncwarp();
    // ---- step 2: row a of the stage matrices G = H + Z' P Z, the (replicated) control block and the gains ----
    double PA[NX * NX], PB[NX * NU], PV[NX * NU], Y[NX];
    MMPC_UNROLL for (int q = 0; q < NX * NX; ++q) PA[q] = xs[C::XS_PA + q];
    MMPC_UNROLL for (int q = 0; q < NX * NU; ++q) { PB[q] = xs[C::XS_PB + q]; PV[q] = xs[C::XS_PV + q]; }
    MMPC_UNROLL for (int q = 0; q < NX; ++q) Y[q] = xs[C::XS_Y + q];
    double Gxx[NX], Gxu[NU], gx = qxa;
    MMPC_UNROLL for (int q = 0; q < NX; ++q) gx = fma(Fxc[q], Y[q], gx);
    MMPC_UNROLL for (int b = 0; b < NX; ++b) {
        double t = Hx[b];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(Fxc[q], PA[q * NX + b], t);
        Gxx[b] = t;
    }
    MMPC_UNROLL for (int b = 0; b < NU; ++b) {
        double t = Hu[b];
        MMPC_UNROLL for (int q = 0; q < NX; ++q) t = fma(Fxc[q], PB[q * NU + b], t);
        Gxu[b] = t;
    }
    double Guu[NU * NU], gu[NU];
    MMPC_UNROLL for (int p = 0; p < NU; ++p) {
        MMPC_UNROLL for (int d = p; d < NU; ++d) {
            double t = Huu[tri(p, d, NU)] + 0.5 * (V.Pvv[p * NU + d] + V.Pvv[d * NU + p]), t2 = 0.0;  // two chains
            MMPC_UNROLL for (int q = 0; q < NX; ++q) {
                if (T4_FUNZ(q, p)) t = fma(Fu[q * NU + p], PB[q * NU + d], t);
                if (T4_FUNZ(q, d)) t2 = fma(PV[q * NU + p], Fu[q * NU + d], t2);
            }
            Guu[p * NU + d] = t + t2;
            Guu[d * NU + p] = t + t2;
        }
        double t = qu[p] + V.pv[p], t2 = 0.0;
        MMPC_UNROLL for (int q = 0; q < NX; ++q) {
            t = fma(PV[q * NU + p], cres[q], t);
            if (T4_FUNZ(q, p)) t2 = fma(Fu[q * NU + p], Y[q], t2);
        }
        gu[p] = t + t2;
    }
    bool ok;
    double Kxc[NU], Kv[NU * NU], k0[NU];  // column a of Kx; Kv and k0 replicated
    if constexpr (NU == 2) {
        const double det = fma(Guu[0], Guu[3], -Guu[1] * Guu[1]);
        ok = (Guu[0] > 0.0) && (det > 0.0) && isfinite(det);
        const double rd = 1.0 / (ok ? det : 1.0);
        const double i00 = Guu[3] * rd, i01 = -Guu[1] * rd, i11 = Guu[0] * rd;
        Kxc[0] = -fma(i00, Gxu[0], i01 * Gxu[1]);
        Kxc[1] = -fma(i01, Gxu[0], i11 * Gxu[1]);
        Kv[0] = i00 * (2.0 * R[0]); Kv[2] = i01 * (2.0 * R[0]);
        Kv[1] = i01 * (2.0 * R[1]); Kv[3] = i11 * (2.0 * R[1]);
        k0[0] = -fma(i00, gu[0], i01 * gu[1]);
        k0[1] = -fma(i01, gu[0], i11 * gu[1]);
    } else {
        ok = chol<NU>(Guu);
        MMPC_UNROLL for (int b = 0; b < NU; ++b) Kxc[b] = -Gxu[b];
        chol_solve<NU>(Guu, Kxc);
        MMPC_UNROLL for (int l = 0; l < NU; ++l) {
            double col[NU];
            MMPC_UNROLL for (int b = 0; b < NU; ++b) col[b] = (b == l) ? 2.0 * R[l] : 0.0;
            chol_solve<NU>(Guu, col);
            MMPC_UNROLL for (int b = 0; b < NU; ++b) Kv[b * NU + l] = col[b];
        }
        MMPC_UNROLL for (int b = 0; b < NU; ++b) k0[b] = -gu[b];
        chol_solve<NU>(Guu, k0);
    }
    MMPC_UNROLL for (int l = 0; l < NU; ++l) xs[C::XS_KX + a * NU + l] = Kxc[l];
    __syncwarp();
    // ---- step 3: row a of the value function of stage k ----
    double KX[NX * NU];
    MMPC_UNROLL for (int q = 0; q < NX * NU; ++q) KX[q] = xs[C::XS_KX + q];
    if (wr) {  // gains -> global record (after the exchange: off the stage's dependent chain)
        double* kg = recg + (size_t)C::RF_K * I;
        MMPC_UNROLL for (int l = 0; l < NU; ++l) kg[(C::KG_KX + l * NX + a) * I] = Kxc[l];
        if (a == 0) {
            MMPC_UNROLL for (int q = 0; q < NU * NU; ++q) kg[(C::KG_KV + q) * I] = Kv[q];
            MMPC_UNROLL for (int q = 0; q < NU; ++q) kg[(C::KG_K0 + q) * I] = k0[q];
        }
    }
    MMPC_UNROLL for (int b = 0; b < NX; ++b) {
        double t = Gxx[b];
        MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Gxu[l], KX[b * NU + l], t);
        V.Pxx[b] = t;
    }
    MMPC_UNROLL for (int b = 0; b < NU; ++b) {
        double t = 0.0;
        MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Gxu[l], Kv[l * NU + b], t);
        V.Pxv[b] = t;
    }
    {
        double t = gx;
        MMPC_UNROLL for (int l = 0; l < NU; ++l) t = fma(Gxu[l], k0[l], t);
        V.px = t;
    }
    MMPC_UNROLL for (int p = 0; p < NU; ++p) {
        MMPC_UNROLL for (int d = 0; d < NU; ++d) V.Pvv[p * NU + d] = (p == d ? 2.0 * R[p] : 0.0) - 2.0 * R[p] * Kv[p * NU + d];
        V.pv[p] = qvv[p] - 2.0 * R[p] * k0[p];
    }
    return ok;
}

// The ring's barriers are initialised ONCE per kernel (full: the I threads of the slot's worker group; empty: every lane of the
// serial warps that own instances) and never re-armed: on the GPU, mbarrier.inval + mbarrier.init of live objects between
// sweeps left some of them in their old phase (slots whose use count in a sweep was odd then looked "already complete").
// Instead every slot goes through exactly R = ceil(N / NSLOT) phases per sweep -- slots without a node in the last round do an
// empty hand-shake -- and each thread carries the running phase number `ring_phase` from sweep to sweep.
template <class C>
__device__ __forceinline__ void t4_ring_init(double* smem, int tid) {
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
    if (tid < 2 * C::NSLOT) t4_mbar_init(&bar[tid], tid < C::NSLOT ? (unsigned)C::I : (unsigned)(32 * C::NSER));
    t4_fence_barrier_init();
    t4_cta_sync();
}
// entry of a sweep: generic-proxy writes of the previous phase (st.shared into the ring, st.global into the node records) are
// ordered before the bulk copies of this one
template <class C>
__device__ __forceinline__ void t4_sweep_entry() {
    t4_fence_proxy_async();
    t4_cta_sync();
}

// ---- backward sweep: node evaluation (workers) -> ring -> Riccati recursion (serial warps) ----------------------------------
// The two roles are separate non-inlined functions: ptxas allocates registers and schedules per function, and the recursion's
// short dependent chains were scheduled for minimum register use while they shared a function with the evaluator.
template <class M, class C>
__device__ __noinline__ void t4_bwd_worker(const T3Ctx<M, C::I, 0>& cin, double* smem, int tid, double* rec_tile, int base, int R) {
    constexpr int I = C::I;
    const T3Ctx<M, C::I, 0> c = cin;
    const int N = c.N;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
    unsigned long long* empty = full + C::NSLOT;
    const int w = C::widx(tid), wi = w % I, wj = w / I;
    double* red_i = smem + C::SM_RED + wj * I + wi;  // row n of this thread's residual pieces: red_i[n * CB * I]
    MMPC_UNROLL for (int n = 0; n < C::NRED; ++n) red_i[n * C::CB * I] = 0.0;
    const double* cst_i = smem + C::SM_CST + wi;
    const long long tw0 = T4_CLK();
    for (int r = 0; r < R; ++r) {
        const int u = wj + r * C::CB, ph = base + r;  // u-th node of the sweep: k = N-1-u
        if (u < N) {
            const int k = N - 1 - u;
            if (u + C::CB < N) t3_prefetch_node<M, C>(c, k - C::CB);
            t4_eval_node<M, C>(c, k, cst_i, smem + C::SM_REC + (size_t)wj * C::SLOT + wi, rec_tile + (size_t)k * C::RF * I + wi, red_i, &empty[wj],
                               ph > 0 ? ((ph - 1) & 1) : -1);
        } else if (ph > 0) {
            t4_mbar_wait(&empty[wj], (unsigned)((ph - 1) & 1), 1, -1);
        }
        t4_mbar_arrive(&full[wj]);
    }
#ifdef MMPC_T4_PROFILE
    if (blockIdx.x == 0 && (w == 0 || w == C::WORKERS - 1)) printf("[t4 bwd worker %d] %d rounds, %lld cycles\n", w, R, T4_CLK() - tw0);
#endif
}

template <class M, class C>
__device__ __noinline__ bool t4_bwd_serial(double* smem, int tid, double* rec_tile, int N, int base, int R) {
    constexpr int NX = M::NX, NU = M::NU, I = C::I, S = C::S;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
    unsigned long long* empty = full + C::NSLOT;
    const int l = tid & 31, isub = l / C::LPI, a = l % C::LPI, q = tid >> 5;
    const int ii = min(q * C::IPW + isub, I - 1);
    const bool wr = q * C::IPW + isub < I;
    T4Lane<C> ln;
    ln.a = a;
    MMPC_UNROLL for (int b = 0; b < S; ++b) ln.hoff[b] = (C::RB_H + trisym(a, b, S)) * I;
    MMPC_UNROLL for (int b = 0; b < NX; ++b) ln.pxoff[b] = (C::PG_XX + trisym(a, b, NX)) * I;
    T4Riccati<M> V;
    MMPC_UNROLL for (int b = 0; b < NX; ++b) V.Pxx[b] = 0.0;
    MMPC_UNROLL for (int b = 0; b < NU; ++b) V.Pxv[b] = 0.0;
    V.px = 0.0;
    MMPC_UNROLL for (int b = 0; b < NU * NU; ++b) V.Pvv[b] = 0.0;
    MMPC_UNROLL for (int b = 0; b < NU; ++b) V.pv[b] = 0.0;
    double R2[NU];
    MMPC_UNROLL for (int b = 0; b < NU; ++b) R2[b] = smem[C::SM_CST + (C::CS_R + b) * I + ii];
    double* xs = smem + C::SM_XS + (size_t)(q * C::IPW + isub) * C::XSI;
    bool ok = true;
    long long t_wait = 0, t_stage = 0;
    for (int u = 0; u < R * C::NSLOT; ++u) {
        const int k = N - 1 - u, sl = u % C::NSLOT;
        const long long t0 = T4_CLK();
        t4_mbar_wait(&full[sl], (unsigned)((base + u / C::NSLOT) & 1), 2, u);
        const long long t1 = T4_CLK();
        if (u < N) ok = t4_riccati_stage<M, C>(V, ln, smem + C::SM_REC + (size_t)sl * C::SLOT + ii, xs, R2, rec_tile + (size_t)k * C::RF * I + ii, wr) && ok;
        t4_mbar_arrive_relaxed(&empty[sl]);
        t_wait += t1 - t0; t_stage += T4_CLK() - t1;
    }
#ifdef MMPC_T4_PROFILE
    if (blockIdx.x == 0 && tid == 0) printf("[t4 bwd serial] waiting %lld, stages %lld cycles (%d nodes)\n", t_wait, t_stage, N);
#endif
    return ok;
}

// returns (serial lanes) whether every G_uu was positive definite; residual norms of the iterate on the master lanes
template <class M, class C>
__device__ __forceinline__ bool t4_backward(const T3Ctx<M, C::I, 0>& c, double* smem, int tid, double delta_self, bool master, int inst, T3Norms& nrm_out,
                                            double* rec_tile, int& ring_phase) {
    constexpr int I = C::I;
    const int N = c.N;
    double* red = smem + C::SM_RED;
    if (master) smem[C::SM_CST + C::CS_DELTA * I + inst] = delta_self;
    t4_sweep_entry<C>();
    const int R = (N + C::NSLOT - 1) / C::NSLOT, base = ring_phase;  // phases of every slot in this sweep
    ring_phase += R;
    bool ok = true;
    if (tid >= C::STHREADS) {
        if (C::widx(tid) < C::WORKERS) t4_bwd_worker<M, C>(c, smem, tid, rec_tile, base, R);
    } else if ((tid >> 5) < C::NSER) {
        ok = t4_bwd_serial<M, C>(smem, tid, rec_tile, N, base, R);
    }
    t4_cta_sync();
    T3Norms o;
    o.dual_inf = o.prim_inf = o.theta = o.fval = o.lam_l1 = 0.0;
    if (master) {
        for (int s = 0; s < C::CB; ++s) {
            o.dual_inf = fmax(o.dual_inf, red[(0 * C::CB + s) * I + inst]);
            o.prim_inf = fmax(o.prim_inf, red[(1 * C::CB + s) * I + inst]);
            o.theta += red[(2 * C::CB + s) * I + inst];
            o.fval += red[(3 * C::CB + s) * I + inst];
            o.lam_l1 += red[(4 * C::CB + s) * I + inst];
        }
    }
    __syncthreads();
    nrm_out = o;
    return ok;
}

// ---- forward sweep: bulk-copied node records -> ring -> roll-out of the step (serial warps); merit of the trial point (workers) ----
template <class M, class C>
__device__ __noinline__ void t4_fwd_worker(const T3Ctx<M, C::I, 0>& cin, double* smem, int tid, const double* rec_tile, int base, int R) {
    constexpr int NX = M::NX, S = C::S, I = C::I;
    const T3Ctx<M, C::I, 0> c = cin;
    const int N = c.N;
    double* red = smem + C::SM_RED;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
    unsigned long long* empty = full + C::NSLOT;
    const int w = C::widx(tid), wi = w % I, wj = w / I;
    const bool act = smem[C::SM_INST + 2 * I + wi] != 0.0;
    double f1 = 0.0, th1 = 0.0;
    T3TrialBD bd;
    bd.bar1 = 0.0; bd.amax = bd.az = 1.0; bd.full = true; bd.mu = 0.0; bd.tau = 0.99;
    // Group wj brings the records of ITS nodes (k = wj, wj + CB, ...) into ITS slot: the record of the next round is
    // requested right after the wait for "the roll-out has passed node k", which is also the slot's "empty" condition.
    // Every thread of the group arrives on the slot's full barrier (count I, as in the backward sweep); the first one
    // carries the transaction bytes of the bulk copy.
    auto produce = [&](int r) {
        const int kk = wj + r * C::CB;
#ifdef MMPC_T4_NO_TMA
        if (kk < N) {  // developer build: the node record is copied with plain loads / stores (bisecting aid)
            const double* src = rec_tile + (size_t)kk * C::RF * I + wi;
            double* dst = smem + C::SM_REC + (size_t)wj * C::SLOT + wi;
            for (int f = 0; f < C::RF; ++f) dst[f * I] = src[f * I];
        }
        t4_mbar_arrive(&full[wj]);
#else
        if (wi == 0 && kk < N) {
            t4_mbar_arrive_expect_tx(&full[wj], C::REC_BYTES);
            t4_bulk_load(smem + C::SM_REC + (size_t)wj * C::SLOT, rec_tile + (size_t)kk * C::RF * I, C::REC_BYTES, &full[wj]);
        } else {
            t4_mbar_arrive(&full[wj]);
        }
#endif
    };
    const long long tw0 = T4_CLK();
    long long t_merit = 0;
    if (base > 0) t4_mbar_wait(&empty[wj], (unsigned)((base - 1) & 1), 3, -1);  // the slot's last use of the previous sweep (long complete)
    produce(0);
    for (int r = 0; r < R; ++r) {
        const int k = wj + r * C::CB;
        if (k < N) {
            MMPC_UNROLL for (int a = 0; a < S + NX; ++a) t3_prefetch(c.V + ((size_t)k * S + a) * I);  // iterate rows of the merit evaluation -> L1
            t3_prefetch(c.traj + (size_t)k * NX);
        }
        t4_mbar_wait(&empty[wj], (unsigned)((base + r) & 1), 3, k);  // the roll-out has passed node k: dz_k, dz_{k+1} are written
        if (r + 1 < R) produce(r + 1);
        const long long tm0 = T4_CLK();
        if (k < N) t3_trial_node<M, C>(c, k, 1.0, f1, th1, bd, !act);
        t_merit += T4_CLK() - tm0;
    }
#ifdef MMPC_T4_PROFILE
    if (blockIdx.x == 0 && (w == 0 || w == C::WORKERS - 1)) printf("[t4 fwd worker %d] total %lld, merit %lld cycles\n", w, T4_CLK() - tw0, t_merit);
#endif
    red[(0 * C::CB + wj) * I + wi] = f1;
    red[(1 * C::CB + wj) * I + wi] = th1;
}

struct T4FwdSerial {
    double gradf_d, lam_inf;
    bool finite;
};

template <class M, class C>
__device__ __noinline__ void t4_fwd_serial(double* smem, int tid, double* DV, double* LAMN, int N, int base, int R, T4FwdSerial& out) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
    unsigned long long* empty = full + C::NSLOT;
    const int l = tid & 31, isub = l / C::LPI, a = l % C::LPI, q = tid >> 5;
    const int ii = min(q * C::IPW + isub, I - 1);
    const bool wr = q * C::IPW + isub < I;
    const bool act = smem[C::SM_INST + 2 * I + ii] != 0.0;
    int pxoff[NX];
    MMPC_UNROLL for (int b = 0; b < NX; ++b) pxoff[b] = (C::RF_P + C::PG_XX + trisym(a, b, NX)) * I;
    double dx[NX], dv[NU];
    MMPC_UNROLL for (int b = 0; b < NX; ++b) dx[b] = 0.0;
    MMPC_UNROLL for (int b = 0; b < NU; ++b) dv[b] = 0.0;
    double gradf_d = 0.0, lam_inf = 0.0;
    bool finite = true;
    long long t_wait = 0;
    const long long ts0 = T4_CLK();
    for (int k = 0; k < R * C::NSLOT; ++k) {
        const int sl = k % C::NSLOT;
        const long long t0 = T4_CLK();
        t4_mbar_wait(&full[sl], (unsigned)((base + k / C::NSLOT) & 1), 4, k);
        t_wait += T4_CLK() - t0;
        if (k >= N) {  // slot without a node in the last round: empty hand-shake
            t4_mbar_arrive(&empty[sl]);
            continue;
        }
        const double* rp = smem + C::SM_REC + (size_t)sl * C::SLOT + ii;
        // all loads of the stage first (their latencies overlap), then the two short dependent chains du -> dx+
        double K0[NU], KX[NU * NX], KV[NU * NU], Cc[NX], FXv[C::NFX], FUv[C::NFU], Qt[C::N2], Pr[NX], Pv[NU];
        MMPC_UNROLL for (int p = 0; p < NU; ++p) K0[p] = rp[(C::RF_K + C::KG_K0 + p) * I];
        MMPC_UNROLL for (int p = 0; p < NU * NX; ++p) KX[p] = rp[(C::RF_K + C::KG_KX + p) * I];
        MMPC_UNROLL for (int p = 0; p < NU * NU; ++p) KV[p] = rp[(C::RF_K + C::KG_KV + p) * I];
        MMPC_UNROLL for (int p = 0; p < NX; ++p) Cc[p] = rp[(C::FG_C + p) * I];
        MMPC_UNROLL for (int p = 0; p < C::NFX; ++p) FXv[p] = rp[(C::FG_FX + p) * I];
        MMPC_UNROLL for (int p = 0; p < C::NFU; ++p) FUv[p] = rp[(C::FG_FU + p) * I];
        MMPC_UNROLL for (int p = 0; p < C::N2; ++p) Qt[p] = rp[(C::FG_Q + p) * I];
        const double pxa = rp[(C::RF_P + C::PG_X + a) * I];
        MMPC_UNROLL for (int p = 0; p < NX; ++p) Pr[p] = rp[pxoff[p]];
        MMPC_UNROLL for (int p = 0; p < NU; ++p) Pv[p] = rp[(C::RF_P + C::PG_XV + a * NU + p) * I];
        double du[NU], dxn[NX];
        MMPC_UNROLL for (int p = 0; p < NU; ++p) {
            double t = K0[p], t2 = 0.0;
            MMPC_UNROLL for (int r = 0; r < NX; ++r) t = fma(KX[p * NX + r], dx[r], t);
            MMPC_UNROLL for (int r = 0; r < NU; ++r) t2 = fma(KV[p * NU + r], dv[r], t2);
            du[p] = t + t2;
        }
        MMPC_UNROLL for (int p = 0; p < NX; ++p) {
            double t = Cc[p], t2 = 0.0;
            MMPC_UNROLL for (int r = 0; r < NX; ++r) if (T4_FXNZ(p, r)) t = fma(FXv[T4_FXSLOT(p * NX + r)], dx[r], t);
            MMPC_UNROLL for (int r = 0; r < NU; ++r) if (T4_FUNZ(p, r)) t2 = fma(FUv[T4_FUSLOT(p * NU + r)], du[r], t2);
            dxn[p] = t + t2;
        }
        // directional derivative of the objective: q~ . (dx, dv, du)  (every lane of the instance computes it; the master's copy is used)
        MMPC_UNROLL for (int p = 0; p < NX; ++p) gradf_d = fma(Qt[p], dx[p], gradf_d);
        MMPC_UNROLL for (int p = 0; p < NU; ++p) {
            gradf_d = fma(Qt[NX + p], dv[p], gradf_d);
            gradf_d = fma(Qt[S + p], du[p], gradf_d);
        }
        // lam_k+ = (P_{k+1} (dx_{k+1}, du_k) + p_{k+1})_x : entry a in lane a
        double lam = pxa;
        MMPC_UNROLL for (int r = 0; r < NX; ++r) lam = fma(Pr[r], dxn[r], lam);
        MMPC_UNROLL for (int r = 0; r < NU; ++r) lam = fma(Pv[r], du[r], lam);
        const double dxa = t4_pick<NX>(dxn, a);
        if (wr) {
            if (act) LAMN[((size_t)k * NX + a) * I] = lam;
            DV[((size_t)(k + 1) * S + a) * I] = dxa;
            if (a < NU) DV[((size_t)k * S + NX + a) * I] = t4_pick<NU>(du, a);
        }
        lam_inf = fmax(lam_inf, fabs(lam));
        finite = finite && isfinite(lam) && isfinite(dxa);
        MMPC_UNROLL for (int b = 0; b < NX; ++b) dx[b] = dxn[b];
        MMPC_UNROLL for (int b = 0; b < NU; ++b) dv[b] = du[b];
        t4_mbar_arrive(&empty[sl]);
    }
#ifdef MMPC_T4_PROFILE
    if (blockIdx.x == 0 && tid == 0) printf("[t4 fwd serial] total %lld, waiting %lld cycles\n", T4_CLK() - ts0, t_wait);
#endif
    // combine the rows of the instance (lanes a = 0..LPI-1 are adjacent)
    MMPC_UNROLL for (int m = 1; m < C::LPI; m <<= 1) {
        lam_inf = fmax(lam_inf, __shfl_xor_sync(0xffffffffu, lam_inf, m));
        finite = __shfl_xor_sync(0xffffffffu, (int)finite, m) && finite;
    }
    out.gradf_d = gradf_d; out.lam_inf = lam_inf; out.finite = finite;
}

template <class M, class C>
__device__ __forceinline__ void t4_forward(const T3Ctx<M, C::I, 0>& c, double* smem, int tid, bool master, int inst, T3Fwd& o_out, const double* rec_tile,
                                           int& ring_phase) {
    constexpr int I = C::I;
    const int N = c.N;
    double* red = smem + C::SM_RED;
    T3Fwd o;
    o.gradf_d = o.lam_inf = o.f1 = o.th1 = o.bar1 = 0.0;
    o.amax = o.az = 1.0;
    o.finite = true;
    t4_sweep_entry<C>();
    const int R = (N + C::NSLOT - 1) / C::NSLOT, base = ring_phase;
    ring_phase += R;
    if (tid >= C::STHREADS) {
        if (C::widx(tid) < C::WORKERS) t4_fwd_worker<M, C>(c, smem, tid, rec_tile, base, R);
    } else if ((tid >> 5) < C::NSER) {
        T4FwdSerial fs;
        t4_fwd_serial<M, C>(smem, tid, c.DV, c.LAMN, N, base, R, fs);
        o.gradf_d = fs.gradf_d; o.lam_inf = fs.lam_inf; o.finite = fs.finite;
    }
    t4_cta_sync();
    if (master) {
        for (int s = 0; s < C::CB; ++s) {
            o.f1 += red[(0 * C::CB + s) * I + inst];
            o.th1 += red[(1 * C::CB + s) * I + inst];
        }
    }
    __syncthreads();
    o_out = o;
}

template <class M, int I_>
__global__ void __launch_bounds__(T4<M, I_>::THREADS, 1) solve_kernel_t4(mmpc_problem P, mmpc_workspace W, mmpc_solve_io io) {
    using C = T4<M, I_>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, tile = io.b0 / I + blockIdx.x;  // b0 is a multiple of the tile width (the ABI layer cuts sub-batches on tile boundaries)
    const bool serial = tid < C::STHREADS;
#ifdef MMPC_T4_PROFILE
    const long long t_entry = T4_CLK();
#endif
    t4_ring_init<C>(smem, tid);
    int ring_phase = 0;  // phases every ring slot has completed so far (same value in every thread)
    // register budgets by role: warpgroup 0 = serial warps, warpgroups 1-3 = evaluators
#if MMPC_T4_SETMAXNREG
    if (serial) t4_reg_dec<MMPC_T4_SERIAL_REGS>();
    else t4_reg_inc<MMPC_T4_WORKER_REGS>();
#endif
    const bool worker = C::is_worker(tid);
    // a thread's instance: worker w -> instance w % I (node slot w / I); serial lane -> instance of its lane group
    const int sinst = (tid >> 5) * C::IPW + (tid & 31) / C::LPI;
    const bool master = serial && (tid & 31) % C::LPI == 0 && sinst < I;  // holds the per-instance solver state
    const int i = worker ? C::widx(tid) % I : (serial ? min(sinst, I - 1) : 0);
    const int j = worker ? C::widx(tid) / I : 0;
    const int N = P.N, NV = NX * (N + 1) + NU * N, Mc = N * NX;
    const int b = tile * I + i;
    const bool valid = b >= io.b0 && b < io.b1;
    const int bi = valid ? b : io.b1 - 1;  // surplus slots read the last instance's inputs and never write outputs

    T3Ctx<M, I, 0> c;
    c.N = N;
    c.h = P.h; c.is_linear = P.is_linear; c.substeps = P.substeps;
    c.V = W.V2 + (size_t)tile * I * NV + i; c.LAM = W.LAM2 + (size_t)tile * I * Mc + i;
    c.DV = W.DV2 + (size_t)tile * I * NV + i; c.LAMN = W.LAMN2 + (size_t)tile * I * Mc + i;
    c.LINp = W.LIN2 + (size_t)b * C::LIN;
    c.traj = io.traj + (size_t)bi * N * NX;
    double* rec_tile = W.FG2 + (size_t)tile * I * N * C::RF;  // [node][RF fields][I]; FG2, KG2, PG2 are one contiguous region (mmpc_api.cu)
    MMPC_UNROLL for (int a = 0; a < NX; ++a) c.Q[a] = W.Qw[(size_t)a * W.Bp + bi];
    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
        c.R[a] = W.Rw[(size_t)a * W.Bp + bi];
        c.Rm[a] = W.Rmw[(size_t)a * W.Bp + bi];
        c.uinit[a] = io.control[(size_t)bi * NU + a];
    }
    if (worker && j == 0) {  // per-instance constants of the evaluators
        double* cs = smem + C::SM_CST + i;
        MMPC_UNROLL for (int a = 0; a < NX; ++a) cs[(C::CS_Q + a) * I] = c.Q[a];
        MMPC_UNROLL for (int a = 0; a < NU; ++a) { cs[(C::CS_R + a) * I] = c.R[a]; cs[(C::CS_RM + a) * I] = c.Rm[a]; cs[(C::CS_U0 + a) * I] = c.uinit[a]; }
        cs[C::CS_DELTA * I] = 0.0;
    }
    // ---- start point (ModelControl.cpp:28-50,68 cold; :161 warm), x_0 pinned (:144-145) ----
    const bool warm = io.warm_start && W.have_warm;
    if (worker) {
        if (!warm) for (int e = NX + j; e < NV; e += C::CB) c.V[(size_t)e * I] = 0.0;
        if (!(warm && io.warm_start_duals)) for (int e = j; e < Mc; e += C::CB) c.LAM[(size_t)e * I] = 0.0;
        for (int e = j; e < NX; e += C::CB) {
            c.V[(size_t)e * I] = io.state[(size_t)bi * NX + e];
            c.DV[(size_t)e * I] = 0.0;
        }
    }
    t4_cta_sync();
    if (P.is_linear && worker && j == 0) {
        double x0[NX], xd[NX], A[NX * NX], Bm[NX * NU];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) x0[a] = c.V[(size_t)a * I];
        M::fAB(x0, c.uinit, xd, A, Bm);
        MMPC_UNROLL for (int a = 0; a < NX * NX; ++a) c.LINp[a] = A[a];
        MMPC_UNROLL for (int a = 0; a < NX * NU; ++a) c.LINp[NX * NX + a] = Bm[a];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) c.LINp[NX * NX + NX * NU + a] = xd[a];
    }
    t4_cta_sync();

    double* const Vp = c.V; const double* const DVp = c.DV;
    double nu_merit = 1.0, delta_last = 0.0, err = INFINITY, fval = 0.0;
    int it = 0, it_done = 0, status = MMPC_SOLVE_MAX_ITER, ls_total = 0, reg_total = 0;
    bool active = valid && master;
    double* sh_alpha = smem + C::SM_INST + I;
    double* sh_flag = smem + C::SM_INST + 2 * I;
    const int nfree = N * S, ncon = N * NX;
    bool expect_converged = false;
#ifdef MMPC_T4_PROFILE
    const long long tk0 = T4_CLK();
    long long tp_res = 0, tp_bwd = 0, tp_fwd = 0, tp_post = 0;
#endif
    for (it = 0;; ++it) {
        T3Norms nrm;
        double delta = 0.0;
        bool ok = true;
        if (expect_converged) {
            // every instance of the tile that still iterates took a full Newton step from a small KKT error: test the new iterate
            // with the cheap first-derivative residual sweep; the full sweep runs only if some instance fails (as in T3)
            T3NormsBD nb;
#ifdef MMPC_T4_PROFILE
            const long long tr0 = T4_CLK();
#endif
            t3_residuals<M, C>(c, smem, tid, nb);
#ifdef MMPC_T4_PROFILE
            tp_res += T4_CLK() - tr0;
#endif
            // t3_residuals reduces into the threads tid < I; here the per-instance state lives in the master lanes
            double* rr = smem + C::SM_RED;
            if (tid < I) { rr[0 * I + tid] = nb.dual_inf; rr[1 * I + tid] = nb.prim_inf; rr[2 * I + tid] = nb.fval; rr[3 * I + tid] = nb.lam_l1; }
            __syncthreads();
            if (active) {
                const double sd = fmax(100.0, rr[3 * I + sinst] / (double)(nfree + ncon)) / 100.0;
                const double e1 = fmax(rr[0 * I + sinst] / sd, rr[1 * I + sinst]);
                if (isfinite(e1) && e1 <= io.tol) {
                    err = e1; fval = rr[2 * I + sinst]; it_done = it;
                    status = MMPC_SOLVE_CONVERGED;
                    active = false;
                }
            }
            if (!__syncthreads_or(active)) break;
        }
#ifdef MMPC_T4_PROFILE
        const long long tb0 = T4_CLK();
#endif
        ok = t4_backward<M, C>(c, smem, tid, delta, master, sinst, nrm, rec_tile, ring_phase);
#ifdef MMPC_T4_PROFILE
        tp_bwd += T4_CLK() - tb0;
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
        if (!__syncthreads_or(active)) break;
        int tries = 0;
        while (__syncthreads_or(active && !ok)) {  // inertia correction for the instances that need it
            if (active && !ok) {
                delta = (delta == 0.0) ? (delta_last > 0.0 ? fmax(1e-20, delta_last / 3.0) : 1e-4) : delta * 8.0;
                ++reg_total;
            }
            T3Norms dummy;
            const bool ok2 = t4_backward<M, C>(c, smem, tid, delta, master, sinst, dummy, rec_tile, ring_phase);
            if (active && !ok) ok = ok2;
            if (++tries >= 24) {
                if (active && !ok) { status = MMPC_SOLVE_NUMERICAL; active = false; }
                break;
            }
        }
        if (delta > 0.0) delta_last = delta;
        if (master) sh_flag[sinst] = active ? 1.0 : 0.0;  // which instances still iterate: read by the forward sweep
        if (!__syncthreads_or(active)) break;
        T3Fwd fo;
#ifdef MMPC_T4_PROFILE
        const long long tf0 = T4_CLK();
#endif
        t4_forward<M, C>(c, smem, tid, master, sinst, fo, rec_tile, ring_phase);
#ifdef MMPC_T4_PROFILE
        const long long tf1 = T4_CLK();
        tp_fwd += tf1 - tf0;
#endif
        if (active && !(fo.finite && isfinite(fo.f1))) { status = MMPC_SOLVE_NUMERICAL; active = false; }
        nu_merit = fmax(nu_merit, 1.1 * fo.lam_inf + 1e-8);
        const double m0 = nrm.fval + nu_merit * nrm.theta;
        const double D = fo.gradf_d - nu_merit * nrm.theta;
        double alpha = 1.0;
        double f1 = fo.f1, th1 = fo.th1;
        int ls = 0;
        bool need = active && !(f1 + nu_merit * th1 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0));
        while (__syncthreads_or(need)) {
            if (need) { alpha *= 0.5; ++ls_total; }
            if (master) sh_alpha[sinst] = alpha;
            __syncthreads();
            double f2, t2, b2;
            t3_trial<M, C>(c, smem, tid, f2, t2, b2);
            double* rr = smem + C::SM_RED;  // t3_trial reduces into the threads tid < I: hand the sums to the master lanes
            if (tid < I) { rr[0 * I + tid] = f2; rr[1 * I + tid] = t2; }
            __syncthreads();
            if (need) need = !(rr[0 * I + sinst] + nu_merit * rr[1 * I + sinst] <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0)) && (++ls < 40);
            __syncthreads();
        }
        // accept the step.  Multipliers: LAMN already holds lam+ (instances that iterate) or the old values (the others), so a
        // full step only swaps the two arrays; a shortened step (some instance of the tile) blends in place first.
        if (master) sh_alpha[sinst] = active ? alpha : 0.0;
        const bool shortened = __syncthreads_or(active && alpha != 1.0);
        if (worker) {
            const double al = sh_alpha[i];
            constexpr int UB = 8;
            if (al != 0.0) for (int e0 = NX + j; e0 < NV; e0 += C::CB * UB) {
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
        t4_cta_sync();
        { double* t = c.LAM; c.LAM = c.LAMN; c.LAMN = t; }
        {
            const bool expect = alpha == 1.0 && (P.is_linear || err <= T3_EXPECT_CONVERGED_BELOW);
            expect_converged = !__syncthreads_or(active && !expect);
        }
#ifdef MMPC_T4_PROFILE
        tp_post += T4_CLK() - tf1;
#endif
    }
#ifdef MMPC_T4_PROFILE
    const long long tk1 = T4_CLK();
#endif
    // ---- outputs: the reference's V layout, instance-major.  The warm start lives in LAM2: after an odd number of accepted
    // steps the multipliers are copied back there ----
    double* const Lcan = W.LAM2 + (size_t)tile * I * Mc + i;
    const bool back = c.LAM != Lcan;  // tile-uniform
    if (worker) {
        const double* LAMp = c.LAM;
        constexpr int UB = 8;
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
#ifdef MMPC_T4_PROFILE
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0)
        printf("[t4 kernel] prologue %lld  loop %lld (residual sweeps %lld, backward %lld, forward %lld, line search + step %lld)  epilogue %lld cycles, %d iterations\n",
               tk0 - t_entry, tk1 - tk0, tp_res, tp_bwd, tp_fwd, tp_post, T4_CLK() - tk1, it);
#endif
    if (valid && master && io.stats) {
        mmpc_stats s;
        s.objective = fval; s.kkt_error = err; s.mu = 0.0; s.iterations = it_done; s.status = status;
        s.ls_steps = ls_total; s.reg_steps = reg_total;
        io.stats[b] = s;
    }
}

template <class M>
struct LaunchT4 {
    template <int I_>
    static cudaError_t go(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        using C = T4<M, I_>;
        static_assert(C::SMEM_BYTES <= 227 * 1024, "ring + scratch must fit the 227 KB of shared memory a CTA can opt into");
        static std::atomic<unsigned long long> attr_done{0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ULL)) {
            cudaError_t e = cudaFuncSetAttribute(solve_kernel_t4<M, I_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            attr_done.fetch_or(1ULL << (dev & 63), std::memory_order_release);
        }
        if (io.b0 % C::I != 0) return cudaErrorInvalidValue;  // sub-batches must start on a tile boundary
        const int t0 = io.b0 / C::I, t1 = (io.b1 + C::I - 1) / C::I;
        solve_kernel_t4<M, I_><<<t1 - t0, C::THREADS, C::SMEM_BYTES, st>>>(P, W, io);
        return cudaGetLastError();
    }
    // tile width from the batch of the handle (it fixes the layout of the persistent iterate): the narrowest tile that still
    // covers the batch with one CTA per SM (148 SMs), so that every SM works and every instance gets the most threads.
    //   I = 4  (96 node slots: a whole horizon of N <= 96 is evaluated in ONE round)   B <= 592     -- the latency configuration
    //   I = 14 (27 node slots)                                                         B <= 2072
    //   I = 28 (13 node slots, 147 CTAs at B = 4096)                                   beyond
    // kernel_variant = 4 + 256*k forces configuration k (1, 2, 3).
    static cudaError_t run(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        if (P.bounded || (P.integrator == MMPC_RK4 && !P.is_linear)) return cudaErrorNotSupported;
        if constexpr (M::NX == 4 && M::NX + 2 * M::NU <= 8) {
            const int I = mmpc_t4_tile_width(W.B, io.variant >> 8);
            if (I == 4) return go<4>(P, W, io, st);
            if (I == 14) return go<14>(P, W, io, st);
            return go<28>(P, W, io, st);
        } else {
            return cudaErrorNotSupported;
        }
    }
};

}  // namespace mmpc

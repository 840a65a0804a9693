// mahi-mpc-b200 :: solve kernel, variant T4 -- T3's CTA-per-tile organisation (warp 0: one lane per instance runs the
// Riccati recursion, worker warps evaluate the model) with the iteration re-ordered around what the in-kernel timers
// showed for T3 at the headline batch (profiles/r1_t3_phase_timers.txt: per Newton iteration backward 84 k, forward 50 k,
// line search + step application 25 k cycles; the forward sweep was bound by the merit evaluation of the trial point in
// the worker warps and by a 760-cycle serial stage):
//
//   * SPECULATIVE FULL STEP ("lazy merit").  The Armijo test of the full Newton step needs the objective and the constraint
//     violation of z + dz -- exactly what the NEXT sweep (backward sweep or convergence-test residual sweep) computes at
//     its iterate anyway.  So the forward sweep applies the full step at once and evaluates nothing; the test is made one
//     sweep later from that sweep's own norms.  If it fails for some instance of the tile (no instance of the benchmark
//     workload: `ls_steps` is 0 there) the tile takes the slow path: ell_1-merit backtracking on z + alpha dz for the
//     instances that failed -- dz and the previous multipliers are still there --, then the sweep is repeated.  The
//     iterates are those of T1-T3 (same acceptance rule, same step lengths).
//   * FORWARD SWEEP = the dependent chain and nothing else.  The serial lane of an instance rolls out
//     du_k = K_k (dx_k, du_{k-1}) + k_k, dx_{k+1} = Fx dx_k + Fu du_k + c_k from records staged with cp.async (Fx, Fu, c
//     and the gains only: 32 of the 64 doubles of a node record) -- ~26 DFMA per stage with short accumulation trees --
//     and writes dz.  Everything that hangs off dz is done one chunk later by the worker threads, one (instance, node)
//     each, straight from global memory: new multipliers lam_k = (P_{k+1} (dx_{k+1}, du_k) + p_{k+1})_x, directional
//     derivative of the objective, the step itself (z += dz) and the multiplier hand-over.  No separate pass over the
//     iterate, no merit evaluation, no value-function rows in shared memory.
//   * tile widths 4 (64 threads, 8 node slots per instance: batches up to 1184) and 32 (larger batches).
// Backward sweep, residual sweep and merit pass are T3's (mmpc_solve_t3.cuh).  Euler / linearised / RK4 shooting, no
// finite bounds (those stay on T3 / T1).  Replaces IPOPT+MUMPS at ModelControl.cpp:159.
#pragma once
#if defined(MMPC_T4_PROFILE)
#include <cstdio>
#define T4_CLK() clock64()
#else
#define T4_CLK() 0LL
#endif

namespace mmpc {

template <class M, int I_, int CB_, int CF_, int RK_ = 0, int SK_ = 0>
struct T4 : T3<M, I_, CB_, CF_, 0, RK_> {
    using Base = T3<M, I_, CB_, CF_, 0, RK_>;
    static_assert((RK_ & 2) == 0, "T4 has no interior-point mode");
    // SK: the warps that share the serial warp's SM sub-partition (warp ids 4, 8, ...) stay idle, so that the Riccati recursion
    // -- bound by the issue rate of ONE warp's fp64 instructions -- has its sub-partition's fp64 pipe to itself.  Worker warp n
    // then is warp 1 + n + n / 3.
    static constexpr int SK = SK_;
    static constexpr int NWW = (I_ * CB_ + 31) / 32;
    static_assert(!SK_ || (I_ * CB_) % 32 == 0, "SK: whole worker warps");
    static constexpr int THREADS = SK_ ? 32 * (2 + (NWW - 1) + (NWW - 1) / 3) : 32 + I_ * CB_;
    __device__ static __forceinline__ bool is_worker(int tid) { return SK_ ? (((tid >> 5) & 3) != 0) : tid >= 32; }
    __device__ static __forceinline__ int widx(int tid) { return SK_ ? tid - 32 * (1 + (tid >> 7)) : tid - 32; }
    __device__ static __forceinline__ int slot_of(int g) { return SK_ ? g : Base::slot_of(g); }
    // ONE global record per node and tile, [field][instance]:  Fx | Fu | c | gains  ||  q~ | value-function rows  -- the first
    // REC_C rows are what the forward roll-out needs (chain part), the other REC_P rows what the post-processing of the step
    // needs (post part): each part of a node is one contiguous block, moved by one bulk copy (TMA).  T3's sweeps write the
    // record through the same field names: FG_Q (where q~ starts) and the node strides are overridden accordingly.
    static constexpr int RS_F = Base::FG_C + Base::NX;  // Fx | Fu | c
    static constexpr int REC_C = RS_F + Base::KG, REC_P = Base::N2 + Base::PG, REC = REC_C + REC_P;
    static constexpr int FG_Q = REC_C, FG_NS = REC, KG_NS = REC, PG_NS = REC;
    static constexpr int REC_K = RS_F, REC_PG = REC_C + Base::N2;  // row offsets of the gains and the value-function rows
    // Forward sweep, shared memory (an overlay of the backward sweep's record buffers).  Per chunk parity:
    //   chain part  CF x REC_C x I
    //   post part   CF x REC_P x I
    //   dz          (CF + 1) x S x I   slot 0 = (du, dx) entering the chunk, slot s + 1 = (du_k, dx_{k+1}) of its node s
    static constexpr int RC = REC_C, RC_K = RS_F, RP_Q = 0, RP_P = Base::N2, RP = REC_P;
    static constexpr int FW_CHAIN = 0, FW_POST = CF_ * RC * I_, FW_DZ = FW_POST + CF_ * RP * I_, FWD = FW_DZ + (CF_ + 1) * Base::S * I_;
    // shared-memory map: T3's, with record buffers sized for the backward record and THIS forward layout (these shadow the
    // base's constants: the sweeps T4 borrows from T3 read them through C::)
    static constexpr int BUF = CB_ * Base::RB * I_ > FWD ? CB_ * Base::RB * I_ : FWD;
    static constexpr int SM_REC = 0, SM_RED = 2 * BUF, SM_INST = SM_RED + 2 * Base::REDBUF, SM_BAR = SM_INST + 6 * I_, SM_END = SM_BAR + 4;
    static constexpr size_t SMEM_BYTES = sizeof(double) * SM_END;
    static_assert((I_ % 2) == 0, "bulk copies move multiples of 16 bytes: rows of an even number of instances");
};

#ifdef MMPC_T4_PROFILE
__device__ long long t4_dbg[16];
#define T4_STAMP(slot, t_from) do { if (blockIdx.x == 0 && tid == 0) { const long long t_now = clock64(); t4_dbg[slot] += t_now - (t_from); (t_from) = t_now; } } while (0)
#define T4_STAMPW(slot, t_from) do { if (blockIdx.x == 0 && tid == 32) { const long long t_now = clock64(); t4_dbg[slot] += t_now - (t_from); (t_from) = t_now; } } while (0)
#else
#define T4_STAMP(slot, t_from) do { } while (0)
#define T4_STAMPW(slot, t_from) do { } while (0)
#endif

__device__ __forceinline__ void t4_discard_l2(const void* line128) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(line128) : "memory"); }

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ unsigned t4_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t4_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t4_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void t4_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t4_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t4_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "T4_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra T4_DONE;\n"
        "bra T4_WAIT;\n"
        "T4_DONE:\n"
        "}\n" ::"r"(t4_smem_u32(bar)),
        "r"(parity), "r"(0x989680u)  // suspend-time hint: the waiting warp sleeps in hardware instead of polling shared memory
        : "memory");
}
// one contiguous block global -> shared through the TMA unit; completion (bytes) is signalled on `bar`
__device__ __forceinline__ void t4_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(t4_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(t4_smem_u32(bar))
                 : "memory");
}
// generic-proxy accesses (st.shared / ld.shared of the record buffers, st.global of the node records) are ordered before the
// async proxy (the bulk copies) touches the same bytes
__device__ __forceinline__ void t4_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void t4_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct T4Fwd {
    double gradf_d, lam_inf;
    bool finite;
};

// Forward sweep for the tile.  Step ch of the pipeline (one CTA barrier per step):
//   TMA     chain part of chunk ch+1 and post part of chunk ch -> shared memory (issued by the first worker thread)
//   warp 0  rolls the Newton step through chunk ch: lane i = instance i, chain part of parity ch & 1 -> dz (shared + global)
//   workers post-process chunk ch-1, thread = (instance, node): lam+, grad f . d, z += dz -- from shared memory but for z
// `phase` counts the completed uses of the four mbarriers (chain / post x parity): they are initialised once per kernel.
template <class M, class C>
__device__ __noinline__ void t4_forward(const T3Ctx<M, C::I, C::MODE>& cin, double* smem, int tid, T4Fwd& o_out, const double* fg_tile,
                                        const double* kg_tile, const double* pg_tile, unsigned (&phase)[4]) {
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I, N2 = C::N2;
#ifdef MMPC_T4_PROFILE
    long long t_st = clock64();
    if (blockIdx.x == 0 && tid == 0) t4_dbg[0] += t_st - t4_dbg[15];  // call -> first statement
#endif
    const T3Ctx<M, C::I, C::MODE> c = cin;  // register-resident copy (see t3_backward)
    const int N = c.N;
    const bool worker = C::is_worker(tid);
    const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
    double* red = smem + C::SM_RED;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);  // [0,1] chain, [2,3] post
    const int nch = (N + C::CF - 1) / C::CF;
    constexpr unsigned B_CHAIN = C::REC_C * I * 8, B_POST = C::REC_P * I * 8;
    const double* rec_tile = fg_tile;  // FG | KG | PG are ONE record per node here (kg_tile, pg_tile point into it)
    (void)kg_tile; (void)pg_tile;
    // bulk copies of step ch: the chain parts of chunk ch+1 and the post parts of chunk ch, one copy per lane (a single thread
    // needs ~90 cycles per copy to issue them back to back); the first lane announces the bytes.  A copy may complete before the
    // announcement: the phase cannot end before the announcing thread's arrival.  The issuing lanes are the workers of the LAST
    // node slot when that slot has no node to post-process (CF < CB: it is beyond the forward chunk), so that issuing
    // (~0.9 k cycles per step) is not in front of any thread's post-processing.
#ifdef MMPC_T4_ISSUE_FIRST_WARP
    constexpr bool SPARE_SLOT = false;
#else
    constexpr bool SPARE_SLOT = C::CB - 1 >= C::CF && I >= 2 * C::CF;
#endif
    const int wl = SPARE_SLOT ? (wj == C::CB - 1 ? wi : -1) : w;  // lane of the issuing group (-1: not an issuer)
    auto issue_loads = [&](int ch_chain, int ch_post) {
        const int knc = ch_chain < nch ? min(C::CF, N - ch_chain * C::CF) : 0, knp = ch_post >= 0 ? min(C::CF, N - ch_post * C::CF) : 0;
        if (wl == 0) {
            if (knc) t4_mbar_arrive_expect_tx(&bars[ch_chain & 1], (unsigned)knc * B_CHAIN);
            if (knp) t4_mbar_arrive_expect_tx(&bars[2 + (ch_post & 1)], (unsigned)knp * B_POST);
        }
        if (wl >= 0 && wl < knc) {
            const int k = ch_chain * C::CF + wl;
            t4_bulk_load(smem + C::SM_REC + (ch_chain & 1) * C::BUF + C::FW_CHAIN + (size_t)wl * C::RC * I, rec_tile + (size_t)k * C::REC * I, B_CHAIN, &bars[ch_chain & 1]);
        } else if (wl >= C::CF && wl - C::CF < knp) {
            const int s = wl - C::CF, k = ch_post * C::CF + s;
            t4_bulk_load(smem + C::SM_REC + (ch_post & 1) * C::BUF + C::FW_POST + (size_t)s * C::RP * I, rec_tile + ((size_t)k * C::REC + C::REC_C) * I, B_POST,
                         &bars[2 + (ch_post & 1)]);
        }
    };
    // does this thread's instance still iterate?  (instances that do not: their multipliers are handed over unchanged)
    const bool act = smem[C::SM_INST + 2 * I + (worker ? wi : (tid < I ? tid : 0))] != 0.0;
    double gsum = 0.0, lmax = 0.0;
    bool fin = true;
    // worker (i, node k = k0 + s of chunk ch): everything that depends on dz_k, dz_{k+1} but is not on the roll-out's chain
    auto post_node = [&](int ch, int s) {
        const int k = ch * C::CF + s;
        const double* pb = smem + C::SM_REC + (ch & 1) * C::BUF + C::FW_POST + (size_t)s * C::RP * I + wi;
        const double* dzb = smem + C::SM_REC + (ch & 1) * C::BUF + C::FW_DZ + (size_t)s * S * I + wi;  // slot s: (dv, dx); slot s+1: (du, dx+)
        double* z = c.V + (size_t)k * S * I;
        double zu[NU], zx[NX], lk[NX];
        MMPC_UNROLL for (int a = 0; a < NU; ++a) zu[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NX; ++a) zx[a] = lk[a] = 0.0;
        if (act) {
            MMPC_UNROLL for (int a = 0; a < NU; ++a) zu[a] = z[(NX + a) * I];
            MMPC_UNROLL for (int a = 0; a < NX; ++a) zx[a] = z[(S + a) * I];
        } else {
            MMPC_UNROLL for (int a = 0; a < NX; ++a) lk[a] = c.LAM[((size_t)k * NX + a) * I];
        }
        double dx[NX], dv[NU], du[NU], dxn[NX];
        MMPC_UNROLL for (int a = 0; a < NU; ++a) { dv[a] = dzb[a * I]; du[a] = dzb[(S + a) * I]; }
        MMPC_UNROLL for (int a = 0; a < NX; ++a) { dx[a] = dzb[(NU + a) * I]; dxn[a] = dzb[(S + NU + a) * I]; }
        // directional derivative of the objective: q~ . (dx, dv, du)
        MMPC_UNROLL for (int a = 0; a < NX; ++a) gsum = fma(pb[(C::RP_Q + a) * I], dx[a], gsum);
        MMPC_UNROLL for (int a = 0; a < NU; ++a) {
            gsum = fma(pb[(C::RP_Q + NX + a) * I], dv[a], gsum);
            gsum = fma(pb[(C::RP_Q + S + a) * I], du[a], gsum);
        }
        if (c.node_cost && k == N - 1) {  // MMPC_COST_NODE: x_N's tracking term is in no node record (terminal value function)
            MMPC_UNROLL for (int a = 0; a < NX; ++a)
                gsum = fma(2.0 * c.Q[a] * (z[(S + a) * I] - c.traj[(size_t)k * NX + a]), dxn[a], gsum);
        }
        // lam_k+ = (P_{k+1} (dx_{k+1}, du_k) + p_{k+1})_x
        MMPC_UNROLL for (int a = 0; a < NX; ++a) {
            double t = pb[(C::RP_P + C::PG_X + a) * I];
            MMPC_UNROLL for (int b = 0; b < NX; ++b) t = fma(pb[(C::RP_P + C::PG_XX + trisym(a, b, NX)) * I], dxn[b], t);
            MMPC_UNROLL for (int b = 0; b < NU; ++b) t = fma(pb[(C::RP_P + C::PG_XV + a * NU + b) * I], du[b], t);
            lmax = fmax(lmax, fabs(t));
            fin = fin && isfinite(t) && isfinite(dxn[a]);
            c.LAMN[((size_t)k * NX + a) * I] = act ? t : lk[a];
        }
        if (act) {  // the full step, applied at once (its Armijo test follows the next sweep)
            MMPC_UNROLL for (int a = 0; a < NU; ++a) z[(NX + a) * I] = zu[a] + du[a];
            MMPC_UNROLL for (int a = 0; a < NX; ++a) z[(S + a) * I] = zx[a] + dxn[a];
            // global copy of the step: only the (rare) backtracking of t4_armijo reads it
            MMPC_UNROLL for (int a = 0; a < NU; ++a) c.DV[((size_t)k * S + NX + a) * I] = du[a];
            MMPC_UNROLL for (int a = 0; a < NX; ++a) c.DV[((size_t)(k + 1) * S + a) * I] = dxn[a];
        }
#ifndef MMPC_T4_NO_DISCARD
        // The node's record (FG | KG | PG rows of this tile) is dead now -- its bulk copies completed before this step -- and
        // the next backward sweep overwrites every byte of it.  Tell L2 so: the lines are dropped instead of being written back
        // to HBM when they are evicted (the record stream was ~1 GB of DRAM traffic per launch,
        // profiles/r1_bench_solve_t3_ncu_full.txt).  Whole 128-byte lines only: rows of 32 instances.
        if constexpr (I == 32) {
            const char* recb = reinterpret_cast<const char*>(rec_tile + (size_t)k * C::REC * I);
            MMPC_UNROLL for (int l = 0; l < 2 * C::REC; l += 32) if (l + wi < 2 * C::REC) t4_discard_l2(recb + (size_t)(l + wi) * 128);
        }
#endif
    };
    auto post_chunk = [&](int ch) {
        const int kn = min(C::CF, N - ch * C::CF);
        t4_mbar_wait(&bars[2 + (ch & 1)], (phase[2 + (ch & 1)] + (unsigned)(ch >> 1)) & 1u);
        if (wj < kn) post_node(ch, wj);
    };
#ifdef MMPC_T4_PROFILE
    long long t_work = 0;
#endif
    // the record buffers were last touched by the backward sweep's st.shared / ld.shared and the node records were written
    // with st.global: order all of it before the bulk copies
    t4_fence_proxy_async();
    __syncwarp();
    __syncthreads();
    if (worker) issue_loads(0, -1);
    T4_STAMP(1, t_st);
    if (worker) {
        for (int ch = 0; ch < nch; ++ch) {
            const long long t0 = T4_CLK();
#ifdef MMPC_T4_PROFILE
            long long tw = t0;
#endif
            issue_loads(ch + 1, ch);
            if (wj < min(C::CF, N - ch * C::CF)) {  // what the NEXT step's post-processing reads from global memory -> L1, one step ahead
                const int k = ch * C::CF + wj;
                if (act) { MMPC_UNROLL for (int a = 0; a < S; ++a) t3_prefetch(c.V + ((size_t)k * S + NX + a) * I); }
                else { MMPC_UNROLL for (int a = 0; a < NX; ++a) t3_prefetch(c.LAM + ((size_t)k * NX + a) * I); }
            }
            T4_STAMPW(6, tw);
            if (ch > 0) post_chunk(ch - 1);
            T4_STAMPW(7, tw);
#ifdef MMPC_T4_PROFILE
            t_work += T4_CLK() - t0;
#else
            (void)t0;
#endif
            t3_cta_barrier();  // dz of chunk ch is in shared memory; the buffers of parity (ch+1)&1 are free again
            T4_STAMPW(9, tw);
        }
        post_chunk(nch - 1);
        red[(0 * C::CB + wj) * I + wi] = gsum;
        red[(1 * C::CB + wj) * I + wi] = lmax;
        red[(2 * C::CB + wj) * I + wi] = fin ? 1.0 : 0.0;
    } else {
        double dx[NX], dv[NU];
        MMPC_UNROLL for (int a = 0; a < NX; ++a) dx[a] = 0.0;
        MMPC_UNROLL for (int a = 0; a < NU; ++a) dv[a] = 0.0;
        for (int ch = 0; ch < nch; ++ch) {
            const long long t0 = T4_CLK();
#ifdef MMPC_T4_PROFILE
            long long ts = t0;
#endif
            if (tid < I) {
                const int k0 = ch * C::CF, kn = min(C::CF, N - k0);
                const double* __restrict__ rec = smem + C::SM_REC + (ch & 1) * C::BUF + C::FW_CHAIN;
                double* __restrict__ dzb = smem + C::SM_REC + (ch & 1) * C::BUF + C::FW_DZ + tid;
                MMPC_UNROLL for (int a = 0; a < NU; ++a) dzb[a * I] = dv[a];
                MMPC_UNROLL for (int a = 0; a < NX; ++a) dzb[(NU + a) * I] = dx[a];
                t4_mbar_wait(&bars[ch & 1], (phase[ch & 1] + (unsigned)(ch >> 1)) & 1u);
                T4_STAMP(12, ts);
                MMPC_UNROLL for (int s = 0; s < C::CF; ++s) if (s < kn) {
                    const double* __restrict__ const recp = rec + (size_t)s * C::RC * I + tid;
#define T4_F(f) recp[(f) * I]
                    // short accumulation trees: the stage's dependent chain is du (2 fma + 2 add after dx) -> dx+ (<= NU fma)
                    double du[NU], dxn[NX];
                    MMPC_UNROLL for (int a = 0; a < NU; ++a) {
                        double ta = T4_F(C::RC_K + C::KG_K0 + a), tb = 0.0, tc = 0.0;
                        MMPC_UNROLL for (int q = 0; q < NX; ++q) {
                            if (q < (NX + 1) / 2) ta = fma(T4_F(C::RC_K + C::KG_KX + a * NX + q), dx[q], ta);
                            else tb = fma(T4_F(C::RC_K + C::KG_KX + a * NX + q), dx[q], tb);
                        }
                        MMPC_UNROLL for (int q = 0; q < NU; ++q) tc = fma(T4_F(C::RC_K + C::KG_KV + a * NU + q), dv[q], tc);
                        du[a] = (ta + tb) + tc;
                    }
                    MMPC_UNROLL for (int a = 0; a < NX; ++a) {
                        double t = T4_F(C::FG_C + a);
                        MMPC_UNROLL for (int q = 0; q < NX; ++q) if (T3_FXNZ(a, q)) t = fma(T4_F(C::FG_FX + T3_FXSLOT(a * NX + q)), dx[q], t);
                        MMPC_UNROLL for (int q = 0; q < NU; ++q) if (T3_FUNZ(a, q)) t = fma(T4_F(C::FG_FU + T3_FUSLOT(a * NU + q)), du[q], t);
                        dxn[a] = t;
                    }
#undef T4_F
                    double* const dzs = dzb + (size_t)(s + 1) * S * I;
                    MMPC_UNROLL for (int a = 0; a < NU; ++a) dzs[a * I] = du[a];
                    MMPC_UNROLL for (int a = 0; a < NX; ++a) dzs[(NU + a) * I] = dxn[a];
                    MMPC_UNROLL for (int a = 0; a < NX; ++a) dx[a] = dxn[a];
                    MMPC_UNROLL for (int a = 0; a < NU; ++a) dv[a] = du[a];
                }
            }
#ifdef MMPC_T4_PROFILE
            t_work += T4_CLK() - t0;
            T4_STAMP(10, ts);
#else
            (void)t0;
#endif
            t3_cta_barrier();
            T4_STAMP(11, ts);
        }
    }
    // every barrier has completed ceil((nch - parity) / 2) more phases
    phase[0] += (unsigned)((nch + 1) >> 1); phase[1] += (unsigned)(nch >> 1);
    phase[2] += (unsigned)((nch + 1) >> 1); phase[3] += (unsigned)(nch >> 1);
    T4_STAMP(2, t_st);  // pipeline loop
    __syncwarp();
    __syncthreads();
    T4_STAMP(3, t_st);  // last chunk's post-processing (workers)
    T4Fwd o;
    o.gradf_d = o.lam_inf = 0.0;
    o.finite = true;
    if (tid < I) MMPC_UNROLL for (int s = 0; s < C::CF; ++s) {
        o.gradf_d += red[(0 * C::CB + s) * I + tid];
        o.lam_inf = fmax(o.lam_inf, red[(1 * C::CB + s) * I + tid]);
        o.finite = o.finite && red[(2 * C::CB + s) * I + tid] != 0.0;
    }
    __syncthreads();
    o_out = o;
    T4_STAMP(4, t_st);  // reduction
#ifdef MMPC_T4_PROFILE
    if (blockIdx.x == 0 && tid == 0) t4_dbg[14] = clock64();
    (void)t_work;
#endif
}

// Armijo test of the speculative full step, made from the norms (f1, th1) the sweep at z + dz has just produced.  Returns
// (tile-uniform) whether some instance failed it; then those instances have been moved to z + alpha dz by backtracking --
// V holds z + dz, so the trial points are V + (alpha - 1) dz -- their multipliers blended accordingly (LAM holds lam+, LAMN
// the previous ones), and the caller repeats the sweep.  Rare (never on the benchmark workload), hence not inlined.
template <class M, class C>
__device__ __noinline__ bool t4_armijo(const T3Ctx<M, C::I, C::MODE>& c, double* smem, int tid, bool pending, double f1, double th1, double nu_merit,
                                       double m0, double D, int NV, int Mc, int& ls_total) {
    constexpr int NX = M::NX, I = C::I;
    const bool failed = pending && !(f1 + nu_merit * th1 <= m0 + 1e-4 * D + 1e-12 * fabs(m0));
    if (!__syncthreads_or(failed)) return false;
    double* sh_alpha = smem + C::SM_INST + I;
    double alpha = 1.0;
    int ls = 0;
    bool need = failed;
    while (__syncthreads_or(need)) {
        if (need) { alpha *= 0.5; ++ls_total; }
        if (tid < I) sh_alpha[tid] = need ? alpha - 1.0 : 0.0;
        __syncthreads();
        double f2, t2, b2;
        t3_trial<M, C>(c, smem, tid, f2, t2, b2);
        if (need) need = !(f2 + nu_merit * t2 <= m0 + 1e-4 * alpha * D + 1e-12 * fabs(m0)) && (++ls < 40);
    }
    if (tid < I) sh_alpha[tid] = failed ? alpha - 1.0 : 0.0;
    __syncthreads();
    if (C::is_worker(tid)) {
        const int w = C::widx(tid), wi = w % I, wj = C::slot_of(w / I);
        const double al = sh_alpha[wi];
        if (al != 0.0) {
            double* Vp = c.V; const double* DVp = c.DV; double* L = c.LAM; const double* Lo = c.LAMN;
            for (int e = NX + wj; e < NV; e += C::CB) Vp[(size_t)e * I] = fma(al, DVp[(size_t)e * I], Vp[(size_t)e * I]);
            for (int e = wj; e < Mc; e += C::CB) L[(size_t)e * I] = fma(al + 1.0, L[(size_t)e * I] - Lo[(size_t)e * I], Lo[(size_t)e * I]);
        }
    }
    __syncwarp();
    __syncthreads();
    return true;
}

// HO_: the instantiation that can write its outputs to host memory in place (zero-copy epilogue).  The kernels sit at 255
// registers, and the mere presence of the epilogue call costs the device-resident solve 2 % (profiles/r2_zero_copy_experiments.txt):
// the default tile shapes are therefore instantiated twice, and the launcher picks by io.host_out.
template <class M, int I_, int CB_, int CF_, int RK_, int SK_, bool HO_ = true>
__global__ void __launch_bounds__(T4<M, I_, CB_, CF_, RK_, SK_>::THREADS) solve_kernel_t4(mmpc_problem P, mmpc_workspace W, mmpc_solve_io io) {
    using C = T4<M, I_, CB_, CF_, RK_, SK_>;
    constexpr int NX = M::NX, NU = M::NU, S = C::S, I = C::I;
    extern __shared__ __align__(16) double smem[];
#ifdef MMPC_T4_PROFILE
    const long long t_entry = clock64();
    long long tp_res = 0, tp_bwd = 0, tp_fwd = 0, tp_post = 0;
#endif
    const int tid = threadIdx.x, tile = io.b0 / C::I + blockIdx.x;  // b0 is a multiple of the tile width (the ABI layer cuts sub-batches on tile boundaries)
    const bool worker = C::is_worker(tid);
    // mbarriers of the forward sweep's bulk copies: initialised once, `fwd_phase` counts their completed phases
    if (tid == 0) {
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + C::SM_BAR);
        for (int q = 0; q < 4; ++q) t4_mbar_init(&bars[q], 1u);
        t4_fence_barrier_init();
    }
    unsigned fwd_phase[4] = {0u, 0u, 0u, 0u};
    // a thread's instance: serial lane tid -> instance tid; worker w -> instance w % I (node slot w / I)
    const int i = worker ? C::widx(tid) % I : (tid < I ? tid : 0);
    const int j = worker ? C::slot_of(C::widx(tid) / I) : 0;
    const int N = P.N, NV = NX * (N + 1) + NU * N, Mc = N * NX;
    const int b = tile * I + i;
    const bool valid = b >= io.b0 && b < io.b1;
    const int bi = valid ? b : io.b1 - 1;  // surplus slots read the last instance's inputs and never write outputs

    T3Ctx<M, I, C::MODE> c;
    c.N = N;
    c.h = P.h; c.is_linear = P.is_linear; c.substeps = P.substeps; c.gauss_newton = P.gauss_newton; c.node_cost = P.node_cost;
    // the iterate lives tile-wise, [element][instance] (T3's layout)
    c.V = W.V2 + (size_t)tile * I * NV + i; c.LAM = W.LAM2 + (size_t)tile * I * Mc + i;
    c.DV = W.DV2 + (size_t)tile * I * NV + i; c.LAMN = W.LAMN2 + (size_t)tile * I * Mc + i;
    c.LINp = W.LIN2 + (size_t)b * C::LIN;
    const double *in_state = io.state, *in_control = io.control;
    c.traj = io.traj + (size_t)bi * N * NX;
    if (io.traj_stage) {  // zero-copy inputs (uniform over the grid): the tile's rows are brought on the device first
        t3_stage_inputs(io.state, io.control, io.traj, io.state_stage, io.control_stage, io.traj_stage, max(tile * I, io.b0),
                        min(tile * I + I, io.b1), NX, NU, N);
        in_state = io.state_stage; in_control = io.control_stage;
        c.traj = io.traj_stage + (size_t)bi * N * NX;
    }
    // node records of the tile: FG2, KG2, PG2 are one contiguous region of the arena (mmpc_api.cu), used here as [node][REC rows][I]
    double* fg_tile = W.FG2 + (size_t)tile * I * N * C::REC;
    double* kg_tile = fg_tile + (size_t)C::REC_K * I;
    double* pg_tile = fg_tile + (size_t)C::REC_PG * I;
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
    __syncwarp();
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

    const double* const Vp = c.V;
    // per-instance solver state lives in the serial lane (tid < I) of the instance
    double nu_merit = 1.0, delta_last = 0.0, err = INFINITY, fval = 0.0;
    double arm_m0 = 0.0, arm_D = 0.0;  // merit and its directional derivative at the point the pending step was taken from
    bool pending = false;              // a full step has been applied whose Armijo test is still to be made
    int it = 0, it_done = 0, status = MMPC_SOLVE_MAX_ITER, ls_total = 0, reg_total = 0;
    bool active = valid && tid < I;
    double* sh_flag = smem + C::SM_INST + 2 * I;  // per-instance "still iterates" flag
    const int nfree = N * S, ncon = N * NX;
    bool expect_converged = false;  // tile-uniform
#ifdef MMPC_T4_PROFILE
    const long long tk0 = clock64();
#endif
    for (it = 0;;) {
        T3Norms nrm;
        double delta = 0.0;
        bool ok = true;
        if (expect_converged) {
            // every instance of the tile that still iterates took its step from a small KKT error: the new iterate is expected to
            // pass the convergence test, which needs first derivatives only (T3).  The same sweep settles the step's Armijo test.
            T3NormsBD nb;
            const long long tr0 = T4_CLK();
            t3_residuals<M, C>(c, smem, tid, nb);
#ifdef MMPC_T4_PROFILE
            tp_res += T4_CLK() - tr0;
#else
            (void)tr0;
#endif
            if (t4_armijo<M, C>(c, smem, tid, pending && active, nb.fval, nb.theta, nu_merit, arm_m0, arm_D, NV, Mc, ls_total)) {
                pending = false;
                expect_converged = false;
                continue;
            }
            pending = false;
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
        const long long tb0 = T4_CLK();
        ok = t3_backward<M, C>(c, smem, tid, delta, nrm, fg_tile, kg_tile, pg_tile);
#ifdef MMPC_T4_PROFILE
        tp_bwd += T4_CLK() - tb0;
#else
        (void)tb0;
#endif
        if (t4_armijo<M, C>(c, smem, tid, pending && active, nrm.fval, nrm.theta, nu_merit, arm_m0, arm_D, NV, Mc, ls_total)) {
            pending = false;
            continue;  // some instance was moved back along its step: the sweep is repeated at the tile's new iterates
        }
        pending = false;
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
            const bool ok2 = t3_backward<M, C>(c, smem, tid, delta, dummy, fg_tile, kg_tile, pg_tile);
            if (active && !ok) ok = ok2;
            if (++tries >= 24) {
                if (active && !ok) { status = MMPC_SOLVE_NUMERICAL; active = false; }
                break;
            }
        }
        if (delta > 0.0) delta_last = delta;
        if (tid < I) sh_flag[tid] = active ? 1.0 : 0.0;  // which instances still iterate: read by the forward sweep
        if (!__syncthreads_or(active)) break;
        T4Fwd fo;
        const long long tf0 = T4_CLK();
#ifdef MMPC_T4_PROFILE
        if (blockIdx.x == 0 && tid == 0) t4_dbg[15] = tf0;
#endif
        t4_forward<M, C>(c, smem, tid, fo, fg_tile, kg_tile, pg_tile, fwd_phase);
        const long long tf1 = T4_CLK();
#ifdef MMPC_T4_PROFILE
        if (blockIdx.x == 0 && tid == 0) t4_dbg[5] += tf1 - t4_dbg[14];  // last statement -> back in the caller
#endif
#ifdef MMPC_T4_PROFILE
        tp_fwd += tf1 - tf0;
#else
        (void)tf0; (void)tf1;
#endif
        if (active && !fo.finite) { status = MMPC_SOLVE_NUMERICAL; active = false; }
        nu_merit = fmax(nu_merit, 1.1 * fo.lam_inf + 1e-8);
        arm_m0 = nrm.fval + nu_merit * nrm.theta;
        arm_D = fo.gradf_d - nu_merit * nrm.theta;
        pending = active;
        { double* t = c.LAM; c.LAM = c.LAMN; c.LAMN = t; }  // LAMN holds lam+ (instances that iterate) or the old values (the others)
        ++it;
        // Is the next iterate expected to meet tol?  Then its test is done by the cheap residual sweep first (T3's rule).
        {
            const bool newton = C::RK == 0 || !P.gauss_newton;  // Gauss-Newton under RK4: linear rate, not predicted
            const bool expect = newton && (P.is_linear || err <= T3_EXPECT_CONVERGED_BELOW);
            expect_converged = !__syncthreads_or(active && !expect);
        }
#ifdef MMPC_T4_PROFILE
        tp_post += T4_CLK() - tf1;
#endif
    }
#ifdef MMPC_T4_PROFILE
    const long long tk1 = clock64();
#endif
    // ---- outputs: the reference's V layout, instance-major.  The warm start lives in LAM2: after an odd number of accepted
    // steps the multipliers are copied back there ----
    double* const Lcan = W.LAM2 + (size_t)tile * I * Mc + i;
    const bool back = c.LAM != Lcan;  // tile-uniform
    // Zero-copy outputs (io.host_out, grid-uniform): the transposing epilogue t3_outputs_to_host_s runs instead of the plain one.  It is
    // called at the very end, without an early return and without solver state among its arguments (the statistics go through
    // shared memory): measured on one box at B = 4096, this form leaves the device-resident solve where it was (0.307 ms), an early
    // `if (host_out) { ...; return; }` in front of the plain epilogue cost it 2.5 % (0.315 ms) -- profiles/r2_latency.txt.
#ifdef MMPC_WIDE_TILES_NO_HOSTOUT
    constexpr bool HOST_OUT = HO_ && I <= 8;  // experiment: the ABI layer fetches wide tiles' outputs with a copy kernel (MMPC_ZC_OUT_KERNEL)
#else
    constexpr bool HOST_OUT = HO_;
#endif
    const bool host_out = HOST_OUT && io.host_out != 0;
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
        s.objective = fval; s.kkt_error = err; s.mu = 0.0; s.iterations = it_done; s.status = status;
        s.ls_steps = ls_total; s.reg_steps = reg_total;
        mmpc_stats* dst = host_out ? reinterpret_cast<mmpc_stats*>(smem + I * 65) + tid : io.stats + b;
        *dst = s;
    }
    if constexpr (HOST_OUT) if (host_out)
        t3_outputs_to_host_s(smem, io.V_out, io.lam_out, io.stats, io.b1, tile, I, NV, Mc, Vp - i, c.LAM - i, back ? Lcan - i : nullptr);
#ifdef MMPC_T4_PROFILE
    __syncthreads();
#ifdef MMPC_T4_PROFILE_CTAS
    if (tid == 0) printf("[t4 cta] %d total %lld cycles, %d iterations\n", (int)blockIdx.x, clock64() - t_entry, it);
#endif
    if (blockIdx.x == 0 && tid == 0) {
        printf("[t4 forward, all sweeps] call %lld  first chunk %lld  loop %lld  last post %lld  reduce %lld  return %lld\n", t4_dbg[0], t4_dbg[1], t4_dbg[2], t4_dbg[3], t4_dbg[4], t4_dbg[5]);
        printf("[t4 forward loop, all sweeps] worker: issue %lld  post (incl. wait for its bulk copies) %lld  barrier %lld | serial: wait for bulk copies %lld  chain %lld  barrier %lld\n", t4_dbg[6], t4_dbg[7], t4_dbg[9], t4_dbg[12], t4_dbg[10], t4_dbg[11]);
        for (int q = 0; q < 16; ++q) t4_dbg[q] = 0;
    }
    if (blockIdx.x == 0 && tid == 0)
        printf("[t4 kernel] prologue %lld  loop %lld (residual sweeps %lld, backward %lld, forward %lld, votes %lld)  epilogue %lld cycles, %d iterations\n",
               tk0 - t_entry, tk1 - tk0, tp_res, tp_bwd, tp_fwd, tp_post, clock64() - tk1, it);
#endif
}

template <class M>
struct LaunchT4 {
    template <int I_, int CB_, int CF_, int RK_ = 0, int SK_ = 0, bool HO_ = true>
    static cudaError_t go(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        using C = T4<M, I_, CB_, CF_, RK_, SK_>;
        static_assert(C::SMEM_BYTES <= 227 * 1024, "record buffers must fit the 227 KB of shared memory a CTA can opt into");
        if (io.host_out && !HO_) return cudaErrorInvalidValue;  // the launcher's part selection pairs host_out with an HO_ instantiation
        static std::atomic<unsigned long long> attr_done{0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (C::SMEM_BYTES > 48 * 1024 && !((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ULL)) {
            cudaError_t e = cudaFuncSetAttribute(solve_kernel_t4<M, I_, CB_, CF_, RK_, SK_, HO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            attr_done.fetch_or(1ULL << (dev & 63), std::memory_order_release);
        }
        if (io.b0 % C::I != 0) return cudaErrorInvalidValue;  // sub-batches must start on a tile boundary
        const int t0 = io.b0 / C::I, t1 = (io.b1 + C::I - 1) / C::I;
        solve_kernel_t4<M, I_, CB_, CF_, RK_, SK_, HO_><<<t1 - t0, C::THREADS, C::SMEM_BYTES, st>>>(P, W, io);
        return cudaGetLastError();
    }
    // Tile width from the batch of the handle (mmpc_t4_tile_width; it fixes the layout of the persistent iterate):
    //   4 x 8 x 8    64 threads (warp 0 + one worker warp: 8 node slots per instance), B <= 8 * 148: the latency configuration --
    //               the backward sweep is bound by the recursion, not by evaluation rounds (N = 50: 7 rounds instead of 13)
    //  32 x 7 x 5   256 threads, 1 CTA / SM    beyond: full serial warps, 256-byte rows (whole L2 lines: see the discards)
    //   8 x 4 x 4 and 28 x 8 x 6   on request only (tests of other tile shapes; 28: 147 CTAs at B = 4096, but the serial warp then
    //               shares its SM sub-partition with a full worker warp: 0.377 vs 0.352 ms)
    // kernel_variant = 4 + 256*k forces configuration k (1: 8, 2: 28, 3: 32, 4: 4).  Measured at N = 50 (profiles/r2_t4_configs.txt):
    //   B = 4 / 32 / 296 / 592 / 1184 / 2368:  4 x 8 x 8: 0.190 / 0.212 / 0.225 / 0.224 / 0.232 / 0.342 ms,  8 x 4 x 4: 0.218 / 0.250 /
    //   0.251 / 0.266 / 0.268 / 0.302 ms.
    // The instantiations are compiled as four parts, one translation unit each (MMPC_DEFINE_SOLVE_T4 / _T4_PART in mmpc_kernels.cuh:
    // ModelGenerator::compile_model and the library build compile the units side by side):
    //   part 0  Euler / linearised, tile widths 4 and 32 (the defaults)     part 1  Euler / linearised, widths 8 and 28
    //   part 2  RK4 shooting, widths 4 and 32                               part 3  RK4 shooting, widths 8 and 28
    //   part 4  = part 0 with the zero-copy output epilogue (io.host_out); part 0 itself is instantiated without it
    static int part_of(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io) {
        const int I = mmpc_t4_tile_width(W.B, io.variant >> 8);
        const int part = ((P.integrator == MMPC_RK4 && !P.is_linear) ? 2 : 0) + ((I == 4 || I == 32) ? 0 : 1);
        return part == 0 && io.host_out ? 4 : part;
    }
    template <int PART>
    static cudaError_t run_part(const mmpc_problem& P, const mmpc_workspace& W, const mmpc_solve_io& io, cudaStream_t st) {
        if (P.bounded) return cudaErrorNotSupported;
        if constexpr (M::NX + 2 * M::NU <= 8) {
            const int I = mmpc_t4_tile_width(W.B, io.variant >> 8);
            if constexpr (PART == 0) return I == 4 ? go<4, 8, 8, 0, 0, false>(P, W, io, st) : go<32, 7, 5, 0, 0, false>(P, W, io, st);
            if constexpr (PART == 4) return I == 4 ? go<4, 8, 8>(P, W, io, st) : go<32, 7, 5>(P, W, io, st);
            if constexpr (PART == 1) return I == 8 ? go<8, 4, 4>(P, W, io, st) : go<28, 8, 6>(P, W, io, st);
            if constexpr (PART == 2) return I == 4 ? go<4, 8, 5, 1>(P, W, io, st) : go<32, 7, 5, 1>(P, W, io, st);
            if constexpr (PART == 3) return I == 8 ? go<8, 4, 4, 1>(P, W, io, st) : go<28, 8, 5, 1>(P, W, io, st);
        }
        return cudaErrorNotSupported;
    }
};

}  // namespace mmpc

// hc_sorted.cuh -- the phase-sorted HeatCool kernel (included by nyx_hc.cu).
//
// Same per-lane state machine as hc_integrate_kernel (hc_device.cuh), different execution plan:
//
//   * the integrator state of every lane lives in SHARED MEMORY (structure of arrays, [field][lane]); the rate tables move to
//     global memory (L1/L2) -- the register row cache of iterate_ne makes their latency a minor cost;
//   * each round has three CTA-wide phases:
//       R  every thread evaluates the pending RHS/EOS request of "its" lane (thread t <-> lane t): full warps, small live state;
//       S  the lanes are counting-sorted by integrator phase (the `pc` they wait in, split by "will request a Jacobian setup");
//       B  thread t runs the BDF bookkeeping of lane order[t]: it loads that lane's scalars into registers, resumes the state
//          machine and writes them back.  Lanes of one phase sit in the same warps, so the bookkeeping -- in the register-resident
//          kernel executed by 3..10 lanes per instruction -- runs in (nearly) full warps; finished lanes are stored, idle lanes
//          (sorted to the end: whole warps of them) are refilled together.
//
// This is the CTA-wide version of "warp-level ballot/compaction": compaction by integrator phase across the 384 cells in
// flight of an SM, every round.
#ifndef NYXB200_HC_SORTED_CUH
#define NYXB200_HC_SORTED_CUH

namespace sorted {

// ---- scalar fields of a lane kept in shared memory between phases (the Nordsieck arrays occupy slots 0..ARR_DOUBLES-1)
#define HC_SD_COMMON(X) X(req_t) X(req_y) X(rho) X(e0) X(lastT) X(lastNe) X(ewt) X(acor) X(ftemp) X(tn) X(h) X(hprime) X(eta) X(hscale) \
    X(etamax) X(rl1) X(gamma) X(gammap) X(crate) X(delp) X(saved_tq5) X(M) X(gammasv) X(saved_t) X(delta) X(yy_ft) X(hg) X(hub) X(hlb) X(e_final)
// (outT, outNe, IR -- the SDC finalize results that wait for the final EOS solve when reionization heating is on -- have no slots of their
//  own: a lane in PC_FINAL_EOS keeps them in the slots of the cvHin locals hg, hub, hlb, which are dead after the initial step.  592 instead
//  of 616 bytes per lane: 384 instead of 352 lanes fit.)
#if !defined(HC_STRUCT_ALIAS_OUT)
#define HC_STRUCT_ALIAS_OUT 1
#endif
#if HC_STRUCT_ALIAS_OUT
#define HC_SD_STRUCT(X) X(jh) X(rho_src) X(rhoe_src) X(e_src) X(rho_out) X(rhoe_new) X(reset_src) X(zhi) X(lastNh) X(lastRho) X(eos_nhe0) \
    X(eos_nhepp)
#else
#define HC_SD_STRUCT(X) X(jh) X(rho_src) X(rhoe_src) X(e_src) X(rho_out) X(rhoe_new) X(reset_src) X(zhi) X(lastNh) X(lastRho) X(eos_nhe0) \
    X(eos_nhepp) X(outT) X(outNe) X(IR)
#endif
#define HC_COUNT(name) +1
constexpr int ND_COMMON = 0 HC_SD_COMMON(HC_COUNT);
constexpr int ND_STRUCT = 0 HC_SD_STRUCT(HC_COUNT);
template <int PATH> constexpr int nd_total() { return ARR_DOUBLES + ND_COMMON + (PATH == PATH_STRUCT ? ND_STRUCT : 0) + 1 /* fval */; }
// 32-bit words per lane: two packed words of small integers, 12 counters, 4 words of cell coordinates
#define HC_SI_WORDS(X) X(nst) X(nstlp) X(nfe) X(nfe_ls) X(netf) X(nni) X(nnf) X(nsetups) X(ne_iters) X(attempts) X(n_eos) X(flag)
constexpr int NI_WORDS = 2 + (0 HC_SI_WORDS(HC_COUNT)) + 4;
#if !defined(HC_SORT_FINE)
#define HC_SORT_FINE 2
#endif
#if HC_SORT_FINE
// the step-completing lanes are split further by (order q, qwait), which decide loop trip counts and branches of their chain
// (measured in one gpurun call, best of 6, twice: 68.43 / 68.48 ms against 69.48 / 71.87 ms with the 8 plain keys; HC_SORT_FINE == 2
// interleaves the two step-completing phases within each (q, qwait) class: 62.15 / 61.60 ms against 64.66 / 65.86 ms for == 1)
constexpr int NSUB = 9;   // (min(q,3)-1) * 3 + (min(qwait,3)-1)
enum Key { K_NEWTON = 0, K_LSETUP = NSUB, K_SETUP_REQ = 2 * NSUB, K_HIN, K_INIT, K_ETEST, K_FINAL, K_IDLE, NKEY };
#else
constexpr int NSUB = 1;
constexpr int NKEY = 8;   // sort keys, in the order the warps will process them
enum Key { K_NEWTON = 0, K_SETUP_REQ, K_LSETUP, K_HIN, K_INIT, K_ETEST, K_FINAL, K_IDLE };
#endif
static_assert(NKEY <= 32, "one lane per key in the base computation");
// the 8 integrator phases behind the keys (diagnostics): NEWTON, SETUP_REQ, LSETUP, HIN, INIT, ETEST, FINAL, IDLE
__device__ __forceinline__ int key_class(int key) {
#if HC_SORT_FINE == 2
    return (key < K_SETUP_REQ) ? ((key & 1) ? 2 : 0) : (key == K_SETUP_REQ) ? 1 : 3 + (key - K_HIN);
#elif HC_SORT_FINE
    return (key < K_LSETUP) ? 0 : (key < K_SETUP_REQ) ? 2 : (key == K_SETUP_REQ) ? 1 : 3 + (key - K_HIN);
#else
    return key;
#endif
}

template <int PATH, int LANES>
struct Layout {
    static constexpr int ND = nd_total<PATH>();
    static constexpr int WARPS = LANES / 32;
    static constexpr size_t bytes_d = (size_t)ND * LANES * sizeof(double);
    static constexpr size_t bytes_i = (size_t)NI_WORDS * LANES * sizeof(int);
    static constexpr size_t bytes_order = (size_t)LANES * sizeof(unsigned short);
    static constexpr size_t bytes_cnt = (size_t)NKEY * WARPS * sizeof(int);
    static constexpr size_t total = bytes_d + bytes_i + bytes_order + bytes_cnt;
    static_assert((total + 2048) * HC_SORTED_CTAS <= 228 * 1024, "shared memory budget of one sm_100 SM");
};

template <int STRIDE>
struct ArrSmemT {
    double* p;   // this lane's slot 0
    __device__ __forceinline__ double& at(int slot) const { return p[slot * STRIDE]; }
};

template <class LaneT>
__device__ __forceinline__ int sort_key(unsigned w0, unsigned w1) {
    const int pc = (int)(w0 & 15u);
    const bool callSetup = (w1 >> 8) & 1u, res_at_top = (w1 >> 9) & 1u;
#if HC_SORT_FINE
    const int q = (int)((w0 >> 4) & 15u), qwait = (int)((w0 >> 12) & 15u);
    const int sub = (min(max(q, 1), 3) - 1) * 3 + (min(max(qwait, 1), 3) - 1);
#else
    const int sub = 0;
#endif
    switch (pc) {
#if HC_SORT_FINE == 2
    // the two step-completing phases interleaved: key = 2 * (q, qwait class) + phase, so that a Newton-residual lane sits next to the
    // Jacobian-setup lanes of the same order -- they differ in their first two stages only and share the long tail of the chain
    case PC_NLS_RES: return (res_at_top && callSetup) ? K_SETUP_REQ : K_NEWTON + 2 * sub;
    case PC_LSETUP_F: return K_NEWTON + 2 * sub + 1;
#else
    case PC_NLS_RES: return (res_at_top && callSetup) ? K_SETUP_REQ : K_NEWTON + sub;
    case PC_LSETUP_F: return K_LSETUP + sub;
#endif
    case PC_HIN_F: return K_HIN;
    case PC_INIT_F0: return K_INIT;
    case PC_ETEST_F: return K_ETEST;
    case PC_FINAL_EOS: return K_FINAL;
    default: return K_IDLE;
    }
}

template <class LaneT>
__device__ __forceinline__ void pack_ints(const LaneT& ln, unsigned& w0, unsigned& w1) {
    w0 = (unsigned)ln.pc | ((unsigned)ln.q << 4) | ((unsigned)ln.qprime << 8) | ((unsigned)ln.qwait << 12) | ((unsigned)ln.L << 16) |
         ((unsigned)ln.ncf << 20) | ((unsigned)ln.nef << 24) | ((unsigned)ln.curiter << 28);
    w1 = (unsigned)ln.nflag | ((unsigned)ln.hin_count << 4) | ((unsigned)ln.callSetup << 8) | ((unsigned)ln.res_at_top << 9) |
         ((unsigned)ln.jcur << 10) | ((unsigned)ln.nls_jcur << 11) | ((unsigned)ln.floor_hit << 12);
}
template <class LaneT>
__device__ __forceinline__ void unpack_ints(LaneT& ln, unsigned w0, unsigned w1) {
    ln.pc = (int)(w0 & 15u); ln.q = (int)((w0 >> 4) & 15u); ln.qprime = (int)((w0 >> 8) & 15u); ln.qwait = (int)((w0 >> 12) & 15u);
    ln.L = (int)((w0 >> 16) & 15u); ln.ncf = (int)((w0 >> 20) & 15u); ln.nef = (int)((w0 >> 24) & 15u); ln.curiter = (int)((w0 >> 28) & 15u);
    ln.nflag = (int)(w1 & 15u); ln.hin_count = (int)((w1 >> 4) & 15u); ln.callSetup = (w1 >> 8) & 1u; ln.res_at_top = (w1 >> 9) & 1u;
    ln.jcur = (w1 >> 10) & 1u; ln.nls_jcur = (w1 >> 11) & 1u; ln.floor_hit = (int)((w1 >> 12) & 1u);
}

#if defined(HC_PHASE_TIMING)
// diagnostics build only: per-phase clock64 totals summed over warps: [0] rounds*warps [1] sort [2] B work [3] B barrier wait [4] R work [5] R barrier wait
// [6] active lanes at R (sum over rounds) [7] kernel cycles*warps [8..15] lanes per sort key (sum over rounds)
// B work split: [16] load lane [17] resume + store_cell [18] refill [19] write back; [24..31] B work cycles of the warps whose first lane has
// sort key 0..7, [32..39] number of such warp-rounds, [40..47] their active lanes in resume()
__device__ unsigned long long g_phase[48];
#define HC_TICK(slot) do { const long long t_ = clock64(); ph[slot] += (unsigned long long)(t_ - t_last); t_last = t_; } while (0)
#else
#define HC_TICK(slot) do { } while (0)
#endif

template <int PATH, int LANES>
__global__ void __launch_bounds__(LANES, HC_SORTED_CTAS) hc_sorted_kernel(const __grid_constant__ KernelArgs a) {
    using L = Layout<PATH, LANES>;
    using LaneT = Lane<PATH, ArrSmemT<LANES>>;
    __shared__ unsigned long long s_stats[S_COUNT];
    double* sd = reinterpret_cast<double*>(s_raw);
    int* si = reinterpret_cast<int*>(s_raw + L::bytes_d);
    unsigned short* s_order = reinterpret_cast<unsigned short*>(s_raw + L::bytes_d + L::bytes_i);
    int* s_cnt = reinterpret_cast<int*>(s_raw + L::bytes_d + L::bytes_i + L::bytes_order);   // [NKEY][WARPS] counts

    const int tid = threadIdx.x;
    const unsigned lane_id = tid & 31u, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane_id) - 1u;
    const Tables tb{a.ionx, a.iony, a.cool, a.logtab};   // all three in global memory (L1/L2)
    const Consts& c = a.k;

    // slot indices
    enum : int { SD0 = ARR_DOUBLES };
    constexpr int SD_FVAL = L::ND - 1;
    enum : int { SI_W0 = 0, SI_W1 = 1, SI_CNT0 = 2, SI_TILE = NI_WORDS - 4, SI_CI = NI_WORDS - 3, SI_CJ = NI_WORDS - 2, SI_CK = NI_WORDS - 1 };

    if (tid < S_COUNT) s_stats[tid] = 0ull;
    si[SI_W0 * LANES + tid] = PC_IDLE;
    si[SI_W1 * LANES + tid] = 0;
    sd[(SD0 + 1) * LANES + tid] = 200.0;   // req_y of an idle lane: never evaluated, but keep it benign
    Totals tot;
#pragma unroll
    for (int i = 0; i < 7; ++i) tot.w[i] = 0ull;
    tot.max_nst = 0u;
    int w_tile = 0, w_j = 0, w_k = 0, w_x = 0, w_xend = 0;   // the warp's current chunk (warp-uniform)
    bool queue_empty = false;
    __syncthreads();
#if defined(HC_PHASE_TIMING)
    unsigned long long ph[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) ph[i] = 0ull;
    long long t_last = clock64();
    const long long t_begin = t_last;
#endif

    for (;;) {
        // ================= phase S: stable counting sort of the lanes by integrator phase
        {
            const int key = sort_key<LaneT>((unsigned)si[SI_W0 * LANES + tid], (unsigned)si[SI_W1 * LANES + tid]);
            const unsigned same = __match_any_sync(0xffffffffu, key);
            const int rank = __popc(same & lt_mask);
            if (lane_id < NKEY) s_cnt[lane_id * L::WARPS + warp] = 0;
            __syncwarp();
            if (rank == 0) s_cnt[key * L::WARPS + warp] = __popc(same);
            HC_TICK(20);
            __syncthreads();
            HC_TICK(21);
            // every warp computes its own bases: lane kk sums the counts of key kk over the warps (and over the warps before this one)
            int tot_k = 0, pre_k = 0;
            if (lane_id < NKEY) {
#pragma unroll
                for (int w = 0; w < L::WARPS; ++w) {
                    const int cw = s_cnt[lane_id * L::WARPS + w];
                    tot_k += cw;
                    if (w < (int)warp) pre_k += cw;
                }
            }
            int incl = tot_k;
#pragma unroll
            for (int o = 1; o < NKEY; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane_id >= o) incl += t; }
            const int base_k = incl - tot_k + pre_k;   // first position of this warp's lanes with key == lane_id
            s_order[__shfl_sync(0xffffffffu, base_k, key) + rank] = (unsigned short)tid;
            HC_TICK(22);
            __syncthreads();
#if defined(HC_PHASE_TIMING)
            if (rank == 0) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) if (key_class(key) == kk) ph[8 + kk] += __popc(same);
            }
            ph[0] += 1;
#endif
            HC_TICK(1);
        }

        // ================= phase B: bookkeeping of lane order[tid]
        bool active_after;
        {
            const int my = s_order[tid];
            LaneT ln;
            ln.arr.p = sd + my;
            unsigned w0 = (unsigned)si[SI_W0 * LANES + my], w1 = (unsigned)si[SI_W1 * LANES + my];
            unpack_ints(ln, w0, w1);
            {
                int s = SD0;
#define HC_LD(name) ln.name = sd[(s++) * LANES + my];
                HC_SD_COMMON(HC_LD)
                if (PATH == PATH_STRUCT) { HC_SD_STRUCT(HC_LD) }
#undef HC_LD
#if HC_STRUCT_ALIAS_OUT
                if (PATH == PATH_STRUCT) { ln.outT = ln.hg; ln.outNe = ln.hub; ln.IR = ln.hlb; }   // meaningful in PC_FINAL_EOS only
#endif
                int w = SI_CNT0;
#define HC_LDI(name) ln.name = si[(w++) * LANES + my];
                HC_SI_WORDS(HC_LDI)
#undef HC_LDI
            }
            if (PATH != PATH_STRUCT) {
                ln.jh = 1.0; ln.rho_src = ln.rhoe_src = ln.e_src = ln.rho_out = ln.rhoe_new = ln.reset_src = ln.zhi = 0.0;
                ln.lastNh = 1.0; ln.lastRho = ln.rho; ln.eos_nhe0 = ln.eos_nhepp = 0.0; ln.outT = ln.outNe = ln.IR = 0.0;
            }
            ln.abstol = nv_scale(c.atol_factor, ln.e0);
            ln.y = 0.0; ln.gamrat = 0.0; ln.acnrm = 0.0;
            int c_tile = si[SI_TILE * LANES + my], c_i = si[SI_CI * LANES + my], c_j = si[SI_CJ * LANES + my], c_k = si[SI_CK * LANES + my];
            const double f = sd[SD_FVAL * LANES + my];

            HC_TICK(16);
#if defined(HC_PHASE_TIMING)
            const long long t_b0 = clock64();
            const int key0 = key_class(__shfl_sync(0xffffffffu, sort_key<LaneT>(w0, w1), 0));
#endif
            const bool act0 = ln.active();
            const unsigned rmask = __ballot_sync(0xffffffffu, act0);
#if defined(HC_PHASE_TIMING)
#if !defined(HC_DBG_CLASS)
#define HC_DBG_CLASS 2   // 0: Newton-residual lanes, 2: Jacobian-setup lanes
#endif
            ln.dbg_on = (key0 == HC_DBG_CLASS) && (key_class(__shfl_sync(0xffffffffu, sort_key<LaneT>(w0, w1), 31)) == HC_DBG_CLASS);   // stage timing: warps made of one phase only
            ln.dbg_last = clock64();
#endif
            if (act0) {
                ln.resume(c, f, rmask);
                if (!ln.active()) store_cell(ln, a, a.tiles[c_tile], c_i, c_j, c_k, tot);
            }
            __syncwarp();
            HC_TICK(17);
            // ---- refill: idle lanes take the next cells of the warp's chunk (new chunks from the global queue)
            if (!queue_empty) {
                bool need = !ln.active();
                unsigned m = __ballot_sync(0xffffffffu, need);
                while (m) {
                    if (w_x >= w_xend) {
                        unsigned long long chunk = 0;
                        if (lane_id == 0) chunk = atomicAdd(a.queue, 1ull);
                        chunk = __shfl_sync(0xffffffffu, chunk, 0);
                        if (chunk >= (unsigned long long)a.nchunks) { queue_empty = true; break; }
                        w_tile = find_tile_by_chunk(a.tiles, a.ntiles, (long long)chunk);
                        const TileDesc& t = a.tiles[w_tile];
                        const unsigned local = (unsigned)((long long)chunk - t.chunk_begin);
                        const unsigned row = local / (unsigned)t.cpr, piece = local - row * (unsigned)t.cpr;
                        const unsigned kk = row / (unsigned)t.ny;
                        w_k = t.lo[2] + (int)kk;
                        w_j = t.lo[1] + (int)(row - kk * (unsigned)t.ny);
                        w_x = t.lo[0] + (int)piece * t.chunk_len;
                        w_xend = min(w_x + t.chunk_len, t.lo[0] + t.nx);
                    }
                    const int avail = w_xend - w_x;
                    const int rk = __popc(m & lt_mask);
                    if (need && rk < avail) {
                        c_tile = w_tile; c_i = w_x + rk; c_j = w_j; c_k = w_k;
                        load_cell(ln, a, a.tiles[c_tile], c_i, c_j, c_k);
                        if (ln.active()) need = false;
                        else store_cell(ln, a, a.tiles[c_tile], c_i, c_j, c_k, tot);
                    }
                    w_x += min(avail, __popc(m));
                    m = __ballot_sync(0xffffffffu, need);
                }
            }
            HC_TICK(18);
            // ---- write the lane back
            pack_ints(ln, w0, w1);
            si[SI_W0 * LANES + my] = (int)w0; si[SI_W1 * LANES + my] = (int)w1;
            {
#if HC_STRUCT_ALIAS_OUT
                if (PATH == PATH_STRUCT && ln.pc == PC_FINAL_EOS) { ln.hg = ln.outT; ln.hub = ln.outNe; ln.hlb = ln.IR; }
#endif
                int s = SD0;
#define HC_ST(name) sd[(s++) * LANES + my] = ln.name;
                HC_SD_COMMON(HC_ST)
                if (PATH == PATH_STRUCT) { HC_SD_STRUCT(HC_ST) }
#undef HC_ST
                int w = SI_CNT0;
#define HC_STI(name) si[(w++) * LANES + my] = ln.name;
                HC_SI_WORDS(HC_STI)
#undef HC_STI
            }
            si[SI_TILE * LANES + my] = c_tile; si[SI_CI * LANES + my] = c_i; si[SI_CJ * LANES + my] = c_j; si[SI_CK * LANES + my] = c_k;
            active_after = ln.active();
#if defined(HC_PHASE_TIMING)
            {
                const long long t_b1 = clock64();
                const int nact = __popc(rmask);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) if (key0 == kk) { ph[24 + kk] += (unsigned long long)(t_b1 - t_b0); ph[32 + kk] += 1; ph[40 + kk] += nact; }
            }
#endif
        }
        HC_TICK(19);
        const bool any_active = __syncthreads_or(active_after);
        HC_TICK(3);
        if (!any_active) break;   // nothing in flight and the queue is empty (also publishes the lanes for phase R)

        // ================= phase R: thread t evaluates the request of lane t
        {
            LaneT ln;
            ln.arr.p = sd + tid;
            ln.pc = (int)((unsigned)si[SI_W0 * LANES + tid] & 15u);
            if (ln.active()) {
                ln.req_t = sd[(SD0 + 0) * LANES + tid];
                ln.req_y = sd[(SD0 + 1) * LANES + tid];
                ln.rho = sd[(SD0 + 2) * LANES + tid];
                ln.ne_iters = 0; ln.n_eos = 0;
                ln.jh = 1.0; ln.rho_src = 0.0; ln.e_src = 0.0; ln.lastRho = ln.rho;
                if (PATH == PATH_STRUCT) {
                    constexpr int S0 = SD0 + ND_COMMON;   // order of HC_SD_STRUCT: jh rho_src rhoe_src e_src ... lastNh(8) lastRho(9) eos_nhe0(10) eos_nhepp(11)
                    ln.jh = sd[(S0 + 0) * LANES + tid]; ln.rho_src = sd[(S0 + 1) * LANES + tid]; ln.e_src = sd[(S0 + 3) * LANES + tid];
                    ln.lastRho = sd[(S0 + 9) * LANES + tid];
                }
                const bool is_eos = (ln.pc == PC_FINAL_EOS);
                const double f = ln.eval_request(tb, c);
                sd[SD_FVAL * LANES + tid] = f;
                sd[(SD0 + 1) * LANES + tid] = ln.req_y;                 // the RHS clamps its argument in place (f_rhs.H:167)
                sd[(SD0 + 4) * LANES + tid] = ln.lastT;
                sd[(SD0 + 5) * LANES + tid] = ln.lastNe;
                if (PATH == PATH_STRUCT) {
                    constexpr int S0 = SD0 + ND_COMMON;
                    if (is_eos) { sd[(S0 + 10) * LANES + tid] = ln.eos_nhe0; sd[(S0 + 11) * LANES + tid] = ln.eos_nhepp; }
                    else { sd[(S0 + 8) * LANES + tid] = ln.lastNh; sd[(S0 + 9) * LANES + tid] = ln.lastRho; }
                }
                // counters: ne_iters is word 8, n_eos word 10 of HC_SI_WORDS
                si[(SI_CNT0 + 8) * LANES + tid] += ln.ne_iters;
                si[(SI_CNT0 + 10) * LANES + tid] += ln.n_eos;
            }
#if defined(HC_PHASE_TIMING)
            ph[6] += ln.active() ? 1 : 0;
#endif
        }
        HC_TICK(4);
        __syncthreads();
        HC_TICK(5);
    }
#if defined(HC_PHASE_TIMING)
    ph[7] = (unsigned long long)(clock64() - t_begin);
#pragma unroll
    for (int i = 0; i < 48; ++i) {
        unsigned long long v = ph[i];
        if (i == 6 || (i >= 8 && i < 16)) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); }   // per-thread quantities
        if (lane_id == 0) atomicAdd(&g_phase[i], v);
    }
#endif

    flush_totals(tot, s_stats, a.dstats);
}

}  // namespace sorted
#endif

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

// The lane state between phases is what Lane::save() writes (hc_device.cuh, "persistence between rounds"): LaneT::ND doubles
// (the 22 Nordsieck / coefficient doubles first) and LaneT::WS_N 32-bit words per lane, structure of arrays [slot][lane].
// 364 bytes per lane on the Strang path, 420 on the SDC path: 384 lanes fit the 164 KB shared-memory carve-out, which leaves the L1 92 KB
// for the rate tables and the stack (round 1: 496 / 592 bytes, 196 / 228 KB carve-out, 60 / 28 KB of L1).
// Sort keys.  The step-completing lanes (Newton-residual and Jacobian-setup phases) are split further by (order q, qwait), which decide loop
// trip counts and branches of their chain, and the two phases are interleaved within each (q, qwait) class: a Newton-residual lane sits
// next to the Jacobian-setup lanes of the same class -- they differ in their first two stages only and share the long tail of the chain.
// Measured, each within one gpurun call: 8 plain keys 69.5 ms -> (q, qwait) classes 68.4 ms -> interleaved 61.6 ms
// (256^3 Strang, round 1); no gain from splitting further by "this attempt reaches tout" (profiles/r2_s10_sortkey_laststep.log) or by
// first / later Newton iteration (profiles/r2_s20_sortkey_variants.log); (qwait, q) instead of (q, qwait) order: 0.99x / 0.965x
// (profiles/r2_s22_subkey_transposed.log).
constexpr int NSUB = 9;   // (min(q,3)-1) * 3 + (min(qwait,3)-1)
enum Key { K_NEWTON = 0, K_SETUP_REQ = 2 * NSUB, K_HIN, K_INIT, K_ETEST, K_FINAL, K_IDLE, NKEY };
static_assert(NKEY <= 32, "one lane per key in the base computation");
// the 8 integrator phases behind the keys (diagnostics): NEWTON, SETUP_REQ, LSETUP, HIN, INIT, ETEST, FINAL, IDLE
__device__ __forceinline__ int key_class(int key) {
    return (key < K_SETUP_REQ) ? ((key & 1) ? 2 : 0) : (key == K_SETUP_REQ) ? 1 : 3 + (key - K_HIN);
}

// Both accessors index the extern __shared__ array ITSELF with an integer lane number: through pointers kept in a struct the compiler
// lost the address space and emitted generic loads / stores for the lane state (long-scoreboard stalls all over the bookkeeping phase,
// +25 % kernel time: profiles/r2a_*).
template <int STRIDE>
struct ArrSmemT {
    int my;   // this lane's index
    __device__ __forceinline__ double& at(int slot) const { return reinterpret_cast<double*>(s_raw)[slot * STRIDE + my]; }
};
// slot access of one lane for Lane::save / load / load_request / save_result; BYTES_D = size of the double slots (the words follow them)
template <int STRIDE, int ND_SLOTS>
struct SmemIO {
    int my;
    __device__ __forceinline__ double& d(int slot) const { return reinterpret_cast<double*>(s_raw)[slot * STRIDE + my]; }
    __device__ __forceinline__ unsigned& w(int slot) const {
        return reinterpret_cast<unsigned*>(s_raw + (size_t)ND_SLOTS * STRIDE * sizeof(double))[slot * STRIDE + my];
    }
};

template <int PATH, int LANES>
struct Layout {
    using LaneT = Lane<PATH, ArrSmemT<LANES>>;
    static constexpr int ND = LaneT::ND;
    static constexpr int NI_WORDS = LaneT::WS_N;
    static constexpr int WARPS = LANES / 32;
    static constexpr size_t bytes_d = (size_t)ND * LANES * sizeof(double);
    static constexpr size_t bytes_i = (size_t)NI_WORDS * LANES * sizeof(int);
    static constexpr size_t bytes_order = (size_t)LANES * sizeof(unsigned short);
    static constexpr size_t bytes_cnt = (size_t)NKEY * WARPS * sizeof(int);
    static constexpr size_t total = bytes_d + bytes_i + bytes_order + bytes_cnt;
    static_assert((total + 2048) * HC_SORTED_CTAS <= 228 * 1024, "shared memory budget of one sm_100 SM");
};

template <class LaneT>
__device__ __forceinline__ int sort_key(unsigned w0, unsigned w1) {
    const int pc = (int)(w0 & 15u);
    const bool callSetup = (w1 >> 8) & 1u, res_at_top = (w1 >> 9) & 1u;
    const int q = (int)((w0 >> 4) & 15u), qwait = (int)((w0 >> 12) & 15u);
    const int sub = (min(max(q, 1), 3) - 1) * 3 + (min(max(qwait, 1), 3) - 1);
    switch (pc) {
    case PC_NLS_RES: return (res_at_top && callSetup) ? K_SETUP_REQ : K_NEWTON + 2 * sub;
    case PC_LSETUP_F: return K_NEWTON + 2 * sub + 1;
    case PC_HIN_F: return K_HIN;
    case PC_INIT_F0: return K_INIT;
    case PC_ETEST_F: return K_ETEST;
    case PC_FINAL_EOS: return K_FINAL;
    default: return K_IDLE;
    }
}

#if defined(HC_PHASE_TIMING)
// diagnostics build only: per-phase clock64 totals summed over warps: [0] rounds*warps [1] sort [2] B work [3] B barrier wait [4] R work [5] R barrier wait
// [6] active lanes at R (sum over rounds) [7] kernel cycles*warps [8..15] lanes per sort key (sum over rounds)
// B work split: [16] load lane [17] resume + store_cell [18] refill [19] write back; [24..31] B work cycles of the warps whose first lane has
// sort key 0..7, [32..39] number of such warp-rounds, [40..47] their active lanes in resume()
__device__ unsigned long long g_phase[48];
// bookkeeping phase by the classes of a warp's first and last lane (idx = 8 * first + last): [idx] cycles, [64 + idx] warp-rounds,
// [128 + idx] rounds in which such a warp was the slowest of its CTA, [192 + idx] its cycles then
__device__ unsigned long long g_mix[256];
#define HC_TICK(slot) do { const long long t_ = clock64(); ph[slot] += (unsigned long long)(t_ - t_last); t_last = t_; } while (0)
#else
#define HC_TICK(slot) do { } while (0)
#endif

// REACT (SDC path only): the instantiation behind the SAVE_REACT dumps (integrate_state_with_source_3d.cpp:126-183,602-631) also records,
// per cell, what CVODE returned before the finalize step touched it (raw solution, estimated local error) and what the finalize step
// left (density of the last RHS evaluation, final energy): a.react_raw, 4 doubles per cell.  The production instantiations are unchanged.
template <int PATH, int LANES, bool REACT = false>
__global__ void __launch_bounds__(LANES, HC_SORTED_CTAS) hc_sorted_kernel(const __grid_constant__ KernelArgs a) {
    using L = Layout<PATH, LANES>;
    using LaneT = typename L::LaneT;
    using IO = SmemIO<LANES, L::ND>;
    __shared__ unsigned long long s_pair[P_COUNT];        // packed diagnostics of the CTA
    __shared__ int s_chunk[L::WARPS][6];                  // per warp: the current work-queue chunk (tile, j, k, x, x_end) and "queue empty"
    double* sd = reinterpret_cast<double*>(s_raw);
    unsigned* si = reinterpret_cast<unsigned*>(s_raw + L::bytes_d);
    unsigned short* s_order = reinterpret_cast<unsigned short*>(s_raw + L::bytes_d + L::bytes_i);
    int* s_cnt = reinterpret_cast<int*>(s_raw + L::bytes_d + L::bytes_i + L::bytes_order);   // [NKEY][WARPS] counts

    const int tid = threadIdx.x;
    const unsigned lane_id = tid & 31u, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane_id) - 1u;
    const Tables tb{a.ionx, a.iony, a.cool, a.logtab};   // all three in global memory (L1/L2)
    const Consts& c = a.k;

    if (a.timing && tid == 0) atomicMax(&a.timing[2], ~global_ns());
    if (tid < P_COUNT) s_pair[tid] = 0ull;
    if (lane_id < 6) s_chunk[warp][lane_id] = 0;
    si[LaneT::WS_W0 * LANES + tid] = PC_IDLE;
    si[LaneT::WS_W1 * LANES + tid] = 0u;
    Totals tot;
    tot.iters_attempts = 0ull; tot.n_eos = 0u; tot.s_pair = s_pair;
    __syncthreads();
#if defined(HC_PHASE_TIMING)
    __shared__ unsigned long long s_slow;
    if (tid == 0) s_slow = 0ull;
    unsigned long long ph[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) ph[i] = 0ull;
    long long t_last = clock64();
    const long long t_begin = t_last;
#endif

    // Round structure (two CTA barriers per round):
    //   B   thread t runs the bookkeeping of lane `my` (stores a finished cell, refills an idle lane);
    //   S   counting sort by the keys the threads hold in registers: counts, barrier (which also decides termination), bases + order,
    //       barrier; thread t adopts lane order[t];
    //   R   thread t evaluates the request of ITS lane -- the lane whose bookkeeping it runs next, so no barrier separates R from B.
    // (Sorting only every 2nd / 3rd round lets the warps drift apart, evaluation code then runs next to bookkeeping code: 0.81x / 0.72x,
    // profiles/r2_s6_round_structure.log.)
    int my = tid;
    for (;;) {
        // ================= phase B: bookkeeping of lane `my`
        bool active_after;
        int key;
        {
            const IO io{my};
            LaneT ln;
            ln.arr.my = my;
            double f = 0.0;
            // the lane's cell, packed: tile (20 bits) | k (12 bits), i (16) | j (16), relative to the tile.  Decoded only where the FABs
            // are touched (finalize data, store of a finished cell): once per cell, not once per round
            unsigned cell0 = 0u, cell1 = 0u;
            ln.load(io, c, f);   // unconditional, idle lanes included: a conditional load turns every lane field into a phi and the kernel spills
            cell0 = io.w(LaneT::WS_CELL0); cell1 = io.w(LaneT::WS_CELL1);
            HC_TICK(16);
#if defined(HC_PHASE_TIMING)
            const long long t_b0 = clock64();
            const int key0 = key_class(__shfl_sync(0xffffffffu, sort_key<LaneT>(io.w(LaneT::WS_W0), io.w(LaneT::WS_W1)), 0));
            const int key31 = key_class(__shfl_sync(0xffffffffu, sort_key<LaneT>(io.w(LaneT::WS_W0), io.w(LaneT::WS_W1)), 31));
#endif
            const bool act0 = ln.active();
            const unsigned rmask = __ballot_sync(0xffffffffu, act0);
#if defined(HC_PHASE_TIMING)
#if !defined(HC_DBG_CLASS)
#define HC_DBG_CLASS 2   // 0: Newton-residual lanes, 2: Jacobian-setup lanes
#endif
            ln.dbg_on = (key0 == HC_DBG_CLASS) && (key_class(__shfl_sync(0xffffffffu, sort_key<LaneT>(io.w(LaneT::WS_W0), io.w(LaneT::WS_W1)), 31)) == HC_DBG_CLASS);   // stage timing: warps made of one phase only
            ln.dbg_last = clock64();
#endif
            if (act0) {
                // (SDC path) the cell data only the finalize step reads is fetched again when a lane gets there: it is not part of the lane state
                if (PATH == PATH_STRUCT && ln.pc == PC_FINAL_EOS) load_finalize_cell(ln, a, cell0, cell1);
                ln.resume(c, f, rmask);
                if (ln.fin_pending) {
                    if (PATH == PATH_STRUCT) load_finalize_cell(ln, a, cell0, cell1);
                    if (REACT) store_react_cvode(a, cell0, cell1, ln.e_final, ln.acor);   // dptr[idx], CVodeGetEstLocalErrors (= cv_acor, cvode_io.c:1346-1360)
                    ln.begin_finalize(c);
                }
                tot.iters_attempts += (unsigned long long)(unsigned)ln.attempts << 32;
                if (!ln.active()) store_cell_packed<LaneT, REACT>(ln, a, cell0, cell1, tot);
            }
            __syncwarp();
            HC_TICK(17);
            // ---- refill: idle lanes take the next cells of the warp's chunk (new chunks from the global queue).  The chunk cursor is
            // warp-uniform state that is touched once per finished cell: it lives in shared memory, not in registers
            if (!s_chunk[warp][5]) {
                bool need = !ln.active();
                unsigned m = __ballot_sync(0xffffffffu, need);
                if (m) {
                    int w_tile = s_chunk[warp][0], w_j = s_chunk[warp][1], w_k = s_chunk[warp][2], w_x = s_chunk[warp][3], w_xend = s_chunk[warp][4];
                    bool queue_empty = false;
                    while (m) {
                        if (w_x >= w_xend) {
                            unsigned long long chunk = 0;
                            if (lane_id == 0) chunk = atomicAdd(a.queue, 1ull);
                            chunk = __shfl_sync(0xffffffffu, chunk, 0);
                            if (chunk >= (unsigned long long)a.nchunks) {
                                queue_empty = true;
                                if (a.timing && lane_id == 0) atomicMax(&a.timing[0], ~global_ns());   // the drain tail starts here
                                break;
                            }
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
                            const TileDesc& t = a.tiles[w_tile];
                            const int c_i = w_x + rk;
                            cell0 = ((unsigned)w_tile << 12) | (unsigned)(w_k - t.lo[2]);
                            cell1 = ((unsigned)(c_i - t.lo[0]) << 16) | (unsigned)(w_j - t.lo[1]);
                            load_cell(ln, a, t, c_i, w_j, w_k);
                            tot.iters_attempts += (unsigned long long)(unsigned)ln.attempts << 32;
                            if (ln.active()) { need = false; io.w(LaneT::WS_CELL0) = cell0; io.w(LaneT::WS_CELL1) = cell1; }
                            else {
                                if (REACT) store_react_cvode(a, cell0, cell1, ln.e0, 0.0);   // CVode refused the input: the solution vector still holds e(t0)
                                store_cell<LaneT, REACT>(ln, a, t, c_i, w_j, w_k, tot);
                            }
                        }
                        w_x += min(avail, __popc(m));
                        m = __ballot_sync(0xffffffffu, need);
                    }
                    __syncwarp();
                    if (lane_id == 0) {
                        s_chunk[warp][0] = w_tile; s_chunk[warp][1] = w_j; s_chunk[warp][2] = w_k; s_chunk[warp][3] = w_x; s_chunk[warp][4] = w_xend;
                        if (queue_empty) s_chunk[warp][5] = 1;
                    }
                    __syncwarp();
                }
            }
            HC_TICK(18);
            // ---- write the lane back
            active_after = ln.active();
            if (active_after) ln.save(io);
            else {
                io.w(LaneT::WS_W0) = PC_IDLE; io.w(LaneT::WS_W1) = 0u;
            }
            key = sort_key<LaneT>(io.w(LaneT::WS_W0), io.w(LaneT::WS_W1));
#if defined(HC_PHASE_TIMING)
            {
                const long long t_b1 = clock64();
                const int nact = __popc(rmask);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) if (key0 == kk) { ph[24 + kk] += (unsigned long long)(t_b1 - t_b0); ph[32 + kk] += 1; ph[40 + kk] += nact; }
                if (lane_id == 0) {
                    const int idx = key0 * 8 + key31;
                    atomicAdd(&g_mix[idx], (unsigned long long)(t_b1 - t_b0)); atomicAdd(&g_mix[64 + idx], 1ull);
                    atomicMax(&s_slow, ((unsigned long long)(t_b1 - t_b0) << 6) | (unsigned long long)idx);
                }
            }
#endif
        }
        HC_TICK(19);

        // ================= phase S: stable counting sort of the lanes by integrator phase
        {
            const unsigned same = __match_any_sync(0xffffffffu, key);
            const int rank = __popc(same & lt_mask);
            if (lane_id < NKEY) s_cnt[lane_id * L::WARPS + warp] = 0;
            __syncwarp();
            if (rank == 0) s_cnt[key * L::WARPS + warp] = __popc(same);
            HC_TICK(20);
            const bool any_active = __syncthreads_or(active_after);
#if defined(HC_PHASE_TIMING)
            if (tid == 0) {
                const unsigned long long v = s_slow;
                atomicAdd(&g_mix[128 + (int)(v & 63ull)], 1ull); atomicAdd(&g_mix[192 + (int)(v & 63ull)], v >> 6);
                s_slow = 0ull;
            }
#endif
            HC_TICK(3);
            if (!any_active) break;   // nothing in flight and the queue is empty
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
            s_order[__shfl_sync(0xffffffffu, base_k, key) + rank] = (unsigned short)my;
            HC_TICK(22);
            __syncthreads();
            my = s_order[tid];
#if defined(HC_PHASE_TIMING)
            if (rank == 0) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) if (key_class(key) == kk) ph[8 + kk] += __popc(same);
            }
#endif
            HC_TICK(1);
        }
#if defined(HC_PHASE_TIMING)
        ph[0] += 1;
#endif

        // ================= phase R: thread t evaluates the request of its lane
        {
            const IO io{my};
            LaneT ln;
            ln.arr.my = my;
            ln.pc = (int)(io.w(LaneT::WS_W0) & 15u);
            if (ln.active()) {
                ln.load_request(io);
                const bool is_eos = (ln.pc == PC_FINAL_EOS);
                const double f = ln.eval_request(tb, c);
                ln.save_result(io, f, is_eos);
                tot.iters_attempts += (unsigned long long)(unsigned)ln.ne_iters;
                tot.n_eos += (unsigned)ln.n_eos;
            }
#if defined(HC_PHASE_TIMING)
            ph[6] += ln.active() ? 1 : 0;
#endif
        }
        HC_TICK(4);
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

    flush_totals(tot, a.dstats);
    if (a.timing && tid == 0) atomicMax(&a.timing[1], global_ns());
}

}  // namespace sorted
#endif

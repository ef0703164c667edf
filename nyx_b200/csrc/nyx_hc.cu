// nyx_hc.cu -- sm_100a kernels and the C-ABI of include/nyx_hc.h.
//
// Kernel design (DESIGN.md has the numbers):
//   * the integrator kernel is sorted::hc_sorted_kernel (hc_sorted.cuh): one persistent CTA of 384 threads per SM; the whole integrator
//     state of the 384 cells in flight lives in SHARED MEMORY (364 / 420 bytes per lane, structure of arrays), and every round has three
//     CTA-wide phases -- R: thread t evaluates the pending right-hand side / EOS request of lane t (the FP64-heavy part, full warps);
//     S: the lanes are counting-sorted by integrator phase; B: thread t runs the CVODE bookkeeping of lane order[t] (hc_device.cuh:
//     a resumable per-cell BDF state machine), stores finished cells and refills idle lanes from a global work queue of x-row pieces;
//   * the rate tables stay in global memory behind the L1 (92 KB next to the 164 KB shared-memory carve-out) and L2; the rows of one
//     temperature bin are cached in registers across the evaluation points of an ionization-equilibrium solve; the UV-background row is
//     interpolated once per call on the host (z is uniform) and travels as kernel constants;
//   * per-cell outputs are scattered straight to the FABs; diagnostics are per-thread totals, reduced once at kernel end;
//   * hc_eos_kernel (compute_new_temp / EOS rows: one ionization-equilibrium solve per cell, tables staged in shared memory) and the
//     streaming kernels of the rows either side of the path (hc_reset_e_kernel, hc_sources.cuh) follow.
// No AMReX, no SUNDIALS, no library kernels, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

// Build-time tuning knobs (measured on B200, profiles/): 384 lanes per SM leave 168 registers per thread
#ifndef HC_SORTED_LANES_VEC
#define HC_SORTED_LANES_VEC 384            // lanes (cells in flight) per SM, one persistent CTA per SM; 364 B of shared memory per lane on the Strang path
#endif
#ifndef HC_SORTED_CTAS
#define HC_SORTED_CTAS 1                   // CTAs per SM (each with its own phase barriers; LANES x CTAS lanes in flight per SM)
#endif
#ifndef HC_SORTED_LANES_STRUCT
#define HC_SORTED_LANES_STRUCT 384         // 420 B per lane on the SDC path
#endif
#if defined(HC_PHASE_TIMING)
// diagnostics build: cycles between the stage boundaries of Lane::resume(), summed over the warps the kernel switches on
__device__ unsigned long long g_stage[16];
__host__ __device__ __forceinline__ void hc_stage_tick(long long& last, bool on, int slot) {
#if defined(__CUDA_ARCH__)
    if (on) { const long long t_ = clock64(); if ((threadIdx.x & 31u) == 0u) atomicAdd(&g_stage[slot], (unsigned long long)(t_ - last)); last = t_; }
#endif
}
#define HC_STAGE_TICK(ln, slot) hc_stage_tick((ln).dbg_last, (ln).dbg_on, slot)
#endif
#include "hc_host.hpp"

namespace {

using namespace hc;

#ifndef HC_EOS_THREADS
#define HC_EOS_THREADS 512                 // threads per CTA of hc_eos_kernel, one CTA per SM (tables: 112 KB of shared memory); measured 256: 6.9, 384: 5.5, 512: 4.9, 768: 4.9 ms per 3.4e7 cells
#endif
constexpr int EOS_THREADS = HC_EOS_THREADS;
constexpr int TAB_ROWS = NTAB + 1;         // one padding row: row j+1 always exists
constexpr int CHUNK_MAX = 256;             // cells per work-queue chunk (a piece of one x-row of a tile)
// dynamic shared memory of hc_eos_kernel: [ionx 2002 x 48 B][iony 2003 x 8 B, padded to 16]
constexpr size_t SM_IONX = (size_t)TAB_ROWS * IONX_ROW * sizeof(double);
constexpr size_t SM_IONY = (((size_t)TAB_ROWS + 1) * sizeof(double) + 15) / 16 * 16;
constexpr size_t SMEM_EOS = SM_IONX + SM_IONY;

// dynamic shared memory of the kernels
extern __shared__ __align__(16) unsigned char s_raw[];

enum Comp { DENS = 0, EDEN = 4, EINT = 5, TEMP = 0, NE = 1, ZHI = 2 };
enum FabSlot { F_STATE = 0, F_DIAG = 1, F_SNEW = 2, F_HSRC = 3, F_RSRC = 4, F_IR = 5 };

struct TileDesc {
    HcFab f[6];
    int lo[3];
    int nx, ny, nz;
    int cpr;                  // chunks per x-row
    int chunk_len;            // cells per chunk (last chunk of a row may be shorter)
    long long chunk_begin;    // global index of this tile's first chunk
    long long offset;         // global index of this tile's first cell (cell_stats ordering)
};

struct KernelArgs {
    Consts k;
    const TileDesc* tiles;
    int ntiles;
    long long ncells;
    long long nchunks;
    unsigned long long* queue;
    unsigned long long* dstats;   // HcStats as 14 x u64
    unsigned long long* timing;   // [0] ~(first time a warp found the work queue empty) [1] last CTA exit [2] ~(first CTA start), %globaltimer ns
                                  // (complemented values under atomicMax: the scratch is zero-initialised); nullptr: not recorded
    HcCellStat* cell_stats;
    double* react_raw;            // REACT instantiation only: per cell {CVODE's solution, its estimated local error, rho of the last RHS evaluation, final e}
    const double* ionx;
    const double* iony;
    const double* cool;
    const double* logtab;
    // hc_eos_kernel / hc_reset_e_kernel only
    int eos_mode;             // 0: T, ne from (rho, e = rho_e / rho);  1: the cell body of Nyx::compute_new_temp
    int max_temp_dt, interp;
    double small_temp, large_temp;
};

enum StatSlot { S_CELLS = 0, S_FAILED, S_FLOOR, S_NST, S_MAXNST, S_NFE, S_NFELS, S_NETF, S_NNI, S_NCFN, S_NSETUPS, S_NEITERS, S_ATTEMPTS, S_EOS, S_COUNT };
static_assert(S_COUNT * sizeof(long long) == sizeof(HcStats), "HcStats layout");

__device__ __forceinline__ long long fab_off(const HcFab& f, int i, int j, int k) {
    return (long long)(i - f.lo[0]) + (long long)(j - f.lo[1]) * f.jstride + (long long)(k - f.lo[2]) * f.kstride;
}

__device__ __forceinline__ int find_tile_by_chunk(const TileDesc* tiles, int ntiles, long long chunk) {
    int lo = 0, hi = ntiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tiles[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// the tile that holds global cell `id` (tiles are concatenated in `offset` order)
__device__ __forceinline__ int find_tile_by_cell(const TileDesc* tiles, int ntiles, long long id) {
    int lo = 0, hi = ntiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tiles[mid].offset <= id) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// grid-stride loops visit the cells in increasing order: walk forward from the tile of the previous cell (binary search the first time)
__device__ __forceinline__ int next_tile(const KernelArgs& a, int t, long long id) {
    if (t < 0) return find_tile_by_cell(a.tiles, a.ntiles, id);
    while (t + 1 < a.ntiles && a.tiles[t + 1].offset <= id) ++t;
    return t;
}
__device__ __forceinline__ void cell_of(const TileDesc& t, long long id, int& i, int& j, int& k) {
    const unsigned local = (unsigned)(id - t.offset);
    const unsigned row = local / (unsigned)t.nx, kk = row / (unsigned)t.ny;
    i = t.lo[0] + (int)(local - row * (unsigned)t.nx);
    j = t.lo[1] + (int)(row - kk * (unsigned)t.ny);
    k = t.lo[2] + (int)kk;
}

// stage the ionization tables into shared memory (16-byte vector copies)
__device__ __forceinline__ void stage_tables(const KernelArgs& a, double* s_ionx, double* s_iony) {
    const double2* src = reinterpret_cast<const double2*>(a.ionx);
    double2* dst = reinterpret_cast<double2*>(s_ionx);
#pragma unroll 2
    for (int i = threadIdx.x; i < TAB_ROWS * IONX_ROW / 2; i += blockDim.x) dst[i] = __ldg(src + i);
#pragma unroll 1
    for (int i = threadIdx.x; i < TAB_ROWS + 1; i += blockDim.x) s_iony[i] = __ldg(a.iony + i);
}

// Diagnostics.  Per CELL counters are added to the CTA's totals in shared memory when the cell is stored (packed two 32-bit counters per
// 64-bit word: 5 shared-memory atomics per cell); only the two per-ROUND quantities ride in registers.  (Fifteen registers of per-thread
// totals were live through both phases of the kernel, next to a lane state that already spills at 168 registers per thread.)
enum PairSlot { P_CELLS_FAILED = 0, P_FLOOR_NST, P_NFE_NFELS, P_NETF_NNI, P_NCFN_NSETUPS, P_NEITERS_ATTEMPTS, P_EOS, P_MAXNST, P_COUNT };
struct Totals {
    unsigned long long iters_attempts;   // sum of iterate_ne Newton iterations | step attempts << 32
    unsigned int n_eos;
    unsigned long long* s_pair;          // the CTA's packed totals (shared memory, P_COUNT words)
};

// gather one cell into a lane (HOT LOOP A of the reference: integrate_state_vec_3d.cpp:227-233,
// ode_eos_initialize_arrays f_rhs_struct.H:180-209) and start its integration
template <class LaneT>
__device__ __forceinline__ void load_cell(LaneT& ln, const KernelArgs& a, const TileDesc& t, int i, int j, int k) {
    constexpr int PATH = LaneT::path;
    const Consts& c = a.k;
    const long long so = fab_off(t.f[F_STATE], i, j, k);
    ln.rho = t.f[F_STATE].p[so + DENS * t.f[F_STATE].nstride];
    const double rhoe0 = t.f[F_STATE].p[so + EINT * t.f[F_STATE].nstride];
    ln.e0 = rhoe0 / ln.rho;
    ln.abstol = nv_scale(c.atol_factor, ln.e0);
    ln.jh = (double)c.JH0;
    if (PATH == PATH_STRUCT) {
        const long long dof = fab_off(t.f[F_DIAG], i, j, k);
        ln.lastT = t.f[F_DIAG].p[dof + TEMP * t.f[F_DIAG].nstride];
        ln.lastNe = t.f[F_DIAG].p[dof + NE * t.f[F_DIAG].nstride];
        ln.rho_src = ln.rhoe_src = ln.e_src = ln.reset_src = 0.0; ln.zhi = 0.0;
        if (c.sdc_has_src) {
            const long long ho = fab_off(t.f[F_HSRC], i, j, k);
            ln.rho_src = t.f[F_HSRC].p[ho + DENS * t.f[F_HSRC].nstride] / c.dt;
            ln.rhoe_src = t.f[F_HSRC].p[ho + EINT * t.f[F_HSRC].nstride] / c.dt;
            ln.reset_src = t.f[F_RSRC].p[fab_off(t.f[F_RSRC], i, j, k)];
            ln.e_src = (((c.asq * rhoe0 + c.dt * ln.rhoe_src) / c.aendsq + ln.reset_src) / (ln.rho + c.dt * ln.rho_src) - ln.e0) / c.dt;
        }
        if (c.inhomo) { ln.zhi = t.f[F_DIAG].p[dof + ZHI * t.f[F_DIAG].nstride]; ln.jh = (c.z > ln.zhi) ? 0.0 : 1.0; }
        const long long no = fab_off(t.f[F_SNEW], i, j, k);
        ln.rho_out = t.f[F_SNEW].p[no + DENS * t.f[F_SNEW].nstride];
        ln.rhoe_new = t.f[F_SNEW].p[no + EINT * t.f[F_SNEW].nstride];
    } else {
        ln.lastT = 0.0; ln.lastNe = 0.0;   // diag(Temp, Ne) are dead inputs on the Strang path (always overwritten, eos_hc.H:151)
    }
    ln.start(c);
}

// (SDC path) the cell data only the finalize step reads -- rhoe_src, reset_src of the source construction and S_new(rho, rho e) -- is not
// part of the lane state between rounds: a lane that reaches ode_eos_finalize_struct fetches it again (once per cell)
__device__ __forceinline__ void unpack_cell(const KernelArgs& a, unsigned cell0, unsigned cell1, int& tile, int& i, int& j, int& k) {
    tile = (int)(cell0 >> 12);
    const TileDesc& t = a.tiles[tile];
    k = t.lo[2] + (int)(cell0 & 0xfffu); i = t.lo[0] + (int)(cell1 >> 16); j = t.lo[1] + (int)(cell1 & 0xffffu);
}
template <class LaneT>
__device__ __forceinline__ void load_finalize_cell(LaneT& ln, const KernelArgs& a, unsigned cell0, unsigned cell1) {
    const Consts& c = a.k;
    int tile, i, j, k;
    unpack_cell(a, cell0, cell1, tile, i, j, k);
    const TileDesc& t = a.tiles[tile];
    ln.rhoe_src = 0.0; ln.reset_src = 0.0;
    if (c.sdc_has_src) {
        ln.rhoe_src = t.f[F_HSRC].p[fab_off(t.f[F_HSRC], i, j, k) + EINT * t.f[F_HSRC].nstride] / c.dt;
        ln.reset_src = t.f[F_RSRC].p[fab_off(t.f[F_RSRC], i, j, k)];
    }
    const long long no = fab_off(t.f[F_SNEW], i, j, k);
    ln.rho_out = t.f[F_SNEW].p[no + DENS * t.f[F_SNEW].nstride];
    ln.rhoe_new = t.f[F_SNEW].p[no + EINT * t.f[F_SNEW].nstride];
}

// scatter a finished cell (HOT LOOP C: integrate_state_vec_3d.cpp:317-321, f_rhs_struct.H:290-291,438-444)
__device__ __forceinline__ long long cell_index(const TileDesc& t, int i, int j, int k) {
    return t.offset + ((long long)(k - t.lo[2]) * t.ny + (j - t.lo[1])) * t.nx + (i - t.lo[0]);
}

template <class LaneT, bool REACT = false>
__device__ __forceinline__ void store_cell(const LaneT& ln, const KernelArgs& a, const TileDesc& t, int i, int j, int k, Totals& tot) {
    constexpr int PATH = LaneT::path;
    const Consts& c = a.k;
    const long long dof = fab_off(t.f[F_DIAG], i, j, k);
    t.f[F_DIAG].p[dof + TEMP * t.f[F_DIAG].nstride] = ln.outT;
    t.f[F_DIAG].p[dof + NE * t.f[F_DIAG].nstride] = ln.outNe;
    if (PATH == PATH_VEC || !c.sdc_has_src) {
        const double d = ln.rho * (ln.e_final - ln.e0);
        double* ps = t.f[F_STATE].p + fab_off(t.f[F_STATE], i, j, k);
        ps[EINT * t.f[F_STATE].nstride] += d;
        ps[EDEN * t.f[F_STATE].nstride] += d;
    } else {
        t.f[F_IR].p[fab_off(t.f[F_IR], i, j, k)] = ln.IR;
        const double d = c.dt * c.ahalf * ln.IR / c.aendsq;
        double* pn = t.f[F_SNEW].p + fab_off(t.f[F_SNEW], i, j, k);
        pn[EINT * t.f[F_SNEW].nstride] = ln.rhoe_new + d;
        pn[EDEN * t.f[F_SNEW].nstride] = pn[EDEN * t.f[F_SNEW].nstride] + d;
    }
    if (a.cell_stats) a.cell_stats[cell_index(t, i, j, k)] = HcCellStat{ln.nst, ln.netf, ln.nfe, ln.nni, ln.nnf, ln.nsetups, ln.nfe_ls, ln.flag};
    if (REACT) {
        double* raw = a.react_raw + 4 * cell_index(t, i, j, k);
        raw[2] = ln.lastRho; raw[3] = ln.e_final;
    }
    atomicAdd(&tot.s_pair[P_CELLS_FAILED], 1ull | ((unsigned long long)(ln.flag < 0) << 32));
    atomicAdd(&tot.s_pair[P_FLOOR_NST], (unsigned long long)(unsigned)ln.floor_hit | ((unsigned long long)(unsigned)ln.nst << 32));
    atomicAdd(&tot.s_pair[P_NFE_NFELS], (unsigned long long)(unsigned)ln.nfe | ((unsigned long long)(unsigned)ln.nfe_ls << 32));
    atomicAdd(&tot.s_pair[P_NETF_NNI], (unsigned long long)(unsigned)ln.netf | ((unsigned long long)(unsigned)ln.nni << 32));
    atomicAdd(&tot.s_pair[P_NCFN_NSETUPS], (unsigned long long)(unsigned)ln.nnf | ((unsigned long long)(unsigned)ln.nsetups << 32));
    // (ne_iters, attempts and n_eos go to the totals round by round: they are not part of the lane state)
    atomicMax(&tot.s_pair[P_MAXNST], (unsigned long long)(unsigned)ln.nst);
}

template <class LaneT, bool REACT = false>
__device__ __forceinline__ void store_cell_packed(const LaneT& ln, const KernelArgs& a, unsigned cell0, unsigned cell1, Totals& tot) {
    int tile, i, j, k;
    unpack_cell(a, cell0, cell1, tile, i, j, k);
    store_cell<LaneT, REACT>(ln, a, a.tiles[tile], i, j, k, tot);
}

// REACT instantiation: what CVode handed back for this cell, before ode_eos_finalize_struct
__device__ __forceinline__ void store_react_cvode(const KernelArgs& a, unsigned cell0, unsigned cell1, double e_cvode, double ele) {
    int tile, i, j, k;
    unpack_cell(a, cell0, cell1, tile, i, j, k);
    double* raw = a.react_raw + 4 * cell_index(a.tiles[tile], i, j, k);
    raw[0] = e_cvode; raw[1] = ele;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __noinline__ void flush_totals(const Totals& tot, unsigned long long* dstats) {
    // once per thread at the end of the kernel: the per-round counters join the CTA's packed totals, then one global atomic per counter per CTA
    // (per-CTA 32-bit halves: 134 M cells over 148 CTAs x ~35 Newton iterations per cell = 3e7)
    if (tot.iters_attempts) atomicAdd(&tot.s_pair[P_NEITERS_ATTEMPTS], tot.iters_attempts);
    if (tot.n_eos) atomicAdd(&tot.s_pair[P_EOS], (unsigned long long)tot.n_eos);
    __syncthreads();
    const int lo_slot[7] = {S_CELLS, S_FLOOR, S_NFE, S_NETF, S_NCFN, S_NEITERS, S_EOS};
    const int hi_slot[7] = {S_FAILED, S_NST, S_NFELS, S_NNI, S_NSETUPS, S_ATTEMPTS, -1};
    if (threadIdx.x < 7) {
        const unsigned long long v = tot.s_pair[threadIdx.x];
        const unsigned long long lo = v & 0xffffffffull, hi = v >> 32;
        if (lo) atomicAdd(&dstats[lo_slot[threadIdx.x]], lo);
        if (hi && hi_slot[threadIdx.x] >= 0) atomicAdd(&dstats[hi_slot[threadIdx.x]], hi);
    }
    if (threadIdx.x == 7) atomicMax(&dstats[S_MAXNST], tot.s_pair[P_MAXNST]);
}

// EOS kernel: one thread per cell, grid-stride over the cells of all tiles; ionization tables staged in shared memory.
//   eos_mode 0: diag(Temp, Ne) = nyx_eos_T_given_Re(rho, rho_e / rho)                       (eos_hc.H:204-220)
//   eos_mode 1: the cell body of Nyx::compute_new_temp (Source/Driver/Nyx.cpp:2473-2519): e = rho_e * (1 / rho); cells at or above
//               large_temp are clipped (max_temp_dt), cells with rho_e <= 0 are reset to small_temp; both rewrite (rho e, rho E).
// dstats: S_CELLS, S_EOS, S_NEITERS as for the integrators; S_FLOOR counts the cells reset to small_temp, S_FAILED the clipped ones.
__global__ void __launch_bounds__(EOS_THREADS, 1) hc_eos_kernel(const __grid_constant__ KernelArgs a) {
    __shared__ unsigned long long s_stats[S_COUNT];
    double* s_ionx = reinterpret_cast<double*>(s_raw);
    double* s_iony = reinterpret_cast<double*>(s_raw + SM_IONX);
    stage_tables(a, s_ionx, s_iony);
    if (threadIdx.x < S_COUNT) s_stats[threadIdx.x] = 0ull;
    __syncthreads();
    const Tables tb{s_ionx, s_iony, a.cool, a.logtab};
    const Consts& c = a.k;
    unsigned long long iters = 0, cells = 0, n_eos = 0, n_small = 0, n_large = 0;
    int ti = -1;
    for (long long id = (long long)blockIdx.x * EOS_THREADS + threadIdx.x; id < a.ncells; id += (long long)gridDim.x * EOS_THREADS) {
        ti = next_tile(a, ti, id);
        const TileDesc& t = a.tiles[ti];
        const HcFab& S = t.f[F_STATE];
        const HcFab& D = t.f[F_DIAG];
        int i, j, k;
        cell_of(t, id, i, j, k);
        const long long so = fab_off(S, i, j, k), dof = fab_off(D, i, j, k);
        const double R = S.p[so + DENS * S.nstride];
        const double rhoe = S.p[so + EINT * S.nstride];
        cells++;
        if (a.eos_mode == 0 || rhoe > 0.0) {
            const double e = (a.eos_mode == 0) ? rhoe / R : rhoe * (1.0 / R);
            const double rho_cgs = R * density_to_cgs / c.a3_eos;
            const double U = e * e_to_cgs;
            const double nh = rho_cgs * c.h_species / MPROTON;
            EosOut s;
            iterate_ne(tb, c, c.uvb_eos, 1.0, 1.0, U, nh, s);
            iters += s.iters; n_eos++;
            double T = s.T;
            if (a.eos_mode == 1 && T >= a.large_temp && a.max_temp_dt == 1) {   // Nyx.cpp:2488-2505
                T = a.large_temp;
                const double mu = c.c_mu_num / (c.c_mu_den + s.ne);             // nyx_eos_given_RT, eos_hc.H:222-231
                const double eint = T / (c.gm1 * mp_over_kb * mu);
                const double rhoInv = 1.0 / R;
                const double mx = S.p[so + 1 * S.nstride], my = S.p[so + 2 * S.nstride], mz = S.p[so + 3 * S.nstride];
                const double ke = 0.5e0 * (mx * mx + my * my + mz * mz) * rhoInv;
                const double re = R * eint;
                S.p[so + EINT * S.nstride] = re;
                S.p[so + EDEN * S.nstride] = re + ke;
                n_large++;
            }
            D.p[dof + TEMP * D.nstride] = T;
            D.p[dof + NE * D.nstride] = s.ne;
        } else {   // Nyx.cpp:2507-2519: rho e <= 0 -> small_temp with the cell's current ne
            const double ne_old = D.p[dof + NE * D.nstride];
            const double mu = c.c_mu_num / (c.c_mu_den + ne_old);
            const double eint = a.small_temp / (c.gm1 * mp_over_kb * mu);
            const double rhoInv = 1.0 / R;
            const double mx = S.p[so + 1 * S.nstride], my = S.p[so + 2 * S.nstride], mz = S.p[so + 3 * S.nstride];
            const double ke = 0.5e0 * (mx * mx + my * my + mz * mz) * rhoInv;
            const double re = R * eint;
            D.p[dof + TEMP * D.nstride] = a.small_temp;
            S.p[so + EINT * S.nstride] = re;
            S.p[so + EDEN * S.nstride] = re + ke;
            n_small++;
        }
    }
    atomicAdd(&s_stats[S_CELLS], cells);
    atomicAdd(&s_stats[S_EOS], n_eos);
    atomicAdd(&s_stats[S_NEITERS], iters);
    if (n_small) atomicAdd(&s_stats[S_FLOOR], n_small);
    if (n_large) atomicAdd(&s_stats[S_FAILED], n_large);
    __syncthreads();
    if (threadIdx.x < S_COUNT) atomicAdd(&a.dstats[threadIdx.x], s_stats[threadIdx.x]);
}

// SAVE_REACT dumps (ode_eos_save_react_arrays, f_rhs_struct.H:213-267, called right after ode_eos_finalize_struct): the three diagnostic
// FABs react_in (7 components), react_out (7), react_out_work (9), assembled per cell from the call's inputs (S_old is read-only when
// sdc_iter >= 0), the per-cell record of the REACT instantiation of the integrator kernel and its per-cell counters.
//   react_in        eptr (= e(t0)), rho_init_vode, rhoe_src_vode, e_src_vode, abstol, a, time_in (= 0)
//   react_out       dptr (CVODE's solution, untouched by the finalize step), rho_init + dt * rho_src, T_vode, ne_vode, estimated local error, a_end, dt
//   react_out_work  nst, netf, nfe, nni, ncfn, nsetups, nje, ncfl, nfeLS -- per CELL here, per tile-wide CVODE instance in the reference; nje and ncfl are
//                   never assigned by the reference's GetFinalStats (integrate_state_with_source_3d.cpp:792-810: uninitialised values): written as 0
// T_vode / ne_vode after the finalize step are the EOS solve of its LAST nyx_eos_T_given_Re_device call (f_rhs_struct.H:346-348, or :423-426 after
// instantaneous reionization heating): a function of (rho of the last RHS evaluation, final e, J_H of the cell) only -- the solve starts from
// ne = 1 whatever the caller holds (eos_hc.H:151) -- so it is evaluated here rather than carried through the integrator kernel.
struct ReactArgs {
    const HcFab* rf;          // [ntiles][3]: react_in, react_out, react_out_work of each tile
    const double* raw;        // [ncells][4]
    const HcCellStat* cs;     // [ncells]
};
__global__ void __launch_bounds__(EOS_THREADS, 1) hc_react_kernel(const __grid_constant__ KernelArgs a, const __grid_constant__ ReactArgs r) {
    double* s_ionx = reinterpret_cast<double*>(s_raw);
    double* s_iony = reinterpret_cast<double*>(s_raw + SM_IONX);
    stage_tables(a, s_ionx, s_iony);
    __syncthreads();
    const Tables tb{s_ionx, s_iony, a.cool, a.logtab};
    const Consts& c = a.k;
    int ti = -1;
    for (long long id = (long long)blockIdx.x * EOS_THREADS + threadIdx.x; id < a.ncells; id += (long long)gridDim.x * EOS_THREADS) {
        ti = next_tile(a, ti, id);
        const TileDesc& t = a.tiles[ti];
        int i, j, k;
        cell_of(t, id, i, j, k);
        // ode_eos_initialize_arrays f_rhs_struct.H:180-193, the expressions of load_cell
        const long long so = fab_off(t.f[F_STATE], i, j, k);
        const double rho = t.f[F_STATE].p[so + DENS * t.f[F_STATE].nstride];
        const double rhoe0 = t.f[F_STATE].p[so + EINT * t.f[F_STATE].nstride];
        const double e0 = rhoe0 / rho;
        const long long ho = fab_off(t.f[F_HSRC], i, j, k);
        const double rho_src = t.f[F_HSRC].p[ho + DENS * t.f[F_HSRC].nstride] / c.dt;
        const double rhoe_src = t.f[F_HSRC].p[ho + EINT * t.f[F_HSRC].nstride] / c.dt;
        const double reset_src = t.f[F_RSRC].p[fab_off(t.f[F_RSRC], i, j, k)];
        const double e_src = (((c.asq * rhoe0 + c.dt * rhoe_src) / c.aendsq + reset_src) / (rho + c.dt * rho_src) - e0) / c.dt;
        double jh = (double)c.JH0;
        if (c.inhomo) jh = (c.z > t.f[F_DIAG].p[fab_off(t.f[F_DIAG], i, j, k) + ZHI * t.f[F_DIAG].nstride]) ? 0.0 : 1.0;
        const double* raw = r.raw + 4 * id;
        // nyx_eos_T_given_Re_device(rho_vode, e_out) eos_hc.H:190-220, as Lane::eval_request evaluates it for PC_FINAL_EOS
        const double rho_cgs = raw[2] * density_to_cgs / c.a3_eos;
        const double U = raw[3] * e_to_cgs;
        const double nh = rho_cgs * c.h_species / MPROTON;
        EosOut s;
        iterate_ne(tb, c, c.uvb_eos, jh, (double)c.JHe0, U, nh, s);
        const HcFab& RI = r.rf[3 * ti + 0];
        const HcFab& RO = r.rf[3 * ti + 1];
        const HcFab& RW = r.rf[3 * ti + 2];
        double* pi = RI.p + fab_off(RI, i, j, k);
        pi[0 * RI.nstride] = e0; pi[1 * RI.nstride] = rho; pi[2 * RI.nstride] = rhoe_src; pi[3 * RI.nstride] = e_src;
        pi[4 * RI.nstride] = nv_scale(c.atol_factor, e0); pi[5 * RI.nstride] = c.a; pi[6 * RI.nstride] = 0.0;
        double* po = RO.p + fab_off(RO, i, j, k);
        po[0 * RO.nstride] = raw[0]; po[1 * RO.nstride] = rho + c.dt * rho_src; po[2 * RO.nstride] = s.T; po[3 * RO.nstride] = s.ne;
        po[4 * RO.nstride] = raw[1]; po[5 * RO.nstride] = c.a_end; po[6 * RO.nstride] = c.dt;
        const HcCellStat cs = r.cs[id];
        double* pw = RW.p + fab_off(RW, i, j, k);
        pw[0 * RW.nstride] = (double)cs.nst; pw[1 * RW.nstride] = (double)cs.netf; pw[2 * RW.nstride] = (double)cs.nfe; pw[3 * RW.nstride] = (double)cs.nni;
        pw[4 * RW.nstride] = (double)cs.ncfn; pw[5 * RW.nstride] = (double)cs.nsetups; pw[6 * RW.nstride] = 0.0; pw[7 * RW.nstride] = 0.0;
        pw[8 * RW.nstride] = (double)cs.nfe_ls;
    }
}

// reset_internal_e (Source/EOS/reset_internal_e.H:16-68) over all tiles: synchronises (rho e) and (rho E), records the change of
// (rho e) in the reset source.  Pure streaming: 8 doubles read, up to 3 written per cell.
__global__ void __launch_bounds__(256) hc_reset_e_kernel(const __grid_constant__ KernelArgs a) {
    const Consts& c = a.k;
    constexpr int U4 = 4;   // cells per thread and pass: all loads of a pass are issued before the first store
    int t0 = -1;
    for (long long base = (long long)blockIdx.x * (256 * U4); base < a.ncells; base += (long long)gridDim.x * (256 * U4)) {
        t0 = next_tile(a, t0, base);
        double v[U4][8];
        double* pu[U4]; double* pr[U4];
        long long ns[U4];
        bool on[U4];
#pragma unroll
        for (int u = 0; u < U4; ++u) {
            const long long id = base + u * 256 + threadIdx.x;
            on[u] = id < a.ncells;
            pu[u] = nullptr; pr[u] = nullptr; ns[u] = 0;
            if (on[u]) {
                int ti = t0;
                while (ti + 1 < a.ntiles && a.tiles[ti + 1].offset <= id) ++ti;
                const TileDesc& t = a.tiles[ti];
                const HcFab& Uf = t.f[F_STATE];
                const HcFab& D = t.f[F_DIAG];
                const HcFab& Rs = t.f[2];
                int i, j, k;
                cell_of(t, id, i, j, k);
                pu[u] = Uf.p + fab_off(Uf, i, j, k); ns[u] = Uf.nstride;
                pr[u] = Rs.p + fab_off(Rs, i, j, k);
#pragma unroll
                for (int n = 0; n < 6; ++n) v[u][n] = pu[u][n * ns[u]];
                v[u][6] = D.p[fab_off(D, i, j, k) + NE * D.nstride];
                v[u][7] = *pr[u];
            }
        }
#pragma unroll
        for (int u = 0; u < U4; ++u) {
            if (!on[u]) continue;
            const double rho = v[u][DENS];
            const double rhoInv = 1.0 / rho;
            const double Up = v[u][1] * rhoInv, Vp = v[u][2] * rhoInv, Wp = v[u][3] * rhoInv;
            const double ke = 0.5 * rho * (Up * Up + Vp * Vp + Wp * Wp);
            const double eden = v[u][EDEN], eint = v[u][EINT];
            const double rho_eint = eden - ke;
            if (rho_eint > 0.0 && rho_eint / eden > 1.0e-6 && a.interp == 0) {
                *pr[u] = rho_eint - eint;
                pu[u][EINT * ns[u]] = rho_eint;
            } else if (eint > 0.0) {
                *pr[u] = v[u][7] + 0.0;
                pu[u][EDEN * ns[u]] = eint + ke;
            } else if (eint <= 0.0) {
                const double mu = c.c_mu_num / (c.c_mu_den + v[u][6]);
                const double eint_new = a.small_temp / (c.gm1 * mp_over_kb * mu);
                const double re = rho * eint_new;
                *pr[u] = re - eint;
                pu[u][EINT * ns[u]] = re;
                pu[u][EDEN * ns[u]] = re + ke;
            }
        }
    }
}

#include "hc_sources.cuh"
#include "hc_sorted.cuh"

// FP64 FMA throughput probe: 8 independent chains per thread, explicit __fma_rn (unaffected by -fmad=false)
__global__ void __launch_bounds__(256) hc_dfma_peak_kernel(double* out, int iters, double seed) {
    double x0 = seed + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const double m = 1.0 + 1e-9, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fma_rn(x0, m, b); x1 = __fma_rn(x1, m, b); x2 = __fma_rn(x2, m, b); x3 = __fma_rn(x3, m, b);
            x4 = __fma_rn(x4, m, b); x5 = __fma_rn(x5, m, b); x6 = __fma_rn(x6, m, b); x7 = __fma_rn(x7, m, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void hc_log10_selftest_kernel(const double* logtab, const double* x, double* y, int* bad, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bool b = false;
        y[i] = fast_log10(logtab, x[i], b);
        bad[i] = b ? 1 : 0;
    }
}

__global__ void hc_divdelta_selftest_kernel(const double* x, double* y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = div_delta_t(x[i]);
}

// ---------------------------------------------------------------------------------------------- host state
thread_local char g_err[512] = "";
thread_local double g_last_kernel_ms = 0.0, g_last_drain_ms = 0.0;
void set_err(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_err("%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return HC_ERR_CUDA; } } while (0)

struct DeviceTables {
    double* ionx = nullptr;
    double* iony = nullptr;
    double* cool = nullptr;
    double* logtab = nullptr;
    int sm_count = 0;
    std::atomic<bool> attr_set[5] = {{false}, {false}, {false}, {false}, {false}};   // set-once flags (the attribute calls themselves are idempotent)
};
std::mutex g_mu;
// host copy of the rates image: replaced as a whole by hc_tables_upload under g_mu; every reader takes a reference-counted snapshot under the
// same mutex, so a concurrent upload (another device, another thread) cannot pull the image from under a launch that is building its constants
std::shared_ptr<const std::vector<double>> g_rates_ptr;
std::shared_ptr<const std::vector<double>> rates_snapshot() {
    std::lock_guard<std::mutex> lock(g_mu);
    return g_rates_ptr;
}
DeviceTables g_dev[64];               // indexed by CUDA device ordinal

int current_device(int& dev) {
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_err("unsupported device ordinal %d", dev); return HC_ERR_CUDA; }
    return HC_OK;
}

// Small host-to-device copies (tile descriptors) go through a ring of PINNED staging slots: cudaMemcpyAsync from pageable memory lets the
// driver drain the stream before it stages the data, which exposes the host side of every call on an otherwise back-to-back stream (measured
// on the streaming kernels of SURVEY 8f rank 2: 1.45 ms per call for a 1.15 ms kernel).  A slot is reused only after the copy that read it
// last has completed (one event per slot).  Also raises the release threshold of the device's stream-ordered pool once, so that the scratch
// of a call is not handed back to the driver at every synchronisation.
__global__ void hc_copy_words_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
static_assert(sizeof(TileDesc) % 8 == 0, "tile descriptors are copied in 8-byte words");
struct StageRing {
    static constexpr int SLOTS = 8;
    static constexpr size_t SLOT_BYTES = 256 * 1024;
    char* base = nullptr;       // host address of the pinned, device-mapped ring
    char* dev_base = nullptr;   // the same memory as the device sees it (equal to `base` under unified addressing)
    cudaEvent_t ev[SLOTS] = {};
    int next = 0;
    bool ready = false, pool_set = false;
};
StageRing g_ring[64];
std::mutex g_ring_mu;

int copy_small_h2d(int dev, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_ring_mu);
    StageRing& r = g_ring[dev];
    if (!r.pool_set) {
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        r.pool_set = true;
    }
    static const bool pageable = std::getenv("NYX_HC_PAGEABLE_DESC") != nullptr;   // diagnostics: the old behaviour, for A/B timing
    if (pageable || bytes > StageRing::SLOT_BYTES) {   // thousands of tiles: the pageable path (a one-off drain is small against such a launch)
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return HC_OK;
    }
    if (!r.ready) {
        CUDA_TRY(cudaHostAlloc((void**)&r.base, StageRing::SLOTS * StageRing::SLOT_BYTES, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer((void**)&r.dev_base, r.base, 0));
        for (int i = 0; i < StageRing::SLOTS; ++i) CUDA_TRY(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
        r.ready = true;
    }
    const int slot = r.next;
    r.next = (r.next + 1) % StageRing::SLOTS;
    CUDA_TRY(cudaEventSynchronize(r.ev[slot]));   // a never-recorded event is complete
    char* pin = r.base + (size_t)slot * StageRing::SLOT_BYTES;
    std::memcpy(pin, src, bytes);
    // a KERNEL reads the pinned slot (zero-copy) and writes the device copy: stream-ordered on the compute engine.  A DMA copy of these few
    // kilobytes shares a copy engine with the bulk FAB transfers of the host-buffer pipeline and waited behind them (measured: 518-530 ms
    // instead of 507 ms per 512^3 step end to end).
    const int words = (int)((bytes + 7) / 8);
    hc_copy_words_kernel<<<(words + 255) / 256, 256, 0, stream>>>(reinterpret_cast<unsigned long long*>(dst),
                                                                  reinterpret_cast<const unsigned long long*>(r.dev_base + (size_t)slot * StageRing::SLOT_BYTES), words);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(r.ev[slot], stream));
    return HC_OK;
}

// Grid of the streaming kernels (rows either side of the path): experiment knob NYX_HC_STREAM_CTAS = CTAs per SM of the grid-stride launch
// (default 8); 0 = one pass per CTA (as many CTAs as the cells need)
int stream_grid(long long ncells, long long per_cta, int sms, int dflt_per_sm = 8) {
    static const int knob = [] { const char* e = std::getenv("NYX_HC_STREAM_CTAS"); return e ? std::atoi(e) : -1; }();
    const long long want = (ncells + per_cta - 1) / per_cta;
    const int per_sm = knob >= 0 ? knob : dflt_per_sm;
    if (per_sm == 0) return (int)std::min<long long>(want, 0x7fffffffLL);
    return (int)std::min<long long>(want, (long long)sms * per_sm);
}

// multiprocessor count, queried once per device (the rank-2/4 streaming launchers do not need the rate tables of DeviceTables)
int sm_count_of(int dev, int& sms) {
    static std::atomic<int> cached[64];
    int v = cached[dev].load(std::memory_order_relaxed);
    if (!v) { CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); cached[dev].store(v, std::memory_order_relaxed); }
    sms = v;
    return HC_OK;
}

bool valid_params(const HcParams* p) {
    return p && p->rtol > 0.0 && p->atol_factor >= 0.0 && p->h_species > 0.0 && p->h_species <= 1.0;
}

TileDesc make_tile(const HcFab* const* fabs, int nf, int idx, const HcBox& b, long long offset, long long chunk_begin) {
    TileDesc t{};
    for (int s = 0; s < nf; ++s) t.f[s] = fabs[s][idx];
    for (int d = 0; d < 3; ++d) t.lo[d] = b.lo[d];
    t.nx = b.hi[0] - b.lo[0] + 1; t.ny = b.hi[1] - b.lo[1] + 1; t.nz = b.hi[2] - b.lo[2] + 1;
    t.offset = offset;
    t.chunk_begin = chunk_begin;
    if (t.nx > 0) {
        // a chunk is a piece of one x-row: rows longer than CHUNK_MAX are cut into equal pieces (NYX_HC_CHUNK: measurement knob)
        static const int chunk_max = [] { const char* e = std::getenv("NYX_HC_CHUNK"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : CHUNK_MAX; }();
        t.cpr = (t.nx + chunk_max - 1) / chunk_max;
        t.chunk_len = (t.nx + t.cpr - 1) / t.cpr;
    }
    return t;
}

bool tile_inside(const TileDesc& t, int nf) {
    for (int s = 0; s < nf; ++s) {
        const HcFab& f = t.f[s];
        if (!f.p) return false;
        const int hi[3] = {t.lo[0] + t.nx - 1, t.lo[1] + t.ny - 1, t.lo[2] + t.nz - 1};
        for (int d = 0; d < 3; ++d) if (t.lo[d] < f.lo[d] || hi[d] > f.hi[d]) return false;
    }
    return true;
}

template <typename KernelT>
int set_smem_attr(KernelT kernel, DeviceTables& dt, int slot, size_t bytes) {
    if (!dt.attr_set[slot]) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        // ask for the smallest shared-memory carve-out that holds one CTA (+ the 1 KB the system reserves + static): the rest of the 256 KB
        // is L1 for the rate tables (carve-out steps ... 132, 164, 196, 228 KB)
        static const int knob = [] { const char* e = std::getenv("NYX_HC_CARVEOUT_KB"); return e ? std::atoi(e) : 0; }();
        const int want_kb = knob > 0 ? knob : (int)((bytes + 4096 + 1023) / 1024);
        const int pct = std::min(100, (want_kb * 100 + 227) / 228);
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        dt.attr_set[slot] = true;
    }
    return HC_OK;
}

// stream-ordered scratch of a launcher, released on EVERY exit path (early error returns included)
struct StreamScratch {
    cudaStream_t stream;
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    explicit StreamScratch(cudaStream_t s) : stream(s) {}
    StreamScratch(const StreamScratch&) = delete;
    StreamScratch& operator=(const StreamScratch&) = delete;
    void own(void* q) { if (q && n < 4) p[n++] = q; }
    ~StreamScratch() { for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], stream); }
};

// Common launcher: build tile descriptors, stream-ordered scratch, launch, optionally read the statistics back.
// `ext_dstats` (device, 14 x u64, zeroed by the caller): accumulate the statistics there instead (several launches of one call).
struct EosOpts {
    int mode = 0, max_temp_dt = 0, interp = 0;
    double small_temp = 0.0, large_temp = 0.0;
};
constexpr int PATH_RESET_E = 3;   // hc_reset_e_kernel (PATH_EOS = 2: hc_eos_kernel)
// the three SAVE_REACT FABs of every tile (device memory)
struct ReactFabs {
    const HcFab* in; const HcFab* out; const HcFab* work;
};

int launch(int path, int ntiles, const HcFab* const* fabs, int nf, const HcBox* tiles, const Consts& k, HcStats* stats,
           HcCellStat* cell_stats, cudaStream_t stream, unsigned long long* ext_dstats = nullptr, const EosOpts* eos = nullptr,
           const ReactFabs* react = nullptr) {
    int dev; if (int rc = current_device(dev)) return rc;
    DeviceTables& dt = g_dev[dev];
    if (!dt.ionx) { set_err("hc_tables_upload has not been called on device %d", dev); return HC_ERR_NO_TABLES; }
    if (ntiles < 0) { set_err("ntiles < 0"); return HC_ERR_ARG; }
    std::vector<TileDesc> h_tiles; h_tiles.reserve(ntiles);
    std::vector<HcFab> h_react;
    long long ncells = 0, nchunks = 0;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc td = make_tile(fabs, nf, t, tiles[t], ncells, nchunks);
        if (td.nx <= 0 || td.ny <= 0 || td.nz <= 0) continue;   // empty tile: nothing to do (as an empty MFIter tile)
        if (!tile_inside(td, nf)) { set_err("tile %d is not contained in its FABs (or a FAB pointer is null)", t); return HC_ERR_ARG; }
        if (react) {
            const HcFab rf[3] = {react->in[t], react->out[t], react->work[t]};
            const int need[3] = {7, 7, 9};
            for (int s = 0; s < 3; ++s) {
                bool ok = rf[s].p && rf[s].ncomp >= need[s];
                for (int d = 0; d < 3 && ok; ++d) ok = tiles[t].lo[d] >= rf[s].lo[d] && tiles[t].hi[d] <= rf[s].hi[d];
                if (!ok) { set_err("tile %d: react FAB %d is null, has fewer than %d components or does not contain the tile", t, s, need[s]); return HC_ERR_ARG; }
                h_react.push_back(rf[s]);
            }
        }
        // a lane remembers its cell as (tile: 20 bits, k: 12 bits, i, j: 16 bits each, relative to the tile)
        if ((path == PATH_VEC || path == PATH_STRUCT) && (td.nx > 65536 || td.ny > 65536 || td.nz > 4096 || h_tiles.size() >= (1u << 20))) {
            set_err("tile %d: the integrator kernels take tiles of at most 65536 x 65536 x 4096 cells and at most 2^20 tiles per call", t); return HC_ERR_ARG;
        }
        ncells += (long long)td.nx * td.ny * td.nz;
        nchunks += (long long)td.cpr * td.ny * td.nz;
        h_tiles.push_back(td);
    }
    if (k.max_steps > 65535) { set_err("max_steps > 65535 (the step counters of a lane are 16-bit; CVodeSetMaxNumSteps is 2000 in Nyx)"); return HC_ERR_ARG; }
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (ncells == 0) return HC_OK;

    const size_t tiles_bytes = h_tiles.size() * sizeof(TileDesc);
    const size_t react_bytes = h_react.size() * sizeof(HcFab);
    const size_t scratch_bytes = 256 + tiles_bytes + react_bytes;   // [queue u64][pad][stats 14 x u64][pad] [tiles] [react FABs]
    char* scratch = nullptr;
    StreamScratch owned(stream);
    CUDA_TRY(cudaMallocAsync((void**)&scratch, scratch_bytes, stream));
    owned.own(scratch);
    CUDA_TRY(cudaMemsetAsync(scratch, 0, 256, stream));
    if (int rc = copy_small_h2d(dev, scratch + 256, h_tiles.data(), tiles_bytes, stream)) return rc;
    // SAVE_REACT: the per-cell record of the REACT kernel and (unless the caller asked for them anyway) the per-cell counters
    double* react_raw = nullptr;
    HcCellStat* own_cell_stats = nullptr;
    if (react) {
        if (int rc = copy_small_h2d(dev, scratch + 256 + tiles_bytes, h_react.data(), react_bytes, stream)) return rc;
        CUDA_TRY(cudaMallocAsync((void**)&react_raw, (size_t)ncells * 4 * sizeof(double), stream));
        owned.own(react_raw);
        if (!cell_stats) {
            CUDA_TRY(cudaMallocAsync((void**)&own_cell_stats, (size_t)ncells * sizeof(HcCellStat), stream));
            owned.own(own_cell_stats);
            cell_stats = own_cell_stats;
        }
    }

    KernelArgs a{};
    a.k = k;
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = (int)h_tiles.size();
    a.ncells = ncells;
    a.nchunks = nchunks;
    a.queue = reinterpret_cast<unsigned long long*>(scratch);
    a.dstats = ext_dstats ? ext_dstats : reinterpret_cast<unsigned long long*>(scratch + 64);
    a.timing = (stats && !ext_dstats) ? reinterpret_cast<unsigned long long*>(scratch + 192) : nullptr;
    a.cell_stats = cell_stats;
    a.react_raw = react_raw;
    a.ionx = dt.ionx; a.iony = dt.iony; a.cool = dt.cool; a.logtab = dt.logtab;
    if (eos) { a.eos_mode = eos->mode; a.max_temp_dt = eos->max_temp_dt; a.interp = eos->interp; a.small_temp = eos->small_temp; a.large_temp = eos->large_temp; }

    constexpr int LV = HC_SORTED_LANES_VEC, LS = HC_SORTED_LANES_STRUCT;
    if (path == PATH_VEC) {
        const int g = (int)std::min<long long>((ncells + LV - 1) / LV, (long long)dt.sm_count * HC_SORTED_CTAS);
        if (int rc = set_smem_attr(sorted::hc_sorted_kernel<PATH_VEC, LV>, dt, 0, sorted::Layout<PATH_VEC, LV>::total)) return rc;
        sorted::hc_sorted_kernel<PATH_VEC, LV><<<g, LV, sorted::Layout<PATH_VEC, LV>::total, stream>>>(a);
    } else if (path == PATH_STRUCT && react) {
        const int g = (int)std::min<long long>((ncells + LS - 1) / LS, (long long)dt.sm_count * HC_SORTED_CTAS);
        if (int rc = set_smem_attr(sorted::hc_sorted_kernel<PATH_STRUCT, LS, true>, dt, 3, sorted::Layout<PATH_STRUCT, LS>::total)) return rc;
        sorted::hc_sorted_kernel<PATH_STRUCT, LS, true><<<g, LS, sorted::Layout<PATH_STRUCT, LS>::total, stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (int rc = set_smem_attr(hc_react_kernel, dt, 4, SMEM_EOS)) return rc;
        ReactArgs ra{reinterpret_cast<const HcFab*>(scratch + 256 + tiles_bytes), react_raw, cell_stats};
        const int ge = (int)std::min<long long>((ncells + EOS_THREADS - 1) / EOS_THREADS, dt.sm_count);
        hc_react_kernel<<<ge, EOS_THREADS, SMEM_EOS, stream>>>(a, ra);
    } else if (path == PATH_STRUCT) {
        const int g = (int)std::min<long long>((ncells + LS - 1) / LS, (long long)dt.sm_count * HC_SORTED_CTAS);
        if (int rc = set_smem_attr(sorted::hc_sorted_kernel<PATH_STRUCT, LS>, dt, 1, sorted::Layout<PATH_STRUCT, LS>::total)) return rc;
        sorted::hc_sorted_kernel<PATH_STRUCT, LS><<<g, LS, sorted::Layout<PATH_STRUCT, LS>::total, stream>>>(a);
    } else {
        if (path == PATH_RESET_E) {
            const int g = stream_grid(ncells, 1024, dt.sm_count);
            hc_reset_e_kernel<<<g, 256, 0, stream>>>(a);
        } else {
            if (int rc = set_smem_attr(hc_eos_kernel, dt, 2, SMEM_EOS)) return rc;
            const int ge = (int)std::min<long long>((ncells + EOS_THREADS - 1) / EOS_THREADS, dt.sm_count);
            hc_eos_kernel<<<ge, EOS_THREADS, SMEM_EOS, stream>>>(a);
        }
    }
    CUDA_TRY(cudaGetLastError());
    if (stats && !ext_dstats) {
        // the pageable `stats` target makes this copy synchronous with respect to the host
        CUDA_TRY(cudaMemcpyAsync(stats, scratch + 64, sizeof(HcStats), cudaMemcpyDeviceToHost, stream));
        unsigned long long tm[3] = {0, 0, 0};
        CUDA_TRY(cudaMemcpyAsync(tm, scratch + 192, sizeof tm, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (path == PATH_VEC || path == PATH_STRUCT) {
            // kernel span and drain tail (first "work queue empty" -> last CTA exit) of this launch, for hc_last_launch_timing
            const unsigned long long t_empty = ~tm[0], t_end = tm[1], t_start = ~tm[2];
            g_last_kernel_ms = (t_end > t_start) ? 1e-6 * (double)(t_end - t_start) : 0.0;
            g_last_drain_ms = (tm[0] != 0 && t_end > t_empty) ? 1e-6 * (double)(t_end - t_empty) : 0.0;
        }
    }
    return HC_OK;   // `owned` releases the scratch, stream-ordered after the kernels
}

// ---- host-buffer entry points: staged and pipelined --------------------------------------------------------------------
// The tiles are cut into groups; group g+1 is copied to the device (stream h2d) while group g is integrated (streams comp / comp2,
// alternating) and group g-1 is copied back (stream d2h): except for the first H2D and the last D2H the transfers hide behind the kernel.
// Consecutive groups run on two different compute streams so that the CTAs of kernel g+1 move onto the multiprocessors as the CTAs of
// the persistent kernel g run out of work: its drain tail (~1.8 ms) overlaps the next kernel instead of idling the device once per group.
// Device buffers come from the stream-ordered pool (its release threshold is raised once, so that repeated calls do not
// re-acquire memory from the driver).
struct HostSlot {
    const HcFab* host;          // ntiles host FABs
    std::vector<int> in, out;   // components copied to the device before / back to the host after the kernel
    int halo = 0;               // cells around a tile its kernel reads from this FAB (the border copy of the conservative density fix: 2)
    bool whole = false;         // the kernel indexes this FAB outside the tile's index range (the coarse zhi FAB of init_zhi): always copied whole
};
struct HostPipe {
    cudaStream_t h2d = nullptr, comp = nullptr, comp2 = nullptr, d2h = nullptr;
    bool ready = false;
    // Three persistent device slabs, used round-robin by the groups of a call: one copying in, one computing, one copying out.  A slab is
    // handed to group g once the D2H of group g-3 has finished (event, waited for by the h2d STREAM: the host never blocks).  No allocator in
    // the loop: with stream-ordered allocation the blocks a group released on the d2h stream came back to the h2d stream with a dependency
    // on that release, and the H2D of group g+1 stopped overlapping kernel g as soon as the host no longer paced the loop by accident.
    static constexpr int NSLAB = 3;
    char* slab[NSLAB] = {nullptr, nullptr, nullptr};
    size_t cap[NSLAB] = {0, 0, 0};
    cudaEvent_t slab_free[NSLAB] = {};
    bool slab_busy[NSLAB] = {false, false, false};
    std::mutex call_mu;   // host-buffer calls on one device run one after the other (they share the slabs and the three streams)
};
HostPipe g_pipe[64];
#if !defined(HC_HOST_GROUPS)
#define HC_HOST_GROUPS 8
#endif
constexpr int HOST_GROUPS = HC_HOST_GROUPS;
// measurement knob: NYX_HC_HOST_ONE_COMP_STREAM=1 puts every group's kernel on the same stream (the round-1 pipeline)
#if !defined(HC_HOST_TAPER)
#define HC_HOST_TAPER 2   // 0: equal groups; 1: first group half a share; 2: geometric ramps at both ends (512^3 step end to end, profiles/r2_s23_host_taper.log:
                          //    Strang 440.0 against 441.3 ms, SDC 703.5 against 715.5 ms with 1)
#endif
constexpr bool HOST_TAPER = (HC_HOST_TAPER != 0);
size_t fab_doubles(const HcFab& f) { return (size_t)f.nstride * f.ncomp; }

int host_pipe(int dev, HostPipe*& hp) {
    std::lock_guard<std::mutex> lock(g_mu);
    hp = &g_pipe[dev];
    if (!hp->ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->h2d, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->comp, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->comp2, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&hp->d2h, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        for (int i = 0; i < HostPipe::NSLAB; ++i) CUDA_TRY(cudaEventCreateWithFlags(&hp->slab_free[i], cudaEventDisableTiming));
        hp->ready = true;
    }
    return HC_OK;
}

using GroupLauncher = std::function<int(int n, const HcFab* const* fabs, const HcBox* tiles, cudaStream_t stream)>;

// Copy of component c of one FAB restricted to the cells of `b`, device <-> host (same strides on both sides).  D2H of the tile: components
// that were NOT uploaded (pure outputs), so that host cells outside the tile -- ghost cells, other tiles of the same FAB -- keep their values.
// Both directions: FABs with ghost cells, of which only the tiles' cells (+ halo) are staged.  Contiguous when the box spans the FAB in x
// and y (one block of nz planes), a pitched 3-D copy otherwise.
int copy_box(const HcFab& h, const double* dbase, int c, const HcBox& b, cudaMemcpyKind kind, cudaStream_t stream);
int copy_tile_d2h(const HcFab& h, const double* dbase, int c, const HcBox& b, cudaStream_t stream) {
    return copy_box(h, dbase, c, b, cudaMemcpyDeviceToHost, stream);
}
int copy_box(const HcFab& h, const double* dbase, int c, const HcBox& b, cudaMemcpyKind kind, cudaStream_t stream) {
    const bool d2h = (kind == cudaMemcpyDeviceToHost);
    const long long nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
    if (nx <= 0 || ny <= 0 || nz <= 0) return HC_OK;
    const long long off = (b.lo[0] - h.lo[0]) + (b.lo[1] - h.lo[1]) * h.jstride + (b.lo[2] - h.lo[2]) * h.kstride + (long long)c * h.nstride;
    double* dev = const_cast<double*>(dbase) + off;
    double* hst = h.p + off;
    if (nx == h.jstride && ny * h.jstride == h.kstride) {
        CUDA_TRY(cudaMemcpyAsync(d2h ? hst : dev, d2h ? dev : hst, (size_t)(nz * h.kstride) * sizeof(double), kind, stream));
        return HC_OK;
    }
    if (h.kstride % h.jstride != 0) { set_err("FAB strides are not those of a box (kstride is not a multiple of jstride)"); return HC_ERR_ARG; }
    cudaMemcpy3DParms prm{};
    const size_t pitch = (size_t)h.jstride * sizeof(double), rows = (size_t)(h.kstride / h.jstride);
    // the pitched pointers start at the tile's first cell: the row pitch and the rows per plane are the FAB's
    prm.srcPtr = make_cudaPitchedPtr(d2h ? dev : hst, pitch, (size_t)nx * sizeof(double), rows);
    prm.dstPtr = make_cudaPitchedPtr(d2h ? hst : dev, pitch, (size_t)nx * sizeof(double), rows);
    prm.extent = make_cudaExtent((size_t)nx * sizeof(double), (size_t)ny, (size_t)nz);
    prm.kind = kind;
    CUDA_TRY(cudaMemcpy3DAsync(&prm, stream));
    return HC_OK;
}

// everything a host-buffer call must hand back on EVERY exit path (early error returns included)
struct HostCallGuard {
    HostPipe* hp;
    unsigned long long* dstats = nullptr;
    std::vector<cudaEvent_t> events;
    explicit HostCallGuard(HostPipe* p) : hp(p) {}
    ~HostCallGuard() {
        // drain the three streams first: the slabs and the events may still be in use by queued work
        cudaStreamSynchronize(hp->h2d); cudaStreamSynchronize(hp->comp); cudaStreamSynchronize(hp->comp2); cudaStreamSynchronize(hp->d2h);
        for (int i = 0; i < HostPipe::NSLAB; ++i) hp->slab_busy[i] = false;
        if (dstats) cudaFreeAsync(dstats, hp->comp);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    cudaError_t new_event(cudaEvent_t& e) {
        cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (r == cudaSuccess) events.push_back(e);
        return r;
    }
};

int run_host(int path, int ntiles, std::vector<HostSlot>& slots, const HcBox* tiles, const Consts& k, HcStats* stats, const EosOpts* eos = nullptr,
             const GroupLauncher* custom = nullptr, bool with_react = false) {
    // with_react: the last three slots are the SAVE_REACT FABs (pure outputs), the ones before them the integrator's
    int dev; if (int rc = current_device(dev)) return rc;
    HostPipe* hp = nullptr;
    if (int rc = host_pipe(dev, hp)) return rc;
    const int nf = (int)slots.size();
    for (const HostSlot& sl : slots)
        for (int t = 0; t < ntiles; ++t) if (!sl.host[t].p) { set_err("null host FAB"); return HC_ERR_ARG; }
    // Several tiles may share a FAB (a tiled MFIter of a CPU build of AMReX): such tiles get ONE device copy, uploaded once, and stay in
    // one group.  They must be consecutive (as MFIter delivers them); anything else is rejected rather than silently losing updates.
    auto shares_fab = [&](int t, int u) {
        for (int s = 0; s < nf; ++s) if (slots[s].host[t].p == slots[s].host[u].p) return true;
        return false;
    };
    {
        std::vector<const double*> seen;
        for (int s = 0; s < nf; ++s) {
            seen.clear();
            for (int t = 0; t < ntiles; ++t) {
                const double* p = slots[s].host[t].p;
                if (t > 0 && p == slots[s].host[t - 1].p) continue;
                if (std::find(seen.begin(), seen.end(), p) != seen.end()) {
                    set_err("tiles %d...: tiles that share a host FAB must be consecutive in the tile list", t); return HC_ERR_ARG;
                }
                seen.push_back(p);
            }
        }
    }
    // groups of consecutive tiles with about 1/HOST_GROUPS of the cells each
    long long total = 0;
    std::vector<long long> cells(ntiles);
    for (int t = 0; t < ntiles; ++t) {
        const HcBox& b = tiles[t];
        cells[t] = (long long)std::max(0, b.hi[0] - b.lo[0] + 1) * std::max(0, b.hi[1] - b.lo[1] + 1) * std::max(0, b.hi[2] - b.lo[2] + 1);
        total += cells[t];
    }
    const long long per_group = std::max<long long>((total + HOST_GROUPS - 1) / HOST_GROUPS, 1);
    std::lock_guard<std::mutex> call_lock(hp->call_mu);
    HostCallGuard guard(hp);
    CUDA_TRY(cudaMallocAsync((void**)&guard.dstats, 128, hp->comp));
    unsigned long long* dstats = guard.dstats;
    CUDA_TRY(cudaMemsetAsync(dstats, 0, 128, hp->comp));
    static const bool one_comp_stream = [] { const char* e = std::getenv("NYX_HC_HOST_ONE_COMP_STREAM"); return e && std::atoi(e) != 0; }();
    cudaEvent_t e_zero, e_last2 = nullptr;
    CUDA_TRY(guard.new_event(e_zero));
    CUDA_TRY(cudaEventRecord(e_zero, hp->comp));
    CUDA_TRY(cudaStreamWaitEvent(hp->comp2, e_zero, 0));     // the statistics block is cleared before either stream's first kernel
    int rc = HC_OK;
    std::vector<std::vector<HcFab>> dfab(nf);
    int group = 0;
    for (int t0 = 0; t0 < ntiles && rc == HC_OK; ++group) {
        // the first group is half a share: its H2D is not hidden behind any kernel; the last group then is the remaining half share, whose
        // D2H is not hidden either (measured: 496.8 ms instead of 507.0 ms per 512^3 step; a finer ramp 1/32, 1/16, 1/8 ... 3/32, 1/16: 495.5 ms, not kept)
        long long share = (group == 0 && HOST_TAPER) ? std::max<long long>(per_group / 2, 1) : per_group;
#if HC_HOST_TAPER == 2
        // geometric ramps at both ends: the first H2D and the last D2H are the only transfers no kernel hides, so the first groups take
        // 1/64, 1/32, 1/16 of the cells and the last ones half of what is left each (the drain tail of a small group's kernel overlaps the next
        // kernel: consecutive groups run on alternating streams)
        {
            long long done = 0;
            for (int t = 0; t < t0; ++t) done += cells[t];
            const long long unit = std::max<long long>(total / 64, 1);
            const long long head = (group < 30) ? (unit << std::min(group, 20)) : per_group;
            const long long tail = std::max<long long>((total - done) / 2, unit);
            share = std::min(per_group, std::min(head, tail));
        }
#endif
        int t1 = t0; long long acc = 0;
        while (t1 < ntiles && (t1 == t0 || acc + cells[t1] <= share || shares_fab(t1, t1 - 1))) acc += cells[t1++];
        const int n = t1 - t0;
        // first tile of each FAB of this group (per slot): the one that owns the device copy
        auto first_of = [&](int s, int i) { int j = i; while (j > 0 && slots[s].host[t0 + j - 1].p == slots[s].host[t0 + i].p) --j; return j; };
        // this group's slab: large enough for all its FABs (256-byte aligned), free once its previous user has copied out
        const int b = group % HostPipe::NSLAB;
        size_t need = 0;
        for (int s = 0; s < nf; ++s)
            for (int i = 0; i < n; ++i) if (first_of(s, i) == i) need += (fab_doubles(slots[s].host[t0 + i]) * sizeof(double) + 255) / 256 * 256;
        if (need > hp->cap[b]) {
            if (hp->slab_busy[b]) CUDA_TRY(cudaEventSynchronize(hp->slab_free[b]));
            if (hp->slab[b]) CUDA_TRY(cudaFree(hp->slab[b]));
            hp->slab[b] = nullptr; hp->cap[b] = 0;
            CUDA_TRY(cudaMalloc((void**)&hp->slab[b], need));
            hp->cap[b] = need;
            hp->slab_busy[b] = false;
        }
        if (hp->slab_busy[b]) CUDA_TRY(cudaStreamWaitEvent(hp->h2d, hp->slab_free[b], 0));
        // What travels of a FAB: the bounding box of its tiles in this group (+ the slot's halo), clipped to the FAB -- or the whole
        // components (one contiguous copy each) when that box is most of the FAB anyway.  A state FAB of Nyx carries 4 ghost cells: a 32^3
        // box is 51 % of its 40^3 FAB, a 128^3 box 83 % of 136^3.
        std::vector<std::vector<HcBox>> stage(nf, std::vector<HcBox>(n));
        std::vector<std::vector<char>> whole(nf, std::vector<char>(n, 1));
        for (int s = 0; s < nf; ++s)
            for (int i = 0; i < n; ++i) {
                if (first_of(s, i) != i) continue;
                const HcFab& h = slots[s].host[t0 + i];
                HcBox bb{{INT_MAX, INT_MAX, INT_MAX}, {INT_MIN, INT_MIN, INT_MIN}};
                for (int j = i; j < n && first_of(s, j) == i; ++j)
                    for (int d = 0; d < 3; ++d) {
                        if (tiles[t0 + j].hi[d] < tiles[t0 + j].lo[d]) continue;
                        bb.lo[d] = std::min(bb.lo[d], tiles[t0 + j].lo[d] - slots[s].halo);
                        bb.hi[d] = std::max(bb.hi[d], tiles[t0 + j].hi[d] + slots[s].halo);
                    }
                long long vol = 1, fvol = 1;
                for (int d = 0; d < 3; ++d) {
                    bb.lo[d] = std::max(bb.lo[d], h.lo[d]); bb.hi[d] = std::min(bb.hi[d], h.hi[d]);
                    vol *= std::max(0, bb.hi[d] - bb.lo[d] + 1); fvol *= std::max(1, h.hi[d] - h.lo[d] + 1);
                }
                stage[s][i] = bb;
                // (strides that are not those of the FAB's own box: whole components, the pitched copy assumes them)
                const bool box_strides = h.jstride == (long long)(h.hi[0] - h.lo[0] + 1) && h.kstride == h.jstride * (h.hi[1] - h.lo[1] + 1);
                whole[s][i] = (slots[s].whole || !box_strides || vol * 100 >= fvol * 85) ? 1 : 0;
            }
        size_t off = 0;
        for (int s = 0; s < nf; ++s) {
            dfab[s].assign(slots[s].host + t0, slots[s].host + t1);
            for (int i = 0; i < n; ++i) {
                const HcFab& h = slots[s].host[t0 + i];
                const int own = first_of(s, i);
                if (own != i) { dfab[s][i].p = dfab[s][own].p; continue; }   // another tile of the same FAB: share its device copy
                double* d = reinterpret_cast<double*>(hp->slab[b] + off);
                off += (fab_doubles(h) * sizeof(double) + 255) / 256 * 256;
                dfab[s][i].p = d;
                for (int c : slots[s].in) {
                    if (c >= h.ncomp) continue;
                    if (whole[s][i]) CUDA_TRY(cudaMemcpyAsync(d + (size_t)c * h.nstride, h.p + (size_t)c * h.nstride, (size_t)h.nstride * sizeof(double),
                                                              cudaMemcpyHostToDevice, hp->h2d));
                    else if (int rc2 = copy_box(h, d, c, stage[s][i], cudaMemcpyHostToDevice, hp->h2d)) return rc2;
                }
            }
        }
        cudaEvent_t e_in, e_k;
        CUDA_TRY(guard.new_event(e_in)); CUDA_TRY(guard.new_event(e_k));
        // custom launchers (the two-pass source assembly) order their groups through one stream
        cudaStream_t cs = ((group & 1) && !one_comp_stream && !custom) ? hp->comp2 : hp->comp;
        CUDA_TRY(cudaEventRecord(e_in, hp->h2d));
        CUDA_TRY(cudaStreamWaitEvent(cs, e_in, 0));
        std::vector<const HcFab*> fabs(nf);
        for (int s = 0; s < nf; ++s) fabs[s] = dfab[s].data();
        if (custom) rc = (*custom)(n, fabs.data(), tiles + t0, cs);
        else if (with_react) {
            const ReactFabs rf{fabs[nf - 3], fabs[nf - 2], fabs[nf - 1]};
            rc = launch(path, n, fabs.data(), nf - 3, tiles + t0, k, nullptr, nullptr, cs, dstats, eos, &rf);
        } else rc = launch(path, n, fabs.data(), nf, tiles + t0, k, nullptr, nullptr, cs, dstats, eos);
        if (rc != HC_OK) break;
        CUDA_TRY(cudaEventRecord(e_k, cs));
        if (cs == hp->comp2) e_last2 = e_k;
        CUDA_TRY(cudaStreamWaitEvent(hp->d2h, e_k, 0));
        for (int s = 0; s < nf; ++s)
            for (int i = 0; i < n; ++i) {
                const HcFab& h = slots[s].host[t0 + i];
                const bool owner = (first_of(s, i) == i);
                for (int c : slots[s].out) {
                    if (c >= h.ncomp) continue;
                    const bool uploaded = std::find(slots[s].in.begin(), slots[s].in.end(), c) != slots[s].in.end();
                    if (uploaded && whole[s][first_of(s, i)]) {
                        // the device copy holds the whole component: one contiguous copy per FAB
                        if (owner) CUDA_TRY(cudaMemcpyAsync(h.p + (size_t)c * h.nstride, dfab[s][i].p + (size_t)c * h.nstride, (size_t)h.nstride * sizeof(double),
                                                            cudaMemcpyDeviceToHost, hp->d2h));
                    } else if (int rc2 = copy_tile_d2h(h, dfab[s][i].p, c, tiles[t0 + i], hp->d2h)) return rc2;   // the tile's cells only (pure outputs; FABs staged by sub-box)
                }
            }
        CUDA_TRY(cudaEventRecord(hp->slab_free[b], hp->d2h));
        hp->slab_busy[b] = true;
        t0 = t1;
    }
    if (rc == HC_OK && stats) {
        if (e_last2) CUDA_TRY(cudaStreamWaitEvent(hp->comp, e_last2, 0));
        CUDA_TRY(cudaMemcpyAsync(stats, dstats, sizeof(HcStats), cudaMemcpyDeviceToHost, hp->comp));
    }
    return rc;   // the guard drains the streams and releases the scratch
}


// ---- update_state_with_sources / enforce_minimum_density / MultiFab component operations ----------------------------
bool valid_src_params(const HcSrcParams* p) {
    return p && p->gamma_minus_1 > 0.0 && p->h_species > 0.0 && p->h_species <= 1.0;
}

// tile descriptors of `nf` FAB slots in stream-ordered scratch: [256 B header][tiles]; returns the number of cells
int stage_tiles(int ntiles, const HcFab* const* fabs, int nf, const HcBox* tiles, cudaStream_t stream, char*& scratch, int& n_used, long long& ncells,
                int nf_check = -1, long long* nchunks_out = nullptr) {
    std::vector<TileDesc> h_tiles; h_tiles.reserve(ntiles);
    ncells = 0; long long nchunks = 0;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc td = make_tile(fabs, nf, t, tiles[t], ncells, nchunks);
        if (td.nx <= 0 || td.ny <= 0 || td.nz <= 0) continue;
        if (!tile_inside(td, nf_check < 0 ? nf : nf_check)) { set_err("tile %d is not contained in its FABs (or a FAB pointer is null)", t); return HC_ERR_ARG; }
        ncells += (long long)td.nx * td.ny * td.nz;
        nchunks += (long long)td.cpr * td.ny * td.nz;
        h_tiles.push_back(td);
    }
    n_used = (int)h_tiles.size();
    if (nchunks_out) *nchunks_out = nchunks;
    scratch = nullptr;
    if (ncells == 0) return HC_OK;
    const size_t tiles_bytes = h_tiles.size() * sizeof(TileDesc);
    CUDA_TRY(cudaMallocAsync((void**)&scratch, 256 + tiles_bytes, stream));
    if (cudaMemsetAsync(scratch, 0xff, 256, stream) != cudaSuccess) {     // the minimum key starts at its largest value
        cudaFreeAsync(scratch, stream); scratch = nullptr;
        set_err("cudaMemsetAsync failed"); return HC_ERR_CUDA;
    }
    int dev = 0;
    int rc = current_device(dev);
    if (rc == HC_OK) rc = copy_small_h2d(dev, scratch + 256, h_tiles.data(), tiles_bytes, stream);
    if (rc != HC_OK) { cudaFreeAsync(scratch, stream); scratch = nullptr; }
    return rc;
}

SrcArgs make_src_args(double dt, double a_old, double a_new, const HcSrcParams& p) {
    SrcArgs a{};
    // the scalars of Nyx_update_state_with_sources.cpp:25-31, same expressions
    a.dt = dt; a.a_old = a_old;
    a.a_half = 0.5 * (a_old + a_new);
    a.a_half_inv = 1 / a.a_half;
    a.a_oldsq = a_old * a_old;
    a.a_newsq = a_new * a_new;
    a.a_new_inv = 1.0 / a_new;
    a.a_newsq_inv = 1.0 / a.a_newsq;
    a.dt_a_new = dt / a_new;
    a.a_half_dt = a.a_half * dt;
    a.dt_a_half = dt * a.a_half;
    a.small_dens = p.small_dens;
    // floor_density: e from nyx_eos_given_RT(small_temp, Ne = 0) (eos_hc.H:222-231), (rho e) = small_dens * e
    const double YHELIUM = (1.0 - p.h_species) / (4.0 * p.h_species);
    const double mu = (1.0 + 4.0 * YHELIUM) / (1.0 + YHELIUM + 0.0);
    const double eint_new = p.small_temp / (p.gamma_minus_1 * mp_over_kb * mu);
    a.floor_rhoe = p.small_dens * eint_new;
    a.sdc = p.sdc;
    return a;
}

// mode 0: sweep (1)+(3) and the minimum, then the enforce kernel predicated on this call's own minimum (single rank);
// mode 1: sweep (1)+(3) and the minimum accumulated into ext_min (device, caller-owned), no enforce kernel;
// mode 2: the enforce kernel unconditionally (the caller has reduced the minimum over ranks / groups)
// mode 3: ("conservative" variant) sweep (1) alone and the minimum;  mode 4 / 5: sweep (3) alone, 5 also resets hydro_src(rho) (SDC build)
int launch_sources(int mode, int ntiles, const HcFab* const* fabs, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams& p,
                   double* min_dens_out, unsigned long long* ext_min, cudaStream_t stream) {
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; if (int rc = sm_count_of(dev, sms)) return rc;
    char* scratch; int n_used; long long ncells;
    if (int rc = stage_tiles(ntiles, fabs, 5, tiles, stream, scratch, n_used, ncells)) return rc;
    StreamScratch owned(stream);
    owned.own(scratch);
    if (min_dens_out) *min_dens_out = DBL_MAX;
    if (ncells == 0) return HC_OK;
    SrcArgs a = make_src_args(dt, a_old, a_new, p);
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = n_used; a.ncells = ncells;
    a.min_key = ext_min ? ext_min : reinterpret_cast<unsigned long long*>(scratch);
    const long long per_cta = SRC_THREADS * SRC_U;
    const int grid = stream_grid(ncells, per_cta, sms, 0);   // one pass per CTA: measured 1.19 / 2.25 ms against 1.22 / 2.48 ms with 8 CTAs per SM striding
    if (mode == 3) hc_sources_kernel<false, 1><<<grid, SRC_THREADS, 0, stream>>>(a);                 // conservative variant: sweep (1) alone + minimum
    else if (mode == 4 || mode == 5) { a.reset_hsrc = (mode == 5); hc_sources_kernel<false, 2><<<grid, SRC_THREADS, 0, stream>>>(a); }   // sweep (3) alone
    else {
        if (mode != 2) hc_sources_kernel<false><<<grid, SRC_THREADS, 0, stream>>>(a);
        if (mode == 0) { a.use_flag = 1; hc_sources_kernel<true><<<grid, SRC_THREADS, 0, stream>>>(a); }
        if (mode == 2) { a.use_flag = 0; hc_sources_kernel<true><<<grid, SRC_THREADS, 0, stream>>>(a); }
    }
    CUDA_TRY(cudaGetLastError());
    if (min_dens_out && (mode == 0 || mode == 1 || mode == 3)) {
        unsigned long long key = 0;
        CUDA_TRY(cudaMemcpyAsync(&key, a.min_key, sizeof key, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        *min_dens_out = dens_from_key(key);
    }
    return HC_OK;
}

int check_src_fabs(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext, const HcFab* hs, const HcFab* grav) {
    for (int t = 0; t < ntiles; ++t) {
        if (s_old[t].ncomp != 6 || s_new[t].ncomp != 6 || ext[t].ncomp != 6 || hs[t].ncomp != 6 || grav[t].ncomp < 3) {
            set_err("tile %d: state, source FABs need 6 components (CONST_SPECIES build of the reference), grav_vector 3", t); return HC_ERR_ARG;
        }
    }
    return HC_OK;
}

int launch_fab_op(int op, int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, cudaStream_t stream) {
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; if (int rc = sm_count_of(dev, sms)) return rc;
    for (int t = 0; t < ntiles; ++t)
        if (dcomp < 0 || scomp < 0 || ncomp < 0 || dcomp + ncomp > dst[t].ncomp || scomp + ncomp > src[t].ncomp) { set_err("component range outside FAB %d", t); return HC_ERR_ARG; }
    const HcFab* fabs[2] = {dst, src};
    char* scratch; int n_used; long long ncells, nchunks = 0;
    if (int rc = stage_tiles(ntiles, fabs, 2, tiles, stream, scratch, n_used, ncells, -1, &nchunks)) return rc;
    StreamScratch owned(stream);
    owned.own(scratch);
    if (ncells == 0 || ncomp == 0) return HC_OK;
    FabOpArgs a{};
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = n_used; a.ncells = ncells; a.nchunks = nchunks; a.scomp = scomp; a.dcomp = dcomp; a.ncomp = ncomp; a.op = op;
    const int grid = stream_grid(nchunks, 8, sms);   // 8 warps per CTA, one chunk per warp and pass
    hc_fab_op_kernel<<<grid, 256, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return HC_OK;
}

}  // namespace

// ================================================================================================ C-ABI
extern "C" {

const char* hc_last_error(void) { return g_err; }
const char* hc_version(void) { return "nyx_hc 0.1 (sm_100a, strict-FP64)"; }

void hc_default_params(HcParams* p) { if (p) hc::default_params(p); }

int hc_tabulate_rates(const char* treecool_file, double mean_rhob, double* rates_out) {
    if (!treecool_file || !rates_out) { set_err("null argument"); return HC_ERR_ARG; }
    const int rc = hc::tabulate_rates(treecool_file, mean_rhob, rates_out);
    if (rc == HC_ERR_IO) set_err("cannot read %d rows x 7 columns from TREECOOL file %s", HC_NCOOLFILE, treecool_file);
    if (rc == HC_ERR_TREECOOL_LEN) set_err("TREECOOL file %s is longer than NCOOLFILE=%d rows", treecool_file, HC_NCOOLFILE);
    return rc;
}

int hc_tables_upload(const double* rates, size_t n_doubles) {
    if (!rates || n_doubles != (size_t)HC_RATES_DOUBLES) { set_err("rates image must hold %d doubles", HC_RATES_DOUBLES); return HC_ERR_ARG; }
    int dev; if (int rc = current_device(dev)) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    g_rates_ptr = std::make_shared<const std::vector<double>>(rates, rates + n_doubles);
    std::vector<double> ionx, iony, cool, logtab;
    interleave_tables(rates, ionx, iony, cool);
    build_log10_table(logtab);
    DeviceTables& dt = g_dev[dev];
    if (!dt.ionx) {
        CUDA_TRY(cudaMalloc((void**)&dt.ionx, ionx.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&dt.iony, iony.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&dt.cool, cool.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&dt.logtab, logtab.size() * sizeof(double)));
        CUDA_TRY(cudaDeviceGetAttribute(&dt.sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    CUDA_TRY(cudaMemcpy(dt.ionx, ionx.data(), ionx.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dt.iony, iony.data(), iony.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dt.cool, cool.data(), cool.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dt.logtab, logtab.data(), logtab.size() * sizeof(double), cudaMemcpyHostToDevice));
    return HC_OK;
}

int hc_uvb_at_z(double z, double* out6) {
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Uvb u = uvb_at_z(rates_sp->data(), z);
    out6[0] = u.ggh0; out6[1] = u.gghe0; out6[2] = u.gghep; out6[3] = u.eh0; out6[4] = u.ehe0; out6[5] = u.ehep;
    return HC_OK;
}

int hc_integrate_vec_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                           const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream) {
    if (!valid_params(prm) || (ntiles > 0 && (!state || !diag || !tiles)) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_vec(rates_sp->data(), *prm, a, dt);
    const HcFab* fabs[2] = {state, diag};
    return launch(PATH_VEC, ntiles, fabs, 2, tiles, k, stats, cell_stats, (cudaStream_t)stream);
}

int hc_integrate_vec(const HcFab* state, const HcFab* diag, HcBox tile, double a, double dt, const HcParams* prm,
                     HcStats* stats, HcCellStat* cell_stats, void* stream) {
    return hc_integrate_vec_batch(1, state, diag, &tile, a, dt, prm, stats, cell_stats, stream);
}

int hc_integrate_struct_batch(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                              const HcFab* reset_src, const HcFab* ir, const HcBox* tiles, double a, double a_end, double dt,
                              int sdc_iter, const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream) {
    if (!valid_params(prm) || (ntiles > 0 && (!s_old || !diag || !s_new || !hydro_src || !reset_src || !ir || !tiles)) || !(a > 0.0) || !(a_end > 0.0)) {
        set_err("bad argument"); return HC_ERR_ARG;
    }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_struct(rates_sp->data(), *prm, a, a_end, dt, sdc_iter);
    const HcFab* fabs[6] = {s_old, diag, s_new, hydro_src, reset_src, ir};
    return launch(PATH_STRUCT, ntiles, fabs, 6, tiles, k, stats, cell_stats, (cudaStream_t)stream);
}

int hc_integrate_struct_react_batch(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                                    const HcFab* reset_src, const HcFab* ir, const HcFab* react_in, const HcFab* react_out,
                                    const HcFab* react_out_work, const HcBox* tiles, double a, double a_end, double dt, int sdc_iter,
                                    const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream) {
    if (!valid_params(prm) || (ntiles > 0 && (!s_old || !diag || !s_new || !hydro_src || !reset_src || !ir || !react_in || !react_out || !react_out_work || !tiles)) ||
        !(a > 0.0) || !(a_end > 0.0)) {
        set_err("bad argument"); return HC_ERR_ARG;
    }
    // without sources the reference's ode_eos_save_react_arrays reads rhoe_src_vode / e_src_vode through null pointers (f_rhs_struct.H:197-198,251-252)
    if (sdc_iter < 0) { set_err("the SAVE_REACT dumps need the SDC sources (sdc_iter >= 0)"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_struct(rates_sp->data(), *prm, a, a_end, dt, sdc_iter);
    const HcFab* fabs[6] = {s_old, diag, s_new, hydro_src, reset_src, ir};
    const ReactFabs rf{react_in, react_out, react_out_work};
    return launch(PATH_STRUCT, ntiles, fabs, 6, tiles, k, stats, cell_stats, (cudaStream_t)stream, nullptr, nullptr, &rf);
}

int hc_integrate_struct(const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src, const HcFab* reset_src,
                        const HcFab* ir, HcBox tile, double a, double a_end, double dt, int sdc_iter, const HcParams* prm,
                        HcStats* stats, HcCellStat* cell_stats, void* stream) {
    return hc_integrate_struct_batch(1, s_old, diag, s_new, hydro_src, reset_src, ir, &tile, a, a_end, dt, sdc_iter, prm, stats,
                                     cell_stats, stream);
}

int hc_eos_T_given_Re(const HcFab* state, const HcFab* diag, HcBox tile, double a, const HcParams* prm, HcStats* stats, void* stream) {
    if (!valid_params(prm) || !state || !diag || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_eos(rates_sp->data(), *prm, a);
    const HcFab* fabs[2] = {state, diag};
    return launch(PATH_EOS, 1, fabs, 2, &tile, k, stats, nullptr, (cudaStream_t)stream);
}

namespace {
int launch_eos(int path, int ntiles, const HcFab* const* fabs, int nf, const HcBox* tiles, const Consts& k, const EosOpts& eos, HcStats* stats,
               cudaStream_t stream) {
    return launch(path, ntiles, fabs, nf, tiles, k, stats, nullptr, stream, nullptr, &eos);
}
}  // namespace

int hc_compute_new_temp_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, const HcParams* prm,
                              double small_temp, double large_temp, int max_temp_dt, HcStats* stats, void* stream) {
    if (!valid_params(prm) || ntiles < 0 || (ntiles > 0 && (!state || !diag || !tiles)) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_eos(rates_sp->data(), *prm, a);
    EosOpts eos; eos.mode = 1; eos.small_temp = small_temp; eos.large_temp = large_temp; eos.max_temp_dt = max_temp_dt;
    const HcFab* fabs[2] = {state, diag};
    return launch_eos(PATH_EOS, ntiles, fabs, 2, tiles, k, eos, stats, (cudaStream_t)stream);
}

int hc_reset_internal_energy_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcFab* reset_src, const HcBox* tiles, double a,
                                   const HcParams* prm, double small_temp, int interp, void* stream) {
    if (!valid_params(prm) || ntiles < 0 || (ntiles > 0 && (!state || !diag || !reset_src || !tiles)) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_eos(rates_sp->data(), *prm, a);
    EosOpts eos; eos.small_temp = small_temp; eos.interp = interp;
    const HcFab* fabs[3] = {state, diag, reset_src};
    return launch_eos(PATH_RESET_E, ntiles, fabs, 3, tiles, k, eos, nullptr, (cudaStream_t)stream);
}

int hc_integrate_vec_host(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                          const HcParams* prm, HcStats* stats) {
    if (ntiles <= 0 || !state || !diag || !tiles || !valid_params(prm) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    if (stats) std::memset(stats, 0, sizeof *stats);
    const Consts k = make_consts_vec(rates_sp->data(), *prm, a, dt);
    // diag(Temp, Ne) are dead inputs of the Strang path (eos_hc.H:151; load_cell never reads them): pure outputs, not uploaded
    std::vector<HostSlot> slots = {{state, {DENS, EDEN, EINT}, {EDEN, EINT}}, {diag, {}, {TEMP, NE}}};
    return run_host(PATH_VEC, ntiles, slots, tiles, k, stats);
}

int hc_integrate_struct_host(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                             const HcFab* reset_src, const HcFab* ir, const HcBox* tiles, double a, double a_end, double dt,
                             int sdc_iter, const HcParams* prm, HcStats* stats) {
    if (ntiles <= 0 || !s_old || !diag || !s_new || !hydro_src || !reset_src || !ir || !tiles || !valid_params(prm) || !(a > 0.0) || !(a_end > 0.0)) {
        set_err("bad argument"); return HC_ERR_ARG;
    }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    if (stats) std::memset(stats, 0, sizeof *stats);
    const Consts k = make_consts_struct(rates_sp->data(), *prm, a, a_end, dt, sdc_iter);
    const bool src = (sdc_iter >= 0);   // with sdc_iter < 0 the update goes to S_old (f_rhs_struct.H:430-444 mirrored in store_cell)
    std::vector<HostSlot> slots = {
        {s_old, src ? std::vector<int>{DENS, EINT} : std::vector<int>{DENS, EDEN, EINT}, src ? std::vector<int>{} : std::vector<int>{EDEN, EINT}},   // with sources S_old is read-only and its (rho E) is never read
        {diag, prm->inhomo_reion ? std::vector<int>{TEMP, NE, ZHI} : std::vector<int>{TEMP, NE}, {TEMP, NE}},
        {s_new, {DENS, EDEN, EINT}, src ? std::vector<int>{EDEN, EINT} : std::vector<int>{}},
        {hydro_src, {DENS, EINT}, {}},
        {reset_src, {0}, {}},
        {ir, {}, src ? std::vector<int>{0} : std::vector<int>{}}};   // pure output: only the tile's cells travel back (copy_tile_d2h), host ghost cells keep their values
    return run_host(PATH_STRUCT, ntiles, slots, tiles, k, stats);
}

int hc_integrate_struct_react_host(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                                   const HcFab* reset_src, const HcFab* ir, const HcFab* react_in, const HcFab* react_out,
                                   const HcFab* react_out_work, const HcBox* tiles, double a, double a_end, double dt, int sdc_iter,
                                   const HcParams* prm, HcStats* stats) {
    if (ntiles <= 0 || !s_old || !diag || !s_new || !hydro_src || !reset_src || !ir || !react_in || !react_out || !react_out_work || !tiles ||
        !valid_params(prm) || !(a > 0.0) || !(a_end > 0.0)) {
        set_err("bad argument"); return HC_ERR_ARG;
    }
    if (sdc_iter < 0) { set_err("the SAVE_REACT dumps need the SDC sources (sdc_iter >= 0)"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    if (stats) std::memset(stats, 0, sizeof *stats);
    const Consts k = make_consts_struct(rates_sp->data(), *prm, a, a_end, dt, sdc_iter);
    std::vector<HostSlot> slots = {
        {s_old, {DENS, EINT}, {}},
        {diag, prm->inhomo_reion ? std::vector<int>{TEMP, NE, ZHI} : std::vector<int>{TEMP, NE}, {TEMP, NE}},
        {s_new, {DENS, EDEN, EINT}, {EDEN, EINT}},
        {hydro_src, {DENS, EINT}, {}},
        {reset_src, {0}, {}},
        {ir, {}, {0}},
        {react_in, {}, {0, 1, 2, 3, 4, 5, 6}},
        {react_out, {}, {0, 1, 2, 3, 4, 5, 6}},
        {react_out_work, {}, {0, 1, 2, 3, 4, 5, 6, 7, 8}}};
    return run_host(PATH_STRUCT, ntiles, slots, tiles, k, stats, nullptr, nullptr, true);
}

int hc_compute_new_temp_host(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, const HcParams* prm,
                             double small_temp, double large_temp, int max_temp_dt, HcStats* stats) {
    if (ntiles <= 0 || !state || !diag || !tiles || !valid_params(prm) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    if (stats) std::memset(stats, 0, sizeof *stats);
    const Consts k = make_consts_eos(rates_sp->data(), *prm, a);
    EosOpts eos; eos.mode = 1; eos.small_temp = small_temp; eos.large_temp = large_temp; eos.max_temp_dt = max_temp_dt;
    std::vector<HostSlot> slots = {{state, {0, 1, 2, 3, EDEN, EINT}, {EDEN, EINT}}, {diag, {TEMP, NE}, {TEMP, NE}}};
    return run_host(PATH_EOS, ntiles, slots, tiles, k, stats, &eos);
}

int hc_reset_internal_energy_host(int ntiles, const HcFab* state, const HcFab* diag, const HcFab* reset_src, const HcBox* tiles, double a,
                                  const HcParams* prm, double small_temp, int interp) {
    if (ntiles <= 0 || !state || !diag || !reset_src || !tiles || !valid_params(prm) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    const auto rates_sp = rates_snapshot();
    if (!rates_sp) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_eos(rates_sp->data(), *prm, a);
    EosOpts eos; eos.small_temp = small_temp; eos.interp = interp;
    std::vector<HostSlot> slots = {{state, {0, 1, 2, 3, EDEN, EINT}, {EDEN, EINT}}, {diag, {NE}, {}}, {reset_src, {0}, {0}}};
    return run_host(PATH_RESET_E, ntiles, slots, tiles, k, nullptr, &eos);
}

void hc_default_src_params(HcSrcParams* p) {
    if (!p) return;
    p->small_dens = -1.e200; p->small_temp = -1.e200;   // Source/Driver/Nyx.cpp:98-99
    p->gamma_minus_1 = 5.0 / 3.0 - 1.0; p->h_species = 0.76;
    p->min_density_type = HC_MIN_DENSITY_FLOOR; p->sdc = 1;
}

#define HC_SRC_CHECK() \
    if (ntiles < 0 || (ntiles > 0 && (!s_old || !s_new || !ext_src_old || !hydro_src || !grav || !tiles)) || !valid_src_params(prm) || !(a_old > 0.0) || \
        !(a_new > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; } \
    if (prm->min_density_type != HC_MIN_DENSITY_FLOOR && prm->min_density_type != HC_MIN_DENSITY_CONSERVATIVE) { \
        set_err("unknown min_density_type %d (Nyx: Don't know this enforce_min_density_type)", prm->min_density_type); return HC_ERR_ARG; } \
    if (int rc = check_src_fabs(ntiles, s_old, s_new, ext_src_old, hydro_src, grav)) return rc;

int hc_update_state_with_sources_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                       const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                       double* min_dens, void* stream) {
    HC_SRC_CHECK();
    const HcFab* fabs[5] = {s_old, s_new, ext_src_old, hydro_src, grav};
    // conservative variant: the source update alone; the caller iterates hc_enforce_min_density_cons_iter_* and ends with hc_finish_state_with_sources_*
    const int mode = (prm->min_density_type == HC_MIN_DENSITY_CONSERVATIVE) ? 3 : 0;
    return launch_sources(mode, ntiles, fabs, tiles, dt, a_old, a_new, *prm, min_dens, nullptr, (cudaStream_t)stream);
}

#define HC_FLOOR_ONLY() \
    if (prm->min_density_type != HC_MIN_DENSITY_FLOOR) { \
        set_err("the conservative variant is enforced by hc_enforce_min_density_cons_iter_* between the caller's FillPatch calls"); return HC_ERR_ARG; }

namespace {
int check_cons_fabs(int ntiles, const HcFab* sborder, const HcFab* s_new, const HcFab* reset_src, const HcBox* tiles, const HcSrcParams* prm) {
    if (ntiles < 0 || (ntiles > 0 && (!sborder || !s_new || !tiles)) || !valid_src_params(prm)) { set_err("bad argument"); return HC_ERR_ARG; }
    if (!(prm->small_dens > 0.0)) { set_err("the conservative variant needs small_dens > 0 (a cell below it must have nothing to give)"); return HC_ERR_ARG; }
    if (prm->sdc && ntiles > 0 && !reset_src) { set_err("SDC build: reset_e_src is written by every iteration"); return HC_ERR_ARG; }
    for (int t = 0; t < ntiles; ++t) {
        if (tiles[t].hi[0] < tiles[t].lo[0] || tiles[t].hi[1] < tiles[t].lo[1] || tiles[t].hi[2] < tiles[t].lo[2]) continue;
        if (!sborder[t].p || sborder[t].ncomp < 6 || !s_new[t].p || s_new[t].ncomp < 6 || (prm->sdc && (!reset_src[t].p || reset_src[t].ncomp < 1))) {
            set_err("tile %d: Sborder and S_new need the 6 state components, reset_e_src one", t); return HC_ERR_ARG;
        }
        for (int d = 0; d < 3; ++d)
            if (sborder[t].lo[d] > tiles[t].lo[d] - 2 || sborder[t].hi[d] < tiles[t].hi[d] + 2) {
                set_err("tile %d: Sborder must hold two filled ghost cells around the tile (Nyx_enforce_minimum_density.cpp:116-118)", t); return HC_ERR_ARG;
            }
    }
    return HC_OK;
}

// one iteration on device FABs; dmm (device, 16 bytes: [minimum key, initialised to all ones][bad faces, initialised to 0]) accumulates over calls
int launch_cons_iter(int ntiles, const HcFab* const* fabs, const HcBox* tiles, const HcSrcParams& p, unsigned long long* dmm, cudaStream_t stream) {
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; if (int rc = sm_count_of(dev, sms)) return rc;
    char* scratch; int n_used; long long ncells;
    if (int rc = stage_tiles(ntiles, fabs, p.sdc ? 3 : 2, tiles, stream, scratch, n_used, ncells)) return rc;
    StreamScratch owned(stream);
    owned.own(scratch);
    if (ncells == 0) return HC_OK;
    ConsArgs a{};
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = n_used; a.ncells = ncells; a.min_key = dmm; a.n_bad = dmm + 1; a.small_dens = p.small_dens; a.sdc = p.sdc;
    const int grid = (int)std::min<long long>((ncells + 255) / 256, (long long)sms * 8);
    hc_min_dens_cons_kernel<<<grid, 256, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return HC_OK;
}
int finish_cons(unsigned long long* dmm, cudaStream_t stream, double* min_after) {
    unsigned long long h[2] = {~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(h, dmm, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaFreeAsync(dmm, stream));
    if (min_after) *min_after = dens_from_key(h[0]);
    if (h[1]) { set_err("enforce_minimum_density (conservative): %llu faces with a negative coefficient (the reference aborts: mu < 0)", h[1]); return HC_ERR_ARG; }
    return HC_OK;
}
int new_cons_words(unsigned long long*& dmm, cudaStream_t stream) {
    CUDA_TRY(cudaMallocAsync((void**)&dmm, 16, stream));
    CUDA_TRY(cudaMemsetAsync(dmm, 0xff, 8, stream));
    CUDA_TRY(cudaMemsetAsync(dmm + 1, 0, 8, stream));
    return HC_OK;
}
}  // namespace

int hc_enforce_min_density_cons_iter_batch(int ntiles, const HcFab* sborder, const HcFab* s_new, const HcFab* reset_src, const HcBox* tiles,
                                           const HcSrcParams* prm, double* min_dens_after, void* stream_) {
    if (int rc = check_cons_fabs(ntiles, sborder, s_new, reset_src, tiles, prm)) return rc;
    if (min_dens_after) *min_dens_after = DBL_MAX;
    if (ntiles == 0) return HC_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long* dmm = nullptr;
    if (int rc = new_cons_words(dmm, stream)) return rc;
    const HcFab* fabs[3] = {sborder, s_new, reset_src};
    if (int rc = launch_cons_iter(ntiles, fabs, tiles, *prm, dmm, stream)) return rc;
    return finish_cons(dmm, stream, min_dens_after);
}

int hc_enforce_min_density_cons_iter_host(int ntiles, const HcFab* sborder, const HcFab* s_new, const HcFab* reset_src, const HcBox* tiles,
                                          const HcSrcParams* prm, double* min_dens_after) {
    if (int rc = check_cons_fabs(ntiles, sborder, s_new, reset_src, tiles, prm)) return rc;
    if (min_dens_after) *min_dens_after = DBL_MAX;
    if (ntiles == 0) return HC_OK;
    int dev; if (int rc = current_device(dev)) return rc;
    HostPipe* hp = nullptr;
    if (int rc = host_pipe(dev, hp)) return rc;
    unsigned long long* dmm = nullptr;
    if (int rc = new_cons_words(dmm, hp->comp)) return rc;
    const std::vector<int> all6 = {0, 1, 2, 3, 4, 5};
    const HcSrcParams p = *prm;
    std::vector<HostSlot> slots = {{sborder, all6, {}, 2}, {s_new, all6, all6}};   // the kernel reads the border copy two cells around a tile
    if (p.sdc) slots.push_back({reset_src, {}, {0}});
    GroupLauncher iter = [&](int n, const HcFab* const* fabs, const HcBox* tl, cudaStream_t st) { return launch_cons_iter(n, fabs, tl, p, dmm, st); };
    const int rc = run_host(-1, ntiles, slots, tiles, Consts{}, nullptr, nullptr, &iter);
    const int rc2 = finish_cons(dmm, hp->comp, min_dens_after);
    return rc != HC_OK ? rc : rc2;
}

// sweep (3) of update_state_with_sources on the S_new the conservative iterations left; density_enforced: they ran (SDC build: hydro_src(rho) is reset)
int hc_finish_state_with_sources_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                       const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                       int density_enforced, void* stream) {
    HC_SRC_CHECK();
    const HcFab* fabs[5] = {s_old, s_new, ext_src_old, hydro_src, grav};
    return launch_sources(density_enforced ? 5 : 4, ntiles, fabs, tiles, dt, a_old, a_new, *prm, nullptr, nullptr, (cudaStream_t)stream);
}

int hc_finish_state_with_sources_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                      const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                      int density_enforced) {
    HC_SRC_CHECK();
    if (ntiles == 0) return HC_OK;
    const std::vector<int> all6 = {0, 1, 2, 3, 4, 5};
    const HcSrcParams p = *prm;
    const bool reset = density_enforced && p.sdc;
    std::vector<HostSlot> slots = {{s_old, {0, 1, 2, 3}, {}}, {s_new, all6, all6}, {ext_src_old, {}, {}},
                                   {hydro_src, reset ? std::vector<int>{0} : std::vector<int>{}, reset ? std::vector<int>{0} : std::vector<int>{}},
                                   {grav, {0, 1, 2}, {}}};
    GroupLauncher fin = [&](int n, const HcFab* const* fabs, const HcBox* tl, cudaStream_t st) {
        return launch_sources(density_enforced ? 5 : 4, n, fabs, tl, dt, a_old, a_new, p, nullptr, nullptr, st);
    };
    return run_host(-1, ntiles, slots, tiles, Consts{}, nullptr, nullptr, &fin);
}

int hc_enforce_minimum_density_batch(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                     const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                     void* stream) {
    HC_SRC_CHECK();
    HC_FLOOR_ONLY();
    const HcFab* fabs[5] = {s_old, s_new, ext_src_old, hydro_src, grav};
    return launch_sources(2, ntiles, fabs, tiles, dt, a_old, a_new, *prm, nullptr, nullptr, (cudaStream_t)stream);
}

int hc_enforce_minimum_density_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                    const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm) {
    HC_SRC_CHECK();
    HC_FLOOR_ONLY();
    if (ntiles == 0) return HC_OK;
    // every cell again from the untouched inputs with the floor in between; hydro_src(rho) goes back as well
    // (S_new travels in too: whole components travel back, and its ghost cells must keep their values)
    const std::vector<int> all6 = {0, 1, 2, 3, 4, 5};
    const HcSrcParams p = *prm;
    std::vector<HostSlot> slots2 = {{s_old, all6, {}}, {s_new, all6, all6}, {ext_src_old, all6, {}},
                                    {hydro_src, all6, p.sdc ? std::vector<int>{0} : std::vector<int>{}}, {grav, {0, 1, 2}, {}}};
    GroupLauncher pass2 = [&](int n, const HcFab* const* fabs, const HcBox* tl, cudaStream_t st) {
        return launch_sources(2, n, fabs, tl, dt, a_old, a_new, p, nullptr, nullptr, st);
    };
    return run_host(-1, ntiles, slots2, tiles, Consts{}, nullptr, nullptr, &pass2);
}

int hc_update_state_with_sources_host(int ntiles, const HcFab* s_old, const HcFab* s_new, const HcFab* ext_src_old, const HcFab* hydro_src,
                                      const HcFab* grav, const HcBox* tiles, double dt, double a_old, double a_new, const HcSrcParams* prm,
                                      double* min_dens) {
    HC_SRC_CHECK();
    if (min_dens) *min_dens = DBL_MAX;
    if (ntiles == 0) return HC_OK;
    int dev; if (int rc = current_device(dev)) return rc;
    HostPipe* hp = nullptr;
    if (int rc = host_pipe(dev, hp)) return rc;
    unsigned long long* dmin = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&dmin, 8, hp->comp));
    CUDA_TRY(cudaMemsetAsync(dmin, 0xff, 8, hp->comp));
    const std::vector<int> all6 = {0, 1, 2, 3, 4, 5};
    // pass 1: sweeps (1)+(3) per group of tiles, the minimum accumulated over the groups on the device
    std::vector<HostSlot> slots = {{s_old, all6, {}}, {s_new, all6, all6}, {ext_src_old, all6, {}}, {hydro_src, all6, {}}, {grav, {0, 1, 2}, {}}};
    // (S_new travels in as well: whole components travel back, and its ghost cells must keep their values)
    const HcSrcParams p = *prm;
    GroupLauncher pass1 = [&](int n, const HcFab* const* fabs, const HcBox* tl, cudaStream_t st) {
        return launch_sources(p.min_density_type == HC_MIN_DENSITY_CONSERVATIVE ? 3 : 1, n, fabs, tl, dt, a_old, a_new, p, nullptr, dmin, st);
    };
    int rc = run_host(-1, ntiles, slots, tiles, Consts{}, nullptr, nullptr, &pass1);
    unsigned long long key = ~0ull;
    if (rc == HC_OK) {
        CUDA_TRY(cudaMemcpyAsync(&key, dmin, 8, cudaMemcpyDeviceToHost, hp->comp));
        CUDA_TRY(cudaStreamSynchronize(hp->comp));
    }
    CUDA_TRY(cudaFreeAsync(dmin, hp->comp));
    if (rc != HC_OK) return rc;
    const double m = dens_from_key(key);
    if (min_dens) *min_dens = m;
    if (m < p.small_dens && p.min_density_type == HC_MIN_DENSITY_FLOOR)   // pass 2 (rare)
        rc = hc_enforce_minimum_density_host(ntiles, s_old, s_new, ext_src_old, hydro_src, grav, tiles, dt, a_old, a_new, prm);
    return rc;
}

int hc_init_zhi_batch(int ntiles, const HcFab* diag, const HcFab* zhi, int ratio, const HcBox* tiles, void* stream_) {
    if (ntiles < 0 || (ntiles > 0 && (!diag || !zhi || !tiles)) || ratio < 1) { set_err("bad argument"); return HC_ERR_ARG; }
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; if (int rc = sm_count_of(dev, sms)) return rc;
    for (int t = 0; t < ntiles; ++t) {
        if (diag[t].ncomp <= ZHI) { set_err("diag FAB %d has no Zhi component (nyx.inhomo_reion = 1 allocates 3 components)", t); return HC_ERR_ARG; }
        if (!zhi[t].p) { set_err("null zhi FAB %d", t); return HC_ERR_ARG; }
        for (int d = 0; d < 3; ++d) {   // the coarse FAB must cover the tile coarsened by ratio
            if (tiles[t].hi[d] < tiles[t].lo[d]) continue;
            if (tiles[t].lo[d] / ratio < zhi[t].lo[d] || tiles[t].hi[d] / ratio > zhi[t].hi[d]) { set_err("zhi FAB %d does not cover tile/ratio", t); return HC_ERR_ARG; }
        }
    }
    const HcFab* fabs[2] = {diag, zhi};
    char* scratch; int n_used; long long ncells;
    if (int rc = stage_tiles(ntiles, fabs, 2, tiles, stream, scratch, n_used, ncells, 1)) return rc;   // containment of the fine tile: diag only
    StreamScratch owned(stream);
    owned.own(scratch);
    if (ncells == 0) return HC_OK;
    ZhiArgs a{};
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = n_used; a.ncells = ncells; a.ratio = ratio; a.zcomp = ZHI;
    const int grid = (int)std::min<long long>((ncells + 255) / 256, (long long)sms * 16);
    hc_init_zhi_kernel<<<grid, 256, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    return HC_OK;
}

int hc_init_zhi_host(int ntiles, const HcFab* diag, const HcFab* zhi, int ratio, const HcBox* tiles) {
    if (ntiles < 0 || (ntiles > 0 && (!diag || !zhi || !tiles)) || ratio < 1) { set_err("bad argument"); return HC_ERR_ARG; }
    if (ntiles == 0) return HC_OK;
    // diag(Zhi) is a pure output (only the tile's cells travel back), the coarse zhi FAB a pure input
    std::vector<HostSlot> slots = {{diag, {}, {ZHI}}, {zhi, {0}, {}, 0, true}};   // zhi lives on the coarse index space: whole
    GroupLauncher fill = [&](int n, const HcFab* const* fabs, const HcBox* tl, cudaStream_t st) {
        return hc_init_zhi_batch(n, fabs[0], fabs[1], ratio, tl, st);
    };
    return run_host(-1, ntiles, slots, tiles, Consts{}, nullptr, nullptr, &fill);
}

int hc_fab_copy_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream) {
    if (ntiles < 0 || (ntiles > 0 && (!dst || !src || !tiles))) { set_err("bad argument"); return HC_ERR_ARG; }
    return launch_fab_op(0, ntiles, dst, dcomp, src, scomp, ncomp, tiles, (cudaStream_t)stream);
}
int hc_fab_add_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream) {
    if (ntiles < 0 || (ntiles > 0 && (!dst || !src || !tiles))) { set_err("bad argument"); return HC_ERR_ARG; }
    return launch_fab_op(1, ntiles, dst, dcomp, src, scomp, ncomp, tiles, (cudaStream_t)stream);
}
int hc_fab_subtract_batch(int ntiles, const HcFab* dst, int dcomp, const HcFab* src, int scomp, int ncomp, const HcBox* tiles, void* stream) {
    if (ntiles < 0 || (ntiles > 0 && (!dst || !src || !tiles))) { set_err("bad argument"); return HC_ERR_ARG; }
    return launch_fab_op(2, ntiles, dst, dcomp, src, scomp, ncomp, tiles, (cudaStream_t)stream);
}

int hc_measure_fp64_peak(double* flops_per_s) {
    if (!flops_per_s) { set_err("null argument"); return HC_ERR_ARG; }
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; if (int rc = sm_count_of(dev, sms)) return rc;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1; CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        hc_dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3);
        if (rep > 0 && flops > best) best = flops;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *flops_per_s = best;
    return HC_OK;
}

#if defined(HC_PHASE_TIMING)
// diagnostics build only (tools/build_variants.sh): read and reset the per-phase cycle totals of the sorted kernel
int hc_debug_phase(unsigned long long* out48) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out48, sorted::g_phase, 48 * sizeof(unsigned long long)));
    unsigned long long zero[48] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(sorted::g_phase, zero, sizeof zero));
    return HC_OK;
}
int hc_debug_mix(unsigned long long* out256) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out256, sorted::g_mix, 256 * sizeof(unsigned long long)));
    static unsigned long long zero[256] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(sorted::g_mix, zero, sizeof zero));
    return HC_OK;
}
int hc_debug_stage(unsigned long long* out16) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out16, g_stage, 16 * sizeof(unsigned long long)));
    unsigned long long zero[16] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(g_stage, zero, sizeof zero));
    return HC_OK;
}
#endif

int hc_selftest_log10(const double* x, double* y, int* bad, long long n) {
    if (!x || !y || !bad || n < 0) { set_err("bad argument"); return HC_ERR_ARG; }
    int dev; if (int rc = current_device(dev)) return rc;
    if (!g_dev[dev].logtab) { set_err("hc_tables_upload has not been called on device %d", dev); return HC_ERR_NO_TABLES; }
    if (n == 0) return HC_OK;
    double *dx = nullptr, *dy = nullptr; int* db = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dx, n * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&dy, n * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&db, n * sizeof(int)));
    CUDA_TRY(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
    hc_log10_selftest_kernel<<<148, 256>>>(g_dev[dev].logtab, dx, dy, db, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(bad, db, n * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(db);
    return HC_OK;
}

int hc_selftest_div_delta_t(const double* x, double* y, long long n) {
    if (!x || !y || n < 0) { set_err("bad argument"); return HC_ERR_ARG; }
    int dev; if (int rc = current_device(dev)) return rc;
    if (n == 0) return HC_OK;
    double *dx = nullptr, *dy = nullptr;
    CUDA_TRY(cudaMalloc((void**)&dx, n * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&dy, n * sizeof(double)));
    CUDA_TRY(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
    hc_divdelta_selftest_kernel<<<148, 256>>>(dx, dy, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy);
    return HC_OK;
}

int hc_last_launch_timing(double* kernel_ms, double* drain_ms) {
    if (kernel_ms) *kernel_ms = g_last_kernel_ms;
    if (drain_ms) *drain_ms = g_last_drain_ms;
    return HC_OK;
}

int hc_sync(void* stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return HC_OK;
}

}  // extern "C"

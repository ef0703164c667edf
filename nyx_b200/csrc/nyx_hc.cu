// nyx_hc.cu -- sm_100a kernels and the C-ABI of include/nyx_hc.h.
//
// Kernel design (DESIGN.md has the numbers):
//   * one persistent CTA per SM, 227 KB-class shared memory: the 7 ionization-rate tables (read ~6x per RHS
//     evaluation) are staged interleaved in shared memory, one 64-byte row per temperature index, so one lookup
//     is two adjacent rows = 128 contiguous bytes; the 8 cooling tables (read once per RHS) stay in L1/L2 (__ldg);
//     the UV-background row is interpolated once per call on the host (z is uniform) and travels as kernel constants;
//   * one thread (lane) per cell in flight, CVODE-equivalent BDF state in registers (hc_device.cuh);
//   * a global work queue of cells: a lane that finishes its cell immediately pulls the next one
//     (warp-aggregated atomicAdd on ballot of free lanes), and the integrator is a resumable state machine so that
//     the 32 lanes of a warp evaluate their right-hand sides together whatever BDF phase each is in;
//   * per-cell outputs are scattered straight to the FABs; diagnostics are reduced in shared memory, then one
//     atomicAdd per CTA per counter.
// No AMReX, no SUNDIALS, no library kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <vector>

#include "hc_host.hpp"

namespace {

using namespace hc;

constexpr int THREADS = 512;               // 16 warps per SM
constexpr int ION_ROWS = NTAB + 1;         // one padding row: row j+1 always exists
constexpr size_t ION_SMEM_BYTES = (size_t)ION_ROWS * TABLE_ROW * sizeof(double);   // 128,128 B

enum Comp { DENS = 0, EDEN = 4, EINT = 5, TEMP = 0, NE = 1, ZHI = 2 };
enum FabSlot { F_STATE = 0, F_DIAG = 1, F_SNEW = 2, F_HSRC = 3, F_RSRC = 4, F_IR = 5 };

struct TileDesc {
    HcFab f[6];
    int lo[3];
    int nx, ny, nz;
    long long offset;   // global index of this tile's first cell
};

struct KernelArgs {
    Consts k;
    const TileDesc* tiles;
    int ntiles;
    long long ncells;
    unsigned long long* queue;
    unsigned long long* dstats;   // HcStats as 14 x u64
    HcCellStat* cell_stats;
    const double* ion;
    const double* cool;
};

enum StatSlot { S_CELLS = 0, S_FAILED, S_FLOOR, S_NST, S_MAXNST, S_NFE, S_NFELS, S_NETF, S_NNI, S_NCFN, S_NSETUPS, S_NEITERS, S_ATTEMPTS, S_EOS, S_COUNT };
static_assert(S_COUNT * sizeof(long long) == sizeof(HcStats), "HcStats layout");

__device__ __forceinline__ double& fab_at(const HcFab& f, int i, int j, int k, int n) {
    return f.p[(i - f.lo[0]) + (long long)(j - f.lo[1]) * f.jstride + (long long)(k - f.lo[2]) * f.kstride + (long long)n * f.nstride];
}

__device__ __forceinline__ const TileDesc& find_tile(const TileDesc* tiles, int ntiles, long long id) {
    int lo = 0, hi = ntiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tiles[mid].offset <= id) lo = mid; else hi = mid - 1;
    }
    return tiles[lo];
}

__device__ __forceinline__ void cell_ijk(const TileDesc& t, long long id, int& i, int& j, int& k) {
    const long long loc = id - t.offset;
    const int plane = t.nx * t.ny;
    const int kk = (int)(loc / plane);
    const int rem = (int)(loc - (long long)kk * plane);
    const int jj = rem / t.nx;
    i = t.lo[0] + (rem - jj * t.nx); j = t.lo[1] + jj; k = t.lo[2] + kk;
}

// gather one cell into a lane (HOT LOOP A of the reference: integrate_state_vec_3d.cpp:227-233,
// ode_eos_initialize_arrays f_rhs_struct.H:180-209) and start its integration
template <int PATH>
__device__ __forceinline__ void load_cell(Lane<PATH>& ln, const KernelArgs& a, long long id) {
    const TileDesc& t = find_tile(a.tiles, a.ntiles, id);
    int i, j, k; cell_ijk(t, id, i, j, k);
    const Consts& c = a.k;
    ln.rho = fab_at(t.f[F_STATE], i, j, k, DENS);
    const double rhoe0 = fab_at(t.f[F_STATE], i, j, k, EINT);
    ln.e0 = rhoe0 / ln.rho;
    ln.abstol = nv_scale(c.atol_factor, ln.e0);
    ln.lastT = fab_at(t.f[F_DIAG], i, j, k, TEMP);
    ln.lastNe = fab_at(t.f[F_DIAG], i, j, k, NE);
    ln.lastNh = 1.0;
    ln.jh = (double)c.JH0;
    if (PATH == PATH_STRUCT) {
        ln.rho_src = ln.rhoe_src = ln.e_src = ln.reset_src = 0.0; ln.zhi = 0.0;
        if (c.sdc_has_src) {
            ln.rho_src = fab_at(t.f[F_HSRC], i, j, k, DENS) / c.dt;
            ln.rhoe_src = fab_at(t.f[F_HSRC], i, j, k, EINT) / c.dt;
            ln.reset_src = fab_at(t.f[F_RSRC], i, j, k, 0);
            ln.e_src = (((c.asq * rhoe0 + c.dt * ln.rhoe_src) / c.aendsq + ln.reset_src) / (ln.rho + c.dt * ln.rho_src) - ln.e0) / c.dt;
        }
        if (c.inhomo) { ln.zhi = fab_at(t.f[F_DIAG], i, j, k, ZHI); ln.jh = (c.z > ln.zhi) ? 0.0 : 1.0; }
        ln.rho_out = fab_at(t.f[F_SNEW], i, j, k, DENS);
        ln.rhoe_new = fab_at(t.f[F_SNEW], i, j, k, EINT);
    }
    ln.start(c);
}

// scatter a finished cell (HOT LOOP C: integrate_state_vec_3d.cpp:317-321, f_rhs_struct.H:290-291,438-444)
template <int PATH>
__device__ __forceinline__ void store_cell(const Lane<PATH>& ln, const KernelArgs& a, long long id, unsigned long long* sstats) {
    const TileDesc& t = find_tile(a.tiles, a.ntiles, id);
    int i, j, k; cell_ijk(t, id, i, j, k);
    const Consts& c = a.k;
    fab_at(t.f[F_DIAG], i, j, k, TEMP) = ln.outT;
    fab_at(t.f[F_DIAG], i, j, k, NE) = ln.outNe;
    if (PATH == PATH_VEC || !c.sdc_has_src) {
        const double d = ln.rho * (ln.e_final - ln.e0);
        fab_at(t.f[F_STATE], i, j, k, EINT) += d;
        fab_at(t.f[F_STATE], i, j, k, EDEN) += d;
    } else {
        fab_at(t.f[F_IR], i, j, k, 0) = ln.IR;
        const double d = c.dt * c.ahalf * ln.IR / c.aendsq;
        fab_at(t.f[F_SNEW], i, j, k, EINT) = ln.rhoe_new + d;
        double& eden = fab_at(t.f[F_SNEW], i, j, k, EDEN);
        eden = eden + d;
    }
    if (a.cell_stats) a.cell_stats[id] = HcCellStat{ln.nst, ln.netf, ln.nfe, ln.nni, ln.nnf, ln.nsetups, ln.nfe_ls, ln.flag};
    atomicAdd(&sstats[S_CELLS], 1ull);
    if (ln.flag < 0) atomicAdd(&sstats[S_FAILED], 1ull);
    if (ln.floor_hit) atomicAdd(&sstats[S_FLOOR], 1ull);
    atomicAdd(&sstats[S_NST], (unsigned long long)ln.nst);
    atomicMax(&sstats[S_MAXNST], (unsigned long long)ln.nst);
    atomicAdd(&sstats[S_NFE], (unsigned long long)ln.nfe);
    atomicAdd(&sstats[S_NFELS], (unsigned long long)ln.nfe_ls);
    atomicAdd(&sstats[S_NETF], (unsigned long long)ln.netf);
    atomicAdd(&sstats[S_NNI], (unsigned long long)ln.nni);
    atomicAdd(&sstats[S_NCFN], (unsigned long long)ln.nnf);
    atomicAdd(&sstats[S_NSETUPS], (unsigned long long)ln.nsetups);
    atomicAdd(&sstats[S_NEITERS], (unsigned long long)ln.ne_iters);
    atomicAdd(&sstats[S_ATTEMPTS], (unsigned long long)ln.attempts);
    atomicAdd(&sstats[S_EOS], (unsigned long long)ln.n_eos);
}

template <int PATH>
__global__ void __launch_bounds__(THREADS, 1) hc_integrate_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(16) double s_ion[];
    __shared__ unsigned long long s_stats[S_COUNT];

    // stage the ionization tables (16-byte vector copies)
    {
        const double2* src = reinterpret_cast<const double2*>(a.ion);
        double2* dst = reinterpret_cast<double2*>(s_ion);
        for (int i = threadIdx.x; i < ION_ROWS * TABLE_ROW / 2; i += THREADS) dst[i] = __ldg(src + i);
        if (threadIdx.x < S_COUNT) s_stats[threadIdx.x] = 0ull;
    }
    __syncthreads();

    const Tables tb{s_ion, a.cool};
    const unsigned lane_id = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane_id) - 1u;

    Lane<PATH> ln;
    ln.pc = PC_IDLE;
    long long cell_id = -1;
    bool queue_empty = false;

    for (;;) {
        // ---- refill free lanes from the global queue (warp-aggregated)
        if (!queue_empty) {
            const bool need = !ln.active();
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                const int cnt = __popc(m);
                unsigned long long base = 0;
                if (lane_id == 0) base = atomicAdd(a.queue, (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + cnt >= (unsigned long long)a.ncells) queue_empty = true;
                if (need) {
                    const long long id = (long long)(base + __popc(m & lt_mask));
                    if (id < a.ncells) { cell_id = id; load_cell<PATH>(ln, a, id); }
                }
            }
        }
        // a cell can complete inside start() only through the early-failure path, which still requests an EOS solve
        if (!__any_sync(0xffffffffu, ln.active())) break;

        // ---- all lanes evaluate their pending request together
        double f = 0.0;
        if (ln.active()) f = ln.eval_request(tb, a.k);
        __syncwarp();
        // ---- integrator bookkeeping until the next request (cheap, divergent)
        if (ln.active()) {
            ln.resume(a.k, f);
            if (!ln.active()) store_cell<PATH>(ln, a, cell_id, s_stats);
        }
        __syncwarp();
    }

    __syncthreads();
    if (threadIdx.x < S_COUNT) {
        if (threadIdx.x == S_MAXNST) atomicMax(&a.dstats[threadIdx.x], s_stats[threadIdx.x]);
        else atomicAdd(&a.dstats[threadIdx.x], s_stats[threadIdx.x]);
    }
}

// compute_new_temp core: one thread per cell, grid-stride; same table staging
__global__ void __launch_bounds__(THREADS, 1) hc_eos_kernel(const __grid_constant__ KernelArgs a) {
    extern __shared__ __align__(16) double s_ion[];
    __shared__ unsigned long long s_stats[S_COUNT];
    {
        const double2* src = reinterpret_cast<const double2*>(a.ion);
        double2* dst = reinterpret_cast<double2*>(s_ion);
        for (int i = threadIdx.x; i < ION_ROWS * TABLE_ROW / 2; i += THREADS) dst[i] = __ldg(src + i);
        if (threadIdx.x < S_COUNT) s_stats[threadIdx.x] = 0ull;
    }
    __syncthreads();
    const Tables tb{s_ion, a.cool};
    const Consts& c = a.k;
    unsigned long long iters = 0, cells = 0;
    for (long long id = (long long)blockIdx.x * THREADS + threadIdx.x; id < a.ncells; id += (long long)gridDim.x * THREADS) {
        const TileDesc& t = find_tile(a.tiles, a.ntiles, id);
        int i, j, k; cell_ijk(t, id, i, j, k);
        const double R = fab_at(t.f[F_STATE], i, j, k, DENS);
        const double e = fab_at(t.f[F_STATE], i, j, k, EINT) / R;
        const double rho_cgs = R * density_to_cgs / c.a3_eos;
        const double U = e * e_to_cgs;
        const double nh = rho_cgs * c.h_species / MPROTON;
        EosOut s;
        iterate_ne(tb, c, c.uvb_eos, 1.0, 1.0, U, nh, s);
        fab_at(t.f[F_DIAG], i, j, k, TEMP) = s.T;
        fab_at(t.f[F_DIAG], i, j, k, NE) = s.ne;
        iters += s.iters; cells++;
    }
    atomicAdd(&s_stats[S_CELLS], cells);
    atomicAdd(&s_stats[S_EOS], cells);
    atomicAdd(&s_stats[S_NEITERS], iters);
    __syncthreads();
    if (threadIdx.x < S_COUNT) atomicAdd(&a.dstats[threadIdx.x], s_stats[threadIdx.x]);
}

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

// ---------------------------------------------------------------------------------------------- host state
thread_local char g_err[512] = "";
void set_err(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_err("%s: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); return HC_ERR_CUDA; } } while (0)

struct DeviceTables {
    double* ion = nullptr;
    double* cool = nullptr;
    int sm_count = 0;
    bool attr_set[3] = {false, false, false};
};
std::mutex g_mu;
std::vector<double> g_rates;          // host copy of the rates image
DeviceTables g_dev[64];               // indexed by CUDA device ordinal

int current_device(int& dev) {
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_err("unsupported device ordinal %d", dev); return HC_ERR_CUDA; }
    return HC_OK;
}

bool valid_params(const HcParams* p) {
    return p && p->rtol > 0.0 && p->atol_factor >= 0.0 && p->h_species > 0.0 && p->h_species <= 1.0;
}

TileDesc make_tile(const HcFab* const* fabs, int nf, int idx, const HcBox& b, long long offset) {
    TileDesc t{};
    for (int s = 0; s < nf; ++s) t.f[s] = fabs[s][idx];
    for (int d = 0; d < 3; ++d) t.lo[d] = b.lo[d];
    t.nx = b.hi[0] - b.lo[0] + 1; t.ny = b.hi[1] - b.lo[1] + 1; t.nz = b.hi[2] - b.lo[2] + 1;
    t.offset = offset;
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
int set_smem_attr(KernelT kernel, DeviceTables& dt, int slot) {
    if (!dt.attr_set[slot]) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ION_SMEM_BYTES));
        dt.attr_set[slot] = true;
    }
    return HC_OK;
}

// Common launcher: build tile descriptors, stream-ordered scratch, launch, optionally read the statistics back.
int launch(int path, int ntiles, const HcFab* const* fabs, int nf, const HcBox* tiles, const Consts& k, HcStats* stats,
           HcCellStat* cell_stats, cudaStream_t stream) {
    int dev; if (int rc = current_device(dev)) return rc;
    DeviceTables& dt = g_dev[dev];
    if (!dt.ion) { set_err("hc_tables_upload has not been called on device %d", dev); return HC_ERR_NO_TABLES; }
    if (ntiles < 0) { set_err("ntiles < 0"); return HC_ERR_ARG; }
    std::vector<TileDesc> h_tiles; h_tiles.reserve(ntiles);
    long long ncells = 0;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc td = make_tile(fabs, nf, t, tiles[t], ncells);
        if (td.nx <= 0 || td.ny <= 0 || td.nz <= 0) continue;   // empty tile: nothing to do (as an empty MFIter tile)
        if (!tile_inside(td, nf)) { set_err("tile %d is not contained in its FABs (or a FAB pointer is null)", t); return HC_ERR_ARG; }
        ncells += (long long)td.nx * td.ny * td.nz;
        h_tiles.push_back(td);
    }
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (ncells == 0) return HC_OK;

    const size_t tiles_bytes = h_tiles.size() * sizeof(TileDesc);
    const size_t scratch_bytes = 256 + tiles_bytes;   // [queue u64][pad][stats 14 x u64][pad] [tiles]
    char* scratch = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&scratch, scratch_bytes, stream));
    CUDA_TRY(cudaMemsetAsync(scratch, 0, 256, stream));
    CUDA_TRY(cudaMemcpyAsync(scratch + 256, h_tiles.data(), tiles_bytes, cudaMemcpyHostToDevice, stream));

    KernelArgs a{};
    a.k = k;
    a.tiles = reinterpret_cast<const TileDesc*>(scratch + 256);
    a.ntiles = (int)h_tiles.size();
    a.ncells = ncells;
    a.queue = reinterpret_cast<unsigned long long*>(scratch);
    a.dstats = reinterpret_cast<unsigned long long*>(scratch + 64);
    a.cell_stats = cell_stats;
    a.ion = dt.ion; a.cool = dt.cool;

    const long long want = (ncells + THREADS - 1) / THREADS;
    const int grid = (int)std::min<long long>(want, dt.sm_count);
    if (path == PATH_VEC) {
        if (int rc = set_smem_attr(hc_integrate_kernel<PATH_VEC>, dt, 0)) return rc;
        hc_integrate_kernel<PATH_VEC><<<grid, THREADS, ION_SMEM_BYTES, stream>>>(a);
    } else if (path == PATH_STRUCT) {
        if (int rc = set_smem_attr(hc_integrate_kernel<PATH_STRUCT>, dt, 1)) return rc;
        hc_integrate_kernel<PATH_STRUCT><<<grid, THREADS, ION_SMEM_BYTES, stream>>>(a);
    } else {
        if (int rc = set_smem_attr(hc_eos_kernel, dt, 2)) return rc;
        hc_eos_kernel<<<grid, THREADS, ION_SMEM_BYTES, stream>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    if (stats) {
        // the pageable `stats` target makes this copy synchronous with respect to the host
        CUDA_TRY(cudaMemcpyAsync(stats, scratch + 64, sizeof(HcStats), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    CUDA_TRY(cudaFreeAsync(scratch, stream));
    return HC_OK;
}

// ---- host-buffer staging -----------------------------------------------------------------------------------
struct Staged {
    std::vector<HcFab> dev;       // device-side fabs (same geometry)
    std::vector<double*> bufs;
};
size_t fab_doubles(const HcFab& f) { return (size_t)f.nstride * f.ncomp; }

int stage_in(int n, const HcFab* host, const std::vector<int>& comps, Staged& st, cudaStream_t stream) {
    st.dev.assign(host, host + n);
    st.bufs.assign(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (!host[i].p) { set_err("null host FAB"); return HC_ERR_ARG; }
        double* d = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&d, fab_doubles(host[i]) * sizeof(double), stream));
        st.bufs[i] = d; st.dev[i].p = d;
        for (int c : comps) {
            if (c >= host[i].ncomp) continue;
            CUDA_TRY(cudaMemcpyAsync(d + (size_t)c * host[i].nstride, host[i].p + (size_t)c * host[i].nstride,
                                     (size_t)host[i].nstride * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
    }
    return HC_OK;
}
int stage_out(int n, const HcFab* host, const std::vector<int>& comps, Staged& st, cudaStream_t stream) {
    for (int i = 0; i < n; ++i)
        for (int c : comps) {
            if (c >= host[i].ncomp) continue;
            CUDA_TRY(cudaMemcpyAsync(host[i].p + (size_t)c * host[i].nstride, st.bufs[i] + (size_t)c * host[i].nstride,
                                     (size_t)host[i].nstride * sizeof(double), cudaMemcpyDeviceToHost, stream));
        }
    return HC_OK;
}
int stage_free(Staged& st, cudaStream_t stream) {
    for (double* d : st.bufs) if (d) CUDA_TRY(cudaFreeAsync(d, stream));
    st.bufs.clear();
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
    g_rates.assign(rates, rates + n_doubles);
    std::vector<double> ion, cool;
    interleave_tables(rates, ion, cool);
    DeviceTables& dt = g_dev[dev];
    if (!dt.ion) {
        CUDA_TRY(cudaMalloc((void**)&dt.ion, ion.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc((void**)&dt.cool, cool.size() * sizeof(double)));
        CUDA_TRY(cudaDeviceGetAttribute(&dt.sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    CUDA_TRY(cudaMemcpy(dt.ion, ion.data(), ion.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dt.cool, cool.data(), cool.size() * sizeof(double), cudaMemcpyHostToDevice));
    return HC_OK;
}

int hc_uvb_at_z(double z, double* out6) {
    if (g_rates.empty()) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Uvb u = uvb_at_z(g_rates.data(), z);
    out6[0] = u.ggh0; out6[1] = u.gghe0; out6[2] = u.gghep; out6[3] = u.eh0; out6[4] = u.ehe0; out6[5] = u.ehep;
    return HC_OK;
}

int hc_integrate_vec_batch(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                           const HcParams* prm, HcStats* stats, HcCellStat* cell_stats, void* stream) {
    if (!valid_params(prm) || (ntiles > 0 && (!state || !diag || !tiles)) || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    if (g_rates.empty()) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_vec(g_rates.data(), *prm, a, dt);
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
    if (g_rates.empty()) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_struct(g_rates.data(), *prm, a, a_end, dt, sdc_iter);
    const HcFab* fabs[6] = {s_old, diag, s_new, hydro_src, reset_src, ir};
    return launch(PATH_STRUCT, ntiles, fabs, 6, tiles, k, stats, cell_stats, (cudaStream_t)stream);
}

int hc_integrate_struct(const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src, const HcFab* reset_src,
                        const HcFab* ir, HcBox tile, double a, double a_end, double dt, int sdc_iter, const HcParams* prm,
                        HcStats* stats, HcCellStat* cell_stats, void* stream) {
    return hc_integrate_struct_batch(1, s_old, diag, s_new, hydro_src, reset_src, ir, &tile, a, a_end, dt, sdc_iter, prm, stats,
                                     cell_stats, stream);
}

int hc_eos_T_given_Re(const HcFab* state, const HcFab* diag, HcBox tile, double a, const HcParams* prm, HcStats* stats, void* stream) {
    if (!valid_params(prm) || !state || !diag || !(a > 0.0)) { set_err("bad argument"); return HC_ERR_ARG; }
    if (g_rates.empty()) { set_err("hc_tables_upload has not been called"); return HC_ERR_NO_TABLES; }
    const Consts k = make_consts_eos(g_rates.data(), *prm, a);
    const HcFab* fabs[2] = {state, diag};
    return launch(PATH_EOS, 1, fabs, 2, &tile, k, stats, nullptr, (cudaStream_t)stream);
}

int hc_integrate_vec_host(int ntiles, const HcFab* state, const HcFab* diag, const HcBox* tiles, double a, double dt,
                          const HcParams* prm, HcStats* stats) {
    if (ntiles <= 0 || !state || !diag || !tiles) { set_err("bad argument"); return HC_ERR_ARG; }
    cudaStream_t s = nullptr;
    Staged S, D;
    if (int rc = stage_in(ntiles, state, {DENS, EDEN, EINT}, S, s)) return rc;
    if (int rc = stage_in(ntiles, diag, {TEMP, NE}, D, s)) return rc;
    int rc = hc_integrate_vec_batch(ntiles, S.dev.data(), D.dev.data(), tiles, a, dt, prm, stats, nullptr, s);
    if (rc == HC_OK) rc = stage_out(ntiles, state, {EDEN, EINT}, S, s);
    if (rc == HC_OK) rc = stage_out(ntiles, diag, {TEMP, NE}, D, s);
    stage_free(S, s); stage_free(D, s);
    if (rc == HC_OK) CUDA_TRY(cudaStreamSynchronize(s));
    return rc;
}

int hc_integrate_struct_host(int ntiles, const HcFab* s_old, const HcFab* diag, const HcFab* s_new, const HcFab* hydro_src,
                             const HcFab* reset_src, const HcFab* ir, const HcBox* tiles, double a, double a_end, double dt,
                             int sdc_iter, const HcParams* prm, HcStats* stats) {
    if (ntiles <= 0 || !s_old || !diag || !s_new || !hydro_src || !reset_src || !ir || !tiles) { set_err("bad argument"); return HC_ERR_ARG; }
    cudaStream_t s = nullptr;
    Staged SO, D, SN, H, R, I;
    const bool inhomo = prm && prm->inhomo_reion;
    int rc = stage_in(ntiles, s_old, {DENS, EDEN, EINT}, SO, s);
    if (rc == HC_OK) rc = stage_in(ntiles, diag, inhomo ? std::vector<int>{TEMP, NE, ZHI} : std::vector<int>{TEMP, NE}, D, s);
    if (rc == HC_OK) rc = stage_in(ntiles, s_new, {DENS, EDEN, EINT}, SN, s);
    if (rc == HC_OK) rc = stage_in(ntiles, hydro_src, {DENS, EINT}, H, s);
    if (rc == HC_OK) rc = stage_in(ntiles, reset_src, {0}, R, s);
    if (rc == HC_OK) rc = stage_in(ntiles, ir, {}, I, s);
    if (rc == HC_OK) rc = hc_integrate_struct_batch(ntiles, SO.dev.data(), D.dev.data(), SN.dev.data(), H.dev.data(), R.dev.data(), I.dev.data(),
                                                    tiles, a, a_end, dt, sdc_iter, prm, stats, nullptr, s);
    if (rc == HC_OK) {
        if (sdc_iter >= 0) { rc = stage_out(ntiles, s_new, {EDEN, EINT}, SN, s); if (rc == HC_OK) rc = stage_out(ntiles, ir, {0}, I, s); }
        else rc = stage_out(ntiles, s_old, {EDEN, EINT}, SO, s);
    }
    if (rc == HC_OK) rc = stage_out(ntiles, diag, {TEMP, NE}, D, s);
    stage_free(SO, s); stage_free(D, s); stage_free(SN, s); stage_free(H, s); stage_free(R, s); stage_free(I, s);
    if (rc == HC_OK) CUDA_TRY(cudaStreamSynchronize(s));
    return rc;
}

int hc_measure_fp64_peak(double* flops_per_s) {
    if (!flops_per_s) { set_err("null argument"); return HC_ERR_ARG; }
    int dev; if (int rc = current_device(dev)) return rc;
    int sms = 0; CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
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

int hc_sync(void* stream) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return HC_OK;
}

}  // extern "C"

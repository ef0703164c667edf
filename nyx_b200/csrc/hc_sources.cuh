// hc_sources.cuh -- SURVEY section 8f rank 2: the SDC source assembly either side of sdc_reactions, as ONE streaming sweep.
//
// Reference behaviour (three sweeps over the level, Source/TimeStep/Nyx_update_state_with_sources.cpp):
//   (1) :33-76    S_new(n) = f_n(S_old(n), hydro_src(n), ext_src_old(n)) for every state component
//   (2) :79-84    Nyx::enforce_minimum_density (Nyx_enforce_minimum_density.cpp:8-65): ONLY IF the minimum of the new density over the
//                 whole level is below small_dens: floor_density (Nyx_enforce_minimum_density.H:8-58) in the cells below it and, in the
//                 SDC build, hydro_src(rho) = S_new(rho) - S_old(rho) in EVERY cell (:40-63)
//   (3) :90-120   gravity: momenta += rho_old g dt / a_new, rho E += (rho u)_old . g dt a_half / a_new^2
// All three are cell-local; only the decision of (2) is global.  Here: hc_sources_kernel<false> does (1) + (3) in registers, stores S_new
// once and reduces the minimum of the sweep-(1) density into one word; hc_sources_kernel<true> recomputes a cell from the untouched
// inputs with (2) in between, and is either predicated on that word (single rank: no host round trip) or launched by the caller after
// the minimum has been reduced over ranks.  Per cell: 21 doubles read, 6 written = 216 algorithmic bytes (the reference's three sweeps
// move 192 + 120 + (2) bytes); HBM-bound, arithmetic is the reference's expression by expression (-fmad=false).
//
// Included by nyx_hc.cu inside its anonymous namespace (uses TileDesc, fab_off, cell_of).
#ifndef NYXB200_HC_SOURCES_CUH
#define NYXB200_HC_SOURCES_CUH

enum SrcSlot { SRC_UIN = 0, SRC_UOUT = 1, SRC_EXT = 2, SRC_HSRC = 3, SRC_GRAV = 4 };

struct SrcArgs {
    const TileDesc* tiles;
    int ntiles;
    long long ncells;
    unsigned long long* min_key;   // ordered-integer image of the minimum new density (see dens_key)
    int use_flag;                  // <true> kernel: return at once unless *min_key decodes below small_dens
    int sdc;                       // SDC build of the reference: (2) resets hydro_src(rho)
    int reset_hsrc;                // PHASE 2 only: the density was enforced (conservative variant) -> hydro_src(rho) = S_new(rho) - S_old(rho)
    double dt, a_old, a_half, a_half_inv, a_oldsq, a_newsq, a_new_inv, a_newsq_inv, dt_a_new, a_half_dt, dt_a_half;
    double small_dens, floor_rhoe; // floor_rhoe = small_dens * e(small_temp, Ne = 0)
};

// doubles ordered as unsigned integers: key(x) < key(y) <=> x < y (no NaNs reach this: they fail the `<` that guards the update)
__host__ __device__ __forceinline__ unsigned long long dens_key(double x) {
#if defined(__CUDA_ARCH__)
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b; memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dens_from_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x; memcpy(&x, &b, 8); return x;
#endif
}

constexpr int SRC_THREADS = 256;
constexpr int SRC_U = 2;            // cells per thread and pass: 42 loads in flight before the first store

// PHASE 0: sweeps (1) + (3) fused (the floor variant's path).  The "conservative" variant of enforce_minimum_density moves density between
// neighbour cells of the sweep-(1) state through the caller's FillPatch, so there the two sweeps stay apart:
// PHASE 1: sweep (1) only (+ the minimum);  PHASE 2: sweep (3) only, on the S_new the conservative iterations left (+ the SDC reset of
// hydro_src(rho), Nyx_enforce_minimum_density.cpp:40-63, when a.reset_hsrc).
template <bool ENFORCE, int PHASE = 0>
__global__ void __launch_bounds__(SRC_THREADS) hc_sources_kernel(const __grid_constant__ SrcArgs a) {
    if (ENFORCE && a.use_flag) {
        if (!(dens_from_key(*a.min_key) < a.small_dens)) return;
    }
    double vmin = DBL_MAX;   // amrex::MultiFab::min starts from std::numeric_limits<Real>::max()
    int t0 = -1;
    for (long long base = (long long)blockIdx.x * (SRC_THREADS * SRC_U); base < a.ncells; base += (long long)gridDim.x * (SRC_THREADS * SRC_U)) {
        if (t0 < 0) t0 = find_tile_by_cell(a.tiles, a.ntiles, base);
        while (t0 + 1 < a.ntiles && a.tiles[t0 + 1].offset <= base) ++t0;
        double ui[SRC_U][6], hs[SRC_U][6], ex[SRC_U][6], g[SRC_U][3];
        double* po[SRC_U]; double* ph[SRC_U];
        long long nso[SRC_U];
        bool on[SRC_U];
#pragma unroll
        for (int u = 0; u < SRC_U; ++u) {
            const long long id = base + u * SRC_THREADS + threadIdx.x;
            on[u] = id < a.ncells;
            po[u] = nullptr; ph[u] = nullptr; nso[u] = 0;
            if (on[u]) {
                int ti = t0;
                while (ti + 1 < a.ntiles && a.tiles[ti + 1].offset <= id) ++ti;
                const TileDesc& t = a.tiles[ti];
                int i, j, k;
                cell_of(t, id, i, j, k);
                const HcFab& Fi = t.f[SRC_UIN]; const HcFab& Fo = t.f[SRC_UOUT]; const HcFab& Fe = t.f[SRC_EXT];
                const HcFab& Fh = t.f[SRC_HSRC]; const HcFab& Fg = t.f[SRC_GRAV];
                const double* pi = Fi.p + fab_off(Fi, i, j, k);
                const double* pe = Fe.p + fab_off(Fe, i, j, k);
                const double* pg = Fg.p + fab_off(Fg, i, j, k);
                ph[u] = Fh.p + fab_off(Fh, i, j, k);
                po[u] = Fo.p + fab_off(Fo, i, j, k); nso[u] = Fo.nstride;
                if (PHASE == 2) {
                    // sweep (3) alone: the OLD state (rho, momenta), the gravity vector and the S_new of the sweeps before (kept in hs[])
#pragma unroll
                    for (int n = 0; n < 4; ++n) ui[u][n] = __ldg(pi + n * Fi.nstride);
#pragma unroll
                    for (int n = 0; n < 6; ++n) hs[u][n] = po[u][n * nso[u]];
                } else {
#pragma unroll
                    for (int n = 0; n < 6; ++n) { ui[u][n] = __ldg(pi + n * Fi.nstride); hs[u][n] = ph[u][n * Fh.nstride]; ex[u][n] = __ldg(pe + n * Fe.nstride); }
                }
                if (PHASE != 1) {
#pragma unroll
                    for (int n = 0; n < 3; ++n) g[u][n] = __ldg(pg + n * Fg.nstride);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SRC_U; ++u) {
            if (!on[u]) continue;
            double o[6];
            if (PHASE == 2) {
#pragma unroll
                for (int n = 0; n < 6; ++n) o[n] = hs[u][n];
                if (a.sdc && a.reset_hsrc) ph[u][0] = o[0] - ui[u][0];
            } else {
            // sweep (1), Nyx_update_state_with_sources.cpp:47-73
            o[0] = ui[u][0] + hs[u][0] + a.dt * ex[u][0] * a.a_half_inv;
#pragma unroll
            for (int n = 1; n <= 3; ++n) {
                o[n] = a.a_old * ui[u][n] + hs[u][n] + a.dt * ex[u][n];
                o[n] = o[n] * a.a_new_inv;
            }
#pragma unroll
            for (int n = 4; n <= 5; ++n) {
                o[n] = a.a_oldsq * ui[u][n] + hs[u][n] + a.a_half_dt * ex[u][n];
                o[n] = o[n] * a.a_newsq_inv;
            }
            }
            if (PHASE == 2) {
            } else if (!ENFORCE) {
                if (o[0] < vmin) vmin = o[0];
            } else {
                // sweep (2): floor_density, then (SDC) hydro_src(rho) = A_rho in every cell
                if (o[0] < a.small_dens) {
                    o[0] = a.small_dens;
                    o[1] = 0.0; o[2] = 0.0; o[3] = 0.0;
                    o[5] = a.floor_rhoe;
                    o[4] = o[5];
                }
                if (a.sdc) ph[u][0] = o[0] - ui[u][0];
            }
            if (PHASE == 1) {
#pragma unroll
                for (int n = 0; n < 6; ++n) po[u][n * nso[u]] = o[n];
                continue;
            }
            // sweep (3), :97-119 (rho and the momenta of the OLD state)
            const double rho = ui[u][0];
            const double SrU = rho * g[u][0], SrV = rho * g[u][1], SrW = rho * g[u][2];
            o[1] += SrU * a.dt_a_new;
            o[2] += SrV * a.dt_a_new;
            o[3] += SrW * a.dt_a_new;
            const double SrE = ui[u][1] * g[u][0] + ui[u][2] * g[u][1] + ui[u][3] * g[u][2];
            o[4] = (a.a_newsq * o[4] + SrE * a.dt_a_half) * a.a_newsq_inv;
#pragma unroll
            for (int n = 0; n < 6; ++n) po[u][n * nso[u]] = o[n];
        }
    }
    if (!ENFORCE && PHASE != 2) {
        // block minimum -> one atomicMin per CTA
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, vmin, s); if (w < vmin) vmin = w; }
        __shared__ double s_min[SRC_THREADS / 32];
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = vmin;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = s_min[0];
#pragma unroll
            for (int w = 1; w < SRC_THREADS / 32; ++w) if (s_min[w] < m) m = s_min[w];
            atomicMin(a.min_key, dens_key(m));
        }
    }
}

// ---- enforce_minimum_density, "conservative" variant: ONE iteration of the loop of Nyx::enforce_minimum_density_cons
// (Source/TimeStep/Nyx_enforce_minimum_density.cpp:179-236) on a border-filled copy of the new state.  The reference does it in two sweeps
// with three face-centred work arrays: compute_mu_for_enforce_min (Nyx_enforce_minimum_density.H:60-157) SCATTERS, from every cell below
// small_dens of the tile grown by one, a diffusion coefficient to the faces it draws density through; create_update_for_minimum (:159-200)
// then forms update(n) = div(mu grad state(n)) in the valid cells; then S_new += update, reset_e_src = update(rho e).  A face is written by at
// most one of its two cells (a cell below small_dens < target = 1.01 small_dens has nothing to give), so the coefficient of a face is a pure
// function of the densities within two cells of it: here every valid cell GATHERS its six coefficients and applies the update at once -- no
// work arrays, no second sweep; same expressions, same order (-fmad=false).  FillPatch between iterations stays with the caller (it is the
// host framework's ghost exchange); the new minimum density comes back for the caller's loop test.
enum ConsSlot { CONS_SBORD = 0, CONS_SNEW = 1, CONS_RSRC = 2 };
struct ConsArgs {
    const TileDesc* tiles;
    int ntiles;
    long long ncells;
    unsigned long long* min_key;   // minimum of the new density over the valid cells (dens_key image)
    unsigned long long* n_bad;     // faces with a negative coefficient (the reference aborts: "mu_x(i+1,j,k) < 0")
    double small_dens;
    int sdc;                       // SDC build: reset_e_src = update(rho e)
};
__device__ __forceinline__ double cons_max0(double v) { return (v < 0.0) ? 0.0 : v; }   // amrex::max(v, 0.0)
// the fraction of its neighbours' offers a cell below small_dens takes (:84-124); r points at the cell's density
__device__ __forceinline__ double cons_fac(const double* r, long long js, long long ks, double target) {
    const double total_need = target - r[0];
    const double a_ihi = cons_max0((r[1] - target) / 6.0), a_ilo = cons_max0((r[-1] - target) / 6.0);
    const double a_jhi = cons_max0((r[js] - target) / 6.0), a_jlo = cons_max0((r[-js] - target) / 6.0);
    const double a_khi = cons_max0((r[ks] - target) / 6.0), a_klo = cons_max0((r[-ks] - target) / 6.0);
    const double total_avail = a_ihi + a_ilo + a_jhi + a_jlo + a_khi + a_klo;
    return (total_need < total_avail) ? total_need / total_avail : 1.0;
}
// coefficient of the face between the cell at r1 - st (lower) and the cell at r1 (upper) along the direction of stride st (:128-156)
__device__ __forceinline__ double cons_mu(const double* r1, long long st, long long js, long long ks, double small, double target, unsigned& bad) {
    const double lo = r1[-st], hi = r1[0];
    double mu = 0.0;
    if (lo < small) {          // the lower cell draws from its "hi" neighbour
        const double from = cons_fac(r1 - st, js, ks, target) * cons_max0((hi - target) / 6.0);
        if (from > 0) { mu = from / (hi - lo); if (mu < 0.) bad++; }
    }
    if (hi < small) {          // the upper cell draws from its "lo" neighbour
        const double from = cons_fac(r1, js, ks, target) * cons_max0((lo - target) / 6.0);
        if (from > 0) { mu = -from / (hi - lo); if (mu < 0.) bad++; }
    }
    return mu;
}
__global__ void __launch_bounds__(256) hc_min_dens_cons_kernel(const __grid_constant__ ConsArgs a) {
    const double small = a.small_dens, target = 1.01 * a.small_dens;
    double vmin = DBL_MAX;
    unsigned bad = 0;
    int ti = -1;
    for (long long id = (long long)blockIdx.x * 256 + threadIdx.x; id < a.ncells; id += (long long)gridDim.x * 256) {
        if (ti < 0) ti = find_tile_by_cell(a.tiles, a.ntiles, id);
        while (ti + 1 < a.ntiles && a.tiles[ti + 1].offset <= id) ++ti;
        const TileDesc& t = a.tiles[ti];
        int i, j, k;
        cell_of(t, id, i, j, k);
        const HcFab& B = t.f[CONS_SBORD];
        const HcFab& N = t.f[CONS_SNEW];
        const long long js = B.jstride, ks = B.kstride;
        const double* r = B.p + fab_off(B, i, j, k);   // component 0 = density
        const double mx0 = cons_mu(r, 1, js, ks, small, target, bad), mx1 = cons_mu(r + 1, 1, js, ks, small, target, bad);
        const double my0 = cons_mu(r, js, js, ks, small, target, bad), my1 = cons_mu(r + js, js, js, ks, small, target, bad);
        const double mz0 = cons_mu(r, ks, js, ks, small, target, bad), mz1 = cons_mu(r + ks, ks, js, ks, small, target, bad);
        double* pn = N.p + fab_off(N, i, j, k);
#pragma unroll
        for (int n = 0; n < 6; ++n) {
            const double* s = r + (long long)n * B.nstride;
            const double upd = mx1 * (s[1] - s[0]) - mx0 * (s[0] - s[-1]) + my1 * (s[js] - s[0]) - my0 * (s[0] - s[-js])
                             + mz1 * (s[ks] - s[0]) - mz0 * (s[0] - s[-ks]);
            const double v = pn[(long long)n * N.nstride] + upd;       // S_new.plus(update, 0, nComp, 0)
            pn[(long long)n * N.nstride] = v;
            if (n == 0 && v < vmin) vmin = v;
            if (n == 5 && a.sdc) t.f[CONS_RSRC].p[fab_off(t.f[CONS_RSRC], i, j, k)] = upd;   // MultiFab::Copy(reset_e_src, update, Eint_comp, 0, 1, 0)
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, vmin, s); if (w < vmin) vmin = w; }
    if ((threadIdx.x & 31) == 0) atomicMin(a.min_key, dens_key(vmin));
    if (bad) atomicAdd(a.n_bad, (unsigned long long)bad);
}

// MultiFab::Copy / Add / Subtract of one component range (Source/Hydro/sdc_hydro.cpp:83-84,94-95,112,135): dst(dcomp + n) (=, +=, -=) src(scomp + n)
struct FabOpArgs {
    const TileDesc* tiles;
    int ntiles;
    long long ncells, nchunks;
    int scomp, dcomp, ncomp, op;   // op 0: copy, 1: add, 2: subtract
};
// One warp per chunk (a piece of an x-row, at most CHUNK_MAX = 256 cells): the tile lookup and the (j, k) decode cost three integer divisions
// per CHUNK instead of two per CELL -- with three memory operations per cell the per-cell index arithmetic was what bounded this kernel
// (0.255 ms = 48 % of HBM peak for 3.4e7 cells whatever the grid).  Lanes stride over the row: loads of up to 8 cells per lane first, then stores.
__global__ void __launch_bounds__(256) hc_fab_op_kernel(const __grid_constant__ FabOpArgs a) {
    const unsigned lane = threadIdx.x & 31u;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long ch = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); ch < a.nchunks; ch += nwarps) {
        const TileDesc& t = a.tiles[find_tile_by_chunk(a.tiles, a.ntiles, ch)];
        const unsigned local = (unsigned)(ch - t.chunk_begin);
        const unsigned row = local / (unsigned)t.cpr, piece = local - row * (unsigned)t.cpr;
        const unsigned kk = row / (unsigned)t.ny;
        const int k = t.lo[2] + (int)kk, j = t.lo[1] + (int)(row - kk * (unsigned)t.ny);
        const int x0 = t.lo[0] + (int)piece * t.chunk_len;
        const int len = min(t.chunk_len, t.lo[0] + t.nx - x0);
        const HcFab& D = t.f[0]; const HcFab& S = t.f[1];
        double* pd = D.p + fab_off(D, x0, j, k) + (long long)a.dcomp * D.nstride + lane;
        const double* ps = S.p + fab_off(S, x0, j, k) + (long long)a.scomp * S.nstride + lane;
        for (int n = 0; n < a.ncomp; ++n, pd += D.nstride, ps += S.nstride) {
            double sv[8], dv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                sv[u] = 0.0; dv[u] = 0.0;
                if ((int)lane + 32 * u < len) { sv[u] = __ldg(ps + 32 * u); if (a.op != 0) dv[u] = pd[32 * u]; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if ((int)lane + 32 * u < len) pd[32 * u] = (a.op == 0) ? sv[u] : (a.op == 1) ? dv[u] + sv[u] : dv[u] - sv[u];
        }
    }
}

// Nyx::init_zhi, the cell loop (Source/Initialization/Nyx_initdata.cpp:198-209): piecewise-constant injection of the coarse reionization-redshift
// field into diag(Zhi_comp): D_new(i,j,k,Zhi) = zhi(i/ratio, j/ratio, k/ratio) (C++ integer division).  f[0] = diag, f[1] = coarse zhi.
struct ZhiArgs {
    const TileDesc* tiles;
    int ntiles;
    long long ncells;
    int ratio, zcomp;
};
__global__ void __launch_bounds__(256) hc_init_zhi_kernel(const __grid_constant__ ZhiArgs a) {
    int t0 = -1;
    for (long long id = (long long)blockIdx.x * 256 + threadIdx.x; id < a.ncells; id += (long long)gridDim.x * 256) {
        if (t0 < 0) t0 = find_tile_by_cell(a.tiles, a.ntiles, id);
        while (t0 + 1 < a.ntiles && a.tiles[t0 + 1].offset <= id) ++t0;
        const TileDesc& t = a.tiles[t0];
        int i, j, k;
        cell_of(t, id, i, j, k);
        const HcFab& D = t.f[0]; const HcFab& Z = t.f[1];
        D.p[fab_off(D, i, j, k) + (long long)a.zcomp * D.nstride] = __ldg(Z.p + fab_off(Z, i / a.ratio, j / a.ratio, k / a.ratio));
    }
}

#endif

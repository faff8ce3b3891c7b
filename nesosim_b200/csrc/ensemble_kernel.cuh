// ensemble_kernel.cuh -- season-resident path for small grids (the 100 km calibration ensemble).
//
// One thread-block CLUSTER of 4 CTAs owns one ensemble member for the whole season; each CTA holds a strip of
// rows (strips are cut by the host so that their ocean-cell counts balance).  What stays on chip for all T-1 days:
//   * both snow layers (h0,h1) of the strip + a 2-row halo, interleaved as double2 in a double-buffered shared
//     memory tile; the two boundary rows of each neighbour strip are pushed into the neighbour's halo through
//     distributed shared memory (st.shared::cluster) -- one cluster barrier per day, no global round trip;
//   * the seven member-dependent accumulators (snowAdv, snowDiv, snowLead, snowAtm, snowWindPackLoss/Gain/Net)
//     in REGISTERS of the thread that owns the cell.
// The land mask is compiled into per-strip cell lists once per context: only ocean cells are owned and
// advanced; only cells with an ocean cell in their 3x3 neighbourhood get raw advection/divergence; land cells
// (56 % of the 100 km grid) have closed-form outputs after the first step -- h and density NaN, every
// accumulator NaN if its switch is on and 0 otherwise (NaN + anything = NaN, x + 0 = x) -- and are only stored.
// HBM therefore sees what SURVEY.md §8d counts as algorithmic traffic: the 12 output planes written once
// (96 B per member-cell-day) plus the member-independent forcing, which is pre-digested once per season by two
// small kernels (drift gradients; snowfall -> accumulation/ocean flux and their running sums) and then shared
// by every member through L2.
//
// Arithmetic is the per-cell code of cell_math.cuh, shared with the general path: results are value-identical.
#pragma once
#include <cooperative_groups.h>

#include "cell_math.cuh"
#include "day_kernels.cuh"

namespace nesosim {

namespace cg = cooperative_groups;

// ------------------------------------------------------------------ member-independent pre-pass (per season)
// DA[x][cell][2] = (ut, vt), (gx(ut), gy(vt))   drift displacement and its gradients   (NESOSIM.py:204-205)
// DB[x][cell]    = (acc, 1-C)                   accumulation delta (NESOSIM.py:260-263), open-water fraction
// DC[x][cell]    = (snowAcc[x+1], snowOcean[x+1])   running sums (NESOSIM.py:264,268)

struct DeriveArgs {
    int ny, nx, steps;                 // steps = T-1
    const double *P, *C, *UV;          // [T][plane], [T][plane], [T][2][plane]
    double2 *DA;                       // [steps][plane][2]
    double2 *DB, *DC;                  // [steps][plane]
    ModelConsts k;
    GradConsts g;
    ConstDiv rho_new;
};

__global__ void derive_pointwise_kernel(const __grid_constant__ DeriveArgs a) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (gx >= a.nx || gy >= a.ny) return;
    const long long plane = (long long)a.ny * a.nx, o = (long long)gy * a.nx + gx;
    const double *U = a.UV + (long long)x * 2 * plane, *V = U + plane;
    const int xm = max(gx - 1, 0), xp = min(gx + 1, a.nx - 1), ym = max(gy - 1, 0), yp = min(gy + 1, a.ny - 1);
    auto ut = [&](int r, int c) { return mul(U[(long long)r * a.nx + c], a.k.deltaT); };
    auto vt = [&](int r, int c) { return mul(V[(long long)r * a.nx + c], a.k.deltaT); };
    const double utc = ut(gy, gx), vtc = vt(gy, gx);
    double2 *da = a.DA + ((long long)x * plane + o) * 2;
    da[0] = make_double2(utc, vtc);
    da[1] = make_double2(gradient1d(ut(gy, xm), utc, ut(gy, xp), gx, a.nx, a.g),
                         gradient1d(vt(ym, gx), vtc, vt(yp, gx), gy, a.ny, a.g));
    const double C = a.C[(long long)x * plane + o];
    const double pd = div_const(a.P[(long long)x * plane + o], a.rho_new);
    const double omc = sub(1.0, C);
    a.DB[(long long)x * plane + o] = make_double2(mul(pd, C), omc);
    a.DC[(long long)x * plane + o] = make_double2(0.0, -mul(pd, omc));   // .y parks oc until the scan
}

// One thread per cell, sequential in time, loads batched 8 days ahead (add-latency bound, not load-latency bound).
__global__ void derive_scan_kernel(const double2 *DB, double2 *DC, long long plane, int steps) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= plane) return;
    double sa = 0.0, so = 0.0;
    constexpr int B = 8;
    for (int x0 = 0; x0 < steps; x0 += B) {
        double da[B], dq[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int x = min(x0 + j, steps - 1);
            da[j] = DB[(long long)x * plane + o].x;
            dq[j] = DC[(long long)x * plane + o].y;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (x0 + j < steps) {
                sa = add(sa, da[j]);
                so = add(so, dq[j]);
                DC[(long long)(x0 + j) * plane + o] = make_double2(sa, so);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ the season kernel

constexpr int ENS_CLUSTER = 4;
constexpr int ENS_NT = 512;            // 16 warps, up to 128 registers per thread
constexpr int ENS_MAXR = 28;           // rows one CTA can own
constexpr int ENS_SX = 96;             // h tile row stride (double2 elements); columns 0..nx-1
constexpr int ENS_SXR = 98;            // raw tile row stride; column c lives at c+1 (zero pad each side)
constexpr int ENS_MAX_NX = 96;
constexpr int ENS_KLB = 3;             // land cells handled per thread per batch

// Shared memory (double2 units), sized by the host for the tallest strip (`rows`) and the longest raw list:
//   h tiles  [2 parity][rows+4][SX]   own rows + 2 halo rows each side
//   raw adv  [rows+2][SXR], raw div [rows+2][SXR]
//   staged drift terms [2][n_raw]     next day's (ut,vt),(gxu,gyv) per raw-list entry (cp.async), optional
//   raw and land code lists (uint16)  (L1 is invalidated by every cluster barrier, so they must not live there)
inline size_t ens_smem_bytes(int rows, int n_stage, int n_codes) {
    return (size_t)(2 * (rows + 4) * ENS_SX + 2 * (rows + 2) * ENS_SXR + 2 * n_stage) * sizeof(double2) +
           (size_t)((n_codes + 7) / 8 * 8) * sizeof(unsigned short);
}

// Per-strip cell lists (uint16 code = row*128 + col; row is global for raw lists, strip-local for owned cells).
struct StripTables {
    const unsigned short *codes;       // all lists concatenated (device)
    int row0[ENS_CLUSTER + 1];         // strip k owns rows row0[k] .. row0[k+1]-1
    int raw_off[ENS_CLUSTER], raw_int_n[ENS_CLUSTER], raw_n[ENS_CLUSTER];   // interior entries first, then edge entries
    int ocean_off[ENS_CLUSTER], ocean_n[ENS_CLUSTER];
    int land_off[ENS_CLUSTER], land_n[ENS_CLUSTER];
    int rows_alloc;                    // tallest strip
    int stage_alloc;                   // staged entries per CTA (0: read the drift terms straight from L2)
    int raw_alloc, land_alloc;         // longest raw / land list (shared-memory copies)
};

struct EnsArgs {
    int ny, nx, T, M;
    const double2 *DA, *DB, *DC;
    const double *W;                   // wind [T][plane]
    const double *ic;                  // NULL -> zero depth
    long long ic_stride;               // 0 shared, plane per member
    const double *conc0;
    double *out[NVAR];                 // member 0, slot 0 of each array (NULL: not stored)
    long long mstride[NVAR];
    const MemberCoef *coef;
    ModelConsts k;
    GradConsts g;
    ConstDiv conv_div;
    double w[9];
    Switches sw;
    StripTables st;
    long long *timing;                 // debug: [gridDim.x][8] phase cycle totals of thread 0 (NULL = off)
};

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// streaming store: every output element is written exactly once and never read back by this kernel
__device__ __forceinline__ void st_out(double *p, double v) { __stcs(p, v); }

// KO = owned ocean cells per thread (capacity KO*512 per strip); ALLOUT = all twelve outputs requested;
// STAGE = next day's drift terms are staged in shared memory with cp.async during the current day.
template <int KO, bool ALLOUT, bool STAGE>
__global__ void __cluster_dims__(ENS_CLUSTER, 1, 1) __launch_bounds__(ENS_NT, 1)
ensemble_season_kernel(const __grid_constant__ EnsArgs a) {
    constexpr int NT = ENS_NT, SX = ENS_SX, SXR = ENS_SXR, KLB = ENS_KLB;
    extern __shared__ __align__(16) double2 smem2[];
    const int TR = a.st.rows_alloc + 4, RR = a.st.rows_alloc + 2;
    double2 *s_h = smem2;                              // [2 parity][TR][SX]   (h0,h1)
    double2 *s_adv = s_h + 2 * TR * SX;                // [RR][SXR]            (adv0,adv1) after NaN->0
    double2 *s_div = s_adv + RR * SXR;                 // [RR][SXR]            (div0,div1) after NaN->0
    double2 *s_da = s_div + RR * SXR;                  // [n_raw][2]           staged (ut,vt),(gxu,gyv)
    unsigned short *s_raw_code = reinterpret_cast<unsigned short *>(s_da + 2 * a.st.stage_alloc);
    unsigned short *s_land_code = s_raw_code + (a.st.raw_alloc + 7) / 8 * 8;

    cg::cluster_group cluster = cg::this_cluster();
    const int k = (int)cluster.block_rank();
    const int cid = blockIdx.x / ENS_CLUSTER, ncl = gridDim.x / ENS_CLUSTER;
    const int ny = a.ny, nx = a.nx;
    const long long plane = (long long)ny * nx;
    const int ra = a.st.row0[k], rb = a.st.row0[k + 1], nrow = rb - ra;
    const int tid = threadIdx.x;
    const int steps = a.T - 1;

    // neighbours' tiles through distributed shared memory; a global row r sits at tile row r - ra_nb + 2
    double2 *nb_up = (k > 0) ? cluster.map_shared_rank(s_h, k - 1) : nullptr;
    double2 *nb_dn = (k < ENS_CLUSTER - 1) ? cluster.map_shared_rank(s_h, k + 1) : nullptr;
    const int up_shift = (k > 0) ? (ra - a.st.row0[k - 1]) : 0;   // my local row lr -> their tile row lr + 2 + up_shift
    const int dn_shift = nrow;                                     // my local row lr -> their tile row lr + 2 - dn_shift

    const unsigned short *ocean = a.st.codes + a.st.ocean_off[k];
    const int n_raw_int = a.st.raw_int_n[k], n_raw = a.st.raw_n[k];
    const int n_ocean = a.st.ocean_n[k], n_land = a.st.land_n[k];
    for (int i = tid; i < n_raw; i += NT) s_raw_code[i] = a.st.codes[a.st.raw_off[k] + i];
    for (int i = tid; i < n_land; i += NT) s_land_code[i] = a.st.codes[a.st.land_off[k] + i];
    const unsigned short *raw = s_raw_code, *land = s_land_code;

    // cells this thread owns for the whole season: ocean[tid + j*NT]
    int own_t[KO];       // tile offset (lr+2)*SX + col, or -1
    int own_o[KO];       // global cell offset
#pragma unroll
    for (int j = 0; j < KO; ++j) {
        const int idx = tid + j * NT;
        own_t[j] = -1;
        own_o[j] = 0;
        if (idx < n_ocean) {
            const int code = ocean[idx], lr = code >> 7, c = code & 127;
            own_t[j] = (lr + 2) * SX + c;
            own_o[j] = (ra + lr) * nx + c;
        }
    }

    for (int i = tid; i < RR * SXR; i += NT) {         // zero padding of convolve(boundary='fill')
        s_adv[i] = make_double2(0.0, 0.0);
        s_div[i] = make_double2(0.0, 0.0);
    }
    const double nan = qnan();
    // closed-form land values from slot 2 on (slot 1 is computed from the initial depths)
    const double landAdv = a.sw.dynamics ? nan : 0.0, landLead = a.sw.leadloss ? nan : 0.0;
    const double landAtm = a.sw.atmloss ? nan : 0.0, landWp = a.sw.windpack ? nan : 0.0;

    // write (h0,h1) of local row lr, column c into the tiles of parity `par` (own copy + neighbours' halos)
    auto put_h = [&](int par, int lr, int c, double2 v) {
        s_h[(par * TR + lr + 2) * SX + c] = v;
        if (lr < 2 && nb_up) nb_up[(par * TR + lr + 2 + up_shift) * SX + c] = v;
        if (lr >= nrow - 2 && nb_dn) nb_dn[(par * TR + lr + 2 - dn_shift) * SX + c] = v;
    };
    auto want = [&](int v) { return ALLOUT || a.out[v] != nullptr; };
    // stage day x's drift terms for this thread's raw-list entries
    auto stage_day = [&](int x) {
        const double2 *DAx = a.DA + (long long)x * plane * 2;
        for (int i = tid; i < n_raw; i += NT) {
            const int code = raw[i], r = code >> 7, c = code & 127;
            const double2 *src = DAx + (r * nx + c) * 2;
            cp_async16(s_da + 2 * i, src);
            cp_async16(s_da + 2 * i + 1, src + 1);
        }
    };

    cluster.sync();   // every CTA of the cluster is running before anyone writes into a neighbour's shared memory

    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool timing = a.timing != nullptr && tid == 0;
    long long tlast = 0;
#define ENS_TICK(slot)                                   \
    if (timing) {                                        \
        const long long now_ = clock64();                \
        tacc[slot] += now_ - tlast;                      \
        tlast = now_;                                    \
    }

    for (int m = cid; m < a.M; m += ncl) {
        const MemberCoef mc = a.coef[m];
        double accAdv[KO], accDiv[KO], accLead[KO], accAtm[KO], accWpl[KO], accWpg[KO], accWp[KO];
#pragma unroll
        for (int j = 0; j < KO; ++j) accAdv[j] = accDiv[j] = accLead[j] = accAtm[j] = accWpl[j] = accWpg[j] = accWp[j] = 0.0;

        // output address of variable v, time slot `slot` of member m (add the cell offset)
        const long long mo_plane = (long long)m * a.mstride[V_DENS], mo_depth = (long long)m * a.mstride[V_H0];
        auto outp = [&](int v, int slot) -> double * {
            return (v == V_H0 || v == V_H1) ? a.out[v] + (mo_depth + (long long)slot * 2 * plane)
                                            : a.out[v] + (mo_plane + (long long)slot * plane);
        };

        // ---- slot 0: genEmptyArrays zeros + the IC split of main (NESOSIM.py:604-609), every cell of the strip.
        // (No barrier against the previous member: its last day ended with a cluster barrier after all reads.)
        if (STAGE && a.sw.dynamics) stage_day(0);
        for (int i = tid; i < nrow * nx; i += NT) {
            const int lr = i / nx, c = i - lr * nx;
            const long long o = (long long)(ra + lr) * nx + c;
            double half = 0.0;
            if (a.ic) {
                double v = a.ic[(long long)m * a.ic_stride + o];
                if (a.conc0[o] < a.k.minConc) v = 0.0;
                half = mul(v, 0.5);
            }
            put_h(0, lr, c, make_double2(half, half));
#pragma unroll
            for (int v = 0; v < NVAR; ++v)
                if (want(v)) st_out(outp(v, 0) + o, (v == V_H0 || v == V_H1) ? half : 0.0);
        }
        if (STAGE) cp_async_wait_all();
        cluster.sync();

        if (timing) tlast = clock64();
        for (int x = 0; x < steps; ++x) {
            const int par = x & 1;
            const double2 *hcur = s_h + par * TR * SX;
            const double2 *DAx = a.DA + (long long)x * plane * 2;
            const double2 *DBx = a.DB + (long long)x * plane;
            const double2 *DCx = a.DC + (long long)x * plane;
            const double *Wx = a.W + (long long)x * plane;

            // member-independent inputs of the owned cells, requested before phase A so L2 latency overlaps it
            double2 f_b[KO], f_c[KO];
            double f_W[KO];
#pragma unroll
            for (int j = 0; j < KO; ++j) {
                f_b[j] = f_c[j] = make_double2(0.0, 0.0);
                f_W[j] = 0.0;
                if (own_t[j] >= 0) {
                    f_b[j] = __ldg(DBx + own_o[j]);
                    f_c[j] = __ldg(DCx + own_o[j]);
                    f_W[j] = __ldg(Wx + own_o[j]);
                }
            }

            double2 l_c[KLB];   // running sums for the first batch of land cells (copied to the outputs)
#pragma unroll
            for (int q = 0; q < KLB; ++q) {
                const int idx = q * NT + tid;
                l_c[q] = make_double2(0.0, 0.0);
                if (idx < n_land) {
                    const int code = land[idx];
                    l_c[q] = __ldg(DCx + (ra + (code >> 7)) * nx + (code & 127));
                }
            }

            // ---------------- phase A: raw advection / divergence (calcDynamics, NESOSIM.py:189-222) where an
            // ocean cell of this strip will read it
            if (a.sw.dynamics) {
                for (int i = tid; i < n_raw; i += NT) {
                    const int code = raw[i], r = code >> 7, c = code & 127;
                    double2 d01, d23;
                    if (STAGE) {
                        d01 = s_da[2 * i];
                        d23 = s_da[2 * i + 1];
                    } else {
                        d01 = __ldg(DAx + (r * nx + c) * 2);
                        d23 = __ldg(DAx + (r * nx + c) * 2 + 1);
                    }
                    const double2 *hp = hcur + (r - ra + 2) * SX + c;
                    const double2 hc = hp[0];
                    double gx0, gy0, gx1, gy1;
                    if (i < n_raw_int) {
                        const double2 hl = hp[-1], hr = hp[1], hu = hp[-SX], hd = hp[SX];
                        gx0 = div_const(sub(hr.x, hl.x), a.g.two_dx);
                        gy0 = div_const(sub(hd.x, hu.x), a.g.two_dx);
                        gx1 = div_const(sub(hr.y, hl.y), a.g.two_dx);
                        gy1 = div_const(sub(hd.y, hu.y), a.g.two_dx);
                    } else {   // first/last row or column: one-sided differences (np.gradient edge_order=1)
                        const double2 hl = hp[c > 0 ? -1 : 0], hr = hp[c < nx - 1 ? 1 : 0];
                        const double2 hu = hp[r > 0 ? -SX : 0], hd = hp[r < ny - 1 ? SX : 0];
                        gx0 = gradient1d(hl.x, hc.x, hr.x, c, nx, a.g);
                        gy0 = gradient1d(hu.x, hc.x, hd.x, r, ny, a.g);
                        gx1 = gradient1d(hl.y, hc.y, hr.y, c, nx, a.g);
                        gy1 = gradient1d(hu.y, hc.y, hd.y, r, ny, a.g);
                    }
                    const int ro = (r - ra + 1) * SXR + c + 1;
                    s_adv[ro] = make_double2(zero_if_nonfinite(adv_term(d01.x, d01.y, gx0, gy0)),
                                             zero_if_nonfinite(adv_term(d01.x, d01.y, gx1, gy1)));
                    s_div[ro] = make_double2(zero_if_nonfinite(div_term(hc.x, d23.x, d23.y)),
                                             zero_if_nonfinite(div_term(hc.y, d23.x, d23.y)));
                }
            }
            ENS_TICK(0)   // prefetch issue + phase A
            __syncthreads();
            ENS_TICK(1)   // wait for the CTA
            // ---------------- land cells: no state.  Step 0 sees the initial depths; afterwards h is NaN, so every
            // switched-on term is NaN and every switched-off term adds 0 (NESOSIM.py:287-322): closed form.
            for (int base = 0; base < n_land; base += KLB * NT) {
                int lo[KLB], lt[KLB];
                double2 cum[KLB];
#pragma unroll
                for (int q = 0; q < KLB; ++q) {
                    const int idx = base + q * NT + tid;
                    lo[q] = -1;
                    lt[q] = 0;
                    if (idx < n_land) {
                        const int code = land[idx], lr = code >> 7, c = code & 127;
                        lo[q] = (ra + lr) * nx + c;
                        lt[q] = (lr + 2) * SX + c;
                    }
                }
#pragma unroll
                for (int q = 0; q < KLB; ++q)   // the first batch was requested at the top of the day
                    cum[q] = (base == 0) ? l_c[q] : ((lo[q] >= 0) ? __ldg(DCx + lo[q]) : make_double2(0.0, 0.0));
#pragma unroll
                for (int q = 0; q < KLB; ++q) {
                    if (lo[q] < 0) continue;
                    const int o = lo[q];
                    double vLead = landLead, vAtm = landAtm, vWpl = landWp, vWpg = landWp, vWp = landWp;
                    if (x == 0) {
                        const double2 h = hcur[lt[q]];
                        const double W = __ldg(Wx + o);
                        const double omc = __ldg(DBx + o).y;
                        const double wt = wind_flag(W, mc.wpt);
                        vLead = add(0.0, a.sw.leadloss ? -mul(mul(mul(mul(mul(wt, mc.llf), a.k.deltaT), h.x), W), omc) : 0.0);
                        vAtm = add(0.0, a.sw.atmloss ? atm_loss(wt, h.x, W, mc, a.k) : 0.0);
                        double wpl = 0.0, wpg = 0.0, wpn = 0.0;
                        if (a.sw.windpack) wind_packing(wt, h.x, mc, a.k, wpl, wpg, wpn);
                        vWpl = add(0.0, wpl);
                        vWpg = add(0.0, wpg);
                        vWp = add(0.0, wpn);
                    }
                    if (want(V_ACC)) st_out(outp(V_ACC, x + 1) + o, cum[q].x);
                    if (want(V_OCEAN)) st_out(outp(V_OCEAN, x + 1) + o, cum[q].y);
                    if (want(V_LEAD)) st_out(outp(V_LEAD, x + 1) + o, vLead);
                    if (want(V_ATM)) st_out(outp(V_ATM, x + 1) + o, vAtm);
                    if (want(V_WPL)) st_out(outp(V_WPL, x + 1) + o, vWpl);
                    if (want(V_WPG)) st_out(outp(V_WPG, x + 1) + o, vWpg);
                    if (want(V_WP)) st_out(outp(V_WP, x + 1) + o, vWp);
                    if (want(V_ADV)) st_out(outp(V_ADV, x + 1) + o, landAdv);
                    if (want(V_DIV)) st_out(outp(V_DIV, x + 1) + o, landAdv);
                    if (want(V_H0)) st_out(outp(V_H0, x + 1) + o, nan);
                    if (want(V_H1)) st_out(outp(V_H1, x + 1) + o, nan);
                    if (want(V_DENS)) st_out(outp(V_DENS, x + 1) + o, nan);
                }
            }
            ENS_TICK(5)   // land stores

            // ---------------- phase B: owned ocean cells -- point-wise terms, 3x3 smoothing, update
            double h0n[KO], h1n[KO];
#pragma unroll
            for (int j = 0; j < KO; ++j) {
                h0n[j] = h1n[j] = 0.0;
                if (own_t[j] < 0) continue;
                const int to = own_t[j];
                const int lr = to / SX - 2, c = to - (lr + 2) * SX;
                const double2 h = hcur[to];
                const double W = f_W[j];
                const double wt = wind_flag(W, mc.wpt);
                const double lead = a.sw.leadloss ? -mul(mul(mul(mul(mul(wt, mc.llf), a.k.deltaT), h.x), W), f_b[j].y) : 0.0;
                const double atm = a.sw.atmloss ? atm_loss(wt, h.x, W, mc, a.k) : 0.0;
                double wpl = 0.0, wpg = 0.0, wpn = 0.0;
                if (a.sw.windpack) wind_packing(wt, h.x, mc, a.k, wpl, wpg, wpn);
                accLead[j] = add(accLead[j], lead);
                accAtm[j] = add(accAtm[j], atm);
                accWpl[j] = add(accWpl[j], wpl);
                accWpg[j] = add(accWpg[j], wpg);
                accWp[j] = add(accWp[j], wpn);
                double t0 = add(add(add(add(h.x, f_b[j].x), wpl), lead), atm);   // NESOSIM.py:327 before the dynamics terms
                double t1 = add(h.y, wpg);                                      // NESOSIM.py:329
                if (a.sw.dynamics) {
                    // astropy tap order: rows outer, columns inner, flipped kernel, accumulators start at 0.0
                    const double2 *pa = s_adv + lr * SXR + c, *pd = s_div + lr * SXR + c;
                    double a0 = 0.0, a1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) {
                            const double wgt = a.w[(2 - ii) * 3 + (2 - jj)];
                            const double2 va = pa[ii * SXR + jj], vd = pd[ii * SXR + jj];
                            a0 = add(a0, mul(va.x, wgt));
                            a1 = add(a1, mul(va.y, wgt));
                            d0 = add(d0, mul(vd.x, wgt));
                            d1 = add(d1, mul(vd.y, wgt));
                        }
                    // smooth_snow's division, then fill_nan_no_negative on an ocean cell (NESOSIM.py:276-284)
                    a0 = mask_nan(div_const(a0, a.conv_div), false, false);
                    a1 = mask_nan(div_const(a1, a.conv_div), false, false);
                    d0 = mask_nan(div_const(d0, a.conv_div), false, false);
                    d1 = mask_nan(div_const(d1, a.conv_div), false, false);
                    accAdv[j] = add(add(accAdv[j], a0), a1);     // NESOSIM.py:290
                    accDiv[j] = add(add(accDiv[j], d0), d1);     // NESOSIM.py:291
                    t0 = add(add(t0, a0), d0);
                    t1 = add(add(t1, a1), d1);
                } else {
                    accAdv[j] = add(add(accAdv[j], 0.0), 0.0);
                    accDiv[j] = add(add(accDiv[j], 0.0), 0.0);
                    t0 = add(add(t0, 0.0), 0.0);
                    t1 = add(add(t1, 0.0), 0.0);
                }
                h0n[j] = mask_nan(t0, false, true);              // NESOSIM.py:332-333
                h1n[j] = mask_nan(t1, false, true);
                put_h(par ^ 1, lr, c, make_double2(h0n[j], h1n[j]));
            }
            if (x <= 1) {   // land: h is NaN from slot 1 on; two steps put it into both tile parities
                for (int i = tid; i < n_land; i += NT) {
                    const int code = land[i];
                    put_h(par ^ 1, code >> 7, code & 127, make_double2(nan, nan));
                }
            }
            // Tiles of day x+1 are written: arrive now, wait after the global stores have been issued, so the
            // barrier's release fence never has this day's HBM stores to drain.
            ENS_TICK(2)   // phase B compute
            cluster_arrive_release();
            ENS_TICK(3)   // arrive (release fence)
            if (STAGE && a.sw.dynamics && x + 1 < steps) stage_day(x + 1);

            // ---------------- outputs of the owned cells (slot x+1)
#pragma unroll
            for (int j = 0; j < KO; ++j) {
                if (own_t[j] < 0) continue;
                const int o = own_o[j];
                if (want(V_ACC)) st_out(outp(V_ACC, x + 1) + o, f_c[j].x);
                if (want(V_OCEAN)) st_out(outp(V_OCEAN, x + 1) + o, f_c[j].y);
                if (want(V_LEAD)) st_out(outp(V_LEAD, x + 1) + o, accLead[j]);
                if (want(V_ATM)) st_out(outp(V_ATM, x + 1) + o, accAtm[j]);
                if (want(V_WPL)) st_out(outp(V_WPL, x + 1) + o, accWpl[j]);
                if (want(V_WPG)) st_out(outp(V_WPG, x + 1) + o, accWpg[j]);
                if (want(V_WP)) st_out(outp(V_WP, x + 1) + o, accWp[j]);
                if (want(V_ADV)) st_out(outp(V_ADV, x + 1) + o, accAdv[j]);
                if (want(V_DIV)) st_out(outp(V_DIV, x + 1) + o, accDiv[j]);
                if (want(V_H0)) st_out(outp(V_H0, x + 1) + o, h0n[j]);
                if (want(V_H1)) st_out(outp(V_H1, x + 1) + o, h1n[j]);
                if (want(V_DENS)) st_out(outp(V_DENS, x + 1) + o, density_variable(h0n[j], h1n[j], false, a.k));
            }

            ENS_TICK(4)   // staging issue + owned-cell stores
            if (STAGE) cp_async_wait_all();
            ENS_TICK(6)   // staged copies landed
            cluster_wait_acquire();   // day x+1 tiles (own rows and pushed halos) are complete; raw tiles are free
            ENS_TICK(7)   // wait for the cluster
        }
    }
#undef ENS_TICK
    if (timing)
        for (int q = 0; q < 8; ++q) a.timing[(long long)blockIdx.x * 8 + q] = tacc[q];
}

// host-side state of this path
struct EnsembleState {
    bool derived_valid = false;
    void *derived = nullptr;           // DA | DB | DC
    size_t derived_bytes = 0;
    unsigned short *codes_dev = nullptr;
    StripTables tables;
    bool tables_ready = false;
    int ko_needed = 0;                 // ceil(max ocean cells per strip / ENS_NT)
    int max_clusters = 0;
};

inline void ensemble_release(EnsembleState &e) {
    cudaFree(e.derived);
    cudaFree(e.codes_dev);
    e.derived = nullptr;
    e.codes_dev = nullptr;
    e.derived_bytes = 0;
    e.tables_ready = false;
}

}  // namespace nesosim

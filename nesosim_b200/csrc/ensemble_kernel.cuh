// ensemble_kernel.cuh -- season-resident path for small grids (the 100 km calibration ensemble).
//
// One thread-block CLUSTER (4 or 8 CTAs) owns one ensemble member for the whole season; each CTA holds a strip of
// rows (strips are cut by the host so that their work balances).  Everything a member needs stays on chip for all
// T-1 days, and HBM only ever sees the output planes, written once, as large contiguous TMA bulk stores:
//   * ten output planes of the strip live in shared memory IN THE OUTPUT LAYOUT (rows x nx doubles, contiguous):
//     h0, h1 (also the state the stencils read), density, snowAdv, snowDiv, snowLead, snowAtm, snowWindPackLoss/
//     Gain/Net.  Each day the owning threads overwrite their cells and one thread issues one
//     cp.async.bulk.global.shared::cta per plane (8-18 KB each).  Measured on B200 (tools/micro): fragmented
//     per-thread stores of this pattern reach 1.6 TB/s, full-row st.global 4.4 TB/s, these bulk stores 5.8 TB/s.
//   * the two boundary rows of each neighbour strip are pushed into double-buffered halo rows through
//     distributed shared memory (st.shared::cluster), one split (arrive ... wait) cluster barrier per day;
//   * the seven member-dependent accumulators are carried in REGISTERS of the thread that owns the cell.
// The land mask is compiled into per-strip cell lists once per context: only ocean cells are owned and advanced;
// only cells with an ocean cell in their 3x3 neighbourhood get raw advection/divergence; land cells (56 % of the
// 100 km grid) have closed-form outputs after the first step -- h and density NaN, every accumulator NaN if its
// switch is on and 0 otherwise (NaN + anything = NaN, x + 0 = x) -- so their plane entries are written twice per
// season and then simply ride along in every bulk store.
// The member-independent forcing is pre-digested once per season by two small kernels (drift gradients;
// snowfall -> accumulation/ocean flux and their running sums) and then shared by every member through L2.
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

constexpr int ENS_MAX_CLUSTER = 8;
constexpr int ENS_NT = 512;            // default CTA size (16 warps); the kernel is templated on it
constexpr int ENS_SXR = 98;            // raw tile row stride; column c lives at c+1 (zero pad each side)
constexpr int ENS_MAX_NX = 96;
constexpr int ENS_MAX_OCEAN = 1024;    // per-strip capacities of the largest kernel variant
constexpr int ENS_MAX_RAW = 1536;
constexpr int ENS_NPLANE = 10;         // staged output planes: h0, h1, density, adv, div, lead, atm, wpl, wpg, wp

// plane index -> output variable
__device__ __constant__ const int ENS_PLANE_VAR[ENS_NPLANE] = {V_H0, V_H1, V_DENS, V_ADV, V_DIV, V_LEAD, V_ATM, V_WPL, V_WPG, V_WP};
enum { PL_H0 = 0, PL_H1, PL_DENS, PL_ADV, PL_DIV, PL_LEAD, PL_ATM, PL_WPL, PL_WPG, PL_WP };

// Shared memory, sized by the host for the tallest strip (`rows`):
//   planes   [10][rows*nx] doubles   output layout; planes 0,1 are also the h the stencils read
//   halos    [2 parity][2 side][2 layer][2 rows][nx] doubles
//   raw adv  [rows+2][SXR] double2, raw div [rows+2][SXR] double2
//   raw and land code lists (uint16)
inline size_t ens_plane_elems(int rows, int nx) { return (size_t)((rows * nx + 1) / 2 * 2); }
inline size_t ens_smem_bytes(int rows, int nx, int n_codes) {
    return (ENS_NPLANE * ens_plane_elems(rows, nx) + 16 * (size_t)nx) * sizeof(double) +
           (size_t)2 * (rows + 2) * ENS_SXR * sizeof(double2) + (size_t)((n_codes + 7) / 8 * 8) * sizeof(unsigned short);
}

// Per-strip cell lists (uint16 code = row*128 + col; row is global for the raw list, strip-local for cells).
struct StripTables {
    const unsigned short *codes;       // all lists concatenated (device)
    int cluster;                       // CTAs per member (4 or 8)
    int row0[ENS_MAX_CLUSTER + 1];     // strip k owns rows row0[k] .. row0[k+1]-1
    int raw_off[ENS_MAX_CLUSTER], raw_int_n[ENS_MAX_CLUSTER], raw_n[ENS_MAX_CLUSTER];   // interior entries first
    int ocean_off[ENS_MAX_CLUSTER], ocean_n[ENS_MAX_CLUSTER];
    int land_off[ENS_MAX_CLUSTER], land_n[ENS_MAX_CLUSTER];
    int rows_alloc;                    // tallest strip
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
// shared -> global bulk copy (TMA), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(double *gdst, const double *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory must be fenced before the async proxy (TMA) reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// NT compute threads per CTA (+ one warp that only issues and drains the bulk stores, so nobody who computes
// ever blocks on the TMA queue); KO owned ocean cells and KR raw-list entries per compute thread.
template <int NT, int KO, int KR>
__global__ void __launch_bounds__(NT + 32, 1) ensemble_season_kernel(const __grid_constant__ EnsArgs a) {
    constexpr int SXR = ENS_SXR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ny = a.ny, nx = a.nx, CL = a.st.cluster;
    const int RA = a.st.rows_alloc;
    const int PE = (RA * nx + 1) / 2 * 2;              // plane stride (doubles)
    double *s_plane = reinterpret_cast<double *>(smem_raw);   // [10][PE]
    double *s_halo = s_plane + ENS_NPLANE * PE;        // [2 parity][2 side: 0 above, 1 below][2 layer][2 rows][nx]
    double2 *s_adv = reinterpret_cast<double2 *>(s_halo + 16 * nx);   // [RA+2][SXR] (adv0,adv1) after NaN->0
    double2 *s_div = s_adv + (RA + 2) * SXR;           // [RA+2][SXR] (div0,div1) after NaN->0
    unsigned short *s_raw_code = reinterpret_cast<unsigned short *>(s_div + (RA + 2) * SXR);
    unsigned short *s_land_code = s_raw_code + (a.st.raw_alloc + 7) / 8 * 8;

    cg::cluster_group cluster = cg::this_cluster();
    const int k = (int)cluster.block_rank();
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const long long plane = (long long)ny * nx;
    const int ra = a.st.row0[k], rb = a.st.row0[k + 1], nrow = rb - ra;
    const int ncell = nrow * nx;
    const int tid = threadIdx.x;
    const bool comp = tid < NT;        // compute thread; the last warp is the store warp
    const bool store_lane = (tid == NT);
    const int steps = a.T - 1;

    // neighbours' halo rows through distributed shared memory
    double *nb_up = (k > 0) ? cluster.map_shared_rank(s_halo, k - 1) : nullptr;        // their side 1 (rows below them)
    double *nb_dn = (k < CL - 1) ? cluster.map_shared_rank(s_halo, k + 1) : nullptr;   // their side 0 (rows above them)
    auto halo_off = [&](int par, int side, int l, int row) { return (((par * 2 + side) * 2 + l) * 2 + row) * nx; };

    const int n_raw_int = a.st.raw_int_n[k], n_raw = a.st.raw_n[k];
    const int n_ocean = a.st.ocean_n[k], n_land = a.st.land_n[k];
    for (int i = tid; i < n_raw; i += NT + 32) s_raw_code[i] = a.st.codes[a.st.raw_off[k] + i];
    for (int i = tid; i < n_land; i += NT + 32) s_land_code[i] = a.st.codes[a.st.land_off[k] + i];
    for (int i = tid; i < (RA + 2) * SXR; i += NT + 32) {   // zero padding of convolve(boundary='fill')
        s_adv[i] = make_double2(0.0, 0.0);
        s_div[i] = make_double2(0.0, 0.0);
    }

    // cells this thread owns for the whole season: ocean list entries tid + j*NT
    int own_lr[KO], own_c[KO];
#pragma unroll
    for (int j = 0; j < KO; ++j) {
        const int idx = tid + j * NT;
        own_lr[j] = -1;
        own_c[j] = 0;
        if (comp && idx < n_ocean) {
            const int code = a.st.codes[a.st.ocean_off[k] + idx];
            own_lr[j] = code >> 7;
            own_c[j] = code & 127;
        }
    }
    // raw-list entries this thread computes every day: tid + q*NT.  For each, the shared-memory offsets (doubles,
    // relative to s_plane, layer 0) of the cell in rows r-1, r, r+1; bit 30 marks a halo row, whose offset moves
    // by one parity block every other day.  Layer 1 sits PE (own rows) or 2*nx (halo rows) further.
    int raw_r[KR], raw_c[KR], raw_up[KR], raw_ce[KR], raw_dn[KR];
    const int HALO0 = ENS_NPLANE * PE;                 // s_halo - s_plane
    constexpr int HALO_FLAG = 1 << 30;
    auto row_off = [&](int r, int c) -> int {          // parity-0 offset of (layer 0, global row r, column c)
        if (r < ra) return (HALO0 + (((0 * 2 + 0) * 2 + 0) * 2 + (r - (ra - 2))) * nx + c) | HALO_FLAG;
        if (r >= rb) return (HALO0 + (((0 * 2 + 1) * 2 + 0) * 2 + (r - rb)) * nx + c) | HALO_FLAG;
        return (r - ra) * nx + c;
    };
#pragma unroll
    for (int q = 0; q < KR; ++q) {
        const int idx = tid + q * NT;
        raw_r[q] = -1;
        raw_c[q] = raw_up[q] = raw_ce[q] = raw_dn[q] = 0;
        if (comp && idx < n_raw) {
            const int code = a.st.codes[a.st.raw_off[k] + idx];
            const int r = code >> 7, c = code & 127;
            raw_r[q] = r;
            raw_c[q] = c;
            raw_up[q] = row_off(r > 0 ? r - 1 : r, c);
            raw_ce[q] = row_off(r, c);
            raw_dn[q] = row_off(r < ny - 1 ? r + 1 : r, c);
        }
    }

    const double nan = qnan();
    // closed-form land values from slot 2 on (slot 1 is computed from the initial depths)
    const double landAdv = a.sw.dynamics ? nan : 0.0, landLead = a.sw.leadloss ? nan : 0.0;
    const double landAtm = a.sw.atmloss ? nan : 0.0, landWp = a.sw.windpack ? nan : 0.0;

    // write (h0,h1) of local row lr, column c: own planes (in place) + the neighbours' halo rows of parity `par`
    auto put_h = [&](int par, int lr, int c, double h0, double h1) {
        s_plane[PL_H0 * PE + lr * nx + c] = h0;
        s_plane[PL_H1 * PE + lr * nx + c] = h1;
        if (lr < 2 && nb_up) {
            nb_up[halo_off(par, 1, 0, lr) + c] = h0;
            nb_up[halo_off(par, 1, 1, lr) + c] = h1;
        }
        if (lr >= nrow - 2 && nb_dn) {
            nb_dn[halo_off(par, 0, 0, lr - (nrow - 2)) + c] = h0;
            nb_dn[halo_off(par, 0, 1, lr - (nrow - 2)) + c] = h1;
        }
    };
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool timing = a.timing != nullptr && tid == 0;
    long long tlast = 0;
#define ENS_TICK(slot)                                   \
    if (timing) {                                        \
        const long long now_ = clock64();                \
        tacc[slot] += now_ - tlast;                      \
        tlast = now_;                                    \
    }

    cluster.sync();   // every CTA of the cluster is running before anyone writes into a neighbour's shared memory

    for (int m = cid; m < a.M; m += ncl) {
        const MemberCoef mc = a.coef[m];
        double accAdv[KO], accDiv[KO], accLead[KO], accAtm[KO], accWpl[KO], accWpg[KO], accWp[KO];
#pragma unroll
        for (int j = 0; j < KO; ++j) accAdv[j] = accDiv[j] = accLead[j] = accAtm[j] = accWpl[j] = accWpg[j] = accWp[j] = 0.0;

        // output address of variable v, time slot `slot` of member m, first cell of this strip
        const long long mo_plane = (long long)m * a.mstride[V_DENS] + (long long)ra * nx;
        const long long mo_depth = (long long)m * a.mstride[V_H0] + (long long)ra * nx;
        auto outp = [&](int v, int slot) -> double * {
            return (v == V_H0 || v == V_H1) ? a.out[v] + (mo_depth + (long long)slot * 2 * plane)
                                            : a.out[v] + (mo_plane + (long long)slot * plane);
        };

        // ---- slot 0: genEmptyArrays zeros + the IC split of main (NESOSIM.py:604-609), every cell of the strip.
        // The previous member's bulk stores may still be reading the planes.
        if (store_lane) bulk_wait_read<0>();
        __syncthreads();
        for (int i = tid; comp && i < ncell; i += NT) {
            const int lr = i / nx, c = i - lr * nx;
            const long long o = (long long)(ra + lr) * nx + c;
            double half = 0.0;
            if (a.ic) {
                double v = a.ic[(long long)m * a.ic_stride + o];
                if (a.conc0[o] < a.k.minConc) v = 0.0;
                half = mul(v, 0.5);
            }
            put_h(0, lr, c, half, half);
#pragma unroll
            for (int v = 0; v < NVAR; ++v)
                if (a.out[v]) outp(v, 0)[i] = (v == V_H0 || v == V_H1) ? half : 0.0;
        }
        cluster.sync();

        // member-independent inputs are always requested one phase ahead, so their L2 latency is never exposed
        double2 d01[KR], d23[KR];   // (ut,vt), (gxu,gyv) of this thread's raw entries
        double2 f_b[KO];            // (acc, 1-C) of the owned cells
        double f_W[KO];
        auto load_raw_inputs = [&](int x) {
            const double2 *DAx = a.DA + (long long)x * plane * 2;
#pragma unroll
            for (int q = 0; q < KR; ++q)
                if (raw_r[q] >= 0) {
                    d01[q] = __ldg(DAx + (raw_r[q] * nx + raw_c[q]) * 2);
                    d23[q] = __ldg(DAx + (raw_r[q] * nx + raw_c[q]) * 2 + 1);
                }
        };
        auto load_cell_inputs = [&](int x) {
#pragma unroll
            for (int j = 0; j < KO; ++j)
                if (own_lr[j] >= 0) {
                    const long long o = (long long)x * plane + (ra + own_lr[j]) * nx + own_c[j];
                    f_b[j] = __ldg(a.DB + o);
                    f_W[j] = __ldg(a.W + o);
                }
        };
#pragma unroll
        for (int q = 0; q < KR; ++q) d01[q] = d23[q] = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < KO; ++j) { f_b[j] = make_double2(0.0, 0.0); f_W[j] = 0.0; }
        if (a.sw.dynamics) load_raw_inputs(0);
        load_cell_inputs(0);

        if (timing) tlast = clock64();
        for (int x = 0; x < steps; ++x) {
            const int par = x & 1;

            // ---------------- phase A: raw advection / divergence (calcDynamics, NESOSIM.py:189-222) where an
            // ocean cell of this strip will read it
            if (a.sw.dynamics) {
                const int hpar = par * 8 * nx;   // halo offset of today's parity block
#pragma unroll
                for (int q = 0; q < KR; ++q) {
                    if (raw_r[q] < 0) continue;
                    const int r = raw_r[q], c = raw_c[q];
                    auto rowp = [&](int off) -> const double * {
                        return s_plane + ((off & HALO_FLAG) ? (off & ~HALO_FLAG) + hpar : off);
                    };
                    auto layer1 = [&](int off) -> int { return (off & HALO_FLAG) ? 2 * nx : PE; };
                    const double *c0 = rowp(raw_ce[q]), *u0 = rowp(raw_up[q]), *w0 = rowp(raw_dn[q]);
                    const double *c1 = c0 + layer1(raw_ce[q]), *u1 = u0 + layer1(raw_up[q]), *w1 = w0 + layer1(raw_dn[q]);
                    const double h0 = c0[0], h1 = c1[0];
                    double gx0, gy0, gx1, gy1;
                    if (tid + q * NT < n_raw_int) {
                        gx0 = div_const(sub(c0[1], c0[-1]), a.g.two_dx);
                        gy0 = div_const(sub(w0[0], u0[0]), a.g.two_dx);
                        gx1 = div_const(sub(c1[1], c1[-1]), a.g.two_dx);
                        gy1 = div_const(sub(w1[0], u1[0]), a.g.two_dx);
                    } else {   // first/last row or column: one-sided differences (np.gradient edge_order=1)
                        const int cm = c > 0 ? -1 : 0, cp = c < nx - 1 ? 1 : 0;
                        gx0 = gradient1d(c0[cm], h0, c0[cp], c, nx, a.g);
                        gy0 = gradient1d(u0[0], h0, w0[0], r, ny, a.g);
                        gx1 = gradient1d(c1[cm], h1, c1[cp], c, nx, a.g);
                        gy1 = gradient1d(u1[0], h1, w1[0], r, ny, a.g);
                    }
                    const int ro = (r - ra + 1) * SXR + c + 1;
                    s_adv[ro] = make_double2(zero_if_nonfinite(adv_term(d01[q].x, d01[q].y, gx0, gy0)),
                                             zero_if_nonfinite(adv_term(d01[q].x, d01[q].y, gx1, gy1)));
                    s_div[ro] = make_double2(zero_if_nonfinite(div_term(h0, d23[q].x, d23[q].y)),
                                             zero_if_nonfinite(div_term(h1, d23[q].x, d23[q].y)));
                }
                if (x + 1 < steps) load_raw_inputs(x + 1);
            }
            ENS_TICK(0)   // phase A
            // the bulk stores of the previous day must have finished READING the planes before they change
            if (store_lane) bulk_wait_read<0>();
            ENS_TICK(1)   // (store warp: bulk stores drained)
            __syncthreads();
            ENS_TICK(2)   // wait for the CTA

            // ---------------- phase B: owned ocean cells -- point-wise terms, 3x3 smoothing, update, plane entries
#pragma unroll
            for (int j = 0; j < KO; ++j) {
                if (own_lr[j] < 0) continue;
                const int lr = own_lr[j], c = own_c[j], ci = lr * nx + c;
                const double h0 = s_plane[PL_H0 * PE + ci], h1 = s_plane[PL_H1 * PE + ci];
                const double W = f_W[j];
                const double wt = wind_flag(W, mc.wpt);
                const double lead = a.sw.leadloss ? -mul(mul(mul(mul(mul(wt, mc.llf), a.k.deltaT), h0), W), f_b[j].y) : 0.0;
                const double atm = a.sw.atmloss ? atm_loss(wt, h0, W, mc, a.k) : 0.0;
                double wpl = 0.0, wpg = 0.0, wpn = 0.0;
                if (a.sw.windpack) wind_packing(wt, h0, mc, a.k, wpl, wpg, wpn);
                accLead[j] = add(accLead[j], lead);
                accAtm[j] = add(accAtm[j], atm);
                accWpl[j] = add(accWpl[j], wpl);
                accWpg[j] = add(accWpg[j], wpg);
                accWp[j] = add(accWp[j], wpn);
                double t0 = add(add(add(add(h0, f_b[j].x), wpl), lead), atm);   // NESOSIM.py:327 before the dynamics terms
                double t1 = add(h1, wpg);                                      // NESOSIM.py:329
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
                const double h0n = mask_nan(t0, false, true), h1n = mask_nan(t1, false, true);   // NESOSIM.py:332-333
                put_h(par ^ 1, lr, c, h0n, h1n);
                s_plane[PL_DENS * PE + ci] = density_variable(h0n, h1n, false, a.k);
                s_plane[PL_ADV * PE + ci] = accAdv[j];
                s_plane[PL_DIV * PE + ci] = accDiv[j];
                s_plane[PL_LEAD * PE + ci] = accLead[j];
                s_plane[PL_ATM * PE + ci] = accAtm[j];
                s_plane[PL_WPL * PE + ci] = accWpl[j];
                s_plane[PL_WPG * PE + ci] = accWpg[j];
                s_plane[PL_WP * PE + ci] = accWp[j];
            }
            // ---------------- land cells.  Step 0 sees the initial depths; afterwards h is NaN, so every switched-on
            // term is NaN and every switched-off term adds 0 (NESOSIM.py:287-322): closed form, written on the
            // first two days (both halo parities) and then left alone in the planes.
            if (x <= 1 && comp) {
                for (int i = tid; i < n_land; i += NT) {
                    const int code = s_land_code[i], lr = code >> 7, c = code & 127, ci = lr * nx + c;
                    double vLead = landLead, vAtm = landAtm, vWpl = landWp, vWpg = landWp, vWp = landWp;
                    if (x == 0) {
                        const long long o = (long long)(ra + lr) * nx + c;
                        const double h0 = s_plane[PL_H0 * PE + ci];
                        const double W = __ldg(a.W + o);
                        const double omc = __ldg(a.DB + o).y;
                        const double wt = wind_flag(W, mc.wpt);
                        vLead = add(0.0, a.sw.leadloss ? -mul(mul(mul(mul(mul(wt, mc.llf), a.k.deltaT), h0), W), omc) : 0.0);
                        vAtm = add(0.0, a.sw.atmloss ? atm_loss(wt, h0, W, mc, a.k) : 0.0);
                        double wpl = 0.0, wpg = 0.0, wpn = 0.0;
                        if (a.sw.windpack) wind_packing(wt, h0, mc, a.k, wpl, wpg, wpn);
                        vWpl = add(0.0, wpl);
                        vWpg = add(0.0, wpg);
                        vWp = add(0.0, wpn);
                    }
                    put_h(par ^ 1, lr, c, nan, nan);
                    s_plane[PL_DENS * PE + ci] = nan;
                    s_plane[PL_ADV * PE + ci] = landAdv;
                    s_plane[PL_DIV * PE + ci] = landAdv;
                    s_plane[PL_LEAD * PE + ci] = vLead;
                    s_plane[PL_ATM * PE + ci] = vAtm;
                    s_plane[PL_WPL * PE + ci] = vWpl;
                    s_plane[PL_WPG * PE + ci] = vWpg;
                    s_plane[PL_WP * PE + ci] = vWp;
                }
            }
            fence_async_smem();
            ENS_TICK(3)   // phase B
            // Planes and pushed halos of day x+1 are written: arrive now, wait at the end of the day.
            cluster_arrive_release();
            __syncthreads();
            ENS_TICK(4)   // arrive + CTA barrier
            if (store_lane) {
                const unsigned bytes = (unsigned)ncell * 8u;
#pragma unroll
                for (int p = 0; p < ENS_NPLANE; ++p) {
                    const int v = ENS_PLANE_VAR[p];
                    if (a.out[v]) bulk_store(outp(v, x + 1), s_plane + p * PE, bytes);
                }
                bulk_commit();
            }
            // snowAcc / snowOcean: member-independent running sums, copied row-contiguously
            if (comp && (a.out[V_ACC] || a.out[V_OCEAN])) {
                const double2 *DCx = a.DC + (long long)x * plane + (long long)ra * nx;
                double *pa_ = a.out[V_ACC] ? outp(V_ACC, x + 1) : nullptr;
                double *po_ = a.out[V_OCEAN] ? outp(V_OCEAN, x + 1) : nullptr;
                for (int base = 0; base < ncell; base += 4 * NT) {   // all loads of a batch in flight before the stores
                    double2 cum[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = base + q * NT + tid;
                        cum[q] = (i < ncell) ? __ldg(DCx + i) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = base + q * NT + tid;
                        if (i < ncell) {
                            if (pa_) __stcs(pa_ + i, cum[q].x);
                            if (po_) __stcs(po_ + i, cum[q].y);
                        }
                    }
                }
            }
            if (x + 1 < steps) load_cell_inputs(x + 1);
            ENS_TICK(5)   // bulk issue + running-sum copies
            cluster_wait_acquire();   // pushed halos of day x+1 are visible; raw tiles are free again
            ENS_TICK(6)   // wait for the cluster
        }
    }
    if (store_lane) bulk_wait_all();
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
    int variant = 0;                   // index into the host's kernel-variant table
    size_t smem_bytes = 0;
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

// ensemble_kernel.cuh -- season-resident path for small grids (the 100 km calibration ensemble).
//
// One thread-block CLUSTER of 4 CTAs owns one ensemble member for the whole season: each CTA holds a strip of
// ~ny/4 rows.  What stays on chip for all T-1 days:
//   * both snow layers h0,h1 of the strip in shared memory (neighbours need them: radius-2 dependency),
//     with the two boundary rows of each neighbour pushed into double-buffered halo rows through distributed
//     shared memory (st.shared::cluster) -- one cluster barrier per day, no global-memory round trip;
//   * the seven member-dependent accumulators (snowAdv, snowDiv, snowLead, snowAtm, snowWindPackLoss/Gain/Net)
//     in REGISTERS of the thread that owns the cell (a fixed R-row x 1-column strip per thread).
// HBM therefore sees only what SURVEY.md §8d counts as algorithmic traffic: the 12 output planes written once
// (96 B per member-cell-day) plus the member-independent forcing, which is pre-digested once per season by two
// small kernels (drift gradients, snowfall -> accumulation/ocean flux and their running sums) and is then
// shared by every member through L2.
//
// Arithmetic is the same per-cell code as the general path (cell_math.cuh), so results are value-identical.
#pragma once
#include <cooperative_groups.h>

#include "cell_math.cuh"
#include "day_kernels.cuh"

namespace nesosim {

namespace cg = cooperative_groups;

enum Derived { D_UT = 0, D_VT, D_GXU, D_GYV, D_ACC, D_OMC, D_SACC, D_SOCE, ND };

// ------------------------------------------------------------------ member-independent pre-pass (per season)

struct DeriveArgs {
    int ny, nx, steps;                 // steps = T-1
    const double *P, *C, *UV;          // [T][plane], [T][plane], [T][2][plane]
    const double *rho_clim;            // unused (variable density only on this path)
    double *D;                         // [steps][ND][plane]
    ModelConsts k;
    GradConsts g;
    ConstDiv rho_new;
};

// ut = U*dT, vt = V*dT, gx(ut), gy(vt) (NESOSIM.py:204-205); acc = (P/rho)*C (260-263); omc = 1-C;
// oc = -((P/rho)*(1-C)) (267) parked in the D_SOCE plane until the scan turns it into the running sum.
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
    double *D = a.D + (long long)x * ND * plane + o;
    const double utc = ut(gy, gx), vtc = vt(gy, gx);
    D[D_UT * plane] = utc;
    D[D_VT * plane] = vtc;
    D[D_GXU * plane] = gradient1d(ut(gy, xm), utc, ut(gy, xp), gx, a.nx, a.g);
    D[D_GYV * plane] = gradient1d(vt(ym, gx), vtc, vt(yp, gx), gy, a.ny, a.g);
    const double C = a.C[(long long)x * plane + o];
    const double pd = div_const(a.P[(long long)x * plane + o], a.rho_new);
    const double omc = sub(1.0, C);
    D[D_ACC * plane] = mul(pd, C);
    D[D_OMC * plane] = omc;
    D[D_SOCE * plane] = -mul(pd, omc);
}

// snowAcc[x+1] = snowAcc[x] + acc, snowOcean[x+1] = snowOcean[x] + oc (NESOSIM.py:264,268): one thread per cell,
// sequential in time, loads batched 8 days ahead so the chain is add-latency bound, not load-latency bound.
__global__ void derive_scan_kernel(double *D, long long plane, int steps) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= plane) return;
    double sa = 0.0, so = 0.0;
    constexpr int B = 8;
    for (int x0 = 0; x0 < steps; x0 += B) {
        double da[B], dq[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int x = min(x0 + j, steps - 1);
            const double *Dx = D + (long long)x * ND * plane + o;
            da[j] = Dx[D_ACC * plane];
            dq[j] = Dx[D_SOCE * plane];
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (x0 + j < steps) {
                sa = add(sa, da[j]);
                so = add(so, dq[j]);
                double *Dx = D + (long long)(x0 + j) * ND * plane + o;
                Dx[D_SACC * plane] = sa;
                Dx[D_SOCE * plane] = so;
            }
        }
    }
}

// ------------------------------------------------------------------------------------ the season kernel

struct EnsArgs {
    int ny, nx, T, M;
    const double *D;                   // derived forcing [T-1][ND][plane]
    const double *W;                   // wind [T][plane]
    const uint8_t *mask;
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
};

constexpr int ENS_CLUSTER = 4;
constexpr int ENS_SX = 96;             // row stride of the h tiles (doubles); columns 0..nx-1
constexpr int ENS_SXR = 98;            // row stride of the raw tiles; column c lives at c+1 (zero pad each side)
constexpr int ENS_MAX_NX = 96;

// R rows per thread, NG row groups per CTA; PSM: keep the five point-wise accumulators (lead, atm, wind-pack
// loss/gain/net) in shared memory instead of registers (trades shared-memory traffic for register pressure).
template <int R, int NG, bool PSM>
struct EnsLayout {
    static constexpr int MAXR = R * NG;                       // rows a CTA can own
    static constexpr int NT = NG * 96;                        // threads: NG row groups x 3 warps (96 columns)
    static constexpr int H_ELEMS = 2 * MAXR * ENS_SX;         // own rows of h0,h1
    static constexpr int HALO_ELEMS = 2 * 2 * 2 * 2 * ENS_SX; // [parity][side][layer][2 rows]
    static constexpr int RAW_ELEMS = 4 * (MAXR + 2) * ENS_SXR;
    static constexpr int PACC_ELEMS = PSM ? 5 * MAXR * ENS_SX : 0;
    static constexpr size_t SMEM_BYTES = (size_t)(H_ELEMS + HALO_ELEMS + RAW_ELEMS + PACC_ELEMS) * sizeof(double);
};

template <int R, int NG, bool PSM>
__global__ void __cluster_dims__(ENS_CLUSTER, 1, 1) __launch_bounds__(NG * 96, 1)
ensemble_season_kernel(const __grid_constant__ EnsArgs a) {
    using L = EnsLayout<R, NG, PSM>;
    constexpr int MAXR = L::MAXR, NT = L::NT, NWARP = NT / 32;
    constexpr int RP = PSM ? 1 : R;              // register copies of the point-wise accumulators
    extern __shared__ __align__(16) double smem[];
    double *s_h = smem;                          // [2][MAXR][SX]
    double *s_halo = s_h + L::H_ELEMS;           // [2 parity][2 side: 0 top,1 bottom][2 layer][2 rows][SX]
    double *s_raw = s_halo + L::HALO_ELEMS;      // [4][MAXR+2][SXR]: adv0, adv1, div0, div1 (after NaN->0)
    double *s_pacc = s_raw + L::RAW_ELEMS;       // [5][MAXR][SX] when PSM

    cg::cluster_group cluster = cg::this_cluster();
    const int k = (int)cluster.block_rank();
    const int cid = blockIdx.x / ENS_CLUSTER, ncl = gridDim.x / ENS_CLUSTER;
    const int ny = a.ny, nx = a.nx;
    const long long plane = (long long)ny * nx;
    const int ra = (k * ny) / ENS_CLUSTER, rb = ((k + 1) * ny) / ENS_CLUSTER, nrow = rb - ra;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // neighbours' halo buffers through distributed shared memory
    double *halo_up = (k > 0) ? cluster.map_shared_rank(s_halo, k - 1) : nullptr;                  // their side 1
    double *halo_dn = (k < ENS_CLUSTER - 1) ? cluster.map_shared_rank(s_halo, k + 1) : nullptr;    // their side 0
    auto halo_off = [](int par, int side, int l, int row) { return (((par * 2 + side) * 2 + l) * 2 + row) * ENS_SX; };

    // this thread's cells: column `col`, local rows lr0 .. lr0+R-1
    const int col = (warp % 3) * 32 + lane;
    const int lr0 = (warp / 3) * R;
    const bool col_ok = col < nx;
    unsigned land_bits = 0, valid_bits = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (col_ok && lr0 + i < nrow) {
            valid_bits |= 1u << i;
            if (is_land(a.mask[(long long)(ra + lr0 + i) * nx + col])) land_bits |= 1u << i;
        }
    }

    for (int i = tid; i < L::RAW_ELEMS; i += NT) s_raw[i] = 0.0;   // zero padding of convolve(boundary='fill')

    const int r_lo = max(ra - 1, 0), r_hi = min(rb, ny - 1);       // raw rows this CTA needs
    const int nraw = r_hi - r_lo + 1;
    const int steps = a.T - 1;
    cluster.sync();   // every CTA of the cluster is running before anyone writes into a neighbour's shared memory

    for (int m = cid; m < a.M; m += ncl) {
        const MemberCoef mc = a.coef[m];
        double accAdv[R], accDiv[R], accLead[RP], accAtm[RP], accWpl[RP], accWpg[RP], accWp[RP];

        auto outp = [&](int v, int slot) -> double * {
            if (!a.out[v]) return nullptr;
            const long long per_slot = (v == V_H0 || v == V_H1) ? 2 * plane : plane;
            return a.out[v] + (long long)m * a.mstride[v] + (long long)slot * per_slot;
        };
        auto put_h = [&](int par, int l, int lr, double val) {
            s_h[(l * MAXR + lr) * ENS_SX + col] = val;
            if (lr < 2 && halo_up) halo_up[halo_off(par, 1, l, lr) + col] = val;
            if (lr >= nrow - 2 && halo_dn) halo_dn[halo_off(par, 0, l, lr - (nrow - 2)) + col] = val;
        };

        // ---- slot 0: genEmptyArrays zeros + the IC split of main (NESOSIM.py:604-609).  No barrier is needed
        // against the previous member: its last day ended with a cluster barrier after every halo read.
#pragma unroll
        for (int i = 0; i < R; ++i) {
            accAdv[i] = accDiv[i] = 0.0;
            if (!PSM) accLead[i % RP] = accAtm[i % RP] = accWpl[i % RP] = accWpg[i % RP] = accWp[i % RP] = 0.0;
            if (!(valid_bits >> i & 1)) continue;
            if (PSM) {
#pragma unroll
                for (int q = 0; q < 5; ++q) s_pacc[(q * MAXR + lr0 + i) * ENS_SX + col] = 0.0;
            }
            const long long o = (long long)(ra + lr0 + i) * nx + col;
            double half = 0.0;
            if (a.ic) {
                double v = a.ic[(long long)m * a.ic_stride + o];
                if (a.conc0[o] < a.k.minConc) v = 0.0;
                half = mul(v, 0.5);
            }
            put_h(0, 0, lr0 + i, half);
            put_h(0, 1, lr0 + i, half);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) {
                double *p = outp(v, 0);
                if (p) p[o] = (v == V_H0 || v == V_H1) ? half : 0.0;
            }
        }
        cluster.sync();

        for (int x = 0; x < steps; ++x) {
            const int par = x & 1;
            const double *Dx = a.D + (long long)x * ND * plane;

            // member-independent inputs of this thread's cells, requested before phase A so L2 latency overlaps it
            double f_acc[R], f_omc[R], f_W[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                f_acc[i] = f_omc[i] = f_W[i] = 0.0;
                if (valid_bits >> i & 1) {
                    const long long o = (long long)(ra + lr0 + i) * nx + col;
                    f_acc[i] = __ldg(Dx + D_ACC * plane + o);
                    f_omc[i] = __ldg(Dx + D_OMC * plane + o);
                    f_W[i] = __ldg(a.W + (long long)x * plane + o);
                }
            }

            // ---------------- phase A: raw advection / divergence on own rows +-1 (calcDynamics, NESOSIM.py:189-222)
            if (a.sw.dynamics) {
                auto hrow = [&](int l, int r) -> const double * {
                    if (r < ra) return s_halo + halo_off(par, 0, l, r - (ra - 2));
                    if (r >= rb) return s_halo + halo_off(par, 1, l, r - rb);
                    return s_h + (l * MAXR + (r - ra)) * ENS_SX;
                };
                for (int t = warp; t < nraw * 3; t += NWARP) {
                    const int r = r_lo + t / 3;
                    const int c = (t % 3) * 32 + lane;
                    if (c >= nx) continue;
                    const long long o = (long long)r * nx + c;
                    const double ut = __ldg(Dx + D_UT * plane + o), vt = __ldg(Dx + D_VT * plane + o);
                    const double gxu = __ldg(Dx + D_GXU * plane + o), gyv = __ldg(Dx + D_GYV * plane + o);
                    const int cm = max(c - 1, 0), cp = min(c + 1, nx - 1);
                    const int rm = max(r - 1, 0), rp = min(r + 1, ny - 1);
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        const double *hc = hrow(l, r);
                        const double h = hc[c];
                        const double gxh = gradient1d(hc[cm], h, hc[cp], c, nx, a.g);
                        const double gyh = gradient1d(hrow(l, rm)[c], h, hrow(l, rp)[c], r, ny, a.g);
                        const int ro = (r - (ra - 1)) * ENS_SXR + c + 1;
                        s_raw[(l * (MAXR + 2)) * ENS_SXR + ro] = zero_if_nonfinite(adv_term(ut, vt, gxh, gyh));
                        s_raw[((2 + l) * (MAXR + 2)) * ENS_SXR + ro] = zero_if_nonfinite(div_term(h, gxu, gyv));
                    }
                }
            }
            __syncthreads();

            // ---------------- phase B: point-wise terms, 3x3 smoothing, update, outputs
            double t0[R], t1[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                t0[i] = t1[i] = 0.0;
                if (!(valid_bits >> i & 1)) continue;
                const long long o = (long long)(ra + lr0 + i) * nx + col;
                const double h0 = s_h[(0 * MAXR + lr0 + i) * ENS_SX + col];
                const double h1 = s_h[(1 * MAXR + lr0 + i) * ENS_SX + col];
                const double W = f_W[i];
                const double wt = wind_flag(W, mc.wpt);
                const double lead = a.sw.leadloss ? -mul(mul(mul(mul(mul(wt, mc.llf), a.k.deltaT), h0), W), f_omc[i]) : 0.0;
                const double atm = a.sw.atmloss ? atm_loss(wt, h0, W, mc, a.k) : 0.0;
                double wpl = 0.0, wpg = 0.0, wpn = 0.0;
                if (a.sw.windpack) wind_packing(wt, h0, mc, a.k, wpl, wpg, wpn);
                const int q = i % RP;
                double *sp = s_pacc + (lr0 + i) * ENS_SX + col;
                if (PSM) {
                    accLead[q] = sp[0 * MAXR * ENS_SX];
                    accAtm[q] = sp[1 * MAXR * ENS_SX];
                    accWpl[q] = sp[2 * MAXR * ENS_SX];
                    accWpg[q] = sp[3 * MAXR * ENS_SX];
                    accWp[q] = sp[4 * MAXR * ENS_SX];
                }
                accLead[q] = add(accLead[q], lead);
                accAtm[q] = add(accAtm[q], atm);
                accWpl[q] = add(accWpl[q], wpl);
                accWpg[q] = add(accWpg[q], wpg);
                accWp[q] = add(accWp[q], wpn);
                if (PSM) {
                    sp[0 * MAXR * ENS_SX] = accLead[q];
                    sp[1 * MAXR * ENS_SX] = accAtm[q];
                    sp[2 * MAXR * ENS_SX] = accWpl[q];
                    sp[3 * MAXR * ENS_SX] = accWpg[q];
                    sp[4 * MAXR * ENS_SX] = accWp[q];
                }
                double *p;
                if ((p = outp(V_ACC, x + 1))) p[o] = __ldg(Dx + D_SACC * plane + o);     // member-independent sums
                if ((p = outp(V_OCEAN, x + 1))) p[o] = __ldg(Dx + D_SOCE * plane + o);
                if ((p = outp(V_LEAD, x + 1))) p[o] = accLead[q];
                if ((p = outp(V_ATM, x + 1))) p[o] = accAtm[q];
                if ((p = outp(V_WPL, x + 1))) p[o] = accWpl[q];
                if ((p = outp(V_WPG, x + 1))) p[o] = accWpg[q];
                if ((p = outp(V_WP, x + 1))) p[o] = accWp[q];
                t0[i] = add(add(add(add(h0, f_acc[i]), wpl), lead), atm);   // NESOSIM.py:327 up to the dynamics terms
                t1[i] = add(h1, wpg);                                      // NESOSIM.py:329
            }
            if (a.sw.dynamics) {
                // planes in the order the reference adds them: adv0, adv1 (snowAdv), div0, div1 (snowDiv)
#pragma unroll
                for (int p4 = 0; p4 < 4; ++p4) {
                    const double *rp = s_raw + (p4 * (MAXR + 2) + lr0) * ENS_SXR + col;   // raw row lr0-1, column col-1
                    double w0[3], w1[3], w2[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) { w0[j] = rp[j]; w1[j] = rp[ENS_SXR + j]; }
#pragma unroll
                    for (int i = 0; i < R; ++i) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) w2[j] = rp[(i + 2) * ENS_SXR + j];
                        double top = 0.0;
#pragma unroll
                        for (int j = 0; j < 3; ++j) top = add(top, mul(w0[j], a.w[8 - j]));
#pragma unroll
                        for (int j = 0; j < 3; ++j) top = add(top, mul(w1[j], a.w[5 - j]));
#pragma unroll
                        for (int j = 0; j < 3; ++j) top = add(top, mul(w2[j], a.w[2 - j]));
                        const double sm = mask_nan(div_const(top, a.conv_div), land_bits >> i & 1, false);
                        if (p4 == 0) { accAdv[i] = add(accAdv[i], sm); t0[i] = add(t0[i], sm); }
                        if (p4 == 1) { accAdv[i] = add(accAdv[i], sm); t1[i] = add(t1[i], sm); }
                        if (p4 == 2) { accDiv[i] = add(accDiv[i], sm); }
                        if (p4 == 3) { accDiv[i] = add(accDiv[i], sm); }
                        if (p4 == 2) t0[i] = add(t0[i], sm);
                        if (p4 == 3) t1[i] = add(t1[i], sm);
#pragma unroll
                        for (int j = 0; j < 3; ++j) { w0[j] = w1[j]; w1[j] = w2[j]; }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; ++i) {   // + zeros, as the reference adds them (NESOSIM.py:287-291,327-329)
                    accAdv[i] = add(add(accAdv[i], 0.0), 0.0);
                    accDiv[i] = add(add(accDiv[i], 0.0), 0.0);
                    t0[i] = add(add(t0[i], 0.0), 0.0);
                    t1[i] = add(add(t1[i], 0.0), 0.0);
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (!(valid_bits >> i & 1)) continue;
                const long long o = (long long)(ra + lr0 + i) * nx + col;
                const bool land = land_bits >> i & 1;
                const double h0n = mask_nan(t0[i], land, true);
                const double h1n = mask_nan(t1[i], land, true);
                put_h(par ^ 1, 0, lr0 + i, h0n);
                put_h(par ^ 1, 1, lr0 + i, h1n);
                double *p;
                if ((p = outp(V_ADV, x + 1))) p[o] = accAdv[i];
                if ((p = outp(V_DIV, x + 1))) p[o] = accDiv[i];
                if ((p = outp(V_H0, x + 1))) p[o] = h0n;
                if ((p = outp(V_H1, x + 1))) p[o] = h1n;
                if ((p = outp(V_DENS, x + 1))) p[o] = density_variable(h0n, h1n, land, a.k);
            }
            cluster.sync();   // halos of day x+1 are visible; raw tile and own rows are free for the next day
        }
    }
}

// host-side state of this path
struct EnsembleState {
    bool derived_valid = false;
    double *derived = nullptr;
    size_t derived_elems = 0;
    int max_clusters = 0;
};

inline void ensemble_release(EnsembleState &e) {
    cudaFree(e.derived);
    e.derived = nullptr;
    e.derived_elems = 0;
}

}  // namespace nesosim

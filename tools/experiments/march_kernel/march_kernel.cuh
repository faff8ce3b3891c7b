// march_kernel.cuh -- the day step for LARGE grids (the 5 km pan-Arctic grid and its row strips): one launch per day
// like day_step_kernel, same arguments, same values, but organised as a streaming pipeline instead of tiles.
//
// Why a second kernel.  The tile kernel (day_kernels.cuh) gives every 32x16 tile to a short-lived CTA: load, barrier,
// raw dynamics, barrier, budget, store.  On a grid of thousands of tiles that costs (measured, round 1): 41 % more
// depth/drift cells loaded than owned (two-cell halo on all four sides), 20 % more raw dynamics computed than used,
// 256-byte row fragments per store instruction on rows that are only 8-byte aligned (partial sectors at both ends of
// every fragment), a dependent load -> compute -> store chain per CTA with two CTAs per SM to hide it, and ~1000
// instructions per warp and ocean cell.  Here instead:
//   * a WARP owns a strip of 62 columns and MARCHES down its rows (a "chain"); the two-cell halo exists only sideways
//     (64 raw-dynamics columns and 66 depth/drift columns for 62 owned ones: 3 % / 6 % instead of 20 % / 41 %);
//     vertically, the three raw rows and four depth/drift rows a row needs roll through a small per-warp ring in shared
//     memory, every row is loaded and its raw dynamics computed exactly once;
//   * warps are autonomous -- no CTA barrier anywhere, only __syncwarp -- so the twelve warps of an SM drift apart and
//     some are always waiting on memory while others compute: the loads of row y+3 and the point-wise inputs of row y
//     are in flight while the raw dynamics of row y+1 and the budget of row y run;
//   * a lane handles columns lane and lane+32: every load/store instruction of a warp is one contiguous 256-byte run and
//     the two of a row together 496 contiguous bytes per plane;
//   * rows whose output cells are all land (57 % of the polar grid is land, in large blocks) take a closed-form path
//     and skip the depth/drift loads and the raw dynamics at ROW granularity (per strip and row, compiled from the mask
//     once per context), not only for whole 32x16 tiles;
//   * the constant divisions of the stencils run as the bare three-operation sequence; their operand window is guarded
//     once per loaded depth / drift value when the row is staged (cell_math.cuh out_of_guard), and a row that fails the
//     guard -- never on physical data -- is redone by the same warp with the fully general divisions;
//   * chains are cut on the host so that every warp of the launch gets the same number of (modelled) bytes; rows within
//     two cells of the grid edge (one-sided differences, zero padding) and the first / last strip run a general,
//     bounds-checked instantiation of the same code.
// Reference semantics are those of day_kernels.cuh: calcDynamics NESOSIM.py:189-222, smooth_snow :170-187,
// calcBudget :260-347, densityCalc :458-473; arithmetic from cell_math.cuh, so results are value-identical.
#pragma once
#include "day_kernels.cuh"

namespace nesosim {

// CPL = columns per lane (1 or 2): a warp strip has 32*CPL raw-dynamics columns, of which the inner 32*CPL-2 are owned
// (3x3 smoothing), and 32*CPL+2 depth / drift columns (gradient of the outermost raw columns).  CPL = 2: 62-column
// strips, 496 contiguous bytes per plane and row, twelve warps per SM; CPL = 1: 30-column strips, half the registers and
// shared memory per warp, twenty warps per SM.
constexpr int march_owned(int cpl) { return 32 * cpl - 2; }
constexpr int MARCH_WARPS = 4;    // warps per CTA
constexpr int MARCH_CAP = 4;      // the first / last MARCH_CAP rows of every strip are chains of their own

enum { MC_EDGE = 1, MC_MAIL_TOP = 2, MC_MAIL_BOT = 4, MC_READS_MAIL = 8 };
enum { RF_OCEAN = 1, RF_RAW = 2, RF_H = 4, RF_IN = 8 };
constexpr int MARCH_PAD = 4;      // row descriptors exist for rows -MARCH_PAD .. ny-1+MARCH_PAD (flags 0 outside the grid)

struct MarchChain {
    int strip, y0, y1, flags;     // output rows [y0, y1) of strip `strip`
};

struct MarchArgs {
    // Row descriptors [nstrips][ny + 2*MARCH_PAD] (row r at index r + MARCH_PAD): .x/.y = land bits of the strip's 64
    // raw-dynamics columns (bit i: column c0-1+i is land, lake or outside the grid), .z = flags: RF_OCEAN some owned cell
    // of the row is ocean; RF_RAW / RF_H the raw row / the depth+drift row is read by some ocean cell's stencil; RF_IN the
    // row is inside the grid.  Warps fetch them a full row ahead of their use: nothing on a row's critical path depends
    // on a load of the mask.
    const uint4 *rowdesc;
    const MarchChain *chains;
    const int *worker_first;      // [nworkers + 1]: chains of worker (warp) w are worker_first[w] .. worker_first[w+1]-1
    int nworkers;
    int shortcut;                 // slot x was written by this library in this call: depths on land are NaN
};

template <int CPL>
struct MarchWarpSmem {
    static constexpr int MRAW = 32 * CPL, MHS = 32 * CPL + 4;
    double h[4][2][MHS];          // depth rows r & 3, both layers; column index = gx - (c0 - 2)
    double ut[4][MHS], vt[4][MHS];   // drift * deltaT
    double2 ra[4][MRAW];          // raw advection (layer 0, layer 1) after the NaN/inf -> 0 fill; column index = gx - (c0 - 1)
    double2 rd[4][MRAW];          // raw divergence
};

template <bool STRIP, int CPL>
struct MarchCtx {
    static constexpr int MW = march_owned(CPL);
    const DayArgs &a;
    const MarchArgs &q;
    const StripLink &s;
    MarchWarpSmem<CPL> &S;
    const int lane, m;
    const long long moffp, moffd;
    const MemberCoef mc;
    unsigned hbad, rbad;          // per ring slot: the staged row / the raw row failed the operand guard
    const uint4 *rd;              // this strip's descriptors, row 0
    int c0;

    __device__ __forceinline__ uint4 desc(int r) const { return __ldg(rd + r); }
    // before this library's own first step of a call the depths on land may be finite: every row inside the grid then
    // runs the full arithmetic
    __device__ __forceinline__ unsigned eff(unsigned f) const { return q.shortcut ? f : ((f & RF_IN) ? (f | 7u) : 0u); }

    // ---- staging of one depth / drift row: asynchronous copies straight into the ring slot (issued a step early, no
    // registers held) ...
    template <bool EDGE>
    __device__ __forceinline__ void stage_issue(int r) {
        const int nx = a.nx, slot = r & 3;
        const long long ro = (long long)r * nx + (c0 - 2);
        const double *p0 = a.prev[V_H0] + moffd + ro, *p1 = a.prev[V_H1] + moffd + ro;
        const double *pu = a.U + ro, *pv = a.V + ro;
        bool mail = false;
        if (STRIP && s.use_mail) {      // ghost rows of a neighbouring strip: depths come from the mailbox
            const double *mb = nullptr;
            int mrow = 0;
            if (s.has_up && r < STRIP_GHOST) { mb = s.mail_top; mrow = r; }
            else if (s.has_dn && r >= a.ny - STRIP_GHOST) { mb = s.mail_bot; mrow = r - (a.ny - STRIP_GHOST); }
            if (mb) {
                mail = true;
                p0 = mb + ((long long)(a.x & 1) * 2 * STRIP_GHOST + mrow) * nx + (c0 - 2);
                p1 = p0 + (long long)STRIP_GHOST * nx;
            }
        }
        // every lane copies columns lane and lane+32, lanes 0 and 1 also 64 and 65; columns outside the grid are clamped
        // into it -- whatever they hold is never used: cells outside the grid get zero raw dynamics, and the one-sided
        // differences at the grid edge ignore the outer neighbour
#pragma unroll
        for (int k = 0; k <= CPL; ++k) {
            const int hc = lane + 32 * k;
            if (k == CPL && lane >= 2) break;
            int sc = hc;
            if (EDGE) sc = min(max(hc, 2 - c0), nx - 1 - (c0 - 2));
            if (STRIP && mail) {        // peer-written memory: read through L2 (the neighbour GPU stored it there)
                S.h[slot][0][hc] = __ldcg(p0 + sc);
                S.h[slot][1][hc] = __ldcg(p1 + sc);
            } else {
                cp_async8(&S.h[slot][0][hc], p0 + sc, true);
                cp_async8(&S.h[slot][1][hc], p1 + sc, true);
            }
            cp_async8(&S.ut[slot][hc], pu + sc, true);
            cp_async8(&S.vt[slot][hc], pv + sc, true);
        }
    }
    // ... and the commit once they have landed: driftGday*deltaT (NESOSIM.py:204,210) is formed once, in place, and every
    // value is checked against the operand guard of the bare constant divisions
    __device__ __forceinline__ void stage_commit(int r) {
        const int slot = r & 3;
        cp_async_wait_all();
        unsigned bad = 0;
#pragma unroll
        for (int k = 0; k <= CPL; ++k) {
            const int hc = lane + 32 * k;
            if (k == CPL && lane >= 2) break;
            const double ut = mul(S.ut[slot][hc], a.k.deltaT), vt = mul(S.vt[slot][hc], a.k.deltaT);
            S.ut[slot][hc] = ut;
            S.vt[slot][hc] = vt;
            bad |= out_of_guard(S.h[slot][0][hc]) | out_of_guard(S.h[slot][1][hc]) | out_of_guard(ut) | out_of_guard(vt);
        }
        const unsigned any = __any_sync(0xffffffffu, bad) ? 1u : 0u;
        hbad = (hbad & ~(1u << slot)) | (any << slot);
    }

    // ---- raw advection / divergence of row r (calcDynamics + fillMaskAndNaNWithZero) into the raw ring
    template <bool EDGE>
    __device__ __forceinline__ void raw_row(int r, bool needed) {
        const int slot = r & 3, up = (r - 1) & 3, dn = (r + 1) & 3;
        const int ny = a.ny, nx = a.nx;
        rbad &= ~(1u << slot);
        if (r < 0 || r >= ny) {            // outside the grid: the zero padding of convolve(boundary='fill')
            if (EDGE) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    S.ra[slot][lane + 32 * j] = make_double2(0.0, 0.0);
                    S.rd[slot][lane + 32 * j] = make_double2(0.0, 0.0);
                }
            }
            return;
        }
        if (!needed) return;               // no ocean cell's stencil reads this row
        const bool guard_failed = (((hbad >> up) | (hbad >> slot) | (hbad >> dn)) & 1u) != 0u;
        if (!EDGE && guard_failed) rbad |= 1u << slot;
        const bool fast = !EDGE && !guard_failed;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int i = lane + 32 * j, hc = i + 1;
            const int gx = c0 - 1 + i;
            double2 adv = make_double2(0.0, 0.0), dv = make_double2(0.0, 0.0);
            if (!EDGE || (gx >= 0 && gx < nx)) {
                const double ut = S.ut[slot][hc], vt = S.vt[slot][hc];
                double gxu, gyv, gxh[2], gyh[2];
                if (fast) {
                    gxu = div_const_bare(sub(S.ut[slot][hc + 1], S.ut[slot][hc - 1]), a.g.two_dx);
                    gyv = div_const_bare(sub(S.vt[dn][hc], S.vt[up][hc]), a.g.two_dx);
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        gxh[l] = div_const_bare(sub(S.h[slot][l][hc + 1], S.h[slot][l][hc - 1]), a.g.two_dx);
                        gyh[l] = div_const_bare(sub(S.h[dn][l][hc], S.h[up][l][hc]), a.g.two_dx);
                    }
                } else if (!EDGE) {
                    gxu = div_const(sub(S.ut[slot][hc + 1], S.ut[slot][hc - 1]), a.g.two_dx);
                    gyv = div_const(sub(S.vt[dn][hc], S.vt[up][hc]), a.g.two_dx);
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        gxh[l] = div_const(sub(S.h[slot][l][hc + 1], S.h[slot][l][hc - 1]), a.g.two_dx);
                        gyh[l] = div_const(sub(S.h[dn][l][hc], S.h[up][l][hc]), a.g.two_dx);
                    }
                } else {                   // np.gradient's one-sided differences at the grid edge
                    gxu = gradient1d(S.ut[slot][hc - 1], ut, S.ut[slot][hc + 1], gx, nx, a.g);
                    gyv = gradient1d(S.vt[up][hc], vt, S.vt[dn][hc], r, ny, a.g);
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        gxh[l] = gradient1d(S.h[slot][l][hc - 1], S.h[slot][l][hc], S.h[slot][l][hc + 1], gx, nx, a.g);
                        gyh[l] = gradient1d(S.h[up][l][hc], S.h[slot][l][hc], S.h[dn][l][hc], r, ny, a.g);
                    }
                }
                const double h0 = S.h[slot][0][hc], h1 = S.h[slot][1][hc];
                adv = make_double2(zero_if_nonfinite(adv_term(ut, vt, gxh[0], gyh[0])), zero_if_nonfinite(adv_term(ut, vt, gxh[1], gyh[1])));
                dv = make_double2(zero_if_nonfinite(div_term(h0, gxu, gyv)), zero_if_nonfinite(div_term(h1, gxu, gyv)));
            }
            S.ra[slot][i] = adv;
            S.rd[slot][i] = dv;
        }
    }

    // ---- the point-wise inputs of output row y, requested before the raw dynamics of row y+1 are computed and first
    // touched behind the smoothing of row y.  Straight-line code: every lane loads every value into a register of its own
    // (lanes outside the strip's owned columns load a neighbouring column and store nothing), so no load is followed by
    // a conditional overwrite of its destination -- which would wait for the load on the spot.
    struct Point {
        double P, C, W, prev[9];
    };
    template <bool EDGE>
    __device__ __forceinline__ long long cell_of(int y, int j) const {      // (clamped into the grid for the loads)
        int gx = c0 - 1 + lane + 32 * j;
        if (EDGE) gx = min(max(gx, 0), a.nx - 1);
        return (long long)y * a.nx + gx;
    }
    __device__ __forceinline__ bool owned(int j, bool edge) const {
        const int i = lane + 32 * j;
        return i >= 1 && i <= MW && (!edge || c0 - 1 + i < a.nx);
    }
    template <bool EDGE>
    __device__ __forceinline__ void point_load_ocean(int y, Point (&pt)[CPL]) const {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const long long o = cell_of<EDGE>(y, j), ip = moffp + o;
            pt[j].P = __ldg(a.P + o);
            pt[j].C = __ldg(a.C + o);
            pt[j].W = __ldg(a.W + o);
#pragma unroll
            for (int v = 0; v < 9; ++v) pt[j].prev[v] = a.prev[V_ACC + v][ip];
        }
    }
    // a row whose owned cells are all land, after this library's own step: the depths are NaN, only snowAcc / snowOcean
    // (and accumulators whose switch is off, which add a literal zero) still need their previous value
    template <bool EDGE>
    __device__ __forceinline__ void point_load_land(int y, Point (&pt)[CPL]) const {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const long long o = cell_of<EDGE>(y, j), ip = moffp + o;
            pt[j].P = __ldg(a.P + o);
            pt[j].C = __ldg(a.C + o);
            pt[j].prev[V_ACC - V_ACC] = a.prev[V_ACC][ip];
            pt[j].prev[V_OCEAN - V_ACC] = a.prev[V_OCEAN][ip];
            if (!a.sw.leadloss) pt[j].prev[V_LEAD - V_ACC] = a.prev[V_LEAD][ip];
            if (!a.sw.atmloss) pt[j].prev[V_ATM - V_ACC] = a.prev[V_ATM][ip];
            if (!a.sw.windpack) {
                pt[j].prev[V_WPL - V_ACC] = a.prev[V_WPL][ip];
                pt[j].prev[V_WPG - V_ACC] = a.prev[V_WPG][ip];
                pt[j].prev[V_WP - V_ACC] = a.prev[V_WP][ip];
            }
        }
    }

    __device__ __forceinline__ void mail_store(int y, int gx, double h0n, double h1n) const {
        if (STRIP) {      // first / last owned rows -> the neighbour's mailbox for slot x+1
            const int nx = a.nx, ny = a.ny;
            const long long par = (long long)((a.x + 1) & 1) * 2 * STRIP_GHOST;
            if (s.has_up && y >= STRIP_GHOST && y < 2 * STRIP_GHOST) {
                const long long mo = (par + (y - STRIP_GHOST)) * nx + gx;
                s.peer_up_mail[mo] = h0n;
                s.peer_up_mail[mo + (long long)STRIP_GHOST * nx] = h1n;
            }
            if (s.has_dn && y >= ny - 2 * STRIP_GHOST && y < ny - STRIP_GHOST) {
                const long long mo = (par + (y - (ny - 2 * STRIP_GHOST))) * nx + gx;
                s.peer_dn_mail[mo] = h0n;
                s.peer_dn_mail[mo + (long long)STRIP_GHOST * nx] = h1n;
            }
        }
    }

    // ---- output row y with ocean cells: 3x3 smoothing of the raw planes, the budget terms, the twelve stores
    // (calcBudget).  Land cells of the row run the same arithmetic -- with NaN depths every masked or h-dependent term
    // comes out NaN exactly as the reference's own 0*NaN algebra does.
    template <bool EDGE>
    __device__ __forceinline__ void out_row_ocean(int y, unsigned lb_lo, unsigned lb_hi, const Point (&pt)[CPL]) const {
        const int nx = a.nx;
        const int s0 = (y - 1) & 3, s1 = y & 3, s2 = (y + 1) & 3;
        const bool general = EDGE || ((((rbad >> s0) | (rbad >> s1) | (rbad >> s2)) & 1u) != 0u);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const Point &p = pt[j];
            const bool own = owned(j, EDGE);
            const int i = min(max(lane + 32 * j, 1), MW);       // (lanes outside the owned columns compute a neighbour's cell)
            const int gx = c0 - 1 + lane + 32 * j;
            const bool land = (((j == 0 ? lb_lo : lb_hi) >> lane) & 1u) != 0u;     // (bit 32*j + lane of the row's land bits)
            const long long o = (long long)y * nx + gx;
            const long long ip = moffp + o, id = moffd + o;
            auto prev = [&](int v) { return p.prev[v - V_ACC]; };
            auto store = [&](int v, double val) {       // only the density output is optional
                if (own && (v != V_DENS || a.next[v])) a.next[v][(v == V_H0 || v == V_H1) ? id : ip] = val;
            };
            // astropy tap order: rows outer, columns inner, flipped kernel, accumulators start at 0.0
            double a0 = 0.0, a1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) {
                const int sl = (ii == 0) ? s0 : (ii == 1 ? s1 : s2);
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const double wgt = a.w[(2 - ii) * 3 + (2 - jj)];
                    const double2 va = S.ra[sl][i - 1 + jj], vd = S.rd[sl][i - 1 + jj];
                    a0 = add(a0, mul(va.x, wgt));
                    a1 = add(a1, mul(va.y, wgt));
                    d0 = add(d0, mul(vd.x, wgt));
                    d1 = add(d1, mul(vd.y, wgt));
                }
            }
            double adv0, adv1, div0, div1;
            if (general) {
                adv0 = div_const(a0, a.conv_div); adv1 = div_const(a1, a.conv_div);
                div0 = div_const(d0, a.conv_div); div1 = div_const(d1, a.conv_div);
            } else {
                adv0 = div_const_bare(a0, a.conv_div); adv1 = div_const_bare(a1, a.conv_div);
                div0 = div_const_bare(d0, a.conv_div); div1 = div_const_bare(d1, a.conv_div);
            }
            adv0 = mask_nan(adv0, land, false);               // NESOSIM.py:276-284
            adv1 = mask_nan(adv1, land, false);
            div0 = mask_nan(div0, land, false);
            div1 = mask_nan(div1, land, false);
            const double h0 = S.h[s1][0][i + 1], h1 = S.h[s1][1][i + 1];
            // (the loaded point-wise inputs are first touched here, behind the smoothing: their latency is covered)
            const double pd = div_const(p.P, a.rho_new);           // precipDayT/snowDensityNew (NESOSIM.py:260)
            const double acc = mul(pd, p.C);                       // NESOSIM.py:263
            const double oc = -mul(pd, sub(1.0, p.C));             // NESOSIM.py:267
            const double W = p.W, C = p.C;
            const double wt = wind_flag(W, mc.wpt);
            const double lead = a.sw.leadloss ? lead_loss(wt, h0, W, C, mc, a.k) : 0.0;
            const double atm = a.sw.atmloss ? atm_loss(wt, h0, W, mc, a.k) : 0.0;
            double wpl = 0.0, wpg = 0.0, wpn = 0.0;
            if (a.sw.windpack) wind_packing(wt, h0, mc, a.k, wpl, wpg, wpn);
            store(V_ACC, add(prev(V_ACC), acc));
            store(V_OCEAN, add(prev(V_OCEAN), oc));
            store(V_ADV, add(add(prev(V_ADV), adv0), adv1));     // NESOSIM.py:290
            store(V_DIV, add(add(prev(V_DIV), div0), div1));     // NESOSIM.py:291
            store(V_LEAD, add(prev(V_LEAD), lead));
            store(V_ATM, add(prev(V_ATM), atm));
            store(V_WPL, add(prev(V_WPL), wpl));
            store(V_WPG, add(prev(V_WPG), wpg));
            store(V_WP, add(prev(V_WP), wpn));
            // NESOSIM.py:327,329 (left to right), then fill_nan_no_negative (332-333)
            double h0n = add(add(add(add(add(add(h0, acc), wpl), lead), atm), adv0), div0);
            double h1n = add(add(add(h1, wpg), adv1), div1);
            h0n = mask_nan(h0n, land, true);
            h1n = mask_nan(h1n, land, true);
            store(V_H0, h0n);
            store(V_H1, h1n);
            store(V_DENS, density_variable(h0n, h1n, land, a.k));
            if (STRIP && own) mail_store(y, gx, h0n, h1n);
        }
    }

    // ---- output row y without an ocean cell, depths already NaN: closed form (same algebra as day_step_land_tile)
    template <bool EDGE>
    __device__ __forceinline__ void out_row_land(int y, const Point (&pt)[CPL]) const {
        const int nx = a.nx;
        const double nan = qnan();
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const Point &p = pt[j];
            if (!owned(j, EDGE)) continue;
            const int gx = c0 - 1 + lane + 32 * j;
            const long long o = (long long)y * nx + gx;
            const long long ip = moffp + o, id = moffd + o;
            auto prev = [&](int v) { return p.prev[v - V_ACC]; };
            auto store = [&](int v, double val) {
                if (v != V_DENS || a.next[v]) a.next[v][(v == V_H0 || v == V_H1) ? id : ip] = val;
            };
            const double pd = div_const(p.P, a.rho_new);           // precipDayT/snowDensityNew (NESOSIM.py:260)
            store(V_ACC, add(prev(V_ACC), mul(pd, p.C)));           // NESOSIM.py:263-264
            store(V_OCEAN, add(prev(V_OCEAN), -mul(pd, sub(1.0, p.C))));   // NESOSIM.py:267-268
            store(V_ADV, nan);
            store(V_DIV, nan);
            store(V_LEAD, a.sw.leadloss ? nan : add(prev(V_LEAD), 0.0));
            store(V_ATM, a.sw.atmloss ? nan : add(prev(V_ATM), 0.0));
            store(V_WPL, a.sw.windpack ? nan : add(prev(V_WPL), 0.0));
            store(V_WPG, a.sw.windpack ? nan : add(prev(V_WPG), 0.0));
            store(V_WP, a.sw.windpack ? nan : add(prev(V_WP), 0.0));
            store(V_H0, nan);
            store(V_H1, nan);
            store(V_DENS, nan);
            if (STRIP) mail_store(y, gx, nan, nan);
        }
    }

    // ---- one chain: output rows [y0, y1) of one strip
    template <bool EDGE>
    __device__ __forceinline__ void chain(const MarchChain &ch, int ny_pad) {
        c0 = ch.strip * MW;
        rd = q.rowdesc + (size_t)ch.strip * ny_pad + MARCH_PAD;
        hbad = rbad = 0u;
        const int y0 = ch.y0, y1 = ch.y1;
        // prologue: depth / drift rows y0-2 .. y0+1 fill the four ring slots, raw row y0-1 is formed from them, then row
        // y0+2 takes the slot of row y0-2 and raw row y0 follows
        auto stage_now = [&](int r) {
            if (eff(desc(r).z) & RF_H) {
                stage_issue<EDGE>(r);
                stage_commit(r);
            } else {
                hbad &= ~(1u << (r & 3));
            }
        };
        for (int r = y0 - 2; r <= y0 + 1; ++r) stage_now(r);
        __syncwarp();
        raw_row<EDGE>(y0 - 1, (eff(desc(y0 - 1).z) & RF_RAW) != 0u);
        __syncwarp();
        stage_now(y0 + 2);
        __syncwarp();
        raw_row<EDGE>(y0, (eff(desc(y0).z) & RF_RAW) != 0u);
        __syncwarp();
        // descriptors of the rows the first step looks at; every later step fetches those of the step after it
        uint4 d0 = desc(y0);
        unsigned f0 = eff(d0.z), f1 = eff(desc(y0 + 1).z), f2 = eff(desc(y0 + 2).z), f3 = eff(desc(y0 + 3).z);
        unsigned lb_lo = d0.x, lb_hi = d0.y;
        for (int y = y0; y < y1; ++y) {
            const uint4 dn = desc(y + 1);                          // land bits of the next row ...
            const unsigned fn = desc(y + 4).z;                     // ... and the flags of the row it will stage
            const int rs = y + 3;                                  // staged for the NEXT step's raw row y+2
            const bool do_stage = rs <= y1 + 1 && (f3 & RF_H) != 0u;
            if (do_stage) stage_issue<EDGE>(rs);
            const bool ocean_row = (f0 & RF_OCEAN) != 0u;
            Point pt[CPL];
            if (ocean_row) {
                point_load_ocean<EDGE>(y, pt);
                raw_row<EDGE>(y + 1, (f1 & RF_RAW) != 0u);
                __syncwarp();
                out_row_ocean<EDGE>(y, lb_lo, lb_hi, pt);
            } else {
                // a row without an ocean cell only keeps the rings going here; its own (closed-form) outputs are written
                // by the land pass below, several rows per memory round trip
                raw_row<EDGE>(y + 1, (f1 & RF_RAW) != 0u);
                __syncwarp();
            }
            if (do_stage) stage_commit(rs);
            else hbad &= ~(1u << (rs & 3));
            __syncwarp();
            f0 = f1; f1 = f2; f2 = f3; f3 = eff(fn);
            lb_lo = dn.x; lb_hi = dn.y;
        }
        // ---- land pass: the chain's rows without an ocean cell.  With every loss term switched on (the usual
        // configuration) such a row needs four values per cell -- snowfall, concentration and the two accumulators that
        // never see the land mask -- and LB rows are fetched per memory round trip; otherwise row by row with the
        // accumulators whose switch is off as well.
        if (a.sw.leadloss && a.sw.atmloss && a.sw.windpack) {
            constexpr int LB = 4;
            for (int y = y0; y < y1;) {
                int ys[LB], n = 0;
#pragma unroll
                for (int b = 0; b < LB; ++b) {
                    while (y < y1 && (eff(desc(y).z) & RF_OCEAN)) ++y;
                    ys[b] = y;
                    if (y < y1) { ++n; ++y; }
                }
                if (n == 0) break;
#pragma unroll
                for (int b = 1; b < LB; ++b)
                    if (b >= n) ys[b] = ys[b - 1];        // a short batch repeats its last row: the same values again
                double P[LB][CPL], C[LB][CPL], pa[LB][CPL], po[LB][CPL];
#pragma unroll
                for (int b = 0; b < LB; ++b)
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        const long long o = cell_of<EDGE>(ys[b], j), ip = moffp + o;
                        P[b][j] = __ldg(a.P + o);
                        C[b][j] = __ldg(a.C + o);
                        pa[b][j] = a.prev[V_ACC][ip];
                        po[b][j] = a.prev[V_OCEAN][ip];
                    }
                const double nan = qnan();
#pragma unroll
                for (int b = 0; b < LB; ++b)
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        if (!owned(j, EDGE)) continue;
                        const int gx = c0 - 1 + lane + 32 * j;
                        const long long o = (long long)ys[b] * a.nx + gx;
                        const long long ip = moffp + o, id = moffd + o;
                        const double pd = div_const(P[b][j], a.rho_new);             // NESOSIM.py:260
                        a.next[V_ACC][ip] = add(pa[b][j], mul(pd, C[b][j]));          // NESOSIM.py:263-264
                        a.next[V_OCEAN][ip] = add(po[b][j], -mul(pd, sub(1.0, C[b][j])));   // NESOSIM.py:267-268
                        a.next[V_ADV][ip] = nan;
                        a.next[V_DIV][ip] = nan;
                        a.next[V_LEAD][ip] = nan;
                        a.next[V_ATM][ip] = nan;
                        a.next[V_WPL][ip] = nan;
                        a.next[V_WPG][ip] = nan;
                        a.next[V_WP][ip] = nan;
                        a.next[V_H0][id] = nan;
                        a.next[V_H1][id] = nan;
                        if (a.next[V_DENS]) a.next[V_DENS][ip] = nan;
                        if (STRIP) mail_store(ys[b], gx, nan, nan);
                    }
            }
        } else {
            for (int y = y0; y < y1; ++y) {
                if (eff(desc(y).z) & RF_OCEAN) continue;
                Point pt[CPL];
                point_load_land<EDGE>(y, pt);
                out_row_land<EDGE>(y, pt);
            }
        }
    }
};

// MINB = CTAs per SM the build is compiled for (register budget 65536 / (128 * MINB) per thread)
template <bool STRIP, int CPL, int MINB>
__global__ void __launch_bounds__(MARCH_WARPS * 32, MINB)
day_march_kernel(const __grid_constant__ DayArgs a, const __grid_constant__ MarchArgs q, const __grid_constant__ StripLink s) {
    extern __shared__ __align__(16) unsigned char march_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int worker = blockIdx.x * MARCH_WARPS + warp;
    const int m = blockIdx.y;
    int cbeg = 0, cend = 0;
    if (worker < q.nworkers) {
        cbeg = q.worker_first[worker];
        cend = q.worker_first[worker + 1];
    }
    if (STRIP) {
        // Warps whose chains read ghost rows wait for the neighbours' delivery of slot x first; only then may the next
        // day's CTAs take SM slots (see day_step_strip): every warp of this launch that depends on another strip has
        // what it needs, so the launch finishes on its own.  All other CTAs release their dependents at once.
        bool reads = false;
        for (int c = cbeg; c < cend; ++c) reads = reads || (q.chains[c].flags & MC_READS_MAIL);
        if (reads && s.use_mail && lane == 0) {
            if (s.has_up) strip_wait(s.flag_top, s.base + (unsigned long long)a.x, s);
            if (s.has_dn) strip_wait(s.flag_bot, s.base + (unsigned long long)a.x, s);
        }
        __syncthreads();
    }
    pdl_launch_dependents();
    MarchWarpSmem<CPL> &S = reinterpret_cast<MarchWarpSmem<CPL> *>(march_smem)[warp];
    MarchCtx<STRIP, CPL> ctx{a, q, s, S, lane, m, (long long)m * a.plane_mstride, (long long)m * a.depth_mstride, a.coef[m], 0u, 0u, nullptr, 0};
    const int ny_pad = a.ny + 2 * MARCH_PAD;
    pdl_wait();      // ---- from here on yesterday's slot may be read
    for (int c = cbeg; c < cend; ++c) {
        const MarchChain ch = q.chains[c];
        if (ch.flags & MC_EDGE) ctx.template chain<true>(ch, ny_pad);
        else ctx.template chain<false>(ch, ny_pad);
        if (STRIP && (ch.flags & (MC_MAIL_TOP | MC_MAIL_BOT))) {
            __threadfence_system();      // every lane's mailbox stores ...
            __syncwarp();
            if (lane == 0) {             // ... before the last chain of the side publishes slot x+1 to the neighbour
                const unsigned long long v = s.base + (unsigned long long)a.x + 1ull;
                if (ch.flags & MC_MAIL_TOP) strip_signal(s.cnt_top, s.expect_top, s.peer_up_flag, v);
                if (ch.flags & MC_MAIL_BOT) strip_signal(s.cnt_bot, s.expect_bot, s.peer_dn_flag, v);
            }
        }
    }
}

}  // namespace nesosim

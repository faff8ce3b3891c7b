// cell_math.cuh -- per-cell fp64 arithmetic of NESOSIM's daily budget, in the reference's evaluation order.
//
// Every expression below mirrors one line of /root/reference/source/NESOSIM.py (cited) as a tree of single
// IEEE-754 double operations.  The explicit __dmul_rn/__dadd_rn/... intrinsics are never contracted into FMAs
// by nvcc, which is what makes finite results bit-identical to numpy (SURVEY.md §7 "bit-exactness discipline").
// NaN algebra is never short-circuited: 0*NaN stays NaN (NESOSIM.py:69-71).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nesosim {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
// exponent-field tests run on the integer pipe, leaving the fp64 pipe to the arithmetic
__device__ __forceinline__ unsigned biased_exp(double x) { return ((unsigned)__double2hiint(x) >> 20) & 0x7ffu; }
__device__ __forceinline__ bool finite(double x) { return biased_exp(x) != 0x7ffu; }

// x / c for a divisor that is fixed for the whole launch (c finite, non-zero, normal).  `rc` = RN(1/c) (host).
// Fast path (Markstein / Brisebarre-Muller-Raina "division by a known constant"): q0 = RN(x*rc);
// r = fma(-c, q0, x); q = RN(q0 + r*rc).  Before the last rounding the value is x/c*(1+eta), |eta| <= 3*2^-107
// relative to the quotient's binade, so q is the correctly rounded quotient unless x/c lies that close to a
// rounding midpoint.  The host only sets `fast` for divisors where that cannot happen (const_div_host() in
// nesosim_abi.cu proves it per divisor).  Zero, infinite and NaN dividends -- more than half of every plane is
// NaN land -- are exactly x*rc (same sign rules as IEEE division by a finite non-zero constant) and never reach
// the division subroutine; only finite dividends outside the magnitude window (or a divisor without the proof)
// (outside 2^-623..2^624) use the IEEE division sequence.  The result always equals numpy's `x / c`.
struct ConstDiv {
    double c, rc;
    int fast;
};
// The IEEE division sequence is kept behind a real call: inlined, the compiler predicates it and every warp
// that holds a NaN or zero lane (every warp near land) would run the division subroutine's special-operand path.
__device__ __noinline__ double div_generic(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double div_const(double x, const ConstDiv &d) {
    const unsigned e = biased_exp(x);
    const double q0 = __dmul_rn(x, d.rc);
    if (d.fast && (e - 400u) <= 1246u)           // 2^-623 <= |x| < 2^624: x, q0 and the residual are normal
        return __fma_rn(__fma_rn(-d.c, q0, x), d.rc, q0);
    if (e == 0x7ffu || (e == 0u && ((__double2hiint(x) & 0x000fffff) | __double2loint(x)) == 0)) return q0;   // inf, NaN, +-0
    return div_generic(x, d.c);
}

// Bare form for the season-resident kernel's loops (requires d.fast): exactly the three operations of the fast
// path, no operand test.  Correct for x = 0 (gives 0), NaN (NaN) and every finite x with 2^-623 <= |x| < 2^624; the
// kernel guarantees that range structurally by guarding its INPUTS (see out_of_guard) instead of every dividend.
__device__ __forceinline__ double div_const_bare(double x, const ConstDiv &d) {
    const double q0 = __dmul_rn(x, d.rc);
    return __fma_rn(__fma_rn(-d.c, q0, x), d.rc, q0);
}
// 1 if x is finite, non-zero and outside [2^-150, 2^150].  With every depth, drift displacement and drift gradient
// inside that guard (or zero / NaN) each dividend of the season-resident kernel is zero, NaN or within
// 2^-506 .. 2^400: depth differences are multiples of 2^-202, gradients divide by at most 2^30, products with a
// guarded factor stay above 2^-382, sums of such products above 2^-434, nine-tap sums with weights in
// [2^-20, 2^20] above 2^-506 -- far inside the window div_const_bare needs.
__device__ __forceinline__ unsigned out_of_guard(double x) {
    const unsigned hi = (unsigned)__double2hiint(x);
    const unsigned e = (hi >> 20) & 0x7ffu;
    const bool zero = ((hi & 0x7fffffffu) | (unsigned)__double2loint(x)) == 0u;
    return (unsigned)((e - 873u) > 300u && e != 0x7ffu && !zero);
}

// np.gradient(f, dx, axis) with edge_order=1 and uniform spacing (numpy; call sites NESOSIM.py:204-211):
// interior (f[+1]-f[-1])/(2.*dx), first (f[1]-f[0])/dx, last (f[n-1]-f[n-2])/dx.
struct GradConsts {
    ConstDiv dx, two_dx;
};
// `fm`, `fc`, `fp` are the values at index-1, index, index+1 (fm/fp ignored where they fall off the grid).
// One subtraction and one division per call: the operands and the divisor are selected first.
__device__ __forceinline__ double gradient1d(double fm, double fc, double fp, int idx, int n, const GradConsts &g) {
    const bool first = (idx == 0), last = (idx == n - 1);
    const double hi = last ? fc : fp;
    const double lo = first ? fc : fm;
    ConstDiv d;
    d.c = (first || last) ? g.dx.c : g.two_dx.c;
    d.rc = (first || last) ? g.dx.rc : g.two_dx.rc;
    d.fast = g.dx.fast & g.two_dx.fast;
    return div_const(sub(hi, lo), d);
}

// a / b for two variable operands with IEEE results.  Lanes whose operands are not both normal numbers (NaN
// land, 0/0 on snow-free cells, ...) never reach the division sequence: the common special cases are written out
// and the sequence itself runs on substituted normal operands, so its special-operand subroutine is never called.
__device__ __forceinline__ double div_ieee(double a, double b) {
    const unsigned ea = biased_exp(a), eb = biased_exp(b);
    const bool normal = (ea - 1u) <= 0x7fdu && (eb - 1u) <= 0x7fdu;
    const double q = __ddiv_rn(normal ? a : 1.0, normal ? b : 1.0);
    if (normal) return q;
    if (a != a || b != b) return qnan();
    const bool neg = (__double2hiint(a) ^ __double2hiint(b)) < 0;
    const bool a0 = (a == 0.0), b0 = (b == 0.0), ai = (ea == 0x7ffu), bi = (eb == 0x7ffu);
    if ((a0 && b0) || (ai && bi)) return qnan();
    if (b0 || ai) return __longlong_as_double(neg ? 0xfff0000000000000LL : 0x7ff0000000000000LL);
    if (a0 || bi) return __longlong_as_double(neg ? 0x8000000000000000LL : 0LL);
    return div_generic(a, b);   // a denormal operand
}

// fillMaskAndNaNWithZero (NESOSIM.py:127-139): NaN -> 0, +-inf -> 0
__device__ __forceinline__ double zero_if_nonfinite(double x) { return finite(x) ? x : 0.0; }

// fill_nan_no_negative (NESOSIM.py:141-166): non-finite -> NaN; land/coast/lakes -> NaN; optionally <0 -> 0.
// `land` = (mask > 10) || (mask < 1).
__device__ __forceinline__ double mask_nan(double x, bool land, bool negative_to_zero) {
    if (!finite(x) || land) return qnan();
    if (negative_to_zero && x < 0.0) return 0.0;
    return x;
}
__device__ __forceinline__ bool is_land(uint8_t m) { return m > 10 || m < 1; }

// calcDynamics (NESOSIM.py:204-213) for one layer at one cell, before the NaN->0 fill.
//   div = -((h*gx(U*dT)) + (h*gy(V*dT)));  adv = -(((U*dT)*gx(h)) + ((V*dT)*gy(h)))
__device__ __forceinline__ double div_term(double h, double gxu, double gyv) { return -add(mul(h, gxu), mul(h, gyv)); }
__device__ __forceinline__ double adv_term(double ut, double vt, double gxh, double gyh) {
    return -add(mul(ut, gxh), mul(vt, gyh));
}

// Per-member coefficients with the scalar sub-products the reference forms in Python floats first.
struct MemberCoef {
    double llf;       // leadLossFactor
    double alf;       // atmLossFactor
    double wpt;       // windPackThresh
    double neg_wpf_dt;  // (-windPackFactor)*deltaT   (NESOSIM.py:119, scalar*scalar before the array multiply)
    double wpf_dt;      // windPackFactor*deltaT      (NESOSIM.py:122)
};

struct ModelConsts {
    double deltaT, rhoFresh, rhoOld, rho_ratio /* rhoFresh/rhoOld, NESOSIM.py:122 */, minSnowD, minConc;
};

// windT = np.where(W > thresh, 1, 0) as a double (NaN > thr is False)
__device__ __forceinline__ double wind_flag(double W, double thr) { return (W > thr) ? 1.0 : 0.0; }

// calcLeadLoss (NESOSIM.py:71): -(windT*LLF*dT*h0*W*(1-C))
__device__ __forceinline__ double lead_loss(double wt, double h0, double W, double C, const MemberCoef &m,
                                            const ModelConsts &k) {
    return -mul(mul(mul(mul(mul(wt, m.llf), k.deltaT), h0), W), sub(1.0, C));
}
// calcAtmLoss (NESOSIM.py:94): -(windT*dT*h0*W*ALF)
__device__ __forceinline__ double atm_loss(double wt, double h0, double W, const MemberCoef &m, const ModelConsts &k) {
    return -mul(mul(mul(mul(wt, k.deltaT), h0), W), m.alf);
}
// calcWindPacking (NESOSIM.py:119-124)
__device__ __forceinline__ void wind_packing(double wt, double h0, const MemberCoef &m, const ModelConsts &k,
                                             double &loss, double &gain, double &net) {
    loss = mul(mul(m.neg_wpf_dt, wt), h0);
    gain = mul(mul(mul(m.wpf_dt, wt), h0), k.rho_ratio);
    net = add(loss, gain);
}

// densityCalc (NESOSIM.py:464-471) on the updated depths
__device__ __forceinline__ double density_variable(double h0, double h1, bool land, const ModelConsts &k) {
    const double den = add(h0, h1);
    double rho = div_ieee(add(mul(h0, k.rhoFresh), mul(h1, k.rhoOld)), den);
    if (rho > k.rhoOld) rho = k.rhoOld;
    if (rho < k.rhoFresh) rho = k.rhoFresh;
    if (land) rho = qnan();
    if (den < k.minSnowD) rho = qnan();
    return rho;
}
// densityCalc for a cell that is known to be ocean, branch-free: the quotient only matters where the total depth
// is a number >= minSnowD (everything else ends as NaN, NESOSIM.py:471); there both operands are normal numbers
// unless minSnowD <= 0 or the depths overflow -- those lanes raise `bad` and are redone with density_variable().
__device__ __forceinline__ double density_ocean_flagged(double h0, double h1, const ModelConsts &k, unsigned &bad) {
    const double den = add(h0, h1);
    const double num = add(mul(h0, k.rhoFresh), mul(h1, k.rhoOld));
    const bool normal = (biased_exp(num) - 1u) <= 0x7fdu && (biased_exp(den) - 1u) <= 0x7fdu;
    const bool need = !(den < k.minSnowD) && den == den;
    double rho = __ddiv_rn(normal ? num : 1.0, normal ? den : 1.0);
    bad |= (unsigned)(need && !normal);
    if (rho > k.rhoOld) rho = k.rhoOld;
    if (rho < k.rhoFresh) rho = k.rhoFresh;
    return need ? rho : qnan();
}
// clim branch (NESOSIM.py:339-344)
__device__ __forceinline__ double density_clim(double rho_new, double h0, double h1, double C, bool land,
                                               const ModelConsts &k) {
    double rho = rho_new;
    if (land) rho = qnan();
    if (C < k.minConc) rho = qnan();
    if (add(h0, h1) < k.minSnowD) rho = qnan();
    return rho;
}

// 3x3 convolution tap order of astropy's C loop (rows outer, columns inner, flipped kernel, accumulator
// starting at 0.0): see oracle/astropy_restated.py.  w[] is the 3x3 kernel row-major; v(r,c) fetches the
// zero-padded input at offset (r-1, c-1) from the output cell.
template <typename Fetch>
__device__ __forceinline__ double conv3x3(const double *w, Fetch v) {
    double top = 0.0;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) top = add(top, mul(v(ii, jj), w[(2 - ii) * 3 + (2 - jj)]));
    return top;
}

}  // namespace nesosim

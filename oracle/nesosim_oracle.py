"""CPU oracle: numpy restatement of NESOSIM v1.1's daily two-layer snow-budget step.

TEST INFRASTRUCTURE ONLY -- never imported by the product (``nesosim_b200``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` use it, as the checker / the baseline,
never as the thing shipped.

Every function restates one function of ``/root/reference/source/NESOSIM.py`` (cited per function) with the
reference's *evaluation order* preserved, so that on the same inputs every finite value is identical to what
the reference's numpy code produces.  Differences from the reference are structural only:
parameters travel in an explicit :class:`Params` instead of module globals, and state lives in a dict.

Pinning: the reference has no tests or golden vectors.  This oracle is pinned instead against
the *reference's own functions executed verbatim* (``oracle/ref_loader.py`` stub-imports
``/root/reference/source/NESOSIM.py`` in the build container; ``tests/golden/make_golden.py`` writes the
fixtures, ``tests/test_oracle_golden.py`` checks them).  The one boundary that stays "parity unpinned" is
astropy's ``convolve``/``Gaussian2DKernel`` (not installed anywhere here) -- see ``oracle/astropy_restated.py``.
"""
from dataclasses import dataclass, field

import numpy as np

from .astropy_restated import convolve_fill0, gaussian2d_kernel

ACCUMULATORS = ("snowAcc", "snowOcean", "snowAdv", "snowDiv", "snowLead", "snowAtm",
                "snowWindPackLoss", "snowWindPackGain", "snowWindPack")
FORCING_COPIES = ("precipDays", "iceConcDays", "windDays", "tempDays")
ALL_ARRAYS = FORCING_COPIES + ("snowDepths", "density") + ACCUMULATORS


@dataclass
class Params:
    """The module globals ``main`` sets at ``NESOSIM.py:527-541``."""
    windPackFactor: float = 5.8e-7
    windPackThresh: float = 5.
    leadLossFactor: float = 2.9e-7
    atmLossFactor: float = 2.2e-8
    snowDensityFresh: float = 200.
    snowDensityOld: float = 350.
    minSnowD: float = 0.02
    minConc: float = 0.15
    deltaT: float = 60. * 60. * 24.


@dataclass
class Flags:
    """Keyword switches of ``calcBudget`` (``NESOSIM.py:227``)."""
    dynamicsInc: int = 1
    leadlossInc: int = 1
    windpackInc: int = 1
    atmlossInc: int = 0
    densityType: str = "variable"
    conv_variant: str = "post_divide"


# --------------------------------------------------------------------------------------- point-wise terms

def lead_loss(h0, wind, conc, p):
    """``calcLeadLoss`` (NESOSIM.py:51-72): -(windT*LLF*dT*h0*W*(1-C)), evaluated left to right."""
    windT = np.where(wind > p.windPackThresh, 1, 0)
    return -(windT * p.leadLossFactor * p.deltaT * h0 * wind * (1 - conc))


def atm_loss(h0, wind, p):
    """``calcAtmLoss`` (NESOSIM.py:74-95): -(windT*dT*h0*W*ALF)."""
    windT = np.where(wind > p.windPackThresh, 1, 0)
    return -(windT * p.deltaT * h0 * wind * p.atmLossFactor)


def wind_packing(wind, h0, p):
    """``calcWindPacking`` (NESOSIM.py:97-125): loss from the new layer, gain to the old layer, net."""
    windT = np.where(wind > p.windPackThresh, 1, 0)
    loss = -p.windPackFactor * p.deltaT * windT * h0
    gain = p.windPackFactor * p.deltaT * windT * h0 * (p.snowDensityFresh / p.snowDensityOld)
    return loss, gain, loss + gain


def fill_mask_nan_zero(arr):
    """``fillMaskAndNaNWithZero`` (NESOSIM.py:127-139): NaN -> 0 then +-inf -> 0, in place."""
    arr[np.isnan(arr)] = 0.
    arr[~np.isfinite(arr)] = 0.


def fill_nan_no_negative(arr, mask, negative_to_zero=True):
    """``fill_nan_no_negative`` (NESOSIM.py:141-166), in place, in the reference's order."""
    arr[~np.isfinite(arr)] = np.nan
    arr[np.where(mask > 10)] = np.nan
    arr[np.where(mask < 1)] = np.nan
    if negative_to_zero:
        with np.errstate(invalid="ignore"):
            arr[np.where(arr < 0.)] = 0.


def smooth_snow(arr, variant="post_divide"):
    """``smooth_snow`` (NESOSIM.py:170-187): 3x3 sigma=1 Gaussian, astropy ``convolve`` defaults."""
    return convolve_fill0(arr, gaussian2d_kernel(x_stddev=1, x_size=3, y_size=3), variant=variant)


def gradient_explicit(f, dx, axis):
    """What ``np.gradient(f, dx, axis=axis)`` computes (edge_order=1, uniform spacing) -- SURVEY.md §8 row a3.

    interior (f[i+1]-f[i-1])/(2.*dx); first (f[1]-f[0])/dx; last (f[-1]-f[-2])/dx.
    """
    f = np.asarray(f, dtype=float)
    out = np.empty_like(f)
    fm = np.moveaxis(f, axis, 0)
    om = np.moveaxis(out, axis, 0)
    with np.errstate(all="ignore"):
        om[1:-1] = (fm[2:] - fm[:-2]) / (2. * dx)
        om[0] = (fm[1] - fm[0]) / dx
        om[-1] = (fm[-1] - fm[-2]) / dx
    return out


def calc_dynamics(drift, h, dx, p):
    """``calcDynamics`` (NESOSIM.py:189-222).  ``drift`` (2,ny,nx), ``h`` (2,ny,nx)."""
    with np.errstate(all="ignore"):
        divx = h * np.gradient(drift[0] * p.deltaT, dx, axis=(1))
        divy = h * np.gradient(drift[1] * p.deltaT, dx, axis=(0))
        div = -(divx + divy)
        advx = drift[0] * p.deltaT * np.gradient(h, dx, axis=(2))
        advy = drift[1] * p.deltaT * np.gradient(h, dx, axis=(1))
        adv = -(advx + advy)
    for plane in (adv[0], adv[1], div[0], div[1]):
        fill_mask_nan_zero(plane)
    return adv, div


def density_calc(h, conc, mask, p):
    """``densityCalc`` (NESOSIM.py:458-473); ``conc`` is accepted and unused, as in the reference."""
    with np.errstate(all="ignore"):
        rho = ((h[0] * p.snowDensityFresh) + (h[1] * p.snowDensityOld)) / (h[0] + h[1])
        rho[np.where(rho > p.snowDensityOld)] = p.snowDensityOld
        rho[np.where(rho < p.snowDensityFresh)] = p.snowDensityFresh
        rho[np.where(mask < 1)] = np.nan
        rho[np.where(mask > 10)] = np.nan
        rho[np.where((h[0] + h[1]) < p.minSnowD)] = np.nan
    return rho


# ------------------------------------------------------------------------------------------- state + step

def gen_empty_arrays(num_days, ny, nx):
    """``genEmptyArrays`` (NESOSIM.py:350-376) as a dict (the reference returns a 15-tuple)."""
    s = {k: np.zeros((num_days, ny, nx)) for k in ALL_ARRAYS if k != "snowDepths"}
    s["snowDepths"] = np.zeros((num_days, 2, ny, nx))
    return s


def calc_budget(s, conc, precip, drift, wind, temp, mask, dx, x, p, f, rho_clim_day=None):
    """``calcBudget`` (NESOSIM.py:224-347): advance day x -> x+1 in place in the state dict ``s``."""
    s["precipDays"][x] = precip
    s["iceConcDays"][x] = conc
    s["windDays"][x] = wind
    s["tempDays"][x] = temp
    h = s["snowDepths"]
    zeros = lambda: np.zeros(conc.shape)

    rho_new = rho_clim_day if f.densityType == "clim" else p.snowDensityFresh
    with np.errstate(all="ignore"):
        pd = precip / rho_new
        acc = pd * conc
        s["snowAcc"][x + 1] = s["snowAcc"][x] + acc
        oc = -(pd * (1 - conc))
        s["snowOcean"][x + 1] = s["snowOcean"][x] + oc

        if f.dynamicsInc == 1:
            adv, div = calc_dynamics(drift, h[x], dx, p)
            adv[0] = smooth_snow(adv[0], f.conv_variant)
            adv[1] = smooth_snow(adv[1], f.conv_variant)
            div[0] = smooth_snow(div[0], f.conv_variant)
            div[1] = smooth_snow(div[1], f.conv_variant)
            for plane in (adv[0], adv[1], div[0], div[1]):
                fill_nan_no_negative(plane, mask, negative_to_zero=False)
        else:
            # the reference makes (ny,nx) zeros and indexes rows 0/1, which broadcast as zeros (Appendix A)
            adv = np.zeros((2,) + conc.shape)
            div = np.zeros((2,) + conc.shape)
        s["snowAdv"][x + 1] = s["snowAdv"][x] + adv[0] + adv[1]
        s["snowDiv"][x + 1] = s["snowDiv"][x] + div[0] + div[1]

        lead = lead_loss(h[x, 0], wind, conc, p) if f.leadlossInc == 1 else zeros()
        s["snowLead"][x + 1] = s["snowLead"][x] + lead
        atm = atm_loss(h[x, 0], wind, p) if f.atmlossInc == 1 else zeros()
        s["snowAtm"][x + 1] = s["snowAtm"][x] + atm
        if f.windpackInc == 1:
            wpl, wpg, wpn = wind_packing(wind, h[x, 0], p)
        else:
            wpl, wpg, wpn = zeros(), zeros(), zeros()
        s["snowWindPackLoss"][x + 1] = s["snowWindPackLoss"][x] + wpl
        s["snowWindPackGain"][x + 1] = s["snowWindPackGain"][x] + wpg
        s["snowWindPack"][x + 1] = s["snowWindPack"][x] + wpn

        h[x + 1, 0] = h[x, 0] + acc + wpl + lead + atm + adv[0] + div[0]
        h[x + 1, 1] = h[x, 1] + wpg + adv[1] + div[1]
        fill_nan_no_negative(h[x + 1, 0], mask)
        fill_nan_no_negative(h[x + 1, 1], mask)

        if f.densityType == "clim":
            d = s["density"]
            d[x + 1] = rho_new
            d[x + 1][np.where(mask > 10)] = np.nan
            d[x + 1][np.where(mask < 1)] = np.nan
            d[x + 1][np.where(conc < p.minConc)] = np.nan
            d[x + 1][np.where((h[x + 1][0] + h[x + 1][1]) < p.minSnowD)] = np.nan
        else:
            s["density"][x + 1] = density_calc(h[x + 1], conc, mask, p)


def initial_depths(ic, conc0, p):
    """IC handling of ``main`` (NESOSIM.py:604-609): zero where conc<minConc, split 50/50."""
    ic = np.array(ic, dtype=float)
    with np.errstate(invalid="ignore"):
        ic[np.where(conc0 < p.minConc)] = 0
    return ic * 0.5


def run_season(forcing, ic, mask, dx, p=None, f=None, rho_clim=None, num_steps=None):
    """Day loop of ``main`` (NESOSIM.py:586-649) on pre-loaded forcing.

    ``forcing``: dict with ``precip, conc, wind`` (T,ny,nx), ``drift`` (T,2,ny,nx), optional ``temp``; slot
    T-1 only feeds the forcing copies (NESOSIM.py:645-649).  ``ic``: (ny,nx) total initial depth or None (IC=0).
    ``rho_clim``: (T,) fresh-snow density per step for densityType='clim' (``utils.densityClim`` value).
    """
    p = p or Params()
    f = f or Flags()
    T, ny, nx = forcing["precip"].shape
    s = gen_empty_arrays(T, ny, nx)
    temp = forcing.get("temp")
    if temp is None:
        temp = np.full((T, ny, nx), np.nan)
    if ic is not None:
        half = initial_depths(ic, forcing["conc"][0], p)
        s["snowDepths"][0, 0] = half
        s["snowDepths"][0, 1] = half
    steps = T - 1 if num_steps is None else num_steps
    for x in range(steps):
        calc_budget(s, forcing["conc"][x], forcing["precip"][x], forcing["drift"][x], forcing["wind"][x],
                    temp[x], mask, dx, x, p, f,
                    rho_clim_day=None if rho_clim is None else rho_clim[x])
    if steps == T - 1:
        s["precipDays"][T - 1] = forcing["precip"][T - 1]
        s["iceConcDays"][T - 1] = forcing["conc"][T - 1]
        s["windDays"][T - 1] = forcing["wind"][T - 1]
        s["tempDays"][T - 1] = temp[T - 1]
    return s

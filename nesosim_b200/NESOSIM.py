"""Drop-in host module for the reference's ``source/NESOSIM.py`` with the daily budget loop on the GPU.

Same public names, signatures and side effects as the reference for the hot path and its two neighbours:

* ``main(...)``            -- ``NESOSIM.main`` (reference ``source/NESOSIM.py:495-659``): same keyword arguments, same
  output directories and file names, same NetCDF products.  Grid, region mask, calendar, NetCDF writers and plots
  stay the reference's own ``utils`` (imported as ``cF`` exactly like the reference does, so this file can sit next
  to ``utils.py`` in ``source/`` or anywhere with ``source/`` on ``sys.path``).  What changes: the
  ``for x in range(numDays-1): loadData; calcBudget`` loop (``NESOSIM.py:614-639``) becomes "stage the season's
  forcing once -> one native season call" (``nesosim_run_season`` through ``nesosim_b200.engine``).
* ``calcBudget(...)``      -- the reference's in-place operator (``NESOSIM.py:224-347``) on caller-owned numpy arrays:
  slot ``x`` of the forcing copies and slot ``x+1`` of the 11 budget arrays are written, nothing else.
* ``loadData(...)``        -- the daily forcing reader (``NESOSIM.py:379-456``) with its fallbacks.
* ``genEmptyArrays``, ``doyToMonth``, ``applyScaling`` -- same contracts.
* ``calcLeadLoss``, ``calcAtmLoss``, ``calcWindPacking``, ``fillMaskAndNaNWithZero``, ``fill_nan_no_negative``,
  ``smooth_snow``, ``calcDynamics``, ``densityCalc`` -- the functions ``calcBudget`` is made of (``NESOSIM.py:51-222,
  458-473``), one GPU entry point each, for callers and tests that use them one by one.

Model constants are module globals with the reference's names (``NESOSIM.py:517,527-541``) because callers of the
reference set them that way; every native call receives them explicitly.  There is no numpy fallback for the
budget arithmetic: without ``libnesosim_b200.so`` and a CUDA device ``main``/``calcBudget`` raise.
"""
import datetime
import os

import numpy as np
import numpy.ma as ma

# ---- module globals of the reference (NESOSIM.py:517,527-541); main() assigns them, calcBudget() reads them
forcingPath = './'
outPath = './'
ancDataPath = '../anc_data/'
figpath = './'
snowDensityFresh = 200.
snowDensityOld = 350.
minSnowD = 0.02
minConc = 0.15
deltaT = 60. * 60. * 24.
leadLossFactor = 0.1
atmLossFactor = 2.2e-8
windPackThresh = 5.
windPackFactor = 0.1

VERBOSE = True          # the reference prints one line per file it opens; set False to silence the stager


def _say(*a):
    if VERBOSE:
        print(*a)


def _cF():
    """The reference's own helper module (grid, mask, calendar, NetCDF writers, plots)."""
    import utils as cF      # noqa: the reference imports it the same way (NESOSIM.py:44)
    return cF


# ------------------------------------------------------------------------------------------ state arrays

def genEmptyArrays(numDaysT, nxT, nyT):
    """The 15 zero-initialised arrays of the reference (NESOSIM.py:350-376), same order."""
    def z():
        return np.zeros((numDaysT, nxT, nyT))
    precipDays, iceConcDays, windDays, tempDays = z(), z(), z(), z()
    snowDepths = np.zeros((numDaysT, 2, nxT, nyT))
    density = z()
    snowDiv, snowAdv, snowAcc, snowOcean = z(), z(), z(), z()
    snowWindPack, snowWindPackLoss, snowWindPackGain = z(), z(), z()
    snowLead, snowAtm = z(), z()
    return (precipDays, iceConcDays, windDays, tempDays, snowDepths, density, snowDiv, snowAdv, snowAcc, snowOcean,
            snowWindPack, snowWindPackLoss, snowWindPackGain, snowLead, snowAtm)


# ------------------------------------------------------------------------------------------ forcing reader

def _daily_file(kind, var, yearT, dayStr, dxStr, extraStr, precipVar=None, sf_fallback=False):
    """On-disk names written by the reference's gridding scripts (NESOSIM.py:389-448)."""
    y = str(yearT)
    if kind == 'precip':
        mid = 'Precip/' + var + ('/sf/' if sf_fallback else '/') + y + '/'
        return forcingPath + mid + var + 'sf' + dxStr + '-' + y + '_d' + dayStr + extraStr
    if kind == 'wind':
        return forcingPath + 'Winds/' + var + '/' + y + '/' + var + 'winds' + dxStr + '-' + y + '_d' + dayStr + extraStr
    if kind == 'conc':
        return forcingPath + 'IceConc/' + var + '/' + y + '/iceConcG_' + var + dxStr + '-' + y + '_d' + dayStr + extraStr
    if kind == 'drift':
        return forcingPath + 'IceDrift/' + var + '/' + y + '/' + var + '_driftG' + dxStr + '-' + y + '_d' + dayStr + extraStr
    if kind == 'temp':
        return forcingPath + 'Temp/' + precipVar + '/t2m/' + y + '/t2m' + dxStr + '-' + y + '_d' + dayStr + extraStr
    raise ValueError(kind)


def _required(kind, var, yearT, dayStr, dxStr, extraStr, label):
    """A forcing the reference cannot run without: day 365 falls back to day 364 (no leap-day file), anything else
    missing ends the run like the reference's ``print(...); exit()`` (NESOSIM.py:392-427)."""
    path = _daily_file(kind, var, yearT, dayStr, dxStr, extraStr)
    _say('Loading gridded %s forcing from:' % label, path)
    try:
        return np.load(path, allow_pickle=True)
    except Exception:
        if dayStr == '365':
            _say('no leap year data, using data from the previous day')
            # (the reference's precipitation fallback looks under an extra 'sf/' directory, NESOSIM.py:395)
            return np.load(_daily_file(kind, var, yearT, '364', dxStr, extraStr, sf_fallback=(kind == 'precip')),
                           allow_pickle=True)
        print('No %s data so exiting!' % label)
        raise SystemExit()


def loadData(yearT, dayT, precipVar, windVar, concVar, driftVar, dxStr, extraStr):
    """Daily forcing planes exactly as the reference's ``loadData`` returns them (NESOSIM.py:379-456):
    ``(iceConcDayG, precipDayG, driftGdayG, windDayG, tempDayG)``.  Concentration: non-finite -> 0.  Drift: a missing
    file is an all-NaN day, a masked array is filled with NaN.  Temperature: optional, all-NaN when absent."""
    dayStr = '%03d' % dayT
    precipDayG = _required('precip', precipVar, yearT, dayStr, dxStr, extraStr, 'snowfall')
    windDayG = _required('wind', windVar, yearT, dayStr, dxStr, extraStr, 'wind')
    iceConcDayG = _required('conc', concVar, yearT, dayStr, dxStr, extraStr, 'ice conc')
    iceConcDayG[~np.isfinite(iceConcDayG)] = 0.

    path = _daily_file('drift', driftVar, yearT, dayStr, dxStr, extraStr)
    _say('Loading gridded ice drift forcing from:', path)
    try:
        driftGdayG = np.load(path, allow_pickle=True)
    except Exception:
        _say('No drift data')
        driftGdayG = np.full((2,) + iceConcDayG.shape, np.nan)
    driftGdayG = ma.filled(driftGdayG, np.nan)

    try:
        tempDayG = np.load(_daily_file('temp', None, yearT, dayStr, dxStr, extraStr, precipVar=precipVar), allow_pickle=True)
    except Exception:
        tempDayG = np.full(iceConcDayG.shape, np.nan)
    return iceConcDayG, precipDayG, driftGdayG, windDayG, tempDayG


def doyToMonth(day, year):
    """Month of a day-of-year (NESOSIM.py:475-480).  As in the reference the argument is treated as 1-based although
    the model's day counter is 0-based, so the first day of a month maps to the previous month."""
    d = np.datetime64('{}-01-01'.format(year)) + np.timedelta64(day - 1, 'D')
    return d.astype(object).month


def applyScaling(product, factor, scaling_type='mul'):
    """Multiplicative scaling of a daily product (NESOSIM.py:482-492)."""
    if scaling_type == 'mul':
        product_scaled = product * factor
    return product_scaled


def _open_scale_factors(path):
    """CloudSat monthly scaling factors as a (12, ny, nx) array indexed by month-1 (NESOSIM.py:559-563)."""
    import xarray as xr
    f = xr.open_dataset(path)['scale_factors']
    return np.stack([np.asarray(f.loc[m, :, :].values, dtype=np.float64) for m in range(1, 13)])


def _clim_density_table():
    """Daily Warren-climatology densities [kg m-3] (utils.py:1336-1343): 1000 * W99_density.csv 'Density' column."""
    import pandas as pd
    t = pd.read_csv(ancDataPath + '/W99_density.csv', header=0, names=['Day', 'Density'])
    return 1000 * np.asarray(t['Density'], dtype=np.float64)


def stage_season(year1, year2, startDay, numDays, numDaysYear1, precipVar, windVar, concVar, driftVar, dxStr, extraStr,
                 scale_factors=None, clim_table=None):
    """Everything the day loop of ``main`` reads (NESOSIM.py:614-649), stacked over the season:
    precip/conc/wind/temp ``(T, ny, nx)`` -- slots 0..T-2 are the days the steps use, slot T-1 is the day after the
    last step, which the reference loads only to fill the last slot of its forcing copies (NESOSIM.py:645-649) --
    drift ``(T, 2, ny, nx)`` (slot T-1 unused, NaN) and, for densityType='clim', the fresh-snow density per step."""
    staged = None
    rho = np.zeros(numDays)
    yearCurrent = year1
    day = startDay
    for x in range(numDays):
        if x < numDays - 1:
            day = x + startDay
            if day >= numDaysYear1:          # jump into the second year (NESOSIM.py:617-620)
                day = day - numDaysYear1
                yearCurrent = year2
            load_day = day
        else:
            load_day = day + 1               # "load last data": no wrap, as in the reference (NESOSIM.py:645)
        conc, precip, drift, wind, temp = loadData(yearCurrent, load_day, precipVar, windVar, concVar, driftVar, dxStr, extraStr)
        if staged is None:
            ny, nx = conc.shape
            staged = {"precip": np.zeros((numDays, ny, nx)), "conc": np.zeros((numDays, ny, nx)),
                      "wind": np.zeros((numDays, ny, nx)), "temp": np.full((numDays, ny, nx), np.nan),
                      "drift": np.full((numDays, 2, ny, nx), np.nan)}
        if scale_factors is not None and x < numDays - 1:
            precip = applyScaling(precip, scale_factors[doyToMonth(load_day, yearCurrent) - 1], scaling_type='mul')
        staged["precip"][x] = precip
        staged["conc"][x] = conc
        staged["wind"][x] = wind
        staged["temp"][x] = temp
        if x < numDays - 1:
            staged["drift"][x] = drift
            if clim_table is not None:
                rho[x] = clim_table[load_day - 1]      # densityClim(dayT): .iloc[dayT-1] (day 0 wraps to the last row)
    staged["rho_clim"] = rho if clim_table is not None else None
    return staged


# ------------------------------------------------------------------------------------------ the operator

_ENGINES = {}


def _engine(region_maskG, num_days, dx, densityType, dynamicsInc, leadlossInc, windpackInc, atmlossInc):
    from .engine import SnowBudgetEngine, region_codes_u8
    codes = region_codes_u8(region_maskG)
    key = (codes.tobytes(), codes.shape, int(num_days), float(dx), densityType, int(dynamicsInc), int(leadlossInc),
           int(windpackInc), int(atmlossInc), snowDensityFresh, snowDensityOld, minSnowD, minConc, deltaT)
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) > 4:
            for e in _ENGINES.values():
                e.close()
            _ENGINES.clear()
        eng = SnowBudgetEngine(codes, num_days, dx, n_members=1, dynamicsInc=dynamicsInc, leadlossInc=leadlossInc,
                               windpackInc=windpackInc, atmlossInc=atmlossInc, densityType=densityType,
                               snowDensityFresh=snowDensityFresh, snowDensityOld=snowDensityOld, minSnowD=minSnowD,
                               minConc=minConc, deltaT=deltaT)
        _ENGINES[key] = eng
    return eng


def _params_row():
    return [[windPackFactor, windPackThresh, leadLossFactor, atmLossFactor]]


_STATE_NAMES = ("snowDepths", "density", "snowAcc", "snowOcean", "snowAdv", "snowDiv", "snowLead", "snowAtm",
                "snowWindPackLoss", "snowWindPackGain", "snowWindPack")


def calcBudget(xptsG, yptsG, snowDepths, iceConcDayT, precipDayT, driftGdayT, windDayT, tempDayT,
               density, precipDays, iceConcDays, windDays, tempDays, snowAcc, snowOcean, snowAdv,
               snowDiv, snowLead, snowAtm, snowWindPackLoss, snowWindPackGain, snowWindPack, region_maskG, dx, x, dayT,
               densityType='variable', dynamicsInc=1, leadlossInc=1, windpackInc=1, atmlossInc=0):
    """One day of the budget with the reference's in-place contract (NESOSIM.py:224-347): copies the day's forcing
    into slot ``x`` of the four forcing arrays and advances slot ``x`` -> ``x+1`` of the eleven budget arrays.
    The arithmetic runs in ``nesosim_step_day`` on the GPU; only the two touched slots cross the bus."""
    import torch
    precipDays[x] = precipDayT
    iceConcDays[x] = iceConcDayT
    windDays[x] = windDayT
    tempDays[x] = tempDayT
    T = snowDepths.shape[0]
    eng = _engine(region_maskG, T, dx, densityType, dynamicsInc, leadlossInc, windpackInc, atmlossInc)
    host = dict(snowDepths=snowDepths, density=density, snowAcc=snowAcc, snowOcean=snowOcean, snowAdv=snowAdv,
                snowDiv=snowDiv, snowLead=snowLead, snowAtm=snowAtm, snowWindPackLoss=snowWindPackLoss,
                snowWindPackGain=snowWindPackGain, snowWindPack=snowWindPack)
    dev = getattr(eng, "_shim_state", None)
    if dev is None:
        dev = eng.alloc_outputs(zero=True)
        eng._shim_state = dev
    for n in _STATE_NAMES:                          # slot x of the caller's arrays is the state the step reads
        dev[n][0, x].copy_(torch.from_numpy(np.ascontiguousarray(host[n][x])))
    rho_new = float(_clim_density_table()[dayT - 1]) if densityType == 'clim' else snowDensityFresh
    eng.step_day(x, iceConcDayT, precipDayT, driftGdayT, windDayT, _params_row(), dev, rho_new=rho_new)
    torch.cuda.synchronize()
    for n in _STATE_NAMES:
        host[n][x + 1] = dev[n][0, x + 1].cpu().numpy()


# ------------------------------------------------------------------------------------------ the per-function operators
# The functions calcBudget is made of, with the reference's names, arguments and return values (NESOSIM.py:51-222,
# 458-473), each on the GPU through the matching per-function entry point of the C ABI (nesosim_op_*, nesosim_smooth):
# what a caller -- or a test written against the reference -- that uses them one by one gets.  numpy in, numpy out; the
# model constants are the module globals above, as in the reference.  (The season path does not go through these: it
# runs the fused kernels.)

def _ops():
    """The native per-function operators (nesosim_b200.engine); tests on a machine without a GPU swap in a stand-in."""
    from . import engine
    return engine


def _plain(arr, what):
    if isinstance(arr, ma.MaskedArray):
        raise TypeError(what + ": plain ndarrays only (what calcBudget passes); fill the masked array first")
    return arr


def calcLeadLoss(snowDepthT, windDayT, iceConcDaysT):
    """Snow lost to leads from the new-snow layer (NESOSIM.py:51-72): ``-(windT*LLF*deltaT*h0*W*(1-C))``."""
    return _ops().op_wind_terms(snowDepthT, windDayT, iceConcDaysT, _params_row(), deltaT, snowDensityFresh,
                                snowDensityOld)[0]


def calcAtmLoss(snowDepthT, windDayT):
    """Snow lost to the atmosphere (NESOSIM.py:74-95): ``-(windT*deltaT*h0*W*ALF)``."""
    return _ops().op_wind_terms(snowDepthT, windDayT, np.zeros_like(np.asarray(windDayT, dtype=np.float64)), _params_row(),
                                deltaT, snowDensityFresh, snowDensityOld)[1]


def calcWindPacking(windDayT, snowDepthT0):
    """Wind packing (NESOSIM.py:97-125): loss from the new-snow layer, gain of the old-snow layer, net."""
    out = _ops().op_wind_terms(snowDepthT0, windDayT, np.zeros_like(np.asarray(windDayT, dtype=np.float64)), _params_row(),
                               deltaT, snowDensityFresh, snowDensityOld)
    return out[2], out[3], out[4]


def fillMaskAndNaNWithZero(arr):
    """NaN and +-inf -> 0, in place (NESOSIM.py:127-139); returns None like the reference."""
    _plain(arr, "fillMaskAndNaNWithZero")[...] = _ops().op_fill_zero(arr)


def fill_nan_no_negative(arr, region_maskG, negative_to_zero=True):
    """Non-finite, land/coast (mask > 10) and lake (mask < 1) cells -> NaN, then negative -> 0 if asked; in place
    (NESOSIM.py:141-166); returns None like the reference."""
    _plain(arr, "fill_nan_no_negative")[...] = _ops().op_fill_nan_no_negative(arr, region_maskG, negative_to_zero)


def smooth_snow(arr, stddev_val=1, x_size_val=3, y_size_val=3):
    """astropy ``convolve(arr, Gaussian2DKernel(stddev_val, x_size=3, y_size=3))`` (NESOSIM.py:170-187): both branches
    of astropy's routine (plain, and NaN-interpolating when the input holds a NaN); returns a new array."""
    if int(x_size_val) != 3 or int(y_size_val) != 3:
        raise ValueError("the native smoother is the reference's 3x3 kernel (x_size_val = y_size_val = 3)")
    return _ops().smooth(arr, stddev=float(stddev_val))


def calcDynamics(driftGday, snowDepthsT, dx):
    """Advection and divergence of both layers for one day (NESOSIM.py:189-222), NaN/inf filled with zero: returns
    ``(snowAdvAllT, snowDivAllT)``, each ``(2, ny, nx)``."""
    return _ops().op_dynamics(driftGday, snowDepthsT, dx, deltaT)


def densityCalc(snowDepthsT, iceConcDayT, region_maskT):
    """Bulk density of the two layers (NESOSIM.py:458-473); ``iceConcDayT`` is accepted and unused, as in the reference."""
    return _ops().op_density(snowDepthsT, region_maskT, snowDensityFresh, snowDensityOld, minSnowD)


# ------------------------------------------------------------------------------------------ the driver

def main(year1, month1, day1, year2, month2, day2, outPathT='.', forcingPathT='.', anc_data_pathT='../anc_data/', figPathT='../Figures/',
         precipVar='ERA5', windVar='ERA5', driftVar='OSISAF', concVar='CDR', icVar='ERAI', densityTypeT='variable',
         outStr='', extraStr='', IC=2, windPackFactorT=0.1, windPackThreshT=5., leadLossFactorT=0.1, atmLossFactorT=2.2e-8, dynamicsInc=1, leadlossInc=1,
         windpackInc=1, atmlossInc=0, saveData=1, plotBudgets=1, plotdaily=1, saveFolder='', dx=50000, scaleCS=False):
    """The reference's model driver (NESOSIM.py:495-659) with the day loop replaced by one GPU season call.

    Returns ``None`` like the reference; the products are ``budgets/<saveStr>.nc`` and
    ``final/NESOSIMv11_<dates>.nc`` written by the reference's own writers, plus its figures."""
    from scipy.interpolate import griddata
    cF = _cF()
    global forcingPath, outPath, ancDataPath, figpath
    global snowDensityFresh, snowDensityOld, minSnowD, minConc, leadLossFactor, atmLossFactor, windPackThresh, windPackFactor, deltaT

    xptsG, yptsG, latG, lonG, proj = cF.create_grid(dxRes=dx)
    nx = xptsG.shape[0]
    ny = xptsG.shape[1]
    dxStr = str(int(dx / 1000)) + 'km'
    outPath = outPathT + dxStr + '/'
    forcingPath = forcingPathT + dxStr + '/'
    ancDataPath = anc_data_pathT
    _say(nx, ny, dxStr)
    _say('OutPath:', outPath)
    _say('forcingPath:', forcingPath)
    _say('ancDataPath:', ancDataPath)

    snowDensityFresh = 200.
    snowDensityOld = 350.
    minSnowD = 0.02
    minConc = 0.15
    deltaT = 60. * 60. * 24.
    # (the reference unpacks three values from a function that returns five, NESOSIM.py:535 / utils.py:1378)
    region_mask, xptsI, yptsI = cF.get_region_mask_pyproj(anc_data_pathT, proj, xypts_return=1)[:3]
    region_maskG = griddata((xptsI.flatten(), yptsI.flatten()), region_mask.flatten(), (xptsG, yptsG), method='nearest')
    leadLossFactor = leadLossFactorT
    windPackThresh = windPackThreshT
    windPackFactor = windPackFactorT
    atmLossFactor = atmLossFactorT

    startDay, numDays, numDaysYear1, dateOut = cF.getDays(year1, month1, day1, year2, month2, day2)
    _say(startDay, numDays, numDaysYear1, dateOut)
    dates = [int((datetime.datetime(year1, month1 + 1, day1 + 1) + datetime.timedelta(x)).strftime('%Y%m%d'))
             for x in range(numDays)]

    CSstr = ''
    scale_factors = None
    if scaleCS:
        scale_factors = _open_scale_factors('{}scale_coeffs_{}_{}_v2.nc'.format(ancDataPath, precipVar, dxStr))
        CSstr = 'CSscaled'

    saveStrNoDate = (precipVar + CSstr + 'sf' + windVar + 'winds' + driftVar + 'drifts' + concVar + 'sic' + 'rho' + densityTypeT +
                     '_IC' + str(IC) + '_DYN' + str(dynamicsInc) + '_WP' + str(windpackInc) + '_LL' + str(leadlossInc) +
                     '_AL' + str(atmlossInc) + '_WPF' + str(windPackFactorT) + '_WPT' + str(windPackThreshT) +
                     '_LLF' + str(leadLossFactorT) + '-' + dxStr + extraStr + outStr)
    saveStr = saveStrNoDate + '-' + dateOut
    _say('Saving to:', saveStr)
    savePath = outPath + saveFolder + '/' + saveStrNoDate
    for sub in ('/budgets/', '/final/'):
        if not os.path.exists(savePath + sub):
            os.makedirs(savePath + sub)
    figpath = figPathT + '/Diagnostic/' + dxStr + '/' + saveStrNoDate + '/'
    for d in (figpath, figpath + '/daily_snow_depths/'):
        if not os.path.exists(d):
            os.makedirs(d)

    # ---- initial conditions (NESOSIM.py:589-609); the concentration mask uses the first model day
    _say('IC:', IC)
    ICSnowDepth = None
    if IC > 0:
        if IC == 1:
            ICSnowDepth = np.load(forcingPath + 'InitialConditions/AugSnow' + dxStr, allow_pickle=True)
            _say('Initialize with August Warren climatology')
        elif IC == 2:
            ICSnowDepth = np.load(forcingPath + 'InitialConditions/' + icVar + '/ICsnow' + dxStr + '-' + str(year1) + extraStr,
                                  allow_pickle=True)
            _say('Initialize with new v1.1 scaled initial conditions')
        ICSnowDepth = np.array(ma.filled(ICSnowDepth, np.nan), dtype=np.float64)

    # ---- stage the whole season once (the reference reads five files per day inside its loop)
    staged = stage_season(year1, year2, startDay, numDays, numDaysYear1, precipVar, windVar, concVar, driftVar, dxStr, extraStr,
                          scale_factors=scale_factors, clim_table=_clim_density_table() if densityTypeT == 'clim' else None)

    # ---- the season on the GPU: IC masking/split, numDays-1 budget steps, all eleven budget arrays
    eng = _engine(region_maskG, numDays, dx, densityTypeT, dynamicsInc, leadlossInc, windpackInc, atmlossInc)
    eng.set_forcing(staged["precip"], staged["conc"], staged["wind"], staged["drift"], staged["rho_clim"])
    out = eng.run_season(_params_row(), ICSnowDepth)
    res = {k: v[0].cpu().numpy() for k, v in out.items()}
    snowDepths, density = res["snowDepths"], res["density"]
    snowAcc, snowOcean, snowAdv, snowDiv = res["snowAcc"], res["snowOcean"], res["snowAdv"], res["snowDiv"]
    snowLead, snowAtm = res["snowLead"], res["snowAtm"]
    snowWindPack, snowWindPackLoss, snowWindPackGain = res["snowWindPack"], res["snowWindPackLoss"], res["snowWindPackGain"]
    precipDays, iceConcDays, windDays, tempDays = staged["precip"], staged["conc"], staged["wind"], staged["temp"]
    x = numDays - 2

    if plotdaily == 1:
        import cartopy.crs as ccrs
        for d in range(numDays - 1):
            cF.plot_gridded_cartopy(lonG, latG, snowDepths[d + 1, 0] + snowDepths[d + 1, 1],
                                    proj=ccrs.NorthPolarStereo(central_longitude=-45), date_string='',
                                    out=figpath + 'daily_snow_depths/snowTot_' + saveStrNoDate + str(d), units_lab='m',
                                    varStr='Snow depth', minval=0., maxval=0.6)
    if saveData == 1:
        cF.OutputSnowModelRaw(savePath, saveStr, snowDepths, density, precipDays, iceConcDays, windDays, snowAcc, snowOcean,
                              snowAdv, snowDiv, snowLead, snowAtm, snowWindPack)
        with np.errstate(divide='ignore', invalid='ignore'):
            depth_over_ice = (snowDepths[:, 0] + snowDepths[:, 1]) / iceConcDays
        cF.OutputSnowModelFinal(savePath, 'NESOSIMv11_' + dateOut, lonG, latG, xptsG, yptsG, snowDepths[:, 0] + snowDepths[:, 1],
                                depth_over_ice, density, iceConcDays, precipDays, windDays, tempDays, dates)
    if plotBudgets == 1:
        cF.plot_budgets_cartopy(lonG, latG, precipDays[x + 1], windDays[x + 1], snowDepths[x + 1], snowOcean[x + 1], snowAcc[x + 1],
                                snowDiv[x + 1], snowAdv[x + 1], snowLead[x + 1], snowAtm[x + 1], snowWindPack[x + 1],
                                snowWindPackLoss[x + 1], snowWindPackGain[x + 1], density[x + 1], dates[-1], figpath,
                                totalOutStr=saveStr)


def run_multiseason(yearS, yearE, month1, day1, month2, day2, rank=0, world=1, **main_kwargs):
    """The loop of the reference's ``run_multiseason.py`` (``source/run_multiseason.py:39-50``: one independent
    ``main`` per start year, year2 = year1 + 1), with the seasons dealt round-robin over ``world`` ranks -- one process
    per GPU, no communication.  Returns the start years this rank ran."""
    from . import sharding
    mine = sharding.season_assignment(list(range(yearS, yearE + 1)), rank, world)
    for y in mine:
        main(year1=y, month1=month1, day1=day1, year2=y + 1, month2=month2, day2=day2, **main_kwargs)
    return mine

"""Grid, calendar and region-mask helpers the driver needs around the hot path.

These are *host-side, once-per-season* pieces that ``NESOSIM.main`` calls before the
day loop.  The reference gets them from pyproj (``utils.create_grid``, utils.py:870-890;
``utils.get_region_mask_pyproj``, utils.py:1345-1380) and ``utils.getDays`` (utils.py:241-256).  pyproj is
optional here: EPSG:3413 (WGS84 polar stereographic, lat_ts=70N, lon_0=-45) has a closed form, written out
below, which reproduces the reference's 100 km corners +-4460030.963 m (source/gridding/nohup.out:3-4).
"""
import datetime
import os

import numpy as np

_A = 6378137.0
_F = 1.0 / 298.257223563
_E2 = 2 * _F - _F * _F
_E = np.sqrt(_E2)
_LAT_TS = np.deg2rad(70.0)
_LON_0 = -45.0


def _t(phi):
    s = np.sin(phi)
    return np.tan(np.pi / 4 - phi / 2) / ((1 - _E * s) / (1 + _E * s)) ** (_E / 2)


_MC = np.cos(_LAT_TS) / np.sqrt(1 - _E2 * np.sin(_LAT_TS) ** 2)
_TC = _t(_LAT_TS)


class Proj3413:
    """Callable with pyproj.Proj's ``p(lon, lat)`` / ``p(x, y, inverse=True)`` convention for EPSG:3413."""

    def __call__(self, a, b, inverse=False):
        a = np.asarray(a, dtype=float)
        b = np.asarray(b, dtype=float)
        if not inverse:
            lam = np.deg2rad(a - _LON_0)
            rho = _A * _MC * _t(np.deg2rad(b)) / _TC
            return rho * np.sin(lam), -rho * np.cos(lam)
        rho = np.hypot(a, b)
        t = rho * _TC / (_A * _MC)
        phi = np.pi / 2 - 2 * np.arctan(t)
        for _ in range(8):
            s = np.sin(phi)
            phi = np.pi / 2 - 2 * np.arctan(t * ((1 - _E * s) / (1 + _E * s)) ** (_E / 2))
        lon = _LON_0 + np.rad2deg(np.arctan2(a, -b))
        lon = (lon + 180.0) % 360.0 - 180.0
        return lon, np.rad2deg(phi)


def create_grid(epsg_string='3413', dxRes=50000, lllat=36, llon=-90, urlat=36, urlon=90, verbose=False):
    """Square model grid covering the corner lat/lons (utils.py:870-890): returns x, y, lats, lons, proj.

    ``x``/``y`` are float32 like the reference's (they come from ``np.indices(..., np.float32)``); the hot path
    only ever uses the scalar ``dxRes``.
    """
    if str(epsg_string) != '3413':
        raise ValueError("only EPSG:3413 has a closed form here")
    p = Proj3413()
    llcrn = tuple(float(v) for v in p(llon, lllat))
    urcrn = tuple(float(v) for v in p(urlon, urlat))
    if verbose:
        print(llcrn)
        print(urcrn)
    nx = int((urcrn[0] - llcrn[0]) / dxRes) + 1
    ny = int((urcrn[1] - llcrn[1]) / dxRes) + 1
    if verbose:
        print(nx, ny)
    x = llcrn[0] + dxRes * np.indices((ny, nx), np.float32)[1]
    y = llcrn[1] + dxRes * np.indices((ny, nx), np.float32)[0]
    lons, lats = p(x, y, inverse=True)
    return x, y, lats, lons, p


def grid_shape(dx):
    """(ny, nx) of the reference grid at spacing ``dx`` metres: 90^2 @100 km, 357^2 @25 km, 1785^2 @5 km."""
    p = Proj3413()
    ll = p(-90, 36)
    ur = p(90, 36)
    return (int((float(ur[1]) - float(ll[1])) / dx) + 1, int((float(ur[0]) - float(ll[0])) / dx) + 1)


_LEAP = (1976, 1980, 1984, 1988, 1992, 1996, 2000, 2004, 2008, 2012, 2016, 2020)


def getLeapYr(year):
    """Days in ``year`` by the reference's table (utils.py:232-238) -- it ends at 2020, kept as is."""
    return 366 if year in _LEAP else 365


def getDays(year1, month1, day1, year2, month2, day2):
    """0-based month/day in, (startDay, numDays, numDaysYear1, 'ddmmYYYY-ddmmYYYY') out (utils.py:241-256)."""
    d0 = datetime.datetime(year1, 1, 1)
    d1 = datetime.datetime(year1, month1 + 1, day1 + 1)
    d2 = datetime.datetime(year2, month2 + 1, day2 + 1)
    fmt = '%d%m%Y'
    return (d1 - d0).days, (d2 - d1).days + 1, getLeapYr(year1), d1.strftime(fmt) + '-' + d2.strftime(fmt)


def get_region_mask(anc_data_path, proj, xypts_return=0):
    """NSIDC region mask (448x304 uint8 after a 300-byte header) and its projected coordinates
    (utils.py:1345-1380).  Returns the reference's 5-tuple when ``xypts_return==1``."""
    raw = np.fromfile(os.path.join(anc_data_path, 'region_n.msk'), dtype='uint8')
    region_mask = np.reshape(raw[300:], [448, 304])
    if xypts_return != 1:
        return region_mask
    lats = np.reshape(np.fromfile(os.path.join(anc_data_path, 'psn25lats_v3.dat'), dtype='<i4') / 100000., [448, 304])
    lons = np.reshape(np.fromfile(os.path.join(anc_data_path, 'psn25lons_v3.dat'), dtype='<i4') / 100000., [448, 304])
    xpts, ypts = proj(lons, lats)
    return region_mask, xpts, ypts, lons, lats


def region_mask_on_grid(anc_data_path, xptsG, yptsG, proj=None):
    """Nearest-neighbour regrid of the region mask onto the model grid, as ``main`` does (NESOSIM.py:535-536)."""
    from scipy.interpolate import griddata
    proj = proj or Proj3413()
    region_mask, xptsI, yptsI = get_region_mask(anc_data_path, proj, xypts_return=1)[:3]
    return griddata((xptsI.flatten(), yptsI.flatten()), region_mask.flatten(), (xptsG, yptsG), method='nearest')


_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def bundled_region_mask(dx):
    """Region mask fixture shipped with the package (made by tools/make_mask_fixtures.py from
    anc_data/region_n.msk exactly as ``region_mask_on_grid`` does) for the synthetic benchmark shapes.
    100 km and 25 km are stored; finer grids are the 25 km mask sampled at nearest cell centres."""
    km = int(dx / 1000)
    path = os.path.join(_DATA_DIR, 'region_mask_%dkm.npy' % km)
    if os.path.isfile(path):
        return np.load(path)
    base = np.load(os.path.join(_DATA_DIR, 'region_mask_25km.npy'))
    ny, nx = grid_shape(dx)
    by, bx = base.shape
    # cell centres of both grids share the lower-left corner; nearest 25 km cell
    iy = np.clip(np.rint(np.arange(ny) * (dx / 25000.0)).astype(int), 0, by - 1)
    ix = np.clip(np.rint(np.arange(nx) * (dx / 25000.0)).astype(int), 0, bx - 1)
    return base[np.ix_(iy, ix)]

"""Seeded synthetic forcing of the reference's shapes.

The reference's test forcing (``test_forcings.zip``) is absent from the checkout, so every parity test and
every benchmark number in this repository runs on forcing made here.  The fields are built to exercise what
``loadData`` (NESOSIM.py:379-456) can hand to ``calcBudget``: concentration with non-finite -> 0 and an
open-water margin, NaN snowfall outside the reanalysis hull (grid corners), winds straddling the 5 m/s packing
threshold (some cells exactly 5.0), drift that is NaN over open water / land and on ~2 % of days entirely
missing, and an all-NaN temperature placeholder.
"""
import numpy as np

from . import grid as _grid

__all__ = ["region_mask", "make_season", "make_ic", "ensemble_params", "season_lengths"]


def season_lengths():
    """numDays for the two season definitions in BASELINE.md §2."""
    return {"run_oneseason": 242, "aug15_may1": 260}


def region_mask(dx=None, shape=None, kind="auto"):
    """uint8 region codes (0 lakes, 1-10 ocean regions, 8 Arctic Ocean, 11 land, 12 coast).

    ``kind='auto'``: the bundled regrid of ``anc_data/region_n.msk`` for the reference grids (give ``dx``).
    ``kind='disc'``: SURVEY.md §8d's fallback for arbitrary ``shape`` -- a disc r < 0.75*n/2 of code 8 inside
    land (11), with a short arc of 0 ("lakes").
    """
    if kind == "auto" and dx is not None:
        return np.ascontiguousarray(_grid.bundled_region_mask(dx).astype(np.uint8))
    ny, nx = shape
    yy, xx = np.mgrid[0:ny, 0:nx]
    cy, cx = (ny - 1) / 2.0, (nx - 1) / 2.0
    r = np.hypot((yy - cy) / (ny / 2.0), (xx - cx) / (nx / 2.0))
    m = np.where(r < 0.75, 8, 11).astype(np.uint8)
    ang = np.arctan2(yy - cy, xx - cx)
    m[(np.abs(r - 0.80) < 1.5 / max(ny, nx)) & (np.abs(ang) < 0.3)] = 0
    return m


def _smooth(a, sigma):
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(a, sigma=sigma, mode="nearest")


def make_season(mask, num_days, seed=0, missing_drift_frac=0.02, dtype=np.float64):
    """Forcing dict ``precip, conc, wind`` (T,ny,nx), ``drift`` (T,2,ny,nx), ``temp`` (T,ny,nx) = NaN.

    Slot ``T-1`` is the day after the last step (NESOSIM.py:645-649).
    """
    mask = np.asarray(mask)
    ny, nx = mask.shape
    T = int(num_days)
    rng = np.random.default_rng(seed)
    land = (mask > 10) | (mask < 1)
    yy, xx = np.mgrid[0:ny, 0:nx]
    r = np.hypot((yy - (ny - 1) / 2.0) / (ny / 2.0), (xx - (nx - 1) / 2.0) / (nx / 2.0))

    # ice concentration: radial pack + noise, growing through the season
    grow = np.linspace(0.0, 0.35, T)[:, None, None]
    conc = np.clip(1.05 + grow - 1.5 * r[None] + 0.05 * rng.standard_normal((T, ny, nx)), 0.0, 1.0)
    conc[conc < 0.15] = 0.0
    conc[:, land] = np.nan                       # product files carry NaN over land ...
    conc[~np.isfinite(conc)] = 0.0               # ... and loadData zeroes it (NESOSIM.py:430)

    precip = rng.gamma(0.5, 2.0, size=(T, ny, nx))
    k = 0 if min(ny, nx) < 6 else max(2, ny // 30)   # NaN corners (skipped on toy grids)
    for sy in (slice(0, k), slice(ny - k, ny)):
        for sx in (slice(0, k), slice(nx - k, nx)):
            precip[:, sy, sx] = np.nan

    wind = rng.gamma(4.0, 1.5, size=(T, ny, nx))
    n_exact = max(4, (ny * nx) // 2000)
    for t in range(T):
        idx = rng.integers(0, ny * nx, size=n_exact)
        wind[t].reshape(-1)[idx] = 5.0

    drift = np.empty((T, 2, ny, nx))
    sig = max(1.0, ny / 45.0)
    for t in range(T):
        for c in range(2):
            f = _smooth(rng.standard_normal((ny, nx)), sig)
            drift[t, c] = 0.1 * f / max(f.std(), 1e-12)
    nodrift = (conc == 0.0) | land[None]
    drift[:, 0][nodrift] = np.nan
    drift[:, 1][nodrift] = np.nan
    missing = rng.random(T) < missing_drift_frac
    if T > 3:
        missing[3] = True                       # always exercise the missing-file day (NESOSIM.py:437-441)
    drift[missing] = np.nan

    temp = np.full((T, ny, nx), np.nan)
    out = {"precip": precip, "conc": conc, "wind": wind, "drift": drift, "temp": temp}
    if dtype != np.float64:
        out = {k_: v.astype(dtype) for k_, v in out.items()}
    return out


def make_ic(mask, seed=0):
    """Initial total snow depth: smooth 0-0.10 m field on the Arctic Ocean (code 8), zero elsewhere."""
    mask = np.asarray(mask)
    rng = np.random.default_rng(seed + 7919)
    f = _smooth(rng.random(mask.shape), max(1.0, mask.shape[0] / 30.0))
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
    return np.where(mask == 8, 0.10 * f, 0.0)


def ensemble_params(n_members, seed=0):
    """(M,4) float64 rows [windPackFactor, windPackThresh, leadLossFactor, atmLossFactor], log-uniform in the
    calibration ranges of SURVEY.md §8d: WPF in [1e-7,1e-6], LLF in [5e-8,6e-7], ALF in [1e-9,1e-7], WPT=5."""
    rng = np.random.default_rng(seed + 104729)

    def logu(lo, hi):
        return np.exp(rng.uniform(np.log(lo), np.log(hi), size=n_members))

    p = np.empty((n_members, 4))
    p[:, 0] = logu(1e-7, 1e-6)
    p[:, 1] = 5.0
    p[:, 2] = logu(5e-8, 6e-7)
    p[:, 3] = logu(1e-9, 1e-7)
    return p

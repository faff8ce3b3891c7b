"""Generate the golden fixtures in this directory from the REFERENCE'S OWN FUNCTIONS executed verbatim.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these files are the pinning: every output
array below is produced by ``/root/reference/source/NESOSIM.py``'s functions (imported through
``oracle/ref_loader.py``, astropy replaced by ``oracle/astropy_restated.py``), never by the oracle restatement
or the CUDA code.  Inputs are stored next to the outputs so the fixtures do not depend on the random generator.

  kat_functions.npz      per-function known answers for rows a2-a10 of SURVEY.md §8 (edge cases listed there)
  season_small.npz       42x37 grid, 16 days, all inputs + all 15 arrays (two parameter sets, switches, clim)
  season_100km_digest.npz  90x90, 242 days (run_oneseason.py dates): inputs by seed, outputs as final/selected
                           planes + sha256 of every full array
  season_25km_digest.npz   357x357, 260 days (Aug 15 - May 1), run_multiseason parameters: sha256 of every full array,
                           of the last slot, and sampled rows of the last slot
  steps_5km_digest.npz     1785x1785 (the bundled 5 km mask), 3 steps: sha256 of every full array + sampled rows
(the two large cases take a few minutes of CPU: `python tests/golden/make_golden.py large`)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader                     # noqa: E402
from nesosim_b200 import synthetic as S            # noqa: E402

ref = ref_loader.load_reference()
ANC = "/root/reference/anc_data/"


def canon_sha(a):
    """sha256 of the array with every NaN replaced by one canonical quiet NaN and -0.0 by +0.0."""
    b = np.array(a, dtype=np.float64, copy=True)
    b[np.isnan(b)] = np.nan
    b = b + 0.0
    return hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest()


def kat_functions():
    rng = np.random.default_rng(42)
    out = {}
    ref_loader.set_globals(ref, 5.8e-7, 5., 1.45e-7, 2.2e-8)
    # ---- a7-a9: wind terms, including W == threshold, NaN wind (0*NaN), NaN depth, inf
    n = 256
    h0 = np.abs(rng.standard_normal(n)) * 0.3
    W = rng.gamma(4.0, 1.5, n)
    C = rng.random(n)
    W[:8] = 5.0
    W[8:12] = np.nan
    h0[12:16] = np.nan
    W[16] = np.inf
    h0[17], W[17] = 0.0, np.inf
    C[18] = 1.0
    C[19] = 0.0
    with np.errstate(all="ignore"):
        wl, wg, wn = ref.calcWindPacking(W, h0)
        out.update(wt_h0=h0, wt_W=W, wt_C=C, wt_lead=ref.calcLeadLoss(h0, W, C), wt_atm=ref.calcAtmLoss(h0, W),
                   wt_wpl=wl, wt_wpg=wg, wt_wpn=wn, wt_params=np.array([5.8e-7, 5., 1.45e-7, 2.2e-8]))
    # ---- a2-a4: dynamics with NaN blocks, a NaN drift row, inf, first/last row/col gradients
    ny, nx = 11, 14
    h = np.abs(rng.standard_normal((2, ny, nx))) * 0.2
    h[:, 4:7, 5:8] = np.nan
    d = 0.1 * rng.standard_normal((2, ny, nx))
    d[:, 9, :] = np.nan
    d[0, 2, 2] = np.inf
    for dx in (100000, 25000):
        with np.errstate(all="ignore"):
            adv, div = ref.calcDynamics(d.copy(), h.copy(), dx)
        out["dyn_adv_%d" % dx] = adv
        out["dyn_div_%d" % dx] = div
    out.update(dyn_h=h, dyn_drift=d)
    # all-NaN drift day -> zero dynamics
    dn = np.full((2, ny, nx), np.nan)
    with np.errstate(all="ignore"):
        adv, div = ref.calcDynamics(dn, h.copy(), 100000)
    out.update(dyn_adv_nandrift=adv, dyn_div_nandrift=div)
    # ---- a4/a6 fills
    a = rng.standard_normal(300)
    a[::7] = np.nan
    a[::11] = np.inf
    a[::13] = -np.inf
    a[5] = -0.0
    mask = rng.integers(0, 13, 300).astype(np.float64)
    z = a.copy()
    ref.fillMaskAndNaNWithZero(z)
    f1 = a.copy()
    ref.fill_nan_no_negative(f1, mask, negative_to_zero=True)
    f0 = a.copy()
    ref.fill_nan_no_negative(f0, mask, negative_to_zero=False)
    out.update(fill_in=a, fill_mask=mask, fill_zero=z, fill_nan_neg=f1, fill_nan_noneg=f0)
    # ---- a10 density: 0/0, straddling minSnowD, clamps, NaN, land/lake
    hd = np.abs(rng.standard_normal((2, 300))) * 0.05
    hd[0, :50] = 0.0
    hd[1, :25] = 0.0
    hd[:, 60:70] = 0.01
    hd[0, 70], hd[1, 70] = 0.02, 0.0
    hd[0, 71], hd[1, 71] = 0.0, 0.02
    hd[0, 72], hd[1, 72] = 0.019999999999999997, 0.0
    hd[:, 80:90] = np.nan
    with np.errstate(all="ignore"):
        out.update(dens_h=hd, dens_out=ref.densityCalc(hd, None, mask))
    # ---- a5 smooth_snow: plain branch, corners; interpolate branch: isolated NaN, 4x4 NaN block, NaN at edges
    s = rng.standard_normal((13, 17))
    out.update(sm_plain_in=s, sm_plain_out=ref.smooth_snow(s.copy()))
    sn = s.copy()
    sn[5, 7] = np.nan
    sn[8:12, 2:6] = np.nan
    sn[0, 0] = sn[-1, -1] = sn[0, 9] = np.nan
    out.update(sm_nan_in=sn, sm_nan_out=ref.smooth_snow(sn.copy()))
    si = s.copy()
    si[3, 3], si[9, 9] = np.inf, -np.inf
    with np.errstate(all="ignore"):
        out.update(sm_inf_in=si, sm_inf_out=ref.smooth_snow(si.copy()))
    k = ref.Gaussian2DKernel(x_stddev=1, x_size=3, y_size=3).array
    out.update(kernel=k, kernel_sum=np.array(k.sum()))
    np.savez_compressed(os.path.join(HERE, "kat_functions.npz"), **out)
    print("kat_functions.npz", len(out), "arrays")


def season_small():
    ny, nx, T = 42, 37, 16
    mask = S.region_mask(shape=(ny, nx), kind="disc")
    mask[3:6, 30:33] = 12       # a coast patch
    F = S.make_season(mask, T, seed=7)
    ic = S.make_ic(mask, seed=7) * 4.0
    out = {"mask": mask, "ic": ic, "dx": np.array(100000), "precip": F["precip"], "conc": F["conc"],
           "wind": F["wind"], "drift": F["drift"]}
    w99 = np.loadtxt(os.path.join(ANC, "W99_density.csv"), delimiter=",", skiprows=1)
    days = [(243 + x) % 365 for x in range(T)]          # dayT passed to calcBudget (0-based DOY), NESOSIM.py:615-620
    cases = {
        "oneseason": dict(p=(5.8e-7, 5, 2.9e-7, 2.2e-8), flags=dict(atmlossInc=0)),
        "multiseason": dict(p=(5.8e-7, 5., 1.45e-7, 2.2e-8), flags=dict(atmlossInc=1)),
        "nodyn": dict(p=(5.8e-7, 5., 1.45e-7, 2.2e-8), flags=dict(atmlossInc=1, dynamicsInc=0)),
        "clim": dict(p=(5.8e-7, 5., 1.45e-7, 2.2e-8), flags=dict(atmlossInc=1, densityType="clim")),
    }
    for name, c in cases.items():
        ref_loader.set_globals(ref, *c["p"], ancDataPath=ANC)
        R = ref_loader.run_reference_season(ref, F, ic, mask.astype(np.float64), 100000, c["flags"], day_of_year=days)
        out[name + "__params"] = np.array(c["p"], dtype=np.float64)
        for k, v in R.items():
            if k in ("precipDays", "iceConcDays", "windDays", "tempDays"):
                continue
            out[name + "__" + k] = v
    # what utils.densityClim(dayT) returns for the days above: 1000*Density.iloc[dayT-1]
    out["clim__rho"] = np.array([1000 * w99[(d - 1) % len(w99), 1] for d in days])
    np.savez_compressed(os.path.join(HERE, "season_small.npz"), **out)
    print("season_small.npz", os.path.getsize(os.path.join(HERE, "season_small.npz")) // 1024, "KiB")


def season_100km_digest():
    mask = S.region_mask(dx=100000)
    T = 242
    seed = 2018
    F = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed)
    out = {"seed": np.array(seed), "T": np.array(T)}
    for k in ("precip", "conc", "wind", "drift"):
        out["in_sha__" + k] = np.array(canon_sha(F[k]))
    out["in_sha__ic"] = np.array(canon_sha(ic))
    for name, p, atm in (("oneseason", (5.8e-7, 5, 2.9e-7, 2.2e-8), 0), ("multiseason", (5.8e-7, 5., 1.45e-7, 2.2e-8), 1)):
        ref_loader.set_globals(ref, *p)
        R = ref_loader.run_reference_season(ref, F, ic, mask.astype(np.float64), 100000, dict(atmlossInc=atm))
        out[name + "__params"] = np.array(p, dtype=np.float64)
        for k, v in R.items():
            if k in ("precipDays", "iceConcDays", "windDays", "tempDays"):
                continue
            out[name + "__sha__" + k] = np.array(canon_sha(v))
            out[name + "__last__" + k] = v[-1]
            out[name + "__day60__" + k] = v[60]
    np.savez_compressed(os.path.join(HERE, "season_100km_digest.npz"), **out)
    print("season_100km_digest.npz", os.path.getsize(os.path.join(HERE, "season_100km_digest.npz")) // 1024, "KiB")


def large_digest(fname, dx, T, seed, params, atm, sample_rows):
    """A configuration at BASELINE's full grid size through the reference's own loop: only digests are kept (sha256 of
    every full array and of its last slot) plus a few sampled rows of the last slot for diagnosis."""
    mask = S.region_mask(dx=dx)
    F = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed)
    out = {"seed": np.array(seed), "T": np.array(T), "dx": np.array(dx), "params": np.array(params, dtype=np.float64),
           "atmlossInc": np.array(atm), "sample_rows": np.array(sample_rows), "mask_sha": np.array(canon_sha(mask.astype(np.float64)))}
    for k in ("precip", "conc", "wind", "drift"):
        out["in_sha__" + k] = np.array(canon_sha(F[k]))
    out["in_sha__ic"] = np.array(canon_sha(ic))
    ref_loader.set_globals(ref, *params)
    R = ref_loader.run_reference_season(ref, F, ic, mask.astype(np.float64), dx, dict(atmlossInc=atm))
    for k, v in R.items():
        if k in ("precipDays", "iceConcDays", "windDays", "tempDays"):
            continue
        out["sha__" + k] = np.array(canon_sha(v))
        out["sha_last__" + k] = np.array(canon_sha(v[-1]))
        out["rows_last__" + k] = v[-1][..., sample_rows, :]
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, os.path.getsize(os.path.join(HERE, fname)) // 1024, "KiB")


def season_25km_digest():
    large_digest("season_25km_digest.npz", 25000, 260, 2025, (5.8e-7, 5., 1.45e-7, 2.2e-8), 1, list(range(5, 357, 50)))


def steps_5km_digest():
    large_digest("steps_5km_digest.npz", 5000, 4, 5005, (5.8e-7, 5., 1.45e-7, 2.2e-8), 1, list(range(7, 1785, 300)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        season_25km_digest()
        steps_5km_digest()
    else:
        kat_functions()
        season_small()
        season_100km_digest()

"""Final-product diagnostics: the fused GPU pass against the numpy oracle, and the oracle against the
reference's own OutputSnowModelFinal executed verbatim with a recording stand-in for netCDF4."""
import sys
import types

import numpy as np
import pytest

from nesosim_b200 import synthetic as S
from oracle import final_products as FP
from oracle import nesosim_oracle as O
from oracle import ref_loader


def season(T=9, seed=13):
    mask = S.region_mask(dx=100000)
    forcing = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed) * 4
    out = O.run_season(forcing, ic, mask, 100000, O.Params(), O.Flags(atmlossInc=1))
    conc = forcing["conc"].copy()
    conc[2, 40:44, 40:44] = 0.15            # thresholds exactly
    conc[3, 40:44, 40:44] = 0.5
    conc[4, 10, 10] = np.nan
    h = out["snowDepths"].copy()
    h[5, 0, 45, 45] = 0.123456789           # rounding cases: half-way and float32-inexact values
    h[5, 1, 45, 45] = 0.00005
    h[6, :, 46, 46] = [0.00125, 0.0]
    return h, out["density"], conc, forcing["precip"], forcing["wind"]


def same32(a, b):
    return a.dtype == np.float32 and b.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


class _Var:
    def __init__(self, dtype):
        self.dtype, self.value = dtype, None

    def __setitem__(self, key, val):
        self.value = np.asarray(np.ma.filled(val, np.nan) if np.ma.isMaskedArray(val) else val, dtype=self.dtype)

    def __setattr__(self, k, v):
        object.__setattr__(self, k, v)


class _Dataset:
    last = None

    def __init__(self, *a, **k):
        self.vars = {}
        _Dataset.last = self

    def createVariable(self, name, dtype, dims=()):
        self.vars[name] = _Var(dtype)
        return self.vars[name]

    def createDimension(self, *a):
        pass

    def close(self):
        pass


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
def test_oracle_matches_reference_writer(tmp_path, capsys):
    fake = types.ModuleType("netCDF4")
    fake.Dataset = _Dataset
    saved = sys.modules.get("netCDF4")
    sys.modules["netCDF4"] = fake
    try:
        ref_loader.load_reference()
        utils = sys.modules["utils"]
        h, dens, conc, precip, wind = season()
        exp = FP.final_fields(h, dens, conc, precip, wind)
        T, ny, nx = conc.shape
        lon = np.zeros((ny, nx))
        with np.errstate(all="ignore"):
            utils.OutputSnowModelFinal(str(tmp_path), "x", lon, lon, lon, lon, h[:, 0] + h[:, 1], (h[:, 0] + h[:, 1]) / conc,
                                       dens.copy(), conc.copy(), precip, wind, np.full(conc.shape, np.nan), list(range(T)))
        got = _Dataset.last.vars
        for ours, theirs in (("snow_volume", "snow_volume"), ("snow_depth", "snow_depth"), ("snow_density", "snow_density"),
                             ("ice_concentration", "ice_concentration"), ("precipitation", "precipitation"),
                             ("wind_speed", "wind_speed")):
            assert same32(exp[ours], got[theirs].value), ours
    finally:
        if saved is None:
            sys.modules.pop("netCDF4", None)
        else:
            sys.modules["netCDF4"] = saved
        sys.modules.pop("utils", None)
        sys.modules.pop("NESOSIM", None)
    capsys.readouterr()


def test_oracle_masks_and_rounding():
    h, dens, conc, precip, wind = season()
    f = FP.final_fields(h, dens, conc, precip, wind)
    assert all(v.dtype == np.float32 for v in f.values())
    with np.errstate(invalid="ignore"):
        low = conc < 0.5
    assert np.isnan(f["snow_volume"][low]).all() and np.isnan(f["snow_depth"][low]).all() and np.isnan(f["snow_density"][low]).all()
    assert not np.isnan(f["snow_volume"][3, 40:44, 40:44]).all()          # conc == 0.5 is kept
    assert np.isnan(f["ice_concentration"][conc < 0.15]).all() and f["ice_concentration"][2, 41, 41] == np.float32(0.15)
    assert np.isnan(f["snow_depth"][4, 10, 10])                              # NaN concentration: depth NaN, volume kept
    keep = FP.final_fields(h, dens, conc, precip, wind, ice_conc_mask=0)
    assert np.isfinite(keep["ice_concentration"][conc < 0.15]).all()


@pytest.mark.gpu
def test_gpu_final_products_match_oracle(cuda):
    from nesosim_b200 import engine as E
    h, dens, conc, precip, wind = season()
    for m in (0.5, 0.0, 0.3):
        exp = FP.final_fields(h, dens, conc, precip, wind, ice_conc_mask=m)
        got = {k: v.cpu().numpy() for k, v in E.final_products(h, dens, conc, precip, wind, ice_conc_mask=m).items()}
        for k in exp:
            assert same32(got[k], exp[k]), (k, m)
    # an ensemble: members stacked, one shared forcing
    h3 = np.stack([h, h * 0.5, h * 2.0])
    d3 = np.stack([dens, dens, dens])
    got = {k: v.cpu().numpy() for k, v in E.final_products(h3, d3, conc, precip, wind).items()}
    for m in range(3):
        exp = FP.final_fields(h3[m], d3[m], conc, precip, wind)
        for k in exp:
            assert same32(got[k][m], exp[k]), (k, m)
    # straight from a season still resident on the device
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    forcing = S.make_season(mask, 9, seed=13)
    ic = S.make_ic(mask, seed=13) * 4
    eng = SnowBudgetEngine(mask, 9, 100000, atmlossInc=1)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = eng.run_season([[5.8e-7, 5., 2.9e-7, 2.2e-8]], ic)
    got = E.final_products(out["snowDepths"][0], out["density"][0], forcing["conc"], forcing["precip"], forcing["wind"])
    ref = O.run_season(forcing, ic, mask, 100000, O.Params(), O.Flags(atmlossInc=1))
    exp = FP.final_fields(ref["snowDepths"], ref["density"], forcing["conc"], forcing["precip"], forcing["wind"])
    for k in exp:
        assert same32(got[k].cpu().numpy(), exp[k]), k

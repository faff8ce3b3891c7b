"""GPU: the eight functions `calcBudget` is made of, called under the reference's names through the drop-in module
(`nesosim_b200.NESOSIM.calcLeadLoss` ... `densityCalc`, NESOSIM.py:51-222, 458-473), against the oracle on 2-D planes
with land, lakes, NaN / inf forcing, the wind threshold and empty cells.  (The native operators underneath are tested
one by one in test_gpu_parity.py::test_op_*; the plumbing of the wrappers against the reference run verbatim in
test_host_module.py::test_per_function_operators_have_the_reference_contracts.)"""
import numpy as np
import pytest

from oracle import astropy_restated as ar
from oracle import nesosim_oracle as O

pytestmark = pytest.mark.gpu


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a, b, equal_nan=True)


def test_reference_named_operators_on_the_gpu(cuda, monkeypatch):
    from nesosim_b200 import NESOSIM as N
    consts = dict(windPackFactor=5.8e-7, windPackThresh=5., leadLossFactor=1.45e-7, atmLossFactor=2.2e-8)
    for k, v in consts.items():
        monkeypatch.setattr(N, k, v)
    p = O.Params(**consts)
    rng = np.random.default_rng(17)
    ny, nx = 23, 31
    mask = rng.choice(np.array([0, 3, 8, 8, 8, 11, 12]), size=(ny, nx)).astype(float)
    h = np.abs(rng.normal(0.1, 0.1, (2, ny, nx)))
    h[:, mask > 10] = np.nan
    h[:, 2, 3] = 0.0                                   # 0/0 in the density
    h[0, 5, 5], h[1, 5, 5] = 0.012, 0.007              # sum just under minSnowD
    W = rng.gamma(4.0, 1.5, (ny, nx))
    W[0, :3] = 5.0                                     # exactly the threshold: not packed
    W[1, 1] = np.nan
    W[1, 2] = np.inf
    C = np.clip(rng.random((ny, nx)), 0, 1)
    drift = 0.1 * rng.normal(size=(2, ny, nx))
    drift[:, 4:6, 4:7] = np.nan
    with np.errstate(all="ignore"):
        assert same(N.calcLeadLoss(h[0], W, C), O.lead_loss(h[0], W, C, p))
        assert same(N.calcAtmLoss(h[0], W), O.atm_loss(h[0], W, p))
        for g, e in zip(N.calcWindPacking(W, h[0]), O.wind_packing(W, h[0], p)):
            assert same(g, e)
        for g, e in zip(N.calcDynamics(drift, h, 100000), O.calc_dynamics(drift, h, 100000, p)):
            assert same(g, e)
        assert same(N.densityCalc(h, C, mask), O.density_calc(h, C, mask, p))
        clean = np.nan_to_num(h[0], nan=0.0)
        assert same(N.smooth_snow(clean), O.smooth_snow(clean))
        k = ar.gaussian2d_kernel(x_stddev=0.5, y_stddev=0.5, theta=0.0, x_size=3, y_size=3)
        assert same(N.smooth_snow(h[0], stddev_val=0.5), ar.convolve_fill0(h[0], k))      # NaNs: the interpolating branch
        for negz in (True, False):
            a = rng.normal(size=(ny, nx))
            a[3, 3], a[4, 4] = np.inf, np.nan
            b = a.copy()
            assert N.fill_nan_no_negative(a, mask, negative_to_zero=negz) is None
            O.fill_nan_no_negative(b, mask, negative_to_zero=negz)
            assert same(a, b)
        a = rng.normal(size=(ny, nx))
        a[1, 2], a[2, 1], a[0, 0] = np.nan, np.inf, -np.inf
        b = a.copy()
        assert N.fillMaskAndNaNWithZero(a) is None
        O.fill_mask_nan_zero(b)
        assert same(a, b)

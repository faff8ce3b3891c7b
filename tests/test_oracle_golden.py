"""CPU: the oracle restatement against the golden vectors made from the reference's own functions
(tests/golden/make_golden.py), plus property checks on the restated astropy convolution."""
import numpy as np
import pytest

from golden_util import OUT_NAMES, SMALL_CASES, assert_identical, canon_sha, load
from nesosim_b200 import synthetic as S
from oracle import nesosim_oracle as O
from oracle.astropy_restated import convolve_fill0, convolve_fill0_scalar, gaussian2d_kernel


def P(row):
    return O.Params(windPackFactor=row[0], windPackThresh=row[1], leadLossFactor=row[2], atmLossFactor=row[3])


def test_kernel_constants():
    g = load("kat_functions.npz")
    k = gaussian2d_kernel(x_stddev=1, x_size=3, y_size=3)
    assert np.array_equal(k, g["kernel"]) and k.sum() == float(g["kernel_sum"])
    # SURVEY.md §8 row a5 values
    assert abs(k[0, 0] - 0.05854983) < 1e-8 and abs(k[0, 1] - 0.09653235) < 1e-8 and abs(k[1, 1] - 0.15915494) < 1e-8
    assert k.sum() == 0.7794836797093876
    assert np.array_equal(k, k.T) and np.array_equal(k, k[::-1, ::-1])


def test_kat_wind_terms():
    g = load("kat_functions.npz")
    p = P(g["wt_params"])
    with np.errstate(all="ignore"):
        assert_identical(O.lead_loss(g["wt_h0"], g["wt_W"], g["wt_C"], p), g["wt_lead"], "lead")
        assert_identical(O.atm_loss(g["wt_h0"], g["wt_W"], p), g["wt_atm"], "atm")
        l, gn, n = O.wind_packing(g["wt_W"], g["wt_h0"], p)
    assert_identical(l, g["wt_wpl"], "wpl")
    assert_identical(gn, g["wt_wpg"], "wpg")
    assert_identical(n, g["wt_wpn"], "wpn")
    assert g["wt_lead"][0] == 0 and np.isnan(g["wt_lead"][8])      # W == thresh not packed; 0*NaN is NaN


def test_kat_dynamics_and_gradient_formula():
    g = load("kat_functions.npz")
    for dx in (100000, 25000):
        adv, div = O.calc_dynamics(g["dyn_drift"].copy(), g["dyn_h"].copy(), dx, O.Params())
        assert_identical(adv, g["dyn_adv_%d" % dx], "adv")
        assert_identical(div, g["dyn_div_%d" % dx], "div")
    adv, div = O.calc_dynamics(np.full_like(g["dyn_drift"], np.nan), g["dyn_h"].copy(), 100000, O.Params())
    assert not adv.any() and not div.any() and not g["dyn_adv_nandrift"].any()
    rng = np.random.default_rng(0)
    f = rng.standard_normal((2, 9, 12))
    for dx in (100000, 25000, 5000, 7.25):
        for ax in (0, 1, 2):
            assert np.array_equal(O.gradient_explicit(f, dx, ax), np.gradient(f, dx, axis=ax))


def test_kat_fills_density():
    g = load("kat_functions.npz")
    a = g["fill_in"].copy()
    O.fill_mask_nan_zero(a)
    assert_identical(a, g["fill_zero"], "fill zero")
    for neg, key in ((True, "fill_nan_neg"), (False, "fill_nan_noneg")):
        a = g["fill_in"].copy()
        O.fill_nan_no_negative(a, g["fill_mask"], negative_to_zero=neg)
        assert_identical(a, g[key], key)
    assert_identical(O.density_calc(g["dens_h"], None, g["fill_mask"], O.Params()), g["dens_out"], "density")


def test_kat_smooth_both_branches():
    g = load("kat_functions.npz")
    with np.errstate(all="ignore"):
        for tag in ("plain", "nan", "inf"):
            assert_identical(O.smooth_snow(g["sm_%s_in" % tag]), g["sm_%s_out" % tag], tag)
    out = g["sm_nan_out"]
    assert np.isnan(out[9, 3]) and np.isnan(out[10, 4])     # centre of the 4x4 NaN block: bot == 0 keeps NaN
    assert np.isfinite(out[5, 7]) and np.isfinite(out[0, 0])


@pytest.mark.parametrize("variant", ["post_divide", "pre_normalised"])
def test_vectorised_convolve_equals_literal_loop(variant):
    rng = np.random.default_rng(5)
    g = gaussian2d_kernel()
    for shape in ((1, 1), (1, 7), (6, 1), (8, 9)):
        a = rng.standard_normal(shape)
        assert_identical(convolve_fill0(a, g, variant), convolve_fill0_scalar(a, g, variant), "plain %r" % (shape,))
        if a.size > 2:
            a.reshape(-1)[::3] = np.nan
            assert_identical(convolve_fill0(a, g, variant), convolve_fill0_scalar(a, g, variant), "nan %r" % (shape,))
    z = np.zeros((0, 5))
    assert convolve_fill0(z, g, variant).shape == (0, 5)      # empty input


def test_convolve_properties():
    rng = np.random.default_rng(6)
    g = gaussian2d_kernel()
    a = rng.standard_normal((12, 15))
    # interior of a constant field is preserved (normalised kernel); edges lose the padded weight
    c = convolve_fill0(np.full((7, 7), 3.0), g)
    assert np.allclose(c[1:-1, 1:-1], 3.0, rtol=1e-14) and c[0, 0] < 3.0
    # the two published normalisation orders agree far below the 1e-10 parity tolerance
    assert np.allclose(convolve_fill0(a, g), convolve_fill0(a, g, "pre_normalised"), rtol=1e-12, atol=1e-15)
    # linear in the input up to rounding
    b = rng.standard_normal((12, 15))
    assert np.allclose(convolve_fill0(a + b, g), convolve_fill0(a, g) + convolve_fill0(b, g), rtol=1e-12, atol=1e-14)
    # NaN branch on an all-NaN plane: the interior keeps NaN (bot == 0); edge cells see the zero padding, which
    # is not NaN and therefore counts in `bot`, so they come out 0/bot = 0
    n = convolve_fill0(np.full((4, 4), np.nan), g)
    assert np.isnan(n[1:-1, 1:-1]).all() and (n[0] == 0).all() and (n[:, -1] == 0).all()


@pytest.mark.parametrize("case", list(SMALL_CASES))
def test_season_small_golden(case):
    g = load("season_small.npz")
    forcing = {k: g[k] for k in ("precip", "conc", "wind", "drift")}
    fl = O.Flags(**SMALL_CASES[case])
    rho = g["clim__rho"] if case == "clim" else None
    out = O.run_season(forcing, g["ic"], g["mask"], int(g["dx"]), P(g[case + "__params"]), fl, rho_clim=rho)
    for name in OUT_NAMES:
        assert_identical(out[name], g[case + "__" + name], case + ":" + name)
    if case == "nodyn":            # Appendix A: land stays 0.0 (not NaN) in snowAdv/snowDiv without dynamics
        assert not np.isnan(out["snowAdv"]).any()
    assert (out["density"][0] == 0).all() and (out["snowAcc"][0] == 0).all()


@pytest.mark.parametrize("case", ["oneseason", "multiseason"])
def test_season_100km_digest(case):
    g = load("season_100km_digest.npz")
    mask = S.region_mask(dx=100000)
    T, seed = int(g["T"]), int(g["seed"])
    F = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed)
    for k in ("precip", "conc", "wind", "drift"):
        if canon_sha(F[k]) != str(g["in_sha__" + k]):
            pytest.skip("synthetic generator output differs from the fixture (numpy/scipy version); "
                        "season_small.npz carries its own inputs")
    out = O.run_season(F, ic, mask, 100000, P(g[case + "__params"]), O.Flags(atmlossInc=int(case == "multiseason")))
    for name in OUT_NAMES:
        assert_identical(out[name][-1], g[case + "__last__" + name], name + "[-1]")
        assert_identical(out[name][60], g[case + "__day60__" + name], name + "[60]")
        assert canon_sha(out[name]) == str(g[case + "__sha__" + name]), name

"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): NaN/land/ice masks bit-exact (identical ``isnan`` patterns) and every value
within 1e-10 relative -- the tests below assert the stronger property the design aims for: value-identical
(``np.array_equal(..., equal_nan=True)``) on every array, and additionally state the 1e-10 tolerance check.
"""
import numpy as np
import pytest

from nesosim_b200 import synthetic as S
from oracle import nesosim_oracle as O
from oracle.astropy_restated import convolve_fill0, gaussian2d_kernel

pytestmark = pytest.mark.gpu

RTOL = 1e-10   # tolerance stated by north_star for snow volume / depth / density / both layers
ATOL = 1e-300


def assert_parity(got, ref, name):
    assert got.shape == ref.shape, name
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN mask differs: " + name
    fin = ~np.isnan(ref)
    assert np.allclose(got[fin], ref[fin], rtol=RTOL, atol=ATOL), "outside 1e-10: " + name
    assert np.array_equal(got, ref, equal_nan=True), "not value-identical: " + name


def oracle_params(row, **kw):
    return O.Params(windPackFactor=row[0], windPackThresh=row[1], leadLossFactor=row[2], atmLossFactor=row[3], **kw)


PATHS = ["general", "ensemble"]


def run_both(mask, T, dx, params, flags, seed=0, rho_clim=None, ic_scale=1.0, conv_variant="post_divide",
             path="auto", expect_path=None):
    from nesosim_b200.engine import SnowBudgetEngine
    forcing = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed) * ic_scale
    params = np.asarray(params, dtype=float).reshape(-1, 4)
    eng = SnowBudgetEngine(mask, T, dx, n_members=len(params), conv_variant=conv_variant, **flags)
    eng.set_path(path)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"], rho_clim)
    out = eng.run_season(params, ic)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    assert eng.last_path() == (expect_path or (path if path != "auto" else eng.last_path()))
    eng.close()
    refs = []
    fl = O.Flags(dynamicsInc=flags.get("dynamicsInc", 1), leadlossInc=flags.get("leadlossInc", 1),
                 windpackInc=flags.get("windpackInc", 1), atmlossInc=flags.get("atmlossInc", 0),
                 densityType=flags.get("densityType", "variable"), conv_variant=conv_variant)
    for row in params:
        refs.append(O.run_season(forcing, ic, mask, dx, oracle_params(row), fl, rho_clim=rho_clim))
    return got, refs


def compare_all(got, refs):
    for m, ref in enumerate(refs):
        for name in got:
            assert_parity(got[name][m], ref[name], "%s[member %d]" % (name, m))


ONESEASON = [5.8e-7, 5., 2.9e-7, 2.2e-8]      # run_oneseason.py:40-48
MULTISEASON = [5.8e-7, 5., 1.45e-7, 2.2e-8]   # run_multiseason.py:42-50


@pytest.mark.parametrize("path", PATHS)
def test_season_100km_oneseason_params(cuda, path):
    mask = S.region_mask(dx=100000)
    got, refs = run_both(mask, 62, 100000, [ONESEASON], dict(atmlossInc=0), seed=11, path=path)
    compare_all(got, refs)
    h = got["snowDepths"][0]
    assert np.isfinite(h).sum() > 1000 and np.nanmax(h) > 0.01      # the comparison is not vacuous


@pytest.mark.parametrize("path", PATHS)
def test_season_100km_multiseason_params_ensemble(cuda, path):
    mask = S.region_mask(dx=100000)
    params = np.vstack([MULTISEASON, S.ensemble_params(4, seed=5)])
    got, refs = run_both(mask, 40, 100000, params, dict(atmlossInc=1), seed=12, path=path)
    compare_all(got, refs)
    # members really differ
    assert not np.array_equal(got["snowDepths"][1], got["snowDepths"][2], equal_nan=True)


@pytest.mark.parametrize("case", ["oneseason", "multiseason"])
@pytest.mark.parametrize("path", PATHS)
def test_full_season_against_reference_digest(cuda, path, case):
    """A full Aug 15 - May 1 season (260 days) on the 100 km grid against the digest the REFERENCE's own calcBudget
    loop produced for it (tests/golden/season_100km_digest.npz, made by tests/golden/make_golden.py)."""
    from golden_util import OUT_NAMES, assert_identical, canon_sha, load
    from nesosim_b200.engine import SnowBudgetEngine
    g = load("season_100km_digest.npz")
    mask = S.region_mask(dx=100000)
    T, seed = int(g["T"]), int(g["seed"])
    F = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed)
    for k in ("precip", "conc", "wind", "drift"):
        if canon_sha(F[k]) != str(g["in_sha__" + k]):
            pytest.skip("synthetic generator output differs from the fixture (numpy/scipy version)")
    eng = SnowBudgetEngine(mask, T, 100000, n_members=1, atmlossInc=int(case == "multiseason"))
    eng.set_path(path)
    eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    out = {k: v[0].cpu().numpy() for k, v in eng.run_season([g[case + "__params"]], ic).items()}
    assert eng.last_path() == path and eng.rerun_count() == 0
    for name in OUT_NAMES:
        assert_identical(out[name][-1], g[case + "__last__" + name], name + "[-1]")
        assert_identical(out[name][60], g[case + "__day60__" + name], name + "[60]")
        assert canon_sha(out[name]) == str(g[case + "__sha__" + name]), name


def test_full_size_ensemble_properties(cuda):
    """BASELINE's per-GPU shard at full size (128 members x 260 days x 90x90; 25.9 GB of outputs): too large for the
    CPU oracle, so size-independent properties -- determinism, member-independence of the accumulation planes,
    identical members for identical parameters, land mask on every day, and agreement of a sample of members with
    the general kernels (themselves checked against the oracle above)."""
    from nesosim_b200.engine import SnowBudgetEngine
    torch = cuda
    mask = S.region_mask(dx=100000)
    T, M = 260, 128
    F = S.make_season(mask, T, seed=2024)
    ic = S.make_ic(mask, seed=2024)
    params = S.ensemble_params(M, seed=2024)
    params[77] = params[3]                                   # two members with the same parameter set
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    out = eng.run_season(params, ic)
    assert eng.rerun_count() == 0

    def same(a, b):
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())

    sums = {k: torch.nan_to_num(v, nan=3.0).sum(dtype=torch.float64).item() for k, v in out.items()}
    out2 = eng.run_season(params, ic)                        # determinism: a second run is identical
    for k in out:
        assert torch.nan_to_num(out2[k], nan=3.0).sum(dtype=torch.float64).item() == sums[k], k
    assert same(out["snowDepths"], out2["snowDepths"])
    del out2
    for k in ("snowAcc", "snowOcean"):                        # member-independent planes
        assert same(out[k], out[k][:1].expand_as(out[k])), k
    for k in out:                                             # identical parameters, identical members
        assert same(out[k][77], out[k][3]), k
    assert not same(out["snowDepths"][5], out["snowDepths"][6])
    land = torch.from_numpy((mask > 10) | (mask < 1)).cuda()
    h = out["snowDepths"]
    assert bool(torch.isnan(h[:, 1:, :, land]).all())          # land is NaN from slot 1 on
    ho = h[:, :, :, ~land]
    assert bool(((ho >= 0) | torch.isnan(ho)).all())           # fill_nan_no_negative: no negative depth survives
    assert float(torch.isnan(ho).double().mean()) < 0.05       # (NaN only where the forcing itself is NaN)
    dens = out["density"][:, 1:]
    ok = torch.isnan(dens) | ((dens >= 200.0) & (dens <= 350.0))
    assert bool(ok.all())
    # a sample of members through the general kernels
    pick = [0, 3, 64, 127]
    eng2 = SnowBudgetEngine(mask, T, 100000, n_members=len(pick), atmlossInc=1)
    eng2.set_path("general")
    eng2.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    ref = eng2.run_season(params[pick], ic)
    for k in out:
        for i, m in enumerate(pick):
            assert same(out[k][m], ref[k][i]), (k, m)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("flags", [dict(dynamicsInc=0), dict(leadlossInc=0, atmlossInc=1), dict(windpackInc=0),
                                   dict(dynamicsInc=0, leadlossInc=0, windpackInc=0, atmlossInc=0)])
def test_switches(cuda, flags, path):
    mask = S.region_mask(dx=100000)
    got, refs = run_both(mask, 12, 100000, [MULTISEASON], flags, seed=13, path=path)
    compare_all(got, refs)


def test_clim_density(cuda):
    mask = S.region_mask(dx=100000)
    T = 15
    rho = 1000 * (0.29 + 0.0003 * np.arange(T))      # shape of W99_density.csv values (utils.py:1336-1343)
    got, refs = run_both(mask, T, 100000, [ONESEASON], dict(densityType="clim"), seed=14, rho_clim=rho,
                         expect_path="general")     # the season-resident kernel is variable-density only
    compare_all(got, refs)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(45, 70), (9, 96), (94, 33)])
def test_ragged_grid_not_multiple_of_tile(cuda, path, shape):
    mask = S.region_mask(shape=shape, kind="disc")
    got, refs = run_both(mask, 10, 50000, [MULTISEASON, ONESEASON], dict(atmlossInc=1), seed=15, ic_scale=3.0,
                         path=path)
    compare_all(got, refs)


def test_odd_cell_count_falls_back_to_general_kernel(cuda):
    """95x33 has an odd number of cells: planes are not 16-byte aligned, so the bulk-store kernel steps aside."""
    from nesosim_b200 import _lib
    mask = S.region_mask(shape=(95, 33), kind="disc")
    got, refs = run_both(mask, 6, 50000, [MULTISEASON], dict(atmlossInc=1), seed=15, expect_path="general")
    compare_all(got, refs)
    with pytest.raises(_lib.NesosimError, match="not applicable"):
        run_both(mask, 6, 50000, [MULTISEASON], dict(atmlossInc=1), seed=15, path="ensemble")


def test_minimum_grid_2x2(cuda):
    mask = np.full((2, 2), 8, dtype=np.uint8)
    got, refs = run_both(mask, 5, 100000, [MULTISEASON], dict(atmlossInc=1), seed=16, expect_path="general")
    compare_all(got, refs)


@pytest.mark.parametrize("path", PATHS)
def test_prenormalised_kernel_variant(cuda, path):
    mask = S.region_mask(dx=100000)
    got, refs = run_both(mask, 8, 100000, [ONESEASON], dict(), seed=17, conv_variant="pre_normalised", path=path)
    compare_all(got, refs)


def _blocky_mask(ny=100, nx=150):
    """Large land blocks (whole 32x16 tiles), a lake-only tile, a coast running through tiles, ocean elsewhere."""
    mask = np.full((ny, nx), 8, dtype=np.uint8)
    mask[:40, :70] = 11          # land: tiles (0..1, 0..1) entirely, and partly the ones around them
    mask[64:80, 96:128] = 0      # one tile of lake cells exactly
    mask[50:, 130:] = 12         # coast code, ragged last tile column
    mask[90:, :] = 11
    return mask


@pytest.mark.parametrize("threads", ["256", "512"])
@pytest.mark.parametrize("flags", [dict(atmlossInc=1), dict(dynamicsInc=0), dict(leadlossInc=0, windpackInc=0),
                                   dict(dynamicsInc=0, leadlossInc=0, windpackInc=0, atmlossInc=0),
                                   dict(densityType="clim")])
def test_land_tile_shortcut_and_thread_variants(cuda, flags, threads, monkeypatch):
    """Tiles without an ocean cell take the closed-form land path from the second step on (day_step_land_tile); the
    IC deliberately puts snow on land, so step 0 must still be the general arithmetic.  Both CTA shapes."""
    monkeypatch.setenv("NESOSIM_DAY_THREADS", threads)
    mask = _blocky_mask()
    T = 7
    rho = 1000 * (0.29 + 0.0003 * np.arange(T)) if flags.get("densityType") == "clim" else None
    from nesosim_b200.engine import SnowBudgetEngine
    forcing = S.make_season(mask, T, seed=21)
    ic = np.full(mask.shape, 0.08)               # finite depth everywhere at slot 0, land included
    eng = SnowBudgetEngine(mask, T, 50000, n_members=2, **flags)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"], rho)
    got = {k: v.cpu().numpy() for k, v in eng.run_season([MULTISEASON, ONESEASON], ic).items()}
    fl = O.Flags(dynamicsInc=flags.get("dynamicsInc", 1), leadlossInc=flags.get("leadlossInc", 1),
                 windpackInc=flags.get("windpackInc", 1), atmlossInc=flags.get("atmlossInc", 0),
                 densityType=flags.get("densityType", "variable"))
    refs = [O.run_season(forcing, ic, mask, 50000, oracle_params(row), fl, rho_clim=rho) for row in (MULTISEASON, ONESEASON)]
    compare_all(got, refs)
    # resumed in pieces: the first step of every piece reads a caller-provided slot and must not take the shortcut
    out = eng.alloc_outputs(zero=True)
    for a, b in ((0, 1), (1, 3), (4, 2)):
        eng.run_season([MULTISEASON, ONESEASON], ic, out, first_step=a, num_steps=b)
    for k in got:
        assert np.array_equal(out[k].cpu().numpy(), got[k], equal_nan=True), k
    eng.close()


def test_land_shortcut_is_not_taken_on_a_caller_provided_slot(cuda):
    """nesosim_run_season(first_step=k) on a state the caller made up: finite depths on land at slot k are data."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = _blocky_mask()
    T = 4
    forcing = S.make_season(mask, T, seed=22)
    eng = SnowBudgetEngine(mask, T, 50000, n_members=1, atmlossInc=1)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = eng.alloc_outputs(zero=True)
    rng = np.random.default_rng(5)
    state = O.gen_empty_arrays(T, *mask.shape)
    for k in out:
        made_up = rng.random(state[k][1].shape) * 0.2
        state[k][1] = made_up
        out[k][0, 1] = cuda.from_numpy(made_up).cuda()
    eng.run_season([MULTISEASON], None, out, first_step=1, num_steps=2)
    p = oracle_params(MULTISEASON)
    for x in (1, 2):
        O.calc_budget(state, forcing["conc"][x], forcing["precip"][x], forcing["drift"][x], forcing["wind"][x],
                      np.full(mask.shape, np.nan), mask, 50000, x, p, O.Flags(atmlossInc=1))
    for k in out:
        assert np.array_equal(out[k][0, 2:].cpu().numpy(), state[k][2:], equal_nan=True), k
    eng.close()


def test_25km_short_season(cuda):
    mask = S.region_mask(dx=25000)
    got, refs = run_both(mask, 6, 25000, [MULTISEASON], dict(atmlossInc=1), seed=18, expect_path="general")
    compare_all(got, refs)


@pytest.mark.parametrize("cluster", ["4", "5", "6", "7", "8"])
def test_ensemble_kernel_cluster_sizes_many_members_per_cluster(cuda, cluster, monkeypatch):
    """Every cluster size of the season-resident kernel; 3 clusters share 11 members with per-member ICs."""
    from nesosim_b200.engine import SnowBudgetEngine
    variant = "cluster" + cluster
    monkeypatch.setenv("NESOSIM_ENS_CLUSTERS", "3")
    monkeypatch.setenv("NESOSIM_ENS_CLUSTER", cluster)
    # the full 100 km grid needs >= 5 CTAs per member (shared memory); a 48x90 cut of it also fits 4
    mask = S.region_mask(dx=100000) if cluster != "4" else np.ascontiguousarray(S.region_mask(dx=100000)[20:68])
    T, M = 9, 11
    forcing = S.make_season(mask, T, seed=23)
    rng = np.random.default_rng(23)
    ic = S.make_ic(mask, seed=23)[None] * rng.uniform(0.5, 3.0, (M, 1, 1))
    params = S.ensemble_params(M, seed=23)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
    for m in range(M):
        ref = O.run_season(forcing, ic[m], mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m], ref[name], "%s[%d] %s" % (name, m, variant))


@pytest.mark.parametrize("variant", ["t608r2o1", "t480r2o1", "t480r2o2", "t352r3o2", "t736r2o1", "t224r4o3", "t352r6o5"])
def test_ensemble_kernel_build_variants(cuda, variant, monkeypatch):
    """Every build variant (threads x owned cells per thread) of the season-resident kernel on the 100 km grid."""
    from nesosim_b200.engine import SnowBudgetEngine
    monkeypatch.setenv("NESOSIM_ENS_VARIANT", variant)
    monkeypatch.setenv("NESOSIM_ENS_CLUSTERS", "2")
    mask = S.region_mask(dx=100000)
    T, M = 8, 5
    forcing = S.make_season(mask, T, seed=29)
    ic = S.make_ic(mask, seed=29)
    params = S.ensemble_params(M, seed=29)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
    for m in range(M):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m], ref[name], "%s[%d] %s" % (name, m, variant))


def test_ensemble_kernel_hands_out_of_range_operands_to_the_general_kernels(cuda):
    """Depths around 1e-300 leave the proven range of the fast constant divisions: the season-resident kernel
    must notice, and the season must still come out value-identical (redone by the per-day kernels)."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 6
    forcing = S.make_season(mask, T, seed=31)
    forcing["precip"] = forcing["precip"] * 1e-300          # snowfall and depths of order 1e-300 m
    ic = S.make_ic(mask, seed=31) * 1e-298
    params = S.ensemble_params(3, seed=31)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=3, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
    assert eng.rerun_count() == 1
    for m in range(3):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m], ref[name], "%s[%d]" % (name, m))


def test_ensemble_kernel_with_drift_defined_over_land(cuda):
    """Real drift products are NaN over land, which lets the season kernel compute the land entries of its raw lists
    on day 0 only; a forcing with finite drift everywhere must switch that shortcut off (pre-pass flag)."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, M = 7, 3
    forcing = S.make_season(mask, T, seed=37)
    rng = np.random.default_rng(37)
    d = forcing["drift"]
    forcing["drift"] = np.where(np.isnan(d), 0.05 * rng.standard_normal(d.shape), d)
    forcing["drift"][3] = np.nan                       # one missing-drift day in between
    ic = S.make_ic(mask, seed=37) + 0.01               # snow on land in the initial condition as well
    params = S.ensemble_params(M, seed=37)
    for path in PATHS:
        eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
        eng.set_path(path)
        eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
        out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
        assert eng.rerun_count() == 0
        for m in range(M):
            ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
            for name in out:
                assert_parity(out[name][m], ref[name], "%s[%d] %s" % (name, m, path))
        eng.close()


@pytest.mark.parametrize("path", PATHS)
def test_batch_of_seasons_with_their_own_forcing_and_length(cuda, path):
    """run_multiseason.py as one native call: three seasons of different length stacked in one context, five members
    spread over them; every member must equal the oracle run on its own season."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, days = 12, [12, 9, 11]
    seasons = [S.make_season(mask, T, seed=50 + i) for i in range(3)]
    stack = {k: np.stack([f[k] for f in seasons]) for k in ("precip", "conc", "wind", "drift")}
    member_set = [0, 1, 2, 1, 0]
    M = len(member_set)
    params = S.ensemble_params(M, seed=51)
    ic = S.make_ic(mask, seed=52)[None] * np.linspace(0.5, 2.0, M)[:, None, None]
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path(path)
    eng.set_forcing_sets(stack["precip"], stack["conc"], stack["wind"], stack["drift"], member_set, days)
    out = eng.alloc_outputs()
    for t in out.values():
        t.fill_(-5.0)                                  # slots past a member's last day must stay untouched
    out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic, out).items()}
    assert eng.last_path() == path
    for m in range(M):
        s_ = member_set[m]
        d = days[s_]
        f = {k: v[:d] for k, v in seasons[s_].items()}
        ref = O.run_season(f, ic[m], mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m][:d], ref[name], "%s[%d] %s" % (name, m, path))
            assert (out[name][m][d:] == -5.0).all(), name
    eng.set_forcing(seasons[0]["precip"], seasons[0]["conc"], seasons[0]["wind"], seasons[0]["drift"])   # back to one season
    again = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
    ref = O.run_season(seasons[0], ic[1], mask, 100000, oracle_params(params[1]), O.Flags(atmlossInc=1))
    assert_parity(again["snowDepths"][1], ref["snowDepths"], "single season after a batch")


def test_step_day_matches_calc_budget(cuda):
    """nesosim_step_day has calcBudget's in-place contract (NESOSIM.py:224-347)."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 6
    forcing = S.make_season(mask, T, seed=19)
    ic = S.make_ic(mask, seed=19)
    p = oracle_params(MULTISEASON)
    fl = O.Flags(atmlossInc=1)
    ref = O.run_season(forcing, ic, mask, 100000, p, fl)
    eng = SnowBudgetEngine(mask, T, 100000, atmlossInc=1)
    out = eng.alloc_outputs(zero=True)
    half = O.initial_depths(ic, forcing["conc"][0], p)
    out["snowDepths"][0, 0, 0] = cuda.from_numpy(half).cuda()
    out["snowDepths"][0, 0, 1] = cuda.from_numpy(half).cuda()
    for x in range(T - 1):
        eng.step_day(x, forcing["conc"][x], forcing["precip"][x], forcing["drift"][x], forcing["wind"][x],
                     [MULTISEASON], out)
    for name, t in out.items():
        assert_parity(t[0].cpu().numpy(), ref[name], name)


def test_resume_in_two_halves(cuda):
    """first_step/num_steps: running [0,k) then [k,T-1) equals one run."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 14
    forcing = S.make_season(mask, T, seed=20)
    ic = S.make_ic(mask, seed=20)
    eng = SnowBudgetEngine(mask, T, 100000)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    a = eng.run_season([ONESEASON], ic)
    b = eng.alloc_outputs()
    eng.run_season([ONESEASON], ic, b, 0, 6)
    eng.run_season([ONESEASON], ic, b, 6, -1)
    for n in a:
        assert cuda.equal(a[n].nan_to_num(nan=-7.0), b[n].nan_to_num(nan=-7.0)), n


def test_outputs_subset_uses_internal_state(cuda):
    """NULL outputs are carried in internal scratch; the requested ones are unchanged."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 9
    forcing = S.make_season(mask, T, seed=21)
    ic = S.make_ic(mask, seed=21)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=2, atmlossInc=1)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    params = S.ensemble_params(2, seed=2)
    full = eng.run_season(params, ic)
    part = eng.alloc_outputs(names=("snowDepths", "density"))
    eng.run_season(params, ic, part)
    for n in part:
        assert cuda.equal(full[n].nan_to_num(nan=-7.0), part[n].nan_to_num(nan=-7.0)), n


def test_outputs_subset_on_the_general_path_with_mixed_member_strides(cuda):
    """Some plane arrays from the caller (member stride T*ny*nx), the rest in the internal scratch (stride ny*nx): the
    day kernel carries one stride per array kind, so such a step is launched member by member."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(shape=(40, 70), kind="disc")
    T = 6
    forcing = S.make_season(mask, T, seed=31)
    ic = S.make_ic(mask, seed=31)
    eng = SnowBudgetEngine(mask, T, 50000, n_members=3, atmlossInc=1)
    eng.set_path("general")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    params = S.ensemble_params(3, seed=4)
    full = eng.run_season(params, ic)
    for names in (("snowDepths", "density"), ("snowLead", "snowAcc", "density"), ("snowDepths", "snowAdv")):
        part = eng.alloc_outputs(names=names)
        eng.run_season(params, ic, part)
        for n in part:
            assert cuda.equal(full[n].nan_to_num(nan=-7.0), part[n].nan_to_num(nan=-7.0)), (n, names)
    eng.close()


def test_host_path_end_to_end(cuda, monkeypatch):
    from nesosim_b200.engine import SnowBudgetEngine
    monkeypatch.setenv("NESOSIM_HOST_COMPACT", "0")     # the plain drain: every byte of the contract crosses the link
    mask = S.region_mask(dx=100000)
    T = 10
    forcing = S.make_season(mask, T, seed=22)
    ic = S.make_ic(mask, seed=22)
    params = S.ensemble_params(5, seed=3)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=5, atmlossInc=1)
    import os
    os.environ["NESOSIM_HOST_BATCH_GB"] = "0.02"       # force several member batches
    try:
        out, up, down = eng.run_season_host(forcing, params, ic)
    finally:
        del os.environ["NESOSIM_HOST_BATCH_GB"]
    # snowAcc / snowOcean are member-independent: one copy crosses the link, the host replicates it
    assert up >= 5 * T * mask.size * 8 and down == (10 * 5 + 2) * T * mask.size * 8
    for m in range(5):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m], ref[name], name)


def test_host_path_regrows_its_staging_when_more_outputs_are_asked_for(cuda, monkeypatch):
    """One engine, first a season with only snowDepths requested, then with all eleven arrays and the same member
    batch: the cached device staging is sized in bytes, so the second call must regrow it (it used to be cached by
    the batch count and the second season wrote up to 6x past the allocation)."""
    from nesosim_b200.engine import SnowBudgetEngine
    monkeypatch.setenv("NESOSIM_HOST_COMPACT", "0")
    mask = S.region_mask(dx=100000)
    T, M = 9, 3
    forcing = S.make_season(mask, T, seed=31)
    ic = S.make_ic(mask, seed=31)
    params = S.ensemble_params(M, seed=31)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    few, _, down_few = eng.run_season_host(forcing, params, ic, names=("snowDepths",))
    full, _, down_full = eng.run_season_host(forcing, params, ic)
    again, _, _ = eng.run_season_host(forcing, params, ic, names=("snowDepths", "density"))
    assert down_few == 2 * M * T * mask.size * 8 and down_full > 5 * down_few
    for m in range(M):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        assert_parity(few["snowDepths"][m], ref["snowDepths"], "snowDepths (first call)")
        for name in full:
            assert_parity(full[name][m], ref[name], name + " (second call)")
        for name in again:
            assert_parity(again[name][m], ref[name], name + " (third call)")
    eng.close()


def _host_season(eng, forcing, params, ic, monkeypatch, compact, names=None, **env):
    monkeypatch.setenv("NESOSIM_HOST_COMPACT", "1" if compact else "0")
    for k in ("NESOSIM_HOST_BATCH_GB", "NESOSIM_HOST_CHUNK_MB", "NESOSIM_DRAIN_HEAD", "NESOSIM_HOST_THREADS", "NESOSIM_HOST_HYBRID",
              "NESOSIM_HOST_RING"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    kw = {} if names is None else {"names": names}
    # poisoned destination: every cell the drain does not write would show
    out = None
    if names is None:
        from nesosim_b200 import _lib
        names = _lib.OUTPUT_NAMES
    out = {n: np.full((eng.M, eng.T, 2, eng.ny, eng.nx) if n == "snowDepths" else (eng.M, eng.T, eng.ny, eng.nx), -3.25)
           for n in names}
    res, up, down = eng.run_season_host(forcing, params, ic, outputs=out, **kw)
    return res, down, eng.host_drain_info()


@pytest.mark.parametrize("path", ["ensemble", "general"])
def test_host_path_compacted_drain_matches_the_plain_one_bit_for_bit(cuda, monkeypatch, path):
    """nesosim_run_season_host with the compacted drain (ocean cells + the land cells of the first three slots cross the
    link, host threads scatter) against the plain copy of every array: same bits in the caller's arrays -- NaNs
    included -- through one batch and one chunk, through several batches, and through more chunks than the pinned ring
    has slots; and well under half of the bytes on the link."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, M = 12, 7
    forcing = S.make_season(mask, T, seed=41)
    ic = S.make_ic(mask, seed=41)
    ic[:3] = 0.07                                   # snow on land in the initial condition: slot 1 differs from slot 2 there
    params = S.ensemble_params(M, seed=41)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path(path)
    plain, down_plain, info = _host_season(eng, forcing, params, ic, monkeypatch, False)
    assert info[0] is False
    ref0 = O.run_season(forcing, ic, mask, 100000, oracle_params(params[0]), O.Flags(atmlossInc=1))
    for name in plain:
        assert_parity(plain[name][0], ref0[name], name)
    seen_plain = 0
    for env in ({"NESOSIM_HOST_HYBRID": "0"}, {"NESOSIM_HOST_BATCH_GB": "0.03", "NESOSIM_HOST_HYBRID": "0"},
                {"NESOSIM_HOST_CHUNK_MB": "1.5"}, {"NESOSIM_HOST_CHUNK_MB": "0.1", "NESOSIM_HOST_THREADS": "3"},
                {"NESOSIM_HOST_BATCH_GB": "0.05", "NESOSIM_HOST_CHUNK_MB": "0.1", "NESOSIM_HOST_THREADS": "1", "NESOSIM_HOST_RING": "2"},
                {"NESOSIM_HOST_BATCH_GB": "0.03", "NESOSIM_HOST_CHUNK_MB": "0.1", "NESOSIM_HOST_THREADS": "1", "NESOSIM_HOST_RING": "2"}):
        packed, down, info = _host_season(eng, forcing, params, ic, monkeypatch, True, **env)
        assert info == (True, 0), (env, info)
        n_packed, n_plain = eng.host_drain_blocks()
        assert n_packed + n_plain == 9 * M, (env, n_packed, n_plain)
        if env.get("NESOSIM_HOST_HYBRID") == "0":       # every block packed: well under half of the bytes on the link
            assert n_plain == 0 and down < 0.62 * down_plain, (env, down, down_plain, n_plain)
        seen_plain += n_plain
        for name in plain:
            assert np.array_equal(packed[name].view(np.uint64), plain[name].view(np.uint64)), (name, env)
    # with one slow host thread and a ring of two small slots the link must have taken whole blocks at some point
    assert seen_plain > 0
    eng.close()


def test_host_path_compacted_drain_falls_back_when_land_cells_change(cuda, monkeypatch):
    """The packed form repeats a land cell's value of the last head slot; that land cells are constant from there on is
    checked on the device, chunk by chunk.  With a single head slot (test knob) slot 1 differs from slot 0 on land, every
    chunk is flagged and copied in full: the arrays are the same again, and the call says how many chunks it gave up."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, M = 8, 5
    forcing = S.make_season(mask, T, seed=43)
    ic = S.make_ic(mask, seed=43)
    params = S.ensemble_params(M, seed=43)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M)
    plain, _, _ = _host_season(eng, forcing, params, ic, monkeypatch, False)
    packed, _, info = _host_season(eng, forcing, params, ic, monkeypatch, True, NESOSIM_DRAIN_HEAD="1", NESOSIM_HOST_CHUNK_MB="0.5")
    assert info[0] is True and info[1] >= 2, info
    for name in plain:
        assert np.array_equal(packed[name].view(np.uint64), plain[name].view(np.uint64)), name
    eng.close()


def test_host_path_compacted_drain_of_selected_arrays(cuda, monkeypatch):
    """Only some arrays requested (with and without the member-independent snowAcc), one member (nothing to share)."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 9
    forcing = S.make_season(mask, T, seed=47)
    ic = S.make_ic(mask, seed=47)
    for M in (1, 4):
        params = S.ensemble_params(M, seed=47)
        eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
        for names in (("snowDepths",), ("density", "snowAcc", "snowLead"), ("snowOcean",)):
            plain, _, _ = _host_season(eng, forcing, params, ic, monkeypatch, False, names=names)
            packed, _, info = _host_season(eng, forcing, params, ic, monkeypatch, True, names=names)
            assert info[0] is (names != ("snowOcean",)), (names, info)
            for name in names:
                assert np.array_equal(packed[name].view(np.uint64), plain[name].view(np.uint64)), (name, names, M)
        eng.close()


@pytest.mark.parametrize("cluster,rows,land_cols", [("5", "18,36,54,72", 3), ("6", "16,30,44,60,76", 55),
                                                    ("6", "14,30,46,62,78", 55)])
def test_ensemble_kernel_all_land_rows_at_a_strip_boundary(cuda, cluster, rows, land_cols, monkeypatch):
    """A strip whose two rows facing a neighbour hold no ocean cell never pushes to that neighbour, and the neighbour,
    expecting no bytes from it, may run a day ahead.  The "done reading" handshake is therefore kept PER NEIGHBOUR (a
    shared barrier completed on two arrivals of the fast neighbour and let this strip overwrite the other neighbour's
    halo while it was being read), and the next member's slot-0 pushes wait for a cluster barrier.  Masks with all-land
    row bands straddling forced strip boundaries, several members per cluster, per-member ICs with snow on land."""
    from nesosim_b200.engine import SnowBudgetEngine
    monkeypatch.setenv("NESOSIM_ENS_CLUSTERS", "2")
    monkeypatch.setenv("NESOSIM_ENS_CLUSTER", cluster)
    monkeypatch.setenv("NESOSIM_ENS_ROWS", rows)
    ny, nx = 90, 90
    mask = np.full((ny, nx), 8, dtype=np.uint8)
    cuts = [int(r) for r in rows.split(",")]
    mask[cuts[0] - 2:cuts[0] + 2] = 11            # land on both sides of the first boundary
    mask[cuts[1]:cuts[1] + 2] = 11                # only the lower strip's top rows are land
    mask[cuts[2] - 2:cuts[2]] = 11                # only the upper strip's bottom rows are land
    mask[:, :land_cols] = 11                      # (55 land columns: the strips' lists fit the default build variant)
    mask[40:43, 60:70] = 0
    T, M = 12, 7
    forcing = S.make_season(mask, T, seed=37)
    rng = np.random.default_rng(37)
    ic = (0.05 + S.make_ic(mask, seed=37))[None] * rng.uniform(0.5, 3.0, (M, 1, 1))      # snow on land in slot 0 too
    params = S.ensemble_params(M, seed=37)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    for rep in range(3):                          # a race does not show on every run
        out = {k: v.cpu().numpy() for k, v in eng.run_season(params, ic).items()}
        for m in range(M):
            ref = O.run_season(forcing, ic[m], mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
            for name in out:
                assert_parity(out[name][m], ref[name], "%s[%d] rep %d" % (name, m, rep))
    assert eng.rerun_count() == 0
    eng.close()


def test_tall_narrow_grid_is_not_taken_by_the_season_kernel(cuda):
    """ny > 511 does not fit the 16-bit row*128+col cell codes of the season-resident kernel: general path."""
    mask = np.full((520, 4), 8, dtype=np.uint8)
    mask[::7, 0] = 11
    got, refs = run_both(mask, 4, 100000, [MULTISEASON], dict(atmlossInc=1), seed=41, path="auto", expect_path="general")
    compare_all(got, refs)


def test_asynchronous_mode_back_to_back_seasons(cuda):
    """nesosim_set_async: run_season never synchronises; three seasons with different parameters are enqueued back to
    back into three output sets, nesosim_sync then resolves their operand-range flags (none raised)."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, M = 20, 4
    forcing = S.make_season(mask, T, seed=61)
    ic = S.make_ic(mask, seed=61)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path("ensemble")
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    eng.set_async(True)
    ic_dev = eng._dev(ic)
    sets = [S.ensemble_params(M, seed=61 + i) for i in range(3)]
    outs = [eng.run_season(p, ic_dev, eng.alloc_outputs()) for p in sets]
    assert eng.sync() == 0
    k0 = eng.season_kernel_time()
    assert k0[1] == 3 and k0[0] > 0
    for p, out in zip(sets, outs):
        for m in range(M):
            ref = O.run_season(forcing, ic, mask, 100000, oracle_params(p[m]), O.Flags(atmlossInc=1))
            for name in out:
                assert_parity(out[name][m].cpu().numpy(), ref[name], name)
    # the host-buffer call suspends asynchronous mode for its duration (it copies the results out itself)
    host, _, _ = eng.run_season_host(forcing, sets[0], ic, names=("snowDepths", "density"))
    ref = O.run_season(forcing, ic, mask, 100000, oracle_params(sets[0][1]), O.Flags(atmlossInc=1))
    for name in host:
        assert_parity(host[name][1], ref[name], name + " (host call in asynchronous mode)")
    more = eng.run_season(sets[1], ic_dev, eng.alloc_outputs())          # ... and the mode is still on afterwards
    assert eng.sync() == 0
    assert_parity(more["density"][0].cpu().numpy(), O.run_season(forcing, ic, mask, 100000, oracle_params(sets[1][0]),
                                                                 O.Flags(atmlossInc=1))["density"], "density")
    eng.set_async(False)
    eng.close()


def test_asynchronous_mode_redoes_an_out_of_range_season_at_sync(cuda):
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T = 6
    forcing = S.make_season(mask, T, seed=62)
    ic = S.make_ic(mask, seed=62) * 1e-299
    eng = SnowBudgetEngine(mask, T, 100000, n_members=2, atmlossInc=1)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    eng.set_async(True)
    params = S.ensemble_params(2, seed=62)
    out = eng.run_season(params, eng._dev(ic))
    assert eng.sync() == 1 and eng.rerun_count() == 1
    for m in range(2):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        for name in out:
            assert_parity(out[name][m].cpu().numpy(), ref[name], name)
    eng.close()


# ------------------------------------------------------------------------------------ per-function KATs

def test_smooth_plain_and_nan_branch(cuda):
    from nesosim_b200 import engine as E
    rng = np.random.default_rng(0)
    g = gaussian2d_kernel()
    a = rng.standard_normal((37, 53))
    assert_parity(E.smooth(a), convolve_fill0(a, g), "plain")
    b = a.copy()
    b[5, 7] = np.nan                    # isolated NaN
    b[20:24, 30:34] = np.nan            # block >= 3x3 -> bot == 0 -> NaN preserved in the middle
    b[0, 0] = np.nan
    b[-1, -1] = np.nan                  # corners / edges
    b[0, 20] = np.nan
    ref = convolve_fill0(b, g)
    assert np.isnan(ref[21, 31]) and np.isfinite(ref[5, 7])
    assert_parity(E.smooth(b), ref, "interpolate")
    c = a.copy()
    c[3, 3] = np.inf
    c[9, 9] = -np.inf                   # +inf and -inf make arr.sum() NaN -> interpolate branch
    assert_parity(E.smooth(c), convolve_fill0(c, g), "inf pair")
    d = a.copy()
    d[3, 3] = np.inf                    # a single inf keeps the plain branch
    assert_parity(E.smooth(d), convolve_fill0(d, g), "single inf")
    assert_parity(E.smooth(b, "pre_normalised"), convolve_fill0(b, g, "pre_normalised"), "pre-normalised")


def test_drift_smoothing_tail_of_the_offline_regridders(cuda):
    """utils.int_smooth_drifts_v2/v3 (utils.py:283-291): sigma 0.5 and 1 kernels, NaNs present (interpolating branch),
    result masked where the gridded input was NaN."""
    from nesosim_b200 import engine as E
    rng = np.random.default_rng(4)
    gx = 0.1 * rng.standard_normal((41, 37))
    gy = 0.1 * rng.standard_normal((41, 37))
    gx[:6, :] = np.nan
    gy[:6, :] = np.nan                  # outside the convex hull of the source grid
    gx[20:24, 10:14] = np.nan
    gy[30, 30] = np.nan
    for sigma in (1, 0.5):
        got = E.smooth_gridded_drift(gx, gy, sigma_factor=sigma)
        k = gaussian2d_kernel(x_stddev=sigma, y_stddev=sigma, x_size=3, y_size=3)
        for i, comp in enumerate((gx, gy)):
            ref = np.ma.masked_where(np.isnan(comp), convolve_fill0(comp, k))
            assert np.array_equal(np.ma.getmaskarray(got[i]), np.ma.getmaskarray(ref))
            assert np.array_equal(got[i].compressed(), ref.compressed())


def test_ensemble_driver_reduces_on_the_device(cuda):
    """N3: per-member misfit against point observations, only snowDepths computed, nothing but M scalars leaves."""
    from nesosim_b200 import ensemble as ENS
    mask = S.region_mask(dx=100000)
    T, M = 10, 6
    forcing = S.make_season(mask, T, seed=61)
    ic = S.make_ic(mask, seed=61) * 3
    params = S.ensemble_params(M, seed=61)
    rng = np.random.default_rng(61)
    ocean = np.argwhere((mask <= 10) & (mask >= 1))
    pick = ocean[rng.choice(len(ocean), 200, replace=False)]
    day = rng.integers(1, T, 200)
    obs = (day, pick[:, 0], pick[:, 1], 0.2 * rng.random(200))
    mis, used = ENS.run_ensemble(mask, forcing, ic, params, 100000, obs, atmlossInc=1, fused=False)
    for m in range(M):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        with np.errstate(all="ignore"):
            model = (ref["snowDepths"][day, 0, pick[:, 0], pick[:, 1]] + ref["snowDepths"][day, 1, pick[:, 0], pick[:, 1]]) \
                / forcing["conc"][day, pick[:, 0], pick[:, 1]]
        d = model - obs[3]
        ok = np.isfinite(d)
        assert used[m] == ok.sum() and used[m] > 20
        assert np.isclose(mis[m], np.sum(d[ok] ** 2), rtol=1e-12)


@pytest.mark.parametrize("cluster", [None, "8"])
def test_misfit_fused_into_the_season_kernel(cuda, cluster, monkeypatch):
    """N3 fused: the season-resident kernel stores no output array at all; the thread that owns an observed cell writes
    the total depth of the observed days (8 bytes per observed cell-day), a one-CTA-per-member epilogue applies the
    observation operator (snow depth over ice at a day, row, col) and sums the squares.  Observations on land, on day 0, on the last
    day, repeated at one cell and day, in every strip; result equal to the oracle's to 1e-12 (the sums are formed in a
    different order), counts exact, and bit-identical from run to run."""
    from nesosim_b200.engine import SnowBudgetEngine
    if cluster:
        monkeypatch.setenv("NESOSIM_ENS_CLUSTER", cluster)
    mask = S.region_mask(dx=100000)
    T, M = 30, 9
    forcing = S.make_season(mask, T, seed=63)
    ic = S.make_ic(mask, seed=63) * 3
    params = S.ensemble_params(M, seed=63)
    rng = np.random.default_rng(63)
    n = 1500
    row, col = rng.integers(0, 90, n), rng.integers(0, 90, n)          # a good half of them on land
    day = rng.integers(0, T, n)
    day[:40] = 0
    day[40:80] = T - 1
    row[100:110], col[100:110], day[100:110] = row[100], col[100], day[100]      # ten observations of one cell and day
    depth = 0.3 * rng.random(n)
    obs = (day, row, col, depth)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    l0 = eng.launch_count()
    mis, used = eng.run_season_misfit(params, ic, obs)
    mis, used = mis.cpu().numpy(), used.cpu().numpy()
    assert eng.launch_count() - l0 == 4          # pre-pass (2), season kernel, the per-member epilogue: nothing else
    again = eng.run_season_misfit(params, ic, obs)[0].cpu().numpy()
    assert np.array_equal(mis, again)
    eng.close()
    for m in range(M):
        ref = O.run_season(forcing, ic, mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1))
        with np.errstate(all="ignore"):
            model = (ref["snowDepths"][day, 0, row, col] + ref["snowDepths"][day, 1, row, col]) / forcing["conc"][day, row, col]
        d = model - depth
        ok = np.isfinite(d)
        assert used[m] == ok.sum() and used[m] > 100
        assert np.isclose(mis[m], np.sum(d[ok] ** 2), rtol=1e-12, atol=0.0)
    # the driver takes the fused path by default and agrees with the HBM round trip
    from nesosim_b200 import ensemble as ENS
    a = ENS.run_ensemble(mask, forcing, ic, params, 100000, obs, atmlossInc=1, fused=True)
    b = ENS.run_ensemble(mask, forcing, ic, params, 100000, obs, atmlossInc=1, fused=False)
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=0.0) and np.array_equal(a[1], b[1])


def test_misfit_with_per_member_initial_conditions_and_replaced_observations(cuda):
    """Per-member ICs; no observation at all (zero misfit, zero count); a second set of observations replaces the
    first on the same context."""
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(dx=100000)
    T, M = 12, 5
    forcing = S.make_season(mask, T, seed=65)
    rng = np.random.default_rng(65)
    ic = S.make_ic(mask, seed=65)[None] * rng.uniform(0.5, 3.0, (M, 1, 1))
    params = S.ensemble_params(M, seed=65)
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    mis, used = eng.run_season_misfit(params, ic, ([], [], [], []))
    assert not mis.cpu().numpy().any() and not used.cpu().numpy().any()
    refs = [O.run_season(forcing, ic[m], mask, 100000, oracle_params(params[m]), O.Flags(atmlossInc=1)) for m in range(M)]
    for n in (40, 700):
        day, row, col = rng.integers(0, T, n), rng.integers(0, 90, n), rng.integers(0, 90, n)
        depth = 0.3 * rng.random(n)
        mis, used = eng.run_season_misfit(params, ic, (day, row, col, depth))
        mis, used = mis.cpu().numpy(), used.cpu().numpy()
        for m in range(M):
            with np.errstate(all="ignore"):
                model = (refs[m]["snowDepths"][day, 0, row, col] + refs[m]["snowDepths"][day, 1, row, col]) / forcing["conc"][day, row, col]
            d = model - depth
            ok = np.isfinite(d)
            assert used[m] == ok.sum()
            assert np.isclose(mis[m], np.sum(d[ok] ** 2), rtol=1e-12, atol=0.0)
    eng.close()


def test_misfit_mode_refuses_grids_the_season_kernel_does_not_take(cuda):
    from nesosim_b200 import _lib
    from nesosim_b200.engine import SnowBudgetEngine
    mask = S.region_mask(shape=(40, 130), kind="disc")
    forcing = S.make_season(mask, 4, seed=64)
    eng = SnowBudgetEngine(mask, 4, 50000, n_members=2)
    eng.set_forcing(forcing["precip"], forcing["conc"], forcing["wind"], forcing["drift"])
    with pytest.raises(_lib.NesosimError) as e:
        eng.run_season_misfit(S.ensemble_params(2, seed=1), None, ([1], [5], [5], [0.1]))
    assert e.value.code == _lib.ERR_ARG and "nx > 96" in str(e.value)
    eng.close()


def test_op_dynamics(cuda):
    from nesosim_b200 import engine as E
    rng = np.random.default_rng(1)
    ny, nx = 19, 45
    h = np.abs(rng.standard_normal((2, ny, nx))) * 0.2
    h[:, 4:7, 10:13] = np.nan
    d = 0.1 * rng.standard_normal((2, ny, nx))
    d[:, 12, :] = np.nan
    d[0, 2, 2] = np.inf
    for dx in (100000, 25000, 12345.678):
        adv, div = E.op_dynamics(d, h, dx)
        radv, rdiv = O.calc_dynamics(d, h, dx, O.Params())
        assert_parity(adv, radv, "adv dx=%r" % dx)
        assert_parity(div, rdiv, "div dx=%r" % dx)
        assert np.isfinite(adv).all() and np.isfinite(div).all()


def test_op_wind_terms(cuda):
    from nesosim_b200 import engine as E
    rng = np.random.default_rng(2)
    n = 4096
    h0 = np.abs(rng.standard_normal(n)) * 0.3
    W = rng.gamma(4.0, 1.5, n)
    C = rng.random(n)
    W[:8] = 5.0                          # exactly the threshold: not packed
    W[8:12] = np.nan                     # 0*NaN stays NaN (NESOSIM.py:69-71)
    h0[12:16] = np.nan
    W[16] = np.inf
    h0[17] = 0.0
    W[17] = np.inf                       # 0*inf = NaN
    p = O.Params(windPackFactor=5.8e-7, windPackThresh=5., leadLossFactor=1.45e-7, atmLossFactor=2.2e-8)
    lead, atm, wpl, wpg, wpn = E.op_wind_terms(h0, W, C, [MULTISEASON])
    with np.errstate(all="ignore"):
        rl, rg, rn = O.wind_packing(W, h0, p)
        assert_parity(lead, O.lead_loss(h0, W, C, p), "lead")
        assert_parity(atm, O.atm_loss(h0, W, p), "atm")
    assert_parity(wpl, rl, "wpl")
    assert_parity(wpg, rg, "wpg")
    assert_parity(wpn, rn, "wpn")
    assert np.isnan(lead[8]) and lead[0] == 0.0


def test_op_fills_and_density(cuda):
    from nesosim_b200 import engine as E
    rng = np.random.default_rng(3)
    n = 2000
    a = rng.standard_normal(n)
    a[::7] = np.nan
    a[::11] = np.inf
    a[::13] = -np.inf
    a[5] = -0.0
    mask = rng.integers(0, 13, n).astype(np.uint8)
    z = a.copy()
    O.fill_mask_nan_zero(z)
    assert_parity(E.op_fill_zero(a), z, "fill zero")
    for neg in (True, False):
        r = a.copy()
        O.fill_nan_no_negative(r, mask, negative_to_zero=neg)
        assert_parity(E.op_fill_nan_no_negative(a, mask, neg), r, "fill nan %s" % neg)
    h = np.abs(rng.standard_normal((2, n))) * 0.05
    h[0, :50] = 0.0
    h[1, :25] = 0.0                      # 0/0
    h[0, 60:70] = 0.01
    h[1, 60:70] = 0.01                   # sum straddles minSnowD
    h[0, 70] = 0.02
    h[1, 70] = 0.0
    h[:, 80:90] = np.nan
    assert_parity(E.op_density(h, mask), O.density_calc(h, None, mask, O.Params()), "density")

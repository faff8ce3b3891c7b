"""GPU parity at BASELINE's full grid sizes against digests the REFERENCE's own loop produced
(tests/golden/season_25km_digest.npz: 357x357, the whole 260-day season; tests/golden/steps_5km_digest.npz: 1785x1785,
three steps; both made by ``python tests/golden/make_golden.py large`` from /root/reference executed verbatim).
Every full array is compared by sha256 (NaNs canonicalised), the last slot additionally as sampled rows, so a failure
says where.  The general path -- both CTA shapes of the day kernel, with and without the land-tile shortcut -- and the
row-strip decomposition must reproduce the digests."""
import numpy as np
import pytest

from golden_util import OUT_NAMES, assert_identical, canon_sha, load
from nesosim_b200 import domain, synthetic as S

pytestmark = pytest.mark.gpu


def _inputs(g):
    dx, T, seed = int(g["dx"]), int(g["T"]), int(g["seed"])
    mask = S.region_mask(dx=dx)
    if canon_sha(mask.astype(np.float64)) != str(g["mask_sha"]):
        pytest.skip("bundled mask differs from the fixture's")
    F = S.make_season(mask, T, seed=seed)
    ic = S.make_ic(mask, seed=seed)
    for k in ("precip", "conc", "wind", "drift"):
        if canon_sha(F[k]) != str(g["in_sha__" + k]):
            pytest.skip("synthetic generator output differs from the fixture (numpy/scipy version)")
    return mask, dx, T, F, ic


def _check(out, g, what):
    rows = g["sample_rows"]
    for name in OUT_NAMES:
        assert_identical(out[name][-1][..., rows, :], g["rows_last__" + name], "%s[-1] sampled rows (%s)" % (name, what))
        assert canon_sha(out[name][-1]) == str(g["sha_last__" + name]), "%s[-1] (%s)" % (name, what)
        assert canon_sha(out[name]) == str(g["sha__" + name]), "%s (%s)" % (name, what)


_CACHE = {}


def _case(fname):
    if fname not in _CACHE:
        g = load(fname)
        _CACHE.clear()                      # one large case in memory at a time
        _CACHE[fname] = (g,) + _inputs(g)
    return _CACHE[fname]


@pytest.mark.parametrize("threads,shortcut", [("256", "1"), ("512", "1"), ("512", "0")])
@pytest.mark.parametrize("fname", ["season_25km_digest.npz", "steps_5km_digest.npz"])
def test_general_path_reproduces_the_reference_digest(cuda, monkeypatch, fname, threads, shortcut):
    from nesosim_b200.engine import SnowBudgetEngine
    g, mask, dx, T, F, ic = _case(fname)
    monkeypatch.setenv("NESOSIM_DAY_THREADS", threads)
    monkeypatch.setenv("NESOSIM_LAND_SHORTCUT", shortcut)
    eng = SnowBudgetEngine(mask, T, dx, n_members=1, atmlossInc=int(g["atmlossInc"]))
    eng.set_path("general")
    eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    out = {k: v[0].cpu().numpy() for k, v in eng.run_season([g["params"]], ic).items()}
    eng.close()
    _check(out, g, "%s threads, land shortcut %s" % (threads, shortcut))


@pytest.mark.parametrize("n_strips", [2, 3, 4])
def test_5km_row_strips_reproduce_the_reference_digest(cuda, n_strips):
    """The 5 km grid cut into row strips (all on this GPU, each with its own stream, ghost rows through the peer-memory
    mailboxes exactly as across GPUs): the assembled owned rows equal the reference's three steps."""
    g, mask, dx, T, F, ic = _case("steps_5km_digest.npz")
    out = domain.run_decomposed_season_peer_one_process(mask, T, dx, F, list(g["params"]), ic, n_strips,
                                                        whole_season_per_strip=True, atmlossInc=int(g["atmlossInc"]))
    _check(out, g, "%d strips" % n_strips)

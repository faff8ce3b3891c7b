"""Smallest run through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
season-resident kernel (plain, forcing sets, misfit mode), general day kernel (both CTA shapes, land shortcut),
row strips over the mailboxes (two strips, one stream each), smooth / final products / per-function kernels.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nesosim_b200 import domain, engine as E, synthetic as S
from nesosim_b200.engine import SnowBudgetEngine
mask = S.region_mask(dx=100000)
T, M = 5, 3
F = S.make_season(mask, T, seed=2)
ic = S.make_ic(mask, seed=2)
params = S.ensemble_params(M, seed=2)
res = {}
for path in ("ensemble", "general"):
    eng = SnowBudgetEngine(mask, T, 100000, n_members=M, atmlossInc=1)
    eng.set_path(path)
    eng.set_forcing(F["precip"], F["conc"], F["wind"], F["drift"])
    out = eng.run_season(params, ic)
    torch.cuda.synchronize()
    res[path] = {k: v.cpu().numpy() for k, v in out.items()}
    if path == "ensemble":
        rng = np.random.default_rng(2)
        obs = (rng.integers(0, T, 300), rng.integers(0, 90, 300), rng.integers(0, 90, 300), rng.random(300))
        mis, used = eng.run_season_misfit(params, ic, obs)
        print("misfit", mis.cpu().numpy(), used.cpu().numpy())
    eng.close()
ok = all(np.array_equal(res["ensemble"][k], res["general"][k], equal_nan=True) for k in res["general"])
print("paths identical:", ok)
# forcing sets (season kernel, SETS instantiation)
eng = SnowBudgetEngine(mask, T, 100000, n_members=2, atmlossInc=1)
stack = {k: np.stack([F[k], F[k][::-1].copy()]) for k in ("precip", "conc", "wind", "drift")}
eng.set_forcing_sets(stack["precip"], stack["conc"], stack["wind"], stack["drift"], [0, 1], [T, T - 1])
eng.run_season(params[:2], ic)
torch.cuda.synchronize()
eng.close()
# 512-thread day kernel with the land shortcut on a grid with all-land tiles, then two strips through the mailboxes
os.environ["NESOSIM_DAY_THREADS"] = "512"
os.environ["NESOSIM_LAND_SHORTCUT"] = "1"
m2 = S.region_mask(shape=(70, 100), kind="disc")
F2 = S.make_season(m2, T, seed=3)
ic2 = S.make_ic(m2, seed=3)
one = SnowBudgetEngine(m2, T, 50000, n_members=1, atmlossInc=1)
one.set_path("general")
one.set_forcing(F2["precip"], F2["conc"], F2["wind"], F2["drift"])
ref = {k: v[0].cpu().numpy() for k, v in one.run_season([params[0]], ic2).items()}
one.close()
got = domain.run_decomposed_season_peer_one_process(m2, T, 50000, F2, list(params[0]), ic2, 2, whole_season_per_strip=True, atmlossInc=1)
print("strips identical:", all(np.array_equal(got[k], ref[k], equal_nan=True) for k in ref))
# small kernels
a = np.random.default_rng(1).standard_normal((20, 30))
a[3, 4] = np.nan
E.smooth(a)
E.final_products(res["general"]["snowDepths"][0], res["general"]["density"][0], F["conc"], F["precip"], F["wind"])
torch.cuda.synchronize()
print("done")

"""Smallest season through both kernel paths (for compute-sanitizer --tool memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nesosim_b200 import synthetic as S
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
    eng.close()
ok = all(np.array_equal(res["ensemble"][k], res["general"][k], equal_nan=True) for k in res["general"])
print("paths identical:", ok)
